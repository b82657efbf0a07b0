// goprobe -- TEST INFRASTRUCTURE (oracle/): pins the hash KATs against real Go.
//
// The reference cannot be built in the build container (no Go toolchain, no module cache), so the block-key
// values in tests/golden/hash_kats.json are pinned against an independent Python implementation only
// ("parity unpinned", DESIGN.md).  bench.py probes for `go` at run time; if a toolchain AND the one third-party
// module the hash depends on (github.com/fxamacker/cbor/v2 v2.7.0, go.mod:11 of the reference) are on the
// box, this program re-computes every KAT with the reference's own call shape
// (pkg/kvcache/kvblock/token_processor.go:94-123: CanonicalEncOptions().EncMode(), Marshal of
// []interface{}{parent uint64, tokens []uint32, nil}, FNV-64a over the bytes, keys chained block by block) and
// reports whether the fixture agrees.  Input: hash_kats.json on stdin.  Output: one JSON object.
package main

import (
	"encoding/json"
	"fmt"
	"hash/fnv"
	"io"
	"os"

	"github.com/fxamacker/cbor/v2"
)

type kat struct {
	Name      string   `json:"name"`
	Seed      string   `json:"seed"`
	BlockSize int      `json:"block_size"`
	Parent    *uint64  `json:"parent"`
	Tokens    []uint32 `json:"tokens"`
	Keys      []uint64 `json:"keys"`
}

type fixture struct {
	Fnv64a map[string]uint64 `json:"fnv64a"`
	Cases  []kat             `json:"cases"`
}

func hashBlock(parent uint64, tokens []uint32) (uint64, error) {
	em, err := cbor.CanonicalEncOptions().EncMode()
	if err != nil {
		return 0, err
	}
	b, err := em.Marshal([]interface{}{parent, tokens, nil})
	if err != nil {
		return 0, err
	}
	h := fnv.New64a()
	_, _ = h.Write(b)
	return h.Sum64(), nil
}

func main() {
	raw, err := io.ReadAll(os.Stdin)
	if err != nil {
		fmt.Fprintln(os.Stderr, err)
		os.Exit(2)
	}
	var fx fixture
	if err := json.Unmarshal(raw, &fx); err != nil {
		fmt.Fprintln(os.Stderr, err)
		os.Exit(2)
	}
	mismatches := []string{}
	for s, want := range fx.Fnv64a {
		h := fnv.New64a()
		_, _ = h.Write([]byte(s))
		if h.Sum64() != want {
			mismatches = append(mismatches, "fnv64a:"+s)
		}
	}
	for _, c := range fx.Cases {
		h := fnv.New64a()
		_, _ = h.Write([]byte(c.Seed))
		parent := h.Sum64()
		if c.Parent != nil {
			parent = *c.Parent
		}
		n := 0
		if c.BlockSize > 0 {
			n = len(c.Tokens) / c.BlockSize
		}
		ok := n == len(c.Keys)
		for i := 0; i < n && ok; i++ {
			k, err := hashBlock(parent, c.Tokens[i*c.BlockSize:(i+1)*c.BlockSize])
			if err != nil || k != c.Keys[i] {
				ok = false
			}
			parent = k
		}
		if !ok {
			mismatches = append(mismatches, c.Name)
		}
	}
	out := map[string]interface{}{"cases": len(fx.Cases), "all_equal": len(mismatches) == 0, "mismatches": mismatches}
	_ = json.NewEncoder(os.Stdout).Encode(out)
}
