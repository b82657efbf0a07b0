"""CPU oracle for the Score() read path and the KVEvents write path.

TEST INFRASTRUCTURE ONLY.  Nothing outside ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import this
module; the product path (``llm-d-kv-cache-manager_b200``) never does and fails
loudly when its CUDA library is missing.

This is a plain-Python restatement of the reference's Go algorithm
(llm-d/llm-d-kv-cache-manager @ a378f5b3).  Every function cites the Go source
(path:line relative to the reference root) it follows.

PARITY STATUS
  * hash values ("request keys"): **parity unpinned** -- the reference's only
    golden-hash test is skipped (tests/integration/prompt_to_block_test.go:59)
    and Go is not installed here, so the FNV-64a/canonical-CBOR chain is pinned
    against (a) an independent CBOR implementation (python ``cbor2``,
    canonical=True; see tests/golden/make_golden.py) and (b) the published
    FNV-1a offset/prime and RFC 7049 canonical-form rules.
  * index semantics and scores: pinned by the reference's own unit tests
    (pkg/kvcache/kvblock/index_test.go:66-211, in_memory_test.go:44-116,
    pkg/kvcache/kvblock_scorer_test.go:34-99), re-expressed in
    tests/test_oracle_reference_cases.py.

Third-party behaviour restated (sources not vendored in the reference):
  * hashicorp/golang-lru/v2 v2.0.7 (go.mod:13)  -> class ``LRU``
  * fxamacker/cbor/v2 v2.7.0 CanonicalEncOptions (go.mod:11) -> ``cbor_*``
  * stdlib hash/fnv New64a / New32a -> ``fnv64a`` / ``fnv32a``
  * vmihailenco/msgpack/v5 v5.4.1 array-encoded structs (go.mod:19)
    -> ``decode_event_batch`` (python ``msgpack`` does the byte parsing)
"""
from __future__ import annotations

from collections import OrderedDict
from dataclasses import dataclass
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

MASK64 = (1 << 64) - 1
FNV64_OFFSET = 0xCBF29CE484222325
FNV64_PRIME = 0x100000001B3
FNV32_OFFSET = 0x811C9DC5
FNV32_PRIME = 0x01000193

DEFAULT_BLOCK_SIZE = 16          # token_processor.go:31
DEFAULT_INDEX_SIZE = 10 ** 8     # in_memory.go:33
DEFAULT_PODS_PER_KEY = 10        # in_memory.go:34
DEFAULT_DEVICE_TIER = "gpu"      # kvevents/pool.go (DefaultDeviceTier)
DEFAULT_TIER_WEIGHTS = {"gpu": 1.0, "cpu": 0.8}   # backend.go:26-31


# --------------------------------------------------------------------------
# hashing: stdlib hash/fnv + fxamacker/cbor canonical encoding
# --------------------------------------------------------------------------
def fnv64a(data: bytes, h: int = FNV64_OFFSET) -> int:
    """hash/fnv New64a: h = (h ^ b) * prime mod 2^64 per byte."""
    for b in data:
        h = ((h ^ b) * FNV64_PRIME) & MASK64
    return h


def fnv32a(data: bytes) -> int:
    """hash/fnv New32a (used by kvevents/pool.go:135-142 for queue sharding)."""
    h = FNV32_OFFSET
    for b in data:
        h = ((h ^ b) * FNV32_PRIME) & 0xFFFFFFFF
    return h


def cbor_head(major: int, v: int) -> bytes:
    """RFC 7049 shortest-form head (what CanonicalEncOptions emits)."""
    m = major << 5
    if v < 24:
        return bytes([m | v])
    if v < 1 << 8:
        return bytes([m | 24, v])
    if v < 1 << 16:
        return bytes([m | 25]) + v.to_bytes(2, "big")
    if v < 1 << 32:
        return bytes([m | 26]) + v.to_bytes(4, "big")
    return bytes([m | 27]) + v.to_bytes(8, "big")


def cbor_block_payload(parent: int, tokens: Sequence[int]) -> bytes:
    """Canonical CBOR of ``[]interface{}{parent uint64, tokens []uint32, nil}``.

    token_processor.go:94-103.  ``[]uint32`` is an *array of uints* (major 4),
    the nil interface is 0xf6, the outer slice a 3-element array (0x83).
    """
    out = bytearray(b"\x83")
    out += cbor_head(0, parent)
    out += cbor_head(4, len(tokens))
    for t in tokens:
        out += cbor_head(0, t)
    out += b"\xf6"
    return bytes(out)


def block_hash(parent: int, tokens: Sequence[int]) -> int:
    """ChunkedTokenDatabase.hash  (token_processor.go:94-112)."""
    return fnv64a(cbor_block_payload(parent, tokens))


@dataclass(frozen=True)
class Key:
    """kvblock.Key (index.go:138-141): model name is identity, not hashed."""
    model: str
    chunk_hash: int


@dataclass(frozen=True)
class PodEntry:
    """kvblock.PodEntry (index.go:149-154): identity is the (pod, tier) pair."""
    pod: str
    tier: str


class ChunkedTokenDatabase:
    """token_processor.go:63-162."""

    def __init__(self, block_size: int = DEFAULT_BLOCK_SIZE, hash_seed: str = ""):
        self.block_size = block_size
        self.hash_seed = hash_seed

    def init_hash(self) -> int:
        """getInitHash (token_processor.go:81-90): FNV64a of the seed string."""
        return fnv64a(self.hash_seed.encode())

    def chunk_tokens(self, tokens: Sequence[int]) -> List[Sequence[int]]:
        """chunkTokens (:126-138): full blocks only, the tail is dropped."""
        bs = self.block_size
        return [tokens[i:i + bs] for i in range(0, len(tokens) - bs + 1, bs)]

    def prefix_hashes(self, parent: int, chunks) -> List[int]:
        """prefixHashes (:115-123): serial chain."""
        out = []
        for c in chunks:
            parent = block_hash(parent, c)
            out.append(parent)
        return out

    def tokens_to_kv_block_keys(self, parent_key: Optional[Key], tokens: Sequence[int],
                                model: str) -> List[Key]:
        """TokensToKVBlockKeys (:141-162); returns [] where Go returns nil."""
        parent = parent_key.chunk_hash if parent_key is not None else self.init_hash()
        chunks = self.chunk_tokens(tokens)
        if not chunks:
            return []
        return [Key(model, h) for h in self.prefix_hashes(parent, chunks)]


# --------------------------------------------------------------------------
# hashicorp/golang-lru/v2 v2.0.7  (simplelru + Cache wrapper)
# --------------------------------------------------------------------------
class LRU:
    """golang-lru Cache semantics used by in_memory.go.

    ``OrderedDict`` order is oldest -> newest, which is what ``Keys()`` returns.
    """

    def __init__(self, size: int):
        if size <= 0:
            raise ValueError("must provide a positive size")   # lru.New error
        self.size = size
        self.d: "OrderedDict" = OrderedDict()

    def add(self, k, v) -> bool:
        """Add: update+refresh an existing key, else insert; evict oldest if over."""
        if k in self.d:
            self.d[k] = v
            self.d.move_to_end(k)
            return False
        self.d[k] = v
        if len(self.d) > self.size:
            self.d.popitem(last=False)
            return True
        return False

    def get(self, k):
        """Get: refreshes recency on hit."""
        if k in self.d:
            self.d.move_to_end(k)
            return self.d[k], True
        return None, False

    def contains_or_add(self, k, v) -> Tuple[bool, bool]:
        """ContainsOrAdd: no recency refresh on hit."""
        if k in self.d:
            return True, False
        return False, self.add(k, v)

    def remove(self, k) -> bool:
        return self.d.pop(k, _MISSING) is not _MISSING

    def keys(self) -> list:
        return list(self.d.keys())

    def __len__(self):
        return len(self.d)


_MISSING = object()


# --------------------------------------------------------------------------
# kvblock.InMemoryIndex
# --------------------------------------------------------------------------
class IndexError_(Exception):
    """Stands in for the Go ``error`` values returned by the index."""


class InMemoryIndex:
    """in_memory.go:54-270 (single-threaded restatement; the mutex-protected
    double-checked paths collapse to their uncontended branch)."""

    def __init__(self, size: int = DEFAULT_INDEX_SIZE, pod_cache_size: int = DEFAULT_PODS_PER_KEY):
        self.data = LRU(size)                    # requestKey -> LRU[PodEntry]
        self.engine_to_request = LRU(size)       # engineKey  -> requestKey
        self.pod_cache_size = pod_cache_size

    def lookup(self, request_keys: Sequence[Key], pod_filter: Iterable[str] = ()) -> Dict[Key, List[PodEntry]]:
        """Lookup (in_memory.go:105-146)."""
        if len(request_keys) == 0:
            raise IndexError_("no requestKeys provided for lookup")
        filt = set(pod_filter or ())
        out: Dict[Key, List[PodEntry]] = {}
        for k in request_keys:
            pods, found = self.data.get(k)                   # :118 refreshes recency
            if found:
                if pods is None or len(pods) == 0:           # :119-122 cut
                    return out
                if not filt:
                    out[k] = pods.keys()                     # :126-128 oldest->newest
                else:
                    for e in pods.keys():                    # :130-135
                        if e.pod in filt:
                            out.setdefault(k, []).append(e)
            # not found: continue (no cut)  :137-139
        return out

    def add(self, engine_keys: Sequence[Key], request_keys: Sequence[Key], entries: Sequence[PodEntry]) -> None:
        """Add (in_memory.go:149-209)."""
        if len(engine_keys) == 0 or len(request_keys) == 0 or len(entries) == 0:
            raise IndexError_("no keys or entries provided for adding to index")
        if len(engine_keys) != len(request_keys):
            raise IndexError_("mismatch between engine keys and request keys length")
        for ek, rk in zip(engine_keys, request_keys):
            self.engine_to_request.add(ek, rk)               # :163
            pc, found = self.data.get(rk)                    # :170
            if not found:
                pc = LRU(self.pod_cache_size)                # :174 (size<=0 -> error)
                self.data.contains_or_add(rk, pc)            # :186
            for e in entries:                                # :199-203
                pc.add(e, None)

    def evict(self, engine_key: Key, entries: Sequence[PodEntry]) -> None:
        """Evict (in_memory.go:212-260)."""
        if len(entries) == 0:
            raise IndexError_("no entries provided for eviction from index")
        rk, found = self.engine_to_request.get(engine_key)   # :219
        if not found:
            return
        pc, found = self.data.get(rk)                        # :225
        if not found or pc is None:
            self.engine_to_request.remove(engine_key)        # :228
            return
        for e in entries:
            pc.remove(e)                                     # :234
        if len(pc) == 0:                                     # :243-256
            cur, still = self.data.get(rk)
            if still and cur is not None and len(cur) == 0:
                self.data.remove(rk)
                self.engine_to_request.remove(engine_key)

    def get_request_key(self, engine_key: Key) -> Key:
        """GetRequestKey (in_memory.go:264-270): miss is an error."""
        rk, found = self.engine_to_request.get(engine_key)
        if not found:
            raise IndexError_("engine key not found: %s@%d" % (engine_key.model, engine_key.chunk_hash))
        return rk


# --------------------------------------------------------------------------
# kvcache.LongestPrefixScorer
# --------------------------------------------------------------------------
def get_max_weight(entries: Sequence[PodEntry], pod: str, weights: Optional[Dict[str, float]]) -> float:
    """getMaxWeight (kvblock_scorer.go:89-105). Starts at 0.0, unknown tier -> 1.0."""
    mx = 0.0
    for e in entries:
        if e.pod == pod:
            w = 1.0
            if weights is not None and e.tier in weights:
                w = weights[e.tier]
            if w > mx:
                mx = w
    return mx


class LongestPrefixScorer:
    """kvblock_scorer.go:77-151."""

    def __init__(self, weights: Optional[Dict[str, float]] = None):
        self.weights = weights

    def score(self, keys: Sequence[Key], key_to_pods: Dict[Key, List[PodEntry]]) -> Dict[str, float]:
        scores: Dict[str, float] = {}
        if len(keys) == 0:
            return scores
        first = key_to_pods.get(keys[0], [])
        active = {e.pod for e in first}
        for p in active:
            scores[p] = get_max_weight(first, p, self.weights)       # :126-128
        for i in range(1, len(keys)):
            if not active:                                           # :131-133
                break
            cur = key_to_pods.get(keys[i], [])
            active = active & {e.pod for e in cur}                   # :142
            for p in active:
                scores[p] = scores[p] + get_max_weight(cur, p, self.weights)   # :143-146 in-order f64 add
        return scores


# --------------------------------------------------------------------------
# tokenization prefix store (pkg/tokenization/prefixstore/lru_store.go): the cache in front of the tokenizer
# --------------------------------------------------------------------------
_XP1, _XP2, _XP3 = 0x9E3779B185EBCA87, 0xC2B2AE3D27D4EB4F, 0x165667B19E3779F9
_XP4, _XP5 = 0x85EBCA77C2B2AE63, 0x27D4EB2F165667C5


def _rotl64(x: int, r: int) -> int:
    return ((x << r) | (x >> (64 - r))) & MASK64


def _xx_round(acc: int, inp: int) -> int:
    return (_rotl64((acc + inp * _XP2) & MASK64, 31) * _XP1) & MASK64


def _xx_merge(h: int, v: int) -> int:
    return (((h ^ _xx_round(0, v)) * _XP1) + _XP4) & MASK64


def xxhash64(data: bytes, seed: int = 0) -> int:
    """XXH64 (github.com/cespare/xxhash/v2 v2.3.0 in the reference's go.mod; algorithm: the published xxHash spec)."""
    n, i = len(data), 0
    if n >= 32:
        v1, v2, v3, v4 = (seed + _XP1 + _XP2) & MASK64, (seed + _XP2) & MASK64, seed, (seed - _XP1) & MASK64
        while i + 32 <= n:
            v1 = _xx_round(v1, int.from_bytes(data[i:i + 8], "little")); v2 = _xx_round(v2, int.from_bytes(data[i + 8:i + 16], "little"))
            v3 = _xx_round(v3, int.from_bytes(data[i + 16:i + 24], "little")); v4 = _xx_round(v4, int.from_bytes(data[i + 24:i + 32], "little"))
            i += 32
        h = (_rotl64(v1, 1) + _rotl64(v2, 7) + _rotl64(v3, 12) + _rotl64(v4, 18)) & MASK64
        for v in (v1, v2, v3, v4):
            h = _xx_merge(h, v)
    else:
        h = (seed + _XP5) & MASK64
    h = (h + n) & MASK64
    while i + 8 <= n:
        h = (_rotl64(h ^ _xx_round(0, int.from_bytes(data[i:i + 8], "little")), 27) * _XP1 + _XP4) & MASK64
        i += 8
    if i + 4 <= n:
        h = (_rotl64(h ^ (int.from_bytes(data[i:i + 4], "little") * _XP1 & MASK64), 23) * _XP2 + _XP3) & MASK64
        i += 4
    while i < n:
        h = (_rotl64(h ^ (data[i] * _XP5 & MASK64), 11) * _XP1) & MASK64
        i += 1
    h ^= h >> 33; h = (h * _XP2) & MASK64
    h ^= h >> 29; h = (h * _XP3) & MASK64
    h ^= h >> 32
    return h


PREFIX_STORE_BLOCK_SIZE = 256        # lru_store.go:29 (bytes of text per block, despite the field's comment)
PREFIX_STORE_CACHE_SIZE = 500000     # lru_store.go:31


class LRUTokenStore:
    """prefixstore.LRUTokenStore (lru_store.go:54-190): xxhash64-chained text blocks -> the tokens that end inside them."""

    def __init__(self, cache_size: int = PREFIX_STORE_CACHE_SIZE, block_size: int = PREFIX_STORE_BLOCK_SIZE):
        self.block_size = block_size
        self.cache = LRU(cache_size)

    def _hash(self, prev: int, chunk: bytes) -> int:
        return xxhash64(prev.to_bytes(8, "little") + chunk)              # :111-118 binary.Write(LE, previousHash) then the chunk

    def add_tokenization(self, prompt: bytes, tokens: Sequence[int], offsets: Sequence[Tuple[int, int]]) -> None:
        """AddTokenization (:89-141).  prompt is the UTF-8 byte string; offsets are byte offsets [low, high)."""
        if len(prompt) == 0 or len(tokens) == 0:
            return
        it, prev = 0, 0
        for start in range(0, len(prompt), self.block_size):
            end = start + self.block_size
            if end > len(prompt):
                break                                                    # no partial blocks
            prev = self._hash(prev, prompt[start:end])
            blk = []
            while it < len(tokens) and offsets[it][1] <= end:            # :128-135 a token belongs to the block its END falls in
                blk.append(tokens[it]); it += 1
            self.cache.add(prev, blk)

    def find_longest_contained_tokens(self, prompt: bytes) -> Tuple[List[int], float]:
        """FindLongestContainedTokens (:143-190)."""
        out: List[int] = []
        prev, ratio = 0, 0.0
        for i in range(0, len(prompt), self.block_size):
            end = i + self.block_size
            if end > len(prompt):
                break
            prev = self._hash(prev, prompt[i:end])
            blk, ok = self.cache.get(prev)                               # refreshes recency
            if not ok:
                break                                                    # early stop
            out.extend(blk)
            ratio = end / len(prompt)
        return out, ratio


def tokenize_with_prefix_store(store: LRUTokenStore, prompt: bytes, encode, min_overlap: float = 0.8) -> List[int]:
    """tokenization.Pool.processTask after chat templating (pool.go:209-225): reuse the contained tokens when the cached
    prefix covers at least min_overlap of the prompt, else tokenize and remember.  encode(prompt) -> (tokens, offsets)."""
    toks, ratio = store.find_longest_contained_tokens(prompt)
    if ratio < min_overlap:
        tokens, offsets = encode(prompt)
        store.add_tokenization(prompt, tokens, offsets)
        return list(tokens)
    return toks


# --------------------------------------------------------------------------
# metrics: InstrumentedIndex (kvblock/instrumented_index.go:35-92) over the collectors of metrics/collector.go:28-59
# --------------------------------------------------------------------------
LATENCY_BUCKETS = (0.005, 0.01, 0.025, 0.05, 0.1, 0.25, 0.5, 1.0, 2.5, 5.0, 10.0)    # prometheus.DefBuckets


@dataclass
class IndexMetrics:
    admissions_total: int = 0            # kvcache_index_admissions_total
    evictions_total: int = 0             # kvcache_index_evictions_total
    lookup_requests_total: int = 0       # kvcache_index_lookup_requests_total
    max_pod_hit_count_total: int = 0     # kvcache_index_max_pod_hit_count_total
    lookup_hits_total: int = 0           # kvcache_index_lookup_hits_total
    lookup_latency_count: int = 0        # kvcache_index_lookup_latency_seconds (count only: values are wall clock)


class InstrumentedIndex:
    """NewInstrumentedIndex (instrumented_index.go:30-92): counts around the wrapped index."""

    def __init__(self, nxt: "InMemoryIndex", metrics: Optional[IndexMetrics] = None):
        self.next = nxt
        self.metrics = metrics if metrics is not None else IndexMetrics()
        self.data = nxt.data
        self.engine_to_request = nxt.engine_to_request

    def add(self, engine_keys, request_keys, entries) -> None:
        try:
            self.next.add(engine_keys, request_keys, entries)
        finally:
            self.metrics.admissions_total += len(request_keys)          # :35-39 counted whatever Add returned

    def evict(self, engine_key, entries) -> None:
        try:
            self.next.evict(engine_key, entries)
        finally:
            self.metrics.evictions_total += len(entries)                # :41-45

    def get_request_key(self, engine_key):
        return self.next.get_request_key(engine_key)

    def lookup(self, request_keys, pod_filter=()):
        self.metrics.lookup_latency_count += 1                          # :52-53 timer observed on every return
        self.metrics.lookup_requests_total += 1                         # :55
        pods = self.next.lookup(request_keys, pod_filter)               # an error returns before the hit metrics (:57-60)
        count: Dict[str, int] = {}
        for entries in pods.values():                                   # recordHitMetrics :71-92
            for e in entries:
                count[e.pod] = count.get(e.pod, 0) + 1
        mx = max(count.values()) if count else 0
        self.metrics.max_pod_hit_count_total += mx
        self.metrics.lookup_hits_total += mx
        return pods


# --------------------------------------------------------------------------
# kvcache.Indexer (steps 2-4 of GetPodScores; tokenisation is out of path)
# --------------------------------------------------------------------------
class Indexer:
    """indexer.go:75-166, starting from token ids."""

    def __init__(self, block_size: int = DEFAULT_BLOCK_SIZE, hash_seed: str = "",
                 size: int = DEFAULT_INDEX_SIZE, pod_cache_size: int = DEFAULT_PODS_PER_KEY,
                 weights: Optional[Dict[str, float]] = None):
        self.tokens_processor = ChunkedTokenDatabase(block_size, hash_seed)
        self.index = InMemoryIndex(size, pod_cache_size)
        self.scorer = LongestPrefixScorer(dict(DEFAULT_TIER_WEIGHTS) if weights is None else weights)

    def get_pod_scores(self, tokens: Sequence[int], model: str, pods: Sequence[str] = ()) -> Optional[Dict[str, float]]:
        """GetPodScores (indexer.go:132-166) after step 1.  None == Go's (nil, nil)."""
        keys = self.tokens_processor.tokens_to_kv_block_keys(None, tokens, model)
        if not keys:
            return None
        hits = self.index.lookup(keys, set(pods))
        return self.scorer.score(keys, hits)


# --------------------------------------------------------------------------
# kvevents: wire structs, digestEvents, AddTask sharding
# --------------------------------------------------------------------------
@dataclass
class BlockStored:           # events.go:49-58
    block_hashes: list       # uint64 | int64 | bytes each
    parent_block_hash: object
    token_ids: List[int]
    block_size: int = 0      # ignored by the indexer
    lora_id: Optional[int] = None
    medium: Optional[str] = None


@dataclass
class BlockRemoved:          # events.go:77-81
    block_hashes: list
    medium: Optional[str] = None


@dataclass
class AllBlocksCleared:      # events.go:93-96
    pass


def get_hash_as_uint64(h) -> int:
    """getHashAsUint64 (kvevents/pool.go:343-367). Raises on unsupported input."""
    if isinstance(h, bool):
        raise TypeError("unsupported hash type: bool")
    if isinstance(h, int):
        return h & MASK64                       # uint64 as is; int64 reinterpreted
    if isinstance(h, (bytes, bytearray)):
        if len(h) == 0:
            raise ValueError("hash byte slice is empty")
        if len(h) >= 8:
            return int.from_bytes(h[-8:], "big")
        return int.from_bytes(bytes(h), "big")  # left-zero-padded
    raise TypeError("unsupported hash type: %r" % type(h))


class TInt(int):
    """An integer together with the msgpack type code it arrived with.  vmihailenco/msgpack's DecodeInterface returns
    int8..int64 / uint8..uint64 according to the wire code, and getHashAsUint64 (kvevents/pool.go:343-367) only accepts
    the uint64 (0xcf) and int64 (0xd3) forms -- so the width matters for parity."""
    code = 0

    def __new__(cls, v, code):
        o = super().__new__(cls, v)
        o.code = code
        return o


class _MpError(Exception):
    pass


def _mp_read(buf: bytes, pos: int, depth: int = 0):
    """Minimal msgpack reader: returns (value, new_pos).  ints -> TInt, str/bin -> ("str"|"bin", bytes) tuples are NOT
    used; str -> str (bytes if not UTF-8), bin -> bytes, arrays -> list, maps -> dict, ext -> ("ext", bytes)."""
    if depth > 64 or pos >= len(buf):
        raise _MpError("truncated")
    b = buf[pos]; pos += 1

    def take(n):
        nonlocal pos
        if pos + n > len(buf):
            raise _MpError("truncated")
        v = buf[pos:pos + n]; pos += n
        return v

    if b <= 0x7F:
        return TInt(b, b), pos
    if b >= 0xE0:
        return TInt(b - 256, b), pos
    if 0xA0 <= b <= 0xBF:
        return _mp_str(take(b & 0x1F)), pos
    if 0x90 <= b <= 0x9F or b in (0xDC, 0xDD):
        n = b & 0x0F if b <= 0x9F else int.from_bytes(take(2 if b == 0xDC else 4), "big")
        out = []
        for _ in range(n):
            v, pos = _mp_read(buf, pos, depth + 1)
            out.append(v)
        return out, pos
    if 0x80 <= b <= 0x8F or b in (0xDE, 0xDF):
        n = b & 0x0F if b <= 0x8F else int.from_bytes(take(2 if b == 0xDE else 4), "big")
        out = {}
        for _ in range(n):
            k, pos = _mp_read(buf, pos, depth + 1)
            v, pos = _mp_read(buf, pos, depth + 1)
            out[repr(k)] = v
        return out, pos
    if b == 0xC0:
        return None, pos
    if b in (0xC2, 0xC3):
        return b == 0xC3, pos
    if b in (0xC4, 0xC5, 0xC6):
        n = int.from_bytes(take(1 << (b - 0xC4)), "big")
        return bytes(take(n)), pos
    if b in (0xC7, 0xC8, 0xC9):
        n = int.from_bytes(take(1 << (b - 0xC7)), "big"); take(1)
        return ("ext", bytes(take(n))), pos
    if b == 0xCA:
        import struct
        return struct.unpack(">f", take(4))[0], pos
    if b == 0xCB:
        import struct
        return struct.unpack(">d", take(8))[0], pos
    if 0xCC <= b <= 0xCF:
        return TInt(int.from_bytes(take(1 << (b - 0xCC)), "big"), b), pos
    if 0xD0 <= b <= 0xD3:
        return TInt(int.from_bytes(take(1 << (b - 0xD0)), "big", signed=True), b), pos
    if 0xD4 <= b <= 0xD8:
        take(1)
        return ("ext", bytes(take(1 << (b - 0xD4)))), pos
    if b in (0xD9, 0xDA, 0xDB):
        n = int.from_bytes(take(1 << (b - 0xD9)), "big")
        return _mp_str(take(n)), pos
    raise _MpError("invalid code 0x%02x" % b)


class _MpStr(str):
    """str that remembers its raw bytes (Go strings are byte strings; lower() must not choke on invalid UTF-8)."""
    raw = b""


def _mp_str(raw: bytes):
    o = _MpStr(raw.decode("utf-8", "surrogateescape"))
    o.raw = bytes(raw)
    return o


def typed_hash(v):
    """getHashAsUint64 on a DecodeInterface-typed value: only uint64 / int64 coded ints and byte slices pass."""
    if isinstance(v, TInt):
        if v.code in (0xCF, 0xD3):
            return int(v) & MASK64
        raise TypeError("unsupported hash type: int code 0x%02x" % v.code)
    if isinstance(v, (bytes, bytearray)) and not isinstance(v, str):
        return get_hash_as_uint64(bytes(v))
    raise TypeError("unsupported hash type: %r" % type(v))


def _is_int(v):
    return isinstance(v, TInt)


def decode_event_batch(payload: bytes) -> list:
    """processEvent's msgpack decoding (kvevents/pool.go:177-244, events.go:38-96) with vmihailenco/msgpack v5 typing.

    EventBatch = [ts, [event...], dp_rank?] decoded all-or-nothing (any error drops the message, pool.go:182-187);
    each event is an array-tagged union [tag, fields...]; array-encoded structs take fields positionally, missing
    trailing fields stay zero, extra ones are skipped; a field of the wrong type skips that event (pool.go:233-237).
    Hash lists are returned already filtered through getHashAsUint64's type rules (pool.go:270-277): the elements are
    plain ints (the uint64 value) so that digest_events sees exactly what the Go code would append."""
    try:
        batch, _ = _mp_read(bytes(payload), 0)
    except _MpError:
        return []
    if not isinstance(batch, list):
        return []
    n = len(batch)
    if n >= 1 and not (isinstance(batch[0], float) or _is_int(batch[0]) or batch[0] is None):
        return []
    events_raw = []
    if n >= 2:
        if isinstance(batch[1], list):
            events_raw = batch[1]
        elif batch[1] is not None:
            return []
    if n >= 3 and not (_is_int(batch[2]) or batch[2] is None):
        return []
    out = []
    for ev in events_raw:
        if not isinstance(ev, list) or len(ev) < 1:
            continue
        tag = ev[0]
        if isinstance(tag, (bytes, bytearray)) and not isinstance(tag, str):
            tag = bytes(tag).decode("utf-8", "surrogateescape")
        if not isinstance(tag, str):
            continue
        f = ev[1:]

        def medium_of(idx):
            if len(f) <= idx or f[idx] is None:
                return None, True
            m = f[idx]
            if isinstance(m, (bytes, bytearray)) and not isinstance(m, str):
                return bytes(m).decode("utf-8", "surrogateescape"), True
            if isinstance(m, str):
                return str(m), True
            return None, False

        def hashes_of(v):
            if v is None:
                return [], True
            if not isinstance(v, list):
                return None, False
            res = []
            for x in v:
                try:
                    res.append(typed_hash(x))
                except (TypeError, ValueError):
                    continue                       # pool.go:272-275: unsupported / empty hash is skipped
            return res, True

        if tag == "BlockStored":
            hs, ok = hashes_of(f[0]) if len(f) > 0 else ([], True)
            if not ok:
                continue
            parent = None
            parent_bad = False
            if len(f) > 1 and f[1] is not None:
                try:
                    parent = typed_hash(f[1])
                except (TypeError, ValueError):
                    parent_bad = True
            toks = []
            if len(f) > 2 and f[2] is not None:
                if not isinstance(f[2], list) or not all(_is_int(t) for t in f[2]):
                    continue
                toks = [int(t) & 0xFFFFFFFF for t in f[2]]
            if len(f) > 3 and not (_is_int(f[3]) or f[3] is None):
                continue
            if len(f) > 4 and not (_is_int(f[4]) or f[4] is None):
                continue
            med, ok = medium_of(5)
            if not ok:
                continue
            if parent_bad:
                continue                           # pool.go:283-287: the whole event is skipped
            out.append(BlockStored(hs, parent, toks, 0, None, med))
        elif tag == "BlockRemoved":
            hs, ok = hashes_of(f[0]) if len(f) > 0 else ([], True)
            if not ok:
                continue
            med, ok = medium_of(1)
            if not ok:
                continue
            out.append(BlockRemoved(hs, med))
        elif tag == "AllBlocksCleared":
            out.append(AllBlocksCleared())
    return out


def medium_tier(medium):
    """lower(Medium) or "gpu" (kvevents/pool.go:258-261).  Go's strings.ToLower folds Unicode letters (and rewrites invalid
    UTF-8); vLLM media are ASCII ("GPU", "CPU", ...), so only ASCII letters are folded here and in the C++ host mirror."""
    if medium is None:
        return DEFAULT_DEVICE_TIER
    return "".join(chr(ord(c) + 32) if "A" <= c <= "Z" else c for c in medium)


class EventsPool:
    """kvevents.Pool without ZMQ: AddTask sharding + digestEvents (pool.go:132-338)."""

    def __init__(self, index: InMemoryIndex, token_processor: ChunkedTokenDatabase, concurrency: int = 4):
        self.index = index
        self.tp = token_processor
        self.concurrency = concurrency

    def queue_index(self, pod: str) -> int:
        """AddTask (pool.go:132-144): FNV-32a(pod) % concurrency."""
        return fnv32a(pod.encode()) % self.concurrency

    def process_event(self, pod: str, model: str, payload: bytes) -> None:
        self.digest_events(pod, model, decode_event_batch(payload))

    def digest_events(self, pod: str, model: str, events: Sequence[object]) -> None:
        """digestEvents (pool.go:246-338)."""
        for ev in events:
            if isinstance(ev, BlockStored):
                tier = medium_tier(ev.medium)                                 # :258-261
                entries = [PodEntry(pod, tier)]
                engine_keys = []
                for raw in ev.block_hashes:                                  # :270-277 bad hashes skipped
                    try:
                        engine_keys.append(Key(model, get_hash_as_uint64(raw)))
                    except (TypeError, ValueError):
                        continue
                parent_rk = None
                if ev.parent_block_hash is not None:                         # :279-294
                    try:
                        ph = get_hash_as_uint64(ev.parent_block_hash)
                    except (TypeError, ValueError):
                        continue                                             # event skipped
                    try:
                        parent_rk = self.index.get_request_key(Key(model, ph))
                    except IndexError_:
                        parent_rk = None                                     # chain restarts at the seed
                request_keys = self.tp.tokens_to_kv_block_keys(parent_rk, ev.token_ids, model)   # :296
                if engine_keys:                                              # :299-305
                    try:
                        self.index.add(engine_keys, request_keys, entries)
                    except IndexError_:
                        continue                                             # event dropped
            elif isinstance(ev, BlockRemoved):
                tier = medium_tier(ev.medium)
                entries = [PodEntry(pod, tier)]
                for raw in ev.block_hashes:                                  # :317-330
                    try:
                        h = get_hash_as_uint64(raw)
                    except (TypeError, ValueError):
                        continue
                    self.index.evict(Key(model, h), entries)
            # AllBlocksCleared: no-op (:332-333)
