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


def decode_event_batch(payload: bytes) -> list:
    """processEvent's msgpack decoding (kvevents/pool.go:177-244, events.go:38-96).

    EventBatch = [ts, [event...], dp_rank?]; each event is an array-tagged
    union [tag, fields...]; trailing omitempty fields may be absent.  Malformed
    events are skipped, a malformed batch yields [] (poison pill dropped).
    """
    import msgpack
    try:
        batch = msgpack.unpackb(payload, raw=False, strict_map_key=False)
    except Exception:
        return []
    if not isinstance(batch, (list, tuple)) or len(batch) < 2 or not isinstance(batch[1], (list, tuple)):
        return []
    out = []
    for ev in batch[1]:
        if not isinstance(ev, (list, tuple)) or len(ev) < 1 or not isinstance(ev[0], str):
            continue
        tag, f = ev[0], list(ev[1:])
        try:
            if tag == "BlockStored":
                if len(f) < 4:
                    continue
                out.append(BlockStored(list(f[0] or []), f[1], [int(t) for t in (f[2] or [])], int(f[3] or 0),
                                       f[4] if len(f) > 4 else None, f[5] if len(f) > 5 else None))
            elif tag == "BlockRemoved":
                if len(f) < 1:
                    continue
                out.append(BlockRemoved(list(f[0] or []), f[1] if len(f) > 1 else None))
            elif tag == "AllBlocksCleared":
                out.append(AllBlocksCleared())
        except Exception:
            continue
    return out


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
                tier = ev.medium.lower() if ev.medium is not None else DEFAULT_DEVICE_TIER   # :258-261
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
                tier = ev.medium.lower() if ev.medium is not None else DEFAULT_DEVICE_TIER
                entries = [PodEntry(pod, tier)]
                for raw in ev.block_hashes:                                  # :317-330
                    try:
                        h = get_hash_as_uint64(raw)
                    except (TypeError, ValueError):
                        continue
                    self.index.evict(Key(model, h), entries)
            # AllBlocksCleared: no-op (:332-333)
