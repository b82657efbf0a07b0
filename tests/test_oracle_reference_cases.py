"""CPU: the reference's own unit tests for this path, re-expressed against both oracles.

  pkg/kvcache/kvblock_scorer_test.go:34-99         scorer known answers
  pkg/kvcache/kvblock/index_test.go:66-211         index conformance suite
  pkg/kvcache/kvblock/in_memory_test.go:44-116     key-LRU cap and pod-LRU cap
  tests/e2e/redis_mock/e2e_test.go:134-205         score == number of cached prefix blocks
"""
import numpy as np
import pytest

from oracle import kvoracle as ko
from oracle.kvoracle_c import COracle

M = "test-model"


def K(h):
    return ko.Key(M, h)


def E(p, t="gpu"):
    return ko.PodEntry(p, t)


# ---------------- scorer (kvblock_scorer_test.go) ----------------
def test_longest_prefix_scorer():                      # :34-67
    s = ko.LongestPrefixScorer({"gpu": 1.0, "cpu": 0.5})
    keys = [K(h) for h in (1001, 1002, 1003, 1004, 1005, 1006)]
    hit = {K(1001): [E("pod-a")], K(1002): [E("pod-a")], K(1003): [E("pod-a"), E("pod-a", "cpu")],
           K(1004): [E("pod-b", "cpu")], K(1005): [E("pod-b", "cpu")], K(1006): [E("pod-a")]}
    assert s.score(keys, hit) == {"pod-a": 3.0}


def test_longest_prefix_scorer_tiers():                # :69-99
    s = ko.LongestPrefixScorer({"gpu": 1.0, "cpu": 0.5})
    keys = [K(h) for h in (1001, 1002, 1003, 1004, 1005, 1006)]
    hit = {K(1001): [E("pod-a")], K(1002): [E("pod-a")], K(1003): [E("pod-a", "cpu")],
           K(1004): [E("pod-b", "cpu")], K(1005): [E("pod-b", "cpu")], K(1006): [E("pod-a")]}
    assert s.score(keys, hit) == {"pod-a": 2.5}


def test_scorer_unknown_tier_and_nil_weights():        # kvblock_scorer.go:93-98
    keys = [K(1), K(2)]
    hit = {K(1): [E("a", "disk")], K(2): [E("a", "disk")]}
    assert ko.LongestPrefixScorer({"gpu": 1.0}).score(keys, hit) == {"a": 2.0}
    assert ko.LongestPrefixScorer(None).score(keys, hit) == {"a": 2.0}
    assert ko.LongestPrefixScorer({"disk": -3.0}).score(keys, hit) == {"a": 0.0}   # max starts at 0.0 (:90)


# ---------------- index conformance (index_test.go) ----------------
class PyBackend:
    """string-level view of the Python oracle."""

    def __init__(self, size=10 ** 6, pods=10):
        self.ix = ko.InMemoryIndex(size, pods)

    def add(self, ek, rk, ents):
        self.ix.add([K(h) for h in ek], [K(h) for h in rk], [E(*e) for e in ents])

    def evict(self, ek, ents):
        self.ix.evict(K(ek), [E(*e) for e in ents])

    def lookup(self, rks, filt=()):
        return {k.chunk_hash: [(e.pod, e.tier) for e in v] for k, v in self.ix.lookup([K(h) for h in rks], set(filt)).items()}


class CBackend:
    """same view over the C++ oracle (ids interned here)."""

    def __init__(self, size=10 ** 6, pods=10):
        self.co = COracle(size=size, pod_cache_size=pods, max_pods=256)
        self.pods, self.tiers = {}, {"gpu": 0, "cpu": 1}

    def _pt(self, e):
        p = self.pods.setdefault(e[0], len(self.pods))
        t = self.tiers.setdefault(e[1] if len(e) > 1 else "gpu", len(self.tiers))
        return (p << 4) | t

    def add(self, ek, rk, ents):
        assert self.co.add(0, ek, rk, [self._pt(e) for e in ents]) == 0

    def evict(self, ek, ents):
        assert self.co.evict(0, ek, [self._pt(e) for e in ents]) == 0

    def lookup(self, rks, filt=()):
        inv_p = {v: k for k, v in self.pods.items()}
        inv_t = {v: k for k, v in self.tiers.items()}
        fm = None
        if filt:
            fm = np.zeros(self.co.filter_words, np.uint64)
            for f in filt:
                p = self.pods.setdefault(f, len(self.pods))
                fm[p // 64] |= np.uint64(1 << (p % 64))
        rc, pt, cnt = self.co.lookup(0, np.array(rks, np.uint64), fm)
        assert rc == 0
        return {h: [(inv_p[int(e) >> 4], inv_t[int(e) & 15]) for e in pt[i, :cnt[i]]] for i, h in enumerate(rks) if cnt[i]}


BACKENDS = [PyBackend, CBackend]


@pytest.mark.parametrize("B", BACKENDS)
def test_basic_add_and_lookup(B):                      # index_test.go:66-88
    b = B()
    b.add([55269488], [10633516], [("pod1", "gpu"), ("pod2", "gpu")])
    got = b.lookup([10633516])
    assert list(got) == [10633516] and sorted(got[10633516]) == [("pod1", "gpu"), ("pod2", "gpu")]


@pytest.mark.parametrize("B", BACKENDS)
def test_duplicate_pod_handling(B):                    # index_test.go:93-133
    b = B()
    b.add([91642125], [61519471], [("pod1", "gpu"), ("pod2", "gpu")])
    b.add([91642125], [61519471], [("pod1", "gpu"), ("pod2", "cpu"), ("pod3", "gpu")])
    got = b.lookup([61519471])[61519471]
    assert sorted(got) == sorted([("pod1", "gpu"), ("pod2", "gpu"), ("pod2", "cpu"), ("pod3", "gpu")])
    assert got == [("pod2", "gpu"), ("pod1", "gpu"), ("pod2", "cpu"), ("pod3", "gpu")]   # lru Keys(): oldest -> newest


@pytest.mark.parametrize("B", BACKENDS)
def test_filtered_lookup(B):                           # index_test.go:137-173
    b = B()
    b.add([93788608], [55204205], [("pod1", "gpu"), ("pod2", "gpu"), ("pod3", "gpu")])
    assert b.lookup([55204205], ["pod1"]) == {55204205: [("pod1", "gpu")]}
    assert sorted(b.lookup([55204205], ["pod1", "pod3"])[55204205]) == [("pod1", "gpu"), ("pod3", "gpu")]
    assert b.lookup([55204205], ["pod999"]) == {}


@pytest.mark.parametrize("B", BACKENDS)
def test_evict_basic(B):                               # index_test.go:177-211
    b = B()
    b.add([17434655], [59244875], [("pod1", "gpu"), ("pod2", "gpu"), ("pod3", "gpu")])
    b.evict(17434655, [("pod1", "gpu"), ("pod3", "cpu")])
    assert sorted(b.lookup([59244875])[59244875]) == [("pod2", "gpu"), ("pod3", "gpu")]


@pytest.mark.parametrize("B", BACKENDS)
def test_index_size_cap(B):                            # in_memory_test.go:44-83
    b = B(size=2, pods=1)
    b.add([72735753], [79215516], [("pod1", "gpu")])
    b.add([41341092], [12871930], [("pod2", "gpu")])
    b.add([34012886], [69914638], [("pod3", "cpu")])
    got = b.lookup([79215516, 12871930, 69914638])
    assert got == {12871930: [("pod2", "gpu")], 69914638: [("pod3", "cpu")]}   # missing first key does not cut


@pytest.mark.parametrize("B", BACKENDS)
def test_pod_cache_size_cap(B):                        # in_memory_test.go:85-116
    b = B(size=1, pods=2)
    b.add([28409753], [51374550], [("pod1", "gpu"), ("pod2", "gpu"), ("pod3", "cpu")])
    assert b.lookup([51374550]) == {51374550: [("pod2", "gpu"), ("pod3", "cpu")]}


def test_error_cases():
    ix = ko.InMemoryIndex()
    with pytest.raises(ko.IndexError_):
        ix.lookup([])                                  # in_memory.go:108-110
    with pytest.raises(ko.IndexError_):
        ix.add([], [], [E("p")])                       # :150-152
    with pytest.raises(ko.IndexError_):
        ix.add([K(1), K(2)], [K(3)], [E("p")])         # :153-155
    with pytest.raises(ko.IndexError_):
        ix.evict(K(1), [])                             # :213-215
    with pytest.raises(ko.IndexError_):
        ix.get_request_key(K(404))                     # :266-268
    ix.evict(K(404), [E("p")])                         # unknown engine key: silent no-op (:219-223)
    with pytest.raises(ValueError):
        ko.InMemoryIndex(size=0)                       # lru.New(size<=0)
    co = COracle()
    assert co.add(0, [], [], [1]) == -22 and co.add(0, [1, 2], [3], [1]) == -22 and co.evict(0, 1, []) == -22
    assert co.get_request_key(0, 404)[0] == -2


def test_evict_last_entry_removes_key_and_engine_mapping():   # in_memory.go:243-256
    ix = ko.InMemoryIndex()
    ix.add([K(1)], [K(100)], [E("a")])
    ix.evict(K(1), [E("a")])
    assert ix.lookup([K(100)]) == {} and len(ix.index if False else ix.data) == 0
    with pytest.raises(ko.IndexError_):
        ix.get_request_key(K(1))


def test_e2e_shape_score_equals_cached_prefix_blocks():      # e2e_test.go:134-205 (block size 4)
    ix = ko.Indexer(block_size=4)
    toks = list(range(100, 121))                             # 21 tokens -> 5 keys
    keys = ix.tokens_processor.tokens_to_kv_block_keys(None, toks, M)
    assert len(keys) == 5
    assert ix.get_pod_scores(toks, M) == {}                  # CacheMiss: empty map, not nil
    ix.index.add([ko.Key(M, 1000 + i) for i in range(5)], keys, [E("pod1")])
    assert ix.get_pod_scores(toks, M, ["pod1"]) == {"pod1": 5.0}       # CacheHit
    assert ix.get_pod_scores(toks[:9], M) == {"pod1": 2.0}             # PrefixReduction
    assert ix.get_pod_scores(toks + [7] * 8, M) == {"pod1": 5.0}       # PrefixExpansion
    assert ix.get_pod_scores(toks[:3], M) is None                      # no full block -> (nil, nil)


def test_get_hash_as_uint64():                               # kvevents/pool.go:343-367
    g = ko.get_hash_as_uint64
    assert g(5) == 5 and g(-1) == (1 << 64) - 1
    assert g(bytes(range(1, 11))) == int.from_bytes(bytes(range(3, 11)), "big")
    assert g(b"\x01\x02") == 0x0102
    with pytest.raises(ValueError):
        g(b"")
    with pytest.raises(TypeError):
        g("str")
    with pytest.raises(TypeError):
        g(1.5)


def test_queue_sharding():                                   # kvevents/pool.go:132-144
    pool = ko.EventsPool(ko.InMemoryIndex(), ko.ChunkedTokenDatabase(), concurrency=4)
    assert pool.queue_index("pod-1") == ko.fnv32a(b"pod-1") % 4
    assert len({pool.queue_index("pod-%d" % i) for i in range(64)}) == 4


def test_instrumented_index_counts():
    """InstrumentedIndex (kvblock/instrumented_index.go:35-92): admissions = len(requestKeys) per Add (also when Add fails),
    evictions = len(entries) per Evict, one request per Lookup, hits = max over pods of that pod's entries in the result."""
    ix = ko.InstrumentedIndex(ko.InMemoryIndex(size=100, pod_cache_size=10))
    m = ix.metrics
    k = [ko.Key("m", h) for h in (11, 12, 13)]
    e = [ko.Key("m", h) for h in (21, 22, 23)]
    ix.add(e, k, [ko.PodEntry("a", "gpu"), ko.PodEntry("b", "gpu")])
    ix.add(e[:2], k[:2], [ko.PodEntry("a", "cpu")])                     # pod a on two tiers of keys 11, 12
    assert m.admissions_total == 5
    with pytest.raises(ko.IndexError_):
        ix.add(e[:2], k[:1], [ko.PodEntry("a", "gpu")])                 # length mismatch: still counted (len(requestKeys) = 1)
    assert m.admissions_total == 6
    hits = ix.lookup(k + [ko.Key("m", 99)])                             # a: 2 + 2 + 1 entries, b: 3
    assert len(hits) == 3 and (m.lookup_requests_total, m.max_pod_hit_count_total, m.lookup_hits_total, m.lookup_latency_count) == (1, 5, 5, 1)
    ix.lookup(k, {"b"})
    assert (m.lookup_requests_total, m.max_pod_hit_count_total) == (2, 8)
    with pytest.raises(ko.IndexError_):
        ix.lookup([])                                                   # an error: request and latency counted, no hits
    assert (m.lookup_requests_total, m.lookup_latency_count, m.lookup_hits_total) == (3, 3, 8)
    ix.evict(e[0], [ko.PodEntry("a", "gpu"), ko.PodEntry("a", "cpu")])
    ix.evict(ko.Key("m", 404), [ko.PodEntry("a", "gpu")])               # unknown engine key: silent no-op, still counted
    assert m.evictions_total == 3
