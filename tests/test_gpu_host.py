"""GPU: the string-level host mirror (kvidx_host.h: Indexer / Index / Pool with msgpack KVEvents) end to end against the
Python oracle fed with the SAME wire payloads."""
import msgpack
import numpy as np
import pytest

from helpers import golden
from kvidx.host import HostIndexer
from oracle import kvoracle as ko

pytestmark = pytest.mark.gpu


def test_pool_and_indexer_strings_end_to_end():
    sc = golden("scenario_small.json")
    tiers = [(t, sc["weights"].get(t, 1.0)) for t in ("gpu", "cpu")]
    h = HostIndexer(block_size=sc["block_size"], hash_seed=sc["hash_seed"], capacity=4096, pods_per_key=sc["pod_cache_size"], tiers=tiers, max_pods=64)
    ix = ko.Indexer(block_size=sc["block_size"], hash_seed=sc["hash_seed"], size=10 ** 6, pod_cache_size=sc["pod_cache_size"], weights=sc["weights"])
    pool = ko.EventsPool(ix.index, ix.tokens_processor)
    model = sc["model"]
    n_msgs = 0
    for i, e in enumerate(sc["events"]):
        if e["type"] == "BlockStored":
            body = ["BlockStored", e["hashes"], e["parent"], e["tokens"], sc["block_size"], None, e["medium"]]
        elif e["type"] == "BlockRemoved":
            body = ["BlockRemoved", e["hashes"], e["medium"]]
        else:
            body = ["AllBlocksCleared"]
        payload = msgpack.packb([1700000000.0 + i, [body], 0])
        pool.process_event(e["pod"], model, payload)                      # oracle: processEvent + digestEvents
        assert h.add_task(e["pod"], model, payload) == 0                 # Pool.AddTask
        assert h.queue_index(e["pod"]) == pool.queue_index(e["pod"])
        n, dropped = h.process()
        assert n >= 0
        n_msgs += 1
    from kvidx import _native
    import ctypes as C
    st = _native.Stats()
    assert _native.load().kvidx_get_stats(C.c_void_p(h.L.kvhost_index(h.h)), C.byref(st)) == 0
    assert st.request_keys == len(ix.index.data) and st.engine_keys == len(ix.index.engine_to_request)
    for p in sc["prompts"]:
        exp = ix.get_pod_scores(p["tokens"], model, p["filter"])
        got = h.get_pod_scores(p["tokens"], model, p["filter"])
        assert got == exp, (p["filter"], exp, got)
        keys = [k.chunk_hash for k in ix.tokens_processor.tokens_to_kv_block_keys(None, p["tokens"], model)]
        if keys:
            uk = list(dict.fromkeys(keys))
            rc, look = h.lookup(model, uk, p["filter"])
            want = {k.chunk_hash: [(e.pod, e.tier) for e in v] for k, v in ix.index.lookup([ko.Key(model, x) for x in uk], set(p["filter"])).items()}
            assert rc == 0 and look == want
    # Index interface with strings + error cases (in_memory.go:108-110,150-155,213-215,266-268)
    assert h.add(model, [], [], [("p", "gpu")]) == -22 and h.add(model, [1, 2], [3], [("p", "gpu")]) == -22 and h.evict(model, 1, []) == -22
    assert h.get_request_key(model, 404040)[0] == -2
    assert h.add("other-model", [2 ** 40 + 1], [2 ** 41 + 7], [("pod-a", "GPU"), ("pod-b", "cpu")]) == 0
    assert h.lookup("other-model", [2 ** 41 + 7])[1] == {2 ** 41 + 7: [("pod-a", "gpu"), ("pod-b", "cpu")]}
    assert h.lookup(model, [2 ** 41 + 7])[1] == {}                       # model name is part of the key identity (index.go:138-141)
    assert h.get_request_key("other-model", 2 ** 40 + 1) == (0, 2 ** 41 + 7)
    assert h.evict("other-model", 2 ** 40 + 1, [("pod-a", "gpu"), ("pod-b", "cpu")]) == 0
    assert h.lookup("other-model", [2 ** 41 + 7])[1] == {}
    assert h.get_pod_scores([1, 2, 3], model) is None                    # (nil, nil): no full block


def test_metrics_match_the_instrumented_index():
    """kvcache_index_* counters of the host mirror (enable_metrics) against the oracle's InstrumentedIndex fed with the same
    events, GetPodScores calls and Index calls (kvblock/instrumented_index.go:35-92)."""
    sc = golden("scenario_small.json")
    tiers = [(t, sc["weights"].get(t, 1.0)) for t in ("gpu", "cpu")]
    h = HostIndexer(block_size=sc["block_size"], hash_seed=sc["hash_seed"], capacity=4096, pods_per_key=sc["pod_cache_size"], tiers=tiers, max_pods=64,
                    enable_metrics=True)
    ix = ko.Indexer(block_size=sc["block_size"], hash_seed=sc["hash_seed"], size=10 ** 6, pod_cache_size=sc["pod_cache_size"], weights=sc["weights"])
    ix.index = ko.InstrumentedIndex(ix.index)
    pool = ko.EventsPool(ix.index, ix.tokens_processor)
    model = sc["model"]
    for i, e in enumerate(sc["events"]):
        if e["type"] == "BlockStored":
            body = ["BlockStored", e["hashes"], e["parent"], e["tokens"], sc["block_size"], None, e["medium"]]
        elif e["type"] == "BlockRemoved":
            body = ["BlockRemoved", e["hashes"], e["medium"]]
        else:
            body = ["AllBlocksCleared"]
        payload = msgpack.packb([1700000000.0 + i, [body], 0])
        pool.process_event(e["pod"], model, payload)
        assert h.add_task(e["pod"], model, payload) == 0
        if i % 3 == 2:
            h.process()
    h.process()
    for p in sc["prompts"]:
        assert h.get_pod_scores(p["tokens"], model, p["filter"]) == ix.get_pod_scores(p["tokens"], model, p["filter"])
    longest = max(sc["prompts"], key=lambda p: len(p["tokens"]))["tokens"]
    keys = [k.chunk_hash for k in ix.tokens_processor.tokens_to_kv_block_keys(None, longest, model)]
    uk = list(dict.fromkeys(keys))
    assert uk
    h.lookup(model, uk); ix.index.lookup([ko.Key(model, x) for x in uk])
    h.lookup(model, uk, ["no-such-pod"]); ix.index.lookup([ko.Key(model, x) for x in uk], {"no-such-pod"})
    assert h.lookup(model, [])[0] == -22
    with pytest.raises(ko.IndexError_):
        ix.index.lookup([])
    assert h.add(model, [1, 2], [3], [("p", "gpu")]) == -22
    with pytest.raises(ko.IndexError_):
        ix.index.add([ko.Key(model, 1), ko.Key(model, 2)], [ko.Key(model, 3)], [ko.PodEntry("p", "gpu")])
    assert h.evict(model, 777, [("p", "gpu"), ("q", "cpu")]) == 0
    ix.index.evict(ko.Key(model, 777), [ko.PodEntry("p", "gpu"), ko.PodEntry("q", "cpu")])
    got, want = h.metrics(), ix.index.metrics
    assert want.admissions_total > 0 and want.evictions_total > 0 and want.max_pod_hit_count_total > 0
    for name in ("admissions_total", "evictions_total", "lookup_requests_total", "max_pod_hit_count_total", "lookup_hits_total", "lookup_latency_count"):
        assert got[name] == getattr(want, name), (name, got[name], getattr(want, name))
    assert got["lookup_latency_bucket"][-1] <= got["lookup_latency_count"] and got["lookup_latency_sum"] > 0
    assert "kvcache_index_lookup_requests_total %d\n" % want.lookup_requests_total in h.metrics_text()
    plain = HostIndexer(block_size=sc["block_size"], capacity=1024, max_pods=64)           # metrics off: nothing is counted
    plain.add(model, [1], [2], [("p", "gpu")]); plain.lookup(model, [2])
    assert plain.metrics()["admissions_total"] == 0 and plain.metrics()["lookup_requests_total"] == 0
