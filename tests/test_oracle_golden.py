"""CPU: pin both oracles (pure-Python restatement and the fast C++ port) against the committed
golden vectors.  hash_kats.json / kv_event_base_keys.json were produced WITHOUT the oracle
(cbor2 + a from-scratch FNV loop, tests/golden/make_golden.py), scenario_small.json by the
Python oracle."""
import numpy as np
import pytest

from helpers import Interner, csr, dense_from_map, filter_mask, golden, scenario_events
from oracle import kvoracle as ko
from oracle.kvoracle_c import COracle, lib as colib

KATS = golden("hash_kats.json")


def test_fnv64a_vectors():
    for s, v in KATS["fnv64a"].items():
        assert ko.fnv64a(s.encode()) == v
        assert colib().ko_fnv64a(s.encode(), len(s)) == v
    # FNV-32a reference vectors (queue sharding, kvevents/pool.go:135-142)
    assert ko.fnv32a(b"") == 0x811C9DC5 and ko.fnv32a(b"a") == 0xE40C292C and ko.fnv32a(b"foobar") == 0xBF9CF968


def test_cbor_rfc8949_appendix_a_vectors():
    """RFC 8949 Appendix A (the published CBOR examples; RFC 7049 has the same table): unsigned integers in shortest
    form, definite-length arrays, null -- every item kind of the block payload [parent, [tokens...], nil]
    (token_processor.go:94-103, fxamacker/cbor CanonicalEncOptions)."""
    uints = {0: "00", 1: "01", 10: "0a", 23: "17", 24: "1818", 25: "1819", 100: "1864", 1000: "1903e8", 1000000: "1a000f4240",
             1000000000000: "1b000000e8d4a51000", 18446744073709551615: "1bffffffffffffffff"}
    for v, hx in uints.items():
        assert ko.cbor_head(0, v).hex() == hx
    assert ko.cbor_head(4, 0).hex() == "80" and ko.cbor_head(4, 3).hex() == "83" and ko.cbor_head(4, 25).hex() == "9819"
    # [1, 2, 3] -> 83010203 ; [1, [2, 3], [4, 5]] -> 8301820203820405 ; [1..25] -> 98190102..1718181819 ; null -> f6
    assert (ko.cbor_head(4, 3) + b"".join(ko.cbor_head(0, v) for v in (1, 2, 3))).hex() == "83010203"
    assert (ko.cbor_head(4, 25) + b"".join(ko.cbor_head(0, v) for v in range(1, 26))).hex() == \
        "98190102030405060708090a0b0c0d0e0f101112131415161718181819"
    # the block payload is the Appendix A nesting [1, [2, 3], <third item>] with null (f6) as the third item
    assert ko.cbor_block_payload(1, [2, 3]).hex() == "8301820203f6"
    assert ko.cbor_block_payload(1000000000000, list(range(1, 26))).hex() == \
        "831b000000e8d4a5100098190102030405060708090a0b0c0d0e0f101112131415161718181819f6"
    # and its hash is FNV-64a over exactly those bytes, in both oracles
    assert ko.block_hash(1, [2, 3]) == ko.fnv64a(bytes.fromhex("8301820203f6"))
    co = COracle(block_size=2, init_hash=1)
    keys, _ = co.hash_keys(np.array([2, 3], np.uint32), np.array([0, 2], np.int64))
    assert int(keys[0]) == ko.fnv64a(bytes.fromhex("8301820203f6"))


@pytest.mark.parametrize("case", KATS["cases"], ids=[c["name"] for c in KATS["cases"]])
def test_hash_chain_kats(case):
    tp = ko.ChunkedTokenDatabase(case["block_size"], case["seed"])
    parent = None if case["parent"] is None else ko.Key("m", case["parent"])
    got = [k.chunk_hash for k in tp.tokens_to_kv_block_keys(parent, case["tokens"], "m")]
    assert got == case["keys"]
    co = COracle(block_size=case["block_size"], init_hash=ko.fnv64a(case["seed"].encode()))
    tok, off = csr([case["tokens"]])
    par = None if case["parent"] is None else np.array([case["parent"]], np.uint64)
    keys, koff = co.hash_keys(tok, off, par)
    assert [int(k) for k in keys] == case["keys"] and koff[-1] == len(case["keys"])


def test_reference_fixture_tokens():
    g = golden("kv_event_base_keys.json")
    tp = ko.ChunkedTokenDatabase(g["block_size"], g["hash_seed"])
    keys = [k.chunk_hash for k in tp.tokens_to_kv_block_keys(None, g["token_ids"], "Qwen/Qwen3-0.6B")]
    assert keys == g["request_keys"] and len(keys) == 25 == len(g["engine_hashes"])


def _replay_py(sc):
    ix = ko.Indexer(block_size=sc["block_size"], hash_seed=sc["hash_seed"], size=10 ** 6,
                    pod_cache_size=sc["pod_cache_size"], weights=sc["weights"])
    pool = ko.EventsPool(ix.index, ix.tokens_processor)
    for e in sc["events"]:
        if e["type"] == "BlockStored":
            pool.digest_events(e["pod"], sc["model"], [ko.BlockStored(e["hashes"], e["parent"], e["tokens"], 0, None, e["medium"])])
        elif e["type"] == "BlockRemoved":
            pool.digest_events(e["pod"], sc["model"], [ko.BlockRemoved(e["hashes"], e["medium"])])
    return ix


def test_scenario_python_oracle_is_deterministic():
    sc = golden("scenario_small.json")
    ix = _replay_py(sc)
    assert len(ix.index.data) == sc["final_request_keys"] and len(ix.index.engine_to_request) == sc["final_engine_keys"]
    for p in sc["prompts"]:
        got = ix.get_pod_scores(p["tokens"], sc["model"], p["filter"])
        assert got == p["scores"]


def test_scenario_cpp_oracle_matches_golden():
    sc = golden("scenario_small.json")
    pods, tiers = Interner(sc["pods"]), Interner(sc["tiers"])
    P = 64
    w = [sc["weights"].get(t, 1.0) for t in sc["tiers"]]
    co = COracle(block_size=sc["block_size"], init_hash=ko.fnv64a(sc["hash_seed"].encode()), size=10 ** 6,
                 pod_cache_size=sc["pod_cache_size"], tier_weights=w, max_pods=P)
    ev, hs, tk = scenario_events(sc, pods, tiers)
    rc, dropped = co.apply_events(ev, hs, tk)
    assert rc == 0 and dropped > 0
    assert co.len_request() == sc["final_request_keys"] and co.len_engine() == sc["final_engine_keys"]
    tok, off = csr([p["tokens"] for p in sc["prompts"]])
    fm = np.stack([filter_mask([pods.ids[x] for x in p["filter"]], co.filter_words) for p in sc["prompts"]])
    scores, has, _, _ = co.score_batch(tok, off, filter_mask=fm)
    keys, koff = co.hash_keys(tok, off)
    for i, p in enumerate(sc["prompts"]):
        assert [int(k) for k in keys[koff[i]:koff[i + 1]]] == p["keys"]
        assert bool(has[i]) == bool(p["keys"])
        exp = dense_from_map(p["scores"], pods, P)
        assert np.array_equal(scores[i], exp), (i, p["scores"])       # bit-exact float64
        if p["keys"]:
            rc, pt, cnt = co.lookup(0, np.array(p["keys"], np.uint64), fm[i] if p["filter"] else None)
            assert rc == 0
            for j, ents in enumerate(p["lookup"]):
                got = [(int(e) >> 4, int(e) & 15) for e in pt[j, :cnt[j]]]
                assert got == [(pods.ids[a], tiers.ids[b]) for a, b in ents]   # oldest -> newest order
