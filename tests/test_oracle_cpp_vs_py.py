"""CPU: randomised differential test, C++ oracle vs Python oracle (same id-level inputs)."""
import numpy as np
import pytest

from helpers import EVENT_DTYPE, csr
from oracle import kvoracle as ko
from oracle.kvoracle_c import COracle


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_random_streams(seed):
    rng = np.random.default_rng(seed)
    BS, P, NT = 4, 12, 3
    w = [1.0, 0.8, 0.3]
    tiers = ["gpu", "cpu", "disk"]
    py = ko.Indexer(block_size=BS, hash_seed="x", size=40, pod_cache_size=3, weights=dict(zip(tiers, w)))
    pool = ko.EventsPool(py.index, py.tokens_processor)
    co = COracle(block_size=BS, init_hash=ko.fnv64a(b"x"), size=40, pod_cache_size=3, tier_weights=w, max_pods=64)
    docs = [rng.integers(0, 1 << 17, size=BS * int(rng.integers(1, 9))).tolist() for _ in range(10)]
    for step in range(300):
        pod = int(rng.integers(0, P)); tier = int(rng.integers(0, NT))
        d = int(rng.integers(0, len(docs))); nb = len(docs[d]) // BS
        if rng.random() < 0.65:
            b0 = int(rng.integers(0, nb)); b1 = int(rng.integers(b0 + 1, nb + 1))
            hashes = [d * 100 + b for b in range(b0, b1)]
            toks = docs[d][b0 * BS:b1 * BS]
            parent = d * 100 + b0 - 1 if b0 > 0 else None
            pool.digest_events("p%d" % pod, "m", [ko.BlockStored(hashes, parent, toks, 0, None, tiers[tier].upper())])
            ev = np.zeros(1, EVENT_DTYPE)
            ev["op"] = 0; ev["has_parent"] = parent is not None; ev["parent_hash"] = parent or 0
            ev["podtier"] = (pod << 4) | tier; ev["n_hashes"] = len(hashes); ev["n_tokens"] = len(toks)
            assert co.apply_events(ev, np.array(hashes, np.uint64), np.array(toks, np.uint32))[0] == 0
        else:
            hashes = [d * 100 + int(rng.integers(0, nb)) for _ in range(int(rng.integers(1, 4)))]
            pool.digest_events("p%d" % pod, "m", [ko.BlockRemoved(hashes, tiers[tier])])
            ev = np.zeros(1, EVENT_DTYPE)
            ev["op"] = 1; ev["podtier"] = (pod << 4) | tier; ev["n_hashes"] = len(hashes)
            assert co.apply_events(ev, np.array(hashes, np.uint64), np.zeros(0, np.uint32))[0] == 0
        if step % 10 == 9:      # key-LRU cap (size=40) is exercised: recency from Lookup matters
            q = docs[int(rng.integers(0, len(docs)))]
            q = q[:BS * int(rng.integers(0, len(q) // BS + 1))] + rng.integers(0, 1 << 17, size=int(rng.integers(0, 6))).tolist()
            exp = py.get_pod_scores(q, "m")
            tok, off = csr([q])
            scores, has, _, _ = co.score_batch(tok, off)
            row = np.full(64, -1.0)
            for p, s in (exp or {}).items():
                row[int(p[1:])] = s
            assert np.array_equal(scores[0], row) and bool(has[0]) == (exp is not None)
            assert co.len_request() == len(py.index.data) and co.len_engine() == len(py.index.engine_to_request)
