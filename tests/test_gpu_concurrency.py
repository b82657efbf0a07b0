"""GPU (-m gpu): concurrency of the C ABI -- the read path next to the write path, many OS threads, coalesced Score()
callers -- and the two-phase write path, against the oracle."""
import threading

import numpy as np
import pytest

from helpers import EVENT_DTYPE, csr, filter_mask
import kvidx
from kvidx import synth
from oracle.kvoracle_c import COracle
from test_gpu_parity import PT, _index_pair, _random_stream

pytestmark = pytest.mark.gpu


def test_concurrent_operations_100_threads():
    """pkg/kvcache/kvblock/index_test.go:214-278 (testConcurrentOperations): 100 workers x 10 operations -- Add, Lookup
    (must see the entry the worker just added), Evict, Lookup (must not see it any more) -- all at once through the C ABI.
    The reference raises PodCacheSize to 500 so that 400 pods fit ONE key; a slot here holds the reference's DEFAULT of 10
    (in_memory.go:34), so every worker asserts read-your-write on a key of its own while all 100 also hammer one shared
    key, whose state is checked structurally (<= 10 entries, no duplicates, only pods somebody added)."""
    ix = kvidx.Index(capacity=4096, max_pods=1024)
    ENG_S, REQ_S = 38894120, 72568158
    errs = []
    start = threading.Barrier(100)

    def worker(wid):
        try:
            eng, req = 10_000_000 + wid, 20_000_000 + wid
            start.wait()
            for op in range(10):
                pod = wid * 10 + op if op % 3 == 0 else None
                if op % 3 == 0:                                   # Add
                    assert ix.add(0, [eng], [req], [PT(pod)]) == 0
                    assert ix.add(0, [ENG_S], [REQ_S], [PT(pod)]) == 0
                elif op % 3 == 1:                                 # Lookup: contains what this worker added one op ago
                    rc, pt, cnt = ix.lookup(0, [req])
                    assert rc == 0 and PT(wid * 10 + op - 1) in pt[0, :cnt[0]].tolist(), (wid, op, pt[0, :cnt[0]].tolist())
                    rc, pt, cnt = ix.lookup(0, [REQ_S])
                    ent = pt[0, :cnt[0]].tolist()
                    assert rc == 0 and cnt[0] <= 10 and len(set(ent)) == len(ent) and all((e >> 4) < 1000 and (e >> 4) % 10 in (0, 3, 6, 9) for e in ent), ent
                else:                                             # Evict: gone afterwards
                    mine = PT(wid * 10 + op - 2)
                    assert ix.evict(0, eng, [mine]) == 0 and ix.evict(0, ENG_S, [mine]) == 0
                    rc, pt, cnt = ix.lookup(0, [req])
                    assert rc == 0 and mine not in pt[0, :cnt[0]].tolist()
                    rc, pt, cnt = ix.lookup(0, [REQ_S])
                    assert rc == 0 and mine not in pt[0, :cnt[0]].tolist()
        except Exception as ex:                                   # noqa: BLE001
            errs.append(ex)

    ths = [threading.Thread(target=worker, args=(i,)) for i in range(100)]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    assert not errs, errs[:3]
    # the index still works, and every worker's own key holds exactly its last Add (op 9)
    keys = np.arange(100, dtype=np.uint64) + 20_000_000
    rc, pt, cnt = ix.lookup(0, keys)
    assert rc == 0 and (cnt == 1).all() and pt[:, 0].tolist() == [PT(w * 10 + 9) for w in range(100)]


def test_duplicate_hashes_inside_one_event_and_one_add_call():
    """The reference adds the pairs of an event / an Add call IN ORDER (in_memory.go:159-203): an engine key that appears
    twice keeps its LAST request key.  Lanes of one warp take the pairs in parallel here, so duplicates are resolved
    explicitly; BlockRemoved with a repeated hash must not hang on the slot it already owns."""
    BS = 16
    rng = np.random.default_rng(8)
    for phase1 in ("0", "1"):
        import os
        os.environ["KVIDX_WRITE_PHASE1"] = phase1
        try:
            ix, co = _index_pair(capacity=4096, max_pods=16)
        finally:
            del os.environ["KVIDX_WRITE_PHASE1"]
        toks = rng.integers(0, 128256, size=BS * 40).astype(np.uint32)
        hashes = np.arange(40, dtype=np.uint64) + 500
        hashes[7] = hashes[3]; hashes[35] = hashes[3]; hashes[20] = hashes[19]      # same warp pass, and across two passes
        ev = np.zeros(2, EVENT_DTYPE)
        ev[0]["op"] = 0; ev[0]["podtier"] = PT(2); ev[0]["n_hashes"] = 40; ev[0]["n_tokens"] = BS * 40
        ev[1]["op"] = 0; ev[1]["podtier"] = PT(3, 1); ev[1]["n_hashes"] = 40; ev[1]["n_tokens"] = BS * 40      # same blocks, other pod
        for one in (ev[:1], ev[1:]):                       # one call per pod: entry order inside a slot is then the oracle's
            assert ix.apply_events(one, hashes, toks) == (0, 0) and co.apply_events(one, hashes, toks) == (0, 0)
        for h in np.unique(hashes):
            assert ix.get_request_key(0, int(h)) == co.get_request_key(0, int(h)), int(h)
        st = ix.stats()
        assert st["request_keys"] == co.len_request() and st["engine_keys"] == co.len_engine()
        # remove with duplicates: [3, 3, 19, 19, 3]
        rm = np.zeros(1, EVENT_DTYPE)
        rm[0]["op"] = 1; rm[0]["podtier"] = PT(2); rm[0]["n_hashes"] = 5
        rh = np.array([hashes[3], hashes[3], hashes[19], hashes[20], hashes[35]], np.uint64)
        assert ix.apply_events(rm, rh, np.zeros(0, np.uint32))[0] == 0 and co.apply_events(rm, rh, np.zeros(0, np.uint32))[0] == 0
        keys, _ = ix.hash_keys(toks, [0, len(toks)])
        r1, r2 = ix.lookup(0, keys), co.lookup(0, keys)
        assert np.array_equal(r1[2], r2[2])
        for i in range(len(keys)):
            assert np.array_equal(r1[1][i, :r1[2][i]], r2[1][i, :r2[2][i]])
    # Index.Add with repeated engine keys far apart in one call (different warps)
    ix, co = _index_pair(capacity=1 << 14, max_pods=16)
    eng = rng.integers(1, 1 << 60, 3000, dtype=np.uint64)
    req = rng.integers(1, 1 << 60, 3000, dtype=np.uint64)
    eng[2500] = eng[10]; eng[2999] = eng[10]; eng[1700] = eng[1699]
    assert ix.add(0, eng, req, [PT(1), PT(2, 1)]) == 0 and co.add(0, eng, req, [PT(1), PT(2, 1)]) == 0
    for e in (eng[10], eng[1699], eng[5]):
        assert ix.get_request_key(0, int(e)) == co.get_request_key(0, int(e))
    assert ix.stats()["engine_keys"] == co.len_engine() and ix.stats()["request_keys"] == co.len_request()


@pytest.mark.parametrize("phase1", ["1", "0"])
def test_two_phase_write_path_vs_oracle(phase1, monkeypatch):
    """One big kvidx_apply_events call per step, pods on disjoint documents (every schedule gives the same index): chains
    continued through parents stored EARLIER IN THE SAME BATCH (at the last block of the producer, in its middle, by a
    re-stored block), parents only the index knows, unknown parents (chain restarts at the seed), removals in between
    that take the parent away again -- phase 1 predicts, phase 2 verifies.  Index and scores equal the oracle's."""
    monkeypatch.setenv("KVIDX_WRITE_PHASE1", phase1)
    rng = np.random.default_rng(33)
    BS, P = 16, 48
    ix, co = _index_pair(block_size=BS, capacity=1 << 17, max_pods=64)
    docs = [rng.integers(0, 128256, size=BS * int(rng.integers(2, 120))).tolist() for _ in range(P)]
    for step in range(4):
        evs, hss, tks = [], [], []
        ho = to = 0
        for pod in range(P):
            e, h, t = _random_stream(rng, 40, BS, 1, 2, [docs[pod]])
            e["podtier"] = (pod << 4) | (e["podtier"] & 15)
            e["hash_off"] += ho; e["tok_off"] += to
            e["parent_hash"] += np.uint64(pod) * np.uint64(1 << 32); h = h + np.uint64(pod) * np.uint64(1 << 32)
            ho += len(h); to += len(t)
            evs.append(e); hss.append(h); tks.append(t)
        ev = np.stack(evs, axis=1).reshape(-1)
        hs, tk = np.concatenate(hss), np.concatenate(tks)
        rc, d1 = ix.apply_events(ev, hs, tk)
        rc2, d2 = co.apply_events(ev, hs, tk)
        assert rc == rc2 == 0 and d1 == d2
        st = ix.stats()
        assert st["request_keys"] == co.len_request() and st["engine_keys"] == co.len_engine(), step
    if phase1 == "1":
        # most predictions hold; the ones that do not are events whose parent was REMOVED by an earlier event of the same batch
        # (this stream removes 30 % of the time and every document lives on one pod, so a removal often takes the mapping away)
        assert 0 < ix.stats()["rehashed_events"] < 0.35 * 4 * 40 * P
    tok, off = csr([d[:BS * int(rng.integers(0, len(d) // BS + 1))] for d in docs])
    s1, _ = ix.score_batch(tok, off)
    s2, _, _, _ = co.score_batch(tok, off)
    assert np.array_equal(s1, s2)
    uk = np.unique(ix.hash_keys(tok, off)[0])
    r1, r2 = ix.lookup(0, uk), co.lookup(0, uk)
    assert np.array_equal(r1[2], r2[2])
    for i in range(len(uk)):
        assert np.array_equal(r1[1][i, :r1[2][i]], r2[1][i, :r2[2][i]])


def test_fill_workload_through_both_write_paths_equal():
    """The benchmark's index fill (4 pods per document, chunks chained through parents inside one batch) through phase 1 and
    through the in-place path: same keys, same entries, no event re-hashed."""
    wl = synth.Workload(7, 2048, 1 << 15, 64)
    ev, hs, tk = wl.fill_events(0, wl.D)
    import os
    res = []
    for phase1 in ("1", "0"):
        os.environ["KVIDX_WRITE_PHASE1"] = phase1
        try:
            ix = kvidx.Index(capacity=1 << 16, max_pods=64)
        finally:
            del os.environ["KVIDX_WRITE_PHASE1"]
        assert ix.apply_events(ev, hs, tk) == (0, 0)
        st = ix.stats()
        assert st["request_keys"] == wl.n_blocks and st["rehashed_events"] == 0
        toks, doc, m = wl.queries(0, 500)
        off = np.arange(0, (len(toks) + 1) * wl.T, wl.T, dtype=np.int64)
        s, _ = ix.score_batch(toks.reshape(-1), off)
        assert np.array_equal(s, wl.expected_scores(doc, m))
        res.append(s)
    assert np.array_equal(res[0], res[1])


def test_pod_ids_beyond_max_pods_are_refused():
    """Score rows and filter rows are max_pods wide: an entry of a pod id beyond that could be neither scored nor filtered."""
    ix = kvidx.Index(capacity=1024, max_pods=8)
    assert ix.add(0, [1], [2], [PT(7)]) == 0
    assert ix.add(0, [1], [2], [PT(8)]) == kvidx.ERANGE and "max_pods" in ix.last_error()
    assert ix.evict(0, 1, [PT(9)]) == kvidx.ERANGE
    ev = np.zeros(1, EVENT_DTYPE)
    ev[0]["op"] = 1; ev[0]["podtier"] = PT(8); ev[0]["n_hashes"] = 1
    assert ix.apply_events(ev, np.array([1], np.uint64), np.zeros(0, np.uint32))[0] == kvidx.ERANGE
    rc, pt, cnt = ix.lookup(0, [2], filter_mask([7], ix.filter_words))
    assert rc == 0 and pt[0, :cnt[0]].tolist() == [PT(7)]


def test_readers_see_whole_slot_images_while_writers_update():
    """A writer thread keeps adding / refreshing / evicting entries of a handful of keys while reader threads look them up
    and score a prompt over them.  Every slot image a reader gets must be a state some prefix of the writer's operations
    produces: entries unique, count consistent, the permanent entry always present -- never a half-updated slot."""
    BS = 16
    rng = np.random.default_rng(4)
    ix = kvidx.Index(capacity=1 << 12, max_pods=64)
    toks = rng.integers(0, 128256, size=BS * 8).astype(np.uint32)
    keys, _ = ix.hash_keys(toks, [0, len(toks)])
    eng = (keys ^ np.uint64(0x99)).astype(np.uint64)
    assert ix.add(0, eng, keys, [PT(1)]) == 0                      # permanent entry: pod 1 on every block
    stop = threading.Event()
    errs = []

    def writer():
        try:
            i = 0
            while not stop.is_set():
                pod = 2 + (i % 12)
                assert ix.add(0, eng, keys, [PT(pod), PT(1)]) == 0     # new pod + refresh of the permanent one (entries move)
                if i % 3 == 2:
                    for e in eng[: 4]:
                        assert ix.evict(0, int(e), [PT(2 + ((i - 2) % 12))]) == 0
                i += 1
        except Exception as ex:                                       # noqa: BLE001
            errs.append(ex)

    def reader():
        try:
            while not stop.is_set():
                rc, pt, cnt = ix.lookup(0, keys)
                assert rc == 0
                for j in range(len(keys)):
                    ent = pt[j, :cnt[j]].tolist()
                    assert 1 <= cnt[j] <= 10 and PT(1) in ent and len(set(ent)) == len(ent), ent
                    assert all(1 <= (e >> 4) <= 13 and (e & 15) == 0 for e in ent), ent
                s, has = ix.score_batch(toks, [0, len(toks)])
                assert has[0] == 1 and s[0, 1] == 8.0, s[0][s[0] >= 0]      # pod 1 holds all 8 blocks at every instant
        except Exception as ex:                                       # noqa: BLE001
            errs.append(ex)

    ths = [threading.Thread(target=writer)] + [threading.Thread(target=reader) for _ in range(3)]
    for t in ths:
        t.start()
    import time
    time.sleep(3.0)
    stop.set()
    for t in ths:
        t.join()
    assert not errs, errs[:2]


def test_concurrent_score_callers_are_coalesced():
    """64 OS threads, one prompt per call (the shape of the gRPC server: one goroutine per RPC, server.go:70-96), dense and
    sparse calls mixed.  Callers that arrive while a batch is on the device are served together by the next launch; every
    caller gets exactly its own rows."""
    wl = synth.Workload(8, 1024, 1 << 14, 32)
    ix, co = _index_pair(capacity=1 << 15, max_pods=32)
    ev, hs, tk = wl.fill_events(0, wl.D)
    assert ix.apply_events(ev, hs, tk) == (0, 0) and co.apply_events(ev, hs, tk) == (0, 0)
    toks, doc, m = wl.queries(0, 64 * 20)
    want = wl.expected_scores(doc, m)
    errs = []
    go = threading.Barrier(64)

    def caller(c):
        try:
            go.wait()
            for it in range(20):
                q = c * 20 + it
                L = wl.T if it % 4 else int(wl.T * 0.6)               # ragged lengths across callers
                exp = want[q] if it % 4 else co.score_batch(toks[q, :L], [0, L])[0][0]
                if c % 2 == 0:
                    s, has = ix.score_batch(toks[q, :L], [0, L])
                    assert np.array_equal(s[0], exp), (c, it)
                else:
                    pods, sc, cnt, has = ix.score_batch_sparse(toks[q, :L], [0, L])
                    got = {int(pods[0, j]): float(sc[0, j]) for j in range(cnt[0])}
                    assert got == {int(p): float(exp[p]) for p in np.nonzero(exp >= 0)[0]}, (c, it)
        except Exception as ex:                                       # noqa: BLE001
            errs.append(ex)

    ths = [threading.Thread(target=caller, args=(c,)) for c in range(64)]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    assert not errs, errs[:2]
    assert ix.stats()["coalesced_calls"] > 0


def test_apply_events_dev_and_sparse_dev():
    """Device-resident write batch (pod-sorted events, queue offsets) followed, without any host synchronisation, by a
    device-resident sparse Score(): the read is ordered after the asynchronous write batch."""
    import torch
    wl = synth.Workload(4, 1024, 1 << 13, 32)
    ix, co = _index_pair(capacity=1 << 14, max_pods=32)
    ev, hs, tk = wl.fill_events(0, wl.D)
    assert co.apply_events(ev, hs, tk) == (0, 0)
    order = np.argsort(ev["podtier"] >> 4, kind="stable")
    evs = ev[order]
    pods = (evs["podtier"] >> 4).astype(np.int64)
    qoff = np.concatenate([[0], np.nonzero(np.diff(pods))[0] + 1, [len(evs)]]).astype(np.int64)
    dev = torch.device("cuda", 0)
    d_ev = torch.from_numpy(evs.view(np.uint8).reshape(-1).copy()).to(dev)
    d_q = torch.from_numpy(qoff).to(dev)
    d_hs = torch.from_numpy(hs.view(np.int64).copy()).to(dev)
    d_tk = torch.from_numpy(tk.view(np.int32).copy()).to(dev)
    toks, doc, m = wl.queries(0, 300)
    d_tok = torch.from_numpy(toks.reshape(-1).view(np.int32).copy()).to(dev)
    d_off = torch.arange(0, (len(toks) + 1) * wl.T, wl.T, dtype=torch.int64, device=dev)
    d_pods = torch.zeros((len(toks), 10), dtype=torch.int16, device=dev)
    d_sc = torch.zeros((len(toks), 10), dtype=torch.float64, device=dev)
    d_cnt = torch.zeros(len(toks), dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()
    ix.apply_events_dev(d_ev.data_ptr(), d_q.data_ptr(), len(qoff) - 1, len(evs), d_hs.data_ptr(), len(hs), d_tk.data_ptr())
    ix.score_batch_sparse_dev(d_tok.data_ptr(), d_off.data_ptr(), len(toks), d_pods.data_ptr(), d_sc.data_ptr(), d_cnt.data_ptr())
    ix.synchronize()
    exp = wl.expected_scores(doc, m)
    cnt, pods_h, sc_h = d_cnt.cpu().numpy(), d_pods.cpu().numpy().view(np.uint16), d_sc.cpu().numpy()
    for i in range(len(toks)):
        got = {int(pods_h[i, j]): float(sc_h[i, j]) for j in range(cnt[i])}
        assert got == {int(p): float(exp[i, p]) for p in np.nonzero(exp[i] >= 0)[0]}, i
    assert ix.stats()["request_keys"] == wl.n_blocks


def test_coop_kernel_short_parent_heads_and_block_sizes():
    """The cooperative kernel's fast path assumes a 9-byte parent head (parent >= 2^32); chains that start from a tiny
    parent (the e2e suite's ChunkHash 1, e2e_suite_test.go:125) take the in-kernel plain chain for that block.  Seeds whose
    FNV is small cannot be constructed, so the init hash is set directly."""
    rng = np.random.default_rng(12)
    for init in (1, 23, 24, 255, 65536, 2 ** 32 - 1, 2 ** 32, 0xCBF29CE484222325):
        ix = kvidx.Index(capacity=1 << 12, init_hash=init, max_pods=16)
        co = COracle(init_hash=init, max_pods=16)
        toks = rng.integers(0, 70000, size=16 * 70 + 5).astype(np.uint32)
        toks[:16] = [0, 23, 24, 255, 256, 65535, 65536, 2 ** 32 - 1, 1, 2, 3, 4, 5, 6, 7, 8]
        keys, _ = co.hash_keys(toks, [0, len(toks)])
        eng = (keys ^ np.uint64(3)).astype(np.uint64)
        assert ix.add(0, eng[:50], keys[:50], [PT(3), PT(4, 1)]) == 0 and co.add(0, eng[:50], keys[:50], [PT(3), PT(4, 1)]) == 0
        k2, _ = ix.hash_keys(toks, [0, len(toks)])
        assert np.array_equal(keys, k2)
        for n in (len(toks), 16 * 33, 16 * 32, 16, 15, 0):
            s1, h1 = ix.score_batch(toks[:n], [0, n])
            s2, h2, _, _ = co.score_batch(toks[:n], [0, n])
            assert np.array_equal(s1, s2) and np.array_equal(h1, h2), (init, n)
