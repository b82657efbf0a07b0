"""GPU, >= 2 devices: hash-range sharded index.  Two handles in one process (kvidx_shard_attach) -- the same peer-memory
mechanism the one-process-per-GPU deployment gets through CUDA IPC (kvidx_shard_export / _import)."""
import numpy as np
import pytest

import kvidx
from kvidx import dist as kd
from kvidx import synth
from oracle.kvoracle_c import COracle

pytestmark = pytest.mark.gpu


def _n_gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.skipif(_n_gpus() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("path", ["fused", "rounds", "classes"])
def test_two_shards_one_process(path, monkeypatch):
    monkeypatch.setenv("KVIDX_SCORE_PATH", path)   # fused | rounds | classes
    wl = synth.Workload(5, 1024, 1 << 14, 32)
    world = 2
    shards = [kvidx.Index(capacity=1 << 15, max_pods=32, device=r, shard_rank=r, shard_count=world) for r in range(world)]
    with pytest.raises(kvidx.KvidxError):                      # not connected yet
        shards[0].score_batch(np.zeros(16, np.uint32), [0, 16])
    for r in range(world):
        for q in range(world):
            if q != r:
                shards[r].shard_attach(q, shards[q])
    co = COracle(size=10 ** 6, max_pods=32)
    ev, hs, tk = wl.fill_events(0, wl.D)
    assert co.apply_events(ev, hs, tk) == (0, 0)
    for r in range(world):                                     # each rank ingests its pods' events
        mine = kd.events_for_rank(ev, r, world)
        assert len(mine) > 0
        rc, dropped = shards[r].apply_events(mine, hs, tk)
        assert (rc, dropped) == (0, 0)
    st = [s.stats() for s in shards]
    assert sum(s["request_keys"] for s in st) == wl.n_blocks == co.len_request()
    assert all(s["request_keys"] > wl.n_blocks // 4 for s in st)          # both shards hold a real share
    toks, doc, m = wl.queries(0, 3000)
    off = np.arange(0, (len(toks) + 1) * wl.T, wl.T, dtype=np.int64)
    exp, _, _, _ = co.score_batch(toks.reshape(-1), off, n_threads=4)
    for r in range(world):                                     # any rank can score any prompt
        lo, hi = kd.shard_range(len(toks), r, world)
        got, has = shards[r].score_batch(toks[lo:hi].reshape(-1), off[: hi - lo + 1])
        assert np.array_equal(got, exp[lo:hi]) and has.all()
    # Index API across shards: lookup / evict / get_request_key reach the owning shard from either rank
    k0, _ = shards[0].hash_keys(toks[0], [0, wl.T])
    rc, pt, cnt = shards[1].lookup(0, k0[:8])
    rc2, pt2, cnt2 = co.lookup(0, k0[:8])
    # (entry ORDER inside a slot depends on how the two ranks' pods interleaved -- unspecified in the reference too)
    assert rc == rc2 == 0 and np.array_equal(cnt, cnt2) and all(sorted(pt[i, :cnt[i]]) == sorted(pt2[i, :cnt2[i]]) for i in range(8))
    e0 = int(wl.engine_hashes(int(doc[0]), int(doc[0]) + 1)[0, 0])
    assert shards[1].get_request_key(0, e0) == co.get_request_key(0, e0)
