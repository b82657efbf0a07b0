"""CPU, world_size 2, gloo: the replica-mode plumbing (prompt sharding, event replication, max-over-ranks timing).
The index on each rank is the C++ oracle here (no GPU in this container); on the GPU box the same helpers carry
libkvidx handles (bench.py)."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from kvidx import dist as kd
from kvidx import synth


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle.kvoracle_c import COracle
        wl = synth.Workload(3, 512, 1 << 11, 16)
        co = COracle(size=10 ** 6, max_pods=16)
        # rank 0 ingests the events, every rank applies the replicated batch
        ev, hs, tk = wl.fill_events(0, wl.D) if rank == 0 else (None, None, None)
        ev, hs, tk = kd.broadcast_event_batch(ev, hs, tk, src=0)
        assert co.apply_events(ev, hs, tk) == (0, 0)
        assert co.len_request() == wl.n_blocks
        # prompts are sharded; each rank scores its slice; the gathered result equals the single-rank result
        n = 203
        toks, doc, m = wl.queries(0, n)
        lo, hi = kd.shard_range(n, rank, world)
        off = np.arange(0, (hi - lo + 1) * wl.T, wl.T, dtype=np.int64)
        local, _, _, _ = co.score_batch(toks[lo:hi].reshape(-1), off)
        full = kd.gather_scores(local, n)
        assert np.array_equal(full, wl.expected_scores(doc, m))
        # timing reduction: max over ranks
        assert kd.max_over_ranks(10.0 + rank) == 10.0 + world - 1
        q.put((rank, lo, hi, True))
    except Exception as e:      # noqa: BLE001
        q.put((rank, -1, -1, repr(e)))
    finally:
        dist.destroy_process_group()


def _worker_sharded(rank, world, port, q):
    """One process per rank, gloo for the handle exchange, CUDA IPC for the shards (both ranks may share one GPU)."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import torch
        import kvidx
        from oracle.kvoracle_c import COracle
        dev = rank % torch.cuda.device_count()
        wl = synth.Workload(3, 512, 1 << 12, 16)
        ix = kvidx.Index(capacity=1 << 13, max_pods=16, device=dev, shard_rank=rank, shard_count=world)
        kd.connect_shards(ix)                                   # all_gather of the 192-byte blobs + cudaIpcOpenMemHandle
        ev, hs, tk = wl.fill_events(0, wl.D)
        mine = kd.events_for_rank(ev, rank, world)
        assert ix.apply_events(mine, hs, tk) == (0, 0)
        dist.barrier()
        tot = torch.tensor([ix.stats()["request_keys"]], dtype=torch.int64)
        dist.all_reduce(tot)
        assert int(tot.item()) == wl.n_blocks
        co = COracle(size=10 ** 6, max_pods=16)
        assert co.apply_events(ev, hs, tk) == (0, 0)
        n = 301
        toks, doc, m = wl.queries(0, n)
        lo, hi = kd.shard_range(n, rank, world)
        off = np.arange(0, (hi - lo + 1) * wl.T, wl.T, dtype=np.int64)
        local, _ = ix.score_batch(toks[lo:hi].reshape(-1), off)
        exp, _, _, _ = co.score_batch(toks[lo:hi].reshape(-1), off)
        assert np.array_equal(local, exp)
        full = kd.gather_scores(local, n)
        assert np.array_equal(full, wl.expected_scores(doc, m))
        # the ROUTED form of the same Score() (keys to their owners, slot images back: two all-to-alls) gives the same bits
        tdev = torch.device("cuda", dev)
        torch.cuda.set_device(dev)
        st = torch.cuda.Stream(device=tdev)                     # a real stream: the library and the torch ops of the routed
        with torch.cuda.stream(st):                             # form must share one (handle 0 means "the library's own")
            d_tok = torch.from_numpy(toks[lo:hi].reshape(-1).view(np.int32).copy()).to(tdev)
            d_off = torch.from_numpy(off).to(tdev)
            d_sc = torch.empty((hi - lo, 16), dtype=torch.float64, device=tdev)
            ix.set_stream(st.cuda_stream)
            vol = kd.score_alltoall(ix, d_tok, d_off, hi - lo, d_sc)
        torch.cuda.synchronize()
        ix.set_stream(0)
        assert np.array_equal(d_sc.cpu().numpy(), exp) and vol["keys"] == (hi - lo) * wl.n and vol["bytes_back"] == 4 * vol["bytes_out"] > 0
        # churn (every rank removes all its pods' blocks of the first half of the documents), the collective compaction,
        # and the index still answers like the oracle from both ranks
        half = ev[ev["hash_off"] < np.uint64((wl.D // 2) * wl.n)].copy()
        half["op"] = 1; half["has_parent"] = 0; half["n_tokens"] = 0
        assert ix.apply_events(kd.events_for_rank(half, rank, world), hs, tk)[0] == 0
        assert co.apply_events(half, hs, tk)[0] == 0
        kd.compact_shards(ix)
        assert ix.stats()["request_tombs"] == 0
        local2, _ = ix.score_batch(toks[lo:hi].reshape(-1), off)
        exp2, _, _, _ = co.score_batch(toks[lo:hi].reshape(-1), off)
        assert np.array_equal(local2, exp2) and not np.array_equal(exp2, exp)
        q.put((rank, lo, hi, True))
    except Exception as e:      # noqa: BLE001
        import sys
        import traceback
        tb = traceback.format_exc()
        print("rank %d failed:\n%s" % (rank, tb), file=sys.stderr, flush=True)
        q.put((rank, -1, -1, tb[-700:]))
    finally:
        dist.destroy_process_group()


@pytest.mark.gpu
def test_sharded_mode_two_ranks_cuda_ipc():
    """The deployment shape of the sharded index: one process per rank, shards exchanged as CUDA IPC handles (here over
    gloo; bench.py does the same over NCCL).  Runs on a one-GPU box too -- both ranks then keep their shard on device 0."""
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_sharded, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert all(r[3] is True for r in res), res


def test_shard_range_partitions():
    for n in (0, 1, 7, 64, 203, 1000):
        for world in (1, 2, 3, 8):
            spans = [kd.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_replica_mode_two_ranks_gloo():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert all(r[3] is True for r in res), res
    spans = sorted((r[1], r[2]) for r in res)
    assert spans == [(0, 102), (102, 203)]
