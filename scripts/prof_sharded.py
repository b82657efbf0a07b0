#!/usr/bin/env python
"""Two hash-range shards on two GPUs driven by ONE process (kvidx_shard_attach), so that ncu can watch the kernels of a sharded
Score() -- NVLink bytes moved by the walk's peer probes.  usage: python scripts/prof_sharded.py [blocks] [prompts] [world]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "llm-d-kv-cache-manager_b200")]
import torch          # noqa: E402
import kvidx          # noqa: E402
from kvidx import dist as kd, synth   # noqa: E402

nblocks = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 21
nq = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 19
world = int(sys.argv[3]) if len(sys.argv) > 3 else 2
wl = synth.Workload(6, 4096, nblocks, 256)
shards = [kvidx.Index(capacity=wl.n_blocks + (1 << 18), max_pods=256, device=r, shard_rank=r, shard_count=world) for r in range(world)]
for r in range(world):
    for q in range(world):
        if q != r:
            shards[r].shard_attach(q, shards[q])
for d0 in range(0, wl.D, 2048):
    ev, hs, tk = wl.fill_events(d0, min(wl.D, d0 + 2048))
    for r in range(world):
        assert shards[r].apply_events(kd.events_for_rank(ev, r, world), hs, tk) == (0, 0)
sys.path.insert(0, ROOT)
from bench import device_queries   # noqa: E402
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
d_tok, doc, m = device_queries(wl, 0, nq, dev)
d_off = torch.arange(0, (nq + 1) * wl.T, wl.T, dtype=torch.int64, device=dev)
d_sc = torch.empty((nq, 256), dtype=torch.float64, device=dev)
torch.cuda.synchronize()
ix = shards[0]
for it in range(3):
    t0 = time.perf_counter()
    ix.score_batch_dev(d_tok.data_ptr(), d_off.data_ptr(), nq, d_sc.data_ptr())
    ix.synchronize()
    dt = time.perf_counter() - t0
exp = wl.expected_scores(doc[:2048], m[:2048])
assert np.array_equal(d_sc[:2048].cpu().numpy(), exp)
n_probe = np.minimum(wl.n, m + 1)
print("sharded step: %d prompts in %.2f ms (%.3g prompts/s on one GPU of %d shards); mean probes/prompt needed %.1f; all-to-all would move %.0f B/prompt"
      % (nq, dt * 1e3, nq / dt, world, n_probe.mean(), wl.n * 40.0 * (world - 1) / world))
