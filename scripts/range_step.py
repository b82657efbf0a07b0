#!/usr/bin/env python
"""One Score() step inside cudaProfilerStart/Stop, for `ncu --replay-mode range`: the step's kernels run concurrently on several
streams, so per-kernel profiles (serialised) cannot say how busy DRAM / L2 / the SMs are while it runs; a range profile can.
usage: ncu --replay-mode range --set full -o out python scripts/range_step.py [blocks] [prompts]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "llm-d-kv-cache-manager_b200")]
import torch          # noqa: E402
import kvidx          # noqa: E402
from kvidx import synth   # noqa: E402
from bench import device_queries   # noqa: E402

nblocks = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
nq = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 20
wl = synth.Workload(6, 4096, nblocks, 256)
ix = kvidx.Index(capacity=wl.n_blocks + (1 << 18), max_pods=256, device=0)
for d0 in range(0, wl.D, 2048):
    ev, hs, tk = wl.fill_events(d0, min(wl.D, d0 + 2048))
    assert ix.apply_events(ev, hs, tk) == (0, 0)
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
d_tok, doc, m = device_queries(wl, 0, nq, dev)
d_off = torch.arange(0, (nq + 1) * wl.T, wl.T, dtype=torch.int64, device=dev)
d_sc = torch.empty((nq, 256), dtype=torch.float64, device=dev)
torch.cuda.synchronize()
for _ in range(3):
    ix.score_batch_dev(d_tok.data_ptr(), d_off.data_ptr(), nq, d_sc.data_ptr())
    ix.synchronize()
torch.cuda.synchronize()
torch.cuda.profiler.start()
ix.score_batch_dev(d_tok.data_ptr(), d_off.data_ptr(), nq, d_sc.data_ptr())
ix.synchronize()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
assert np.array_equal(d_sc[:2048].cpu().numpy(), wl.expected_scores(doc[:2048], m[:2048]))
print("ok")
