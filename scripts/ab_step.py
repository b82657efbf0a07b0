#!/usr/bin/env python
"""A/B of tuning knobs on the default Score() step: one index per configuration (the knobs are read at kvidx_create), same
queries, step time from CUDA events.  usage: python scripts/ab_step.py [blocks] [prompts] CONF...   CONF = K=V,K=V or 'base'"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "llm-d-kv-cache-manager_b200")]
import torch          # noqa: E402
import kvidx          # noqa: E402
from kvidx import synth   # noqa: E402
from bench import device_queries   # noqa: E402

nblocks, nq = int(sys.argv[1]), int(sys.argv[2])
wl = synth.Workload(6, 4096, nblocks, 256)
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
d_tok, doc, m = device_queries(wl, 0, nq, dev)
d_off = torch.arange(0, (nq + 1) * wl.T, wl.T, dtype=torch.int64, device=dev)
d_sc = torch.empty((nq, 256), dtype=torch.float64, device=dev)
exp = wl.expected_scores(doc[:4096], m[:4096])
fills = [wl.fill_events(d0, min(wl.D, d0 + 2048)) for d0 in range(0, wl.D, 2048)]
for conf in sys.argv[3:]:
    kv = dict(x.split("=", 1) for x in conf.split(",") if "=" in x)
    os.environ.update(kv)
    ix = kvidx.Index(capacity=wl.n_blocks + (1 << 18), max_pods=256, device=0)
    for ev, hs, tk in fills:
        assert ix.apply_events(ev, hs, tk) == (0, 0)
    st = torch.cuda.Stream()
    ix.set_stream(st.cuda_stream)
    for _ in range(4):
        ix.score_batch_dev(d_tok.data_ptr(), d_off.data_ptr(), nq, d_sc.data_ptr())
    st.synchronize()
    ts = []
    for _ in range(8):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        ix.score_batch_dev(d_tok.data_ptr(), d_off.data_ptr(), nq, d_sc.data_ptr())
        e1.record(st)
        e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    ok = np.array_equal(d_sc[:4096].cpu().numpy(), exp)
    print("%-60s  step ms: min %.3f  median %.3f   (%.3g prompts/s)  parity %s" % (conf, min(ts), float(np.median(ts)), nq / (np.median(ts) / 1e3), ok), flush=True)
    ix.set_stream(0)
    del ix
    for k in kv:
        os.environ.pop(k, None)
