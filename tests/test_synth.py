"""CPU: the synthetic workload of SURVEY 8(d) -- the generator's closed-form scores against the oracle fed through the
write path, and bench.py's torch implementation of the query stream against the numpy one."""
import importlib.util
import os

import numpy as np
import pytest

from kvidx import synth
from oracle.kvoracle_c import COracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_closed_form_scores_equal_the_oracle():
    wl = synth.Workload(2, 1024, 1 << 13, 32)
    co = COracle(size=10 ** 6, max_pods=32)
    ev, hs, tk = wl.fill_events(0, wl.D)
    assert co.apply_events(ev, hs, tk) == (0, 0) and co.len_request() == wl.n_blocks
    toks, doc, m = wl.queries(0, 400)
    off = np.arange(0, (len(toks) + 1) * wl.T, wl.T, dtype=np.int64)
    s, has, _, _ = co.score_batch(toks.reshape(-1), off, n_threads=2)
    assert has.all() and np.array_equal(s, wl.expected_scores(doc, m))
    full, _, mf = wl.queries(0, 16, full_depth=True)
    assert (mf == wl.n).all()


def test_bench_device_query_stream_matches_numpy():
    torch = pytest.importorskip("torch")
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    wl = synth.Workload(6, 256, 1 << 12, 64)
    for q0, q1 in ((0, 37), (1000, 1100)):
        d_tok, doc, m = bench.device_queries(wl, q0, q1, torch.device("cpu"), chunk=16)
        toks, doc2, m2 = wl.queries(q0, q1)
        assert np.array_equal(d_tok.numpy().view(np.uint32).reshape(toks.shape), toks)
        assert np.array_equal(np.asarray(doc), doc2) and np.array_equal(np.asarray(m), m2)
