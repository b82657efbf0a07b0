"""CPU: libkvidx.so loads, exports every symbol include/kvidx.h declares, and refuses to run
without a GPU (no CPU fallback).  No compute calls."""
import ctypes as C
import os
import re

import pytest

import conftest
import kvidx
from kvidx import _native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "kvidx.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(kvidx_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_all_exported_and_bound():
    names = _declared_symbols()
    assert len(names) >= 24
    L = C.CDLL(_native.LIB_PATH)
    for n in names:
        assert hasattr(L, n), "libkvidx.so does not export %s" % n
        assert n in _native.SYMBOLS, "ctypes binding misses %s" % n
    assert sorted(_native.SYMBOLS) == names


def test_host_mirror_symbols_exported_and_bound():
    from kvidx import host
    src = open(os.path.join(ROOT, "include", "kvidx_host.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = sorted(set(re.findall(r"\b(kvhost_[a-z0-9_]+)\s*\(", src)))
    assert len(names) >= 17
    L = C.CDLL(_native.LIB_PATH)
    for n in names:
        assert hasattr(L, n), "libkvidx.so does not export %s" % n
    assert sorted(host.SYMBOLS) == names
    h = host.HostIndexer(no_device=True)                  # host-only instance works without a GPU ...
    assert h.pod_id("pod-a") == 0 and h.pod_id("pod-b") == 1 and h.pod_id("pod-a") == 0 and h.tier_id("CPU") == 1
    with pytest.raises(kvidx.KvidxError) as ei:           # ... but scoring does not silently fall back to the CPU
        h.get_pod_scores(list(range(32)), "m")
    assert ei.value.code == kvidx.ECUDA and "no CPU fallback" in str(ei.value)


def test_host_metrics_exposition_format():
    """kvhost_metrics_text: Prometheus text format with the reference's metric names (metrics/collector.go:28-59)."""
    from kvidx import host
    h = host.HostIndexer(no_device=True, enable_metrics=True)
    m = h.metrics()
    assert m["admissions_total"] == 0 and m["lookup_latency_count"] == 0 and len(m["lookup_latency_bucket"]) == 11
    text = h.metrics_text()
    for name in ("admissions_total", "evictions_total", "lookup_requests_total", "max_pod_hit_count_total", "lookup_hits_total"):
        assert "# TYPE kvcache_index_%s counter\nkvcache_index_%s 0\n" % (name, name) in text
    assert "# TYPE kvcache_index_lookup_latency_seconds histogram" in text
    les = re.findall(r'kvcache_index_lookup_latency_seconds_bucket\{le="([^"]+)"\} 0', text)
    assert les == ["0.005", "0.01", "0.025", "0.05", "0.1", "0.25", "0.5", "1", "2.5", "5", "10", "+Inf"]
    assert text.endswith("kvcache_index_lookup_latency_seconds_count 0\n")
    assert C.sizeof(host.HostMetrics) == 8 * (5 + 11 + 1) + 8


def test_abi_version_and_struct_layout():
    L = kvidx.load()
    assert L.kvidx_abi_version() == 2
    cfg = _native.Config()
    L.kvidx_config_default(C.byref(cfg))
    assert cfg.struct_size == C.sizeof(_native.Config) == 208
    assert (cfg.block_size, cfg.pods_per_key, cfg.max_pods, cfg.n_tier_weights) == (16, 10, 256, 2)
    assert cfg.init_hash == 0xCBF29CE484222325 and (cfg.tier_weight[0], cfg.tier_weight[1]) == (1.0, 0.8)
    assert _native.EVENT_DTYPE.itemsize == 40 and C.sizeof(_native.Stats) == 80


def test_host_helpers():
    assert kvidx.fnv64a(b"") == 0xCBF29CE484222325 and kvidx.fnv64a(b"42") == 571532774284038691
    L = kvidx.load()
    assert L.kvidx_fnv32a(b"a", 1) == 0xE40C292C
    assert L.kvidx_queue_index(b"pod-1", 5, 4) == L.kvidx_fnv32a(b"pod-1", 5) % 4
    assert kvidx.podtier(5, 1) == 0x51


@pytest.mark.skipif(conftest.HAS_CUDA, reason="checks the no-GPU failure mode")
def test_create_fails_loudly_without_gpu():
    with pytest.raises(kvidx.KvidxError) as ei:
        kvidx.Index()
    assert ei.value.code == kvidx.ECUDA and "no CPU fallback" in str(ei.value)
