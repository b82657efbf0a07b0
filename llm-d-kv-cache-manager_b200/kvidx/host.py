"""ctypes binding of the host mirror (include/kvidx_host.h): string-level Indexer / Index / Pool over libkvidx."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _native
from ._native import Config, EVENT_DTYPE

MAX_TIERS = 16
NO_KEYS = -1000


class HostConfig(C.Structure):
    _fields_ = [("index", Config), ("concurrency", C.c_uint32), ("n_tiers", C.c_uint32), ("tier_names", C.c_char_p * MAX_TIERS),
                ("tier_weights", C.c_double * MAX_TIERS), ("no_device", C.c_int32), ("enable_metrics", C.c_int32)]


LATENCY_BUCKETS = 11


class HostMetrics(C.Structure):
    _fields_ = [("admissions_total", C.c_uint64), ("evictions_total", C.c_uint64), ("lookup_requests_total", C.c_uint64),
                ("max_pod_hit_count_total", C.c_uint64), ("lookup_hits_total", C.c_uint64), ("lookup_latency_bucket", C.c_uint64 * LATENCY_BUCKETS),
                ("lookup_latency_count", C.c_uint64), ("lookup_latency_sum", C.c_double)]


_cpp = C.POINTER(C.c_char_p)
_u64p, _u32p, _u8p, _f64p = C.POINTER(C.c_uint64), C.POINTER(C.c_uint32), C.POINTER(C.c_uint8), C.POINTER(C.c_double)
SYMBOLS = {
    "kvhost_config_default": (None, [C.POINTER(HostConfig)]),
    "kvhost_create": (C.c_int, [C.POINTER(HostConfig), C.c_char_p, C.POINTER(C.c_void_p)]),
    "kvhost_destroy": (None, [C.c_void_p]),
    "kvhost_prefix_store_create": (C.c_int, [C.c_int64, C.c_int32, C.POINTER(C.c_void_p)]),
    "kvhost_prefix_store_destroy": (None, [C.c_void_p]),
    "kvhost_prefix_store_add": (C.c_int, [C.c_void_p, C.c_char_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_size_t]),
    "kvhost_prefix_store_find": (C.c_int64, [C.c_void_p, C.c_char_p, C.c_size_t, C.c_void_p, C.c_size_t, C.POINTER(C.c_double)]),
    "kvhost_prefix_store_len": (C.c_int64, [C.c_void_p]),
    "kvhost_xxhash64": (C.c_uint64, [C.c_char_p, C.c_size_t, C.c_uint64]),
    "kvhost_get_metrics": (C.c_int, [C.c_void_p, C.POINTER(HostMetrics)]),
    "kvhost_metrics_text": (C.c_int64, [C.c_void_p, C.c_char_p, C.c_size_t]),
    "kvhost_last_error": (C.c_char_p, []),
    "kvhost_index": (C.c_void_p, [C.c_void_p]),
    "kvhost_get_pod_scores": (C.c_int, [C.c_void_p, _u32p, C.c_size_t, C.c_char_p, _cpp, C.c_size_t, _cpp, _f64p]),
    "kvhost_index_add": (C.c_int, [C.c_void_p, C.c_char_p, _u64p, C.c_size_t, _u64p, C.c_size_t, _cpp, _cpp, C.c_size_t]),
    "kvhost_index_evict": (C.c_int, [C.c_void_p, C.c_char_p, C.c_uint64, _cpp, _cpp, C.c_size_t]),
    "kvhost_index_get_request_key": (C.c_int, [C.c_void_p, C.c_char_p, C.c_uint64, _u64p]),
    "kvhost_index_lookup": (C.c_int, [C.c_void_p, C.c_char_p, _u64p, C.c_size_t, _cpp, C.c_size_t, _cpp, _cpp, _u8p]),
    "kvhost_pool_add_task": (C.c_int, [C.c_void_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_size_t]),
    "kvhost_pool_queue_index": (C.c_int, [C.c_void_p, C.c_char_p]),
    "kvhost_pool_process": (C.c_int64, [C.c_void_p, C.POINTER(C.c_int64)]),
    "kvhost_decode_event_batch": (C.c_int64, [C.c_void_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_size_t, C.c_void_p, C.c_size_t, _u64p, C.c_size_t,
                                              C.POINTER(C.c_size_t), _u32p, C.c_size_t, C.POINTER(C.c_size_t)]),
    "kvhost_pod_id": (C.c_int, [C.c_void_p, C.c_char_p]),
    "kvhost_tier_id": (C.c_int, [C.c_void_p, C.c_char_p]),
    "kvhost_model_id": (C.c_int, [C.c_void_p, C.c_char_p]),
}
_bound = False


def _lib():
    global _bound
    L = _native.load()
    if not _bound:
        for name, (res, args) in SYMBOLS.items():
            f = getattr(L, name)
            f.restype, f.argtypes = res, args
        _bound = True
    return L


def _strs(items):
    arr = (C.c_char_p * max(len(items), 1))()
    for i, s in enumerate(items):
        arr[i] = s.encode() if isinstance(s, str) else s
    return arr


class HostIndexer:
    """kvcache.Indexer + kvblock.Index + kvevents.Pool with strings (pods, tiers, models are names)."""

    def __init__(self, block_size=16, hash_seed="", capacity=1 << 20, pods_per_key=10, tiers=(("gpu", 1.0), ("cpu", 0.8)), max_pods=256,
                 concurrency=4, device=0, no_device=False, lru_exact=0, enable_metrics=False):
        self.L = _lib()
        cfg = HostConfig()
        self.L.kvhost_config_default(C.byref(cfg))
        cfg.index.block_size, cfg.index.capacity, cfg.index.pods_per_key = block_size, capacity, pods_per_key
        cfg.index.max_pods, cfg.index.device, cfg.index.lru_exact = max_pods, device, lru_exact
        cfg.concurrency, cfg.n_tiers, cfg.no_device = concurrency, len(tiers), 1 if no_device else 0
        cfg.enable_metrics = 1 if enable_metrics else 0
        self._tier_names = [t[0].encode() for t in tiers]
        for i, (n, w) in enumerate(tiers):
            cfg.tier_names[i] = self._tier_names[i]
            cfg.tier_weights[i] = float(w)
        h = C.c_void_p()
        rc = self.L.kvhost_create(C.byref(cfg), hash_seed.encode(), C.byref(h))
        if rc:
            raise _native.KvidxError(rc, self.L.kvhost_last_error().decode())
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.L.kvhost_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def err(self):
        return self.L.kvhost_last_error().decode()

    def get_pod_scores(self, tokens, model, pods=()):
        """-> dict pod -> score, or None for the reference's (nil, nil)."""
        tok = np.ascontiguousarray(tokens, np.uint32)
        po = (C.c_char_p * 10)(); so = (C.c_double * 10)()
        n = self.L.kvhost_get_pod_scores(self.h, tok.ctypes.data_as(_u32p), len(tok), model.encode(), _strs(list(pods)), len(pods), po, so)
        if n == NO_KEYS:
            return None
        if n < 0:
            raise _native.KvidxError(n, self.err())
        return {po[i].decode(): so[i] for i in range(n)}

    def add(self, model, engine, request, entries):
        e = np.ascontiguousarray(engine, np.uint64); r = np.ascontiguousarray(request, np.uint64)
        return self.L.kvhost_index_add(self.h, model.encode(), e.ctypes.data_as(_u64p), len(e), r.ctypes.data_as(_u64p), len(r),
                                       _strs([p for p, _ in entries]), _strs([t for _, t in entries]), len(entries))

    def evict(self, model, engine, entries):
        return self.L.kvhost_index_evict(self.h, model.encode(), int(engine), _strs([p for p, _ in entries]), _strs([t for _, t in entries]), len(entries))

    def get_request_key(self, model, engine):
        out = C.c_uint64(0)
        rc = self.L.kvhost_index_get_request_key(self.h, model.encode(), int(engine), C.byref(out))
        return rc, out.value

    def lookup(self, model, keys, pods=()):
        k = np.ascontiguousarray(keys, np.uint64)
        n = len(k)
        po = (C.c_char_p * (10 * max(n, 1)))(); to = (C.c_char_p * (10 * max(n, 1)))(); cnt = np.zeros(max(n, 1), np.uint8)
        rc = self.L.kvhost_index_lookup(self.h, model.encode(), k.ctypes.data_as(_u64p), n, _strs(list(pods)), len(pods), po, to, cnt.ctypes.data_as(_u8p))
        if rc:
            return rc, {}
        return 0, {int(k[i]): [(po[i * 10 + j].decode(), to[i * 10 + j].decode()) for j in range(cnt[i])] for i in range(n) if cnt[i]}

    def metrics(self):
        """kvcache_index_* counters as a dict (all zero unless enable_metrics)."""
        m = HostMetrics()
        rc = self.L.kvhost_get_metrics(self.h, C.byref(m))
        if rc:
            raise _native.KvidxError(rc, self.err())
        d = {k: getattr(m, k) for k, _ in HostMetrics._fields_ if k != "lookup_latency_bucket"}
        d["lookup_latency_bucket"] = list(m.lookup_latency_bucket)
        return d

    def metrics_text(self):
        n = self.L.kvhost_metrics_text(self.h, None, 0)
        buf = C.create_string_buffer(int(n) + 1)
        self.L.kvhost_metrics_text(self.h, buf, len(buf))
        return buf.value.decode()

    def add_task(self, pod, model, payload: bytes):
        return self.L.kvhost_pool_add_task(self.h, pod.encode(), model.encode(), payload, len(payload))

    def queue_index(self, pod):
        return self.L.kvhost_pool_queue_index(self.h, pod.encode())

    def process(self):
        nd = C.c_int64(0)
        n = self.L.kvhost_pool_process(self.h, C.byref(nd))
        return n, nd.value

    def decode(self, pod, model, payload: bytes, cap=4096):
        ev = np.zeros(cap, EVENT_DTYPE); hs = np.zeros(cap * 64, np.uint64); tk = np.zeros(cap * 1024, np.uint32)
        nh, nt = C.c_size_t(0), C.c_size_t(0)
        n = self.L.kvhost_decode_event_batch(self.h, pod.encode(), model.encode(), payload, len(payload), ev.ctypes.data_as(C.c_void_p), cap,
                                             hs.ctypes.data_as(_u64p), len(hs), C.byref(nh), tk.ctypes.data_as(_u32p), len(tk), C.byref(nt))
        if n < 0:
            raise _native.KvidxError(int(n), self.err())
        return ev[:n], hs[:nh.value], tk[:nt.value]

    def pod_id(self, s):
        return self.L.kvhost_pod_id(self.h, s.encode())

    def tier_id(self, s):
        return self.L.kvhost_tier_id(self.h, s.encode("utf-8", "surrogateescape"))


def xxhash64(data: bytes, seed: int = 0) -> int:
    return int(_lib().kvhost_xxhash64(data, len(data), seed))


class PrefixStore:
    """prefixstore.LRUTokenStore (the cache in front of the tokenizer): text blocks -> the tokens that end inside them."""

    def __init__(self, cache_size=500000, block_size=256):
        self.L = _lib()
        h = C.c_void_p()
        rc = self.L.kvhost_prefix_store_create(cache_size, block_size, C.byref(h))
        if rc:
            raise _native.KvidxError(rc, self.L.kvhost_last_error().decode())
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.L.kvhost_prefix_store_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __len__(self):
        return int(self.L.kvhost_prefix_store_len(self.h))

    def add_tokenization(self, prompt: bytes, tokens, offsets):
        tok = np.ascontiguousarray(tokens, np.uint32)
        off = np.ascontiguousarray(offsets, np.uint64).reshape(-1, 2) if len(tok) else np.zeros((0, 2), np.uint64)
        assert len(off) == len(tok)
        rc = self.L.kvhost_prefix_store_add(self.h, prompt, len(prompt), tok.ctypes.data_as(C.c_void_p), off.ctypes.data_as(C.c_void_p), len(tok))
        if rc:
            raise _native.KvidxError(rc, self.L.kvhost_last_error().decode())

    def find_longest_contained_tokens(self, prompt: bytes):
        ratio = C.c_double(0.0)
        n = self.L.kvhost_prefix_store_find(self.h, prompt, len(prompt), None, 0, C.byref(ratio))      # count only (and refresh)
        out = np.zeros(max(int(n), 1), np.uint32)
        n = self.L.kvhost_prefix_store_find(self.h, prompt, len(prompt), out.ctypes.data_as(C.c_void_p), len(out), C.byref(ratio))
        return out[: int(n)].tolist(), ratio.value
