"""ctypes binding of libkvidx.so (include/kvidx.h).

This is the only way Python reaches the product: there is NO Python or CPU
implementation of the path behind it.  If the library is missing or no CUDA
device is present, loading / ``Index()`` raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_PKG = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.environ.get("KVIDX_LIB") or os.path.join(_PKG, "lib", "libkvidx.so")   # KVIDX_LIB: experiment builds only

E = 10            # KVIDX_MAX_PODS_PER_KEY
MAX_TIERS = 16
SCORE_ABSENT = -1.0
OK, ENOENT, ECUDA, ENOMEM, EINVAL, ENOSPC, ERANGE = 0, -2, -5, -12, -22, -28, -34
EV_BLOCK_STORED, EV_BLOCK_REMOVED = 0, 1


class Config(C.Structure):
    """kvidx_config_t."""
    _fields_ = [("struct_size", C.c_uint32), ("device", C.c_int32), ("block_size", C.c_uint32),
                ("pods_per_key", C.c_uint32), ("init_hash", C.c_uint64), ("capacity", C.c_uint64),
                ("table_slots", C.c_uint64), ("max_pods", C.c_uint32), ("n_tier_weights", C.c_uint32),
                ("tier_weight", C.c_double * MAX_TIERS), ("lru_exact", C.c_uint32), ("shard_rank", C.c_uint32),
                ("shard_count", C.c_uint32), ("reserved", C.c_uint32 * 5)]


class Stats(C.Structure):
    """kvidx_stats_t."""
    _fields_ = [(n, C.c_uint64) for n in ("request_keys", "engine_keys", "request_tombs", "engine_tombs",
                                          "request_slots", "engine_slots", "rebuilds", "kernel_launches",
                                          "rehashed_events", "coalesced_calls")]


EVENT_DTYPE = np.dtype([("op", "u1"), ("has_parent", "u1"), ("podtier", "<u2"), ("model", "<u4"),
                        ("parent_hash", "<u8"), ("hash_off", "<u8"), ("tok_off", "<u8"),
                        ("n_hashes", "<u4"), ("n_tokens", "<u4")])
assert EVENT_DTYPE.itemsize == 40

_u32p, _u64p, _i64p = C.POINTER(C.c_uint32), C.POINTER(C.c_uint64), C.POINTER(C.c_int64)
_u16p, _u8p, _f64p = C.POINTER(C.c_uint16), C.POINTER(C.c_uint8), C.POINTER(C.c_double)

# every symbol include/kvidx.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "kvidx_abi_version": (C.c_int, []),
    "kvidx_config_default": (None, [C.POINTER(Config)]),
    "kvidx_create": (C.c_int, [C.POINTER(Config), C.POINTER(C.c_void_p)]),
    "kvidx_destroy": (None, [C.c_void_p]),
    "kvidx_last_error": (C.c_char_p, [C.c_void_p]),
    "kvidx_fnv64a": (C.c_uint64, [C.c_char_p, C.c_size_t]),
    "kvidx_fnv32a": (C.c_uint32, [C.c_char_p, C.c_size_t]),
    "kvidx_queue_index": (C.c_uint32, [C.c_char_p, C.c_size_t, C.c_uint32]),
    "kvidx_set_tier_weight": (C.c_int, [C.c_void_p, C.c_uint32, C.c_double]),
    "kvidx_host_alloc": (C.c_void_p, [C.c_size_t]),
    "kvidx_host_free": (None, [C.c_void_p]),
    "kvidx_hash_keys": (C.c_int, [C.c_void_p, _u32p, _i64p, C.c_int64, _u64p, _u8p, _u64p, _i64p]),
    "kvidx_lookup": (C.c_int, [C.c_void_p, C.c_uint32, _u64p, C.c_int64, _u64p, _u16p, _u8p]),
    "kvidx_score_batch": (C.c_int, [C.c_void_p, _u32p, _i64p, C.c_int64, _u32p, C.c_uint32, _u64p, _f64p, _u8p]),
    "kvidx_score_batch_sparse": (C.c_int, [C.c_void_p, _u32p, _i64p, C.c_int64, _u32p, C.c_uint32, _u64p, _u16p, _f64p, _u8p, _u8p]),
    "kvidx_add": (C.c_int, [C.c_void_p, C.c_uint32, _u64p, _u64p, C.c_int64, _u16p, C.c_int32]),
    "kvidx_evict": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint64, _u16p, C.c_int32]),
    "kvidx_get_request_key": (C.c_int, [C.c_void_p, C.c_uint32, C.c_uint64, _u64p]),
    "kvidx_apply_events": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, _u64p, C.c_int64, _u32p, C.c_int64, _i64p]),
    "kvidx_set_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
    "kvidx_synchronize": (C.c_int, [C.c_void_p]),
    "kvidx_score_batch_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "kvidx_hash_keys_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "kvidx_apply_events_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "kvidx_score_batch_sparse_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "kvidx_shard_compact": (C.c_int, [C.c_void_p]),
    "kvidx_key_owners_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_int64, C.c_void_p]),
    "kvidx_probe_slots_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_int64, C.c_void_p]),
    "kvidx_score_slots_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "kvidx_get_stats": (C.c_int, [C.c_void_p, C.POINTER(Stats)]),
    "kvidx_shard_export": (C.c_int, [C.c_void_p, C.c_char_p]),
    "kvidx_shard_import": (C.c_int, [C.c_void_p, C.c_uint32, C.c_char_p]),
    "kvidx_shard_attach": (C.c_int, [C.c_void_p, C.c_uint32, C.c_void_p]),
}

_lib = None


def load():
    """dlopen libkvidx.so and type every exported symbol.  Raises if the library is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise OSError("libkvidx.so not built (%s): run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "or ./build.sh -- there is no CPU fallback" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            f = getattr(L, name)      # AttributeError if the header and the library disagree
            f.restype, f.argtypes = res, args
        _lib = L
    return _lib


class KvidxError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("kvidx error %d: %s" % (code, msg))
        self.code = code


def _p(a, t):
    return None if a is None else a.ctypes.data_as(t)


def fnv64a(data: bytes) -> int:
    return load().kvidx_fnv64a(data, len(data))


def podtier(pod: int, tier: int) -> int:
    return ((pod << 4) | (tier & 15)) & 0xFFFF


class Index:
    """One libkvidx handle (== kvblock.Index + TokenProcessor + scorer of the reference, ids not strings)."""

    def __init__(self, block_size=16, init_hash=None, hash_seed="", capacity=1 << 20, pods_per_key=10,
                 tier_weights=(1.0, 0.8), max_pods=256, table_slots=0, device=0, lru_exact=0, shard_rank=0, shard_count=0):
        self.L = load()
        cfg = Config()
        self.L.kvidx_config_default(C.byref(cfg))
        cfg.device = device
        cfg.block_size = block_size
        cfg.pods_per_key = pods_per_key
        cfg.init_hash = fnv64a(hash_seed.encode()) if init_hash is None else init_hash
        cfg.capacity = capacity
        cfg.table_slots = table_slots
        cfg.max_pods = max_pods
        cfg.n_tier_weights = len(tier_weights)
        for i, w in enumerate(tier_weights):
            cfg.tier_weight[i] = float(w)
        cfg.lru_exact = lru_exact
        cfg.shard_rank, cfg.shard_count = shard_rank, shard_count
        h = C.c_void_p()
        rc = self.L.kvidx_create(C.byref(cfg), C.byref(h))
        if rc:
            raise KvidxError(rc, self.L.kvidx_last_error(None).decode())
        self.h = h
        self.block_size, self.max_pods = block_size, max_pods
        self.filter_words = (max_pods + 63) // 64

    # -- plumbing --
    def close(self):
        if getattr(self, "h", None):
            self.L.kvidx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc:
            raise KvidxError(rc, self.L.kvidx_last_error(self.h).decode())

    def last_error(self) -> str:
        return self.L.kvidx_last_error(self.h).decode()

    def set_stream(self, cuda_stream: int):
        self._ck(self.L.kvidx_set_stream(self.h, C.c_void_p(cuda_stream)))

    def synchronize(self):
        self._ck(self.L.kvidx_synchronize(self.h))

    def set_tier_weight(self, tier, w):
        self._ck(self.L.kvidx_set_tier_weight(self.h, tier, float(w)))

    def stats(self) -> dict:
        s = Stats()
        self._ck(self.L.kvidx_get_stats(self.h, C.byref(s)))
        return {n: getattr(s, n) for n, _ in Stats._fields_}

    # -- hash-range sharding (one handle per GPU) --
    SHARD_HANDLE_BYTES = 192

    def shard_export(self) -> bytes:
        buf = C.create_string_buffer(self.SHARD_HANDLE_BYTES)
        self._ck(self.L.kvidx_shard_export(self.h, buf))
        return buf.raw

    def shard_import(self, rank: int, handle: bytes):
        self._ck(self.L.kvidx_shard_import(self.h, rank, handle))

    def shard_attach(self, rank: int, other: "Index"):
        self._ck(self.L.kvidx_shard_attach(self.h, rank, other.h))

    def shard_compact(self):
        """Drop this shard's tombstones.  Collective by contract (see kvidx.dist.compact_shards)."""
        self._ck(self.L.kvidx_shard_compact(self.h))

    # -- read path --
    def hash_keys(self, tok, tok_off, parent=None, parent_valid=None):
        tok = np.ascontiguousarray(tok, np.uint32)
        tok_off = np.ascontiguousarray(tok_off, np.int64)
        n = len(tok_off) - 1
        nk = int(((tok_off[1:] - tok_off[:-1]) // self.block_size).sum())
        keys = np.zeros(max(nk, 1), np.uint64)
        koff = np.zeros(n + 1, np.int64)
        parent = None if parent is None else np.ascontiguousarray(parent, np.uint64)
        parent_valid = None if parent_valid is None else np.ascontiguousarray(parent_valid, np.uint8)
        self._ck(self.L.kvidx_hash_keys(self.h, _p(tok, _u32p), _p(tok_off, _i64p), n, _p(parent, _u64p),
                                        _p(parent_valid, _u8p), _p(keys, _u64p), _p(koff, _i64p)))
        return keys[:nk], koff

    def lookup(self, model, keys, filter_mask=None):
        keys = np.ascontiguousarray(keys, np.uint64)
        n = len(keys)
        pt = np.zeros((max(n, 1), E), np.uint16)
        cnt = np.zeros(max(n, 1), np.uint8)
        filter_mask = None if filter_mask is None else np.ascontiguousarray(filter_mask, np.uint64)
        rc = self.L.kvidx_lookup(self.h, model, _p(keys, _u64p), n, _p(filter_mask, _u64p), _p(pt, _u16p), _p(cnt, _u8p))
        return rc, pt[:n], cnt[:n]

    def score_batch(self, tok, tok_off, model=None, model0=0, filter_mask=None, out=None):
        tok = np.ascontiguousarray(tok, np.uint32) if not isinstance(tok, _Raw) else tok
        tok_off = np.ascontiguousarray(tok_off, np.int64)
        n = len(tok_off) - 1
        scores = np.empty((n, self.max_pods), np.float64) if out is None else out
        has = np.zeros(n, np.uint8)
        model = None if model is None else np.ascontiguousarray(model, np.uint32)
        filter_mask = None if filter_mask is None else np.ascontiguousarray(filter_mask, np.uint64)
        tp = C.cast(tok.ptr, _u32p) if isinstance(tok, _Raw) else _p(tok, _u32p)
        self._ck(self.L.kvidx_score_batch(self.h, tp, _p(tok_off, _i64p), n, _p(model, _u32p), model0,
                                          _p(filter_mask, _u64p), _p(scores, _f64p), _p(has, _u8p)))
        return scores, has

    def score_batch_sparse(self, tok, tok_off, model=None, model0=0, filter_mask=None, out=None):
        """out: optional preallocated (pods (n,10) uint16, scores (n,10) float64, cnt (n,) uint8, has (n,) uint8)."""
        tok = np.ascontiguousarray(tok, np.uint32)
        tok_off = np.ascontiguousarray(tok_off, np.int64)
        n = len(tok_off) - 1
        if out is None:
            pods = np.zeros((n, E), np.uint16)
            scores = np.zeros((n, E), np.float64)
            cnt = np.zeros(n, np.uint8)
            has = np.zeros(n, np.uint8)
        else:
            pods, scores, cnt, has = out
        model = None if model is None else np.ascontiguousarray(model, np.uint32)
        filter_mask = None if filter_mask is None else np.ascontiguousarray(filter_mask, np.uint64)
        self._ck(self.L.kvidx_score_batch_sparse(self.h, _p(tok, _u32p), _p(tok_off, _i64p), n, _p(model, _u32p), model0,
                                                 _p(filter_mask, _u64p), _p(pods, _u16p), _p(scores, _f64p), _p(cnt, _u8p),
                                                 _p(has, _u8p)))
        return pods, scores, cnt, has

    def score_batch_dev(self, d_tok, d_tok_off, n, d_scores, d_model=0, model0=0, d_filter=0, d_has_keys=0):
        """All pointers are device addresses (ints).  Asynchronous on the handle's stream."""
        self._ck(self.L.kvidx_score_batch_dev(self.h, d_tok, d_tok_off, n, d_model or None, model0, d_filter or None,
                                              d_scores, d_has_keys or None))

    def score_batch_sparse_dev(self, d_tok, d_tok_off, n, d_pods, d_scores, d_cnt, d_model=0, model0=0, d_filter=0, d_has_keys=0):
        """Sparse result rows (<= 10 (pod, score) pairs per prompt) on the device; pointers are device addresses."""
        self._ck(self.L.kvidx_score_batch_sparse_dev(self.h, d_tok, d_tok_off, n, d_model or None, model0, d_filter or None,
                                                     d_pods, d_scores, d_cnt, d_has_keys or None))

    def hash_keys_dev(self, d_tok, d_tok_off, n, d_key_off, d_keys, d_parent=0, d_parent_valid=0):
        self._ck(self.L.kvidx_hash_keys_dev(self.h, d_tok, d_tok_off, n, d_parent or None, d_parent_valid or None, d_key_off, d_keys))

    def key_owners_dev(self, d_keys, n, d_owner, model0=0, d_model=0):
        self._ck(self.L.kvidx_key_owners_dev(self.h, d_keys, d_model or None, model0, n, d_owner))

    def probe_slots_dev(self, d_keys, n, d_slots, model0=0, d_model=0):
        self._ck(self.L.kvidx_probe_slots_dev(self.h, d_keys, d_model or None, model0, n, d_slots))

    def score_slots_dev(self, d_slots, d_key_off, n_prompts, d_scores, d_filter=0, d_has_keys=0):
        self._ck(self.L.kvidx_score_slots_dev(self.h, d_slots, d_key_off, n_prompts, d_filter or None, d_scores, d_has_keys or None))

    def apply_events_dev(self, d_ev_sorted, d_queue_off, n_queues, n_events, d_hashes, n_hashes, d_tokens, d_n_dropped=0):
        """Device-resident, pod-sorted event batch; asynchronous on the handle's write stream."""
        self._ck(self.L.kvidx_apply_events_dev(self.h, d_ev_sorted, d_queue_off, n_queues, n_events, d_hashes, n_hashes, d_tokens,
                                               d_n_dropped or None))

    # -- write path --
    def add(self, model, engine, request, podtiers):
        engine = np.ascontiguousarray(engine, np.uint64)
        request = np.ascontiguousarray(request, np.uint64)
        pt = np.ascontiguousarray(podtiers, np.uint16)
        if len(engine) != len(request):
            return EINVAL           # in_memory.go:153-155 (the C ABI takes one n for both arrays)
        return self.L.kvidx_add(self.h, model, _p(engine, _u64p), _p(request, _u64p), len(engine), _p(pt, _u16p), len(pt))

    def evict(self, model, engine, podtiers):
        pt = np.ascontiguousarray(podtiers, np.uint16)
        return self.L.kvidx_evict(self.h, model, int(engine), _p(pt, _u16p), len(pt))

    def get_request_key(self, model, engine):
        out = C.c_uint64(0)
        rc = self.L.kvidx_get_request_key(self.h, model, int(engine), C.byref(out))
        return rc, out.value

    def apply_events(self, events, hashes, tokens):
        events = np.ascontiguousarray(events, EVENT_DTYPE)
        hashes = np.ascontiguousarray(hashes, np.uint64)
        tokens = np.ascontiguousarray(tokens, np.uint32)
        nd = C.c_int64(0)
        rc = self.L.kvidx_apply_events(self.h, events.ctypes.data_as(C.c_void_p), len(events), _p(hashes, _u64p), len(hashes),
                                       _p(tokens, _u32p), len(tokens), C.byref(nd))
        return rc, nd.value


class _Raw:
    """A raw host pointer (e.g. pinned memory from kvidx_host_alloc) passed through untouched."""

    def __init__(self, ptr: int):
        self.ptr = ptr


def host_alloc(nbytes: int) -> int:
    p = load().kvidx_host_alloc(nbytes)
    if not p:
        raise MemoryError("kvidx_host_alloc(%d) failed" % nbytes)
    return p


def host_free(p: int):
    load().kvidx_host_free(C.c_void_p(p))


def pinned_array(shape, dtype):
    """numpy array backed by pinned host memory (freed when the array's base is collected)."""
    dtype = np.dtype(dtype)
    n = int(np.prod(shape)) * dtype.itemsize
    p = host_alloc(max(n, 1))
    buf = (C.c_char * max(n, 1)).from_address(p)
    arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
    _PINNED[arr.ctypes.data] = (p, buf)
    return arr


_PINNED = {}
