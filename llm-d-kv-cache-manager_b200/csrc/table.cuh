// table.cuh -- HBM-resident open-addressed tables that replace the reference's two LRU maps
// (pkg/kvcache/kvblock/in_memory.go:77-84):
//   request table : Key{model, chunk hash} -> pod-entry set   (lru.Cache[Key,*PodCache])
//   engine  table : Key{model, engine hash} -> request hash    (lru.Cache[Key,Key])
//
// A request slot is exactly one 32-byte DRAM sector, so a Score() probe that hits its home
// slot costs one sector:
//   bytes  0..7   tag      chunk hash
//   bytes  8..27  ent[10]  packed (pod<<4|tier), physically ordered oldest -> newest, which is
//                          the order golang-lru's Keys() returns (in_memory.go:128) and makes the
//                          per-key pod LRU (in_memory.go:199-203, cap = PodCacheSize) a shift.
//   bytes 28..31  meta     [1:0] state  [2] lock  [7:4] count  [31:16] model id
// Linear probing, power-of-two slot count, load factor <= 0.5 by construction (slots >= 2*capacity).
// Deletion leaves a tombstone; the host rebuilds when tombstones pile up.
#pragma once
#include <stdint.h>
#include "fnv_cbor.cuh"

namespace kvx {

constexpr uint32_t kStateEmpty = 0, kStateFull = 1, kStateTomb = 2;
constexpr uint32_t kStateMask = 3u, kLockBit = 4u;
constexpr int kMaxEnt = 10;

struct __align__(32) ReqSlot {
    uint64_t tag;
    uint16_t ent[kMaxEnt];
    uint32_t meta;
};
static_assert(sizeof(ReqSlot) == 32, "request slot must be one 32-byte sector");

struct __align__(32) EngSlot {
    uint64_t ehash;
    uint64_t rhash;
    uint32_t meta;     // [1:0] state [2] lock [31:16] model
    uint32_t pad;
    uint64_t stamp;    // recency stamp (exact-LRU mode)
};
static_assert(sizeof(EngSlot) == 32, "engine slot is 32 bytes");

// Home slot: even-aligned so that probe positions 0 and 1 share one 64-byte segment and the read path can
// fetch both with a single coalesced 64-byte access (then linear probing continues at home+2).
__host__ __device__ __forceinline__ uint64_t slot_home(uint64_t hash, uint32_t model, uint64_t mask) {
    return home_of(hash, model) & mask & ~1ull;
}
__host__ __device__ __forceinline__ uint32_t shard_of(uint64_t mixed, uint32_t shard_bits) {
    return shard_bits ? (uint32_t)(mixed >> (64 - shard_bits)) : 0u;
}
__host__ __device__ __forceinline__ uint32_t meta_state(uint32_t m) { return m & kStateMask; }
__host__ __device__ __forceinline__ uint32_t meta_count(uint32_t m) { return (m >> 4) & 0xfu; }
__host__ __device__ __forceinline__ uint32_t meta_model(uint32_t m) { return m >> 16; }
__host__ __device__ __forceinline__ uint32_t make_meta(uint32_t state, uint32_t count, uint32_t model) {
    return state | (count << 4) | (model << 16);
}

struct Counters {
    unsigned long long req_full, req_tomb, eng_full, eng_tomb;
    unsigned long long dropped_events;     // BlockStored with key-count mismatch (in_memory.go:153-155)
    unsigned long long nospc;              // inserts refused at capacity
    unsigned long long clock;              // recency clock (exact-LRU mode)
    unsigned long long pad;
};

struct TableView {
    ReqSlot* req;
    EngSlot* eng;
    unsigned long long* req_stamp;   // per-request-slot recency stamp: exact-LRU mode only (else nullptr).  Engine
                                     // slots carry their stamp inline.  Larger = more recent (golang-lru front).
    uint64_t req_mask, eng_mask;
    uint64_t capacity;
    uint64_t init_hash;
    uint32_t block_size;
    uint32_t pods_per_key;
    uint32_t max_pods;
    uint32_t filter_words;
    Counters* cnt;
    double weight[16];
    // hash-range sharding over the GPUs of one NVSwitch domain: shard = top bits of the mixed key, slot = low bits.
    // *_peer[r] is shard r's table mapped into this process (CUDA IPC); entry `shard_rank` is the local one.
    // Unsharded handles have shard_bits == 0 and *_peer[0] == the local tables.
    ReqSlot* req_peer[8];
    EngSlot* eng_peer[8];
    Counters* cnt_peer[8];
    uint32_t shard_bits, shard_rank;
};

#ifdef __CUDACC__

__device__ __forceinline__ uint32_t ld_volatile_u32(const uint32_t* p) { return *(const volatile uint32_t*)p; }

// ---- read-only probes (kernels that never run concurrently with writers: every call on a
//      handle is ordered on one stream) ----------------------------------------------------

struct SlotWords { uint4 a, b; };   // a = {tag.lo, tag.hi, ent0|1, ent2|3}; b = {ent4|5, ent6|7, ent8|9, meta}

// One 32-byte slot as a single 256-bit load (sm_100: LDG.E.256).  Local tables go through the read-only path; a
// peer GPU's shard is read with a plain (weak) load -- NVLink-mapped memory is not a read-only-cache target.
__device__ __forceinline__ void ld_slot(const ReqSlot* s, bool peer, uint4& a, uint4& b) {
    if (peer)
        asm volatile("ld.global.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w) : "l"(s) : "memory");
    else
        asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w) : "l"(s));
}
__device__ __forceinline__ SlotWords load_slot(const ReqSlot* s, bool peer = false) {
    SlotWords w;
    ld_slot(s, peer, w.a, w.b);
    return w;
}
__device__ __forceinline__ uint32_t slot_ent(const SlotWords& w, int j) {
    // j in [0,10): entries packed two per word starting at a.z
    const uint32_t word = j < 2 ? w.a.z : j < 4 ? w.a.w : j < 6 ? w.b.x : j < 8 ? w.b.y : w.b.z;
    return (j & 1) ? (word >> 16) : (word & 0xffffu);
}

// Finds the FULL slot holding (model, tag).  Returns false at the first EMPTY slot.
__device__ __forceinline__ bool req_find(const TableView& t, uint32_t model, uint64_t tag, SlotWords& w, uint64_t* slot_out = nullptr) {
    const uint64_t hm = home_of(tag, model);
    const ReqSlot* base = t.req_peer[shard_of(hm, t.shard_bits)];
    uint64_t i = hm & t.req_mask & ~1ull;
    const uint32_t tlo = (uint32_t)tag, thi = (uint32_t)(tag >> 32);
    for (;;) {
        w = load_slot(base + i, t.shard_bits != 0);
        const uint32_t st = meta_state(w.b.w);
        if (st == kStateEmpty) return false;
        if (st == kStateFull && w.a.x == tlo && w.a.y == thi && meta_model(w.b.w) == model) {
            if (slot_out) *slot_out = i;
            return true;
        }
        i = (i + 1) & t.req_mask;
    }
}

__device__ __forceinline__ bool eng_find(const TableView& t, uint32_t model, uint64_t ehash, uint64_t* rhash, uint64_t* slot_out = nullptr,
                                         unsigned long long stamp = 0) {
    const uint64_t hm = home_of(ehash, model);
    const EngSlot* base = t.eng_peer[shard_of(hm, t.shard_bits)];
    uint64_t i = hm & t.eng_mask & ~1ull;
    for (;;) {
        const EngSlot* s = base + i;
        const uint32_t m = ld_volatile_u32(&s->meta);
        const uint32_t st = meta_state(m);
        if (st == kStateEmpty) return false;
        if (st == kStateFull && !(m & kLockBit) && meta_model(m) == model && *(const volatile uint64_t*)&s->ehash == ehash) {
            *rhash = *(const volatile uint64_t*)&s->rhash;
            if (slot_out) *slot_out = i;
            if (stamp && t.req_stamp) *(volatile uint64_t*)&const_cast<EngSlot*>(s)->stamp = stamp;   // lru Get refreshes recency
            return true;
        }
        if (st == kStateFull && (m & kLockBit)) continue;   // being written: re-read this slot
        i = (i + 1) & t.eng_mask;
    }
}

// ---- writer-side protocol ---------------------------------------------------------------
// All mutation of a slot happens between a successful lock CAS on its meta word and the
// releasing store.  EMPTY -> FULL|LOCK is the claim; tags are written under the lock and are
// immutable while the slot stays FULL.

// Returns the slot of (model, tag) with the lock held; *created tells whether the slot was claimed
// fresh (count 0, tag written).  If must_exist and the key is absent returns nullptr.  The slot may
// live in a peer GPU's shard: atomics and fences are system scope (NVLink carries both).
__device__ __forceinline__ ReqSlot* req_lock(const TableView& t, uint32_t model, uint64_t tag, bool must_exist, bool* created) {
    const uint64_t hm = home_of(tag, model);
    ReqSlot* base = t.req_peer[shard_of(hm, t.shard_bits)];
    uint64_t i = hm & t.req_mask & ~1ull;
    *created = false;
    for (;;) {
        ReqSlot* s = base + i;
        const uint32_t m = ld_volatile_u32(&s->meta);
        if (m & kLockBit) continue;                                    // spin on this slot
        const uint32_t st = meta_state(m);
        if (st == kStateEmpty) {
            if (must_exist) return nullptr;
            const uint32_t want = make_meta(kStateFull, 0, model) | kLockBit;
            if (atomicCAS_system(&s->meta, m, want) == m) {
                *(volatile uint64_t*)&s->tag = tag;
                __threadfence_system();
                *created = true;
                return s;
            }
            continue;                                                  // lost the race: re-examine
        }
        if (st == kStateFull && meta_model(m) == model) {
            __threadfence_system();
            if (*(const volatile uint64_t*)&s->tag == tag) {
                if (atomicCAS_system(&s->meta, m, m | kLockBit) == m) { __threadfence_system(); return s; }
                continue;
            }
        }
        i = (i + 1) & t.req_mask;
    }
}
__device__ __forceinline__ void req_unlock(ReqSlot* s, uint32_t new_meta) {
    __threadfence_system();
    *(volatile uint32_t*)&s->meta = new_meta & ~kLockBit;
}

__device__ __forceinline__ EngSlot* eng_lock(const TableView& t, uint32_t model, uint64_t ehash, bool must_exist, bool* created) {
    const uint64_t hm = home_of(ehash, model);
    EngSlot* base = t.eng_peer[shard_of(hm, t.shard_bits)];
    uint64_t i = hm & t.eng_mask & ~1ull;
    *created = false;
    for (;;) {
        EngSlot* s = base + i;
        const uint32_t m = ld_volatile_u32(&s->meta);
        if (m & kLockBit) continue;
        const uint32_t st = meta_state(m);
        if (st == kStateEmpty) {
            if (must_exist) return nullptr;
            const uint32_t want = make_meta(kStateFull, 0, model) | kLockBit;
            if (atomicCAS_system(&s->meta, m, want) == m) {
                *(volatile uint64_t*)&s->ehash = ehash;
                __threadfence_system();
                *created = true;
                return s;
            }
            continue;
        }
        if (st == kStateFull && meta_model(m) == model) {
            __threadfence_system();
            if (*(const volatile uint64_t*)&s->ehash == ehash) {
                if (atomicCAS_system(&s->meta, m, m | kLockBit) == m) { __threadfence_system(); return s; }
                continue;
            }
        }
        i = (i + 1) & t.eng_mask;
    }
}
__device__ __forceinline__ void eng_unlock(EngSlot* s, uint32_t new_meta) {
    __threadfence_system();
    *(volatile uint32_t*)&s->meta = new_meta & ~kLockBit;
}

__device__ __forceinline__ Counters* cnt_of(const TableView& t, uint64_t hash, uint32_t model) {
    return t.cnt_peer[shard_of(home_of(hash, model), t.shard_bits)];
}

#endif  // __CUDACC__
}  // namespace kvx
