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
    uint32_t* req_stamp;       // per-request-slot recency stamp, exact-LRU mode only (else nullptr)
    uint64_t req_mask, eng_mask;
    uint64_t capacity;
    uint64_t init_hash;
    uint32_t block_size;
    uint32_t pods_per_key;
    uint32_t max_pods;
    uint32_t filter_words;
    Counters* cnt;
    double weight[16];
};

#ifdef __CUDACC__

__device__ __forceinline__ uint32_t ld_volatile_u32(const uint32_t* p) { return *(const volatile uint32_t*)p; }

// ---- read-only probes (kernels that never run concurrently with writers: every call on a
//      handle is ordered on one stream) ----------------------------------------------------

struct SlotWords { uint4 a, b; };   // a = {tag.lo, tag.hi, ent0|1, ent2|3}; b = {ent4|5, ent6|7, ent8|9, meta}

__device__ __forceinline__ SlotWords load_slot(const ReqSlot* s) {
    SlotWords w;
    const uint4* p = reinterpret_cast<const uint4*>(s);
    w.a = __ldg(p);
    w.b = __ldg(p + 1);
    return w;
}
__device__ __forceinline__ uint32_t slot_ent(const SlotWords& w, int j) {
    // j in [0,10): entries packed two per word starting at a.z
    const uint32_t word = j < 2 ? w.a.z : j < 4 ? w.a.w : j < 6 ? w.b.x : j < 8 ? w.b.y : w.b.z;
    return (j & 1) ? (word >> 16) : (word & 0xffffu);
}

// Finds the FULL slot holding (model, tag).  Returns false at the first EMPTY slot.
__device__ __forceinline__ bool req_find(const TableView& t, uint32_t model, uint64_t tag, SlotWords& w, uint64_t* slot_out = nullptr) {
    uint64_t i = slot_home(tag, model, t.req_mask);
    const uint32_t tlo = (uint32_t)tag, thi = (uint32_t)(tag >> 32);
    for (;;) {
        w = load_slot(t.req + i);
        const uint32_t st = meta_state(w.b.w);
        if (st == kStateEmpty) return false;
        if (st == kStateFull && w.a.x == tlo && w.a.y == thi && meta_model(w.b.w) == model) {
            if (slot_out) *slot_out = i;
            return true;
        }
        i = (i + 1) & t.req_mask;
    }
}

__device__ __forceinline__ bool eng_find(const TableView& t, uint32_t model, uint64_t ehash, uint64_t* rhash, uint64_t* slot_out = nullptr) {
    uint64_t i = slot_home(ehash, model, t.eng_mask);
    for (;;) {
        const EngSlot* s = t.eng + i;
        const uint32_t m = ld_volatile_u32(&s->meta);
        const uint32_t st = meta_state(m);
        if (st == kStateEmpty) return false;
        if (st == kStateFull && !(m & kLockBit) && meta_model(m) == model && *(const volatile uint64_t*)&s->ehash == ehash) {
            *rhash = *(const volatile uint64_t*)&s->rhash;
            if (slot_out) *slot_out = i;
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

// Returns the slot index of (model, tag) with the lock held; *created tells whether the slot
// was claimed fresh (count 0, tag written).  If must_exist and the key is absent returns ~0.
__device__ __forceinline__ uint64_t req_lock(const TableView& t, uint32_t model, uint64_t tag, bool must_exist, bool* created) {
    uint64_t i = slot_home(tag, model, t.req_mask);
    *created = false;
    for (;;) {
        ReqSlot* s = t.req + i;
        const uint32_t m = ld_volatile_u32(&s->meta);
        if (m & kLockBit) continue;                                    // spin on this slot
        const uint32_t st = meta_state(m);
        if (st == kStateEmpty) {
            if (must_exist) return ~0ull;
            const uint32_t want = make_meta(kStateFull, 0, model) | kLockBit;
            if (atomicCAS(&s->meta, m, want) == m) {
                *(volatile uint64_t*)&s->tag = tag;
                __threadfence();
                *created = true;
                return i;
            }
            continue;                                                  // lost the race: re-examine
        }
        if (st == kStateFull && meta_model(m) == model) {
            __threadfence();
            if (*(const volatile uint64_t*)&s->tag == tag) {
                if (atomicCAS(&s->meta, m, m | kLockBit) == m) { __threadfence(); return i; }
                continue;
            }
        }
        i = (i + 1) & t.req_mask;
    }
}
__device__ __forceinline__ void req_unlock(ReqSlot* s, uint32_t new_meta) {
    __threadfence();
    *(volatile uint32_t*)&s->meta = new_meta & ~kLockBit;
}

__device__ __forceinline__ uint64_t eng_lock(const TableView& t, uint32_t model, uint64_t ehash, bool must_exist, bool* created) {
    uint64_t i = slot_home(ehash, model, t.eng_mask);
    *created = false;
    for (;;) {
        EngSlot* s = t.eng + i;
        const uint32_t m = ld_volatile_u32(&s->meta);
        if (m & kLockBit) continue;
        const uint32_t st = meta_state(m);
        if (st == kStateEmpty) {
            if (must_exist) return ~0ull;
            const uint32_t want = make_meta(kStateFull, 0, model) | kLockBit;
            if (atomicCAS(&s->meta, m, want) == m) {
                *(volatile uint64_t*)&s->ehash = ehash;
                __threadfence();
                *created = true;
                return i;
            }
            continue;
        }
        if (st == kStateFull && meta_model(m) == model) {
            __threadfence();
            if (*(const volatile uint64_t*)&s->ehash == ehash) {
                if (atomicCAS(&s->meta, m, m | kLockBit) == m) { __threadfence(); return i; }
                continue;
            }
        }
        i = (i + 1) & t.eng_mask;
    }
}
__device__ __forceinline__ void eng_unlock(EngSlot* s, uint32_t new_meta) {
    __threadfence();
    *(volatile uint32_t*)&s->meta = new_meta & ~kLockBit;
}

#endif  // __CUDACC__
}  // namespace kvx
