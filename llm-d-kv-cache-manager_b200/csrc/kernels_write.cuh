// kernels_write.cuh -- the kvevents.Pool -> Index.Add / Index.Evict write path on the device.
//
// Concurrency model: the reference shards messages by FNV-32a(pod) so that one pod's events are
// applied in arrival order and pods are unordered with respect to each other
// (pkg/kvcache/kvevents/pool.go:129-144).  Here ONE WARP owns ONE pod's queue and walks it in
// order; warps of different pods run concurrently and meet only on slot locks.  Within an event
// the lanes of the warp take the event's blocks in parallel (they are distinct keys), after all
// lanes have redundantly evaluated the (inherently serial) hash chain.
#pragma once
#include "kernels_v1.cuh"
#include "../../include/kvidx.h"

namespace kvx {

// PodCache.Add under podCache.mu (in_memory.go:199-203) with golang-lru semantics: an existing
// entry is refreshed to newest, a new one is appended and the oldest dropped beyond the cap.
__device__ __forceinline__ uint32_t slot_add_entry(ReqSlot* s, uint32_t count, uint16_t pt, uint32_t cap) {
    volatile uint16_t* e = s->ent;
    int pos = -1;
    for (uint32_t j = 0; j < count; ++j) if (e[j] == pt) pos = (int)j;
    if (pos >= 0) {
        for (uint32_t j = (uint32_t)pos; j + 1 < count; ++j) e[j] = e[j + 1];
        e[count - 1] = pt;
        return count;
    }
    if (count >= cap) {
        for (uint32_t j = 0; j + 1 < count; ++j) e[j] = e[j + 1];
        --count;
    }
    e[count] = pt;
    return count + 1;
}

// PodCache.Remove (in_memory.go:233-235): exact (pod,tier) match only.
__device__ __forceinline__ uint32_t slot_remove_entry(ReqSlot* s, uint32_t count, uint16_t pt) {
    volatile uint16_t* e = s->ent;
    for (uint32_t j = 0; j < count; ++j) {
        if (e[j] == pt) {
            for (uint32_t q = j; q + 1 < count; ++q) e[q] = e[q + 1];
            e[count - 1] = 0;      // vacated position is zeroed: slots with equal live entries are bitwise equal
            return count - 1;
        }
    }
    return count;
}

// One (engineKey, requestKey) pair of Index.Add (in_memory.go:159-206).
__device__ __forceinline__ void do_add(const TableView& t, uint32_t model, uint64_t ehash, uint64_t rhash,
                                       const uint16_t* __restrict__ pts, int m, unsigned long long stamp = 0) {
    bool created;
    // 1. engineToRequestKeys.Add(engineKey, requestKey)   (in_memory.go:163)
    EngSlot* es = eng_lock(t, model, ehash, false, &created);
    *(volatile uint64_t*)&es->rhash = rhash;
    if (t.req_stamp) *(volatile uint64_t*)&es->stamp = stamp;                         // lru Add: insert or refresh
    if (created) atomicAdd_system(&cnt_of(t, ehash, model)->eng_full, 1ull);
    eng_unlock(es, make_meta(kStateFull, 0, model));
    // 2. get-or-create the PodCache and add the entries    (in_memory.go:170-203)
    ReqSlot* rs = req_lock(t, model, rhash, false, &created);
    uint32_t count = created ? 0u : meta_count(ld_volatile_u32(&rs->meta));
    if (t.req_stamp) t.req_stamp[rs - t.req] = stamp + 1;                             // data.Get refresh / ContainsOrAdd insert
    if (created) atomicAdd_system(&cnt_of(t, rhash, model)->req_full, 1ull);
    for (int j = 0; j < m; ++j) count = slot_add_entry(rs, count, pts[j], t.pods_per_key);
    req_unlock(rs, make_meta(kStateFull, count, model));
}

// Index.Evict (in_memory.go:212-260).
__device__ __forceinline__ void do_evict(const TableView& t, uint32_t model, uint64_t ehash,
                                         const uint16_t* __restrict__ pts, int m, unsigned long long stamp = 0) {
    uint64_t rhash;
    if (!eng_find(t, model, ehash, &rhash, nullptr, stamp)) return;      // :219-223 silent no-op (a hit refreshes recency)
    bool created;
    ReqSlot* rs = req_lock(t, model, rhash, true, &created);
    bool drop_engine = false;
    if (!rs) {
        drop_engine = true;                                              // :225-230 stale engine mapping
    } else {
        uint32_t count = meta_count(ld_volatile_u32(&rs->meta));
        if (t.req_stamp) t.req_stamp[rs - t.req] = stamp + 1;                         // data.Get (:225)
        for (int j = 0; j < m; ++j) count = slot_remove_entry(rs, count, pts[j]);
        if (count == 0) {                                                // :243-256 last entry gone
            Counters* c = cnt_of(t, rhash, model);
            atomicAdd_system(&c->req_tomb, 1ull);
            atomicAdd_system(&c->req_full, ~0ull);
            req_unlock(rs, make_meta(kStateTomb, 0, model));
            drop_engine = true;
        } else {
            req_unlock(rs, make_meta(kStateFull, count, model));
        }
    }
    if (drop_engine) {
        EngSlot* e2 = eng_lock(t, model, ehash, true, &created);
        if (e2) {
            Counters* c = cnt_of(t, ehash, model);
            atomicAdd_system(&c->eng_tomb, 1ull);
            atomicAdd_system(&c->eng_full, ~0ull);
            eng_unlock(e2, make_meta(kStateTomb, 0, model));
        }
    }
}

// Index.Add for one call: thread per key pair.
__global__ void add_kernel(TableView t, uint32_t model, const uint64_t* __restrict__ engine,
                           const uint64_t* __restrict__ request, int64_t n, const uint16_t* __restrict__ pts, int m,
                           unsigned long long stamp_base) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) do_add(t, model, engine[i], request[i], pts, m, stamp_base + 2ull * (unsigned long long)i);
}

__global__ void evict_kernel(TableView t, uint32_t model, uint64_t engine, const uint16_t* __restrict__ pts, int m,
                             unsigned long long stamp_base) {
    if (blockIdx.x == 0 && threadIdx.x == 0) do_evict(t, model, engine, pts, m, stamp_base);
}

__global__ void get_request_key_kernel(TableView t, uint32_t model, uint64_t engine, uint64_t* out, int* found, unsigned long long stamp) {
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        uint64_t r = 0;
        const bool ok = eng_find(t, model, engine, &r, nullptr, stamp);
        *out = r; *found = ok ? 1 : 0;
    }
}

// Pool.digestEvents (kvevents/pool.go:246-338): one warp per pod queue.
//   ev        events stably sorted by pod; queue q owns ev[queue_off[q] .. queue_off[q+1])
__global__ void apply_events_kernel(TableView t, const kvidx_event_t* __restrict__ ev, const int64_t* __restrict__ queue_off,
                                    int64_t n_queues, const uint64_t* __restrict__ hashes, const uint32_t* __restrict__ tokens,
                                    unsigned long long stamp_base) {
    const int lane = threadIdx.x & 31;
    const int64_t q = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (q >= n_queues) return;
    const uint32_t B = t.block_size;
    for (int64_t e = queue_off[q]; e < queue_off[q + 1]; ++e) {
        const kvidx_event_t evt = ev[e];
        const uint16_t pt = evt.podtier;
        // exact-LRU mode: event e owns the stamp range [sb, sb + 2^20): parent lookup first, then two stamps per block
        const unsigned long long sb = stamp_base + ((unsigned long long)e << 20);
        if (evt.op == KVIDX_EV_BLOCK_STORED) {
            // parent request key: GetRequestKey(parent engine key); a miss restarts at the seed (pool.go:279-294)
            uint64_t h = t.init_hash;
            if (evt.has_parent) { uint64_t r; if (eng_find(t, evt.model, evt.parent_hash, &r, nullptr, lane == 0 ? sb + 1 : 0)) h = r; }
            if (evt.n_hashes == 0) continue;                                 // pool.go:299 `if len(engineKeys) > 0`
            const uint32_t nblk = evt.n_tokens / B;
            if (nblk == 0 || nblk != evt.n_hashes) {                         // in_memory.go:150-155 -> event dropped
                if (lane == 0) atomicAdd(&t.cnt->dropped_events, 1ull);
                continue;
            }
            const uint32_t* tk = tokens + evt.tok_off;
            const uint64_t* eh = hashes + evt.hash_off;
            for (uint32_t base = 0; base < nblk; base += 32) {
                uint64_t mine = 0;
                const uint32_t lim = min(32u, nblk - base);
                for (uint32_t j = 0; j < lim; ++j) {                         // every lane walks the chain
                    h = hash_block_global(h, tk + (size_t)(base + j) * B, B);
                    if ((uint32_t)lane == j) mine = h;
                }
                if ((uint32_t)lane < lim) do_add(t, evt.model, eh[base + lane], mine, &pt, 1, sb + 2 + 2ull * (base + lane));
                __syncwarp();
            }
        } else if (evt.op == KVIDX_EV_BLOCK_REMOVED) {
            const uint64_t* eh = hashes + evt.hash_off;
            for (uint32_t base = 0; base < evt.n_hashes; base += 32) {
                if (base + lane < evt.n_hashes) do_evict(t, evt.model, eh[base + lane], &pt, 1, sb + 2 + 2ull * (base + lane));
                __syncwarp();
            }
        }
        __syncwarp();
    }
}

// ---- exact key-LRU (InMemoryIndexConfig.Size, in_memory.go:59,64): evict the least recently used key -----------
// golang-lru evicts the back of its list when Len() exceeds the size.  Here recency is a stamp per slot; the oldest
// live slot is found with a grid-wide atomicMin and tombstoned.  Evicting a request key does NOT touch the engine map
// (the stale mapping is cleaned up lazily by Evict, in_memory.go:225-230), exactly as in the reference.
__global__ void lru_min_req_kernel(const ReqSlot* __restrict__ tab, const unsigned long long* __restrict__ stamp, uint64_t slots,
                                   unsigned long long* out_min) {
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i < slots && meta_state(tab[i].meta) == kStateFull) atomicMin(out_min, stamp[i]);
}
__global__ void lru_drop_req_kernel(ReqSlot* tab, const unsigned long long* __restrict__ stamp, uint64_t slots,
                                    const unsigned long long* __restrict__ victim, Counters* cnt) {
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i < slots && meta_state(tab[i].meta) == kStateFull && stamp[i] == *victim) {
        for (int j = 0; j < kMaxEnt; ++j) tab[i].ent[j] = 0;
        tab[i].meta = make_meta(kStateTomb, 0, meta_model(tab[i].meta));
        atomicAdd(&cnt->req_full, ~0ull); atomicAdd(&cnt->req_tomb, 1ull);
    }
}
__global__ void lru_min_eng_kernel(const EngSlot* __restrict__ tab, uint64_t slots, unsigned long long* out_min) {
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i < slots && meta_state(tab[i].meta) == kStateFull) atomicMin(out_min, (unsigned long long)tab[i].stamp);
}
__global__ void lru_drop_eng_kernel(EngSlot* tab, uint64_t slots, const unsigned long long* __restrict__ victim, Counters* cnt) {
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i < slots && meta_state(tab[i].meta) == kStateFull && tab[i].stamp == *victim) {
        tab[i].meta = make_meta(kStateTomb, 0, meta_model(tab[i].meta));
        atomicAdd(&cnt->eng_full, ~0ull); atomicAdd(&cnt->eng_tomb, 1ull);
    }
}

// Re-insert every FULL slot of an old table into a fresh one (drops tombstones).
__global__ void rebuild_req_kernel(const ReqSlot* __restrict__ old_tab, uint64_t old_slots, ReqSlot* new_tab, uint64_t new_mask,
                                   const unsigned long long* __restrict__ old_stamp, unsigned long long* new_stamp) {
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= old_slots) return;
    const uint4* src = reinterpret_cast<const uint4*>(old_tab + i);
    const uint4 a = src[0], b = src[1];
    if (meta_state(b.w) != kStateFull) return;
    const uint64_t tag = ((uint64_t)a.y << 32) | a.x;
    uint64_t j = slot_home(tag, meta_model(b.w), new_mask);
    for (;;) {
        if (atomicCAS(&new_tab[j].meta, 0u, b.w | kLockBit) == 0u) {
            uint4* dst = reinterpret_cast<uint4*>(new_tab + j);
            dst[0] = a;
            uint32_t* d1 = reinterpret_cast<uint32_t*>(dst + 1);
            d1[0] = b.x; d1[1] = b.y; d1[2] = b.z;
            if (new_stamp) new_stamp[j] = old_stamp[i];
            __threadfence();
            *(volatile uint32_t*)&new_tab[j].meta = b.w & ~kLockBit;
            return;
        }
        j = (j + 1) & new_mask;
    }
}
__global__ void rebuild_eng_kernel(const EngSlot* __restrict__ old_tab, uint64_t old_slots, EngSlot* new_tab, uint64_t new_mask) {
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= old_slots) return;
    const EngSlot s = old_tab[i];
    if (meta_state(s.meta) != kStateFull) return;
    uint64_t j = slot_home(s.ehash, meta_model(s.meta), new_mask);
    for (;;) {
        if (atomicCAS(&new_tab[j].meta, 0u, s.meta | kLockBit) == 0u) {
            new_tab[j].ehash = s.ehash; new_tab[j].rhash = s.rhash; new_tab[j].stamp = s.stamp;
            __threadfence();
            *(volatile uint32_t*)&new_tab[j].meta = s.meta & ~kLockBit;
            return;
        }
        j = (j + 1) & new_mask;
    }
}

}  // namespace kvx
