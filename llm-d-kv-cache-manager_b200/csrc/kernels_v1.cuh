// kernels_v1.cuh -- first-correct kernels: one thread per prompt / key / queue, tokens read
// straight from global memory.  Kept as the in-library cross-check for the tuned kernels
// (KVIDX_SCORE_KERNEL=v1) and used for the small, latency-insensitive entry points.
#pragma once
#include "table.cuh"

namespace kvx {

// ChunkedTokenDatabase.prefixHashes step (token_processor.go:115-123) for one block.
__device__ __forceinline__ uint64_t hash_block_global(uint64_t parent, const uint32_t* __restrict__ tok, uint32_t bs) {
    Fnv f;
    f.begin_block(parent, bs);
    for (uint32_t j = 0; j < bs; ++j) f.uint32(__ldg(tok + j));
    return f.end_block();
}

// TokensToKVBlockKeys (token_processor.go:141-162) for a CSR batch: thread per prompt.
__global__ void hash_keys_kernel_v1(TableView t, const uint32_t* __restrict__ tok, const int64_t* __restrict__ tok_off,
                                    int64_t tok_base, int64_t n_prompts, const uint64_t* __restrict__ parent,
                                    const uint8_t* __restrict__ parent_valid, const int64_t* __restrict__ key_off,
                                    int64_t key_base, uint64_t* __restrict__ keys_out) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n_prompts) return;
    const int64_t b = tok_off[i] - tok_base, e = tok_off[i + 1] - tok_base;
    const int64_t nblk = (e - b) / t.block_size;
    uint64_t h = t.init_hash;
    if (parent && (!parent_valid || parent_valid[i])) h = parent[i];
    uint64_t* out = keys_out + (key_off[i] - key_base);
    for (int64_t k = 0; k < nblk; ++k) {
        h = hash_block_global(h, tok + b + k * t.block_size, t.block_size);
        out[k] = h;
    }
}

__device__ __forceinline__ bool filter_has(const uint64_t* __restrict__ frow, uint32_t pod) {
    return (frow[pod >> 6] >> (pod & 63)) & 1ull;
}
__device__ __forceinline__ const uint64_t* filter_row(const uint64_t* __restrict__ filter, int64_t i, uint32_t words) {
    if (!filter) return nullptr;
    const uint64_t* r = filter + i * words;
    uint64_t any = 0;
    for (uint32_t w = 0; w < words; ++w) any |= r[w];
    return any ? r : nullptr;      // empty set == all pods (in_memory.go:126)
}

// Index.Lookup (in_memory.go:105-146): thread per key.  (A FULL slot always holds >= 1 entry --
// Evict tombstones a slot the moment it empties -- so the present-but-empty cut of
// in_memory.go:119-122 cannot trigger; it is still honoured via cut_out for completeness.)
__global__ void lookup_kernel_v1(TableView t, uint32_t model, const uint64_t* __restrict__ keys, int64_t n,
                                 const uint64_t* __restrict__ filter, uint16_t* __restrict__ podtier_out,
                                 uint8_t* __restrict__ cnt_out, int* __restrict__ cut_out, unsigned long long stamp_base) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t* frow = filter_row(filter, 0, t.filter_words);
    SlotWords w;
    uint32_t c = 0;
    uint64_t slot = 0;
    if (req_find(t, model, keys[i], w, &slot)) {
        if (t.req_stamp) t.req_stamp[slot] = stamp_base + (unsigned long long)i;      // data.Get refreshes recency (in_memory.go:118)
        const uint32_t cnt = meta_count(w.b.w);
        if (cnt == 0) atomicMin(cut_out, (int)min(i, (int64_t)0x7fffffff));
        for (uint32_t j = 0; j < cnt; ++j) {
            const uint32_t pt = slot_ent(w, j);
            if (frow && !filter_has(frow, pt >> 4)) continue;
            podtier_out[i * kMaxEnt + c++] = (uint16_t)pt;
        }
    }
    cnt_out[i] = (uint8_t)c;
}

// Per-prompt scoring state == LongestPrefixScorer.Score's podScores/activePods
// (kvblock_scorer.go:108-151), restricted to the <= 10 pods block 0 can hold.
struct ScoreState {
    uint16_t pod[kMaxEnt];
    double sc[kMaxEnt];
    uint32_t k;        // pods present at block 0 (after the filter)
    uint32_t alive;    // bitmask over [0,k): still on the consecutive prefix

    __device__ __forceinline__ void first(const TableView& t, const SlotWords& w, const uint64_t* frow) {
        k = 0; alive = 0;
        const uint32_t cnt = meta_count(w.b.w);
        for (uint32_t j = 0; j < cnt; ++j) {
            const uint32_t pt = slot_ent(w, j);
            const uint32_t p = pt >> 4;
            if (frow && !filter_has(frow, p)) continue;
            const double wt = t.weight[pt & 15u];
            uint32_t q = 0;
            for (; q < k; ++q) if (pod[q] == p) break;
            if (q == k) { pod[k] = (uint16_t)p; sc[k] = 0.0; ++k; }
            if (wt > sc[q]) sc[q] = wt;                      // getMaxWeight starts at 0.0 (kvblock_scorer.go:90)
        }
        alive = (1u << k) - 1u;
    }
    __device__ __forceinline__ void next(const TableView& t, const SlotWords& w) {
        const uint32_t cnt = meta_count(w.b.w);
        for (uint32_t q = 0; q < k; ++q) {
            if (!((alive >> q) & 1u)) continue;
            double mx = 0.0; bool present = false;
            for (uint32_t j = 0; j < cnt; ++j) {
                const uint32_t pt = slot_ent(w, j);
                if ((pt >> 4) == pod[q]) { present = true; const double wt = t.weight[pt & 15u]; if (wt > mx) mx = wt; }
            }
            if (present) sc[q] = __dadd_rn(sc[q], mx);       // in-order f64 add (kvblock_scorer.go:143-146)
            else alive &= ~(1u << q);
        }
    }
};

// Indexer.GetPodScores steps 2-4 (indexer.go:141-163): thread per prompt.
__global__ void score_kernel_v1(TableView t, const uint32_t* __restrict__ tok, const int64_t* __restrict__ tok_off,
                                int64_t tok_base, int64_t n_prompts, const uint32_t* __restrict__ model, uint32_t model0,
                                const uint64_t* __restrict__ filter, double* __restrict__ dense_out,
                                uint16_t* __restrict__ sp_pods, double* __restrict__ sp_scores, uint8_t* __restrict__ sp_cnt,
                                uint8_t* __restrict__ has_keys, unsigned long long stamp_base = 0, unsigned long long stamp_stride = 0) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n_prompts) return;
    const int64_t b = tok_off[i] - tok_base, e = tok_off[i + 1] - tok_base;
    const int64_t nblk = (e - b) / t.block_size;
    const uint32_t mdl = model ? model[i] : model0;
    const uint64_t* frow = filter_row(filter, i, t.filter_words);
    ScoreState s; s.k = 0; s.alive = 0;
    uint64_t h = t.init_hash;
    bool walking = true;
    for (int64_t k = 0; k < nblk; ++k) {
        h = hash_block_global(h, tok + b + k * t.block_size, t.block_size);
        SlotWords w;
        uint64_t slot = 0;
        const bool hit = req_find(t, mdl, h, w, &slot);
        // exact-LRU mode: the reference's Lookup touches EVERY key of the prompt that is present, also past the end of
        // the consecutive prefix (in_memory.go:117-139), so the probe loop runs to the last block and stamps them.
        if (hit && t.req_stamp) t.req_stamp[slot] = stamp_base + (unsigned long long)i * stamp_stride + (unsigned long long)k;
        if (walking) {
            if (!hit) walking = false;
            else { if (k == 0) s.first(t, w, frow); else s.next(t, w); if (!s.alive) walking = false; }
        }
        if (!walking && !t.req_stamp) break;
    }
    if (has_keys) has_keys[i] = nblk > 0;
    if (dense_out) {
        double* row = dense_out + i * (int64_t)t.max_pods;
        for (uint32_t p = 0; p < t.max_pods; ++p) row[p] = -1.0;
        for (uint32_t q = 0; q < s.k; ++q) if (s.pod[q] < t.max_pods) row[s.pod[q]] = s.sc[q];
    }
    if (sp_cnt) {
        sp_cnt[i] = (uint8_t)s.k;
        for (uint32_t q = 0; q < s.k; ++q) { sp_pods[i * kMaxEnt + q] = s.pod[q]; sp_scores[i * kMaxEnt + q] = s.sc[q]; }
    }
}

}  // namespace kvx
