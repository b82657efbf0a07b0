// kernels_score.cuh -- the fused Score() kernel for sm_100a.
//
// Replaces steps 2-4 of kvcache.Indexer.GetPodScores (pkg/kvcache/indexer.go:141-163):
//   TokensToKVBlockKeys  (kvblock/token_processor.go:141-162)   -> chain hash in registers
//   Index.Lookup         (kvblock/in_memory.go:105-146)         -> one 32-byte sector probe per key
//   LongestPrefixScorer  (kvblock_scorer.go:108-151)            -> in-order f64 accumulate
// Keys never touch memory: a key is produced in registers, turned into a slot address, and dropped.
//
// Shape of the kernel (pure integer / HBM work -- no tensor cores):
//   * persistent grid, one CTA-set per SM; every LANE is an independent chain worker that owns one
//     prompt at a time (the FNV chain is serial per prompt, so the only parallelism is across
//     prompts); a lane that finishes its prompt -- early, because its last active pod dropped, or
//     at the end -- pulls the next prompt index from a global counter, so early exits turn into
//     throughput instead of idle lanes;
//   * tokens are staged global->shared with per-lane 1-D TMA bulk copies (cp.async.bulk, 64 B =
//     one 16-token block, completion on a per-warp mbarrier), double buffered one block ahead; the
//     shared layout pads each lane's row to 80 B so the four LDS.128 per block are conflict free;
//   * the table probe for block b is issued (2 x LDG.128 = one sector) right after key b exists
//     and consumed after block b+1 has been hashed, which hides the DRAM latency behind ~450
//     issue slots of hashing;
//   * per-prompt score state (<= 10 pods: the block-0 slot bounds the result map) lives in shared
//     memory, transposed [entry][lane] so lanes never conflict and, at retirement, the whole warp
//     can read one lane's state and write that prompt's dense row with coalesced 16-byte stores.
#pragma once
#include <cuda_runtime.h>
#include "kernels_v1.cuh"

namespace kvx {

constexpr int kScoreWarps = 4;                 // warps per CTA
constexpr int kScoreThreads = kScoreWarps * 32;
constexpr int kScoreCtasPerSm = 5;

template <int BS> struct ScoreSmem {
    static constexpr int kRow = BS * 4 + 16;   // bytes per lane per stage (pad 16 -> conflict-free LDS.128)
    struct __align__(16) Warp {
        unsigned char tok[2][32 * kRow];
        double sc[kMaxEnt][32];
        uint16_t pod[kMaxEnt][32];
        unsigned long long bar[2];
    };
    Warp w[kScoreWarps];
    double weight[16];
};

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) { while (!mbar_try_wait(bar, parity)) {} }
// 1-D TMA bulk copy global -> shared, completion counted in bytes on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void tma_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ uint4 ld_nc_v4(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}

struct ScoreArgs {
    const uint32_t* tok; const int64_t* tok_off; int64_t tok_base; int64_t n_prompts;
    const uint32_t* model; uint32_t model0; const uint64_t* filter;
    double* dense; uint16_t* sp_pods; double* sp_scores; uint8_t* sp_cnt; uint8_t* has_keys;
    unsigned long long* next;      // global work counter (zeroed before launch)
};

template <int BS>
__global__ void __launch_bounds__(kScoreThreads, kScoreCtasPerSm)
score_kernel_tuned(const TableView t, const ScoreArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    using SM = ScoreSmem<BS>;
    SM& sm = *reinterpret_cast<SM*>(smem_raw);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    typename SM::Warp& W = sm.w[wid];
    const uint32_t bar0 = smem_addr(&W.bar[0]), bar1 = smem_addr(&W.bar[1]);
    constexpr uint32_t kBlkBytes = BS * 4;
    if (lane == 0) { mbar_init(bar0, 1); mbar_init(bar1, 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    if (threadIdx.x < 16) sm.weight[threadIdx.x] = t.weight[threadIdx.x];
    __syncthreads();

    // ---- per-lane worker state ----
    long long pi = -1;                 // prompt index, -1 = no prompt
    bool exhausted = false;            // the global counter ran past n_prompts
    const uint32_t* tokp = nullptr;    // first token of the prompt
    int nblk = 0, cblk = 0, hblk = 0;  // blocks: total, copies issued, hashed
    bool aligned = true;               // prompt start is 16-byte aligned (TMA path) else direct global loads
    bool staged_cur = false;           // a block for this lane sits in the current stage
    uint64_t h = t.init_hash;
    bool pend = false; uint64_t pkey = 0, pslot = 0; uint4 pa = {0, 0, 0, 0}, pb = {0, 0, 0, 0};
    int pblk = 0;
    uint32_t k = 0, alive = 0, mdl = a.model0;
    const uint64_t* frow = nullptr;
    bool finished = false;             // prompt result is final, waiting for the cooperative write-out

    auto refill = [&]() {
        // write out finished prompts and hand new prompts to lanes that need one
        for (;;) {
            __syncwarp();
            const uint32_t done = __ballot_sync(0xffffffffu, finished);
            uint32_t dm = done;
            while (dm) {
                const int p = __ffs(dm) - 1; dm &= dm - 1;
                const long long ppi = __shfl_sync(0xffffffffu, pi, p);
                const uint32_t pk = __shfl_sync(0xffffffffu, k, p);
                const int pn = __shfl_sync(0xffffffffu, nblk, p);
                if (a.dense) {
                    double* row = a.dense + ppi * (long long)t.max_pods;
                    const uint32_t P = t.max_pods;
                    if ((P & 1u) == 0 && ((reinterpret_cast<uintptr_t>(row) & 15u) == 0)) {
                        for (uint32_t c = lane * 2; c < P; c += 64) *reinterpret_cast<double2*>(row + c) = make_double2(-1.0, -1.0);
                    } else {
                        for (uint32_t c = lane; c < P; c += 32) row[c] = -1.0;
                    }
                    __syncwarp();
                    if ((uint32_t)lane < pk) { const uint32_t pd = W.pod[lane][p]; if (pd < P) row[pd] = W.sc[lane][p]; }
                }
                if (a.sp_cnt) {
                    if ((uint32_t)lane < pk) {
                        a.sp_pods[ppi * kMaxEnt + lane] = W.pod[lane][p];
                        a.sp_scores[ppi * kMaxEnt + lane] = W.sc[lane][p];
                    }
                    if (lane == 0) a.sp_cnt[ppi] = (uint8_t)pk;
                }
                if (a.has_keys && lane == 0) a.has_keys[ppi] = pn > 0;
            }
            __syncwarp();
            const bool want = (finished || pi < 0) && !exhausted;
            const uint32_t wm = __ballot_sync(0xffffffffu, want);
            if (finished) { finished = false; pi = -1; }
            if (!wm) break;
            unsigned long long base = 0;
            if (lane == 0) base = atomicAdd(a.next, (unsigned long long)__popc(wm));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (want) {
                const long long idx = (long long)base + __popc(wm & ((1u << lane) - 1u));
                if (idx >= a.n_prompts) { exhausted = true; pi = -1; }
                else {
                    pi = idx;
                    const int64_t b = a.tok_off[idx] - a.tok_base, e = a.tok_off[idx + 1] - a.tok_base;
                    tokp = a.tok + b;
                    nblk = (int)((e - b) / BS);
                    cblk = 0; hblk = 0; h = t.init_hash; pend = false; staged_cur = false; k = 0; alive = 0;
                    aligned = (reinterpret_cast<uintptr_t>(tokp) & 15u) == 0;
                    mdl = a.model ? a.model[idx] : a.model0;
                    frow = filter_row(a.filter, idx, t.filter_words);
                    finished = (nblk == 0);
                }
            }
            if (!__any_sync(0xffffffffu, finished)) break;
        }
    };

    // issue the staging copy of each lane's next block into stage s (bar = its mbarrier)
    auto issue_copy = [&](int s, uint32_t bar) -> bool {
        const bool issue = pi >= 0 && !finished && cblk < nblk;
        const bool tma = issue && aligned;
        const uint32_t nb = __popc(__ballot_sync(0xffffffffu, tma));
        if (lane == 0) mbar_expect_tx(bar, nb * kBlkBytes);
        __syncwarp();
        if (tma) tma_bulk_g2s(smem_addr(&W.tok[s][lane * SM::kRow]), tokp + (size_t)cblk * BS, kBlkBytes, bar);
        if (issue) ++cblk;
        return issue;
    };

    refill();
    bool staged_next = issue_copy(0, bar0);
    uint32_t it = 0;
    for (;; ++it) {
        const int sC = it & 1, sN = sC ^ 1;
        staged_cur = staged_next;
        // 1. stage the following block
        staged_next = issue_copy(sN, sN ? bar1 : bar0);
        // 2. wait for the current stage
        mbar_wait(sC ? bar1 : bar0, (it >> 1) & 1u);
        // 3. hash the current block -> request key
        uint64_t key = 0;
        if (__any_sync(0xffffffffu, staged_cur)) {
            Fnv f;
            f.begin_block(h, BS);
            if (aligned) {
                const uint4* tp = reinterpret_cast<const uint4*>(&W.tok[sC][lane * SM::kRow]);
#pragma unroll
                for (int c = 0; c < BS / 4; ++c) {
                    const uint4 v = tp[c];
                    f.uint32(v.x); f.uint32(v.y); f.uint32(v.z); f.uint32(v.w);
                }
            } else if (staged_cur) {
                const uint32_t* g = tokp + (size_t)hblk * BS;
                for (int c = 0; c < BS; ++c) f.uint32(__ldg(g + c));
            }
            key = f.end_block();
            if (staged_cur) { h = key; }
        }
        // 4. consume the probe issued one block ago
        if (pend) {
            pend = false;
            uint32_t st = meta_state(pb.w);
            bool hit = st == kStateFull && pa.x == (uint32_t)pkey && pa.y == (uint32_t)(pkey >> 32) && meta_model(pb.w) == mdl;
            while (!hit && st != kStateEmpty) {                   // linear probing past a collision
                pslot = (pslot + 1) & t.req_mask;
                const uint4* sp = reinterpret_cast<const uint4*>(t.req + pslot);
                pa = ld_nc_v4(sp); pb = ld_nc_v4(sp + 1);
                st = meta_state(pb.w);
                hit = st == kStateFull && pa.x == (uint32_t)pkey && pa.y == (uint32_t)(pkey >> 32) && meta_model(pb.w) == mdl;
            }
            if (!hit) { finished = true; }
            else {
                SlotWords w; w.a = pa; w.b = pb;
                const uint32_t cnt = meta_count(pb.w);
                if (pblk == 0) {
                    // activePods := pods of block 0 (after the filter); score = max weight   (kvblock_scorer.go:118-128)
                    k = 0;
                    for (uint32_t j = 0; j < cnt; ++j) {
                        const uint32_t pt = slot_ent(w, j), p = pt >> 4;
                        if (frow && !filter_has(frow, p)) continue;
                        const double wt = sm.weight[pt & 15u];
                        uint32_t q = 0;
                        for (; q < k; ++q) if (W.pod[q][lane] == p) break;
                        if (q == k) { W.pod[k][lane] = (uint16_t)p; W.sc[k][lane] = 0.0; ++k; }
                        if (wt > W.sc[q][lane]) W.sc[q][lane] = wt;
                    }
                    alive = (1u << k) - 1u;
                } else {
                    // activePods &= pods(block); score[p] += max weight, in block order   (kvblock_scorer.go:130-147)
                    uint32_t am = alive;
                    while (am) {
                        const int q = __ffs(am) - 1; am &= am - 1;
                        const uint32_t want = W.pod[q][lane];
                        double mx = 0.0; bool present = false;
                        for (uint32_t j = 0; j < cnt; ++j) {
                            const uint32_t pt = slot_ent(w, j);
                            if ((pt >> 4) == want) { present = true; const double wt = sm.weight[pt & 15u]; if (wt > mx) mx = wt; }
                        }
                        if (present) W.sc[q][lane] = __dadd_rn(W.sc[q][lane], mx);
                        else alive &= ~(1u << q);
                    }
                }
                if (!alive || pblk == nblk - 1) finished = true;
            }
        }
        // 5. issue the probe for the block just hashed
        if (staged_cur) {
            if (!finished) {
                pkey = key; pblk = hblk;
                pslot = home_of(key, mdl) & t.req_mask;
                const uint4* sp = reinterpret_cast<const uint4*>(t.req + pslot);
                pa = ld_nc_v4(sp); pb = ld_nc_v4(sp + 1);
                pend = true;
            }
            ++hblk;
        }
        // 6. retire + refill
        if (__any_sync(0xffffffffu, finished)) {
            if (finished) { staged_next = false; pend = false; }
            refill();
        }
        if (!__any_sync(0xffffffffu, pi >= 0)) break;
    }
    // every issued bulk copy must have landed before the CTA's shared memory is released: the
    // copies issued in step 1 of the last iteration target stage (it+1)&1.
    mbar_wait(((it + 1) & 1u) ? bar1 : bar0, ((it + 1) >> 1) & 1u);
}

// Host side -------------------------------------------------------------------------------------
inline int score_tuned_init() {
    cudaError_t e = cudaFuncSetAttribute(score_kernel_tuned<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ScoreSmem<16>));
    if (e != cudaSuccess) return -1;
    return 0;
}

inline int launch_score_tuned(const TableView& t, int sm_count, const uint32_t* d_tok, const int64_t* d_off, int64_t tok_base, int64_t n,
                              const uint32_t* d_model, uint32_t model0, const uint64_t* d_filter, double* dense, uint16_t* sp_pods,
                              double* sp_scores, uint8_t* sp_cnt, uint8_t* has_keys, unsigned long long* d_next, cudaStream_t st,
                              uint64_t* launches) {
    if (t.block_size != 16) {
        const int T = 128;
        score_kernel_v1<<<(unsigned)((n + T - 1) / T), T, 0, st>>>(t, d_tok, d_off, tok_base, n, d_model, model0, d_filter, dense, sp_pods, sp_scores, sp_cnt, has_keys);
        *launches += 1;
        return cudaGetLastError() == cudaSuccess ? 0 : -1;
    }
    if (cudaMemsetAsync(d_next, 0, sizeof(unsigned long long), st) != cudaSuccess) return -1;
    ScoreArgs a{d_tok, d_off, tok_base, n, d_model, model0, d_filter, dense, sp_pods, sp_scores, sp_cnt, has_keys, d_next};
    const int64_t lanes_per_cta = kScoreThreads;
    int64_t ctas = (n + lanes_per_cta - 1) / lanes_per_cta;
    const int64_t max_ctas = (int64_t)sm_count * kScoreCtasPerSm;
    if (ctas > max_ctas) ctas = max_ctas;
    score_kernel_tuned<16><<<(unsigned)ctas, kScoreThreads, sizeof(ScoreSmem<16>), st>>>(t, a);
    *launches += 1;
    if (cudaGetLastError() != cudaSuccess) return -1;
    return 0;
}

}  // namespace kvx
