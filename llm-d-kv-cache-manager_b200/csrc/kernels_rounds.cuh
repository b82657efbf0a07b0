// kernels_rounds.cuh -- Score() for LARGE batches as alternating rounds of two specialised kernels.
//
// Same reference path as kernels_score.cuh (GetPodScores steps 2-4, pkg/kvcache/indexer.go:141-163);
// different decomposition.  The fused persistent kernel makes every lane a little state machine (stage,
// hash, probe, score, retire, refill); ncu showed it stuck at ~0.4 IPC per scheduler with ~20 resident
// warps/SM, because the state each lane carries (~94 registers + ~300 B shared) caps occupancy while
// the non-hash phases add serial latency per block.  Here the two halves of the work get the shape
// each one wants:
//
//   round r, kernel H (hash_round_kernel):  prompts that are still on the consecutive-prefix walk hash
//       their next kRoundBlocks blocks.  Lockstep, thin lanes (no score / probe state), the branch-free
//       FNV/CBOR code at the pipe-saturating occupancy measured by scripts/ubench_hash.cu.  Keys go to
//       a per-round buffer in HBM (8 B per block -- versus 64 B of tokens and 64 B of slot read).
//   round r, kernel P (probe_round_kernel): one WARP per prompt: the 32 lanes probe the round's 32 keys
//       at once (32 independent 64-byte reads in flight per warp -> the memory system sees full
//       parallelism instead of one dependent probe per lane), a ballot finds the first miss, and the
//       longest-prefix walk + in-order f64 accumulation runs on the hits (lane q owns block-0 pod q).
//       Prompts whose walk continues are appended to the next round's list.
//
// Early termination costs at most one round of extra hashing per prompt.  Results are bit-identical to
// the fused kernel and the oracle (tests run both paths).
#pragma once
#include <cuda_runtime.h>
#include "kernels_score.cuh"

namespace kvx {

constexpr int kRoundBlocks = 32;         // blocks hashed per prompt per round (== lanes per warp in kernel P)
constexpr int kHashThreads = 256;
constexpr int kProbeThreads = 256;

struct __align__(8) PromptState {        // walk state carried between rounds (only for prompts that continue)
    double sc[kMaxEnt];
    uint16_t pod[kMaxEnt];
    uint16_t alive;                      // bitmask over [0,k)
    uint8_t k;
    uint8_t pad;
};

struct RoundBufs {
    uint32_t* act[2];                    // active prompt lists (ping-pong)
    unsigned int* n_act;                 // [2] list lengths
    uint64_t* hstate;                    // chain hash after the last hashed block, per prompt
    uint64_t* keys;                      // [n_act][kRoundBlocks] keys of the current round
    PromptState* pst;                    // per prompt
};

template <int BS> struct HashSmem {
    static constexpr int kRow = BS * 4 + 16;
    unsigned char tok[kHashThreads / 32][2][32 * kRow];
};

// ---- kernel H ---------------------------------------------------------------------------------
template <int BS>
__global__ void __launch_bounds__(kHashThreads, 4)
hash_round_kernel(const TableView t, const ScoreArgs a, const RoundBufs rb, const int cur, const int round) {
    static_assert(BS == 16, "staging pattern is written for 16-token blocks");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    using SM = HashSmem<BS>;
    SM& sm = *reinterpret_cast<SM*>(smem_raw);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const unsigned int n_act = rb.n_act[cur];
    if (blockIdx.x == 0 && threadIdx.x == 0) rb.n_act[cur ^ 1] = 0;      // next round's list starts empty
    const unsigned int total_warps = gridDim.x * (kHashThreads / 32);
    for (unsigned int w = blockIdx.x * (kHashThreads / 32) + wid; w * 32u < n_act; w += total_warps) {
        const unsigned int i = w * 32u + lane;                            // slot in the active list
        const bool have = i < n_act;
        const uint32_t p = have ? rb.act[cur][i] : 0u;
        int nb = 0;                                                      // blocks of this prompt in this round
        const uint32_t* src = nullptr;
        bool aligned = true;
        uint64_t h = 0;
        if (have) {
            const int64_t b = a.tok_off[p] - a.tok_base, e = a.tok_off[p + 1] - a.tok_base;
            const int64_t nblk = (e - b) / BS;
            const int64_t first = (int64_t)round * kRoundBlocks;
            nb = (int)max((int64_t)0, min((int64_t)kRoundBlocks, nblk - first));
            src = a.tok + b + first * BS;
            aligned = (reinterpret_cast<uintptr_t>(src) & 15u) == 0;
            h = round == 0 ? t.init_hash : rb.hstate[p];
        }
        const int nb_max = __reduce_max_sync(0xffffffffu, nb);
        uint64_t* krow = rb.keys + (size_t)i * kRoundBlocks;
        auto stage = [&](int s, int b) {
            const bool issue = b < nb;
            const unsigned long long srcv = (issue && aligned) ? (unsigned long long)(uintptr_t)(src + (size_t)b * BS) : 0ull;
            __syncwarp();
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const int q = 8 * r + (lane >> 2);
                const unsigned long long sp = __shfl_sync(0xffffffffu, srcv, q);
                if (sp) cp_async_16(smem_addr(&sm.tok[wid][s][q * SM::kRow + (lane & 3) * 16]), reinterpret_cast<const char*>(sp) + (lane & 3) * 16);
            }
            cp_async_commit();
            if (issue && !aligned) {
                uint32_t* dst = reinterpret_cast<uint32_t*>(&sm.tok[wid][s][lane * SM::kRow]);
                const uint32_t* g = src + (size_t)b * BS;
                for (int j = 0; j < BS; ++j) dst[j] = __ldg(g + j);
            }
        };
        stage(0, 0);
        for (int b = 0; b < nb_max; ++b) {
            stage((b + 1) & 1, b + 1);
            cp_async_wait<1>();
            __syncwarp();
            Fnv f;
            f.begin_block(h, BS);
            const uint4* tp = reinterpret_cast<const uint4*>(&sm.tok[wid][b & 1][lane * SM::kRow]);
            const uint4 v0 = tp[0], v1 = tp[1];
            f.token(v0.x); f.token(v0.y); f.token(v0.z); f.token(v0.w);
            const uint4 v2 = tp[2];
            f.token(v1.x); f.token(v1.y); f.token(v1.z); f.token(v1.w);
            const uint4 v3 = tp[3];
            f.token(v2.x); f.token(v2.y); f.token(v2.z); f.token(v2.w);
            f.token(v3.x); f.token(v3.y); f.token(v3.z); f.token(v3.w);
            const uint64_t key = f.end_block();
            if (b < nb) { h = key; krow[b] = key; }
        }
        cp_async_wait<0>();
        __syncwarp();
        if (have && nb > 0) rb.hstate[p] = h;
    }
}

// ---- kernel P ---------------------------------------------------------------------------------
__device__ __forceinline__ void ld_slot_pair(const ReqSlot* s, uint4& a0, uint4& b0, uint4& a1, uint4& b1) {
    asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(a0.x), "=r"(a0.y), "=r"(a0.z), "=r"(a0.w), "=r"(b0.x), "=r"(b0.y), "=r"(b0.z), "=r"(b0.w) : "l"(s));
    asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(a1.x), "=r"(a1.y), "=r"(a1.z), "=r"(a1.w), "=r"(b1.x), "=r"(b1.y), "=r"(b1.z), "=r"(b1.w) : "l"(s + 1));
}

// entry j (0..9) of a slot whose words live in this lane's registers
__device__ __forceinline__ uint32_t ent_of(uint32_t e0, uint32_t e1, uint32_t e2, uint32_t e3, uint32_t e4, int j) {
    const uint32_t word = j < 2 ? e0 : j < 4 ? e1 : j < 6 ? e2 : j < 8 ? e3 : e4;
    return (j & 1) ? (word >> 16) : (word & 0xffffu);
}

__global__ void __launch_bounds__(kProbeThreads)
probe_round_kernel(const TableView t, const ScoreArgs a, const RoundBufs rb, const int cur, const int round) {
    __shared__ double s_weight[16];
    if (threadIdx.x < 16) s_weight[threadIdx.x] = t.weight[threadIdx.x];
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const unsigned int n_act = rb.n_act[cur];
    const unsigned int total_warps = gridDim.x * (kProbeThreads / 32);
    for (unsigned int i = blockIdx.x * (kProbeThreads / 32) + wid; i < n_act; i += total_warps) {
        const uint32_t p = rb.act[cur][i];
        const int64_t tb = a.tok_off[p] - a.tok_base, te = a.tok_off[p + 1] - a.tok_base;
        const int64_t nblk = (te - tb) / t.block_size;
        const int64_t first = (int64_t)round * kRoundBlocks;
        const int nb = (int)max((int64_t)0, min((int64_t)kRoundBlocks, nblk - first));
        const uint32_t mdl = a.model ? a.model[p] : a.model0;

        // ---- 1. all keys of the round probed at once: lane j <-> block first+j ----
        bool hit = false;
        uint32_t e0 = 0, e1 = 0, e2 = 0, e3 = 0, e4 = 0, cnt = 0;
        if (lane < nb) {
            const uint64_t key = rb.keys[(size_t)i * kRoundBlocks + lane];
            uint64_t slot = slot_home(key, mdl, t.req_mask);
            for (;;) {
                uint4 a0, b0, a1, b1;
                ld_slot_pair(t.req + slot, a0, b0, a1, b1);
                if (slot_matches(a0, b0, key, mdl)) { hit = true; e0 = a0.z; e1 = a0.w; e2 = b0.x; e3 = b0.y; e4 = b0.z; cnt = meta_count(b0.w); break; }
                if (meta_state(b0.w) == kStateEmpty) break;
                if (slot_matches(a1, b1, key, mdl)) { hit = true; e0 = a1.z; e1 = a1.w; e2 = b1.x; e3 = b1.y; e4 = b1.z; cnt = meta_count(b1.w); break; }
                if (meta_state(b1.w) == kStateEmpty) break;
                slot = (slot + 2) & t.req_mask;
            }
        }
        const uint32_t valid = nb >= 32 ? 0xffffffffu : ((1u << nb) - 1u);
        const uint32_t hitmask = __ballot_sync(0xffffffffu, hit) & valid;
        const int nhit = __ffs(~hitmask) - 1;                 // consecutive hits from the round's first block (32 if all)
        // ---- 2. walk state: lane q owns block-0 pod q ----
        uint32_t k = 0, alive = 0;
        uint32_t mypod = 0xffffffffu; double mysc = 0.0;
        int start = 0;                                        // first block of this round still to be scored
        if (round == 0) {
            if (nhit > 0) {
                // activePods := pods of block 0 after the filter; score = max weight      (kvblock_scorer.go:118-128)
                const uint32_t f0 = __shfl_sync(0xffffffffu, e0, 0), f1 = __shfl_sync(0xffffffffu, e1, 0), f2 = __shfl_sync(0xffffffffu, e2, 0),
                               f3 = __shfl_sync(0xffffffffu, e3, 0), f4 = __shfl_sync(0xffffffffu, e4, 0), fc = __shfl_sync(0xffffffffu, cnt, 0);
                const uint64_t* frow = filter_row(a.filter, p, t.filter_words);
                const uint32_t myent = lane < (int)fc ? ent_of(f0, f1, f2, f3, f4, lane) : 0u;
                const bool pass = lane < (int)fc && (!frow || filter_has(frow, myent >> 4));
                bool firstocc = pass;                          // first passing occurrence of this pod id in the slot
                for (int j = 0; j < kMaxEnt; ++j) {
                    const uint32_t oe = __shfl_sync(0xffffffffu, myent, j);
                    const bool op = __shfl_sync(0xffffffffu, (int)pass, j);
                    if (j < lane && op && (oe >> 4) == (myent >> 4)) firstocc = false;
                }
                const uint32_t fm = __ballot_sync(0xffffffffu, firstocc);
                k = __popc(fm);
                const int srcl = lane < (int)k ? __fns(fm, 0, lane + 1) : 0;
                const uint32_t pe = __shfl_sync(0xffffffffu, myent, srcl);
                if (lane < (int)k) {
                    mypod = pe >> 4;
                    double mx = 0.0;
                    for (int j = 0; j < (int)fc; ++j) {
                        const uint32_t ee = ent_of(f0, f1, f2, f3, f4, j);
                        if ((ee >> 4) == mypod) { const double wt = s_weight[ee & 15u]; if (wt > mx) mx = wt; }
                    }
                    mysc = mx;
                }
                alive = (1u << k) - 1u;
                start = 1;
            }
        } else {
            const PromptState& ps = rb.pst[p];
            k = ps.k; alive = ps.alive;
            if (lane < (int)k) { mypod = ps.pod[lane]; mysc = ps.sc[lane]; }
        }
        // ---- 3. blocks [start, nhit): activePods &= pods(block); score += max weight, in block order (kvblock_scorer.go:130-147)
        //         consecutive blocks with bitwise-equal entry words form a run: one scan, then `len` ordered adds.
        {
            const uint32_t pe0 = __shfl_up_sync(0xffffffffu, e0, 1), pe1 = __shfl_up_sync(0xffffffffu, e1, 1), pe2 = __shfl_up_sync(0xffffffffu, e2, 1),
                           pe3 = __shfl_up_sync(0xffffffffu, e3, 1), pe4 = __shfl_up_sync(0xffffffffu, e4, 1), pc = __shfl_up_sync(0xffffffffu, cnt, 1);
            const bool same_as_prev = lane > 0 && (((pe0 ^ e0) | (pe1 ^ e1) | (pe2 ^ e2) | (pe3 ^ e3) | (pe4 ^ e4) | (pc ^ cnt)) == 0u);
            const uint32_t samemask = __ballot_sync(0xffffffffu, same_as_prev);
            int b = start;
            while (b < nhit && alive) {
                // run [b, e): block b plus the following blocks whose pattern equals their predecessor's
                const uint32_t follow = (samemask >> (b + 1 < 32 ? b + 1 : 31)) ;
                int len = 1;
                if (b + 1 < 32) len += __ffs(~follow) - 1;
                if (b + len > nhit) len = nhit - b;
                const uint32_t r0 = __shfl_sync(0xffffffffu, e0, b), r1 = __shfl_sync(0xffffffffu, e1, b), r2 = __shfl_sync(0xffffffffu, e2, b),
                               r3 = __shfl_sync(0xffffffffu, e3, b), r4 = __shfl_sync(0xffffffffu, e4, b), rc = __shfl_sync(0xffffffffu, cnt, b);
                bool present = false; double mx = 0.0;
                if (lane < (int)k && ((alive >> lane) & 1u)) {
                    for (int j = 0; j < (int)rc; ++j) {
                        const uint32_t ee = ent_of(r0, r1, r2, r3, r4, j);
                        if ((ee >> 4) == mypod) { present = true; const double wt = s_weight[ee & 15u]; if (wt > mx) mx = wt; }
                    }
                    if (present) for (int x = 0; x < len; ++x) mysc = __dadd_rn(mysc, mx);
                }
                alive &= __ballot_sync(0xffffffffu, present);
                b += len;
            }
        }
        // ---- 4. finished, or continue next round ----
        const bool more = alive != 0 && nhit == nb && nb == kRoundBlocks && first + nb < nblk;
        if (more) {
            PromptState& ps = rb.pst[p];
            if (lane < (int)k) { ps.pod[lane] = (uint16_t)mypod; ps.sc[lane] = mysc; }
            if (lane == 0) { ps.k = (uint8_t)k; ps.alive = (uint16_t)alive; rb.act[cur ^ 1][atomicAdd(&rb.n_act[cur ^ 1], 1u)] = p; }
        } else {
            if (a.dense) {
                double* row = a.dense + (long long)p * t.max_pods;
                const uint32_t P = t.max_pods;
                if ((P & 1u) == 0 && ((reinterpret_cast<uintptr_t>(row) & 15u) == 0)) {
                    for (uint32_t c = lane * 2; c < P; c += 64) *reinterpret_cast<double2*>(row + c) = make_double2(-1.0, -1.0);
                } else {
                    for (uint32_t c = lane; c < P; c += 32) row[c] = -1.0;
                }
                __syncwarp();
                if (lane < (int)k && mypod < P) row[mypod] = mysc;
            }
            if (a.sp_cnt) {
                if (lane < (int)k) { a.sp_pods[(long long)p * kMaxEnt + lane] = (uint16_t)mypod; a.sp_scores[(long long)p * kMaxEnt + lane] = mysc; }
                if (lane == 0) a.sp_cnt[p] = (uint8_t)k;
            }
            if (a.has_keys && lane == 0) a.has_keys[p] = nblk > 0;
        }
        __syncwarp();
    }
}

__global__ void rounds_init_kernel(const ScoreArgs a, const RoundBufs rb, uint32_t block_size, unsigned long long* max_blocks) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    unsigned long long nb = 0;
    if (i < a.n_prompts) {
        rb.act[0][i] = (uint32_t)i;
        nb = (unsigned long long)((a.tok_off[i + 1] - a.tok_off[i]) / block_size);
    }
    nb = __reduce_max_sync(0xffffffffu, (unsigned)min(nb, 0xffffffffull));
    if ((threadIdx.x & 31) == 0 && nb) atomicMax(max_blocks, nb);
    if (i == 0) { rb.n_act[0] = (unsigned int)a.n_prompts; rb.n_act[1] = 0; }
}

inline int rounds_init() {
    return cudaFuncSetAttribute(hash_round_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(HashSmem<16>)) == cudaSuccess ? 0 : -1;
}

}  // namespace kvx
