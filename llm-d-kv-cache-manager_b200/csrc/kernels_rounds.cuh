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
constexpr int kGroupThreads = 256;
constexpr uint32_t kRoleSelf = 0xffffffffu;

struct __align__(8) PromptState {        // walk state carried between rounds (only for prompts that continue)
    double sc[kMaxEnt];
    uint16_t pod[kMaxEnt];
    uint32_t pat[6];                     // entry words + count of the last scored block's slot
    uint16_t alive;                      // bitmask over [0,k)
    uint8_t k;
    uint8_t pad;
    uint8_t bt[kMaxEnt];                 // tier giving pod q its max weight in that slot (0xff: 0.0)
    uint8_t pad2[2];
};

constexpr int kMaxParts = 4;              // the sorted batch runs as up to this many independent parts on their own streams

struct RoundBufs {
    uint32_t* act[2];                    // live prompt lists (ping-pong)
    unsigned int* n_act;                 // [2] list lengths
    uint64_t* hstate;                    // per prompt: chain hash after the last block of the previous round
    uint32_t* src;                       // per prompt: the prompt whose PromptState (previous round's buffer) is this prompt's
                                         // walk state -- itself, or the representative of the class it was in
    PromptState* pst[2];                 // per prompt, by round parity (round r reads [r&1 ^ 1], writes [r&1])
    uint64_t* keys;                      // [kRoundBlocks][n_prompts] keys of the current round, block-major, by position in
                                         // the representative list: key of block j of entry a at keys[j * n_prompts + a]
    uint32_t* nbr;                       // [n_act] per live slot: blocks in this round | (more blocks follow) << 8
    // prefix classes (group_round_kernel): live prompts with the same walk state, model and filter whose next chunk of
    // tokens is identical get identical keys, probes, scores and fate this round.  One representative per class is
    // hashed and walked; the others wait in the follower list and take the representative's outcome.
    uint32_t* role;                      // [n_act] kRoleSelf: representative (or alone); else the live slot of its representative
    uint8_t* fate;                       // [n_act] written for representatives by kernel P: kFateMore / kFateDone
    uint32_t* hl;                        // [n_act] representative list (live slots), compacted
    uint32_t* fl;                        // [n_act] follower list (live slots), compacted
    unsigned int* n_hl;                  // n_hl[0] = representatives, n_hl[1] = followers (zeroed before every round)
    uint32_t* map;                       // class election: bucket -> live slot (kRoleSelf = empty), cleared before every round
    uint32_t map_mask;
};
constexpr uint8_t kFateMore = 1, kFateDone = 2;

constexpr int kHashChunk = 1;            // blocks staged per copy step (2 = 128-byte accesses was measured: fewer, longer DRAM
                                         // accesses but only 24 resident warps/SM -> 6 % slower; the kernel is pipe bound)
template <int BS> struct HashSmem {
    static constexpr int kRow = kHashChunk * BS * 4 + 16;   // 144 B: the four LDS.128 of a block stay conflict free
    unsigned char tok[kHashThreads / 32][2][32 * kRow];
};

// ---- kernel G: prefix classes ------------------------------------------------------------------------------
// A warp takes 32 live slots.  Metadata is lane-parallel; each prompt's next chunk (<= 32 blocks = 2 KB) is then read by
// the whole warp (coalesced) and folded into a 64-bit fingerprint.  Lane-parallel again, every prompt tries to claim
// bucket hash(state source, chain state, model, fingerprint) of a small map: the winner represents the class; a loser
// compares itself with the bucket owner -- state source, chain state, block count, model, filter row, and the chunk
// token by token (warp-wide, 16 bytes per lane).  Only an exact match joins the class, so fingerprint and bucket
// collisions can cost sharing, never correctness.  Representatives and followers are compacted into two lists.
__device__ __forceinline__ void chunk_load(const uint32_t* sp, int nw, int lane, uint4 (&v)[4]) {
    const bool al = (reinterpret_cast<uintptr_t>(sp) & 15u) == 0;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const int w0 = (c * 32 + lane) * 4;
        v[c] = make_uint4(0, 0, 0, 0);
        if (w0 < nw) {
            if (al) v[c] = __ldg(reinterpret_cast<const uint4*>(sp + w0));
            else { v[c].x = __ldg(sp + w0); v[c].y = __ldg(sp + w0 + 1); v[c].z = __ldg(sp + w0 + 2); v[c].w = __ldg(sp + w0 + 3); }
        }
    }
}

template <int BS>
__global__ void __launch_bounds__(kGroupThreads)
group_round_kernel(const TableView t, const ScoreArgs a, const RoundBufs rb, const int cur, const int round, const int dedup) {
    static_assert(BS == 16 && kRoundBlocks == 32, "a chunk is 4 x 32 lanes x 16 bytes");
    const int lane = threadIdx.x & 31;
    const unsigned int n_act = rb.n_act[cur];
    if (blockIdx.x == 0 && threadIdx.x == 0) rb.n_act[cur ^ 1] = 0;      // next round's list starts empty
    const unsigned int total_warps = gridDim.x * (kGroupThreads / 32);
    for (unsigned int w = blockIdx.x * (kGroupThreads / 32) + (threadIdx.x >> 5); w * 32u < n_act; w += total_warps) {
        const unsigned int i = w * 32u + lane;
        const bool have = i < n_act;
        uint32_t p = 0, srcp = kRoleSelf, mdl = a.model0;
        int nb = 0; bool more = false;
        uint64_t hprev = t.init_hash;
        const uint32_t* tp = a.tok;
        int64_t first = (int64_t)round * kRoundBlocks;
        if (have) {
            p = rb.act[cur][i];
            const int64_t b = a.tok_off[p] - a.tok_base, e = a.tok_off[p + 1] - a.tok_base;
            const int64_t nblk = (e - b) / BS;
            nb = (int)max((int64_t)0, min((int64_t)kRoundBlocks, nblk - first));
            more = first + nb < nblk;
            rb.nbr[i] = (uint32_t)nb | (more ? 0x100u : 0u);
            tp = a.tok + b + first * BS;
            if (round > 0) { srcp = rb.src[p]; hprev = rb.hstate[p]; }
            if (a.model) mdl = a.model[p];
        }
        const bool cls = have && nb > 0 && dedup;
        // Does the prompt equal its predecessor in the list?  (The list is sorted by prefix, so mostly yes.)  Everything
        // but the tokens lane-parallel here; the tokens as the chunks stream by below.
        uint64_t fsum = 0;
        if (cls && round == 0 && a.filter) { const uint64_t* fr = a.filter + (int64_t)p * t.filter_words; for (uint32_t x = 0; x < t.filter_words; ++x) fsum = (fsum ^ fr[x]) * 0x9E3779B97F4A7C15ull; }
        bool eqm;
        {
            const int nb_u = __shfl_up_sync(0xffffffffu, nb, 1); const int more_u = __shfl_up_sync(0xffffffffu, (int)more, 1);
            const uint32_t mdl_u = __shfl_up_sync(0xffffffffu, mdl, 1), src_u = __shfl_up_sync(0xffffffffu, srcp, 1);
            const uint64_t h_u = __shfl_up_sync(0xffffffffu, hprev, 1), f_u = __shfl_up_sync(0xffffffffu, fsum, 1);
            const int cls_u = __shfl_up_sync(0xffffffffu, (int)cls, 1);
            eqm = cls && lane > 0 && cls_u && nb_u == nb && more_u == (int)more && mdl_u == mdl && src_u == srcp && h_u == hprev && f_u == fsum;
            if (eqm && round == 0 && a.filter) {       // equal filter fingerprints: compare the rows themselves
                const uint32_t pu = rb.act[cur][i - 1];
                const uint64_t* fa = a.filter + (int64_t)p * t.filter_words; const uint64_t* fb = a.filter + (int64_t)pu * t.filter_words;
                for (uint32_t x = 0; x < t.filter_words; ++x) eqm = eqm && fa[x] == fb[x];
            }
        }
        // chunks, one prompt at a time, the next one in flight while this one is compared with the previous one (still in
        // registers) and -- only if it differs -- folded into a fingerprint for the election
        uint32_t f0 = 0, f1 = 0;
        bool eqprev = false;                         // same class as the previous list entry
        uint32_t todo = __ballot_sync(0xffffffffu, cls);
        const uint32_t eqm_mask = __ballot_sync(0xffffffffu, eqm);
        uint4 v[4], vn[4], vp[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) vp[c] = make_uint4(0, 0, 0, 0);
        int q = todo ? __ffs(todo) - 1 : -1, qprev = -2;
        if (q >= 0) chunk_load(reinterpret_cast<const uint32_t*>(__shfl_sync(0xffffffffu, (unsigned long long)(uintptr_t)tp, q)),
                               __shfl_sync(0xffffffffu, nb, q) * BS, lane, v);
        while (q >= 0) {
            todo &= todo - 1;
            const int qn = todo ? __ffs(todo) - 1 : -1;
            if (qn >= 0) chunk_load(reinterpret_cast<const uint32_t*>(__shfl_sync(0xffffffffu, (unsigned long long)(uintptr_t)tp, qn)),
                                    __shfl_sync(0xffffffffu, nb, qn) * BS, lane, vn);
            bool same = false;
            if (qprev == q - 1 && ((eqm_mask >> q) & 1u)) {          // warp-uniform
                uint32_t d = 0;
#pragma unroll
                for (int c = 0; c < 4; ++c) d |= (v[c].x ^ vp[c].x) | (v[c].y ^ vp[c].y) | (v[c].z ^ vp[c].z) | (v[c].w ^ vp[c].w);
                same = __all_sync(0xffffffffu, d == 0u);
            }
            if (same) { if (lane == q) eqprev = true; }
            else {
                uint32_t a0 = 0, a1 = 0;
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const uint32_t w0 = (uint32_t)(c * 32 + lane) * 4u;
                    uint32_t x = (v[c].x ^ (w0 * 0x9E3779B1u + 0x7F4A7C15u)) * 0x85EBCA6Bu; x = (x ^ v[c].y) * 0xC2B2AE35u; x ^= x >> 15;
                    uint32_t y = (v[c].z ^ (w0 * 0x7FEB352Du + 0x165667B1u)) * 0x846CA68Bu; y = (y ^ v[c].w) * 0x9E3779B1u; y ^= y >> 13;
                    a0 ^= x + y; a1 ^= x * 0x27D4EB2Fu ^ y;
                }
                a0 = __reduce_xor_sync(0xffffffffu, a0);
                a1 = __reduce_xor_sync(0xffffffffu, a1);
                if (lane == q) { f0 = a0; f1 = a1; }
            }
#pragma unroll
            for (int c = 0; c < 4; ++c) { vp[c] = v[c]; v[c] = vn[c]; }
            qprev = q; q = qn;
        }
        // election among the prompts that differ from their predecessor
        uint32_t cand = kRoleSelf;
        if (cls && !eqprev) {
            const uint64_t key = mix64((((uint64_t)f1 << 32) | f0) ^ (hprev * 0xFF51AFD7ED558CCDull) ^ ((uint64_t)srcp << 17) ^ (uint64_t)(nb | (more ? 64 : 0)) ^
                                       ((uint64_t)mdl * 0xC2B2AE3D27D4EB4Full) ^ fsum);
            cand = atomicCAS(&rb.map[(uint32_t)key & rb.map_mask], kRoleSelf, i);
        }
        // a loser checks everything but the tokens lane-parallel ...
        bool okm = false;
        const uint32_t* lp = a.tok;
        if (cand != kRoleSelf) {
            const uint32_t pl = rb.act[cur][cand];
            const int64_t bl = a.tok_off[pl] - a.tok_base, el = a.tok_off[pl + 1] - a.tok_base;
            const int64_t nblk_l = (el - bl) / BS;
            const int nbl = (int)max((int64_t)0, min((int64_t)kRoundBlocks, nblk_l - first));
            okm = nbl == nb && (first + nbl < nblk_l) == more && (a.model ? a.model[pl] : a.model0) == mdl;
            if (round > 0) okm = okm && rb.src[pl] == srcp && rb.hstate[pl] == hprev;
            else if (a.filter) {
                const uint64_t* fa = a.filter + (int64_t)p * t.filter_words; const uint64_t* fb = a.filter + (int64_t)pl * t.filter_words;
                for (uint32_t x = 0; x < t.filter_words; ++x) okm = okm && fa[x] == fb[x];
            }
            lp = a.tok + bl + first * BS;
        }
        // ... and the tokens warp-wide
        bool shared = false;
        uint32_t vm = __ballot_sync(0xffffffffu, okm);
        while (vm) {
            const int l = __ffs(vm) - 1; vm &= vm - 1;
            const uint32_t* s1 = reinterpret_cast<const uint32_t*>(__shfl_sync(0xffffffffu, (unsigned long long)(uintptr_t)tp, l));
            const uint32_t* s2 = reinterpret_cast<const uint32_t*>(__shfl_sync(0xffffffffu, (unsigned long long)(uintptr_t)lp, l));
            const int nw = __shfl_sync(0xffffffffu, nb, l) * BS;
            chunk_load(s1, nw, lane, v);
            chunk_load(s2, nw, lane, vn);
            uint32_t d = 0;
#pragma unroll
            for (int c = 0; c < 4; ++c) d |= (v[c].x ^ vn[c].x) | (v[c].y ^ vn[c].y) | (v[c].z ^ vn[c].z) | (v[c].w ^ vn[c].w);
            const bool same = __all_sync(0xffffffffu, d == 0u);
            if (lane == l) shared = same;
        }
        // a prompt equal to its predecessor takes the class of the nearest preceding prompt that went through the election
        {
            const uint32_t heads = ~__ballot_sync(0xffffffffu, eqprev);            // lane 0 is always a head
            const int hd = 31 - __clz(heads & ((2u << lane) - 1u));
            const uint32_t cand_h = __shfl_sync(0xffffffffu, cand, hd);
            const int shared_h = __shfl_sync(0xffffffffu, (int)shared, hd);
            const unsigned int i_h = __shfl_sync(0xffffffffu, i, hd);
            if (eqprev) { shared = true; cand = shared_h ? cand_h : i_h; }
        }
        if (have) rb.role[i] = shared ? cand : kRoleSelf;
        // compaction: representatives (and prompts on their own) / followers
        const uint32_t ma = __ballot_sync(0xffffffffu, have && !shared), mf = __ballot_sync(0xffffffffu, shared);
        unsigned int ba = 0, bf = 0;
        if (lane == 0) { if (ma) ba = atomicAdd(&rb.n_hl[0], (unsigned int)__popc(ma)); if (mf) bf = atomicAdd(&rb.n_hl[1], (unsigned int)__popc(mf)); }
        ba = __shfl_sync(0xffffffffu, ba, 0); bf = __shfl_sync(0xffffffffu, bf, 0);
        const uint32_t below = (1u << lane) - 1u;
        if (have && !shared) rb.hl[ba + __popc(ma & below)] = i;
        if (shared) rb.fl[bf + __popc(mf & below)] = i;
    }
}

// ---- kernel H ---------------------------------------------------------------------------------
template <int BS>
__global__ void __launch_bounds__(kHashThreads, 4)
hash_round_kernel(const TableView t, const ScoreArgs a, const RoundBufs rb, const int cur, const int round) {
    static_assert(BS == 16, "staging pattern is written for 16-token blocks");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    using SM = HashSmem<BS>;
    SM& sm = *reinterpret_cast<SM*>(smem_raw);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const unsigned int n_hl = rb.n_hl[0];                                 // representatives this round (kernel G)
    const unsigned int total_warps = gridDim.x * (kHashThreads / 32);
    for (unsigned int w = blockIdx.x * (kHashThreads / 32) + wid; w * 32u < n_hl; w += total_warps) {
        const unsigned int i = w * 32u + lane;                            // position in the representative list
        const bool have = i < n_hl;
        const uint32_t p = have ? rb.act[cur][rb.hl[i]] : 0u;
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
        // stage chunk c (kHashChunk blocks) of every lane's prompt: 4*kHashChunk lanes move one prompt's contiguous bytes.
        auto stage = [&](int s, int c) {
            const int b0 = c * kHashChunk;
            const bool issue = b0 < nb;
            const int nbytes = issue ? min(kHashChunk, nb - b0) * BS * 4 : 0;
            const unsigned long long srcv = (issue && aligned) ? (unsigned long long)(uintptr_t)(src + (size_t)b0 * BS) : 0ull;
            __syncwarp();
            constexpr int LPP = 4 * kHashChunk;             // lanes that move one prompt's chunk (16 B each)
            constexpr int PPI = 32 / LPP;                   // prompts per copy instruction
#pragma unroll
            for (int r = 0; r < 32 / PPI; ++r) {
                const int q = PPI * r + lane / LPP;
                const unsigned long long sp = __shfl_sync(0xffffffffu, srcv, q);
                const int nby = kHashChunk == 1 ? BS * 4 : __shfl_sync(0xffffffffu, nbytes, q);
                if (sp && (lane % LPP) * 16 < nby)
                    cp_async_16(smem_addr(&sm.tok[wid][s][q * SM::kRow + (lane % LPP) * 16]), reinterpret_cast<const char*>(sp) + (lane % LPP) * 16);
            }
            cp_async_commit();
            if (issue && !aligned) {
                uint32_t* dst = reinterpret_cast<uint32_t*>(&sm.tok[wid][s][lane * SM::kRow]);
                const uint32_t* g = src + (size_t)b0 * BS;
                for (int j = 0; j < nbytes / 4; ++j) dst[j] = __ldg(g + j);
            }
        };
        stage(0, 0);
        const int nchunks = (nb_max + kHashChunk - 1) / kHashChunk;
        for (int c = 0; c < nchunks; ++c) {
            stage((c + 1) & 1, c + 1);
            cp_async_wait<1>();
            __syncwarp();
#pragma unroll
            for (int u = 0; u < kHashChunk; ++u) {
                const int b = c * kHashChunk + u;
                Fnv f;
                f.begin_block(h, BS);
                const uint4* tp = reinterpret_cast<const uint4*>(&sm.tok[wid][c & 1][lane * SM::kRow + u * BS * 4]);
                const uint4 v0 = tp[0], v1 = tp[1];
                f.token(v0.x); f.token(v0.y); f.token(v0.z); f.token(v0.w);
                const uint4 v2 = tp[2];
                f.token(v1.x); f.token(v1.y); f.token(v1.z); f.token(v1.w);
                const uint4 v3 = tp[3];
                f.token(v2.x); f.token(v2.y); f.token(v2.z); f.token(v2.w);
                f.token(v3.x); f.token(v3.y); f.token(v3.z); f.token(v3.w);
                const uint64_t key = f.end_block();
                if (b < nb) { h = key; rb.keys[(size_t)b * a.n_prompts + i] = key; }
            }
        }
        cp_async_wait<0>();
        __syncwarp();
    }
}

// ---- kernel P ---------------------------------------------------------------------------------
__device__ __forceinline__ void ld_slot_pair(const ReqSlot* s, bool peer, uint4& a0, uint4& b0, uint4& a1, uint4& b1) {
    ld_slot(s, peer, a0, b0);
    ld_slot(s + 1, peer, a1, b1);
}

// entry j (0..9) of a slot whose words live in this lane's registers
__device__ __forceinline__ uint32_t ent_of(uint32_t e0, uint32_t e1, uint32_t e2, uint32_t e3, uint32_t e4, int j) {
    const uint32_t word = j < 2 ? e0 : j < 4 ? e1 : j < 6 ? e2 : j < 8 ? e3 : e4;
    return (j & 1) ? (word >> 16) : (word & 0xffffu);
}

// Kernel P, lane-per-prompt.  (Two warp-per-prompt versions came first -- 32 lanes probing a prompt's 32 keys at
// once, then the same software-pipelined.  Both cost ~700 issue slots per prompt-round because the walk itself
// runs on <= 10 lanes while the other 22 idle, and stayed at ~25 % issue utilisation: profiles/r1d_*.)  Here a
// warp walks 32 prompts in lockstep, one block per iteration: every instruction serves 32 prompts, the 32 lanes'
// probes are 32 independent 64-byte reads, and the slot pair for block j+1 is in flight while block j is scored.
struct WalkSmem {
    struct Warp {
        double sc[kMaxEnt][32];
        uint16_t pod[kMaxEnt][32];
        uint8_t bt[kMaxEnt][32];
    };
    Warp w[kProbeThreads / 32];
    double weight[16];
};

__global__ void __launch_bounds__(kProbeThreads, 4)
probe_round_kernel(const TableView t, const ScoreArgs a, const RoundBufs rb, const int cur, const int round) {
    extern __shared__ __align__(16) unsigned char smem_raw_w[];
    WalkSmem& sm = *reinterpret_cast<WalkSmem*>(smem_raw_w);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    WalkSmem::Warp& W = sm.w[wid];
    if (threadIdx.x < 16) sm.weight[threadIdx.x] = t.weight[threadIdx.x];
    __syncthreads();
    const unsigned int n_act = rb.n_hl[0];                     // representatives this round
    const size_t kstride = (size_t)a.n_prompts;
    const bool peer = t.shard_bits != 0;
    const PromptState* pst_prev = rb.pst[(round & 1) ^ 1];
    PromptState* pst_cur = rb.pst[round & 1];
    const unsigned int total_warps = gridDim.x * (kProbeThreads / 32);
    for (unsigned int w = blockIdx.x * (kProbeThreads / 32) + wid; w * 32u < n_act; w += total_warps) {
        const unsigned int i = w * 32u + lane;                 // position in the representative list
        const bool have = i < n_act;
        const unsigned int li = have ? rb.hl[i] : 0u;          // its live slot
        const uint32_t p = have ? rb.act[cur][li] : 0u;
        const uint32_t meta = have ? rb.nbr[li] : 0u;
        const int nb = (int)(meta & 63u);
        const bool has_more = (meta >> 8) & 1u;
        const uint32_t mdl = (have && a.model) ? a.model[p] : a.model0;
        const unsigned int ks = i;
        uint64_t lastkey = 0;
        uint32_t k = 0, alive = 0;
        uint32_t pv0 = 0, pv1 = 0, pv2 = 0, pv3 = 0, pv4 = 0, pvc = 0;      // last scored pattern
        if (round > 0 && have) {
            const PromptState& ps = pst_prev[rb.src[p]];
            k = ps.k; alive = ps.alive;
            pv0 = ps.pat[0]; pv1 = ps.pat[1]; pv2 = ps.pat[2]; pv3 = ps.pat[3]; pv4 = ps.pat[4]; pvc = ps.pat[5];
            for (uint32_t q = 0; q < k; ++q) { W.sc[q][lane] = ps.sc[q]; W.pod[q][lane] = ps.pod[q]; W.bt[q][lane] = ps.bt[q]; }
        }
        const uint64_t* frow = nullptr;
        const int nb_max = __reduce_max_sync(0xffffffffu, nb);
        bool done = !have || nb == 0;                          // walk ended (miss / no live pod); scores are final
        // slot pair of block 0
        // (An L2 prefetch running 8 blocks ahead of the walk was tried and made this kernel 2x slower -- DRAM reads
        //  grew 50 % and issue utilisation fell to 11 %; profiles/r1d_*.  The SM's outstanding-miss capacity, not the
        //  DRAM latency of a single chain, is what bounds it; the batch is sorted by prefix instead so that lanes of
        //  a warp ask for the same slots.)
        uint64_t key = 0, slot = 0;
        const ReqSlot* base = t.req;                           // table (shard) of the current block's key
        uint4 A0 = {0, 0, 0, 0}, B0 = {0, 0, 0, 0}, A1 = {0, 0, 0, 0}, B1 = {0, 0, 0, 0};
        if (!done) {
            key = rb.keys[ks];
            const uint64_t hm = home_of(key, mdl);
            base = t.req_peer[shard_of(hm, t.shard_bits)]; slot = hm & t.req_mask & ~1ull;
            ld_slot_pair(base + slot, peer, A0, B0, A1, B1);
        }
        uint64_t key1 = (!done && nb > 1) ? rb.keys[kstride + ks] : 0ull;                         // key of block j+1
        for (int j = 0; j < nb_max; ++j) {
            // next block's pair is requested before this block is scored; keys are read one iteration ahead of their use
            uint64_t nkey = key1, nslot = 0;
            const ReqSlot* nbase = t.req;
            uint4 nA0 = {0, 0, 0, 0}, nB0 = {0, 0, 0, 0}, nA1 = {0, 0, 0, 0}, nB1 = {0, 0, 0, 0};
            const bool nextv = !done && (j + 1 < nb);
            if (nextv) {
                const uint64_t hm = home_of(nkey, mdl);
                nbase = t.req_peer[shard_of(hm, t.shard_bits)]; nslot = hm & t.req_mask & ~1ull;
                ld_slot_pair(nbase + nslot, peer, nA0, nB0, nA1, nB1);
            }
            key1 = (!done && j + 2 < nb) ? rb.keys[(size_t)(j + 2) * kstride + ks] : 0ull;
            if (!done && j < nb) {
                lastkey = key;
                uint4 A = A0, B = B0;
                bool hit = slot_matches(A, B, key, mdl);
                if (!hit && meta_state(B.w) != kStateEmpty) {
                    A = A1; B = B1;
                    hit = slot_matches(A, B, key, mdl);
                    while (!hit && meta_state(B.w) != kStateEmpty) {          // rare: displaced past the home pair
                        slot = (slot + 2) & t.req_mask;
                        ld_slot_pair(base + slot, peer, A0, B0, A1, B1);
                        A = A0; B = B0; hit = slot_matches(A, B, key, mdl);
                        if (!hit && meta_state(B.w) != kStateEmpty) { A = A1; B = B1; hit = slot_matches(A, B, key, mdl); }
                    }
                }
                if (!hit) { done = true; }
                else {
                    SlotWords sw; sw.a = A; sw.b = B;
                    const uint32_t cnt = meta_count(B.w);
                    const bool first_block = (round == 0 && j == 0);
                    const bool same = !first_block && (((pv0 ^ A.z) | (pv1 ^ A.w) | (pv2 ^ B.x) | (pv3 ^ B.y) | (pv4 ^ B.z) | (pvc ^ cnt)) == 0u);
                    if (same) {
                        uint32_t am = alive;
                        while (am) {
                            const int q = __ffs(am) - 1; am &= am - 1;
                            const uint32_t bt = W.bt[q][lane];
                            const double mx = bt == 0xffu ? 0.0 : sm.weight[bt];
                            W.sc[q][lane] = __dadd_rn(W.sc[q][lane], mx);
                        }
                    } else if (first_block) {
                        // activePods := pods of block 0 (after the filter); score = max weight   (kvblock_scorer.go:118-128)
                        frow = filter_row(a.filter, p, t.filter_words);
                        k = 0;
                        for (uint32_t e = 0; e < cnt; ++e) {
                            const uint32_t pt = slot_ent(sw, e), pd = pt >> 4;
                            if (frow && !filter_has(frow, pd)) continue;
                            const double wt = sm.weight[pt & 15u];
                            uint32_t q = 0;
                            for (; q < k; ++q) if (W.pod[q][lane] == pd) break;
                            if (q == k) { W.pod[k][lane] = (uint16_t)pd; W.sc[k][lane] = 0.0; W.bt[k][lane] = 0xffu; ++k; }
                            if (wt > W.sc[q][lane]) { W.sc[q][lane] = wt; W.bt[q][lane] = (uint8_t)(pt & 15u); }
                        }
                        alive = (1u << k) - 1u;
                    } else {
                        // activePods &= pods(block); score[p] += max weight, in block order   (kvblock_scorer.go:130-147)
                        uint32_t am = alive;
                        while (am) {
                            const int q = __ffs(am) - 1; am &= am - 1;
                            const uint32_t want = W.pod[q][lane];
                            double mx = 0.0; bool present = false; uint32_t bt = 0xffu;
                            for (uint32_t e = 0; e < cnt; ++e) {
                                const uint32_t pt = slot_ent(sw, e);
                                if ((pt >> 4) == want) { present = true; const double wt = sm.weight[pt & 15u]; if (wt > mx) { mx = wt; bt = pt & 15u; } }
                            }
                            if (present) { W.sc[q][lane] = __dadd_rn(W.sc[q][lane], mx); W.bt[q][lane] = (uint8_t)bt; }
                            else alive &= ~(1u << q);
                        }
                    }
                    pv0 = A.z; pv1 = A.w; pv2 = B.x; pv3 = B.y; pv4 = B.z; pvc = cnt;
                    if (!alive) done = true;
                }
            }
            key = nkey; slot = nslot; base = nbase; A0 = nA0; B0 = nB0; A1 = nA1; B1 = nB1;
        }
        // ---- continue next round, or write the result ----
        const bool more = have && !done && has_more;           // all blocks of the round hit, pods still live, blocks left
        const uint32_t mm = __ballot_sync(0xffffffffu, more);
        if (mm) {
            unsigned int base = 0;
            if (lane == 0) base = atomicAdd(&rb.n_act[cur ^ 1], (unsigned int)__popc(mm));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (more) {
                rb.act[cur ^ 1][base + __popc(mm & ((1u << lane) - 1u))] = p;
                PromptState& ps = pst_cur[p];
                rb.hstate[p] = lastkey;                          // chain state for the next round (key of this round's last block)
                rb.src[p] = p;
                ps.k = (uint8_t)k; ps.alive = (uint16_t)alive;
                ps.pat[0] = pv0; ps.pat[1] = pv1; ps.pat[2] = pv2; ps.pat[3] = pv3; ps.pat[4] = pv4; ps.pat[5] = pvc;
                for (uint32_t q = 0; q < k; ++q) { ps.sc[q] = W.sc[q][lane]; ps.pod[q] = W.pod[q][lane]; ps.bt[q] = W.bt[q][lane]; }
            }
        }
        if (have) rb.fate[li] = more ? kFateMore : kFateDone;
        if (have && !more) {                                   // final pods and scores, for followers of this class (kernel R)
            PromptState& ps = pst_cur[p];
            ps.k = (uint8_t)k;
            for (uint32_t q = 0; q < k; ++q) { ps.sc[q] = W.sc[q][lane]; ps.pod[q] = W.pod[q][lane]; }
        }
        __syncwarp();
        uint32_t dm = __ballot_sync(0xffffffffu, have && !more);
        while (dm) {                                           // the whole warp writes each finished prompt's row
            const int l = __ffs(dm) - 1; dm &= dm - 1;
            const uint32_t pp = __shfl_sync(0xffffffffu, p, l);
            const uint32_t pk = __shfl_sync(0xffffffffu, k, l);
            const uint32_t pmeta = __shfl_sync(0xffffffffu, meta, l);
            if (a.dense) {
                double* row = a.dense + (long long)pp * t.max_pods;
                const uint32_t P = t.max_pods;
                if ((P & 1u) == 0 && ((reinterpret_cast<uintptr_t>(row) & 15u) == 0)) {
                    for (uint32_t c = lane * 2; c < P; c += 64) *reinterpret_cast<double2*>(row + c) = make_double2(-1.0, -1.0);
                } else {
                    for (uint32_t c = lane; c < P; c += 32) row[c] = -1.0;
                }
                __syncwarp();
                if ((uint32_t)lane < pk) { const uint32_t pd = W.pod[lane][l]; if (pd < P) row[pd] = W.sc[lane][l]; }
            }
            if (a.sp_cnt) {
                if ((uint32_t)lane < pk) { a.sp_pods[(long long)pp * kMaxEnt + lane] = W.pod[lane][l]; a.sp_scores[(long long)pp * kMaxEnt + lane] = W.sc[lane][l]; }
                if (lane == 0) a.sp_cnt[pp] = (uint8_t)pk;
            }
            if (a.has_keys && lane == 0) a.has_keys[pp] = (round > 0) || (pmeta & 63u) > 0;
        }
        __syncwarp();
    }
}

// ---- kernel R: followers take their representative's outcome --------------------------------------------------
// Lane per follower.  Representative continues: so does the follower, with the representative's chain state, and its
// walk state stays where the representative left it (src).  Representative finished: same pods, same scores -- the warp
// writes the follower's result from the representative's final state.
__global__ void __launch_bounds__(256)
resolve_round_kernel(const TableView t, const ScoreArgs a, const RoundBufs rb, const int cur, const int round) {
    const int lane = threadIdx.x & 31;
    const unsigned int n_fl = rb.n_hl[1];
    const PromptState* pst_cur = rb.pst[round & 1];
    const unsigned int total_warps = gridDim.x * (blockDim.x / 32);
    for (unsigned int w = blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5); w * 32u < n_fl; w += total_warps) {
        const unsigned int f = w * 32u + lane;
        const bool have = f < n_fl;
        uint32_t p = 0, pl = 0; uint8_t ft = 0;
        if (have) {
            const unsigned int li = rb.fl[f], lj = rb.role[li];
            p = rb.act[cur][li]; pl = rb.act[cur][lj]; ft = rb.fate[lj];
        }
        const bool more = have && ft == kFateMore;
        const uint32_t mm = __ballot_sync(0xffffffffu, more);
        if (mm) {
            unsigned int base = 0;
            if (lane == 0) base = atomicAdd(&rb.n_act[cur ^ 1], (unsigned int)__popc(mm));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (more) {
                rb.act[cur ^ 1][base + __popc(mm & ((1u << lane) - 1u))] = p;
                rb.src[p] = pl;
                rb.hstate[p] = rb.hstate[pl];
            }
        }
        uint32_t dm = __ballot_sync(0xffffffffu, have && !more);
        while (dm) {
            const int l = __ffs(dm) - 1; dm &= dm - 1;
            const uint32_t pp = __shfl_sync(0xffffffffu, p, l);
            const PromptState& ps = pst_cur[__shfl_sync(0xffffffffu, pl, l)];
            const uint32_t pk = ps.k;
            if (a.dense) {
                double* row = a.dense + (long long)pp * t.max_pods;
                const uint32_t P = t.max_pods;
                if ((P & 1u) == 0 && ((reinterpret_cast<uintptr_t>(row) & 15u) == 0)) {
                    for (uint32_t c = lane * 2; c < P; c += 64) *reinterpret_cast<double2*>(row + c) = make_double2(-1.0, -1.0);
                } else {
                    for (uint32_t c = lane; c < P; c += 32) row[c] = -1.0;
                }
                __syncwarp();
                if ((uint32_t)lane < pk) { const uint32_t pd = ps.pod[lane]; if (pd < P) row[pd] = ps.sc[lane]; }
            }
            if (a.sp_cnt) {
                if ((uint32_t)lane < pk) { a.sp_pods[(long long)pp * kMaxEnt + lane] = ps.pod[lane]; a.sp_scores[(long long)pp * kMaxEnt + lane] = ps.sc[lane]; }
                if (lane == 0) a.sp_cnt[pp] = (uint8_t)pk;
            }
            if (a.has_keys && lane == 0) a.has_keys[pp] = 1;     // followers always have a block in the round
        }
        __syncwarp();
    }
}

// List setup + a 64-bit fingerprint of every prompt's first block.  The batch is then radix-sorted by fingerprint so
// that prompts sharing a prefix sit in neighbouring lanes: their probes are the same 64-byte segments, which the
// load unit merges within a warp and L2 serves across warps.  (Any order gives the same results; this one lets the
// prefix sharing the system exists for -- system prompts, shared documents -- show up as memory locality.)
struct PartSizes { unsigned int n[kMaxParts]; };
__global__ void rounds_init_kernel(const ScoreArgs a, uint32_t block_size, unsigned long long* max_blocks, uint64_t* fp, uint32_t* idx,
                                   unsigned int* cnt, const PartSizes ps) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    unsigned long long nb = 0;
    if (i < a.n_prompts) {
        const int64_t b = a.tok_off[i] - a.tok_base, e = a.tok_off[i + 1] - a.tok_base;
        nb = (unsigned long long)((e - b) / block_size);
        uint64_t f = ~0ull;                                   // prompts without a full block sort last
        if (nb > 0) {
            f = 0x9E3779B97F4A7C15ull;
            const uint32_t* tk = a.tok + b;
            for (uint32_t j = 0; j < block_size && j < 16u; ++j) f = mix64(f ^ __ldg(tk + j));
            f &= ~(1ull << 63);
        }
        fp[i] = f; idx[i] = (uint32_t)i;
    }
    nb = __reduce_max_sync(0xffffffffu, (unsigned)min(nb, 0xffffffffull));
    if ((threadIdx.x & 31) == 0 && nb) atomicMax(max_blocks, nb);
    if (i == 0) for (int q = 0; q < kMaxParts; ++q) { cnt[4 * q] = ps.n[q]; cnt[4 * q + 1] = 0; }      // live-list lengths per part
}

inline int rounds_init() {
    if (cudaFuncSetAttribute(hash_round_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(HashSmem<16>)) != cudaSuccess) return -1;
    if (cudaFuncSetAttribute(probe_round_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(WalkSmem)) != cudaSuccess) return -1;
    return 0;
}

}  // namespace kvx
