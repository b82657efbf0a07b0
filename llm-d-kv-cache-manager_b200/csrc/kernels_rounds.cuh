// kernels_rounds.cuh -- Score() for LARGE batches: rounds over PREFIX CLASSES.
//
// Same reference path as kernels_score.cuh (GetPodScores steps 2-4, pkg/kvcache/indexer.go:141-163): chain keys
// (token_processor.go:94-162), Lookup (in_memory.go:105-146), LongestPrefixScorer.Score (kvblock_scorer.go:108-151).
//
// A large batch repeats itself: thousands of prompts share a system prompt or a document, and the per-prompt kernels
// hash and probe that shared prefix once per prompt.  Here the batch is processed in rounds of kRoundBlocks blocks, and
// in every round the live prompts are partitioned into CLASSES -- same walk state, same model and filter, and a
// token-for-token identical next chunk.  One representative per class is hashed and walked; the other members take its
// outcome.  Class membership is never decided by a hash: every member has been compared with its class's chunk, so the
// results are bit-identical to the per-prompt kernels and the oracle for any input (tests run all paths).
//
//   G   group_round_kernel   stream every live prompt's chunk once, find the classes (compare with the running class's
//                            anchor chunk; election through a small map for the rest)
//   G2  group_lists_kernel   prompts that leave a class in the middle of the chunk become partial followers (class, d);
//                            compact the lists: representatives / followers / partial followers
//   H   hash_round_kernel    FNV-64a / CBOR chain keys of the representatives' chunks
//   P   walk_round_kernel    warp per representative: 32 probes at once, ordered scoring of the hits, run snapshots
//   R+D finish_round_kernel  followers copy their representative's fate; partial followers resume from its snapshot and
//                            walk their own blocks; next round's live list
//
// The per-round cost is a fixed ~0.2 ms of short latency-bound kernels (hidden by running the batch as independent parts
// on separate streams), so medium batches use kernels_rounds_plain.cuh instead (see kvidx.cu: launch_score).
#pragma once
#include <cuda_runtime.h>
#include "kernels_score.cuh"

namespace kvx {

constexpr int kRoundBlocks = 32;         // blocks hashed per prompt per round (== lanes per warp in kernel P)
#ifndef KVIDX_HASH_STAGES
#define KVIDX_HASH_STAGES 2
#define KVIDX_HASH_THREADS 256
#endif
constexpr int kHashThreads = KVIDX_HASH_THREADS;
constexpr int kHashStages = KVIDX_HASH_STAGES;   // token blocks staged per prompt.  (4 stages x 4 warps was measured: slower -- what a launch of this kernel
                                         // waits for is room on an SM beside kernel G's CTAs, so the smaller footprint per warp wins: scripts/ab_step.py)
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

constexpr int kMaxParts = 16;             // the sorted batch runs as up to this many independent parts on their own streams

struct RoundBufs {
    uint32_t* act[2];                    // live prompt lists (ping-pong)
    unsigned int* n_act;                 // [2] list lengths
    uint64_t* hstate;                    // per prompt: chain hash after the last block of the previous round
    uint32_t* pos;                       // per prompt: first block of its next chunk (32 x round unless it walked part of a chunk alone)
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
    uint32_t* dl;                        // [n_act] partial-follower list (live slots), compacted
    unsigned int* n_hl;                  // [0] representatives, [1] followers, [2] partial followers (zeroed before every round)
    // partial followers: a prompt alone in its class whose chunk starts like a neighbouring class's chunk (it leaves a
    // popular prefix inside this chunk) shares the first dmin blocks with that class -- keys and walk -- and continues on
    // its own from there (detach_round_kernel)
    uint8_t* nfol;                       // [n_act] live slot has followers (it must stay a representative); zeroed per round
    uint8_t* need_snap;                  // [n_act] representative must record its walk state after every block; zeroed per round
    uint32_t* anch;                      // [n_act] live slot of the representative it could share a chunk prefix with
    uint8_t* dmin;                       // [n_act] ... and the number of leading blocks it shares (0: none)
    uint32_t* apos;                      // [n_act] live slot -> position in the representative list
    uint32_t* lslot;                     // per prompt: its slot in the current live list (written when the list is built)
    uint4* rec;                          // [n_hl] per representative {prompt, live slot, nbr word, src}: one load instead of a chain
    uint4* drec;                         // [n_dl] per partial follower {prompt, representative's live slot, shared blocks << 16 | nbr word, representative's prompt}
    uint32_t* grp;                       // [n_act] by live slot of last round's representative R: a class of this round whose members
                                         // had R's state (kRoleSelf: none); cleared with the map
    uint8_t* nwalk;                      // [n_hl] blocks of the round after which the representative's walk was still alive
    // walk snapshots, one per RUN of blocks with the same slot pattern: state after the run's first block; the state after a
    // later block of the run is that plus (distance) in-order additions of the same per-pod addend
    uint8_t* snap_run;                   // [n_hl][32] first block of the run block j belongs to
    uint16_t* snap_alive;                // [n_hl][32] alive mask after block j (stored at run starts)
    uint8_t* snap_bt;                    // [n_hl][32][kMaxEnt] tier that gives pod q its addend (0xff: 0.0) (at run starts)
    double* snap_sc;                     // [n_hl][32][kMaxEnt] scores after block j (at run starts)
    uint32_t part_size;                  // prompts in this part (rounded up to 32): length of the per-slot arrays
    uint32_t* map;                       // class election: bucket -> live slot (kRoleSelf = empty), cleared before every round
    uint32_t map_mask;
};
constexpr uint8_t kFateMore = 1, kFateDone = 2;

constexpr int kHashChunk = 1;            // blocks staged per copy step (2 = 128-byte accesses was measured: fewer, longer DRAM
                                         // accesses but only 24 resident warps/SM -> 6 % slower; the kernel is pipe bound)
template <int BS> struct HashSmem {
    static constexpr int kRow = kHashChunk * BS * 4 + 16;   // 144 B: the four LDS.128 of a block stay conflict free
    unsigned char tok[kHashThreads / 32][kHashStages][32 * kRow];     // [warps of the CTA]: a launch with fewer warps passes that much less
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

#ifndef KVIDX_GROUP_TILE
#define KVIDX_GROUP_TILE 32
#endif
constexpr int kGroupTile = KVIDX_GROUP_TILE;   // live prompts a warp of kernel G takes at a time (their chunks stream through the warp one after the other)
#ifndef KVIDX_GROUP_RING
#define KVIDX_GROUP_RING 4
#endif
constexpr int kGroupRing = KVIDX_GROUP_RING;            // chunk slots per warp: one being examined, two in flight, the anchor (deeper rings measured: slower)
struct GroupSmem {
    uint4 ring[kGroupThreads / 32][kGroupRing][4][32];                       // 64 KB
    unsigned long long bar[kGroupThreads / 32][kGroupRing];                  // TMA variant: one mbarrier per rotating slot
};

// TMA = true: a chunk (<= 2 KB, contiguous, 16-byte aligned) is ONE cp.async.bulk issued by an elected lane onto the slot's
// mbarrier (UBLKCP in SASS) instead of four 16-byte cp.async per lane (LDGSTS); chunks that do not start on a 16-byte
// boundary are copied with plain loads in both variants.
template <int BS, bool TMA>
__global__ void __launch_bounds__(kGroupThreads, kGroupRing <= 4 ? 3 : 2)
group_round_kernel(const TableView t, const ScoreArgs a, const RoundBufs rb, const int cur, const int round, const int dedup) {
    static_assert(BS == 16 && kRoundBlocks == 32, "a chunk is 4 x 32 lanes x 16 bytes");
    extern __shared__ __align__(128) unsigned char smem_raw_g[];
    uint4 (*ring)[4][32] = reinterpret_cast<GroupSmem*>(smem_raw_g)->ring[threadIdx.x >> 5];
    unsigned long long* bar = reinterpret_cast<GroupSmem*>(smem_raw_g)->bar[threadIdx.x >> 5];
    const int lane = threadIdx.x & 31;
    uint32_t tma_pending = 0, tma_phase = 0;          // per rotating slot: a bulk copy is in flight / parity of its next wait
    const uint64_t l2pol = l2_policy_stream();
    if (TMA) {
        if (lane == 0) for (int s = 0; s < kGroupRing - 1; ++s) mbar_init(&bar[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        fence_proxy_async();
        __syncwarp();
    }
    const unsigned int n_act = rb.n_act[cur];
    if (blockIdx.x == 0 && threadIdx.x == 0) { rb.n_act[cur ^ 1] = 0; rb.n_hl[0] = 0; rb.n_hl[1] = 0; rb.n_hl[2] = 0; rb.n_hl[3] = 0; }   // lists built below / by G2
    const unsigned int total_warps = gridDim.x * (kGroupThreads / 32);
    for (unsigned int w = blockIdx.x * (kGroupThreads / 32) + (threadIdx.x >> 5); w * (unsigned)kGroupTile < n_act; w += total_warps) {
        const unsigned int i = w * (unsigned)kGroupTile + lane;
        const bool have = lane < kGroupTile && i < n_act;
        uint32_t p = 0, srcp = kRoleSelf, mdl = a.model0;
        int nb = 0; bool more = false;
        uint64_t hprev = t.init_hash;
        const uint32_t* tp = a.tok;
        int64_t first = 0;
        if (have) {
            p = rb.act[cur][i];
            if (round > 0) first = rb.pos[p];
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
        uint64_t fsum = 0;
        if (cls && round == 0 && a.filter) { const uint64_t* fr = a.filter + (int64_t)p * t.filter_words; for (uint32_t x = 0; x < t.filter_words; ++x) fsum = (fsum ^ fr[x]) * 0x9E3779B97F4A7C15ull; }
        // The chunks stream through a 3-slot ring (one being examined, two in flight); a fourth slot keeps the ANCHOR: the chunk
        // of the class the list is currently running through (the list is sorted by prefix, so class-mates are neighbours).
        // A chunk equal to the anchor (and with the same walk state, model, filter, length) is a member of the anchor's
        // class: nothing else to do for it.  A chunk that starts like the anchor but differs is a prompt leaving that class
        // inside the chunk: it goes through the election like any new chunk, and remembers (class, shared blocks) as an exact
        // hint for kernel G2.  A chunk unrelated to the anchor -- or one that differs from an anchor nobody has joined yet --
        // becomes the anchor.
        uint32_t f0 = 0, f1 = 0;
        bool eqanch = false;                          // member of the class of lane `alane`
        int alane = -1;
        int dh = 0, hlane = -1;                       // hint: shares dh leading blocks with the chunk of lane hlane
        int anchor = -1, members = 0;                 // warp-uniform: the anchor's lane, prompts that matched it so far
        uint32_t todo = __ballot_sync(0xffffffffu, cls);
        auto issue = [&](int slot, int qi) {
            if (qi >= 0) {
                const uint32_t* sp = reinterpret_cast<const uint32_t*>(__shfl_sync(0xffffffffu, (unsigned long long)(uintptr_t)tp, qi));
                const int nw = __shfl_sync(0xffffffffu, nb, qi) * BS;
                const bool al = (reinterpret_cast<uintptr_t>(sp) & 15u) == 0;
                if (TMA && al) {
                    __syncwarp();                              // every lane has taken the slot's previous chunk into registers
                    if (lane == 0) { fence_proxy_async(); mbar_expect_tx(&bar[slot], (uint32_t)nw * 4u); tma_load_1d(&ring[slot][0][0], sp, (uint32_t)nw * 4u, &bar[slot]); }
                    tma_pending |= 1u << slot;
                }
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const int w0 = (c * 32 + lane) * 4;
                    uint4* dst = &ring[slot][c][lane];
                    if (w0 < nw) {
                        if (TMA && al) {}
                        else if (al) cp_async_16_stream(smem_addr(dst), sp + w0, l2pol);
                        else *dst = make_uint4(__ldg(sp + w0), __ldg(sp + w0 + 1), __ldg(sp + w0 + 2), __ldg(sp + w0 + 3));
                    } else *dst = make_uint4(0, 0, 0, 0);
                }
            }
            if (!TMA) cp_async_commit();
        };
        constexpr int RS = kGroupRing - 1;            // rotating slots; slot RS holds the anchor
        constexpr int PD = RS - 1;                    // chunks in flight
        uint32_t pend = todo;
        int issued = 0;
#pragma unroll
        for (int u = 0; u < PD; ++u) {
            const int qi = pend ? __ffs(pend) - 1 : -1;
            pend &= pend - 1;
            issue(issued % RS, qi); ++issued;
        }
        int ord = 0;
        while (todo) {
            const int q = __ffs(todo) - 1;
            todo &= todo - 1;
            {
                const int qi = pend ? __ffs(pend) - 1 : -1;
                pend &= pend - 1;
                issue(issued % RS, qi); ++issued;
            }
            if (TMA) {
                const int s_ = ord % RS;
                if ((tma_pending >> s_) & 1u) { mbar_wait(&bar[s_], (tma_phase >> s_) & 1u); tma_phase ^= 1u << s_; tma_pending &= ~(1u << s_); }
                __syncwarp();                                  // plain-store chunks (unaligned starts, zero tails) are visible too
            } else cp_async_wait<PD>();
            uint4 v[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) v[c] = ring[ord % RS][c][lane];
            bool same = false, compat = false;
            int fbc = 0;
            if (anchor >= 0) {
                const int nb_q = __shfl_sync(0xffffffffu, nb, q), nb_a = __shfl_sync(0xffffffffu, nb, anchor);
                compat = __shfl_sync(0xffffffffu, mdl, q) == __shfl_sync(0xffffffffu, mdl, anchor) &&
                         __shfl_sync(0xffffffffu, srcp, q) == __shfl_sync(0xffffffffu, srcp, anchor) &&
                         __shfl_sync(0xffffffffu, hprev, q) == __shfl_sync(0xffffffffu, hprev, anchor) &&
                         __shfl_sync(0xffffffffu, fsum, q) == __shfl_sync(0xffffffffu, fsum, anchor);
                if (compat && round == 0 && a.filter) {           // equal filter fingerprints: compare the rows themselves
                    const uint32_t pq = __shfl_sync(0xffffffffu, p, q), pa = __shfl_sync(0xffffffffu, p, anchor);
                    bool eqr = true;
                    for (uint32_t x = lane; x < t.filter_words; x += 32) eqr = eqr && a.filter[(int64_t)pq * t.filter_words + x] == a.filter[(int64_t)pa * t.filter_words + x];
                    compat = __all_sync(0xffffffffu, eqr);
                }
                if (compat) {
                    int fb = 32;                                  // first block in which this lane's pieces differ
#pragma unroll
                    for (int c = 3; c >= 0; --c) {
                        const uint4 u = ring[RS][c][lane];
                        if (((v[c].x ^ u.x) | (v[c].y ^ u.y) | (v[c].z ^ u.z) | (v[c].w ^ u.w)) != 0u) fb = c * 8 + (lane >> 2);
                    }
                    fb = __reduce_min_sync(0xffffffffu, fb);
                    fbc = min(fb, min(nb_q, nb_a));
                    same = fb == 32 && nb_q == nb_a && __shfl_sync(0xffffffffu, (int)more, q) == __shfl_sync(0xffffffffu, (int)more, anchor);
                }
            }
            if (same) {
                if (lane == q) { eqanch = true; alane = anchor; }
                ++members;
            } else {
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
                bool take = true;                                // does this chunk become the anchor?
                if (anchor >= 0 && compat && fbc >= 1) {
                    if (members > 0) { take = false; if (lane == q) { dh = fbc; hlane = anchor; } }      // leaves the anchor's class
                    else if (lane == anchor) { dh = fbc; hlane = q; }                                      // ... or the old anchor leaves this one's
                }
                if (take) {
#pragma unroll
                    for (int c = 0; c < 4; ++c) ring[RS][c][lane] = v[c];
                    anchor = q; members = 0;
                }
            }
            ++ord;
        }
        if (!TMA) cp_async_wait<0>();
        // election among the prompts that are not members of an anchor's class
        uint32_t cand = kRoleSelf;
        if (cls && !eqanch) {
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
            uint4 v[4], vn[4];
            chunk_load(s1, nw, lane, v);
            chunk_load(s2, nw, lane, vn);
            uint32_t d = 0;
#pragma unroll
            for (int c = 0; c < 4; ++c) d |= (v[c].x ^ vn[c].x) | (v[c].y ^ vn[c].y) | (v[c].z ^ vn[c].z) | (v[c].w ^ vn[c].w);
            const bool same = __all_sync(0xffffffffu, d == 0u);
            if (lane == l) shared = same;
        }
        // members take the outcome of their anchor (which went through the election); hints are resolved the same way
        {
            const int al_ = eqanch ? alane : lane;
            const uint32_t cand_h = __shfl_sync(0xffffffffu, cand, al_);
            const int shared_h = __shfl_sync(0xffffffffu, (int)shared, al_);
            const unsigned int i_h = __shfl_sync(0xffffffffu, i, al_);
            if (eqanch) { shared = true; cand = shared_h ? cand_h : i_h; }
        }
        if (have) { rb.role[i] = shared ? cand : kRoleSelf; if (shared) rb.nfol[cand] = 1; }
        // Every member of a class tells the prompts that shared its state last round (same src) where the class lives:
        // kernel G2 points the ones that stayed alone at this class.
        if (shared && round > 0) rb.grp[rb.lslot[srcp]] = cand;
        {
            const int hl_ = hlane >= 0 ? hlane : lane;
            const uint32_t cand_h = __shfl_sync(0xffffffffu, cand, hl_);
            const int shared_h = __shfl_sync(0xffffffffu, (int)shared, hl_);
            const unsigned int i_h = __shfl_sync(0xffffffffu, i, hl_);
            if (have) {
                const bool hint = cls && !shared && hlane >= 0 && dh > 0;
                rb.anch[i] = hint ? (shared_h ? cand_h : i_h) : kRoleSelf;
                rb.dmin[i] = (uint8_t)(hint ? (dh | 0x80) : 0);     // exact: measured against a chunk of that very class
            }
        }
    }
}

// ---- kernel G2: partial followers, and the three lists -----------------------------------------------------------
// A prompt that ended up alone in its class and that nobody follows looks for a class whose chunk starts like its own:
// the class that prompts with its state joined this round (grp), else last round's representative itself, else -- round
// 0, or nothing better -- the class of a list neighbour (kernel G's hint).  The shared prefix is then measured exactly,
// token by token against that class's chunk (warp-wide), so that the prompt detaches at the right block.
template <int BS>
__global__ void __launch_bounds__(256)
group_lists_kernel(const TableView t, const ScoreArgs a, const RoundBufs rb, const int cur, const int round, const int partial) {
    const int lane = threadIdx.x & 31;
    const unsigned int n_act = rb.n_act[cur];
    const unsigned int total_warps = gridDim.x * (blockDim.x / 32);
    for (unsigned int w = blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5); w * 32u < n_act; w += total_warps) {
        const unsigned int i = w * 32u + lane;
        const bool have = i < n_act;
        int kind = -1;                               // 0 representative / alone, 1 follower, 2 partial follower
        uint32_t target = kRoleSelf;
        const uint32_t* tp = a.tok; const uint32_t* rp = a.tok;
        int nbc = 0;                                 // blocks both chunks have
        int dknown = 0;
        if (have) {
            kind = rb.role[i] == kRoleSelf ? 0 : 1;
            const int nb = (int)(rb.nbr[i] & 63u);
            if (kind == 0 && partial && nb > 0 && !rb.nfol[i]) {
                const uint32_t p = rb.act[cur][i];
                // (Last round's representatives never become partial followers: the prompts that shared their state may be
                //  attaching to them right now.)
                bool allow = true;
                if (round > 0) {
                    const uint32_t srcp = rb.src[p];
                    allow = srcp != p;
                    if (allow) { const uint32_t rs = rb.lslot[srcp]; const uint32_t g = rb.grp[rs]; target = g != kRoleSelf ? g : rs; }
                }
                const uint32_t hint = rb.dmin[i];                  // kernel G: shared blocks with a neighbour's class | 0x80 if exact
                // a hinted class is only safe to lean on if it has followers (then its representative stays one)
                if (allow && target == kRoleSelf && (hint & 0x7fu) > 0 && rb.anch[i] != kRoleSelf && rb.nfol[rb.anch[i]]) target = rb.anch[i];
                if (target != kRoleSelf && rb.role[target] != kRoleSelf) target = rb.role[target];   // someone with the same chunk represents it
                if (target == i) target = kRoleSelf;
                if (target != kRoleSelf) {
                    const uint32_t pl = rb.act[cur][target];
                    const int64_t first = round > 0 ? (int64_t)rb.pos[p] : 0;      // the class's too: same state
                    const int64_t b = a.tok_off[p] - a.tok_base;
                    const int64_t bl = a.tok_off[pl] - a.tok_base, el = a.tok_off[pl + 1] - a.tok_base;
                    nbc = min(nb, (int)max((int64_t)0, min((int64_t)kRoundBlocks, (el - bl) / BS - first)));
                    tp = a.tok + b + first * BS; rp = a.tok + bl + first * BS;
                    // same walk state, model and filter as the class?  (grp / src: by construction; the neighbour hint: checked by kernel G)
                    if (round > 0 && !(rb.src[pl] == rb.src[p] && rb.hstate[pl] == rb.hstate[p])) nbc = 0;
                    // kernel G compared this prompt with a member of exactly this class already: no need to read the chunks again
                    if (nbc > 0 && (hint & 0x80u) && rb.anch[i] == target) { dknown = min((int)(hint & 0x7fu), nbc); nbc = 0; }
                }
            }
        }
        int dshare = dknown;
        uint32_t wm = __ballot_sync(0xffffffffu, nbc > 0);
        while (wm) {
            const int l = __ffs(wm) - 1; wm &= wm - 1;
            const uint32_t* s1 = reinterpret_cast<const uint32_t*>(__shfl_sync(0xffffffffu, (unsigned long long)(uintptr_t)tp, l));
            const uint32_t* s2 = reinterpret_cast<const uint32_t*>(__shfl_sync(0xffffffffu, (unsigned long long)(uintptr_t)rp, l));
            const int nw = __shfl_sync(0xffffffffu, nbc, l) * BS;
            uint4 v[4], vn[4];
            chunk_load(s1, nw, lane, v);
            chunk_load(s2, nw, lane, vn);
            int fb = 32;
#pragma unroll
            for (int c = 3; c >= 0; --c)
                if (((v[c].x ^ vn[c].x) | (v[c].y ^ vn[c].y) | (v[c].z ^ vn[c].z) | (v[c].w ^ vn[c].w)) != 0u) fb = c * 8 + (lane >> 2);
            fb = __reduce_min_sync(0xffffffffu, fb);
            if (lane == l) dshare = min(fb, nw / BS);
        }
        if (dshare > 0) { kind = 2; rb.anch[i] = target; rb.dmin[i] = (uint8_t)dshare; rb.need_snap[target] = 1; }
        const uint32_t m0 = __ballot_sync(0xffffffffu, kind == 0), m1 = __ballot_sync(0xffffffffu, kind == 1), m2 = __ballot_sync(0xffffffffu, kind == 2);
        unsigned int b0 = 0, b1 = 0, b2 = 0;
        if (lane == 0) {
            if (m0) b0 = atomicAdd(&rb.n_hl[0], (unsigned int)__popc(m0));
            if (m1) b1 = atomicAdd(&rb.n_hl[1], (unsigned int)__popc(m1));
            if (m2) b2 = atomicAdd(&rb.n_hl[2], (unsigned int)__popc(m2));
        }
        b0 = __shfl_sync(0xffffffffu, b0, 0); b1 = __shfl_sync(0xffffffffu, b1, 0); b2 = __shfl_sync(0xffffffffu, b2, 0);
        const uint32_t below = (1u << lane) - 1u;
        if (kind == 0) {
            const unsigned int ap = b0 + __popc(m0 & below);
            const uint32_t p = rb.act[cur][i];
            rb.hl[ap] = i; rb.apos[i] = ap;
            rb.rec[ap] = make_uint4(p, i, rb.nbr[i], round > 0 ? rb.src[p] : p);
        } else if (kind == 1) rb.fl[b1 + __popc(m1 & below)] = i;
        else if (kind == 2) {
            const unsigned int dp = b2 + __popc(m2 & below);
            rb.dl[dp] = i;
            rb.drec[dp] = make_uint4(rb.act[cur][i], target, ((uint32_t)dshare << 16) | rb.nbr[i], rb.act[cur][target]);
        }
    }
}

// ---- kernel H ---------------------------------------------------------------------------------
template <int BS>
__global__ void __launch_bounds__(kHashThreads, 3)
hash_round_kernel(const TableView t, const ScoreArgs a, const RoundBufs rb, const int cur, const int round, const int prefetch) {
    static_assert(BS == 16, "staging pattern is written for 16-token blocks");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    using SM = HashSmem<BS>;
    SM& sm = *reinterpret_cast<SM*>(smem_raw);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const unsigned int n_hl = rb.n_hl[0];                                 // representatives this round (kernel G)
    const uint64_t l2pol = l2_policy_stream();
    {   // the election map, grp and nfol are dead until the next round's kernel G: clear them here instead of three memsets
        uint4* m4 = reinterpret_cast<uint4*>(rb.map);
        const size_t nm = ((size_t)rb.map_mask + 1 + rb.part_size) / 4;   // map + grp are contiguous; sizes are multiples of 4 words
        for (size_t x = blockIdx.x * (size_t)blockDim.x + threadIdx.x; x < nm; x += (size_t)gridDim.x * blockDim.x) m4[x] = make_uint4(~0u, ~0u, ~0u, ~0u);
        uint32_t* f4 = reinterpret_cast<uint32_t*>(rb.nfol);
        for (size_t x = blockIdx.x * (size_t)blockDim.x + threadIdx.x; x < rb.part_size / 4; x += (size_t)gridDim.x * blockDim.x) f4[x] = 0u;
    }
    const unsigned int total_warps = gridDim.x * (blockDim.x / 32);
    for (unsigned int w = blockIdx.x * (blockDim.x / 32) + wid; w * 32u < n_hl; w += total_warps) {
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
            const int64_t first = round > 0 ? (int64_t)rb.pos[p] : 0;
            nb = (int)max((int64_t)0, min((int64_t)kRoundBlocks, nblk - first));
            src = a.tok + b + first * BS;
            aligned = (reinterpret_cast<uintptr_t>(src) & 15u) == 0;
            h = round == 0 ? t.init_hash : rb.hstate[p];
        }
        const int nb_max = __reduce_max_sync(0xffffffffu, nb);
        // The chain below consumes one 64-byte block per prompt per ~0.6 us and stages only one block ahead: pull the whole
        // chunk (<= 2 KB per prompt) towards L2 now, so that the staging copies find it there instead of paying a loaded DRAM
        // round trip per block.
        if (prefetch && have) {
            const char* pb = reinterpret_cast<const char*>(src);
            for (int o = 128; o < nb * BS * 4; o += 128) {
                if (prefetch == 2) asm volatile("prefetch.global.L2::evict_last [%0];" ::"l"(pb + o));
                else asm volatile("prefetch.global.L2 [%0];" ::"l"(pb + o));
            }
        }
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
                    cp_async_16_stream(smem_addr(&sm.tok[wid][s][q * SM::kRow + (lane % LPP) * 16]), reinterpret_cast<const char*>(sp) + (lane % LPP) * 16, l2pol);
            }
            cp_async_commit();
            if (issue && !aligned) {
                uint32_t* dst = reinterpret_cast<uint32_t*>(&sm.tok[wid][s][lane * SM::kRow]);
                const uint32_t* g = src + (size_t)b0 * BS;
                for (int j = 0; j < nbytes / 4; ++j) dst[j] = __ldg(g + j);
            }
        };
#pragma unroll
        for (int c = 0; c < kHashStages - 1; ++c) stage(c, c);
        const int nchunks = (nb_max + kHashChunk - 1) / kHashChunk;
        for (int c = 0; c < nchunks; ++c) {
            stage((c + kHashStages - 1) % kHashStages, c + kHashStages - 1);
            cp_async_wait<kHashStages - 1>();
            __syncwarp();
#pragma unroll
            for (int u = 0; u < kHashChunk; ++u) {
                const int b = c * kHashChunk + u;
                Fnv f;
                f.begin_block(h, BS);
                const uint4* tp = reinterpret_cast<const uint4*>(&sm.tok[wid][c % kHashStages][lane * SM::kRow + u * BS * 4]);
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

// ---- helpers shared with kernels_rounds_plain.cuh ---------------------------------------------
__device__ __forceinline__ void ld_slot_pair(const ReqSlot* s, bool peer, uint4& a0, uint4& b0, uint4& a1, uint4& b1) {
    ld_slot(s, peer, a0, b0);
    ld_slot(s + 1, peer, a1, b1);
}

// entry j (0..9) of a slot whose words live in this lane's registers
__device__ __forceinline__ uint32_t ent_of(uint32_t e0, uint32_t e1, uint32_t e2, uint32_t e3, uint32_t e4, int j) {
    const uint32_t word = j < 2 ? e0 : j < 4 ? e1 : j < 6 ? e2 : j < 8 ? e3 : e4;
    return (j & 1) ? (word >> 16) : (word & 0xffffu);
}

// ---- kernel P: warp per representative ------------------------------------------------------------------------
// With prefix classes the representatives of a round are few (one per distinct prefix), so a lane-per-prompt walk
// (kernels_rounds_plain.cuh: 32 dependent iterations of divergent code per warp) would leave the machine idle: its time
// is a latency chain that does not shrink with the list.  Here a warp takes one representative: the 32 lanes probe the
// round's 32 keys at once (one DRAM latency for the whole chunk), a ballot finds the first miss, and the blocks before
// it are scored in order with lane q owning pod q of the first block (<= 10 pods), the block's entries broadcast by
// shuffle.  Same arithmetic in the same order as everywhere else: max weight per pod per block, added in block order.
__global__ void __launch_bounds__(256)
walk_round_kernel(const TableView t, const ScoreArgs a, const RoundBufs rb, const int cur, const int round) {
    const int lane = threadIdx.x & 31;
    const unsigned int n_a = rb.n_hl[0];
    const uint64_t l2pol = l2_policy_stream();
    const size_t kstride = (size_t)a.n_prompts;
    const bool peer = t.shard_bits != 0;
    const PromptState* pst_prev = rb.pst[(round & 1) ^ 1];
    PromptState* pst_cur = rb.pst[round & 1];
    const unsigned int total_warps = gridDim.x * (blockDim.x / 32);
    for (unsigned int i = blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5); i < n_a; i += total_warps) {
        const uint4 rc = __ldg(rb.rec + i);
        uint64_t key = rb.keys[(size_t)lane * kstride + i];   // lanes >= nb read a stale key and ignore it
        const unsigned int li = rc.y;
        const uint32_t p = rc.x;
        const uint32_t meta = rc.z;
        const int nb = (int)(meta & 63u);
        const bool has_more = (meta >> 8) & 1u;
        const uint32_t mdl = a.model ? a.model[p] : a.model0;
        const bool snapf = rb.need_snap[li];
        // walk state: lane q < k owns pod q
        uint32_t k = 0, alive = 0, mypod = 0xffffffffu, mybt = 0xffu;
        double mysc = 0.0;
        uint32_t pv0 = 0, pv1 = 0, pv2 = 0, pv3 = 0, pv4 = 0, pvc = 0xffffffffu;
        if (round > 0) {
            const PromptState& ps = pst_prev[rc.w];
            k = ps.k; alive = ps.alive;
            pv0 = ps.pat[0]; pv1 = ps.pat[1]; pv2 = ps.pat[2]; pv3 = ps.pat[3]; pv4 = ps.pat[4]; pvc = ps.pat[5];
            if ((uint32_t)lane < k) { mysc = ps.sc[lane]; mypod = ps.pod[lane]; mybt = ps.bt[lane]; }
        }
        // probe: lane j looks up block j's key
        uint32_t e0 = 0, e1 = 0, e2 = 0, e3 = 0, e4 = 0, cnt = 0;
        bool hit = false;
        if (lane < nb) {
            const uint64_t hm = home_of(key, mdl);
            const ReqSlot* base = t.req_peer[shard_of(hm, t.shard_bits)];
            uint64_t slot = hm & t.req_mask & ~1ull;
            // A local probe fetches the aligned slot PAIR (one 64-byte DRAM access).  On a peer GPU's shard every byte crosses
            // NVLink, and at load <= 0.25 the home slot alone decides ~9 probes out of 10: fetch it first, its neighbour only
            // if needed (halves the NVLink bytes of a sharded step, profiles/r2_sharded_nvlink.json).
            const bool remote = base != t.req && !t.peer_pair;
            for (;;) {
                uint4 A0, B0, A1, B1;
                ld_slot(base + slot, peer, A0, B0);
                if (!remote) ld_slot(base + slot + 1, peer, A1, B1);
                uint4 A = A0, B = B0;
                hit = slot_matches(A, B, key, mdl);
                bool stop = hit || meta_state(B.w) == kStateEmpty;
                if (!stop) {
                    if (remote) ld_slot(base + slot + 1, peer, A1, B1);
                    A = A1; B = B1; hit = slot_matches(A, B, key, mdl); stop = hit || meta_state(B.w) == kStateEmpty;
                }
                if (hit) { e0 = A.z; e1 = A.w; e2 = B.x; e3 = B.y; e4 = B.z; cnt = meta_count(B.w); }
                if (stop) break;
                slot = (slot + 2) & t.req_mask;                         // rare: displaced past the home pair
            }
        }
        const uint32_t hm_ = __ballot_sync(0xffffffffu, hit);
        const int nhit = hm_ == 0xffffffffu ? 32 : __ffs(~hm_) - 1;   // consecutive hits from block 0 (lanes >= nb never hit)
        uint32_t nwalk = 0;
        bool done = nb == 0;
        // blocks whose slot holds the same pods and tiers as the block before: every live pod gets the same addend again,
        // so a run of them is a run of in-order additions (most documents live on the same pods block after block)
        uint32_t samemask;
        {
            const uint32_t u0 = __shfl_up_sync(0xffffffffu, e0, 1), u1 = __shfl_up_sync(0xffffffffu, e1, 1), u2 = __shfl_up_sync(0xffffffffu, e2, 1),
                           u3 = __shfl_up_sync(0xffffffffu, e3, 1), u4 = __shfl_up_sync(0xffffffffu, e4, 1), uc = __shfl_up_sync(0xffffffffu, cnt, 1);
            bool sm_ = lane > 0 ? (((u0 ^ e0) | (u1 ^ e1) | (u2 ^ e2) | (u3 ^ e3) | (u4 ^ e4) | (uc ^ cnt)) == 0u)
                                : (round > 0 && ((pv0 ^ e0) | (pv1 ^ e1) | (pv2 ^ e2) | (pv3 ^ e3) | (pv4 ^ e4) | (pvc ^ cnt)) == 0u);
            samemask = __ballot_sync(0xffffffffu, sm_ && lane < nhit);
        }
        int j = 0;
        while (j < nhit && !done) {
            if ((samemask >> j) & 1u) {
                const uint32_t rest = ~(samemask >> j);
                const int run = min(rest ? __ffs(rest) - 1 : 32, nhit - j);           // >= 1
                const bool mine = (alive >> lane) & 1u;
                const double add = mybt == 0xffu ? 0.0 : t.weight[mybt & 15u];
                int u = 0;
                if (j == 0) {                                  // the chunk opens inside a run: its first block is the snapshot
                    if (mine) mysc = __dadd_rn(mysc, add);
                    if (snapf) {
                        if ((uint32_t)lane < k) { rb.snap_sc[(size_t)i * kRoundBlocks * kMaxEnt + lane] = mysc; rb.snap_bt[(size_t)i * kRoundBlocks * kMaxEnt + lane] = (uint8_t)mybt; }
                        if (lane == 0) rb.snap_alive[(size_t)i * kRoundBlocks] = (uint16_t)alive;
                    }
                    u = 1;
                }
                if (mine) for (; u < run; ++u) mysc = __dadd_rn(mysc, add);
                j += run;
                nwalk = (uint32_t)j;
                continue;
            }
            const uint32_t w0 = __shfl_sync(0xffffffffu, e0, j), w1 = __shfl_sync(0xffffffffu, e1, j), w2 = __shfl_sync(0xffffffffu, e2, j),
                           w3 = __shfl_sync(0xffffffffu, e3, j), w4 = __shfl_sync(0xffffffffu, e4, j), c = __shfl_sync(0xffffffffu, cnt, j);
            const bool first_block = round == 0 && j == 0;
            if (first_block) {
                // activePods := pods of block 0 (after the filter), in entry order; score = max weight   (kvblock_scorer.go:118-128)
                const uint64_t* frow = filter_row(a.filter, p, t.filter_words);
                k = 0;
                for (uint32_t e = 0; e < c; ++e) {
                    const uint32_t pt = ent_of(w0, w1, w2, w3, w4, (int)e), pd = pt >> 4;
                    if (frow && !filter_has(frow, pd)) continue;
                    const double wt = t.weight[pt & 15u];
                    const uint32_t own = __ballot_sync(0xffffffffu, (uint32_t)lane < k && mypod == pd);
                    const int q = own ? __ffs(own) - 1 : (int)k;
                    if (!own) { if (lane == q) { mypod = pd; mysc = 0.0; mybt = 0xffu; } ++k; }
                    if (lane == q && wt > mysc) { mysc = wt; mybt = pt & 15u; }
                }
                alive = (1u << k) - 1u;
            } else {
                // activePods &= pods(block); score[p] += max weight, in block order   (kvblock_scorer.go:130-147)
                bool present = false; double mx = 0.0; uint32_t bt = 0xffu;
                if ((alive >> lane) & 1u) {
                    for (uint32_t e = 0; e < c; ++e) {
                        const uint32_t pt = ent_of(w0, w1, w2, w3, w4, (int)e);
                        if ((pt >> 4) == mypod) { present = true; const double wt = t.weight[pt & 15u]; if (wt > mx) { mx = wt; bt = pt & 15u; } }
                    }
                    if (present) { mysc = __dadd_rn(mysc, mx); mybt = bt; }
                }
                alive = __ballot_sync(0xffffffffu, present);
            }
            if (!alive) done = true;
            else {
                nwalk = (uint32_t)j + 1u;
                if (snapf) {
                    if ((uint32_t)lane < k) { rb.snap_sc[((size_t)i * kRoundBlocks + j) * kMaxEnt + lane] = mysc; rb.snap_bt[((size_t)i * kRoundBlocks + j) * kMaxEnt + lane] = (uint8_t)mybt; }
                    if (lane == 0) rb.snap_alive[(size_t)i * kRoundBlocks + j] = (uint16_t)alive;
                }
            }
            ++j;
        }
        if (snapf) {                                           // block -> first block of its run
            const uint32_t starts = ~samemask & ((2u << lane) - 1u);
            rb.snap_run[(size_t)i * kRoundBlocks + lane] = (uint8_t)(starts ? 31 - __clz(starts) : 0);
        }
        // pattern of the last scored block, for the next round's first comparison
        if (j > 0) {
            const int lj = j - 1;
            pv0 = __shfl_sync(0xffffffffu, e0, lj); pv1 = __shfl_sync(0xffffffffu, e1, lj); pv2 = __shfl_sync(0xffffffffu, e2, lj);
            pv3 = __shfl_sync(0xffffffffu, e3, lj); pv4 = __shfl_sync(0xffffffffu, e4, lj); pvc = __shfl_sync(0xffffffffu, cnt, lj);
        }
        if (nhit < nb) done = true;                            // a block of the round is not in the index
        const bool more = !done && has_more;
        PromptState& ps = pst_cur[p];
        if ((uint32_t)lane < k) { ps.sc[lane] = mysc; ps.pod[lane] = (uint16_t)mypod; ps.bt[lane] = (uint8_t)mybt; }
        if (lane == 0) {
            ps.k = (uint8_t)k; ps.alive = (uint16_t)alive;
            ps.pat[0] = pv0; ps.pat[1] = pv1; ps.pat[2] = pv2; ps.pat[3] = pv3; ps.pat[4] = pv4; ps.pat[5] = pvc;
            rb.fate[li] = more ? kFateMore : kFateDone; rb.nwalk[i] = (uint8_t)nwalk;
            if (more) { rb.src[p] = p; rb.pos[p] = (round > 0 ? rb.pos[p] : 0u) + (uint32_t)nb; }
        }
        if (more) { if (lane == nb - 1) rb.hstate[p] = key; }   // chain state for the next round: key of the round's last block
        else {
            if (a.dense) {
                double* row = a.dense + (long long)p * t.max_pods;
                const uint32_t P = t.max_pods;
                if ((P & 1u) == 0 && ((reinterpret_cast<uintptr_t>(row) & 15u) == 0)) {
                    for (uint32_t c2 = lane * 2; c2 < P; c2 += 64) st_stream_f64x2(row + c2, -1.0, -1.0, l2pol);
                } else {
                    for (uint32_t c2 = lane; c2 < P; c2 += 32) row[c2] = -1.0;
                }
                __syncwarp();
                if ((uint32_t)lane < k && mypod < P) row[mypod] = mysc;
            }
            if (a.sp_cnt) {
                if ((uint32_t)lane < k) { a.sp_pods[(long long)p * kMaxEnt + lane] = (uint16_t)mypod; a.sp_scores[(long long)p * kMaxEnt + lane] = mysc; }
                if (lane == 0) a.sp_cnt[p] = (uint8_t)k;
            }
            if (a.has_keys && lane == 0) a.has_keys[p] = (round > 0) || nb > 0;
        }
        __syncwarp();
    }
}

// ---- kernel R: followers take their representative's outcome --------------------------------------------------
// Lane per follower.  Representative continues: so does the follower, with the representative's chain state, and its
// walk state stays where the representative left it (src).  Representative finished: same pods, same scores -- the warp
// writes the follower's result from the representative's final state.
__device__ __forceinline__ void resolve_round(const TableView& t, const ScoreArgs& a, const RoundBufs& rb, const int cur, const int round,
                                              const unsigned int bid, const unsigned int nbl, const uint64_t l2pol) {
    const int lane = threadIdx.x & 31;
    {   // kernel P is done with need_snap: clear it for the next round
        uint32_t* f4 = reinterpret_cast<uint32_t*>(rb.need_snap);
        for (size_t x = bid * (size_t)blockDim.x + threadIdx.x; x < rb.part_size / 4; x += (size_t)nbl * blockDim.x) f4[x] = 0u;
    }
    {                                                          // representatives that continue (walk_round_kernel leaves the list to us)
        const unsigned int n_a = rb.n_hl[0];
        const unsigned int tw = nbl * (blockDim.x / 32);
        for (unsigned int w = bid * (blockDim.x / 32) + (threadIdx.x >> 5); w * 32u < n_a; w += tw) {
            const unsigned int i = w * 32u + lane;
            const unsigned int li = i < n_a ? rb.hl[i] : 0u;
            const bool more = i < n_a && rb.fate[li] == kFateMore;
            const uint32_t mm = __ballot_sync(0xffffffffu, more);
            if (!mm) continue;
            unsigned int base = 0;
            if (lane == 0) base = atomicAdd(&rb.n_act[cur ^ 1], (unsigned int)__popc(mm));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (more) { const unsigned int ns = base + __popc(mm & ((1u << lane) - 1u)); const uint32_t pp = rb.act[cur][li]; rb.act[cur ^ 1][ns] = pp; rb.lslot[pp] = ns; }
        }
    }
    const unsigned int n_fl = rb.n_hl[1];
    const PromptState* pst_cur = rb.pst[round & 1];
    const unsigned int total_warps = nbl * (blockDim.x / 32);
    for (unsigned int w = bid * (blockDim.x / 32) + (threadIdx.x >> 5); w * 32u < n_fl; w += total_warps) {
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
                { const unsigned int ns = base + __popc(mm & ((1u << lane) - 1u)); rb.act[cur ^ 1][ns] = p; rb.lslot[p] = ns; }
                rb.src[p] = pl;
                rb.hstate[p] = rb.hstate[pl]; rb.pos[p] = rb.pos[pl];
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
                    for (uint32_t c = lane * 2; c < P; c += 64) st_stream_f64x2(row + c, -1.0, -1.0, l2pol);
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

// ---- kernel D: partial followers leave their class inside the chunk ----------------------------------------------
// Lane per partial follower.  It shares the first d blocks of the chunk with representative R (same walk state before,
// same tokens): if R's walk ended inside those blocks, so does this one, with R's result.  Otherwise it resumes from
// R's snapshot after block d-1 and R's key of that block, and walks the rest of its chunk by itself -- hash one block,
// probe it, score it (the order of the fused kernel) -- which normally ends at the first block it does not share.
__device__ __forceinline__ uint64_t hash_block16(uint64_t parent, const uint32_t* __restrict__ tk) {
    Fnv f;
    f.begin_block(parent, 16);
    if ((reinterpret_cast<uintptr_t>(tk) & 15u) == 0) {
        const uint4* t4 = reinterpret_cast<const uint4*>(tk);
#pragma unroll
        for (int c = 0; c < 4; ++c) { const uint4 v = __ldg(t4 + c); f.token(v.x); f.token(v.y); f.token(v.z); f.token(v.w); }
    } else {
#pragma unroll
        for (int c = 0; c < 16; ++c) f.token(__ldg(tk + c));
    }
    return f.end_block();
}
#ifndef KVIDX_DETACH_BLOCKS
#define KVIDX_DETACH_BLOCKS 3
#endif
constexpr int kDetachBlocks = KVIDX_DETACH_BLOCKS;          // blocks a partial follower walks alone inside the round before it is re-queued
struct DetachWarp { double sc[kMaxEnt][32]; uint16_t pod[kMaxEnt][32]; };     // per warp (dynamic shared memory: warps of the CTA x this)
template <int BS>
__device__ __forceinline__ void detach_round(const TableView& t, const ScoreArgs& a, const RoundBufs& rb, const int cur, const int round, const int trace, DetachWarp* smw,
                                             const unsigned int bid, const unsigned int nbl, const uint64_t l2pol) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    DetachWarp& sm = smw[wid];
    const unsigned int n_dl = rb.n_hl[2];
    const PromptState* pst_rd = rb.pst[round & 1];
    PromptState* pst_cur = rb.pst[round & 1];
    const size_t kstride = (size_t)a.n_prompts;
    const bool peer = t.shard_bits != 0;
    const unsigned int total_warps = nbl * (blockDim.x / 32);
    for (unsigned int w = bid * (blockDim.x / 32) + wid; w * 32u < n_dl; w += total_warps) {
        const unsigned int f = w * 32u + lane;
        const bool have = f < n_dl;
        uint32_t p = 0, meta = 0;
        bool more = false;
        ScoreState s; s.k = 0; s.alive = 0;
        uint64_t h = 0;
        uint32_t pos0 = 0, adv = 0;
        if (have) {
            const uint4 rc = __ldg(rb.drec + f);
            p = rc.x; meta = rc.z & 0xffffu;
            const unsigned int lr = rc.y;
            const int d = (int)(rc.z >> 16);
            const uint32_t pl = rc.w;
            const unsigned int ar = rb.apos[lr];
            const int nb = (int)(meta & 63u);
            const PromptState& R = pst_rd[pl];
            s.k = R.k;
            for (uint32_t q = 0; q < s.k; ++q) s.pod[q] = R.pod[q];
            if ((int)rb.nwalk[ar] < d) {                       // R stopped inside the shared blocks: same stop, same scores
                for (uint32_t q = 0; q < s.k; ++q) s.sc[q] = R.sc[q];
            } else {
                const int j0 = rb.snap_run[(size_t)ar * kRoundBlocks + (size_t)(d - 1)];
                const size_t sj = (size_t)ar * kRoundBlocks + (size_t)j0;
                s.alive = rb.snap_alive[sj];
                for (uint32_t q = 0; q < s.k; ++q) {
                    double v = rb.snap_sc[sj * kMaxEnt + q];
                    if (j0 < d - 1 && ((s.alive >> q) & 1u)) {          // later blocks of the run: the same addend again, in order
                        const uint32_t bt = rb.snap_bt[sj * kMaxEnt + q];
                        const double add = bt == 0xffu ? 0.0 : t.weight[bt & 15u];
                        for (int u = j0; u < d - 1; ++u) v = __dadd_rn(v, add);
                    }
                    s.sc[q] = v;
                }
                h = rb.keys[(size_t)(d - 1) * kstride + ar];
                const uint32_t mdl = a.model ? a.model[p] : a.model0;
                pos0 = round > 0 ? rb.pos[p] : 0u;
                const uint32_t* tk = a.tok + (a.tok_off[p] - a.tok_base) + (int64_t)pos0 * BS;
                bool walking = true;
                const int bend = min(nb, d + kDetachBlocks);          // walk at most this far alone; the rest waits for the next round
                uint64_t hk = d < nb ? hash_block16(h, tk + (size_t)d * BS) : 0ull;       // key of the first block of its own
                int b = d;
                for (; b < bend; ++b) {
                    if (trace) atomicAdd(&rb.n_hl[3], 1u);
                    const uint64_t hm = home_of(hk, mdl);
                    const ReqSlot* base = t.req_peer[shard_of(hm, t.shard_bits)];
                    uint64_t slot = hm & t.req_mask & ~1ull;
                    const bool remote = base != t.req && !t.peer_pair;                  // a peer's shard: home slot first, its neighbour only if needed
                    uint4 A0, B0, A1, B1;
                    ld_slot(base + slot, peer, A0, B0);
                    if (!remote) ld_slot(base + slot + 1, peer, A1, B1);
                    // the next key does not depend on the probe: hash it while the probe is in flight
                    const uint64_t hn = b + 1 < nb ? hash_block16(hk, tk + (size_t)(b + 1) * BS) : 0ull;
                    SlotWords sw;
                    bool hit;
                    for (;;) {
                        sw.a = A0; sw.b = B0;
                        hit = slot_matches(sw.a, sw.b, hk, mdl);
                        bool stop = hit || meta_state(sw.b.w) == kStateEmpty;
                        if (!stop) {
                            if (remote) ld_slot(base + slot + 1, peer, A1, B1);
                            sw.a = A1; sw.b = B1; hit = slot_matches(sw.a, sw.b, hk, mdl); stop = hit || meta_state(sw.b.w) == kStateEmpty;
                        }
                        if (stop) break;
                        slot = (slot + 2) & t.req_mask;                 // rare: displaced past the home pair
                        ld_slot(base + slot, peer, A0, B0);
                        if (!remote) ld_slot(base + slot + 1, peer, A1, B1);
                    }
                    if (!hit) { walking = false; break; }
                    s.next(t, sw);
                    if (!s.alive) { walking = false; break; }
                    h = hk; hk = hn;
                }
                more = walking && (b < nb || ((meta >> 8) & 1u));
                adv = (uint32_t)b;
            }
        }
        // continue on its own next round, or write the result
        const uint32_t mm = __ballot_sync(0xffffffffu, more);
        if (mm) {
            unsigned int base = 0;
            if (lane == 0) base = atomicAdd(&rb.n_act[cur ^ 1], (unsigned int)__popc(mm));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (more) {
                { const unsigned int ns = base + __popc(mm & ((1u << lane) - 1u)); rb.act[cur ^ 1][ns] = p; rb.lslot[p] = ns; }
                rb.hstate[p] = h; rb.src[p] = p; rb.pos[p] = pos0 + adv;
                PromptState& ps = pst_cur[p];
                ps.k = (uint8_t)s.k; ps.alive = (uint16_t)s.alive;
                ps.pat[5] = 0xffffffffu;                        // no cached slot pattern: the next block is scored the long way
                for (uint32_t q = 0; q < s.k; ++q) { ps.sc[q] = s.sc[q]; ps.pod[q] = s.pod[q]; ps.bt[q] = 0xffu; }
            }
        }
        for (uint32_t q = 0; q < s.k; ++q) { sm.sc[q][lane] = s.sc[q]; sm.pod[q][lane] = s.pod[q]; }
        __syncwarp();
        uint32_t dm = __ballot_sync(0xffffffffu, have && !more);
        while (dm) {
            const int l = __ffs(dm) - 1; dm &= dm - 1;
            const uint32_t pp = __shfl_sync(0xffffffffu, p, l);
            const uint32_t pk = __shfl_sync(0xffffffffu, s.k, l);
            if (a.dense) {
                double* row = a.dense + (long long)pp * t.max_pods;
                const uint32_t P = t.max_pods;
                if ((P & 1u) == 0 && ((reinterpret_cast<uintptr_t>(row) & 15u) == 0)) {
                    for (uint32_t c = lane * 2; c < P; c += 64) st_stream_f64x2(row + c, -1.0, -1.0, l2pol);
                } else {
                    for (uint32_t c = lane; c < P; c += 32) row[c] = -1.0;
                }
                __syncwarp();
                if ((uint32_t)lane < pk) { const uint32_t pd = sm.pod[lane][l]; if (pd < P) row[pd] = sm.sc[lane][l]; }
            }
            if (a.sp_cnt) {
                if ((uint32_t)lane < pk) { a.sp_pods[(long long)pp * kMaxEnt + lane] = sm.pod[lane][l]; a.sp_scores[(long long)pp * kMaxEnt + lane] = sm.sc[lane][l]; }
                if (lane == 0) a.sp_cnt[pp] = (uint8_t)pk;
            }
            if (a.has_keys && lane == 0) a.has_keys[pp] = 1;
        }
        __syncwarp();
    }
}

// Kernels R and D in one launch (they work on different lists and neither reads what the other writes, except the next
// live list, which both append to).
template <int BS>
__global__ void __launch_bounds__(256)
finish_round_kernel(const TableView t, const ScoreArgs a, const RoundBufs rb, const int cur, const int round, const int detach, const int trace) {
    extern __shared__ __align__(16) unsigned char smem_raw_f[];
    DetachWarp* sm = reinterpret_cast<DetachWarp*>(smem_raw_f);
    const uint64_t l2pol = l2_policy_stream();
    if (!detach) { resolve_round(t, a, rb, cur, round, blockIdx.x, gridDim.x, l2pol); return; }
    // odd CTAs take the followers, even CTAs the partial followers: the two latency chains run side by side
    const unsigned int half = gridDim.x / 2;
    if (gridDim.x < 2) { resolve_round(t, a, rb, cur, round, 0, 1, l2pol); detach_round<BS>(t, a, rb, cur, round, trace, sm, 0, 1, l2pol); }
    else if (blockIdx.x & 1) { if (blockIdx.x / 2 < half) resolve_round(t, a, rb, cur, round, blockIdx.x / 2, half, l2pol); }
    else detach_round<BS>(t, a, rb, cur, round, trace, sm, blockIdx.x / 2, (gridDim.x + 1) / 2, l2pol);
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
            // (independent mixes of token pairs, not a chain: this kernel is one thread per prompt and latency bound)
            const uint32_t nt = block_size < 16u ? block_size : 16u;
            uint64_t g = 0;
            for (uint32_t j = 0; j + 1 < nt; j += 2) g ^= mix64(((uint64_t)__ldg(tk + j + 1) << 32 | __ldg(tk + j)) + 0xC2B2AE3D27D4EB4Full * (j + 1));
            if (nt & 1u) g ^= mix64(__ldg(tk + nt - 1) + 0x165667B19E3779F9ull);
            f = mix64(f ^ g);
            f &= ~(1ull << 63);
        }
        fp[i] = f; idx[i] = (uint32_t)i;
    }
    nb = __reduce_max_sync(0xffffffffu, (unsigned)min(nb, 0xffffffffull));
    if ((threadIdx.x & 31) == 0 && nb) atomicMax(max_blocks, nb);
    if (i == 0) for (int q = 0; q < kMaxParts; ++q) { cnt[8 * q] = ps.n[q]; cnt[8 * q + 1] = 0; }      // live-list lengths per part
}

// Distinct values among the sorted fingerprints (the sort orders by the top 32 bits, so those are compared).
__global__ void count_distinct_prefixes_kernel(const uint64_t* __restrict__ fp_sorted, int64_t n, unsigned long long* out) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const bool edge = i < n && (i == 0 || (fp_sorted[i] >> 32) != (fp_sorted[i - 1] >> 32));
    const unsigned int m = __ballot_sync(0xffffffffu, edge);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(out, (unsigned long long)__popc(m));
}

inline int rounds_init() {
    if (cudaFuncSetAttribute(hash_round_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(HashSmem<16>)) != cudaSuccess) return -1;
    if (cudaFuncSetAttribute(group_round_kernel<16, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(GroupSmem)) != cudaSuccess) return -1;
    if (cudaFuncSetAttribute(group_round_kernel<16, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(GroupSmem)) != cudaSuccess) return -1;
    return 0;
}

}  // namespace kvx
