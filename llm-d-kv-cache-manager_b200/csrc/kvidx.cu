// kvidx.cu -- C ABI (include/kvidx.h) over the sm_100a kernels.  No torch types, no CPU fallback.
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <shared_mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/kvidx.h"
#include "kernels_v1.cuh"
#include "kernels_write.cuh"
#include "kernels_score.cuh"
#include "kernels_rounds.cuh"
#include "kernels_rounds_plain.cuh"
#include "kernels_coop.cuh"
#include "kernels_route.cuh"
#include <cub/device/device_radix_sort.cuh>

using namespace kvx;

namespace {

thread_local std::string g_err;

int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
    g_err = buf;
    return code;
}

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) return fail(KVIDX_ECUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

uint64_t pow2ceil(uint64_t v) { uint64_t p = 1; while (p < v) p <<= 1; return p; }

// growable device / pinned-host scratch
struct DevBuf {
    void* p = nullptr; size_t cap = 0;
    cudaError_t need(size_t n) {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = std::max(n, (size_t)4096);
        want = want + want / 4;
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <class T> T* as() { return static_cast<T*>(p); }
};
struct PinBuf {
    void* p = nullptr; size_t cap = 0;
    cudaError_t need(size_t n) {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        size_t want = std::max(n, (size_t)4096);
        want = want + want / 4;
        cudaError_t e = cudaMallocHost(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
    template <class T> T* as() { return static_cast<T*>(p); }
};

bool is_device_accessible_host(const void* p) {
    cudaPointerAttributes a{};
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost;     // pinned: cudaMemcpyAsync from it is truly async
}

}  // namespace

// Concurrent host-buffer Score() callers (one goroutine / OS thread per RPC, server.go:70-96) are coalesced: whoever finds
// no batch in progress becomes the leader, takes everything that has queued up meanwhile, runs it as ONE batch through the
// pipeline and wakes the owners.  No dedicated thread, no artificial delay: a lone caller goes straight through with its
// own buffers; under load the batch size grows by itself to what arrived during the previous launch.
struct SubmitReq {
    const uint32_t* tok; const int64_t* tok_off; int64_t n; const uint32_t* model; uint32_t model0; const uint64_t* filter;
    double* dense; uint16_t* sp_pods; double* sp_scores; uint8_t* sp_cnt; uint8_t* has_keys;
    int rc = 0; bool done = false, promote = false; std::string err;
    std::mutex m; std::condition_variable cv;      // each waiter sleeps on its OWN condition variable: finishing a batch wakes
                                                   // exactly its owners (and the next leader), not every queued thread
};
struct SubmitQueue {
    std::mutex mu;
    std::deque<SubmitReq*> pending;
    bool leader_active = false;
    bool enabled = true;
    int64_t max_prompts = 65536, max_tokens = 64ll << 20;
    std::atomic<uint64_t> coalesced{0}, batches{0};
    // combined host-side batch (pinned), owned by the current leader
    void* h_in = nullptr; size_t h_in_cap = 0;
    void* h_out = nullptr; size_t h_out_cap = 0;
};

struct kvidx {
    kvidx_config_t cfg{};
    TableView tv{};
    int device = 0;
    int sm_count = 148;
    // Read path (Score / Lookup / hash) and write path (Add / Evict / events) run on their own streams and behind their own
    // mutexes: a Score() call and an event batch overlap on the device (table.cuh: readers never wait for writers).
    cudaStream_t own_stream = nullptr, stream = nullptr, copy_stream = nullptr, d2h_stream = nullptr, own_wstream = nullptr;
    cudaEvent_t ev_h2d[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr}, ev_k[2] = {nullptr, nullptr};
    cudaEvent_t ev_write = nullptr;   // last asynchronous (device-resident) write batch: later reads are ordered after it
    Counters* d_cnt = nullptr;
    Counters* h_cnt = nullptr;     // pinned mirror (write side)
    Counters* d_cnt_all = nullptr; // sharded: every owner's counters, gathered before a write batch
    Counters* h_cnt_all = nullptr;
    uint64_t rebuilds = 0;
    std::atomic<uint64_t> launches{0};
    int64_t last_batch_events = 1 << 20;   // events in the batch being applied (sizes the stamp range)
    std::atomic<unsigned long long> clock{1};   // exact-LRU mode: recency stamps handed to kernels
    std::mutex mu_r, mu_w;          // read-side / write-side scratch and stream
    std::shared_mutex tables;       // shared: any call that launches on the tables; exclusive: rebuild swaps them
    // read-side scratch (guarded by mu_r)
    DevBuf d_tok[2], d_off[2], d_model[2], d_filter[2], d_out[2], d_aux[2], d_rmisc;
    PinBuf h_stage[2], h_out[2];
    // write-side scratch (guarded by mu_w)
    DevBuf d_misc, d_ev, d_hash, d_evtok, d_qoff, d_wkeys, d_wpred, d_wready, d_wmap;
    PinBuf h_misc;
    int write_phase1 = -1;          // -1: by batch size, 0: never, 1: always (phase 1 = hash_events_kernel)
    int64_t write_phase1_min = 1024; // engine hashes in a batch from which phase 1 pays
    int score_kernel = 2;          // 1 = v1 (thread per prompt, global tokens), 2 = tuned
    int score_path = 0;            // 0 = by batch size, 1 = always the fused persistent kernel, 2 / 3 = always the plain / class round pipeline,
                                   // 4 = always the warp-per-prompt cooperative kernel
    int64_t coop_max = 2048;       // batches up to this many prompts use the warp-per-prompt cooperative kernel
    int rounds_spec = 0;           // per-prompt rounds: run hash(t+1) beside walk(t) for small batches (measured at 65 536 prompts: 3 % slower, so off)
    int64_t rounds_spec_max = 131072;
    cudaEvent_t ev_spec_h[2][2] = {}, ev_spec_p[2][2] = {};
    int64_t zerocopy_max = 32;     // host-buffer calls up to this many prompts skip the copy engine (tokens read from pinned host memory)
    int group_tma = 0;             // class pipeline: token chunks by TMA bulk copy instead of cp.async (measured: 2 % slower per step)
    struct SubmitQueue* queue = nullptr;   // coalesces concurrent host-buffer Score() callers (submit.cuh)
    int64_t rounds_min = 4096;     // batches at least this large use the round pipeline (measured: warp-per-prompt rounds beat the fused
                                   // kernel from 4096 prompts, profiles/r2_medium_batch_ab.txt)
    int64_t classes_min = 393216;  // ... and at least this large, the prefix-class round pipeline
    double classes_min_sharing = 0.85;   // ... if at least this fraction of the batch follows a representative in round 0
    DevBuf r_act0, r_act1, r_cnt, r_hstate, r_keys, r_pst, r_nbr, r_fp, r_sort, r_role, r_hl, r_map, r_src, r_fate, r_anch, r_snap, r_rec;
    int64_t host_chunk_tokens = 32ll << 20;   // tokens per H2D chunk of the host-buffer pipeline
    int rounds_trace = 0;
    int rounds_grid[5] = {2, 4, 2, 4, 4};   // CTAs per SM a part's group / lists / hash / walk / finish kernel may occupy (multi-part runs)
    int rounds_parts = 8;          // parts (streams) a large batch is split into
    int rounds_dedup = 2;          // round pipeline: 0 every prompt on its own, 1 prefix classes, 2 + partial followers
    int sort_prefix = 1;           // sort the batch by first-block fingerprint before the rounds
    int rounds_overlap = 1;        // run the two halves of a large batch on two streams
    int64_t rounds_overlap_min = 65536;
    cudaStream_t aux_stream[kMaxParts - 1] = {}; cudaEvent_t ev_fork = nullptr, ev_join[kMaxParts - 1] = {};
    // group_serial: every part's kernel G runs on ONE stream, one after the other at full bandwidth, while the other parts'
    // latency-bound kernels (G2, H, P, R) run beside it on their own streams (otherwise the parts drift into lock step: all in
    // G sharing the bandwidth, then all in the short kernels with DRAM idle -- scripts/timeline.py)
    int group_serial = 2, group_serial_grid = 2;
    int rounds_lane_stages = 2; int64_t rounds_lane_stages_max = 1 << 30;      // lane-per-prompt rounds: stages of kernel H (experiment knob)
    int rounds_warp = 1, rounds_warp_stages = 3; int64_t rounds_warp_max = 57344;       // per-prompt rounds: kernel P warp per prompt up to this many prompts
                                   // (measured crossover with the lane-per-prompt kernel P between 49152 and 65536)
    int group_ctas = 0;            // > 0: CTAs of one part's kernel G (all parts' G CTAs resident at once: nothing queues behind them)
    int hash_prefetch = 0;         // kernel H pulls its chunks towards L2 before the chain starts (1; 2: evict_last)
    int small_cta = 256;           // threads per CTA of the short kernels (G2, H, P, R): small CTAs fit into what kernel G's CTAs leave of an SM
    cudaStream_t g_stream = nullptr, hp_stream[kMaxParts] = {}, lp_stream[kMaxParts] = {}; cudaEvent_t ev_g[kMaxParts] = {}, ev_r[kMaxParts] = {};
};

namespace {

// a sharded handle is usable once every peer shard has been mapped
int check_shards(kvidx* x) {
    for (uint32_t r = 0; r < (1u << x->tv.shard_bits); ++r)
        if (!x->tv.req_peer[r]) return fail(KVIDX_EINVAL, "shard %u of %u is not connected (kvidx_shard_import / kvidx_shard_attach)", r, 1u << x->tv.shard_bits);
    return 0;
}

unsigned long long reserve_stamps(kvidx* x, unsigned long long n) { return x->clock.fetch_add(n + 1); }

// The stream write kernels go to.  Exact-LRU mode keeps reads and writes on ONE stream (reads write recency stamps, so the
// two sides are not independent there); otherwise the write path has its own.
cudaStream_t wstream(kvidx* x) { return x->tv.req_stamp ? x->stream : x->own_wstream; }

// Locks of a call.  Read-side calls take mu_r (+ mu_w in exact-LRU mode, where everything is serialised as in round 1),
// write-side calls mu_w; both hold `tables` shared while they launch.  Only rebuild() takes `tables` exclusively.
struct ReadGuard {
    std::unique_lock<std::mutex> w, r;
    std::shared_lock<std::shared_mutex> t;
    explicit ReadGuard(kvidx* x) {
        if (x->tv.req_stamp) w = std::unique_lock<std::mutex>(x->mu_w);
        r = std::unique_lock<std::mutex>(x->mu_r);
        t = std::shared_lock<std::shared_mutex>(x->tables);
    }
};

int refresh_counters(kvidx* x) {
    CK(cudaMemcpyAsync(x->h_cnt, x->d_cnt, sizeof(Counters), cudaMemcpyDeviceToHost, wstream(x)));
    CK(cudaStreamSynchronize(wstream(x)));
    return 0;
}

int alloc_tables(kvidx* x, uint64_t req_slots, uint64_t eng_slots, ReqSlot** req, EngSlot** eng, cudaStream_t st) {
    *req = nullptr; *eng = nullptr;
    CK(cudaMalloc((void**)req, req_slots * sizeof(ReqSlot)));
    cudaError_t e = cudaMalloc((void**)eng, eng_slots * sizeof(EngSlot));
    if (e != cudaSuccess) { cudaFree(*req); *req = nullptr; return fail(KVIDX_ENOMEM, "cudaMalloc of %llu engine slots: %s", (unsigned long long)eng_slots, cudaGetErrorString(e)); }
    CK(cudaMemsetAsync(*req, 0, req_slots * sizeof(ReqSlot), st));
    CK(cudaMemsetAsync(*eng, 0, eng_slots * sizeof(EngSlot), st));
    return 0;
}

// Drop tombstones: re-insert the live slots into fresh tables.  Caller holds mu_w and `tables` EXCLUSIVELY; every stream of
// the device is drained first (asynchronous Score() calls may still be reading the old tables).
//   unsharded: the fresh tables replace the old ones;
//   sharded  : peers have this shard mapped (CUDA IPC), so the fresh tables are copied back into the same allocation.  The
//              caller of kvidx_shard_compact guarantees that no rank touches the index meanwhile (a barrier either side).
int rebuild(kvidx* x, bool in_place) {
    CK(cudaDeviceSynchronize());
    cudaStream_t st = wstream(x);
    const uint64_t rs = x->tv.req_mask + 1, es = x->tv.eng_mask + 1;
    ReqSlot* nreq; EngSlot* neng;
    int rc = alloc_tables(x, rs, es, &nreq, &neng, st);
    if (rc) return rc;
    const int T = 256;
    unsigned long long* nstamp = nullptr;
    if (x->tv.req_stamp) { CK(cudaMalloc((void**)&nstamp, rs * 8)); CK(cudaMemsetAsync(nstamp, 0, rs * 8, st)); }
    rebuild_req_kernel<<<(unsigned)((rs + T - 1) / T), T, 0, st>>>(x->tv.req, rs, nreq, rs - 1, x->tv.req_stamp, nstamp);
    rebuild_eng_kernel<<<(unsigned)((es + T - 1) / T), T, 0, st>>>(x->tv.eng, es, neng, es - 1);
    x->launches += 2;
    CK(cudaGetLastError());
    // zero the tombstone counters on the device
    CK(cudaMemsetAsync(&x->d_cnt->req_tomb, 0, sizeof(unsigned long long), st));
    CK(cudaMemsetAsync(&x->d_cnt->eng_tomb, 0, sizeof(unsigned long long), st));
    if (in_place) {
        CK(cudaMemcpyAsync(x->tv.req, nreq, rs * sizeof(ReqSlot), cudaMemcpyDeviceToDevice, st));
        CK(cudaMemcpyAsync(x->tv.eng, neng, es * sizeof(EngSlot), cudaMemcpyDeviceToDevice, st));
        CK(cudaStreamSynchronize(st));
        cudaFree(nreq); cudaFree(neng);
    } else {
        CK(cudaStreamSynchronize(st));
        cudaFree(x->tv.req); cudaFree(x->tv.eng);
        if (x->tv.req_stamp) { cudaFree(x->tv.req_stamp); x->tv.req_stamp = nstamp; }
        x->tv.req = nreq; x->tv.eng = neng;
        x->tv.req_peer[0] = nreq; x->tv.eng_peer[0] = neng;
    }
    ++x->rebuilds;
    return refresh_counters(x);
}

// Make room for up to `incoming` new keys in each table before a write batch.  Caller holds mu_w and NOT `tables`.
//   unsharded: this handle's counters decide; tombstones are compacted away here when they are what fills the table;
//   sharded  : the keys may land on any owner, so every owner's counters are read (through the mapped peer memory) and
//              each must have room for the whole batch; compaction is the collective kvidx_shard_compact.
int ensure_room(kvidx* x, uint64_t incoming) {
    const uint64_t rs = x->tv.req_mask + 1, es = x->tv.eng_mask + 1;
    auto over = [&](uint64_t full, uint64_t tomb, uint64_t slots) { return (full + tomb + incoming) * 10 > slots * 8; };
    if (x->tv.shard_bits) {
        const uint32_t ns = 1u << x->tv.shard_bits;
        gather_counters_kernel<<<1, 32, 0, wstream(x)>>>(x->tv, x->d_cnt_all);
        x->launches += 1;
        CK(cudaMemcpyAsync(x->h_cnt_all, x->d_cnt_all, sizeof(Counters) * ns, cudaMemcpyDeviceToHost, wstream(x)));
        CK(cudaStreamSynchronize(wstream(x)));
        for (uint32_t r = 0; r < ns; ++r) {
            const Counters& c = x->h_cnt_all[r];
            if ((c.req_full + c.req_tomb + incoming) * 10 > rs * 9 || (c.eng_full + c.eng_tomb + incoming) * 10 > es * 9)
                return fail(KVIDX_ENOSPC, "shard %u full: %llu keys + %llu tombstones + %llu incoming in %llu slots%s", r,
                            (unsigned long long)c.req_full, (unsigned long long)c.req_tomb, (unsigned long long)incoming, (unsigned long long)rs,
                            (c.req_tomb || c.eng_tomb) ? " (kvidx_shard_compact on every rank drops the tombstones)" : "");
        }
        return 0;
    }
    const Counters& c = *x->h_cnt;
    if (over(c.req_full, c.req_tomb, rs) || over(c.eng_full, c.eng_tomb, es)) {
        if (c.req_tomb || c.eng_tomb) {
            std::unique_lock<std::shared_mutex> tl(x->tables);
            int rc = rebuild(x, false);
            if (rc) return rc;
        }
        const Counters& d = *x->h_cnt;
        if ((d.req_full + incoming) * 10 > rs * 9 || (d.eng_full + incoming) * 10 > es * 9)
            return fail(KVIDX_ENOSPC, "table full: %llu request keys + %llu incoming in %llu slots",
                        (unsigned long long)d.req_full, (unsigned long long)incoming, (unsigned long long)rs);
    }
    return 0;
}

// exact-LRU mode: after a write call, evict least-recently-used keys until both maps respect InMemoryIndexConfig.Size
// (lru.Cache evicts the back of its list when Len() > size; in_memory.go:59,64).  Exact at call granularity.
int enforce_caps(kvidx* x) {
    if (!x->tv.req_stamp) return 0;
    int rc = refresh_counters(x);
    if (rc) return rc;
    CK(x->d_misc.need(64));
    unsigned long long* d_v = x->d_misc.as<unsigned long long>();
    const uint64_t rs = x->tv.req_mask + 1, es = x->tv.eng_mask + 1;
    const int T = 256;
    cudaStream_t st = wstream(x);
    while (x->h_cnt->req_full > x->tv.capacity) {
        CK(cudaMemsetAsync(d_v, 0xff, 8, st));
        lru_min_req_kernel<<<(unsigned)((rs + T - 1) / T), T, 0, st>>>(x->tv.req, x->tv.req_stamp, rs, d_v);
        lru_drop_req_kernel<<<(unsigned)((rs + T - 1) / T), T, 0, st>>>(x->tv.req, x->tv.req_stamp, rs, d_v, x->d_cnt);
        x->launches += 2;
        CK(cudaGetLastError());
        if ((rc = refresh_counters(x))) return rc;
    }
    while (x->h_cnt->eng_full > x->tv.capacity) {
        CK(cudaMemsetAsync(d_v, 0xff, 8, st));
        lru_min_eng_kernel<<<(unsigned)((es + T - 1) / T), T, 0, st>>>(x->tv.eng, es, d_v);
        lru_drop_eng_kernel<<<(unsigned)((es + T - 1) / T), T, 0, st>>>(x->tv.eng, es, d_v, x->d_cnt);
        x->launches += 2;
        CK(cudaGetLastError());
        if ((rc = refresh_counters(x))) return rc;
    }
    return 0;
}

struct ScoreOut { double* dense; uint16_t* sp_pods; double* sp_scores; uint8_t* sp_cnt; uint8_t* has_keys; };

int launch_score_rounds_plain(kvidx* x, const uint32_t* d_tok, const int64_t* d_off, int64_t tok_base, int64_t n,
                              const uint32_t* d_model, uint32_t model0, const uint64_t* d_filter, const ScoreOut& o, cudaStream_t st,
                              int64_t max_blocks);

// Large batches: rounds of group / hash / walk / resolve kernels (kernels_rounds.cuh).  max_blocks < 0: computed on the
// device.  The sorted batch is split into parts that run their rounds on separate streams, so that one part's
// latency-bound kernels (the serial FNV chain of few representatives, the dependent probes) share the SMs with
// another part's bandwidth-bound ones.
int launch_score_rounds(kvidx* x, const uint32_t* d_tok, const int64_t* d_off, int64_t tok_base, int64_t n,
                        const uint32_t* d_model, uint32_t model0, const uint64_t* d_filter, const ScoreOut& o, cudaStream_t st,
                        int64_t max_blocks) {
    CK(x->r_act0.need((size_t)n * 4)); CK(x->r_act1.need((size_t)n * 4)); CK(x->r_cnt.need(1024));
    CK(x->r_hstate.need((size_t)n * 8)); CK(x->r_keys.need((size_t)n * kRoundBlocks * 8)); CK(x->r_pst.need((size_t)n * sizeof(PromptState) * 2));
    CK(x->r_nbr.need((size_t)n * 4 * 2)); CK(x->r_role.need((size_t)n * 4)); CK(x->r_hl.need((size_t)n * 4 * 2)); CK(x->r_src.need((size_t)n * 4));
    const int64_t n_al = (n + 63) & ~63ll;
    CK(x->r_fate.need((size_t)n_al * 4 + 64 * kMaxParts));      // fate, dmin, then per part nfol | need_snap
    CK(x->r_anch.need((size_t)n * 4 * 4));  // anch, apos, dl, lslot
    CK(x->r_rec.need((size_t)n * 16 * 2));  // rec, drec
    CK(x->r_snap.need((size_t)n * (kRoundBlocks * kMaxEnt * 8 + kRoundBlocks * 2 + 1 + kRoundBlocks + kRoundBlocks * kMaxEnt)));
    int np = 1;
    if (x->rounds_overlap && n >= x->rounds_overlap_min) np = std::max(1, std::min(kMaxParts, x->rounds_parts));
    int64_t psz[kMaxParts] = {}, poff[kMaxParts] = {};
    {
        const int64_t per = ((n + np - 1) / np + 31) & ~31ll;
        int64_t at = 0;
        for (int q = 0; q < np; ++q) { poff[q] = at; psz[q] = std::max<int64_t>(0, std::min(per, n - at)); at += psz[q]; }
    }
    const uint64_t map_slots = pow2ceil((uint64_t)std::max<int64_t>(4 * ((n + np - 1) / np), 1024));
    const size_t gstride = map_slots + (size_t)((((n + np - 1) / np) + 32 + 3) & ~3ll);      // election map + grp per part, cleared together
    CK(x->r_map.need(gstride * 4 * np));
    unsigned int* cnt = x->r_cnt.as<unsigned int>();
    unsigned long long* d_maxb = reinterpret_cast<unsigned long long*>(cnt + 8 * kMaxParts);
    RoundBufs rb[kMaxParts]{};
    for (int q = 0; q < np; ++q) {
        const int64_t off = poff[q];
        rb[q].act[0] = x->r_act0.as<uint32_t>() + off; rb[q].act[1] = x->r_act1.as<uint32_t>() + off;
        rb[q].n_act = cnt + 8 * q;
        rb[q].n_hl = cnt + 8 * q + 2;
        rb[q].hstate = x->r_hstate.as<uint64_t>(); rb[q].src = x->r_src.as<uint32_t>();
        rb[q].pst[0] = x->r_pst.as<PromptState>(); rb[q].pst[1] = x->r_pst.as<PromptState>() + n;
        rb[q].keys = x->r_keys.as<uint64_t>() + off; rb[q].nbr = x->r_nbr.as<uint32_t>() + off; rb[q].pos = x->r_nbr.as<uint32_t>() + n;
        rb[q].role = x->r_role.as<uint32_t>() + off; rb[q].fate = x->r_fate.as<uint8_t>() + off;
        rb[q].part_size = (uint32_t)((psz[q] + 31) & ~31ll);
        rb[q].nfol = x->r_fate.as<uint8_t>() + 2 * n_al + 2 * off + 64 * q; rb[q].need_snap = rb[q].nfol + rb[q].part_size;
        rb[q].dmin = x->r_fate.as<uint8_t>() + n_al + off;
        rb[q].anch = x->r_anch.as<uint32_t>() + off; rb[q].apos = x->r_anch.as<uint32_t>() + n + off; rb[q].dl = x->r_anch.as<uint32_t>() + 2 * n + off; rb[q].lslot = x->r_anch.as<uint32_t>() + 3 * n;
        rb[q].rec = x->r_rec.as<uint4>() + off; rb[q].drec = x->r_rec.as<uint4>() + n + off;
        rb[q].snap_sc = x->r_snap.as<double>() + (size_t)off * kRoundBlocks * kMaxEnt;
        rb[q].snap_alive = reinterpret_cast<uint16_t*>(x->r_snap.as<double>() + (size_t)n * kRoundBlocks * kMaxEnt) + (size_t)off * kRoundBlocks;
        uint8_t* snap_bytes = reinterpret_cast<uint8_t*>(reinterpret_cast<uint16_t*>(x->r_snap.as<double>() + (size_t)n * kRoundBlocks * kMaxEnt) + (size_t)n * kRoundBlocks);
        rb[q].nwalk = snap_bytes + off;
        rb[q].snap_run = snap_bytes + n + (size_t)off * kRoundBlocks;
        rb[q].snap_bt = snap_bytes + n + (size_t)n * kRoundBlocks + (size_t)off * kRoundBlocks * kMaxEnt;
        rb[q].hl = x->r_hl.as<uint32_t>() + off; rb[q].fl = x->r_hl.as<uint32_t>() + n + off;
        rb[q].map = x->r_map.as<uint32_t>() + (size_t)q * gstride; rb[q].map_mask = (uint32_t)(map_slots - 1);
        rb[q].grp = rb[q].map + map_slots;
    }
    ScoreArgs a{d_tok, d_off, tok_base, n, d_model, model0, d_filter, o.dense, o.sp_pods, o.sp_scores, o.sp_cnt, o.has_keys, nullptr};
    CK(cudaMemsetAsync(x->r_cnt.p, 0, 1024, st));
    CK(x->r_fp.need((size_t)n * 8 * 2 + (size_t)n * 4));
    uint64_t* fp_in = x->r_fp.as<uint64_t>(); uint64_t* fp_out = fp_in + n;
    uint32_t* idx_in = reinterpret_cast<uint32_t*>(fp_out + n);
    PartSizes ps{};
    for (int q = 0; q < kMaxParts; ++q) ps.n[q] = (unsigned int)psz[q];
    rounds_init_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(a, x->tv.block_size, d_maxb, fp_in, x->sort_prefix ? idx_in : rb[0].act[0], cnt, ps);
    x->launches += 1;
    CK(cudaGetLastError());
    if (x->sort_prefix) {
        size_t tmp = 0;
        CK(cub::DeviceRadixSort::SortPairs(nullptr, tmp, fp_in, fp_out, idx_in, rb[0].act[0], (int)n, 32, 64, st));
        CK(x->r_sort.need(tmp));
        CK(cub::DeviceRadixSort::SortPairs(x->r_sort.p, tmp, fp_in, fp_out, idx_in, rb[0].act[0], (int)n, 32, 64, st));
        x->launches += 5;     // cub: histogram + 4 onesweep passes (the top 32 fingerprint bits order the groups well enough)
    }
    // Is the batch worth it?  The class pipeline pays a token pass and five kernels per round; with few prompts per distinct
    // prefix (measured crossover: ~6 prompts per cached document) the per-prompt rounds are faster.  The sorted first-block
    // fingerprints tell before anything is spent: count the distinct ones, and hand the batch over if more than 15 % of the
    // prompts start differently from their neighbour.  (The device-resident entry point waits for the block count here anyway.)
    const bool adaptive = x->score_path == 0 && x->rounds_dedup && x->sort_prefix && x->classes_min_sharing > 0.0;
    if (adaptive) {
        count_distinct_prefixes_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(fp_out, n, d_maxb + 1);
        x->launches += 1;
    }
    if (max_blocks < 0 || adaptive) {
        unsigned long long mb[2] = {0, 0};
        CK(cudaMemcpyAsync(mb, d_maxb, sizeof mb, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        if (max_blocks < 0) max_blocks = (int64_t)mb[0];
        if (adaptive && (double)mb[1] > (1.0 - x->classes_min_sharing) * (double)n) {
            if (x->rounds_trace) fprintf(stderr, "[kvidx rounds] %llu distinct first blocks among %lld prompts: per-prompt rounds instead\n", mb[1], (long long)n);
            return launch_score_rounds_plain(x, d_tok, d_off, tok_base, n, d_model, model0, d_filter, o, st, max_blocks);
        }
    }
    int64_t rounds = (max_blocks + kRoundBlocks - 1) / kRoundBlocks;
    if (x->rounds_dedup >= 2) rounds += 1;            // a prompt that left its class mid-chunk runs unaligned from there: one more round
    if (rounds < 1) rounds = 1;                       // round 0 also retires the prompts that have no full block
    cudaStream_t strm[kMaxParts];
    strm[0] = st;
    for (int q = 1; q < kMaxParts; ++q) strm[q] = x->aux_stream[q - 1];
    CK(cudaMemsetAsync(x->r_map.p, 0xff, gstride * 4 * np, st));
    CK(cudaMemsetAsync(x->r_fate.as<uint8_t>() + 2 * n_al, 0, (size_t)n_al * 2 + 64 * kMaxParts, st));     // nfol, need_snap
    if (np > 1) {
        CK(cudaEventRecord(x->ev_fork, st));
        for (int q = 1; q < np; ++q) CK(cudaStreamWaitEvent(strm[q], x->ev_fork, 0));
    }
    const auto t_enq = std::chrono::steady_clock::now();
    // Where kernel G runs: 0 the part's stream; 1 ONE stream for every part's G; 2 a low-priority stream per part, with the part's
    // short kernels on a high-priority one (a CTA slot that frees up goes to a waiting latency-bound kernel before the next G CTA).
    const int gmode = np > 1 ? x->group_serial : 0;
    cudaStream_t sS[kMaxParts], sG[kMaxParts];
    for (int q = 0; q < kMaxParts; ++q) { sS[q] = gmode == 2 ? x->hp_stream[q] : strm[q]; sG[q] = gmode == 2 ? x->lp_stream[q] : gmode == 1 ? x->g_stream : strm[q]; }
    if (gmode == 1) CK(cudaStreamWaitEvent(x->g_stream, x->ev_fork, 0));
    if (gmode == 2) for (int q = 0; q < np; ++q) { CK(cudaStreamWaitEvent(sS[q], x->ev_fork, 0)); CK(cudaStreamWaitEvent(sG[q], x->ev_fork, 0)); }
    const int T = x->small_cta, W = T / 32, kf = 256 / T;        // threads / warps of the short kernels' CTAs; CTAs per 256 threads
    for (int64_t r = 0; r < rounds; ++r) {
        const int cur = (int)(r & 1);
        for (int q = 0; q < np; ++q) {
            const int64_t m = psz[q];
            if (m <= 0) continue;
            unsigned ggrid = (unsigned)std::min<int64_t>((m + kGroupThreads / 32 * kGroupTile - 1) / (kGroupThreads / 32 * kGroupTile), (int64_t)x->sm_count * (np == 1 ? 3 : gmode == 1 ? x->group_serial_grid : x->rounds_grid[0]));
            if (x->group_ctas > 0 && np > 1) ggrid = std::min<unsigned>(ggrid, (unsigned)x->group_ctas);
            if (gmode && r > 0) CK(cudaStreamWaitEvent(sG[q], x->ev_r[q], 0));       // this part's previous round has built the live list
            if (x->group_tma) group_round_kernel<16, true><<<ggrid, kGroupThreads, sizeof(GroupSmem), sG[q]>>>(x->tv, a, rb[q], cur, (int)r, x->rounds_dedup);
            else group_round_kernel<16, false><<<ggrid, kGroupThreads, sizeof(GroupSmem), sG[q]>>>(x->tv, a, rb[q], cur, (int)r, x->rounds_dedup);
            if (gmode) { CK(cudaEventRecord(x->ev_g[q], sG[q])); CK(cudaStreamWaitEvent(sS[q], x->ev_g[q], 0)); }
            const unsigned lgrid = (unsigned)std::min<int64_t>((m + T - 1) / T, (int64_t)x->sm_count * kf * (np <= 2 ? 8 : x->rounds_grid[1]));
            group_lists_kernel<16><<<lgrid, T, 0, sS[q]>>>(x->tv, a, rb[q], cur, (int)r, x->rounds_dedup >= 2);
        }
        for (int q = 0; q < np; ++q) {
            const int64_t m = psz[q];
            if (m <= 0) continue;
            const int TH = std::min(T, kHashThreads);
            const unsigned hgrid = (unsigned)std::min<int64_t>((m + TH - 1) / TH, (int64_t)x->sm_count * (256 / TH) * (np == 1 ? 4 : x->rounds_grid[2]));
            hash_round_kernel<16><<<hgrid, TH, sizeof(HashSmem<16>) / (kHashThreads / 32) * (TH / 32), sS[q]>>>(x->tv, a, rb[q], cur, (int)r, x->hash_prefetch);
            const unsigned wgrid = (unsigned)std::min<int64_t>((m + W - 1) / W, (int64_t)x->sm_count * kf * (np == 1 ? 8 : x->rounds_grid[3]));
            walk_round_kernel<<<wgrid, T, 0, sS[q]>>>(x->tv, a, rb[q], cur, (int)r);
            const unsigned rgrid = (unsigned)std::min<int64_t>((m + T - 1) / T, (int64_t)x->sm_count * kf * (np <= 2 ? 8 : x->rounds_grid[4]));
            finish_round_kernel<16><<<rgrid, T, sizeof(DetachWarp) * W, sS[q]>>>(x->tv, a, rb[q], cur, (int)r, x->rounds_dedup >= 2, x->rounds_trace);
            if (gmode && r + 1 < rounds) CK(cudaEventRecord(x->ev_r[q], sS[q]));
            x->launches += 5;
            if (x->rounds_trace == 1) {      // debugging aid: list sizes of this round (synchronises)
                unsigned int c[8];
                CK(cudaMemcpyAsync(c, rb[q].n_act, sizeof c, cudaMemcpyDeviceToHost, sS[q]));
                CK(cudaStreamSynchronize(sS[q]));
                fprintf(stderr, "[kvidx rounds] round %lld part %d: live %u -> representatives %u, followers %u, partial %u (%u blocks walked alone) -> next %u\n",
                        (long long)r, q, c[cur], c[2], c[3], c[4], c[5], c[cur ^ 1]);
            }
        }
    }
    if (x->rounds_trace == 2) fprintf(stderr, "[kvidx rounds] host enqueue of %lld rounds x %d parts: %.3f ms\n", (long long)rounds, np,
                                      std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_enq).count());
    CK(cudaGetLastError());
    for (int q = 1; q < np; ++q) { CK(cudaEventRecord(x->ev_join[q - 1], sS[q])); CK(cudaStreamWaitEvent(st, x->ev_join[q - 1], 0)); }
    if (gmode == 2) { CK(cudaEventRecord(x->ev_g[0], sS[0])); CK(cudaStreamWaitEvent(st, x->ev_g[0], 0)); }
    return 0;
}

// Medium batches: hash / walk rounds, every prompt on its own (kernels_rounds_plain.cuh); two halves on two streams.
// Below rounds_spec_max prompts the rounds are latency bound (a few warps per SM walking dependent chains), so kernel H of
// round t+1 is launched BESIDE kernel P of round t on a second stream per half: it hashes the next 32 blocks of every prompt
// of round t's list without waiting to learn which of them survive (at most one round of hashing per prompt is wasted).
int launch_score_rounds_plain(kvidx* x, const uint32_t* d_tok, const int64_t* d_off, int64_t tok_base, int64_t n,
                        const uint32_t* d_model, uint32_t model0, const uint64_t* d_filter, const ScoreOut& o, cudaStream_t st,
                        int64_t max_blocks) {
    const bool spec = x->rounds_spec && n <= x->rounds_spec_max;
    const bool warp_walk = !spec && x->rounds_warp && n <= x->rounds_warp_max;      // kernel P: warp per prompt (batch too small to fill the machine lane per prompt)
    CK(x->r_act0.need((size_t)n * 4)); CK(x->r_act1.need((size_t)n * 4)); CK(x->r_cnt.need(64));
    CK(x->r_hstate.need((size_t)n * 8)); CK(x->r_keys.need((size_t)n * plain::kRoundBlocks * 8 * (spec ? 2 : 1))); CK(x->r_pst.need((size_t)n * sizeof(plain::PromptState)));
    CK(x->r_nbr.need((size_t)n * 4 * (spec ? 2 : 1)));
    if (spec) CK(x->r_hl.need((size_t)n * 4 * 2));
    const bool overlap = x->rounds_overlap && n >= x->rounds_overlap_min;
    const int64_t nA = overlap ? ((n / 2 + 31) & ~31ll) : n, nB = n - nA;
    unsigned int* cnt = x->r_cnt.as<unsigned int>();
    unsigned long long* d_maxb = reinterpret_cast<unsigned long long*>(x->r_cnt.as<unsigned char>() + 16);
    plain::RoundBufs rb[2]{};
    for (int hlf = 0; hlf < 2; ++hlf) {
        const int64_t off = hlf ? nA : 0;
        rb[hlf].act[0] = x->r_act0.as<uint32_t>() + off; rb[hlf].act[1] = x->r_act1.as<uint32_t>() + off;
        rb[hlf].n_act = cnt + 2 * hlf;
        rb[hlf].hstate = x->r_hstate.as<uint64_t>(); rb[hlf].pst = x->r_pst.as<plain::PromptState>();
        rb[hlf].keys = x->r_keys.as<uint64_t>() + (warp_walk ? off * plain::kRoundBlocks : off); rb[hlf].nbr = x->r_nbr.as<uint32_t>() + off;      // prompt-major keys: a half's rows are contiguous
        rb[hlf].spec = spec ? 1 : 0;
        if (spec) { rb[hlf].prev[0] = x->r_hl.as<uint32_t>() + off; rb[hlf].prev[1] = x->r_hl.as<uint32_t>() + n + off; }
    }
    ScoreArgs a{d_tok, d_off, tok_base, n, d_model, model0, d_filter, o.dense, o.sp_pods, o.sp_scores, o.sp_cnt, o.has_keys, nullptr};
    CK(cudaMemsetAsync(x->r_cnt.p, 0, 64, st));
    CK(x->r_fp.need((size_t)n * 8 * 2 + (size_t)n * 4));
    uint64_t* fp_in = x->r_fp.as<uint64_t>(); uint64_t* fp_out = fp_in + n;
    uint32_t* idx_in = reinterpret_cast<uint32_t*>(fp_out + n);
    plain::rounds_init_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(a, x->tv.block_size, d_maxb, fp_in, x->sort_prefix ? idx_in : rb[0].act[0], cnt,
                                                                    (unsigned int)nA);
    x->launches += 1;
    CK(cudaGetLastError());
    if (x->sort_prefix) {
        size_t tmp = 0;
        CK(cub::DeviceRadixSort::SortPairs(nullptr, tmp, fp_in, fp_out, idx_in, rb[0].act[0], (int)n, 32, 64, st));
        CK(x->r_sort.need(tmp));
        CK(cub::DeviceRadixSort::SortPairs(x->r_sort.p, tmp, fp_in, fp_out, idx_in, rb[0].act[0], (int)n, 32, 64, st));
        x->launches += 5;     // cub: histogram + 4 onesweep passes (the top 32 fingerprint bits order the groups well enough)
    }
    if (max_blocks < 0) {
        unsigned long long mb = 0;
        CK(cudaMemcpyAsync(&mb, d_maxb, sizeof mb, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        max_blocks = (int64_t)mb;
    }
    int64_t rounds = (max_blocks + plain::kRoundBlocks - 1) / plain::kRoundBlocks;
    if (rounds < 1) rounds = 1;                       // round 0 also retires the prompts that have no full block
    const int nh = nB > 0 ? 2 : 1;
    const int per_sm_h = nh == 2 ? 2 : 4, per_sm_p = nh == 2 ? 2 : 4;
    if (!spec) {
        cudaStream_t strm[2] = {st, x->aux_stream[0]};
        if (nh == 2) { CK(cudaEventRecord(x->ev_fork, st)); CK(cudaStreamWaitEvent(x->aux_stream[0], x->ev_fork, 0)); }
        for (int64_t r = 0; r < rounds; ++r) {
            const int cur = (int)(r & 1);
            for (int hlf = 0; hlf < nh; ++hlf) {
                const int64_t m = hlf ? nB : nA;
                const unsigned hgrid = (unsigned)std::min<int64_t>((m + plain::kHashThreads - 1) / plain::kHashThreads, (int64_t)x->sm_count * per_sm_h);
                if (warp_walk && x->rounds_warp_stages == 3) {      // 128-thread CTAs, three stages: 55 KB, four CTAs per SM
                    const unsigned hg = (unsigned)std::min<int64_t>((m + 127) / 128, (int64_t)x->sm_count * 4);
                    plain::hash_round_kernel<16, true, 3><<<hg, 128, sizeof(plain::HashSmem<16>) / 8 / 2 * 3 * 4, strm[hlf]>>>(x->tv, a, rb[hlf], cur, (int)r);
                } else if (warp_walk) plain::hash_round_kernel<16, true, 2><<<hgrid, plain::kHashThreads, sizeof(plain::HashSmem<16>), strm[hlf]>>>(x->tv, a, rb[hlf], cur, (int)r);
                else if (x->rounds_lane_stages == 3 && n <= x->rounds_lane_stages_max) {
                    const unsigned hg = (unsigned)std::min<int64_t>((m + 127) / 128, (int64_t)x->sm_count * 4);
                    plain::hash_round_kernel<16, false, 3><<<hg, 128, sizeof(plain::HashSmem<16>) / 8 / 2 * 3 * 4, strm[hlf]>>>(x->tv, a, rb[hlf], cur, (int)r);
                } else plain::hash_round_kernel<16, false, 2><<<hgrid, plain::kHashThreads, sizeof(plain::HashSmem<16>), strm[hlf]>>>(x->tv, a, rb[hlf], cur, (int)r);
            }
            for (int hlf = 0; hlf < nh; ++hlf) {
                const int64_t m = hlf ? nB : nA;
                if (warp_walk) {      // warp per prompt: 8 prompts per CTA iteration
                    const unsigned pgrid = (unsigned)std::min<int64_t>((m + 7) / 8, (int64_t)x->sm_count * 8);
                    plain::probe_round_warp_kernel<<<pgrid, plain::kProbeThreads, 0, strm[hlf]>>>(x->tv, a, rb[hlf], cur, (int)r);
                    continue;
                }
                const unsigned pgrid = (unsigned)std::min<int64_t>((m + plain::kProbeThreads - 1) / plain::kProbeThreads, (int64_t)x->sm_count * per_sm_p);
                plain::probe_round_kernel<<<pgrid, plain::kProbeThreads, sizeof(plain::WalkSmem), strm[hlf]>>>(x->tv, a, rb[hlf], cur, (int)r);
            }
            x->launches += 2 * nh;
        }
        CK(cudaGetLastError());
        if (nh == 2) { CK(cudaEventRecord(x->ev_join[0], x->aux_stream[0])); CK(cudaStreamWaitEvent(st, x->ev_join[0], 0)); }
        return 0;
    }
    // speculative schedule: per half one H stream and one P stream; H(t+1) waits for P(t-1) (its list), P(t) for H(t) (its keys)
    cudaStream_t sH[2] = {st, x->aux_stream[0]}, sP[2] = {x->aux_stream[1], x->aux_stream[2]};
    CK(cudaEventRecord(x->ev_fork, st));
    for (int q = 0; q < 3; ++q) CK(cudaStreamWaitEvent(x->aux_stream[q], x->ev_fork, 0));
    for (int64_t r = 0; r < rounds; ++r) {
        const int cur = (int)(r & 1);
        for (int hlf = 0; hlf < nh; ++hlf) {
            const int64_t m = hlf ? nB : nA;
            plain::RoundBufs rr = rb[hlf];
            rr.keys += (size_t)cur * (size_t)n * plain::kRoundBlocks; rr.nbr += (size_t)cur * (size_t)n;
            if (r >= 2) CK(cudaStreamWaitEvent(sH[hlf], x->ev_spec_p[hlf][cur], 0));                 // P(r-2) built the list H(r) runs over ... (r-1)&1 == cur^1; see below
            const unsigned hgrid = (unsigned)std::min<int64_t>((m + plain::kHashThreads - 1) / plain::kHashThreads, (int64_t)x->sm_count * per_sm_h);
            plain::hash_round_kernel<16, false, 2><<<hgrid, plain::kHashThreads, sizeof(plain::HashSmem<16>), sH[hlf]>>>(x->tv, a, rr, cur, (int)r);
            CK(cudaEventRecord(x->ev_spec_h[hlf][cur], sH[hlf]));
            CK(cudaStreamWaitEvent(sP[hlf], x->ev_spec_h[hlf][cur], 0));
            CK(cudaMemsetAsync(rb[hlf].n_act + (cur ^ 1), 0, sizeof(unsigned int), sP[hlf]));      // the list P(r) appends to starts empty
            const unsigned pgrid = (unsigned)std::min<int64_t>((m + plain::kProbeThreads - 1) / plain::kProbeThreads, (int64_t)x->sm_count * per_sm_p);
            plain::probe_round_kernel<<<pgrid, plain::kProbeThreads, sizeof(plain::WalkSmem), sP[hlf]>>>(x->tv, a, rr, cur, (int)r);
            CK(cudaEventRecord(x->ev_spec_p[hlf][cur], sP[hlf]));
        }
        x->launches += 2 * nh;
    }
    CK(cudaGetLastError());
    if (nh == 2) { CK(cudaEventRecord(x->ev_join[0], sH[1])); CK(cudaStreamWaitEvent(st, x->ev_join[0], 0)); }
    for (int hlf = 0; hlf < nh; ++hlf) { CK(cudaEventRecord(x->ev_join[1 + hlf], sP[hlf])); CK(cudaStreamWaitEvent(st, x->ev_join[1 + hlf], 0)); }
    return 0;
}

// Launch the score kernel(s) over device-resident inputs.
int launch_score(kvidx* x, const uint32_t* d_tok, const int64_t* d_off, int64_t tok_base, int64_t n,
                 const uint32_t* d_model, uint32_t model0, const uint64_t* d_filter, const ScoreOut& o, cudaStream_t st,
                 int64_t max_blocks = -1) {
    if (n <= 0) return 0;
    if (int rc = check_shards(x)) return rc;
    if (x->score_kernel == 1 || x->tv.req_stamp) {
        // exact-LRU mode always takes this kernel: Lookup must touch every present key of a prompt (no early exit)
        const int T = 128;
        const unsigned long long stride = 1ull << 20;
        const unsigned long long sb = x->tv.req_stamp ? reserve_stamps(x, (unsigned long long)n * stride) : 0;
        score_kernel_v1<<<(unsigned)((n + T - 1) / T), T, 0, st>>>(x->tv, d_tok, d_off, tok_base, n, d_model, model0, d_filter,
                                                                  o.dense, o.sp_pods, o.sp_scores, o.sp_cnt, o.has_keys, sb, stride);
        x->launches += 1;
    } else if (x->tv.block_size == 16 && (x->score_path == 4 || (x->score_path == 0 && n <= x->coop_max))) {
        // small batches: a warp per prompt, warp-cooperative hashing, TMA-staged tokens (kernels_coop.cuh)
        ScoreArgs a{d_tok, d_off, tok_base, n, d_model, model0, d_filter, o.dense, o.sp_pods, o.sp_scores, o.sp_cnt, o.has_keys, nullptr};
        const int64_t ctas = std::min<int64_t>((n + kCoopWarps - 1) / kCoopWarps, (int64_t)x->sm_count * 4);
        coop_score_kernel<16><<<(unsigned)ctas, kCoopThreads, sizeof(CoopSmem), st>>>(x->tv, a);
        x->launches += 1;
    } else if (x->tv.block_size == 16 && n < (1ll << 31) && ((x->score_path >= 2 && x->score_path <= 3) || (x->score_path == 0 && n >= x->rounds_min))) {
        // path 2: plain rounds, 3: prefix classes; automatic: by batch size (the class pipeline has a fixed cost per round)
        const bool classes = x->score_path == 3 || (x->score_path == 0 && n >= x->classes_min);
        return classes ? launch_score_rounds(x, d_tok, d_off, tok_base, n, d_model, model0, d_filter, o, st, max_blocks)
                       : launch_score_rounds_plain(x, d_tok, d_off, tok_base, n, d_model, model0, d_filter, o, st, max_blocks);
    } else {
        uint64_t nl = 0;
        int rc = launch_score_tuned(x->tv, x->sm_count, d_tok, d_off, tok_base, n, d_model, model0, d_filter,
                                    o.dense, o.sp_pods, o.sp_scores, o.sp_cnt, o.has_keys, &x->d_cnt->pad, st, &nl);
        x->launches += nl;
        if (rc) return fail(KVIDX_ECUDA, "score launch failed: %s", cudaGetErrorString(cudaGetLastError()));
    }
    CK(cudaGetLastError());
    return 0;
}

int check_csr(const int64_t* off, int64_t n) {
    for (int64_t i = 0; i < n; ++i) if (off[i + 1] < off[i]) return fail(KVIDX_EINVAL, "tok_off not monotone at %lld", (long long)i);
    return 0;
}

// Host-buffer scoring.  Chunks of prompts flow through two slots and three streams so that the H2D copy
// of chunk c+1, the kernel of chunk c and the D2H of chunk c-1 overlap (two DMA engines + SMs):
//   copy_stream : staging -> d_tok/d_off/...      (records ev_h2d[slot])
//   x->stream   : score kernel                      (waits ev_h2d, records ev_k[slot])
//   d2h_stream  : results -> pinned host            (waits ev_k, records ev_done[slot])
// Caller buffers that are already pinned are used directly (no staging memcpy).
int score_host_locked(kvidx* x, const uint32_t* tok, const int64_t* tok_off, int64_t n, const uint32_t* model, uint32_t model0,
                      const uint64_t* filter, double* dense, uint16_t* sp_pods, double* sp_scores, uint8_t* sp_cnt, uint8_t* has_keys);

int score_host(kvidx* x, const uint32_t* tok, const int64_t* tok_off, int64_t n, const uint32_t* model, uint32_t model0,
               const uint64_t* filter, double* dense, uint16_t* sp_pods, double* sp_scores, uint8_t* sp_cnt, uint8_t* has_keys) {
    if (n < 0 || !tok_off) return fail(KVIDX_EINVAL, "bad arguments");
    if (n == 0) return 0;
    int rc = check_csr(tok_off, n);
    if (rc) return rc;
    if (tok_off[n] > tok_off[0] && !tok) return fail(KVIDX_EINVAL, "NULL tokens");
    ReadGuard g(x);
    CK(cudaSetDevice(x->device));
    rc = score_host_locked(x, tok, tok_off, n, model, model0, filter, dense, sp_pods, sp_scores, sp_cnt, has_keys);
    if (rc) {
        // an error in the middle of the pipeline: copies may still be reading the caller's buffers and this handle's staging
        // slots -- drain the three streams before the caller gets its buffers back (the error text is kept)
        const std::string keep = g_err;
        cudaStreamSynchronize(x->copy_stream); cudaStreamSynchronize(x->stream); cudaStreamSynchronize(x->d2h_stream);
        cudaGetLastError();
        g_err = keep;
    }
    return rc;
}

// A handful of prompts (the size of one RPC, or of what the submission queue gathers at low load): no copy engine, no second
// stream.  The cooperative kernel reads the tokens straight from PINNED HOST memory -- its TMA bulk copies fetch each 2 KB
// chunk over PCIe one chunk ahead of the chain -- and writes the result rows straight back into pinned host memory; the call
// is one launch and one stream synchronisation.  (cudaMallocHost memory is device-addressable under UVA.)
int score_host_zerocopy(kvidx* x, const uint32_t* tok, const int64_t* tok_off, int64_t n, const uint32_t* model, uint32_t model0,
                        const uint64_t* filter, double* dense, uint16_t* sp_pods, double* sp_scores, uint8_t* sp_cnt, uint8_t* has_keys) {
    const uint32_t P = x->tv.max_pods, FW = x->tv.filter_words;
    const bool sparse = sp_cnt != nullptr;
    const int64_t tb = tok_off[0], nt = tok_off[n] - tb;
    const bool tok_pinned = nt == 0 || is_device_accessible_host(tok);
    const size_t a8 = 63;
    const size_t in_bytes = (((size_t)(n + 1) * 8 + a8) & ~a8) + (model ? (((size_t)n * 4 + a8) & ~a8) : 0) + (filter ? (((size_t)n * FW * 8 + a8) & ~a8) : 0) +
                            (tok_pinned ? 0 : (size_t)nt * 4 + 64);
    CK(x->h_stage[0].need(in_bytes + 64));
    uint8_t* hs = x->h_stage[0].as<uint8_t>();
    size_t o = 0;
    int64_t* h_off = reinterpret_cast<int64_t*>(hs); memcpy(h_off, tok_off, (size_t)(n + 1) * 8); o += ((size_t)(n + 1) * 8 + a8) & ~a8;
    const uint32_t* h_model = nullptr; const uint64_t* h_filter = nullptr;
    if (model) { memcpy(hs + o, model, (size_t)n * 4); h_model = reinterpret_cast<const uint32_t*>(hs + o); o += ((size_t)n * 4 + a8) & ~a8; }
    if (filter) { memcpy(hs + o, filter, (size_t)n * FW * 8); h_filter = reinterpret_cast<const uint64_t*>(hs + o); o += ((size_t)n * FW * 8 + a8) & ~a8; }
    const uint32_t* h_tok = tok ? tok + tb : nullptr;                     // token of absolute index tb
    if (!tok_pinned) { memcpy(hs + o, tok + tb, (size_t)nt * 4); h_tok = reinterpret_cast<const uint32_t*>(hs + o); }
    // results: straight into the caller's buffers when those are pinned, else through the pinned staging
    const bool out_direct = sparse ? (is_device_accessible_host(sp_pods) && is_device_accessible_host(sp_scores) && is_device_accessible_host(sp_cnt) &&
                                      (!has_keys || is_device_accessible_host(has_keys)))
                                   : (is_device_accessible_host(dense) && (!has_keys || is_device_accessible_host(has_keys)));
    ScoreArgs a{h_tok, h_off, tb, n, h_model, model0, h_filter, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    uint8_t* ho = nullptr;
    if (out_direct) { a.dense = dense; a.sp_pods = sp_pods; a.sp_scores = sp_scores; a.sp_cnt = sp_cnt; a.has_keys = has_keys; }
    else {
        const size_t out_bytes = sparse ? (size_t)n * (kMaxEnt * 10 + 2) : (size_t)n * (P * 8 + 1);
        CK(x->h_out[0].need(out_bytes + 64));
        ho = x->h_out[0].as<uint8_t>();
        if (!sparse) { a.dense = reinterpret_cast<double*>(ho); a.has_keys = ho + (size_t)n * P * 8; }
        else { a.sp_scores = reinterpret_cast<double*>(ho); a.sp_pods = reinterpret_cast<uint16_t*>(ho + (size_t)n * kMaxEnt * 8); a.sp_cnt = ho + (size_t)n * kMaxEnt * 10; a.has_keys = a.sp_cnt + n; }
    }
    if (int rc = check_shards(x)) return rc;
    CK(cudaStreamWaitEvent(x->stream, x->ev_write, 0));
    const int64_t ctas = std::min<int64_t>((n + kCoopWarps - 1) / kCoopWarps, (int64_t)x->sm_count * 4);
    coop_score_kernel<16><<<(unsigned)ctas, kCoopThreads, sizeof(CoopSmem), x->stream>>>(x->tv, a);
    x->launches += 1;
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(x->stream));
    if (!out_direct) {
        if (!sparse) { memcpy(dense, ho, (size_t)n * P * 8); if (has_keys) memcpy(has_keys, ho + (size_t)n * P * 8, (size_t)n); }
        else {
            memcpy(sp_scores, ho, (size_t)n * kMaxEnt * 8); memcpy(sp_pods, ho + (size_t)n * kMaxEnt * 8, (size_t)n * kMaxEnt * 2);
            memcpy(sp_cnt, ho + (size_t)n * kMaxEnt * 10, (size_t)n);
            if (has_keys) memcpy(has_keys, ho + (size_t)n * kMaxEnt * 10 + n, (size_t)n);
        }
    }
    return 0;
}

int score_host_locked(kvidx* x, const uint32_t* tok, const int64_t* tok_off, int64_t n, const uint32_t* model, uint32_t model0,
                      const uint64_t* filter, double* dense, uint16_t* sp_pods, double* sp_scores, uint8_t* sp_cnt, uint8_t* has_keys) {
    int rc = 0;
    if (n <= x->zerocopy_max && x->tv.block_size == 16 && !x->tv.req_stamp && x->score_kernel != 1 && (x->score_path == 0 || x->score_path == 4) &&
        tok_off[n] - tok_off[0] <= (4ll << 20))
        return score_host_zerocopy(x, tok, tok_off, n, model, model0, filter, dense, sp_pods, sp_scores, sp_cnt, has_keys);
    const uint32_t P = x->tv.max_pods, FW = x->tv.filter_words;
    const bool sparse = sp_cnt != nullptr;
    const size_t out_row = sparse ? (size_t)kMaxEnt * (sizeof(double) + sizeof(uint16_t)) + 2 : (size_t)P * sizeof(double) + 1;
    // chunk = up to 32 Mi tokens (128 MiB): the H2D copy of a chunk (2.3 ms at 55 GB/s) must outlast its kernel
    // (~1 ms for 8192 x 4K-token prompts) for the pipeline to be PCIe bound; 8 Mi-token chunks were kernel bound.
    const int64_t kMaxTokChunk = x->host_chunk_tokens;
    const int64_t kMaxRowsChunk = std::max<int64_t>(1, (128ll << 20) / (int64_t)out_row);
    const bool tok_pinned = tok && is_device_accessible_host(tok);
    const bool out_pinned = !sparse && dense && is_device_accessible_host(dense) && (!has_keys || is_device_accessible_host(has_keys));
    cudaStream_t s_k = x->stream, s_in = x->copy_stream, s_out = x->d2h_stream;
    struct Pending { bool live = false; int64_t i0 = 0, cnt = 0; } pend[2];
    auto drain = [&](int s) -> int {
        if (!pend[s].live) return 0;
        CK(cudaEventSynchronize(x->ev_done[s]));
        pend[s].live = false;
        if (out_pinned) return 0;                 // results were written straight into the caller's buffers
        const int64_t c = pend[s].cnt, b = pend[s].i0;
        const uint8_t* ho = x->h_out[s].as<uint8_t>();
        if (!sparse) {
            memcpy(dense + b * (int64_t)P, ho, (size_t)c * P * sizeof(double));
            if (has_keys) memcpy(has_keys + b, ho + (size_t)c * P * sizeof(double), (size_t)c);
        } else {
            size_t o = 0;
            memcpy(sp_scores + b * kMaxEnt, ho + o, (size_t)c * kMaxEnt * sizeof(double)); o += (size_t)c * kMaxEnt * sizeof(double);
            memcpy(sp_pods + b * kMaxEnt, ho + o, (size_t)c * kMaxEnt * sizeof(uint16_t)); o += (size_t)c * kMaxEnt * sizeof(uint16_t);
            memcpy(sp_cnt + b, ho + o, (size_t)c); o += (size_t)c;
            if (has_keys) memcpy(has_keys + b, ho + o, (size_t)c);
        }
        return 0;
    };
    // the first chunk must see everything already queued on the handle's stream, and asynchronous write batches issued before
    CK(cudaStreamWaitEvent(s_k, x->ev_write, 0));
    CK(cudaEventRecord(x->ev_k[0], s_k));
    CK(cudaStreamWaitEvent(s_in, x->ev_k[0], 0));
    int64_t i0 = 0;
    int slot = 0;
    while (i0 < n) {
        int64_t i1 = i0;
        while (i1 < n && (i1 - i0) < kMaxRowsChunk && (tok_off[i1 + 1] - tok_off[i0]) <= kMaxTokChunk) ++i1;
        if (i1 == i0) i1 = i0 + 1;                // a single prompt larger than the chunk budget
        const int64_t c = i1 - i0, tb = tok_off[i0], nt = tok_off[i1] - tb;
        rc = drain(slot);                         // slot's previous results are home; its device buffers are free
        if (rc) return rc;
        CK(x->d_tok[slot].need((size_t)std::max<int64_t>(nt, 1) * 4 + 64));
        CK(x->d_off[slot].need((size_t)(c + 1) * 8));
        const size_t stage_bytes = (size_t)(c + 1) * 8 + (model ? (size_t)c * 4 : 0) + (filter ? (size_t)c * FW * 8 : 0) +
                                   (tok_pinned ? 0 : (size_t)nt * 4);
        CK(x->h_stage[slot].need(stage_bytes + 64));
        uint8_t* hs = x->h_stage[slot].as<uint8_t>();
        size_t o = 0;
        memcpy(hs + o, tok_off + i0, (size_t)(c + 1) * 8);
        CK(cudaMemcpyAsync(x->d_off[slot].p, hs + o, (size_t)(c + 1) * 8, cudaMemcpyHostToDevice, s_in)); o += (size_t)(c + 1) * 8;
        const uint32_t* dm = nullptr; const uint64_t* df = nullptr;
        if (model) {
            CK(x->d_model[slot].need((size_t)c * 4));
            memcpy(hs + o, model + i0, (size_t)c * 4);
            CK(cudaMemcpyAsync(x->d_model[slot].p, hs + o, (size_t)c * 4, cudaMemcpyHostToDevice, s_in)); o += (size_t)c * 4;
            dm = x->d_model[slot].as<uint32_t>();
        }
        if (filter) {
            CK(x->d_filter[slot].need((size_t)c * FW * 8));
            memcpy(hs + o, filter + i0 * FW, (size_t)c * FW * 8);
            CK(cudaMemcpyAsync(x->d_filter[slot].p, hs + o, (size_t)c * FW * 8, cudaMemcpyHostToDevice, s_in)); o += (size_t)c * FW * 8;
            df = x->d_filter[slot].as<uint64_t>();
        }
        if (nt > 0) {
            const uint32_t* src = tok + tb;
            if (!tok_pinned) { memcpy(hs + o, src, (size_t)nt * 4); src = reinterpret_cast<const uint32_t*>(hs + o); }
            CK(cudaMemcpyAsync(x->d_tok[slot].p, src, (size_t)nt * 4, cudaMemcpyHostToDevice, s_in));
        }
        CK(cudaEventRecord(x->ev_h2d[slot], s_in));
        // kernel
        CK(x->d_out[slot].need((size_t)c * out_row + 64));
        if (!out_pinned) CK(x->h_out[slot].need((size_t)c * out_row + 64));
        uint8_t* dout = x->d_out[slot].as<uint8_t>();
        ScoreOut so{};
        if (!sparse) {
            so.dense = reinterpret_cast<double*>(dout);
            so.has_keys = dout + (size_t)c * P * sizeof(double);
        } else {
            size_t q = 0;
            so.sp_scores = reinterpret_cast<double*>(dout + q); q += (size_t)c * kMaxEnt * sizeof(double);
            so.sp_pods = reinterpret_cast<uint16_t*>(dout + q); q += (size_t)c * kMaxEnt * sizeof(uint16_t);
            so.sp_cnt = dout + q; q += (size_t)c;
            so.has_keys = dout + q; q += (size_t)c;
        }
        CK(cudaStreamWaitEvent(s_k, x->ev_h2d[slot], 0));
        int64_t maxb = 0;
        for (int64_t q = i0; q < i1; ++q) maxb = std::max<int64_t>(maxb, (tok_off[q + 1] - tok_off[q]) / x->tv.block_size);
        rc = launch_score(x, x->d_tok[slot].as<uint32_t>(), x->d_off[slot].as<int64_t>(), tb, c, dm, model0, df, so, s_k, maxb);
        if (rc) return rc;
        CK(cudaEventRecord(x->ev_k[slot], s_k));
        // (slot reuse is safe without further stream waits: drain(slot) host-waits on ev_done[slot], which is
        //  recorded after this kernel and its read-back)
        // read back
        CK(cudaStreamWaitEvent(s_out, x->ev_k[slot], 0));
        if (out_pinned) {
            CK(cudaMemcpyAsync(dense + i0 * (int64_t)P, so.dense, (size_t)c * P * sizeof(double), cudaMemcpyDeviceToHost, s_out));
            if (has_keys) CK(cudaMemcpyAsync(has_keys + i0, so.has_keys, (size_t)c, cudaMemcpyDeviceToHost, s_out));
        } else {
            const size_t total = sparse ? (size_t)c * (kMaxEnt * (sizeof(double) + sizeof(uint16_t)) + 2) : (size_t)c * (P * sizeof(double) + 1);
            CK(cudaMemcpyAsync(x->h_out[slot].p, dout, total, cudaMemcpyDeviceToHost, s_out));
        }
        CK(cudaEventRecord(x->ev_done[slot], s_out));
        pend[slot].live = true; pend[slot].i0 = i0; pend[slot].cnt = c;
        slot ^= 1;
        i0 = i1;
    }
    rc = drain(slot); if (rc) return rc;
    rc = drain(slot ^ 1); if (rc) return rc;
    return 0;
}


int pin_need(void** p, size_t* cap, size_t n) {
    if (n <= *cap) return 0;
    if (*p) cudaFreeHost(*p);
    *p = nullptr; *cap = 0;
    const size_t want = n + n / 4 + 4096;
    CK(cudaMallocHost(p, want));
    *cap = want;
    return 0;
}

// Several callers' requests as one batch: tokens are gathered into one pinned buffer (CSR rebuilt), per-prompt model ids and
// filter rows materialised if any request carries them, one pass through score_host, results scattered to the owners.
int run_combined(kvidx* x, SubmitQueue* q, const std::vector<SubmitReq*>& batch) {
    const uint32_t P = x->tv.max_pods, FW = x->tv.filter_words;
    int64_t n = 0, nt = 0;
    bool any_model = false, any_filter = false, want_has = false;
    for (const SubmitReq* r : batch) {
        n += r->n; nt += r->tok_off[r->n] - r->tok_off[0];
        any_model |= r->model != nullptr || r->model0 != batch[0]->model0; any_filter |= r->filter != nullptr; want_has |= r->has_keys != nullptr;
    }
    const bool sparse = batch[0]->sp_cnt != nullptr;
    const size_t in_bytes = (size_t)nt * 4 + 64 + (size_t)(n + 1) * 8 + (any_model ? (size_t)n * 4 : 0) + 64 + (any_filter ? (size_t)n * FW * 8 : 0);
    const size_t out_bytes = sparse ? (size_t)n * (kMaxEnt * 10 + 2) + 64 : (size_t)n * (P * 8 + 1) + 64;
    int rc = pin_need(&q->h_in, &q->h_in_cap, in_bytes); if (rc) return rc;
    rc = pin_need(&q->h_out, &q->h_out_cap, out_bytes); if (rc) return rc;
    uint8_t* hi = static_cast<uint8_t*>(q->h_in);
    uint32_t* tok = reinterpret_cast<uint32_t*>(hi); size_t o = ((size_t)nt * 4 + 63) & ~(size_t)63;
    int64_t* off = reinterpret_cast<int64_t*>(hi + o); o += (size_t)(n + 1) * 8;
    uint32_t* model = nullptr; uint64_t* filter = nullptr;
    if (any_model) { model = reinterpret_cast<uint32_t*>(hi + o); o += ((size_t)n * 4 + 63) & ~(size_t)63; }
    if (any_filter) { filter = reinterpret_cast<uint64_t*>(hi + o); }
    int64_t at = 0, tat = 0;
    off[0] = 0;
    for (const SubmitReq* r : batch) {
        const int64_t tb = r->tok_off[0], rt = r->tok_off[r->n] - tb;
        if (rt > 0) memcpy(tok + tat, r->tok + tb, (size_t)rt * 4);
        for (int64_t i = 0; i < r->n; ++i) off[at + i + 1] = tat + (r->tok_off[i + 1] - tb);
        if (model) for (int64_t i = 0; i < r->n; ++i) model[at + i] = r->model ? r->model[i] : r->model0;
        if (filter) { if (r->filter) memcpy(filter + (size_t)at * FW, r->filter, (size_t)r->n * FW * 8); else memset(filter + (size_t)at * FW, 0, (size_t)r->n * FW * 8); }
        at += r->n; tat += rt;
    }
    uint8_t* ho = static_cast<uint8_t*>(q->h_out);
    if (!sparse) {
        double* dense = reinterpret_cast<double*>(ho); uint8_t* has = ho + (size_t)n * P * 8;
        rc = score_host(x, tok, off, n, model, batch[0]->model0, filter, dense, nullptr, nullptr, nullptr, has);
        if (rc) return rc;
        at = 0;
        for (SubmitReq* r : batch) {
            memcpy(r->dense, dense + (size_t)at * P, (size_t)r->n * P * 8);
            if (r->has_keys) memcpy(r->has_keys, has + at, (size_t)r->n);
            at += r->n;
        }
    } else {
        double* sc = reinterpret_cast<double*>(ho); uint16_t* pods = reinterpret_cast<uint16_t*>(ho + (size_t)n * kMaxEnt * 8);
        uint8_t* cnt = ho + (size_t)n * kMaxEnt * 10; uint8_t* has = cnt + n;
        rc = score_host(x, tok, off, n, model, batch[0]->model0, filter, nullptr, pods, sc, cnt, has);
        if (rc) return rc;
        at = 0;
        for (SubmitReq* r : batch) {
            memcpy(r->sp_scores, sc + (size_t)at * kMaxEnt, (size_t)r->n * kMaxEnt * 8);
            memcpy(r->sp_pods, pods + (size_t)at * kMaxEnt, (size_t)r->n * kMaxEnt * 2);
            memcpy(r->sp_cnt, cnt + at, (size_t)r->n);
            if (r->has_keys) memcpy(r->has_keys, has + at, (size_t)r->n);
            at += r->n;
        }
    }
    (void)want_has;
    return 0;
}

int submit_score(kvidx* x, const uint32_t* tok, const int64_t* tok_off, int64_t n, const uint32_t* model, uint32_t model0,
                 const uint64_t* filter, double* dense, uint16_t* sp_pods, double* sp_scores, uint8_t* sp_cnt, uint8_t* has_keys) {
    SubmitQueue* q = x->queue;
    // big batches and malformed calls go straight through (score_host validates); so does everything when the queue is off
    if (!q || !q->enabled || n <= 0 || !tok_off || n > q->max_prompts / 4)
        return score_host(x, tok, tok_off, n, model, model0, filter, dense, sp_pods, sp_scores, sp_cnt, has_keys);
    if (int rc = check_csr(tok_off, n)) return rc;
    if (tok_off[n] > tok_off[0] && !tok) return fail(KVIDX_EINVAL, "NULL tokens");
    SubmitReq me;
    me.tok = tok; me.tok_off = tok_off; me.n = n; me.model = model; me.model0 = model0; me.filter = filter;
    me.dense = dense; me.sp_pods = sp_pods; me.sp_scores = sp_scores; me.sp_cnt = sp_cnt; me.has_keys = has_keys;
    bool lead;
    {
        std::lock_guard<std::mutex> lk(q->mu);
        q->pending.push_back(&me);
        lead = !q->leader_active;
        if (lead) q->leader_active = true;
    }
    if (!lead) {
        std::unique_lock<std::mutex> lk(me.m);
        me.cv.wait(lk, [&] { return me.done || me.promote; });
        if (me.done) { if (me.rc) g_err = me.err; return me.rc; }
        // promoted: the previous leader left with work still queued (this request among it)
    }
    for (;;) {
        std::vector<SubmitReq*> batch;
        {
            std::unique_lock<std::mutex> lk(q->mu);
            if (me.done) {                                    // my own request is served: hand the queue to the next owner
                SubmitReq* next = nullptr;
                if (q->pending.empty()) q->leader_active = false;
                else next = q->pending.front();
                lk.unlock();
                if (next) { std::lock_guard<std::mutex> g2(next->m); next->promote = true; next->cv.notify_one(); }
                if (me.rc) g_err = me.err;
                return me.rc;
            }
            const bool sparse = q->pending.front()->sp_cnt != nullptr;
            int64_t np = 0, ntok = 0;
            for (auto it = q->pending.begin(); it != q->pending.end();) {
                SubmitReq* r = *it;
                const int64_t rt = r->tok_off[r->n] - r->tok_off[0];
                if ((r->sp_cnt != nullptr) != sparse) { ++it; continue; }
                if (!batch.empty() && (np + r->n > q->max_prompts || ntok + rt > q->max_tokens)) break;
                batch.push_back(r); np += r->n; ntok += rt;
                it = q->pending.erase(it);
            }
        }
        int rc;
        if (batch.size() == 1) {
            SubmitReq* r = batch[0];
            rc = score_host(x, r->tok, r->tok_off, r->n, r->model, r->model0, r->filter, r->dense, r->sp_pods, r->sp_scores, r->sp_cnt, r->has_keys);
        } else {
            rc = run_combined(x, q, batch);
            q->coalesced += batch.size();
        }
        q->batches += 1;
        for (SubmitReq* r : batch) {
            if (r == &me) { me.rc = rc; if (rc) me.err = g_err; me.done = true; continue; }
            std::lock_guard<std::mutex> g2(r->m);             // (notify under the lock: the owner cannot return -- and destroy r -- before we are done with it)
            r->rc = rc; if (rc) r->err = g_err; r->done = true;
            r->cv.notify_one();
        }
    }
}

SubmitQueue* new_submit_queue() {
    SubmitQueue* q = new SubmitQueue();
    if (const char* k = getenv("KVIDX_SUBMIT_QUEUE")) q->enabled = atoi(k) != 0;
    return q;
}
void delete_submit_queue(SubmitQueue* q) {
    if (!q) return;
    if (q->h_in) cudaFreeHost(q->h_in);
    if (q->h_out) cudaFreeHost(q->h_out);
    delete q;
}
uint64_t submit_queue_coalesced(SubmitQueue* q) { return q->coalesced.load(); }

// TokensToKVBlockKeys for a CSR batch on the device (token_processor.go:141-162)
int launch_hash_keys(kvidx* x, const uint32_t* d_tok, const int64_t* d_off, int64_t tok_base, int64_t n, const uint64_t* d_parent,
                     const uint8_t* d_pv, const int64_t* d_koff, uint64_t* d_keys, cudaStream_t st) {
    const int T = 128;
    hash_keys_kernel_v1<<<(unsigned)((n + T - 1) / T), T, 0, st>>>(x->tv, d_tok, d_off, tok_base, n, d_parent, d_pv, d_koff, 0, d_keys);
    x->launches += 1;
    CK(cudaGetLastError());
    return 0;
}

}  // namespace

// ------------------------------------------------------------------------------------------
extern "C" {

int kvidx_abi_version(void) { return KVIDX_ABI_VERSION; }

void kvidx_config_default(kvidx_config_t* c) {
    if (!c) return;
    memset(c, 0, sizeof *c);
    c->struct_size = sizeof *c;
    c->device = 0;
    c->block_size = 16;                      // token_processor.go:31
    c->pods_per_key = KVIDX_MAX_PODS_PER_KEY; // in_memory.go:34
    c->init_hash = kFnvOffset;               // FNV64a("")  (HashSeed default "", token_processor.go:48)
    c->capacity = 1ull << 20;
    c->table_slots = 0;
    c->max_pods = 256;
    c->n_tier_weights = 2;
    for (auto& w : c->tier_weight) w = 1.0;
    c->tier_weight[0] = 1.0;                 // "gpu"  backend.go:28
    c->tier_weight[1] = 0.8;                 // "cpu"  backend.go:29
    c->lru_exact = 0;
}

const char* kvidx_last_error(kvidx_t*) { return g_err.c_str(); }

uint64_t kvidx_fnv64a(const void* data, size_t n) {
    const uint8_t* p = static_cast<const uint8_t*>(data);
    uint64_t h = kFnvOffset;
    for (size_t i = 0; i < n; ++i) h = (h ^ p[i]) * kFnvPrime;
    return h;
}
uint32_t kvidx_fnv32a(const void* data, size_t n) {
    const uint8_t* p = static_cast<const uint8_t*>(data);
    uint32_t h = 0x811C9DC5u;
    for (size_t i = 0; i < n; ++i) h = (h ^ p[i]) * 0x01000193u;
    return h;
}
uint32_t kvidx_queue_index(const char* pod, size_t n, uint32_t concurrency) {
    return concurrency ? kvidx_fnv32a(pod, n) % concurrency : 0;
}

void* kvidx_host_alloc(size_t bytes) { void* p = nullptr; if (cudaMallocHost(&p, bytes) != cudaSuccess) { cudaGetLastError(); return nullptr; } return p; }
void kvidx_host_free(void* p) { if (p) cudaFreeHost(p); }

namespace {
int create_impl(const kvidx_config_t& c, kvidx* x) {
    CK(cudaSetDevice(c.device));
    x->cfg = c;
    x->device = c.device;
    cudaDeviceProp prop{};
    CK(cudaGetDeviceProperties(&prop, c.device));
    x->sm_count = prop.multiProcessorCount;
    if (prop.major < 10) return fail(KVIDX_ECUDA, "device sm_%d%d is not Blackwell; libkvidx is built for sm_100a only", prop.major, prop.minor);
    CK(cudaStreamCreateWithFlags(&x->own_stream, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&x->own_wstream, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&x->copy_stream, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&x->d2h_stream, cudaStreamNonBlocking));
    for (int q = 0; q < kMaxParts - 1; ++q) {
        CK(cudaStreamCreateWithFlags(&x->aux_stream[q], cudaStreamNonBlocking));
        CK(cudaEventCreateWithFlags(&x->ev_join[q], cudaEventDisableTiming));
    }
    CK(cudaEventCreateWithFlags(&x->ev_fork, cudaEventDisableTiming));
    CK(cudaStreamCreateWithFlags(&x->g_stream, cudaStreamNonBlocking));
    {
        int lo = 0, hi = 0;
        CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        for (int q = 0; q < kMaxParts; ++q) {
            CK(cudaStreamCreateWithPriority(&x->hp_stream[q], cudaStreamNonBlocking, hi));
            CK(cudaStreamCreateWithPriority(&x->lp_stream[q], cudaStreamNonBlocking, lo));
        }
    }
    for (int q = 0; q < kMaxParts; ++q) { CK(cudaEventCreateWithFlags(&x->ev_g[q], cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&x->ev_r[q], cudaEventDisableTiming)); }
    CK(cudaEventCreateWithFlags(&x->ev_write, cudaEventDisableTiming));
    for (int a_ = 0; a_ < 2; ++a_) for (int b_ = 0; b_ < 2; ++b_) {
        CK(cudaEventCreateWithFlags(&x->ev_spec_h[a_][b_], cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&x->ev_spec_p[a_][b_], cudaEventDisableTiming));
    }
    x->stream = x->own_stream;
    for (int i = 0; i < 2; ++i) {
        CK(cudaEventCreateWithFlags(&x->ev_h2d[i], cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&x->ev_done[i], cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&x->ev_k[i], cudaEventDisableTiming));
    }
    uint32_t shard_bits = 0;
    if (c.shard_count > 1) {
        if (c.shard_count > 8 || (c.shard_count & (c.shard_count - 1)) || c.shard_rank >= c.shard_count) return fail(KVIDX_EINVAL, "shard_count must be a power of two <= 8 and shard_rank < shard_count");
        while ((1u << shard_bits) < c.shard_count) ++shard_bits;
    }
    const uint64_t per_shard = c.shard_count > 1 ? (c.capacity + c.shard_count - 1) / c.shard_count : c.capacity;
    uint64_t slots = c.table_slots ? pow2ceil(c.table_slots) : pow2ceil(std::max<uint64_t>(4 * per_shard, 1024));
    if (slots < 1024) slots = 1024;
    TableView& t = x->tv;
    t.req_mask = slots - 1; t.eng_mask = slots - 1;
    t.capacity = c.capacity; t.init_hash = c.init_hash; t.block_size = c.block_size; t.pods_per_key = c.pods_per_key;
    t.max_pods = c.max_pods; t.filter_words = (c.max_pods + 63) / 64;
    for (int i = 0; i < 16; ++i) t.weight[i] = (uint32_t)i < c.n_tier_weights ? c.tier_weight[i] : 1.0;
    t.req_stamp = nullptr;
    int rc = alloc_tables(x, slots, slots, &t.req, &t.eng, x->stream);
    if (rc) return rc;
    if (c.lru_exact) {
        if (c.shard_count > 1) return fail(KVIDX_EINVAL, "lru_exact is not supported on a sharded index");
        CK(cudaMalloc((void**)&t.req_stamp, slots * 8));
        CK(cudaMemsetAsync(t.req_stamp, 0, slots * 8, x->stream));
    }
    CK(cudaMalloc((void**)&x->d_cnt, sizeof(Counters)));
    CK(cudaMemsetAsync(x->d_cnt, 0, sizeof(Counters), x->stream));
    CK(cudaMallocHost((void**)&x->h_cnt, sizeof(Counters)));
    memset(x->h_cnt, 0, sizeof(Counters));
    CK(cudaMalloc((void**)&x->d_cnt_all, sizeof(Counters) * 8));
    CK(cudaMallocHost((void**)&x->h_cnt_all, sizeof(Counters) * 8));
    t.cnt = x->d_cnt;
    t.shard_bits = shard_bits; t.shard_rank = c.shard_count > 1 ? c.shard_rank : 0;
    for (int i = 0; i < 8; ++i) { t.req_peer[i] = nullptr; t.eng_peer[i] = nullptr; t.cnt_peer[i] = nullptr; }
    t.req_peer[t.shard_rank] = t.req; t.eng_peer[t.shard_rank] = t.eng; t.cnt_peer[t.shard_rank] = t.cnt;
    CK(cudaStreamSynchronize(x->stream));
    if (const char* k = getenv("KVIDX_SCORE_KERNEL")) x->score_kernel = (k[0] == 'v' ? atoi(k + 1) : atoi(k)) == 1 ? 1 : 2;
    if (const char* k = getenv("KVIDX_SCORE_PATH")) x->score_path = !strcmp(k, "fused") ? 1 : !strcmp(k, "rounds") ? 2 : !strcmp(k, "classes") ? 3 : !strcmp(k, "coop") ? 4 : 0;
    if (const char* k = getenv("KVIDX_COOP_MAX")) x->coop_max = atoll(k);
    if (const char* k = getenv("KVIDX_ZEROCOPY_MAX")) x->zerocopy_max = atoll(k);
    if (const char* k = getenv("KVIDX_ROUNDS_SPEC")) x->rounds_spec = atoi(k) != 0;
    if (const char* k = getenv("KVIDX_ROUNDS_SPEC_MAX")) x->rounds_spec_max = atoll(k);
    if (const char* k = getenv("KVIDX_ROUNDS_MIN")) x->rounds_min = atoll(k);
    if (const char* k = getenv("KVIDX_CLASSES_MIN")) x->classes_min = atoll(k);
    if (const char* k = getenv("KVIDX_CLASSES_SHARING")) x->classes_min_sharing = atof(k);
    if (const char* k = getenv("KVIDX_SORT_PREFIX")) x->sort_prefix = atoi(k) != 0;
    if (const char* k = getenv("KVIDX_ROUNDS_OVERLAP")) x->rounds_overlap = atoi(k) != 0;
    if (const char* k = getenv("KVIDX_ROUNDS_GRID")) sscanf(k, "%d,%d,%d,%d,%d", &x->rounds_grid[0], &x->rounds_grid[1], &x->rounds_grid[2], &x->rounds_grid[3], &x->rounds_grid[4]);
    if (const char* k = getenv("KVIDX_HOST_CHUNK_TOKENS")) x->host_chunk_tokens = std::max<int64_t>(1 << 16, atoll(k));
    if (const char* k = getenv("KVIDX_ROUNDS_TRACE")) x->rounds_trace = atoi(k);
    if (const char* k = getenv("KVIDX_ROUNDS_PARTS")) x->rounds_parts = atoi(k);
    if (const char* k = getenv("KVIDX_GROUP_SERIAL")) x->group_serial = atoi(k);
    if (const char* k = getenv("KVIDX_HASH_PREFETCH")) x->hash_prefetch = atoi(k);
    if (const char* k = getenv("KVIDX_PEER_PAIR")) x->tv.peer_pair = atoi(k) != 0;
    if (const char* k = getenv("KVIDX_GROUP_CTAS")) x->group_ctas = atoi(k);
    if (const char* k = getenv("KVIDX_ROUNDS_WARP")) x->rounds_warp = atoi(k) != 0;
    if (const char* k = getenv("KVIDX_ROUNDS_WARP_MAX")) x->rounds_warp_max = atoll(k);
    if (const char* k = getenv("KVIDX_ROUNDS_LANE_STAGES")) x->rounds_lane_stages = atoi(k) == 3 ? 3 : 2;
    if (const char* k = getenv("KVIDX_ROUNDS_WARP_STAGES")) x->rounds_warp_stages = atoi(k) == 3 ? 3 : 2;
    if (const char* k = getenv("KVIDX_SMALL_CTA")) { const int v = atoi(k); if (v == 32 || v == 64 || v == 128 || v == 256) x->small_cta = v; }
    if (const char* k = getenv("KVIDX_GROUP_SERIAL_GRID")) x->group_serial_grid = std::max(1, atoi(k));
    if (const char* k = getenv("KVIDX_ROUNDS_DEDUP")) x->rounds_dedup = atoi(k);
    if (const char* k = getenv("KVIDX_ROUNDS_OVERLAP_MIN")) x->rounds_overlap_min = atoll(k);
    if (const char* k = getenv("KVIDX_WRITE_PHASE1")) x->write_phase1 = atoi(k);
    if (const char* k = getenv("KVIDX_GROUP_TMA")) x->group_tma = atoi(k) != 0;
    x->queue = new_submit_queue();
    if (rounds_init() || plain::rounds_init() || coop_init()) return fail(KVIDX_ECUDA, "kernel attribute setup failed: %s", cudaGetErrorString(cudaGetLastError()));
    if (score_tuned_init()) return fail(KVIDX_ECUDA, "kernel attribute setup failed: %s", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
}  // namespace

int kvidx_create(const kvidx_config_t* cfg_in, kvidx_t** out) {
    if (!out) return fail(KVIDX_EINVAL, "out is NULL");
    *out = nullptr;
    kvidx_config_t c;
    kvidx_config_default(&c);
    if (cfg_in) {
        if (cfg_in->struct_size == 0 || cfg_in->struct_size > sizeof c) return fail(KVIDX_EINVAL, "bad struct_size");
        memcpy(&c, cfg_in, cfg_in->struct_size);
        c.struct_size = sizeof c;
    }
    if (c.block_size == 0) c.block_size = 16;
    if (c.pods_per_key == 0) c.pods_per_key = KVIDX_MAX_PODS_PER_KEY;
    if (c.pods_per_key > KVIDX_MAX_PODS_PER_KEY) return fail(KVIDX_EINVAL, "pods_per_key %u > %d", c.pods_per_key, KVIDX_MAX_PODS_PER_KEY);
    if (c.capacity == 0) c.capacity = 1ull << 20;
    if (c.max_pods == 0) c.max_pods = 256;
    if (c.max_pods > KVIDX_MAX_PODS) return fail(KVIDX_ERANGE, "max_pods %u > %u", c.max_pods, KVIDX_MAX_PODS);
    if (c.n_tier_weights > KVIDX_MAX_TIERS) return fail(KVIDX_ERANGE, "n_tier_weights > 16");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(KVIDX_ECUDA, "no CUDA device: libkvidx has no CPU fallback");
    }
    if (c.device < 0 || c.device >= ndev) return fail(KVIDX_EINVAL, "device %d out of range (%d devices)", c.device, ndev);
    kvidx* x = new kvidx();
    const int rc = create_impl(c, x);
    if (rc) {                       // whatever was created so far (streams, events, tables) goes away with the handle
        const std::string keep = g_err;
        kvidx_destroy(x);
        cudaGetLastError();
        g_err = keep;
        return rc;
    }
    *out = x;
    return 0;
}

void kvidx_destroy(kvidx_t* x) {
    if (!x) return;
    cudaSetDevice(x->device);
    cudaDeviceSynchronize();
    if (x->queue) { delete_submit_queue(x->queue); x->queue = nullptr; }
    for (int i = 0; i < 2; ++i) {
        x->d_tok[i].release(); x->d_off[i].release(); x->d_model[i].release(); x->d_filter[i].release();
        x->d_out[i].release(); x->d_aux[i].release(); x->h_stage[i].release(); x->h_out[i].release();
        if (x->ev_h2d[i]) cudaEventDestroy(x->ev_h2d[i]);
        if (x->ev_done[i]) cudaEventDestroy(x->ev_done[i]);
        if (x->ev_k[i]) cudaEventDestroy(x->ev_k[i]);
    }
    x->r_act0.release(); x->r_act1.release(); x->r_cnt.release(); x->r_hstate.release(); x->r_keys.release(); x->r_pst.release(); x->r_nbr.release(); x->r_fp.release(); x->r_sort.release(); x->r_role.release(); x->r_hl.release(); x->r_map.release(); x->r_src.release(); x->r_fate.release(); x->r_anch.release(); x->r_snap.release(); x->r_rec.release();
    x->d_misc.release(); x->d_rmisc.release(); x->d_ev.release(); x->d_hash.release(); x->d_evtok.release(); x->d_qoff.release(); x->h_misc.release();
    x->d_wkeys.release(); x->d_wpred.release(); x->d_wready.release(); x->d_wmap.release();
    if (x->tv.req) cudaFree(x->tv.req);
    if (x->tv.eng) cudaFree(x->tv.eng);
    if (x->tv.req_stamp) cudaFree(x->tv.req_stamp);
    if (x->d_cnt) cudaFree(x->d_cnt);
    if (x->h_cnt) cudaFreeHost(x->h_cnt);
    if (x->d_cnt_all) cudaFree(x->d_cnt_all);
    if (x->h_cnt_all) cudaFreeHost(x->h_cnt_all);
    if (x->own_stream) cudaStreamDestroy(x->own_stream);
    if (x->own_wstream) cudaStreamDestroy(x->own_wstream);
    if (x->copy_stream) cudaStreamDestroy(x->copy_stream);
    if (x->d2h_stream) cudaStreamDestroy(x->d2h_stream);
    for (int q = 0; q < kMaxParts - 1; ++q) {
        if (x->aux_stream[q]) cudaStreamDestroy(x->aux_stream[q]);
        if (x->ev_join[q]) cudaEventDestroy(x->ev_join[q]);
    }
    if (x->ev_fork) cudaEventDestroy(x->ev_fork);
    if (x->g_stream) cudaStreamDestroy(x->g_stream);
    for (int q = 0; q < kMaxParts; ++q) { if (x->hp_stream[q]) cudaStreamDestroy(x->hp_stream[q]); if (x->lp_stream[q]) cudaStreamDestroy(x->lp_stream[q]); }
    for (int q = 0; q < kMaxParts; ++q) { if (x->ev_g[q]) cudaEventDestroy(x->ev_g[q]); if (x->ev_r[q]) cudaEventDestroy(x->ev_r[q]); }
    if (x->ev_write) cudaEventDestroy(x->ev_write);
    for (int a_ = 0; a_ < 2; ++a_) for (int b_ = 0; b_ < 2; ++b_) {
        if (x->ev_spec_h[a_][b_]) cudaEventDestroy(x->ev_spec_h[a_][b_]);
        if (x->ev_spec_p[a_][b_]) cudaEventDestroy(x->ev_spec_p[a_][b_]);
    }
    cudaGetLastError();
    delete x;
}

int kvidx_set_tier_weight(kvidx_t* x, uint32_t tier, double w) {
    if (!x) return fail(KVIDX_EINVAL, "NULL handle");
    if (tier >= KVIDX_MAX_TIERS) return fail(KVIDX_ERANGE, "tier %u out of range", tier);
    std::unique_lock<std::shared_mutex> tl(x->tables);
    x->tv.weight[tier] = w;
    return 0;
}

int kvidx_set_stream(kvidx_t* x, void* s) {
    if (!x) return fail(KVIDX_EINVAL, "NULL handle");
    ReadGuard g(x);
    x->stream = s ? static_cast<cudaStream_t>(s) : x->own_stream;
    return 0;
}
int kvidx_synchronize(kvidx_t* x) {
    if (!x) return fail(KVIDX_EINVAL, "NULL handle");
    CK(cudaSetDevice(x->device));
    CK(cudaStreamSynchronize(x->stream));
    CK(cudaStreamSynchronize(x->own_wstream));
    return 0;
}

// ---- read path ------------------------------------------------------------------------------

int kvidx_hash_keys(kvidx_t* x, const uint32_t* tok, const int64_t* tok_off, int64_t n, const uint64_t* parent,
                    const uint8_t* parent_valid, uint64_t* keys_out, int64_t* key_off_out) {
    if (!x || !tok_off || !key_off_out || n < 0) return fail(KVIDX_EINVAL, "bad arguments");
    int rc = check_csr(tok_off, n);
    if (rc) return rc;
    const uint32_t B = x->tv.block_size;
    int64_t nk = 0;
    for (int64_t i = 0; i < n; ++i) { key_off_out[i] = nk; nk += (tok_off[i + 1] - tok_off[i]) / B; }
    key_off_out[n] = nk;
    if (n == 0 || nk == 0) return 0;
    if (!keys_out || !tok) return fail(KVIDX_EINVAL, "NULL buffer");
    ReadGuard g(x);
    CK(cudaSetDevice(x->device));
    const int64_t tb = tok_off[0], nt = tok_off[n] - tb;
    CK(x->d_tok[0].need((size_t)nt * 4 + 64));
    CK(x->d_off[0].need((size_t)(n + 1) * 8));
    CK(x->d_aux[0].need((size_t)(n + 1) * 8 + (size_t)n * 9));
    CK(x->d_out[0].need((size_t)nk * 8));
    CK(cudaMemcpyAsync(x->d_tok[0].p, tok + tb, (size_t)nt * 4, cudaMemcpyHostToDevice, x->stream));
    CK(cudaMemcpyAsync(x->d_off[0].p, tok_off, (size_t)(n + 1) * 8, cudaMemcpyHostToDevice, x->stream));
    uint8_t* aux = x->d_aux[0].as<uint8_t>();
    int64_t* d_koff = reinterpret_cast<int64_t*>(aux);
    uint64_t* d_parent = nullptr; uint8_t* d_pv = nullptr;
    CK(cudaMemcpyAsync(d_koff, key_off_out, (size_t)(n + 1) * 8, cudaMemcpyHostToDevice, x->stream));
    if (parent) {
        d_parent = reinterpret_cast<uint64_t*>(aux + (size_t)(n + 1) * 8);
        CK(cudaMemcpyAsync(d_parent, parent, (size_t)n * 8, cudaMemcpyHostToDevice, x->stream));
        if (parent_valid) {
            d_pv = aux + (size_t)(n + 1) * 8 + (size_t)n * 8;
            CK(cudaMemcpyAsync(d_pv, parent_valid, (size_t)n, cudaMemcpyHostToDevice, x->stream));
        }
    }
    rc = launch_hash_keys(x, x->d_tok[0].as<uint32_t>(), x->d_off[0].as<int64_t>(), tb, n, d_parent, d_pv, d_koff, x->d_out[0].as<uint64_t>(), x->stream);
    if (rc) return rc;
    CK(cudaMemcpyAsync(keys_out, x->d_out[0].p, (size_t)nk * 8, cudaMemcpyDeviceToHost, x->stream));
    CK(cudaStreamSynchronize(x->stream));
    return 0;
}

int kvidx_hash_keys_dev(kvidx_t* x, const uint32_t* d_tok, const int64_t* d_tok_off, int64_t n, const uint64_t* d_parent,
                        const uint8_t* d_parent_valid, const int64_t* d_key_off, uint64_t* d_keys_out) {
    if (!x || n < 0) return fail(KVIDX_EINVAL, "bad arguments");
    if (n == 0) return 0;
    ReadGuard g(x);
    CK(cudaSetDevice(x->device));
    return launch_hash_keys(x, d_tok, d_tok_off, 0, n, d_parent, d_parent_valid, d_key_off, d_keys_out, x->stream);
}

int kvidx_lookup(kvidx_t* x, uint32_t model, const uint64_t* keys, int64_t n, const uint64_t* filter,
                 kvidx_podtier_t* podtier_out, uint8_t* cnt_out) {
    if (!x) return fail(KVIDX_EINVAL, "NULL handle");
    if (n <= 0 || !keys) return fail(KVIDX_EINVAL, "no requestKeys provided for lookup");   // in_memory.go:108-110
    if (!podtier_out || !cnt_out) return fail(KVIDX_EINVAL, "NULL output");
    if (model > 0xffffu) return fail(KVIDX_ERANGE, "model id %u > 65535", model);
    ReadGuard g(x);
    CK(cudaSetDevice(x->device));
    if (int rc0 = check_shards(x)) return rc0;
    const uint32_t FW = x->tv.filter_words;
    CK(x->d_aux[0].need((size_t)n * 8 + (size_t)FW * 8 + 16));
    CK(x->d_out[0].need((size_t)n * kMaxEnt * 2 + (size_t)n + 16));
    uint8_t* aux = x->d_aux[0].as<uint8_t>();
    uint64_t* d_keys = reinterpret_cast<uint64_t*>(aux);
    uint64_t* d_f = nullptr;
    int* d_cut = reinterpret_cast<int*>(aux + (size_t)n * 8 + (size_t)FW * 8);
    CK(cudaStreamWaitEvent(x->stream, x->ev_write, 0));
    CK(cudaMemcpyAsync(d_keys, keys, (size_t)n * 8, cudaMemcpyHostToDevice, x->stream));
    if (filter) {
        d_f = reinterpret_cast<uint64_t*>(aux + (size_t)n * 8);
        CK(cudaMemcpyAsync(d_f, filter, (size_t)FW * 8, cudaMemcpyHostToDevice, x->stream));
    }
    const int big = 0x7fffffff;
    CK(cudaMemcpyAsync(d_cut, &big, sizeof(int), cudaMemcpyHostToDevice, x->stream));
    uint16_t* d_pt = x->d_out[0].as<uint16_t>();
    uint8_t* d_cnt = x->d_out[0].as<uint8_t>() + (size_t)n * kMaxEnt * 2;
    const int T = 128;
    lookup_kernel_v1<<<(unsigned)((n + T - 1) / T), T, 0, x->stream>>>(x->tv, model, d_keys, n, d_f, d_pt, d_cnt, d_cut,
                                                                      reserve_stamps(x, (unsigned long long)n));
    x->launches += 1;
    CK(cudaGetLastError());
    int cut = big;
    CK(cudaMemcpyAsync(podtier_out, d_pt, (size_t)n * kMaxEnt * 2, cudaMemcpyDeviceToHost, x->stream));
    CK(cudaMemcpyAsync(cnt_out, d_cnt, (size_t)n, cudaMemcpyDeviceToHost, x->stream));
    CK(cudaMemcpyAsync(&cut, d_cut, sizeof(int), cudaMemcpyDeviceToHost, x->stream));
    CK(cudaStreamSynchronize(x->stream));
    if (cut != big) for (int64_t i = cut; i < n; ++i) cnt_out[i] = 0;   // in_memory.go:119-122 early return
    return 0;
}

int kvidx_score_batch(kvidx_t* x, const uint32_t* tok, const int64_t* tok_off, int64_t n, const uint32_t* model, uint32_t model0,
                      const uint64_t* filter, double* scores_out, uint8_t* has_keys_out) {
    if (!x || !scores_out) return fail(KVIDX_EINVAL, "bad arguments");
    return submit_score(x, tok, tok_off, n, model, model0, filter, scores_out, nullptr, nullptr, nullptr, has_keys_out);
}

int kvidx_score_batch_sparse(kvidx_t* x, const uint32_t* tok, const int64_t* tok_off, int64_t n, const uint32_t* model, uint32_t model0,
                             const uint64_t* filter, uint16_t* pods_out, double* scores_out, uint8_t* cnt_out, uint8_t* has_keys_out) {
    if (!x || !pods_out || !scores_out || !cnt_out) return fail(KVIDX_EINVAL, "bad arguments");
    return submit_score(x, tok, tok_off, n, model, model0, filter, nullptr, pods_out, scores_out, cnt_out, has_keys_out);
}

int kvidx_score_batch_dev(kvidx_t* x, const uint32_t* d_tok, const int64_t* d_tok_off, int64_t n, const uint32_t* d_model,
                          uint32_t model0, const uint64_t* d_filter, double* d_scores_out, uint8_t* d_has_keys_out) {
    if (!x || n < 0) return fail(KVIDX_EINVAL, "bad arguments");
    ReadGuard g(x);
    CK(cudaSetDevice(x->device));
    CK(cudaStreamWaitEvent(x->stream, x->ev_write, 0));
    ScoreOut so{};
    so.dense = d_scores_out; so.has_keys = d_has_keys_out;
    return launch_score(x, d_tok, d_tok_off, 0, n, d_model, model0, d_filter, so, x->stream);
}

int kvidx_score_batch_sparse_dev(kvidx_t* x, const uint32_t* d_tok, const int64_t* d_tok_off, int64_t n, const uint32_t* d_model,
                                 uint32_t model0, const uint64_t* d_filter, uint16_t* d_pods_out, double* d_scores_out,
                                 uint8_t* d_cnt_out, uint8_t* d_has_keys_out) {
    if (!x || n < 0 || !d_pods_out || !d_scores_out || !d_cnt_out) return fail(KVIDX_EINVAL, "bad arguments");
    ReadGuard g(x);
    CK(cudaSetDevice(x->device));
    CK(cudaStreamWaitEvent(x->stream, x->ev_write, 0));
    ScoreOut so{};
    so.sp_pods = d_pods_out; so.sp_scores = d_scores_out; so.sp_cnt = d_cnt_out; so.has_keys = d_has_keys_out;
    return launch_score(x, d_tok, d_tok_off, 0, n, d_model, model0, d_filter, so, x->stream);
}

// ---- routed (all-to-all) form of the sharded Score(): the device steps around the exchange (kernels_route.cuh) ----

int kvidx_key_owners_dev(kvidx_t* x, const uint64_t* d_keys, const uint32_t* d_model, uint32_t model0, int64_t n, uint8_t* d_owner_out) {
    if (!x || n < 0) return fail(KVIDX_EINVAL, "bad arguments");
    if (n == 0) return 0;
    ReadGuard g(x);
    CK(cudaSetDevice(x->device));
    key_owners_kernel<<<(unsigned)((n + 255) / 256), 256, 0, x->stream>>>(x->tv, d_keys, d_model, model0, n, d_owner_out);
    x->launches += 1;
    CK(cudaGetLastError());
    return 0;
}

int kvidx_probe_slots_dev(kvidx_t* x, const uint64_t* d_keys, const uint32_t* d_model, uint32_t model0, int64_t n, void* d_slots_out) {
    if (!x || n < 0) return fail(KVIDX_EINVAL, "bad arguments");
    if (n == 0) return 0;
    ReadGuard g(x);
    CK(cudaSetDevice(x->device));
    if (int rc0 = check_shards(x)) return rc0;
    CK(cudaStreamWaitEvent(x->stream, x->ev_write, 0));
    probe_slots_kernel<<<(unsigned)((n + 255) / 256), 256, 0, x->stream>>>(x->tv, d_keys, d_model, model0, n, static_cast<uint4*>(d_slots_out));
    x->launches += 1;
    CK(cudaGetLastError());
    return 0;
}

int kvidx_score_slots_dev(kvidx_t* x, const void* d_slots, const int64_t* d_key_off, int64_t n_prompts, const uint64_t* d_filter,
                          double* d_scores_out, uint8_t* d_has_keys_out) {
    if (!x || n_prompts < 0 || !d_scores_out) return fail(KVIDX_EINVAL, "bad arguments");
    if (n_prompts == 0) return 0;
    ReadGuard g(x);
    CK(cudaSetDevice(x->device));
    score_slots_kernel<<<(unsigned)((n_prompts + 127) / 128), 128, 0, x->stream>>>(x->tv, static_cast<const uint4*>(d_slots), d_key_off, n_prompts, d_filter,
                                                                                   d_scores_out, d_has_keys_out);
    x->launches += 1;
    CK(cudaGetLastError());
    return 0;
}

// ---- write path -----------------------------------------------------------------------------

namespace {
int check_podtiers(kvidx* x, const kvidx_podtier_t* pts, int32_t m) {
    // dense score rows and filter rows are max_pods wide: a pod id beyond that cannot be scored or filtered
    for (int32_t j = 0; j < m; ++j)
        if (KVIDX_PT_POD(pts[j]) >= x->tv.max_pods) return fail(KVIDX_ERANGE, "pod id %u >= max_pods %u", KVIDX_PT_POD(pts[j]), x->tv.max_pods);
    return 0;
}
}  // namespace

int kvidx_add(kvidx_t* x, uint32_t model, const uint64_t* engine, const uint64_t* request, int64_t n,
              const kvidx_podtier_t* pts, int32_t m) {
    if (!x) return fail(KVIDX_EINVAL, "NULL handle");
    if (n <= 0 || m <= 0 || !engine || !request || !pts) return fail(KVIDX_EINVAL, "no keys or entries provided for adding to index");
    if (model > 0xffffu) return fail(KVIDX_ERANGE, "model id %u > 65535", model);
    if (int rc0 = check_podtiers(x, pts, m)) return rc0;
    std::lock_guard<std::mutex> g(x->mu_w);
    CK(cudaSetDevice(x->device));
    if (int rc0 = check_shards(x)) return rc0;
    int rc = ensure_room(x, (uint64_t)n);
    if (rc) return rc;
    std::shared_lock<std::shared_mutex> tl(x->tables);
    cudaStream_t st = wstream(x);
    CK(x->d_misc.need((size_t)n * 17 + (size_t)m * 2 + 32));
    CK(x->h_misc.need((size_t)n + 16));
    // pairs are added in order (in_memory.go:159): an engine key that appears again later in the call keeps the later mapping
    uint8_t* skip = x->h_misc.as<uint8_t>();
    bool any_dup = false;
    {
        std::unordered_map<uint64_t, int64_t> last;
        last.reserve((size_t)n * 2);
        for (int64_t i = 0; i < n; ++i) { skip[i] = 0; auto it = last.find(engine[i]); if (it != last.end()) { skip[it->second] = 1; any_dup = true; it->second = i; } else last.emplace(engine[i], i); }
    }
    uint8_t* d = x->d_misc.as<uint8_t>();
    uint8_t* d_skip = d + (size_t)n * 16 + (((size_t)m * 2 + 15) & ~(size_t)15);
    CK(cudaMemcpyAsync(d, engine, (size_t)n * 8, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d + (size_t)n * 8, request, (size_t)n * 8, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d + (size_t)n * 16, pts, (size_t)m * 2, cudaMemcpyHostToDevice, st));
    if (any_dup) CK(cudaMemcpyAsync(d_skip, skip, (size_t)n, cudaMemcpyHostToDevice, st));
    const int T = 128;
    add_kernel<<<(unsigned)((n + T - 1) / T), T, 0, st>>>(x->tv, model, reinterpret_cast<uint64_t*>(d),
                                                         reinterpret_cast<uint64_t*>(d + (size_t)n * 8), n,
                                                         reinterpret_cast<uint16_t*>(d + (size_t)n * 16), m, any_dup ? d_skip : nullptr,
                                                         reserve_stamps(x, 2ull * (unsigned long long)n + 2));
    x->launches += 1;
    CK(cudaGetLastError());
    const unsigned long long nospc_before = x->h_cnt->nospc;
    rc = refresh_counters(x);
    if (rc) return rc;
    if (x->h_cnt->nospc != nospc_before) return fail(KVIDX_ENOSPC, "%llu inserts refused: the owning shard is full", x->h_cnt->nospc - nospc_before);
    return enforce_caps(x);
}

int kvidx_evict(kvidx_t* x, uint32_t model, uint64_t engine, const kvidx_podtier_t* pts, int32_t m) {
    if (!x) return fail(KVIDX_EINVAL, "NULL handle");
    if (m <= 0 || !pts) return fail(KVIDX_EINVAL, "no entries provided for eviction from index");
    if (model > 0xffffu) return fail(KVIDX_ERANGE, "model id %u > 65535", model);
    if (int rc0 = check_podtiers(x, pts, m)) return rc0;
    std::lock_guard<std::mutex> g(x->mu_w);
    CK(cudaSetDevice(x->device));
    if (int rc0 = check_shards(x)) return rc0;
    std::shared_lock<std::shared_mutex> tl(x->tables);
    cudaStream_t st = wstream(x);
    CK(x->d_misc.need((size_t)m * 2 + 16));
    CK(cudaMemcpyAsync(x->d_misc.p, pts, (size_t)m * 2, cudaMemcpyHostToDevice, st));
    evict_kernel<<<1, 32, 0, st>>>(x->tv, model, engine, x->d_misc.as<uint16_t>(), m, reserve_stamps(x, 4));
    x->launches += 1;
    CK(cudaGetLastError());
    return refresh_counters(x);
}

int kvidx_get_request_key(kvidx_t* x, uint32_t model, uint64_t engine, uint64_t* out) {
    if (!x || !out) return fail(KVIDX_EINVAL, "bad arguments");
    ReadGuard g(x);
    CK(cudaSetDevice(x->device));
    if (int rc0 = check_shards(x)) return rc0;
    CK(x->d_rmisc.need(32));
    uint64_t* d_out = x->d_rmisc.as<uint64_t>();
    int* d_found = reinterpret_cast<int*>(d_out + 1);
    CK(cudaStreamWaitEvent(x->stream, x->ev_write, 0));
    get_request_key_kernel<<<1, 32, 0, x->stream>>>(x->tv, model, engine, d_out, d_found, x->tv.req_stamp ? reserve_stamps(x, 1) : 0);
    x->launches += 1;
    CK(cudaGetLastError());
    struct { uint64_t r; int f; int pad; } h{};
    CK(cudaMemcpyAsync(&h, d_out, 16, cudaMemcpyDeviceToHost, x->stream));
    CK(cudaStreamSynchronize(x->stream));
    if (!h.f) return fail(KVIDX_ENOENT, "engine key not found: %u@%llu", model, (unsigned long long)engine);
    *out = h.r;
    return 0;
}

namespace {
// Both phases of a device-resident, pod-sorted event batch on the write stream (kernels_write.cuh).  Caller holds mu_w and
// `tables` shared.
int launch_apply_events(kvidx* x, const kvidx_event_t* d_ev_sorted, const int64_t* d_queue_off, int64_t n_queues, int64_t n_events,
                        const uint64_t* d_hashes, int64_t n_hashes, const uint32_t* d_tokens, cudaStream_t st) {
    const uint64_t* d_keys = nullptr; const uint64_t* d_pred = nullptr;
    const bool phase1 = x->write_phase1 == 1 || (x->write_phase1 < 0 && n_hashes >= x->write_phase1_min);
    if (phase1 && n_hashes > 0) {
        const uint64_t map_slots = pow2ceil((uint64_t)std::max<int64_t>(2 * n_events, 64));
        CK(x->d_wkeys.need((size_t)n_hashes * 8));
        CK(x->d_wpred.need((size_t)n_events * 8));
        // one zeroed block: work counter (16 B) | ready[n] | next[n] | best[n] (8-byte aligned)
        const size_t n4 = ((size_t)n_events + 3) & ~(size_t)3;
        const size_t zbytes = 16 + n4 * 4 + n4 * 4 + (size_t)n_events * 8;
        CK(x->d_wready.need(zbytes));
        CK(x->d_wmap.need((size_t)map_slots * sizeof(WantEnt)));
        uint8_t* zb = x->d_wready.as<uint8_t>();
        unsigned long long* d_next = reinterpret_cast<unsigned long long*>(zb);
        unsigned int* d_ready = reinterpret_cast<unsigned int*>(zb + 16);
        unsigned int* d_chain = reinterpret_cast<unsigned int*>(zb + 16 + n4 * 4);
        unsigned long long* d_best = reinterpret_cast<unsigned long long*>(zb + 16 + n4 * 8);
        CK(cudaMemsetAsync(zb, 0, zbytes, st));
        CK(cudaMemsetAsync(x->d_wmap.p, 0, (size_t)map_slots * sizeof(WantEnt), st));
        const int T = 256;
        want_parents_kernel<<<(unsigned)((n_events + T - 1) / T), T, 0, st>>>(d_ev_sorted, n_events, x->tv.block_size, x->d_wmap.as<WantEnt>(), (uint32_t)(map_slots - 1), d_chain);
        offer_blocks_kernel<<<(unsigned)((n_events * 32 + T - 1) / T), T, 0, st>>>(d_ev_sorted, n_events, x->tv.block_size, d_hashes, x->d_wmap.as<WantEnt>(), (uint32_t)(map_slots - 1), d_chain, d_best);
        const int64_t ctas = std::min<int64_t>((n_events + kHashEvThreads - 1) / kHashEvThreads, (int64_t)x->sm_count * 8);
        hash_events_kernel<<<(unsigned)ctas, kHashEvThreads, 0, st>>>(x->tv, d_ev_sorted, n_events, d_hashes, d_tokens, d_best,
                                                                     x->d_wkeys.as<uint64_t>(), x->d_wpred.as<uint64_t>(), d_ready, d_next);
        x->launches += 3;
        CK(cudaGetLastError());
        d_keys = x->d_wkeys.as<uint64_t>(); d_pred = x->d_wpred.as<uint64_t>();
    }
    const int T = 128;   // 4 queues per CTA
    apply_events_kernel<<<(unsigned)((n_queues * 32 + T - 1) / T), T, 0, st>>>(x->tv, d_ev_sorted, d_queue_off, n_queues, d_hashes, d_tokens, d_keys, d_pred,
                                                                               x->tv.req_stamp ? reserve_stamps(x, (1ull << 20) * (unsigned long long)std::max<int64_t>(n_events, 1)) : 0);
    x->launches += 1;
    CK(cudaGetLastError());
    return 0;
}
}  // namespace

int kvidx_apply_events_dev(kvidx_t* x, const kvidx_event_t* d_ev_sorted, const int64_t* d_queue_off, int64_t n_queues, int64_t n_events,
                           const uint64_t* d_hashes, int64_t n_hashes, const uint32_t* d_tokens, int64_t* d_n_dropped) {
    if (!x || n_queues < 0 || n_events < 0 || n_hashes < 0) return fail(KVIDX_EINVAL, "bad arguments");
    if (n_queues == 0 || n_events == 0) return 0;
    std::lock_guard<std::mutex> g(x->mu_w);
    CK(cudaSetDevice(x->device));
    if (int rc0 = check_shards(x)) return rc0;
    int rc = ensure_room(x, (uint64_t)n_hashes);       // every hash of the batch may be a new key
    if (rc) return rc;
    std::shared_lock<std::shared_mutex> tl(x->tables);
    cudaStream_t st = wstream(x);
    rc = launch_apply_events(x, d_ev_sorted, d_queue_off, n_queues, n_events, d_hashes, n_hashes, d_tokens, st);
    if (rc) return rc;
    // events dropped so far on this handle (cumulative, like the counter behind kvidx_apply_events' n_dropped_out)
    if (d_n_dropped) CK(cudaMemcpyAsync(d_n_dropped, &x->d_cnt->dropped_events, sizeof(int64_t), cudaMemcpyDeviceToDevice, st));
    CK(cudaEventRecord(x->ev_write, st));              // reads issued after this call are ordered after the batch
    return 0;
}

int kvidx_apply_events(kvidx_t* x, const kvidx_event_t* ev, int64_t n, const uint64_t* hashes, int64_t n_hashes,
                       const uint32_t* tokens, int64_t n_tokens, int64_t* n_dropped_out) {
    if (!x || n < 0) return fail(KVIDX_EINVAL, "bad arguments");
    if (n_dropped_out) *n_dropped_out = 0;
    if (n == 0) return 0;
    if (!ev) return fail(KVIDX_EINVAL, "NULL events");
    // validate + stable counting sort by pod -> per-pod FIFO queues (pool.go:129-144)
    std::vector<int64_t> qcount(KVIDX_MAX_PODS + 1, 0);
    uint64_t new_keys = 0;
    for (int64_t i = 0; i < n; ++i) {
        const kvidx_event_t& e = ev[i];
        if (e.op > KVIDX_EV_BLOCK_REMOVED) return fail(KVIDX_EINVAL, "event %lld: unknown op %u", (long long)i, e.op);
        if (e.model > 0xffffu) return fail(KVIDX_ERANGE, "event %lld: model id %u > 65535", (long long)i, e.model);
        if (KVIDX_PT_POD(e.podtier) >= x->tv.max_pods) return fail(KVIDX_ERANGE, "event %lld: pod id %u >= max_pods %u", (long long)i, KVIDX_PT_POD(e.podtier), x->tv.max_pods);
        if (e.hash_off + e.n_hashes > (uint64_t)n_hashes) return fail(KVIDX_EINVAL, "event %lld: hashes out of range", (long long)i);
        if (e.op == KVIDX_EV_BLOCK_STORED) {
            if (e.tok_off + e.n_tokens > (uint64_t)n_tokens) return fail(KVIDX_EINVAL, "event %lld: tokens out of range", (long long)i);
            new_keys += e.n_hashes;
        }
        qcount[KVIDX_PT_POD(e.podtier) + 1]++;
    }
    std::lock_guard<std::mutex> g(x->mu_w);
    CK(cudaSetDevice(x->device));
    if (int rc0 = check_shards(x)) return rc0;
    int rc = ensure_room(x, new_keys);
    if (rc) return rc;
    std::shared_lock<std::shared_mutex> tl(x->tables);
    cudaStream_t st = wstream(x);
    // compact non-empty queues
    std::vector<int64_t> start(KVIDX_MAX_PODS + 1, 0);
    for (uint32_t p = 0; p < KVIDX_MAX_PODS; ++p) start[p + 1] = start[p] + qcount[p + 1];
    CK(x->h_misc.need((size_t)n * sizeof(kvidx_event_t) + (size_t)(KVIDX_MAX_PODS + 2) * 8));
    kvidx_event_t* sorted = x->h_misc.as<kvidx_event_t>();
    int64_t* qoff = reinterpret_cast<int64_t*>(x->h_misc.as<uint8_t>() + (size_t)n * sizeof(kvidx_event_t));
    {
        std::vector<int64_t> cur(start.begin(), start.end() - 1);
        for (int64_t i = 0; i < n; ++i) sorted[cur[KVIDX_PT_POD(ev[i].podtier)]++] = ev[i];
    }
    int64_t nq = 0;
    for (uint32_t p = 0; p < KVIDX_MAX_PODS; ++p) if (qcount[p + 1]) { qoff[nq++] = start[p]; }
    qoff[nq] = n;
    CK(x->d_ev.need((size_t)n * sizeof(kvidx_event_t)));
    CK(x->d_qoff.need((size_t)(nq + 1) * 8));
    CK(x->d_hash.need((size_t)std::max<int64_t>(n_hashes, 1) * 8));
    CK(x->d_evtok.need((size_t)std::max<int64_t>(n_tokens, 1) * 4 + 16));
    CK(cudaMemcpyAsync(x->d_ev.p, sorted, (size_t)n * sizeof(kvidx_event_t), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(x->d_qoff.p, qoff, (size_t)(nq + 1) * 8, cudaMemcpyHostToDevice, st));
    if (n_hashes > 0) CK(cudaMemcpyAsync(x->d_hash.p, hashes, (size_t)n_hashes * 8, cudaMemcpyHostToDevice, st));
    if (n_tokens > 0) CK(cudaMemcpyAsync(x->d_evtok.p, tokens, (size_t)n_tokens * 4, cudaMemcpyHostToDevice, st));
    const unsigned long long dropped_before = x->h_cnt->dropped_events, nospc_before = x->h_cnt->nospc;
    x->last_batch_events = n;
    rc = launch_apply_events(x, x->d_ev.as<kvidx_event_t>(), x->d_qoff.as<int64_t>(), nq, n, x->d_hash.as<uint64_t>(), n_hashes,
                             x->d_evtok.as<uint32_t>(), st);
    if (rc) return rc;
    rc = refresh_counters(x);
    if (rc) return rc;
    if (n_dropped_out) *n_dropped_out = (int64_t)(x->h_cnt->dropped_events - dropped_before);
    if (x->h_cnt->nospc != nospc_before) return fail(KVIDX_ENOSPC, "%llu inserts refused: the owning shard is full", x->h_cnt->nospc - nospc_before);
    return enforce_caps(x);
}

int kvidx_shard_export(kvidx_t* x, void* out) {
    if (!x || !out) return fail(KVIDX_EINVAL, "bad arguments");
    std::lock_guard<std::mutex> g(x->mu_w);
    CK(cudaSetDevice(x->device));
    cudaIpcMemHandle_t h[3];
    CK(cudaIpcGetMemHandle(&h[0], x->tv.req));
    CK(cudaIpcGetMemHandle(&h[1], x->tv.eng));
    CK(cudaIpcGetMemHandle(&h[2], x->d_cnt));
    static_assert(sizeof(h) == KVIDX_SHARD_HANDLE_BYTES, "handle blob size");
    memcpy(out, h, sizeof h);
    return 0;
}

int kvidx_shard_import(kvidx_t* x, uint32_t rank, const void* handle) {
    if (!x || !handle) return fail(KVIDX_EINVAL, "bad arguments");
    if (rank >= (1u << x->tv.shard_bits)) return fail(KVIDX_ERANGE, "rank %u outside shard_count", rank);
    if (rank == x->tv.shard_rank) return 0;
    std::lock_guard<std::mutex> g(x->mu_w);
    std::unique_lock<std::shared_mutex> tl(x->tables);
    CK(cudaSetDevice(x->device));
    cudaIpcMemHandle_t h[3];
    memcpy(h, handle, sizeof h);
    void* p[3] = {nullptr, nullptr, nullptr};
    for (int i = 0; i < 3; ++i) CK(cudaIpcOpenMemHandle(&p[i], h[i], cudaIpcMemLazyEnablePeerAccess));
    x->tv.req_peer[rank] = static_cast<ReqSlot*>(p[0]); x->tv.eng_peer[rank] = static_cast<EngSlot*>(p[1]);
    x->tv.cnt_peer[rank] = static_cast<Counters*>(p[2]);
    return 0;
}

int kvidx_shard_attach(kvidx_t* x, uint32_t rank, kvidx_t* other) {
    if (!x || !other) return fail(KVIDX_EINVAL, "bad arguments");
    if (rank >= (1u << x->tv.shard_bits)) return fail(KVIDX_ERANGE, "rank %u outside shard_count", rank);
    if (rank == x->tv.shard_rank) return 0;
    if (other->tv.req_mask != x->tv.req_mask || other->tv.shard_bits != x->tv.shard_bits || other->tv.shard_rank != rank)
        return fail(KVIDX_EINVAL, "shard geometry mismatch");
    std::lock_guard<std::mutex> g(x->mu_w);
    std::unique_lock<std::shared_mutex> tl(x->tables);
    CK(cudaSetDevice(x->device));
    if (other->device != x->device) {
        int can = 0;
        CK(cudaDeviceCanAccessPeer(&can, x->device, other->device));
        if (!can) return fail(KVIDX_ECUDA, "device %d cannot access device %d", x->device, other->device);
        cudaError_t e = cudaDeviceEnablePeerAccess(other->device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return fail(KVIDX_ECUDA, "cudaDeviceEnablePeerAccess: %s", cudaGetErrorString(e));
        cudaGetLastError();
    }
    x->tv.req_peer[rank] = other->tv.req; x->tv.eng_peer[rank] = other->tv.eng; x->tv.cnt_peer[rank] = other->d_cnt;
    return 0;
}

int kvidx_shard_compact(kvidx_t* x) {
    if (!x) return fail(KVIDX_EINVAL, "NULL handle");
    std::lock_guard<std::mutex> g(x->mu_w);
    std::unique_lock<std::shared_mutex> tl(x->tables);
    CK(cudaSetDevice(x->device));
    return rebuild(x, x->tv.shard_bits != 0);
}

int kvidx_get_stats(kvidx_t* x, kvidx_stats_t* out) {
    if (!x || !out) return fail(KVIDX_EINVAL, "bad arguments");
    std::lock_guard<std::mutex> g(x->mu_w);
    CK(cudaSetDevice(x->device));
    int rc = refresh_counters(x);
    if (rc) return rc;
    out->request_keys = x->h_cnt->req_full; out->engine_keys = x->h_cnt->eng_full;
    out->request_tombs = x->h_cnt->req_tomb; out->engine_tombs = x->h_cnt->eng_tomb;
    out->request_slots = x->tv.req_mask + 1; out->engine_slots = x->tv.eng_mask + 1;
    out->rebuilds = x->rebuilds; out->kernel_launches = x->launches.load();
    out->rehashed_events = x->h_cnt->rehashed; out->coalesced_calls = x->queue ? submit_queue_coalesced(x->queue) : 0;
    return 0;
}

}  // extern "C"
