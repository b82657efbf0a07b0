/*
 * kvidx.h -- C ABI of libkvidx: the B200-resident KV-block locality index.
 *
 * This is the drop-in boundary for ONE hot path of llm-d/llm-d-kv-cache-manager:
 * kvcache.Indexer.GetPodScores() steps 2-4 (token blocks -> chain-hashed request
 * keys -> index probe -> longest-consecutive-prefix per-pod score) and the
 * kvevents.Pool -> kvblock.Index.Add/Evict write path that feeds the same table.
 * Every entry point names the reference interface (path:line under the
 * reference root) it replaces.  A Go maintainer binds these with cgo
 * (INTEGRATION.md shows the stub); the tests bind them with ctypes.
 *
 * Conventions
 *   - plain C types only; caller owns every buffer; nothing is retained after
 *     return (cgo rule: no Go pointer survives the call);
 *   - return 0 on success or a negative KVIDX_E* code; kvidx_last_error() gives
 *     the message for the calling thread's last failure on that handle;
 *   - strings are interned on the host side of the boundary: model -> uint32,
 *     pod -> 12-bit id, device tier -> 4-bit id.  (pod,tier) travels packed in a
 *     kvidx_podtier_t.  The host mirror (kvidx_host.h) owns the string maps;
 *   - all entry points may be called concurrently from many OS threads (cgo
 *     pins one per call), as kvblock.Index requires (kvblock/index.go:118).
 *     The READ path (hash / lookup / score / get_request_key) and the WRITE path
 *     (add / evict / apply_events) run on separate CUDA streams and overlap on the
 *     device: a reader sees every slot either before or after an individual
 *     Add / Evict of that key (table.cuh), which is the guarantee the reference's
 *     per-key mutex gives.  A write call has taken effect when it returns, so a
 *     thread reads its own writes (index_test.go:214-278).  Concurrent host-buffer
 *     kvidx_score_batch* callers are coalesced into one launch;
 *   - there is NO CPU fallback: without a CUDA device kvidx_create fails with
 *     KVIDX_ECUDA.
 */
#ifndef KVIDX_H
#define KVIDX_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KVIDX_ABI_VERSION 2   /* 2: read and write path run concurrently; kvidx_apply_events_dev takes the batch sizes;
                                  kvidx_score_batch_sparse_dev, kvidx_shard_compact; two more stats fields */

/* error codes (negative errno style) */
#define KVIDX_OK        0
#define KVIDX_ENOENT   (-2)   /* engine key not found (in_memory.go:266-268)                 */
#define KVIDX_ECUDA    (-5)   /* CUDA runtime / device failure, or no device                  */
#define KVIDX_ENOMEM   (-12)
#define KVIDX_EINVAL   (-22)  /* empty args, length mismatch ... (in_memory.go:108-110,150-155,213-215) */
#define KVIDX_ENOSPC   (-28)  /* table full                                                    */
#define KVIDX_ERANGE   (-34)  /* id outside the configured pod / tier / model range            */

/* packed (pod, tier) entry == kvblock.PodEntry (index.go:149-154) after interning */
#define KVIDX_TIER_BITS        4
#define KVIDX_POD_BITS         12
#define KVIDX_MAX_TIERS        (1u << KVIDX_TIER_BITS)   /* 16   */
#define KVIDX_MAX_PODS         (1u << KVIDX_POD_BITS)    /* 4096 */
#define KVIDX_MAX_PODS_PER_KEY 10                        /* in_memory.go:34 defaultPodsPerKey */
typedef uint16_t kvidx_podtier_t;
#define KVIDX_PODTIER(pod, tier) ((kvidx_podtier_t)((((uint32_t)(pod)) << KVIDX_TIER_BITS) | ((uint32_t)(tier) & (KVIDX_MAX_TIERS - 1))))
#define KVIDX_PT_POD(pt)  ((uint32_t)(pt) >> KVIDX_TIER_BITS)
#define KVIDX_PT_TIER(pt) ((uint32_t)(pt) & (KVIDX_MAX_TIERS - 1))

/* value written into a dense score row for a pod that is NOT in the result map.
 * The reference's scores are sums of max(0, weight) (kvblock_scorer.go:90,100), so
 * every real score is >= 0 and -1.0 is unambiguous. */
#define KVIDX_SCORE_ABSENT (-1.0)

typedef struct kvidx kvidx_t;   /* opaque */

/* Configuration == TokenProcessorConfig (token_processor.go:34-42)
 *                + InMemoryIndexConfig  (in_memory.go:38-43)
 *                + KVCacheBackendConfig weights (backend.go:19-31). */
typedef struct kvidx_config {
    uint32_t struct_size;      /* sizeof(kvidx_config_t), for ABI growth                        */
    int32_t  device;           /* CUDA device ordinal                                           */
    uint32_t block_size;       /* tokens per block; 0 -> 16 (token_processor.go:31)             */
    uint32_t pods_per_key;     /* PodCacheSize; 0 -> 10; must be 1..10                          */
    uint64_t init_hash;        /* FNV-64a(HashSeed) (token_processor.go:81-90); use kvidx_fnv64a */
    uint64_t capacity;         /* InMemoryIndexConfig.Size: max resident keys; 0 -> 1<<20       */
    uint64_t table_slots;      /* open-addressing slots (power of two); 0 -> 4*capacity rounded up
                                  (load factor <= 0.25: a probe almost always ends in its home pair) */
    uint32_t max_pods;         /* dense score-row width P (1..4096); 0 -> 256                   */
    uint32_t n_tier_weights;   /* entries of tier_weight[] that are configured                  */
    double   tier_weight[KVIDX_MAX_TIERS]; /* weight of tier id i; ids >= n_tier_weights score 1.0
                                              (unknown tier => 1.0, kvblock_scorer.go:93-98)    */
    uint32_t lru_exact;        /* 1: track key recency (a stamp per slot, written by Lookup / Score / Add / Evict /
                                  GetRequestKey exactly where golang-lru refreshes) and evict the least recently used
                                  request / engine keys when a write call leaves more than `capacity` of them
                                  (in_memory.go:59,64,118,163,170).  Exact at call granularity.  Score() then probes
                                  every key of a prompt (the reference's Lookup has no early exit), so it is the slow,
                                  semantics-first mode.  0 (default): recency is not tracked and `capacity` only sizes
                                  the table -- right whenever the fleet's block count stays below Size (default 1e8). */
    uint32_t shard_rank;       /* hash-range sharding over the GPUs of one NVSwitch domain: this handle */
    uint32_t shard_count;      /* owns shard `shard_rank` of `shard_count` (power of two <= 8; 0 or 1 = unsharded).   */
                               /* `capacity` stays the TOTAL key budget; each shard holds capacity/shard_count.        */
    uint32_t reserved[5];
} kvidx_config_t;

void        kvidx_config_default(kvidx_config_t* cfg);
int         kvidx_create(const kvidx_config_t* cfg, kvidx_t** out);   /* NewInMemoryIndex + NewChunkedTokenDatabase + NewKVBlockScorer */
void        kvidx_destroy(kvidx_t* idx);
const char* kvidx_last_error(kvidx_t* idx);
int         kvidx_abi_version(void);

/* host-only helpers */
uint64_t kvidx_fnv64a(const void* data, size_t n);                 /* getInitHash, token_processor.go:81-90 */
uint32_t kvidx_fnv32a(const void* data, size_t n);
uint32_t kvidx_queue_index(const char* pod, size_t n, uint32_t concurrency); /* Pool.AddTask, kvevents/pool.go:132-144 */
int      kvidx_set_tier_weight(kvidx_t* idx, uint32_t tier, double weight);

/* pinned staging a caller may use for zero-copy submission (optional) */
void* kvidx_host_alloc(size_t bytes);
void  kvidx_host_free(void* p);

/* ---- read path ---------------------------------------------------------------- */

/* TokenProcessor.TokensToKVBlockKeys (token_processor.go:141-162) for a batch.
 * tok/tok_off: CSR, prompt i owns tok[tok_off[i] .. tok_off[i+1]).
 * parent: NULL -> every chain starts at init_hash; else parent[i] is the parent
 *         request-key hash of prompt i (parent_valid[i]==0 -> init_hash; parent_valid
 *         NULL -> all valid).
 * keys_out: sum_i floor(len_i/B) hashes, prompt-major; key_off_out[n_prompts+1]. */
int kvidx_hash_keys(kvidx_t* idx, const uint32_t* tok, const int64_t* tok_off, int64_t n_prompts,
                    const uint64_t* parent, const uint8_t* parent_valid,
                    uint64_t* keys_out, int64_t* key_off_out);

/* Index.Lookup (index.go:125, in_memory.go:105-146).  filter: NULL = all pods, else a
 * bitmask of ceil(max_pods/64) words over pod ids.  For key i, cnt_out[i] entries are
 * written to podtier_out[i*10 ..] oldest->newest (lru Keys() order); cnt_out[i]==0 means
 * "key absent from the result map" (missing, filtered to nothing, or after the
 * present-but-empty cut).  n==0 -> KVIDX_EINVAL (in_memory.go:108-110). */
int kvidx_lookup(kvidx_t* idx, uint32_t model, const uint64_t* keys, int64_t n,
                 const uint64_t* filter, kvidx_podtier_t* podtier_out, uint8_t* cnt_out);

/* Indexer.GetPodScores steps 2-4 fused (indexer.go:141-163) for a batch of prompts.
 * Callers that arrive while another call's batch is on the device are queued and served together by one launch
 * (each gets exactly the rows of its own prompts; KVIDX_SUBMIT_QUEUE=0 turns the coalescing off).
 * model: per-prompt model ids, or NULL to use model0 for all.
 * filter: NULL, or n_prompts * ceil(max_pods/64) words (all-zero row == empty set == all pods,
 *         indexer.go:151 / in_memory.go:126).
 * scores_out: n_prompts * max_pods doubles; KVIDX_SCORE_ABSENT where the pod is not in the
 *         reference's result map.  A prompt with no full block yields an all-absent row and
 *         has_keys_out[i]=0 (reference returns (nil,nil), indexer.go:142-146); has_keys_out
 *         may be NULL. */
int kvidx_score_batch(kvidx_t* idx, const uint32_t* tok, const int64_t* tok_off, int64_t n_prompts,
                      const uint32_t* model, uint32_t model0, const uint64_t* filter,
                      double* scores_out, uint8_t* has_keys_out);

/* Same, sparse result: for prompt i cnt_out[i] (<=10) pairs pods_out[i*10+j], scores_out[i*10+j]. */
int kvidx_score_batch_sparse(kvidx_t* idx, const uint32_t* tok, const int64_t* tok_off, int64_t n_prompts,
                             const uint32_t* model, uint32_t model0, const uint64_t* filter,
                             uint16_t* pods_out, double* scores_out, uint8_t* cnt_out, uint8_t* has_keys_out);

/* ---- write path --------------------------------------------------------------- */

/* Index.Add (index.go:127, in_memory.go:149-209): n (engine,request) key pairs of one model,
 * m entries added to every key in order.  n==0 or m==0 -> KVIDX_EINVAL. */
int kvidx_add(kvidx_t* idx, uint32_t model, const uint64_t* engine, const uint64_t* request, int64_t n,
              const kvidx_podtier_t* podtier, int32_t m);

/* Index.Evict (index.go:129, in_memory.go:212-260).  Unknown engine key is a silent no-op. */
int kvidx_evict(kvidx_t* idx, uint32_t model, uint64_t engine, const kvidx_podtier_t* podtier, int32_t m);

/* Index.GetRequestKey (index.go:131, in_memory.go:264-270); miss -> KVIDX_ENOENT. */
int kvidx_get_request_key(kvidx_t* idx, uint32_t model, uint64_t engine, uint64_t* request_out);

/* One decoded KV event == one element of Pool.digestEvents' loop (kvevents/pool.go:246-338)
 * after msgpack decoding, getHashAsUint64 (pool.go:343-367) and string interning. */
#define KVIDX_EV_BLOCK_STORED  0
#define KVIDX_EV_BLOCK_REMOVED 1
typedef struct kvidx_event {
    uint8_t  op;            /* KVIDX_EV_*                                                       */
    uint8_t  has_parent;    /* BlockStored: ParentBlockHash != nil                              */
    kvidx_podtier_t podtier;/* {pod of the message, lower(Medium) or "gpu"} (pool.go:258-265)   */
    uint32_t model;
    uint64_t parent_hash;   /* engine hash of the parent block                                  */
    uint64_t hash_off;      /* first engine hash of this event in `hashes`                      */
    uint64_t tok_off;       /* first token id of this event in `tokens` (BlockStored)           */
    uint32_t n_hashes;      /* engine hashes that survived getHashAsUint64                      */
    uint32_t n_tokens;
} kvidx_event_t;

/* Apply a batch of events.  Events of one pod are applied in array order (pool.go:129-144);
 * events of different pods are applied in an unspecified interleaving, as in the reference.
 * BlockStored whose request-key count differs from n_hashes is dropped (in_memory.go:153-155)
 * and counted in *n_dropped_out (may be NULL). */
int kvidx_apply_events(kvidx_t* idx, const kvidx_event_t* ev, int64_t n_events,
                       const uint64_t* hashes, int64_t n_hashes,
                       const uint32_t* tokens, int64_t n_tokens, int64_t* n_dropped_out);

/* ---- device-resident variants (inputs already in HBM; used by bench `value`, by the
 *      multi-GPU router and by callers that keep token buffers on the device) ------- */

/* Use an external CUDA stream (cudaStream_t as void*) for subsequent READ-path calls; NULL restores the
 * handle's own stream.  The write path always uses the handle's own write stream. */
int kvidx_set_stream(kvidx_t* idx, void* cuda_stream);
int kvidx_synchronize(kvidx_t* idx);

/* Score() over device-resident inputs and outputs, launched on the handle's stream.  Batches below 4096 prompts
 * (KVIDX_ROUNDS_MIN) are a single asynchronous kernel; larger ones run the round pipelines, which wait on the stream ONCE at the start (the number
 * of rounds is the longest prompt's block count, computed on the device) and are asynchronous after that.  Results are
 * complete when the stream reaches the point after the call (kvidx_synchronize, or an event recorded by the caller). */
int kvidx_score_batch_dev(kvidx_t* idx, const uint32_t* d_tok, const int64_t* d_tok_off, int64_t n_prompts,
                          const uint32_t* d_model, uint32_t model0, const uint64_t* d_filter,
                          double* d_scores_out, uint8_t* d_has_keys_out);
int kvidx_score_batch_sparse_dev(kvidx_t* idx, const uint32_t* d_tok, const int64_t* d_tok_off, int64_t n_prompts,
                                 const uint32_t* d_model, uint32_t model0, const uint64_t* d_filter,
                                 uint16_t* d_pods_out, double* d_scores_out, uint8_t* d_cnt_out, uint8_t* d_has_keys_out);
int kvidx_hash_keys_dev(kvidx_t* idx, const uint32_t* d_tok, const int64_t* d_tok_off, int64_t n_prompts,
                        const uint64_t* d_parent, const uint8_t* d_parent_valid,
                        const int64_t* d_key_off, uint64_t* d_keys_out);
/* Pool.digestEvents for a device-resident batch: events stably sorted by pod (queue q owns
 * d_ev_sorted[d_queue_off[q] .. d_queue_off[q+1])), n_events in all, n_hashes engine hashes behind them.  Asynchronous on
 * the handle's WRITE stream; read calls issued after it returns are ordered after the batch.  d_n_dropped (may be NULL)
 * receives the handle's cumulative dropped-event count once the batch is done. */
int kvidx_apply_events_dev(kvidx_t* idx, const kvidx_event_t* d_ev_sorted, const int64_t* d_queue_off,
                           int64_t n_queues, int64_t n_events, const uint64_t* d_hashes, int64_t n_hashes,
                           const uint32_t* d_tokens, int64_t* d_n_dropped);

/* ---- hash-range sharding across GPUs (one process per GPU) ------------------------
 * The request and engine tables are partitioned by the top bits of the mixed key.  Every rank maps its peers'
 * shards (CUDA IPC over NVLink / NVSwitch peer memory) and the SAME kernels then probe, lock and update slots
 * wherever they live: a Score() probe of a remote key is a 64-byte peer load issued from the fused walk, an
 * Add / Evict is a system-scope CAS on the owner's slot.  No collective sits on the data path; NCCL (or any
 * transport) is only needed to exchange the 192-byte handle blobs once.
 * Ingest rule: all events of one pod must be applied through ONE rank (per-pod order, kvevents/pool.go:129-144).
 * Before a write batch every owner's fill level is read through the mapped memory; a full owner fails the call with
 * KVIDX_ENOSPC (and an insert that still finds no slot is refused and counted), nothing spins. */
#define KVIDX_SHARD_HANDLE_BYTES 192
int kvidx_shard_export(kvidx_t* idx, void* handle_out /* KVIDX_SHARD_HANDLE_BYTES */);
int kvidx_shard_import(kvidx_t* idx, uint32_t rank, const void* handle /* from that rank's kvidx_shard_export */);
/* same-process variant (several GPUs driven by one process, e.g. tests): map `other`'s shard directly */
int kvidx_shard_attach(kvidx_t* idx, uint32_t rank, kvidx_t* other);
/* Drop the tombstones of THIS rank's shard (steady BlockStored / BlockRemoved churn leaves them behind; a write call
 * answers KVIDX_ENOSPC when an owner runs out of room).  Collective by contract: every rank calls it between two barriers
 * of the embedding program, no rank touches the index meanwhile (kvidx.dist.compact_shards does exactly that).  On an
 * unsharded handle it is the compaction the write path runs by itself when needed. */
int kvidx_shard_compact(kvidx_t* idx);

/* The ROUTED form of a sharded Score() (all-to-all of keys to their owners, slot images back), as three
 * device steps around an exchange the embedding program performs (kvidx.dist.score_alltoall does it with NCCL
 * all_to_all_single).  Provided so that the two designs can be measured against each other; the library's own sharded
 * Score() probes the owner's shard directly from the walk (see above) and needs none of this.
 *   key_owners : owner rank of every key (the routing decision)
 *   probe_slots: owner side -- the 32-byte slot image of every key of THIS shard (all zero: not in the index)
 *   score_slots: origin side -- consecutive-prefix walk + score over the returned images, prompt i owning images
 *                d_key_off[i] .. d_key_off[i+1]; dense rows like kvidx_score_batch_dev. */
int kvidx_key_owners_dev(kvidx_t* idx, const uint64_t* d_keys, const uint32_t* d_model, uint32_t model0, int64_t n, uint8_t* d_owner_out);
int kvidx_probe_slots_dev(kvidx_t* idx, const uint64_t* d_keys, const uint32_t* d_model, uint32_t model0, int64_t n, void* d_slots_out);
int kvidx_score_slots_dev(kvidx_t* idx, const void* d_slots, const int64_t* d_key_off, int64_t n_prompts, const uint64_t* d_filter,
                          double* d_scores_out, uint8_t* d_has_keys_out);

/* ---- introspection ------------------------------------------------------------ */
typedef struct kvidx_stats {
    uint64_t request_keys;      /* resident request keys (lru data.Len())       */
    uint64_t engine_keys;       /* resident engine->request mappings            */
    uint64_t request_tombs;
    uint64_t engine_tombs;
    uint64_t request_slots;
    uint64_t engine_slots;
    uint64_t rebuilds;
    uint64_t kernel_launches;   /* kernels this handle has launched so far      */
    uint64_t rehashed_events;   /* BlockStored events whose parent resolved differently at apply time than when their
                                   keys were precomputed (kernels_write.cuh phase 1): re-hashed in place             */
    uint64_t coalesced_calls;   /* host-buffer Score() calls that were served as part of another call's launch      */
} kvidx_stats_t;
int kvidx_get_stats(kvidx_t* idx, kvidx_stats_t* out);

#ifdef __cplusplus
}
#endif
#endif /* KVIDX_H */
