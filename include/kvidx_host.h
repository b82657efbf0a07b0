/*
 * kvidx_host.h -- string-level host mirror of the reference's Go interfaces for the hot path, in C++ behind a C ABI.
 *
 * The reference's host code is Go (kvcache.Indexer, kvblock.Index, kvevents.Pool); Go is not available in this build
 * image, so the layer a Go shim would provide -- string interning, the per-pod sharded event queues and the msgpack
 * decoding of KVEvents -- is provided here in C++ over libkvidx's id-level ABI (kvidx.h), with the same names, argument
 * meaning and error behaviour.  The decoder is pure host code and works without a GPU (tests/test_host_decode.py).
 *
 *   kvhost_create            <-> kvcache.NewKVCacheIndexer            pkg/kvcache/indexer.go:75-113
 *   kvhost_get_pod_scores    <-> Indexer.GetPodScores (steps 2-4)     pkg/kvcache/indexer.go:132-166
 *   kvhost_index_*           <-> kvblock.Index Add/Evict/Lookup/GetRequestKey   pkg/kvcache/kvblock/index.go:119-135
 *   kvhost_pool_add_task     <-> kvevents.Pool.AddTask                pkg/kvcache/kvevents/pool.go:132-144
 *   kvhost_pool_process      <-> worker loop: processEvent + digestEvents       pkg/kvcache/kvevents/pool.go:149-338
 *   kvhost_decode_event_batch<-> processEvent's msgpack decoding      pkg/kvcache/kvevents/pool.go:177-244, events.go:38-96
 *   kvhost_prefix_store_*    <-> prefixstore.LRUTokenStore            pkg/tokenization/prefixstore/lru_store.go:54-190
 *   kvhost_get_metrics / _text <-> InstrumentedIndex + collectors      pkg/kvcache/kvblock/instrumented_index.go:30-92, metrics/collector.go:28-59
 */
#ifndef KVIDX_HOST_H
#define KVIDX_HOST_H
#include "kvidx.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct kvhost kvhost_t;

typedef struct kvhost_config {
    kvidx_config_t index;            /* device index (tier weights are taken from tier_names / tier_weights below)      */
    uint32_t concurrency;            /* kvevents.Config.Concurrency (default 4): number of FNV-32a(pod) sharded queues  */
    uint32_t n_tiers;                /* KVCacheBackendConfig list: names[i] scores weights[i]; ids follow this order     */
    const char* tier_names[KVIDX_MAX_TIERS];
    double tier_weights[KVIDX_MAX_TIERS];
    int32_t no_device;               /* 1: host-only instance (interners, queues, decoder) without a kvidx handle       */
    int32_t enable_metrics;          /* IndexConfig.EnableMetrics (index.go:41-43): wrap the index in the instrumented one */
} kvhost_config_t;

void kvhost_config_default(kvhost_config_t* cfg);       /* block 16, seed "", gpu 1.0 / cpu 0.8, concurrency 4 */
int  kvhost_create(const kvhost_config_t* cfg, const char* hash_seed, kvhost_t** out);
void kvhost_destroy(kvhost_t* h);
const char* kvhost_last_error(void);
kvidx_t* kvhost_index(kvhost_t* h);                      /* Indexer.KVBlockIndex(), indexer.go:121-123 */

/* Indexer.GetPodScores after tokenisation.  pods[n_pods] is the filter (n_pods == 0: all pods).  Results: up to 10
 * (pod name, score) pairs; the name pointers stay valid for the life of the handle.  Returns the number of pairs,
 * -1000 for "no full block" (the reference's (nil, nil)), or a negative KVIDX_E* code. */
int kvhost_get_pod_scores(kvhost_t* h, const uint32_t* tokens, size_t n_tokens, const char* model,
                          const char* const* pods, size_t n_pods, const char** pod_out, double* score_out);

/* kvblock.Index with strings.  entries are (pod, tier) name pairs. */
int kvhost_index_add(kvhost_t* h, const char* model, const uint64_t* engine, size_t n_engine, const uint64_t* request, size_t n_request,
                     const char* const* pods, const char* const* tiers, size_t n_entries);
int kvhost_index_evict(kvhost_t* h, const char* model, uint64_t engine, const char* const* pods, const char* const* tiers, size_t n_entries);
int kvhost_index_get_request_key(kvhost_t* h, const char* model, uint64_t engine, uint64_t* out);
/* Lookup: for key i, cnt_out[i] entries; entry j's names at pod_out / tier_out [i*10 + j]. */
int kvhost_index_lookup(kvhost_t* h, const char* model, const uint64_t* keys, size_t n, const char* const* pods, size_t n_pods,
                        const char** pod_out, const char** tier_out, uint8_t* cnt_out);

/* kvevents.Pool.  add_task copies the payload into queue FNV-32a(pod) % concurrency; process drains every queue in
 * FIFO order, decodes the msgpack EventBatch of each message and applies all resulting events with ONE
 * kvidx_apply_events call.  Returns the number of events applied (or a negative code). */
int kvhost_pool_add_task(kvhost_t* h, const char* pod, const char* model, const void* payload, size_t len);
int kvhost_pool_queue_index(kvhost_t* h, const char* pod);
int64_t kvhost_pool_process(kvhost_t* h, int64_t* n_dropped_out);

/* Pure host decoding of one message payload into kvidx_event_t records (+ flat hash / token arrays).  Writes at most
 * the given capacities and returns the number of events, or -KVIDX_ENOSPC-style negatives.  Malformed batches decode
 * to 0 events (poison pill dropped, pool.go:182-187); malformed events are skipped (pool.go:190-239).
 * Hashes rejected by getHashAsUint64 (pool.go:343-367) are left out of a BlockStored's list (pool.go:272-275); a rejected
 * parent hash skips the event (pool.go:283-287).  AllBlocksCleared decodes to nothing (pool.go:332-333). */
int64_t kvhost_decode_event_batch(kvhost_t* h, const char* pod, const char* model, const void* payload, size_t len,
                                  kvidx_event_t* ev_out, size_t ev_cap, uint64_t* hash_out, size_t hash_cap, size_t* n_hash_out,
                                  uint32_t* tok_out, size_t tok_cap, size_t* n_tok_out);

/* kvcache_index_* metrics (metrics/collector.go:28-59) with the instrumented index's semantics (instrumented_index.go:35-92):
 * Add counts len(requestKeys) admissions and Evict len(entries) evictions whatever the call returned; every Lookup counts
 * one request and one latency observation; a successful Lookup adds max over pods of that pod's entries in the result to
 * both max_pod_hit_count and lookup_hits.  The event pool goes through the same Add / Evict (one Evict per removed hash,
 * pool.go:317-330).  With enable_metrics, kvhost_get_pod_scores issues the reference's full-depth Lookup
 * (kvidx_hash_keys + kvidx_lookup on the device) next to the fused score call, because the hit count needs every key of
 * the prompt, not only the consecutive prefix the scorer walks.  All zero when enable_metrics is 0. */
#define KVHOST_LATENCY_BUCKETS 11          /* prometheus.DefBuckets: .005 .01 .025 .05 .1 .25 .5 1 2.5 5 10 (+Inf = count) */
typedef struct kvhost_metrics {
    uint64_t admissions_total, evictions_total, lookup_requests_total, max_pod_hit_count_total, lookup_hits_total;
    uint64_t lookup_latency_bucket[KVHOST_LATENCY_BUCKETS];      /* cumulative, le = bucket bound */
    uint64_t lookup_latency_count;
    double lookup_latency_sum;                                   /* seconds */
} kvhost_metrics_t;
int kvhost_get_metrics(kvhost_t* h, kvhost_metrics_t* out);
/* Prometheus text exposition of the above (same metric names and help strings as the reference).  Returns the number of
 * bytes needed (excluding the NUL); writes at most cap - 1 bytes + NUL. */
int64_t kvhost_metrics_text(kvhost_t* h, char* buf, size_t cap);

/* Tokenization prefix store (the cache in front of the tokenizer; the tokenizer itself stays with the
 * embedding program).  Text is cut into block_size-BYTE blocks, block i is keyed by XXH64(seed 0) of the little-endian
 * previous key followed by the block's bytes (lru_store.go:111-121), and maps to the tokens whose END offset falls inside
 * the text up to the block's end (:124-137).  A partial last block is never stored or looked up.
 *   create: NewLRUTokenStore (:72-87); cache_size <= 0 is an error like lru.New; defaults 500000 blocks of 256 bytes.
 *   add:    AddTokenization (:89-141); offsets are n_tokens pairs [low, high) of byte offsets; empty prompt or no
 *           tokens is a no-op.
 *   find:   FindLongestContainedTokens (:143-190): walks the blocks until the first one missing from the cache (a hit
 *           refreshes its recency); returns the number of contained tokens (all of them are written if cap allows, else
 *           the first cap) and the covered fraction of the prompt, end of the last matched block / prompt length.
 * tokenization.Pool.processTask (pool.go:209-225) is: find; if the ratio is below 0.8 tokenize, add, use the fresh tokens. */
typedef struct kvhost_prefix_store kvhost_prefix_store_t;
int  kvhost_prefix_store_create(int64_t cache_size, int32_t block_size, kvhost_prefix_store_t** out);
void kvhost_prefix_store_destroy(kvhost_prefix_store_t* s);
int  kvhost_prefix_store_add(kvhost_prefix_store_t* s, const char* prompt, size_t prompt_len, const uint32_t* tokens,
                             const uint64_t* offsets, size_t n_tokens);
int64_t kvhost_prefix_store_find(kvhost_prefix_store_t* s, const char* prompt, size_t prompt_len, uint32_t* tokens_out, size_t cap,
                                 double* overlap_ratio_out);
int64_t kvhost_prefix_store_len(kvhost_prefix_store_t* s);           /* blocks resident */
uint64_t kvhost_xxhash64(const void* data, size_t len, uint64_t seed);

/* interning introspection (ids are append-only and never reused) */
int kvhost_pod_id(kvhost_t* h, const char* pod);
int kvhost_tier_id(kvhost_t* h, const char* tier);
int kvhost_model_id(kvhost_t* h, const char* model);

#ifdef __cplusplus
}
#endif
#endif
