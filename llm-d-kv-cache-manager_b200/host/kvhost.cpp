// kvhost.cpp -- host mirror of the reference's Go layer for the hot path (include/kvidx_host.h).
// String interning, FNV-32a sharded per-pod queues (kvevents.Pool) and the msgpack decoding of KVEvents, over the
// id-level C ABI of libkvidx.  No CUDA in this file; the decoder runs without a GPU.
#include <algorithm>
#include <cctype>
#include <cstring>
#include <chrono>
#include <deque>
#include <list>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/kvidx_host.h"

namespace {

thread_local std::string g_herr;
int hfail(int code, const std::string& m) { g_herr = m; return code; }

struct Interner {
    std::unordered_map<std::string, uint32_t> ids;
    std::deque<std::string> names;                   // deque: element addresses are stable
    int id(const std::string& s, uint32_t limit) {
        auto it = ids.find(s);
        if (it != ids.end()) return (int)it->second;
        if (names.size() >= limit) return -1;
        names.push_back(s);
        ids.emplace(s, (uint32_t)names.size() - 1);
        return (int)names.size() - 1;
    }
    int find(const std::string& s) const { auto it = ids.find(s); return it == ids.end() ? -1 : (int)it->second; }
};

// ---- minimal msgpack reader (vmihailenco/msgpack v5 type rules are applied by the callers) -------------------
enum Kind { K_NIL, K_BOOL, K_INT, K_FLOAT, K_STR, K_BIN, K_ARRAY, K_MAP, K_EXT, K_BAD };
struct Val { Kind k = K_BAD; uint8_t code = 0; uint64_t u = 0; double f = 0; const uint8_t* s = nullptr; size_t n = 0; };
struct Cur { const uint8_t* p; const uint8_t* e; };

bool need(Cur& c, size_t n) { return (size_t)(c.e - c.p) >= n; }
uint64_t be(const uint8_t* p, int n) { uint64_t v = 0; for (int i = 0; i < n; ++i) v = (v << 8) | p[i]; return v; }

// Reads one element head; scalars, strings and bins are consumed completely, containers leave the cursor at their
// first child (n = element count; maps count pairs).
bool head(Cur& c, Val& v) {
    if (!need(c, 1)) return false;
    const uint8_t b = *c.p++;
    v.code = b; v.u = 0; v.n = 0; v.s = nullptr;
    auto rd = [&](int n, uint64_t& out) { if (!need(c, n)) return false; out = be(c.p, n); c.p += n; return true; };
    auto blob = [&](size_t n, Kind k) { if (!need(c, n)) return false; v.k = k; v.s = c.p; v.n = n; c.p += n; return true; };
    uint64_t t = 0;
    if (b <= 0x7f) { v.k = K_INT; v.u = b; return true; }
    if (b >= 0xe0) { v.k = K_INT; v.u = (uint64_t)(int64_t)(int8_t)b; return true; }
    if (b >= 0xa0 && b <= 0xbf) return blob(b & 0x1f, K_STR);
    if (b >= 0x90 && b <= 0x9f) { v.k = K_ARRAY; v.n = b & 0x0f; return true; }
    if (b >= 0x80 && b <= 0x8f) { v.k = K_MAP; v.n = b & 0x0f; return true; }
    switch (b) {
        case 0xc0: v.k = K_NIL; return true;
        case 0xc2: case 0xc3: v.k = K_BOOL; v.u = b & 1; return true;
        case 0xc4: return rd(1, t) && blob(t, K_BIN);
        case 0xc5: return rd(2, t) && blob(t, K_BIN);
        case 0xc6: return rd(4, t) && blob(t, K_BIN);
        case 0xc7: return rd(1, t) && need(c, 1) && (c.p++, blob(t, K_EXT));
        case 0xc8: return rd(2, t) && need(c, 1) && (c.p++, blob(t, K_EXT));
        case 0xc9: return rd(4, t) && need(c, 1) && (c.p++, blob(t, K_EXT));
        case 0xca: { if (!rd(4, t)) return false; uint32_t w = (uint32_t)t; float f; memcpy(&f, &w, 4); v.k = K_FLOAT; v.f = f; return true; }
        case 0xcb: { if (!rd(8, t)) return false; double d; memcpy(&d, &t, 8); v.k = K_FLOAT; v.f = d; return true; }
        case 0xcc: v.k = K_INT; return rd(1, v.u);
        case 0xcd: v.k = K_INT; return rd(2, v.u);
        case 0xce: v.k = K_INT; return rd(4, v.u);
        case 0xcf: v.k = K_INT; return rd(8, v.u);
        case 0xd0: v.k = K_INT; if (!rd(1, t)) return false; v.u = (uint64_t)(int64_t)(int8_t)t; return true;
        case 0xd1: v.k = K_INT; if (!rd(2, t)) return false; v.u = (uint64_t)(int64_t)(int16_t)t; return true;
        case 0xd2: v.k = K_INT; if (!rd(4, t)) return false; v.u = (uint64_t)(int64_t)(int32_t)t; return true;
        case 0xd3: v.k = K_INT; return rd(8, v.u);
        case 0xd4: return need(c, 1) && (c.p++, blob(1, K_EXT));
        case 0xd5: return need(c, 1) && (c.p++, blob(2, K_EXT));
        case 0xd6: return need(c, 1) && (c.p++, blob(4, K_EXT));
        case 0xd7: return need(c, 1) && (c.p++, blob(8, K_EXT));
        case 0xd8: return need(c, 1) && (c.p++, blob(16, K_EXT));
        case 0xd9: return rd(1, t) && blob(t, K_STR);
        case 0xda: return rd(2, t) && blob(t, K_STR);
        case 0xdb: return rd(4, t) && blob(t, K_STR);
        case 0xdc: v.k = K_ARRAY; if (!rd(2, t)) return false; v.n = t; return true;
        case 0xdd: v.k = K_ARRAY; if (!rd(4, t)) return false; v.n = t; return true;
        case 0xde: v.k = K_MAP; if (!rd(2, t)) return false; v.n = t; return true;
        case 0xdf: v.k = K_MAP; if (!rd(4, t)) return false; v.n = t; return true;
        default: v.k = K_BAD; return false;     // 0xc1: never used
    }
}
bool skip(Cur& c, int depth = 0) {
    Val v;
    if (depth > 64 || !head(c, v)) return false;
    if (v.k == K_ARRAY) { for (size_t i = 0; i < v.n; ++i) if (!skip(c, depth + 1)) return false; }
    else if (v.k == K_MAP) { for (size_t i = 0; i < 2 * v.n; ++i) if (!skip(c, depth + 1)) return false; }
    return true;
}

// getHashAsUint64 (kvevents/pool.go:343-367) on a value decoded the way msgpack's DecodeInterface types it:
// only a uint64-coded (0xcf) or int64-coded (0xd3) integer or a byte slice (bin) is accepted.
bool hash_of(Cur& c, uint64_t* out, bool* is_nil) {
    const uint8_t* start = c.p;
    Val v;
    if (is_nil) *is_nil = false;
    Cur probe = c;
    if (!head(probe, v)) { c.p = c.e; return false; }
    if (v.k == K_ARRAY || v.k == K_MAP) { c.p = start; if (!skip(c)) c.p = c.e; return false; }
    c = probe;
    if (v.k == K_NIL) { if (is_nil) *is_nil = true; return false; }
    if (v.k == K_INT && (v.code == 0xcf || v.code == 0xd3)) { *out = v.u; return true; }
    if (v.k == K_BIN) {
        if (v.n == 0) return false;
        if (v.n >= 8) { *out = be(v.s + v.n - 8, 8); return true; }
        *out = be(v.s, (int)v.n);
        return true;
    }
    return false;
}

struct Msg { uint32_t pod, model; std::string payload; };

}  // namespace

struct kvhost {
    kvidx_t* ix = nullptr;
    kvhost_config_t cfg{};
    Interner models, pods, tiers;
    std::mutex mu;
    std::vector<std::deque<Msg>> queues;
    uint32_t filter_words = 4;
    uint32_t pod_cap = 256;                          // == index.max_pods: ids of pods that own entries
    kvhost_metrics_t met{};              // guarded by mu
};

namespace {

std::string lower(std::string s) { for (auto& ch : s) ch = (char)tolower((unsigned char)ch); return s; }

// one event slice [p, e): a tagged union array ["BlockStored", ...]
void decode_event(kvhost* h, uint32_t pod, uint32_t model, const uint8_t* p, const uint8_t* e, std::vector<kvidx_event_t>& evs,
                  std::vector<uint64_t>& hashes, std::vector<uint32_t>& toks) {
    Cur c{p, e};
    Val v;
    if (!head(c, v)) return;
    if (v.k != K_ARRAY || v.n < 1) return;                               // not an array / no tag element: skipped
    size_t rest = v.n - 1;
    Val tag;
    if (!head(c, tag)) return;
    if (tag.k != K_STR && tag.k != K_BIN) return;                         // tag must decode as string
    const std::string t((const char*)tag.s, tag.n);
    auto medium_tier = [&](Cur& cc, bool present, int* tier_out) -> bool {  // *string, lower-cased; default "gpu"
        std::string name = "gpu";
        if (present) {
            Val m;
            if (!head(cc, m)) return false;
            if (m.k == K_STR || m.k == K_BIN) name = lower(std::string((const char*)m.s, m.n));
            else if (m.k != K_NIL) return false;
        }
        const int id = h->tiers.id(name, KVIDX_MAX_TIERS);
        if (id < 0) return false;
        *tier_out = id;
        return true;
    };
    if (t == "BlockStored") {
        const size_t h0 = hashes.size(), t0 = toks.size();
        kvidx_event_t ev{};
        ev.op = KVIDX_EV_BLOCK_STORED; ev.model = model; ev.hash_off = h0; ev.tok_off = t0;
        bool ok = true, parent_bad = false;
        size_t f = 0;
        if (f < rest) {                                                   // BlockHashes []any
            Val a;
            if (!head(c, a)) return;
            if (a.k == K_ARRAY) { for (size_t i = 0; i < a.n && ok; ++i) { uint64_t hv; const uint8_t* before = c.p; if (hash_of(c, &hv, nullptr)) hashes.push_back(hv); if (c.p == before || c.p > e) ok = false; } }
            else if (a.k != K_NIL) ok = false;
            ++f;
        }
        if (ok && f < rest) {                                             // ParentBlockHash any
            uint64_t hv; bool nil = false;
            if (hash_of(c, &hv, &nil)) { ev.has_parent = 1; ev.parent_hash = hv; }
            else if (!nil) parent_bad = true;                             // present but unsupported type: event skipped (pool.go:283-287)
            ++f;
        }
        if (ok && f < rest) {                                             // TokenIds []uint32
            Val a;
            if (!head(c, a)) ok = false;
            else if (a.k == K_ARRAY) { for (size_t i = 0; i < a.n && ok; ++i) { Val x; if (!head(c, x) || x.k != K_INT) ok = false; else toks.push_back((uint32_t)x.u); } }
            else if (a.k != K_NIL) ok = false;
            ++f;
        }
        if (ok && f < rest) { Val x; if (!head(c, x) || (x.k != K_INT && x.k != K_NIL)) ok = false; ++f; }   // BlockSize int (ignored)
        if (ok && f < rest) { Val x; if (!head(c, x) || (x.k != K_INT && x.k != K_NIL)) ok = false; ++f; }   // LoraID *int
        int tier = 0;
        if (ok) { ok = medium_tier(c, f < rest, &tier); if (f < rest) ++f; }
        for (; ok && f < rest; ++f) ok = skip(c);                         // extra fields are skipped
        if (!ok || parent_bad) { hashes.resize(h0); toks.resize(t0); return; }
        ev.podtier = KVIDX_PODTIER(pod, tier);
        ev.n_hashes = (uint32_t)(hashes.size() - h0); ev.n_tokens = (uint32_t)(toks.size() - t0);
        evs.push_back(ev);
    } else if (t == "BlockRemoved") {
        const size_t h0 = hashes.size();
        kvidx_event_t ev{};
        ev.op = KVIDX_EV_BLOCK_REMOVED; ev.model = model; ev.hash_off = h0; ev.tok_off = toks.size();
        bool ok = true;
        size_t f = 0;
        if (f < rest) {
            Val a;
            if (!head(c, a)) return;
            if (a.k == K_ARRAY) { for (size_t i = 0; i < a.n && ok; ++i) { uint64_t hv; const uint8_t* before = c.p; if (hash_of(c, &hv, nullptr)) hashes.push_back(hv); if (c.p == before || c.p > e) ok = false; } }
            else if (a.k != K_NIL) ok = false;
            ++f;
        }
        int tier = 0;
        if (ok) { ok = medium_tier(c, f < rest, &tier); if (f < rest) ++f; }
        for (; ok && f < rest; ++f) ok = skip(c);
        if (!ok) { hashes.resize(h0); return; }
        ev.podtier = KVIDX_PODTIER(pod, tier);
        ev.n_hashes = (uint32_t)(hashes.size() - h0);
        evs.push_back(ev);
    }
    // "AllBlocksCleared": no-op (pool.go:332-333); unknown tags: skipped (pool.go:229-231)
}

// EventBatch = [ts, events[], data_parallel_rank?] (events.go:38-43).  Any decode error drops the whole message.
void decode_batch(kvhost* h, uint32_t pod, uint32_t model, const uint8_t* p, size_t len, std::vector<kvidx_event_t>& evs,
                  std::vector<uint64_t>& hashes, std::vector<uint32_t>& toks) {
    Cur c{p, p + len};
    Val v;
    if (!head(c, v) || v.k != K_ARRAY) return;
    const size_t n = v.n;
    std::vector<std::pair<const uint8_t*, const uint8_t*>> raw;
    if (n >= 1) { Val ts; if (!head(c, ts) || (ts.k != K_FLOAT && ts.k != K_INT && ts.k != K_NIL)) return; }
    if (n >= 2) {
        Val a;
        if (!head(c, a)) return;
        if (a.k == K_ARRAY) {
            for (size_t i = 0; i < a.n; ++i) { const uint8_t* s = c.p; if (!skip(c)) return; raw.emplace_back(s, c.p); }
        } else if (a.k != K_NIL) return;
    }
    if (n >= 3) { Val r; if (!head(c, r) || (r.k != K_INT && r.k != K_NIL)) return; }
    for (size_t i = 3; i < n; ++i) if (!skip(c)) return;
    for (auto& r : raw) decode_event(h, pod, model, r.first, r.second, evs, hashes, toks);
}

}  // namespace

static const double kLatencyBuckets[KVHOST_LATENCY_BUCKETS] = {0.005, 0.01, 0.025, 0.05, 0.1, 0.25, 0.5, 1.0, 2.5, 5.0, 10.0};   // prometheus.DefBuckets

extern "C" {

const char* kvhost_last_error(void) { return g_herr.c_str(); }

void kvhost_config_default(kvhost_config_t* c) {
    if (!c) return;
    memset(c, 0, sizeof *c);
    kvidx_config_default(&c->index);
    c->concurrency = 4;                      // kvevents/pool.go:52
    c->n_tiers = 2;
    c->tier_names[0] = "gpu"; c->tier_weights[0] = 1.0;     // backend.go:28
    c->tier_names[1] = "cpu"; c->tier_weights[1] = 0.8;     // backend.go:29
}

int kvhost_create(const kvhost_config_t* cfg_in, const char* hash_seed, kvhost_t** out) {
    if (!out) return hfail(KVIDX_EINVAL, "out is NULL");
    *out = nullptr;
    kvhost_config_t c;
    kvhost_config_default(&c);
    if (cfg_in) c = *cfg_in;
    if (c.concurrency == 0) c.concurrency = 4;
    if (c.n_tiers > KVIDX_MAX_TIERS) return hfail(KVIDX_ERANGE, "too many tiers");
    kvhost* h = new kvhost();
    h->cfg = c;
    h->queues.resize(c.concurrency);
    for (uint32_t i = 0; i < c.n_tiers; ++i) h->tiers.id(lower(c.tier_names[i] ? c.tier_names[i] : ""), KVIDX_MAX_TIERS);
    const char* seed = hash_seed ? hash_seed : "";
    c.index.init_hash = kvidx_fnv64a(seed, strlen(seed));           // getInitHash, token_processor.go:81-90
    c.index.n_tier_weights = c.n_tiers;
    for (uint32_t i = 0; i < KVIDX_MAX_TIERS; ++i) c.index.tier_weight[i] = i < c.n_tiers ? c.tier_weights[i] : 1.0;
    h->filter_words = ((c.index.max_pods ? c.index.max_pods : 256) + 63) / 64;
    h->pod_cap = c.index.max_pods ? c.index.max_pods : 256;     // pods that own index entries must fit the score / filter row width
    if (!c.no_device) {
        const int rc = kvidx_create(&c.index, &h->ix);
        if (rc) { g_herr = kvidx_last_error(nullptr); delete h; return rc; }
    }
    *out = h;
    return 0;
}

void kvhost_destroy(kvhost_t* h) { if (!h) return; if (h->ix) kvidx_destroy(h->ix); delete h; }
kvidx_t* kvhost_index(kvhost_t* h) { return h ? h->ix : nullptr; }

int kvhost_get_metrics(kvhost_t* h, kvhost_metrics_t* out) {
    if (!h || !out) return hfail(KVIDX_EINVAL, "bad arguments");
    std::lock_guard<std::mutex> g(h->mu);
    *out = h->met;
    return 0;
}

int64_t kvhost_metrics_text(kvhost_t* h, char* buf, size_t cap) {
    kvhost_metrics_t m;
    if (kvhost_get_metrics(h, &m)) return KVIDX_EINVAL;
    std::string s;
    auto counter = [&](const char* name, const char* help, uint64_t v) {
        s += "# HELP kvcache_index_"; s += name; s += " "; s += help; s += "\n# TYPE kvcache_index_"; s += name; s += " counter\nkvcache_index_";
        s += name; s += " " + std::to_string(v) + "\n";
    };
    counter("admissions_total", "Total number of KV-block admissions", m.admissions_total);
    counter("evictions_total", "Total number of KV-block evictions", m.evictions_total);
    counter("lookup_requests_total", "Total number of lookup calls", m.lookup_requests_total);
    counter("max_pod_hit_count_total", "Maximum cache hits on a single pod on Lookup()", m.max_pod_hit_count_total);
    counter("lookup_hits_total", "Number of keys found in the cache on Lookup()", m.lookup_hits_total);
    s += "# HELP kvcache_index_lookup_latency_seconds Latency of Lookup calls in seconds\n# TYPE kvcache_index_lookup_latency_seconds histogram\n";
    char tmp[96];
    for (int b = 0; b < KVHOST_LATENCY_BUCKETS; ++b) {
        snprintf(tmp, sizeof tmp, "kvcache_index_lookup_latency_seconds_bucket{le=\"%g\"} %llu\n", kLatencyBuckets[b], (unsigned long long)m.lookup_latency_bucket[b]);
        s += tmp;
    }
    snprintf(tmp, sizeof tmp, "kvcache_index_lookup_latency_seconds_bucket{le=\"+Inf\"} %llu\n", (unsigned long long)m.lookup_latency_count); s += tmp;
    snprintf(tmp, sizeof tmp, "kvcache_index_lookup_latency_seconds_sum %.9g\n", m.lookup_latency_sum); s += tmp;
    snprintf(tmp, sizeof tmp, "kvcache_index_lookup_latency_seconds_count %llu\n", (unsigned long long)m.lookup_latency_count); s += tmp;
    if (buf && cap) { const size_t n = std::min(cap - 1, s.size()); memcpy(buf, s.data(), n); buf[n] = 0; }
    return (int64_t)s.size();
}

int kvhost_pod_id(kvhost_t* h, const char* s) { std::lock_guard<std::mutex> g(h->mu); return h->pods.id(s, KVIDX_MAX_PODS); }
int kvhost_tier_id(kvhost_t* h, const char* s) { std::lock_guard<std::mutex> g(h->mu); return h->tiers.id(lower(s), KVIDX_MAX_TIERS); }
int kvhost_model_id(kvhost_t* h, const char* s) { std::lock_guard<std::mutex> g(h->mu); return h->models.id(s, 65536); }

static int need_dev(kvhost* h) { return h && h->ix ? 0 : hfail(KVIDX_ECUDA, "host-only instance: no device index (there is no CPU fallback)"); }

// One Lookup as the instrumented index sees it (instrumented_index.go:47-69): request + latency always, hits on success.
static void observe_lookup(kvhost* h, double seconds, bool ok, const kvidx_podtier_t* pt, const uint8_t* cnt, size_t n) {
    if (!h->cfg.enable_metrics) return;
    uint64_t mx = 0;
    if (ok) {                                           // recordHitMetrics: entries per pod over the whole result
        std::vector<uint32_t> per(KVIDX_MAX_PODS, 0);
        for (size_t i = 0; i < n; ++i)
            for (int j = 0; j < cnt[i]; ++j) { const uint32_t c = ++per[KVIDX_PT_POD(pt[i * KVIDX_MAX_PODS_PER_KEY + j])]; if (c > mx) mx = c; }
    }
    std::lock_guard<std::mutex> g(h->mu);
    h->met.lookup_requests_total += 1;
    h->met.lookup_latency_count += 1; h->met.lookup_latency_sum += seconds;
    for (int b = 0; b < KVHOST_LATENCY_BUCKETS; ++b) if (seconds <= kLatencyBuckets[b]) h->met.lookup_latency_bucket[b] += 1;
    h->met.max_pod_hit_count_total += mx; h->met.lookup_hits_total += mx;
}

static bool build_filter(kvhost* h, const char* const* pods, size_t n_pods, std::vector<uint64_t>& mask) {
    // sets.New(podIdentifiers...): a name never seen by the index cannot match any entry, but still makes the set non-empty
    mask.assign(h->filter_words, 0);
    if (n_pods == 0) return false;
    for (size_t i = 0; i < n_pods; ++i) {
        const int id = h->pods.id(pods[i], KVIDX_MAX_PODS);
        if (id >= 0 && (uint32_t)id < h->filter_words * 64) mask[id >> 6] |= 1ull << (id & 63);
    }
    return true;
}

int kvhost_get_pod_scores(kvhost_t* h, const uint32_t* tokens, size_t n_tokens, const char* model, const char* const* pods, size_t n_pods,
                          const char** pod_out, double* score_out) {
    if (int rc = need_dev(h)) return rc;
    std::vector<uint64_t> mask;
    bool filtered; int mid;
    { std::lock_guard<std::mutex> g(h->mu); filtered = build_filter(h, pods, n_pods, mask); mid = h->models.id(model, 65536); }
    if (mid < 0) return hfail(KVIDX_ERANGE, "too many models");
    if (filtered) {                                 // a filter whose pods all lie outside the mask width selects nothing
        bool any = false; for (uint64_t w : mask) any |= w != 0;
        if (!any) {
            const bool has_keys = n_tokens >= (h->cfg.index.block_size ? h->cfg.index.block_size : 16);
            if (has_keys) observe_lookup(h, 0.0, true, nullptr, nullptr, 0);      // the Lookup still happens; it just finds no listed pod
            return has_keys ? 0 : -1000;
        }
    }
    const int64_t off[2] = {0, (int64_t)n_tokens};
    uint16_t ids[KVIDX_MAX_PODS_PER_KEY]; double sc[KVIDX_MAX_PODS_PER_KEY]; uint8_t cnt = 0, has = 0;
    const int rc = kvidx_score_batch_sparse(h->ix, tokens, off, 1, nullptr, (uint32_t)mid, filtered ? mask.data() : nullptr, ids, sc, &cnt, &has);
    if (rc) { g_herr = kvidx_last_error(h->ix); return rc; }
    if (!has) return -1000;                         // (nil, nil): no full block (indexer.go:142-146)
    if (h->cfg.enable_metrics) {                    // the instrumented Lookup of indexer.go:150: every key of the prompt
        const size_t nk = n_tokens / (h->cfg.index.block_size ? h->cfg.index.block_size : 16);
        std::vector<uint64_t> keys(nk); int64_t koff[2] = {0, 0};
        std::vector<kvidx_podtier_t> pt(nk * KVIDX_MAX_PODS_PER_KEY); std::vector<uint8_t> kc(nk);
        int r2 = kvidx_hash_keys(h->ix, tokens, off, 1, nullptr, nullptr, keys.data(), koff);
        const auto t0 = std::chrono::steady_clock::now();
        if (!r2) r2 = kvidx_lookup(h->ix, (uint32_t)mid, keys.data(), (int64_t)nk, filtered ? mask.data() : nullptr, pt.data(), kc.data());
        observe_lookup(h, std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(), r2 == 0, pt.data(), kc.data(), nk);
    }
    std::lock_guard<std::mutex> g(h->mu);
    for (int i = 0; i < cnt; ++i) { pod_out[i] = h->pods.names[ids[i]].c_str(); score_out[i] = sc[i]; }
    return cnt;
}

static int entries_of(kvhost* h, const char* const* pods, const char* const* tiers, size_t n, std::vector<kvidx_podtier_t>& out) {
    out.clear();
    for (size_t i = 0; i < n; ++i) {
        const int p = h->pods.id(pods[i], h->pod_cap), t = h->tiers.id(lower(tiers[i]), KVIDX_MAX_TIERS);
        if (p < 0 || t < 0) return hfail(KVIDX_ERANGE, "pod / tier id space exhausted");
        out.push_back(KVIDX_PODTIER(p, t));
    }
    return 0;
}

int kvhost_index_add(kvhost_t* h, const char* model, const uint64_t* engine, size_t n_engine, const uint64_t* request, size_t n_request,
                     const char* const* pods, const char* const* tiers, size_t n_entries) {
    if (int rc = need_dev(h)) return rc;
    if (h->cfg.enable_metrics) { std::lock_guard<std::mutex> g(h->mu); h->met.admissions_total += n_request; }   // whatever Add returns
    if (n_engine == 0 || n_request == 0 || n_entries == 0) return hfail(KVIDX_EINVAL, "no keys or entries provided for adding to index");
    if (n_engine != n_request) return hfail(KVIDX_EINVAL, "mismatch between engine keys and request keys length");
    std::vector<kvidx_podtier_t> pt; int mid;
    { std::lock_guard<std::mutex> g(h->mu); if (int rc = entries_of(h, pods, tiers, n_entries, pt)) return rc; mid = h->models.id(model, 65536); }
    const int rc = kvidx_add(h->ix, (uint32_t)mid, engine, request, (int64_t)n_engine, pt.data(), (int32_t)pt.size());
    if (rc) g_herr = kvidx_last_error(h->ix);
    return rc;
}

int kvhost_index_evict(kvhost_t* h, const char* model, uint64_t engine, const char* const* pods, const char* const* tiers, size_t n_entries) {
    if (int rc = need_dev(h)) return rc;
    if (h->cfg.enable_metrics) { std::lock_guard<std::mutex> g(h->mu); h->met.evictions_total += n_entries; }
    if (n_entries == 0) return hfail(KVIDX_EINVAL, "no entries provided for eviction from index");
    std::vector<kvidx_podtier_t> pt; int mid;
    { std::lock_guard<std::mutex> g(h->mu); if (int rc = entries_of(h, pods, tiers, n_entries, pt)) return rc; mid = h->models.id(model, 65536); }
    const int rc = kvidx_evict(h->ix, (uint32_t)mid, engine, pt.data(), (int32_t)pt.size());
    if (rc) g_herr = kvidx_last_error(h->ix);
    return rc;
}

int kvhost_index_get_request_key(kvhost_t* h, const char* model, uint64_t engine, uint64_t* out) {
    if (int rc = need_dev(h)) return rc;
    int mid; { std::lock_guard<std::mutex> g(h->mu); mid = h->models.id(model, 65536); }
    const int rc = kvidx_get_request_key(h->ix, (uint32_t)mid, engine, out);
    if (rc) g_herr = kvidx_last_error(h->ix);
    return rc;
}

int kvhost_index_lookup(kvhost_t* h, const char* model, const uint64_t* keys, size_t n, const char* const* pods, size_t n_pods,
                        const char** pod_out, const char** tier_out, uint8_t* cnt_out) {
    if (int rc = need_dev(h)) return rc;
    const auto t0 = std::chrono::steady_clock::now();
    auto secs = [&] { return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); };
    if (n == 0) { observe_lookup(h, secs(), false, nullptr, nullptr, 0); return hfail(KVIDX_EINVAL, "no requestKeys provided for lookup"); }
    std::vector<uint64_t> mask; bool filtered; int mid;
    { std::lock_guard<std::mutex> g(h->mu); filtered = build_filter(h, pods, n_pods, mask); mid = h->models.id(model, 65536); }
    std::vector<kvidx_podtier_t> pt(n * KVIDX_MAX_PODS_PER_KEY);
    if (filtered) { bool any = false; for (uint64_t w : mask) any |= w != 0; if (!any) { memset(cnt_out, 0, n); observe_lookup(h, secs(), true, pt.data(), cnt_out, n); return 0; } }
    const int rc = kvidx_lookup(h->ix, (uint32_t)mid, keys, (int64_t)n, filtered ? mask.data() : nullptr, pt.data(), cnt_out);
    observe_lookup(h, secs(), rc == 0, pt.data(), cnt_out, n);
    if (rc) { g_herr = kvidx_last_error(h->ix); return rc; }
    std::lock_guard<std::mutex> g(h->mu);
    for (size_t i = 0; i < n; ++i)
        for (int j = 0; j < cnt_out[i]; ++j) {
            const kvidx_podtier_t e = pt[i * KVIDX_MAX_PODS_PER_KEY + j];
            pod_out[i * KVIDX_MAX_PODS_PER_KEY + j] = h->pods.names[KVIDX_PT_POD(e)].c_str();
            tier_out[i * KVIDX_MAX_PODS_PER_KEY + j] = h->tiers.names[KVIDX_PT_TIER(e)].c_str();
        }
    return 0;
}

int kvhost_pool_queue_index(kvhost_t* h, const char* pod) { return (int)kvidx_queue_index(pod, strlen(pod), h->cfg.concurrency); }

int kvhost_pool_add_task(kvhost_t* h, const char* pod, const char* model, const void* payload, size_t len) {
    if (!h || !pod || !model) return hfail(KVIDX_EINVAL, "bad arguments");
    std::lock_guard<std::mutex> g(h->mu);
    const int p = h->pods.id(pod, h->pod_cap), m = h->models.id(model, 65536);
    if (p < 0 || m < 0) return hfail(KVIDX_ERANGE, "pod / model id space exhausted");
    h->queues[kvidx_queue_index(pod, strlen(pod), h->cfg.concurrency)].push_back(Msg{(uint32_t)p, (uint32_t)m, std::string((const char*)payload, len)});
    return 0;
}

int64_t kvhost_pool_process(kvhost_t* h, int64_t* n_dropped_out) {
    if (int rc = need_dev(h)) return rc;
    std::vector<kvidx_event_t> evs; std::vector<uint64_t> hashes; std::vector<uint32_t> toks;
    {
        std::lock_guard<std::mutex> g(h->mu);
        for (auto& q : h->queues) {                      // each queue is FIFO; one pod always maps to one queue
            for (auto& m : q) decode_batch(h, m.pod, m.model, (const uint8_t*)m.payload.data(), m.payload.size(), evs, hashes, toks);
            q.clear();
        }
    }
    if (n_dropped_out) *n_dropped_out = 0;
    if (h->cfg.enable_metrics) {                       // digestEvents calls the instrumented Add / Evict (pool.go:299-330)
        const uint32_t bs = h->cfg.index.block_size ? h->cfg.index.block_size : 16;
        uint64_t adm = 0, evi = 0;
        for (const auto& e : evs) {
            if (e.op == KVIDX_EV_BLOCK_STORED) { if (e.n_hashes > 0) adm += e.n_tokens / bs; }      // Add(len(requestKeys)) only if there are engine keys
            else if (e.op == KVIDX_EV_BLOCK_REMOVED) evi += e.n_hashes;                             // one Evict with one entry per hash
        }
        std::lock_guard<std::mutex> g(h->mu);
        h->met.admissions_total += adm; h->met.evictions_total += evi;
    }
    if (evs.empty()) return 0;
    const int rc = kvidx_apply_events(h->ix, evs.data(), (int64_t)evs.size(), hashes.data(), (int64_t)hashes.size(), toks.data(), (int64_t)toks.size(), n_dropped_out);
    if (rc) { g_herr = kvidx_last_error(h->ix); return rc; }
    return (int64_t)evs.size();
}

int64_t kvhost_decode_event_batch(kvhost_t* h, const char* pod, const char* model, const void* payload, size_t len, kvidx_event_t* ev_out,
                                  size_t ev_cap, uint64_t* hash_out, size_t hash_cap, size_t* n_hash_out, uint32_t* tok_out, size_t tok_cap,
                                  size_t* n_tok_out) {
    if (!h || !pod || !model) return hfail(KVIDX_EINVAL, "bad arguments");
    std::vector<kvidx_event_t> evs; std::vector<uint64_t> hashes; std::vector<uint32_t> toks;
    {
        std::lock_guard<std::mutex> g(h->mu);
        const int p = h->pods.id(pod, h->pod_cap), m = h->models.id(model, 65536);
        if (p < 0 || m < 0) return hfail(KVIDX_ERANGE, "pod / model id space exhausted");
        decode_batch(h, (uint32_t)p, (uint32_t)m, (const uint8_t*)payload, len, evs, hashes, toks);
    }
    if (evs.size() > ev_cap || hashes.size() > hash_cap || toks.size() > tok_cap) return hfail(KVIDX_ENOSPC, "output buffers too small");
    if (!evs.empty()) memcpy(ev_out, evs.data(), evs.size() * sizeof(kvidx_event_t));
    if (!hashes.empty()) memcpy(hash_out, hashes.data(), hashes.size() * 8);
    if (!toks.empty()) memcpy(tok_out, toks.data(), toks.size() * 4);
    if (n_hash_out) *n_hash_out = hashes.size();
    if (n_tok_out) *n_tok_out = toks.size();
    return (int64_t)evs.size();
}

}  // extern "C"

// ---- tokenization prefix store (prefixstore/lru_store.go) -------------------------------------------------------------
namespace {
constexpr uint64_t XP1 = 0x9E3779B185EBCA87ull, XP2 = 0xC2B2AE3D27D4EB4Full, XP3 = 0x165667B19E3779F9ull, XP4 = 0x85EBCA77C2B2AE63ull,
                   XP5 = 0x27D4EB2F165667C5ull;
inline uint64_t rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }
inline uint64_t rd64(const uint8_t* p) { uint64_t v; memcpy(&v, p, 8); return v; }      // little-endian hosts only (x86-64, aarch64)
inline uint32_t rd32(const uint8_t* p) { uint32_t v; memcpy(&v, p, 4); return v; }
inline uint64_t xx_round(uint64_t acc, uint64_t in) { return rotl64(acc + in * XP2, 31) * XP1; }
inline uint64_t xx_merge(uint64_t h, uint64_t v) { return (h ^ xx_round(0, v)) * XP1 + XP4; }

// XXH64 as published (cespare/xxhash/v2 in the reference's go.mod); fed in two pieces (the 8-byte chain prefix, then the block)
uint64_t xxh64(const uint8_t* p, size_t n, uint64_t seed) {
    const uint8_t* e = p + n;
    uint64_t h;
    if (n >= 32) {
        uint64_t v1 = seed + XP1 + XP2, v2 = seed + XP2, v3 = seed, v4 = seed - XP1;
        for (; p + 32 <= e; p += 32) { v1 = xx_round(v1, rd64(p)); v2 = xx_round(v2, rd64(p + 8)); v3 = xx_round(v3, rd64(p + 16)); v4 = xx_round(v4, rd64(p + 24)); }
        h = rotl64(v1, 1) + rotl64(v2, 7) + rotl64(v3, 12) + rotl64(v4, 18);
        h = xx_merge(h, v1); h = xx_merge(h, v2); h = xx_merge(h, v3); h = xx_merge(h, v4);
    } else h = seed + XP5;
    h += (uint64_t)n;
    for (; p + 8 <= e; p += 8) h = rotl64(h ^ xx_round(0, rd64(p)), 27) * XP1 + XP4;
    if (p + 4 <= e) { h = rotl64(h ^ ((uint64_t)rd32(p) * XP1), 23) * XP2 + XP3; p += 4; }
    for (; p < e; ++p) h = rotl64(h ^ ((uint64_t)*p * XP5), 11) * XP1;
    h ^= h >> 33; h *= XP2; h ^= h >> 29; h *= XP3; h ^= h >> 32;
    return h;
}
}  // namespace

struct kvhost_prefix_store {
    int64_t cache_size; int32_t block_size;
    std::mutex mu;
    // golang-lru semantics: Add inserts at the front (or updates and refreshes), evicts the back beyond cache_size; Get refreshes
    std::list<std::pair<uint64_t, std::vector<uint32_t>>> order;
    std::unordered_map<uint64_t, std::list<std::pair<uint64_t, std::vector<uint32_t>>>::iterator> map;
    std::vector<uint8_t> scratch;

    uint64_t chain(uint64_t prev, const uint8_t* blk) {
        scratch.resize(8 + (size_t)block_size);
        memcpy(scratch.data(), &prev, 8);
        memcpy(scratch.data() + 8, blk, (size_t)block_size);
        return xxh64(scratch.data(), scratch.size(), 0);
    }
};

extern "C" {

uint64_t kvhost_xxhash64(const void* data, size_t len, uint64_t seed) { return xxh64((const uint8_t*)data, len, seed); }

int kvhost_prefix_store_create(int64_t cache_size, int32_t block_size, kvhost_prefix_store_t** out) {
    if (!out) return hfail(KVIDX_EINVAL, "bad arguments");
    if (cache_size <= 0) return hfail(KVIDX_EINVAL, "failed to initialize in-memory index: must provide a positive size");     // lru.New
    if (block_size <= 0) return hfail(KVIDX_EINVAL, "block size must be positive");
    auto* s = new kvhost_prefix_store();
    s->cache_size = cache_size; s->block_size = block_size;
    *out = s;
    return 0;
}

void kvhost_prefix_store_destroy(kvhost_prefix_store_t* s) { delete s; }

int64_t kvhost_prefix_store_len(kvhost_prefix_store_t* s) { if (!s) return KVIDX_EINVAL; std::lock_guard<std::mutex> g(s->mu); return (int64_t)s->map.size(); }

int kvhost_prefix_store_add(kvhost_prefix_store_t* s, const char* prompt, size_t prompt_len, const uint32_t* tokens, const uint64_t* offsets, size_t n_tokens) {
    if (!s || (prompt_len && !prompt) || (n_tokens && (!tokens || !offsets))) return hfail(KVIDX_EINVAL, "bad arguments");
    if (prompt_len == 0 || n_tokens == 0) return 0;                                  // lru_store.go:96-98
    std::lock_guard<std::mutex> g(s->mu);
    const size_t bs = (size_t)s->block_size;
    size_t it = 0;
    uint64_t prev = 0;
    for (size_t start = 0; start + bs <= prompt_len; start += bs) {                  // no partial blocks
        const size_t end = start + bs;
        prev = s->chain(prev, (const uint8_t*)prompt + start);
        std::vector<uint32_t> blk;
        for (; it < n_tokens && offsets[2 * it + 1] <= end; ++it) blk.push_back(tokens[it]);      // a token goes with the block its END falls in
        auto f = s->map.find(prev);
        if (f != s->map.end()) { f->second->second = std::move(blk); s->order.splice(s->order.begin(), s->order, f->second); }
        else {
            s->order.emplace_front(prev, std::move(blk));
            s->map[prev] = s->order.begin();
            if ((int64_t)s->map.size() > s->cache_size) { s->map.erase(s->order.back().first); s->order.pop_back(); }
        }
    }
    return 0;
}

int64_t kvhost_prefix_store_find(kvhost_prefix_store_t* s, const char* prompt, size_t prompt_len, uint32_t* tokens_out, size_t cap, double* ratio_out) {
    if (!s || (prompt_len && !prompt)) return hfail(KVIDX_EINVAL, "bad arguments");
    std::lock_guard<std::mutex> g(s->mu);
    const size_t bs = (size_t)s->block_size;
    uint64_t prev = 0;
    double ratio = 0.0;
    int64_t n = 0;
    for (size_t i = 0; i + bs <= prompt_len; i += bs) {
        prev = s->chain(prev, (const uint8_t*)prompt + i);
        auto f = s->map.find(prev);
        if (f == s->map.end()) break;                                                // early stop
        s->order.splice(s->order.begin(), s->order, f->second);                      // Get refreshes recency
        for (uint32_t t : f->second->second) { if (tokens_out && (size_t)n < cap) tokens_out[n] = t; ++n; }
        ratio = (double)(i + bs) / (double)prompt_len;
    }
    if (ratio_out) *ratio_out = ratio;
    return n;
}

}  // extern "C"
