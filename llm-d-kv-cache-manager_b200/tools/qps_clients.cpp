// qps_clients.cpp -- N OS threads, ONE prompt per call, through the C ABI: the load shape of the reference's gRPC server
// (one goroutine per RPC, no batching anywhere, examples/kv_cache_index_service/server/server.go:70-96).  Reports calls/s and
// latency percentiles; the library's submission queue is what turns the concurrent callers into shared launches.
//   usage: kvidx_qps <threads> <seconds> [documents] [tokens_per_prompt]       (prints one JSON line)
// Every returned score is checked against what the index must hold (pod p of a document caches its first depth_p blocks).
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>

#include "../../include/kvidx.h"

static uint64_t mix(uint64_t z) { z ^= z >> 30; z *= 0xBF58476D1CE4E5B9ull; z ^= z >> 27; z *= 0x94D049BB133111EBull; return z ^ (z >> 31); }

int main(int argc, char** argv) {
    const int threads = argc > 1 ? atoi(argv[1]) : 1000;
    const double seconds = argc > 2 ? atof(argv[2]) : 2.0;
    const int D = argc > 3 ? atoi(argv[3]) : 4096;
    const int T = argc > 4 ? atoi(argv[4]) : 4096;
    const int B = 16, n = T / B, P = 64;
    kvidx_config_t cfg; kvidx_config_default(&cfg);
    cfg.capacity = (uint64_t)D * n + (1 << 16); cfg.max_pods = P;
    kvidx_t* ix = nullptr;
    if (kvidx_create(&cfg, &ix)) { fprintf(stderr, "kvidx_create: %s\n", kvidx_last_error(nullptr)); return 2; }
    // documents: tokens from a counter-based generator; document d is cached on pods d%P (all blocks) and (d+7)%P (half)
    std::vector<uint32_t> tok((size_t)D * T);
    for (size_t i = 0; i < tok.size(); ++i) tok[i] = (uint32_t)(mix(i * 0x9E3779B97F4A7C15ull + 12345) % 128256);
    {
        std::vector<int64_t> off(D + 1), koff(D + 1);
        for (int d = 0; d <= D; ++d) off[d] = (int64_t)d * T;
        std::vector<uint64_t> keys((size_t)D * n), eng((size_t)D * n);
        if (kvidx_hash_keys(ix, tok.data(), off.data(), D, nullptr, nullptr, keys.data(), koff.data())) { fprintf(stderr, "hash_keys: %s\n", kvidx_last_error(ix)); return 2; }
        for (size_t i = 0; i < keys.size(); ++i) eng[i] = ~keys[i];
        for (int d = 0; d < D; ++d) {
            kvidx_podtier_t a = KVIDX_PODTIER(d % P, 0), b = KVIDX_PODTIER((d + 7) % P, 1);
            if (kvidx_add(ix, 0, eng.data() + (size_t)d * n, keys.data() + (size_t)d * n, n, &a, 1)) { fprintf(stderr, "add: %s\n", kvidx_last_error(ix)); return 2; }
            if (kvidx_add(ix, 0, eng.data() + (size_t)d * n, keys.data() + (size_t)d * n, n / 2, &b, 1)) { fprintf(stderr, "add: %s\n", kvidx_last_error(ix)); return 2; }
        }
    }
    std::atomic<bool> go{false}, stop{false};
    std::atomic<long long> errors{0};
    std::vector<std::vector<float>> lat(threads);
    std::vector<std::thread> th;
    for (int t = 0; t < threads; ++t) th.emplace_back([&, t] {
        std::vector<uint32_t> q(T);
        uint64_t s = 0xABCDEFull * (t + 1);
        lat[t].reserve(1 << 14);
        while (!go.load()) std::this_thread::yield();
        while (!stop.load()) {
            s = mix(s + 0x9E3779B97F4A7C15ull);
            const int d = (int)(s % D), m = (int)((s >> 32) % (n + 1));           // first m blocks of document d, then fresh tokens
            for (int i = 0; i < m * B; ++i) q[i] = tok[(size_t)d * T + i];
            for (int i = m * B; i < T; ++i) q[i] = (uint32_t)(mix(s + i) % 128256);
            const int64_t off[2] = {0, T};
            uint16_t pods[10]; double sc[10]; uint8_t cnt = 0, has = 0;
            const auto t0 = std::chrono::steady_clock::now();
            const int rc = kvidx_score_batch_sparse(ix, q.data(), off, 1, nullptr, 0, nullptr, pods, sc, &cnt, &has);
            const auto t1 = std::chrono::steady_clock::now();
            lat[t].push_back(std::chrono::duration<float, std::micro>(t1 - t0).count());
            // expected: pod d%P holds all blocks -> score m; pod (d+7)%P (tier cpu 0.8) min(m, n/2) adds of 0.8
            bool ok = rc == 0 && has == 1 && (m == 0 ? cnt == 0 : cnt == 2);
            if (ok && m > 0) {
                double e2 = 0.0; for (int i = 0; i < std::min(m, n / 2); ++i) e2 = i ? e2 + 0.8 : 0.8;
                for (int j = 0; j < 2; ++j) {
                    if (pods[j] == d % P) ok = ok && sc[j] == (double)m;
                    else if (pods[j] == (d + 7) % P) ok = ok && sc[j] == e2;
                    else ok = false;
                }
            }
            if (!ok) errors++;
        }
    });
    go.store(true);
    const auto t0 = std::chrono::steady_clock::now();
    std::this_thread::sleep_for(std::chrono::duration<double>(seconds));
    stop.store(true);
    for (auto& x : th) x.join();
    const double el = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    std::vector<float> all;
    for (auto& v : lat) all.insert(all.end(), v.begin(), v.end());
    std::sort(all.begin(), all.end());
    kvidx_stats_t st{}; kvidx_get_stats(ix, &st);
    auto pct = [&](double p) { return all.empty() ? 0.0 : (double)all[std::min(all.size() - 1, (size_t)(p * all.size()))] / 1e3; };
    printf("{\"threads\": %d, \"seconds\": %.3f, \"calls\": %zu, \"calls_per_s\": %.1f, \"p50_ms\": %.4f, \"p99_ms\": %.4f, \"max_ms\": %.4f, "
           "\"coalesced_calls\": %llu, \"kernel_launches\": %llu, \"wrong_results\": %lld, \"prompt_tokens\": %d, \"index_blocks\": %lld}\n",
           threads, el, all.size(), all.size() / el, pct(0.5), pct(0.99), all.empty() ? 0.0 : all.back() / 1e3,
           (unsigned long long)st.coalesced_calls, (unsigned long long)st.kernel_launches, errors.load(), T, (long long)D * n);
    kvidx_destroy(ix);
    return errors.load() ? 1 : 0;
}
