// Instrumentation for bench.py: counts this library's kernel launches and, on request, brackets
// every tcgen05 GEMM launch with CUDA events on the launching stream so the kernel's time share
// inside a real step can be reported (roofline.achieved).  Off by default: no events, no syncs.
#include <atomic>
#include <mutex>
#include <vector>

#include "common.cuh"
#include "prof.cuh"

namespace clipppo {

namespace {
std::atomic<long long> g_launches{0};
std::atomic<int> g_timing{0};
std::mutex g_mu;
struct Span { cudaEvent_t a, b; double flops; long long tag; };
struct Bucket { long long tag; double ms, flops; long long n; };
std::vector<Bucket> g_buckets;   // per-shape totals of the last prof_end
std::vector<Span> g_spans;
std::vector<cudaEvent_t> g_pool;

cudaEvent_t get_event() {
    if (!g_pool.empty()) { cudaEvent_t e = g_pool.back(); g_pool.pop_back(); return e; }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
}
}  // namespace

void prof_count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
bool prof_timing_enabled() { return g_timing.load(std::memory_order_relaxed) != 0; }

void prof_span_begin(cudaStream_t s, double flops, long long tag, void** token) {
    std::lock_guard<std::mutex> lk(g_mu);
    Span sp{get_event(), get_event(), flops, tag};
    cudaEventRecord(sp.a, s);
    g_spans.push_back(sp);
    *token = reinterpret_cast<void*>(g_spans.size());
}
void prof_span_end(cudaStream_t s, void* token) {
    std::lock_guard<std::mutex> lk(g_mu);
    const size_t i = reinterpret_cast<size_t>(token) - 1;
    if (i < g_spans.size()) cudaEventRecord(g_spans[i].b, s);
}

}  // namespace clipppo

using namespace clipppo;

extern "C" int clipppo_prof_begin(int time_gemms) {
    std::lock_guard<std::mutex> lk(g_mu);
    for (auto& sp : g_spans) { g_pool.push_back(sp.a); g_pool.push_back(sp.b); }
    g_spans.clear();
    g_launches.store(0);
    g_timing.store(time_gemms ? 1 : 0);
    return CLIPPPO_OK;
}

extern "C" int clipppo_prof_end(long long* launches, double* gemm_ms, double* gemm_flops, long long* gemm_launches) {
    std::lock_guard<std::mutex> lk(g_mu);
    g_timing.store(0);
    double ms = 0.0, fl = 0.0;
    g_buckets.clear();
    for (auto& sp : g_spans) {
        CLIPPPO_CUDA_TRY(cudaEventSynchronize(sp.b));
        float t = 0.f;
        CLIPPPO_CUDA_TRY(cudaEventElapsedTime(&t, sp.a, sp.b));
        ms += t;
        fl += sp.flops;
        size_t bi = 0;
        while (bi < g_buckets.size() && g_buckets[bi].tag != sp.tag) ++bi;
        if (bi == g_buckets.size()) g_buckets.push_back(Bucket{sp.tag, 0.0, 0.0, 0});
        g_buckets[bi].ms += t; g_buckets[bi].flops += sp.flops; g_buckets[bi].n += 1;
    }
    if (launches) *launches = g_launches.load();
    if (gemm_ms) *gemm_ms = ms;
    if (gemm_flops) *gemm_flops = fl;
    if (gemm_launches) *gemm_launches = static_cast<long long>(g_spans.size());
    for (auto& sp : g_spans) { g_pool.push_back(sp.a); g_pool.push_back(sp.b); }
    g_spans.clear();
    return CLIPPPO_OK;
}

extern "C" int clipppo_prof_bucket(int index, long long* tag, double* ms, double* flops, long long* launches) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (index < 0 || static_cast<size_t>(index) >= g_buckets.size()) return CLIPPPO_ERR_BAD_SHAPE;
    const Bucket& b = g_buckets[index];
    if (tag) *tag = b.tag;
    if (ms) *ms = b.ms;
    if (flops) *flops = b.flops;
    if (launches) *launches = b.n;
    return CLIPPPO_OK;
}
