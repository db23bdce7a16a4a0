// Library-wide entry points: version, thread-local error string, launch counter.
#include <atomic>
#include <map>
#include <mutex>
#include <stdarg.h>
#include <string>
#include <vector>
#include "common.cuh"

namespace snb {
static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

struct TimingRec { const char* name; cudaEvent_t a, b; };
static std::atomic<bool> g_timing{false};
static std::mutex g_timing_mu;
static std::vector<TimingRec> g_timing_recs;
static std::vector<cudaEvent_t> g_event_pool;   // events are recycled: creating them costs more than the kernels being timed
static bool pool_get(cudaEvent_t* e) {
    if (!g_event_pool.empty()) {
        *e = g_event_pool.back();
        g_event_pool.pop_back();
        return true;
    }
    return cudaEventCreate(e) == cudaSuccess;
}
void timing_begin(const char* name, cudaStream_t stream) {
    if (!g_timing.load(std::memory_order_relaxed)) return;
    std::lock_guard<std::mutex> lk(g_timing_mu);
    TimingRec r;
    r.name = name;
    if (!pool_get(&r.a) || !pool_get(&r.b)) return;
    cudaEventRecord(r.a, stream);
    g_timing_recs.push_back(r);
}
void timing_end(cudaStream_t stream) {
    if (!g_timing.load(std::memory_order_relaxed)) return;
    std::lock_guard<std::mutex> lk(g_timing_mu);
    if (!g_timing_recs.empty()) cudaEventRecord(g_timing_recs.back().b, stream);
}
}  // namespace snb

extern "C" {
int sn_version(void) { return SN_VERSION; }
const char* sn_last_error_string(void) { return snb::g_err; }
long long sn_launch_count(void) { return snb::g_launches.load(); }
void sn_reset_launch_count(void) { snb::g_launches.store(0); }
void sn_timing_enable(int on) { snb::g_timing.store(on != 0); }
/* "name count total_ms\n" per kernel for every launch recorded since the last report; the caller synchronises the device first */
int sn_timing_report(char* buf, int cap) {
    std::lock_guard<std::mutex> lk(snb::g_timing_mu);
    std::map<std::string, std::pair<int, double>> agg;
    for (auto& r : snb::g_timing_recs) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) {
            auto& e = agg[r.name];
            e.first += 1;
            e.second += ms;
        }
        snb::g_event_pool.push_back(r.a);
        snb::g_event_pool.push_back(r.b);
    }
    snb::g_timing_recs.clear();
    std::string out;
    for (auto& kv : agg) out += kv.first + " " + std::to_string(kv.second.first) + " " + std::to_string(kv.second.second) + "\n";
    if (buf != nullptr && cap > 0) {
        snprintf(buf, (size_t)cap, "%s", out.c_str());
    }
    return (int)out.size();
}
}
