// prof.cu -- per-kernel timing and launch accounting for libfqss_sm100.
//
// Every kernel launch site of the library sits inside a ProfScope.  The scope always adds its kernel
// count to a process-wide launch counter (bench.py's `gpu_launches`); when collection is enabled
// (fqss_prof_enable(1)) it also brackets the launch with a pair of CUDA events recorded ON THE
// LAUNCHING STREAM, so a whole QAT step can be attributed kernel class by kernel class without a
// profiler attached (bench.py's `roofline` object and profiles/step_breakdown_*.txt come from this).
// Event pairs are resolved lazily in fqss_prof_read (after a device synchronise).
#include <string.h>

#include <atomic>
#include <mutex>
#include <vector>

#include "fqss_common.cuh"

namespace fqss {

namespace {

constexpr int MAX_SLOTS = 96;

struct Slot {
    char name[48];
    double ms;
    int64_t count;      // scopes
    int64_t kernels;    // kernels launched inside those scopes
};

struct Pending {
    int slot;
    cudaEvent_t e0, e1;
};

std::mutex g_mu;
Slot g_slots[MAX_SLOTS];
int g_nslots = 0;
std::vector<Pending> g_pending;
std::vector<cudaEvent_t> g_pool;
std::atomic<int> g_enabled{0};
std::atomic<long long> g_launches{0};

int slot_of(const char* name) {
    for (int i = 0; i < g_nslots; ++i)
        if (strcmp(g_slots[i].name, name) == 0) return i;
    if (g_nslots == MAX_SLOTS) return MAX_SLOTS - 1;
    Slot& s = g_slots[g_nslots];
    strncpy(s.name, name, sizeof(s.name) - 1);
    s.name[sizeof(s.name) - 1] = 0;
    s.ms = 0.0;
    s.count = 0;
    s.kernels = 0;
    return g_nslots++;
}

cudaEvent_t take_event() {
    if (!g_pool.empty()) {
        cudaEvent_t e = g_pool.back();
        g_pool.pop_back();
        return e;
    }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
}

void resolve_locked() {
    if (g_pending.empty()) return;
    cudaDeviceSynchronize();
    for (const Pending& p : g_pending) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, p.e0, p.e1) == cudaSuccess) g_slots[p.slot].ms += (double)ms;
        g_pool.push_back(p.e0);
        g_pool.push_back(p.e1);
    }
    g_pending.clear();
    cudaGetLastError();
}

}  // namespace

ProfScope::ProfScope(const char* name, cudaStream_t s, int nkernels) : stream_(s), e1_(nullptr), slot_(-1) {
    g_launches.fetch_add(nkernels, std::memory_order_relaxed);
    if (!g_enabled.load(std::memory_order_relaxed)) return;
    std::lock_guard<std::mutex> lk(g_mu);
    slot_ = slot_of(name);
    g_slots[slot_].count += 1;
    g_slots[slot_].kernels += nkernels;
    cudaEvent_t e0 = take_event();
    e1_ = take_event();
    cudaEventRecord(e0, stream_);
    g_pending.push_back(Pending{slot_, e0, (cudaEvent_t)e1_});
}

ProfScope::~ProfScope() {
    if (slot_ >= 0) cudaEventRecord((cudaEvent_t)e1_, stream_);
}

}  // namespace fqss

using namespace fqss;

extern "C" {

int fqss_prof_enable(int on) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (!on) resolve_locked();
    g_enabled.store(on ? 1 : 0);
    return 0;
}

int fqss_prof_reset(void) {
    std::lock_guard<std::mutex> lk(g_mu);
    resolve_locked();
    g_nslots = 0;
    return 0;
}

int fqss_prof_nslots(void) {
    std::lock_guard<std::mutex> lk(g_mu);
    resolve_locked();
    return g_nslots;
}

int fqss_prof_read(int slot, char* name, int name_cap, double* total_ms, int64_t* scopes, int64_t* kernels) {
    std::lock_guard<std::mutex> lk(g_mu);
    resolve_locked();
    FQSS_REQUIRE(slot >= 0 && slot < g_nslots && name && name_cap > 0 && total_ms && scopes && kernels, -1, "prof_read: bad argument");
    strncpy(name, g_slots[slot].name, (size_t)name_cap - 1);
    name[name_cap - 1] = 0;
    *total_ms = g_slots[slot].ms;
    *scopes = g_slots[slot].count;
    *kernels = g_slots[slot].kernels;
    return 0;
}

int64_t fqss_launch_count(void) { return (int64_t)g_launches.load(); }

}  // extern "C"
