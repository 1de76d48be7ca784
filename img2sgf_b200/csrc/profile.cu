// profile.cu -- section timers and launch counter (see profile.cuh)
#include <atomic>
#include <mutex>
#include <vector>
#include "common.cuh"
#include "profile.cuh"

namespace i2s {

static const char *kNames[SEC_COUNT] = {
    "grey", "enhance", "sobel_nms_rgb", "sobel_nms", "hysteresis", "state_to_edges", "gauss357", "median", "edge_list", "vote",
    "radius", "circles_finish", "stack", "mask", "line_vote", "line_peaks", "cluster", "validate", "classify"};

struct Pair { cudaEvent_t a, b; int id; };
static std::mutex g_mu;
static bool g_on = false;
static std::vector<Pair> g_live, g_free;
static std::atomic<long long> g_launches{0};

void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

ScopedSection::ScopedSection(int id_, cudaStream_t st_) : id(id_), st(st_), stop(nullptr)
{
    if (!g_on) return;
    std::lock_guard<std::mutex> lk(g_mu);
    Pair p;
    if (!g_free.empty()) { p = g_free.back(); g_free.pop_back(); }
    else { cudaEventCreate(&p.a); cudaEventCreate(&p.b); }
    p.id = id;
    cudaEventRecord(p.a, st);
    stop = (void *)p.b;
    g_live.push_back(p);
}

ScopedSection::~ScopedSection()
{
    if (stop) cudaEventRecord((cudaEvent_t)stop, st);
}

}  // namespace i2s

using namespace i2s;

extern "C" int i2s_profile_enable(int on)
{
    std::lock_guard<std::mutex> lk(g_mu);
    g_on = on != 0;
    return SEC_COUNT;
}

extern "C" const char *i2s_profile_section_name(int id) { return id >= 0 && id < SEC_COUNT ? kNames[id] : ""; }

// Waits for every recorded section, adds its elapsed time (ms) and count per section, recycles events.
extern "C" int i2s_profile_read(double *ms, long long *counts, int nsections)
{
    std::lock_guard<std::mutex> lk(g_mu);
    for (int i = 0; i < nsections; i++) { ms[i] = 0; counts[i] = 0; }
    for (Pair &p : g_live) {
        float t = 0;
        if (cudaEventSynchronize(p.b) != cudaSuccess || cudaEventElapsedTime(&t, p.a, p.b) != cudaSuccess) {
            set_error("i2s_profile_read: event error");
            return I2S_E_CUDA;
        }
        if (p.id < nsections) { ms[p.id] += t; counts[p.id]++; }
        g_free.push_back(p);
    }
    g_live.clear();
    return I2S_OK;
}

// Number of kernels of this library launched since the last call with reset != 0.
extern "C" long long i2s_launch_count(int reset)
{
    long long v = g_launches.load();
    if (reset) g_launches.store(0);
    return v;
}
