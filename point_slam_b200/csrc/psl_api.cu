// error string, version, device helpers
#include <stdarg.h>
#include <stdlib.h>

#include "psl_common.cuh"

namespace psl {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
int sm_count() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (!cached[dev]) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

static bool g_timing = false;
static unsigned long long g_launches = 0;
struct EvPair { int id; cudaEvent_t a, b; };
static EvPair* g_ev = nullptr;
static int g_ev_cap = 0, g_ev_n = 0;
TimingScope::TimingScope(int id, cudaStream_t s, int n_kernels) : slot(-1), st(s) {
    g_launches += (unsigned long long)n_kernels;
    if (!g_timing) return;
    if (g_ev_n >= g_ev_cap) {
        const int ncap = g_ev_cap ? g_ev_cap * 2 : 1024;
        EvPair* n = static_cast<EvPair*>(realloc(g_ev, sizeof(EvPair) * ncap));
        if (!n) return;
        for (int i = g_ev_cap; i < ncap; ++i) { cudaEventCreate(&n[i].a); cudaEventCreate(&n[i].b); }
        g_ev = n; g_ev_cap = ncap;
    }
    slot = g_ev_n++;
    g_ev[slot].id = id;
    cudaEventRecord(g_ev[slot].a, st);
}
TimingScope::~TimingScope() {
    if (slot >= 0) cudaEventRecord(g_ev[slot].b, st);
}
}  // namespace psl

extern "C" unsigned long long psl_launch_count(void) { return psl::g_launches; }
extern "C" int psl_timing_enable(int on) {
    psl::g_timing = on != 0;
    psl::g_ev_n = 0;
    return 0;
}
extern "C" int psl_timing_collect(float* ms_out, int* count_out) {
    using namespace psl;
    for (int i = 0; i < T_COUNT; ++i) { ms_out[i] = 0.f; count_out[i] = 0; }
    PSL_CHECK_CUDA(cudaDeviceSynchronize());
    for (int i = 0; i < g_ev_n; ++i) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, g_ev[i].a, g_ev[i].b) == cudaSuccess) { ms_out[g_ev[i].id] += ms; count_out[g_ev[i].id] += 1; }
    }
    g_ev_n = 0;
    return 0;
}

extern "C" int psl_version(void) { return 100; }
extern "C" const char* psl_last_error(void) { return psl::g_err; }
extern "C" int psl_device_sm_count(void) { return psl::sm_count(); }
