// xl_core.cu -- state shared by the translation units of libxlprop.so: thread-local error text, launch counter, per-kernel
// event profiling, per-device twiddle table and SM count.  All of it is safe to call from several host threads and with
// several devices (include/xlprop.h: the library is re-entrant; work is ordered by the caller's streams only).
#include "xl_common.h"
#include <mutex>
#include <vector>
#include <math.h>

#ifdef XL_HOST_EMU
thread_local xl_dim3 xl_emu_blockIdx;
thread_local xl_dim3 xl_emu_gridDim;
#endif

// ------------------------------------------------------------------------------------------------ errors
static thread_local char g_err[512] = "";
int xl_fail(int code, const char* fmt, const char* a, long long b) {
    snprintf(g_err, sizeof(g_err), fmt, a, b);
    return code;
}
extern "C" int xl_version(void) { return XLPROP_VERSION; }
extern "C" const char* xl_last_error(void) { return g_err; }

// ------------------------------------------------------------------------------------------------ instrumentation
static std::atomic<long long> g_launches{0};
static std::atomic<int> g_prof_on{0};
static std::mutex g_prof_mutex;
void xl_count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
static bool xl_prof_active() { return g_prof_on.load(std::memory_order_relaxed) != 0; }
extern "C" long long xl_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }
#ifndef XL_HOST_EMU
struct XlProfRec { const char* name; cudaEvent_t e0, e1; };
static std::vector<XlProfRec*> g_prof;
void* xl_prof_begin(const char* name, xl_stream_t stream) {
    if (!xl_prof_active()) return 0;
    XlProfRec* r = new XlProfRec;
    r->name = name;
    cudaEventCreate(&r->e0); cudaEventCreate(&r->e1);
    cudaEventRecord(r->e0, stream);
    return r;
}
void xl_prof_end(void* rec, xl_stream_t stream) {
    XlProfRec* r = (XlProfRec*)rec;
    cudaEventRecord(r->e1, stream);
    std::lock_guard<std::mutex> lk(g_prof_mutex);
    g_prof.push_back(r);
}
#else
void* xl_prof_begin(const char*, xl_stream_t) { return 0; }
void xl_prof_end(void*, xl_stream_t) {}
#endif
extern "C" void xl_prof_enable(int on) {
#ifndef XL_HOST_EMU
    if (on) {
        std::lock_guard<std::mutex> lk(g_prof_mutex);
        for (auto* r : g_prof) { cudaEventDestroy(r->e0); cudaEventDestroy(r->e1); delete r; }
        g_prof.clear();
    }
#endif
    g_prof_on.store(on, std::memory_order_relaxed);
}
// Writes lines "name count total_ms\n" into buf (after synchronising the recorded events); returns bytes written.
extern "C" int xl_prof_report(char* buf, int cap) {
    int n = 0;
    if (cap > 0) buf[0] = 0;
#ifndef XL_HOST_EMU
    struct Acc { const char* name; int count; double ms; };
    std::vector<Acc> acc;
    std::lock_guard<std::mutex> lk(g_prof_mutex);
    for (auto* r : g_prof) {
        cudaEventSynchronize(r->e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, r->e0, r->e1);
        bool found = false;
        for (auto& a : acc) if (a.name == r->name || !strcmp(a.name, r->name)) { a.count++; a.ms += ms; found = true; break; }
        if (!found) acc.push_back(Acc{r->name, 1, (double)ms});
    }
    for (auto& a : acc) {
        int w = snprintf(buf + n, cap > n ? cap - n : 0, "%s %d %.6f\n", a.name, a.count, a.ms);
        if (w < 0 || n + w >= cap) break;
        n += w;
    }
#endif
    return n;
}

// ------------------------------------------------------------------------------------------------ per-device tables
static std::mutex g_dev_mutex;
static float2* g_tw[64] = {0};
static int g_sms[64] = {0};
#define XL_TWN_HOST 32768     // == XL_TWN (xl_fft.cuh): tw[k] = exp(-2*pi*i*k/XL_TWN), generated in fp64
const float2* xl_twiddles() {
    int dev = 0;
#ifndef XL_HOST_EMU
    cudaGetDevice(&dev);
#endif
    if (dev < 0 || dev >= 64) return 0;
    std::lock_guard<std::mutex> lk(g_dev_mutex);
    if (g_tw[dev]) return g_tw[dev];
    std::vector<float2> h(XL_TWN_HOST);
    for (int k = 0; k < XL_TWN_HOST; ++k) {
        double a = 2.0 * M_PI * (double)k / (double)XL_TWN_HOST;
        h[k].x = (float)cos(a);
        h[k].y = (float)(-sin(a));
    }
#ifdef XL_HOST_EMU
    g_tw[dev] = (float2*)malloc(sizeof(float2) * XL_TWN_HOST);
    memcpy(g_tw[dev], h.data(), sizeof(float2) * XL_TWN_HOST);
#else
    float2* d = 0;
    if (cudaMalloc(&d, sizeof(float2) * XL_TWN_HOST) != cudaSuccess) return 0;
    if (cudaMemcpy(d, h.data(), sizeof(float2) * XL_TWN_HOST, cudaMemcpyHostToDevice) != cudaSuccess) return 0;
    g_tw[dev] = d;
#endif
    return g_tw[dev];
}
int xl_sm_count() {
#ifdef XL_HOST_EMU
    return 1;
#else
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) return 148;
    std::lock_guard<std::mutex> lk(g_dev_mutex);
    if (!g_sms[dev]) {
        int sms = 0;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        g_sms[dev] = sms > 0 ? sms : 148;
    }
    return g_sms[dev];
#endif
}

int zero_async(void* p, size_t bytes, xl_stream_t s) {
#ifdef XL_HOST_EMU
    (void)s; memset(p, 0, bytes); return XL_OK;
#else
    cudaError_t e = cudaMemsetAsync(p, 0, bytes, s);
    return e == cudaSuccess ? XL_OK : xl_fail(XL_E_CUDA, "cudaMemsetAsync: %s", cudaGetErrorString(e));
#endif
}

extern "C" int xl_rs_padded_length(int N) {
    if (N < 2) return 0;
    int L = next_pow2(2 * N - 1);
    if (L < 32) L = 32;
    return L <= 4096 ? L : 0;
}
extern "C" int xl_czt_padded_length(int m, int M) {
    if (m < 1 || M < 2) return 0;
    int mp = m + M - 1;
    int L = next_pow2(mp);
    if (L == mp) return 0;  // the reference slices b[m:mp+1] out of np2 == mp rows and raises; out of contract
    if (L < 32) L = 32;
    return L <= 4096 ? L : 0;
}
