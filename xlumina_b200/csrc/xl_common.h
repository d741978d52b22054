// xl_common.h -- host-side plumbing shared by the translation units of libxlprop.so (xl_core.cu, xl_rs.cu, xl_slab.cu,
// xl_czt.cu): error reporting, the launch wrapper with its instrumentation, dispatch on the padded length, workspace
// carving.  One translation unit per kernel family keeps the build reproducible (no -split-compile partitioning) and
// parallel.  With -DXL_HOST_EMU (g++, tests only) xl_api.cu includes all of them as one unit.
#pragma once
#include "xl_platform.h"
#include "../../include/xlprop.h"
#include <stdio.h>
#include <string.h>
#include <stdlib.h>
#include <atomic>

#ifdef XL_HOST_EMU
typedef void* xl_stream_t;
#include <vector>
#else
typedef cudaStream_t xl_stream_t;
#endif

// ------------------------------------------------------------------------------------------------ errors (xl_core.cu)
int xl_fail(int code, const char* fmt, const char* a = "", long long b = 0);

// ------------------------------------------------------------------------------------------------ launch
struct XlDim { int x, y; };

// Instrumentation for bench.py (xl_core.cu): a launch counter (always on, atomic) and optional per-kernel CUDA-event timing
// recorded on the launching stream (xl_prof_enable(1); ...; xl_prof_report()); the records are guarded by a mutex.
void xl_count_launch();
void* xl_prof_begin(const char* name, xl_stream_t stream);    // returns a record handle (null when profiling is off)
void xl_prof_end(void* rec, xl_stream_t stream);
int xl_sm_count();                                            // SMs of the current device (cached per device)
const float2* xl_twiddles();                                  // per-device master twiddle table

#ifndef XL_HOST_EMU
// register budget: at least 512/NT CTAs per SM (128 registers per thread), so two L=4096 CTAs overlap their phases on an SM
template <class Body, class = void> struct XlMinBlocks { static constexpr int value = (512 / Body::NT) > 16 ? 16 : (512 / Body::NT); };
template <class Body> struct XlMinBlocks<Body, decltype((void)Body::MINB)> { static constexpr int value = Body::MINB; };   // per-kernel override
template <class Body> __global__ void __launch_bounds__(Body::NT, XlMinBlocks<Body>::value) xl_kernel(const typename Body::Params p) {
    extern __shared__ float4 xl_smem[];
    Body::run(p, (float2*)xl_smem);
}
#endif

template <class Body> static int xl_launch(XlDim grid, xl_stream_t stream, const typename Body::Params& p) {
    if (grid.x <= 0 || grid.y <= 0) return XL_OK;
    const size_t smem = Body::smem();
    xl_count_launch();
#ifdef XL_HOST_EMU
    (void)stream;
    std::vector<char> buf(smem + 64);
    xl_emu_gridDim.x = grid.x; xl_emu_gridDim.y = grid.y; xl_emu_gridDim.z = 1;
    for (int by = 0; by < grid.y; ++by)
        for (int bx = 0; bx < grid.x; ++bx) {
            xl_emu_blockIdx.x = bx; xl_emu_blockIdx.y = by; xl_emu_blockIdx.z = 0;
            Body::run(p, (float2*)buf.data());
        }
    return XL_OK;
#else
    static std::atomic<unsigned long long> attr_mask{0};      // devices on which this kernel's shared-memory limit is set
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 64 && !((attr_mask.load(std::memory_order_acquire) >> dev) & 1ull)) {
        cudaError_t e = cudaFuncSetAttribute(xl_kernel<Body>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return xl_fail(XL_E_CUDA, "cudaFuncSetAttribute: %s (smem %lld)", cudaGetErrorString(e), (long long)smem);
        attr_mask.fetch_or(1ull << dev, std::memory_order_release);
    }
    void* rec = xl_prof_begin(Body::name(), stream);
    xl_kernel<Body><<<dim3(grid.x, grid.y, 1), dim3(Body::NT, 1, 1), smem, stream>>>(p);
    if (rec) xl_prof_end(rec, stream);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return xl_fail(XL_E_CUDA, "kernel launch: %s", cudaGetErrorString(e));
    return XL_OK;
#endif
}

// Cluster kernels: the grid.x CTAs of one blockIdx.y form a thread-block cluster (grid.x <= 8).  Body::run1 is everything
// before the cluster barrier (it leaves its result in the CTA's shared memory), Body::run2 reads the peers' shared memory
// (xl_peer_ld4); a second cluster barrier keeps every CTA's shared memory alive until its peers have finished reading.
#ifndef XL_HOST_EMU
template <class Body> __global__ void __launch_bounds__(Body::NT, XlMinBlocks<Body>::value) xl_kernel_cluster(const typename Body::Params p) {
    extern __shared__ float4 xl_smem[];
    Body::run1(p, (float2*)xl_smem);
    xl_cluster_sync();
    const XlPeers pr{(unsigned)__cvta_generic_to_shared(xl_smem)};
    Body::run2(p, (float2*)xl_smem, pr);
    xl_cluster_sync();
}
#endif
template <class Body> static int xl_launch_cluster(XlDim grid, xl_stream_t stream, const typename Body::Params& p) {
    if (grid.x <= 0 || grid.y <= 0) return XL_OK;
    if (grid.x > 8) return xl_fail(XL_E_UNSUPPORTED, "cluster of %s%lld CTAs (portable limit 8)", "", (long long)grid.x);
    const size_t smem = Body::smem();
    xl_count_launch();
#ifdef XL_HOST_EMU
    (void)stream;
    std::vector<std::vector<char>> bufs((size_t)grid.x, std::vector<char>(smem + 64));
    float2* ptrs[8];
    for (int bx = 0; bx < grid.x; ++bx) ptrs[bx] = (float2*)bufs[(size_t)bx].data();
    xl_emu_gridDim.x = grid.x; xl_emu_gridDim.y = grid.y; xl_emu_gridDim.z = 1;
    for (int by = 0; by < grid.y; ++by) {
        for (int bx = 0; bx < grid.x; ++bx) {
            xl_emu_blockIdx.x = bx; xl_emu_blockIdx.y = by; xl_emu_blockIdx.z = 0;
            Body::run1(p, ptrs[bx]);
        }
        for (int bx = 0; bx < grid.x; ++bx) {
            xl_emu_blockIdx.x = bx; xl_emu_blockIdx.y = by; xl_emu_blockIdx.z = 0;
            Body::run2(p, ptrs[bx], XlPeers{ptrs});
        }
    }
    return XL_OK;
#else
    static std::atomic<unsigned long long> attr_mask{0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 64 && !((attr_mask.load(std::memory_order_acquire) >> dev) & 1ull)) {
        cudaError_t e = cudaFuncSetAttribute(xl_kernel_cluster<Body>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return xl_fail(XL_E_CUDA, "cudaFuncSetAttribute: %s (smem %lld)", cudaGetErrorString(e), (long long)smem);
        attr_mask.fetch_or(1ull << dev, std::memory_order_release);
    }
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(grid.x, grid.y, 1);
    cfg.blockDim = dim3(Body::NT, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = (unsigned)grid.x; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    void* rec = xl_prof_begin(Body::name(), stream);
    cudaError_t e = cudaLaunchKernelEx(&cfg, xl_kernel_cluster<Body>, p);
    if (rec) xl_prof_end(rec, stream);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e != cudaSuccess) return xl_fail(XL_E_CUDA, "cluster kernel launch: %s", cudaGetErrorString(e));
    return XL_OK;
#endif
}

// Persistent kernels: `per_sm` resident CTAs per SM walk `items` work items (blockIdx.x, blockIdx.x + gridDim.x, ...).
// The host emulation uses 3 CTAs so that every CTA walks several items.
template <class Body> static int xl_launch_persistent(int items, int per_sm, xl_stream_t stream, const typename Body::Params& p) {
#ifdef XL_HOST_EMU
    const int slots = 3;
#else
    // resident CTAs per SM: what the register budget of the launch bounds (XlMinBlocks: 512 threads per SM) and the shared
    // memory (227 KB per SM, 1 KB reserved per CTA) allow -- 2 at L = 4096, 4 at L = 2048, 8 at L = 1024
    int fit = (int)((227u * 1024u) / (Body::smem() + 1024u));
    if (fit > XlMinBlocks<Body>::value) fit = XlMinBlocks<Body>::value;
    if (fit > per_sm) per_sm = fit;
    const int slots = per_sm * xl_sm_count();
#endif
    (void)per_sm;
    return xl_launch<Body>(XlDim{items < slots ? items : slots, 1}, stream, p);
}

// XL_DEV_FAST (development builds only, `XL_FAST=1 python -m xlumina_b200.build`): instantiate the two large sizes only
#ifdef XL_DEV_FAST
#define XL_SMALL_L_CASES(...)
#else
#define XL_SMALL_L_CASES(...)                                        \
        case 32: { constexpr int XL = 32; __VA_ARGS__; } break;      \
        case 64: { constexpr int XL = 64; __VA_ARGS__; } break;      \
        case 128: { constexpr int XL = 128; __VA_ARGS__; } break;    \
        case 256: { constexpr int XL = 256; __VA_ARGS__; } break;    \
        case 512: { constexpr int XL = 512; __VA_ARGS__; } break;    \
        case 1024: { constexpr int XL = 1024; __VA_ARGS__; } break;
#endif
#define XL_FOR_L(L, ...)                                      \
    switch (L) {                                              \
        XL_SMALL_L_CASES(__VA_ARGS__)                         \
        case 2048: { constexpr int XL = 2048; __VA_ARGS__; } break;  \
        case 4096: { constexpr int XL = 4096; __VA_ARGS__; } break;  \
        default: return xl_fail(XL_E_UNSUPPORTED, "padded length %s%lld outside [32,4096]", "", (long long)(L)); \
    }

static inline int xl_groups(int n) { return (n + 2 - 1) / 2; }   // CTAs needed for n lines (XL_V = 2 lines per CTA)

// ------------------------------------------------------------------------------------------------ helpers
static inline int next_pow2(int v) { int p = 1; while (p < v) p <<= 1; return p; }
static inline size_t align_up(size_t v) { return (v + 255) & ~(size_t)255; }
struct Carver {
    char* base; size_t off, cap;
    void* take(size_t bytes) { void* p = base + off; off += align_up(bytes); return p; }
    bool ok() const { return off <= cap; }
};
int zero_async(void* p, size_t bytes, xl_stream_t s);
static inline bool aligned16(const void* p) { return ((uintptr_t)p & 15) == 0; }
static inline int pointwise_grid(size_t n, int nt) { return (int)((n + nt - 1) / nt); }
