// xl_platform.h -- one source, two compilations.
//
//  * nvcc (product): kernels for sm_100a.
//  * g++ -DXL_HOST_EMU (tests only): the SAME kernel bodies run as plain C++ so that index math, butterflies,
//    twiddles and the fused pre/post factors can be unit-tested in a container without a GPU
//    (tests/emu/).  A CTA becomes a sequence of "phases"; inside a phase every thread id is visited by a loop, and
//    __syncthreads() falls between phases.  Kernel bodies therefore never carry per-thread registers across a
//    barrier.  The emulation library is never loaded by the xlumina_b200 package.
#pragma once
#include <stdint.h>
#include <stddef.h>
#include <math.h>

#ifdef XL_HOST_EMU
struct float2 { float x, y; };
struct double2 { double x, y; };
static inline float2 make_float2(float x, float y) { float2 r; r.x = x; r.y = y; return r; }
static inline double2 make_double2(double x, double y) { double2 r; r.x = x; r.y = y; return r; }
#define XL_DEV inline
#define XL_DEVFN static inline
#define XL_RESTRICT
struct xl_dim3 { int x, y, z; };
extern thread_local xl_dim3 xl_emu_blockIdx;
#define XL_BLOCK_X (xl_emu_blockIdx.x)
#define XL_BLOCK_Y (xl_emu_blockIdx.y)
#define XL_BLOCK_Z (xl_emu_blockIdx.z)
#define XL_THREADS(tid, nthr) for (int tid = 0; tid < (nthr); ++tid)
#define XL_SYNC() do { } while (0)
static inline void xl_sincospi(double a, double* s, double* c) { *s = sin(M_PI * a); *c = cos(M_PI * a); }
static inline void xl_sincospif(float a, float* s, float* c) { *s = (float)sin(M_PI * (double)a); *c = (float)cos(M_PI * (double)a); }
static inline void xl_sincosf(float a, float* s, float* c) { *s = sinf(a); *c = cosf(a); }
static inline float xl_ldg(const float* p) { return *p; }
static inline float2 xl_ldg(const float2* p) { return *p; }
static inline double xl_ldg(const double* p) { return *p; }
static inline void xl_atomic_add(double* p, double v) { *p += v; }
#else
#include <cuda_runtime.h>
#define XL_DEV __device__ __forceinline__
#define XL_DEVFN __device__
#define XL_RESTRICT __restrict__
#define XL_BLOCK_X ((int)blockIdx.x)
#define XL_BLOCK_Y ((int)blockIdx.y)
#define XL_BLOCK_Z ((int)blockIdx.z)
#define XL_THREADS(tid, nthr) for (int tid = (int)threadIdx.x, _xl_once = 1; _xl_once; _xl_once = 0)
#define XL_SYNC() __syncthreads()
XL_DEV void xl_sincospi(double a, double* s, double* c) { sincospi(a, s, c); }
XL_DEV void xl_sincospif(float a, float* s, float* c) { sincospif(a, s, c); }
XL_DEV void xl_sincosf(float a, float* s, float* c) { sincosf(a, s, c); }
XL_DEV float xl_ldg(const float* p) { return __ldg(p); }
XL_DEV float2 xl_ldg(const float2* p) { return __ldg(p); }
XL_DEV double xl_ldg(const double* p) { return __ldg(p); }
XL_DEV void xl_atomic_add(double* p, double v) { atomicAdd(p, v); }
#endif

typedef float2 cf;

XL_DEV cf cf_make(float x, float y) { return make_float2(x, y); }
XL_DEV cf cf_add(cf a, cf b) { return make_float2(a.x + b.x, a.y + b.y); }
XL_DEV cf cf_sub(cf a, cf b) { return make_float2(a.x - b.x, a.y - b.y); }
XL_DEV cf cf_mul(cf a, cf b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
XL_DEV cf cf_mulc(cf a, cf b) { return make_float2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y); }  // a * conj(b)
XL_DEV cf cf_conj(cf a) { return make_float2(a.x, -a.y); }
XL_DEV cf cf_scale(cf a, float s) { return make_float2(a.x * s, a.y * s); }
XL_DEV cf cf_fma(cf a, cf b, cf acc) {  // acc + a*b
    return make_float2(acc.x + a.x * b.x - a.y * b.y, acc.y + a.x * b.y + a.y * b.x);
}
