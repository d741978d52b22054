// ubench.cu -- standalone B200 microbenchmarks that drive design decisions of the FFT engine (not part of the product):
//   1. FADD/FFMA vs FADD2/FFMA2 issue throughput per SM        2. LDS.128/STS.128 shared-memory bandwidth
//   3. raw XlFft<L>::conv (forward -> multiply -> inverse) throughput per line for several L
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo scripts/ubench.cu -o gpurun_out/ubench
#include "../xlumina_b200/csrc/xl_fft.cuh"
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); return 1; } } while (0)

template <int MODE> __global__ void __launch_bounds__(256) k_fp(float* out, int iters, float a, float b) {
    float2 acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = make_float2(threadIdx.x + i, threadIdx.x - i);
    float2 va = make_float2(a, a * 1.0001f), vb = make_float2(b, b * 0.999f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            if (MODE == 0) { acc[i].x = acc[i].x + va.x; acc[i].y = acc[i].y + va.y; }                       // 2 FADD
            if (MODE == 1) { acc[i] = __fadd2_rn(acc[i], va); }                                              // 1 FADD2
            if (MODE == 2) { acc[i].x = fmaf(acc[i].x, va.x, vb.x); acc[i].y = fmaf(acc[i].y, va.y, vb.y); } // 2 FFMA
            if (MODE == 3) { acc[i] = __ffma2_rn(acc[i], va, vb); }                                          // 1 FFMA2
            if (MODE == 4) { acc[i] = __fadd2_rn(va, make_float2(-acc[i].y, acc[i].x)); }                    // FADD2 with .LO_HI.NP operand
            if (MODE == 5) { acc[i] = __ffma2_rn(make_float2(-acc[i].y, acc[i].x), make_float2(va.y, va.y), vb); }  // FFMA2 swizzle + broadcast
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += acc[i].x + acc[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE> __global__ void __launch_bounds__(256) k_smem(float* out, int iters) {
    extern __shared__ float4 sm[];
    const int t = threadIdx.x;
    for (int i = t; i < 2048; i += 256) sm[i] = make_float4(i, 1, 2, 3);
    __syncthreads();
    float4 acc = make_float4(0, 0, 0, 0);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (MODE == 0) { float4 v = sm[(t + 256 * j + it) & 2047]; acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w; }
            if (MODE == 1) { sm[(t + 256 * j + it) & 2047] = acc; acc.x += 1.f; }
        }
    }
    __syncthreads();
    out[blockIdx.x * blockDim.x + t] = acc.x + acc.y + acc.z + acc.w + sm[t].x;
}

// raw engine: V lines per CTA, out of place; PRUNE: input upper half is zero and output upper half is dropped
template <int L, int V, bool PRUNE> struct ConvOp : XlOpBase {
    static constexpr bool kInLoHalf = PRUNE, kOutLoHalf = PRUNE;
    static constexpr int R1 = xl_first_radix(L), S1 = L / R1;
    const cf* in; cf* out; float s;
    XL_DEV void load(int i, cf* v, int stride) const {
#pragma unroll
        for (int l = 0; l < V; ++l) v[l * stride] = in[(size_t)l * L + i];
    }
    XL_DEV void spec(int beta, cf* v) const {
#pragma unroll
        for (int q = 0; q < 16 * V; ++q) v[q] = cf_scale(v[q], s);
    }
    XL_DEV void store_vec(int n, const cf* v) const {
#pragma unroll
        for (int l = 0; l < V; ++l)
#pragma unroll
            for (int j = 0; j < (PRUNE ? R1 / 2 : R1); ++j) out[(size_t)l * L + n + S1 * j] = v[l * R1 + j];
    }
};
// compute-only probe: operands come from / go to the shared-memory tile itself (values are garbage, timing is what counts)
template <int L, int V> struct SmemOp : XlOpBase {
    static constexpr int R1 = xl_first_radix(L), S1 = L / R1;
    cf* s; float sc;
    XL_DEV void load(int i, cf* v, int stride) const { XlTile<V>::ld(s, i, v, stride); }
    XL_DEV void spec(int beta, cf* v) const {
#pragma unroll
        for (int q = 0; q < 16 * V; ++q) v[q] = cf_scale(v[q], sc);
    }
    XL_DEV void store_vec(int n, const cf* v) const {
#pragma unroll
        for (int j = 0; j < R1; ++j) XlTile<V>::st(s, n + S1 * j, v + j, R1);
    }
};
template <int L, int V, int MINB, bool PRUNE, int REPS = 1> __global__ void __launch_bounds__(xl_threads(L), MINB) k_conv(const cf* in, cf* out, const cf* tw, int nsets) {
    extern __shared__ float4 sm4[];
    cf* sm = (cf*)sm4;
    cf* t = sm + xl_tile_elems(L, V);
    XlFft<L, V>::init_tw(t, tw);
    for (int p = blockIdx.x; p < nsets; p += gridDim.x) {
        ConvOp<L, V, PRUNE> op;
        op.in = in + (size_t)(V * p) * L; op.out = out + (size_t)(V * p) * L; op.s = 1.0f / L;
        XlFft<L, V>::conv(sm, t, op);
        __syncthreads();
        if (REPS > 1) {
            SmemOp<L, V> sop; sop.s = sm; sop.sc = 1.0f / L;
            for (int r = 1; r < REPS; ++r) { XlFft<L, V>::conv(sm, t, sop); __syncthreads(); }
        }
    }
}

// forward-only probe (the shape of the row kernels): PAT 0 = contiguous row-major spectra, 1 = blocked [g/2][y][2] scatter
// (8-byte stores), 2 = blocked scatter with lane-pair exchange (16-byte stores)
template <int L, int PAT> struct FwdOp : XlOpBase {
    static constexpr bool kInLoHalf = true;
    const cf* in; cf* out; int y0, nrows;
    XL_DEV void load(int i, cf* v, int stride) const {
#pragma unroll
        for (int l = 0; l < 2; ++l) v[l * stride] = i < L / 2 ? in[(size_t)(y0 + l) * (L / 2) + i] : cf_zero();
    }
    XL_DEV void spec(int beta, const cf* v) const {
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            const int g = q * (L / 16) + beta;
            if (PAT == 0) { out[(size_t)y0 * L + g] = v[q]; out[(size_t)(y0 + 1) * L + g] = v[16 + q]; }
            if (PAT == 1) {
                out[((size_t)(g / 2) * nrows + y0) * 2 + (g % 2)] = v[q];
                out[((size_t)(g / 2) * nrows + y0 + 1) * 2 + (g % 2)] = v[16 + q];
            }
            if (PAT == 2) {
                // even lane keeps row y0 (its slot g and the neighbour's g+1), odd lane keeps row y0+1
                const bool odd = threadIdx.x & 1;
                const cf give = odd ? v[q] : v[16 + q], keep = odd ? v[16 + q] : v[q];
                cf got;
                got.x = __shfl_xor_sync(0xffffffffu, give.x, 1);
                got.y = __shfl_xor_sync(0xffffffffu, give.y, 1);
                const cf lo = odd ? got : keep, hi = odd ? keep : got;
                float4* dst = reinterpret_cast<float4*>(out + ((size_t)(g / 2) * nrows + y0 + (odd ? 1 : 0)) * 2);
                *dst = make_float4(lo.x, lo.y, hi.x, hi.y);
            }
        }
    }
    XL_DEV void store_vec(int, const cf*) const {}
};
template <int L, int PAT, int PERSIST> __global__ void __launch_bounds__(xl_threads(L), 2) k_fwd(const cf* in, cf* out, const cf* tw, int nrows) {
    extern __shared__ float4 sm4[];
    cf* sm = (cf*)sm4;
    cf* t = sm + xl_tile_elems(L, 2);
    XlFft<L, 2>::init_tw(t, tw);
    for (int p = blockIdx.x; p < nrows / 2; p += gridDim.x) {
        FwdOp<L, PAT> op;
        op.in = in; op.out = out; op.y0 = 2 * p; op.nrows = nrows;
        XlFft<L, 2>::forward(sm, t, op);
        if (PERSIST) __syncthreads();
    }
}

template <class F> static float time_ms(F f, int reps) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(e0);
    for (int i = 0; i < reps; ++i) f();
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    return ms / reps;
}

template <int L, int V, int MINB, bool PRUNE, int REPS = 1> static int run_conv(const cf* tw, int persistent) {
    const int nlines = (4096 / L) * 8192;  // 256 MiB in + 256 MiB out
    const int nsets = nlines / V;
    cf *in, *out;
    CK(cudaMalloc(&in, (size_t)nlines * L * sizeof(cf)));
    CK(cudaMalloc(&out, (size_t)nlines * L * sizeof(cf)));
    CK(cudaMemset(out, 0, (size_t)nlines * L * sizeof(cf)));
    std::vector<cf> h((size_t)L);
    for (int i = 0; i < L; ++i) h[i] = (PRUNE && i >= L / 2) ? make_float2(0.f, 0.f) : make_float2((float)((i * 37) % 101) / 101.f, (float)((i * 11) % 17) / 17.f);
    for (int p = 0; p < nlines; ++p) CK(cudaMemcpy(in + (size_t)p * L, h.data(), sizeof(cf) * L, cudaMemcpyHostToDevice));
    size_t smem = xl_smem_bytes(L, V);
    auto kern = k_conv<L, V, MINB, PRUNE, REPS>;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, xl_threads(L), smem));
    cudaFuncAttributes fa; CK(cudaFuncGetAttributes(&fa, kern));
    int grid = persistent ? 148 * occ : nsets;
    float ms = time_ms([&] { kern<<<grid, xl_threads(L), smem>>>(in, out, tw, nsets); }, 10);
    CK(cudaGetLastError());
    std::vector<cf> r((size_t)L);
    CK(cudaMemcpy(r.data(), out + (size_t)(nlines - 1) * L, sizeof(cf) * L, cudaMemcpyDeviceToHost));
    double err = 0, nrm = 0;
    for (int i = 0; i < (PRUNE ? L / 2 : L); ++i) { err += (r[i].x - h[i].x) * (r[i].x - h[i].x) + (r[i].y - h[i].y) * (r[i].y - h[i].y); nrm += h[i].x * h[i].x + h[i].y * h[i].y; }
    double lines = nlines;
    printf("conv L=%4d V=%d minb=%d regs=%3d occ=%d %s %s: %7.1f us  %6.1f ns/line(fwd+inv)  %6.1f GB/s  relerr %.1e  clk/FFT/SM=%.0f reps=%d\n",
           L, V, MINB, fa.numRegs, occ, PRUNE ? "pruned" : "full  ", persistent ? "persist" : "grid   ", ms * 1e3, ms * 1e6 / lines,
           (PRUNE ? 1.0 : 2.0) * lines * L * 8 / ms / 1e6, sqrt(err / nrm), ms * 1e-3 * 1.965e9 * 148 / lines / 2 / REPS, REPS);
    cudaFree(in); cudaFree(out);
    return 0;
}

template <int L, int PAT, int PERSIST> static int run_fwd(const cf* tw) {
    const int nrows = 2048 * 4;   // 4 fields of 2048 rows, input half-length rows (zero padded on the fly)
    cf *in, *out;
    CK(cudaMalloc(&in, (size_t)nrows * (L / 2) * sizeof(cf)));
    CK(cudaMalloc(&out, (size_t)nrows * L * sizeof(cf)));
    CK(cudaMemset(in, 0, (size_t)nrows * (L / 2) * sizeof(cf)));
    size_t smem = xl_smem_bytes(L, 2);
    auto kern = k_fwd<L, PAT, PERSIST>;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, xl_threads(L), smem));
    int grid = PERSIST ? 148 * occ : nrows / 2;
    float ms = time_ms([&] { kern<<<grid, xl_threads(L), smem>>>(in, out, tw, nrows); }, 10);
    CK(cudaGetLastError());
    printf("fwd  L=%4d pat=%d %s occ=%d: %7.1f us for %d rows  -> %6.1f us per 2048 rows, clk/FFT/SM=%.0f\n", L, PAT, PERSIST ? "persist" : "grid   ", occ,
           ms * 1e3, nrows, ms * 1e3 * 2048 / nrows, ms * 1e-3 * 1.965e9 * 148 / nrows);
    cudaFree(in); cudaFree(out);
    return 0;
}

int main(int argc, char** argv) {
    const int only = argc > 1 ? atoi(argv[1]) : -1;  // run a single conv variant (for ncu)
    int vi = 0;
    cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr, 0));
    printf("device %s, %d SMs, clock %d kHz\n", pr.name, pr.multiProcessorCount, pr.clockRate);
    float* out; CK(cudaMalloc(&out, 148 * 8 * 256 * sizeof(float)));
    const int iters = only >= 0 ? 1 : 4096;
    const char* names[6] = {"FADD  x2", "FADD2   ", "FFMA  x2", "FFMA2   ", "FADD2 swz", "FFMA2 swz"};
    for (int m = 0; m < 6; ++m) {
        float ms = 0;
        if (m == 0) ms = time_ms([&] { k_fp<0><<<148 * 8, 256>>>(out, iters, 1.0001f, 0.5f); }, 5);
        if (m == 1) ms = time_ms([&] { k_fp<1><<<148 * 8, 256>>>(out, iters, 1.0001f, 0.5f); }, 5);
        if (m == 2) ms = time_ms([&] { k_fp<2><<<148 * 8, 256>>>(out, iters, 1.0001f, 0.5f); }, 5);
        if (m == 3) ms = time_ms([&] { k_fp<3><<<148 * 8, 256>>>(out, iters, 1.0001f, 0.5f); }, 5);
        if (m == 4) ms = time_ms([&] { k_fp<4><<<148 * 8, 256>>>(out, iters, 1.0001f, 0.5f); }, 5);
        if (m == 5) ms = time_ms([&] { k_fp<5><<<148 * 8, 256>>>(out, iters, 1.0001f, 0.5f); }, 5);
        double lane_ops = 148.0 * 8 * 256 * (double)iters * 16 * 2;  // fp32 lane-operations
        printf("%s: %.3f ms  %.1f lane-ops/clk/SM (@1.965 GHz)  %.2f T lane-op/s\n", names[m], ms, lane_ops / (ms * 1e-3) / 1.965e9 / 148, lane_ops / ms / 1e9);
    }
    CK(cudaFuncSetAttribute(k_smem<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768));
    CK(cudaFuncSetAttribute(k_smem<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768));
    for (int m = 0; m < 2; ++m) {
        float ms = m == 0 ? time_ms([&] { k_smem<0><<<148 * 6, 256, 32768>>>(out, 4096); }, 5)
                          : time_ms([&] { k_smem<1><<<148 * 6, 256, 32768>>>(out, 4096); }, 5);
        double bytes = 148.0 * 6 * 256 * 4096.0 * 8 * 16;
        printf("%s: %.3f ms  %.1f B/clk/SM\n", m == 0 ? "LDS.128" : "STS.128", ms, bytes / (ms * 1e-3) / 1.965e9 / 148);
    }
    // twiddles
    std::vector<cf> h(XL_TWN);
    for (int k = 0; k < XL_TWN; ++k) { double a = 2.0 * M_PI * k / XL_TWN; h[k] = make_float2((float)cos(a), (float)(-sin(a))); }
    cf* tw; CK(cudaMalloc(&tw, sizeof(cf) * XL_TWN)); CK(cudaMemcpy(tw, h.data(), sizeof(cf) * XL_TWN, cudaMemcpyHostToDevice));
    if ((only < 0 || only == vi) && run_fwd<4096, 0, 0>(tw)) return 1; ++vi;
    if ((only < 0 || only == vi) && run_fwd<4096, 1, 0>(tw)) return 1; ++vi;
    if ((only < 0 || only == vi) && run_fwd<4096, 2, 0>(tw)) return 1; ++vi;
    if ((only < 0 || only == vi) && run_fwd<4096, 0, 1>(tw)) return 1; ++vi;
    if ((only < 0 || only == vi) && run_fwd<4096, 1, 1>(tw)) return 1; ++vi;
    if ((only < 0 || only == vi) && run_fwd<4096, 2, 1>(tw)) return 1; ++vi;
    if ((only < 0 || only == vi) && run_conv<4096, 2, 1, false>(tw, 0)) return 1; ++vi;
    if ((only < 0 || only == vi) && run_conv<4096, 2, 2, false>(tw, 0)) return 1; ++vi;
    if ((only < 0 || only == vi) && run_conv<4096, 2, 2, true>(tw, 0)) return 1; ++vi;
    if ((only < 0 || only == vi) && run_conv<4096, 2, 2, false>(tw, 1)) return 1; ++vi;
    if ((only < 0 || only == vi) && run_conv<4096, 1, 2, false>(tw, 0)) return 1; ++vi;
    if ((only < 0 || only == vi) && run_conv<4096, 1, 3, false>(tw, 0)) return 1; ++vi;
    if ((only < 0 || only == vi) && run_conv<4096, 1, 4, false>(tw, 0)) return 1; ++vi;
    if ((only < 0 || only == vi) && run_conv<4096, 1, 4, true>(tw, 0)) return 1; ++vi;
    if ((only < 0 || only == vi) && run_conv<4096, 1, 4, false>(tw, 1)) return 1; ++vi;
    if ((only < 0 || only == vi) && run_conv<2048, 2, 4, false>(tw, 0)) return 1; ++vi;
    if ((only < 0 || only == vi) && run_conv<2048, 1, 8, false>(tw, 0)) return 1; ++vi;
    if ((only < 0 || only == vi) && run_conv<1024, 2, 8, false>(tw, 0)) return 1; ++vi;
    if ((only < 0 || only == vi) && run_conv<256, 2, 16, false>(tw, 0)) return 1; ++vi;
    if ((only < 0 || only == vi) && run_conv<4096, 2, 2, false, 8>(tw, 0)) return 1; ++vi;
    if ((only < 0 || only == vi) && run_conv<4096, 1, 4, false, 8>(tw, 0)) return 1; ++vi;
    if ((only < 0 || only == vi) && run_conv<4096, 1, 3, false, 8>(tw, 0)) return 1; ++vi;
    return 0;
}
