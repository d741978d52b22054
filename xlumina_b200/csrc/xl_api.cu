// xl_api.cu -- C ABI (include/xlprop.h) over the kernels in xl_kernels.cuh: workspace carving, dispatch on the padded
// length, launches.  Compiled by nvcc for sm_100a (product) and, with -DXL_HOST_EMU, by g++ for the test-only host
// emulation of the kernel bodies (tests/emu).
#include "xl_kernels.cuh"
#include "xl_long.cuh"
#include "../../include/xlprop.h"
#include <stdio.h>
#include <string.h>
#include <stdlib.h>
#include <mutex>
#include <vector>

#ifdef XL_HOST_EMU
thread_local xl_dim3 xl_emu_blockIdx;
thread_local xl_dim3 xl_emu_gridDim;
typedef void* xl_stream_t;
#else
typedef cudaStream_t xl_stream_t;
#endif

// ------------------------------------------------------------------------------------------------ errors
static thread_local char g_err[512] = "";
static int xl_fail(int code, const char* fmt, const char* a = "", long long b = 0) {
    snprintf(g_err, sizeof(g_err), fmt, a, b);
    return code;
}
extern "C" int xl_version(void) { return XLPROP_VERSION; }
extern "C" const char* xl_last_error(void) { return g_err; }

// ------------------------------------------------------------------------------------------------ launch
struct XlDim { int x, y; };

// Instrumentation for bench.py: a launch counter (always on) and optional per-kernel CUDA-event timing recorded on the
// launching stream (xl_prof_enable(1); ...; xl_prof_report()).  Event pairs are pooled; nothing is allocated when off.
static long long g_launches = 0;
static int g_prof_on = 0;
#ifndef XL_HOST_EMU
struct XlProfRec { const char* name; cudaEvent_t e0, e1; };
static std::vector<XlProfRec> g_prof;
#endif
extern "C" long long xl_launch_count(void) { return g_launches; }
extern "C" void xl_prof_enable(int on) {
    g_prof_on = on;
#ifndef XL_HOST_EMU
    if (on) {
        for (auto& r : g_prof) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
        g_prof.clear();
    }
#endif
}
// Writes lines "name count total_ms\n" into buf (after synchronising the recorded events); returns bytes written.
extern "C" int xl_prof_report(char* buf, int cap) {
    int n = 0;
    if (cap > 0) buf[0] = 0;
#ifndef XL_HOST_EMU
    struct Acc { const char* name; int count; double ms; };
    std::vector<Acc> acc;
    for (auto& r : g_prof) {
        cudaEventSynchronize(r.e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, r.e0, r.e1);
        bool found = false;
        for (auto& a : acc) if (a.name == r.name) { a.count++; a.ms += ms; found = true; break; }
        if (!found) acc.push_back(Acc{r.name, 1, (double)ms});
    }
    for (auto& a : acc) {
        int w = snprintf(buf + n, cap > n ? cap - n : 0, "%s %d %.6f\n", a.name, a.count, a.ms);
        if (w < 0 || n + w >= cap) break;
        n += w;
    }
#endif
    return n;
}

#ifndef XL_HOST_EMU
// register budget: at least 512/NT CTAs per SM (128 registers per thread), so two L=4096 CTAs overlap their phases on an SM
template <class Body, class = void> struct XlMinBlocks { static constexpr int value = (512 / Body::NT) > 16 ? 16 : (512 / Body::NT); };
template <class Body> struct XlMinBlocks<Body, decltype((void)Body::MINB)> { static constexpr int value = Body::MINB; };   // per-kernel override
template <class Body> __global__ void __launch_bounds__(Body::NT, XlMinBlocks<Body>::value) xl_kernel(const typename Body::Params p) {
    extern __shared__ float4 xl_smem[];
    Body::run(p, (cf*)xl_smem);
}
#endif

template <class Body> static int xl_launch(XlDim grid, xl_stream_t stream, const typename Body::Params& p) {
    if (grid.x <= 0 || grid.y <= 0) return XL_OK;
    const size_t smem = Body::smem();
    ++g_launches;
#ifdef XL_HOST_EMU
    (void)stream;
    std::vector<char> buf(smem + 64);
    xl_emu_gridDim.x = grid.x; xl_emu_gridDim.y = grid.y; xl_emu_gridDim.z = 1;
    for (int by = 0; by < grid.y; ++by)
        for (int bx = 0; bx < grid.x; ++bx) {
            xl_emu_blockIdx.x = bx; xl_emu_blockIdx.y = by; xl_emu_blockIdx.z = 0;
            Body::run(p, (cf*)buf.data());
        }
    return XL_OK;
#else
    static bool attr_set[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 64 && !attr_set[dev]) {
        cudaError_t e = cudaFuncSetAttribute(xl_kernel<Body>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return xl_fail(XL_E_CUDA, "cudaFuncSetAttribute: %s (smem %lld)", cudaGetErrorString(e), (long long)smem);
        attr_set[dev] = true;
    }
    XlProfRec rec;
    if (g_prof_on) {
        rec.name = Body::name();
        cudaEventCreate(&rec.e0); cudaEventCreate(&rec.e1);
        cudaEventRecord(rec.e0, stream);
    }
    xl_kernel<Body><<<dim3(grid.x, grid.y, 1), dim3(Body::NT, 1, 1), smem, stream>>>(p);
    if (g_prof_on) { cudaEventRecord(rec.e1, stream); g_prof.push_back(rec); }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return xl_fail(XL_E_CUDA, "kernel launch: %s", cudaGetErrorString(e));
    return XL_OK;
#endif
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

static int xl_groups(int n) { return (n + XL_V - 1) / XL_V; }   // CTAs needed for n lines

// ------------------------------------------------------------------------------------------------ twiddles
static std::mutex g_tw_mutex;
static cf* g_tw[64] = {0};
static const cf* xl_twiddles() {
    int dev = 0;
#ifndef XL_HOST_EMU
    cudaGetDevice(&dev);
#endif
    if (dev < 0 || dev >= 64) return 0;
    std::lock_guard<std::mutex> lk(g_tw_mutex);
    if (g_tw[dev]) return g_tw[dev];
    std::vector<cf> h(XL_TWN);
    for (int k = 0; k < XL_TWN; ++k) {
        double a = 2.0 * M_PI * (double)k / (double)XL_TWN;
        h[k].x = (float)cos(a);
        h[k].y = (float)(-sin(a));
    }
#ifdef XL_HOST_EMU
    g_tw[dev] = (cf*)malloc(sizeof(cf) * XL_TWN);
    memcpy(g_tw[dev], h.data(), sizeof(cf) * XL_TWN);
#else
    cf* d = 0;
    if (cudaMalloc(&d, sizeof(cf) * XL_TWN) != cudaSuccess) return 0;
    if (cudaMemcpy(d, h.data(), sizeof(cf) * XL_TWN, cudaMemcpyHostToDevice) != cudaSuccess) return 0;
    g_tw[dev] = d;
#endif
    return g_tw[dev];
}

// ------------------------------------------------------------------------------------------------ helpers
static int next_pow2(int v) { int p = 1; while (p < v) p <<= 1; return p; }
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
static size_t align_up(size_t v) { return (v + 255) & ~(size_t)255; }
struct Carver {
    char* base; size_t off, cap;
    void* take(size_t bytes) { void* p = base + off; off += align_up(bytes); return p; }
    bool ok() const { return off <= cap; }
};

static int zero_async(void* p, size_t bytes, xl_stream_t s) {
#ifdef XL_HOST_EMU
    (void)s; memset(p, 0, bytes); return XL_OK;
#else
    cudaError_t e = cudaMemsetAsync(p, 0, bytes, s);
    return e == cudaSuccess ? XL_OK : xl_fail(XL_E_CUDA, "cudaMemsetAsync: %s", cudaGetErrorString(e));
#endif
}

// ================================================================================================ RS / VRS
extern "C" size_t xl_rs_transfer_bytes(int N) {
    size_t L = (size_t)xl_rs_padded_length(N);
    return L * L * sizeof(cf);
}
extern "C" size_t xl_rs_workspace_bytes(int N, int nfields, int want_grad_z) {
    size_t L = (size_t)xl_rs_padded_length(N);
    if (!L || nfields < 1) return 0;
    size_t spec = align_up((size_t)nfields * L * N * sizeof(cf));
    size_t total = spec;
    if (want_grad_z) total += spec + align_up(L * L * sizeof(cf));
    total += align_up((size_t)3 * N * N * sizeof(cf));  // VRS backward: adjoint of the 3 components before the fold
    return total;
}

static int rs_base_params(XlRsParams& p, int N, double dx, double dy, double k) {
    memset(&p, 0, sizeof(p));
    p.N = N;
    p.L = xl_rs_padded_length(N);
    if (!p.L) return xl_fail(XL_E_UNSUPPORTED, "RS: N=%s%lld unsupported (padded length must be in [32,4096])", "", N);
    p.rows = N; p.chunk_rows = N;
    p.dx = dx; p.dy = dy; p.k = k;
    p.hscale = (float)(dx * dy / ((double)p.L * (double)p.L));
    p.tw = xl_twiddles();
    if (!p.tw) return xl_fail(XL_E_CUDA, "twiddle table allocation failed%s", "");
    return XL_OK;
}

static int rs_transfer_impl(XlRsParams p, cf* H, const double* z, int deriv, xl_stream_t st) {
    p.H = H; p.z = z;
    p.flags = deriv ? XL_F_DERIV : 0;
    const int L = p.L;
    p.rows = L; p.hrow0 = 0; p.hstore_all = 0;   // row spectra of h live inside H: [L/2][L][2], rows 0..L/2
    int rc;
    XL_FOR_L(L, rc = xl_launch<XlHRows<XL>>(XlDim{xl_groups(L / 2 + 1), 1}, st, p));
    if (rc) return rc;
    XL_FOR_L(L, rc = xl_launch<XlHCols<XL>>(XlDim{L / XL_V, 1}, st, p));
    return rc;
}

extern "C" int xl_rs_transfer(void* H, const double* z, int N, double dx, double dy, double k, int deriv, void* stream) {
    if (!H || !z) return xl_fail(XL_E_BAD_ARG, "xl_rs_transfer: null pointer%s", "");
    XlRsParams p;
    int rc = rs_base_params(p, N, dx, dy, k);
    if (rc) return rc;
    return rs_transfer_impl(p, (cf*)H, z, deriv, (xl_stream_t)stream);
}

// rows fwd -> cols conv -> rows inv on `nfields` planes
// rows fwd -> cols conv -> rows inv on `nfields` planes, all fields of a stage in ONE launch.  Measured alternatives that
// were slower or no faster: one launch per field and stage (keeps a field's spectra L2-resident but quantises each
// 1024-CTA launch into 3.5 waves of 296), and a per-field software pipeline over auxiliary streams (grids this large do
// not co-schedule: the second kernel only fills the first one's tail).
static int rs_apply_impl(XlRsParams p, xl_stream_t st, cf* keep = 0) {
    const int L = p.L, N = p.N;
    int rc;
    p.f0 = 0;
#ifdef XL_EXP_KEEP_SPECTRA
    if (keep) {   // row spectra into `keep` (left intact), column pass keep -> p.spec, inverse rows from p.spec
        XlRsParams q = p;
        q.spec = keep; q.spec2 = p.spec;
        XL_FOR_L(L, rc = xl_launch<XlRsRowsFwd<XL>>(XlDim{xl_groups(N), p.nfields}, st, q));
        if (rc) return rc;
        XL_FOR_L(L, rc = xl_launch<XlRsColsKeep<XL>>(XlDim{L / XL_V, p.nfields}, st, q));
        if (rc) return rc;
        XL_FOR_L(L, rc = xl_launch<XlRsRowsInv<XL>>(XlDim{xl_groups(N), p.nfields}, st, p));
        return rc;
    }
#endif
    (void)keep;
    XL_FOR_L(L, rc = xl_launch<XlRsRowsFwd<XL>>(XlDim{xl_groups(N), p.nfields}, st, p));
    if (rc) return rc;
#if defined(XL_EXP_K2_PERSIST)
    {   // persistent CTAs, two per SM (the emulation uses 3 CTAs so that every CTA walks several items)
        int slots = 3;
#ifndef XL_HOST_EMU
        int dev = 0, sms = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        slots = 2 * (sms > 0 ? sms : 148);
#endif
        const int items = (L / XL_V) * p.nfields;
        XL_FOR_L(L, rc = xl_launch<XlRsColsPersist<XL>>(XlDim{items < slots ? items : slots, 1}, st, p));
    }
#elif defined(XL_EXP_K2_STAGE)
    XL_FOR_L(L, rc = xl_launch<XlRsColsStage<XL>>(XlDim{L / XL_V, p.nfields}, st, p));
#else
    XL_FOR_L(L, rc = xl_launch<XlRsCols<XL>>(XlDim{L / XL_V, p.nfields}, st, p));
#endif
    if (rc) return rc;
    XL_FOR_L(L, rc = xl_launch<XlRsRowsInv<XL>>(XlDim{xl_groups(N), p.nfields}, st, p));
    return rc;
}

static int rs_fwd_common(const void* in, void* out, void* H, const double* z, int N, int nfields, int vrs,
                         double x0, double y0, double dx, double dy, double k, int flags,
                         void* ws, size_t ws_bytes, xl_stream_t st, void* keep = 0) {
    if (!in || !out || !H || !z || !ws) return xl_fail(XL_E_BAD_ARG, "rs_fwd: null pointer%s", "");
    XlRsParams p;
    int rc = rs_base_params(p, N, dx, dy, k);
    if (rc) return rc;
    if (ws_bytes < xl_rs_workspace_bytes(N, nfields, 0)) return xl_fail(XL_E_WORKSPACE, "rs_fwd: workspace too small%s", "");
    if (!(flags & XL_REUSE_H)) { rc = rs_transfer_impl(p, (cf*)H, z, 0, st); if (rc) return rc; }
    Carver c{(char*)ws, 0, ws_bytes};
    p.spec = (cf*)c.take((size_t)nfields * p.L * N * sizeof(cf));
    p.in = (const cf*)in; p.out = (cf*)out; p.H = (cf*)H; p.z = z;
    p.nfields = nfields; p.x0 = x0; p.y0 = y0;
    p.flags = (flags & (XL_CONJ_IN | XL_CONJ_OUT)) | (vrs ? XL_F_VRS : 0);
    if (keep && (flags & XL_CONJ_IN)) return xl_fail(XL_E_BAD_ARG, "rs_fwd_keep: kept spectra are those of the unconjugated field%s", "");
    return rs_apply_impl(p, st, (cf*)keep);
}

extern "C" int xl_rs_fwd(const void* in, void* out, void* H, const double* z, int N, int nfields,
                         double dx, double dy, double k, int flags, void* ws, size_t ws_bytes, void* stream) {
    if (nfields < 1) return xl_fail(XL_E_BAD_ARG, "xl_rs_fwd: nfields < 1%s", "");
    return rs_fwd_common(in, out, H, z, N, nfields, 0, 0.0, 0.0, dx, dy, k, flags, ws, ws_bytes, (xl_stream_t)stream);
}
extern "C" int xl_vrs_fwd(const void* exy, void* out, void* H, const double* z, int N, double x0, double y0,
                          double dx, double dy, double k, int flags, void* ws, size_t ws_bytes, void* stream) {
    return rs_fwd_common(exy, out, H, z, N, 3, 1, x0, y0, dx, dy, k, flags, ws, ws_bytes, (xl_stream_t)stream);
}

static int rs_bwd_common(const void* in, const void* out, const void* ct_out, void* ct_in, double* grad_z, const void* H,
                         const double* z, int N, int nfields, int vrs, double x0, double y0, double dx, double dy, double k,
                         int flags, void* ws, size_t ws_bytes, xl_stream_t st, const void* kept = 0) {
    if (!ct_out || !ct_in || !H || !ws || !z) return xl_fail(XL_E_BAD_ARG, "rs_bwd: null pointer%s", "");
    if (grad_z && (!in || !out)) return xl_fail(XL_E_BAD_ARG, "rs_bwd: grad_z needs the primal input and output%s", "");
    XlRsParams p;
    int rc = rs_base_params(p, N, dx, dy, k);
    if (rc) return rc;
    if (ws_bytes < xl_rs_workspace_bytes(N, nfields, grad_z != 0)) return xl_fail(XL_E_WORKSPACE, "rs_bwd: workspace too small%s", "");
    const int L = p.L;
    Carver c{(char*)ws, 0, ws_bytes};
    const size_t spec_bytes = (size_t)nfields * L * N * sizeof(cf);
    p.spec = (cf*)c.take(spec_bytes);
    cf* tmp3 = (cf*)c.take((size_t)3 * N * N * sizeof(cf));
    p.nfields = nfields; p.H = (cf*)H; p.z = z; p.x0 = x0; p.y0 = y0;
    cf* dst = vrs ? tmp3 : (cf*)ct_in;

    if (grad_z) {
        p.spec2 = (cf*)c.take(spec_bytes);
        cf* Hz = (cf*)c.take((size_t)L * L * sizeof(cf));
        rc = rs_transfer_impl(p, Hz, z, 1, st);        // reduced derivative h_z - i k h (xl_rs_h)
        if (rc) return rc;
        {   // the i k h part, exactly, in real space
            XlDotZParams d;
            memset(&d, 0, sizeof(d));
            d.ct = (const cf*)ct_out; d.out = (const cf*)out; d.n = (size_t)nfields * N * N; d.flags = flags & XL_CONJ_IN;
            d.k = k; d.gz = grad_z;
            const size_t per = (size_t)XlDotZ::NT * XlDotZ::PER;
            rc = xl_launch<XlDotZ>(XlDim{(int)((d.n + per - 1) / per), 1}, st, d);
            if (rc) return rc;
        }
        // row spectra of conj(U) -> spec2 (not needed when the forward pass kept its row spectra)
        if (!kept) {
            XlRsParams pw = p;
            pw.in = (const cf*)in; pw.spec = p.spec2;
            pw.flags = XL_F_CONJ_IN | (vrs ? XL_F_VRS : 0);
            XL_FOR_L(L, rc = xl_launch<XlRsRowsFwd<XL>>(XlDim{xl_groups(N), nfields}, st, pw));
            if (rc) return rc;
        }
        // row spectra of the cotangent -> spec
        XlRsParams pc = p;
        pc.in = (const cf*)ct_out;
        pc.flags = (flags & XL_CONJ_IN);
        XL_FOR_L(L, rc = xl_launch<XlRsRowsFwd<XL>>(XlDim{xl_groups(N), nfields}, st, pc));
        if (rc) return rc;
        XlRsParams pg = p;
        pg.H2 = Hz; pg.gz = grad_z;
#ifdef XL_EXP_KEEP_SPECTRA
        if (kept) {
            pg.spec2 = (cf*)kept;
            XL_FOR_L(L, rc = xl_launch<XlRsColsGzKept<XL>>(XlDim{L, nfields}, st, pg));
        } else
#endif
#if defined(XL_EXP_K4_PERSIST)
        {
            int slots = 3;
#ifndef XL_HOST_EMU
            int dev = 0, sms = 0;
            cudaGetDevice(&dev);
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
            slots = 2 * (sms > 0 ? sms : 148);
#endif
            const int items = L * nfields;
            XL_FOR_L(L, rc = xl_launch<XlRsColsGzPersist<XL>>(XlDim{items < slots ? items : slots, 1}, st, pg));
        }
#elif defined(XL_EXP_K4_STAGE)
        XL_FOR_L(L, rc = xl_launch<XlRsColsGzStage<XL>>(XlDim{L, nfields}, st, pg));
#else
        XL_FOR_L(L, rc = xl_launch<XlRsColsGz<XL>>(XlDim{L, nfields}, st, pg));
#endif
        if (rc) return rc;
        XlRsParams po = p;
        po.out = dst;
        po.flags = vrs ? 0 : (flags & XL_CONJ_OUT);
        XL_FOR_L(L, rc = xl_launch<XlRsRowsInv<XL>>(XlDim{xl_groups(N), nfields}, st, po));
        if (rc) return rc;
    } else {
        XlRsParams pa = p;
        pa.in = (const cf*)ct_out; pa.out = dst;
        pa.flags = (flags & XL_CONJ_IN) | (vrs ? 0 : (flags & XL_CONJ_OUT));
        rc = rs_apply_impl(pa, st);
        if (rc) return rc;
    }
    if (vrs) {
        XlFoldParams f;
        memset(&f, 0, sizeof(f));
        f.N = N; f.mode = XL_FOLD_VRS; f.flags = flags & XL_CONJ_OUT;
        f.t = tmp3;
        f.ex = (const cf*)in; f.ey = in ? (const cf*)in + (size_t)N * N : 0;
        f.gx = (cf*)ct_in; f.gy = (cf*)ct_in + (size_t)N * N;
        f.gz = grad_z; f.z = z; f.x0 = x0; f.y0 = y0; f.dx = dx; f.dy = dy;
        const size_t NN = (size_t)N * N;
        rc = xl_launch<XlFold>(XlDim{(int)((NN + XlFold::NT - 1) / XlFold::NT), 1}, st, f);
    }
    return rc;
}

extern "C" int xl_rs_bwd(const void* in, const void* out, const void* ct_out, void* ct_in, double* grad_z, const void* H,
                         const double* z, int N, int nfields, double dx, double dy, double k, int flags,
                         void* ws, size_t ws_bytes, void* stream) {
    if (nfields < 1) return xl_fail(XL_E_BAD_ARG, "xl_rs_bwd: nfields < 1%s", "");
    return rs_bwd_common(in, out, ct_out, ct_in, grad_z, H, z, N, nfields, 0, 0.0, 0.0, dx, dy, k, flags, ws, ws_bytes, (xl_stream_t)stream);
}
extern "C" int xl_vrs_bwd(const void* exy, const void* out, const void* ct_out, void* ct_exy, double* grad_z, const void* H,
                          const double* z, int N, double x0, double y0, double dx, double dy, double k, int flags,
                          void* ws, size_t ws_bytes, void* stream) {
    return rs_bwd_common(exy, out, ct_out, ct_exy, grad_z, H, z, N, 3, 1, x0, y0, dx, dy, k, flags, ws, ws_bytes, (xl_stream_t)stream);
}

#ifdef XL_EXP_KEEP_SPECTRA
// Entry points of the keep-spectra experiment (variant builds only; not part of include/xlprop.h).
extern "C" size_t xl_rs_spectra_bytes(int N, int nfields) {
    const int L = xl_rs_padded_length(N);
    return L ? (size_t)nfields * L * N * sizeof(cf) : 0;
}
extern "C" int xl_rs_fwd_keep(const void* in, void* out, void* H, const double* z, int N, int nfields, double dx, double dy,
                              double k, int flags, void* spectra, void* ws, size_t ws_bytes, void* stream) {
    if (nfields < 1 || !spectra) return xl_fail(XL_E_BAD_ARG, "xl_rs_fwd_keep: bad argument%s", "");
    return rs_fwd_common(in, out, H, z, N, nfields, 0, 0.0, 0.0, dx, dy, k, flags, ws, ws_bytes, (xl_stream_t)stream, spectra);
}
extern "C" int xl_vrs_fwd_keep(const void* exy, void* out, void* H, const double* z, int N, double x0, double y0, double dx,
                               double dy, double k, int flags, void* spectra, void* ws, size_t ws_bytes, void* stream) {
    if (!spectra) return xl_fail(XL_E_BAD_ARG, "xl_vrs_fwd_keep: bad argument%s", "");
    return rs_fwd_common(exy, out, H, z, N, 3, 1, x0, y0, dx, dy, k, flags, ws, ws_bytes, (xl_stream_t)stream, spectra);
}
extern "C" int xl_rs_bwd_kept(const void* in, const void* out, const void* ct_out, void* ct_in, double* grad_z, const void* H,
                              const double* z, int N, int nfields, double dx, double dy, double k, int flags,
                              const void* spectra, void* ws, size_t ws_bytes, void* stream) {
    if (nfields < 1) return xl_fail(XL_E_BAD_ARG, "xl_rs_bwd_kept: nfields < 1%s", "");
    return rs_bwd_common(in, out, ct_out, ct_in, grad_z, H, z, N, nfields, 0, 0.0, 0.0, dx, dy, k, flags, ws, ws_bytes,
                         (xl_stream_t)stream, spectra);
}
extern "C" int xl_vrs_bwd_kept(const void* exy, const void* out, const void* ct_out, void* ct_exy, double* grad_z, const void* H,
                               const double* z, int N, double x0, double y0, double dx, double dy, double k, int flags,
                               const void* spectra, void* ws, size_t ws_bytes, void* stream) {
    return rs_bwd_common(exy, out, ct_out, ct_exy, grad_z, H, z, N, 3, 1, x0, y0, dx, dy, k, flags, ws, ws_bytes,
                         (xl_stream_t)stream, spectra);
}
#endif

// ================================================================================================ slab-decomposed RS
// Stage-level entry points of the multi-GPU RS path (BASELINE.json cfg 5): the N x N field is split
// into row slabs, one per rank; the all-to-all transposes between the stages are the caller's (NCCL via torch.distributed
// in xlumina_b200/slab.py).  Geometry for G ranks, P = xl_slab_padded_length(N):  rows = N/G field rows per rank (even),
// pairs = (P/2)/G x-slot pairs per rank, hrows = xl_slab_h_rows_per_rank(N, G) rows of the y >= 0 half of the impulse
// response per rank.   exchanged layout of a spectra buffer:  [source rank][pairs][rows of that rank][2].
// P <= 4096 runs the single-pass kernels of xl_kernels.cuh; longer lines (up to 32768: N <= 16384) run the split kernels of
// xl_long.cuh (P = R * L0) and need the scratch buffer of xl_slab_scratch_bytes().
static int g_max_line = 4096;   // sub-line length of the split kernels; tests set 32 to exercise them at small sizes
extern "C" void xl_debug_set_max_line(int l) { g_max_line = l == 32 ? 32 : 4096; }
extern "C" int xl_slab_padded_length(int N) {
    if (N < 2) return 0;
    int P = next_pow2(2 * N - 1);
    if (P < 32) P = 32;
    return P <= 8 * g_max_line ? P : 0;
}
struct SlabGeo { int P, L0, R, rows, pairs, hrows; };
static int slab_geo(SlabGeo& g, int N, int G) {
    g.P = xl_slab_padded_length(N);
    if (!g.P) return xl_fail(XL_E_UNSUPPORTED, "slab: N=%s%lld unsupported (padded length must be <= 32768)", "", N);
    if (G < 1 || N % (2 * G) != 0 || (g.P / 2) % G != 0)
        return xl_fail(XL_E_BAD_ARG, "slab: N must be a multiple of 2*G and P/2 a multiple of G (G=%s%lld)", "", G);
    g.L0 = g.P < g_max_line ? g.P : g_max_line;
    g.R = g.P / g.L0;
    g.rows = N / G;
    g.pairs = (g.P / 2) / G;
    int r = (g.P / 2 + 1 + G - 1) / G;
    g.hrows = r + (r & 1);   // row pairs stay on one rank
    return XL_OK;
}
extern "C" int xl_slab_h_rows_per_rank(int N, int G) {
    SlabGeo g;
    return slab_geo(g, N, G) ? 0 : g.hrows;
}
extern "C" size_t xl_slab_scratch_bytes(int N, int G) {
    SlabGeo g;
    if (slab_geo(g, N, G) || g.R == 1) return 256;
    size_t a = (size_t)g.rows * g.P, b = (size_t)g.pairs * g.P * 2, c = (size_t)g.hrows * (g.P / 2 + 1);
    size_t m = a > b ? a : b;
    return (m > c ? m : c) * sizeof(cf);
}
#define XL_FOR_L0(L0, ...)                                             \
    switch (L0) {                                                      \
        case 32: { constexpr int XL = 32; __VA_ARGS__; } break;        \
        case 4096: { constexpr int XL = 4096; __VA_ARGS__; } break;    \
        default: return xl_fail(XL_E_UNSUPPORTED, "split kernels: sub-line length %s%lld", "", (long long)(L0)); \
    }
static int long_params(XlLongParams& q, const SlabGeo& g, int N, double dx, double dy, double k) {
    memset(&q, 0, sizeof(q));
    q.N = N; q.P = g.P; q.R = g.R; q.L0 = g.L0; q.rows = g.rows; q.chunk_rows = g.rows; q.pairs = g.pairs;
    q.dx = dx; q.dy = dy; q.k = k;
    q.hscale = (float)(dx * dy / ((double)g.P * (double)g.P));
    q.tw = xl_twiddles();
    if (!q.tw) return xl_fail(XL_E_CUDA, "twiddle table allocation failed%s", "");
    return XL_OK;
}
static int pointwise_grid(size_t n, int nt) { return (int)((n + nt - 1) / nt); }

// row spectra of this rank's y rows [rank*hrows, (rank+1)*hrows) of the impulse response: R[P/2][hrows][2]
extern "C" int xl_slab_h_rows(void* Rb, const double* z, int N, int G, int rank, double dx, double dy, double k,
                              void* scratch, void* stream) {
    if (!Rb || !z) return xl_fail(XL_E_BAD_ARG, "xl_slab_h_rows: null pointer%s", "");
    SlabGeo g;
    int rc = slab_geo(g, N, G);
    if (rc) return rc;
    xl_stream_t st = (xl_stream_t)stream;
    if (g.R == 1) {
        XlRsParams p;
        if ((rc = rs_base_params(p, N, dx, dy, k))) return rc;
        p.H = (cf*)Rb; p.z = z; p.flags = 0;
        p.rows = g.hrows; p.hrow0 = rank * g.hrows; p.hstore_all = 1;
        const int L = p.L;
        XL_FOR_L(L, rc = xl_launch<XlHRows<XL>>(XlDim{xl_groups(g.hrows), 1}, st, p));
        return rc;
    }
    if (!scratch) return xl_fail(XL_E_BAD_ARG, "xl_slab_h_rows: scratch needed for padded lengths > 4096%s", "");
    XlLongParams q;
    if ((rc = long_params(q, g, N, dx, dy, k))) return rc;
    q.z = z; q.hrow0 = rank * g.hrows; q.hrows = g.hrows; q.scratch = (cf*)scratch; q.spec = (cf*)Rb;
    rc = xl_launch<XlHEval>(XlDim{pointwise_grid((size_t)g.hrows * (g.P / 2 + 1), XlHEval::NT), 1}, st, q);
    if (rc) return rc;
    XL_FOR_L0(g.L0, rc = xl_launch<XlLongHRows<XL>>(XlDim{g.R, xl_groups(g.hrows)}, st, q));
    return rc;
}
// Th = exchanged row spectra of h [G][pairs][hrows][2]  ->  this rank's transfer-function slab Hloc[pairs][P][2]
extern "C" int xl_slab_h_cols(const void* Th, void* Hloc, int N, int G, double dx, double dy, void* stream) {
    if (!Th || !Hloc) return xl_fail(XL_E_BAD_ARG, "xl_slab_h_cols: null pointer%s", "");
    SlabGeo g;
    int rc = slab_geo(g, N, G);
    if (rc) return rc;
    xl_stream_t st = (xl_stream_t)stream;
    if (g.R == 1) {
        XlRsParams p;
        if ((rc = rs_base_params(p, N, dx, dy, 0.0))) return rc;
        const int L = p.L;
        p.spec = (cf*)Th; p.H = (cf*)Hloc; p.chunk_rows = g.hrows; p.nfields = g.pairs;
        XL_FOR_L(L, rc = xl_launch<XlHColsSlab<XL>>(XlDim{g.pairs, 1}, st, p));
        return rc;
    }
    XlLongParams q;
    if ((rc = long_params(q, g, N, dx, dy, 0.0))) return rc;
    q.spec = (cf*)Th; q.H = (cf*)Hloc; q.chunk_rows = g.hrows;
    XL_FOR_L0(g.L0, rc = xl_launch<XlLongHCols<XL>>(XlDim{g.R, g.pairs}, st, q));
    return rc;
}
// this rank's field rows in_local[rows][N]  ->  row spectra S[P/2][rows][2]
extern "C" int xl_slab_rows_fwd(const void* in_local, void* S, int N, int G, int flags, void* stream) {
    if (!in_local || !S) return xl_fail(XL_E_BAD_ARG, "xl_slab_rows_fwd: null pointer%s", "");
    SlabGeo g;
    int rc = slab_geo(g, N, G);
    if (rc) return rc;
    xl_stream_t st = (xl_stream_t)stream;
    if (g.R == 1) {
        XlRsParams p;
        if ((rc = rs_base_params(p, N, 1.0, 1.0, 0.0))) return rc;
        p.in = (const cf*)in_local; p.spec = (cf*)S; p.nfields = 1; p.rows = g.rows; p.flags = flags & XL_CONJ_IN;
        const int L = p.L;
        XL_FOR_L(L, rc = xl_launch<XlRsRowsFwd<XL>>(XlDim{xl_groups(g.rows), 1}, st, p));
        return rc;
    }
    XlLongParams q;
    if ((rc = long_params(q, g, N, 1.0, 1.0, 0.0))) return rc;
    q.in = (const cf*)in_local; q.spec = (cf*)S; q.flags = flags & XL_CONJ_IN;
    XL_FOR_L0(g.L0, rc = xl_launch<XlLongRowsFwd<XL>>(XlDim{g.R, xl_groups(g.rows)}, st, q));
    return rc;
}
// T = exchanged spectra [G][pairs][N/G][2], filtered in place by this rank's transfer-function slab
extern "C" int xl_slab_cols(void* T, const void* Hloc, int N, int G, void* scratch, void* stream) {
    if (!T || !Hloc) return xl_fail(XL_E_BAD_ARG, "xl_slab_cols: null pointer%s", "");
    SlabGeo g;
    int rc = slab_geo(g, N, G);
    if (rc) return rc;
    xl_stream_t st = (xl_stream_t)stream;
    if (g.R == 1) {
        XlRsParams p;
        if ((rc = rs_base_params(p, N, 1.0, 1.0, 0.0))) return rc;
        const int L = p.L;
        p.spec = (cf*)T; p.H = (cf*)Hloc; p.chunk_rows = g.rows; p.nfields = g.pairs;
        XL_FOR_L(L, rc = xl_launch<XlRsColsSlab<XL>>(XlDim{g.pairs, 1}, st, p));
        return rc;
    }
    if (!scratch) return xl_fail(XL_E_BAD_ARG, "xl_slab_cols: scratch needed for padded lengths > 4096%s", "");
    XlLongParams q;
    if ((rc = long_params(q, g, N, 1.0, 1.0, 0.0))) return rc;
    q.spec = (cf*)T; q.H = (cf*)Hloc; q.scratch = (cf*)scratch;
    XL_FOR_L0(g.L0, rc = xl_launch<XlLongCols<XL>>(XlDim{g.R, g.pairs}, st, q));
    if (rc) return rc;
    return xl_launch<XlLongColsCombine>(XlDim{pointwise_grid((size_t)g.pairs * g.L0, XlLongColsCombine::NT), 1}, st, q);
}
// S = spectra exchanged back, [P/2][rows][2]  ->  this rank's output rows out_local[rows][N]
extern "C" int xl_slab_rows_inv(const void* S, void* out_local, int N, int G, int flags, void* scratch, void* stream) {
    if (!S || !out_local) return xl_fail(XL_E_BAD_ARG, "xl_slab_rows_inv: null pointer%s", "");
    SlabGeo g;
    int rc = slab_geo(g, N, G);
    if (rc) return rc;
    xl_stream_t st = (xl_stream_t)stream;
    if (g.R == 1) {
        XlRsParams p;
        if ((rc = rs_base_params(p, N, 1.0, 1.0, 0.0))) return rc;
        p.spec = (cf*)S; p.out = (cf*)out_local; p.nfields = 1; p.rows = g.rows; p.flags = flags & XL_CONJ_OUT;
        const int L = p.L;
        XL_FOR_L(L, rc = xl_launch<XlRsRowsInv<XL>>(XlDim{xl_groups(g.rows), 1}, st, p));
        return rc;
    }
    if (!scratch) return xl_fail(XL_E_BAD_ARG, "xl_slab_rows_inv: scratch needed for padded lengths > 4096%s", "");
    XlLongParams q;
    if ((rc = long_params(q, g, N, 1.0, 1.0, 0.0))) return rc;
    q.spec = (cf*)S; q.out = (cf*)out_local; q.scratch = (cf*)scratch; q.flags = flags & XL_CONJ_OUT;
    XL_FOR_L0(g.L0, rc = xl_launch<XlLongRowsInv<XL>>(XlDim{g.R, xl_groups(g.rows)}, st, q));
    if (rc) return rc;
    return xl_launch<XlLongRowsCombine>(XlDim{pointwise_grid((size_t)g.rows * g.L0, XlLongRowsCombine::NT), 1}, st, q);
}

// ================================================================================================ CZT family
struct CztPlan {
    int N, Mx, My, Ly, Lx, ncomp;
    cf *pre_y, *post_y, *ft_y, *ftT_y, *pre_x, *post_x, *ft_x, *ftT_x;
    cf* mid;   // [ncomp][N][My]
    cf* tmp3;  // [3][N][N]
};
static size_t czt_ws_bytes(int N, int Mx, int My, int ncomp) {
    int Ly = xl_czt_padded_length(N, My), Lx = xl_czt_padded_length(N, Mx);
    if (!Ly || !Lx) return 0;
    size_t t = 0;
    t += 2 * align_up((size_t)N * sizeof(cf)) + align_up((size_t)My * sizeof(cf)) + align_up((size_t)Mx * sizeof(cf));
    t += 2 * align_up((size_t)Ly * sizeof(cf)) + 2 * align_up((size_t)Lx * sizeof(cf));
    t += align_up((size_t)ncomp * N * My * sizeof(cf));
    t += align_up((size_t)3 * N * N * sizeof(cf));
    return t;
}
extern "C" size_t xl_czt_workspace_bytes(int N, int Mx, int My, int vectorial) { return czt_ws_bytes(N, Mx, My, vectorial ? 3 : 1); }
extern "C" size_t xl_highna_workspace_bytes(int N, int Mx, int My) { return czt_ws_bytes(N, Mx, My, 3); }

static int czt_plan(CztPlan& pl, int N, int Mx, int My, int ncomp, void* ws, size_t ws_bytes) {
    if (N < 2 || Mx < 2 || My < 2) return xl_fail(XL_E_BAD_ARG, "czt: sizes must be >= 2%s", "");
    pl.N = N; pl.Mx = Mx; pl.My = My; pl.ncomp = ncomp;
    pl.Ly = xl_czt_padded_length(N, My);
    pl.Lx = xl_czt_padded_length(N, Mx);
    if (!pl.Ly || !pl.Lx) return xl_fail(XL_E_UNSUPPORTED, "czt: m+M-1 is a power of two or padded length outside [32,4096]%s", "");
    if (!ws || ws_bytes < czt_ws_bytes(N, Mx, My, ncomp)) return xl_fail(XL_E_WORKSPACE, "czt: workspace too small%s", "");
    Carver c{(char*)ws, 0, ws_bytes};
    pl.pre_y = (cf*)c.take((size_t)N * sizeof(cf));
    pl.pre_x = (cf*)c.take((size_t)N * sizeof(cf));
    pl.post_y = (cf*)c.take((size_t)My * sizeof(cf));
    pl.post_x = (cf*)c.take((size_t)Mx * sizeof(cf));
    pl.ft_y = (cf*)c.take((size_t)pl.Ly * sizeof(cf));
    pl.ftT_y = (cf*)c.take((size_t)pl.Ly * sizeof(cf));
    pl.ft_x = (cf*)c.take((size_t)pl.Lx * sizeof(cf));
    pl.ftT_x = (cf*)c.take((size_t)pl.Lx * sizeof(cf));
    pl.mid = (cf*)c.take((size_t)ncomp * N * My * sizeof(cf));
    pl.tmp3 = (cf*)c.take((size_t)3 * N * N * sizeof(cf));
    return XL_OK;
}

static int czt_setup(const CztPlan& pl, const double* z, double lambda_over_dx, double Dm_static,
                     double xout0, double xoutl, double yout0, double youtl, const cf* tw, xl_stream_t st) {
    int rc;
    XlCztSetup2Params sp;
    memset(&sp, 0, sizeof(sp));
    for (int ax = 0; ax < 2; ++ax) {
        XlCztSetupParams& s = sp.a[ax];
        s.z = z; s.lambda_over_dx = lambda_over_dx; s.Dm_static = Dm_static; s.tw = tw;
        s.m = pl.N;
        if (ax == 0) {   // y axis (first Bluestein pass, wave_optics.py:349)
            s.L = pl.Ly; s.M = pl.My; s.out0 = yout0; s.outl = youtl;
            s.pre = pl.pre_y; s.post = pl.post_y; s.ft = pl.ft_y; s.ftT = pl.ftT_y;
        } else {         // x axis (second pass, :352)
            s.L = pl.Lx; s.M = pl.Mx; s.out0 = xout0; s.outl = xoutl;
            s.pre = pl.pre_x; s.post = pl.post_x; s.ft = pl.ft_x; s.ftT = pl.ftT_x;
        }
    }
    if (pl.Ly == pl.Lx) {   // both axes in one launch (two CTAs)
        XL_FOR_L(pl.Ly, rc = xl_launch<XlCztSetup<XL>>(XlDim{2, 1}, st, sp));
        return rc;
    }
    XL_FOR_L(pl.Ly, rc = xl_launch<XlCztSetup<XL>>(XlDim{1, 1}, st, sp));
    if (rc) return rc;
    sp.a[0] = sp.a[1];
    XL_FOR_L(pl.Lx, rc = xl_launch<XlCztSetup<XL>>(XlDim{1, 1}, st, sp));
    return rc;
}

struct CztCall {
    int N, Mx, My, mode;  // mode: 0 scalar CZT, 1 VCZT, 2 high-NA
    const double* z; double lambda, k;
    double x0, dx, y0, dy, xout0, xoutl, yout0, youtl;
    double R, f, s2;
    int flags;
};

static void czt_common_params(XlCztParams& a, const CztCall& cc, const cf* tw) {
    memset(&a, 0, sizeof(a));
    a.tw = tw; a.z = cc.z; a.k = cc.k;
    a.lens_R = cc.R; a.lens_f = cc.f; a.lens_s2 = cc.s2;
    a.epi_cr = 1.0; a.epi_ci = 0.0;
}
static void czt_out_const(XlCztParams& a, const CztCall& cc) {
    if (cc.mode == 2) { a.epi_cr = 0.0; a.epi_ci = -cc.s2 / (cc.f * cc.lambda); a.epi_times_z = 0; }   // optical_elements.py:627
    else { a.epi_cr = cc.dx * cc.dy * cc.lambda; a.epi_ci = 0.0; a.epi_times_z = 1; }                   // wave_optics.py:355
}

template <int PRO, int EPI, int ACC> static int czt_axis_launch_t(const XlCztParams& a, XlDim grid, xl_stream_t st) {
    int rc;
#ifdef XL_EXP_CZT_PERSIST
    if constexpr ((PRO == XL_PRO_NONE || PRO == XL_PRO_RSF) && ACC != XL_ACC_GENERIC) {
        int slots = 3;      // the emulation uses 3 CTAs so that every CTA walks several items
#ifndef XL_HOST_EMU
        int dev = 0, sms = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        slots = 2 * (sms > 0 ? sms : 148);
#endif
        XlCztParams q = a;
        q.pairs = grid.x;
        const int items = grid.x * grid.y;
        XL_FOR_L(a.L, rc = xl_launch<XlCztAxisPersist<XL, PRO, EPI, ACC>>(XlDim{items < slots ? items : slots, 1}, st, q));
        return rc;
    }
#endif
    XL_FOR_L(a.L, rc = xl_launch<XlCztAxis<XL, PRO, EPI, ACC>>(grid, st, a));
    return rc;
}
static bool aligned16(const void* p) { return ((uintptr_t)p & 15) == 0; }
// The (prologue, epilogue, access shape) combinations the forward and adjoint chains use, each compiled branch-free.
// Paired 16-byte accesses need an even number of lines and even strides, and the paired variants are compiled with the
// zero-padded / discarded halves pruned (XlCztOp): other sizes take the generic variant.
static int czt_axis_launch(const XlCztParams& a, XlDim grid, xl_stream_t st) {
    const bool even = a.nlines % 2 == 0, in_lo = a.m_in <= a.L / 2;
    const bool out_lo = a.out_off == 0 && a.m_out <= a.L / 2;
    const bool pin = even && in_lo && out_lo && a.in_line == 1 && a.in_pos % 2 == 0 && a.in_comp % 2 == 0 && aligned16(a.in);
    const bool pout = even && in_lo && out_lo && a.out_line == 1 && a.out_pos % 2 == 0 &&
                      a.out_comp % 2 == 0 && aligned16(a.out);
#define XL_CZT_CASE(P, E)                                                                               \
    if (a.pro == P && a.epi == E) {                                                                     \
        if (pin) return czt_axis_launch_t<P, E, XL_ACC_PAIR_IN>(a, grid, st);                           \
        if (pout) return czt_axis_launch_t<P, E, XL_ACC_PAIR_OUT>(a, grid, st);                         \
        return czt_axis_launch_t<P, E, XL_ACC_GENERIC>(a, grid, st);                                    \
    }
    XL_CZT_CASE(XL_PRO_NONE, XL_EPI_NONE)
    XL_CZT_CASE(XL_PRO_NONE, XL_EPI_RSF)
    XL_CZT_CASE(XL_PRO_RSF, XL_EPI_NONE)
#undef XL_CZT_CASE
    // the vectorial prologues only occur in the forward chain (column-direction input)
    if (a.pro == XL_PRO_VCZT && a.epi == XL_EPI_NONE)
        return pin ? czt_axis_launch_t<XL_PRO_VCZT, XL_EPI_NONE, XL_ACC_PAIR_IN>(a, grid, st)
                   : czt_axis_launch_t<XL_PRO_VCZT, XL_EPI_NONE, XL_ACC_GENERIC>(a, grid, st);
    if (a.pro == XL_PRO_HIGHNA && a.epi == XL_EPI_NONE)
        return pin ? czt_axis_launch_t<XL_PRO_HIGHNA, XL_EPI_NONE, XL_ACC_PAIR_IN>(a, grid, st)
                   : czt_axis_launch_t<XL_PRO_HIGHNA, XL_EPI_NONE, XL_ACC_GENERIC>(a, grid, st);
    return xl_fail(XL_E_BAD_ARG, "czt: unsupported prologue/epilogue combination%s", "");
}

static int czt_forward(const CztCall& cc, const void* in, void* out, void* ws, size_t ws_bytes, xl_stream_t st) {
    if (!in || !out) return xl_fail(XL_E_BAD_ARG, "czt_fwd: null pointer%s", "");
    if (cc.mode != 2 && !cc.z) return xl_fail(XL_E_BAD_ARG, "czt_fwd: null z%s", "");
    const int ncomp = cc.mode == 0 ? 1 : 3;
    CztPlan pl;
    int rc = czt_plan(pl, cc.N, cc.Mx, cc.My, ncomp, ws, ws_bytes);
    if (rc) return rc;
    const cf* tw = xl_twiddles();
    if (!tw) return xl_fail(XL_E_CUDA, "twiddle table allocation failed%s", "");
    const double Dm_static = cc.mode == 2 ? cc.f * cc.lambda * (cc.N - 1) / (2 * cc.R) : 0.0;  // optical_elements.py:663
    rc = czt_setup(pl, cc.mode == 2 ? 0 : cc.z, cc.lambda / cc.dx, Dm_static, cc.xout0, cc.xoutl, cc.yout0, cc.youtl, tw, st);
    if (rc) return rc;
    const int N = cc.N, Mx = cc.Mx, My = cc.My;
    const double dxo = (cc.xoutl - cc.xout0) / (Mx - 1), dyo = (cc.youtl - cc.yout0) / (My - 1);
    // pass 1: Bluestein along y for every input column
    XlCztParams a;
    czt_common_params(a, cc, tw);
    a.L = pl.Ly; a.nlines = N; a.ncomp = ncomp; a.m_in = N; a.out_off = 0; a.m_out = My;
    a.in = (const cf*)in; a.in_line = 1; a.in_pos = N; a.in_comp = (long long)N * N;
    a.out = pl.mid; a.out_line = My; a.out_pos = 1; a.out_comp = (long long)N * My;
    a.pre = pl.pre_y; a.ft = pl.ft_y; a.post = pl.post_y;
    a.pro = cc.mode == 0 ? XL_PRO_RSF : (cc.mode == 1 ? XL_PRO_VCZT : XL_PRO_HIGHNA);
    a.gpro = XlGridFactor{cc.x0, cc.dx, cc.y0, cc.dy, 0};
    a.epi = XL_EPI_NONE;
    // pass 2: Bluestein along x for every column of the intermediate
    XlCztParams b;
    czt_common_params(b, cc, tw);
    b.L = pl.Lx; b.nlines = My; b.ncomp = ncomp; b.m_in = N; b.out_off = 0; b.m_out = Mx;
    b.in = pl.mid; b.in_line = 1; b.in_pos = My; b.in_comp = (long long)N * My;
    b.out = (cf*)out; b.out_line = Mx; b.out_pos = 1; b.out_comp = (long long)My * Mx;
    b.pre = pl.pre_x; b.ft = pl.ft_x; b.post = pl.post_x;
    b.pro = XL_PRO_NONE;
    b.epi = cc.mode == 2 ? XL_EPI_NONE : XL_EPI_RSF;
    b.gepi = XlGridFactor{cc.xout0, dxo, cc.yout0, dyo, 1};
    czt_out_const(b, cc);
    b.flags = cc.flags & XL_CONJ_OUT;
    rc = czt_axis_launch(a, XlDim{xl_groups(N), ncomp}, st);
    if (rc) return rc;
    return czt_axis_launch(b, XlDim{xl_groups(My), ncomp}, st);
}

static int czt_backward(const CztCall& cc, const void* ct_out, void* ct_in, void* ws, size_t ws_bytes, xl_stream_t st) {
    if (!ct_out || !ct_in) return xl_fail(XL_E_BAD_ARG, "czt_bwd: null pointer%s", "");
    if (cc.mode != 2 && !cc.z) return xl_fail(XL_E_BAD_ARG, "czt_bwd: null z%s", "");
    const int ncomp = cc.mode == 0 ? 1 : 3;
    CztPlan pl;
    int rc = czt_plan(pl, cc.N, cc.Mx, cc.My, ncomp, ws, ws_bytes);
    if (rc) return rc;
    const cf* tw = xl_twiddles();
    if (!tw) return xl_fail(XL_E_CUDA, "twiddle table allocation failed%s", "");
    const double Dm_static = cc.mode == 2 ? cc.f * cc.lambda * (cc.N - 1) / (2 * cc.R) : 0.0;
    rc = czt_setup(pl, cc.mode == 2 ? 0 : cc.z, cc.lambda / cc.dx, Dm_static, cc.xout0, cc.xoutl, cc.yout0, cc.youtl, tw, st);
    if (rc) return rc;
    const int N = cc.N, Mx = cc.Mx, My = cc.My;
    const double dxo = (cc.xoutl - cc.xout0) / (Mx - 1), dyo = (cc.youtl - cc.yout0) / (My - 1);
    // transpose of pass 2: rows of ct_out (length Mx) -> columns of the intermediate cotangent
    XlCztParams b;
    czt_common_params(b, cc, tw);
    b.L = pl.Lx; b.nlines = My; b.ncomp = ncomp; b.m_in = Mx; b.out_off = 0; b.m_out = N;
    b.in = (const cf*)ct_out; b.in_line = Mx; b.in_pos = 1; b.in_comp = (long long)My * Mx;
    b.out = pl.mid; b.out_line = 1; b.out_pos = My; b.out_comp = (long long)N * My;
    b.pre = pl.post_x; b.ft = pl.ftT_x; b.post = pl.pre_x;
    b.pro = cc.mode == 2 ? XL_PRO_NONE : XL_PRO_RSF;
    b.gpro = XlGridFactor{cc.xout0, dxo, cc.yout0, dyo, 1};
    b.epi = XL_EPI_NONE;
    czt_out_const(b, cc);
    b.flags = cc.flags & XL_CONJ_IN;
    // transpose of pass 1: rows of the intermediate cotangent (length My) -> columns of ct_field
    XlCztParams a;
    czt_common_params(a, cc, tw);
    a.L = pl.Ly; a.nlines = N; a.ncomp = ncomp; a.m_in = My; a.out_off = 0; a.m_out = N;
    a.in = pl.mid; a.in_line = My; a.in_pos = 1; a.in_comp = (long long)N * My;
    a.out = cc.mode == 0 ? (cf*)ct_in : pl.tmp3; a.out_line = 1; a.out_pos = N; a.out_comp = (long long)N * N;
    a.pre = pl.post_y; a.ft = pl.ftT_y; a.post = pl.pre_y;
    a.pro = XL_PRO_NONE;
    a.epi = cc.mode == 2 ? XL_EPI_NONE : XL_EPI_RSF;
    a.gepi = XlGridFactor{cc.x0, cc.dx, cc.y0, cc.dy, 0};
    a.flags = cc.mode == 0 ? (cc.flags & XL_CONJ_OUT) : 0;
    rc = czt_axis_launch(b, XlDim{xl_groups(My), ncomp}, st);
    if (rc) return rc;
    rc = czt_axis_launch(a, XlDim{xl_groups(N), ncomp}, st);
    if (rc || cc.mode == 0) return rc;
    XlFoldParams f;
    memset(&f, 0, sizeof(f));
    f.N = N; f.mode = cc.mode == 1 ? XL_FOLD_VCZT : XL_FOLD_HIGHNA; f.flags = cc.flags & XL_CONJ_OUT;
    f.t = pl.tmp3; f.gx = (cf*)ct_in; f.gy = (cf*)ct_in + (size_t)N * N;
    f.z = cc.mode == 1 ? cc.z : 0; f.x0 = cc.x0; f.y0 = cc.y0; f.dx = cc.dx; f.dy = cc.dy;
    f.lens_R = cc.R; f.lens_f = cc.f; f.lens_s2 = cc.s2;
    const size_t NN = (size_t)N * N;
    return xl_launch<XlFold>(XlDim{(int)((NN + XlFold::NT - 1) / XlFold::NT), 1}, st, f);
}

static CztCall make_czt_call(int mode, const double* z, double lambda, int N, int Mx, int My,
                             double x0, double dx, double y0, double dy, double xout0, double xoutl, double yout0, double youtl,
                             double R, double f, int flags) {
    CztCall c;
    memset(&c, 0, sizeof(c));
    c.N = N; c.Mx = Mx; c.My = My; c.mode = mode; c.z = z; c.lambda = lambda; c.k = 2.0 * M_PI / lambda;
    c.x0 = x0; c.dx = dx; c.y0 = y0; c.dy = dy; c.xout0 = xout0; c.xoutl = xoutl; c.yout0 = yout0; c.youtl = youtl;
    c.R = R; c.f = f;
    if (mode == 2) { double st = R / sqrt(R * R + f * f); c.s2 = st * st; }  // optical_elements.py:528
    c.flags = flags;
    return c;
}

extern "C" int xl_czt_fwd(const void* in, void* out, const double* z, double lambda, int N, int Mx, int My, int vectorial,
                          double x0, double dx, double y0, double dy, double xout0, double xoutl, double yout0, double youtl,
                          int flags, void* ws, size_t ws_bytes, void* stream) {
    CztCall c = make_czt_call(vectorial ? 1 : 0, z, lambda, N, Mx, My, x0, dx, y0, dy, xout0, xoutl, yout0, youtl, 0, 0, flags);
    return czt_forward(c, in, out, ws, ws_bytes, (xl_stream_t)stream);
}
extern "C" int xl_czt_bwd(const void* ct_out, void* ct_in, const double* z, double lambda, int N, int Mx, int My, int vectorial,
                          double x0, double dx, double y0, double dy, double xout0, double xoutl, double yout0, double youtl,
                          int flags, void* ws, size_t ws_bytes, void* stream) {
    CztCall c = make_czt_call(vectorial ? 1 : 0, z, lambda, N, Mx, My, x0, dx, y0, dy, xout0, xoutl, yout0, youtl, 0, 0, flags);
    return czt_backward(c, ct_out, ct_in, ws, ws_bytes, (xl_stream_t)stream);
}
extern "C" int xl_highna_fwd(const void* exy, void* out, int N, int Mx, int My, double radius, double f, double lambda,
                             double x0, double dx, double y0, double dy, double xout0, double xoutl, double yout0, double youtl,
                             int flags, void* ws, size_t ws_bytes, void* stream) {
    CztCall c = make_czt_call(2, 0, lambda, N, Mx, My, x0, dx, y0, dy, xout0, xoutl, yout0, youtl, radius, f, flags);
    return czt_forward(c, exy, out, ws, ws_bytes, (xl_stream_t)stream);
}
extern "C" int xl_highna_bwd(const void* ct_out, void* ct_exy, int N, int Mx, int My, double radius, double f, double lambda,
                             double x0, double dx, double y0, double dy, double xout0, double xoutl, double yout0, double youtl,
                             int flags, void* ws, size_t ws_bytes, void* stream) {
    CztCall c = make_czt_call(2, 0, lambda, N, Mx, My, x0, dx, y0, dy, xout0, xoutl, yout0, youtl, radius, f, flags);
    return czt_backward(c, ct_out, ct_exy, ws, ws_bytes, (xl_stream_t)stream);
}
