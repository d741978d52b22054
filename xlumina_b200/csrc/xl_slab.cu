// xl_slab.cu -- C ABI of the slab-decomposed / large-grid RS path (include/xlprop.h) over xl_kernels.cuh and xl_long.cuh.
#include "xl_host.h"
#include "xl_long.cuh"

// ================================================================================================ slab-decomposed RS
// Stage-level entry points of the multi-GPU RS path (BASELINE.json cfg 5): the N x N field is split
// into row slabs, one per rank; the all-to-all transposes between the stages are the caller's (NCCL via torch.distributed
// in xlumina_b200/slab.py).  Geometry for G ranks, P = xl_slab_padded_length(N):  rows = N/G field rows per rank (even),
// pairs = (P/2)/G x-slot pairs per rank, hrows = xl_slab_h_rows_per_rank(N, G) rows of the y >= 0 half of the impulse
// response per rank.   exchanged layout of a spectra buffer:  [source rank][pairs][rows of that rank][2].
// P <= 4096 runs the single-pass kernels of xl_kernels.cuh; longer lines (up to 32768: N <= 16384) run the split kernels of
// xl_long.cuh (P = R * L0) and need the scratch buffer of xl_slab_scratch_bytes().
// The inverse radix-R step of the split kernels: inside a thread-block cluster over distributed shared memory, or as a second
// launch through the scratch buffer.  Measured on B200 (profiles/long_probe_r02v.txt): clusters of R = 4 CTAs win (8192^2:
// long_cols 1.67 against 1.15 + 0.73 ms, long_rows_inv 0.82 against 0.59 + 0.40 ms), clusters of R = 8 lose (16384^2: 7.70
// against 4.75 + 2.13 ms -- eight 74 KB CTAs must find four SMs of one GPC and wait for each other at the barrier).
// -1 (default): clusters for R <= 4;  0 / 1: force the two-launch / the cluster form (tests, A/B timing).
static int g_long_cluster = -1;
extern "C" void xl_debug_set_long_cluster(int on) { g_long_cluster = on < 0 ? -1 : (on ? 1 : 0); }
static bool long_use_cluster(int R) { return g_long_cluster < 0 ? R <= 4 : g_long_cluster != 0; }
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
    size_t a = (size_t)g.rows * g.P, b = (size_t)g.pairs * g.P * 2, c = (size_t)g.hrows * (g.P / 2 + g.L0);
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
    q.chunk_magic = xl_div_magic(q.chunk_rows);
    q.tw = xl_twiddles();
    if (!q.tw) return xl_fail(XL_E_CUDA, "twiddle table allocation failed%s", "");
    return XL_OK;
}
static void long_set_chunk(XlLongParams& q, int chunk_rows) { q.chunk_rows = chunk_rows; q.chunk_magic = xl_div_magic(chunk_rows); }

// row spectra of this rank's y rows [rank*hrows, (rank+1)*hrows) of the impulse response: R[P/2][hrows][2]
// deriv: the rows of the REDUCED z-derivative h_z - i k h (xl_rs_h, xl_kernels.cuh) instead of h -- even in x and y like h
static int slab_h_rows(void* Rb, const double* z, int N, int G, int rank, double dx, double dy, double k,
                       void* scratch, void* stream, int deriv) {
    if (!Rb || !z) return xl_fail(XL_E_BAD_ARG, "xl_slab_h_rows: null pointer%s", "");
    SlabGeo g;
    int rc = slab_geo(g, N, G);
    if (rc) return rc;
    xl_stream_t st = (xl_stream_t)stream;
    if (g.R == 1) {
        XlRsParams p;
        if ((rc = rs_base_params(p, N, dx, dy, k))) return rc;
        p.H = (cf*)Rb; p.z = z; p.flags = deriv ? XL_F_DERIV : 0;
        p.rows = g.hrows; p.hrow0 = rank * g.hrows; p.hstore_all = 1;
        const int L = p.L;
        XL_FOR_L(L, rc = xl_launch<XlHRows<XL>>(XlDim{xl_groups(g.hrows), 1}, st, p));
        return rc;
    }
    if (!scratch) return xl_fail(XL_E_BAD_ARG, "xl_slab_h_rows: scratch needed for padded lengths > 4096%s", "");
    XlLongParams q;
    if ((rc = long_params(q, g, N, dx, dy, k))) return rc;
    q.z = z; q.hrow0 = rank * g.hrows; q.hrows = g.hrows; q.scratch = (cf*)scratch; q.spec = (cf*)Rb;
    q.flags = deriv ? XL_F_DERIV : 0;
    rc = xl_launch<XlHEval>(XlDim{pointwise_grid((size_t)g.L0 / 2 + 1, XlHEval::NT), g.hrows}, st, q);
    if (rc) return rc;
    XL_FOR_L0(g.L0, rc = xl_launch<XlLongHRows<XL>>(XlDim{g.R / 2 + 1, xl_groups(g.hrows)}, st, q));
    return rc;
}
extern "C" int xl_slab_h_rows(void* Rb, const double* z, int N, int G, int rank, double dx, double dy, double k,
                              void* scratch, void* stream) {
    return slab_h_rows(Rb, z, N, G, rank, dx, dy, k, scratch, stream, 0);
}
extern "C" int xl_slab_h_rows_dz(void* Rb, const double* z, int N, int G, int rank, double dx, double dy, double k,
                                 void* scratch, void* stream) {
    return slab_h_rows(Rb, z, N, G, rank, dx, dy, k, scratch, stream, 1);
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
    q.spec = (cf*)Th; q.H = (cf*)Hloc; long_set_chunk(q, g.hrows);
    rc = xl_launch<XlLongHSplit>(XlDim{pointwise_grid((size_t)g.L0, XlLongHSplit::NT), g.pairs}, st, q);
    if (rc) return rc;
    XL_FOR_L0(g.L0, rc = xl_launch<XlLongHCols<XL>>(XlDim{g.R / 2 + 1, g.pairs}, st, q));
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
    rc = xl_launch<XlLongColsSplit>(XlDim{pointwise_grid((size_t)g.L0, XlLongColsSplit::NT), g.pairs}, st, q);
    if (rc) return rc;
    if (long_use_cluster(g.R)) {      // convolution + radix-R DIT step in one cluster kernel (no scratch round trip)
        XL_FOR_L0(g.L0, rc = xl_launch_cluster<XlLongColsC<XL>>(XlDim{g.R, g.pairs}, st, q));
        return rc;
    }
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
    if (long_use_cluster(g.R)) {
        XL_FOR_L0(g.L0, rc = xl_launch_cluster<XlLongRowsInvC<XL>>(XlDim{g.R, xl_groups(g.rows)}, st, q));
        return rc;
    }
    XL_FOR_L0(g.L0, rc = xl_launch<XlLongRowsInv<XL>>(XlDim{g.R, xl_groups(g.rows)}, st, q));
    if (rc) return rc;
    return xl_launch<XlLongRowsCombine>(XlDim{pointwise_grid((size_t)g.rows * g.L0, XlLongRowsCombine::NT), 1}, st, q);
}
