// xl_long.cuh -- lines longer than one CTA's shared memory (padded length P = R * L0, R in {2,4,8}, L0 = 4096 in
// production): the kernels of the slab-decomposed RS path for grids up to 16384^2 (BASELINE.json cfg 5).
//
// One decimation step splits a length-P DFT into R DFTs of length L0 that the shared-memory engine (xl_fft.cuh) runs:
//   forward   X[q + R k] = FFT_L0( y_q )[k],   y_q[i] = w_P^{i q} * sum_{j<R} x[i + L0 j] w_R^{j q}       (DIF; i < L0, q < R)
//   inverse   x[i + L0 j] = sum_q w_R^{-j q} z_q[i],   z_q[i] = w_P^{-i q} * IDFT_L0( X_q )[i]                (DIT)
// The forward radix-R step needs only the CTA's own inputs, so it is fused into the first pass's LOAD (R, or R/2 when
// the upper half of the line is zero padding, strided global loads per element -- they hit L2, the line was just read by
// the R-1 sibling CTAs).  The inverse radix-R step needs the R sub-lines of R different CTAs: they park z_q in a scratch
// buffer and a pointwise "combine" kernel finishes the line (and drops the discarded half).
// "Super-slot" order of a length-P spectrum: bin q + R k lives at q*L0 + slot_L0(k); adjacent super-slots pair up exactly
// as in the short case, so the blocked layouts, the exchanged layouts of the slab path and the transfer-function layout
// [pair][P][2] carry over (no x/y mirroring of the transfer function in this mode).
#pragma once
#include "xl_kernels.cuh"

struct XlLongParams {
    int N, P, R, L0;        // samples per line, padded length, split factor, sub-line length (P == R * L0)
    int rows;               // row kernels: rows of this launch == row count of the blocked layout
    int chunk_rows;         // column kernels: rows per source rank in the exchanged layout [rank][pairs][chunk_rows][2]
    int pairs;              // column kernels: slot pairs owned by this rank (== gridDim.y; blockIdx.x is the sub-line)
    int flags;
    const cf* in; cf* out;  // field rows [rows][N]
    cf* spec;               // row side: [P/2][rows][2]; column side: exchanged layout
    cf* H;                  // transfer-function slab [pairs][P][2]  (h kernels: destination)
    cf* scratch;            // z_q parking: rows [rows][R][L0], columns [pairs][R][L0][2]; h_eval: plane [hrows][P/2+1]
    const cf* tw;           // master table tw[k] = exp(-2 pi i k / XL_TWN)
    const double* z;
    double dx, dy, k;
    float hscale;
    int hrow0, hrows;       // h kernels: first global y row of this rank, rows per rank
};

// w_den^{num} = exp(-2 pi i num/den), den a power of two <= XL_TWN, num >= 0
XL_DEV cf xl_tw_at(const cf* tw, int num, int den) { return xl_ldg(tw + (size_t)(num & (den - 1)) * (XL_TWN / den)); }

// ------------------------------------------------------------------------------------------------ row kernels
template <int L0> struct XlLongRowsFwdOp : XlOpBase {
    const XlLongParams& p; int q, yb;
    XL_DEV void load(int i, cf* v, int stride) const {
        const cf wiq = xl_tw_at(p.tw, i * q, p.P);
#pragma unroll
        for (int l = 0; l < XL_V; ++l) {
            const int y = yb + l;
            cf acc = cf_zero();
#pragma unroll
            for (int j = 0; j < 4; ++j) {          // x[n] = 0 for n >= N (N <= P/2): j < R/2 <= 4
                const int n = i + L0 * j;
                const bool ok = 2 * j < p.R && n < p.N && y < p.rows;
                cf x = p.in[ok ? (size_t)y * p.N + n : 0];
                if (p.flags & XL_F_CONJ_IN) x = cf_conj(x);
                if (ok) acc = j == 0 ? x : cf_fma(x, xl_tw_at(p.tw, j * q, p.R), acc);
            }
            v[l * stride] = cf_mul(acc, wiq);
        }
    }
    XL_DEV void spec(int beta, const cf* v) const {
#pragma unroll
        for (int qq = 0; qq < 16; ++qq) {
            const int g = q * L0 + qq * (L0 / 16) + beta;
            xl_blocked_store2<xl_lane_mask(L0)>(p.spec + (size_t)(g / 2) * p.rows * 2, yb, p.rows, g, v[qq], v[16 + qq]);
        }
    }
    XL_DEV void store_vec(int, const cf*) const {}
};
template <int L0> struct XlLongRowsFwd {
    static const char* name() { return "long_rows_fwd"; }
    typedef XlLongParams Params;
    static constexpr int NT = xl_threads(L0);
    static size_t smem() { return xl_smem_bytes(L0, XL_V); }
    XL_DEV static void run(const Params& p, cf* s) {
        cf* t = s + xl_tile_elems(L0, XL_V);
        XlFft<L0, XL_V>::init_tw(t, p.tw);
        XlLongRowsFwdOp<L0> op{{}, p, XL_BLOCK_X, XL_BLOCK_Y * XL_V};   // sub-line q varies fastest: the R CTAs that re-read one
                                                                         // row pair are neighbours in launch order (L2 hits)
        XlFft<L0, XL_V>::forward(s, t, op);
    }
};

template <int L0> struct XlLongRowsInvOp : XlOpBase {
    static constexpr int R1 = xl_first_radix(L0), S1 = L0 / R1;
    const XlLongParams& p; int q, yb;
    XL_DEV void load(int, cf*, int) const {}
    XL_DEV void spec(int beta, cf* v) const {
#pragma unroll
        for (int qq = 0; qq < 16; ++qq) {
            const int g = q * L0 + qq * (L0 / 16) + beta;
            xl_blocked_load2<xl_lane_mask(L0)>(p.spec + (size_t)(g / 2) * p.rows * 2, yb, p.rows, g, v + qq, v + 16 + qq);
        }
    }
    XL_DEV void store_vec(int n, const cf* v) const {
#pragma unroll
        for (int j = 0; j < R1; ++j) {
            const int i = n + S1 * j;
            const cf w = cf_conj(xl_tw_at(p.tw, i * q, p.P));     // w_P^{-i q}
#pragma unroll
            for (int l = 0; l < XL_V; ++l) {
                const int y = yb + l;
                if (y < p.rows) p.scratch[((size_t)y * p.R + q) * L0 + i] = cf_mul(v[l * R1 + j], w);
            }
        }
    }
};
template <int L0> struct XlLongRowsInv {
    static const char* name() { return "long_rows_inv"; }
    typedef XlLongParams Params;
    static constexpr int NT = xl_threads(L0);
    static size_t smem() { return xl_smem_bytes(L0, XL_V); }
    XL_DEV static void run(const Params& p, cf* s) {
        cf* t = s + xl_tile_elems(L0, XL_V);
        XlFft<L0, XL_V>::init_tw(t, p.tw);
        XlLongRowsInvOp<L0> op{{}, p, XL_BLOCK_X, XL_BLOCK_Y * XL_V};
        XlFft<L0, XL_V>::inverse(s, t, op);
    }
};
// out[y][i + L0 j] = sum_q w_R^{-j q} z_q[y][i]   for the kept samples (< N)
struct XlLongRowsCombine {
    static const char* name() { return "long_rows_combine"; }
    typedef XlLongParams Params;
    static constexpr int NT = 256;
    static size_t smem() { return 0; }
    XL_DEV static void run(const Params& p, cf*) {
        XL_THREADS(tid, NT) {
            const size_t idx = (size_t)XL_BLOCK_X * NT + tid;
            if (idx < (size_t)p.rows * p.L0) {
                const int y = (int)(idx / p.L0), i = (int)(idx % p.L0);
                cf zq[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) zq[q] = q < p.R ? p.scratch[((size_t)y * p.R + q) * p.L0 + i] : cf_zero();
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int n = i + p.L0 * j;
                    if (2 * j >= p.R || n >= p.N) continue;
                    cf acc = zq[0];
#pragma unroll
                    for (int q = 1; q < 8; ++q)
                        if (q < p.R) acc = cf_fma(zq[q], cf_conj(xl_tw_at(p.tw, j * q, p.R)), acc);
                    if (p.flags & XL_F_CONJ_OUT) acc = cf_conj(acc);
                    p.out[(size_t)y * p.N + n] = acc;
                }
            }
        }
    }
};

// ------------------------------------------------------------------------------------------------ column kernels
// rows of a column pair in the exchanged layout (chunk_rows == N on a single rank)
XL_DEV cf* xl_long_col_row(const XlLongParams& p, cf* tile, int i) {
    return tile + (size_t)(i / p.chunk_rows) * ((size_t)p.pairs * p.chunk_rows * XL_V) + (size_t)(i % p.chunk_rows) * XL_V;
}
template <int L0> struct XlLongColsOp : XlOpBase {
    static constexpr int R1 = xl_first_radix(L0), S1 = L0 / R1;
    const XlLongParams& p; int G, q; cf* tile; const cf* Ht;
    XL_DEV void load(int i, cf* v, int stride) const {
        cf a0 = cf_zero(), a1 = cf_zero();
#pragma unroll
        for (int j = 0; j < 4; ++j) {              // rows >= N are zero padding: j < R/2
            const int n = i + L0 * j;
            const bool ok = 2 * j < p.R && n < p.N;
            cf x0, x1;
            xl_ld4(xl_long_col_row(p, tile, ok ? n : 0), &x0, &x1);
            if (ok) {
                if (j == 0) { a0 = x0; a1 = x1; }
                else { const cf w = xl_tw_at(p.tw, j * q, p.R); a0 = cf_fma(x0, w, a0); a1 = cf_fma(x1, w, a1); }
            }
        }
        const cf wiq = xl_tw_at(p.tw, i * q, p.P);
        v[0] = cf_mul(a0, wiq);
        v[stride] = cf_mul(a1, wiq);
    }
    XL_DEV void spec(int beta, cf* v) const {
#pragma unroll
        for (int qq = 0; qq < 16; ++qq) {
            cf h0, h1;
            xl_ldg4(Ht + (size_t)(q * L0 + qq * (L0 / 16) + beta) * XL_V, &h0, &h1);
            v[qq] = cf_mul(v[qq], h0);
            v[16 + qq] = cf_mul(v[16 + qq], h1);
        }
    }
    XL_DEV void store_vec(int n, const cf* v) const {
        cf* zc = p.scratch + ((size_t)G * p.R + q) * L0 * XL_V;
#pragma unroll
        for (int j = 0; j < R1; ++j) {
            const int i = n + S1 * j;
            const cf w = cf_conj(xl_tw_at(p.tw, i * q, p.P));
            xl_st4(zc + (size_t)i * XL_V, cf_mul(v[j], w), cf_mul(v[R1 + j], w));
        }
    }
};
template <int L0> struct XlLongCols {
    static const char* name() { return "long_cols"; }
    typedef XlLongParams Params;
    static constexpr int NT = xl_threads(L0);
    static size_t smem() { return xl_smem_bytes(L0, XL_V); }
    XL_DEV static void run(const Params& p, cf* s) {
        cf* t = s + xl_tile_elems(L0, XL_V);
        XlFft<L0, XL_V>::init_tw(t, p.tw);
        const int G = XL_BLOCK_Y;   // sub-line on blockIdx.x: the R CTAs of one column pair run back to back
        XlLongColsOp<L0> op{{}, p, G, XL_BLOCK_X, p.spec + (size_t)G * p.chunk_rows * XL_V, p.H + (size_t)G * p.P * XL_V};
        XlFft<L0, XL_V>::conv(s, t, op);
    }
};
// column tile rows i + L0 j (< N)  <-  sum_q w_R^{-j q} z_q[i]
struct XlLongColsCombine {
    static const char* name() { return "long_cols_combine"; }
    typedef XlLongParams Params;
    static constexpr int NT = 256;
    static size_t smem() { return 0; }
    XL_DEV static void run(const Params& p, cf*) {
        XL_THREADS(tid, NT) {
            const size_t idx = (size_t)XL_BLOCK_X * NT + tid;
            if (idx < (size_t)p.pairs * p.L0) {
                const int G = (int)(idx / p.L0), i = (int)(idx % p.L0);
                cf z0[8], z1[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    z0[q] = cf_zero(); z1[q] = cf_zero();
                    if (q < p.R) xl_ld4(p.scratch + (((size_t)G * p.R + q) * p.L0 + i) * XL_V, &z0[q], &z1[q]);
                }
                cf* tile = p.spec + (size_t)G * p.chunk_rows * XL_V;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int n = i + p.L0 * j;
                    if (2 * j >= p.R || n >= p.N) continue;
                    cf a0 = z0[0], a1 = z1[0];
#pragma unroll
                    for (int q = 1; q < 8; ++q)
                        if (q < p.R) { const cf w = cf_conj(xl_tw_at(p.tw, j * q, p.R)); a0 = cf_fma(z0[q], w, a0); a1 = cf_fma(z1[q], w, a1); }
                    xl_st4(xl_long_col_row(p, tile, n), a0, a1);
                }
            }
        }
    }
};

// ------------------------------------------------------------------------------------------------ transfer function
// samples of the impulse response on this rank's y rows, x in [0, P/2] (h is even in x): plane[hrows][P/2+1]
struct XlHEval {
    static const char* name() { return "h_eval"; }
    typedef XlLongParams Params;
    static constexpr int NT = 256;
    static size_t smem() { return 0; }
    XL_DEV static void run(const Params& p, cf*) {
        const XlRsHConst hc = xl_rs_hconst(xl_ldg(p.z), p.k);
        const int W = p.P / 2 + 1;
        XL_THREADS(tid, NT) {
            const size_t idx = (size_t)XL_BLOCK_X * NT + tid;
            if (idx < (size_t)p.hrows * W) {
                const int yl = (int)(idx / W), xi = (int)(idx % W), y = p.hrow0 + yl;
                p.scratch[idx] = y <= p.P / 2 ? xl_rs_h(xi * p.dx, y * p.dy, hc, 0) : cf_zero();
            }
        }
    }
};
// row spectra of the impulse response: all R input blocks are populated (the wrapped kernel fills the whole line)
template <int L0> struct XlLongHRowsOp : XlOpBase {
    const XlLongParams& p; int q, yb;
    XL_DEV void load(int i, cf* v, int stride) const {
        const int W = p.P / 2 + 1;
        const cf wiq = xl_tw_at(p.tw, i * q, p.P);
#pragma unroll
        for (int l = 0; l < XL_V; ++l) {
            const int yl = yb + l;
            const bool rowok = yl < p.hrows;
            cf acc = cf_zero();
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int n = i + L0 * j;
                const bool ok = j < p.R && rowok;
                const int xi = n <= p.P / 2 ? n : p.P - n;
                const cf x = p.scratch[ok ? (size_t)yl * W + xi : 0];
                if (ok) acc = j == 0 ? x : cf_fma(x, xl_tw_at(p.tw, j * q, p.R), acc);
            }
            v[l * stride] = cf_mul(acc, wiq);
        }
    }
    XL_DEV void spec(int beta, const cf* v) const {
#pragma unroll
        for (int qq = 0; qq < 16; ++qq) {
            const int g = q * L0 + qq * (L0 / 16) + beta;
            xl_blocked_store2<xl_lane_mask(L0)>(p.spec + (size_t)(g / 2) * p.hrows * 2, yb, p.hrows, g, v[qq], v[16 + qq]);
        }
    }
    XL_DEV void store_vec(int, const cf*) const {}
};
template <int L0> struct XlLongHRows {
    static const char* name() { return "long_h_rows"; }
    typedef XlLongParams Params;
    static constexpr int NT = xl_threads(L0);
    static size_t smem() { return xl_smem_bytes(L0, XL_V); }
    XL_DEV static void run(const Params& p, cf* s) {
        cf* t = s + xl_tile_elems(L0, XL_V);
        XlFft<L0, XL_V>::init_tw(t, p.tw);
        XlLongHRowsOp<L0> op{{}, p, XL_BLOCK_X, XL_BLOCK_Y * XL_V};
        XlFft<L0, XL_V>::forward(s, t, op);
    }
};
// column spectra of the impulse-response row spectra (even in y: row P - y == row y), exchanged layout in, H slab out
template <int L0> struct XlLongHColsOp : XlOpBase {
    const XlLongParams& p; int q; const cf* src; cf* Ht;
    XL_DEV void load(int i, cf* v, int stride) const {
        cf a0 = cf_zero(), a1 = cf_zero();
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int n = i + L0 * j;
            const bool ok = j < p.R;
            const int r = n <= p.P / 2 ? n : p.P - n;
            cf x0, x1;
            xl_ld4(src + (size_t)((ok ? r : 0) / p.chunk_rows) * ((size_t)p.pairs * p.chunk_rows * XL_V) +
                       (size_t)((ok ? r : 0) % p.chunk_rows) * XL_V, &x0, &x1);
            if (ok) {
                if (j == 0) { a0 = x0; a1 = x1; }
                else { const cf w = xl_tw_at(p.tw, j * q, p.R); a0 = cf_fma(x0, w, a0); a1 = cf_fma(x1, w, a1); }
            }
        }
        const cf wiq = xl_tw_at(p.tw, i * q, p.P);
        v[0] = cf_mul(a0, wiq);
        v[stride] = cf_mul(a1, wiq);
    }
    XL_DEV void spec(int beta, const cf* v) const {
#pragma unroll
        for (int qq = 0; qq < 16; ++qq)
            xl_st4(Ht + (size_t)(q * L0 + qq * (L0 / 16) + beta) * XL_V, cf_scale(v[qq], p.hscale), cf_scale(v[16 + qq], p.hscale));
    }
    XL_DEV void store_vec(int, const cf*) const {}
};
template <int L0> struct XlLongHCols {
    static const char* name() { return "long_h_cols"; }
    typedef XlLongParams Params;
    static constexpr int NT = xl_threads(L0);
    static size_t smem() { return xl_smem_bytes(L0, XL_V); }
    XL_DEV static void run(const Params& p, cf* s) {
        cf* t = s + xl_tile_elems(L0, XL_V);
        XlFft<L0, XL_V>::init_tw(t, p.tw);
        const int G = XL_BLOCK_Y;
        XlLongHColsOp<L0> op{{}, p, XL_BLOCK_X, p.spec + (size_t)G * p.chunk_rows * XL_V, p.H + (size_t)G * p.P * XL_V};
        XlFft<L0, XL_V>::forward(s, t, op);
    }
};
