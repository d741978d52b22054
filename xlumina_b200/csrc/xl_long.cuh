// xl_long.cuh -- lines longer than one CTA's shared memory (padded length P = R * L0, R in {2,4,8}, L0 = 4096 in
// production): the kernels of the slab-decomposed RS path for grids up to 16384^2 (BASELINE.json cfg 5).
//
// One decimation step splits a length-P DFT into R DFTs of length L0 that the shared-memory engine (xl_fft.cuh) runs:
//   forward   X[q + R k] = FFT_L0( y_q )[k],   y_q[i] = w_P^{i q} * sum_{j<R} x[i + L0 j] w_R^{j q}       (DIF; i < L0, q < R)
//   inverse   x[i + L0 j] = sum_q w_R^{-j q} z_q[i],   z_q[i] = w_P^{-i q} * IDFT_L0( X_q )[i]                (DIT)
// Where the two radix-R steps run (round 2, measured per kernel with scripts/long_probe.py):
//   rows, forward      fused into the first pass's LOAD of long_rows_fwd (R/2 strided loads per element: the upper half of the
//                      line is zero padding; they hit L2, the line was just read by the R-1 sibling CTAs)
//   columns, forward   its own pointwise pass (long_cols_split, long_h_split): once per position instead of once per sub-line CTA
//   impulse response   inside h_eval, which evaluates every sample once (positions t and L0 - t share their R samples)
//   inverse (DIT)      needs the R sub-lines of R different CTAs: for R <= 4 they are a thread-block cluster and combine over
//                      distributed shared memory (XlLongColsC, XlLongRowsInvC); for R = 8 they park z_q in a scratch buffer
//                      and a pointwise "combine" kernel finishes the line (and drops the discarded half)
// "Super-slot" order of a length-P spectrum: bin q + R k lives at q*L0 + slot_L0(k); adjacent super-slots pair up exactly
// as in the short case, so the blocked layouts, the exchanged layouts of the slab path and the transfer-function layout
// [pair][P][2] carry over.  The transfer function is even in x and y; in this mode only the evenness along the transformed
// axis is used: sub-line R - q of a spectrum is sub-line q in reverse slot order (XlLongHRows), so the sub-lines q > R/2 are
// neither transformed nor (for the column spectra) stored.
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
    cf* H;                  // transfer-function slab [pairs][P][2], of which the sub-line blocks q <= R/2 are written / read
    cf* scratch;            // z_q parking: rows [rows][R][L0], columns [pairs][R][L0][2]; h_eval: Y[hrows][R/2+1][L0]
    const cf* tw;           // master table tw[k] = exp(-2 pi i k / XL_TWN)
    const double* z;
    double dx, dy, k;
    float hscale;
    int hrow0, hrows;       // h kernels: first global y row of this rank, rows per rank
    unsigned chunk_magic;   // ceil(2^32 / chunk_rows): n / chunk_rows == umulhi(n, chunk_magic) for n, chunk_rows in [2, 65536)
};
XL_HD inline unsigned xl_div_magic(int d) { return (unsigned)((0x100000000ull + (unsigned long long)d - 1) / (unsigned long long)d); }

// The first pass of every split kernel forms  y_q[i] = w_P^{i q} * sum_j x[i + L0 j] w_R^{j q}  from R (or R/2) strided global
// loads per position.  Round 2 (profiles/long_probe_r02q.txt): written with `if (ok)` guards, selected 64-bit addresses, run-time
// integer divisions and a twiddle-table lookup per (position, j), the loads of a thread came out of ptxas as ~100 dependent
// groups behind branches (68-113 BSSY regions per kernel) and the kernels were bound by that serialised L2 latency:
// long_rows_fwd 14.3 ms against 2.3 ms for long_rows_inv over the same 4.3 GB of spectra at 16384^2.  Now every address is a
// masked offset (no select, no branch), out-of-range samples are zeroed by a select on the DATA, w_R^{j q} (CTA-uniform) is
// hoisted into the functor, and n / chunk_rows is a multiply-high by a host-made reciprocal -- all loads of a thread's 16
// positions are independent and issue back to back.
// w_den^{num} = exp(-2 pi i num/den), den a power of two <= XL_TWN, num >= 0
XL_DEV cf xl_tw_at(const cf* tw, int num, int den) { return xl_ldg(tw + (size_t)(num & (den - 1)) * (XL_TWN / den)); }
struct XlLongWr { cf w[8]; };   // w_R^{j q}, j < 8 (entries >= R repeat w^0, never used with a non-zero operand)
XL_DEV XlLongWr xl_long_wr(const XlLongParams& p, int q) {
    XlLongWr r;
#pragma unroll
    for (int j = 0; j < 8; ++j) r.w[j] = xl_tw_at(p.tw, (j < p.R ? j : 0) * q, p.R);
    return r;
}
// x where ok, 0 elsewhere (a select on the loaded DATA; the address was masked); cj = -1 conjugates
XL_DEV cf xl_sel(bool ok, cf x, float cj = 1.f) { return make_float2(ok ? x.x : 0.f, ok ? x.y * cj : 0.f); }
// element offset of row n of a column pair in the exchanged layout [source rank][pairs][chunk_rows][2]; cstride = elements
// per source rank (chunk_rows == N on a single rank: offset n * 2)
XL_DEV size_t xl_long_col_off(const XlLongParams& p, size_t cstride, int n) {
    const unsigned c = xl_umulhi((unsigned)n, p.chunk_magic);
    return (size_t)c * cstride + (size_t)(n - (int)c * p.chunk_rows) * XL_V;
}

// ------------------------------------------------------------------------------------------------ row kernels
template <int L0> struct XlLongRowsFwdOp : XlOpBase {
    const XlLongParams& p; int q, yb; XlLongWr wr; float cj;   // cj: -1 conjugates the input (XL_F_CONJ_IN), +1 otherwise
    XL_DEV void load(int i, cf* v, int stride) const {
        const cf wiq = xl_tw_at(p.tw, i * q, p.P);
        cf x[XL_V][4];
        bool ok[XL_V][4];
#pragma unroll
        for (int l = 0; l < XL_V; ++l)
#pragma unroll
            for (int j = 0; j < 4; ++j) {          // x[n] = 0 for n >= N (N <= P/2): j < R/2 <= 4
                const int n = i + L0 * j;
                ok[l][j] = (2 * j < p.R) & (n < p.N) & (yb + l < p.rows);
                x[l][j] = p.in[((long long)(yb + l) * p.N + n) & -(long long)ok[l][j]];
            }
#pragma unroll
        for (int l = 0; l < XL_V; ++l) {
            cf acc = xl_sel(ok[l][0], x[l][0], cj);
#pragma unroll
            for (int j = 1; j < 4; ++j) acc = cf_fma(xl_sel(ok[l][j], x[l][j], cj), wr.w[j], acc);
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
        // sub-line q varies fastest: the R CTAs that re-read one row pair are neighbours in launch order (L2 hits)
        XlLongRowsFwdOp<L0> op{{}, p, XL_BLOCK_X, XL_BLOCK_Y * XL_V, xl_long_wr(p, XL_BLOCK_X), (p.flags & XL_F_CONJ_IN) ? -1.f : 1.f};
        XlFft<L0, XL_V>::forward(s, t, op);
    }
};

// Measured and dropped (profiles/long_probe_r02w.txt): the sub-lines q and q + R/2 of ONE row as the two lines of a CTA
// (w_R^{j(q+R/2)} = (-1)^j w_R^{jq}: one set of loads and products serves both, half the first-pass loads) -- 5.0 ms against
// 3.6 ms at 16384^2: each lane pair then writes 16 bytes of a 32-byte sector and the other half arrives from another CTA.

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
                for (int q = 0; q < 8; ++q)
                    zq[q] = xl_sel(q < p.R, p.scratch[(((size_t)y * p.R + q) * p.L0 + i) & -(size_t)(q < p.R)]);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int n = i + p.L0 * j;
                    if (2 * j >= p.R || n >= p.N) continue;
                    cf acc = zq[0];
#pragma unroll
                    for (int q = 1; q < 8; ++q) acc = cf_fma(zq[q], cf_conj(xl_tw_at(p.tw, j * q, p.R)), acc);   // zq[q >= R] == 0
                    if (p.flags & XL_F_CONJ_OUT) acc = cf_conj(acc);
                    p.out[(size_t)y * p.N + n] = acc;
                }
            }
        }
    }
};

// long_rows_inv + long_rows_combine as ONE cluster kernel: the R CTAs of a row pair are a thread-block cluster; each leaves
// w_P^{-i q} IDFT_L0(X_q)[i] of both rows in its own tile, and after the cluster barrier CTA c finishes the positions
// i in [c L0/R, (c+1) L0/R): it gathers z_q[i] from the R tiles over distributed shared memory (one 16-byte load per peer
// brings both rows) and writes the kept samples.  No scratch round trip (4.3 GB written + read at 16384^2), no second launch.
template <int L0> struct XlLongRowsInvZOp : XlOpBase {
    static constexpr int R1 = xl_first_radix(L0), S1 = L0 / R1;
    const XlLongParams& p; int q, yb; cf* s;
    XL_DEV void load(int, cf*, int) const {}
    XL_DEV void spec(int beta, cf* v) const {
#pragma unroll
        for (int qq = 0; qq < 16; ++qq) {
            const int g = q * L0 + qq * (L0 / 16) + beta;
            xl_blocked_load2<xl_lane_mask(L0)>(p.spec + (size_t)(g / 2) * p.rows * 2, yb, p.rows, g, v + qq, v + 16 + qq);
        }
    }
    // in place: the thread that owns n reads and writes the positions n + S1 j only
    XL_DEV void store_vec(int n, const cf* v) const {
#pragma unroll
        for (int j = 0; j < R1; ++j) {
            const int i = n + S1 * j;
            const cf w = cf_conj(xl_tw_at(p.tw, i * q, p.P));     // w_P^{-i q}
            const cf o[2] = {cf_mul(v[j], w), cf_mul(v[R1 + j], w)};
            XlTile<2>::st(s, i, o, 1);
        }
    }
};
template <int L0> struct XlLongRowsInvC {
    static const char* name() { return "long_rows_inv"; }
    typedef XlLongParams Params;
    static constexpr int NT = xl_threads(L0);
    static size_t smem() { return xl_smem_bytes(L0, XL_V); }
    XL_DEV static void run1(const Params& p, cf* s) {
        cf* t = s + xl_tile_elems(L0, XL_V);
        XlLongRowsInvZOp<L0> op{{}, p, XL_BLOCK_X, XL_BLOCK_Y * XL_V, s};
        XlFft<L0, XL_V>::inverse_g(s, t, p.tw, op);
    }
    XL_DEV static void run2(const Params& p, cf*, const XlPeers& pr) {
        const int c = XL_BLOCK_X, yb = XL_BLOCK_Y * XL_V, slice = L0 / p.R;
        XL_THREADS(tid, NT) {
            for (int ii = tid; ii < slice; ii += NT) {
                const int i = c * slice + ii;
                cf z0[8], z1[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    z0[q] = cf_zero(); z1[q] = cf_zero();
                    if (q < p.R) xl_peer_ld4(pr, q, 2 * xl_pad(i), &z0[q], &z1[q]);
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int n = i + L0 * j;
                    if (2 * j >= p.R || n >= p.N) continue;
                    cf a0 = z0[0], a1 = z1[0];
#pragma unroll
                    for (int q = 1; q < 8; ++q) {   // z[q >= R] == 0
                        const cf w = cf_conj(xl_tw_at(p.tw, j * q, p.R)); a0 = cf_fma(z0[q], w, a0); a1 = cf_fma(z1[q], w, a1);
                    }
                    if (p.flags & XL_F_CONJ_OUT) { a0 = cf_conj(a0); a1 = cf_conj(a1); }
                    if (yb < p.rows) p.out[(size_t)yb * p.N + n] = a0;
                    if (yb + 1 < p.rows) p.out[(size_t)(yb + 1) * p.N + n] = a1;
                }
            }
        }
    }
};

// ------------------------------------------------------------------------------------------------ column kernels
// Column kernels, three launches per stage:
//   long_cols_split   (pointwise)  Y[G][q][i] = w_P^{i q} sum_{j < R/2} x[i + L0 j] w_R^{j q}      the radix-R DIF step, ONCE per
//                                  position for all R sub-lines (scratch, [pairs][R][L0][2])
//   long_cols         (FFT)        z_q = IDFT_L0( DFT_L0(Y_q) * H_q ), in place in the scratch block of (G, q)
//   long_cols_combine (pointwise)  x[i + L0 j] = sum_q w_R^{-j q} w_P^{-i q} z_q[i]                  the radix-R DIT step
// Until round 2 the DIF step was fused into the first pass of long_cols: every one of the R CTAs of a column pair re-read the
// R/2 input blocks and spent as many instructions on addresses, selects and twiddles as on its convolution (14.4 ms for
// long_cols at 16384^2 against 5.9 ms of convolutions at the rate of rs_cols; profiles/long_probe_r02r.txt).  The split pass
// costs one more write + read of the column tile (contiguous, 16-byte coalesced) and leaves a plain convolution kernel.
struct XlLongColsSplit {
    static const char* name() { return "long_cols_split"; }
    typedef XlLongParams Params;
    static constexpr int NT = 256;
    static size_t smem() { return 0; }
    XL_DEV static void run(const Params& p, cf*) {
        const int G = XL_BLOCK_Y;
        const cf* tile = p.spec + (size_t)G * p.chunk_rows * XL_V;
        const size_t cstride = (size_t)p.pairs * p.chunk_rows * XL_V;
        XL_THREADS(tid, NT) {
            const int i = XL_BLOCK_X * NT + tid;
            if (i < p.L0) {
                cf x0[4], x1[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {              // rows >= N are zero padding: j < R/2
                    const int n = i + p.L0 * j;
                    const bool ok = (2 * j < p.R) & (n < p.N);
                    xl_ld4(tile + (xl_long_col_off(p, cstride, n) & -(size_t)ok), &x0[j], &x1[j]);
                    x0[j] = xl_sel(ok, x0[j]); x1[j] = xl_sel(ok, x1[j]);
                }
                cf* Y = p.scratch + ((size_t)G * p.R * p.L0 + i) * XL_V;
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    if (q >= p.R) break;
                    cf a0 = x0[0], a1 = x1[0];
#pragma unroll
                    for (int j = 1; j < 4; ++j) {
                        const cf w = xl_tw_at(p.tw, j * q, p.R);
                        a0 = cf_fma(x0[j], w, a0); a1 = cf_fma(x1[j], w, a1);
                    }
                    const cf wiq = xl_tw_at(p.tw, i * q, p.P);
                    xl_st4(Y + (size_t)q * p.L0 * XL_V, cf_mul(a0, wiq), cf_mul(a1, wiq));
                }
            }
        }
    }
};
template <int L0> struct XlLongColsOp : XlOpBase {
    static constexpr int R1 = xl_first_radix(L0), S1 = L0 / R1;
    cf* Yq; const cf* Hq; int rev;   // this CTA's block of the scratch buffer (in place); the transfer-function block of this
                                      // sub-line, or (rev = L0 - 1) of its mirror sub-line, read in reverse slot order
    XL_DEV void load(int i, cf* v, int stride) const { xl_ld4(Yq + (size_t)i * XL_V, v, v + stride); }
    XL_DEV void spec(int beta, cf* v) const {
#pragma unroll
        for (int qq = 0; qq < 16; ++qq) {
            cf h0, h1;
            xl_ldg4(Hq + (size_t)((qq * (L0 / 16) + beta) ^ rev) * XL_V, &h0, &h1);   // L0 - 1 - slot == slot ^ (L0 - 1)
            v[qq] = cf_mul(v[qq], h0);
            v[16 + qq] = cf_mul(v[16 + qq], h1);
        }
    }
    // in place: every load() of the CTA happened before the first barrier of the transform
    XL_DEV void store_vec(int n, const cf* v) const {
#pragma unroll
        for (int j = 0; j < R1; ++j) xl_st4(Yq + (size_t)(n + S1 * j) * XL_V, v[j], v[R1 + j]);
    }
};
template <int L0> struct XlLongCols {
    static const char* name() { return "long_cols"; }
    typedef XlLongParams Params;
    static constexpr int NT = xl_threads(L0);
    static size_t smem() { return xl_smem_bytes(L0, XL_V); }
    XL_DEV static void run(const Params& p, cf* s) {
        cf* t = s + xl_tile_elems(L0, XL_V);
        const int q = XL_BLOCK_X;                                                   // (pair G = blockIdx.y, sub-line q = blockIdx.x)
        const size_t blk = ((size_t)XL_BLOCK_Y * p.R + q) * L0 * XL_V;
        // the transfer function is even in y: the slab holds the sub-lines q <= R/2 only, sub-line R - q is sub-line q in
        // reverse slot order (XlLongHRows); the two CTAs run side by side and share the block in L2
        const bool mir = 2 * q > p.R;
        const cf* Hq = p.H + ((size_t)XL_BLOCK_Y * p.R + (mir ? p.R - q : q)) * L0 * XL_V;
        XlLongColsOp<L0> op{{}, p.scratch + blk, Hq, mir ? L0 - 1 : 0};
        XL_THREADS(tid, NT) {      // the transfer-function block is needed in the spectrum phase: ask L2 for it now
            constexpr unsigned BYTES = L0 * XL_V * sizeof(cf), CH = BYTES < 32768 ? BYTES : 32768;
            if (tid == 0)
                for (unsigned o = 0; o < BYTES; o += CH) xl_prefetch_l2_bulk((const char*)Hq + o, CH);
        }
        XlFft<L0, XL_V>::conv_g(s, t, p.tw, op);
    }
};
// column tile rows i + L0 j (< N)  <-  sum_q w_R^{-j q} w_P^{-i q} z_q[i]
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
                    xl_ld4(p.scratch + (((((size_t)G * p.R + q) * p.L0 + i) * XL_V) & -(size_t)(q < p.R)), &z0[q], &z1[q]);
                    const cf w = cf_conj(xl_tw_at(p.tw, i * q, p.P));      // w_P^{-i q} (long_cols stores the bare IDFT)
                    z0[q] = xl_sel(q < p.R, cf_mul(z0[q], w)); z1[q] = xl_sel(q < p.R, cf_mul(z1[q], w));
                }
                cf* tile = p.spec + (size_t)G * p.chunk_rows * XL_V;
                const size_t cstride = (size_t)p.pairs * p.chunk_rows * XL_V;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int n = i + p.L0 * j;
                    if (2 * j >= p.R || n >= p.N) continue;
                    cf a0 = z0[0], a1 = z1[0];
#pragma unroll
                    for (int q = 1; q < 8; ++q) {   // z[q >= R] == 0
                        const cf w = cf_conj(xl_tw_at(p.tw, j * q, p.R)); a0 = cf_fma(z0[q], w, a0); a1 = cf_fma(z1[q], w, a1);
                    }
                    xl_st4(tile + xl_long_col_off(p, cstride, n), a0, a1);
                }
            }
        }
    }
};

// long_cols + long_cols_combine as ONE cluster kernel (see XlLongRowsInvC): the R CTAs of a column pair park
// w_P^{-i q} z_q[i] in their tiles, CTA c combines the positions of its slice from the R tiles over distributed shared memory
// and writes the kept rows of the column tile.  Saves the write + read of the z_q scratch (2 x 8.6 GB at 16384^2).
template <int L0> struct XlLongColsZOp : XlOpBase {
    static constexpr int R1 = xl_first_radix(L0), S1 = L0 / R1;
    const XlLongParams& p; int q; const cf* Yq; const cf* Hq; int rev; cf* s;
    XL_DEV void load(int i, cf* v, int stride) const { xl_ld4(Yq + (size_t)i * XL_V, v, v + stride); }
    XL_DEV void spec(int beta, cf* v) const {
#pragma unroll
        for (int qq = 0; qq < 16; ++qq) {
            cf h0, h1;
            xl_ldg4(Hq + (size_t)((qq * (L0 / 16) + beta) ^ rev) * XL_V, &h0, &h1);   // L0 - 1 - slot == slot ^ (L0 - 1)
            v[qq] = cf_mul(v[qq], h0);
            v[16 + qq] = cf_mul(v[16 + qq], h1);
        }
    }
    XL_DEV void store_vec(int n, const cf* v) const {
#pragma unroll
        for (int j = 0; j < R1; ++j) {
            const int i = n + S1 * j;
            const cf w = cf_conj(xl_tw_at(p.tw, i * q, p.P));     // w_P^{-i q}
            const cf o[2] = {cf_mul(v[j], w), cf_mul(v[R1 + j], w)};
            XlTile<2>::st(s, i, o, 1);
        }
    }
};
template <int L0> struct XlLongColsC {
    static const char* name() { return "long_cols"; }
    typedef XlLongParams Params;
    static constexpr int NT = xl_threads(L0);
    static size_t smem() { return xl_smem_bytes(L0, XL_V); }
    XL_DEV static void run1(const Params& p, cf* s) {
        cf* t = s + xl_tile_elems(L0, XL_V);
        const int q = XL_BLOCK_X;                                                   // (pair G = blockIdx.y, sub-line q = blockIdx.x)
        const size_t blk = ((size_t)XL_BLOCK_Y * p.R + q) * L0 * XL_V;
        const bool mir = 2 * q > p.R;                                               // even in y: see XlLongCols
        const cf* Hq = p.H + ((size_t)XL_BLOCK_Y * p.R + (mir ? p.R - q : q)) * L0 * XL_V;
        XlLongColsZOp<L0> op{{}, p, q, p.scratch + blk, Hq, mir ? L0 - 1 : 0, s};
        XL_THREADS(tid, NT) {
            constexpr unsigned BYTES = L0 * XL_V * sizeof(cf), CH = BYTES < 32768 ? BYTES : 32768;
            if (tid == 0)
                for (unsigned o = 0; o < BYTES; o += CH) xl_prefetch_l2_bulk((const char*)Hq + o, CH);
        }
        XlFft<L0, XL_V>::conv_g(s, t, p.tw, op);
    }
    XL_DEV static void run2(const Params& p, cf*, const XlPeers& pr) {
        const int c = XL_BLOCK_X, G = XL_BLOCK_Y, slice = L0 / p.R;
        cf* tile = p.spec + (size_t)G * p.chunk_rows * XL_V;
        const size_t cstride = (size_t)p.pairs * p.chunk_rows * XL_V;
        XL_THREADS(tid, NT) {
            for (int ii = tid; ii < slice; ii += NT) {
                const int i = c * slice + ii;
                cf z0[8], z1[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    z0[q] = cf_zero(); z1[q] = cf_zero();
                    if (q < p.R) xl_peer_ld4(pr, q, 2 * xl_pad(i), &z0[q], &z1[q]);
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int n = i + L0 * j;
                    if (2 * j >= p.R || n >= p.N) continue;
                    cf a0 = z0[0], a1 = z1[0];
#pragma unroll
                    for (int q = 1; q < 8; ++q) {   // z[q >= R] == 0
                        const cf w = cf_conj(xl_tw_at(p.tw, j * q, p.R)); a0 = cf_fma(z0[q], w, a0); a1 = cf_fma(z1[q], w, a1);
                    }
                    xl_st4(tile + xl_long_col_off(p, cstride, n), a0, a1);
                }
            }
        }
    }
};

// ------------------------------------------------------------------------------------------------ transfer function
// h_eval: the impulse response of this rank's y rows AND the radix-R DIF step of their row transforms in one pointwise pass:
//     Y[y][q][i] = w_P^{i q} sum_{j < R} x_y[i + L0 j] w_R^{j q},   x_y[n] = h(min(n, P - n) dx, y dy)   (h is even in x),
// for the sub-lines q <= R/2 (the others are mirrors, XlLongHRows).  Positions i = t and i = L0 - t need the SAME R samples --
// |x| = t + L0 j (j < R/2) and |x| = L0 j - t (1 <= j <= R/2) -- so one thread serves both and every sample is evaluated
// exactly once.  (Until round 2: a plane of samples written by h_eval and re-read R times by the R sub-line CTAs of
// long_h_rows, 256 strided loads per thread; 1.3 + 3.7 ms at 16384^2.)   scratch: Y[hrows][R/2 + 1][L0]
struct XlHEval {
    static const char* name() { return "h_eval"; }
    typedef XlLongParams Params;
    static constexpr int NT = 256;
    static size_t smem() { return 0; }
    XL_DEV static cf pick(const cf* a, int k) {      // a[k], k in [0, 5), without a run-time register index
        cf r = a[0];
#pragma unroll
        for (int m = 1; m < 5; ++m) r = k == m ? a[m] : r;
        return r;
    }
    XL_DEV static void emit(const Params& p, cf* Yrow, int i, const cf* x) {   // x[j] = x_y[i + L0 j], j < R (0 beyond)
#pragma unroll
        for (int q = 0; q <= 4; ++q) {
            if (2 * q > p.R) break;
            cf acc = x[0];
#pragma unroll
            for (int j = 1; j < 8; ++j) acc = cf_fma(x[j], xl_tw_at(p.tw, (j < p.R ? j : 0) * q, p.R), acc);
            Yrow[(size_t)q * p.L0 + i] = cf_mul(acc, xl_tw_at(p.tw, i * q, p.P));
        }
    }
    XL_DEV static void run(const Params& p, cf*) {
        const XlRsHConst hc = xl_rs_hconst(xl_ldg(p.z), p.k);
        const int half = p.R / 2, yl = XL_BLOCK_Y, y = p.hrow0 + yl, deriv = (p.flags & XL_F_DERIV) ? 1 : 0;
        const bool row = y <= p.P / 2;            // the rows beyond P/2 only pad the row count to an even number
        cf* Yrow = p.scratch + (size_t)yl * (half + 1) * p.L0;
        XL_THREADS(tid, NT) {
            const int t = XL_BLOCK_X * NT + tid;
            if (t <= p.L0 / 2) {
                cf A[5], B[5];
#pragma unroll
                for (int j = 0; j <= 4; ++j) {
                    A[j] = cf_zero(); B[j] = cf_zero();
                    if (row && (j < half || (j == half && t == 0))) A[j] = xl_rs_h((t + p.L0 * j) * p.dx, y * p.dy, hc, deriv);
                    if (row && j >= 1 && j <= half && t > 0) B[j] = xl_rs_h((p.L0 * j - t) * p.dx, y * p.dy, hc, deriv);
                }
                if (t == 0) {                     // |x| = L0 j - 0 is the sample A[j]
#pragma unroll
                    for (int j = 1; j <= 4; ++j) B[j] = A[j];
                }
                cf x[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {     // position t:  n = t + L0 j,  mirrored beyond P/2
                    x[j] = cf_zero();
                    if (j < half) x[j] = A[j < 5 ? j : 0];
                    else if (j == half) x[j] = B[j < 5 ? j : 0];       // t == 0: B == A
                    else if (j < p.R) x[j] = pick(B, p.R - j);
                }
                emit(p, Yrow, t, x);
                if (t > 0 && 2 * t < p.L0) {      // position L0 - t:  n = L0 (j + 1) - t
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        x[j] = cf_zero();
                        if (j + 1 <= half) x[j] = B[(j + 1) < 5 ? (j + 1) : 0];
                        else if (j < p.R) x[j] = pick(A, p.R - j - 1);
                    }
                    emit(p, Yrow, p.L0 - t, x);
                }
            }
        }
    }
};
// row spectra of the impulse response from the blocks h_eval wrote
template <int L0> struct XlLongHRowsOp : XlOpBase {
    const XlLongParams& p; int q, yb; int mq;   // mq: the sub-line that mirrors q (== q: none), see XlLongHRows
    XL_DEV void load(int i, cf* v, int stride) const {
        const size_t rs = (size_t)(p.R / 2 + 1) * L0;          // elements per row of Y (hrows is even: both rows exist)
        const cf* y0 = p.scratch + ((size_t)yb * (p.R / 2 + 1) + q) * L0 + i;
#pragma unroll
        for (int l = 0; l < XL_V; ++l) v[l * stride] = y0[l * rs];
    }
    XL_DEV void spec(int beta, const cf* v) const {
#pragma unroll
        for (int qq = 0; qq < 16; ++qq) {
            const int slot = qq * (L0 / 16) + beta, g = q * L0 + slot;
            xl_blocked_store2<xl_lane_mask(L0)>(p.spec + (size_t)(g / 2) * p.hrows * 2, yb, p.hrows, g, v[qq], v[16 + qq]);
            if (mq != q) {   // CTA-uniform
                const int gm = mq * L0 + (L0 - 1 - slot);   // the two lanes of a pair still hold the two slots of one pair
                xl_blocked_store2<xl_lane_mask(L0)>(p.spec + (size_t)(gm / 2) * p.hrows * 2, yb, p.hrows, gm, v[qq], v[16 + qq]);
            }
        }
    }
    XL_DEV void store_vec(int, const cf*) const {}
};
// The impulse response is even in x (and in y): x[n] == x[P - n], so its spectrum is even too, X[b] == X[P - b].  Sub-line q
// holds the bins q + R k and sub-line R - q the bins (R - q) + R k' == P - (q + R k) for k' = L0 - 1 - k; the slot map is a
// digit permutation, so slot(L0 - 1 - k) == L0 - 1 - slot(k): sub-line R - q is sub-line q in reverse slot order.  The CTAs
// of the sub-lines 0 < q < R/2 therefore store their spectrum twice and those of q > R/2 exit at once (3/8 of the transforms
// of both transfer-function kernels at R = 8).
XL_DEV int xl_long_mirror(int q, int R) { return (q > 0 && 2 * q < R) ? R - q : q; }
template <int L0> struct XlLongHRows {
    static const char* name() { return "long_h_rows"; }
    typedef XlLongParams Params;
    static constexpr int NT = xl_threads(L0);
    static size_t smem() { return xl_smem_bytes(L0, XL_V); }
    XL_DEV static void run(const Params& p, cf* s) {
        cf* t = s + xl_tile_elems(L0, XL_V);
        if (2 * XL_BLOCK_X > p.R) return;   // CTA-uniform: written by the CTA of the mirror sub-line
        XlLongHRowsOp<L0> op{{}, p, XL_BLOCK_X, XL_BLOCK_Y * XL_V, xl_long_mirror(XL_BLOCK_X, p.R)};
        XlFft<L0, XL_V>::forward_g(s, t, p.tw, op);
    }
};
// column spectra of the impulse-response row spectra (even in y: row P - y == row y), exchanged layout in, H slab out.
// Two launches, as for the field: long_h_split writes the radix-R DIF step of every needed sub-line (q <= R/2, the others are
// mirrors) into the block of the transfer-function slab where that sub-line's spectrum will live, long_h_cols transforms the
// blocks in place.
struct XlLongHSplit {
    static const char* name() { return "long_h_split"; }
    typedef XlLongParams Params;
    static constexpr int NT = 256;
    static size_t smem() { return 0; }
    XL_DEV static void run(const Params& p, cf*) {
        const int G = XL_BLOCK_Y;
        const cf* src = p.spec + (size_t)G * p.chunk_rows * XL_V;
        const size_t cstride = (size_t)p.pairs * p.chunk_rows * XL_V;
        XL_THREADS(tid, NT) {
            const int i = XL_BLOCK_X * NT + tid;
            if (i < p.L0) {
                cf x0[8], x1[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int n = i + p.L0 * j;
                    const bool ok = j < p.R;
                    const int r = n <= p.P / 2 ? n : p.P - n;
                    xl_ld4(src + (xl_long_col_off(p, cstride, r) & -(size_t)ok), &x0[j], &x1[j]);
                    x0[j] = xl_sel(ok, x0[j]); x1[j] = xl_sel(ok, x1[j]);
                }
                cf* Y = p.H + ((size_t)G * p.P + i) * XL_V;
#pragma unroll
                for (int q = 0; q <= 4; ++q) {
                    if (2 * q > p.R) break;
                    cf a0 = x0[0], a1 = x1[0];
#pragma unroll
                    for (int j = 1; j < 8; ++j) {          // x[j >= R] == 0
                        const cf w = xl_tw_at(p.tw, j * q, p.R);
                        a0 = cf_fma(x0[j], w, a0); a1 = cf_fma(x1[j], w, a1);
                    }
                    const cf wiq = xl_tw_at(p.tw, i * q, p.P);
                    xl_st4(Y + (size_t)q * p.L0 * XL_V, cf_mul(a0, wiq), cf_mul(a1, wiq));
                }
            }
        }
    }
};
template <int L0> struct XlLongHColsOp : XlOpBase {
    cf* Hq; float hscale;     // this sub-line's block (in place); the mirror sub-lines are never stored (XlLongCols reads q reversed)
    XL_DEV void load(int i, cf* v, int stride) const { xl_ld4(Hq + (size_t)i * XL_V, v, v + stride); }
    XL_DEV void spec(int beta, const cf* v) const {
#pragma unroll
        for (int qq = 0; qq < 16; ++qq) {
            xl_st4(Hq + (size_t)(qq * (L0 / 16) + beta) * XL_V, cf_scale(v[qq], hscale), cf_scale(v[16 + qq], hscale));
        }
    }
    XL_DEV void store_vec(int, const cf*) const {}
};
template <int L0> struct XlLongHCols {
    static const char* name() { return "long_h_cols"; }
    typedef XlLongParams Params;
    static constexpr int NT = xl_threads(L0);
    static size_t smem() { return xl_smem_bytes(L0, XL_V); }
    XL_DEV static void run(const Params& p, cf* s) {
        const int q = XL_BLOCK_X;  // <= R/2 (the grid of xl_slab_h_cols)
        cf* t = s + xl_tile_elems(L0, XL_V);
        XlLongHColsOp<L0> op{{}, p.H + ((size_t)XL_BLOCK_Y * p.R + q) * L0 * XL_V, p.hscale};
        // the in-place update is safe: every load of the first pass happens before the barrier that precedes the stores
        XlFft<L0, XL_V>::forward_g(s, t, p.tw, op);
    }
};
