// xl_fft.cuh -- shared-memory radix-decomposed complex64 FFT engine for one CTA.
//
// A CTA owns a tile of V "lines" (rows or columns of the 2-D problem), each of power-of-two length L, stored in shared
// memory interleaved as tile[pad(i)][c] (c = line, innermost).  Thread work is always indexed (c fastest, then butterfly),
// so that a warp touches V adjacent lines x 32/V adjacent positions: 32-byte (V=4 columns) or 64-byte (V=4 rows)
// global segments and conflict-free shared-memory phases (pad(i) = i + i/16).
//
// Forward = in-place decimation-in-frequency:  natural order in  -> digit-permuted spectrum out.
// Inverse = in-place decimation-in-time:       digit-permuted spectrum in -> natural order out.
// Radix plan: L = r * 16^a, passes (r, 16, ..., 16) with r in {2,4,8,16} FIRST, so every pass but the last has
// stride >= 16 and the last pass works on 16 contiguous elements.  After the last forward pass the thread that owns
// butterfly `beta` holds, in v[q], the DFT bin of "slot" q*(L/16)+beta; the same thread starts the inverse from the
// same registers, so  forward-last / spectrum multiply / inverse-first  are fused in registers (XlConv).
// The slot order is an arbitrary but fixed permutation of the frequency bins: transfer functions are produced by the
// same forward code, hence in the same order, and no reordering pass ever exists.
//
// The first forward pass reads its operands through op.load(c,i) and the last inverse pass emits through
// op.store(c,i,v): zero padding, cropping, analytic factors and layout changes live in those functors and never touch HBM.
#pragma once
#include "xl_platform.h"

#define XL_TWN 16384  // twiddle table: tw[k] = exp(-2*pi*i*k/XL_TWN), generated in fp64

constexpr int xl_first_radix(int L) { return L > 16 ? xl_first_radix(L / 16) : L; }
constexpr int xl_tile_elems(int L, int V) { return (L + L / 16) * V; }
constexpr int xl_threads(int L, int V) { return (L * V / 32) < 32 ? 32 : ((L * V / 32) > 512 ? 512 : (L * V / 32)); }

template <int V> XL_DEV int xl_tidx(int i, int c) { return (i + (i >> 4)) * V + c; }

// multiply by (wr + i*DIR*wi)
template <int DIR> XL_DEV cf xl_mulw(cf a, float wr, float wi) {
    return DIR < 0 ? make_float2(a.x * wr + a.y * wi, a.y * wr - a.x * wi)
                   : make_float2(a.x * wr - a.y * wi, a.y * wr + a.x * wi);
}
// multiply by DIR*i
template <int DIR> XL_DEV cf xl_muli(cf a) { return DIR < 0 ? make_float2(a.y, -a.x) : make_float2(-a.y, a.x); }

template <int DIR> XL_DEV void xl_fft4(cf& a0, cf& a1, cf& a2, cf& a3) {
    cf t0 = cf_add(a0, a2), t1 = cf_sub(a0, a2), t2 = cf_add(a1, a3), t3 = xl_muli<DIR>(cf_sub(a1, a3));
    a0 = cf_add(t0, t2); a2 = cf_sub(t0, t2); a1 = cf_add(t1, t3); a3 = cf_sub(t1, t3);
}

// In-register DFT of R points, sign DIR (-1: exp(-2 pi i jq/R)); result in natural order.
template <int R, int DIR> struct XlBfly;
template <int DIR> struct XlBfly<2, DIR> {
    XL_DEV static void run(cf* v) { cf a = v[0]; v[0] = cf_add(a, v[1]); v[1] = cf_sub(a, v[1]); }
};
template <int DIR> struct XlBfly<4, DIR> {
    XL_DEV static void run(cf* v) { xl_fft4<DIR>(v[0], v[1], v[2], v[3]); }
};
template <int DIR> struct XlBfly<8, DIR> {
    XL_DEV static void run(cf* v) {
        const float r = 0.70710678118654752f;
#pragma unroll
        for (int j = 0; j < 4; ++j) { cf a = v[j]; v[j] = cf_add(a, v[4 + j]); v[4 + j] = cf_sub(a, v[4 + j]); }
        v[5] = xl_mulw<DIR>(v[5], r, r);
        v[6] = xl_muli<DIR>(v[6]);
        v[7] = xl_mulw<DIR>(v[7], -r, r);
        xl_fft4<DIR>(v[0], v[1], v[2], v[3]);
        xl_fft4<DIR>(v[4], v[5], v[6], v[7]);
        cf t[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) t[q] = v[4 * (q % 2) + q / 2];
#pragma unroll
        for (int q = 0; q < 8; ++q) v[q] = t[q];
    }
};
template <int DIR> struct XlBfly<16, DIR> {
    XL_DEV static void run(cf* v) {
        const float c1 = 0.92387953251128674f, s1 = 0.38268343236508977f, r = 0.70710678118654752f;
#pragma unroll
        for (int j = 0; j < 4; ++j) xl_fft4<DIR>(v[j], v[4 + j], v[8 + j], v[12 + j]);
        // v[4*q1 + j2] *= w16^(j2*q1)
        v[5] = xl_mulw<DIR>(v[5], c1, s1);    // e=1
        v[6] = xl_mulw<DIR>(v[6], r, r);      // e=2
        v[7] = xl_mulw<DIR>(v[7], s1, c1);    // e=3
        v[9] = xl_mulw<DIR>(v[9], r, r);      // e=2
        v[10] = xl_muli<DIR>(v[10]);          // e=4
        v[11] = xl_mulw<DIR>(v[11], -r, r);   // e=6
        v[13] = xl_mulw<DIR>(v[13], s1, c1);  // e=3
        v[14] = xl_mulw<DIR>(v[14], -r, r);   // e=6
        v[15] = xl_mulw<DIR>(v[15], -c1, -s1);  // e=9
#pragma unroll
        for (int q1 = 0; q1 < 4; ++q1) xl_fft4<DIR>(v[4 * q1], v[4 * q1 + 1], v[4 * q1 + 2], v[4 * q1 + 3]);
        cf t[16];
#pragma unroll
        for (int q = 0; q < 16; ++q) t[q] = v[4 * (q % 4) + q / 4];
#pragma unroll
        for (int q = 0; q < 16; ++q) v[q] = t[q];
    }
};

// v[q] *= w^(q*e), w = exp(DIR*2*pi*i/XL_TWN); requires (R-1)*e < XL_TWN.  4 table loads + products of depth <= 2.
template <int R, int DIR> XL_DEV void xl_twiddle(cf* v, const cf* XL_RESTRICT tw, int e) {
#define XL_TWMUL(a, w) (DIR < 0 ? cf_mul(a, w) : cf_mulc(a, w))
    cf w1 = xl_ldg(tw + e);
    v[1] = XL_TWMUL(v[1], w1);
    if constexpr (R >= 4) {
        cf w2 = xl_ldg(tw + 2 * e);
        cf w3 = cf_mul(w1, w2);
        v[2] = XL_TWMUL(v[2], w2);
        v[3] = XL_TWMUL(v[3], w3);
        if constexpr (R >= 8) {
            cf w4 = xl_ldg(tw + 4 * e);
            cf w5 = cf_mul(w4, w1), w6 = cf_mul(w4, w2), w7 = cf_mul(w4, w3);
            v[4] = XL_TWMUL(v[4], w4);
            v[5] = XL_TWMUL(v[5], w5);
            v[6] = XL_TWMUL(v[6], w6);
            v[7] = XL_TWMUL(v[7], w7);
            if constexpr (R >= 16) {
                cf w8 = xl_ldg(tw + 8 * e);
                v[8] = XL_TWMUL(v[8], w8);
                v[9] = XL_TWMUL(v[9], cf_mul(w8, w1));
                v[10] = XL_TWMUL(v[10], cf_mul(w8, w2));
                v[11] = XL_TWMUL(v[11], cf_mul(w8, w3));
                v[12] = XL_TWMUL(v[12], cf_mul(w8, w4));
                v[13] = XL_TWMUL(v[13], cf_mul(w8, w5));
                v[14] = XL_TWMUL(v[14], cf_mul(w8, w6));
                v[15] = XL_TWMUL(v[15], cf_mul(w8, w7));
            }
        }
    }
#undef XL_TWMUL
}

template <int L, int V, int NT> struct XlFft {
    static constexpr int R1 = xl_first_radix(L);
    static constexpr int S1 = L / R1;
    static_assert(L >= 32 && (L & (L - 1)) == 0, "L must be a power of two >= 32");

    // ---- forward ----
    template <class Op> XL_DEV static void fwd_first(cf* s, const cf* XL_RESTRICT tw, const Op& op) {
        XL_THREADS(tid, NT) {
            for (int u = tid; u < S1 * V; u += NT) {
                const int c = u % V, n = u / V;
                cf v[R1];
#pragma unroll
                for (int j = 0; j < R1; ++j) v[j] = op.load(c, n + S1 * j);
                XlBfly<R1, -1>::run(v);
                xl_twiddle<R1, -1>(v, tw, n * (XL_TWN / L));
#pragma unroll
                for (int q = 0; q < R1; ++q) s[xl_tidx<V>(n + S1 * q, c)] = v[q];
            }
        }
        XL_SYNC();
    }
    template <int B> XL_DEV static void fwd_mid(cf* s, const cf* XL_RESTRICT tw) {
        constexpr int S = B / 16;
        XL_THREADS(tid, NT) {
            for (int u = tid; u < (L / 16) * V; u += NT) {
                const int c = u % V, beta = u / V, b = beta / S, n = beta % S, base = b * B + n;
                cf v[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = s[xl_tidx<V>(base + S * j, c)];
                XlBfly<16, -1>::run(v);
                xl_twiddle<16, -1>(v, tw, n * (XL_TWN / B));
#pragma unroll
                for (int q = 0; q < 16; ++q) s[xl_tidx<V>(base + S * q, c)] = v[q];
            }
        }
        XL_SYNC();
    }
    template <int B> XL_DEV static void fwd_mids(cf* s, const cf* XL_RESTRICT tw) {
        if constexpr (B >= 256) { fwd_mid<B>(s, tw); fwd_mids<B / 16>(s, tw); }
    }
    // ---- inverse ----
    template <int B> XL_DEV static void inv_mid(cf* s, const cf* XL_RESTRICT tw) {
        constexpr int S = B / 16;
        XL_THREADS(tid, NT) {
            for (int u = tid; u < (L / 16) * V; u += NT) {
                const int c = u % V, beta = u / V, b = beta / S, n = beta % S, base = b * B + n;
                cf v[16];
#pragma unroll
                for (int q = 0; q < 16; ++q) v[q] = s[xl_tidx<V>(base + S * q, c)];
                xl_twiddle<16, +1>(v, tw, n * (XL_TWN / B));
                XlBfly<16, +1>::run(v);
#pragma unroll
                for (int j = 0; j < 16; ++j) s[xl_tidx<V>(base + S * j, c)] = v[j];
            }
        }
        XL_SYNC();
    }
    template <int B> XL_DEV static void inv_mids(cf* s, const cf* XL_RESTRICT tw) {
        if constexpr (B >= 256) { inv_mids<B / 16>(s, tw); inv_mid<B>(s, tw); }
    }
    template <class Op> XL_DEV static void inv_last(cf* s, const cf* XL_RESTRICT tw, const Op& op) {
        XL_THREADS(tid, NT) {
            for (int u = tid; u < S1 * V; u += NT) {
                const int c = u % V, n = u / V;
                cf v[R1];
#pragma unroll
                for (int q = 0; q < R1; ++q) v[q] = s[xl_tidx<V>(n + S1 * q, c)];
                xl_twiddle<R1, +1>(v, tw, n * (XL_TWN / L));
                XlBfly<R1, +1>::run(v);
#pragma unroll
                for (int j = 0; j < R1; ++j) op.store(c, n + S1 * j, v[j]);
            }
        }
    }

    // ---- whole-tile drivers ----
    // FWD: op.load -> spectrum; op.spec(c, beta, v) consumes v[q] = bin at slot q*(L/16)+beta.
    template <class Op> XL_DEV static void forward(cf* s, const cf* XL_RESTRICT tw, const Op& op) {
        fwd_first(s, tw, op);
        fwd_mids<L / R1>(s, tw);
        XL_THREADS(tid, NT) {
            for (int u = tid; u < (L / 16) * V; u += NT) {
                const int c = u % V, beta = u / V;
                cf v[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = s[xl_tidx<V>(16 * beta + j, c)];
                XlBfly<16, -1>::run(v);
                op.spec(c, beta, v);
            }
        }
    }
    // CONV: op.load -> forward -> op.spec multiplies in registers -> inverse -> op.store.  (1/L is the op's business.)
    template <class Op> XL_DEV static void conv(cf* s, const cf* XL_RESTRICT tw, const Op& op) {
        fwd_first(s, tw, op);
        fwd_mids<L / R1>(s, tw);
        XL_THREADS(tid, NT) {
            for (int u = tid; u < (L / 16) * V; u += NT) {
                const int c = u % V, beta = u / V;
                cf v[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = s[xl_tidx<V>(16 * beta + j, c)];
                XlBfly<16, -1>::run(v);
                op.spec(c, beta, v);
                XlBfly<16, +1>::run(v);
#pragma unroll
                for (int j = 0; j < 16; ++j) s[xl_tidx<V>(16 * beta + j, c)] = v[j];
            }
        }
        XL_SYNC();
        inv_mids<L / R1>(s, tw);
        inv_last(s, tw, op);
    }
    // INV: op.spec fills v[q] from the stored spectrum -> inverse -> op.store.
    template <class Op> XL_DEV static void inverse(cf* s, const cf* XL_RESTRICT tw, const Op& op) {
        XL_THREADS(tid, NT) {
            for (int u = tid; u < (L / 16) * V; u += NT) {
                const int c = u % V, beta = u / V;
                cf v[16];
                op.spec(c, beta, v);
                XlBfly<16, +1>::run(v);
#pragma unroll
                for (int j = 0; j < 16; ++j) s[xl_tidx<V>(16 * beta + j, c)] = v[j];
            }
        }
        XL_SYNC();
        inv_mids<L / R1>(s, tw);
        inv_last(s, tw, op);
    }
};
