// xl_fft.cuh -- shared-memory radix-decomposed complex64 FFT engine for one CTA (engine v3).
//
// A CTA owns V "lines" (rows or columns of the 2-D problem; V = 1 or 2), each of power-of-two length L, in a shared-memory
// tile tile[pad(i)][V] (pad(i) = i + i/16; one LDS.64 / LDS.128 per position, conflict-free in every pass).  A thread
// owns the same butterfly of all V lines, so twiddle factors are loaded / derived once per thread and reused V times.
// Complex numbers are (re, im) float2 register pairs and all arithmetic is Blackwell f32x2 (FADD2 / FMUL2 / FFMA2 with
// swap / half-negate / broadcast operand modifiers, xl_platform.h): one instruction per complex add, two per complex multiply.
//
// Forward = in-place decimation-in-frequency:  natural order in  -> digit-permuted spectrum out.
// Inverse = in-place decimation-in-time:       digit-permuted spectrum in -> natural order out.
// Radix plan: L = r * 16^a, passes (r, 16, ..., 16) with r in {2,4,8,16} FIRST, so every pass but the last has stride >= 16
// and the last pass works on 16 contiguous positions.  After the last forward pass the thread that owns butterfly `beta`
// holds, in v[q], the DFT bin of "slot" q*(L/16)+beta; the same thread starts the inverse from the same registers, so
// forward-last / spectrum multiply / inverse-first  are fused in registers (conv()).
// The slot order is an arbitrary but fixed permutation of the frequency bins: transfer functions are produced by the same
// forward code, hence in the same order, and no reordering pass ever exists.
//
// Twiddles: per pass level (block B, radix R) two small tables live in shared memory, A[n] = w_B^n and (R == 16)
// C[n] = w_B^(8n), n < B/R (4.3 KB in total for L = 4096); the other powers are products of depth <= 3.
//
// The first forward pass reads through op.load(i, v, stride) (all V lines of position i) (with Op::kInLoHalf only the lower half: the upper half is zero
// padding, pruned at compile time) and the last inverse pass emits through op.store_vec() (Op::kOutLoHalf: only the lower
// half of the positions is computed): zero padding, cropping, analytic factors and layout changes live in those functors
// and never touch HBM.
#pragma once
#include "xl_platform.h"

#define XL_TWN 32768  // master twiddle table in global memory: tw[k] = exp(-2*pi*i*k/XL_TWN), generated in fp64

XL_HD constexpr int xl_first_radix(int L) { return L > 16 ? xl_first_radix(L / 16) : L; }
XL_HD constexpr int xl_tile_elems(int L, int V) { return (L + L / 16) * V; }          // cf elements of the padded tile
XL_HD constexpr int xl_threads(int L) { return (L / 16) < 32 ? 32 : (L / 16); }
// lanes of a warp that run the butterfly loops `for (beta = tid; beta < L/16; beta += NT)` of XlFft<L>
XL_HD constexpr unsigned xl_lane_mask(int L) { return (L / 16) >= 32 ? 0xffffffffu : ((1u << (L / 16)) - 1u); }
XL_HD constexpr int xl_tw_mids(int B) { return B >= 256 ? 2 * (B / 16) + xl_tw_mids(B / 16) : 0; }
XL_HD constexpr int xl_tw_level0(int L) { return (L / xl_first_radix(L)) * (xl_first_radix(L) == 16 ? 2 : 1); }
XL_HD constexpr int xl_tw_total(int L) { return xl_tw_level0(L) + xl_tw_mids(L / xl_first_radix(L)); }
// offset (in cf) of the tables of the mid level with block size B
XL_HD constexpr int xl_tw_off_mid(int L, int B) {
    int off = xl_tw_level0(L);
    for (int b = L / xl_first_radix(L); b > B; b /= 16) off += 2 * (b / 16);
    return off;
}
XL_HD constexpr size_t xl_smem_bytes(int L, int V) { return (size_t)(xl_tile_elems(L, V) + xl_tw_total(L)) * 8; }

XL_DEV int xl_pad(int i) { return i + (i >> 4); }

// slot <-> DFT bin of XlFft<L>: with radices (r_1 = R1, 16, ..., 16) and k = q_1 + r_1 q_2 + r_1 r_2 q_3 + ..., the bin
// k sits in slot  q_m (L/16) + sum_{i<m} q_i L / (16 r_1 ... r_i)   (for L = 4096: the two low hex digits swapped).
XL_HD constexpr int xl_bin_to_slot(int L, int k) {
    int P = xl_first_radix(L);
    int slot = (k % P) * (L / (16 * P));
    k /= P;
    while (P * 16 < L) { P *= 16; slot += (k % 16) * (L / (16 * P)); k /= 16; }
    return slot + k * (L / 16);
}
// The same maps with L as a template constant: every divisor is a compile-time power of two (shifts and masks in SASS; the
// run-time-L versions above cost real integer divisions when a kernel evaluates them per thread).
template <int L> XL_DEV int xl_slot_to_bin_t(int s) {
    constexpr int NB = L / 16, R1 = xl_first_radix(L), W1 = NB / R1;   // W1 = 16^(number of middle passes)
    static_assert(W1 == 1 || W1 == 16 || W1 == 256, "radix plan deeper than L = 4096 * 16");
    const int qm = s / NB, beta = s % NB, q1 = beta / W1, rem = beta % W1;
    int k = qm * NB + q1;
    if (W1 == 16) k += rem * R1;
    if (W1 == 256) k += (rem / 16) * R1 + (rem % 16) * R1 * 16;
    return k;
}
template <int L> XL_DEV int xl_bin_to_slot_t(int k) {
    constexpr int NB = L / 16, R1 = xl_first_radix(L), W1 = NB / R1;
    const int qm = k / NB, low = k % NB, q1 = low % R1, rest = low / R1;   // rest: the middle digits, least significant first
    int beta = q1 * W1;
    if (W1 == 16) beta += rest;
    if (W1 == 256) beta += (rest % 16) * 16 + rest / 16;
    return qm * NB + beta;
}
XL_HD constexpr int xl_slot_to_bin(int L, int s) {
    int P = xl_first_radix(L);
    int k = s / (L / 16) * (L / 16);          // q_m * (r_1 ... r_{m-1}) == q_m * L/16
    int beta = s % (L / 16);
    int w = L / (16 * P);
    k += beta / w;
    beta %= w;
    while (P * 16 < L) { w /= 16; k += (beta / w) * P; beta %= w; P *= 16; }
    return k;
}

// tile access: all V lines of one position in a single shared-memory transaction; line l lands in v[l * stride]
template <int V> struct XlTile;
template <> struct XlTile<1> {
    XL_DEV static void ld(const cf* s, int i, cf* v, int) { v[0] = s[xl_pad(i)]; }
    XL_DEV static void st(cf* s, int i, const cf* v, int) { s[xl_pad(i)] = v[0]; }
};
// line 0 of a two-line tile, seen as a one-line tile (K4 variant: the inverse runs in place on the cotangent's line)
struct XlTileLine0Of2 {
    XL_DEV static void ld(const cf* s, int i, cf* v, int) { v[0] = s[2 * xl_pad(i)]; }
    XL_DEV static void st(cf* s, int i, const cf* v, int) { s[2 * xl_pad(i)] = v[0]; }
};
template <> struct XlTile<2> {
    XL_DEV static void ld(const cf* s, int i, cf* v, int stride) {
        const float4 t = *reinterpret_cast<const float4*>(s + 2 * xl_pad(i));
        v[0] = make_float2(t.x, t.y);
        v[stride] = make_float2(t.z, t.w);
    }
    XL_DEV static void st(cf* s, int i, const cf* v, int stride) {
        *reinterpret_cast<float4*>(s + 2 * xl_pad(i)) = make_float4(v[0].x, v[0].y, v[stride].x, v[stride].y);
    }
};

// a + DIR*i*b
template <int DIR> XL_DEV cf xl_addi(cf a, cf b) { return DIR > 0 ? cf_add(a, cf_muli(b)) : cf_add(a, cf_mulni(b)); }
// multiply by (wr + i*DIR*wi)
template <int DIR> XL_DEV cf xl_mulw(cf a, float wr, float wi) {
    return DIR < 0 ? cf_mulc(a, make_float2(wr, wi)) : cf_mul(a, make_float2(wr, wi));
}
template <int DIR> XL_DEV cf xl_muli(cf a) { return DIR < 0 ? cf_mulni(a) : cf_muli(a); }

// 4-point DFT, sign DIR.  INLO: a2 = a3 = 0 on entry (never read).  OUTLO: only a0, a1 are produced.
template <int DIR, bool INLO, bool OUTLO> XL_DEV void xl_fft4(cf& a0, cf& a1, cf& a2, cf& a3) {
    if (INLO) {
        const cf x0 = a0, x1 = a1;
        a0 = cf_add(x0, x1);
        a1 = xl_addi<DIR>(x0, x1);
        if (!OUTLO) { a2 = cf_sub(x0, x1); a3 = xl_addi<-DIR>(x0, x1); }
    } else {
        const cf t0 = cf_add(a0, a2), t1 = cf_sub(a0, a2), t2 = cf_add(a1, a3), d = cf_sub(a1, a3);
        a0 = cf_add(t0, t2);
        a1 = xl_addi<DIR>(t1, d);
        if (!OUTLO) { a2 = cf_sub(t0, t2); a3 = xl_addi<-DIR>(t1, d); }
    }
}

// In-register DFT of R points, sign DIR (-1: exp(-2 pi i jq/R)); result in natural order.
// INLO: inputs v[R/2..R) are zero (never read).  OUTLO: only outputs v[0..R/2) are produced.
template <int R, int DIR, bool INLO, bool OUTLO> struct XlBfly;
template <int DIR, bool INLO, bool OUTLO> struct XlBfly<2, DIR, INLO, OUTLO> {
    XL_DEV static void run(cf* v) {
        if (INLO) { if (!OUTLO) v[1] = v[0]; }
        else { const cf a = v[0]; v[0] = cf_add(a, v[1]); if (!OUTLO) v[1] = cf_sub(a, v[1]); }
    }
};
template <int DIR, bool INLO, bool OUTLO> struct XlBfly<4, DIR, INLO, OUTLO> {
    XL_DEV static void run(cf* v) { xl_fft4<DIR, INLO, OUTLO>(v[0], v[1], v[2], v[3]); }
};
template <int DIR, bool INLO, bool OUTLO> struct XlBfly<8, DIR, INLO, OUTLO> {
    XL_DEV static void run(cf* v) {
        const float r = 0.70710678118654752f;
        // layer 1: radix-2 over (j, j+4); outputs q1 = 0 -> v[j], q1 = 1 -> v[4+j]
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (INLO) v[4 + j] = v[j];
            else { const cf a = v[j]; v[j] = cf_add(a, v[4 + j]); v[4 + j] = cf_sub(a, v[4 + j]); }
        }
        v[5] = xl_mulw<DIR>(v[5], r, r);
        v[6] = xl_muli<DIR>(v[6]);
        v[7] = xl_mulw<DIR>(v[7], -r, r);
        // layer 2: 4-point DFTs over j within each q1; output q = q1 + 2*q2 comes from v[4*q1 + q2]
        xl_fft4<DIR, false, OUTLO>(v[0], v[1], v[2], v[3]);
        xl_fft4<DIR, false, OUTLO>(v[4], v[5], v[6], v[7]);
        cf t[8];
#pragma unroll
        for (int q = 0; q < (OUTLO ? 4 : 8); ++q) t[q] = v[4 * (q % 2) + q / 2];
#pragma unroll
        for (int q = 0; q < (OUTLO ? 4 : 8); ++q) v[q] = t[q];
    }
};
template <int DIR, bool INLO, bool OUTLO> struct XlBfly<16, DIR, INLO, OUTLO> {
    XL_DEV static void run(cf* v) {
        const float c1 = 0.92387953251128674f, s1 = 0.38268343236508977f, r = 0.70710678118654752f;
        // layer 1: 4-point DFTs over j1 (inputs 4*j1 + j2); output q1 lands in v[4*q1 + j2]
#pragma unroll
        for (int j = 0; j < 4; ++j) xl_fft4<DIR, INLO, false>(v[j], v[4 + j], v[8 + j], v[12 + j]);
        // v[4*q1 + j2] *= w16^(j2*q1)
        v[5] = xl_mulw<DIR>(v[5], c1, s1);      // e=1
        v[6] = xl_mulw<DIR>(v[6], r, r);        // e=2
        v[7] = xl_mulw<DIR>(v[7], s1, c1);      // e=3
        v[9] = xl_mulw<DIR>(v[9], r, r);        // e=2
        v[10] = xl_muli<DIR>(v[10]);            // e=4
        v[11] = xl_mulw<DIR>(v[11], -r, r);     // e=6
        v[13] = xl_mulw<DIR>(v[13], s1, c1);    // e=3
        v[14] = xl_mulw<DIR>(v[14], -r, r);     // e=6
        v[15] = xl_mulw<DIR>(v[15], -c1, -s1);  // e=9
        // layer 2: 4-point DFTs over j2 within each q1; output q = q1 + 4*q2 comes from v[4*q1 + q2]
#pragma unroll
        for (int q1 = 0; q1 < 4; ++q1) xl_fft4<DIR, false, OUTLO>(v[4 * q1], v[4 * q1 + 1], v[4 * q1 + 2], v[4 * q1 + 3]);
        cf t[16];
#pragma unroll
        for (int q = 0; q < (OUTLO ? 8 : 16); ++q) t[q] = v[4 * (q % 4) + q / 4];
#pragma unroll
        for (int q = 0; q < (OUTLO ? 8 : 16); ++q) v[q] = t[q];
    }
};

// powers w[q] = w1^q, q < R, from w1 and (R == 16) w8 = w1^8; products of depth <= 3
template <int R> XL_DEV void xl_tw_powers(cf* w, cf w1, cf w8) {
    w[1] = w1;
    if (R >= 4) { w[2] = cf_mul(w1, w1); w[3] = cf_mul(w1, w[2]); }
    if (R >= 8) {
        w[4] = cf_mul(w[2], w[2]);
#pragma unroll
        for (int q = 1; q < 4; ++q) w[4 + q] = cf_mul(w[4], w[q]);
    }
    if (R >= 16) {
        w[8] = w8;
#pragma unroll
        for (int q = 1; q < 8; ++q) w[8 + q] = cf_mul(w8, w[q]);
    }
}
// v[q] *= w[q] (DIR<0) or conj(w[q]) (DIR>0) for q in [1, NQ)
template <int NQ, int DIR> XL_DEV void xl_twiddle(cf* v, const cf* w) {
#pragma unroll
    for (int q = 1; q < NQ; ++q) v[q] = DIR < 0 ? cf_mul(v[q], w[q]) : cf_mulc(v[q], w[q]);
}

// Default functor pieces; ops derive from this and override what they need.
struct XlOpBase {
    static constexpr bool kInLoHalf = false;   // load() is identically zero for i >= L/2 (compile-time pruning)
    static constexpr bool kOutLoHalf = false;  // store_vec() only receives positions < L/2
    static constexpr bool kSpecSyncCta = false; // conv(): after_spec_sync() needs every thread of the CTA past the spectrum phase
    // run-time (CTA-uniform) version of kInLoHalf: skips the loads and prologue math of the upper half, keeps the full
    // butterfly -- for ops whose sizes are not known at compile time and whose variants are too many to double
    XL_DEV bool in_lo_rt() const { return false; }
    // Hooks for ops that stage their operands in shared memory with asynchronous copies (xl_async.cuh); all empty here.
    // called by every thread before its first load() of a transform (wait for the staged input)
    XL_DEV void before_first() const {}
    // called by every thread after its first-pass work, before the first barrier of a transform
    XL_DEV void before_first_sync() const {}
    // called by every thread right after that barrier (the inputs of this transform are no longer needed)
    XL_DEV void after_first_sync(int) const {}
    // called by every thread before its first spec() of a transform (wait for staged spectrum factors)
    XL_DEV void before_spec() const {}
    // conv() only: called by every thread right after the barrier that ends the spectrum phase
    XL_DEV void after_spec_sync(int) const {}
};

template <int L, int V, class Tile = XlTile<V>> struct XlFft {
    static constexpr int NT = xl_threads(L);
    static constexpr int R1 = xl_first_radix(L);
    static constexpr int S1 = L / R1;
    static constexpr int TILE = xl_tile_elems(L, V);
    static_assert(L >= 32 && (L & (L - 1)) == 0, "L must be a power of two >= 32");
    static_assert(V == 1 || V == 2, "1 or 2 lines per CTA");

    // copy this L's twiddle tables (xl_tw_total(L) entries) from the global master table into shared memory `t`
    XL_DEV static void init_tw(cf* t, const cf* XL_RESTRICT gtw) {
        XL_THREADS(tid, NT) {
            for (int n = tid; n < S1; n += NT) {
                t[n] = xl_ldg(gtw + n * (XL_TWN / L));
                if (R1 == 16) t[S1 + n] = xl_ldg(gtw + 8 * n * (XL_TWN / L));
            }
            int off = xl_tw_level0(L);
            for (int B = L / R1; B >= 256; B /= 16) {
                const int nb = B / 16;
                for (int n = tid; n < nb; n += NT) {
                    t[off + n] = xl_ldg(gtw + n * (XL_TWN / B));
                    t[off + nb + n] = xl_ldg(gtw + 8 * n * (XL_TWN / B));
                }
                off += 2 * nb;
            }
        }
        XL_SYNC();
    }

    // The same tables, with the global loads (tw_fetch) and the shared-memory stores (tw_store) split so that a kernel can
    // put its own first-pass global loads between them: the CTA start then pays ONE memory latency instead of two (ncu round
    // 2: ~10 % of the samples of every one-item-per-CTA kernel sat on the stores of init_tw waiting for the twiddle loads).
    // Each thread fetches exactly the entries it stores; entries [0, kTw0) are the level-0 twiddles of ITS butterflies
    // n = tid + i NT of the first pass (w1, and w8 when R1 == 16), which fwd_first_g takes from the registers.
    static constexpr int kTw0Iter = (S1 + NT - 1) / NT;
    static constexpr int kTw0 = kTw0Iter * (R1 == 16 ? 2 : 1);
    static constexpr int kTwRegs = kTw0 + 2 * ((L / R1) >= 4096 ? 2 : ((L / R1) >= 256 ? 1 : 0));
    XL_DEV static void tw_fetch(int tid, const cf* XL_RESTRICT gtw, cf* r) {
        int k = 0;
#pragma unroll
        for (int i = 0; i < kTw0Iter; ++i) {
            const int n = tid + i * NT, m = n < S1 ? n : 0;
            r[k++] = xl_ldg(gtw + m * (XL_TWN / L));
            if (R1 == 16) r[k++] = xl_ldg(gtw + 8 * m * (XL_TWN / L));
        }
#pragma unroll
        for (int B = L / R1; B >= 256; B /= 16) {
            const int m = tid < B / 16 ? tid : 0;
            r[k++] = xl_ldg(gtw + m * (XL_TWN / B));
            r[k++] = xl_ldg(gtw + 8 * m * (XL_TWN / B));
        }
    }
    XL_DEV static void tw_store(int tid, cf* t, const cf* r) {
        int k = 0;
#pragma unroll
        for (int i = 0; i < kTw0Iter; ++i) {
            const int n = tid + i * NT;
            if (n < S1) t[n] = r[k];
            ++k;
            if (R1 == 16) { if (n < S1) t[S1 + n] = r[k]; ++k; }
        }
        int off = xl_tw_level0(L);
#pragma unroll
        for (int B = L / R1; B >= 256; B /= 16) {
            const int nb = B / 16;
            if (tid < nb) { t[off + tid] = r[k]; t[off + nb + tid] = r[k + 1]; }
            k += 2;
            off += 2 * nb;
        }
    }

    // Barrier between two passes.  The passes inside a block of 256 positions (the B = 256 level and the contiguous-16 pass,
    // forward and inverse) are private to one half-warp: the thread that owns butterfly beta of the contiguous pass reads
    // positions 16 beta + j, which the B = 256 level wrote from threads 16 (beta / 16) + j -- and the other way round on the
    // way back.  A warp-level barrier is enough there (XL_NO_SYNCWARP: A/B switch back to CTA barriers), so the warps of a
    // CTA drift apart and their shared-memory and butterfly phases overlap instead of alternating in lock-step.
    template <bool HALF_WARP_LOCAL> XL_DEV static void sync_local() {
#ifdef XL_NO_SYNCWARP
        XL_SYNC();
#else
        if (HALF_WARP_LOCAL) XL_SYNCWARP(); else XL_SYNC();
#endif
    }
    static constexpr bool kHasMid = (L / R1) >= 256;   // a B = 256 level exists: spectrum <-> next pass is half-warp local

    // ---- forward ----
    template <class Op> XL_DEV static void fwd_first(cf* s, const cf* t, const Op& op) {
        XL_THREADS(tid, NT) {
            op.before_first();
            for (int n = tid; n < S1; n += NT) {
                cf v[V * R1];
#pragma unroll
                for (int j = 0; j < (Op::kInLoHalf ? R1 / 2 : R1); ++j) {   // line l -> v[l*R1 + j]
                    if (!Op::kInLoHalf && j >= R1 / 2 && op.in_lo_rt()) {
#pragma unroll
                        for (int l = 0; l < V; ++l) v[l * R1 + j] = cf_zero();
                    } else {
                        op.load(n + S1 * j, v + j, R1);
                    }
                }
                cf w[R1];
                xl_tw_powers<R1>(w, t[n], R1 == 16 ? t[S1 + n] : cf_zero());
#pragma unroll
                for (int l = 0; l < V; ++l) {
                    XlBfly<R1, -1, Op::kInLoHalf, false>::run(v + l * R1);
                    xl_twiddle<R1, -1>(v + l * R1, w);
                }
#pragma unroll
                for (int q = 0; q < R1; ++q) Tile::st(s, n + S1 * q, v + q, R1);
            }
            op.before_first_sync();   // empty unless the op has asynchronous copies in flight
        }
        XL_SYNC();
        XL_THREADS(tid, NT) { op.after_first_sync(tid); }   // every load() of this transform has completed in every thread
    }
    // first pass of a CTA that has NOT run init_tw: fetches the tables itself, behind the op's own global loads
    template <class Op> XL_DEV static void fwd_first_g(cf* s, cf* t, const cf* XL_RESTRICT gtw, const Op& op) {
        XL_THREADS(tid, NT) {
            op.before_first();
            cf tr[kTwRegs];
            tw_fetch(tid, gtw, tr);
            int it = 0;
            for (int n = tid; n < S1; n += NT, ++it) {
                cf v[V * R1];
#pragma unroll
                for (int j = 0; j < (Op::kInLoHalf ? R1 / 2 : R1); ++j) {   // line l -> v[l*R1 + j]
                    if (!Op::kInLoHalf && j >= R1 / 2 && op.in_lo_rt()) {
#pragma unroll
                        for (int l = 0; l < V; ++l) v[l * R1 + j] = cf_zero();
                    } else {
                        op.load(n + S1 * j, v + j, R1);
                    }
                }
                cf w[R1];
                cf w1 = tr[0], w8 = cf_zero();
#pragma unroll
                for (int i = 0; i < kTw0Iter; ++i)      // compile-time register index
                    if (i == it) { w1 = tr[i * (R1 == 16 ? 2 : 1)]; if (R1 == 16) w8 = tr[i * 2 + 1]; }
                xl_tw_powers<R1>(w, w1, w8);
#pragma unroll
                for (int l = 0; l < V; ++l) {
                    XlBfly<R1, -1, Op::kInLoHalf, false>::run(v + l * R1);
                    xl_twiddle<R1, -1>(v + l * R1, w);
                }
#pragma unroll
                for (int q = 0; q < R1; ++q) Tile::st(s, n + S1 * q, v + q, R1);
            }
            tw_store(tid, t, tr);
            op.before_first_sync();
        }
        XL_SYNC();
        XL_THREADS(tid, NT) { op.after_first_sync(tid); }
    }
    template <int B> XL_DEV static void fwd_mid(cf* s, const cf* tw) {
        constexpr int S = B / 16;
        const cf* t = tw + xl_tw_off_mid(L, B);
        XL_THREADS(tid, NT) {
            for (int beta = tid; beta < L / 16; beta += NT) {
                const int b = beta / S, n = beta % S, base = b * B + n;
                cf v[V * 16];
#pragma unroll
                for (int j = 0; j < 16; ++j) Tile::ld(s, base + S * j, v + j, 16);
                cf w[16];
                xl_tw_powers<16>(w, t[n], t[S + n]);
#pragma unroll
                for (int l = 0; l < V; ++l) {
                    XlBfly<16, -1, false, false>::run(v + l * 16);
                    xl_twiddle<16, -1>(v + l * 16, w);
                }
#pragma unroll
                for (int q = 0; q < 16; ++q) Tile::st(s, base + S * q, v + q, 16);
            }
        }
        sync_local<B == 256>();   // the pass after the B = 256 level is the contiguous-16 pass
    }
    template <int B> XL_DEV static void fwd_mids(cf* s, const cf* tw) {
        if constexpr (B >= 256) { fwd_mid<B>(s, tw); fwd_mids<B / 16>(s, tw); }
    }
    // ---- inverse ----
    template <int B> XL_DEV static void inv_mid(cf* s, const cf* tw) {
        constexpr int S = B / 16;
        const cf* t = tw + xl_tw_off_mid(L, B);
        XL_THREADS(tid, NT) {
            for (int beta = tid; beta < L / 16; beta += NT) {
                const int b = beta / S, n = beta % S, base = b * B + n;
                cf v[V * 16];
#pragma unroll
                for (int q = 0; q < 16; ++q) Tile::ld(s, base + S * q, v + q, 16);
                cf w[16];
                xl_tw_powers<16>(w, t[n], t[S + n]);
#pragma unroll
                for (int l = 0; l < V; ++l) {
                    xl_twiddle<16, +1>(v + l * 16, w);
                    XlBfly<16, +1, false, false>::run(v + l * 16);
                }
#pragma unroll
                for (int j = 0; j < 16; ++j) Tile::st(s, base + S * j, v + j, 16);
            }
        }
        XL_SYNC();
    }
    template <int B> XL_DEV static void inv_mids(cf* s, const cf* tw) {
        if constexpr (B >= 256) { inv_mids<B / 16>(s, tw); inv_mid<B>(s, tw); }
    }
    template <class Op> XL_DEV static void inv_last(cf* s, const cf* t, const Op& op) {
        XL_THREADS(tid, NT) {
            for (int n = tid; n < S1; n += NT) {
                cf v[V * R1];
#pragma unroll
                for (int q = 0; q < R1; ++q) Tile::ld(s, n + S1 * q, v + q, R1);
                cf w[R1];
                xl_tw_powers<R1>(w, t[n], R1 == 16 ? t[S1 + n] : cf_zero());
#pragma unroll
                for (int l = 0; l < V; ++l) {
                    xl_twiddle<R1, +1>(v + l * R1, w);
                    XlBfly<R1, +1, false, Op::kOutLoHalf>::run(v + l * R1);
                }
                op.store_vec(n, v);   // v[l*R1 + j] is position n + S1*j of line l (j < R1/2 only with kOutLoHalf)
            }
        }
    }

    // ---- whole-tile drivers (s = tile of xl_tile_elems(L,V) cf, t = twiddle tables filled by init_tw) ----
    // The *_g drivers are for CTAs that have not run init_tw: they take the global master table and fill `t` on the way.
    // FWD: op.load -> spectrum; op.spec(beta, v) consumes v[l*16 + q] = bin at slot q*(L/16)+beta of line l.
    template <class Op> XL_DEV static void forward_rest(cf* s, const cf* t, const Op& op) {
        fwd_mids<L / R1>(s, t);
        XL_THREADS(tid, NT) {
            op.before_spec();
            for (int beta = tid; beta < L / 16; beta += NT) {
                cf v[V * 16];
#pragma unroll
                for (int j = 0; j < 16; ++j) Tile::ld(s, 16 * beta + j, v + j, 16);
#pragma unroll
                for (int l = 0; l < V; ++l) XlBfly<16, -1, false, false>::run(v + l * 16);
                op.spec(beta, v);
            }
        }
    }
    template <class Op> XL_DEV static void forward(cf* s, const cf* t, const Op& op) {
        fwd_first(s, t, op);
        forward_rest(s, t, op);
    }
    template <class Op> XL_DEV static void forward_g(cf* s, cf* t, const cf* gtw, const Op& op) {
        fwd_first_g(s, t, gtw, op);
        forward_rest(s, t, op);
    }
    // CONV: op.load -> forward -> op.spec multiplies in registers -> inverse -> op.store_vec.  (1/L is the op's business.)
    template <class Op> XL_DEV static void conv_rest(cf* s, const cf* t, const Op& op) {
        fwd_mids<L / R1>(s, t);
        XL_THREADS(tid, NT) {
            op.before_spec();
            for (int beta = tid; beta < L / 16; beta += NT) {
                cf v[V * 16];
#pragma unroll
                for (int j = 0; j < 16; ++j) Tile::ld(s, 16 * beta + j, v + j, 16);
#pragma unroll
                for (int l = 0; l < V; ++l) XlBfly<16, -1, false, false>::run(v + l * 16);
                op.spec(beta, v);
#pragma unroll
                for (int l = 0; l < V; ++l) XlBfly<16, +1, false, false>::run(v + l * 16);
#pragma unroll
                for (int j = 0; j < 16; ++j) Tile::st(s, 16 * beta + j, v + j, 16);
            }
        }
        // ops that hand their staging buffer to the asynchronous proxy here (after_spec_sync) need the CTA barrier
        sync_local<kHasMid && !Op::kSpecSyncCta>();
        XL_THREADS(tid, NT) { op.after_spec_sync(tid); }
        inv_mids<L / R1>(s, t);
        inv_last(s, t, op);
    }
    template <class Op> XL_DEV static void conv(cf* s, const cf* t, const Op& op) {
        fwd_first(s, t, op);
        conv_rest(s, t, op);
    }
    template <class Op> XL_DEV static void conv_g(cf* s, cf* t, const cf* gtw, const Op& op) {
        fwd_first_g(s, t, gtw, op);
        conv_rest(s, t, op);
    }
    // the inverse passes after the first (whose output the caller has already put into the tile) -> op.store_vec
    template <class Op> XL_DEV static void inverse_tail(cf* s, const cf* t, const Op& op) {
        inv_mids<L / R1>(s, t);
        inv_last(s, t, op);
    }
    // INV: op.spec fills v from a stored spectrum -> inverse -> op.store_vec.  G: fill the tables on the way (no init_tw).
    template <bool G, class Op> XL_DEV static void inverse_impl(cf* s, cf* t, const cf* gtw, const Op& op) {
        XL_THREADS(tid, NT) {
            cf tr[G ? kTwRegs : 1];
            if (G) tw_fetch(tid, gtw, tr);
            for (int beta = tid; beta < L / 16; beta += NT) {
                cf v[V * 16];
                op.spec(beta, v);
#pragma unroll
                for (int l = 0; l < V; ++l) XlBfly<16, +1, false, false>::run(v + l * 16);
#pragma unroll
                for (int j = 0; j < 16; ++j) Tile::st(s, 16 * beta + j, v + j, 16);
            }
            if (G) tw_store(tid, t, tr);
        }
        // with G the CTA barrier also publishes the tables (when there is no B = 256 level the barrier is a CTA barrier anyway)
        sync_local<kHasMid && !G>();
        inv_mids<L / R1>(s, t);
        inv_last(s, t, op);
    }
    template <class Op> XL_DEV static void inverse(cf* s, const cf* t, const Op& op) { inverse_impl<false>(s, const_cast<cf*>(t), (const cf*)0, op); }
    template <class Op> XL_DEV static void inverse_g(cf* s, cf* t, const cf* gtw, const Op& op) { inverse_impl<true>(s, t, gtw, op); }
};
