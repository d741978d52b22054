// xl_kernels.cuh -- kernel bodies of the propagation hot path (RS/VRS convolution, Bluestein CZT/VCZT, high-NA lens).
//
// Every kernel is "one CTA = one set of V = 2 lines through XlFft"; what differs is the functor (op) that feeds the first
// pass, multiplies the spectrum in registers and drains the last pass.  Reference lines replaced are cited per op.
//
// Data layouts (c64 everywhere, geometry in fp64):
//   field            [f][y][x]                     row-major N x N planes (the reference's [..., y, x])
//   row spectra  S   [f][L/2][N][2]                "blocked": 2 adjacent x-slots innermost, so a column-pair CTA reads one
//                                                   contiguous 16*N-byte tile and a row-pair CTA fills whole 32-byte sectors
//   transfer fn  H   [L/2][L][2]                   same blocking, y in slot order, premultiplied by dx*dy/L^2; h is even in x
//                                                   and y, so only the slot pairs holding an x-bin <= L/2 and, inside them, the
//                                                   slots of y-bins <= L/2 (= slots 0..L/2) are ever written or read
//   L = padded length (power of two >= 2N-1); slot order = XlFft's digit permutation (never undone).
#pragma once
#include "xl_fft.cuh"

#define XL_V 2  // lines per CTA == column-group width of the blocked layouts

// flags shared by several kernels
#define XL_F_CONJ_IN 1    // conjugate operand on load   (torch's conjugate-cotangent convention, fused)
#define XL_F_CONJ_OUT 2   // conjugate result on store
#define XL_F_VRS 4        // 3 fields, field 2 = Ez formed from (Ex,Ey) at load      (vectorized_optics.py:258-261)
#define XL_F_DERIV 8      // transfer function of dh/dz instead of h
#define XL_F_NOFIELD 16   // backward column kernel: d/dz only, the field cotangent is not needed (constant input)

XL_DEV void xl_ld4(const cf* p, cf* a, cf* b) {   // two adjacent complex values with one 16-byte load (p 16-byte aligned)
    const float4 t = *reinterpret_cast<const float4*>(p);
    *a = make_float2(t.x, t.y);
    *b = make_float2(t.z, t.w);
}
XL_DEV void xl_ldg4(const cf* p, cf* a, cf* b) {
    const float4 t = xl_ldg(reinterpret_cast<const float4*>(p));
    *a = make_float2(t.x, t.y);
    *b = make_float2(t.z, t.w);
}
XL_DEV void xl_st4(cf* p, cf a, cf b) { *reinterpret_cast<float4*>(p) = make_float4(a.x, a.y, b.x, b.y); }

// red[0] = sum of red[0..NT) by a shared-memory tree (NT a power of two); ends with a barrier.
template <int NT, class T> XL_DEV void xl_block_sum(T* red) {
    for (int st = NT / 2; st >= 1; st >>= 1) {
        XL_THREADS(tid, NT) {
            if (tid < st) red[tid] += red[tid + st];
        }
        XL_SYNC();
    }
}

// Blocked pair layout  grp[y][2]  (grp = base of slot pair g/2).  The calling thread owns slot g (parity of g == parity of
// its lane) of the two adjacent rows y0 (v0) and y0+1 (v1); rows >= nrows are not stored / read as zero.
// GPU: lanes 2k and 2k+1 swap one value so that each issues ONE 16-byte access (even lane: row y0, slots g,g+1; odd lane:
// row y0+1, slots g-1,g) -- 8-byte scattered accesses halve the L1<->L2 request efficiency (profiles/ubench_r01.txt).
template <unsigned MASK> XL_DEV void xl_blocked_store2(cf* grp, int y0, int nrows, int g, cf v0, cf v1) {
#ifdef XL_HOST_EMU
    if (y0 < nrows) grp[(size_t)y0 * 2 + (g & 1)] = v0;
    if (y0 + 1 < nrows) grp[(size_t)(y0 + 1) * 2 + (g & 1)] = v1;
#else
    const bool odd = g & 1;
    const cf got = xl_xchg1<MASK>(odd ? v0 : v1), keep = odd ? v1 : v0;
    const int y = y0 + (odd ? 1 : 0);
    if (y < nrows) xl_st4(grp + (size_t)y * 2, odd ? got : keep, odd ? keep : got);
#endif
}
template <unsigned MASK> XL_DEV void xl_blocked_load2(const cf* grp, int y0, int nrows, int g, cf* v0, cf* v1) {
#ifdef XL_HOST_EMU
    *v0 = y0 < nrows ? grp[(size_t)y0 * 2 + (g & 1)] : cf_zero();
    *v1 = y0 + 1 < nrows ? grp[(size_t)(y0 + 1) * 2 + (g & 1)] : cf_zero();
#else
    const bool odd = g & 1;
    const int y = y0 + (odd ? 1 : 0);
    cf lo = cf_zero(), hi = cf_zero();
    if (y < nrows) xl_ld4(grp + (size_t)y * 2, &lo, &hi);
    const cf got = xl_xchg1<MASK>(odd ? lo : hi);
    *v0 = odd ? got : lo;
    *v1 = odd ? hi : got;
#endif
}

// Two planes with the blocked pair layout: the calling thread owns slot g of row y of plane A (va) and of plane B (vb).
// GPU: lanes 2k and 2k+1 swap one value; the even lane stores slots (g, g+1) of A, the odd lane slots (g-1, g) of B.
template <unsigned MASK> XL_DEV void xl_blocked_store_ab(cf* grpA, cf* grpB, int y, bool ok, int g, cf va, cf vb) {
#ifdef XL_HOST_EMU
    if (ok) { grpA[(size_t)y * 2 + (g & 1)] = va; grpB[(size_t)y * 2 + (g & 1)] = vb; }
#else
    const bool odd = g & 1;
    const cf got = xl_xchg1<MASK>(odd ? va : vb);
    if (ok) xl_st4((odd ? grpB : grpA) + (size_t)y * 2, odd ? got : va, odd ? vb : got);
#endif
}

// ------------------------------------------------------------------------------------------------------------------
// Rayleigh-Sommerfeld impulse response, reference wave_optics.py:291-297:
//   h = (1/2pi) * z/r^2 * (1/r - i k) * exp(sgn(z) i k r),  r = sqrt(X^2+Y^2+z^2)
// The phase k*r reaches 4e5 rad, so r is formed in fp64 (fp32 rsqrt seed + one fp64 Newton step: rel. error ~1e-13),
// reduced to a fraction of a cycle in fp64, and only that fraction goes through fp32 sincospi.
// deriv=1: the REDUCED z-derivative  h_z - i k h,  with  dh/dz = (e/2pi) [ g + (z^2/r)(g' + s i k g) ],  g = 1/r^3 - i k/r^2,
// g' = -3/r^4 + 2 i k/r^3:   h_z - i k h = (e/2pi) [ g + (z^2/r) g' + i k g z (|z|/r - 1) ],  |z|/r - 1 = -rho^2/(r (r + |z|)).
// dh/dz is dominated by i k h (a pure phase rotation), whose contribution to the gradient of any intensity-type loss
// cancels identically; in complex64 the cancellation residue of that term buries the answer.  The library therefore
// evaluates it exactly in real space (gz += -k Im sum ct*out, on the way through rs_rows_dual) and pushes only the reduced kernel, 1e2-1e4 times
// smaller, through the FFT pipeline.
// ------------------------------------------------------------------------------------------------------------------
// 1/sqrt(r2) to ~1e-13 relative, branch-free (fp32 rsqrt seed + one fp64 Newton step); NaN for r2 == 0
XL_DEV double xl_rsqrt64(double r2) {
#ifdef XL_HOST_EMU
    const float y0 = 1.0f / sqrtf((float)r2);
#else
    const float y0 = rsqrtf((float)r2);
#endif
    const double y = (double)y0;
    const double e = fma(-(r2 * y), 0.5 * y, 0.5);   // 0.5 - 0.5*r2*y^2
    return fma(y, e, y);
}
// 1/d for d > 0 to ~1e-15 relative, branch-free (fp32 reciprocal seed + two fp64 Newton steps; the library division carries
// a slow path)
XL_DEV double xl_rcp64(double d) {
    double y = (double)xl_rcpf((float)d);
    y = fma(y, fma(-d, y, 1.0), y);
    return fma(y, fma(-d, y, 1.0), y);
}
struct XlRsHConst { double z, z2, k, kcyc, sg; };   // per-z constants shared by all samples
XL_DEV XlRsHConst xl_rs_hconst(double z, double k) {
    XlRsHConst c;
    c.z = z;
    c.z2 = z * z;
    c.k = k;
    c.kcyc = k * 0.15915494309189535;       // cycles per unit length (1/lambda)
    c.sg = z > 0 ? 1.0 : -1.0;
    return c;
}
// WHICH: bit 0 = h, bit 1 = the reduced derivative; both share r, the phase factor and the powers of 1/r
template <int WHICH> XL_DEV void xl_rs_h_eval(double X, double Y, const XlRsHConst& c, cf* h, cf* hz) {
    const double inv2pi = 0.15915494309189535;
    const double r2 = X * X + Y * Y + c.z2;
    const double y = xl_rsqrt64(r2);                  // 1/r
    const double r = r2 * y;
    const double cyc = r * c.kcyc;
    const double fr = cyc - rint(cyc);
    float sn, cs;
    xl_sincospif(2.0f * (float)fr, &sn, &cs);
    sn *= (float)c.sg;                                // exp(sgn(z) i k r)
    // amplitude (complex, multiplies the phase factor) in fp64, rounded once: gradients of intensity-type losses with
    // respect to z cancel the leading term of dh/dz, which amplifies amplitude rounding by ~k*r
    const double ir2 = y * y, ir3 = ir2 * y;
    if (WHICH & 1) {
        const float far = (float)(c.z * inv2pi * ir3), fai = (float)(-c.z * inv2pi * c.k * ir2);
        *h = make_float2(far * cs - fai * sn, far * sn + fai * cs);
    }
    if (WHICH & 2) {
        const double gr = ir3, gi = -c.k * ir2;
        const double gpr = -3.0 * ir2 * ir2, gpi = 2.0 * c.k * ir3;
        const double f = c.z2 * y;                                   // z^2/r
        const double az = c.z < 0 ? -c.z : c.z;
        const double w = -c.k * c.z * (X * X + Y * Y) * y * xl_rcp64(r + az);   // k z (|z|/r - 1)
        // g + (z^2/r) g' + i w g
        const float far = (float)((gr + f * gpr - w * gi) * inv2pi), fai = (float)((gi + f * gpi + w * gr) * inv2pi);
        *hz = make_float2(far * cs - fai * sn, far * sn + fai * cs);
    }
}
XL_DEV cf xl_rs_h(double X, double Y, const XlRsHConst& c, int deriv) {
    cf h = cf_zero(), hz = cf_zero();
    if (deriv) xl_rs_h_eval<2>(X, Y, c, &h, &hz); else xl_rs_h_eval<1>(X, Y, c, &h, &hz);
    return deriv ? hz : h;
}

// ==================================================================================================================
// RS / VRS  (wave_optics.py:281-289, vectorized_optics.py:364-373)
// ==================================================================================================================
struct XlRsParams {
    int N, L, nfields, flags;
    int f0;            // first field of this launch (blockIdx.y counts from it)
    int rows;          // rows of this launch's slab == row count of the blocked layouts (N on a single GPU)
    int chunk_rows;    // slab column kernels: rows per source rank in the exchanged layout [rank][pair][chunk_rows][2]
    int hrow0, hstore_all;   // slab h_rows: first y row of this rank; store every x slot pair (no x-mirror skipping)
    // several transfer functions in one launch pair (xl_rs_transfer_multi): item = blockIdx.y, buffer H + item * h_stride,
    // distance z[item / h_per_z]; h_per_z == 2: even items hold H, odd items the reduced dH/dz of the same distance
    long long h_stride; int h_per_z;
    int z_item_stride;   // xl_rs_transfer_multi from a batched call: item i (pair i) takes z[i * z_item_stride]; 0 means 1
    const cf* in;      // [nfields][N][N]   (XL_F_VRS: the Ex plane; Ey starts ey_off elements later)
    long long ey_off;  // XL_F_VRS: elements from the Ex plane to the Ey plane of the primal input (N*N when they are stacked)
    const cf* in2;     // rs_rows_dual: the primal field(s) whose conjugate is the second line (same shape rules as `in`)
    cf* out;           // [nfields][N][N]
    cf* spec;          // [nfields][L/2][N][2]
    cf* spec2;         // second spectra set (grad-z: spectra of conj(U))
    cf* H;             // [L/2][L][2]
    const cf* H2;      // dH/dz transfer function (grad-z)
    double* gz;        // scalar accumulator (grad-z)
    const cf* tw;
    const double* z;   // device scalar
    double x0, y0, dx, dy, k;
    float hscale;      // dx*dy/L^2
    // Pointwise elements fused into the first / last pass (xl_rs_fwd_fused / xl_rs_bwd_fused):
    const cf* mod;         // shared complex plane [N][N] multiplied into every field while it is loaded: a phase-only SLM
                           // exp(i phi) (optical_elements.py:87-103) or the beam under a batch of masks; null: none
    int in_real;           // the input planes are float32 (binary object masks), not complex64
    const float* target;   // detection: target intensities [nfields][N][N] (four_f_optical_table.py:129-141)
    double* mse;           // [nfields] += sum (|out|^2 - target)^2 / N^2, accumulated by the last pass
    const cf* seed_out;    // backward: primal output; the cotangent of the fused detection is formed from it and `target`
    const double* ct_mse;  // [nfields] dL/dmse
    cf* ct_mod;            // backward: cotangent of `mod` (summed over the fields)
    const cf* dz_out;      // rs_rows_dual: primal output [nfields][N][N]; gz += -k Im sum ct*out (the i k h part of dh/dz, exactly:
                           // fp32 x fp32 products are exact in fp64) while the cotangent is loaded anyway; null: not wanted
};

// input of the fused forward pass: plane f (complex64 or float32) x the shared complex plane
XL_DEV cf xl_fused_in(const XlRsParams& p, const void* base, size_t fo, size_t o) {
    cf v = p.in_real ? make_float2(((const float*)base)[fo + o], 0.f) : ((const cf*)base)[fo + o];
    if (p.mod) v = cf_mul(v, xl_ldg(p.mod + o));
    return v;
}
// Cotangent seeded by the fused detection, torch convention: L = sum_f ct_mse[f] mse[f], mse = sum (|out|^2 - T)^2 / N^2
//   =>  dL/d out = w (|out|^2 - T) out,  w = 4 ct_mse[f] / N^2   (real factor x out: the i k out part of d/dz vanishes exactly)
XL_DEV cf xl_seed_ct(const XlRsParams& p, float w, size_t fo, size_t o) {
    const cf a = p.seed_out[fo + o];
    const float d = a.x * a.x + a.y * a.y - p.target[fo + o];
    return cf_scale(a, w * d);
}
XL_DEV float xl_seed_weight(const XlRsParams& p, int f) {
    return p.seed_out ? (float)(4.0 * xl_ldg(p.ct_mse + f) / ((double)p.N * (double)p.N)) : 0.f;
}

// K1: rows of the zero-padded field -> blocked row spectra.   replaces the row half of fft2(U), wave_optics.py:286-288
// EZ: this CTA's field is Ez = (Ex X + Ey Y)/r formed from (Ex,Ey) while loading (vectorized_optics.py:258-261).
// The load path is branch-free so that all of a thread's global loads are in flight together.
template <int L, bool EZ, bool FUSE = false> struct XlRsRowsFwdOp : XlOpBase {
    static constexpr bool kInLoHalf = true;   // N <= L/2: the upper half of every padded row is zero
    const XlRsParams& p; int f, yb; double z2; float w;   // w: seed weight of the fused detection (FUSE backward)
    XL_DEV cf load1(int y, int i) const {
        const int N = p.N;
        const bool ok = y < p.rows && i < N;
        const size_t NN = (size_t)p.rows * N, o = ok ? (size_t)y * N + i : 0;
        cf v;
        if (EZ) {
            const cf ex = p.in[o], ey = p.in[p.ey_off + (long long)o];
            const double X = p.x0 + i * p.dx, Y = p.y0 + y * p.dy;
            const double ir = xl_rsqrt64(X * X + Y * Y + z2);
            v = cf_lin2(ex, (float)(X * ir), ey, (float)(Y * ir));
        } else if (FUSE) {
            v = p.seed_out ? xl_seed_ct(p, w, (size_t)f * NN, o) : xl_fused_in(p, p.in, (size_t)f * NN, o);
        } else {
            v = p.in[((p.flags & XL_F_VRS) ? (long long)f * p.ey_off : (long long)((size_t)f * NN)) + (long long)o];   // VRS: f is 0 or 1 here
        }
        if (p.flags & XL_F_CONJ_IN) v = cf_conj(v);
        return ok ? v : cf_zero();
    }
    XL_DEV void load(int i, cf* v, int stride) const {
#pragma unroll
        for (int l = 0; l < XL_V; ++l) v[l * stride] = load1(yb + l, i);
    }
    XL_DEV void spec(int beta, const cf* v) const {
        cf* base = p.spec + (size_t)f * L * p.rows;
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            const int g = q * (L / 16) + beta;
            xl_blocked_store2<xl_lane_mask(L)>(base + (size_t)(g / 2) * p.rows * 2, yb, p.rows, g, v[q], v[16 + q]);
        }
    }
    XL_DEV void store_vec(int, const cf*) const {}
};
template <int L, bool FUSE = false> struct XlRsRowsFwd {
    static const char* name() { return FUSE ? "rs_rows_fwd_f" : "rs_rows_fwd"; }
    typedef XlRsParams Params;
    static constexpr int NT = xl_threads(L);
    static size_t smem() { return xl_smem_bytes(L, XL_V); }
    XL_DEV static void run(const Params& p, cf* s) {
        cf* t = s + xl_tile_elems(L, XL_V);
        const int f = p.f0 + XL_BLOCK_Y, yb = XL_BLOCK_X * XL_V;
        if (FUSE) {
            XlRsRowsFwdOp<L, false, true> op{{}, p, f, yb, 0.0, xl_seed_weight(p, f)};
            XlFft<L, XL_V>::forward_g(s, t, p.tw, op);
        } else if ((p.flags & XL_F_VRS) && f == 2) {   // CTA-uniform
            const double z = xl_ldg(p.z);
            XlRsRowsFwdOp<L, true> op{{}, p, f, yb, z * z, 0.f};
            XlFft<L, XL_V>::forward_g(s, t, p.tw, op);
        } else {
            XlRsRowsFwdOp<L, false> op{{}, p, f, yb, 0.0, 0.f};
            XlFft<L, XL_V>::forward_g(s, t, p.tw, op);
        }
    }
};

// K1d: one row of the cotangent AND the same row of conj(primal field) as the two lines of one transform -> interleaved
// row spectra  CW[f][L/2 slot pairs][y][4] = (C_g, W_g, C_g+1, W_g+1):  the d/dz column kernel needs both column spectra of
// one x frequency in the registers of one thread (XlRsColsGzAsync) and reads them with one 16-byte load per row; two
// adjacent lanes (slots g, g+1) fill one 32-byte sector here.  Replaces the two separate row-spectra launches of the
// backward pass (JAX autodiff of wave_optics.py:286-288).
template <int L, bool EZ, bool FUSE = false> struct XlRsRowsDualOp : XlOpBase {
    static constexpr bool kInLoHalf = true;
    const XlRsParams& p; int f, y; double z2; float sw;   // sw: seed weight of the fused detection
    mutable double dz;   // this thread's partial sum of Im(ct*out) (dz_out); the host emulation shares one op between the threads
    XL_DEV void load(int i, cf* v, int stride) const {
        const int N = p.N;
        const bool ok = i < N;
        const size_t NN = (size_t)p.rows * N, o = ok ? (size_t)y * N + i : 0;
        cf c = (FUSE && p.seed_out) ? xl_seed_ct(p, sw, (size_t)f * NN, o) : p.in[(size_t)f * NN + o], w;
        if (p.flags & XL_F_CONJ_IN) c = cf_conj(c);
        if (p.dz_out) {   // kernel-uniform
            const cf u = p.dz_out[(size_t)f * NN + o];
            const double t = (double)c.x * (double)u.y + (double)c.y * (double)u.x;   // Im(ct*out)
            dz += ok ? t : 0.0;
        }
        if (EZ) {
            const cf ex = p.in2[o], ey = p.in2[p.ey_off + (long long)o];
            const double X = p.x0 + i * p.dx, Y = p.y0 + y * p.dy;
            const double ir = xl_rsqrt64(X * X + Y * Y + z2);
            w = cf_lin2(ex, (float)(X * ir), ey, (float)(Y * ir));
        } else if (FUSE) {
            w = xl_fused_in(p, p.in2, (size_t)f * NN, o);
        } else {
            w = p.in2[((p.flags & XL_F_VRS) ? (long long)f * p.ey_off : (long long)((size_t)f * NN)) + (long long)o];
        }
        v[0] = ok ? c : cf_zero();
        v[stride] = ok ? cf_conj(w) : cf_zero();
    }
    XL_DEV void spec(int beta, const cf* v) const {
        cf* base = p.spec + (size_t)f * L * p.rows * 2 + (size_t)y * 4;
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            const int g = q * (L / 16) + beta;
            xl_st4(base + (size_t)(g >> 1) * p.rows * 4 + (g & 1) * 2, v[q], v[16 + q]);
        }
    }
    XL_DEV void store_vec(int, const cf*) const {}
};
template <int L, bool FUSE = false> struct XlRsRowsDual {
    static const char* name() { return FUSE ? "rs_rows_dual_f" : "rs_rows_dual"; }
    typedef XlRsParams Params;
    static constexpr int NT = xl_threads(L);
    static size_t smem() { return xl_smem_bytes(L, XL_V) + NT * sizeof(double); }
    // gz += -k * (sum over the CTA of the per-thread partial sums `acc`)
    XL_DEV static void dz_flush(const Params& p, double* red, double acc) {
        XL_THREADS(tid, NT) { red[tid] = XL_PER_THREAD(tid, acc); }
        XL_SYNC();
        xl_block_sum<NT>(red);
        XL_THREADS(tid, NT) {
            if (tid == 0) xl_atomic_add(p.gz, -p.k * red[0]);
        }
    }
    XL_DEV static void run(const Params& p, cf* s) {
        cf* t = s + xl_tile_elems(L, XL_V);
        double* red = (double*)(t + xl_tw_total(L));
        const int f = p.f0 + XL_BLOCK_Y, y = XL_BLOCK_X;
        if (FUSE) {
            XlRsRowsDualOp<L, false, true> op{{}, p, f, y, 0.0, xl_seed_weight(p, f), 0.0};
            XlFft<L, XL_V>::forward_g(s, t, p.tw, op);
            if (p.dz_out) dz_flush(p, red, op.dz);     // kernel-uniform
        } else if ((p.flags & XL_F_VRS) && f == 2) {   // CTA-uniform
            const double z = xl_ldg(p.z);
            XlRsRowsDualOp<L, true> op{{}, p, f, y, z * z, 0.f, 0.0};
            XlFft<L, XL_V>::forward_g(s, t, p.tw, op);
            if (p.dz_out) dz_flush(p, red, op.dz);
        } else {
            XlRsRowsDualOp<L, false> op{{}, p, f, y, 0.0, 0.f, 0.0};
            XlFft<L, XL_V>::forward_g(s, t, p.tw, op);
            if (p.dz_out) dz_flush(p, red, op.dz);
        }
    }
};

// The impulse response is even in x, so its transfer function is even in the x frequency: only the columns of bins
// kx <= L/2 are ever computed (h_cols) and the column of bin kx > L/2 is read from its mirror L - kx.
// Returns the address of element [slot_y = 0] of the column that serves x-slot g; consecutive slot_y are 2 cf apart.
template <int L> XL_DEV const cf* xl_h_column(const cf* H, int g) {
    const int k = xl_slot_to_bin_t<L>(g);
    const int sc = xl_bin_to_slot_t<L>(k <= L / 2 ? k : L - k);
    return H + (size_t)(sc / 2) * L * 2 + (sc & 1);
}
template <int L> XL_DEV bool xl_h_pair_needed(int G) {   // does slot pair G hold a column with bin <= L/2 ?
    return xl_slot_to_bin_t<L>(2 * G) <= L / 2 || xl_slot_to_bin_t<L>(2 * G + 1) <= L / 2;
}

// Column copies.  Pair G holds a needed column (x-bin <= L/2) exactly when G <= L/4 (the top digit of a bin is the top digit
// of its slot), so the blocks G > L/4 of the [L/2][L][2] buffer are never touched by the pair layout.  That space holds a
// second copy of every needed column as a CONTIGUOUS array  Hc[slot <= L/2 + 1][xl_hc_stride]  (rows = y slots 0..L/2), the
// shape the bulk-asynchronous column kernels stage with one copy per column (xl_async.cuh).
XL_HD constexpr int xl_hc_stride(int L) { return L / 2 + 4; }                              // cf per column copy (16-byte multiple)
XL_HD constexpr int xl_hc_rows(int L) { return L / 2 + 2; }                                // cf staged per column (16-byte multiple)
XL_HD constexpr size_t xl_hc_offset(int L) { return (size_t)(L / 4 + 1) * L * 2; }         // cf offset of the copies inside H
template <int L> XL_DEV const cf* xl_h_colcopy(const cf* H, int g) {                       // the copy that serves x-slot g
    const int k = xl_slot_to_bin_t<L>(g);
    const int sc = xl_bin_to_slot_t<L>(k <= L / 2 ? k : L - k);
    return H + xl_hc_offset(L) + (size_t)sc * xl_hc_stride(L);
}

// Row (slot_y) of the stored half of the transfer function that serves slot q*(L/16)+beta: the slot itself when its
// bin is <= L/2, else the slot of the mirrored bin L - k.  With k = klow(beta) + q L/16 (q = top digit of the bin):
// L - k = (15 - q) L/16 + (L/16 - klow)  for klow != 0, and (16 - q) L/16 for klow == 0.
template <int L> struct XlHRow {
    int beta, bm;
    XL_DEV explicit XlHRow(int b) : beta(b), bm(0) {
        const int klow = xl_slot_to_bin_t<L>(b);
        if (klow) bm = xl_bin_to_slot_t<L>(L / 16 - klow);
    }
    XL_DEV int row(int q) const {
        if (q < 8) return q * (L / 16) + beta;
        if (beta == 0) return (16 - q) * (L / 16);
        return (15 - q) * (L / 16) + bm;
    }
};

// K2: column FFT of the row spectra, x transfer function, inverse column FFT, keep rows [0,N).   wave_optics.py:288
// SLAB: the tile of a column pair is not contiguous but arrives from the all-to-all as [source rank][pair][chunk_rows][2]
// (row i lives in chunk i / chunk_rows); chunk_stride = pairs_per_rank * chunk_rows * 2 elements.
template <int L, bool SLAB = false> struct XlRsColsOp : XlOpBase {
    static constexpr bool kInLoHalf = true, kOutLoHalf = true;
    static constexpr int R1 = xl_first_radix(L), S1 = L / R1;
    const XlRsParams& p; cf* tile; const cf* H0; const cf* H1;   // transfer-function columns of the two lines (stride 2)
    int hmode;   // 0: H0,H1 are the two halves of one 16-byte pair (H1 == H0+1); 1: swapped pair (H0 == H1+1); 2: unrelated
    size_t chunk_stride;
    XL_DEV cf* row(int i) const {
        if (!SLAB) return tile + (size_t)i * XL_V;
        return tile + (size_t)(i / p.chunk_rows) * chunk_stride + (size_t)(i % p.chunk_rows) * XL_V;
    }
    XL_DEV void load(int i, cf* v, int stride) const {
        if (i < p.N) xl_ld4(row(i), v, v + stride);
        else { v[0] = cf_zero(); v[stride] = cf_zero(); }
    }
    XL_DEV void spec(int beta, cf* v) const {
        const XlHRow<L> hr(beta);
        if (hmode == 2) {
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                const size_t o = (size_t)hr.row(q) * XL_V;
                v[q] = cf_mul(v[q], xl_ldg(H0 + o));
                v[16 + q] = cf_mul(v[16 + q], xl_ldg(H1 + o));
            }
        } else {
            const cf* Hp = hmode == 0 ? H0 : H1;
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                cf lo, hi;
                xl_ldg4(Hp + (size_t)hr.row(q) * XL_V, &lo, &hi);
                v[q] = cf_mul(v[q], hmode == 0 ? lo : hi);
                v[16 + q] = cf_mul(v[16 + q], hmode == 0 ? hi : lo);
            }
        }
    }
    XL_DEV void store_vec(int n, const cf* v) const {
#pragma unroll
        for (int j = 0; j < R1 / 2; ++j) {
            const int i = n + S1 * j;
            if (i < p.N) xl_st4(row(i), v[j], v[R1 + j]);
        }
    }
};
// K2 of the slab-decomposed path: this rank owns gridDim.x slot pairs; its transfer-function slab H is [pairs][L][2]
// (generated for exactly these columns, so no x-mirroring), its spectra arrive as [source rank][pairs][chunk_rows][2].
template <int L> struct XlRsColsSlab {
    static const char* name() { return "rs_cols_slab"; }
    typedef XlRsParams Params;
    static constexpr int NT = xl_threads(L);
    static size_t smem() { return xl_smem_bytes(L, XL_V); }
    XL_DEV static void run(const Params& p, cf* s) {
        cf* t = s + xl_tile_elems(L, XL_V);
        XlFft<L, XL_V>::init_tw(t, p.tw);
        const int G = XL_BLOCK_X;
        const cf* H0 = p.H + (size_t)G * L * XL_V;
        XlRsColsOp<L, true> op{{}, p, p.spec + (size_t)G * p.chunk_rows * XL_V, H0, H0 + 1, 0,
                               (size_t)p.nfields * p.chunk_rows * XL_V};   // nfields carries pairs-per-rank here
        XlFft<L, XL_V>::conv(s, t, op);
    }
};

// K3: inverse row FFT of the filtered spectra, crop columns [0,N).   wave_optics.py:288 (row half of ifft2 + crop)
// DET: fused detection -- while the result is stored, sum (|out|^2 - target)^2 is accumulated per field
// (four_f_optical_table.py:129-141, MSE of intensities): the intensity plane is never materialised.
template <int L, bool DET = false> struct XlRsRowsInvOp : XlOpBase {
    static constexpr bool kOutLoHalf = true;
    static constexpr int R1 = xl_first_radix(L), S1 = L / R1;
    const XlRsParams& p; int f, yb; float* red;   // red[S1]: per-butterfly partial sums (DET)
    XL_DEV void load(int, cf*, int) const {}
    XL_DEV void spec(int beta, cf* v) const {
        const cf* base = p.spec + (size_t)f * L * p.rows;
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            const int g = q * (L / 16) + beta;
            xl_blocked_load2<xl_lane_mask(L)>(base + (size_t)(g / 2) * p.rows * 2, yb, p.rows, g, v + q, v + 16 + q);
        }
    }
    XL_DEV void store_vec(int n, const cf* v) const {
        float acc = 0.f;
#pragma unroll
        for (int l = 0; l < XL_V; ++l) {
            const int y = yb + l;
            if (y >= p.rows) continue;
#pragma unroll
            for (int j = 0; j < R1 / 2; ++j) {
                const int i = n + S1 * j;
                if (i >= p.N) continue;
                cf val = v[l * R1 + j];
                const size_t o = (size_t)f * p.rows * p.N + (size_t)y * p.N + i;
                if (DET) {
                    const float d = val.x * val.x + val.y * val.y - p.target[o];
                    acc += d * d;
                }
                if (p.flags & XL_F_CONJ_OUT) val = cf_conj(val);
                p.out[o] = val;
            }
        }
        if (DET) red[n] = acc;
    }
};
template <int L, bool DET = false> struct XlRsRowsInv {
    static const char* name() { return DET ? "rs_rows_inv_det" : "rs_rows_inv"; }
    typedef XlRsParams Params;
    static constexpr int NT = xl_threads(L);
    static constexpr int S1 = L / xl_first_radix(L);
    static size_t smem() { return xl_smem_bytes(L, XL_V) + (DET ? NT * sizeof(double) + S1 * sizeof(float) : 0); }
    XL_DEV static void run(const Params& p, cf* s) {
        cf* t = s + xl_tile_elems(L, XL_V);
        double* dred = (double*)(t + xl_tw_total(L));
        float* red = (float*)(dred + NT);
        const int f = p.f0 + XL_BLOCK_Y;
        XlRsRowsInvOp<L, DET> op{{}, p, f, XL_BLOCK_X * XL_V, red};
        XlFft<L, XL_V>::inverse_g(s, t, p.tw, op);
        if (DET) {
            XL_SYNC();
            XL_THREADS(tid, NT) {
                double a = 0.0;
                for (int n = tid; n < S1; n += NT) a += (double)red[n];
                dred[tid] = a;
            }
            XL_SYNC();
            xl_block_sum<NT>(dred);
            XL_THREADS(tid, NT) {
                if (tid == 0) xl_atomic_add(p.mse + f, dred[0] / ((double)p.N * (double)p.N));
            }
        }
    }
};

// K3m: the inverse row kernel of the fused BACKWARD pass.  v = in * mod was propagated, so with c = A^T ct_out (what the
// inverse row FFT produces):  ct_in[f] = c[f] * mod  and  ct_mod = sum_f c[f] * in[f]  (JAX convention; XL_F_CONJ_OUT
// conjugates both on store).  One CTA owns a row pair and walks the fields, accumulating ct_mod in shared memory: no atomics
// and no per-field cotangent of the modulated field in HBM.  Replaces the VJP of the SLM multiply (optical_elements.py:87-103).
template <int L> struct XlRsRowsInvModOp : XlOpBase {
    static constexpr bool kOutLoHalf = true;
    static constexpr int R1 = xl_first_radix(L), S1 = L / R1;
    const XlRsParams& p; int f, yb; cf* acc;   // acc[XL_V][N]
    XL_DEV void load(int, cf*, int) const {}
    XL_DEV void spec(int beta, cf* v) const {
        const cf* base = p.spec + (size_t)f * L * p.rows;
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            const int g = q * (L / 16) + beta;
            xl_blocked_load2<xl_lane_mask(L)>(base + (size_t)(g / 2) * p.rows * 2, yb, p.rows, g, v + q, v + 16 + q);
        }
    }
    XL_DEV void store_vec(int n, const cf* v) const {
        const size_t NN = (size_t)p.rows * p.N;
#pragma unroll
        for (int l = 0; l < XL_V; ++l) {
            const int y = yb + l;
            if (y >= p.rows) continue;
#pragma unroll
            for (int j = 0; j < R1 / 2; ++j) {
                const int i = n + S1 * j;
                if (i >= p.N) continue;
                const cf c = v[l * R1 + j];
                const size_t o = (size_t)y * p.N + i;
                if (p.ct_mod) {
                    const cf a = p.in_real ? make_float2(((const float*)p.in)[(size_t)f * NN + o], 0.f) : p.in[(size_t)f * NN + o];
                    acc[l * p.N + i] = cf_fma(c, a, acc[l * p.N + i]);
                }
                if (p.out) {
                    cf r = p.mod ? cf_mul(c, xl_ldg(p.mod + o)) : c;
                    if (p.flags & XL_F_CONJ_OUT) r = cf_conj(r);
                    p.out[(size_t)f * NN + o] = r;
                }
            }
        }
    }
};
template <int L> struct XlRsRowsInvMod {
    static const char* name() { return "rs_rows_inv_mod"; }
    typedef XlRsParams Params;
    static constexpr int NT = xl_threads(L);
    static size_t smem() { return xl_smem_bytes(L, XL_V) + (size_t)L * sizeof(cf); }   // + acc[XL_V][N], N <= L/2
    XL_DEV static void run(const Params& p, cf* s) {
        cf* t = s + xl_tile_elems(L, XL_V);
        cf* acc = t + xl_tw_total(L);
        const int yb = XL_BLOCK_X * XL_V;
        XL_THREADS(tid, NT) {
            for (int e = tid; e < XL_V * p.N; e += NT) acc[e] = cf_zero();
        }
        XlFft<L, XL_V>::init_tw(t, p.tw);          // ends with a barrier
        for (int f = p.f0; f < p.f0 + p.nfields; ++f) {
            XlRsRowsInvModOp<L> op{{}, p, f, yb, acc};
            XlFft<L, XL_V>::inverse(s, t, op);
            XL_SYNC();                             // the last pass has read the tile; acc[] of this field is complete
        }
        if (p.ct_mod) {
            XL_THREADS(tid, NT) {
                for (int e = tid; e < XL_V * p.N; e += NT) {
                    const int y = yb + e / p.N, i = e % p.N;
                    if (y >= p.rows) continue;
                    cf r = acc[e];
                    if (p.flags & XL_F_CONJ_OUT) r = cf_conj(r);
                    p.ct_mod[(size_t)y * p.N + i] = r;
                }
            }
        }
    }
};

// K1h: rows y>=0 of the wrapped, analytically generated impulse response -> row spectra stored inside H
//      (rows 0..L/2 of each group block).   replaces transfer_function_RS + row half of fft2(H), wave_optics.py:285,291-297
// h is even in x: each sample x in [0, L/2] is evaluated once into a staging buffer and read back mirrored.
template <int L> struct XlHRowsOp : XlOpBase {
    const XlRsParams& p; int yb; const cf* stage;   // stage[(L/2+1)][XL_V]
    int nvalid;   // local rows that exist (global row <= L/2)
    cf* H;        // this item's buffer
    cf* Hz;       // dual mode: the two lines are row yb of h and of its reduced z-derivative, which goes here; else null
    XL_DEV void load(int i, cf* v, int stride) const {
        const int xi = i <= L / 2 ? i : L - i;
        xl_ld4(stage + (size_t)xi * XL_V, v, v + stride);
    }
    XL_DEV void spec(int beta, const cf* v) const {
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            // single GPU: x-bins <= L/2 only (q is the top digit of the bin; of q = 8 only the pair holding bin L/2), the
            // rest is mirrored.  slab: every slot pair, rows in a [pair][rows][2] buffer of this rank's y rows.
            const int g = q * (L / 16) + beta;
            if (!p.hstore_all && (q > 8 || (q == 8 && beta >= 2))) continue;
            const size_t o = (size_t)(g / 2) * p.rows * 2;
            if (Hz) xl_blocked_store_ab<xl_lane_mask(L)>(H + o, Hz + o, yb, yb < nvalid, g, v[q], v[16 + q]);
            else xl_blocked_store2<xl_lane_mask(L)>(H + o, yb, nvalid, g, v[q], v[16 + q]);
        }
    }
    XL_DEV void store_vec(int, const cf*) const {}
};
template <int L> struct XlHRows {
    static const char* name() { return "h_rows"; }
    typedef XlRsParams Params;
    static constexpr int NT = xl_threads(L);
    static constexpr int NSTAGE = (L / 2 + 2) * XL_V;   // +1 sample of slack keeps the size even (16-byte rows)
    static size_t smem() { return xl_smem_bytes(L, XL_V) + (size_t)NSTAGE * sizeof(cf); }
    // h_per_z == 2 (H and the reduced dH/dz of the same distance, buffers 2j and 2j+1): one CTA transforms row y of BOTH as
    // its two lines, so r, exp(i k r) and the powers of 1/r of a sample are evaluated once (blockIdx.x = y, blockIdx.y = j).
    XL_DEV static void run(const Params& p, cf* s) {
        cf* t = s + xl_tile_elems(L, XL_V);
        cf* stage = t + xl_tw_total(L);
        const bool dual = p.h_per_z == 2;
        const int yb = dual ? XL_BLOCK_X : XL_BLOCK_X * XL_V;
        const int item = XL_BLOCK_Y;
        const XlRsHConst hc = xl_rs_hconst(xl_ldg(p.z + (long long)item * (p.z_item_stride ? p.z_item_stride : 1)), p.k);
        const int deriv = (p.flags & XL_F_DERIV) ? 1 : 0;
        XL_THREADS(tid, NT) {
            cf tr[XlFft<L, XL_V>::kTwRegs];
            XlFft<L, XL_V>::tw_fetch(tid, p.tw, tr);   // the twiddle loads fly while the samples are evaluated
            if (dual) {
                for (int xi = tid; xi < L / 2 + 1; xi += NT) {
                    cf h, hz;
                    xl_rs_h_eval<3>(xi * p.dx, (p.hrow0 + yb) * p.dy, hc, &h, &hz);
                    xl_st4(stage + (size_t)xi * XL_V, h, hz);
                }
            } else {
                for (int e = tid; e < (L / 2 + 1) * XL_V; e += NT) {
                    const int xi = e / XL_V, l = e % XL_V, yi = p.hrow0 + yb + l;   // global y row
                    stage[e] = yi <= L / 2 ? xl_rs_h(xi * p.dx, yi * p.dy, hc, deriv) : cf_zero();
                }
            }
            XlFft<L, XL_V>::tw_store(tid, t, tr);
        }
        XL_SYNC();                          // stage[] and the tables are visible
        int nvalid = L / 2 + 1 - p.hrow0;
        if (nvalid > p.rows) nvalid = p.rows;
        cf* H = p.H + (long long)item * (dual ? 2 : 1) * p.h_stride;
        XlHRowsOp<L> op{{}, p, yb, stage, nvalid, H, dual ? H + p.h_stride : (cf*)0};
        XlFft<L, XL_V>::forward(s, t, op);
    }
};

// K2h: column FFT of the impulse-response row spectra (even in y: row L-y == row y), result in slot order, scaled.
template <int L> struct XlHColsOp : XlOpBase {
    const XlRsParams& p; cf* Ht; cf* Hc0; cf* Hc1;   // pair block (in place) and the two column copies
    XL_DEV void load(int i, cf* v, int stride) const {
        const int r = i <= L / 2 ? i : L - i;
        xl_ld4(Ht + (size_t)r * XL_V, v, v + stride);
    }
    XL_DEV void spec(int beta, const cf* v) const {
#pragma unroll
        for (int q = 0; q <= 8; ++q) {   // y-bins <= L/2 only: H is even in the y frequency as well (XlHRow)
            if (q == 8 && beta != 0) continue;
            const int r = q * (L / 16) + beta;
            const cf a = cf_scale(v[q], p.hscale), b = cf_scale(v[16 + q], p.hscale);
            xl_st4(Ht + (size_t)r * XL_V, a, b);
            Hc0[r] = a;
            Hc1[r] = b;
        }
    }
    XL_DEV void store_vec(int, const cf*) const {}
};
// K2h of the slab path: rows arrive as [source rank][pairs][chunk_rows][2] (chunk_rows y rows per rank), the result goes
// to this rank's transfer-function slab [pairs][L][2].
template <int L> struct XlHColsSlabOp : XlOpBase {
    const XlRsParams& p; const cf* src; cf* Ht; size_t chunk_stride;
    XL_DEV void load(int i, cf* v, int stride) const {
        const int r = i <= L / 2 ? i : L - i;
        xl_ld4(src + (size_t)(r / p.chunk_rows) * chunk_stride + (size_t)(r % p.chunk_rows) * XL_V, v, v + stride);
    }
    XL_DEV void spec(int beta, const cf* v) const {
#pragma unroll
        for (int q = 0; q <= 8; ++q) {
            if (q == 8 && beta != 0) continue;
            xl_st4(Ht + (size_t)(q * (L / 16) + beta) * XL_V, cf_scale(v[q], p.hscale), cf_scale(v[16 + q], p.hscale));
        }
    }
    XL_DEV void store_vec(int, const cf*) const {}
};
template <int L> struct XlHColsSlab {
    static const char* name() { return "h_cols_slab"; }
    typedef XlRsParams Params;
    static constexpr int NT = xl_threads(L);
    static size_t smem() { return xl_smem_bytes(L, XL_V); }
    XL_DEV static void run(const Params& p, cf* s) {
        cf* t = s + xl_tile_elems(L, XL_V);
        XlFft<L, XL_V>::init_tw(t, p.tw);
        const int G = XL_BLOCK_X;
        XlHColsSlabOp<L> op{{}, p, p.spec + (size_t)G * p.chunk_rows * XL_V, p.H + (size_t)G * L * XL_V,
                            (size_t)p.nfields * p.chunk_rows * XL_V};
        XlFft<L, XL_V>::forward(s, t, op);
    }
};
template <int L> struct XlHCols {
    static const char* name() { return "h_cols"; }
    typedef XlRsParams Params;
    static constexpr int NT = xl_threads(L);
    static size_t smem() { return xl_smem_bytes(L, XL_V); }
    XL_DEV static void run(const Params& p, cf* s) {
        if (!xl_h_pair_needed<L>(XL_BLOCK_X)) return;   // CTA-uniform: mirrored columns are never read
        cf* t = s + xl_tile_elems(L, XL_V);
        cf* H = p.H + (long long)XL_BLOCK_Y * p.h_stride;      // this item's buffer (xl_rs_transfer_multi)
        cf* Hc = H + xl_hc_offset(L) + (size_t)XL_BLOCK_X * XL_V * xl_hc_stride(L);
        XlHColsOp<L> op{{}, p, H + (size_t)XL_BLOCK_X * L * XL_V, Hc, Hc + xl_hc_stride(L)};
        // the in-place update is safe: every load of the first pass happens before the barrier that precedes the stores
        XlFft<L, XL_V>::forward_g(s, t, p.tw, op);
    }
};

// K4 (the backward column kernel with d/dz, XlRsColsGzAsync) lives in xl_async.cuh; this functor drains its one-line inverse
// into the blocked pair layout the inverse row kernel reads.
template <int L> struct XlRsColsGzOutOp : XlOpBase {
    static constexpr bool kOutLoHalf = true;
    static constexpr int R1 = xl_first_radix(L), S1 = L / R1;
    const XlRsParams& p; cf* tile; int c;
    XL_DEV void store_vec(int n, const cf* v) const {
#pragma unroll
        for (int j = 0; j < R1 / 2; ++j) {
            const int i = n + S1 * j;
            if (i < p.N) tile[(size_t)i * XL_V + c] = v[j];
        }
    }
};



// ==================================================================================================================
// Bluestein chirp-z axis pass  (wave_optics.py:385-460), with the fused prologue/epilogue factors of CZT_jit (:333-357),
// VCZT (vectorized_optics.py:341-344) and the high-NA lens (optical_elements.py:528-594).
// One launch transforms, for every line (a column of the input in the forward pass, a row in the adjoint pass):
//     out[line, o] = post[o] * IFFT_L( FFT_L( pre[k] * pro(line,k) * in[line,k] ) * ft )[off + o] * epi(line,o)
// ==================================================================================================================
enum { XL_PRO_NONE = 0, XL_PRO_RSF = 1, XL_PRO_VCZT = 2, XL_PRO_HIGHNA = 3 };
enum { XL_EPI_NONE = 0, XL_EPI_RSF = 1 };

struct XlGridFactor {     // coordinates of (line, pos): swap=0 -> X from line, Y from pos; swap=1 -> X from pos, Y from line
    double x0, dx, y0, dy;
    int swap;
};

// Factor tables.  The pointwise factors of the CZT family -- the RS factors F / F0 of CZT_jit (wave_optics.py:340-341), the
// lens matrix of the high-NA objective (optical_elements.py:528-594) -- depend on the grids and on z only, so they are
// evaluated ONCE per call by czt_tables (fp64 phases, xl_rs_h / xl_lens_row) and the axis passes gather them from L2.
// A grid that is symmetric about 0 (toolbox.space(): x_j = (j - (n-1)/2) dx) needs one quadrant only: entry q of an axis
// stands for |x| = (q + 1/2) dx (n even) or q dx (n odd); 8.4 MB instead of 33.5 MB per factor at 2048^2.
struct XlFacAxis { int n, sym; double x0, dx; };
XL_HD inline int xl_fac_size(int n, int sym) { return sym ? (n + 1) / 2 : n; }
XL_DEV int xl_fac_idx(const XlFacAxis& a, int j) {
    if (!a.sym) return j;
    const int m = 2 * j - (a.n - 1);
    return (m < 0 ? -m : m) >> 1;
}
XL_DEV double xl_fac_coord(const XlFacAxis& a, int q) { return a.sym ? (q + ((a.n & 1) ? 0.0 : 0.5)) * a.dx : a.x0 + q * a.dx; }
XL_DEV float xl_fac_sign(const XlFacAxis& a, int j) { return (a.sym && 2 * j < a.n - 1) ? -1.f : 1.f; }   // sign of x_j on a mirrored axis
// T[(c * Qy + iy) * Qx + ix], or with tr = 1 the transposed planes T[(c * Qx + ix) * Qy + iy]: an axis pass whose lines run
// along y (line = x index, position = y index) reads consecutive entries of a transposed table with consecutive lanes
// (ncu round 2: the strided gathers of the natural layout cost 6 long-scoreboard stall cycles per issue in those passes)
struct XlFacTab { const cf* T; int Qx, Qy; XlFacAxis ax, ay; int tr; };
XL_DEV size_t xl_fac_entry(const XlFacTab& t, int ix, int iy, int c) {
    return t.tr ? ((size_t)c * t.Qx + ix) * t.Qy + iy : ((size_t)c * t.Qy + iy) * t.Qx + ix;
}

struct XlCztParams {
    int L, nlines, ncomp, m_in, out_off, m_out, flags;
    int c0;               // first component of this launch: component = c0 + blockIdx.y selects the Ez / lens row math,
                          // blockIdx.y alone selects the memory planes (in_comp / out_comp strides)
    const cf* in; long long in_line, in_pos, in_comp;
    cf* out;      long long out_line, out_pos, out_comp;
    const cf* pre; const cf* ft; const cf* post;
    const cf* tw;
    int pro, epi;          // host-side selectors of the compiled <PRO, EPI> variant
    int in_weight;         // d/dz chains (xl_czt_bwd_z): multiply the input by its position index (1) or its line index (2)
    XlGridFactor gpro, gepi;
    XlFacTab tpro, tepi;  // factor tables of the prologue / epilogue (czt_tables)
    const double* z;      // device scalar (RSF / VCZT factors), may be null
    double k;             // wavenumber
    double epi_cr, epi_ci; // complex constant on the output
    int epi_times_z;      // multiply the constant by z (CZT: z*dx*dy*lambda)
    double lens_R, lens_f, lens_s2;  // high-NA: radius, focal length, sin^2(theta_max)
};

// lens factor row `comp` applied to (Ex,Ey):  apod*G*(RL[comp][0] Ex + RL[comp][1] Ey + RL[comp][2] Ez), Ez=(Ex X+Ey Y)/rho
// Branch-free; X/rho, Y/rho are 0/0 = NaN at the origin exactly like the reference (optical_elements.py:538).
XL_DEV void xl_lens_row(double X, double Y, double R, double f, double s2, int comp, float* ax, float* ay) {
    const double rho2 = X * X + Y * Y;
    const double irho = xl_rsqrt64(rho2);          // NaN when rho2 == 0
    const double rho = rho2 * irho;
    float st, ct;
    xl_sincospif((float)(rho / f * 0.31830988618379067), &st, &ct);   // theta = rho/f
    const float cp = (float)(X * irho), sp = (float)(Y * irho);       // cos(phi), sin(phi) (also the Ez coefficients)
    const double uv = rho2 / (R * R);
    const float pupil = uv < 1.0 ? 1.f : 0.f;
    const float G = pupil / sqrtf(fabsf((float)(1.0 - uv * s2)));
    const float w = sqrtf(fabsf(ct)) * G;
    float r0, r1, r2;
    if (comp == 0) { r0 = ct * cp * cp + sp * sp; r1 = ct * cp * sp - sp * cp; r2 = -cp * st; }
    else if (comp == 1) { r0 = sp * ct * cp - cp * sp; r1 = ct * sp * sp + cp * cp; r2 = -sp * st; }
    else { r0 = st * cp; r1 = st * sp; r2 = ct; }
    *ax = w * (r0 + r2 * cp);
    *ay = w * (r1 + r2 * sp);
}

// ACC selects the global access shape: 0 generic 8-byte accesses; 1 the two lines are adjacent in the INPUT (in_line == 1,
// 16-byte aligned pairs): one 16-byte load serves both; 2 the same for the OUTPUT (out_line == 1): one 16-byte store.
enum { XL_ACC_GENERIC = 0, XL_ACC_PAIR_IN = 1, XL_ACC_PAIR_OUT = 2 };
template <int L, int PRO, int EPI, int ACC> struct XlCztOp : XlOpBase {
    static constexpr int R1 = xl_first_radix(L), S1 = L / R1;
    // The paired variants are also the pruned ones (the host selects them only when the sizes allow): the input fills at
    // most the lower half of the padded line (m_in <= L/2) and the kept outputs are [0, m_out) with m_out <= L/2 (the
    // forward kernel table is rotated so that its slice starts at 0 too, XlCztSetupOp).  Pruning at compile time also
    // halves the unrolled prologue / epilogue code of these kernels.
    static constexpr bool kInLoHalf = ACC != XL_ACC_GENERIC;
    static constexpr bool kOutLoHalf = ACC != XL_ACC_GENERIC;
    const XlCztParams& p; int lb, cl, comp; float z; cf cst;   // cl: launch-local plane, comp: component
    // Hoisted out of the per-element code (all CTA-uniform): a factor-table entry is  line_part + fac_idx(pos) * pos_stride,
    // whichever way the (line, position) pair maps onto (x, y) and whichever way the table is laid out.
    struct Tab { long long line_part[XL_V]; int pos_stride, pos_n, pos_mask; float sline[XL_V]; };   // pos_mask: -1 on a mirrored axis
    Tab tp, te;              // prologue / epilogue table
    float lcoord[XL_V];      // VCZT: coordinate of each line along its own axis
    double pos0, dpos;       // VCZT: coordinates along the position axis
    // index and sign of position j on the table's position axis, branch-free (a branch here would split the first pass's loads)
    XL_DEV static int pos_idx(const Tab& t, int j) {
        int m = 2 * j - (t.pos_n - 1);
        m = (m < 0 ? -m : m) >> 1;
        return (m & t.pos_mask) | (j & ~t.pos_mask);
    }
    XL_DEV static float pos_sign(const Tab& t, int j) { return ((2 * j < t.pos_n - 1) & (t.pos_mask != 0)) ? -1.f : 1.f; }
    XL_DEV bool in_lo_rt() const { return p.m_in <= L / 2; }   // zero padding fills the upper half: skip its loads and factors
    XL_DEV static Tab tab_split(const XlFacTab& t, const XlGridFactor& g, int lb, int nlines, int c) {
        Tab r;
#pragma unroll
        for (int l = 0; l < XL_V; ++l) {
            const int line = lb + l < nlines ? lb + l : 0;
            if (!g.swap) {       // x from the line, y from the position
                const int ix = xl_fac_idx(t.ax, line);
                r.line_part[l] = t.tr ? ((long long)c * t.Qx + ix) * t.Qy : (long long)c * t.Qy * t.Qx + ix;
                r.sline[l] = xl_fac_sign(t.ax, line);
            } else {             // y from the line, x from the position
                const int iy = xl_fac_idx(t.ay, line);
                r.line_part[l] = t.tr ? (long long)c * t.Qx * t.Qy + iy : ((long long)c * t.Qy + iy) * t.Qx;
                r.sline[l] = xl_fac_sign(t.ay, line);
            }
        }
        r.pos_stride = g.swap ? (t.tr ? t.Qy : 1) : (t.tr ? 1 : t.Qx);
        r.pos_n = g.swap ? t.ax.n : t.ay.n;
        r.pos_mask = (g.swap ? t.ax.sym : t.ay.sym) ? -1 : 0;
        return r;
    }
    XL_DEV void prepare() {
        if (PRO != XL_PRO_NONE) tp = tab_split(p.tpro, p.gpro, lb, p.nlines, PRO == XL_PRO_HIGHNA ? comp : 0);
        if (EPI == XL_EPI_RSF) te = tab_split(p.tepi, p.gepi, lb, p.nlines, 0);
#pragma unroll
        for (int l = 0; l < XL_V; ++l) {
            const int line = lb + l < p.nlines ? lb + l : 0;
            lcoord[l] = p.gpro.swap ? (float)(p.gpro.y0 + line * p.gpro.dy) : (float)(p.gpro.x0 + line * p.gpro.dx);
        }
        pos0 = p.gpro.swap ? p.gpro.x0 : p.gpro.y0;
        dpos = p.gpro.swap ? p.gpro.dx : p.gpro.dy;
    }
    XL_DEV cf fac(const XlFacTab& t, const Tab& tb, int l, int ipos) const {
        return xl_ldg(t.T + tb.line_part[l] + (long long)ipos * tb.pos_stride);
    }
    // raw operand(s) of both lines at position i (branch-free: out-of-range samples read a valid address, zeroed later)
    XL_DEV void fetch(const cf* src, int i, bool ok_i, cf* a) const {
        if (ACC == XL_ACC_PAIR_IN) {
            xl_ld4(src + (((long long)lb + (long long)i * p.in_pos) & -(long long)ok_i), a, a + 1);
        } else {
#pragma unroll
            for (int l = 0; l < XL_V; ++l) {
                const bool ok = ok_i && lb + l < p.nlines;
                // masked, not selected: the compiler turns a select around a 64-bit address into a branch, and a branch
                // here splits the first pass's loads into dependent groups (round 2: +18 % on the row-input passes)
                a[l] = src[((long long)(lb + l) * p.in_line + (long long)i * p.in_pos) & -(long long)ok];
            }
        }
    }
    XL_DEV void load(int i, cf* v, int stride) const {
        const bool ok_i = i < p.m_in;
        cf a[XL_V], b[XL_V];
        if (PRO == XL_PRO_NONE || PRO == XL_PRO_RSF) {
            fetch(p.in + (long long)cl * p.in_comp, i, ok_i, a);
        } else {
            fetch(p.in, i, ok_i, a);                 // Ex
            fetch(p.in + p.in_comp, i, ok_i, b);     // Ey
        }
        const cf pre = xl_ldg(p.pre + (ok_i ? i : 0));
        const int pi = ok_i ? i : 0;
        int ipos = 0;
        float spos = 1.f, cpos = 0.f;
        if (PRO != XL_PRO_NONE) {
            ipos = pos_idx(tp, pi);
            if (PRO == XL_PRO_HIGHNA) spos = pos_sign(tp, pi);
            if (PRO == XL_PRO_VCZT) cpos = (float)(pos0 + pi * dpos);
        }
#pragma unroll
        for (int l = 0; l < XL_V; ++l) {
            const int line = lb + l;
            cf x;
            if (PRO == XL_PRO_NONE) {
                x = (p.flags & XL_F_CONJ_IN) ? cf_conj(a[l]) : a[l];
            } else if (PRO == XL_PRO_RSF) {          // F = h(X, Y; z), wave_optics.py:341,344
                x = (p.flags & XL_F_CONJ_IN) ? cf_conj(a[l]) : a[l];
                x = cf_mul(x, fac(p.tpro, tp, l, ipos));
            } else if (PRO == XL_PRO_VCZT) {
                // comp 0/1: Ex / Ey;  comp 2: Ez = ((Ex X + Ey Y)/r) * z/r     vectorized_optics.py:341-344
                // (an amplitude factor: fp32 coordinates are enough; the phase lives in F)
                const float X = p.gpro.swap ? cpos : lcoord[l], Y = p.gpro.swap ? lcoord[l] : cpos;
                const float ir2 = xl_rcpf(X * X + Y * Y + z * z);   // MUFU.RCP: an IEEE division carries a slow-path branch that splits the loads
                // every component loads both planes: selecting, per CTA, only the plane components 0 / 1 need (the unused
                // operand repeating the used one) was measured in bench.py's per-kernel pass and lost, 174 against 161 us
                // for this pass at 2048^2 (profiles/ab_vczt_loads_r02w.txt)
                const float ax = comp == 0 ? 1.f : (comp == 1 ? 0.f : X * z * ir2);
                const float ay = comp == 0 ? 0.f : (comp == 1 ? 1.f : Y * z * ir2);
                x = cf_mul(cf_lin2(a[l], ax, b[l], ay), fac(p.tpro, tp, l, ipos));
            } else {  // XL_PRO_HIGHNA: row `comp` of apod*G*RL(theta,phi), stored for |x|,|y| with the parity of each entry
                const cf w = fac(p.tpro, tp, l, ipos);
                const float sx = p.gpro.swap ? spos : tp.sline[l], sy = p.gpro.swap ? tp.sline[l] : spos, sxy = sx * sy;
                const float ax = w.x * (comp == 0 ? 1.f : (comp == 1 ? sxy : sx));
                const float ay = w.y * (comp == 0 ? sxy : (comp == 1 ? 1.f : sy));
                x = cf_lin2(a[l], ax, b[l], ay);
            }
            // index weight of the d/dz chains; a multiply by 1 otherwise (a branch here, even a uniform one, splits the loads)
            x = cf_scale(x, p.in_weight == 0 ? 1.f : (float)(p.in_weight == 1 ? i : line));
            x = cf_mul(x, pre);
            v[l * stride] = (ok_i && line < p.nlines) ? x : cf_zero();
        }
    }
    XL_DEV void spec(int beta, cf* v) const {
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            const cf w = xl_ldg(p.ft + q * (L / 16) + beta);
#pragma unroll
            for (int l = 0; l < XL_V; ++l) v[l * 16 + q] = cf_mul(v[l * 16 + q], w);
        }
    }
    // output factor of (line, o): post-chirp * F0 * constant.  Gathered for ALL outputs of a thread before its first store:
    // the compiler may not move a load above a store to another pointer, and a load-use-store chain per output serialises
    // one L2 round trip per output (ncu round 2: 6 long-scoreboard stall cycles per issue in the epilogue kernels).
    XL_DEV cf out_factor(int l, int ipos, cf post) const {
        cf w = post;
        if (EPI == XL_EPI_RSF) w = cf_mul(w, fac(p.tepi, te, l, ipos));   // F0 = h(Xout, Yout; z), wave_optics.py:340,355
        return cf_mul(w, cst);
    }
    XL_DEV cf finish(cf val, cf w) const {
        val = cf_mul(val, w);
        return (p.flags & XL_F_CONJ_OUT) ? cf_conj(val) : val;
    }
    XL_DEV void store_vec(int n, const cf* v) const {
        constexpr int NJ = kOutLoHalf ? R1 / 2 : R1;
        constexpr int CH = NJ < 8 ? NJ : 8;          // outputs per gather-then-store chunk (2*CH factors live at once)
#pragma unroll
        for (int j0 = 0; j0 < NJ; j0 += CH) {
            cf w[XL_V * CH];
#pragma unroll
            for (int jj = 0; jj < CH; ++jj) {
                const int o = n + S1 * (j0 + jj) - p.out_off;
                const bool ok = o >= 0 && o < p.m_out;
                const cf post = xl_ldg(p.post + (ok ? o : 0));
                const int ipos = EPI == XL_EPI_RSF ? pos_idx(te, ok ? o : 0) : 0;
#pragma unroll
                for (int l = 0; l < XL_V; ++l) w[l * CH + jj] = out_factor(l, ipos, post);
            }
#pragma unroll
            for (int jj = 0; jj < CH; ++jj) {
                const int j = j0 + jj, o = n + S1 * j - p.out_off;
                if (o < 0 || o >= p.m_out) continue;
                cf* dst = p.out + (long long)cl * p.out_comp + (long long)o * p.out_pos;
                if (ACC == XL_ACC_PAIR_OUT) {
                    xl_st4(dst + lb, finish(v[j], w[jj]), finish(v[R1 + j], w[CH + jj]));
                } else {
#pragma unroll
                    for (int l = 0; l < XL_V; ++l)
                        if (lb + l < p.nlines) dst[(long long)(lb + l) * p.out_line] = finish(v[l * R1 + j], w[l * CH + jj]);
                }
            }
        }
    }
};
template <int L, int PRO, int EPI, int ACC> struct XlCztAxis {
    static const char* name() {   // "czt_axis<prologue,epilogue,access>": the profiling report keeps the variants apart
        static const char n[] = {'c', 'z', 't', '_', 'a', 'x', 'i', 's', '<', char('0' + PRO), ',', char('0' + EPI), ',', char('0' + ACC), '>', 0};
        return n;
    }
    typedef XlCztParams Params;
    static constexpr int NT = xl_threads(L);
    static size_t smem() { return xl_smem_bytes(L, XL_V); }
    // One CTA per (line pair, component).  A persistent variant (resident CTAs walking the items, twiddles and the kernel
    // spectrum staged once per CTA) was measured in round 2 and lost 20-35 %: without asynchronous staging of the next
    // item's input the two CTAs of an SM stay in lock-step, while CTAs handed out by the hardware start out of phase.
    XL_DEV static void run(const Params& p, cf* s) {
        cf* t = s + xl_tile_elems(L, XL_V);
        const double z = p.z ? xl_ldg(p.z) : 0.0;
        double cr = p.epi_cr, ci = p.epi_ci;
        if (p.epi_times_z) { cr *= z; ci *= z; }
        // component-minor CTA order: the CTAs that read the same (Ex, Ey) samples of a line pair run together (ncu round 2: the
        // vectorial first pass read 3.3x its algorithmic bytes from DRAM when blockIdx.y carried the component)
        const int cl = XL_BLOCK_X % p.ncomp, lb = (XL_BLOCK_X / p.ncomp) * XL_V;
        XlCztOp<L, PRO, EPI, ACC> op{{}, p, lb, cl, p.c0 + cl, (float)z, make_float2((float)cr, (float)ci)};
        op.prepare();
        XlFft<L, XL_V>::conv_g(s, t, p.tw, op);
    }
};


// Bluestein tables for one axis (wave_optics.py:385-410, 430-459), all phases in fp64:
//   pre[k]  = A^-k * h_k                       h_j = W^(j^2/2) on the principal branch of log W
//   post[l] = h_l * exp(-i 2 pi f_l (1/2 - m/2)/Dm)
//   ft      = FFT_L(1/h[:mp+1]) / L            (slot order)         -- forward pass multiplier
//   ftT     = FFT_L(transposed kernel) / L     (slot order)         -- adjoint pass multiplier
// Two launches: czt_tables evaluates every table ENTRY (chirps, the two kernel sequences `kin`, the factor tables) with
// one thread per entry over as many CTAs as it takes; czt_kernel_fft (one CTA per axis) transforms kin -> ft, ftT.
struct XlCztAxisTab {
    int L, m, M;
    double out0, outl;        // first / last output coordinate of this axis
    double Dm_static;         // used when z == null
    double lambda_over_dx;    // Dm = lambda*z/dx   (wave_optics.py:322)
    cf* pre; cf* post; cf* kin; cf* ft; cf* ftT;     // kin: [2][L] = {forward kernel sequence, adjoint kernel sequence}
};
struct XlCztAxisConsts { double Dm, D1, D2, uw; int Lh; };
XL_DEV XlCztAxisConsts xl_czt_consts(const XlCztAxisTab& p, const double* z) {
    XlCztAxisConsts a;
    a.Dm = z ? p.lambda_over_dx * xl_ldg(z) : p.Dm_static;
    const double f1 = p.out0 + a.Dm / 2, f2 = p.outl + a.Dm / 2;          // wave_optics.py:325-329
    a.D1 = f1 + (p.M * a.Dm + f2 - f1) / (2 * p.M);                       // :431
    a.D2 = f2 + (p.M * a.Dm + f2 - f1) / (2 * p.M);                       // :433
    double u = -(a.D2 - a.D1) / (p.M * a.Dm);                             // W = exp(2 pi i u), :393
    a.uw = u - rint(u);                                                   // principal branch of log W
    const int hl = p.m - 1 + (p.M > p.m ? p.M : p.m);                     // len(h), :396
    a.Lh = hl < p.m + p.M ? hl : p.m + p.M;                               // len(h[:mp+1]), :398
    return a;
}
XL_DEV cf xl_chirp(double uw, double j, double sign) {   // exp(sign * 2 pi i * uw * j^2/2)
    double cyc = uw * (j * j) * 0.5;
    cyc -= rint(cyc);
    double s, c;
    xl_sincospi(2.0 * cyc, &s, &c);
    return make_float2((float)c, (float)(sign * s));
}
// entry i of kernel sequence c (0: forward, 1: adjoint) of an axis
XL_DEV cf xl_czt_kin(const XlCztAxisTab& p, const XlCztAxisConsts& a, int c, int i) {
    const int L = p.L;
    int t;
    if (c == 0) t = (i + p.m) & (L - 1);   // kernel rotated left by m: the kept slice b[m : m+M] (wave_optics.py:446) lands
                                           // at positions [0, M) of the inverse transform, i.e. in its prunable lower half
    else {
        int sft;
        if (i <= p.m - 1) sft = i; else if (i >= L - (p.M - 1)) sft = i - L; else return cf_zero();
        t = p.m - sft;
    }
    if (t < 0 || t >= a.Lh) return cf_zero();
    return xl_chirp(a.uw, (double)(t - (p.m - 1)), -1.0);  // 1/h_j = conj(h_j)
}
enum { XL_FAC_NONE = 0, XL_FAC_RS = 1, XL_FAC_LENS = 2 };
struct XlCztTablesParams {
    XlCztAxisTab a[2];
    const double* z; double k;
    int fac_kind[3];           // what the factor tables hold (XL_FAC_*): [0] input grid, [1] output grid, [2] input grid again
                               // in the other orientation (high-NA: the pointwise fold reads it row-major)
    cf* T[3]; int Qx[3], Qy[3], tr[3]; XlFacAxis fx[3], fy[3];
    double lens_R, lens_f, lens_s2;
    long long seg[10];         // cumulative entry counts: pre_y post_y kin_y | pre_x post_x kin_x | T0 T1 T2 | end
};
struct XlCztTables {
    static const char* name() { return "czt_tables"; }
    typedef XlCztTablesParams Params;
    static constexpr int NT = 256;
    static size_t smem() { return 16; }
    XL_DEV static void run(const Params& p, cf*) {
        const double z = p.z ? xl_ldg(p.z) : 0.0;
        XL_THREADS(tid, NT) {
            const long long e = (long long)XL_BLOCK_X * NT + tid;
            if (e < p.seg[6]) {                                     // chirp tables of one axis
                const int ax = e < p.seg[3] ? 0 : 1;
                const XlCztAxisTab& t = p.a[ax];
                const XlCztAxisConsts a = xl_czt_consts(t, p.z);
                const long long base = ax ? p.seg[3] : 0;
                const long long r = e - base;
                const long long n_pre = p.seg[ax * 3 + 1] - base, n_post = p.seg[ax * 3 + 2] - base;
                if (r < n_pre) {                                    // pre[k]
                    const int k = (int)r;
                    double cyc = a.D1 / a.Dm;
                    cyc -= rint(cyc);
                    double ph = -(double)k * cyc;
                    ph -= rint(ph);
                    double sn, cs;
                    xl_sincospi(2.0 * ph, &sn, &cs);
                    t.pre[k] = cf_mul(make_float2((float)cs, (float)sn), xl_chirp(a.uw, (double)k, 1.0));
                } else if (r < n_post) {                            // post[l]
                    const int l = (int)(r - n_pre);
                    double fl = (double)l / t.M * (a.D2 - a.D1) + a.D1;              // :451-453
                    double ph = -fl * (-(double)t.m / 2 + 0.5) / a.Dm;                // :456-457
                    ph -= rint(ph);
                    double sn, cs;
                    xl_sincospi(2.0 * ph, &sn, &cs);
                    t.post[l] = cf_mul(make_float2((float)cs, (float)sn), xl_chirp(a.uw, (double)l, 1.0));
                } else {                                            // kin[c][i]
                    const long long q = r - n_post;
                    t.kin[q] = xl_czt_kin(t, a, (int)(q / t.L), (int)(q % t.L));
                }
            } else if (e < p.seg[9]) {                              // factor tables
                const int w = e < p.seg[7] ? 0 : (e < p.seg[8] ? 1 : 2);
                const long long r = e - p.seg[6 + w];
                const int Qx = p.Qx[w], Qy = p.Qy[w];
                const int c = (int)(r / ((long long)Qx * Qy));
                const int ix = p.tr[w] ? (int)((r / Qy) % Qx) : (int)(r % Qx), iy = p.tr[w] ? (int)(r % Qy) : (int)((r / Qx) % Qy);
                const double X = xl_fac_coord(p.fx[w], ix), Y = xl_fac_coord(p.fy[w], iy);
                if (p.fac_kind[w] == XL_FAC_RS) {
                    p.T[w][r] = xl_rs_h(X, Y, xl_rs_hconst(z, p.k), 0);
                } else {
                    float ax, ay;
                    xl_lens_row(X, Y, p.lens_R, p.lens_f, p.lens_s2, c, &ax, &ay);
                    p.T[w][r] = make_float2(ax, ay);
                }
            }
        }
    }
};
template <int L> struct XlCztKernelFftOp : XlOpBase {
    const XlCztAxisTab& p;
    XL_DEV void load(int i, cf* v, int stride) const { v[0] = xl_ldg(p.kin + i); v[stride] = xl_ldg(p.kin + L + i); }
    XL_DEV void spec(int beta, const cf* v) const {
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            p.ft[q * (L / 16) + beta] = cf_scale(v[q], 1.0f / L);
            p.ftT[q * (L / 16) + beta] = cf_scale(v[16 + q], 1.0f / L);
        }
    }
    XL_DEV void store_vec(int, const cf*) const {}
};
struct XlCztKernelFftParams { XlCztAxisTab a[2]; const cf* tw; };   // one CTA per axis (blockIdx.x), both axes in one launch
template <int L> struct XlCztKernelFft {
    static const char* name() { return "czt_kernel_fft"; }
    typedef XlCztKernelFftParams Params;
    static constexpr int NT = xl_threads(L);
    static size_t smem() { return xl_smem_bytes(L, XL_V); }
    XL_DEV static void run(const Params& pp, cf* s) {
        const XlCztAxisTab& p = pp.a[XL_BLOCK_X];
        cf* t = s + xl_tile_elems(L, XL_V);
        XlFft<L, XL_V>::init_tw(t, pp.tw);
        XlCztKernelFftOp<L> op{{}, p};
        XlFft<L, XL_V>::forward(s, t, op);
    }
};

// ==================================================================================================================
// d/dz of CZT / VCZT.  JAX differentiates CZT_jit through F, F0, the constant z dx dy lambda and, via
// Dm = lambda z / dx, through every Bluestein chirp (wave_optics.py:322, 340-355, 393-403).  In closed form the transform of
// one axis is K[l,k] = exp(i Phi(l,k)) on its valid entries, with (the off-by-one slice included)
//     Phi = theta ((l+1) k - l - 1/2) - 2 pi k D1/Dm - 2 pi f_l (1/2 - m/2)/Dm,   theta = -2 pi Delta/Dm (principal branch),
//     D1 = out0 + Dm + Delta/2,  f_l = l Delta + D1,  Delta = (out_last - out0)/M,
// so   dPhi/dDm = alpha l k + beta k + gamma l + delta   is bilinear in (l, k):
//     alpha = 2 pi Delta/Dm^2,  beta = 2 pi (out0 + 3 Delta/2)/Dm^2,  gamma = -pi Delta (m+1)/Dm^2,
//     delta = 2 pi (-Delta/2 + (1/2 - m/2)(out0 + Delta/2))/Dm^2.
// With A the whole forward operator (out = A U), O1 = A(k_y U) and O2 = A(k_x U) (two more forward chains with an index
// weight on the input, XlCztParams::in_weight) give
//     d out/dz = out [d ln F0/dz + 1/z + i Dm' (gamma_y l_y + delta_y + gamma_x l_x + delta_x)]
//              + i Dm' [(alpha_y l_y + beta_y) O1 + (alpha_x l_x + beta_x) O2]  +  A((d ln F/dz) U),        Dm' = lambda/dx,
// and the last term is folded onto the input side with the field cotangent:  sum ct_out A(D U) = sum ct_in D U.
// This kernel forms  gz += Re[ sum_out ct_out (...) + sum_in ct_in (d ln F/dz) U ]  in fp64 (+ the dEz/dz term of VCZT).
// ==================================================================================================================
// d ln h / dz of the RS factor h = (1/2pi)(z/r^2)(1/r - i k) exp(sgn(z) i k r)   (wave_optics.py:291-297)
XL_DEV void xl_rs_dlnh(double X, double Y, double z, double k, double* re, double* im) {
    const double r2 = X * X + Y * Y + z * z, r = sqrt(r2), ir = 1.0 / r;
    // d/dz ln(1/r - i k) = (-z/r^3) / (1/r - i k) = (-z/r^3) (1/r + i k) / (1/r^2 + k^2)
    const double q = -z * ir * ir * ir / (ir * ir + k * k);
    const double sg = z > 0 ? 1.0 : -1.0;
    *re = 1.0 / z - 2.0 * z / r2 + q * ir;
    *im = q * k + sg * k * z * ir;
}
struct XlCztAxisDz { double out0, Delta; int m; };
struct XlCztDotZParams {
    int N, Mx, My, ncomp, flags;   // XL_F_CONJ_IN: ct_out is conjugated on load; XL_F_CONJ_OUT: ct_in was stored conjugated
    const cf* ct_out; const cf* out; const cf* O1; const cf* O2;   // [ncomp][My][Mx]
    const cf* ct_in;     // [ncomp][N][N] cotangent of the component planes (scalar: the caller's ct_in; vectorial: before the fold)
    const cf* in;        // [N][N], or the Ex plane with the Ey plane ey_off elements later
    long long ey_off;
    const double* z; double k, dDm_dz, lambda_over_dx;
    XlCztAxisDz ay, ax;
    double x0, dx, y0, dy, xo0, dxo, yo0, dyo;
    double* gz;
};
struct XlCztDotZ {
    static const char* name() { return "czt_dot_z"; }
    typedef XlCztDotZParams Params;
    static constexpr int NT = 256;
    static size_t smem() { return NT * sizeof(double); }
    XL_DEV static void run(const Params& p, cf* s) {
        double* red = (double*)s;
        const double z = xl_ldg(p.z), Dm = p.lambda_over_dx * z, w = 6.283185307179586 / (Dm * Dm);
        const double a_y = w * p.ay.Delta, b_y = w * (p.ay.out0 + 1.5 * p.ay.Delta), g_y = -0.5 * w * p.ay.Delta * (p.ay.m + 1),
                     d_y = w * (-0.5 * p.ay.Delta + (0.5 - 0.5 * p.ay.m) * (p.ay.out0 + 0.5 * p.ay.Delta));
        const double a_x = w * p.ax.Delta, b_x = w * (p.ax.out0 + 1.5 * p.ax.Delta), g_x = -0.5 * w * p.ax.Delta * (p.ax.m + 1),
                     d_x = w * (-0.5 * p.ax.Delta + (0.5 - 0.5 * p.ax.m) * (p.ax.out0 + 0.5 * p.ax.Delta));
        const size_t MM = (size_t)p.My * p.Mx, NN = (size_t)p.N * p.N;
        XL_THREADS(tid, NT) {
            const size_t e = (size_t)XL_BLOCK_X * NT + tid;
            double acc = 0.0;
            if (e < MM * p.ncomp) {                                   // output side
                const size_t o = e % MM;
                const int ly = (int)(o / p.Mx), lx = (int)(o % p.Mx);
                double cr = (double)p.ct_out[e].x, ci = (double)p.ct_out[e].y;
                if (p.flags & XL_F_CONJ_IN) ci = -ci;
                double fr, fi;
                xl_rs_dlnh(p.xo0 + lx * p.dxo, p.yo0 + ly * p.dyo, z, p.k, &fr, &fi);
                fr += 1.0 / z;
                fi += p.dDm_dz * (g_y * ly + d_y + g_x * lx + d_x);
                const double w1 = p.dDm_dz * (a_y * ly + b_y), w2 = p.dDm_dz * (a_x * lx + b_x);
                const cf u = p.out[e], o1 = p.O1[e], o2 = p.O2[e];
                // t = out (fr + i fi) + i (w1 O1 + w2 O2)
                const double tr = (double)u.x * fr - (double)u.y * fi - (w1 * (double)o1.y + w2 * (double)o2.y);
                const double ti = (double)u.x * fi + (double)u.y * fr + (w1 * (double)o1.x + w2 * (double)o2.x);
                acc += cr * tr - ci * ti;                             // Re(ct * t)
            }
            if (e < NN * p.ncomp) {                                   // input side
                const int c = (int)(e / NN);
                const size_t o = e % NN;
                const int iy = (int)(o / p.N), ix = (int)(o % p.N);
                const double X = p.x0 + ix * p.dx, Y = p.y0 + iy * p.dy;
                double cr = (double)p.ct_in[e].x, ci = (double)p.ct_in[e].y;
                if (p.flags & XL_F_CONJ_OUT) ci = -ci;
                double ur, ui, er = 0.0, ei = 0.0;                    // component plane U_c and (c == 2) dEz/dz
                if (c < 2) {
                    const cf a = p.in[(long long)c * p.ey_off + (long long)o];
                    ur = (double)a.x; ui = (double)a.y;
                } else {                                              // Ez = (Ex X + Ey Y) z / r^2, vectorized_optics.py:341-344
                    const cf ex = p.in[o], ey = p.in[p.ey_off + (long long)o];
                    const double ir2 = 1.0 / (X * X + Y * Y + z * z);
                    const double sr = (double)ex.x * X + (double)ey.x * Y, si = (double)ex.y * X + (double)ey.y * Y;
                    ur = sr * z * ir2; ui = si * z * ir2;
                    const double dz = ir2 - 2.0 * z * z * ir2 * ir2;
                    er = sr * dz; ei = si * dz;
                }
                double fr, fi;
                xl_rs_dlnh(X, Y, z, p.k, &fr, &fi);
                const double tr = ur * fr - ui * fi + er, ti = ur * fi + ui * fr + ei;
                acc += cr * tr - ci * ti;
            }
            red[tid] = acc;
        }
        XL_SYNC();
        xl_block_sum<NT>(red);
        XL_THREADS(tid, NT) {
            if (tid == 0) xl_atomic_add(p.gz, red[0]);
        }
    }
};

// ==================================================================================================================
// Pointwise adjoint folds (the transposes of the Ez formation / lens matrix), one thread per pixel.
// ==================================================================================================================
enum { XL_FOLD_VRS = 0, XL_FOLD_VCZT = 1, XL_FOLD_HIGHNA = 2 };
struct XlFoldParams {
    int N, mode, flags;
    const cf* t;        // [3][N][N] adjoint of the three propagated components (JAX convention)
    const cf* ex; const cf* ey;   // primal Ex,Ey (VRS: for the dEz/dz term), may be null
    cf* gx; cf* gy;     // cotangents of Ex, Ey
    double* gz;         // += Re sum t_z * dEz/dz  (VRS only)
    const double* z;
    double x0, y0, dx, dy;
    XlFacTab lens;      // high-NA: the lens-matrix table of czt_tables
};
struct XlFold {
    static const char* name() { return "fold"; }
    typedef XlFoldParams Params;
    static constexpr int NT = 256;
    static size_t smem() { return NT * sizeof(float); }
    XL_DEV static void run(const Params& p, cf* s) {
        float* red = (float*)s;
        const double z = p.z ? xl_ldg(p.z) : 0.0;
        const size_t NN = (size_t)p.N * p.N;
        XL_THREADS(tid, NT) {
            const size_t idx = (size_t)XL_BLOCK_X * NT + tid;
            float acc = 0.f;
            if (idx < NN) {
                const int y = (int)(idx / p.N), x = (int)(idx % p.N);
                const double X = p.x0 + x * p.dx, Y = p.y0 + y * p.dy;
                cf t0 = p.t[idx], t1 = p.t[NN + idx], t2 = p.t[2 * NN + idx];
                cf gx, gy;
                if (p.mode == XL_FOLD_VRS) {
                    double ir = xl_rsqrt64(X * X + Y * Y + z * z);
                    float ax = (float)(X * ir), ay = (float)(Y * ir);
                    gx = make_float2(t0.x + ax * t2.x, t0.y + ax * t2.y);
                    gy = make_float2(t1.x + ay * t2.x, t1.y + ay * t2.y);
                    if (p.gz) {  // dEz/dz = -(Ex X + Ey Y) z / r^3
                        cf ex = p.ex[idx], ey = p.ey[idx];
                        if (p.flags & XL_F_CONJ_IN) { ex.y = -ex.y; ey.y = -ey.y; }
                        double w3 = -z * ir * ir * ir;
                        float bx = (float)(X * w3), by = (float)(Y * w3);
                        cf d = make_float2(ex.x * bx + ey.x * by, ex.y * bx + ey.y * by);
                        acc = t2.x * d.x - t2.y * d.y;
                    }
                } else if (p.mode == XL_FOLD_VCZT) {
                    double ir2 = 1.0 / (X * X + Y * Y + z * z);
                    float ax = (float)(X * z * ir2), ay = (float)(Y * z * ir2);
                    gx = make_float2(t0.x + ax * t2.x, t0.y + ax * t2.y);
                    gy = make_float2(t1.x + ay * t2.x, t1.y + ay * t2.y);
                } else {
                    const size_t e = xl_fac_entry(p.lens, xl_fac_idx(p.lens.ax, x), xl_fac_idx(p.lens.ay, y), 0), pl = (size_t)p.lens.Qx * p.lens.Qy;
                    const cf w0 = xl_ldg(p.lens.T + e), w1 = xl_ldg(p.lens.T + pl + e), w2 = xl_ldg(p.lens.T + 2 * pl + e);
                    const float sx = xl_fac_sign(p.lens.ax, x), sy = xl_fac_sign(p.lens.ay, y), sxy = sx * sy;
                    const float a0x = w0.x, a0y = w0.y * sxy, a1x = w1.x * sxy, a1y = w1.y, a2x = w2.x * sx, a2y = w2.y * sy;
                    gx = make_float2(a0x * t0.x + a1x * t1.x + a2x * t2.x, a0x * t0.y + a1x * t1.y + a2x * t2.y);
                    gy = make_float2(a0y * t0.x + a1y * t1.x + a2y * t2.x, a0y * t0.y + a1y * t1.y + a2y * t2.y);
                }
                if (p.flags & XL_F_CONJ_OUT) { gx.y = -gx.y; gy.y = -gy.y; }
                p.gx[idx] = gx;
                p.gy[idx] = gy;
            }
            red[tid] = acc;
        }
        XL_SYNC();
        if (p.gz) {   // kernel-uniform
            xl_block_sum<NT>(red);
            XL_THREADS(tid, NT) {
                if (tid == 0) xl_atomic_add(p.gz, (double)red[0]);
            }
        }
    }
};
