// xl_kernels.cuh -- kernel bodies of the propagation hot path (RS/VRS convolution, Bluestein CZT/VCZT, high-NA lens).
//
// Every kernel is "one CTA = one tile of V=4 lines through XlFft"; what differs is the functor (op) that feeds the first
// pass, multiplies the spectrum in registers and drains the last pass.  Reference lines replaced are cited per op.
//
// Data layouts (c64 everywhere, geometry in fp64):
//   field            [f][y][x]                     row-major N x N planes (the reference's [..., y, x])
//   row spectra  S   [f][L/4][N][4]                "blocked": 4 adjacent x-slots innermost, so the column kernel reads one
//                                                   contiguous 32*N-byte tile and the row kernel writes 128-byte lines
//   transfer fn  H   [L/4][L][4]                   same blocking, y in slot order, premultiplied by dx*dy/L^2
//   L = padded length (power of two >= 2N-1); slot order = XlFft's digit permutation (never undone).
#pragma once
#include "xl_fft.cuh"

#define XL_CW 4  // lines per tile == column-group width of the blocked layouts

// flags shared by several kernels
#define XL_F_CONJ_IN 1    // conjugate operand on load   (torch's conjugate-cotangent convention, fused)
#define XL_F_CONJ_OUT 2   // conjugate result on store
#define XL_F_VRS 4        // 3 fields, field 2 = Ez formed from (Ex,Ey) at load      (vectorized_optics.py:258-261)
#define XL_F_DERIV 8      // transfer function of dh/dz instead of h

// ------------------------------------------------------------------------------------------------------------------
// Rayleigh-Sommerfeld impulse response, reference wave_optics.py:291-297:
//   h = (1/2pi) * z/r^2 * (1/r - i k) * exp(sgn(z) i k r),  r = sqrt(X^2+Y^2+z^2)
// Phase k*r reaches 4e5 rad: reduced in fp64 (r/lambda - rint) and only the fraction goes through fp32 sincospi.
// deriv=1: dh/dz = (e/2pi) * [ g + (z^2/r) * (g' + s i k g) ],  g = 1/r^3 - i k/r^2,  g' = -3/r^4 + 2 i k/r^3.
// ------------------------------------------------------------------------------------------------------------------
XL_DEV cf xl_rs_h(double X, double Y, double z, double k, int deriv) {
    const double inv2pi = 0.15915494309189535;
    double r2 = X * X + Y * Y + z * z;
    double r = sqrt(r2);
    double cyc = r * (k * inv2pi);
    double fr = cyc - rint(cyc);
    float sn, cs;
    xl_sincospif(2.0f * (float)fr, &sn, &cs);
    double sg = z > 0 ? 1.0 : -1.0;
    double ir = 1.0 / r, ir2 = ir * ir, ir3 = ir2 * ir;
    double ar, ai;  // amplitude (complex), multiplies exp(s i k r)
    if (!deriv) {
        ar = z * inv2pi * ir3;
        ai = -z * inv2pi * k * ir2;
    } else {
        double gr = ir3, gi = -k * ir2;
        double gpr = -3.0 * ir2 * ir2, gpi = 2.0 * k * ir3;
        // g' + s*i*k*g = (gpr - s*k*gi) + i (gpi + s*k*gr)
        double tr = gpr - sg * k * gi, ti = gpi + sg * k * gr;
        double f = z * z * ir;
        ar = (gr + f * tr) * inv2pi;
        ai = (gi + f * ti) * inv2pi;
    }
    float c = cs, s = (float)sg * sn;
    float far = (float)ar, fai = (float)ai;
    return make_float2(far * c - fai * s, far * s + fai * c);
}

// ==================================================================================================================
// RS / VRS  (wave_optics.py:281-289, vectorized_optics.py:364-373)
// ==================================================================================================================
struct XlRsParams {
    int N, L, nfields, flags;
    const cf* in;      // [nfields][N][N]   (XL_F_VRS: [2][N][N] = Ex,Ey)
    cf* out;           // [nfields][N][N]
    cf* spec;          // [nfields][L/4][N][4]
    cf* spec2;         // second spectra set (grad-z: spectra of conj(U))
    cf* H;             // [L/4][L][4]
    const cf* H2;      // dH/dz transfer function (grad-z)
    cf* scratch;       // [L/4][nfields][L][4] (grad-z)
    double* gz;        // scalar accumulator (grad-z)
    const cf* tw;
    const double* z;   // device scalar
    double x0, y0, dx, dy, k;
    float hscale;      // dx*dy/L^2
};

// K1: rows of the zero-padded field -> blocked row spectra.   replaces the row half of fft2(U), wave_optics.py:286-288
template <int L> struct XlRsRowsFwdOp {
    const XlRsParams& p; int f, yb; double z;
    XL_DEV cf load(int c, int i) const {
        const int y = yb + c, N = p.N;
        if (y >= N || i >= N) return make_float2(0.f, 0.f);
        const size_t NN = (size_t)N * N, o = (size_t)y * N + i;
        cf v;
        if ((p.flags & XL_F_VRS) && f == 2) {
            cf ex = p.in[o], ey = p.in[NN + o];
            double X = p.x0 + i * p.dx, Y = p.y0 + y * p.dy;
            double ir = 1.0 / sqrt(X * X + Y * Y + z * z);
            float ax = (float)(X * ir), ay = (float)(Y * ir);
            v = make_float2(ex.x * ax + ey.x * ay, ex.y * ax + ey.y * ay);
        } else {
            v = p.in[(size_t)f * NN + o];
        }
        if (p.flags & XL_F_CONJ_IN) v.y = -v.y;
        return v;
    }
    XL_DEV void spec(int c, int beta, cf* v) const {
        const int y = yb + c;
        if (y >= p.N) return;
        cf* base = p.spec + (size_t)f * L * p.N + (size_t)y * XL_CW;
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            const int g = q * (L / 16) + beta;
            base[(size_t)(g / XL_CW) * p.N * XL_CW + (g % XL_CW)] = v[q];
        }
    }
    XL_DEV void store(int, int, cf) const {}
};
template <int L> struct XlRsRowsFwd {
    static const char* name() { return "rs_rows_fwd"; }
    typedef XlRsParams Params;
    static constexpr int NT = xl_threads(L, XL_CW);
    XL_DEV static void run(const Params& p, cf* s) {
        XlRsRowsFwdOp<L> op{p, XL_BLOCK_Y, XL_BLOCK_X * XL_CW, p.z ? xl_ldg(p.z) : 0.0};
        XlFft<L, XL_CW, NT>::forward(s, p.tw, op);
    }
};

// K2: column FFT of the row spectra, x transfer function, inverse column FFT, keep rows [0,N).   wave_optics.py:288
template <int L> struct XlRsColsOp {
    const XlRsParams& p; cf* tile; const cf* Ht;
    XL_DEV cf load(int c, int i) const { return i < p.N ? tile[(size_t)i * XL_CW + c] : make_float2(0.f, 0.f); }
    XL_DEV void spec(int c, int beta, cf* v) const {
#pragma unroll
        for (int q = 0; q < 16; ++q) v[q] = cf_mul(v[q], xl_ldg(Ht + (size_t)(q * (L / 16) + beta) * XL_CW + c));
    }
    XL_DEV void store(int c, int i, cf val) const { if (i < p.N) tile[(size_t)i * XL_CW + c] = val; }
};
template <int L> struct XlRsCols {
    static const char* name() { return "rs_cols"; }
    typedef XlRsParams Params;
    static constexpr int NT = xl_threads(L, XL_CW);
    XL_DEV static void run(const Params& p, cf* s) {
        const int G = XL_BLOCK_X, f = XL_BLOCK_Y;
        XlRsColsOp<L> op{p, p.spec + (size_t)f * L * p.N + (size_t)G * p.N * XL_CW, p.H + (size_t)G * L * XL_CW};
        XlFft<L, XL_CW, NT>::conv(s, p.tw, op);
    }
};

// K3: inverse row FFT of the filtered spectra, crop columns [0,N).   wave_optics.py:288 (row half of ifft2 + crop)
template <int L> struct XlRsRowsInvOp {
    const XlRsParams& p; int f, yb;
    XL_DEV cf load(int, int) const { return make_float2(0.f, 0.f); }
    XL_DEV void spec(int c, int beta, cf* v) const {
        const int y = yb + c;
        const cf* base = p.spec + (size_t)f * L * p.N + (size_t)y * XL_CW;
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            const int g = q * (L / 16) + beta;
            v[q] = y < p.N ? base[(size_t)(g / XL_CW) * p.N * XL_CW + (g % XL_CW)] : make_float2(0.f, 0.f);
        }
    }
    XL_DEV void store(int c, int i, cf val) const {
        const int y = yb + c;
        if (y >= p.N || i >= p.N) return;
        if (p.flags & XL_F_CONJ_OUT) val.y = -val.y;
        p.out[(size_t)f * p.N * p.N + (size_t)y * p.N + i] = val;
    }
};
template <int L> struct XlRsRowsInv {
    static const char* name() { return "rs_rows_inv"; }
    typedef XlRsParams Params;
    static constexpr int NT = xl_threads(L, XL_CW);
    XL_DEV static void run(const Params& p, cf* s) {
        XlRsRowsInvOp<L> op{p, XL_BLOCK_Y, XL_BLOCK_X * XL_CW};
        XlFft<L, XL_CW, NT>::inverse(s, p.tw, op);
    }
};

// K1h: rows y>=0 of the wrapped, analytically generated impulse response -> row spectra stored inside H
//      (rows 0..L/2 of each group block).   replaces transfer_function_RS + row half of fft2(H), wave_optics.py:285,291-297
template <int L> struct XlHRowsOp {
    const XlRsParams& p; int yb; double z;
    XL_DEV cf load(int c, int i) const {
        const int yi = yb + c;
        if (yi > L / 2) return make_float2(0.f, 0.f);
        const int xi = i <= L / 2 ? i : i - L;
        return xl_rs_h(xi * p.dx, yi * p.dy, z, p.k, (p.flags & XL_F_DERIV) ? 1 : 0);
    }
    XL_DEV void spec(int c, int beta, cf* v) const {
        const int yi = yb + c;
        if (yi > L / 2) return;
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            const int g = q * (L / 16) + beta;
            p.H[(size_t)(g / XL_CW) * L * XL_CW + (size_t)yi * XL_CW + (g % XL_CW)] = v[q];
        }
    }
    XL_DEV void store(int, int, cf) const {}
};
template <int L> struct XlHRows {
    static const char* name() { return "h_rows"; }
    typedef XlRsParams Params;
    static constexpr int NT = xl_threads(L, XL_CW);
    XL_DEV static void run(const Params& p, cf* s) {
        XlHRowsOp<L> op{p, XL_BLOCK_X * XL_CW, xl_ldg(p.z)};
        XlFft<L, XL_CW, NT>::forward(s, p.tw, op);
    }
};

// K2h: column FFT of the impulse-response row spectra (even in y: row L-y == row y), result in slot order, scaled.
template <int L> struct XlHColsOp {
    const XlRsParams& p; cf* Ht;
    XL_DEV cf load(int c, int i) const { const int r = i <= L / 2 ? i : L - i; return Ht[(size_t)r * XL_CW + c]; }
    XL_DEV void spec(int c, int beta, cf* v) const {
#pragma unroll
        for (int q = 0; q < 16; ++q) Ht[(size_t)(q * (L / 16) + beta) * XL_CW + c] = cf_scale(v[q], p.hscale);
    }
    XL_DEV void store(int, int, cf) const {}
};
template <int L> struct XlHCols {
    static const char* name() { return "h_cols"; }
    typedef XlRsParams Params;
    static constexpr int NT = xl_threads(L, XL_CW);
    XL_DEV static void run(const Params& p, cf* s) {
        XlHColsOp<L> op{p, p.H + (size_t)XL_BLOCK_X * L * XL_CW};
        XlFft<L, XL_CW, NT>::forward(s, p.tw, op);
    }
};

// K4: backward column kernel.  Phase A: column spectra W of conj(U) (spec2) parked in scratch.  Phase B: column spectra C
// of the cotangent; accumulates Re sum conj(W)*C*Hz (Parseval form of ct_z) and applies H for ct_field.
template <int L> struct XlRsColsWOp {
    const XlRsParams& p; const cf* tile; cf* scr;
    XL_DEV cf load(int c, int i) const { return i < p.N ? tile[(size_t)i * XL_CW + c] : make_float2(0.f, 0.f); }
    XL_DEV void spec(int c, int beta, cf* v) const {
#pragma unroll
        for (int q = 0; q < 16; ++q) scr[(size_t)(q * (L / 16) + beta) * XL_CW + c] = v[q];
    }
    XL_DEV void store(int, int, cf) const {}
};
template <int L> struct XlRsColsGzOp {
    const XlRsParams& p; cf* tile; const cf* Ht; const cf* Hzt; const cf* scr; float* red;
    XL_DEV cf load(int c, int i) const { return i < p.N ? tile[(size_t)i * XL_CW + c] : make_float2(0.f, 0.f); }
    XL_DEV void spec(int c, int beta, cf* v) const {
        float acc = 0.f;
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            const size_t o = (size_t)(q * (L / 16) + beta) * XL_CW + c;
            cf w = scr[o];
            cf t = cf_mul(v[q], xl_ldg(Hzt + o));
            acc += w.x * t.x + w.y * t.y;  // Re(conj(w) * t)
            v[q] = cf_mul(v[q], xl_ldg(Ht + o));
        }
        red[beta * XL_CW + c] = acc;
    }
    XL_DEV void store(int c, int i, cf val) const { if (i < p.N) tile[(size_t)i * XL_CW + c] = val; }
};
template <int L> struct XlRsColsGz {
    static const char* name() { return "rs_cols_gz"; }
    typedef XlRsParams Params;
    static constexpr int NT = xl_threads(L, XL_CW);
    static constexpr int NRED = (L / 16) * XL_CW;
    static constexpr int TILE = xl_tile_elems(L, XL_CW);
    XL_DEV static void run(const Params& p, cf* s) {
        const int G = XL_BLOCK_X, f = XL_BLOCK_Y;
        float* red = (float*)(s + TILE);
        cf* scr = p.scratch + ((size_t)G * p.nfields + f) * L * XL_CW;
        const size_t toff = (size_t)f * L * p.N + (size_t)G * p.N * XL_CW;
        XlRsColsWOp<L> opw{p, p.spec2 + toff, scr};
        XlFft<L, XL_CW, NT>::forward(s, p.tw, opw);
        XL_SYNC();
        XlRsColsGzOp<L> op{p, p.spec + toff, p.H + (size_t)G * L * XL_CW, p.H2 + (size_t)G * L * XL_CW, scr, red};
        XlFft<L, XL_CW, NT>::conv(s, p.tw, op);
        XL_SYNC();
        XL_THREADS(tid, NT) {
            if (tid < 32) {
                float a = 0.f;
                for (int i = tid; i < NRED; i += 32) a += red[i];
                red[NRED + tid] = a;
            }
        }
        XL_SYNC();
        XL_THREADS(tid, NT) {
            if (tid == 0) {
                double a = 0.0;
                for (int i = 0; i < 32; ++i) a += (double)red[NRED + i];
                xl_atomic_add(p.gz, a);
            }
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

struct XlCztParams {
    int L, nlines, ncomp, m_in, out_off, m_out, flags;
    const cf* in; long long in_line, in_pos, in_comp;
    cf* out;      long long out_line, out_pos, out_comp;
    const cf* pre; const cf* ft; const cf* post;
    const cf* tw;
    int pro, epi;
    XlGridFactor gpro, gepi;
    const double* z;      // device scalar (RSF / VCZT factors), may be null
    double k;             // wavenumber
    double epi_cr, epi_ci; // complex constant on the output
    int epi_times_z;      // multiply the constant by z (CZT: z*dx*dy*lambda)
    double lens_R, lens_f, lens_s2;  // high-NA: radius, focal length, sin^2(theta_max)
};

// lens factor row `comp` applied to (Ex,Ey):  apod*G*(RL[comp][0] Ex + RL[comp][1] Ey + RL[comp][2] Ez), Ez=(Ex X+Ey Y)/rho
XL_DEV void xl_lens_row(double X, double Y, double R, double f, double s2, int comp, float* ax, float* ay) {
    double rho2 = X * X + Y * Y;
    double rho = sqrt(rho2);
    float th = (float)(rho / f);
    float st, ct;
    xl_sincosf(th, &st, &ct);
    float cp = rho == 0.0 ? 1.f : (float)(X / rho), sp = rho == 0.0 ? 0.f : (float)(Y / rho);
    // Ez coefficients: X/rho, Y/rho -- 0/0 = NaN at the origin exactly like the reference (optical_elements.py:538)
    float zx = (float)(X / rho), zy = (float)(Y / rho);
    float pupil = (rho2 / (R * R) < 1.0) ? 1.f : 0.f;
    double uv = (X / R) * (X / R) + (Y / R) * (Y / R);
    float G = pupil / sqrtf(fabsf((float)(1.0 - uv * s2)));
    float w = sqrtf(fabsf(ct)) * G;
    float r0, r1, r2;
    if (comp == 0) { r0 = ct * cp * cp + sp * sp; r1 = ct * cp * sp - sp * cp; r2 = -cp * st; }
    else if (comp == 1) { r0 = sp * ct * cp - cp * sp; r1 = ct * sp * sp + cp * cp; r2 = -sp * st; }
    else { r0 = st * cp; r1 = st * sp; r2 = ct; }
    *ax = w * (r0 + r2 * zx);
    *ay = w * (r1 + r2 * zy);
}

template <int L> struct XlCztOp {
    const XlCztParams& p; int lb, comp; double z; cf cst;
    XL_DEV void coords(const XlGridFactor& g, int line, int pos, double* X, double* Y) const {
        if (g.swap) { *X = g.x0 + pos * g.dx; *Y = g.y0 + line * g.dy; }
        else { *X = g.x0 + line * g.dx; *Y = g.y0 + pos * g.dy; }
    }
    XL_DEV cf load(int c, int i) const {
        const int line = lb + c;
        if (line >= p.nlines || i >= p.m_in) return make_float2(0.f, 0.f);
        const long long o = (long long)line * p.in_line + (long long)i * p.in_pos;
        cf v;
        if (p.pro == XL_PRO_NONE) {
            v = p.in[(long long)comp * p.in_comp + o];
            if (p.flags & XL_F_CONJ_IN) v.y = -v.y;
        } else {
            double X, Y;
            coords(p.gpro, line, i, &X, &Y);
            if (p.pro == XL_PRO_RSF) {
                v = p.in[(long long)comp * p.in_comp + o];
                if (p.flags & XL_F_CONJ_IN) v.y = -v.y;
                v = cf_mul(v, xl_rs_h(X, Y, z, p.k, 0));
            } else if (p.pro == XL_PRO_VCZT) {
                if (comp < 2) {
                    v = p.in[(long long)comp * p.in_comp + o];
                } else {  // Ez = ((Ex X + Ey Y)/r) * z/r     vectorized_optics.py:341-344
                    cf ex = p.in[o], ey = p.in[p.in_comp + o];
                    double ir2 = 1.0 / (X * X + Y * Y + z * z);
                    float ax = (float)(X * z * ir2), ay = (float)(Y * z * ir2);
                    v = make_float2(ex.x * ax + ey.x * ay, ex.y * ax + ey.y * ay);
                }
                v = cf_mul(v, xl_rs_h(X, Y, z, p.k, 0));
            } else {  // XL_PRO_HIGHNA
                cf ex = p.in[o], ey = p.in[p.in_comp + o];
                float ax, ay;
                xl_lens_row(X, Y, p.lens_R, p.lens_f, p.lens_s2, comp, &ax, &ay);
                v = make_float2(ex.x * ax + ey.x * ay, ex.y * ax + ey.y * ay);
            }
        }
        return cf_mul(v, xl_ldg(p.pre + i));
    }
    XL_DEV void spec(int c, int beta, cf* v) const {
        (void)c;
#pragma unroll
        for (int q = 0; q < 16; ++q) v[q] = cf_mul(v[q], xl_ldg(p.ft + q * (L / 16) + beta));
    }
    XL_DEV void store(int c, int i, cf val) const {
        const int line = lb + c, o = i - p.out_off;
        if (line >= p.nlines || o < 0 || o >= p.m_out) return;
        val = cf_mul(val, xl_ldg(p.post + o));
        if (p.epi == XL_EPI_RSF) {
            double X, Y;
            coords(p.gepi, line, o, &X, &Y);
            val = cf_mul(val, xl_rs_h(X, Y, z, p.k, 0));
        }
        val = cf_mul(val, cst);
        if (p.flags & XL_F_CONJ_OUT) val.y = -val.y;
        p.out[(long long)comp * p.out_comp + (long long)line * p.out_line + (long long)o * p.out_pos] = val;
    }
};
template <int L> struct XlCztAxis {
    static const char* name() { return "czt_axis"; }
    typedef XlCztParams Params;
    static constexpr int NT = xl_threads(L, XL_CW);
    XL_DEV static void run(const Params& p, cf* s) {
        const double z = p.z ? xl_ldg(p.z) : 0.0;
        double cr = p.epi_cr, ci = p.epi_ci;
        if (p.epi_times_z) { cr *= z; ci *= z; }
        XlCztOp<L> op{p, XL_BLOCK_X * XL_CW, XL_BLOCK_Y, z, make_float2((float)cr, (float)ci)};
        XlFft<L, XL_CW, NT>::conv(s, p.tw, op);
    }
};

// Bluestein tables for one axis (wave_optics.py:385-410, 430-459), all phases in fp64:
//   pre[k]  = A^-k * h_k                       h_j = W^(j^2/2) on the principal branch of log W
//   post[l] = h_l * exp(-i 2 pi f_l (1/2 - m/2)/Dm)
//   ft      = FFT_L(1/h[:mp+1]) / L            (slot order)         -- forward pass multiplier
//   ftT     = FFT_L(transposed kernel) / L     (slot order)         -- adjoint pass multiplier
struct XlCztSetupParams {
    int L, m, M;
    double out0, outl;        // first / last output coordinate of this axis
    double Dm_static;         // used when z == null
    const double* z; double lambda_over_dx;   // Dm = lambda*z/dx   (wave_optics.py:322)
    cf* pre; cf* post; cf* ft; cf* ftT;
    const cf* tw;
};
struct XlCztAxisConsts { double Dm, D1, D2, uw; int Lh; };
XL_DEV XlCztAxisConsts xl_czt_consts(const XlCztSetupParams& p) {
    XlCztAxisConsts a;
    a.Dm = p.z ? p.lambda_over_dx * xl_ldg(p.z) : p.Dm_static;
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
template <int L> struct XlCztSetupOp {
    const XlCztSetupParams& p; XlCztAxisConsts a;
    XL_DEV cf load(int c, int i) const {
        int t;
        if (c == 0) t = i;
        else if (c == 1) {
            int sft;
            if (i <= p.m - 1) sft = i; else if (i >= L - (p.M - 1)) sft = i - L; else return make_float2(0.f, 0.f);
            t = p.m - sft;
        } else return make_float2(0.f, 0.f);
        if (t < 0 || t >= a.Lh) return make_float2(0.f, 0.f);
        return xl_chirp(a.uw, (double)(t - (p.m - 1)), -1.0);  // 1/h_j = conj(h_j)
    }
    XL_DEV void spec(int c, int beta, cf* v) const {
        cf* dst = c == 0 ? p.ft : (c == 1 ? p.ftT : (cf*)0);
        if (!dst) return;
#pragma unroll
        for (int q = 0; q < 16; ++q) dst[q * (L / 16) + beta] = cf_scale(v[q], 1.0f / L);
    }
    XL_DEV void store(int, int, cf) const {}
};
template <int L> struct XlCztSetup {
    static const char* name() { return "czt_setup"; }
    typedef XlCztSetupParams Params;
    static constexpr int NT = xl_threads(L, XL_CW);
    XL_DEV static void run(const Params& p, cf* s) {
        XlCztAxisConsts a = xl_czt_consts(p);
        XL_THREADS(tid, NT) {
            for (int k = tid; k < p.m; k += NT) {  // pre[k]
                double cyc = a.D1 / a.Dm;
                cyc -= rint(cyc);
                double ph = -(double)k * cyc;
                ph -= rint(ph);
                double sn, cs;
                xl_sincospi(2.0 * ph, &sn, &cs);
                cf h = xl_chirp(a.uw, (double)k, 1.0);
                p.pre[k] = cf_mul(make_float2((float)cs, (float)sn), h);
            }
            for (int l = tid; l < p.M; l += NT) {  // post[l]
                double fl = (double)l / p.M * (a.D2 - a.D1) + a.D1;              // :451-453
                double ph = -fl * (-(double)p.m / 2 + 0.5) / a.Dm;                // :456-457
                ph -= rint(ph);
                double sn, cs;
                xl_sincospi(2.0 * ph, &sn, &cs);
                cf h = xl_chirp(a.uw, (double)l, 1.0);
                p.post[l] = cf_mul(make_float2((float)cs, (float)sn), h);
            }
        }
        XlCztSetupOp<L> op{p, a};
        XlFft<L, XL_CW, NT>::forward(s, p.tw, op);
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
    double lens_R, lens_f, lens_s2;
};
struct XlFold {
    static const char* name() { return "fold"; }
    typedef XlFoldParams Params;
    static constexpr int NT = 256;
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
                    double ir = 1.0 / sqrt(X * X + Y * Y + z * z);
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
                    float a0x, a0y, a1x, a1y, a2x, a2y;
                    xl_lens_row(X, Y, p.lens_R, p.lens_f, p.lens_s2, 0, &a0x, &a0y);
                    xl_lens_row(X, Y, p.lens_R, p.lens_f, p.lens_s2, 1, &a1x, &a1y);
                    xl_lens_row(X, Y, p.lens_R, p.lens_f, p.lens_s2, 2, &a2x, &a2y);
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
        if (p.gz) {
            XL_THREADS(tid, NT) {
                if (tid == 0) {
                    double a = 0.0;
                    for (int i = 0; i < NT; ++i) a += (double)red[i];
                    xl_atomic_add(p.gz, a);
                }
            }
        }
    }
};
