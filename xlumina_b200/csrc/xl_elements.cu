// xl_elements.cu -- the pointwise Jones elements that sit between the propagations of a vectorial optical table
// (SURVEY.md 8f-1): sSLM, LCD and BS_symmetric, each as ONE pass over the planes it changes, forward and VJP, with the
// reference's parameter maps and the reductions of the scalar-parameter gradients inside the kernels.
//
// The reference builds an (N^2, 2, 2) Jones tensor per element and batch-multiplies it (optical_elements.py:142-180,
// 208-213, 283-290, 375-391); as host-level torch arithmetic one element of the sharp-focus table is ~15 pointwise /
// reduction launches per direction (round 2 profile: the elements and their scalar-parameter algebra were 48 % of the GPU
// time of one loss+gradient of BASELINE config 3).  Here: one launch forward, one backward (+ a one-thread finish for the
// scalar chain rule).
//
// Conventions: planes are complex64 [n] (n = N*N pixels, any layout: the elements are pointwise); cotangents follow torch
// (g = dL/dRe + i dL/dIm of a real loss L), so  g_in = conj(J)^T g_out  and, for a real parameter t,
// dL/dt = Re sum conj(g_out) * d out/dt.  Scalar parameters are float64 in device memory (they are optimizer parameters).
#include "xl_common.h"
#include "xl_kernels.cuh"

#define XL_EL_NT 256
#define XL_EL_PER 4   // pixels per thread

static inline int el_grid(size_t n) { return (int)((n + (size_t)XL_EL_NT * XL_EL_PER - 1) / ((size_t)XL_EL_NT * XL_EL_PER)); }

// exp(i * (scale * p + offset)) with the phase formed in fp64 and rounded to fp32 once (as the torch host layer did)
XL_DEV cf xl_el_phasor(float p, double scale, double offset) {
    const float ph = (float)((double)p * scale + offset);
    float s, c;
    xl_sincosf(ph, &s, &c);
    return make_float2(c, s);
}

// ------------------------------------------------------------------------------------------------ sSLM
// out_x = ex * exp(i (scale alpha + offset)),  out_y = ey * exp(i (scale phi + offset)).      optical_elements.py:186-222
// VJP:  g_ex = g_ox conj(m_a),  g_alpha = scale Im(g_ox conj(out_x))   (d out_x / d alpha = i scale out_x).
struct XlElSslmParams {
    const cf* ex; const cf* ey; const float* alpha; const float* phi; double scale, offset;
    cf* ox; cf* oy;                                   // forward
    const cf* gox; const cf* goy; cf* gex; cf* gey; float* galpha; float* gphi;   // backward (each may be null)
    size_t n; int backward;
};
struct XlElSslm {
    static const char* name() { return "el_sslm"; }
    typedef XlElSslmParams Params;
    static constexpr int NT = XL_EL_NT;
    static size_t smem() { return 16; }
    XL_DEV static void one(const Params& p, const cf* e, const float* ph, cf* o, const cf* go, cf* ge, float* gph, size_t i) {
        const cf m = xl_el_phasor(ph[i], p.scale, p.offset);
        if (!p.backward) { o[i] = cf_mul(e[i], m); return; }
        const cf g = go ? go[i] : cf_zero();
        if (ge) ge[i] = cf_mulc(g, m);
        if (gph) {
            const cf out = cf_mul(e[i], m);
            gph[i] = (float)p.scale * (g.y * out.x - g.x * out.y);   // Im(g conj(out))
        }
    }
    XL_DEV static void run(const Params& p, cf*) {
        XL_THREADS(tid, NT) {
#pragma unroll
            for (int e = 0; e < XL_EL_PER; ++e) {
                const size_t i = ((size_t)XL_BLOCK_X * XL_EL_PER + e) * NT + tid;
                if (i >= p.n) continue;
                one(p, p.ex, p.alpha, p.ox, p.gox, p.gex, p.galpha, i);
                one(p, p.ey, p.phi, p.oy, p.goy, p.gey, p.gphi, i);
            }
        }
    }
};

// ------------------------------------------------------------------------------------------------ scalar-parameter sums
// Every scalar-parameter gradient of LCD / BS is a fixed real-linear combination of a few complex sums over the pixels.
// The pointwise backward kernels accumulate those sums in fp64 (per-thread -> shared-memory tree -> one atomic per CTA and
// sum) into `sums`, and a one-thread finish kernel applies the chain rule of the parameter map.
#define XL_EL_NSUM 8
XL_DEV void xl_el_reduce(double* red, const double* acc, double* sums, int nsum) {
    for (int s = 0; s < nsum; ++s) {
        XL_THREADS(tid, XL_EL_NT) { red[tid] = XL_PER_THREAD(tid, acc[s]); }
        XL_SYNC();
        xl_block_sum<XL_EL_NT>(red);
        XL_THREADS(tid, XL_EL_NT) {
            if (tid == 0) xl_atomic_add(sums + s, red[0]);
        }
        XL_SYNC();
    }
}

// ------------------------------------------------------------------------------------------------ LCD
// Uniform retarder (eta) with its fast axis at theta, delta = 0:  [[a, b], [b, d]],
//   a = cos(eta/2) - i sin(eta/2) cos 2theta,  b = -i sin(eta/2) sin 2theta,  d = cos(eta/2) + i sin(eta/2) cos 2theta
// (optical_elements.py:123-140, 170-180, 266-305).  The parameters arrive as the optimizer's raw values p: eta = scale p + offset.
struct XlElLcdParams {
    const cf* ex; const cf* ey; const double* eta; const double* theta; double scale, offset;
    cf* ox; cf* oy;
    const cf* gox; const cf* goy; cf* gex; cf* gey;
    double* sums;            // [8]: Re/Im of Pxx, Pxy + Pyx, Pyy  (P_uv = sum conj(g_ou) e_v); [6], [7] unused
    double* geta; double* gtheta;   // += dL/d(raw parameter)
    size_t n; int backward;
};
struct XlLcdJones { cf a, b, d; double ch, sh, c2, s2; };
XL_DEV XlLcdJones xl_lcd_jones(const XlElLcdParams& p) {
    const double eta = xl_ldg(p.eta) * p.scale + p.offset, th = xl_ldg(p.theta) * p.scale + p.offset;
    XlLcdJones j;
    j.ch = cos(0.5 * eta); j.sh = sin(0.5 * eta); j.c2 = cos(2.0 * th); j.s2 = sin(2.0 * th);
    j.a = make_float2((float)j.ch, (float)(-j.sh * j.c2));
    j.b = make_float2(0.f, (float)(-j.sh * j.s2));
    j.d = make_float2((float)j.ch, (float)(j.sh * j.c2));
    return j;
}
struct XlElLcd {
    static const char* name() { return "el_lcd"; }
    typedef XlElLcdParams Params;
    static constexpr int NT = XL_EL_NT;
    static size_t smem() { return NT * sizeof(double); }
    XL_DEV static void run(const Params& p, cf* s) {
        const XlLcdJones j = xl_lcd_jones(p);
        double acc[6] = {0, 0, 0, 0, 0, 0};
        XL_THREADS(tid, NT) {
#pragma unroll
            for (int e = 0; e < XL_EL_PER; ++e) {
                const size_t i = ((size_t)XL_BLOCK_X * XL_EL_PER + e) * NT + tid;
                if (i >= p.n) continue;
                const cf ex = p.ex[i], ey = p.ey[i];
                if (!p.backward) {
                    p.ox[i] = cf_fma(j.b, ey, cf_mul(j.a, ex));
                    p.oy[i] = cf_fma(j.d, ey, cf_mul(j.b, ex));
                    continue;
                }
                const cf gx = p.gox ? p.gox[i] : cf_zero(), gy = p.goy ? p.goy[i] : cf_zero();
                if (p.gex) {   // conj(J)^T g  (J is symmetric)
                    p.gex[i] = cf_fma(gy, cf_conj(j.b), cf_mulc(gx, j.a));
                    p.gey[i] = cf_fma(gy, cf_conj(j.d), cf_mulc(gx, j.b));
                }
                // P_uv = conj(g_u) e_v
                const double gxr = gx.x, gxi = gx.y, gyr = gy.x, gyi = gy.y, xr = ex.x, xi = ex.y, yr = ey.x, yi = ey.y;
                acc[0] += gxr * xr + gxi * xi;              acc[1] += gxr * xi - gxi * xr;                 // Pxx
                acc[2] += gxr * yr + gxi * yi + gyr * xr + gyi * xi;
                acc[3] += gxr * yi - gxi * yr + gyr * xi - gyi * xr;                                       // Pxy + Pyx
                acc[4] += gyr * yr + gyi * yi;              acc[5] += gyr * yi - gyi * yr;                 // Pyy
            }
        }
        if (p.backward && p.sums) xl_el_reduce((double*)s, acc, p.sums, 6);   // kernel-uniform
    }
};
struct XlElLcdFinish {
    static const char* name() { return "el_lcd_finish"; }
    typedef XlElLcdParams Params;
    static constexpr int NT = 32;
    static size_t smem() { return 16; }
    XL_DEV static void run(const Params& p, cf*) {
        const XlLcdJones j = xl_lcd_jones(p);
        XL_THREADS(tid, NT) {
            if (tid != 0) continue;
            const double* S = p.sums;
            // dL/dt = Re(a_t Pxx + b_t (Pxy + Pyx) + d_t Pyy);  Re((u + i v)(Pr + i Pi)) = u Pr - v Pi
            // d/d eta:   a = (-sh/2) - i (ch/2) c2,  b = -i (ch/2) s2,  d = (-sh/2) + i (ch/2) c2
            const double ge = (-0.5 * j.sh) * S[0] + (0.5 * j.ch * j.c2) * S[1] + (0.5 * j.ch * j.s2) * S[3]
                            + (-0.5 * j.sh) * S[4] - (0.5 * j.ch * j.c2) * S[5];
            // d/d theta: a = i 2 sh s2,  b = -i 2 sh c2,  d = -i 2 sh s2
            const double gt = -(2.0 * j.sh * j.s2) * S[1] + (2.0 * j.sh * j.c2) * S[3] + (2.0 * j.sh * j.s2) * S[5];
            if (p.geta) xl_atomic_add(p.geta, ge * p.scale);
            if (p.gtheta) xl_atomic_add(p.gtheta, gt * p.scale);
        }
    }
};

// ------------------------------------------------------------------------------------------------ BS_symmetric
// c = R a + i T b,  d = i T a + R b,  T = 0.99 |cos theta|,  R = |sin theta| - 0.01 |cos theta|   (optical_elements.py:334-392);
// theta = scale p + offset.  VJP: g_a = R g_c - i T g_d,  g_b = -i T g_c + R g_d;
// dL/dR = Re sum (conj(g_c) a + conj(g_d) b),  dL/dT = Re sum i (conj(g_c) b + conj(g_d) a) = -Im sum (conj(g_c) b + conj(g_d) a).
struct XlElBsParams {
    const cf* a[2]; const cf* b[2]; const double* theta; double scale, offset;   // [0] = Ex plane, [1] = Ey plane
    cf* c[2]; cf* d[2];
    const cf* gc[2]; const cf* gd[2]; cf* ga[2]; cf* gb[2];                      // each pair may be null
    double* sums;            // [2]: dL/dR, dL/dT
    double* gtheta;          // += dL/d(raw parameter)
    size_t n; int backward;
};
struct XlBsRT { float R, T; double dR, dT; };
XL_DEV XlBsRT xl_bs_rt(const XlElBsParams& p) {
    const double th = xl_ldg(p.theta) * p.scale + p.offset;
    const double c = cos(th), s = sin(th), ac = fabs(c), as = fabs(s);
    XlBsRT r;
    r.T = (float)(0.99 * ac);
    r.R = (float)(as - 0.01 * ac);
    const double sc = c > 0 ? 1.0 : (c < 0 ? -1.0 : 0.0), ss = s > 0 ? 1.0 : (s < 0 ? -1.0 : 0.0);
    r.dT = -0.99 * sc * s;                 // d|cos|/dth = -sgn(cos) sin
    r.dR = ss * c + 0.01 * sc * s;
    return r;
}
struct XlElBs {
    static const char* name() { return "el_bs"; }
    typedef XlElBsParams Params;
    static constexpr int NT = XL_EL_NT;
    static size_t smem() { return NT * sizeof(double); }
    XL_DEV static void run(const Params& p, cf* s) {
        const XlBsRT rt = xl_bs_rt(p);
        double acc[2] = {0, 0};
        XL_THREADS(tid, NT) {
#pragma unroll
            for (int e = 0; e < XL_EL_PER; ++e) {
                const size_t i = ((size_t)XL_BLOCK_X * XL_EL_PER + e) * NT + tid;
                if (i >= p.n) continue;
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    const cf a = p.a[k][i], b = p.b[k][i];
                    if (!p.backward) {
                        p.c[k][i] = cf_lin2(a, rt.R, cf_muli(b), rt.T);
                        p.d[k][i] = cf_lin2(cf_muli(a), rt.T, b, rt.R);
                        continue;
                    }
                    const cf gc = p.gc[k] ? p.gc[k][i] : cf_zero(), gd = p.gd[k] ? p.gd[k][i] : cf_zero();
                    if (p.ga[k]) p.ga[k][i] = cf_lin2(gc, rt.R, cf_mulni(gd), rt.T);
                    if (p.gb[k]) p.gb[k][i] = cf_lin2(cf_mulni(gc), rt.T, gd, rt.R);
                    const double cr = gc.x, ci = gc.y, dr = gd.x, di = gd.y, ar = a.x, ai = a.y, br = b.x, bi = b.y;
                    acc[0] += cr * ar + ci * ai + dr * br + di * bi;                   // Re(conj(gc) a + conj(gd) b)
                    acc[1] -= cr * bi - ci * br + dr * ai - di * ar;                   // -Im(conj(gc) b + conj(gd) a)
                }
            }
        }
        if (p.backward && p.sums) xl_el_reduce((double*)s, acc, p.sums, 2);   // kernel-uniform
    }
};
struct XlElBsFinish {
    static const char* name() { return "el_bs_finish"; }
    typedef XlElBsParams Params;
    static constexpr int NT = 32;
    static size_t smem() { return 16; }
    XL_DEV static void run(const Params& p, cf*) {
        const XlBsRT rt = xl_bs_rt(p);
        XL_THREADS(tid, NT) {
            if (tid != 0) continue;
            if (p.gtheta) xl_atomic_add(p.gtheta, (p.sums[0] * rt.dR + p.sums[1] * rt.dT) * p.scale);
        }
    }
};

// ================================================================================================ C ABI
extern "C" size_t xl_el_scratch_bytes(void) { return XL_EL_NSUM * sizeof(double); }

extern "C" int xl_el_sslm(const void* ex, const void* ey, const float* alpha, const float* phi, double scale, double offset,
                          void* ox, void* oy, size_t n, void* stream) {
    if (!ex || !ey || !alpha || !phi || !ox || !oy || !n) return xl_fail(XL_E_BAD_ARG, "xl_el_sslm: null pointer%s", "");
    XlElSslmParams p;
    memset(&p, 0, sizeof(p));
    p.ex = (const cf*)ex; p.ey = (const cf*)ey; p.alpha = alpha; p.phi = phi; p.scale = scale; p.offset = offset;
    p.ox = (cf*)ox; p.oy = (cf*)oy; p.n = n;
    return xl_launch<XlElSslm>(XlDim{el_grid(n), 1}, (xl_stream_t)stream, p);
}
extern "C" int xl_el_sslm_bwd(const void* ex, const void* ey, const float* alpha, const float* phi, double scale, double offset,
                              const void* g_ox, const void* g_oy, void* g_ex, void* g_ey, float* g_alpha, float* g_phi,
                              size_t n, void* stream) {
    if (!ex || !ey || !alpha || !phi || !n) return xl_fail(XL_E_BAD_ARG, "xl_el_sslm_bwd: null pointer%s", "");
    XlElSslmParams p;
    memset(&p, 0, sizeof(p));
    p.ex = (const cf*)ex; p.ey = (const cf*)ey; p.alpha = alpha; p.phi = phi; p.scale = scale; p.offset = offset;
    p.gox = (const cf*)g_ox; p.goy = (const cf*)g_oy; p.gex = (cf*)g_ex; p.gey = (cf*)g_ey; p.galpha = g_alpha; p.gphi = g_phi;
    p.n = n; p.backward = 1;
    return xl_launch<XlElSslm>(XlDim{el_grid(n), 1}, (xl_stream_t)stream, p);
}

extern "C" int xl_el_lcd(const void* ex, const void* ey, const double* eta, const double* theta, double scale, double offset,
                         void* ox, void* oy, size_t n, void* stream) {
    if (!ex || !ey || !eta || !theta || !ox || !oy || !n) return xl_fail(XL_E_BAD_ARG, "xl_el_lcd: null pointer%s", "");
    XlElLcdParams p;
    memset(&p, 0, sizeof(p));
    p.ex = (const cf*)ex; p.ey = (const cf*)ey; p.eta = eta; p.theta = theta; p.scale = scale; p.offset = offset;
    p.ox = (cf*)ox; p.oy = (cf*)oy; p.n = n;
    return xl_launch<XlElLcd>(XlDim{el_grid(n), 1}, (xl_stream_t)stream, p);
}
extern "C" int xl_el_lcd_bwd(const void* ex, const void* ey, const double* eta, const double* theta, double scale, double offset,
                             const void* g_ox, const void* g_oy, void* g_ex, void* g_ey, double* g_eta, double* g_theta,
                             void* scratch, size_t n, void* stream) {
    if (!ex || !ey || !eta || !theta || !n) return xl_fail(XL_E_BAD_ARG, "xl_el_lcd_bwd: null pointer%s", "");
    if ((g_eta || g_theta) && !scratch) return xl_fail(XL_E_WORKSPACE, "xl_el_lcd_bwd: parameter gradients need the scratch buffer%s", "");
    if ((g_ex == 0) != (g_ey == 0)) return xl_fail(XL_E_BAD_ARG, "xl_el_lcd_bwd: g_ex and g_ey come together%s", "");
    xl_stream_t st = (xl_stream_t)stream;
    XlElLcdParams p;
    memset(&p, 0, sizeof(p));
    p.ex = (const cf*)ex; p.ey = (const cf*)ey; p.eta = eta; p.theta = theta; p.scale = scale; p.offset = offset;
    p.gox = (const cf*)g_ox; p.goy = (const cf*)g_oy; p.gex = (cf*)g_ex; p.gey = (cf*)g_ey;
    p.geta = g_eta; p.gtheta = g_theta; p.n = n; p.backward = 1;
    int rc;
    if (g_eta || g_theta) {
        p.sums = (double*)scratch;
        rc = zero_async(scratch, XL_EL_NSUM * sizeof(double), st);
        if (rc) return rc;
    }
    rc = xl_launch<XlElLcd>(XlDim{el_grid(n), 1}, st, p);
    if (rc || !p.sums) return rc;
    return xl_launch<XlElLcdFinish>(XlDim{1, 1}, st, p);
}

static void bs_planes(XlElBsParams& p, const void* a_ex, const void* a_ey, const void* b_ex, const void* b_ey) {
    p.a[0] = (const cf*)a_ex; p.a[1] = (const cf*)a_ey; p.b[0] = (const cf*)b_ex; p.b[1] = (const cf*)b_ey;
}
extern "C" int xl_el_bs(const void* a_ex, const void* a_ey, const void* b_ex, const void* b_ey, const double* theta, double scale, double offset,
                        void* c_ex, void* c_ey, void* d_ex, void* d_ey, size_t n, void* stream) {
    if (!a_ex || !a_ey || !b_ex || !b_ey || !theta || !c_ex || !c_ey || !d_ex || !d_ey || !n) return xl_fail(XL_E_BAD_ARG, "xl_el_bs: null pointer%s", "");
    XlElBsParams p;
    memset(&p, 0, sizeof(p));
    bs_planes(p, a_ex, a_ey, b_ex, b_ey);
    p.theta = theta; p.scale = scale; p.offset = offset;
    p.c[0] = (cf*)c_ex; p.c[1] = (cf*)c_ey; p.d[0] = (cf*)d_ex; p.d[1] = (cf*)d_ey; p.n = n;
    return xl_launch<XlElBs>(XlDim{el_grid(n), 1}, (xl_stream_t)stream, p);
}
extern "C" int xl_el_bs_bwd(const void* a_ex, const void* a_ey, const void* b_ex, const void* b_ey, const double* theta, double scale, double offset,
                            const void* g_c_ex, const void* g_c_ey, const void* g_d_ex, const void* g_d_ey,
                            void* g_a_ex, void* g_a_ey, void* g_b_ex, void* g_b_ey, double* g_theta,
                            void* scratch, size_t n, void* stream) {
    if (!a_ex || !a_ey || !b_ex || !b_ey || !theta || !n) return xl_fail(XL_E_BAD_ARG, "xl_el_bs_bwd: null pointer%s", "");
    if (g_theta && !scratch) return xl_fail(XL_E_WORKSPACE, "xl_el_bs_bwd: the parameter gradient needs the scratch buffer%s", "");
    if ((g_c_ex == 0) != (g_c_ey == 0) || (g_d_ex == 0) != (g_d_ey == 0) || (g_a_ex == 0) != (g_a_ey == 0) || (g_b_ex == 0) != (g_b_ey == 0))
        return xl_fail(XL_E_BAD_ARG, "xl_el_bs_bwd: the Ex and Ey planes of a beam come together%s", "");
    xl_stream_t st = (xl_stream_t)stream;
    XlElBsParams p;
    memset(&p, 0, sizeof(p));
    bs_planes(p, a_ex, a_ey, b_ex, b_ey);
    p.theta = theta; p.scale = scale; p.offset = offset;
    p.gc[0] = (const cf*)g_c_ex; p.gc[1] = (const cf*)g_c_ey; p.gd[0] = (const cf*)g_d_ex; p.gd[1] = (const cf*)g_d_ey;
    p.ga[0] = (cf*)g_a_ex; p.ga[1] = (cf*)g_a_ey; p.gb[0] = (cf*)g_b_ex; p.gb[1] = (cf*)g_b_ey;
    p.gtheta = g_theta; p.n = n; p.backward = 1;
    int rc;
    if (g_theta) {
        p.sums = (double*)scratch;
        rc = zero_async(scratch, XL_EL_NSUM * sizeof(double), st);
        if (rc) return rc;
    }
    rc = xl_launch<XlElBs>(XlDim{el_grid(n), 1}, st, p);
    if (rc || !p.sums) return rc;
    return xl_launch<XlElBsFinish>(XlDim{1, 1}, st, p);
}
