// xl_czt.cu -- C ABI of the CZT / VCZT / high-NA path (include/xlprop.h) over the Bluestein kernels in xl_kernels.cuh.
#include "xl_host.h"

// ================================================================================================ CZT family
// Caller-owned table buffer (xl_czt_tables_bytes / xl_highna_tables_bytes): Bluestein chirps and kernel spectra of both axes
// and the pointwise factor tables.  A forward call fills it (unless XL_REUSE_TABLES); the backward call of the same
// propagation, and any later call with the same sizes, grids and z, reuses it.
struct CztPlan {
    int N, Mx, My, Ly, Lx, ncomp, mode;
    cf *pre_y, *post_y, *ft_y, *ftT_y, *kin_y, *pre_x, *post_x, *ft_x, *ftT_x, *kin_x;
    cf* T[3]; int Qx[3], Qy[3], tr[3]; XlFacAxis fx[3], fy[3]; int fac_kind[3];   // [0] input grid, transposed planes (the axis
                                                                                  // passes whose lines run along y); [1] output
                                                                                  // grid; [2] input grid row-major (high-NA fold)
    cf* mid;   // [ncomp][N][My]
    cf* tmp3;  // [3][N][N]
};
struct CztCall {
    int N, Mx, My, mode;  // mode: 0 scalar CZT, 1 VCZT, 2 high-NA
    const double* z; double lambda, k;
    double x0, dx, y0, dy, xout0, xoutl, yout0, youtl;
    double R, f, s2;
    int flags;
    long long ey_off;     // vectorial inputs: elements from the Ex plane to the Ey plane (N*N when stacked)
    int batch;            // scalar CZT only: this many contiguous planes in ONE set of launches (0 or 1: a single plane)
};
static int axis_sym(double x0, double dx, int n) {   // grid symmetric about 0 (toolbox.space): one quadrant of a factor is enough
    const double span = fabs((n - 1) * dx);
    return fabs(x0 + 0.5 * (n - 1) * dx) <= 1e-9 * (span > 0 ? span : 1.0) ? 1 : 0;
}
static size_t czt_ws_bytes(int N, int Mx, int My, int ncomp) {
    int Ly = xl_czt_padded_length(N, My), Lx = xl_czt_padded_length(N, Mx);
    if (!Ly || !Lx) return 0;
    return align_up((size_t)ncomp * N * My * sizeof(cf)) + align_up((size_t)3 * N * N * sizeof(cf));
}
// upper bound that does not depend on the grids' symmetry: full-plane factor tables
static size_t czt_tab_bytes(int N, int Mx, int My, int mode) {
    int Ly = xl_czt_padded_length(N, My), Lx = xl_czt_padded_length(N, Mx);
    if (!Ly || !Lx) return 0;
    size_t t = 2 * align_up((size_t)N * sizeof(cf)) + align_up((size_t)My * sizeof(cf)) + align_up((size_t)Mx * sizeof(cf));
    t += 4 * align_up((size_t)Ly * sizeof(cf)) + 4 * align_up((size_t)Lx * sizeof(cf));
    if (mode == 2) t += 2 * align_up((size_t)3 * N * N * sizeof(cf));
    else t += align_up((size_t)N * N * sizeof(cf)) + align_up((size_t)Mx * My * sizeof(cf));
    return t;
}
extern "C" size_t xl_czt_workspace_bytes(int N, int Mx, int My, int vectorial) { return czt_ws_bytes(N, Mx, My, vectorial ? 3 : 1); }
extern "C" size_t xl_highna_workspace_bytes(int N, int Mx, int My) { return czt_ws_bytes(N, Mx, My, 3); }
extern "C" size_t xl_czt_tables_bytes(int N, int Mx, int My) { return czt_tab_bytes(N, Mx, My, 0); }
extern "C" size_t xl_highna_tables_bytes(int N, int Mx, int My) { return czt_tab_bytes(N, Mx, My, 2); }

static int czt_plan(CztPlan& pl, const CztCall& cc, void* tables, void* ws, size_t ws_bytes) {
    const int N = cc.N, Mx = cc.Mx, My = cc.My;
    if (N < 2 || Mx < 2 || My < 2) return xl_fail(XL_E_BAD_ARG, "czt: sizes must be >= 2%s", "");
    pl.N = N; pl.Mx = Mx; pl.My = My; pl.mode = cc.mode; pl.ncomp = cc.mode == 0 ? (cc.batch > 1 ? cc.batch : 1) : 3;
    pl.Ly = xl_czt_padded_length(N, My);
    pl.Lx = xl_czt_padded_length(N, Mx);
    if (!pl.Ly || !pl.Lx) return xl_fail(XL_E_UNSUPPORTED, "czt: m+M-1 is a power of two or padded length outside [32,4096]%s", "");
    if (!tables) return xl_fail(XL_E_BAD_ARG, "czt: null table buffer%s", "");
    if (!ws || ws_bytes < czt_ws_bytes(N, Mx, My, pl.ncomp)) return xl_fail(XL_E_WORKSPACE, "czt: workspace too small%s", "");
    Carver t{(char*)tables, 0, czt_tab_bytes(N, Mx, My, cc.mode)};
    pl.pre_y = (cf*)t.take((size_t)N * sizeof(cf));
    pl.pre_x = (cf*)t.take((size_t)N * sizeof(cf));
    pl.post_y = (cf*)t.take((size_t)My * sizeof(cf));
    pl.post_x = (cf*)t.take((size_t)Mx * sizeof(cf));
    pl.ft_y = (cf*)t.take((size_t)pl.Ly * sizeof(cf));
    pl.ftT_y = (cf*)t.take((size_t)pl.Ly * sizeof(cf));
    pl.kin_y = (cf*)t.take((size_t)2 * pl.Ly * sizeof(cf));
    pl.ft_x = (cf*)t.take((size_t)pl.Lx * sizeof(cf));
    pl.ftT_x = (cf*)t.take((size_t)pl.Lx * sizeof(cf));
    pl.kin_x = (cf*)t.take((size_t)2 * pl.Lx * sizeof(cf));
    // factor tables: [0] on the input grid, [1] on the output grid (CZT / VCZT: the RS factors F, F0; high-NA: the lens matrix)
    const double dxo = (cc.xoutl - cc.xout0) / (Mx - 1), dyo = (cc.youtl - cc.yout0) / (My - 1);
    pl.fx[0] = XlFacAxis{N, axis_sym(cc.x0, cc.dx, N), cc.x0, cc.dx};
    pl.fy[0] = XlFacAxis{N, axis_sym(cc.y0, cc.dy, N), cc.y0, cc.dy};
    pl.fx[1] = XlFacAxis{Mx, axis_sym(cc.xout0, dxo, Mx), cc.xout0, dxo};
    pl.fy[1] = XlFacAxis{My, axis_sym(cc.yout0, dyo, My), cc.yout0, dyo};
    pl.fx[2] = pl.fx[0]; pl.fy[2] = pl.fy[0];
    for (int w = 0; w < 3; ++w) {
        pl.Qx[w] = xl_fac_size(pl.fx[w].n, pl.fx[w].sym);
        pl.Qy[w] = xl_fac_size(pl.fy[w].n, pl.fy[w].sym);
        pl.tr[w] = w == 0 ? 1 : 0;
        pl.fac_kind[w] = XL_FAC_NONE; pl.T[w] = 0;
    }
    if (cc.mode == 2) {
        pl.fac_kind[0] = XL_FAC_LENS; pl.fac_kind[2] = XL_FAC_LENS;
        pl.T[0] = (cf*)t.take((size_t)3 * pl.Qx[0] * pl.Qy[0] * sizeof(cf));
        pl.T[2] = (cf*)t.take((size_t)3 * pl.Qx[2] * pl.Qy[2] * sizeof(cf));
    } else {
        pl.fac_kind[0] = XL_FAC_RS;
        pl.T[0] = (cf*)t.take((size_t)pl.Qx[0] * pl.Qy[0] * sizeof(cf));
        // Identical input and output grids (the reference's default xout = x, yout = y) whose x and y axes are the same too:
        // F0 is F, and the transposed table IS the row-major one (h depends on X^2 + Y^2 only).
        const bool same = Mx == N && My == N && pl.fx[0].sym == pl.fx[1].sym && pl.fy[0].sym == pl.fy[1].sym &&
                          fabs(cc.xout0 - cc.x0) <= 1e-12 * fabs(cc.x0) && fabs(dxo - cc.dx) <= 1e-12 * fabs(cc.dx) &&
                          fabs(cc.yout0 - cc.y0) <= 1e-12 * fabs(cc.y0) && fabs(dyo - cc.dy) <= 1e-12 * fabs(cc.dy) &&
                          pl.fx[0].sym == pl.fy[0].sym && fabs(cc.x0 - cc.y0) <= 1e-12 * fabs(cc.x0) && fabs(cc.dx - cc.dy) <= 1e-12 * fabs(cc.dx);
        if (same) { pl.T[1] = pl.T[0]; }
        else { pl.fac_kind[1] = XL_FAC_RS; pl.T[1] = (cf*)t.take((size_t)pl.Qx[1] * pl.Qy[1] * sizeof(cf)); }
    }
    Carver c{(char*)ws, 0, ws_bytes};
    pl.mid = (cf*)c.take((size_t)pl.ncomp * N * My * sizeof(cf));
    pl.tmp3 = (cf*)c.take((size_t)3 * N * N * sizeof(cf));
    return XL_OK;
}
static XlFacTab fac_tab(const CztPlan& pl, int w) { return XlFacTab{pl.T[w], pl.Qx[w], pl.Qy[w], pl.fx[w], pl.fy[w], pl.tr[w]}; }

// Fill the table buffer: czt_tables (one thread per entry) then czt_kernel_fft (one CTA per axis).
static int czt_setup(const CztPlan& pl, const CztCall& cc, const cf* tw, xl_stream_t st) {
    int rc;
    XlCztTablesParams tp;
    memset(&tp, 0, sizeof(tp));
    const double Dm_static = cc.mode == 2 ? cc.f * cc.lambda * (cc.N - 1) / (2 * cc.R) : 0.0;  // optical_elements.py:663
    for (int ax = 0; ax < 2; ++ax) {
        XlCztAxisTab& s = tp.a[ax];
        s.lambda_over_dx = cc.lambda / cc.dx; s.Dm_static = Dm_static;
        s.m = pl.N;
        if (ax == 0) {   // y axis (first Bluestein pass, wave_optics.py:349)
            s.L = pl.Ly; s.M = pl.My; s.out0 = cc.yout0; s.outl = cc.youtl;
            s.pre = pl.pre_y; s.post = pl.post_y; s.kin = pl.kin_y; s.ft = pl.ft_y; s.ftT = pl.ftT_y;
        } else {         // x axis (second pass, :352)
            s.L = pl.Lx; s.M = pl.Mx; s.out0 = cc.xout0; s.outl = cc.xoutl;
            s.pre = pl.pre_x; s.post = pl.post_x; s.kin = pl.kin_x; s.ft = pl.ft_x; s.ftT = pl.ftT_x;
        }
    }
    tp.z = cc.mode == 2 ? 0 : cc.z; tp.k = cc.k;
    tp.lens_R = cc.R; tp.lens_f = cc.f; tp.lens_s2 = cc.s2;
    long long e = 0;
    for (int ax = 0; ax < 2; ++ax) {
        tp.seg[ax * 3] = e;     e += tp.a[ax].m;
        tp.seg[ax * 3 + 1] = e; e += tp.a[ax].M;
        tp.seg[ax * 3 + 2] = e; e += 2LL * tp.a[ax].L;
    }
    for (int w = 0; w < 3; ++w) {
        tp.fac_kind[w] = pl.fac_kind[w]; tp.T[w] = pl.T[w]; tp.Qx[w] = pl.Qx[w]; tp.Qy[w] = pl.Qy[w]; tp.tr[w] = pl.tr[w];
        tp.fx[w] = pl.fx[w]; tp.fy[w] = pl.fy[w];
        tp.seg[6 + w] = e;
        if (pl.fac_kind[w] != XL_FAC_NONE) e += (long long)(pl.fac_kind[w] == XL_FAC_LENS ? 3 : 1) * pl.Qx[w] * pl.Qy[w];
    }
    tp.seg[9] = e;
    // seg[] as used by the kernel: [0..2] y axis starts (pre, post, kin), [3..5] x axis, [6],[7],[8] factor tables, [9] end;
    // the kernel reads seg[ax*3+1], seg[ax*3+2] as the ENDS of pre and post relative to the axis base
    rc = xl_launch<XlCztTables>(XlDim{pointwise_grid((size_t)e, XlCztTables::NT), 1}, st, tp);
    if (rc) return rc;
    XlCztKernelFftParams kp;
    kp.a[0] = tp.a[0]; kp.a[1] = tp.a[1]; kp.tw = tw;
    if (pl.Ly == pl.Lx) {   // both axes in one launch (two CTAs)
        XL_FOR_L(pl.Ly, rc = xl_launch<XlCztKernelFft<XL>>(XlDim{2, 1}, st, kp));
        return rc;
    }
    XL_FOR_L(pl.Ly, rc = xl_launch<XlCztKernelFft<XL>>(XlDim{1, 1}, st, kp));
    if (rc) return rc;
    kp.a[0] = kp.a[1];
    XL_FOR_L(pl.Lx, rc = xl_launch<XlCztKernelFft<XL>>(XlDim{1, 1}, st, kp));
    return rc;
}

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
    XL_FOR_L(a.L, rc = xl_launch<XlCztAxis<XL, PRO, EPI, ACC>>(XlDim{grid.x * grid.y, 1}, st, a));   // component-minor order
    return rc;
}
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

static int czt_forward(const CztCall& cc, const void* in, void* out, void* tables, void* ws, size_t ws_bytes, xl_stream_t st, int in_weight = 0) {
    if (!in || !out) return xl_fail(XL_E_BAD_ARG, "czt_fwd: null pointer%s", "");
    if (cc.mode != 2 && !cc.z) return xl_fail(XL_E_BAD_ARG, "czt_fwd: null z%s", "");
    CztPlan pl;
    int rc = czt_plan(pl, cc, tables, ws, ws_bytes);
    if (rc) return rc;
    const int ncomp = pl.ncomp;
    const cf* tw = xl_twiddles();
    if (!tw) return xl_fail(XL_E_CUDA, "twiddle table allocation failed%s", "");
    if (!(cc.flags & XL_REUSE_TABLES)) { rc = czt_setup(pl, cc, tw, st); if (rc) return rc; }
    const int N = cc.N, Mx = cc.Mx, My = cc.My;
    const double dxo = (cc.xoutl - cc.xout0) / (Mx - 1), dyo = (cc.youtl - cc.yout0) / (My - 1);
    // pass 1: Bluestein along y for every input column
    XlCztParams a;
    czt_common_params(a, cc, tw);
    a.L = pl.Ly; a.nlines = N; a.ncomp = ncomp; a.m_in = N; a.out_off = 0; a.m_out = My;
    a.in = (const cf*)in; a.in_line = 1; a.in_pos = N; a.in_comp = cc.mode == 0 ? (long long)N * N : cc.ey_off;
    a.out = pl.mid; a.out_line = My; a.out_pos = 1; a.out_comp = (long long)N * My;
    a.pre = pl.pre_y; a.ft = pl.ft_y; a.post = pl.post_y;
    a.pro = cc.mode == 0 ? XL_PRO_RSF : (cc.mode == 1 ? XL_PRO_VCZT : XL_PRO_HIGHNA);
    a.gpro = XlGridFactor{cc.x0, cc.dx, cc.y0, cc.dy, 0};
    a.tpro = fac_tab(pl, 0);
    a.epi = XL_EPI_NONE;
    a.in_weight = in_weight;
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
    b.tepi = fac_tab(pl, 1);
    czt_out_const(b, cc);
    b.flags = cc.flags & XL_CONJ_OUT;
    rc = czt_axis_launch(a, XlDim{xl_groups(N), ncomp}, st);
    if (rc) return rc;
    return czt_axis_launch(b, XlDim{xl_groups(My), ncomp}, st);
}

static int czt_backward(const CztCall& cc, const void* ct_out, void* ct_in, void* tables, void* ws, size_t ws_bytes, xl_stream_t st) {
    if (!ct_out || !ct_in) return xl_fail(XL_E_BAD_ARG, "czt_bwd: null pointer%s", "");
    if (cc.mode != 2 && !cc.z) return xl_fail(XL_E_BAD_ARG, "czt_bwd: null z%s", "");
    CztPlan pl;
    int rc = czt_plan(pl, cc, tables, ws, ws_bytes);
    if (rc) return rc;
    const int ncomp = pl.ncomp;
    const cf* tw = xl_twiddles();
    if (!tw) return xl_fail(XL_E_CUDA, "twiddle table allocation failed%s", "");
    if (!(cc.flags & XL_REUSE_TABLES)) { rc = czt_setup(pl, cc, tw, st); if (rc) return rc; }
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
    b.tpro = fac_tab(pl, 1);
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
    a.tepi = fac_tab(pl, 0);
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
    if (cc.mode == 2) f.lens = fac_tab(pl, 2);
    const size_t NN = (size_t)N * N;
    return xl_launch<XlFold>(XlDim{(int)((NN + XlFold::NT - 1) / XlFold::NT), 1}, st, f);
}

// Field VJP + d/dz (XlCztDotZ, xl_kernels.cuh): the backward chain, two more forward chains on index-weighted inputs, one
// pointwise reduction.  CZT / VCZT only (the high-NA focus has no distance).
static int czt_backward_z(const CztCall& cc, const void* in, const void* out, const void* ct_out, void* ct_in, double* grad_z,
                          void* tables, void* ws, size_t ws_bytes, xl_stream_t st) {
    if (!in || !out || !grad_z) return xl_fail(XL_E_BAD_ARG, "czt_bwd_z: null pointer%s", "");
    if (cc.mode == 2) return xl_fail(XL_E_BAD_ARG, "czt_bwd_z: the high-NA focus has no propagation distance%s", "");
    const int N = cc.N, Mx = cc.Mx, My = cc.My, ncomp = cc.mode == 0 ? 1 : 3;
    const size_t base = czt_ws_bytes(N, Mx, My, ncomp), plane = align_up((size_t)ncomp * My * Mx * sizeof(cf));
    if (!base || ws_bytes < base + 2 * plane) return xl_fail(XL_E_WORKSPACE, "czt_bwd_z: workspace too small%s", "");
    int rc = czt_backward(cc, ct_out, ct_in, tables, ws, ws_bytes, st);
    if (rc) return rc;
    cf* O1 = (cf*)((char*)ws + base);
    cf* O2 = (cf*)((char*)ws + base + plane);
    CztCall cf_ = cc;
    cf_.flags = XL_REUSE_TABLES;
    rc = czt_forward(cf_, in, O1, tables, ws, ws_bytes, st, 1);   // A(k_y U): the position index of the first (y) pass
    if (rc) return rc;
    rc = czt_forward(cf_, in, O2, tables, ws, ws_bytes, st, 2);   // A(k_x U): its line index
    if (rc) return rc;
    CztPlan pl;
    rc = czt_plan(pl, cc, tables, ws, ws_bytes);
    if (rc) return rc;
    XlCztDotZParams d;
    memset(&d, 0, sizeof(d));
    d.N = N; d.Mx = Mx; d.My = My; d.ncomp = ncomp;
    d.flags = (cc.flags & XL_CONJ_IN) | (cc.mode == 0 ? (cc.flags & XL_CONJ_OUT) : 0);
    d.ct_out = (const cf*)ct_out; d.out = (const cf*)out; d.O1 = O1; d.O2 = O2;
    d.ct_in = cc.mode == 0 ? (const cf*)ct_in : pl.tmp3;
    d.in = (const cf*)in; d.ey_off = cc.ey_off; d.z = cc.z; d.k = cc.k;
    d.lambda_over_dx = cc.lambda / cc.dx; d.dDm_dz = cc.lambda / cc.dx;      // Dm = lambda z / dx, wave_optics.py:322
    d.ay = XlCztAxisDz{cc.yout0, (cc.youtl - cc.yout0) / My, N};
    d.ax = XlCztAxisDz{cc.xout0, (cc.xoutl - cc.xout0) / Mx, N};
    d.x0 = cc.x0; d.dx = cc.dx; d.y0 = cc.y0; d.dy = cc.dy;
    d.xo0 = cc.xout0; d.dxo = (cc.xoutl - cc.xout0) / (Mx - 1); d.yo0 = cc.yout0; d.dyo = (cc.youtl - cc.yout0) / (My - 1);
    d.gz = grad_z;
    const size_t n = (size_t)ncomp * ((size_t)My * Mx > (size_t)N * N ? (size_t)My * Mx : (size_t)N * N);
    return xl_launch<XlCztDotZ>(XlDim{pointwise_grid(n, XlCztDotZ::NT), 1}, st, d);
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
    c.ey_off = (long long)N * N;
    return c;
}
static int set_planes(CztCall& c, const void* ex, const void* ey) {
    if (!ey || !ex) return XL_OK;
    const long long d = (const char*)ey - (const char*)ex;
    if (d % (long long)sizeof(cf)) return xl_fail(XL_E_BAD_ARG, "czt: Ex and Ey must be 8-byte aligned%s", "");
    c.ey_off = d / (long long)sizeof(cf);
    return XL_OK;
}

extern "C" int xl_czt_fwd(const void* in, const void* ey, void* out, const double* z, double lambda, int N, int Mx, int My, int vectorial,
                          double x0, double dx, double y0, double dy, double xout0, double xoutl, double yout0, double youtl,
                          int flags, void* tables, void* ws, size_t ws_bytes, void* stream) {
    CztCall c = make_czt_call(vectorial ? 1 : 0, z, lambda, N, Mx, My, x0, dx, y0, dy, xout0, xoutl, yout0, youtl, 0, 0, flags);
    if (vectorial) { int rc = set_planes(c, in, ey); if (rc) return rc; }
    return czt_forward(c, in, out, tables, ws, ws_bytes, (xl_stream_t)stream);
}
extern "C" int xl_czt_bwd(const void* ct_out, void* ct_in, const double* z, double lambda, int N, int Mx, int My, int vectorial,
                          double x0, double dx, double y0, double dy, double xout0, double xoutl, double yout0, double youtl,
                          int flags, void* tables, void* ws, size_t ws_bytes, void* stream) {
    CztCall c = make_czt_call(vectorial ? 1 : 0, z, lambda, N, Mx, My, x0, dx, y0, dy, xout0, xoutl, yout0, youtl, 0, 0, flags);
    return czt_backward(c, ct_out, ct_in, tables, ws, ws_bytes, (xl_stream_t)stream);
}
extern "C" size_t xl_czt_workspace_bytes_z(int N, int Mx, int My, int vectorial) {
    const int ncomp = vectorial ? 3 : 1;
    const size_t base = czt_ws_bytes(N, Mx, My, ncomp);
    return base ? base + 2 * align_up((size_t)ncomp * My * Mx * sizeof(cf)) : 0;
}
extern "C" int xl_czt_bwd_z(const void* in, const void* ey, const void* out, const void* ct_out, void* ct_in, double* grad_z,
                            const double* z, double lambda, int N, int Mx, int My, int vectorial,
                            double x0, double dx, double y0, double dy, double xout0, double xoutl, double yout0, double youtl,
                            int flags, void* tables, void* ws, size_t ws_bytes, void* stream) {
    CztCall c = make_czt_call(vectorial ? 1 : 0, z, lambda, N, Mx, My, x0, dx, y0, dy, xout0, xoutl, yout0, youtl, 0, 0, flags);
    if (vectorial) { int rc = set_planes(c, in, ey); if (rc) return rc; }
    return czt_backward_z(c, in, out, ct_out, ct_in, grad_z, tables, ws, ws_bytes, (xl_stream_t)stream);
}
extern "C" int xl_highna_fwd(const void* ex, const void* ey, void* out, int N, int Mx, int My, double radius, double f, double lambda,
                             double x0, double dx, double y0, double dy, double xout0, double xoutl, double yout0, double youtl,
                             int flags, void* tables, void* ws, size_t ws_bytes, void* stream) {
    CztCall c = make_czt_call(2, 0, lambda, N, Mx, My, x0, dx, y0, dy, xout0, xoutl, yout0, youtl, radius, f, flags);
    { int rc = set_planes(c, ex, ey); if (rc) return rc; }
    return czt_forward(c, ex, out, tables, ws, ws_bytes, (xl_stream_t)stream);
}
extern "C" int xl_highna_bwd(const void* ct_out, void* ct_exy, int N, int Mx, int My, double radius, double f, double lambda,
                             double x0, double dx, double y0, double dy, double xout0, double xoutl, double yout0, double youtl,
                             int flags, void* tables, void* ws, size_t ws_bytes, void* stream) {
    CztCall c = make_czt_call(2, 0, lambda, N, Mx, My, x0, dx, y0, dy, xout0, xoutl, yout0, youtl, radius, f, flags);
    return czt_backward(c, ct_out, ct_exy, tables, ws, ws_bytes, (xl_stream_t)stream);
}


// ================================================================================================ batched entry points
// `nbatch` independent propagations in ONE call (the reference vmaps the seam functions, include/xlprop.h).  Scalar CZT with a
// shared distance runs all the planes in a single set of launches; the other shapes walk the items inside the library
// (shared distance: the tables are filled once).
extern "C" size_t xl_czt_workspace_bytes_batch(int N, int Mx, int My, int vectorial, int nbatch) {
    if (nbatch < 1) return 0;
    return czt_ws_bytes(N, Mx, My, vectorial ? 3 : nbatch);
}
extern "C" int xl_czt_fwd_batch(const void* in, const void* ey, long long in_bstride, void* out, const double* z, int z_stride, double lambda,
                                int N, int Mx, int My, int vectorial, int nbatch,
                                double x0, double dx, double y0, double dy, double xout0, double xoutl, double yout0, double youtl,
                                int flags, void* tables, void* ws, size_t ws_bytes, void* stream) {
    if (nbatch < 1 || z_stride < 0 || !z) return xl_fail(XL_E_BAD_ARG, "xl_czt_fwd_batch: bad batch shape%s", "");
    xl_stream_t st = (xl_stream_t)stream;
    if (!vectorial && z_stride == 0 && in_bstride == (long long)N * N) {
        CztCall c = make_czt_call(0, z, lambda, N, Mx, My, x0, dx, y0, dy, xout0, xoutl, yout0, youtl, 0, 0, flags);
        c.batch = nbatch;
        return czt_forward(c, in, out, tables, ws, ws_bytes, st);
    }
    const size_t tb = czt_tab_bytes(N, Mx, My, 0), oplane = (size_t)(vectorial ? 3 : 1) * My * Mx * sizeof(cf);
    for (int b = 0; b < nbatch; ++b) {
        const int f = flags | ((z_stride == 0 && b > 0) ? XL_REUSE_TABLES : 0);
        CztCall c = make_czt_call(vectorial ? 1 : 0, z + (size_t)b * z_stride, lambda, N, Mx, My, x0, dx, y0, dy, xout0, xoutl, yout0, youtl, 0, 0, f);
        const char* inb = (const char*)in + (size_t)b * in_bstride * sizeof(cf);
        const char* eyb = ey ? (const char*)ey + (size_t)b * in_bstride * sizeof(cf) : 0;
        if (vectorial) { int rc = set_planes(c, inb, eyb); if (rc) return rc; }
        int rc = czt_forward(c, inb, (char*)out + b * oplane, (char*)tables + (z_stride ? b * tb : 0), ws, ws_bytes, st);
        if (rc) return rc;
    }
    return XL_OK;
}
extern "C" int xl_czt_bwd_batch(const void* ct_out, void* ct_in, const double* z, int z_stride, double lambda,
                                int N, int Mx, int My, int vectorial, int nbatch,
                                double x0, double dx, double y0, double dy, double xout0, double xoutl, double yout0, double youtl,
                                int flags, void* tables, void* ws, size_t ws_bytes, void* stream) {
    if (nbatch < 1 || z_stride < 0 || !z) return xl_fail(XL_E_BAD_ARG, "xl_czt_bwd_batch: bad batch shape%s", "");
    xl_stream_t st = (xl_stream_t)stream;
    if (!vectorial && z_stride == 0) {
        CztCall c = make_czt_call(0, z, lambda, N, Mx, My, x0, dx, y0, dy, xout0, xoutl, yout0, youtl, 0, 0, flags);
        c.batch = nbatch;
        return czt_backward(c, ct_out, ct_in, tables, ws, ws_bytes, st);
    }
    const size_t tb = czt_tab_bytes(N, Mx, My, 0), oplane = (size_t)(vectorial ? 3 : 1) * My * Mx * sizeof(cf),
                 iplane = (size_t)(vectorial ? 2 : 1) * N * N * sizeof(cf);
    for (int b = 0; b < nbatch; ++b) {
        const int f = flags | ((z_stride == 0 && b > 0) ? XL_REUSE_TABLES : 0);
        CztCall c = make_czt_call(vectorial ? 1 : 0, z + (size_t)b * z_stride, lambda, N, Mx, My, x0, dx, y0, dy, xout0, xoutl, yout0, youtl, 0, 0, f);
        int rc = czt_backward(c, (const char*)ct_out + b * oplane, (char*)ct_in + b * iplane, (char*)tables + (z_stride ? b * tb : 0), ws, ws_bytes, st);
        if (rc) return rc;
    }
    return XL_OK;
}
extern "C" int xl_highna_fwd_batch(const void* ex, const void* ey, long long in_bstride, void* out, int N, int Mx, int My, int nbatch,
                                   double radius, double f, double lambda, double x0, double dx, double y0, double dy,
                                   double xout0, double xoutl, double yout0, double youtl, int flags, void* tables, void* ws, size_t ws_bytes, void* stream) {
    if (nbatch < 1) return xl_fail(XL_E_BAD_ARG, "xl_highna_fwd_batch: bad batch shape%s", "");
    const size_t oplane = (size_t)3 * My * Mx * sizeof(cf);
    for (int b = 0; b < nbatch; ++b) {
        const char* exb = (const char*)ex + (size_t)b * in_bstride * sizeof(cf);
        const char* eyb = ey ? (const char*)ey + (size_t)b * in_bstride * sizeof(cf) : 0;
        int rc = xl_highna_fwd(exb, eyb, (char*)out + b * oplane, N, Mx, My, radius, f, lambda, x0, dx, y0, dy, xout0, xoutl, yout0, youtl,
                               flags | (b > 0 ? XL_REUSE_TABLES : 0), tables, ws, ws_bytes, stream);   // the objective's tables serve every item
        if (rc) return rc;
    }
    return XL_OK;
}
extern "C" int xl_highna_bwd_batch(const void* ct_out, void* ct_exy, int N, int Mx, int My, int nbatch, double radius, double f, double lambda,
                                   double x0, double dx, double y0, double dy, double xout0, double xoutl, double yout0, double youtl,
                                   int flags, void* tables, void* ws, size_t ws_bytes, void* stream) {
    if (nbatch < 1) return xl_fail(XL_E_BAD_ARG, "xl_highna_bwd_batch: bad batch shape%s", "");
    const size_t oplane = (size_t)3 * My * Mx * sizeof(cf), iplane = (size_t)2 * N * N * sizeof(cf);
    for (int b = 0; b < nbatch; ++b) {
        int rc = xl_highna_bwd((const char*)ct_out + b * oplane, (char*)ct_exy + b * iplane, N, Mx, My, radius, f, lambda, x0, dx, y0, dy,
                               xout0, xoutl, yout0, youtl, flags | (b > 0 ? XL_REUSE_TABLES : 0), tables, ws, ws_bytes, stream);
        if (rc) return rc;
    }
    return XL_OK;
}
