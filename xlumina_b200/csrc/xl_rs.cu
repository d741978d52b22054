// xl_rs.cu -- C ABI of the RS / VRS path (include/xlprop.h) over the kernels in xl_kernels.cuh / xl_async.cuh.
#include "xl_host.h"
#include "xl_async.cuh"

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
    if (want_grad_z) total += 2 * spec + align_up(L * L * sizeof(cf));   // interleaved (C,W) spectra + the reduced dH/dz
    total += align_up((size_t)3 * N * N * sizeof(cf));  // VRS backward: adjoint of the 3 components before the fold
    return total;
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

// Several transfer functions in ONE launch pair: buffer i (i < count, h_stride_bytes apart) holds, for pairs == 0, H(z[i]) (or
// its reduced z-derivative when deriv), and for pairs == 1 alternately H(z[i/2]) (even i) and the reduced dH/dz (odd i).
extern "C" int xl_rs_transfer_multi(void* H, size_t h_stride_bytes, const double* z, int count, int pairs, int N,
                                    double dx, double dy, double k, int deriv, void* stream) {
    if (!H || !z || count < 1) return xl_fail(XL_E_BAD_ARG, "xl_rs_transfer_multi: bad argument%s", "");
    if (h_stride_bytes % 16 || h_stride_bytes < xl_rs_transfer_bytes(N)) return xl_fail(XL_E_BAD_ARG, "xl_rs_transfer_multi: buffer stride too small or not a multiple of 16%s", "");
    if (pairs && (count & 1)) return xl_fail(XL_E_BAD_ARG, "xl_rs_transfer_multi: pairs need an even count%s", "");
    XlRsParams p;
    int rc = rs_base_params(p, N, dx, dy, k);
    if (rc) return rc;
    xl_stream_t st = (xl_stream_t)stream;
    p.H = (cf*)H; p.z = z;
    p.flags = deriv ? XL_F_DERIV : 0;
    p.h_stride = (long long)(h_stride_bytes / sizeof(cf)); p.h_per_z = pairs ? 2 : 1;
    const int L = p.L;
    p.rows = L; p.hrow0 = 0; p.hstore_all = 0;
    XL_FOR_L(L, rc = xl_launch<XlHRows<XL>>(XlDim{xl_groups(L / 2 + 1), count}, st, p));
    if (rc) return rc;
    XL_FOR_L(L, rc = xl_launch<XlHCols<XL>>(XlDim{L / XL_V, count}, st, p));
    return rc;
}

extern "C" int xl_rs_transfer(void* H, const double* z, int N, double dx, double dy, double k, int deriv, void* stream) {
    if (!H || !z) return xl_fail(XL_E_BAD_ARG, "xl_rs_transfer: null pointer%s", "");
    XlRsParams p;
    int rc = rs_base_params(p, N, dx, dy, k);
    if (rc) return rc;
    return rs_transfer_impl(p, (cf*)H, z, deriv, (xl_stream_t)stream);
}


// rows fwd -> cols conv -> rows inv on `nfields` planes, all fields of a stage in ONE launch.  The column stage is the
// persistent bulk-asynchronous kernel (xl_async.cuh: work items walk the column pairs with the fields of a pair back to
// back; two CTAs per SM).  Measured alternatives that were no faster: one launch per field and stage, a per-field software
// pipeline over auxiliary streams, and bulk-staged row pairs in the row kernels (round 2: 37.7 vs 37.3 us).
static int rs_apply_impl(XlRsParams p, xl_stream_t st) {
    const int L = p.L, N = p.N;
    int rc;
    p.f0 = 0;
    if (!aligned16(p.H) || !aligned16(p.spec)) return xl_fail(XL_E_BAD_ARG, "RS: the transfer-function buffer and the workspace must be 16-byte aligned%s", "");
    XL_FOR_L(L, rc = xl_launch<XlRsRowsFwd<XL>>(XlDim{xl_groups(N), p.nfields}, st, p));
    if (rc) return rc;
    XL_FOR_L(L, rc = xl_launch_persistent<XlRsColsAsync<XL>>((L / XL_V) * p.nfields, 2, st, p));
    if (rc) return rc;
    XL_FOR_L(L, rc = xl_launch<XlRsRowsInv<XL>>(XlDim{xl_groups(N), p.nfields}, st, p));
    return rc;
}

static long long plane_offset(const void* ex, const void* ey, int N) {   // elements from the Ex plane to the Ey plane
    return ey ? (long long)((const cf*)ey - (const cf*)ex) : (long long)N * N;
}
static int rs_fwd_common(const void* in, const void* ey, void* out, void* H, const double* z, int N, int nfields, int vrs,
                         double x0, double y0, double dx, double dy, double k, int flags,
                         void* ws, size_t ws_bytes, xl_stream_t st) {
    if (!in || !out || !H || !z || !ws) return xl_fail(XL_E_BAD_ARG, "rs_fwd: null pointer%s", "");
    XlRsParams p;
    int rc = rs_base_params(p, N, dx, dy, k);
    if (rc) return rc;
    if (ws_bytes < xl_rs_workspace_bytes(N, nfields, 0)) return xl_fail(XL_E_WORKSPACE, "rs_fwd: workspace too small%s", "");
    if (!(flags & XL_REUSE_H)) { rc = rs_transfer_impl(p, (cf*)H, z, 0, st); if (rc) return rc; }
    Carver c{(char*)ws, 0, ws_bytes};
    p.spec = (cf*)c.take((size_t)nfields * p.L * N * sizeof(cf));
    p.in = (const cf*)in; p.out = (cf*)out; p.H = (cf*)H; p.z = z;
    p.ey_off = plane_offset(in, ey, N);
    p.nfields = nfields; p.x0 = x0; p.y0 = y0;
    p.flags = (flags & (XL_CONJ_IN | XL_CONJ_OUT)) | (vrs ? XL_F_VRS : 0);
    return rs_apply_impl(p, st);
}

extern "C" int xl_rs_fwd(const void* in, void* out, void* H, const double* z, int N, int nfields,
                         double dx, double dy, double k, int flags, void* ws, size_t ws_bytes, void* stream) {
    if (nfields < 1) return xl_fail(XL_E_BAD_ARG, "xl_rs_fwd: nfields < 1%s", "");
    return rs_fwd_common(in, 0, out, H, z, N, nfields, 0, 0.0, 0.0, dx, dy, k, flags, ws, ws_bytes, (xl_stream_t)stream);
}
extern "C" int xl_vrs_fwd(const void* ex, const void* ey, void* out, void* H, const double* z, int N, double x0, double y0,
                          double dx, double dy, double k, int flags, void* ws, size_t ws_bytes, void* stream) {
    if (ey && (((const char*)ey - (const char*)ex) % (long long)sizeof(cf))) return xl_fail(XL_E_BAD_ARG, "xl_vrs_fwd: Ex and Ey must be 8-byte aligned%s", "");
    return rs_fwd_common(ex, ey, out, H, z, N, 3, 1, x0, y0, dx, dy, k, flags, ws, ws_bytes, (xl_stream_t)stream);
}

static int rs_bwd_common(const void* in, const void* ey, const void* out, const void* ct_out, void* ct_in, double* grad_z, const void* H,
                         const double* z, int N, int nfields, int vrs, double x0, double y0, double dx, double dy, double k,
                         int flags, void* ws, size_t ws_bytes, xl_stream_t st) {
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
    p.ey_off = in ? plane_offset(in, ey, N) : (long long)N * N;
    cf* dst = vrs ? tmp3 : (cf*)ct_in;

    if (grad_z) {
        p.spec2 = (cf*)c.take(2 * spec_bytes);         // interleaved (cotangent, conj-field) row spectra [f][L][N][2]
        cf* Hz = (cf*)c.take((size_t)L * L * sizeof(cf));
        if (!aligned16(H) || !aligned16(ws)) return xl_fail(XL_E_BAD_ARG, "RS: the transfer-function buffer and the workspace must be 16-byte aligned%s", "");
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
        // row spectra of the cotangent and of conj(U), interleaved per x frequency -> spec2
        {
            XlRsParams pd = p;
            pd.in = (const cf*)ct_out; pd.in2 = (const cf*)in; pd.spec = p.spec2;
            pd.flags = (flags & XL_CONJ_IN) | (vrs ? XL_F_VRS : 0);
            XL_FOR_L(L, rc = xl_launch<XlRsRowsDual<XL>>(XlDim{N, nfields}, st, pd));
            if (rc) return rc;
        }
        XlRsParams pg = p;
        pg.H2 = Hz; pg.gz = grad_z;
        XL_FOR_L(L, rc = xl_launch_persistent<XlRsColsGzAsync<XL>>(L * nfields, 2, st, pg));
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
        f.ex = (const cf*)in; f.ey = in ? (const cf*)in + p.ey_off : 0;
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
    return rs_bwd_common(in, 0, out, ct_out, ct_in, grad_z, H, z, N, nfields, 0, 0.0, 0.0, dx, dy, k, flags, ws, ws_bytes, (xl_stream_t)stream);
}
extern "C" int xl_vrs_bwd(const void* ex, const void* ey, const void* out, const void* ct_out, void* ct_exy, double* grad_z, const void* H,
                          const double* z, int N, double x0, double y0, double dx, double dy, double k, int flags,
                          void* ws, size_t ws_bytes, void* stream) {
    return rs_bwd_common(ex, ey, out, ct_out, ct_exy, grad_z, H, z, N, 3, 1, x0, y0, dx, dy, k, flags, ws, ws_bytes, (xl_stream_t)stream);
}



// ================================================================================================ fused elements (8f-1 / 8f-2)
// xl_rs_fwd_fused / xl_rs_bwd_fused: the scalar RS path with the pointwise elements that bracket it in an optical table
// folded into its first and last pass -- a shared complex modulation plane (phase-only SLM, optical_elements.py:87-103; the
// beam under a batch of real object masks, four_f_optical_table.py:58-65) on the way in, and the intensity-MSE detector
// (four_f_optical_table.py:129-141) on the way out.
static void rs_fuse_params(XlRsParams& p, const xl_rs_fuse* fu) {
    if (!fu) return;
    p.mod = (const cf*)fu->mod; p.in_real = fu->in_real; p.target = fu->target; p.mse = fu->mse;
}

extern "C" int xl_rs_fwd_fused(const void* in, void* out, void* H, const double* z, int N, int nfields,
                               double dx, double dy, double k, int flags, const xl_rs_fuse* fuse,
                               void* ws, size_t ws_bytes, void* stream) {
    xl_stream_t st = (xl_stream_t)stream;
    if (nfields < 1) return xl_fail(XL_E_BAD_ARG, "xl_rs_fwd_fused: nfields < 1%s", "");
    if (!in || !out || !H || !z || !ws) return xl_fail(XL_E_BAD_ARG, "xl_rs_fwd_fused: null pointer%s", "");
    if (fuse && fuse->target && !fuse->mse) return xl_fail(XL_E_BAD_ARG, "xl_rs_fwd_fused: a detection target needs the mse accumulator%s", "");
    XlRsParams p;
    int rc = rs_base_params(p, N, dx, dy, k);
    if (rc) return rc;
    if (ws_bytes < xl_rs_workspace_bytes(N, nfields, 0)) return xl_fail(XL_E_WORKSPACE, "xl_rs_fwd_fused: workspace too small%s", "");
    if (!(flags & XL_REUSE_H)) { rc = rs_transfer_impl(p, (cf*)H, z, 0, st); if (rc) return rc; }
    Carver c{(char*)ws, 0, ws_bytes};
    p.spec = (cf*)c.take((size_t)nfields * p.L * N * sizeof(cf));
    p.in = (const cf*)in; p.out = (cf*)out; p.H = (cf*)H; p.z = z; p.nfields = nfields; p.f0 = 0;
    p.flags = 0;
    rs_fuse_params(p, fuse);
    if (!aligned16(p.H) || !aligned16(p.spec)) return xl_fail(XL_E_BAD_ARG, "RS: the transfer-function buffer and the workspace must be 16-byte aligned%s", "");
    const int L = p.L;
    if (p.mod || p.in_real) { XL_FOR_L(L, rc = xl_launch<XlRsRowsFwd<XL, true>>(XlDim{xl_groups(N), nfields}, st, p)); }
    else { XL_FOR_L(L, rc = xl_launch<XlRsRowsFwd<XL>>(XlDim{xl_groups(N), nfields}, st, p)); }
    if (rc) return rc;
    XL_FOR_L(L, rc = xl_launch_persistent<XlRsColsAsync<XL>>((L / XL_V) * nfields, 2, st, p));
    if (rc) return rc;
    if (p.target) { XL_FOR_L(L, rc = xl_launch<XlRsRowsInv<XL, true>>(XlDim{xl_groups(N), nfields}, st, p)); }
    else { XL_FOR_L(L, rc = xl_launch<XlRsRowsInv<XL>>(XlDim{xl_groups(N), nfields}, st, p)); }
    return rc;
}

extern "C" int xl_rs_bwd_fused(const void* in, const void* out, const void* ct_out, const double* ct_mse, void* ct_in, void* ct_mod,
                               double* grad_z, const void* H, const double* z, int N, int nfields,
                               double dx, double dy, double k, int flags, const xl_rs_fuse* fuse,
                               void* ws, size_t ws_bytes, void* stream) {
    xl_stream_t st = (xl_stream_t)stream;
    if (nfields < 1) return xl_fail(XL_E_BAD_ARG, "xl_rs_bwd_fused: nfields < 1%s", "");
    if (!in || !H || !ws || !z) return xl_fail(XL_E_BAD_ARG, "xl_rs_bwd_fused: null pointer%s", "");
    const bool seed = fuse && fuse->target;
    if (seed ? (!ct_mse || !out) : !ct_out) return xl_fail(XL_E_BAD_ARG, "xl_rs_bwd_fused: missing output cotangent%s", "");
    if (grad_z && !out) return xl_fail(XL_E_BAD_ARG, "xl_rs_bwd_fused: grad_z needs the primal output%s", "");
    const bool need_field = ct_in || ct_mod;
    if (!need_field && !grad_z) return XL_OK;
    XlRsParams p;
    int rc = rs_base_params(p, N, dx, dy, k);
    if (rc) return rc;
    if (ws_bytes < xl_rs_workspace_bytes(N, nfields, grad_z != 0)) return xl_fail(XL_E_WORKSPACE, "xl_rs_bwd_fused: workspace too small%s", "");
    if (!aligned16(H) || !aligned16(ws)) return xl_fail(XL_E_BAD_ARG, "RS: the transfer-function buffer and the workspace must be 16-byte aligned%s", "");
    const int L = p.L;
    Carver c{(char*)ws, 0, ws_bytes};
    const size_t spec_bytes = (size_t)nfields * L * N * sizeof(cf);
    p.spec = (cf*)c.take(spec_bytes);
    c.take((size_t)3 * N * N * sizeof(cf));
    p.nfields = nfields; p.H = (cf*)H; p.z = z; p.f0 = 0;
    XlRsParams pc = p;                              // cotangent-side loads: read ct_out, or seed it from the detector
    pc.in = (const cf*)ct_out;
    pc.flags = flags & XL_CONJ_IN;
    if (seed) { pc.seed_out = (const cf*)out; pc.target = fuse->target; pc.ct_mse = ct_mse; }
    if (grad_z) {
        p.spec2 = (cf*)c.take(2 * spec_bytes);
        cf* Hz = (cf*)c.take((size_t)L * L * sizeof(cf));
        if (fuse && fuse->Hz) {                     // generated ahead by xl_rs_transfer_multi
            if (!aligned16(fuse->Hz)) return xl_fail(XL_E_BAD_ARG, "xl_rs_bwd_fused: Hz must be 16-byte aligned%s", "");
            Hz = (cf*)fuse->Hz;
        } else {
            rc = rs_transfer_impl(p, Hz, z, 1, st);
            if (rc) return rc;
        }
        if (!seed && !(flags & XL_PHASE_BLIND)) {   // the i k out part of d out/dz, exactly; it vanishes identically for the fused
                                                    // detector (xl_seed_ct) and for every phase-blind loss
            XlDotZParams d;
            memset(&d, 0, sizeof(d));
            d.ct = (const cf*)ct_out; d.out = (const cf*)out; d.n = (size_t)nfields * N * N; d.flags = flags & XL_CONJ_IN;
            d.k = k; d.gz = grad_z;
            const size_t per = (size_t)XlDotZ::NT * XlDotZ::PER;
            rc = xl_launch<XlDotZ>(XlDim{(int)((d.n + per - 1) / per), 1}, st, d);
            if (rc) return rc;
        }
        XlRsParams pd = pc;
        pd.in2 = (const cf*)in; pd.spec = p.spec2;
        if (fuse) { pd.mod = (const cf*)fuse->mod; pd.in_real = fuse->in_real; }
        XL_FOR_L(L, rc = xl_launch<XlRsRowsDual<XL, true>>(XlDim{N, nfields}, st, pd));
        if (rc) return rc;
        XlRsParams pg = p;
        pg.H2 = Hz; pg.gz = grad_z;
        pg.flags = need_field ? 0 : XL_F_NOFIELD;
        XL_FOR_L(L, rc = xl_launch_persistent<XlRsColsGzAsync<XL>>(L * nfields, 2, st, pg));
        if (rc) return rc;
    } else {
        XlRsParams pa = pc;
        if (seed) { XL_FOR_L(L, rc = xl_launch<XlRsRowsFwd<XL, true>>(XlDim{xl_groups(N), nfields}, st, pa)); }
        else { XL_FOR_L(L, rc = xl_launch<XlRsRowsFwd<XL>>(XlDim{xl_groups(N), nfields}, st, pa)); }
        if (rc) return rc;
        XL_FOR_L(L, rc = xl_launch_persistent<XlRsColsAsync<XL>>((L / XL_V) * nfields, 2, st, pa));
        if (rc) return rc;
    }
    if (!need_field) return XL_OK;
    XlRsParams po = p;
    po.flags = flags & XL_CONJ_OUT;
    po.out = (cf*)ct_in;
    const cf* mod = fuse ? (const cf*)fuse->mod : 0;
    if (mod || ct_mod) {
        po.in = (const cf*)in; po.in_real = fuse ? fuse->in_real : 0; po.mod = mod; po.ct_mod = (cf*)ct_mod;
        XL_FOR_L(L, rc = xl_launch<XlRsRowsInvMod<XL>>(XlDim{xl_groups(N), 1}, st, po));
    } else {
        XL_FOR_L(L, rc = xl_launch<XlRsRowsInv<XL>>(XlDim{xl_groups(N), nfields}, st, po));
    }
    return rc;
}
