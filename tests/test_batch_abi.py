"""Batched entry points of the C ABI (xl_*_batch, include/xlprop.h; SURVEY.md 8b: the reference vmaps its seam functions over
masks, candidates and noisy distances) on the host-emulated kernels: every batched call must reproduce the single-item calls
bit for bit -- shared distance (z_stride = 0) and one distance per item -- forward, field VJP and d/dz."""
import ctypes

import numpy as np
import pytest

XL_WITH_HZ, XL_REUSE_TABLES = 128, 32


def ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


def cplx(rng, *shape):
    return (rng.standard_normal(shape) + 1j * rng.standard_normal(shape)).astype(np.complex64)


N, B = 24, 3
X0, DX, LAM = -300.0, 600.0 / (N - 1), 0.6328
K = 2 * np.pi / LAM


def ok(emu, rc):
    assert rc == 0, emu.xl_last_error()


@pytest.mark.parametrize("per_item_z", [False, True])
@pytest.mark.parametrize("with_hz", [False, True])
def test_rs_batch_matches_single_calls(emu, per_item_z, with_hz):
    rng = np.random.default_rng(1)
    F = 2
    u, ct = cplx(rng, B, F, N, N), cplx(rng, B, F, N, N)
    zs = np.array([4000.0, 5200.0, -3100.0]) if per_item_z else np.array([4000.0])
    zstr = 1 if per_item_z else 0
    flags = XL_WITH_HZ if with_hz else 0
    tb = emu.xl_rs_transfer_bytes(N) * (2 if with_hz else 1)
    ws = np.zeros(emu.xl_rs_workspace_bytes(N, F * B, 1), np.uint8)
    # single-item calls
    out1, gin1, gz1 = np.zeros_like(u), np.zeros_like(u), np.zeros(B)
    for b in range(B):
        H = np.zeros(tb, np.uint8)
        z = zs[b * zstr:b * zstr + 1].copy()
        ok(emu, emu.xl_rs_fwd(ptr(u[b]), ptr(out1[b]), ptr(H), ptr(z), N, F, DX, DX, K, flags, ptr(ws), ws.size, None))
        ok(emu, emu.xl_rs_bwd(ptr(u[b]), ptr(out1[b]), ptr(ct[b]), ptr(gin1[b]), ptr(gz1[b:b + 1]), ptr(H), ptr(z), N, F, DX, DX, K, flags,
                              ptr(ws), ws.size, None))
    # one batched call each
    Hb = np.zeros(tb * (B if per_item_z else 1), np.uint8)
    out2, gin2 = np.zeros_like(u), np.zeros_like(u)
    gz2 = np.zeros(B if per_item_z else 1)
    ok(emu, emu.xl_rs_fwd_batch(ptr(u), ptr(out2), ptr(Hb), ptr(zs), zstr, N, F, B, DX, DX, K, flags, ptr(ws), ws.size, None))
    ok(emu, emu.xl_rs_bwd_batch(ptr(u), ptr(out2), ptr(ct), ptr(gin2), ptr(gz2), zstr, ptr(Hb), ptr(zs), zstr, N, F, B, DX, DX, K, flags,
                                ptr(ws), ws.size, None))
    assert np.array_equal(out1, out2) and np.array_equal(gin1, gin2)
    if per_item_z:
        assert np.allclose(gz1, gz2, rtol=1e-12, atol=0)
    else:   # the shared distance receives the sum over the batch
        assert abs(gz2[0] - gz1.sum()) <= 1e-6 * np.abs(gz1).sum()   # fp32 partial sums are grouped differently


@pytest.mark.parametrize("per_item_z", [False, True])
@pytest.mark.parametrize("stacked", [False, True])
def test_vrs_batch_matches_single_calls(emu, per_item_z, stacked):
    rng = np.random.default_rng(2)
    if stacked:
        e = cplx(rng, B, 2, N, N)
        ex_of, ey_of, bstride, ey_ptr = (lambda b: e[b, 0]), (lambda b: e[b, 1]), 2 * N * N, None
        ex_all = e
    else:
        ex_all, ey_all = cplx(rng, B, N, N), cplx(rng, B, N, N)
        ex_of, ey_of, bstride, ey_ptr = (lambda b: ex_all[b]), (lambda b: ey_all[b]), N * N, ptr(ey_all)
    ct = cplx(rng, B, 3, N, N)
    zs = np.array([4000.0, 5200.0, 6100.0]) if per_item_z else np.array([4000.0])
    zstr = 1 if per_item_z else 0
    tb = emu.xl_rs_transfer_bytes(N) * 2
    ws = np.zeros(emu.xl_rs_workspace_bytes(N, 3, 1), np.uint8)
    out1, g1, gz1 = np.zeros((B, 3, N, N), np.complex64), np.zeros((B, 2, N, N), np.complex64), np.zeros(B)
    for b in range(B):
        H = np.zeros(tb, np.uint8)
        z = zs[b * zstr:b * zstr + 1].copy()
        exb, eyb = np.ascontiguousarray(ex_of(b)), np.ascontiguousarray(ey_of(b))
        ok(emu, emu.xl_vrs_fwd(ptr(exb), ptr(eyb), ptr(out1[b]), ptr(H), ptr(z), N, X0, X0, DX, DX, K, XL_WITH_HZ, ptr(ws), ws.size, None))
        ok(emu, emu.xl_vrs_bwd(ptr(exb), ptr(eyb), ptr(out1[b]), ptr(ct[b]), ptr(g1[b]), ptr(gz1[b:b + 1]), ptr(H), ptr(z), N, X0, X0, DX, DX, K,
                               XL_WITH_HZ, ptr(ws), ws.size, None))
    Hb = np.zeros(tb * (B if per_item_z else 1), np.uint8)
    out2, g2 = np.zeros_like(out1), np.zeros_like(g1)
    gz2 = np.zeros(B if per_item_z else 1)
    ok(emu, emu.xl_vrs_fwd_batch(ptr(ex_all), ey_ptr, bstride, ptr(out2), ptr(Hb), ptr(zs), zstr, N, B, X0, X0, DX, DX, K, XL_WITH_HZ,
                                 ptr(ws), ws.size, None))
    ok(emu, emu.xl_vrs_bwd_batch(ptr(ex_all), ey_ptr, bstride, ptr(out2), ptr(ct), ptr(g2), ptr(gz2), zstr, ptr(Hb), ptr(zs), zstr, N, B,
                                 X0, X0, DX, DX, K, XL_WITH_HZ, ptr(ws), ws.size, None))
    assert np.array_equal(out1, out2) and np.array_equal(g1, g2)
    if per_item_z:
        assert np.allclose(gz1, gz2, rtol=1e-12, atol=0)
    else:
        assert abs(gz2[0] - gz1.sum()) <= 1e-6 * np.abs(gz1).sum()   # fp32 partial sums are grouped differently


@pytest.mark.parametrize("vectorial", [0, 1])
@pytest.mark.parametrize("per_item_z", [False, True])
def test_czt_batch_matches_single_calls(emu, vectorial, per_item_z):
    rng = np.random.default_rng(3)
    M = 30
    xo0, xol = -40.0, 40.0
    nin, nout = (2, 3) if vectorial else (1, 1)
    u = cplx(rng, B, nin, N, N)
    ct = cplx(rng, B, nout, M, M)
    zs = np.array([9000.0, 9500.0, 10100.0]) if per_item_z else np.array([9000.0])
    zstr = 1 if per_item_z else 0
    tb = emu.xl_czt_tables_bytes(N, M, M)
    ws = np.zeros(emu.xl_czt_workspace_bytes_batch(N, M, M, vectorial, B), np.uint8)
    grid = (X0, DX, X0, DX, xo0, xol, xo0, xol)
    out1, g1 = np.zeros((B, nout, M, M), np.complex64), np.zeros((B, nin, N, N), np.complex64)
    for b in range(B):
        tab = np.zeros(tb, np.uint8)
        z = zs[b * zstr:b * zstr + 1].copy()
        ok(emu, emu.xl_czt_fwd(ptr(u[b]), None, ptr(out1[b]), ptr(z), LAM, N, M, M, vectorial, *grid, 0, ptr(tab), ptr(ws), ws.size, None))
        ok(emu, emu.xl_czt_bwd(ptr(ct[b]), ptr(g1[b]), ptr(z), LAM, N, M, M, vectorial, *grid, XL_REUSE_TABLES, ptr(tab), ptr(ws), ws.size, None))
    tabs = np.zeros(tb * (B if per_item_z else 1), np.uint8)
    out2, g2 = np.zeros_like(out1), np.zeros_like(g1)
    ok(emu, emu.xl_czt_fwd_batch(ptr(u), None, nin * N * N, ptr(out2), ptr(zs), zstr, LAM, N, M, M, vectorial, B, *grid, 0, ptr(tabs),
                                 ptr(ws), ws.size, None))
    ok(emu, emu.xl_czt_bwd_batch(ptr(ct), ptr(g2), ptr(zs), zstr, LAM, N, M, M, vectorial, B, *grid, XL_REUSE_TABLES, ptr(tabs),
                                 ptr(ws), ws.size, None))
    assert np.array_equal(out1, out2) and np.array_equal(g1, g2)


def test_highna_batch_matches_single_calls(emu):
    rng = np.random.default_rng(4)
    M = 20
    e = cplx(rng, B, 2, N, N)
    ct = cplx(rng, B, 3, M, M)
    args = (1800.0, 2000.0, 0.635, -2500.0, 5000.0 / (N - 1), -2500.0, 5000.0 / (N - 1), -10.0, 10.0, -10.0, 10.0)
    tb = emu.xl_highna_tables_bytes(N, M, M)
    ws = np.zeros(emu.xl_highna_workspace_bytes(N, M, M), np.uint8)
    out1, g1 = np.zeros((B, 3, M, M), np.complex64), np.zeros((B, 2, N, N), np.complex64)
    for b in range(B):
        tab = np.zeros(tb, np.uint8)
        ok(emu, emu.xl_highna_fwd(ptr(e[b]), None, ptr(out1[b]), N, M, M, *args, 0, ptr(tab), ptr(ws), ws.size, None))
        ok(emu, emu.xl_highna_bwd(ptr(ct[b]), ptr(g1[b]), N, M, M, *args, XL_REUSE_TABLES, ptr(tab), ptr(ws), ws.size, None))
    tab = np.zeros(tb, np.uint8)
    out2, g2 = np.zeros_like(out1), np.zeros_like(g1)
    ok(emu, emu.xl_highna_fwd_batch(ptr(e), None, 2 * N * N, ptr(out2), N, M, M, B, *args, 0, ptr(tab), ptr(ws), ws.size, None))
    ok(emu, emu.xl_highna_bwd_batch(ptr(ct), ptr(g2), N, M, M, B, *args, XL_REUSE_TABLES, ptr(tab), ptr(ws), ws.size, None))
    assert np.array_equal(out1, out2, equal_nan=True) and np.array_equal(g1, g2, equal_nan=True)
