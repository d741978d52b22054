"""Kernel BODIES (xlumina_b200/csrc/*.cuh compiled with g++ -DXL_HOST_EMU, see conftest.emu) against the golden fixtures
and the oracle, on the CPU.  This exercises every line of the device code's index math, butterflies, twiddles, chirp
tables and fused factors through the same C ABI the GPU library exports; the GPU parity proper is tests/test_gpu_parity.py.
Tolerance: rel-L2 <= 1e-4 vs the complex128 reference (BASELINE.json north_star); observed ~3e-7."""
import ctypes
import os

import numpy as np
import pytest

from conftest import golden, rel_l2, ROOT
from oracle import oracle_np as o

TOL = 1e-4
TIGHT = 5e-6


def ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


def c64(a):
    return np.ascontiguousarray(a, dtype=np.complex64)


def rs_fwd(emu, field, x, y, lam, z, nfields=1):
    N = field.shape[-1]
    fin = c64(field)
    out = np.zeros_like(fin)
    H = np.zeros(emu.xl_rs_transfer_bytes(N), np.uint8)
    ws = np.zeros(emu.xl_rs_workspace_bytes(N, nfields, 1), np.uint8)
    zz = np.array([z], np.float64)
    rc = emu.xl_rs_fwd(ptr(fin), ptr(out), ptr(H), ptr(zz), N, nfields, x[1] - x[0], y[1] - y[0], 2 * np.pi / lam, 0,
                       ptr(ws), ws.size, None)
    assert rc == 0, emu.xl_last_error()
    return out, H, ws, zz


@pytest.mark.parametrize("name", ["rs_n32_zpos", "rs_n32_zneg", "rs_n48_far"])
def test_rs_forward_and_vjp_golden(emu, name):
    g = golden(name)
    x, y, lam, z = g["x"], g["y"], float(g["wavelength"]), float(g["z"])
    N = len(x)
    out, H, ws, zz = rs_fwd(emu, g["field"], x, y, lam, z)
    assert rel_l2(out, g["out"]) < TIGHT
    ct = c64(g["ct"])
    gin = np.zeros((N, N), np.complex64)
    gz = np.zeros(1)
    rc = emu.xl_rs_bwd(ptr(c64(g["field"])), ptr(out), ptr(ct), ptr(gin), ptr(gz), ptr(H), ptr(zz), N, 1, x[1] - x[0], y[1] - y[0],
                       2 * np.pi / lam, 0, ptr(ws), ws.size, None)
    assert rc == 0, emu.xl_last_error()
    if "vjp_field" in g:
        assert rel_l2(gin, g["vjp_field"]) < TIGHT
    assert abs(gz[0] - float(g["vjp_z"])) < TOL * abs(float(g["vjp_z"]))


def test_rs_batched_fields_and_reuse_flag(emu):
    g = golden("rs_n32_zpos")
    x, y, lam, z = g["x"], g["y"], float(g["wavelength"]), float(g["z"])
    rng = np.random.default_rng(0)
    f3 = np.stack([g["field"], rng.standard_normal((32, 32)) + 0j, 1j * g["field"]])
    out, H, ws, zz = rs_fwd(emu, f3, x, y, lam, z, nfields=3)
    assert rel_l2(out[0], g["out"]) < TIGHT and rel_l2(out[2], 1j * g["out"]) < TIGHT
    out2 = np.zeros_like(out)
    rc = emu.xl_rs_fwd(ptr(c64(f3)), ptr(out2), ptr(H), ptr(zz), 32, 3, x[1] - x[0], y[1] - y[0], 2 * np.pi / lam, 16,
                       ptr(ws), ws.size, None)
    assert rc == 0 and np.array_equal(out, out2)


def test_rs_reference_test_config(emu):
    g = golden("scalar_gaussian_n64")
    out, *_ = rs_fwd(emu, g["field"], g["x"], g["y"], float(g["wavelength"]), float(g["z"]))
    assert rel_l2(out, g["rs_out"]) < TIGHT


@pytest.mark.parametrize("N", [300, 600, 1100])
def test_rs_large_padded_lengths(emu, N):
    """Padded lengths 1024, 2048 and 4096 (first radix 4, 8 and 16): pins the radix plans, the slot <-> bin maps behind the
    x/y-mirrored transfer function and the pruned passes of the production sizes on the CPU.  Checked on a band of rows."""
    rng = np.random.default_rng(N)
    x = np.linspace(-3000, 3000, N)
    f = np.zeros((N, N), np.complex128)
    f[N // 3:N // 3 + 40, :] = rng.standard_normal((40, N)) + 1j * rng.standard_normal((40, N))
    ref, _ = o.RS_propagation(f, x, x, 0.6328, 40000.0)
    out, *_ = rs_fwd(emu, f, x, x, 0.6328, 40000.0)
    assert rel_l2(out, ref) < TIGHT


def test_rs_and_czt_at_the_baseline_size_2048(emu):
    """BASELINE.json's size itself (2048 x 2048, lambda = 632.8 nm, window +-15 mm; z = 5 cm for RS, 5 mm for CZT,
    examples/scalar_xlumina.py:22-33): the kernel bodies against the complex128 oracle at full size, on the CPU.  The GPU suite
    checks the same size through properties only (no CPU oracle run on the GPU box's clock)."""
    o.set_workers(8)
    N, lam = 2048, 0.6328
    x = np.linspace(-15000, 15000, N)
    rng = np.random.default_rng(2048)
    X, Y = np.meshgrid(x, x)
    f = np.exp(-(X ** 2 + Y ** 2) / 1200.0 ** 2) * (1 + 0.05 * (rng.standard_normal((N, N)) + 1j * rng.standard_normal((N, N))))
    ref, _ = o.RS_propagation(f, x, x, lam, 50000.0)
    out, *_ = rs_fwd(emu, f, x, x, lam, 50000.0)
    assert rel_l2(out, ref) < TIGHT
    g = dict(x=x, y=x, xout=x, yout=x, z=5000.0, wavelength=lam)
    cref = o.CZT(f, x, x, lam, 5000.0, x, x)
    cout = np.zeros((N, N), np.complex64)
    czt_call(emu, emu.xl_czt_fwd, c64(f), cout, g, 0)
    assert rel_l2(cout, cref) < TIGHT


def test_rs_gradients_at_the_baseline_size_2048(emu):
    """Backward kernels at 2048 x 2048: the field VJP against the oracle (A is complex-symmetric: A^T ct = A ct) and d/dz
    against a central difference of the oracle, with a random cotangent."""
    o.set_workers(8)
    N, lam, z = 2048, 0.6328, 50000.0
    x = np.linspace(-15000, 15000, N)
    rng = np.random.default_rng(4096)
    X, Y = np.meshgrid(x, x)
    env = np.exp(-(X ** 2 + Y ** 2) / 4000.0 ** 2)
    f = env * (rng.standard_normal((N, N)) + 1j * rng.standard_normal((N, N)))
    ct = env * (rng.standard_normal((N, N)) + 1j * rng.standard_normal((N, N)))
    out, H, ws, zz = rs_fwd(emu, f, x, x, lam, z)
    gin = np.zeros((N, N), np.complex64)
    gz = np.zeros(1)
    rc = emu.xl_rs_bwd(ptr(c64(f)), ptr(out), ptr(c64(ct)), ptr(gin), ptr(gz), ptr(H), ptr(zz), N, 1, x[1] - x[0], x[1] - x[0],
                       2 * np.pi / lam, 0, ptr(ws), ws.size, None)
    assert rc == 0, emu.xl_last_error()
    vref, _ = o.RS_propagation(c64(ct).astype(np.complex128), x, x, lam, z)
    assert rel_l2(gin, vref) < TIGHT
    eps = 1e-3
    fp = c64(f).astype(np.complex128)
    lp = np.sum(c64(ct).astype(np.complex128) * o.RS_propagation(fp, x, x, lam, z + eps)[0]).real
    lm = np.sum(c64(ct).astype(np.complex128) * o.RS_propagation(fp, x, x, lam, z - eps)[0]).real
    gz_ref = (lp - lm) / (2 * eps)
    assert abs(gz[0] - gz_ref) < TOL * abs(gz_ref)


def test_vectorial_paths_at_2048(emu):
    """VRS (examples/vectorial_xlumina.py:21-32 shape) and the high-NA focus 2048 -> 400 (examples/examples.ipynb:387-460
    geometry: r = 1800, f = 2000, +-10 um output window) at full size against the oracle."""
    o.set_workers(8)
    N, lam = 2048, 0.6328
    x = np.linspace(-15000, 15000, N)
    rng = np.random.default_rng(6)
    X, Y = np.meshgrid(x, x)
    env = np.exp(-(X ** 2 + Y ** 2) / 3000.0 ** 2)
    ex = env * (1 + 0.1 * (rng.standard_normal((N, N)) + 1j * rng.standard_normal((N, N))))
    ey = env * (1j + 0.1 * (rng.standard_normal((N, N)) + 1j * rng.standard_normal((N, N))))
    exy = c64(np.stack([ex, ey]))
    k = 2 * np.pi / lam
    out = np.zeros((3, N, N), np.complex64)
    H = np.zeros(emu.xl_rs_transfer_bytes(N), np.uint8)
    ws = np.zeros(emu.xl_rs_workspace_bytes(N, 3, 0), np.uint8)
    zz = np.array([50000.0])
    assert emu.xl_vrs_fwd(ptr(exy), None, ptr(out), ptr(H), ptr(zz), N, x[0], x[0], x[1] - x[0], x[1] - x[0], k, 0, ptr(ws), ws.size, None) == 0
    ref, _ = o.VRS_propagation(exy[0].astype(np.complex128), exy[1].astype(np.complex128), x, x, lam, 50000.0)
    assert rel_l2(out, ref) < TIGHT
    del out, H, ws, ref
    x2 = np.linspace(-1500, 1500, N)
    xo = np.linspace(-10, 10, 400)
    foc = np.zeros((3, 400, 400), np.complex64)
    ws = np.zeros(emu.xl_highna_workspace_bytes(N, 400, 400), np.uint8)
    tb = np.zeros(emu.xl_highna_tables_bytes(N, 400, 400), np.uint8)
    assert emu.xl_highna_fwd(ptr(exy), None, ptr(foc), N, 400, 400, 1800.0, 2000.0, 0.65, x2[0], x2[1] - x2[0], x2[0], x2[1] - x2[0],
                             xo[0], xo[-1], xo[0], xo[-1], 0, ptr(tb), ptr(ws), ws.size, None) == 0, emu.xl_last_error()
    fref = o.VCZT_objective_lens(exy[0].astype(np.complex128), exy[1].astype(np.complex128), x2, x2, 0.65, 1800.0, 2000.0, xo, xo)
    assert rel_l2(foc, fref) < TIGHT


@pytest.mark.parametrize("N", [9, 17, 50, 100])
def test_rs_non_power_of_two_sizes(emu, N):
    rng = np.random.default_rng(N)
    x = np.linspace(-300, 300, N)
    f = rng.standard_normal((N, N)) + 1j * rng.standard_normal((N, N))
    ref, _ = o.RS_propagation(f, x, x, 0.6328, 2500.0)
    out, *_ = rs_fwd(emu, f, x, x, 0.6328, 2500.0)
    assert rel_l2(out, ref) < TIGHT


@pytest.mark.parametrize("name", ["vrs_n24", "vrs_n40_zneg"])
def test_vrs_forward_and_vjp_golden(emu, name):
    g = golden(name)
    x, y, lam, z = g["x"], g["y"], float(g["wavelength"]), float(g["z"])
    N = len(x)
    exy = c64(np.stack([g["Ex"], g["Ey"]]))
    out = np.zeros((3, N, N), np.complex64)
    H = np.zeros(emu.xl_rs_transfer_bytes(N), np.uint8)
    ws = np.zeros(emu.xl_rs_workspace_bytes(N, 3, 1), np.uint8)
    zz = np.array([z])
    k = 2 * np.pi / lam
    assert emu.xl_vrs_fwd(ptr(exy), None, ptr(out), ptr(H), ptr(zz), N, x[0], y[0], x[1] - x[0], y[1] - y[0], k, 0, ptr(ws), ws.size, None) == 0
    assert rel_l2(out, g["out"]) < TIGHT
    gin = np.zeros((2, N, N), np.complex64)
    gz = np.zeros(1)
    assert emu.xl_vrs_bwd(ptr(exy), None, ptr(out), ptr(c64(g["ct"])), ptr(gin), ptr(gz), ptr(H), ptr(zz), N, x[0], y[0], x[1] - x[0], y[1] - y[0],
                          k, 0, ptr(ws), ws.size, None) == 0
    if "vjp_field" in g:
        assert rel_l2(gin, g["vjp_field"]) < TIGHT
    assert abs(gz[0] - float(g["vjp_z"])) < TOL * abs(float(g["vjp_z"]))


def czt_call(emu, fn, a, b, g, vect, flags=0, tables=None):
    """One xl_czt_fwd / xl_czt_bwd call; `tables` (returned) can be passed back with flags | 32 (XL_REUSE_TABLES)."""
    x, y, xo, yo = g["x"], g["y"], g["xout"], g["yout"]
    N, Mx, My = len(x), len(xo), len(yo)
    ws = np.zeros(emu.xl_czt_workspace_bytes(N, Mx, My, vect), np.uint8)
    if tables is None:
        tables = np.zeros(emu.xl_czt_tables_bytes(N, Mx, My), np.uint8)
    zz = np.array([float(g["z"])])
    extra = (None,) if fn is emu.xl_czt_fwd or getattr(fn, "__name__", "") == "xl_czt_fwd" else ()   # forward: (in, ey = NULL: stacked pair)
    rc = fn(ptr(a), *extra, ptr(b), ptr(zz), float(g["wavelength"]), N, Mx, My, vect, x[0], x[1] - x[0], y[0], y[1] - y[0],
            xo[0], xo[-1], yo[0], yo[-1], flags, ptr(tables), ptr(ws), ws.size, None)
    assert rc == 0, emu.xl_last_error()
    return tables


@pytest.mark.parametrize("name", ["czt_n32_m24x40", "czt_n24_m50", "czt_n40_same"])
def test_czt_forward_and_vjp_golden(emu, name):
    g = golden(name)
    out = np.zeros(g["out"].shape, np.complex64)
    czt_call(emu, emu.xl_czt_fwd, c64(g["field"]), out, g, 0)
    assert rel_l2(out, g["out"]) < TIGHT
    if "vjp_field" in g:
        gin = np.zeros(g["field"].shape, np.complex64)
        czt_call(emu, emu.xl_czt_bwd, c64(g["ct"]), gin, g, 0)
        assert rel_l2(gin, g["vjp_field"]) < TIGHT


@pytest.mark.parametrize("N,M", [(520, 300), (1100, 1100)])
def test_czt_large_padded_lengths(emu, N, M):
    """Bluestein lengths 1024 and 4096: the paired + pruned kernel variants and the rotated forward kernel table at
    production sizes, forward and adjoint, against the oracle (adjoint through the dot-product identity)."""
    rng = np.random.default_rng(N + M)
    x = np.linspace(-1500, 1500, N)
    xo = np.linspace(-400, 400, M)
    f = np.zeros((N, N), np.complex128)
    f[N // 2 - 8:N // 2 + 8, :] = rng.standard_normal((16, N)) + 1j * rng.standard_normal((16, N))
    g = dict(x=x, y=x, xout=xo, yout=xo, z=30000.0, wavelength=0.6328)
    ref = o.CZT(f, x, x, 0.6328, 30000.0, xo, xo)
    out = np.zeros((M, M), np.complex64)
    czt_call(emu, emu.xl_czt_fwd, c64(f), out, g, 0)
    assert rel_l2(out, ref) < TIGHT
    ct = np.zeros((M, M), np.complex64)
    ct[M // 2 - 4:M // 2 + 4, :] = (rng.standard_normal((8, M)) + 1j * rng.standard_normal((8, M))).astype(np.complex64)
    gin = np.zeros((N, N), np.complex64)
    czt_call(emu, emu.xl_czt_bwd, ct, gin, g, 0)
    lhs = np.sum(ct.astype(np.complex128) * ref)            # <ct, K f> = <K^T ct, f>   (no conjugation: JAX convention)
    rhs = np.sum(gin.astype(np.complex128) * f)
    assert abs(lhs - rhs) < 2e-5 * abs(lhs)


@pytest.mark.parametrize("N,Mx,My", [(25, 31, 31), (27, 20, 33), (30, 45, 21), (22, 9, 9), (40, 70, 64)])
def test_czt_odd_and_unequal_sizes_generic_variants(emu, N, Mx, My):
    """Odd sizes break the 16-byte pairing, M > L/2 breaks the output pruning: these take the generic kernel variants (8-byte
    accesses, run-time pruning).  Forward against the oracle, adjoint through the dot-product identity."""
    if emu.xl_czt_padded_length(N, Mx) == 0 or emu.xl_czt_padded_length(N, My) == 0:
        pytest.skip("m+M-1 is a power of two: the reference raises too")
    rng = np.random.default_rng(N * 1000 + Mx * 10 + My)
    x = np.linspace(-500, 500, N)
    xo, yo = np.linspace(-150, 150, Mx), np.linspace(-100, 120, My)
    f = rng.standard_normal((N, N)) + 1j * rng.standard_normal((N, N))
    g = dict(x=x, y=x, xout=xo, yout=yo, z=9000.0, wavelength=0.6328)
    ref = o.CZT(f, x, x, 0.6328, 9000.0, xo, yo)
    out = np.zeros((My, Mx), np.complex64)
    czt_call(emu, emu.xl_czt_fwd, c64(f), out, g, 0)
    assert rel_l2(out, ref) < TIGHT
    ct = c64(rng.standard_normal((My, Mx)) + 1j * rng.standard_normal((My, Mx)))
    gin = np.zeros((N, N), np.complex64)
    czt_call(emu, emu.xl_czt_bwd, ct, gin, g, 0)
    lhs, rhs = np.sum(ct.astype(np.complex128) * ref), np.sum(gin.astype(np.complex128) * f)
    assert abs(lhs - rhs) < 2e-5 * abs(lhs)


def test_vczt_and_highna_odd_output_sizes(emu):
    rng = np.random.default_rng(77)
    N, Mx, My = 26, 15, 19
    x = np.linspace(-400, 400, N)
    xo, yo = np.linspace(-8, 8, Mx), np.linspace(-6, 7, My)
    ex = rng.standard_normal((N, N)) + 1j * rng.standard_normal((N, N))
    ey = rng.standard_normal((N, N)) + 1j * rng.standard_normal((N, N))
    exy = c64(np.stack([ex, ey]))
    g = dict(x=x, y=x, xout=xo, yout=yo, z=7000.0, wavelength=0.6328)
    out = np.zeros((3, My, Mx), np.complex64)
    czt_call(emu, emu.xl_czt_fwd, exy, out, g, 1)
    assert rel_l2(out, o.VCZT(ex, ey, x, x, 0.6328, 7000.0, xo, yo)) < TIGHT
    ws = np.zeros(emu.xl_highna_workspace_bytes(N, Mx, My), np.uint8)
    tb = np.zeros(emu.xl_highna_tables_bytes(N, Mx, My), np.uint8)
    out2 = np.zeros((3, My, Mx), np.complex64)
    assert emu.xl_highna_fwd(ptr(exy), None, ptr(out2), N, Mx, My, 350.0, 500.0, 0.635, x[0], x[1] - x[0], x[0], x[1] - x[0],
                             xo[0], xo[-1], yo[0], yo[-1], 0, ptr(tb), ptr(ws), ws.size, None) == 0
    assert rel_l2(out2, o.VCZT_objective_lens(ex, ey, x, x, 0.635, 350.0, 500.0, xo, yo)) < TIGHT


def test_czt_reference_test_config(emu):
    g = golden("scalar_gaussian_n64")
    g = dict(g, xout=g["x"], yout=g["y"])
    out = np.zeros((64, 64), np.complex64)
    czt_call(emu, emu.xl_czt_fwd, c64(g["field"]), out, g, 0)
    assert rel_l2(out, g["czt_out"]) < TIGHT


def test_vczt_forward_and_vjp_golden(emu):
    g = golden("vczt_n24_m30")
    out = np.zeros(g["out"].shape, np.complex64)
    exy = c64(np.stack([g["Ex"], g["Ey"]]))
    czt_call(emu, emu.xl_czt_fwd, exy, out, g, 1)
    assert rel_l2(out, g["out"]) < TIGHT
    gin = np.zeros((2, 24, 24), np.complex64)
    czt_call(emu, emu.xl_czt_bwd, c64(g["ct"]), gin, g, 1)
    assert rel_l2(gin, g["vjp_field"]) < TIGHT
    # torch convention through the fused conjugation flags: conj(J^T conj(ct))
    gin2 = np.zeros((2, 24, 24), np.complex64)
    czt_call(emu, emu.xl_czt_bwd, c64(np.conj(g["ct"])), gin2, g, 1, flags=3)
    assert rel_l2(np.conj(gin2), g["vjp_field"]) < TIGHT


@pytest.mark.parametrize("name", ["highna_n24_m20", "highna_n40_m30x26"])
def test_highna_forward_and_vjp_golden(emu, name):
    g = golden(name)
    x, y, xo, yo = g["x"], g["y"], g["xout"], g["yout"]
    N, Mx, My = len(x), len(xo), len(yo)
    exy = c64(np.stack([g["Ex"], g["Ey"]]))
    out = np.zeros((3, My, Mx), np.complex64)
    ws = np.zeros(emu.xl_highna_workspace_bytes(N, Mx, My), np.uint8)
    tb = np.zeros(emu.xl_highna_tables_bytes(N, Mx, My), np.uint8)
    geo = (N, Mx, My, float(g["radius"]), float(g["f"]), float(g["wavelength"]), x[0], x[1] - x[0], y[0], y[1] - y[0],
           xo[0], xo[-1], yo[0], yo[-1])
    assert emu.xl_highna_fwd(ptr(exy), None, ptr(out), *geo, 0, ptr(tb), ptr(ws), ws.size, None) == 0, emu.xl_last_error()
    assert rel_l2(out, g["out"]) < TIGHT
    if "vjp_field" in g:
        gin = np.zeros((2, N, N), np.complex64)
        # the backward call reuses the tables of the forward call (XL_REUSE_TABLES = 32)
        assert emu.xl_highna_bwd(ptr(c64(g["ct"])), ptr(gin), *geo, 32, ptr(tb), ptr(ws), ws.size, None) == 0, emu.xl_last_error()
        assert rel_l2(gin, g["vjp_field"]) < TIGHT


def test_highna_odd_n_nan_like_reference(emu):
    N, M = 9, 6
    x = np.linspace(-100, 100, N)
    xo = np.linspace(-1, 1, M)
    exy = np.ones((2, N, N), np.complex64)
    out = np.zeros((3, M, M), np.complex64)
    ws = np.zeros(emu.xl_highna_workspace_bytes(N, M, M), np.uint8)
    tb = np.zeros(emu.xl_highna_tables_bytes(N, M, M), np.uint8)
    assert emu.xl_highna_fwd(ptr(exy), None, ptr(out), N, M, M, 90.0, 100.0, 0.635, x[0], x[1] - x[0], x[0], x[1] - x[0],
                             xo[0], xo[-1], xo[0], xo[-1], 0, ptr(tb), ptr(ws), ws.size, None) == 0
    assert np.isnan(out).any()


def test_error_codes(emu):
    f = np.zeros((8, 8), np.complex64)
    z = np.array([1.0])
    ws = np.zeros(1 << 16, np.uint8)
    H = np.zeros(1 << 16, np.uint8)
    assert emu.xl_rs_fwd(None, ptr(f), ptr(H), ptr(z), 8, 1, 1.0, 1.0, 1.0, 0, ptr(ws), ws.size, None) == -1     # null
    assert emu.xl_rs_fwd(ptr(f), ptr(f), ptr(H), ptr(z), 8, 1, 1.0, 1.0, 1.0, 0, ptr(ws), 16, None) == -3        # workspace
    assert emu.xl_rs_padded_length(4096) == 0 and emu.xl_rs_padded_length(2048) == 4096
    assert emu.xl_czt_padded_length(32, 33) == 0           # m+M-1 == 64: the reference raises too
    assert emu.xl_czt_padded_length(2048, 2048) == 4096 and emu.xl_czt_padded_length(1024, 400) == 2048
    assert b"workspace" in emu.xl_last_error() or emu.xl_last_error() is not None
