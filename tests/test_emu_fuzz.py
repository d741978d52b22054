"""Randomised parity of the kernel bodies (host emulation, same C ABI as the GPU library) against the complex128 oracle:
grid sizes, windows, wavelengths, distances of both signs, output windows.  Small sizes, a few dozen cases, seconds."""
import ctypes

import numpy as np
from hypothesis import given, settings, strategies as st, HealthCheck

from conftest import rel_l2
from oracle import oracle_np as o

TOL = 2e-5


def ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def c64(a):
    return np.ascontiguousarray(a, dtype=np.complex64)


@settings(max_examples=25, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
@given(N=st.integers(4, 40), half=st.floats(50.0, 3000.0), lam=st.floats(0.4, 1.6), zmag=st.floats(200.0, 2e5),
       neg=st.booleans(), seed=st.integers(0, 2 ** 16))
def test_rs_forward_and_field_vjp_random(emu, N, half, lam, zmag, neg, seed):
    rng = np.random.default_rng(seed)
    z = -zmag if neg else zmag
    x = np.linspace(-half, half, N)
    f = rng.standard_normal((N, N)) + 1j * rng.standard_normal((N, N))
    ct = rng.standard_normal((N, N)) + 1j * rng.standard_normal((N, N))
    fin, out = c64(f), np.zeros((N, N), np.complex64)
    H = np.zeros(emu.xl_rs_transfer_bytes(N), np.uint8)
    ws = np.zeros(emu.xl_rs_workspace_bytes(N, 1, 0), np.uint8)
    zz = np.array([z])
    dx, k = x[1] - x[0], 2 * np.pi / lam
    assert emu.xl_rs_fwd(ptr(fin), ptr(out), ptr(H), ptr(zz), N, 1, dx, dx, k, 0, ptr(ws), ws.size, None) == 0
    ref, _ = o.RS_propagation(fin.astype(np.complex128), x, x, lam, z)
    assert rel_l2(out, ref) < TOL
    gin = np.zeros((N, N), np.complex64)
    assert emu.xl_rs_bwd(None, None, ptr(c64(ct)), ptr(gin), None, ptr(H), ptr(zz), N, 1, dx, dx, k, 0, ptr(ws), ws.size, None) == 0
    lhs = np.sum(c64(ct).astype(np.complex128) * ref)              # <ct, A f> = <A^T ct, f>
    rhs = np.sum(gin.astype(np.complex128) * fin.astype(np.complex128))
    assert abs(lhs - rhs) < 5e-5 * abs(lhs)


@settings(max_examples=20, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
@given(N=st.integers(4, 36), Mx=st.integers(2, 40), My=st.integers(2, 40), half=st.floats(100.0, 2000.0),
       frac=st.floats(0.05, 0.9), zmag=st.floats(2000.0, 1e5), seed=st.integers(0, 2 ** 16))
def test_czt_forward_random(emu, N, Mx, My, half, frac, zmag, seed):
    if emu.xl_czt_padded_length(N, Mx) == 0 or emu.xl_czt_padded_length(N, My) == 0:
        return                                                     # m+M-1 a power of two: the reference raises as well
    rng = np.random.default_rng(seed)
    x = np.linspace(-half, half, N)
    xo, yo = np.linspace(-frac * half, frac * half, Mx), np.linspace(-0.7 * frac * half, frac * half, My)
    f = rng.standard_normal((N, N)) + 1j * rng.standard_normal((N, N))
    ref = o.CZT(f, x, x, 0.6328, zmag, xo, yo)
    out = np.zeros((My, Mx), np.complex64)
    ws = np.zeros(emu.xl_czt_workspace_bytes(N, Mx, My, 0), np.uint8)
    tb = np.zeros(emu.xl_czt_tables_bytes(N, Mx, My), np.uint8)
    zz = np.array([zmag])
    rc = emu.xl_czt_fwd(ptr(c64(f)), None, ptr(out), ptr(zz), 0.6328, N, Mx, My, 0, x[0], x[1] - x[0], x[0], x[1] - x[0],
                        xo[0], xo[-1], yo[0], yo[-1], 0, ptr(tb), ptr(ws), ws.size, None)
    assert rc == 0, emu.xl_last_error()
    assert rel_l2(out, ref) < TOL
