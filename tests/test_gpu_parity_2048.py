"""Device-vs-oracle parity at BASELINE.json's full size (2048 x 2048), run on the B200 with -m gpu.

Every operator of the hot path -- RS, VRS, CZT, VCZT, high-NA focus 2048 -> 400 -- is run through the public API (torch
plumbing -> C ABI -> sm_100a kernels) and compared with the torch-CPU complex128 oracle (oracle/oracle_torch.py, autograd
for the gradients) on the same seeded inputs: forward fields, field VJPs and, for RS / VRS, d/dz.  Parameters are those of
BASELINE configs 1 / 2 (examples/scalar_xlumina.py:22-33, examples/vectorial_xlumina.py:21-32 of the reference: window
+-15 mm, 632.8 nm) at both distances the survey names (z = 5 mm, the example file; z = 5 cm, BASELINE.json).
Tolerance: rel-L2 <= 1e-4 in fields and gradients (BASELINE.json north_star)."""
import numpy as np
import pytest

from conftest import rel_l2

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

TOL = 1e-4
N = 2048
LAM = 0.6328


@pytest.fixture(scope="module")
def xb():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import xlumina_b200 as xb
    from xlumina_b200 import _lib
    _lib.lib()
    torch.set_num_threads(max(1, torch.get_num_threads()))
    return xb


def crand(rng, *shape):
    return (rng.standard_normal(shape) + 1j * rng.standard_normal(shape)).astype(np.complex64)


def dev(a):
    return torch.as_tensor(np.ascontiguousarray(a), device="cuda")


def gaussian(x, w0=1200.0):
    X, Y = np.meshgrid(x, x)
    return np.exp(-(X ** 2 + Y ** 2) / w0 ** 2)


def check(name, got, ref, tol=TOL):
    e = rel_l2(got, ref)
    print(f"2048^2 {name}: rel-L2 {e:.3e}")
    assert e < tol, f"{name}: {e}"


@pytest.mark.parametrize("z", [5000.0, 50000.0])
def test_rs_2048_vs_oracle_forward_vjp_dz(xb, z):
    """cfg 1: Gaussian beam (w0 = 1.2 mm) with a random complex modulation, so that every spatial frequency carries signal."""
    from oracle import oracle_torch as ot
    rng = np.random.default_rng(int(z))
    x, _ = xb.space(15000.0, N)
    dx, k = float(x[1] - x[0]), 2 * np.pi / LAM
    u = (gaussian(x) * (1.0 + 0.5 * crand(rng, N, N))).astype(np.complex64)
    ct = crand(rng, N, N)
    U = dev(u).requires_grad_(True)
    zt = torch.tensor([z], dtype=torch.float64, device="cuda", requires_grad=True)
    out = xb.ops.rs_propagation(U, zt, dx, dx, k)
    torch.real(torch.sum(dev(ct) * out)).backward()
    ru = torch.tensor(u.astype(np.complex128), requires_grad=True)
    rz = torch.tensor(z, dtype=torch.float64, requires_grad=True)
    ref = ot.RS_propagation(ru, x, x, LAM, rz)
    torch.real(torch.sum(torch.tensor(ct.astype(np.complex128)) * ref)).backward()
    check(f"RS z={z} forward", out.detach().cpu().numpy(), ref.detach().numpy())
    check(f"RS z={z} field VJP", U.grad.cpu().numpy(), ru.grad.numpy())
    ez = abs(float(zt.grad) - float(rz.grad)) / abs(float(rz.grad))
    print(f"2048^2 RS z={z} d/dz: rel {ez:.3e} ({float(zt.grad):.6e} vs {float(rz.grad):.6e})")
    assert ez < TOL


@pytest.mark.parametrize("z", [5000.0, 50000.0])
def test_vrs_2048_vs_oracle_forward_vjp_dz(xb, z):
    """cfg 2: radially polarised Gaussian beam (examples/vectorial_xlumina.py) with a random complex modulation."""
    from oracle import oracle_torch as ot
    rng = np.random.default_rng(int(z) + 1)
    x, _ = xb.space(15000.0, N)
    dx, k = float(x[1] - x[0]), 2 * np.pi / LAM
    X, Y = np.meshgrid(x, x)
    rho = np.sqrt(X ** 2 + Y ** 2)
    g = gaussian(x)
    ex = (g * X / rho * (1.0 + 0.5 * crand(rng, N, N))).astype(np.complex64)
    ey = (g * Y / rho * (1.0 + 0.5 * crand(rng, N, N))).astype(np.complex64)
    ct = crand(rng, 3, N, N)
    Ex, Ey = dev(ex).requires_grad_(True), dev(ey).requires_grad_(True)
    zt = torch.tensor([z], dtype=torch.float64, device="cuda", requires_grad=True)
    out = xb.ops.vrs_propagation(Ex, Ey, zt, float(x[0]), float(x[0]), dx, dx, k)
    torch.real(torch.sum(dev(ct) * out)).backward()
    rex = torch.tensor(ex.astype(np.complex128), requires_grad=True)
    rey = torch.tensor(ey.astype(np.complex128), requires_grad=True)
    rz = torch.tensor(z, dtype=torch.float64, requires_grad=True)
    ref = ot.VRS_propagation(rex, rey, x, x, LAM, rz)
    torch.real(torch.sum(torch.tensor(ct.astype(np.complex128)) * ref)).backward()
    check(f"VRS z={z} forward", out.detach().cpu().numpy(), ref.detach().numpy())
    check(f"VRS z={z} VJP Ex", Ex.grad.cpu().numpy(), rex.grad.numpy())
    check(f"VRS z={z} VJP Ey", Ey.grad.cpu().numpy(), rey.grad.numpy())
    ez = abs(float(zt.grad) - float(rz.grad)) / abs(float(rz.grad))
    print(f"2048^2 VRS z={z} d/dz: rel {ez:.3e} ({float(zt.grad):.6e} vs {float(rz.grad):.6e})")
    assert ez < TOL


@pytest.mark.parametrize("z", [5000.0, 50000.0])
def test_czt_2048_vs_oracle_forward_vjp(xb, z):
    from oracle import oracle_torch as ot
    rng = np.random.default_rng(int(z) + 2)
    x, _ = xb.space(15000.0, N)
    u = (gaussian(x) * (1.0 + 0.5 * crand(rng, N, N))).astype(np.complex64)
    ct = crand(rng, N, N)
    U = dev(u).requires_grad_(True)
    out = xb.ops.czt(U, z, LAM, x, x, x, x)
    torch.real(torch.sum(dev(ct) * out)).backward()
    ru = torch.tensor(u.astype(np.complex128), requires_grad=True)
    ref = ot.CZT(ru, x, x, LAM, z, x, x)
    torch.real(torch.sum(torch.tensor(ct.astype(np.complex128)) * ref)).backward()
    check(f"CZT z={z} forward", out.detach().cpu().numpy(), ref.detach().numpy())
    check(f"CZT z={z} field VJP", U.grad.cpu().numpy(), ru.grad.numpy())


def test_vczt_2048_vs_oracle_forward_vjp(xb):
    from oracle import oracle_torch as ot
    z = 5000.0
    rng = np.random.default_rng(33)
    x, _ = xb.space(15000.0, N)
    g = gaussian(x)
    ex = (g * (1.0 + 0.5 * crand(rng, N, N))).astype(np.complex64)
    ey = (g * (0.3 + 0.5 * crand(rng, N, N))).astype(np.complex64)
    ct = crand(rng, 3, N, N)
    Ex, Ey = dev(ex).requires_grad_(True), dev(ey).requires_grad_(True)
    out = xb.ops.vczt(Ex, Ey, z, LAM, x, x, x, x)
    torch.real(torch.sum(dev(ct) * out)).backward()
    rex = torch.tensor(ex.astype(np.complex128), requires_grad=True)
    rey = torch.tensor(ey.astype(np.complex128), requires_grad=True)
    ref = ot.VCZT(rex, rey, x, x, LAM, z, x, x)
    torch.real(torch.sum(torch.tensor(ct.astype(np.complex128)) * ref)).backward()
    check("VCZT forward", out.detach().cpu().numpy(), ref.detach().numpy())
    check("VCZT VJP Ex", Ex.grad.cpu().numpy(), rex.grad.numpy())
    check("VCZT VJP Ey", Ey.grad.cpu().numpy(), rey.grad.numpy())


def test_highna_2048_to_400_vs_oracle_forward_vjp(xb):
    """cfg 3 building block at the size of examples/examples.ipynb (2048^2 -> 400^2, NA 0.9 objective)."""
    from oracle import oracle_torch as ot
    rng = np.random.default_rng(44)
    x, _ = xb.space(2500.0, N)
    xo, _ = xb.space(10.0, 400)
    ex, ey = crand(rng, N, N), crand(rng, N, N)
    ct = crand(rng, 3, 400, 400)
    Ex, Ey = dev(ex).requires_grad_(True), dev(ey).requires_grad_(True)
    out = xb.ops.highna_focus(Ex, Ey, 1800.0, 2000.0, 0.635, x, x, xo, xo)
    torch.real(torch.sum(dev(ct) * out)).backward()
    rex = torch.tensor(ex.astype(np.complex128), requires_grad=True)
    rey = torch.tensor(ey.astype(np.complex128), requires_grad=True)
    ref = ot.VCZT_objective_lens(rex, rey, x, x, 0.635, 1800.0, 2000.0, xo, xo)
    torch.real(torch.sum(torch.tensor(ct.astype(np.complex128)) * ref)).backward()
    check("high-NA forward", out.detach().cpu().numpy(), ref.detach().numpy())
    check("high-NA VJP Ex", Ex.grad.cpu().numpy(), rex.grad.numpy())
    check("high-NA VJP Ey", Ey.grad.cpu().numpy(), rey.grad.numpy())
