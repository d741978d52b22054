"""The oracle (oracle/oracle_np.py) against (1) the golden fixtures produced by executing the reference's own source
(tests/golden/make_golden.py) and (2) the algorithm-independent closed forms of SURVEY.md Appendix A."""
import numpy as np
import pytest

from conftest import golden, rel_l2
from oracle import oracle_np as o

TOL = 1e-11  # float64 restatement vs float64 reference source: only FFT-size / summation-order rounding differs


@pytest.mark.parametrize("name", ["rs_n32_zpos", "rs_n32_zneg", "rs_n48_far"])
def test_rs_matches_reference_source(name):
    g = golden(name)
    out, q = o.RS_propagation(g["field"], g["x"], g["y"], float(g["wavelength"]), float(g["z"]))
    assert rel_l2(out, g["out"]) < TOL
    assert abs(q - float(g["quality"])) < 1e-12 * abs(q)


def test_reference_test_config_gaussian():
    g = golden("scalar_gaussian_n64")  # reference tests/test_wave_optics.py:17-53 at N=64
    lam, z = float(g["wavelength"]), float(g["z"])
    f = o.gaussian_beam(g["x"], g["y"], lam, (1200, 1200), 1.0)
    assert rel_l2(f, g["field"]) < 1e-14
    assert rel_l2(o.RS_propagation(f, g["x"], g["y"], lam, z)[0], g["rs_out"]) < TOL
    assert rel_l2(o.CZT(f, g["x"], g["y"], lam, z), g["czt_out"]) < TOL


@pytest.mark.parametrize("name", ["vrs_n24", "vrs_n40_zneg"])
def test_vrs_matches_reference_source(name):
    g = golden(name)
    out, _ = o.VRS_propagation(g["Ex"], g["Ey"], g["x"], g["y"], float(g["wavelength"]), float(g["z"]))
    assert rel_l2(out, g["out"]) < TOL


@pytest.mark.parametrize("name", ["czt_n32_m24x40", "czt_n24_m50", "czt_n40_same"])
def test_czt_matches_reference_source(name):
    g = golden(name)
    out = o.CZT(g["field"], g["x"], g["y"], float(g["wavelength"]), float(g["z"]), g["xout"], g["yout"])
    assert rel_l2(out, g["out"]) < TOL


@pytest.mark.parametrize("name", ["bluestein_m48_M100", "bluestein_m64_M40"])
def test_bluestein_matches_reference_source_and_closed_form(name):
    g = golden(name)
    x, Dm, f1, f2, M = g["x"], float(g["Dm"]), float(g["f1"]), float(g["f2"]), int(g["M_out"])
    assert rel_l2(o.Bluestein_method(x, f1, f2, Dm, M), g["out"]) < TOL
    K = o.bluestein_matrix(x.shape[0], M, f1, f2, Dm)      # includes the off-by-one and the dropped term
    assert rel_l2((K @ x).T, g["out"]) < 1e-10


def test_vczt_matches_reference_source():
    g = golden("vczt_n24_m30")
    out = o.VCZT(g["Ex"], g["Ey"], g["x"], g["y"], float(g["wavelength"]), float(g["z"]), g["xout"], g["yout"])
    assert rel_l2(out, g["out"]) < TOL


@pytest.mark.parametrize("name", ["highna_n24_m20", "highna_n40_m30x26"])
def test_highna_matches_reference_source(name):
    g = golden(name)
    lens, s = o.high_NA_objective_lens(g["Ex"], g["Ey"], g["x"], g["y"], float(g["radius"]), float(g["f"]))
    assert rel_l2(lens, g["lens"]) < 1e-13 and abs(s - float(g["sin_theta_max"])) < 1e-15
    out = o.VCZT_objective_lens(g["Ex"], g["Ey"], g["x"], g["y"], float(g["wavelength"]), float(g["radius"]), float(g["f"]),
                                g["xout"], g["yout"])
    assert rel_l2(out, g["out"]) < TOL


def test_rs_equals_direct_convolution_sum():
    rng = np.random.default_rng(3)
    N = 24
    x = np.linspace(-50, 50, N)
    f = rng.standard_normal((N, N)) + 1j * rng.standard_normal((N, N))
    out, _ = o.RS_propagation(f, x, x, 0.6328, 300.0)
    pts = [(0, 0), (3, 7), (23, 23), (11, 2)]
    d = o.rs_direct_sum(f, x, x, 0.6328, 300.0, pts)
    for i, (p, q) in enumerate(pts):
        assert abs(out[p, q] - d[i]) < 1e-11 * abs(d[i])


def test_rs_is_complex_symmetric():
    """A = A^T: sum ct*A(u) == sum A(ct)*u -- the property the backward kernels rely on."""
    rng = np.random.default_rng(4)
    N = 20
    x = np.linspace(-80, 80, N)
    u = rng.standard_normal((N, N)) + 1j * rng.standard_normal((N, N))
    ct = rng.standard_normal((N, N)) + 1j * rng.standard_normal((N, N))
    a = np.sum(ct * o.RS_propagation(u, x, x, 0.5, 700.0)[0])
    b = np.sum(o.RS_propagation(ct, x, x, 0.5, 700.0)[0] * u)
    assert abs(a - b) < 1e-12 * abs(a)


def test_highna_odd_n_is_nan_like_reference():
    """0/0 at rho == 0 poisons the output for odd N (optical_elements.py:538)."""
    N = 9
    x = np.linspace(-100, 100, N)
    ex = np.ones((N, N), complex)
    out = o.VCZT_objective_lens(ex, ex, x, x, 0.635, 90.0, 100.0, np.linspace(-1, 1, 6), np.linspace(-1, 1, 6))
    assert np.isnan(out).any()


# ---------------------------------------------------------------------------- torch twin (gradient ground truth)
def test_torch_twin_matches_numpy_oracle_and_golden_vjps():
    import torch
    from oracle import oracle_torch as ot
    g = golden("rs_n32_zpos")
    lam, z = float(g["wavelength"]), float(g["z"])
    u = torch.tensor(g["field"], requires_grad=True)
    zt = torch.tensor(z, dtype=torch.float64, requires_grad=True)
    out = ot.RS_propagation(u, g["x"], g["y"], lam, zt)
    assert rel_l2(out.detach().numpy(), g["out"]) < TOL
    # JAX-convention VJP = conj(torch grad) with cotangent conj'ed:  L = Re sum(ct * out)
    L = torch.real(torch.sum(torch.tensor(g["ct"]) * out))
    L.backward()
    assert rel_l2(np.conj(u.grad.numpy()), g["vjp_field"]) < 1e-10
    assert abs(float(zt.grad) - float(g["vjp_z"])) < 1e-6 * abs(float(g["vjp_z"]))

    g = golden("vrs_n24")
    ex = torch.tensor(g["Ex"], requires_grad=True)
    ey = torch.tensor(g["Ey"], requires_grad=True)
    zt = torch.tensor(float(g["z"]), dtype=torch.float64, requires_grad=True)
    out = ot.VRS_propagation(ex, ey, g["x"], g["y"], float(g["wavelength"]), zt)
    assert rel_l2(out.detach().numpy(), g["out"]) < TOL
    torch.real(torch.sum(torch.tensor(g["ct"]) * out)).backward()
    assert rel_l2(np.conj(np.stack([ex.grad.numpy(), ey.grad.numpy()])), g["vjp_field"]) < 1e-10
    assert abs(float(zt.grad) - float(g["vjp_z"])) < 1e-6 * abs(float(g["vjp_z"]))

    g = golden("czt_n32_m24x40")
    u = torch.tensor(g["field"], requires_grad=True)
    out = ot.CZT(u, g["x"], g["y"], float(g["wavelength"]), float(g["z"]), g["xout"], g["yout"])
    assert rel_l2(out.detach().numpy(), g["out"]) < TOL
    torch.real(torch.sum(torch.tensor(g["ct"]) * out)).backward()
    assert rel_l2(np.conj(u.grad.numpy()), g["vjp_field"]) < 1e-10

    g = golden("highna_n24_m20")
    ex = torch.tensor(g["Ex"], requires_grad=True)
    ey = torch.tensor(g["Ey"], requires_grad=True)
    out = ot.VCZT_objective_lens(ex, ey, g["x"], g["y"], float(g["wavelength"]), float(g["radius"]), float(g["f"]), g["xout"], g["yout"])
    assert rel_l2(out.detach().numpy(), g["out"]) < TOL
    torch.real(torch.sum(torch.tensor(g["ct"]) * out)).backward()
    assert rel_l2(np.conj(np.stack([ex.grad.numpy(), ey.grad.numpy()])), g["vjp_field"]) < 1e-10

    g = golden("vczt_n24_m30")
    ex = torch.tensor(g["Ex"], requires_grad=True)
    ey = torch.tensor(g["Ey"], requires_grad=True)
    out = ot.VCZT(ex, ey, g["x"], g["y"], float(g["wavelength"]), float(g["z"]), g["xout"], g["yout"])
    assert rel_l2(out.detach().numpy(), g["out"]) < TOL
    torch.real(torch.sum(torch.tensor(g["ct"]) * out)).backward()
    assert rel_l2(np.conj(np.stack([ex.grad.numpy(), ey.grad.numpy()])), g["vjp_field"]) < 1e-10
