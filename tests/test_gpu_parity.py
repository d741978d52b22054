"""GPU parity tests (run on the B200 with -m gpu).  Everything goes through the public Python API -> torch plumbing ->
the C ABI of libxlprop.so -> the sm_100a kernels, and is compared with the CPU oracle / golden fixtures.
Tolerance: rel-L2 <= 1e-4 vs the complex128 reference in fields and gradients (BASELINE.json north_star)."""
import numpy as np
import pytest

from conftest import golden, rel_l2

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

TOL = 1e-4


@pytest.fixture(scope="module")
def xb():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import xlumina_b200 as xb
    from xlumina_b200 import _lib
    _lib.lib()           # must load the in-tree CUDA library; no fallback exists
    return xb


def dev_c64(a):
    return torch.as_tensor(np.ascontiguousarray(a, dtype=np.complex64), device="cuda")


def crand(rng, *shape):
    return rng.standard_normal(shape) + 1j * rng.standard_normal(shape)


# ------------------------------------------------------------------------------------------- golden fixtures
@pytest.mark.parametrize("name", ["rs_n32_zpos", "rs_n32_zneg", "rs_n48_far"])
def test_rs_golden_forward_and_gradients(xb, name):
    g = golden(name)
    lam, z = float(g["wavelength"]), float(g["z"])
    light = xb.ScalarLight(g["x"], g["y"], lam)
    light.field = dev_c64(g["field"]).requires_grad_(True)
    zt = torch.tensor(z, dtype=torch.float64, device="cuda", requires_grad=True)
    out, q = light.RS_propagation(zt)
    assert rel_l2(out.field.detach().cpu().numpy(), g["out"]) < TOL
    assert abs(float(q) - float(g["quality"])) < 1e-9 * float(g["quality"])
    L = torch.real(torch.sum(dev_c64(g["ct"]) * out.field))      # JAX cotangent ct  <=>  torch grad conj(ct)
    L.backward()
    if "vjp_field" in g:
        assert rel_l2(np.conj(light.field.grad.cpu().numpy()), g["vjp_field"]) < TOL
    assert abs(float(zt.grad) - float(g["vjp_z"])) < TOL * abs(float(g["vjp_z"]))


def test_reference_test_configs_scalar(xb):
    """tests/test_wave_optics.py:43-53 of the reference (shape assertions) + values from the golden run at N=64."""
    g = golden("scalar_gaussian_n64")
    lam, z = float(g["wavelength"]), float(g["z"])
    src = xb.LightSource(g["x"], g["y"], lam)
    src.gaussian_beam(w0=(1200, 1200), E0=1)
    assert rel_l2(src.field.cpu().numpy(), g["field"]) < 1e-6
    out, _ = src.RS_propagation(z=z)
    assert out.field.shape == (64, 64)
    assert rel_l2(out.field.cpu().numpy(), g["rs_out"]) < TOL
    c = src.CZT(z=z)
    assert c.field.shape == (64, 64)
    assert rel_l2(c.field.cpu().numpy(), g["czt_out"]) < TOL


@pytest.mark.parametrize("name", ["vrs_n24", "vrs_n40_zneg"])
def test_vrs_golden_forward_and_gradients(xb, name):
    g = golden(name)
    lam, z = float(g["wavelength"]), float(g["z"])
    li = xb.VectorizedLight(g["x"], g["y"], lam)
    li.Ex = dev_c64(g["Ex"]).requires_grad_(True)
    li.Ey = dev_c64(g["Ey"]).requires_grad_(True)
    zt = torch.tensor(z, dtype=torch.float64, device="cuda", requires_grad=True)
    out, _ = li.VRS_propagation(zt)
    E = torch.stack([out.Ex, out.Ey, out.Ez])
    assert rel_l2(E.detach().cpu().numpy(), g["out"]) < TOL
    torch.real(torch.sum(dev_c64(g["ct"]) * E)).backward()
    if "vjp_field" in g:
        got = np.conj(np.stack([li.Ex.grad.cpu().numpy(), li.Ey.grad.cpu().numpy()]))
        assert rel_l2(got, g["vjp_field"]) < TOL
    assert abs(float(zt.grad) - float(g["vjp_z"])) < TOL * abs(float(g["vjp_z"]))


@pytest.mark.parametrize("name", ["czt_n32_m24x40", "czt_n24_m50", "czt_n40_same"])
def test_czt_golden_forward_and_gradient(xb, name):
    g = golden(name)
    li = xb.ScalarLight(g["x"], g["y"], float(g["wavelength"]))
    li.field = dev_c64(g["field"]).requires_grad_(True)
    out = li.CZT(float(g["z"]), g["xout"], g["yout"])
    assert out.field.shape == g["out"].shape
    assert rel_l2(out.field.detach().cpu().numpy(), g["out"]) < TOL
    if "vjp_field" in g:
        torch.real(torch.sum(dev_c64(g["ct"]) * out.field)).backward()
        assert rel_l2(np.conj(li.field.grad.cpu().numpy()), g["vjp_field"]) < TOL


def test_vczt_golden_forward_and_gradient(xb):
    g = golden("vczt_n24_m30")
    li = xb.VectorizedLight(g["x"], g["y"], float(g["wavelength"]))
    li.Ex = dev_c64(g["Ex"]).requires_grad_(True)
    li.Ey = dev_c64(g["Ey"]).requires_grad_(True)
    out = li.VCZT(float(g["z"]), g["xout"], g["yout"])
    E = torch.stack([out.Ex, out.Ey, out.Ez])
    assert rel_l2(E.detach().cpu().numpy(), g["out"]) < TOL
    torch.real(torch.sum(dev_c64(g["ct"]) * E)).backward()
    got = np.conj(np.stack([li.Ex.grad.cpu().numpy(), li.Ey.grad.cpu().numpy()]))
    assert rel_l2(got, g["vjp_field"]) < TOL


@pytest.mark.parametrize("name", ["highna_n24_m20", "highna_n40_m30x26"])
def test_highna_golden_forward_and_gradient(xb, name):
    g = golden(name)
    li = xb.VectorizedLight(g["x"], g["y"], float(g["wavelength"]))
    li.Ex = dev_c64(g["Ex"]).requires_grad_(True)
    li.Ey = dev_c64(g["Ey"]).requires_grad_(True)
    out = xb.VCZT_objective_lens(li, float(g["radius"]), float(g["f"]), g["xout"], g["yout"])
    E = torch.stack([out.Ex, out.Ey, out.Ez])
    assert rel_l2(E.detach().cpu().numpy(), g["out"]) < TOL
    if "vjp_field" in g:
        torch.real(torch.sum(dev_c64(g["ct"]) * E)).backward()
        got = np.conj(np.stack([li.Ex.grad.cpu().numpy(), li.Ey.grad.cpu().numpy()]))
        assert rel_l2(got, g["vjp_field"]) < TOL


# ------------------------------------------------------------------------------------------- oracle, mid sizes
@pytest.mark.parametrize("N,z", [(100, 4000.0), (256, 50000.0), (512, -20000.0), (1024, 1000.0)])
def test_rs_vs_oracle(xb, N, z):
    from oracle import oracle_np as o
    o.set_workers(8)
    rng = np.random.default_rng(N)
    x, y = xb.space(1500.0, N)
    u = crand(rng, N, N)
    ref, _ = o.RS_propagation(u, x, y, 0.633, z)
    li = xb.ScalarLight(x, y, 0.633)
    li.field = dev_c64(u)
    out, _ = li.RS_propagation(z)
    assert rel_l2(out.field.cpu().numpy(), ref) < TOL


def test_reference_test_configs_vectorial_n1024(xb):
    """tests/test_vectorized_optics.py:54-74 and tests/test_optical_elements.py:128-136 of the reference (their shapes at
    their sizes), plus values against the oracle."""
    from oracle import oracle_np as o
    o.set_workers(8)
    N, lam = 1024, 633e-3
    x = np.linspace(-1500, 1500, N)
    src = xb.PolarizedLightSource(x, x, lam)
    src.gaussian_beam(w0=(1200, 1200), jones_vector=(1, 1))
    ex, ey = src.Ex.cpu().numpy().astype(complex), src.Ey.cpu().numpy().astype(complex)
    v, _ = src.VRS_propagation(z=1000)
    assert v.Ex.shape == v.Ey.shape == v.Ez.shape == (N, N)
    ref, _ = o.VRS_propagation(ex, ey, x, x, lam, 1000)
    assert rel_l2(torch.stack([v.Ex, v.Ey, v.Ez]).cpu().numpy(), ref) < TOL
    c = src.VCZT(1000, x, x)
    assert c.Ex.shape == (N, N)
    assert rel_l2(torch.stack([c.Ex, c.Ey, c.Ez]).cpu().numpy(), o.VCZT(ex, ey, x, x, lam, 1000, x, x)) < TOL
    N2 = 512
    x2 = np.linspace(-1500, 1500, N2)
    s2 = xb.PolarizedLightSource(x2, x2, lam)
    s2.gaussian_beam(w0=(1200, 1200), jones_vector=(1, 0))
    f = xb.VCZT_objective_lens(s2, 1800.0, 2000.0, x2, x2)
    assert f.Ex.shape == (N2, N2)
    ref = o.VCZT_objective_lens(s2.Ex.cpu().numpy().astype(complex), s2.Ey.cpu().numpy().astype(complex), x2, x2, lam,
                                1800.0, 2000.0, x2, x2)
    assert rel_l2(torch.stack([f.Ex, f.Ey, f.Ez]).cpu().numpy(), ref) < TOL


def test_hybrid_config_highna_1024_to_400_with_gradient(xb):
    """cfg 3 building block (experiments/hybrid_sharp_optical_table.py:26-46): N=1024, 2500 um window, 635 nm, NA 0.9."""
    from oracle import oracle_torch as ot
    rng = np.random.default_rng(5)
    N = 1024
    x, y = xb.space(2500.0, N)
    xo, yo = xb.space(10.0, 400)
    ex, ey = crand(rng, N, N), crand(rng, N, N)
    ct = crand(rng, 3, 400, 400)
    li = xb.VectorizedLight(x, y, 0.635)
    li.Ex = dev_c64(ex).requires_grad_(True)
    li.Ey = dev_c64(ey).requires_grad_(True)
    out = xb.VCZT_objective_lens(li, 1800.0, 2000.0, xo, yo)
    E = torch.stack([out.Ex, out.Ey, out.Ez])
    torch.real(torch.sum(dev_c64(ct) * E)).backward()
    tex = torch.tensor(ex, requires_grad=True)
    tey = torch.tensor(ey, requires_grad=True)
    ref = ot.VCZT_objective_lens(tex, tey, x, y, 0.635, 1800.0, 2000.0, xo, yo)
    torch.real(torch.sum(torch.tensor(ct) * ref)).backward()
    assert rel_l2(E.detach().cpu().numpy(), ref.detach().numpy()) < TOL
    assert rel_l2(li.Ex.grad.cpu().numpy(), tex.grad.numpy()) < TOL
    assert rel_l2(li.Ey.grad.cpu().numpy(), tey.grad.numpy()) < TOL


# ------------------------------------------------------------------------------------------- full size (2048^2): properties
def test_rs_2048_properties(xb):
    """At BASELINE.json's size the oracle is slow, so use size-independent properties:
    linearity, complex symmetry <ct, A u> = <A ct, u>, direct-sum spot checks and dz by finite differences."""
    from oracle import oracle_np as o
    rng = np.random.default_rng(7)
    N, lam = 2048, 0.6328
    x, y = xb.space(15000.0, N)
    dx = x[1] - x[0]
    k = 2 * np.pi / lam
    u = rng.standard_normal((N, N)).astype(np.float32) + 1j * rng.standard_normal((N, N)).astype(np.float32)
    ct = rng.standard_normal((N, N)).astype(np.float32) + 1j * rng.standard_normal((N, N)).astype(np.float32)
    U, CT = dev_c64(u), dev_c64(ct)
    z = torch.tensor([50000.0], dtype=torch.float64, device="cuda", requires_grad=True)
    A = lambda f: xb.ops.rs_propagation(f, z, dx, dx, k)
    Au, Act = A(U), A(CT)
    lin = A(2.5 * U - 1j * CT)
    assert rel_l2(lin.detach().cpu().numpy(), (2.5 * Au - 1j * Act).detach().cpu().numpy()) < 1e-5
    lhs = torch.sum(CT.to(torch.complex128) * Au.detach().to(torch.complex128))
    rhs = torch.sum(Act.detach().to(torch.complex128) * U.to(torch.complex128))
    assert abs(lhs - rhs) / abs(lhs) < 1e-5
    pts = [(0, 0), (5, 2040), (1024, 1024), (2047, 2047), (1300, 17)]
    d = o.rs_direct_sum(u.astype(np.complex128), x, y, lam, 50000.0, pts)
    got = Au.detach().cpu().numpy()
    for i, (p, q) in enumerate(pts):
        assert abs(got[p, q] - d[i]) < TOL * np.abs(d).max()
    # d/dz of Re<ct, A(z) u> against central differences of the GPU forward itself (fp64 accumulation of the contraction)
    L = torch.real(torch.sum(CT * Au))
    L.backward()
    eps = 0.01
    with torch.no_grad():
        fp = torch.real(torch.sum(CT.to(torch.complex128) * xb.ops.rs_propagation(U, z + eps, dx, dx, k).to(torch.complex128)))
        fm = torch.real(torch.sum(CT.to(torch.complex128) * xb.ops.rs_propagation(U, z - eps, dx, dx, k).to(torch.complex128)))
    fd = float((fp - fm) / (2 * eps))
    assert abs(float(z.grad) - fd) < 2e-3 * abs(fd)     # FD truncation (k*eps)^2/6 ~ 1.6e-3 dominates this check


def test_czt_2048_adjoint_identity(xb):
    rng = np.random.default_rng(8)
    N = 2048
    x, y = xb.space(15000.0, N)
    u = dev_c64(crand(rng, N, N)).requires_grad_(True)
    ct = dev_c64(crand(rng, N, N))
    du = dev_c64(crand(rng, N, N))
    out = xb.ops.czt(u, 5000.0, 0.6328, x, y, x, y)
    torch.real(torch.sum(ct * out)).backward()
    with torch.no_grad():
        lhs = torch.sum(ct.to(torch.complex128) * xb.ops.czt(du, 5000.0, 0.6328, x, y, x, y).to(torch.complex128))
        rhs = torch.sum(torch.conj(u.grad).to(torch.complex128) * du.to(torch.complex128))
    assert abs(lhs - rhs) / abs(lhs) < 1e-5


def test_energy_conservation_well_sampled(xb):
    """toolbox.is_conserving_energy (reference toolbox.py:74-96) ~ 1 for a well-sampled propagation."""
    from xlumina_b200.toolbox import is_conserving_energy
    x, y = xb.space(1500.0, 1024)
    src = xb.LightSource(x, y, 0.6328)
    src.gaussian_beam(w0=(400, 400), E0=1)
    out, q = src.RS_propagation(20000.0)
    assert abs(float(is_conserving_energy(src, out)) - 1.0) < 1e-3


def test_errors_are_loud(xb):
    from xlumina_b200._lib import XlpropError
    x, y = xb.space(100.0, 33)
    li = xb.ScalarLight(x, y, 0.5)
    with pytest.raises(XlpropError):
        li.CZT(100.0, np.linspace(-1, 1, 32), np.linspace(-1, 1, 32))   # m+M-1 == 64: reference raises too
    with pytest.raises(XlpropError):
        # above the fused path the stage chain needs whole row pairs: an odd size must fail loudly, not approximately
        xb.ops.rs_propagation(torch.zeros(2049, 2049, dtype=torch.complex64, device="cuda"), 1.0, 1.0, 1.0, 1.0)


@pytest.mark.gpu
@pytest.mark.parametrize("fused", [True, False])
def test_four_f_table_loss_and_shared_parameter_gradients(xb, fused):
    """cfg 4 (experiments/four_f_optical_table.py:36-141): mask -> RS -> SLM -> RS -> SLM -> RS -> |.|^2 -> MSE over a batch
    with shared parameters.  Loss and all parameter gradients (3 distances, 2 phase masks) against the torch-CPU
    complex128 restatement."""
    import math
    import torch
    from oracle import oracle_torch as ot
    from scripts.four_f_sharded import forward_loss, synthetic_circles
    N, B, lam = 64, 3, 0.6328
    x, _ = xb.space(1500.0, N)
    dx, k = float(x[1] - x[0]), 2 * math.pi / lam
    rng = np.random.default_rng(5)
    masks, targets = synthetic_circles(B, x, rng)
    X, Y = np.meshgrid(x, x)
    beam = np.exp(-(X ** 2 + Y ** 2) / 1200.0 ** 2)
    pz = [rng.uniform(0.027, 1) for _ in range(3)]
    ph = [rng.uniform(0, 1, (N, N)) for _ in range(2)]
    dev = torch.device("cuda:0")
    params = [torch.tensor([v], dtype=torch.float64, device=dev, requires_grad=True) for v in pz] + \
             [torch.tensor(v.astype(np.float32), device=dev, requires_grad=True) for v in ph]
    loss = forward_loss(params, torch.as_tensor(masks, device=dev).to(torch.complex64), torch.as_tensor(targets, device=dev),
                        torch.as_tensor(beam.astype(np.complex64), device=dev), dx, k, fused=fused)
    loss.backward()
    # oracle
    rp = [torch.tensor(v, dtype=torch.float64, requires_grad=True) for v in pz] + \
         [torch.tensor(v.astype(np.float32).astype(np.float64), requires_grad=True) for v in ph]
    tot = 0
    for b in range(B):
        f = torch.tensor(beam * masks[b], dtype=torch.complex128)
        f = ot.RS_propagation(f, x, x, lam, (rp[0].abs() * 100 + 1.2) * 1e4)
        f = f * torch.exp(1j * (rp[3] * (2 * math.pi) - math.pi))
        f = ot.RS_propagation(f, x, x, lam, (rp[1].abs() * 100 + 1.2) * 1e4)
        f = f * torch.exp(1j * (rp[4] * (2 * math.pi) - math.pi))
        f = ot.RS_propagation(f, x, x, lam, (rp[2].abs() * 100 + 1.2) * 1e4)
        inten = f.real ** 2 + f.imag ** 2
        tot = tot + ((inten - torch.tensor(targets[b], dtype=torch.float64)) ** 2).sum() / (N * N)
    tot.backward()
    assert abs(float(loss.detach()) - float(tot.detach())) < 1e-4 * abs(float(tot.detach()))
    for i in (3, 4):
        assert rel_l2(params[i].grad.cpu().numpy(), rp[i].grad.numpy()) < 1e-3      # fp32 parameters / fp32 phase factors
    # distances: d loss/d z of an intensity loss is a cancellation residue; the library evaluates the
    # cancelling i*k*out term exactly, which leaves errors ~1e-4 of the LARGEST distance gradient of the table
    gz = np.array([float(params[i].grad) for i in range(3)])
    gz_ref = np.array([float(rp[i].grad) for i in range(3)])
    print("four_f distance gradients", "fused" if fused else "unfused", gz, gz_ref, np.abs(gz / gz_ref - 1))
    if fused:   # every plane declared phase-blind (XL_PHASE_BLIND): no residue, each distance gradient within the north star's 1e-4
        assert np.max(np.abs(gz / gz_ref - 1)) < 1e-4
    else:
        assert np.max(np.abs(gz - gz_ref)) < 1e-3 * np.max(np.abs(gz_ref))


@pytest.mark.gpu
@pytest.mark.parametrize("N", [256, 1024])
def test_slab_stage_kernels_single_rank_equal_fused_path(xb, N):
    """The slab-decomposed chain (xl_slab_* stage entry points, xlumina_b200/slab.py) with one rank must reproduce the
    single-GPU path; the 2- and 4-rank decompositions are covered on CPU by tests/test_slab.py and on 2 GPUs by
    scripts/slab_check.py (profiles/scaling_r01.md)."""
    from xlumina_b200 import ops, slab
    rng = np.random.default_rng(N)
    x, _ = xb.space(1500.0, N)
    dx, k = float(x[1] - x[0]), 2 * np.pi / 0.6328
    u = dev_c64(crand(rng, N, N))
    ref = ops.rs_propagation(u, 30000.0, dx, dx, k)
    out, H = slab.rs_propagation_slab(u, 30000.0, dx, dx, k, return_transfer=True)
    assert rel_l2(out.cpu().numpy(), ref.cpu().numpy()) < 2e-6
    ct = dev_c64(crand(rng, N, N))
    vjp = slab.rs_slab_vjp(ct, H)
    assert rel_l2(vjp.cpu().numpy(), ops.rs_propagation(ct, 30000.0, dx, dx, k).cpu().numpy()) < 2e-6


@pytest.mark.gpu
def test_transfer_function_cache_same_z_object(xb):
    """with the cache on, propagations that receive the same z tensor (unmodified) reuse one transfer function;
    results and gradients are identical to the uncached path, and an in-place change of z invalidates the entry."""
    from xlumina_b200 import ops
    rng = np.random.default_rng(11)
    N = 128
    x, _ = xb.space(1500.0, N)
    dx, k = float(x[1] - x[0]), 2 * np.pi / 0.6328
    u1, u2 = dev_c64(crand(rng, N, N)), dev_c64(crand(rng, 2, N, N))
    z = torch.tensor([25000.0], dtype=torch.float64, device="cuda", requires_grad=True)

    def run():
        a = u1.clone().requires_grad_(True)
        o1 = ops.rs_propagation(a, z, dx, dx, k)
        o2 = ops.vrs_propagation(u2, None, z, float(x[0]), float(x[0]), dx, dx, k)
        (o1.abs().sum() + o2.abs().sum()).backward()
        g = (a.grad.clone(), z.grad.clone())
        z.grad = None
        return o1.detach(), o2.detach(), g

    ref = run()
    ops.set_transfer_cache(2)
    try:
        got = run()
        assert len(ops._transfer_cache) == 1            # RS and VRS shared one entry
        for r, g_ in zip(ref[:2], got[:2]):
            assert torch.equal(r, g_)
        assert torch.equal(ref[2][0], got[2][0]) and torch.allclose(ref[2][1], got[2][1], rtol=1e-12)
        with torch.no_grad():
            z += 1000.0                                  # in-place update bumps z._version: the entry must not be reused
        o_new = ops.rs_propagation(u1, z, dx, dx, k)
        assert len(ops._transfer_cache) == 2 and not torch.equal(o_new, ref[0])
    finally:
        ops.set_transfer_cache(0)


@pytest.mark.gpu
@pytest.mark.parametrize("N,z", [(64, 4000.0), (128, -9000.0)])
def test_split_line_kernels_small_on_device(xb, N, z):
    """csrc/xl_long.cuh on the device at small sizes (sub-line length forced to 32: padded length = 4 and 8 sub-lines), against
    the fused path; production sizes (4096^2, 16384^2 point-source checks) are run by scripts/long_check.py."""
    from xlumina_b200 import ops, slab, _lib
    rng = np.random.default_rng(N)
    x, _ = xb.space(600.0, N)
    dx, k = float(x[1] - x[0]), 2 * np.pi / 0.6328
    u = dev_c64(crand(rng, N, N))
    ref = ops.rs_propagation(u, z, dx, dx, k)
    L = _lib.lib()
    L.xl_debug_set_max_line(32)
    try:
        out = slab.rs_propagation_slab(u, z, dx, dx, k)
    finally:
        L.xl_debug_set_max_line(4096)
    assert rel_l2(out.cpu().numpy(), ref.cpu().numpy()) < 2e-6


@pytest.mark.gpu
def test_public_api_routes_large_grids_through_stage_chain(xb):
    """ops.rs_propagation above FUSED_MAX_N uses the slab / split-line stage chain with one rank (the route a 16384^2 field
    takes); here the threshold is lowered so that a 64^2 batch takes it: forward, field gradient and d/dz equal the fused
    path (d/dz: the fused path's Parseval sum against the stage chain's extra propagation with the reduced kernel)."""
    from xlumina_b200 import ops
    rng = np.random.default_rng(2)
    N = 64
    x, _ = xb.space(600.0, N)
    dx, k = float(x[1] - x[0]), 2 * np.pi / 0.6328
    u = dev_c64(crand(rng, 2, N, N))
    ct = dev_c64(crand(rng, 2, N, N))

    def run():
        a = u.clone().requires_grad_(True)
        zt = torch.tensor([7000.0], dtype=torch.float64, device="cuda", requires_grad=True)
        o = ops.rs_propagation(a, zt, dx, dx, k)
        (o * ct).real.sum().backward()
        return o.detach(), a.grad.detach(), float(zt.grad)

    o_ref, g_ref, gz_ref = run()
    old = ops.FUSED_MAX_N
    ops.FUSED_MAX_N = 32
    try:
        o_big, g_big, gz_big = run()
    finally:
        ops.FUSED_MAX_N = old
    assert rel_l2(o_big.cpu().numpy(), o_ref.cpu().numpy()) < 2e-6
    assert rel_l2(g_big.cpu().numpy(), g_ref.cpu().numpy()) < 2e-6
    assert abs(gz_big - gz_ref) < 1e-4 * abs(gz_ref)


@pytest.mark.gpu
@pytest.mark.parametrize("N", [2304, 4608, 8704])
def test_split_line_chain_point_source_at_production_sub_lines(xb, N):
    """csrc/xl_long.cuh with sub-lines of 4096 (padded length 8192 / 16384 / 32768 = 2 / 4 / 8 sub-lines: cluster kernels for
    the first two, the two-launch inverse step for the last): a point source must reproduce the sampled impulse response,
    out[p,q] = dx dy h((q - j0) dx, (p - i0) dy; z) (wave_optics.py:291-297), and propagation must stay linear."""
    import math
    from xlumina_b200 import slab
    lam, z = 0.6328, 5.0e4
    k = 2 * math.pi / lam
    x, _ = xb.space(15000.0, N)
    dx = float(x[1] - x[0])
    assert slab.SlabPlan(N, 1, __import__("xlumina_b200")._lib.lib()).L == {2304: 8192, 4608: 16384, 8704: 32768}[N]
    i0, j0 = N // 3, (2 * N) // 5
    f = torch.zeros(N, N, dtype=torch.complex64, device="cuda")
    f[i0, j0] = 1.0
    out, H = slab.rs_propagation_slab(f, z, dx, dx, k, return_transfer=True, group=slab._LOCAL)
    q = (torch.arange(N, device="cuda", dtype=torch.float64) - j0) * dx
    p = (torch.arange(N, device="cuda", dtype=torch.float64) - i0) * dx
    r = torch.sqrt(p[:, None] ** 2 + q[None, :] ** 2 + z * z)
    ref = (1 / (2 * math.pi)) * z / r ** 2 * (1 / r - 1j * k) * torch.exp(1j * k * r) * dx * dx
    err = float(torch.linalg.norm(out.to(torch.complex128) - ref) / torch.linalg.norm(ref))
    del ref, r
    assert err < 2e-6
    g = torch.Generator(device="cpu").manual_seed(N)
    u = torch.view_as_complex(torch.randn(N, N, 2, generator=g)).to("cuda")
    a = slab.rs_propagation_slab(u, z, dx, dx, k, transfer=H, group=slab._LOCAL)
    b = slab.rs_propagation_slab(u + float(N) * f, z, dx, dx, k, transfer=H, group=slab._LOCAL)   # a source as strong as the field
    assert float(torch.linalg.norm(b - a - float(N) * out) / torch.linalg.norm(float(N) * out)) < 1e-5
    # adjoint identity of the complex-symmetric operator: sum(ct * A u) == sum(A ct * u)
    ct = torch.view_as_complex(torch.randn(N, N, 2, generator=g)).to("cuda")
    v = slab.rs_slab_vjp(ct, H, group=slab._LOCAL)
    lhs, rhs = (ct.to(torch.complex128) * a).sum(), (v.to(torch.complex128) * u).sum()
    assert abs(lhs - rhs) < 1e-5 * abs(lhs)


@pytest.mark.gpu
def test_lazily_conjugated_inputs_are_resolved(xb):
    """torch's .conj() is a lazy view over the unconjugated storage: the raw-pointer boundary must materialise it."""
    from xlumina_b200 import ops
    rng = np.random.default_rng(4)
    N = 64
    x, _ = xb.space(600.0, N)
    dx, k = float(x[1] - x[0]), 2 * np.pi / 0.6328
    u = dev_c64(crand(rng, N, N))
    a = ops.rs_propagation(u.conj(), 5000.0, dx, dx, k)
    b = ops.rs_propagation(torch.conj_physical(u), 5000.0, dx, dx, k)
    assert torch.equal(a, b)


@pytest.mark.gpu
def test_vrs_large_grid_route_equals_fused_path(xb):
    """vrs_propagation above FUSED_MAX_N: Ez formed pointwise, three components through the stage chain as one batch."""
    from xlumina_b200 import ops
    rng = np.random.default_rng(9)
    N = 64
    x, _ = xb.space(600.0, N)
    dx, k = float(x[1] - x[0]), 2 * np.pi / 0.6328
    exy = dev_c64(crand(rng, 2, N, N))
    ct = dev_c64(crand(rng, 3, N, N))

    def run():
        a = exy.clone().requires_grad_(True)
        zt = torch.tensor([6000.0], dtype=torch.float64, device="cuda", requires_grad=True)
        o = ops.vrs_propagation(a, None, zt, float(x[0]), float(x[0]), dx, dx, k)
        (o * ct).real.sum().backward()
        return o.detach(), a.grad.detach(), float(zt.grad)

    o_ref, g_ref, gz_ref = run()
    old = ops.FUSED_MAX_N
    ops.FUSED_MAX_N = 32
    try:
        o_big, g_big, gz_big = run()
    finally:
        ops.FUSED_MAX_N = old
    assert rel_l2(o_big.cpu().numpy(), o_ref.cpu().numpy()) < 2e-6
    assert rel_l2(g_big.cpu().numpy(), g_ref.cpu().numpy()) < 2e-6
    assert abs(gz_big - gz_ref) < 1e-4 * abs(gz_ref)


@pytest.mark.gpu
@pytest.mark.parametrize("cache", [0, 8])
def test_sharp_focus_table_matches_reference(xb, cache):
    """BASELINE config 3 at reduced size: hybrid_setup_sharp_focus (16 VRS + 6 objective focusings between beam splitters,
    sSLMs and wave plates, optical_elements.py:1503-1649) + small_area_hybrid + softmin against the fixture produced by the
    reference's own source; loss gradients w.r.t. all 29 parameters against the fixture's finite differences.  With the
    transfer-function cache on, the repeated distances reuse one transfer function (same results required)."""
    import torch
    from xlumina_b200 import ops
    from test_elements import directional, sharp_focus_losses, sharp_focus_problem
    g = golden("sharp_focus_n32")
    ops.set_transfer_cache(cache)
    try:
        ls, params, fixed = sharp_focus_problem(g, "cuda", torch.complex64)
        inten, lv, l_soft, l_lin = sharp_focus_losses(g, ls, params, fixed)
        gl = torch.autograd.grad(l_lin, params, retain_graph=True, allow_unused=True)
        gm = torch.autograd.grad(l_soft, params, allow_unused=True)
    finally:
        ops.set_transfer_cache(0)
    e_int = rel_l2(inten.detach().cpu().numpy(), g["intensities"])
    errs = {t: abs(directional(g, params, gl, t) - float(g["dlin_" + t])) / abs(float(g["dlin_" + t])) for t in ("all", "dist", "other")}
    e_soft = abs(directional(g, params, gm, "other") - float(g["dsoft_other"])) / abs(float(g["dsoft_other"]))
    print("sharp focus table: intensities", e_int, "gradients", errs, e_soft)
    assert e_int < 1e-4
    assert np.allclose(lv.detach().cpu().numpy(), g["loss_vec"], rtol=1e-3)
    assert abs(float(l_soft.detach()) - float(g["loss_softmin"])) < 1e-3 * abs(float(g["loss_softmin"]))
    assert max(errs.values()) < 1e-3 and e_soft < 1e-3


@pytest.mark.gpu
def test_four_f_module_matches_reference_fixture(xb):
    """BASELINE config 4 at reduced size through xlumina_b200.four_f against the fixture produced by the reference's
    experiments/four_f_optical_table.py (intensities, batch loss, finite-difference directional derivatives)."""
    import torch
    from xlumina_b200 import four_f
    from test_elements import directional, four_f_problem
    g = golden("four_f_n32")
    src, params, masks, targets = four_f_problem(g, "cuda", torch.complex64)
    inten, _, _ = four_f.vector_dualSLM_4f_system(masks, src, params)
    loss = four_f.loss_dualSLM(params, masks, targets, src)
    grads = torch.autograd.grad(loss, params)
    e_int = rel_l2(inten.detach().cpu().numpy(), g["intensities"])
    errs = {t: abs(directional(g, params, grads, t, "v_%s_%d") - float(g["dloss_" + t])) / abs(float(g["dloss_" + t])) for t in ("dist", "phase")}
    print("4f table: intensities", e_int, "gradients", errs)
    assert e_int < 1e-4
    assert abs(float(loss.detach()) - float(g["loss"])) < 1e-4 * abs(float(g["loss"]))
    assert errs["phase"] < 1e-4
    # distance gradients of a phase-blind (intensity) loss are cancellation residues: see the comment in
    # test_four_f_table_loss_and_shared_parameter_gradients and DESIGN.md section 2
    assert errs["dist"] < 3e-3


@pytest.mark.gpu
def test_four_f_fused_elements_on_device(xb):
    """SURVEY 8f-1 / 8f-2 on the B200: beam x mask and both SLM multiplies folded into the first pass, the intensity MSE and its
    cotangent into the last pass (xl_rs_fwd_fused / xl_rs_bwd_fused), every plane declared phase-blind.  Loss, phase-mask
    gradients AND distance gradients within 1e-4 of the reference fixture (the distance gradients of the unfused table are
    complex64 cancellation residues)."""
    import torch
    from xlumina_b200 import four_f
    from test_elements import directional, four_f_problem
    g = golden("four_f_n32")
    src, params, masks, targets = four_f_problem(g, "cuda", torch.complex64)
    loss_u = four_f.loss_dualSLM(params, masks, targets, src)
    grads_u = torch.autograd.grad(loss_u, params)
    loss = four_f.loss_dualSLM_fused(params, masks, targets, src)
    grads = torch.autograd.grad(loss, params)
    errs = {t: abs(directional(g, params, grads, t, "v_%s_%d") - float(g["dloss_" + t])) / abs(float(g["dloss_" + t])) for t in ("dist", "phase")}
    print("4f fused table: loss", float(loss.detach()) / float(g["loss"]) - 1, "gradients", errs)
    assert abs(float(loss.detach()) - float(g["loss"])) < 1e-4 * abs(float(g["loss"]))
    assert errs["phase"] < 1e-4 and errs["dist"] < 1e-4
    for i in (3, 4):
        assert rel_l2(grads[i].cpu().numpy(), grads_u[i].cpu().numpy()) < 1e-4


@pytest.mark.gpu
def test_rs_fused_pieces_on_device(xb):
    """Each fused piece at 1024^2 (batch of 3) against plain torch arithmetic around the unfused operator on the device:
    complex and float32 inputs, shared modulation plane, detection; gradients in field, z, modulation; d/dz-only backward."""
    import torch
    from xlumina_b200 import ops
    rng = np.random.default_rng(5)
    N, F = 1024, 3
    x = np.linspace(-1500.0, 1500.0, N)
    dx, k = float(x[1] - x[0]), 2 * np.pi / 0.6328
    dev = "cuda"

    def crand(*s):
        return torch.tensor((rng.standard_normal(s) + 1j * rng.standard_normal(s)).astype(np.complex64), device=dev)
    z = torch.tensor([31000.0], dtype=torch.float64, device=dev, requires_grad=True)
    u = crand(F, N, N).requires_grad_(True)
    m = crand(N, N).requires_grad_(True)
    ct = crand(F, N, N)
    tgt = torch.tensor(rng.uniform(0, 2, (F, N, N)).astype(np.float32), device=dev)
    w = torch.tensor(rng.uniform(0.5, 1.5, F), device=dev)
    ref = ops.rs_propagation(u * m[None], z, dx, dx, k)
    out = ops.rs_propagation_fused(u, z, dx, dx, k, mod=m)
    assert rel_l2(out.detach().cpu().numpy(), ref.detach().cpu().numpy()) < 1e-6
    gr = torch.autograd.grad((ct.conj() * ref).real.sum(), (u, z, m))
    gf = torch.autograd.grad((ct.conj() * out).real.sum(), (u, z, m))
    assert rel_l2(gf[0].cpu().numpy(), gr[0].cpu().numpy()) < 1e-5 and rel_l2(gf[2].cpu().numpy(), gr[2].cpu().numpy()) < 1e-5
    assert abs(float(gf[1]) - float(gr[1])) < 1e-4 * abs(float(gr[1]))

    def mse_ref(uu, zz, mm):
        o = ops.rs_propagation(uu * mm[None], zz, dx, dx, k)
        return (((o.real ** 2 + o.imag ** 2 - tgt) ** 2).sum(dim=(-2, -1)) / (N * N) * w).sum()
    lr = mse_ref(u, z, m)
    lf = (ops.rs_propagation_fused(u, z, dx, dx, k, mod=m, target=tgt) * w).sum()
    assert abs(float(lf) - float(lr)) < 1e-5 * abs(float(lr))
    gr = torch.autograd.grad(lr, (u, m))
    gf = torch.autograd.grad(lf, (u, m))
    assert rel_l2(gf[0].cpu().numpy(), gr[0].cpu().numpy()) < 1e-4 and rel_l2(gf[1].cpu().numpy(), gr[1].cpu().numpy()) < 1e-4
    masks = torch.tensor((rng.uniform(0, 1, (F, N, N)) > 0.5).astype(np.float32), device=dev)
    beam = m.detach()
    ref = ops.rs_propagation(masks.to(torch.complex64) * beam[None], z, dx, dx, k)
    out = ops.rs_propagation_fused(masks, z, dx, dx, k, mod=beam)
    assert rel_l2(out.detach().cpu().numpy(), ref.detach().cpu().numpy()) < 1e-6
    (gzr,) = torch.autograd.grad((ct.conj() * ref).real.sum(), (z,))
    (gzf,) = torch.autograd.grad((ct.conj() * out).real.sum(), (z,))
    assert abs(float(gzf) - float(gzr)) < 1e-4 * abs(float(gzr))


@pytest.mark.gpu
@pytest.mark.parametrize("vect", [0, 1])
def test_czt_distance_gradient_on_device(xb, vect):
    """SURVEY.md 8f-4 on the B200: d/dz of CZT / VCZT (256^2 -> 200 x 180 region of interest) against a 4th-order central
    difference of the complex128 NumPy oracle, random cotangent; the field gradient of the same backward call against the
    call without d/dz."""
    import torch
    from oracle import oracle_np as o
    from xlumina_b200 import ops
    rng = np.random.default_rng(13 + vect)
    N, lam, z0 = 256, 0.6328, 9000.0
    x = np.linspace(-1200.0, 1200.0, N)
    xo, yo = np.linspace(-150.0, 190.0, 200), np.linspace(-120.0, 160.0, 180)
    shp = (2, N, N) if vect else (N, N)
    u = (rng.standard_normal(shp) + 1j * rng.standard_normal(shp)).astype(np.complex64)
    oshp = (3, len(yo), len(xo)) if vect else (len(yo), len(xo))
    ct = (rng.standard_normal(oshp) + 1j * rng.standard_normal(oshp)).astype(np.complex64)

    def L_ref(z):
        uu = u.astype(np.complex128)
        out = o.VCZT(uu[0], uu[1], x, x, lam, z, xo, yo) if vect else o.CZT(uu, x, x, lam, z, xo, yo)
        return float(np.real(np.sum(np.conj(ct) * out)))
    eps = 2e-5
    gz_ref = (8 * (L_ref(z0 + eps) - L_ref(z0 - eps)) - (L_ref(z0 + 2 * eps) - L_ref(z0 - 2 * eps))) / (12 * eps)
    ut = torch.tensor(u, device="cuda", requires_grad=True)
    zt = torch.tensor([z0], dtype=torch.float64, device="cuda", requires_grad=True)
    ctt = torch.tensor(ct, device="cuda")
    out = ops.vczt(ut, None, zt, lam, x, x, xo, yo) if vect else ops.czt(ut, zt, lam, x, x, xo, yo)
    gu, gz = torch.autograd.grad((ctt.conj() * out).real.sum(), (ut, zt))
    out0 = ops.vczt(ut, None, z0, lam, x, x, xo, yo) if vect else ops.czt(ut, z0, lam, x, x, xo, yo)
    (gu0,) = torch.autograd.grad((ctt.conj() * out0).real.sum(), (ut,))
    print("CZT d/dz on device:", "vect" if vect else "scalar", float(gz), gz_ref, abs(float(gz) - gz_ref) / abs(gz_ref))
    assert abs(float(gz) - gz_ref) < 1e-4 * abs(gz_ref)
    assert rel_l2(gu.cpu().numpy(), gu0.cpu().numpy()) < 1e-6


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["sslm", "lcd", "bs"])
def test_element_kernels_on_device(kind):
    """SURVEY 8f-1, vectorial tables: xl_el_sslm / xl_el_lcd / xl_el_bs (forward, field VJP, parameter gradients) on the B200
    against the same elements composed from complex128 torch arithmetic."""
    from test_elements import check_element_kernels
    check_element_kernels(kind, "cuda:0")


@pytest.mark.gpu
def test_batched_entry_points_on_device():
    """one library call for a batch (xl_rs_*_batch with a distance per item, xl_vrs_*_batch with a shared distance and
    with one per item): outputs, field gradients and distance gradients equal the single-item calls."""
    from xlumina_b200 import ops
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(21)
    N = 96
    x = np.linspace(-900.0, 900.0, N)
    dx, k = float(x[1] - x[0]), 2 * np.pi / 0.6328
    c = lambda *s: torch.tensor((rng.standard_normal(s) + 1j * rng.standard_normal(s)).astype(np.complex64), device=dev)   # noqa: E731
    u = c(3, N, N).requires_grad_(True)
    zs = torch.tensor([9000.0, 11000.0, -8000.0], dtype=torch.float64, device=dev, requires_grad=True)
    ct = c(3, N, N)
    out = ops.rs_propagation(u, zs, dx, dx, k)
    gu, gz = torch.autograd.grad((out * ct).real.sum(), (u, zs))
    for i in range(3):
        ui = u.detach()[i].clone().requires_grad_(True)
        zi = zs.detach()[i:i + 1].clone().requires_grad_(True)
        oi = ops.rs_propagation(ui, zi, dx, dx, k)
        gui, gzi = torch.autograd.grad((oi * ct[i]).real.sum(), (ui, zi))
        assert torch.equal(oi.detach(), out.detach()[i]) and torch.equal(gui, gu[i])
        assert abs(float(gzi) - float(gz[i])) <= 1e-9 * abs(float(gzi))
    ex, ey = c(2, N, N).requires_grad_(True), c(2, N, N).requires_grad_(True)
    ctv = c(2, 3, N, N)
    for z in (torch.tensor([9000.0], dtype=torch.float64, device=dev, requires_grad=True),
              torch.tensor([9000.0, 9900.0], dtype=torch.float64, device=dev, requires_grad=True)):
        ob = ops.vrs_propagation(ex, ey, z, float(x[0]), float(x[0]), dx, dx, k)
        gx, gy, gzv = torch.autograd.grad((ob * ctv).real.sum(), (ex, ey, z))
        tot = 0.0
        for i in range(2):
            exi, eyi = ex.detach()[i].clone().requires_grad_(True), ey.detach()[i].clone().requires_grad_(True)
            zi = z.detach()[(i if z.numel() > 1 else 0):(i if z.numel() > 1 else 0) + 1].clone().requires_grad_(True)
            oi = ops.vrs_propagation(exi, eyi, zi, float(x[0]), float(x[0]), dx, dx, k)
            gxi, gyi, gzi = torch.autograd.grad((oi * ctv[i]).real.sum(), (exi, eyi, zi))
            assert torch.equal(oi.detach(), ob.detach()[i]) and torch.equal(gxi, gx[i]) and torch.equal(gyi, gy[i])
            if z.numel() > 1:
                assert abs(float(gzi) - float(gzv[i])) <= 1e-9 * abs(float(gzi))
            tot += float(gzi)
        if z.numel() == 1:      # the shared distance receives the sum over the batch
            assert abs(tot - float(gzv)) <= 1e-6 * abs(tot)
