"""Pointwise optical elements, loss functions and the two optical tables of BASELINE configs 3 and 4
against fixtures produced by the reference's own source (tests/golden/make_golden.py --tables-only).

CPU part: the elements are device-agnostic torch arithmetic, so they are compared directly; for the tables the propagation
seam (ops.vrs_propagation / ops.highna_focus / ops.rs_propagation -- CUDA only) is replaced by the complex128 torch oracle,
which pins the table wiring, parameter maps, losses and their gradients.  The same tables on the real CUDA seam are in
tests/test_gpu_parity.py."""
import math

import numpy as np
import pytest
import torch

from conftest import golden, rel_l2

import xlumina_b200 as xb
from xlumina_b200 import four_f, loss_functions, ops, optical_elements as oe
from xlumina_b200.toolbox import softmin
from oracle import oracle_torch as ot


def _vec(x, lam, comps, dtype):
    li = xb.VectorizedLight(x, x, lam, device="cpu", _alloc=False)
    li.Ex, li.Ey, li.Ez = (torch.as_tensor(c).to(dtype) for c in comps)
    return li


def _comps(li):
    return torch.stack([li.Ex, li.Ey, li.Ez]).numpy()


@pytest.mark.parametrize("dtype,tol", [(torch.complex128, 1e-13), (torch.complex64, 2e-6)])
def test_elements_match_reference(dtype, tol):
    g = golden("elements_n16")
    x, lam = g["x"], float(g["wavelength"])
    a, b = _vec(x, lam, g["a"], dtype), _vec(x, lam, g["b"], dtype)
    assert rel_l2(_comps(oe.sSLM(a, g["alpha"], g["phi"])), g["sslm"]) < tol
    assert rel_l2(_comps(oe.sSLM_with_amplitude(a, g["alpha"], g["phi"], g["A1"], g["A2"])), g["sslm_amp"]) < tol
    assert rel_l2(_comps(oe.LCD(a, float(g["eta"]), float(g["theta"]))), g["lcd"]) < tol
    assert rel_l2(_comps(oe.linear_polarizer(a, g["pol"])), g["lp"]) < tol
    c, d = oe.BS_symmetric(a, b, float(g["bs_theta"]))
    assert rel_l2(_comps(c), g["bs_c"]) < tol and rel_l2(_comps(d), g["bs_d"]) < tol
    s = xb.ScalarLight(x, x, lam, device="cpu", _alloc=False)
    s.field = torch.as_tensor(g["u"]).to(dtype)
    out, slm = oe.SLM(s, torch.as_tensor(g["alpha"]), 16)
    assert rel_l2(out.field.numpy(), g["slm_out"]) < tol and rel_l2(slm.numpy(), g["slm"]) < tol
    ls, mask = oe.lens(s, tuple(g["lens_radius"]), tuple(g["lens_focal"]))
    assert rel_l2(ls.field.numpy(), g["lens_scalar"]) < tol and rel_l2(mask.numpy(), g["lens_mask"]) < tol
    lv, _ = oe.lens(a, tuple(g["lens_radius"]), tuple(g["lens_focal"]))
    assert rel_l2(_comps(lv), g["lens_vector"]) < tol
    _, cyl = oe.cylindrical_lens(s, 5.0e4, 1.5, 0.3)
    av, axi = oe.axicon_lens(a, 0.05)
    # cylindrical lens: thickness = R - sqrt(R^2 - x^2) cancels ~3 digits before it is multiplied by k
    assert rel_l2(cyl.numpy(), g["cyl_mask"]) < max(tol, 1e-10) and rel_l2(axi.numpy(), g["axicon_mask"]) < tol
    assert rel_l2(_comps(av), g["axicon_vector"]) < tol


class TestReferenceElementTests:
    """The reference's own element tests (tests/test_optical_elements.py:32-118, 138-163 -- identities and shapes at its
    configuration: 633 nm, +-1500 um; resolution reduced from 512 to 128) on the mirror, CPU tensors."""
    wavelength, resolution = 633e-3, 128
    x = np.linspace(-1500, 1500, resolution)

    def _source(self, jones):
        li = xb.PolarizedLightSource(self.x, self.x, self.wavelength, device="cpu")
        li.gaussian_beam(w0=(1200, 1200), jones_vector=jones)
        return li

    def test_slm(self):
        light = xb.LightSource(self.x, self.x, self.wavelength, device="cpu")
        light.gaussian_beam(w0=(1200, 1200), E0=1)
        out, _ = oe.SLM(light, torch.zeros((self.resolution, self.resolution)), self.resolution)
        assert out.field.shape == (self.resolution, self.resolution)
        assert torch.allclose(out.field, light.field)

    def test_polarization_devices(self):
        n = self.resolution
        light = self._source((1, 1))
        zero = torch.zeros((n, n))
        out = oe.sSLM(light, zero, zero)
        assert out.Ex.shape == out.Ey.shape == out.Ez.shape == (n, n)
        assert torch.allclose(out.Ex, light.Ex) and torch.allclose(out.Ey, light.Ey) and torch.allclose(out.Ez, light.Ez)
        pi = math.pi * torch.ones((n, n))
        out = oe.sSLM(light, pi, pi)
        assert torch.allclose(out.Ex, -light.Ex, atol=1e-6) and torch.allclose(out.Ey, -light.Ey, atol=1e-6)
        light = self._source((1, 0))
        out = oe.LCD(light, 0, 0)
        assert torch.allclose(out.Ex, light.Ex) and torch.allclose(out.Ey, light.Ey) and torch.allclose(out.Ez, light.Ez)
        out = oe.linear_polarizer(light, zero)                      # aligned with the incident polarisation
        assert torch.allclose(out.Ex, light.Ex) and torch.allclose(out.Ey, light.Ey) and torch.allclose(out.Ez, light.Ez)
        out = oe.linear_polarizer(light, math.pi / 2 * torch.ones((n, n)))   # crossed
        assert torch.allclose(out.Ex, torch.zeros_like(out.Ex), atol=1e-6) and torch.allclose(out.Ey, torch.zeros_like(out.Ey), atol=1e-6)

    def test_beam_splitter(self):
        l1, l2 = self._source((1, 0)), self._source((1, 0))
        c, d = oe.BS_symmetric(l1, l2, 0)                           # fully transmissive
        T, R, noise = 1.0, 0.0, 0.01
        assert torch.allclose(c.Ex, (T - noise) * 1j * l2.Ex + (R - noise) * l1.Ex)
        assert torch.allclose(c.Ey, (T - noise) * 1j * l2.Ey + (R - noise) * l1.Ey)
        assert torch.allclose(d.Ex, (T - noise) * 1j * l1.Ex + (R - noise) * l2.Ex)
        assert torch.allclose(d.Ey, (T - noise) * 1j * l1.Ey + (R - noise) * l2.Ey)

    def test_lenses(self):
        n = self.resolution
        light = xb.LightSource(self.x, self.x, self.wavelength, device="cpu")
        light.gaussian_beam(w0=(1200, 1200), E0=1)
        vec = self._source((1, 0))
        for fn, args in ((oe.lens, ((50, 50), (1000, 1000))), (oe.cylindrical_lens, (1000,)), (oe.axicon_lens, (0.1,))):
            assert fn(light, *args)[0].field.shape == (n, n)
            out, _ = fn(vec, *args)
            assert out.Ex.shape == out.Ey.shape == (n, n)
        with pytest.raises(ValueError):
            bad = xb.ScalarLight(self.x, self.x, self.wavelength, device="cpu")
            bad.info = "something else"
            oe.lens(bad, (50, 50), (1000, 1000))


def test_elements_accept_tensor_parameters_and_differentiate():
    g = golden("elements_n16")
    a = _vec(g["x"], float(g["wavelength"]), g["a"], torch.complex128)
    eta = torch.tensor([float(g["eta"])], dtype=torch.float64, requires_grad=True)
    theta = torch.tensor([float(g["theta"])], dtype=torch.float64, requires_grad=True)
    alpha = torch.tensor(g["alpha"], requires_grad=True)
    out = oe.LCD(oe.sSLM(a, alpha, g["phi"]), eta, theta)
    loss = (out.Ex.abs() ** 2 * torch.as_tensor(g["A1"])).sum() + (out.Ey.real * torch.as_tensor(g["A2"])).sum()
    loss.backward()
    # central differences through the same code
    def f(e, t, al):
        o = oe.LCD(oe.sSLM(a, al, g["phi"]), e, t)
        return float((o.Ex.abs() ** 2 * torch.as_tensor(g["A1"])).sum() + (o.Ey.real * torch.as_tensor(g["A2"])).sum())
    h = 1e-6
    e0, t0, a0 = eta.detach(), theta.detach(), alpha.detach()
    assert abs((f(e0 + h, t0, a0) - f(e0 - h, t0, a0)) / (2 * h) - float(eta.grad)) < 1e-6 * abs(float(eta.grad)) + 1e-7
    assert abs((f(e0, t0 + h, a0) - f(e0, t0 - h, a0)) / (2 * h) - float(theta.grad)) < 1e-6 * abs(float(theta.grad)) + 1e-7
    v = torch.as_tensor(g["pol"])
    fd = (f(e0, t0, a0 + h * v) - f(e0, t0, a0 - h * v)) / (2 * h)
    assert abs(fd - float((alpha.grad * v).sum())) < 1e-6 * abs(fd) + 1e-7


def test_loss_functions_match_reference_definitions():
    g = golden("sharp_focus_n32")
    inten = torch.as_tensor(g["intensities"])
    lv = loss_functions.vectorized_loss_hybrid(inten)
    assert np.allclose(lv.numpy(), g["loss_vec"], rtol=1e-12)
    assert abs(float(softmin(lv)) - float(g["loss_softmin"])) < 1e-12 * abs(float(g["loss_softmin"]))
    assert abs(float(loss_functions.small_area_hybrid(inten[2])) - g["loss_vec"][2]) < 1e-12 * g["loss_vec"][2]
    rng = np.random.default_rng(5)
    a = torch.as_tensor(rng.standard_normal((3, 8, 8)) + 1j * rng.standard_normal((3, 8, 8)))
    b = torch.as_tensor(rng.standard_normal((3, 8, 8)) + 1j * rng.standard_normal((3, 8, 8)))
    an, bn = a.numpy(), b.numpy()
    assert np.allclose(loss_functions.MSE_Intensity(a, b).numpy(), ((abs(an) ** 2 - abs(bn) ** 2) ** 2).sum((1, 2)) / 64)
    assert np.allclose(loss_functions.MSE_Amplitude(a, b).numpy(), ((abs(an) - abs(bn)) ** 2).sum((1, 2)) / 64)
    assert np.allclose(loss_functions.MSE_Phase(a, b).numpy(), ((np.angle(an) - np.angle(bn)) ** 2).sum((1, 2)) / 64)
    m, per = loss_functions.mean_batch_MSE_Intensity(a, b)
    assert np.allclose(float(m), per.numpy().mean())


# ----------------------------------------------------------------------------------------------- tables on the oracle seam
@pytest.fixture
def oracle_seam(monkeypatch):
    """Route the three seam functions the tables call to the complex128 torch oracle (CPU)."""
    def vrs(Ex, Ey, z, x0, y0, dx, dy, k):
        n = Ex.shape[-1]
        x = x0 + dx * np.arange(n)
        y = y0 + dy * np.arange(n)
        return ot.VRS_propagation(Ex, Ey, x, y, 2 * math.pi / k, z)

    def focus(Ex, Ey, radius, f, wavelength, x, y, xout, yout):
        return ot.VCZT_objective_lens(Ex, Ey, x, y, wavelength, radius, f, xout, yout)

    def rs(field, z, dx, dy, k):
        n = field.shape[-1]
        x = dx * (np.arange(n) - (n - 1) / 2)
        return torch.stack([ot.RS_propagation(f, x, x, 2 * math.pi / k, z) for f in field.reshape(-1, n, n)]).reshape(field.shape)
    monkeypatch.setattr(ops, "vrs_propagation", vrs)
    monkeypatch.setattr(ops, "highna_focus", focus)
    monkeypatch.setattr(ops, "rs_propagation", rs)


def sharp_focus_problem(g, device, cdtype):
    x, lam = g["x"], float(g["wavelength"])
    ls = xb.PolarizedLightSource(x, x, lam, device=device)
    ls.Ex = torch.as_tensor(g["Ex"]).to(device=device, dtype=cdtype)
    ls.Ey = torch.as_tensor(g["Ey"]).to(device=device, dtype=cdtype)
    ls.Ez = torch.zeros_like(ls.Ex)
    params = [torch.tensor(g["p%02d" % i], dtype=torch.float64, device=device, requires_grad=True) for i in range(29)]
    fixed = [float(g["radius"]), float(g["f"]), g["xout"], g["xout"]]
    return ls, params, fixed


def sharp_focus_losses(g, ls, params, fixed):
    inten, dets = oe.hybrid_setup_sharp_focus(ls, ls, ls, ls, ls, ls, params, fixed)
    lv = loss_functions.vectorized_loss_hybrid(inten.to(torch.float64))
    W = torch.as_tensor(g["W"], device=lv.device)
    return inten, lv, softmin(lv), (W * inten.to(torch.float64)).sum()     # small-area loss; smooth surrogate sum(W * I)


def directional(g, params, grads, tag, fmt="v_%s_%02d"):
    return sum(float((gr.cpu() * torch.as_tensor(g[fmt % (tag, i)])).sum()) for i, gr in enumerate(grads) if gr is not None)


def test_sharp_focus_table_matches_reference_on_oracle_seam(oracle_seam):
    g = golden("sharp_focus_n32")
    ls, params, fixed = sharp_focus_problem(g, "cpu", torch.complex128)
    inten, lv, l_soft, l_lin = sharp_focus_losses(g, ls, params, fixed)
    assert rel_l2(inten.detach().numpy(), g["intensities"]) < 1e-8      # k*z ~ 1e7 rad in float64
    assert np.allclose(lv.detach().numpy(), g["loss_vec"], rtol=1e-7)
    assert abs(float(l_soft.detach()) - float(g["loss_softmin"])) < 1e-7 * abs(float(g["loss_softmin"]))
    assert abs(float(l_lin.detach()) - float(g["loss_linear"])) < 1e-8 * abs(float(g["loss_linear"]))
    gl = torch.autograd.grad(l_lin, params, retain_graph=True, allow_unused=True)
    gm = torch.autograd.grad(l_soft, params, allow_unused=True)
    for tag in ("all", "dist", "other"):
        want = float(g["dlin_" + tag])
        assert abs(directional(g, params, gl, tag) - want) < 1e-5 * abs(want), tag       # finite differences of the fixture
    want = float(g["dsoft_other"])
    assert abs(directional(g, params, gm, "other") - want) < 1e-5 * abs(want)


def four_f_problem(g, device, cdtype):
    x, lam = g["x"], float(g["wavelength"])
    src = xb.LightSource(x, x, lam, device=device)
    src.field = torch.as_tensor(g["beam"]).to(device=device, dtype=cdtype)
    params = [torch.tensor(g["p%d" % i], dtype=torch.float64, device=device, requires_grad=True) for i in range(5)]
    rd = torch.float64 if cdtype == torch.complex128 else torch.float32
    masks = torch.as_tensor(g["masks"]).to(device=device, dtype=cdtype)
    targets = torch.as_tensor(g["targets"]).to(device=device, dtype=rd)
    return src, params, masks, targets


def test_four_f_table_matches_reference_on_oracle_seam(oracle_seam):
    g = golden("four_f_n32")
    src, params, masks, targets = four_f_problem(g, "cpu", torch.complex128)
    inten, slm1, slm2 = four_f.vector_dualSLM_4f_system(masks, src, params)
    assert rel_l2(inten.detach().numpy(), g["intensities"]) < 1e-8      # k*z ~ 1e7 rad in float64
    one, _, _ = four_f.batch_dualSLM_4f(masks[1], src, params)
    assert rel_l2(one.detach().numpy(), g["intensities"][1]) < 1e-8
    loss = four_f.loss_dualSLM(params, masks, targets, src)
    assert abs(float(loss.detach()) - float(g["loss"])) < 1e-8 * abs(float(g["loss"]))
    grads = torch.autograd.grad(loss, params)
    for tag in ("dist", "phase"):
        want = float(g["dloss_" + tag])
        assert abs(directional(g, params, grads, tag, "v_%s_%d") - want) < 2e-4 * abs(want), tag


def test_reference_toolbox_tests_for_the_mirrored_helpers():
    """tests/test_toolbox.py:33-56 of the reference (space, is_conserving_energy, softmin) on the mirror."""
    from xlumina_b200.toolbox import is_conserving_energy, space
    n = 128
    x, y = space(1500, n)
    assert np.allclose(x, np.linspace(-1500, 1500, n)) and np.allclose(y, np.linspace(-1500, 1500, n))
    l1 = xb.VectorizedLight(x, y, 633e-3, device="cpu")
    l2 = xb.VectorizedLight(x, y, 633e-3, device="cpu")
    l1.Ex = torch.ones((n, n), dtype=torch.complex64)
    l2.Ex = torch.ones((n, n), dtype=torch.complex64)
    assert abs(float(is_conserving_energy(l1, l2)) - 1.0) < 1e-6
    l2.Ex = 0 * l2.Ex
    assert float(is_conserving_energy(l1, l2)) == 0
    assert float(softmin(torch.tensor([1.0, 2.0, 3.0], dtype=torch.float64))) == 1.0


def _torch_elements(kind, planes, params):
    """The elements as the reference composes them (complex128 torch arithmetic), for checking the element kernels."""
    from xlumina_b200 import optical_elements as oe

    def light(ex, ey):
        l = oe.VectorizedLight(np.zeros(2), np.zeros(2), 1.0, ex.device, _alloc=False)
        l.Ex, l.Ey, l.Ez = ex, ey, torch.zeros_like(ex)
        return l
    if kind == "sslm":
        o = oe.sSLM(light(planes[0], planes[1]), params[0], params[1])
        return o.Ex, o.Ey
    if kind == "lcd":
        o = oe.LCD(light(planes[0], planes[1]), params[0], params[1])
        return o.Ex, o.Ey
    c, d = oe.BS_symmetric(light(planes[0], planes[1]), light(planes[2], planes[3]), params[0])
    return c.Ex, c.Ey, d.Ex, d.Ey


def check_element_kernels(kind, device):
    """xl_el_sslm / xl_el_lcd / xl_el_bs (forward, field VJP, parameter gradients with the map p -> p*2pi - pi inside) against
    the same elements composed from complex128 torch arithmetic (the path this file pins to the reference).  Shared by
    tests/test_ops_emu.py (host-emulated kernels) and tests/test_gpu_parity.py (CUDA build)."""
    from xlumina_b200 import ops
    rng = np.random.default_rng(5)
    n = 37          # not a multiple of anything in the kernels
    nplanes = 4 if kind == "bs" else 2
    planes64 = [torch.tensor((rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))).astype(np.complex64), requires_grad=True, device=device) for _ in range(nplanes)]
    planes128 = [p.detach().to(torch.complex128).requires_grad_(True) for p in planes64]
    if kind == "sslm":
        raw = [torch.tensor(rng.uniform(0, 1, (n, n)).astype(np.float32), requires_grad=True, device=device) for _ in range(2)]
    else:
        raw = [torch.tensor(rng.uniform(0, 1, (1,)), dtype=torch.float64, requires_grad=True, device=device) for _ in range(1 if kind == "bs" else 2)]
    raw_ref = [r.detach().to(torch.float64).requires_grad_(True) for r in raw]
    two_pi = 2 * np.pi
    fn = {"sslm": ops.el_sslm, "lcd": ops.el_lcd, "bs": ops.el_bs}[kind]
    outs = fn(*planes64, *raw, two_pi, -np.pi)
    refs = _torch_elements(kind, planes128, [r * two_pi - np.pi for r in raw_ref])
    cts = [torch.tensor((rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))).astype(np.complex64), device=device) for _ in outs]
    for o, r in zip(outs, refs):
        assert rel_l2(o.detach().cpu().numpy(), r.detach().cpu().numpy()) < 2e-6
    # drop one output from the loss of the beam splitter: its cotangent arrives as None
    use = range(len(outs)) if kind != "bs" else (0, 1, 2)
    loss = sum((torch.conj(cts[i]) * outs[i]).real.sum() for i in use)
    loss_ref = sum((torch.conj(cts[i].to(torch.complex128)) * refs[i]).real.sum() for i in use)
    g = torch.autograd.grad(loss, planes64 + raw)
    g_ref = torch.autograd.grad(loss_ref, planes128 + raw_ref)
    for a, b in zip(g[:nplanes], g_ref[:nplanes]):
        assert rel_l2(a.cpu().numpy(), b.cpu().numpy()) < 2e-6
    for a, b in zip(g[nplanes:], g_ref[nplanes:]):
        assert rel_l2(a.double().cpu().numpy(), b.cpu().numpy()) < 2e-5
