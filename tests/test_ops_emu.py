"""The torch host layer (xlumina_b200/ops.py: autograd Functions, conjugation flags, workspaces, z handling) and the optical
tables built on it, driven on CPU tensors through the HOST-EMULATED kernel bodies (tests/emu/libxlprop_emu.so, same sources
as the CUDA build compiled with -DXL_HOST_EMU).  Test tooling only: the package itself refuses CPU tensors; here its device
check and stream lookup are patched out so that the host logic and the complex64 numerics of whole tables can be checked
against the reference fixtures without a GPU.  The same assertions run on the CUDA build in tests/test_gpu_parity.py."""
import ctypes

import numpy as np
import pytest
import torch

from conftest import golden, rel_l2

from xlumina_b200 import _lib, four_f, ops
from test_elements import (check_element_kernels, directional, four_f_problem, sharp_focus_losses, sharp_focus_problem)


@pytest.fixture
def ops_on_emu(monkeypatch, emu):
    monkeypatch.setattr(_lib, "_lib", emu)
    monkeypatch.setattr(ops, "_require_device", lambda t: None)
    monkeypatch.setattr(ops, "_stream", lambda t: ctypes.c_void_p(0))
    monkeypatch.setattr(ops, "_stream_key", lambda t: ("cpu", 0))
    monkeypatch.setattr(ops, "_workspaces", {})
    monkeypatch.setattr(ops, "_z_cache", {})
    monkeypatch.setattr(ops, "_cpu_kernels", True)      # complex64 CPU planes go through the emulated element kernels (xl_el_*)
    yield


def test_rs_autograd_matches_reference(ops_on_emu):
    g = golden("rs_n32_zpos")
    x = g["x"]
    dx, k = float(x[1] - x[0]), 2 * np.pi / float(g["wavelength"])
    u = torch.tensor(g["field"].astype(np.complex64), requires_grad=True)
    z = torch.tensor([float(g["z"])], dtype=torch.float64, requires_grad=True)
    out = ops.rs_propagation(u, z, dx, dx, k)
    assert rel_l2(out.detach().numpy(), g["out"]) < 1e-5
    ct = torch.tensor(g["ct"].astype(np.complex64))
    # JAX cotangent convention of the fixture: vjp = J^T ct; torch's gradient of Re sum(conj(c) * out) is conj(J^T conj(c))
    loss = (torch.conj(torch.conj(ct)) * out).real.sum()        # Re sum(ct * out)
    gu, gz = torch.autograd.grad(loss, (u, z))
    assert rel_l2(np.conj(gu.numpy()), g["vjp_field"]) < 1e-5
    assert abs(float(gz) - float(g["vjp_z"])) < 1e-4 * abs(float(g["vjp_z"]))


@pytest.mark.parametrize("name,max_line", [("rs_n32_zpos", 32), ("rs_n32_zpos", 4096), ("rs_n32_zneg", 32), ("rs_n48_far", 32)])
def test_rs_large_grid_route_autograd_matches_reference(ops_on_emu, monkeypatch, emu, name, max_line):
    """ops.rs_propagation above FUSED_MAX_N (lowered here): the slab / split-line stage chain with one rank, differentiable in
    the field AND in z (slab.rs_slab_grad_z) -- against the reference-generated fixture (forward, field VJP, d/dz)."""
    from xlumina_b200 import slab
    g = golden(name)
    x = g["x"]
    dx, k = float(x[1] - x[0]), 2 * np.pi / float(g["wavelength"])
    monkeypatch.setattr(ops, "FUSED_MAX_N", 16)
    monkeypatch.setattr(slab, "_require_device", lambda t, lib: None)
    emu.xl_debug_set_max_line(max_line)          # 32: padded length 64 / 128 = 2 / 4 sub-lines through csrc/xl_long.cuh
    try:
        u = torch.tensor(g["field"].astype(np.complex64), requires_grad=True)
        z = torch.tensor([float(g["z"])], dtype=torch.float64, requires_grad=True)
        out = ops.rs_propagation(u, z, dx, dx, k)
        ct = torch.tensor(g["ct"].astype(np.complex64))
        loss = (ct * out).real.sum()                              # Re sum(ct * out): JAX cotangent convention of the fixture
        gu, gz = torch.autograd.grad(loss, (u, z))
    finally:
        emu.xl_debug_set_max_line(4096)
    assert rel_l2(out.detach().numpy(), g["out"]) < 1e-5
    if "vjp_field" in g:
        assert rel_l2(np.conj(gu.numpy()), g["vjp_field"]) < 1e-5
    assert abs(float(gz) - float(g["vjp_z"])) < 1e-4 * abs(float(g["vjp_z"]))


@pytest.mark.parametrize("name", ["vrs_n24", "vrs_n40_zneg"])
def test_vrs_large_grid_route_autograd_matches_reference(ops_on_emu, monkeypatch, emu, name):
    """vrs_propagation above FUSED_MAX_N (lowered here): Ez formed pointwise, three components through the split-line stage
    chain; forward, field VJP and d/dz (chain + d Ez/dz) against the reference-generated fixture."""
    from xlumina_b200 import slab
    g = golden(name)
    x, y = g["x"], g["y"]
    dx, dy, k = float(x[1] - x[0]), float(y[1] - y[0]), 2 * np.pi / float(g["wavelength"])
    monkeypatch.setattr(ops, "FUSED_MAX_N", 16)
    monkeypatch.setattr(slab, "_require_device", lambda t, lib: None)
    emu.xl_debug_set_max_line(32)
    try:
        e = torch.tensor(np.stack([g["Ex"], g["Ey"]]).astype(np.complex64), requires_grad=True)
        z = torch.tensor([float(g["z"])], dtype=torch.float64, requires_grad=True)
        out = ops.vrs_propagation(e, None, z, float(x[0]), float(y[0]), dx, dy, k)
        ct = torch.tensor(g["ct"].astype(np.complex64))
        ge, gz = torch.autograd.grad((ct * out).real.sum(), (e, z))
    finally:
        emu.xl_debug_set_max_line(4096)
    assert rel_l2(out.detach().numpy(), g["out"]) < 1e-5
    if "vjp_field" in g:
        assert rel_l2(np.conj(ge.numpy()), g["vjp_field"]) < 1e-5
    assert abs(float(gz) - float(g["vjp_z"])) < 1e-4 * abs(float(g["vjp_z"]))


def test_sharp_focus_table_complex64(ops_on_emu):
    g = golden("sharp_focus_n32")
    ls, params, fixed = sharp_focus_problem(g, "cpu", torch.complex64)
    inten, lv, l_soft, l_lin = sharp_focus_losses(g, ls, params, fixed)
    e_int = rel_l2(inten.detach().numpy(), g["intensities"])
    gl = torch.autograd.grad(l_lin, params, retain_graph=True, allow_unused=True)
    gm = torch.autograd.grad(l_soft, params, allow_unused=True)
    errs = {t: abs(directional(g, params, gl, t) - float(g["dlin_" + t])) / abs(float(g["dlin_" + t])) for t in ("all", "dist", "other")}
    e_soft = abs(directional(g, params, gm, "other") - float(g["dsoft_other"])) / abs(float(g["dsoft_other"]))
    print("sharp focus c64: intensities", e_int, "loss_vec", np.max(np.abs(lv.detach().numpy() / g["loss_vec"] - 1)), "dlin", errs, "dsoft", e_soft)
    assert e_int < 1e-4
    assert np.allclose(lv.detach().numpy(), g["loss_vec"], rtol=1e-3)
    assert max(errs.values()) < 1e-3 and e_soft < 1e-3


def test_four_f_table_complex64(ops_on_emu):
    g = golden("four_f_n32")
    src, params, masks, targets = four_f_problem(g, "cpu", torch.complex64)
    inten, _, _ = four_f.vector_dualSLM_4f_system(masks, src, params)
    e_int = rel_l2(inten.detach().numpy(), g["intensities"])
    loss = four_f.loss_dualSLM(params, masks, targets, src)
    grads = torch.autograd.grad(loss, params)
    errs = {t: abs(directional(g, params, grads, t, "v_%s_%d") - float(g["dloss_" + t])) / abs(float(g["dloss_" + t])) for t in ("dist", "phase")}
    print("4f c64: intensities", e_int, "loss", float(loss.detach()) / float(g["loss"]) - 1, "dloss", errs)
    assert e_int < 1e-4
    assert abs(float(loss.detach()) - float(g["loss"])) < 1e-4 * abs(float(g["loss"]))
    assert max(errs.values()) < 1e-3


def test_four_f_fused_elements_match_unfused_and_reference(ops_on_emu):
    """beam x mask and both SLM multiplies folded into the first pass, the intensity MSE and its cotangent
    into the last pass (xl_rs_fwd_fused / xl_rs_bwd_fused).  Same loss and gradients as the unfused table and the reference."""
    g = golden("four_f_n32")
    src, params, masks, targets = four_f_problem(g, "cpu", torch.complex64)
    loss_u = four_f.loss_dualSLM(params, masks, targets, src)
    grads_u = torch.autograd.grad(loss_u, params)
    loss_f = four_f.loss_dualSLM_fused(params, masks, targets, src)
    grads_f = torch.autograd.grad(loss_f, params)
    assert abs(float(loss_f.detach()) - float(g["loss"])) < 1e-4 * abs(float(g["loss"]))
    assert abs(float(loss_f.detach()) - float(loss_u.detach())) < 1e-5 * abs(float(loss_u.detach()))
    for i in (3, 4):     # phase masks
        assert rel_l2(grads_f[i].numpy(), grads_u[i].numpy()) < 1e-4
    errs = {t: abs(directional(g, params, grads_f, t, "v_%s_%d") - float(g["dloss_" + t])) / abs(float(g["dloss_" + t])) for t in ("dist", "phase")}
    print("4f fused c64: loss", float(loss_f.detach()) / float(g["loss"]) - 1, "dloss", errs,
          "dz fused", [float(grads_f[i]) for i in range(3)], "dz unfused", [float(grads_u[i]) for i in range(3)])
    assert max(errs.values()) < 1e-3


def test_rs_fused_pieces(ops_on_emu):
    """Each fused piece against plain torch arithmetic around the unfused operator: complex and float32 inputs, shared
    modulation, detection; gradients with respect to field, z and the modulation plane; the d/dz-only backward."""
    rng = np.random.default_rng(5)
    N, F = 24, 3
    x = np.linspace(-400.0, 400.0, N)
    dx, k = float(x[1] - x[0]), 2 * np.pi / 0.6328
    z = torch.tensor([31000.0], dtype=torch.float64, requires_grad=True)
    u = torch.tensor((rng.standard_normal((F, N, N)) + 1j * rng.standard_normal((F, N, N))).astype(np.complex64), requires_grad=True)
    m = torch.tensor((rng.standard_normal((N, N)) + 1j * rng.standard_normal((N, N))).astype(np.complex64), requires_grad=True)
    ct = torch.tensor((rng.standard_normal((F, N, N)) + 1j * rng.standard_normal((F, N, N))).astype(np.complex64))
    tgt = torch.tensor(rng.uniform(0, 2, (F, N, N)).astype(np.float32))
    w = torch.tensor(rng.uniform(0.5, 1.5, F))
    # (a) modulation, complex cotangent
    ref = ops.rs_propagation(u * m[None], z, dx, dx, k)
    out = ops.rs_propagation_fused(u, z, dx, dx, k, mod=m)
    assert rel_l2(out.detach().numpy(), ref.detach().numpy()) < 1e-6
    gr = torch.autograd.grad((ct.conj() * ref).real.sum(), (u, z, m))
    gf = torch.autograd.grad((ct.conj() * out).real.sum(), (u, z, m))
    assert rel_l2(gf[0].numpy(), gr[0].numpy()) < 1e-5 and rel_l2(gf[2].numpy(), gr[2].numpy()) < 1e-5
    assert abs(float(gf[1]) - float(gr[1])) < 1e-4 * abs(float(gr[1]))
    # (b) detection with modulation
    def mse_ref(uu, zz, mm):
        o = ops.rs_propagation(uu * mm[None], zz, dx, dx, k)
        return (((o.real ** 2 + o.imag ** 2 - tgt) ** 2).sum(dim=(-2, -1)) / (N * N) * w).sum()
    lr = mse_ref(u, z, m)
    lf = (ops.rs_propagation_fused(u, z, dx, dx, k, mod=m, target=tgt) * w).sum()
    assert abs(float(lf) - float(lr)) < 1e-5 * abs(float(lr))
    gr = torch.autograd.grad(lr, (u, z, m))
    gf = torch.autograd.grad(lf, (u, z, m))
    assert rel_l2(gf[0].numpy(), gr[0].numpy()) < 1e-4 and rel_l2(gf[2].numpy(), gr[2].numpy()) < 1e-4
    # (c) float32 object masks under a constant beam: only z needs a gradient (d/dz-only backward, no inverse transforms)
    masks = torch.tensor((rng.uniform(0, 1, (F, N, N)) > 0.5).astype(np.float32))
    beam = m.detach()
    ref = ops.rs_propagation(masks.to(torch.complex64) * beam[None], z, dx, dx, k)
    out = ops.rs_propagation_fused(masks, z, dx, dx, k, mod=beam)
    assert rel_l2(out.detach().numpy(), ref.detach().numpy()) < 1e-6
    (gzr,) = torch.autograd.grad((ct.conj() * ref).real.sum(), (z,))
    (gzf,) = torch.autograd.grad((ct.conj() * out).real.sum(), (z,))
    assert abs(float(gzf) - float(gzr)) < 1e-4 * abs(float(gzr))
    # (d) no gradient with respect to z: forward-operator backward with the detection seed
    z0 = z.detach()
    gr = torch.autograd.grad(mse_ref(u, z0, m), (u, m))
    gf = torch.autograd.grad((ops.rs_propagation_fused(u, z0, dx, dx, k, mod=m, target=tgt) * w).sum(), (u, m))
    assert rel_l2(gf[0].numpy(), gr[0].numpy()) < 1e-4 and rel_l2(gf[1].numpy(), gr[1].numpy()) < 1e-4


def test_leading_batch_axes_and_per_item_distances(ops_on_emu):
    """What the reference reaches with vmap (a table vmapped over masks / noisy distances): batches are independent calls,
    a shared z shares the transfer function, a z per item is honoured, gradients flow to every item and every distance."""
    rng = np.random.default_rng(11)
    N, M = 16, 12
    x = np.linspace(-300.0, 300.0, N)
    xo = np.linspace(-40.0, 40.0, M)
    dx, lam = float(x[1] - x[0]), 0.6328
    k = 2 * np.pi / lam
    c = lambda *s: torch.tensor((rng.standard_normal(s) + 1j * rng.standard_normal(s)).astype(np.complex64))   # noqa: E731
    u = c(3, N, N).requires_grad_(True)
    zs = torch.tensor([2000.0, 2500.0, -1800.0], dtype=torch.float64, requires_grad=True)
    ct = c(3, N, N)
    out = ops.rs_propagation(u, zs, dx, dx, k)
    (out * ct).real.sum().backward()
    for i in range(3):
        ui = u.detach()[i].clone().requires_grad_(True)
        zi = zs.detach()[i:i + 1].clone().requires_grad_(True)
        oi = ops.rs_propagation(ui, zi, dx, dx, k)
        (oi * ct[i]).real.sum().backward()
        assert torch.equal(oi.detach(), out.detach()[i]) and torch.equal(ui.grad, u.grad[i])
        assert float(zi.grad) == float(zs.grad[i])
    with pytest.raises(ValueError):
        ops.rs_propagation(u, zs[:2], dx, dx, k)

    ex, ey = c(2, N, N).requires_grad_(True), c(2, N, N)
    for z in (3000.0, torch.tensor([3000.0, 3300.0], dtype=torch.float64)):
        ob = ops.vrs_propagation(ex, ey, z, float(x[0]), float(x[0]), dx, dx, k)
        assert ob.shape == (2, 3, N, N)
        for i in range(2):
            zi = z if isinstance(z, float) else z[i:i + 1]
            assert torch.equal(ob[i], ops.vrs_propagation(ex[i], ey[i], zi, float(x[0]), float(x[0]), dx, dx, k))
    ob = ops.vrs_propagation(torch.stack([ex, ey], dim=1), None, 3000.0, float(x[0]), float(x[0]), dx, dx, k)
    assert torch.equal(ob[1], ops.vrs_propagation(ex[1], ey[1], 3000.0, float(x[0]), float(x[0]), dx, dx, k))
    g, = torch.autograd.grad((ob * c(2, 3, N, N)).real.sum(), ex)
    assert g.shape == ex.shape and float(g.abs().min()) > 0
    # one batched call (xl_vrs_*_batch) with a distance per item: field and distance gradients equal the single-item calls
    zv = torch.tensor([3000.0, 3300.0], dtype=torch.float64, requires_grad=True)
    eyg = ey.clone().requires_grad_(True)
    ctv = c(2, 3, N, N)
    ob = ops.vrs_propagation(ex, eyg, zv, float(x[0]), float(x[0]), dx, dx, k)
    gx, gy, gzv = torch.autograd.grad((ob * ctv).real.sum(), (ex, eyg, zv))
    for i in range(2):
        exi, eyi = ex.detach()[i].clone().requires_grad_(True), ey[i].clone().requires_grad_(True)
        zi = zv.detach()[i:i + 1].clone().requires_grad_(True)
        oi = ops.vrs_propagation(exi, eyi, zi, float(x[0]), float(x[0]), dx, dx, k)
        gxi, gyi, gzi = torch.autograd.grad((oi * ctv[i]).real.sum(), (exi, eyi, zi))
        assert torch.equal(gxi, gx[i]) and torch.equal(gyi, gy[i]) and float(gzi) == float(gzv[i])

    cb = ops.czt(u.detach(), torch.tensor([9000.0, 9500.0, 8000.0], dtype=torch.float64), lam, x, x, xo, xo)
    assert cb.shape == (3, M, M) and torch.equal(cb[2], ops.czt(u.detach()[2], 8000.0, lam, x, x, xo, xo))
    vb = ops.vczt(ex.detach(), ey, 9000.0, lam, x, x, xo, xo)
    assert vb.shape == (2, 3, M, M) and torch.equal(vb[1], ops.vczt(ex.detach()[1], ey[1], 9000.0, lam, x, x, xo, xo))
    hb = ops.highna_focus(ex.detach(), ey, 250.0, 400.0, lam, x, x, xo, xo)
    assert hb.shape == (2, 3, M, M) and torch.equal(hb[0], ops.highna_focus(ex.detach()[0], ey[0], 250.0, 400.0, lam, x, x, xo, xo))


def test_reference_propagating_element_tests(ops_on_emu):
    """The reference's own tests of the elements that propagate (tests/test_optical_elements.py:128-136, 165-171:
    VCZT_objective_lens and building_block shapes) at resolution 64 instead of 512."""
    import math
    import xlumina_b200 as xb
    n, wavelength = 64, 633e-3
    x = np.linspace(-1500, 1500, n)
    light = xb.PolarizedLightSource(x, x, wavelength, device="cpu")
    light.gaussian_beam(w0=(1200, 1200), jones_vector=(1, 0))
    radius = 3.6 * 1e3 / 2
    out = xb.VCZT_objective_lens(light, radius, radius / 0.9, x, x)
    assert out.Ex.shape == out.Ey.shape == out.Ez.shape == (n, n)
    out = xb.building_block(light, torch.zeros((n, n)), torch.zeros((n, n)), 1000, math.pi / 2, math.pi / 4)
    assert out.Ex.shape == out.Ey.shape == out.Ez.shape == (n, n)
    assert bool(torch.isfinite(out.Ex.abs()).all())


# --------------------------------------------------------------------------- randomised: torch-convention gradients vs autograd
from hypothesis import given, settings, strategies as st, HealthCheck   # noqa: E402
from oracle import oracle_torch as ot   # noqa: E402


def _crand(rng, *s):
    return (rng.standard_normal(s) + 1j * rng.standard_normal(s))


def _grads(loss, leaves):
    return [g.detach().numpy() for g in torch.autograd.grad(loss, leaves)]


@settings(max_examples=12, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
@given(N=st.integers(4, 28), half=st.floats(100.0, 2000.0), zmag=st.floats(500.0, 1e5), neg=st.booleans(),
       vect=st.booleans(), seed=st.integers(0, 2 ** 16))
def test_rs_vrs_autograd_random(ops_on_emu, N, half, zmag, neg, vect, seed):
    """Gradients delivered to torch (field: conj-of-JAX convention; z) equal autograd through the complex128 oracle for a
    generic complex-weighted loss Re sum(w * out)."""
    rng = np.random.default_rng(seed)
    x = np.linspace(-half, half, N)
    dx, lam = float(x[1] - x[0]), 0.6328
    k, zv = 2 * np.pi / lam, (-zmag if neg else zmag)
    nf = 2 if vect else 1
    u0, w0 = _crand(rng, nf, N, N), _crand(rng, 3 if vect else 1, N, N)

    def run(cdtype, fwd):
        u = torch.tensor(u0, dtype=cdtype, requires_grad=True)
        z = torch.tensor([zv], dtype=torch.float64, requires_grad=True)
        out = fwd(u, z)
        loss = (torch.tensor(w0, dtype=cdtype) * out).real.sum()
        return [out.detach().numpy()] + _grads(loss, (u, z))
    if vect:
        got = run(torch.complex64, lambda u, z: ops.vrs_propagation(u[0], u[1], z, float(x[0]), float(x[0]), dx, dx, k))
        ref = run(torch.complex128, lambda u, z: ot.VRS_propagation(u[0], u[1], x, x, lam, z))
    else:
        got = run(torch.complex64, lambda u, z: ops.rs_propagation(u, z, dx, dx, k))
        ref = run(torch.complex128, lambda u, z: ot.RS_propagation(u[0], x, x, lam, z)[None])
    assert rel_l2(got[0], ref[0]) < 2e-5 and rel_l2(got[1], ref[1]) < 2e-5
    scale = k * np.linalg.norm(w0) * np.linalg.norm(ref[0])       # natural size of d/dz (the i*k*out part)
    assert abs(got[2].item() - ref[2].item()) < 2e-5 * scale + 1e-4 * abs(ref[2].item())


@settings(max_examples=12, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
@given(N=st.integers(4, 28), Mx=st.integers(2, 30), My=st.integers(2, 30), kind=st.sampled_from(["czt", "vczt", "highna"]),
       seed=st.integers(0, 2 ** 16))
def test_czt_family_autograd_random(ops_on_emu, emu, N, Mx, My, kind, seed):
    if emu.xl_czt_padded_length(N, Mx) == 0 or emu.xl_czt_padded_length(N, My) == 0:
        return                                                     # m+M-1 a power of two: the reference raises as well
    rng = np.random.default_rng(seed)
    N += N & 1 if kind == "highna" else 0                          # odd N: the reference's 0/0 at the origin pixel (NaN)
    if emu.xl_czt_padded_length(N, Mx) == 0 or emu.xl_czt_padded_length(N, My) == 0:
        return
    x = np.linspace(-600.0, 600.0, N)
    xo, yo = np.linspace(-50.0, 40.0, Mx), np.linspace(-30.0, 50.0, My)
    lam, z = 0.6328, 20000.0
    nf = 1 if kind == "czt" else 2
    u0, w0 = _crand(rng, nf, N, N), _crand(rng, 1 if kind == "czt" else 3, My, Mx)

    def run(cdtype, fwd):
        u = torch.tensor(u0, dtype=cdtype, requires_grad=True)
        out = fwd(u)
        loss = (torch.tensor(w0, dtype=cdtype) * out).real.sum()
        return [out.detach().numpy()] + _grads(loss, (u,))
    if kind == "czt":
        got = run(torch.complex64, lambda u: ops.czt(u[0], z, lam, x, x, xo, yo)[None])
        ref = run(torch.complex128, lambda u: ot.CZT(u[0], x, x, lam, z, xo, yo)[None])
    elif kind == "vczt":
        got = run(torch.complex64, lambda u: ops.vczt(u[0], u[1], z, lam, x, x, xo, yo))
        ref = run(torch.complex128, lambda u: ot.VCZT(u[0], u[1], x, x, lam, z, xo, yo))
    else:
        got = run(torch.complex64, lambda u: ops.highna_focus(u[0], u[1], 500.0, 700.0, lam, x, x, xo, yo))
        ref = run(torch.complex128, lambda u: ot.VCZT_objective_lens(u[0], u[1], x, x, lam, 500.0, 700.0, xo, yo))
    assert rel_l2(got[0], ref[0]) < 2e-5 and rel_l2(got[1], ref[1]) < 2e-5


def test_separate_planes_need_no_stacking(ops_on_emu):
    """The vectorial entry points take Ex and Ey where they live (include/xlprop.h: `ex`, `ey` pointers): planes that are NOT
    adjacent in memory -- what the elements of an optical table produce -- give the same results and gradients as a stacked
    (2,N,N) pair, without a torch.stack copy."""
    rng = np.random.default_rng(21)
    N, lam = 16, 0.6328
    x = np.linspace(-300.0, 300.0, N)
    xo = np.linspace(-40.0, 40.0, 12)
    dx, k = float(x[1] - x[0]), 2 * np.pi / lam
    pool = torch.tensor((rng.standard_normal((5, N, N)) + 1j * rng.standard_normal((5, N, N))).astype(np.complex64))
    calls = {
        "vrs": lambda a, b: ops.vrs_propagation(a, b, torch.tensor([9000.0], dtype=torch.float64, requires_grad=True), x[0], x[0], dx, dx, k),
        "vczt": lambda a, b: ops.vczt(a, b, 9000.0, lam, x, x, xo, xo),
        "highna": lambda a, b: ops.highna_focus(a, b, 500.0, 700.0, lam, x, x, xo, xo),
    }
    for name, fn in calls.items():
        for (ia, ib) in ((3, 0), (0, 4)):          # Ey before Ex in memory, and three planes apart
            ex = pool[ia].clone().requires_grad_(True)
            ey = pool[ib].clone().requires_grad_(True)
            out = fn(ex, ey)
            ct = torch.tensor((rng.standard_normal(tuple(out.shape)) + 1j * rng.standard_normal(tuple(out.shape))).astype(np.complex64))
            gx, gy = torch.autograd.grad((ct.conj() * out).real.sum(), (ex, ey))
            st = torch.stack([pool[ia], pool[ib]]).requires_grad_(True)
            out2 = fn(st, None)
            (gs,) = torch.autograd.grad((ct.conj() * out2).real.sum(), (st,))
            assert rel_l2(out.detach().numpy(), out2.detach().numpy()) < 1e-6, name
            assert rel_l2(torch.stack([gx, gy]).numpy(), gs.numpy()) < 1e-6, name


def _fd_dz(fn, z, eps):
    """4th-order central difference of a real functional of z."""
    return (8 * (fn(z + eps) - fn(z - eps)) - (fn(z + 2 * eps) - fn(z - 2 * eps))) / (12 * eps)


@pytest.mark.parametrize("vect", [0, 1])
@pytest.mark.parametrize("case", ["same_grid", "roi", "zneg"])
def test_czt_distance_gradient(ops_on_emu, vect, case):
    """d/dz of CZT / VCZT (z enters F, F0, the constant and, through Dm, every chirp: wave_optics.py:322,
    340-355, 393-403) against a central difference of the complex128 NumPy oracle, with a random cotangent; the field gradient
    of the same backward call is checked too."""
    from oracle import oracle_np as o
    rng = np.random.default_rng(3 + vect)
    N, lam = 24, 0.6328
    x = np.linspace(-300.0, 300.0, N)
    y = np.linspace(-280.0, 280.0, N) if case != "same_grid" else x
    xo, yo = (x, y) if case == "same_grid" else (np.linspace(-40.0, 55.0, 30), np.linspace(-35.0, 50.0, 28))
    z0 = -9000.0 if case == "zneg" else 9000.0
    shp = (2, N, N) if vect else (N, N)
    u = (rng.standard_normal(shp) + 1j * rng.standard_normal(shp)).astype(np.complex64)
    oshp = (3, len(yo), len(xo)) if vect else (len(yo), len(xo))
    ct = (rng.standard_normal(oshp) + 1j * rng.standard_normal(oshp)).astype(np.complex64)

    def L_ref(z):
        uu = u.astype(np.complex128)
        out = o.VCZT(uu[0], uu[1], x, y, lam, z, xo, yo) if vect else o.CZT(uu, x, y, lam, z, xo, yo)
        return float(np.real(np.sum(np.conj(ct) * out)))
    gz_ref = _fd_dz(L_ref, z0, 2e-5)
    ut = torch.tensor(u, requires_grad=True)
    zt = torch.tensor([z0], dtype=torch.float64, requires_grad=True)
    out = ops.vczt(ut, None, zt, lam, x, y, xo, yo) if vect else ops.czt(ut, zt, lam, x, y, xo, yo)
    loss = (torch.tensor(ct).conj() * out).real.sum()
    gu, gz = torch.autograd.grad(loss, (ut, zt))
    assert abs(float(loss) - L_ref(z0)) < 1e-5 * abs(L_ref(z0)) + 1e-5 * float(np.linalg.norm(ct)) * float(out.detach().abs().pow(2).sum().sqrt())
    scale = float(np.linalg.norm(ct)) * float(out.detach().abs().pow(2).sum().sqrt())   # size of a generic d/dz is k * scale
    print(case, "vect" if vect else "scalar", "gz", float(gz), "ref", gz_ref, "rel", abs(float(gz) - gz_ref) / abs(gz_ref))
    assert abs(float(gz) - gz_ref) < 1e-4 * max(abs(gz_ref), 1e-2 * scale)
    # the field gradient of the same call equals the one of a call without d/dz
    (gu0,) = torch.autograd.grad((torch.tensor(ct).conj() * (ops.vczt(ut, None, z0, lam, x, y, xo, yo) if vect else ops.czt(ut, z0, lam, x, y, xo, yo))).real.sum(), (ut,))
    assert rel_l2(gu.numpy(), gu0.numpy()) < 1e-6


@pytest.mark.parametrize("kind", ["sslm", "lcd", "bs"])
def test_element_kernels_match_torch_arithmetic(ops_on_emu, kind):
    """xl_el_sslm / xl_el_lcd / xl_el_bs on the host-emulated kernels (see test_elements.check_element_kernels)."""
    check_element_kernels(kind, "cpu")
