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
from test_elements import (directional, four_f_problem, sharp_focus_losses, sharp_focus_problem)


@pytest.fixture
def ops_on_emu(monkeypatch, emu):
    monkeypatch.setattr(_lib, "_lib", emu)
    monkeypatch.setattr(ops, "_require_device", lambda t: None)
    monkeypatch.setattr(ops, "_stream", lambda t: ctypes.c_void_p(0))
    monkeypatch.setattr(ops, "_stream_key", lambda t: ("cpu", 0))
    monkeypatch.setattr(ops, "_workspaces", {})
    monkeypatch.setattr(ops, "_z_cache", {})
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
