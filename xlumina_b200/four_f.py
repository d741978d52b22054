"""
The 4f-system optical table of BASELINE.json config 4 (two phase-only SLMs between three RS propagations) and its batch
loss.  Mirror of experiments/four_f_optical_table.py:36-141; the reference reads the light source, grid and resolution from
module globals, here they are arguments, and the whole batch of masks goes through each propagation as ONE library call
sharing the transfer function (the reference vmaps the per-sample function, :98).
"""
import math

import torch

from . import ops

cm = 1e4
OFFSET = 1.2     # from get_RS_minimum_z() at the reference's resolution, four_f_optical_table.py:57


def _distance(p):
    """(|p| * 100 + offset) cm in microns, in float64 (distances reach 1e6 um), four_f_optical_table.py:65,73,81."""
    return (torch.abs(p.to(torch.float64)) * 100 + OFFSET) * cm


def _slm_phasor(p):
    ph = p * (2 * math.pi) - math.pi
    return torch.polar(torch.ones_like(ph), ph)


def vector_dualSLM_4f_system(input_masks, input_light, parameters):
    """Detected intensities (B, N, N) for a batch of input masks (B, N, N) applied to `input_light` (a LightSource):
    mask -> RS(z0) -> SLM(phase1) -> RS(z1) -> SLM(phase2) -> RS(z2) -> |.|^2; parameters = [p_z0, p_z1, p_z2, p_phase1, p_phase2]
    in (0, 1).  four_f_optical_table.py:36-100.  Returns (intensities, slm_1, slm_2)."""
    x = input_light.x
    dx, k = float(x[1] - x[0]), input_light.k
    f = input_light.field[None] * input_masks
    f = ops.rs_propagation(f, _distance(parameters[0]), dx, dx, k)
    slm_1 = _slm_phasor(parameters[3])
    f = ops.rs_propagation(f * slm_1[None], _distance(parameters[1]), dx, dx, k)
    slm_2 = _slm_phasor(parameters[4])
    f = ops.rs_propagation(f * slm_2[None], _distance(parameters[2]), dx, dx, k)
    return f.real ** 2 + f.imag ** 2, slm_1, slm_2


def loss_dualSLM_fused(parameters, input_masks, target_intensities, input_light):
    """loss_dualSLM with the pointwise elements folded into the propagation kernels (ops.rs_propagation_fused): the beam is the
    shared modulation of the real object masks, each SLM the shared modulation of the next propagation, and the intensity MSE
    is accumulated by the last inverse pass -- no element, intensity or cotangent plane is materialised per sample.
    input_masks (B,N,N) real (float32 or complex with zero imaginary part), target_intensities (B,N,N).  Same value and
    gradients as loss_dualSLM (four_f_optical_table.py:36-141)."""
    x = input_light.x
    dx, k = float(x[1] - x[0]), input_light.k
    masks = input_masks.real if input_masks.is_complex() else input_masks
    zs = [_distance(parameters[i]) for i in range(3)]
    # the three transfer functions and their z-derivatives in one launch pair (they do not depend on the batch)
    pre = ops.rs_transfer_pairs(zs, masks.shape[-1], dx, dx, k, masks.device)
    # one path per sample ending in an intensity detector: the loss is blind to the global phase of every plane
    f = ops.rs_propagation_fused(masks.to(torch.float32), zs[0], dx, dx, k, mod=input_light.field, phase_blind=True, pre=pre[0])
    f = ops.rs_propagation_fused(f, zs[1], dx, dx, k, mod=_slm_phasor(parameters[3]), phase_blind=True, pre=pre[1])
    mse = ops.rs_propagation_fused(f, zs[2], dx, dx, k, mod=_slm_phasor(parameters[4]), target=target_intensities, pre=pre[2])
    return mse.mean()


def batch_dualSLM_4f(input_mask, input_light, parameters):
    """One sample of vector_dualSLM_4f_system.  four_f_optical_table.py:36-83."""
    inten, slm_1, slm_2 = vector_dualSLM_4f_system(input_mask[None], input_light, parameters)
    return inten[0], slm_1, slm_2


def MSE_Intensity(input_light, target_light):
    """MSE between two INTENSITY planes.  four_f_optical_table.py:129-141 (loss_functions.MSE_Intensity takes fields)."""
    num_pix = input_light.shape[-2] * input_light.shape[-1]
    return torch.sum((input_light - target_light) ** 2, dim=(-2, -1)) / num_pix


def mean_batch_MSE_Intensity(optimized, target):
    """four_f_optical_table.py:120-127."""
    mse = MSE_Intensity(optimized, target)
    return mse.mean(), mse


def loss_dualSLM(parameters, input_masks, target_intensities, input_light):
    """Mean over the batch of the intensity MSE against the targets.  four_f_optical_table.py:103-118."""
    optimized, _, _ = vector_dualSLM_4f_system(input_masks, input_light, parameters)
    return mean_batch_MSE_Intensity(optimized, target_intensities)[0]
