"""
Detector-side loss functions of the reference's optical tables: closed-form torch reductions over the
propagated planes, differentiable through autograd.  Mirror of xlumina/loss_functions.py (reference lines per docstring).
"""
import torch


def small_area_hybrid(detected_intensity):
    """Pixels above 70 % of the peak of the normalised intensity, divided by the intensity fraction they hold; also
    accepts a stack (..., M, M) and then returns one value per plane.  loss_functions.py:24-40 (the mask is piecewise
    constant: no gradient flows through it, as with jnp.where on a comparison)."""
    epsilon = 0.7
    eps = 1e-08
    total = detected_intensity.sum(dim=(-2, -1), keepdim=True)
    I = detected_intensity / (total + eps)
    mask = (I > epsilon * I.amax(dim=(-2, -1), keepdim=True)).to(I.dtype)
    return mask.sum(dim=(-2, -1)) / ((mask * I).sum(dim=(-2, -1)) + eps)


def vectorized_loss_hybrid(detected_intensities):
    """small_area_hybrid over the leading (detector) axis.  loss_functions.py:42-50."""
    return small_area_hybrid(detected_intensities)


def MSE_Amplitude(input_light, target_light):
    """loss_functions.py:111-123."""
    num_pix = input_light.shape[-2] * input_light.shape[-1]
    return torch.sum((torch.abs(input_light) - torch.abs(target_light)) ** 2, dim=(-2, -1)) / num_pix


def MSE_Phase(input_light, target_light):
    """loss_functions.py:125-137."""
    num_pix = input_light.shape[-2] * input_light.shape[-1]
    return torch.sum((torch.angle(input_light) - torch.angle(target_light)) ** 2, dim=(-2, -1)) / num_pix


def _intensity(t):
    return t.real ** 2 + t.imag ** 2 if torch.is_complex(t) else t ** 2


def MSE_Intensity(input_light, target_light):
    """sum((|a|^2 - |b|^2)^2) / num_pix for field planes a, b (a leading batch axis is allowed).  loss_functions.py:139-151."""
    num_pix = input_light.shape[-2] * input_light.shape[-1]
    return torch.sum((_intensity(input_light) - _intensity(target_light)) ** 2, dim=(-2, -1)) / num_pix


def mean_batch_MSE_Intensity(optimized, target):
    """(mean over the batch, per-sample values) of MSE_Intensity.  loss_functions.py:52-59."""
    mse = MSE_Intensity(optimized, target)
    return mse.mean(), mse


def _stack_components(light):
    return torch.stack([light.Ex, light.Ey, light.Ez])


def vMSE_Amplitude(input_light, target_light):
    """[MSEx, MSEy, MSEz] in amplitude.  loss_functions.py:61-75."""
    return MSE_Amplitude(_stack_components(input_light), _stack_components(target_light))


def vMSE_Phase(input_light, target_light):
    """[MSEx, MSEy, MSEz] in phase.  loss_functions.py:77-91."""
    return MSE_Phase(_stack_components(input_light), _stack_components(target_light))


def vMSE_Intensity(input_light, target_light):
    """[MSEx, MSEy, MSEz] in intensity.  loss_functions.py:93-107."""
    return MSE_Intensity(_stack_components(input_light), _stack_components(target_light))
