"""Host-side helpers of the hot path.  Mirrors xlumina/toolbox.py:49-61 (space) and :74-96 (is_conserving_energy)."""
import numpy as np
import torch


def space(x_total, num_pix):
    """Simulation axes: num_pix samples on [-x_total, x_total] (microns).  Reference: toolbox.py:49-61."""
    x = np.linspace(-x_total, x_total, num_pix, dtype=np.float64)
    y = np.linspace(-x_total, x_total, num_pix, dtype=np.float64)
    return x, y


def is_conserving_energy(light_source, propagated_light):
    """I_propagated / I_source.  Reference: toolbox.py:74-96."""
    if light_source.info in ('Wave optics light', 'Wave optics light source'):
        i0 = torch.sum(torch.abs(light_source.field ** 2))
        i1 = torch.sum(torch.abs(propagated_light.field ** 2))
    else:
        i0 = sum(torch.sum(torch.abs(c ** 2)) for c in (light_source.Ex, light_source.Ey, light_source.Ez))
        i1 = sum(torch.sum(torch.abs(c ** 2)) for c in (propagated_light.Ex, propagated_light.Ey, propagated_light.Ez))
    return i1 / i0


def softmin(args, beta=90):
    """Differentiable min: -logsumexp(-beta * args) / beta.  Reference: toolbox.py:98-103."""
    return -torch.logsumexp(-beta * args.reshape(-1), dim=0) / beta
