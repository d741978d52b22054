"""
High-NA objective focusing on libxlprop.so.  Mirror of xlumina/optical_elements.py:515-672: the lens factor
(high_NA_objective_lens + _high_NA_objective_lens_, :515-594) is fused into the load of the first Bluestein pass and the
constant -i sin^2(theta_max)/(f lambda) (:627) into the store of the second, so the (3,N,N) lens field never exists in HBM.
"""
import time

from . import ops
from .vectorized_optics import VectorizedLight
from . import wave_optics as _wo


def build_high_NA_VCZT_grid(f, r, wavelength, xin, xout, yout):
    """optical_elements.py:640-672 (host scalars)."""
    nx, ny, Din = len(xout), len(yout), len(xin)
    Dm = f * wavelength * (Din - 1) / (2 * r)
    return nx, ny, Dm, yout[0] + Dm / 2, yout[-1] + Dm / 2, xout[0] + Dm / 2, xout[-1] + Dm / 2


def VCZT_objective_lens(input_field, r, f, xout, yout):
    """Focus `input_field` (VectorizedLight) with an objective of radius r and focal length f (microns); returns the
    VectorizedLight in the focal plane sampled at (xout, yout).  Reference: optical_elements.py:600-638."""
    tic = time.perf_counter()
    E = ops.highna_focus(input_field.Ex, input_field.Ey, r, f, input_field.wavelength, input_field.x, input_field.y, xout, yout)
    out = VectorizedLight(xout, yout, input_field.wavelength, input_field.device)
    out.Ex, out.Ey, out.Ez = E[0], E[1], E[2]
    if _wo.VERBOSE:
        print(f"Time taken to perform one VCZT propagation through objective lens (in seconds):  {(time.perf_counter() - tic):.4f}")
    return out
