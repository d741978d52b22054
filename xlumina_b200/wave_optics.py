"""
Scalar fields: ScalarLight / LightSource with RS_propagation and CZT running on libxlprop.so.

Mirror of xlumina/wave_optics.py (reference line numbers in each docstring).  Differences that matter to a user switching
over: arrays are torch CUDA tensors (complex64 by default), coordinate grids must be uniform (they are regenerated
analytically inside the kernels instead of being materialised as Xext/Yext/X/Y, wave_optics.py:265-279,314), and the
per-call "Time taken" print of the reference is off unless xlumina_b200.wave_optics.VERBOSE is set.
"""
import math
import time

import numpy as np
import torch

from . import ops

VERBOSE = False
DEFAULT_DEVICE = "cuda"


def _quality_factor(x, y, wavelength, z):
    """Sampling quality factor, wave_optics.py:188-191 (same formula at vectorized_optics.py:265-270)."""
    dx = float(x[1] - x[0])
    dy = float(y[1] - y[0])
    dr_real = math.sqrt(dx ** 2 + dy ** 2)
    rmax = math.sqrt(float((np.asarray(x) ** 2).max()) + float((np.asarray(y) ** 2).max()))
    if isinstance(z, torch.Tensor):
        dr_ideal = torch.sqrt(wavelength ** 2 + rmax ** 2 + 2 * wavelength * torch.sqrt(rmax ** 2 + z.detach() ** 2)) - rmax
    else:
        dr_ideal = math.sqrt(wavelength ** 2 + rmax ** 2 + 2 * wavelength * math.sqrt(rmax ** 2 + z ** 2)) - rmax
    return dr_ideal / dr_real


def build_grid(x, y):
    """wave_optics.py:265-279 without the (2N-1)^2 meshgrids: returns nx, ny, dx, dy (Xext/Yext are implicit)."""
    return len(x), len(y), float(x[1] - x[0]), float(y[1] - y[0])


def RS_propagation_jit(input_field, z, nx, ny, dx, dy, k):
    """Seam function, wave_optics.py:281-289 (Xext/Yext dropped: generated analytically on the device)."""
    if nx != ny:
        raise ValueError("square grids only")
    return ops.rs_propagation(input_field, z, dx, dy, k)


def build_CZT_grid(z, wavelength, xin, yin, xout, yout):
    """wave_optics.py:299-331 (host scalars only; Xout/Yout are implicit)."""
    nx, ny = len(xout), len(yout)
    dx = float(xin[1] - xin[0])
    dy = float(yin[1] - yin[0])
    Dm = wavelength * z / dx
    return nx, ny, dx, dy, Dm, yout[0] + Dm / 2, yout[-1] + Dm / 2, xout[0] + Dm / 2, xout[-1] + Dm / 2


def CZT_jit(field, z, wavelength, x, y, xout, yout):
    """Seam function, wave_optics.py:333-357: F0 * Bluestein_x(Bluestein_y(field * F)) * z*dx*dy*lambda."""
    return ops.czt(field, z, wavelength, x, y, xout, yout)


class ScalarLight:
    """Scalar complex amplitude on a square grid.  Reference: wave_optics.py:43-54."""

    def __init__(self, x, y, wavelength, device=None, _alloc=True):
        self.x = x
        self.y = y
        self.wavelength = wavelength
        self.k = 2 * math.pi / wavelength
        self.n = 1
        self.device = torch.device(device or DEFAULT_DEVICE)
        # _alloc=False: the caller assigns the field right away (skips one N^2 memset per element of a table)
        self.field = torch.zeros((len(y), len(x)), dtype=torch.complex64, device=self.device) if _alloc else None
        self.info = 'Wave optics light'

    @property
    def X(self):
        return torch.as_tensor(np.meshgrid(self.x, self.y)[0], device=self.device)

    @property
    def Y(self):
        return torch.as_tensor(np.meshgrid(self.x, self.y)[1], device=self.device)

    def RS_propagation(self, z):
        """Rayleigh-Sommerfeld propagation over z (microns); returns (ScalarLight, quality_factor).  wave_optics.py:173-196."""
        tic = time.perf_counter()
        nx, ny, dx, dy = build_grid(self.x, self.y)
        quality_factor = _quality_factor(self.x, self.y, self.wavelength, z)
        out = ScalarLight(self.x, self.y, self.wavelength, self.device, _alloc=False)
        out.field = RS_propagation_jit(self.field, z, nx, ny, dx, dy, self.k)
        if VERBOSE:
            print(f"Time taken to perform one RS propagation (in seconds): {(time.perf_counter() - tic):.4f}")
        return out, quality_factor

    def get_RS_minimum_z(self, n=1, quality_factor=1):
        """wave_optics.py:198-231 (diagnostic; host math)."""
        range_x = self.x[-1] - self.x[0]
        range_y = self.y[-1] - self.y[0]
        dx = range_x / np.size(self.x)
        dy = range_y / np.size(self.y)
        dr_real = np.sqrt(dx ** 2 + dy ** 2)
        rmax = np.sqrt(range_x ** 2 + range_y ** 2)
        factor = (((quality_factor * dr_real + rmax) ** 2 - (self.wavelength / n) ** 2 - rmax ** 2) / (2 * self.wavelength / n)) ** 2 - rmax ** 2
        z_min = np.sqrt(factor) if factor > 0 else 0
        return print("Minimum distance to propagate (in microns):", z_min)

    def CZT(self, z, xout=None, yout=None):
        """Chirped z-transform propagation (Bluestein) to the plane sampled at (xout, yout).  wave_optics.py:233-263."""
        tic = time.perf_counter()
        if xout is None:
            xout = self.x
        if yout is None:
            yout = self.y
        out = ScalarLight(xout, yout, self.wavelength, self.device, _alloc=False)
        out.field = CZT_jit(self.field, z, self.wavelength, self.x, self.y, xout, yout)
        if VERBOSE:
            print(f"Time taken to perform one CZT propagation (in seconds): {(time.perf_counter() - tic):.4f}")
        return out


class LightSource(ScalarLight):
    """Scalar beams.  Reference: wave_optics.py:462-542."""

    def __init__(self, x, y, wavelength, device=None):
        super().__init__(x, y, wavelength, device)
        self.info = 'Wave optics light source'

    def gaussian_beam(self, w0, E0, center=(0, 0), z_w0=(0, 0), alpha=0):
        """wave_optics.py:468-523 (evaluated once on the host in float64, then uploaded as complex64)."""
        self.field = torch.as_tensor(_gaussian_beam(self.x, self.y, self.k, self.n, w0, E0, center, z_w0, alpha)
                                     .astype(np.complex64), device=self.device)

    def plane_wave(self, A=1, theta=0, phi=0, z0=0):
        """wave_optics.py:525-542."""
        X, Y = np.meshgrid(self.x, self.y)
        f = A * np.exp(1j * self.k * (X * np.sin(theta) * np.cos(phi) + Y * np.sin(theta) * np.sin(phi) + z0 * np.cos(theta)))
        self.field = torch.as_tensor(f.astype(np.complex64), device=self.device)


def _gaussian_beam(x, y, k, n, w0, E0, center, z_w0, alpha):
    w0_x, w0_y = w0
    x0, y0 = center
    z_w0x, z_w0y = z_w0
    Rayleigh_x = k * w0_x ** 2 * n / 2
    Rayleigh_y = k * w0_y ** 2 * n / 2
    Gouy_phase_x = np.arctan2(z_w0x, Rayleigh_x)
    Gouy_phase_y = np.arctan2(z_w0y, Rayleigh_y)
    w_x = w0_x * np.sqrt(1 + (z_w0x / Rayleigh_x) ** 2)
    w_y = w0_y * np.sqrt(1 + (z_w0y / Rayleigh_y) ** 2)
    R_x = 1e12 if z_w0x == 0 else z_w0x * (1 + (Rayleigh_x / z_w0x) ** 2)
    R_y = 1e12 if z_w0x == 0 else z_w0y * (1 + (Rayleigh_y / z_w0y) ** 2)   # reference tests z_w0x twice (wave_optics.py:507)
    X, Y = np.meshgrid(x, y)
    x_rot = X * np.cos(alpha) + Y * np.sin(alpha)
    y_rot = -X * np.sin(alpha) + Y * np.cos(alpha)
    phase = np.exp(-1j * ((k * z_w0x + k * X ** 2 / (2 * R_x) - Gouy_phase_x) + (k * z_w0y + k * Y ** 2 / (2 * R_y) - Gouy_phase_y)))
    return (E0 * (w0_x / w_x) * (w0_y / w_y) * np.exp(-(x_rot - x0) ** 2 / (w_x ** 2) - (y_rot - y0) ** 2 / (w_y ** 2))) * phase
