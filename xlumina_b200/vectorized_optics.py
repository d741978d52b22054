"""
Vectorial fields: VectorizedLight / PolarizedLightSource with VRS_propagation and VCZT on libxlprop.so.
Mirror of xlumina/vectorized_optics.py (reference line numbers per docstring).
"""
import math
import time

import numpy as np
import torch

from . import ops
from .wave_optics import _quality_factor, _gaussian_beam, build_grid, DEFAULT_DEVICE
from . import wave_optics as _wo


def VRS_propagation_jit(Ex, Ey, z, nx, ny, x0, y0, dx, dy, k):
    """Seam function, vectorized_optics.py:364-373; takes (Ex,Ey) because the incoming Ez is ignored (:258-261)."""
    if nx != ny:
        raise ValueError("square grids only")
    return ops.vrs_propagation(Ex, Ey, z, x0, y0, dx, dy, k)


def VCZT_jit(Ex, Ey, z, wavelength, x, y, xout, yout):
    """Seam function, vectorized_optics.py:375-384."""
    return ops.vczt(Ex, Ey, z, wavelength, x, y, xout, yout)


class VectorizedLight:
    """(Ex, Ey, Ez) on a square grid.  Reference: vectorized_optics.py:36-49."""

    def __init__(self, x=None, y=None, wavelength=None, device=None, _alloc=True):
        self.x = x
        self.y = y
        self.wavelength = wavelength
        self.k = 2 * math.pi / wavelength
        self.n = 1
        self.device = torch.device(device or DEFAULT_DEVICE)
        if _alloc:
            shape = (len(y), len(x))
            self.Ex = torch.zeros(shape, dtype=torch.complex64, device=self.device)
            self.Ey = torch.zeros(shape, dtype=torch.complex64, device=self.device)
            self.Ez = torch.zeros(shape, dtype=torch.complex64, device=self.device)
        else:   # the caller assigns all three planes right away (skips three N^2 memsets per element of a table)
            self.Ex = self.Ey = self.Ez = None
        self.info = 'Vectorized light'

    @property
    def X(self):
        return torch.as_tensor(np.meshgrid(self.x, self.y)[0], device=self.device)

    @property
    def Y(self):
        return torch.as_tensor(np.meshgrid(self.x, self.y)[1], device=self.device)

    def VRS_propagation(self, z):
        """Vectorial Rayleigh-Sommerfeld propagation; returns (VectorizedLight, quality_factor).  vectorized_optics.py:244-284."""
        tic = time.perf_counter()
        nx, ny, dx, dy = build_grid(self.x, self.y)
        quality_factor = _quality_factor(self.x, self.y, self.wavelength, z)
        E = VRS_propagation_jit(self.Ex, self.Ey, z, nx, ny, float(self.x[0]), float(self.y[0]), dx, dy, self.k)
        out = VectorizedLight(self.x, self.y, self.wavelength, self.device, _alloc=False)
        out.Ex, out.Ey, out.Ez = E[0], E[1], E[2]
        if _wo.VERBOSE:
            print(f"Time taken to perform one VRS propagation (in seconds): {(time.perf_counter() - tic):.4f}")
        return out, quality_factor

    def get_VRS_minimum_z(self, n=1, quality_factor=1):
        """vectorized_optics.py:286-319 (diagnostic; host math)."""
        range_x = self.x[-1] - self.x[0]
        range_y = self.y[-1] - self.y[0]
        dx = range_x / np.size(self.x)
        dy = range_y / np.size(self.y)
        dr_real = np.sqrt(dx ** 2 + dy ** 2)
        rmax = np.sqrt(range_x ** 2 + range_y ** 2)
        factor = (((quality_factor * dr_real + rmax) ** 2 - (self.wavelength / n) ** 2 - rmax ** 2) / (2 * self.wavelength / n)) ** 2 - rmax ** 2
        z_min = np.sqrt(factor) if factor > 0 else 0
        return print("Minimum distance to propagate (in um):", z_min)

    def VCZT(self, z, xout, yout):
        """Vectorial chirped z-transform propagation.  vectorized_optics.py:321-361."""
        tic = time.perf_counter()
        if xout is None:
            xout = self.x
        if yout is None:
            yout = self.y
        E = VCZT_jit(self.Ex, self.Ey, z, self.wavelength, self.x, self.y, xout, yout)
        out = VectorizedLight(xout, yout, self.wavelength, self.device, _alloc=False)
        out.Ex, out.Ey, out.Ez = E[0], E[1], E[2]
        if _wo.VERBOSE:
            print(f"Time taken to perform one VCZT propagation (in seconds):  {(time.perf_counter() - tic):.4f}")
        return out


class PolarizedLightSource(VectorizedLight):
    """Polarised beams.  Reference: vectorized_optics.py:396-490."""

    def __init__(self, x, y, wavelength, device=None):
        super().__init__(x, y, wavelength, device)
        self.info = 'Vectorized light source'

    def gaussian_beam(self, w0, jones_vector, center=(0, 0), z_w0=(0, 0), alpha=0):
        """vectorized_optics.py:402-465."""
        g = _gaussian_beam(self.x, self.y, self.k, self.n, w0, 1.0, center, z_w0, alpha)
        j = np.asarray(jones_vector, dtype=np.complex128)   # complex Jones vectors (circular polarisation) as in the reference
        j = j / np.linalg.norm(j)
        self.Ex = torch.as_tensor((j[0] * g).astype(np.complex64), device=self.device)
        self.Ey = torch.as_tensor((j[1] * g).astype(np.complex64), device=self.device)
        self.Ez = torch.zeros_like(self.Ex)

    def plane_wave(self, jones_vector, theta=0, phi=0, z0=0):
        """vectorized_optics.py:467-490."""
        j = np.asarray(jones_vector, dtype=np.complex128)   # complex Jones vectors (circular polarisation) as in the reference
        j = j / np.linalg.norm(j)
        X, Y = np.meshgrid(self.x, self.y)
        pw = np.exp(1j * self.k * (X * np.sin(theta) * np.cos(phi) + Y * np.sin(theta) * np.sin(phi) + z0 * np.cos(theta)))
        self.Ex = torch.as_tensor((j[0] * pw).astype(np.complex64), device=self.device)
        self.Ey = torch.as_tensor((j[1] * pw).astype(np.complex64), device=self.device)
        self.Ez = torch.zeros_like(self.Ex)
