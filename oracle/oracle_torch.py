"""
oracle_torch -- torch-CPU complex128 twin of oracle_np with autograd, for gradient ground truth and the fwd+grad CPU baseline.

TEST INFRASTRUCTURE ONLY (same rules as oracle_np.py: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs, never by the product package).  Same line-by-line restatement of the reference
(file:line citations in oracle_np.py); checked against oracle_np in tests/test_oracle.py.

Gradient convention: torch returns conj(dL/dz-bar)-style gradients for complex leaves; the JAX cotangent of the reference
is the complex conjugate of torch's.  Tests compare in torch's convention (the product's torch skin uses it too).
"""
import math

import numpy as np
import torch

C = torch.complex128
R = torch.float64


def _t(a, dtype=None):
    if isinstance(a, torch.Tensor):
        return a
    return torch.as_tensor(np.asarray(a), dtype=dtype)


def transfer_function_RS(z, X, Y, k):
    r = torch.sqrt(X ** 2 + Y ** 2 + z ** 2)
    factor = 1 / (2 * math.pi) * z / r ** 2 * (1 / r - 1j * k)
    zpos = (z > 0) if isinstance(z, torch.Tensor) else torch.tensor(z > 0)
    return torch.where(zpos, torch.exp(1j * k * r) * factor, torch.exp(-1j * k * r) * factor)


def _ext_grid(x, y):
    # wave_optics.py:265-279 closed form: Xext[i,j] = (j-(N-1))*dx  (verified in tests/test_oracle.py)
    nx, ny = len(x), len(y)
    dx, dy = x[1] - x[0], y[1] - y[0]
    xe = (torch.arange(2 * nx - 1, dtype=R) - (nx - 1)) * dx
    ye = (torch.arange(2 * ny - 1, dtype=R) - (ny - 1)) * dy
    Yext, Xext = torch.meshgrid(ye, xe, indexing="ij")
    return nx, ny, dx, dy, Xext, Yext


def RS_propagation(field, x, y, wavelength, z):
    k = 2 * math.pi / wavelength
    nx, ny, dx, dy, Xext, Yext = _ext_grid(x, y)
    H = transfer_function_RS(z, Xext, Yext, k)
    U = torch.nn.functional.pad(field.to(C), (0, nx - 1, 0, ny - 1))
    return (torch.fft.ifft2(torch.fft.fft2(U) * torch.fft.fft2(H)) * dx * dy)[ny - 1:, nx - 1:]


def VRS_propagation(Ex, Ey, x, y, wavelength, z):
    k = 2 * math.pi / wavelength
    X, Y = torch.meshgrid(_t(x, R), _t(y, R), indexing="xy")
    r = torch.sqrt(X ** 2 + Y ** 2 + z ** 2)
    Ez = Ex * X / r + Ey * Y / r
    nx, ny, dx, dy, Xext, Yext = _ext_grid(x, y)
    Hh = torch.fft.fft2(transfer_function_RS(z, Xext, Yext, k))
    outs = []
    for comp in (Ex, Ey, Ez):
        U = torch.nn.functional.pad(comp.to(C), (0, nx - 1, 0, ny - 1))
        outs.append((torch.fft.ifft2(torch.fft.fft2(U) * Hh) * dx * dy)[ny - 1:, nx - 1:])
    return torch.stack(outs)


def Bluestein_method(x, f1, f2, Dm, M_out):
    m, n = x.shape
    D1 = f1 + (M_out * Dm + f2 - f1) / (2 * M_out)
    D2 = f2 + (M_out * Dm + f2 - f1) / (2 * M_out)
    mp = m + M_out - 1
    np2 = int(2 ** int(np.ceil(np.log2(mp))))
    # chirps are constants w.r.t. the field: build them in numpy exactly as oracle_np.compute_fft does
    A = np.exp(1j * 2 * np.pi * D1 / Dm)
    W = np.exp(-1j * 2 * np.pi * (D2 - D1) / (M_out * Dm))
    hj = np.arange(-m + 1, max(M_out - 1, m - 1) + 1)
    h = W ** (hj ** 2 / 2)
    ft = torch.as_tensor(np.fft.fft(1 / h[:mp + 1], np2))
    b = torch.as_tensor(A ** (-(np.arange(m))) * h[np.arange(m - 1, 2 * m - 1)])
    bb = torch.fft.fft(x * b[:, None], n=np2, dim=0)
    bb = torch.fft.ifft(bb * ft[:, None], dim=0)
    out = bb[m:mp + 1, :].T * torch.as_tensor(h[m - 1:mp])[None, :]
    l = np.linspace(0, M_out - 1, M_out) / M_out * (D2 - D1) + D1
    shift = torch.as_tensor(np.exp(-1j * 2 * np.pi * l * (-m / 2 + 1 / 2) / Dm))
    return out * shift[None, :]


def CZT(field, x, y, wavelength, z, xout, yout):
    k = 2 * math.pi / wavelength
    X, Y = torch.meshgrid(_t(x, R), _t(y, R), indexing="xy")
    Xo, Yo = torch.meshgrid(_t(xout, R), _t(yout, R), indexing="xy")
    dx, dy = x[1] - x[0], y[1] - y[0]
    Dm = wavelength * z / dx
    F0 = transfer_function_RS(z, Xo, Yo, k)
    F = transfer_function_RS(z, X, Y, k)
    U = Bluestein_method(field.to(C) * F, yout[0] + Dm / 2, yout[-1] + Dm / 2, Dm, len(yout))
    U = Bluestein_method(U, xout[0] + Dm / 2, xout[-1] + Dm / 2, Dm, len(xout))
    return F0 * U * z * dx * dy * wavelength


def VCZT(Ex, Ey, x, y, wavelength, z, xout, yout):
    X, Y = torch.meshgrid(_t(x, R), _t(y, R), indexing="xy")
    r = torch.sqrt(X ** 2 + Y ** 2 + z ** 2)
    Ez = (Ex * X / r + Ey * Y / r) * z / r
    return torch.stack([CZT(c, x, y, wavelength, z, xout, yout) for c in (Ex, Ey, Ez)])


def VCZT_objective_lens(Ex, Ey, x, y, wavelength, radius, f, xout, yout):
    s = radius / math.sqrt(radius ** 2 + f ** 2)
    X, Y = torch.meshgrid(_t(x, R), _t(y, R), indexing="xy")
    r = torch.sqrt(X ** 2 + Y ** 2)
    phi = torch.atan2(Y, X)
    theta = r / f
    Ez = Ex * X / r + Ey * Y / r
    pupil = torch.where((X ** 2 + Y ** 2) / radius ** 2 < 1, 1.0, 0.0)
    G = pupil * (1 / torch.sqrt(torch.abs(1 - ((X / radius) ** 2 + (Y / radius) ** 2) * s ** 2)))
    ct, st, cp, sp = torch.cos(theta), torch.sin(theta), torch.cos(phi), torch.sin(phi)
    apod = torch.sqrt(torch.abs(ct))
    E0x = (ct * cp ** 2 + sp ** 2) * Ex + (ct * cp * sp - sp * cp) * Ey + (-cp * st) * Ez
    E0y = (sp * ct * cp - cp * sp) * Ex + (ct * sp ** 2 + cp ** 2) * Ey + (-sp * st) * Ez
    E0z = (st * cp) * Ex + (st * sp) * Ey + ct * Ez
    Dm = f * wavelength * (len(x) - 1) / (2 * radius)
    outs = []
    for comp in (E0x, E0y, E0z):
        U = Bluestein_method(apod * G * comp, yout[0] + Dm / 2, yout[-1] + Dm / 2, Dm, len(yout))
        outs.append(Bluestein_method(U, xout[0] + Dm / 2, xout[-1] + Dm / 2, Dm, len(xout)))
    return -(1j * s ** 2 / (f * wavelength)) * torch.stack(outs)
