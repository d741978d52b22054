"""
Optical elements.  The propagating element -- high-NA objective focusing -- runs on libxlprop.so; the pointwise elements
that bracket the propagators in the reference's optical tables (SLM, sSLM, LCD, beam splitters, lenses) and the sharp-focus
table built from them are closed-form torch arithmetic on the field planes.

High-NA objective focusing on libxlprop.so.  Mirror of xlumina/optical_elements.py:515-672: the lens factor
(high_NA_objective_lens + _high_NA_objective_lens_, :515-594) is fused into the load of the first Bluestein pass and the
constant -i sin^2(theta_max)/(f lambda) (:627) into the store of the second, so the (3,N,N) lens field never exists in HBM.
"""
import math
import time

import numpy as np
import torch

from . import ops
from .vectorized_optics import VectorizedLight
from .wave_optics import ScalarLight
from . import wave_optics as _wo

cm = 1e4


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
    out = VectorizedLight(xout, yout, input_field.wavelength, input_field.device, _alloc=False)
    out.Ex, out.Ey, out.Ez = E[0], E[1], E[2]
    if _wo.VERBOSE:
        print(f"Time taken to perform one VCZT propagation through objective lens (in seconds):  {(time.perf_counter() - tic):.4f}")
    return out


# ------------------------------------------------------------------------------------------------------------------------
# Pointwise elements that bracket the propagators in the reference's optical tables.  They are written as
# closed-form complex arithmetic on the field planes: the reference materialises an (N^2, 2, 2) Jones tensor per element
# and batch-multiplies it (optical_elements.py:208-213); here the 2x2 product is expanded, so an element is one read and one
# write of the planes it changes.  Parameters may be Python floats, NumPy arrays or torch tensors (autograd flows through).
def _real(v, like, dtype=None):
    """`v` as a real tensor on the device of `like`, in the real dtype matching `like` unless `dtype` is given."""
    rd = dtype or (torch.float64 if like.dtype == torch.complex128 else torch.float32)
    if isinstance(v, torch.Tensor):
        return v.to(device=like.device, dtype=rd)
    return torch.as_tensor(np.asarray(v, dtype=np.float64), device=like.device).to(rd)


def _phasor(phase, like):
    """exp(i*phase) in the complex dtype of `like`."""
    p = _real(phase, like)
    return torch.polar(torch.ones_like(p), p)


def _new_vector(src):
    return VectorizedLight(src.x, src.y, src.wavelength, src.device, _alloc=False)


_zero_planes = {}


def _zeros_like_plane(t):
    """The zero Ez plane the reference attaches to the output of most elements.  Nothing on the path reads the input Ez (the
    propagators recompute it from Ex, Ey), so one shared read-only plane per (device, dtype, shape) stands for all of them."""
    key = (t.device, t.dtype, tuple(t.shape))
    z = _zero_planes.get(key)
    if z is None:
        if len(_zero_planes) > 16:
            _zero_planes.clear()
        z = _zero_planes[key] = torch.zeros_like(t).detach()
    return z


def phase_scalar_SLM(phase):
    """optical_elements.py:76-85."""
    return torch.polar(torch.ones_like(phase), phase)


def SLM(input_field, phase_array, shape=None):
    """Phase-only spatial light modulator for ScalarLight: field * exp(i*phase) pixel-wise; returns (light, slm).
    Reference: optical_elements.py:87-103 (`shape` is implied by the mask and only kept for signature parity)."""
    slm = _phasor(phase_array, input_field.field)
    out = ScalarLight(input_field.x, input_field.y, input_field.wavelength, input_field.device, _alloc=False)
    out.field = input_field.field * slm
    return out, slm


def sSLM(input_field, alpha_array=None, phi_array=None):
    """Super-SLM: independent phase masks on Ex (alpha) and Ey (phi); Ez is carried over.  optical_elements.py:186-222,
    Jones matrix diag(e^{i alpha}, e^{i phi}) (:142-153)."""
    return _sslm(input_field, alpha_array, phi_array, 1.0, 0.0)


def _sslm(input_field, alpha, phi, scale, offset):
    """sSLM with the phases scale * alpha + offset, scale * phi + offset (the tables pass the optimizer's raw masks)."""
    out = _new_vector(input_field)
    if ops.elements_on(input_field.Ex):       # one kernel, forward and backward (xl_el_sslm)
        out.Ex, out.Ey = ops.el_sslm(input_field.Ex, input_field.Ey, alpha, phi, scale, offset)
    else:
        out.Ex = input_field.Ex * _phasor(_real(alpha, input_field.Ex, torch.float64) * scale + offset, input_field.Ex)
        out.Ey = input_field.Ey * _phasor(_real(phi, input_field.Ey, torch.float64) * scale + offset, input_field.Ey)
    out.Ez = input_field.Ez
    return out


def sSLM_with_amplitude(input_field, alpha_array=None, phi_array=None, A1_array=None, A2_array=None):
    """Super-SLM with amplitude masks: diag(A1 e^{i alpha}, A2 e^{i phi}).  optical_elements.py:224-264, :155-168."""
    out = _new_vector(input_field)
    pa = _real(alpha_array, input_field.Ex)
    pp = _real(phi_array, input_field.Ey)
    out.Ex = input_field.Ex * torch.polar(_real(A1_array, input_field.Ex).expand_as(pa), pa)
    out.Ey = input_field.Ey * torch.polar(_real(A2_array, input_field.Ey).expand_as(pp), pp)
    out.Ez = input_field.Ez
    return out


def jones_LCD(eta, theta, like):
    """The three distinct entries (a, b, d) of the symmetric retarder matrix [[a, b], [b, d]] with delta = 0
    (optical_elements.py:123-140, :170-180):  a = e^{-i eta/2} cos^2 + e^{i eta/2} sin^2,  b = (e^{-i eta/2} - e^{i eta/2}) sin cos,
    d = e^{-i eta/2} sin^2 + e^{i eta/2} cos^2."""
    e = _real(eta, like, torch.float64).reshape(())
    t = _real(theta, like, torch.float64).reshape(())
    c, s = torch.cos(t), torch.sin(t)
    ch, sh = torch.cos(e / 2), torch.sin(e / 2)
    a = torch.complex(ch, -sh * (c * c - s * s))
    b = torch.complex(torch.zeros_like(sh), -2 * sh * s * c)
    d = torch.complex(ch, sh * (c * c - s * s))
    return a.to(like.dtype), b.to(like.dtype), d.to(like.dtype)


def LCD(input_field, eta, theta):
    """Liquid-crystal device = uniform linear wave plate of retardance eta with its fast axis at theta; Ez carried over.
    optical_elements.py:266-305 (the constant (N, N) eta/theta cell of toolbox.build_LCD_cell is never built)."""
    return _lcd(input_field, eta, theta, 1.0, 0.0)


def _lcd(input_field, eta, theta, scale, offset):
    out = _new_vector(input_field)
    if ops.elements_on(input_field.Ex):       # xl_el_lcd: Jones matrix, VJP and the (eta, theta) reductions in the kernels
        out.Ex, out.Ey = ops.el_lcd(input_field.Ex, input_field.Ey, eta, theta, scale, offset)
    else:
        like = input_field.Ex
        a, b, d = jones_LCD(_real(eta, like, torch.float64) * scale + offset, _real(theta, like, torch.float64) * scale + offset, like)
        out.Ex = a * input_field.Ex + b * input_field.Ey
        out.Ey = b * input_field.Ex + d * input_field.Ey
    out.Ez = input_field.Ez
    return out


def linear_polarizer(input_field, alpha):
    """Pixel-wise linear polariser with transmission angle alpha[i, j]; the output Ez is zero as in the reference.
    optical_elements.py:307-332, Jones matrix :110-121."""
    al = _real(alpha, input_field.Ex)
    c, s = torch.cos(al), torch.sin(al)
    proj = c * input_field.Ex + s * input_field.Ey
    out = _new_vector(input_field)
    out.Ex = c * proj
    out.Ey = s * proj
    out.Ez = _zeros_like_plane(input_field.Ex)
    return out


def BS_symmetric(a, b, theta):
    """Lossy symmetric beam splitter: c = R a + i T b, d = i T a + R b with T = |cos theta|, R = |sin theta|, both reduced by
    0.01 T; the outputs' Ez are zero.  optical_elements.py:334-392."""
    return _bs(a, b, theta, 1.0, 0.0)


def _bs(a, b, theta, scale, offset):
    if ops.elements_on(a.Ex):                 # xl_el_bs: both outputs, the VJP and the theta reduction in the kernels
        c, d = _new_vector(a), _new_vector(a)
        c.Ex, c.Ey, d.Ex, d.Ey = ops.el_bs(a.Ex, a.Ey, b.Ex, b.Ey, theta, scale, offset)
        c.Ez = _zeros_like_plane(a.Ex)
        d.Ez = _zeros_like_plane(a.Ex)
        return c, d
    th = (_real(theta, a.Ex, torch.float64) * scale + offset).reshape(())
    T = torch.abs(torch.cos(th))
    R = torch.abs(torch.sin(th))
    noise = T * 0.01
    T = T - noise
    R = R - noise
    rd = torch.float64 if a.Ex.dtype == torch.complex128 else torch.float32
    Rr = R.to(rd)
    iT = torch.complex(torch.zeros_like(T), T).to(a.Ex.dtype)
    c, d = _new_vector(a), _new_vector(a)
    c.Ex = Rr * a.Ex + iT * b.Ex
    c.Ey = Rr * a.Ey + iT * b.Ey
    d.Ex = iT * a.Ex + Rr * b.Ex
    d.Ey = iT * a.Ey + Rr * b.Ey
    c.Ez = _zeros_like_plane(a.Ex)
    d.Ez = _zeros_like_plane(a.Ex)
    return c, d


def circular_mask(X, Y, r):
    """optical_elements.py:781-795."""
    rx, ry = r
    return torch.where((X ** 2 / rx ** 2 + Y ** 2 / ry ** 2) < 1, 1, 0)


def _apply_mask(input_field, mask64):
    """field * mask for scalar light, (Ex, Ey) * mask for vectorial light (Ez zero, as in the reference); `mask64` is a
    complex128 plane built in float64 (lens phases reach 1e4-1e5 rad), cast once to the field dtype."""
    if input_field.info in ('Wave optics light', 'Wave optics light source'):
        mask = mask64.to(input_field.field.dtype)
        out = ScalarLight(input_field.x, input_field.y, input_field.wavelength, input_field.device, _alloc=False)
        out.field = input_field.field * mask
    elif input_field.info in ('Vectorized light', 'Vectorized light source'):
        mask = mask64.to(input_field.Ex.dtype)
        out = _new_vector(input_field)
        out.Ex = input_field.Ex * mask
        out.Ey = input_field.Ey * mask
        out.Ez = _zeros_like_plane(input_field.Ex)
    else:
        raise ValueError("Invalid input. Please use ScalarLight or VectorizedLight object.")
    return out, mask


def _unit_phasor(phase):
    return torch.polar(torch.ones_like(phase), phase)


def lens(input_field, radius, focal):
    """Thin lens with a pupil of radii `radius` and focal lengths `focal` (both (x, y) pairs); returns (light, lens mask).
    optical_elements.py:678-703."""
    fx, fy = focal
    X = input_field.X.to(torch.float64)
    Y = input_field.Y.to(torch.float64)
    ph = -input_field.k * (X ** 2 / (2 * fx) + Y ** 2 / (2 * fy))
    return _apply_mask(input_field, circular_mask(X, Y, radius) * _unit_phasor(ph))


def cylindrical_lens(input_field, focal_length, refractive_index=1.5, angle=0):
    """Plano-convex cylindrical lens rotated by `angle`; returns (light, lens mask).  optical_elements.py:705-744."""
    X = input_field.X.to(torch.float64)
    Y = input_field.Y.to(torch.float64)
    ang = _real(angle, X, torch.float64)               # a tensor parameter keeps its gradient (the reference is differentiable in it)
    Xrot = X * torch.cos(ang) + Y * torch.sin(ang)
    R = focal_length * (refractive_index - 1)
    thickness = R - torch.sqrt(R ** 2 - Xrot ** 2)
    phase = input_field.k * (refractive_index - 1) * (Xrot ** 2 / (2 * focal_length) + thickness)
    return _apply_mask(input_field, _unit_phasor(-phase))


def axicon_lens(input_field, alpha, n=1.5):
    """Axicon of angle alpha (Bessel-beam generator); returns (light, axicon mask).  optical_elements.py:746-779."""
    X = input_field.X.to(torch.float64)
    Y = input_field.Y.to(torch.float64)
    r = torch.sqrt(X ** 2 + Y ** 2)
    al = _real(alpha, X, torch.float64)                # a tensor parameter keeps its gradient
    phase_shift = input_field.k * r * (n - 1) * torch.sin(al) * torch.tan(al)
    return _apply_mask(input_field, _unit_phasor(-phase_shift))


# ------------------------------------------------------------------------------------------------------------------------
# Building blocks and the sharp-focus table (BASELINE.json config 3)
def building_block(input_light, alpha, phi, z, eta, theta):
    """sSLM(alpha, phi) -> VRS(z) -> LCD(eta, theta).  optical_elements.py:919-939."""
    l_modulated = sSLM(input_light, alpha, phi)
    l_propagated, _ = l_modulated.VRS_propagation(z)
    return LCD(l_propagated, eta, theta)


def bb_amplitude_and_phase_mod(input_light, alpha, phi, amp1, amp2, z, eta, theta):
    """sSLM_with_amplitude -> VRS(z) -> LCD.  optical_elements.py:894-917."""
    l_modulated = sSLM_with_amplitude(input_light, alpha, phi, amp1, amp2)
    l_propagated, _ = l_modulated.VRS_propagation(z)
    return LCD(l_propagated, eta, theta)


def _param(p, device):
    """Optimizer parameter as a float64 tensor: distances reach 1e6 um, where float32 would lose the optical phase."""
    if isinstance(p, torch.Tensor):
        return p.to(device=device, dtype=torch.float64)
    return torch.as_tensor(np.asarray(p, dtype=np.float64), device=device)


def hybrid_setup_sharp_focus(ls1, ls2, ls3, ls4, ls5, ls6, parameters, fixed_params, distance_offset=8.9):
    """The 3x3 beam-splitter grid with three building blocks and six high-NA detectors of the Dorn-Quabis-Leuchs
    rediscovery (16 VRS propagations + 6 objective-lens focusings): returns (intensities (6, M, M) = |Ez|^2, detector list).

    parameters (29 entries, each in (0, 1)): [phase1_1, phase1_2, eta1, theta1, z1_1, z1_2, (same for blocks 2 and 3),
    bs1..bs9, z4, z5]; fixed_params = [r, f, xout, yout].  Reference: optical_elements.py:1503-1649.

    Distances that occur more than once (z1_1+z1_2, z2_1+z2_2, z3_1+z3_2, z4, z5) are formed ONCE, so with
    ops.set_transfer_cache(n >= 5) every repeat reuses the transfer function."""
    r, f, xout, yout = fixed_params[0], fixed_params[1], fixed_params[2], fixed_params[3]
    dev = ls1.device
    two_pi = 2 * math.pi
    cache = {}

    def P(i):                                  # float64 view of parameter i, made on first use
        if i not in cache:
            cache[i] = _param(parameters[i], dev)
        return cache[i]

    def angle(i):
        return P(i) * two_pi - math.pi

    def dist(i):
        return (torch.abs(P(i)) * 100 + distance_offset) * cm

    # With complex64 planes on the GPU the elements are single kernels that take the optimizer's RAW parameters and apply
    # the map p -> p*2pi - pi themselves (xl_el_*): no per-pixel float64 copies of the masks, no scalar-algebra launches.
    raw = ops.elements_on(ls1.Ex)

    def BS(a, b, i):
        return _bs(a, b, P(18 + i), two_pi, -math.pi) if raw else BS_symmetric(a, b, angle(18 + i))

    def block(light, i0, z):                   # sSLM(masks i0, i0+1) -> VRS(z) -> LCD(i0+2, i0+3): building_block
        if not raw:
            return building_block(light, angle(i0), angle(i0 + 1), z, angle(i0 + 2), angle(i0 + 3))
        l_modulated = _sslm(light, parameters[i0], parameters[i0 + 1], two_pi, -math.pi)
        l_propagated, _ = l_modulated.VRS_propagation(z)
        return _lcd(l_propagated, P(i0 + 2), P(i0 + 3), two_pi, -math.pi)

    z1_1, z1_2, z2_1, z2_2, z3_1, z3_2 = dist(4), dist(5), dist(10), dist(11), dist(16), dist(17)
    z4, z5 = dist(27), dist(28)
    z1s, z2s, z3s = z1_1 + z1_2, z2_1 + z2_2, z3_1 + z3_2

    # 1st row
    c1, d1 = BS(ls1, ls4, 0)
    b2, _ = block(c1, 0, z1_1).VRS_propagation(z1_2)
    c2, d2 = BS(ls2, b2, 1)
    b3, _ = c2.VRS_propagation(z2s)
    c3, d3 = BS(ls3, b3, 2)
    b_det1, _ = c3.VRS_propagation(z3s)
    det_1 = VCZT_objective_lens(b_det1, r, f, xout, yout)
    # mid space
    a5, _ = d2.VRS_propagation(z4)
    a6, _ = d3.VRS_propagation(z4)
    # 2nd row
    c4, d4 = BS(d1, ls5, 3)
    b5, _ = c4.VRS_propagation(z1s)
    c5, d5 = BS(a5, b5, 4)
    b6, _ = block(c5, 6, z2_1).VRS_propagation(z2_2)
    c6, d6 = BS(a6, b6, 5)
    b_det2, _ = c6.VRS_propagation(z3s)
    det_2 = VCZT_objective_lens(b_det2, r, f, xout, yout)
    # mid space
    a8, _ = d5.VRS_propagation(z5)
    a9, _ = d6.VRS_propagation(z5)
    # 3rd row
    c7, d7 = BS(d4, ls6, 6)
    b8, _ = c7.VRS_propagation(z1s)
    c8, d8 = BS(a8, b8, 7)
    b9, _ = c8.VRS_propagation(z2s)
    c9, d9 = BS(a9, b9, 8)
    b_det3, _ = block(c9, 12, z3_1).VRS_propagation(z3_2)
    det_3 = VCZT_objective_lens(b_det3, r, f, xout, yout)
    # detector row
    det_4 = VCZT_objective_lens(d7, r, f, xout, yout)
    det_5 = VCZT_objective_lens(d8, r, f, xout, yout)
    det_6 = VCZT_objective_lens(d9, r, f, xout, yout)

    detector_array = [det_1, det_2, det_3, det_4, det_5, det_6]
    intensities = torch.stack([d.Ez.real ** 2 + d.Ez.imag ** 2 for d in detector_array])
    return intensities, detector_array
