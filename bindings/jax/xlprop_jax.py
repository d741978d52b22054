"""
JAX side of the drop-in: FFI registration, `custom_vjp` operators and `install()`, which rebinds XLuminA's five propagation
entry points to them so that the reference's optical elements, tables, loss functions and Optax loops run unmodified
(`jit`, `vmap`, `value_and_grad` included).

STATUS: NEVER EXECUTED in the build image of this repository (jax / jaxlib are not installable there).  It is the JAX
counterpart of xlumina_b200/ops.py (the torch skin, which IS tested against fixtures made by the reference's own source):
same C entry points, same operand order, same cotangent convention (flags = 0 is already JAX's).  Written for
jax >= 0.5 (`jax.ffi`); on jax 0.4.33 (the reference's pin) use `jax.extend.ffi` -- same functions.

    import xlprop_jax; xlprop_jax.install()          # before the experiment script star-imports xlumina
"""
import ctypes
import os
from functools import partial

import numpy as np
import jax
import jax.numpy as jnp
from jax import ffi

_HERE = os.path.dirname(os.path.abspath(__file__))
_core = ctypes.CDLL(os.path.join(_HERE, "..", "..", "xlumina_b200", "libxlprop.so"), mode=ctypes.RTLD_GLOBAL)
_core.xl_rs_transfer_bytes.restype = ctypes.c_size_t
_core.xl_rs_transfer_bytes.argtypes = [ctypes.c_int]
_so = ctypes.CDLL(os.path.join(_HERE, "libxlprop_jax.so"))
for _name in ("XlRsFwd", "XlRsBwd", "XlVrsFwd", "XlVrsBwd", "XlCztFwd", "XlCztBwd", "XlHighnaFwd", "XlHighnaBwd"):
    ffi.register_ffi_target(_name, ffi.pycapsule(getattr(_so, _name)), platform="CUDA")

C64 = jnp.complex64
_f = np.float64
# vmap_method="sequential" everywhere: always correct (a vmapped distance needs its own transfer function per item).  A batch
# of fields that SHARES z can go to the handler in one call instead (the C entry points take nfields; "broadcast_all" plus a
# check that z was not batched) -- the optimisation the torch skin already has (ops.rs_propagation).


def _z1(z):
    return jnp.reshape(jnp.asarray(z, dtype=jnp.float64), (1,))


def _uniform(coords):
    """(first, step, last, n) of a uniform axis; the kernels regenerate coordinate grids from these."""
    c = np.asarray(coords, dtype=np.float64)
    step = (c[-1] - c[0]) / (len(c) - 1)
    if not np.allclose(np.diff(c), step, rtol=1e-9, atol=1e-12 * max(1.0, abs(c).max())):
        raise ValueError("xlprop: coordinate axes must be uniformly spaced (toolbox.space grids are)")
    return float(c[0]), float(step), float(c[-1]), len(c)


# ------------------------------------------------------------------------------------------------ scalar RS
def _rs_call(field, z, dx, dy, k):
    n = field.shape[-1]
    out_t = jax.ShapeDtypeStruct(field.shape, C64)
    h_t = jax.ShapeDtypeStruct((_core.xl_rs_transfer_bytes(n),), jnp.uint8)
    return ffi.ffi_call("XlRsFwd", (out_t, h_t), vmap_method="sequential")(field, _z1(z), dx=_f(dx), dy=_f(dy), k=_f(k))


@partial(jax.custom_vjp, nondiff_argnums=(2, 3, 4))
def rs(field, z, dx, dy, k):
    """RS_propagation_jit (wave_optics.py:281-289) for field (..., N, N) complex64 and one distance z."""
    return _rs_call(field, z, dx, dy, k)[0]


def _rs_fwd(field, z, dx, dy, k):
    out, H = _rs_call(field, z, dx, dy, k)
    return out, (field, out, z, H)


def _rs_bwd(dx, dy, k, res, ct):
    field, out, z, H = res
    ct_field, ct_z = ffi.ffi_call("XlRsBwd", (jax.ShapeDtypeStruct(field.shape, C64), jax.ShapeDtypeStruct((1,), jnp.float64)),
                                  vmap_method="sequential")(field, out, ct.astype(C64), H, _z1(z), dx=_f(dx), dy=_f(dy), k=_f(k))
    return ct_field, jnp.reshape(ct_z, jnp.shape(z)).astype(jnp.result_type(z))


rs.defvjp(_rs_fwd, _rs_bwd)


# ------------------------------------------------------------------------------------------------ vectorial RS (Ez formed in the kernel)
def _vrs_call(exy, z, x0, y0, dx, dy, k):
    n = exy.shape[-1]
    out_t = jax.ShapeDtypeStruct((3, n, n), C64)
    h_t = jax.ShapeDtypeStruct((_core.xl_rs_transfer_bytes(n),), jnp.uint8)
    return ffi.ffi_call("XlVrsFwd", (out_t, h_t), vmap_method="sequential")(
        exy, _z1(z), x0=_f(x0), y0=_f(y0), dx=_f(dx), dy=_f(dy), k=_f(k))


@partial(jax.custom_vjp, nondiff_argnums=(2, 3, 4, 5, 6))
def vrs(exy, z, x0, y0, dx, dy, k):
    """VectorizedLight.VRS_propagation's arithmetic (vectorized_optics.py:256-275): exy = stack([Ex, Ey]) -> (3, N, N)."""
    return _vrs_call(exy, z, x0, y0, dx, dy, k)[0]


def _vrs_fwd(exy, z, x0, y0, dx, dy, k):
    out, H = _vrs_call(exy, z, x0, y0, dx, dy, k)
    return out, (exy, out, z, H)


def _vrs_bwd(x0, y0, dx, dy, k, res, ct):
    exy, out, z, H = res
    ct_exy, ct_z = ffi.ffi_call("XlVrsBwd", (jax.ShapeDtypeStruct(exy.shape, C64), jax.ShapeDtypeStruct((1,), jnp.float64)),
                                vmap_method="sequential")(exy, out, ct.astype(C64), H, _z1(z),
                                                          x0=_f(x0), y0=_f(y0), dx=_f(dx), dy=_f(dy), k=_f(k))
    return ct_exy, jnp.reshape(ct_z, jnp.shape(z)).astype(jnp.result_type(z))


vrs.defvjp(_vrs_fwd, _vrs_bwd)


# ------------------------------------------------------------------------------------------------ CZT / VCZT
def _czt_attrs(wavelength, vectorial, gin, gout):
    (x0, dx, y0, dy), (xo0, xol, yo0, yol) = gin, gout
    return dict(wavelength=_f(wavelength), vectorial=np.int64(vectorial), x0=_f(x0), dx=_f(dx), y0=_f(y0), dy=_f(dy),
                xo0=_f(xo0), xol=_f(xol), yo0=_f(yo0), yol=_f(yol))


@partial(jax.custom_vjp, nondiff_argnums=(2, 3, 4, 5, 6))
def czt(field, z, wavelength, vectorial, gin, gout, out_shape):
    """CZT_jit / VCZT (wave_optics.py:333-357, vectorized_optics.py:321-361): field (N,N) or stack([Ex,Ey]) -> out_shape.
    gin = (x0, dx, y0, dy), gout = (xout[0], xout[-1], yout[0], yout[-1]) -- hashable tuples of floats.
    The distance has no derivative here: its cotangent is NaN, so a table that optimises a CZT distance
    fails visibly instead of silently ignoring that path (no reference table does)."""
    return ffi.ffi_call("XlCztFwd", jax.ShapeDtypeStruct(out_shape, C64), vmap_method="sequential")(
        field, _z1(z), **_czt_attrs(wavelength, vectorial, gin, gout))


def _czt_fwd(field, z, wavelength, vectorial, gin, gout, out_shape):
    return czt(field, z, wavelength, vectorial, gin, gout, out_shape), (z, field.shape)


def _czt_bwd(wavelength, vectorial, gin, gout, out_shape, res, ct):
    z, in_shape = res
    ct_field = ffi.ffi_call("XlCztBwd", jax.ShapeDtypeStruct(in_shape, C64), vmap_method="sequential")(
        ct.astype(C64), _z1(z), **_czt_attrs(wavelength, vectorial, gin, gout))
    return ct_field, jnp.full_like(jnp.asarray(z, dtype=jnp.float64), jnp.nan)


czt.defvjp(_czt_fwd, _czt_bwd)


# ------------------------------------------------------------------------------------------------ high-NA objective focusing
def _hna_attrs(radius, f, wavelength, gin, gout):
    (x0, dx, y0, dy), (xo0, xol, yo0, yol) = gin, gout
    return dict(radius=_f(radius), f=_f(f), wavelength=_f(wavelength), x0=_f(x0), dx=_f(dx), y0=_f(y0), dy=_f(dy),
                xo0=_f(xo0), xol=_f(xol), yo0=_f(yo0), yol=_f(yol))


@partial(jax.custom_vjp, nondiff_argnums=(1, 2, 3, 4, 5, 6))
def highna(exy, radius, f, wavelength, gin, gout, out_shape):
    """VCZT_objective_lens' arithmetic (optical_elements.py:515-638): stack([Ex, Ey]) -> (3, My, Mx) in the focal plane."""
    return ffi.ffi_call("XlHighnaFwd", jax.ShapeDtypeStruct(out_shape, C64), vmap_method="sequential")(
        exy, **_hna_attrs(radius, f, wavelength, gin, gout))


def _hna_fwd(exy, radius, f, wavelength, gin, gout, out_shape):
    return highna(exy, radius, f, wavelength, gin, gout, out_shape), exy.shape


def _hna_bwd(radius, f, wavelength, gin, gout, out_shape, in_shape, ct):
    return (ffi.ffi_call("XlHighnaBwd", jax.ShapeDtypeStruct(in_shape, C64), vmap_method="sequential")(
        ct.astype(C64), **_hna_attrs(radius, f, wavelength, gin, gout)),)


highna.defvjp(_hna_fwd, _hna_bwd)


# ------------------------------------------------------------------------------------------------ rebinding the reference
def install():
    """Replace the five propagation entry points of XLuminA by the operators above.  The methods are patched on the
    classes (every module shares the class objects), `VCZT_objective_lens` on `xlumina.optical_elements` (the optical
    tables call it through that module's globals): call install() before user scripts do `from xlumina... import *`."""
    import xlumina.wave_optics as wo
    import xlumina.vectorized_optics as vo
    import xlumina.optical_elements as oe

    def quality(self, z):                                  # wave_optics.py:188-191 == vectorized_optics.py:265-270
        dx, dy = self.x[1] - self.x[0], self.y[1] - self.y[0]
        rmax = jnp.sqrt(jnp.max(self.x ** 2) + jnp.max(self.y ** 2))
        ideal = jnp.sqrt(self.wavelength ** 2 + rmax ** 2 + 2 * self.wavelength * jnp.sqrt(rmax ** 2 + z ** 2)) - rmax
        return ideal / jnp.sqrt(dx ** 2 + dy ** 2)

    def grids(self, xout, yout):
        x0, dx, _, _ = _uniform(self.x)
        y0, dy, _, _ = _uniform(self.y)
        xo0, _, xol, mx = _uniform(xout)
        yo0, _, yol, my = _uniform(yout)
        return (x0, dx, y0, dy), (xo0, xol, yo0, yol), mx, my

    def RS_propagation(self, z):
        dt = self.field.dtype
        _, dx, _, _ = _uniform(self.x)
        _, dy, _, _ = _uniform(self.y)
        out = wo.ScalarLight(self.x, self.y, self.wavelength)
        out.field = rs(self.field.astype(C64), z, dx, dy, float(self.k)).astype(dt)
        return out, quality(self, z)

    def CZT(self, z, xout=None, yout=None):
        xout = self.x if xout is None else xout
        yout = self.y if yout is None else yout
        gin, gout, mx, my = grids(self, xout, yout)
        out = wo.ScalarLight(xout, yout, self.wavelength)
        out.field = czt(self.field.astype(C64), z, float(self.wavelength), 0, gin, gout, (my, mx)).astype(self.field.dtype)
        return out

    def VRS_propagation(self, z):
        dt = self.Ex.dtype
        x0, dx, _, _ = _uniform(self.x)
        y0, dy, _, _ = _uniform(self.y)
        e = vrs(jnp.stack([self.Ex, self.Ey]).astype(C64), z, x0, y0, dx, dy, float(self.k)).astype(dt)
        out = vo.VectorizedLight(self.x, self.y, self.wavelength)
        out.Ex, out.Ey, out.Ez = e[0], e[1], e[2]
        return out, quality(self, z)

    def VCZT(self, z, xout, yout):
        xout = self.x if xout is None else xout
        yout = self.y if yout is None else yout
        gin, gout, mx, my = grids(self, xout, yout)
        e = czt(jnp.stack([self.Ex, self.Ey]).astype(C64), z, float(self.wavelength), 1, gin, gout, (3, my, mx)).astype(self.Ex.dtype)
        out = vo.VectorizedLight(xout, yout, self.wavelength)
        out.Ex, out.Ey, out.Ez = e[0], e[1], e[2]
        return out

    def VCZT_objective_lens(input_field, r, f, xout, yout):
        gin, gout, mx, my = grids(input_field, xout, yout)
        e = highna(jnp.stack([input_field.Ex, input_field.Ey]).astype(C64), float(r), float(f), float(input_field.wavelength),
                   gin, gout, (3, my, mx)).astype(input_field.Ex.dtype)
        out = vo.VectorizedLight(xout, yout, input_field.wavelength)
        out.Ex, out.Ey, out.Ez = e[0], e[1], e[2]
        return out

    wo.ScalarLight.RS_propagation = RS_propagation
    wo.ScalarLight.CZT = CZT
    vo.VectorizedLight.VRS_propagation = VRS_propagation
    vo.VectorizedLight.VCZT = VCZT
    oe.VCZT_objective_lens = VCZT_objective_lens
