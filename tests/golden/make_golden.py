"""
Generate the golden fixtures under tests/golden/*.npz by EXECUTING THE REFERENCE'S OWN SOURCE
(/root/reference/xlumina/{wave_optics,vectorized_optics,optical_elements}.py, unmodified, imported from where it lies).

Real JAX cannot be installed in this image, so the reference modules are imported on top of a NumPy-backed stand-in for
`jax` (tests/golden/jaxshim): jit = identity, vmap = loop, jnp = numpy (float64 / complex128, pocketfft instead of XLA's
FFT thunk).  What the fixtures pin is therefore every line of the reference's Python -- grids, padding, crop offsets,
Bluestein slicing and chirps, Ez definitions, lens matrix, constants -- not XLA's floating-point rounding.

Gradients: the propagators are linear in the field, so the reference's exact Jacobian is obtained by pushing basis fields
through the reference itself (small N); golden VJPs are J^T ct (JAX convention, no conjugation).  d/dz goldens are
central differences of the reference output (step 1e-4 um, truncation ~ (k*step)^2/6 ~ 2e-7).

Run (in the build container only; /root/reference does not exist on the GPU box):
    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "jaxshim"))
sys.path.insert(0, "/root/reference")

from xlumina.wave_optics import ScalarLight, LightSource  # noqa: E402
from xlumina.vectorized_optics import VectorizedLight, PolarizedLightSource  # noqa: E402
from xlumina.optical_elements import VCZT_objective_lens, high_NA_objective_lens  # noqa: E402
import xlumina.wave_optics as rwo  # noqa: E402

rng = np.random.default_rng(20261017)


def crand(*shape):
    return rng.standard_normal(shape) + 1j * rng.standard_normal(shape)


def quiet(fn, *a, **k):
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def scalar_light(x, y, lam, field):
    li = ScalarLight(x, y, lam)
    li.field = field
    return li


def vector_light(x, y, lam, ex, ey):
    li = VectorizedLight(x, y, lam)
    li.Ex, li.Ey = ex, ey
    li.Ez = np.zeros_like(ex)
    return li


def jacobian(fn, in_shape):
    """Columns = fn(basis) for a complex-linear fn."""
    n = int(np.prod(in_shape))
    cols = []
    for i in range(n):
        e = np.zeros(n, dtype=complex)
        e[i] = 1.0
        cols.append(np.asarray(fn(e.reshape(in_shape))).ravel())
    return np.stack(cols, axis=1)


def save(name, **arrays):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **arrays)
    print("wrote", path, {k: np.shape(v) for k, v in arrays.items()})


def rs_cases():
    for tag, N, span, lam, z in (("rs_n32_zpos", 32, 400.0, 0.6328, 3000.0), ("rs_n32_zneg", 32, 400.0, 0.6328, -2500.0),
                                 ("rs_n48_far", 48, 1500.0, 0.633, 50000.0)):
        x = np.linspace(-span, span, N)
        y = np.linspace(-span, span, N)
        u = crand(N, N)
        out, q = quiet(scalar_light(x, y, lam, u).RS_propagation, z)
        ct = crand(N, N)
        arrays = dict(x=x, y=y, wavelength=lam, z=z, field=u, out=np.asarray(out.field), quality=float(q), ct=ct)
        if N <= 32:
            J = jacobian(lambda f: quiet(scalar_light(x, y, lam, f).RS_propagation, z)[0].field, (N, N))
            arrays["vjp_field"] = (J.T @ ct.ravel()).reshape(N, N)
        eps = 1e-4
        op = quiet(scalar_light(x, y, lam, u).RS_propagation, z + eps)[0].field
        om = quiet(scalar_light(x, y, lam, u).RS_propagation, z - eps)[0].field
        arrays["vjp_z"] = float(np.real(np.sum(ct * (op - om) / (2 * eps))))
        save(tag, **arrays)
    # the reference's own test configuration (tests/test_wave_optics.py:17-23, 43-47) at reduced N: gaussian beam
    N, lam, z = 64, 633e-3, 1000.0
    x = np.linspace(-1500, 1500, N)
    y = np.linspace(-1500, 1500, N)
    src = LightSource(x, y, lam)
    src.gaussian_beam(w0=(1200, 1200), E0=1)
    out, q = quiet(src.RS_propagation, z)
    cz = quiet(src.CZT, z)
    save("scalar_gaussian_n64", x=x, y=y, wavelength=lam, z=z, field=np.asarray(src.field), rs_out=np.asarray(out.field),
         quality=float(q), czt_out=np.asarray(cz.field))


def vrs_cases():
    for tag, N, span, lam, z in (("vrs_n24", 24, 300.0, 0.6328, 2000.0), ("vrs_n40_zneg", 40, 1500.0, 0.635, -40000.0)):
        x = np.linspace(-span, span, N)
        y = np.linspace(-span, span, N)
        ex, ey = crand(N, N), crand(N, N)

        def fwd(exy, zz=z):
            o, _ = quiet(vector_light(x, y, lam, exy[0], exy[1]).VRS_propagation, zz)
            return np.stack([o.Ex, o.Ey, o.Ez])
        out = fwd(np.stack([ex, ey]))
        ct = crand(3, N, N)
        arrays = dict(x=x, y=y, wavelength=lam, z=z, Ex=ex, Ey=ey, out=out, ct=ct)
        if N <= 24:
            J = jacobian(fwd, (2, N, N))
            arrays["vjp_field"] = (J.T @ ct.ravel()).reshape(2, N, N)
        eps = 1e-4
        arrays["vjp_z"] = float(np.real(np.sum(ct * (fwd(np.stack([ex, ey]), z + eps) - fwd(np.stack([ex, ey]), z - eps)) / (2 * eps))))
        save(tag, **arrays)


def czt_cases():
    for tag, N, Mx, My, span, ospan, lam, z in (("czt_n32_m24x40", 32, 24, 40, 400.0, 60.0, 0.6328, 20000.0),
                                                ("czt_n24_m50", 24, 50, 50, 1500.0, 900.0, 0.633, -30000.0),
                                                ("czt_n40_same", 40, 40, 40, 1500.0, 1500.0, 0.6328, 5000.0)):
        x = np.linspace(-span, span, N)
        y = np.linspace(-span, span, N)
        xo = np.linspace(-ospan, 0.9 * ospan, Mx)
        yo = np.linspace(-0.8 * ospan, ospan, My)
        u = crand(N, N)

        def fwd(f):
            return np.asarray(quiet(scalar_light(x, y, lam, f).CZT, z, xo, yo).field)
        out = fwd(u)
        ct = crand(My, Mx)
        arrays = dict(x=x, y=y, xout=xo, yout=yo, wavelength=lam, z=z, field=u, out=out, ct=ct)
        if N <= 32:
            arrays["vjp_field"] = (jacobian(fwd, (N, N)).T @ ct.ravel()).reshape(N, N)
        save(tag, **arrays)
    # raw Bluestein_method (wave_optics.py:412-460) incl. the M_out > m branch
    for tag, m, n, M in (("bluestein_m48_M100", 48, 5, 100), ("bluestein_m64_M40", 64, 3, 40)):
        xin = crand(m, n)
        Dm, f1, f2 = 37.3, -3.1 + 37.3 / 2, 4.7 + 37.3 / 2
        save(tag, x=xin, Dm=Dm, f1=f1, f2=f2, M_out=M, out=np.asarray(rwo.Bluestein_method(xin, f1, f2, Dm, M)))
    # VCZT
    N, M, span, ospan, lam, z = 24, 30, 300.0, 40.0, 0.6328, 15000.0
    x = np.linspace(-span, span, N)
    y = np.linspace(-span, span, N)
    xo = np.linspace(-ospan, ospan, M)
    yo = np.linspace(-ospan, ospan, M)
    ex, ey = crand(N, N), crand(N, N)

    def vfwd(exy):
        o = quiet(vector_light(x, y, lam, exy[0], exy[1]).VCZT, z, xo, yo)
        return np.stack([o.Ex, o.Ey, o.Ez])
    ct = crand(3, M, M)
    save("vczt_n24_m30", x=x, y=y, xout=xo, yout=yo, wavelength=lam, z=z, Ex=ex, Ey=ey, out=vfwd(np.stack([ex, ey])), ct=ct,
         vjp_field=(jacobian(vfwd, (2, N, N)).T @ ct.ravel()).reshape(2, N, N))


def highna_cases():
    for tag, N, Mx, My, span in (("highna_n24_m20", 24, 20, 20, 2500.0), ("highna_n40_m30x26", 40, 30, 26, 2500.0)):
        lam, R, f = 0.635, 1800.0, 2000.0           # experiments/hybrid_sharp_optical_table.py:26-46
        x = np.linspace(-span, span, N)
        y = np.linspace(-span, span, N)
        xo = np.linspace(-10, 10, Mx)
        yo = np.linspace(-10, 10, My)
        ex, ey = crand(N, N), crand(N, N)

        def fwd(exy):
            o = quiet(VCZT_objective_lens, vector_light(x, y, lam, exy[0], exy[1]), R, f, xo, yo)
            return np.stack([o.Ex, o.Ey, o.Ez])
        out = fwd(np.stack([ex, ey]))
        lens, s = high_NA_objective_lens(vector_light(x, y, lam, ex, ey), R, f)
        ct = crand(3, My, Mx)
        arrays = dict(x=x, y=y, xout=xo, yout=yo, wavelength=lam, radius=R, f=f, Ex=ex, Ey=ey, out=out, lens=np.asarray(lens),
                      sin_theta_max=float(s), ct=ct)
        if N <= 24:
            arrays["vjp_field"] = (jacobian(fwd, (2, N, N)).T @ ct.ravel()).reshape(2, N, N)
        save(tag, **arrays)


if __name__ == "__main__":
    rs_cases()
    vrs_cases()
    czt_cases()
    highna_cases()
