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


def element_cases():
    """Pointwise elements (optical_elements.py:87-392, 678-703) on random fields: one fixture with every element's output."""
    from xlumina.optical_elements import (SLM, sSLM, sSLM_with_amplitude, LCD, linear_polarizer, BS_symmetric, lens,
                                          cylindrical_lens, axicon_lens)
    N, span, lam = 16, 900.0, 0.635
    x = np.linspace(-span, span, N)
    a = vector_light(x, x, lam, crand(N, N), crand(N, N))
    a.Ez = crand(N, N)
    b = vector_light(x, x, lam, crand(N, N), crand(N, N))
    alpha, phi = rng.uniform(-np.pi, np.pi, (N, N)), rng.uniform(-np.pi, np.pi, (N, N))
    A1, A2 = rng.uniform(0, 1, (N, N)), rng.uniform(0, 1, (N, N))
    eta, theta, bs_theta = 1.234, -0.777, 2.2
    pol = rng.uniform(-np.pi, np.pi, (N, N))
    u = crand(N, N)
    comps = lambda li: np.stack([li.Ex, li.Ey, li.Ez])   # noqa: E731
    s_out, slm = SLM(scalar_light(x, x, lam, u), alpha, N)
    c, d = BS_symmetric(a, b, bs_theta)
    l_s, lens_s = lens(scalar_light(x, x, lam, u), (700.0, 500.0), (4.0e4, 6.0e4))
    l_v, _ = lens(a, (700.0, 500.0), (4.0e4, 6.0e4))
    _, cyl = cylindrical_lens(scalar_light(x, x, lam, u), 5.0e4, 1.5, 0.3)
    ax_v, axi = axicon_lens(a, 0.05)
    save("elements_n16", x=x, wavelength=lam, a=comps(a), b=comps(b), alpha=alpha, phi=phi, A1=A1, A2=A2, eta=eta, theta=theta,
         bs_theta=bs_theta, pol=pol, u=u, slm_out=np.asarray(s_out.field), slm=np.asarray(slm),
         sslm=comps(sSLM(a, alpha, phi)), sslm_amp=comps(sSLM_with_amplitude(a, alpha, phi, A1, A2)),
         lcd=comps(LCD(a, eta, theta)), lp=comps(linear_polarizer(a, pol)), bs_c=comps(c), bs_d=comps(d),
         lens_radius=np.array([700.0, 500.0]), lens_focal=np.array([4.0e4, 6.0e4]), lens_scalar=np.asarray(l_s.field),
         lens_mask=np.asarray(lens_s), lens_vector=comps(l_v), cyl_mask=np.asarray(cyl), axicon_mask=np.asarray(axi),
         axicon_vector=comps(ax_v))


def table_cases():
    """BASELINE config 3 at reduced size: hybrid_setup_sharp_focus (optical_elements.py:1503-1649) + small_area_hybrid +
    softmin, and config 4: the dual-SLM 4f table (experiments/four_f_optical_table.py:36-141).  Directional derivatives of the
    losses are finite differences of the reference itself (steps chosen per table, see below)."""
    from xlumina.optical_elements import hybrid_setup_sharp_focus
    from xlumina.loss_functions import vectorized_loss_hybrid
    from xlumina.toolbox import softmin
    N, M, lam = 32, 20, 0.635
    x = np.linspace(-2500.0, 2500.0, N)
    xo = np.linspace(-10.0, 10.0, M)
    ls = PolarizedLightSource(x, x, lam)
    ls.gaussian_beam(w0=(1200.0, 1200.0), jones_vector=(1, 1))
    big = (0, 1, 6, 7, 12, 13)
    params = [rng.uniform(0, 1, (N, N)) if i in big else rng.uniform(0, 1, (1,)) for i in range(29)]
    W = rng.uniform(0.0, 1.0, (6, M, M))      # smooth surrogate loss sum(W * I): the small-area loss is piecewise smooth only

    def losses(p):
        inten, _ = quiet(hybrid_setup_sharp_focus, ls, ls, ls, ls, ls, ls, p, [1800.0, 2000.0, xo, xo])
        lv = np.asarray(vectorized_loss_hybrid(inten))
        return np.asarray(inten), lv, float(softmin(lv)), float(np.sum(W * np.asarray(inten)))
    inten, lv, l_soft, l_lin = losses(params)
    arrays = dict(x=x, xout=xo, wavelength=lam, radius=1800.0, f=2000.0, intensities=inten, loss_vec=lv, loss_softmin=l_soft,
                  W=W, loss_linear=l_lin, Ex=np.asarray(ls.Ex), Ey=np.asarray(ls.Ey))
    for i, p in enumerate(params):
        arrays["p%02d" % i] = p
    dist = (4, 5, 10, 11, 16, 17, 27, 28)
    for tag, which in (("all", range(29)), ("dist", dist), ("other", [i for i in range(29) if i not in dist])):
        v = [rng.standard_normal(np.shape(p)) if i in which else np.zeros(np.shape(p)) for i, p in enumerate(params)]
        # The beams interfere after different path lengths, so the losses oscillate with the distances at the optical
        # period: a 4th-order stencil with h = 1e-9 (1e-3 um) for directions that move distances (stable to ~1e-7 between
        # h and 2h; smaller steps drown in the k*z ~ 1e7 rad rounding of float64, larger ones in truncation), h = 1e-7 otherwise.
        h = 1e-7 if tag == "other" else 1e-9
        at = {t: losses([p + t * h * d for p, d in zip(params, v)]) for t in (-2, -1, 1, 2)}
        for j, name in ((3, "dlin_"), (2, "dsoft_")):
            c1 = (at[1][j] - at[-1][j]) / (2 * h)
            c2 = (at[2][j] - at[-2][j]) / (4 * h)
            if name == "dlin_" or tag == "other":      # the small-area loss: only along directions that keep its mask fixed
                arrays[name + tag] = (4 * c1 - c2) / 3
            print("  sharp focus", name + tag, "h:", c1, "2h:", c2, "richardson:", (4 * c1 - c2) / 3)
        for i, d in enumerate(v):
            arrays["v_%s_%02d" % (tag, i)] = d
    save("sharp_focus_n32", **arrays)

    # 4f table: the experiment module builds its 1024^2 globals at import; they are replaced by a small grid afterwards
    sys.path.insert(0, "/root/reference/experiments")
    import four_f_optical_table as ff
    N, lam, B = 32, 0.6328, 3
    x = np.linspace(-1500.0, 1500.0, N)
    src = LightSource(x, x, lam)
    src.gaussian_beam(w0=(1200.0, 1200.0), E0=1)
    beam = np.asarray(src.field).copy()
    ff.x, ff.y, ff.wavelength, ff.shape, ff.input_light = x, x, lam, N, src
    X, Y = np.meshgrid(x, x)
    masks = np.stack([((X / r1) ** 2 + (Y / r2) ** 2 < 1).astype(float) for r1, r2 in rng.uniform(300, 1200, (B, 2))])
    targets = rng.uniform(0, 0.3, (B, N, N))
    p4 = [rng.uniform(0.027, 1, (1,)) for _ in range(3)] + [rng.uniform(0, 1, (N, N)) for _ in range(2)]

    def run4(p):
        outs = []
        for m in masks:                       # vmap traces the body once, so the mask is applied to the pristine beam
            src.field = beam.copy()
            outs.append(np.asarray(quiet(ff.batch_dualSLM_4f, m, x, x, lam, p)[0]))
        src.field = beam.copy()
        o = np.stack(outs)
        return o, float(ff.mean_batch_MSE_Intensity(o, targets)[0])
    o4, l4 = run4(p4)
    arrays = dict(x=x, wavelength=lam, beam=beam, masks=masks, targets=targets, intensities=o4, loss=l4)
    for i, p in enumerate(p4):
        arrays["p%d" % i] = p
    # single beam path: the e^{ikz} carrier drops out of the intensities, so a 1e-6 step (1 um) sits on the plateau between
    # float64 rounding of k*z and truncation (1e-10 is pure noise here)
    for tag, which, eps in (("dist", (0, 1, 2), 1e-6), ("phase", (3, 4), 1e-6)):
        v = [rng.standard_normal(np.shape(p)) if i in which else np.zeros(np.shape(p)) for i, p in enumerate(p4)]
        arrays["dloss_" + tag] = (run4([p + eps * d for p, d in zip(p4, v)])[1] - run4([p - eps * d for p, d in zip(p4, v)])[1]) / (2 * eps)
        for i, d in enumerate(v):
            arrays["v_%s_%d" % (tag, i)] = d
    save("four_f_n32", **arrays)


if __name__ == "__main__":
    if "--tables-only" in sys.argv:           # adds the element/table fixtures without redrawing the propagator ones
        rng = np.random.default_rng(20261018)
        element_cases()
        if "--elements-only" not in sys.argv:
            table_cases()
        sys.exit(0)
    rs_cases()
    vrs_cases()
    czt_cases()
    highna_cases()
