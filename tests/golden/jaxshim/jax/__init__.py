"""Minimal NumPy-backed stand-in for the parts of `jax` that XLuminA's propagation path touches.
TEST TOOLING ONLY (used by tests/golden/make_golden.py to execute the reference's own source files; real JAX is not
installable in this image).  jit = identity, vmap = Python loop, jnp = numpy with an `.at[...].set()` shim."""
import numpy as _np
from . import numpy  # noqa: F401  (jax.numpy)


class _Config:
    def update(self, *a, **k):
        pass


config = _Config()


def jit(fun=None, static_argnums=None, static_argnames=None, **kw):
    if fun is None:
        return lambda f: f
    return fun


def vmap(fun, in_axes=0, out_axes=0):
    def wrapped(*args):
        axes = in_axes if isinstance(in_axes, (tuple, list)) else (in_axes,) * len(args)
        n = None
        for a, ax in zip(args, axes):
            if ax is not None:
                n = _np.shape(a)[ax]
                break
        outs = []
        for i in range(n):
            call = [(_np.take(a, i, axis=ax) if ax is not None else a) for a, ax in zip(args, axes)]
            outs.append(fun(*call))
        if isinstance(outs[0], tuple):
            return tuple(_np.stack([o[j] for o in outs], axis=out_axes) for j in range(len(outs[0])))
        return _np.stack(outs, axis=out_axes)
    return wrapped


class lax:
    @staticmethod
    def cond(pred, true_fun, false_fun, *operands):
        return true_fun(*operands) if pred else false_fun(*operands)


class nn:
    @staticmethod
    def logsumexp(a, axis=None):
        a = _np.asarray(a)
        m = _np.max(a, axis=axis, keepdims=True)
        out = _np.log(_np.sum(_np.exp(a - m), axis=axis, keepdims=True)) + m
        return out.reshape(()) if axis is None else _np.squeeze(out, axis=axis)


class random:
    @staticmethod
    def PRNGKey(seed):
        return _np.random.default_rng(seed)

    @staticmethod
    def uniform(key, shape=(), minval=0.0, maxval=1.0):
        return key.uniform(minval, maxval, size=shape)
