"""`jax.numpy` stand-in: numpy, in float64/complex128, plus the functional `.at[idx].set(v)` update."""
import numpy as _np
from numpy import *  # noqa: F401,F403
from numpy import fft, linalg  # noqa: F401

pi = _np.pi
complex128 = _np.complex128
complex64 = _np.complex64


class _Setter:
    def __init__(self, arr, idx):
        self.arr, self.idx = arr, idx

    def set(self, value):
        out = _np.array(self.arr, copy=True).view(_Arr)
        out[self.idx] = value
        return out


class _At:
    def __init__(self, arr):
        self.arr = arr

    def __getitem__(self, idx):
        return _Setter(self.arr, idx)


class _Arr(_np.ndarray):
    @property
    def at(self):
        return _At(self)


def zeros(shape, dtype=float):
    return _np.zeros(shape, dtype=dtype).view(_Arr)


def ones(shape, dtype=float):
    return _np.ones(shape, dtype=dtype).view(_Arr)


def array(obj, dtype=None):
    return _np.array(obj, dtype=dtype).view(_Arr)


def fromfunction(function, shape, dtype=float, **kw):
    """jax.numpy.fromfunction vmaps `function` over the index axes, so the callee sees SCALAR indices (the reference relies on
    this to build (N, N, 2, 2) Jones tensors, optical_elements.py:208); numpy's version would pass whole index arrays."""
    idx = _np.indices(shape).reshape(len(shape), -1).astype(dtype)
    vals = [_np.asarray(function(*[ix[j] for ix in idx], **kw)) for j in range(idx.shape[1])]
    return _np.stack(vals).reshape(tuple(shape) + vals[0].shape).view(_Arr)
