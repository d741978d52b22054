"""
Differentiable propagation operators: torch.autograd.Function wrappers over the C ABI (include/xlprop.h).

These are the torch-side equivalents of the reference's jitted seam functions (paths relative to the XLuminA repo):
  rs_propagation      <- RS_propagation_jit          xlumina/wave_optics.py:281-289
  vrs_propagation     <- VRS_propagation_jit         xlumina/vectorized_optics.py:364-373 (+ Ez formation :258-261)
  czt / vczt          <- CZT_jit / VCZT_jit          xlumina/wave_optics.py:333-357, vectorized_optics.py:375-384
  highna_focus        <- high_NA_objective_lens + vectorized_CZT_for_high_NA + cte   xlumina/optical_elements.py:515-638

torch is plumbing only (device memory, streams, autograd graph); all arithmetic happens in libxlprop.so.  Inputs must live
on a CUDA device; there is no CPU path.  complex128 inputs are cast to complex64 on entry and back on exit (the reference
runs in x64; the kernels are complex64 with fp64 phase generation).
"""
import contextlib
import ctypes
import functools
import os

import torch

from . import _lib

__all__ = ["rs_propagation", "rs_propagation_fused", "vrs_propagation", "czt", "vczt", "highna_focus", "rs_transfer", "set_transfer_cache",
           "el_sslm", "el_lcd", "el_bs", "elements_on"]


# ---------------------------------------------------------------------------------------------- plumbing
def _require_device(t):
    if not t.is_cuda:
        raise _lib.XlpropError("xlumina_b200 operators need CUDA tensors (no CPU fallback)")


def _on_device(fn):
    """Run an autograd.Function forward / backward with the CUDA device of its first tensor argument current: the library
    uses the current device for its per-device tables and kernel attributes, the stream comes from the tensor's device."""
    @functools.wraps(fn)
    def wrapped(ctx, *args):
        t = next((a for a in args if isinstance(a, torch.Tensor)), None)
        guard = torch.cuda.device(t.device) if t is not None and t.is_cuda else contextlib.nullcontext()
        with guard:
            return fn(ctx, *args)
    return wrapped


def _stream(t):
    return ctypes.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def _stream_key(t):
    return (t.device, torch.cuda.current_stream(t.device).cuda_stream)


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


_workspaces = {}


def _workspace(ref, nbytes):
    """Per-(device, stream) scratch buffer, grown on demand; stream-ordered reuse makes sharing it between calls safe."""
    key = _stream_key(ref)
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty(max(int(nbytes), 1 << 20), dtype=torch.uint8, device=ref.device)
        _workspaces[key] = ws
    return ws


_z_cache = {}


def _z_per_item(z, batch):
    """None if `z` is one distance shared by the whole batch; else the list of `batch` per-item distances (the reference
    reaches this case by vmapping a table over per-sample distances, e.g. noisy optical tables)."""
    if not isinstance(z, torch.Tensor) or z.numel() == 1:
        return None
    if z.numel() != batch:
        raise ValueError(f"z has {z.numel()} elements for a batch of {batch} fields: pass one distance or one per field")
    zf = z.reshape(-1)
    return [zf[i:i + 1] for i in range(batch)]


def _refuse_z_gradient(z, what):
    """Paths whose d/dz does not exist yet must not drop it silently (the reference differentiates through everything)."""
    if isinstance(z, torch.Tensor) and z.requires_grad and torch.is_grad_enabled():
        raise _lib.XlpropError(f"{what}: the gradient with respect to the distance z is not implemented; "
                               "pass z.detach() if the distance is not being optimised")


def _as_z(z, ref):
    """Propagation distance as ONE float64 on the device (a traced value in the reference, wave_optics.py:281)."""
    if isinstance(z, torch.Tensor):
        if z.numel() != 1:
            raise ValueError("z must have exactly one element here (shape () or (1,))")
        return z.to(device=ref.device, dtype=torch.float64).reshape(1)
    key = (float(z), ref.device)
    t = _z_cache.get(key)
    if t is None:
        if len(_z_cache) > 4096:
            _z_cache.clear()
        # torch.full is a device-side fill: no pageable host-to-device copy, so the host never waits for the stream
        t = torch.full((1,), float(z), dtype=torch.float64, device=ref.device)
        _z_cache[key] = t
    return t


def _c64(t):
    if not torch.is_complex(t):
        t = t.to(torch.complex64)
    elif t.dtype != torch.complex64:
        t = t.to(torch.complex64)
    return t.resolve_conj().contiguous()      # a lazily conjugated view shares storage with the unconjugated data


_grid_cache = {}


def _grid(coords):
    """(first coordinate, spacing, last coordinate, n) of a uniformly spaced 1-D grid given as numpy/torch/list.
    The uniformity check is cached per array object (the hot loop calls this with the same x/y arrays every step)."""
    import numpy as np
    key = id(coords)
    hit = _grid_cache.get(key)
    if hit is not None and hit[0] is coords:
        return hit[1]
    a = coords.detach().cpu().numpy() if isinstance(coords, torch.Tensor) else np.asarray(coords, dtype=np.float64)
    n = int(a.shape[0])
    first, last = float(a[0]), float(a[-1])
    step = (last - first) / (n - 1)
    if n > 2:
        d = np.diff(a)
        if np.max(np.abs(d - step)) > 1e-6 * abs(step):
            raise ValueError("xlumina_b200 regenerates coordinate grids analytically: grids must be uniformly spaced")
    res = (first, step, last, n)
    if len(_grid_cache) > 256:
        _grid_cache.clear()
    _grid_cache[key] = (coords, res)     # keeps the array alive, so the id stays valid
    return res


# ---------------------------------------------------------------------------------------------- transfer-function cache
# Optical tables propagate several beams over the SAME distance inside one loss evaluation (the reference's
# hybrid_setup_sharp_focus uses z1_1 + z1_2 three times, optical_elements.py:1609,1622; the 6x6 ansatz uses each of
# z8..z12 six times): SURVEY.md 8f-3.  With the cache on, calls that pass the same z OBJECT (same tensor, unmodified, or
# the same Python float) on the same grid reuse one transfer function instead of regenerating it (XL_REUSE_H).
# Off by default: every entry pins xl_rs_transfer_bytes(N) of device memory (128 MiB at N = 2048; twice that when z needs a
# gradient, because the reduced dH/dz is generated with H).  Memory model with the cache off: every RS / VRS autograd node
# keeps its own H for its backward call (16 VRS nodes of the sharp-focus table at N = 1024: 16 x 32 MiB, x2 with d/dz).
# Caveat: a tensor distance is keyed on (id(z), z._version) -- an update that bypasses the version counter (writing through
# z.data) would reuse a stale transfer function; optimizers that update parameters in place (torch.optim) bump the version.
import collections

_transfer_cache = collections.OrderedDict()
_transfer_cache_size = 0


def set_transfer_cache(entries):
    """Keep the transfer functions of the last `entries` distinct (z, grid) combinations; 0 disables and frees the cache."""
    global _transfer_cache_size
    _transfer_cache_size = max(0, int(entries))
    while len(_transfer_cache) > _transfer_cache_size:
        _transfer_cache.popitem(last=False)


def _z_key(z):
    return ("t", id(z), z._version) if isinstance(z, torch.Tensor) else ("f", float(z))


def _cached_transfer(zkey, zobj, N, dx, dy, k, device, nbytes, with_hz=0):
    """(H, reuse): H is a device buffer for the transfer function (with_hz: followed by the reduced dH/dz of the same
    distance), reuse says whether it already holds it."""
    if _transfer_cache_size == 0 or zkey is None:
        return torch.empty(nbytes, dtype=torch.uint8, device=device), False
    key = (zkey, N, dx, dy, k, device, torch.cuda.current_stream(device).cuda_stream, with_hz)
    hit = _transfer_cache.get(key)
    if hit is not None:
        _transfer_cache.move_to_end(key)
        return hit[1], True
    H = torch.empty(nbytes, dtype=torch.uint8, device=device)
    _transfer_cache[key] = (zobj, H)          # zobj is kept alive so that id(z) cannot be recycled while the entry lives
    while len(_transfer_cache) > _transfer_cache_size:
        _transfer_cache.popitem(last=False)
    return H, False


# ---------------------------------------------------------------------------------------------- RS / VRS
class _RS(torch.autograd.Function):
    """field (F,N,N) c64, z (1,) f64 -> (F,N,N).  backward = same complex-symmetric operator + Parseval d/dz."""

    @staticmethod
    @_on_device
    def forward(ctx, field, z, dx, dy, k, zkey=None, zobj=None):
        _require_device(field)
        L = _lib.lib()
        F, N = field.shape[0], field.shape[-1]
        out = torch.empty_like(field)
        # z needs a gradient: H and the reduced dH/dz are generated together, in one launch pair (XL_WITH_HZ)
        hz = _lib.XL_WITH_HZ if z.requires_grad else 0
        H, reuse = _cached_transfer(zkey, zobj, N, dx, dy, k, field.device, L.xl_rs_transfer_bytes(N) * (2 if hz else 1), hz)
        need = L.xl_rs_workspace_bytes(N, F, 0)
        ws = _workspace(field, need)
        _lib.check(L.xl_rs_fwd(_ptr(field), _ptr(out), _ptr(H), _ptr(z), N, F, dx, dy, k, (_lib.XL_REUSE_H if reuse else 0) | hz,
                               _ptr(ws), ws.numel(), _stream(field)), "xl_rs_fwd")
        ctx.save_for_backward(field, z, H, out)      # out: the exact i*k*out part of d out/dz (include/xlprop.h)
        ctx.geom = (dx, dy, k)
        ctx.hz = hz
        return out

    @staticmethod
    @_on_device
    def backward(ctx, g):
        field, z, H, out = ctx.saved_tensors
        dx, dy, k = ctx.geom
        L = _lib.lib()
        F, N = field.shape[0], field.shape[-1]
        g = g.resolve_conj().contiguous()
        want_z = ctx.needs_input_grad[1]
        gin = torch.empty_like(field)
        gz = torch.zeros(1, dtype=torch.float64, device=field.device) if want_z else None
        need = L.xl_rs_workspace_bytes(N, F, 1 if want_z else 0)
        ws = _workspace(field, need)
        _lib.check(L.xl_rs_bwd(_ptr(field), _ptr(out), _ptr(g), _ptr(gin), _ptr(gz), _ptr(H), _ptr(z), N, F, dx, dy, k,
                               _lib.XL_CONJ_IN | _lib.XL_CONJ_OUT | ctx.hz, _ptr(ws), ws.numel(), _stream(field)), "xl_rs_bwd")
        return gin, gz, None, None, None, None, None


class _RSItems(torch.autograd.Function):
    """field (B,N,N) c64, z (B,) f64 -- one distance per item (what vmapping a table over noisy distances gives the reference,
    examples/noisy_optimization.ipynb cell 7) -> (B,N,N): ONE library call each way (xl_rs_fwd_batch / xl_rs_bwd_batch with
    z_stride = 1); the B transfer functions (and reduced dH/dz) are generated in one launch pair."""

    @staticmethod
    @_on_device
    def forward(ctx, field, z, dx, dy, k):
        _require_device(field)
        L = _lib.lib()
        B, N = field.shape[0], field.shape[-1]
        out = torch.empty_like(field)
        hz = _lib.XL_WITH_HZ if z.requires_grad else 0
        H = torch.empty(B * L.xl_rs_transfer_bytes(N) * (2 if hz else 1), dtype=torch.uint8, device=field.device)
        ws = _workspace(field, L.xl_rs_workspace_bytes(N, 1, 0))
        _lib.check(L.xl_rs_fwd_batch(_ptr(field), _ptr(out), _ptr(H), _ptr(z), 1, N, 1, B, dx, dy, k, hz, _ptr(ws), ws.numel(), _stream(field)),
                   "xl_rs_fwd_batch")
        ctx.save_for_backward(field, z, H, out)
        ctx.geom = (dx, dy, k, hz)
        return out

    @staticmethod
    @_on_device
    def backward(ctx, g):
        field, z, H, out = ctx.saved_tensors
        dx, dy, k, hz = ctx.geom
        L = _lib.lib()
        B, N = field.shape[0], field.shape[-1]
        g = g.resolve_conj().contiguous()
        want_z = ctx.needs_input_grad[1]
        gin = torch.empty_like(field)
        gz = torch.zeros(B, dtype=torch.float64, device=field.device) if want_z else None
        ws = _workspace(field, L.xl_rs_workspace_bytes(N, 1, 1 if want_z else 0))
        _lib.check(L.xl_rs_bwd_batch(_ptr(field), _ptr(out), _ptr(g), _ptr(gin), _ptr(gz), 1, _ptr(H), _ptr(z), 1, N, 1, B, dx, dy, k,
                                     _lib.XL_CONJ_IN | _lib.XL_CONJ_OUT | hz, _ptr(ws), ws.numel(), _stream(field)), "xl_rs_bwd_batch")
        return gin, gz, None, None, None


class _VRSBatch(torch.autograd.Function):
    """A batch of vectorial RS propagations in ONE library call each way (xl_vrs_fwd_batch / xl_vrs_bwd_batch): `ex` is
    (B,N,N) with `ey` (B,N,N), or a stacked (B,2,N,N) array with ey = None; z is one shared distance (1,) -- one transfer
    function for the batch -- or one per item (B,).  Returns (B,3,N,N)."""

    @staticmethod
    @_on_device
    def forward(ctx, ex, ey, z, x0, y0, dx, dy, k):
        _require_device(ex)
        L = _lib.lib()
        B, N = ex.shape[0], ex.shape[-1]
        zstr = 0 if z.numel() == 1 else 1
        out = torch.empty((B, 3, N, N), dtype=ex.dtype, device=ex.device)
        hz = _lib.XL_WITH_HZ if z.requires_grad else 0
        H = torch.empty((B if zstr else 1) * L.xl_rs_transfer_bytes(N) * (2 if hz else 1), dtype=torch.uint8, device=ex.device)
        bstride = N * N if ey is not None else 2 * N * N
        ws = _workspace(ex, L.xl_rs_workspace_bytes(N, 3, 0))
        _lib.check(L.xl_vrs_fwd_batch(_ptr(ex), _ptr(ey), bstride, _ptr(out), _ptr(H), _ptr(z), zstr, N, B, x0, y0, dx, dy, k, hz,
                                      _ptr(ws), ws.numel(), _stream(ex)), "xl_vrs_fwd_batch")
        ctx.save_for_backward(ex, ey, z, H, out)
        ctx.geom = (x0, y0, dx, dy, k, hz, zstr, bstride)
        return out

    @staticmethod
    @_on_device
    def backward(ctx, g):
        ex, ey, z, H, out = ctx.saved_tensors
        x0, y0, dx, dy, k, hz, zstr, bstride = ctx.geom
        L = _lib.lib()
        B, N = ex.shape[0], ex.shape[-1]
        g = g.resolve_conj().contiguous()
        want_z = ctx.needs_input_grad[2]
        gin = torch.empty((B, 2, N, N), dtype=ex.dtype, device=ex.device)
        gz = torch.zeros(B if zstr else 1, dtype=torch.float64, device=ex.device) if want_z else None
        ws = _workspace(ex, L.xl_rs_workspace_bytes(N, 3, 1 if want_z else 0))
        _lib.check(L.xl_vrs_bwd_batch(_ptr(ex), _ptr(ey), bstride, _ptr(out), _ptr(g), _ptr(gin), _ptr(gz), zstr, _ptr(H), _ptr(z), zstr, N, B,
                                      x0, y0, dx, dy, k, _lib.XL_CONJ_IN | _lib.XL_CONJ_OUT | hz, _ptr(ws), ws.numel(), _stream(ex)),
                   "xl_vrs_bwd_batch")
        return ((gin, None) if ey is None else (gin[:, 0], gin[:, 1])) + (gz, None, None, None, None, None)


class _VRS(torch.autograd.Function):
    """(Ex, Ey) (N,N) each -> (3,N,N); Ez = (Ex X + Ey Y)/r formed at load (vectorized_optics.py:258-261).  The two planes are
    passed to the library where they live: no stacking copy."""

    @staticmethod
    @_on_device
    def forward(ctx, ex, ey, z, x0, y0, dx, dy, k, zkey=None, zobj=None, hshare=None):
        _require_device(ex)
        L = _lib.lib()
        N = ex.shape[-1]
        out = torch.empty((3, N, N), dtype=ex.dtype, device=ex.device)
        hz = _lib.XL_WITH_HZ if z.requires_grad else 0         # H and the reduced dH/dz generated together
        if hshare is not None and hshare[0] is not None:       # later item of a batch that shares z: reuse its transfer function
            H, reuse = hshare[0], True
        else:
            H, reuse = _cached_transfer(zkey, zobj, N, dx, dy, k, ex.device, L.xl_rs_transfer_bytes(N) * (2 if hz else 1), hz)
            if hshare is not None:
                hshare[0] = H
        ws = _workspace(ex, L.xl_rs_workspace_bytes(N, 3, 0))
        _lib.check(L.xl_vrs_fwd(_ptr(ex), _ptr(ey), _ptr(out), _ptr(H), _ptr(z), N, x0, y0, dx, dy, k, (_lib.XL_REUSE_H if reuse else 0) | hz,
                                _ptr(ws), ws.numel(), _stream(ex)), "xl_vrs_fwd")
        ctx.save_for_backward(ex, ey, z, H, out)
        ctx.geom = (x0, y0, dx, dy, k)
        ctx.hz = hz
        return out

    @staticmethod
    @_on_device
    def backward(ctx, g):
        ex, ey, z, H, out = ctx.saved_tensors
        x0, y0, dx, dy, k = ctx.geom
        L = _lib.lib()
        N = ex.shape[-1]
        g = g.resolve_conj().contiguous()
        want_z = ctx.needs_input_grad[2]
        gin = torch.empty((2, N, N), dtype=ex.dtype, device=ex.device)
        gz = torch.zeros(1, dtype=torch.float64, device=ex.device) if want_z else None
        ws = _workspace(ex, L.xl_rs_workspace_bytes(N, 3, 1 if want_z else 0))
        _lib.check(L.xl_vrs_bwd(_ptr(ex), _ptr(ey), _ptr(out), _ptr(g), _ptr(gin), _ptr(gz), _ptr(H), _ptr(z), N, x0, y0, dx, dy, k,
                                _lib.XL_CONJ_IN | _lib.XL_CONJ_OUT | ctx.hz, _ptr(ws), ws.numel(), _stream(ex)), "xl_vrs_bwd")
        return ((gin, None) if ey is None else (gin[0], gin[1])) + (gz, None, None, None, None, None, None, None, None)


class _RSFused(torch.autograd.Function):
    """Scalar RS with the bracketing pointwise elements folded into its first and last pass (xl_rs_fwd_fused /
    xl_rs_bwd_fused):  out[f] = RS(field[f] * mod; z), or -- with `target` -- the per-field MSE of
    |out|^2 against target intensities.  field (F,N,N) complex64 or float32, mod (N,N) complex64 or None, target (F,N,N)
    float32 or None."""

    @staticmethod
    @_on_device
    def forward(ctx, field, z, mod, target, dx, dy, k, phase_blind, pre):
        _require_device(field)
        L = _lib.lib()
        F, N = field.shape[0], field.shape[-1]
        ctx.phase_blind = bool(phase_blind)
        out = torch.empty((F, N, N), dtype=torch.complex64, device=field.device)
        mse = torch.zeros(F, dtype=torch.float64, device=field.device) if target is not None else None
        if pre is not None:       # (H, reduced dH/dz) of this z, generated ahead by rs_transfer_pairs
            H, Hz = pre
        else:
            H, Hz = torch.empty(L.xl_rs_transfer_bytes(N), dtype=torch.uint8, device=field.device), None
        ws = _workspace(field, L.xl_rs_workspace_bytes(N, F, 0))
        fuse = _lib.RsFuse(mod.data_ptr() if mod is not None else None, 0 if field.is_complex() else 1,
                           target.data_ptr() if target is not None else None, mse.data_ptr() if mse is not None else None, None)
        _lib.check(L.xl_rs_fwd_fused(_ptr(field), _ptr(out), _ptr(H), _ptr(z), N, F, dx, dy, k, _lib.XL_REUSE_H if pre is not None else 0,
                                     ctypes.byref(fuse), _ptr(ws), ws.numel(), _stream(field)), "xl_rs_fwd_fused")
        ctx.save_for_backward(field, z, H, out, mod, target, Hz)
        ctx.geom = (dx, dy, k)
        return mse if target is not None else out

    @staticmethod
    @_on_device
    def backward(ctx, g):
        field, z, H, out, mod, target, Hz = ctx.saved_tensors
        dx, dy, k = ctx.geom
        L = _lib.lib()
        F, N = field.shape[0], field.shape[-1]
        want_f, want_z, want_m = ctx.needs_input_grad[0], ctx.needs_input_grad[1], ctx.needs_input_grad[2] and mod is not None
        g = g.resolve_conj().contiguous()
        if target is not None:
            ct_out, ct_mse = None, g.to(torch.float64)
        else:
            ct_out, ct_mse = g, None
        gin = torch.empty((F, N, N), dtype=torch.complex64, device=field.device) if want_f else None
        gmod = torch.empty((N, N), dtype=torch.complex64, device=field.device) if want_m else None
        gz = torch.zeros(1, dtype=torch.float64, device=field.device) if want_z else None
        ws = _workspace(field, L.xl_rs_workspace_bytes(N, F, 1 if want_z else 0))
        fuse = _lib.RsFuse(mod.data_ptr() if mod is not None else None, 0 if field.is_complex() else 1,
                           target.data_ptr() if target is not None else None, None, Hz.data_ptr() if Hz is not None else None)
        _lib.check(L.xl_rs_bwd_fused(_ptr(field), _ptr(out), _ptr(ct_out), _ptr(ct_mse), _ptr(gin), _ptr(gmod), _ptr(gz), _ptr(H), _ptr(z),
                                     N, F, dx, dy, k, _lib.XL_CONJ_IN | _lib.XL_CONJ_OUT | (_lib.XL_PHASE_BLIND if ctx.phase_blind else 0),
                                     ctypes.byref(fuse), _ptr(ws), ws.numel(), _stream(field)), "xl_rs_bwd_fused")
        if want_f and not field.is_complex():
            gin = gin.real
        return gin, gz, gmod, None, None, None, None, None, None


def rs_transfer_pairs(zs, N, dx, dy, k, device):
    """The transfer function AND its reduced z-derivative for every distance in `zs` (floats or one-element tensors), all in
    ONE launch pair (xl_rs_transfer_multi) instead of two launches per propagation and pass: the fixed per-step cost of an
    optimizer whose distances are known when the step starts.  Returns [(H, Hz), ...] for rs_propagation_fused(pre=...)."""
    L = _lib.lib()
    ref = torch.empty(0, device=device)
    _require_device(ref)
    nb = int(L.xl_rs_transfer_bytes(N))
    stride = (nb + 255) // 256 * 256
    zcat = torch.cat([_as_z(z, ref).detach() for z in zs])
    buf = torch.empty(2 * len(zs) * stride, dtype=torch.uint8, device=device)
    with (torch.cuda.device(ref.device) if ref.is_cuda else contextlib.nullcontext()):
        _lib.check(L.xl_rs_transfer_multi(_ptr(buf), stride, _ptr(zcat), 2 * len(zs), 1, N, float(dx), float(dy), float(k), 0,
                                          _stream(ref)), "xl_rs_transfer_multi")
    return [(buf[2 * i * stride:2 * i * stride + nb], buf[(2 * i + 1) * stride:(2 * i + 1) * stride + nb]) for i in range(len(zs))]


def rs_propagation_fused(field, z, dx, dy, k, mod=None, target=None, phase_blind=False, pre=None):
    """Scalar RS propagation of a batch `field` (F,N,N) with the pointwise elements around it fused into the kernels
    (N <= 2048): every field is multiplied by the shared complex plane `mod` (N,N) while it is loaded -- a phase-only SLM
    exp(i phi), an amplitude mask, the beam under a batch of real object masks (then `field` may be float32) -- and, with
    `target` (F,N,N) float32, the detector is folded into the last pass: the call returns the F per-field values
    mean((|out|^2 - target)^2) (float64) instead of the fields.  Differentiable in field, z and mod.
    `phase_blind=True` is the caller's guarantee that the loss is invariant under a global phase of this propagation's output
    (a single path that ends in an intensity detector): the i*k*out part of d/dz, which is then exactly zero, is dropped
    instead of being evaluated as a complex64 cancellation residue.  `pre` = one entry of
    rs_transfer_pairs() for this z: the call then generates no transfer function itself."""
    N = field.shape[-1]
    if field.dim() != 3 or field.shape[-2] != N:
        raise ValueError("rs_propagation_fused needs a batch of square fields (F, N, N)")
    if N > FUSED_MAX_N:
        raise _lib.XlpropError(f"rs_propagation_fused: N <= {FUSED_MAX_N}")
    f = field.contiguous() if field.dtype == torch.float32 else _c64(field)
    m = None if mod is None else _c64(mod)
    t = None if target is None else target.to(torch.float32).contiguous()
    if m is not None and m.shape != (N, N):
        raise ValueError("mod must be one (N, N) plane shared by the batch")
    if t is not None and t.shape != f.shape:
        raise ValueError("target must have the shape of the batch")
    return _RSFused.apply(f, _as_z(z, f), m, t, float(dx), float(dy), float(k), bool(phase_blind), pre)


FUSED_MAX_N = 2048   # largest grid of the fused single-pass path (padded length 4096); above it: the stage chain of slab.py


class _RSLarge(torch.autograd.Function):
    """Grids above FUSED_MAX_N (up to 16384^2): the slab / split-line stage chain of xlumina_b200/slab.py with one rank.
    Differentiable in the field (the operator is complex-symmetric) and in z (slab.rs_slab_grad_z: the exact i k out term
    in float64 plus one more chain with the transfer function of the reduced kernel, per field)."""

    @staticmethod
    @_on_device
    def forward(ctx, field, z, dx, dy, k):
        from . import slab
        _require_device(field)
        outs, H = [], None
        for f in field:                                   # the fields of a batch share one transfer-function slab
            o, H = slab.rs_propagation_slab(f, z, dx, dy, k, transfer=H, return_transfer=True, group=slab._LOCAL)
            outs.append(o)
        out = torch.stack(outs)
        ctx.geo = (dx, dy, k)
        if z.requires_grad:
            ctx.save_for_backward(H, field, out, z)
        else:
            ctx.save_for_backward(H)
        return out

    @staticmethod
    @_on_device
    def backward(ctx, g):
        from . import slab
        H = ctx.saved_tensors[0]
        # torch convention: conj(A^T conj(g)); A^T = A
        gin = None
        if ctx.needs_input_grad[0]:
            gin = torch.stack([torch.conj_physical(slab.rs_slab_vjp(torch.conj_physical(gi), H, group=slab._LOCAL)) for gi in g])
        gz = None
        if ctx.needs_input_grad[1]:
            _, field, out, z = ctx.saved_tensors
            dx, dy, k = ctx.geo
            gz, Hz = None, None
            for i in range(g.shape[0]):                   # the fields of a batch share the slab of the reduced kernel
                gi, Hz = slab.rs_slab_grad_z(field[i], torch.conj_physical(g[i]), out[i], z, dx, dy, k, group=slab._LOCAL,
                                             transfer_dz=Hz, return_transfer=True)
                gz = gi if gz is None else gz + gi
            gz = gz.reshape(z.shape).to(z.dtype)
        return gin, gz, None, None, None



def rs_propagation(field, z, dx, dy, k):
    """Scalar Rayleigh-Sommerfeld propagation of `field` (..., N, N) over distance z (differentiable in field and z).  `z`: one distance shared by all the fields (one transfer function, one library call),
    or a tensor with one distance per field."""
    dt = field.dtype
    N = field.shape[-1]
    if field.shape[-2] != N:
        raise ValueError("RS propagation needs square fields")
    f = _c64(field).reshape(-1, N, N)
    zs = _z_per_item(z, f.shape[0])
    if zs is not None and N <= FUSED_MAX_N:  # one distance per field: ONE batched library call (every item has its own transfer function)
        zv = z.to(device=f.device, dtype=torch.float64).reshape(-1).contiguous()
        out = _RSItems.apply(f, zv, float(dx), float(dy), float(k)).reshape(field.shape)
        return out if dt == torch.complex64 or not torch.is_complex(field) else out.to(dt)
    if zs is not None:
        out = torch.stack([rs_propagation(f[i], zs[i], dx, dy, k) for i in range(f.shape[0])]).reshape(field.shape)
        return out if dt == torch.complex64 or not torch.is_complex(field) else out.to(dt)
    zt = _as_z(z, f)
    if N > FUSED_MAX_N:
        out = _RSLarge.apply(f, zt, float(dx), float(dy), float(k)).reshape(field.shape)
        return out if dt == torch.complex64 or not torch.is_complex(field) else out.to(dt)
    out = _RS.apply(f, zt, float(dx), float(dy), float(k), _z_key(z) if _transfer_cache_size else None, z).reshape(field.shape)
    return out if dt == torch.complex64 or not torch.is_complex(field) else out.to(dt)


def _planes(Ex, Ey):
    """The input of a vectorial operator as contiguous complex64: the caller's two (N,N) planes, or -- `Ex` a stacked (2,N,N)
    pair, Ey=None -- the pair itself with None (the library then takes Ey = Ex + N*N, and the autograd node has ONE input
    whose gradient is the (2,N,N) buffer the library fills: no per-plane select / accumulate kernels).  Never a copy."""
    if Ey is None:
        return _c64(Ex), None
    return _c64(Ex), _c64(Ey)


def _batch_of_pairs(fn, Ex, Ey, z):
    """Leading batch axis for the vectorial operators: Ex, Ey (B,N,N) or a stacked (B,2,N,N) -> (B,3,...); `z` shared
    or one per item.  Items are independent library calls; fn(ex, ey, z, hshare)."""
    B = Ex.shape[0]
    zs = _z_per_item(z, B)
    hshare = [None] if zs is None else None
    return torch.stack([fn(Ex[i], None if Ey is None else Ey[i], z if zs is None else zs[i], hshare) for i in range(B)])


def vrs_propagation(Ex, Ey, z, x0, y0, dx, dy, k, _hshare=None):
    """Vectorial RS: returns (3,N,N) = propagated [Ex, Ey, Ez].  Pass Ey=None if `Ex` is already the stacked (2,N,N) pair.
    With a leading batch axis (Ex, Ey (B,N,N) or stacked (B,2,N,N)) returns (B,3,N,N); z shared or one per item."""
    if Ex.dim() == (4 if Ey is None else 3):
        if Ex.shape[-1] <= FUSED_MAX_N:       # the whole batch in ONE library call (shared or per-item distances)
            dt = Ex.dtype
            ex, ey = _planes(Ex, Ey)
            B = ex.shape[0]
            if isinstance(z, torch.Tensor) and z.numel() > 1:
                if z.numel() != B:
                    raise ValueError(f"z has {z.numel()} elements for a batch of {B} fields: pass one distance or one per field")
                zt = z.to(device=ex.device, dtype=torch.float64).reshape(-1).contiguous()
            else:
                zt = _as_z(z, ex)
            out = _VRSBatch.apply(ex, ey, zt, float(x0), float(y0), float(dx), float(dy), float(k))
            return out if dt == torch.complex64 or not torch.is_complex(Ex) else out.to(dt)
        return _batch_of_pairs(lambda a, b, zz, hs: vrs_propagation(a, b, zz, x0, y0, dx, dy, k, hs), Ex, Ey, z)
    dt = Ex.dtype
    ex, ey = _planes(Ex, Ey)
    zt = _as_z(z, ex)
    N = ex.shape[-1]
    if N > FUSED_MAX_N:
        exy = ex if ey is None else torch.stack([ex, ey])
        # large grids: Ez = (Ex X + Ey Y)/r (vectorized_optics.py:258-261) is formed pointwise here and the three components
        # go through the stage chain as one batch sharing the transfer function; z stays on the autograd tape through r
        # (d Ez/dz) and through the chain (_RSLarge), so the route is differentiable in Ex, Ey and z
        xs = float(x0) + float(dx) * torch.arange(N, dtype=torch.float64, device=exy.device)
        ys = float(y0) + float(dy) * torch.arange(N, dtype=torch.float64, device=exy.device)
        r = torch.sqrt(xs[None, :] ** 2 + ys[:, None] ** 2 + zt ** 2)
        ez = exy[0] * (xs[None, :] / r).to(torch.float32) + exy[1] * (ys[:, None] / r).to(torch.float32)
        out = _RSLarge.apply(torch.stack([exy[0], exy[1], ez]), zt, float(dx), float(dy), float(k))
        return out if dt == torch.complex64 or not torch.is_complex(Ex) else out.to(dt)
    out = _VRS.apply(ex, ey, zt, float(x0), float(y0), float(dx), float(dy), float(k),
                     _z_key(z) if _transfer_cache_size else None, z, _hshare)
    return out if dt == torch.complex64 or not torch.is_complex(Ex) else out.to(dt)


def rs_transfer(z, N, dx, dy, k, device, deriv=False):
    """The transfer function buffer (opaque layout) for distance z -- exposed for caching / inspection."""
    L = _lib.lib()
    H = torch.empty(L.xl_rs_transfer_bytes(N), dtype=torch.uint8, device=device)
    _require_device(H)
    zt = _as_z(z, H)
    with torch.cuda.device(H.device):
        _lib.check(L.xl_rs_transfer(_ptr(H), _ptr(zt), N, float(dx), float(dy), float(k), 1 if deriv else 0, _stream(H)),
                   "xl_rs_transfer")
    return H


# ---------------------------------------------------------------------------------------------- CZT / VCZT / high-NA
class _CZT(torch.autograd.Function):
    @staticmethod
    @_on_device
    def forward(ctx, fin, ey, z, lam, vect, gin, gout):
        """fin: the scalar field, or Ex with `ey` = Ey (vect = 1)."""
        _require_device(fin)
        L = _lib.lib()
        N = fin.shape[-1]
        (x0, dx, y0, dy) = gin
        (xo0, xol, Mx, yo0, yol, My) = gout
        out = torch.empty((3, My, Mx) if vect else (My, Mx), dtype=fin.dtype, device=fin.device)
        need = L.xl_czt_workspace_bytes(N, Mx, My, vect)
        if need == 0:
            raise _lib.XlpropError("CZT: unsupported sizes (m+M-1 must not be a power of two; padded length <= 4096)")
        ws = _workspace(fin, need)
        # tables: Bluestein chirps / kernel spectra and the RS factor tables of this (z, grids); the backward pass reuses them
        tables = torch.empty(L.xl_czt_tables_bytes(N, Mx, My), dtype=torch.uint8, device=fin.device)
        _lib.check(L.xl_czt_fwd(_ptr(fin), _ptr(ey), _ptr(out), _ptr(z), lam, N, Mx, My, vect, x0, dx, y0, dy, xo0, xol, yo0, yol, 0,
                                _ptr(tables), _ptr(ws), ws.numel(), _stream(fin)), "xl_czt_fwd")
        want_z = ctx.needs_input_grad[2]
        ctx.save_for_backward(z, tables, *((fin, ey, out) if want_z else ()))   # d/dz needs the primal input and result
        ctx.meta = (lam, vect, gin, gout, N)
        ctx.stacked = bool(vect) and ey is None
        return out

    @staticmethod
    @_on_device
    def backward(ctx, g):
        z, tables = ctx.saved_tensors[:2]
        lam, vect, gin, gout, N = ctx.meta
        (x0, dx, y0, dy) = gin
        (xo0, xol, Mx, yo0, yol, My) = gout
        L = _lib.lib()
        g = g.resolve_conj().contiguous()
        ct = torch.empty((2, N, N) if vect else (N, N), dtype=g.dtype, device=g.device)
        if vect and not ctx.stacked:
            grads = lambda gz: (ct[0], ct[1], gz, None, None, None, None)
        else:
            grads = lambda gz: (ct, None, gz, None, None, None, None)
        if ctx.needs_input_grad[2]:
            fin, ey, out = ctx.saved_tensors[2:]
            gz = torch.zeros(1, dtype=torch.float64, device=g.device)
            ws = _workspace(g, L.xl_czt_workspace_bytes_z(N, Mx, My, vect))
            _lib.check(L.xl_czt_bwd_z(_ptr(fin), _ptr(ey), _ptr(out), _ptr(g), _ptr(ct), _ptr(gz), _ptr(z), lam, N, Mx, My, vect,
                                      x0, dx, y0, dy, xo0, xol, yo0, yol,
                                      _lib.XL_CONJ_IN | _lib.XL_CONJ_OUT | _lib.XL_REUSE_TABLES, _ptr(tables), _ptr(ws), ws.numel(),
                                      _stream(g)), "xl_czt_bwd_z")
            return grads(gz)
        ws = _workspace(g, L.xl_czt_workspace_bytes(N, Mx, My, vect))
        _lib.check(L.xl_czt_bwd(_ptr(g), _ptr(ct), _ptr(z), lam, N, Mx, My, vect, x0, dx, y0, dy, xo0, xol, yo0, yol,
                                _lib.XL_CONJ_IN | _lib.XL_CONJ_OUT | _lib.XL_REUSE_TABLES, _ptr(tables), _ptr(ws), ws.numel(),
                                _stream(g)), "xl_czt_bwd")
        return grads(None)


# The tables of the high-NA objective (Bluestein tables + lens matrix on the input grid) depend on static arguments only:
# every focusing of an optical table through the same objective shares one buffer (the sharp-focus table of BASELINE
# config 3 focuses six beams through one objective).  Keyed per device and stream; a handful of entries.
_highna_cache = collections.OrderedDict()


def _highna_tables(L, device, N, Mx, My, radius, f, lam, gin, gout):
    key = (device, torch.cuda.current_stream(device).cuda_stream if device.type == "cuda" else 0, N, Mx, My, radius, f, lam, gin, gout)
    hit = _highna_cache.get(key)
    if hit is not None:
        _highna_cache.move_to_end(key)
        return hit, True
    tables = torch.empty(L.xl_highna_tables_bytes(N, Mx, My), dtype=torch.uint8, device=device)
    _highna_cache[key] = tables
    while len(_highna_cache) > 4:
        _highna_cache.popitem(last=False)
    return tables, False


class _HighNA(torch.autograd.Function):
    @staticmethod
    @_on_device
    def forward(ctx, exy, ey, radius, f, lam, gin, gout):
        """exy: the Ex plane, ey: the Ey plane."""
        _require_device(exy)
        L = _lib.lib()
        N = exy.shape[-1]
        (x0, dx, y0, dy) = gin
        (xo0, xol, Mx, yo0, yol, My) = gout
        out = torch.empty((3, My, Mx), dtype=exy.dtype, device=exy.device)
        need = L.xl_highna_workspace_bytes(N, Mx, My)
        if need == 0:
            raise _lib.XlpropError("high-NA: unsupported sizes (m+M-1 must not be a power of two; padded length <= 4096)")
        ws = _workspace(exy, need)
        tables, reuse = _highna_tables(L, exy.device, N, Mx, My, radius, f, lam, gin, gout)
        _lib.check(L.xl_highna_fwd(_ptr(exy), _ptr(ey), _ptr(out), N, Mx, My, radius, f, lam, x0, dx, y0, dy, xo0, xol, yo0, yol,
                                   _lib.XL_REUSE_TABLES if reuse else 0, _ptr(tables), _ptr(ws), ws.numel(), _stream(exy)),
                   "xl_highna_fwd")
        ctx.save_for_backward(tables)
        ctx.meta = (radius, f, lam, gin, gout, N)
        ctx.stacked = ey is None
        return out

    @staticmethod
    @_on_device
    def backward(ctx, g):
        radius, f, lam, gin, gout, N = ctx.meta
        (tables,) = ctx.saved_tensors
        (x0, dx, y0, dy) = gin
        (xo0, xol, Mx, yo0, yol, My) = gout
        L = _lib.lib()
        g = g.resolve_conj().contiguous()
        ct = torch.empty((2, N, N), dtype=g.dtype, device=g.device)
        ws = _workspace(g, L.xl_highna_workspace_bytes(N, Mx, My))
        _lib.check(L.xl_highna_bwd(_ptr(g), _ptr(ct), N, Mx, My, radius, f, lam, x0, dx, y0, dy, xo0, xol, yo0, yol,
                                   _lib.XL_CONJ_IN | _lib.XL_CONJ_OUT | _lib.XL_REUSE_TABLES, _ptr(tables), _ptr(ws), ws.numel(),
                                   _stream(g)), "xl_highna_bwd")
        return ((ct, None) if ctx.stacked else (ct[0], ct[1])) + (None, None, None, None, None)


def _gout(xout, yout):
    xo0, _, xol, Mx = _grid(xout)
    yo0, _, yol, My = _grid(yout)
    return (xo0, xol, Mx, yo0, yol, My)


def _gin(x, y, N):
    x0, dx, _, nx = _grid(x)
    y0, dy, _, ny = _grid(y)
    if nx != N or ny != N:
        raise ValueError("CZT kernels need square N x N input fields with len(x) == len(y) == N")
    return (x0, dx, y0, dy)


def czt(field, z, wavelength, x, y, xout, yout):
    """Scalar chirped z-transform propagation (N,N) -> (len(yout), len(xout)); differentiable in `field` and `z`.
    A leading batch axis (B,N,N) is propagated item by item (z shared or one per item)."""
    if field.dim() == 3:
        zs = _z_per_item(z, field.shape[0])
        return torch.stack([czt(field[i], z if zs is None else zs[i], wavelength, x, y, xout, yout) for i in range(field.shape[0])])
    dt = field.dtype
    f = _c64(field)
    out = _CZT.apply(f, None, _as_z(z, f), float(wavelength), 0, _gin(x, y, f.shape[-1]), _gout(xout, yout))
    return out if dt == torch.complex64 or not torch.is_complex(field) else out.to(dt)


def vczt(Ex, Ey, z, wavelength, x, y, xout, yout):
    """Vectorial CZT: (Ex,Ey) -> (3, len(yout), len(xout)); Ez = ((Ex X + Ey Y)/r) z/r formed at load.
    Pass Ey=None if `Ex` is already the stacked (2,N,N) pair; a leading batch axis gives (B,3,...)."""
    if Ex.dim() == (4 if Ey is None else 3):
        return _batch_of_pairs(lambda a, b, zz, hs: vczt(a, b, zz, wavelength, x, y, xout, yout), Ex, Ey, z)
    dt = Ex.dtype
    ex, ey = _planes(Ex, Ey)
    out = _CZT.apply(ex, ey, _as_z(z, ex), float(wavelength), 1, _gin(x, y, ex.shape[-1]), _gout(xout, yout))
    return out if dt == torch.complex64 or not torch.is_complex(Ex) else out.to(dt)


def highna_focus(Ex, Ey, radius, f, wavelength, x, y, xout, yout):
    """High-NA objective + Debye integral by 2-pass Bluestein: (Ex,Ey) -> focal-plane (3, len(yout), len(xout)).
    Pass Ey=None if `Ex` is already the stacked (2,N,N) pair; a leading batch axis gives (B,3,...)."""
    if Ex.dim() == (4 if Ey is None else 3):
        return _batch_of_pairs(lambda a, b, zz, hs: highna_focus(a, b, radius, f, wavelength, x, y, xout, yout), Ex, Ey, None)
    dt = Ex.dtype
    ex, ey = _planes(Ex, Ey)
    out = _HighNA.apply(ex, ey, float(radius), float(f), float(wavelength), _gin(x, y, ex.shape[-1]), _gout(xout, yout))
    return out if dt == torch.complex64 or not torch.is_complex(Ex) else out.to(dt)


# ---------------------------------------------------------------------------------------------- pointwise Jones elements
# sSLM, LCD and BS_symmetric of a vectorial table as single-pass kernels (xl_el_*, include/xlprop.h): one
# launch forward, one backward, parameter maps and scalar-parameter reductions inside.  Planes are complex64; scalar
# parameters are float64 (1,) device tensors holding the optimizer's raw values (angle = scale * p + offset).
_cpu_kernels = False     # tests only: CPU tensors go to the host-emulated library (tests/test_ops_emu.py)


def elements_on(t):
    """True when the planes of `t`'s table go through the element kernels: complex64 on a CUDA device."""
    return isinstance(t, torch.Tensor) and t.dtype == torch.complex64 and (t.is_cuda or _cpu_kernels)


def _plane(t):
    return t.resolve_conj().contiguous()


def _f64_scalar(v, ref):
    if isinstance(v, torch.Tensor):
        return v.to(device=ref.device, dtype=torch.float64).reshape(1)
    return torch.full((1,), float(v), dtype=torch.float64, device=ref.device)


def _pair_or_none(a, b):
    """Cotangents of the two planes of one beam come together: both None, or a missing one replaced by zeros."""
    if a is None and b is None:
        return None, None
    a = torch.zeros_like(b) if a is None else _plane(a)
    b = torch.zeros_like(a) if b is None else _plane(b)
    return a, b


class _ElSSLM(torch.autograd.Function):
    @staticmethod
    @_on_device
    def forward(ctx, ex, ey, alpha, phi, scale, offset):
        _require_device(ex)
        L = _lib.lib()
        ox, oy = torch.empty_like(ex), torch.empty_like(ey)
        _lib.check(L.xl_el_sslm(_ptr(ex), _ptr(ey), _ptr(alpha), _ptr(phi), scale, offset, _ptr(ox), _ptr(oy), ex.numel(), _stream(ex)), "xl_el_sslm")
        ctx.save_for_backward(ex, ey, alpha, phi)
        ctx.map = (scale, offset)
        ctx.set_materialize_grads(False)
        return ox, oy

    @staticmethod
    @_on_device
    def backward(ctx, gox, goy):
        ex, ey, alpha, phi = ctx.saved_tensors
        scale, offset = ctx.map
        if gox is None and goy is None:
            return None, None, None, None, None, None
        L = _lib.lib()
        gox = None if gox is None else _plane(gox)
        goy = None if goy is None else _plane(goy)
        want_f = ctx.needs_input_grad[0] or ctx.needs_input_grad[1]
        gex = torch.empty_like(ex) if want_f else None
        gey = torch.empty_like(ey) if want_f else None
        ga = torch.empty_like(alpha) if ctx.needs_input_grad[2] else None
        gp = torch.empty_like(phi) if ctx.needs_input_grad[3] else None
        _lib.check(L.xl_el_sslm_bwd(_ptr(ex), _ptr(ey), _ptr(alpha), _ptr(phi), scale, offset, _ptr(gox), _ptr(goy), _ptr(gex), _ptr(gey),
                                    _ptr(ga), _ptr(gp), ex.numel(), _stream(ex)), "xl_el_sslm_bwd")
        return gex, gey, ga, gp, None, None


class _ElLCD(torch.autograd.Function):
    @staticmethod
    @_on_device
    def forward(ctx, ex, ey, eta, theta, scale, offset):
        _require_device(ex)
        L = _lib.lib()
        ox, oy = torch.empty_like(ex), torch.empty_like(ey)
        _lib.check(L.xl_el_lcd(_ptr(ex), _ptr(ey), _ptr(eta), _ptr(theta), scale, offset, _ptr(ox), _ptr(oy), ex.numel(), _stream(ex)), "xl_el_lcd")
        ctx.save_for_backward(ex, ey, eta, theta)
        ctx.map = (scale, offset)
        ctx.set_materialize_grads(False)
        return ox, oy

    @staticmethod
    @_on_device
    def backward(ctx, gox, goy):
        ex, ey, eta, theta = ctx.saved_tensors
        scale, offset = ctx.map
        gox, goy = _pair_or_none(gox, goy)
        if gox is None:
            return None, None, None, None, None, None
        L = _lib.lib()
        want_f = ctx.needs_input_grad[0] or ctx.needs_input_grad[1]
        gex = torch.empty_like(ex) if want_f else None
        gey = torch.empty_like(ey) if want_f else None
        ge = torch.zeros(1, dtype=torch.float64, device=ex.device) if ctx.needs_input_grad[2] else None
        gt = torch.zeros(1, dtype=torch.float64, device=ex.device) if ctx.needs_input_grad[3] else None
        ws = _workspace(ex, L.xl_el_scratch_bytes())
        _lib.check(L.xl_el_lcd_bwd(_ptr(ex), _ptr(ey), _ptr(eta), _ptr(theta), scale, offset, _ptr(gox), _ptr(goy), _ptr(gex), _ptr(gey),
                                   _ptr(ge), _ptr(gt), _ptr(ws), ex.numel(), _stream(ex)), "xl_el_lcd_bwd")
        return gex, gey, ge, gt, None, None


class _ElBS(torch.autograd.Function):
    @staticmethod
    @_on_device
    def forward(ctx, a_ex, a_ey, b_ex, b_ey, theta, scale, offset):
        _require_device(a_ex)
        L = _lib.lib()
        outs = [torch.empty_like(a_ex) for _ in range(4)]
        _lib.check(L.xl_el_bs(_ptr(a_ex), _ptr(a_ey), _ptr(b_ex), _ptr(b_ey), _ptr(theta), scale, offset,
                              _ptr(outs[0]), _ptr(outs[1]), _ptr(outs[2]), _ptr(outs[3]), a_ex.numel(), _stream(a_ex)), "xl_el_bs")
        ctx.save_for_backward(a_ex, a_ey, b_ex, b_ey, theta)
        ctx.map = (scale, offset)
        ctx.set_materialize_grads(False)
        return tuple(outs)

    @staticmethod
    @_on_device
    def backward(ctx, gcx, gcy, gdx, gdy):
        a_ex, a_ey, b_ex, b_ey, theta = ctx.saved_tensors
        scale, offset = ctx.map
        gcx, gcy = _pair_or_none(gcx, gcy)
        gdx, gdy = _pair_or_none(gdx, gdy)
        if gcx is None and gdx is None:
            return (None,) * 7
        L = _lib.lib()
        want_a = ctx.needs_input_grad[0] or ctx.needs_input_grad[1]
        want_b = ctx.needs_input_grad[2] or ctx.needs_input_grad[3]
        gax = torch.empty_like(a_ex) if want_a else None
        gay = torch.empty_like(a_ex) if want_a else None
        gbx = torch.empty_like(a_ex) if want_b else None
        gby = torch.empty_like(a_ex) if want_b else None
        gt = torch.zeros(1, dtype=torch.float64, device=a_ex.device) if ctx.needs_input_grad[4] else None
        ws = _workspace(a_ex, L.xl_el_scratch_bytes())
        _lib.check(L.xl_el_bs_bwd(_ptr(a_ex), _ptr(a_ey), _ptr(b_ex), _ptr(b_ey), _ptr(theta), scale, offset,
                                  _ptr(gcx), _ptr(gcy), _ptr(gdx), _ptr(gdy), _ptr(gax), _ptr(gay), _ptr(gbx), _ptr(gby), _ptr(gt),
                                  _ptr(ws), a_ex.numel(), _stream(a_ex)), "xl_el_bs_bwd")
        return gax, gay, gbx, gby, gt, None, None


def _phase_plane(v, ref):
    """A per-pixel phase parameter as a contiguous float32 plane on the device of `ref` (scalars are broadcast)."""
    if not isinstance(v, torch.Tensor):
        import numpy as np
        v = torch.as_tensor(np.asarray(v, dtype=np.float64))
    v = v.to(device=ref.device, dtype=torch.float32)
    return v.expand(ref.shape).contiguous()


def el_sslm(ex, ey, alpha, phi, scale=1.0, offset=0.0):
    """(ex e^{i(scale alpha + offset)}, ey e^{i(scale phi + offset)}): the super-SLM, optical_elements.py:186-222."""
    ex, ey = _plane(ex), _plane(ey)
    return _ElSSLM.apply(ex, ey, _phase_plane(alpha, ex), _phase_plane(phi, ey), float(scale), float(offset))


def el_lcd(ex, ey, eta, theta, scale=1.0, offset=0.0):
    """Uniform wave plate (retardance scale*eta + offset, fast axis at scale*theta + offset): optical_elements.py:266-305."""
    ex, ey = _plane(ex), _plane(ey)
    return _ElLCD.apply(ex, ey, _f64_scalar(eta, ex), _f64_scalar(theta, ex), float(scale), float(offset))


def el_bs(a_ex, a_ey, b_ex, b_ey, theta, scale=1.0, offset=0.0):
    """Lossy symmetric beam splitter, optical_elements.py:334-392: returns (c_ex, c_ey, d_ex, d_ey)."""
    a_ex, a_ey, b_ex, b_ey = _plane(a_ex), _plane(a_ey), _plane(b_ex), _plane(b_ey)
    return _ElBS.apply(a_ex, a_ey, b_ex, b_ey, _f64_scalar(theta, a_ex), float(scale), float(offset))
