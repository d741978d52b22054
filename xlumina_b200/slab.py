"""
Slab-decomposed scalar RS propagation of ONE N x N field over the G GPUs of a box (BASELINE.json cfg 5).

Rank g owns field rows [g N/G, (g+1) N/G).  The 2-D FFT convolution of xlumina/wave_optics.py:281-297 becomes
    row FFTs of the local rows  ->  ALL-TO-ALL  ->  column convolution on this rank's (L/2)/G slot pairs  ->  ALL-TO-ALL
    ->  inverse row FFTs of the local rows
and the transfer function is built the same way (local y rows of the analytic impulse response -> all-to-all -> column FFTs
of this rank's slot pairs), so no rank ever holds a full padded plane.  The stages are the C-ABI entry points
`xl_slab_*` (include/xlprop.h); the exchanges are `torch.distributed.all_to_all_single` (NCCL over NVLink/NVSwitch on GPUs;
gloo in the CPU tests, which drive the host-emulated kernel bodies through the same code).

The exchanged layouts are chosen so that every peer's chunk is contiguous on both sides and the second exchange is the
exact inverse of the first: row side [L/2 pairs][rows][2], column side [source rank][(L/2)/G pairs][rows][2].

Padded lengths up to 4096 (N <= 2048) run the single-pass FFT kernels; longer lines, up to 32768 (N <= 16384, the 16384^2
configuration), are split into R <= 8 sub-lines of 4096 (csrc/xl_long.cuh: the forward radix-R step is fused into the
loads, the inverse one is a pointwise combine kernel over a scratch buffer).  With one rank the same chain is the
single-GPU path for grids above 2048^2.  Scalar fields; forward, field-VJP (the operator is complex-symmetric:
`rs_slab_vjp`) and d/dz (`rs_slab_grad_z`: one more transfer-function slab -- of the reduced kernel dh/dz - i k h -- and one
more forward chain on the primal input; the fused single-GPU path for N <= 2048 gets d/dz from a Parseval sum instead).
"""
import contextlib
import ctypes

import torch
import torch.distributed as dist

from . import _lib

__all__ = ["rs_propagation_slab", "rs_slab_vjp", "rs_slab_grad_z", "SlabPlan"]


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr())


def _device_of(t):
    """Context that makes the tensor's CUDA device current (the library keys its tables on the current device)."""
    return torch.cuda.device(t.device) if t.is_cuda else contextlib.nullcontext()


def _require_device(t, lib):
    """The product library takes CUDA tensors only (no CPU fallback); only the host-emulation library of the tests, passed
    explicitly as `lib`, runs on CPU tensors."""
    if lib is _lib._lib and not t.is_cuda:
        raise _lib.XlpropError("xlumina_b200 operators need CUDA tensors (no CPU fallback)")


def _stream_of(t):
    return ctypes.c_void_p(torch.cuda.current_stream(t.device).cuda_stream) if t.is_cuda else None


class SlabPlan:
    """Geometry of one slab decomposition (all sizes in complex64 elements)."""

    def __init__(self, N, world, lib):
        self.N, self.G = int(N), int(world)
        self.L = lib.xl_slab_padded_length(self.N)
        if not self.L:
            raise _lib.XlpropError(f"slab RS: N={N} unsupported (padded length must be <= 32768)")
        if self.N % (2 * self.G) or (self.L // 2) % self.G:
            raise _lib.XlpropError(f"slab RS: N={N} must be a multiple of 2*G and L/2={self.L // 2} a multiple of G={world}")
        self.rows = self.N // self.G                    # field rows per rank
        self.pairs = (self.L // 2) // self.G            # x-slot pairs per rank
        self.hrows = lib.xl_slab_h_rows_per_rank(self.N, self.G)
        self.spec_elems = (self.L // 2) * self.rows * 2
        self.hspec_elems = (self.L // 2) * self.hrows * 2
        self.hloc_elems = self.pairs * self.L * 2
        self.scratch_bytes = int(lib.xl_slab_scratch_bytes(self.N, self.G))


_LOCAL = object()   # group sentinel: single-rank chain even inside an initialised multi-rank job (no collective at all)


def _world_rank(group):
    if group is _LOCAL or not dist.is_initialized():
        return 1, 0
    return dist.get_world_size(group), dist.get_rank(group)


def _all_to_all(buf, group):
    """Equal-split all-to-all of a complex64 buffer whose G peer chunks are contiguous; returns the received buffer."""
    if _world_rank(group)[0] == 1:
        return buf            # one rank: the exchanged layout [1][pairs][rows][2] IS the row-side layout
    out = torch.empty_like(buf)
    dist.all_to_all_single(torch.view_as_real(out).reshape(-1), torch.view_as_real(buf).reshape(-1), group=group)
    return out


def _all_to_all_start(buf, group):
    """The same exchange, issued asynchronously: returns (received buffer, work handle or None).  The collective runs on the
    backend's own stream behind everything already queued on the current stream; kernels launched after this call overlap
    with it, and `_wait` makes the current stream wait for the exchanged data."""
    if _world_rank(group)[0] == 1:
        return buf, None
    out = torch.empty_like(buf)
    work = dist.all_to_all_single(torch.view_as_real(out).reshape(-1), torch.view_as_real(buf).reshape(-1), group=group, async_op=True)
    return out, work


def _wait(work):
    if work is not None:
        work.wait()


def _scratch(plan, ref):
    return torch.empty(plan.scratch_bytes, dtype=torch.uint8, device=ref.device)


def _transfer_slab(plan, z, dx, dy, k, rank, ref, lib, group, deriv=False):
    R = torch.empty(plan.hspec_elems, dtype=torch.complex64, device=ref.device)
    scr = _scratch(plan, ref)
    rows = lib.xl_slab_h_rows_dz if deriv else lib.xl_slab_h_rows      # deriv: the reduced kernel dh/dz - i k h
    _lib.check(rows(_ptr(R), _ptr(z), plan.N, plan.G, rank, dx, dy, k, _ptr(scr), _stream_of(ref)), "xl_slab_h_rows")
    Th = _all_to_all(R, group)
    H = torch.empty(plan.hloc_elems, dtype=torch.complex64, device=ref.device)
    _lib.check(lib.xl_slab_h_cols(_ptr(Th), _ptr(H), plan.N, plan.G, dx, dy, _stream_of(ref)), "xl_slab_h_cols")
    return H


def _apply(plan, field_local, H, flags, lib, group):
    S = torch.empty(plan.spec_elems, dtype=torch.complex64, device=field_local.device)
    st = _stream_of(field_local)
    _lib.check(lib.xl_slab_rows_fwd(_ptr(field_local), _ptr(S), plan.N, plan.G, flags, st), "xl_slab_rows_fwd")
    T = _all_to_all(S, group)
    scr = _scratch(plan, field_local)
    _lib.check(lib.xl_slab_cols(_ptr(T), _ptr(H), plan.N, plan.G, _ptr(scr), st), "xl_slab_cols")
    S2 = _all_to_all(T, group)
    out = torch.empty_like(field_local)
    _lib.check(lib.xl_slab_rows_inv(_ptr(S2), _ptr(out), plan.N, plan.G, flags, _ptr(scr), st), "xl_slab_rows_inv")
    return out


def _transfer_and_apply(plan, field_local, z, dx, dy, k, rank, lib, group):
    """Fresh distance on several ranks: the transfer-function chain and the field chain interleaved so that each of the
    first two exchanges has independent work behind it -- the row FFTs of the field run while the row spectra of the impulse
    response are exchanged, the column FFTs of the transfer function while the field's row spectra are.  Same stages, same
    buffers and the same results as _transfer_slab followed by _apply."""
    dev, st = field_local.device, _stream_of(field_local)
    R = torch.empty(plan.hspec_elems, dtype=torch.complex64, device=dev)
    scr = _scratch(plan, field_local)
    _lib.check(lib.xl_slab_h_rows(_ptr(R), _ptr(z), plan.N, plan.G, rank, dx, dy, k, _ptr(scr), st), "xl_slab_h_rows")
    Th, w_h = _all_to_all_start(R, group)
    S = torch.empty(plan.spec_elems, dtype=torch.complex64, device=dev)
    _lib.check(lib.xl_slab_rows_fwd(_ptr(field_local), _ptr(S), plan.N, plan.G, 0, st), "xl_slab_rows_fwd")
    T, w_s = _all_to_all_start(S, group)
    _wait(w_h)
    H = torch.empty(plan.hloc_elems, dtype=torch.complex64, device=dev)
    _lib.check(lib.xl_slab_h_cols(_ptr(Th), _ptr(H), plan.N, plan.G, dx, dy, st), "xl_slab_h_cols")
    _wait(w_s)
    _lib.check(lib.xl_slab_cols(_ptr(T), _ptr(H), plan.N, plan.G, _ptr(scr), st), "xl_slab_cols")
    S2 = _all_to_all(T, group)
    out = torch.empty_like(field_local)
    _lib.check(lib.xl_slab_rows_inv(_ptr(S2), _ptr(out), plan.N, plan.G, 0, _ptr(scr), st), "xl_slab_rows_inv")
    return out, H


def rs_propagation_slab(field_local, z, dx, dy, k, group=None, lib=None, transfer=None, return_transfer=False):
    """Propagate the row slab `field_local` (N/G, N) complex64 of an N x N field by z; every rank of `group` calls this with
    its own slab and the same z, dx, dy, k (`group=slab._LOCAL`: this process alone, whole field, no collective).  Returns this rank's rows of the result (and the transfer-function slab when
    `return_transfer`, to be passed back as `transfer` for another field or the VJP at the same z)."""
    lib = lib or _lib.lib()
    world, rank = _world_rank(group)
    _require_device(field_local, lib)
    f = field_local.to(torch.complex64).resolve_conj().contiguous()
    N = f.shape[-1]
    plan = SlabPlan(N, world, lib)
    if f.shape != (plan.rows, N):
        raise ValueError(f"slab RS: this rank's slab must be ({plan.rows}, {N}), got {tuple(f.shape)}")
    with _device_of(f):
        if transfer is None:
            zt = z if isinstance(z, torch.Tensor) else torch.full((1,), float(z), dtype=torch.float64, device=f.device)
            zt = zt.to(device=f.device, dtype=torch.float64).reshape(1)
            out, transfer = _transfer_and_apply(plan, f, zt, float(dx), float(dy), float(k), rank, lib, group)
        else:
            out = _apply(plan, f, transfer, 0, lib, group)
    return (out, transfer) if return_transfer else out


def rs_slab_vjp(ct_local, transfer, group=None, lib=None):
    """Field VJP (JAX convention: plain transpose) of rs_propagation_slab at the same z: the operator is complex-symmetric,
    so it is the forward chain on the cotangent slab with the saved transfer-function slab."""
    lib = lib or _lib.lib()
    world, _ = _world_rank(group)
    c = ct_local.to(torch.complex64).resolve_conj().contiguous()
    plan = SlabPlan(c.shape[-1], world, lib)
    with _device_of(c):
        return _apply(plan, c, transfer, 0, lib, group)


def rs_slab_grad_z(field_local, ct_local, out_local, z, dx, dy, k, group=None, lib=None, transfer_dz=None, return_transfer=False):
    """d/dz of  Re sum(ct * out)  for  out = rs_propagation_slab(field, z)  -- `ct_local` is this rank's slab of the cotangent
    in the JAX convention (no conjugate; torch callers pass conj(grad_output)), `out_local` the saved primal output.
        d out/dz = i k out + field (*) (dh/dz - i k h)
    The first term is a pure phase rotation: its contribution, -k Im sum(ct * out), is accumulated in float64 from complex64
    operands (every product is exact), so it cancels exactly for intensity-type losses; only the
    reduced kernel, 1e2-1e4 times smaller, goes through the complex64 FFT chain.  Returns a float64 scalar tensor; with
    more than one rank the partial sums are all-reduced (every rank gets the total).  `transfer_dz` / `return_transfer`: the
    transfer-function slab of the reduced kernel, to be reused for the other fields of a batch at the same z."""
    lib = lib or _lib.lib()
    world, rank = _world_rank(group)
    _require_device(field_local, lib)
    f = field_local.to(torch.complex64).resolve_conj().contiguous()
    c = ct_local.to(torch.complex64).resolve_conj().contiguous()
    o = out_local.to(torch.complex64).resolve_conj().contiguous()
    plan = SlabPlan(f.shape[-1], world, lib)
    with _device_of(f):
        zt = z if isinstance(z, torch.Tensor) else torch.full((1,), float(z), dtype=torch.float64, device=f.device)
        zt = zt.detach().to(device=f.device, dtype=torch.float64).reshape(1)
        Hz = transfer_dz
        if Hz is None:
            Hz = _transfer_slab(plan, zt, float(dx), float(dy), float(k), rank, f, lib, group, deriv=True)
        d = _apply(plan, f, Hz, 0, lib, group)
        gz = torch.zeros((), dtype=torch.float64, device=f.device)
        step = max(1, (1 << 24) // f.shape[-1])                       # rows per chunk: bounds the float64 temporaries
        for r0 in range(0, f.shape[0], step):
            cc = c[r0:r0 + step].to(torch.complex128)
            gz += (cc * d[r0:r0 + step]).real.sum() - float(k) * (cc * o[r0:r0 + step]).imag.sum()
    if world > 1:
        dist.all_reduce(gz, group=group)
    return (gz, Hz) if return_transfer else gz
