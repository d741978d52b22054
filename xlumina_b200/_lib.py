"""
ctypes binding of libxlprop.so (C ABI: include/xlprop.h).

The library is built in-tree by `__graft_entry__.build()` / `python -m xlumina_b200.build`.  There is no fallback: if the
shared object is missing, or a tensor is not on a CUDA device, the propagators raise.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# XLPROP_LIB: development only -- an alternative CUDA build of the same sources (an experiment variant made by
# `python -m xlumina_b200.build --exp ...`) for A/B timing; it is still the C ABI of include/xlprop.h on the GPU.
LIB_PATH = os.environ.get("XLPROP_LIB") or os.path.join(_HERE, "libxlprop.so")

_vp = ctypes.c_void_p
_i = ctypes.c_int
_d = ctypes.c_double
_sz = ctypes.c_size_t
_ll = ctypes.c_longlong

# name -> (restype, argtypes); mirrors include/xlprop.h one to one
SIGNATURES = {
    "xl_version": (_i, []),
    "xl_last_error": (ctypes.c_char_p, []),
    "xl_rs_padded_length": (_i, [_i]),
    "xl_czt_padded_length": (_i, [_i, _i]),
    "xl_rs_transfer_bytes": (_sz, [_i]),
    "xl_rs_workspace_bytes": (_sz, [_i, _i, _i]),
    "xl_rs_transfer": (_i, [_vp, _vp, _i, _d, _d, _d, _i, _vp]),
    "xl_rs_transfer_multi": (_i, [_vp, _sz, _vp, _i, _i, _i, _d, _d, _d, _i, _vp]),
    "xl_rs_fwd": (_i, [_vp, _vp, _vp, _vp, _i, _i, _d, _d, _d, _i, _vp, _sz, _vp]),
    "xl_rs_bwd": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _d, _d, _d, _i, _vp, _sz, _vp]),
    "xl_vrs_fwd": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _d, _d, _d, _d, _d, _i, _vp, _sz, _vp]),
    "xl_vrs_bwd": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _d, _d, _d, _d, _d, _i, _vp, _sz, _vp]),
    "xl_rs_fwd_fused": (_i, [_vp, _vp, _vp, _vp, _i, _i, _d, _d, _d, _i, _vp, _vp, _sz, _vp]),
    "xl_rs_bwd_fused": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _d, _d, _d, _i, _vp, _vp, _sz, _vp]),
    "xl_slab_padded_length": (_i, [_i]),
    "xl_slab_h_rows_per_rank": (_i, [_i, _i]),
    "xl_slab_scratch_bytes": (_sz, [_i, _i]),
    "xl_slab_h_rows": (_i, [_vp, _vp, _i, _i, _i, _d, _d, _d, _vp, _vp]),
    "xl_slab_h_rows_dz": (_i, [_vp, _vp, _i, _i, _i, _d, _d, _d, _vp, _vp]),
    "xl_slab_h_cols": (_i, [_vp, _vp, _i, _i, _d, _d, _vp]),
    "xl_slab_rows_fwd": (_i, [_vp, _vp, _i, _i, _i, _vp]),
    "xl_slab_cols": (_i, [_vp, _vp, _i, _i, _vp, _vp]),
    "xl_slab_rows_inv": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp]),
    "xl_debug_set_max_line": (None, [_i]),
    "xl_debug_set_long_cluster": (None, [_i]),
    "xl_czt_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "xl_czt_tables_bytes": (_sz, [_i, _i, _i]),
    "xl_czt_fwd": (_i, [_vp, _vp, _vp, _vp, _d, _i, _i, _i, _i, _d, _d, _d, _d, _d, _d, _d, _d, _i, _vp, _vp, _sz, _vp]),
    "xl_czt_bwd": (_i, [_vp, _vp, _vp, _d, _i, _i, _i, _i, _d, _d, _d, _d, _d, _d, _d, _d, _i, _vp, _vp, _sz, _vp]),
    "xl_czt_workspace_bytes_z": (_sz, [_i, _i, _i, _i]),
    "xl_czt_bwd_z": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _d, _i, _i, _i, _i, _d, _d, _d, _d, _d, _d, _d, _d, _i, _vp, _vp, _sz, _vp]),
    "xl_highna_workspace_bytes": (_sz, [_i, _i, _i]),
    "xl_highna_tables_bytes": (_sz, [_i, _i, _i]),
    "xl_highna_fwd": (_i, [_vp, _vp, _vp, _i, _i, _i, _d, _d, _d, _d, _d, _d, _d, _d, _d, _d, _d, _i, _vp, _vp, _sz, _vp]),
    "xl_rs_fwd_batch": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _d, _d, _d, _i, _vp, _sz, _vp]),
    "xl_rs_bwd_batch": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _vp, _vp, _i, _i, _i, _i, _d, _d, _d, _i, _vp, _sz, _vp]),
    "xl_vrs_fwd_batch": (_i, [_vp, _vp, _ll, _vp, _vp, _vp, _i, _i, _i, _d, _d, _d, _d, _d, _i, _vp, _sz, _vp]),
    "xl_vrs_bwd_batch": (_i, [_vp, _vp, _ll, _vp, _vp, _vp, _vp, _i, _vp, _vp, _i, _i, _i, _d, _d, _d, _d, _d, _i, _vp, _sz, _vp]),
    "xl_czt_workspace_bytes_batch": (_sz, [_i, _i, _i, _i, _i]),
    "xl_czt_fwd_batch": (_i, [_vp, _vp, _ll, _vp, _vp, _i, _d, _i, _i, _i, _i, _i, _d, _d, _d, _d, _d, _d, _d, _d, _i, _vp, _vp, _sz, _vp]),
    "xl_czt_bwd_batch": (_i, [_vp, _vp, _vp, _i, _d, _i, _i, _i, _i, _i, _d, _d, _d, _d, _d, _d, _d, _d, _i, _vp, _vp, _sz, _vp]),
    "xl_highna_fwd_batch": (_i, [_vp, _vp, _ll, _vp, _i, _i, _i, _i, _d, _d, _d, _d, _d, _d, _d, _d, _d, _d, _d, _i, _vp, _vp, _sz, _vp]),
    "xl_highna_bwd_batch": (_i, [_vp, _vp, _i, _i, _i, _i, _d, _d, _d, _d, _d, _d, _d, _d, _d, _d, _d, _i, _vp, _vp, _sz, _vp]),
    "xl_el_scratch_bytes": (_sz, []),
    "xl_el_sslm": (_i, [_vp, _vp, _vp, _vp, _d, _d, _vp, _vp, _sz, _vp]),
    "xl_el_sslm_bwd": (_i, [_vp, _vp, _vp, _vp, _d, _d, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "xl_el_lcd": (_i, [_vp, _vp, _vp, _vp, _d, _d, _vp, _vp, _sz, _vp]),
    "xl_el_lcd_bwd": (_i, [_vp, _vp, _vp, _vp, _d, _d, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "xl_el_bs": (_i, [_vp, _vp, _vp, _vp, _vp, _d, _d, _vp, _vp, _vp, _vp, _sz, _vp]),
    "xl_el_bs_bwd": (_i, [_vp, _vp, _vp, _vp, _vp, _d, _d, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "xl_launch_count": (ctypes.c_longlong, []),
    "xl_prof_enable": (None, [_i]),
    "xl_prof_report": (_i, [ctypes.c_char_p, _i]),
    "xl_highna_bwd": (_i, [_vp, _vp, _i, _i, _i, _d, _d, _d, _d, _d, _d, _d, _d, _d, _d, _d, _i, _vp, _vp, _sz, _vp]),
}

XL_CONJ_IN = 1
XL_CONJ_OUT = 2
XL_REUSE_H = 16
XL_REUSE_TABLES = 32
XL_PHASE_BLIND = 64
XL_WITH_HZ = 128


class RsFuse(ctypes.Structure):
    """struct xl_rs_fuse of include/xlprop.h (pointwise elements fused into the scalar RS path)."""
    _fields_ = [("mod", ctypes.c_void_p), ("in_real", ctypes.c_int), ("target", ctypes.c_void_p), ("mse", ctypes.c_void_p),
                ("Hz", ctypes.c_void_p)]


class XlpropError(RuntimeError):
    pass


def declare(cdll):
    """Attach restype/argtypes for every exported symbol; raises AttributeError if one is missing."""
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(cdll, name)
        fn.restype = res
        fn.argtypes = args
    return cdll


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise XlpropError(
                f"{LIB_PATH} not found: build the CUDA extension first (python -c 'import __graft_entry__ as g; g.build()'). "
                "xlumina_b200 has no CPU fallback.")
        _lib = declare(ctypes.CDLL(LIB_PATH))
    return _lib


def check(rc, what=""):
    if rc != 0:
        msg = lib().xl_last_error().decode("utf-8", "replace") if _lib is not None else ""
        raise XlpropError(f"{what} failed with code {rc}: {msg}")
