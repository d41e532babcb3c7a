"""ctypes binding of libstk.so (include/stk.h).

There is no CPU fallback: if the library is missing or cannot be loaded, every
product entry point raises.  Build it with `python -m
spacetime_fullgrid_parallel_b200.build` (or `__graft_entry__.build()`).
"""
import ctypes
import os

from .build import LIB

_c = ctypes
_vp, _int, _i64, _dbl = _c.c_void_p, _c.c_int, _c.c_int64, _c.c_double

# name -> (restype, argtypes); mirrors include/stk.h one to one.
SIGNATURES = {
    'stk_version': (_int, []),
    'stk_last_error': (_c.c_char_p, []),
    'stk_sync': (_int, [_vp]),
    'stk_launch_count': (_i64, []),
    'stk_block_from_rowmajor': (_int, [_vp, _int, _int, _vp, _int, _vp]),
    'stk_block_to_rowmajor': (_int, [_vp, _int, _int, _int, _vp, _vp]),
    'stk_block_upload_host': (_int, [_vp, _int, _int, _vp, _int, _vp, _vp]),
    'stk_block_download_host': (_int,
                                [_vp, _int, _int, _int, _vp, _vp, _vp]),
    'stk_axpy': (_int, [_dbl, _vp, _vp, _i64, _vp]),
    'stk_scale': (_int, [_dbl, _vp, _i64, _vp]),
    'stk_xpay': (_int, [_vp, _dbl, _vp, _i64, _vp]),
    'stk_pcg_update': (_int, [_dbl, _vp, _vp, _vp, _vp, _i64, _vp]),
    'stk_xpay_dev': (_int, [_vp, _vp, _vp, _vp, _i64, _vp]),
    'stk_pcg_update_dev': (_int, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp]),
    'stk_dot': (_int, [_vp, _vp, _i64, _vp, _vp, _vp]),
    'stk_space_spmm': (_int, [
        _int, _vp, _vp, _int, _vp, _vp, _vp, _vp, _vp, _dbl, _dbl, _vp, _vp,
        _int, _vp
    ]),
    'stk_csr_set_row_order': (_int, [_vp, _int, _vp]),
    'stk_time_apply': (_int, [
        _int, _int, _int, _vp, _vp, _vp, _vp, _int, _int, _vp, _int, _dbl,
        _dbl, _vp, _int, _vp
    ]),
    'stk_time_apply2': (_int, [
        _int, _int, _vp, _vp, _vp, _vp, _vp, _int, _int, _vp, _int, _vp, _dbl,
        _dbl, _vp, _int, _int, _vp
    ]),
    'stk_time_tridiag_pair': (_int, [
        _int, _int, _int, _vp, _vp, _vp, _int, _vp, _vp, _vp, _vp, _vp, _vp,
        _int, _vp
    ]),
    'stk_space_spmm_split': (_int, [
        _int, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _int, _vp
    ]),
    'stk_space_spmm_pair': (_int, [
        _int, _vp, _vp, _vp, _vp, _vp, _vp, _int, _dbl, _dbl, _vp, _vp, _int,
        _vp
    ]),
    'stk_copy_cols': (_int, [_int, _int, _vp, _int, _int, _vp, _vp, _int, _int,
                             _vp]),
    'stk_outer': (_int, [_int, _int, _vp, _vp, _vp, _vp]),
    'stk_pack_slices': (_int, [_vp, _int, _int, _vp, _int, _vp, _vp]),
    'stk_unpack_slices': (_int,
                          [_vp, _int, _int, _vp, _int, _vp, _dbl, _dbl, _vp]),
    'stk_wavelet_lift': (_int, [_int, _int, _int, _vp, _vp, _int, _vp]),
    'stk_time_chain': (_int, [
        _int, _int, _int, _int, _vp, _vp, _vp, _vp, _vp, _int, _int, _vp, _int,
        _vp, _vp, _int, _vp, _vp
    ]),
    'stk_mg_create': (_vp, [_int, _int, _int, _int]),
    'stk_mg_destroy': (None, [_vp]),
    'stk_mg_set_level': (_int, [
        _vp, _int, _int, _int, _vp, _vp, _vp, _vp, _vp, _vp, _int
    ]),
    'stk_mg_set_transfer': (_int, [_vp, _int, _vp, _vp, _vp, _vp, _vp, _vp]),
    'stk_mg_workspace': (_i64, [_vp, _int]),
    'stk_mg_apply': (_int, [_vp, _vp, _vp, _vp, _vp, _int, _vp, _vp]),
    'stk_mg_smooth': (_int, [_vp, _int, _int, _int, _vp, _vp, _vp, _int, _vp]),
    'stk_gs_wavefronts': (_int, [_int, _vp, _vp, _vp]),
    'stk_gs_alloc_slots': (_int, [_int, _vp, _vp, _vp]),
    'stk_gs_prog_create': (_vp, [
        _int, _int, _int, _int, _int, _int, _vp, _vp, _vp, _vp, _vp
    ]),
    'stk_gs_prog_destroy': (None, [_vp]),
    'stk_mg_set_fused': (_int, [_vp, _int, _vp, _vp, _vp, _int, _int, _vp,
                                _int, _int, _vp, _vp]),
    'stk_mg_set_fused_wide': (_int, [_vp, _int, _vp, _vp, _int]),
    'stk_gs_fused': (_int, [
        _vp, _int, _int, _vp, _int, _int, _vp, _i64, _vp, _vp, _vp, _vp, _int,
        _vp
    ]),
}

_lib = None


class StkError(RuntimeError):
    pass


def lib():
    """The loaded library; raises if it was not built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            raise StkError(
                'libstk.so is not built (%s): run `python -m '
                'spacetime_fullgrid_parallel_b200.build`; there is no CPU '
                'fallback' % LIB)
        handle = ctypes.CDLL(LIB)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)  # AttributeError if not exported
            fn.restype, fn.argtypes = res, args
        _lib = handle
    return _lib


def check(rc):
    if rc != 0:
        raise StkError('libstk error %d: %s' %
                       (rc, lib().stk_last_error().decode()))


def stream():
    import torch
    return torch.cuda.current_stream().cuda_stream


def ptr(t):
    """Device (or host) address of a torch tensor / numpy array, or NULL."""
    if t is None:
        return None
    if hasattr(t, 'data_ptr'):
        return t.data_ptr()
    return t.ctypes.data
