"""Host-buffer drop-in for the reference's Cython extension module
`model.monotonic_align.model.monotonic_align.core` (core.pyx).

    maximum_path_c(paths, values, t_xs, t_ys, max_neg_val=-1e9)      core.pyx:40

Same argument meaning: numpy int32 `paths` [B,Tx,Ty] filled with {0,1},
float32 `values` [B,Tx,Ty], int32 `t_xs`, `t_ys` [B], all C-contiguous host
arrays.  Differences, both deliberate: `values` is NOT clobbered (the reference
accumulates in place, core.pyx:30) and `paths` need not be pre-zeroed.
The search runs on the GPU through mas_b200_maximum_path_host.
"""
from __future__ import annotations

import ctypes

import numpy as np

from .. import _lib


def _chk(a, dtype, ndim, name):
    if not (isinstance(a, np.ndarray) and a.dtype == dtype and a.ndim == ndim and a.flags.c_contiguous):
        raise TypeError(f"{name}: need a C-contiguous {np.dtype(dtype).name} ndarray with ndim={ndim}")


def maximum_path_c(paths, values, t_xs, t_ys, max_neg_val: float = _lib.MAX_NEG_VAL) -> int:
    _chk(paths, np.int32, 3, "paths")
    _chk(values, np.float32, 3, "values")
    _chk(t_xs, np.int32, 1, "t_xs")
    _chk(t_ys, np.int32, 1, "t_ys")
    B, Tx, Ty = values.shape
    if paths.shape != values.shape or t_xs.shape != (B,) or t_ys.shape != (B,):
        raise ValueError("shape mismatch")
    rc = _lib.lib().mas_b200_maximum_path_host(paths.ctypes.data, values.ctypes.data, t_xs.ctypes.data,
                                               t_ys.ctypes.data, B, Tx, Ty, ctypes.c_float(max_neg_val))
    if rc < 0:
        raise _lib.MasB200Error(rc, "mas_b200_maximum_path_host")
    return rc      # number of rejected items (t_x > t_y etc.), 0 normally
