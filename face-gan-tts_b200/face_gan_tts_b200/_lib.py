"""ctypes binding of libmas_b200.so (include/mas_b200.h).

The library is the product: if it is missing this module raises, loudly --
there is no Python/CPU fallback for any compute entry point.
"""
from __future__ import annotations

import ctypes
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_PKG)
LIB_PATH = os.environ.get("MAS_B200_LIB", os.path.join(_ROOT, "lib", "libmas_b200.so"))

# status codes (include/mas_b200.h)
OK, ERR_ARG, ERR_UNSUPPORTED, ERR_WORKSPACE, ERR_CUDA, ERR_ALIGN = 0, -1, -2, -3, -4, -5
ITEM_OK, ITEM_BAD_LENGTH = 0, 1
PATH_NONE, PATH_F32, PATH_I32 = 0, 1, 2
LP_AUTO, LP_FFMA, LP_TCGEN05 = 0, 1, 2
WS_PREPARED = 0x100
MAX_NEG_VAL = -1e9

c_int, c_float, c_size_t, c_void_p, c_char_p, c_ll = (
    ctypes.c_int, ctypes.c_float, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_char_p, ctypes.c_longlong)

# name -> (restype, argtypes): every symbol include/mas_b200.h declares
SIGNATURES = {
    "mas_b200_abi_version": (c_int, []),
    "mas_b200_error_string": (c_char_p, [c_int]),
    "mas_b200_last_cuda_error": (c_int, []),
    "mas_b200_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "mas_b200_lengths_from_mask": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "mas_b200_maximum_path": (c_int, [c_void_p, c_ll, c_ll, c_void_p, c_void_p, c_int, c_int, c_int, c_float,
                                      c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "mas_b200_log_prior": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_void_p]),
    "mas_b200_packed_batch_bytes": (c_size_t, [c_void_p, c_void_p, c_int, c_int]),
    "mas_b200_pack_batch_host": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_size_t]),
    "mas_b200_unpack_batch": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "mas_b200_set_pointer_option": (c_int, [c_char_p, c_void_p]),
    "mas_b200_debug_occupy_sms": (c_int, [c_int, c_ll, c_void_p]),
    "mas_b200_fused_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "mas_b200_fused_workspace_prepare": (c_int, [c_void_p, c_size_t, c_int, c_int, c_int, c_int, c_void_p]),
    "mas_b200_log_prior_maximum_path": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                                                c_float, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                                c_size_t, c_int, c_void_p]),
    "mas_b200_put_durations": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_void_p]),
    "mas_b200_generate_path": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_void_p]),
    "mas_b200_generate_path_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_void_p,
                                           c_void_p]),
    "mas_b200_sequence_mask": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "mas_b200_crop_frames": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p,
                                     c_void_p, c_void_p, c_void_p, c_void_p]),
    "mas_b200_gather_mu_y": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "mas_b200_gather_mu_y_backward": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                                              c_int, c_void_p, c_void_p]),
    "mas_b200_prior_loss_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "mas_b200_prior_loss": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p,
                                    c_void_p, c_void_p, c_size_t, c_void_p]),
    "mas_b200_prior_loss_backward": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                             c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "mas_b200_duration_loss": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p,
                                       c_void_p]),
    "mas_b200_upload_batch": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p,
                                      c_void_p, c_void_p, c_void_p, c_void_p]),
    "mas_b200_maximum_path_host": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_float]),
    "mas_b200_log_prior_maximum_path_host": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                                                     c_int, c_float, c_void_p, c_void_p, c_void_p]),
    "mas_b200_set_option": (c_int, [c_char_p, c_int]),
    "mas_b200_get_option": (c_int, [c_char_p]),
}

_lib = None


class MasB200Error(RuntimeError):
    def __init__(self, status, where):
        self.status = status
        msg = lib().mas_b200_error_string(status).decode()
        if status == ERR_CUDA:
            msg += f" [cudaError {lib().mas_b200_last_cuda_error()}]"
        super().__init__(f"{where}: {msg} (status {status})")


def lib():
    """Load libmas_b200.so.  Raises if it has not been built: the CUDA library IS the
    implementation (build it with `python face-gan-tts_b200/build.py`)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: build the sm_100a library with "
                f"`python face-gan-tts_b200/build.py` (there is no CPU fallback)")
        l = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(l, name)          # AttributeError if the .so does not export it
            fn.restype = res
            fn.argtypes = args
        if l.mas_b200_abi_version() != 1:
            raise RuntimeError("libmas_b200.so ABI version mismatch")
        _lib = l
    return _lib


def check(status, where):
    if status != OK:
        raise MasB200Error(status, where)


def set_option(key: str, value: int) -> int:
    prev = lib().mas_b200_set_option(key.encode(), int(value))
    if prev == -2 ** 31:
        raise KeyError(key)
    return prev


def set_pointer_option(key: str, tensor) -> None:
    """Diagnostics / tests: hand a device buffer (torch tensor, or None to switch off) to the library."""
    check(lib().mas_b200_set_pointer_option(key.encode(), None if tensor is None else tensor.data_ptr()),
          "mas_b200_set_pointer_option")


def get_option(key: str) -> int:
    v = lib().mas_b200_get_option(key.encode())
    if v == -2 ** 31:
        raise KeyError(key)
    return v
