"""Drop-in for the reference package `model/monotonic_align`.

    from face_gan_tts_b200 import monotonic_align
    attn = monotonic_align.maximum_path(log_prior, attn_mask.squeeze(1))      # face_tts.py:173

Same name, arguments and return contract as reference
model/monotonic_align/__init__.py:8-23: `path` has value's shape, dtype and
device and holds exactly {0,1}.  The work is done by libmas_b200.so on the
GPU (no D2H/H2D bounce, no host sync); CPU tensors are copied to the current
CUDA device, aligned there and copied back -- the GPU still does the search,
there is no CPU implementation.
"""
from __future__ import annotations

import torch

from .. import alignment as _al
from . import core  # noqa: F401  (maximum_path_c, the host-buffer entry point)


def maximum_path(value: torch.Tensor, mask: torch.Tensor, *, check: bool = False) -> torch.Tensor:
    """value: [b, t_x, t_y]; mask: [b, t_x, t_y] prefix (rectangular) mask.
    Lengths are read from the mask exactly like the reference (column 0 / row 0,
    __init__.py:20-21).  The reference's `value * mask` (:13) only zeroes cells the
    search never reads, so it is skipped.  Non-differentiable, like the reference
    (called under no_grad and detached, face_tts.py:165,174)."""
    if value.dim() != 3 or mask.shape != value.shape:
        raise ValueError("value and mask must both be [b, t_x, t_y]")
    device, dtype = value.device, value.dtype
    if not torch.cuda.is_available():
        raise RuntimeError("maximum_path needs a CUDA device: this library has no CPU fallback")
    v = value.detach()
    m = mask.detach()
    if not v.is_cuda:
        v = v.cuda()
    if m.device != v.device:
        m = m.to(v.device)
    if v.dtype != torch.float32:
        v = v.to(torch.float32)          # the reference also computes in float32 (:16)
    t_x, t_y = _al.lengths_from_mask(m)
    res = _al.align(v, t_x, t_y, dense_path=True, path_dtype=torch.float32, check=check)
    path = res.path
    if dtype != torch.float32:
        path = path.to(dtype)
    if path.device != device:
        path = path.to(device)
    return path


def maximum_path_from_lengths(value: torch.Tensor, t_x, t_y, **kw) -> "_al.AlignmentResult":
    """Richer entry point: explicit lengths in, (path, durations, frame_token, status) out."""
    return _al.align(value, t_x, t_y, **kw)
