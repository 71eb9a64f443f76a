"""PyTorch-facing host layer over libmas_b200.so.

Every function here only validates arguments, allocates outputs/workspace
with torch (device memory + streams are torch's job) and calls the C ABI on
`torch.cuda.current_stream()`.  Nothing synchronises the host unless asked
to (`check=True`).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Optional

import torch

from . import _lib

_IMPL = {"auto": _lib.LP_AUTO, "ffma": _lib.LP_FFMA, "tcgen05": _lib.LP_TCGEN05}


@dataclass
class AlignmentResult:
    """Outputs of one alignment call (all on the inputs' CUDA device).

    path         [B,Tx,Ty] {0,1} in the requested dtype, or None
    durations    [B,Tx] int32  frames per token (= path.sum(-1), reference face_tts.py:176)
    frame_token  [B,Ty] int32  token index of every frame, -1 beyond t_y
    status       [B]    int32  0 ok, 1 rejected (t_x > t_y, ...: undefined in the reference)
    """
    path: Optional[torch.Tensor]
    durations: torch.Tensor
    frame_token: torch.Tensor
    status: torch.Tensor


def _need_cuda(t: torch.Tensor, name: str):
    if not t.is_cuda:
        raise ValueError(f"{name} must be a CUDA tensor (this library has no CPU path)")


def _stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def _lengths(t, B, device, name):
    if not torch.is_tensor(t):
        t = torch.as_tensor(t)
    if t.shape != (B,):
        raise ValueError(f"{name} must have shape [{B}]")
    return t.to(device=device, dtype=torch.int32).contiguous()


def _path_dtype_code(dtype):
    if dtype == torch.float32:
        return _lib.PATH_F32
    if dtype == torch.int32:
        return _lib.PATH_I32
    raise ValueError("dense path dtype must be float32 or int32")


def _raise_on_bad(status: torch.Tensor, where: str):
    bad = torch.nonzero(status != 0).flatten().tolist()     # synchronises
    if bad:
        raise ValueError(f"{where}: items {bad} rejected (need 1 <= t_x <= t_y, t_x <= Tx, t_y <= Ty; "
                         f"the reference is undefined for these, core.pyx:34)")


def align(value: torch.Tensor, t_x, t_y, *, dense_path: bool = True, path_dtype=torch.float32,
          max_neg_val: float = _lib.MAX_NEG_VAL, check: bool = False) -> AlignmentResult:
    """Monotonic Alignment Search from explicit lengths.

    Replaces maximum_path_c (reference model/monotonic_align/core.pyx:40-45) and the
    D2H/H2D bounce around it.  `value` [B,Tx,Ty] float32 CUDA, unit stride along Ty,
    is not modified.  Bit-exact with the reference path for every defined input.
    """
    _need_cuda(value, "value")
    if value.dim() != 3 or value.dtype != torch.float32:
        raise ValueError("value must be a float32 [B,Tx,Ty] tensor")
    if value.stride(2) != 1 or value.stride(1) < value.shape[2]:
        value = value.contiguous()
    B, Tx, Ty = value.shape
    dev = value.device
    L = _lib.lib()
    with torch.cuda.device(dev):
        t_x = _lengths(t_x, B, dev, "t_x")
        t_y = _lengths(t_y, B, dev, "t_y")
        path = torch.empty((B, Tx, Ty), dtype=path_dtype, device=dev) if dense_path else None
        dur = torch.empty((B, Tx), dtype=torch.int32, device=dev)
        ft = torch.empty((B, Ty), dtype=torch.int32, device=dev)
        status = torch.empty((B,), dtype=torch.int32, device=dev)
        ws_bytes = L.mas_b200_workspace_bytes(B, Tx, Ty)
        ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev)
        rc = L.mas_b200_maximum_path(
            value.data_ptr(), value.stride(0), value.stride(1), t_x.data_ptr(), t_y.data_ptr(), B, Tx, Ty,
            max_neg_val, path.data_ptr() if dense_path else None,
            _path_dtype_code(path_dtype) if dense_path else _lib.PATH_NONE,
            dur.data_ptr(), ft.data_ptr(), status.data_ptr(), ws.data_ptr(), ws_bytes, _stream_ptr(dev))
        _lib.check(rc, "mas_b200_maximum_path")
        if check:
            _raise_on_bad(status, "align")
    return AlignmentResult(path, dur, ft, status)


def lengths_from_mask(mask: torch.Tensor):
    """t_x, t_y (int32 [B]) from a dense prefix mask [B,Tx,Ty], as reference
    model/monotonic_align/__init__.py:20-21: column 0 summed over x, row 0 summed over y."""
    _need_cuda(mask, "mask")
    B, Tx, Ty = mask.shape
    dev = mask.device
    if mask.dtype == torch.float32 and mask.is_contiguous():
        with torch.cuda.device(dev):
            t_x = torch.empty((B,), dtype=torch.int32, device=dev)
            t_y = torch.empty((B,), dtype=torch.int32, device=dev)
            rc = _lib.lib().mas_b200_lengths_from_mask(mask.data_ptr(), B, Tx, Ty, t_x.data_ptr(), t_y.data_ptr(),
                                                       _stream_ptr(dev))
            _lib.check(rc, "mas_b200_lengths_from_mask")
        return t_x, t_y
    # other dtypes / layouts: two thin strided reductions on the device (plumbing, not the hot path)
    t_x = mask[:, :, 0].to(torch.float32).sum(1).to(torch.int32)
    t_y = mask[:, 0, :].to(torch.float32).sum(1).to(torch.int32)
    return t_x, t_y


def log_prior(mu_x: torch.Tensor, y: torch.Tensor, impl: str = "auto") -> torch.Tensor:
    """Grad-TTS log-prior [B,Tx,Ty] (reference model/face_tts.py:165-171), fp32, within 1e-4
    relative of the torch expression.  mu_x [B,F,Tx], y [B,F,Ty] float32 CUDA."""
    _need_cuda(mu_x, "mu_x")
    _need_cuda(y, "y")
    if mu_x.dim() != 3 or y.dim() != 3 or mu_x.shape[:2] != y.shape[:2]:
        raise ValueError("mu_x [B,F,Tx] and y [B,F,Ty] must agree on B and F")
    mu_x = mu_x.detach().to(torch.float32).contiguous()
    y = y.detach().to(torch.float32).contiguous()
    B, F, Tx = mu_x.shape
    Ty = y.shape[2]
    dev = mu_x.device
    with torch.cuda.device(dev):
        out = torch.empty((B, Tx, Ty), dtype=torch.float32, device=dev)
        rc = _lib.lib().mas_b200_log_prior(mu_x.data_ptr(), y.data_ptr(), B, F, Tx, Ty, out.data_ptr(),
                                           _IMPL[impl], _stream_ptr(dev))
        _lib.check(rc, "mas_b200_log_prior")
    return out


def log_prior_maximum_path(mu_x: torch.Tensor, y: torch.Tensor, x_lengths, y_lengths, *,
                           dense_path: bool = True, path_dtype=torch.float32, impl: str = "auto",
                           max_neg_val: float = _lib.MAX_NEG_VAL, check: bool = False) -> AlignmentResult:
    """Fused log-prior + MAS: the whole block reference model/face_tts.py:165-174 in one call.
    mu_x [B,F,Tx], y [B,F,Ty] float32 CUDA; x_lengths/y_lengths [B] true lengths."""
    _need_cuda(mu_x, "mu_x")
    _need_cuda(y, "y")
    if mu_x.dim() != 3 or y.dim() != 3 or mu_x.shape[:2] != y.shape[:2]:
        raise ValueError("mu_x [B,F,Tx] and y [B,F,Ty] must agree on B and F")
    mu_x = mu_x.detach().to(torch.float32).contiguous()
    y = y.detach().to(torch.float32).contiguous()
    B, F, Tx = mu_x.shape
    Ty = y.shape[2]
    dev = mu_x.device
    L = _lib.lib()
    with torch.cuda.device(dev):
        t_x = _lengths(x_lengths, B, dev, "x_lengths")
        t_y = _lengths(y_lengths, B, dev, "y_lengths")
        path = torch.empty((B, Tx, Ty), dtype=path_dtype, device=dev) if dense_path else None
        dur = torch.empty((B, Tx), dtype=torch.int32, device=dev)
        ft = torch.empty((B, Ty), dtype=torch.int32, device=dev)
        status = torch.empty((B,), dtype=torch.int32, device=dev)
        ws_bytes = L.mas_b200_fused_workspace_bytes(B, F, Tx, Ty)
        ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev)
        rc = L.mas_b200_log_prior_maximum_path(
            mu_x.data_ptr(), y.data_ptr(), t_x.data_ptr(), t_y.data_ptr(), B, F, Tx, Ty, max_neg_val,
            path.data_ptr() if dense_path else None,
            _path_dtype_code(path_dtype) if dense_path else _lib.PATH_NONE,
            dur.data_ptr(), ft.data_ptr(), status.data_ptr(), ws.data_ptr(), ws_bytes, _IMPL[impl],
            _stream_ptr(dev))
        _lib.check(rc, "mas_b200_log_prior_maximum_path")
        if check:
            _raise_on_bad(status, "log_prior_maximum_path")
    return AlignmentResult(path, dur, ft, status)


class AlignmentPlan:
    """Reusable buffers for repeated fused calls of one shape (a training loop): outputs and workspace are
    allocated once, so a call is argument checks + ONE C-ABI call (~20 us of host time instead of ~45 us for
    `log_prior_maximum_path`, which allocates its outputs every time).  The returned AlignmentResult aliases the
    plan's buffers: it is overwritten by the next call on the plan (stream-ordered)."""

    def __init__(self, B: int, F: int, Tx: int, Ty: int, *, device=None, dense_path: bool = True,
                 path_dtype=torch.float32, impl: str = "auto", max_neg_val: float = _lib.MAX_NEG_VAL):
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.shape = (int(B), int(F), int(Tx), int(Ty))
        self.device, self.impl, self.neg = dev, _IMPL[impl], float(max_neg_val)
        L = _lib.lib()
        with torch.cuda.device(dev):
            self.path = torch.empty((B, Tx, Ty), dtype=path_dtype, device=dev) if dense_path else None
            self.path_code = _path_dtype_code(path_dtype) if dense_path else _lib.PATH_NONE
            self.durations = torch.empty((B, Tx), dtype=torch.int32, device=dev)
            self.frame_token = torch.empty((B, Ty), dtype=torch.int32, device=dev)
            self.status = torch.empty((B,), dtype=torch.int32, device=dev)
            self.ws_bytes = L.mas_b200_fused_workspace_bytes(B, F, Tx, Ty)
            self.ws = torch.empty((self.ws_bytes,), dtype=torch.uint8, device=dev)
        self._fn = L.mas_b200_log_prior_maximum_path
        self._result = AlignmentResult(self.path, self.durations, self.frame_token, self.status)

    def __call__(self, mu_x: torch.Tensor, y: torch.Tensor, x_lengths: torch.Tensor, y_lengths: torch.Tensor,
                 check: bool = False) -> AlignmentResult:
        """mu_x [B,F,Tx], y [B,F,Ty] float32 contiguous CUDA; x_lengths / y_lengths int32 [B] CUDA (no conversions
        are done here -- that is the point)."""
        B, F, Tx, Ty = self.shape
        if mu_x.shape != (B, F, Tx) or y.shape != (B, F, Ty) or mu_x.dtype != torch.float32 or y.dtype != torch.float32 \
                or not mu_x.is_contiguous() or not y.is_contiguous() or mu_x.device != self.device or y.device != self.device:
            raise ValueError("AlignmentPlan: mu_x / y must be contiguous float32 CUDA tensors of the planned shape on the plan's device")
        if x_lengths.dtype != torch.int32 or y_lengths.dtype != torch.int32 or x_lengths.device != self.device \
                or y_lengths.device != self.device or x_lengths.shape != (B,) or y_lengths.shape != (B,):
            raise ValueError("AlignmentPlan: lengths must be int32 [B] tensors on the plan's device")
        if torch.cuda.current_device() != self.device.index:       # the library works on the CURRENT device
            with torch.cuda.device(self.device):
                return self.__call__(mu_x, y, x_lengths, y_lengths, check)
        rc = self._fn(mu_x.data_ptr(), y.data_ptr(), x_lengths.data_ptr(), y_lengths.data_ptr(), B, F, Tx, Ty, self.neg,
                      self.path.data_ptr() if self.path is not None else None, self.path_code,
                      self.durations.data_ptr(), self.frame_token.data_ptr(), self.status.data_ptr(),
                      self.ws.data_ptr(), self.ws_bytes, self.impl, torch.cuda.current_stream(self.device).cuda_stream)
        _lib.check(rc, "mas_b200_log_prior_maximum_path")
        if check:
            _raise_on_bad(self.status, "AlignmentPlan")
        return self._result


def generate_path(duration: torch.Tensor, mask: torch.Tensor, *, return_index: bool = False):
    """Drop-in for reference model/utils.py:27-40 generate_path(duration, mask):
    duration [B,Tx] (float as at the call site face_tts.py:126, or integer), mask [B,Tx,Ty] prefix mask
    -> path [B,Tx,Ty] in mask.dtype.  `return_index=True` also returns frame_token [B,Ty] int32."""
    _need_cuda(duration, "duration")
    _need_cuda(mask, "mask")
    B, Tx, Ty = mask.shape
    dev = mask.device
    t_x, t_y = lengths_from_mask(mask)
    with torch.cuda.device(dev):
        path = torch.empty((B, Tx, Ty), dtype=torch.float32, device=dev)
        ft = torch.empty((B, Ty), dtype=torch.int32, device=dev) if return_index else None
        if duration.dtype.is_floating_point:
            d = duration.detach().to(torch.float32).contiguous()
            rc = _lib.lib().mas_b200_generate_path_f32(d.data_ptr(), t_x.data_ptr(), t_y.data_ptr(), B, Tx, Ty,
                                                       path.data_ptr(), _lib.PATH_F32,
                                                       ft.data_ptr() if ft is not None else None, _stream_ptr(dev))
        else:
            d = duration.detach().to(torch.int32).contiguous()
            rc = _lib.lib().mas_b200_generate_path(d.data_ptr(), t_x.data_ptr(), t_y.data_ptr(), B, Tx, Ty,
                                                   path.data_ptr(), _lib.PATH_F32, _stream_ptr(dev))
            if rc == 0 and ft is not None:
                rc = _lib.lib().mas_b200_generate_path_f32(d.to(torch.float32).data_ptr(), t_x.data_ptr(), t_y.data_ptr(),
                                                           B, Tx, Ty, None, _lib.PATH_NONE, ft.data_ptr(), _stream_ptr(dev))
        _lib.check(rc, "mas_b200_generate_path")
    path = path if mask.dtype == torch.float32 else path.to(mask.dtype)
    return (path, ft) if return_index else path


def expand_durations(mu_x: torch.Tensor, duration: torch.Tensor, x_lengths, y_lengths, Ty: int):
    """The inference-side expansion of reference model/face_tts.py:124-129 without the dense path:
    float durations [B,Tx] -> (mu_y [B,F,Ty], frame_token [B,Ty]); mu_y[:,:,t] = mu_x[:,:,token of frame t]."""
    _need_cuda(mu_x, "mu_x")
    B, F, Tx = mu_x.shape
    dev = mu_x.device
    with torch.cuda.device(dev):
        t_x = _lengths(x_lengths, B, dev, "x_lengths")
        t_y = _lengths(y_lengths, B, dev, "y_lengths")
        d = duration.detach().reshape(B, Tx).to(torch.float32).contiguous()
        mx = mu_x.detach().to(torch.float32).contiguous()
        ft = torch.empty((B, Ty), dtype=torch.int32, device=dev)
        mu_y = torch.empty((B, F, Ty), dtype=torch.float32, device=dev)
        L = _lib.lib()
        _lib.check(L.mas_b200_generate_path_f32(d.data_ptr(), t_x.data_ptr(), t_y.data_ptr(), B, Tx, Ty, None,
                                                _lib.PATH_NONE, ft.data_ptr(), _stream_ptr(dev)), "mas_b200_generate_path_f32")
        _lib.check(L.mas_b200_gather_mu_y(mx.data_ptr(), ft.data_ptr(), B, F, Tx, Ty, mu_y.data_ptr(), _stream_ptr(dev)),
                   "mas_b200_gather_mu_y")
    return mu_y, ft


_side_streams = {}


def _side_stream(dev):
    key = (dev.type, dev.index)
    if key not in _side_streams:
        _side_streams[key] = torch.cuda.Stream(dev)
    return _side_streams[key]


def upload_batch(mu_x: torch.Tensor, y: torch.Tensor, x_lengths: torch.Tensor, y_lengths: torch.Tensor, *,
                 device=None, out=None, mu_on_copy_engine: bool = False):
    """Host -> device transfer of one padded batch that moves only its valid part over PCIe
    (mas_b200_upload_batch): mu_x [B,F,Tx], y [B,F,Ty] float32 and int32 lengths [B], all in PINNED host
    memory.  Returns (mu_x, y, t_x, t_y) on the device (`out` = the same 4-tuple to reuse buffers), padding
    zero-filled.  Asynchronous on the current stream: do not touch the host tensors until it has completed.
    mu_on_copy_engine: move the small padded mu_x with a plain copy-engine transfer on a side stream while the
    zero-copy kernel pulls the ragged y (measured: the two requesters do not add up on the link -- 207 vs 218 us
    standalone, and the extra stream hops cost more than that in a pipelined loop -- so it is off by default)."""
    for t, name in ((mu_x, "mu_x"), (y, "y"), (x_lengths, "x_lengths"), (y_lengths, "y_lengths")):
        if t.is_cuda or not t.is_pinned() or not t.is_contiguous():
            raise ValueError(f"{name} must be a contiguous tensor in pinned host memory")
    if mu_x.dtype != torch.float32 or y.dtype != torch.float32 or x_lengths.dtype != torch.int32 or \
            y_lengths.dtype != torch.int32:
        raise ValueError("mu_x / y must be float32 and the lengths int32")
    B, F, Tx = mu_x.shape
    Ty = y.shape[2]
    if y.shape[:2] != (B, F) or x_lengths.shape != (B,) or y_lengths.shape != (B,):
        raise ValueError("mu_x [B,F,Tx], y [B,F,Ty] and lengths [B] must agree")
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    with torch.cuda.device(dev):
        if out is None:
            out = (torch.empty((B, F, Tx), dtype=torch.float32, device=dev),
                   torch.empty((B, F, Ty), dtype=torch.float32, device=dev),
                   torch.empty((B,), dtype=torch.int32, device=dev), torch.empty((B,), dtype=torch.int32, device=dev))
        if mu_on_copy_engine:
            cur, side = torch.cuda.current_stream(dev), _side_stream(dev)
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                out[0].copy_(mu_x, non_blocking=True)
        rc = _lib.lib().mas_b200_upload_batch(None if mu_on_copy_engine else mu_x.data_ptr(), y.data_ptr(),
                                              x_lengths.data_ptr(), y_lengths.data_ptr(),
                                              B, F, Tx, Ty, out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr(),
                                              out[3].data_ptr(), _stream_ptr(dev))
        _lib.check(rc, "mas_b200_upload_batch")
        if mu_on_copy_engine:
            cur.wait_stream(side)
    return out


def pack_batch(mu_x: torch.Tensor, y: torch.Tensor, x_lengths: torch.Tensor, y_lengths: torch.Tensor, *, out=None,
               pin: bool = True) -> torch.Tensor:
    """Host-side collate WITHOUT padding (mas_b200_pack_batch_host): padded HOST tensors mu_x [B,F,Tx], y [B,F,Ty]
    (float32) + int32 lengths -> one contiguous uint8 buffer
        [t_x: B int32][t_y: B int32][pad to 16 B][mu: for b, for f: t_x[b] floats][y: for b, for f: t_y[b] floats]
    holding only the valid data (what a dataset collate that never pads would produce directly).  Pinned by default
    so that `upload_packed_batch` moves it with one asynchronous copy."""
    for t, name in ((mu_x, "mu_x"), (y, "y"), (x_lengths, "x_lengths"), (y_lengths, "y_lengths")):
        if t.is_cuda or not t.is_contiguous():
            raise ValueError(f"{name} must be a contiguous host tensor")
    if mu_x.dtype != torch.float32 or y.dtype != torch.float32 or x_lengths.dtype != torch.int32 or \
            y_lengths.dtype != torch.int32:
        raise ValueError("mu_x / y must be float32 and the lengths int32")
    B, F, Tx = mu_x.shape
    Ty = y.shape[2]
    L = _lib.lib()
    nbytes = L.mas_b200_packed_batch_bytes(x_lengths.data_ptr(), y_lengths.data_ptr(), B, F)
    if out is None:
        out = torch.empty((nbytes,), dtype=torch.uint8)
        if pin:
            out = out.pin_memory()
    elif out.numel() < nbytes or out.dtype != torch.uint8 or out.is_cuda:
        raise ValueError("pack_batch: `out` must be a host uint8 tensor of at least packed_batch_bytes")
    _lib.check(L.mas_b200_pack_batch_host(mu_x.data_ptr(), y.data_ptr(), x_lengths.data_ptr(), y_lengths.data_ptr(),
                                          B, F, Tx, Ty, out.data_ptr(), out.numel()), "mas_b200_pack_batch_host")
    return out[:nbytes]


def upload_packed_batch(packed: torch.Tensor, B: int, F: int, Tx: int, Ty: int, *, device=None, out=None, staging=None):
    """Packed ragged batch (pack_batch layout, pinned host memory) -> zero-padded device tensors
    (mu_x [B,F,Tx], y [B,F,Ty], t_x [B], t_y [B]): ONE copy-engine transfer of the valid bytes + one device kernel
    (mas_b200_unpack_batch).  Asynchronous on the current stream.  `out` / `staging` (a CUDA uint8 buffer of at least
    packed.numel() bytes) let a training loop reuse its buffers."""
    if packed.is_cuda or packed.dtype != torch.uint8 or not packed.is_contiguous():
        raise ValueError("packed must be a contiguous uint8 host tensor (pack_batch)")
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    with torch.cuda.device(dev):
        if staging is None:
            staging = torch.empty((packed.numel(),), dtype=torch.uint8, device=dev)
        elif staging.numel() < packed.numel() or not staging.is_cuda:
            raise ValueError("staging must be a CUDA uint8 buffer of at least packed.numel() bytes")
        if out is None:
            out = (torch.empty((B, F, Tx), dtype=torch.float32, device=dev),
                   torch.empty((B, F, Ty), dtype=torch.float32, device=dev),
                   torch.empty((B,), dtype=torch.int32, device=dev), torch.empty((B,), dtype=torch.int32, device=dev))
        staging[:packed.numel()].copy_(packed, non_blocking=True)
        return unpack_batch(staging, B, F, Tx, Ty, out=out)


def unpack_batch(staging: torch.Tensor, B: int, F: int, Tx: int, Ty: int, *, out=None):
    """The device half of `upload_packed_batch` on its own (mas_b200_unpack_batch): a packed ragged batch already in
    device memory (`staging`, CUDA uint8, pack_batch layout) -> zero-padded (mu_x [B,F,Tx], y [B,F,Ty], t_x [B], t_y [B]).
    Asynchronous on the current stream.  A pipelined loop keeps the copy stream for the host->device copies alone (so
    that the copy engine runs back to back) and calls this on the compute stream in front of the step."""
    if not staging.is_cuda or staging.dtype != torch.uint8 or not staging.is_contiguous():
        raise ValueError("staging must be a contiguous CUDA uint8 tensor holding a packed batch")
    dev = staging.device
    with torch.cuda.device(dev):
        if out is None:
            out = (torch.empty((B, F, Tx), dtype=torch.float32, device=dev),
                   torch.empty((B, F, Ty), dtype=torch.float32, device=dev),
                   torch.empty((B,), dtype=torch.int32, device=dev), torch.empty((B,), dtype=torch.int32, device=dev))
        _lib.check(_lib.lib().mas_b200_unpack_batch(staging.data_ptr(), B, F, Tx, Ty, out[0].data_ptr(), out[1].data_ptr(),
                                                    out[2].data_ptr(), out[3].data_ptr(), _stream_ptr(dev)),
                   "mas_b200_unpack_batch")
    return out


def durations_to_logw(durations: torch.Tensor, x_mask: torch.Tensor) -> torch.Tensor:
    """logw_ = log(1e-8 + sum_t attn) * x_mask  (reference model/face_tts.py:176) straight from the
    integer durations the backtrack emits -- no dense re-read.  x_mask [B,1,Tx] -> [B,1,Tx]."""
    return torch.log(1e-8 + durations.to(x_mask.dtype).unsqueeze(1)) * x_mask


LOG_PRIOR_CONST_PER_FEAT = -0.5 * math.log(2 * math.pi)
