"""The consumers of the alignment inside FaceTTS.compute_loss, on the index form of the path.

Reference model/face_tts.py:161-234 materialises a dense [B,Tx,Ty] `attn`, re-reads it for the
durations (:176), slices it per utterance in a Python loop (:204-209), multiplies it with mu_x as a
K=Tx GEMM (:217-218) and reduces the prior loss from the result (:233-234).  Here the same numbers
come from durations [B,Tx] / frame_token [B,Ty] (what the backtrack already knows) through the small
CUDA kernels of csrc/loss_ops.cu; gradients (w.r.t. `logw` and `mu_x` only -- `attn` is detached in
the reference, :174) are custom autograd Functions over the matching backward kernels.

`alignment_losses` is the drop-in for the whole block :161-218 + :233-234.
"""
from __future__ import annotations

import random
from dataclasses import dataclass
from typing import Optional

import torch

from . import _lib
from .alignment import AlignmentResult, _lengths, _need_cuda, _stream_ptr, log_prior_maximum_path


def _f32c(t):
    return t.to(torch.float32).contiguous()


def token_starts(durations: torch.Tensor) -> torch.Tensor:
    """Exclusive prefix sum of the durations: first frame of every token, int32 [B,Tx]."""
    return (torch.cumsum(durations, 1, dtype=torch.int32) - durations).contiguous()


def sequence_mask(lengths: torch.Tensor, max_length: Optional[int] = None, dtype=torch.float32) -> torch.Tensor:
    """Drop-in for reference model/utils.py:6-11 (plus the `.to(x_mask)` cast of face_tts.py:161):
    [B,T] mask of `t < lengths[b]` built on the device.  `max_length=None` syncs, like the reference."""
    _need_cuda(lengths, "lengths")
    if max_length is None:
        max_length = int(lengths.max())
    B, T, dev = lengths.shape[0], int(max_length), lengths.device
    with torch.cuda.device(dev):
        ln = lengths.to(torch.int32).contiguous()
        out = torch.empty((B, T), dtype=torch.float32, device=dev)
        if T > 0:
            _lib.check(_lib.lib().mas_b200_sequence_mask(ln.data_ptr(), B, T, out.data_ptr(), _stream_ptr(dev)),
                       "mas_b200_sequence_mask")
    return out if dtype == torch.float32 else out.to(dtype)


def crop_frames(y: torch.Tensor, frame_token: Optional[torch.Tensor], y_lengths, offsets, out_size: int):
    """One launch for the per-utterance crop loop of reference face_tts.py:204-211.
    Returns (y_cut [B,F,out_size], frame_token_cut [B,out_size] or None, cut_lengths [B] int32,
    y_cut_mask [B,1,out_size])."""
    _need_cuda(y, "y")
    B, F, Ty = y.shape
    dev = y.device
    with torch.cuda.device(dev):
        yc = _f32c(y.detach())
        ln = _lengths(y_lengths, B, dev, "y_lengths")
        off = _lengths(offsets, B, dev, "offsets")
        y_cut = torch.empty((B, F, out_size), dtype=torch.float32, device=dev)
        ft_cut = torch.empty((B, out_size), dtype=torch.int32, device=dev) if frame_token is not None else None
        cut_len = torch.empty((B,), dtype=torch.int32, device=dev)
        mask = torch.empty((B, 1, out_size), dtype=torch.float32, device=dev)
        ft = frame_token.contiguous() if frame_token is not None else None
        rc = _lib.lib().mas_b200_crop_frames(
            yc.data_ptr(), ft.data_ptr() if ft is not None else None, ln.data_ptr(), off.data_ptr(), B, F, Ty,
            int(out_size), y_cut.data_ptr(), ft_cut.data_ptr() if ft_cut is not None else None, cut_len.data_ptr(),
            mask.data_ptr(), _stream_ptr(dev))
        _lib.check(rc, "mas_b200_crop_frames")
    return y_cut, ft_cut, cut_len, mask


class _GatherMuY(torch.autograd.Function):
    """mu_y[b,f,t] = mu_x[b,f,frame_token[b,t]]  == attn^T @ mu_x^T (reference face_tts.py:217-218)."""

    @staticmethod
    def forward(ctx, mu_x, frame_token, start, durations, offsets, lengths):
        B, F, Tx = mu_x.shape
        T = frame_token.shape[1]
        dev = mu_x.device
        with torch.cuda.device(dev):
            mx = _f32c(mu_x)
            mu_y = torch.empty((B, F, T), dtype=torch.float32, device=dev)
            rc = _lib.lib().mas_b200_gather_mu_y(mx.data_ptr(), frame_token.data_ptr(), B, F, Tx, T, mu_y.data_ptr(),
                                                 _stream_ptr(dev))
            _lib.check(rc, "mas_b200_gather_mu_y")
        ctx.save_for_backward(start, durations, offsets, lengths)
        ctx.dims = (B, F, Tx, T)
        ctx.in_dtype = mu_x.dtype
        return mu_y

    @staticmethod
    def backward(ctx, grad):
        start, durations, offsets, lengths = ctx.saved_tensors
        B, F, Tx, T = ctx.dims
        dev = grad.device
        with torch.cuda.device(dev):
            g = _f32c(grad)
            gx = torch.empty((B, F, Tx), dtype=torch.float32, device=dev)
            rc = _lib.lib().mas_b200_gather_mu_y_backward(
                g.data_ptr(), start.data_ptr(), durations.data_ptr(),
                offsets.data_ptr() if offsets is not None else None,
                lengths.data_ptr() if lengths is not None else None, B, F, Tx, T, gx.data_ptr(), _stream_ptr(dev))
            _lib.check(rc, "mas_b200_gather_mu_y_backward")
        return gx.to(ctx.in_dtype), None, None, None, None, None


def gather_mu_y(mu_x: torch.Tensor, frame_token: torch.Tensor, durations: torch.Tensor, *,
                start: Optional[torch.Tensor] = None, offsets: Optional[torch.Tensor] = None,
                lengths: Optional[torch.Tensor] = None) -> torch.Tensor:
    """mu_y [B,F,T] from mu_x [B,F,Tx] and the frame->token index [B,T] (differentiable w.r.t. mu_x).
    `offsets`/`lengths` describe the crop window `frame_token` was cut with (None: the whole utterance)."""
    _need_cuda(mu_x, "mu_x")
    if start is None:
        start = token_starts(durations)
    return _GatherMuY.apply(mu_x, frame_token.contiguous(), start, durations.contiguous(), offsets, lengths)


class _PriorLoss(torch.autograd.Function):
    """(prior_loss, mu_y) in one pass over y (reference face_tts.py:217-218, 233-234)."""

    @staticmethod
    def forward(ctx, mu_x, y, frame_token, y_lengths, start, durations, offsets):
        B, F, Tx = mu_x.shape
        T = y.shape[2]
        dev = mu_x.device
        L = _lib.lib()
        with torch.cuda.device(dev):
            mx, yc = _f32c(mu_x), _f32c(y)
            mu_y = torch.empty((B, F, T), dtype=torch.float32, device=dev)
            loss = torch.empty((1,), dtype=torch.float32, device=dev)
            ws_bytes = L.mas_b200_prior_loss_workspace_bytes(B, F, T)
            ws = torch.empty((ws_bytes // 8,), dtype=torch.float64, device=dev)
            rc = L.mas_b200_prior_loss(yc.data_ptr(), mx.data_ptr(), frame_token.data_ptr(), y_lengths.data_ptr(), B, F,
                                       Tx, T, mu_y.data_ptr(), loss.data_ptr(), ws.data_ptr(), ws_bytes, _stream_ptr(dev))
            _lib.check(rc, "mas_b200_prior_loss")
        ctx.save_for_backward(mx, yc, y_lengths, start, durations, offsets)
        ctx.dims = (B, F, Tx, T)
        ctx.in_dtype = mu_x.dtype
        ctx.mark_non_differentiable(mu_y)
        return loss.reshape(()), mu_y

    @staticmethod
    def backward(ctx, grad_loss, _grad_mu_y):
        mx, yc, y_lengths, start, durations, offsets = ctx.saved_tensors
        B, F, Tx, T = ctx.dims
        dev = mx.device
        with torch.cuda.device(dev):
            g = grad_loss.to(torch.float32).reshape(1).contiguous()
            gx = torch.empty((B, F, Tx), dtype=torch.float32, device=dev)
            rc = _lib.lib().mas_b200_prior_loss_backward(
                yc.data_ptr(), mx.data_ptr(), start.data_ptr(), durations.data_ptr(),
                offsets.data_ptr() if offsets is not None else None, y_lengths.data_ptr(), g.data_ptr(), B, F, Tx, T,
                gx.data_ptr(), _stream_ptr(dev))
            _lib.check(rc, "mas_b200_prior_loss_backward")
        return gx.to(ctx.in_dtype), None, None, None, None, None, None


def prior_loss(mu_x: torch.Tensor, y: torch.Tensor, frame_token: torch.Tensor, y_lengths, durations: torch.Tensor, *,
               start: Optional[torch.Tensor] = None, offsets: Optional[torch.Tensor] = None) -> torch.Tensor:
    """sum(0.5*((y-mu_y)^2 + log 2pi) * y_mask) / (sum(y_mask)*F)  (reference face_tts.py:233-234) without
    materialising mu_y for the loss; differentiable w.r.t. mu_x."""
    _need_cuda(mu_x, "mu_x")
    B = mu_x.shape[0]
    if start is None:
        start = token_starts(durations)
    ln = _lengths(y_lengths, B, mu_x.device, "y_lengths")
    loss, _ = _PriorLoss.apply(mu_x, y.detach(), frame_token.contiguous(), ln, start, durations.contiguous(), offsets)
    return loss


class _DurationLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logw, durations, x_lengths):
        B, Tx = durations.shape
        dev = logw.device
        with torch.cuda.device(dev):
            lw = _f32c(logw).reshape(B, Tx)
            loss = torch.empty((1,), dtype=torch.float32, device=dev)
            grad = torch.empty((B, Tx), dtype=torch.float32, device=dev)
            rc = _lib.lib().mas_b200_duration_loss(lw.data_ptr(), durations.data_ptr(), x_lengths.data_ptr(), B, Tx,
                                                   loss.data_ptr(), None, grad.data_ptr(), _stream_ptr(dev))
            _lib.check(rc, "mas_b200_duration_loss")
        ctx.save_for_backward(grad)
        ctx.shape = logw.shape
        ctx.in_dtype = logw.dtype
        return loss.reshape(())

    @staticmethod
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        return (grad * g).reshape(ctx.shape).to(ctx.in_dtype), None, None


def duration_loss(logw: torch.Tensor, durations: torch.Tensor, x_lengths) -> torch.Tensor:
    """duration_loss(logw, log(1e-8 + durations) * x_mask, x_lengths) of reference model/utils.py:43-45 /
    face_tts.py:176-179, from the integer durations.  logw [B,1,Tx] or [B,Tx]; differentiable w.r.t. logw."""
    _need_cuda(logw, "logw")
    B, Tx = durations.shape
    if logw.numel() != B * Tx:
        raise ValueError("logw must hold B*Tx elements")
    ln = _lengths(x_lengths, B, logw.device, "x_lengths")
    return _DurationLoss.apply(logw, durations.contiguous(), ln)


def logw_target(durations: torch.Tensor, x_lengths) -> torch.Tensor:
    """logw_ [B,1,Tx] = log(1e-8 + durations) * x_mask (reference face_tts.py:176)."""
    _need_cuda(durations, "durations")
    B, Tx = durations.shape
    dev = durations.device
    ln = _lengths(x_lengths, B, dev, "x_lengths")
    with torch.cuda.device(dev):
        zero = torch.zeros((B, Tx), dtype=torch.float32, device=dev)
        loss = torch.empty((1,), dtype=torch.float32, device=dev)
        tgt = torch.empty((B, Tx), dtype=torch.float32, device=dev)
        rc = _lib.lib().mas_b200_duration_loss(zero.data_ptr(), durations.contiguous().data_ptr(), ln.data_ptr(), B, Tx,
                                               loss.data_ptr(), tgt.data_ptr(), None, _stream_ptr(dev))
        _lib.check(rc, "mas_b200_duration_loss")
    return tgt.unsqueeze(1)


def draw_crop_offsets(y_lengths_host, out_size: int, rng=random):
    """The reference's window draw (face_tts.py:182-194): per utterance `random.choice(range(0, max(len-out_size,0)))`
    or 0 -- same consumption of Python's `random` stream, so a seeded run crops the same windows."""
    out = []
    for n in y_lengths_host:
        end = max(int(n) - int(out_size), 0)
        out.append(rng.choice(range(0, end)) if end > 0 else 0)
    return out


@dataclass
class AlignmentLosses:
    """What reference FaceTTS.compute_loss holds after face_tts.py:218 (+ the prior loss of :233-234).

    dur_loss, prior_loss   scalars (differentiable w.r.t. logw / mu_x)
    mu_y    [B,F,T]        aligned encoder means for the decoder (differentiable w.r.t. mu_x); T = out_size if cropped
    y       [B,F,T]        (cropped) target
    y_mask  [B,1,T]        (cropped) frame mask
    y_lengths [B] int32    (cropped) lengths
    offsets [B] int32 or None   crop offsets used
    alignment              AlignmentResult (durations, frame_token, status; dense path only if requested)
    """
    dur_loss: torch.Tensor
    prior_loss: torch.Tensor
    mu_y: torch.Tensor
    y: torch.Tensor
    y_mask: torch.Tensor
    y_lengths: torch.Tensor
    offsets: Optional[torch.Tensor]
    alignment: AlignmentResult


def alignment_losses(mu_x: torch.Tensor, logw: torch.Tensor, x_lengths, y: torch.Tensor, y_lengths, *,
                     out_size: Optional[int] = None, out_offset=None, dense_path: bool = False,
                     impl: str = "auto", rng=random, alignment: Optional[AlignmentResult] = None) -> AlignmentLosses:
    """Drop-in for the alignment block of FaceTTS.compute_loss, reference model/face_tts.py:161-218 + 233-234:
    masks -> log-prior -> MAS -> durations/duration loss -> random crop -> mu_y -> prior loss, with lengths
    instead of dense masks and the index form of the path instead of `attn`.

    mu_x [B,F,Tx], logw [B,1,Tx] (encoder outputs, may require grad), y [B,F,Ty]; x_lengths / y_lengths [B].
    out_size: crop window in frames (None: no crop).  out_offset: explicit [B] offsets; None draws them like the
    reference (needs y_lengths on the host -- pass the CPU tensor the data loader produced to avoid a sync).
    alignment: a precomputed AlignmentResult (e.g. from maximum_path_from_lengths on a caller-supplied value
    matrix); None runs the fused log-prior + MAS here."""
    _need_cuda(mu_x, "mu_x")
    _need_cuda(y, "y")
    B, F, Tx = mu_x.shape
    dev = mu_x.device
    res = alignment
    if res is None:
        with torch.no_grad():                                          # face_tts.py:165, attn detached :174
            res = log_prior_maximum_path(mu_x, y, x_lengths, y_lengths, dense_path=dense_path, impl=impl)
    dur = res.durations
    start = token_starts(dur)
    x_len = _lengths(x_lengths, B, dev, "x_lengths")
    y_len = _lengths(y_lengths, B, dev, "y_lengths")
    d_loss = _DurationLoss.apply(logw, dur, x_len)                     # :176-179

    offsets = None
    ft = res.frame_token
    y_used = y.detach()
    if out_size is not None:                                           # :181-215
        if out_offset is None:
            host_len = y_lengths if (torch.is_tensor(y_lengths) and not y_lengths.is_cuda) else \
                (y_lengths.cpu() if torch.is_tensor(y_lengths) else y_lengths)
            out_offset = draw_crop_offsets([int(v) for v in host_len], out_size, rng)
        offsets = _lengths(out_offset, B, dev, "out_offset")
        y_used, ft, y_len, y_mask = crop_frames(y_used, ft, y_len, offsets, out_size)
    else:
        y_mask = sequence_mask(y_len, y.shape[2]).unsqueeze(1)         # :161
    p_loss, mu_y_nograd = _PriorLoss.apply(mu_x, y_used, ft, y_len, start, dur, offsets)      # :217-218, :233-234
    # the decoder's mu_y carries its own gradient path (diffusion loss); same values as the fused pass produced
    mu_y = _GatherMuY.apply(mu_x, ft, start, dur, offsets, y_len if offsets is not None else None) \
        if mu_x.requires_grad else mu_y_nograd
    return AlignmentLosses(d_loss, p_loss, mu_y, y_used, y_mask, y_len, offsets, res)
