"""CPU oracle for the log-prior + Monotonic Alignment Search hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package
(face-gan-tts_b200/) imports this; only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / `--impl reference` legs do, and only as the checker
or as the timed CPU baseline -- never as the thing shipped.

Parity status: PINNED against the reference's own compiled core.pyx
(oracle/build_ref.py -> oracle/_ref/) and the fixtures that build generated
(tests/golden/).  The reference has no golden vectors of its own
(SURVEY.md section 4).

Three layers, each citing the reference lines it restates:

  maximum_path_c        plain-C restatement (oracle/mas_oracle.c) of
                        model/monotonic_align/core.pyx:9-45, same in-place numpy
                        signature as the Cython `maximum_path_c` (core.pyx:40)
  maximum_path_numpy    independent numpy restatement (rolling column + one
                        direction bit per cell -- the formulation the CUDA
                        kernels use), core.pyx:9-35
  maximum_path          the Python wrapper model/monotonic_align/__init__.py:8-23
  log_prior_reference   model/face_tts.py:165-171 verbatim in torch
  reference_core        the reference's real compiled module, when built
"""
from __future__ import annotations

import ctypes
import math
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libmas_oracle.so")
_lib = None

MAX_NEG_VAL = -1e9  # core.pyx:40 default


def build_c(force: bool = False) -> str:
    """gcc the plain-C restatement into oracle/libmas_oracle.so."""
    src = os.path.join(_HERE, "mas_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-Wall", "-shared", "-o", _LIB_PATH, src])
    return _LIB_PATH


def _load():
    global _lib
    if _lib is None:
        build_c()
        lib = ctypes.CDLL(_LIB_PATH)
        lib.mas_oracle_batch.restype = ctypes.c_int
        lib.mas_oracle_batch.argtypes = [
            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
            ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_float,
        ]
        lib.mas_oracle_durations.restype = None
        lib.mas_oracle_durations.argtypes = [
            ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
        ]
        _lib = lib
    return _lib


def _check(a, dtype, ndim, name):
    if not (isinstance(a, np.ndarray) and a.dtype == dtype and a.ndim == ndim and a.flags.c_contiguous):
        raise TypeError(f"{name}: need C-contiguous {np.dtype(dtype).name} ndarray with ndim={ndim}")


def maximum_path_c(paths, values, t_xs, t_ys, max_neg_val: float = MAX_NEG_VAL) -> int:
    """Same contract as the Cython `maximum_path_c` (core.pyx:40): `paths`
    int32 [B,Tx,Ty] pre-zeroed, `values` float32 [B,Tx,Ty] CLOBBERED in place,
    `t_xs`,`t_ys` int32 [B].  Returns the number of rejected items
    (t_x > t_y or < 1; the reference is undefined there)."""
    _check(paths, np.int32, 3, "paths")
    _check(values, np.float32, 3, "values")
    _check(t_xs, np.int32, 1, "t_xs")
    _check(t_ys, np.int32, 1, "t_ys")
    B, Tx, Ty = values.shape
    assert paths.shape == values.shape and t_xs.shape == (B,) and t_ys.shape == (B,)
    return _load().mas_oracle_batch(
        paths.ctypes.data, values.ctypes.data, t_xs.ctypes.data, t_ys.ctypes.data,
        B, Tx, Ty, ctypes.c_float(max_neg_val),
    )


def durations_and_frame_token(path_i32):
    """[B,Tx,Ty] int32 path -> (durations [B,Tx] i32, frame_token [B,Ty] i32, -1 where no token)."""
    _check(path_i32, np.int32, 3, "path")
    B, Tx, Ty = path_i32.shape
    dur = np.zeros((B, Tx), np.int32)
    ft = np.zeros((B, Ty), np.int32)
    lib = _load()
    for b in range(B):
        lib.mas_oracle_durations(path_i32[b].ctypes.data, Tx, Ty, dur[b].ctypes.data, ft[b].ctypes.data)
    return dur, ft


def lengths_from_mask(mask_np):
    """model/monotonic_align/__init__.py:20-21: t_x from mask column 0, t_y from row 0."""
    t_x = mask_np.sum(1)[:, 0].astype(np.int32)
    t_y = mask_np.sum(2)[:, 0].astype(np.int32)
    return t_x, t_y


def maximum_path(value, mask, core=None):
    """Restatement of the reference wrapper, model/monotonic_align/__init__.py:8-23.
    `value`, `mask`: torch tensors [B,Tx,Ty] on any device; returns a tensor of
    value.dtype on value.device.  `core` = module/object providing
    maximum_path_c (default: the C restatement in this package)."""
    import torch

    value = value * mask                                            # __init__.py:13
    device, dtype = value.device, value.dtype
    value_np = value.data.cpu().numpy().astype(np.float32)          # :16
    path = np.zeros_like(value_np).astype(np.int32)                 # :17
    mask_np = mask.data.cpu().numpy()                               # :18
    t_x_max, t_y_max = lengths_from_mask(mask_np)                   # :20-21
    fn = maximum_path_c if core is None else core.maximum_path_c
    fn(path, np.ascontiguousarray(value_np), t_x_max, t_y_max)      # :22
    return torch.from_numpy(path).to(device=device, dtype=dtype)    # :23


def maximum_path_numpy(values, t_xs, t_ys, max_neg_val: float = MAX_NEG_VAL):
    """Independent numpy restatement of core.pyx:9-35 in the form the CUDA
    kernels use: a rolling previous column Q[:,y-1], ONE direction bit per cell
    (bit = v_prev > v_cur, the same predicate the backtrack re-derives at
    core.pyx:34), no lower band bound (the in-band recursion is closed).
    `values` is NOT clobbered.  Returns int32 paths [B,Tx,Ty]."""
    values = np.asarray(values, np.float32)
    B, Tx, Ty = values.shape
    paths = np.zeros((B, Tx, Ty), np.int32)
    neg = np.float32(max_neg_val)
    for b in range(B):
        t_x, t_y = int(t_xs[b]), int(t_ys[b])
        if t_x < 1 or t_y < 1 or t_x > t_y:
            continue
        v = values[b, :t_x, :t_y]
        bits = np.zeros((t_x, t_y), bool)
        q = np.full(t_x, neg, np.float32)          # column y-1 (unused at y=0 except via masks)
        xs = np.arange(t_x)
        for y in range(t_y):
            v_cur = np.where(xs >= y, neg, q)                      # x==y -> max_neg_val (x>y: never read)
            v_prev = np.empty(t_x, np.float32)
            v_prev[1:] = q[:-1]
            v_prev[0] = np.float32(0.0) if y == 0 else neg         # core.pyx:23-27
            with np.errstate(invalid="ignore"):
                d = v_prev > v_cur                                  # NaN -> False -> v_cur
            q = (np.where(d, v_prev, v_cur) + v[:, y]).astype(np.float32)
            bits[:, y] = d
        index = t_x - 1
        for y in range(t_y - 1, -1, -1):
            paths[b, index, y] = 1
            if index != 0 and (index == y or bits[index, y]):
                index -= 1
    return paths


def log_prior_reference(mu_x, y):
    """model/face_tts.py:165-171 verbatim (torch, whatever device the inputs are on).
    mu_x [B,F,Tx], y [B,F,Ty] -> log_prior [B,Tx,Ty]."""
    import torch

    n_feats = mu_x.shape[1]
    with torch.no_grad():
        const = -0.5 * math.log(2 * math.pi) * n_feats                                     # :166
        factor = -0.5 * torch.ones(mu_x.shape, dtype=mu_x.dtype, device=mu_x.device)      # :167
        y_square = torch.matmul(factor.transpose(1, 2), y ** 2)                            # :168
        y_mu_double = torch.matmul(2.0 * (factor * mu_x).transpose(1, 2), y)               # :169
        mu_square = torch.sum(factor * (mu_x ** 2), 1).unsqueeze(-1)                       # :170
        log_prior = y_square - y_mu_double + mu_square + const                             # :171
    return log_prior


def log_prior_direct(mu_x, y):
    """Same quantity in its un-expanded form, -0.5*sum_f (y-mu)^2 - 0.5*F*log(2*pi), in
    float64 -- the ground truth both the reference expansion and the CUDA kernel are
    compared with (tolerance 1e-4 relative, BASELINE.json north_star)."""
    import torch

    n_feats = mu_x.shape[1]
    mu = mu_x.double()
    yy = y.double()
    d = yy.unsqueeze(2) - mu.unsqueeze(3)            # [B,F,Tx,Ty]
    return -0.5 * (d * d).sum(1) - 0.5 * math.log(2 * math.pi) * n_feats


def reference_core(variant: str = "asis"):
    """The reference's own compiled core (oracle/_ref/<variant>/core.*.so) or None."""
    from . import build_ref

    return build_ref.load(variant)


# ---------------------------------------------------------------------------------------------
# The alignment block of FaceTTS.compute_loss (model/face_tts.py:161-218, 233-234) restated in torch.
# Pinned against the reference's real FaceTTS.compute_loss by tests/golden/make_compute_loss_golden.py
# (fixture tests/golden/compute_loss_block.npz).
# ---------------------------------------------------------------------------------------------
def sequence_mask(length, max_length=None):
    """model/utils.py:6-11."""
    import torch

    if max_length is None:
        max_length = length.max()
    x = torch.arange(int(max_length), dtype=length.dtype, device=length.device)
    return x.unsqueeze(0) < length.unsqueeze(1)


def duration_loss(logw, logw_, lengths):
    """model/utils.py:43-45."""
    import torch

    return torch.sum((logw - logw_) ** 2) / torch.sum(lengths)


def generate_path(duration, mask):
    """model/utils.py:27-40."""
    import torch

    b, t_x, t_y = mask.shape
    cum_duration = torch.cumsum(duration, 1)
    path = sequence_mask(cum_duration.view(b * t_x), t_y).to(mask.dtype).view(b, t_x, t_y)
    path = path - torch.nn.functional.pad(path, [0, 0, 1, 0, 0, 0])[:, :-1]
    return path * mask


def compute_loss_block(mu_x, logw, x_mask, y, y_lengths, x_lengths, n_feats, out_size=None, out_offset=None,
                       maximum_path_fn=None):
    """model/face_tts.py:159-218 + 233-234, dense formulation, autograd-differentiable w.r.t. mu_x and logw.
    mu_x [B,F,Tx], logw [B,1,Tx], x_mask [B,1,Tx] float, y [B,F,Ty]; lengths int64 [B].
    `out_offset` (int64 [B]) replaces the reference's `random.choice` draw (:186-194) when given.
    Returns dict(dur_loss, prior_loss, mu_y, y, y_mask, attn, logw_)."""
    import random

    import torch

    if maximum_path_fn is None:
        maximum_path_fn = maximum_path
    y_max_length = y.shape[-1]                                                            # :159
    y_mask = sequence_mask(y_lengths, y_max_length).unsqueeze(1).to(x_mask)               # :161
    attn_mask = x_mask.unsqueeze(-1) * y_mask.unsqueeze(2)                                # :162
    with torch.no_grad():
        log_prior = log_prior_reference(mu_x, y)                                          # :166-171
        attn = maximum_path_fn(log_prior, attn_mask.squeeze(1))                           # :173
        attn = attn.detach()                                                              # :174
    logw_ = torch.log(1e-8 + torch.sum(attn.unsqueeze(1), -1)) * x_mask                   # :176
    dur_loss = duration_loss(logw, logw_, x_lengths)                                      # :179
    attn_full = attn
    if out_size is not None:                                                              # :181
        max_offset = (y_lengths - out_size).clamp(0)                                      # :182
        if out_offset is None:
            offset_ranges = list(zip([0] * max_offset.shape[0], max_offset.cpu().numpy()))    # :183-185
            out_offset = torch.LongTensor(
                [torch.tensor(random.choice(range(start, end)) if end > start else 0)
                 for start, end in offset_ranges]).to(y_lengths)                          # :186-191
        attn_cut = torch.zeros(attn.shape[0], attn.shape[1], out_size, dtype=attn.dtype, device=attn.device)
        y_cut = torch.zeros(y.shape[0], n_feats, out_size, dtype=y.dtype, device=y.device)
        y_cut_lengths = []
        for i, (y_, out_offset_) in enumerate(zip(y, out_offset)):                        # :204
            y_cut_length = out_size + (y_lengths[i] - out_size).clamp(None, 0)            # :205
            y_cut_lengths.append(y_cut_length)
            cut_lower, cut_upper = out_offset_, out_offset_ + y_cut_length                # :207
            y_cut[i, :, :y_cut_length] = y_[:, cut_lower:cut_upper]                       # :208
            attn_cut[i, :, :y_cut_length] = attn[i, :, cut_lower:cut_upper]               # :209
        y_cut_lengths = torch.LongTensor(y_cut_lengths)
        y_cut_mask = sequence_mask(y_cut_lengths).unsqueeze(1).to(y_mask)                 # :211
        attn, y, y_mask = attn_cut, y_cut, y_cut_mask.to(y.device)                        # :213-215
    mu_y = torch.matmul(attn.squeeze(1).transpose(1, 2), mu_x.transpose(1, 2))            # :217
    mu_y = mu_y.transpose(1, 2)                                                           # :218
    prior_loss = torch.sum(0.5 * ((y - mu_y) ** 2 + math.log(2 * math.pi)) * y_mask)      # :233
    prior_loss = prior_loss / (torch.sum(y_mask) * n_feats)                               # :234
    return dict(dur_loss=dur_loss, prior_loss=prior_loss, mu_y=mu_y, y=y, y_mask=y_mask, attn=attn_full,
                attn_cut=attn, logw_=logw_, out_offset=out_offset)
