"""Synthetic inputs of the shapes BASELINE.json names.

Everything is generated on the CPU from a seeded torch.Generator so that the
bench, the tests and the golden-vector script see bit-identical inputs.
Lengths follow the reference's data conventions: text length odd (tokens
interspersed with blanks, reference utils/tts_util.py:17-21), mel length
padded to a multiple of 4 (reference data/lrs2_dataset.py:251-252), element
0 of every batch at full length.
"""
from __future__ import annotations

import torch

LOG_MEL_FLOOR = -11.512925  # log(1e-5): reference utils/mel_spectrogram.py:26-27


def _gen(seed: int) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed(int(seed))
    return g


def _randint(g, lo, hi, n):
    """n ints uniform in [lo, hi] inclusive."""
    return torch.randint(int(lo), int(hi) + 1, (n,), generator=g, dtype=torch.int64)


def lengths_mas(B, Tx, Ty, tx_lo, ty_lo, seed):
    """t_x ~ U{tx_lo..Tx}, t_y ~ U{max(t_x, ty_lo)..Ty}; element 0 full length."""
    g = _gen(seed)
    t_x = _randint(g, tx_lo, Tx, B)
    t_y = torch.empty(B, dtype=torch.int64)
    for b in range(B):
        t_y[b] = _randint(g, max(int(t_x[b]), ty_lo), Ty, 1)[0]
    t_x[0], t_y[0] = Tx, Ty
    return t_x.to(torch.int32), t_y.to(torch.int32)


def mas_value(B=16, Tx=200, Ty=800, seed=1234, tx_lo=None, ty_lo=None):
    """configs[0]: fp32 value ~ N(0,1) with ragged lengths (MAS only).
    Returns value [B,Tx,Ty], t_x [B] i32, t_y [B] i32."""
    g = _gen(seed)
    value = torch.randn(B, Tx, Ty, generator=g, dtype=torch.float32)
    t_x, t_y = lengths_mas(B, Tx, Ty, tx_lo if tx_lo is not None else Tx // 2,
                           ty_lo if ty_lo is not None else Ty // 2, seed + 1)
    return value, t_x, t_y


def lrs2_batch(B=32, F=80, Tx=190, Ty=1000, seed=1234, tx_lo=60, ty_lo=300):
    """configs[1]: LRS2-shaped encoder means and log-mels.
    mu_x [B,F,Tx] ~ N(0,1), zero beyond t_x (reference text_encoder.py:417);
    y [B,F,Ty] ~ clamp(N(-5,2), log(1e-5), 2), zero-padded beyond t_y
    (reference lrs2_dataset.py:256,265); t_x odd."""
    g = _gen(seed)
    mu_x = torch.randn(B, F, Tx, generator=g, dtype=torch.float32)
    y = (torch.randn(B, F, Ty, generator=g, dtype=torch.float32) * 2.0 - 5.0).clamp_(LOG_MEL_FLOOR, 2.0)
    t_x = _randint(g, tx_lo, Tx, B)
    t_x = t_x - (1 - t_x % 2)           # make odd (2n+1 tokens with blanks)
    t_x.clamp_(min=1)
    t_y = torch.empty(B, dtype=torch.int64)
    for b in range(B):
        t_y[b] = _randint(g, max(int(t_x[b]), ty_lo), Ty, 1)[0]
    t_x[0] = Tx                         # element 0 at full padded length
    t_y[0] = Ty
    ar_x = torch.arange(Tx)[None, None, :]
    ar_y = torch.arange(Ty)[None, None, :]
    mu_x = mu_x * (ar_x < t_x[:, None, None])
    y = y * (ar_y < t_y[:, None, None])
    return mu_x.contiguous(), y.contiguous(), t_x.to(torch.int32), t_y.to(torch.int32)


def prefix_mask(t_x, t_y, Tx, Ty, dtype=torch.float32):
    """attn_mask as the call site builds it (reference face_tts.py:161-162), squeezed to [B,Tx,Ty]."""
    xm = (torch.arange(Tx)[None, :] < t_x[:, None].long()).to(dtype)
    ym = (torch.arange(Ty)[None, :] < t_y[:, None].long()).to(dtype)
    return xm[:, :, None] * ym[:, None, :]
