"""Known-answer cases for maximum_path (SURVEY.md section 8(c) list).

Each case is a function of nothing (fixed seeds) returning
  value [B,Tx,Ty] float32 ndarray, t_x [B] int32, t_y [B] int32.
Small cases are rebuilt from numpy arithmetic that cannot drift (integers,
constants); random cases use seeded torch CPU generators and the fixture
stores a sha256 of the input bytes so generator drift is detected, not
silently accepted.

The expected outputs live in tests/golden/mas_kats.npz and were produced by
the REFERENCE's compiled core.pyx via tests/golden/make_golden.py.
"""
from __future__ import annotations

import numpy as np
import torch


def _randn(shape, seed, scale=1.0, shift=0.0):
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    return (torch.randn(*shape, generator=g, dtype=torch.float32) * scale + shift).numpy()


def _randint(lo, hi, shape, seed):
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    return torch.randint(lo, hi + 1, shape, generator=g).numpy()


def _i32(*v):
    return np.asarray(v, np.int32)


def case_rand_small():
    v = _randn((4, 13, 37), 11)
    return v, _i32(13, 7, 1, 12), _i32(37, 20, 5, 12)


def case_all_zero():
    # tie rule: strict '<' in the backtrack -> stay; 4x8 zeros -> durations [1,1,1,5]
    return np.zeros((1, 4, 8), np.float32), _i32(4), _i32(8)


def case_small_int_ties():
    v = _randint(-2, 2, (3, 17, 50), 12).astype(np.float32)
    return v, _i32(17, 9, 16), _i32(50, 31, 17)


def case_tx_eq_ty():
    v = _randn((2, 9, 9), 13)
    return v, _i32(9, 5), _i32(9, 5)


def case_tx_eq_1():
    v = _randn((2, 6, 11), 14)
    return v, _i32(1, 1), _i32(11, 1)


def case_tx_eq_ty_minus_1():
    v = _randn((2, 12, 13), 15)
    return v, _i32(12, 7), _i32(13, 8)


def case_padding_garbage():
    # t_y < Ty and t_x < Tx with NaN/inf/huge values in the padding: never read.
    v = _randn((3, 20, 40), 16)
    t_x, t_y = _i32(11, 20, 5), _i32(23, 39, 40)
    for b in range(3):
        v[b, t_x[b]:, :] = np.nan
        v[b, :, t_y[b]:] = np.inf
    v[1, 19, 39] = 7.0
    return v, t_x, t_y


def case_clamp_const():
    # accumulations fall below max_neg_val=-1e9: the x==0 / x==y substitutions become observable
    return np.full((1, 6, 12), -5e8, np.float32), _i32(6), _i32(12)


def case_clamp_random():
    v = _randn((2, 10, 30), 17, scale=2e8, shift=-4e8)
    return v, _i32(10, 6), _i32(30, 19)


def case_nan_cell():
    v = _randn((2, 8, 20), 18)
    v[0, 3, 9] = np.nan
    v[1, 0, 0] = np.nan
    return v, _i32(8, 8), _i32(20, 20)


def case_inf_cells():
    v = _randn((2, 8, 20), 19)
    v[0, 2, 7] = np.inf
    v[0, 5, 11] = -np.inf
    v[1, 4, 4] = -np.inf
    return v, _i32(8, 8), _i32(20, 18)


def case_signed_zeros():
    v = np.zeros((2, 7, 15), np.float32)
    v[0] = -0.0
    v[1, ::2, ::3] = -0.0
    return v, _i32(7, 7), _i32(15, 15)


def case_large_accum():
    # values ~ -1e4 over Ty=4096: large accumulations, fp32 rounding order matters
    v = _randn((1, 33, 4096), 20, scale=3e3, shift=-1e4)
    return v, _i32(33), _i32(4096)


def case_edge_33x65():
    v = _randn((2, 33, 65), 21)
    return v, _i32(33, 32), _i32(65, 64)


def case_edge_64x128():
    v = _randn((2, 64, 128), 22)
    return v, _i32(64, 63), _i32(128, 127)


def case_edge_65x129():
    v = _randn((2, 65, 129), 23)
    return v, _i32(65, 33), _i32(129, 97)


def case_edge_31x33():
    v = _randn((2, 31, 33), 24)
    return v, _i32(31, 2), _i32(33, 32)


def case_lrs2_shape():
    # one utterance at the LRS2 padded shape (Tx=190 not /32, Ty=1000 not /32)
    v = _randn((2, 190, 1000), 25, scale=30.0, shift=-1200.0)
    return v, _i32(190, 101), _i32(1000, 637)


def case_cfg1():
    # BASELINE.json configs[0]: B=16, Tx=200, Ty=800, N(0,1)
    import os
    import sys

    here = os.path.dirname(os.path.abspath(__file__))
    pkg = os.path.join(here, "..", "..", "face-gan-tts_b200")
    if pkg not in sys.path:
        sys.path.insert(0, pkg)
    from face_gan_tts_b200 import synthetic

    v, t_x, t_y = synthetic.mas_value(16, 200, 800, seed=1234)
    return v.numpy(), t_x.numpy(), t_y.numpy()


def case_wide_513x1030():
    # Tx > 512 rows: exercises the widest kernel instantiation
    v = _randn((1, 513, 1030), 26)
    return v, _i32(513), _i32(1030)


CASES = {
    "rand_small": case_rand_small,
    "all_zero": case_all_zero,
    "small_int_ties": case_small_int_ties,
    "tx_eq_ty": case_tx_eq_ty,
    "tx_eq_1": case_tx_eq_1,
    "tx_eq_ty_minus_1": case_tx_eq_ty_minus_1,
    "padding_garbage": case_padding_garbage,
    "clamp_const": case_clamp_const,
    "clamp_random": case_clamp_random,
    "nan_cell": case_nan_cell,
    "inf_cells": case_inf_cells,
    "signed_zeros": case_signed_zeros,
    "large_accum": case_large_accum,
    "edge_33x65": case_edge_33x65,
    "edge_64x128": case_edge_64x128,
    "edge_65x129": case_edge_65x129,
    "edge_31x33": case_edge_31x33,
    "lrs2_shape": case_lrs2_shape,
    "cfg1": case_cfg1,
    "wide_513x1030": case_wide_513x1030,
}


def compute_loss_inputs(B, Tx, Ty, out_size, seed, n_vocab, n_feats=128):
    """Inputs of one FaceTTS.compute_loss fixture case (tests/golden/make_compute_loss_golden.py), drawn from ONE seeded
    CPU generator in a fixed order -- shared by the generator script and by the tests, which regenerate `y` for the
    LRS2-sized case instead of storing 8 MB of noise in the fixture.
    Returns x_len, y_len (int64), x [B,Tx] int64, y [B,n_feats,Ty] float32, face [B,3,224,224] float32."""
    g = torch.Generator().manual_seed(seed)
    x_len = torch.randint(max(3, Tx // 3), Tx + 1, (B,), generator=g)
    x_len = x_len - (1 - x_len % 2)
    x_len[0] = Tx
    y_len = torch.stack([torch.randint(max(int(x_len[b]), Ty // 4), Ty + 1, (1,), generator=g)[0] for b in range(B)])
    y_len[0] = Ty
    if out_size is not None:
        y_len[1] = min(out_size - 20, Ty)            # one utterance shorter than the crop window
        x_len[1] = min(int(x_len[1]), int(y_len[1]))
    x = torch.randint(0, n_vocab - 1, (B, Tx), generator=g)
    x = x * (torch.arange(Tx)[None] < x_len[:, None])
    y = (torch.randn(B, n_feats, Ty, generator=g) * 2.0 - 5.0).clamp_(-11.512925, 2.0)
    y = y * (torch.arange(Ty)[None, None] < y_len[:, None, None])
    face = torch.rand(B, 3, 224, 224, generator=g) * 255.0
    return x_len, y_len, x, y, face
