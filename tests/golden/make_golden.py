#!/usr/bin/env python
"""Generate tests/golden/mas_kats.npz from the REFERENCE's compiled core.pyx.

Run in the build container (needs /root/reference to build oracle/_ref, or a
prebuilt oracle/_ref):

    python tests/golden/make_golden.py

For every case in tests/golden/cases.py it stores
    <name>/t_x, <name>/t_y           the lengths fed to the reference
    <name>/input_sha256              sha256 of the float32 value bytes
    <name>/dur   [B,Tx] int32        row sums of the reference path
    <name>/frame_token [B,Ty] int32  x with path[x,y]==1 (-1 if none)
    <name>/path_sha256               sha256 of the int32 dense path bytes
The dense path is fully determined by frame_token (one 1 per valid frame),
so the fixture stays small while pinning the whole output.

The log-prior fixture (tests/golden/logprior_small.npz) holds a tiny
(mu_x, y) pair with the reference expression face_tts.py:165-171 evaluated by
torch fp32 on the CPU and the float64 direct form.
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import oracle  # noqa: E402
from oracle import build_ref  # noqa: E402
import cases  # noqa: E402


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    build_ref.build()
    core = build_ref.load("asis")
    if core is None:
        raise SystemExit("reference core not built (need /root/reference or oracle/_ref)")
    out = {}
    for name, fn in cases.CASES.items():
        value, t_x, t_y = fn()
        value = np.ascontiguousarray(value, np.float32)
        paths = np.zeros(value.shape, np.int32)
        work = value.copy()
        core.maximum_path_c(paths, work, t_x, t_y)      # reference core.pyx:40
        dur, ft = oracle.durations_and_frame_token(paths)
        out[f"{name}/t_x"] = t_x
        out[f"{name}/t_y"] = t_y
        out[f"{name}/shape"] = np.asarray(value.shape, np.int32)
        out[f"{name}/input_sha256"] = np.asarray(sha(value))
        out[f"{name}/dur"] = dur
        out[f"{name}/frame_token"] = ft
        out[f"{name}/path_sha256"] = np.asarray(sha(paths))
        print(f"{name:20s} shape={value.shape} ones={int(paths.sum())}")
    np.savez_compressed(os.path.join(HERE, "mas_kats.npz"), **out)

    # log-prior fixture
    import torch

    sys.path.insert(0, os.path.join(ROOT, "face-gan-tts_b200"))
    from face_gan_tts_b200 import synthetic

    mu_x, y, t_x, t_y = synthetic.lrs2_batch(B=2, F=80, Tx=21, Ty=36, seed=7, tx_lo=9, ty_lo=20)
    lp = oracle.log_prior_reference(mu_x, y).numpy()
    lp64 = oracle.log_prior_direct(mu_x, y).numpy()
    mu128, y128, _, _ = synthetic.lrs2_batch(B=1, F=128, Tx=9, Ty=16, seed=8, tx_lo=5, ty_lo=9)
    np.savez_compressed(
        os.path.join(HERE, "logprior_small.npz"),
        mu_x=mu_x.numpy(), y=y.numpy(), t_x=t_x.numpy(), t_y=t_y.numpy(),
        log_prior_ref_fp32=lp, log_prior_direct_fp64=lp64,
        mu_x_f128=mu128.numpy(), y_f128=y128.numpy(),
        log_prior_ref_fp32_f128=oracle.log_prior_reference(mu128, y128).numpy(),
    )
    print("wrote", os.path.join(HERE, "mas_kats.npz"), "and logprior_small.npz")


if __name__ == "__main__":
    main()
