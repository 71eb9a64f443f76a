#!/usr/bin/env python
"""Diagnostics: which phase bounds the tcgen05 log-prior kernel -- time it with phases switched off (lp_debug_skip)."""
import os, sys
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "face-gan-tts_b200"))
import torch
import face_gan_tts_b200 as fgt
from face_gan_tts_b200 import _lib, synthetic
for B, F in ((32, 80), (1024, 80), (32, 128)):
    n = 6 if B == 32 else 2
    sets = []
    for s in range(n):
        mu, y, _, _ = synthetic.lrs2_batch(B, F, 190, 1000, seed=5 + s)
        sets.append((mu.cuda(), y.cuda()))
    for skip, name in ((0, "full"), (1, "no global stores"), (9, "no staging + no stores"), (2, "no MMA"), (4, "no split math"),
                       (15, "waits only")):
        _lib.set_option("lp_debug_skip", skip)
        for i in range(3): fgt.log_prior(*sets[i % n], impl="tcgen05")
        torch.cuda.synchronize()
        a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(20): fgt.log_prior(*sets[i % n], impl="tcgen05")
        b_.record(); torch.cuda.synchronize()
        print(f"B={B} F={F} {name:24s}: {a.elapsed_time(b_)/20*1e3:8.1f} us")
    _lib.set_option("lp_debug_skip", 0)
