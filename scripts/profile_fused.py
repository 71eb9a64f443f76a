#!/usr/bin/env python
"""Small driver for ncu: a few fused log-prior + MAS calls at the cfg2 shape (B=32, F=80, 190x1000)."""
import os, sys
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(ROOT, "face-gan-tts_b200"))
import torch
import face_gan_tts_b200 as fgt
from face_gan_tts_b200 import synthetic
B = int(os.environ.get("PB", "32")); F = int(os.environ.get("PF", "80"))
Tx = int(os.environ.get("PTX", "190")); Ty = int(os.environ.get("PTY", "1000"))
mu_x, y, t_x, t_y = synthetic.lrs2_batch(B, F, Tx, Ty, seed=1234)
mu_x, y = mu_x.cuda(), y.cuda()
for i in range(int(os.environ.get("PN", "4"))):
    r = fgt.log_prior_maximum_path(mu_x, y, t_x, t_y, dense_path=True)
torch.cuda.synchronize()
print("ok", int(r.durations.sum()))
