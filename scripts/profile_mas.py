#!/usr/bin/env python
"""Small driver for ncu: a few MAS launches at the cfg2 shape (B=32, 190x1000)."""
import os, sys
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(ROOT, "face-gan-tts_b200"))
import torch
import face_gan_tts_b200 as fgt
from face_gan_tts_b200 import synthetic, _lib
B = int(os.environ.get("PB", "32")); Tx = int(os.environ.get("PTX", "190")); Ty = int(os.environ.get("PTY", "1000"))
for k, v in os.environ.items():
    if k.startswith("MASOPT_"):
        _lib.set_option(k[7:].lower(), int(v))
v, t_x, t_y = synthetic.mas_value(B, Tx, Ty, seed=3, tx_lo=Tx // 3, ty_lo=Ty // 3)
v = v.cuda()
for i in range(int(os.environ.get("PN", "4"))):
    r = fgt.align(v, t_x, t_y, dense_path=bool(int(os.environ.get("PDENSE", "0"))))
torch.cuda.synchronize()
print("ok", int(r.durations.sum()))
