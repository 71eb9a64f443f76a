#!/usr/bin/env python
"""Diagnostics: per-CTA clock64 phase stamps of the MAS kernel (start, first tile ready, DP end,
ring-full wait cycles, backtrack end, end)."""
import os, sys
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(ROOT, "face-gan-tts_b200"))
import torch
import face_gan_tts_b200 as fgt
from face_gan_tts_b200 import synthetic, _lib

def run(B, Tx, Ty, opts, dense=False):
    for k, v in opts.items():
        _lib.set_option(k, v)
    v, t_x, t_y = synthetic.mas_value(B, Tx, Ty, seed=3, tx_lo=Tx // 3, ty_lo=Ty // 3)
    v = v.cuda()
    dbg = torch.zeros((B, 16), dtype=torch.int64, device="cuda")
    for _ in range(3):
        fgt.align(v, t_x, t_y, dense_path=dense)
    _lib.set_pointer_option("mas_debug_ptr", dbg)
    fgt.align(v, t_x, t_y, dense_path=dense)
    torch.cuda.synchronize()
    _lib.set_pointer_option("mas_debug_ptr", None)
    d = dbg.cpu()
    print(f"--- B={B} Tx={Tx} Ty={Ty} opts={opts}")
    print(" b   t_x  t_y |  dp_total  cyc/frame | warps_done backtrack outputs | total cyc")
    for b in list(range(min(B, 6))) + ([B - 1] if B > 6 else []):
        s = d[b].tolist()
        tx, ty = s[7] >> 32, s[7] & 0xFFFFFFFF
        print(f"{b:3d} {tx:5d} {ty:5d} | {s[2]-s[0]:9d} {(s[2]-s[0])/max(ty,1):10.1f} | "
              f"{s[4]-s[2]:9d} {s[5]-s[4]:9d} {s[6]-s[5]:7d} | {s[6]-s[0]:9d}")
    for k in opts:
        _lib.set_option(k, 0)

run(32, 190, 1000, {})
run(32, 190, 1000, {"mas_rows_per_lane": 8, "mas_dp_warps": 1})
run(32, 190, 1000, {"mas_rows_per_lane": 4, "mas_dp_warps": 2})
run(64, 512, 4096, {})
