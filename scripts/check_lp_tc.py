#!/usr/bin/env python
"""GPU check of the tcgen05 log-prior against torch fp64 / the FFMA kernel (+ timing)."""
import os, sys
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "face-gan-tts_b200"))
import torch
import face_gan_tts_b200 as fgt
from face_gan_tts_b200 import synthetic

def ref64(mu, y):
    mu, y = mu.double(), y.double()
    F = mu.shape[1]
    return -0.5 * ((y.unsqueeze(2) - mu.unsqueeze(3)) ** 2).sum(1) - 0.5 * F * torch.log(torch.tensor(2 * torch.pi, dtype=torch.float64))

def t(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3

for (B, F, Tx, Ty) in [(2, 80, 61, 200), (4, 80, 190, 1000), (32, 80, 190, 1000), (3, 64, 129, 136), (2, 96, 256, 420), (32, 128, 190, 1000)]:
    mu, y, tx, ty = synthetic.lrs2_batch(B, F, Tx, Ty, seed=5, tx_lo=max(1, Tx // 3), ty_lo=max(Tx // 3, Ty // 3))
    mu, y = mu.cuda(), y.cuda()
    r = ref64(mu, y)
    for impl in ("ffma", "tcgen05"):
        try:
            o = fgt.log_prior(mu, y, impl=impl)
            torch.cuda.synchronize()
        except Exception as e:
            print(B, F, Tx, Ty, impl, "ERR", e); continue
        rel = ((o.double() - r).abs() / r.abs()).max().item()
        ab = (o.double() - r).abs().max().item()
        us = t(lambda: fgt.log_prior(mu, y, impl=impl))
        print(f"B={B} F={F} Tx={Tx} Ty={Ty} {impl:8s} max rel {rel:.3e} max abs {ab:.3e}  {us:.1f} us", flush=True)

# raw kernel timing through the C ABI (no allocation in the loop)
from face_gan_tts_b200 import _lib
L = _lib.lib()
for (B, F, Tx, Ty) in [(32, 80, 190, 1000), (64, 80, 190, 1000), (148, 80, 190, 1000), (16, 80, 200, 800)]:
    mu, y, tx, ty = synthetic.lrs2_batch(B, F, Tx, Ty, seed=5, tx_lo=Tx // 3, ty_lo=Ty // 3)
    mu, y = mu.cuda(), y.cuda()
    out = torch.empty((B, Tx, Ty), device="cuda")
    sp = torch.cuda.current_stream().cuda_stream
    for name, impl in (("ffma", 1), ("tcgen05", 2)):
        def call():
            rc = L.mas_b200_log_prior(mu.data_ptr(), y.data_ptr(), B, F, Tx, Ty, out.data_ptr(), impl, sp)
            assert rc == 0, rc
        print(f"raw B={B} F={F} Tx={Tx} Ty={Ty} {name:8s} {t(call, 50):.1f} us", flush=True)
