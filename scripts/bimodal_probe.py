#!/usr/bin/env python
"""Diagnostics: run-to-run spread of the fused call (dense path / durations only), direct C calls, rotating sets."""
import os, sys
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "face-gan-tts_b200"))
import torch
from face_gan_tts_b200 import _lib, synthetic
B, F, TX, TY = 32, 80, 190, 1000
L = _lib.lib(); dev = torch.device("cuda", 0); NS = 6
sets = []
for s in range(NS):
    mu_x, y, t_x, t_y = synthetic.lrs2_batch(B, F, TX, TY, seed=1234 + s)
    sets.append(dict(mu=mu_x.to(dev), y=y.to(dev), tx=t_x.to(dev), ty=t_y.to(dev), dur=torch.empty((B, TX), dtype=torch.int32, device=dev),
                     ft=torch.empty((B, TY), dtype=torch.int32, device=dev), status=torch.empty((B,), dtype=torch.int32, device=dev),
                     path=torch.empty((B, TX, TY), device=dev)))
ws_bytes = L.mas_b200_fused_workspace_bytes(B, F, TX, TY)
wss = [torch.empty((ws_bytes,), dtype=torch.uint8, device=dev) for _ in range(NS)]
sp = torch.cuda.current_stream(dev).cuda_stream
for w in wss: assert L.mas_b200_fused_workspace_prepare(w.data_ptr(), ws_bytes, B, F, TX, TY, None) == 0
torch.cuda.synchronize()
PREP = [0]
def call(i, dense):
    d = sets[i]
    rc = L.mas_b200_log_prior_maximum_path(d["mu"].data_ptr(), d["y"].data_ptr(), d["tx"].data_ptr(), d["ty"].data_ptr(), B, F, TX, TY, -1e9,
                                           d["path"].data_ptr() if dense else None, _lib.PATH_F32 if dense else _lib.PATH_NONE,
                                           d["dur"].data_ptr(), d["ft"].data_ptr(), d["status"].data_ptr(), wss[i].data_ptr(), ws_bytes, _lib.LP_AUTO | (_lib.WS_PREPARED if PREP[0] else 0), sp)
    assert rc == 0
for dense, prep in ((True, 0), (True, 1), (False, 0), (False, 1), (True, 0), (True, 1)):
    PREP[0] = prep
    ts = []
    for r in range(8):
        for i in range(6): call(i % NS, dense)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(100): call(i % NS, dense)
        b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 10)
    print("prepared ws (no memset)" if prep else "memset per call        ", "dense" if dense else "durations only", " ".join(f"{t:.1f}" for t in ts), "us/step")
