#!/usr/bin/env python
"""Soak: N back-to-back overlapped fused calls (direct C calls, rotating sets), results checked against the serial
pipeline every 1000 steps.  Any protocol race shows up as a mismatch or a trap (bounded waits)."""
import os, sys, time
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "face-gan-tts_b200"))
import torch
from face_gan_tts_b200 import _lib, synthetic
N = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
B, F, TX, TY = 32, int(os.environ.get("PF", "80")), 190, 1000
L = _lib.lib(); dev = torch.device("cuda", 0); NS = 6
sets = []
for s in range(NS):
    mu_x, y, t_x, t_y = synthetic.lrs2_batch(B, F, TX, TY, seed=77 + s)
    sets.append(dict(mu=mu_x.to(dev), y=y.to(dev), tx=t_x.to(dev), ty=t_y.to(dev), dur=torch.empty((B, TX), dtype=torch.int32, device=dev),
                     ft=torch.empty((B, TY), dtype=torch.int32, device=dev), status=torch.empty((B,), dtype=torch.int32, device=dev),
                     path=torch.empty((B, TX, TY), device=dev)))
ws_bytes = L.mas_b200_fused_workspace_bytes(B, F, TX, TY)
wss = [torch.empty((ws_bytes,), dtype=torch.uint8, device=dev) for _ in range(NS)]
sp = torch.cuda.current_stream(dev).cuda_stream
def call(i):
    d = sets[i]
    rc = L.mas_b200_log_prior_maximum_path(d["mu"].data_ptr(), d["y"].data_ptr(), d["tx"].data_ptr(), d["ty"].data_ptr(), B, F, TX, TY, -1e9,
                                           d["path"].data_ptr(), _lib.PATH_F32, d["dur"].data_ptr(), d["ft"].data_ptr(), d["status"].data_ptr(),
                                           wss[i].data_ptr(), ws_bytes, _lib.LP_AUTO, sp)
    assert rc == 0
_lib.set_option("fused_impl", 1)
for i in range(NS): call(i)
torch.cuda.synchronize()
want = [(d["dur"].clone(), d["ft"].clone(), d["path"].clone()) for d in sets]
_lib.set_option("fused_impl", 0)
t0 = time.time(); bad = 0
for step in range(N):
    call(step % NS)
    if (step + 1) % 1000 == 0:
        torch.cuda.synchronize()
        for d, w in zip(sets, want):
            if not (torch.equal(d["dur"], w[0]) and torch.equal(d["ft"], w[1]) and torch.equal(d["path"], w[2])): bad += 1
        for d in sets: d["dur"].zero_(); d["ft"].zero_()
        for i in range(NS): call(i)      # refill after zeroing so the next check sees fresh results
torch.cuda.synchronize()
print(f"soak F={F}: {N} overlapped steps in {time.time()-t0:.1f} s, mismatching checks: {bad}")
