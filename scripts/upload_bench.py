#!/usr/bin/env python
"""Diagnostics for the e2e leg: PCIe rate of the ragged zero-copy upload vs CTA count, the padded copy-engine
transfer, and the host-side cost of one public-API call."""
import os, sys, time
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "face-gan-tts_b200"))
import torch
import face_gan_tts_b200 as fgt
from face_gan_tts_b200 import _lib, synthetic
B, F, TX, TY = 32, 80, 190, 1000
dev = torch.device("cuda", 0)
host = [[t.pin_memory() for t in synthetic.lrs2_batch(B, F, TX, TY, seed=4321 + k)] for k in range(3)]
out = fgt.upload_batch(*host[0])
valid = [4 * F * int((h[2].long() + ((h[3].long() + 3) // 4) * 4).sum()) for h in host]

def ev_time(fn, reps=50):
    for i in range(5): fn(i)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(reps): fn(i)
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e-3

_lib.set_option("upload_impl", 2)
s = ev_time(lambda i: fgt.upload_batch(*host[i % 3], out=out, mu_on_copy_engine=False))
print(f"ragged upload, copy engine 2-D per utterance: {s*1e6:7.1f} us  {sum(valid)/3/s/1e9:6.1f} GB/s over PCIe ({sum(valid)/3/1e6:.2f} MB valid)")
t0 = time.perf_counter()
for i in range(200): fgt.upload_batch(*host[i % 3], out=out)
t1 = time.perf_counter(); torch.cuda.synchronize()
print(f"  host enqueue {1e6*(t1-t0)/200:.1f} us/call")
_lib.set_option("upload_impl", 3)
s = ev_time(lambda i: fgt.upload_batch(*host[i % 3], out=out, mu_on_copy_engine=False))
print(f"ragged upload, y by TMA bulk copies (+ pull kernel for mu_x, tails, padding, serial): {s*1e6:7.1f} us  {sum(valid)/3/s/1e9:6.1f} GB/s")
_lib.set_option("upload_impl", 1)
for hint in (0,):
  _lib.set_option("upload_l2_256b", hint); print("L2::256B hint", hint)
  for ctas in (148,):
    _lib.set_option("upload_ctas", ctas)
    s = ev_time(lambda i: fgt.upload_batch(*host[i % 3], out=out, mu_on_copy_engine=False))
    print(f"  ragged upload ctas={ctas:5d}: {s*1e6:7.1f} us  {sum(valid)/3/s/1e9:6.1f} GB/s over PCIe ({sum(valid)/3/1e6:.2f} MB valid)")
_lib.set_option("upload_ctas", 0)
for m in (True, False):
    s = ev_time(lambda i: fgt.upload_batch(*host[i % 3], out=out, mu_on_copy_engine=m))
    print(f"ragged upload, mu_on_copy_engine={m}: {s*1e6:7.1f} us")
def padded(i):
    h = host[i % 3]
    for d, s in zip(out, h): d.copy_(s, non_blocking=True)
s = ev_time(padded)
print(f"padded cudaMemcpyAsync x4: {s*1e6:7.1f} us  {(4*F*B*(TX+TY))/s/1e9:6.1f} GB/s")
# host cost of the public API call (no sync inside the loop)
d = [t.to(dev) for t in host[0]]
for _ in range(10): fgt.log_prior_maximum_path(*d, dense_path=True)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(200): fgt.log_prior_maximum_path(*d, dense_path=True)
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print(f"fgt.log_prior_maximum_path host enqueue {1e6*(t1-t0)/200:.1f} us/call, with drain {1e6*(t2-t0)/200:.1f} us/call")
t0 = time.perf_counter()
for i in range(200): fgt.upload_batch(*host[i % 3], out=out)
t1 = time.perf_counter(); torch.cuda.synchronize()
print(f"fgt.upload_batch host enqueue {1e6*(t1-t0)/200:.1f} us/call")
