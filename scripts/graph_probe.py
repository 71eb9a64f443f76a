#!/usr/bin/env python
"""Probe: is the two-stream overlapped pipeline capturable in a CUDA graph, and what does replay cost per step?"""
import os, sys, time
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "face-gan-tts_b200"))
import torch
from face_gan_tts_b200 import _lib, synthetic
B, F, TX, TY = 32, 80, 190, 1000
L = _lib.lib(); dev = torch.device("cuda", 0)
NS = 6
sets = []
for s in range(NS):
    mu_x, y, t_x, t_y = synthetic.lrs2_batch(B, F, TX, TY, seed=1234 + s)
    sets.append(dict(mu=mu_x.to(dev), y=y.to(dev), tx=t_x.to(dev), ty=t_y.to(dev), dur=torch.empty((B, TX), dtype=torch.int32, device=dev),
                     ft=torch.empty((B, TY), dtype=torch.int32, device=dev), status=torch.empty((B,), dtype=torch.int32, device=dev),
                     path=torch.empty((B, TX, TY), device=dev)))
ws_bytes = L.mas_b200_fused_workspace_bytes(B, F, TX, TY)
wss = [torch.empty((ws_bytes,), dtype=torch.uint8, device=dev) for _ in range(NS)]
def call(i, stream):
    d = sets[i]
    rc = L.mas_b200_log_prior_maximum_path(d["mu"].data_ptr(), d["y"].data_ptr(), d["tx"].data_ptr(), d["ty"].data_ptr(), B, F, TX, TY, -1e9,
                                           d["path"].data_ptr(), _lib.PATH_F32, d["dur"].data_ptr(), d["ft"].data_ptr(), d["status"].data_ptr(),
                                           wss[i].data_ptr(), ws_bytes, _lib.LP_AUTO, stream.cuda_stream)
    assert rc == 0, rc
cur = torch.cuda.current_stream(dev)
for i in range(NS): call(i, cur)
torch.cuda.synchronize()
want = [(d["dur"].clone(), d["ft"].clone(), d["path"].clone()) for d in sets]
def timeit(fn, n=200):
    for i in range(12): fn(i % NS)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); t0 = time.perf_counter()
    for i in range(n): fn(i % NS)
    t1 = time.perf_counter(); b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3, (t1 - t0) / n * 1e6
print("eager: %.1f us/step (host enqueue %.1f us)" % timeit(lambda i: call(i, torch.cuda.current_stream(dev))))
graphs = []
cap = torch.cuda.Stream(dev)
cap.wait_stream(cur)
for i in range(NS):
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=cap):
        call(i, torch.cuda.current_stream(dev))
    graphs.append(g)
torch.cuda.synchronize()
print("captured", len(graphs), "graphs")
for d in sets:
    d["dur"].zero_(); d["ft"].zero_(); d["path"].zero_()
for i in range(NS): graphs[i].replay()
torch.cuda.synchronize()
ok = all(torch.equal(d["dur"], w[0]) and torch.equal(d["ft"], w[1]) and torch.equal(d["path"], w[2]) for d, w in zip(sets, want))
print("replay results identical to eager:", ok)
print("graph replay: %.1f us/step (host enqueue %.1f us)" % timeit(lambda i: graphs[i].replay()))
# one graph holding all NS steps back to back
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g, stream=cap):
    for i in range(NS): call(i, torch.cuda.current_stream(dev))
torch.cuda.synchronize()
t, h = timeit(lambda i: g.replay(), n=50)
print("graph of %d steps: %.1f us/step (host enqueue %.1f us per replay)" % (NS, t / NS, h))
