#!/usr/bin/env python
"""Throughput of the other BASELINE configs (parity-test shapes, not bench lines), one JSON line each:
cfg1 maximum_path drop-in (B=16, 200x800), cfg2 fused at F=80/128, cfg4 streamed stress (B=64, 512x4096),
cfg5 batch sweep B=64..1024 on one GPU.  CUDA events on the launching stream, buffer sets rotated so that the
footprint exceeds L2.  HBM roofline: algorithmic bytes / time against MEASURED_PEAKS.json."""
import json, os, sys
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "face-gan-tts_b200"))
import torch
import face_gan_tts_b200 as fgt
from face_gan_tts_b200 import monotonic_align, synthetic

try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    PEAK = 6650.0
dev = "cuda:0"


def timeit(fn, nsets, warm=5, reps=20):
    for i in range(warm):
        fn(i % nsets)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(reps):
        fn(i % nsets)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e-3


def nsets_for(bytes_per_set):
    return max(2, min(8, int(400e6 // max(bytes_per_set, 1)) + 1))


def report(name, sec, cells, alg_bytes, **kw):
    print(json.dumps({"config": name, "ms": sec * 1e3, "gcells_per_s": cells / sec / 1e9,
                      "algorithmic_GBps": alg_bytes / sec / 1e9, "hbm_frac": alg_bytes / sec / 1e9 / PEAK, **kw}), flush=True)


def run_mas(name, B, Tx, Ty, dense, via_mask=False):
    per = B * Tx * Ty * 4 * (2 if dense else 1)
    n = nsets_for(per)
    vals, tx, ty = [], None, None
    for s in range(n):
        v, tx, ty = synthetic.mas_value(B, Tx, Ty, seed=10 + s)
        vals.append(v.to(dev))
    if via_mask:
        mask = synthetic.prefix_mask(tx, ty, Tx, Ty).to(dev)
        fn = lambda i: monotonic_align.maximum_path(vals[i], mask)
    else:
        txd, tyd = tx.to(dev), ty.to(dev)
        fn = lambda i: fgt.align(vals[i], txd, tyd, dense_path=dense)
    sec = timeit(fn, n)
    report(name, sec, B * Tx * Ty, B * Tx * Ty * 4 * (2 if dense else 1) + (B * Tx * Ty * 4 if via_mask else 0),
           B=B, Tx=Tx, Ty=Ty, dense_path=dense, valid_cells=int((tx.long() * ty.long()).sum()), buffer_sets=n)


def run_fused(name, B, F, Tx, Ty, dense):
    per = 4 * F * B * (Tx + Ty) + B * Tx * Ty * 4 * (2 if dense else 1)
    n = nsets_for(per)
    sets = []
    for s in range(n):
        mu, y, tx, ty = synthetic.lrs2_batch(B, F, Tx, Ty, seed=20 + s)
        sets.append((mu.to(dev), y.to(dev), tx.to(dev), ty.to(dev)))
    fn = lambda i: fgt.log_prior_maximum_path(*sets[i], dense_path=dense)
    sec = timeit(fn, n)
    report(name, sec, B * Tx * Ty, 4 * F * B * (Tx + Ty) + (B * Tx * Ty * 4 if dense else 0), B=B, F=F, Tx=Tx, Ty=Ty,
           dense_path=dense, buffer_sets=n)


which = sys.argv[1:] or ["cfg1", "cfg2", "cfg4", "cfg5"]
if "cfg1" in which:
    run_mas("cfg1 maximum_path(value, mask) drop-in", 16, 200, 800, True, via_mask=True)
    run_mas("cfg1 align(value, lengths) dense", 16, 200, 800, True)
    run_mas("cfg1 align(value, lengths) durations only", 16, 200, 800, False)
if "cfg2" in which:
    run_fused("cfg2 fused F=80 dense", 32, 80, 190, 1000, True)
    run_fused("cfg2 fused F=80 durations only", 32, 80, 190, 1000, False)
    run_fused("cfg2 fused F=128 dense", 32, 128, 190, 1000, True)
if "cfg4" in which:
    run_mas("cfg4 streamed MAS durations only", 64, 512, 4096, False)
    run_mas("cfg4 streamed MAS dense", 64, 512, 4096, True)
    run_fused("cfg4 fused F=80 dense", 64, 80, 512, 4096, True)
if "cfg5" in which:
    for B in (64, 128, 256, 512, 1024):
        run_fused(f"cfg5 fused F=80 dense B={B}", B, 80, 190, 1000, True)
        run_fused(f"cfg5 fused F=80 durations only B={B}", B, 80, 190, 1000, False)
    for B in (148, 296, 1024):
        run_mas(f"cfg5 MAS only durations B={B}", B, 190, 1000, False)
