#!/usr/bin/env python
"""GPU experiment: time the MAS kernel under every (rows-per-lane, DP-warps, cell impl) plan
for the BASELINE shapes, plus the other kernels.  Writes gpurun_out/tune_mas.json.

    python scripts/tune_mas.py [--quick]
"""
import json
import os
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "face-gan-tts_b200"))

import torch  # noqa: E402

import face_gan_tts_b200 as fgt  # noqa: E402
from face_gan_tts_b200 import _lib, synthetic  # noqa: E402

DEV = "cuda:0"


def time_call(fn, reps=30, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3      # us


def mas_case(B, Tx, Ty, tx_lo, ty_lo, plans, results, tag, nsets=4):
    sets = []
    for s in range(nsets):
        v, t_x, t_y = synthetic.mas_value(B, Tx, Ty, seed=10 + s, tx_lo=tx_lo, ty_lo=ty_lo)
        sets.append((v.to(DEV), t_x.to(DEV), t_y.to(DEV)))
    L = _lib.lib()
    ws_bytes = L.mas_b200_workspace_bytes(B, Tx, Ty)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=DEV)
    dur = torch.empty((B, Tx), dtype=torch.int32, device=DEV)
    ft = torch.empty((B, Ty), dtype=torch.int32, device=DEV)
    st = torch.empty((B,), dtype=torch.int32, device=DEV)
    path = torch.empty((B, Tx, Ty), dtype=torch.float32, device=DEV)
    sp = torch.cuda.current_stream().cuda_stream
    it = [0]

    def call(dense):
        v, t_x, t_y = sets[it[0] % nsets]
        it[0] += 1
        rc = L.mas_b200_maximum_path(v.data_ptr(), Tx * Ty, Ty, t_x.data_ptr(), t_y.data_ptr(), B, Tx, Ty, -1e9,
                                     path.data_ptr() if dense else None, 1 if dense else 0, dur.data_ptr(),
                                     ft.data_ptr(), st.data_ptr(), ws.data_ptr(), ws_bytes, sp)
        assert rc == 0, rc

    for (R, W, extra) in plans:
        opts = dict(mas_rows_per_lane=R, mas_dp_warps=W)
        opts.update(extra)
        prev = {k: _lib.set_option(k, val) for k, val in opts.items()}
        try:
            us = time_call(lambda: call(False))
            us_dense = time_call(lambda: call(True))
            err = None
        except Exception as e:  # unsupported plan
            us = us_dense = None
            err = repr(e)
        for k, val in prev.items():
            _lib.set_option(k, val)
        rec = dict(tag=tag, B=B, Tx=Tx, Ty=Ty, R=R, W=W, extra=extra, mas_us=us, mas_dense_us=us_dense,
                   gcells_s=(B * Tx * Ty / us / 1e3) if us else None, err=err)
        results.append(rec)
        print(rec, flush=True)


def main():
    quick = "--quick" in sys.argv
    results = []
    plans_190 = [(R, W, {}) for (R, W) in [(2, 3), (4, 2), (2, 4), (8, 1), (4, 3), (1, 4)]]
    plans_190 += [(4, 2, {"mas_ring_stages": 2}), (4, 2, {"mas_ring_stages": 3}),
                  (4, 2, {"mas_force_global_bits": 1}), (4, 2, {"mas_force_unaligned": 1})]
    mas_case(32, 190, 1000, 60, 300, plans_190, results, "cfg2-shape B=32")
    mas_case(16, 200, 800, 100, 400, [(4, 2, {}), (2, 4, {}), (8, 1, {})], results, "cfg1 B=16")
    if not quick:
        big = [(R, W, {"mas_ctas_per_sm": k, "mas_fused_path_write": 0}) for (R, W) in [(2, 3), (4, 2), (8, 1)]
               for k in (1, 2, 3, 4)]
        big += [(4, 2, {"mas_ctas_per_sm": 3, "mas_fused_path_write": 1})]
        mas_case(1024, 190, 1000, 60, 300, big, results, "cfg5 B=1024", nsets=1)
        mas_case(64, 512, 4096, 256, 2048, [(4, 4, {}), (8, 2, {}), (8, 4, {})], results, "cfg4 B=64", nsets=1)

    # log-prior kernels
    for (B, F, Tx, Ty) in [(32, 80, 190, 1000), (32, 128, 190, 1000)]:
        mu_x, y, t_x, t_y = synthetic.lrs2_batch(B, F, Tx, Ty, seed=1)
        mu_d, y_d = mu_x.to(DEV), y.to(DEV)
        for impl in ("ffma", "auto"):
            us = time_call(lambda: fgt.log_prior(mu_d, y_d, impl=impl))
            rec = dict(tag="log_prior", B=B, F=F, Tx=Tx, Ty=Ty, impl=impl, us=us,
                       gflops=2 * F * B * Tx * Ty / us / 1e3)
            results.append(rec)
            print(rec, flush=True)
        us = time_call(lambda: fgt.log_prior_maximum_path(mu_d, y_d, t_x, t_y))
        rec = dict(tag="fused_api", B=B, F=F, Tx=Tx, Ty=Ty, us=us, gcells_s=B * Tx * Ty / us / 1e3)
        results.append(rec)
        print(rec, flush=True)

    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(results, open(os.path.join(ROOT, "gpurun_out", "tune_mas.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
