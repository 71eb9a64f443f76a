"""Fused kernel: value tiles vs torch fp32, path bit-exact vs the oracle MAS of those very values, timing vs the serial form (B200)."""
import os, sys, torch, numpy as np
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(ROOT, "face-gan-tts_b200")); sys.path.insert(0, ROOT)
import face_gan_tts_b200 as fgt
from face_gan_tts_b200 import synthetic, _lib
import oracle

dev = "cuda:0"

def set_ptr(name, t):
    _lib.set_pointer_option(name, t)

def run(B, F, Tx, Ty, dense, reps=30, check=True):
    mu, y, tx, ty = synthetic.lrs2_batch(B=B, F=F, Tx=Tx, Ty=Ty, seed=1234)
    mu, y, tx, ty = mu.to(dev), y.to(dev), tx.to(dev).int(), ty.to(dev).int()
    out, msg = {}, ""
    for mode in (0, 1):
        prev = _lib.set_option("fused_impl", mode)
        plan = fgt.AlignmentPlan(B, F, Tx, Ty, device=dev, dense_path=dense)
        if mode == 0 and check and B <= 64:
            dump = torch.full((B, Tx, Ty), float("nan"), device=dev)
            set_ptr("fused_dump_ptr", dump)
            r = plan(mu, y, tx, ty); torch.cuda.synchronize()
            set_ptr("fused_dump_ptr", None)
            ref = oracle.log_prior_reference(mu, y)
            txc, tyc = tx.cpu().numpy(), ty.cpu().numpy()
            rel = 0.0
            for b in range(B):
                a, c = dump[b, :txc[b], :tyc[b]].double(), ref[b, :txc[b], :tyc[b]].double()
                rel = max(rel, ((a - c).abs() / c.abs().clamp_min(1e-30)).max().item())
            own = np.zeros((B, Tx, Ty), np.int32)
            oracle.maximum_path_c(own, torch.nan_to_num(dump).cpu().numpy().copy(), txc, tyc)
            dur, ft = oracle.durations_and_frame_token(own)
            ok = np.array_equal(r.durations.cpu().numpy(), dur) and np.array_equal(r.frame_token.cpu().numpy(), ft)
            if dense: ok = ok and np.array_equal(r.path.cpu().numpy().astype(np.int32), own)
            msg = f" value rel err {rel:.2e}  path==oracle(own value): {ok}"
        r = plan(mu, y, tx, ty); torch.cuda.synchronize()
        res = (r.durations.clone(), r.frame_token.clone())
        for _ in range(5): plan(mu, y, tx, ty)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps): plan(mu, y, tx, ty)
        e1.record(); torch.cuda.synchronize()
        out[mode] = (res, e0.elapsed_time(e1) / reps * 1e3)
        _lib.set_option("fused_impl", prev)
    agree = (out[0][0][1] == out[1][0][1]).float().mean().item() * 100
    print(f"B={B} F={F} Tx={Tx} Ty={Ty} dense={dense}: fused {out[0][1]:.1f} us  serial {out[1][1]:.1f} us  frames agreeing with serial {agree:.3f}%{msg}", flush=True)

if __name__ == "__main__":
    run(32, 80, 190, 1000, False)
    run(32, 80, 190, 1000, True)
    run(8, 80, 100, 400, False)
    run(32, 64, 190, 1000, False)
    run(32, 128, 128, 1000, False)
    run(32, 96, 100, 600, False)
    run(148, 80, 190, 1000, False)
    run(296, 80, 190, 1000, True)
    run(1024, 80, 190, 1000, False)
    run(1024, 80, 190, 1000, True)
