"""Fused kernel vs serial form: equality + CUDA-event timing (B200)."""
import sys, torch, time
sys.path.insert(0, "face-gan-tts_b200")
import face_gan_tts_b200 as fgt
from face_gan_tts_b200 import synthetic, _lib

dev = "cuda:0"
def run(B, F, Tx, Ty, dense, reps=30):
    mu, y, tx, ty = synthetic.lrs2_batch(B=B, F=F, Tx=Tx, Ty=Ty, seed=1234)
    mu, y, tx, ty = mu.to(dev), y.to(dev), tx.to(dev).int(), ty.to(dev).int()
    out = {}
    for mode in (0, 1):
        prev = _lib.set_option("fused_impl", mode)
        plan = fgt.AlignmentPlan(B, F, Tx, Ty, device=dev, dense_path=dense)
        r = plan(mu, y, tx, ty)
        torch.cuda.synchronize()
        res = (r.durations.clone(), r.frame_token.clone(), r.status.clone(), r.path.clone() if dense else None)
        for _ in range(5): plan(mu, y, tx, ty)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps): plan(mu, y, tx, ty)
        e1.record(); torch.cuda.synchronize()
        out[mode] = (res, e0.elapsed_time(e1) / reps * 1e3)
        _lib.set_option("fused_impl", prev)
    a, b = out[0][0], out[1][0]
    eq = all(torch.equal(p, q) for p, q in zip(a[:3], b[:3])) and (not dense or torch.equal(a[3], b[3]))
    print(f"B={B} F={F} Tx={Tx} Ty={Ty} dense={dense}: fused {out[0][1]:.1f} us  serial {out[1][1]:.1f} us  equal={eq}", flush=True)

if __name__ == "__main__":
    run(32, 80, 190, 1000, False)
    run(32, 80, 190, 1000, True)
    run(8, 80, 100, 400, False)
    run(32, 64, 190, 1000, False)
    run(32, 96, 256, 1000, False)
    run(148, 80, 190, 1000, False)
    run(296, 80, 190, 1000, True)
    run(1024, 80, 190, 1000, False)
    run(1024, 80, 190, 1000, True)
