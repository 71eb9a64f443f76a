"""One fused log-prior + MAS call at the headline shape (for ncu): B=32, F=80, Tx=190, Ty=1000."""
import os, sys, torch
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(ROOT, "face-gan-tts_b200"))
import face_gan_tts_b200 as fgt
from face_gan_tts_b200 import synthetic

B = int(os.environ.get("FUSED_B", "32")); full = os.environ.get("FUSED_FULL", "0") == "1"
mu, y, tx, ty = synthetic.lrs2_batch(B=B, F=80, Tx=190, Ty=1000, seed=1234)
if full:
    tx[:] = 190; ty[:] = 1000
mu, y, tx, ty = mu.cuda(), y.cuda(), tx.cuda().int(), ty.cuda().int()
plan = fgt.AlignmentPlan(B, 80, 190, 1000, device="cuda:0", dense_path=os.environ.get("FUSED_DENSE", "0") == "1")
for _ in range(4):
    r = plan(mu, y, tx, ty)
torch.cuda.synchronize()
print("durations sum", int(r.durations.sum()))
