import os, sys
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "face-gan-tts_b200"))
import torch
import face_gan_tts_b200 as fgt
from face_gan_tts_b200 import _lib, synthetic
skip = int(sys.argv[1]) if len(sys.argv) > 1 else 0
B, F, Tx, Ty = 2, 80, 190, 256
mu, y, _, _ = synthetic.lrs2_batch(B, F, Tx, Ty, seed=5, tx_lo=50, ty_lo=200)
_lib.set_option("lp_debug_skip", skip)
try:
    out = fgt.log_prior(mu.cuda(), y.cuda(), impl="tcgen05")
    torch.cuda.synchronize()
    ref = fgt.log_prior(mu.cuda(), y.cuda(), impl="ffma")
    torch.cuda.synchronize()
    print("skip", skip, "ok; max abs diff", (out - ref).abs().max().item())
except Exception as e:
    print("skip", skip, "FAILED", repr(e)[:300])
