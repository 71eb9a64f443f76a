import os, sys, time
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "face-gan-tts_b200"))
import torch
from face_gan_tts_b200 import _lib, synthetic
B, F, TX, TY = int(os.environ.get("PB", "8")), 80, 190, 1000
DENSE = int(os.environ.get("DENSE", "1"))
L = _lib.lib()
dev = torch.device("cuda", 0)
mu_x, y, t_x, t_y = synthetic.lrs2_batch(B, F, TX, TY, seed=1234)
d = dict(mu=mu_x.to(dev), y=y.to(dev), tx=t_x.to(dev), ty=t_y.to(dev), dur=torch.zeros((B, TX), dtype=torch.int32, device=dev),
         ft=torch.empty((B, TY), dtype=torch.int32, device=dev), status=torch.full((B,), -7, dtype=torch.int32, device=dev),
         path=torch.full((B, TX, TY), -3.0, device=dev))
ws_bytes = L.mas_b200_fused_workspace_bytes(B, F, TX, TY)
ws = torch.zeros((ws_bytes,), dtype=torch.uint8, device=dev)
sp = torch.cuda.current_stream(dev).cuda_stream
dbg = torch.zeros((B, 16), dtype=torch.int64, device=dev)
p = dbg.data_ptr(); lo, hi = p & 0xFFFFFFFF, p >> 32
_lib.set_option("mas_debug_ptr_lo", lo - (1 << 32) if lo >= (1 << 31) else lo); _lib.set_option("mas_debug_ptr_hi", hi)
t0 = time.time()
rc = L.mas_b200_log_prior_maximum_path(d["mu"].data_ptr(), d["y"].data_ptr(), d["tx"].data_ptr(), d["ty"].data_ptr(), B, F, TX, TY, -1e9,
                                       d["path"].data_ptr() if DENSE else None, _lib.PATH_F32 if DENSE else _lib.PATH_NONE, d["dur"].data_ptr(), d["ft"].data_ptr(), d["status"].data_ptr(), ws.data_ptr(), ws_bytes, _lib.LP_AUTO, sp)
torch.cuda.synchronize()
print("rc", rc, f"{(time.time()-t0)*1e3:.1f} ms")
mas_ws = (L.mas_b200_workspace_bytes(B, TX, TY) + 255) // 256 * 256
val_bytes = (4 * B * TX * TY + 255) // 256 * 256

print("B", B, "dense", DENSE, "spins", dbg[:8, 10].tolist())
print("path sum", float(d["path"].sum()), "min", float(d["path"].min()))
