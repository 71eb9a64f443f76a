"""Diagnostics: globaltimer timeline of one overlapped log-prior || MAS step (ns relative to the first stamp)."""
import os, sys
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "face-gan-tts_b200"))
import torch
from face_gan_tts_b200 import _lib, synthetic
B, F, TX, TY = 32, 80, 190, 1000
L = _lib.lib(); dev = torch.device("cuda", 0)
mu_x, y, t_x, t_y = synthetic.lrs2_batch(B, F, TX, TY, seed=1234)
d = dict(mu=mu_x.to(dev), y=y.to(dev), tx=t_x.to(dev), ty=t_y.to(dev), dur=torch.empty((B, TX), dtype=torch.int32, device=dev),
         ft=torch.empty((B, TY), dtype=torch.int32, device=dev), status=torch.empty((B,), dtype=torch.int32, device=dev),
         path=torch.empty((B, TX, TY), device=dev))
ws_bytes = L.mas_b200_fused_workspace_bytes(B, F, TX, TY)
ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev)
PREP = int(os.environ.get("PREPARED", "1"))
if PREP: assert L.mas_b200_fused_workspace_prepare(ws.data_ptr(), ws_bytes, B, F, TX, TY, None) == 0
sp = torch.cuda.current_stream(dev).cuda_stream
def call():
    rc = L.mas_b200_log_prior_maximum_path(d["mu"].data_ptr(), d["y"].data_ptr(), d["tx"].data_ptr(), d["ty"].data_ptr(), B, F, TX, TY, -1e9,
                                           d["path"].data_ptr(), _lib.PATH_F32, d["dur"].data_ptr(), d["ft"].data_ptr(), d["status"].data_ptr(), ws.data_ptr(), ws_bytes, _lib.LP_AUTO | (_lib.WS_PREPARED if PREP else 0), sp)
    assert rc == 0, rc
for _ in range(5): call()
torch.cuda.synchronize()
dm = torch.zeros((B, 16), dtype=torch.int64, device=dev); dl = torch.zeros((B * 8, 32), dtype=torch.int64, device=dev)
def setp(name, t):
    p = t.data_ptr(); lo, hi = p & 0xFFFFFFFF, p >> 32
    _lib.set_option(name + "_lo", lo - (1 << 32) if lo >= (1 << 31) else lo); _lib.set_option(name + "_hi", hi)
setp("mas_debug_ptr", dm); setp("lp_debug_ptr", dl)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); call(); e1.record(); torch.cuda.synchronize()
for n in ("mas_debug_ptr", "lp_debug_ptr"):
    _lib.set_option(n + "_lo", 0); _lib.set_option(n + "_hi", 0)
dm, dl = dm.cpu(), dl.cpu()
lp = dl[dl[:, 0] != 0]
t0 = min(int(lp[:, 0].min()), int(dm[:, 12].min()))
print(f"event time {e0.elapsed_time(e1)*1e3:.1f} us; {lp.shape[0]} LP CTAs")
print("LP  start  min/max us:", (int(lp[:, 0].min()) - t0) / 1e3, (int(lp[:, 0].max()) - t0) / 1e3)
print("LP  gemm end min/max :", (int(lp[:, 1].min()) - t0) / 1e3, (int(lp[:, 1].max()) - t0) / 1e3)
print("LP  prologue end min/max:", (int(lp[:, 4].min()) - t0) / 1e3, (int(lp[:, 4].max()) - t0) / 1e3)
import numpy as np
m = lp[:, 5:14].double().mean(0).tolist()
print("LP mean cycles/CTA: E drain-wait %.0f, E loop %.0f | S raw-wait %.0f, S bfree-wait %.0f, S loop %.0f | M split-wait %.0f, M dempty-wait %.0f, M loop %.0f, M aready-wait %.0f" % tuple(m))
print("LP E staging-free wait mean cycles/CTA: %.0f" % lp[:, 14].double().mean().item())
print("LP  path end min/max :", (int(lp[:, 2].min()) - t0) / 1e3, (int(lp[:, 2].max()) - t0) / 1e3)
print("MAS start  min/max   :", (int(dm[:, 12].min()) - t0) / 1e3, (int(dm[:, 12].max()) - t0) / 1e3)
print("MAS end    min/max   :", (int(dm[:, 13].min()) - t0) / 1e3, (int(dm[:, 13].max()) - t0) / 1e3)
print("MAS gate spins       :", dm[:8, 10].tolist())
i = int(torch.argmax(dm[:, 13]))
s_ = dm[i].tolist()
print(f"slowest MAS CTA b={i} t_x={s_[7] >> 32} t_y={s_[7] & 0xffffffff}: dp_done {s_[2]-s_[0]} all_warps {s_[4]-s_[0]} backtrack {s_[5]-s_[4]} tail {s_[6]-s_[5]} total {s_[6]-s_[0]} cycles; spins {s_[10]}")
print(f"  helper done {s_[14]-s_[0]}  producer done {s_[15]-s_[0]}; gate wait cycles total {s_[9]} (first group {s_[8]}); producer built upper transfer tables of {s_[11]} tiles in its slack")
# same batch through the serial pipeline (MAS ungated)
_lib.set_option("fused_impl", 1)
for _ in range(3): call()
dm2 = torch.zeros((B, 16), dtype=torch.int64, device=dev)
setp("mas_debug_ptr", dm2); call(); torch.cuda.synchronize()
_lib.set_option("mas_debug_ptr_lo", 0); _lib.set_option("mas_debug_ptr_hi", 0); _lib.set_option("fused_impl", 0)
s2 = dm2.cpu()[i].tolist()
print(f"ungated    b={i}: dp_done {s2[2]-s2[0]} all_warps {s2[4]-s2[0]} backtrack {s2[5]-s2[4]} tail {s2[6]-s2[5]} total {s2[6]-s2[0]}; helper done {s2[14]-s2[0]} producer done {s2[15]-s2[0]}")


# gated MAS with every flag preset (log-prior run before it): cost of the gating code path alone
_lib.set_option("fused_impl", 3)
for _ in range(3): call()
dm3 = torch.zeros((B, 16), dtype=torch.int64, device=dev)
setp("mas_debug_ptr", dm3); call(); torch.cuda.synchronize()
_lib.set_option("mas_debug_ptr_lo", 0); _lib.set_option("mas_debug_ptr_hi", 0); _lib.set_option("fused_impl", 0)
s3 = dm3.cpu()[i].tolist()
print(f"gated, flags preset b={i}: dp_done {s3[2]-s3[0]} all_warps {s3[4]-s3[0]} backtrack {s3[5]-s3[4]} tail {s3[6]-s3[5]} total {s3[6]-s3[0]}; spins {s3[10]} producer done {s3[15]-s3[0]}")

# inter-step gap: two back-to-back overlapped calls with separate stamp buffers
dlA = torch.zeros((B * 8, 32), dtype=torch.int64, device=dev); dlB = torch.zeros_like(dlA)
dmA = torch.zeros((B, 16), dtype=torch.int64, device=dev); dmB = torch.zeros_like(dmA)
for _ in range(3): call()
torch.cuda.synchronize()
setp("lp_debug_ptr", dlA); setp("mas_debug_ptr", dmA); call()
setp("lp_debug_ptr", dlB); setp("mas_debug_ptr", dmB); call()
torch.cuda.synchronize()
for n in ("mas_debug_ptr", "lp_debug_ptr"):
    _lib.set_option(n + "_lo", 0); _lib.set_option(n + "_hi", 0)
a, b2 = dlA.cpu(), dlB.cpu(); ma, mb = dmA.cpu(), dmB.cpu()
a = a[a[:, 0] != 0]; b2 = b2[b2[:, 0] != 0]
endA = max(int(a[:, 2].max()), int(a[:, 1].max()), int(ma[:, 13].max()))
print("back-to-back: step A LP start %.2f .. last stamp %.2f us; step B LP start %.2f (gap %.2f us), B first MAS start %.2f" % (
    0.0, (endA - int(a[:, 0].min())) / 1e3, (int(b2[:, 0].min()) - int(a[:, 0].min())) / 1e3,
    (int(b2[:, 0].min()) - endA) / 1e3, (int(mb[:, 12].min()) - int(a[:, 0].min())) / 1e3))
