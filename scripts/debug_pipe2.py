import os, sys, time
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "face-gan-tts_b200"))
import torch
from face_gan_tts_b200 import _lib, synthetic
B, F, TX, TY = int(os.environ.get("PB", "32")), 80, 190, 1000
L = _lib.lib()
dev = torch.device("cuda", 0)
mu_x, y, t_x, t_y = synthetic.lrs2_batch(B, F, TX, TY, seed=1234)
d = dict(mu=mu_x.to(dev), y=y.to(dev), tx=t_x.to(dev), ty=t_y.to(dev), dur=torch.empty((B, TX), dtype=torch.int32, device=dev),
         ft=torch.empty((B, TY), dtype=torch.int32, device=dev), status=torch.empty((B,), dtype=torch.int32, device=dev))
ws_bytes = L.mas_b200_fused_workspace_bytes(B, F, TX, TY)
ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev)
st = torch.cuda.Stream(dev) if os.environ.get("OWN_STREAM") else torch.cuda.current_stream(dev)
sp = st.cuda_stream
def call():
    rc = L.mas_b200_log_prior_maximum_path(d["mu"].data_ptr(), d["y"].data_ptr(), d["tx"].data_ptr(), d["ty"].data_ptr(), B, F, TX, TY, -1e9,
                                           None, _lib.PATH_NONE, d["dur"].data_ptr(), d["ft"].data_ptr(), d["status"].data_ptr(), ws.data_ptr(), ws_bytes, _lib.LP_AUTO, sp)
    assert rc == 0, rc
for mode in (1, 0):
    _lib.set_option("fused_impl", mode)
    for _ in range(5): call()
    torch.cuda.synchronize()
    a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ts = []
    for _ in range(10):
        a.record(st); call(); b_.record(st); torch.cuda.synchronize(); ts.append(a.elapsed_time(b_) * 1e3)
    print("fused_impl", mode, "(1 = serial lp->mas, 0 = overlapped)", " ".join(f"{t:.1f}" for t in ts), "us")
dbg = torch.zeros((B, 16), dtype=torch.int64, device=dev)
p = dbg.data_ptr(); lo, hi = p & 0xFFFFFFFF, p >> 32
_lib.set_option("mas_debug_ptr_lo", lo - (1 << 32) if lo >= (1 << 31) else lo); _lib.set_option("mas_debug_ptr_hi", hi)
call(); torch.cuda.synchronize()
_lib.set_option("mas_debug_ptr_lo", 0); _lib.set_option("mas_debug_ptr_hi", 0)
dd = dbg.cpu()
print("gate spins per CTA:", dd[:8, 10].tolist(), " total cycles:", (dd[:8, 6] - dd[:8, 0]).tolist())
