#!/usr/bin/env python
"""Diagnostics: per-CTA clock64 phase stamps of the fused log-prior + MAS kernel.
dbg: [0] start, [1] A parked in TMEM (prologue done), [2] first value tile in the ring (DP starts), [4] last DP warp
done, [3] all roles joined, [5] backtrack done, [6] end, [12]/[13] globaltimer at start / end."""
import os, sys
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(ROOT, "face-gan-tts_b200"))
import torch
import face_gan_tts_b200 as fgt
from face_gan_tts_b200 import synthetic, _lib

def run(B, F, Tx, Ty, full=False):
    mu, y, tx, ty = synthetic.lrs2_batch(B=B, F=F, Tx=Tx, Ty=Ty, seed=1234)
    if full:
        tx[:] = Tx; ty[:] = Ty
    mu, y, tx, ty = mu.cuda(), y.cuda(), tx.cuda().int(), ty.cuda().int()
    plan = fgt.AlignmentPlan(B, F, Tx, Ty, device="cuda:0", dense_path=os.environ.get("FUSED_DENSE", "0") == "1")
    for _ in range(3):
        plan(mu, y, tx, ty)
    dbg = torch.zeros((2 * B, 32), dtype=torch.int64, device="cuda")
    _lib.set_pointer_option("mas_debug_ptr", dbg)
    plan(mu, y, tx, ty)
    torch.cuda.synchronize()
    _lib.set_pointer_option("mas_debug_ptr", None)
    d = dbg.cpu()
    print(f"--- fused B={B} F={F} Tx={Tx} Ty={Ty} full={full}")
    print("  b  t_x   t_y | (setup mu_load park_A) prologue first_tile   dp_total  cyc/frame |  join backtrack (chain sync walk sync) outputs (heads) | total cyc   wall us")
    t0 = int(d[:B, 12].min())
    for b in list(range(min(B, 8))) + ([B - 1] if B > 8 else []):
        s = d[b].tolist()
        txb, tyb = s[7] >> 32, s[7] & 0xFFFFFFFF
        print(f"{b:3d} {txb:4d} {tyb:5d} | ({s[8]-s[0]:5d} {s[9]-s[8]:5d} {s[1]-s[9]:5d}) {s[1]-s[0]:8d} {s[2]-s[0]:10d} {s[4]-s[2]:10d} {(s[4]-s[2])/max(tyb,1):10.1f} | "
              f"{s[3]-s[4]:5d} {s[5]-s[3]:9d} ({s[10]-s[3]:5d} {s[14]-s[10]:4d} {s[15]-s[14]:4d} {s[5]-s[15]:4d}) {s[6]-s[5]:7d} ({s[11]-s[5]:5d}) | {s[6]-s[0]:9d} {(s[13]-s[12])/1e3:8.1f}  (start +{(s[12]-t0)/1e3:.1f} us)  dp0/helpA/helpB done at {s[30]-s[0]} {s[28]-s[0]} {s[29]-s[0]}, DP done {s[4]-s[0]}, walk rounds of 8 for {s[31]} tokens")
    if int(d[B, 0]) != 0:      # pair kernel: the second CTA of each utterance
        for b in range(min(B, 4)):
            s = d[B + b].tolist(); r0 = d[b].tolist()
            if s[0]:
                print(f"    rank 1 of {b}: start {s[0]-r0[0]:+d} vs rank 0 (different SM clocks!), first tile at {s[2]-s[0]}, DP {s[4]-s[2]} cycles, joined at {s[3]-s[0]}; rank 0: tile 0 released at {r0[26]-r0[0]}, forwarded at {r0[27]-r0[0]}")
    print("  wait cycles per frame:  mma<-split mma<-dempty | split<-bfree split<-raw | epi<-dfull epi<-ring_empty | dp0<-ring_full dp0<-flag dp1<-ring_full dp1<-flag | dp0 body dp1 body")
    for b in list(range(min(B, 8))):
        s = d[b].tolist(); tyb = max(s[7] & 0xFFFFFFFF, 1)
        print(f"{b:3d} " + " ".join(f"{s[k]/tyb:9.1f}" for k in range(16, 28)))
    print(f"  kernel span (globaltimer): {(int(d[:B,13].max())-t0)/1e3:.1f} us")

if __name__ == "__main__":
    if os.environ.get("FUSED_EXP"):      # MASB200_PROF builds: role-parking experiments (option fused_exp)
        for e in [int(v) for v in os.environ["FUSED_EXP"].split(",")]:
            _lib.set_option("fused_exp", e)
            print(f"=== fused_exp = {e}")
            run(32, 80, 190, 1000, full=True)
        sys.exit(0)
    run(32, 80, 190, 1000)
    run(32, 80, 190, 1000, full=True)
    run(8, 80, 100, 400)
    run(296, 80, 190, 1000)
