"""Per-step device time of the headline step right after a synchronize (what a --steps 20 bench window looks like from
inside): 6 rotating buffer sets, one event per step."""
import os, sys, torch
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, os.path.join(ROOT, "face-gan-tts_b200"))
import face_gan_tts_b200 as fgt
from face_gan_tts_b200 import synthetic

B, N = 32, 6
sets = []
for s in range(N):
    mu, y, tx, ty = synthetic.lrs2_batch(B=B, F=80, Tx=190, Ty=1000, seed=1234 + s)
    sets.append((mu.cuda(), y.cuda(), tx.cuda().int(), ty.cuda().int(), fgt.AlignmentPlan(B, 80, 190, 1000, device="cuda:0", dense_path=True)))
st = torch.cuda.current_stream()
for rep in range(3):
    for i in range(5):
        m, y, a, b, p = sets[i % N]; p(m, y, a, b)
    torch.cuda.synchronize()
    K = 24
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(K + 1)]
    ev[0].record(st)
    for i in range(K):
        m, y, a, b, p = sets[(5 + i) % N]; p(m, y, a, b)
        ev[i + 1].record(st)
    torch.cuda.synchronize()
    d = [ev[i].elapsed_time(ev[i + 1]) * 1e3 for i in range(K)]
    print("rep", rep, "total/K %.2f us" % (ev[0].elapsed_time(ev[K]) * 1e3 / K), " per step:", " ".join("%.1f" % x for x in d))
    # without per-step events
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for i in range(20):
        m, y, a, b, p = sets[(5 + i) % N]; p(m, y, a, b)
    e1.record(st)
    torch.cuda.synchronize()
    print("   20 steps, no inner events: %.2f us/step" % (e0.elapsed_time(e1) * 1e3 / 20))
