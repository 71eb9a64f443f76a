import os, sys, math
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "face-gan-tts_b200"))
import torch
import face_gan_tts_b200 as fgt
torch.set_printoptions(linewidth=200, precision=2, sci_mode=False)
B, F, Tx, Ty = 1, 80, 64, 64
def cross(mu, y):
    o = fgt.log_prior(mu.cuda(), y.cuda(), impl="tcgen05").cpu().double()
    ysq = -0.5 * (y.double() ** 2).sum(1)            # [B,Ty]
    musq = -0.5 * (mu.double() ** 2).sum(1)          # [B,Tx]
    c = -0.5 * math.log(2 * math.pi) * F
    return o - ysq[:, None, :] - musq[:, :, None] - c
# test 1: mu = 1, y[f0,t] = t+1
for f0 in (0, 1, 8, 9):
    mu = torch.ones(B, F, Tx); y = torch.zeros(B, F, Ty); y[0, f0, :] = torch.arange(1, Ty + 1).float()
    c = cross(mu, y)
    print(f"T1 f0={f0}: expect row = 1..Ty for all x\n", c[0, :3, :40])
# test 2: y = 1 (all), mu[f0,x] = x+1
for f0 in (0, 1, 8):
    mu = torch.zeros(B, F, Tx); mu[0, f0, :] = torch.arange(1, Tx + 1).float(); y = torch.ones(B, F, Ty)
    c = cross(mu, y)
    print(f"T2 f0={f0}: expect col = x+1 for all t\n", c[0, :40, :3].T)
# test 3: mu[f,x] = f+1, y[f,t] = delta(f==f0)
for f0 in (0, 1, 2, 7, 8, 15, 79):
    mu = torch.arange(1, F + 1).float()[None, :, None].expand(B, F, Tx).contiguous(); y = torch.zeros(B, F, Ty); y[0, f0, :] = 1
    c = cross(mu, y)
    print(f"T3 f0={f0}: expect {f0+1} everywhere: got", c[0, 0, :4].tolist(), c[0, 5, 10:12].tolist())
