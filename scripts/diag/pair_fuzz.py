import os, sys, random
sys.path.insert(0, "/root/repo/face-gan-tts_b200"); sys.path.insert(0, "/root/repo")
import torch, numpy as np
import face_gan_tts_b200 as fgt
from face_gan_tts_b200 import _lib, synthetic
random.seed(7)
DEV="cuda:0"
bad=0; n=0
for it in range(60):
    B=random.choice([1,2,3,5,9,33,74])
    F=random.choice([64,80])
    Tx=random.choice([129,130,160,191,192,193,224,255,256])
    Ty=random.choice([Tx+ (4-Tx%4)%4, 260, 288, 320, 516, 1000, 1028, 1404])
    if Ty < Tx: Ty = ((Tx+3)//4)*4
    mu,y,tx,ty = synthetic.lrs2_batch(B=B,F=F,Tx=Tx,Ty=Ty,seed=100+it,tx_lo=1,ty_lo=min(Ty,max(4,Tx//2)))
    # force some edge lengths
    tx[0]=Tx; ty[0]=Ty
    if B>1: tx[1]=min(129,Tx); ty[1]=max(int(ty[1]),129)
    if B>2: tx[2]=128; ty[2]=max(int(ty[2]),128)
    ty=torch.minimum(ty, torch.tensor(Ty)); 
    outs=[]
    for mode in (0,2):
        prev=_lib.set_option("fused_pair",mode)
        try:
            r=fgt.log_prior_maximum_path(mu.to(DEV),y.to(DEV),tx,ty,path_dtype=torch.int32)
            torch.cuda.synchronize()
            outs.append(r)
        finally:
            _lib.set_option("fused_pair",prev)
    a,b=outs
    ok = torch.equal(a.durations,b.durations) and torch.equal(a.frame_token,b.frame_token) and torch.equal(a.path,b.path) and torch.equal(a.status,b.status)
    n+=1
    if not ok:
        bad+=1
        print("MISMATCH", B,F,Tx,Ty, tx.tolist()[:4], ty.tolist()[:4])
print("cases", n, "mismatches", bad)
