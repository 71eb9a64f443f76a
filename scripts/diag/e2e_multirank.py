"""Why does the end-to-end step (packed host batch -> H2D -> unpack -> fused kernel -> D2H of durations / frame_token) take
longer per rank when 4 or 8 ranks run on one node, although the node's PCIe delivers 54 GB/s to every GPU at once
(scripts/diag/h2d_concurrency.py)?  torchrun --nproc-per-node N scripts/diag/e2e_multirank.py; every rank prints its own
per-step time for a series of variants of the bench's e2e loop (gloo barrier in front of each)."""
import os, sys, time
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, os.path.join(ROOT, "face-gan-tts_b200"))
import torch
import torch.distributed as dist

rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); lr = int(os.environ.get("LOCAL_RANK", "0"))
os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29533")
ALL_CORES = sorted(os.sched_getaffinity(0))


def bind(mode):
    if mode == "slice":
        per = max(1, len(ALL_CORES) // world)
        os.sched_setaffinity(0, ALL_CORES[lr * per:(lr + 1) * per] or ALL_CORES)
    else:
        os.sched_setaffinity(0, ALL_CORES)


bind(os.environ.get("E2E_BIND", "slice"))
dist.init_process_group("gloo", rank=rank, world_size=world)
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
import face_gan_tts_b200 as fgt
from face_gan_tts_b200 import synthetic

B, F, TX, TY = 32, 80, 190, 1000
NH, NSETS = 3, 6
host_in = [[t.pin_memory() for t in synthetic.lrs2_batch(B, F, TX, TY, seed=4321 + 100 * rank + k)] for k in range(NH)]
host_packed = [fgt.pack_batch(*h) for h in host_in]
staging = [torch.empty((max(p.numel() for p in host_packed),), dtype=torch.uint8, device=dev) for _ in range(2)]
sets = [dict(mu=torch.empty((B, F, TX), device=dev), y=torch.empty((B, F, TY), device=dev),
             tx=torch.empty((B,), dtype=torch.int32, device=dev), ty=torch.empty((B,), dtype=torch.int32, device=dev)) for _ in range(NSETS)]
dur_h = [torch.empty((B, TX), dtype=torch.int32).pin_memory() for _ in range(2)]
ft_h = [torch.empty((B, TY), dtype=torch.int32).pin_memory() for _ in range(2)]
stream = torch.cuda.current_stream(dev)
copy_stream = torch.cuda.Stream(dev)
plans = [fgt.AlignmentPlan(B, F, TX, TY, device=dev, dense_path=False) for _ in range(2)]
for i in range(NSETS):      # valid contents everywhere
    fgt.upload_packed_batch(host_packed[i % NH], B, F, TX, TY, device=dev, out=(sets[i]["mu"], sets[i]["y"], sets[i]["tx"], sets[i]["ty"]), staging=staging[0])
torch.cuda.synchronize(dev)


def run_e2e(nsteps, h2d=True, compute=True, d2h=True, blocking_events=False, stamps=None):
    h2d_done = [torch.cuda.Event() for _ in range(nsteps)]
    res_done = [torch.cuda.Event(blocking=blocking_events) for _ in range(2)]

    def enqueue_h2d(i):
        d = sets[i % NSETS]
        with torch.cuda.stream(copy_stream):
            if h2d:
                fgt.upload_packed_batch(host_packed[i % NH], B, F, TX, TY, device=dev, out=(d["mu"], d["y"], d["tx"], d["ty"]), staging=staging[i & 1])
            h2d_done[i].record(copy_stream)

    enqueue_h2d(0)
    for i in range(nsteps):
        if stamps is not None:
            stamps.append(time.perf_counter())
        if i + 1 < nsteps:
            if i >= 1:
                copy_stream.wait_event(res_done[(i - 1) & 1])
            enqueue_h2d(i + 1)
        d = sets[i % NSETS]
        stream.wait_event(h2d_done[i])
        if compute:
            res = plans[i & 1](d["mu"], d["y"], d["tx"], d["ty"])
            if d2h:
                dur_h[i & 1].copy_(res.durations, non_blocking=True)
                ft_h[i & 1].copy_(res.frame_token, non_blocking=True)
        res_done[i & 1].record(stream)
        if i >= 1:
            res_done[(i - 1) & 1].synchronize()
    res_done[(nsteps - 1) & 1].synchronize()


def leg(name, K, **kw):
    run_e2e(4, **{k: v for k, v in kw.items() if k != "stamps"})
    torch.cuda.synchronize(dev)
    dist.barrier()
    stamps = []
    t0 = time.perf_counter()
    run_e2e(K, stamps=stamps, **kw)
    torch.cuda.synchronize(dev)
    us = (time.perf_counter() - t0) / K * 1e6
    gaps = sorted((b - a) * 1e6 for a, b in zip(stamps, stamps[1:]))
    t = torch.tensor([us], dtype=torch.float64)
    allt = [torch.zeros(1, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(allt, t)
    if rank == 0:
        print(f"{name:58s} K={K:4d}  per-rank us/step: " + " ".join(f"{float(x):6.1f}" for x in allt) +
              f"   rank0 iteration gaps us: median {gaps[len(gaps)//2]:.0f} p90 {gaps[int(len(gaps)*0.9)]:.0f} max {gaps[-1]:.0f}", flush=True)


leg("e2e packed (bench loop), cores sliced per rank", 20)
leg("e2e packed (bench loop), cores sliced per rank", 200)
leg("H2D + unpack only", 200, compute=False)
leg("compute + D2H only (no H2D)", 200, h2d=False)
leg("compute only (no H2D, no D2H)", 200, h2d=False, d2h=False)
leg("e2e packed, blocking-sync events", 200, blocking_events=True)
bind("all")
leg("e2e packed, no core binding", 200)
bind("slice")
# one rank at a time, the others idle: is it the concurrency at all?
for r in range(world):
    dist.barrier()
    if r == rank:
        run_e2e(4); torch.cuda.synchronize(dev)
        t0 = time.perf_counter(); run_e2e(200); torch.cuda.synchronize(dev)
        print(f"rank {r} alone: {(time.perf_counter() - t0) / 200 * 1e6:.1f} us/step", flush=True)
dist.barrier()
# the NCCL process group beside it (its proxy / watchdog threads)
if world > 1:
    pg = dist.new_group(backend="nccl")
    x = torch.ones(1, device=dev); dist.all_reduce(x, group=pg); torch.cuda.synchronize(dev)
    leg("e2e packed, NCCL group alive", 200)
dist.barrier()
dist.destroy_process_group()
