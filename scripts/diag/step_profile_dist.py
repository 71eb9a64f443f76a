"""Where do the extra ~50-70 us of a 20-step window go at N > 1 (bench: 42.1 us/step at N = 2 in 20 steps, 38.7 in 200)?
torchrun --nproc-per-node 2 scripts/diag/step_profile_dist.py : per-step device time of the bench's multi-GPU step on
rank 0 (events created before the window), with / without the duration put, behind a NCCL barrier or a plain sync."""
import os, sys, time
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, os.path.join(ROOT, "face-gan-tts_b200"))
import torch
import torch.distributed as dist

rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); lr = int(os.environ.get("LOCAL_RANK", "0"))
os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29535")
os.environ.setdefault("NCCL_MAX_NCHANNELS", "1"); os.environ.setdefault("NCCL_MIN_NCHANNELS", "1")
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
import face_gan_tts_b200 as fgt
from face_gan_tts_b200 import sharding, synthetic

B, F, TX, TY, NSETS, K = 32, 80, 190, 1000, 6, 20
sets = []
for s in range(NSETS):
    mu, y, tx, ty = synthetic.lrs2_batch(B, F, TX, TY, seed=1234 + 1000 * rank + s)
    sets.append((mu.to(dev), y.to(dev), tx.to(dev), ty.to(dev), fgt.AlignmentPlan(B, F, TX, TY, device=dev, dense_path=True)))
put = sharding.OneSidedDurationGather(B, TX, dev)
stream = torch.cuda.current_stream(dev)
comm = torch.cuda.Stream(dev)
step_done = [torch.cuda.Event() for _ in range(4)]
marks = [torch.cuda.Event(enable_timing=True) for _ in range(K + 2)]
last = {}


def step(i, with_put):
    m, y, a, b, p = sets[i % NSETS]
    r = p(m, y, a, b)
    if with_put:
        ev = step_done[i % 4]
        ev.record(stream)
        comm.wait_event(ev)
        put.put(r.durations, comm)


def window(name, with_put, nccl_barrier, per_step_marks):
    for i in range(5):
        step(i, with_put)
    if nccl_barrier:
        dist.barrier()
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    marks[0].record(stream)
    for i in range(K):
        step(5 + i, with_put)
        if per_step_marks:
            marks[i + 1].record(stream)
    host_us = (time.perf_counter() - t0) / K * 1e6
    if with_put:
        stream.wait_stream(comm)
    marks[K + 1].record(stream)
    torch.cuda.synchronize(dev)
    dist.barrier()
    total = marks[0].elapsed_time(marks[K + 1]) * 1e3
    if rank == 0:
        line = f"{name:64s} total {total:7.1f} us = {total / K:5.2f} us/step, host enqueue {host_us:5.1f} us/step"
        if per_step_marks:
            d = [marks[i].elapsed_time(marks[i + 1]) * 1e3 for i in range(K)]
            line += "\n      per step: " + " ".join(f"{x:.1f}" for x in d) + f" | tail after the last step {marks[K].elapsed_time(marks[K + 1]) * 1e3:.1f}"
        print(line, flush=True)


for rep in range(2):
    window("put on a side stream, NCCL barrier in front (the bench)", True, True, False)
    window("put on a side stream, NCCL barrier in front, per-step marks", True, True, True)
    window("no put, NCCL barrier in front", False, True, False)
    window("no put, NCCL barrier in front, per-step marks", False, True, True)
    window("put on a side stream, plain synchronize in front", True, False, False)
    window("no put, plain synchronize in front", False, False, False)
torch.cuda.synchronize(dev)
dist.barrier()
dist.destroy_process_group()
