"""H2D from pinned memory, N PROCESSES (one per GPU) at once -- where does the per-GPU bandwidth go?
torchrun --nproc-per-node N scripts/diag/h2d_multiproc.py.  gloo barrier in front of every variant; every rank reports."""
import ctypes, os, sys, time
import torch
import torch.distributed as dist

rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); lr = int(os.environ.get("LOCAL_RANK", "0"))
os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29534")
dist.init_process_group("gloo", rank=rank, world_size=world)
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
rt = ctypes.CDLL("libcudart.so.12")
rt.cudaHostAlloc.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_size_t, ctypes.c_uint]
rt.cudaMemcpyAsync.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p]
NB = 8 * 1024 * 1024
dst = torch.empty(64 * 1024 * 1024, dtype=torch.uint8, device=dev)
src_t = [torch.empty(NB, dtype=torch.uint8).pin_memory() for _ in range(3)]
for t in src_t:
    t.fill_(3)
big_t = torch.empty(64 * 1024 * 1024, dtype=torch.uint8).pin_memory(); big_t.fill_(1)
raw = []
for k in range(3):
    p = ctypes.c_void_p(); assert rt.cudaHostAlloc(ctypes.byref(p), NB, 1) == 0; ctypes.memset(p.value, 2, NB); raw.append(p.value)
s = torch.cuda.Stream(dev)
s2 = torch.cuda.Stream(dev)


def report(name, gbs):
    t = torch.tensor([gbs], dtype=torch.float64)
    allt = [torch.zeros(1, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(allt, t)
    if rank == 0:
        print(f"{name:74s} GB/s per rank: " + " ".join(f"{float(x):5.1f}" for x in allt), flush=True)


def timed(name, fn, reps, nbytes=NB):
    fn(4); torch.cuda.synchronize(dev)
    dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(s)
    fn(reps)
    b.record(s)
    torch.cuda.synchronize(dev)
    report(name, nbytes * reps / (a.elapsed_time(b) * 1e-3) / 1e9)


def v_raw(reps):
    for r in range(reps):
        rt.cudaMemcpyAsync(dst.data_ptr(), raw[r % 3], NB, 1, s.cuda_stream)


def v_torch(reps):
    with torch.cuda.stream(s):
        for r in range(reps):
            dst[:NB].copy_(src_t[r % 3], non_blocking=True)


def v_torch_sync_each(reps):
    with torch.cuda.stream(s):
        for r in range(reps):
            dst[:NB].copy_(src_t[r % 3], non_blocking=True)
            s.synchronize()


def v_torch_event_chain(reps):
    # copy on s, then a cross-stream hop (s2 waits, records) the next copy depends on -- the bench loop's shape without kernels
    ev = [torch.cuda.Event() for _ in range(2)]
    ev2 = [torch.cuda.Event() for _ in range(2)]
    for r in range(reps):
        if r >= 2:
            s.wait_event(ev2[r & 1])
        with torch.cuda.stream(s):
            dst[:NB].copy_(src_t[r % 3], non_blocking=True)
        ev[r & 1].record(s)
        s2.wait_event(ev[r & 1])
        ev2[r & 1].record(s2)
        if r >= 1:
            ev2[(r - 1) & 1].synchronize()


def v_big(reps):
    with torch.cuda.stream(s):
        for r in range(reps):
            dst.copy_(big_t, non_blocking=True)


sys.path.insert(0, os.path.join(os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "..")), "face-gan-tts_b200"))
import face_gan_tts_b200 as fgt
from face_gan_tts_b200 import synthetic
B, F, TX, TY = 32, 80, 190, 1000
packed = [fgt.pack_batch(*[t for t in synthetic.lrs2_batch(B, F, TX, TY, seed=4321 + 100 * rank + k)]) for k in range(3)]
outs = (torch.empty((B, F, TX), device=dev), torch.empty((B, F, TY), device=dev), torch.empty((B,), dtype=torch.int32, device=dev), torch.empty((B,), dtype=torch.int32, device=dev))
stg = [torch.empty((max(p.numel() for p in packed),), dtype=torch.uint8, device=dev) for _ in range(2)]
PB = int(sum(p.numel() for p in packed) / 3)


def v_upload(reps):
    with torch.cuda.stream(s):
        for r in range(reps):
            fgt.upload_packed_batch(packed[r % 3], B, F, TX, TY, device=dev, out=outs, staging=stg[r & 1])


def v_copy_packed_only(reps):
    with torch.cuda.stream(s):
        for r in range(reps):
            stg[r & 1][:packed[r % 3].numel()].copy_(packed[r % 3], non_blocking=True)


timed("raw cudaMemcpyAsync 8 MB back to back (cudaHostAlloc portable)", v_raw, 100)
timed("torch copy_ 8 MB back to back (pin_memory)", v_torch, 100)
timed("torch copy_ 8 MB, stream synchronize after each", v_torch_sync_each, 100)
timed("torch copy_ 8 MB in an event chain with a host sync per step", v_torch_event_chain, 100)
timed("torch copy_ 64 MB back to back", v_big, 20, 64 * 1024 * 1024)
timed("packed batch: copy only, back to back", v_copy_packed_only, 100, PB)
timed("packed batch: upload_packed_batch (copy + unpack kernel) back to back", v_upload, 100, PB)
timed("raw cudaMemcpyAsync 8 MB back to back, again", v_raw, 100)
dist.barrier()
dist.destroy_process_group()
