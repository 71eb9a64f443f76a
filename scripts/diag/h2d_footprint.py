"""Does the host-side FOOTPRINT of the pinned source buffers limit concurrent H2D?  Every GPU of the node copies 8 MB
batches, rotating over k pinned buffers (k x 8 MB per GPU): k = 1 re-reads one buffer (served from the CPU's last-level
cache), larger k has to come from host DRAM.  One process, one stream per GPU."""
import ctypes, os, sys
import torch

N = torch.cuda.device_count()
NBYTES = 8 * 1024 * 1024
REPS = 96
rt = ctypes.CDLL("libcudart.so.12")
rt.cudaHostAlloc.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_size_t, ctypes.c_uint]
rt.cudaMemcpyAsync.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p]
KMAX = 12
devs = [torch.device("cuda", i) for i in range(N)]
dst = [torch.empty(NBYTES, dtype=torch.uint8, device=d) for d in devs]
streams = [torch.cuda.Stream(d) for d in devs]
bufs = []
for i in range(N):
    row = []
    for k in range(KMAX):
        p = ctypes.c_void_p()
        assert rt.cudaHostAlloc(ctypes.byref(p), NBYTES, 1) == 0
        ctypes.memset(p.value, k + 1, NBYTES)
        row.append(p.value)
    bufs.append(row)


def run(active, k, touch=False):
    for i in active:
        torch.cuda.synchronize(i)
    ev = {i: (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for i in active}
    for i in active:
        torch.cuda.set_device(i)
        ev[i][0].record(streams[i])
    for r in range(REPS):
        for i in active:
            torch.cuda.set_device(i)
            if touch:      # the host (re)writes the batch just before it is copied, like a collate does
                ctypes.memset(bufs[i][r % k], r & 0xff, NBYTES)
            rt.cudaMemcpyAsync(dst[i].data_ptr(), bufs[i][r % k], NBYTES, 1, streams[i].cuda_stream)
    for i in active:
        torch.cuda.set_device(i)
        ev[i][1].record(streams[i])
    for i in active:
        torch.cuda.synchronize(i)
    return [NBYTES * REPS / (ev[i][0].elapsed_time(ev[i][1]) * 1e-3) / 1e9 for i in active]


print("lscpu cache:", os.popen("lscpu | grep -i 'L3\\|L2\\|Model name' | tr -s ' ' | tr '\\n' ';'").read())
for active in ([0], list(range(min(N, 2))), list(range(N))):
    for k in (1, 2, 3, 6, 12):
        run(active, k)
        bw = run(active, k)
        print(f"gpus {len(active)}  k={k:2d}  footprint {len(active) * k * 8:4d} MB   per-GPU GB/s " + " ".join(f"{b:5.1f}" for b in bw) + f"   sum {sum(bw):6.1f}")
