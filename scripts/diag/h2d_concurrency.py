"""What limits host->device copies when several GPUs of one node pull their batches at once?
One process, one stream per GPU; 8 MB copies (the packed LRS2 batch) from pinned host memory.
Prints the topology the box exposes, solo / pairwise / all-GPU bandwidth, and the same with the pinned buffer
(a) allocated write-combined, (b) backed by transparent huge pages, (c) first-touched on each NUMA node."""
import ctypes, mmap, os, subprocess, sys, time
import torch

N = torch.cuda.device_count()
NBYTES = 8 * 1024 * 1024
REPS = 60
rt = ctypes.CDLL("libcudart.so.12")
rt.cudaHostAlloc.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_size_t, ctypes.c_uint]
rt.cudaHostRegister.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_uint]
rt.cudaMemcpyAsync.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p]


def sh(cmd):
    try:
        return subprocess.run(cmd, shell=True, capture_output=True, text=True, timeout=20).stdout.strip()
    except Exception as ex:
        return f"<{ex!r}>"


print("== topology")
print(sh("nvidia-smi topo -m"))
print("cpus:", os.cpu_count(), "affinity:", len(os.sched_getaffinity(0)))
print("numa nodes:", sh("ls -d /sys/devices/system/node/node* 2>/dev/null | tr '\\n' ' '"))
print(sh("for n in /sys/devices/system/node/node*; do echo $n cpus=$(cat $n/cpulist) $(grep MemTotal $n/meminfo); done"))
print("THP:", sh("cat /sys/kernel/mm/transparent_hugepage/enabled"), "| hugepages:", sh("grep -i huge /proc/meminfo | tr '\\n' ';'"))
for i in range(N):
    bus = torch.cuda.get_device_properties(i).pci_bus_id if hasattr(torch.cuda.get_device_properties(i), "pci_bus_id") else None
    print("gpu", i, sh(f"nvidia-smi -i {i} --query-gpu=pci.bus_id,pcie.link.gen.current,pcie.link.width.current --format=csv,noheader"),
          "numa_node:", sh(f"cat /sys/bus/pci/devices/$(nvidia-smi -i {i} --query-gpu=pci.bus_id --format=csv,noheader | cut -c5- | tr A-Z a-z)/numa_node"))

devs = [torch.device("cuda", i) for i in range(N)]
dst = [torch.empty(NBYTES, dtype=torch.uint8, device=d) for d in devs]
streams = [torch.cuda.Stream(d) for d in devs]


def host_buffers(kind, node_cpus=None):
    """N pinned host buffers of NBYTES; returns list of raw pointers (leaked: diagnostics)."""
    out = []
    for i in range(N):
        if node_cpus is not None:
            os.sched_setaffinity(0, node_cpus[i % len(node_cpus)])
        p = ctypes.c_void_p()
        if kind == "default":
            assert rt.cudaHostAlloc(ctypes.byref(p), NBYTES, 1) == 0          # portable
        elif kind == "wc":
            assert rt.cudaHostAlloc(ctypes.byref(p), NBYTES, 1 | 4) == 0      # portable | write-combined
        elif kind == "thp":
            m = mmap.mmap(-1, NBYTES + (2 << 20), flags=mmap.MAP_PRIVATE | mmap.MAP_ANONYMOUS)
            addr = ctypes.addressof(ctypes.c_char.from_buffer(m))
            al = (addr + (2 << 20) - 1) & ~((2 << 20) - 1)
            try:
                m.madvise(mmap.MADV_HUGEPAGE)
            except Exception as ex:
                print("madvise failed", ex)
            ctypes.memset(al, 1, NBYTES)                                       # first touch
            assert rt.cudaHostRegister(al, NBYTES, 1) == 0
            p = ctypes.c_void_p(al)
            host_buffers.keep.append(m)
        if kind != "thp":
            ctypes.memset(p.value, 1, NBYTES)
        out.append(p.value)
    if node_cpus is not None:
        os.sched_setaffinity(0, set(range(os.cpu_count())))
    return out


host_buffers.keep = []


def run(ptrs, active):
    ev = {}
    for i in active:
        torch.cuda.set_device(i)
        for _ in range(5):
            rt.cudaMemcpyAsync(dst[i].data_ptr(), ptrs[i], NBYTES, 1, streams[i].cuda_stream)
    for i in active:
        torch.cuda.synchronize(i)
    for i in active:
        ev[i] = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
    for i in active:
        torch.cuda.set_device(i)
        ev[i][0].record(streams[i])
    for _ in range(REPS):
        for i in active:
            torch.cuda.set_device(i)
            rt.cudaMemcpyAsync(dst[i].data_ptr(), ptrs[i], NBYTES, 1, streams[i].cuda_stream)
    for i in active:
        torch.cuda.set_device(i)
        ev[i][1].record(streams[i])
    for i in active:
        torch.cuda.synchronize(i)
    bw = [NBYTES * REPS / (ev[i][0].elapsed_time(ev[i][1]) * 1e-3) / 1e9 for i in active]
    return bw


def report(tag, ptrs):
    solo = [run(ptrs, [i])[0] for i in range(N)]
    print(f"-- {tag}: solo GB/s", " ".join(f"{b:.1f}" for b in solo))
    if N >= 2:
        for j in range(1, min(N, 8)):
            bw = run(ptrs, [0, j])
            print(f"   pair (0,{j}):", " ".join(f"{b:.1f}" for b in bw), f"sum {sum(bw):.1f}")
    if N >= 4:
        bw = run(ptrs, list(range(4)))
        print("   gpus 0-3:", " ".join(f"{b:.1f}" for b in bw), f"sum {sum(bw):.1f}")
    if N >= 8:
        bw = run(ptrs, list(range(4, 8)))
        print("   gpus 4-7:", " ".join(f"{b:.1f}" for b in bw), f"sum {sum(bw):.1f}")
    bw = run(ptrs, list(range(N)))
    print(f"   all {N}:", " ".join(f"{b:.1f}" for b in bw), f"sum {sum(bw):.1f}")


report("cudaHostAlloc", host_buffers("default"))
report("cudaHostAlloc write-combined", host_buffers("wc"))
report("mmap + MADV_HUGEPAGE + cudaHostRegister", host_buffers("thp"))
nodes = []
for n in sorted(os.listdir("/sys/devices/system/node")) if os.path.isdir("/sys/devices/system/node") else []:
    if n.startswith("node") and n[4:].isdigit():
        cl = open(f"/sys/devices/system/node/{n}/cpulist").read().strip()
        cpus = set()
        for part in cl.split(","):
            if "-" in part:
                a, b = part.split("-"); cpus |= set(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        if cpus:
            nodes.append(cpus)
if len(nodes) > 1:
    for k, cp in enumerate(nodes):
        report(f"cudaHostAlloc, first touch on NUMA node {k}", host_buffers("default", [cp]))
    report("cudaHostAlloc, buffer i on node i % nodes", host_buffers("default", nodes))
