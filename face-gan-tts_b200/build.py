#!/usr/bin/env python
"""Build libmas_b200.so (sm_100a only) in-tree with nvcc.

    python face-gan-tts_b200/build.py [--force] [--verbose]

Output: face-gan-tts_b200/lib/libmas_b200.so  (git-ignored; travels to the GPU box
with the gpurun snapshot).  One build target, no fallback architectures.
"""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libmas_b200.so")
# (source, object stem, extra flags)
UNITS = [
    ("abi.cu", "abi", []),
    ("mas_forward.cu", "mas_forward", []),
    ("mas_forward_inst.cu", "mas_forward_r1", ["-DMASB200_INST_R=1"]),
    ("mas_forward_inst.cu", "mas_forward_r2", ["-DMASB200_INST_R=2"]),
    ("mas_forward_inst.cu", "mas_forward_r4", ["-DMASB200_INST_R=4"]),
    ("mas_forward_inst.cu", "mas_forward_r8", ["-DMASB200_INST_R=8"]),
    ("path_ops.cu", "path_ops", []),
    ("loss_ops.cu", "loss_ops", []),
    ("upload.cu", "upload", []),
    ("log_prior_ffma.cu", "log_prior_ffma", []),
    ("log_prior_tc.cu", "log_prior_tc", []),
    ("lp_mas_fused.cu", "lp_mas_fused", []),
]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-O2",
    "--expt-relaxed-constexpr",
    "-Xptxas", "-v",
]
# diagnostics build: in-loop wait / body cycle counters of the fused kernel (scripts/fused_phases.py)
if os.environ.get("MASB200_XFLAGS"):
    FLAGS += os.environ["MASB200_XFLAGS"].split()
if os.environ.get("MASB200_HELP_PROF"):
    FLAGS.append("-DMASB200_HELP_PROF=1")
if os.environ.get("MASB200_PROF"):
    FLAGS.append("-DMASB200_FUSED_PROF=1")


def _digest():
    h = hashlib.sha256()
    names = sorted(os.listdir(CSRC)) + ["../../include/mas_b200.h"]
    for n in names:
        p = os.path.join(CSRC, n)
        if os.path.isfile(p):
            h.update(n.encode())
            h.update(open(p, "rb").read())
    h.update(" ".join(FLAGS).encode())
    return h.hexdigest()


def build(force=False, verbose=False):
    os.makedirs(LIBDIR, exist_ok=True)
    stamp = os.path.join(LIBDIR, "build.sha256")
    dig = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read().strip() == dig:
        return LIB
    from concurrent.futures import ThreadPoolExecutor

    def compile_one(unit):
        src, stem, extra = unit
        o = os.path.join(LIBDIR, stem + ".o")
        cmd = [NVCC] + FLAGS + extra + ["-c", os.path.join(CSRC, src), "-o", o]
        r = subprocess.run(cmd, capture_output=True, text=True)
        return o, f"$ {' '.join(cmd)}\n{r.stdout}{r.stderr}", r.returncode

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        results = list(ex.map(compile_one, UNITS))
    objs = [r[0] for r in results]
    logs = [r[1] for r in results]
    for (o, log, rcode), unit in zip(results, UNITS):
        if rcode != 0:
            sys.stderr.write(log)
            raise RuntimeError(f"nvcc failed on {unit[0]} ({unit[1]})")
    cmd = [NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs + ["-cudart", "static"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    logs.append(f"$ {' '.join(cmd)}\n{r.stdout}{r.stderr}")
    if r.returncode != 0:
        sys.stderr.write(logs[-1])
        raise RuntimeError("link failed")
    with open(os.path.join(LIBDIR, "build.log"), "w") as f:
        f.write("\n".join(logs))
    with open(stamp, "w") as f:
        f.write(dig)
    if verbose:
        print("\n".join(logs))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
