#!/usr/bin/env python
"""Build recipe for the REFERENCE's own MAS implementation (test infrastructure).

TEST INFRASTRUCTURE ONLY -- nothing under oracle/ may be imported by the
product package (face-gan-tts_b200/).  Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs use it, as the checker.

What it does: cythonizes the reference's `model/monotonic_align/core.pyx`
*where it lies* under /root/reference (nothing is copied into the repo) and
compiles the generated C into `oracle/_ref/` (git-ignored, NOT
gpurun-ignored, so the built .so travels to the GPU box):

  oracle/_ref/asis/core.*.so   flags exactly as the reference's setup.py gives
                               (distutils defaults, i.e. sysconfig CFLAGS, no
                               -fopenmp -> `prange` runs serially; reference
                               model/monotonic_align/setup.py:1-11)
  oracle/_ref/omp/core.*.so    same source, `-O3 -fopenmp` (the "generous"
                               all-cores CPU baseline)

Run:  python oracle/build_ref.py           (needs /root/reference; no-op with a
                                            message when it is absent, e.g. on
                                            the GPU box, where the prebuilt .so
                                            files are used)
"""
import os
import shutil
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.environ.get("FACE_GAN_TTS_REFERENCE", "/root/reference")
REF_PYX = os.path.join(REF_ROOT, "model", "monotonic_align", "core.pyx")
OUT = os.path.join(HERE, "_ref")


def _ext_suffix():
    return sysconfig.get_config_var("EXT_SUFFIX")


def built_paths():
    return {
        "asis": os.path.join(OUT, "asis", "core" + _ext_suffix()),
        "omp": os.path.join(OUT, "omp", "core" + _ext_suffix()),
    }


def build(force=False, verbose=True):
    """Returns dict variant -> path of the built extension (or {} if the
    reference tree is not present and nothing was prebuilt)."""
    paths = built_paths()
    if not os.path.exists(REF_PYX):
        have = {k: v for k, v in paths.items() if os.path.exists(v)}
        if verbose:
            print(f"[oracle/build_ref] {REF_PYX} absent; prebuilt: {sorted(have)}")
        return have
    if not force and all(os.path.exists(p) for p in paths.values()):
        newest_src = os.path.getmtime(REF_PYX)
        if all(os.path.getmtime(p) >= newest_src for p in paths.values()):
            return paths

    import numpy
    from Cython.Build import cythonize  # noqa: F401  (presence check)

    gen_dir = os.path.join(OUT, "gen")
    os.makedirs(gen_dir, exist_ok=True)
    c_file = os.path.join(gen_dir, "core.c")
    # cython -o writes ONLY the generated C to our directory; the .pyx is read in place.
    subprocess.check_call(
        [sys.executable, "-m", "cython", "-3", "-o", c_file, REF_PYX],
        stdout=subprocess.DEVNULL if not verbose else None,
        stderr=subprocess.DEVNULL,
    )
    cc = (sysconfig.get_config_var("CC") or "gcc").split()
    if shutil.which(cc[0]) is None:
        cc = ["gcc"]
    base_cflags = (sysconfig.get_config_var("CFLAGS") or "-O2").split()
    incs = ["-I" + sysconfig.get_paths()["include"], "-I" + numpy.get_include()]
    variants = {
        # reference setup.py passes no extra flags: distutils' defaults only.
        "asis": [],
        "omp": ["-O3", "-fopenmp"],
    }
    for name, extra in variants.items():
        os.makedirs(os.path.dirname(paths[name]), exist_ok=True)
        cmd = cc + base_cflags + ["-fPIC", "-shared", "-w"] + extra + incs + [c_file, "-o", paths[name]]
        if verbose:
            print("[oracle/build_ref]", " ".join(cmd))
        subprocess.check_call(cmd)
    return paths


def load(variant="asis"):
    """Import the compiled reference module; returns module with
    `maximum_path_c(paths, values, t_xs, t_ys, max_neg_val=-1e9)`
    (reference core.pyx:40).  None if not built."""
    import importlib.util

    p = built_paths()[variant]
    if not os.path.exists(p):
        return None
    spec = importlib.util.spec_from_file_location("core", p)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    out = build(force="--force" in sys.argv)
    for k, v in out.items():
        print(k, v)
