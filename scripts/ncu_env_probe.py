#!/usr/bin/env python
"""Which environment variables does a profiler set in the profiled process?  (abi.cu keys the serial pipeline on them.)"""
import os
import torch
keys = sorted(k for k in os.environ if any(s in k for s in ("NV", "CUDA", "INJECT", "NSIGHT", "SANITIZER")))
print("ENVKEYS", keys)
print("detect", any(k in os.environ for k in ("CUDA_INJECTION64_PATH", "NV_COMPUTE_PROFILER_PERFWORKS_DIR", "NV_SANITIZER_INJECTION_PORT_BASE")))
x = torch.ones(1024, device="cuda") + 1
torch.cuda.synchronize()
print("ok", float(x.sum()))
