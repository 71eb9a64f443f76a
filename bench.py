#!/usr/bin/env python
"""bench.py -- alignment-cells/s of the log-prior + MAS hot path on B200 (one JSON line).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (BASELINE.json configs[1]): fused log_prior + MAS at the LRS2 train batch shape,
B=32 utterances per GPU, n_feats=80, T_text=190, T_mel=1000, synthetic lengths
(face_gan_tts_b200.synthetic.lrs2_batch, seed 1234).  A "step" is one pass of the hot
path over one batch: mu_x, y, lengths -> dense fp32 path + durations + frame->token index.
Metric: alignment cells/s = B*T_text*T_mel / time (padded cells), whole job over all GPUs.

  value     inputs already resident in HBM; K steps back to back on one stream between two
            CUDA events; a step is ONE kernel (lp_mas_fused_kernel: tcgen05 log-prior -> shared-memory
            ring -> MAS -> backtrack -> dense path; at this batch size a 2-CTA cluster per utterance); the steps rotate over NSETS independent buffer
            sets whose footprint exceeds L2, so no step finds its inputs in cache.
  e2e       the same step through the public API with HOST buffers: every step moves the batch from
            pinned host memory (packed ragged form, one copy-engine transfer + device unpack), runs
            the fused call, and reads durations + frame->token index back to pinned host memory
            (`e2e`); `e2e_padded_copy` = padded tensors with plain copies, `e2e_dense_path` = also
            the whole dense path back.
  roofline  the step's kernel: algorithmic bytes / its CUDA-event time against MEASURED_PEAKS.json
            hbm_gbs, on padded and on valid bytes, `traffic` from the committed ncu capture.
  sweep     the throughput-regime workloads (configs[3] B=64 512x4096, configs[4] B=1024), each with
            its own roofline.       path_agreement   % of frames agreeing with torch fp32 -> core.pyx.
  cpu_baseline  the reference's own implementation (oracle/_ref: core.pyx as shipped, serial)
            + the torch log-prior expression on this box's host cores, same batch.
  --impl reference  times that CPU path as its own arm (rank 0 only under torchrun).

Multi-GPU (torchrun, one rank per GPU): utterances are independent, so each rank aligns its
own 32-utterance shard (weak scaling, no data-path collective); the per-token durations are
returned to every rank every step (loss bookkeeping) by a ONE-SIDED gather: every rank's
[world*B, Tx] buffer is symmetric memory mapped into all peers, and a rank's durations are
stored into all of them over NVLink -- by a small copy kernel on a side stream (default: off the
step's critical path) or by the fused kernel's own output stage (--put-in-kernel) -- overlapped
with the next step and completed inside the timed region.  No collective kernel, no
torch.distributed call per step (an NCCL all-gather per step costs more host time than the
step takes on the GPU; --nccl-gather keeps that path).  `strong_scaling_configs4` adds BASELINE
configs[4]: a fixed total batch (64..1024) split over the ranks.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "face-gan-tts_b200"))

import numpy as np  # noqa: E402
import torch  # noqa: E402

B, F, TX, TY = 32, 80, 190, 1000
METRIC = "MAS alignment cells/s (B*T_text*T_mel)"
UNIT = "cells/s"
CELLS = B * TX * TY


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region.

    The timed region of the headline workload is a few milliseconds long, so the samples come from NVML called in-process
    (`pynvml`, one sample per millisecond from a thread that holds no Python lock while it waits); `nvidia-smi -lms` (the
    profiling recipe's clocks line, >= 100 ms per sample) is only the fallback.  `hold()` keeps the same step running untimed
    until at least two samples under load exist, for the case that even the NVML thread saw none in the window."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    REASONS = ((0x8, "hw_slowdown"), (0x40, "hw_thermal_slowdown"), (0x20, "sw_thermal_slowdown"), (0x4, "sw_power_cap"))

    def __init__(self, index=0, uuid=None):
        self.index, self.uuid = index, uuid
        self.rows = []            # (sm_mhz, max_mhz, [reasons])
        self.proc = self.thread = self.nvml = None
        self.source = None
        self._stop = threading.Event()

    # -- NVML, in-process
    def _nvml_open(self):
        import pynvml
        pynvml.nvmlInit()
        h = None
        if self.uuid:
            try:
                h = pynvml.nvmlDeviceGetHandleByUUID(self.uuid)
            except Exception:
                try:
                    h = pynvml.nvmlDeviceGetHandleByUUID(self.uuid.encode())
                except Exception:
                    h = None
        if h is None:
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
        get_reasons = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
            getattr(pynvml, "nvmlDeviceGetCurrentClocksThrottleReasons")
        mx = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))

        def sample():
            sm = float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
            bits = int(get_reasons(h))
            return sm, mx, [n for b, n in self.REASONS if bits & b]

        sample()                  # fails here, not in the thread, if the device does not answer
        return sample

    def _nvml_loop(self, sample):
        while not self._stop.is_set():
            try:
                self.rows.append(sample())
            except Exception:
                pass
            self._stop.wait(0.001)

    # -- nvidia-smi, fallback
    def _smi_loop(self):
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.proc.stdout:
            c = [x.strip() for x in line.strip().split(",")]
            if len(c) < 7:
                continue
            try:
                self.rows.append((float(c[0]), float(c[1]),
                                  [n for n, v in zip(names, c[3:7]) if v.lower().startswith("active")]))
            except ValueError:
                continue

    def start(self):
        try:
            sample = self._nvml_open()
            self.source = "nvml"
            self.thread = threading.Thread(target=self._nvml_loop, args=(sample,), daemon=True)
            self.thread.start()
            return
        except Exception:
            pass
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.source = "nvidia-smi -lms 100"
            self.thread = threading.Thread(target=self._smi_loop, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def mark(self):
        return len(self.rows)

    def hold(self, since, run_step, sync, want=2, limit_s=1.5):
        """Run `run_step(i)` (untimed, the timed region's own step) until `want` samples newer than `since` exist."""
        if self.thread is None:
            return 0
        t0, i = time.perf_counter(), 0
        while len(self.rows) - since < want and time.perf_counter() - t0 < limit_s:
            for _ in range(8):
                run_step(i)
                i += 1
            sync()
        return i

    def stop(self, since=0, held_steps=0):
        if self.thread is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "samples": 0, "reasons": ["no NVML and no nvidia-smi on this host"]}
        self._stop.set()
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        self.thread.join(timeout=2)
        rows = self.rows[since:] or self.rows
        sm = [r[0] for r in rows]
        mx = [r[1] for r in rows]
        reasons = sorted({n for r in rows for n in r[2]})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": reasons, "source": self.source,
                "window": "timed region" + (f" + {held_steps} untimed steps of the same workload held for the sampler"
                                            if held_steps else "")}


# ----------------------------------------------------------------------------- reference (CPU) arm
def reference_step_fn(variant="asis"):
    """The reference's CPU implementation of the path: torch log-prior (face_tts.py:165-171, all torch
    threads) + its compiled core.pyx maximum_path_c as shipped (serial) through the wrapper's numpy steps."""
    import oracle
    from face_gan_tts_b200 import synthetic

    # torchrun exports OMP_NUM_THREADS=1: give the CPU arm every host thread it can use
    try:
        torch.set_num_threads(max(1, len(os.sched_getaffinity(0))))
    except Exception:
        torch.set_num_threads(max(1, os.cpu_count() or 1))
    core = oracle.reference_core(variant)
    kind = "reference"
    if core is None:
        core, kind = oracle, "port"
    mu_x, y, t_x, t_y = synthetic.lrs2_batch(B, F, TX, TY, seed=1234)
    mask = synthetic.prefix_mask(t_x, t_y, TX, TY)
    t_x_np, t_y_np = t_x.numpy(), t_y.numpy()

    def step():
        log_prior = oracle.log_prior_reference(mu_x, y)
        value = (log_prior * mask).numpy().astype(np.float32)          # __init__.py:13,16
        path = np.zeros_like(value).astype(np.int32)                     # :17
        core.maximum_path_c(path, value, t_x_np, t_y_np)                 # :22
        return torch.from_numpy(path).to(dtype=log_prior.dtype)         # :23

    cores = torch.get_num_threads()
    sample = (f"full configs[1] batch per step (B={B}, F={F}, Tx={TX}, Ty={TY}): torch CPU log-prior on "
              f"{cores} threads + reference core.pyx maximum_path_c serial (as shipped, no OpenMP)")
    return step, kind, cores, sample


def time_cpu(step, steps, warmup):
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    return (time.perf_counter() - t0) / steps


def workload_config():
    """The keys both arms of the bench print under `config` (the GPU arm adds what only it has)."""
    return {"workload": "configs[1]: log_prior+MAS, LRS2 train batch shape", "B_per_gpu": B, "n_feats": F,
            "T_text": TX, "T_mel": TY}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    step, kind, cores, sample = reference_step_fn()
    # exactly the K steps / W warm-up steps asked for (a step is one full configs[1] batch, ~20-50 ms of CPU work); the
    # caps only keep an absurd request within minutes
    steps, warmup = max(1, min(args.steps, 1000)), max(0, min(args.warmup, 100))
    sec = time_cpu(step, steps, warmup)
    v = CELLS / sec
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(),
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def time_compute_loss_block(dev, out_size=128):
    """ms per forward+backward of the alignment block at the bench shape: this library vs the reference's way."""
    import random

    import oracle
    from face_gan_tts_b200 import losses, synthetic

    mu_x, y, t_x, t_y = synthetic.lrs2_batch(B, F, TX, TY, seed=1234)
    g = torch.Generator().manual_seed(7)
    x_mask = (torch.arange(TX)[None, :] < t_x[:, None]).float().unsqueeze(1)
    logw = torch.randn(B, 1, TX, generator=g) * x_mask
    mu_d, y_d, lw_d, xm_d = mu_x.to(dev), y.to(dev), logw.to(dev), x_mask.to(dev)
    off = losses.draw_crop_offsets(t_y.tolist(), out_size, random.Random(5))

    def ours():
        mu = mu_d.detach().requires_grad_(True)
        lw = lw_d.detach().requires_grad_(True)
        o = losses.alignment_losses(mu, lw, t_x, y_d, t_y, out_size=out_size, out_offset=off)
        (o.dur_loss + o.prior_loss + o.mu_y.sum()).backward()
        return o.prior_loss.detach()

    core = oracle.reference_core("asis")
    ty_l, tx_l = t_y.long().to(dev), t_x.long().to(dev)
    off_t = torch.as_tensor(off).long()

    def ref():
        mu = mu_d.detach().requires_grad_(True)
        lw = lw_d.detach().requires_grad_(True)
        r = oracle.compute_loss_block(mu, lw, xm_d, y_d, ty_l, tx_l, F, out_size=out_size, out_offset=off_t,
                                      maximum_path_fn=lambda v, m: oracle.maximum_path(v, m, core=core))
        (r["dur_loss"] + r["prior_loss"] + r["mu_y"].sum()).backward()
        return r["prior_loss"].detach()

    def wall(fn, n):
        fn(); torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        for _ in range(n):
            v = fn()
        float(v)                                   # the loss is read back, as a training step logs it
        return (time.perf_counter() - t0) / n * 1e3

    ms_ours = wall(ours, 50)
    out = {"shape": f"B={B} F={F} Tx={TX} Ty={TY} out_size={out_size}", "this_library_ms": ms_ours,
           "what": "masks -> log-prior -> MAS -> duration loss -> crop -> mu_y -> prior loss, forward + backward"}
    if core is not None:
        ms_ref = wall(ref, 3)
        out.update({"reference_formulation_ms": ms_ref, "speedup": ms_ref / ms_ours})
        a, b_ = float(ours()), float(ref())
        out["prior_loss_rel_diff"] = abs(a - b_) / abs(b_)
    return out


# ----------------------------------------------------------------------------- CUDA arm
def bind_rank_to_cores(local_rank, local_world):
    """Give every rank of the node its own slice of the host cores BEFORE it allocates pinned memory / starts copy
    threads: 8 ranks x pinned H2D on one host otherwise share (and migrate between) the same cores."""
    try:
        cores = sorted(os.sched_getaffinity(0))
        per = max(1, len(cores) // max(1, local_world))
        mine = cores[local_rank * per:(local_rank + 1) * per] or cores
        os.sched_setaffinity(0, mine)
        torch.set_num_threads(max(1, len(mine)))
        return len(mine)
    except Exception:
        return None


def cuda_time(stream, dev, fn, reps, warm=3):
    """ms per call of fn(i): CUDA events on the launching stream, synchronised on both sides."""
    for i in range(warm):
        fn(i)
    torch.cuda.synchronize(dev)
    a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    for i in range(reps):
        fn(warm + i)
    b_.record(stream)
    torch.cuda.synchronize(dev)
    return a.elapsed_time(b_) / reps


def ncu_dram_bytes(name):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from a committed `ncu --set full` summary
    (profiles/<name>.json, written by scripts/tools/ncu_summary.py); None if it is not there."""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", name)))
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        tot = 0.0
        for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            v, u = d[k]
            tot += float(str(v).replace(",", "")) * scale[u]
        return tot
    except Exception:
        return None


def path_agreement(dev, n_feats):
    """north_star: 'end-to-end path agreement is reported against the reference on the same synthetic mu_x/y' --
    % of valid frames of the bench batch whose token agrees with (torch fp32 log-prior -> the reference's compiled
    core.pyx MAS); the C restatement stands in when the compiled reference is not on the box."""
    import oracle
    import face_gan_tts_b200 as fgt
    from face_gan_tts_b200 import synthetic

    mu_x, y, t_x, t_y = synthetic.lrs2_batch(B, n_feats, TX, TY, seed=1234)
    mu_d, y_d = mu_x.to(dev), y.to(dev)
    res = fgt.log_prior_maximum_path(mu_d, y_d, t_x, t_y, dense_path=False)
    ref_lp = oracle.log_prior_reference(mu_d, y_d).cpu().numpy()
    core = oracle.reference_core("asis")
    paths = np.zeros(ref_lp.shape, np.int32)
    (core or oracle).maximum_path_c(paths, ref_lp.copy(), t_x.numpy(), t_y.numpy())
    _, ref_ft = oracle.durations_and_frame_token(paths)
    ft = res.frame_token.cpu().numpy()
    valid = ref_ft >= 0
    return {"n_feats": n_feats, "frames": int(valid.sum()), "agree_pct": float((ft[valid] == ref_ft[valid]).mean() * 100.0),
            "against": "torch fp32 log-prior -> " + ("reference core.pyx (compiled)" if core is not None else "C restatement of core.pyx")}


def run_sweep(dev, peak, K):
    """Throughput-regime workloads of BASELINE.json (configs[3], configs[4] at one GPU), each CUDA-event timed on buffers
    larger than L2, with the roofline on padded cells, on VALID cells, and -- where a capture is committed -- on the
    DRAM bytes ncu saw."""
    import face_gan_tts_b200 as fgt
    from face_gan_tts_b200 import _lib, synthetic

    L = _lib.lib()
    stream = torch.cuda.current_stream(dev)
    sp = stream.cuda_stream
    out = []

    def entry(name, form, Bn, Fn, Tx, Ty, ms, bytes_padded, bytes_valid, valid_cells, ncu=None, note=None, extra=None):
        cells = Bn * Tx * Ty
        e = {"workload": name, "form": form, "B": Bn, "n_feats": Fn, "T_text": Tx, "T_mel": Ty, "ms": ms,
             "cells_per_s": cells / (ms * 1e-3), "valid_cells_per_s": valid_cells / (ms * 1e-3),
             "roofline": {"bound": "hbm", "peak": peak, "unit": "GB/s",
                          "achieved": bytes_padded / (ms * 1e-3) / 1e9, "frac": bytes_padded / (ms * 1e-3) / 1e9 / peak,
                          "achieved_valid": bytes_valid / (ms * 1e-3) / 1e9,
                          "frac_valid": bytes_valid / (ms * 1e-3) / 1e9 / peak}}
        tr = ncu_dram_bytes(ncu) if ncu else None
        e["roofline"]["traffic"] = tr
        if tr:
            e["roofline"]["frac_dram"] = tr / (ms * 1e-3) / 1e9 / peak
        if note:
            e["note"] = note
        if extra:
            e.update(extra)
        out.append(e)

    reps = max(5, min(K, 20))
    # ---- the reference's default n_feats = 128 at the LRS2 batch shape: the fused kernel's pair form (a 2-CTA cluster per
    # utterance; one CTA cannot hold two M-tiles of A in tensor memory at F = 128)
    Bn, Fn, Tx, Ty = B, 128, TX, TY
    mu_x, y, t_x, t_y = synthetic.lrs2_batch(Bn, Fn, Tx, Ty, seed=98)
    mu_d, y_d, tx_d, ty_d = mu_x.to(dev), y.to(dev), t_x.to(dev), t_y.to(dev)
    vc = int((t_x.long() * t_y.long()).sum())
    plans = [fgt.AlignmentPlan(Bn, Fn, Tx, Ty, device=dev, dense_path=True) for _ in range(4)]     # rotate > L2
    ms = cuda_time(stream, dev, lambda i: plans[i % 4](mu_d, y_d, tx_d, ty_d), max(reps, 20))
    entry("configs[1] shape at the reference default n_feats=128", "fused kernel (pair form), dense fp32 path", Bn, Fn, Tx, Ty, ms,
          4 * Fn * Bn * (Tx + Ty) + 4 * Bn * Tx * Ty + 4 * Bn * (Tx + Ty),
          4 * Fn * int((t_x.long() + t_y.long()).sum()) + 4 * Bn * Tx * Ty + 4 * Bn * (Tx + Ty), vc,
          note="round 1 / the one-CTA kernel: serial form (split-M log-prior -> HBM -> MAS -> expand), 69 us")
    del plans, mu_d, y_d
    # ---- configs[4] at one GPU: B = 1024 of the LRS2 shape
    Bn, Fn, Tx, Ty = 1024, F, TX, TY
    mu_x, y, t_x, t_y = synthetic.lrs2_batch(Bn, Fn, Tx, Ty, seed=99)
    mu_d, y_d, tx_d, ty_d = mu_x.to(dev), y.to(dev), t_x.to(dev), t_y.to(dev)
    vc = int((t_x.long() * t_y.long()).sum())
    in_pad = 4 * Fn * Bn * (Tx + Ty)
    in_val = 4 * Fn * int((t_x.long() + t_y.long()).sum())
    io_small = 4 * Bn * (Tx + Ty)
    for dense in (True, False):
        plan = fgt.AlignmentPlan(Bn, Fn, Tx, Ty, device=dev, dense_path=dense)
        ms = cuda_time(stream, dev, lambda i: plan(mu_d, y_d, tx_d, ty_d), reps)
        pathb = 4 * Bn * Tx * Ty if dense else 0
        entry("configs[4] @1 GPU: LRS2 shape, B=1024", "fused kernel, " + ("dense fp32 path" if dense else "index outputs only"),
              Bn, Fn, Tx, Ty, ms, in_pad + pathb + io_small, in_val + pathb + io_small, vc,
              ncu="r2_ncu_fused_B1024.json" if dense else None,
              note="algorithmic bytes = 4F(Tx+Ty) inputs + durations/frame_token" + (" + 4 B/cell dense path" if dense else "") +
                   "; the [Tx,Ty] value matrix never exists in memory")
        del plan
    value = fgt.log_prior(mu_d, y_d)
    ws_bytes = L.mas_b200_workspace_bytes(Bn, Tx, Ty)
    ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev)
    dur = torch.empty((Bn, Tx), dtype=torch.int32, device=dev)
    ft = torch.empty((Bn, Ty), dtype=torch.int32, device=dev)
    st = torch.empty((Bn,), dtype=torch.int32, device=dev)

    def mas_only(i):
        _lib.check(L.mas_b200_maximum_path(value.data_ptr(), Tx * Ty, Ty, tx_d.data_ptr(), ty_d.data_ptr(), Bn, Tx, Ty, -1e9,
                                           None, _lib.PATH_NONE, dur.data_ptr(), ft.data_ptr(), st.data_ptr(), ws.data_ptr(),
                                           ws_bytes, sp), "maximum_path")

    ms = cuda_time(stream, dev, mas_only, reps)
    entry("configs[4] @1 GPU: LRS2 shape, B=1024", "maximum_path alone (value matrix resident in HBM), index outputs", Bn, Fn, Tx, Ty,
          ms, 4 * Bn * Tx * Ty + io_small, 4 * vc + io_small, vc, ncu="r2_ncu_mas_forward_B1024.json",
         
          note="algorithmic bytes = 4 B per value cell read; cells outside [0,t_x)x[0,t_y) are never read")
    del value, ws, mu_d, y_d
    torch.cuda.empty_cache()

    # ---- configs[3]: long-utterance stress, value matrix streamed (B=64, 512 x 4096)
    Bn, Tx, Ty = 64, 512, 4096
    v, t_x, t_y = synthetic.mas_value(Bn, Tx, Ty, seed=5, tx_lo=256, ty_lo=2048)
    v_d, tx_d, ty_d = v.to(dev), t_x.to(dev), t_y.to(dev)
    vc = int((t_x.long() * t_y.long()).sum())
    for dense in (True, False):
        ms = cuda_time(stream, dev, lambda i: fgt.align(v_d, tx_d, ty_d, dense_path=dense), reps)
        pathb = 4 * Bn * Tx * Ty if dense else 0
        entry("configs[3]: long-utterance stress", "maximum_path(value) streamed, " + ("dense fp32 path" if dense else "index outputs only"),
              Bn, 0, Tx, Ty, ms, 4 * Bn * Tx * Ty + pathb, 4 * vc + pathb, vc,
              note="direction bits (262 KB/utterance) in L2-resident global scratch; host time of the functional API included")
    del v_d
    mu_x, y, t_x, t_y = synthetic.lrs2_batch(Bn, F, Tx, Ty, seed=6, tx_lo=256, ty_lo=2048)
    mu_d, y_d, tx_d, ty_d = mu_x.to(dev), y.to(dev), t_x.to(dev), t_y.to(dev)
    vc = int((t_x.long() * t_y.long()).sum())
    plan = fgt.AlignmentPlan(Bn, F, Tx, Ty, device=dev, dense_path=True)
    ms = cuda_time(stream, dev, lambda i: plan(mu_d, y_d, tx_d, ty_d), reps)
    entry("configs[3]: long-utterance stress", "log-prior + MAS, serial form (Tx > 256: split-M tcgen05 log-prior -> HBM -> MAS -> expand), dense",
          Bn, F, Tx, Ty, ms, 4 * F * Bn * (Tx + Ty) + 4 * Bn * Tx * Ty, 4 * F * int((t_x.long() + t_y.long()).sum()) + 4 * Bn * Tx * Ty, vc,
          note="algorithmic bytes as for the fused form (6.0 B/cell class); the serial form really moves 3 x 4 B/cell")
    del plan, mu_d, y_d
    torch.cuda.empty_cache()
    return out


def run_cuda(args):
    import face_gan_tts_b200 as fgt
    from face_gan_tts_b200 import _lib, sharding, synthetic

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    local_world = int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))
    host_cores = bind_rank_to_cores(local_rank, local_world) if world > 1 else None
    dist = None
    if world > 1:
        import torch.distributed as dist_mod

        dist = dist_mod
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # the only collective is a 24 KB all-gather that runs beside the next step's kernel: one channel (one CTA) is plenty
        os.environ.setdefault("NCCL_MAX_NCHANNELS", os.environ.get("MAS_B200_NCCL_CHANNELS", "1"))
        os.environ.setdefault("NCCL_MIN_NCHANNELS", "1")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    L = _lib.lib()
    K, W = args.steps, max(args.warmup, 3)
    SETTLE = 96    # untimed steps ahead of the warm-up (a multiple of the buffer-set and event rings)
    NSETS = 6      # ~37 MB per set (inputs + dense path): 6 sets = 220 MB > 126 MB L2

    # ---- device-resident buffer sets (each rank its own utterances: independent shards)
    sets = []
    for s in range(NSETS):
        mu_x, y, t_x, t_y = synthetic.lrs2_batch(B, F, TX, TY, seed=1234 + 1000 * rank + s)
        d = dict(mu=mu_x.to(dev), y=y.to(dev), tx=t_x.to(dev), ty=t_y.to(dev),
                 path=torch.empty((B, TX, TY), dtype=torch.float32, device=dev),
                 dur=torch.empty((B, TX), dtype=torch.int32, device=dev),
                 ft=torch.empty((B, TY), dtype=torch.int32, device=dev),
                 status=torch.empty((B,), dtype=torch.int32, device=dev))
        sets.append(d)
    ws_bytes = L.mas_b200_fused_workspace_bytes(B, F, TX, TY)
    wss = [torch.empty((ws_bytes,), dtype=torch.uint8, device=dev) for _ in range(NSETS)]
    torch.cuda.synchronize(dev)
    stream = torch.cuda.current_stream(dev)
    sp = stream.cuda_stream

    def fused(d, ws, dense=True, on=None):
        rc = L.mas_b200_log_prior_maximum_path(
            d["mu"].data_ptr(), d["y"].data_ptr(), d["tx"].data_ptr(), d["ty"].data_ptr(), B, F, TX, TY, -1e9,
            d["path"].data_ptr() if dense else None, _lib.PATH_F32 if dense else _lib.PATH_NONE,
            d["dur"].data_ptr(), d["ft"].data_ptr(), d["status"].data_ptr(), ws.data_ptr(), ws_bytes,
            _lib.LP_AUTO, sp if on is None else on.cuda_stream)
        _lib.check(rc, "mas_b200_log_prior_maximum_path")

    # multi-GPU duration gather: one-sided (the fused kernel stores its durations into every rank's symmetric-memory
    # buffer over NVLink: no collective kernel, no torch.distributed call per step) when symmetric memory is available,
    # else an asynchronous NCCL all-gather per step on a side stream
    put = None
    if dist and not args.nccl_gather and not os.environ.get("MAS_B200_BENCH_NO_GATHER"):
        try:
            put = sharding.OneSidedDurationGather(B, TX, dev)
            if args.put_in_kernel:
                put.enable()         # the fused kernel stores its durations itself (+ the NVLink round trip at its end)
        except Exception as ex:
            sys.stderr.write(f"bench.py: one-sided duration gather unavailable ({ex!r}); NCCL all-gather per step\n")
            put = None
    graphs = [None]      # multi-GPU: CUDA graphs of the fused call, one per buffer set (see below)
    gl = g_ = None
    gathered = [torch.empty((world * B, TX), dtype=torch.int32, device=dev) for _ in range(2)] if dist else None
    comm_stream = torch.cuda.Stream(dev) if dist else None
    step_done = [torch.cuda.Event() for _ in range(4)] if dist else None

    def step(i):
        d = sets[i % NSETS]
        if graphs[0] is not None:
            graphs[0][i % NSETS].replay()
        else:
            fused(d, wss[i % NSETS])
        if dist and put is not None and not args.put_in_kernel:
            # one-sided gather off the step's critical path: a small copy kernel on a side stream, ordered behind the step
            ev = step_done[i % 4]
            ev.record(stream)
            comm_stream.wait_event(ev)
            put.put(d["dur"], comm_stream)
        if dist and put is None and not os.environ.get("MAS_B200_BENCH_NO_GATHER"):
            # loss-bookkeeping collective: durations of every rank, asynchronous on a side stream
            ev = step_done[i % 4]
            ev.record(stream)
            comm_stream.wait_event(ev)
            with torch.cuda.stream(comm_stream):
                sharding.all_gather_durations_into(gathered[i % 2], d["dur"])

    def barrier():
        if dist:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # the clock sampler opens NVML (tens of ms) BEFORE the warm-up, so that nothing but the barrier sits between the warm-up
    # steps and the timed region
    try:
        dev_uuid = "GPU-" + str(torch.cuda.get_device_properties(dev).uuid)
    except Exception:
        dev_uuid = None
    sampler = ClockSampler(local_rank, dev_uuid)
    if rank == 0 and os.environ.get("MAS_B200_BENCH_CLOCKS", "on") != "off":
        sampler.start()
    # SETTLE untimed steps ahead of the W warm-up steps: the first window of a process is slower than every later one
    # (N = 2, 20 steps: 824 us against 780 us for the same window ~100 steps later, scripts/diag/step_profile_dist.py), and
    # W = 5 steps do not even touch all NSETS buffer sets.  Reported in the JSON line (`settle_steps_untimed`).
    for i in range(SETTLE):
        step(i)
    for i in range(W):
        step(i)
    barrier()
    # eager launches keep the programmatic dependent launch between consecutive steps (~2 us per step); a CUDA graph per
    # step does not, but costs almost no host time.  Replay graphs only when the host cannot enqueue a step (fused call +
    # NCCL all-gather) in less time than the GPU needs for it.
    use_graph = False
    if dist and not args.no_graph and put is None:
        e_a, e_b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e_a.record(stream)
        t0 = time.perf_counter()
        for i in range(24):
            step(W + i)
        host_us = (time.perf_counter() - t0) / 24 * 1e6
        e_b.record(stream)
        barrier()
        dev_us = e_a.elapsed_time(e_b) / 24 * 1e3
        flag = torch.tensor([1.0 if host_us > 0.85 * dev_us else 0.0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MAX)          # every rank takes the same decision
        use_graph = bool(flag.item() > 0) or args.graph
    if dist and use_graph:
        # the fused call (ONE kernel node) replayed from a CUDA graph per buffer set; the all-gather stays an eager NCCL
        # call on its own stream.
        try:
            cap = torch.cuda.Stream(dev)
            cap.wait_stream(stream)
            gl = []
            for s_ in range(NSETS):
                g_ = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g_, stream=cap):
                    fused(sets[s_], wss[s_], on=torch.cuda.current_stream(dev))
                gl.append(g_)
            torch.cuda.synchronize(dev)
            graphs[0] = gl
            for i in range(NSETS):
                step(i)
            barrier()
        except Exception as ex:          # capture not possible: keep the eager launches
            sys.stderr.write(f"bench.py: CUDA-graph capture failed ({ex!r}); eager launches\n")
            graphs[0] = None
            torch.cuda.synchronize(dev)
    mark = sampler.mark()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    host_t0 = time.perf_counter()
    for i in range(K):
        step(W + i)
    host_enqueue_us = (time.perf_counter() - host_t0) / K * 1e6      # host time to enqueue one step (no sync inside)
    if dist and (put is None or not args.put_in_kernel):
        stream.wait_stream(comm_stream)      # all gathers / puts complete inside the timed region
    e1.record(stream)
    barrier()                                # one-sided gather: every rank's puts are complete and visible after this
    ms_total = e0.elapsed_time(e1)
    clocks = None
    if rank == 0:
        # fewer than two samples inside a window of a few milliseconds: keep the same kernel running (untimed, no gather) until
        # the sampler has them
        held = sampler.hold(mark, lambda i: fused(sets[i % NSETS], wss[i % NSETS]), lambda: torch.cuda.synchronize(dev))
        clocks = sampler.stop(mark, held)
    if put is not None:
        # every rank's block of the gather buffer holds real durations (each utterance's sum to its t_y: 1..TY), and this
        # rank's block is what its own last step produced
        g = put.gathered.view(world, B, TX)
        sums = g.sum(-1)
        assert bool(((sums >= 1) & (sums <= TY)).all()), "one-sided gather: a rank's durations did not arrive"
        assert torch.equal(g[rank], sets[(W + K - 1) % NSETS]["dur"]), "one-sided gather: own block differs"
        put.disable()
    if dist:
        t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_step = ms_total / K
    value = world * CELLS / (ms_step * 1e-3)

    # ---- configs[4]: strong scaling of a fixed total batch over the ranks (every rank: B_total / world utterances of the
    # LRS2 shape + the asynchronous all-gather of the durations), device-timed, max over ranks
    strong = None
    if dist:
        strong = []
        for b_total in (64, 128, 256, 512, 1024):
            if b_total % world:
                continue
            bl = b_total // world
            mu_x, y, t_x, t_y = synthetic.lrs2_batch(bl, F, TX, TY, seed=777 + rank)
            args_d = (mu_x.to(dev), y.to(dev), t_x.to(dev), t_y.to(dev))
            plan = fgt.AlignmentPlan(bl, F, TX, TY, device=dev, dense_path=True)
            gat = torch.empty((b_total, TX), dtype=torch.int32, device=dev)
            sput = None
            if put is not None:
                try:        # the fused kernel stores its durations into every rank's buffer itself: one launch per step
                    sput = sharding.OneSidedDurationGather(bl, TX, dev)
                    sput.enable()
                except Exception:
                    sput = None

            def sstep(i):
                r = plan(*args_d)
                if sput is None:
                    comm_stream.wait_stream(stream)
                    with torch.cuda.stream(comm_stream):
                        sharding.all_gather_durations_into(gat, r.durations)
                    stream.wait_stream(comm_stream)      # the next step overwrites plan.durations

            for i in range(3):
                sstep(i)
            barrier()
            a_, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 20
            a_.record(stream)
            for i in range(reps):
                sstep(i)
            b_.record(stream)
            barrier()
            t = torch.tensor([a_.elapsed_time(b_) / reps], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            strong.append({"B_total": b_total, "B_per_gpu": bl, "ms_per_step": float(t.item()),
                           "cells_per_s": b_total * TX * TY / (float(t.item()) * 1e-3),
                           "gather": "one-sided, by the fused kernel" if sput is not None else "NCCL all-gather"})
            if sput is not None:
                assert bool((sput.gathered.view(world, bl, TX).sum(-1) >= 1).all()), "strong scaling: a rank's durations did not arrive"
                sput.disable()
            del plan, gat, args_d, sput
        torch.cuda.empty_cache()

    # ---- e2e (every rank): HOST buffers in, results back to pinned host memory, every step, inside the timed region.
    # Software-pipelined like a training loop: the H2D of step i+1 runs on a copy stream while step i computes, and the
    # host consumes the result of step i-1 (event sync) while step i is in flight.
    #   packed   the batch arrives as a collate that does not pad produces it (fgt.pack_batch layout: lengths + the valid
    #            part of every row, one pinned buffer): ONE copy-engine transfer + the device unpack kernel
    #   padded   the padded pinned tensors, four cudaMemcpyAsync
    NH = 3
    host_in = [[t.pin_memory() for t in synthetic.lrs2_batch(B, F, TX, TY, seed=4321 + 100 * rank + k)] for k in range(NH)]
    host_packed = [fgt.pack_batch(*h) for h in host_in]          # outside the timed region: this IS the input format
    dur_h = [torch.empty((B, TX), dtype=torch.int32).pin_memory() for _ in range(2)]
    ft_h = [torch.empty((B, TY), dtype=torch.int32).pin_memory() for _ in range(2)]
    path_h = [torch.empty((B, TX, TY), dtype=torch.float32).pin_memory() for _ in range(2)]
    copy_stream = torch.cuda.Stream(dev)
    plans = [fgt.AlignmentPlan(B, F, TX, TY, device=dev, dense_path=True) for _ in range(2)]
    plans_idx = [fgt.AlignmentPlan(B, F, TX, TY, device=dev, dense_path=False) for _ in range(2)]

    NST = 3                                               # device staging buffers of the packed batch
    staging = [torch.empty((max(p.numel() for p in host_packed),), dtype=torch.uint8, device=dev) for _ in range(NST)]
    # events are REUSED (small rings): creating a fresh event per step inside the loop costs a driver call per step, and
    # with 4-8 ranks on one host those calls serialise across the processes (scripts/diag/e2e_multirank.py: 169 -> 285 us
    # per step at 4 ranks although every GPU still gets its 54 GB/s of PCIe, scripts/diag/h2d_multiproc.py)
    EV = 4
    h2d_done = [torch.cuda.Event() for _ in range(EV)]
    unpacked = [torch.cuda.Event() for _ in range(NST)]
    res_done = [torch.cuda.Event() for _ in range(2)]

    def run_e2e(nsteps, dense_d2h, mode):
        checksum = 0

        def enqueue_h2d(i):
            d = sets[i % NSETS]
            with torch.cuda.stream(copy_stream):
                if mode == "packed":
                    # the copy stream carries NOTHING but the host->device copies, so the copy engine runs back to back;
                    # the unpack kernel of step i runs on the compute stream in front of the step.  Staging buffer reuse:
                    # copy i overwrites what unpack i - NST read.
                    if i >= NST:
                        copy_stream.wait_event(unpacked[i % NST])
                    p = host_packed[i % NH]
                    staging[i % NST][:p.numel()].copy_(p, non_blocking=True)
                else:
                    if i >= 2:
                        copy_stream.wait_event(res_done[i & 1])        # input set reuse: step i-2 is done with its tensors
                    mu_h, y_h, tx_h, ty_h = host_in[i % NH]
                    d["mu"].copy_(mu_h, non_blocking=True)
                    d["y"].copy_(y_h, non_blocking=True)
                    d["tx"].copy_(tx_h, non_blocking=True)
                    d["ty"].copy_(ty_h, non_blocking=True)
                h2d_done[i % EV].record(copy_stream)

        pl = plans if dense_d2h else plans_idx
        enqueue_h2d(0)
        for i in range(nsteps):
            if i + 1 < nsteps:
                enqueue_h2d(i + 1)
            d = sets[i % NSETS]
            stream.wait_event(h2d_done[i % EV])
            if mode == "packed":
                fgt.unpack_batch(staging[i % NST], B, F, TX, TY, out=(d["mu"], d["y"], d["tx"], d["ty"]))
                unpacked[i % NST].record(stream)
            res = pl[i & 1](d["mu"], d["y"], d["tx"], d["ty"])          # AlignmentPlan: reusable outputs + workspace
            dur_h[i & 1].copy_(res.durations, non_blocking=True)
            ft_h[i & 1].copy_(res.frame_token, non_blocking=True)
            if dense_d2h:
                path_h[i & 1].copy_(res.path, non_blocking=True)
            res_done[i & 1].record(stream)
            if i >= 1:
                res_done[(i - 1) & 1].synchronize()        # the caller consumes result i-1
                checksum += int(dur_h[(i - 1) & 1][0, 0])
        res_done[(nsteps - 1) & 1].synchronize()
        return checksum

    def time_e2e(dense_d2h, mode):
        run_e2e(4, dense_d2h, mode)
        barrier()
        t0 = time.perf_counter()
        run_e2e(K, dense_d2h, mode)
        torch.cuda.synchronize(dev)
        sec = (time.perf_counter() - t0) / K
        if dist:
            t = torch.tensor([sec], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            sec = float(t.item())
        return sec

    h2d_padded = 4 * F * B * (TX + TY) + 8 * B
    h2d_packed = int(sum(p.numel() for p in host_packed) / NH)
    sec_e2e = time_e2e(False, "packed")
    sec_e2e_padded = time_e2e(False, "padded")
    sec_e2e_dense = time_e2e(True, "packed")
    d2h_idx = 4 * B * (TX + TY)
    e2e = {"value": world * CELLS / sec_e2e, "unit": UNIT, "h2d_bytes_per_step": h2d_packed,
           "d2h_bytes_per_step": d2h_idx, "ms_per_step": sec_e2e * 1e3,
           "input": "packed ragged batch in pinned host memory (lengths + valid part of every mu_x / y row: what a collate "
                    "that does not pad produces, fgt.pack_batch layout)",
           "h2d": "ONE cudaMemcpyAsync per step on a stream that carries nothing else (copy engine back to back) + mas_b200_unpack_batch on the compute stream in front of the step (zero-padded tensors)",
           "result": "durations [B,Tx] + frame->token index [B,Ty] to pinned host memory (the dense path stays in HBM for mu_y)",
           "pipelining": "H2D of step i+1 on a copy stream under step i; host consumes result i-1 while step i runs; events reused from small rings"}
    e2e_padded = {"value": world * CELLS / sec_e2e_padded, "unit": UNIT, "h2d_bytes_per_step": h2d_padded,
                  "d2h_bytes_per_step": d2h_idx, "ms_per_step": sec_e2e_padded * 1e3,
                  "h2d": "cudaMemcpyAsync of the padded pinned tensors (what relocate_input does, face_tts.py:85-89)"}
    e2e_dense = {"value": world * CELLS / sec_e2e_dense, "unit": UNIT, "h2d_bytes_per_step": h2d_packed,
                 "d2h_bytes_per_step": d2h_idx + 4 * CELLS, "ms_per_step": sec_e2e_dense * 1e3,
                 "result": "as e2e + the dense fp32 path [B,Tx,Ty] copied back (what maximum_path literally returns)"}

    out = None
    if rank == 0:
        mas_ws = L.mas_b200_workspace_bytes(B, TX, TY)
        vals = [torch.empty((B, TX, TY), dtype=torch.float32, device=dev) for _ in range(NSETS)]

        def lp_only(i):
            d = sets[i % NSETS]
            _lib.check(L.mas_b200_log_prior(d["mu"].data_ptr(), d["y"].data_ptr(), B, F, TX, TY,
                                            vals[i % NSETS].data_ptr(), _lib.LP_AUTO, sp), "log_prior")

        def mas_only(i, dense=False):
            d = sets[i % NSETS]
            _lib.check(L.mas_b200_maximum_path(
                vals[i % NSETS].data_ptr(), TX * TY, TY, d["tx"].data_ptr(), d["ty"].data_ptr(), B, TX, TY, -1e9,
                d["path"].data_ptr() if dense else None, _lib.PATH_F32 if dense else _lib.PATH_NONE,
                d["dur"].data_ptr(), d["ft"].data_ptr(), d["status"].data_ptr(), wss[i % NSETS].data_ptr(), mas_ws,
                sp), "maximum_path")

        def serial_form(i):
            prev = _lib.set_option("fused_impl", 1)
            try:
                fused(sets[i % NSETS], wss[i % NSETS])
            finally:
                _lib.set_option("fused_impl", prev)

        for i in range(NSETS):
            lp_only(i)
        t_fused = cuda_time(stream, dev, lambda i: fused(sets[i % NSETS], wss[i % NSETS], True), K)
        t_fused_idx = cuda_time(stream, dev, lambda i: fused(sets[i % NSETS], wss[i % NSETS], False), K)
        t_serial = cuda_time(stream, dev, serial_form, K)
        t_lp = cuda_time(stream, dev, lp_only, K)
        t_mas = cuda_time(stream, dev, lambda i: mas_only(i, False), K)
        t_mas_dense = cuda_time(stream, dev, lambda i: mas_only(i, True), K)
        valid_cells = int(sum((d["tx"].long() * d["ty"].long()).sum().item() for d in sets) / NSETS)
        valid_in = int(sum((d["tx"].long() + d["ty"].long()).sum().item() for d in sets) / NSETS) * 4 * F
        peak, peak_src = load_peaks()
        # dominant (only) kernel of the step: lp_mas_fused_kernel.  Algorithmic bytes (fused form):
        # 4F(Tx+Ty) per utterance in + 4 B/cell dense path out (+ the two index outputs); the value matrix costs nothing.
        step_bytes = 4 * F * B * (TX + TY) + 4 * CELLS + 4 * B * (TX + TY)
        step_bytes_valid = valid_in + 4 * CELLS + 4 * B * (TX + TY)
        achieved = step_bytes / (t_fused * 1e-3) / 1e9
        traffic = ncu_dram_bytes("r2_ncu_fused_pair_B32.json") or ncu_dram_bytes("r2_ncu_fused_B32.json")
        roofline = {
            "bound": "hbm", "kernel": "lp_mas_fused_kernel<10,1,pair> (log-prior + MAS + dense path; a 2-CTA cluster per utterance: one 128-row "
                                      "M-tile and one DP warp per CTA, halo row / direction words cross with st.async)",
            "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
            "achieved_valid": step_bytes_valid / (t_fused * 1e-3) / 1e9,
            "frac_valid": step_bytes_valid / (t_fused * 1e-3) / 1e9 / peak,
            "peak_source": peak_src, "algorithmic_bytes_per_launch": step_bytes,
            "kernel_ms": t_fused,
            "kernels_ms": {"lp_mas_fused (dense path)": t_fused, "lp_mas_fused (index outputs only)": t_fused_idx,
                           "serial form: log_prior + mas_forward + path_expand": t_serial,
                           "log_prior alone (tcgen05, -> HBM)": t_lp, "mas_forward alone (index outputs)": t_mas,
                           "mas_forward + path_expand": t_mas_dense},
            "tensor_pipe": "3xTF32 tcgen05.mma, A in TMEM: 31 MMAs (M=128,N=32,K=8) per 32-frame tile and CTA",
            "note": "2B=64 CTAs on 148 SMs: the step is bound by the T_mel-long dependency chain of the DP "
                    "(~50 cycles/frame in one warp per 128 text rows; profiles/r2_phase_cycles.txt), not by HBM; the throughput "
                    "regime is in `sweep`",
        }
        if traffic:
            roofline["frac_dram"] = traffic / (t_fused * 1e-3) / 1e9 / peak

        # ---- CPU baseline on this box's host cores (bounded sample: a few full-batch steps)
        try:
            cstep, kind, cores, sample = reference_step_fn()
            sec = time_cpu(cstep, 3, 1)
            cpu = {"value": CELLS / sec, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample + "; 3 steps",
                   "ms_per_step": sec * 1e3}
            try:    # informational: the same core.pyx rebuilt with -fopenmp (its prange over the batch then uses every core)
                ostep, okind, ocores, _ = reference_step_fn("omp")
                if okind == "reference":
                    osec = time_cpu(ostep, 3, 1)
                    cpu["openmp_rebuild"] = {"value": CELLS / osec, "ms_per_step": osec * 1e3, "cores": ocores,
                                             "note": "not how the reference ships (setup.py links no OpenMP)"}
            except Exception:
                pass
        except Exception as ex:  # the oracle is a reported baseline, never a dependency of the product
            cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "unavailable", "sample": repr(ex)}

        # ---- end-to-end path agreement with the reference on the bench batch (north_star), F = 80 and the reference
        # default n_feats = 128
        try:
            agreement = [path_agreement(dev, 80), path_agreement(dev, 128)]
        except Exception as ex:
            agreement = {"error": repr(ex)[:200]}

        # ---- configs[2] at the bench shape: the whole alignment block of compute_loss (face_tts.py:159-218,233-234),
        # forward + backward w.r.t. mu_x / logw.  "reference" = the reference's formulation on the SAME GPU tensors
        # (torch log-prior GEMMs on the device, its maximum_path wrapper bouncing value/mask to the host for the
        # compiled core.pyx and back, dense attn consumers) -- what a training step pays today.
        block, sweep = None, None
        if world == 1:
            try:
                block = time_compute_loss_block(dev)
            except Exception as ex:
                block = {"error": repr(ex)[:200]}
            try:
                sweep = run_sweep(dev, peak, K)
            except Exception as ex:
                sweep = {"error": repr(ex)[:300]}

        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W, "settle_steps_untimed": SETTLE,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {**workload_config(), "valid_cells_per_step": valid_cells,
                       "outputs": "dense fp32 path [B,Tx,Ty] + durations [B,Tx] + frame->token index [B,Ty]",
                       "cache": f"inputs rotate over {NSETS} buffer sets (~{NSETS * 37} MB) larger than the 126 MB L2",
                       "parallelism": (f"utterance shards x{world}, " + (("one-sided duration gather over NVLink into every rank's symmetric-memory buffer: " + ("by the fused kernel itself" if args.put_in_kernel else "a copy kernel on a side stream, no collective")) if put is not None else "async NCCL all-gather of durations")) if world > 1
                       else "single GPU"},
            "roofline": roofline, "cpu_baseline": cpu, "path_agreement": agreement, "sweep": sweep,
            "strong_scaling_configs4": strong, "compute_loss_block": block,
            "e2e": e2e, "e2e_padded_copy": e2e_padded, "e2e_dense_path": e2e_dense,
            "gpu_launches": K * (2 if (dist and put is not None and not args.put_in_kernel) else 1),
            "launches_per_step": 2 if (dist and put is not None and not args.put_in_kernel) else 1,
            "pipeline": "ONE kernel per step: lp_mas_fused_kernel (tcgen05 log-prior -> shared-memory ring -> MAS -> backtrack "
                        "-> dense path), programmatic dependent launch" + ("; + the duration put kernel on a side stream" if (dist and put is not None and not args.put_in_kernel) else ""),
            "host_enqueue_us_per_step": host_enqueue_us, "host_cores_per_rank": host_cores,
            "launch": "CUDA graph replay of the fused call + eager NCCL all-gather" if graphs[0] is not None else "eager",
            "clocks": clocks,
        }
    if dist:
        dist.barrier()
        # CUDA graphs go before the communicator does: destroy_process_group() never returns while a graph that captured
        # NCCL work is alive (scripts/diag/graph_nccl_probe.py) -- none of the graphs here does, but keep the order
        graphs[0] = None
        gl = g_ = None
        import gc
        gc.collect()
        torch.cuda.synchronize(dev)
        dist.destroy_process_group()
    if out is not None:
        print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--no-graph", action="store_true", help="multi-GPU: always eager launches")
    ap.add_argument("--graph", action="store_true", help="multi-GPU: always replay the fused call from a CUDA graph")
    ap.add_argument("--put-in-kernel", action="store_true", help="multi-GPU: the fused kernel stores its durations into the peers' buffers itself")
    ap.add_argument("--nccl-gather", action="store_true", help="multi-GPU: NCCL all-gather of the durations per step instead of the one-sided gather")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product has no CPU path; use --impl reference for the CPU arm)")
    run_cuda(args)


if __name__ == "__main__":
    main()
