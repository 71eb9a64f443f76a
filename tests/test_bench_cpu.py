"""bench.py on a box without a GPU: the reference arm (the driver launches it exactly like the GPU arm) and the host-side
pieces of the GPU arm that need no device (clock sampler, the config both arms print)."""
import importlib.util
import json
import os
import subprocess
import sys
import threading
import time

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench_module():
    spec = importlib.util.spec_from_file_location("bench_under_test", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    argv = sys.argv
    sys.argv = ["bench.py"]
    try:
        spec.loader.exec_module(mod)
    finally:
        sys.argv = argv
    return mod


def _run_reference(extra_env=None, steps=1, warmup=0, gpus=1):
    env = dict(os.environ)
    env.update(extra_env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", str(gpus),
                           "--steps", str(steps), "--warmup", str(warmup)], capture_output=True, text=True, env=env,
                          timeout=600, cwd=ROOT)


def test_reference_arm_prints_one_json_line_on_the_gpu_arms_config():
    r = _run_reference(steps=2, warmup=1)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    bench = _bench_module()
    assert d["impl"] == "reference" and d["metric"] == bench.METRIC and d["unit"] == bench.UNIT
    assert d["higher_is_better"] is True and d["steps"] == 2 and d["warmup"] == 1 and d["n_gpus"] == 1
    assert d["config"] == bench.workload_config()                     # the keys the GPU arm prints too
    assert d["config"]["B_per_gpu"] == 32 and d["config"]["T_text"] == 190 and d["config"]["T_mel"] == 1000
    assert d["value"] > 0 and abs(d["value"] - 32 * 190 * 1000 / (d["ms_per_step"] * 1e-3)) < 1e-6 * d["value"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_runs_on_rank_0_only():
    r = _run_reference({"RANK": "1", "LOCAL_RANK": "1", "WORLD_SIZE": "2"}, gpus=2)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_clock_sampler_without_nvml_or_nvidia_smi_reports_why():
    bench = _bench_module()
    s = bench.ClockSampler(0, "GPU-00000000-0000-0000-0000-000000000000")
    s.start()
    if s.thread is not None:
        pytest.skip("this box has NVML or nvidia-smi")
    m = s.mark()
    assert s.hold(m, lambda i: None, lambda: None) == 0
    c = s.stop(m, 0)
    assert c["sm_mhz"] is None and c["samples"] == 0 and c["reasons"]


def test_clock_sampler_holds_the_step_until_two_samples_exist():
    """A 20-step window is shorter than any sampling period: hold() keeps the same step running (untimed) until the
    sampler has two samples newer than the mark, and stop() reports the window and the throttle reasons it saw."""
    bench = _bench_module()
    s = bench.ClockSampler(0, None)
    feed = iter([(1965.0, 1965.0, []), (1950.0, 1965.0, ["sw_power_cap"])] + [(1965.0, 1965.0, [])] * 10000)
    s.source = "nvml"
    s.thread = threading.Thread(target=s._nvml_loop, args=(lambda: next(feed),), daemon=True)
    mark = s.mark()
    assert mark == 0
    steps = []
    s.thread.start()
    held = s.hold(mark, steps.append, lambda: time.sleep(0.0005), want=2, limit_s=5.0)
    c = s.stop(mark, held)
    assert held == len(steps) and held % 8 == 0
    assert c["samples"] >= 2 and c["sm_max_mhz"] == 1965.0 and 1950.0 <= c["sm_mhz"] <= 1965.0
    assert c["reasons"] == ["sw_power_cap"] and c["source"] == "nvml"
    assert ("held" in c["window"]) == (held > 0)
