#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/pytest_gpu.log
echo "== bench"; timeout 600 python bench.py > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "rc=$?"; cat gpurun_out/bench.log; tail -5 gpurun_out/bench.err
echo "== sweep"; timeout 900 python scripts/sweep.py > gpurun_out/sweep.jsonl 2> gpurun_out/sweep.err; echo "rc=$?"; cat gpurun_out/sweep.jsonl; tail -5 gpurun_out/sweep.err
