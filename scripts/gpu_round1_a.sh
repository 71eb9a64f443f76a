#!/bin/bash
# first GPU call: smoke, parity tests, tuning sweep, bench, ncu launch list
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
echo "== smoke" ; timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -5 gpurun_out/smoke.log
echo "== pytest mas"; timeout 1500 python -m pytest tests/test_gpu_mas.py -m gpu -q --timeout 600 > gpurun_out/pytest_mas.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/pytest_mas.log
echo "== pytest logprior"; timeout 900 python -m pytest tests/test_gpu_logprior.py -m gpu -q -s --timeout 600 > gpurun_out/pytest_lp.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/pytest_lp.log
echo "== tune"; timeout 900 python scripts/tune_mas.py > gpurun_out/tune.log 2>&1; echo "rc=$?"; tail -60 gpurun_out/tune.log
echo "== bench"; timeout 600 python bench.py --steps 100 --warmup 10 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "rc=$?"; cat gpurun_out/bench.log; tail -5 gpurun_out/bench.err
echo "== bench reference"; timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.log 2>&1; cat gpurun_out/bench_ref.log
echo "== ncu launch list"
timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv python bench.py --steps 5 --warmup 3 > gpurun_out/ncu.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu.log
