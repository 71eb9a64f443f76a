#!/bin/bash
# round-1 evidence run (overlapped pipeline): smoke, all GPU tests, bench (both arms), ncu launch list,
# full captures of the MAS kernel and the tcgen05 log-prior kernel (serial pipeline under ncu: the overlapped
# pipeline needs both kernels co-resident, which a replaying profiler does not allow).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
echo "== smoke" ; timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/pytest_gpu.log
echo "== bench"; timeout 600 python bench.py > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "rc=$?"; cat gpurun_out/bench.log; tail -5 gpurun_out/bench.err
echo "== bench reference"; timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.log 2>&1; cat gpurun_out/bench_ref.log
if [ -z "$SKIP_NCU" ]; then
export MAS_B200_PIPELINE=serial
echo "== ncu launch list (serial pipeline)"
timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches.csv python bench.py --steps 5 --warmup 3 > gpurun_out/ncu.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/ncu.log
echo "== ncu full (MAS kernel + log-prior kernel, cfg2 shape)"
timeout 300 python scripts/profile_fused.py > gpurun_out/prof_plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mas_forward -s 2 -c 1 -f -o gpurun_out/mas_full python scripts/profile_fused.py > gpurun_out/ncu_mas.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/ncu_mas.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:log_prior -s 2 -c 1 -f -o gpurun_out/lp_full python scripts/profile_fused.py > gpurun_out/ncu_lp.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/ncu_lp.log
fi
