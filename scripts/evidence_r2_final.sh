set -x
cd $GRAFT_REPO_ROOT
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -1 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
python bench.py --steps 100 --warmup 10 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_n1_steps20.json 2> gpurun_out/r2_bench_n1_steps20.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_reference_n1.json 2> gpurun_out/r2_bench_reference_n1.err
python scripts/fused_phases.py > gpurun_out/r2_phase_cycles_fused.txt 2>&1
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_bench_steps5.csv python bench.py --steps 5 --warmup 3 > gpurun_out/ncu_bench.log 2>&1
tail -c 300 gpurun_out/r2_bench_n1.json
