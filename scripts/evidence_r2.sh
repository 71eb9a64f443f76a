set -x
cd $GRAFT_REPO_ROOT
python bench.py --steps 100 --warmup 10 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_reference_n1.json 2> gpurun_out/r2_bench_reference_n1.err
(python scripts/fused_phases.py; python scripts/phase_times.py) > gpurun_out/r2_phase_cycles.txt 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_bench_steps5.csv python bench.py --steps 5 --warmup 3 > gpurun_out/ncu_bench.log 2>&1
FUSED_DENSE=1 timeout 200 ncu --set full --clock-control none --import-source on -k regex:lp_mas_fused -c 1 -s 3 -f -o gpurun_out/r2_fused_pair_B32 python scripts/fused_once.py > gpurun_out/ncu1.log 2>&1
FUSED_B=1024 timeout 200 ncu --set full --clock-control none --import-source on -k regex:lp_mas_fused -c 1 -s 3 -f -o gpurun_out/r2_fused_B1024 python scripts/fused_once.py > gpurun_out/ncu2.log 2>&1
PB=1024 timeout 200 ncu --set full --clock-control none --import-source on -k regex:mas_forward -c 1 -s 3 -f -o gpurun_out/r2_mas_forward_B1024 python scripts/profile_mas.py > gpurun_out/ncu3.log 2>&1
PB=32 timeout 200 ncu --set full --clock-control none --import-source on -k regex:mas_forward -c 1 -s 3 -f -o gpurun_out/r2_mas_forward_B32 python scripts/profile_mas.py > gpurun_out/ncu5.log 2>&1
FUSED_FULL=1 FUSED_B=74 timeout 300 ncu --set full --warp-sampling-interval 0 --clock-control none --import-source on -k regex:lp_mas_fused -c 1 -s 3 -f -o gpurun_out/r2_fused_pair_b74_si0 python scripts/fused_once.py > gpurun_out/ncu4.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -8
tail -n 2 gpurun_out/ncu1.log; tail -n 2 gpurun_out/ncu4.log
