#!/bin/bash
# GPU check: parity tests + phase stamps + tuning sweep
mkdir -p gpurun_out
echo "== pytest mas"; timeout 1500 python -m pytest tests/test_gpu_mas.py -m gpu -q --timeout 600 > gpurun_out/pytest_mas.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/pytest_mas.log
echo "== phase"; timeout 300 python scripts/phase_times.py > gpurun_out/phase.log 2>&1; echo "rc=$?"; cat gpurun_out/phase.log
echo "== tune"; timeout 900 python scripts/tune_mas.py $TUNE_ARGS > gpurun_out/tune.log 2>&1; echo "rc=$?"; grep -v "^$" gpurun_out/tune.log | python -c "
import sys,ast
for l in sys.stdin:
    try: d=ast.literal_eval(l.strip())
    except Exception: print(l.strip()[:200]); continue
    print({k:(round(v,1) if isinstance(v,float) else v) for k,v in d.items() if k not in ('err',) or v})
"
