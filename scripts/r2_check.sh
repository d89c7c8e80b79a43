#!/bin/bash
# Round-2 GPU check: panel kernel unit check, erank-related parity tests, then a short bench.
mkdir -p gpurun_out
timeout 120 python scripts/dbg_panel.py sym > gpurun_out/r2_dbg_sym.log 2>&1; echo "dbg rc=$?"; tail -25 gpurun_out/r2_dbg_sym.log
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "erank or jacobi or altern or graph" > gpurun_out/r2_pytest_erank.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_pytest_erank.log
tail -5 gpurun_out/r2_pytest_erank.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_a.json 2> gpurun_out/r2_bench_a.err
echo "bench rc=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_a.json'))
print({k:d.get(k) for k in ('value','ms_per_step','gpu_launches')}, d.get('roofline'))
for k,v in d['stages'].items():
    if v.get('ms_per_step',0)>0.2: print(k, {a:(round(b,3) if isinstance(b,float) else b) for a,b in v.items() if a in ('ms_per_step','launches_per_step','avg_launch_us','frac')})
PY
