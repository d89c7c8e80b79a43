#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "gemm or golden_forward or train_step" 2>&1 | tail -6
python scripts/gemm_bench.py 2>&1 | grep -E "fc1|dH|proj"
ROWS=4 python scripts/prof_full_fuser.py 2>&1 | tail -4
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r2_bench_full.json 2> gpurun_out/r2_bench_full.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_full.json'))
print({k:d.get(k) for k in ('value','ms_per_step','gpu_launches','e2e')})
print(d['roofline']); print(d.get('gpu_library_yardstick')); print(d.get('cpu_baseline')); print(d.get('full_fuser_fwd_bwd'))
for k,v in d['stages'].items():
    if v.get('ms_per_step',0)>0.3: print('   ',k, {a:(round(b,3) if isinstance(b,float) else b) for a,b in v.items() if a in ('ms_per_step','launches_per_step','avg_launch_us','frac')})
PY
tail -5 gpurun_out/r2_bench_full.err
