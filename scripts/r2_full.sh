#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -8
python __graft_entry__.py smoke 2>&1 | tail -3
timeout 600 python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err; echo "ref rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_n1.json'))
print({k:d.get(k) for k in ('value','ms_per_step','gpu_launches','clocks')}, d['e2e'])
print(d['roofline']); print(d.get('full_fuser_fwd_bwd')); print(d.get('cpu_baseline',{}).get('value'))
r=json.load(open('gpurun_out/r2_bench_reference.json')); print('reference', r['value'], r['cpu_baseline']['kind'], r['cpu_baseline']['cores'])
PY
