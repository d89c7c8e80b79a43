#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -4
python __graft_entry__.py smoke 2>&1 | tail -2
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err; echo "ref rc=$?"
timeout 400 python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench rc=$?"
python scripts/gemm_bench.py > gpurun_out/gemm_bench.log 2>&1; tail -3 gpurun_out/gemm_bench.log
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_n1.json'))
print({k:d.get(k) for k in ('value','ms_per_step','gpu_launches','clocks')}, d['e2e']['value'], d.get('e2e_c_abi',{}).get('value'))
print(d['roofline']); print(d.get('full_fuser_fwd_bwd')); print(d.get('cpu_baseline',{}).get('value'), d.get('gpu_library_yardstick'))
for k,v in d['stages'].items():
    if v.get('ms_per_step',0)>0.3: print('   ',k, {a:(round(b,3) if isinstance(b,float) else b) for a,b in v.items() if a in ('ms_per_step','launches_per_step','avg_launch_us','frac')})
print({k:(round(v['avg_launch_us'],1), round(v['frac'],3)) for k,v in d.get('stages_isolated',{}).items() if 'frac' in v})
r=json.load(open('gpurun_out/r02_bench_reference.json')); print('reference', r['value'], r['cpu_baseline']['kind'], r['cpu_baseline']['cores'])
PY
