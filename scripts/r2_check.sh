#!/bin/bash
# Round-2 GPU check: erank-related parity tests, then short benches with a few knobs.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "erank or jacobi or altern or graph or packed or empty" > gpurun_out/r2_pytest_erank.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_pytest_erank.log
tail -8 gpurun_out/r2_pytest_erank.log
run() {
  name=$1; shift
  timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/r2_bench_$name.json 2> gpurun_out/r2_bench_$name.err
  echo "bench $name rc=$?"; python - "$name" <<'PY'
import json,sys
d=json.load(open(f'gpurun_out/r2_bench_{sys.argv[1]}.json'))
print({k:d.get(k) for k in ('value','ms_per_step','gpu_launches')}, d['config'].get('jacobi_sweeps_mean'))
for k,v in d['stages'].items():
    if v.get('ms_per_step',0)>0.5: print('   ',k, {a:(round(b,3) if isinstance(b,float) else b) for a,b in v.items() if a in ('ms_per_step','launches_per_step','avg_launch_us','frac')})
PY
}
run default
run vafterg --opt jacobi_v_after_g=1
run noverlap --opt jacobi_overlap_v=0
run spread --opt jacobi_schedule=1
