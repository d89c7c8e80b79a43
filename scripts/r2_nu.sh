#!/bin/bash
python scripts/dbg_floor_sweep.py 2>&1 | tail -12
for o in "jacobi_nu_pass1=1024 --opt jacobi_nu_pass2=-1e-10 --opt erank_pass2_sweeps=6" "jacobi_nu_pass1=2048 --opt jacobi_nu_pass2=-1e-10 --opt erank_pass2_sweeps=6" "jacobi_nu_pass1=512 --opt jacobi_nu_pass2=-1e-10 --opt erank_pass2_sweeps=6"; do
  timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --opt $o 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$o', round(d['ms_per_step'],2), d['config'].get('jacobi_sweeps_mean'), d['config'].get('jacobi_not_converged'))"
done
python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "token_axis or erank_golden or strict or threads or packed or empty" 2>&1 | tail -8
