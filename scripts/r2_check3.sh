#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "panel_round or chained or chunked or golden_fixtures" 2>&1 | tail -3
rm -f gpurun_out/r2_exp_*
bash scripts/r2_exp.sh pf "" pf_c1 "--opt jacobi_chunks=1"
