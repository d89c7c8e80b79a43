#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 100 python scripts/dbg_vchain.py 2>&1 | tail -2; timeout 100 python scripts/dbg_panel.py b 2>&1 | tail -3; timeout 200 python scripts/dbg_panel.py sym 2>&1 | grep -v "^  H" | tail -10
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
rm -f gpurun_out/r2_exp_*
bash scripts/r2_exp.sh rawhi "" rawhi_nov "--opt jacobi_overlap_v=0"
