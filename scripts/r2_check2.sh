#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
rm -f gpurun_out/r2_exp_*
bash scripts/r2_exp.sh base "" 
