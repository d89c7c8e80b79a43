#!/bin/bash
# 2-GPU call: NCCL parity tests + weak / strong scaling
cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_parity.py -x -q -m gpu -k "multi or two_devices or dataparallel or nccl" 2>&1 | tail -4
bash scripts/r2_scale.sh 2
