#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "gemm or block or golden_forward or embed or front or futr or multi or train_step" 2>&1 | tail -4
timeout 300 python scripts/gemm_bench.py > gpurun_out/gemm_bench.log 2>&1; tail -14 gpurun_out/gemm_bench.log | cut -c1-160
