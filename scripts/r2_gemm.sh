#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "gemm or golden_forward or train_step or embed or front or multi_modality or host_entry or torch_ops or futr" 2>&1 | tail -6
python scripts/gemm_bench.py 2>&1 | tail -13
ROWS=6 python scripts/prof_full_fuser.py 2>&1 | tail -14
