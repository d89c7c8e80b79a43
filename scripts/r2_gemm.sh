#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "multi_modality or embed or front" 2>&1 | tail -12
