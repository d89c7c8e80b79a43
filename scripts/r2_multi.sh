#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | head -4
timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_parity.py -q -m gpu -k "nccl or futr or two_devices or dataparallel" 2>&1 | tail -8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2_bench_n2_weak.json 2> gpurun_out/r2_bench_n2_weak.err; echo "n2 weak rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 --global-batch 128 > gpurun_out/r2_bench_n2_strong.json 2> gpurun_out/r2_bench_n2_strong.err; echo "n2 strong rc=$?"
python - <<'PY'
import json
for f in ('r2_bench_n2_weak','r2_bench_n2_strong'):
    try:
        d=json.load(open(f'gpurun_out/{f}.json'))
        print(f, {k:d.get(k) for k in ('value','ms_per_step','n_gpus','scaling')}, d.get('grad_allreduce'), d['e2e']['value'])
    except Exception as e:
        print(f, 'ERR', e); print(open(f'gpurun_out/{f}.err').read()[-1500:])
PY
