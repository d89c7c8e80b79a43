#!/bin/bash
# usage: r2_scale.sh N   -- weak (B=64/GPU) and strong (global batch 512) scaling runs on N GPUs of one box
cd /root/repo
N=$1
mkdir -p gpurun_out
nvidia-smi -L | wc -l
run() {  # name, extra args
  name=$1; shift
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29520 \
      bench.py --gpus $N --steps 10 --warmup 3 "$@" > gpurun_out/r02_bench_n${N}_$name.json 2> gpurun_out/r02_bench_n${N}_$name.err
  echo "n$N $name rc=$?"
}
run weak
run strong512 --global-batch 512
python - <<PY
import json
for f in ('weak','strong512'):
    try:
        d=json.load(open(f'gpurun_out/r02_bench_n${N}_{f}.json'))
        print(f, {k:d.get(k) for k in ('value','ms_per_step','n_gpus','scaling')}, 'e2e', round(d['e2e']['value'],1), d.get('grad_allreduce',{}).get('ms'), d['config']['B_per_gpu'])
    except Exception as e:
        print(f, 'ERR', e); print(open(f'gpurun_out/r02_bench_n${N}_{f}.err').read()[-1200:])
PY
