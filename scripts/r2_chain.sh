#!/bin/bash
# chained V update: kernel check, alternative-path tests, bench with jacobi_schedule 0 / 2
cd /root/repo
mkdir -p gpurun_out
timeout 300 python scripts/dbg_vchain.py > gpurun_out/r2_vchain.log 2>&1; echo "vchain rc=$?"; tail -12 gpurun_out/r2_vchain.log
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "alternative or chained" 2>&1 | tail -8
timeout 600 python bench.py --steps 5 --warmup 3 --no-yardstick --no-extras --no-cpu-baseline --opt jacobi_schedule=2 > gpurun_out/r2_bench_chain.json 2> gpurun_out/r2_bench_chain.err; echo "bench chain rc=$?"
timeout 600 python bench.py --steps 5 --warmup 3 --no-yardstick --no-extras --no-cpu-baseline > gpurun_out/r2_bench_base.json 2> gpurun_out/r2_bench_base.err; echo "bench base rc=$?"
python - <<'PY'
import json
for n in ("chain", "base"):
    try:
        d = json.loads(open(f"gpurun_out/r2_bench_{n}.json").read().strip().splitlines()[-1])
        print(n, d["value"], d["ms_per_step"], {k: (v["ms_per_step"], v.get("avg_launch_us")) for k, v in d.get("stages", {}).items() if k.startswith("jacobi")})
    except Exception as e:
        print(n, "failed", e)
        print(open(f"gpurun_out/r2_bench_{n}.err").read()[-2000:])
PY
