#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -6
timeout 600 python bench.py --steps 5 --warmup 3 --no-yardstick --no-extras --no-cpu-baseline --opt jacobi_overlap_v=0 > gpurun_out/r2_bench_chain_nov.json 2> gpurun_out/r2_bench_chain_nov.err; echo "bench nov rc=$?"
timeout 600 python bench.py --steps 5 --warmup 3 --no-yardstick --no-extras --no-cpu-baseline > gpurun_out/r2_bench_chain.json 2> gpurun_out/r2_bench_chain.err; echo "bench rc=$?"
python - <<'PY'
import json
for n in ("chain_nov", "chain"):
    try:
        d = json.loads(open(f"gpurun_out/r2_bench_{n}.json").read().strip().splitlines()[-1])
        print(n, d["value"], d["ms_per_step"], d.get("erank_sweeps"), {k: (round(v["ms_per_step"], 2), round(v.get("avg_launch_us", 0), 1), v.get("launches_per_step")) for k, v in d.get("stages", {}).items()})
    except Exception as e:
        print(n, "failed", e)
        print(open(f"gpurun_out/r2_bench_{n}.err").read()[-2000:])
PY
