#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "panel_round or chained" 2>&1 | tail -4
timeout 150 python bench.py --steps 5 --warmup 3 --no-yardstick --no-extras --no-cpu-baseline --opt jacobi_overlap_v=0 > gpurun_out/r2_bench_ps_nov.json 2> gpurun_out/r2_bench_ps_nov.err; echo "bench nov rc=$?"
timeout 150 python bench.py --steps 5 --warmup 3 --no-yardstick --no-extras --no-cpu-baseline > gpurun_out/r2_bench_ps.json 2> gpurun_out/r2_bench_ps.err; echo "bench rc=$?"
python - <<'PY'
import json
for n in ("ps_nov", "ps"):
    try:
        d = json.loads(open(f"gpurun_out/r2_bench_{n}.json").read().strip().splitlines()[-1])
        print(n, d["value"], d["ms_per_step"], {k: (round(v["ms_per_step"], 2), round(v.get("avg_launch_us", 0), 1), v.get("launches_per_step")) for k, v in d.get("stages", {}).items() if k.startswith("jacobi")})
    except Exception as e:
        print(n, "failed", e)
        print(open(f"gpurun_out/r2_bench_{n}.err").read()[-2000:])
PY
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
