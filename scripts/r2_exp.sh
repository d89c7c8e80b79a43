#!/bin/bash
# usage: r2_exp.sh name "--opt a=b --opt c=d" ...  (pairs) -- short bench runs with option sets
cd /root/repo
mkdir -p gpurun_out
while [ $# -gt 1 ]; do
  name=$1; opts=$2; shift 2
  timeout 150 python bench.py --steps 5 --warmup 3 --no-yardstick --no-extras --no-cpu-baseline $opts > gpurun_out/r2_exp_$name.json 2> gpurun_out/r2_exp_$name.err; echo "$name rc=$?"
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2_exp_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("r2_exp_")[1][:-5], round(d["value"], 1), round(d["ms_per_step"], 3), d["config"].get("jacobi_sweeps_mean"), {k: round(v["ms_per_step"], 2) for k, v in d.get("stages", {}).items() if v["ms_per_step"] > 0.3})
    except Exception as e:
        print(f, "failed", e)
PY
