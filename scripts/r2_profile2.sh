#!/bin/bash
# launch list of exactly the bench steps (no extras): 3 warm-up + 1 timed + 1 with stage events + 3 e2e steps
cd /root/repo
mkdir -p gpurun_out
BENCH="python bench.py --steps 1 --warmup 3 --no-extras --no-cpu-baseline --no-yardstick"
timeout 300 $BENCH > gpurun_out/plain.log 2>&1 && \
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_all.csv $BENCH > gpurun_out/ncu_launch.log 2>&1
echo "launch list rc=$?"; tail -c 400 gpurun_out/ncu_launch.log; wc -l gpurun_out/r02_launches_all.csv
