#!/bin/bash
# Round-2 ncu evidence: launch list of one bench step + --set full captures of the dominant kernels.
mkdir -p gpurun_out
BENCH="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-yardstick"
$BENCH > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 5600 -c 2600 --csv --log-file gpurun_out/r02_launches.csv $BENCH > gpurun_out/ncu_launch.log 2>&1
echo "launch list rc=$?"; tail -2 gpurun_out/ncu_launch.log
python scripts/prof_erank.py > gpurun_out/plain2.log 2>&1 && \
for k in panel_sym_kernel panel_update_tc_kernel jacobi_inner_cross_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 40 -c 1 -f -o gpurun_out/r02_prof_$k python scripts/prof_erank.py > gpurun_out/ncu_$k.log 2>&1
  echo "$k rc=$?"
done
python scripts/prof_block.py > gpurun_out/plain3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:lin_kernel -c 12 -f -o gpurun_out/r02_prof_lin python scripts/prof_block.py > gpurun_out/ncu_lin.log 2>&1
echo "lin rc=$?"
ls -la gpurun_out/*.ncu-rep
