#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
K=${1:-panel_vchain_kernel}
SKIP=${2:-10}
export R3D_OPTS=jacobi_chunks=1
python scripts/prof_erank.py > gpurun_out/plain_chain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:$K -s $SKIP -c 1 -f -o gpurun_out/r02_prof_$K python scripts/prof_erank.py > gpurun_out/ncu_$K.log 2>&1
echo "$K rc=$?"; tail -3 gpurun_out/ncu_$K.log
ls -la gpurun_out/*.ncu-rep
