set -x
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -2
timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/r01_bench_n1.json 2> gpurun_out/bench_n1.err; tail -c 300 gpurun_out/bench_n1.err
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r01_bench_reference.json 2>/dev/null
timeout 300 python scripts/erank_sweep.py > gpurun_out/erank_sweep.log 2>&1; tail -3 gpurun_out/erank_sweep.log
timeout 200 python scripts/gram_sweep.py gpurun_out/gram_sweep.json > gpurun_out/gram_sweep.log 2>&1
timeout 100 python scripts/stage_bench.py > gpurun_out/stage_bench.log 2>&1; cat gpurun_out/stage_bench.log
python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -s 3600 -c 1250 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
tail -2 gpurun_out/ncu_launch.log
