#!/bin/bash
# First-contact GPU script: parity report, gpu tests, a tiny bench, the C3 bench, and an ncu launch list.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu_info.txt 2>&1
nproc >> gpurun_out/gpu_info.txt
echo "== parity report"; timeout 600 python tools/parity_report.py > gpurun_out/parity_report.txt 2> gpurun_out/parity_report.err; echo "rc=$?"; tail -5 gpurun_out/parity_report.err
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.txt 2>&1; echo "rc=$?"; tail -15 gpurun_out/pytest_gpu.txt
echo "== bench tiny"; timeout 300 python bench.py --workload tiny --steps 2 --warmup 1 > gpurun_out/bench_tiny.json 2> gpurun_out/bench_tiny.err; echo "rc=$?"; cat gpurun_out/bench_tiny.json | cut -c1-1500; tail -3 gpurun_out/bench_tiny.err
echo "== bench C3"; timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; echo "rc=$?"; cat gpurun_out/bench_c3.json | cut -c1-3000; tail -3 gpurun_out/bench_c3.err
echo "== ncu launch list (tiny)"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_tiny.csv python bench.py --workload tiny --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_tiny.log 2>&1; echo "rc=$?"
