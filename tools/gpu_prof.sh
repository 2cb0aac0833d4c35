#!/bin/bash
mkdir -p gpurun_out
for k in simulate_kernel superres_scatter_kernel gaussian_scatter_kernel gaussian_sume_kernel; do
echo "== ncu full $k (C3 geometry, 2 stacks)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -c 1 -o gpurun_out/prof_$k -f python tools/profile_c3.py 2 > gpurun_out/ncu_$k.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/ncu_$k.log
done
ls -la gpurun_out/*.ncu-rep
