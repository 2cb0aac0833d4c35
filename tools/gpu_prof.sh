#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.txt 2>&1; echo "rc=$?"; tail -5 gpurun_out/pytest_gpu.txt
echo "== ncu full (K1,K2,K3 at C3, 4 stacks)"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'(gaussian_scatter_kernel|simulate_kernel|superres_scatter_kernel)' -c 3 -o gpurun_out/prof_psf -f python tools/profile_c3.py 4 > gpurun_out/ncu_full.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out/*.ncu-rep
