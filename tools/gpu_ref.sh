#!/bin/bash
# Runs the reference's own CUDA path (oracle/_ref) on the golden cases and the three-way parity report.
mkdir -p gpurun_out/ref
nvidia-smi --query-gpu=name,driver_version --format=csv,noheader
for c in steps svr reg; do
  echo "== ref_runner $c"; timeout 600 python -m oracle.ref_runner --out gpurun_out/ref --cases $c > gpurun_out/ref/run_$c.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/ref/run_$c.log
done
echo "== three-way parity"; timeout 900 python tools/ref_parity.py gpurun_out/ref gpurun_out/ref_parity.json > gpurun_out/ref_parity.txt 2> gpurun_out/ref_parity.err; echo "rc=$?"; cat gpurun_out/ref_parity.txt; tail -5 gpurun_out/ref_parity.err
