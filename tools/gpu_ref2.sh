#!/bin/bash
# ref svr (vol multiple of 8) + reg intermediates + reference-vs-ours at C3 geometry
mkdir -p gpurun_out/ref
echo "== ref_runner svr"; timeout 600 python -m oracle.ref_runner --out gpurun_out/ref --cases svr > gpurun_out/ref/run_svr.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/ref/run_svr.log
echo "== reg debug"; timeout 600 python tools/ref_reg_debug.py gpurun_out/ref/ref_reg_debug.npz > gpurun_out/ref/run_regdbg.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/ref/run_regdbg.log
echo "== three-way parity (svr)"; timeout 900 python tools/ref_parity.py gpurun_out/ref gpurun_out/ref_parity.json 2> gpurun_out/ref_parity.err | grep "^svr"; tail -3 gpurun_out/ref_parity.err
echo "== ref bench"; timeout 300 python tools/ref_bench.py gen --stacks ${1:-2} 2>&1 | tail -1
timeout 1500 python tools/ref_bench.py ref --out gpurun_out/refbench_ref.npz > gpurun_out/refbench_ref.log 2>&1; echo "rc=$?"; grep -A12 "^ref S" gpurun_out/refbench_ref.log
timeout 600 python tools/ref_bench.py cuda --out gpurun_out/refbench_cuda.npz > gpurun_out/refbench_cuda.log 2>&1; echo "rc=$?"; grep -A12 "^cuda S" gpurun_out/refbench_cuda.log
timeout 600 python tools/ref_bench.py cmp gpurun_out/refbench_ref.npz gpurun_out/refbench_cuda.npz gpurun_out/refbench.json > gpurun_out/refbench_cmp.log 2>&1; echo "rc=$?"; cat gpurun_out/refbench_cmp.log | head -80
rm -f gpurun_out/refbench_ref.npz gpurun_out/refbench_cuda.npz
