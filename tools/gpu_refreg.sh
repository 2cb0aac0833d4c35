#!/bin/bash
# registration: ours vs the reference's CUDA code on the same GPU (tools/ref_bench_reg.py)
mkdir -p gpurun_out
timeout 600 python tools/ref_bench_reg.py ours gpurun_out/refbench_reg_ours.json 2> gpurun_out/refbench_reg_ours.err | tail -1
timeout 480 python tools/ref_bench_reg.py ref gpurun_out/refbench_reg_ref.json 2> gpurun_out/refbench_reg_ref.err | tail -1
tail -n 3 gpurun_out/refbench_reg_ours.err; tail -n 3 gpurun_out/refbench_reg_ref.err
