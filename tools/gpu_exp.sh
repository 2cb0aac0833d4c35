#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/kernel_times.txt
for lib in "$@"; do
  SVR_B200_LIB=$PWD/fetalreconstruction_b200/csrc/$lib timeout 600 python tools/kernel_times.py >> gpurun_out/kernel_times.txt 2>> gpurun_out/kernel_times.err
done
cat gpurun_out/kernel_times.txt; tail -3 gpurun_out/kernel_times.err
