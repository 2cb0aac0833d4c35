#!/bin/bash
mkdir -p gpurun_out/ref
echo "== reg trace"
timeout 600 python tools/ref_reg_trace.py ref gpurun_out/ref/regtrace_ref.npz > gpurun_out/ref/regtrace_ref.log 2>&1; echo "rc=$?"
timeout 600 python tools/ref_reg_trace.py ref gpurun_out/ref/regtrace_ref2.npz > gpurun_out/ref/regtrace_ref2.log 2>&1; echo "rc=$?"
timeout 600 python tools/ref_reg_trace.py cuda gpurun_out/ref/regtrace_cuda.npz > gpurun_out/ref/regtrace_cuda.log 2>&1; echo "rc=$?"
echo "== ref bench"; timeout 300 python tools/ref_bench.py gen --stacks ${1:-2} 2>&1 | tail -1
timeout 1500 python tools/ref_bench.py ref --rec-iters ${2:-2} --out /tmp/refbench_ref.npz > gpurun_out/refbench_ref.log 2>&1; echo "rc=$?"; grep -A6 "^ref S" gpurun_out/refbench_ref.log
timeout 600 python tools/ref_bench.py cuda --rec-iters ${2:-2} --out /tmp/refbench_cuda.npz > gpurun_out/refbench_cuda.log 2>&1; echo "rc=$?"; grep -A6 "^cuda S" gpurun_out/refbench_cuda.log
timeout 600 python tools/ref_bench.py cmp /tmp/refbench_ref.npz /tmp/refbench_cuda.npz gpurun_out/refbench.json > gpurun_out/refbench_cmp.log 2>&1; echo "rc=$?"; python - <<'PY'
import json
d = json.load(open('gpurun_out/refbench.json'))
for k, v in d['parity_cuda_vs_reference'].items(): print(k, {a: ('%.2e' % b if isinstance(b, float) else b) for a, b in v.items()})
print(json.dumps(d['values'], indent=0))
for k, v in d['times_ms_per_call'].items(): print(k, v)
print(d['outer_iteration_s'])
PY
