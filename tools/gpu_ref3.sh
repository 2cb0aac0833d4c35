#!/bin/bash
# three-way parity against the committed reference goldens + reference-vs-ours at C3 geometry
mkdir -p gpurun_out
echo "== three-way parity"; timeout 900 python tools/ref_parity.py tests/golden gpurun_out/ref_parity.json 2> gpurun_out/ref_parity.err > gpurun_out/ref_parity.txt; grep "^reg" gpurun_out/ref_parity.txt; tail -3 gpurun_out/ref_parity.err
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.txt 2>&1; echo "rc=$?"; tail -15 gpurun_out/pytest_gpu.txt
echo "== ref bench"; timeout 300 python tools/ref_bench.py gen --stacks ${1:-2} 2>&1 | tail -1
timeout 1500 python tools/ref_bench.py ref --rec-iters ${2:-2} --out gpurun_out/refbench_ref.npz > gpurun_out/refbench_ref.log 2>&1; echo "rc=$?"; grep -A8 "^ref S" gpurun_out/refbench_ref.log
timeout 600 python tools/ref_bench.py cuda --rec-iters ${2:-2} --out gpurun_out/refbench_cuda.npz > gpurun_out/refbench_cuda.log 2>&1; echo "rc=$?"; grep -A8 "^cuda S" gpurun_out/refbench_cuda.log
timeout 600 python tools/ref_bench.py cmp gpurun_out/refbench_ref.npz gpurun_out/refbench_cuda.npz gpurun_out/refbench.json > gpurun_out/refbench_cmp.log 2>&1; echo "rc=$?"; python - <<'PY'
import json
d = json.load(open('gpurun_out/refbench.json'))
for k, v in d['parity_cuda_vs_reference'].items(): print(k, {a: ('%.2e' % b if isinstance(b, float) else b) for a, b in v.items()})
for k, v in d['times_ms_per_call'].items(): print(k, v)
print(d['outer_iteration_s'])
PY
rm -f gpurun_out/refbench_ref.npz gpurun_out/refbench_cuda.npz
