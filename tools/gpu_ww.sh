#!/bin/bash
# Round 2: warp-window kernels: quick sanity on one stack, the GPU test-suite under the variants, per-stack A/B, short bench.
mkdir -p gpurun_out
echo "== ww_ab stack 0"; timeout 600 python tools/ww_ab.py 0 > gpurun_out/ww_ab_0.txt 2> gpurun_out/ww_ab_0.err; echo "rc=$?"; cat gpurun_out/ww_ab_0.txt; tail -5 gpurun_out/ww_ab_0.err
echo "== pytest gpu (defaults: scatter=2 simulate=0)"; timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.txt 2>&1; echo "rc=$?"; tail -8 gpurun_out/pytest_gpu.txt
echo "== pytest gpu parity (scatter=1 simulate=1)"; SVR_TUNE_SCATTER=1 SVR_TUNE_SIMULATE=1 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_ref_golden.py tests/test_gpu_fullsize.py tests/test_gpu_pvr.py -m gpu -q > gpurun_out/pytest_gpu_v11.txt 2>&1; echo "rc=$?"; tail -8 gpurun_out/pytest_gpu_v11.txt
echo "== ww_ab all stacks"; timeout 1500 python tools/ww_ab.py 1 2 3 4 5 6 7 > gpurun_out/ww_ab.txt 2> gpurun_out/ww_ab.err; echo "rc=$?"; cat gpurun_out/ww_ab.txt; tail -3 gpurun_out/ww_ab.err
echo "== bench C3 (default tuning)"; timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-reference-cuda --no-registration > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; echo "rc=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_c3.json'))
print('value',d['value'],'ms_per_step',d['ms_per_step'],'vph',d['volumes_per_hour'],'e2e',d['e2e']['value'] if d['e2e'] else None)
for k,v in d['roofline']['kernels'].items(): print(k, v)
print(d['clocks'])
PY
tail -3 gpurun_out/bench_c3.err
