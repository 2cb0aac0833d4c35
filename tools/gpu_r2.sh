#!/bin/bash
# Round 2 check: GPU tests, the three bench workloads, the CPU arm.
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.txt 2>&1; echo "rc=$?"; tail -15 gpurun_out/pytest_gpu.txt
echo "== pytest gpu, two-kernel regulariser"; SVR_TUNE_REGULARIZE=0 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_ref_golden.py -m gpu -q > gpurun_out/pytest_gpu_reg0.txt 2>&1; echo "rc=$?"; tail -3 gpurun_out/pytest_gpu_reg0.txt
for wl in C2 C4; do
echo "== bench $wl"; timeout 1200 python bench.py --workload $wl --steps 3 --warmup 3 > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.err; echo "rc=$?"; tail -3 gpurun_out/bench_$wl.err; python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_$wl.json'))
    print('value',d['value'],d['unit'],'ms_per_step',d['ms_per_step'],'e2e',d['e2e'])
    for k,v in d['roofline']['kernels'].items(): print('  ',k, {a:(round(b,3) if isinstance(b,float) else b) for a,b in v.items()})
    print('  cpu', d.get('cpu_baseline') and {k:v for k,v in d['cpu_baseline'].items() if k!='sample'}, d['clocks'], d.get('patches'), d.get('patch_enumeration_s'))
except Exception as e: print('parse failed', e)
PY
done
echo "== bench C3 (full default line)"; timeout 1500 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; echo "rc=$?"; tail -3 gpurun_out/bench_c3.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_c3.json'))
print('value',d['value'],'ms_per_step',d['ms_per_step'],'vph',d['volumes_per_hour'],'e2e',d['e2e']['value'] if d['e2e'] else None)
for k,v in d['roofline']['kernels'].items(): print('  ',k, {a:(round(b,3) if isinstance(b,float) else b) for a,b in v.items()})
print('cpu', {k:v for k,v in (d['cpu_baseline'] or {}).items() if k!='sample'})
rc=d['reference_cuda']; print('refcuda', {k:v for k,v in (rc or {}).items() if k not in ('sample',)})
print('ours per stack', d.get('ours_per_stack'))
print(d['clocks'])
PY
echo "== bench reference arm C3"; timeout 900 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref_c3.json 2> gpurun_out/bench_ref_c3.err; echo "rc=$?"; cut -c1-400 gpurun_out/bench_ref_c3.json; tail -2 gpurun_out/bench_ref_c3.err
