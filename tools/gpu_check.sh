#!/bin/bash
# GPU check: parity report, gpu tests, the C3 bench, and an ncu launch list restricted to our kernels.
mkdir -p gpurun_out
KREGEX='regex:(scatter_kernel|sume_kernel|simulate_kernel|reg_|estep|mstep|scale_|robust|pack_volume|equalize|init_em|mask_|build_geom|fold_|potential|DeviceSelect|DeviceCompact)'
echo "== parity report"; timeout 600 python tools/parity_report.py > gpurun_out/parity_report.txt 2> gpurun_out/parity_report.err; echo "rc=$?"; tail -3 gpurun_out/parity_report.err
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.txt 2>&1; echo "rc=$?"; tail -15 gpurun_out/pytest_gpu.txt
echo "== bench C3"; timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; echo "rc=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_c3.json'))
print('value',d['value'],'ms_per_step',d['ms_per_step'],'vph',d['volumes_per_hour'],'e2e',d['e2e']['value'] if d['e2e'] else None)
for k,v in d['roofline']['kernels'].items(): print(k, v)
print(d['cpu_baseline']); print(d['clocks'])
PY
tail -3 gpurun_out/bench_c3.err
if [ "$1" == "ncu" ]; then
echo "== ncu launch list (C3, one step)"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREGEX" -c 200 --csv --log-file gpurun_out/launches_c3.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --no-reference-cuda --no-registration > gpurun_out/ncu_c3.log 2>&1; echo "rc=$?"
fi
