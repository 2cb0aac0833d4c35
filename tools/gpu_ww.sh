#!/bin/bash
# Round 2: warp-window kernels: per-stack A/B of the variants, the GPU test-suite under each variant, ncu of the window kernels.
mkdir -p gpurun_out
echo "== ww_ab all stacks"; timeout 1500 python tools/ww_ab.py > gpurun_out/ww_ab.txt 2> gpurun_out/ww_ab.err; echo "rc=$?"; python - <<'PY'
import json
for l in open('gpurun_out/ww_ab.txt'):
    d=json.loads(l)
    print(d['stack'], d['slice_x_axis'], d['slice_y_axis'], 'K3', [d.get(f'K3_scatter{v}_ms') for v in range(4)], 'K1', [d.get(f'K1_scatter{v}_ms') for v in range(4)], 'K2', [d.get(f'K2_sim{v}_ms') for v in range(3)])
    for v in (1,2): print('   plan', v, d.get(f'K3_scatter{v}_plan'))
    print('   parity K3', {v: (d[f'K3_scatter{v}_vs0']['addon']['rms'], d[f'K3_scatter{v}_vs0']['addon']['max']) for v in (1,2,3)}, 'K1', {v: (d[f'K1_scatter{v}_vs0']['volume']['rms'], d[f'K1_scatter{v}_vs0']['volume']['max'], d[f'K1_scatter{v}_vs0']['voxel_num_diff']) for v in (1,2,3)}, 'K2', {v: (d[f'K2_sim{v}_vs0']['sim']['rms'], d[f'K2_sim{v}_vs0']['sim']['max'], d[f'K2_sim{v}_vs0']['inside_diff']) for v in (1,2)})
PY
tail -3 gpurun_out/ww_ab.err
echo "== pytest gpu (defaults: scatter=3 simulate=0)"; timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.txt 2>&1; echo "rc=$?"; tail -5 gpurun_out/pytest_gpu.txt
for cfg in "1 1" "2 2"; do set -- $cfg
echo "== pytest gpu parity (scatter=$1 simulate=$2)"; SVR_TUNE_SCATTER=$1 SVR_TUNE_SIMULATE=$2 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_ref_golden.py tests/test_gpu_fullsize.py tests/test_gpu_pvr.py -m gpu -q > gpurun_out/pytest_gpu_v$1$2.txt 2>&1; echo "rc=$?"; tail -12 gpurun_out/pytest_gpu_v$1$2.txt
done
for st in 0 7; do
echo "== ncu window_scatter_kernel stack $st"; SVR_TUNE_SCATTER=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:window_scatter_kernel -s 1 -c 1 -o gpurun_out/prof_window_scatter_s$st -f python tools/profile_c3.py stack=$st > gpurun_out/ncu_ws_$st.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/ncu_ws_$st.log
done
echo "== ncu window_simulate_kernel stack 7"; SVR_TUNE_SIMULATE=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:window_simulate_kernel -s 1 -c 1 -o gpurun_out/prof_window_simulate_s7 -f python tools/profile_c3.py stack=7 > gpurun_out/ncu_wsim_7.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/ncu_wsim_7.log
echo "== bench C3 (default tuning)"; timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-reference-cuda --no-registration > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; echo "rc=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_c3.json'))
print('value',d['value'],'ms_per_step',d['ms_per_step'],'vph',d['volumes_per_hour'],'e2e',d['e2e']['value'] if d['e2e'] else None)
for k,v in d['roofline']['kernels'].items(): print(k, v)
print(d['clocks'])
PY
tail -3 gpurun_out/bench_c3.err
