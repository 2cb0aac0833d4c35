#!/bin/bash
# Two GPUs: the C++ host's -d 0 1 against one device, bench at N = 1 and N = 2.
mkdir -p gpurun_out
nvidia-smi -L
echo "== pytest gpu cli + pipeline"; timeout 1200 python -m pytest tests/test_gpu_cli.py tests/test_gpu_parity.py tests/test_ref_golden.py -m gpu -q -x > gpurun_out/pytest_gpu_n2.txt 2>&1; echo "rc=$?"; tail -12 gpurun_out/pytest_gpu_n2.txt
echo "== bench C3 N=1"; timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-reference-cuda --no-registration > gpurun_out/bench_c3_n1.json 2> gpurun_out/bench_c3_n1.err; echo "rc=$?"; tail -2 gpurun_out/bench_c3_n1.err
echo "== bench C3 N=2"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --no-registration > gpurun_out/bench_c3_n2.json 2> gpurun_out/bench_c3_n2.err; echo "rc=$?"; tail -2 gpurun_out/bench_c3_n2.err
echo "== bench C4 N=2"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload C4 --steps 2 --warmup 1 > gpurun_out/bench_c4_n2.json 2> gpurun_out/bench_c4_n2.err; echo "rc=$?"; tail -2 gpurun_out/bench_c4_n2.err
python - <<'PY'
import json
for f in ('bench_c3_n1','bench_c3_n2','bench_c4_n2'):
    try:
        d=json.loads([l for l in open(f'gpurun_out/{f}.json') if l.startswith('{')][-1])
        print(f, 'value',round(d['value'],1),'ms_per_step',round(d['ms_per_step'],2),'e2e',d['e2e'] and round(d['e2e']['value'],1),'replicas',d['replicas_identical'], 'launches', d['gpu_launches'])
        for k,v in d['roofline']['kernels'].items(): print('    ',k, round(v['ms_per_launch'],3), v['launches'])
    except Exception as e: print(f, 'failed', e)
PY
