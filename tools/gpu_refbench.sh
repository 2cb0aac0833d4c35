#!/bin/bash
mkdir -p gpurun_out/ref
echo "== reg trace"
for r in ref ref2; do for i in 0 1 2 3 4 5 6; do timeout 300 python tools/ref_reg_trace.py ref gpurun_out/ref/regtrace_$r.npz $i > gpurun_out/ref/regtrace_$r.log 2>&1 || echo "fail $r $i"; done; done
timeout 600 python tools/ref_reg_trace.py cuda gpurun_out/ref/regtrace_cuda.npz > gpurun_out/ref/regtrace_cuda.log 2>&1; echo "rc=$?"
echo "== ref bench"; timeout 300 python tools/ref_bench.py gen --stacks ${1:-2} 2>&1 | tail -1
timeout 1500 python tools/ref_bench.py ref --rec-iters ${2:-2} --out /tmp/refbench_ref.npz > gpurun_out/refbench_ref.log 2>&1; echo "rc=$?"
timeout 1500 python tools/ref_bench.py ref --rec-iters ${2:-2} --out /tmp/refbench_ref2.npz > gpurun_out/refbench_ref2.log 2>&1; echo "rc=$?"
timeout 600 python tools/ref_bench.py cuda --rec-iters ${2:-2} --out /tmp/refbench_cuda.npz > gpurun_out/refbench_cuda.log 2>&1; echo "rc=$?"
for pair in "ref cuda refbench" "ref ref2 refbench_selfcheck"; do set -- $pair
timeout 600 python tools/ref_bench.py cmp /tmp/refbench_$1.npz /tmp/refbench_$2.npz gpurun_out/$3.json > gpurun_out/$3_cmp.log 2>&1; echo "rc=$?"; echo "---- $1 vs $2"; python - $3 <<'PY'
import json, sys
d = json.load(open('gpurun_out/%s.json' % sys.argv[1]))
for k, v in d['parity_cuda_vs_reference'].items(): print(k, {a: ('%.2e' % b if isinstance(b, float) else b) for a, b in v.items() if a != 'n'})
for k, v in d['times_ms_per_call'].items(): print(k, {a: round(b, 2) for a, b in v.items()})
print(d['outer_iteration_s'])
PY
done
