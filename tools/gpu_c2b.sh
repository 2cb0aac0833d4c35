#!/bin/bash
# C2 with the GPU registration in the loop: reference x2, ours x2; pairwise volume statistics and registration differences (mm)
mkdir -p gpurun_out/c2
D=data_local/c2_setup
n=0
for arm in ref ref cuda cuda; do n=$((n+1)); timeout 900 python tools/c2_parity.py run $arm $D /tmp/c2r_$n.npz --register > gpurun_out/c2/reg_$n.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/c2/reg_$n.log; done
for pair in "1 3 parity_reg ours-vs-reference" "1 2 selfcheck_ref_reg reference-vs-reference" "3 4 selfcheck_cuda_reg ours-vs-ours"; do set -- $pair
echo "---- $4 (registered)"; timeout 300 python tools/c2_parity.py cmp /tmp/c2r_$1.npz /tmp/c2r_$2.npz gpurun_out/c2/$3.json --setup $D 2>&1 | grep -E "^volume |^TRE|slice weights"
done
