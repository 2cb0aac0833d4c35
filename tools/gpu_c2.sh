#!/bin/bash
# C2 (BASELINE.json configs[1]): bundled 4-stack 3T data, full iteration count: ours vs the reference's CUDA path, and
# the SVRreconstructionGPU CLI end to end from the .nii.gz files.
mkdir -p gpurun_out/c2
D=data_local/c2_setup
if [ "$1" != "regonly" ]; then
echo "== C2 parity, no registration (PSF forward/back-projection + EM only)"
timeout 900 python tools/c2_parity.py run ref $D /tmp/c2_ref.npz > gpurun_out/c2/ref.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/c2/ref.log
timeout 900 python tools/c2_parity.py run ref $D /tmp/c2_ref2.npz > gpurun_out/c2/ref2.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/c2/ref2.log
timeout 900 python tools/c2_parity.py run cuda $D /tmp/c2_cuda.npz > gpurun_out/c2/cuda.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/c2/cuda.log
echo "---- ours vs reference"; timeout 300 python tools/c2_parity.py cmp /tmp/c2_ref.npz /tmp/c2_cuda.npz gpurun_out/c2/parity_noreg.json 2>&1 | tail -20
echo "---- reference vs reference (second run)"; timeout 300 python tools/c2_parity.py cmp /tmp/c2_ref.npz /tmp/c2_ref2.npz gpurun_out/c2/selfcheck_noreg.json 2>&1 | tail -20
fi
echo "== C2 parity with GPU registration"
timeout 900 python tools/c2_parity.py run ref $D /tmp/c2r_ref.npz --register > gpurun_out/c2/ref_reg.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/c2/ref_reg.log
timeout 900 python tools/c2_parity.py run cuda $D /tmp/c2r_cuda.npz --register > gpurun_out/c2/cuda_reg.log 2>&1; echo "rc=$?"; tail -1 gpurun_out/c2/cuda_reg.log
echo "---- ours vs reference (registered)"; timeout 300 python tools/c2_parity.py cmp /tmp/c2r_ref.npz /tmp/c2r_cuda.npz gpurun_out/c2/parity_reg.json 2>&1 | tail -12
echo "== CLI end to end"
mkdir -p /tmp/c2cli && cd /tmp/c2cli && R=$GRAFT_REPO_ROOT && $R/host/SVRreconstructionGPU -o 3TReconstruction.nii.gz -i $R/data_local/14_3T_nody_001.nii.gz $R/data_local/10_3T_nody_001.nii.gz $R/data_local/21_3T_nody_001.nii.gz $R/data_local/23_3T_nody_001.nii.gz -m $R/data_local/mask_10_3T_brain_smooth.nii.gz --resolution 1.0 --useGPUReg > $R/gpurun_out/c2/cli.log 2>&1; echo "rc=$?"; tail -45 $R/gpurun_out/c2/cli.log | head -40; ls -la /tmp/c2cli | head -30; cp /tmp/c2cli/performance_GPU_*.txt /tmp/c2cli/log-evaluation.txt $R/gpurun_out/c2/ 2>/dev/null
cd $R; python - <<'PY'
import sys, numpy as np
sys.path.insert(0, '.')
# compare the CLI's final volume with the python-driven cuda arm (registered)
import gzip, struct
def read_nii(path):
    raw = gzip.open(path).read() if path.endswith('.gz') else open(path, 'rb').read()
    dim = struct.unpack('<8h', raw[40:56]); dt = struct.unpack('<h', raw[70:72])[0]; off = int(struct.unpack('<f', raw[108:112])[0])
    n = dim[1] * dim[2] * dim[3]
    return np.frombuffer(raw, {16: np.float32, 64: np.float64}[dt], n, off).reshape(dim[3], dim[2], dim[1])
v = read_nii('/tmp/c2cli/3TReconstruction.nii.gz')
c = np.load('/tmp/c2r_cuda.npz')['volume'].reshape(v.shape)
sc = np.sqrt(np.mean(c[c != 0] ** 2))
print('CLI volume', v.shape, 'finite', np.isfinite(v).all(), 'rms diff vs python-driven path (registered): %.3e rel' % (np.sqrt(np.mean((v - c) ** 2)) / sc))
PY
