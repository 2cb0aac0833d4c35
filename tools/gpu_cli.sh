#!/bin/bash
# SVRreconstructionGPU end to end on the bundled 3T data (C2); optional: library variant via SVR_LIB_VARIANT
mkdir -p gpurun_out/c2 /tmp/c2cli && cd /tmp/c2cli && R=$GRAFT_REPO_ROOT
$R/host/SVRreconstructionGPU -o 3TReconstruction.nii.gz -i $R/data_local/14_3T_nody_001.nii.gz $R/data_local/10_3T_nody_001.nii.gz $R/data_local/21_3T_nody_001.nii.gz $R/data_local/23_3T_nody_001.nii.gz -m $R/data_local/mask_10_3T_brain_smooth.nii.gz --resolution 1.0 --useGPUReg > $R/gpurun_out/c2/cli.log 2>&1; echo "rc=$?"
grep -E "^(GaussianReconstruction|SimulateSlices|Superresolution|Registration|EStep|MStep)" $R/gpurun_out/c2/cli.log; grep "overall" $R/gpurun_out/c2/cli.log
