#!/usr/bin/env python
"""Packs the set-up dump of BASELINE.json configs[1] ("C2": the reference's bundled 4-stack 3T data at --resolution 1.0)
into tests/golden/c2_setup.npz (2.3 MB): what irtkReconstruction::SyncGPU hands to the device library after the
reference's set-up pipeline (mask crop, template, intensity matching, slice creation + masking) -- the padded slice cube,
the mask, the per-slice matrices and voxel sizes.  Derived DATA of the reference's data/*.nii.gz, no reference code.

    host/SVRreconstructionGPU -o out.nii.gz -i data/14_3T_nody_001.nii.gz data/10_3T_nody_001.nii.gz \
        data/21_3T_nody_001.nii.gz data/23_3T_nody_001.nii.gz -m data/mask_10_3T_brain_smooth.nii.gz \
        --resolution 1.0 --dump_setup DIR          # (GPU box; the command of bin/linux64/runTestDataset.sh + --dump_setup)
    python tests/golden/make_c2_setup.py DIR
"""
import os
import sys

import numpy as np

d = sys.argv[1]
arrs = {}
for f in sorted(os.listdir(d)):
    n, ext = os.path.splitext(f)
    if ext in (".f32", ".i32", ".f64"):
        arrs[n] = np.fromfile(os.path.join(d, f), {".f32": np.float32, ".i32": np.int32, ".f64": np.float64}[ext])
    elif ext == ".txt":
        arrs[n] = np.array(open(os.path.join(d, f)).read())
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "c2_setup.npz")
np.savez_compressed(out, **arrs)
print(out, os.path.getsize(out), "bytes")
