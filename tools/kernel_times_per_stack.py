"""Per-kernel device times at the C3 geometry, one stack at a time: the stacks differ only in orientation, so this shows
how the PSF kernels' time depends on how a slice's pixel rows run through the volume.  GPU-box tooling."""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fetalreconstruction_b200.phantom import make_dataset, c3_config, _STACK_ANGLES
from fetalreconstruction_b200.pipeline import SVRPipeline, SVRParams, upload_dataset
from fetalreconstruction_b200.reconstruction import Reconstruction

for st in range(8):
    ds = make_dataset(c3_config(), device="cuda", stacks=[st])
    b = Reconstruction(0)
    upload_dataset(b, ds)
    p = SVRPipeline(b, ds.S, 0, ds.S, params=SVRParams())
    p.InitializeEMGPU(ds.slices)
    p.outer_iteration(0)
    b.profile_reset(); b.profile_enable(True)
    p.outer_iteration(0)
    torch.cuda.synchronize()
    prof = b.profile_read()
    n_px = int(np.count_nonzero(b.debugv_PSF_sums()))
    a = ds.stack_attrs[0]
    print(json.dumps({"stack": st, "angles": list(_STACK_ANGLES[st]) if st < len(_STACK_ANGLES) else None,
                      "slice_x_axis_in_volume": [round(float(v), 3) for v in a.xaxis], "slice_y_axis": [round(float(v), 3) for v in a.yaxis],
                      "pixels": n_px, **{k: round(v[0] / max(v[1], 1), 3) for k, v in prof.items() if k in ("gaussian", "simulate", "superres")},
                      "simulate_ns_per_pixel": round(prof["simulate"][0] / max(prof["simulate"][1], 1) * 1e6 / max(n_px, 1), 3)}))
    b.close()
