"""Per-kernel device times (CUDA events inside the library) at the C3 size for the library selected by
$SVR_B200_LIB.  Used to compare build variants on the GPU box."""
import os, sys, json, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fetalreconstruction_b200.phantom import make_dataset, c3_config
from fetalreconstruction_b200.pipeline import SVRPipeline, SVRParams, upload_dataset
from fetalreconstruction_b200.reconstruction import Reconstruction

cache = "/tmp/c3_ds.pt"
if os.path.exists(cache):
    ds = torch.load(cache, weights_only=False)
else:
    ds = make_dataset(c3_config(), device="cuda")
    torch.save(ds, cache)
b = Reconstruction(0)
upload_dataset(b, ds)
p = SVRPipeline(b, ds.S, 0, ds.S, params=SVRParams())
p.InitializeEMGPU(ds.slices)
p.outer_iteration(0)          # warm-up
b.profile_reset(); b.profile_enable(True)
t = time.perf_counter()
p.outer_iteration(0)
torch.cuda.synchronize()
wall = time.perf_counter() - t
prof = b.profile_read()
print(json.dumps({"lib": os.path.basename(os.environ.get("SVR_B200_LIB", "default")), "step_s": round(wall, 4),
                  **{k: round(v[0] / max(v[1], 1), 3) for k, v in prof.items()}}))
