"""One pass over the hot kernels at the C3 size, for ncu (no timing here; never report numbers taken under a profiler).
   python tools/profile_c3.py [n_stacks]          the first n stacks
   python tools/profile_c3.py stack=K             stack K alone (e.g. 2 = the through-plane stack of the phantom)"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fetalreconstruction_b200.phantom import make_dataset, c3_config
from fetalreconstruction_b200.pipeline import SVRPipeline, SVRParams, upload_dataset
from fetalreconstruction_b200.reconstruction import Reconstruction

arg = sys.argv[1] if len(sys.argv) > 1 else "8"
cfg = c3_config()
ds = make_dataset(cfg, device="cuda", stacks=[int(arg[6:])] if arg.startswith("stack=") else range(int(arg)))
b = Reconstruction(0)
upload_dataset(b, ds)
p = SVRPipeline(b, ds.S, 0, ds.S, params=SVRParams())
p.InitializeEMGPU(ds.slices)
p.set_schedule(0)
p.InitializeEMValuesGPU()
p.GaussianReconstructionGPU()
p.SimulateSlicesGPU()
p.InitializeRobustStatisticsGPU()
p.EStepGPU()
p.reconstruction_iteration(0)
print("n_valid fraction", float((ds.slices != -1).mean()), "S", ds.S)
