#!/usr/bin/env python
"""Registration optimiser trace on the golden registration case for a list of shortened schedules (levels, steps,
iterations): final transforms + evaluation counts per schedule, for the reference (own process), ours and the oracle.
Test tooling.   python tools/ref_reg_trace.py ref|cuda|oracle OUT.npz [schedule-index]
The reference keeps process-global state (legacy texture references, cudaDeviceReset in its constructor): one process per
schedule for the `ref` arm (pass the schedule index; the outputs are merged by the caller)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import ref_runner  # noqa: E402

SCHEDULES = [(1, 1, 1), (1, 1, 2), (1, 1, 4), (1, 2, 1), (1, 2, 4), (2, 1, 1), (2, 2, 4)]


def main():
    arm, out_path = sys.argv[1], sys.argv[2]
    only = int(sys.argv[3]) if len(sys.argv) > 3 else None
    mg = ref_runner._mg()
    out = dict(np.load(out_path)) if only is not None and os.path.exists(out_path) else {}
    for sched in (SCHEDULES if only is None else [SCHEDULES[only]]):
        if arm == "ref":
            from oracle.ref_backend import RefReconstruction
            b = RefReconstruction(0)
        elif arm == "cuda":
            from fetalreconstruction_b200.reconstruction import Reconstruction
            b = Reconstruction(0)
        else:
            b = None
        orig = None
        if b is not None:
            orig = b.setRegSchedule
            b.setRegSchedule = lambda *a, _o=orig, _s=sched: _o(*_s)
            d = mg.reg_case(b)
        else:
            from oracle.oracle_backend import OracleReconstruction

            class O(OracleReconstruction):
                def setRegSchedule(self, *a):
                    super().setRegSchedule(*sched)
            # reg_case(None) builds its own oracle backend; go through the non-oracle branch with recon_w2i preset
            d = reg_case_oracle(mg, O())
        tag = "%d%d%d" % sched
        out["T_" + tag] = d["transforms_out"]
        out["ev_" + tag] = d["evaluations"]
        out["sim0"] = d["sim_level0"]
    np.savez(out_path, **out)
    print("wrote", out_path)


def reg_case_oracle(mg, backend):
    # the oracle branch of reg_case sets recon_w2i directly instead of the slice set-up calls
    import types
    backend.setMask = lambda *a, **k: None
    backend.initStorageVolumes = lambda *a, **k: None
    backend.setSliceDims = lambda *a, **k: None

    def ssm(T, Tinv, a3, a4, a5, a6, ri2w, rw2i):
        backend.recon_w2i = np.asarray(rw2i, np.float32)
    backend.SetSliceMatrices = ssm
    return mg.reg_case(backend)


if __name__ == "__main__":
    main()
