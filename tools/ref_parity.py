#!/usr/bin/env python
"""Three-way parity report: reference CUDA path (outputs written by oracle/ref_runner.py) vs our CUDA path vs the
CPU oracle, on the seeded golden cases.  Test tooling; run on a GPU box:

    python -m oracle.ref_runner --out gpurun_out/ref && python tools/ref_parity.py gpurun_out/ref gpurun_out/ref_parity.json
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import ref_runner  # noqa: E402


def stats(a, b):
    a = np.asarray(a, np.float64).ravel(); b = np.asarray(b, np.float64).ravel()
    if a.shape != b.shape:
        return {"shape_mismatch": [list(a.shape), list(b.shape)]}
    nz = b[(b != 0) & np.isfinite(b)]
    scale = float(np.sqrt(np.mean(nz ** 2))) if nz.size else 1.0
    fin = np.isfinite(a) & np.isfinite(b)
    d = np.abs(a[fin] - b[fin]) / scale
    return {"rms": float(np.sqrt(np.mean(d ** 2))) if d.size else 0.0, "max": float(d.max()) if d.size else 0.0,
            "p999": float(np.quantile(d, 0.999)) if d.size else 0.0, "nonfinite_a": int((~np.isfinite(a)).sum()),
            "nonfinite_ref": int((~np.isfinite(b)).sum()), "n": int(a.size), "scale": scale}


def main():
    ref_dir, out_path = sys.argv[1], sys.argv[2]
    use_gpu = "--no-gpu" not in sys.argv
    mg = ref_runner._mg()
    from oracle.oracle_backend import OracleReconstruction
    report = {}
    cases = {"svr": lambda b: mg.svr_case(b, slice_size=mg.REF_SVR_SIZE), "reg": mg.reg_case,
             "steps": lambda b: ref_runner.steps_case(b or OracleReconstruction())}
    for case, fn in cases.items():
        path = os.path.join(ref_dir, f"ref_{case}_small.npz")
        if not os.path.exists(path):
            report[case] = "missing"
            continue
        ref = dict(np.load(path))
        arms = {}
        arms["oracle"] = fn(None)
        if use_gpu:
            from fetalreconstruction_b200.reconstruction import Reconstruction
            arms["cuda"] = fn(Reconstruction(0))
        rep = {}
        for k, v in ref.items():
            if k in ("config", "inplane", "spacing", "transforms_in", "evaluations", "m", "sigma"):
                continue
            rep[k] = {arm: stats(out[k], v) for arm, out in arms.items() if k in out}
            if "cuda" in arms and k in arms["cuda"] and k in arms["oracle"]:
                rep[k]["cuda_vs_oracle"] = stats(arms["cuda"][k], arms["oracle"][k])
        report[case] = rep
    with open(out_path, "w") as f:
        json.dump(report, f, indent=1)
    for case, rep in report.items():
        if not isinstance(rep, dict):
            print(case, rep); continue
        for k, r in rep.items():
            print(f"{case:6s}{k:18s}" + "  ".join(f"{arm}: rms {s.get('rms', -1):.2e} max {s.get('max', -1):.2e}" for arm, s in r.items()))


if __name__ == "__main__":
    main()
