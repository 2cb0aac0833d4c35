"""Shared builder of a small PVR case (stacks -> patches) for the CPU and GPU PVR tests (lives in the package since bench.py's
C4 workload uses it too)."""
from fetalreconstruction_b200.pvr_case import make_pvr_case, setup_backend, shard_case  # noqa: F401
