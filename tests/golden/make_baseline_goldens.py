"""Fixtures for the BASELINE configs at their OWN size (tests/test_gpu_baseline_configs.py).

The reference ships no golden value at these sizes; the numbers below are outputs of the CPU
oracle (oracle/ed_oracle.c: ora_stored_lanczos_gs = pass 1 of sp_lanc_eigh on the stored
ED_SPARSE_H=T operator), which is itself pinned to the reference's test/src/*.check files.  They
are cached here because the Ns=16 recurrence takes minutes on 8 host cores; the GPU test can
re-run them live with EDGPU_LIVE_ORACLE_LANCZOS=1.

    python tests/golden/make_baseline_goldens.py          # writes tests/golden/baseline_cfg2.json
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import edipack_oracle as O  # noqa: E402
from models import star_kwargs  # noqa: E402


def cfg2(nthreads):
    """BASELINE config 2: Norb=1, Nbath=15 (Ns=16), sector (8,8), 165 636 900 states; start vector
    = start_vector(dim, 4321) (the seeded start shared by oracle and product)."""
    mo = O.Model(**star_kwargs(15))
    du, dd = O.sector_dims(16, 8, 8)
    v0 = O.start_vector(du * dd, 4321)
    egs, nit, a, b, sec = O.stored_lanczos_gs(mo, 8, 8, v0, 300, 1e-12, P=nthreads, nthreads=nthreads)
    return {"config": "cfg2: Norb=1 Nbath=15 Ns=16 sector (8,8)", "dim": du * dd, "seed": 4321,
            "threshold": 1e-12, "ncheck": 10, "egs": egs, "niter": nit, "alanc": list(a), "blanc": list(b),
            "oracle_seconds": sec, "oracle_threads": nthreads,
            "made_by": "tests/golden/make_baseline_goldens.py (oracle ora_stored_lanczos_gs)"}


if __name__ == "__main__":
    nthreads = len(os.sched_getaffinity(0))
    out = cfg2(nthreads)
    with open(os.path.join(HERE, "baseline_cfg2.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(out["egs"], out["niter"], out["oracle_seconds"])
