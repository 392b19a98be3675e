"""Generates tests/golden/scns_acoustic_oracle.npz: the CPU oracle (oracle/scns.py) on the reference's two time-dependent
SCnsIM cases (tests/acoustic_cases.py). Stored per case: max velocity every 50 steps over the whole run (the last one is
what the reference's driver asserts on) and the full solution after 100 steps (what the short CPU / GPU parity tests
compare with). Runs about two minutes:  python scripts/make_acoustic_fixture.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import acoustic_cases  # noqa: E402

out = {}
for case in ("duct", "pml"):
    s = acoustic_cases.make_oracle(case)
    n_steps = int(round(s.prm.end_time / s.dt))
    vmax = []
    for k in range(n_steps):
        if s.hard_coded:  # SUPGFluidSolver::run (oracle/scns.py run())
            s.bc_time += s.dt
            s.make_constraints()
        s.run_one_step(True)
        if (k + 1) % 50 == 0:
            vmax.append(s.velocity().max())
        if k + 1 == 100:
            out[case + "_solution_100"] = s.present.copy()
    out[case + "_vmax_every_50"] = np.asarray(vmax)
    print(case, "steps", s.timestep, "final max velocity", vmax[-1])
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "scns_acoustic_oracle.npz"), **out)
