"""Time-dependent hard-coded boundary values through the device path: SUPGFluidSolver::run (reference
source/mpi_supg_solver.cpp:427-486) advances the clock of the boundary functions by dt before every make_constraints() and
applies the nonzero constraints in every step. Cases: the reference's acoustic_duct_wave_mpi and acoustic_pml_mpi
(tests/acoustic_cases.py), checked against the oracle fixture tests/golden/scns_acoustic_oracle.npz and the reference goldens.

STATUS: written after the round's GPU budget was spent. On the emulated device (tests/cpu_emul) all four tests pass -
both 100-step parity runs, the acoustic_duct_wave_mpi golden (1000 steps, 53 min there) and the acoustic_pml_mpi golden (500
steps); nothing here has run on a B200 yet. The file sorts after the verified suites.

Tolerances: fields after 100 steps 1e-5 relative (200 FGMRES solves, tightened to 1e-10 |rhs| on the device, sparse direct
in the oracle); goldens as in the reference's drivers (5.93 +- 1e-3; |v| < 5e-2)."""
import os

import numpy as np
import pytest

import acoustic_cases

pytestmark = pytest.mark.gpu


def _flow(case, n_steps=None):
    import openifem_b200 as ifem

    c = acoustic_cases.CASES[case]
    params = ifem.Parameters.AllParameters(text=acoustic_cases.prm_text(case, n_steps))
    tria = ifem.Triangulation(2)
    ifem.GridGenerator.subdivided_hyper_rectangle(tria, c["reps"], (0, 0), c["hi"], True)
    flow = ifem.Fluid.MPI.SCnsIM(tria, params)
    flow.add_hard_coded_boundary_condition(0, acoustic_cases.gaussian_pulse(case, 1e-7))
    if c["pml"]:
        flow.set_sigma_pml_field(acoustic_cases.sigma_pml_field)
    return flow


@pytest.mark.parametrize("case", ["duct", "pml"])
def test_acoustic_100_steps_match_oracle(golden_dir, case):
    z = np.load(os.path.join(golden_dir, "scns_acoustic_oracle.npz"))
    flow = _flow(case, n_steps=100)
    flow.set_control(fgmres_rel=1e-10)  # parity run: linear solves tightened as in tests/test_scns_gpu.py
    flow.run()  # refines 3 times, then 100 steps with the constraints re-made every step
    sol = flow.get_current_solution()
    ref = z[case + "_solution_100"]
    assert sol.size == ref.size
    assert np.linalg.norm(sol - ref) / np.linalg.norm(ref) < 1e-5
    h = flow.history()
    assert h[-1]["timestep"] == 100


def test_acoustic_duct_wave_reference_golden():
    """tests/acoustic_duct_wave_mpi/acoustic_duct_wave_mpi.cpp:60-68: max velocity 5.93 +- 1e-3 after 1000 steps"""
    flow = _flow("duct")
    # the converged value (oracle: 5.93536) sits 9e-4 from the rounded golden, i.e. 1e-4 inside its tolerance: the linear solves
    # are tightened from the reference's 1e-6 |rhs| so that solver noise over 1000 steps cannot decide the assertion
    flow.set_control(fgmres_rel=1e-8)
    flow.run()
    vmax = flow.get_current_solution()[: flow.n_u].max()
    assert abs(vmax - 5.93) / 5.93 < 1e-3, vmax
    assert abs(vmax - 5.935360717) / 5.93 < 1e-4, vmax  # the oracle's value


def test_acoustic_pml_reference_golden():
    """tests/acoustic_pml_mpi/acoustic_pml_mpi.cpp:79-85: the pulse is absorbed, |max velocity| < 5e-2 after 500 steps"""
    flow = _flow("pml")
    flow.run()
    vmax = flow.get_current_solution()[: flow.n_u].max()
    assert abs(vmax) < 5e-2, vmax
