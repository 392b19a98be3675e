"""BASELINE config 4 - the reference's tests/fsi_leaflet_mpi case as its driver meshes it (fsi_leaflet_mpi.cpp:47-92, .prm verbatim in
tests/golden/fsi_leaflet_2d.prm): MPI::FSI<2>(SCnsIM Q1/Q1 on an 80 x 20 channel whose band around the leaflet is refined once -
hanging nodes -, SharedHyperElasticity NeoHookean leaflet on 8 x 32 cells, use_dirichlet_bc = true), hard-coded inflow u_x = 1.5.

The reference pins nothing for this case (smoke test: "parity unpinned"); the device loop is compared with the oracle's FSI loop
(oracle/fsi.py on oracle/scns.py + oracle/solid.py) step by step: fluid velocity / pressure and solid displacement to 1e-6
relative (device FGMRES to 1e-10 |rhs| for the comparison, oracle sparse direct), indicator field exact."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

L, H, A, B, HH, U = 4.0, 1.0, 0.1, 0.4, 0.05, 1.5


def inflow(p, c, t):
    return U if c == 0 and abs(p[0]) < 1e-10 else 0.0


def leaflet_case(golden_dir, fluid_h=HH):
    """both sides of the case as the reference driver builds it; fluid_h = 0.05 is the reference's mesh"""
    import openifem_b200 as ifem
    from oracle import fem, fsi, grid, prm, scns, solid

    path = os.path.join(golden_dir, "fsi_leaflet_2d.prm")
    params = ifem.Parameters.AllParameters(path)
    ftria = ifem.Triangulation(2)
    ifem.GridGenerator.subdivided_hyper_rectangle(ftria, (int(round(L / fluid_h)), int(round(H / fluid_h))), (0, 0), (L, H), True)
    v, c, _ = ftria.get_mesh()
    cx = v[c].mean(axis=1)[:, 0]
    ftria.execute_refinement(((cx >= L / 4 - 2 * A) & (cx <= L / 4 + 3 * A)).astype(np.uint8))
    fluid = ifem.Fluid.MPI.SCnsIM(ftria, params)
    fluid.add_hard_coded_boundary_condition(0, inflow)
    fluid.setup()
    stria = ifem.Triangulation(2)
    ifem.GridGenerator.subdivided_hyper_rectangle(stria, (int(round(A / HH)), int(round(B / HH))), (L / 4, 0), (A + L / 4, B), True)
    stria.refine_global(2)  # Global refinements = 0, 2 (FSI::run, mpi_fsi.cpp:1126)
    sol = ifem.Solid.MPI.SharedHyperElasticity(stria, params)
    sol.setup()
    coupling = ifem.MPI.FSI(fluid, sol, params, True)

    P = prm.Params(path)
    v, c, b = ftria.get_mesh()
    o_fluid = scns.SCnsIM(grid.QuadMesh(v, c, b), P, hard_coded={0: inflow})
    o_solid = solid.HyperElasticity(fem.BoxMesh((4 * int(round(A / HH)), 4 * int(round(B / HH))), (L / 4, 0), (A + L / 4, B)), P)
    loop = fsi.FSI(o_fluid, o_solid, True)
    return ftria, fluid, sol, coupling, o_fluid, o_solid, loop


def _rel(a, b):
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(np.asarray(b)), 1e-300)


def test_leaflet_mesh_is_the_reference_mesh(golden_dir):
    ftria, fluid, sol, coupling, o_fluid, o_solid, _ = leaflet_case(golden_dir)
    # 80 x 20 cells, the 10 columns of the band replaced by their four children (~2 200 cells, ~7 k dofs)
    assert ftria.n_active_cells() == 80 * 20 + 3 * 10 * 20 == o_fluid.mesh.n_cells
    assert len(o_fluid.dofs.hanging_u) == 2 * 20 and ftria.hanging()[0].size == 40
    assert fluid.n_dofs == o_fluid.n and sol.n_dofs == o_solid.n == 2 * 9 * 33


@pytest.mark.parametrize("steps", [4])
def test_leaflet_steps_match_oracle(golden_dir, steps):
    ftria, fluid, sol, coupling, o_fluid, o_solid, loop = leaflet_case(golden_dir)
    fluid.set_control(fgmres_rel=1e-10)
    for k in range(steps):
        loop.run_one_step(k == 0)
        coupling.run_one_step(k == 0)
        assert np.array_equal(coupling.get_indicator(), o_fluid.indicator)
        fsol = fluid.get_current_solution()
        eu, ep = _rel(fsol[: o_fluid.n_u], o_fluid.velocity()), _rel(fsol[o_fluid.n_u:], o_fluid.pressure())
        es = _rel(sol.get_current_solution(), o_solid.cur_u)
        assert eu < 1e-6 and ep < 1e-6 and es < 1e-6, (k, eu, ep, es)
    assert o_fluid.indicator.sum() > 0 and np.abs(o_solid.cur_u).max() > 0
    assert np.abs(o_fluid.velocity()).max() >= U  # the inflow has entered the channel (and accelerates over the leaflet)


def test_cpp_leaflet_driver_runs_the_reference_driver_body(golden_dir, tmp_path):
    """the reference's own driver body (tests/cpp/fsi_leaflet_mpi.cpp: band refinement through cell iterators, FSI::run) compiled
    against the C++ facade, End time cut to six steps, against the same six steps through the Python mirror"""
    if os.environ.get("IFEM_CPU_EMULATION"):
        pytest.skip("compiled drivers link the product library: not replayable on the emulated device")
    import subprocess
    import sys

    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from test_cpp_facade import ROOT, _build

    _build("fsi_leaflet_mpi")
    text = open(os.path.join(golden_dir, "fsi_leaflet_2d.prm")).read().replace("set End time = 2e0", "set End time = 3e-2")
    prm_file = tmp_path / "leaflet_short.prm"
    prm_file.write_text(text)
    r = subprocess.run([os.path.join(ROOT, "tests", "cpp", "_build", "fsi_leaflet_mpi"), str(prm_file)], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr + r.stdout
    line = [l for l in r.stdout.splitlines() if l.startswith("cells ")][-1].split()
    cells, vmax, pmax, umax = int(line[1]), float(line[3]), float(line[5]), float(line[7])
    ftria, fluid, sol, coupling, o_fluid, _, _ = leaflet_case(golden_dir)
    for k in range(6):
        coupling.run_one_step(k == 0)
    fsol = fluid.get_current_solution()
    assert cells == ftria.n_active_cells() == 2200
    assert abs(vmax - fsol[: o_fluid.n_u].max()) <= 1e-9 * abs(vmax)
    assert abs(pmax - fsol[o_fluid.n_u:].max()) <= 1e-9 * abs(pmax)
    assert abs(umax - np.abs(sol.get_current_solution()).max()) <= 1e-9 * umax
