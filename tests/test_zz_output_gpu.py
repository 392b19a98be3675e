"""Result files and checkpoints written by the solvers on the device path (openifem_b200/csrc/solver_io.cu; formats checked
on the CPU in tests/test_output_cpu.py): FluidSolver::output_results / save_checkpoint / load_checkpoint (reference
source/mpi_fluid_solver.cpp:491-713) and the solid's (source/mpi_shared_solid_solver.cpp:237-337, 452-571).

STATUS: written after the round's GPU budget was spent; host-side file code verified on the CPU, all three tests pass on the
emulated device (tests/cpu_emul); not run on a B200 yet. The file sorts after the verified suites.

Properties: the written velocity / pressure equal get_current_solution() at the vertices; a run restarted from the latest
checkpoint continues to the same fields as an uninterrupted run (1e-9: a time step depends on present_solution only)."""
import os
import xml.etree.ElementTree as ET

import numpy as np
import pytest

from util import cavity_prm, rel

pytestmark = pytest.mark.gpu


def _prm(dim, n_steps, dt=1e-2):
    text = cavity_prm(dim, dt=dt, end_time=n_steps * dt)
    return text.replace("set Output interval = 1e6", "set Output interval = %g" % dt).replace("set Save interval = 1e6", "set Save interval = %g" % (2 * dt))


def _fluid(dim, n_steps, directory):
    import openifem_b200 as ifem

    reps = (6, 6) if dim == 2 else (3, 3, 3)
    tria = ifem.Triangulation(dim)
    ifem.GridGenerator.subdivided_hyper_rectangle(tria, reps, (0,) * dim, (1.0,) * dim, True)
    s = ifem.Fluid.MPI.InsIM(tria, ifem.Parameters.AllParameters(text=_prm(dim, n_steps)))
    if directory is not None:
        s.set_output_directory(str(directory))
    s.set_control(a_inv_rel=1e-10, a_inv_max_it=5000, fgmres_rel=1e-10)
    return s


def _point_arrays(path):
    piece = ET.parse(path).getroot().find("UnstructuredGrid/Piece")
    out = {}
    for sec in ("Points", "PointData"):
        for a in piece.find(sec).findall("DataArray"):
            nc = int(a.attrib.get("NumberOfComponents", 1))
            v = np.array(a.text.split(), dtype=np.float64)
            out[a.attrib.get("Name", "points")] = v.reshape(-1, nc) if nc > 1 else v
    return out


@pytest.mark.parametrize("dim", [2, 3])
def test_fluid_output_and_restart(tmp_path, dim):
    a = _fluid(dim, 4, tmp_path)
    a.run()
    files = sorted(os.listdir(tmp_path))
    for k in range(5):  # step 0 and every step after it
        assert "fluid_%06d.pvtu" % k in files and "fluid_%06d.proc0000.vtu" % k in files
    assert "fluid.pvd" in files and "000002.fluid_checkpoint" in files and "000004.fluid_checkpoint" in files
    sets = ET.parse(str(tmp_path / "fluid.pvd")).getroot().findall("Collection/DataSet")
    assert [s.attrib["file"] for s in sets] == ["fluid_%06d.pvtu" % k for k in range(5)]
    # the last file holds the current solution at the vertices
    arr = _point_arrays(str(tmp_path / "fluid_000004.proc0000.vtu"))
    sol, pts = a.get_current_solution(), a.support_points()
    lookup = {tuple(np.round(p, 12)): i for i, p in enumerate(pts[: a.n_u: dim])}
    for x, v in zip(arr["points"][:, :dim], arr["velocity"][:, :dim]):
        node = lookup[tuple(np.round(x, 12))]
        assert np.allclose(v, sol[dim * node: dim * node + dim], atol=1e-15)
    plookup = {tuple(np.round(p, 12)): i for i, p in enumerate(pts[a.n_u:])}
    for x, p in zip(arr["points"][:, :dim], arr["pressure"]):
        assert abs(p - sol[a.n_u + plookup[tuple(np.round(x, 12))]]) < 1e-15
    # restart: a fresh solver in the same directory picks up 000004 and runs steps 5 and 6
    b = _fluid(dim, 6, tmp_path)
    b.run()
    assert b.get_time()[1] == 6
    assert [r["timestep"] for r in b.history()][0] == 5
    c = _fluid(dim, 6, None)
    c.run()
    assert rel(b.get_current_solution()[: b.n_u], c.get_current_solution()[: c.n_u]) < 1e-9
    # only the two newest checkpoints are left (the rotation keeps one before writing the next)
    left = sorted(f for f in os.listdir(tmp_path) if f.endswith(".fluid_checkpoint"))
    assert left == ["000004.fluid_checkpoint", "000006.fluid_checkpoint"]


def test_solid_output_and_restart(tmp_path):
    import openifem_b200 as ifem
    import test_linear_elasticity_gpu as T

    def make(n_steps, directory):
        text = T._prm(2).replace("set End time = 1.0", "set End time = %g" % (0.05 * n_steps)).replace(
            "set Output interval = 1.0", "set Output interval = 0.05").replace("set Save interval = 100", "set Save interval = 0.1")
        tria = ifem.Triangulation(2)
        ifem.GridGenerator.subdivided_hyper_rectangle(tria, (7, 3), (0, 0), (4.0, 1.0), True)
        s = ifem.Solid.MPI.LinearElasticity(tria, ifem.Parameters.AllParameters(text=text))
        if directory is not None:
            s.set_output_directory(str(directory))
        return s

    a = make(4, tmp_path)
    a.run()
    files = sorted(os.listdir(tmp_path))
    assert "solid.pvd" in files and "solid_000004.pvtu" in files and "solid_000004.proc0000.vtu" in files
    for ext in ("displacement", "velocity", "acceleration"):
        assert "000004.solid_checkpoint_" + ext in files
    u = a.get_current_solution()
    assert np.array_equal(ifem.io.block_read(str(tmp_path / "000004.solid_checkpoint_displacement"), u.size), u)
    arr = _point_arrays(str(tmp_path / "solid_000004.proc0000.vtu"))
    assert np.array_equal(arr["displacements"][:, :2].ravel(), u)
    b = make(6, tmp_path)
    b.run()
    assert b.get_time()[1] == 6
    c = make(6, None)
    c.run()
    assert rel(b.get_current_solution(), c.get_current_solution()) < 1e-9


def test_fsi_restart(tmp_path):
    """FSI::run with checkpoints of both solvers at the save interval and a restart from them (reference source/mpi_fsi.cpp:
    1127-1151, 1176-1179, 1219-1223): four coupled steps, then a fresh pair of solvers in the same directories continues to step
    six and must agree with six uninterrupted steps"""
    import openifem_b200 as ifem
    from test_fsi_gpu import _fsi_text

    def pair(n_steps, dirs):
        text = _fsi_text(2).replace("set End time = 1.0", "set End time = %g" % (n_steps * 1e-3)).replace("set Save interval = 1e6", "set Save interval = 2e-3")
        assert "End time = %g" % (n_steps * 1e-3) in text and "Save interval = 2e-3" in text
        params = ifem.Parameters.AllParameters(text=text)
        ftria, stria = ifem.Triangulation(2), ifem.Triangulation(2)
        ifem.GridGenerator.subdivided_hyper_rectangle(ftria, (12, 12), (0.0, 0.0), (1.0, 1.0), True)
        ifem.GridGenerator.subdivided_hyper_rectangle(stria, (4, 6), (0.3125, 0.0), (0.5625, 0.6875), True)
        fluid, solid = ifem.Fluid.MPI.SCnsIM(ftria, params), ifem.Solid.MPI.HyperElasticity(stria, params)
        fluid.setup()
        solid.setup()
        fluid.set_control(fgmres_rel=1e-10)
        if dirs:
            fluid.set_output_directory(str(dirs[0]))
            solid.set_output_directory(str(dirs[1]))
        return fluid, solid, ifem.MPI.FSI(fluid, solid, params, False)

    dirs = (tmp_path / "fluid", tmp_path / "solid")
    fluid, solid, fsi = pair(4, dirs)
    fsi.run()
    assert "000004.fluid_checkpoint" in os.listdir(dirs[0]) and "000004.solid_checkpoint_displacement" in os.listdir(dirs[1])
    fluid_b, solid_b, fsi_b = pair(6, dirs)
    fsi_b.run()
    assert fluid_b.get_time()[1] == 6 and solid_b.get_time()[1] == 6
    fluid_c, solid_c, fsi_c = pair(6, None)
    fsi_c.run()
    assert np.abs(solid_c.get_current_solution()).max() > 0
    assert rel(solid_b.get_current_solution(), solid_c.get_current_solution()) < 1e-8
    assert rel(fluid_b.get_current_solution(), fluid_c.get_current_solution()) < 1e-8
