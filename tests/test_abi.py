"""The C-ABI library loads and exports every symbol include/openifem_b200.h declares
(no compute without a GPU; compute entry points must fail loudly, not fall back)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "openifem_b200.h")).read()
    return sorted(set(re.findall(r"\b(ifem_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from openifem_b200 import build
    from openifem_b200._lib import lib

    build.build()
    L = lib()
    names = _declared()
    assert len(names) > 30
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/openifem_b200.h but not exported"


def test_host_side_objects_work_without_gpu(golden_dir):
    import openifem_b200 as ifem

    tria = ifem.Triangulation(2)
    ifem.GridGenerator.subdivided_hyper_rectangle(tria, (4, 2), (0, 0), (2, 1), True)
    assert tria.n_active_cells() == 8 and tria.n_vertices() == 15
    tria.refine_global(1)
    assert tria.n_active_cells() == 32 and tria.n_vertices() == 45
    ifem.Parameters.AllParameters(os.path.join(golden_dir, "ins_cavity_2d.prm"))
    with pytest.raises(ifem.IfemError):
        ifem.Parameters.AllParameters(text="subsection Simulation\n set No such key = 1\nend\n")


def test_compute_fails_loudly_without_gpu(golden_dir):
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import openifem_b200 as ifem

    tria = ifem.Triangulation(2)
    ifem.GridGenerator.hyper_cube(tria, 0, 1, True)
    prm = ifem.Parameters.AllParameters(os.path.join(golden_dir, "ins_cavity_2d.prm"))
    with pytest.raises(ifem.IfemError, match="no CUDA device"):
        ifem.Fluid.MPI.InsIM(tria, prm)
