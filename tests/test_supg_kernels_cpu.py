"""The kernel bodies of the SUPGInsIM device path (openifem_b200/csrc/insim_supg.cuh) run on the CPU: the same source the
CUDA kernel wraps is compiled with g++ (tests/cpp/supg_kernels_cpu.cpp) and walked over its launch grid phase by phase,
then compared with the oracle (oracle/scns.py SUPGInsIM, pinned on the reference goldens
fluid_pressure_driven_mpi_insim_supg and fluid_plane_wall_driven_mpi_insim_supg). This checks the device arithmetic, the
row-plane BCSR indexing of the four blocks and the constrained scatter without a GPU; the launches themselves are covered
by tests/test_zz_supg_insim_gpu.py."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest
import scipy.sparse as sp

from oracle import fem, prm, scns
from util import cavity_prm

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def harness():
    out = os.path.join(ROOT, "tests", "cpp", "_build", "libsupg_kernels_cpu.so")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    src = os.path.join(ROOT, "tests", "cpp", "supg_kernels_cpu.cpp")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-shared", "-fPIC", "-Wno-unknown-pragmas", src, "-o", out])
    return C.CDLL(out)


def _p(a, t=C.c_double):
    return None if a is None else a.ctypes.data_as(C.POINTER(t))


def _pattern(rows_nodes, cols_nodes, n_rows, n_cols):
    k = rows_nodes.shape[1]
    r = np.repeat(rows_nodes, k, axis=1).ravel()
    c = np.tile(cols_nodes, (1, k)).ravel()
    P = sp.coo_matrix((np.ones(r.size), (r, c)), shape=(n_rows, n_cols)).tocsr()
    P.sort_indices()
    return P.indptr.astype(np.int64), P.indices.astype(np.int32)


def _slots(rows_nodes, cols_nodes, rowptr, col):
    nc, k = rows_nodes.shape
    s = np.zeros((nc, k, k), dtype=np.uint8)
    for c in range(nc):
        for a in range(k):
            lst = col[rowptr[rows_nodes[c, a]]: rowptr[rows_nodes[c, a] + 1]]
            s[c, a] = np.searchsorted(lst, cols_nodes[c])
    return s


def _blocks_to_coo(val, rowptr, col, R, Cc, row_of, col_of):
    """row-plane BCSR (val[rowptr[i] R C + (r C + c) nb + j]) -> (rows, cols, vals) in the global [u | p] numbering"""
    rows, cols, vals = [], [], []
    for i in range(rowptr.size - 1):
        nb = rowptr[i + 1] - rowptr[i]
        blk = val[rowptr[i] * R * Cc: rowptr[i + 1] * R * Cc].reshape(R, Cc, nb)
        cj = col[rowptr[i]: rowptr[i + 1]]
        for r in range(R):
            for c in range(Cc):
                rows.append(np.full(nb, row_of(i, r))), cols.append(col_of(cj, c)), vals.append(blk[r, c])
    return np.concatenate(rows), np.concatenate(cols), np.concatenate(vals)


def _device_layout_assemble(h, o, nonzero):
    dim, d = o.dim, o.dofs
    un, pn = np.ascontiguousarray(d.unodes, dtype=np.int32), np.ascontiguousarray(d.pnodes, dtype=np.int32)
    nU, nP = d.n_unodes, d.n_pnodes
    cell_x = np.ascontiguousarray(o.mesh.vertices[o.mesh.cells], dtype=np.float64)
    tables = np.concatenate([o.Nu.ravel(), o.dNu.ravel(), o.Np.ravel(), o.dNgeo.ravel(), o.qw.ravel()])
    pats = {"uu": _pattern(un, un, nU, nU), "up": _pattern(un, pn, nU, nP), "pu": _pattern(pn, un, nP, nU), "pp": _pattern(pn, pn, nP, nP)}
    slots = np.ascontiguousarray(np.concatenate([
        _slots(un, un, *pats["uu"]).reshape(len(un), -1), _slots(un, pn, *pats["up"]).reshape(len(un), -1),
        _slots(pn, un, *pats["pu"]).reshape(len(un), -1), _slots(pn, pn, *pats["pp"]).reshape(len(un), -1)], axis=1))
    uu = np.zeros(pats["uu"][1].size * dim * dim)
    up = np.zeros(pats["up"][1].size * dim)
    pu = np.zeros(pats["pu"][1].size * dim)
    pp = np.zeros(pats["pp"][1].size)
    rhs = np.zeros(o.n)
    grav = np.asarray(o.prm.gravity, dtype=np.float64)
    inhom = np.ascontiguousarray(o.nonzero_val) if nonzero else None
    rc = h.cpu_supg_assemble(
        C.c_int(dim), C.c_int(o.mesh.n_cells), _p(un, C.c_int), _p(pn, C.c_int), _p(cell_x), _p(tables), _p(slots, C.c_ubyte),
        _p(np.ascontiguousarray(o.con), C.c_ubyte), _p(o.evaluation_point), _p(o.present), _p(o.body_force), _p(inhom),
        C.c_int64(o.n_u), C.c_int(nU), C.c_int(nP), C.c_double(o.prm.viscosity), C.c_double(o.prm.fluid_rho), C.c_double(o.dt),
        _p(grav), _p(pats["uu"][0], C.c_int64), _p(pats["up"][0], C.c_int64), _p(pats["pu"][0], C.c_int64),
        _p(pats["pp"][0], C.c_int64), _p(uu), _p(up), _p(pu), _p(pp), _p(rhs))
    assert rc == 0
    nu_ = o.n_u
    parts = [
        _blocks_to_coo(uu, *pats["uu"], dim, dim, lambda i, r: i * dim + r, lambda cj, c: cj * dim + c),
        _blocks_to_coo(up, *pats["up"], dim, 1, lambda i, r: i * dim + r, lambda cj, c: nu_ + cj),
        # A_pu: one row per pressure node, dim planes = the velocity component of the column
        _blocks_to_coo(pu, *pats["pu"], 1, dim, lambda i, r: nu_ + i, lambda cj, c: cj * dim + c),
        _blocks_to_coo(pp, *pats["pp"], 1, 1, lambda i, r: nu_ + i, lambda cj, c: nu_ + cj),
    ]
    rows, cols, vals = (np.concatenate([p[k] for p in parts]) for k in range(3))
    return sp.coo_matrix((vals, (rows, cols)), shape=(o.n, o.n)).tocsr(), rhs


CASES = [
    (cavity_prm(2), (6, 5), (0, 0), (1.0, 0.8)),
    (cavity_prm(3, mu=0.05), (3, 4, 3), (0, 0, 0), (1.0, 1.2, 0.9)),
    (cavity_prm(2, gravity=[10.0, -3.0], dirichlet={2: (3, [0, 0]), 3: (3, [0.5, 0])}), (7, 4), (0, 0), (2.0, 0.2)),
]


@pytest.mark.parametrize("case", range(len(CASES)))
@pytest.mark.parametrize("nonzero", [True, False])
def test_supg_kernel_bodies_match_oracle(harness, case, nonzero):
    text, reps, lo, hi = CASES[case]
    text = text.replace("set Velocity degree = 2", "set Velocity degree = 1")
    bf = (lambda x, c: 0.3 * (c + 1) * x[0] - 0.2 * x[1]) if case == 2 else None
    o = scns.SUPGInsIM(fem.BoxMesh(tuple(reps), lo, hi), prm.Params(text, is_text=True), body_force=bf)
    rng = np.random.default_rng(20 + case)
    o.evaluation_point[:] = rng.uniform(-1, 1, o.n)
    o.present[:] = rng.uniform(-1, 1, o.n)
    if case == 1:  # a cell with zero previous-step velocity: the h == 0 branch of the UGN parameters
        o.present[:] = 0.0
    A_ref, rhs_ref = o.assemble(nonzero)
    A, rhs = _device_layout_assemble(harness, o, nonzero)
    assert sp.linalg.norm(A - A_ref) / sp.linalg.norm(A_ref) < 1e-13
    assert np.linalg.norm(rhs - rhs_ref) / np.linalg.norm(rhs_ref) < 1e-13
