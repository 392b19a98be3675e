"""Multi-GPU parity: the slab-partitioned path (owner-computes assembly, NCCL halo exchange,
all-reduced Krylov dot products, matrix-free S_m) must reproduce the single-GPU path, which the other GPU
tests tie to the oracle. Needs >= 2 GPUs (run with `gpurun --gpus 2`); skipped otherwise.

Tolerances: block SpMV 1e-13 relative; Newton residual history and final fields 1e-6 relative with the
linear solves tightened on both sides (same rule as tests/test_ins_gpu.py)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _n_gpus():
    import torch

    return torch.cuda.device_count()


def _worker(rank, size, idfile, dim, reps, steps, q, inner_mode=0):
    try:
        sys.path.insert(0, ROOT)
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import time

        import torch

        torch.cuda.set_device(rank)
        from util import cavity_prm

        import openifem_b200 as ifem

        ifem.init(rank)
        if size > 1:
            if rank == 0:
                uid = ifem.comm_unique_id()
                with open(idfile + ".tmp", "wb") as f:
                    f.write(uid)
                os.replace(idfile + ".tmp", idfile)
            else:
                while not os.path.exists(idfile):
                    time.sleep(0.05)
                uid = open(idfile, "rb").read()
            ifem.comm_init(rank, size, uid)
        tria = ifem.Triangulation(dim)
        ifem.GridGenerator.subdivided_hyper_rectangle(tria, reps, (0,) * dim, (1,) * dim, True)
        flow = ifem.Fluid.MPI.InsIM(tria, ifem.Parameters.AllParameters(text=cavity_prm(dim, newton_tol=1e-9)))
        flow.setup()
        if inner_mode >= 2:
            # fp32 inner solvers on the sliced copies of A_uu / S_m (2: fp32 values, 3: fp16 values): fp32 cannot reach
            # 1e-10, FGMRES (flexible) still converges
            flow.set_control(a_inv_rel=1e-3, a_inv_max_it=500, fgmres_rel=1e-9, a_inv_fp32=inner_mode, cg_sm_fp32=inner_mode - 1)
        else:
            flow.set_control(a_inv_rel=1e-10, a_inv_max_it=5000, fgmres_rel=1e-9)
        n_un_glob = int(np.prod([2 * k + 1 for k in reps]))
        n_pn_glob = int(np.prod([k + 1 for k in reps]))
        n_glob = dim * n_un_glob + n_pn_glob
        loc, glo = flow.owned_global_dofs(n_un_glob)
        # block SpMV of a global vector known to every rank
        gu, gp = flow.local_to_global(0).astype(np.int64), flow.local_to_global(1).astype(np.int64)
        xg = np.sin(0.11 * np.arange(n_glob)) + 0.3
        ev = 0.1 * np.cos(0.05 * np.arange(n_glob))

        def localise(vg):
            vu = vg[(gu[:, None] * dim + np.arange(dim)[None, :]).ravel()]
            return np.concatenate([vu, vg[dim * n_un_glob + gp]])

        flow.set_vector(flow.EVALUATION_POINT, localise(ev))
        flow.set_vector(flow.PRESENT, localise(0.5 * ev))
        flow.assemble(True)
        y = flow.vmult(localise(xg))
        rhs = flow.get_vector(flow.SYSTEM_RHS)
        sell_err = None
        if inner_mode >= 2:
            # the product of the inner solver (SELL-32 copy, ghosts refreshed by the peer push or by NCCL) against the fp64 product
            sell_err = flow.bench_spmv_uu_sell(1, check_error=True)[3]
        # time steps from rest
        zero = np.zeros(flow.n_dofs)
        flow.set_vector(flow.EVALUATION_POINT, zero)
        flow.set_vector(flow.PRESENT, zero)
        for k in range(steps):
            flow.run_one_step(k == 0)
        sol = flow.get_current_solution()
        hist = [(h["timestep"], h["iteration"], h["abs_res"], h["gmres_its"]) for h in flow.history()]
        q.put((rank, "ok", glo, y[loc], rhs[loc], sol[loc], hist, sell_err))
        if size > 1:
            ifem.comm_finalize()
    except Exception as e:  # pragma: no cover
        import traceback

        q.put((rank, "fail", traceback.format_exc(), None, None, None, None, None))


def _run(size, dim, reps, steps, tmp_path, inner_mode=0):
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    idfile = str(tmp_path / f"nccl_id_{size}_{inner_mode}")
    procs = [ctx.Process(target=_worker, args=(r, size, idfile, dim, reps, steps, q, inner_mode)) for r in range(size)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for r in res:
        assert r[1] == "ok", f"rank {r[0]} failed:\n{r[2]}"
    n = sum(len(r[2]) for r in res)
    y, rhs, sol = np.zeros(n), np.zeros(n), np.zeros(n)
    seen = np.zeros(n, dtype=int)
    for r in res:
        y[r[2]], rhs[r[2]], sol[r[2]] = r[3], r[4], r[5]
        seen[r[2]] += 1
    assert np.all(seen == 1)  # owned dofs of the ranks tile the global vector exactly once
    hist = [r for r in res if r[0] == 0][0][6]
    _run.sell_errors = [r[7] for r in res]
    return y, rhs, sol, hist


@pytest.mark.parametrize("dim,reps,size", [(3, (4, 4, 6), 2), (2, (6, 8), 2)])
def test_two_ranks_match_one_rank(dim, reps, size, tmp_path):
    if _n_gpus() < size:
        pytest.skip(f"needs {size} GPUs")
    y1, rhs1, sol1, h1 = _run(1, dim, reps, 2, tmp_path)
    y2, rhs2, sol2, h2 = _run(size, dim, reps, 2, tmp_path)
    rel = lambda a, b: np.linalg.norm(a - b) / np.linalg.norm(b)
    assert rel(y2, y1) < 1e-13
    assert rel(rhs2, rhs1) < 1e-13
    assert len(h1) == len(h2)
    for a, b in zip(h2, h1):
        assert a[:2] == b[:2]
        assert abs(a[2] - b[2]) <= 1e-6 * max(b[2], 1e-9)
    nu = dim * int(np.prod([2 * k + 1 for k in reps]))
    assert rel(sol2[:nu], sol1[:nu]) < 1e-6
    p2, p1 = sol2[nu:] - sol2[nu:].mean(), sol1[nu:] - sol1[nu:].mean()
    assert rel(p2, p1) < 1e-6


@pytest.mark.parametrize("dim,reps,inner_mode", [(3, (4, 4, 6), 3), (2, (6, 8), 2)])
def test_fp32_inner_solver_two_ranks_match_one_rank(dim, reps, inner_mode, tmp_path):
    """a_inv_fp32 = 2 / 3 and cg_sm_fp32 = 1 / 2 (fp32 BiCGStab and CG on the SELL-32 copies of A_uu and S_m, halo
    exchange in the permuted numbering) on two ranks against the fp64 inner solves on one rank: same converged Newton
    states to 1e-6"""
    if _n_gpus() < 2:
        pytest.skip("needs 2 GPUs")
    _, _, sol1, h1 = _run(1, dim, reps, 2, tmp_path)
    _, _, _, h1i = _run(1, dim, reps, 2, tmp_path, inner_mode=inner_mode)
    _, _, sol2, h2 = _run(2, dim, reps, 2, tmp_path, inner_mode=inner_mode)
    # the ghost entries of the inner solver's gather source are right on both ranks (fp16 values: 2^-11 per entry) ...
    assert all(e is not None and e < (2e-3 if inner_mode == 3 else 1e-5) for e in _run.sell_errors), _run.sell_errors
    # ... and the preconditioner is as good as on one rank: a broken halo or all-reduce inside the inner solvers still lets the
    # flexible outer iteration converge to the same fields, only with many more iterations
    assert all(abs(a[3] - b[3]) <= 2 for a, b in zip(h2, h1i)), ([a[3] for a in h2], [b[3] for b in h1i])
    rel = lambda a, b: np.linalg.norm(a - b) / np.linalg.norm(b)
    assert [a[:2] for a in h2] == [b[:2] for b in h1]
    for a, b in zip(h2, h1):
        assert abs(a[2] - b[2]) <= 1e-6 * b[2] + 1e-11  # floor: residuals at the linear-solver tolerance
    nu = dim * int(np.prod([2 * k + 1 for k in reps]))
    assert rel(sol2[:nu], sol1[:nu]) < 1e-6
    p2, p1 = sol2[nu:] - sol2[nu:].mean(), sol1[nu:] - sol1[nu:].mean()
    assert rel(p2, p1) < 1e-6


def test_four_ranks_match_one_rank(tmp_path):
    if _n_gpus() < 4:
        pytest.skip("needs 4 GPUs")
    y1, rhs1, sol1, h1 = _run(1, 3, (4, 4, 12), 1, tmp_path)
    y4, rhs4, sol4, h4 = _run(4, 3, (4, 4, 12), 1, tmp_path)
    rel = lambda a, b: np.linalg.norm(a - b) / np.linalg.norm(b)
    assert rel(y4, y1) < 1e-13 and rel(rhs4, rhs1) < 1e-13
    nu = 3 * 9 * 9 * 25
    assert rel(sol4[:nu], sol1[:nu]) < 1e-6


def _scns_worker(rank, size, idfile, dim, reps, steps, q, mode="plain"):
    try:
        sys.path.insert(0, ROOT)
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import time

        import torch

        torch.cuda.set_device(rank)
        from test_scns_gpu import scns_prm

        import openifem_b200 as ifem

        ifem.init(rank)
        if size > 1:
            if rank == 0:
                uid = ifem.comm_unique_id()
                with open(idfile + ".tmp", "wb") as f:
                    f.write(uid)
                os.replace(idfile + ".tmp", idfile)
            else:
                while not os.path.exists(idfile):
                    time.sleep(0.05)
                uid = open(idfile, "rb").read()
            ifem.comm_init(rank, size, uid)
        tria = ifem.Triangulation(dim)
        solid = None
        if mode.startswith("fsi"):
            # the coupled loop of MPI::FSI::run (tests/test_fsi_gpu.py's 3-D problem): partitioned fluid, replicated solid
            from test_fsi_gpu import _fsi_text

            params = ifem.Parameters.AllParameters(text=_fsi_text(dim, solid_v0=(0.02, 0.0, 0.01)))
            ifem.GridGenerator.subdivided_hyper_rectangle(tria, reps, (0.0,) * dim, (1.0,) * dim, True)
            stria = ifem.Triangulation(dim)
            ifem.GridGenerator.subdivided_hyper_rectangle(stria, (3, 3, 4), (0.25, 0.0, 0.25), (0.7, 0.6, 0.75), True)
            flow, solid = ifem.Fluid.MPI.SCnsIM(tria, params), ifem.Solid.MPI.HyperElasticity(stria, params)
            flow.setup()
            solid.setup()
            flow.set_control(fgmres_rel=1e-10, supg_ilu=0)  # the same (Jacobi) factors on one rank as on several: partition parity only
            coupling = ifem.MPI.FSI(flow, solid, params, mode == "fsi_dirichlet")
        else:
            hi = (2.0,) + (1.0,) * (dim - 1)
            ifem.GridGenerator.subdivided_hyper_rectangle(tria, reps, (0,) * dim, hi, True)
            if mode == "refined":  # a band across the last axis refined once: hanging nodes (the meshes of BASELINE configs 4 / 5)
                v, c, _ = tria.get_mesh()
                z = v[c].mean(axis=1)[:, dim - 1]
                tria.execute_refinement(((z > 0.34) & (z < 0.67)).astype(np.uint8))
            flow = ifem.Fluid.MPI.SCnsIM(tria, ifem.Parameters.AllParameters(text=scns_prm(dim, dt=1e-3)))
            flow.set_body_force(lambda p, c: 5.0 if c == 0 else 0.0)
            flow.setup()
            flow.set_control(fgmres_rel=1e-10, supg_ilu=0)  # the same (Jacobi) factors on one rank as on several: partition parity only
        n_un_glob = tria.n_vertices()
        loc, glo = flow.owned_global_dofs(n_un_glob)
        for k in range(steps):
            if solid is not None:
                coupling.run_one_step(k == 0)
            else:
                flow.run_one_step(k == 0)
        sol = flow.get_current_solution()
        ou = flow.partition(0)[0]
        stress = flow.get_stress()[:, :ou]
        gu = flow.local_to_global(0)[:ou]
        hist = [(h["timestep"], h["iteration"], h["abs_res"]) for h in flow.history()]
        q.put((rank, "ok", glo, sol[loc], gu, stress, hist, None if solid is None else solid.get_current_solution()))
        if size > 1:
            ifem.comm_finalize()
    except Exception:  # pragma: no cover
        import traceback

        q.put((rank, "fail", traceback.format_exc(), None, None, None, None, None))


def _run_scns(size, dim, reps, steps, tmp_path, mode="plain"):
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    idfile = str(tmp_path / f"nccl_id_scns_{mode}_{size}")
    procs = [ctx.Process(target=_scns_worker, args=(r, size, idfile, dim, reps, steps, q, mode)) for r in range(size)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for r in res:
        assert r[1] == "ok", f"rank {r[0]} failed:\n{r[2]}"
    n = sum(len(r[2]) for r in res)
    sol = np.zeros(n)
    n_nodes = sum(len(r[4]) for r in res)
    stress = np.zeros((dim * dim, n_nodes))
    for r in res:
        sol[r[2]] = r[3]
        stress[:, r[4]] = r[5]
    hist = [r for r in res if r[0] == 0][0][6]
    if mode.startswith("fsi"):
        for r in res[1:]:  # the solid is replicated: every rank holds the same displacement
            # (not bitwise on hardware: the solid kernels add with atomics, the CG solves stop at 1e-8 |b|)
            assert np.abs(r[7] - res[0][7]).max() <= 1e-6 * np.abs(res[0][7]).max()
        return sol, stress, hist, res[0][7]
    return sol, stress, hist


@pytest.mark.parametrize("dim,reps", [(3, (5, 4, 6)), (2, (10, 8))])
def test_scnsim_two_ranks_match_one_rank(dim, reps, tmp_path):
    """the slightly compressible solver (assembly with nodal-stress gradients, update_stress, SUPG preconditioner) on
    two ranks against one rank"""
    if _n_gpus() < 2:
        pytest.skip("needs 2 GPUs")
    sol1, st1, h1 = _run_scns(1, dim, reps, 3, tmp_path)
    sol2, st2, h2 = _run_scns(2, dim, reps, 3, tmp_path)
    rel = lambda a, b: np.linalg.norm(a - b) / np.linalg.norm(b)
    assert len(h1) == len(h2)
    for a, b in zip(h2, h1):
        assert a[:2] == b[:2]
        assert abs(a[2] - b[2]) <= 1e-6 * max(b[2], 1e-9)
    assert rel(sol2, sol1) < 1e-6
    assert rel(st2, st1) < 1e-6


@pytest.mark.parametrize("dim,reps,size", [(3, (3, 3, 9), 2), (2, (4, 9), 2), (3, (2, 2, 12), 4)])
def test_scnsim_on_a_band_refined_mesh_ranks_match_one_rank(dim, reps, size, tmp_path):
    """hanging nodes on several GPUs: slabs cut along mesh planes that carry no hanging node or master, condensation local to the
    owner of a line (csrc/hanging.cu, csrc/partition.cpp plane_slabs)"""
    if _n_gpus() < size:
        pytest.skip(f"needs {size} GPUs")
    sol1, st1, h1 = _run_scns(1, dim, reps, 3, tmp_path, "refined")
    sol2, st2, h2 = _run_scns(size, dim, reps, 3, tmp_path, "refined")
    rel = lambda a, b: np.linalg.norm(a - b) / np.linalg.norm(b)
    assert len(h1) == len(h2)
    for a, b in zip(h2, h1):
        assert a[:2] == b[:2] and abs(a[2] - b[2]) <= 1e-6 * max(b[2], 1e-9)
    assert rel(sol2, sol1) < 1e-6 and rel(st2, st1) < 1e-6


@pytest.mark.parametrize("mode", ["fsi", "fsi_dirichlet"])
def test_coupled_fsi_two_ranks_match_one_rank(mode, tmp_path):
    """the full IFEM step of BASELINE config 5's path on real NCCL: partitioned SCnsIM fluid, replicated NeoHookean solid, solid-side
    interpolation summed over the ranks (source/mpi_fsi.cpp:849-865)"""
    if _n_gpus() < 2:
        pytest.skip("needs 2 GPUs")
    sol1, _, h1, us1 = _run_scns(1, 3, (6, 6, 6), 2, tmp_path, mode)
    sol2, _, h2, us2 = _run_scns(2, 3, (6, 6, 6), 2, tmp_path, mode)
    rel = lambda a, b: np.linalg.norm(a - b) / np.linalg.norm(b)
    assert len(h1) == len(h2) and np.abs(us1).max() > 0
    assert rel(sol2, sol1) < 1e-6 and rel(us2, us1) < 1e-6
