"""GPU tests of the mixed-precision inner solver (openifem_b200/csrc/inner32.*): the fp32 SELL-32 copy of the
velocity block and the fp32 BiCGStab that stands in for the reference's MUMPS factorisation of
system_matrix.block(0,0) (source/mpi_insim.cpp:111-127).

Tolerances: product kernel against the fp64 BCSR product 2e-6 of max|y| (fp32 rounding of ~100 terms);
converged Newton states against the oracle 1e-6 relative, as in tests/test_ins_gpu.py - the inner solve is only
a preconditioner of the flexible GMRES, so its precision must not show in converged quantities."""
import numpy as np
import pytest

from util import cavity_prm, make_gpu, make_oracle, rel

pytestmark = pytest.mark.gpu

VARIANTS = [13, 16, 23, 24, 26, 43]
VARIANTS_H = [13, 16, 23, 24, 26, 43]


@pytest.mark.parametrize("dim,reps,hi", [(3, (5, 4, 6), (1.0, 1.2, 0.9)), (2, (9, 7), (1.0, 0.8)), (3, (12, 12, 12), (1, 1, 1))])
def test_sell_product_matches_fp64_product(dim, reps, hi):
    g = make_gpu(cavity_prm(dim), reps, (0,) * dim, hi)
    rng = np.random.default_rng(11)
    ev = rng.uniform(-1, 1, g.n_dofs)
    g.set_vector(g.EVALUATION_POINT, ev)
    g.set_vector(g.PRESENT, 0.5 * ev)
    g.assemble(True)
    g.set_vector(g.SYSTEM_RHS, rng.uniform(-1, 1, g.n_dofs))  # x of the product
    for v in VARIANTS:
        ms, nbytes, pad, err = g.bench_spmv_uu_sell(1, variant=v, precision=32)
        assert 1.0 <= pad < 2.5
        assert 0.0 <= err < 2e-6, (v, err)
    # row-scaled fp16 storage of the values (a_inv_fp32 = 3): 2^-11 relative rounding per entry
    for v in VARIANTS_H:
        ms, nbytes16, pad, err = g.bench_spmv_uu_sell(1, variant=v, precision=16)
        assert 1.0 <= pad < 2.5 and nbytes16 < nbytes
        assert 0.0 <= err < 2e-3, (v, err)


@pytest.mark.parametrize("dim,reps", [(3, (5, 4, 6)), (2, (9, 7))])
def test_fp32_cg_for_mass_schur(dim, reps):
    """'CG for Sm' (mpi_insim.cpp:88-109) in fp32 on the SELL-32 copy of S_m, driven from device-resident scalars, against
    the assembled S_m: residual of its solution in fp64 on the host"""
    g = make_gpu(cavity_prm(dim), reps, (0,) * dim, (1,) * dim)
    g.run_one_step(True)  # forms S_m
    S = g.get_matrix(2)
    rng = np.random.default_rng(13)
    b = S @ rng.uniform(-1, 1, S.shape[0])  # in the range of S_m (closed cavity: constant-pressure null space)
    nb = np.linalg.norm(b)
    x0, it0, _ = g.solve_mass_schur(b, mode=0, rel_tol=1e-5)
    assert np.linalg.norm(S @ x0 - b) / nb < 2e-5
    # fp32 values: solved to 1e-5; fp16 values: the copy differs from S_m by 2^-11 relative per entry
    for mode, tol in [(1, 5e-5), (2, 1e-2)]:
        x, it, res = g.solve_mass_schur(b, mode=mode, rel_tol=1e-5)
        assert np.linalg.norm(S @ x - b) / nb < tol, (mode, it, res)
        assert it <= 2 * it0 + 10, (mode, it, it0)  # the two-level preconditioner of the fp32 path needs FEWER iterations than plain CG
        assert res <= 1.01e-5 * nb


@pytest.mark.parametrize("mode,sm_mode", [(2, 0), (3, 0), (3, 2), (2, 1)])
@pytest.mark.parametrize("dim,reps,steps", [(3, (4, 4, 4), 2), (2, (8, 8), 3)])
def test_time_steps_with_fp32_inner_solver_match_oracle(dim, reps, steps, mode, sm_mode):
    prm = cavity_prm(dim, newton_tol=1e-9)
    o = make_oracle(prm, reps, (0,) * dim, (1,) * dim)
    g = make_gpu(prm, reps, (0,) * dim, (1,) * dim)
    o.fgmres_rel = 1e-9
    g.set_control(a_inv_rel=1e-3, a_inv_max_it=500, fgmres_rel=1e-9, a_inv_fp32=mode, cg_sm_fp32=sm_mode)
    for k in range(steps):
        o.run_one_step(k == 0)
        g.run_one_step(k == 0)
    sol = g.get_current_solution()
    nu = o.n_u
    assert rel(sol[:nu], o.velocity()) < 1e-6
    pg, po = sol[nu:] - sol[nu:].mean(), o.pressure() - o.pressure().mean()
    assert rel(pg, po) < 1e-6
    hg, ho = g.history(), o.history
    assert [(a["timestep"], a["iteration"]) for a in hg] == [(b[0], b[1]) for b in ho]
    # residual norms below ~1e-11 are set by the linear-solver tolerance (1e-9 x the previous residual), not by the
    # discretisation: compare them with that absolute floor
    for a, b in zip(hg, ho):
        assert abs(a["abs_res"] - b[2]) <= 1e-6 * b[2] + 1e-11
    assert sum(h["a_inv_its"] for h in hg) > 0
