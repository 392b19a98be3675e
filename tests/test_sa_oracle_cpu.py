"""The oracle's Spalart-Allmaras restatement (oracle/spalart_allmaras.py; reference source/mpi_spalart_allmaras.cpp) checked
against properties of the model itself - no reference test attaches the model, so there is no golden value to pin it on
("parity unpinned", see the module's header):
  * the matrix of assemble() is the exact derivative of its right-hand side with respect to evaluation_point (the reference
    linearises the transport equation about evaluation_point with the coefficients frozen at present_solution, :799-835);
  * far from any wall and without shear a uniform nu~ is a steady state;
  * eddy viscosity mu_t = f_v1 nu~ rho (:864-889) and the sublayer branch of get_shear_velocity (:227-293)."""
import numpy as np

from test_scns_gpu import scns_prm

SA = """
subsection Spalart Allmaras model
  set Number of S-A model BCs = {n}
  set S-A model boundary id = {ids}
  set S-A model boundary types = {types}
  set Initial condition coefficient = 3.0
  set Wall function image distance = 0.02
end
"""


def _oracle(dim, bcs, **kw):
    from oracle import fem, prm, scns

    text = scns_prm(dim, **kw) + SA.format(n=len(bcs), ids=", ".join(str(i) for i in sorted(bcs)) or "0",
                                            types=", ".join(str(bcs[i]) for i in sorted(bcs)) or "0")
    reps, hi = ((6, 5), (2.0, 1.0)) if dim == 2 else ((3, 3, 3), (2.0, 1.0, 0.9))
    o = scns.SCnsIM(fem.BoxMesh(reps, (0.0,) * dim, hi), prm.Params(text, is_text=True))
    return o, o.attach_turbulence_model("Spalart-Allmaras")


def test_matrix_is_the_derivative_of_the_residual():
    for dim in (2, 3):
        o, t = _oracle(dim, {}, mu=1e-3, rho=1.2, dt=1e-2)  # no lines: the raw system
        t.fixed_wall_distance = 0.05 + np.abs(t.coords[:, 1])  # a wall along y = 0, kept at a distance
        rng = np.random.default_rng(3)
        t.present[:] = t.nu_laminar * rng.uniform(0.2, 5.0, t.n)
        t.evaluation_point[:] = t.present + t.nu_laminar * rng.uniform(-0.2, 0.2, t.n)
        o.present[:] = rng.uniform(-1, 1, o.n)
        A, b = t.assemble(False)
        delta = t.nu_laminar * rng.uniform(-1, 1, t.n)
        errs = []
        for eps in (1e-3, 1e-4):
            e0 = t.evaluation_point.copy()
            t.evaluation_point = e0 + eps * delta
            _, b1 = t.assemble(False)
            t.evaluation_point = e0
            errs.append(np.linalg.norm((b1 - b) / eps + A @ delta) / np.linalg.norm(A @ delta))
        assert errs[0] < 1e-2 and errs[1] < 0.2 * errs[0], errs  # first-order difference quotient: the error shrinks with eps


def test_uniform_state_far_from_walls_is_steady():
    o, t = _oracle(2, {}, mu=1e-3, rho=1.0, dt=1e-2)
    assert (t.fixed_wall_distance == np.finfo(np.float64).max).all()  # no wall boundary at all (:521)
    assert np.allclose(t.present, 3.0 * t.nu_laminar, rtol=0, atol=0)
    _, b = t.assemble(False)
    assert np.abs(b).max() < 1e-18
    t.run_one_step(True)
    assert np.allclose(t.present, 3.0 * t.nu_laminar, rtol=1e-14, atol=0)


def test_wall_and_inflow_lines():
    o, t = _oracle(2, {0: 1, 2: 0, 3: 0}, mu=1e-3, rho=2.0)
    y, x = t.coords[:, 1], t.coords[:, 0]
    inflow = x == 0.0  # boundary id 0 is visited first: the corner nodes keep the inflow value
    wall = ((y == 0.0) | (y == 1.0)) & ~inflow
    assert (t.con[wall] == 1).all() and (t.nonzero_val[wall] == 0.0).all()
    assert (t.con[inflow] == 1).all() and np.allclose(t.nonzero_val[inflow], 5.0 * 1e-3 / 2.0)
    assert t.con.sum() == wall.sum() + inflow.sum()
    # distance to the nearest wall VERTEX (:497-551), not to the wall itself
    h = 2.0 / 6
    for n in np.nonzero((y == 0.4))[0]:
        nearest = min(np.hypot(x[n] - k * h, 0.4) for k in range(7))
        assert abs(t.fixed_wall_distance[n] - nearest) < 1e-15
    # first step applies the inflow value: nu~ = 5 nu on the inflow nodes, 0 on the walls
    t.run_one_step(True)
    assert np.allclose(t.present[inflow], 5.0 * t.nu_laminar, rtol=1e-12) and np.abs(t.present[wall]).max() == 0.0


def test_eddy_viscosity_and_wall_law():
    o, t = _oracle(2, {2: 0}, mu=1.8e-5, rho=1.2)
    t.present[:] = t.nu_laminar * np.linspace(0.0, 30.0, t.n)
    mu_t = t.update_eddy_viscosity()
    chi = t.present / t.nu_laminar
    assert np.allclose(mu_t, chi ** 3 / (chi ** 3 + 7.1 ** 3) * t.present * 1.2, rtol=1e-15)
    assert mu_t[0] == 0.0 and abs(mu_t[-1] / (1.2 * t.present[-1]) - 1.0) < 0.02  # f_v1 -> 1 for chi >> c_v1
    nu, dist = t.nu_laminar, 0.02
    # viscous sublayer: u+ = y+  =>  u_tau = sqrt(u nu / y)
    vel = 1e-3
    assert vel * dist / nu < np.sqrt(5.0)
    assert abs(t.get_shear_velocity(vel, 0.0) - np.sqrt(vel * nu / dist)) < 1e-15
    assert t.get_shear_velocity(0.0, 1.0) == 0.0
    # beyond the sublayer the reference iterates on the composite law with a loose stopping rule (|step| < 1e-2 |u_tau|, :283) and
    # an approximate slope (:266-270): the iterate it stops at is what is restated, not the root of the law
    ut = t.get_shear_velocity(20.0, 0.5)
    assert ut * dist / nu > 100 and 0.3 < ut < 1.5
    assert t.get_shear_velocity(20.0, 0.0) == t.get_shear_velocity(20.0, 1e-9)  # the initial guess is floored at 5 nu / y
