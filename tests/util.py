"""Shared helpers of the parity tests: build the same problem in the oracle (CPU) and
in the product (CUDA, through the C ABI)."""
import os

import numpy as np

PRM_TEMPLATE = """
subsection Simulation
  set Simulation type = Fluid
  set Dimension = {dim}
  set Global refinements = 0, 0
  set End time = {end_time}
  set Time step size = {dt}
  set Output interval = 1e6
  set Refinement interval = 1e6
  set Save interval = 1e6
  set Gravity = {gravity}
  set Initial velocity = {zeros}
end
subsection Fluid finite element system
  set Pressure degree = 1
  set Velocity degree = 2
end
subsection Fluid material properties
  set Dynamic viscosity = {mu}
  set Fluid density = {rho}
end
subsection Fluid solver control
  set Grad-Div stabilization = {gamma}
  set Max Newton iterations = {max_newton}
  set Nonlinear system tolerance = {newton_tol}
end
subsection Fluid Dirichlet BCs
  set Use hard-coded boundary values = 0
  set Number of Dirichlet BCs = {n_dir}
  set Dirichlet boundary id = {dir_ids}
  set Dirichlet boundary components = {dir_comps}
  set Dirichlet boundary values = {dir_vals}
end
subsection Fluid Neumann BCs
  set Number of Neumann BCs = {n_neu}
  set Neumann boundary id = {neu_ids}
  set Neumann boundary values = {neu_vals}
end
"""


def cavity_prm(dim, dt=1e-2, mu=0.01, rho=1.0, gamma=1.0, end_time=1.0, newton_tol=1e-6, max_newton=8, gravity=None,
               dirichlet=None, neumann=None):
    """Lid-driven cavity in the style of tests/fluid_cavity/fluid_cavity.prm: all walls no-slip, lid
    (last boundary id) moves with u_x = 1."""
    if dirichlet is None:
        full = 3 if dim == 2 else 7
        dirichlet = {i: (full, [0.0] * dim) for i in range(2 * dim)}
        dirichlet[2 * dim - 1] = (full, [1.0] + [0.0] * (dim - 1))
    neumann = neumann or {}
    gravity = gravity or [0.0] * dim
    ids = sorted(dirichlet)
    return PRM_TEMPLATE.format(
        dim=dim, zeros=", ".join(["0.0"] * dim), end_time=end_time, dt=dt, gravity=", ".join(str(g) for g in gravity), mu=mu, rho=rho, gamma=gamma,
        max_newton=max_newton, newton_tol=newton_tol, n_dir=len(ids), dir_ids=", ".join(str(i) for i in ids),
        dir_comps=", ".join(str(dirichlet[i][0]) for i in ids),
        dir_vals=", ".join(str(v) for i in ids for v in dirichlet[i][1]),
        n_neu=len(neumann), neu_ids=", ".join(str(i) for i in sorted(neumann)),
        neu_vals=", ".join(str(neumann[i]) for i in sorted(neumann)))


def make_oracle(prm_text, reps, lo, hi, mode="mpi", a_inv="lu"):
    from oracle import fem, ins, prm

    p = prm.Params(prm_text, is_text=True)
    mesh = fem.BoxMesh(tuple(reps), lo, hi)
    return ins.InsIM(mesh, p, mode=mode, a_inv=a_inv)


def make_gpu(prm_text, reps, lo, hi):
    import openifem_b200 as ifem

    dim = len(reps)
    tria = ifem.Triangulation(dim)
    ifem.GridGenerator.subdivided_hyper_rectangle(tria, reps, lo, hi, True)
    params = ifem.Parameters.AllParameters(text=prm_text)
    s = ifem.Fluid.MPI.InsIM(tria, params)
    s.setup()
    return s


def rel(a, b):
    d = np.linalg.norm(np.asarray(a) - np.asarray(b))
    n = np.linalg.norm(np.asarray(b))
    return d / n if n > 0 else d
