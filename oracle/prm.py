"""ORACLE (test infrastructure, NOT product code).

Minimal reader for deal.II ParameterHandler `.prm` files with the semantics of
Parameters::AllParameters (reference source/parameters.cpp:7-658): defaults as
declared there, comma-separated lists, Dirichlet values consumed 1/2/3 per
component flag (parameters.cpp:209-238).
"""
from __future__ import annotations


def parse_prm(path_or_text: str, is_text: bool = False) -> dict:
    text = path_or_text if is_text else open(path_or_text).read()
    out: dict = {}
    stack = []
    for raw in text.splitlines():
        line = raw.split("#", 1)[0].strip()
        if not line:
            continue
        if line.startswith("subsection"):
            stack.append(line[len("subsection"):].strip())
        elif line == "end":
            stack.pop()
        elif line.startswith("set"):
            k, v = line[3:].split("=", 1)
            out[("/".join(stack), " ".join(k.split()))] = v.strip()
    return out


def _lst(s, f):
    return [f(x) for x in s.split(",") if x.strip()] if s is not None else []


class Params:
    def __init__(self, path_or_text: str, is_text: bool = False):
        d = parse_prm(path_or_text, is_text)
        self.raw = d
        g = lambda sec, key, default: d.get((sec, key), default)
        S = "Simulation"
        self.simulation_type = g(S, "Simulation type", "FSI")
        self.dimension = int(g(S, "Dimension", "2"))
        self.global_refinements = _lst(g(S, "Global refinements", ""), int)
        self.end_time = float(g(S, "End time", "1.0"))
        self.time_step = float(g(S, "Time step size", "1.0"))
        self.output_interval = float(g(S, "Output interval", "1.0"))
        self.refinement_interval = float(g(S, "Refinement interval", "1.0"))
        self.save_interval = float(g(S, "Save interval", "1.0"))
        self.gravity = _lst(g(S, "Gravity", ""), float)
        self.initial_velocity = _lst(g(S, "Initial velocity", ""), float)
        F = "Fluid finite element system"
        self.fluid_pressure_degree = int(g(F, "Pressure degree", "1"))
        self.fluid_velocity_degree = int(g(F, "Velocity degree", "2"))
        M = "Fluid material properties"
        self.viscosity = float(g(M, "Dynamic viscosity", "1e-3"))
        self.fluid_rho = float(g(M, "Fluid density", "1.0"))
        C = "Fluid solver control"
        self.grad_div = float(g(C, "Grad-Div stabilization", "1.0"))
        self.fluid_max_iterations = int(g(C, "Max Newton iterations", "8"))
        self.fluid_tolerance = float(g(C, "Nonlinear system tolerance", "1e-10"))
        D = "Fluid Dirichlet BCs"
        n = int(g(D, "Number of Dirichlet BCs", "0"))
        ids = _lst(g(D, "Dirichlet boundary id", ""), int)
        comps = _lst(g(D, "Dirichlet boundary components", ""), int)
        vals = _lst(g(D, "Dirichlet boundary values", ""), float)
        self.use_hard_coded_values = int(g(D, "Use hard-coded boundary values", "0"))
        self.fluid_dirichlet_bcs = {}
        k = 0
        for i in range(n):
            m = {1: 1, 2: 1, 4: 1, 3: 2, 5: 2, 6: 2, 7: 3}[comps[i]]
            self.fluid_dirichlet_bcs[ids[i]] = (comps[i], vals[k:k + m])
            k += m
        if n and k != len(vals):
            raise ValueError("Inconsistent boundary values!")
        N = "Fluid Neumann BCs"
        n = int(g(N, "Number of Neumann BCs", "0"))
        ids = _lst(g(N, "Neumann boundary id", ""), int)
        vals = _lst(g(N, "Neumann boundary values", ""), float)
        self.fluid_neumann_bcs = {ids[i]: vals[i] for i in range(n)}
        SF = "Solid finite element system"
        self.solid_degree = int(g(SF, "Degree", "1"))
        SM = "Solid material properties"
        self.solid_type = g(SM, "Solid type", "LinearElastic")
        self.n_solid_parts = int(g(SM, "Number of solid parts", "1"))
        self.solid_rho = float(g(SM, "Solid density", "1.0"))
        self.E = _lst(g(SM, "Young's modulus", "0.0"), float)
        self.nu = _lst(g(SM, "Poisson's ratio", "0.0"), float)
        self.eta = _lst(g(SM, "Viscosity", "0.0"), float)
        c = _lst(g(SM, "Hyperelastic parameters", ""), float)
        per = 2 if self.solid_type == "NeoHookean" else 1
        self.C = [c[i * per:(i + 1) * per] for i in range(self.n_solid_parts)] if c else []
        SS = "Solid solver control"
        self.damping = float(g(SS, "Damping", "0.0"))
        self.solid_max_iterations = int(g(SS, "Max Newton iterations", "8"))
        self.tol_d = float(g(SS, "Displacement tolerance", "1e-10"))
        self.tol_f = float(g(SS, "Force tolerance", "1e-10"))
        self.contact_force_multiplier = float(g(SS, "Contact force multiplier", "1e8"))
        SD = "Solid Dirichlet BCs"
        n = int(g(SD, "Number of Dirichlet BCs", "0"))
        ids = _lst(g(SD, "Dirichlet boundary id", ""), int)
        comps = _lst(g(SD, "Dirichlet boundary components", ""), int)
        self.solid_dirichlet_bcs = {ids[i]: comps[i] for i in range(n)}
        SN = "Solid Neumann BCs"
        n = int(g(SN, "Number of Neumann BCs", "0"))
        ids = _lst(g(SN, "Neumann boundary id", ""), int)
        self.solid_neumann_bc_type = g(SN, "Neumann boundary type", "Traction")
        vals = _lst(g(SN, "Neumann boundary values", ""), float)
        per = self.dimension if self.solid_neumann_bc_type == "Traction" else 1
        self.solid_neumann_bcs = {ids[i]: vals[i * per:(i + 1) * per] for i in range(n)}
