"""The two time-dependent SCnsIM cases of the reference (test infrastructure shared by the oracle goldens, the fixture
generator scripts/make_acoustic_fixture.py and the GPU tests):

  acoustic_duct_wave_mpi (tests/acoustic_duct_wave_mpi/acoustic_duct_wave_mpi.cpp:33-68): 4 x 1 duct, 8 x 2 cells refined 3
      times, Gaussian velocity pulse on x = 0 given as per-step increments through a hard-coded boundary function,
      1000 steps of 1e-7 s; golden: max velocity 5.93 +- 1e-3
  acoustic_pml_mpi (tests/acoustic_pml_mpi/acoustic_pml_mpi.cpp:33-85): 1.4 x 0.4 tube, 7 x 2 cells refined 3 times, the
      same pulse (100 x shorter) and a quartic PML over the last 1.2 of the tube, 500 steps; golden: |max velocity| < 5e-2
"""
import math
import os

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

CASES = {
    "duct": dict(prm="scns_acoustic_duct_2d.prm", reps=(8, 2), hi=(4.0, 1.0), t0=0.5e-4, width=0.15e-4, pml=False),
    "pml": dict(prm="scns_acoustic_pml_2d.prm", reps=(7, 2), hi=(1.4, 0.4), t0=0.5e-6, width=0.15e-6, pml=True),
}


def gaussian_pulse(case, dt):
    """the gaussian_pulse lambdas: increment of 6 exp(-((t - t0) / w)^2 / 2) over the step, on u_x at x = 0"""
    t0, w = CASES[case]["t0"], CASES[case]["width"]

    def time_value(t):
        return 6.0 * math.exp(-0.5 * ((t - t0) / w) ** 2)

    def f(p, component, time):
        if component == 0 and abs(p[0]) < 1e-10:
            return time_value(time) - (0.0 if time < 2 * dt else time_value(time - dt))
        return 0.0

    return f


def sigma_pml_field(p, component=0):
    """acoustic_pml_mpi.cpp:36-50: SigmaMax ((x + L_pml - boundary) / L_pml)^4 inside the layer"""
    pml_length, sigma_max, boundary = 1.2, 340000.0, 1.4
    if p[0] > boundary - pml_length:
        return sigma_max * ((p[0] + pml_length - boundary) / pml_length) ** 4
    return 0.0


def prm_text(case, n_steps=None):
    """the reference's parameter file; n_steps shortens the run (End time = n_steps dt)"""
    text = open(os.path.join(GOLDEN, CASES[case]["prm"])).read()
    if n_steps is not None:
        out = []
        for line in text.splitlines():
            if line.strip().startswith("set End time"):
                line = "  set End time = %.17g" % (n_steps * 1e-7)
            out.append(line)
        text = "\n".join(out) + "\n"
    return text


def make_oracle(case, n_steps=None):
    from oracle import fem, prm, scns

    c = CASES[case]
    p = prm.Params(prm_text(case, n_steps), is_text=True)
    mesh = fem.BoxMesh(c["reps"], (0, 0), c["hi"]).refine_global(p.global_refinements[0])
    return scns.SCnsIM(mesh, p, hard_coded={0: gaussian_pulse(case, p.time_step)},
                       sigma_pml_field=sigma_pml_field if c["pml"] else None)
