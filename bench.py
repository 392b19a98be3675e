#!/usr/bin/env python
"""bench.py - BASELINE.json's metric on BASELINE.json's config.

Metric : wall seconds per time step of the 3-D INS lid-driven cavity "256^3" (= 128^3 hex cells, 257^3
         velocity nodes, Q2/Q1, 53 070 468 DoF, 1.13e10 matrix entries; SURVEY.md 8 "config 3") through the
         whole hot path - Newton iterations of { cell-loop assembly -> FGMRES + block-Schur preconditioner },
         plus the achieved HBM GB/s of the dominant kernel (the velocity-block SpMV) against the measured
         copy peak.
A step : one call of Fluid::MPI::InsIM::run_one_step (one time step = ~3 Newton iterations).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--cells C] [--impl ours|reference]

N = 1 : config 3 on one B200.  N > 1 : the same mesh (strong scaling) split into z-slabs, one rank per GPU,
launched by torch.distributed.run; NCCL carries only ghost-DoF halos and Krylov dot products.
`--impl reference` times the reference's CPU path: /root/reference cannot be built (deal.II / PETSc /
p4est absent), so this arm runs oracle/ - the CPU restatement of the same algorithm - on all host cores on
bounded samples of the same workload (one time step at 24^3, 32^3 and 48^3 cells), fits the growth exponent of
s/step in the number of cells and extrapolates to 128^3 with it (the line says so: same_config false).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "time_step_wall_s (3D INS cavity 128^3 cells Q2/Q1, 53.07M DoF)"
METRIC2 = "time_step_wall_s (3D flow past cylinder, InsIM Q2/Q1, ~1.35M DoF)"
METRIC4 = "time_step_wall_s (fsi_leaflet_mpi 2D: SCnsIM Q1/Q1 on the band-refined channel + NeoHookean leaflet, full IFEM step)"
METRIC5 = "time_step_wall_s (fsi-wall-3D: SCnsIM Q1/Q1 ~10M fluid DoF on the band-refined box + NeoHookean plate, full IFEM step)"
UNIT = "s/step"
CPU_BASELINE_CELLS = [16, 24]  # cpu_baseline leg of the default run (bounded); --impl reference: --ref-cells


def _peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def _ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch at config 3 on one GPU, from the committed
    `ncu --set full` captures (profiles/ncu_traffic.json names the capture each number comes from)."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            return {k: v["bytes_per_launch"] for k, v in json.load(f).items()}
    except Exception:
        return {}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.samples, self.stop_flag, self.proc = index, [], False, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE, text=True)
        except Exception:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            parts = [p.strip() for p in s.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons)}


CYL_H, CYL_UMAX = 0.41, 9 * 0.2 / 4


def cylinder_inflow(p, c, t):
    """3-D Schaefer-Turek inflow profile on x = -0.3 (boundary id 0), see tests/test_zz_config2_gpu.py"""
    return 16 * CYL_UMAX * p[1] * (CYL_H - p[1]) * p[2] * (CYL_H - p[2]) / CYL_H ** 4 if c == 0 and abs(p[0] + 0.3) < 1e-10 else 0.0


def cylinder_prm_path():
    return os.path.join(ROOT, "tests", "golden", "ins_cylinder_3d.prm")


def cpu_cylinder_fit(levels, target_level, threads=None):
    """config 2 on the CPU: one time step of the oracle on the 3-D cylinder mesh at each refinement level of `levels`, extrapolated to
    target_level with the fitted exponent of s/step in the number of cells (8x cells per level)."""
    import math

    import openifem_b200 as ifem
    from oracle import grid, ins as oracle_ins, prm

    team = oracle_ins.set_threads(threads)
    secs, cells = [], []
    for lv in levels:
        tria = ifem.Triangulation(3)  # host-side mesh generator of the product (no device work): the same mesh arrays as the GPU arm
        ifem.GridCreator.flow_around_cylinder(tria)
        tria.refine_global(lv)
        v, c, b = tria.get_mesh()
        o = oracle_ins.InsIM(grid.HexMesh(v, c, b), prm.Params(cylinder_prm_path()), mode="mpi", hard_coded={0: cylinder_inflow},
                             a_inv=("bicgstab", 1e-1, 2000))
        t0 = time.perf_counter()
        o.run_one_step(True)
        secs.append(time.perf_counter() - t0)
        cells.append(c.shape[0])
    p = math.log(secs[-1] / secs[0]) / math.log(cells[-1] / cells[0]) if len(levels) > 1 else 1.0
    target_cells = 832 * 8 ** target_level
    value = secs[-1] * (target_cells / cells[-1]) ** p
    desc = ("one time step (from rest, hard-coded inflow) of the 3-D cylinder case at "
            + ", ".join(f"refinement {lv} ({n} cells): {t:.1f} s" for lv, n, t in zip(levels, cells, secs))
            + f" with oracle/ on an OpenMP team of {team} threads; fitted s/step ~ cells^{p:.3f}; extrapolated to refinement {target_level} "
              f"({target_cells} cells)")
    return value, desc, team, secs, p


def cpu_reference_step(cells, steps=1, warmup=0, threads=None):
    """The oracle (CPU restatement of mpi_insim.cpp) on all host cores: seconds per time step on a cells^3 cavity, same .prm and
    same solver settings as the GPU arm (A~^-1 = BiCGStab + node-block Jacobi to 1e-1). Returns (s/step, oracle, team size)."""
    from oracle import ins as oracle_ins
    from util import cavity_prm, make_oracle

    team = oracle_ins.set_threads(threads)  # torchrun exports OMP_NUM_THREADS=1: ask the runtime explicitly
    o = make_oracle(cavity_prm(3), (cells,) * 3, (0, 0, 0), (1, 1, 1), a_inv=("bicgstab", 1e-1, 2000))
    k = 0
    for _ in range(warmup):
        o.run_one_step(k == 0)
        k += 1
    t0 = time.perf_counter()
    for _ in range(steps):
        o.run_one_step(k == 0)
        k += 1
    return (time.perf_counter() - t0) / steps, o, team


def cpu_reference_fit(sizes, target_cells, threads=None):
    """Time one step at each sample size, fit s/step = c * cells^(3 p) by least squares in log space and extrapolate to
    target_cells^3. Returns (value, description, team size, per-size seconds, p)."""
    import math

    secs, team = [], 0
    for c in sizes:
        sec, o, team = cpu_reference_step(c, threads=threads)
        secs.append(sec)
        del o
    xs = [3.0 * math.log(c) for c in sizes]
    ys = [math.log(t) for t in secs]
    if len(sizes) > 1:
        mx, my = sum(xs) / len(xs), sum(ys) / len(ys)
        p = sum((x - mx) * (y - my) for x, y in zip(xs, ys)) / sum((x - mx) ** 2 for x in xs)
    else:
        p = 1.0
    value = secs[-1] * (target_cells / sizes[-1]) ** (3.0 * p)
    desc = ("one time step (step 1 from rest: 3 Newton iterations, nothing cached) of the same cavity .prm at "
            + ", ".join(f"{c}^3 cells: {t:.1f} s" for c, t in zip(sizes, secs))
            + f" with oracle/ (CPU restatement of mpi_insim.cpp, same inner solvers and tolerances as the GPU arm) on an OpenMP team of "
              f"{team} threads; fitted s/step ~ cells^{p:.3f}; value = {secs[-1]:.1f} s x ({target_cells}/{sizes[-1]})^(3 x {p:.3f}) "
              f"- an extrapolation, the {target_cells}^3 system (135 GB of CSR) is not run on the CPU")
    return value, desc, team, secs, p


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    os.environ["OMP_NUM_THREADS"] = str(cores)  # unconditional: torchrun sets it to 1 for N > 1
    if args.config == 2:
        sizes = [0, 1]
        value, sample, team, secs, p = cpu_cylinder_fit(sizes, args.refine, threads=cores)
    else:
        sizes = [int(c) for c in args.ref_cells.split(",")]
        value, sample, team, secs, p = cpu_reference_fit(sizes, args.cells, threads=cores)
    line = {
        "impl": "reference", "metric": METRIC if args.config == 3 else METRIC2, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": value * 1e3, "higher_is_better": False, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": (f"3D INS lid-driven cavity {args.cells}^3 hex cells Q2/Q1 (config 3), one time step" if args.config == 3 else
                                f"3D flow past a cylinder, InsIM Q2/Q1, Global refinements = {args.refine} (config 2), one time step"),
                   "same_config": False, "extrapolated": True, "sample_cells": sizes, "sample_seconds": secs, "fitted_exponent": p},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": team, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------------------------
# BASELINE configs 4 and 5: the immersed FSI step (find_solid_bc -> solid step -> solid box -> indicator -> constraints ->
# find_fluid_bc -> fluid step; reference source/mpi_fsi.cpp:1172-1214)
# ---------------------------------------------------------------------------------------------------------------------------
LEAF_L, LEAF_H, LEAF_A, LEAF_B, LEAF_h, LEAF_U = 4.0, 1.0, 0.1, 0.4, 0.05, 1.5


def leaflet_inflow(p, c, t):
    return LEAF_U if c == 0 and abs(p[0]) < 1e-10 else 0.0


def fsi_prm_path(config):
    return os.path.join(ROOT, "tests", "golden", "fsi_leaflet_2d.prm" if config == 4 else "fsi_wall_3d.prm")


def fsi_meshes(config, scale, solid_scale, half=False):
    """(fluid triangulation, solid triangulation) of tests/fsi_leaflet_mpi/fsi_leaflet_mpi.cpp:47-92 (config 4; scale divides h)
    or tests/fsi-wall-3D/fsi-wall-3D.cpp:33-57 (config 5; scale multiplies the {10,10,40} fluid subdivisions, solid_scale the
    {20,20,8} solid ones) - host-side mesh generators of the product, the same arrays feed the oracle in the CPU legs"""
    import numpy as np

    import openifem_b200 as ifem

    if config == 4:
        h = LEAF_h / scale
        ftria = ifem.Triangulation(2)
        ifem.GridGenerator.subdivided_hyper_rectangle(ftria, (int(round(LEAF_L / h)), int(round(LEAF_H / h))), (0, 0), (LEAF_L, LEAF_H), True)
        v, c, _ = ftria.get_mesh()
        cx = v[c].mean(axis=1)[:, 0]
        ftria.execute_refinement(((cx >= LEAF_L / 4 - 2 * LEAF_A) & (cx <= LEAF_L / 4 + 3 * LEAF_A)).astype(np.uint8))
        stria = ifem.Triangulation(2)
        ifem.GridGenerator.subdivided_hyper_rectangle(stria, (int(round(LEAF_A / LEAF_h)), int(round(LEAF_B / LEAF_h))), (LEAF_L / 4, 0),
                                                      (LEAF_A + LEAF_L / 4, LEAF_B), True)
        stria.refine_global(2)  # Global refinements = 0, 2
        return ftria, stria
    ftria = ifem.Triangulation(3)
    if half:  # CPU sample below the reference's own resolution: {5,5,20} fluid and {10,10,4} solid subdivisions
        fr, sr = (5, 5, 20), (10, 10, 4)
    else:
        fr, sr = (10 * scale, 10 * scale, 40 * scale), (20 * solid_scale, 20 * solid_scale, 8 * solid_scale)
    ifem.GridGenerator.subdivided_hyper_rectangle(ftria, fr, (0, 0, 0), (1, 1, 4), True)
    v, c, _ = ftria.get_mesh()
    cz = v[c].mean(axis=1)[:, 2]
    ftria.execute_refinement(((cz >= 2) & (cz <= 2.4)).astype(np.uint8))
    stria = ifem.Triangulation(3)
    ifem.GridGenerator.subdivided_hyper_rectangle(stria, sr, (0, 0, 2), (1, 1, 2.4), True)
    return ftria, stria


def cpu_fsi_step(config, scale, solid_scale, steps=1, half=False):
    """the oracle's coupled loop (oracle/fsi.py on oracle/scns.py + oracle/solid.py: NumPy / C restatement, one core for the Python
    parts) on the same meshes: seconds per IFEM step"""
    import numpy as np

    from oracle import fem, fsi, grid, prm, scns, solid

    ftria, stria = fsi_meshes(config, scale, solid_scale, half)
    P = prm.Params(fsi_prm_path(config))
    v, c, b = ftria.get_mesh()
    sv, sc, _ = stria.get_mesh()
    if config == 4:
        o_fluid = scns.SCnsIM(grid.QuadMesh(v, c, b), P, hard_coded={0: leaflet_inflow})
        n = (4 * int(round(LEAF_A / LEAF_h)), 4 * int(round(LEAF_B / LEAF_h)))
        o_solid = solid.HyperElasticity(fem.BoxMesh(n, (LEAF_L / 4, 0), (LEAF_A + LEAF_L / 4, LEAF_B)), P)
    else:
        o_fluid = scns.SCnsIM(grid.HexMesh(v, c, b), P)
        sr = (10, 10, 4) if half else (20 * solid_scale, 20 * solid_scale, 8 * solid_scale)
        o_solid = solid.HyperElasticity(fem.BoxMesh(sr, (0, 0, 2), (1, 1, 2.4)), P)
    loop = fsi.FSI(o_fluid, o_solid, config == 4)
    t0 = time.perf_counter()
    for k in range(steps):
        loop.run_one_step(k == 0)
    return (time.perf_counter() - t0) / steps, c.shape[0], sc.shape[0]


def cpu_fsi_baseline(config, scale, solid_scale, full=False):
    """(s/step, description, cores). full = the reference arm: also the reference's own resolution of fsi-wall-3D (about two minutes)"""
    import math

    if config == 4:
        sec, nf, ns = cpu_fsi_step(4, 1, 1, steps=2)
        return sec, (f"two IFEM steps of the fsi_leaflet_mpi case itself ({nf} fluid cells, {ns} solid cells) with oracle/ (fsi.py loop in Python / "
                     f"NumPy, cell loops in C, sparse direct solves): {sec:.1f} s/step on one core"), 1
    target = 6800 * scale ** 3
    sec, nf, ns = cpu_fsi_step(5, 1, 1, steps=1, half=True)
    text = (f"one IFEM step of fsi-wall-3D at half the reference's resolution ({nf} fluid cells, {ns} solid cells): {sec:.1f} s")
    if full:
        sec2, nf2, ns2 = cpu_fsi_step(5, 1, 1, steps=1, half=False)
        p = math.log(sec2 / sec) / math.log(nf2 / nf)
        text += (f", at the reference's own resolution ({nf2} fluid cells, {ns2} solid cells): {sec2:.1f} s (s/step ~ cells^{p:.2f}: the reference's "
                 f"brute-force 3-D point_in_solid grows with fluid cells x solid cells)")
        sec, nf = sec2, nf2
    return sec * target / nf, (text + f" with oracle/ (fsi.py loop in Python / NumPy, cell loops in C, sparse direct solves) on one core; value = the "
                               f"larger sample x {target}/{nf} fluid cells - a LINEAR extrapolation in the fluid cells, optimistic for the CPU; "
                               f"it says nothing about a real MPI run of the reference"), 1


def run_fsi_reference(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    value, sample, cores = cpu_fsi_baseline(args.config, args.scale, args.solid_scale, full=True)
    line = {"impl": "reference", "metric": METRIC4 if args.config == 4 else METRIC5, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": value * 1e3, "higher_is_better": False, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": fsi_workload(args), "same_config": args.config == 4, "extrapolated": args.config != 4},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# inner T_pp solve on fp32 copies of the blocks: off until measured on hardware
TPP_FP32_DEFAULT = {4: 0, 5: 0}


def fsi_workload(args):
    if args.config == 4:
        return ("fsi_leaflet_mpi (config 4): MPI::FSI<2>(SCnsIM Q1/Q1 on the 80 x 20 channel with the band 0.8 <= x <= 1.3 refined once, "
                "SharedHyperElasticity NeoHookean leaflet 8 x 32 cells, use_dirichlet_bc = true), reference .prm verbatim, dt 5e-3"
                + (f", fluid h divided by {args.scale}" if args.scale != 1 else ""))
    return (f"fsi-wall-3D (config 5): MPI::FSI<3>(SCnsIM Q1/Q1 on {{10,10,40}} x {args.scale} cells of [0,1]^2 x [0,4] with the band 2 <= z <= 2.4 "
            f"refined once, SharedHyperElasticity NeoHookean plate {{20,20,8}} x {args.solid_scale} cells), reference .prm (solid type NeoHookean), dt 1e-6")


def run_fsi(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    import openifem_b200 as ifem

    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    rank, world = ifem.init_distributed(local_rank)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce(x, op):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=op)
        return t.item()

    t_setup = time.perf_counter()
    ftria, stria = fsi_meshes(args.config, args.scale, args.solid_scale)
    params = ifem.Parameters.AllParameters(fsi_prm_path(args.config))
    fluid = ifem.Fluid.MPI.SCnsIM(ftria, params)
    if args.config == 4:
        fluid.add_hard_coded_boundary_condition(0, leaflet_inflow)
    fluid.setup()
    tpp_fp32 = args.tpp_fp32 if args.tpp_fp32 >= 0 else TPP_FP32_DEFAULT[args.config]
    if tpp_fp32:
        fluid.set_control(a_inv_fp32=1)  # SUPG solvers: the inner T_pp solve streams fp32 copies of A_vp / A_pv / A_pp
    solid = ifem.Solid.MPI.SharedHyperElasticity(stria, params)
    solid.setup()
    coupling = ifem.MPI.FSI(fluid, solid, params, args.config == 4)
    barrier()
    t_setup = time.perf_counter() - t_setup
    n_cells, n_scells = ftria.n_active_cells(), stria.n_active_cells()
    dim = 2 if args.config == 4 else 3
    n_u, n_p, nnz_local, _, _ = fluid.sizes()
    ou, op = fluid.partition(0)[0], fluid.partition(1)[0]
    n_dofs = int(reduce(dim * ou + op, dist.ReduceOp.SUM))
    nnz = int(reduce(nnz_local, dist.ReduceOp.SUM))
    n_local = n_u + n_p
    host = torch.zeros(n_local, dtype=torch.float64).pin_memory()
    host_np = host.numpy()
    sections = ["Find solid BC", "Run solid solver", "Update solid box", "Update indicator", "Find fluid BC", "Run fluid solver"]
    fsections = ["Assemble system", "Solve linear system", "Solving Tpp"]

    step_no = 0
    for _ in range(args.warmup):
        coupling.run_one_step(step_no == 0)
        step_no += 1
    host_np[:] = fluid.get_current_solution()
    t_before = {k: coupling.timer_ms(k) for k in sections}
    f_before = {k: fluid.timer_ms(k) for k in fsections}
    n_hist = len(fluid.history())
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = ifem.kernel_launches()
    barrier()
    t_e2e = t_dev_ms = 0.0
    for _ in range(args.steps):
        t0 = time.perf_counter()
        fluid.set_vector(fluid.PRESENT, host_np)                     # H2D of the fluid state the step starts from (pinned)
        t_dev_ms += coupling.bench_steps(1, step_no == 0)            # CUDA events on the library stream around the coupled pass
        host_np[:] = fluid.get_current_solution()                    # D2H of the step's result
        disp = solid.get_current_solution()
        torch.cuda.synchronize()
        t_e2e += time.perf_counter() - t0
        step_no += 1
    barrier()
    launches = ifem.kernel_launches() - launches0
    clocks = sampler.stop()
    sec_dev = reduce(t_dev_ms * 1e-3 / args.steps, dist.ReduceOp.MAX)
    sec_e2e = reduce(t_e2e / args.steps, dist.ReduceOp.MAX)
    h2d = int(reduce(n_local * 8, dist.ReduceOp.SUM))
    d2h = h2d + disp.size * 8
    sec = {k: (coupling.timer_ms(k) - t_before[k]) / args.steps for k in sections}
    fsec = {k: (fluid.timer_ms(k) - f_before[k]) / args.steps for k in fsections}
    hist = fluid.history()[n_hist:]
    ms_blk, bytes_blk = fluid.bench_vmult(20)
    ms_blk = reduce(ms_blk, dist.ReduceOp.MAX)
    peak, peak_src = _peaks()
    sol = host_np
    u_own, p_own = sol[:dim * ou], sol[n_u:n_u + op]
    su2 = reduce(float(np.dot(u_own, u_own)), dist.ReduceOp.SUM)
    sp2 = reduce(float(np.dot(p_own, p_own)), dist.ReduceOp.SUM)
    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return
    cpu = None
    if not args.no_cpu_baseline and world == 1:
        v, sample, cores = cpu_fsi_baseline(args.config, args.scale, args.solid_scale)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}
    nv = 1 << dim
    # FSI kernels: algorithmic work per pass: update_indicator tests the 2^d vertices of every local fluid cell
    # (2^d d 8 B of coordinates in, 4 B out per cell), find_fluid_bc visits every velocity node of the indicator-1 cells
    ind_ms = max(sec["Update indicator"], 1e-9)
    fsi_k = {"update_indicator": {"ms": sec["Update indicator"], "points_per_s": n_cells / world * nv / (ind_ms * 1e-3),
                                  "algorithmic_bytes": n_cells / world * (nv * dim * 8 + 4),
                                  "GB/s": n_cells / world * (nv * dim * 8 + 4) / (ind_ms * 1e-3) / 1e9,
                                  "note": "bounding-box reject, then binned solid-cell search (3-D) / crossing number (2-D); latency bound at these sizes"},
             "find_fluid_bc": {"ms": sec["Find fluid BC"]}, "find_solid_bc": {"ms": sec["Find solid BC"]},
             "update_solid_box": {"ms": sec["Update solid box"]}}
    line = {
        "metric": METRIC4 if args.config == 4 else METRIC5, "value": sec_dev, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": sec_dev * 1e3, "higher_is_better": False, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": fsi_workload(args) + f": {n_cells} fluid cells, {n_dofs} fluid DoF, {nnz} matrix entries, {n_scells} solid cells, "
                                                    f"{solid.n_dofs} solid DoF (replicated on every rank)",
                   "l2": ("inputs larger than L2 (the fluid matrix is %.2f GB per GPU)" % (bytes_blk / 1e9)) if bytes_blk > 126e6 else
                         "the whole problem fits in L2 (%.1f MB of matrix): a launch-latency-bound step, the reference's own size" % (bytes_blk / 1e6),
                   "parallelism": f"{world} slab(s) of the fluid along mesh planes, one rank per GPU; the solid is replicated; NCCL: ghost halos, "
                                  "dot products and the sum of the solid-side interpolation",
                   "setup_s": round(t_setup, 1),
                   "tpp_fp32": int(tpp_fp32),
                   "section_ms_per_step": {**sec, **{"fluid: " + k: v for k, v in fsec.items()}},
                   "newton_its_per_step": len(hist) / max(1, args.steps),
                   "fgmres_its": [h["gmres_its"] for h in hist], "inner_tpp_its": [h["a_inv_its"] for h in hist]},
        "e2e": {"value": sec_e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": launches, "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": "block SpMV of the SCnsIM system (FGMRES operator: bcsr_spmv_row_kernel on the four Q1 blocks), per GPU",
                     "achieved": bytes_blk / (ms_blk * 1e-3) / 1e9, "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                     "frac": bytes_blk / (ms_blk * 1e-3) / 1e9 / peak, "traffic": None, "algorithmic_bytes": bytes_blk, "ms": ms_blk,
                     "csr_equivalent_bytes_per_gpu": 12.0 * nnz_local + 20.0 * (dim * ou + op), "fsi_kernels": fsi_k},
        "cpu_baseline": cpu,
        "parity_pins": {"u_l2": su2 ** 0.5, "p_l2": sp2 ** 0.5, "solid_u_max": float(np.abs(disp).max()),
                        "final_newton_abs_res": hist[-1]["abs_res"] if hist else None,
                        "note": "field norms over all owned dofs after the last timed step; tests/test_zz_config4_gpu.py and test_fsi_gpu.py compare "
                                "the same loop with the oracle to 1e-6"},
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()



def run_ours(args):
    import numpy as np
    import torch
    from util import cavity_prm

    import openifem_b200 as ifem

    import torch.distributed as dist

    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    rank, world = ifem.init_distributed(local_rank)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return t.item()

    n = args.cells
    t_setup = time.perf_counter()
    tria = ifem.Triangulation(3)
    if args.config == 2:
        ifem.GridCreator.flow_around_cylinder(tria)
        tria.refine_global(args.refine)
        params = ifem.Parameters.AllParameters(cylinder_prm_path())
        flow = ifem.Fluid.MPI.InsIM(tria, params)
        flow.add_hard_coded_boundary_condition(0, cylinder_inflow)
    else:
        ifem.GridGenerator.subdivided_hyper_rectangle(tria, (n, n, n), (0, 0, 0), (1, 1, 1), True)
        params = ifem.Parameters.AllParameters(text=cavity_prm(3))
        flow = ifem.Fluid.MPI.InsIM(tria, params)
    flow.setup()
    n_cells_global = tria.n_active_cells()
    # fp32 inner solver on the SELL-32 copy of A_uu with row-scaled fp16 matrix values (preconditioner only)
    flow.set_control(a_inv_rel=1e-1, a_inv_fp32=args.inner_mode, cg_sm_fp32=args.sm_mode, a_inv_max_it=400)
    barrier()
    t_setup = time.perf_counter() - t_setup
    n_u, n_p, nnz_local, _, _ = flow.sizes()
    n_local = n_u + n_p  # local block vector (owned + ghost entries) exchanged with the host
    ou, op = flow.partition(0)[0], flow.partition(1)[0]
    n_dofs = int(sum_over_ranks(3 * ou + op))
    nnz = int(sum_over_ranks(nnz_local))
    host = torch.zeros(n_local, dtype=torch.float64).pin_memory()
    host_np = host.numpy()

    step_no = 0
    for _ in range(args.warmup):
        flow.run_one_step(step_no == 0)
        step_no += 1
    host_np[:] = flow.get_current_solution()

    SECTIONS = ["Assemble system", "Solve linear system", "CG for Mp", "CG for Sm", "A_inv", "CG for Sm fp64 fallbacks (count)",
                "A_inv block-Jacobi fallbacks (count)"]
    sections_before = {k: flow.timer_ms(k) for k in SECTIONS}
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = ifem.kernel_launches()
    barrier()
    t_e2e = t_dev_ms = 0.0
    for _ in range(args.steps):
        t0 = time.perf_counter()
        flow.set_vector(flow.PRESENT, host_np)                      # H2D of the step's input (pinned)
        t_dev_ms += flow.bench_steps(1, step_no == 0)               # CUDA events on the library stream
        host_np[:] = flow.get_current_solution()                    # D2H of the step's result
        torch.cuda.synchronize()
        t_e2e += time.perf_counter() - t0
        step_no += 1
    barrier()
    launches = ifem.kernel_launches() - launches0
    clocks = sampler.stop()
    sec_dev = max_over_ranks(t_dev_ms * 1e-3 / args.steps)
    sec_e2e = max_over_ranks(t_e2e / args.steps)
    h2d = int(sum_over_ranks(n_local * 8))

    # kernels timed live on the library stream (matrix >> L2, so every launch streams from HBM):
    #  * the DOMINANT kernel of a step: product kernel of the fp32 inner A~^-1 solves on the SELL-32 copy of A_uu
    #    (~70 % of a step)
    #  * the FGMRES operator SpMV on the same block in fp64 (north_star's ">= 40 % of HBM roofline" kernel)
    ms_uu, bytes_uu = flow.bench_spmv_uu(20)
    if args.inner_mode >= 2:
        ms_32, bytes_32, sell_padding, sell_err = flow.bench_spmv_uu_sell(20, check_error=False)
    else:
        (ms_32, bytes_32), sell_padding = (flow.bench_spmv_uu_fp32(20) if args.inner_mode == 1 else (ms_uu, bytes_uu)), 1.0
    ms_blk, bytes_blk = flow.bench_vmult(10)
    # cell-loop assembly (InsIM::assemble, all colours + Neumann faces + zeroing of the matrices), timed live; algorithmic work per
    # cell: 27^3 (row node, column node, quadrature point) triples x 25 FMA for the velocity-velocity blocks + ~3e4
    # FMA for the other phases; 89^2 matrix entries read-modify-written (16 B each)
    flow.bench_assemble(1)
    ms_asm = flow.bench_assemble(3)
    n_cells_local = n_cells_global // world  # cells a rank owns (the duplicated interface layer of the owner-computes scheme is overhead)
    asm_flop = n_cells_local * (27 ** 3 * 25 * 2 + 60000.0)
    asm_bytes = n_cells_local * 89 * 89 * 16.0
    fp64_peak = ifem.bench_fp64_peak()
    peak, peak_src = _peaks()
    achieved = bytes_32 / (ms_32 * 1e-3) / 1e9
    achieved64 = bytes_uu / (ms_uu * 1e-3) / 1e9
    traffic = _ncu_traffic()
    hist = flow.history()
    last = [h for h in hist if h["timestep"] == hist[-1]["timestep"]]
    sections = {k: flow.timer_ms(k) for k in SECTIONS}  # totals since set-up (warm-up included)
    sections_timed = {k: (sections[k] - sections_before[k]) / (1 if "count" in k else args.steps) for k in SECTIONS}
    # parity pins, computed outside the timed region in fp64: true residual |b - A x| / |b| of every linear solve of the timed
    # steps (recomputed with the fp64 operator), the Newton residual the reference prints (mpi_insim.cpp:456-460), and norms of
    # the final fields summed over the owned dofs of all ranks - comparable across N = 1 / 2 / 4 / 8 and with an all-fp64 run
    sol = host_np  # solution after the last timed step
    u_own, p_own = sol[:3 * ou], sol[n_u:n_u + op]
    su2, sp1, sp2, cnt = (sum_over_ranks(float(v)) for v in (np.dot(u_own, u_own), p_own.sum(), np.dot(p_own, p_own), p_own.size))
    timed = [h for h in hist if h["timestep"] > args.warmup]
    pins = {"u_l2": su2 ** 0.5, "p_meanfree_l2": max(0.0, sp2 - sp1 * sp1 / cnt) ** 0.5,
            "final_newton_abs_res": last[-1]["abs_res"], "final_newton_rel_res": last[-1]["rel_res"],
            "max_true_res_timed_steps": max(h["true_res"] for h in timed) if timed else None,
            # solves whose tolerance was the relative one, 1e-4 |rhs| (the last Newton iteration of a step runs into the absolute floor 1e-12)
            "max_true_res_relative_tolerance_solves": max([h["true_res"] for h in timed if 1e-4 * h["abs_res"] > 1e-12] or [None]),
            "last_step": [{"newton_it": h["iteration"], "abs_res": h["abs_res"], "fgmres_its": h["gmres_its"], "fgmres_res": h["gmres_res"],
                           "true_res": h["true_res"], "a_inv_its": h["a_inv_its"], "cg_sm_its": h["cg_sm_its"], "cg_mp_its": h["cg_mp_its"]}
                          for h in last],
            "a_inv_block_jacobi_fallbacks": int(sections["A_inv block-Jacobi fallbacks (count)"]),
            "cg_sm_fp64_fallbacks": int(sections["CG for Sm fp64 fallbacks (count)"]),
            "note": "true_res = |b - A x|_2 / |b|_2 with the fp64 block operator after FGMRES returned (tolerance 1e-4); "
                    "field norms over all owned dofs after the last timed step"}

    if world > 1:
        ms_uu, ms_blk, ms_32, ms_asm = max_over_ranks(ms_uu), max_over_ranks(ms_blk), max_over_ranks(ms_32), max_over_ranks(ms_asm)
        achieved = bytes_32 / (ms_32 * 1e-3) / 1e9
        achieved64 = bytes_uu / (ms_uu * 1e-3) / 1e9
        dist.barrier()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    cpu = None
    if not args.no_cpu_baseline and world == 1:
        # bounded sample (about 20-30 s of CPU work); the reference arm (--impl reference) runs the larger sizes
        if args.config == 2:
            v, sample, team, _, _ = cpu_cylinder_fit([0, 1], args.refine, threads=os.cpu_count())
        else:
            v, sample, team, _, _ = cpu_reference_fit(CPU_BASELINE_CELLS, n, threads=os.cpu_count())
        cpu = {"value": v, "unit": UNIT, "cores": team, "kind": "port", "sample": sample}

    line = {
        "metric": METRIC if args.config == 3 else METRIC2, "value": sec_dev, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": sec_dev * 1e3, "higher_is_better": False, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": (f"3D INS lid-driven cavity {n}^3 hex cells Q2/Q1 (config 3): {n_dofs} DoF, {nnz} matrix entries, "
                                f"Re 100, dt 1e-2, from rest; step = run_one_step (Newton x (assembly + FGMRES/Schur))" if args.config == 3 else
                                f"3D flow past a cylinder (config 2): GridCreator<3>::flow_around_cylinder, Global refinements = {args.refine}, "
                                f"{n_cells_global} cells, {n_dofs} DoF, {nnz} matrix entries, InsIM Q2/Q1, hard-coded parabolic inflow, mu 1e-3, "
                                f"gamma 0.1, dt 1e-2, from rest; step = run_one_step"),
                   "l2": "inputs larger than L2 (A_uu alone is %.1f GB per GPU)" % (bytes_uu / 1e9),
                   "a_inv": {0: "A~^-1 = BiCGStab(node-block Jacobi) to 1e-1 in fp64 on the BCSR matrix",
                             1: "A~^-1 = BiCGStab(node-block Jacobi) to 1e-1, A_uu streamed as fp32 (BCSR)",
                             2: "A~^-1 = BiCGStab(node-block Jacobi) to 1e-1 in fp32 on a sliced-ELL (SELL-32) copy of A_uu",
                             3: "A~^-1 = BiCGStab(node-block Jacobi) to 1e-1 in fp32 on a sliced-ELL (SELL-32) copy of A_uu "
                                "whose values are stored as row-scaled fp16"}[args.inner_mode]
                            + {0: "; CG for Sm in fp64", 1: "; CG for Sm in fp32 on a SELL-32 copy of S_m",
                               2: "; CG for Sm in fp32 on a SELL-32 copy of S_m with row-scaled fp16 values"}[args.sm_mode]
                            + "; inside the preconditioner only - FGMRES operator, residuals and basis are fp64, Newton/FGMRES "
                              "iteration counts and converged fields equal the fp64 path's (tests/test_inner32_gpu.py: 1e-6 vs oracle)",
                   "parallelism": f"{world} z-slab(s), one rank per GPU; NCCL: ghost halos + dot-product all-reduces only",
                   "setup_s": round(t_setup, 1), "setup_note": "mesh, patterns, partition and the SELL-32 layouts are built once before the "
                   "warm-up steps and are outside `value`; the per-solve refresh of the SELL copies from the assembled matrix is inside", "newton_its_last_step": len(last),
                   "fgmres_its_last_step": [h["gmres_its"] for h in last], "section_ms_total": sections,
                   "section_ms_per_timed_step": sections_timed},
        "e2e": {"value": sec_e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": h2d},
        "gpu_launches": launches, "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": ("sell_spmv_h_kernel<3,2,4>" if args.inner_mode == 3 else "sell_spmv_pipe_kernel<3,2,4>")
                                               + " (product of the fp32 inner A~^-1 solves on the SELL-32 copy of A_uu; dominant "
                                                 "kernel of a step), per GPU (rank 0's rows)",
                     "achieved": achieved, "peak": peak, "peak_source": peak_src, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic.get("sell_spmv_h_kernel" if args.inner_mode == 3 else "sell_spmv_pipe_kernel") if world == 1 and n == 128 and args.config == 3 else None,
                     "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture of this kernel at this "
                                       "config (profiles/ncu_traffic.json names the capture); not re-measured in this run",
                     "algorithmic_bytes": bytes_32, "ms": ms_32, "sell_padding": sell_padding,
                     "fgmres_operator_spmv": {"kernel": "bcsr_spmv_row_kernel<3,3,32,double,1,4,0> (A_uu, fp64 operator of FGMRES)", "ms": ms_uu,
                                              "algorithmic_bytes": bytes_uu, "achieved": achieved64, "frac": achieved64 / peak,
                                              "traffic": traffic.get("bcsr_spmv_kernel<3,3,32,double>") if world == 1 and n == 128 and args.config == 3 else None},
                     "assembly": {"kernel": "ins_assemble_kernel<3> (team of 3 warps per cell, 8 colour launches) + zeroing + Neumann faces",
                                  "ms": ms_asm, "bound": "fp64 + hbm (matrix read-modify-write)", "algorithmic_flop": asm_flop,
                                  "achieved_tflops": asm_flop / (ms_asm * 1e-3) / 1e12, "fp64_peak_tflops": fp64_peak,
                                  "fp64_peak_source": "measured in this run (register-only FMA chains, ifem_bench_fp64_peak)",
                                  "frac_fp64": asm_flop / (ms_asm * 1e-3) / 1e12 / fp64_peak, "rmw_bytes": asm_bytes,
                                  "rmw_GBps": asm_bytes / (ms_asm * 1e-3) / 1e9, "frac_hbm": asm_bytes / (ms_asm * 1e-3) / 1e9 / peak},
                     "block_vmult": {"ms": ms_blk, "bytes": bytes_blk, "GB/s": bytes_blk / (ms_blk * 1e-3) / 1e9,
                                     "csr_equivalent_bytes_per_gpu": 12.0 * nnz_local + 20.0 * (3 * ou + op)}},
        "cpu_baseline": cpu, "parity_pins": pins,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--cells", type=int, default=128, help="cells per direction (config 3 = 128)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=3, choices=[2, 3, 4, 5],
                    help="BASELINE config: 3 = 3-D cavity 128^3 (the metric's config, default), 2 = 3-D flow past a cylinder (~1.35 M dofs, 1 GPU), "
                         "4 = fsi_leaflet_mpi (2-D FSI, 1 GPU), 5 = fsi-wall-3D scaled to ~10 M fluid dofs (3-D FSI, 8 GPUs)")
    ap.add_argument("--scale", type=int, default=0, help="configs 4 / 5: refinement factor of the fluid mesh (default 1 for config 4, 7 for config 5)")
    ap.add_argument("--solid-scale", type=int, default=1, help="config 5: factor on the {20,20,8} solid subdivisions")
    ap.add_argument("--refine", type=int, default=2, help="config 2: Global refinements of the cylinder mesh")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--tpp-fp32", type=int, default=-1, choices=[-1, 0, 1],
                    help="configs 4 / 5: inner T_pp solve of the SUPG block preconditioner on fp32 copies of the blocks (-1: the config's default)")
    ap.add_argument("--ref-cells", default="24,32,48",
                    help="--impl reference: cells per direction of the timed samples (the fit is extrapolated to --cells)")
    ap.add_argument("--sm-mode", type=int, default=1, choices=[0, 1, 2],
                    help="'CG for Sm': 0 fp64 CG on CSR, 1 fp32 CG on SELL-32, 2 fp32 CG on fp16 SELL-32 values")
    ap.add_argument("--inner-mode", type=int, default=3, choices=[0, 1, 2, 3],
                    help="A~^-1 inner solve: 0 fp64 BCSR, 1 fp32-streamed BCSR, 2 fp32 SELL-32, 3 fp32 solver on fp16 SELL-32 values")
    args = ap.parse_args()
    if args.config in (4, 5):
        args.scale = args.scale or (1 if args.config == 4 else 7)
        (run_fsi_reference if args.impl == "reference" else run_fsi)(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
