"""openifem_b200: B200-native hot path of OpenIFEM behind the reference's class surface.

Python mirror (over the C ABI, via ctypes) of the classes the reference's test
drivers use: `Triangulation` + `GridGenerator` (deal.II stand-ins),
`Parameters.AllParameters` (include/parameters.h:191) and `Fluid.MPI.InsIM`
(include/mpi_insim.h:42-87). All arithmetic runs in the CUDA library
openifem_b200/lib/libopenifem_b200.so; there is no CPU path in this package.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import IfemError, InsControl, NewtonRecord, SolidRecord, check, dptr, iptr, lptr, lib

__all__ = ["Triangulation", "GridGenerator", "Parameters", "Fluid", "Solid", "MPI", "io", "Partition", "IfemError", "init", "init_distributed",
           "comm_unique_id", "comm_init", "comm_finalize", "peer_selftest", "bench_fp64_peak", "kernel_launches", "set_host_threads"]


def init(device: int = 0):
    check(lib().ifem_init(C.c_int(device)))


def comm_unique_id() -> bytes:
    """Rank 0: create the 128-byte NCCL unique id to broadcast to the other ranks."""
    buf = (C.c_ubyte * 128)()
    check(lib().ifem_comm_unique_id(buf))
    return bytes(buf)


def comm_init(rank: int, size: int, unique_id: bytes):
    """One process per GPU: join the NCCL communicator (replaces MPI_COMM_WORLD of the reference)."""
    buf = (C.c_ubyte * 128).from_buffer_copy(unique_id)
    check(lib().ifem_comm_init(C.c_int(rank), C.c_int(size), buf))


def bench_fp64_peak():
    """measured FP64 FMA throughput of the device in TFLOP/s (register-only FMA chains)"""
    v = C.c_double()
    check(lib().ifem_bench_fp64_peak(C.byref(v)))
    return v.value


def peer_selftest(rounds=200):
    """collective: mismatches seen by this rank in the peer-memory self test (0 = fine, -1 = link inactive)"""
    bad = C.c_int64()
    check(lib().ifem_peer_selftest(C.c_int(rounds), C.byref(bad)))
    return bad.value


def comm_finalize():
    check(lib().ifem_comm_finalize())


def init_distributed(local_rank: int = None):
    """Convenience for torchrun launches: bind the device, then build the library's communicator from
    torch.distributed's rendezvous (the id travels through the process group, nothing else does)."""
    import os

    import torch
    import torch.distributed as dist

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0")) if local_rank is None else local_rank
    torch.cuda.set_device(local_rank)
    init(local_rank)
    # torchrun pins OMP_NUM_THREADS=1; the host-side setup (patterns, partition) is OpenMP code
    set_host_threads(max(1, (os.cpu_count() or 1) // max(1, world)))
    if world > 1:
        if not dist.is_initialized():
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        box = [comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        comm_init(rank, world, box[0])
    return rank, world


class Partition:
    """Host-side domain decomposition of a triangulation for FE_Q(pu)^dim x FE_Q(pp) (no device needed)."""

    def __init__(self, tria, velocity_degree, pressure_degree, rank, size):
        self._h = C.c_void_p()
        check(lib().ifem_partition_create(tria._h, C.c_int(velocity_degree), C.c_int(pressure_degree), C.c_int(rank),
                                          C.c_int(size), C.byref(self._h)))

    def __del__(self):
        if getattr(self, "_h", None) and _lib._lib is not None:
            _lib._lib.ifem_partition_destroy(self._h)
            self._h = None

    def counts(self, which):
        v = [C.c_int() for _ in range(5)]
        check(lib().ifem_partition_counts(self._h, C.c_int(which), *[C.byref(x) for x in v]))
        return dict(n_owned=v[0].value, n_layer1=v[1].value, n_local=v[2].value, n_neighbours=v[3].value, n_local_cells=v[4].value)

    def local_to_global(self, which):
        out = np.empty(self.counts(which)["n_local"], dtype=np.int32)
        check(lib().ifem_partition_local_to_global(self._h, C.c_int(which), iptr(out)))
        return out

    def neighbours(self, which):
        res = []
        for k in range(self.counts(which)["n_neighbours"]):
            v = [C.c_int() for _ in range(4)]
            check(lib().ifem_partition_neighbour(self._h, C.c_int(which), C.c_int(k), *[C.byref(x) for x in v]))
            send = np.empty(v[1].value, dtype=np.int32)
            check(lib().ifem_partition_send_list(self._h, C.c_int(which), C.c_int(k), iptr(send)))
            res.append(dict(rank=v[0].value, send_local=send, recv_offset=v[2].value, recv_count=v[3].value))
        return res


def set_host_threads(n: int):
    check(lib().ifem_set_host_threads(C.c_int(n)))


def kernel_launches() -> int:
    n = C.c_int64(0)
    check(lib().ifem_kernel_launches(C.byref(n)))
    return n.value


def set_spmv_short_variant(key: int):
    check(lib().ifem_set_spmv_short_variant(C.c_int(key)))


def ilu0_apply(A, b):
    """ILU(0) of the scipy CSR matrix A on the device and x = U^-1 L^-1 b; returns (factors in A's pattern, x, (levels L, levels U))"""
    A = A.tocsr()
    A.sort_indices()
    rp = np.ascontiguousarray(A.indptr, dtype=np.int64)
    ci = np.ascontiguousarray(A.indices, dtype=np.int32)
    v = np.ascontiguousarray(A.data, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    f, x = np.empty_like(v), np.empty_like(b)
    nl, nu = C.c_int(), C.c_int()
    check(lib().ifem_ilu0_apply(C.c_int(A.shape[0]), lptr(rp), iptr(ci), dptr(v), dptr(b), dptr(f), dptr(x), C.byref(nl), C.byref(nu)))
    return f, x, (nl.value, nu.value)


class Triangulation:
    def __init__(self, dim: int):
        self.dim = dim
        self._h = C.c_void_p()
        check(lib().ifem_tria_create(C.c_int(dim), C.byref(self._h)))

    def refine_global(self, times: int):
        check(lib().ifem_tria_refine_global(self._h, C.c_int(times)))

    def execute_refinement(self, flags):
        """set_refine_flag() on the cells with a nonzero flag + execute_coarsening_and_refinement()"""
        flags = np.ascontiguousarray(flags, dtype=np.uint8)
        check(lib().ifem_tria_execute_refinement(self._h, flags.ctypes.data_as(C.POINTER(C.c_ubyte)), C.c_int64(flags.size)))

    def execute_coarsening_and_refinement(self, refine_flags, coarsen_flags=None):
        """flagged refinement and coarsening with 2:1 balancing (one level per call); returns the transfer plan
        (ptr, old_vertex, weight): value at new vertex k = sum of weight * old value over ptr[k]:ptr[k+1]"""
        rf = np.ascontiguousarray(refine_flags, dtype=np.uint8)
        cf = None if coarsen_flags is None else np.ascontiguousarray(coarsen_flags, dtype=np.uint8)
        check(lib().ifem_tria_execute_coarsening_and_refinement(self._h, rf.ctypes.data_as(C.POINTER(C.c_ubyte)),
                                                                None if cf is None else cf.ctypes.data_as(C.POINTER(C.c_ubyte)), C.c_int64(rf.size)))
        nv, ne = C.c_int64(), C.c_int64()
        check(lib().ifem_tria_get_transfer_plan(self._h, C.byref(nv), C.byref(ne), None, None, None))
        ptr, old, w = np.empty(nv.value + 1, np.int64), np.empty(ne.value, np.int32), np.empty(ne.value)
        check(lib().ifem_tria_get_transfer_plan(self._h, None, None, lptr(ptr), iptr(old), dptr(w)))
        return ptr, old, w

    def levels(self):
        out = np.empty(self.n_active_cells(), dtype=np.int32)
        check(lib().ifem_tria_get_levels(self._h, iptr(out)))
        return out

    def hanging(self):
        """(vertex [n], n_masters [n], masters [n][4]) of the hanging vertices of the active mesh"""
        n = C.c_int64()
        check(lib().ifem_tria_get_hanging(self._h, C.byref(n), None, None, None))
        v, k, m = np.empty(n.value, np.int32), np.empty(n.value, np.int32), np.empty((n.value, 4), np.int32)
        if n.value:
            check(lib().ifem_tria_get_hanging(self._h, C.byref(n), iptr(v), iptr(k), iptr(m)))
        return v, k, m

    def set_material_ids(self, ids):
        """cell->set_material_id() of every active cell (1-based solid part numbers)"""
        ids = np.ascontiguousarray(ids, dtype=np.int32)
        check(lib().ifem_tria_set_material_ids(self._h, iptr(ids), C.c_int64(ids.size)))

    def _counts(self):
        a, b, c = C.c_int64(), C.c_int64(), C.c_int64()
        check(lib().ifem_tria_counts(self._h, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def n_vertices(self):
        return self._counts()[0]

    def get_mesh(self):
        """(vertices [nv][dim], cells [nc][2^dim], boundary_faces [nbf][3] = (cell, face_no, boundary id))"""
        nv, nc, nbf = self._counts()
        v = np.empty((nv, self.dim))
        c = np.empty((nc, 1 << self.dim), dtype=np.int32)
        b = np.empty((nbf, 3), dtype=np.int32)
        check(lib().ifem_tria_get_mesh(self._h, dptr(v), iptr(c), iptr(b)))
        return v, c, b

    def n_active_cells(self):
        return self._counts()[1]

    def set_mesh(self, vertices, cells, boundary_faces, material_ids=None):
        """import a mesh from plain arrays (the layout get_mesh() returns)"""
        v = np.ascontiguousarray(vertices, dtype=np.float64)
        c = np.ascontiguousarray(cells, dtype=np.int32)
        b = np.ascontiguousarray(boundary_faces, dtype=np.int32).reshape(-1, 3)
        m = None if material_ids is None else np.ascontiguousarray(material_ids, dtype=np.int32)
        check(lib().ifem_tria_set_mesh(self._h, C.c_int64(v.shape[0]), dptr(v), C.c_int64(c.shape[0]), iptr(c), C.c_int64(b.shape[0]),
                                       iptr(b) if b.size else None, None if m is None else iptr(m)))

    def __del__(self):
        if getattr(self, "_h", None) and _lib._lib is not None:
            _lib._lib.ifem_tria_destroy(self._h)
            self._h = None


class GridCreator:
    """Utils::GridCreator<dim> (reference source/utilities.cpp:343-574)"""

    @staticmethod
    def flow_around_cylinder(tria: Triangulation):
        check(lib().ifem_tria_flow_around_cylinder(tria._h))


class GridGenerator:
    @staticmethod
    def subdivided_hyper_rectangle(tria: Triangulation, repetitions, p1, p2, colorize=False):
        reps = (C.c_uint * tria.dim)(*[int(r) for r in repetitions])
        a = (C.c_double * tria.dim)(*[float(x) for x in p1])
        b = (C.c_double * tria.dim)(*[float(x) for x in p2])
        check(lib().ifem_tria_subdivided_hyper_rectangle(tria._h, reps, a, b, C.c_int(1 if colorize else 0)))

    @staticmethod
    def hyper_cube(tria: Triangulation, left=0.0, right=1.0, colorize=False):
        check(lib().ifem_tria_hyper_cube(tria._h, C.c_double(left), C.c_double(right), C.c_int(1 if colorize else 0)))


class Parameters:
    class AllParameters:
        def __init__(self, prm_file: str = None, text: str = None):
            self._h = C.c_void_p()
            if text is not None:
                check(lib().ifem_params_from_text(text.encode(), C.byref(self._h)))
            else:
                check(lib().ifem_params_from_file(prm_file.encode(), C.byref(self._h)))

        def __del__(self):
            if getattr(self, "_h", None) and _lib._lib is not None:
                _lib._lib.ifem_params_destroy(self._h)
                self._h = None


BC_FN = C.CFUNCTYPE(C.c_double, C.POINTER(C.c_double), C.c_uint, C.c_double, C.c_void_p)


class _InsIM:
    """Fluid::MPI::InsIM<dim>(triangulation, parameters)."""

    PRESENT, EVALUATION_POINT, FSI_ACCELERATION, NEWTON_UPDATE, SYSTEM_RHS, DIAG_MU = range(6)

    def __init__(self, tria: Triangulation, params: "Parameters.AllParameters"):
        self.tria, self.params = tria, params  # the solver keeps a reference to the caller's triangulation
        self._h = C.c_void_p()
        check(lib().ifem_insim_create(tria._h, params._h, C.byref(self._h)))

    def __del__(self):
        if getattr(self, "_h", None) and _lib._lib is not None:
            _lib._lib.ifem_insim_destroy(self._h)
            self._h = None

    # -- reference surface --------------------------------------------------
    def run(self):
        check(lib().ifem_insim_run(self._h))

    def add_hard_coded_boundary_condition(self, boundary_id: int, f):
        """FluidSolver::add_hard_coded_boundary_condition(id, f(point, component, time)); call before setup() / run()"""
        dim = self.tria.dim
        cb = BC_FN(lambda p, c, t, _u: float(f([p[i] for i in range(dim)], int(c), float(t))))
        if not hasattr(self, "_keep_bc"):
            self._keep_bc = []
        self._keep_bc.append(cb)  # the library calls it later: keep the trampoline alive
        check(lib().ifem_insim_add_hard_coded_boundary_condition(self._h, C.c_int(boundary_id), cb, None))

    def run_one_step(self, apply_nonzero_constraints: bool, assemble_system: bool = True):
        check(lib().ifem_insim_run_one_step(self._h, C.c_int(1 if apply_nonzero_constraints else 0)))

    def get_current_solution(self) -> np.ndarray:
        out = np.empty(self.n_dofs)
        check(lib().ifem_insim_get_current_solution(self._h, dptr(out)))
        return out

    # -- result files and checkpoints (FluidSolver::output_results / save_checkpoint / load_checkpoint) ----------
    def set_output_directory(self, directory):
        """switches the .vtu / .pvtu / .pvd output and the checkpoints of run() / run_one_step() on (None: off)"""
        check(lib().ifem_insim_set_output_directory(self._h, directory.encode() if directory else None))

    def output_results(self, output_index: int):
        check(lib().ifem_insim_output_results(self._h, C.c_uint(output_index)))

    def save_checkpoint(self, output_index: int):
        check(lib().ifem_insim_save_checkpoint(self._h, C.c_int(output_index)))

    def load_checkpoint(self) -> bool:
        loaded = C.c_int()
        check(lib().ifem_insim_load_checkpoint(self._h, C.byref(loaded)))
        return bool(loaded.value)

    def get_time(self):
        """(time.current(), time.get_timestep())"""
        t, k = C.c_double(), C.c_uint()
        check(lib().ifem_insim_get_time(self._h, C.byref(t), C.byref(k)))
        return t.value, k.value

    def setup(self):
        """setup_dofs(); make_constraints(); initialize_system()"""
        check(lib().ifem_insim_setup(self._h))

    def assemble(self, use_nonzero_constraints: bool):
        check(lib().ifem_insim_assemble(self._h, C.c_int(1 if use_nonzero_constraints else 0)))

    def solve(self, use_nonzero_constraints: bool):
        its, res = C.c_uint(), C.c_double()
        check(lib().ifem_insim_solve(self._h, C.c_int(1 if use_nonzero_constraints else 0), C.byref(its), C.byref(res)))
        return its.value, res.value

    # -- inspection -----------------------------------------------------------
    def sizes(self):
        v = [C.c_int64() for _ in range(5)]
        check(lib().ifem_insim_sizes(self._h, *[C.byref(x) for x in v]))
        return tuple(x.value for x in v)

    @property
    def n_u(self):
        return self.sizes()[0]

    @property
    def n_p(self):
        return self.sizes()[1]

    @property
    def n_dofs(self):
        s = self.sizes()
        return s[0] + s[1]

    def support_points(self):
        pts = np.empty((self.n_dofs, self.tria.dim))
        check(lib().ifem_insim_support_points(self._h, dptr(pts)))
        return pts

    def set_control(self, serial_twin=False, **kw):
        """change some tolerances; the others keep the solver's current values (serial_twin: start from the serial InsIM's)"""
        c = InsControl()
        if serial_twin:
            check(lib().ifem_insim_default_control(C.c_int(1), C.byref(c)))
        else:
            check(lib().ifem_insim_get_control(self._h, C.byref(c)))
        for k, v in kw.items():
            if not hasattr(c, k):
                raise KeyError(k)
            setattr(c, k, v)
        check(lib().ifem_insim_set_control(self._h, C.byref(c)))

    def set_verbose(self, v=True):
        check(lib().ifem_insim_set_verbose(self._h, C.c_int(1 if v else 0)))

    def set_vector(self, which: int, host: np.ndarray):
        host = np.ascontiguousarray(host, dtype=np.float64)
        check(lib().ifem_insim_set_vector(self._h, C.c_int(which), dptr(host)))

    def get_vector(self, which: int) -> np.ndarray:
        n = self.n_u if which == self.DIAG_MU else self.n_dofs
        out = np.empty(n)
        check(lib().ifem_insim_get_vector(self._h, C.c_int(which), dptr(out)))
        return out

    def set_indicator(self, ind: np.ndarray):
        ind = np.ascontiguousarray(ind, dtype=np.int32)
        check(lib().ifem_insim_set_indicator(self._h, iptr(ind)))

    def get_matrix(self, which=0):
        """0 system_matrix, 1 M_p, 2 mass_schur, 3 system matrix of the attached turbulence model -> scipy.sparse.csr_matrix"""
        import scipy.sparse as sp

        n_u, n_p, nnz, nnz_mp, nnz_s = self.sizes()
        n, z = {0: (n_u + n_p, nnz), 1: (n_p, nnz_mp), 2: (n_p, nnz_s), 3: (n_p, nnz_mp)}[which]
        rp, ci, v = np.empty(n + 1, dtype=np.int64), np.empty(z, dtype=np.int32), np.empty(z)
        check(lib().ifem_insim_get_matrix(self._h, C.c_int(which), lptr(rp), iptr(ci), dptr(v)))
        return sp.csr_matrix((v, ci, rp), shape=(n, n))

    def vmult(self, x: np.ndarray) -> np.ndarray:
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.empty_like(x)
        check(lib().ifem_insim_vmult(self._h, dptr(x), dptr(y)))
        return y

    def history(self, max_records=4096):
        buf = (NewtonRecord * max_records)()
        n = C.c_int()
        check(lib().ifem_insim_history(self._h, C.c_int(max_records), buf, C.byref(n)))
        k = min(n.value, max_records)
        return [{f: getattr(buf[i], f) for f, _ in NewtonRecord._fields_} for i in range(k)]

    def partition(self, which):
        """(n_owned_nodes, n_local_nodes) of velocity (0) / pressure (1) nodes on this rank"""
        a, b = C.c_int(), C.c_int()
        check(lib().ifem_insim_partition(self._h, C.c_int(which), C.byref(a), C.byref(b)))
        return a.value, b.value

    def local_to_global(self, which):
        out = np.empty(self.partition(which)[1], dtype=np.int32)
        check(lib().ifem_insim_local_to_global(self._h, C.c_int(which), iptr(out)))
        return out

    def owned_global_dofs(self, n_unodes_global):
        """(local dof indices, global dof indices) of the dofs this rank owns, global numbering [u | p]."""
        dim = self.tria.dim
        ou, lu = self.partition(0)
        op, _ = self.partition(1)
        gu, gp = self.local_to_global(0), self.local_to_global(1)
        loc_u = (np.arange(ou)[:, None] * dim + np.arange(dim)[None, :]).ravel()
        glo_u = (gu[:ou, None].astype(np.int64) * dim + np.arange(dim)[None, :]).ravel()
        loc_p = dim * lu + np.arange(op)
        glo_p = dim * n_unodes_global + gp[:op].astype(np.int64)
        return np.concatenate([loc_u, loc_p]), np.concatenate([glo_u, glo_p])

    def timer_ms(self, section: str) -> float:
        ms = C.c_double()
        check(lib().ifem_insim_timer_ms(self._h, section.encode(), C.byref(ms)))
        return ms.value

    def time(self):
        ts, cur = C.c_uint(), C.c_double()
        check(lib().ifem_insim_time(self._h, C.byref(ts), C.byref(cur)))
        return ts.value, cur.value

    def update_stress(self):
        """FluidSolver::update_stress: nodal viscous stress from present_solution"""
        check(lib().ifem_scnsim_update_stress(self._h))

    def get_stress(self):
        dim = self.tria.dim
        out = np.empty((dim * dim, self.partition(0)[1]))
        check(lib().ifem_scnsim_get_field(self._h, C.c_int(0), dptr(out)))
        return out

    # -- measurement hooks ---------------------------------------------------
    def bench_vmult(self, reps):
        ms, b = C.c_double(), C.c_double()
        check(lib().ifem_insim_bench_vmult(self._h, C.c_int(reps), C.byref(ms), C.byref(b)))
        return ms.value, b.value

    def bench_spmv_uu(self, reps):
        ms, b = C.c_double(), C.c_double()
        check(lib().ifem_insim_bench_spmv_uu(self._h, C.c_int(reps), C.byref(ms), C.byref(b)))
        return ms.value, b.value

    def bench_spmv_uu_fp32(self, reps):
        ms, b = C.c_double(), C.c_double()
        check(lib().ifem_insim_bench_spmv_uu_fp32(self._h, C.c_int(reps), C.byref(ms), C.byref(b)))
        return ms.value, b.value

    def solve_mass_schur(self, b, mode=0, rel_tol=1e-3, max_it=100000):
        """'CG for Sm' on its own: (x, iterations, residual) with S_m x = b in cg_sm_fp32 mode 0 / 1 / 2"""
        b = np.ascontiguousarray(b, dtype=np.float64)
        x = np.empty_like(b)
        its, res = C.c_int(), C.c_double()
        check(lib().ifem_insim_solve_mass_schur(self._h, C.c_int(mode), dptr(b), C.c_double(rel_tol), C.c_int(max_it), dptr(x),
                                                C.byref(its), C.byref(res)))
        return x, its.value, res.value

    def set_inner_variant(self, variant):
        check(lib().ifem_insim_set_inner_variant(self._h, C.c_int(variant)))

    def bench_spmv_uu_sell(self, reps, variant=0, check_error=True, precision=0):
        """(ms, algorithmic bytes, padding ratio, max rel. error vs the fp64 product) of the SELL-32 product kernel of the
        fp32 inner solver; precision = storage of the matrix values (32, 16; 0 = as built)"""
        ms, b, pad, err = C.c_double(), C.c_double(), C.c_double(), C.c_double(-1.0)
        check(lib().ifem_insim_bench_spmv_uu_sell(self._h, C.c_int(precision), C.c_int(variant), C.c_int(reps), C.byref(ms), C.byref(b), C.byref(pad),
                                                  C.byref(err) if check_error else None))
        return ms.value, b.value, pad.value, err.value

    def bench_steps(self, n_steps, first_applies_nonzero_constraints=False):
        ms = C.c_double()
        check(lib().ifem_insim_bench_steps(self._h, C.c_int(n_steps), C.c_int(1 if first_applies_nonzero_constraints else 0), C.byref(ms)))
        return ms.value

    def bench_assemble(self, reps):
        ms = C.c_double()
        check(lib().ifem_insim_bench_assemble(self._h, C.c_int(reps), C.byref(ms)))
        return ms.value


FIELD_FN = C.CFUNCTYPE(C.c_double, C.POINTER(C.c_double), C.c_uint, C.c_void_p)


class _InsIMEX(_InsIM):
    """Fluid::MPI::InsIMEX<dim>(triangulation, parameters): implicit-explicit twin of InsIM (source/mpi_insimex.cpp); shares
    the InsIM handle API, NEWTON_UPDATE holds solution_time_increment."""

    def __init__(self, tria, params):
        self.tria, self.params = tria, params
        self._h = C.c_void_p()
        check(lib().ifem_insimex_create(tria._h, params._h, C.byref(self._h)))

    def assemble(self, use_nonzero_constraints: bool, assemble_system: bool = True):
        check(lib().ifem_insimex_assemble(self._h, C.c_int(1 if use_nonzero_constraints else 0), C.c_int(1 if assemble_system else 0)))

    def solve(self, use_nonzero_constraints: bool, assemble_system: bool = True):
        its, res = C.c_uint(), C.c_double()
        check(lib().ifem_insimex_solve(self._h, C.c_int(1 if use_nonzero_constraints else 0), C.c_int(1 if assemble_system else 0),
                                       C.byref(its), C.byref(res)))
        return its.value, res.value

    def run_one_step(self, apply_nonzero_constraints: bool, assemble_system: bool = True):
        check(lib().ifem_insimex_run_one_step(self._h, C.c_int(1 if apply_nonzero_constraints else 0),
                                              C.c_int(1 if assemble_system else 0)))


class _SCnsIM(_InsIM):
    """Fluid::MPI::SCnsIM<dim>(triangulation, parameters); shares the InsIM handle API."""

    def __init__(self, tria, params):
        self.tria, self.params = tria, params
        self._h = C.c_void_p()
        self._keep = []
        check(lib().ifem_scnsim_create(tria._h, params._h, C.byref(self._h)))

    def _wrap(self, f):
        dim = self.tria.dim
        cb = FIELD_FN(lambda p, c, _u: float(f([p[i] for i in range(dim)], int(c))))
        self._keep.append(cb)
        return cb

    def set_body_force(self, f):
        check(lib().ifem_scnsim_set_body_force(self._h, self._wrap(f), None))

    def set_sigma_pml_field(self, f):
        check(lib().ifem_scnsim_set_sigma_pml_field(self._h, self._wrap(f), None))

    def set_initial_condition(self, f):
        check(lib().ifem_scnsim_set_initial_condition(self._h, self._wrap(f), None))

    def _field_shape(self, which):
        dim = self.tria.dim
        return (dim * dim if which == 0 else dim * (dim + 1) // 2, self.partition(0)[1])

    def get_field(self, which):
        out = np.empty(self._field_shape(which))
        check(lib().ifem_scnsim_get_field(self._h, C.c_int(which), dptr(out)))
        return out

    def set_field(self, which, host):
        host = np.ascontiguousarray(host, dtype=np.float64).reshape(self._field_shape(which))
        check(lib().ifem_scnsim_set_field(self._h, C.c_int(which), dptr(host)))

    def attach_turbulence_model(self, model_name: str):
        """FluidSolver::attach_turbulence_model; returns the model (a view of this solver's handle)"""
        check(lib().ifem_insim_attach_turbulence_model(self._h, model_name.encode()))
        self.turbulence_model = _TurbulenceModel(self)
        return self.turbulence_model


class _TurbulenceModel:
    """Fluid::MPI::SpalartAllmaras<dim> attached to an SCnsIM solver (lives inside the solver's handle)"""

    PRESENT, EVALUATION_POINT, EDDY_VISCOSITY, SYSTEM_RHS, NEWTON_UPDATE, WALL_DISTANCE = range(6)

    def __init__(self, fluid):
        self.fluid = fluid

    @property
    def n_dofs(self):
        return self.fluid.partition(1)[1]

    def get_vector(self, which):
        out = np.empty(self.n_dofs)
        check(lib().ifem_turbulence_get_vector(self.fluid._h, C.c_int(which), dptr(out)))
        return out

    def set_vector(self, which, host):
        host = np.ascontiguousarray(host, dtype=np.float64)
        assert host.size == self.n_dofs
        check(lib().ifem_turbulence_set_vector(self.fluid._h, C.c_int(which), dptr(host)))

    def get_eddy_viscosity(self):
        return self.get_vector(self.EDDY_VISCOSITY)

    def assemble(self, use_nonzero_constraints):
        check(lib().ifem_turbulence_assemble(self.fluid._h, C.c_int(1 if use_nonzero_constraints else 0)))

    def get_matrix(self):
        return self.fluid.get_matrix(3)

    def run_one_step(self, apply_nonzero_constraints):
        check(lib().ifem_turbulence_run_one_step(self.fluid._h, C.c_int(1 if apply_nonzero_constraints else 0)))

    def update_boundary_condition(self, first_step):
        check(lib().ifem_turbulence_update_boundary_condition(self.fluid._h, C.c_int(1 if first_step else 0)))

    def get_shear_velocity(self, vel, init_guess):
        out = C.c_double()
        check(lib().ifem_turbulence_get_shear_velocity(self.fluid._h, C.c_double(vel), C.c_double(init_guess), C.byref(out)))
        return out.value

    def history(self, max_records=1024):
        res, its, n = np.empty(max_records), np.empty(max_records, dtype=np.int32), C.c_int()
        check(lib().ifem_turbulence_history(self.fluid._h, C.c_int(max_records), dptr(res), iptr(its), C.byref(n)))
        k = min(n.value, max_records)
        return list(zip(res[:k].tolist(), its[:k].tolist()))


class _HyperElasticity:
    """Solid::MPI::HyperElasticity<dim>(triangulation, parameters) - NeoHookean, Newmark-beta."""

    CUR_U, CUR_V, CUR_A, PREV_U, PREV_V, PREV_A, SYSTEM_RHS = range(7)

    def __init__(self, tria: Triangulation, params: "Parameters.AllParameters"):
        self.tria, self.params = tria, params
        self._h = C.c_void_p()
        check(lib().ifem_hyper_create(tria._h, params._h, C.byref(self._h)))

    def __del__(self):
        if getattr(self, "_h", None) and _lib._lib is not None:
            _lib._lib.ifem_hyper_destroy(self._h)
            self._h = None

    def run(self):
        check(lib().ifem_hyper_run(self._h))

    def run_one_step(self, first_step: bool):
        check(lib().ifem_hyper_run_one_step(self._h, C.c_int(1 if first_step else 0)))

    # -- result files and checkpoints (SharedSolidSolver::output_results / save_checkpoint / load_checkpoint) ----
    def set_output_directory(self, directory):
        check(lib().ifem_hyper_set_output_directory(self._h, directory.encode() if directory else None))

    def output_results(self, output_index: int):
        check(lib().ifem_hyper_output_results(self._h, C.c_uint(output_index)))

    def save_checkpoint(self, output_index: int):
        check(lib().ifem_hyper_save_checkpoint(self._h, C.c_int(output_index)))

    def load_checkpoint(self) -> bool:
        loaded = C.c_int()
        check(lib().ifem_hyper_load_checkpoint(self._h, C.byref(loaded)))
        return bool(loaded.value)

    def get_time(self):
        t, k = C.c_double(), C.c_uint()
        check(lib().ifem_hyper_get_time(self._h, C.byref(t), C.byref(k)))
        return t.value, k.value

    def setup(self):
        check(lib().ifem_hyper_setup(self._h))

    def set_verbose(self, v=True):
        check(lib().ifem_hyper_set_verbose(self._h, C.c_int(1 if v else 0)))

    def sizes(self):
        a, b, c, d = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int()
        check(lib().ifem_hyper_sizes(self._h, C.byref(a), C.byref(b), C.byref(c), C.byref(d)))
        return a.value, b.value, c.value, d.value

    @property
    def n_dofs(self):
        return self.sizes()[0]

    def get_current_solution(self):
        out = np.empty(self.n_dofs)
        check(lib().ifem_hyper_get_current_solution(self._h, dptr(out)))
        return out

    def set_vector(self, which, host):
        host = np.ascontiguousarray(host, dtype=np.float64)
        check(lib().ifem_hyper_set_vector(self._h, C.c_int(which), dptr(host)))

    def get_vector(self, which):
        out = np.empty(self.n_dofs)
        check(lib().ifem_hyper_get_vector(self._h, C.c_int(which), dptr(out)))
        return out

    def update_qph(self):
        check(lib().ifem_hyper_update_qph(self._h))

    def assemble_system(self, initial_step: bool):
        check(lib().ifem_hyper_assemble_system(self._h, C.c_int(1 if initial_step else 0)))

    def get_matrix(self, which=0):
        import scipy.sparse as sp

        n, nnz, _, _ = self.sizes()
        rp, ci, v = np.empty(n + 1, dtype=np.int64), np.empty(nnz, dtype=np.int32), np.empty(nnz)
        check(lib().ifem_hyper_get_matrix(self._h, C.c_int(which), lptr(rp), iptr(ci), dptr(v)))
        return sp.csr_matrix((v, ci, rp), shape=(n, n))

    def get_qph(self):
        _, _, nqp, nsym = self.sizes()
        dim = self.tria.dim
        Finv, tau = np.empty((nqp, dim, dim)), np.empty((nqp, dim, dim))
        Jc, det = np.empty((nqp, nsym, nsym)), np.empty(nqp)
        check(lib().ifem_hyper_get_qph(self._h, dptr(Finv), dptr(tau), dptr(Jc), dptr(det)))
        return Finv, tau, Jc, det

    def update_strain_and_stress(self):
        check(lib().ifem_hyper_update_strain_and_stress(self._h))

    def get_fsi_inputs(self):
        dim = self.tria.dim
        n = self.n_dofs
        rows, vel, pres = np.empty((dim, n)), np.empty(n), np.empty(n // dim)
        check(lib().ifem_hyper_get_fsi_inputs(self._h, dptr(rows), dptr(vel), dptr(pres)))
        return rows, vel, pres

    def _tensor_shape(self):
        dim = self.tria.dim
        return (dim * dim, self.n_dofs // dim)

    def get_nodal_tensor(self, which):
        out = np.empty(self._tensor_shape())
        check(lib().ifem_hyper_get_nodal_tensor(self._h, C.c_int(which), dptr(out)))
        return out

    def set_nodal_tensor(self, which, host):
        host = np.ascontiguousarray(host, dtype=np.float64).reshape(self._tensor_shape())
        check(lib().ifem_hyper_set_nodal_tensor(self._h, C.c_int(which), dptr(host)))

    def history(self, max_records=4096):
        buf = (SolidRecord * max_records)()
        n = C.c_int()
        check(lib().ifem_hyper_history(self._h, C.c_int(max_records), buf, C.byref(n)))
        return [{f: getattr(buf[i], f) for f, _ in SolidRecord._fields_} for i in range(min(n.value, max_records))]


class _SharedHyperElasticity(_HyperElasticity):
    """Solid::MPI::SharedHyperElasticity<dim>(triangulation, parameters): the replicated twin MPI::FSI takes (extra Newton stop on a
    vanishing update, nodal strain / stress after every step) whatever `Simulation type` says."""

    def __init__(self, tria: Triangulation, params: "Parameters.AllParameters"):
        self.tria, self.params = tria, params
        self._h = C.c_void_p()
        check(lib().ifem_hyper_create_twin(tria._h, params._h, C.c_int(1), C.byref(self._h)))


class _LinearElasticity(_HyperElasticity):
    """Solid::MPI::LinearElasticity<dim>(triangulation, parameters): small-strain elasticity, Newmark-beta in
    acceleration form. Same handle type as the other solid solvers (update_qph / get_qph do not apply)."""

    SHARED = False
    SYSTEM, MASS, STIFFNESS, DAMPING = range(4)

    def __init__(self, tria: Triangulation, params: "Parameters.AllParameters"):
        self.tria, self.params = tria, params
        self._h = C.c_void_p()
        check(lib().ifem_linear_elasticity_create(tria._h, params._h, C.c_int(1 if self.SHARED else 0), C.byref(self._h)))


class _SharedLinearElasticity(_LinearElasticity):
    """Solid::MPI::SharedLinearElasticity<dim>(triangulation, parameters): the replicated twin MPI::FSI drives."""

    SHARED = True


POINT_FN = C.CFUNCTYPE(C.c_double, C.POINTER(C.c_double), C.c_void_p)


class _FSI:
    """MPI::FSI<dim>(fluid_solver, solid_solver, parameters, use_dirichlet_bc) - coupling kernels."""

    def __init__(self, fluid, solid, params, use_dirichlet_bc=False):
        self.fluid, self.solid, self.params = fluid, solid, params
        self._h = C.c_void_p()
        check(lib().ifem_fsi_create(fluid._h, solid._h, params._h, C.c_int(1 if use_dirichlet_bc else 0), C.byref(self._h)))

    def __del__(self):
        if getattr(self, "_h", None) and _lib._lib is not None:
            _lib._lib.ifem_fsi_destroy(self._h)
            self._h = None

    def update_solid_box(self):
        box = np.empty(2 * self.fluid.tria.dim)
        check(lib().ifem_fsi_update_solid_box(self._h, dptr(box)))
        return box

    def update_indicator(self):
        check(lib().ifem_fsi_update_indicator(self._h))
        ind = np.empty(self.fluid.tria.n_active_cells() if self.fluid.partition(0)[0] == self.fluid.partition(0)[1] else self._n_local_cells(), dtype=np.int32)
        check(lib().ifem_fsi_get_indicator(self._h, iptr(ind)))
        return ind

    def get_indicator(self):
        """CellProperty::indicator of the local fluid cells as the last update_indicator() left it"""
        ind = np.empty(self.fluid.tria.n_active_cells() if self.fluid.partition(0)[0] == self.fluid.partition(0)[1] else self._n_local_cells(), dtype=np.int32)
        check(lib().ifem_fsi_get_indicator(self._h, iptr(ind)))
        return ind

    def _n_local_cells(self):
        raise NotImplementedError("indicator download on multi-rank runs: use the C ABI with the local cell count")

    def find_solid_bc(self):
        check(lib().ifem_fsi_find_solid_bc(self._h))
        return self.solid.get_fsi_inputs()

    def run_one_step(self, first_step: bool):
        check(lib().ifem_fsi_run_one_step(self._h, C.c_int(1 if first_step else 0)))

    def prepare_fluid_step(self, first_step: bool):
        """run_one_step up to and including find_fluid_bc, without the fluid time step"""
        check(lib().ifem_fsi_prepare_fluid_step(self._h, C.c_int(1 if first_step else 0)))

    def run(self):
        check(lib().ifem_fsi_run(self._h))

    def set_penetration_criterion(self, criterion, direction):
        """FSI::set_penetration_criterion(criterion(point) -> double, direction): enables apply_contact_model"""
        dim = self.fluid.tria.dim
        cb = POINT_FN(lambda p, _u: float(criterion([p[i] for i in range(dim)])))
        self._keep_criterion = cb  # the library calls it during run(): keep the trampoline alive
        d = np.ascontiguousarray(direction, dtype=np.float64)
        check(lib().ifem_fsi_set_penetration_criterion(self._h, cb, None, dptr(d)))

    def contact_iterations(self):
        n = C.c_int()
        check(lib().ifem_fsi_contact_iterations(self._h, C.byref(n)))
        return n.value

    def find_fluid_bc(self):
        check(lib().ifem_fsi_find_fluid_bc(self._h))
        return self.fluid.get_vector(self.fluid.FSI_ACCELERATION)

    def inner_constraints(self):
        n = self.fluid.n_dofs
        flags, inhom = np.empty(n, dtype=np.uint8), np.empty(n)
        check(lib().ifem_fsi_get_inner_constraints(self._h, flags.ctypes.data_as(C.POINTER(C.c_ubyte)), dptr(inhom)))
        return flags, inhom

    def point_in_solid(self, pts):
        pts = np.ascontiguousarray(pts, dtype=np.float64)
        out = np.empty(pts.shape[0], dtype=np.int32)
        check(lib().ifem_fsi_point_in_solid(self._h, C.c_int(pts.shape[0]), dptr(pts), iptr(out)))
        return out.astype(bool)

    def interpolate(self, which, pts):
        pts = np.ascontiguousarray(pts, dtype=np.float64)
        vals = np.empty_like(pts)
        found = np.empty(pts.shape[0], dtype=np.int32)
        check(lib().ifem_fsi_interpolate(self._h, C.c_int(which), C.c_int(pts.shape[0]), dptr(pts), dptr(vals), iptr(found)))
        return vals, found

    def timer_ms(self, section):
        ms = C.c_double()
        check(lib().ifem_fsi_timer_ms(self._h, section.encode(), C.byref(ms)))
        return ms.value

    def refine_mesh(self, min_grid_level, max_grid_level):
        check(lib().ifem_fsi_refine_mesh(self._h, C.c_uint(min_grid_level), C.c_uint(max_grid_level)))

    def bench_steps(self, n_steps, first_step=False):
        """n_steps coupled passes between CUDA events on the library's stream; total milliseconds"""
        ms = C.c_double()
        check(lib().ifem_fsi_bench_steps(self._h, C.c_int(n_steps), C.c_int(1 if first_step else 0), C.byref(ms)))
        return ms.value


class MPI:
    FSI = _FSI


class _SUPGInsIM(_SCnsIM):
    """Fluid::MPI::SUPGInsIM<dim>(triangulation, parameters): stabilised incompressible solver (source/mpi_insim_supg.cpp)."""

    def __init__(self, tria, params):
        self.tria, self.params = tria, params
        self._h = C.c_void_p()
        self._keep = []
        check(lib().ifem_supg_insim_create(tria._h, params._h, C.byref(self._h)))


class io:
    """Host-side pieces of the on-disk formats (openifem_b200/csrc/output.h); no device needed."""

    @staticmethod
    def write_vtu(path, dim, points, cells, point_fields=(), cell_fields=()):
        """points [n][dim]; cells [m][2^dim] lexicographic; point_fields [(name, array [n] or [n][k])]; cell_fields [(name, [m])]"""
        points = np.ascontiguousarray(points, dtype=np.float64)
        cells = np.ascontiguousarray(cells, dtype=np.int32)
        pf = [(n, np.ascontiguousarray(a, dtype=np.float64)) for n, a in point_fields]
        cf = [(n, np.ascontiguousarray(a, dtype=np.float64)) for n, a in cell_fields]

        def pack(fields):
            names = (C.c_char_p * max(len(fields), 1))(*[n.encode() for n, _ in fields])
            data = (C.POINTER(C.c_double) * max(len(fields), 1))(*[dptr(a) for _, a in fields])
            return names, data

        pn, pd = pack(pf)
        cn, cd = pack(cf)
        nc = (C.c_int * max(len(pf), 1))(*[1 if a.ndim == 1 else a.shape[1] for _, a in pf])
        check(lib().ifem_write_vtu(path.encode(), C.c_int(dim), C.c_int64(points.shape[0]), dptr(points), C.c_int64(cells.shape[0]),
                                   iptr(cells), C.c_int(len(pf)), pn, nc, pd, C.c_int(len(cf)), cn, cd))

    @staticmethod
    def write_pvd(path, prefix, times, timesteps):
        t = np.ascontiguousarray(times, dtype=np.float64)
        k = np.ascontiguousarray(timesteps, dtype=np.uint32)
        check(lib().ifem_write_pvd(path.encode(), prefix.encode(), C.c_int(t.size), dptr(t), k.ctypes.data_as(C.POINTER(C.c_uint))))

    @staticmethod
    def block_write(path, values):
        v = np.ascontiguousarray(values, dtype=np.float64)
        check(lib().ifem_block_write(path.encode(), C.c_int64(v.size), dptr(v)))

    @staticmethod
    def block_read(path, capacity):
        out, n = np.empty(capacity), C.c_int64()
        check(lib().ifem_block_read(path.encode(), C.c_int64(capacity), dptr(out), C.byref(n)))
        return out[: n.value]

    @staticmethod
    def fluid_write_results_host(tria, pu, pp, present, fsi_acceleration=None, indicator=None, stress=None, directory=".", output_index=0):
        present = np.ascontiguousarray(present, dtype=np.float64)
        acc = None if fsi_acceleration is None else np.ascontiguousarray(fsi_acceleration, dtype=np.float64)
        ind = None if indicator is None else np.ascontiguousarray(indicator, dtype=np.int32)
        st = None if stress is None else np.ascontiguousarray(stress, dtype=np.float64)
        check(lib().ifem_fluid_write_results_host(tria._h, C.c_int(pu), C.c_int(pp), dptr(present), None if acc is None else dptr(acc),
                                                  None if ind is None else iptr(ind), None if st is None else dptr(st),
                                                  directory.encode(), C.c_uint(output_index)))


class Fluid:
    class MPI:
        InsIM = _InsIM
        InsIMEX = _InsIMEX
        SCnsIM = _SCnsIM
        SUPGInsIM = _SUPGInsIM


class Solid:
    class MPI:
        HyperElasticity = _HyperElasticity
        SharedHyperElasticity = _SharedHyperElasticity
        LinearElasticity = _LinearElasticity
        SharedLinearElasticity = _SharedLinearElasticity
