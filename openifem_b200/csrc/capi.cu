// extern "C" layer: see include/openifem_b200.h for the contract.
#include "../../include/openifem_b200.h"

#include <omp.h>

#include <cstring>
#include <memory>
#include <string>

#include "comm.h"
#include "peer.h"
#include "fsi.h"
#include "ilu0.h"
#include "insim.h"
#include "insimex.h"
#include "output.h"
#include "partition.h"
#include "scnsim.h"
#include "solid.h"

using namespace ifem;

struct ifem_tria
{
  Triangulation t;
  Triangulation::TransferPlan last_plan; // of the last ifem_tria_execute_coarsening_and_refinement
};
struct ifem_params
{
  std::unique_ptr<Parameters::AllParameters> p;
};
struct ifem_insim
{
  std::unique_ptr<InsIM> s;
};
struct ifem_hyper
{
  std::unique_ptr<SolidSolver> s; // HyperElasticity or LinearElasticity
};
struct ifem_fsi
{
  std::unique_ptr<FsiCoupling> f;
};
struct ifem_partition
{
  Partition p;
};

namespace
{
  thread_local std::string g_error;
  bool g_initialised = false;

  template <typename F>
  int guard(F &&f)
  {
    try
      {
        f();
        return IFEM_OK;
      }
    catch (const std::exception &e)
      {
        g_error = e.what();
        return g_error.find("no CUDA device") != std::string::npos ? IFEM_ERR_NO_DEVICE : IFEM_ERR;
      }
    catch (...)
      {
        g_error = "unknown error";
        return IFEM_ERR;
      }
  }

  void require_device()
  {
    if (g_initialised) return;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0)
      throw std::runtime_error("openifem_b200: no CUDA device available - this library has no CPU fallback");
    g_initialised = true;
  }

  // scalar CSR of a BCSR matrix appended at a row/column offset
  struct HostCsr
  {
    std::vector<int64_t> rp;
    std::vector<int> ci;
    std::vector<double> v;
  };
} // namespace

template <typename F>
static double time_reps(Context &ctx, int reps, F &&f)
{
  cudaEvent_t e0, e1;
  IFEM_CUDA(cudaEventCreate(&e0));
  IFEM_CUDA(cudaEventCreate(&e1));
  IFEM_CUDA(cudaEventRecord(e0, ctx.stream));
  for (int i = 0; i < reps; ++i) f();
  IFEM_CUDA(cudaEventRecord(e1, ctx.stream));
  IFEM_CUDA(cudaEventSynchronize(e1));
  float ms = 0;
  IFEM_CUDA(cudaEventElapsedTime(&ms, e0, e1));
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  return (double)ms / reps;
}

static DevBuf<double> *pick_vector(InsIM &m, int which, int64_t &n)
{
  n = m.fs.n_dofs;
  switch (which)
    {
    case 0: return &m.present_solution;
    case 1: return &m.evaluation_point;
    case 2: return &m.fsi_acceleration;
    case 3: return &m.newton_update;
    case 4: return &m.fs.rhs;
    case 5: n = m.fs.n_u; return &m.fs.diag_Mu;
    default: throw std::runtime_error("unknown vector id");
    }
}
static DevBuf<double> *pick_solid_vector(SolidSolver &m, int which)
{
  switch (which)
    {
    case 0: return &m.current_displacement;
    case 1: return &m.current_velocity;
    case 2: return &m.current_acceleration;
    case 3: return &m.previous_displacement;
    case 4: return &m.previous_velocity;
    case 5: return &m.previous_acceleration;
    case 6: return &m.ss.rhs;
    default: throw std::runtime_error("unknown vector id");
    }
}

static HyperElasticity &as_hyper(ifem_hyper *s)
{
  auto *p = dynamic_cast<HyperElasticity *>(s->s.get());
  if (!p) throw std::runtime_error("this solid solver is not a HyperElasticity");
  return *p;
}

static SCnsIM &as_scns(ifem_insim *s)
{
  auto *p = dynamic_cast<SCnsIM *>(s->s.get());
  if (!p) throw std::runtime_error("handle is not a SCnsIM solver");
  return *p;
}

static const NodePartition &pick_np(const Partition &p, int which) { return which == 0 ? p.u : p.p; }

extern "C" {

const char *ifem_last_error(void) { return g_error.c_str(); }
int ifem_version(void) { return 100; }

int ifem_init(int device)
{
  return guard([&] {
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0)
      throw std::runtime_error("openifem_b200: no CUDA device available - this library has no CPU fallback");
    IFEM_CUDA(cudaSetDevice(device));
    g_initialised = true;
    (void)default_context();
  });
}

int ifem_set_host_threads(int n)
{
  return guard([&] {
    if (n > 0) omp_set_num_threads(n);
  });
}

int ifem_kernel_launches(int64_t *count)
{
  return guard([&] {
    require_device();
    *count = default_context().kernel_launches;
  });
}

int ifem_peer_selftest(int rounds, int64_t *mismatches)
{
  return guard([&] {
    require_device();
    *mismatches = peer_selftest(default_context(), rounds);
  });
}

int ifem_comm_unique_id(unsigned char id[128])
{
  return guard([&] { comm_get_unique_id(id); });
}
int ifem_comm_init(int rank, int size, const unsigned char id[128])
{
  return guard([&] {
    require_device();
    Context &ctx = default_context();
    peer_link_reset();
    if (ctx.comm) comm_destroy(ctx.comm);
    ctx.comm = comm_create(rank, size, id);
  });
}
int ifem_comm_finalize(void)
{
  return guard([&] {
    if (!g_initialised) return;
    Context &ctx = default_context();
    peer_link_reset();
    if (ctx.comm) comm_destroy(ctx.comm);
    ctx.comm = nullptr;
  });
}

int ifem_partition_create(const ifem_tria *t, int pu, int pp, int rank, int size, ifem_partition **out)
{
  return guard([&] {
    const NodeTable un = build_node_table(t->t, pu), pn = build_node_table(t->t, pp);
    auto *h = new ifem_partition;
    h->p = build_partition(t->t, un, pn, rank, size);
    *out = h;
  });
}
int ifem_partition_destroy(ifem_partition *p)
{
  delete p;
  return IFEM_OK;
}
int ifem_partition_counts(const ifem_partition *p, int which, int *n_owned, int *n_layer1, int *n_local, int *n_nb, int *n_cells)
{
  return guard([&] {
    const NodePartition &np = pick_np(p->p, which);
    if (n_owned) *n_owned = np.n_owned;
    if (n_layer1) *n_layer1 = np.n_layer1;
    if (n_local) *n_local = np.n_local;
    if (n_nb) *n_nb = (int)np.neighbours.size();
    if (n_cells) *n_cells = (int)p->p.local_cells.size();
  });
}
int ifem_partition_local_to_global(const ifem_partition *p, int which, int *ids)
{
  return guard([&] {
    const NodePartition &np = pick_np(p->p, which);
    std::copy(np.local_to_global.begin(), np.local_to_global.end(), ids);
  });
}
int ifem_partition_neighbour(const ifem_partition *p, int which, int k, int *rank, int *n_send, int *recv_offset, int *recv_count)
{
  return guard([&] {
    const NodePartition &np = pick_np(p->p, which);
    *rank = np.neighbours.at(k);
    *n_send = (int)np.send_local.at(k).size();
    *recv_offset = np.recv_offset.at(k);
    *recv_count = np.recv_count.at(k);
  });
}
int ifem_partition_send_list(const ifem_partition *p, int which, int k, int *ids)
{
  return guard([&] {
    const auto &l = pick_np(p->p, which).send_local.at(k);
    std::copy(l.begin(), l.end(), ids);
  });
}

int ifem_tria_create(int dim, ifem_tria **out)
{
  return guard([&] {
    if (dim != 2 && dim != 3) throw std::runtime_error("Triangulation: dim must be 2 or 3");
    *out = new ifem_tria;
    (*out)->t.dim = dim;
  });
}
int ifem_tria_destroy(ifem_tria *t)
{
  delete t;
  return IFEM_OK;
}
int ifem_tria_subdivided_hyper_rectangle(ifem_tria *t, const unsigned int *reps, const double *p1, const double *p2, int colorize)
{
  return guard([&] {
    const int dim = t->t.dim;
    GridGenerator::subdivided_hyper_rectangle(t->t, std::vector<unsigned int>(reps, reps + dim), p1, p2, colorize != 0);
  });
}
int ifem_tria_hyper_cube(ifem_tria *t, double left, double right, int colorize)
{
  return guard([&] { GridGenerator::hyper_cube(t->t, t->t.dim, left, right, colorize != 0); });
}
int ifem_tria_refine_global(ifem_tria *t, int times)
{
  return guard([&] { t->t.refine_global(times); });
}
int ifem_tria_shift(ifem_tria *t, const double *offset)
{
  return guard([&] {
    const int dim = t->t.dim;
    for (size_t i = 0; i < t->t.vertices.size(); ++i) t->t.vertices[i] += offset[i % dim];
  });
}
int ifem_tria_execute_refinement(ifem_tria *t, const unsigned char *flags, int64_t n)
{
  return guard([&] {
    if (n != (int64_t)t->t.n_cells()) throw std::runtime_error("ifem_tria_execute_refinement: one flag per active cell expected");
    t->t.execute_refinement(std::vector<unsigned char>(flags, flags + n));
  });
}
int ifem_tria_execute_coarsening_and_refinement(ifem_tria *t, const unsigned char *refine_flags, const unsigned char *coarsen_flags, int64_t n)
{
  return guard([&] {
    if (n != (int64_t)t->t.n_cells()) throw std::runtime_error("ifem_tria_execute_coarsening_and_refinement: one flag per active cell expected");
    std::vector<unsigned char> rf(refine_flags, refine_flags + n), cf;
    if (coarsen_flags) cf.assign(coarsen_flags, coarsen_flags + n);
    t->t.execute_coarsening_and_refinement(rf, cf, &t->last_plan);
  });
}
int ifem_tria_get_transfer_plan(const ifem_tria *t, int64_t *n_new_vertices, int64_t *n_entries, int64_t *ptr, int *old_vertex, double *weight)
{
  return guard([&] {
    const auto &p = t->last_plan;
    if (n_new_vertices) *n_new_vertices = p.ptr.empty() ? 0 : (int64_t)p.ptr.size() - 1;
    if (n_entries) *n_entries = (int64_t)p.old_vertex.size();
    if (ptr) std::copy(p.ptr.begin(), p.ptr.end(), ptr);
    if (old_vertex) std::copy(p.old_vertex.begin(), p.old_vertex.end(), old_vertex);
    if (weight) std::copy(p.weight.begin(), p.weight.end(), weight);
  });
}
int ifem_tria_get_levels(const ifem_tria *t, int *levels)
{
  return guard([&] {
    const int nc = t->t.n_cells();
    for (int c = 0; c < nc; ++c) levels[c] = t->t.cell_level.empty() ? 0 : t->t.cell_level[c];
  });
}
int ifem_tria_get_hanging(const ifem_tria *t, int64_t *n_hanging, int *vertex, int *n_masters, int *masters)
{
  return guard([&] {
    const auto &h = t->t.hanging;
    *n_hanging = (int64_t)h.size();
    if (!vertex) return;
    for (size_t k = 0; k < h.size(); ++k)
      {
        vertex[k] = h[k].vertex;
        n_masters[k] = h[k].n_masters;
        for (int j = 0; j < 4; ++j) masters[4 * k + j] = h[k].master[j];
      }
  });
}
int ifem_tria_set_material_ids(ifem_tria *t, const int *ids, int64_t n)
{
  return guard([&] {
    if (n != (int64_t)t->t.n_cells()) throw std::runtime_error("ifem_tria_set_material_ids: one id per active cell expected");
    t->t.material_id.assign(ids, ids + n);
  });
}
int ifem_tria_flow_around_cylinder(ifem_tria *t)
{
  return guard([&] { GridCreator::flow_around_cylinder(t->t, t->t.dim); });
}
int ifem_tria_get_mesh(const ifem_tria *t, double *vertices, int *cells, int *boundary_faces)
{
  return guard([&] {
    if (vertices) std::copy(t->t.vertices.begin(), t->t.vertices.end(), vertices);
    if (cells) std::copy(t->t.cells.begin(), t->t.cells.end(), cells);
    if (boundary_faces) std::copy(t->t.boundary_faces.begin(), t->t.boundary_faces.end(), boundary_faces);
  });
}
int ifem_tria_set_mesh(ifem_tria *t, int64_t n_vertices, const double *vertices, int64_t n_cells, const int *cells, int64_t n_boundary_faces,
                       const int *boundary_faces, const int *material_ids)
{
  return guard([&] {
    Triangulation &T = t->t;
    const int dim = T.dim, vpc = 1 << dim;
    if (n_vertices <= 0 || n_cells <= 0 || !vertices || !cells) throw std::runtime_error("ifem_tria_set_mesh: empty mesh");
    for (int64_t k = 0; k < n_cells * vpc; ++k)
      if (cells[k] < 0 || cells[k] >= n_vertices) throw std::runtime_error("ifem_tria_set_mesh: vertex index out of range");
    for (int64_t f = 0; f < n_boundary_faces; ++f)
      if (boundary_faces[3 * f] < 0 || boundary_faces[3 * f] >= n_cells || boundary_faces[3 * f + 1] < 0 || boundary_faces[3 * f + 1] >= 2 * dim)
        throw std::runtime_error("ifem_tria_set_mesh: boundary face out of range");
    // every cell must be positively oriented in the lexicographic corner order (the Q1 map of the assembly kernels)
    for (int64_t c = 0; c < n_cells; ++c)
      {
        double e[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
        for (int d = 0; d < dim; ++d)
          for (int k = 0; k < dim; ++k) e[d][k] = vertices[(size_t)cells[c * vpc + (1 << d)] * dim + k] - vertices[(size_t)cells[c * vpc] * dim + k];
        const double det = dim == 2 ? e[0][0] * e[1][1] - e[0][1] * e[1][0]
                                    : e[0][0] * (e[1][1] * e[2][2] - e[1][2] * e[2][1]) - e[0][1] * (e[1][0] * e[2][2] - e[1][2] * e[2][0]) +
                                        e[0][2] * (e[1][0] * e[2][1] - e[1][1] * e[2][0]);
        if (!(det > 0)) throw std::runtime_error("ifem_tria_set_mesh: cell " + std::to_string(c) + " is inverted or degenerate in lexicographic corner order");
      }
    T = Triangulation();
    T.dim = dim;
    T.vertices.assign(vertices, vertices + n_vertices * dim);
    T.cells.assign(cells, cells + n_cells * vpc);
    if (n_boundary_faces) T.boundary_faces.assign(boundary_faces, boundary_faces + 3 * n_boundary_faces);
    if (material_ids) T.material_id.assign(material_ids, material_ids + n_cells);
    else T.material_id.assign((size_t)n_cells, 1);
    T.find_hanging_vertices(); // a mesh that arrives locally refined keeps its hanging vertices
  });
}
int ifem_tria_counts(const ifem_tria *t, int64_t *nv, int64_t *nc, int64_t *nbf)
{
  return guard([&] {
    if (nv) *nv = t->t.n_vertices();
    if (nc) *nc = t->t.n_cells();
    if (nbf) *nbf = t->t.n_boundary_faces();
  });
}

int ifem_params_from_file(const char *prm_file, ifem_params **out)
{
  return guard([&] {
    auto *p = new ifem_params;
    p->p.reset(new Parameters::AllParameters(std::string(prm_file)));
    *out = p;
  });
}
int ifem_params_from_text(const char *text, ifem_params **out)
{
  return guard([&] {
    auto *p = new ifem_params;
    p->p.reset(new Parameters::AllParameters(Parameters::AllParameters::from_text(std::string(text))));
    *out = p;
  });
}
int ifem_params_destroy(ifem_params *p)
{
  delete p;
  return IFEM_OK;
}

int ifem_insim_create(ifem_tria *tria, const ifem_params *params, ifem_insim **out)
{
  return guard([&] {
    require_device();
    auto *h = new ifem_insim;
    h->s.reset(new InsIM(default_context(), tria->t, *params->p));
    *out = h;
  });
}
int ifem_insim_destroy(ifem_insim *s)
{
  delete s;
  return IFEM_OK;
}
int ifem_insim_add_hard_coded_boundary_condition(ifem_insim *s, int boundary_id, ifem_bc_fn f, void *user)
{
  return guard([&] {
    s->s->add_hard_coded_boundary_condition(boundary_id, [f, user](const double *p, unsigned int c, double time) { return f(p, c, time, user); });
  });
}
int ifem_insim_default_control(int serial_twin, ifem_ins_control *out)
{
  return guard([&] {
    const InsSolverControl c = serial_twin ? InsSolverControl::serial() : InsSolverControl();
    out->fgmres_rel = c.fgmres_rel;
    out->fgmres_floor = c.fgmres_floor;
    out->cg_mp_rel = c.cg_mp_rel;
    out->cg_sm_rel = c.cg_sm_rel;
    out->cg_floor = c.cg_floor;
    out->a_inv_rel = c.a_inv_rel;
    out->a_inv_max_it = c.a_inv_max_it;
    out->basis_size = c.basis_size;
    out->a_inv_fp32 = c.a_inv_fp32;
    out->cg_sm_fp32 = c.cg_sm_fp32;
    out->supg_ilu = c.supg_ilu;
  });
}
int ifem_insim_get_control(const ifem_insim *s, ifem_ins_control *out)
{
  return guard([&] {
    const InsSolverControl &c = s->s->control;
    out->fgmres_rel = c.fgmres_rel;
    out->fgmres_floor = c.fgmres_floor;
    out->cg_mp_rel = c.cg_mp_rel;
    out->cg_sm_rel = c.cg_sm_rel;
    out->cg_floor = c.cg_floor;
    out->a_inv_rel = c.a_inv_rel;
    out->a_inv_max_it = c.a_inv_max_it;
    out->basis_size = c.basis_size;
    out->a_inv_fp32 = c.a_inv_fp32;
    out->cg_sm_fp32 = c.cg_sm_fp32;
    out->supg_ilu = c.supg_ilu;
  });
}
int ifem_insim_set_control(ifem_insim *s, const ifem_ins_control *c)
{
  return guard([&] {
    InsSolverControl &k = s->s->control;
    k.fgmres_rel = c->fgmres_rel;
    k.fgmres_floor = c->fgmres_floor;
    k.cg_mp_rel = c->cg_mp_rel;
    k.cg_sm_rel = c->cg_sm_rel;
    k.cg_floor = c->cg_floor;
    k.a_inv_rel = c->a_inv_rel;
    k.a_inv_max_it = c->a_inv_max_it;
    k.basis_size = c->basis_size;
    k.a_inv_fp32 = c->a_inv_fp32;
    k.cg_sm_fp32 = c->cg_sm_fp32;
    k.supg_ilu = c->supg_ilu;
  });
}
int ifem_insim_set_verbose(ifem_insim *s, int verbose)
{
  s->s->verbose = verbose != 0;
  return IFEM_OK;
}
int ifem_insim_setup(ifem_insim *s)
{
  return guard([&] {
    s->s->setup_dofs();
    s->s->make_constraints();
    s->s->initialize_system();
  });
}
int ifem_insim_setup_with_refinement(ifem_insim *s)
{
  return guard([&] {
    InsIM &m = *s->s;
    if (m.dofs_ready) return;
    m.triangulation.refine_global(m.parameters.global_refinements.empty() ? 0 : m.parameters.global_refinements[0]);
    m.setup_dofs();
    m.make_constraints();
    m.initialize_system();
  });
}
int ifem_insim_run(ifem_insim *s)
{
  return guard([&] { s->s->run(); });
}
int ifem_insim_run_one_step(ifem_insim *s, int nz)
{
  return guard([&] { s->s->run_one_step(nz != 0); });
}
int ifem_insim_assemble(ifem_insim *s, int nz)
{
  return guard([&] {
    s->s->assemble(nz != 0);
    IFEM_CUDA(cudaStreamSynchronize(s->s->ctx.stream));
  });
}
int ifem_insim_solve(ifem_insim *s, int nz, unsigned int *its, double *res)
{
  return guard([&] {
    fill(s->s->ctx, s->s->fs.n_dofs, 0.0, s->s->newton_update.p);
    const auto r = s->s->solve(nz != 0);
    IFEM_CUDA(cudaStreamSynchronize(s->s->ctx.stream));
    if (its) *its = r.first;
    if (res) *res = r.second;
  });
}
int ifem_insim_sizes(const ifem_insim *s, int64_t *n_u, int64_t *n_p, int64_t *nnz, int64_t *nnz_mp, int64_t *nnz_schur)
{
  return guard([&] {
    const FluidSpace &fs = s->s->fs;
    if (n_u) *n_u = fs.n_u;
    if (n_p) *n_p = fs.n_p;
    if (nnz) *nnz = fs.A_uu.nnz() + fs.A_up.nnz() + fs.A_pu.nnz() + fs.A_pp.nnz();
    if (nnz_mp) *nnz_mp = fs.M_p.nnz();
    if (nnz_schur) *nnz_schur = fs.S_m.nnz();
  });
}
int ifem_insim_support_points(const ifem_insim *s, double *pts)
{
  return guard([&] {
    const FluidSpace &fs = s->s->fs;
    const int dim = fs.dim;
    for (int n = 0; n < fs.un.n_nodes; ++n)
      for (int c = 0; c < dim; ++c)
        for (int d = 0; d < dim; ++d) pts[((size_t)dim * n + c) * dim + d] = fs.un.coords[(size_t)n * dim + d];
    for (int n = 0; n < fs.pn.n_nodes; ++n)
      for (int d = 0; d < dim; ++d) pts[((size_t)fs.n_u + n) * dim + d] = fs.pn.coords[(size_t)n * dim + d];
  });
}
int ifem_insim_get_current_solution(ifem_insim *s, double *host)
{
  return guard([&] { s->s->present_solution.download(host, s->s->fs.n_dofs, s->s->ctx.stream); });
}

int ifem_insim_set_vector(ifem_insim *s, int which, const double *host)
{
  return guard([&] {
    int64_t n;
    DevBuf<double> *v = pick_vector(*s->s, which, n);
    v->upload(host, n, s->s->ctx.stream);
    IFEM_CUDA(cudaStreamSynchronize(s->s->ctx.stream));
  });
}
int ifem_insim_get_vector(ifem_insim *s, int which, double *host)
{
  return guard([&] {
    int64_t n;
    DevBuf<double> *v = pick_vector(*s->s, which, n);
    v->download(host, n, s->s->ctx.stream);
  });
}
int ifem_insim_set_indicator(ifem_insim *s, const int *ind)
{
  return guard([&] {
    s->s->fs.d_indicator.upload(ind, s->s->fs.n_cells, s->s->ctx.stream);
    IFEM_CUDA(cudaStreamSynchronize(s->s->ctx.stream));
  });
}

int ifem_insim_get_matrix(ifem_insim *s, int which, int64_t *rowptr, int *col, double *val)
{
  return guard([&] {
    FluidSpace &fs = s->s->fs;
    cudaStream_t st = s->s->ctx.stream;
    std::vector<int64_t> rp;
    std::vector<int> ci;
    std::vector<double> v;
    if (which == 1 || which == 2)
      {
        (which == 1 ? fs.M_p : fs.S_m).to_host_csr(st, rp, ci, v);
        std::copy(rp.begin(), rp.end(), rowptr);
        std::copy(ci.begin(), ci.end(), col);
        std::copy(v.begin(), v.end(), val);
        return;
      }
    if (which == 3) // system matrix of the attached turbulence model (pattern of M_p)
      {
        SCnsIM *f = dynamic_cast<SCnsIM *>(s->s.get());
        if (!f || !f->turbulence_model) throw std::runtime_error("no turbulence model attached");
        f->turbulence_model->system_matrix.to_host_csr(st, rp, ci, v);
        std::copy(rp.begin(), rp.end(), rowptr);
        std::copy(ci.begin(), ci.end(), col);
        std::copy(v.begin(), v.end(), val);
        return;
      }
    if (which != 0) throw std::runtime_error("unknown matrix id");
    HostCsr uu, up, pu, ppm;
    fs.A_uu.to_host_csr(st, uu.rp, uu.ci, uu.v);
    fs.A_up.to_host_csr(st, up.rp, up.ci, up.v);
    fs.A_pu.to_host_csr(st, pu.rp, pu.ci, pu.v);
    const bool has_pp = fs.A_pp.n_brows > 0;
    if (has_pp) fs.A_pp.to_host_csr(st, ppm.rp, ppm.ci, ppm.v);
    int64_t pos = 0;
    for (int64_t r = 0; r < fs.n_u; ++r)
      {
        rowptr[r] = pos;
        for (int64_t k = uu.rp[r]; k < uu.rp[r + 1]; ++k, ++pos) { col[pos] = uu.ci[k]; val[pos] = uu.v[k]; }
        for (int64_t k = up.rp[r]; k < up.rp[r + 1]; ++k, ++pos) { col[pos] = (int)(fs.n_u + up.ci[k]); val[pos] = up.v[k]; }
      }
    for (int64_t r = 0; r < fs.n_p; ++r)
      {
        rowptr[fs.n_u + r] = pos;
        for (int64_t k = pu.rp[r]; k < pu.rp[r + 1]; ++k, ++pos) { col[pos] = pu.ci[k]; val[pos] = pu.v[k]; }
        if (has_pp)
          for (int64_t k = ppm.rp[r]; k < ppm.rp[r + 1]; ++k, ++pos) { col[pos] = (int)(fs.n_u + ppm.ci[k]); val[pos] = ppm.v[k]; }
      }
    rowptr[fs.n_dofs] = pos;
  });
}

int ifem_insim_vmult(ifem_insim *s, const double *x_host, double *y_host)
{
  return guard([&] {
    InsIM &m = *s->s;
    DevBuf<double> x(m.fs.n_dofs), y(m.fs.n_dofs);
    x.upload(x_host, m.fs.n_dofs, m.ctx.stream);
    block_vmult(m.ctx, m.fs, x.p, y.p);
    y.download(y_host, m.fs.n_dofs, m.ctx.stream);
  });
}

int ifem_insim_history(const ifem_insim *s, int max_records, ifem_newton_record *out, int *n_records)
{
  return guard([&] {
    const auto &h = s->s->history;
    *n_records = (int)h.size();
    const int first = std::max(0, (int)h.size() - max_records);
    for (int i = first; i < (int)h.size(); ++i)
      {
        ifem_newton_record &r = out[i - first];
        r.timestep = h[i].timestep; r.iteration = h[i].iteration; r.abs_res = h[i].abs_res; r.rel_res = h[i].rel_res;
        r.gmres_its = h[i].gmres_its; r.gmres_res = h[i].gmres_res; r.cg_mp_its = h[i].cg_mp_its; r.cg_sm_its = h[i].cg_sm_its;
        r.a_inv_its = h[i].a_inv_its; r.precond_applies = h[i].precond_applies; r.true_res = h[i].true_res;
      }
  });
}

int ifem_insim_timer_ms(const ifem_insim *s, const char *section, double *ms)
{
  return guard([&] {
    auto it = s->s->timer_ms.find(section);
    *ms = it == s->s->timer_ms.end() ? 0.0 : it->second;
  });
}
int ifem_insim_time(const ifem_insim *s, unsigned int *timestep, double *current)
{
  return guard([&] {
    if (timestep) *timestep = s->s->time.get_timestep();
    if (current) *current = s->s->time.current();
  });
}

int ifem_insim_bench_vmult(ifem_insim *s, int reps, double *ms, double *bytes)
{
  return guard([&] {
    InsIM &m = *s->s;
    DevBuf<double> y(m.fs.n_dofs);
    *ms = time_reps(m.ctx, reps, [&] { block_vmult(m.ctx, m.fs, m.fs.rhs.p, y.p); });
    *bytes = m.fs.A_uu.spmv_bytes() + m.fs.A_up.spmv_bytes() + m.fs.A_pu.spmv_bytes() + (m.fs.A_pp.n_brows ? m.fs.A_pp.spmv_bytes() : 0.0);
  });
}
int ifem_insim_bench_spmv_uu(ifem_insim *s, int reps, double *ms, double *bytes)
{
  return guard([&] {
    InsIM &m = *s->s;
    DevBuf<double> y(m.fs.n_u);
    *ms = time_reps(m.ctx, reps, [&] { spmv(m.ctx, m.fs.A_uu, m.fs.rhs.p, y.p); });
    *bytes = m.fs.A_uu.spmv_bytes();
  });
}
int ifem_insim_partition(const ifem_insim *s, int which, int *n_owned, int *n_local)
{
  return guard([&] {
    const FluidSpace &fs = s->s->fs;
    *n_owned = which == 0 ? fs.n_owned_unodes : fs.n_owned_pnodes;
    *n_local = which == 0 ? fs.un.n_nodes : fs.pn.n_nodes;
  });
}
int ifem_insim_local_to_global(const ifem_insim *s, int which, int *ids)
{
  return guard([&] {
    const FluidSpace &fs = s->s->fs;
    const int n = which == 0 ? fs.un.n_nodes : fs.pn.n_nodes;
    if (fs.n_ranks == 1)
      for (int i = 0; i < n; ++i) ids[i] = i;
    else
      {
        const auto &l = (which == 0 ? fs.part.u : fs.part.p).local_to_global;
        std::copy(l.begin(), l.end(), ids);
      }
  });
}
int ifem_insim_bench_steps(ifem_insim *s, int n_steps, int first_nz, double *ms_total)
{
  return guard([&] {
    InsIM &m = *s->s;
    int k = 0;
    *ms_total = n_steps * time_reps(m.ctx, n_steps, [&] { m.run_one_step(first_nz != 0 && k++ == 0); });
  });
}
int ifem_insim_bench_spmv_uu_fp32(ifem_insim *s, int reps, double *ms, double *bytes)
{
  return guard([&] {
    InsIM &m = *s->s;
    DevBuf<double> y(m.fs.n_u);
    make_fp32_copy(m.ctx, m.fs.A_uu);
    *ms = time_reps(m.ctx, reps, [&] { spmv_fp32(m.ctx, m.fs.A_uu, m.fs.rhs.p, y.p); });
    // algorithmic bytes of the fp32 stream (unpadded): 4 B per value instead of 8
    *bytes = m.fs.A_uu.spmv_bytes() - 4.0 * m.fs.A_uu.nnz();
  });
}
int ifem_insim_bench_spmv_uu_sell(ifem_insim *s, int precision, int variant, int reps, double *ms, double *bytes, double *padding,
                                  double *max_rel_err)
{
  return guard([&] {
    InsIM &m = *s->s;
    if (precision == 0) precision = m.inner32.S.built() ? m.inner32.S.precision : 32;
    if (!m.inner32.S.built() || m.inner32.S.precision != precision)
      m.inner32.setup(m.ctx, m.fs.A_uu, m.fs.un, m.fs.n_ranks > 1 ? &m.fs.halo_u : nullptr, precision);
    m.inner32.refresh(m.ctx, m.fs.A_uu, nullptr);
    const int keep = m.inner32.S.variant;
    if (variant > 0) m.inner32.S.variant = variant;
    m.inner32.probe_load(m.ctx, m.fs.rhs.p);
    for (int i = 0; i < 2; ++i) m.inner32.probe_apply(m.ctx);
    *ms = time_reps(m.ctx, reps, [&] { m.inner32.probe_apply(m.ctx); });
    m.inner32.S.variant = keep;
    *bytes = m.inner32.S.spmv_bytes();
    if (padding) *padding = m.inner32.S.padding();
    if (max_rel_err)
      {
        // against the fp64 product on the BCSR matrix (owned rows)
        DevBuf<double> y32(m.fs.n_u), y64(m.fs.n_u);
        m.inner32.probe_store(m.ctx, y32.p);
        m.fs.halo_u.update(m.ctx, m.fs.rhs.p);
        spmv(m.ctx, m.fs.A_uu, m.fs.rhs.p, y64.p);
        const std::vector<double> a = y32.to_host(m.ctx.stream), b = y64.to_host(m.ctx.stream);
        const size_t n_owned = (size_t)m.fs.n_owned_unodes * m.fs.dim;
        double scale = 0.0, err = 0.0;
        for (size_t i = 0; i < n_owned; ++i) scale = std::max(scale, std::fabs(b[i]));
        for (size_t i = 0; i < n_owned; ++i) err = std::max(err, std::fabs(a[i] - b[i]));
        *max_rel_err = scale > 0.0 ? err / scale : err;
      }
  });
}
int ifem_insim_solve_mass_schur(ifem_insim *s, int mode, const double *b, double rel_tol, int max_it, double *x, int *its, double *residual)
{
  return guard([&] {
    InsIM &m = *s->s;
    if (!m.fs.schur_valid) throw std::runtime_error("solve_mass_schur: S_m has not been formed yet (call solve first)");
    const size_t n = (size_t)m.fs.n_p;
    DevBuf<double> db(n), dx(n);
    db.upload(b, n, m.ctx.stream);
    dx.zero(m.ctx.stream);
    const double nrm = nrm2(m.ctx, m.fs.vs_p, db.p);
    const SolveResult r = m.solve_mass_schur(mode, db.p, nrm, dx.p, rel_tol * nrm, max_it);
    dx.download(x, n, m.ctx.stream);
    if (its) *its = r.iterations;
    if (residual) *residual = r.residual;
  });
}
int ifem_insim_set_inner_variant(ifem_insim *s, int variant)
{
  s->s->inner32.S.variant = variant;
  return IFEM_OK;
}
namespace
{
  // 16 independent FMA chains per thread, no memory traffic: the FP64 pipe's issue rate
  __global__ void __launch_bounds__(256) fp64_peak_kernel(int iters, double seed, double *out)
  {
    double a[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) a[k] = seed + threadIdx.x * 1e-6 + k;
    const double m = 1.0 - 1e-9, c = 1e-9;
    for (int i = 0; i < iters; ++i)
      {
#pragma unroll
        for (int k = 0; k < 16; ++k) a[k] = fma(a[k], m, c);
      }
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < 16; ++k) s += a[k];
    if (s == 12345.678) out[0] = s; // never true: keeps the chains alive
  }
} // namespace

int ifem_bench_fp64_peak(double *tflops)
{
  return guard([&] {
    require_device();
    Context &ctx = default_context();
    DevBuf<double> out(1);
    const int blocks = ctx.sm_count * 8, iters = 20000;
    fp64_peak_kernel<<<blocks, 256, 0, ctx.stream>>>(iters, 1.0, out.p);
    const double ms = time_reps(ctx, 3, [&] { fp64_peak_kernel<<<blocks, 256, 0, ctx.stream>>>(iters, 1.0, out.p); });
    IFEM_KERNEL_CHECK();
    *tflops = 2.0 * 16.0 * iters * 256.0 * blocks / (ms * 1e-3) / 1e12;
  });
}

int ifem_insim_bench_assemble(ifem_insim *s, int reps, double *ms)
{
  return guard([&] {
    InsIM &m = *s->s;
    *ms = time_reps(m.ctx, reps, [&] { m.assemble(false); });
  });
}

int ifem_hyper_create(ifem_tria *tria, const ifem_params *params, ifem_hyper **out)
{
  return guard([&] {
    require_device();
    auto *h = new ifem_hyper;
    h->s.reset(new HyperElasticity(default_context(), tria->t, *params->p));
    *out = h;
  });
}
int ifem_hyper_create_twin(ifem_tria *tria, const ifem_params *params, int shared, ifem_hyper **out)
{
  return guard([&] {
    require_device();
    auto *h = new ifem_hyper;
    h->s.reset(new HyperElasticity(default_context(), tria->t, *params->p, shared ? 1 : 0));
    *out = h;
  });
}
int ifem_linear_elasticity_create(ifem_tria *tria, const ifem_params *params, int shared, ifem_hyper **out)
{
  return guard([&] {
    require_device();
    auto *h = new ifem_hyper;
    h->s.reset(new LinearElasticity(default_context(), tria->t, *params->p, shared != 0));
    *out = h;
  });
}
int ifem_hyper_destroy(ifem_hyper *s)
{
  delete s;
  return IFEM_OK;
}
int ifem_hyper_set_verbose(ifem_hyper *s, int v)
{
  s->s->verbose = v != 0;
  return IFEM_OK;
}
int ifem_hyper_setup(ifem_hyper *s)
{
  return guard([&] {
    s->s->setup_dofs();
    s->s->initialize_system();
  });
}
int ifem_hyper_setup_with_refinement(ifem_hyper *s)
{
  return guard([&] {
    SolidSolver &m = *s->s;
    if (m.dofs_ready) return;
    m.triangulation.refine_global(m.parameters.global_refinements.size() > 1 ? m.parameters.global_refinements[1] : 0);
    m.setup_dofs();
    m.initialize_system();
  });
}
int ifem_hyper_run(ifem_hyper *s)
{
  return guard([&] { s->s->run(); });
}
int ifem_hyper_run_one_step(ifem_hyper *s, int first)
{
  return guard([&] { s->s->run_one_step(first != 0); });
}
int ifem_hyper_update_qph(ifem_hyper *s)
{
  return guard([&] {
    as_hyper(s).update_qph(s->s->current_displacement.p);
    IFEM_CUDA(cudaStreamSynchronize(s->s->ctx.stream));
  });
}
int ifem_hyper_assemble_system(ifem_hyper *s, int initial)
{
  return guard([&] {
    s->s->assemble_system(initial != 0);
    IFEM_CUDA(cudaStreamSynchronize(s->s->ctx.stream));
  });
}
int ifem_hyper_sizes(const ifem_hyper *s, int64_t *n_dofs, int64_t *nnz, int64_t *nqp, int *nsym)
{
  return guard([&] {
    const SolidSpace &ss = s->s->ss;
    if (n_dofs) *n_dofs = ss.n_dofs;
    if (nnz) *nnz = ss.K.nnz();
    if (nqp) *nqp = (int64_t)ss.n_cells * ss.nq;
    if (nsym) *nsym = ss.nsym;
  });
}
int ifem_hyper_get_current_solution(ifem_hyper *s, double *host)
{
  return guard([&] { s->s->current_displacement.download(host, s->s->ss.n_dofs, s->s->ctx.stream); });
}
int ifem_hyper_set_vector(ifem_hyper *s, int which, const double *host)
{
  return guard([&] {
    pick_solid_vector(*s->s, which)->upload(host, s->s->ss.n_dofs, s->s->ctx.stream);
    IFEM_CUDA(cudaStreamSynchronize(s->s->ctx.stream));
  });
}
int ifem_hyper_get_vector(ifem_hyper *s, int which, double *host)
{
  return guard([&] { pick_solid_vector(*s->s, which)->download(host, s->s->ss.n_dofs, s->s->ctx.stream); });
}
int ifem_hyper_get_matrix(ifem_hyper *s, int which, int64_t *rowptr, int *col, double *val)
{
  return guard([&] {
    std::vector<int64_t> rp;
    std::vector<int> ci;
    std::vector<double> v;
    Bcsr *A = which == 0 ? &s->s->ss.K : which == 1 ? &s->s->ss.M : nullptr;
    if (which == 2 || which == 3)
      {
        auto *lin = dynamic_cast<LinearElasticity *>(s->s.get());
        if (!lin) throw std::runtime_error("stiffness / damping matrices exist for LinearElasticity only");
        A = which == 2 ? &lin->stiffness_matrix : &lin->damping_matrix;
        if (which == 3 && !lin->shared) throw std::runtime_error("the damping matrix exists for SharedLinearElasticity only");
      }
    if (!A) throw std::runtime_error("unknown matrix id");
    A->to_host_csr(s->s->ctx.stream, rp, ci, v);
    std::copy(rp.begin(), rp.end(), rowptr);
    std::copy(ci.begin(), ci.end(), col);
    std::copy(v.begin(), v.end(), val);
  });
}
int ifem_hyper_get_qph(ifem_hyper *s, double *F_inv, double *tau, double *Jc, double *det_F)
{
  return guard([&] {
    const SolidSpace &ss = s->s->ss;
    cudaStream_t st = s->s->ctx.stream;
    if (F_inv) ss.d_Finv.download(F_inv, ss.d_Finv.n, st);
    if (tau) ss.d_tau.download(tau, ss.d_tau.n, st);
    if (Jc) ss.d_Jc.download(Jc, ss.d_Jc.n, st);
    if (det_F) ss.d_detF.download(det_F, ss.d_detF.n, st);
  });
}
int ifem_hyper_update_strain_and_stress(ifem_hyper *s)
{
  return guard([&] {
    s->s->update_strain_and_stress();
    IFEM_CUDA(cudaStreamSynchronize(s->s->ctx.stream));
  });
}
int ifem_hyper_get_nodal_tensor(ifem_hyper *s, int which, double *host)
{
  return guard([&] {
    DevBuf<double> &v = which == 0 ? s->s->stress : s->s->strain;
    v.download(host, v.n, s->s->ctx.stream);
  });
}
int ifem_hyper_set_nodal_tensor(ifem_hyper *s, int which, const double *host)
{
  return guard([&] {
    DevBuf<double> &v = which == 0 ? s->s->stress : s->s->strain;
    v.upload(host, v.n, s->s->ctx.stream);
    IFEM_CUDA(cudaStreamSynchronize(s->s->ctx.stream));
  });
}
int ifem_hyper_get_fsi_inputs(ifem_hyper *s, double *rows, double *vel, double *pres)
{
  return guard([&] {
    SolidSolver &m = *s->s;
    if (rows) m.fsi_stress_rows.download(rows, m.fsi_stress_rows.n, m.ctx.stream);
    if (vel) m.fluid_velocity.download(vel, m.fluid_velocity.n, m.ctx.stream);
    if (pres) m.fluid_pressure.download(pres, m.fluid_pressure.n, m.ctx.stream);
  });
}
int ifem_hyper_history(const ifem_hyper *s, int max_records, ifem_solid_record *out, int *n_records)
{
  return guard([&] {
    const auto &h = s->s->history;
    *n_records = (int)h.size();
    const int first = std::max(0, (int)h.size() - max_records);
    for (int i = first; i < (int)h.size(); ++i)
      {
        ifem_solid_record &r = out[i - first];
        r.timestep = h[i].timestep; r.iteration = h[i].iteration; r.res_F = h[i].res_F; r.res_U = h[i].res_U; r.cg_its = h[i].cg_its;
      }
  });
}

int ifem_fsi_create(ifem_insim *fluid, ifem_hyper *solid, const ifem_params *params, int use_dirichlet_bc, ifem_fsi **out)
{
  return guard([&] {
    require_device();
    auto *h = new ifem_fsi;
    h->f.reset(new FsiCoupling(default_context(), *fluid->s, *solid->s, *params->p, use_dirichlet_bc != 0));
    *out = h;
  });
}
int ifem_fsi_destroy(ifem_fsi *f)
{
  delete f;
  return IFEM_OK;
}
int ifem_fsi_update_solid_box(ifem_fsi *f, double *box)
{
  return guard([&] {
    const std::vector<double> b = f->f->update_solid_box();
    if (box) std::copy(b.begin(), b.end(), box);
  });
}
int ifem_fsi_update_indicator(ifem_fsi *f)
{
  return guard([&] {
    f->f->update_indicator();
    IFEM_CUDA(cudaStreamSynchronize(f->f->ctx.stream));
  });
}
int ifem_fsi_get_indicator(ifem_fsi *f, int *host)
{
  return guard([&] { f->f->fluid.fs.d_indicator.download(host, f->f->fluid.fs.n_cells, f->f->ctx.stream); });
}
int ifem_fsi_find_fluid_bc(ifem_fsi *f)
{
  return guard([&] {
    f->f->find_fluid_bc();
    IFEM_CUDA(cudaStreamSynchronize(f->f->ctx.stream));
  });
}
int ifem_fsi_get_inner_constraints(ifem_fsi *f, unsigned char *flags, double *inhom)
{
  return guard([&] {
    f->f->d_inner_con.download(flags, f->f->fluid.fs.n_dofs, f->f->ctx.stream);
    f->f->d_inner_inhom.download(inhom, f->f->fluid.fs.n_dofs, f->f->ctx.stream);
  });
}
int ifem_fsi_point_in_solid(ifem_fsi *f, int n, const double *points, int *inside)
{
  return guard([&] { f->f->point_in_solid(n, points, inside); });
}
int ifem_fsi_interpolate(ifem_fsi *f, int which, int n, const double *points, double *values, int *found)
{
  return guard([&] { f->f->interpolate(which, n, points, values, found); });
}
int ifem_fsi_find_solid_bc(ifem_fsi *f)
{
  return guard([&] {
    f->f->find_solid_bc();
    IFEM_CUDA(cudaStreamSynchronize(f->f->ctx.stream));
  });
}
int ifem_fsi_set_penetration_criterion(ifem_fsi *f, ifem_point_fn criterion, void *user, const double *direction)
{
  return guard([&] { f->f->set_penetration_criterion([criterion, user](const double *p) { return criterion(p, user); }, direction); });
}
int ifem_fsi_contact_iterations(const ifem_fsi *f, int *n)
{
  return guard([&] { *n = f->f->contact_iterations; });
}
int ifem_fsi_run_one_step(ifem_fsi *f, int first_step)
{
  return guard([&] { f->f->run_one_step(first_step != 0); });
}
int ifem_fsi_prepare_fluid_step(ifem_fsi *f, int first_step)
{
  return guard([&] { f->f->run_one_step(first_step != 0, true); });
}
int ifem_fsi_run(ifem_fsi *f)
{
  return guard([&] { f->f->run(); });
}
int ifem_set_spmv_short_variant(int key)
{
  return guard([&] { default_context().spmv_short = key; });
}
int ifem_ilu0_apply(int n, const int64_t *rowptr, const int *col, const double *val, const double *b, double *factors, double *x,
                    int *n_levels_lower, int *n_levels_upper)
{
  return guard([&] {
    require_device();
    Context &ctx = default_context();
    Ilu0 ilu;
    ilu.setup(ctx, std::vector<int64_t>(rowptr, rowptr + n + 1), std::vector<int>(col, col + rowptr[n]));
    ilu.val.upload(val, (size_t)rowptr[n], ctx.stream);
    ilu.factor(ctx);
    DevBuf<double> db((size_t)n), dx((size_t)n);
    db.upload(b, (size_t)n, ctx.stream);
    ilu.solve(ctx, db.p, dx.p);
    if (factors) ilu.val.download(factors, (size_t)rowptr[n], ctx.stream);
    dx.download(x, (size_t)n, ctx.stream);
    if (n_levels_lower) *n_levels_lower = ilu.n_levels_lower;
    if (n_levels_upper) *n_levels_upper = ilu.n_levels_upper;
  });
}
int ifem_fsi_refine_mesh(ifem_fsi *f, unsigned int min_grid_level, unsigned int max_grid_level)
{
  return guard([&] { f->f->refine_mesh(min_grid_level, max_grid_level); });
}
int ifem_fsi_bench_steps(ifem_fsi *f, int n_steps, int first_step, double *ms_total)
{
  return guard([&] {
    FsiCoupling &c = *f->f;
    int k = 0;
    *ms_total = n_steps * time_reps(c.ctx, n_steps, [&] { c.run_one_step(first_step != 0 && k++ == 0); });
  });
}
int ifem_fsi_timer_ms(const ifem_fsi *f, const char *section, double *ms)
{
  return guard([&] {
    auto it = f->f->timer_ms.find(section);
    *ms = it == f->f->timer_ms.end() ? 0.0 : it->second;
  });
}

int ifem_insim_set_output_directory(ifem_insim *s, const char *dir)
{
  return guard([&] { s->s->set_output_directory(dir ? dir : ""); });
}
int ifem_insim_output_results(ifem_insim *s, unsigned int output_index)
{
  return guard([&] { s->s->output_results(output_index); });
}
int ifem_insim_save_checkpoint(ifem_insim *s, int output_index)
{
  return guard([&] { s->s->save_checkpoint(output_index); });
}
int ifem_insim_load_checkpoint(ifem_insim *s, int *loaded)
{
  return guard([&] { *loaded = s->s->load_checkpoint() ? 1 : 0; });
}
int ifem_insim_get_time(const ifem_insim *s, double *time, unsigned int *timestep)
{
  return guard([&] {
    if (time) *time = s->s->time.current();
    if (timestep) *timestep = s->s->time.get_timestep();
  });
}
int ifem_write_vtu(const char *path, int dim, int64_t n_points, const double *points, int64_t n_cells, const int *cells,
                   int n_point_fields, const char *const *point_names, const int *point_ncomp, const double *const *point_data,
                   int n_cell_fields, const char *const *cell_names, const double *const *cell_data)
{
  return guard([&] {
    if (dim != 2 && dim != 3) throw std::runtime_error("ifem_write_vtu: dim must be 2 or 3");
    std::vector<io::Field> pd, cd;
    for (int k = 0; k < n_point_fields; ++k)
      pd.push_back({point_names[k], point_ncomp[k], std::vector<double>(point_data[k], point_data[k] + n_points * point_ncomp[k])});
    for (int k = 0; k < n_cell_fields; ++k) cd.push_back({cell_names[k], 1, std::vector<double>(cell_data[k], cell_data[k] + n_cells)});
    io::write_vtu(path, dim, std::vector<double>(points, points + n_points * dim), std::vector<int>(cells, cells + n_cells * (1 << dim)), pd, cd);
  });
}
int ifem_write_pvd(const char *path, const char *pvtu_prefix, int n, const double *times, const unsigned int *timesteps)
{
  return guard([&] {
    io::PVDWriter w(path);
    for (int k = 0; k < n; ++k) w.write_current_timestep(times[k], timesteps[k], pvtu_prefix, 6);
  });
}
int ifem_block_write(const char *path, int64_t n, const double *values)
{
  return guard([&] { io::block_write(path, std::vector<double>(values, values + n)); });
}
int ifem_block_read(const char *path, int64_t capacity, double *values, int64_t *n)
{
  return guard([&] {
    const std::vector<double> v = io::block_read(path);
    *n = (int64_t)v.size();
    if ((int64_t)v.size() > capacity) throw std::runtime_error("ifem_block_read: buffer too small");
    std::copy(v.begin(), v.end(), values);
  });
}
int ifem_fluid_write_results_host(const ifem_tria *tria, int pu, int pp, const double *present, const double *fsi_acceleration,
                                  const int *indicator, const double *stress, const char *dir, unsigned int output_index)
{
  return guard([&] {
    const Triangulation &t = tria->t;
    const NodeTable un = build_node_table(t, pu), pn = build_node_table(t, pp);
    const size_t n = (size_t)t.dim * un.n_nodes + pn.n_nodes;
    std::vector<int> all(t.n_cells());
    for (int c = 0; c < t.n_cells(); ++c) all[c] = c;
    io::write_fluid_results(dir ? dir : ".", output_index, 0, 1, t.dim, un, pn, all, std::vector<double>(present, present + n),
                            fsi_acceleration ? std::vector<double>(fsi_acceleration, fsi_acceleration + n) : std::vector<double>(),
                            indicator ? std::vector<int>(indicator, indicator + t.n_cells()) : std::vector<int>(),
                            stress ? std::vector<double>(stress, stress + (size_t)t.dim * t.dim * un.n_nodes) : std::vector<double>());
  });
}
int ifem_hyper_set_output_directory(ifem_hyper *s, const char *dir)
{
  return guard([&] { s->s->set_output_directory(dir ? dir : ""); });
}
int ifem_hyper_output_results(ifem_hyper *s, unsigned int output_index)
{
  return guard([&] { s->s->output_results(output_index); });
}
int ifem_hyper_save_checkpoint(ifem_hyper *s, int output_index)
{
  return guard([&] { s->s->save_checkpoint(output_index); });
}
int ifem_hyper_load_checkpoint(ifem_hyper *s, int *loaded)
{
  return guard([&] { *loaded = s->s->load_checkpoint() ? 1 : 0; });
}
int ifem_hyper_get_time(const ifem_hyper *s, double *time, unsigned int *timestep)
{
  return guard([&] {
    if (time) *time = s->s->time.current();
    if (timestep) *timestep = s->s->time.get_timestep();
  });
}

int ifem_insimex_create(ifem_tria *tria, const ifem_params *params, ifem_insim **out)
{
  return guard([&] {
    require_device();
    auto *h = new ifem_insim;
    h->s.reset(new InsIMEX(default_context(), tria->t, *params->p));
    *out = h;
  });
}
static InsIMEX &as_imex(ifem_insim *s)
{
  auto *p = dynamic_cast<InsIMEX *>(s->s.get());
  if (!p) throw std::runtime_error("this fluid solver is not an InsIMEX");
  return *p;
}
int ifem_insimex_assemble(ifem_insim *s, int nz, int assemble_system)
{
  return guard([&] {
    as_imex(s).assemble(nz != 0, assemble_system != 0);
    IFEM_CUDA(cudaStreamSynchronize(s->s->ctx.stream));
  });
}
int ifem_insimex_solve(ifem_insim *s, int nz, int assemble_system, unsigned int *its, double *res)
{
  return guard([&] {
    fill(s->s->ctx, s->s->fs.n_dofs, 0.0, s->s->newton_update.p);
    const auto r = as_imex(s).solve(nz != 0, assemble_system != 0);
    IFEM_CUDA(cudaStreamSynchronize(s->s->ctx.stream));
    if (its) *its = r.first;
    if (res) *res = r.second;
  });
}
int ifem_insimex_run_one_step(ifem_insim *s, int nz, int assemble_system)
{
  return guard([&] { as_imex(s).run_one_step(nz != 0, assemble_system != 0); });
}

int ifem_scnsim_create(ifem_tria *tria, const ifem_params *params, ifem_insim **out)
{
  return guard([&] {
    require_device();
    auto *h = new ifem_insim;
    h->s.reset(new SCnsIM(default_context(), tria->t, *params->p));
    *out = h;
  });
}
int ifem_supg_insim_create(ifem_tria *tria, const ifem_params *params, ifem_insim **out)
{
  return guard([&] {
    require_device();
    auto *h = new ifem_insim;
    h->s.reset(new SUPGInsIM(default_context(), tria->t, *params->p));
    *out = h;
  });
}
int ifem_scnsim_set_body_force(ifem_insim *s, ifem_field_fn f, void *user)
{
  return guard([&] { as_scns(s).set_body_force([f, user](const double *p, unsigned c) { return f(p, c, user); }); });
}
int ifem_scnsim_set_sigma_pml_field(ifem_insim *s, ifem_field_fn f, void *user)
{
  return guard([&] { as_scns(s).set_sigma_pml_field([f, user](const double *p, unsigned c) { return f(p, c, user); }); });
}
int ifem_scnsim_set_initial_condition(ifem_insim *s, ifem_field_fn f, void *user)
{
  return guard([&] { as_scns(s).set_initial_condition([f, user](const double *p, unsigned c) { return f(p, c, user); }); });
}
int ifem_scnsim_update_stress(ifem_insim *s)
{
  return guard([&] {
    s->s->update_stress();
    IFEM_CUDA(cudaStreamSynchronize(s->s->ctx.stream));
  });
}
namespace
{
  SpalartAllmaras &as_turbulence(ifem_insim *s)
  {
    SCnsIM &f = as_scns(s);
    if (!f.turbulence_model) throw std::runtime_error("no turbulence model attached");
    return *f.turbulence_model;
  }
  DevBuf<double> &turbulence_vector(SpalartAllmaras &t, int which)
  {
    switch (which)
      {
      case 0: return t.present_solution;
      case 1: return t.evaluation_point;
      case 2: return t.eddy_viscosity;
      case 3: return t.system_rhs;
      case 4: return t.newton_update;
      case 5: return t.fixed_wall_distance;
      default: throw std::runtime_error("unknown turbulence vector id");
      }
  }
} // namespace
int ifem_insim_attach_turbulence_model(ifem_insim *s, const char *model_name)
{
  return guard([&] { as_scns(s).attach_turbulence_model(model_name ? model_name : ""); });
}
int ifem_turbulence_get_vector(ifem_insim *s, int which, double *host)
{
  return guard([&] {
    SpalartAllmaras &t = as_turbulence(s);
    DevBuf<double> &v = turbulence_vector(t, which);
    v.download(host, v.n, t.ctx.stream);
  });
}
int ifem_turbulence_set_vector(ifem_insim *s, int which, const double *host)
{
  return guard([&] {
    if (which < 0 || which > 2) throw std::runtime_error("turbulence vector is read-only");
    SpalartAllmaras &t = as_turbulence(s);
    DevBuf<double> &v = turbulence_vector(t, which);
    v.upload(host, v.n, t.ctx.stream);
    IFEM_CUDA(cudaStreamSynchronize(t.ctx.stream));
  });
}
int ifem_turbulence_assemble(ifem_insim *s, int use_nonzero_constraints)
{
  return guard([&] {
    SpalartAllmaras &t = as_turbulence(s);
    t.assemble(use_nonzero_constraints != 0);
    IFEM_CUDA(cudaStreamSynchronize(t.ctx.stream));
  });
}
int ifem_turbulence_run_one_step(ifem_insim *s, int apply_nonzero_constraints)
{
  return guard([&] {
    SpalartAllmaras &t = as_turbulence(s);
    t.run_one_step(apply_nonzero_constraints != 0);
    IFEM_CUDA(cudaStreamSynchronize(t.ctx.stream));
  });
}
int ifem_turbulence_update_boundary_condition(ifem_insim *s, int first_step)
{
  return guard([&] {
    SpalartAllmaras &t = as_turbulence(s);
    t.update_boundary_condition(first_step != 0);
    IFEM_CUDA(cudaStreamSynchronize(t.ctx.stream));
  });
}
int ifem_turbulence_get_shear_velocity(ifem_insim *s, double vel, double init_guess, double *out)
{
  return guard([&] { *out = as_turbulence(s).get_shear_velocity(vel, init_guess); });
}
int ifem_turbulence_history(ifem_insim *s, int max_records, double *abs_res, int *gmres_its, int *n_records)
{
  return guard([&] {
    const SpalartAllmaras &t = as_turbulence(s);
    const int n = (int)t.history.size(), k = std::min(n, max_records);
    for (int i = 0; i < k; ++i)
      {
        abs_res[i] = t.history[n - k + i].abs_res;
        gmres_its[i] = t.history[n - k + i].gmres_its;
      }
    *n_records = n;
  });
}
int ifem_scnsim_get_field(ifem_insim *s, int which, double *host)
{
  return guard([&] {
    InsIM &m = *s->s;
    DevBuf<double> &v = which == 0 ? m.stress : as_scns(s).fsi_stress;
    v.download(host, v.n, m.ctx.stream);
  });
}
int ifem_scnsim_set_field(ifem_insim *s, int which, const double *host)
{
  return guard([&] {
    InsIM &m = *s->s;
    DevBuf<double> &v = which == 0 ? m.stress : as_scns(s).fsi_stress;
    v.upload(host, v.n, m.ctx.stream);
    IFEM_CUDA(cudaStreamSynchronize(m.ctx.stream));
  });
}
} // extern "C"
