#include "fluid.h"

#include "comm.h"

#include <algorithm>
#include <cmath>
#include <cstring>

namespace ifem
{
  // ===========================================================================
  // Setup: setup_dofs + initialize_system (mpi_fluid_solver.cpp:116-162, 305-365)
  // ===========================================================================
  namespace
  {
    // slot of column node B in the sorted column list of block row A
    __global__ void build_slots_kernel(int n_cells, int nr, int nc, const int *__restrict__ row_tab, const int *__restrict__ col_tab,
                                       const int64_t *__restrict__ rowptr, const int *__restrict__ col, int n_rows_owned,
                                       unsigned char *__restrict__ slots, int stride, int offset, int *__restrict__ err)
    {
      const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
      const int64_t total = (int64_t)n_cells * nr * nc;
      if (t >= total) return;
      const int cell = (int)(t / (nr * nc));
      const int rem = (int)(t % (nr * nc));
      const int a = rem / nc, b = rem % nc;
      const int A = row_tab[(int64_t)cell * nr + a], B = col_tab[(int64_t)cell * nc + b];
      if (A >= n_rows_owned) // ghost row: assembled by its owner
        {
          slots[(int64_t)cell * stride + offset + rem] = 0;
          return;
        }
      const int64_t base = rowptr[A];
      int lo = 0, hi = (int)(rowptr[A + 1] - base) - 1, j = -1;
      while (lo <= hi)
        {
          const int mid = (lo + hi) >> 1;
          const int c = col[base + mid];
          if (c == B) { j = mid; break; }
          if (c < B) lo = mid + 1; else hi = mid - 1;
        }
      if (j < 0 || j > 255) { atomicExch(err, 1); j = 0; }
      slots[(int64_t)cell * stride + offset + rem] = (unsigned char)j;
    }
  } // namespace

  void FluidSpace::setup(Context &ctx, const Triangulation &tria, int pu_, int pp_, bool with_App)
  {
    dim = tria.dim;
    pu = pu_;
    pp = pp_;
    n_cells = tria.n_cells();
    nv = 1 << dim;
    fe_u = FEQ(dim, pu);
    fe_p = FEQ(dim, pp);
    fe_geo = FEQ(dim, 1);
    nu = fe_u.n;
    np = fe_p.n;
    quad = Quadrature(dim, pu + 1);
    nq = quad.nq;
    tab_u = ShapeTable(fe_u, quad.points, nq);
    tab_p = ShapeTable(fe_p, quad.points, nq);
    tab_geo = ShapeTable(fe_geo, quad.points, nq);
    rank = ctx.comm ? ctx.comm->rank : 0;
    n_ranks = ctx.comm ? ctx.comm->size : 1;
    if (n_ranks > 1)
      {
        un_global = build_node_table(tria, pu);
        pn_global = build_node_table(tria, pp);
        part = build_partition(tria, un_global, pn_global, rank, n_ranks);
        local_cells = part.local_cells;
        un = localise(un_global, local_cells, part.u);
        pn = localise(pn_global, local_cells, part.p);
        n_owned_unodes = part.u.n_owned;
        n_owned_pnodes = part.p.n_owned;
        n_layer1_unodes = part.u.n_layer1;
        n_layer1_pnodes = part.p.n_layer1;
      }
    else
      {
        un = build_node_table(tria, pu);
        pn = build_node_table(tria, pp);
        local_cells.resize(n_cells);
        for (int c = 0; c < n_cells; ++c) local_cells[c] = c;
        n_owned_unodes = un.n_nodes;
        n_owned_pnodes = pn.n_nodes;
        n_layer1_unodes = un.n_nodes;
        n_layer1_pnodes = pn.n_nodes;
        part = Partition();
        part.u.n_owned = part.u.n_layer1 = part.u.n_local = un.n_nodes;
        part.p.n_owned = part.p.n_layer1 = part.p.n_local = pn.n_nodes;
        part.cell_layer.assign(n_cells, 1);
      }
    n_cells = (int)local_cells.size();
    n_u = (int64_t)dim * un.n_nodes;
    n_p = pn.n_nodes;
    n_dofs = n_u + n_p;
    vs_all = VecSpace((int64_t)dim * n_owned_unodes, n_u, n_owned_pnodes, n_dofs);
    vs_u = VecSpace((int64_t)dim * n_owned_unodes, 0, 0, n_u);
    vs_p = VecSpace(n_owned_pnodes, 0, 0, n_p);
    halo_u.init(ctx, part.u, dim);
    halo_p.init(ctx, part.p, 1);
    halo_s.init(ctx, part.u, 1); // scalar fields on the velocity nodes (nodal stress)

    // patterns: rows = owned nodes (they come first in the local numbering), columns = local nodes
    auto owned_rows = [](Pattern P, int n_owned) {
      P.n_rows = n_owned;
      P.rowptr.resize((size_t)n_owned + 1);
      P.col.resize(P.rowptr[n_owned]);
      return P;
    };
    // locally refined mesh: the masters of a cell's hanging nodes couple with the cell's nodes (extended per-cell lists)
    std::vector<int> ext_un, ext_pn;
    int ext_w = 0;
    hanging.find(tria, *this, ext_un, ext_pn, ext_w);
    const int *tu = hanging.active ? ext_un.data() : un.cell_nodes.data(), *tp = hanging.active ? ext_pn.data() : pn.cell_nodes.data();
    const int wu = hanging.active ? ext_w : nu, wp = hanging.active ? ext_w : np;
    P_uu = owned_rows(build_pattern(n_cells, tu, wu, un.n_nodes, tu, wu, un.n_nodes), n_owned_unodes);
    // A_up keeps the rows of the layer-1 ghost velocity nodes as well: B^T rows the explicit Schur complement of
    // the owned pressure rows needs (complete, because every cell around a layer-1 node is local)
    P_up = owned_rows(build_pattern(n_cells, tu, wu, un.n_nodes, tp, wp, pn.n_nodes), n_layer1_unodes);
    P_pu = owned_rows(build_pattern(n_cells, tp, wp, pn.n_nodes, tu, wu, un.n_nodes), n_owned_pnodes);
    P_pp = owned_rows(build_pattern(n_cells, tp, wp, pn.n_nodes, tp, wp, pn.n_nodes), n_owned_pnodes);
    colour_cells(n_cells, un.cell_nodes.data(), nu, un.n_nodes, colour_order, colour_offsets);
    // inside every colour: cells that touch an owned node first (the ones every assembly visits)
    colour_n1.assign(colour_offsets.size() - 1, 0);
    for (size_t k = 0; k + 1 < colour_offsets.size(); ++k)
      {
        auto b = colour_order.begin() + colour_offsets[k], e = colour_order.begin() + colour_offsets[k + 1];
        auto mid = std::stable_partition(b, e, [&](int c) { return part.cell_layer[c] == 1; });
        colour_n1[k] = (int)(mid - b);
      }

    cudaStream_t s = ctx.stream;
    d_cell_un.upload(un.cell_nodes, s);
    d_cell_pn.upload(pn.cell_nodes, s);
    d_colour_order.upload(colour_order, s);
    {
      std::vector<double> cx((size_t)n_cells * nv * dim);
      for (int c = 0; c < n_cells; ++c)
        for (int v = 0; v < nv; ++v)
          for (int d = 0; d < dim; ++d)
            cx[((size_t)c * nv + v) * dim + d] = tria.vertices[(size_t)tria.cells[(size_t)local_cells[c] * nv + v] * dim + d];
      d_cell_x.upload(cx, s);
      IFEM_CUDA(cudaStreamSynchronize(s));
    }
    {
      std::vector<double> t;
      t.insert(t.end(), tab_u.N.begin(), tab_u.N.end());
      t.insert(t.end(), tab_u.dN.begin(), tab_u.dN.end());
      t.insert(t.end(), tab_p.N.begin(), tab_p.N.end());
      t.insert(t.end(), tab_geo.dN.begin(), tab_geo.dN.end());
      t.insert(t.end(), quad.weights.begin(), quad.weights.end());
      d_tables.upload(t, s);
      std::vector<double> ts;
      ts.insert(ts.end(), tab_u.N.begin(), tab_u.N.end());
      ts.insert(ts.end(), tab_p.N.begin(), tab_p.N.end());
      ts.insert(ts.end(), tab_geo.dN.begin(), tab_geo.dN.end());
      ts.insert(ts.end(), quad.weights.begin(), quad.weights.end());
      if (ts.size() % 2) ts.push_back(0.0);
      d_tables_s.upload(ts, s);
      IFEM_CUDA(cudaStreamSynchronize(s));
    }
    A_uu.init(P_uu, dim, dim, s);
    A_up.init(P_up, dim, 1, s);
    A_up.n_brows_spmv = n_owned_unodes;
    A_pu.init(P_pu, 1, dim, s);
    if (with_App) A_pp.init(P_pp, 1, 1, s);
    M_p.init(P_pp, 1, 1, s);
    diag_Mu.alloc(n_u);
    rhs.alloc(n_dofs);
    d_indicator.alloc(n_cells);
    d_indicator.zero(s);

    // per-cell slot tables
    const int spc = slots_per_cell();
    d_slots.alloc((size_t)n_cells * spc);
    DevBuf<int> err(1);
    err.zero(s);
    auto launch = [&](int nr, int nc, const DevBuf<int> &rt, const DevBuf<int> &ct, const Bcsr &M, int offset) {
      const int64_t total = (int64_t)n_cells * nr * nc;
      const int threads = 256;
      build_slots_kernel<<<(unsigned)((total + threads - 1) / threads), threads, 0, s>>>(n_cells, nr, nc, rt.p, ct.p, M.rowptr.p,
                                                                                         M.col.p, M.n_brows, d_slots.p, spc, offset, err.p);
      IFEM_KERNEL_CHECK();
    };
    launch(nu, nu, d_cell_un, d_cell_un, A_uu, 0);
    launch(nu, np, d_cell_un, d_cell_pn, A_up, nu * nu);
    launch(np, nu, d_cell_pn, d_cell_un, A_pu, nu * nu + nu * np);
    launch(np, np, d_cell_pn, d_cell_pn, M_p, nu * nu + 2 * nu * np);
    if (err.to_host(s)[0]) throw std::runtime_error("FluidSpace::setup: a matrix row has more than 256 block columns");
    hanging.plan(ctx, *this);

    // pattern of B B^T on the owned pressure rows (compute_mmult_pattern, mpi_fluid_solver.cpp:326-329)
    P_schur = product_pattern(P_pu, P_up, pn.n_nodes);
    S_m.init(P_schur, 1, 1, s);
    schur_valid = false;

    // qpt_to_dof = M^-1 Q^T W on the reference cell (FETools::compute_projection_from_quadrature_points_matrix,
    // mpi_fluid_solver.cpp:740-743) for the scalar FE_Q(pu) space
    {
      const int n = nu;
      std::vector<double> M((size_t)n * n, 0.0), R((size_t)n * nq, 0.0);
      for (int q = 0; q < nq; ++q)
        for (int i = 0; i < n; ++i)
          {
            R[(size_t)i * nq + q] = tab_u.N[(size_t)q * n + i] * quad.weights[q];
            for (int j = 0; j < n; ++j) M[(size_t)i * n + j] += tab_u.N[(size_t)q * n + i] * tab_u.N[(size_t)q * n + j] * quad.weights[q];
          }
      for (int c = 0; c < n; ++c) // Gauss-Jordan on [M | R]
        {
          int piv = c;
          for (int r2 = c + 1; r2 < n; ++r2)
            if (std::fabs(M[(size_t)r2 * n + c]) > std::fabs(M[(size_t)piv * n + c])) piv = r2;
          if (piv != c)
            {
              for (int k = 0; k < n; ++k) std::swap(M[(size_t)c * n + k], M[(size_t)piv * n + k]);
              for (int k = 0; k < nq; ++k) std::swap(R[(size_t)c * nq + k], R[(size_t)piv * nq + k]);
            }
          const double d = 1.0 / M[(size_t)c * n + c];
          for (int k = 0; k < n; ++k) M[(size_t)c * n + k] *= d;
          for (int k = 0; k < nq; ++k) R[(size_t)c * nq + k] *= d;
          for (int r2 = 0; r2 < n; ++r2)
            {
              if (r2 == c) continue;
              const double f = M[(size_t)r2 * n + c];
              if (f == 0.0) continue;
              for (int k = 0; k < n; ++k) M[(size_t)r2 * n + k] -= f * M[(size_t)c * n + k];
              for (int k = 0; k < nq; ++k) R[(size_t)r2 * nq + k] -= f * R[(size_t)c * nq + k];
            }
        }
      d_qpt_to_dof.upload(R, s);
      d_stress_count.alloc(un.n_nodes);
    }

    con.assign(n_dofs, 0);
    nonzero_val.assign(n_dofs, 0.0);
    base_valid = false; // cached constraint lines belong to the previous space
    flags_merged = false;
    d_con.upload(con, s);
    d_nonzero_val.upload(nonzero_val, s);
    IFEM_CUDA(cudaStreamSynchronize(s));
  }

  void FluidSpace::make_constraints(Context &ctx, const Triangulation &tria,
                                    const std::map<unsigned int, std::pair<unsigned int, std::vector<double>>> &dirichlet,
                                    const std::function<bool(int, const double *, int, double &)> &hard_coded)
  {
    // The pass runs over the GLOBAL boundary faces and node table (every rank holds the whole
    // triangulation), so "first boundary id wins" cannot depend on the partition; the flags are then
    // restricted to the local dofs.
    const NodeTable &g = n_ranks > 1 ? un_global : un;
    const std::vector<char> hflag = hanging.active ? hanging_node_flags(tria, g) : std::vector<char>();
    std::vector<unsigned char> gcon((size_t)dim * g.n_nodes, 0);
    std::vector<double> gval((size_t)dim * g.n_nodes, 0.0);
    std::vector<std::vector<int>> face_nodes(2 * dim);
    for (int f = 0; f < 2 * dim; ++f) face_nodes[f] = face_local_nodes(dim, pu, f);
    // std::map iterates boundary ids in ascending order; a dof already constrained
    // keeps its first value (VectorTools::interpolate_boundary_values semantics)
    for (const auto &bc : dirichlet)
      {
        const int id = (int)bc.first;
        const unsigned flag = bc.second.first;
        double aug[3] = {0, 0, 0};
        int k = 0;
        for (int c = 0; c < dim; ++c)
          if (flag & (1u << c)) aug[c] = bc.second.second.at(k++);
        for (int f = 0; f < tria.n_boundary_faces(); ++f)
          {
            if (tria.boundary_faces[3 * f + 2] != id) continue;
            const int cell = tria.boundary_faces[3 * f], face = tria.boundary_faces[3 * f + 1];
            for (int a : face_nodes[face])
              {
                const int node = g.cell_nodes[(size_t)cell * nu + a];
                if (!hflag.empty() && hflag[node]) continue; // keeps its hanging-node line (made first, mpi_fluid_solver.cpp:182-184)
                for (int c = 0; c < dim; ++c)
                  {
                    if (!(flag & (1u << c))) continue;
                    const int64_t gd = (int64_t)dim * node + c;
                    if (gcon[gd]) continue;
                    gcon[gd] = 1;
                    double v = aug[c];
                    if (hard_coded) hard_coded(id, &g.coords[(size_t)node * dim], c, v);
                    gval[gd] = v;
                  }
              }
          }
      }
    std::fill(con.begin(), con.end(), 0);
    std::fill(nonzero_val.begin(), nonzero_val.end(), 0.0);
    for (int l = 0; l < un.n_nodes; ++l)
      {
        const int gn = n_ranks > 1 ? part.u.local_to_global[l] : l;
        for (int c = 0; c < dim; ++c)
          {
            con[(size_t)dim * l + c] = gcon[(size_t)dim * gn + c];
            nonzero_val[(size_t)dim * l + c] = gval[(size_t)dim * gn + c];
          }
      }
    std::vector<int> idx;
    for (int64_t gd = 0; gd < n_dofs; ++gd)
      if (con[gd]) idx.push_back((int)gd);
    n_con = (int)idx.size();
    cudaStream_t s = ctx.stream;
    d_con.upload(con, s);
    d_nonzero_val.upload(nonzero_val, s);
    if (n_con) d_con_idx.upload(idx, s);
    if (d_base_con.n != d_con.n) d_base_con.alloc(d_con.n);
    if (d_base_val.n != d_nonzero_val.n) d_base_val.alloc(d_nonzero_val.n);
    IFEM_CUDA(cudaMemcpyAsync(d_base_con.p, d_con.p, d_con.n, cudaMemcpyDeviceToDevice, s));
    IFEM_CUDA(cudaMemcpyAsync(d_base_val.p, d_nonzero_val.p, d_nonzero_val.n * sizeof(double), cudaMemcpyDeviceToDevice, s));
    base_valid = true;
    flags_merged = false;
    schur_valid = false;
    IFEM_CUDA(cudaStreamSynchronize(s));
  }

  void FluidSpace::restore_base_constraints(Context &ctx)
  {
    cudaStream_t s = ctx.stream;
    IFEM_CUDA(cudaMemcpyAsync(d_con.p, d_base_con.p, d_con.n, cudaMemcpyDeviceToDevice, s));
    IFEM_CUDA(cudaMemcpyAsync(d_nonzero_val.p, d_base_val.p, d_nonzero_val.n * sizeof(double), cudaMemcpyDeviceToDevice, s));
    if (flags_merged) schur_valid = false; // S_m was formed for flags that held merged lines
    flags_merged = false;
  }

  void FluidSpace::set_neumann_faces(Context &ctx, const Triangulation &tria, const std::map<unsigned int, double> &neumann)
  {
    std::vector<int> fc;
    std::vector<double> fv;
    if (neumann.empty())
      {
        n_nfaces = 0;
        return;
      }
    std::vector<int> g2l(tria.n_cells(), -1);
    for (int l = 0; l < n_cells; ++l) g2l[local_cells[l]] = l;
    for (int f = 0; f < tria.n_boundary_faces(); ++f)
      {
        auto it = neumann.find((unsigned)tria.boundary_faces[3 * f + 2]);
        if (it == neumann.end()) continue;
        const int lc = g2l[tria.boundary_faces[3 * f]];
        if (lc < 0) continue; // not a local cell of this rank
        fc.push_back(lc);
        fc.push_back(tria.boundary_faces[3 * f + 1]);
        fv.push_back(it->second);
      }
    n_nfaces = (int)fv.size();
    if (!n_nfaces) return;
    // face quadrature QGauss<dim-1>(pu+1) placed on each of the 2*dim reference faces
    Quadrature fq(dim - 1, pu + 1);
    nqf = fq.nq;
    std::vector<double> t;
    std::vector<double> Nf((size_t)2 * dim * nqf * nu), Gf((size_t)2 * dim * nqf * nv * dim);
    std::vector<double> N(nu), dN((size_t)nu * dim), g(nv), dg((size_t)nv * dim);
    for (int face = 0; face < 2 * dim; ++face)
      for (int q = 0; q < nqf; ++q)
        {
          double xi[3];
          int k = 0;
          for (int d = 0; d < dim; ++d) xi[d] = (d == face / 2) ? double(face % 2) : fq.points[(size_t)q * (dim - 1) + k++];
          fe_u.eval(xi, N.data(), dN.data());
          fe_geo.eval(xi, g.data(), dg.data());
          std::copy(N.begin(), N.end(), Nf.begin() + ((size_t)face * nqf + q) * nu);
          std::copy(dg.begin(), dg.end(), Gf.begin() + ((size_t)face * nqf + q) * nv * dim);
        }
    t.insert(t.end(), Nf.begin(), Nf.end());
    t.insert(t.end(), Gf.begin(), Gf.end());
    t.insert(t.end(), fq.weights.begin(), fq.weights.end());
    cudaStream_t s = ctx.stream;
    d_face_tables.upload(t, s);
    d_nface_cell.upload(fc, s);
    d_nface_val.upload(fv, s);
    IFEM_CUDA(cudaStreamSynchronize(s));
  }

  // ===========================================================================
  // INS cell assembly kernel: one cell per team of warps (3-D: three warps, 2-D: one), cells of one colour per launch (no
  // two cells of a colour share a velocity node, so the scatter is a plain read-modify-write and the result is bitwise
  // reproducible). Shape values / pressure shape values / geometry gradients / weights are staged in shared memory by one
  // TMA bulk copy per CTA; the per-cell scratch (physical gradients, fields at the quadrature points) is shared by the team.
  // ===========================================================================
  namespace
  {
    template <int DIM>
    struct InsT
    {
      static constexpr int NU = DIM == 2 ? 9 : 27;
      static constexpr int NP = DIM == 2 ? 4 : 8;
      static constexpr int NQ = NU;
      static constexpr int NV = 1 << DIM;
      static constexpr int DPC = NU * DIM + NP;
      // a TEAM of warps works on one cell (the cell's scratch in shared memory is shared by the team), CELLS cells per CTA
      static constexpr int TEAM = DIM == 2 ? 1 : 3;
      static constexpr int CELLS = DIM == 2 ? 8 : 5;
      static constexpr int THREADS = TEAM * CELLS * 32;
      static constexpr int MIN_CTAS = 1;
      static constexpr int TAB = NQ * NU + NQ * NU * DIM + NQ * NP + NQ * NV * DIM + NQ; // doubles, layout of FluidSpace::d_tables
      static constexpr int STAB = (NQ * NU + NQ * NP + NQ * NV * DIM + NQ + 1) & ~1;     // staged in shared memory (d_tables_s)
    };

    template <int DIM>
    struct CellScratch
    {
      using T = InsT<DIM>;
      double g[T::NQ][T::NU][DIM]; // physical gradients of the velocity shape functions
      double Jinv[T::NQ][DIM * DIM];
      double JxW[T::NQ];
      double u[T::NQ][DIM], G[T::NQ][DIM * DIM], p[T::NQ], du[T::NQ][DIM], acc[T::NQ][DIM];
      double Ue[T::NU][DIM], Up[T::NU][DIM], Ua[T::NU][DIM], Pe[T::NP];
      double lrhs[T::DPC], ldiag[T::DPC], inh[T::DPC];
      int con[T::DPC];
      int un[T::NU], pn[T::NP];
    };

    struct BcsrView
    {
      const int64_t *rowptr;
      double *val;
    };

    struct InsArgs
    {
      int n_list;
      const int *cell_list;
      const int *cell_un, *cell_pn;
      const double *cell_x;
      const double *tables;   // FluidSpace::d_tables (global: the reference gradients dN are read from here)
      const double *tables_s; // FluidSpace::d_tables_s (staged in shared memory)
      const unsigned char *slots;
      const double *eval_pt, *present, *fsi_acc;
      const int *indicator;
      int64_t n_u;
      int n_owned_u, n_owned_p; // rows of nodes >= these are ghosts: assembled by their owner
      int lim_up, lim_diag;      // row limits of A_up / diag(M_u): owned, or owned + layer-1 ghosts in the Schur pass
      int do_uu, do_rhs;         // Schur pass: only the (solution independent) coupling blocks and diag(M_u)
      double mu, gamma, rho, inv_dt, grav[3];
      const unsigned char *con;
      const double *inhom; // null: homogeneous (zero_constraints)
      BcsrView uu, up, pu, mp;
      double *diag_Mu, *rhs;
      int assemble_mass;
      // InsIMEX (mpi_insimex.cpp:150-355): convection stays out of the matrix; rhs_only = the `assemble_system == false`
      // pass (no matrix entry is touched, constrained rows receive nothing)
      int explicit_convection, rhs_only;
    };

    template <int DIM>
    __device__ __forceinline__ void invert(const double *J, double *Ji, double &det)
    {
      if (DIM == 2)
        {
          det = J[0] * J[3] - J[1] * J[2];
          const double d = 1.0 / det;
          Ji[0] = J[3] * d; Ji[1] = -J[1] * d; Ji[2] = -J[2] * d; Ji[3] = J[0] * d;
        }
      else
        {
          const double c00 = J[4] * J[8] - J[5] * J[7], c01 = J[5] * J[6] - J[3] * J[8], c02 = J[3] * J[7] - J[4] * J[6];
          det = J[0] * c00 + J[1] * c01 + J[2] * c02;
          const double d = 1.0 / det;
          Ji[0] = c00 * d; Ji[1] = (J[2] * J[7] - J[1] * J[8]) * d; Ji[2] = (J[1] * J[5] - J[2] * J[4]) * d;
          Ji[3] = c01 * d; Ji[4] = (J[0] * J[8] - J[2] * J[6]) * d; Ji[5] = (J[2] * J[3] - J[0] * J[5]) * d;
          Ji[6] = c02 * d; Ji[7] = (J[1] * J[6] - J[0] * J[7]) * d; Ji[8] = (J[0] * J[4] - J[1] * J[3]) * d;
        }
    }

    // The cells of a CTA move through the phases in lock step (every team reaches every barrier, also a team without a cell
    // in the last round), so one CTA-wide barrier separates the phases; a single-warp team only needs the warp barrier.
    template <int TEAM>
    __device__ __forceinline__ void team_sync()
    {
      if (TEAM == 1)
        __syncwarp();
      else
        __syncthreads();
    }

    // One cell per team of warps. Phases (tl = thread of the team, TT threads):
    //   0  node ids, local dof values, constraints (items = local dofs) | geometry at the quadrature points (items = q)
    //   2  physical gradients g[q][b][k]                                  (items = (q, b))
    //   3  u, u - u_present, fsi acceleration, grad u, p at q             (items = (q, component))
    //   4  local right-hand side, diag(M_u)                               (items = (node, component))
    //   5  velocity-velocity blocks: lane = column node b, a warp takes three row nodes per pass; K accumulates over q in
    //      registers and goes straight to the BCSR planes of its row (read-modify-write: cells of a colour share no node)
    //   6  velocity-pressure coupling and M_p                             (items = (node, pressure node))
    //   8  local rhs through the constraints
    template <int DIM>
    __global__ void __launch_bounds__(InsT<DIM>::THREADS, InsT<DIM>::MIN_CTAS) ins_assemble_kernel(const InsArgs a)
    {
      using T = InsT<DIM>;
      constexpr int NU = T::NU, NP = T::NP, NQ = T::NQ, NV = T::NV, TEAM = T::TEAM, TT = T::TEAM * 32;
      extern __shared__ double smem[];
      __shared__ unsigned long long tma_bar;
      const double *tN = smem;                  // [NQ][NU]
      const double *tNp = tN + NQ * NU;         // [NQ][NP]
      const double *tdG = tNp + NQ * NP;        // [NQ][NV][DIM]
      const double *tqw = tdG + NQ * NV * DIM;  // [NQ]
      const double *tdN = a.tables + NQ * NU;   // [NQ][NU][DIM], global (phase 2 only)
      tma_stage_1d(smem, a.tables_s, T::STAB * (unsigned int)sizeof(double), &tma_bar);
      const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
      const int team = warp / TEAM, tw = warp % TEAM, tl = tw * 32 + lane;
      CellScratch<DIM> &S = *reinterpret_cast<CellScratch<DIM> *>(smem + T::STAB + (size_t)team * ((sizeof(CellScratch<DIM>) + 7) / 8));
      const double mu = a.mu, rho = a.rho, gam_rho = a.gamma * a.rho, rho_dt = a.rho * a.inv_dt;
      const double rho_conv = a.explicit_convection ? 0.0 : a.rho; // matrix side of the convection terms
      constexpr int SPC = NU * NU + 2 * NU * NP + NP * NP;

      for (int base_li = blockIdx.x * T::CELLS; base_li < a.n_list; base_li += gridDim.x * T::CELLS)
        {
          const int li = base_li + team;
          const bool active = li < a.n_list;
          const int cell = active ? a.cell_list[li] : 0;
          const int ind = (active && a.indicator) ? a.indicator[cell] : 0;
          const unsigned char *slots = a.slots + (int64_t)cell * SPC;
          // ---- phase 0: node ids | geometry at the quadrature points ----
          if (active)
            {
              if (tl < NU) S.un[tl] = a.cell_un[(int64_t)cell * NU + tl];
              if (tl < NP) S.pn[tl] = a.cell_pn[(int64_t)cell * NP + tl];
              // the last warp of the team takes the geometry (one lane per quadrature point)
              if (tw == TEAM - 1 && lane < NQ)
                {
                  const int q = lane;
                  const double *X = a.cell_x + (int64_t)cell * NV * DIM;
                  double J[DIM * DIM];
#pragma unroll
                  for (int i = 0; i < DIM * DIM; ++i) J[i] = 0.0;
                  for (int v = 0; v < NV; ++v)
#pragma unroll
                    for (int i = 0; i < DIM; ++i)
#pragma unroll
                      for (int j = 0; j < DIM; ++j) J[i * DIM + j] = fma(X[v * DIM + i], tdG[(q * NV + v) * DIM + j], J[i * DIM + j]);
                  double Ji[DIM * DIM], det;
                  invert<DIM>(J, Ji, det);
#pragma unroll
                  for (int i = 0; i < DIM * DIM; ++i) S.Jinv[q][i] = Ji[i];
                  S.JxW[q] = det * tqw[q];
                }
            }
          team_sync<TEAM>();
          // ---- phase 1: local dof values and constraints | phase 2: physical gradients ----
          if (active)
            {
              for (int i = tl; i < NU * DIM; i += TT)
                {
                  const int b = i / DIM, c = i % DIM;
                  const int64_t g = (int64_t)DIM * S.un[b] + c;
                  S.Ue[b][c] = a.eval_pt[g];
                  S.Up[b][c] = a.present[g];
                  S.Ua[b][c] = (ind && a.fsi_acc) ? a.fsi_acc[g] : 0.0;
                  const int cf = a.con[g];
                  S.con[i] = cf;
                  S.inh[i] = (cf && a.inhom) ? a.inhom[g] : 0.0;
                  S.lrhs[i] = 0.0;
                  S.ldiag[i] = 0.0;
                }
              if (tl < NP)
                {
                  const int64_t g = a.n_u + S.pn[tl];
                  S.Pe[tl] = a.eval_pt[g];
                  const int cf = a.con[g];
                  S.con[NU * DIM + tl] = cf;
                  S.inh[NU * DIM + tl] = (cf && a.inhom) ? a.inhom[g] : 0.0;
                  S.lrhs[NU * DIM + tl] = 0.0;
                  S.ldiag[NU * DIM + tl] = 0.0;
                }
              // g[q][b][k] = sum_j dN[q][b][j] Jinv[q][j][k]
              for (int i = tl; i < NQ * NU; i += TT)
                {
                  const int q = i / NU;
                  double r[DIM];
#pragma unroll
                  for (int j = 0; j < DIM; ++j) r[j] = __ldg(tdN + (size_t)i * DIM + j);
#pragma unroll
                  for (int k = 0; k < DIM; ++k)
                    {
                      double s = 0.0;
#pragma unroll
                      for (int j = 0; j < DIM; ++j) s = fma(r[j], S.Jinv[q][j * DIM + k], s);
                      (&S.g[0][0][0])[(size_t)i * DIM + k] = s;
                    }
                }
            }
          team_sync<TEAM>();
          // ---- phase 3: field values at q: get_function_values / gradients (:219-232), one item per (q, component) ----
          if (active)
            {
              for (int i = tl; i < NQ * DIM; i += TT)
                {
                  const int q = i / DIM, c = i % DIM;
                  double u = 0.0, up = 0.0, ac = 0.0, G[DIM];
#pragma unroll
                  for (int k = 0; k < DIM; ++k) G[k] = 0.0;
                  for (int b = 0; b < NU; ++b)
                    {
                      const double N = tN[q * NU + b];
                      const double ue = S.Ue[b][c];
                      u = fma(N, ue, u);
                      up = fma(N, S.Up[b][c], up);
                      ac = fma(N, S.Ua[b][c], ac);
#pragma unroll
                      for (int k = 0; k < DIM; ++k) G[k] = fma(ue, S.g[q][b][k], G[k]);
                    }
                  S.u[q][c] = u;
                  S.du[q][c] = u - up;
                  S.acc[q][c] = ac;
#pragma unroll
                  for (int k = 0; k < DIM; ++k) S.G[q][c * DIM + k] = G[k];
                }
              // pressure at q (the items above leave the tail of the last pass idle: give it to the threads from the end)
              for (int i = TT - 1 - tl; i < NQ; i += TT)
                {
                  double p = 0.0;
                  for (int j = 0; j < NP; ++j) p = fma(tNp[i * NP + j], S.Pe[j], p);
                  S.p[i] = p;
                }
            }
          team_sync<TEAM>();
          // ---- phase 4: local rhs (:281-304) and diag(M_u), one item per (node, component); pressure rows ----
          if (active)
            {
              for (int i = tl; i < NU * DIM; i += TT)
                {
                  const int aN = i / DIM, c = i % DIM;
                  double r = 0.0, m = 0.0;
                  for (int q = 0; q < NQ; ++q)
                    {
                      const double w = S.JxW[q], N = tN[q * NU + aN];
                      m = fma(w * N, N, m);
                      double div = 0.0;
#pragma unroll
                      for (int k = 0; k < DIM; ++k) div += S.G[q][k * DIM + k];
                      const double pd = S.p[q] - gam_rho * div;
                      double t = 0.0, conv = 0.0;
#pragma unroll
                      for (int k = 0; k < DIM; ++k)
                        {
                          t = fma(S.G[q][c * DIM + k], S.g[q][aN][k], t);    // grad u : grad phi_i
                          conv = fma(S.G[q][c * DIM + k], S.u[q][k], conv);  // (grad u) u
                        }
                      double v = -mu * t + pd * S.g[q][aN][c] + N * (-rho * conv - rho_dt * S.du[q][c] + rho * a.grav[c]);
                      if (ind == 1) v += rho * S.acc[q][c] * N;
                      r = fma(w, v, r);
                    }
                  S.lrhs[i] = r;
                  if (a.assemble_mass && S.un[aN] < a.lim_diag) a.diag_Mu[(int64_t)DIM * S.un[aN] + c] += m;
                }
              for (int i = TT - 1 - tl; i < NP; i += TT)
                {
                  double r = 0.0;
                  for (int q = 0; q < NQ; ++q)
                    {
                      double div = 0.0;
#pragma unroll
                      for (int k = 0; k < DIM; ++k) div += S.G[q][k * DIM + k];
                      r = fma(S.JxW[q] * div, tNp[q * NP + i], r);
                    }
                  S.lrhs[NU * DIM + i] = r;
                }
            }
          team_sync<TEAM>();
          // ---- phase 5: velocity-velocity blocks (:263-273), lane = column node b, 3 row nodes per pass and warp ----
          //   K_ab[c][d] = sum_q  delta_cd (w mu ga.gb + Na (w rho u.gb + w rho/dt Nb)) + (Na Nb) (w rho G[c][d]) + ga[c] (w gamma rho gb[d])
          if (active && a.do_uu && !a.rhs_only)
            {
              const int b = lane < NU ? lane : NU - 1;
              int cb[DIM];
              double ib[DIM];
#pragma unroll
              for (int d = 0; d < DIM; ++d)
                {
                  cb[d] = S.con[b * DIM + d];
                  ib[d] = S.inh[b * DIM + d];
                }
              constexpr int TA = 3;
              for (int a0 = tw * TA; a0 < NU; a0 += TEAM * TA)
                {
                  double K[TA][DIM * DIM], sK[TA];
                  // the planes this pass will read-modify-write: on their way to L2 while K is being summed
                  if (lane < NU)
                    {
#pragma unroll
                      for (int t = 0; t < TA; ++t)
                        {
                          const int A = S.un[a0 + t];
                          if (A >= a.n_owned_u) continue;
                          const int64_t rp = a.uu.rowptr[A];
                          const int nb = (int)(a.uu.rowptr[A + 1] - rp);
                          const double *base = a.uu.val + rp * (DIM * DIM) + slots[(a0 + t) * NU + b];
#pragma unroll
                          for (int i = 0; i < DIM * DIM; ++i) prefetch_l2(base + (int64_t)i * nb);
                        }
                    }
#pragma unroll
                  for (int t = 0; t < TA; ++t)
                    {
                      sK[t] = 0.0;
#pragma unroll
                      for (int i = 0; i < DIM * DIM; ++i) K[t][i] = 0.0;
                    }
                  for (int q = 0; q < NQ; ++q)
                    {
                      const double w = S.JxW[q];
                      const double Nb = tN[q * NU + b];
                      double gb[DIM], wmg[DIM], wgg[DIM], wG[DIM * DIM];
                      double ugb = 0.0;
#pragma unroll
                      for (int k = 0; k < DIM; ++k)
                        {
                          gb[k] = S.g[q][b][k];
                          ugb = fma(S.u[q][k], gb[k], ugb);
                        }
                      const double wm = w * mu, wg = w * gam_rho, wr = w * rho_conv;
#pragma unroll
                      for (int k = 0; k < DIM; ++k)
                        {
                          wmg[k] = wm * gb[k];
                          wgg[k] = wg * gb[k];
                        }
#pragma unroll
                      for (int i = 0; i < DIM * DIM; ++i) wG[i] = wr * S.G[q][i];
                      const double c12 = fma(wr, ugb, w * rho_dt * Nb);
#pragma unroll
                      for (int t = 0; t < TA; ++t)
                        {
                          const int aN = a0 + t;
                          const double Na = tN[q * NU + aN];
                          double ga[DIM];
#pragma unroll
                          for (int k = 0; k < DIM; ++k) ga[k] = S.g[q][aN][k];
                          double s = fma(Na, c12, sK[t]);
#pragma unroll
                          for (int k = 0; k < DIM; ++k) s = fma(ga[k], wmg[k], s);
                          sK[t] = s;
                          const double NaNb = Na * Nb;
#pragma unroll
                          for (int c = 0; c < DIM; ++c)
#pragma unroll
                            for (int d = 0; d < DIM; ++d)
                              K[t][c * DIM + d] = fma(ga[c], wgg[d], fma(NaNb, wG[c * DIM + d], K[t][c * DIM + d]));
                        }
                    }
                  // scatter the TA x (DIM x DIM) blocks of this lane
#pragma unroll
                  for (int t = 0; t < TA; ++t)
                    {
                      const int aN = a0 + t;
                      const int A = S.un[aN];
                      if (A >= a.n_owned_u) continue; // warp-uniform
#pragma unroll
                      for (int c = 0; c < DIM; ++c) K[t][c * DIM + c] += sK[t];
                      const int64_t rp = a.uu.rowptr[A];
                      const int nb = (int)(a.uu.rowptr[A + 1] - rp);
                      double *base = a.uu.val + rp * (DIM * DIM) + slots[aN * NU + b];
                      double corr[DIM];
#pragma unroll
                      for (int c = 0; c < DIM; ++c) corr[c] = 0.0;
                      if (lane < NU)
                        {
                          // what goes into plane (c, d) of this block: the entry, |diagonal| of a constrained row, or nothing
                          // (constrained row off the diagonal; constrained column: lifted to the right-hand side)
                          double add[DIM * DIM];
                          bool put[DIM * DIM];
#pragma unroll
                          for (int c = 0; c < DIM; ++c)
                            {
                              const int rc = S.con[aN * DIM + c];
#pragma unroll
                              for (int d = 0; d < DIM; ++d)
                                {
                                  const double v = K[t][c * DIM + d];
                                  add[c * DIM + d] = v;
                                  put[c * DIM + d] = false;
                                  if (rc)
                                    {
                                      if (b == aN && c == d)
                                        {
                                          add[c * DIM + d] = fabs(v);
                                          put[c * DIM + d] = true;
                                          S.ldiag[aN * DIM + c] = fabs(v);
                                        }
                                    }
                                  else if (cb[d])
                                    corr[c] = fma(v, ib[d], corr[c]);
                                  else
                                    put[c * DIM + d] = true;
                                }
                            }
                          // read-modify-write of the DIM x DIM planes: all loads of the block in flight before the first store
                          // (the planes of a block never alias; written as one chain the compiler has to serialise them)
                          double old[DIM * DIM];
#pragma unroll
                          for (int i = 0; i < DIM * DIM; ++i) old[i] = put[i] ? __ldcg(base + (int64_t)i * nb) : 0.0;
#pragma unroll
                          for (int i = 0; i < DIM * DIM; ++i)
                            if (put[i]) __stcg(base + (int64_t)i * nb, old[i] + add[i]);
                        }
                      if (a.inhom)
                        {
#pragma unroll
                          for (int c = 0; c < DIM; ++c)
                            {
                              const double sc = warp_sum(corr[c]);
                              if (lane == 0) S.lrhs[aN * DIM + c] -= sc;
                            }
                        }
                    }
                }
            }
          team_sync<TEAM>();
          // ---- phase 6: velocity-pressure coupling  -div(phi_i) psi_j  and its transpose, one item per (a, j);
          //      phase 7: pressure mass matrix (:274-276), one item per (i, j) ----
          if (active && !a.rhs_only)
            {
              const unsigned char *s_up = slots + NU * NU;
              const unsigned char *s_pu = slots + NU * NU + NU * NP;
              for (int i = tl; i < NU * NP; i += TT)
                {
                  const int aN = i / NP, j = i % NP;
                  double B[DIM];
#pragma unroll
                  for (int c = 0; c < DIM; ++c) B[c] = 0.0;
                  for (int q = 0; q < NQ; ++q)
                    {
                      const double wpsi = -S.JxW[q] * tNp[q * NP + j];
#pragma unroll
                      for (int c = 0; c < DIM; ++c) B[c] = fma(wpsi, S.g[q][aN][c], B[c]);
                    }
                  const int A = S.un[aN];
                  const bool own_row = A < a.lim_up;
                  const int64_t rp = own_row ? a.up.rowptr[A] : 0;
                  const int nb = own_row ? (int)(a.up.rowptr[A + 1] - rp) : 0;
                  double *base = a.up.val + rp * DIM;
                  const int slot = s_up[aN * NP + j];
                  const int pc = S.con[NU * DIM + j];
                  // row = pressure node j, column = velocity node a
                  const int Pn = S.pn[j];
                  const bool own_p = Pn < a.n_owned_p;
                  const int64_t rq = own_p ? a.pu.rowptr[Pn] : 0;
                  const int nbq = own_p ? (int)(a.pu.rowptr[Pn + 1] - rq) : 0;
                  double *baseq = a.pu.val + rq * DIM;
                  const int slotq = s_pu[j * NU + aN];
                  // both read-modify-writes of every component in flight before the first store (see phase 5)
                  bool put[DIM];
                  double old_up[DIM], old_pu[DIM];
#pragma unroll
                  for (int c = 0; c < DIM; ++c)
                    {
                      put[c] = !S.con[aN * DIM + c] && !pc;
                      old_up[c] = (put[c] && own_row) ? __ldcg(base + (int64_t)c * nb + slot) : 0.0;
                      old_pu[c] = (put[c] && own_p) ? __ldcg(baseq + (int64_t)c * nbq + slotq) : 0.0;
                    }
#pragma unroll
                  for (int c = 0; c < DIM; ++c)
                    {
                      const int uc = S.con[aN * DIM + c];
                      const double v = B[c];
                      if (put[c])
                        {
                          if (own_row) __stcg(base + (int64_t)c * nb + slot, old_up[c] + v);   // A_up block row a, plane c
                          if (own_p) __stcg(baseq + (int64_t)c * nbq + slotq, old_pu[c] + v);  // A_pu row j, plane c
                        }
                      else if (uc && !pc && a.inhom)
                        atomicAdd(&S.lrhs[NU * DIM + j], -v * S.inh[aN * DIM + c]); // column (a,c) constrained
                    }
                }
              if (a.assemble_mass && a.do_rhs)
                for (int e = TT - 1 - tl; e < NP * NP; e += TT)
                  {
                    const int i = e / NP, j = e % NP;
                    double m = 0.0;
                    for (int q = 0; q < NQ; ++q) m = fma(S.JxW[q] * tNp[q * NP + i], tNp[q * NP + j], m);
                    const int Pn = S.pn[i];
                    if (Pn >= a.n_owned_p) continue;
                    const int64_t rp = a.mp.rowptr[Pn];
                    a.mp.val[rp + slots[NU * NU + 2 * NU * NP + e]] += m;
                  }
            }
          team_sync<TEAM>();
          // ---- scatter local rhs through the constraints (distribute_local_to_global) ----
          if (active && a.do_rhs)
            for (int i = tl; i < T::DPC; i += TT)
              {
                const bool own = i < NU * DIM ? S.un[i / DIM] < a.n_owned_u : S.pn[i - NU * DIM] < a.n_owned_p;
                if (!own) continue;
                const int64_t g = i < NU * DIM ? (int64_t)DIM * S.un[i / DIM] + i % DIM : a.n_u + S.pn[i - NU * DIM];
                if (!S.con[i])
                  a.rhs[g] += S.lrhs[i];
                else if (a.inhom)
                  a.rhs[g] += S.ldiag[i] * S.inh[i];
              }
          team_sync<TEAM>();
        }
    }

    // Pressure Neumann faces (:313-341): rhs_i -= phi_i . n  p  JxW_face on unconstrained rows.
    template <int DIM, int NU>
    __global__ void ins_neumann_kernel(int n_faces, int nqf, const int *__restrict__ face_cell, const double *__restrict__ face_val,
                                       const double *__restrict__ ftab, const int *__restrict__ cell_un,
                                       const double *__restrict__ cell_x, const unsigned char *__restrict__ con, int n_owned_u,
                                       double *__restrict__ rhs)
    {
      constexpr int NV = 1 << DIM;
      const int f = blockIdx.x;
      if (f >= n_faces) return;
      const int cell = face_cell[2 * f], face = face_cell[2 * f + 1];
      const int axis = face / 2, side = face % 2;
      const double pbar = face_val[f];
      const double *Nf = ftab;
      const double *Gf = ftab + (size_t)2 * DIM * nqf * NU;
      const double *qwf = Gf + (size_t)2 * DIM * nqf * NV * DIM;
      const double *X = cell_x + (int64_t)cell * NV * DIM;
      for (int i = threadIdx.x; i < NU * DIM; i += blockDim.x)
        {
          const int node = i / DIM, c = i % DIM;
          double r = 0.0;
          for (int q = 0; q < nqf; ++q)
            {
              const size_t fq = (size_t)face * nqf + q;
              double J[DIM * DIM], Ji[DIM * DIM], det;
              for (int k = 0; k < DIM * DIM; ++k) J[k] = 0.0;
              for (int v = 0; v < NV; ++v)
                for (int ii = 0; ii < DIM; ++ii)
                  for (int jj = 0; jj < DIM; ++jj) J[ii * DIM + jj] = fma(X[v * DIM + ii], Gf[(fq * NV + v) * DIM + jj], J[ii * DIM + jj]);
              invert<DIM>(J, Ji, det);
              // n dS = det(J) J^{-T} n_ref
              const double nds = det * Ji[axis * DIM + c] * (side ? 1.0 : -1.0) * qwf[q];
              r -= Nf[fq * NU + node] * nds * pbar;
            }
          const int gn = cell_un[(int64_t)cell * NU + node];
          const int64_t g = (int64_t)DIM * gn + c;
          if (gn < n_owned_u && !con[g] && r != 0.0) atomicAdd(&rhs[g], r);
        }
    }
  } // namespace

  void neumann_faces(Context &ctx, FluidSpace &fs)
  {
    if (!fs.n_nfaces) return;
    auto go = [&](auto dim_tag, auto nu_tag) {
      constexpr int DIM = decltype(dim_tag)::value, NU = decltype(nu_tag)::value;
      ins_neumann_kernel<DIM, NU><<<fs.n_nfaces, 64, 0, ctx.stream>>>(fs.n_nfaces, fs.nqf, fs.d_nface_cell.p, fs.d_nface_val.p,
                                                                      fs.d_face_tables.p, fs.d_cell_un.p, fs.d_cell_x.p, fs.d_con.p,
                                                                      fs.n_owned_unodes, fs.rhs.p);
    };
    using I2 = std::integral_constant<int, 2>;
    using I3 = std::integral_constant<int, 3>;
    if (fs.dim == 2 && fs.nu == 9) go(I2(), std::integral_constant<int, 9>());
    else if (fs.dim == 2 && fs.nu == 4) go(I2(), std::integral_constant<int, 4>());
    else if (fs.dim == 3 && fs.nu == 27) go(I3(), std::integral_constant<int, 27>());
    else if (fs.dim == 3 && fs.nu == 8) go(I3(), std::integral_constant<int, 8>());
    else throw std::runtime_error("neumann_faces: unsupported element");
    IFEM_KERNEL_CHECK();
    ctx.kernel_launches++;
  }

  template <int DIM>
  static void ins_assemble_dim(Context &ctx, FluidSpace &fs, const InsAssembleParams &prm, const double *eval_pt,
                               const double *present, const double *fsi_acc, bool use_nonzero, bool assemble_mass, bool schur_pass)
  {
    using T = InsT<DIM>;
    if (fs.nu != T::NU || fs.np != T::NP) throw std::runtime_error("ins_assemble: only Q2/Q1 elements are supported");
    cudaStream_t s = ctx.stream;
    if (prm.rhs_only)
      {
        fs.rhs.zero(s);
        assemble_mass = false;
      }
    else if (schur_pass)
      {
        // only the solution-independent coupling blocks and diag(M_u), on the rows of owned + layer-1 nodes
        fs.A_up.zero(s);
        fs.A_pu.zero(s);
        fs.diag_Mu.zero(s);
        assemble_mass = true;
      }
    else
      {
        fs.A_uu.zero(s);
        fs.A_up.zero(s);
        fs.A_pu.zero(s);
        fs.rhs.zero(s);
        if (assemble_mass)
          {
            fs.M_p.zero(s);
            fs.diag_Mu.zero(s);
          }
      }
    InsArgs a;
    a.cell_un = fs.d_cell_un.p;
    a.cell_pn = fs.d_cell_pn.p;
    a.cell_x = fs.d_cell_x.p;
    a.tables = fs.d_tables.p;
    a.tables_s = fs.d_tables_s.p;
    a.slots = fs.d_slots.p;
    a.eval_pt = eval_pt;
    a.present = present;
    a.fsi_acc = fsi_acc;
    a.indicator = fs.d_indicator.p;
    a.n_u = fs.n_u;
    a.n_owned_u = fs.n_owned_unodes;
    a.n_owned_p = fs.n_owned_pnodes;
    a.lim_up = schur_pass ? fs.n_layer1_unodes : fs.n_owned_unodes;
    a.lim_diag = a.lim_up;
    a.do_uu = schur_pass ? 0 : 1;
    a.do_rhs = schur_pass ? 0 : 1;
    a.mu = prm.viscosity;
    a.gamma = prm.gamma;
    a.rho = prm.rho;
    a.inv_dt = 1.0 / prm.dt;
    for (int d = 0; d < 3; ++d) a.grav[d] = prm.gravity[d];
    a.con = fs.d_con.p;
    a.inhom = use_nonzero && !prm.rhs_only ? fs.d_nonzero_val.p : nullptr;
    a.explicit_convection = prm.explicit_convection;
    a.rhs_only = prm.rhs_only;
    a.uu = {fs.A_uu.rowptr.p, fs.A_uu.val.p};
    a.up = {fs.A_up.rowptr.p, fs.A_up.val.p};
    a.pu = {fs.A_pu.rowptr.p, fs.A_pu.val.p};
    a.mp = {fs.M_p.rowptr.p, fs.M_p.val.p};
    a.diag_Mu = fs.diag_Mu.p;
    a.rhs = fs.rhs.p;
    a.assemble_mass = assemble_mass ? 1 : 0;
    const size_t smem = (size_t)T::STAB * 8 + (size_t)T::CELLS * (((sizeof(CellScratch<DIM>) + 7) / 8) * 8);
    static bool attr_set[4] = {false, false, false, false};
    if (!attr_set[DIM])
      {
        IFEM_CUDA(cudaFuncSetAttribute(ins_assemble_kernel<DIM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set[DIM] = true;
      }
    const int n_colours = (int)fs.colour_offsets.size() - 1;
    for (int k = 0; k < n_colours; ++k)
      {
        // layer-1 cells come first inside every colour; the Schur pass also visits the layer-2 cells
        a.n_list = schur_pass ? fs.colour_offsets[k + 1] - fs.colour_offsets[k] : fs.colour_n1[k];
        a.cell_list = fs.d_colour_order.p + fs.colour_offsets[k];
        if (a.n_list == 0) continue;
        // persistent CTAs: as many as fit on the device at once (shared memory bound), each loops over its share of the colour
        const int blocks = std::min((a.n_list + T::CELLS - 1) / T::CELLS, ctx.sm_count * (DIM == 2 ? 4 : T::MIN_CTAS));
        ins_assemble_kernel<DIM><<<blocks, T::THREADS, smem, s>>>(a);
        IFEM_KERNEL_CHECK();
        ctx.kernel_launches++;
      }
    if (!schur_pass) neumann_faces(ctx, fs);
  }

  void ins_assemble(Context &ctx, FluidSpace &fs, const InsAssembleParams &prm, const double *eval_pt, const double *present,
                    const double *fsi_acc, bool use_nonzero_constraints, bool assemble_mass, bool schur_pass)
  {
    if (fs.dim == 2)
      ins_assemble_dim<2>(ctx, fs, prm, eval_pt, present, fsi_acc, use_nonzero_constraints, assemble_mass, schur_pass);
    else
      ins_assemble_dim<3>(ctx, fs, prm, eval_pt, present, fsi_acc, use_nonzero_constraints, assemble_mass, schur_pass);
  }

  // ===========================================================================
  // update_stress: one thread per (cell, node a): tau = 2 mu sym grad v at the quadrature points, projected to node a
  // with qpt_to_dof, scatter-added (cells of one colour per launch share no node), then divided by the cell count.
  // ===========================================================================
  namespace
  {
    template <int DIM, int NU>
    __global__ void stress_kernel(int n_list, const int *__restrict__ cell_list, const int *__restrict__ cell_un,
                                  const double *__restrict__ cell_x, const double *__restrict__ tdN, const double *__restrict__ tdG,
                                  const double *__restrict__ qpt_to_dof, const double *__restrict__ present, double mu, int n_unodes,
                                  int n_owned_u, double *__restrict__ stress, double *__restrict__ count)
    {
      constexpr int NQ = NU, NV = 1 << DIM;
      const int t = blockIdx.x * blockDim.x + threadIdx.x;
      if (t >= n_list * NU) return;
      const int cell = cell_list[t / NU], a = t % NU;
      const int un_a = cell_un[(int64_t)cell * NU + a];
      if (un_a >= n_owned_u) return;
      const double *X = cell_x + (int64_t)cell * NV * DIM;
      double acc[DIM * DIM];
#pragma unroll
      for (int i = 0; i < DIM * DIM; ++i) acc[i] = 0.0;
      for (int q = 0; q < NQ; ++q)
        {
          double J[DIM * DIM], Ji[DIM * DIM], G[DIM * DIM];
#pragma unroll
          for (int i = 0; i < DIM * DIM; ++i) J[i] = G[i] = 0.0;
          for (int v = 0; v < NV; ++v)
#pragma unroll
            for (int i = 0; i < DIM; ++i)
#pragma unroll
              for (int j = 0; j < DIM; ++j) J[i * DIM + j] = fma(X[v * DIM + i], tdG[(q * NV + v) * DIM + j], J[i * DIM + j]);
          double det;
          invert<DIM>(J, Ji, det);
          for (int b = 0; b < NU; ++b)
            {
              const int un = cell_un[(int64_t)cell * NU + b];
              double g[DIM];
#pragma unroll
              for (int k = 0; k < DIM; ++k)
                {
                  double sacc = 0.0;
#pragma unroll
                  for (int j = 0; j < DIM; ++j) sacc = fma(tdN[(q * NU + b) * DIM + j], Ji[j * DIM + k], sacc);
                  g[k] = sacc;
                }
#pragma unroll
              for (int c = 0; c < DIM; ++c)
#pragma unroll
                for (int k = 0; k < DIM; ++k) G[c * DIM + k] = fma(present[(int64_t)DIM * un + c], g[k], G[c * DIM + k]);
            }
          const double w = qpt_to_dof[a * NQ + q];
#pragma unroll
          for (int i = 0; i < DIM; ++i)
#pragma unroll
            for (int j = 0; j < DIM; ++j) acc[i * DIM + j] = fma(w, mu * (G[i * DIM + j] + G[j * DIM + i]), acc[i * DIM + j]);
        }
#pragma unroll
      for (int ij = 0; ij < DIM * DIM; ++ij) stress[(int64_t)ij * n_unodes + un_a] += acc[ij];
      count[un_a] += 1.0;
    }

    __global__ void stress_average_kernel(int n, int ncomp, int n_unodes, const double *__restrict__ count, double *__restrict__ stress)
    {
      const int i = blockIdx.x * blockDim.x + threadIdx.x;
      if (i >= n) return;
      const double c = count[i];
      if (c > 0)
        for (int k = 0; k < ncomp; ++k) stress[(int64_t)k * n_unodes + i] /= c;
    }
  } // namespace

  void update_nodal_stress(Context &ctx, FluidSpace &fs, const double *present, double viscosity, double *stress)
  {
    cudaStream_t s = ctx.stream;
    const int dim = fs.dim, nn = fs.un.n_nodes;
    IFEM_CUDA(cudaMemsetAsync(stress, 0, (size_t)dim * dim * nn * sizeof(double), s));
    fs.d_stress_count.zero(s);
    const double *tdN = fs.d_tables.p + (size_t)fs.nq * fs.nu;
    const double *tdG = tdN + (size_t)fs.nq * fs.nu * dim + (size_t)fs.nq * fs.np;
    const int n_colours = (int)fs.colour_offsets.size() - 1;
    for (int k = 0; k < n_colours; ++k)
      {
        const int n = fs.colour_n1[k]; // every cell around an owned node is a layer-1 cell
        if (!n) continue;
        const int *list = fs.d_colour_order.p + fs.colour_offsets[k];
        auto go = [&](auto dim_tag, auto nu_tag) {
          constexpr int DIM = decltype(dim_tag)::value, NU = decltype(nu_tag)::value;
          const int total = n * NU;
          stress_kernel<DIM, NU><<<(total + 127) / 128, 128, 0, s>>>(n, list, fs.d_cell_un.p, fs.d_cell_x.p, tdN, tdG, fs.d_qpt_to_dof.p, present,
                                                                     viscosity, nn, fs.n_owned_unodes, stress, fs.d_stress_count.p);
        };
        using I2 = std::integral_constant<int, 2>;
        using I3 = std::integral_constant<int, 3>;
        if (dim == 2 && fs.nu == 9) go(I2(), std::integral_constant<int, 9>());
        else if (dim == 2 && fs.nu == 4) go(I2(), std::integral_constant<int, 4>());
        else if (dim == 3 && fs.nu == 27) go(I3(), std::integral_constant<int, 27>());
        else if (dim == 3 && fs.nu == 8) go(I3(), std::integral_constant<int, 8>());
        else throw std::runtime_error("update_nodal_stress: unsupported element");
        IFEM_KERNEL_CHECK();
        ctx.kernel_launches++;
      }
    stress_average_kernel<<<(nn + 255) / 256, 256, 0, s>>>(nn, dim * dim, nn, fs.d_stress_count.p, stress);
    IFEM_KERNEL_CHECK();
    ctx.kernel_launches++;
    if (fs.n_ranks > 1)
      {
        // relevant_partition_stress = stress (ghosted copy, mpi_scnsim.cpp:36-45): scalar halo on the velocity node set
        Halo &h = fs.halo_s;
        for (int k = 0; k < dim * dim; ++k) h.update(ctx, stress + (size_t)k * nn);
      }
  }

  void block_vmult(Context &ctx, const FluidSpace &fs, const double *x, double *y)
  {
    // ghost entries of x are scratch by design: refresh them from their owners first
    const_cast<FluidSpace &>(fs).halo_update(ctx, const_cast<double *>(x));
    spmv(ctx, fs.A_uu, x, y, false);
    spmv(ctx, fs.A_up, x + fs.n_u, y, true);
    spmv(ctx, fs.A_pu, x, y + fs.n_u, false);
    if (fs.A_pp.n_brows) spmv(ctx, fs.A_pp, x + fs.n_u, y + fs.n_u, true);
  }

  // ===========================================================================
  // S_m = B diag(M_u)^-1 B^T, B = A_pu, B^T = A_up (mpi_insim.cpp:44-49), numeric
  // phase on the fixed pattern: one warp per pressure row, accumulation in shared
  // memory, slots located by binary search in the (sorted) Schur row.
  // ===========================================================================
  namespace
  {
    template <int DIM>
    __global__ void __launch_bounds__(128) mass_schur_kernel(int n_p, const int64_t *__restrict__ pu_rp, const int *__restrict__ pu_col,
                                                             const double *__restrict__ pu_val, const int64_t *__restrict__ up_rp,
                                                             const int *__restrict__ up_col, const double *__restrict__ up_val,
                                                             const double *__restrict__ diag_Mu, const int64_t *__restrict__ s_rp,
                                                             const int *__restrict__ s_col, double *__restrict__ s_val)
    {
      __shared__ double acc[4][256];
      const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
      const int row = blockIdx.x * 4 + warp;
      if (row >= n_p) return;
      const int64_t sb = s_rp[row];
      const int sn = (int)(s_rp[row + 1] - sb);
      for (int i = lane; i < sn; i += 32) acc[warp][i] = 0.0;
      __syncwarp();
      const int64_t rb = pu_rp[row];
      const int rn = (int)(pu_rp[row + 1] - rb);
      // lanes walk the velocity nodes k of B's row; each visits the B^T rows of (k, d)
      for (int jk = lane; jk < rn; jk += 32)
        {
          const int k = pu_col[rb + jk];
          const int64_t ub = up_rp[k];
          const int un = (int)(up_rp[k + 1] - ub);
#pragma unroll
          for (int d = 0; d < DIM; ++d)
            {
              const double bkd = pu_val[rb * DIM + (int64_t)d * rn + jk] / diag_Mu[(int64_t)DIM * k + d];
              if (bkd == 0.0) continue;
              for (int m = 0; m < un; ++m)
                {
                  const int pj = up_col[ub + m];
                  const double v = bkd * up_val[ub * DIM + (int64_t)d * un + m];
                  int lo = 0, hi = sn - 1;
                  while (lo < hi)
                    {
                      const int mid = (lo + hi) >> 1;
                      if (s_col[sb + mid] < pj) lo = mid + 1; else hi = mid;
                    }
                  atomicAdd(&acc[warp][lo], v);
                }
            }
        }
      __syncwarp();
      for (int i = lane; i < sn; i += 32) s_val[sb + i] = acc[warp][i];
    }
  } // namespace

  void apply_mass_schur_matrix_free(Context &ctx, FluidSpace &fs, const double *x_p, double *y_p, double *tmp_u)
  {
    fs.halo_p.update(ctx, const_cast<double *>(x_p));
    spmv(ctx, fs.A_up, x_p, tmp_u);
    divide(ctx, fs.vs_u, fs.diag_Mu.p, tmp_u);
    fs.halo_u.update(ctx, tmp_u);
    spmv(ctx, fs.A_pu, tmp_u, y_p);
  }

  void compute_mass_schur(Context &ctx, FluidSpace &fs)
  {
    const int n_p = fs.n_owned_pnodes;
    const int blocks = (n_p + 3) / 4;
    auto go = [&](auto tag) {
      constexpr int DIM = decltype(tag)::value;
      mass_schur_kernel<DIM><<<blocks, 128, 0, ctx.stream>>>(n_p, fs.A_pu.rowptr.p, fs.A_pu.col.p, fs.A_pu.val.p, fs.A_up.rowptr.p,
                                                             fs.A_up.col.p, fs.A_up.val.p, fs.diag_Mu.p, fs.S_m.rowptr.p,
                                                             fs.S_m.col.p, fs.S_m.val.p);
    };
    if (fs.dim == 2) go(std::integral_constant<int, 2>()); else go(std::integral_constant<int, 3>());
    IFEM_KERNEL_CHECK();
    ctx.kernel_launches++;
  }
} // namespace ifem
