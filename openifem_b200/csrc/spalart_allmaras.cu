// Spalart-Allmaras turbulence model on the device: see spalart_allmaras.h for the reference map.
//
// Kernels:
//   sa_assemble_kernel<DIM>   the cell loop of SpalartAllmaras::assemble (source/mpi_spalart_allmaras.cpp:681-827): per colour one
//                             launch, one thread per (cell, test node a, trial node b); the quadrature-point quantities (model
//                             coefficients P, D, f_n, the diffusivity, physical gradients) are prepared once per (cell, q) in shared
//                             memory; Dirichlet lines are eliminated in the scatter like AffineConstraints::distribute_local_to_global
//   wall_distance_kernel      setup_cell_property (:497-551): nearest wall VERTEX per scalar support point, wall points tiled through
//                             shared memory
//   eddy_viscosity_kernel     update_eddy_viscosity (:864-889)
//   solid_lines_kernel        update_boundary_condition (:159-181): nodes of cells inside the immersed solid carry nu~ -> 0
// HBM-bound scalar work of n_nodes doubles; the transport system has the pattern of the pressure block.
#include "spalart_allmaras.h"

#include <cfloat>
#include <cmath>
#include <cstdio>
#include <set>

#include "scnsim.h"

namespace ifem
{
  namespace
  {
    // constants of source/mpi_spalart_allmaras.cpp:624-631
    constexpr double kCv1 = 7.1, kCv2 = 0.7, kCv3 = 0.9;
    constexpr double kCb1 = 0.1355, kCb2 = 0.622, kCt3 = 1.2, kCt4 = 0.5, kKappa = 0.41;
    constexpr double kCw2 = 0.3, kCw3 = 2.0, kCn1 = 16.0;
    constexpr double kSigma = 2.0 / 3.0;
    constexpr double kCw1 = kCb1 / (kKappa * kKappa) + (1.0 + kCb2) / kSigma;

    struct SaArgs
    {
      int n_list;
      const int *cell_list, *cell_un, *cell_pn, *indicator;
      const double *cell_x, *tables;
      const unsigned char *slots, *con;
      const double *fluid_present; // block vector [u | p] of the fluid solver (fluid_present_solution)
      const double *present, *eval_pt, *wall_d, *inhom;
      int n_owned;
      double nu_fluid, inv_rho, dt;
      const int64_t *rp;
      double *val, *rhs;
    };

    template <int DIM>
    struct SaQ
    {
      double JxW, N[1 << DIM], g[1 << DIM][DIM], vel[DIM], gnu[DIM];
      double nu_p, nu_c, P, D, diff;
    };

    template <int DIM>
    __device__ __forceinline__ void invert_jacobian(const double *J, double *Ji, double &det)
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

    // everything that depends on the quadrature point only (:689-790)
    template <int DIM>
    __device__ void sa_fill_qpoint(const SaArgs &A, int cell, int q, SaQ<DIM> &Q)
    {
      constexpr int NU = 1 << DIM, NQ = NU, NV = NU;
      const double *tN = A.tables, *tdN = tN + NQ * NU, *tdG = tdN + NQ * NU * DIM + NQ * NU, *tqw = tdG + NQ * NV * DIM;
      const double *X = A.cell_x + (int64_t)cell * NV * DIM;
      double J[DIM * DIM], Ji[DIM * DIM], det;
#pragma unroll
      for (int i = 0; i < DIM * DIM; ++i) J[i] = 0.0;
      for (int v = 0; v < NV; ++v)
#pragma unroll
        for (int i = 0; i < DIM; ++i)
#pragma unroll
          for (int j = 0; j < DIM; ++j) J[i * DIM + j] = fma(X[v * DIM + i], tdG[(q * NV + v) * DIM + j], J[i * DIM + j]);
      invert_jacobian<DIM>(J, Ji, det);
      Q.JxW = det * tqw[q];
      double gradv[DIM * DIM], d = 0.0;
#pragma unroll
      for (int i = 0; i < DIM * DIM; ++i) gradv[i] = 0.0;
#pragma unroll
      for (int c = 0; c < DIM; ++c) Q.vel[c] = Q.gnu[c] = 0.0;
      Q.nu_p = Q.nu_c = 0.0;
      for (int b = 0; b < NU; ++b)
        {
          const double N = tN[q * NU + b];
          Q.N[b] = N;
#pragma unroll
          for (int k = 0; k < DIM; ++k)
            {
              double s = 0.0;
#pragma unroll
              for (int j = 0; j < DIM; ++j) s = fma(tdN[(q * NU + b) * DIM + j], Ji[j * DIM + k], s);
              Q.g[b][k] = s;
            }
          const int un = A.cell_un[(int64_t)cell * NU + b], pn = A.cell_pn[(int64_t)cell * NU + b];
#pragma unroll
          for (int c = 0; c < DIM; ++c)
            {
              const double u = A.fluid_present[(int64_t)DIM * un + c];
              Q.vel[c] = fma(N, u, Q.vel[c]);
#pragma unroll
              for (int k = 0; k < DIM; ++k) gradv[c * DIM + k] = fma(u, Q.g[b][k], gradv[c * DIM + k]);
            }
          const double ne = A.eval_pt[pn];
          Q.nu_p = fma(N, A.present[pn], Q.nu_p);
          Q.nu_c = fma(N, ne, Q.nu_c);
#pragma unroll
          for (int k = 0; k < DIM; ++k) Q.gnu[k] = fma(ne, Q.g[b][k], Q.gnu[k]);
          d = fma(N, A.wall_d[pn], d); // nearest wall distance interpolated from the support points (:728-733)
        }
      // |curl v| (:747)
      double S;
      if (DIM == 2)
        S = fabs(gradv[1 * DIM + 0] - gradv[0 * DIM + 1]);
      else
        {
          const double c0 = gradv[2 * DIM + 1] - gradv[1 * DIM + 2], c1 = gradv[0 * DIM + 2] - gradv[2 * DIM + 0],
                       c2 = gradv[1 * DIM + 0] - gradv[0 * DIM + 1];
          S = sqrt(c0 * c0 + c1 * c1 + c2 * c2);
        }
      const double laminar_nu = A.indicator[cell] == 1 ? A.inv_rho : A.nu_fluid; // :713-722
      const double nu = Q.nu_p, k2d2 = kKappa * kKappa * d * d;
      const double chi = nu / laminar_nu, chi3 = chi * chi * chi;
      const double ft2 = kCt3 * exp(-kCt4 * chi * chi);
      const double fv1 = chi3 / (chi3 + kCv1 * kCv1 * kCv1);
      const double fv2 = 1.0 - chi / (1.0 + chi * fv1);
      const double S_bar = nu / k2d2 * fv2;
      const double S_tilde = S_bar >= -kCv2 * S ? S + S_bar : S + S * (kCv2 * kCv2 * S - kCv3 * S_bar) / ((kCv3 - 2 * kCv2) * S - S_bar);
      // :757-770 - see the header: the value the reference's lambda computes but does not assign
      const double r = fabs(S_tilde) > 1e-8 ? fmin(nu / (S_tilde * k2d2), 10.0) : 10.0;
      const double r2 = r * r, gg = r + kCw2 * (r2 * r2 * r2 - r), g2 = gg * gg, cw36 = kCw3 * kCw3 * kCw3 * kCw3 * kCw3 * kCw3;
      const double fw = gg * pow((1.0 + cw36) / (g2 * g2 * g2 + cw36), 1.0 / 6.0);
      const bool positive = nu >= 0; // negative S-A branch otherwise (:776-790)
      Q.P = positive ? kCb1 * (1 - ft2) * S_tilde : kCb1 * (1 - kCt3) * S;
      Q.D = positive ? (kCw1 * fw - kCb1 / (kKappa * kKappa) * ft2) / (d * d) : -kCw1 / (d * d);
      const double fn = positive ? 1.0 : (kCn1 + chi3) / (kCn1 - chi3);
      Q.diff = (laminar_nu + fn * nu) / kSigma;
    }

    template <int DIM>
    __global__ void __launch_bounds__(64) sa_assemble_kernel(SaArgs A)
    {
      constexpr int NU = 1 << DIM, NQ = NU, PAIRS = NU * NU, CPB = 64 / PAIRS, SPC = 4 * PAIRS;
      __shared__ SaQ<DIM> sq[CPB][NQ];
      __shared__ double lrhs[CPB][NU], ldiag[CPB][NU];
      const int cl = threadIdx.x / PAIRS, pr = threadIdx.x % PAIRS;
      const int li = blockIdx.x * CPB + cl;
      const bool active = li < A.n_list;
      const int cell = active ? A.cell_list[li] : 0;
      if (threadIdx.x < CPB * NQ)
        {
          const int c2 = threadIdx.x / NQ, q = threadIdx.x % NQ, l2 = blockIdx.x * CPB + c2;
          if (l2 < A.n_list) sa_fill_qpoint<DIM>(A, A.cell_list[l2], q, sq[c2][q]);
        }
      if (pr < NU)
        {
          lrhs[cl][pr] = 0.0;
          ldiag[cl][pr] = 0.0;
        }
      __syncthreads();
      const int a = pr / NU, b = pr % NU;
      if (active)
        {
          double K = 0.0, r = 0.0;
          for (int q = 0; q < NQ; ++q)
            {
              const SaQ<DIM> &Q = sq[cl][q];
              const double Na = Q.N[a], Nb = Q.N[b];
              double gagb = 0.0, ugb = 0.0, gb_gnu = 0.0;
#pragma unroll
              for (int k = 0; k < DIM; ++k)
                {
                  gagb = fma(Q.g[a][k], Q.g[b][k], gagb);
                  ugb = fma(Q.vel[k], Q.g[b][k], ugb);
                  gb_gnu = fma(Q.g[b][k], Q.gnu[k], gb_gnu);
                }
              // :799-815
              const double m = Na * Nb / A.dt + Na * ugb + Q.diff * gagb - 2 * kCb2 / kSigma * Na * gb_gnu - Q.P * Na * Nb +
                               2 * Q.D * Na * Nb * Q.nu_c;
              K = fma(m, Q.JxW, K);
              if (b == 0)
                {
                  double u_gnu = 0.0, ga_gnu = 0.0, gnu2 = 0.0;
#pragma unroll
                  for (int k = 0; k < DIM; ++k)
                    {
                      u_gnu = fma(Q.vel[k], Q.gnu[k], u_gnu);
                      ga_gnu = fma(Q.g[a][k], Q.gnu[k], ga_gnu);
                      gnu2 = fma(Q.gnu[k], Q.gnu[k], gnu2);
                    }
                  // :818-835
                  const double v = Na * (Q.nu_c - Q.nu_p) / A.dt + Na * u_gnu + Q.diff * ga_gnu - kCb2 / kSigma * Na * gnu2 - Q.P * Na * Q.nu_c +
                                   Q.D * Na * Q.nu_c * Q.nu_c;
                  r = fma(-v, Q.JxW, r);
                }
            }
          // scatter through the constraints (:817-826)
          const int nA = A.cell_pn[(int64_t)cell * NU + a], nB = A.cell_pn[(int64_t)cell * NU + b];
          if (nA < A.n_owned)
            {
              const int slot = A.slots[(int64_t)cell * SPC + 3 * PAIRS + pr];
              double *dst = A.val + A.rp[nA] + slot;
              if (A.con[nA])
                {
                  if (a == b)
                    {
                      const double dv = fabs(K);
                      *dst += dv;
                      ldiag[cl][a] = dv;
                    }
                }
              else
                {
                  double add = b == 0 ? r : 0.0;
                  if (A.con[nB])
                    {
                      if (A.inhom) add -= K * A.inhom[nB];
                    }
                  else
                    *dst += K;
                  if (add != 0.0) atomicAdd(&lrhs[cl][a], add);
                }
            }
        }
      __syncthreads();
      if (active && pr < NU)
        {
          const int n = A.cell_pn[(int64_t)cell * NU + pr];
          if (n < A.n_owned)
            {
              if (!A.con[n]) A.rhs[n] += lrhs[cl][pr];
              else if (A.inhom) A.rhs[n] += ldiag[cl][pr] * A.inhom[n];
            }
        }
    }

    template <int DIM>
    __global__ void wall_distance_kernel(int n_nodes, const double *__restrict__ coords, int n_wall, const double *__restrict__ wall,
                                         double *__restrict__ dist)
    {
      constexpr int TILE = 256;
      __shared__ double tile[TILE * DIM];
      const int i = blockIdx.x * blockDim.x + threadIdx.x;
      double x[DIM];
#pragma unroll
      for (int d = 0; d < DIM; ++d) x[d] = i < n_nodes ? coords[(int64_t)i * DIM + d] : 0.0;
      double best = DBL_MAX; // std::numeric_limits<double>::max() when there is no wall (:521)
      for (int w0 = 0; w0 < n_wall; w0 += TILE)
        {
          const int nt = min(TILE, n_wall - w0);
          __syncthreads();
          for (int t = threadIdx.x; t < nt * DIM; t += blockDim.x) tile[t] = wall[(int64_t)w0 * DIM + t];
          __syncthreads();
          for (int t = 0; t < nt; ++t)
            {
              double s = 0.0;
#pragma unroll
              for (int d = 0; d < DIM; ++d)
                {
                  const double e = tile[t * DIM + d] - x[d];
                  s = fma(e, e, s);
                }
              best = fmin(best, s);
            }
        }
      if (i < n_nodes) dist[i] = n_wall ? sqrt(best) : DBL_MAX;
    }

    __global__ void eddy_viscosity_kernel(int n, const double *__restrict__ nu_tilde, double laminar_nu, double rho, double *__restrict__ out)
    {
      const int i = blockIdx.x * blockDim.x + threadIdx.x;
      if (i >= n) return;
      const double v = nu_tilde[i], chi = v / laminar_nu, chi3 = chi * chi * chi;
      out[i] = chi3 / (chi3 + kCv1 * kCv1 * kCv1) * v * rho;
    }

    // every node of a cell inside the solid: line nu~_new = 0, i.e. update = -present (:159-181). Several cells may write the same
    // node; they write the same values.
    __global__ void solid_lines_kernel(int n_cells, int npc, const int *__restrict__ cell_pn, const int *__restrict__ indicator,
                                       const double *__restrict__ present, unsigned char *__restrict__ con, double *__restrict__ val)
    {
      const int c = blockIdx.x * blockDim.x + threadIdx.x;
      if (c >= n_cells || indicator[c] != 1) return;
      for (int a = 0; a < npc; ++a)
        {
          const int n = cell_pn[(int64_t)c * npc + a];
          con[n] = 1;
          val[n] = -present[n];
        }
    }

    __global__ void jacobi_inverse_kernel(int n, const int64_t *__restrict__ rp, const int *__restrict__ col, const double *__restrict__ val,
                                          double *__restrict__ out)
    {
      const int i = blockIdx.x * blockDim.x + threadIdx.x;
      if (i >= n) return;
      double d = 0.0;
      for (int64_t k = rp[i]; k < rp[i + 1]; ++k)
        if (col[k] == i) d = val[k];
      out[i] = d != 0.0 ? 1.0 / d : 1.0;
    }

    inline unsigned blocks(int64_t n, int t = 128) { return (unsigned)((n + t - 1) / t); }
  } // namespace

  SpalartAllmaras::SpalartAllmaras(Context &ctx_, SCnsIM &fluid_) : ctx(ctx_), fluid(fluid_) {}

  void SpalartAllmaras::make_constraints()
  {
    FluidSpace &fs = fluid.fs;
    const Triangulation &tria = fluid.triangulation;
    const Parameters::AllParameters &prm = fluid.parameters;
    if (constraints_made && base_nodes == (size_t)fs.pn.n_nodes) // same mesh, same .prm: restore the device copies
      {
        cudaStream_t s = ctx.stream;
        IFEM_CUDA(cudaMemcpyAsync(d_con.p, d_base_con.p, base_nodes, cudaMemcpyDeviceToDevice, s));
        IFEM_CUDA(cudaMemcpyAsync(d_nonzero_val.p, d_base_val.p, base_nodes * sizeof(double), cudaMemcpyDeviceToDevice, s));
        return;
      }
    // like FluidSpace::make_constraints: over the GLOBAL boundary faces in ascending boundary id, first line on a node wins,
    // hanging nodes keep their hanging-node line (made first, :371-374)
    const NodeTable &g = fs.n_ranks > 1 ? fs.pn_global : fs.pn;
    const std::vector<char> hflag = fs.hanging.active ? hanging_node_flags(tria, g) : std::vector<char>();
    std::vector<unsigned char> gcon((size_t)g.n_nodes, 0);
    std::vector<double> gval((size_t)g.n_nodes, 0.0);
    const int dim = fs.dim, np = fs.np;
    for (const auto &bc : prm.spalart_allmaras_model_bcs)
      {
        double value = 0.0;
        if (bc.second == 1)
          value = 5.0 * prm.viscosity / prm.fluid_rho;
        else if (bc.second != 0)
          throw std::runtime_error("Unrecogonized Spalart-Allmaras BC type!");
        for (int f = 0; f < tria.n_boundary_faces(); ++f)
          {
            if (tria.boundary_faces[3 * f + 2] != (int)bc.first) continue;
            const int cell = tria.boundary_faces[3 * f];
            for (int a : face_local_nodes(dim, fs.pp, tria.boundary_faces[3 * f + 1]))
              {
                const int node = g.cell_nodes[(size_t)cell * np + a];
                if ((!hflag.empty() && hflag[node]) || gcon[node]) continue;
                gcon[node] = 1;
                gval[node] = value;
              }
          }
      }
    std::vector<unsigned char> con((size_t)fs.pn.n_nodes);
    std::vector<double> val((size_t)fs.pn.n_nodes);
    for (int l = 0; l < fs.pn.n_nodes; ++l)
      {
        const int gn = fs.n_ranks > 1 ? fs.part.p.local_to_global[l] : l;
        con[l] = gcon[gn];
        val[l] = gval[gn];
      }
    cudaStream_t s = ctx.stream;
    d_base_con.upload(con, s);
    d_base_val.upload(val, s);
    base_nodes = con.size();
    d_con.upload(con, s);
    d_nonzero_val.upload(val, s);
    IFEM_CUDA(cudaStreamSynchronize(s));
    constraints_made = true;
  }

  void SpalartAllmaras::update_boundary_condition(bool first_step)
  {
    FluidSpace &fs = fluid.fs;
    cudaStream_t s = ctx.stream;
    // the caller re-made the fluid's constraints for this step (source/mpi_fsi.cpp:1192), which re-makes the model's (:276-279)
    IFEM_CUDA(cudaMemcpyAsync(d_con.p, d_base_con.p, d_con.n, cudaMemcpyDeviceToDevice, s));
    if (first_step)
      IFEM_CUDA(cudaMemcpyAsync(d_nonzero_val.p, d_base_val.p, d_nonzero_val.n * sizeof(double), cudaMemcpyDeviceToDevice, s));
    else
      d_nonzero_val.zero(s); // nonzero_constraints.copy_from(zero_constraints), :143-147
    solid_lines_kernel<<<blocks(fs.n_cells), 128, 0, s>>>(fs.n_cells, fs.np, fs.d_cell_pn.p, fs.d_indicator.p, present_solution.p, d_con.p,
                                                         d_nonzero_val.p);
    IFEM_KERNEL_CHECK();
    ctx.kernel_launches++;
  }

  void SpalartAllmaras::initialize_system()
  {
    FluidSpace &fs = fluid.fs;
    if (fs.pu != 1 || fs.pp != 1) throw std::runtime_error("SpalartAllmaras: equal-order Q1/Q1 fluid solvers only");
    if (!constraints_made) make_constraints();
    cudaStream_t s = ctx.stream;
    const size_t n = (size_t)fs.pn.n_nodes;
    system_matrix.init(fs.P_pp, 1, 1, s);
    system_matrix.n_brows_spmv = fs.A_pp.n_brows_spmv;
    for (DevBuf<double> *v : {&present_solution, &evaluation_point, &newton_update, &eddy_viscosity, &fixed_wall_distance, &system_rhs,
                              &d_diag_inv})
      {
        v->alloc(n);
        v->zero(s);
      }
    ilu = Ilu0();
    // initial condition (:573-579): coefficient x laminar nu, then zero_constraints.distribute
    const Parameters::AllParameters &prm = fluid.parameters;
    fill(ctx, VecSpace((int64_t)n), prm.spalart_allmaras_initial_condition_coefficient * prm.viscosity / prm.fluid_rho, newton_update.p);
    set_flagged(ctx, (int64_t)n, d_con.p, nullptr, newton_update.p);
    fs.hanging.distribute_scalar(ctx, newton_update.p);
    copy(ctx, VecSpace((int64_t)n), newton_update.p, present_solution.p);
    setup_cell_property();
    history.clear();
    ready = true;
  }

  void SpalartAllmaras::setup_cell_property()
  {
    FluidSpace &fs = fluid.fs;
    const Triangulation &tria = fluid.triangulation;
    const Parameters::AllParameters &prm = fluid.parameters;
    const int dim = fs.dim;
    // vertices of the wall faces (type 0) of the whole mesh (:436-482: collected per rank, then all-reduced)
    std::set<int> wall_vertices;
    for (int f = 0; f < tria.n_boundary_faces(); ++f)
      {
        const auto it = prm.spalart_allmaras_model_bcs.find((unsigned)tria.boundary_faces[3 * f + 2]);
        if (it == prm.spalart_allmaras_model_bcs.end() || it->second != 0) continue;
        const int cell = tria.boundary_faces[3 * f];
        for (int a : face_local_nodes(dim, 1, tria.boundary_faces[3 * f + 1])) wall_vertices.insert(tria.cells[(size_t)cell * tria.verts_per_cell() + a]);
      }
    std::vector<double> wall;
    for (int v : wall_vertices)
      for (int d = 0; d < dim; ++d) wall.push_back(tria.vertices[(size_t)v * dim + d]);
    cudaStream_t s = ctx.stream;
    DevBuf<double> d_wall, d_coords;
    d_wall.upload(wall, s);
    d_coords.upload(fs.pn.coords, s);
    const int n = fs.pn.n_nodes, nw = (int)wall_vertices.size();
    if (dim == 2)
      wall_distance_kernel<2><<<blocks(n, 256), 256, 0, s>>>(n, d_coords.p, nw, d_wall.p, fixed_wall_distance.p);
    else
      wall_distance_kernel<3><<<blocks(n, 256), 256, 0, s>>>(n, d_coords.p, nw, d_wall.p, fixed_wall_distance.p);
    IFEM_KERNEL_CHECK();
    ctx.kernel_launches++;
    IFEM_CUDA(cudaStreamSynchronize(s));
  }

  void SpalartAllmaras::assemble(bool use_nonzero_constraints)
  {
    FluidSpace &fs = fluid.fs;
    cudaStream_t s = ctx.stream;
    if (fs.n_ranks > 1)
      {
        fs.halo_update(ctx, fluid.present_solution.p);
        fs.halo_p.update(ctx, present_solution.p);
        fs.halo_p.update(ctx, evaluation_point.p);
      }
    system_matrix.zero(s);
    system_rhs.zero(s);
    SaArgs a{};
    a.cell_un = fs.d_cell_un.p;
    a.cell_pn = fs.d_cell_pn.p;
    a.indicator = fs.d_indicator.p;
    a.cell_x = fs.d_cell_x.p;
    a.tables = fs.d_tables.p;
    a.slots = fs.d_slots.p;
    a.con = d_con.p;
    a.fluid_present = fluid.present_solution.p;
    a.present = present_solution.p;
    a.eval_pt = evaluation_point.p;
    a.wall_d = fixed_wall_distance.p;
    a.inhom = use_nonzero_constraints ? d_nonzero_val.p : nullptr;
    a.n_owned = fs.n_owned_pnodes;
    a.nu_fluid = fluid.parameters.viscosity / fluid.parameters.fluid_rho;
    a.inv_rho = 1.0 / fluid.parameters.fluid_rho;
    a.dt = fluid.time.get_delta_t();
    a.rp = system_matrix.rowptr.p;
    a.val = system_matrix.val.p;
    a.rhs = system_rhs.p;
    const int n_colours = (int)fs.colour_offsets.size() - 1;
    for (int k = 0; k < n_colours; ++k)
      {
        a.n_list = fs.colour_offsets[k + 1] - fs.colour_offsets[k];
        a.cell_list = fs.d_colour_order.p + fs.colour_offsets[k];
        if (!a.n_list) continue;
        if (fs.dim == 2)
          sa_assemble_kernel<2><<<(a.n_list + 3) / 4, 64, 0, s>>>(a);
        else
          sa_assemble_kernel<3><<<a.n_list, 64, 0, s>>>(a);
        IFEM_KERNEL_CHECK();
        ctx.kernel_launches++;
      }
    fs.hanging.condense_scalar(ctx, fs, system_matrix, system_rhs.p, d_con.p, a.inhom);
  }

  bool SpalartAllmaras::use_ilu() const
  {
    // the reference's Euclid ILU(0) on one rank up to the size the one-CTA sweeps pay (ilu0.h), Jacobi otherwise - the transport
    // matrix is dominated by its mass term M / dt
    const FluidSpace &fs = fluid.fs;
    if (fluid.control.supg_ilu == 0 || fs.n_ranks > 1) return false;
    return fluid.control.supg_ilu == 1 || fs.pn.n_nodes <= 60000;
  }

  std::pair<unsigned int, double> SpalartAllmaras::solve(bool use_nonzero_constraints)
  {
    FluidSpace &fs = fluid.fs;
    cudaStream_t s = ctx.stream;
    const VecSpace &vp = fs.vs_p;
    const bool with_ilu = use_ilu();
    if (with_ilu)
      {
        if (!ilu.ready()) ilu.setup(ctx, fs.P_pp.rowptr, fs.P_pp.col);
        bcsr_to_scalar(ctx, system_matrix, ilu.rowptr.p, ilu.val.p);
        ilu.factor(ctx);
      }
    else
      {
        jacobi_inverse_kernel<<<blocks(fs.n_owned_pnodes), 128, 0, s>>>(fs.n_owned_pnodes, system_matrix.rowptr.p, system_matrix.col.p,
                                                                        system_matrix.val.p, d_diag_inv.p);
        IFEM_KERNEL_CHECK();
        ctx.kernel_launches++;
      }
    const double nrm = nrm2(ctx, vp, system_rhs.p);
    const double tol = 1e-8 * nrm; // SolverControl(2 n, 1e-8 |rhs|), :840-841
    int64_t n_global = fs.n_owned_pnodes;
    if (fs.n_ranks > 1) n_global = fs.pn_global.n_nodes;
    LinOp A = [&](const double *x, double *y) {
      fs.halo_p.update(ctx, const_cast<double *>(x));
      spmv(ctx, system_matrix, x, y);
    };
    LinOp P = [&](const double *x, double *y) {
      if (with_ilu) ilu.solve(ctx, x, y);
      else hadamard(ctx, vp, d_diag_inv.p, x, y);
    };
    SolveResult r;
    if (nrm > 0)
      r = fgmres(ctx, vp, A, P, system_rhs.p, newton_update.p, tol, 2 * n_global, 30, pool);
    else
      fill(ctx, vp, 0.0, newton_update.p);
    set_flagged(ctx, fs.pn.n_nodes, d_con.p, use_nonzero_constraints ? d_nonzero_val.p : nullptr, newton_update.p);
    fs.hanging.distribute_scalar(ctx, newton_update.p); // constraints_used.distribute(newton_update), :855-858
    return {(unsigned)r.iterations, r.residual};
  }

  void SpalartAllmaras::run_one_step(bool apply_nonzero_constraints)
  {
    if (!ready) throw std::runtime_error("SpalartAllmaras::run_one_step before initialize_system");
    FluidSpace &fs = fluid.fs;
    const Parameters::AllParameters &prm = fluid.parameters;
    const VecSpace &vp = fs.vs_p;
    if (verbose && fs.rank == 0) std::printf("%s\nSolving for S-A turbulence model...\n", std::string(96, '*').c_str());
    double current_residual = 1.0, initial_residual = 1.0, relative_residual = 1.0;
    unsigned int outer_iteration = 0;
    copy(ctx, vp, present_solution.p, evaluation_point.p);
    while (relative_residual > prm.fluid_tolerance && current_residual > 1e-14)
      {
        if (outer_iteration >= prm.fluid_max_iterations) throw std::runtime_error("Too many Newton iterations!");
        fill(ctx, vp, 0.0, newton_update.p);
        const bool nz = apply_nonzero_constraints && outer_iteration == 0;
        assemble(nz);
        const auto state = solve(nz);
        current_residual = nrm2(ctx, vp, system_rhs.p);
        axpy(ctx, vp, 1.0, newton_update.p, evaluation_point.p);
        if (outer_iteration == 0) initial_residual = current_residual;
        relative_residual = current_residual / initial_residual;
        history.push_back({outer_iteration, current_residual, relative_residual, (int)state.first, state.second});
        if (verbose && fs.rank == 0)
          std::printf(" ITR = %-2u ABS_RES = %e REL_RES = %e GMRES_ITR = %-3u GMRES_RES = %e\n", outer_iteration, current_residual,
                      relative_residual, state.first, state.second);
        outer_iteration++;
      }
    copy(ctx, vp, evaluation_point.p, present_solution.p);
    fs.halo_p.update(ctx, present_solution.p);
    update_eddy_viscosity();
  }

  void SpalartAllmaras::update_eddy_viscosity()
  {
    FluidSpace &fs = fluid.fs;
    const Parameters::AllParameters &prm = fluid.parameters;
    const int n = fs.pn.n_nodes; // ghost entries too: present_solution was refreshed, the formula is pointwise
    eddy_viscosity_kernel<<<blocks(n), 128, 0, ctx.stream>>>(n, present_solution.p, prm.viscosity / prm.fluid_rho, prm.fluid_rho, eddy_viscosity.p);
    IFEM_KERNEL_CHECK();
    ctx.kernel_launches++;
  }

  // Newton iteration on the composite law of the wall u+ (y+) (:227-293). Host arithmetic: FSI::find_solid_bc calls it once per
  // solid boundary vertex (source/mpi_fsi.cpp:835-838).
  double SpalartAllmaras::get_shear_velocity(double vel, double init_guess) const
  {
    const Parameters::AllParameters &prm = fluid.parameters;
    if (std::fabs(vel) < 1e-10) return 0.0;
    const double nu = prm.viscosity / prm.fluid_rho, dist = prm.spalart_allmaras_image_distance;
    if (vel * dist / nu < std::sqrt(5.0)) return vel / std::sqrt(vel * dist / nu);
    init_guess = std::max(init_guess, 5.0 * nu / dist);
    const double B = 5.03339088, a1 = 8.14822158, a2 = -6.92870938, b1 = 7.46008761, b2 = 7.46814579, c1 = 2.54967735, c2 = 1.33016516,
                 c3 = 3.59945911, c4 = 3.63975319;
    auto sq = [](double x) { return x * x; };
    auto u_plus = [&](double yp) {
      return B + c1 * std::log(sq(yp + a1) + sq(b1)) - c2 * std::log(sq(yp + a2) + sq(b2)) - c3 * std::atan2(b1, yp + a1) -
             c4 * std::atan2(b2, yp + a2);
    };
    const double kappa3 = kKappa * kKappa * kKappa, cv13 = kCv1 * kCv1 * kCv1;
    auto dup_dyp = [&](double yp) { return (kappa3 * yp * yp * yp) / (cv13 + kappa3 * yp * yp * yp); };
    double ut = init_guess;
    for (int i = 0; i < 30; ++i)
      {
        const double yp = ut * dist / nu, up = u_plus(yp);
        const double next = ut - (ut * up - vel) / (up + ut * dist / nu * dup_dyp(yp));
        const bool done = std::fabs(next - ut) < 1e-2 * std::fabs(ut);
        ut = next;
        if (done) break;
      }
    return ut;
  }
} // namespace ifem
