#include "solid.h"

#include "hyper_materials.cuh"
#include "solid_linear.cuh"

#include <algorithm>
#include <chrono>
#include <cmath>

namespace ifem
{
  namespace
  {
    // ---- NeoHookean point update: PointHistory::update + HyperElasticMaterial::update_data ----------
    // (mpi_hyper_elasticity.cpp:37-65, hyper_elastic_material.cpp:8-39, hyper_elastic_material.h:50-70,
    //  neo_hookean.h:26-34). Voigt pair order: (0,0),(1,1)[,(2,2)],(0,1)[,(0,2),(1,2)].
    template <int DIM>
    struct Voigt;
    template <>
    struct Voigt<2>
    {
      static constexpr int N = 3;
      __host__ __device__ static int idx(int i, int j) { return i == j ? i : 2; }
    };
    template <>
    struct Voigt<3>
    {
      static constexpr int N = 6;
      __host__ __device__ static int idx(int i, int j)
      {
        if (i == j) return i;
        const int a = i < j ? i : j, b = i < j ? j : i;
        return a == 0 ? (b == 1 ? 3 : 4) : 5;
      }
    };

    template <int DIM>
    __device__ void neo_hookean_point(const double *gu /*[DIM*DIM] Grad u*/, double c1, double kappa, double *Finv, double *tau,
                                      double *Jc /*[N*N]*/, double &detF)
    {
      constexpr int NS = Voigt<DIM>::N;
      double F[DIM * DIM];
#pragma unroll
      for (int i = 0; i < DIM; ++i)
#pragma unroll
        for (int j = 0; j < DIM; ++j) F[i * DIM + j] = gu[i * DIM + j] + (i == j ? 1.0 : 0.0);
      double J;
      if (DIM == 2)
        {
          J = F[0] * F[3] - F[1] * F[2];
          const double d = 1.0 / J;
          Finv[0] = F[3] * d; Finv[1] = -F[1] * d; Finv[2] = -F[2] * d; Finv[3] = F[0] * d;
        }
      else
        {
          const double c00 = F[4] * F[8] - F[5] * F[7], c01 = F[5] * F[6] - F[3] * F[8], c02 = F[3] * F[7] - F[4] * F[6];
          J = F[0] * c00 + F[1] * c01 + F[2] * c02;
          const double d = 1.0 / J;
          Finv[0] = c00 * d; Finv[1] = (F[2] * F[7] - F[1] * F[8]) * d; Finv[2] = (F[1] * F[5] - F[2] * F[4]) * d;
          Finv[3] = c01 * d; Finv[4] = (F[0] * F[8] - F[2] * F[6]) * d; Finv[5] = (F[2] * F[3] - F[0] * F[5]) * d;
          Finv[6] = c02 * d; Finv[7] = (F[1] * F[6] - F[0] * F[7]) * d; Finv[8] = (F[0] * F[4] - F[1] * F[3]) * d;
        }
      detF = J;
      const double s = pow(J, -2.0 / DIM); // (J^(-1/dim))^2 on b = F F^T
      double tb[DIM * DIM], tr = 0.0;      // tau_bar = 2 c1 b_bar
#pragma unroll
      for (int i = 0; i < DIM; ++i)
#pragma unroll
        for (int j = 0; j < DIM; ++j)
          {
            double b = 0.0;
#pragma unroll
            for (int k = 0; k < DIM; ++k) b = fma(F[i * DIM + k], F[j * DIM + k], b);
            tb[i * DIM + j] = 2.0 * c1 * s * b;
          }
#pragma unroll
      for (int i = 0; i < DIM; ++i) tr += tb[i * DIM + i];
      const double p = kappa * (J - 1.0), pt = p + J * kappa;
      double tiso[DIM * DIM];
#pragma unroll
      for (int i = 0; i < DIM; ++i)
#pragma unroll
        for (int j = 0; j < DIM; ++j)
          {
            tiso[i * DIM + j] = tb[i * DIM + j] - (i == j ? tr / DIM : 0.0);
            tau[i * DIM + j] = tiso[i * DIM + j] + (i == j ? J * p : 0.0);
          }
      // Jc_ijkl = (2/d) tr (S - IxI/d) - (2/d)(tiso_ij d_kl + d_ij tiso_kl) + J (pt IxI - 2 p S)
#pragma unroll
      for (int i = 0; i < DIM; ++i)
#pragma unroll
        for (int j = i; j < DIM; ++j)
#pragma unroll
          for (int k = 0; k < DIM; ++k)
#pragma unroll
            for (int l = k; l < DIM; ++l)
              {
                const double dij = i == j, dkl = k == l;
                const double S = 0.5 * ((i == k && j == l ? 1.0 : 0.0) + (i == l && j == k ? 1.0 : 0.0));
                const double v = (2.0 / DIM) * tr * (S - dij * dkl / DIM) - (2.0 / DIM) * (tiso[i * DIM + j] * dkl + dij * tiso[k * DIM + l]) +
                                 J * (pt * dij * dkl - 2.0 * p * S);
                Jc[Voigt<DIM>::idx(i, j) * NS + Voigt<DIM>::idx(k, l)] = v;
              }
    }

    template <int DIM, int NPC>
    __global__ void update_qph_kernel(int n_cells, int nq, const int *__restrict__ cell_nodes, const double *__restrict__ G,
                                      const double *__restrict__ u, int material, const double *__restrict__ cell_mat, double *__restrict__ Finv,
                                      double *__restrict__ tau, double *__restrict__ Jc, double *__restrict__ detF)
    {
      constexpr int NS = Voigt<DIM>::N;
      const int t = blockIdx.x * blockDim.x + threadIdx.x;
      if (t >= n_cells * nq) return;
      const int cell = t / nq;
      const double *g = G + (int64_t)t * NPC * DIM;
      double gu[DIM * DIM];
#pragma unroll
      for (int i = 0; i < DIM * DIM; ++i) gu[i] = 0.0;
      for (int a = 0; a < NPC; ++a)
        {
          const int node = cell_nodes[(int64_t)cell * NPC + a];
#pragma unroll
          for (int c = 0; c < DIM; ++c)
            {
              const double uc = u[(int64_t)DIM * node + c];
#pragma unroll
              for (int k = 0; k < DIM; ++k) gu[c * DIM + k] = fma(uc, g[a * DIM + k], gu[c * DIM + k]);
            }
        }
      double fi[DIM * DIM], ta[DIM * DIM], jc[NS * NS], dj;
      // material of the cell: parameters of part cell->material_id() (mpi_hyper_elasticity.cpp:226-228)
      const double c1 = cell_mat[2 * cell], kappa = cell_mat[2 * cell + 1];
      if (material == 0)
        neo_hookean_point<DIM>(gu, c1, kappa, fi, ta, jc, dj);
      else
        kirchhoff_point<DIM>(gu, c1, kappa, fi, ta, jc, dj); // (c1, kappa) carry (Young's modulus, Poisson's ratio)
#pragma unroll
      for (int i = 0; i < DIM * DIM; ++i)
        {
          Finv[(int64_t)t * DIM * DIM + i] = fi[i];
          tau[(int64_t)t * DIM * DIM + i] = ta[i];
        }
#pragma unroll
      for (int i = 0; i < NS * NS; ++i) Jc[(int64_t)t * NS * NS + i] = jc[i];
      detF[t] = dj;
    }

    __global__ void solid_slots_kernel(int n_cells, int npc, const int *__restrict__ tab, const int64_t *__restrict__ rowptr,
                                       const int *__restrict__ col, unsigned char *__restrict__ slots, int *__restrict__ err)
    {
      const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
      if (t >= (int64_t)n_cells * npc * npc) return;
      const int cell = (int)(t / (npc * npc)), rem = (int)(t % (npc * npc));
      const int A = tab[(int64_t)cell * npc + rem / npc], B = tab[(int64_t)cell * npc + rem % npc];
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
      slots[t] = (unsigned char)j;
    }

    struct HyperArgs
    {
      int n_list;
      const int *cell_list, *cell_nodes;
      const unsigned char *slots, *con;
      const double *N, *G, *JxW, *Finv, *tau, *Jc;
      int nq;
      double rho, inv_beta_dt2, grav[3];
      int initial_step;
      const int64_t *rowptr;
      double *val, *rhs;
    };

    // HyperElasticity::assemble_system cell loop (mpi_hyper_elasticity.cpp:362-438, 507-522): one thread per
    // (cell, row node a, column node b), cells of one colour per launch (race-free read-modify-write).
    template <int DIM, int NPC>
    __global__ void __launch_bounds__(64) hyper_assemble_kernel(const HyperArgs A)
    {
      constexpr int NS = Voigt<DIM>::N, PAIRS = NPC * NPC, CPB = 64 / PAIRS;
      const int li = blockIdx.x * CPB + threadIdx.x / PAIRS;
      if (li >= A.n_list) return;
      const int cell = A.cell_list[li];
      const int pr = threadIdx.x % PAIRS, a = pr / NPC, b = pr % NPC;
      const int nq = A.nq;
      double K[DIM * DIM], r[DIM];
#pragma unroll
      for (int i = 0; i < DIM * DIM; ++i) K[i] = 0.0;
#pragma unroll
      for (int i = 0; i < DIM; ++i) r[i] = 0.0;
      for (int q = 0; q < nq; ++q)
        {
          const int64_t cq = (int64_t)cell * nq + q;
          const double w = A.JxW[cq];
          const double *Fi = A.Finv + cq * DIM * DIM, *ta = A.tau + cq * DIM * DIM, *jc = A.Jc + cq * NS * NS;
          const double *g0 = A.G + cq * NPC * DIM;
          const double Na = A.N[q * NPC + a], Nb = A.N[q * NPC + b];
          // pushed-forward gradients: grad_phi = Grad N * F_inv (:384-386)
          double ga[DIM], gb[DIM];
#pragma unroll
          for (int k = 0; k < DIM; ++k)
            {
              double sa = 0.0, sb = 0.0;
#pragma unroll
              for (int m = 0; m < DIM; ++m)
                {
                  sa = fma(g0[a * DIM + m], Fi[m * DIM + k], sa);
                  sb = fma(g0[b * DIM + m], Fi[m * DIM + k], sb);
                }
              ga[k] = sa;
              gb[k] = sb;
            }
          if (A.initial_step)
            {
#pragma unroll
              for (int c = 0; c < DIM; ++c) K[c * DIM + c] = fma(A.rho * Na * Nb, w, K[c * DIM + c]);
            }
          else
            {
              // geometric term for equal components: grad_phi_i[c] . tau . grad_phi_j[c]
              double geo = 0.0;
#pragma unroll
              for (int k = 0; k < DIM; ++k)
#pragma unroll
                for (int l = 0; l < DIM; ++l) geo = fma(ga[k] * ta[k * DIM + l], gb[l], geo);
              const double mass = A.rho * A.inv_beta_dt2 * Na * Nb;
#pragma unroll
              for (int c = 0; c < DIM; ++c)
#pragma unroll
                for (int d = 0; d < DIM; ++d)
                  {
                    // sym_grad_phi_i : Jc : sym_grad_phi_j = sum_kl ga[k] Jc[c k d l] gb[l]   (minor symmetries)
                    double s = 0.0;
#pragma unroll
                    for (int k = 0; k < DIM; ++k)
#pragma unroll
                      for (int l = 0; l < DIM; ++l) s = fma(ga[k] * jc[Voigt<DIM>::idx(c, k) * NS + Voigt<DIM>::idx(d, l)], gb[l], s);
                    if (c == d) s += mass + geo;
                    K[c * DIM + d] = fma(s, w, K[c * DIM + d]);
                  }
            }
          if (b == 0)
            {
              // -internal force + body force (:421-424)
#pragma unroll
              for (int c = 0; c < DIM; ++c)
                {
                  double s = 0.0;
#pragma unroll
                  for (int k = 0; k < DIM; ++k) s = fma(ga[k], ta[c * DIM + k], s);
                  r[c] = fma(w, -s + A.rho * A.grav[c] * Na, r[c]);
                }
            }
        }
      const int nA = A.cell_nodes[(int64_t)cell * NPC + a], nB = A.cell_nodes[(int64_t)cell * NPC + b];
      const int64_t rp = A.rowptr[nA];
      const int nb = (int)(A.rowptr[nA + 1] - rp);
      double *base = A.val + rp * DIM * DIM;
      const int slot = A.slots[(int64_t)cell * PAIRS + pr];
#pragma unroll
      for (int c = 0; c < DIM; ++c)
        {
          const int rc = A.con[(int64_t)DIM * nA + c];
#pragma unroll
          for (int d = 0; d < DIM; ++d)
            {
              const int cc = A.con[(int64_t)DIM * nB + d];
              if (rc)
                {
                  if (a == b && c == d) base[(int64_t)(c * DIM + d) * nb + slot] += fabs(K[c * DIM + d]);
                }
              else if (!cc)
                base[(int64_t)(c * DIM + d) * nb + slot] += K[c * DIM + d];
            }
          if (b == 0 && !rc) A.rhs[(int64_t)DIM * nA + c] += r[c];
        }
    }

    // Neumann faces (:445-505): traction or pressure (normal w.r.t. the reference configuration)
    template <int DIM, int NPC>
    __global__ void solid_neumann_kernel(int n_faces, int nqf, const int *__restrict__ faces, const double *__restrict__ vals,
                                         int is_pressure, const double *__restrict__ ftab, const int *__restrict__ cell_nodes,
                                         const double *__restrict__ node_x, const unsigned char *__restrict__ con,
                                         double *__restrict__ rhs, const double *__restrict__ fsi_rows, const double *__restrict__ disp,
                                         int64_t n_dofs)
    {
      // fsi_rows != null: FSI traction sigma_f n on the DEFORMED face (mpi_shared_hyper_elasticity.cpp:495-554),
      // sigma_f(q) interpolated from the vertex values MPI::FSI::find_solid_bc wrote into fsi_stress_rows
      constexpr int NV = 1 << DIM;
      const int f = blockIdx.x;
      if (f >= n_faces) return;
      const int cell = faces[2 * f], face = faces[2 * f + 1], axis = face / 2, side = face % 2;
      const double *Nf = ftab, *Gf = ftab + (size_t)2 * DIM * nqf * NPC, *qwf = Gf + (size_t)2 * DIM * nqf * NV * DIM;
      for (int i = threadIdx.x; i < NPC * DIM; i += blockDim.x)
        {
          const int a = i / DIM, c = i % DIM;
          double r = 0.0;
          for (int q = 0; q < nqf; ++q)
            {
              const size_t fq = (size_t)face * nqf + q;
              double J[DIM * DIM];
              for (int k = 0; k < DIM * DIM; ++k) J[k] = 0.0;
              // vertices of a Q1 cell are its nodes (degree 1): geometry from the node coordinates
              for (int v = 0; v < NV; ++v)
                {
                  const int64_t vn = cell_nodes[(int64_t)cell * NPC + v];
                  const double *X = node_x + vn * DIM;
                  for (int ii = 0; ii < DIM; ++ii)
                    {
                      const double xv = X[ii] + (fsi_rows ? disp[vn * DIM + ii] : 0.0);
                      for (int jj = 0; jj < DIM; ++jj) J[ii * DIM + jj] = fma(xv, Gf[(fq * NV + v) * DIM + jj], J[ii * DIM + jj]);
                    }
                }
              double nds[DIM], det;
              if (DIM == 2)
                {
                  det = J[0] * J[3] - J[1] * J[2];
                  const double ji[4] = {J[3] / det, -J[1] / det, -J[2] / det, J[0] / det};
                  for (int k = 0; k < DIM; ++k) nds[k] = det * ji[axis * DIM + k] * (side ? 1.0 : -1.0);
                }
              else
                {
                  // cofactor row `axis` of J = det * Jinv[axis][:]
                  const int r1 = (axis + 1) % 3, r2 = (axis + 2) % 3;
                  // det * Jinv[axis][k] = cofactor(J)[k][axis]
                  for (int k = 0; k < 3; ++k)
                    {
                      const int k1 = (k + 1) % 3, k2 = (k + 2) % 3;
                      nds[k] = (J[k1 * 3 + r1] * J[k2 * 3 + r2] - J[k1 * 3 + r2] * J[k2 * 3 + r1]) * (side ? 1.0 : -1.0);
                    }
                  det = 1.0;
                }
              double dS = 0.0;
              for (int k = 0; k < DIM; ++k) dS += nds[k] * nds[k];
              dS = sqrt(dS);
              double tr;
              if (fsi_rows)
                {
                  tr = 0.0;
                  for (int d2 = 0; d2 < DIM; ++d2)
                    {
                      double sg = 0.0; // sigma[c][d2] at q
                      for (int b = 0; b < NPC; ++b)
                        sg = fma(Nf[fq * NPC + b], fsi_rows[(int64_t)c * n_dofs + (int64_t)DIM * cell_nodes[(int64_t)cell * NPC + b] + d2], sg);
                      tr = fma(sg, nds[d2] / dS, tr);
                    }
                }
              else
                tr = is_pressure ? nds[c] / dS * vals[f * DIM] : vals[f * DIM + c];
              r = fma(Nf[fq * NPC + a] * tr, dS * qwf[q], r);
            }
          const int64_t g = (int64_t)DIM * cell_nodes[(int64_t)cell * NPC + a] + c;
          if (!con[g] && r != 0.0) atomicAdd(&rhs[g], r);
        }
    }

    // one thread per cell (cells of one colour per launch): quadrature values -> qpt_to_dof -> nodal scatter-add
    template <int DIM, int NPC>
    __global__ void solid_stress_kernel(int n_list, const int *__restrict__ cell_list, int nq, const int *__restrict__ cell_nodes,
                                        const double *__restrict__ qpt_to_dof, const double *__restrict__ Finv,
                                        const double *__restrict__ tau, const double *__restrict__ detF, int n_nodes,
                                        double *__restrict__ stress, double *__restrict__ strain, double *__restrict__ count)
    {
      const int li = blockIdx.x * blockDim.x + threadIdx.x;
      if (li >= n_list) return;
      const int cell = cell_list[li];
      for (int a = 0; a < NPC; ++a)
        {
          const int node = cell_nodes[(int64_t)cell * NPC + a];
          double st[DIM * DIM], sn[DIM * DIM];
#pragma unroll
          for (int i = 0; i < DIM * DIM; ++i) st[i] = sn[i] = 0.0;
          for (int q = 0; q < nq; ++q)
            {
              const int64_t cq = (int64_t)cell * nq + q;
              const double w = qpt_to_dof[a * nq + q], J = detF[cq];
              // F = (F^-1)^-1
              const double *fi = Finv + cq * DIM * DIM;
              double F[DIM * DIM];
              if (DIM == 2)
                {
                  const double d = 1.0 / (fi[0] * fi[3] - fi[1] * fi[2]);
                  F[0] = fi[3] * d; F[1] = -fi[1] * d; F[2] = -fi[2] * d; F[3] = fi[0] * d;
                }
              else
                {
                  const double c00 = fi[4] * fi[8] - fi[5] * fi[7], c01 = fi[5] * fi[6] - fi[3] * fi[8], c02 = fi[3] * fi[7] - fi[4] * fi[6];
                  const double d = 1.0 / (fi[0] * c00 + fi[1] * c01 + fi[2] * c02);
                  F[0] = c00 * d; F[1] = (fi[2] * fi[7] - fi[1] * fi[8]) * d; F[2] = (fi[1] * fi[5] - fi[2] * fi[4]) * d;
                  F[3] = c01 * d; F[4] = (fi[0] * fi[8] - fi[2] * fi[6]) * d; F[5] = (fi[2] * fi[3] - fi[0] * fi[5]) * d;
                  F[6] = c02 * d; F[7] = (fi[1] * fi[6] - fi[0] * fi[7]) * d; F[8] = (fi[0] * fi[4] - fi[1] * fi[3]) * d;
                }
#pragma unroll
              for (int i = 0; i < DIM * DIM; ++i)
                {
                  st[i] = fma(w, tau[cq * DIM * DIM + i] / J, st[i]);
                  sn[i] = fma(w, F[i], sn[i]);
                }
            }
#pragma unroll
          for (int i = 0; i < DIM * DIM; ++i)
            {
              stress[(int64_t)i * n_nodes + node] += st[i];
              strain[(int64_t)i * n_nodes + node] += sn[i];
            }
          count[node] += 1.0;
        }
    }

    __global__ void solid_average_kernel(int n_nodes, int ncomp, const double *__restrict__ count, double *__restrict__ a, double *__restrict__ b)
    {
      const int i = blockIdx.x * blockDim.x + threadIdx.x;
      if (i >= n_nodes) return;
      const double c = count[i];
      if (c > 0)
        for (int k = 0; k < ncomp; ++k)
          {
            a[(int64_t)k * n_nodes + i] /= c;
            b[(int64_t)k * n_nodes + i] /= c;
          }
    }

    // ---- LinearElasticity / SharedLinearElasticity: bodies in solid_linear.cuh -----------------------------------
    template <int DIM, int NPC>
    __global__ void __launch_bounds__(64) linear_assemble_kernel(const LinearArgs A)
    {
      linear_assemble_body<DIM, NPC>(A, blockIdx.x, threadIdx.x);
    }

    template <int DIM, int NPC>
    __global__ void linear_stress_kernel(int n_list, const int *__restrict__ cell_list, int nq, const int *__restrict__ cell_nodes,
                                         const double *__restrict__ qpt_to_dof, const double *__restrict__ G,
                                         const double *__restrict__ u, double lambda, double mu, int n_nodes,
                                         double *__restrict__ stress, double *__restrict__ strain, double *__restrict__ count)
    {
      linear_stress_body<DIM, NPC>(blockIdx.x * blockDim.x + threadIdx.x, n_list, cell_list, nq, cell_nodes, qpt_to_dof, G, u, lambda, mu,
                                   n_nodes, stress, strain, count);
    }

    struct ScopedTimer
    {
      Context &ctx;
      double &acc;
      std::chrono::steady_clock::time_point t0;
      ScopedTimer(Context &c, double &a) : ctx(c), acc(a)
      {
        cudaStreamSynchronize(ctx.stream);
        t0 = std::chrono::steady_clock::now();
      }
      ~ScopedTimer()
      {
        cudaStreamSynchronize(ctx.stream);
        acc += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
      }
    };
  } // namespace

  // ===========================================================================
  void SolidSpace::setup(Context &ctx, const Triangulation &tria, const Parameters::AllParameters &prm)
  {
    dim = tria.dim;
    degree = (int)prm.solid_degree;
    if (degree != 1) throw std::runtime_error("SolidSpace: only FE_Q(1) solids are implemented on the device");
    n_cells = tria.n_cells();
    nv = 1 << dim;
    nsym = dim == 2 ? 3 : 6;
    FEQ fe(dim, degree), feg(dim, 1);
    npc = fe.n;
    Quadrature quad(dim, degree + 1);
    nq = quad.nq;
    ShapeTable tab(fe, quad.points, nq), tabg(feg, quad.points, nq);
    nt = build_node_table(tria, degree);
    n_dofs = (int64_t)dim * nt.n_nodes;
    P = build_pattern(n_cells, nt.cell_nodes.data(), npc, nt.n_nodes, nt.cell_nodes.data(), npc, nt.n_nodes);
    colour_cells(n_cells, nt.cell_nodes.data(), npc, nt.n_nodes, colour_order, colour_offsets);
    // geometry on the reference (undeformed) configuration: G[c][q][a][k], JxW[c][q]
    std::vector<double> G((size_t)n_cells * nq * npc * dim), JxW((size_t)n_cells * nq);
#pragma omp parallel for schedule(static)
    for (int c = 0; c < n_cells; ++c)
      for (int q = 0; q < nq; ++q)
        {
          double J[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, Ji[9];
          for (int v = 0; v < nv; ++v)
            {
              const double *X = &tria.vertices[(size_t)tria.cells[(size_t)c * nv + v] * dim];
              for (int i = 0; i < dim; ++i)
                for (int j = 0; j < dim; ++j) J[i * dim + j] += X[i] * tabg.dN[((size_t)q * nv + v) * dim + j];
            }
          double det;
          if (dim == 2)
            {
              det = J[0] * J[3] - J[1] * J[2];
              Ji[0] = J[3] / det; Ji[1] = -J[1] / det; Ji[2] = -J[2] / det; Ji[3] = J[0] / det;
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
          JxW[(size_t)c * nq + q] = det * quad.weights[q];
          for (int a = 0; a < npc; ++a)
            for (int k = 0; k < dim; ++k)
              {
                double s = 0;
                for (int j = 0; j < dim; ++j) s += tab.dN[((size_t)q * npc + a) * dim + j] * Ji[j * dim + k];
                G[(((size_t)c * nq + q) * npc + a) * dim + k] = s;
              }
        }
    // homogeneous Dirichlet constraints (mpi_solid_solver.cpp:67-94)
    con.assign(n_dofs, 0);
    for (const auto &bc : prm.solid_dirichlet_bcs)
      for (int f = 0; f < tria.n_boundary_faces(); ++f)
        {
          if (tria.boundary_faces[3 * f + 2] != (int)bc.first) continue;
          const int cell = tria.boundary_faces[3 * f], face = tria.boundary_faces[3 * f + 1];
          for (int a : face_local_nodes(dim, degree, face))
            for (int c = 0; c < dim; ++c)
              if (bc.second & (1u << c)) con[(size_t)dim * nt.cell_nodes[(size_t)cell * npc + a] + c] = 1;
        }
    std::vector<int> idx;
    for (int64_t g = 0; g < n_dofs; ++g)
      if (con[g]) idx.push_back((int)g);
    n_con = (int)idx.size();
    // Neumann faces
    std::vector<int> nf;
    std::vector<double> nfv;
    neumann_is_pressure = prm.solid_neumann_bc_type == "Pressure";
    if (prm.simulation_type == "FSI")
      {
        // every boundary face carries the fluid traction (rows of constrained dofs are dropped in the scatter)
        for (int f = 0; f < tria.n_boundary_faces(); ++f)
          {
            nf.push_back(tria.boundary_faces[3 * f]);
            nf.push_back(tria.boundary_faces[3 * f + 1]);
            for (int c = 0; c < dim; ++c) nfv.push_back(0.0);
          }
      }
    else
      for (int f = 0; f < tria.n_boundary_faces(); ++f)
        {
          const unsigned id = (unsigned)tria.boundary_faces[3 * f + 2];
          if (neumann_skips_dirichlet_faces && prm.solid_dirichlet_bcs.count(id)) continue;
          auto it = prm.solid_neumann_bcs.find(id);
          if (it == prm.solid_neumann_bcs.end()) continue;
          nf.push_back(tria.boundary_faces[3 * f]);
          nf.push_back(tria.boundary_faces[3 * f + 1]);
          for (int c = 0; c < dim; ++c) nfv.push_back(c < (int)it->second.size() ? it->second[c] : 0.0);
        }
    n_nfaces = (int)nf.size() / 2;

    cudaStream_t s = ctx.stream;
    d_cell_nodes.upload(nt.cell_nodes, s);
    d_colour_order.upload(colour_order, s);
    d_con.upload(con, s);
    if (n_con) d_con_idx.upload(idx, s);
    d_N.upload(tab.N, s);
    d_G.upload(G, s);
    d_JxW.upload(JxW, s);
    d_node_x.upload(nt.coords, s);
    const size_t nqp = (size_t)n_cells * nq;
    d_Finv.alloc(nqp * dim * dim);
    d_tau.alloc(nqp * dim * dim);
    d_Jc.alloc(nqp * nsym * nsym);
    d_detF.alloc(nqp);
    if (n_nfaces)
      {
        Quadrature fq(dim - 1, degree + 1);
        nqf = fq.nq;
        std::vector<double> t, Nf((size_t)2 * dim * nqf * npc), Gf((size_t)2 * dim * nqf * nv * dim);
        std::vector<double> N(npc), dN((size_t)npc * dim), g(nv), dg((size_t)nv * dim);
        for (int face = 0; face < 2 * dim; ++face)
          for (int q = 0; q < nqf; ++q)
            {
              double xi[3];
              int k = 0;
              for (int d = 0; d < dim; ++d) xi[d] = (d == face / 2) ? double(face % 2) : fq.points[(size_t)q * (dim - 1) + k++];
              fe.eval(xi, N.data(), dN.data());
              feg.eval(xi, g.data(), dg.data());
              std::copy(N.begin(), N.end(), Nf.begin() + ((size_t)face * nqf + q) * npc);
              std::copy(dg.begin(), dg.end(), Gf.begin() + ((size_t)face * nqf + q) * nv * dim);
            }
        t.insert(t.end(), Nf.begin(), Nf.end());
        t.insert(t.end(), Gf.begin(), Gf.end());
        t.insert(t.end(), fq.weights.begin(), fq.weights.end());
        d_face_tables.upload(t, s);
        d_nface.upload(nf, s);
        d_nface_val.upload(nfv, s);
      }
    K.init(P, dim, dim, s);
    M.init(P, dim, dim, s);
    rhs.alloc(n_dofs);
    d_slots.alloc((size_t)n_cells * npc * npc);
    DevBuf<int> err(1);
    err.zero(s);
    const int64_t total = (int64_t)n_cells * npc * npc;
    solid_slots_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(n_cells, npc, d_cell_nodes.p, K.rowptr.p, K.col.p, d_slots.p, err.p);
    IFEM_KERNEL_CHECK();
    if (err.to_host(s)[0]) throw std::runtime_error("SolidSpace::setup: a matrix row has more than 256 block columns");
  }

  // ===========================================================================
  SolidSolver::SolidSolver(Context &ctx_, Triangulation &tria, const Parameters::AllParameters &params)
    : ctx(ctx_), triangulation(tria), parameters(params),
      time(params.end_time, params.time_step, params.output_interval, params.refinement_interval, params.save_interval)
  {
  }

  void SolidSolver::setup_dofs()
  {
    ss.setup(ctx, triangulation, parameters);
    dofs_ready = true;
  }

  void SolidSolver::initialize_system()
  {
    for (DevBuf<double> *v : {&current_displacement, &current_velocity, &current_acceleration, &previous_displacement,
                              &previous_velocity, &previous_acceleration, &d_tmp, &d_pred, &d_update})
      {
        v->alloc(ss.n_dofs);
        v->zero(ctx.stream);
      }
    d_binv.alloc((size_t)ss.nt.n_nodes * ss.dim * ss.dim);
    stress.alloc((size_t)ss.dim * ss.dim * ss.nt.n_nodes);
    strain.alloc((size_t)ss.dim * ss.dim * ss.nt.n_nodes);
    stress.zero(ctx.stream);
    strain.zero(ctx.stream);
    fsi_stress_rows.alloc((size_t)ss.dim * ss.n_dofs);
    fluid_velocity.alloc(ss.n_dofs);
    fluid_pressure.alloc(ss.nt.n_nodes);
    fsi_stress_rows.zero(ctx.stream);
    fluid_velocity.zero(ctx.stream);
    fluid_pressure.zero(ctx.stream);
    d_count.alloc(ss.nt.n_nodes);
    {
      // qpt_to_dof = M^-1 Q^T W on the reference cell (FETools::compute_projection_from_quadrature_points_matrix)
      FEQ fe(ss.dim, ss.degree);
      Quadrature quad(ss.dim, ss.degree + 1);
      ShapeTable tab(fe, quad.points, quad.nq);
      const int n = fe.n, nq = quad.nq;
      std::vector<double> M((size_t)n * n, 0.0), R((size_t)n * nq, 0.0);
      for (int q = 0; q < nq; ++q)
        for (int i = 0; i < n; ++i)
          {
            R[(size_t)i * nq + q] = tab.N[(size_t)q * n + i] * quad.weights[q];
            for (int j = 0; j < n; ++j) M[(size_t)i * n + j] += tab.N[(size_t)q * n + i] * tab.N[(size_t)q * n + j] * quad.weights[q];
          }
      for (int c = 0; c < n; ++c)
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
      d_qpt_to_dof.upload(R, ctx.stream);
    }
    // initial velocity (mpi_solid_solver.cpp:116-137, mpi_shared_solid_solver.cpp:152-196), constraints distributed
    bool any = false;
    for (int c = 0; c < ss.dim && c < (int)parameters.initial_velocity.size(); ++c) any = any || parameters.initial_velocity[c] != 0.0;
    if (any)
      {
        std::vector<double> v(ss.n_dofs, 0.0);
        for (int64_t g = 0; g < ss.n_dofs; ++g)
          if (!ss.con[g] && (int)(g % ss.dim) < (int)parameters.initial_velocity.size()) v[g] = parameters.initial_velocity[g % ss.dim];
        previous_velocity.upload(v, ctx.stream);
        current_velocity.upload(v, ctx.stream);
      }
    IFEM_CUDA(cudaStreamSynchronize(ctx.stream));
  }

  void SolidSolver::neumann_rhs()
  {
    cudaStream_t s = ctx.stream;
    if (ss.n_nfaces)
      {
        const bool fsi = parameters.simulation_type == "FSI";
        const double *rows = fsi ? fsi_stress_rows.p : nullptr, *disp = fsi ? current_displacement.p : nullptr;
        if (ss.dim == 2)
          solid_neumann_kernel<2, 4><<<ss.n_nfaces, 32, 0, s>>>(ss.n_nfaces, ss.nqf, ss.d_nface.p, ss.d_nface_val.p, ss.neumann_is_pressure,
                                                                ss.d_face_tables.p, ss.d_cell_nodes.p, ss.d_node_x.p, ss.d_con.p, ss.rhs.p,
                                                                rows, disp, ss.n_dofs);
        else
          solid_neumann_kernel<3, 8><<<ss.n_nfaces, 32, 0, s>>>(ss.n_nfaces, ss.nqf, ss.d_nface.p, ss.d_nface_val.p, ss.neumann_is_pressure,
                                                                ss.d_face_tables.p, ss.d_cell_nodes.p, ss.d_node_x.p, ss.d_con.p, ss.rhs.p,
                                                                rows, disp, ss.n_dofs);
        IFEM_KERNEL_CHECK();
        ctx.kernel_launches++;
      }
  }

  // SolidSolver::solve (mpi_solid_solver.cpp:143-161): CG to 1e-8 |b|; the reference's per-rank ILU(0) block
  // Jacobi is replaced by the node-block Jacobi preconditioner (rank-count independent).
  std::pair<unsigned int, double> SolidSolver::solve(Bcsr &A, double *x, const double *b)
  {
    ScopedTimer t(ctx, timer_ms["Solve linear system"]);
    const VecSpace n(ss.n_dofs);
    block_diag_inverse(ctx, A, d_binv.p);
    LinOp op = [&](const double *v, double *y) { spmv(ctx, A, v, y); };
    LinOp pc = [&](const double *v, double *y) { block_diag_apply(ctx, ss.nt.n_nodes, ss.dim, d_binv.p, v, y); };
    const double tol = 1e-8 * nrm2(ctx, n, b);
    const SolveResult r = pcg(ctx, n, op, pc, b, x, tol, (int)ss.n_dofs, pool);
    if (ss.n_con) set_indexed(ctx, ss.n_con, ss.d_con_idx.p, nullptr, x); // constraints.distribute (homogeneous)
    return {(unsigned)r.iterations, r.residual};
  }

  double SolidSolver::get_error(const double *v)
  {
    const VecSpace n(ss.n_dofs);
    copy(ctx, n, v, d_tmp.p);
    if (ss.n_con) set_indexed(ctx, ss.n_con, ss.d_con_idx.p, nullptr, d_tmp.p);
    return nrm2(ctx, n, d_tmp.p);
  }

  void SolidSolver::run()
  {
    if (!dofs_ready) triangulation.refine_global(parameters.global_refinements.size() > 1 ? parameters.global_refinements[1] : 0);
    const bool success_load = load_checkpoint(); // mpi_solid_solver.cpp:318-319; false unless an output directory is set
    if (!dofs_ready)
      {
        setup_dofs();
        initialize_system();
      }
    if (!success_load)
      run_one_step(true);
    else
      after_restart(); // "if we load from previous task, we need to assemble the mass matrix" (mpi_shared_solid_solver.cpp:433-437)
    while (time.end() - time.current() > 1e-12) run_one_step(false);
  }

  std::vector<double> SolidSolver::get_current_solution() { return current_displacement.to_host(ctx.stream); }

  // ===========================================================================
  HyperElasticity::HyperElasticity(Context &ctx_, Triangulation &tria, const Parameters::AllParameters &params, int shared)
    : SolidSolver(ctx_, tria, params)
  {
    shared_twin = shared < 0 ? params.simulation_type == "FSI" : shared != 0;
    // PointHistory::setup (mpi_hyper_elasticity.cpp:8-35): NeoHookean or Kirchhoff
    if (parameters.solid_type == "Kirchhoff")
      {
        if (parameters.E.empty() || parameters.nu.empty()) throw std::runtime_error("HyperElasticity: Kirchhoff requires Young's modulus and Poisson's ratio");
      }
    else if (parameters.solid_type != "NeoHookean")
      throw std::runtime_error("HyperElasticity: Solid type must be NeoHookean or Kirchhoff");
    else if (parameters.C.empty() || parameters.C[0].size() < 2)
      throw std::runtime_error("HyperElasticity: NeoHookean requires C1, kappa");
  }

  void HyperElasticity::initialize_system()
  {
    SolidSolver::initialize_system();
    // setup_qph (:217-239): PointHistory::setup calls update with a zero displacement gradient
    update_qph(current_displacement.p);
    IFEM_CUDA(cudaStreamSynchronize(ctx.stream));
  }

  void HyperElasticity::update_qph(const double *u)
  {
    ScopedTimer t(ctx, timer_ms["Update QPH data"]);
    const int total = ss.n_cells * ss.nq;
    // NeoHookean: (C1, kappa) of "Hyperelastic parameters"; Kirchhoff: (Young's modulus, Poisson's ratio) (mpi_hyper_elasticity.cpp:13-27)
    const bool kirchhoff = parameters.solid_type == "Kirchhoff";
    const int material = kirchhoff ? 1 : 0;
    if ((int)d_cell_mat.n != 2 * ss.n_cells)
      {
        // (C1, kappa) or (E, nu) of every cell from its material id; one part: id 1 for all cells (:226-228)
        std::vector<double> cm((size_t)2 * ss.n_cells);
        const size_t n_parts = kirchhoff ? std::min(parameters.E.size(), parameters.nu.size()) : parameters.C.size();
        for (int c = 0; c < ss.n_cells; ++c)
          {
            unsigned int mat_id = parameters.n_solid_parts == 1 ? 1u : (unsigned int)triangulation.material_id[c];
            if (mat_id < 1 || mat_id > n_parts) throw std::runtime_error("HyperElasticity: no material parameters for material id " + std::to_string(mat_id));
            if (!kirchhoff && parameters.C[mat_id - 1].size() < 2) throw std::runtime_error("HyperElasticity: NeoHookean needs two parameters per part");
            cm[2 * c] = kirchhoff ? parameters.E[mat_id - 1] : parameters.C[mat_id - 1][0];
            cm[2 * c + 1] = kirchhoff ? parameters.nu[mat_id - 1] : parameters.C[mat_id - 1][1];
          }
        d_cell_mat.upload(cm, ctx.stream);
      }
    if (ss.dim == 2)
      update_qph_kernel<2, 4><<<(total + 127) / 128, 128, 0, ctx.stream>>>(ss.n_cells, ss.nq, ss.d_cell_nodes.p, ss.d_G.p, u, material, d_cell_mat.p,
                                                                           ss.d_Finv.p, ss.d_tau.p, ss.d_Jc.p, ss.d_detF.p);
    else
      update_qph_kernel<3, 8><<<(total + 127) / 128, 128, 0, ctx.stream>>>(ss.n_cells, ss.nq, ss.d_cell_nodes.p, ss.d_G.p, u, material, d_cell_mat.p,
                                                                           ss.d_Finv.p, ss.d_tau.p, ss.d_Jc.p, ss.d_detF.p);
    IFEM_KERNEL_CHECK();
    ctx.kernel_launches++;
  }

  void HyperElasticity::update_strain_and_stress()
  {
    cudaStream_t s = ctx.stream;
    stress.zero(s);
    strain.zero(s);
    d_count.zero(s);
    const int n_colours = (int)ss.colour_offsets.size() - 1;
    for (int k = 0; k < n_colours; ++k)
      {
        const int n = ss.colour_offsets[k + 1] - ss.colour_offsets[k];
        if (!n) continue;
        const int *list = ss.d_colour_order.p + ss.colour_offsets[k];
        if (ss.dim == 2)
          solid_stress_kernel<2, 4><<<(n + 127) / 128, 128, 0, s>>>(n, list, ss.nq, ss.d_cell_nodes.p, d_qpt_to_dof.p, ss.d_Finv.p, ss.d_tau.p,
                                                                    ss.d_detF.p, ss.nt.n_nodes, stress.p, strain.p, d_count.p);
        else
          solid_stress_kernel<3, 8><<<(n + 127) / 128, 128, 0, s>>>(n, list, ss.nq, ss.d_cell_nodes.p, d_qpt_to_dof.p, ss.d_Finv.p, ss.d_tau.p,
                                                                    ss.d_detF.p, ss.nt.n_nodes, stress.p, strain.p, d_count.p);
        IFEM_KERNEL_CHECK();
        ctx.kernel_launches++;
      }
    solid_average_kernel<<<(ss.nt.n_nodes + 255) / 256, 256, 0, s>>>(ss.nt.n_nodes, ss.dim * ss.dim, d_count.p, stress.p, strain.p);
    IFEM_KERNEL_CHECK();
    ctx.kernel_launches++;
  }

  void HyperElasticity::assemble_system(bool initial_step)
  {
    ScopedTimer t(ctx, timer_ms["Assemble tangent matrix"]);
    cudaStream_t s = ctx.stream;
    Bcsr &A = initial_step ? ss.M : ss.K;
    A.zero(s);
    ss.rhs.zero(s);
    const double gamma = 0.5 + parameters.damping, beta = gamma / 2, dt = time.get_delta_t();
    HyperArgs a;
    a.cell_nodes = ss.d_cell_nodes.p;
    a.slots = ss.d_slots.p;
    a.con = ss.d_con.p;
    a.N = ss.d_N.p;
    a.G = ss.d_G.p;
    a.JxW = ss.d_JxW.p;
    a.Finv = ss.d_Finv.p;
    a.tau = ss.d_tau.p;
    a.Jc = ss.d_Jc.p;
    a.nq = ss.nq;
    a.rho = parameters.solid_rho;
    a.inv_beta_dt2 = 1.0 / (beta * dt * dt);
    for (int d = 0; d < 3; ++d) a.grav[d] = d < (int)parameters.gravity.size() ? parameters.gravity[d] : 0.0;
    a.initial_step = initial_step ? 1 : 0;
    a.rowptr = A.rowptr.p;
    a.val = A.val.p;
    a.rhs = ss.rhs.p;
    const int n_colours = (int)ss.colour_offsets.size() - 1;
    for (int k = 0; k < n_colours; ++k)
      {
        a.n_list = ss.colour_offsets[k + 1] - ss.colour_offsets[k];
        a.cell_list = ss.d_colour_order.p + ss.colour_offsets[k];
        if (!a.n_list) continue;
        if (ss.dim == 2)
          hyper_assemble_kernel<2, 4><<<(a.n_list + 3) / 4, 64, 0, s>>>(a);
        else
          hyper_assemble_kernel<3, 8><<<a.n_list, 64, 0, s>>>(a);
        IFEM_KERNEL_CHECK();
        ctx.kernel_launches++;
      }
    neumann_rhs();
  }

  void HyperElasticity::run_one_step(bool first_step)
  {
    io_before_step();
    const VecSpace n(ss.n_dofs);
    const double gamma = 0.5 + parameters.damping, beta = gamma / 2;
    if (first_step)
      {
        assemble_system(true);
        solve(ss.M, previous_acceleration.p, ss.rhs.p);
      }
    time.increment();
    const double dt = time.get_delta_t();
    fill(ctx, n, 0.0, d_update.p);
    double err_res = 1.0, err_res0 = 1.0, nerr_res = 1.0, err_upd = 1.0, err_upd0 = 1.0, nerr_upd = 1.0;
    unsigned int it = 0;
    // predicted = previous_u + dt v + (0.5 - beta) dt^2 a
    lin3(ctx, n, d_pred.p, previous_displacement.p, dt, previous_velocity.p, (0.5 - beta) * dt * dt, previous_acceleration.p);
    auto kinematics = [&] {
      // a = (u - predicted) / (beta dt^2);  v = v_prev + dt (1 - gamma) a_prev + dt gamma a
      lin3(ctx, n, current_acceleration.p, current_displacement.p, -1.0, d_pred.p, 0.0, d_pred.p);
      scale(ctx, n, 1.0 / (beta * dt * dt), current_acceleration.p);
      lin3(ctx, n, current_velocity.p, previous_velocity.p, dt * (1 - gamma), previous_acceleration.p, dt * gamma, current_acceleration.p);
    };
    // the replicated twin MPI::FSI uses also stops on a vanishing update (mpi_shared_hyper_elasticity.cpp:125-127)
    while ((nerr_upd > parameters.tol_d || nerr_res > parameters.tol_f) && (!shared_twin || err_upd > 1e-12))
      {
        if (it >= parameters.solid_max_iterations) throw std::runtime_error("Too many Newton iterations!");
        kinematics();
        assemble_system(false);
        spmv(ctx, ss.M, current_acceleration.p, d_tmp.p);
        axpy(ctx, n, -1.0, d_tmp.p, ss.rhs.p);
        const auto lin = solve(ss.K, d_update.p, ss.rhs.p);
        err_res = get_error(ss.rhs.p);
        if (it == 0) err_res0 = err_res;
        nerr_res = err_res / err_res0;
        err_upd = get_error(d_update.p);
        if (it == 0) err_upd0 = err_upd;
        nerr_upd = err_upd / err_upd0;
        axpy(ctx, n, 1.0, d_update.p, current_displacement.p);
        update_qph(current_displacement.p);
        history.push_back({time.get_timestep(), it, err_res, err_upd, (int)lin.first});
        if (verbose)
          std::printf("Newton iteration = %u, CG itr = %u, CG res = %.3e, res_F = %.3e, res_U = %.3e\n", it, lin.first, lin.second,
                      err_res, err_upd);
        it++;
      }
    kinematics();
    copy(ctx, n, current_acceleration.p, previous_acceleration.p);
    copy(ctx, n, current_velocity.p, previous_velocity.p);
    copy(ctx, n, current_displacement.p, previous_displacement.p);
    if (shared_twin) update_strain_and_stress(); // the twin used by MPI::FSI does this every step (mpi_shared_hyper_elasticity.cpp:204-205)
    io_after_step();
  }

  // ===========================================================================
  LinearElasticity::LinearElasticity(Context &ctx_, Triangulation &tria, const Parameters::AllParameters &params, bool shared_)
    : SolidSolver(ctx_, tria, params), shared(shared_)
  {
    if (parameters.solid_type != "LinearElastic") throw std::runtime_error("LinearElasticity: Solid type must be LinearElastic");
    if (parameters.E.empty() || parameters.nu.empty()) throw std::runtime_error("LinearElasticity: Young's modulus and Poisson's ratio are required");
    if (parameters.n_solid_parts != 1) throw std::runtime_error("LinearElasticity: one solid part only on the device");
    // LinearElasticMaterial (linear_elastic_material.cpp:5-14)
    const double E = parameters.E[0], nu = parameters.nu[0];
    lambda = E * nu / ((1 + nu) * (1 - 2 * nu));
    mu = E / (2 * (1 + nu));
    eta = parameters.eta.empty() ? 0.0 : parameters.eta[0];
    // the linear solvers integrate every face with a Neumann id (mpi_linear_elasticity.cpp:139-142); the hyperelastic one
    // skips faces that also carry a Dirichlet id (mpi_hyper_elasticity.cpp:452-456)
    ss.neumann_skips_dirichlet_faces = false;
  }

  void LinearElasticity::initialize_system()
  {
    SolidSolver::initialize_system();
    stiffness_matrix.init(ss.P, ss.dim, ss.dim, ctx.stream);
    if (shared) damping_matrix.init(ss.P, ss.dim, ss.dim, ctx.stream);
    d_tmp2.alloc(ss.n_dofs);
    d_tmp3.alloc(ss.n_dofs);
    IFEM_CUDA(cudaStreamSynchronize(ctx.stream));
  }

  void LinearElasticity::assemble_system(bool is_initial)
  {
    ScopedTimer t(ctx, timer_ms["Assemble system"]);
    cudaStream_t s = ctx.stream;
    const double dt = time.get_delta_t();
    LinearArgs a;
    a.cell_nodes = ss.d_cell_nodes.p;
    a.slots = ss.d_slots.p;
    a.con = ss.d_con.p;
    a.N = ss.d_N.p;
    a.G = ss.d_G.p;
    a.JxW = ss.d_JxW.p;
    a.nq = ss.nq;
    a.rho = parameters.solid_rho;
    a.lambda = lambda;
    a.mu = mu;
    a.eta = eta;
    for (int d = 0; d < 3; ++d) a.grav[d] = d < (int)parameters.gravity.size() ? parameters.gravity[d] : 0.0;
    a.rowptr = ss.K.rowptr.p;
    a.sys = a.mass = a.stiff = a.damp = nullptr;
    a.c_mass = 1.0;
    a.c_damp = a.c_stiff = 0.0;
    a.rhs = ss.rhs.p;
    ss.rhs.zero(s);
    if (!shared)
      {
        // mpi_linear_elasticity.cpp:31-36, 96-121: is_initial -> system_matrix = mass; else system = M + beta dt^2 K and K
        const double gamma = 0.5 + parameters.damping, beta = gamma / 2;
        ss.K.zero(s);
        stiffness_matrix.zero(s);
        a.sys = ss.K.val.p;
        if (!is_initial)
          {
            a.stiff = stiffness_matrix.val.p;
            a.c_stiff = beta * dt * dt;
          }
      }
    else if (is_initial)
      {
        // mpi_shared_linear_elasticity.cpp:30-40, 126-149: the four matrices are assembled once; beta = (1 + alpha)^2 / 4 HERE
        const double alpha = -parameters.damping, gamma = 0.5 - alpha, beta = (1 + alpha) * (1 + alpha) / 4;
        ss.K.zero(s);
        ss.M.zero(s);
        stiffness_matrix.zero(s);
        damping_matrix.zero(s);
        a.sys = ss.K.val.p;
        a.mass = ss.M.val.p;
        a.stiff = stiffness_matrix.val.p;
        a.damp = damping_matrix.val.p;
        a.c_damp = gamma * dt * (1 + alpha);
        a.c_stiff = beta * dt * dt * (1 + alpha);
      }
    const int n_colours = (int)ss.colour_offsets.size() - 1;
    for (int k = 0; k < n_colours; ++k)
      {
        a.n_list = ss.colour_offsets[k + 1] - ss.colour_offsets[k];
        a.cell_list = ss.d_colour_order.p + ss.colour_offsets[k];
        if (!a.n_list) continue;
        if (ss.dim == 2)
          linear_assemble_kernel<2, 4><<<(a.n_list + 3) / 4, 64, 0, s>>>(a);
        else
          linear_assemble_kernel<3, 8><<<a.n_list, 64, 0, s>>>(a);
        IFEM_KERNEL_CHECK();
        ctx.kernel_launches++;
      }
    neumann_rhs();
  }

  void LinearElasticity::run_one_step(bool first_step)
  {
    io_before_step();
    const VecSpace n(ss.n_dofs);
    const double dt = time.get_delta_t();
    double gamma, beta;
    if (!shared)
      {
        // mpi_linear_elasticity.cpp:199-262
        gamma = 0.5 + parameters.damping;
        beta = gamma / 2;
        if (first_step)
          {
            assemble_system(true);
            solve(ss.K, previous_acceleration.p, ss.rhs.p); // M a_0 = F
            assemble_system(false);
          }
        time.increment();
        // tmp1 = rhs - K (u_n + dt v_n + (1/2 - beta) dt^2 a_n)
        lin3(ctx, n, d_tmp2.p, previous_displacement.p, dt, previous_velocity.p, (0.5 - beta) * dt * dt, previous_acceleration.p);
        spmv(ctx, stiffness_matrix, d_tmp2.p, d_tmp3.p);
        copy(ctx, n, ss.rhs.p, d_tmp.p);
        axpy(ctx, n, -1.0, d_tmp3.p, d_tmp.p);
      }
    else
      {
        // mpi_shared_linear_elasticity.cpp:300-348; beta = (1 - alpha)^2 / 4 HERE (the reference's own inconsistency, kept)
        const double alpha = -parameters.damping;
        gamma = 0.5 - alpha;
        beta = (1 - alpha) * (1 - alpha) / 4;
        if (first_step)
          {
            assemble_system(true);
            solve(ss.M, previous_acceleration.p, ss.rhs.p);
          }
        else if (parameters.simulation_type == "FSI")
          assemble_system(false);
        time.increment();
        lin3(ctx, n, d_tmp2.p, previous_displacement.p, (1 + alpha) * dt, previous_velocity.p, (0.5 - beta) * dt * dt * (1 + alpha),
             previous_acceleration.p);
        spmv(ctx, stiffness_matrix, d_tmp2.p, d_tmp3.p);
        copy(ctx, n, ss.rhs.p, d_tmp.p);
        axpy(ctx, n, -1.0, d_tmp3.p, d_tmp.p);
        lin3(ctx, n, d_tmp2.p, previous_velocity.p, (1 + alpha) * (1 - gamma) * dt, previous_acceleration.p, 0.0, previous_acceleration.p);
        spmv(ctx, damping_matrix, d_tmp2.p, d_tmp3.p);
        axpy(ctx, n, -1.0, d_tmp3.p, d_tmp.p);
      }
    const auto lin = solve(ss.K, current_acceleration.p, d_tmp.p);
    // v_{n+1} = v_n + (1 - gamma) dt a_n + gamma dt a_{n+1};  u_{n+1} = u_n + dt v_n + dt^2 ((1/2 - beta) a_n + beta a_{n+1})
    lin3(ctx, n, current_velocity.p, previous_velocity.p, dt * (1 - gamma), previous_acceleration.p, dt * gamma, current_acceleration.p);
    lin3(ctx, n, current_displacement.p, previous_displacement.p, dt, previous_velocity.p, dt * dt * (0.5 - beta), previous_acceleration.p);
    axpy(ctx, n, dt * dt * beta, current_acceleration.p, current_displacement.p);
    copy(ctx, n, current_acceleration.p, previous_acceleration.p);
    copy(ctx, n, current_velocity.p, previous_velocity.p);
    copy(ctx, n, current_displacement.p, previous_displacement.p);
    history.push_back({time.get_timestep(), 0u, lin.second, 0.0, (int)lin.first});
    if (verbose) std::printf(" CG iteration: %u CG residual: %.6e\n", lin.first, lin.second);
    if (shared) update_strain_and_stress();
    io_after_step();
  }

  void LinearElasticity::update_strain_and_stress()
  {
    cudaStream_t s = ctx.stream;
    stress.zero(s);
    strain.zero(s);
    d_count.zero(s);
    const int n_colours = (int)ss.colour_offsets.size() - 1;
    for (int k = 0; k < n_colours; ++k)
      {
        const int n = ss.colour_offsets[k + 1] - ss.colour_offsets[k];
        if (!n) continue;
        const int *list = ss.d_colour_order.p + ss.colour_offsets[k];
        if (ss.dim == 2)
          linear_stress_kernel<2, 4><<<(n + 127) / 128, 128, 0, s>>>(n, list, ss.nq, ss.d_cell_nodes.p, d_qpt_to_dof.p, ss.d_G.p,
                                                                     current_displacement.p, lambda, mu, ss.nt.n_nodes, stress.p, strain.p,
                                                                     d_count.p);
        else
          linear_stress_kernel<3, 8><<<(n + 127) / 128, 128, 0, s>>>(n, list, ss.nq, ss.d_cell_nodes.p, d_qpt_to_dof.p, ss.d_G.p,
                                                                     current_displacement.p, lambda, mu, ss.nt.n_nodes, stress.p, strain.p,
                                                                     d_count.p);
        IFEM_KERNEL_CHECK();
        ctx.kernel_launches++;
      }
    solid_average_kernel<<<(ss.nt.n_nodes + 255) / 256, 256, 0, s>>>(ss.nt.n_nodes, ss.dim * ss.dim, d_count.p, stress.p, strain.p);
    IFEM_KERNEL_CHECK();
    ctx.kernel_launches++;
  }
} // namespace ifem
