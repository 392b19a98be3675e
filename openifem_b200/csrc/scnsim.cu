#include "scnsim.h"

#include <chrono>
#include <cstdio>

#include "comm.h"

namespace ifem
{
  namespace
  {
    constexpr double kCpToCv = 1.4, kAtm = 1013250, kKappaS = 1e4; // mpi_scnsim.cpp:124-126

    template <int DIM>
    struct QPoint
    {
      static constexpr int NU = 1 << DIM;
      double JxW, N[NU], g[NU][DIM];
      double u[DIM], dv[DIM], G[DIM * DIM], p, dp, gradp[DIM], divu, sdiv[DIM];
      double u_gradu[DIM], gradu_u[DIM], res[DIM], g_bf[DIM], acc[DIM], fsis[DIM * DIM];
      double rho, visc, sigma, tau_supg, tau_pspg, tau_lsic;
    };

    struct ScnsArgs
    {
      int n_list;
      const int *cell_list, *cell_un, *cell_pn, *indicator;
      const double *cell_x, *tables; // N[nq][nu] | dN[nq][nu][dim] | Np[nq][np] | dNgeo[nq][nv][dim] | qw[nq]
      const unsigned char *slots, *con;
      const double *eval_pt, *present, *fsi_acc, *stress, *fsi_stress, *sigma_pml, *body_force, *inhom;
      const double *eddy; // nodal eddy viscosity of an attached turbulence model (pressure-node numbering) or null
      int64_t n_u;
      int n_unodes, n_owned_u, n_owned_p, n_h, h_node[8];
      double mu, rho_f, rho_s, dt, grav[3];
      const int64_t *uu_rp, *up_rp, *pu_rp, *pp_rp;
      double *uu, *up, *pu, *pp, *rhs;
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

    template <int DIM>
    __device__ __forceinline__ double dotd(const double *a, const double *b)
    {
      double s = 0.0;
#pragma unroll
      for (int d = 0; d < DIM; ++d) s = fma(a[d], b[d], s);
      return s;
    }

    // Everything that depends on the quadrature point only (mpi_scnsim.cpp:153-289): one thread per (cell, q).
    template <int DIM>
    __device__ void fill_qpoint(const ScnsArgs &A, int cell, int q, int ind, QPoint<DIM> &Q)
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
      invert<DIM>(J, Ji, det);
      Q.JxW = det * tqw[q];
      double up[DIM], pp = 0.0, eddy = 0.0;
#pragma unroll
      for (int c = 0; c < DIM; ++c) Q.u[c] = up[c] = Q.acc[c] = Q.gradp[c] = Q.sdiv[c] = 0.0;
#pragma unroll
      for (int i = 0; i < DIM * DIM; ++i) Q.G[i] = Q.fsis[i] = 0.0;
      Q.p = 0.0;
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
          const double pe = A.eval_pt[A.n_u + pn];
          Q.p = fma(N, pe, Q.p);
          pp = fma(N, A.present[A.n_u + pn], pp);
          if (A.eddy) eddy = fma(N, A.eddy[pn], eddy);
#pragma unroll
          for (int c = 0; c < DIM; ++c)
            {
              const double ue = A.eval_pt[(int64_t)DIM * un + c];
              Q.u[c] = fma(N, ue, Q.u[c]);
              up[c] = fma(N, A.present[(int64_t)DIM * un + c], up[c]);
              if (A.fsi_acc) Q.acc[c] = fma(N, A.fsi_acc[(int64_t)DIM * un + c], Q.acc[c]);
              Q.gradp[c] = fma(pe, Q.g[b][c], Q.gradp[c]);
#pragma unroll
              for (int k = 0; k < DIM; ++k) Q.G[c * DIM + k] = fma(ue, Q.g[b][k], Q.G[c * DIM + k]);
            }
          if (A.stress)
#pragma unroll
            for (int i = 0; i < DIM; ++i)
#pragma unroll
              for (int j = 0; j < DIM; ++j) Q.sdiv[i] = fma(A.stress[(int64_t)(i * DIM + j) * A.n_unodes + un], Q.g[b][j], Q.sdiv[i]);
          if (ind != 0 && A.fsi_stress)
            {
              int si = 0;
#pragma unroll
              for (int k = 0; k < DIM; ++k)
#pragma unroll
                for (int m = 0; m <= k; ++m)
                  {
                    const double v = A.fsi_stress[(int64_t)si * A.n_unodes + un] * N;
                    Q.fsis[k * DIM + m] += v;
                    if (m != k) Q.fsis[m * DIM + k] += v;
                    ++si;
                  }
            }
        }
      Q.dp = Q.p - pp;
      Q.sigma = A.sigma_pml ? A.sigma_pml[(int64_t)cell * NQ + q] : 0.0;
      Q.rho = A.rho_f * (1 + pp / kAtm) * (1 - ind) + ind * A.rho_s; // :210-213
      Q.visc = (ind == 1 ? 1.0 : A.mu) + (eddy > 0.0 ? eddy : 0.0);    // :198-203, :214-216
      // UGN stabilisation parameters (:247-274) from the previous-step velocity
      double h = 0.0;
      for (int k = 0; k < A.n_h; ++k) h += fabs(dotd<DIM>(up, Q.g[A.h_node[k]]));
      const double v_norm = sqrt(dotd<DIM>(up, up));
      h = h != 0.0 ? 2 * v_norm / h : 0.0;
      const double nu = Q.visc / Q.rho;
      if (h != 0.0)
        {
          const double t1 = 2 / A.dt, t2 = 2 * v_norm / h, t3 = 4 * nu / (h * h);
          Q.tau_supg = 1 / sqrt(t1 * t1 + t2 * t2 + t3 * t3);
        }
      else
        Q.tau_supg = A.dt / 2;
      Q.tau_pspg = Q.tau_supg / Q.rho;
      const double localRe = v_norm * h / (2 * nu);
      Q.tau_lsic = h / 2 * v_norm * (localRe <= 3 ? localRe / 3 : 1.0);
      Q.divu = 0.0;
#pragma unroll
      for (int c = 0; c < DIM; ++c)
        {
          Q.divu += Q.G[c * DIM + c];
          Q.dv[c] = Q.u[c] - up[c];
          Q.sdiv[c] *= Q.visc / A.mu; // :288
          Q.g_bf[c] = A.grav[c] + (A.body_force ? A.body_force[((int64_t)cell * NQ + q) * DIM + c] : 0.0);
        }
#pragma unroll
      for (int l = 0; l < DIM; ++l)
        {
          double a = 0.0, b = 0.0;
#pragma unroll
          for (int k = 0; k < DIM; ++k)
            {
              a = fma(Q.u[k], Q.G[k * DIM + l], a); // u * grad u   (Tensor<1> * Tensor<2>: first index contracted)
              b = fma(Q.G[l * DIM + k], Q.u[k], b); // (grad u) u
            }
          Q.u_gradu[l] = a;
          Q.gradu_u[l] = b;
        }
      // strong momentum residual used by SUPG / PSPG (:474-492)
#pragma unroll
      for (int d = 0; d < DIM; ++d)
        Q.res[d] = Q.rho * (Q.dv[d] / A.dt + Q.u_gradu[d]) + Q.gradp[d] - Q.sdiv[d] - Q.rho * Q.g_bf[d] + Q.rho * Q.sigma * Q.u[d];
    }

    // SCnsIM::assemble cell loop. CTA = CPB cells; per cell NU*NU threads, thread (a, b) owns the (dim+1)x(dim+1)
    // block coupling node a (test) with node b (trial); the per-quadrature-point state is staged in shared
    // memory by the first NQ threads of the cell. Cells of one colour per launch.
    template <int DIM>
    __global__ void __launch_bounds__(64) scns_assemble_kernel(const ScnsArgs A)
    {
      constexpr int NU = 1 << DIM, NQ = NU, PAIRS = NU * NU, CPB = 64 / PAIRS, D1 = DIM + 1, DPC = NU * D1;
      __shared__ QPoint<DIM> sq[CPB][NQ];
      __shared__ double lrhs[CPB][DPC], ldiag[CPB][DPC];
      const int cl = threadIdx.x / PAIRS, pr = threadIdx.x % PAIRS;
      const int li = blockIdx.x * CPB + cl;
      const bool active = li < A.n_list;
      const int cell = active ? A.cell_list[li] : 0;
      const int ind = (active && A.indicator) ? A.indicator[cell] : 0;
      if (active && pr < NQ) fill_qpoint<DIM>(A, cell, pr, ind, sq[cl][pr]);
      if (active && pr < DPC)
        {
          lrhs[cl][pr] = 0.0;
          ldiag[cl][pr] = 0.0;
        }
      __syncthreads();
      if (!active) return;
      const int a = pr / NU, b = pr % NU;
      const double dt = A.dt, cp = kCpToCv, atm = kAtm, ks = kKappaS, om = 1.0 - ind;
      double K[D1][D1], r[D1];
#pragma unroll
      for (int i = 0; i < D1; ++i)
        {
          r[i] = 0.0;
#pragma unroll
          for (int j = 0; j < D1; ++j) K[i][j] = 0.0;
        }
      for (int q = 0; q < NQ; ++q)
        {
          const QPoint<DIM> &Q = sq[cl][q];
          const double w = Q.JxW, Na = Q.N[a], Nb = Q.N[b], rho = Q.rho, ts = Q.tau_supg, tp = Q.tau_pspg, tl = Q.tau_lsic;
          const double *ga = Q.g[a], *gb = Q.g[b];
          const double gagb = dotd<DIM>(ga, gb), ugb = dotd<DIM>(Q.u, gb);
          // ga . X for the vectors that meet "phi_u[j] * grad_phi_u[i]" (only when the components agree)
          const double ga_ugu = dotd<DIM>(ga, Q.u_gradu), ga_dv = dotd<DIM>(ga, Q.dv), ga_gp = dotd<DIM>(ga, Q.gradp),
                       ga_sd = dotd<DIM>(ga, Q.sdiv), ga_bf = dotd<DIM>(ga, Q.g_bf), ga_u = dotd<DIM>(ga, Q.u), ga_acc = dotd<DIM>(ga, Q.acc);
          const double same = ts * rho * Nb * ga_ugu + ts * rho * Nb * ga_dv / dt + ts * Nb * ga_gp - ts * Nb * ga_sd - ts * Nb * ga_bf * rho +
                              ts * rho * Nb * ga_u * Q.sigma - (ind == 1 ? ts * Nb * ga_acc * rho : 0.0);
#pragma unroll
          for (int c = 0; c < DIM; ++c)
            {
              const double uc = Q.u[c];
              // ---- velocity test (a,c) x velocity trial (b,d) ----
#pragma unroll
              for (int d = 0; d < DIM; ++d)
                {
                  double m = rho * Q.G[c * DIM + d] * Nb * Na;                      // (grad u phi_j) . phi_i
                  m += ts * rho * uc * Nb * dotd<DIM>(ga, &Q.G[d * DIM]);            // SUPG: (u grad phi_i).(phi_j grad u)
                  m += ts * rho * uc * Q.u[d] * gagb;                                // SUPG: (u grad phi_i).(u grad phi_j)
                  m += ts * rho * uc * ga[d] * Nb / dt;                              // SUPG acceleration
                  m += ts * rho * uc * ga[d] * Nb * Q.sigma;                         // SUPG PML
                  m += tl * rho * cp * ga[c] * gb[d] * (1.0 + Q.p * om / atm);       // LSIC velocity divergence (2 terms)
                  m += tl * rho * ga[c] * Nb * Q.gradp[d] / atm * om;                // LSIC pressure gradient (phi_j . grad p)
                  if (c == d)
                    m += Q.visc * gagb + rho * ugb * Na + rho * Na * Nb / dt + rho * Q.sigma * Nb * Na + same;
                  K[c][d] = fma(m, w, K[c][d]);
                }
              // ---- velocity test (a,c) x pressure trial b ----
              {
                double m = -ga[c] * Nb + ts * uc * gagb;
                m += tl * rho * ga[c] * Nb / dt * om / atm + tl * rho / ks * ga[c] * Nb / dt * ind;
                m += tl * rho * cp * ga[c] * Nb * om * Q.divu / atm + tl * rho * ga[c] * ugb / atm * om;
                K[c][DIM] = fma(m, w, K[c][DIM]);
              }
              // ---- pressure test a x velocity trial (b,c) ----
              {
                double m = tp * rho * Nb * dotd<DIM>(ga, &Q.G[c * DIM]) + tp * rho * Q.u[c] * gagb + tp * rho * ga[c] * Nb / dt +
                           tp * rho * ga[c] * Nb * Q.sigma;
                m += (cp * (atm + Q.p * om) * gb[c] * Na + Nb * Q.gradp[c] * Na * om) / atm;
                K[DIM][c] = fma(m, w, K[DIM][c]);
              }
              if (b == 0)
                {
                  // rhs of velocity row (a,c) (:429-512)
                  double v = -Q.visc * dotd<DIM>(&Q.G[c * DIM], ga) - rho * Q.gradu_u[c] * Na + Q.p * ga[c] - rho * Q.dv[c] * Na / dt +
                             Q.g_bf[c] * Na * rho;
                  v += -rho * Q.sigma * uc * Na;
                  v += -ts * uc * dotd<DIM>(ga, Q.res);
                  v += -(tl * rho * ga[c]) * ((Q.dp / dt * om + cp * atm * Q.divu + cp * Q.p * Q.divu * om + dotd<DIM>(Q.u, Q.gradp) * om) / atm +
                                              (1 / ks * Q.dp / dt) * ind);
                  if (ind == 1) v += dotd<DIM>(ga, &Q.fsis[c * DIM]) + rho * (Q.acc[c] * Na + ts * uc * ga_acc);
                  r[c] = fma(v, w, r[c]);
                }
            }
          // ---- pressure test a x pressure trial b ----
          {
            double m = Q.sigma * Nb * Na / atm + tp * gagb;
            m += (Nb * Q.divu * Na * om + ugb * Na * om + Na * Nb / dt * om) / atm + 1 / ks * Na * Nb * ind / dt;
            K[DIM][DIM] = fma(m, w, K[DIM][DIM]);
          }
          if (b == 0)
            {
              double v = -Q.sigma * Q.p * Na / atm;
              v += -(cp * (atm + Q.p * om) * Q.divu * Na + dotd<DIM>(Q.u, Q.gradp) * Na * om + Q.dp * Na / dt * om) / atm - 1 / ks * Q.dp * Na * ind / dt;
              v += -tp * dotd<DIM>(ga, Q.res);
              if (ind == 1) v += rho * tp * ga_acc;
              r[DIM] = fma(v, w, r[DIM]);
            }
        }
      // ---- scatter through the constraints (distribute_local_to_global, :548-560) ----
      constexpr int SPC = 4 * PAIRS; // uu | up | pu | pp slot tables, NU x NU each
      const unsigned char *slots = A.slots + (int64_t)cell * SPC;
      const int nAu = A.cell_un[(int64_t)cell * NU + a], nBu = A.cell_un[(int64_t)cell * NU + b];
      const int nAp = A.cell_pn[(int64_t)cell * NU + a], nBp = A.cell_pn[(int64_t)cell * NU + b];
      int rcon[D1], ccon[D1];
      double cinh[D1];
#pragma unroll
      for (int c = 0; c < DIM; ++c)
        {
          rcon[c] = A.con[(int64_t)DIM * nAu + c];
          ccon[c] = A.con[(int64_t)DIM * nBu + c];
          cinh[c] = (ccon[c] && A.inhom) ? A.inhom[(int64_t)DIM * nBu + c] : 0.0;
        }
      rcon[DIM] = A.con[A.n_u + nAp];
      ccon[DIM] = A.con[A.n_u + nBp];
      cinh[DIM] = (ccon[DIM] && A.inhom) ? A.inhom[A.n_u + nBp] : 0.0;
      const bool own_u = nAu < A.n_owned_u, own_p = nAp < A.n_owned_p;
      // row pointers of the four blocks
      const int64_t uu0 = own_u ? A.uu_rp[nAu] : 0, up0 = own_u ? A.up_rp[nAu] : 0, pu0 = own_p ? A.pu_rp[nAp] : 0, pp0 = own_p ? A.pp_rp[nAp] : 0;
      const int uun = own_u ? (int)(A.uu_rp[nAu + 1] - uu0) : 0, upn = own_u ? (int)(A.up_rp[nAu + 1] - up0) : 0;
      const int pun = own_p ? (int)(A.pu_rp[nAp + 1] - pu0) : 0;
      const int s_uu = slots[pr], s_up = slots[PAIRS + pr], s_pu = slots[2 * PAIRS + pr], s_pp = slots[3 * PAIRS + pr];
#pragma unroll
      for (int i = 0; i < D1; ++i)
        {
          const bool own = i < DIM ? own_u : own_p;
          if (!own) continue;
          double corr = 0.0;
#pragma unroll
          for (int j = 0; j < D1; ++j)
            {
              const double v = K[i][j];
              if (rcon[i])
                {
                  if (a == b && i == j)
                    {
                      const double dv = fabs(v);
                      if (i < DIM) A.uu[uu0 * DIM * DIM + (int64_t)(i * DIM + j) * uun + s_uu] += dv;
                      else A.pp[pp0 + s_pp] += dv;
                      ldiag[cl][a * D1 + i] = dv;
                    }
                  continue;
                }
              if (ccon[j])
                {
                  corr = fma(v, cinh[j], corr);
                  continue;
                }
              if (i < DIM && j < DIM) A.uu[uu0 * DIM * DIM + (int64_t)(i * DIM + j) * uun + s_uu] += v;
              else if (i < DIM) A.up[up0 * DIM + (int64_t)i * upn + s_up] += v;
              else if (j < DIM) A.pu[pu0 * DIM + (int64_t)j * pun + s_pu] += v;
              else A.pp[pp0 + s_pp] += v;
            }
          if (!rcon[i])
            {
              double add = -corr;
              if (b == 0) add += r[i];
              if (add != 0.0) atomicAdd(&lrhs[cl][a * D1 + i], add);
            }
        }
      __syncthreads();
      if (pr < DPC)
        {
          const int aa = pr / D1, i = pr % D1;
          const int nu_ = A.cell_un[(int64_t)cell * NU + aa], np_ = A.cell_pn[(int64_t)cell * NU + aa];
          const bool own = i < DIM ? nu_ < A.n_owned_u : np_ < A.n_owned_p;
          const int64_t g = i < DIM ? (int64_t)DIM * nu_ + i : A.n_u + np_;
          if (own)
            {
              if (!A.con[g]) A.rhs[g] += lrhs[cl][pr];
              else if (A.inhom) A.rhs[g] += ldiag[cl][pr] * A.inhom[g];
            }
        }
    }


    // 3-D default: variant 1 - 18.3 ms against 42.7 ms per assembly of 435 200 cells on a B200, results equal to 1e-16
    // (profiles/r02_scns_asm_variants.json); the 2-D kernel (1 ms per assembly at config 4) keeps the original until it is measured
    constexpr int kScnsAssembleVariant = 1;

    template <int V>
    struct Divisor
    {
      double d, inv;
    };
    template <int V>
    __device__ __forceinline__ double operator*(double x, const Divisor<V> &r)
    {
      return V ? x * r.inv : x / r.d;
    }

    template <int DIM, int V, int MINB>
    __global__ void __launch_bounds__(64, MINB) scns_assemble_v1_kernel(const ScnsArgs A)
    {
      constexpr int NU = 1 << DIM, NQ = NU, PAIRS = NU * NU, CPB = 64 / PAIRS, D1 = DIM + 1, DPC = NU * D1;
      __shared__ QPoint<DIM> sq[CPB][NQ];
      __shared__ double lrhs[CPB][DPC], ldiag[CPB][DPC];
      const int cl = threadIdx.x / PAIRS, pr = threadIdx.x % PAIRS;
      const int li = blockIdx.x * CPB + cl;
      const bool active = li < A.n_list;
      const int cell = active ? A.cell_list[li] : 0;
      const int ind = (active && A.indicator) ? A.indicator[cell] : 0;
      if (active && pr < NQ) fill_qpoint<DIM>(A, cell, pr, ind, sq[cl][pr]);
      if (active && pr < DPC)
        {
          lrhs[cl][pr] = 0.0;
          ldiag[cl][pr] = 0.0;
        }
      __syncthreads();
      if (!active) return;
      const int a = pr / NU, b = pr % NU;
      const double cp = kCpToCv, atm = kAtm, om = 1.0 - ind;
      const Divisor<V> idt{A.dt, 1.0 / A.dt}, iatm{kAtm, 1.0 / kAtm}, iks{kKappaS, 1.0 / kKappaS};
      double K[D1][D1], r[D1];
#pragma unroll
      for (int i = 0; i < D1; ++i)
        {
          r[i] = 0.0;
#pragma unroll
          for (int j = 0; j < D1; ++j) K[i][j] = 0.0;
        }
      for (int q = 0; q < NQ; ++q)
        {
          const QPoint<DIM> &Q = sq[cl][q];
          const double w = Q.JxW, Na = Q.N[a], Nb = Q.N[b], rho = Q.rho, ts = Q.tau_supg, tp = Q.tau_pspg, tl = Q.tau_lsic;
          const double *ga = Q.g[a], *gb = Q.g[b];
          const double gagb = dotd<DIM>(ga, gb), ugb = dotd<DIM>(Q.u, gb);
          // ga . X for the vectors that meet "phi_u[j] * grad_phi_u[i]" (only when the components agree)
          const double ga_ugu = dotd<DIM>(ga, Q.u_gradu), ga_dv = dotd<DIM>(ga, Q.dv), ga_gp = dotd<DIM>(ga, Q.gradp),
                       ga_sd = dotd<DIM>(ga, Q.sdiv), ga_bf = dotd<DIM>(ga, Q.g_bf), ga_u = dotd<DIM>(ga, Q.u), ga_acc = dotd<DIM>(ga, Q.acc);
          const double same = ts * rho * Nb * ga_ugu + ts * rho * Nb * ga_dv * idt + ts * Nb * ga_gp - ts * Nb * ga_sd - ts * Nb * ga_bf * rho +
                              ts * rho * Nb * ga_u * Q.sigma - (ind == 1 ? ts * Nb * ga_acc * rho : 0.0);
#pragma unroll
          for (int c = 0; c < DIM; ++c)
            {
              const double uc = Q.u[c];
              // ---- velocity test (a,c) x velocity trial (b,d) ----
#pragma unroll
              for (int d = 0; d < DIM; ++d)
                {
                  double m = rho * Q.G[c * DIM + d] * Nb * Na;                      // (grad u phi_j) . phi_i
                  m += ts * rho * uc * Nb * dotd<DIM>(ga, &Q.G[d * DIM]);            // SUPG: (u grad phi_i).(phi_j grad u)
                  m += ts * rho * uc * Q.u[d] * gagb;                                // SUPG: (u grad phi_i).(u grad phi_j)
                  m += ts * rho * uc * ga[d] * Nb * idt;                              // SUPG acceleration
                  m += ts * rho * uc * ga[d] * Nb * Q.sigma;                         // SUPG PML
                  m += tl * rho * cp * ga[c] * gb[d] * (1.0 + Q.p * om * iatm);       // LSIC velocity divergence (2 terms)
                  m += tl * rho * ga[c] * Nb * Q.gradp[d] * iatm * om;                // LSIC pressure gradient (phi_j . grad p)
                  if (c == d)
                    m += Q.visc * gagb + rho * ugb * Na + rho * Na * Nb * idt + rho * Q.sigma * Nb * Na + same;
                  K[c][d] = fma(m, w, K[c][d]);
                }
              // ---- velocity test (a,c) x pressure trial b ----
              {
                double m = -ga[c] * Nb + ts * uc * gagb;
                m += tl * rho * ga[c] * Nb * idt * om * iatm + tl * rho * iks * ga[c] * Nb * idt * ind;
                m += tl * rho * cp * ga[c] * Nb * om * Q.divu * iatm + tl * rho * ga[c] * ugb * iatm * om;
                K[c][DIM] = fma(m, w, K[c][DIM]);
              }
              // ---- pressure test a x velocity trial (b,c) ----
              {
                double m = tp * rho * Nb * dotd<DIM>(ga, &Q.G[c * DIM]) + tp * rho * Q.u[c] * gagb + tp * rho * ga[c] * Nb * idt +
                           tp * rho * ga[c] * Nb * Q.sigma;
                m += (cp * (atm + Q.p * om) * gb[c] * Na + Nb * Q.gradp[c] * Na * om) * iatm;
                K[DIM][c] = fma(m, w, K[DIM][c]);
              }
              if (b == 0)
                {
                  // rhs of velocity row (a,c) (:429-512)
                  double v = -Q.visc * dotd<DIM>(&Q.G[c * DIM], ga) - rho * Q.gradu_u[c] * Na + Q.p * ga[c] - rho * Q.dv[c] * Na * idt +
                             Q.g_bf[c] * Na * rho;
                  v += -rho * Q.sigma * uc * Na;
                  v += -ts * uc * dotd<DIM>(ga, Q.res);
                  v += -(tl * rho * ga[c]) * ((Q.dp * idt * om + cp * atm * Q.divu + cp * Q.p * Q.divu * om + dotd<DIM>(Q.u, Q.gradp) * om) * iatm +
                                              (1 * iks * Q.dp * idt) * ind);
                  if (ind == 1) v += dotd<DIM>(ga, &Q.fsis[c * DIM]) + rho * (Q.acc[c] * Na + ts * uc * ga_acc);
                  r[c] = fma(v, w, r[c]);
                }
            }
          // ---- pressure test a x pressure trial b ----
          {
            double m = Q.sigma * Nb * Na * iatm + tp * gagb;
            m += (Nb * Q.divu * Na * om + ugb * Na * om + Na * Nb * idt * om) * iatm + 1 * iks * Na * Nb * ind * idt;
            K[DIM][DIM] = fma(m, w, K[DIM][DIM]);
          }
          if (b == 0)
            {
              double v = -Q.sigma * Q.p * Na * iatm;
              v += -(cp * (atm + Q.p * om) * Q.divu * Na + dotd<DIM>(Q.u, Q.gradp) * Na * om + Q.dp * Na * idt * om) * iatm - 1 * iks * Q.dp * Na * ind * idt;
              v += -tp * dotd<DIM>(ga, Q.res);
              if (ind == 1) v += rho * tp * ga_acc;
              r[DIM] = fma(v, w, r[DIM]);
            }
        }
      // ---- scatter through the constraints (distribute_local_to_global, :548-560) ----
      constexpr int SPC = 4 * PAIRS; // uu | up | pu | pp slot tables, NU x NU each
      const unsigned char *slots = A.slots + (int64_t)cell * SPC;
      const int nAu = A.cell_un[(int64_t)cell * NU + a], nBu = A.cell_un[(int64_t)cell * NU + b];
      const int nAp = A.cell_pn[(int64_t)cell * NU + a], nBp = A.cell_pn[(int64_t)cell * NU + b];
      int rcon[D1], ccon[D1];
      double cinh[D1];
#pragma unroll
      for (int c = 0; c < DIM; ++c)
        {
          rcon[c] = A.con[(int64_t)DIM * nAu + c];
          ccon[c] = A.con[(int64_t)DIM * nBu + c];
          cinh[c] = (ccon[c] && A.inhom) ? A.inhom[(int64_t)DIM * nBu + c] : 0.0;
        }
      rcon[DIM] = A.con[A.n_u + nAp];
      ccon[DIM] = A.con[A.n_u + nBp];
      cinh[DIM] = (ccon[DIM] && A.inhom) ? A.inhom[A.n_u + nBp] : 0.0;
      const bool own_u = nAu < A.n_owned_u, own_p = nAp < A.n_owned_p;
      // row pointers of the four blocks
      const int64_t uu0 = own_u ? A.uu_rp[nAu] : 0, up0 = own_u ? A.up_rp[nAu] : 0, pu0 = own_p ? A.pu_rp[nAp] : 0, pp0 = own_p ? A.pp_rp[nAp] : 0;
      const int uun = own_u ? (int)(A.uu_rp[nAu + 1] - uu0) : 0, upn = own_u ? (int)(A.up_rp[nAu + 1] - up0) : 0;
      const int pun = own_p ? (int)(A.pu_rp[nAp + 1] - pu0) : 0;
      const int s_uu = slots[pr], s_up = slots[PAIRS + pr], s_pu = slots[2 * PAIRS + pr], s_pp = slots[3 * PAIRS + pr];
      // where entry (i, j) of this thread's block goes (nullptr: not stored)
      auto target = [&](int i, int j) -> double * {
        if (i < DIM && j < DIM) return A.uu + uu0 * DIM * DIM + (int64_t)(i * DIM + j) * uun + s_uu;
        if (i < DIM) return A.up + up0 * DIM + (int64_t)i * upn + s_up;
        if (j < DIM) return A.pu + pu0 * DIM + (int64_t)j * pun + s_pu;
        return A.pp + pp0 + s_pp;
      };
      if (V == 0)
        {
#pragma unroll
          for (int i = 0; i < D1; ++i)
            {
              const bool own = i < DIM ? own_u : own_p;
              if (!own) continue;
              double corr = 0.0;
#pragma unroll
              for (int j = 0; j < D1; ++j)
                {
                  const double v = K[i][j];
                  if (rcon[i])
                    {
                      if (a == b && i == j)
                        {
                          const double dv = fabs(v);
                          *target(i, j) += dv;
                          ldiag[cl][a * D1 + i] = dv;
                        }
                      continue;
                    }
                  if (ccon[j])
                    {
                      corr = fma(v, cinh[j], corr);
                      continue;
                    }
                  *target(i, j) += v;
                }
              if (!rcon[i])
                {
                  double add = -corr;
                  if (b == 0) add += r[i];
                  if (add != 0.0) atomicAdd(&lrhs[cl][a * D1 + i], add);
                }
            }
        }
      else
        {
          // pass 1: decide per entry (stored value in K, bit in `store`), right-hand side corrections
          unsigned store = 0;
#pragma unroll
          for (int i = 0; i < D1; ++i)
            {
              const bool own = i < DIM ? own_u : own_p;
              if (!own) continue;
              double corr = 0.0;
#pragma unroll
              for (int j = 0; j < D1; ++j)
                {
                  if (rcon[i])
                    {
                      if (a == b && i == j)
                        {
                          K[i][j] = fabs(K[i][j]);
                          ldiag[cl][a * D1 + i] = K[i][j];
                          store |= 1u << (i * D1 + j);
                        }
                    }
                  else if (ccon[j])
                    corr = fma(K[i][j], cinh[j], corr);
                  else
                    store |= 1u << (i * D1 + j);
                }
              if (!rcon[i])
                {
                  double add = -corr;
                  if (b == 0) add += r[i];
                  if (add != 0.0) atomicAdd(&lrhs[cl][a * D1 + i], add);
                }
            }
          // pass 2: all loads, then all stores
          double cur[D1][D1];
#pragma unroll
          for (int i = 0; i < D1; ++i)
#pragma unroll
            for (int j = 0; j < D1; ++j) cur[i][j] = (store >> (i * D1 + j)) & 1u ? *target(i, j) : 0.0;
#pragma unroll
          for (int i = 0; i < D1; ++i)
#pragma unroll
            for (int j = 0; j < D1; ++j)
              if ((store >> (i * D1 + j)) & 1u) *target(i, j) = cur[i][j] + K[i][j];
        }
      __syncthreads();
      if (pr < DPC)
        {
          const int aa = pr / D1, i = pr % D1;
          const int nu_ = A.cell_un[(int64_t)cell * NU + aa], np_ = A.cell_pn[(int64_t)cell * NU + aa];
          const bool own = i < DIM ? nu_ < A.n_owned_u : np_ < A.n_owned_p;
          const int64_t g = i < DIM ? (int64_t)DIM * nu_ + i : A.n_u + np_;
          if (own)
            {
              if (!A.con[g]) A.rhs[g] += lrhs[cl][pr];
              else if (A.inhom) A.rhs[g] += ldiag[cl][pr] * A.inhom[g];
            }
        }
    }

    // rowsum(|A_vv|)^-1 per velocity dof (mpi_supg_solver.cpp:68-118)
    template <int DIM>
    __global__ void abs_rowsum_inv_kernel(int n_brows, const int64_t *__restrict__ rp, const double *__restrict__ val, double *__restrict__ out)
    {
      const int row = blockIdx.x * blockDim.x + threadIdx.x;
      if (row >= n_brows) return;
      const int64_t base = rp[row];
      const int nb = (int)(rp[row + 1] - base);
#pragma unroll
      for (int r = 0; r < DIM; ++r)
        {
          double s = 0.0;
          for (int c = 0; c < DIM; ++c)
            for (int j = 0; j < nb; ++j) s += fabs(val[base * DIM * DIM + (int64_t)(r * DIM + c) * nb + j]);
          out[(int64_t)DIM * row + r] = 1.0 / s;
        }
    }

    // diagonal of B2pp = App - Apv diag(rowsum|Avv|)^-1 Avp (:120-127); its inverse is the preconditioner of Tpp
    template <int DIM>
    __global__ void b2pp_diag_inv_kernel(int n_p, const int64_t *__restrict__ pu_rp, const int *__restrict__ pu_col,
                                         const double *__restrict__ pu_val, const int64_t *__restrict__ up_rp,
                                         const int *__restrict__ up_col, const double *__restrict__ up_val,
                                         const int64_t *__restrict__ pp_rp, const int *__restrict__ pp_col,
                                         const double *__restrict__ pp_val, const double *__restrict__ rinv, int n_owned_u,
                                         double *__restrict__ out)
    {
      const int i = blockIdx.x * blockDim.x + threadIdx.x;
      if (i >= n_p) return;
      double d = 0.0;
      for (int64_t k = pp_rp[i]; k < pp_rp[i + 1]; ++k)
        if (pp_col[k] == i) d = pp_val[k];
      const int64_t rb = pu_rp[i];
      const int rn = (int)(pu_rp[i + 1] - rb);
      for (int jk = 0; jk < rn; ++jk)
        {
          const int k = pu_col[rb + jk];
          if (k >= n_owned_u) continue; // rows of ghost velocity nodes live on their owner (multi-rank: diagonal is approximate there)
          const int64_t ub = up_rp[k];
          const int un = (int)(up_rp[k + 1] - ub);
          int m = -1;
          for (int t = 0; t < un; ++t)
            if (up_col[ub + t] == i) m = t;
          if (m < 0) continue;
#pragma unroll
          for (int c = 0; c < DIM; ++c)
            d -= pu_val[rb * DIM + (int64_t)c * rn + jk] * rinv[(int64_t)DIM * k + c] * up_val[ub * DIM + (int64_t)c * un + m];
        }
      out[i] = d != 0.0 ? 1.0 / d : 1.0;
    }

    // B2pp = A_pp - A_pv diag(rowsum|A_vv|)^-1 A_vp (mpi_supg_solver.cpp:120-127) on the pattern of A_pv A_vp (scalar CSR, the
    // layout Ilu0 factorises); one thread per pressure row
    template <int DIM>
    __global__ void b2pp_matrix_kernel(int n_p, const int64_t *__restrict__ pu_rp, const int *__restrict__ pu_col,
                                       const double *__restrict__ pu_val, const int64_t *__restrict__ up_rp,
                                       const int *__restrict__ up_col, const double *__restrict__ up_val,
                                       const int64_t *__restrict__ pp_rp, const int *__restrict__ pp_col,
                                       const double *__restrict__ pp_val, const double *__restrict__ rinv, const int *__restrict__ b_rp,
                                       const int *__restrict__ b_col, double *__restrict__ b_val)
    {
      const int i = blockIdx.x * blockDim.x + threadIdx.x;
      if (i >= n_p) return;
      const int b0 = b_rp[i], b1 = b_rp[i + 1];
      auto find = [&](int c) {
        int lo = b0, hi = b1 - 1;
        while (lo <= hi)
          {
            const int mid = (lo + hi) >> 1;
            if (b_col[mid] == c) return mid;
            if (b_col[mid] < c) lo = mid + 1; else hi = mid - 1;
          }
        return -1;
      };
      for (int p = b0; p < b1; ++p) b_val[p] = 0.0;
      for (int64_t k = pp_rp[i]; k < pp_rp[i + 1]; ++k)
        {
          const int pos = find(pp_col[k]);
          if (pos >= 0) b_val[pos] += pp_val[k];
        }
      const int64_t rb = pu_rp[i];
      const int rn = (int)(pu_rp[i + 1] - rb);
      for (int jk = 0; jk < rn; ++jk)
        {
          const int k = pu_col[rb + jk];
          const int64_t ub = up_rp[k];
          const int un = (int)(up_rp[k + 1] - ub);
          double w[DIM];
#pragma unroll
          for (int c = 0; c < DIM; ++c) w[c] = pu_val[rb * DIM + (int64_t)c * rn + jk] * rinv[(int64_t)DIM * k + c];
          for (int t = 0; t < un; ++t)
            {
              double v = 0.0;
#pragma unroll
              for (int c = 0; c < DIM; ++c) v += w[c] * up_val[ub * DIM + (int64_t)c * un + t];
              if (v == 0.0) continue;
              const int pos = find(up_col[ub + t]);
              if (pos >= 0) b_val[pos] -= v;
            }
        }
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
  SCnsIM::SCnsIM(Context &ctx_, Triangulation &tria, const Parameters::AllParameters &params) : InsIM(ctx_, tria, params, false)
  {
    const unsigned pu = parameters.fluid_velocity_degree, pp = parameters.fluid_pressure_degree;
    if (pu < 1 || pu > 2 || !(pp == 1 || pp == pu))
      throw std::runtime_error("SCnsIM: velocity degree 1 or 2 with pressure degree 1 or equal to it");
    control.fgmres_rel = 1e-6; // SUPGFluidSolver::solve: SolverControl(m, 1e-6 * |rhs|) (mpi_supg_solver.cpp:311-312)
  }

  void SCnsIM::setup_dofs()
  {
    fs.setup(ctx, triangulation, (int)parameters.fluid_velocity_degree, (int)parameters.fluid_pressure_degree, true);
    dofs_ready = true;
    ilu_vv = Ilu0(); // factors belong to the old pattern (refine_mesh calls setup_dofs again)
    ilu_b2 = Ilu0();
  }

  void SCnsIM::initialize_system()
  {
    InsIM::initialize_system();
    const int dim = fs.dim, nsym = dim * (dim + 1) / 2;
    cudaStream_t s = ctx.stream;
    fsi_stress.alloc((size_t)nsym * fs.un.n_nodes);
    fsi_stress.zero(s);
    d_rowsum_inv.alloc(fs.n_u);
    d_b2pp_diag_inv.alloc(fs.n_p);
    d_pt1.alloc(fs.n_p);
    d_pt2.alloc(fs.n_p);
    d_ut1.alloc(fs.n_u);
    d_ut2.alloc(fs.n_u);
    // user fields at the quadrature points of the local cells (Q1 map)
    if (sigma_pml_field || body_force)
      {
        const int nq = fs.nq, nv = fs.nv;
        std::vector<double> sp, bf;
        if (sigma_pml_field) sp.resize((size_t)fs.n_cells * nq);
        if (body_force) bf.resize((size_t)fs.n_cells * nq * dim);
        for (int c = 0; c < fs.n_cells; ++c)
          for (int q = 0; q < nq; ++q)
            {
              double x[3] = {0, 0, 0};
              for (int v = 0; v < nv; ++v)
                for (int d = 0; d < dim; ++d)
                  x[d] += fs.tab_geo.N[(size_t)q * nv + v] * triangulation.vertices[(size_t)triangulation.cells[(size_t)fs.local_cells[c] * nv + v] * dim + d];
              if (sigma_pml_field) sp[(size_t)c * nq + q] = sigma_pml_field(x, 0);
              if (body_force)
                for (int d = 0; d < dim; ++d) bf[((size_t)c * nq + q) * dim + d] = body_force(x, (unsigned)d);
            }
        if (sigma_pml_field) d_sigma_pml.upload(sp, s);
        if (body_force) d_body_force.upload(bf, s);
      }
    if (initial_condition) // apply_initial_condition (mpi_fluid_solver.cpp:368-414)
      {
        std::vector<double> ic(fs.n_dofs);
        for (int n = 0; n < fs.un.n_nodes; ++n)
          for (int c = 0; c < dim; ++c) ic[(size_t)dim * n + c] = initial_condition(&fs.un.coords[(size_t)n * dim], (unsigned)c);
        for (int n = 0; n < fs.pn.n_nodes; ++n) ic[(size_t)fs.n_u + n] = initial_condition(&fs.pn.coords[(size_t)n * dim], (unsigned)dim);
        present_solution.upload(ic, s);
      }
    IFEM_CUDA(cudaStreamSynchronize(s));
    if (turbulence_model) turbulence_model->initialize_system(); // mpi_supg_solver.cpp:290-293
  }

  void SCnsIM::attach_turbulence_model(const std::string &model_name)
  {
    if (model_name != "Spalart-Allmaras") throw std::runtime_error("attach_turbulence_model: model <" + model_name + "> is not implemented");
    turbulence_model = std::make_unique<SpalartAllmaras>(ctx, *this);
    after_make_constraints = [this] { turbulence_model->make_constraints(); };
    on_mesh_change = [this] { turbulence_model->mesh_changed(); };
    if (dofs_ready) // attached after setup: what make_constraints() / initialize_system() would have done (mpi_supg_solver.cpp:290-293)
      {
        turbulence_model->make_constraints();
        turbulence_model->initialize_system();
      }
  }

  void SCnsIM::assemble(bool use_nonzero_constraints)
  {
    ScopedTimer t(ctx, timer_ms["Assemble system"]);
    if (fs.n_ranks > 1)
      {
        fs.halo_update(ctx, evaluation_point.p);
        fs.halo_update(ctx, present_solution.p);
        fs.halo_update(ctx, fsi_acceleration.p);
      }
    cudaStream_t s = ctx.stream;
    fs.A_uu.zero(s);
    fs.A_up.zero(s);
    fs.A_pu.zero(s);
    fs.A_pp.zero(s);
    fs.rhs.zero(s);
    if (fs.pu != 1 || fs.pp != 1)
      {
        ScnsGenericInput in{};
        in.eval_pt = evaluation_point.p;
        in.present = present_solution.p;
        in.fsi_acc = fsi_acceleration.p;
        in.stress = stress.p;
        in.fsi_stress = fsi_stress.p;
        in.sigma_pml = d_sigma_pml.n ? d_sigma_pml.p : nullptr;
        in.body_force = d_body_force.n ? d_body_force.p : nullptr;
        in.mu = parameters.viscosity;
        in.rho_f = parameters.fluid_rho;
        in.rho_s = parameters.solid_rho;
        in.dt = time.get_delta_t();
        for (int d = 0; d < 3; ++d) in.grav[d] = d < (int)parameters.gravity.size() ? parameters.gravity[d] : 0.0;
        scns_assemble_generic(ctx, fs, in, use_nonzero_constraints);
        neumann_faces(ctx, fs);
        fs.hanging.condense(ctx, fs, use_nonzero_constraints ? fs.d_nonzero_val.p : nullptr);
        return;
      }
    ScnsArgs a{};
    a.cell_un = fs.d_cell_un.p;
    a.cell_pn = fs.d_cell_pn.p;
    a.indicator = fs.d_indicator.p;
    a.cell_x = fs.d_cell_x.p;
    a.tables = fs.d_tables.p;
    a.slots = fs.d_slots.p;
    a.con = fs.d_con.p;
    a.eval_pt = evaluation_point.p;
    a.present = present_solution.p;
    a.fsi_acc = fsi_acceleration.p;
    a.stress = stress.p;
    a.fsi_stress = fsi_stress.p;
    a.sigma_pml = d_sigma_pml.n ? d_sigma_pml.p : nullptr;
    a.body_force = d_body_force.n ? d_body_force.p : nullptr;
    a.inhom = use_nonzero_constraints ? fs.d_nonzero_val.p : nullptr;
    a.eddy = turbulence_model ? turbulence_model->eddy_viscosity.p : nullptr;
    a.n_u = fs.n_u;
    a.n_unodes = fs.un.n_nodes;
    a.n_owned_u = fs.n_owned_unodes;
    a.n_owned_p = fs.n_owned_pnodes;
    // the first dofs_per_cell / dofs_per_vertex system shape functions (mpi_scnsim.cpp:251-257): vertex v
    // carries dim velocity components then the pressure, all with the Q1 shape of that vertex
    const int dim = fs.dim;
    a.n_h = 1 << dim;
    for (int k = 0; k < a.n_h; ++k) a.h_node[k] = k / (dim + 1);
    a.mu = parameters.viscosity;
    a.rho_f = parameters.fluid_rho;
    a.rho_s = parameters.solid_rho;
    a.dt = time.get_delta_t();
    for (int d = 0; d < 3; ++d) a.grav[d] = d < (int)parameters.gravity.size() ? parameters.gravity[d] : 0.0;
    a.uu_rp = fs.A_uu.rowptr.p;
    a.up_rp = fs.A_up.rowptr.p;
    a.pu_rp = fs.A_pu.rowptr.p;
    a.pp_rp = fs.A_pp.rowptr.p;
    a.uu = fs.A_uu.val.p;
    a.up = fs.A_up.val.p;
    a.pu = fs.A_pu.val.p;
    a.pp = fs.A_pp.val.p;
    a.rhs = fs.rhs.p;
    const int n_colours = (int)fs.colour_offsets.size() - 1;
    // kernel variant (see scns_assemble_kernel): IFEM_SCNS_ASM = 0 / 1 overrides the default
    const char *env = std::getenv("IFEM_SCNS_ASM");
    const int variant = env ? std::atoi(env) : (dim == 3 ? kScnsAssembleVariant : 0);
    for (int k = 0; k < n_colours; ++k)
      {
        a.n_list = fs.colour_offsets[k + 1] - fs.colour_offsets[k];
        a.cell_list = fs.d_colour_order.p + fs.colour_offsets[k];
        if (!a.n_list) continue;
        if (dim == 2 && variant)
          scns_assemble_v1_kernel<2, 1, 7><<<(a.n_list + 3) / 4, 64, 0, s>>>(a);
        else if (dim == 2)
          scns_assemble_kernel<2><<<(a.n_list + 3) / 4, 64, 0, s>>>(a);
        else if (variant == 2) // variant 1 held to 128 registers: 8 CTAs per SM instead of 6
          scns_assemble_v1_kernel<3, 1, 8><<<a.n_list, 64, 0, s>>>(a);
        else if (variant == 3) // variant 1 with the register allocation left to the compiler (5 CTAs per SM)
          scns_assemble_v1_kernel<3, 1, 1><<<a.n_list, 64, 0, s>>>(a);
        else if (variant == 4) // the reference arithmetic (divisions) with the batched scatter only
          scns_assemble_v1_kernel<3, 0, 1><<<a.n_list, 64, 0, s>>>(a);
        else if (variant)
          scns_assemble_v1_kernel<3, 1, 6><<<a.n_list, 64, 0, s>>>(a);
        else
          scns_assemble_kernel<3><<<a.n_list, 64, 0, s>>>(a);
        IFEM_KERNEL_CHECK();
        ctx.kernel_launches++;
      }
    neumann_faces(ctx, fs); // same pressure face term as InsIM (:516-546)
    fs.hanging.condense(ctx, fs, a.inhom); // hanging-node lines of a locally refined mesh (distribute_local_to_global, :548-560)
  }

  // BlockIncompSchurPreconditioner::vmult (mpi_supg_solver.cpp:137-192). The two Hypre-Euclid ILU(0) factors
  // (rank-count dependent in the reference) are replaced by rank-independent Jacobi factors: P_vv^-1 = inverse of
  // the node-diagonal blocks of A_vv, and diag(B2pp)^-1 as the preconditioner of the T_pp solve.
  bool SCnsIM::use_ilu() const
  {
    constexpr int64_t kIluMaxRows = 60000; // scalar rows of A_vv up to which the one-CTA level-scheduled sweeps pay (ilu0.h)
    if (control.supg_ilu == 0 || fs.n_ranks > 1) return false;
    return control.supg_ilu == 1 || fs.n_u <= kIluMaxRows;
  }

  void SCnsIM::precondition_supg(const double *src, double *dst)
  {
    const int64_t n_u = fs.n_u;
    const VecSpace &vu = fs.vs_u, &vp = fs.vs_p;
    const double *src_u = src, *src_p = src + n_u;
    double *dst_u = dst, *dst_p = dst + n_u;
    const bool ilu = use_ilu();
    auto Pvv = [&](const double *x, double *y) {
      if (ilu) ilu_vv.solve(ctx, x, y);
      else block_diag_apply(ctx, fs.n_owned_unodes, fs.dim, d_binv.p, x, y);
    };
    // ptmp = src_p - A_pv P_vv^-1 src_u
    Pvv(src_u, d_ut1.p);
    fs.halo_u.update(ctx, d_ut1.p);
    spmv(ctx, fs.A_pu, d_ut1.p, d_pt1.p);
    axpby(ctx, vp, 1.0, src_p, -1.0, d_pt1.p);
    // dst_p = T_pp^-1 ptmp,  T_pp = A_pp - A_pv P_vv^-1 A_vp  (matrix-free, :20-32)
    {
      ScopedTimer t(ctx, timer_ms["Solving Tpp"]);
      // control.a_inv_fp32 != 0: the three products of the INNER solve stream fp32 copies of the blocks (x, y and the sums stay
      // fp64). T_pp^-1 is applied to 1e-3 inside a preconditioner of a flexible GMRES - the perturbation of the operator (6e-8
      // relative per entry) is far below that; the outer operator, residuals and bases stay fp64
      const bool f32 = control.a_inv_fp32 != 0 && fs.A_up.val32.p != nullptr;
      LinOp Tpp = [&](const double *x, double *y) {
        fs.halo_p.update(ctx, const_cast<double *>(x));
        if (f32) spmv_fp32(ctx, fs.A_up, x, d_ut2.p); else spmv(ctx, fs.A_up, x, d_ut2.p);
        Pvv(d_ut2.p, d_ut1.p);
        fs.halo_u.update(ctx, d_ut1.p);
        if (f32) spmv_fp32(ctx, fs.A_pu, d_ut1.p, d_pt2.p); else spmv(ctx, fs.A_pu, d_ut1.p, d_pt2.p);
        if (f32) spmv_fp32(ctx, fs.A_pp, x, y); else spmv(ctx, fs.A_pp, x, y);
        axpy(ctx, vp, -1.0, d_pt2.p, y);
      };
      LinOp B2 = [&](const double *x, double *y) {
        if (ilu) ilu_b2.solve(ctx, x, y);
        else hadamard(ctx, vp, d_b2pp_diag_inv.p, x, y);
      };
      const double tol = 1e-3 * nrm2(ctx, vp, d_pt1.p);
      if (tol > 0)
        {
          const SolveResult r = fgmres(ctx, vp, Tpp, B2, d_pt1.p, dst_p, tol, n_p_global, 50, pool_tpp, /*fused_orthogonalisation=*/true);
          tpp_its += r.iterations;
        }
      else
        fill(ctx, vp, 0.0, dst_p);
    }
    // dst_u = P_vv^-1 src_u - P_vv^-1 A_vp dst_p
    fs.halo_p.update(ctx, dst_p);
    spmv(ctx, fs.A_up, dst_p, d_ut2.p);
    axpby(ctx, vu, 1.0, src_u, -1.0, d_ut2.p);
    Pvv(d_ut2.p, dst_u);
    cur.precond_applies++;
  }

  std::pair<unsigned int, double> SCnsIM::solve(bool use_nonzero_constraints)
  {
    ScopedTimer t(ctx, timer_ms["Solve linear system"]);
    cudaStream_t s = ctx.stream;
    block_diag_inverse(ctx, fs.A_uu, d_binv.p);
    const int nbr = fs.A_uu.n_brows;
    if (fs.dim == 2)
      {
        abs_rowsum_inv_kernel<2><<<(nbr + 127) / 128, 128, 0, s>>>(nbr, fs.A_uu.rowptr.p, fs.A_uu.val.p, d_rowsum_inv.p);
        b2pp_diag_inv_kernel<2><<<(fs.n_owned_pnodes + 127) / 128, 128, 0, s>>>(
          fs.n_owned_pnodes, fs.A_pu.rowptr.p, fs.A_pu.col.p, fs.A_pu.val.p, fs.A_up.rowptr.p, fs.A_up.col.p, fs.A_up.val.p,
          fs.A_pp.rowptr.p, fs.A_pp.col.p, fs.A_pp.val.p, d_rowsum_inv.p, fs.n_owned_unodes, d_b2pp_diag_inv.p);
      }
    else
      {
        abs_rowsum_inv_kernel<3><<<(nbr + 127) / 128, 128, 0, s>>>(nbr, fs.A_uu.rowptr.p, fs.A_uu.val.p, d_rowsum_inv.p);
        b2pp_diag_inv_kernel<3><<<(fs.n_owned_pnodes + 127) / 128, 128, 0, s>>>(
          fs.n_owned_pnodes, fs.A_pu.rowptr.p, fs.A_pu.col.p, fs.A_pu.val.p, fs.A_up.rowptr.p, fs.A_up.col.p, fs.A_up.val.p,
          fs.A_pp.rowptr.p, fs.A_pp.col.p, fs.A_pp.val.p, d_rowsum_inv.p, fs.n_owned_unodes, d_b2pp_diag_inv.p);
      }
    IFEM_KERNEL_CHECK();
    ctx.kernel_launches += 2;
    if (use_ilu())
      {
        // the reference's two Euclid factorisations (mpi_supg_solver.cpp:51, 130-133), redone for every Newton matrix
        if (!ilu_vv.ready())
          {
            std::vector<int64_t> rp;
            std::vector<int> ci;
            scalar_pattern(fs.P_uu, fs.dim, rp, ci);
            ilu_vv.setup(ctx, rp, ci);
            ilu_b2.setup(ctx, fs.P_schur.rowptr, fs.P_schur.col);
          }
        bcsr_to_scalar(ctx, fs.A_uu, ilu_vv.rowptr.p, ilu_vv.val.p);
        ilu_vv.factor(ctx);
        const int np_ = fs.n_owned_pnodes;
        if (fs.dim == 2)
          b2pp_matrix_kernel<2><<<(np_ + 127) / 128, 128, 0, s>>>(np_, fs.A_pu.rowptr.p, fs.A_pu.col.p, fs.A_pu.val.p, fs.A_up.rowptr.p, fs.A_up.col.p,
                                                                 fs.A_up.val.p, fs.A_pp.rowptr.p, fs.A_pp.col.p, fs.A_pp.val.p, d_rowsum_inv.p,
                                                                 ilu_b2.rowptr.p, ilu_b2.col.p, ilu_b2.val.p);
        else
          b2pp_matrix_kernel<3><<<(np_ + 127) / 128, 128, 0, s>>>(np_, fs.A_pu.rowptr.p, fs.A_pu.col.p, fs.A_pu.val.p, fs.A_up.rowptr.p, fs.A_up.col.p,
                                                                 fs.A_up.val.p, fs.A_pp.rowptr.p, fs.A_pp.col.p, fs.A_pp.val.p, d_rowsum_inv.p,
                                                                 ilu_b2.rowptr.p, ilu_b2.col.p, ilu_b2.val.p);
        IFEM_KERNEL_CHECK();
        ctx.kernel_launches++;
        ilu_b2.factor(ctx);
      }
    if (control.a_inv_fp32 != 0) // fp32 copies of the blocks the inner T_pp solve streams (precondition_supg)
      {
        make_fp32_copy(ctx, fs.A_up);
        make_fp32_copy(ctx, fs.A_pu);
        make_fp32_copy(ctx, fs.A_pp);
      }
    const VecSpace &va = fs.vs_all;
    const double nrm = nrm2(ctx, va, fs.rhs.p);
    const double tol = control.fgmres_rel * nrm; // SolverControl(m, 1e-6 * |rhs|), mpi_supg_solver.cpp:311-312
    LinOp A = [&](const double *x, double *y) { block_vmult(ctx, fs, x, y); };
    LinOp P = [&](const double *x, double *y) { precondition_supg(x, y); };
    SolveResult r;
    if (nrm > 0)
      r = fgmres(ctx, va, A, P, fs.rhs.p, newton_update.p, tol, n_dofs_global, control.basis_size, pool_fgmres);
    else
      fill(ctx, va, 0.0, newton_update.p);
    set_flagged(ctx, fs.n_dofs, fs.d_con.p, use_nonzero_constraints ? fs.d_nonzero_val.p : nullptr, newton_update.p);
    fs.hanging.distribute(ctx, fs, newton_update.p); // constraints.distribute(newton_update), mpi_supg_solver.cpp:323-325
    return {(unsigned)r.iterations, r.residual};
  }

  void SCnsIM::run_one_step(bool apply_nonzero_constraints, bool /*assemble_system*/)
  {
    io_before_step();
    time.increment();
    if (verbose && fs.rank == 0)
      std::printf("%s\nTime step = %u, at t = %e\n", std::string(96, '*').c_str(), time.get_timestep(), time.current());
    double current_residual = 1.0, initial_residual = 1.0, relative_residual = 1.0;
    unsigned int outer_iteration = 0;
    const VecSpace &n = fs.vs_all;
    copy(ctx, n, present_solution.p, evaluation_point.p);
    while (relative_residual > parameters.fluid_tolerance && current_residual > 1e-14) // mpi_supg_solver.cpp:354-355
      {
        if (outer_iteration >= parameters.fluid_max_iterations) throw std::runtime_error("Too many Newton iterations!");
        fill(ctx, n, 0.0, newton_update.p);
        cur = NewtonRecord{};
        tpp_its = 0;
        const bool nz = apply_nonzero_constraints && outer_iteration == 0;
        assemble(nz);
        const auto state = solve(nz);
        current_residual = nrm2(ctx, n, fs.rhs.p);
        axpy(ctx, n, 1.0, newton_update.p, evaluation_point.p);
        fs.halo_update(ctx, evaluation_point.p);
        if (outer_iteration == 0) initial_residual = current_residual;
        relative_residual = current_residual / initial_residual;
        cur.timestep = time.get_timestep();
        cur.iteration = outer_iteration;
        cur.abs_res = current_residual;
        cur.rel_res = relative_residual;
        cur.gmres_its = (int)state.first;
        cur.gmres_res = state.second;
        cur.a_inv_its = tpp_its; // INNER_GMRES_ITR of the reference's log line
        history.push_back(cur);
        if (verbose && fs.rank == 0)
          std::printf(" ITR = %-2u ABS_RES = %e REL_RES = %e GMRES_ITR = %-3u GMRES_RES = %e INNER_GMRES_ITR = %-3d\n", outer_iteration,
                      current_residual, relative_residual, state.first, state.second, tpp_its);
        outer_iteration++;
      }
    lin3(ctx, n, solution_increment.p, present_solution.p, -1.0, evaluation_point.p, 0.0, evaluation_point.p);
    copy(ctx, n, evaluation_point.p, present_solution.p);
    update_stress(); // :417
    io_after_step();
  }

  // SUPGFluidSolver::run (mpi_supg_solver.cpp:427-486). With hard-coded boundary values the clock of the boundary
  // functions runs one step ahead of the solver's (advance_time before the first make_constraints, :438-444) and every
  // step re-makes the constraints and applies the nonzero ones (:470-480): the functions return the increment of the
  // boundary value over the step. Only the values change between steps - flags, pattern and colouring stay.
  void SCnsIM::run()
  {
    const bool time_dependent = !hard_coded.empty();
    const bool success_load = load_checkpoint(); // :433; false unless an output directory is set
    // the clock of the boundary functions runs one step ahead (:438-444) whether or not setup() was called before run()
    const bool advance_clock = time_dependent && !bc_clock_started && time.get_timestep() == 0;
    if (advance_clock)
      {
        bc_time += time.get_delta_t();
        bc_clock_started = true;
      }
    if (!dofs_ready)
      {
        triangulation.refine_global(parameters.global_refinements.empty() ? 0 : parameters.global_refinements[0]);
        setup_dofs();
        make_constraints();
        initialize_system();
      }
    else if (advance_clock)
      make_constraints();
    if (!success_load)
      {
        if (turbulence_model) turbulence_model->run_one_step(true); // :456-459
        run_one_step(true);
      }
    while (time.end() - time.current() > 1e-12)
      {
        if (turbulence_model) turbulence_model->run_one_step(false); // :464-467
        if (time_dependent)
          {
            bc_time += time.get_delta_t();
            make_constraints();
            run_one_step(true);
          }
        else
          run_one_step(false);
      }
  }
} // namespace ifem
