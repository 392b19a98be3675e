// SCnsIM::assemble for velocity / pressure degrees other than Q1/Q1 (reference source/mpi_scnsim.cpp:15-568 is degree-generic;
// every reference case uses Q1/Q1, which keeps its own kernel in scnsim.cu). Correctness first: one CTA per cell, the
// per-quadrature-point state and the physical shape gradients of both spaces staged in shared memory by the first n_q threads,
// then one task per (test node, trial node) pair of each of the four blocks - the integrands are those of scnsim.cu with the
// velocity shapes (N^u, grad N^u) and the pressure shapes (N^p, grad N^p) told apart. Cells of one colour per launch.
#include "scnsim.h"

namespace ifem
{
  namespace
  {
    constexpr double kCpToCv = 1.4, kAtm = 1013250, kKappaS = 1e4; // mpi_scnsim.cpp:124-126
    constexpr int kMaxH = 32;

    template <int DIM>
    struct QState
    {
      double JxW;
      double u[DIM], dv[DIM], G[DIM * DIM], p, dp, gradp[DIM], divu, sdiv[DIM];
      double u_gradu[DIM], gradu_u[DIM], res[DIM], g_bf[DIM], acc[DIM], fsis[DIM * DIM];
      double rho, visc, sigma, tau_supg, tau_pspg, tau_lsic;
    };

    struct GenArgs
    {
      int n_list;
      const int *cell_list, *cell_un, *cell_pn, *indicator;
      const double *cell_x;
      const double *tN, *tdN, *tNp, *tdNp, *tdG, *tqw; // N[nq][nu], dN[nq][nu][dim], Np[nq][np], dNp[nq][np][dim], dNgeo[nq][nv][dim], qw[nq]
      const unsigned char *slots, *con;
      const double *eval_pt, *present, *fsi_acc, *stress, *fsi_stress, *sigma_pml, *body_force, *inhom;
      int64_t n_u;
      int nu, np, nq, nv, n_unodes, n_owned_u, n_owned_p;
      int n_h, h_type[kMaxH], h_node[kMaxH]; // the first dofs_per_cell / dofs_per_vertex system shape functions (:251-257)
      double mu, rho_f, rho_s, dt, grav[3];
      const int64_t *uu_rp, *up_rp, *pu_rp, *pp_rp;
      double *uu, *up, *pu, *pp, *rhs;
    };

    template <int DIM>
    __device__ __forceinline__ double dotd(const double *a, const double *b)
    {
      double s = 0.0;
#pragma unroll
      for (int d = 0; d < DIM; ++d) s = fma(a[d], b[d], s);
      return s;
    }

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

    // everything that depends on the quadrature point only (mpi_scnsim.cpp:153-289)
    template <int DIM>
    __device__ void fill_qpoint(const GenArgs &A, int cell, int q, int ind, QState<DIM> &Q, double *gU /*[nu][DIM]*/, double *gP /*[np][DIM]*/)
    {
      const int nu = A.nu, np = A.np, nv = A.nv;
      const double *X = A.cell_x + (int64_t)cell * nv * DIM;
      double J[DIM * DIM], Ji[DIM * DIM], det;
#pragma unroll
      for (int i = 0; i < DIM * DIM; ++i) J[i] = 0.0;
      for (int v = 0; v < nv; ++v)
#pragma unroll
        for (int i = 0; i < DIM; ++i)
#pragma unroll
          for (int j = 0; j < DIM; ++j) J[i * DIM + j] = fma(X[v * DIM + i], A.tdG[((int64_t)q * nv + v) * DIM + j], J[i * DIM + j]);
      invert<DIM>(J, Ji, det);
      Q.JxW = det * A.tqw[q];
      double up[DIM], pp = 0.0;
#pragma unroll
      for (int c = 0; c < DIM; ++c) Q.u[c] = up[c] = Q.acc[c] = Q.gradp[c] = Q.sdiv[c] = 0.0;
#pragma unroll
      for (int i = 0; i < DIM * DIM; ++i) Q.G[i] = Q.fsis[i] = 0.0;
      Q.p = 0.0;
      for (int b = 0; b < nu; ++b)
        {
          const double N = A.tN[(int64_t)q * nu + b];
          double *g = gU + b * DIM;
#pragma unroll
          for (int k = 0; k < DIM; ++k)
            {
              double s = 0.0;
#pragma unroll
              for (int j = 0; j < DIM; ++j) s = fma(A.tdN[((int64_t)q * nu + b) * DIM + j], Ji[j * DIM + k], s);
              g[k] = s;
            }
          const int un = A.cell_un[(int64_t)cell * nu + b];
#pragma unroll
          for (int c = 0; c < DIM; ++c)
            {
              const double ue = A.eval_pt[(int64_t)DIM * un + c];
              Q.u[c] = fma(N, ue, Q.u[c]);
              up[c] = fma(N, A.present[(int64_t)DIM * un + c], up[c]);
              if (A.fsi_acc) Q.acc[c] = fma(N, A.fsi_acc[(int64_t)DIM * un + c], Q.acc[c]);
#pragma unroll
              for (int k = 0; k < DIM; ++k) Q.G[c * DIM + k] = fma(ue, g[k], Q.G[c * DIM + k]);
            }
          if (A.stress)
#pragma unroll
            for (int i = 0; i < DIM; ++i)
#pragma unroll
              for (int j = 0; j < DIM; ++j) Q.sdiv[i] = fma(A.stress[(int64_t)(i * DIM + j) * A.n_unodes + un], g[j], Q.sdiv[i]);
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
      for (int b = 0; b < np; ++b)
        {
          const double N = A.tNp[(int64_t)q * np + b];
          double *g = gP + b * DIM;
#pragma unroll
          for (int k = 0; k < DIM; ++k)
            {
              double s = 0.0;
#pragma unroll
              for (int j = 0; j < DIM; ++j) s = fma(A.tdNp[((int64_t)q * np + b) * DIM + j], Ji[j * DIM + k], s);
              g[k] = s;
            }
          const int pn = A.cell_pn[(int64_t)cell * np + b];
          const double pe = A.eval_pt[A.n_u + pn];
          Q.p = fma(N, pe, Q.p);
          pp = fma(N, A.present[A.n_u + pn], pp);
#pragma unroll
          for (int c = 0; c < DIM; ++c) Q.gradp[c] = fma(pe, g[c], Q.gradp[c]);
        }
      Q.dp = Q.p - pp;
      Q.sigma = A.sigma_pml ? A.sigma_pml[(int64_t)cell * A.nq + q] : 0.0;
      Q.rho = A.rho_f * (1 + pp / kAtm) * (1 - ind) + ind * A.rho_s; // :210-213
      Q.visc = ind == 1 ? 1.0 : A.mu;                                  // :214-216 (no turbulence model)
      // UGN stabilisation parameters (:247-274) from the previous-step velocity
      double h = 0.0;
      for (int k = 0; k < A.n_h; ++k) h += fabs(dotd<DIM>(up, (A.h_type[k] == 0 ? gU : gP) + A.h_node[k] * DIM));
      const double v_norm = sqrt(dotd<DIM>(up, up));
      h = h != 0.0 ? 2 * v_norm / h : 0.0;
      const double nu_k = Q.visc / Q.rho;
      if (h != 0.0)
        {
          const double t1 = 2 / A.dt, t2 = 2 * v_norm / h, t3 = 4 * nu_k / (h * h);
          Q.tau_supg = 1 / sqrt(t1 * t1 + t2 * t2 + t3 * t3);
        }
      else
        Q.tau_supg = A.dt / 2;
      Q.tau_pspg = Q.tau_supg / Q.rho;
      const double localRe = v_norm * h / (2 * nu_k);
      Q.tau_lsic = h / 2 * v_norm * (localRe <= 3 ? localRe / 3 : 1.0);
      Q.divu = 0.0;
#pragma unroll
      for (int c = 0; c < DIM; ++c)
        {
          Q.divu += Q.G[c * DIM + c];
          Q.dv[c] = Q.u[c] - up[c];
          Q.sdiv[c] *= Q.visc / A.mu; // :288
          Q.g_bf[c] = A.grav[c] + (A.body_force ? A.body_force[((int64_t)cell * A.nq + q) * DIM + c] : 0.0);
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
#pragma unroll
      for (int d = 0; d < DIM; ++d)
        Q.res[d] = Q.rho * (Q.dv[d] / A.dt + Q.u_gradu[d]) + Q.gradp[d] - Q.sdiv[d] - Q.rho * Q.g_bf[d] + Q.rho * Q.sigma * Q.u[d];
    }

    template <int DIM>
    __global__ void __launch_bounds__(256) scns_generic_kernel(const GenArgs A)
    {
      extern __shared__ double smem[];
      const int nu = A.nu, np = A.np, nq = A.nq, dpc = nu * DIM + np;
      constexpr int QS = (sizeof(QState<DIM>) + 7) / 8;
      QState<DIM> *sq = reinterpret_cast<QState<DIM> *>(smem);
      double *gU = smem + (size_t)nq * QS;            // [nq][nu][DIM]
      double *gP = gU + (size_t)nq * nu * DIM;        // [nq][np][DIM]
      double *lrhs = gP + (size_t)nq * np * DIM;      // [dpc]: velocity (a * DIM + c), then pressure
      double *ldiag = lrhs + dpc;
      const int cell = A.cell_list[blockIdx.x];
      const int ind = A.indicator ? A.indicator[cell] : 0;
      for (int q = threadIdx.x; q < nq; q += blockDim.x) fill_qpoint<DIM>(A, cell, q, ind, sq[q], gU + (size_t)q * nu * DIM, gP + (size_t)q * np * DIM);
      for (int k = threadIdx.x; k < dpc; k += blockDim.x) lrhs[k] = ldiag[k] = 0.0;
      __syncthreads();
      const double dt = A.dt, cp = kCpToCv, atm = kAtm, ks = kKappaS, om = 1.0 - ind;
      const unsigned char *slots = A.slots + (int64_t)cell * (nu * nu + 2 * nu * np + np * np);
      const int n_uu = nu * nu, n_up = nu * np, n_pu = np * nu, n_pp = np * np;
      for (int task = threadIdx.x; task < n_uu + n_up + n_pu + n_pp; task += blockDim.x)
        {
          if (task < n_uu)
            {
              // ---- velocity test (a, c) x velocity trial (b, d) (+ the velocity rhs when b == 0) ----
              const int a = task / nu, b = task % nu;
              double K[DIM][DIM], r[DIM];
#pragma unroll
              for (int c = 0; c < DIM; ++c)
                {
                  r[c] = 0.0;
#pragma unroll
                  for (int d = 0; d < DIM; ++d) K[c][d] = 0.0;
                }
              for (int q = 0; q < nq; ++q)
                {
                  const QState<DIM> &Q = sq[q];
                  const double w = Q.JxW, Na = A.tN[(int64_t)q * nu + a], Nb = A.tN[(int64_t)q * nu + b], rho = Q.rho, ts = Q.tau_supg, tl = Q.tau_lsic;
                  const double *ga = gU + ((size_t)q * nu + a) * DIM, *gb = gU + ((size_t)q * nu + b) * DIM;
                  const double gagb = dotd<DIM>(ga, gb), ugb = dotd<DIM>(Q.u, gb);
                  const double ga_ugu = dotd<DIM>(ga, Q.u_gradu), ga_dv = dotd<DIM>(ga, Q.dv), ga_gp = dotd<DIM>(ga, Q.gradp),
                               ga_sd = dotd<DIM>(ga, Q.sdiv), ga_bf = dotd<DIM>(ga, Q.g_bf), ga_u = dotd<DIM>(ga, Q.u), ga_acc = dotd<DIM>(ga, Q.acc);
                  const double same = ts * rho * Nb * ga_ugu + ts * rho * Nb * ga_dv / dt + ts * Nb * ga_gp - ts * Nb * ga_sd - ts * Nb * ga_bf * rho +
                                      ts * rho * Nb * ga_u * Q.sigma - (ind == 1 ? ts * Nb * ga_acc * rho : 0.0);
#pragma unroll
                  for (int c = 0; c < DIM; ++c)
                    {
                      const double uc = Q.u[c];
#pragma unroll
                      for (int d = 0; d < DIM; ++d)
                        {
                          double m = rho * Q.G[c * DIM + d] * Nb * Na;
                          m += ts * rho * uc * Nb * dotd<DIM>(ga, &Q.G[d * DIM]);
                          m += ts * rho * uc * Q.u[d] * gagb;
                          m += ts * rho * uc * ga[d] * Nb / dt;
                          m += ts * rho * uc * ga[d] * Nb * Q.sigma;
                          m += tl * rho * cp * ga[c] * gb[d] * (1.0 + Q.p * om / atm);
                          m += tl * rho * ga[c] * Nb * Q.gradp[d] / atm * om;
                          if (c == d) m += Q.visc * gagb + rho * ugb * Na + rho * Na * Nb / dt + rho * Q.sigma * Nb * Na + same;
                          K[c][d] = fma(m, w, K[c][d]);
                        }
                      if (b == 0)
                        {
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
                }
              const int nA = A.cell_un[(int64_t)cell * nu + a], nB = A.cell_un[(int64_t)cell * nu + b];
              if (nA < A.n_owned_u)
                {
                  const int64_t r0 = A.uu_rp[nA];
                  const int rn = (int)(A.uu_rp[nA + 1] - r0), s = slots[a * nu + b];
#pragma unroll
                  for (int c = 0; c < DIM; ++c)
                    {
                      if (A.con[(int64_t)DIM * nA + c])
                        {
                          if (a == b)
                            {
                              const double dv = fabs(K[c][c]);
                              A.uu[r0 * DIM * DIM + (int64_t)(c * DIM + c) * rn + s] += dv;
                              ldiag[a * DIM + c] = dv;
                            }
                          continue;
                        }
                      double corr = 0.0;
#pragma unroll
                      for (int d = 0; d < DIM; ++d)
                        {
                          if (A.con[(int64_t)DIM * nB + d])
                            {
                              if (A.inhom) corr = fma(K[c][d], A.inhom[(int64_t)DIM * nB + d], corr);
                              continue;
                            }
                          A.uu[r0 * DIM * DIM + (int64_t)(c * DIM + d) * rn + s] += K[c][d];
                        }
                      const double add = (b == 0 ? r[c] : 0.0) - corr;
                      if (add != 0.0) atomicAdd(&lrhs[a * DIM + c], add);
                    }
                }
            }
          else if (task < n_uu + n_up)
            {
              // ---- velocity test (a, c) x pressure trial b ----
              const int t = task - n_uu, a = t / np, b = t % np;
              double K[DIM];
#pragma unroll
              for (int c = 0; c < DIM; ++c) K[c] = 0.0;
              for (int q = 0; q < nq; ++q)
                {
                  const QState<DIM> &Q = sq[q];
                  const double w = Q.JxW, Nb = A.tNp[(int64_t)q * np + b], rho = Q.rho, ts = Q.tau_supg, tl = Q.tau_lsic;
                  const double *ga = gU + ((size_t)q * nu + a) * DIM, *gb = gP + ((size_t)q * np + b) * DIM;
                  const double gagb = dotd<DIM>(ga, gb), ugb = dotd<DIM>(Q.u, gb);
#pragma unroll
                  for (int c = 0; c < DIM; ++c)
                    {
                      double m = -ga[c] * Nb + ts * Q.u[c] * gagb;
                      m += tl * rho * ga[c] * Nb / dt * om / atm + tl * rho / ks * ga[c] * Nb / dt * ind;
                      m += tl * rho * cp * ga[c] * Nb * om * Q.divu / atm + tl * rho * ga[c] * ugb / atm * om;
                      K[c] = fma(m, w, K[c]);
                    }
                }
              const int nA = A.cell_un[(int64_t)cell * nu + a], nB = A.cell_pn[(int64_t)cell * np + b];
              if (nA < A.n_owned_u)
                {
                  const int64_t r0 = A.up_rp[nA];
                  const int rn = (int)(A.up_rp[nA + 1] - r0), s = slots[n_uu + a * np + b];
                  const bool ccon = A.con[A.n_u + nB];
#pragma unroll
                  for (int c = 0; c < DIM; ++c)
                    {
                      if (A.con[(int64_t)DIM * nA + c]) continue;
                      if (ccon)
                        {
                          if (A.inhom && K[c] != 0.0) atomicAdd(&lrhs[a * DIM + c], -K[c] * A.inhom[A.n_u + nB]);
                        }
                      else
                        A.up[r0 * DIM + (int64_t)c * rn + s] += K[c];
                    }
                }
            }
          else if (task < n_uu + n_up + n_pu)
            {
              // ---- pressure test a x velocity trial (b, c) ----
              const int t = task - n_uu - n_up, a = t / nu, b = t % nu;
              double K[DIM];
#pragma unroll
              for (int c = 0; c < DIM; ++c) K[c] = 0.0;
              for (int q = 0; q < nq; ++q)
                {
                  const QState<DIM> &Q = sq[q];
                  const double w = Q.JxW, Na = A.tNp[(int64_t)q * np + a], Nb = A.tN[(int64_t)q * nu + b], rho = Q.rho, tp = Q.tau_pspg;
                  const double *ga = gP + ((size_t)q * np + a) * DIM, *gb = gU + ((size_t)q * nu + b) * DIM;
                  const double gagb = dotd<DIM>(ga, gb);
#pragma unroll
                  for (int c = 0; c < DIM; ++c)
                    {
                      double m = tp * rho * Nb * dotd<DIM>(ga, &Q.G[c * DIM]) + tp * rho * Q.u[c] * gagb + tp * rho * ga[c] * Nb / dt +
                                 tp * rho * ga[c] * Nb * Q.sigma;
                      m += (cp * (atm + Q.p * om) * gb[c] * Na + Nb * Q.gradp[c] * Na * om) / atm;
                      K[c] = fma(m, w, K[c]);
                    }
                }
              const int nA = A.cell_pn[(int64_t)cell * np + a], nB = A.cell_un[(int64_t)cell * nu + b];
              if (nA < A.n_owned_p && !A.con[A.n_u + nA])
                {
                  const int64_t r0 = A.pu_rp[nA];
                  const int rn = (int)(A.pu_rp[nA + 1] - r0), s = slots[n_uu + n_up + a * nu + b];
                  double corr = 0.0;
#pragma unroll
                  for (int c = 0; c < DIM; ++c)
                    {
                      if (A.con[(int64_t)DIM * nB + c])
                        {
                          if (A.inhom) corr = fma(K[c], A.inhom[(int64_t)DIM * nB + c], corr);
                          continue;
                        }
                      A.pu[r0 * DIM + (int64_t)c * rn + s] += K[c];
                    }
                  if (corr != 0.0) atomicAdd(&lrhs[nu * DIM + a], -corr);
                }
            }
          else
            {
              // ---- pressure test a x pressure trial b (+ the pressure rhs when b == 0) ----
              const int t = task - n_uu - n_up - n_pu, a = t / np, b = t % np;
              double K = 0.0, r = 0.0;
              for (int q = 0; q < nq; ++q)
                {
                  const QState<DIM> &Q = sq[q];
                  const double w = Q.JxW, Na = A.tNp[(int64_t)q * np + a], Nb = A.tNp[(int64_t)q * np + b], rho = Q.rho, tp = Q.tau_pspg;
                  const double *ga = gP + ((size_t)q * np + a) * DIM, *gb = gP + ((size_t)q * np + b) * DIM;
                  const double gagb = dotd<DIM>(ga, gb), ugb = dotd<DIM>(Q.u, gb);
                  double m = Q.sigma * Nb * Na / atm + tp * gagb;
                  m += (Nb * Q.divu * Na * om + ugb * Na * om + Na * Nb / dt * om) / atm + 1 / ks * Na * Nb * ind / dt;
                  K = fma(m, w, K);
                  if (b == 0)
                    {
                      double v = -Q.sigma * Q.p * Na / atm;
                      v += -(cp * (atm + Q.p * om) * Q.divu * Na + dotd<DIM>(Q.u, Q.gradp) * Na * om + Q.dp * Na / dt * om) / atm -
                           1 / ks * Q.dp * Na * ind / dt;
                      v += -tp * dotd<DIM>(ga, Q.res);
                      if (ind == 1) v += rho * tp * dotd<DIM>(ga, Q.acc);
                      r = fma(v, w, r);
                    }
                }
              const int nA = A.cell_pn[(int64_t)cell * np + a], nB = A.cell_pn[(int64_t)cell * np + b];
              if (nA < A.n_owned_p)
                {
                  const int64_t r0 = A.pp_rp[nA];
                  const int s = slots[n_uu + n_up + n_pu + a * np + b];
                  if (A.con[A.n_u + nA])
                    {
                      if (a == b)
                        {
                          A.pp[r0 + s] += fabs(K);
                          ldiag[nu * DIM + a] = fabs(K);
                        }
                    }
                  else
                    {
                      double add = b == 0 ? r : 0.0;
                      if (A.con[A.n_u + nB])
                        {
                          if (A.inhom) add -= K * A.inhom[A.n_u + nB];
                        }
                      else
                        A.pp[r0 + s] += K;
                      if (add != 0.0) atomicAdd(&lrhs[nu * DIM + a], add);
                    }
                }
            }
        }
      __syncthreads();
      for (int k = threadIdx.x; k < dpc; k += blockDim.x)
        {
          const bool is_u = k < nu * DIM;
          const int node = is_u ? A.cell_un[(int64_t)cell * nu + k / DIM] : A.cell_pn[(int64_t)cell * np + (k - nu * DIM)];
          const bool own = is_u ? node < A.n_owned_u : node < A.n_owned_p;
          if (!own) continue;
          const int64_t g = is_u ? (int64_t)DIM * node + k % DIM : A.n_u + node;
          if (!A.con[g]) A.rhs[g] += lrhs[k];
          else if (A.inhom) A.rhs[g] += ldiag[k] * A.inhom[g];
        }
    }
  } // namespace

  // host side: SCnsIM::assemble for general degrees (called from SCnsIM::assemble when the space is not Q1/Q1)
  void scns_assemble_generic(Context &ctx, FluidSpace &fs, const ScnsGenericInput &in, bool use_nonzero_constraints)
  {
    cudaStream_t s = ctx.stream;
    if (!(fs.pp == 1 || fs.pp == fs.pu)) throw std::runtime_error("SCnsIM: pressure degree must be 1 or equal to the velocity degree");
    GenArgs a{};
    a.cell_un = fs.d_cell_un.p;
    a.cell_pn = fs.d_cell_pn.p;
    a.indicator = fs.d_indicator.p;
    a.cell_x = fs.d_cell_x.p;
    const int nu = fs.nu, np = fs.np, nq = fs.nq, nv = fs.nv, dim = fs.dim;
    // d_tables: N[nq][nu] | dN[nq][nu][dim] | Np[nq][np] | dNgeo[nq][nv][dim] | qw[nq]
    a.tN = fs.d_tables.p;
    a.tdN = a.tN + (size_t)nq * nu;
    a.tNp = a.tdN + (size_t)nq * nu * dim;
    a.tdG = a.tNp + (size_t)nq * np;
    a.tqw = a.tdG + (size_t)nq * nv * dim;
    a.tdNp = fs.pp == 1 ? a.tdG : a.tdN; // FE_Q(1) pressure = the geometry element; equal order = the velocity element
    a.slots = fs.d_slots.p;
    a.con = fs.d_con.p;
    a.eval_pt = in.eval_pt;
    a.present = in.present;
    a.fsi_acc = in.fsi_acc;
    a.stress = in.stress;
    a.fsi_stress = in.fsi_stress;
    a.sigma_pml = in.sigma_pml;
    a.body_force = in.body_force;
    a.inhom = use_nonzero_constraints ? fs.d_nonzero_val.p : nullptr;
    a.n_u = fs.n_u;
    a.nu = nu;
    a.np = np;
    a.nq = nq;
    a.nv = nv;
    a.n_unodes = fs.un.n_nodes;
    a.n_owned_u = fs.n_owned_unodes;
    a.n_owned_p = fs.n_owned_pnodes;
    // the first dofs_per_cell / dofs_per_vertex system shape functions in deal.II's cell-local numbering: vertex dofs come
    // first, per vertex dim velocity components then the pressure (mpi_scnsim.cpp:251-257)
    {
      const int dpc = nu * dim + np, n_h = dpc / (dim + 1);
      if (n_h > kMaxH) throw std::runtime_error("SCnsIM: too many shape functions in the element-length sum");
      a.n_h = n_h;
      int k = 0;
      for (int v = 0; v < nv && k < n_h; ++v)
        {
          int un = 0, pn = 0, su = 1, sp = 1;
          for (int d = 0; d < dim; ++d)
            {
              un += ((v >> d) & 1) * fs.pu * su;
              pn += ((v >> d) & 1) * fs.pp * sp;
              su *= fs.pu + 1;
              sp *= fs.pp + 1;
            }
          for (int c = 0; c < dim && k < n_h; ++c, ++k)
            {
              a.h_type[k] = 0;
              a.h_node[k] = un;
            }
          if (k < n_h)
            {
              a.h_type[k] = 1;
              a.h_node[k] = pn;
              ++k;
            }
        }
    }
    a.mu = in.mu;
    a.rho_f = in.rho_f;
    a.rho_s = in.rho_s;
    a.dt = in.dt;
    for (int d = 0; d < 3; ++d) a.grav[d] = in.grav[d];
    a.uu_rp = fs.A_uu.rowptr.p;
    a.up_rp = fs.A_up.rowptr.p;
    a.pu_rp = fs.A_pu.rowptr.p;
    a.pp_rp = fs.A_pp.rowptr.p;
    a.uu = fs.A_uu.val.p;
    a.up = fs.A_up.val.p;
    a.pu = fs.A_pu.val.p;
    a.pp = fs.A_pp.val.p;
    a.rhs = fs.rhs.p;
    const size_t qs = dim == 2 ? (sizeof(QState<2>) + 7) / 8 : (sizeof(QState<3>) + 7) / 8;
    const size_t shared = ((size_t)nq * qs + (size_t)nq * (nu + np) * dim + 2 * (size_t)(nu * dim + np)) * sizeof(double);
    if (shared > 200 * 1024) throw std::runtime_error("SCnsIM: element too large for the generic assembly kernel");
    static bool opted = false;
    if (!opted)
      {
        IFEM_CUDA(cudaFuncSetAttribute(scns_generic_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        IFEM_CUDA(cudaFuncSetAttribute(scns_generic_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        opted = true;
      }
    const int n_colours = (int)fs.colour_offsets.size() - 1;
    for (int k = 0; k < n_colours; ++k)
      {
        a.n_list = fs.colour_offsets[k + 1] - fs.colour_offsets[k];
        a.cell_list = fs.d_colour_order.p + fs.colour_offsets[k];
        if (!a.n_list) continue;
        if (dim == 2)
          scns_generic_kernel<2><<<a.n_list, 256, shared, s>>>(a);
        else
          scns_generic_kernel<3><<<a.n_list, 256, shared, s>>>(a);
        IFEM_KERNEL_CHECK();
        ctx.kernel_launches++;
      }
  }
} // namespace ifem
