// Kernel bodies of Fluid::MPI::SUPGInsIM<dim>::assemble (reference source/mpi_insim_supg.cpp:15-328) for equal-order Q1/Q1
// elements: incompressible Navier-Stokes with SUPG / PSPG / LSIC stabilisation. Work decomposition as in the SCnsIM
// kernel (scnsim.cu): CTA = CPB cells, per cell NU*NU threads, thread (a, b) owns the (dim+1) x (dim+1) block coupling
// test node a with trial node b; the state of every quadrature point is computed once by the first NQ threads of the cell
// and staged in shared memory.
//
// The three phases are written as IFEM_HD functions of (cell, thread-in-cell) on explicit staging arrays, without warp
// intrinsics: insim_supg.cu wraps them in the __global__ kernel (phases separated by __syncthreads), and
// tests/cpp/supg_kernels_cpu.cpp compiles the same bodies with g++ and walks the launch grid phase by phase so that the
// arithmetic, the row-plane BCSR indexing and the constrained scatter are checked against the oracle on a machine without
// a GPU (test infrastructure - the product path is the CUDA launch).
#pragma once
#include <cmath>
#include <cstdint>

#ifdef __CUDACC__
#define IFEM_HD __host__ __device__ __forceinline__
#else
#define IFEM_HD inline
#endif

namespace ifem
{
  struct SupgArgs
  {
    int n_list;
    const int *cell_list, *cell_un, *cell_pn;
    const double *cell_x, *tables; // N[nq][nu] | dN[nq][nu][dim] | Np[nq][np] | dNgeo[nq][nv][dim] | qw[nq]
    const unsigned char *slots, *con;
    const double *eval_pt, *present, *body_force, *inhom; // body_force [cell][q][dim] or null; inhom null: zero constraints
    int64_t n_u;
    int n_owned_u, n_owned_p, n_h, h_node[8];
    double mu, rho, dt, grav[3];
    const int64_t *uu_rp, *up_rp, *pu_rp, *pp_rp;
    double *uu, *up, *pu, *pp, *rhs;
  };

  template <int DIM>
  struct SupgQPoint
  {
    static constexpr int NU = 1 << DIM;
    double JxW, N[NU], g[NU][DIM];
    double u[DIM], dv[DIM], G[DIM * DIM], p, gradp[DIM], divu;
    double u_gradu[DIM], gradu_u[DIM], res[DIM], g_bf[DIM];
    double tau_supg, tau_pspg, tau_lsic;
  };

  namespace supg_detail
  {
    template <int DIM>
    IFEM_HD void invert(const double *J, double *Ji, double &det)
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
    IFEM_HD double dotd(const double *a, const double *b)
    {
      double s = 0.0;
#pragma unroll
      for (int d = 0; d < DIM; ++d) s = std::fma(a[d], b[d], s);
      return s;
    }

    IFEM_HD void shared_add(double *p, double v)
    {
#ifdef __CUDA_ARCH__
      atomicAdd(p, v);
#else
      *p += v;
#endif
    }
  } // namespace supg_detail

  // phase 1, thread q < NQ of the cell: everything that depends on the quadrature point only (:84-152)
  template <int DIM>
  IFEM_HD void supg_fill_qpoint(const SupgArgs &A, int cell, int q, SupgQPoint<DIM> &Q)
  {
    using namespace supg_detail;
    using std::fma;
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
    double up[DIM];
#pragma unroll
    for (int c = 0; c < DIM; ++c) Q.u[c] = up[c] = Q.gradp[c] = 0.0;
#pragma unroll
    for (int i = 0; i < DIM * DIM; ++i) Q.G[i] = 0.0;
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
#pragma unroll
        for (int c = 0; c < DIM; ++c)
          {
            const double ue = A.eval_pt[(int64_t)DIM * un + c];
            Q.u[c] = fma(N, ue, Q.u[c]);
            up[c] = fma(N, A.present[(int64_t)DIM * un + c], up[c]);
            Q.gradp[c] = fma(pe, Q.g[b][c], Q.gradp[c]);
#pragma unroll
            for (int k = 0; k < DIM; ++k) Q.G[c * DIM + k] = fma(ue, Q.g[b][k], Q.G[c * DIM + k]);
          }
      }
    // UGN stabilisation parameters from the previous-step velocity (:122-152)
    double h = 0.0;
    for (int k = 0; k < A.n_h; ++k) h += std::fabs(dotd<DIM>(up, Q.g[A.h_node[k]]));
    const double v_norm = std::sqrt(dotd<DIM>(up, up));
    h = h != 0.0 ? 2 * v_norm / h : 0.0;
    const double nu = A.mu / A.rho;
    if (h != 0.0)
      {
        const double t1 = 2 / A.dt, t2 = 2 * v_norm / h, t3 = 4 * nu / (h * h);
        Q.tau_supg = 1 / std::sqrt(t1 * t1 + t2 * t2 + t3 * t3);
      }
    else
      Q.tau_supg = A.dt / 2;
    Q.tau_pspg = Q.tau_supg / A.rho;
    const double localRe = v_norm * h / (2 * nu);
    Q.tau_lsic = h / 2 * v_norm * (localRe <= 3 ? localRe / 3 : 1.0);
    Q.divu = 0.0;
#pragma unroll
    for (int c = 0; c < DIM; ++c)
      {
        Q.divu += Q.G[c * DIM + c];
        Q.dv[c] = Q.u[c] - up[c];
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
    // strong momentum residual used by the SUPG / PSPG right-hand sides (:257-279)
#pragma unroll
    for (int d = 0; d < DIM; ++d) Q.res[d] = A.rho * (Q.dv[d] / A.dt + Q.u_gradu[d]) + Q.gradp[d] - A.rho * Q.g_bf[d];
  }

  // phase 2, thread pr = a * NU + b of the cell: the (dim+1) x (dim+1) block of the pair over all quadrature points
  // (:154-285), scattered through the constraints (distribute_local_to_global, :323-332) into the four row-plane BCSR
  // blocks; right-hand-side contributions are collected in lrhs / ldiag [NU * (DIM+1)] of the cell
  template <int DIM>
  IFEM_HD void supg_pair_body(const SupgArgs &A, int cell, int pr, const SupgQPoint<DIM> *sq, double *lrhs, double *ldiag)
  {
    using namespace supg_detail;
    using std::fma;
    constexpr int NU = 1 << DIM, NQ = NU, PAIRS = NU * NU, D1 = DIM + 1;
    const int a = pr / NU, b = pr % NU;
    const double dt = A.dt, rho = A.rho, mu = A.mu;
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
        const SupgQPoint<DIM> &Q = sq[q];
        const double w = Q.JxW, Na = Q.N[a], Nb = Q.N[b], ts = Q.tau_supg, tp = Q.tau_pspg, tl = Q.tau_lsic;
        const double *ga = Q.g[a], *gb = Q.g[b];
        const double gagb = dotd<DIM>(ga, gb), ugb = dotd<DIM>(Q.u, gb);
        // the terms with "phi_u[j] * grad_phi_u[i]" act only when trial and test components agree
        const double same = ts * rho * Nb * dotd<DIM>(ga, Q.u_gradu) + ts * rho * Nb * dotd<DIM>(ga, Q.dv) / dt +
                            ts * Nb * dotd<DIM>(ga, Q.gradp) - ts * Nb * dotd<DIM>(ga, Q.g_bf) * rho;
#pragma unroll
        for (int c = 0; c < DIM; ++c)
          {
            const double uc = Q.u[c];
            // velocity test (a, c) x velocity trial (b, d)
#pragma unroll
            for (int d = 0; d < DIM; ++d)
              {
                double m = rho * Q.G[c * DIM + d] * Nb * Na;            // (grad u phi_j) . phi_i
                m += ts * rho * uc * Nb * dotd<DIM>(ga, &Q.G[d * DIM]);  // SUPG: (u grad phi_i) . (phi_j grad u)
                m += ts * rho * uc * Q.u[d] * gagb;                      // SUPG: (u grad phi_i) . (u grad phi_j)
                m += ts * rho * uc * ga[d] * Nb / dt;                    // SUPG acceleration
                m += tl * rho * ga[c] * gb[d];                           // LSIC
                if (c == d) m += mu * gagb + rho * ugb * Na + rho * Na * Nb / dt + same;
                K[c][d] = fma(m, w, K[c][d]);
              }
            // velocity test (a, c) x pressure trial b: -div phi_i psi_j + SUPG pressure
            K[c][DIM] = fma(-ga[c] * Nb + ts * uc * gagb, w, K[c][DIM]);
            // pressure test a x velocity trial (b, c): PSPG convection / acceleration + continuity div phi_j psi_i
            K[DIM][c] = fma(tp * rho * Nb * dotd<DIM>(ga, &Q.G[c * DIM]) + tp * rho * uc * gagb + tp * rho * ga[c] * Nb / dt + gb[c] * Na, w,
                            K[DIM][c]);
            if (b == 0)
              {
                double v = -mu * dotd<DIM>(&Q.G[c * DIM], ga) - rho * Q.gradu_u[c] * Na + Q.p * ga[c] - rho * Q.dv[c] * Na / dt +
                           Q.g_bf[c] * Na * rho;
                v += -ts * uc * dotd<DIM>(ga, Q.res);
                v += -tl * rho * ga[c] * Q.divu;
                r[c] = fma(v, w, r[c]);
              }
          }
        K[DIM][DIM] = fma(tp * gagb, w, K[DIM][DIM]); // PSPG pressure
        if (b == 0) r[DIM] = fma(-Q.divu * Na - tp * dotd<DIM>(ga, Q.res), w, r[DIM]);
      }
    // ---- scatter through the constraints ----
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
                // constrained row: |local diagonal| on the diagonal, remembered for rhs = diag * inhomogeneity
                if (a == b && i == j)
                  {
                    const double dv = std::fabs(v);
                    if (i < DIM) A.uu[uu0 * DIM * DIM + (int64_t)(i * DIM + j) * uun + s_uu] += dv;
                    else A.pp[pp0 + s_pp] += dv;
                    ldiag[a * D1 + i] = dv;
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
            if (add != 0.0) shared_add(&lrhs[a * D1 + i], add);
          }
      }
  }

  // phase 3, thread pr < NU * (DIM+1) of the cell: local right-hand side to the global vector
  template <int DIM>
  IFEM_HD void supg_rhs_body(const SupgArgs &A, int cell, int pr, const double *lrhs, const double *ldiag)
  {
    constexpr int NU = 1 << DIM, D1 = DIM + 1;
    const int aa = pr / D1, i = pr % D1;
    const int nu_ = A.cell_un[(int64_t)cell * NU + aa], np_ = A.cell_pn[(int64_t)cell * NU + aa];
    const bool own = i < DIM ? nu_ < A.n_owned_u : np_ < A.n_owned_p;
    const int64_t g = i < DIM ? (int64_t)DIM * nu_ + i : A.n_u + np_;
    if (!own) return;
    if (!A.con[g]) A.rhs[g] += lrhs[pr];
    else if (A.inhom) A.rhs[g] += ldiag[pr] * A.inhom[g];
  }
} // namespace ifem
