// Kernel bodies of Solid::MPI::[Shared]LinearElasticity (reference source/mpi_linear_elasticity.cpp,
// source/mpi_shared_linear_elasticity.cpp). They use no warp intrinsics and no shared memory, so they are written as
// IFEM_HD functions of (block, thread): solid.cu wraps them in __global__ kernels; tests/cpp/linear_kernels_cpu.cpp
// compiles the same bodies with g++ and walks the launch grid sequentially, so that the arithmetic and the indexing of
// the device code are checked against the oracle on a machine without a GPU (test infrastructure - the product path is
// the CUDA launch).
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
  // ---- LinearElasticity / SharedLinearElasticity ---------------------------------------------------------------
  struct LinearArgs
  {
    int n_list;
    const int *cell_list, *cell_nodes;
    const unsigned char *slots, *con;
    const double *N, *G, *JxW;
    int nq;
    double rho, lambda, mu, eta, grav[3];
    // system = c_mass * M + c_damp * C + c_stiff * K; any of sys / mass / stiff / damp may be null (not assembled)
    double c_mass, c_damp, c_stiff;
    const int64_t *rowptr;
    double *sys, *mass, *stiff, *damp, *rhs;
  };

  // cell loop of [Shared]LinearElasticity::assemble_system (mpi_linear_elasticity.cpp:73-130,
  // mpi_shared_linear_elasticity.cpp:94-152): one thread per (cell, row node a, column node b), cells of one colour per
  // launch. With sym grad phi_(a,c) = sym(e_c x g_a) and C = mu (ik jl + il jk) + lambda ij kl:
  //   sym_(a,c) : C : sym_(b,d) = mu (delta_cd g_a.g_b + g_a[d] g_b[c]) + lambda g_a[c] g_b[d]
  // and with the viscosity tensor eta/2 (ik jl + il jk):  eta/2 (delta_cd g_a.g_b + g_a[d] g_b[c]).
  template <int DIM, int NPC>
  IFEM_HD void linear_assemble_body(const LinearArgs &A, int block, int thread)
  {
    using std::fma;
    using std::fabs;
    constexpr int PAIRS = NPC * NPC, CPB = 64 / PAIRS;
    const int li = block * CPB + thread / PAIRS;
    if (li >= A.n_list) return;
    const int cell = A.cell_list[li];
    const int pr = thread % PAIRS, a = pr / NPC, b = pr % NPC;
    double m = 0.0, gg = 0.0, outer[DIM * DIM], r[DIM];
#pragma unroll
    for (int i = 0; i < DIM * DIM; ++i) outer[i] = 0.0;
#pragma unroll
    for (int i = 0; i < DIM; ++i) r[i] = 0.0;
    for (int q = 0; q < A.nq; ++q)
      {
        const int64_t cq = (int64_t)cell * A.nq + q;
        const double w = A.JxW[cq];
        const double *g0 = A.G + cq * NPC * DIM;
        const double Na = A.N[q * NPC + a], Nb = A.N[q * NPC + b];
        m = fma(A.rho * Na * Nb, w, m);
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < DIM; ++k) s = fma(g0[a * DIM + k], g0[b * DIM + k], s);
        gg = fma(s, w, gg);
#pragma unroll
        for (int c = 0; c < DIM; ++c)
#pragma unroll
          for (int d = 0; d < DIM; ++d) outer[c * DIM + d] = fma(g0[a * DIM + c] * g0[b * DIM + d], w, outer[c * DIM + d]);
        if (b == 0)
#pragma unroll
          for (int c = 0; c < DIM; ++c) r[c] = fma(w, A.rho * A.grav[c] * Na, r[c]);
      }
    const int nA = A.cell_nodes[(int64_t)cell * NPC + a], nB = A.cell_nodes[(int64_t)cell * NPC + b];
    const int64_t rp = A.rowptr[nA];
    const int nb = (int)(A.rowptr[nA + 1] - rp);
    const int slot = A.slots[(int64_t)cell * PAIRS + pr];
    const int64_t base = rp * DIM * DIM + slot;
#pragma unroll
    for (int c = 0; c < DIM; ++c)
      {
        const int rc = A.con[(int64_t)DIM * nA + c];
#pragma unroll
        for (int d = 0; d < DIM; ++d)
          {
            const int cc = A.con[(int64_t)DIM * nB + d];
            const double sym = (c == d ? gg : 0.0) + outer[d * DIM + c];
            const double ke = A.mu * sym + A.lambda * outer[c * DIM + d];
            const double ce = 0.5 * A.eta * sym;
            const double me = c == d ? m : 0.0;
            const double se = A.c_mass * me + A.c_damp * ce + A.c_stiff * ke;
            const int64_t at = base + (int64_t)(c * DIM + d) * nb;
            if (rc)
              {
                // constrained row: distribute_local_to_global keeps |local diagonal| on the diagonal
                if (a == b && c == d)
                  {
                    if (A.sys) A.sys[at] += fabs(se);
                    if (A.mass) A.mass[at] += fabs(me);
                    if (A.stiff) A.stiff[at] += fabs(ke);
                    if (A.damp) A.damp[at] += fabs(ce);
                  }
              }
            else if (!cc)
              {
                if (A.sys) A.sys[at] += se;
                if (A.mass) A.mass[at] += me;
                if (A.stiff) A.stiff[at] += ke;
                if (A.damp) A.damp[at] += ce;
              }
          }
        if (b == 0 && !rc && A.rhs) A.rhs[(int64_t)DIM * nA + c] += r[c];
      }
  }

  // SharedLinearElasticity::update_strain_and_stress (mpi_shared_linear_elasticity.cpp:401-531): one thread per cell
  // (cells of one colour per launch): strain = sym grad u, stress = C : strain at the quadrature points -> qpt_to_dof ->
  // nodal scatter-add; the average over the surrounding cells is taken by solid_average_kernel
  template <int DIM, int NPC>
  IFEM_HD void linear_stress_body(int li, int n_list, const int *cell_list, int nq, const int *cell_nodes, const double *qpt_to_dof,
                                  const double *G, const double *u, double lambda, double mu, int n_nodes, double *stress,
                                  double *strain, double *count)
  {
    using std::fma;
    if (li >= n_list) return;
    const int cell = cell_list[li];
    double ue[NPC * DIM];
    for (int b = 0; b < NPC; ++b)
      for (int c = 0; c < DIM; ++c) ue[b * DIM + c] = u[(int64_t)DIM * cell_nodes[(int64_t)cell * NPC + b] + c];
    for (int a = 0; a < NPC; ++a)
      {
        double st[DIM * DIM], sn[DIM * DIM];
#pragma unroll
        for (int i = 0; i < DIM * DIM; ++i) st[i] = sn[i] = 0.0;
        for (int q = 0; q < nq; ++q)
          {
            const int64_t cq = (int64_t)cell * nq + q;
            const double *g = G + cq * NPC * DIM;
            double gu[DIM * DIM];
#pragma unroll
            for (int i = 0; i < DIM * DIM; ++i) gu[i] = 0.0;
            for (int b = 0; b < NPC; ++b)
#pragma unroll
              for (int c = 0; c < DIM; ++c)
#pragma unroll
                for (int k = 0; k < DIM; ++k) gu[c * DIM + k] = fma(ue[b * DIM + c], g[b * DIM + k], gu[c * DIM + k]);
            double tr = 0.0;
#pragma unroll
            for (int i = 0; i < DIM; ++i) tr += gu[i * DIM + i];
            const double w = qpt_to_dof[a * nq + q];
#pragma unroll
            for (int i = 0; i < DIM; ++i)
#pragma unroll
              for (int j = 0; j < DIM; ++j)
                {
                  const double e = 0.5 * (gu[i * DIM + j] + gu[j * DIM + i]);
                  sn[i * DIM + j] = fma(w, e, sn[i * DIM + j]);
                  st[i * DIM + j] = fma(w, 2.0 * mu * e + (i == j ? lambda * tr : 0.0), st[i * DIM + j]);
                }
          }
        const int node = cell_nodes[(int64_t)cell * NPC + a];
#pragma unroll
        for (int i = 0; i < DIM * DIM; ++i)
          {
            stress[(int64_t)i * n_nodes + node] += st[i];
            strain[(int64_t)i * n_nodes + node] += sn[i];
          }
        count[node] += 1.0;
      }
  }
} // namespace ifem
