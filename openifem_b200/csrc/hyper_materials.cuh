// Material point update of Solid::MPI::HyperElasticity for solid_type = Kirchhoff (reference
// include/kirchhoff_elastic_material.h:36-76 through PointHistory::update, source/mpi_hyper_elasticity.cpp:37-65):
//   F = I + Grad u,  E = (F^T F - I) / 2,  S = lambda tr(E) I + 2 mu E,  tau = F S F^T (contravariant push-forward),
//   Jc = lambda I x I + 2 mu S4  (the constant tensor the reference returns, not pushed forward).
// Written as an IFEM_HD function so that tests/cpp/hyper_materials_cpu.cpp can compile it with g++ and check it against
// the oracle without a GPU; solid.cu calls it from update_qph_kernel. Voigt pair order as everywhere in solid.cu:
// (0,0),(1,1)[,(2,2)],(0,1)[,(0,2),(1,2)].
#pragma once
#include <cmath>

#ifdef __CUDACC__
#define IFEM_HD __host__ __device__ __forceinline__
#else
#define IFEM_HD inline
#endif

namespace ifem
{
  template <int DIM>
  IFEM_HD int voigt_index(int i, int j)
  {
    if (i == j) return i;
    if (DIM == 2) return 2;
    const int a = i < j ? i : j, b = i < j ? j : i;
    return a == 0 ? (b == 1 ? 3 : 4) : 5;
  }

  template <int DIM>
  IFEM_HD void kirchhoff_point(const double *gu /*[DIM*DIM] Grad u*/, double young, double poisson, double *Finv, double *tau,
                               double *Jc /*[NS*NS]*/, double &detF)
  {
    using std::fma;
    constexpr int NS = DIM * (DIM + 1) / 2;
    const double lambda = young * poisson / ((1 + poisson) * (1 - 2 * poisson)), mu = young / (2 * (1 + poisson));
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
    // Green-Lagrange strain and second Piola-Kirchhoff stress
    double S[DIM * DIM], trE = 0.0;
#pragma unroll
    for (int i = 0; i < DIM; ++i)
#pragma unroll
      for (int j = 0; j < DIM; ++j)
        {
          double c = 0.0;
#pragma unroll
          for (int k = 0; k < DIM; ++k) c = fma(F[k * DIM + i], F[k * DIM + j], c);
          S[i * DIM + j] = 0.5 * (c - (i == j ? 1.0 : 0.0)); // E_ij for now
        }
#pragma unroll
    for (int i = 0; i < DIM; ++i) trE += S[i * DIM + i];
#pragma unroll
    for (int i = 0; i < DIM; ++i)
#pragma unroll
      for (int j = 0; j < DIM; ++j) S[i * DIM + j] = 2.0 * mu * S[i * DIM + j] + (i == j ? lambda * trE : 0.0);
    // tau = F S F^T
    double FS[DIM * DIM];
#pragma unroll
    for (int i = 0; i < DIM; ++i)
#pragma unroll
      for (int j = 0; j < DIM; ++j)
        {
          double s = 0.0;
#pragma unroll
          for (int k = 0; k < DIM; ++k) s = fma(F[i * DIM + k], S[k * DIM + j], s);
          FS[i * DIM + j] = s;
        }
#pragma unroll
    for (int i = 0; i < DIM; ++i)
#pragma unroll
      for (int j = 0; j < DIM; ++j)
        {
          double s = 0.0;
#pragma unroll
          for (int k = 0; k < DIM; ++k) s = fma(FS[i * DIM + k], F[j * DIM + k], s);
          tau[i * DIM + j] = s;
        }
    // Jc_ijkl = lambda d_ij d_kl + 2 mu (d_ik d_jl + d_il d_jk) / 2
#pragma unroll
    for (int i = 0; i < DIM; ++i)
#pragma unroll
      for (int j = i; j < DIM; ++j)
#pragma unroll
        for (int k = 0; k < DIM; ++k)
#pragma unroll
          for (int l = k; l < DIM; ++l)
            {
              const double S4 = 0.5 * ((i == k && j == l ? 1.0 : 0.0) + (i == l && j == k ? 1.0 : 0.0));
              Jc[voigt_index<DIM>(i, j) * NS + voigt_index<DIM>(k, l)] = lambda * (i == j ? 1.0 : 0.0) * (k == l ? 1.0 : 0.0) + 2.0 * mu * S4;
            }
  }
} // namespace ifem
