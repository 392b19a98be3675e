// TEST INFRASTRUCTURE: openifem_b200/csrc/hyper_materials.cuh compiled with g++ so that tests/test_hyper_materials_cpu.py can
// check the device function of the Kirchhoff material point against the oracle without a GPU. Nothing in the product links this.
#include "../../openifem_b200/csrc/hyper_materials.cuh"

extern "C" int cpu_kirchhoff_points(int dim, int n, const double *grad_u, double young, double poisson, double *Finv, double *tau,
                                    double *Jc, double *detF)
{
  const int ns = dim * (dim + 1) / 2;
  for (int t = 0; t < n; ++t)
    {
      if (dim == 2)
        ifem::kirchhoff_point<2>(grad_u + (long)t * 4, young, poisson, Finv + (long)t * 4, tau + (long)t * 4, Jc + (long)t * ns * ns, detF[t]);
      else if (dim == 3)
        ifem::kirchhoff_point<3>(grad_u + (long)t * 9, young, poisson, Finv + (long)t * 9, tau + (long)t * 9, Jc + (long)t * ns * ns, detF[t]);
      else
        return 1;
    }
  return 0;
}
