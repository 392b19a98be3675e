// TEST INFRASTRUCTURE: the kernel bodies of openifem_b200/csrc/solid_linear.cuh compiled with g++ and run over their
// launch grid sequentially, so that tests/test_linear_kernels_cpu.py can check the device arithmetic and indexing against
// the oracle without a GPU. Nothing in the product links this file.
#include "../../openifem_b200/csrc/solid_linear.cuh"

#include <numeric>
#include <vector>

using namespace ifem;

extern "C" int cpu_linear_assemble(int dim, int n_cells, const int *cell_nodes, const unsigned char *slots, const unsigned char *con,
                                   const double *N, const double *G, const double *JxW, int nq, double rho, double lambda, double mu,
                                   double eta, const double *grav, double c_mass, double c_damp, double c_stiff, const int64_t *rowptr,
                                   double *sys, double *mass, double *stiff, double *damp, double *rhs)
{
  std::vector<int> list(n_cells);
  std::iota(list.begin(), list.end(), 0);
  LinearArgs A;
  A.n_list = n_cells;
  A.cell_list = list.data();
  A.cell_nodes = cell_nodes;
  A.slots = slots;
  A.con = con;
  A.N = N;
  A.G = G;
  A.JxW = JxW;
  A.nq = nq;
  A.rho = rho;
  A.lambda = lambda;
  A.mu = mu;
  A.eta = eta;
  for (int d = 0; d < 3; ++d) A.grav[d] = d < dim ? grav[d] : 0.0;
  A.c_mass = c_mass;
  A.c_damp = c_damp;
  A.c_stiff = c_stiff;
  A.rowptr = rowptr;
  A.sys = sys;
  A.mass = mass;
  A.stiff = stiff;
  A.damp = damp;
  A.rhs = rhs;
  // the launch configuration of LinearElasticity::assemble_system: 64 threads, 2-D: 4 cells per block, 3-D: 1
  const int blocks = dim == 2 ? (n_cells + 3) / 4 : n_cells;
  for (int b = 0; b < blocks; ++b)
    for (int t = 0; t < 64; ++t)
      {
        if (dim == 2)
          linear_assemble_body<2, 4>(A, b, t);
        else if (dim == 3)
          linear_assemble_body<3, 8>(A, b, t);
        else
          return 1;
      }
  return 0;
}

extern "C" int cpu_linear_stress(int dim, int n_cells, int nq, const int *cell_nodes, const double *qpt_to_dof, const double *G,
                                 const double *u, double lambda, double mu, int n_nodes, double *stress, double *strain, double *count)
{
  std::vector<int> list(n_cells);
  std::iota(list.begin(), list.end(), 0);
  const int block = 128, blocks = (n_cells + block - 1) / block;
  for (int b = 0; b < blocks; ++b)
    for (int t = 0; t < block; ++t)
      {
        if (dim == 2)
          linear_stress_body<2, 4>(b * block + t, n_cells, list.data(), nq, cell_nodes, qpt_to_dof, G, u, lambda, mu, n_nodes, stress, strain, count);
        else if (dim == 3)
          linear_stress_body<3, 8>(b * block + t, n_cells, list.data(), nq, cell_nodes, qpt_to_dof, G, u, lambda, mu, n_nodes, stress, strain, count);
        else
          return 1;
      }
  return 0;
}
