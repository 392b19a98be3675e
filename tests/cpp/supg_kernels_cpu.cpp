// TEST INFRASTRUCTURE: the kernel bodies of openifem_b200/csrc/insim_supg.cuh compiled with g++ and run over the launch
// grid of supg_ins_assemble_kernel (insim_supg.cu) phase by phase - a __syncthreads() becomes the end of a loop over the
// threads of the block - so that tests/test_supg_kernels_cpu.py can check the device arithmetic, indexing and constrained
// scatter against the oracle without a GPU. Nothing in the product links this file.
#include "../../openifem_b200/csrc/insim_supg.cuh"

#include <numeric>
#include <vector>

using namespace ifem;

template <int DIM>
static void run(SupgArgs A)
{
  constexpr int NU = 1 << DIM, NQ = NU, PAIRS = NU * NU, CPB = 64 / PAIRS, DPC = NU * (DIM + 1);
  const int blocks = (A.n_list + CPB - 1) / CPB;
  for (int blk = 0; blk < blocks; ++blk)
    {
      SupgQPoint<DIM> sq[CPB][NQ];
      double lrhs[CPB][DPC], ldiag[CPB][DPC];
      auto each = [&](auto &&f) {
        for (int t = 0; t < 64; ++t)
          {
            const int cl = t / PAIRS, pr = t % PAIRS, li = blk * CPB + cl;
            if (li < A.n_list) f(cl, pr, A.cell_list[li]);
          }
      };
      each([&](int cl, int pr, int cell) {
        if (pr < NQ) supg_fill_qpoint<DIM>(A, cell, pr, sq[cl][pr]);
        if (pr < DPC) lrhs[cl][pr] = ldiag[cl][pr] = 0.0;
      });
      each([&](int cl, int pr, int cell) { supg_pair_body<DIM>(A, cell, pr, sq[cl], lrhs[cl], ldiag[cl]); });
      each([&](int cl, int pr, int cell) {
        if (pr < DPC) supg_rhs_body<DIM>(A, cell, pr, lrhs[cl], ldiag[cl]);
      });
    }
}

extern "C" int cpu_supg_assemble(int dim, int n_cells, const int *cell_un, const int *cell_pn, const double *cell_x, const double *tables,
                                 const unsigned char *slots, const unsigned char *con, const double *eval_pt, const double *present,
                                 const double *body_force, const double *inhom, int64_t n_u, int n_unodes, int n_pnodes, double mu,
                                 double rho, double dt, const double *grav, const int64_t *uu_rp, const int64_t *up_rp,
                                 const int64_t *pu_rp, const int64_t *pp_rp, double *uu, double *up, double *pu, double *pp, double *rhs)
{
  std::vector<int> list(n_cells);
  std::iota(list.begin(), list.end(), 0);
  SupgArgs A{};
  A.n_list = n_cells;
  A.cell_list = list.data();
  A.cell_un = cell_un;
  A.cell_pn = cell_pn;
  A.cell_x = cell_x;
  A.tables = tables;
  A.slots = slots;
  A.con = con;
  A.eval_pt = eval_pt;
  A.present = present;
  A.body_force = body_force;
  A.inhom = inhom;
  A.n_u = n_u;
  A.n_owned_u = n_unodes;
  A.n_owned_p = n_pnodes;
  A.n_h = 1 << dim; // as SUPGInsIM::assemble sets it
  for (int k = 0; k < A.n_h; ++k) A.h_node[k] = k / (dim + 1);
  A.mu = mu;
  A.rho = rho;
  A.dt = dt;
  for (int d = 0; d < 3; ++d) A.grav[d] = d < dim ? grav[d] : 0.0;
  A.uu_rp = uu_rp; A.up_rp = up_rp; A.pu_rp = pu_rp; A.pp_rp = pp_rp;
  A.uu = uu; A.up = up; A.pu = pu; A.pp = pp; A.rhs = rhs;
  if (dim == 2) run<2>(A);
  else if (dim == 3) run<3>(A);
  else return 1;
  return 0;
}
