// ORACLE (test infrastructure, NOT product code).
// CPU restatement of Fluid::MPI::InsIM<dim>::assemble
// (reference source/mpi_insim.cpp:152-362; serial twin source/insim.cpp:142-333):
// the same q / i / j loops over ALL dofs_per_cell with dense Tensor products,
// so both the numbers and the cost profile follow the reference.
// FE tables (shape values / reference gradients at quadrature points) are
// passed in from oracle/fem.py.
#include "oracle_common.h"

using namespace oracle;

namespace
{
  struct InsArgs
  {
    int pu_nodes, pp_nodes, dpc;
    int n_cells;
    const double *vertices;
    const int *cells;
    const int *cell_dofs;
    int nq;
    const double *qw, *Nu, *dNu, *Np, *dNgeo;
    int nqf;
    const double *qwf, *Nu_face, *dNgeo_face; // [2*dim][nqf][...]
    const double *eval_pt, *present, *fsi_acc;
    const int *indicator;
    const double *cell_fsi_stress; // [n_cells][dim*dim] or null (CellProperty::fsi_stress)
    double viscosity, gamma, rho, dt;
    const double *gravity;
    int n_bfaces;
    const int *bfaces; // (cell, face_no, boundary id)
    int n_neumann;
    const int *neumann_ids;
    const double *neumann_vals;
    const unsigned char *con;
    const double *inhom; // null -> zero constraints
    const int64_t *rowptr;
    const int *col;
    double *A, *M, *rhs;
  };

  template <int dim>
  void assemble(const InsArgs &a)
  {
    const int nu = a.pu_nodes, np = a.pp_nodes, dpc = a.dpc, nq = a.nq;
    const int nv = 1 << dim;
    T1<dim> gravity;
    for (int d = 0; d < dim; ++d) gravity[d] = a.gravity[d];

    // boundary faces per cell (for the pressure Neumann term, :313-341)
    std::vector<std::vector<std::pair<int, double>>> cell_nfaces;
    if (a.n_neumann)
      {
        cell_nfaces.resize(a.n_cells);
        for (int f = 0; f < a.n_bfaces; ++f)
          for (int k = 0; k < a.n_neumann; ++k)
            if (a.bfaces[3 * f + 2] == a.neumann_ids[k])
              cell_nfaces[a.bfaces[3 * f]].push_back({a.bfaces[3 * f + 1], a.neumann_vals[k]});
      }

#pragma omp parallel
    {
      std::vector<double> local_matrix(dpc * dpc), local_mass(dpc * dpc), local_rhs(dpc);
      std::vector<double> div_phi_u(dpc), phi_p(dpc);
      std::vector<T1<dim>> phi_u(dpc);
      std::vector<T2<dim>> grad_phi_u(dpc);
      std::vector<T1<dim>> cur_v(nq), pre_v(nq), acc_v(nq);
      std::vector<T2<dim>> cur_g(nq);
      std::vector<double> cur_p(nq);

#pragma omp for schedule(dynamic, 4)
      for (int cell = 0; cell < a.n_cells; ++cell)
        {
          const int *dofs = a.cell_dofs + (size_t)cell * dpc;
          const int *cv = a.cells + (size_t)cell * nv;
          std::fill(local_matrix.begin(), local_matrix.end(), 0.0);
          std::fill(local_mass.begin(), local_mass.end(), 0.0);
          std::fill(local_rhs.begin(), local_rhs.end(), 0.0);
          const int ind = a.indicator ? a.indicator[cell] : 0;
          T2<dim> fsi_stress;
          if (a.cell_fsi_stress)
            for (int i = 0; i < dim; ++i)
              for (int j = 0; j < dim; ++j) fsi_stress[i][j] = a.cell_fsi_stress[(size_t)cell * dim * dim + i * dim + j];

          for (int q = 0; q < nq; ++q)
            {
              const T2<dim> J = jacobian<dim>(a.vertices, cv, a.dNgeo + (size_t)q * nv * dim);
              const T2<dim> Jinv = invert(J);
              const double JxW = det(J) * a.qw[q];
              // shape functions of the FESystem at q (primitive: one component each)
              for (int k = 0; k < dpc; ++k)
                {
                  phi_u[k] = T1<dim>();
                  grad_phi_u[k] = T2<dim>();
                  div_phi_u[k] = 0;
                  phi_p[k] = 0;
                  if (k < nu * dim)
                    {
                      const int node = k / dim, c = k % dim;
                      phi_u[k][c] = a.Nu[q * nu + node];
                      const double *dr = a.dNu + ((size_t)q * nu + node) * dim;
                      for (int i = 0; i < dim; ++i)
                        {
                          double g = 0;
                          for (int j = 0; j < dim; ++j) g += dr[j] * Jinv[j][i];
                          grad_phi_u[k][c][i] = g;
                        }
                      div_phi_u[k] = grad_phi_u[k][c][c];
                    }
                  else
                    phi_p[k] = a.Np[q * np + (k - nu * dim)];
                }
              // get_function_values / gradients (:219-232)
              cur_v[q] = T1<dim>(); pre_v[q] = T1<dim>(); acc_v[q] = T1<dim>();
              cur_g[q] = T2<dim>(); cur_p[q] = 0;
              for (int k = 0; k < dpc; ++k)
                {
                  const double ue = a.eval_pt[dofs[k]], up = a.present[dofs[k]];
                  const double fa = a.fsi_acc ? a.fsi_acc[dofs[k]] : 0.0;
                  for (int i = 0; i < dim; ++i)
                    {
                      cur_v[q][i] += ue * phi_u[k][i];
                      pre_v[q][i] += up * phi_u[k][i];
                      acc_v[q][i] += fa * phi_u[k][i];
                      for (int j = 0; j < dim; ++j) cur_g[q][i][j] += ue * grad_phi_u[k][i][j];
                    }
                  cur_p[q] += ue * phi_p[k];
                }

              const double rho = a.rho, viscosity = a.viscosity, gamma = a.gamma, dt = a.dt;
              for (int i = 0; i < dpc; ++i)
                {
                  for (int j = 0; j < dpc; ++j)
                    {
                      // mpi_insim.cpp:263-273
                      local_matrix[i * dpc + j] +=
                        (viscosity * scalar_product(grad_phi_u[j], grad_phi_u[i]) +
                         dot(mul(cur_g[q], phi_u[j]), phi_u[i]) * rho +
                         dot(mul(grad_phi_u[j], cur_v[q]), phi_u[i]) * rho -
                         div_phi_u[i] * phi_p[j] - phi_p[i] * div_phi_u[j] +
                         gamma * div_phi_u[j] * div_phi_u[i] * rho +
                         dot(phi_u[i], phi_u[j]) / dt * rho) *
                        JxW;
                      // :274-276
                      local_mass[i * dpc + j] += (dot(phi_u[i], phi_u[j]) + phi_p[i] * phi_p[j]) * JxW;
                    }
                  // :281-297
                  const double div_cur = trace(cur_g[q]);
                  T1<dim> dv;
                  for (int d = 0; d < dim; ++d) dv[d] = cur_v[q][d] - pre_v[q][d];
                  local_rhs[i] +=
                    ((-viscosity * scalar_product(cur_g[q], grad_phi_u[i]) -
                      dot(mul(cur_g[q], cur_v[q]), phi_u[i]) * rho + cur_p[q] * div_phi_u[i] +
                      div_cur * phi_p[i] - gamma * div_cur * div_phi_u[i] * rho) -
                     dot(dv, phi_u[i]) / dt * rho + dot(gravity, phi_u[i]) * rho) *
                    JxW;
                  if (ind == 1) // :298-304
                    local_rhs[i] += (scalar_product(grad_phi_u[i], fsi_stress) + dot(acc_v[q], phi_u[i]) * rho) * JxW;
                }
            }

          // pressure Neumann faces (:313-341): rhs_i -= phi_i . n * p * JxW_face
          if (a.n_neumann)
            for (auto &fp : cell_nfaces[cell])
              {
                const int face = fp.first, axis = face / 2, side = face % 2;
                for (int q = 0; q < a.nqf; ++q)
                  {
                    const size_t fq = (size_t)face * a.nqf + q;
                    const T2<dim> J = jacobian<dim>(a.vertices, cv, a.dNgeo_face + fq * nv * dim);
                    const T2<dim> Jinv = invert(J);
                    const double dJ = det(J);
                    // n dS = det(J) J^{-T} n_ref  (n_ref = -/+ e_axis)
                    T1<dim> nds;
                    for (int i = 0; i < dim; ++i) nds[i] = dJ * Jinv[axis][i] * (side ? 1.0 : -1.0) * a.qwf[q];
                    for (int i = 0; i < nu * dim; ++i)
                      {
                        const int node = i / dim, c = i % dim;
                        local_rhs[i] += -(a.Nu_face[fq * nu + node] * nds[c] * fp.second);
                      }
                  }
              }

          distribute_local_to_global(dpc, local_matrix.data(), local_rhs.data(), dofs, a.con, a.inhom, a.rowptr, a.col,
                                     a.A, a.rhs, true);
          if (a.M)
            distribute_local_to_global(dpc, local_mass.data(), nullptr, dofs, a.con, a.inhom, a.rowptr, a.col, a.M,
                                       nullptr, false);
        }
    }
  }
} // namespace

extern "C" int oracle_ins_assemble(
  int dim, int pu_nodes, int pp_nodes, int n_cells, const double *vertices, const int *cells, const int *cell_dofs, int nq,
  const double *qw, const double *Nu, const double *dNu, const double *Np, const double *dNgeo, int nqf,
  const double *qwf, const double *Nu_face, const double *dNgeo_face, const double *eval_pt, const double *present,
  const double *fsi_acc, const int *indicator, const double *cell_fsi_stress, double viscosity, double gamma, double rho,
  double dt, const double *gravity, int n_bfaces, const int *bfaces, int n_neumann, const int *neumann_ids,
  const double *neumann_vals, const unsigned char *con, const double *inhom, const int64_t *rowptr, const int *col,
  double *A, double *M, double *rhs)
{
  InsArgs a;
  a.pu_nodes = pu_nodes; a.pp_nodes = pp_nodes; a.dpc = pu_nodes * dim + pp_nodes;
  a.n_cells = n_cells; a.vertices = vertices; a.cells = cells; a.cell_dofs = cell_dofs;
  a.nq = nq; a.qw = qw; a.Nu = Nu; a.dNu = dNu; a.Np = Np; a.dNgeo = dNgeo;
  a.nqf = nqf; a.qwf = qwf; a.Nu_face = Nu_face; a.dNgeo_face = dNgeo_face;
  a.eval_pt = eval_pt; a.present = present; a.fsi_acc = fsi_acc; a.indicator = indicator;
  a.cell_fsi_stress = cell_fsi_stress;
  a.viscosity = viscosity; a.gamma = gamma; a.rho = rho; a.dt = dt; a.gravity = gravity;
  a.n_bfaces = n_bfaces; a.bfaces = bfaces; a.n_neumann = n_neumann; a.neumann_ids = neumann_ids;
  a.neumann_vals = neumann_vals; a.con = con; a.inhom = inhom; a.rowptr = rowptr; a.col = col;
  a.A = A; a.M = M; a.rhs = rhs;
  if (dim == 2) assemble<2>(a);
  else if (dim == 3) assemble<3>(a);
  else return 1;
  return 0;
}

// y = A x, CSR, one OpenMP thread team standing in for the MPI ranks of
// PETSc MatMult (reference call sites: mpi_insim.cpp:388 via SolverFGMRES, :117).
// Constraint lines with masters for the next assemblies of every oracle cell loop (n_dofs = 0 clears them); see
// oracle_common.h ConstraintLines.
extern "C" void oracle_set_constraint_lines(int64_t n_dofs, const int64_t *ptr, const int *master, const double *weight)
{
  oracle::ConstraintLines &L = oracle::constraint_lines();
  L.ptr.clear();
  L.master.clear();
  L.weight.clear();
  if (n_dofs <= 0) return;
  L.ptr.assign(ptr, ptr + n_dofs + 1);
  L.master.assign(master, master + ptr[n_dofs]);
  L.weight.assign(weight, weight + ptr[n_dofs]);
}

extern "C" void oracle_spmv_csr(int64_t n_rows, const int64_t *rowptr, const int *col, const double *val, const double *x,
                                double *y)
{
#pragma omp parallel for schedule(static) if (n_rows > 100000)
  for (int64_t r = 0; r < n_rows; ++r)
    {
      double s = 0;
      for (int64_t k = rowptr[r]; k < rowptr[r + 1]; ++k) s += val[k] * x[col[k]];
      y[r] = s;
    }
}
