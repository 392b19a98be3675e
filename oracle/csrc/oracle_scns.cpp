// ORACLE (test infrastructure, NOT product code).
// CPU restatement of Fluid::MPI::SCnsIM<dim>::assemble (reference source/mpi_scnsim.cpp:15-568):
// slightly compressible Navier-Stokes with SUPG / PSPG / LSIC stabilisation, PML attenuation, isentropic
// continuity equation and the artificial-fluid (indicator == 1) terms. Same q / i / j loops over ALL
// dofs_per_cell with dense tensor expressions, term by term in the reference's order.
// FE tables come from oracle/fem.py; the element is FESystem(FE_Q(pu)^dim, FE_Q(pp)), the nodal stress
// fields live on FE_Q(pu) (scalar_fe, mpi_fluid_solver.cpp:27-35).
#include "oracle_common.h"

using namespace oracle;

namespace
{
  struct ScnsArgs
  {
    int nu, np, dpc, n_cells, nq;
    const double *vertices;
    const int *cells, *cell_dofs, *cell_unodes;
    const double *qw, *Nu, *dNu, *Np, *dNp, *dNgeo;
    int nqf;
    const double *qwf, *Nu_face, *dNgeo_face;
    const double *eval_pt, *present, *fsi_acc;
    const int *indicator;
    const double *stress;     // [dim*dim][n_unodes] nodal viscous stress (relevant_partition_stress) or null
    const double *fsi_stress; // [dim(dim+1)/2][n_unodes] or null
    int n_unodes;
    const double *sigma_pml;  // [n_cells][nq] or null
    const double *body_force; // [n_cells][nq][dim] or null
    int n_h;
    const int *h_type, *h_node; // the first dofs_per_cell/dofs_per_vertex shape functions (:251-257)
    double viscosity, rho_f, rho_s, dt;
    const double *gravity;
    int n_bfaces;
    const int *bfaces;
    int n_neumann;
    const int *neumann_ids;
    const double *neumann_vals;
    const unsigned char *con;
    const double *inhom;
    const int64_t *rowptr;
    const int *col;
    double *A, *rhs;
  };

  const double *g_eddy_viscosity = nullptr; // nodal eddy viscosity of an attached turbulence model (oracle_scns_set_eddy_viscosity)

  template <int dim>
  inline T1<dim> vecmat(const T1<dim> &a, const T2<dim> &B) // a * B  (contract a with first index of B)
  {
    T1<dim> r;
    for (int j = 0; j < dim; ++j)
      for (int i = 0; i < dim; ++i) r[j] += a[i] * B[i][j];
    return r;
  }

  template <int dim>
  void assemble(const ScnsArgs &a)
  {
    const int nu = a.nu, np = a.np, dpc = a.dpc, nq = a.nq, nv = 1 << dim;
    const double cp_to_cv = 1.4, atm = 1013250, kappa_s = 1e4; // :124-126
    T1<dim> gravity;
    for (int d = 0; d < dim; ++d) gravity[d] = a.gravity[d];
    const double dt = a.dt;

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
      std::vector<double> local_matrix(dpc * dpc), local_rhs(dpc);
      std::vector<double> div_phi_u(dpc), phi_p(dpc);
      std::vector<T1<dim>> phi_u(dpc), grad_phi_p(dpc);
      std::vector<T2<dim>> grad_phi_u(dpc);
      std::vector<T1<dim>> gu(nu), gp(np); // scalar shape gradients in real space

#pragma omp for schedule(dynamic, 4)
      for (int cell = 0; cell < a.n_cells; ++cell)
        {
          const int *dofs = a.cell_dofs + (size_t)cell * dpc;
          const int *un = a.cell_unodes + (size_t)cell * nu;
          const int *cv = a.cells + (size_t)cell * nv;
          std::fill(local_matrix.begin(), local_matrix.end(), 0.0);
          std::fill(local_rhs.begin(), local_rhs.end(), 0.0);
          const int ind = a.indicator ? a.indicator[cell] : 0;

          for (int q = 0; q < nq; ++q)
            {
              const T2<dim> J = jacobian<dim>(a.vertices, cv, a.dNgeo + (size_t)q * nv * dim);
              const T2<dim> Jinv = invert(J);
              const double JxW = det(J) * a.qw[q];
              for (int b = 0; b < nu; ++b)
                {
                  const double *dr = a.dNu + ((size_t)q * nu + b) * dim;
                  gu[b] = T1<dim>();
                  for (int i = 0; i < dim; ++i)
                    for (int j = 0; j < dim; ++j) gu[b][i] += dr[j] * Jinv[j][i];
                }
              for (int b = 0; b < np; ++b)
                {
                  const double *dr = a.dNp + ((size_t)q * np + b) * dim;
                  gp[b] = T1<dim>();
                  for (int i = 0; i < dim; ++i)
                    for (int j = 0; j < dim; ++j) gp[b][i] += dr[j] * Jinv[j][i];
                }
              for (int k = 0; k < dpc; ++k)
                {
                  phi_u[k] = T1<dim>();
                  grad_phi_u[k] = T2<dim>();
                  grad_phi_p[k] = T1<dim>();
                  div_phi_u[k] = 0;
                  phi_p[k] = 0;
                  if (k < nu * dim)
                    {
                      const int node = k / dim, c = k % dim;
                      phi_u[k][c] = a.Nu[q * nu + node];
                      for (int i = 0; i < dim; ++i) grad_phi_u[k][c][i] = gu[node][i];
                      div_phi_u[k] = gu[node][c];
                    }
                  else
                    {
                      phi_p[k] = a.Np[q * np + (k - nu * dim)];
                      grad_phi_p[k] = gp[k - nu * dim];
                    }
                }
              // function values / gradients at q (:153-171, :207-208)
              T1<dim> current_velocity_values, present_velocity_values, fsi_acc_values, current_pressure_gradients;
              T2<dim> current_velocity_gradients;
              double current_pressure_values = 0, present_pressure_values = 0;
              for (int k = 0; k < dpc; ++k)
                {
                  const double ue = a.eval_pt[dofs[k]], up = a.present[dofs[k]];
                  const double fa = a.fsi_acc ? a.fsi_acc[dofs[k]] : 0.0;
                  for (int i = 0; i < dim; ++i)
                    {
                      current_velocity_values[i] += ue * phi_u[k][i];
                      present_velocity_values[i] += up * phi_u[k][i];
                      fsi_acc_values[i] += fa * phi_u[k][i];
                      current_pressure_gradients[i] += ue * grad_phi_p[k][i];
                      for (int j = 0; j < dim; ++j) current_velocity_gradients[i][j] += ue * grad_phi_u[k][i][j];
                    }
                  current_pressure_values += ue * phi_p[k];
                  present_pressure_values += up * phi_p[k];
                }
              // nodal stress gradients and fsi stress values on the scalar FE_Q(pu) space (:173-186)
              T2<dim> stress_grad[dim]; // stress_grad[i][j][k] = d sigma_ij / d x_k
              if (a.stress)
                for (int i = 0; i < dim; ++i)
                  for (int j = 0; j < dim; ++j)
                    for (int b = 0; b < nu; ++b)
                      {
                        const double s = a.stress[(size_t)(i * dim + j) * a.n_unodes + un[b]];
                        for (int k = 0; k < dim; ++k) stress_grad[i][j][k] += s * gu[b][k];
                      }
              T2<dim> fsi_stress_tensor;
              if (ind != 0 && a.fsi_stress)
                {
                  int stress_index = 0;
                  for (int k = 0; k < dim; ++k)
                    for (int m = 0; m < k + 1; ++m)
                      {
                        double v = 0;
                        for (int b = 0; b < nu; ++b) v += a.fsi_stress[(size_t)stress_index * a.n_unodes + un[b]] * a.Nu[q * nu + b];
                        fsi_stress_tensor[k][m] = v;
                        fsi_stress_tensor[m][k] = v; // SymmetricTensor
                        stress_index++;
                      }
                }
              const double sigma_pml = a.sigma_pml ? a.sigma_pml[(size_t)cell * nq + q] : 0.0;
              T1<dim> artificial_bf;
              if (a.body_force)
                for (int d = 0; d < dim; ++d) artificial_bf[d] = a.body_force[((size_t)cell * nq + q) * dim + d];

              // :210-216
              const double rho = a.rho_f * (1 + present_pressure_values / atm) * (1 - ind) + ind * a.rho_s;
              // turbulence_model->get_eddy_viscosity() on scalar FE_Q(pu), clipped at zero (:198-203, :214-216)
              double eddy_viscosity = 0.0;
              if (g_eddy_viscosity)
                for (int b = 0; b < nu; ++b) eddy_viscosity += g_eddy_viscosity[un[b]] * a.Nu[q * nu + b];
              const double viscosity = (ind == 1 ? 1 : a.viscosity) + (eddy_viscosity > 0.0 ? eddy_viscosity : 0.0);

              // UGN stabilisation parameters (:247-274)
              double tau_SUPG, tau_PSPG, tau_LSIC;
              double h = 0.0;
              for (int k = 0; k < a.n_h; ++k)
                h += std::fabs(dot(present_velocity_values, a.h_type[k] == 0 ? gu[a.h_node[k]] : gp[a.h_node[k]]));
              const double v_norm = std::sqrt(dot(present_velocity_values, present_velocity_values));
              if (h)
                h = 2 * v_norm / h;
              else
                h = 0;
              const double nu_k = viscosity / rho;
              if (h)
                tau_SUPG = 1 / std::sqrt((std::pow(2 / dt, 2) + std::pow(2 * v_norm / h, 2) + std::pow(4 * nu_k / std::pow(h, 2), 2)));
              else
                tau_SUPG = dt / 2;
              tau_PSPG = tau_SUPG / rho;
              const double localRe = v_norm * h / (2 * nu_k);
              const double z = localRe <= 3 ? (localRe / 3) : 1;
              tau_LSIC = h / 2 * v_norm * z;

              // divergence of the nodal stress (:278-289)
              T1<dim> current_stress_divergence;
              for (int i = 0; i < dim; ++i)
                {
                  for (int j = 0; j < dim; ++j) current_stress_divergence[i] += stress_grad[i][j][j];
                  current_stress_divergence[i] *= viscosity / a.viscosity;
                }

              const double current_velocity_divergence = trace(current_velocity_gradients);
              T1<dim> g_plus_bf, dv;
              for (int d = 0; d < dim; ++d)
                {
                  g_plus_bf[d] = gravity[d] + artificial_bf[d];
                  dv[d] = current_velocity_values[d] - present_velocity_values[d];
                }
              // u . grad u  (Tensor<1> * Tensor<2> contracts the first index)
              const T1<dim> u_gradu = vecmat(current_velocity_values, current_velocity_gradients);
              const T1<dim> gradu_u = mul(current_velocity_gradients, current_velocity_values);

              for (int i = 0; i < dpc; ++i)
                {
                  // current_velocity_values * grad_phi_u[i]  and  phi_u[j] * grad_phi_u[i]
                  const T1<dim> u_gphi_i = vecmat(current_velocity_values, grad_phi_u[i]);
                  for (int j = 0; j < dpc; ++j)
                    {
                      const T1<dim> phij_gphi_i = vecmat(phi_u[j], grad_phi_u[i]);
                      const T1<dim> phij_gradu = vecmat(phi_u[j], current_velocity_gradients);
                      const T1<dim> u_gphi_j = vecmat(current_velocity_values, grad_phi_u[j]);
                      double m = 0;
                      // :307-319 Galerkin
                      m += ((viscosity * scalar_product(grad_phi_u[j], grad_phi_u[i]) +
                             rho * dot(mul(current_velocity_gradients, phi_u[j]), phi_u[i]) +
                             rho * dot(mul(grad_phi_u[j], current_velocity_values), phi_u[i]) - div_phi_u[i] * phi_p[j]) +
                            rho * dot(phi_u[i], phi_u[j]) / dt) *
                           JxW;
                      // :320-324 PML attenuation
                      m += (rho * sigma_pml * dot(phi_u[j], phi_u[i]) + sigma_pml * phi_p[j] * phi_p[i] / atm) * JxW;
                      // :325-398 SUPG / PSPG / LSIC
                      m += (tau_SUPG * rho * dot(u_gphi_i, phij_gradu) + tau_SUPG * rho * dot(u_gphi_i, u_gphi_j) +
                            tau_SUPG * rho * dot(phij_gphi_i, u_gradu) +
                            tau_SUPG * rho * dot(u_gphi_i, phi_u[j]) / dt + tau_SUPG * rho * dot(phij_gphi_i, dv) / dt +
                            tau_SUPG * dot(u_gphi_i, grad_phi_p[j]) + tau_SUPG * dot(phij_gphi_i, current_pressure_gradients) -
                            tau_SUPG * dot(phij_gphi_i, current_stress_divergence) -
                            tau_SUPG * dot(phij_gphi_i, g_plus_bf) * rho +
                            tau_SUPG * rho * dot(u_gphi_i, phi_u[j]) * sigma_pml +
                            tau_SUPG * rho * dot(phij_gphi_i, current_velocity_values) * sigma_pml +
                            tau_PSPG * rho * dot(grad_phi_p[i], phij_gradu) + tau_PSPG * rho * dot(grad_phi_p[i], u_gphi_j) +
                            tau_PSPG * rho * dot(grad_phi_p[i], phi_u[j]) / dt + tau_PSPG * dot(grad_phi_p[i], grad_phi_p[j]) +
                            tau_PSPG * rho * dot(grad_phi_p[i], phi_u[j]) * sigma_pml +
                            tau_LSIC * rho * div_phi_u[i] * phi_p[j] / dt * (1 - ind) / atm +
                            tau_LSIC * rho * 1 / kappa_s * div_phi_u[i] * phi_p[j] / dt * ind +
                            tau_LSIC * rho * cp_to_cv * div_phi_u[i] * div_phi_u[j] +
                            tau_LSIC * rho * cp_to_cv * div_phi_u[i] * current_pressure_values * (1 - ind) * div_phi_u[j] / atm +
                            tau_LSIC * rho * cp_to_cv * div_phi_u[i] * phi_p[j] * (1 - ind) * current_velocity_divergence / atm +
                            tau_LSIC * rho * div_phi_u[i] * dot(current_velocity_values, grad_phi_p[j]) / atm * (1 - ind) +
                            tau_LSIC * rho * div_phi_u[i] * dot(phi_u[j], current_pressure_gradients) / atm * (1 - ind)) *
                           JxW;
                      // :405-418 continuity
                      m += (cp_to_cv * (atm + current_pressure_values * (1 - ind)) * div_phi_u[j] * phi_p[i] +
                            phi_p[j] * current_velocity_divergence * phi_p[i] * (1 - ind) +
                            dot(current_velocity_values, grad_phi_p[j]) * phi_p[i] * (1 - ind) +
                            dot(phi_u[j], current_pressure_gradients) * phi_p[i] * (1 - ind) +
                            phi_p[i] * phi_p[j] / dt * (1 - ind)) /
                             atm * JxW +
                           1 / kappa_s * phi_p[i] * phi_p[j] * ind / dt * JxW;
                      if (ind == 1) // :419-425
                        m += -(tau_SUPG * dot(phij_gphi_i, fsi_acc_values) * rho) * JxW;
                      local_matrix[i * dpc + j] += m;
                    }

                  // rhs :429-512
                  double r = 0;
                  r += ((-viscosity * scalar_product(current_velocity_gradients, grad_phi_u[i]) -
                         rho * dot(gradu_u, phi_u[i]) + current_pressure_values * div_phi_u[i]) -
                        rho * dot(dv, phi_u[i]) / dt + dot(g_plus_bf, phi_u[i]) * rho) *
                       JxW;
                  r += -(rho * sigma_pml * dot(current_velocity_values, phi_u[i]) + sigma_pml * current_pressure_values * phi_p[i] / atm) * JxW;
                  r += -(cp_to_cv * (atm + current_pressure_values * (1 - ind)) * current_velocity_divergence * phi_p[i] +
                         dot(current_velocity_values, current_pressure_gradients) * phi_p[i] * (1 - ind) +
                         (current_pressure_values - present_pressure_values) * phi_p[i] / dt * (1 - ind)) /
                         atm * JxW -
                       1 / kappa_s * (current_pressure_values - present_pressure_values) * phi_p[i] * ind / dt * JxW;
                  // momentum residual used by SUPG / PSPG
                  T1<dim> res;
                  for (int d = 0; d < dim; ++d)
                    res[d] = rho * (dv[d] / dt + u_gradu[d]) + current_pressure_gradients[d] - current_stress_divergence[d] -
                             rho * g_plus_bf[d] + rho * sigma_pml * current_velocity_values[d];
                  r += -(tau_SUPG * dot(u_gphi_i, res) + tau_PSPG * dot(grad_phi_p[i], res)) * JxW;
                  r += -((tau_LSIC * rho * div_phi_u[i]) *
                           ((current_pressure_values - present_pressure_values) / dt * (1 - ind) +
                            cp_to_cv * atm * current_velocity_divergence +
                            cp_to_cv * current_pressure_values * current_velocity_divergence * (1 - ind) +
                            dot(current_velocity_values, current_pressure_gradients) * (1 - ind)) /
                           atm +
                         (tau_LSIC * rho * div_phi_u[i]) * (1 / kappa_s * (current_pressure_values - present_pressure_values) / dt) * ind) *
                       JxW;
                  if (ind == 1)
                    {
                      T1<dim> w;
                      for (int d = 0; d < dim; ++d) w[d] = phi_u[i][d] + tau_PSPG * grad_phi_p[i][d] + tau_SUPG * u_gphi_i[d];
                      r += (scalar_product(grad_phi_u[i], fsi_stress_tensor) + dot(fsi_acc_values, w) * rho) * JxW;
                    }
                  local_rhs[i] += r;
                }
            }

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
                    T1<dim> nds;
                    for (int i = 0; i < dim; ++i) nds[i] = dJ * Jinv[axis][i] * (side ? 1.0 : -1.0) * a.qwf[q];
                    for (int i = 0; i < nu * dim; ++i)
                      local_rhs[i] += -(a.Nu_face[fq * nu + i / dim] * nds[i % dim] * fp.second);
                  }
              }

          distribute_local_to_global(dpc, local_matrix.data(), local_rhs.data(), dofs, a.con, a.inhom, a.rowptr, a.col, a.A,
                                     a.rhs, true);
        }
    }
  }
} // namespace

extern "C" void oracle_scns_set_eddy_viscosity(const double *nodal) { g_eddy_viscosity = nodal; }

extern "C" int oracle_scns_assemble(
  int dim, int nu, int np, int n_cells, const double *vertices, const int *cells, const int *cell_dofs, const int *cell_unodes,
  int nq, const double *qw, const double *Nu, const double *dNu, const double *Np, const double *dNp, const double *dNgeo, int nqf,
  const double *qwf, const double *Nu_face, const double *dNgeo_face, const double *eval_pt, const double *present,
  const double *fsi_acc, const int *indicator, const double *stress, const double *fsi_stress, int n_unodes,
  const double *sigma_pml, const double *body_force, int n_h, const int *h_type, const int *h_node, double viscosity, double rho_f,
  double rho_s, double dt, const double *gravity, int n_bfaces, const int *bfaces, int n_neumann, const int *neumann_ids,
  const double *neumann_vals, const unsigned char *con, const double *inhom, const int64_t *rowptr, const int *col, double *A,
  double *rhs)
{
  ScnsArgs a;
  a.nu = nu; a.np = np; a.dpc = nu * dim + np; a.n_cells = n_cells; a.nq = nq;
  a.vertices = vertices; a.cells = cells; a.cell_dofs = cell_dofs; a.cell_unodes = cell_unodes;
  a.qw = qw; a.Nu = Nu; a.dNu = dNu; a.Np = Np; a.dNp = dNp; a.dNgeo = dNgeo;
  a.nqf = nqf; a.qwf = qwf; a.Nu_face = Nu_face; a.dNgeo_face = dNgeo_face;
  a.eval_pt = eval_pt; a.present = present; a.fsi_acc = fsi_acc; a.indicator = indicator;
  a.stress = stress; a.fsi_stress = fsi_stress; a.n_unodes = n_unodes; a.sigma_pml = sigma_pml; a.body_force = body_force;
  a.n_h = n_h; a.h_type = h_type; a.h_node = h_node;
  a.viscosity = viscosity; a.rho_f = rho_f; a.rho_s = rho_s; a.dt = dt; a.gravity = gravity;
  a.n_bfaces = n_bfaces; a.bfaces = bfaces; a.n_neumann = n_neumann; a.neumann_ids = neumann_ids; a.neumann_vals = neumann_vals;
  a.con = con; a.inhom = inhom; a.rowptr = rowptr; a.col = col; a.A = A; a.rhs = rhs;
  if (dim == 2) assemble<2>(a);
  else if (dim == 3) assemble<3>(a);
  else return 1;
  return 0;
}
