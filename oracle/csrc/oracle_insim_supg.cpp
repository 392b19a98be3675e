// ORACLE (test infrastructure, NOT product code).
// CPU restatement of Fluid::MPI::SUPGInsIM<dim>::assemble (reference source/mpi_insim_supg.cpp:15-328): incompressible
// Navier-Stokes on equal-order elements with SUPG / PSPG / LSIC stabilisation (UGN parameters from the previous-step
// velocity, :122-152), backward Euler, Newton-linearised. Same q / i / j loops over all dofs_per_cell with the dense
// tensor expressions of the reference, term by term in its order. The solver around it is SUPGFluidSolver (oracle/scns.py).
#include "oracle_common.h"

using namespace oracle;

namespace
{
  template <int dim>
  inline T1<dim> vecmat(const T1<dim> &a, const T2<dim> &B) // Tensor<1> * Tensor<2>: first index of B contracted
  {
    T1<dim> r;
    for (int j = 0; j < dim; ++j)
      for (int i = 0; i < dim; ++i) r[j] += a[i] * B[i][j];
    return r;
  }

  template <int dim>
  void assemble_supg(int nu, int np, int n_cells, const double *vertices, const int *cells, const int *cell_dofs, int nq,
                     const double *qw, const double *Nu, const double *dNu, const double *Np, const double *dNp, const double *dNgeo,
                     int nqf, const double *qwf, const double *Nu_face, const double *dNgeo_face, const double *eval_pt,
                     const double *present, const double *body_force, int n_h, const int *h_type, const int *h_node, double viscosity,
                     double rho, double dt, const double *grav, int n_bfaces, const int *bfaces, int n_neumann, const int *neumann_ids,
                     const double *neumann_vals, const unsigned char *con, const double *inhom, const int64_t *rowptr, const int *col,
                     double *A, double *rhs)
  {
    const int dpc = nu * dim + np, nv = 1 << dim;
    T1<dim> gravity;
    for (int d = 0; d < dim; ++d) gravity[d] = grav[d];
    std::vector<std::vector<std::pair<int, double>>> cell_nfaces(n_neumann ? n_cells : 0);
    for (int f = 0; f < (n_neumann ? n_bfaces : 0); ++f)
      for (int k = 0; k < n_neumann; ++k)
        if (bfaces[3 * f + 2] == neumann_ids[k]) cell_nfaces[bfaces[3 * f]].push_back({bfaces[3 * f + 1], neumann_vals[k]});

#pragma omp parallel
    {
      std::vector<double> local_matrix(dpc * dpc), local_rhs(dpc), div_phi_u(dpc), phi_p(dpc);
      std::vector<T1<dim>> phi_u(dpc), grad_phi_p(dpc), gu(nu), gp(np);
      std::vector<T2<dim>> grad_phi_u(dpc);
#pragma omp for schedule(dynamic, 4)
      for (int cell = 0; cell < n_cells; ++cell)
        {
          const int *dofs = cell_dofs + (size_t)cell * dpc, *cv = cells + (size_t)cell * nv;
          std::fill(local_matrix.begin(), local_matrix.end(), 0.0);
          std::fill(local_rhs.begin(), local_rhs.end(), 0.0);
          for (int q = 0; q < nq; ++q)
            {
              const T2<dim> J = jacobian<dim>(vertices, cv, dNgeo + (size_t)q * nv * dim);
              const T2<dim> Jinv = invert(J);
              const double JxW = det(J) * qw[q];
              for (int b = 0; b < nu; ++b)
                {
                  gu[b] = T1<dim>();
                  for (int i = 0; i < dim; ++i)
                    for (int j = 0; j < dim; ++j) gu[b][i] += dNu[((size_t)q * nu + b) * dim + j] * Jinv[j][i];
                }
              for (int b = 0; b < np; ++b)
                {
                  gp[b] = T1<dim>();
                  for (int i = 0; i < dim; ++i)
                    for (int j = 0; j < dim; ++j) gp[b][i] += dNp[((size_t)q * np + b) * dim + j] * Jinv[j][i];
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
                      phi_u[k][c] = Nu[q * nu + node];
                      for (int i = 0; i < dim; ++i) grad_phi_u[k][c][i] = gu[node][i];
                      div_phi_u[k] = gu[node][c];
                    }
                  else
                    {
                      phi_p[k] = Np[q * np + (k - nu * dim)];
                      grad_phi_p[k] = gp[k - nu * dim];
                    }
                }
              // :84-103
              T1<dim> current_velocity_values, present_velocity_values, current_pressure_gradients, artificial_bf;
              T2<dim> current_velocity_gradients;
              double current_pressure_values = 0;
              for (int k = 0; k < dpc; ++k)
                {
                  const double ue = eval_pt[dofs[k]], up = present[dofs[k]];
                  for (int i = 0; i < dim; ++i)
                    {
                      current_velocity_values[i] += ue * phi_u[k][i];
                      present_velocity_values[i] += up * phi_u[k][i];
                      current_pressure_gradients[i] += ue * grad_phi_p[k][i];
                      for (int j = 0; j < dim; ++j) current_velocity_gradients[i][j] += ue * grad_phi_u[k][i][j];
                    }
                  current_pressure_values += ue * phi_p[k];
                }
              if (body_force)
                for (int d = 0; d < dim; ++d) artificial_bf[d] = body_force[((size_t)cell * nq + q) * dim + d];
              // UGN parameters (:122-152)
              double tau_SUPG, tau_PSPG, tau_LSIC, h = 0;
              for (int k = 0; k < n_h; ++k) h += std::fabs(dot(present_velocity_values, h_type[k] == 0 ? gu[h_node[k]] : gp[h_node[k]]));
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

              const double current_velocity_divergence = trace(current_velocity_gradients);
              T1<dim> g_plus_bf, dv;
              for (int d = 0; d < dim; ++d)
                {
                  g_plus_bf[d] = gravity[d] + artificial_bf[d];
                  dv[d] = current_velocity_values[d] - present_velocity_values[d];
                }
              const T1<dim> u_gradu = vecmat(current_velocity_values, current_velocity_gradients);
              const T1<dim> gradu_u = mul(current_velocity_gradients, current_velocity_values);
              for (int i = 0; i < dpc; ++i)
                {
                  const T1<dim> u_gphi_i = vecmat(current_velocity_values, grad_phi_u[i]);
                  for (int j = 0; j < dpc; ++j)
                    {
                      const T1<dim> phij_gphi_i = vecmat(phi_u[j], grad_phi_u[i]);
                      const T1<dim> phij_gradu = vecmat(phi_u[j], current_velocity_gradients);
                      const T1<dim> u_gphi_j = vecmat(current_velocity_values, grad_phi_u[j]);
                      double m = 0;
                      // :170-181 Galerkin
                      m += ((viscosity * scalar_product(grad_phi_u[j], grad_phi_u[i]) +
                             rho * dot(mul(current_velocity_gradients, phi_u[j]), phi_u[i]) +
                             rho * dot(mul(grad_phi_u[j], current_velocity_values), phi_u[i]) - div_phi_u[i] * phi_p[j]) +
                            rho * dot(phi_u[i], phi_u[j]) / dt) *
                           JxW;
                      // :183-226 SUPG / PSPG / LSIC
                      m += (tau_SUPG * rho * dot(u_gphi_i, phij_gradu) + tau_SUPG * rho * dot(u_gphi_i, u_gphi_j) +
                            tau_SUPG * rho * dot(phij_gphi_i, u_gradu) + tau_SUPG * rho * dot(u_gphi_i, phi_u[j]) / dt +
                            tau_SUPG * rho * dot(phij_gphi_i, dv) / dt + tau_SUPG * dot(u_gphi_i, grad_phi_p[j]) +
                            tau_SUPG * dot(phij_gphi_i, current_pressure_gradients) - tau_SUPG * dot(phij_gphi_i, g_plus_bf) * rho +
                            tau_PSPG * rho * dot(grad_phi_p[i], phij_gradu) + tau_PSPG * rho * dot(grad_phi_p[i], u_gphi_j) +
                            tau_PSPG * rho * dot(grad_phi_p[i], phi_u[j]) / dt + tau_PSPG * dot(grad_phi_p[i], grad_phi_p[j]) +
                            tau_LSIC * rho * div_phi_u[i] * div_phi_u[j]) *
                           JxW;
                      // :234-235 continuity
                      m += div_phi_u[j] * phi_p[i] * JxW;
                      local_matrix[i * dpc + j] += m;
                    }
                  // :240-285 rhs
                  double r = 0;
                  r += ((-viscosity * scalar_product(current_velocity_gradients, grad_phi_u[i]) - rho * dot(gradu_u, phi_u[i]) +
                         current_pressure_values * div_phi_u[i]) -
                        rho * dot(dv, phi_u[i]) / dt + dot(g_plus_bf, phi_u[i]) * rho) *
                       JxW;
                  r += -(current_velocity_divergence * phi_p[i]) * JxW;
                  T1<dim> res;
                  for (int d = 0; d < dim; ++d)
                    res[d] = rho * (dv[d] / dt + u_gradu[d]) + current_pressure_gradients[d] - rho * g_plus_bf[d];
                  r += -(tau_SUPG * dot(u_gphi_i, res) + tau_PSPG * dot(grad_phi_p[i], res)) * JxW;
                  r += -(tau_LSIC * rho * div_phi_u[i]) * current_velocity_divergence * JxW;
                  local_rhs[i] += r;
                }
            }
          if (n_neumann) // :292-321
            for (auto &fp : cell_nfaces[cell])
              {
                const int face = fp.first, axis = face / 2, side = face % 2;
                for (int q = 0; q < nqf; ++q)
                  {
                    const size_t fq = (size_t)face * nqf + q;
                    const T2<dim> J = jacobian<dim>(vertices, cv, dNgeo_face + fq * nv * dim);
                    const T2<dim> Jinv = invert(J);
                    const double dJ = det(J);
                    for (int i = 0; i < nu * dim; ++i)
                      local_rhs[i] += -(Nu_face[fq * nu + i / dim] * dJ * Jinv[axis][i % dim] * (side ? 1.0 : -1.0) * qwf[q] * fp.second);
                  }
              }
          distribute_local_to_global(dpc, local_matrix.data(), local_rhs.data(), dofs, con, inhom, rowptr, col, A, rhs, true);
        }
    }
  }
} // namespace

extern "C" int oracle_insim_supg_assemble(int dim, int nu, int np, int n_cells, const double *vertices, const int *cells,
                                          const int *cell_dofs, int nq, const double *qw, const double *Nu, const double *dNu,
                                          const double *Np, const double *dNp, const double *dNgeo, int nqf, const double *qwf,
                                          const double *Nu_face, const double *dNgeo_face, const double *eval_pt, const double *present,
                                          const double *body_force, int n_h, const int *h_type, const int *h_node, double viscosity,
                                          double rho, double dt, const double *gravity, int n_bfaces, const int *bfaces, int n_neumann,
                                          const int *neumann_ids, const double *neumann_vals, const unsigned char *con,
                                          const double *inhom, const int64_t *rowptr, const int *col, double *A, double *rhs)
{
  if (dim == 2)
    assemble_supg<2>(nu, np, n_cells, vertices, cells, cell_dofs, nq, qw, Nu, dNu, Np, dNp, dNgeo, nqf, qwf, Nu_face, dNgeo_face, eval_pt,
                     present, body_force, n_h, h_type, h_node, viscosity, rho, dt, gravity, n_bfaces, bfaces, n_neumann, neumann_ids,
                     neumann_vals, con, inhom, rowptr, col, A, rhs);
  else if (dim == 3)
    assemble_supg<3>(nu, np, n_cells, vertices, cells, cell_dofs, nq, qw, Nu, dNu, Np, dNp, dNgeo, nqf, qwf, Nu_face, dNgeo_face, eval_pt,
                     present, body_force, n_h, h_type, h_node, viscosity, rho, dt, gravity, n_bfaces, bfaces, n_neumann, neumann_ids,
                     neumann_vals, con, inhom, rowptr, col, A, rhs);
  else
    return 1;
  return 0;
}
