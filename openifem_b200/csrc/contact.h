// The penetration scan of FSI::apply_contact_model (reference source/mpi_fsi.cpp:897-956) as a host function with no
// CUDA dependency: the criterion is a host callback and the solid is small and replicated, so the scan walks the boundary
// faces of the moved solid on the host (the solid steps around it run on the device, fsi.cu). Kept free of device types
// so that tests/cpp/contact_scan_cpu.cpp can run the very same code against the oracle without a GPU.
#pragma once
#include <cmath>
#include <cstdint>
#include <functional>
#include <vector>

#include "fe_tables.h"
#include "mesh.h"

namespace ifem
{
  struct ContactScan
  {
    int dim, degree, nv;
    std::vector<double> dG; // [2 dim faces][nv][dim]: Q1 geometry gradients at the first face quadrature point

    ContactScan(int dim_, int degree_) : dim(dim_), degree(degree_), nv(1 << dim_)
    {
      FEQ feg(dim, 1);
      Quadrature fq(dim - 1, degree + 1);
      dG.resize((size_t)2 * dim * nv * dim);
      std::vector<double> N(nv);
      for (int face = 0; face < 2 * dim; ++face)
        {
          double xi[3];
          int k = 0;
          for (int d = 0; d < dim; ++d) xi[d] = (d == face / 2) ? double(face % 2) : fq.points[k++];
          feg.eval(xi, N.data(), &dG[(size_t)face * nv * dim]);
        }
    }

    // For every (cell, boundary face, face vertex): penetration = criterion(moved vertex); if > 1e-5 add
    // extra_stress[d][dim - 1] = (multiplier * penetration / |direction| * direction[d]) / n[d] (0 unless n[d] > 1e-5, n = the
    // face normal at its first quadrature point on the moved mesh) to rows[d][dim * node + dim - 1]. Returns still_penetrate.
    bool run(int n_bfaces, const int *boundary_faces /*[n][3] = cell, face, id*/, const int *cell_nodes, int npc, const double *coords,
             const double *u, int64_t n_dofs, const std::function<double(const double *)> &criterion, const double *direction,
             double multiplier, double *rows /*[dim][n_dofs]*/) const
    {
      double dir_norm = 0.0;
      for (int d = 0; d < dim; ++d) dir_norm += direction[d] * direction[d];
      dir_norm = std::sqrt(dir_norm);
      bool still_penetrate = false;
      for (int f = 0; f < n_bfaces; ++f)
        {
          const int cell = boundary_faces[3 * f], face = boundary_faces[3 * f + 1], axis = face / 2, side = face % 2;
          const int *cn = cell_nodes + (size_t)cell * npc;
          // normal of the moved face at its first quadrature point: row `axis` of det(J) J^-1, outward
          double J[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, nds[3] = {0, 0, 0};
          for (int v = 0; v < nv; ++v)
            for (int i = 0; i < dim; ++i)
              {
                const double xv = coords[(size_t)cn[v] * dim + i] + u[(size_t)cn[v] * dim + i];
                for (int j = 0; j < dim; ++j) J[i * dim + j] += xv * dG[((size_t)face * nv + v) * dim + j];
              }
          const double sgn = side ? 1.0 : -1.0;
          if (dim == 2)
            {
              // det * Jinv = [[J11, -J01], [-J10, J00]]
              const double adj[4] = {J[3], -J[1], -J[2], J[0]};
              for (int k = 0; k < 2; ++k) nds[k] = adj[axis * 2 + k] * sgn;
            }
          else
            {
              // det * Jinv[axis][k] = cofactor(J)[k][axis]
              const int r1 = (axis + 1) % 3, r2 = (axis + 2) % 3;
              for (int k = 0; k < 3; ++k)
                {
                  const int k1 = (k + 1) % 3, k2 = (k + 2) % 3;
                  nds[k] = (J[k1 * 3 + r1] * J[k2 * 3 + r2] - J[k1 * 3 + r2] * J[k2 * 3 + r1]) * sgn;
                }
            }
          double dS = 0.0;
          for (int k = 0; k < dim; ++k) dS += nds[k] * nds[k];
          dS = std::sqrt(dS);
          for (int a : face_local_nodes(dim, degree, face))
            {
              const int node = cn[a];
              double x[3] = {0, 0, 0};
              for (int d = 0; d < dim; ++d) x[d] = coords[(size_t)node * dim + d] + u[(size_t)node * dim + d];
              const double penetration_value = criterion(x);
              if (!(penetration_value > 1e-5)) continue;
              still_penetrate = true;
              for (int d1 = 0; d1 < dim; ++d1)
                {
                  const double traction = multiplier * penetration_value / dir_norm * direction[d1];
                  const double nd = nds[d1] / dS;
                  const double extra = nd > 1e-5 ? traction / nd : 0.0; // extra_stress[d1][dim - 1]
                  rows[(size_t)d1 * n_dofs + (size_t)dim * node + dim - 1] += extra;
                }
            }
        }
      return still_penetrate;
    }
  };
} // namespace ifem
