// Host-side finite element tables: Lagrange FE_Q(p) on [0,1]^dim with equidistant
// support points (p <= 2), tensor Gauss-Legendre quadrature QGauss(n), and the
// evaluation of shape values / reference gradients at arbitrary points.
// These replace the deal.II objects the reference constructs in
// source/mpi_fluid_solver.cpp:27-35 (FESystem(FE_Q(pu)^dim, FE_Q(pp)), QGauss(pu+1))
// and source/mpi_solid_solver.cpp:17-20.
#pragma once
#include <array>
#include <vector>

namespace ifem
{
  // 1-D Gauss-Legendre on [0,1]
  void gauss_legendre_01(int n, std::vector<double> &x, std::vector<double> &w);

  // values / derivatives of the p+1 equidistant Lagrange polynomials at x
  void lagrange_1d(int p, double x, double *val, double *der);

  struct FEQ
  {
    int dim = 0, p = 0, n1 = 0, n = 0;
    std::vector<std::array<int, 3>> lattice; // local node -> (ix,iy,iz), x fastest
    FEQ() = default;
    FEQ(int dim, int p);
    // N[n], dN[n][dim] at a reference point
    void eval(const double *xi, double *N, double *dN) const;
  };

  struct Quadrature
  {
    int dim = 0, nq = 0;
    std::vector<double> points; // [nq][dim], x fastest
    std::vector<double> weights;
    Quadrature() = default;
    Quadrature(int dim, int n_1d);
  };

  // Tables of an FEQ evaluated on a quadrature: N[nq][n], dN[nq][n][dim]
  struct ShapeTable
  {
    int nq = 0, n = 0, dim = 0;
    std::vector<double> N, dN;
    ShapeTable() = default;
    ShapeTable(const FEQ &fe, const std::vector<double> &points, int npts);
  };
} // namespace ifem
