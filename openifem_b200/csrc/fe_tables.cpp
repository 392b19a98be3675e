#include "fe_tables.h"

#include <cmath>
#include <stdexcept>

namespace ifem
{
  void gauss_legendre_01(int n, std::vector<double> &x, std::vector<double> &w)
  {
    // Newton iteration on the Legendre polynomial P_n (roots in (-1,1)), then
    // mapped to [0,1]. Same nodes/weights as dealii::QGauss<1>(n).
    x.assign(n, 0.0);
    w.assign(n, 0.0);
    const double pi = 3.14159265358979323846264338327950288;
    for (int i = 0; i < (n + 1) / 2; ++i)
      {
        long double z = std::cos(pi * (i + 0.75) / (n + 0.5));
        long double pp = 0;
        for (int it = 0; it < 100; ++it)
          {
            long double p1 = 1.0, p2 = 0.0;
            for (int j = 0; j < n; ++j)
              {
                const long double p3 = p2;
                p2 = p1;
                p1 = ((2.0L * j + 1.0L) * z * p2 - j * p3) / (j + 1.0L);
              }
            pp = n * (z * p1 - p2) / (z * z - 1.0L);
            const long double z1 = z;
            z = z1 - p1 / pp;
            if (std::fabs((double)(z - z1)) < 1e-18) break;
          }
        const double xm = (double)z;
        const double wt = (double)(2.0L / ((1.0L - z * z) * pp * pp));
        x[i] = 0.5 * (1.0 - xm);
        x[n - 1 - i] = 0.5 * (1.0 + xm);
        w[i] = w[n - 1 - i] = 0.5 * wt;
      }
  }

  void lagrange_1d(int p, double x, double *val, double *der)
  {
    if (p < 1 || p > 4) throw std::runtime_error("lagrange_1d: unsupported degree");
    double nodes[5];
    for (int i = 0; i <= p; ++i) nodes[i] = double(i) / p;
    for (int i = 0; i <= p; ++i)
      {
        double v = 1.0;
        for (int j = 0; j <= p; ++j)
          if (j != i) v *= (x - nodes[j]) / (nodes[i] - nodes[j]);
        val[i] = v;
        double d = 0.0;
        for (int k = 0; k <= p; ++k)
          {
            if (k == i) continue;
            double t = 1.0 / (nodes[i] - nodes[k]);
            for (int j = 0; j <= p; ++j)
              if (j != i && j != k) t *= (x - nodes[j]) / (nodes[i] - nodes[j]);
            d += t;
          }
        der[i] = d;
      }
  }

  FEQ::FEQ(int dim_, int p_) : dim(dim_), p(p_), n1(p_ + 1)
  {
    n = 1;
    for (int d = 0; d < dim; ++d) n *= n1;
    lattice.resize(n);
    for (int a = 0; a < n; ++a)
      {
        int r = a;
        std::array<int, 3> l{0, 0, 0};
        for (int d = 0; d < dim; ++d)
          {
            l[d] = r % n1;
            r /= n1;
          }
        lattice[a] = l;
      }
  }

  void FEQ::eval(const double *xi, double *N, double *dN) const
  {
    double V[3][5], D[3][5];
    for (int d = 0; d < dim; ++d) lagrange_1d(p, xi[d], V[d], D[d]);
    for (int a = 0; a < n; ++a)
      {
        double v = 1.0;
        for (int d = 0; d < dim; ++d) v *= V[d][lattice[a][d]];
        N[a] = v;
        for (int e = 0; e < dim; ++e)
          {
            double g = 1.0;
            for (int d = 0; d < dim; ++d) g *= (d == e) ? D[d][lattice[a][d]] : V[d][lattice[a][d]];
            dN[a * dim + e] = g;
          }
      }
  }

  Quadrature::Quadrature(int dim_, int n_1d) : dim(dim_)
  {
    std::vector<double> x, w;
    gauss_legendre_01(n_1d, x, w);
    nq = 1;
    for (int d = 0; d < dim; ++d) nq *= n_1d;
    points.resize((size_t)nq * dim);
    weights.resize(nq);
    for (int q = 0; q < nq; ++q)
      {
        int r = q;
        double wt = 1.0;
        for (int d = 0; d < dim; ++d)
          {
            const int i = r % n_1d;
            r /= n_1d;
            points[(size_t)q * dim + d] = x[i];
            wt *= w[i];
          }
        weights[q] = wt;
      }
  }

  ShapeTable::ShapeTable(const FEQ &fe, const std::vector<double> &points, int npts) : nq(npts), n(fe.n), dim(fe.dim)
  {
    N.resize((size_t)nq * n);
    dN.resize((size_t)nq * n * dim);
    for (int q = 0; q < nq; ++q) fe.eval(&points[(size_t)q * dim], &N[(size_t)q * n], &dN[(size_t)q * n * dim]);
  }
} // namespace ifem
