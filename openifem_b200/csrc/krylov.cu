#include "krylov.h"

namespace ifem
{
  SolveResult cg(Context &ctx, const VecSpace &n, const LinOp &A, const double *b, double *x, bool x_is_zero, double tol_abs, int max_it,
                 VecPool &pool)
  {
    double *r = pool.get(0, n.n_alloc), *p = pool.get(1, n.n_alloc), *Ap = pool.get(2, n.n_alloc);
    SolveResult out;
    if (x_is_zero)
      copy(ctx, n, b, r);
    else
      {
        A(x, Ap);
        lin3(ctx, n, r, b, -1.0, Ap, 0.0, Ap);
      }
    double rr = dot(ctx, n, r, r);
    out.residual = std::sqrt(rr);
    if (out.residual <= tol_abs)
      {
        out.converged = true;
        return out;
      }
    copy(ctx, n, r, p);
    while (out.iterations < max_it)
      {
        A(p, Ap);
        const double alpha = rr / dot(ctx, n, p, Ap);
        axpy(ctx, n, alpha, p, x);
        // r -= alpha Ap ; rr_new = r . r   (fused)
        const double rr_new = add_and_dot(ctx, n, r, -alpha, Ap, r);
        out.iterations++;
        out.residual = std::sqrt(rr_new);
        if (out.residual <= tol_abs)
          {
            out.converged = true;
            break;
          }
        axpby(ctx, n, 1.0, r, rr_new / rr, p); // p = r + beta p
        rr = rr_new;
      }
    return out;
  }

  SolveResult pcg(Context &ctx, const VecSpace &n, const LinOp &A, const LinOp &prec, const double *b, double *x, double tol_abs,
                  int max_it, VecPool &pool)
  {
    double *r = pool.get(0, n.n_alloc), *p = pool.get(1, n.n_alloc), *Ap = pool.get(2, n.n_alloc), *z = pool.get(3, n.n_alloc);
    SolveResult out;
    fill(ctx, n, 0.0, x);
    copy(ctx, n, b, r);
    out.residual = nrm2(ctx, n, r);
    if (out.residual <= tol_abs)
      {
        out.converged = true;
        return out;
      }
    prec(r, z);
    copy(ctx, n, z, p);
    double rz = dot(ctx, n, r, z);
    while (out.iterations < max_it)
      {
        A(p, Ap);
        const double alpha = rz / dot(ctx, n, p, Ap);
        axpy(ctx, n, alpha, p, x);
        const double rr = add_and_dot(ctx, n, r, -alpha, Ap, r);
        out.iterations++;
        out.residual = std::sqrt(rr);
        if (out.residual <= tol_abs)
          {
            out.converged = true;
            break;
          }
        prec(r, z);
        const double rz_new = dot(ctx, n, r, z);
        axpby(ctx, n, 1.0, z, rz_new / rz, p);
        rz = rz_new;
      }
    return out;
  }

  SolveResult bicgstab(Context &ctx, const VecSpace &n, const LinOp &A, const LinOp &prec, const double *b, double *x, double tol_abs,
                       int max_it, VecPool &pool)
  {
    double *r = pool.get(0, n.n_alloc), *r0 = pool.get(1, n.n_alloc), *p = pool.get(2, n.n_alloc), *v = pool.get(3, n.n_alloc), *ph = pool.get(4, n.n_alloc),
           *s = pool.get(5, n.n_alloc), *sh = pool.get(6, n.n_alloc), *t = pool.get(7, n.n_alloc);
    SolveResult out;
    fill(ctx, n, 0.0, x);
    copy(ctx, n, b, r);
    out.residual = nrm2(ctx, n, r);
    if (out.residual <= tol_abs)
      {
        out.converged = true;
        return out;
      }
    copy(ctx, n, r, r0);
    fill(ctx, n, 0.0, v);
    fill(ctx, n, 0.0, p);
    double rho = 1.0, alpha = 1.0, omega = 1.0;
    while (out.iterations < max_it)
      {
        const double rho_new = dot(ctx, n, r0, r);
        const double beta = (rho_new / rho) * (alpha / omega);
        // p = r + beta (p - omega v)
        lin3(ctx, n, p, r, beta, p, -beta * omega, v);
        prec(p, ph);
        A(ph, v);
        alpha = rho_new / dot(ctx, n, r0, v);
        lin3(ctx, n, s, r, -alpha, v, 0.0, v);
        out.iterations++;
        out.residual = nrm2(ctx, n, s);
        if (out.residual <= tol_abs)
          {
            axpy(ctx, n, alpha, ph, x);
            out.converged = true;
            break;
          }
        prec(s, sh);
        A(sh, t);
        omega = dot(ctx, n, t, s) / dot(ctx, n, t, t);
        lin3(ctx, n, x, x, alpha, ph, omega, sh);
        lin3(ctx, n, r, s, -omega, t, 0.0, t);
        out.residual = nrm2(ctx, n, r);
        rho = rho_new;
        if (out.residual <= tol_abs)
          {
            out.converged = true;
            break;
          }
      }
    return out;
  }

  SolveResult fgmres(Context &ctx, const VecSpace &n, const LinOp &A, const LinOp &prec, const double *b, double *x, double tol_abs,
                     int64_t max_it, int m, VecPool &pool, bool fused_orthogonalisation)
  {
    fused_orthogonalisation = fused_orthogonalisation && m <= 64;
    std::vector<const double *> basis;
    std::vector<double> hcol;
    // pool slots: 0 aux, 1..m V, m+1..2m Z
    double *aux = pool.get(0, n.n_alloc);
    auto V = [&](int j) { return pool.get(1 + j, n.n_alloc); };
    auto Z = [&](int j) { return pool.get(1 + m + j, n.n_alloc); };
    SolveResult out;
    fill(ctx, n, 0.0, x);
    int64_t accumulated = 0;
    bool first = true;
    std::vector<double> H((size_t)(m + 1) * m), R((size_t)(m + 1) * m), cs(m), sn(m), g(m + 1), y;
    while (true)
      {
        // aux = b - A x
        if (first)
          copy(ctx, n, b, aux);
        else
          {
            A(x, aux);
            axpby(ctx, n, 1.0, b, -1.0, aux);
          }
        first = false;
        const double beta = nrm2(ctx, n, aux);
        if (!std::isfinite(beta)) throw std::runtime_error("FGMRES: non-finite residual (the preconditioner returned NaN/Inf)");
        out.residual = beta;
        if (beta <= tol_abs)
          {
            out.converged = true;
            break;
          }
        if (accumulated >= max_it) break;
        std::fill(g.begin(), g.end(), 0.0);
        g[0] = beta;
        double a = beta;
        y.clear();
        bool stop = false;
        for (int j = 0; j < m && !stop; ++j)
          {
            if (a != 0.0)
              equ(ctx, n, 1.0 / a, aux, V(j));
            else
              fill(ctx, n, 0.0, V(j));
            prec(V(j), Z(j));
            A(Z(j), aux);
            // modified Gram-Schmidt via add_and_dot
            auto h = [&](int i) -> double & { return H[(size_t)i * m + j]; };
            if (fused_orthogonalisation)
              {
                basis.resize(j + 1);
                hcol.resize(j + 1);
                for (int i = 0; i <= j; ++i) basis[i] = V(i);
                a = orthogonalise_cgs2(ctx, n, j + 1, basis.data(), aux, hcol.data());
                for (int i = 0; i <= j; ++i) h(i) = hcol[i];
              }
            else
              {
                h(0) = dot(ctx, n, aux, V(0));
                for (int i = 1; i <= j; ++i) h(i) = add_and_dot(ctx, n, aux, -h(i - 1), V(i - 1), V(i));
                a = std::sqrt(add_and_dot(ctx, n, aux, -h(j), V(j), aux));
              }
            if (!std::isfinite(a)) throw std::runtime_error("FGMRES: non-finite Arnoldi vector (the preconditioner returned NaN/Inf)");
            h(j + 1) = a;
            // least squares on the (j+1) x j block = all columns before this one:
            // rotate the new column with the previous rotations, residual = |g[j]|
            auto r = [&](int i) -> double & { return R[(size_t)i * m + j]; };
            for (int i = 0; i <= j + 1; ++i) r(i) = h(i);
            for (int i = 0; i < j; ++i)
              {
                const double t0 = cs[i] * r(i) + sn[i] * r(i + 1), t1 = -sn[i] * r(i) + cs[i] * r(i + 1);
                r(i) = t0;
                r(i + 1) = t1;
              }
            if (j > 0)
              {
                out.residual = std::fabs(g[j]);
                ++accumulated;
                // y from the j x j triangular system
                y.assign(j, 0.0);
                for (int i = j - 1; i >= 0; --i)
                  {
                    double sacc = g[i];
                    for (int k = i + 1; k < j; ++k) sacc -= R[(size_t)i * m + k] * y[k];
                    y[i] = sacc / R[(size_t)i * m + i];
                  }
                if (out.residual <= tol_abs)
                  {
                    out.converged = true;
                    stop = true;
                  }
                else if (accumulated >= max_it)
                  stop = true;
              }
            if (!stop)
              {
                // new rotation eliminating r(j+1)
                const double d = std::hypot(r(j), r(j + 1));
                cs[j] = d != 0.0 ? r(j) / d : 1.0;
                sn[j] = d != 0.0 ? r(j + 1) / d : 0.0;
                r(j) = d;
                r(j + 1) = 0.0;
                g[j + 1] = -sn[j] * g[j];
                g[j] = cs[j] * g[j];
              }
          }
        for (size_t j = 0; j < y.size(); ++j) axpy(ctx, n, y[j], Z((int)j), x);
        if (stop) break;
      }
    out.iterations = (int)accumulated;
    if (!out.converged) throw std::runtime_error("FGMRES: no convergence");
    return out;
  }
} // namespace ifem
