#include "hanging.h"

#include <algorithm>
#include <cmath>
#include <unordered_map>

#include "fluid.h"

namespace ifem
{
  namespace
  {
    // nodes of a table by quantised position (the hanging vertices of the triangulation are matched to FE_Q(1) nodes by
    // their coordinates: the node tables are renumbered spatially and, on several ranks, localised)
    struct NodeLookup
    {
      int dim = 0;
      double lo[3] = {0, 0, 0}, ext[3] = {1, 1, 1};
      std::unordered_map<uint64_t, int> at;
      uint64_t key(const double *x) const
      {
        uint64_t k = 0;
        for (int d = dim - 1; d >= 0; --d) k = (k << 21) | (uint64_t)std::llround((x[d] - lo[d]) / ext[d] * double((1u << 21) - 2));
        return k;
      }
      NodeLookup(const Triangulation &tria, const NodeTable &nt) : dim(tria.dim)
      {
        double hi[3] = {0, 0, 0};
        for (int d = 0; d < dim; ++d) lo[d] = hi[d] = tria.vertices[d];
        for (int i = 0; i < tria.n_vertices(); ++i)
          for (int d = 0; d < dim; ++d)
            {
              lo[d] = std::min(lo[d], tria.vertices[(size_t)i * dim + d]);
              hi[d] = std::max(hi[d], tria.vertices[(size_t)i * dim + d]);
            }
        for (int d = 0; d < dim; ++d) ext[d] = hi[d] > lo[d] ? hi[d] - lo[d] : 1.0;
        at.reserve((size_t)nt.n_nodes * 2);
        for (int i = 0; i < nt.n_nodes; ++i) at.emplace(key(&nt.coords[(size_t)i * dim]), i);
      }
      int find(const double *x) const
      {
        auto it = at.find(key(x));
        return it == at.end() ? -1 : it->second;
      }
    };
  } // namespace

  std::vector<char> hanging_node_flags(const Triangulation &tria, const NodeTable &nt)
  {
    std::vector<char> flag((size_t)nt.n_nodes, 0);
    if (tria.hanging.empty()) return flag;
    if (nt.p != 1) throw std::runtime_error("hanging-node constraints are implemented for FE_Q(1) spaces only");
    const NodeLookup look(tria, nt);
    for (const auto &h : tria.hanging)
      {
        const int n = look.find(&tria.vertices[(size_t)h.vertex * tria.dim]);
        if (n >= 0) flag[n] = 1;
      }
    return flag;
  }

  void HangingConstraints::find(const Triangulation &tria, const FluidSpace &fs, std::vector<int> &cell_un_ext, std::vector<int> &cell_pn_ext,
                                int &width)
  {
    active = !tria.hanging.empty();
    width = 0;
    if (!active) return;
    if (fs.pu != 1 || fs.pp != 1)
      throw std::runtime_error("locally refined mesh: hanging-node constraints are implemented for FE_Q(1) velocity and pressure only "
                               "(SCnsIM / SUPGInsIM, the solvers of the reference's refined cases)");
    const int dim = tria.dim;
    auto build = [&](const NodeTable &nt, int n_layer1, HangingNodes &H) {
      const NodeLookup look(tria, nt);
      std::vector<std::array<int, 6>> lines; // node, n_masters, masters
      for (const auto &h : tria.hanging)
        {
          const int n = look.find(&tria.vertices[(size_t)h.vertex * dim]);
          if (n < 0) continue; // not on this rank
          std::array<int, 6> l{n, h.n_masters, -1, -1, -1, -1};
          bool complete = true;
          for (int k = 0; k < h.n_masters; ++k)
            {
              l[2 + k] = look.find(&tria.vertices[(size_t)h.master[k] * dim]);
              complete = complete && l[2 + k] >= 0;
            }
          if (!complete)
            {
              if (n < n_layer1) throw std::runtime_error("hanging node next to the owned range of this rank whose masters are not local");
              continue; // outer ghost layer: never a column of an owned row
            }
          std::sort(l.begin() + 2, l.begin() + 2 + h.n_masters);
          lines.push_back(l);
        }
      std::sort(lines.begin(), lines.end());
      H.n = (int)lines.size();
      H.node.resize(H.n);
      H.n_masters.resize(H.n);
      H.master.assign((size_t)H.n * 4, -1);
      for (int i = 0; i < H.n; ++i)
        {
          H.node[i] = lines[i][0];
          H.n_masters[i] = lines[i][1];
          for (int k = 0; k < 4; ++k) H.master[(size_t)i * 4 + k] = lines[i][2 + k];
        }
    };
    build(fs.un, fs.n_layer1_unodes, u);
    build(fs.pn, fs.n_layer1_pnodes, p);

    // per-cell node lists extended by the masters of the cell's hanging nodes: C^T A C couples them with every node of the cell
    auto extend = [&](const NodeTable &nt, const HangingNodes &H, std::vector<std::vector<int>> &lists) {
      std::vector<int> idx((size_t)nt.n_nodes, -1);
      for (int i = 0; i < H.n; ++i) idx[H.node[i]] = i;
      const int npc = nt.nodes_per_cell;
      lists.resize(fs.n_cells);
      for (int c = 0; c < fs.n_cells; ++c)
        {
          std::vector<int> &l = lists[c];
          l.assign(nt.cell_nodes.begin() + (size_t)c * npc, nt.cell_nodes.begin() + (size_t)(c + 1) * npc);
          for (int a = 0; a < npc; ++a)
            {
              const int i = idx[l[a]];
              if (i < 0) continue;
              for (int k = 0; k < H.n_masters[i]; ++k)
                {
                  const int m = H.master[(size_t)i * 4 + k];
                  if (std::find(l.begin(), l.end(), m) == l.end()) l.push_back(m);
                }
            }
          width = std::max(width, (int)l.size());
        }
    };
    std::vector<std::vector<int>> lu, lp;
    extend(fs.un, u, lu);
    extend(fs.pn, p, lp);
    auto flatten = [&](const std::vector<std::vector<int>> &lists, std::vector<int> &flat) {
      flat.resize((size_t)fs.n_cells * width);
      for (int c = 0; c < fs.n_cells; ++c)
        for (int a = 0; a < width; ++a) flat[(size_t)c * width + a] = lists[c][a < (int)lists[c].size() ? a : 0]; // padded with a repeat
    };
    flatten(lu, cell_un_ext);
    flatten(lp, cell_pn_ext);
    is_hanging_dof.assign((size_t)fs.n_dofs, 0);
    for (int i = 0; i < u.n; ++i)
      for (int c = 0; c < dim; ++c) is_hanging_dof[(size_t)dim * u.node[i] + c] = 1;
    for (int i = 0; i < p.n; ++i) is_hanging_dof[(size_t)fs.n_u + p.node[i]] = 1;
  }

  // ---------------------------------------------------------------------------------------------------------------
  // fold plans
  // ---------------------------------------------------------------------------------------------------------------
  namespace
  {
    void make_plan(Context &ctx, const Pattern &P, int n_rows, const HangingNodes &rows, const HangingNodes &cols, int n_row_nodes,
                   int n_col_nodes, FoldPlan &plan)
    {
      std::vector<int> cidx((size_t)n_col_nodes, -1), ridx((size_t)n_row_nodes, -1);
      for (int i = 0; i < cols.n; ++i) cidx[cols.node[i]] = i;
      for (int i = 0; i < rows.n; ++i) ridx[rows.node[i]] = i;
      std::vector<int> row, ptr{0}, item;
      for (int r = 0; r < n_rows; ++r)
        {
          const int *b = &P.col[P.rowptr[r]], *e = &P.col[P.rowptr[r + 1]];
          bool any = false;
          for (const int *c = b; c != e; ++c)
            {
              const int h = cidx[*c];
              if (h < 0) continue;
              any = true;
              item.push_back((int)(c - b));
              item.push_back(h);
              for (int k = 0; k < 4; ++k)
                {
                  int pos = -1;
                  if (k < cols.n_masters[h])
                    {
                      const int *it = std::lower_bound(b, e, cols.master[(size_t)h * 4 + k]);
                      if (it == e || *it != cols.master[(size_t)h * 4 + k])
                        throw std::runtime_error("hanging-node condensation: the column of a master is missing from the pattern");
                      pos = (int)(it - b);
                    }
                  item.push_back(pos);
                }
            }
          if (any)
            {
              row.push_back(r);
              ptr.push_back((int)item.size() / 6);
            }
        }
      plan.n_rows = (int)row.size();
      // master rows: the rows of hanging nodes are added to the rows of their masters
      std::vector<std::pair<int, int>> ms; // (master row, hanging index)
      for (int i = 0; i < rows.n; ++i)
        {
          const bool owned = rows.node[i] < n_rows;
          for (int k = 0; k < rows.n_masters[i]; ++k)
            {
              const int m = rows.master[(size_t)i * 4 + k];
              if ((m < n_rows) != owned)
                throw std::runtime_error("hanging-node condensation: a hanging node and one of its masters are owned by different ranks "
                                         "(cut the partition along the refinement interface)");
              if (owned) ms.emplace_back(m, i);
            }
        }
      std::sort(ms.begin(), ms.end());
      std::vector<int> mrow, mptr{0}, mslave;
      for (size_t k = 0; k < ms.size(); ++k)
        {
          if (k == 0 || ms[k].first != ms[k - 1].first)
            {
              if (k) mptr.push_back((int)mslave.size());
              mrow.push_back(ms[k].first);
            }
          mslave.push_back(ms[k].second);
        }
      mptr.push_back((int)mslave.size());
      plan.n_mrows = (int)mrow.size();
      cudaStream_t s = ctx.stream;
      if (plan.n_rows)
        {
          plan.d_row.upload(row, s);
          plan.d_ptr.upload(ptr, s);
          plan.d_item.upload(item, s);
        }
      if (plan.n_mrows)
        {
          plan.d_mrow.upload(mrow, s);
          plan.d_mptr.upload(mptr, s);
          plan.d_mslave.upload(mslave, s);
        }
      IFEM_CUDA(cudaStreamSynchronize(s));
    }

    void upload_nodes(Context &ctx, HangingNodes &H, int ncomp)
    {
      if (!H.n) return;
      H.d_node.upload(H.node, ctx.stream);
      H.d_n_masters.upload(H.n_masters, ctx.stream);
      H.d_master.upload(H.master, ctx.stream);
      H.d_diag.alloc((size_t)H.n * ncomp);
      IFEM_CUDA(cudaStreamSynchronize(ctx.stream));
    }
  } // namespace

  void HangingConstraints::plan(Context &ctx, const FluidSpace &fs)
  {
    if (!active) return;
    with_pp = fs.A_pp.n_brows > 0;
    if (!with_pp) throw std::runtime_error("hanging-node condensation needs the pressure-pressure block (SCnsIM / SUPGInsIM)");
    upload_nodes(ctx, u, fs.dim);
    upload_nodes(ctx, p, 1);
    d_is_hanging_dof.upload(is_hanging_dof, ctx.stream);
    make_plan(ctx, fs.P_uu, fs.n_owned_unodes, u, u, fs.un.n_nodes, fs.un.n_nodes, uu);
    make_plan(ctx, fs.P_up, fs.n_owned_unodes, u, p, fs.un.n_nodes, fs.pn.n_nodes, up);
    make_plan(ctx, fs.P_pu, fs.n_owned_pnodes, p, u, fs.pn.n_nodes, fs.un.n_nodes, pu);
    make_plan(ctx, fs.P_pp, fs.n_owned_pnodes, p, p, fs.pn.n_nodes, fs.pn.n_nodes, pp);
  }

  // ---------------------------------------------------------------------------------------------------------------
  // device passes
  // ---------------------------------------------------------------------------------------------------------------
  namespace
  {
    struct Mat
    {
      double *val;
      const int64_t *rowptr;
      const int *col;
      int R, C;
      __device__ double &at(int row, int r, int c, int j) const
      {
        const int64_t base = rowptr[row];
        const int nb = (int)(rowptr[row + 1] - base);
        return val[base * R * C + (int64_t)(r * C + c) * nb + j];
      }
      __device__ int find(int row, int column) const
      {
        const int64_t base = rowptr[row];
        int lo = 0, hi = (int)(rowptr[row + 1] - base) - 1;
        while (lo <= hi)
          {
            const int mid = (lo + hi) >> 1;
            const int c = col[base + mid];
            if (c == column) return mid;
            if (c < column) lo = mid + 1; else hi = mid - 1;
          }
        return -1;
      }
    };
    // dofs of a node space inside the block vector: dof = offset + ncomp * node + component
    struct Space
    {
      int ncomp;
      int64_t offset;
      const int *h_node, *h_nm, *h_master; // hanging nodes of the space
      __device__ int64_t dof(int node, int c) const { return offset + (int64_t)ncomp * node + c; }
    };

    Mat view(const Bcsr &A) { return Mat{A.val.p, A.rowptr.p, A.col.p, A.R, A.C}; }

    // |diagonal| of the hanging rows before anything is folded (the value a constrained row keeps)
    __global__ void hanging_diag_kernel(int n, int n_owned, Mat M, Space X, double *__restrict__ diag)
    {
      const int i = blockIdx.x * blockDim.x + threadIdx.x;
      if (i >= n) return;
      const int h = X.h_node[i];
      if (h < n_owned)
        {
          const int j = M.find(h, h);
          for (int r = 0; r < M.R; ++r)
            {
              const double d = j >= 0 ? fabs(M.at(h, r, r, j)) : 0.0;
              diag[(size_t)i * M.R + r] = d > 0.0 ? d : 1.0;
            }
        }
    }

    // columns of hanging nodes -> columns of their masters; a master with a Dirichlet line moves to the right-hand side.
    // One thread per row, items in pattern order.
    __global__ void fold_columns_kernel(int n_rows, const int *__restrict__ rows, const int *__restrict__ ptr, const int *__restrict__ item,
                                        Mat M, Space X, Space Y, const unsigned char *__restrict__ con, const double *__restrict__ inhom,
                                        double *__restrict__ rhs)
    {
      const int t = blockIdx.x * blockDim.x + threadIdx.x;
      if (t >= n_rows) return;
      const int row = rows[t];
      for (int it = ptr[t]; it < ptr[t + 1]; ++it)
        {
          const int *I = item + (size_t)it * 6;
          const int jh = I[0], h = I[1], nm = Y.h_nm[h];
          const double w = 1.0 / nm;
          for (int r = 0; r < M.R; ++r)
            for (int c = 0; c < M.C; ++c)
              {
                double &src = M.at(row, r, c, jh);
                const double v = src;
                if (v == 0.0) continue;
                for (int k = 0; k < nm; ++k)
                  {
                    const int64_t md = Y.dof(Y.h_master[h * 4 + k], c);
                    if (con[md])
                      {
                        if (inhom) rhs[X.dof(row, r)] -= w * v * inhom[md];
                      }
                    else
                      M.at(row, r, c, I[2 + k]) += w * v;
                  }
                src = 0.0;
              }
        }
    }

    // rows of hanging nodes -> rows of their masters (components whose master dof carries a Dirichlet line are skipped).
    // One thread per master row, slaves in ascending order.
    __global__ void fold_rows_kernel(int n_mrows, const int *__restrict__ mrow, const int *__restrict__ mptr, const int *__restrict__ mslave,
                                     Mat M, Space X, const unsigned char *__restrict__ con, double *__restrict__ rhs, int with_rhs)
    {
      const int t = blockIdx.x * blockDim.x + threadIdx.x;
      if (t >= n_mrows) return;
      const int m = mrow[t];
      for (int s = mptr[t]; s < mptr[t + 1]; ++s)
        {
          const int hi = mslave[s], h = X.h_node[hi];
          const double w = 1.0 / X.h_nm[hi];
          const int64_t base = M.rowptr[h];
          const int nb = (int)(M.rowptr[h + 1] - base);
          for (int j = 0; j < nb; ++j)
            {
              bool any = false;
              for (int r = 0; r < M.R && !any; ++r)
                for (int c = 0; c < M.C; ++c)
                  if (M.at(h, r, c, j) != 0.0) any = true;
              if (!any) continue;
              const int jm = M.find(m, M.col[base + j]);
              if (jm < 0) continue; // cannot happen: the pattern holds every master coupling (checked when the plan was made)
              for (int r = 0; r < M.R; ++r)
                {
                  if (con[X.dof(m, r)]) continue;
                  for (int c = 0; c < M.C; ++c) M.at(m, r, c, jm) += w * M.at(h, r, c, j);
                }
            }
          if (with_rhs)
            for (int r = 0; r < M.R; ++r)
              if (!con[X.dof(m, r)]) rhs[X.dof(m, r)] += w * rhs[X.dof(h, r)];
        }
    }

    // a hanging row keeps its diagonal only; rhs = diagonal x inhomogeneity of the line (the Dirichlet values of its masters)
    __global__ void hanging_rows_kernel(int n, int n_owned, Mat Md, Mat Mo, Space X, const double *__restrict__ diag,
                                        const unsigned char *__restrict__ con, const double *__restrict__ inhom, double *__restrict__ rhs)
    {
      const int i = blockIdx.x * blockDim.x + threadIdx.x;
      if (i >= n) return;
      const int h = X.h_node[i];
      if (h >= n_owned) return; // the row lives on the owner of the node
      const int nm = X.h_nm[i];
      for (int pass = 0; pass < 2; ++pass)
        {
          const Mat &M = pass ? Mo : Md;
          const int nb = (int)(M.rowptr[h + 1] - M.rowptr[h]);
          for (int r = 0; r < M.R; ++r)
            for (int c = 0; c < M.C; ++c)
              for (int j = 0; j < nb; ++j) M.at(h, r, c, j) = 0.0;
        }
      const int jd = Md.find(h, h);
      for (int r = 0; r < Md.R; ++r)
        {
          const double d = diag[(size_t)i * Md.R + r];
          Md.at(h, r, r, jd) = d;
          double g = 0.0;
          if (inhom)
            for (int k = 0; k < nm; ++k)
              {
                const int64_t md = X.dof(X.h_master[i * 4 + k], r);
                if (con[md]) g += inhom[md] / nm;
              }
          rhs[X.dof(h, r)] = d * g;
        }
    }

    __global__ void hanging_distribute_kernel(int n, Space X, double *__restrict__ x)
    {
      const int i = blockIdx.x * blockDim.x + threadIdx.x;
      if (i >= n) return;
      const int h = X.h_node[i], nm = X.h_nm[i];
      for (int c = 0; c < X.ncomp; ++c)
        {
          double s = 0.0;
          for (int k = 0; k < nm; ++k) s += x[X.dof(X.h_master[i * 4 + k], c)];
          x[X.dof(h, c)] = s / nm;
        }
    }

    inline unsigned blocks(int n) { return (unsigned)((n + 127) / 128); }
  } // namespace

  void HangingConstraints::condense(Context &ctx, FluidSpace &fs, const double *inhom) const
  {
    if (!active) return;
    cudaStream_t s = ctx.stream;
    const Space U{fs.dim, 0, u.d_node.p, u.d_n_masters.p, u.d_master.p}, P{1, fs.n_u, p.d_node.p, p.d_n_masters.p, p.d_master.p};
    const Mat uu_ = view(fs.A_uu), up_ = view(fs.A_up), pu_ = view(fs.A_pu), pp_ = view(fs.A_pp);
    const unsigned char *con = fs.d_con.p;
    double *rhs = fs.rhs.p;
    if (u.n) hanging_diag_kernel<<<blocks(u.n), 128, 0, s>>>(u.n, fs.n_owned_unodes, uu_, U, u.d_diag.p);
    if (p.n) hanging_diag_kernel<<<blocks(p.n), 128, 0, s>>>(p.n, fs.n_owned_pnodes, pp_, P, p.d_diag.p);
    auto columns = [&](const FoldPlan &f, const Mat &M, const Space &X, const Space &Y) {
      if (f.n_rows) fold_columns_kernel<<<blocks(f.n_rows), 128, 0, s>>>(f.n_rows, f.d_row.p, f.d_ptr.p, f.d_item.p, M, X, Y, con, inhom, rhs);
    };
    // the two column folds of a row space touch the same rhs entries: they run one after the other on the stream
    columns(uu, uu_, U, U);
    columns(up, up_, U, P);
    columns(pu, pu_, P, U);
    columns(pp, pp_, P, P);
    auto rows = [&](const FoldPlan &f, const Mat &M, const Space &X, int with_rhs) {
      if (f.n_mrows) fold_rows_kernel<<<blocks(f.n_mrows), 128, 0, s>>>(f.n_mrows, f.d_mrow.p, f.d_mptr.p, f.d_mslave.p, M, X, con, rhs, with_rhs);
    };
    rows(uu, uu_, U, 1);
    rows(up, up_, U, 0);
    rows(pu, pu_, P, 0);
    rows(pp, pp_, P, 1);
    if (u.n) hanging_rows_kernel<<<blocks(u.n), 128, 0, s>>>(u.n, fs.n_owned_unodes, uu_, up_, U, u.d_diag.p, con, inhom, rhs);
    if (p.n) hanging_rows_kernel<<<blocks(p.n), 128, 0, s>>>(p.n, fs.n_owned_pnodes, pp_, pu_, P, p.d_diag.p, con, inhom, rhs);
    IFEM_KERNEL_CHECK();
    ctx.kernel_launches += (u.n ? 2 : 0) + (p.n ? 2 : 0) + (uu.n_rows > 0) + (up.n_rows > 0) + (pu.n_rows > 0) + (pp.n_rows > 0) +
                           (uu.n_mrows > 0) + (up.n_mrows > 0) + (pu.n_mrows > 0) + (pp.n_mrows > 0);
  }

  void HangingConstraints::condense_scalar(Context &ctx, const FluidSpace &fs, Bcsr &A, double *rhs, const unsigned char *con,
                                           const double *inhom) const
  {
    if (!active || !p.n) return;
    if (!with_pp) throw std::runtime_error("HangingConstraints::condense_scalar needs the fold plan of A_pp");
    cudaStream_t s = ctx.stream;
    const Space P{1, 0, p.d_node.p, p.d_n_masters.p, p.d_master.p};
    const Mat M = view(A);
    hanging_diag_kernel<<<blocks(p.n), 128, 0, s>>>(p.n, fs.n_owned_pnodes, M, P, p.d_diag.p);
    if (pp.n_rows) fold_columns_kernel<<<blocks(pp.n_rows), 128, 0, s>>>(pp.n_rows, pp.d_row.p, pp.d_ptr.p, pp.d_item.p, M, P, P, con, inhom, rhs);
    if (pp.n_mrows) fold_rows_kernel<<<blocks(pp.n_mrows), 128, 0, s>>>(pp.n_mrows, pp.d_mrow.p, pp.d_mptr.p, pp.d_mslave.p, M, P, con, rhs, 1);
    hanging_rows_kernel<<<blocks(p.n), 128, 0, s>>>(p.n, fs.n_owned_pnodes, M, M, P, p.d_diag.p, con, inhom, rhs);
    IFEM_KERNEL_CHECK();
    ctx.kernel_launches += 2 + (pp.n_rows > 0) + (pp.n_mrows > 0);
  }

  void HangingConstraints::distribute_scalar(Context &ctx, double *x) const
  {
    if (!active || !p.n) return;
    const Space P{1, 0, p.d_node.p, p.d_n_masters.p, p.d_master.p};
    hanging_distribute_kernel<<<blocks(p.n), 128, 0, ctx.stream>>>(p.n, P, x);
    IFEM_KERNEL_CHECK();
    ctx.kernel_launches++;
  }

  void HangingConstraints::distribute(Context &ctx, const FluidSpace &fs, double *x) const
  {
    if (!active) return;
    cudaStream_t s = ctx.stream;
    const Space U{fs.dim, 0, u.d_node.p, u.d_n_masters.p, u.d_master.p}, P{1, fs.n_u, p.d_node.p, p.d_n_masters.p, p.d_master.p};
    if (u.n) hanging_distribute_kernel<<<blocks(u.n), 128, 0, s>>>(u.n, U, x);
    if (p.n) hanging_distribute_kernel<<<blocks(p.n), 128, 0, s>>>(p.n, P, x);
    IFEM_KERNEL_CHECK();
    ctx.kernel_launches += (u.n > 0) + (p.n > 0);
  }
} // namespace ifem
