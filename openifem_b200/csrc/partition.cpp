#include "partition.h"

#include <algorithm>
#include <cmath>
#include <numeric>
#include <stdexcept>

namespace ifem
{
  std::vector<int> slab_cell_ranks(const Triangulation &tria, int size)
  {
    const int dim = tria.dim, nc = tria.n_cells(), vpc = tria.verts_per_cell();
    std::vector<int> rank_of(nc, 0);
    if (size <= 1) return rank_of;
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
    for (int v = 0; v < tria.n_vertices(); ++v)
      for (int d = 0; d < dim; ++d)
        {
          lo[d] = std::min(lo[d], tria.vertices[(size_t)v * dim + d]);
          hi[d] = std::max(hi[d], tria.vertices[(size_t)v * dim + d]);
        }
    std::vector<std::pair<uint64_t, int>> keyed(nc);
    const double Q = double((1u << 21) - 1);
    for (int c = 0; c < nc; ++c)
      {
        uint64_t key = 0;
        for (int d = dim - 1; d >= 0; --d)
          {
            double x = 0;
            for (int v = 0; v < vpc; ++v) x += tria.vertices[(size_t)tria.cells[(size_t)c * vpc + v] * dim + d];
            x /= vpc;
            const double ext = hi[d] - lo[d];
            key = (key << 21) | (uint64_t)std::llround((ext > 0 ? (x - lo[d]) / ext : 0.0) * Q);
          }
        keyed[c] = {key, c};
      }
    std::sort(keyed.begin(), keyed.end());
    for (int pos = 0; pos < nc; ++pos) rank_of[keyed[pos].second] = (int)(((int64_t)pos * size) / nc);
    return rank_of;
  }

  namespace
  {
    std::vector<int> node_owners(const NodeTable &nt, int n_cells, const std::vector<int> &cell_rank, int size)
    {
      std::vector<int> owner(nt.n_nodes, size);
      for (int c = 0; c < n_cells; ++c)
        for (int a = 0; a < nt.nodes_per_cell; ++a)
          {
            int &o = owner[nt.cell_nodes[(size_t)c * nt.nodes_per_cell + a]];
            o = std::min(o, cell_rank[c]);
          }
      return owner;
    }

    NodePartition partition_nodes(const NodeTable &nt, int n_cells, const std::vector<int> &owner,
                                  const std::vector<char> &cell_local_on_me, const std::vector<std::vector<int>> &cell_ranks_touching,
                                  int rank, int size)
    {
      NodePartition np;
      const int npc = nt.nodes_per_cell;
      // local nodes
      std::vector<char> is_local(nt.n_nodes, 0);
      for (int c = 0; c < n_cells; ++c)
        if (cell_local_on_me[c])
          for (int a = 0; a < npc; ++a) is_local[nt.cell_nodes[(size_t)c * npc + a]] = 1;
      std::vector<int> owned, ghosts;
      for (int n = 0; n < nt.n_nodes; ++n)
        if (is_local[n]) (owner[n] == rank ? owned : ghosts).push_back(n);
      std::stable_sort(ghosts.begin(), ghosts.end(), [&](int a, int b) { return owner[a] < owner[b]; });
      np.n_owned = (int)owned.size();
      np.local_to_global = owned;
      np.local_to_global.insert(np.local_to_global.end(), ghosts.begin(), ghosts.end());
      np.n_local = (int)np.local_to_global.size();
      // receive ranges
      std::vector<std::vector<int>> recv_from(size), send_to(size);
      for (size_t k = 0; k < ghosts.size(); ++k) recv_from[owner[ghosts[k]]].push_back(np.n_owned + (int)k);
      // send lists: my owned nodes inside cells that are local on another rank
      std::vector<int> g2l(nt.n_nodes, -1);
      for (int l = 0; l < np.n_local; ++l) g2l[np.local_to_global[l]] = l;
      for (int c = 0; c < n_cells; ++c)
        {
          const auto &rs = cell_ranks_touching[c];
          if (rs.size() < 2) continue;
          for (int s : rs)
            {
              if (s == rank) continue;
              for (int a = 0; a < npc; ++a)
                {
                  const int n = nt.cell_nodes[(size_t)c * npc + a];
                  if (owner[n] == rank) send_to[s].push_back(n);
                }
            }
        }
      for (int s = 0; s < size; ++s)
        {
          auto &v = send_to[s];
          std::sort(v.begin(), v.end());
          v.erase(std::unique(v.begin(), v.end()), v.end());
          if (v.empty() && recv_from[s].empty()) continue;
          np.neighbours.push_back(s);
          std::vector<int> loc(v.size());
          for (size_t k = 0; k < v.size(); ++k) loc[k] = g2l[v[k]];
          np.send_local.push_back(loc);
          np.recv_offset.push_back(recv_from[s].empty() ? np.n_local : recv_from[s].front());
          np.recv_count.push_back((int)recv_from[s].size());
        }
      return np;
    }
  } // namespace

  Partition build_partition(const Triangulation &tria, const NodeTable &un, const NodeTable &pn, int rank, int size)
  {
    Partition P;
    P.rank = rank;
    P.size = size;
    const int nc = tria.n_cells();
    const std::vector<int> cell_rank = slab_cell_ranks(tria, size);
    const std::vector<int> owner_u = node_owners(un, nc, cell_rank, size);
    const std::vector<int> owner_p = node_owners(pn, nc, cell_rank, size);
    // a cell is local on every rank that owns one of its velocity nodes (pressure nodes coincide with
    // velocity vertex nodes and get the same owner, so this covers the pressure rows as well)
    std::vector<std::vector<int>> touching(nc);
    std::vector<char> local_on_me(nc, 0);
    for (int c = 0; c < nc; ++c)
      {
        auto &rs = touching[c];
        for (int a = 0; a < un.nodes_per_cell; ++a)
          {
            const int o = owner_u[un.cell_nodes[(size_t)c * un.nodes_per_cell + a]];
            if (std::find(rs.begin(), rs.end(), o) == rs.end()) rs.push_back(o);
          }
        std::sort(rs.begin(), rs.end());
        if (std::binary_search(rs.begin(), rs.end(), rank)) local_on_me[c] = 1;
      }
    // local cells in slab order
    {
      std::vector<int> order(nc);
      std::iota(order.begin(), order.end(), 0);
      for (int c : order)
        if (local_on_me[c]) P.local_cells.push_back(c);
    }
    P.u = partition_nodes(un, nc, owner_u, local_on_me, touching, rank, size);
    P.p = partition_nodes(pn, nc, owner_p, local_on_me, touching, rank, size);
    return P;
  }

  NodeTable localise(const NodeTable &global, const std::vector<int> &local_cells, const NodePartition &np)
  {
    NodeTable nt;
    nt.p = global.p;
    nt.dim = global.dim;
    nt.nodes_per_cell = global.nodes_per_cell;
    nt.n_nodes = np.n_local;
    std::vector<int> g2l(global.n_nodes, -1);
    for (int l = 0; l < np.n_local; ++l) g2l[np.local_to_global[l]] = l;
    nt.cell_nodes.resize(local_cells.size() * (size_t)nt.nodes_per_cell);
    for (size_t k = 0; k < local_cells.size(); ++k)
      for (int a = 0; a < nt.nodes_per_cell; ++a)
        {
          const int l = g2l[global.cell_nodes[(size_t)local_cells[k] * nt.nodes_per_cell + a]];
          if (l < 0) throw std::runtime_error("localise: node of a local cell is not local");
          nt.cell_nodes[k * nt.nodes_per_cell + a] = l;
        }
    nt.coords.resize((size_t)np.n_local * nt.dim);
    for (int l = 0; l < np.n_local; ++l)
      for (int d = 0; d < nt.dim; ++d) nt.coords[(size_t)l * nt.dim + d] = global.coords[(size_t)np.local_to_global[l] * nt.dim + d];
    return nt;
  }
} // namespace ifem
