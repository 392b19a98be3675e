#include "partition.h"

#include <algorithm>
#include <cmath>
#include <numeric>
#include <stdexcept>

namespace ifem
{
  namespace
  {
    // Slabs bounded by mesh planes for a locally refined mesh. A hanging node and its masters must end up with the same
    // owner (hanging.h: the rows of a hanging node are folded into the rows of its masters on one rank); with "a node
    // belongs to the lowest rank of its cells" that holds when no cut plane carries a hanging node or a master and no cell
    // straddles a cut. The slab axis is the last one along which enough such planes exist (the band-refined channels of the
    // reference: the z planes of fsi-wall-3D except z = 2 and z = 2.4, the x planes of fsi_leaflet except x = 0.8 and 1.3).
    bool plane_slabs(const Triangulation &tria, int size, int axis, std::vector<int> &rank_of)
    {
      const int dim = tria.dim, nc = tria.n_cells(), vpc = tria.verts_per_cell();
      double lo = 1e300, hi = -1e300;
      for (int v = 0; v < tria.n_vertices(); ++v)
        {
          lo = std::min(lo, tria.vertices[(size_t)v * dim + axis]);
          hi = std::max(hi, tria.vertices[(size_t)v * dim + axis]);
        }
      const double Q = double((1u << 24) - 1) / std::max(hi - lo, 1e-300);
      auto quant = [&](double x) { return (int64_t)std::llround((x - lo) * Q); };
      std::vector<int64_t> bottom(nc), top(nc), planes;
      for (int c = 0; c < nc; ++c)
        {
          double b = 1e300, t = -1e300;
          for (int v = 0; v < vpc; ++v)
            {
              const double x = tria.vertices[(size_t)tria.cells[(size_t)c * vpc + v] * dim + axis];
              b = std::min(b, x);
              t = std::max(t, x);
            }
          bottom[c] = quant(b);
          top[c] = quant(t);
          planes.push_back(top[c]);
        }
      std::sort(planes.begin(), planes.end());
      planes.erase(std::unique(planes.begin(), planes.end()), planes.end());
      planes.pop_back(); // the far side of the domain
      const int np = (int)planes.size();
      if (np < size - 1) return false;
      auto index_of = [&](int64_t q) { return (int)(std::lower_bound(planes.begin(), planes.end(), q) - planes.begin()); };
      std::vector<int> straddle(np + 1, 0), below(np + 1, 0);
      for (int c = 0; c < nc; ++c)
        {
          // planes strictly between bottom and top of the cell
          const int a = (int)(std::upper_bound(planes.begin(), planes.end(), bottom[c]) - planes.begin());
          const int b = index_of(top[c]);
          if (b > a)
            {
              straddle[a]++;
              straddle[b]--;
            }
          below[std::min(b, np)]++; // cells whose top is plane b lie below every plane >= b
        }
      std::vector<char> allowed(np, 1);
      int run = 0, cum = 0;
      std::vector<int> count_below(np, 0);
      for (int k = 0; k < np; ++k)
        {
          run += straddle[k];
          if (run > 0) allowed[k] = 0;
          cum += below[k];
          count_below[k] = cum;
        }
      for (const auto &h : tria.hanging)
        {
          auto forbid = [&](int vertex) {
            const int64_t q = quant(tria.vertices[(size_t)vertex * dim + axis]);
            const int k = index_of(q);
            if (k < np && planes[k] == q) allowed[k] = 0;
          };
          forbid(h.vertex);
          for (int k = 0; k < h.n_masters; ++k) forbid(h.master[k]);
        }
      std::vector<int64_t> cuts;
      int last = -1;
      for (int r = 1; r < size; ++r)
        {
          const double want = (double)nc * r / size;
          int best = -1;
          for (int k = last + 1; k < np; ++k)
            if (allowed[k] && (best < 0 || std::fabs(count_below[k] - want) < std::fabs(count_below[best] - want))) best = k;
          if (best < 0) return false;
          cuts.push_back(planes[best]);
          last = best;
        }
      rank_of.assign(nc, 0);
      for (int c = 0; c < nc; ++c) rank_of[c] = (int)(std::upper_bound(cuts.begin(), cuts.end(), bottom[c]) - cuts.begin());
      // every rank needs cells
      std::vector<int> cnt(size, 0);
      for (int c = 0; c < nc; ++c) cnt[rank_of[c]]++;
      for (int r = 0; r < size; ++r)
        if (!cnt[r]) return false;
      return true;
    }
  } // namespace

  std::vector<int> slab_cell_ranks(const Triangulation &tria, int size)
  {
    const int dim = tria.dim, nc = tria.n_cells(), vpc = tria.verts_per_cell();
    std::vector<int> rank_of(nc, 0);
    if (size <= 1) return rank_of;
    if (!tria.hanging.empty())
      {
        for (int axis = dim - 1; axis >= 0; --axis)
          if (plane_slabs(tria, size, axis, rank_of)) return rank_of;
        throw std::runtime_error("partition: no slab decomposition of the locally refined mesh keeps every hanging node with its masters "
                                 "on one rank (fewer ranks, or a refinement band across one axis)");
      }
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

    // ghost sets of rank r for one node table: layer-1 and layer-2 ghost nodes, each sorted by (owner, global id)
    struct GhostSets
    {
      std::vector<int> owned, g1, g2;
    };

    GhostSets ghost_sets(const NodeTable &nt, int n_cells, const std::vector<int> &owner, const std::vector<char> &layer_of_cell, int rank)
    {
      GhostSets G;
      const int npc = nt.nodes_per_cell;
      std::vector<char> cls(nt.n_nodes, 0); // 1 = in a layer-1 cell, 2 = only in layer-2 cells
      for (int c = 0; c < n_cells; ++c)
        {
          if (!layer_of_cell[c]) continue;
          for (int a = 0; a < npc; ++a)
            {
              char &k = cls[nt.cell_nodes[(size_t)c * npc + a]];
              if (k == 0 || layer_of_cell[c] < k) k = layer_of_cell[c];
            }
        }
      for (int n = 0; n < nt.n_nodes; ++n)
        {
          if (!cls[n]) continue;
          if (owner[n] == rank) G.owned.push_back(n);
          else (cls[n] == 1 ? G.g1 : G.g2).push_back(n);
        }
      auto by_owner = [&](int a, int b) { return owner[a] < owner[b]; };
      std::stable_sort(G.g1.begin(), G.g1.end(), by_owner);
      std::stable_sort(G.g2.begin(), G.g2.end(), by_owner);
      return G;
    }

    // layer of every cell as seen from rank r: 1 = touches a velocity node owned by r, 2 = touches a node of a
    // layer-1 cell, 0 = not local
    std::vector<char> cell_layers(const NodeTable &un, int n_cells, const std::vector<int> &owner_u, int r)
    {
      const int npc = un.nodes_per_cell;
      std::vector<char> layer(n_cells, 0), node1(un.n_nodes, 0);
      for (int c = 0; c < n_cells; ++c)
        for (int a = 0; a < npc; ++a)
          if (owner_u[un.cell_nodes[(size_t)c * npc + a]] == r)
            {
              layer[c] = 1;
              break;
            }
      for (int c = 0; c < n_cells; ++c)
        if (layer[c] == 1)
          for (int a = 0; a < npc; ++a) node1[un.cell_nodes[(size_t)c * npc + a]] = 1;
      for (int c = 0; c < n_cells; ++c)
        if (!layer[c])
          for (int a = 0; a < npc; ++a)
            if (node1[un.cell_nodes[(size_t)c * npc + a]])
              {
                layer[c] = 2;
                break;
              }
      return layer;
    }

    NodePartition partition_nodes(const NodeTable &nt, int n_cells, const std::vector<int> &owner,
                                  const std::vector<std::vector<char>> &layers_of_rank, int rank, int size)
    {
      NodePartition np;
      const GhostSets mine = ghost_sets(nt, n_cells, owner, layers_of_rank[rank], rank);
      np.n_owned = (int)mine.owned.size();
      np.n_layer1 = np.n_owned + (int)mine.g1.size();
      np.local_to_global = mine.owned;
      np.local_to_global.insert(np.local_to_global.end(), mine.g1.begin(), mine.g1.end());
      np.local_to_global.insert(np.local_to_global.end(), mine.g2.begin(), mine.g2.end());
      np.n_local = (int)np.local_to_global.size();
      std::vector<int> g2l(nt.n_nodes, -1);
      for (int l = 0; l < np.n_local; ++l) g2l[np.local_to_global[l]] = l;
      // what every other rank expects from me, per layer (its ghost sets, computed from the same global data);
      // only ranks whose local cells contain a node I own can expect anything
      std::vector<char> wants(size, 0);
      {
        const int npc = nt.nodes_per_cell;
        for (int c = 0; c < n_cells; ++c)
          {
            bool mine_here = false, checked = false;
            for (int s = 0; s < size; ++s)
              {
                if (s == rank || wants[s] || !layers_of_rank[s][c]) continue;
                if (!checked)
                  {
                    for (int a = 0; a < npc && !mine_here; ++a) mine_here = owner[nt.cell_nodes[(size_t)c * npc + a]] == rank;
                    checked = true;
                  }
                if (mine_here) wants[s] = 1;
              }
          }
      }
      std::vector<GhostSets> theirs(size);
      for (int s = 0; s < size; ++s)
        if (s != rank && wants[s]) theirs[s] = ghost_sets(nt, n_cells, owner, layers_of_rank[s], s);
      for (int layer = 1; layer <= 2; ++layer)
        {
          const std::vector<int> &my_ghosts = layer == 1 ? mine.g1 : mine.g2;
          const int base = layer == 1 ? np.n_owned : np.n_layer1;
          for (int s = 0; s < size; ++s)
            {
              if (s == rank) continue;
              // receive: my layer-`layer` ghosts owned by s (contiguous, ascending global id)
              int first = -1, count = 0;
              for (size_t k = 0; k < my_ghosts.size(); ++k)
                if (owner[my_ghosts[k]] == s)
                  {
                    if (first < 0) first = (int)k;
                    ++count;
                  }
              // send: s's layer-`layer` ghosts owned by me, in s's order
              std::vector<int> send;
              for (int n : (layer == 1 ? theirs[s].g1 : theirs[s].g2))
                if (owner[n] == rank) send.push_back(g2l[n]);
              if (!count && send.empty()) continue;
              np.neighbours.push_back(s);
              np.send_local.push_back(send);
              np.recv_offset.push_back(count ? base + first : np.n_local);
              np.recv_count.push_back(count);
            }
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
    // pressure nodes coincide with velocity vertex nodes and get the same owner, so the cell layers defined
    // through the velocity nodes cover the pressure rows as well
    std::vector<std::vector<char>> layers(size);
    for (int s = 0; s < size; ++s) layers[s] = cell_layers(un, nc, owner_u, s);
    for (int c = 0; c < nc; ++c)
      if (layers[rank][c])
        {
          P.local_cells.push_back(c);
          P.cell_layer.push_back(layers[rank][c]);
        }
    P.u = partition_nodes(un, nc, owner_u, layers, rank, size);
    P.p = partition_nodes(pn, nc, owner_p, layers, rank, size);
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
