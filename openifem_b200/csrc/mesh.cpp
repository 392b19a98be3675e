#include "mesh.h"

#include <map>

#include <algorithm>
#include <unordered_map>
#include <cmath>
#include <numeric>
#include <stdexcept>

#include "fe_tables.h"

namespace ifem
{
  // ---------------------------------------------------------------------------
  // GridGenerator
  // ---------------------------------------------------------------------------
  void GridGenerator::subdivided_hyper_rectangle(Triangulation &tria, const std::vector<unsigned int> &reps,
                                                 const double *p1, const double *p2, bool colorize)
  {
    const int dim = (int)reps.size();
    if (dim != 2 && dim != 3) throw std::runtime_error("subdivided_hyper_rectangle: dim must be 2 or 3");
    tria = Triangulation();
    tria.dim = dim;
    int n[3] = {1, 1, 1}, nv[3] = {1, 1, 1};
    for (int d = 0; d < dim; ++d)
      {
        n[d] = (int)reps[d];
        nv[d] = n[d] + 1;
      }
    const size_t n_vert = (size_t)nv[0] * nv[1] * nv[2];
    tria.vertices.resize(n_vert * dim);
    for (int k = 0; k < nv[2]; ++k)
      for (int j = 0; j < nv[1]; ++j)
        for (int i = 0; i < nv[0]; ++i)
          {
            const size_t v = (size_t)i + (size_t)nv[0] * (j + (size_t)nv[1] * k);
            const int ijk[3] = {i, j, k};
            for (int d = 0; d < dim; ++d)
              tria.vertices[v * dim + d] = p1[d] + ijk[d] * ((p2[d] - p1[d]) / n[d]);
          }
    const int vpc = 1 << dim;
    const size_t nc = (size_t)n[0] * n[1] * n[2];
    tria.cells.resize(nc * vpc);
    tria.material_id.assign(nc, 1);
    for (int k = 0; k < n[2]; ++k)
      for (int j = 0; j < n[1]; ++j)
        for (int i = 0; i < n[0]; ++i)
          {
            const size_t c = (size_t)i + (size_t)n[0] * (j + (size_t)n[1] * k);
            for (int v = 0; v < vpc; ++v)
              {
                const int oi = v & 1, oj = (v >> 1) & 1, ok = (v >> 2) & 1;
                tria.cells[c * vpc + v] = (int)((i + oi) + (size_t)nv[0] * ((j + oj) + (size_t)nv[1] * (k + ok)));
              }
            const int ijk[3] = {i, j, k};
            for (int axis = 0; axis < dim; ++axis)
              for (int side = 0; side < 2; ++side)
                if (ijk[axis] == (side ? n[axis] - 1 : 0))
                  {
                    tria.boundary_faces.push_back((int)c);
                    tria.boundary_faces.push_back(2 * axis + side);
                    tria.boundary_faces.push_back(colorize ? 2 * axis + side : 0);
                  }
          }
  }

  void PolarArcChart::init()
  {
    const double d0[2] = {v[0][0] - c[0], v[0][1] - c[1]}, d1[2] = {v[1][0] - c[0], v[1][1] - c[1]};
    r0 = std::hypot(d0[0], d0[1]);
    r1 = std::hypot(d1[0], d1[1]);
    a0 = std::atan2(d0[1], d0[0]);
    da = std::atan2(d1[1], d1[0]) - a0;
    const double pi = 3.14159265358979323846;
    while (da > pi) da -= 2 * pi; // the shortest way round: the polar manifold is periodic in the angle
    while (da < -pi) da += 2 * pi;
  }

  void PolarArcChart::eval(double s, double t, double *x) const
  {
    const double r = (1 - s) * r0 + s * r1, a = a0 + s * da;
    const double B[2] = {c[0] + r * std::cos(a), c[1] + r * std::sin(a)};
    for (int d = 0; d < 2; ++d)
      {
        const double T = (1 - s) * v[2][d] + s * v[3][d], L = (1 - t) * v[0][d] + t * v[2][d], R = (1 - t) * v[1][d] + t * v[3][d];
        x[d] = (1 - s) * L + s * R + (1 - t) * B[d] + t * T
               - ((1 - s) * (1 - t) * v[0][d] + s * (1 - t) * v[1][d] + (1 - s) * t * v[2][d] + s * t * v[3][d]);
      }
  }

  namespace
  {
    // the 2-D mesh of flow_around_cylinder_2d (utilities.cpp:343-486); boundary ids of GridCreator<2>
    void cylinder_2d(Triangulation &tria, bool compute_in_2d, bool with_charts)
    {
      tria = Triangulation();
      tria.dim = 2;
      const double left = compute_in_2d ? 0.0 : -0.3;
      const int nx = compute_in_2d ? 22 : 25, ny = 4;
      std::vector<double> xs(nx + 1), ys(ny + 1);
      for (int i = 0; i <= nx; ++i) xs[i] = left + i * ((2.2 - left) / nx);
      for (int j = 0; j <= ny; ++j) ys[j] = j * (0.41 / ny);
      xs[nx] = 2.2;
      ys[ny] = 0.41;
      std::vector<int> vid((size_t)(nx + 1) * (ny + 1), -1);
      auto vertex = [&](int i, int j) {
        int &id = vid[(size_t)j * (nx + 1) + i];
        if (id < 0)
          {
            id = tria.n_vertices();
            tria.vertices.push_back(xs[i]);
            tria.vertices.push_back(ys[j]);
          }
        return id;
      };
      // bulk cells whose centre is closer than 0.15 to (0.2, 0.2) are removed (the 2 x 2 block around the hole)
      int i0 = nx, j0 = ny, n_removed = 0;
      for (int j = 0; j < ny; ++j)
        for (int i = 0; i < nx; ++i)
          {
            const double cx = 0.5 * (xs[i] + xs[i + 1]) - 0.2, cy = 0.5 * (ys[j] + ys[j + 1]) - 0.2;
            if (std::hypot(cx, cy) < 0.15)
              {
                i0 = std::min(i0, i);
                j0 = std::min(j0, j);
                ++n_removed;
                continue;
              }
            const int cv[4] = {vertex(i, j), vertex(i + 1, j), vertex(i, j + 1), vertex(i + 1, j + 1)};
            tria.cells.insert(tria.cells.end(), cv, cv + 4);
          }
      if (n_removed != 4) throw std::runtime_error("flow_around_cylinder: unexpected hole in the bulk mesh");
      const int n_bulk = tria.n_cells();
      // hyper_cube_with_cylindrical_hole(0.05, 0.41 / 4): 8 cells; its outer ring is merged onto the bulk lattice
      // (merge_triangulations keeps the bulk coordinates), ring vertex k sits at the angle k * 45 degrees
      const int ring[8][2] = {{2, 1}, {2, 2}, {1, 2}, {0, 2}, {0, 1}, {0, 0}, {1, 0}, {2, 0}};
      int outer[8], inner[8];
      for (int k = 0; k < 8; ++k) outer[k] = vertex(i0 + ring[k][0], j0 + ring[k][1]);
      const double pi = 3.14159265358979323846;
      for (int k = 0; k < 8; ++k)
        {
          inner[k] = tria.n_vertices();
          // the circle's vertices are re-centred at (0.2, 0.2) (utilities.cpp:452-480)
          tria.vertices.push_back(0.2 + 0.05 * std::cos(2 * pi * k / 8));
          tria.vertices.push_back(0.2 + 0.05 * std::sin(2 * pi * k / 8));
        }
      if (with_charts)
        {
          tria.chart_of_cell.assign(n_bulk, -1);
          tria.chart_box.assign((size_t)n_bulk * 4, 0.0);
        }
      for (int k = 0; k < 8; ++k)
        {
          const int k1 = (k + 1) % 8;
          int cv[4] = {inner[k1], inner[k], outer[k1], outer[k]}; // arc = edge v0 -> v1 (local face 2)
          auto X = [&](int v, int d) { return tria.vertices[(size_t)cv[v] * 2 + d]; };
          const double area = (X(1, 0) - X(0, 0)) * (X(2, 1) - X(0, 1)) - (X(2, 0) - X(0, 0)) * (X(1, 1) - X(0, 1));
          if (area < 0)
            {
              std::swap(cv[0], cv[1]);
              std::swap(cv[2], cv[3]);
            }
          tria.cells.insert(tria.cells.end(), cv, cv + 4);
          const int c = tria.n_cells() - 1;
          const int bf[3] = {c, 2, 4};
          tria.boundary_faces.insert(tria.boundary_faces.end(), bf, bf + 3);
          if (with_charts)
            {
              PolarArcChart ch;
              for (int v = 0; v < 4; ++v)
                for (int d = 0; d < 2; ++d) ch.v[v][d] = tria.vertices[(size_t)cv[v] * 2 + d];
              ch.c[0] = ch.c[1] = 0.2;
              ch.init();
              tria.charts.push_back(ch);
              tria.chart_of_cell.push_back(k);
              const double box[4] = {0.0, 0.0, 1.0, 1.0};
              tria.chart_box.insert(tria.chart_box.end(), box, box + 4);
            }
        }
      tria.material_id.assign(tria.n_cells(), 1);
      // boundary faces of the bulk: edges that belong to one cell only
      const int fv[4][2] = {{0, 2}, {1, 3}, {0, 1}, {2, 3}};
      std::map<std::pair<int, int>, std::pair<int, int>> edges; // edge -> (owner count, cell * 4 + face)
      for (int c = 0; c < tria.n_cells(); ++c)
        for (int f = 0; f < 4; ++f)
          {
            int a = tria.cells[(size_t)c * 4 + fv[f][0]], b = tria.cells[(size_t)c * 4 + fv[f][1]];
            if (a > b) std::swap(a, b);
            auto &e = edges[{a, b}];
            e.first++;
            e.second = c * 4 + f;
          }
      for (const auto &kv : edges)
        {
          if (kv.second.first != 1) continue;
          const int c = kv.second.second / 4, f = kv.second.second % 4;
          if (c >= n_bulk) continue; // the arcs were added with their cells
          const double mx = 0.5 * (tria.vertices[(size_t)kv.first.first * 2] + tria.vertices[(size_t)kv.first.second * 2]);
          const double my = 0.5 * (tria.vertices[(size_t)kv.first.first * 2 + 1] + tria.vertices[(size_t)kv.first.second * 2 + 1]);
          int id = 4;
          if (std::fabs(mx - 2.2) < 1e-12) id = 1;
          else if (std::fabs(mx - left) < 1e-12) id = 0;
          else if (std::fabs(my - 0.41) < 1e-12) id = 3;
          else if (std::fabs(my) < 1e-12) id = 2;
          const int bf[3] = {c, f, id};
          tria.boundary_faces.insert(tria.boundary_faces.end(), bf, bf + 3);
        }
    }
  } // namespace

  void GridCreator::flow_around_cylinder(Triangulation &tria, int dim)
  {
    if (dim == 2)
      {
        cylinder_2d(tria, true, true);
        return;
      }
    if (dim != 3) throw std::runtime_error("flow_around_cylinder: dim must be 2 or 3");
    // GridGenerator::extrude_triangulation(tria_2d, 9, 0.41, tria): 9 slices = 8 layers; manifolds are not copied, so
    // the 3-D mesh refines flat
    Triangulation t2;
    cylinder_2d(t2, false, false);
    const int nv2 = t2.n_vertices(), nc2 = t2.n_cells(), layers = 8;
    tria = Triangulation();
    tria.dim = 3;
    tria.vertices.reserve((size_t)nv2 * (layers + 1) * 3);
    for (int k = 0; k <= layers; ++k)
      for (int v = 0; v < nv2; ++v)
        {
          tria.vertices.push_back(t2.vertices[(size_t)v * 2]);
          tria.vertices.push_back(t2.vertices[(size_t)v * 2 + 1]);
          tria.vertices.push_back(k == layers ? 0.41 : k * (0.41 / layers));
        }
    for (int k = 0; k < layers; ++k)
      for (int c = 0; c < nc2; ++c)
        {
          for (int up = 0; up < 2; ++up)
            for (int v = 0; v < 4; ++v) tria.cells.push_back((k + up) * nv2 + t2.cells[(size_t)c * 4 + v]);
          const int c3 = k * nc2 + c;
          if (k == 0)
            {
              const int bf[3] = {c3, 4, 4};
              tria.boundary_faces.insert(tria.boundary_faces.end(), bf, bf + 3);
            }
          if (k == layers - 1)
            {
              const int bf[3] = {c3, 5, 5};
              tria.boundary_faces.insert(tria.boundary_faces.end(), bf, bf + 3);
            }
        }
    for (int k = 0; k < layers; ++k)
      for (int f = 0; f < t2.n_boundary_faces(); ++f)
        {
          const int id2 = t2.boundary_faces[3 * f + 2];
          const int bf[3] = {k * nc2 + t2.boundary_faces[3 * f], t2.boundary_faces[3 * f + 1], id2 == 4 ? 6 : id2};
          tria.boundary_faces.insert(tria.boundary_faces.end(), bf, bf + 3);
        }
    tria.material_id.assign(tria.n_cells(), 1);
  }

  void GridGenerator::hyper_cube(Triangulation &tria, int dim, double left, double right, bool colorize)
  {
    const double a[3] = {left, left, left}, b[3] = {right, right, right};
    subdivided_hyper_rectangle(tria, std::vector<unsigned int>(dim, 1u), a, b, colorize);
  }

  // ---------------------------------------------------------------------------
  // Node numbering
  // ---------------------------------------------------------------------------
  namespace
  {
    struct Key
    {
      int v[4];
      bool operator==(const Key &o) const { return v[0] == o.v[0] && v[1] == o.v[1] && v[2] == o.v[2] && v[3] == o.v[3]; }
    };
    inline uint64_t hash_key(const Key &k)
    {
      uint64_t h = 0x9E3779B97F4A7C15ull;
      for (int i = 0; i < 4; ++i)
        {
          h ^= (uint64_t)(uint32_t)k.v[i] + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2);
          h *= 0xff51afd7ed558ccdull;
          h ^= h >> 33;
        }
      return h;
    }
    struct FlatMap
    {
      std::vector<int> slot; // index into keys, -1 empty
      std::vector<Key> keys;
      uint64_t mask;
      explicit FlatMap(size_t expected)
      {
        size_t cap = 16;
        while (cap < 2 * expected) cap <<= 1;
        slot.assign(cap, -1);
        mask = cap - 1;
        keys.reserve(expected);
      }
      int find_or_insert(const Key &k)
      {
        uint64_t h = hash_key(k) & mask;
        while (true)
          {
            const int s = slot[h];
            if (s < 0)
              {
                slot[h] = (int)keys.size();
                keys.push_back(k);
                return (int)keys.size() - 1;
              }
            if (keys[s] == k) return s;
            h = (h + 1) & mask;
          }
      }
    };

    // renumber nodes in lexicographic (z,y,x) order of quantised coordinates
    void spatial_renumber(NodeTable &nt)
    {
      const int dim = nt.dim, n = nt.n_nodes;
      double lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0};
      for (int d = 0; d < dim; ++d) lo[d] = hi[d] = nt.coords[d];
      for (int i = 0; i < n; ++i)
        for (int d = 0; d < dim; ++d)
          {
            lo[d] = std::min(lo[d], nt.coords[(size_t)i * dim + d]);
            hi[d] = std::max(hi[d], nt.coords[(size_t)i * dim + d]);
          }
      std::vector<std::pair<uint64_t, int>> keyed(n);
      const double Q = double((1u << 21) - 1);
#pragma omp parallel for schedule(static)
      for (int i = 0; i < n; ++i)
        {
          uint64_t key = 0;
          for (int d = dim - 1; d >= 0; --d)
            {
              const double ext = hi[d] - lo[d];
              const double t = ext > 0 ? (nt.coords[(size_t)i * dim + d] - lo[d]) / ext : 0.0;
              const uint64_t q = (uint64_t)std::llround(t * Q);
              key = (key << 21) | q;
            }
          keyed[i] = {key, i};
        }
      std::sort(keyed.begin(), keyed.end());
      std::vector<int> new_id(n);
      std::vector<double> c2((size_t)n * dim);
      for (int r = 0; r < n; ++r)
        {
          const int old = keyed[r].second;
          new_id[old] = r;
          for (int d = 0; d < dim; ++d) c2[(size_t)r * dim + d] = nt.coords[(size_t)old * dim + d];
        }
      nt.coords.swap(c2);
#pragma omp parallel for schedule(static)
      for (size_t k = 0; k < nt.cell_nodes.size(); ++k) nt.cell_nodes[k] = new_id[nt.cell_nodes[k]];
    }
  } // namespace

  NodeTable build_node_table(const Triangulation &tria, int p)
  {
    if (p != 1 && p != 2) throw std::runtime_error("build_node_table: only FE_Q(1) and FE_Q(2) are supported");
    const int dim = tria.dim, vpc = tria.verts_per_cell(), nc = tria.n_cells();
    NodeTable nt;
    nt.p = p;
    nt.dim = dim;
    FEQ fe(dim, p);
    nt.nodes_per_cell = fe.n;
    nt.cell_nodes.resize((size_t)nc * fe.n);
    if (p == 1)
      {
        nt.n_nodes = tria.n_vertices();
        nt.coords = tria.vertices;
        for (size_t k = 0; k < tria.cells.size(); ++k) nt.cell_nodes[k] = tria.cells[k];
        spatial_renumber(nt);
        return nt;
      }
    // p == 2: a node is identified by the set of cell corners it averages
    const int nv = tria.n_vertices();
    FlatMap map((size_t)nc * (dim == 3 ? 8 : 4) + 16);
    // vertices keep ids 0..nv-1, shared entities follow, cell-interior nodes last
    std::vector<int> corner_sets((size_t)fe.n * 8, -1), corner_cnt(fe.n, 0);
    for (int a = 0; a < fe.n; ++a)
      {
        int cnt = 0;
        for (int v = 0; v < vpc; ++v)
          {
            bool ok = true;
            for (int d = 0; d < dim; ++d)
              {
                const int l = fe.lattice[a][d], bit = (v >> d) & 1;
                if (l == 0 && bit != 0) ok = false;
                if (l == 2 && bit != 1) ok = false;
              }
            if (ok) corner_sets[(size_t)a * 8 + cnt++] = v;
          }
        corner_cnt[a] = cnt;
      }
    int n_shared = 0;
    std::vector<int> tmp_ids((size_t)nc * fe.n);
    for (int c = 0; c < nc; ++c)
      {
        const int *cv = &tria.cells[(size_t)c * vpc];
        for (int a = 0; a < fe.n; ++a)
          {
            const int cnt = corner_cnt[a];
            int id;
            if (cnt == 1)
              id = cv[corner_sets[(size_t)a * 8]];
            else if (cnt == vpc)
              id = -1 - c; // cell-interior node, resolved below
            else
              {
                Key k{{-1, -1, -1, -1}};
                for (int i = 0; i < cnt; ++i) k.v[i] = cv[corner_sets[(size_t)a * 8 + i]];
                std::sort(k.v, k.v + cnt);
                id = nv + map.find_or_insert(k);
              }
            tmp_ids[(size_t)c * fe.n + a] = id;
          }
      }
    n_shared = (int)map.keys.size();
    nt.n_nodes = nv + n_shared + nc;
    nt.coords.assign((size_t)nt.n_nodes * dim, 0.0);
    for (int c = 0; c < nc; ++c)
      {
        const int *cv = &tria.cells[(size_t)c * vpc];
        for (int a = 0; a < fe.n; ++a)
          {
            int id = tmp_ids[(size_t)c * fe.n + a];
            if (id < 0) id = nv + n_shared + c;
            nt.cell_nodes[(size_t)c * fe.n + a] = id;
            const int cnt = corner_cnt[a];
            double x[3] = {0, 0, 0};
            for (int i = 0; i < cnt; ++i)
              for (int d = 0; d < dim; ++d) x[d] += tria.vertices[(size_t)cv[corner_sets[(size_t)a * 8 + i]] * dim + d];
            for (int d = 0; d < dim; ++d) nt.coords[(size_t)id * dim + d] = x[d] / cnt;
          }
      }
    spatial_renumber(nt);
    return nt;
  }

  std::vector<int> face_local_nodes(int dim, int p, int face_no)
  {
    FEQ fe(dim, p);
    const int axis = face_no / 2, side = face_no % 2;
    std::vector<int> out;
    for (int a = 0; a < fe.n; ++a)
      if (fe.lattice[a][axis] == (side ? p : 0)) out.push_back(a);
    return out;
  }

  void Triangulation::refine_global(int times)
  {
    for (int t = 0; t < times; ++t)
      {
        const NodeTable nt = build_node_table(*this, 2);
        const int vpc = verts_per_cell(), nc = n_cells();
        FEQ fe(dim, 2);
        std::vector<int> new_cells((size_t)nc * vpc * vpc);
        std::vector<int> new_mat((size_t)nc * vpc);
        for (int c = 0; c < nc; ++c)
          for (int child = 0; child < vpc; ++child)
            {
              new_mat[(size_t)c * vpc + child] = material_id[c];
              for (int v = 0; v < vpc; ++v)
                {
                  int a = 0, stride = 1;
                  for (int d = 0; d < dim; ++d)
                    {
                      a += (((child >> d) & 1) + ((v >> d) & 1)) * stride;
                      stride *= 3;
                    }
                  new_cells[((size_t)c * vpc + child) * vpc + v] = nt.cell_nodes[(size_t)c * fe.n + a];
                }
            }
        std::vector<int> new_bf;
        for (int f = 0; f < n_boundary_faces(); ++f)
          {
            const int c = boundary_faces[3 * f], face = boundary_faces[3 * f + 1], id = boundary_faces[3 * f + 2];
            const int axis = face / 2, side = face % 2;
            for (int child = 0; child < vpc; ++child)
              if (((child >> axis) & 1) == side)
                {
                  new_bf.push_back(c * vpc + child);
                  new_bf.push_back(face);
                  new_bf.push_back(id);
                }
          }
        std::vector<double> new_vertices = nt.coords;
        if (dim == 2 && !chart_of_cell.empty())
          {
            // curved placement: lattice point (i, j) of a chart cell lies at chart(s0 + i/2 ds, t0 + j/2 dt)
            std::vector<int> new_chart((size_t)nc * vpc);
            std::vector<double> new_box((size_t)nc * vpc * 4);
            for (int c = 0; c < nc; ++c)
              {
                const int ch = chart_of_cell[c];
                const double *b = &chart_box[(size_t)c * 4];
                const double ds = 0.5 * (b[2] - b[0]), dt = 0.5 * (b[3] - b[1]);
                if (ch >= 0)
                  for (int j = 0; j < 3; ++j)
                    for (int i = 0; i < 3; ++i)
                      {
                        if (i != 1 && j != 1) continue; // corners exist already
                        charts[ch].eval(b[0] + i * ds, b[1] + j * dt, &new_vertices[(size_t)nt.cell_nodes[(size_t)c * fe.n + i + 3 * j] * 2]);
                      }
                for (int child = 0; child < vpc; ++child)
                  {
                    const int cx = child & 1, cy = child >> 1;
                    new_chart[(size_t)c * vpc + child] = ch;
                    double *nb = &new_box[((size_t)c * vpc + child) * 4];
                    nb[0] = b[0] + cx * ds;
                    nb[1] = b[1] + cy * dt;
                    nb[2] = b[0] + (cx + 1) * ds;
                    nb[3] = b[1] + (cy + 1) * dt;
                  }
              }
            chart_of_cell.swap(new_chart);
            chart_box.swap(new_box);
          }
        vertices.swap(new_vertices);
        cells.swap(new_cells);
        boundary_faces.swap(new_bf);
        {
          // every cell becomes the parent of a family
          if (cell_level.empty()) cell_level.assign((size_t)nc, 0);
          if (cell_family.empty())
            {
              cell_family.assign((size_t)nc, -1);
              cell_child_no.assign((size_t)nc, 0);
            }
          std::vector<int> lv((size_t)nc * vpc), fam((size_t)nc * vpc), cno((size_t)nc * vpc);
          for (int c = 0; c < nc; ++c)
            {
              Family f;
              f.parent_family = cell_family[c];
              f.parent_child_no = cell_child_no[c];
              f.parent_level = cell_level[c];
              f.material = material_id[c];
              families.push_back(f);
              for (int child = 0; child < vpc; ++child)
                {
                  lv[(size_t)c * vpc + child] = cell_level[c] + 1;
                  fam[(size_t)c * vpc + child] = (int)families.size() - 1;
                  cno[(size_t)c * vpc + child] = child;
                }
            }
          cell_level.swap(lv);
          cell_family.swap(fam);
          cell_child_no.swap(cno);
        }
        material_id.swap(new_mat);
        if (!hanging.empty()) find_hanging_vertices();
      }
  }

  void Triangulation::execute_refinement(const std::vector<unsigned char> &flags)
  {
    const int vpc = verts_per_cell(), nc = n_cells();
    if ((int)flags.size() != nc) throw std::runtime_error("execute_refinement: one flag per active cell expected");
    if (!chart_of_cell.empty()) throw std::runtime_error("execute_refinement: local refinement of meshes with curved charts is not implemented");
    bool any = false;
    for (unsigned char f : flags) any = any || f;
    if (!any) return;
    const NodeTable nt = build_node_table(*this, 2);
    FEQ fe(dim, 2);
    // Q2 nodes that become vertices: the corners of every cell and all nodes of the flagged cells
    std::vector<int> new_id((size_t)nt.n_nodes, -1);
    std::vector<char> used((size_t)nt.n_nodes, 0);
    auto is_corner = [&](int a) {
      for (int d = 0; d < dim; ++d)
        if (fe.lattice[a][d] == 1) return false;
      return true;
    };
    for (int c = 0; c < nc; ++c)
      for (int a = 0; a < fe.n; ++a)
        if (flags[c] || is_corner(a)) used[nt.cell_nodes[(size_t)c * fe.n + a]] = 1;
    int nv_new = 0;
    for (int i = 0; i < nt.n_nodes; ++i)
      if (used[i]) new_id[i] = nv_new++;
    std::vector<double> new_vertices((size_t)nv_new * dim);
    for (int i = 0; i < nt.n_nodes; ++i)
      if (used[i])
        for (int d = 0; d < dim; ++d) new_vertices[(size_t)new_id[i] * dim + d] = nt.coords[(size_t)i * dim + d];
    std::vector<int> new_cells, new_mat, new_level, new_fam, new_cno, first_child((size_t)nc, -1);
    if (cell_level.empty()) cell_level.assign((size_t)nc, 0);
    if (cell_family.empty())
      {
        cell_family.assign((size_t)nc, -1);
        cell_child_no.assign((size_t)nc, 0);
      }
    for (int c = 0; c < nc; ++c)
      {
        first_child[c] = (int)new_mat.size();
        if (!flags[c])
          {
            for (int v = 0; v < vpc; ++v)
              {
                int a = 0, stride = 1;
                for (int d = 0; d < dim; ++d)
                  {
                    a += 2 * ((v >> d) & 1) * stride;
                    stride *= 3;
                  }
                new_cells.push_back(new_id[nt.cell_nodes[(size_t)c * fe.n + a]]);
              }
            new_mat.push_back(material_id[c]);
            new_level.push_back(cell_level[c]);
            new_fam.push_back(cell_family[c]);
            new_cno.push_back(cell_child_no[c]);
            continue;
          }
        {
          Family f;
          f.parent_family = cell_family[c];
          f.parent_child_no = cell_child_no[c];
          f.parent_level = cell_level[c];
          f.material = material_id[c];
          families.push_back(f);
        }
        for (int child = 0; child < vpc; ++child)
          {
            for (int v = 0; v < vpc; ++v)
              {
                int a = 0, stride = 1;
                for (int d = 0; d < dim; ++d)
                  {
                    a += (((child >> d) & 1) + ((v >> d) & 1)) * stride;
                    stride *= 3;
                  }
                new_cells.push_back(new_id[nt.cell_nodes[(size_t)c * fe.n + a]]);
              }
            new_mat.push_back(material_id[c]);
            new_level.push_back(cell_level[c] + 1);
            new_fam.push_back((int)families.size() - 1);
            new_cno.push_back(child);
          }
      }
    std::vector<int> new_bf;
    for (int f = 0; f < n_boundary_faces(); ++f)
      {
        const int c = boundary_faces[3 * f], face = boundary_faces[3 * f + 1], id = boundary_faces[3 * f + 2];
        const int axis = face / 2, side = face % 2;
        if (!flags[c])
          {
            new_bf.insert(new_bf.end(), {first_child[c], face, id});
            continue;
          }
        for (int child = 0; child < vpc; ++child)
          if (((child >> axis) & 1) == side) new_bf.insert(new_bf.end(), {first_child[c] + child, face, id});
      }
    vertices.swap(new_vertices);
    cells.swap(new_cells);
    material_id.swap(new_mat);
    cell_level.swap(new_level);
    cell_family.swap(new_fam);
    cell_child_no.swap(new_cno);
    boundary_faces.swap(new_bf);
    find_hanging_vertices();
  }

  int Triangulation::n_levels() const
  {
    int m = 0;
    for (int l : cell_level) m = std::max(m, l);
    return m + 1;
  }

  void Triangulation::execute_coarsening_and_refinement(const std::vector<unsigned char> &refine_flags,
                                                         const std::vector<unsigned char> &coarsen_flags, TransferPlan *plan)
  {
    const int vpc = verts_per_cell(), nc = n_cells(), nv = n_vertices();
    if ((int)refine_flags.size() != nc || (!coarsen_flags.empty() && (int)coarsen_flags.size() != nc))
      throw std::runtime_error("execute_coarsening_and_refinement: one flag per active cell expected");
    if (!chart_of_cell.empty()) throw std::runtime_error("execute_coarsening_and_refinement: meshes with curved charts are not supported");
    if (cell_level.empty()) cell_level.assign((size_t)nc, 0);
    if (cell_family.empty())
      {
        cell_family.assign((size_t)nc, -1);
        cell_child_no.assign((size_t)nc, 0);
      }
    // ---- target level of every cell ----
    std::vector<int> desired(cell_level);
    for (int c = 0; c < nc; ++c)
      if (refine_flags[c]) desired[c] = cell_level[c] + 1;
      else if (!coarsen_flags.empty() && coarsen_flags[c] && cell_family[c] >= 0) desired[c] = cell_level[c] - 1;
    // members of every family among the active cells
    std::vector<std::vector<int>> members(families.size());
    for (int c = 0; c < nc; ++c)
      if (cell_family[c] >= 0) members[cell_family[c]].push_back(c);
    auto fix_families = [&]() {
      bool changed = false;
      for (size_t f = 0; f < members.size(); ++f)
        {
          const std::vector<int> &m = members[f];
          if (m.empty()) continue;
          bool all = (int)m.size() == vpc;
          for (int c : m) all = all && desired[c] == cell_level[c] - 1;
          if (all) continue;
          for (int c : m)
            if (desired[c] < cell_level[c])
              {
                desired[c] = cell_level[c];
                changed = true;
              }
        }
      return changed;
    };
    fix_families();
    // ---- 2:1 balance over shared vertices ----
    std::vector<int> vmax((size_t)nv);
    for (int sweep = 0; sweep < 64; ++sweep)
      {
        std::fill(vmax.begin(), vmax.end(), -1);
        for (int c = 0; c < nc; ++c)
          for (int v = 0; v < vpc; ++v) vmax[cells[(size_t)c * vpc + v]] = std::max(vmax[cells[(size_t)c * vpc + v]], desired[c]);
        // (a fine cell next to a coarser one always shares one of the coarse cell's corners with it, so the corners carry every
        // demand across a refinement interface; hanging vertices need no special treatment)
        bool changed = false;
        for (int c = 0; c < nc; ++c)
          {
            int need = -1;
            for (int v = 0; v < vpc; ++v) need = std::max(need, vmax[cells[(size_t)c * vpc + v]] - 1);
            need = std::min(need, cell_level[c] + 1);
            if (desired[c] < need)
              {
                desired[c] = need;
                changed = true;
              }
          }
        changed = fix_families() || changed;
        if (!changed) break;
      }
    // ---- new cells ----
    double lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0};
    for (int d = 0; d < dim; ++d) lo[d] = hi[d] = vertices[d];
    for (int i = 0; i < nv; ++i)
      for (int d = 0; d < dim; ++d)
        {
          lo[d] = std::min(lo[d], vertices[(size_t)i * dim + d]);
          hi[d] = std::max(hi[d], vertices[(size_t)i * dim + d]);
        }
    auto key_of = [&](const double *x) {
      uint64_t key = 0;
      for (int d = dim - 1; d >= 0; --d)
        {
          const double ext = hi[d] - lo[d];
          key = (key << 21) | (uint64_t)std::llround((ext > 0 ? (x[d] - lo[d]) / ext : 0.0) * double((1u << 21) - 2));
        }
      return key;
    };
    std::unordered_map<uint64_t, int> at;
    at.reserve((size_t)nv * 2);
    for (int i = 0; i < nv; ++i) at.emplace(key_of(&vertices[(size_t)i * dim]), i);
    std::vector<double> verts(vertices);
    // transfer weights of vertices created here: (old vertex, weight) lists, old vertices are the identity
    std::vector<std::vector<std::pair<int, double>>> created;
    auto lattice_vertex = [&](int c, const int *ijk) {
      // point (i, j, k) / 2 of cell c under its Q1 map = mean of the corners whose bits agree where the index is 0 or 2
      int ids[8], n = 0;
      for (int v = 0; v < vpc; ++v)
        {
          bool ok = true;
          for (int d = 0; d < dim; ++d)
            {
              const int bit = (v >> d) & 1;
              if ((ijk[d] == 0 && bit) || (ijk[d] == 2 && !bit)) ok = false;
            }
          if (ok) ids[n++] = cells[(size_t)c * vpc + v];
        }
      if (n == 1) return ids[0];
      double x[3] = {0, 0, 0};
      for (int k = 0; k < n; ++k)
        for (int d = 0; d < dim; ++d) x[d] += vertices[(size_t)ids[k] * dim + d] / n;
      const uint64_t key = key_of(x);
      auto it = at.find(key);
      if (it != at.end()) return it->second;
      const int id = (int)(verts.size() / dim);
      for (int d = 0; d < dim; ++d) verts.push_back(x[d]);
      at.emplace(key, id);
      std::vector<std::pair<int, double>> w;
      for (int k = 0; k < n; ++k) w.emplace_back(ids[k], 1.0 / n);
      created.push_back(w);
      return id;
    };
    std::vector<int> new_cells, new_mat, new_level, new_fam, new_cno;
    std::vector<int> new_index_of_old((size_t)nc, -1), n_new_of_old((size_t)nc, 0);
    std::vector<char> family_done(families.size(), 0);
    for (int c = 0; c < nc; ++c)
      {
        if (desired[c] == cell_level[c] - 1)
          {
            const int f = cell_family[c];
            if (family_done[f]) continue;
            family_done[f] = 1;
            // the parent: corner k is corner k of child k
            std::vector<int> child_cell(vpc, -1);
            for (int m : members[f]) child_cell[cell_child_no[m]] = m;
            new_index_of_old[c] = (int)new_mat.size();
            for (int k = 0; k < vpc; ++k) new_cells.push_back(cells[(size_t)child_cell[k] * vpc + k]);
            new_mat.push_back(families[f].material);
            new_level.push_back(families[f].parent_level);
            new_fam.push_back(families[f].parent_family);
            new_cno.push_back(families[f].parent_child_no);
            for (int m : members[f]) new_index_of_old[m] = new_index_of_old[c]; // children -> the parent (boundary faces below)
            continue;
          }
        new_index_of_old[c] = (int)new_mat.size();
        if (desired[c] == cell_level[c])
          {
            for (int v = 0; v < vpc; ++v) new_cells.push_back(cells[(size_t)c * vpc + v]);
            new_mat.push_back(material_id[c]);
            new_level.push_back(cell_level[c]);
            new_fam.push_back(cell_family[c]);
            new_cno.push_back(cell_child_no[c]);
            n_new_of_old[c] = 1;
            continue;
          }
        Family f;
        f.parent_family = cell_family[c];
        f.parent_child_no = cell_child_no[c];
        f.parent_level = cell_level[c];
        f.material = material_id[c];
        families.push_back(f);
        for (int child = 0; child < vpc; ++child)
          {
            for (int v = 0; v < vpc; ++v)
              {
                int ijk[3] = {0, 0, 0};
                for (int d = 0; d < dim; ++d) ijk[d] = ((child >> d) & 1) + ((v >> d) & 1);
                new_cells.push_back(lattice_vertex(c, ijk));
              }
            new_mat.push_back(material_id[c]);
            new_level.push_back(cell_level[c] + 1);
            new_fam.push_back((int)families.size() - 1);
            new_cno.push_back(child);
          }
        n_new_of_old[c] = vpc;
      }
    // ---- boundary faces ----
    std::vector<int> new_bf;
    {
      std::vector<char> seen; // (new cell, face) already emitted (a coarsened family meets it once per child on that side)
      seen.assign(new_mat.size() * 2 * (size_t)dim, 0);
      for (int f = 0; f < n_boundary_faces(); ++f)
        {
          const int c = boundary_faces[3 * f], face = boundary_faces[3 * f + 1], id = boundary_faces[3 * f + 2];
          const int axis = face / 2, side = face % 2;
          if (desired[c] == cell_level[c] + 1)
            {
              for (int child = 0; child < vpc; ++child)
                if (((child >> axis) & 1) == side) new_bf.insert(new_bf.end(), {new_index_of_old[c] + child, face, id});
              continue;
            }
          const int nc2 = new_index_of_old[c];
          if (desired[c] == cell_level[c] - 1 && ((cell_child_no[c] >> axis) & 1) != side) continue; // an inner face of the parent
          char &s = seen[(size_t)nc2 * 2 * dim + face];
          if (s) continue;
          s = 1;
          new_bf.insert(new_bf.end(), {nc2, face, id});
        }
    }
    // ---- drop unused vertices ----
    const int nv_all = (int)(verts.size() / dim);
    std::vector<int> new_id((size_t)nv_all, -1);
    for (int v : new_cells) new_id[v] = 0;
    int nv_new = 0;
    for (int i = 0; i < nv_all; ++i)
      if (new_id[i] == 0) new_id[i] = nv_new++;
    std::vector<double> new_vertices((size_t)nv_new * dim);
    for (int i = 0; i < nv_all; ++i)
      if (new_id[i] >= 0)
        for (int d = 0; d < dim; ++d) new_vertices[(size_t)new_id[i] * dim + d] = verts[(size_t)i * dim + d];
    for (int &v : new_cells) v = new_id[v];
    if (plan)
      {
        plan->ptr.assign((size_t)nv_new + 1, 0);
        plan->old_vertex.clear();
        plan->weight.clear();
        std::vector<int> old_of_new((size_t)nv_new, -1);
        for (int i = 0; i < nv_all; ++i)
          if (new_id[i] >= 0) old_of_new[new_id[i]] = i;
        for (int k = 0; k < nv_new; ++k)
          {
            const int i = old_of_new[k];
            if (i < nv)
              {
                plan->old_vertex.push_back(i);
                plan->weight.push_back(1.0);
              }
            else
              for (const auto &pw : created[(size_t)i - nv])
                {
                  plan->old_vertex.push_back(pw.first);
                  plan->weight.push_back(pw.second);
                }
            plan->ptr[(size_t)k + 1] = (int64_t)plan->old_vertex.size();
          }
      }
    vertices.swap(new_vertices);
    cells.swap(new_cells);
    material_id.swap(new_mat);
    cell_level.swap(new_level);
    cell_family.swap(new_fam);
    cell_child_no.swap(new_cno);
    boundary_faces.swap(new_bf);
    find_hanging_vertices();
  }

  void Triangulation::find_hanging_vertices()
  {
    hanging.clear();
    const int nv = n_vertices(), vpc = verts_per_cell(), nc = n_cells();
    if (!nv) return;
    // vertex lookup by quantised position
    double lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0};
    for (int d = 0; d < dim; ++d) lo[d] = hi[d] = vertices[d];
    for (int i = 0; i < nv; ++i)
      for (int d = 0; d < dim; ++d)
        {
          lo[d] = std::min(lo[d], vertices[(size_t)i * dim + d]);
          hi[d] = std::max(hi[d], vertices[(size_t)i * dim + d]);
        }
    const double Q = double((1u << 20) - 1);
    auto key_of = [&](const double *x) {
      uint64_t key = 0;
      for (int d = dim - 1; d >= 0; --d)
        {
          const double ext = hi[d] - lo[d];
          const double t = ext > 0 ? (x[d] - lo[d]) / ext : 0.0;
          key = (key << 21) | (uint64_t)std::llround(t * Q * 2.0); // half steps: a midpoint never rounds onto a vertex of its edge
        }
      return key;
    };
    std::unordered_map<uint64_t, int> at;
    at.reserve((size_t)nv * 2);
    for (int i = 0; i < nv; ++i) at.emplace(key_of(&vertices[(size_t)i * dim]), i);
    std::unordered_map<int, Hanging> found;
    auto consider = [&](const int *ids, int n) {
      double x[3] = {0, 0, 0};
      for (int k = 0; k < n; ++k)
        for (int d = 0; d < dim; ++d) x[d] += vertices[(size_t)ids[k] * dim + d];
      for (int d = 0; d < dim; ++d) x[d] /= n;
      auto it = at.find(key_of(x));
      if (it == at.end()) return;
      const int h = it->second;
      for (int k = 0; k < n; ++k)
        if (ids[k] == h) return;
      Hanging hn;
      hn.vertex = h;
      hn.n_masters = n;
      for (int k = 0; k < n; ++k) hn.master[k] = ids[k];
      std::sort(hn.master, hn.master + n);
      auto ins = found.emplace(h, hn);
      if (!ins.second && ins.first->second.n_masters < n) ins.first->second = hn; // (cannot happen on a 2:1 mesh)
    };
    for (int c = 0; c < nc; ++c)
      {
        const int *cv = &cells[(size_t)c * vpc];
        // edges: pairs of corners that differ in one direction
        for (int v = 0; v < vpc; ++v)
          for (int d = 0; d < dim; ++d)
            if (!((v >> d) & 1))
              {
                const int e[2] = {cv[v], cv[v | (1 << d)]};
                consider(e, 2);
              }
        if (dim == 3)
          for (int face = 0; face < 6; ++face)
            {
              const int axis = face / 2, side = face % 2;
              int f[4], k = 0;
              for (int v = 0; v < 8; ++v)
                if (((v >> axis) & 1) == side) f[k++] = cv[v];
              consider(f, 4);
            }
      }
    for (auto &kv : found) hanging.push_back(kv.second);
    std::sort(hanging.begin(), hanging.end(), [](const Hanging &a, const Hanging &b) { return a.vertex < b.vertex; });
    for (const Hanging &h : hanging)
      for (int k = 0; k < h.n_masters; ++k)
        if (found.count(h.master[k]))
          throw std::runtime_error("local refinement: a hanging vertex would depend on another hanging vertex (more than one level of "
                                   "difference across an edge); refine the neighbouring cells as well");
  }

  // ---------------------------------------------------------------------------
  // Patterns
  // ---------------------------------------------------------------------------
  namespace
  {
    // node -> cells adjacency (CSR)
    void node_to_cells(int n_cells, const int *table, int per_cell, int n_nodes, std::vector<int64_t> &ptr,
                       std::vector<int> &adj)
    {
      ptr.assign((size_t)n_nodes + 1, 0);
      for (size_t k = 0; k < (size_t)n_cells * per_cell; ++k) ptr[table[k] + 1]++;
      for (int i = 0; i < n_nodes; ++i) ptr[i + 1] += ptr[i];
      adj.resize(ptr[n_nodes]);
      std::vector<int64_t> pos(ptr.begin(), ptr.end() - 1);
      for (int c = 0; c < n_cells; ++c)
        for (int a = 0; a < per_cell; ++a) adj[pos[table[(size_t)c * per_cell + a]]++] = c;
    }
  } // namespace

  Pattern build_pattern(int n_cells, const int *row_table, int nr, int n_rows, const int *col_table, int ncl, int n_cols)
  {
    Pattern P;
    P.n_rows = n_rows;
    P.n_cols = n_cols;
    std::vector<int64_t> ptr;
    std::vector<int> adj;
    node_to_cells(n_cells, row_table, nr, n_rows, ptr, adj);
    P.rowptr.assign((size_t)n_rows + 1, 0);
    for (int pass = 0; pass < 2; ++pass)
      {
#pragma omp parallel
        {
          std::vector<int> buf;
#pragma omp for schedule(dynamic, 4096)
          for (int r = 0; r < n_rows; ++r)
            {
              buf.clear();
              for (int64_t k = ptr[r]; k < ptr[r + 1]; ++k)
                {
                  const int *cn = col_table + (size_t)adj[k] * ncl;
                  buf.insert(buf.end(), cn, cn + ncl);
                }
              std::sort(buf.begin(), buf.end());
              const int cnt = (int)(std::unique(buf.begin(), buf.end()) - buf.begin());
              if (pass == 0)
                P.rowptr[r + 1] = cnt;
              else
                std::copy(buf.begin(), buf.begin() + cnt, P.col.begin() + P.rowptr[r]);
            }
        }
        if (pass == 0)
          {
            for (int r = 0; r < n_rows; ++r) P.rowptr[r + 1] += P.rowptr[r];
            P.col.resize(P.rowptr[n_rows]);
          }
      }
    return P;
  }

  Pattern build_schur_pattern(const Triangulation &tria, const NodeTable &pn)
  {
    const int nc = tria.n_cells(), vpc = tria.verts_per_cell();
    // cells around each vertex, then cells around each p-node
    std::vector<int64_t> vptr, pptr;
    std::vector<int> vadj, padj;
    node_to_cells(nc, tria.cells.data(), vpc, tria.n_vertices(), vptr, vadj);
    node_to_cells(nc, pn.cell_nodes.data(), pn.nodes_per_cell, pn.n_nodes, pptr, padj);
    Pattern P;
    P.n_rows = P.n_cols = pn.n_nodes;
    P.rowptr.assign((size_t)pn.n_nodes + 1, 0);
    for (int pass = 0; pass < 2; ++pass)
      {
#pragma omp parallel
        {
          std::vector<int> cbuf, buf;
#pragma omp for schedule(dynamic, 1024)
          for (int r = 0; r < pn.n_nodes; ++r)
            {
              cbuf.clear();
              for (int64_t k = pptr[r]; k < pptr[r + 1]; ++k)
                {
                  const int *cv = &tria.cells[(size_t)padj[k] * vpc];
                  for (int v = 0; v < vpc; ++v)
                    for (int64_t m = vptr[cv[v]]; m < vptr[cv[v] + 1]; ++m) cbuf.push_back(vadj[m]);
                }
              std::sort(cbuf.begin(), cbuf.end());
              cbuf.erase(std::unique(cbuf.begin(), cbuf.end()), cbuf.end());
              buf.clear();
              for (int c : cbuf)
                {
                  const int *cn = &pn.cell_nodes[(size_t)c * pn.nodes_per_cell];
                  buf.insert(buf.end(), cn, cn + pn.nodes_per_cell);
                }
              std::sort(buf.begin(), buf.end());
              const int cnt = (int)(std::unique(buf.begin(), buf.end()) - buf.begin());
              if (pass == 0)
                P.rowptr[r + 1] = cnt;
              else
                std::copy(buf.begin(), buf.begin() + cnt, P.col.begin() + P.rowptr[r]);
            }
        }
        if (pass == 0)
          {
            for (int r = 0; r < pn.n_nodes; ++r) P.rowptr[r + 1] += P.rowptr[r];
            P.col.resize(P.rowptr[pn.n_nodes]);
          }
      }
    return P;
  }

  Pattern product_pattern(const Pattern &A, const Pattern &B, int n_cols)
  {
    Pattern P;
    P.n_rows = A.n_rows;
    P.n_cols = n_cols;
    P.rowptr.assign((size_t)A.n_rows + 1, 0);
    for (int pass = 0; pass < 2; ++pass)
      {
#pragma omp parallel
        {
          std::vector<int> buf;
#pragma omp for schedule(dynamic, 1024)
          for (int r = 0; r < A.n_rows; ++r)
            {
              buf.clear();
              for (int64_t k = A.rowptr[r]; k < A.rowptr[r + 1]; ++k)
                {
                  const int m = A.col[k];
                  if (m >= B.n_rows) continue;
                  buf.insert(buf.end(), B.col.begin() + B.rowptr[m], B.col.begin() + B.rowptr[m + 1]);
                }
              std::sort(buf.begin(), buf.end());
              const int cnt = (int)(std::unique(buf.begin(), buf.end()) - buf.begin());
              if (pass == 0)
                P.rowptr[r + 1] = cnt;
              else
                std::copy(buf.begin(), buf.begin() + cnt, P.col.begin() + P.rowptr[r]);
            }
        }
        if (pass == 0)
          {
            for (int r = 0; r < A.n_rows; ++r) P.rowptr[r + 1] += P.rowptr[r];
            P.col.resize(P.rowptr[A.n_rows]);
          }
      }
    return P;
  }

  void colour_cells(int n_cells, const int *table, int per_cell, int n_nodes, std::vector<int> &order,
                    std::vector<int> &offsets)
  {
    std::vector<int64_t> ptr;
    std::vector<int> adj;
    node_to_cells(n_cells, table, per_cell, n_nodes, ptr, adj);
    std::vector<int> colour(n_cells, -1);
    int n_colours = 0;
    std::vector<int> stamp(64, -1);
    for (int c = 0; c < n_cells; ++c)
      {
        for (int a = 0; a < per_cell; ++a)
          {
            const int node = table[(size_t)c * per_cell + a];
            for (int64_t k = ptr[node]; k < ptr[node + 1]; ++k)
              {
                const int col = colour[adj[k]];
                if (col >= 0)
                  {
                    if (col >= (int)stamp.size()) stamp.resize(col + 1, -1);
                    stamp[col] = c;
                  }
              }
          }
        int pick = 0;
        while (pick < (int)stamp.size() && stamp[pick] == c) ++pick;
        if (pick >= (int)stamp.size()) stamp.resize(pick + 1, -1);
        colour[c] = pick;
        n_colours = std::max(n_colours, pick + 1);
      }
    offsets.assign(n_colours + 1, 0);
    for (int c = 0; c < n_cells; ++c) offsets[colour[c] + 1]++;
    for (int k = 0; k < n_colours; ++k) offsets[k + 1] += offsets[k];
    order.resize(n_cells);
    std::vector<int> pos(offsets.begin(), offsets.end() - 1);
    for (int c = 0; c < n_cells; ++c) order[pos[colour[c]]++] = c;
  }
} // namespace ifem
