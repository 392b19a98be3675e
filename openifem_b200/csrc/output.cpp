// See output.h. Reference: source/mpi_fluid_solver.cpp:491-713, source/mpi_shared_solid_solver.cpp:237-337, 452-571,
// source/utilities.cpp:38-81.
#include "output.h"

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <ctime>
#include <filesystem>
#include <map>
#include <set>
#include <stdexcept>

namespace fs = std::filesystem;

namespace ifem
{
  namespace io
  {
    namespace
    {
      void write_array(std::ofstream &out, const char *type, const std::string &name, int ncomp_file, size_t n, int ncomp,
                       const double *v)
      {
        out << "<DataArray type=\"" << type << "\"";
        if (!name.empty()) out << " Name=\"" << name << "\"";
        if (ncomp_file > 1) out << " NumberOfComponents=\"" << ncomp_file << "\"";
        out << " format=\"ascii\">\n";
        char buf[40];
        for (size_t i = 0; i < n; ++i)
          {
            for (int c = 0; c < ncomp_file; ++c)
              {
                std::snprintf(buf, sizeof buf, "%.17g", c < ncomp ? v[i * ncomp + c] : 0.0);
                out << buf << (c + 1 < ncomp_file ? ' ' : '\n');
              }
          }
        out << "</DataArray>\n";
      }
      int file_components(int n) { return n > 1 ? 3 : 1; }
    } // namespace

    std::string counter_name(const std::string &base, unsigned int index, int n_digits)
    {
      char buf[64];
      std::snprintf(buf, sizeof buf, "%s_%0*u", base.c_str(), n_digits, index);
      return buf;
    }
    std::string piece_name(const std::string &base, unsigned int index, int rank)
    {
      char buf[32];
      std::snprintf(buf, sizeof buf, ".proc%04d.vtu", rank);
      return counter_name(base, index) + buf;
    }

    void write_vtu(const std::string &path, int dim, const std::vector<double> &points, const std::vector<int> &cells,
                   const std::vector<Field> &point_data, const std::vector<Field> &cell_data)
    {
      const int nv = 1 << dim;
      const size_t n_points = points.size() / dim, n_cells = cells.size() / nv;
      std::ofstream out(path);
      if (!out) throw std::runtime_error("cannot open " + path + " for writing");
      out << "<?xml version=\"1.0\"?>\n<VTKFile type=\"UnstructuredGrid\" version=\"0.1\" byte_order=\"LittleEndian\">\n"
          << "<UnstructuredGrid>\n<Piece NumberOfPoints=\"" << n_points << "\" NumberOfCells=\"" << n_cells << "\">\n<Points>\n";
      write_array(out, "Float64", "", 3, n_points, dim, points.data());
      out << "</Points>\n<Cells>\n<DataArray type=\"Int32\" Name=\"connectivity\" format=\"ascii\">\n";
      static const int vtk2[4] = {0, 1, 3, 2}, vtk3[8] = {0, 1, 3, 2, 4, 5, 7, 6}; // lexicographic -> VTK_QUAD / VTK_HEXAHEDRON
      for (size_t c = 0; c < n_cells; ++c)
        for (int v = 0; v < nv; ++v) out << cells[c * nv + (dim == 2 ? vtk2[v] : vtk3[v])] << (v + 1 < nv ? ' ' : '\n');
      out << "</DataArray>\n<DataArray type=\"Int32\" Name=\"offsets\" format=\"ascii\">\n";
      for (size_t c = 0; c < n_cells; ++c) out << (c + 1) * nv << '\n';
      out << "</DataArray>\n<DataArray type=\"UInt8\" Name=\"types\" format=\"ascii\">\n";
      for (size_t c = 0; c < n_cells; ++c) out << (dim == 2 ? 9 : 12) << '\n';
      out << "</DataArray>\n</Cells>\n<PointData Scalars=\"scalars\">\n";
      for (const Field &f : point_data)
        {
          if (f.values.size() != n_points * f.n_components) throw std::runtime_error("write_vtu: point field " + f.name + " has the wrong size");
          write_array(out, "Float64", f.name, file_components(f.n_components), n_points, f.n_components, f.values.data());
        }
      out << "</PointData>\n<CellData>\n";
      for (const Field &f : cell_data)
        {
          if (f.values.size() != n_cells * f.n_components) throw std::runtime_error("write_vtu: cell field " + f.name + " has the wrong size");
          write_array(out, "Float64", f.name, file_components(f.n_components), n_cells, f.n_components, f.values.data());
        }
      out << "</CellData>\n</Piece>\n</UnstructuredGrid>\n</VTKFile>\n";
    }

    void write_pvtu(const std::string &path, const std::vector<std::string> &pieces, const std::vector<Field> &point_data,
                    const std::vector<Field> &cell_data)
    {
      std::ofstream out(path);
      if (!out) throw std::runtime_error("cannot open " + path + " for writing");
      out << "<?xml version=\"1.0\"?>\n<VTKFile type=\"PUnstructuredGrid\" version=\"0.1\" byte_order=\"LittleEndian\">\n"
          << "<PUnstructuredGrid GhostLevel=\"0\">\n<PPointData Scalars=\"scalars\">\n";
      auto decl = [&](const Field &f) {
        out << "<PDataArray type=\"Float64\" Name=\"" << f.name << "\"";
        if (f.n_components > 1) out << " NumberOfComponents=\"3\"";
        out << " format=\"ascii\"/>\n";
      };
      for (const Field &f : point_data) decl(f);
      out << "</PPointData>\n<PCellData>\n";
      for (const Field &f : cell_data) decl(f);
      out << "</PCellData>\n<PPoints>\n<PDataArray type=\"Float64\" NumberOfComponents=\"3\"/>\n</PPoints>\n";
      for (const std::string &p : pieces) out << "<Piece Source=\"" << p << "\"/>\n";
      out << "</PUnstructuredGrid>\n</VTKFile>\n";
    }

    PVDWriter::PVDWriter(const std::string &filename)
    {
      doc.open(filename, std::ios::out);
      if (!doc) throw std::runtime_error("cannot open " + filename + " for writing");
      doc.precision(12);
    }

    void PVDWriter::write_header()
    {
      const std::time_t t = std::time(nullptr);
      char date[32], clock[32];
      std::strftime(date, sizeof date, "%Y/%m/%d", std::localtime(&t));
      std::strftime(clock, sizeof clock, "%H:%M:%S", std::localtime(&t));
      doc << "<?xml version=\"1.0\"?>\n<!--\n#This file was generated by OpenIFEM on " << date << " at " << clock << "\n-->\n";
      doc << "<VTKFile type=\"Collection\" version=\"0.1\" ByteOrder=\"LittleEndian\">\n  <Collection>" << std::endl;
      write_pos = doc.tellp();
      header_written = true;
    }

    void PVDWriter::write_current_timestep(double time, unsigned int timestep, const std::string &pvtu_prefix, unsigned int n_digits)
    {
      if (time == 0 || !header_written) write_header(); // utilities.cpp:55-58 (a restart replays from time 0)
      doc.seekp(write_pos);
      char num[32];
      std::snprintf(num, sizeof num, "%0*u", (int)n_digits, timestep);
      doc << "    <DataSet timestep=\"" << time << "\" group=\"\" part=\"0\" file=\"" << pvtu_prefix << num << ".pvtu\"/>\n";
      write_pos = doc.tellp();
      doc << "  </Collection>\n</VTKFile>" << std::endl;
    }

    void block_write(const std::string &path, const std::vector<double> &v)
    {
      std::ofstream out(path, std::ios::binary);
      if (!out) throw std::runtime_error("cannot open " + path + " for writing");
      char buf[32];
      std::snprintf(buf, sizeof buf, "%llu\n[", (unsigned long long)v.size());
      out.write(buf, (std::streamsize)std::strlen(buf));
      out.write(reinterpret_cast<const char *>(v.data()), (std::streamsize)(v.size() * sizeof(double)));
      out.write("]", 1);
    }

    std::vector<double> block_read(const std::string &path)
    {
      std::ifstream in(path, std::ios::binary);
      if (!in) throw std::runtime_error("cannot open " + path);
      std::string line;
      std::getline(in, line);
      const unsigned long long n = std::stoull(line);
      char c = 0;
      in.read(&c, 1);
      if (c != '[') throw std::runtime_error(path + ": not a block_write stream");
      std::vector<double> v(n);
      in.read(reinterpret_cast<char *>(v.data()), (std::streamsize)(n * sizeof(double)));
      in.read(&c, 1);
      if (!in || c != ']') throw std::runtime_error(path + ": truncated block_write stream");
      return v;
    }

    static const char kMagic[8] = {'I', 'F', 'E', 'M', 'C', 'K', 'P', '1'};

    void save_fluid_checkpoint(const std::string &path, const FluidCheckpoint &c)
    {
      std::ofstream out(path, std::ios::binary);
      if (!out) throw std::runtime_error("cannot open " + path + " for writing");
      auto put = [&](const void *p, size_t n) { out.write(reinterpret_cast<const char *>(p), (std::streamsize)n); };
      const int32_t dim = c.dim;
      const uint32_t ts = c.timestep;
      const int64_t n = (int64_t)c.present_solution.size();
      put(kMagic, 8);
      put(&dim, 4);
      put(&ts, 4);
      put(&c.time, 8);
      put(&c.bc_time, 8);
      put(&c.n_vertices, 8);
      put(&c.n_cells, 8);
      put(&n, 8);
      put(c.present_solution.data(), (size_t)n * 8);
    }

    FluidCheckpoint load_fluid_checkpoint(const std::string &path)
    {
      std::ifstream in(path, std::ios::binary);
      if (!in) throw std::runtime_error("cannot open " + path);
      auto get = [&](void *p, size_t n) { in.read(reinterpret_cast<char *>(p), (std::streamsize)n); };
      char magic[8];
      get(magic, 8);
      if (!in || std::memcmp(magic, kMagic, 8) != 0) throw std::runtime_error(path + ": not a fluid checkpoint of this library");
      FluidCheckpoint c;
      int32_t dim;
      uint32_t ts;
      int64_t n;
      get(&dim, 4);
      get(&ts, 4);
      get(&c.time, 8);
      get(&c.bc_time, 8);
      get(&c.n_vertices, 8);
      get(&c.n_cells, 8);
      get(&n, 8);
      if (!in || n < 0) throw std::runtime_error(path + ": truncated fluid checkpoint");
      c.dim = dim;
      c.timestep = ts;
      c.present_solution.resize((size_t)n);
      get(c.present_solution.data(), (size_t)n * 8);
      if (!in) throw std::runtime_error(path + ": truncated fluid checkpoint");
      return c;
    }

    std::string latest_with_extension(const std::string &dir, const std::string &extension)
    {
      std::string best;
      for (const auto &p : fs::directory_iterator(dir.empty() ? "." : dir))
        if (p.path().extension() == extension && (best.empty() || p.path().stem().string() > fs::path(best).stem().string())) best = p.path().string();
      return best;
    }

    void rotate_checkpoints(const std::string &dir, const std::string &extension, const std::vector<std::string> &siblings)
    {
      std::set<fs::path> found;
      for (const auto &p : fs::directory_iterator(dir.empty() ? "." : dir))
        if (p.path().extension() == extension) found.insert(p.path());
      while (found.size() > 1)
        {
          fs::path victim(*found.begin());
          fs::remove(victim);
          for (const std::string &e : siblings)
            {
              fs::path s(victim);
              s.replace_extension(e);
              fs::remove(s);
            }
          found.erase(found.begin());
        }
    }

    namespace
    {
      // local index of vertex v (bits = position along each axis) in the lexicographic FE_Q(p) lattice
      int vertex_local_node(int dim, int p, int v)
      {
        int idx = 0, stride = 1;
        for (int d = 0; d < dim; ++d)
          {
            idx += ((v >> d) & 1) * p * stride;
            stride *= p + 1;
          }
        return idx;
      }
      std::string join(const std::string &dir, const std::string &file) { return dir.empty() || dir == "." ? file : dir + "/" + file; }
      const char *const kComp[3] = {"x", "y", "z"};
    } // namespace

    void write_fluid_results(const std::string &dir, unsigned int index, int rank, int n_ranks, int dim, const NodeTable &un,
                             const NodeTable &pn, const std::vector<int> &cells_to_write, const std::vector<double> &present,
                             const std::vector<double> &fsi_acceleration, const std::vector<int> &indicator,
                             const std::vector<double> &stress)
    {
      const int nv = 1 << dim;
      const int64_t n_u = (int64_t)dim * un.n_nodes;
      std::map<int, int> point_of_unode;
      std::vector<int> unode_of_point, pnode_of_point, cells;
      for (int c : cells_to_write)
        for (int v = 0; v < nv; ++v)
          {
            const int u = un.cell_nodes[(size_t)c * un.nodes_per_cell + vertex_local_node(dim, un.p, v)];
            const int p = pn.cell_nodes[(size_t)c * pn.nodes_per_cell + vertex_local_node(dim, pn.p, v)];
            auto it = point_of_unode.find(u);
            if (it == point_of_unode.end())
              {
                it = point_of_unode.emplace(u, (int)unode_of_point.size()).first;
                unode_of_point.push_back(u);
                pnode_of_point.push_back(p);
              }
            cells.push_back(it->second);
          }
      const size_t np = unode_of_point.size();
      std::vector<double> points(np * dim);
      Field vel{"velocity", dim, std::vector<double>(np * dim)}, pres{"pressure", 1, std::vector<double>(np)};
      Field force{"fsi_force", dim, std::vector<double>(np * dim)}, dummy{"dummy_fsi_force", 1, std::vector<double>(np)};
      for (size_t i = 0; i < np; ++i)
        {
          const int u = unode_of_point[i], p = pnode_of_point[i];
          for (int d = 0; d < dim; ++d)
            {
              points[i * dim + d] = un.coords[(size_t)u * dim + d];
              vel.values[i * dim + d] = present[(size_t)dim * u + d];
              force.values[i * dim + d] = fsi_acceleration.empty() ? 0.0 : fsi_acceleration[(size_t)dim * u + d];
            }
          pres.values[i] = present[(size_t)n_u + p];
          dummy.values[i] = fsi_acceleration.empty() ? 0.0 : fsi_acceleration[(size_t)n_u + p];
        }
      std::vector<Field> pd{vel, pres, force, dummy};
      // Txx, Txy, Tyy [, Txz, Tyz, Tzz] (:551-560)
      static const int order2[3][2] = {{0, 0}, {0, 1}, {1, 1}}, order3[6][2] = {{0, 0}, {0, 1}, {1, 1}, {0, 2}, {1, 2}, {2, 2}};
      for (int k = 0; k < (dim == 2 ? 3 : 6); ++k)
        {
          const int i = dim == 2 ? order2[k][0] : order3[k][0], j = dim == 2 ? order2[k][1] : order3[k][1];
          Field t{std::string("T") + kComp[i] + kComp[j], 1, std::vector<double>(np)};
          for (size_t q = 0; q < np; ++q) t.values[q] = stress.empty() ? 0.0 : stress[(size_t)(i * dim + j) * un.n_nodes + unode_of_point[q]];
          pd.push_back(std::move(t));
        }
      Field sub{"subdomain", 1, std::vector<double>(cells_to_write.size(), (double)rank)}, ind{"Indicator", 1, std::vector<double>(cells_to_write.size())};
      for (size_t k = 0; k < cells_to_write.size(); ++k) ind.values[k] = indicator.empty() ? 0.0 : indicator[cells_to_write[k]];
      std::vector<Field> cd{sub, ind};
      write_vtu(join(dir, piece_name("fluid", index, rank)), dim, points, cells, pd, cd);
      if (rank == 0)
        {
          std::vector<std::string> pieces;
          for (int r = 0; r < n_ranks; ++r) pieces.push_back(piece_name("fluid", index, r));
          write_pvtu(join(dir, counter_name("fluid", index) + ".pvtu"), pieces, pd, cd);
        }
    }

    void write_solid_results(const std::string &dir, unsigned int index, int dim, const NodeTable &nt, const std::vector<int> &material_id,
                             const std::vector<double> &displacement, const std::vector<double> &velocity,
                             const std::vector<double> &strain, const std::vector<double> &stress)
    {
      const int nv = 1 << dim, n_cells = (int)(nt.cell_nodes.size() / nt.nodes_per_cell);
      std::vector<int> cells((size_t)n_cells * nv);
      for (int c = 0; c < n_cells; ++c)
        for (int v = 0; v < nv; ++v) cells[(size_t)c * nv + v] = nt.cell_nodes[(size_t)c * nt.nodes_per_cell + vertex_local_node(dim, nt.p, v)];
      const size_t np = (size_t)nt.n_nodes;
      std::vector<Field> pd{{"displacements", dim, displacement}, {"velocities", dim, velocity}};
      static const int order2[3][2] = {{0, 0}, {0, 1}, {1, 1}}, order3[6][2] = {{0, 0}, {0, 1}, {1, 1}, {0, 2}, {1, 2}, {2, 2}};
      for (int which = 0; which < 2; ++which) // Exx Exy Eyy, Sxx Sxy Syy [, .xz .yz .zz] (:299-325)
        {
          const std::vector<double> &src = which == 0 ? strain : stress;
          for (int pass = 0; pass < 2; ++pass) // the reference lists the in-plane components first, then the z ones
            for (int k = 0; k < (dim == 2 ? 3 : 6); ++k)
              {
                if ((pass == 0) != (k < 3)) continue;
                const int i = dim == 2 ? order2[k][0] : order3[k][0], j = dim == 2 ? order2[k][1] : order3[k][1];
                Field t{std::string(which == 0 ? "E" : "S") + kComp[i] + kComp[j], 1, std::vector<double>(np)};
                for (size_t q = 0; q < np; ++q) t.values[q] = src.empty() ? 0.0 : src[(size_t)(i * dim + j) * np + q];
                pd.push_back(std::move(t));
              }
        }
      Field sub{"subdomain", 1, std::vector<double>((size_t)n_cells, 0.0)}, mat{"material_id", 1, std::vector<double>((size_t)n_cells)};
      for (int c = 0; c < n_cells; ++c) mat.values[c] = material_id.empty() ? 0.0 : material_id[c];
      std::vector<Field> cd{sub, mat};
      write_vtu(join(dir, piece_name("solid", index, 0)), dim, nt.coords, cells, pd, cd);
      write_pvtu(join(dir, counter_name("solid", index) + ".pvtu"), {piece_name("solid", index, 0)}, pd, cd);
    }
  } // namespace io
} // namespace ifem
