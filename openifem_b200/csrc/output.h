// On-disk formats either side of the hot path, host code only:
//   * .vtu / .pvtu / .pvd result files with the reference's names and fields (FluidSolver::output_results,
//     source/mpi_fluid_solver.cpp:491-579; SharedSolidSolver::output_results, source/mpi_shared_solid_solver.cpp:237-337;
//     Utils::PVDWriter, source/utilities.cpp:38-81). deal.II's DataOut writes one patch per cell (duplicated vertices,
//     zlib-compressed binary); we write the same fields as an ASCII UnstructuredGrid with shared vertices, sampled like
//     build_patches(fluid_pressure_degree = 1) at the cell vertices; piecewise-constant fields go to CellData.
//   * checkpoints: the solid's three files are deal.II Vector<double>::block_write streams ("<size>\n[" + raw doubles + "]",
//     source/mpi_shared_solid_solver.cpp:452-571) and are written byte-for-byte in that format; the fluid's
//     Triangulation::save / SolutionTransfer serialisation (p4est, :582-713) has no stand-alone specification, so
//     NNNNNN.fluid_checkpoint is our own little-endian record (magic, sizes, time step, present_solution) with the
//     reference's naming, rotation (only the latest is kept) and restart semantics (time and .pvd replayed).
#pragma once
#include <cstdint>
#include <fstream>
#include <string>
#include <vector>

#include "mesh.h"

namespace ifem
{
  namespace io
  {
    struct Field
    {
      std::string name;
      int n_components; // vectors are padded to 3 components in the file, as deal.II does
      std::vector<double> values; // [n][n_components]
    };

    // cells: [n_cells][2^dim] in lexicographic vertex order (converted to VTK_QUAD / VTK_HEXAHEDRON order)
    void write_vtu(const std::string &path, int dim, const std::vector<double> &points, const std::vector<int> &cells,
                   const std::vector<Field> &point_data, const std::vector<Field> &cell_data);
    void write_pvtu(const std::string &path, const std::vector<std::string> &pieces, const std::vector<Field> &point_data,
                    const std::vector<Field> &cell_data);
    // "fluid_000012" / "fluid_000012.proc0003.vtu"
    std::string counter_name(const std::string &base, unsigned int index, int n_digits = 6);
    std::string piece_name(const std::string &base, unsigned int index, int rank);

    // Utils::PVDWriter (source/utilities.cpp:38-81): one <DataSet timestep file> line per output, the collection is
    // closed after every write so the file is always valid
    class PVDWriter
    {
    public:
      explicit PVDWriter(const std::string &filename);
      void write_current_timestep(double time, unsigned int timestep, const std::string &pvtu_prefix, unsigned int n_digits = 6);

    private:
      void write_header();
      std::ofstream doc;
      std::streampos write_pos;
      bool header_written = false;
    };

    // deal.II Vector<double>::block_write / block_read
    void block_write(const std::string &path, const std::vector<double> &v);
    std::vector<double> block_read(const std::string &path);

    struct FluidCheckpoint
    {
      int dim = 0;
      unsigned int timestep = 0;
      double time = 0, bc_time = 0;
      int64_t n_vertices = 0, n_cells = 0;
      std::vector<double> present_solution;
    };
    void save_fluid_checkpoint(const std::string &path, const FluidCheckpoint &c);
    FluidCheckpoint load_fluid_checkpoint(const std::string &path);
    // the file with the given extension whose stem is largest (the reference's "latest checkpoint" rule); "" if none
    std::string latest_with_extension(const std::string &dir, const std::string &extension);
    // keep only the newest file with this extension; the siblings with the other extensions go with it
    void rotate_checkpoints(const std::string &dir, const std::string &extension, const std::vector<std::string> &sibling_extensions);

    // FluidSolver::output_results on host copies of the solver state (local numbering of one rank): velocity / pressure /
    // fsi_force / dummy_fsi_force at the vertices, subdomain / Indicator per cell, Txx Txy Tyy [Txz Tyz Tzz]
    void write_fluid_results(const std::string &dir, unsigned int index, int rank, int n_ranks, int dim, const NodeTable &un,
                             const NodeTable &pn, const std::vector<int> &cells_to_write, const std::vector<double> &present,
                             const std::vector<double> &fsi_acceleration, const std::vector<int> &indicator,
                             const std::vector<double> &stress);
    // SharedSolidSolver::output_results: displacements / velocities, subdomain / material_id, Exx.. / Sxx..
    void write_solid_results(const std::string &dir, unsigned int index, int dim, const NodeTable &nt, const std::vector<int> &material_id,
                             const std::vector<double> &displacement, const std::vector<double> &velocity,
                             const std::vector<double> &strain, const std::vector<double> &stress);
  } // namespace io
} // namespace ifem
