// Result files and checkpoints of the solvers (formats in output.h): FluidSolver::output_results / save_checkpoint /
// load_checkpoint (reference source/mpi_fluid_solver.cpp:491-713) and SharedSolidSolver::output_results / save_checkpoint /
// load_checkpoint (source/mpi_shared_solid_solver.cpp:237-337, 452-571). Nothing is written unless an output directory has
// been set (the library never writes into the working directory on its own; the reference always does).
#include <algorithm>
#include <filesystem>

#include "comm.h"
#include "insim.h"
#include "output.h"
#include "partition.h"
#include "solid.h"

namespace ifem
{
  void InsIM::set_output_directory(const std::string &dir)
  {
    output_directory = dir;
    pvd_writer.reset();
    if (dir.empty()) return;
    std::filesystem::create_directories(dir);
    if ((ctx.comm ? ctx.comm->rank : 0) == 0) pvd_writer.reset(new io::PVDWriter((dir == "." ? std::string() : dir + "/") + "fluid.pvd"));
  }

  void InsIM::output_results(unsigned int output_index)
  {
    if (output_directory.empty()) throw std::runtime_error("output_results: no output directory set");
    const std::vector<double> present = present_solution.to_host(ctx.stream), acc = fsi_acceleration.to_host(ctx.stream);
    const std::vector<double> st = stress.n ? stress.to_host(ctx.stream) : std::vector<double>();
    std::vector<int> ind(fs.n_cells, 0);
    if (fs.d_indicator.n) fs.d_indicator.download(ind.data(), fs.n_cells, ctx.stream);
    IFEM_CUDA(cudaStreamSynchronize(ctx.stream));
    // every rank writes the cells of its own slab (cell->is_locally_owned())
    std::vector<int> mine;
    if (fs.n_ranks > 1)
      {
        const std::vector<int> ranks = slab_cell_ranks(triangulation, fs.n_ranks);
        for (int l = 0; l < fs.n_cells; ++l)
          if (ranks[fs.local_cells[l]] == fs.rank) mine.push_back(l);
      }
    else
      for (int l = 0; l < fs.n_cells; ++l) mine.push_back(l);
    io::write_fluid_results(output_directory, output_index, fs.rank, fs.n_ranks, fs.dim, fs.un, fs.pn, mine, present, acc, ind, st);
    if (fs.rank == 0 && pvd_writer) pvd_writer->write_current_timestep(time.current(), time.get_timestep(), "fluid_", 6);
  }

  void InsIM::save_checkpoint(int output_index)
  {
    if (output_directory.empty()) throw std::runtime_error("save_checkpoint: no output directory set");
    io::FluidCheckpoint c;
    c.dim = fs.dim;
    c.timestep = time.get_timestep();
    c.time = time.current();
    c.bc_time = bc_time;
    c.n_vertices = triangulation.n_vertices();
    c.n_cells = triangulation.n_cells();
    c.present_solution = present_solution.to_host(ctx.stream);
    if (fs.n_ranks > 1)
      {
        // the record holds the solution in the GLOBAL numbering, so a run can be continued on any number of ranks (as the
        // reference's p4est-based checkpoints can): every rank contributes its owned entries to a sum all-reduce
        const int dim = fs.dim;
        const int64_t nug = (int64_t)dim * fs.un_global.n_nodes, ng = nug + fs.pn_global.n_nodes;
        std::vector<double> g((size_t)ng, 0.0);
        for (int l = 0; l < fs.n_owned_unodes; ++l)
          for (int d = 0; d < dim; ++d) g[(size_t)dim * fs.part.u.local_to_global[l] + d] = c.present_solution[(size_t)dim * l + d];
        for (int l = 0; l < fs.n_owned_pnodes; ++l) g[(size_t)nug + fs.part.p.local_to_global[l]] = c.present_solution[(size_t)fs.n_u + l];
        DevBuf<double> d((size_t)ng);
        d.upload(g, ctx.stream);
        for (int64_t off = 0; off < ng; off += (1 << 28)) // all-reduce counts are ints
          comm_allreduce_sum(*ctx.comm, d.p + off, (int)std::min<int64_t>(ng - off, 1 << 28), ctx.stream);
        c.present_solution = d.to_host(ctx.stream);
      }
    if (fs.rank != 0) return;
    io::rotate_checkpoints(output_directory, ".fluid_checkpoint", {});
    char name[64];
    std::snprintf(name, sizeof name, "%06d.fluid_checkpoint", output_index);
    io::save_fluid_checkpoint(output_directory + "/" + name, c);
  }

  bool InsIM::load_checkpoint()
  {
    if (output_directory.empty()) return false;
    const std::string file = io::latest_with_extension(output_directory, ".fluid_checkpoint");
    if (file.empty()) return false; // "Did not find fluid checkpoint files. Start from the beginning !"
    const io::FluidCheckpoint c = io::load_fluid_checkpoint(file);
    // triangulation.load() restores the refined mesh in the reference; here the caller's mesh is refined as run() does and
    // must then agree with the one the checkpoint was written on
    if (!dofs_ready)
      {
        if (triangulation.n_cells() != c.n_cells) triangulation.refine_global(parameters.global_refinements.empty() ? 0 : parameters.global_refinements[0]);
        if (triangulation.n_cells() != c.n_cells || triangulation.n_vertices() != c.n_vertices || triangulation.dim != c.dim)
          throw std::runtime_error("load_checkpoint: " + file + " was written on a different mesh");
        setup_dofs();
        make_constraints();
        initialize_system();
      }
    if (fs.n_ranks > 1)
      {
        // global record -> owned and ghost entries of this rank
        const int dim = fs.dim;
        const int64_t nug = (int64_t)dim * fs.un_global.n_nodes, ng = nug + fs.pn_global.n_nodes;
        if ((int64_t)c.present_solution.size() != ng) throw std::runtime_error("load_checkpoint: " + file + " has a different number of dofs");
        std::vector<double> local((size_t)fs.n_dofs);
        for (int l = 0; l < fs.un.n_nodes; ++l)
          for (int d = 0; d < dim; ++d) local[(size_t)dim * l + d] = c.present_solution[(size_t)dim * fs.part.u.local_to_global[l] + d];
        for (int l = 0; l < fs.pn.n_nodes; ++l) local[(size_t)fs.n_u + l] = c.present_solution[(size_t)nug + fs.part.p.local_to_global[l]];
        present_solution.upload(local, ctx.stream);
      }
    else
      {
        if ((int64_t)c.present_solution.size() != fs.n_dofs) throw std::runtime_error("load_checkpoint: " + file + " has a different number of dofs");
        present_solution.upload(c.present_solution, ctx.stream);
      }
    IFEM_CUDA(cudaStreamSynchronize(ctx.stream));
    // the nodal viscous stress is a function of present_solution; the reference leaves it zero until the next step, which
    // makes the first FSI pass after a restart (find_solid_bc reads it) differ from the uninterrupted run
    update_stress();
    // set the current time and write a correct .pvd (:689-708); the clock of time-dependent boundary functions follows
    const int stem = std::stoi(std::filesystem::path(file).stem().string());
    for (int i = 0; i <= stem; ++i)
      {
        if ((time.current() == 0 || time.time_to_output()) && fs.rank == 0 && pvd_writer)
          pvd_writer->write_current_timestep(time.current(), time.get_timestep(), "fluid_", 6);
        if (i == stem) break;
        time.increment();
        if (!hard_coded.empty()) bc_time += time.get_delta_t();
      }
    return true;
  }

  void InsIM::io_before_step()
  {
    if (!output_directory.empty() && time.get_timestep() == 0) output_results(0); // mpi_insim.cpp:403-406
  }

  void InsIM::io_after_step()
  {
    if (output_directory.empty()) return;
    if (time.time_to_output()) output_results(time.get_timestep());
    if (parameters.simulation_type == "Fluid" && time.time_to_save()) save_checkpoint((int)time.get_timestep());
  }

  // ---- solids ---------------------------------------------------------------------------------------------------------
  void SolidSolver::set_output_directory(const std::string &dir)
  {
    output_directory = dir;
    pvd_writer.reset();
    if (dir.empty()) return;
    std::filesystem::create_directories(dir);
    // the solid is replicated on every rank: only rank 0 writes (mpi_shared_solid_solver.cpp:243-246, :460)
    if ((ctx.comm ? ctx.comm->rank : 0) == 0) pvd_writer.reset(new io::PVDWriter((dir == "." ? std::string() : dir + "/") + "solid.pvd"));
  }

  void SolidSolver::output_results(unsigned int output_index)
  {
    if (output_directory.empty()) throw std::runtime_error("output_results: no output directory set");
    if (ctx.comm && ctx.comm->rank != 0) return;
    const std::vector<double> u = current_displacement.to_host(ctx.stream), v = current_velocity.to_host(ctx.stream);
    const std::vector<double> e = strain.n ? strain.to_host(ctx.stream) : std::vector<double>();
    const std::vector<double> s = stress.n ? stress.to_host(ctx.stream) : std::vector<double>();
    io::write_solid_results(output_directory, output_index, ss.dim, ss.nt, triangulation.material_id, u, v, e, s);
    if (pvd_writer) pvd_writer->write_current_timestep(time.current(), time.get_timestep(), "solid_", 6);
  }

  void SolidSolver::save_checkpoint(int output_index)
  {
    if (output_directory.empty()) throw std::runtime_error("save_checkpoint: no output directory set");
    if (ctx.comm && ctx.comm->rank != 0) return;
    io::rotate_checkpoints(output_directory, ".solid_checkpoint_displacement", {".solid_checkpoint_velocity", ".solid_checkpoint_acceleration"});
    char stem[32];
    std::snprintf(stem, sizeof stem, "%06d", output_index);
    const std::string base = output_directory + "/" + stem;
    io::block_write(base + ".solid_checkpoint_displacement", current_displacement.to_host(ctx.stream));
    io::block_write(base + ".solid_checkpoint_velocity", current_velocity.to_host(ctx.stream));
    io::block_write(base + ".solid_checkpoint_acceleration", current_acceleration.to_host(ctx.stream));
  }

  bool SolidSolver::load_checkpoint()
  {
    if (output_directory.empty()) return false;
    const std::string file = io::latest_with_extension(output_directory, ".solid_checkpoint_displacement");
    if (file.empty()) return false;
    if (!dofs_ready)
      {
        setup_dofs();
        initialize_system();
      }
    std::filesystem::path p(file);
    const std::vector<double> u = io::block_read(p.string());
    p.replace_extension(".solid_checkpoint_velocity");
    const std::vector<double> v = io::block_read(p.string());
    p.replace_extension(".solid_checkpoint_acceleration");
    const std::vector<double> a = io::block_read(p.string());
    if ((int64_t)u.size() != ss.n_dofs || v.size() != u.size() || a.size() != u.size())
      throw std::runtime_error("load_checkpoint: " + file + " has a different number of dofs");
    cudaStream_t s = ctx.stream;
    current_displacement.upload(u, s);
    previous_displacement.upload(u, s);
    current_velocity.upload(v, s);
    previous_velocity.upload(v, s);
    current_acceleration.upload(a, s);
    previous_acceleration.upload(a, s);
    IFEM_CUDA(cudaStreamSynchronize(s));
    const int stem = std::stoi(std::filesystem::path(file).stem().string());
    for (int i = 0; i <= stem; ++i)
      {
        if ((time.current() == 0 || time.time_to_output()) && pvd_writer) pvd_writer->write_current_timestep(time.current(), time.get_timestep(), "solid_", 6);
        if (i == stem) break;
        time.increment();
      }
    return true;
  }

  void SolidSolver::io_before_step()
  {
    if (!output_directory.empty() && time.get_timestep() == 0) output_results(0);
  }

  void SolidSolver::io_after_step()
  {
    if (output_directory.empty()) return;
    if (time.time_to_output()) output_results(time.get_timestep());
    if (parameters.simulation_type == "Solid" && time.time_to_save()) save_checkpoint((int)time.get_timestep());
  }
} // namespace ifem
