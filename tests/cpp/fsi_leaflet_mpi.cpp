// The reference's driver tests/fsi_leaflet_mpi/fsi_leaflet_mpi.cpp (2-D branch, lines 18-92; BASELINE config 4): same meshes, same
// band refinement through cell iterators, same hard-coded inflow, MPI::FSI<2>(fluid, solid, params, true).run(). The reference
// driver checks nothing; this one prints the maxima tests/test_zz_config4_gpu.py compares with the same run through the Python mirror.
// Built by tests/test_cpp_facade.py's helper with g++ against libopenifem_b200.so.
#include <openifem/openifem.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <iostream>

const double L = 4, H = 1, a = 0.1, b = 0.4, h = 0.05, U = 1.5;

int main(int argc, char *argv[])
{
  using namespace dealii;
  try
    {
      Utilities::MPI::MPI_InitFinalize mpi_initialization(argc, argv, 1);
      std::string infile("parameters.prm");
      if (argc > 1) infile = argv[1];
      Parameters::AllParameters params(infile);

      auto inflow_bc = [U = U](const Point<2> &p, const unsigned int component, const double time) -> double {
        (void)time;
        if (component == 0 && std::abs(p[0]) < 1e-10) return U;
        return 0.0;
      };

      parallel::distributed::Triangulation<2> fluid_tria(MPI_COMM_WORLD);
      dealii::GridGenerator::subdivided_hyper_rectangle(fluid_tria, {static_cast<unsigned int>(L / h), static_cast<unsigned int>(H / h)},
                                                        Point<2>(0, 0), Point<2>(L, H), true);
      // Refine the middle part
      for (auto cell : fluid_tria.active_cell_iterators())
        {
          auto center = cell->center();
          if (center[0] >= L / 4 - 2 * a && center[0] <= L / 4 + 3 * a && cell->is_locally_owned()) cell->set_refine_flag();
        }
      fluid_tria.execute_coarsening_and_refinement();

      Fluid::MPI::SCnsIM<2> fluid(fluid_tria, params);
      fluid.add_hard_coded_boundary_condition(0, inflow_bc);

      Triangulation<2> solid_tria;
      dealii::GridGenerator::subdivided_hyper_rectangle(solid_tria, {static_cast<unsigned int>(a / h), static_cast<unsigned int>(b / h)},
                                                        Point<2>(L / 4, 0), Point<2>(a + L / 4, b), true);
      Solid::MPI::SharedHyperElasticity<2> solid(solid_tria, params);

      MPI::FSI<2> fsi(fluid, solid, params, true);
      fsi.run();

      auto solution = fluid.get_current_solution();
      const double vmax = Utils::PETScVectorMax(solution.block(0)), pmax = Utils::PETScVectorMax(solution.block(1));
      Vector<double> u(solid.get_current_solution());
      double umax = 0;
      for (double x : u) umax = std::max(umax, std::abs(x));
      std::printf("cells %u vmax %.15e pmax %.15e umax %.15e\n", fluid_tria.n_active_cells(), vmax, pmax, umax);
    }
  catch (std::exception &exc)
    {
      std::cerr << "Exception on processing: " << exc.what() << std::endl;
      return 1;
    }
  return 0;
}
