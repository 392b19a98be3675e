// The reference's regression driver tests/fsi_contact_model_mpi/fsi_contact_model_mpi.cpp (2-D branch, lines 28-60): same
// meshes, same .prm, same penetration criterion, same golden (minimum solid displacement -0.01999, 1e-3).
// Built by tests/test_cpp_facade.py with g++ against libopenifem_b200.so.
#include <openifem/openifem.h>

#include <algorithm>
#include <cmath>
#include <iostream>

int main(int argc, char *argv[])
{
  try
    {
      dealii::Utilities::MPI::MPI_InitFinalize mpi_initialization(argc, argv, 1);
      std::string infile("parameters.prm");
      if (argc > 1) infile = argv[1];
      Parameters::AllParameters params(infile);

      parallel::distributed::Triangulation<2> tria_fluid(MPI_COMM_WORLD);
      GridGenerator::subdivided_hyper_rectangle(tria_fluid, {50, 25}, Point<2>(0, 0), Point<2>(2.0, 1.0), true);

      Triangulation<2> tria_solid;
      GridGenerator::subdivided_hyper_rectangle(tria_solid, {10, 11}, Point<2>(0, 0), Point<2>(1.0, 1.02), true);
      Tensor<1, 2> offset({0.25, 0});
      GridTools::shift(offset, tria_solid);

      Fluid::MPI::SCnsIM<2> fluid(tria_fluid, params);
      Solid::MPI::SharedLinearElasticity<2> solid(tria_solid, params);

      auto penetration_criterion = [](const Point<2> &p) -> double {
        double wall_height = 1.0;
        return (p[1] - wall_height);
      };

      MPI::FSI<2> fsi(fluid, solid, params);
      fsi.set_penetration_criterion(penetration_criterion, Tensor<1, 2>({0, -1}));
      fsi.run();
      Vector<double> u(solid.get_current_solution());
      double umin = *std::min_element(u.begin(), u.end());
      double uerror = std::abs(umin + 0.01999) / 0.01999;
      std::cout << "umin = " << umin << std::endl;
      if (!(uerror < 1e-3)) throw std::runtime_error("Minimum displacement is incorrect!");
    }
  catch (std::exception &exc)
    {
      std::cerr << "Exception on processing: " << exc.what() << std::endl;
      return 1;
    }
  return 0;
}
