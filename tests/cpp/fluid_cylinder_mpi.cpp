// The reference's regression driver tests/fluid_cylinder_mpi/fluid_cylinder_mpi.cpp (2-D branch, lines 25-93):
// same mesh generator, same .prm, same hard-coded inflow, same goldens (max velocity 0.374235, max pressure 46.5226,
// 1e-3). Built by tests/test_cpp_facade.py with g++ against libopenifem_b200.so.
#include <openifem/openifem.h>

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
      auto inflow_bc = [](const dealii::Point<2> &p, const unsigned int component, const double time) -> double {
        (void)time;
        if (component == 0 && std::abs(p[0] - 0.0) < 1e-10)
          {
            double Uavg = 0.2;
            double Umax = 3 * Uavg / 2;
            return 4 * Umax * p[1] * (0.41 - p[1]) / (0.41 * 0.41);
          }
        return 0.0;
      };
      parallel::distributed::Triangulation<2> tria(MPI_COMM_WORLD);
      Utils::GridCreator<2>::flow_around_cylinder(tria);
      Fluid::MPI::InsIM<2> flow(tria, params);
      flow.add_hard_coded_boundary_condition(0, inflow_bc);
      flow.run();
      auto solution = flow.get_current_solution();
      auto v = solution.block(0), p = solution.block(1);
      double vmax = Utils::PETScVectorMax(v);
      double pmax = Utils::PETScVectorMax(p);
      double verror = std::abs(vmax - 0.374235) / 0.374235;
      double perror = std::abs(pmax - 46.5226) / 46.5226;
      std::cout << "vmax = " << vmax << " pmax = " << pmax << std::endl;
      if (!(verror < 1e-3 && perror < 1e-3)) throw std::runtime_error("Maximum velocity or pressure is incorrect!");
    }
  catch (std::exception &exc)
    {
      std::cerr << "Exception on processing: " << exc.what() << std::endl;
      return 1;
    }
  return 0;
}
