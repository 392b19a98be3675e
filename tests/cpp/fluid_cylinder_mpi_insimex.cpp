// Driver in the style of the reference's tests/fluid_cylinder_mpi_insimex/fluid_cylinder_mpi_insimex.cpp (2-D branch, :27-97):
// flow around a cylinder with the implicit-explicit solver, parabolic inflow on x = 0 (Uavg = 0.2), goldens max velocity
// 0.374062 and max pressure 46.5308 to 1e-3. Built by tests/test_cpp_facade.py with g++ against libopenifem_b200.so.
#include <openifem/openifem.h>

#include <cmath>
#include <iostream>

int main(int argc, char *argv[])
{
  try
    {
      dealii::Utilities::MPI::MPI_InitFinalize mpi_initialization(argc, argv, 1);
      Parameters::AllParameters params(argc > 1 ? argv[1] : "parameters.prm");
      auto inflow_bc = [](const dealii::Point<2> &p, const unsigned int component, const double) -> double {
        const double Umax = 3 * 0.2 / 2;
        return component == 0 && std::abs(p[0]) < 1e-10 ? 4 * Umax * p[1] * (0.41 - p[1]) / (0.41 * 0.41) : 0.0;
      };
      parallel::distributed::Triangulation<2> tria(MPI_COMM_WORLD);
      Utils::GridCreator<2>::flow_around_cylinder(tria);
      Fluid::MPI::InsIMEX<2> flow(tria, params);
      flow.add_hard_coded_boundary_condition(0, inflow_bc);
      flow.run();
      auto solution = flow.get_current_solution();
      const double vmax = Utils::PETScVectorMax(solution.block(0)), pmax = Utils::PETScVectorMax(solution.block(1));
      std::cout << "vmax = " << vmax << " pmax = " << pmax << std::endl;
      if (!(std::abs(vmax - 0.374062) / 0.374062 < 1e-3 && std::abs(pmax - 46.5308) / 46.5308 < 1e-3))
        throw std::runtime_error("Maximum velocity or pressure is incorrect!");
    }
  catch (std::exception &exc)
    {
      std::cerr << "Exception on processing: " << exc.what() << std::endl;
      return 1;
    }
  return 0;
}
