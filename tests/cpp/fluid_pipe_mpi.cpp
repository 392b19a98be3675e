// The reference's own regression driver tests/fluid_pipe_mpi/fluid_pipe_mpi.cpp (2-D branch, lines 30-56),
// rewritten only where deal.II's AssertThrow / includes are concerned: same mesh, same .prm, same golden
// (max velocity 1.5 +- 1e-2). Built by tests/test_cpp_facade.py with g++ against libopenifem_b200.so.
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
      double L = 2.0, D = 0.2, h = 0.04;
      parallel::distributed::Triangulation<2> tria(MPI_COMM_WORLD);
      dealii::GridGenerator::subdivided_hyper_rectangle(tria, {static_cast<unsigned int>(L / h), static_cast<unsigned int>(D / h)},
                                                        dealii::Point<2>(0, 0), dealii::Point<2>(L, D), true);
      Fluid::MPI::InsIM<2> flow(tria, params);
      flow.run();
      auto solution = flow.get_current_solution();
      auto v = solution.block(0);
      double vmax = Utils::PETScVectorMax(v);
      double verror = std::abs(vmax - 1.5) / 1.5;
      std::cout << "vmax = " << vmax << std::endl;
      if (!(verror < 1e-2)) throw std::runtime_error("Maximum velocity is incorrect!");
    }
  catch (std::exception &exc)
    {
      std::cerr << "Exception on processing: " << exc.what() << std::endl;
      return 1;
    }
  return 0;
}
