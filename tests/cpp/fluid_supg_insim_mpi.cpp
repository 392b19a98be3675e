// Driver in the style of the reference's regression programs tests/fluid_pressure_driven_mpi_insim_supg/...cpp:32-58 and
// tests/fluid_plane_wall_driven_mpi_insim_supg/...cpp:32-51 (2-D branches), both cases in one executable:
//   argv[1] = "pressure": 100 x 10 cells on 2 x 0.2 (refined once by the .prm), largest velocity within 2 % and the 30th largest
//             within 1e-3 of 2.5e-2;   argv[1] = "wall": 20 x 16 cells on 2 x 0.4, l2 norm of the velocity 4.7112 +- 1e-3.
// Built by tests/test_cpp_facade.py with g++ against libopenifem_b200.so.
#include <openifem/openifem.h>

#include <algorithm>
#include <cmath>
#include <functional>
#include <iostream>

int main(int argc, char *argv[])
{
  try
    {
      dealii::Utilities::MPI::MPI_InitFinalize mpi_initialization(argc, argv, 1);
      if (argc < 3) throw std::runtime_error("usage: fluid_supg_insim_mpi pressure|wall parameters.prm");
      const std::string which(argv[1]);
      Parameters::AllParameters params(argv[2]);
      parallel::distributed::Triangulation<2> tria(MPI_COMM_WORLD);
      if (which == "pressure")
        dealii::GridGenerator::subdivided_hyper_rectangle(tria, {100, 10}, dealii::Point<2>(0, 0), dealii::Point<2>(2, 0.2), true);
      else
        dealii::GridGenerator::subdivided_hyper_rectangle(tria, {static_cast<unsigned int>(2 / 0.1), static_cast<unsigned int>(0.4 / 0.025)},
                                                          dealii::Point<2>(0, 0), dealii::Point<2>(2, 0.4), true);
      Fluid::MPI::SUPGInsIM<2> flow(tria, params);
      flow.run();
      auto solution = flow.get_current_solution();
      auto v = solution.block(0);
      if (which == "pressure")
        {
          dealii::Vector<double> serialized_v(v);
          std::sort(serialized_v.begin(), serialized_v.end(), std::greater<double>());
          const double vmax = serialized_v[0], vmax_30th = serialized_v[29];
          std::cout << "vmax = " << vmax << " 30th = " << vmax_30th << std::endl;
          if (!(std::abs(vmax - 2.5e-2) / 2.5e-2 < 2e-2)) throw std::runtime_error("Maximum velocity is incorrect!");
          if (!(std::abs(vmax_30th - 2.5e-2) / 2.5e-2 < 1e-3)) throw std::runtime_error("Maximum velocity is incorrect!");
        }
      else
        {
          const double l2_norm = v.l2_norm();
          std::cout << "l2 norm = " << l2_norm << std::endl;
          if (!(std::abs(l2_norm - 4.7112) / 4.7112 < 1e-3)) throw std::runtime_error("The l2 norm of velocity is incorrect!");
        }
    }
  catch (std::exception &exc)
    {
      std::cerr << "Exception on processing: " << exc.what() << std::endl;
      return 1;
    }
  return 0;
}
