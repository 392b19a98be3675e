// C++ facade with the reference's class surface over the C ABI (include/openifem_b200.h).
// A driver written against OpenIFEM's headers keeps its body; only the includes change:
//
//   reference                                          here
//   ---------------------------------------------------------------------------------------
//   #include <deal.II/grid/grid_generator.h>            #include <openifem/openifem.h>
//   #include "mpi_insim.h", "parameters.h", ...         (same header)
//   parallel::distributed::Triangulation<dim> tria(c)   parallel::distributed::Triangulation<dim> tria(c)
//   dealii::GridGenerator::subdivided_hyper_rectangle   dealii::GridGenerator::subdivided_hyper_rectangle
//   Parameters::AllParameters params(infile)            Parameters::AllParameters params(infile)
//   Fluid::MPI::InsIM<dim> flow(tria, params)           Fluid::MPI::InsIM<dim> flow(tria, params)
//   flow.run(); flow.get_current_solution().block(0)    flow.run(); flow.get_current_solution().block(0)
//
// Reference declarations mirrored: include/mpi_fluid_solver.h:99-183, include/mpi_insim.h:42-87,
// include/parameters.h:191, include/utilities.h:219-224 (PETScVectorMax/Min).
// deal.II itself is not available, so the handful of deal.II names the test drivers touch
// (Point, Triangulation, GridGenerator, Utilities::MPI::MPI_InitFinalize) are provided as thin shims.
#pragma once
#include <openifem_b200.h>

#include <algorithm>
#include <array>
#include <stdexcept>
#include <string>
#include <vector>

namespace openifem_detail
{
  inline void check(int rc)
  {
    if (rc != 0) throw std::runtime_error(ifem_last_error());
  }
} // namespace openifem_detail

#define MPI_COMM_WORLD 0

namespace dealii
{
  template <int dim>
  class Point
  {
  public:
    Point() { x.fill(0.0); }
    Point(double a, double b) { static_assert(dim == 2, "dim"); x = {a, b}; }
    Point(double a, double b, double c) { static_assert(dim == 3, "dim"); x = {a, b, c}; }
    double operator[](unsigned i) const { return x[i]; }
    double &operator[](unsigned i) { return x[i]; }
    const double *data() const { return x.data(); }

  private:
    std::array<double, dim> x;
  };

  template <int dim>
  class Triangulation
  {
  public:
    Triangulation() { openifem_detail::check(ifem_tria_create(dim, &h)); }
    explicit Triangulation(int /*mpi_communicator*/) : Triangulation() {}
    ~Triangulation() { ifem_tria_destroy(h); }
    Triangulation(const Triangulation &) = delete;
    void refine_global(unsigned times) { openifem_detail::check(ifem_tria_refine_global(h, (int)times)); }
    unsigned n_active_cells() const
    {
      int64_t c = 0;
      openifem_detail::check(ifem_tria_counts(h, nullptr, &c, nullptr));
      return (unsigned)c;
    }
    ifem_tria *handle() const { return h; }

  private:
    ifem_tria *h = nullptr;
  };

  namespace parallel
  {
    namespace distributed
    {
      template <int dim>
      using Triangulation = dealii::Triangulation<dim>;
    }
  } // namespace parallel

  namespace GridGenerator
  {
    template <int dim>
    void subdivided_hyper_rectangle(Triangulation<dim> &tria, const std::vector<unsigned int> &repetitions, const Point<dim> &p1,
                                    const Point<dim> &p2, bool colorize = false)
    {
      openifem_detail::check(ifem_tria_subdivided_hyper_rectangle(tria.handle(), repetitions.data(), p1.data(), p2.data(), colorize));
    }
    template <int dim>
    void hyper_cube(Triangulation<dim> &tria, double left = 0.0, double right = 1.0, bool colorize = false)
    {
      openifem_detail::check(ifem_tria_hyper_cube(tria.handle(), left, right, colorize));
    }
  } // namespace GridGenerator

  namespace Utilities
  {
    namespace MPI
    {
      // one process per GPU: binds the device (rank / NCCL id distribution is the launcher's job)
      struct MPI_InitFinalize
      {
        MPI_InitFinalize(int &, char **&, unsigned = 1, int device = 0) { openifem_detail::check(ifem_init(device)); }
      };
    } // namespace MPI
  } // namespace Utilities
} // namespace dealii

namespace parallel = dealii::parallel;

namespace Parameters
{
  class AllParameters
  {
  public:
    explicit AllParameters(const std::string &prm_file) { openifem_detail::check(ifem_params_from_file(prm_file.c_str(), &h)); }
    ~AllParameters() { ifem_params_destroy(h); }
    AllParameters(const AllParameters &) = delete;
    const ifem_params *handle() const { return h; }

  private:
    ifem_params *h = nullptr;
  };
} // namespace Parameters

// PETScWrappers::MPI::BlockVector as returned by get_current_solution(): block(0) velocity, block(1) pressure
class BlockVector
{
public:
  BlockVector(std::vector<double> u, std::vector<double> p) : b{std::move(u), std::move(p)} {}
  const std::vector<double> &block(unsigned i) const { return b[i]; }
  unsigned n_blocks() const { return 2; }

private:
  std::vector<double> b[2];
};

namespace Utils
{
  inline double PETScVectorMax(const std::vector<double> &v) { return *std::max_element(v.begin(), v.end()); }
  inline double PETScVectorMin(const std::vector<double> &v) { return *std::min_element(v.begin(), v.end()); }
} // namespace Utils

namespace Fluid
{
  namespace MPI
  {
    template <int dim>
    class InsIM
    {
    public:
      InsIM(dealii::Triangulation<dim> &tria, const Parameters::AllParameters &params)
      {
        openifem_detail::check(ifem_insim_create(tria.handle(), params.handle(), &h));
      }
      ~InsIM() { ifem_insim_destroy(h); }
      InsIM(const InsIM &) = delete;
      void run() { openifem_detail::check(ifem_insim_run(h)); }
      void run_one_step(bool apply_nonzero_constraints, bool /*assemble_system*/ = true)
      {
        openifem_detail::check(ifem_insim_run_one_step(h, apply_nonzero_constraints));
      }
      BlockVector get_current_solution() const
      {
        int64_t n_u = 0, n_p = 0;
        openifem_detail::check(ifem_insim_sizes(h, &n_u, &n_p, nullptr, nullptr, nullptr));
        std::vector<double> all((size_t)(n_u + n_p));
        openifem_detail::check(ifem_insim_get_current_solution(h, all.data()));
        return BlockVector(std::vector<double>(all.begin(), all.begin() + n_u), std::vector<double>(all.begin() + n_u, all.end()));
      }
      ifem_insim *handle() const { return h; }

    private:
      ifem_insim *h = nullptr;
    };
  } // namespace MPI
} // namespace Fluid
