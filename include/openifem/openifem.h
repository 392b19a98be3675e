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
#include <cmath>
#include <functional>
#include <memory>
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

  // Tensor<1, dim> as the drivers use it: brace-initialised from a list, indexed, norm()
  template <int rank, int dim>
  class Tensor
  {
    static_assert(rank == 1, "only Tensor<1, dim> is provided");

  public:
    Tensor() { x.fill(0.0); }
    Tensor(std::initializer_list<double> v)
    {
      x.fill(0.0);
      unsigned i = 0;
      for (double a : v)
        if (i < (unsigned)dim) x[i++] = a;
    }
    double operator[](unsigned i) const { return x[i]; }
    double &operator[](unsigned i) { return x[i]; }
    const double *data() const { return x.data(); }

  private:
    std::array<double, dim> x;
  };

  // Vector<double>(solid.get_current_solution()) in the FSI drivers
  template <typename T>
  using Vector = std::vector<T>;

  template <int dim>
  class Triangulation
  {
  public:
    Triangulation() { openifem_detail::check(ifem_tria_create(dim, &h)); }
    explicit Triangulation(int /*mpi_communicator*/) : Triangulation() {}
    ~Triangulation() { ifem_tria_destroy(h); }
    Triangulation(const Triangulation &) = delete;
    void refine_global(unsigned times)
    {
      openifem_detail::check(ifem_tria_refine_global(h, (int)times));
      flags.clear();
    }
    unsigned n_active_cells() const
    {
      int64_t c = 0;
      openifem_detail::check(ifem_tria_counts(h, nullptr, &c, nullptr));
      return (unsigned)c;
    }
    ifem_tria *handle() const { return h; }

    // What the reference's drivers do with cell iterators before handing the mesh to a solver
    // (tests/fsi_leaflet_mpi/fsi_leaflet_mpi.cpp:66-76, tests/fsi-wall-3D/fsi-wall-3D.cpp:47-53):
    //   for (auto cell : tria.active_cell_iterators()) if (cell->center()[0] ...) cell->set_refine_flag();
    //   tria.execute_coarsening_and_refinement();
    class CellAccessor
    {
    public:
      CellAccessor(Triangulation *t, unsigned i) : tria(t), index(i) {}
      Point<dim> center() const
      {
        Point<dim> c;
        const int nv = 1 << dim;
        for (int v = 0; v < nv; ++v)
          for (int d = 0; d < dim; ++d) c[d] += tria->vertices[(size_t)tria->cells[(size_t)index * nv + v] * dim + d] / nv;
        return c;
      }
      bool is_locally_owned() const { return true; } // every rank flags the whole (replicated) coarse mesh
      void set_refine_flag() const { tria->flags[index] = 1; }
      void clear_refine_flag() const { tria->flags[index] = 0; }
      unsigned active_cell_index() const { return index; }
      const CellAccessor *operator->() const { return this; }

    private:
      Triangulation *tria;
      unsigned index;
    };
    class CellIterator
    {
    public:
      CellIterator(Triangulation *t, unsigned i) : acc(t, i), tria(t), index(i) {}
      const CellAccessor &operator*() const { return acc; }
      const CellAccessor *operator->() const { return &acc; }
      CellIterator &operator++()
      {
        acc = CellAccessor(tria, ++index);
        return *this;
      }
      bool operator!=(const CellIterator &o) const { return index != o.index; }

    private:
      CellAccessor acc;
      Triangulation *tria;
      unsigned index;
    };
    struct CellRange
    {
      CellIterator b, e;
      CellIterator begin() const { return b; }
      CellIterator end() const { return e; }
    };
    CellRange active_cell_iterators()
    {
      int64_t nv = 0, nc = 0, nb = 0;
      openifem_detail::check(ifem_tria_counts(h, &nv, &nc, &nb));
      vertices.resize((size_t)nv * dim);
      cells.resize((size_t)nc << dim);
      std::vector<int> bf((size_t)nb * 3);
      openifem_detail::check(ifem_tria_get_mesh(h, vertices.data(), cells.data(), bf.data()));
      if (flags.size() != (size_t)nc) flags.assign((size_t)nc, 0);
      return CellRange{CellIterator(this, 0), CellIterator(this, (unsigned)nc)};
    }
    void execute_coarsening_and_refinement()
    {
      if (flags.empty()) return;
      openifem_detail::check(ifem_tria_execute_refinement(h, flags.data(), (int64_t)flags.size()));
      flags.clear();
    }

  private:
    ifem_tria *h = nullptr;
    std::vector<double> vertices;
    std::vector<int> cells;
    std::vector<unsigned char> flags;
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

  namespace GridTools
  {
    template <int dim>
    void shift(const Tensor<1, dim> &offset, Triangulation<dim> &tria)
    {
      openifem_detail::check(ifem_tria_shift(tria.handle(), offset.data()));
    }
  } // namespace GridTools

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
using dealii::Point;
using dealii::Tensor;
using dealii::Triangulation;
using dealii::Vector;
namespace GridGenerator = dealii::GridGenerator;
namespace GridTools = dealii::GridTools;

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

// PETScWrappers::MPI::Vector as the drivers use one block of a solution: size, element access, norms, and the copy into a
// serial dealii::Vector<double> (= std::vector here; tests/fluid_pressure_driven_mpi_insim_supg/...cpp:41-46, fluid_plane_wall_driven...cpp:45-47)
namespace PETScWrappers
{
  namespace MPI
  {
    class Vector : public std::vector<double>
    {
    public:
      using std::vector<double>::vector;
      Vector() = default;
      Vector(std::vector<double> v) : std::vector<double>(std::move(v)) {}
      double l2_norm() const
      {
        double s = 0;
        for (double x : *this) s += x * x;
        return std::sqrt(s);
      }
      double linfty_norm() const
      {
        double s = 0;
        for (double x : *this) s = std::max(s, std::abs(x));
        return s;
      }
      double max() const { return *std::max_element(begin(), end()); }
      double min() const { return *std::min_element(begin(), end()); }
    };
  } // namespace MPI
} // namespace PETScWrappers

// PETScWrappers::MPI::BlockVector as returned by get_current_solution(): block(0) velocity, block(1) pressure
class BlockVector
{
public:
  BlockVector(std::vector<double> u, std::vector<double> p) : b{PETScWrappers::MPI::Vector(std::move(u)), PETScWrappers::MPI::Vector(std::move(p))} {}
  const PETScWrappers::MPI::Vector &block(unsigned i) const { return b[i]; }
  unsigned n_blocks() const { return 2; }

private:
  PETScWrappers::MPI::Vector b[2];
};

namespace Utils
{
  // Utils::GridCreator<dim>::flow_around_cylinder (include/utilities.h, source/utilities.cpp:343-574)
  template <int dim>
  struct GridCreator
  {
    static void flow_around_cylinder(dealii::Triangulation<dim> &tria) { openifem_detail::check(ifem_tria_flow_around_cylinder(tria.handle())); }
  };
  inline double PETScVectorMax(const std::vector<double> &v) { return *std::max_element(v.begin(), v.end()); }
  inline double PETScVectorMin(const std::vector<double> &v) { return *std::min_element(v.begin(), v.end()); }
} // namespace Utils

namespace Fluid
{
  namespace MPI
  {
    // common surface of the fluid solvers (include/mpi_fluid_solver.h:99-183)
    template <int dim>
    class FluidSolver
    {
    public:
      virtual ~FluidSolver() { ifem_insim_destroy(h); }
      FluidSolver(const FluidSolver &) = delete;
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
      // include/mpi_fluid_solver.h:104-108
      using BoundaryValue = std::function<double(const dealii::Point<dim> &, const unsigned int, const double)>;
      void add_hard_coded_boundary_condition(const int id, const BoundaryValue &f)
      {
        bcs.emplace_back(new BoundaryValue(f));
        openifem_detail::check(ifem_insim_add_hard_coded_boundary_condition(h, id, &bc_thunk, bcs.back().get()));
      }
      ifem_insim *handle() const { return h; }

    protected:
      FluidSolver() = default;
      ifem_insim *h = nullptr;

    private:
      static double bc_thunk(const double *p, unsigned int c, double time, void *user)
      {
        dealii::Point<dim> x;
        for (int d = 0; d < dim; ++d) x[d] = p[d];
        return (*static_cast<BoundaryValue *>(user))(x, c, time);
      }
      std::vector<std::unique_ptr<BoundaryValue>> bcs;
    };

    template <int dim>
    class InsIM : public FluidSolver<dim>
    {
    public:
      InsIM(dealii::Triangulation<dim> &tria, const Parameters::AllParameters &params)
      {
        openifem_detail::check(ifem_insim_create(tria.handle(), params.handle(), &this->h));
      }
    };

    // include/mpi_insimex.h: the implicit-explicit twin (same public surface as InsIM)
    template <int dim>
    class InsIMEX : public FluidSolver<dim>
    {
    public:
      InsIMEX(dealii::Triangulation<dim> &tria, const Parameters::AllParameters &params)
      {
        openifem_detail::check(ifem_insimex_create(tria.handle(), params.handle(), &this->h));
      }
    };

    // include/mpi_scnsim.h + the set_* hooks of include/mpi_fluid_solver.h:120-143
    template <int dim>
    class SCnsIM : public FluidSolver<dim>
    {
    public:
      using Field = std::function<double(const dealii::Point<dim> &, const unsigned int)>;
      SCnsIM(dealii::Triangulation<dim> &tria, const Parameters::AllParameters &params)
      {
        openifem_detail::check(ifem_scnsim_create(tria.handle(), params.handle(), &this->h));
      }
      void set_body_force(const Field &f) { openifem_detail::check(ifem_scnsim_set_body_force(this->h, &thunk, keep(f))); }
      void set_sigma_pml_field(const Field &f) { openifem_detail::check(ifem_scnsim_set_sigma_pml_field(this->h, &thunk, keep(f))); }
      void set_initial_condition(const Field &f) { openifem_detail::check(ifem_scnsim_set_initial_condition(this->h, &thunk, keep(f))); }
      // FluidSolver::attach_turbulence_model (include/mpi_fluid_solver.h:113): "Spalart-Allmaras"
      void attach_turbulence_model(const std::string &model_name)
      {
        openifem_detail::check(ifem_insim_attach_turbulence_model(this->h, model_name.c_str()));
      }
      // turbulence_model->get_eddy_viscosity() (include/mpi_turbulence_model.h:40) over the scalar support points of this rank
      std::vector<double> get_eddy_viscosity(std::size_t n_scalar_dofs)
      {
        std::vector<double> v(n_scalar_dofs);
        openifem_detail::check(ifem_turbulence_get_vector(this->h, 2, v.data()));
        return v;
      }

    private:
      static double thunk(const double *p, unsigned int c, void *user)
      {
        dealii::Point<dim> x;
        for (int d = 0; d < dim; ++d) x[d] = p[d];
        return (*static_cast<Field *>(user))(x, c);
      }
      void *keep(const Field &f)
      {
        fields.emplace_back(new Field(f));
        return fields.back().get();
      }
      std::vector<std::unique_ptr<Field>> fields;

    protected:
      SCnsIM() = default; // for solvers sharing the SUPGFluidSolver surface
    };

    // include/mpi_insim_supg.h: stabilised incompressible solver on the SUPGFluidSolver surface
    template <int dim>
    class SUPGInsIM : public SCnsIM<dim>
    {
    public:
      SUPGInsIM(dealii::Triangulation<dim> &tria, const Parameters::AllParameters &params)
      {
        openifem_detail::check(ifem_supg_insim_create(tria.handle(), params.handle(), &this->h));
      }
    };
  } // namespace MPI
} // namespace Fluid

namespace Solid
{
  namespace MPI
  {
    // include/mpi_solid_solver.h:75-79, include/mpi_shared_solid_solver.h:91-101: run(), get_current_solution()
    template <int dim>
    class SolidSolver
    {
    public:
      ~SolidSolver() { ifem_hyper_destroy(h); }
      SolidSolver(const SolidSolver &) = delete;
      void run() { openifem_detail::check(ifem_hyper_run(h)); }
      void run_one_step(bool first_step) { openifem_detail::check(ifem_hyper_run_one_step(h, first_step)); }
      std::vector<double> get_current_solution() const
      {
        int64_t n = 0;
        openifem_detail::check(ifem_hyper_sizes(h, &n, nullptr, nullptr, nullptr));
        std::vector<double> u((size_t)n);
        openifem_detail::check(ifem_hyper_get_current_solution(h, u.data()));
        return u;
      }
      ifem_hyper *handle() const { return h; }

    protected:
      SolidSolver() = default;
      ifem_hyper *h = nullptr;
    };
    template <int dim>
    using SharedSolidSolver = SolidSolver<dim>;

    // include/mpi_hyper_elasticity.h:96-98 (and the replicated twin include/mpi_shared_hyper_elasticity.h)
    template <int dim>
    class HyperElasticity : public SolidSolver<dim>
    {
    public:
      HyperElasticity(dealii::Triangulation<dim> &tria, const Parameters::AllParameters &params)
      {
        openifem_detail::check(ifem_hyper_create_twin(tria.handle(), params.handle(), 0, &this->h));
      }
    };
    // include/mpi_shared_hyper_elasticity.h: the replicated twin MPI::FSI takes
    template <int dim>
    class SharedHyperElasticity : public SolidSolver<dim>
    {
    public:
      SharedHyperElasticity(dealii::Triangulation<dim> &tria, const Parameters::AllParameters &params)
      {
        openifem_detail::check(ifem_hyper_create_twin(tria.handle(), params.handle(), 1, &this->h));
      }
    };

    // include/mpi_linear_elasticity.h
    template <int dim>
    class LinearElasticity : public SolidSolver<dim>
    {
    public:
      LinearElasticity(dealii::Triangulation<dim> &tria, const Parameters::AllParameters &params)
      {
        openifem_detail::check(ifem_linear_elasticity_create(tria.handle(), params.handle(), 0, &this->h));
      }
    };

    // include/mpi_shared_linear_elasticity.h
    template <int dim>
    class SharedLinearElasticity : public SolidSolver<dim>
    {
    public:
      SharedLinearElasticity(dealii::Triangulation<dim> &tria, const Parameters::AllParameters &params)
      {
        openifem_detail::check(ifem_linear_elasticity_create(tria.handle(), params.handle(), 1, &this->h));
      }
    };
  } // namespace MPI
} // namespace Solid

namespace MPI
{
  // include/mpi_fsi.h:39-47: FSI(fluid, solid, parameters, use_dirichlet_bc), run(), set_penetration_criterion()
  template <int dim>
  class FSI
  {
  public:
    FSI(Fluid::MPI::FluidSolver<dim> &f, Solid::MPI::SolidSolver<dim> &s, const Parameters::AllParameters &p, bool use_dirichlet_bc = false)
      : fluid(f), solid(s), params(p), dirichlet(use_dirichlet_bc)
    {
    }
    ~FSI()
    {
      if (h) ifem_fsi_destroy(h);
    }
    FSI(const FSI &) = delete;
    // source/mpi_fsi.cpp:1120-1227: refine both meshes, set up both solvers, then the coupled time loop
    void run()
    {
      ensure();
      openifem_detail::check(ifem_fsi_run(h));
    }
    void set_penetration_criterion(const std::function<double(const dealii::Point<dim> &)> &c, dealii::Tensor<1, dim> direction)
    {
      criterion = c;
      dir = direction;
      has_criterion = true;
    }
    void update_solid_box() { ensure(); openifem_detail::check(ifem_fsi_update_solid_box(h, nullptr)); }
    void update_indicator() { ensure(); openifem_detail::check(ifem_fsi_update_indicator(h)); }
    void find_fluid_bc() { ensure(); openifem_detail::check(ifem_fsi_find_fluid_bc(h)); }

  private:
    // the reference's FSI::run sets the solvers up itself (:1127-1143); the coupling object needs them set up, so it is
    // created on first use
    void ensure()
    {
      if (h) return;
      openifem_detail::check(ifem_hyper_setup_with_refinement(solid.handle()));
      openifem_detail::check(ifem_insim_setup_with_refinement(fluid.handle()));
      openifem_detail::check(ifem_fsi_create(fluid.handle(), solid.handle(), params.handle(), dirichlet, &h));
      if (has_criterion) openifem_detail::check(ifem_fsi_set_penetration_criterion(h, &thunk, this, dir.data()));
    }
    static double thunk(const double *p, void *user)
    {
      dealii::Point<dim> x;
      for (int d = 0; d < dim; ++d) x[d] = p[d];
      return static_cast<FSI *>(user)->criterion(x);
    }
    Fluid::MPI::FluidSolver<dim> &fluid;
    Solid::MPI::SolidSolver<dim> &solid;
    const Parameters::AllParameters &params;
    bool dirichlet, has_criterion = false;
    std::function<double(const dealii::Point<dim> &)> criterion;
    dealii::Tensor<1, dim> dir;
    ifem_fsi *h = nullptr;
  };
} // namespace MPI
