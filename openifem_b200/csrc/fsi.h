// Immersed coupling kernels of MPI::FSI<dim> (reference include/mpi_fsi.h:39-47, source/mpi_fsi.cpp):
// update_solid_box (:95-119), point_in_solid (:143-224), update_indicator (:292-319) and find_fluid_bc
// (:324-663), plus the point location / interpolation they lean on (Utils::CellLocator,
// Utils::GridInterpolator, source/utilities.cpp:193-341).
//
// The reference moves the solid mesh forward and back around every query loop (move_solid_mesh, :40-75)
// and searches all solid cells per 3-D query. Here the deformed vertex positions live in one device array
// that is refreshed once per coupling step, solid cells are binned on a uniform grid over the solid box,
// and every fluid vertex / support point is one thread: bin lookup -> bounding-box reject -> Newton inverse
// of the Q1 map -> unit-cell test. The 2-D crossing-number test is kept statement for statement (its
// on-edge / on-vertex rules are exact floating-point comparisons).
#pragma once
#include <functional>

#include "insim.h"
#include "solid.h"

namespace ifem
{
  class FsiCoupling
  {
  public:
    FsiCoupling(Context &ctx, InsIM &fluid, SolidSolver &solid, const Parameters::AllParameters &params, bool use_dirichlet_bc);

    // bounding box of the deformed solid, [min0, max0, min1, max1, ...] (host copy returned)
    std::vector<double> update_solid_box();
    // CellProperty::indicator of every local fluid cell (owned and ghost)
    void update_indicator();
    // fluid.fsi_acceleration (and, with use_dirichlet_bc, the inner constraints merged into the fluid's)
    void find_fluid_bc();

    // find_solid_bc (source/mpi_fsi.cpp:666-867): fluid stress (-p I + viscous), velocity and pressure interpolated
    // at the vertices of the solid's non-fixed boundary faces -> solid.fsi_stress_rows / fluid_velocity / fluid_pressure
    void find_solid_bc();
    // one coupled time step / the time loop of FSI::run (source/mpi_fsi.cpp:1172-1226), without refinement / checkpoints
    void run_one_step(bool first_step, bool stop_before_fluid_step = false);
    // FSI::refine_mesh(min_grid_level, max_grid_level) (source/mpi_fsi.cpp:1024-1117): fluid cells whose centre is closer to the
    // boundary of the (deformed) solid than their diameter are refined, the others coarsened, between the two levels; the fluid
    // solution is transferred and the fluid solver set up again
    void refine_mesh(unsigned int min_grid_level, unsigned int max_grid_level);
    void run();
    Time time;
    // FSI::set_penetration_criterion (source/mpi_fsi.cpp:1229-1237) / apply_contact_model (:869-970): while a boundary
    // vertex of the moved solid penetrates (criterion > 1e-5), the solid step is redone with an extra stress
    // Contact force multiplier * penetration along `direction` added to fsi_stress_rows at that vertex
    void set_penetration_criterion(std::function<double(const double *)> criterion, const double *direction);
    void apply_contact_model(bool first_step);
    int contact_iterations = 0; // solid steps taken inside apply_contact_model so far
    bool restarted = false;     // run() restored both solvers from checkpoints (mpi_fsi.cpp:1127-1151)

    // batch queries on the current deformed solid (tests / diagnostics)
    void point_in_solid(int n, const double *pts_host, int *inside_host);
    // which: 0 current_velocity, 1 current_acceleration, 2 current_displacement
    void interpolate(int which, int n, const double *pts_host, double *values_host, int *found_host);

    Context &ctx;
    InsIM &fluid;
    SolidSolver &solid;
    Parameters::AllParameters parameters;
    bool use_dirichlet_bc;
    std::vector<double> solid_box; // host mirror
    // inner constraints of the last find_fluid_bc (before the merge), per local fluid dof
    DevBuf<unsigned char> d_inner_con;
    DevBuf<double> d_inner_inhom;
    std::map<std::string, double> timer_ms;

  private:
    std::function<double(const double *)> penetration_criterion;
    double penetration_direction[3] = {0, 0, 0};
    void refresh_deformed();
    void build_bins();
    int dim;
    DevBuf<double> d_x;        // deformed solid vertices
    DevBuf<double> d_box;      // 2*dim
    DevBuf<int> d_bseg;        // 2-D: boundary segments as node pairs
    int n_bseg = 0;
    int nbin[3] = {1, 1, 1};
    DevBuf<int> d_bin_start, d_bin_cursor, d_bin_items;
    // fluid velocity node -> (cell, local index) adjacency, cells ascending
    DevBuf<int> d_n2c_ptr, d_n2c_cell, d_n2c_loc;
    DevBuf<unsigned char> d_cell_owned;    // 1 for locally owned fluid cells (is_locally_owned)
    DevBuf<unsigned char> d_node_interior; // 1 for cell-centre nodes of the Q2 element
    DevBuf<double> d_un_coords;            // support point of every local velocity node
    DevBuf<double> d_sp_tables;            // dN_u at the unit support points [nu][nu][dim] | dN_geo [nu][nv][dim]
    // uniform-grid bins of the (static) fluid cells, for locating solid vertices in the fluid mesh
    void build_fluid_bins();
    void build_fluid_side(); // everything the coupling derives from the fluid mesh / dofs (constructor, refine_mesh)
    int fbin[3] = {1, 1, 1};
    double fluid_box[6] = {0, 0, 0, 0, 0, 0};
    DevBuf<double> d_fluid_box;
    DevBuf<int> d_fbin_start, d_fbin_items;
    DevBuf<int> d_solid_bvert; // vertices (solid nodes) of the non-fixed boundary faces, unique
    int n_solid_bvert = 0;
    bool deformed_valid = false;
  };
} // namespace ifem
