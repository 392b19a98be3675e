// Parameters::AllParameters with the same member names, defaults and validation as
// the reference (include/parameters.h:14-194, source/parameters.cpp:7-658), read
// from the same deal.II ParameterHandler `.prm` text format ("subsection X / set K
// = V / end", '#' comments). deal.II's ParameterHandler is replaced by a small
// parser; unknown subsections / keys raise like ParameterHandler does.
#pragma once
#include <map>
#include <string>
#include <utility>
#include <vector>

namespace Parameters
{
  struct Simulation
  {
    std::string simulation_type;
    int dimension;
    std::vector<int> global_refinements;
    double end_time;
    double time_step;
    double output_interval;
    double refinement_interval;
    double save_interval;
    std::vector<double> gravity;
    std::vector<double> initial_velocity;
  };

  struct FluidFESystem
  {
    unsigned int fluid_pressure_degree;
    unsigned int fluid_velocity_degree;
  };

  struct FluidMaterial
  {
    double viscosity;
    double fluid_rho;
  };

  struct FluidSolver
  {
    double grad_div;
    unsigned int fluid_max_iterations;
    double fluid_tolerance;
  };

  struct FluidDirichlet
  {
    int use_hard_coded_values;
    unsigned int n_fluid_dirichlet_bcs;
    std::map<unsigned int, std::pair<unsigned int, std::vector<double>>> fluid_dirichlet_bcs;
  };

  struct FluidNeumann
  {
    unsigned int n_fluid_neumann_bcs;
    std::map<unsigned int, double> fluid_neumann_bcs;
  };

  struct SpalartAllmarasModel
  {
    unsigned int n_spalart_allmaras_model_bcs;
    std::map<unsigned int, unsigned int> spalart_allmaras_model_bcs;
    double spalart_allmaras_initial_condition_coefficient;
    double spalart_allmaras_wall_function_distance;
    double spalart_allmaras_image_distance;
  };

  struct SolidFESystem
  {
    unsigned int solid_degree;
  };

  struct SolidMaterial
  {
    std::string solid_type;
    unsigned int n_solid_parts;
    double solid_rho;
    std::vector<double> E;
    std::vector<double> nu;
    std::vector<double> eta;
    std::vector<std::vector<double>> C;
  };

  struct SolidSolver
  {
    double damping;
    unsigned int solid_max_iterations;
    double tol_f;
    double tol_d;
    double contact_force_multiplier;
  };

  struct SolidDirichlet
  {
    unsigned int n_solid_dirichlet_bcs;
    std::map<unsigned int, unsigned int> solid_dirichlet_bcs;
  };

  struct SolidNeumann
  {
    unsigned int n_solid_neumann_bcs;
    std::string solid_neumann_bc_type;
    std::map<unsigned int, std::vector<double>> solid_neumann_bcs;
    int solid_neumann_bc_dim;
  };

  struct AllParameters : public Simulation,
                         public FluidFESystem,
                         public FluidMaterial,
                         public FluidSolver,
                         public FluidDirichlet,
                         public FluidNeumann,
                         public SpalartAllmarasModel,
                         public SolidFESystem,
                         public SolidMaterial,
                         public SolidSolver,
                         public SolidDirichlet,
                         public SolidNeumann
  {
    explicit AllParameters(const std::string &prm_file);
    // parse from text already in memory (used by the C ABI)
    static AllParameters from_text(const std::string &text);

  private:
    AllParameters() = default;
    void parse(const std::string &text);
  };
} // namespace Parameters
