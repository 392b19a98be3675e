#include "parameters.h"

#include <cstdlib>
#include <fstream>
#include <sstream>
#include <stdexcept>

namespace Parameters
{
  namespace
  {
    using Table = std::map<std::pair<std::string, std::string>, std::string>;

    std::string trim(const std::string &s)
    {
      size_t b = s.find_first_not_of(" \t\r\n"), e = s.find_last_not_of(" \t\r\n");
      return b == std::string::npos ? std::string() : s.substr(b, e - b + 1);
    }
    // collapse runs of blanks inside a key ("Time  step size" == "Time step size")
    std::string squeeze(const std::string &s)
    {
      std::string out;
      bool blank = false;
      for (char c : trim(s))
        {
          if (c == ' ' || c == '\t')
            blank = true;
          else
            {
              if (blank && !out.empty()) out += ' ';
              blank = false;
              out += c;
            }
        }
      return out;
    }
    std::vector<std::string> split_list(const std::string &s)
    {
      std::vector<std::string> out;
      std::stringstream ss(s);
      std::string tok;
      while (std::getline(ss, tok, ','))
        {
          tok = trim(tok);
          if (!tok.empty()) out.push_back(tok);
        }
      return out;
    }
    double to_double(const std::string &s, const std::string &what)
    {
      char *end = nullptr;
      const double v = std::strtod(s.c_str(), &end);
      if (end == s.c_str() || !trim(end).empty()) throw std::runtime_error("Cannot parse '" + s + "' as a number for <" + what + ">");
      return v;
    }
    long to_int(const std::string &s, const std::string &what)
    {
      char *end = nullptr;
      const long v = std::strtol(s.c_str(), &end, 10);
      if (end == s.c_str() || !trim(end).empty()) throw std::runtime_error("Cannot parse '" + s + "' as an integer for <" + what + ">");
      return v;
    }
    std::vector<double> to_doubles(const std::string &s, const std::string &what)
    {
      std::vector<double> out;
      for (auto &t : split_list(s)) out.push_back(to_double(t, what));
      return out;
    }
    std::vector<int> to_ints(const std::string &s, const std::string &what)
    {
      std::vector<int> out;
      for (auto &t : split_list(s)) out.push_back((int)to_int(t, what));
      return out;
    }
    void require(bool cond, const char *msg)
    {
      if (!cond) throw std::runtime_error(msg);
    }

    // declared entries with the reference's defaults (parameters.cpp declareParameters)
    Table declared()
    {
      Table t;
      auto D = [&](const char *sec, const char *key, const char *def) { t[{sec, key}] = def; };
      D("Simulation", "Simulation type", "FSI");
      D("Simulation", "Dimension", "2");
      D("Simulation", "Global refinements", "");
      D("Simulation", "End time", "1.0");
      D("Simulation", "Time step size", "1.0");
      D("Simulation", "Output interval", "1.0");
      D("Simulation", "Refinement interval", "1.0");
      D("Simulation", "Save interval", "1.0");
      D("Simulation", "Gravity", "");
      D("Simulation", "Initial velocity", "");
      D("Fluid finite element system", "Pressure degree", "1");
      D("Fluid finite element system", "Velocity degree", "2");
      D("Fluid material properties", "Dynamic viscosity", "1e-3");
      D("Fluid material properties", "Fluid density", "1.0");
      D("Fluid solver control", "Grad-Div stabilization", "1.0");
      D("Fluid solver control", "Max Newton iterations", "8");
      D("Fluid solver control", "Nonlinear system tolerance", "1e-10");
      D("Fluid Dirichlet BCs", "Use hard-coded boundary values", "0");
      D("Fluid Dirichlet BCs", "Number of Dirichlet BCs", "0");
      D("Fluid Dirichlet BCs", "Dirichlet boundary id", "");
      D("Fluid Dirichlet BCs", "Dirichlet boundary components", "");
      D("Fluid Dirichlet BCs", "Dirichlet boundary values", "");
      D("Fluid Neumann BCs", "Number of Neumann BCs", "0");
      D("Fluid Neumann BCs", "Neumann boundary id", "");
      D("Fluid Neumann BCs", "Neumann boundary values", "");
      D("Spalart Allmaras model", "Number of S-A model BCs", "0");
      D("Spalart Allmaras model", "S-A model boundary id", "");
      D("Spalart Allmaras model", "S-A model boundary types", "");
      D("Spalart Allmaras model", "Initial condition coefficient", "0.0");
      D("Spalart Allmaras model", "Wall function effective distance", "0.0");
      D("Spalart Allmaras model", "Wall function image distance", "0.0");
      D("Solid finite element system", "Degree", "1");
      D("Solid material properties", "Solid type", "LinearElastic");
      D("Solid material properties", "Number of solid parts", "1");
      D("Solid material properties", "Solid density", "1.0");
      D("Solid material properties", "Young's modulus", "0.0");
      D("Solid material properties", "Poisson's ratio", "0.0");
      D("Solid material properties", "Viscosity", "0.0");
      D("Solid material properties", "Hyperelastic parameters", "");
      D("Solid solver control", "Damping", "0.0");
      D("Solid solver control", "Max Newton iterations", "8");
      D("Solid solver control", "Displacement tolerance", "1e-10");
      D("Solid solver control", "Force tolerance", "1e-10");
      D("Solid solver control", "Contact force multiplier", "1e8");
      D("Solid Dirichlet BCs", "Number of Dirichlet BCs", "0");
      D("Solid Dirichlet BCs", "Dirichlet boundary id", "");
      D("Solid Dirichlet BCs", "Dirichlet boundary components", "");
      D("Solid Neumann BCs", "Number of Neumann BCs", "0");
      D("Solid Neumann BCs", "Neumann boundary id", "");
      D("Solid Neumann BCs", "Neumann boundary type", "Traction");
      D("Solid Neumann BCs", "Neumann boundary values", "");
      return t;
    }

    Table read_prm(const std::string &text)
    {
      Table t = declared();
      std::vector<std::string> stack;
      std::stringstream ss(text);
      std::string raw;
      int lineno = 0;
      while (std::getline(ss, raw))
        {
          ++lineno;
          const size_t hash = raw.find('#');
          std::string line = trim(hash == std::string::npos ? raw : raw.substr(0, hash));
          if (line.empty()) continue;
          if (line.compare(0, 10, "subsection") == 0)
            stack.push_back(squeeze(line.substr(10)));
          else if (line == "end")
            {
              require(!stack.empty(), "prm: 'end' without matching 'subsection'");
              stack.pop_back();
            }
          else if (line.compare(0, 3, "set") == 0)
            {
              const size_t eq = line.find('=');
              require(eq != std::string::npos, "prm: 'set' without '='");
              std::string sec;
              for (size_t i = 0; i < stack.size(); ++i) sec += (i ? "/" : "") + stack[i];
              const std::string key = squeeze(line.substr(3, eq - 3));
              auto it = t.find({sec, key});
              if (it == t.end())
                throw std::runtime_error("prm line " + std::to_string(lineno) + ": no entry <" + key + "> declared in subsection <" + sec + ">");
              it->second = trim(line.substr(eq + 1));
            }
          else
            throw std::runtime_error("prm line " + std::to_string(lineno) + ": cannot parse '" + line + "'");
        }
      require(stack.empty(), "prm: unbalanced 'subsection' / 'end'");
      return t;
    }
  } // namespace

  AllParameters::AllParameters(const std::string &prm_file)
  {
    std::ifstream in(prm_file);
    if (!in) throw std::runtime_error("Cannot open parameter file " + prm_file);
    std::stringstream ss;
    ss << in.rdbuf();
    parse(ss.str());
  }

  AllParameters AllParameters::from_text(const std::string &text)
  {
    AllParameters p;
    p.parse(text);
    return p;
  }

  void AllParameters::parse(const std::string &text)
  {
    const Table t = read_prm(text);
    auto get = [&](const char *sec, const char *key) -> const std::string & { return t.at({sec, key}); };
    auto getd = [&](const char *sec, const char *key) { return to_double(get(sec, key), key); };
    auto geti = [&](const char *sec, const char *key) { return to_int(get(sec, key), key); };

    // Simulation (parameters.cpp:46-76)
    simulation_type = get("Simulation", "Simulation type");
    require(simulation_type == "FSI" || simulation_type == "Fluid" || simulation_type == "Solid",
            "Simulation type must be FSI|Fluid|Solid");
    dimension = (int)geti("Simulation", "Dimension");
    require(dimension >= 2, "Dimension must be >= 2");
    global_refinements = to_ints(get("Simulation", "Global refinements"), "Global refinements");
    require((int)global_refinements.size() == 2, "Incorrect dimension of global_refinements!");
    end_time = getd("Simulation", "End time");
    time_step = getd("Simulation", "Time step size");
    output_interval = getd("Simulation", "Output interval");
    refinement_interval = getd("Simulation", "Refinement interval");
    save_interval = getd("Simulation", "Save interval");
    gravity = to_doubles(get("Simulation", "Gravity"), "Gravity");
    require((int)gravity.size() == dimension, "Inconsistent dimension of gravity!");
    initial_velocity = to_doubles(get("Simulation", "Initial velocity"), "Initial velocity");
    require((int)initial_velocity.size() == dimension, "Inconsistent dimension of initial velocity!");

    fluid_pressure_degree = (unsigned)geti("Fluid finite element system", "Pressure degree");
    fluid_velocity_degree = (unsigned)geti("Fluid finite element system", "Velocity degree");
    viscosity = getd("Fluid material properties", "Dynamic viscosity");
    fluid_rho = getd("Fluid material properties", "Fluid density");
    grad_div = getd("Fluid solver control", "Grad-Div stabilization");
    fluid_max_iterations = (unsigned)geti("Fluid solver control", "Max Newton iterations");
    fluid_tolerance = getd("Fluid solver control", "Nonlinear system tolerance");

    // Fluid Dirichlet BCs (parameters.cpp:186-241)
    {
      const char *S = "Fluid Dirichlet BCs";
      use_hard_coded_values = (int)geti(S, "Use hard-coded boundary values");
      n_fluid_dirichlet_bcs = (unsigned)geti(S, "Number of Dirichlet BCs");
      const std::vector<int> ids = to_ints(get(S, "Dirichlet boundary id"), "Dirichlet boundary id");
      require(!n_fluid_dirichlet_bcs || ids.size() == n_fluid_dirichlet_bcs, "Inconsistent boundary ids!");
      const std::vector<int> comps = to_ints(get(S, "Dirichlet boundary components"), "Dirichlet boundary components");
      require(!n_fluid_dirichlet_bcs || comps.size() == n_fluid_dirichlet_bcs, "Inconsistent boundary components!");
      const std::vector<double> values = to_doubles(get(S, "Dirichlet boundary values"), "Dirichlet boundary values");
      unsigned n = 0;
      for (unsigned i = 0; i < n_fluid_dirichlet_bcs; ++i)
        {
          const int flag = comps[i];
          require(flag >= 1 && flag <= 7, "Dirichlet boundary components must be in [1,7]");
          require(n < values.size(), "Inconsistent boundary values!");
          const unsigned m = (flag == 1 || flag == 2 || flag == 4) ? 1 : (flag == 7 ? 3 : 2);
          require(n + m <= values.size(), "Inconsistent boundary values!");
          std::vector<double> value(values.begin() + n, values.begin() + n + m);
          n += m;
          fluid_dirichlet_bcs[ids[i]] = {(unsigned)flag, value};
        }
      require(n == values.size(), "Inconsistent boundary values!");
    }
    // Fluid Neumann BCs (:268-291)
    {
      const char *S = "Fluid Neumann BCs";
      n_fluid_neumann_bcs = (unsigned)geti(S, "Number of Neumann BCs");
      const std::vector<int> ids = to_ints(get(S, "Neumann boundary id"), "Neumann boundary id");
      require(!n_fluid_neumann_bcs || ids.size() == n_fluid_neumann_bcs, "Inconsistent boundary ids!");
      const std::vector<double> values = to_doubles(get(S, "Neumann boundary values"), "Neumann boundary values");
      require(!n_fluid_neumann_bcs || values.size() == n_fluid_neumann_bcs, "Inconsistent boundary values!");
      for (unsigned i = 0; i < n_fluid_neumann_bcs; ++i) fluid_neumann_bcs[ids[i]] = values[i];
    }
    // Spalart-Allmaras (:328-361)
    {
      const char *S = "Spalart Allmaras model";
      n_spalart_allmaras_model_bcs = (unsigned)geti(S, "Number of S-A model BCs");
      const std::vector<int> ids = to_ints(get(S, "S-A model boundary id"), "S-A model boundary id");
      require(!n_spalart_allmaras_model_bcs || ids.size() == n_spalart_allmaras_model_bcs, "Inconsistent boundary ids!");
      const std::vector<int> types = to_ints(get(S, "S-A model boundary types"), "S-A model boundary types");
      require(!n_spalart_allmaras_model_bcs || types.size() == n_spalart_allmaras_model_bcs, "Inconsistent boundary values!");
      for (unsigned i = 0; i < n_spalart_allmaras_model_bcs; ++i) spalart_allmaras_model_bcs[ids[i]] = types[i];
      spalart_allmaras_initial_condition_coefficient = getd(S, "Initial condition coefficient");
      spalart_allmaras_wall_function_distance = getd(S, "Wall function effective distance");
      spalart_allmaras_image_distance = getd(S, "Wall function image distance");
    }
    solid_degree = (unsigned)geti("Solid finite element system", "Degree");
    // Solid material (:421-464)
    {
      const char *S = "Solid material properties";
      solid_type = get(S, "Solid type");
      require(solid_type == "LinearElastic" || solid_type == "NeoHookean" || solid_type == "Kirchhoff",
              "Solid type must be LinearElastic|NeoHookean|Kirchhoff");
      n_solid_parts = (unsigned)geti(S, "Number of solid parts");
      require(n_solid_parts > 0, "Number of solid part less than 1!");
      solid_rho = getd(S, "Solid density");
      E = to_doubles(get(S, "Young's modulus"), "Young's modulus");
      require(E.size() == n_solid_parts, "Inconsistent Youngs' moduli!");
      nu = to_doubles(get(S, "Poisson's ratio"), "Poisson's ratio");
      require(nu.size() == n_solid_parts, "Inconsistent Poisson's ratios!");
      eta = to_doubles(get(S, "Viscosity"), "Viscosity");
      require(eta.size() == n_solid_parts, "Inconsistent viscosity!");
      const std::vector<double> c = to_doubles(get(S, "Hyperelastic parameters"), "Hyperelastic parameters");
      const unsigned per = solid_type == "NeoHookean" ? 2 : 1;
      C.assign(n_solid_parts, std::vector<double>(per, 0.0));
      if (solid_type == "NeoHookean" || !c.empty())
        {
          require(c.size() >= per * n_solid_parts, "Insufficient material properties input!");
          for (unsigned i = 0; i < n_solid_parts; ++i)
            for (unsigned j = 0; j < per; ++j) C[i][j] = c[i * per + j];
        }
    }
    {
      const char *S = "Solid solver control";
      damping = getd(S, "Damping");
      solid_max_iterations = (unsigned)geti(S, "Max Newton iterations");
      tol_d = getd(S, "Displacement tolerance");
      tol_f = getd(S, "Force tolerance");
      contact_force_multiplier = getd(S, "Contact force multiplier");
    }
    {
      const char *S = "Solid Dirichlet BCs";
      n_solid_dirichlet_bcs = (unsigned)geti(S, "Number of Dirichlet BCs");
      const std::vector<int> ids = to_ints(get(S, "Dirichlet boundary id"), "Dirichlet boundary id");
      require(!n_solid_dirichlet_bcs || ids.size() == n_solid_dirichlet_bcs, "Inconsistent boundary ids!");
      const std::vector<int> comps = to_ints(get(S, "Dirichlet boundary components"), "Dirichlet boundary components");
      require(!n_solid_dirichlet_bcs || comps.size() == n_solid_dirichlet_bcs, "Inconsistent boundary components!");
      for (unsigned i = 0; i < n_solid_dirichlet_bcs; ++i) solid_dirichlet_bcs[ids[i]] = comps[i];
    }
    {
      const char *S = "Solid Neumann BCs";
      solid_neumann_bc_dim = dimension;
      n_solid_neumann_bcs = (unsigned)geti(S, "Number of Neumann BCs");
      const std::vector<int> ids = to_ints(get(S, "Neumann boundary id"), "Neumann boundary id");
      require(!n_solid_neumann_bcs || ids.size() == n_solid_neumann_bcs, "Inconsistent boundary ids!");
      solid_neumann_bc_type = get(S, "Neumann boundary type");
      require(solid_neumann_bc_type == "Traction" || solid_neumann_bc_type == "Pressure", "Neumann boundary type must be Traction|Pressure");
      const unsigned per = solid_neumann_bc_type == "Traction" ? (unsigned)solid_neumann_bc_dim : 1u;
      const std::vector<double> values = to_doubles(get(S, "Neumann boundary values"), "Neumann boundary values");
      require(!n_solid_neumann_bcs || values.size() == per * n_solid_neumann_bcs, "Inconsistent boundary values!");
      for (unsigned i = 0; i < n_solid_neumann_bcs; ++i)
        solid_neumann_bcs[ids[i]] = std::vector<double>(values.begin() + i * per, values.begin() + (i + 1) * per);
    }
  }
} // namespace Parameters
