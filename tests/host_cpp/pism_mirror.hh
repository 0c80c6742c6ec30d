// pism_mirror.hh -- PETSc-free mirrors of the reference containers SIAFD touches, so that the host class
// SIAFD_B200 (SIAFD_B200.hh) compiles and is tested without PISM's toolchain (MPI, PETSc, NetCDF, UDUNITS).
//
// Same names, accessors and memory layout as the reference (juliusgarbe/pism v1.2.1):
//   pism::IceGrid                      src/util/IceGrid.hh:170-394
//   pism::IceModelVec2S/2Int/2CellType/2Stag/2V/3   src/util/iceModelVec.hh:370-618
//     storage = the DMDA local (ghosted) array [j][i][dof], dof fastest  src/util/IceModelVec_inline.hh:28-40
//   pism::Geometry                     src/geometry/Geometry.hh, Geometry.cc:30-42
//   pism::stressbalance::Inputs        src/stressbalance/StressBalance.hh:41-65
//   pism::RuntimeError                 src/util/error_handling.hh:47-68
//   pism::Config (only the parameters on the SIAFD path, defaults of src/pism_config.cdl)
// Under real PISM these headers are NOT used: INTEGRATION.md shows the same class over PISM's own types.
// TEST SCAFFOLDING (it lives with the tests): one rank: update_ghosts() is the periodic self-wrap a one-process DMDA
// performs (IceGrid.cc:870-872); several ranks (one process each, PISM's decomposition IceGrid.cc:443-499): the driver
// installs a ghost exchanger on the grid, which stands for DMLocalToLocal over MPI.
#pragma once
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <functional>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace pism {

class RuntimeError : public std::runtime_error {
public:
  explicit RuntimeError(const std::string &message, int status = -1) : std::runtime_error(message), m_status(status) {}
  static RuntimeError formatted(int status, const char *format, ...) {
    char buffer[2048];
    va_list argp;
    va_start(argp, format);
    vsnprintf(buffer, sizeof(buffer), format, argp);
    va_end(argp);
    return RuntimeError(buffer, status);
  }
  int status() const { return m_status; } // the C ABI status code that raised it (siafd_b200.h)
private:
  int m_status;
};

// Config: name -> value with the reference's parameter names and defaults (src/pism_config.cdl; line numbers
// in SURVEY.md section 5.6).
class Config {
public:
  typedef std::shared_ptr<Config> Ptr;
  Config() {
    const double secpera = 365.242198781 * 86400.0; // UDUNITS-2 year
    m_num = {{"grid.max_stencil_width", 2},
             {"constants.ice.density", 910.0},
             {"constants.ice.grain_size", 1.0e-3},
             {"constants.ice.beta_Clausius_Clapeyron", 7.9e-8},
             {"constants.ice.specific_heat_capacity", 2009.0},
             {"constants.standard_gravity", 9.81},
             {"constants.ideal_gas_constant", 8.31441},
             {"constants.fresh_water.melting_point_temperature", 273.15},
             {"constants.fresh_water.latent_heat_of_fusion", 3.34e5},
             {"constants.fresh_water.specific_heat_capacity", 4170.0},
             {"constants.sea_water.density", 1028.0},
             {"surface.pressure", 0.0},
             {"enthalpy_converter.T_reference", 223.15},
             {"flow_law.Paterson_Budd.A_cold", 3.61e-13},
             {"flow_law.Paterson_Budd.A_warm", 1.73e3},
             {"flow_law.Paterson_Budd.Q_cold", 6.0e4},
             {"flow_law.Paterson_Budd.Q_warm", 13.9e4},
             {"flow_law.Paterson_Budd.T_critical", 263.15},
             {"flow_law.gpbld.water_frac_coeff", 181.25},
             {"flow_law.gpbld.water_frac_observed_limit", 0.01},
             {"flow_law.isothermal_Glen.ice_softness", 3.1689e-24},
             {"flow_law.Hooke.Q", 7.88e4},
             {"flow_law.Hooke.A", 4.42165e-9},
             {"flow_law.Hooke.C", 0.16612},
             {"flow_law.Hooke.k", 1.17},
             {"flow_law.Hooke.Tr", 273.39},
             {"stress_balance.sia.Glen_exponent", 3.0},
             {"stress_balance.sia.enhancement_factor", 1.0},
             {"stress_balance.sia.enhancement_factor_interglacial", 1.0},
             {"stress_balance.sia.max_diffusivity", 100.0},
             {"stress_balance.sia.bed_smoother.range", 5.0e3},
             {"stress_balance.sia.bed_smoother.theta_min", 0.0},
             {"time.eemian_start", -132000.0 * secpera},
             {"time.eemian_end", -114500.0 * secpera},
             {"time.holocene_start", -11000.0 * secpera},
             {"geometry.ice_free_thickness_standard", 0.01},
             {"stress_balance.ssa.Glen_exponent", 3.0},
             {"stress_balance.ssa.enhancement_factor", 1.0},
             {"time_stepping.maximum_time_step", 60.0}, // years (pism_config.cdl:2626)
             {"time_stepping.adaptive_ratio", 0.12}};
    m_str = {{"stress_balance.sia.flow_law", "gpbld"},
             {"stress_balance.sia.surface_gradient_method", "haseloff"},
             {"stress_balance.ssa.flow_law", "gpbld"},
             {"stress_balance.vertical_velocity_approximation", "centered"}};
    m_flag = {{"stress_balance.sia.limit_diffusivity", false},
              {"stress_balance.sia.grain_size_age_coupling", false},
              {"stress_balance.sia.e_age_coupling", false},
              {"ocean.always_grounded", false},
              {"geometry.update.use_basal_melt_rate", true},
              {"enthalpy_converter.cold_mode", false}}; // ColdEnthalpyConverter (EnthalpyConverter.cc:287-296)
  }
  double get_number(const std::string &name) const { return find(m_num, name); }
  std::string get_string(const std::string &name) const { return find(m_str, name); }
  bool get_flag(const std::string &name) const { return find(m_flag, name); }
  void set_number(const std::string &name, double v) { m_num[name] = v; }
  void set_string(const std::string &name, const std::string &v) { m_str[name] = v; }
  void set_flag(const std::string &name, bool v) { m_flag[name] = v; }

private:
  template <class M> static typename M::mapped_type find(const M &m, const std::string &name) {
    typename M::const_iterator it = m.find(name);
    if (it == m.end()) {
      throw RuntimeError::formatted(-1, "parameter '%s' is unset", name.c_str());
    }
    return it->second;
  }
  std::map<std::string, double> m_num;
  std::map<std::string, std::string> m_str;
  std::map<std::string, bool> m_flag;
};

enum SpacingType { EQUAL, QUADRATIC };

class IceGrid {
public:
  typedef std::shared_ptr<IceGrid> Ptr;
  typedef std::shared_ptr<const IceGrid> ConstPtr;

  // IceGrid::compute_vertical_levels, src/util/IceGrid.cc:381-425
  static std::vector<double> compute_vertical_levels(double Lz, unsigned int Mz, SpacingType spacing, double lambda = 4.0) {
    std::vector<double> z(Mz);
    if (spacing == EQUAL) {
      const double dz = Lz / ((double)Mz - 1);
      for (unsigned int k = 0; k < Mz - 1; k++) z[k] = dz * ((double)k);
    } else {
      for (unsigned int k = 0; k < Mz - 1; k++) {
        const double zeta = ((double)k) / ((double)Mz - 1);
        z[k] = Lz * ((zeta / lambda) * (1.0 + (lambda - 1.0) * zeta));
      }
    }
    z[Mz - 1] = Lz;
    return z;
  }

  // cell-corner registered domain, dx = 2 Lx / (Mx - 1) (IceGrid.cc:545-560), this rank's patch of PISM's DMDA
  // decomposition: processor grid from compute_nprocs (IceGrid.cc:443-484), ownership ranges :489-499, rank = px + Nx py
  IceGrid(Config::Ptr config, unsigned int Mx, unsigned int My, double Lx, double Ly, const std::vector<double> &z,
          int rank = 0, int size = 1)
      : m_config(config), m_Mx(Mx), m_My(My), m_Lx(Lx), m_Ly(Ly), m_z(z), m_time(0.0), m_rank(rank), m_size(size) {
    m_dx = 2.0 * Lx / (Mx - 1);
    m_dy = 2.0 * Ly / (My - 1);
    m_x.resize(Mx);
    m_y.resize(My);
    for (unsigned int i = 0; i < Mx; ++i) m_x[i] = -Lx + i * m_dx;
    for (unsigned int j = 0; j < My; ++j) m_y[j] = -Ly + j * m_dy;
    m_x[Mx - 1] = Lx;
    m_y[My - 1] = Ly;
    unsigned int Nx = 1, Ny = 1;
    compute_nprocs(Mx, My, (unsigned int)size, Nx, Ny);
    const std::vector<unsigned int> px = ownership_ranges(Mx, Nx), py = ownership_ranges(My, Ny);
    const unsigned int ix = (unsigned int)rank % Nx, iy = (unsigned int)rank / Nx;
    m_xs = m_ys = 0;
    for (unsigned int k = 0; k < ix; ++k) m_xs += (int)px[k];
    for (unsigned int k = 0; k < iy; ++k) m_ys += (int)py[k];
    m_xm = (int)px[ix], m_ym = (int)py[iy];
  }
  // IceGrid.cc:443-484
  static void compute_nprocs(unsigned int Mx, unsigned int My, unsigned int size, unsigned int &Nx, unsigned int &Ny) {
    if (My <= 0) throw RuntimeError::formatted(-1, "'My' is invalid.");
    Nx = (unsigned int)(0.5 + sqrt(((double)Mx) * ((double)size) / ((double)My)));
    Ny = 0;
    if (Nx == 0) Nx = 1;
    while (Nx > 0) {
      Ny = size / Nx;
      if (Nx * Ny == size) break;
      Nx--;
    }
    if (Mx > My and Nx < Ny) { // Swap Nx and Ny
      unsigned int tmp = Nx;
      Nx = Ny;
      Ny = tmp;
    }
    if ((Mx / Nx) < 2) throw RuntimeError::formatted(-1, "Can't split %d grid points into %d parts (X-direction).", Mx, (int)Nx);
    if ((My / Ny) < 2) throw RuntimeError::formatted(-1, "Can't split %d grid points into %d parts (Y-direction).", My, (int)Ny);
  }
  // IceGrid.cc:489-499
  static std::vector<unsigned int> ownership_ranges(unsigned int Mx, unsigned int Nx) {
    std::vector<unsigned int> result(Nx);
    for (unsigned int i = 0; i < Nx; i++) result[i] = Mx / Nx + ((Mx % Nx) > i);
    return result;
  }
  Config::Ptr config() const { return m_config; }
  unsigned int Mx() const { return m_Mx; }
  unsigned int My() const { return m_My; }
  unsigned int Mz() const { return (unsigned int)m_z.size(); }
  int xs() const { return m_xs; }
  int ys() const { return m_ys; }
  int xm() const { return m_xm; }
  int ym() const { return m_ym; }
  int rank() const { return m_rank; }
  int size() const { return m_size; }
  // DMLocalToLocal between ranks (util/iceModelVec.cc:630-643): installed by the driver of a multi-process run;
  // gets the vector's name, local array, ghost width and dof
  typedef std::function<void(const std::string &, double *, int, unsigned int)> GhostExchanger;
  void set_ghost_exchanger(GhostExchanger f) const { m_exchanger = f; }
  const GhostExchanger &ghost_exchanger() const { return m_exchanger; }
  double dx() const { return m_dx; }
  double dy() const { return m_dy; }
  double Lx() const { return m_Lx; }
  double Ly() const { return m_Ly; }
  double Lz() const { return m_z.back(); }
  double x(int i) const { return m_x[i]; }
  double y(int j) const { return m_y[j]; }
  double z(int k) const { return m_z[k]; }
  const std::vector<double> &z() const { return m_z; }
  double current_time() const { return m_time; } // grid->ctx()->time()->current(), SIAFD.cc:564
  void set_current_time(double t) { m_time = t; }

private:
  Config::Ptr m_config;
  unsigned int m_Mx, m_My;
  double m_Lx, m_Ly, m_dx, m_dy;
  std::vector<double> m_x, m_y, m_z;
  double m_time;
  int m_rank, m_size, m_xs, m_xm, m_ys, m_ym;
  mutable GhostExchanger m_exchanger;
};

inline double radius(const IceGrid &grid, int i, int j) { return sqrt(grid.x(i) * grid.x(i) + grid.y(j) * grid.y(j)); }

enum IceModelVecKind { WITHOUT_GHOSTS = 0, WITH_GHOSTS = 1 };

// Base: a ghosted local array [j][i][dof] (util/iceModelVec.cc:85-140).
class IceModelVec {
public:
  IceModelVec() : m_dof(1), m_width(0) {}
  virtual ~IceModelVec() {}
  void create(IceGrid::ConstPtr grid, const std::string &name, IceModelVecKind ghostedp, unsigned int dof, int width) {
    m_grid = grid, m_name = name, m_dof = dof, m_width = (ghostedp == WITH_GHOSTS) ? width : 0;
    m_data.assign((size_t)(grid->xm() + 2 * m_width) * (grid->ym() + 2 * m_width) * dof, 0.0);
  }
  IceGrid::ConstPtr grid() const { return m_grid; }
  const std::string &get_name() const { return m_name; }
  unsigned int ndof() const { return m_dof; }
  unsigned int stencil_width() const { return (unsigned int)m_width; }
  double *get_array() { return m_data.data(); } // the local array PISM's get_array() exposes
  const double *get_array() const { return m_data.data(); }
  size_t size() const { return m_data.size(); }
  void set(double c) { std::fill(m_data.begin(), m_data.end(), c); }
  void copy_from(const IceModelVec &other) {
    for (int j = m_grid->ys(); j < m_grid->ys() + m_grid->ym(); ++j)
      for (int i = m_grid->xs(); i < m_grid->xs() + m_grid->xm(); ++i)
        for (unsigned int d = 0; d < m_dof; ++d) at(i, j, d) = other.at(i, j, d);
    update_ghosts();
  }
  // DMLocalToLocal (util/iceModelVec.cc:630-643): the periodic self-wrap of a one-rank DMDA (IceGrid.cc:870-872), or
  // the exchanger the driver of a multi-process run installed on the grid
  void update_ghosts() {
    if (m_width == 0) return;
    if (m_grid->size() > 1) {
      if (!m_grid->ghost_exchanger()) throw RuntimeError::formatted(-1, "no ghost exchanger installed on a decomposed grid");
      m_grid->ghost_exchanger()(m_name, m_data.data(), m_width, m_dof);
      return;
    }
    const int Mx = m_grid->xm(), My = m_grid->ym(), w = m_width;
    for (int j = -w; j < My + w; ++j)
      for (int i = -w; i < Mx + w; ++i) {
        if (i >= 0 && i < Mx && j >= 0 && j < My) continue;
        const int is = ((i % Mx) + Mx) % Mx, js = ((j % My) + My) % My;
        for (unsigned int d = 0; d < m_dof; ++d) at(i, j, d) = at(is, js, d);
      }
  }
  double &at(int i, int j, unsigned int d) { return m_data[index(i, j) * m_dof + d]; }
  const double &at(int i, int j, unsigned int d) const { return m_data[index(i, j) * m_dof + d]; }

protected:
  // (i, j) are GLOBAL grid indices, valid on [xs - w, xs + xm + w) x [ys - w, ys + ym + w), as in PISM
  size_t index(int i, int j) const {
    return (size_t)(j - m_grid->ys() + m_width) * (m_grid->xm() + 2 * m_width) + (i - m_grid->xs() + m_width);
  }
  IceGrid::ConstPtr m_grid;
  std::string m_name;
  unsigned int m_dof;
  int m_width;
  std::vector<double> m_data;
};

class IceModelVec2S : public IceModelVec {
public:
  IceModelVec2S() {}
  IceModelVec2S(IceGrid::ConstPtr grid, const std::string &name, IceModelVecKind ghostedp, int width = 1) {
    create(grid, name, ghostedp, 1, width);
  }
  double &operator()(int i, int j) { return at(i, j, 0); }
  const double &operator()(int i, int j) const { return at(i, j, 0); }
};

// integers stored as doubles, decoded by floor(x + 0.5) (util/IceModelVec_inline.hh:95-101)
class IceModelVec2Int : public IceModelVec2S {
public:
  IceModelVec2Int() {}
  IceModelVec2Int(IceGrid::ConstPtr grid, const std::string &name, IceModelVecKind ghostedp, int width = 1)
      : IceModelVec2S(grid, name, ghostedp, width) {}
  int as_int(int i, int j) const { return (int)floor(at(i, j, 0) + 0.5); }
};
enum MaskValue { MASK_ICE_FREE_BEDROCK = 0, MASK_GROUNDED = 2, MASK_FLOATING = 3, MASK_ICE_FREE_OCEAN = 4 }; // util/Mask.hh:29-35
typedef IceModelVec2Int IceModelVec2CellType;

class IceModelVec2Stag : public IceModelVec {
public:
  IceModelVec2Stag() {}
  IceModelVec2Stag(IceGrid::ConstPtr grid, const std::string &name, IceModelVecKind ghostedp, int width = 1) {
    create(grid, name, ghostedp, 2, width);
  }
  double &operator()(int i, int j, int o) { return at(i, j, (unsigned int)o); }
  const double &operator()(int i, int j, int o) const { return at(i, j, (unsigned int)o); }
};

struct Vector2 {
  double u, v;
};
class IceModelVec2V : public IceModelVec {
public:
  IceModelVec2V() {}
  IceModelVec2V(IceGrid::ConstPtr grid, const std::string &name, IceModelVecKind ghostedp, int width = 1) {
    create(grid, name, ghostedp, 2, width);
  }
  Vector2 &operator()(int i, int j) { return *reinterpret_cast<Vector2 *>(&at(i, j, 0)); }
  const Vector2 &operator()(int i, int j) const { return *reinterpret_cast<const Vector2 *>(&at(i, j, 0)); }
};

class IceModelVec3 : public IceModelVec {
public:
  IceModelVec3() {}
  IceModelVec3(IceGrid::ConstPtr grid, const std::string &name, IceModelVecKind ghostedp, int width = 1) {
    create(grid, name, ghostedp, grid->Mz(), width);
  }
  double *get_column(int i, int j) { return &at(i, j, 0); }
  const double *get_column(int i, int j) const { return &at(i, j, 0); }
  void set_column(int i, int j, double c) {
    for (unsigned int k = 0; k < m_dof; ++k) at(i, j, k) = c;
  }
  void set_column(int i, int j, const double *values) {
    for (unsigned int k = 0; k < m_dof; ++k) at(i, j, k) = values[k];
  }
  // linear interpolation in z, util/iceModelVec3.cc:149-175 (kBelowHeight: largest k in [0, Mz-2] with z[k] <= height)
  double getValZ(int i, int j, double height) const {
    const std::vector<double> &z = m_grid->z();
    const double *column = get_column(i, j);
    if (height >= z.back()) return column[z.size() - 1];
    if (height <= z.front()) return column[0];
    size_t mcurr = 0;
    while (mcurr + 2 < z.size() && z[mcurr + 1] <= height) mcurr++;
    const double incr = (height - z[mcurr]) / (z[mcurr + 1] - z[mcurr]);
    return column[mcurr] + incr * (column[mcurr + 1] - column[mcurr]);
  }
};

// Geometry.cc:30-42: the five 2D fields with ghost width grid.max_stencil_width
class Geometry {
public:
  explicit Geometry(IceGrid::ConstPtr grid)
      : bed_elevation(grid, "topg", WITH_GHOSTS, w(grid)), sea_level_elevation(grid, "sea_level", WITH_GHOSTS, w(grid)),
        ice_thickness(grid, "thk", WITH_GHOSTS, w(grid)), ice_surface_elevation(grid, "usurf", WITH_GHOSTS, w(grid)),
        cell_type(grid, "mask", WITH_GHOSTS, w(grid)) {}
  IceModelVec2S bed_elevation, sea_level_elevation, ice_thickness, ice_surface_elevation;
  IceModelVec2CellType cell_type;

  // Geometry::ensure_consistency (geometry/Geometry.cc:121-187) with GeometryCalculator::compute (util/Mask.hh:96-133);
  // under PISM this is PISM's own host code -- the device version is siafd_b200_ensure_consistency
  void ensure_consistency(double ice_free_thickness_threshold) {
    IceGrid::ConstPtr grid = ice_thickness.grid();
    const Config &config = *grid->config();
    const double alpha = 1 - config.get_number("constants.ice.density") / config.get_number("constants.sea_water.density");
    const bool is_dry_simulation = config.get_flag("ocean.always_grounded");
    for (int j = grid->ys(); j < grid->ys() + grid->ym(); ++j) {
      for (int i = grid->xs(); i < grid->xs() + grid->xm(); ++i) {
        const double thickness = ice_thickness(i, j);
        if (thickness < 0.0) {
          throw RuntimeError::formatted(1, "Thickness is negative at point i=%d, j=%d", i, j);
        }
        const double hgrounded = bed_elevation(i, j) + thickness, hfloating = sea_level_elevation(i, j) + alpha * thickness;
        const bool is_floating = (hfloating > hgrounded), ice_free = (thickness <= ice_free_thickness_threshold);
        if (is_floating && (not is_dry_simulation)) {
          ice_surface_elevation(i, j) = hfloating;
          cell_type(i, j) = ice_free ? MASK_ICE_FREE_OCEAN : MASK_FLOATING;
        } else {
          ice_surface_elevation(i, j) = hgrounded;
          cell_type(i, j) = ice_free ? MASK_ICE_FREE_BEDROCK : MASK_GROUNDED;
        }
      }
    }
    ice_thickness.update_ghosts();
    cell_type.update_ghosts();
    ice_surface_elevation.update_ghosts();
  }

private:
  static int w(IceGrid::ConstPtr grid) { return (int)grid->config()->get_number("grid.max_stencil_width"); }
};

namespace stressbalance {
// StressBalance.hh:41-65 (only the members SIAFD reads)
class Inputs {
public:
  Inputs() : geometry(NULL), new_bed_elevation(true), basal_melt_rate(NULL), enthalpy(NULL), age(NULL) {} // StressBalance.cc:36-46
  const Geometry *geometry;
  bool new_bed_elevation;
  const IceModelVec2S *basal_melt_rate; // WITHOUT_GHOSTS; may be NULL
  const IceModelVec3 *enthalpy;
  const IceModelVec3 *age;
};
} // namespace stressbalance
} // namespace pism
