// SIAFD_B200.hh -- C++ host side of the B200 SIAFD: the reference's SSB_Modifier / SIAFD interface
// (src/stressbalance/SSB_Modifier.hh:39-72, src/stressbalance/sia/SIAFD.hh:50-131) over the C ABI of
// include/siafd_b200.h.  Same method names, argument meaning and error behaviour (pism::RuntimeError) as the
// reference, so that it drops in under StressBalance (StressBalance.cc:169-211) and GeometryEvolution.
// Nothing here computes physics: update() marshals raw pointers of the (PISM-owned, host) ghosted arrays to
// siafd_b200_update and converts the status code back into an exception.
//
// The host types (IceGrid, IceModelVec*, Geometry, Inputs, Config, RuntimeError) come from SIAFD_B200_HOST_TYPES: PISM's
// own headers in a PISM build ("The shim a PISM maintainer adds"), or the PETSc-free mirror the tests
// compile against (tests/host_cpp/pism_mirror.hh, on the include path of tests/host_cpp/Makefile).
//
// Decomposed runs (one process per GPU, PISM's DMDA patches): pass a rendezvous prefix to the constructor; the class
// then forms the library's communicator (siafd_b200_comm_init: no MPI / NCCL underneath), update() runs every ghost
// update of the reference between the GPUs, max_diffusivity() / high_diffusivity_counter() are global, and an error
// on any rank throws on every rank (ParallelSection, util/error_handling.cc:189-214).
#pragma once
#ifndef SIAFD_B200_HOST_TYPES
#define SIAFD_B200_HOST_TYPES "pism_mirror.hh"
#endif
#include SIAFD_B200_HOST_TYPES

#include "../../include/siafd_b200.h"

namespace pism {
namespace stressbalance {

// SSB_Modifier.hh:39-72: owned outputs and getters
class SSB_Modifier {
public:
  explicit SSB_Modifier(IceGrid::ConstPtr g)
      : m_grid(g), m_config(g->config()), m_D_max(0.0), m_diffusive_flux(g, "diffusive_flux", WITH_GHOSTS, 1),
        m_u(g, "uvel", WITH_GHOSTS, 1), m_v(g, "vvel", WITH_GHOSTS, 1), m_strain_heating(g, "strainheat", WITHOUT_GHOSTS) {
  } // SSB_Modifier.cc:30-58
  virtual ~SSB_Modifier() {}
  virtual void init() {}
  virtual void update(const IceModelVec2V &sliding_velocity, const Inputs &inputs, bool full_update) = 0;
  virtual const IceModelVec2Stag &diffusive_flux() { return m_diffusive_flux; }
  virtual double max_diffusivity() const { return m_D_max; }
  const IceModelVec3 &velocity_u() const { return m_u; }
  const IceModelVec3 &velocity_v() const { return m_v; }
  // SSB_Modifier.cc:81-87: the modifier's own strain heating field (SIAFD leaves it at zero: the SIA's strain heating is
  // StressBalance::compute_volumetric_strain_heating, StressBalance_B200 below) and an empty report
  const IceModelVec3 &volumetric_strain_heating() const { return m_strain_heating; }
  virtual std::string stdout_report() const { return ""; }

protected:
  IceGrid::ConstPtr m_grid;
  Config::Ptr m_config;
  double m_D_max;
  IceModelVec2Stag m_diffusive_flux;
  IceModelVec3 m_u, m_v, m_strain_heating;
};

} // namespace stressbalance

// rheology::FlowLaw as SSB_Modifier::flow_law() hands it out (rheology/FlowLaw.hh:70-116; used by
// IceCompModel::reportErrors, iceCompModel.cc:642): name, exponent, enhancement factor and flow(), which evaluates the
// library's own device implementation (siafd_b200_flow_host), so that host-side diagnostics see the same law
namespace rheology {
class FlowLaw_B200 {
public:
  FlowLaw_B200(siafd_b200_handle *h, const std::string &name, double n, double e) : m_h(h), m_name(name), m_n(n), m_e(e) {}
  std::string name() const { return m_name; }
  double exponent() const { return m_n; }
  double enhancement_factor() const { return m_e; }
  // FlowLaw::flow(stress, enthalpy, pressure, grainsize), FlowLaw.cc:97-105
  double flow(double stress, double enthalpy, double pressure, double grainsize) const {
    double result = 0.0;
    const int status = siafd_b200_flow_host(m_h, 1, &stress, &enthalpy, &pressure, &grainsize, &result);
    if (status != SIAFD_B200_OK) throw RuntimeError::formatted(status, "%s", siafd_b200_last_error(m_h));
    return result;
  }

private:
  siafd_b200_handle *m_h;
  std::string m_name;
  double m_n, m_e;
};
} // namespace rheology

namespace stressbalance {

// BedSmoother as SIAFD::bed_smoother() hands it out (sia/BedSmoother.hh:70-110): the smoothed bed and the fields of the
// last update, read back from the device on demand
class BedSmoother_B200 {
public:
  BedSmoother_B200(IceGrid::ConstPtr g, siafd_b200_handle *h, int width)
      : m_h(h), m_topgsmooth(g, "topgsmooth", WITH_GHOSTS, width), m_theta(g, "theta", WITH_GHOSTS, width),
        m_thk_smooth(g, "thksmooth", WITH_GHOSTS, width) {}
  const IceModelVec2S &smoothed_bed() const { return fetch(SIAFD_B200_F_TOPGSMOOTH, m_topgsmooth); } // BedSmoother.cc:155
  const IceModelVec2S &theta() const { return fetch(SIAFD_B200_F_THETA, m_theta); }                   // :351-404
  const IceModelVec2S &smoothed_thk() const { return fetch(SIAFD_B200_F_THK_SMOOTH, m_thk_smooth); }  // :284-327

private:
  const IceModelVec2S &fetch(int field, IceModelVec2S &v) const {
    const int status = siafd_b200_download(m_h, field, v.get_array());
    if (status != SIAFD_B200_OK) throw RuntimeError::formatted(status, "%s", siafd_b200_last_error(m_h));
    return v;
  }
  siafd_b200_handle *m_h;
  mutable IceModelVec2S m_topgsmooth, m_theta, m_thk_smooth;
};

class SIAFD_B200 : public SSB_Modifier {
public:
  // SIAFD::SIAFD, SIAFD.cc:42-91: everything the reference's constructor, FlowLaw (rheology/FlowLaw.cc:33-58) and
  // EnthalpyConverter (util/EnthalpyConverter.cc:55-69) read from Config goes into siafd_b200_config
  // comm_prefix (decomposed runs, g->size() > 1): a path prefix unique to this run that every rank can write to (e.g.
  // "/dev/shm/pism_<jobid>"), the rendezvous of siafd_b200_comm_init
  explicit SIAFD_B200(IceGrid::ConstPtr g, int device = -1, const char *comm_prefix = NULL)
      : SSB_Modifier(g), m_handle(NULL), m_stencil_width((int)m_config->get_number("grid.max_stencil_width")),
        m_h_x(g, "h_x", WITH_GHOSTS, 1), m_h_y(g, "h_y", WITH_GHOSTS, 1), m_D(g, "diffusivity", WITH_GHOSTS, 1),
        m_high_diffusivity_counter(0) {
    siafd_b200_config c;
    siafd_b200_default_config(&c);
    c.Mx = (int)g->Mx(), c.My = (int)g->My(), c.Mz = (int)g->Mz();
    c.xs = g->xs(), c.xm = g->xm(), c.ys = g->ys(), c.ym = g->ym();
    c.dx = g->dx(), c.dy = g->dy();
    c.z = g->z().data();
    c.w_geom = m_stencil_width, c.w_3d_in = m_stencil_width, c.w_stag = 1, c.w_uv = 1, c.w_sliding = 1;
    const Config &cf = *m_config;
    c.ec_p_air = cf.get_number("surface.pressure");
    c.ec_g = cf.get_number("constants.standard_gravity");
    c.ec_beta = cf.get_number("constants.ice.beta_Clausius_Clapeyron");
    c.ec_rho_i = cf.get_number("constants.ice.density");
    c.ec_c_i = cf.get_number("constants.ice.specific_heat_capacity");
    c.ec_c_w = cf.get_number("constants.fresh_water.specific_heat_capacity");
    c.ec_L = cf.get_number("constants.fresh_water.latent_heat_of_fusion");
    c.ec_T_melting = cf.get_number("constants.fresh_water.melting_point_temperature");
    c.ec_T_0 = cf.get_number("enthalpy_converter.T_reference");
    if (cf.get_flag("enthalpy_converter.cold_mode")) { // ColdEnthalpyConverter, EnthalpyConverter.cc:287-296
      c.ec_T_melting = 1e6;
      c.ec_beta = 0.0;
    }
    c.flow_law = flow_law_id(cf.get_string("stress_balance.sia.flow_law"));
    c.fl_n = cf.get_number("stress_balance.sia.Glen_exponent");
    c.fl_e = cf.get_number("stress_balance.sia.enhancement_factor");
    c.fl_e_interglacial = cf.get_number("stress_balance.sia.enhancement_factor_interglacial");
    c.fl_A_cold = cf.get_number("flow_law.Paterson_Budd.A_cold");
    c.fl_A_warm = cf.get_number("flow_law.Paterson_Budd.A_warm");
    c.fl_Q_cold = cf.get_number("flow_law.Paterson_Budd.Q_cold");
    c.fl_Q_warm = cf.get_number("flow_law.Paterson_Budd.Q_warm");
    c.fl_T_crit = cf.get_number("flow_law.Paterson_Budd.T_critical");
    c.fl_R = cf.get_number("constants.ideal_gas_constant");
    c.fl_rho = cf.get_number("constants.ice.density");
    c.fl_g = cf.get_number("constants.standard_gravity");
    c.fl_beta = cf.get_number("constants.ice.beta_Clausius_Clapeyron");
    c.fl_T_melting = cf.get_number("constants.fresh_water.melting_point_temperature");
    c.gpbld_T_0 = cf.get_number("constants.fresh_water.melting_point_temperature");
    c.gpbld_water_frac_coeff = cf.get_number("flow_law.gpbld.water_frac_coeff");
    c.gpbld_water_frac_limit = cf.get_number("flow_law.gpbld.water_frac_observed_limit");
    c.iso_softness_A = cf.get_number("flow_law.isothermal_Glen.ice_softness");
    c.hooke_Q = cf.get_number("flow_law.Hooke.Q"), c.hooke_A = cf.get_number("flow_law.Hooke.A");
    c.hooke_C = cf.get_number("flow_law.Hooke.C"), c.hooke_K = cf.get_number("flow_law.Hooke.k");
    c.hooke_Tr = cf.get_number("flow_law.Hooke.Tr");
    c.grain_size = cf.get_number("constants.ice.grain_size");
    c.gradient_method = gradient_id(cf.get_string("stress_balance.sia.surface_gradient_method"));
    c.limit_diffusivity = cf.get_flag("stress_balance.sia.limit_diffusivity") ? 1 : 0;
    c.grain_size_age_coupling = cf.get_flag("stress_balance.sia.grain_size_age_coupling") ? 1 : 0;
    c.e_age_coupling = cf.get_flag("stress_balance.sia.e_age_coupling") ? 1 : 0;
    c.D_limit = cf.get_number("stress_balance.sia.max_diffusivity");
    c.eemian_start = cf.get_number("time.eemian_start");
    c.eemian_end = cf.get_number("time.eemian_end");
    c.holocene_start = cf.get_number("time.holocene_start");
    c.smoother_range = cf.get_number("stress_balance.sia.bed_smoother.range");
    c.theta_min = cf.get_number("stress_balance.sia.bed_smoother.theta_min");
    c.sea_water_density = cf.get_number("constants.sea_water.density");
    c.ice_free_thickness = cf.get_number("geometry.ice_free_thickness_standard");
    c.dry_simulation = cf.get_flag("ocean.always_grounded") ? 1 : 0;
    int status = siafd_b200_create(&c, device, &m_handle);
    if (status != SIAFD_B200_OK) {
      throw RuntimeError::formatted(status, "%s", siafd_b200_last_error(NULL));
    }
    if (g->size() > 1) {
      if (comm_prefix == NULL) {
        siafd_b200_destroy(m_handle);
        throw RuntimeError::formatted(SIAFD_B200_ERR_BAD_ARGUMENT, "SIAFD_B200 on %d ranks needs a rendezvous prefix", g->size());
      }
      status = siafd_b200_comm_init(m_handle, g->rank(), g->size(), comm_prefix, 120.0);
      if (status != SIAFD_B200_OK) {
        const std::string why = siafd_b200_last_error(m_handle);
        siafd_b200_destroy(m_handle);
        throw RuntimeError::formatted(status, "%s", why.c_str());
      }
    }
    m_flow_law.reset(new rheology::FlowLaw_B200(m_handle, cf.get_string("stress_balance.sia.flow_law"), c.fl_n, c.fl_e));
    m_bed_smoother.reset(new BedSmoother_B200(g, m_handle, m_stencil_width));
  }
  virtual ~SIAFD_B200() { siafd_b200_destroy(m_handle); }

  virtual void init() { SSB_Modifier::init(); } // SIAFD.cc:98-118 (log messages only)

  // SIAFD::update, SIAFD.cc:122-155.  `full_update == false` leaves m_u, m_v untouched (:149-154).
  virtual void update(const IceModelVec2V &sliding_velocity, const Inputs &inputs, bool full_update) {
    const Geometry &geometry = *inputs.geometry;
    if (inputs.new_bed_elevation) { // :130-134 (BedSmoother::preprocess_bed)
      if (m_grid->size() > 1 && m_config->get_number("stress_balance.sia.bed_smoother.range") > 0.0) {
        // the reference gathers the bed on rank 0 (BedSmoother.cc:157-267); a decomposed PISM build hands the five
        // fields its own BedSmoother computed to siafd_b200_set_smoothed_bed instead
        throw RuntimeError::formatted(SIAFD_B200_ERR_BAD_CONFIG, "bed smoother on several ranks: use siafd_b200_set_smoothed_bed");
      }
      std::vector<double> bed((size_t)m_grid->Mx() * m_grid->My());
      for (int j = m_grid->ys(); j < m_grid->ys() + m_grid->ym(); ++j)
        for (int i = m_grid->xs(); i < m_grid->xs() + m_grid->xm(); ++i) bed[(size_t)j * m_grid->Mx() + i] = geometry.bed_elevation(i, j);
      check(siafd_b200_preprocess_bed(m_handle, bed.data()));
    }
    siafd_b200_inputs in;
    in.surface = geometry.ice_surface_elevation.get_array();
    in.thickness = geometry.ice_thickness.get_array();
    in.mask = geometry.cell_type.get_array();
    in.bed = geometry.bed_elevation.get_array();
    in.enthalpy = inputs.enthalpy->get_array();
    in.age = inputs.age ? inputs.age->get_array() : NULL;
    in.sliding = sliding_velocity.get_array();
    in.current_time = m_grid->current_time(); // :564
    in.memory_space = 0;
    in.ghosts_valid = 1; // PISM's Vecs always carry valid ghosts at this point
    siafd_b200_outputs out;
    out.h_x = m_h_x.get_array(), out.h_y = m_h_y.get_array(), out.D = m_D.get_array();
    out.flux = m_diffusive_flux.get_array();
    out.u = m_u.get_array(), out.v = m_v.get_array();
    out.memory_space = 0, out.pad = 0;
    // on several ranks the status is the same everywhere (the library reduces the error flags with D_max), so that
    // every rank throws, as under ParallelSection (util/error_handling.cc:189-214)
    check(siafd_b200_update(m_handle, &in, &out, full_update ? 1 : 0));
    m_D_max = siafd_b200_max_diffusivity(m_handle);                           // :748 GlobalMax
    m_high_diffusivity_counter = siafd_b200_high_diffusivity_count(m_handle); // :750 GlobalSum
  }

  const IceModelVec2Stag &surface_gradient_x() const { return m_h_x; } // SIAFD.cc:963-973
  const IceModelVec2Stag &surface_gradient_y() const { return m_h_y; }
  const IceModelVec2Stag &diffusivity() const { return m_D; }
  const BedSmoother_B200 &bed_smoother() const { return *m_bed_smoother; }                       // SIAFD.cc:975-977
  std::shared_ptr<const rheology::FlowLaw_B200> flow_law() const { return m_flow_law; }           // SSB_Modifier.cc:89-91
  int high_diffusivity_counter() const { return m_high_diffusivity_counter; }                    // SIAFD.cc:750, summed over ranks
  siafd_b200_handle *handle() { return m_handle; }

  // rheology/FlowLawFactory.cc:71-87
  static int flow_law_id(const std::string &name) {
    const char *names[] = {"isothermal_glen", "pb", "gpbld", "hooke", "arr", "arrwarm", "gk"};
    for (int k = 0; k < 7; ++k)
      if (name == names[k]) return k;
    throw RuntimeError::formatted(SIAFD_B200_ERR_BAD_CONFIG, "Selected ice flow law \"%s\" is not available", name.c_str());
  }

private:
  // status code -> RuntimeError, with the message the reference would have thrown
  void check(int status) {
    if (status != SIAFD_B200_OK) {
      throw RuntimeError::formatted(status, "%s", siafd_b200_last_error(m_handle));
    }
  }
  // SIAFD.cc:197-220
  static int gradient_id(const std::string &name) {
    if (name == "haseloff") return SIAFD_B200_GRAD_HASELOFF;
    if (name == "mahaffy") return SIAFD_B200_GRAD_MAHAFFY;
    if (name == "eta") return SIAFD_B200_GRAD_ETA;
    throw RuntimeError::formatted(SIAFD_B200_ERR_BAD_CONFIG, "value of sia.surface_gradient_method, option '-gradient %s', is not valid",
                                  name.c_str());
  }
  siafd_b200_handle *m_handle;
  const int m_stencil_width;
  IceModelVec2Stag m_h_x, m_h_y, m_D;
  int m_high_diffusivity_counter;
  std::shared_ptr<rheology::FlowLaw_B200> m_flow_law;
  std::shared_ptr<BedSmoother_B200> m_bed_smoother;
};

// stressbalance/timestepping.hh: what max_timestep_cfl_2d / _3d return
struct CFLData {
  CFLData() : dt_max(0.0), u_max(0.0), v_max(0.0), w_max(0.0) {}
  double dt_max, u_max, v_max, w_max;
};

// The SIA-only StressBalance container (stressbalance/StressBalance.cc:140-212 with ZeroSliding as the shallow
// stress balance, factory.cc:59-64): update() runs the modifier and, on a full update, the vertical velocity
// from incompressibility (StressBalance.cc:283-424), the volumetric strain heating (:426-642,
// N3) and the CFL reductions (timestepping.cc:42-153, N3) on the device fields the modifier left there.
class StressBalance_B200 {
public:
  StressBalance_B200(IceGrid::ConstPtr g, SIAFD_B200 *modifier)
      : m_grid(g), m_modifier(modifier), m_zero_sliding(g, "velbar", WITH_GHOSTS, 1), m_w(g, "wvel_rel", WITHOUT_GHOSTS),
        m_strain_heating(g, "strain_heating", WITHOUT_GHOSTS) {
    m_zero_sliding.set(0.0); // ZeroSliding
  }
  ~StressBalance_B200() { delete m_modifier; } // StressBalance.cc:157-160
  void init() { m_modifier->init(); }
  void update(const Inputs &inputs, bool full_update) {
    m_modifier->update(m_zero_sliding, inputs, full_update);
    siafd_b200_handle *h = m_modifier->handle();
    const Config &cf = *m_grid->config();
    int status = SIAFD_B200_OK;
    if (full_update) {
      if (inputs.basal_melt_rate) {
        status = siafd_b200_upload(h, SIAFD_B200_F_BASAL_MELT, inputs.basal_melt_rate->get_array());
      }
      const bool upstream = cf.get_string("stress_balance.vertical_velocity_approximation") == "upstream";
      if (status == SIAFD_B200_OK) {
        status = siafd_b200_compute_vertical_velocity(h, upstream ? 1 : 0, inputs.basal_melt_rate ? 1 : 0);
      }
      if (status == SIAFD_B200_OK) {
        status = siafd_b200_download(h, SIAFD_B200_F_W, m_w.get_array());
      }
      if (status == SIAFD_B200_OK) { // the SHALLOW stress balance's flow law (StressBalance.cc:508-524)
        status = siafd_b200_compute_strain_heating(h, SIAFD_B200::flow_law_id(cf.get_string("stress_balance.ssa.flow_law")),
                                                   cf.get_number("stress_balance.ssa.Glen_exponent"),
                                                   cf.get_number("stress_balance.ssa.enhancement_factor"));
      }
      if (status == SIAFD_B200_OK) {
        status = siafd_b200_download(h, SIAFD_B200_F_STRAIN_HEATING, m_strain_heating.get_array());
      }
    }
    if (status == SIAFD_B200_OK) { // StressBalance.cc:196-205
      double out[8];
      status = siafd_b200_cfl(h, cf.get_number("time_stepping.maximum_time_step") * seconds_per_year_udunits(),
                              full_update ? 1 : 0, out);
      if (status == SIAFD_B200_OK && m_grid->size() > 1) { // GlobalMin / GlobalMax, timestepping.cc:85-99, :145-151
        double dts[2] = {out[0], out[4]}, vel[6] = {out[1], out[2], out[3], out[5], out[6], 0.0};
        status = siafd_b200_comm_allreduce(h, 1, 2, dts);
        if (status == SIAFD_B200_OK) status = siafd_b200_comm_allreduce(h, 0, 6, vel);
        out[0] = dts[0], out[4] = dts[1], out[1] = vel[0], out[2] = vel[1], out[3] = vel[2], out[5] = vel[3], out[6] = vel[4];
      }
      if (status == SIAFD_B200_OK) {
        if (full_update) m_cfl_3d.dt_max = out[0], m_cfl_3d.u_max = out[1], m_cfl_3d.v_max = out[2], m_cfl_3d.w_max = out[3];
        m_cfl_2d.dt_max = out[4], m_cfl_2d.u_max = out[5], m_cfl_2d.v_max = out[6], m_cfl_2d.w_max = 0.0;
      }
    }
    if (status != SIAFD_B200_OK) {
      throw RuntimeError::formatted(status, "%s", siafd_b200_last_error(h));
    }
  }
  static double seconds_per_year_udunits() { return 365.242198781 * 86400.0; } // convert(sys, 1, "year", "seconds")
  const IceModelVec3 &velocity_u() const { return m_modifier->velocity_u(); }
  const IceModelVec3 &velocity_v() const { return m_modifier->velocity_v(); }
  const IceModelVec3 &velocity_w() const { return m_w; }
  const IceModelVec3 &volumetric_strain_heating() const { return m_strain_heating; }
  const IceModelVec2V &advective_velocity() const { return m_zero_sliding; }
  const IceModelVec2Stag &diffusive_flux() { return m_modifier->diffusive_flux(); }
  double max_diffusivity() const { return m_modifier->max_diffusivity(); }
  CFLData max_timestep_cfl_2d() const { return m_cfl_2d; }
  CFLData max_timestep_cfl_3d() const { return m_cfl_3d; }
  const SIAFD_B200 *modifier() const { return m_modifier; }
  SIAFD_B200 *modifier() { return m_modifier; }

private:
  IceGrid::ConstPtr m_grid;
  SIAFD_B200 *m_modifier;
  IceModelVec2V m_zero_sliding;
  IceModelVec3 m_w, m_strain_heating;
  CFLData m_cfl_2d, m_cfl_3d;
};

} // namespace stressbalance
} // namespace pism
