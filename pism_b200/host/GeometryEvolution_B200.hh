// GeometryEvolution_B200.hh -- C++ host side of the mass-continuity consumer of SIAFD's outputs:
// the reference's GeometryEvolution interface (src/geometry/GeometryEvolution.hh; flow_step / apply_flux_divergence /
// source_term_step / apply_mass_fluxes and the getters of their results, GeometryEvolution.cc:241-390) over the C ABI.
// The thickness, bed and diffusive flux are on the device since the stress-balance update, so a step uploads only
// what changed on the host and downloads the 2D results PISM's diagnostics read.  Default configuration only
// (geometry.part_grid.enabled = no).  Written against pism_mirror.hh; nothing here computes physics.
#pragma once
#include "SIAFD_B200.hh"

namespace pism {

class GeometryEvolution_B200 {
public:
  GeometryEvolution_B200(IceGrid::ConstPtr g, siafd_b200_handle *handle)
      : m_grid(g), m_h(handle), m_flux_divergence(g, "flux_divergence", WITHOUT_GHOSTS),
        m_thickness_change(g, "thickness_change", WITHOUT_GHOSTS), m_conservation_error(g, "conservation_error", WITHOUT_GHOSTS),
        m_effective_SMB(g, "effective_SMB", WITHOUT_GHOSTS), m_effective_BMB(g, "effective_BMB", WITHOUT_GHOSTS),
        m_ice_density(g->config()->get_number("constants.ice.density")),
        m_use_bmr(g->config()->get_flag("geometry.update.use_basal_melt_rate")) {}

  // GeometryEvolution.cc:241-324.  The bc masks may be NULL (= no Dirichlet locations).
  void flow_step(const Geometry &geometry, double dt, const IceModelVec2V &advective_velocity,
                 const IceModelVec2Stag &diffusive_flux, const IceModelVec2Int *velocity_bc_mask,
                 const IceModelVec2Int *thickness_bc_mask) {
    check(siafd_b200_upload(m_h, SIAFD_B200_F_THICKNESS, geometry.ice_thickness.get_array()));
    check(siafd_b200_upload(m_h, SIAFD_B200_F_BED, geometry.bed_elevation.get_array()));
    check(siafd_b200_upload(m_h, SIAFD_B200_F_SEA_LEVEL, geometry.sea_level_elevation.get_array()));
    check(siafd_b200_upload(m_h, SIAFD_B200_F_SLIDING, advective_velocity.get_array()));
    check(siafd_b200_upload(m_h, SIAFD_B200_F_FLUX, diffusive_flux.get_array()));
    if (velocity_bc_mask) check(siafd_b200_upload(m_h, SIAFD_B200_F_VEL_BC_MASK, velocity_bc_mask->get_array()));
    if (thickness_bc_mask) check(siafd_b200_upload(m_h, SIAFD_B200_F_THK_BC_MASK, thickness_bc_mask->get_array()));
    check(siafd_b200_mass_flow_step(m_h, dt));
    check(siafd_b200_download(m_h, SIAFD_B200_F_FLUX_DIV, m_flux_divergence.get_array()));
    check(siafd_b200_download(m_h, SIAFD_B200_F_THK_CHANGE, m_thickness_change.get_array()));
    check(siafd_b200_download(m_h, SIAFD_B200_F_CONS_ERR, m_conservation_error.get_array()));
  }
  // :347-350 (the area-specific volume only changes with part_grid)
  void apply_flux_divergence(Geometry &geometry) const {
    for (int j = m_grid->ys(); j < m_grid->ys() + m_grid->ym(); ++j)
      for (int i = m_grid->xs(); i < m_grid->xs() + m_grid->xm(); ++i) geometry.ice_thickness(i, j) = geometry.ice_thickness(i, j) + 1.0 * m_thickness_change(i, j);
  }
  // :327-343
  void source_term_step(const Geometry &geometry, double dt, const IceModelVec2Int *thickness_bc_mask,
                        const IceModelVec2S &surface_mass_balance_rate, const IceModelVec2S *basal_melt_rate) {
    check(siafd_b200_upload(m_h, SIAFD_B200_F_THICKNESS, geometry.ice_thickness.get_array()));
    check(siafd_b200_upload(m_h, SIAFD_B200_F_MASK, geometry.cell_type.get_array()));
    check(siafd_b200_upload(m_h, SIAFD_B200_F_SMB, surface_mass_balance_rate.get_array()));
    const bool bmr = m_use_bmr && basal_melt_rate != NULL;
    if (bmr) check(siafd_b200_upload(m_h, SIAFD_B200_F_BASAL_MELT, basal_melt_rate->get_array()));
    if (thickness_bc_mask) check(siafd_b200_upload(m_h, SIAFD_B200_F_THK_BC_MASK, thickness_bc_mask->get_array()));
    check(siafd_b200_mass_source_step(m_h, dt, m_ice_density, bmr ? 1 : 0));
    check(siafd_b200_download(m_h, SIAFD_B200_F_EFF_SMB, m_effective_SMB.get_array()));
    check(siafd_b200_download(m_h, SIAFD_B200_F_EFF_BMB, m_effective_BMB.get_array()));
  }
  // :360-390: the same order of additions as the non-negativity code
  void apply_mass_fluxes(Geometry &geometry) const {
    for (int j = m_grid->ys(); j < m_grid->ys() + m_grid->ym(); ++j)
      for (int i = m_grid->xs(); i < m_grid->xs() + m_grid->xm(); ++i) {
        const double H_new = (geometry.ice_thickness(i, j) + m_effective_SMB(i, j)) + m_effective_BMB(i, j);
        geometry.ice_thickness(i, j) = H_new;
      }
  }
  const IceModelVec2S &flux_divergence() const { return m_flux_divergence; }
  const IceModelVec2S &thickness_change_due_to_flow() const { return m_thickness_change; }
  const IceModelVec2S &conservation_error() const { return m_conservation_error; }
  const IceModelVec2S &top_surface_mass_balance() const { return m_effective_SMB; }
  const IceModelVec2S &bottom_surface_mass_balance() const { return m_effective_BMB; }

private:
  void check(int status) const {
    if (status != SIAFD_B200_OK) {
      throw RuntimeError::formatted(status, "%s", siafd_b200_last_error(m_h));
    }
  }
  IceGrid::ConstPtr m_grid;
  siafd_b200_handle *m_h;
  IceModelVec2S m_flux_divergence, m_thickness_change, m_conservation_error, m_effective_SMB, m_effective_BMB;
  double m_ice_density;
  bool m_use_bmr;
};

} // namespace pism
