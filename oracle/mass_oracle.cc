// mass_oracle.cc -- TEST INFRASTRUCTURE ONLY (part of liboracle_siafd.so).
//
// CPU restatement of the consumers of SIAFD's outputs, SURVEY.md 8(f) rows N1 and N3 (CFL part):
//   * GeometryEvolution::flow_step / apply_flux_divergence   src/geometry/GeometryEvolution.cc:241-350
//       compute_interface_fluxes :535-654, limit_diffusive_flux :462-525, limit_advective_velocity :395-457,
//       compute_flux_divergence :644-669, update_in_place :689-771 (part_grid off, the default
//       geometry.part_grid.enabled = no), ensure_nonnegativity :960-1000; and, for pinning only, the same step with
//       part_grid on (orc_mass_flow_step_part_grid: residual redistribution :777-944,
//       part_grid_threshold_thickness.cc) against the golden numbers of test/mass_transport.py
//   * GeometryEvolution::source_term_step / apply_mass_fluxes  :327-343, :360-390, effective_change :1005-1011,
//       compute_surface_and_basal_mass_balance :1028-1076
//   * max_timestep_cfl_3d / max_timestep_cfl_2d                src/stressbalance/timestepping.cc:42-101, :113-153
// Same loops, same expression order, no FMA contraction (-ffp-contract=off), plain arrays in PISM's local
// ghosted layout.  Pinned by the reference's golden output of `pismv -test C` (test/regression/test_15.sh)
// through tests/test_oracle_pismv.py, which time-steps these functions with the reference's step logic.
#include <cmath>
#include <vector>

#include "siafd_oracle.h"

namespace {

// util/Mask.hh:29-66
enum { M_BEDROCK = 0, M_GROUNDED = 2, M_FLOATING = 3, M_OCEAN = 4 };
inline bool m_grounded(int M) { return not(M >= M_FLOATING); }
inline bool m_ocean(int M) { return M >= M_FLOATING; }
inline bool m_icy(int M) { return (M == M_GROUNDED) || (M == M_FLOATING); }
inline bool m_ice_free(int M) { return not m_icy(M); }
inline bool grounded_ice(int M) { return m_icy(M) && m_grounded(M); }
inline bool floating_ice(int M) { return m_icy(M) && m_ocean(M); }
inline bool ice_free_ocean(int M) { return m_ocean(M) && m_ice_free(M); }
inline bool ice_free_land(int M) { return m_grounded(M) && m_ice_free(M); }

// GeometryEvolution.cc:395-457; returns NaN for the "cannot handle the case" throw (unreachable for the four mask values)
double limit_advective_velocity(int current, int neighbor, double velocity) {
  if (grounded_ice(current) and grounded_ice(neighbor)) return velocity;
  if ((grounded_ice(current) and floating_ice(neighbor)) or (floating_ice(current) and grounded_ice(neighbor)))
    return velocity;
  if ((grounded_ice(current) and ice_free_land(neighbor)) or (ice_free_land(current) and grounded_ice(neighbor)))
    return velocity;
  if ((grounded_ice(current) and ice_free_ocean(neighbor)) or (ice_free_ocean(current) and grounded_ice(neighbor)))
    return velocity;
  if (floating_ice(current) and floating_ice(neighbor)) return velocity;
  if ((floating_ice(current) and ice_free_land(neighbor)) or (ice_free_land(current) and floating_ice(neighbor)))
    return 0.0;
  if ((floating_ice(current) and ice_free_ocean(neighbor)) or (ice_free_ocean(current) and floating_ice(neighbor)))
    return velocity;
  if (ice_free_land(current) and ice_free_land(neighbor)) return 0.0;
  if ((ice_free_land(current) and ice_free_ocean(neighbor)) or (ice_free_ocean(current) and ice_free_land(neighbor)))
    return 0.0;
  if (ice_free_ocean(current) and ice_free_ocean(neighbor)) return 0.0;
  return NAN;
}

// GeometryEvolution.cc:462-525
double limit_diffusive_flux(int current, int neighbor, double flux) {
  if (grounded_ice(current) and grounded_ice(neighbor)) return flux;
  if ((grounded_ice(current) and floating_ice(neighbor)) or (floating_ice(current) and grounded_ice(neighbor)))
    return flux;
  if ((grounded_ice(current) and ice_free_land(neighbor)) or (ice_free_land(current) and grounded_ice(neighbor)))
    return flux;
  if ((grounded_ice(current) and ice_free_ocean(neighbor)) or (ice_free_ocean(current) and grounded_ice(neighbor)))
    return flux;
  if (floating_ice(current) and floating_ice(neighbor)) return 0.0;
  if ((floating_ice(current) and ice_free_land(neighbor)) or (ice_free_land(current) and floating_ice(neighbor)))
    return 0.0;
  if ((floating_ice(current) and ice_free_ocean(neighbor)) or (ice_free_ocean(current) and floating_ice(neighbor)))
    return 0.0;
  if (ice_free_land(current) and ice_free_land(neighbor)) return 0.0;
  if ((ice_free_land(current) and ice_free_ocean(neighbor)) or (ice_free_ocean(current) and ice_free_land(neighbor)))
    return 0.0;
  if (ice_free_ocean(current) and ice_free_ocean(neighbor)) return 0.0;
  return NAN;
}

// GeometryEvolution.cc:1005-1011
inline double effective_change(double H, double dH) {
  if (H + dH <= 0) {
    return -H;
  } else {
    return dH;
  }
}

} // namespace

extern "C" {

// flow_step + apply_flux_divergence on one patch.  bed, sea_level (NULL = 0), thickness: 2D w_geom with valid
// ghosts; velocity: Vector2 with ghost width w_sliding >= 1 and valid ghosts, or NULL (= 0, ZeroSliding);
// velocity_bc_mask / thickness_bc_mask: 2D w_geom or NULL (= 0); Q: Stag w_stag, ghosts valid (SIAFD computes
// it on owned + 1).  Outputs (owned points only, [ym][xm]): flux_divergence, thickness_change (after
// ensure_nonnegativity), conservation_error; thickness is updated in place on the owned points
// (H_old + thickness_change, GeometryEvolution.cc:347-350); the caller refreshes its ghosts.
// The first half of flow_step (:247-282): gc.compute on the ghosted copies, compute_interface_fluxes,
// compute_flux_divergence.  cell_type (w_geom, every local point) and flux_divergence (owned) are outputs.
static int flow_step_flux_divergence(const orc_params *p, const double *sea_level, const double *bed,
                                     const double *thickness, const double *velocity, const double *velocity_bc_mask,
                                     const double *thickness_bc_mask, const double *Q, std::vector<double> &cell_type,
                                     double *flux_divergence) {
  const int wg = p->w_geom, ws = p->w_stag, wv = p->w_sliding;
  const long nxg = p->xm + 2 * wg, nyg = p->ym + 2 * wg, nxs = p->xm + 2 * ws, nxv = p->xm + 2 * wv;
  if (velocity != NULL and wv < 1) return ORC_ERR_BAD_CONFIG;
  auto G = [&](int i, int j) { return (long)(j - (p->ys - wg)) * nxg + (i - (p->xs - wg)); };
  auto S = [&](int i, int j, int o) { return ((long)(j - (p->ys - ws)) * nxs + (i - (p->xs - ws))) * 2 + o; };
  auto V = [&](int i, int j, int c) {
    return velocity ? velocity[((long)(j - (p->ys - wv)) * nxv + (i - (p->xs - wv))) * 2 + c] : 0.0;
  };
  // :262-266 gc.compute on the ghosted copies (pointwise; ghosts of the result = result on the ghosts)
  std::vector<double> zero;
  cell_type.assign(nxg * nyg, 0.0);
  if (sea_level == NULL) {
    zero.assign(nxg * nyg, 0.0);
    sea_level = zero.data();
  }
  orc_geometry_compute(p, (int)(nxg * nyg), sea_level, bed, thickness, cell_type.data(), NULL);
  auto M = [&](int i, int j) { return (int)floor(cell_type[G(i, j)] + 0.5); };
  auto BC = [&](int i, int j) { return velocity_bc_mask ? (int)floor(velocity_bc_mask[G(i, j)] + 0.5) : 0; };

  // compute_interface_fluxes (:535-654) on owned + 1 (= owned, then update_ghosts :279)
  std::vector<double> flux((p->xm + 2) * (long)(p->ym + 2) * 2);
  auto FL = [&](int i, int j, int o) -> double & {
    return flux[((long)(j - (p->ys - 1)) * (p->xm + 2) + (i - (p->xs - 1))) * 2 + o];
  };
  for (int j = p->ys - 1; j < p->ys + p->ym + 1; ++j) {
    for (int i = p->xs - 1; i < p->xs + p->xm + 1; ++i) {
      const int Mc = M(i, j), BCc = BC(i, j);
      const double H = thickness[G(i, j)];
      const double Vu = V(i, j, 0), Vv = V(i, j, 1);
      for (int n = 0; n < 2; ++n) {
        const int oi = 1 - n, oj = n, i_n = i + oi, j_n = j + oj;
        const int M_n = M(i_n, j_n);
        double v = 0.0;
        {
          const double Vnu = V(i_n, j_n, 0), Vnv = V(i_n, j_n, 1);
          {
            if (m_icy(Mc) and m_icy(M_n)) {
              v = (n == 0 ? 0.5 * (Vu + Vnu) : 0.5 * (Vv + Vnv));
            } else if (m_icy(Mc) and m_ice_free(M_n)) {
              v = (n == 0 ? Vu : Vv);
            } else if (m_ice_free(Mc) and m_icy(M_n)) {
              v = (n == 0 ? Vnu : Vnv);
            } else if (m_ice_free(Mc) and m_ice_free(M_n)) {
              v = 0.0;
            }
          }
          {
            const int BC_n = BC(i_n, j_n);
            if (BCc == 1 and BC_n == 1) {
              v = (n == 0 ? 0.5 * (Vu + Vnu) : 0.5 * (Vv + Vnv));
            } else if (BCc == 1 and BC_n == 0) {
              v = (n == 0 ? Vu : Vv);
            } else if (BCc == 0 and BC_n == 1) {
              v = (n == 0 ? Vnu : Vnv);
            }
          }
          v = limit_advective_velocity(Mc, M_n, v);
        }
        const double H_n = thickness[G(i_n, j_n)], Q_advective = v * (v > 0.0 ? H : H_n);
        const double Q_diffusive = limit_diffusive_flux(Mc, M_n, Q[S(i, j, n)]);
        FL(i, j, n) = Q_diffusive + Q_advective;
      }
    }
  }
  // compute_flux_divergence (:644-669)
  const double dx = p->dx, dy = p->dy;
  for (int j = p->ys; j < p->ys + p->ym; ++j) {
    for (int i = p->xs; i < p->xs + p->xm; ++i) {
      const long o = (long)(j - p->ys) * p->xm + (i - p->xs);
      double divQ;
      if (thickness_bc_mask != NULL and thickness_bc_mask[G(i, j)] > 0.5) {
        divQ = 0.0;
      } else {
        const double Qe = FL(i, j, 0), Qw = FL(i - 1, j, 0), Qn = FL(i, j, 1), Qs = FL(i, j - 1, 1);
        divQ = (Qe - Qw) / dx + (Qn - Qs) / dy;
      }
      flux_divergence[o] = divQ;
    }
  }
  return ORC_OK;
}

int orc_mass_flow_step(const orc_params *p, double dt, const double *sea_level, const double *bed, double *thickness,
                       const double *velocity, const double *velocity_bc_mask, const double *thickness_bc_mask,
                       const double *Q, double *flux_divergence, double *thickness_change,
                       double *conservation_error) {
  const int wg = p->w_geom;
  const long nxg = p->xm + 2 * wg;
  auto G = [&](int i, int j) { return (long)(j - (p->ys - wg)) * nxg + (i - (p->xs - wg)); };
  std::vector<double> cell_type;
  const int status = flow_step_flux_divergence(p, sea_level, bed, thickness, velocity, velocity_bc_mask,
                                               thickness_bc_mask, Q, cell_type, flux_divergence);
  if (status != ORC_OK) return status;
  // update_in_place (:689-771, no part_grid), compute changes (:295-300), ensure_nonnegativity (:960-1000),
  // apply_flux_divergence (:347-350)
  for (int j = p->ys; j < p->ys + p->ym; ++j) {
    for (int i = p->xs; i < p->xs + p->xm; ++i) {
      const long o = (long)(j - p->ys) * p->xm + (i - p->xs);
      const double divQ = flux_divergence[o];
      const double H_old = thickness[G(i, j)];
      double H_new = H_old;
      H_new += -dt * divQ;
      double dH = H_new + (-1.0) * H_old;
      conservation_error[o] = 0.0;
      if (H_old + dH < 0.0) {
        // (sic) GeometryEvolution.cc:980-983 assigns H, not -H
        conservation_error[o] += -(H_old + dH);
        dH = H_old;
      }
      thickness_change[o] = dH;
    }
  }
  for (int j = p->ys; j < p->ys + p->ym; ++j) {
    for (int i = p->xs; i < p->xs + p->xm; ++i) {
      thickness[G(i, j)] = thickness[G(i, j)] + 1.0 * thickness_change[(long)(j - p->ys) * p->xm + (i - p->xs)];
    }
  }
  return ORC_OK;
}

// geometry/part_grid_threshold_thickness.cc:33-70; star stencils as {ij, e, w, n, s}
static double part_grid_threshold_thickness(const int M[5], const double H[5], const double h[5],
                                            double bed_elevation) {
  double H_average = 0.0, h_average = 0.0, H_threshold = 0.0;
  int N = 0;
  const int dirs[] = {3, 1, 4, 2}; // North, East, South, West
  for (int n = 0; n < 4; ++n) {
    const int d = dirs[n];
    if (m_icy(M[d])) {
      H_average += H[d];
      h_average += h[d];
      N++;
    }
  }
  if (N == 0) {
    return 0.0;
  }
  H_average = H_average / N;
  h_average = h_average / N;
  if (bed_elevation + H_average > h_average) {
    H_threshold = h_average - bed_elevation;
  } else {
    H_threshold = H_average;
  }
  return std::max(H_threshold, 0.0);
}

// flow_step + apply_flux_divergence WITH geometry.part_grid.enabled (GeometryEvolution.cc:241-357, update_in_place
// :689-816, residual_redistribution_iteration :819-944).  The CUDA path implements the default (part_grid off); this
// variant exists so that the restated interface fluxes (advective part, bc masks, flux divergence) -- shared with
// orc_mass_flow_step -- can be pinned against the golden numbers of the reference's test/mass_transport.py:169-173,
// which runs with part_grid on.  One patch covering the whole domain (ghost updates are periodic self-wraps).
// thickness, area_specific_volume: w_geom, ghosts valid, updated in place (H_old + thickness_change etc., ghosts
// wrapped); outputs owned-only.
int orc_mass_flow_step_part_grid(const orc_params *p, double dt, const double *sea_level, const double *bed,
                                 double *thickness, double *area_specific_volume, const double *velocity,
                                 const double *velocity_bc_mask, const double *thickness_bc_mask, const double *Q,
                                 int max_iterations, double *flux_divergence, double *thickness_change,
                                 double *area_specific_volume_change, double *conservation_error) {
  if (p->xm != p->Mx or p->ym != p->My or p->xs != 0 or p->ys != 0) return ORC_ERR_BAD_CONFIG;
  const int wg = p->w_geom, xm = p->xm, ym = p->ym;
  if (wg < 1) return ORC_ERR_BAD_CONFIG;
  const long nxg = xm + 2 * wg, nyg = ym + 2 * wg, n_all = nxg * nyg;
  auto G = [&](int i, int j) { return (long)(j + wg) * nxg + (i + wg); };
  std::vector<double> zero;
  if (sea_level == NULL) {
    zero.assign(n_all, 0.0);
    sea_level = zero.data();
  }
  // :247-256 ghosted copies
  std::vector<double> H(thickness, thickness + n_all), V(area_specific_volume, area_specific_volume + n_all);
  std::vector<double> cell_type, surface(n_all), residual(n_all, 0.0), H_copy;
  int status = flow_step_flux_divergence(p, sea_level, bed, H.data(), velocity, velocity_bc_mask, thickness_bc_mask, Q,
                                         cell_type, flux_divergence);
  if (status != ORC_OK) return status;
  auto M = [&](int i, int j) { return (int)floor(cell_type[G(i, j)] + 0.5); };
  auto threshold_at = [&](int i, int j) {
    const int Ms[5] = {M(i, j), M(i + 1, j), M(i - 1, j), M(i, j + 1), M(i, j - 1)};
    const double Hs[5] = {H_copy[G(i, j)], H_copy[G(i + 1, j)], H_copy[G(i - 1, j)], H_copy[G(i, j + 1)],
                          H_copy[G(i, j - 1)]};
    const double hs[5] = {surface[G(i, j)], surface[G(i + 1, j)], surface[G(i - 1, j)], surface[G(i, j + 1)],
                          surface[G(i, j - 1)]};
    return part_grid_threshold_thickness(Ms, Hs, hs, bed[G(i, j)]);
  };
  // ---- update_in_place :689-771 ----
  orc_geometry_compute(p, (int)n_all, sea_level, bed, H.data(), cell_type.data(), surface.data());
  H_copy = H;
  for (int j = 0; j < ym; ++j) {
    for (int i = 0; i < xm; ++i) {
      double divQ = flux_divergence[(long)j * xm + i];
      const long g = G(i, j);
      const bool next_to_ice = m_icy(M(i + 1, j)) or m_icy(M(i - 1, j)) or m_icy(M(i, j + 1)) or m_icy(M(i, j - 1));
      if (ice_free_ocean(M(i, j)) and next_to_ice) {
        V[g] += -divQ * dt;
        double threshold = threshold_at(i, j);
        if (threshold == 0.0) {
          threshold = V[g];
        }
        if (V[g] >= threshold) {
          H[g] += threshold;
          residual[g] = V[g] - threshold;
          V[g] = 0.0;
        }
        divQ = 0.0;
      }
      H[g] += -dt * divQ;
    }
  }
  orc_wrap_ghosts(xm, ym, wg, 1, H.data());
  orc_geometry_compute(p, (int)n_all, sea_level, bed, H.data(), cell_type.data(), NULL); // compute_mask :770
  // ---- residual redistribution :777-815 ----
  bool done = false;
  for (int it = 0; it < max_iterations and not done; ++it) {
    // residual_redistribution_iteration :819-944
    orc_geometry_compute(p, (int)n_all, sea_level, bed, H.data(), cell_type.data(), NULL);
    for (int j = 0; j < ym; ++j) {
      for (int i = 0; i < xm; ++i) {
        const long g = G(i, j);
        if (residual[g] <= 0.0) {
          continue;
        }
        const int nb[4] = {M(i, j + 1), M(i + 1, j), M(i, j - 1), M(i - 1, j)};
        int N = 0;
        for (int n = 0; n < 4; ++n) {
          if (ice_free_ocean(nb[n])) {
            N++;
          }
        }
        if (N > 0) {
          residual[g] /= N;
        } else {
          H[g] += residual[g];
          residual[g] = 0.0;
        }
      }
    }
    orc_wrap_ghosts(xm, ym, wg, 1, residual.data());
    for (int j = 0; j < ym; ++j) {
      for (int i = 0; i < xm; ++i) {
        if (ice_free_ocean(M(i, j))) {
          V[G(i, j)] += (residual[G(i + 1, j)] + residual[G(i - 1, j)] + residual[G(i, j + 1)] + residual[G(i, j - 1)]);
        }
      }
    }
    std::fill(residual.begin(), residual.end(), 0.0);
    orc_wrap_ghosts(xm, ym, wg, 1, H.data());
    H_copy = H;
    orc_geometry_compute(p, (int)n_all, sea_level, bed, H.data(), cell_type.data(), surface.data());
    double remaining_residual = 0.0;
    for (int j = 0; j < ym; ++j) {
      for (int i = 0; i < xm; ++i) {
        const long g = G(i, j);
        if (V[g] <= 0.0) {
          continue;
        }
        double threshold = threshold_at(i, j);
        if (threshold == 0.0) {
          threshold = V[g];
        }
        if (V[g] >= threshold) {
          H[g] += threshold;
          residual[g] = V[g] - threshold;
          V[g] = 0.0;
          remaining_residual += residual[g];
        }
      }
    }
    done = not(remaining_residual > 0.0);
    orc_wrap_ghosts(xm, ym, wg, 1, H.data());
  }
  if (not done and max_iterations > 0) {
    for (long g = 0; g < n_all; ++g) {
      H[g] = H[g] + 1.0 * residual[g];
    }
  }
  // ---- changes :293-300, ensure_nonnegativity :960-1000, apply_flux_divergence :347-357 ----
  for (int j = 0; j < ym; ++j) {
    for (int i = 0; i < xm; ++i) {
      const long g = G(i, j), o = (long)j * xm + i;
      double dH = H[g] + (-1.0) * thickness[g], dV = V[g] + (-1.0) * area_specific_volume[g];
      conservation_error[o] = 0.0;
      if (thickness[g] + dH < 0.0) {
        conservation_error[o] += -(thickness[g] + dH);
        dH = thickness[g];
      }
      if (area_specific_volume[g] + dV < 0.0) {
        conservation_error[o] += -(area_specific_volume[g] + dV);
        dV = area_specific_volume[g];
      }
      thickness_change[o] = dH;
      area_specific_volume_change[o] = dV;
    }
  }
  for (int j = 0; j < ym; ++j) {
    for (int i = 0; i < xm; ++i) {
      const long g = G(i, j), o = (long)j * xm + i;
      thickness[g] = thickness[g] + 1.0 * thickness_change[o];
      area_specific_volume[g] = area_specific_volume[g] + 1.0 * area_specific_volume_change[o];
    }
  }
  orc_wrap_ghosts(xm, ym, wg, 1, thickness);
  orc_wrap_ghosts(xm, ym, wg, 1, area_specific_volume);
  return ORC_OK;
}

// source_term_step + apply_mass_fluxes on one patch.  thickness, mask: 2D w_geom; smb_flux [kg m-2 s-1] and
// basal_melt_rate [m s-1]: owned only; thickness_bc_mask: w_geom or NULL.  Outputs owned only.
int orc_mass_source_step(const orc_params *p, double dt, double ice_density, int use_bmr, double *thickness,
                         const double *mask, const double *thickness_bc_mask, const double *smb_flux,
                         const double *basal_melt_rate, double *effective_SMB, double *effective_BMB) {
  const int wg = p->w_geom;
  const long nxg = p->xm + 2 * wg;
  for (int j = p->ys; j < p->ys + p->ym; ++j) {
    for (int i = p->xs; i < p->xs + p->xm; ++i) {
      const long g = (long)(j - (p->ys - wg)) * nxg + (i - (p->xs - wg)), o = (long)(j - p->ys) * p->xm + (i - p->xs);
      const int Mc = (int)floor(mask[g] + 0.5);
      const int bc = thickness_bc_mask ? (int)floor(thickness_bc_mask[g] + 0.5) : 0;
      if (bc == 1 or ice_free_ocean(Mc)) {
        effective_SMB[o] = 0.0;
        effective_BMB[o] = 0.0;
        continue;
      }
      const double H = thickness[g];
      double dH_SMB = effective_change(H, dt * smb_flux[o] / ice_density);
      double dH_BMB = effective_change(H + dH_SMB, dt * (use_bmr ? -basal_melt_rate[o] : 0.0));
      effective_SMB[o] = dH_SMB;
      effective_BMB[o] = dH_BMB;
    }
  }
  for (int j = p->ys; j < p->ys + p->ym; ++j) {
    for (int i = p->xs; i < p->xs + p->xm; ++i) {
      const long g = (long)(j - (p->ys - wg)) * nxg + (i - (p->xs - wg)), o = (long)(j - p->ys) * p->xm + (i - p->xs);
      const double H_new = (thickness[g] + effective_SMB[o]) + effective_BMB[o];
      thickness[g] = H_new;
    }
  }
  return ORC_OK;
}

// max_timestep_cfl_3d (timestepping.cc:42-101) on one patch: thickness, mask w_geom; u, v w_uv; w owned only.
// out = {dt_max, u_max, v_max, w_max} (local values; the caller reduces over ranks: min, max, max, max).
int orc_cfl_3d(const orc_params *p, double max_dt_seconds, const double *thickness, const double *mask,
               const double *u3, const double *v3, const double *w3, double *out) {
  const int Mz = p->Mz, wg = p->w_geom, wuv = p->w_uv;
  const long nxg = p->xm + 2 * wg, nxu = p->xm + 2 * wuv;
  double dt_max = max_dt_seconds;
  const double one_over_dx = 1.0 / p->dx, one_over_dy = 1.0 / p->dy;
  double u_max = 0.0, v_max = 0.0, w_max = 0.0;
  for (int j = p->ys; j < p->ys + p->ym; ++j) {
    for (int i = p->xs; i < p->xs + p->xm; ++i) {
      const long g = (long)(j - (p->ys - wg)) * nxg + (i - (p->xs - wg));
      if (m_icy((int)floor(mask[g] + 0.5))) {
        int status = 0;
        const int ks = orc_k_below_height(p->z, Mz, thickness[g], &status);
        if (status != 0) return status;
        const double *u = u3 + ((long)(j - (p->ys - wuv)) * nxu + (i - (p->xs - wuv))) * Mz;
        const double *v = v3 + ((long)(j - (p->ys - wuv)) * nxu + (i - (p->xs - wuv))) * Mz;
        const double *w = w3 + ((long)(j - p->ys) * p->xm + (i - p->xs)) * Mz;
        for (int k = 0; k <= ks; ++k) {
          const double u_abs = fabs(u[k]), v_abs = fabs(v[k]);
          u_max = std::max(u_max, u_abs);
          v_max = std::max(v_max, v_abs);
          const double denom = fabs(u_abs * one_over_dx) + fabs(v_abs * one_over_dy);
          if (denom > 0.0) {
            dt_max = std::min(dt_max, 1.0 / denom);
          }
        }
        for (int k = 0; k <= ks; ++k) {
          w_max = std::max(w_max, fabs(w[k]));
        }
      }
    }
  }
  out[0] = dt_max;
  out[1] = u_max;
  out[2] = v_max;
  out[3] = w_max;
  return ORC_OK;
}

// max_timestep_cfl_2d (timestepping.cc:113-153): velocity = Vector2 with ghost width w_sliding, or NULL (= 0).
int orc_cfl_2d(const orc_params *p, double max_dt_seconds, const double *mask, const double *velocity, double *out) {
  const int wg = p->w_geom, wv = p->w_sliding;
  const long nxg = p->xm + 2 * wg, nxv = p->xm + 2 * wv;
  double dt_max = max_dt_seconds, u_max = 0.0, v_max = 0.0;
  const double dx = p->dx, dy = p->dy;
  for (int j = p->ys; j < p->ys + p->ym; ++j) {
    for (int i = p->xs; i < p->xs + p->xm; ++i) {
      if (m_icy((int)floor(mask[(long)(j - (p->ys - wg)) * nxg + (i - (p->xs - wg))] + 0.5))) {
        const long o = ((long)(j - (p->ys - wv)) * nxv + (i - (p->xs - wv))) * 2;
        const double u_abs = velocity ? fabs(velocity[o]) : 0.0, v_abs = velocity ? fabs(velocity[o + 1]) : 0.0;
        u_max = std::max(u_max, u_abs);
        v_max = std::max(v_max, v_abs);
        const double denom = u_abs / dx + v_abs / dy;
        if (denom > 0.0) {
          dt_max = std::min(dt_max, 1.0 / denom);
        }
      }
    }
  }
  out[0] = dt_max;
  out[1] = u_max;
  out[2] = v_max;
  out[3] = 0.0;
  return ORC_OK;
}


// SIAFD_Regional::compute_surface_gradient, the override loop (src/regional/SIAFD_Regional.cc:63-116), SURVEY.md 8(f)
// N4.  no_model: the no_model_mask (2D w_geom, ghosts valid); hx_nm / hy_nm: the haseloff gradient of
// no_model_surface_elevation (Stag w_stag, ghosts valid; the caller computes it with orc_siafd_gradient on a
// field set whose surface is the stored one and whose method is haseloff, SIAFD_Regional.cc:52-55); h_x, h_y:
// the regular gradient, overridden in place on owned + 1.
int orc_regional_gradient_override(const orc_params *p, const double *no_model, const double *hx_nm,
                                   const double *hy_nm, double *h_x, double *h_y) {
  const int wg = p->w_geom, ws = p->w_stag, Mx = p->Mx, My = p->My;
  const long nxg = p->xm + 2 * wg, nxs = p->xm + 2 * ws;
  auto NM = [&](int i, int j) { return (int)floor(no_model[(long)(j - (p->ys - wg)) * nxg + (i - (p->xs - wg))] + 0.5); };
  auto S = [&](int i, int j, int o) { return ((long)(j - (p->ys - ws)) * nxs + (i - (p->xs - ws))) * 2 + o; };
  for (int j = p->ys - 1; j < p->ys + p->ym + 1; ++j) {
    for (int i = p->xs - 1; i < p->xs + p->xm + 1; ++i) {
      // BoxStencil (iceModelVec.hh:54-57): ij, n, nw, w, sw, s, se, e, ne
      const int ij = NM(i, j), n = NM(i, j + 1), nw = NM(i - 1, j + 1), w = NM(i - 1, j), s = NM(i, j - 1),
                se = NM(i + 1, j - 1), e = NM(i + 1, j), ne = NM(i + 1, j + 1);
      if (ij > 0.5 or e > 0.5) { // x-component, i-offset
        if (i < 0 or i + 1 > Mx - 1) {
          h_x[S(i, j, 0)] = 0.0;
        } else {
          h_x[S(i, j, 0)] = hx_nm[S(i, j, 0)];
        }
      }
      if (nw > 0.5 or ne > 0.5 or w > 0.5 or e > 0.5) { // x-component, j-offset
        if (i - 1 < 0 or j + 1 > My - 1 or i + 1 > Mx - 1) {
          h_x[S(i, j, 1)] = 0.0;
        } else {
          h_x[S(i, j, 1)] = hx_nm[S(i, j, 1)];
        }
      }
      if (n > 0.5 or ne > 0.5 or s > 0.5 or se > 0.5) { // y-component, i-offset
        if (i < 0 or j + 1 > My - 1 or i + 1 > Mx - 1 or j - 1 < 0) {
          h_y[S(i, j, 0)] = 0.0;
        } else {
          h_y[S(i, j, 0)] = hy_nm[S(i, j, 0)];
        }
      }
      if (ij > 0.5 or n > 0.5) { // y-component, j-offset
        if (j < 0 or j + 1 > My - 1) {
          h_y[S(i, j, 1)] = 0.0;
        } else {
          h_y[S(i, j, 1)] = hy_nm[S(i, j, 1)];
        }
      }
    }
  }
  return ORC_OK;
}

} // extern "C"
