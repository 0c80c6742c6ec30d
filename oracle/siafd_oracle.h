/*
 * siafd_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement ("oracle") of PISM v1.2.1's SIAFD hot path, used ONLY as the
 * checker in tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs.  Nothing under pism_b200/ may include, link or dlopen
 * this.  Every function cites the reference file:line it restates (paths are
 * relative to the reference tree, /root/reference in the build container).
 *
 * Parity pinning: see oracle/README.md.  Pinned by the reference's own
 * known-answer tests (flow-law table test/miscellaneous.py:634-671,
 * bed-smoother ranges test/bed_smoother.py:125-128, enthalpy-converter
 * identities test/enthalpy/converter.py) and by the reference's exact
 * solutions compiled from its own sources (oracle/_ref).  Per-field D / u / v
 * arrays have no checked-in golden values in the reference; for those this
 * restatement is the only comparator ("parity unpinned" for: eta and mahaffy
 * gradients, Vostok grain-size table, e_age_coupling time constants).
 */
#ifndef SIAFD_ORACLE_H
#define SIAFD_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Flow-law ids: keywords of stress_balance.sia.flow_law (src/pism_config.cdl:2093-2094,
 * src/rheology/FlowLawFactory.cc:71-87). */
enum {
  ORC_FLOW_ISOTHERMAL_GLEN = 0,
  ORC_FLOW_PB = 1,
  ORC_FLOW_GPBLD = 2,
  ORC_FLOW_HOOKE = 3,
  ORC_FLOW_ARR = 4,
  ORC_FLOW_ARRWARM = 5,
  ORC_FLOW_GK = 6
};

/* stress_balance.sia.surface_gradient_method (src/pism_config.cdl:2114-2115). */
enum { ORC_GRAD_HASELOFF = 0, ORC_GRAD_MAHAFFY = 1, ORC_GRAD_ETA = 2 };

/* Status codes: one per reference exception on the path. */
enum {
  ORC_OK = 0,
  ORC_ERR_NEGATIVE_THICKNESS = 1, /* BedSmoother.cc:303-305 */
  ORC_ERR_OMEGA_NEGATIVE = 2,     /* BedSmoother.cc:383-387 */
  ORC_ERR_HEIGHT_BELOW_BASE = 3,  /* IceGrid.cc:429-432 */
  ORC_ERR_HEIGHT_ABOVE_TOP = 4,   /* IceGrid.cc:434-437 */
  ORC_ERR_DIFFUSIVITY = 5,        /* SIAFD.cc:752-760 */
  ORC_ERR_BAD_CONFIG = 6          /* SIAFD.cc:216-219, :69-86, BedSmoother.cc:134-137 */
};

/* All physical constants the path reads (defaults: src/pism_config.cdl, lines in
 * SURVEY.md section 5.6).  Layout is mirrored by tests/oracle_lib.py. */
typedef struct orc_params {
  /* grid: IceGrid (src/util/IceGrid.hh) */
  int32_t Mx, My, Mz;     /* global sizes */
  int32_t xs, xm, ys, ym; /* owned patch of this rank (DMDA corners) */
  double dx, dy;
  const double *z; /* Mz vertical levels */
  /* ghost (stencil) widths of the local arrays */
  int32_t w_geom;    /* h, H, mask, bed, smoothed-bed fields: Geometry.cc:33-42 (2) */
  int32_t w_3d_in;   /* enthalpy, age (>= 2) */
  int32_t w_stag;    /* h_x, h_y, D, Q (1) */
  int32_t w_uv;      /* u, v (1) */
  int32_t w_sliding; /* sliding velocity */
  int32_t pad0;
  /* EnthalpyConverter (src/util/EnthalpyConverter.cc:55-69, cold variant :287-296) */
  double ec_p_air, ec_g, ec_beta, ec_rho_i, ec_c_i, ec_c_w, ec_L, ec_T_melting, ec_T_0;
  /* FlowLaw base (src/rheology/FlowLaw.cc:33-58) */
  int32_t flow_law;
  int32_t pad1;
  double fl_n, fl_e, fl_e_interglacial;
  double fl_A_cold, fl_A_warm, fl_Q_cold, fl_Q_warm, fl_T_crit;
  double fl_R, fl_rho, fl_g, fl_beta, fl_T_melting;
  double gpbld_T_0, gpbld_water_frac_coeff, gpbld_water_frac_limit; /* GPBLD.cc:36-39 */
  double iso_softness_A;                                            /* IsothermalGlen.cc:33 */
  double hooke_Q, hooke_A, hooke_C, hooke_K, hooke_Tr;              /* Hooke.cc:35-39 */
  double grain_size;                                                /* constants.ice.grain_size [m] */
  /* SIAFD (src/stressbalance/sia/SIAFD.cc:42-91, :555-570) */
  int32_t gradient_method;
  int32_t limit_diffusivity;
  int32_t grain_size_age_coupling;
  int32_t e_age_coupling;
  double D_limit;
  double eemian_start, eemian_end, holocene_start; /* seconds */
  double years_per_second;                         /* m_seconds_per_year, SIAFD.cc:58 */
  /* BedSmoother (BedSmoother.cc:74-75, :370) */
  double smoother_range, theta_min;
  /* GeometryCalculator (src/util/Mask.hh:71-79) */
  double sea_water_density, ice_free_thickness;
  int32_t dry_simulation;
  int32_t pad2;
} orc_params;

/* Local ghosted arrays in PISM's DMDA layout [j][i][dof], dof fastest
 * (src/util/IceModelVec_inline.hh:28-40; 3D = dof Mz, iceModelVec3.cc:85). */
typedef struct orc_fields {
  /* inputs */
  const double *surface;   /* h     2D  w_geom */
  const double *thickness; /* H     2D  w_geom */
  const double *mask;      /* cell type stored as double, w_geom */
  const double *bed;       /* topg  2D  w_geom (only eta gradient + smoother-off copy) */
  const double *enthalpy;  /* 3D w_3d_in */
  const double *age;       /* 3D w_3d_in or NULL */
  const double *sliding;   /* Vector2 {u,v} interleaved, w_sliding */
  /* bed smoother state (filled by orc_preprocess_bed), all w_geom */
  const double *topgsmooth, *maxtl, *C2, *C3, *C4;
  int32_t smoother_active; /* m_Nx >= 0 */
  int32_t pad;
  double current_time; /* seconds */
  /* outputs */
  double *h_x, *h_y; /* Stag w_stag */
  double *D, *Q;     /* Stag w_stag: diffusivity, diffusive_flux */
  double *u, *v;     /* 3D w_uv (owned points only; ghosts via orc_wrap_ghosts) */
  /* scratch the reference also keeps: 2D work (w_geom) and 3D work (w_stag) */
  double *work2d_0, *work2d_1;
  double *delta_0, *delta_1, *I_0, *I_1;
  /* scalar results */
  double D_max;
  int32_t high_diffusivity_counter;
  int32_t pad3;
} orc_fields;

void orc_default_params(orc_params *p);

/* rheology + converter scalars (FlowLaw.cc:97-105, EnthalpyConverter.cc) */
double orc_flow(const orc_params *p, double stress, double E, double pressure, double gs);
double orc_ec_pressure(const orc_params *p, double depth);
double orc_ec_melting_temperature(const orc_params *p, double P);
double orc_ec_enthalpy_cts(const orc_params *p, double P);
double orc_ec_temperature(const orc_params *p, double E, double P);
double orc_ec_pressure_adjusted_temperature(const orc_params *p, double E, double P);
double orc_ec_water_fraction(const orc_params *p, double E, double P);
double orc_ec_enthalpy(const orc_params *p, double T, double omega, double P);
double orc_ec_enthalpy_permissive(const orc_params *p, double T, double omega, double P);
double orc_grain_size_vostok(double age_years);

/* grid helpers (IceGrid.cc:381-499) */
void orc_vertical_levels(double Lz, int Mz, int quadratic, double lambda, double *z);
int orc_k_below_height(const double *z, int Mz, double height, int *status);
int orc_compute_nprocs(int Mx, int My, int size, int *Nx, int *Ny);
void orc_ownership_ranges(int M, int N, int *out);

/* Mask.hh:96-133 applied pointwise to n values. */
void orc_geometry_compute(const orc_params *p, int n, const double *sea_level, const double *bed,
                          const double *thickness, double *mask_out, double *surface_out);

/* BedSmoother::preprocess_bed (BedSmoother.cc:99-267) on GLOBAL Mx*My arrays
 * ([j][i], no ghosts); outputs global too.  Returns Nx, Ny through pointers. */
int orc_preprocess_bed(const orc_params *p, const double *topg, double *topgsmooth, double *maxtl,
                       double *C2, double *C3, double *C4, int *Nx_out, int *Ny_out);
/* BedSmoother::theta (BedSmoother.cc:351-404) and ::smoothed_thk (:284-327) on the
 * local patch incl. w_geom ghosts. */
int orc_theta(const orc_params *p, const orc_fields *f, double *theta_out);
int orc_smoothed_thk(const orc_params *p, const orc_fields *f, double *result);

/* Periodic self-wrap of the ghosts of a local array that covers the WHOLE domain
 * (single rank): DMDA is always periodic (IceGrid.cc:870-872). */
void orc_wrap_ghosts(int Mx, int My, int w, int dof, double *a);

/* SIAFD::update (SIAFD.cc:122-155) on one patch.  Ghosts of h_x/h_y (haseloff) and
 * u/v are NOT exchanged here; the caller does it (orc_wrap_ghosts on a single
 * patch, or patch-to-patch copies in the decomposition tests).  To make that
 * possible the update is split at the reference's communication points: */
int orc_siafd_gradient(const orc_params *p, orc_fields *f);                /* SIAFD.cc:137 (before :498) */
int orc_siafd_flux_velocity(const orc_params *p, orc_fields *f, int full); /* SIAFD.cc:141-153 (before :946) */
/* Convenience for a single whole-domain patch: gradient, wrap, flux+velocity, wrap. */
int orc_siafd_update_single(const orc_params *p, orc_fields *f, int full);
/* Many independent patches at once, OpenMP over patches (CPU baseline). */
/* StressBalance::compute_vertical_velocity (StressBalance.cc:283-424): mask w_geom, u / v w_uv (ghosts valid),
 * basal_melt_rate owned-only or NULL, w owned-only. */
int orc_vertical_velocity(const orc_params *p, const double *mask, const double *u, const double *v,
                          const double *basal_melt_rate, int use_upstream_fd, double *w);
/* IceModelVec3::getSurfaceValues / getHorSlice (util/iceModelVec3.cc:153-240): a (3D, ghost width wa) at the heights
 * in `heights` (2D, ghost width wh) or, with heights == NULL, at z0; out owned-only. */
void orc_value_at_height(const orc_params *p, const double *a, int wa, const double *heights, int wh, double z0,
                         double *out);
int orc_siafd_update_many(int n, const orc_params *p, orc_fields *f, int full, int nthreads);
/* SIAFD::update on the n patches of one domain (PISM's decomposition): the reference's passes, its two ghost updates
 * (SIAFD.cc:498-499, :946-947) as copies from the owning patch, D_max over all patches (:748); one OpenMP thread per
 * patch stands for one MPI rank. */
int orc_siafd_update_decomposed(int n, const orc_params *p, orc_fields *f, int full, int nthreads);
/* StressBalance::compute_volumetric_strain_heating (StressBalance.cc:426-642), SURVEY.md 8(f) N3: p carries the flow
 * law of the SHALLOW stress balance (id, fl_n, fl_e); thickness, mask w_geom; enthalpy w_3d_in; u, v w_uv (ghosts
 * valid); Sigma owned only. */
int orc_strain_heating(const orc_params *p, const double *thickness, const double *mask, const double *enthalpy,
                       const double *u, const double *v, double *Sigma_out);

/* SURVEY.md 8(f) N1 / N3 (mass_oracle.cc): GeometryEvolution::flow_step + apply_flux_divergence
 * (geometry/GeometryEvolution.cc:241-350), source_term_step + apply_mass_fluxes (:327-390, :1005-1076) and the CFL
 * reductions (stressbalance/timestepping.cc:42-153).  Array conventions in mass_oracle.cc. */
int orc_mass_flow_step(const orc_params *p, double dt, const double *sea_level, const double *bed, double *thickness,
                       const double *velocity, const double *velocity_bc_mask, const double *thickness_bc_mask,
                       const double *Q, double *flux_divergence, double *thickness_change,
                       double *conservation_error);
/* The same step with geometry.part_grid.enabled (GeometryEvolution.cc:689-944, part_grid_threshold_thickness.cc): used
 * only to pin the shared interface-flux code against test/mass_transport.py's golden numbers.  Whole-domain patch. */
int orc_mass_flow_step_part_grid(const orc_params *p, double dt, const double *sea_level, const double *bed,
                                 double *thickness, double *area_specific_volume, const double *velocity,
                                 const double *velocity_bc_mask, const double *thickness_bc_mask, const double *Q,
                                 int max_iterations, double *flux_divergence, double *thickness_change,
                                 double *area_specific_volume_change, double *conservation_error);
int orc_mass_source_step(const orc_params *p, double dt, double ice_density, int use_bmr, double *thickness,
                         const double *mask, const double *thickness_bc_mask, const double *smb_flux,
                         const double *basal_melt_rate, double *effective_SMB, double *effective_BMB);
int orc_cfl_3d(const orc_params *p, double max_dt_seconds, const double *thickness, const double *mask,
               const double *u3, const double *v3, const double *w3, double *out);
int orc_cfl_2d(const orc_params *p, double max_dt_seconds, const double *mask, const double *velocity, double *out);
/* SIAFD_Regional::compute_surface_gradient's override loop (regional/SIAFD_Regional.cc:63-116), SURVEY.md 8(f) N4. */
int orc_regional_gradient_override(const orc_params *p, const double *no_model, const double *hx_nm,
                                   const double *hy_nm, double *h_x, double *h_y);

#ifdef __cplusplus
}
#endif
#endif
