/*
 * siafd_b200.h -- C ABI of the B200-native SIAFD hot path.
 *
 * Drop-in boundary for PISM v1.2.1's `stressbalance::SIAFD` (an `SSB_Modifier`):
 * a C++ subclass in the host model marshals `SIAFD::update()` to these entry points
 * (INTEGRATION.md shows the shim).  Plain pointers and sizes only; no C++ or torch types.
 * All reference citations are file:line in the juliusgarbe/pism tree (v1.2.1).
 *
 *   reference interface                                       replaced by
 *   -------------------------------------------------------   ----------------------------------
 *   SIAFD::SIAFD(grid)        sia/SIAFD.cc:42-91               siafd_b200_create
 *   SIAFD::~SIAFD             sia/SIAFD.cc:93-95               siafd_b200_destroy
 *   SIAFD::init               sia/SIAFD.cc:98-118              siafd_b200_create (no separate state)
 *   BedSmoother::preprocess_bed  sia/BedSmoother.cc:99-153     siafd_b200_preprocess_bed /
 *                                                              siafd_b200_set_smoothed_bed
 *   SIAFD::update             sia/SIAFD.cc:122-155             siafd_b200_update  (one call), or the
 *                                                              split form  _upload / _compute_gradient /
 *                                                              _compute_flux_velocity / _download
 *   SSB_Modifier::max_diffusivity  SSB_Modifier.cc:69-71       siafd_b200_max_diffusivity
 *   SSB_Modifier::diffusive_flux / velocity_u / velocity_v     output pointers of siafd_b200_update
 *   SIAFD::surface_gradient_x/y, diffusivity  SIAFD.cc:963-973 output pointers (h_x, h_y, D)
 *   GeometryCalculator::compute  util/Mask.hh:96-133           siafd_b200_geometry_compute
 *   StressBalance::compute_vertical_velocity  StressBalance.cc:283-424   siafd_b200_compute_vertical_velocity
 *   max_timestep_cfl_3d / _2d  stressbalance/timestepping.cc:42-153      siafd_b200_cfl
 *   StressBalance::compute_volumetric_strain_heating  StressBalance.cc:426-642  siafd_b200_compute_strain_heating
 *   SIAFD_Regional::compute_surface_gradient  regional/SIAFD_Regional.cc:46-116
 *                                               siafd_b200_compute_gradient_no_model + _apply_no_model_gradient
 *   GeometryEvolution::flow_step + apply_flux_divergence  geometry/GeometryEvolution.cc:241-350
 *                                                              siafd_b200_mass_flow_step
 *   GeometryEvolution::source_term_step + apply_mass_fluxes  :327-390    siafd_b200_mass_source_step
 *   Geometry::ensure_consistency  geometry/Geometry.cc:121-187 siafd_b200_ensure_consistency
 *   IceModelVec3::getSurfaceValues / getHorSlice  util/iceModelVec3.cc:209-240
 *                                               siafd_b200_surface_values / siafd_b200_hor_slice
 *
 * Array layout is PISM's DMDA local (ghosted) layout, unchanged: [j][i][dof] with dof
 * fastest (util/IceModelVec_inline.hh:28-40); a 3D field is dof = Mz (util/iceModelVec3.cc:85);
 * IceModelVec2Stag is dof = 2 (offset o); IceModelVec2V is dof = 2 {u, v}.  The array for a
 * field with ghost width w covers i in [xs-w, xs+xm+w), j in [ys-w, ys+ym+w).
 * The cell-type mask is the reference's double-stored integer (IceModelVec_inline.hh:95-101).
 *
 * Threading: one host thread per handle; calls on a handle are stream-ordered and, unless
 * stated otherwise, synchronous at return.  No exception crosses this ABI: every entry
 * point returns a status code, and siafd_b200_last_error() has the message.
 */
#ifndef SIAFD_B200_H
#define SIAFD_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SIAFD_B200_ABI_VERSION 2

/* stress_balance.sia.flow_law keywords (pism_config.cdl:2093-2094, rheology/FlowLawFactory.cc:71-87) */
enum {
  SIAFD_B200_FLOW_ISOTHERMAL_GLEN = 0,
  SIAFD_B200_FLOW_PB = 1,
  SIAFD_B200_FLOW_GPBLD = 2,
  SIAFD_B200_FLOW_HOOKE = 3,
  SIAFD_B200_FLOW_ARR = 4,
  SIAFD_B200_FLOW_ARRWARM = 5,
  SIAFD_B200_FLOW_GK = 6
};

/* stress_balance.sia.surface_gradient_method (pism_config.cdl:2114-2115, sia/SIAFD.cc:197-220) */
enum { SIAFD_B200_GRAD_HASELOFF = 0, SIAFD_B200_GRAD_MAHAFFY = 1, SIAFD_B200_GRAD_ETA = 2 };

/* Status codes.  1..5 are the reference's RuntimeError conditions on this path. */
enum {
  SIAFD_B200_OK = 0,
  SIAFD_B200_ERR_NEGATIVE_THICKNESS = 1, /* sia/BedSmoother.cc:303-305 */
  SIAFD_B200_ERR_OMEGA_NEGATIVE = 2,     /* sia/BedSmoother.cc:383-387 */
  SIAFD_B200_ERR_HEIGHT_BELOW_BASE = 3,  /* util/IceGrid.cc:429-432 */
  SIAFD_B200_ERR_HEIGHT_ABOVE_TOP = 4,   /* util/IceGrid.cc:434-437 */
  SIAFD_B200_ERR_DIFFUSIVITY = 5,        /* sia/SIAFD.cc:752-760 */
  SIAFD_B200_ERR_BAD_CONFIG = 6,         /* sia/SIAFD.cc:69-86,216-219; BedSmoother.cc:134-137 */
  SIAFD_B200_ERR_CUDA = 7,               /* any CUDA runtime failure, incl. "no device" */
  SIAFD_B200_ERR_BAD_ARGUMENT = 8,
  SIAFD_B200_ERR_COMM = 9 /* a peer never arrived at a ghost update of a decomposed run */
};

/* Field ids for upload / download / bind / device_ptr / wrap_ghosts / halo pack. */
enum {
  SIAFD_B200_F_SURFACE = 0,   /* h,    2D, w_geom   (Geometry::ice_surface_elevation) */
  SIAFD_B200_F_THICKNESS = 1, /* H,    2D, w_geom */
  SIAFD_B200_F_MASK = 2,      /* cell_type as double, 2D, w_geom */
  SIAFD_B200_F_BED = 3,       /* bed_elevation, 2D, w_geom */
  SIAFD_B200_F_ENTHALPY = 4,  /* 3D, w_3d_in */
  SIAFD_B200_F_AGE = 5,       /* 3D, w_3d_in (only with age coupling) */
  SIAFD_B200_F_SLIDING = 6,   /* Vector2, w_sliding */
  SIAFD_B200_F_TOPGSMOOTH = 7,
  SIAFD_B200_F_MAXTL = 8,
  SIAFD_B200_F_C2 = 9,
  SIAFD_B200_F_C3 = 10,
  SIAFD_B200_F_C4 = 11, /* 7..11: BedSmoother state, 2D, w_geom */
  SIAFD_B200_F_H_X = 12,
  SIAFD_B200_F_H_Y = 13, /* Stag, w_stag */
  SIAFD_B200_F_D = 14,   /* Stag, w_stag: diffusivity */
  SIAFD_B200_F_FLUX = 15, /* Stag, w_stag: diffusive_flux */
  SIAFD_B200_F_U = 16,
  SIAFD_B200_F_V = 17,          /* 3D, w_uv */
  SIAFD_B200_F_THK_SMOOTH = 18, /* scratch 2D, w_geom (m_work_2d_0 in compute_diffusivity) */
  SIAFD_B200_F_THETA = 19,      /* scratch 2D, w_geom (m_work_2d_1) */
  SIAFD_B200_F_W_I = 20,
  SIAFD_B200_F_W_J = 21, /* scratch 2D, w_geom: haseloff weights / eta */
  SIAFD_B200_F_W = 22,          /* vertical velocity, 3D, no ghosts (StressBalance.cc:142) */
  SIAFD_B200_F_BASAL_MELT = 23, /* basal melt rate (Inputs::basal_melt_rate), 2D, no ghosts */
  /* 24..32: the mass-continuity consumer, geometry/GeometryEvolution.cc */
  SIAFD_B200_F_SEA_LEVEL = 24,   /* Geometry::sea_level_elevation, 2D, w_geom (never uploaded = 0) */
  SIAFD_B200_F_SMB = 25,         /* surface mass balance rate [kg m-2 s-1], 2D, no ghosts */
  SIAFD_B200_F_THK_CHANGE = 26,  /* GeometryEvolution::thickness_change_due_to_flow, 2D, no ghosts */
  SIAFD_B200_F_FLUX_DIV = 27,    /* GeometryEvolution::flux_divergence, 2D, no ghosts */
  SIAFD_B200_F_CONS_ERR = 28,    /* GeometryEvolution::conservation_error, 2D, no ghosts */
  SIAFD_B200_F_EFF_SMB = 29,     /* GeometryEvolution::top_surface_mass_balance [m], 2D, no ghosts */
  SIAFD_B200_F_EFF_BMB = 30,     /* GeometryEvolution::bottom_surface_mass_balance [m], 2D, no ghosts */
  SIAFD_B200_F_VEL_BC_MASK = 31, /* velocity Dirichlet B.C. mask, 2D, w_geom (never uploaded = 0) */
  SIAFD_B200_F_THK_BC_MASK = 32, /* thickness Dirichlet B.C. mask, 2D, w_geom (never uploaded = 0) */
  SIAFD_B200_F_STRAIN_HEATING = 33, /* StressBalance::volumetric_strain_heating, 3D, no ghosts */
  /* 34..37: SIAFD_Regional (regional/SIAFD_Regional.cc), SURVEY.md 8(f) N4 */
  SIAFD_B200_F_NO_MODEL_MASK = 34,    /* Inputs::no_model_mask, 2D, w_geom */
  SIAFD_B200_F_NO_MODEL_SURFACE = 35, /* Inputs::no_model_surface_elevation, 2D, w_geom */
  SIAFD_B200_F_H_X_NO_MODEL = 36,
  SIAFD_B200_F_H_Y_NO_MODEL = 37,     /* Stag, w_stag: SIAFD_Regional::m_h_x_no_model / m_h_y_no_model */
  SIAFD_B200_F_COUNT = 38
};

/* Everything SIAFD's constructor and update() read from Config/IceGrid
 * (defaults: siafd_b200_default_config, values from src/pism_config.cdl). */
typedef struct siafd_b200_config {
  /* IceGrid (util/IceGrid.hh): global sizes, this rank's owned patch, spacing, levels */
  int32_t Mx, My, Mz;
  int32_t xs, xm, ys, ym;
  double dx, dy;
  const double *z; /* Mz levels; copied by create */
  /* ghost widths of the caller's arrays */
  int32_t w_geom;    /* 2 (geometry/Geometry.cc:33-42) */
  int32_t w_3d_in;   /* >= 2 (sia/SIAFD.cc:587,602) */
  int32_t w_stag;    /* 1 */
  int32_t w_uv;      /* 1 */
  int32_t w_sliding; /* >= 0 */
  int32_t pad0;
  /* EnthalpyConverter (util/EnthalpyConverter.cc:55-69; cold mode :287-296 = T_melting 1e6, beta 0) */
  double ec_p_air, ec_g, ec_beta, ec_rho_i, ec_c_i, ec_c_w, ec_L, ec_T_melting, ec_T_0;
  /* FlowLaw (rheology/FlowLaw.cc:33-58) */
  int32_t flow_law;
  int32_t pad1;
  double fl_n, fl_e, fl_e_interglacial;
  double fl_A_cold, fl_A_warm, fl_Q_cold, fl_Q_warm, fl_T_crit;
  double fl_R, fl_rho, fl_g, fl_beta, fl_T_melting;
  double gpbld_T_0, gpbld_water_frac_coeff, gpbld_water_frac_limit;
  double iso_softness_A;
  double hooke_Q, hooke_A, hooke_C, hooke_K, hooke_Tr;
  double grain_size; /* m */
  /* SIAFD (sia/SIAFD.cc:555-570) */
  int32_t gradient_method;
  int32_t limit_diffusivity;
  int32_t grain_size_age_coupling;
  int32_t e_age_coupling;
  double D_limit;
  double eemian_start, eemian_end, holocene_start; /* s */
  double years_per_second;
  /* BedSmoother (sia/BedSmoother.cc:74-75,370) */
  double smoother_range, theta_min;
  /* GeometryCalculator (util/Mask.hh:71-79) */
  double sea_water_density, ice_free_thickness;
  int32_t dry_simulation;
  int32_t pad2;
} siafd_b200_config;

/* Host (or device, see memory_space) arrays of one update() call. NULL = not provided. */
typedef struct siafd_b200_inputs {
  const double *surface, *thickness, *mask, *bed; /* Inputs::geometry, StressBalance.hh:45 */
  const double *enthalpy;                         /* Inputs::enthalpy */
  const double *age;                              /* Inputs::age (may be NULL) */
  const double *sliding;                          /* sliding_velocity argument of update() */
  double current_time;                            /* grid->ctx()->time()->current(), SIAFD.cc:564 */
  int32_t memory_space;                           /* 0 = host pointers, 1 = device pointers */
  int32_t ghosts_valid;                           /* 1: caller filled ghosts (PISM always does);
                                                     0: single-rank whole-domain patch, library
                                                     wraps the input ghosts periodically itself */
} siafd_b200_inputs;

typedef struct siafd_b200_outputs {
  double *h_x, *h_y; /* SIAFD::surface_gradient_x/y */
  double *D;         /* SIAFD::diffusivity */
  double *flux;      /* SSB_Modifier::diffusive_flux */
  double *u, *v;     /* SSB_Modifier::velocity_u/v (full_update only) */
  int32_t memory_space;
  int32_t pad;
} siafd_b200_outputs;

typedef struct siafd_b200_handle siafd_b200_handle;

int siafd_b200_abi_version(void);
void siafd_b200_default_config(siafd_b200_config *cfg);
const char *siafd_b200_status_string(int status);

/* device < 0: use the current CUDA device.  Fails with SIAFD_B200_ERR_CUDA when no GPU; *out is NULL then.
 * Every other entry point accepts a NULL handle and reports it: status-returning calls give
 * SIAFD_B200_ERR_BAD_ARGUMENT (text through siafd_b200_last_error(NULL)), getters -1 / NaN / NULL. */
int siafd_b200_create(const siafd_b200_config *cfg, int device, siafd_b200_handle **out);
void siafd_b200_destroy(siafd_b200_handle *h);
const char *siafd_b200_last_error(const siafd_b200_handle *h);

/* Local array size (in doubles) and ghost width of a field for this handle's patch. */
int64_t siafd_b200_field_size(const siafd_b200_handle *h, int field);
int siafd_b200_field_width(const siafd_b200_handle *h, int field);
int siafd_b200_field_dof(const siafd_b200_handle *h, int field);

/* Device storage.  By default the handle owns one device buffer per field (allocated on
 * first use).  bind() makes the handle use caller-owned device memory (e.g. a torch tensor's
 * data_ptr) for that field instead; it must stay valid until re-bound or destroy. */
int siafd_b200_bind(siafd_b200_handle *h, int field, void *device_ptr);
void *siafd_b200_device_ptr(siafd_b200_handle *h, int field);
/* CUDA stream (cudaStream_t as void*) all work of this handle is enqueued on; NULL resets
 * to the handle's own stream. */
int siafd_b200_set_stream(siafd_b200_handle *h, void *cuda_stream);

int siafd_b200_upload(siafd_b200_handle *h, int field, const double *host);
int siafd_b200_download(siafd_b200_handle *h, int field, double *host);

/* Periodic self-wrap of a field's ghosts on device (single rank owning the whole domain in
 * the wrapped direction; util/IceGrid.cc:870-872: the DMDA is always periodic). */
int siafd_b200_wrap_ghosts(siafd_b200_handle *h, int field);
/* Several fields (1..6) in ONE launch: all four edge and four corner strips of each. */
int siafd_b200_wrap_ghosts_many(siafd_b200_handle *h, int n, const int *fields);
/* One direction only: dir 0 = x over the owned rows, dir 1 = y over all columns (x ghosts included).
 * A rank whose periodic neighbour in that direction is itself uses this in place of an exchange. */
int siafd_b200_wrap_ghosts_dir(siafd_b200_handle *h, int field, int dir);
/* Multi-rank halo exchange support: pack the owned strip that a neighbour needs into a
 * contiguous device buffer / unpack a received strip into the ghost region.
 * dir_x, dir_y in {-1,0,1} name the neighbour; width = ghost width to exchange (<= field
 * width).  `stage` 0 exchanges in x only (rows = owned rows), stage 1 exchanges in y
 * including the x ghosts, so that two stages fill BOX-stencil corners. */
int64_t siafd_b200_halo_count(const siafd_b200_handle *h, int field, int dir_x, int dir_y, int width);
int siafd_b200_halo_pack(siafd_b200_handle *h, int field, int dir_x, int dir_y, int width, double *device_buf);
int siafd_b200_halo_unpack(siafd_b200_handle *h, int field, int dir_x, int dir_y, int width, const double *device_buf);

/* Multi-rank halo exchange over peer memory (one process per GPU of one NVLink node): the neighbours' local
 * arrays are mapped with CUDA IPC, and a phase's strips (BOX stencil: 4 edges + 4 corners per field) are stored
 * straight into their ghost cells by ONE kernel; arrival counters in a small pad order the phases.
 *   dir = 0..7 numbers the neighbours (dx,dy) = (-1,-1),(0,-1),(1,-1),(-1,0),(1,0),(-1,1),(0,1),(1,1).
 * Setup: every rank exports the fields it exchanges (handle-owned storage only) and its pad (field = -1),
 * the 64-byte handles travel by any host channel, ipc_open maps a peer's handle once, halo_attach records the
 * mapped base (NULL = the neighbour is this rank itself) and the neighbour's patch size for a direction.
 * Per phase: halo_push(fields, widths, phase) then halo_wait(phase); both are stream-ordered, no host sync. */
int siafd_b200_ipc_export(siafd_b200_handle *h, int field, void *handle64);
int siafd_b200_ipc_open(siafd_b200_handle *h, const void *handle64, void **peer_ptr);
int siafd_b200_halo_attach(siafd_b200_handle *h, int field, int dir, void *peer_base, int peer_xm, int peer_ym);
int siafd_b200_halo_push(siafd_b200_handle *h, int n, const int *fields, const int *widths, int phase);
int siafd_b200_halo_wait(siafd_b200_handle *h, int phase);

/* Communicator of a decomposed run: one process per GPU of one NVLink node (or, for tests, several handles in one
 * process), no MPI / NCCL underneath.  It replaces IceModelVec::update_ghosts (util/iceModelVec.cc:630-643:
 * DMLocalToLocal on the periodic BOX-stencil DMDA, util/IceGrid.cc:863-885), GlobalMax / GlobalSum
 * (util/pism_utilities.cc:140-167; SIAFD.cc:748-750) and ParallelSection (util/error_handling.cc:189-214) for this path.
 *   comm_init: collective over the `size` ranks.  Every rank writes its patch and the CUDA IPC handles of its ghosted
 *     fields (handle-owned storage; allocated here if they are not yet) and of its pad to `<prefix>.rank.<rank>`, reads
 *     the others' files (any shared directory, e.g. /dev/shm; polls for up to timeout_seconds), maps its neighbours'
 *     arrays and every rank's pad, and leaves when all ranks have done so.  Neighbours are found from the patches
 *     (periodic in x and y); the decomposition must be a tensor product of 1D ranges, as PISM's is (IceGrid.cc:489-499).
 *     The prefix must be unique to this communicator; the files are removed by siafd_b200_destroy.
 *   comm_init_local: the same for `size` handles of ONE process (any devices with peer access; one device for tests).
 *   comm_exchange: update_ghosts of 1..6 fields at once: a "ready to receive" round (one single-CTA launch: a neighbour
 *     stores into this rank's ghost cells only after everything this rank enqueued before the call has run, e.g. an
 *     upload that rewrote the whole array), then ONE launch that stores the eight strips of every field straight into the
 *     neighbours' ghost cells (BOX corners included); its last CTA raises the neighbours' arrival counters and waits for
 *     this rank's own.  Stream-ordered, no host synchronisation.  Collective: all ranks make the same calls.
 *   comm_allreduce: max (op 0) / min (1) / sum in rank order (2) of 1..8 doubles over all ranks, through the pads;
 *     values in / out on the host; synchronises the stream. */
int siafd_b200_comm_init(siafd_b200_handle *h, int rank, int size, const char *rendezvous_prefix, double timeout_seconds);
int siafd_b200_comm_init_local(siafd_b200_handle **handles, int size);
int siafd_b200_comm_rank(const siafd_b200_handle *h);
int siafd_b200_comm_size(const siafd_b200_handle *h);
int siafd_b200_comm_exchange(siafd_b200_handle *h, int n, const int *fields, const int *widths);
int siafd_b200_comm_allreduce(siafd_b200_handle *h, int op, int n, double *values);
/* SIAFD::update (sia/SIAFD.cc:122-155) of one rank of a decomposed run on device-resident fields, every communication
 * of the reference inside: [exchange_inputs: ghosts of surface, thickness, mask, bed, enthalpy (age) -- device-resident
 * callers; under PISM the host arrays already carry them]; gradient, with the ghost update of h_x, h_y (:498-499) stored
 * by the gradient kernel itself; thk_smooth / theta; the fused diffusivity / flux / velocity kernel, which also stores
 * the rim of u, v into the neighbours' ghost cells (:946-947); one last single-CTA kernel that waits for the
 * neighbours' u, v and reduces {D_max, error bits, high-diffusivity counter} over ALL ranks (:748-750) into pinned host
 * memory.  4-6 launches, captured once as a CUDA graph and replayed.  Asynchronous; siafd_b200_finish waits and returns
 * the status, which is the same on every rank (ParallelSection), as are siafd_b200_max_diffusivity and
 * _high_diffusivity_count afterwards.  Works with a communicator of one rank (periodic self-wrap).
 * exchange_inputs = 1 has no "ready to receive" round (the step stays 4-6 launches): between two updates the caller must
 * change the input fields through their OWNED cells only (kernels of its own, or siafd_b200_upload with ghosts_valid
 * host arrays followed by exchange_inputs = 0) -- a neighbour that is already in its next update may be storing into
 * this rank's ghost cells. */
int siafd_b200_update_decomposed(siafd_b200_handle *h, int full_update, double current_time, int exchange_inputs);

/* BedSmoother::preprocess_bed on the GLOBAL bed (Mx*My doubles, [j][i], no ghosts; host
 * pointer).  Every rank passes the same array and gets its own patch (+ghosts) of
 * topgsmooth, maxtl, C2, C3, C4 on device.  Call when Inputs::new_bed_elevation. */
int siafd_b200_preprocess_bed(siafd_b200_handle *h, const double *global_bed_host);
/* Alternative: hand over the five local ghosted arrays PISM's own BedSmoother computed. */
int siafd_b200_set_smoothed_bed(siafd_b200_handle *h, const double *topgsmooth, const double *maxtl,
                                const double *C2, const double *C3, const double *C4, int smoother_active);

/* The update, split at the reference's two communication points (SIAFD.cc:498-499, :946-947)
 * so that a multi-rank caller can exchange ghosts in between.  All asynchronous on the
 * handle's stream. */
int siafd_b200_compute_gradient(siafd_b200_handle *h);                 /* SIAFD.cc:137 */
int siafd_b200_compute_flux_velocity(siafd_b200_handle *h, int full_update,
                                     double current_time);             /* SIAFD.cc:141-153 */
/* SURVEY.md 8(f) N2 -- StressBalance::compute_vertical_velocity (stressbalance/StressBalance.cc:283-424), the next
 * consumer of u, v: w from incompressibility, w(0) = -basal_melt_rate (0 without one), centered differences with
 * one-sided ones at ice margins, or first-order "upstream" ones (stress_balance.vertical_velocity_approximation).
 * Reads the handle's mask, u, v (ghosts valid, i.e. after the wrap / exchange of SIAFD.cc:946-947) and, if
 * use_basal_melt, SIAFD_B200_F_BASAL_MELT; writes SIAFD_B200_F_W.  Asynchronous on the handle's stream. */
int siafd_b200_compute_vertical_velocity(siafd_b200_handle *h, int upstream, int use_basal_melt);

/* SURVEY.md 8(f) N4 -- SIAFD_Regional::compute_surface_gradient (regional/SIAFD_Regional.cc:46-116), the regional
 * model's override of the protected phase SIAFD.hh:72-75.  After siafd_b200_compute_gradient and the wrap / exchange
 * of H_X, H_Y:  _compute_gradient_no_model = surface_gradient_haseloff on NO_MODEL_SURFACE (always haseloff, whatever
 * the configured method) into H_X_NO_MODEL / H_Y_NO_MODEL; the caller wraps / exchanges those two (width 1); then
 * _apply_no_model_gradient overrides H_X, H_Y next to NO_MODEL_MASK cells on owned + 1.  Asynchronous. */
int siafd_b200_compute_gradient_no_model(siafd_b200_handle *h);
int siafd_b200_apply_no_model_gradient(siafd_b200_handle *h);

/* SURVEY.md 8(f) N3 -- StressBalance::compute_volumetric_strain_heating (stressbalance/StressBalance.cc:426-642):
 * Sigma = 2 e^(-1/n) B(E, p) D2^((1/n + 1)/2) below the surface, 0 above, from the handle's THICKNESS, MASK, ENTHALPY,
 * U, V (ghosts valid); writes STRAIN_HEATING.  The reference takes the flow law of the SHALLOW stress balance here
 * (`stress_balance.ssa.flow_law`, `.Glen_exponent`, `.enhancement_factor`, even under ZeroSliding), so the law id,
 * n and e are arguments; the Paterson-Budd / GPBLD / Hooke constants are the handle's.  gk has no softness
 * (GoldsbyKohlstedt.cc:102-108): SIAFD_B200_ERR_BAD_CONFIG.  Asynchronous on the handle's stream. */
int siafd_b200_compute_strain_heating(siafd_b200_handle *h, int flow_law, double glen_exponent,
                                      double enhancement_factor);

/* IceModelVec3::getSurfaceValues (util/iceModelVec3.cc:226-240) and ::getHorSlice (:209-223), the reads of the
 * path's 3D outputs that its callers make: PISM.sia.computeSIASurfaceVelocities (site-packages/PISM/sia.py:63-72,
 * SURVEY.md 3.4) and siafd_test's surface-speed errors (sia/siafd_test.cc:105-151) take u, v at z = H; the velsurf /
 * velbase diagnostics take the same.  Per owned column: IceModelVec3D::getValZ (:153-182), i.e. the end levels
 * outside [z_0, z_{Mz-1}], else linear interpolation between the levels around the height, in the reference's
 * expression order (bit-identical to the CPU's).  field3d is one of ENTHALPY, AGE, U, V, W, STRAIN_HEATING;
 * _surface_values evaluates at the handle's THICKNESS, _hor_slice at the constant height z.  out_dev: DEVICE pointer
 * to ym * xm doubles, [j][i] without ghosts (an IceModelVec2S created WITHOUT_GHOSTS).  Asynchronous on the handle's
 * stream. */
int siafd_b200_surface_values(siafd_b200_handle *h, int field3d, double *out_dev);
int siafd_b200_hor_slice(siafd_b200_handle *h, int field3d, double z, double *out_dev);

/* SURVEY.md 8(f) N1 -- the consumer of diffusive_flux(): GeometryEvolution (geometry/GeometryEvolution.cc), default
 * configuration (geometry.part_grid.enabled = no).  All asynchronous on the handle's stream, device-resident:
 *   mass_flow_step  = flow_step(geometry, dt, advective_velocity, diffusive_flux, bc masks) :241-324 followed by
 *                     apply_flux_divergence :347-350.  Reads THICKNESS, BED, [SEA_LEVEL] (ghosts valid), FLUX (ghosts
 *                     valid: SIAFD computes it on owned + 1), SLIDING as the advective velocity (ghost width >= 1 with
 *                     valid ghosts, else it must be all zero = ZeroSliding), [VEL_BC_MASK], [THK_BC_MASK]; writes
 *                     FLUX_DIV, THK_CHANGE, CONS_ERR and updates THICKNESS on the owned points.
 *   mass_source_step = source_term_step :327-343 + apply_mass_fluxes :360-390.  Reads MASK, SMB, [BASAL_MELT if
 *                     use_basal_melt], [THK_BC_MASK]; writes EFF_SMB, EFF_BMB, updates THICKNESS on the owned points.
 *   ensure_consistency = Geometry::ensure_consistency (geometry/Geometry.cc:121-187) with the handle's
 *                     ice_free_thickness: MASK and SURFACE from THICKNESS, BED, [SEA_LEVEL] on every local point.
 *                     wrap_thickness = 1 first fills THICKNESS's ghosts by periodic self-wrap (single rank); a
 *                     multi-rank caller exchanges them itself and passes 0.  H < 0 raises
 *                     SIAFD_B200_ERR_NEGATIVE_THICKNESS at the next siafd_b200_finish. */
int siafd_b200_mass_flow_step(siafd_b200_handle *h, double dt);
int siafd_b200_mass_source_step(siafd_b200_handle *h, double dt, double ice_density, int use_basal_melt);
int siafd_b200_ensure_consistency(siafd_b200_handle *h, int wrap_thickness);
/* SURVEY.md 8(f) N3 (CFL part) -- max_timestep_cfl_3d / _2d (stressbalance/timestepping.cc:42-101, :113-153) on the
 * handle's THICKNESS, MASK, U, V, W and SLIDING.  out[0..3] = {dt_max, u_max, v_max, w_max} of the 3D criterion (only
 * if do_3d), out[4..7] = {dt_max, u_max, v_max, 0} of the 2D one; local values: multi-rank callers reduce them
 * (min, max, max, max).  max_dt_seconds = time_stepping.maximum_time_step.  Synchronises the stream. */
int siafd_b200_cfl(siafd_b200_handle *h, double max_dt_seconds, int do_3d, double *out8);

/* Waits for the stream, then evaluates the reference's error conditions; returns the status
 * (collective callers reduce it over ranks before acting on it, cf. ParallelSection). */
int siafd_b200_finish(siafd_b200_handle *h);
/* Local maximum of D over owned+ghost staggered points of the last update (SIAFD.cc:729);
 * multi-rank callers take the max over ranks (SIAFD.cc:748).  Synchronises the stream. */
double siafd_b200_max_diffusivity(siafd_b200_handle *h);
int siafd_b200_high_diffusivity_count(siafd_b200_handle *h);

/* One-call drop-in form: upload inputs (if host), gradient, [wrap h_x,h_y], flux+velocity,
 * [wrap u,v], download outputs (if host), finish.  With ghosts_valid = 1 and more than one
 * rank the caller must use the split form instead (the wraps become exchanges). */
int siafd_b200_update(siafd_b200_handle *h, const siafd_b200_inputs *in, siafd_b200_outputs *out, int full_update);

/* GeometryCalculator::compute (util/Mask.hh:96-133) for n points, on device memory. */
int siafd_b200_geometry_compute(siafd_b200_handle *h, int64_t n, const double *sea_level_dev, const double *bed_dev,
                                const double *thickness_dev, double *mask_out_dev, double *surface_out_dev);

/* FlowLaw::flow for n points on device (rheology/FlowLaw.cc:97-105); known-answer tests. */
int siafd_b200_flow_n(siafd_b200_handle *h, int64_t n, const double *stress_dev, const double *enthalpy_dev,
                      const double *pressure_dev, const double *grainsize_dev, double *result_dev);

/* The same for HOST arrays (staged through the device): SSB_Modifier::flow_law()->flow(...) for host-side diagnostics
 * (IceCompModel::reportErrors, verification/iceCompModel.cc:642).  grainsize may be NULL (the configured constant). */
int siafd_b200_flow_host(siafd_b200_handle *h, int64_t n, const double *stress, const double *enthalpy,
                         const double *pressure, const double *grainsize, double *result);

/* Kernel tuning knobs, for benchmarking: rows_per_cta > 0 sets the row-segment length one CTA
 * marches over (0 keeps it); use_bulk_copy / skip_ice_free_rows: 0 or 1 sets, -1 keeps. */
int siafd_b200_set_tuning(siafd_b200_handle *h, int rows_per_cta, int use_bulk_copy, int skip_ice_free_rows);
/* Number of kernel launches issued by this handle since create (bench.py's gpu_launches). */
int64_t siafd_b200_launch_count(const siafd_b200_handle *h);
/* Bytes the host-pointer calls (upload / download / update with host arrays) have moved over PCIe since create.
 * siafd_b200_update with host arrays moves only the parts of the 3D arrays within 3 cells of ice (no enthalpy is read
 * and u = v = sliding velocity elsewhere, SIAFD.cc:631-637, :935-942: those parts of u, v are filled on the host);
 * afterwards the DEVICE copy of the enthalpy is only current in those parts (with the level cut below: on the levels
 * up to the cut), which is all that the path and its device-resident consumers read (strain heating stops at the
 * surface); a caller that wants the whole field on the device uploads it (siafd_b200_upload). */
int siafd_b200_transfer_bytes(const siafd_b200_handle *h, int64_t *h2d, int64_t *d2h);
/* The level cut of that call (single rank, bed smoother off): of a column near ice only the levels [0, n) cross PCIe,
 * n = siafd_b200_host_levels_needed(z, Mz, T) with T the largest thk_smooth (BedSmoother.cc:306-320 with the
 * smoother off: 0 without ice, max(H, usurf - topg) where grounded, H where floating) over the column and the columns
 * next to it.  A staggered point reads the enthalpy on the levels k <= ks = kBelowHeight(thk) only
 * (SIAFD.cc:613-627, IceGrid.cc:427-440) and I is constant above ks (SIAFD.cc:857-859), so u and v (SIAFD.cc:935-942)
 * are constant from level n - 1 up: the host replicates that value, and the host arrays are bit-identical to a full
 * transfer.  Pure host arithmetic (no GPU needed): returns Mz when nothing can be cut.  Environment:
 * SIAFD_B200_LEVEL_CUT = 0 (off) / 1 (default: single rank) / 2 (also with several ranks),
 * SIAFD_B200_CUT_COLS (columns that share one n, default 256), SIAFD_B200_CUT_ROWS (rows that share one n, default 0
 * = the rows of a band), SIAFD_B200_REPL_THREADS (default 8); SIAFD_B200_TRACE=1 prints the timeline of each call;
 * SIAFD_B200_ZERO_COPY=1 (default 0): with u, v in pinned host memory a kernel stores the pieces there itself instead of
 * the copy engine (same bits, measured slower: DESIGN.md section 7). */
int siafd_b200_host_levels_needed(const double *z, int Mz, double max_thickness);
/* Dry run of what siafd_b200_update with host arrays moves and fills, on HOST arrays only (no GPU; for tests of the
 * host logic): the same plan (pism_b200/csrc/siafd_hostplan.hh) -- row bands of `band` segments of rows_per_segment
 * rows, sparse rectangles, level cut (level_cut = 1; 1 + k: chunks of k rows) with cut_cols columns per chunk, patch = 1
 * for one patch of a decomposed domain -- executed with memcpy.  It computes nothing of the update (this is not a CPU
 * path: the "device" arrays are the caller's).  enthalpy_dev (in/out) stands for the device copy of the enthalpy: it receives what would be
 * uploaded and nothing else; u_dev, v_dev stand for the device's result (ghosts valid); u, v receive what the call would
 * leave in the host arrays (downloaded pieces, host fills from `sliding` (may be NULL = zero), values replicated above
 * the cut, ghost rows / columns).  All arrays in the local ghosted layout of cfg.  h2d / d2h (may be NULL): bytes of the
 * 3D arrays that would cross PCIe. */
int siafd_b200_host_plan_emulate(const siafd_b200_config *cfg, int rows_per_segment, int band, int sparse, int level_cut,
                                 int cut_cols, int patch, const double *thickness, const double *surface, const double *bed,
                                 const double *mask, const double *sliding, const double *enthalpy, double *enthalpy_dev,
                                 const double *u_dev, const double *v_dev, double *u, double *v, int64_t *h2d, int64_t *d2h);
/* CUDA-event timing of the fused kernel alone, on the handle's stream (bench.py's roofline):
 * enable, run updates (<= 256), then read the accumulated milliseconds and launch count. */
int siafd_b200_kernel_timing(siafd_b200_handle *h, int enable);
double siafd_b200_kernel_time_ms(siafd_b200_handle *h, int *launches_out);
/* In the same mode siafd_b200_update_decomposed also times its sections (ungraphed launches, up to 64 steps): average
 * milliseconds of {ghosts of the 2D inputs, gradient + thk_smooth/theta pass, wait before the fused kernel (ghosts of
 * the 3D inputs on their side stream; the gradient's ghost update where it is not computed locally), fused kernel,
 * arrival of u, v + all-rank reduction of D_max / status / counter} -- what bench.py reports as "step_breakdown_ms". */
int siafd_b200_step_breakdown_ms(siafd_b200_handle *h, double *out5, int *steps_out);

#ifdef __cplusplus
}
#endif
#endif /* SIAFD_B200_H */
