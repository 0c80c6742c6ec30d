// siafd_capi.cu -- the C ABI declared in include/siafd_b200.h.
//
// Host-side orchestration of SIAFD::update (sia/SIAFD.cc:122-155): device buffers in PISM's
// local ghosted layout, one CUDA stream per handle, error flags and D_max reduced on device and
// read back once per update.  No CPU fallback: without a CUDA device every entry point that
// needs one fails with SIAFD_B200_ERR_CUDA.
#include "siafd_handle.cuh"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#if defined(__x86_64__) || defined(_M_X64)
#include <emmintrin.h>
#endif
#include <cstdint>
#include <string>
#include <thread>
#include <vector>

using namespace siafd;

namespace siafd_host {

thread_local std::string g_create_error;

int fail(siafd_b200_handle *h, int code, const char *fmt, ...) {
  char msg[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(msg, sizeof(msg), fmt, ap);
  va_end(ap);
  if (h) {
    h->err = msg;
  } else {
    g_create_error = msg;
  }
  return code;
}

// every entry point tolerates a NULL handle (e.g. after a failed siafd_b200_create): status codes report
// ERR_BAD_ARGUMENT (message through siafd_b200_last_error(NULL)), getters return -1 / NAN / NULL
int null_handle() { return fail(nullptr, SIAFD_B200_ERR_BAD_ARGUMENT, "NULL handle"); }

FieldMeta meta(const siafd_b200_config &c, int f) {
  switch (f) {
  case SIAFD_B200_F_SURFACE:
  case SIAFD_B200_F_THICKNESS:
  case SIAFD_B200_F_MASK:
  case SIAFD_B200_F_BED:
  case SIAFD_B200_F_TOPGSMOOTH:
  case SIAFD_B200_F_MAXTL:
  case SIAFD_B200_F_C2:
  case SIAFD_B200_F_C3:
  case SIAFD_B200_F_C4:
  case SIAFD_B200_F_THK_SMOOTH:
  case SIAFD_B200_F_THETA:
  case SIAFD_B200_F_W_I:
  case SIAFD_B200_F_W_J:
  case SIAFD_B200_F_SEA_LEVEL:
  case SIAFD_B200_F_NO_MODEL_MASK:
  case SIAFD_B200_F_NO_MODEL_SURFACE:
  case SIAFD_B200_F_VEL_BC_MASK:
  case SIAFD_B200_F_THK_BC_MASK:
    return {c.w_geom, 1};
  case SIAFD_B200_F_ENTHALPY:
  case SIAFD_B200_F_AGE:
    return {c.w_3d_in, c.Mz};
  case SIAFD_B200_F_SLIDING:
    return {c.w_sliding, 2};
  case SIAFD_B200_F_H_X:
  case SIAFD_B200_F_H_Y:
  case SIAFD_B200_F_D:
  case SIAFD_B200_F_FLUX:
  case SIAFD_B200_F_H_X_NO_MODEL:
  case SIAFD_B200_F_H_Y_NO_MODEL:
    return {c.w_stag, 2};
  case SIAFD_B200_F_U:
  case SIAFD_B200_F_V:
    return {c.w_uv, c.Mz};
  case SIAFD_B200_F_W:
  case SIAFD_B200_F_STRAIN_HEATING:
    return {0, c.Mz};
  case SIAFD_B200_F_BASAL_MELT:
  case SIAFD_B200_F_SMB:
  case SIAFD_B200_F_THK_CHANGE:
  case SIAFD_B200_F_FLUX_DIV:
  case SIAFD_B200_F_CONS_ERR:
  case SIAFD_B200_F_EFF_SMB:
  case SIAFD_B200_F_EFF_BMB:
    return {0, 1};
  default:
    return {-1, 0};
  }
}

int64_t field_cells(const siafd_b200_config &c, int w) { return (int64_t)(c.xm + 2 * w) * (c.ym + 2 * w); }

int ensure(siafd_b200_handle *h, int f) {
  if (f < 0 || f >= SIAFD_B200_F_COUNT) {
    return fail(h, SIAFD_B200_ERR_BAD_ARGUMENT, "bad field id %d", f);
  }
  if (h->buf[f]) {
    return SIAFD_B200_OK;
  }
  const FieldMeta m = meta(h->cfg, f);
  // +2 doubles of slack: bulk row copies are issued on 16-byte boundaries and may touch one
  // double past the last element (pism_b200/csrc/siafd_kernels.cu, issue_row)
  const size_t bytes = (size_t)(field_cells(h->cfg, m.width) * m.dof + 2) * sizeof(double);
  CU(h, cudaMalloc(&h->buf[f], bytes));
  // PETSc Vecs start zeroed and the reference relies on it (maxtl with the smoother off, G4)
  CU(h, cudaMemsetAsync(h->buf[f], 0, bytes, h->stream));
  h->owned[f] = true;
  return SIAFD_B200_OK;
}

Fields fields_of(siafd_b200_handle *h) {
  Fields F;
  auto D = [&](int f) { return (double *)h->buf[f]; };
  F.h = D(SIAFD_B200_F_SURFACE);
  F.H = D(SIAFD_B200_F_THICKNESS);
  F.mask = D(SIAFD_B200_F_MASK);
  F.bed = D(SIAFD_B200_F_BED);
  F.E = D(SIAFD_B200_F_ENTHALPY);
  F.age = D(SIAFD_B200_F_AGE);
  F.sliding = D(SIAFD_B200_F_SLIDING);
  // smoother off: topgsmooth is a ghosted copy of the bed (sia/BedSmoother.cc:101-109) -- read the bed itself
  F.topgsmooth = (!(h->cfg.smoother_range > 0.0) && !h->smoother_set) ? D(SIAFD_B200_F_BED) : D(SIAFD_B200_F_TOPGSMOOTH);
  F.maxtl = D(SIAFD_B200_F_MAXTL);
  F.C2 = D(SIAFD_B200_F_C2);
  F.C3 = D(SIAFD_B200_F_C3);
  F.C4 = D(SIAFD_B200_F_C4);
  F.thk_smooth = D(SIAFD_B200_F_THK_SMOOTH);
  F.theta = D(SIAFD_B200_F_THETA);
  F.w_i = D(SIAFD_B200_F_W_I);
  F.w_j = D(SIAFD_B200_F_W_J);
  F.h_x = D(SIAFD_B200_F_H_X);
  F.h_y = D(SIAFD_B200_F_H_Y);
  F.D = D(SIAFD_B200_F_D);
  F.Q = D(SIAFD_B200_F_FLUX);
  F.u = D(SIAFD_B200_F_U);
  F.v = D(SIAFD_B200_F_V);
  F.z = h->d_z;
  F.err = h->d_err;
  F.dmax = h->d_dmax;
  F.hdc = h->d_hdc;
  F.segw = h->d_segw;
  F.segdone = h->d_segdone;
  return F;
}

int check_config(const siafd_b200_config &c, std::string &why) {
  char b[256];
  if (c.Mx < 2 || c.My < 2 || c.Mz < 2 || c.xm < 1 || c.ym < 1 || c.xs < 0 || c.ys < 0 || c.xs + c.xm > c.Mx ||
      c.ys + c.ym > c.My || c.z == nullptr || !(c.dx > 0) || !(c.dy > 0)) {
    why = "invalid grid description";
    return SIAFD_B200_ERR_BAD_CONFIG;
  }
  if (c.w_geom < 2 || c.w_3d_in < 2 || c.w_stag < 1 || c.w_uv < 1 || c.w_sliding < 0) {
    // sia/SIAFD.cc:587-602 asserts
    why = "stencil widths too small (need geometry>=2, enthalpy>=2, staggered>=1, u/v>=1)";
    return SIAFD_B200_ERR_BAD_CONFIG;
  }
  if (c.gradient_method < 0 || c.gradient_method > 2) {
    // sia/SIAFD.cc:216-219
    snprintf(b, sizeof(b), "value of sia.surface_gradient_method (%d) is not valid", c.gradient_method);
    why = b;
    return SIAFD_B200_ERR_BAD_CONFIG;
  }
  if (c.flow_law < 0 || c.flow_law > 6) {
    // rheology/FlowLawFactory.cc:96-103
    snprintf(b, sizeof(b), "Selected ice flow law (%d) is not available", c.flow_law);
    why = b;
    return SIAFD_B200_ERR_BAD_CONFIG;
  }
  if (c.grain_size_age_coupling && c.flow_law != SIAFD_B200_FLOW_GK) {
    // sia/SIAFD.cc:69-76
    why = "flow law does not use grain size but sia.grain_size_age_coupling was set";
    return SIAFD_B200_ERR_BAD_CONFIG;
  }
  for (int k = 1; k < c.Mz; ++k) {
    if (!(c.z[k] > c.z[k - 1])) {
      why = "vertical levels must be strictly increasing";
      return SIAFD_B200_ERR_BAD_CONFIG;
    }
  }
  return SIAFD_B200_OK;
}

void fill_dp(siafd_b200_handle *h) {
  const siafd_b200_config &c = h->cfg;
  DP &P = h->P;
  std::memset(&P, 0, sizeof(P));
  P.Mx = c.Mx, P.My = c.My, P.Mz = c.Mz;
  P.xs = c.xs, P.xm = c.xm, P.ys = c.ys, P.ym = c.ym;
  P.wg = c.w_geom, P.we = c.w_3d_in, P.wst = c.w_stag, P.wuv = c.w_uv, P.wsl = c.w_sliding;
  P.dx = c.dx, P.dy = c.dy;
  P.inv_dx = 1.0 / c.dx, P.inv_dy = 1.0 / c.dy;
  P.p_air = c.ec_p_air;
  P.rg = c.ec_rho_i * c.ec_g; // first product of "m_rho_i * m_g * depth", EnthalpyConverter.cc:150
  P.ec_beta = c.ec_beta, P.c_i = c.ec_c_i, P.inv_c_i = 1.0 / c.ec_c_i, P.c_w = c.ec_c_w, P.L0 = c.ec_L;
  P.T_melting = c.ec_T_melting, P.T_0 = c.ec_T_0;
  P.law = c.flow_law;
  P.n = c.fl_n, P.nm1 = c.fl_n - 1, P.n_is_3 = (c.fl_n == 3.0);
  P.e = c.fl_e, P.e_inter = c.fl_e_interglacial;
  P.A_cold = c.fl_A_cold, P.A_warm = c.fl_A_warm, P.Q_cold = c.fl_Q_cold, P.Q_warm = c.fl_Q_warm;
  P.T_crit = c.fl_T_crit, P.R = c.fl_R;
  P.QoR_cold = c.fl_Q_cold / c.fl_R, P.QoR_warm = c.fl_Q_warm / c.fl_R;
  P.lnA_cold = log(c.fl_A_cold), P.lnA_warm = log(c.fl_A_warm);
  {
    const double c16 = 16.0 / log(2.0);
    P.lnA2_cold = P.lnA_cold * c16, P.lnA2_warm = P.lnA_warm * c16;
    P.QoR2_cold = P.QoR_cold * c16, P.QoR2_warm = P.QoR_warm * c16;
  }
  P.hic = 0.5 / c.ec_c_i;
  P.cts2_a = 2.0 * c.ec_c_i * (c.ec_T_melting - c.ec_T_0), P.cts2_b = 2.0 * c.ec_c_i * c.ec_beta;
  {
    // rheology/FlowLaw.cc:45 and PatersonBudd.cc:57, evaluated like the reference does
    const double beta_CC_grad = c.fl_beta * c.fl_rho * c.fl_g;
    P.beta_ratio = beta_CC_grad / (c.fl_rho * c.fl_g);
  }
  P.gp_T0 = c.gpbld_T_0, P.gp_coeff = c.gpbld_water_frac_coeff, P.gp_limit = c.gpbld_water_frac_limit;
  {
    // softness_paterson_budd(T_0), rheology/FlowLaw.cc:89-94: a constant, evaluated once on the host
    const double T = c.gpbld_T_0;
    const double A = T < c.fl_T_crit ? c.fl_A_cold : c.fl_A_warm, Q = T < c.fl_T_crit ? c.fl_Q_cold : c.fl_Q_warm;
    P.gp_softness_T0 = A * exp(-Q / (c.fl_R * T));
  }
  P.iso_A = c.iso_softness_A;
  P.hk_Q = c.hooke_Q, P.hk_A = c.hooke_A, P.hk_C = c.hooke_C, P.hk_K = c.hooke_K, P.hk_Tr = c.hooke_Tr;
  P.grain_size = c.grain_size;
  P.limit_diffusivity = c.limit_diffusivity;
  P.gs_age = c.grain_size_age_coupling, P.e_age = c.e_age_coupling;
  P.use_age = (c.grain_size_age_coupling || c.e_age_coupling) ? 1 : 0;
  P.D_limit = c.D_limit;
  P.eemian_start = c.eemian_start, P.eemian_end = c.eemian_end, P.holocene_start = c.holocene_start;
  P.years_per_second = c.years_per_second;
  P.smoother_active = 0;
  P.grad = c.gradient_method;
  P.theta_min = c.theta_min;
  P.gc_alpha = 1 - c.ec_rho_i / c.sea_water_density; // util/Mask.hh:72
  P.gc_icefree = c.ice_free_thickness;
  P.gc_dry = c.dry_simulation;
}

// periodic self-wrap in one direction (0: x over owned rows, 1: y over all columns)
int wrap_dir(siafd_b200_handle *h, int f, int dir) {
  const FieldMeta m = meta(h->cfg, f);
  const siafd_b200_config &c = h->cfg;
  const int w = m.width;
  if (w <= 0) {
    return SIAFD_B200_OK;
  }
  double *a = (double *)h->buf[f];
  const long rowc = c.xm + 2 * w;
  if (dir == 0) {
    if (c.xm != c.Mx) {
      return fail(h, SIAFD_B200_ERR_BAD_ARGUMENT, "x self-wrap needs a patch spanning the whole x range");
    }
    // west ghosts <- east owned columns; east ghosts <- west owned columns
    h->launches += launch_copy_region(a, rowc, 0, w, a, rowc, c.xm, w, w, c.ym, m.dof, h->stream);
    h->launches += launch_copy_region(a, rowc, c.xm + w, w, a, rowc, w, w, w, c.ym, m.dof, h->stream);
  } else {
    if (c.ym != c.My) {
      return fail(h, SIAFD_B200_ERR_BAD_ARGUMENT, "y self-wrap needs a patch spanning the whole y range");
    }
    h->launches += launch_copy_region(a, rowc, 0, 0, a, rowc, 0, c.ym, (int)rowc, w, m.dof, h->stream);
    h->launches += launch_copy_region(a, rowc, 0, c.ym + w, a, rowc, 0, w, (int)rowc, w, m.dof, h->stream);
  }
  CU(h, cudaGetLastError());
  return SIAFD_B200_OK;
}

int status_from_bits(unsigned bits) {
  if (bits & EB_NEG_THK) return SIAFD_B200_ERR_NEGATIVE_THICKNESS;
  if (bits & EB_OMEGA) return SIAFD_B200_ERR_OMEGA_NEGATIVE;
  if (bits & EB_BELOW) return SIAFD_B200_ERR_HEIGHT_BELOW_BASE;
  if (bits & EB_ABOVE) return SIAFD_B200_ERR_HEIGHT_ABOVE_TOP;
  if (bits & EB_COMM) return SIAFD_B200_ERR_COMM;
  return SIAFD_B200_OK;
}

void invalidate_graphs(siafd_b200_handle *h) {
  for (int q = 0; q < 4; ++q) h->comm.graph_valid[q] = false;
}

int fetch_result(siafd_b200_handle *h) {
  if (h->comm.result_from_comm && !h->result_pending) {
    // siafd_b200_update_decomposed: the reduction over all ranks was written to pinned host memory by the last kernel
    CU(h, cudaStreamSynchronize(h->stream));
    h->h_res->dmax = h->comm.h_res[0];
    h->h_res->err = (unsigned)h->comm.h_res[1];
    h->h_res->hdc = (int)h->comm.h_res[2];
    h->comm.result_from_comm = false;
    return SIAFD_B200_OK;
  }
  h->comm.result_from_comm = false;
  if (h->result_pending) {
    // (several ranks: D_max and the counter of the last update stay the reduced, global ones; the device copies are
    // this rank's own.  Only the error bits a later call raised are picked up here.)
    if (!(h->comm.active && h->comm.size > 1)) {
      CU(h, cudaMemcpyAsync(&h->h_res->dmax, h->d_dmax, sizeof(unsigned long long), cudaMemcpyDeviceToHost, h->stream));
      CU(h, cudaMemcpyAsync(&h->h_res->hdc, h->d_hdc, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    }
    CU(h, cudaMemcpyAsync(&h->h_res->err, h->d_err, sizeof(unsigned), cudaMemcpyDeviceToHost, h->stream));
    // error bits are sticky on the device until they have been read: an error raised by a call that follows an
    // update (ensure_consistency, cfl) is reported by the next finish, not erased by the next update
    CU(h, cudaMemsetAsync(h->d_err, 0, sizeof(unsigned), h->stream));
    h->result_pending = false;
  }
  CU(h, cudaStreamSynchronize(h->stream));
  return SIAFD_B200_OK;
}

} // namespace siafd_host

using namespace siafd_host;

extern "C" {

int siafd_b200_abi_version(void) { return SIAFD_B200_ABI_VERSION; }

// defaults: src/pism_config.cdl (line numbers in SURVEY.md section 5.6)
void siafd_b200_default_config(siafd_b200_config *c) {
  std::memset(c, 0, sizeof(*c));
  c->w_geom = 2, c->w_3d_in = 2, c->w_stag = 1, c->w_uv = 1, c->w_sliding = 1;
  c->ec_p_air = 0.0, c->ec_g = 9.81, c->ec_beta = 7.9e-8, c->ec_rho_i = 910.0, c->ec_c_i = 2009.0;
  c->ec_c_w = 4170.0, c->ec_L = 3.34e5, c->ec_T_melting = 273.15, c->ec_T_0 = 223.15;
  c->flow_law = SIAFD_B200_FLOW_GPBLD;
  c->fl_n = 3.0, c->fl_e = 1.0, c->fl_e_interglacial = 1.0;
  c->fl_A_cold = 3.61e-13, c->fl_A_warm = 1.73e3, c->fl_Q_cold = 6.0e4, c->fl_Q_warm = 13.9e4, c->fl_T_crit = 263.15;
  c->fl_R = 8.31441, c->fl_rho = 910.0, c->fl_g = 9.81, c->fl_beta = 7.9e-8, c->fl_T_melting = 273.15;
  c->gpbld_T_0 = 273.15, c->gpbld_water_frac_coeff = 181.25, c->gpbld_water_frac_limit = 0.01;
  c->iso_softness_A = 3.1689e-24;
  c->hooke_Q = 7.88e4, c->hooke_A = 4.42165e-9, c->hooke_C = 0.16612, c->hooke_K = 1.17, c->hooke_Tr = 273.39;
  c->grain_size = 1.0e-3;
  c->gradient_method = SIAFD_B200_GRAD_HASELOFF;
  c->limit_diffusivity = 0, c->grain_size_age_coupling = 0, c->e_age_coupling = 0;
  c->D_limit = 100.0;
  const double secpera = 365.242198781 * 86400.0; // UDUNITS-2 year
  c->eemian_start = -132000.0 * secpera, c->eemian_end = -114500.0 * secpera, c->holocene_start = -11000.0 * secpera;
  c->years_per_second = 1.0 / secpera;
  c->smoother_range = 5.0e3, c->theta_min = 0.0;
  c->sea_water_density = 1028.0, c->ice_free_thickness = 0.01, c->dry_simulation = 0;
}

const char *siafd_b200_status_string(int s) {
  switch (s) {
  case SIAFD_B200_OK:
    return "ok";
  case SIAFD_B200_ERR_NEGATIVE_THICKNESS:
    return "BedSmoother detects negative original thickness";
  case SIAFD_B200_ERR_OMEGA_NEGATIVE:
    return "omega is negative in BedSmoother.theta()";
  case SIAFD_B200_ERR_HEIGHT_BELOW_BASE:
    return "height is below base of ice (height must be non-negative)";
  case SIAFD_B200_ERR_HEIGHT_ABOVE_TOP:
    return "height is above top of computational grid Lz";
  case SIAFD_B200_ERR_DIFFUSIVITY:
    return "Maximum diffusivity of SIA flow is too high";
  case SIAFD_B200_ERR_BAD_CONFIG:
    return "invalid configuration";
  case SIAFD_B200_ERR_CUDA:
    return "CUDA failure";
  case SIAFD_B200_ERR_BAD_ARGUMENT:
    return "bad argument";
  case SIAFD_B200_ERR_COMM:
    return "a rank of the communicator did not arrive at a ghost update (timed out)";
  default:
    return "unknown status";
  }
}

int siafd_b200_create(const siafd_b200_config *cfg, int device, siafd_b200_handle **out) {
  if (!cfg || !out) {
    return fail(nullptr, SIAFD_B200_ERR_BAD_ARGUMENT, "null argument");
  }
  *out = nullptr;
  std::string why;
  int st = check_config(*cfg, why);
  if (st != SIAFD_B200_OK) {
    return fail(nullptr, st, "%s", why.c_str());
  }
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    return fail(nullptr, SIAFD_B200_ERR_CUDA, "no CUDA device: %s (this library has no CPU path)",
                e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
  }
  if (device < 0) {
    if (cudaGetDevice(&device) != cudaSuccess) {
      return fail(nullptr, SIAFD_B200_ERR_CUDA, "cudaGetDevice failed");
    }
  }
  if (device >= ndev || cudaSetDevice(device) != cudaSuccess) {
    return fail(nullptr, SIAFD_B200_ERR_CUDA, "cannot select CUDA device %d", device);
  }
  siafd_b200_handle *h = new siafd_b200_handle();
  h->cfg = *cfg;
  h->z.assign(cfg->z, cfg->z + cfg->Mz);
  h->cfg.z = h->z.data();
  h->device = device;
  for (int f = 0; f < SIAFD_B200_F_COUNT; ++f) {
    h->buf[f] = nullptr;
    h->owned[f] = false;
  }
  fill_dp(h);
  // rows one CTA of the fused kernel marches over: shorter segments on small patches, so that the grid still has
  // several waves of CTAs per SM (8 ranks at 4096^2: 2.07 -> 1.96 ms per step)
  h->tuning.rows_per_cta = ((long)cfg->xm * cfg->ym <= 2048L * 1024L) ? 32 : 64;
  h->tuning.use_bulk_copy = 1;
  h->tuning.skip_ice_free = 1;
  h->tuning.wz = 4;
  h->tuning.pipeline_host = 1;
  h->tuning.pipeline_band = 1; // 4096^2, segments per band 1 / 2 / 4 / 8: 368 / 375 / 386 / 411 ms (tools/e2e_sweep.py)
  h->tuning.sparse_host = 1;
  // 4096^2 dome, one rank, 16 host cores (tools/e2e_sweep.py, profiles/e2e_sweep_r02.json): no cut 363 ms; cut with 4 / 8 /
  // 11 replicating threads 357 / 357 / 359 ms, with 8 threads and chunks of 256 columns 350 ms (wider chunks and
  // row-by-row 2D copies are slower: copy_piece)
  h->tuning.level_cut = 1;
  h->tuning.cut_cols = 256;
  h->tuning.cut_rows = 0;
  h->tuning.graph_step = 1;
  h->tuning.order_segments = 1;
  if (const char *e = getenv("SIAFD_B200_ORDER")) h->tuning.order_segments = atoi(e);
  if (const char *e = getenv("SIAFD_B200_GRAPH")) h->tuning.graph_step = atoi(e);
  if (const char *e = getenv("SIAFD_B200_SPARSE")) h->tuning.sparse_host = atoi(e);
  if (const char *e = getenv("SIAFD_B200_FILL_THREADS")) h->fill_threads = std::max(1, atoi(e));
  if (const char *e = getenv("SIAFD_B200_LEVEL_CUT")) h->tuning.level_cut = atoi(e);
  if (const char *e = getenv("SIAFD_B200_CUT_COLS")) h->tuning.cut_cols = std::max(8, atoi(e));
  if (const char *e = getenv("SIAFD_B200_CUT_ROWS")) h->tuning.cut_rows = std::max(0, atoi(e));
  if (const char *e = getenv("SIAFD_B200_ZERO_COPY")) h->zero_copy = atoi(e);
  if (const char *e = getenv("SIAFD_B200_REPL_THREADS")) h->repl_threads = std::max(1, atoi(e));
  if (const char *e = getenv("SIAFD_B200_PIPELINE")) h->tuning.pipeline_host = atoi(e);
  if (const char *e = getenv("SIAFD_B200_BAND")) h->tuning.pipeline_band = atoi(e);
  if (const char *e = getenv("SIAFD_B200_WZ")) h->tuning.wz = atoi(e);
  if (const char *e = getenv("SIAFD_B200_VVEL_KIND")) h->vvel_kind = atoi(e);
  if (const char *e = getenv("SIAFD_B200_VVEL_WZ")) h->vvel_wz = atoi(e);
  if (const char *e = getenv("SIAFD_B200_VVEL_ROWS")) h->vvel_rows = atoi(e) > 0 ? atoi(e) : 64;
  if (const char *e = getenv("SIAFD_B200_ROWS")) h->tuning.rows_per_cta = atoi(e) > 0 ? atoi(e) : 64;
  {
    const double Lz = cfg->z[cfg->Mz - 1] - cfg->z[0], dz = Lz / (cfg->Mz - 1);
    bool uniform = cfg->z[0] == 0.0;
    for (int k = 0; k < cfg->Mz && uniform; ++k) {
      uniform = std::fabs(cfg->z[k] - k * dz) <= 1e-9 * Lz;
    }
    h->inv_dz = uniform ? 1.0 / dz : 0.0;
  }
  if (slab_smem_need(h->P, true, (cfg->Mz & 1) != 0) > (size_t)227 * 1024) {
    delete h;
    return fail(nullptr, SIAFD_B200_ERR_BAD_CONFIG, "Mz = %d is too large for the shared-memory column pipeline",
                cfg->Mz);
  }
#define CUC(call)                                                                                                      \
  do {                                                                                                                 \
    cudaError_t e_ = (call);                                                                                           \
    if (e_ != cudaSuccess) {                                                                                           \
      int code_ = fail(nullptr, SIAFD_B200_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_));                  \
      siafd_b200_destroy(h);                                                                                           \
      return code_;                                                                                                    \
    }                                                                                                                  \
  } while (0)
  CUC(cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking));
  h->stream = h->own_stream;
  CUC(cudaMalloc(&h->d_z, sizeof(double) * cfg->Mz));
  CUC(cudaMemcpy(h->d_z, h->z.data(), sizeof(double) * cfg->Mz, cudaMemcpyHostToDevice));
  CUC(cudaMalloc(&h->d_err, sizeof(unsigned)));
  CUC(cudaMalloc(&h->d_dmax, sizeof(unsigned long long)));
  CUC(cudaMalloc(&h->d_hdc, sizeof(int)));
  CUC(cudaMalloc(&h->d_segw, 256 * sizeof(int)));
  CUC(cudaMalloc(&h->d_segdone, sizeof(unsigned)));
  CUC(cudaMemset(h->d_segw, 0, 256 * sizeof(int)));
  CUC(cudaMemset(h->d_segdone, 0, sizeof(unsigned)));
  CUC(cudaMemset(h->d_err, 0, sizeof(unsigned)));
  CUC(cudaMemset(h->d_dmax, 0, sizeof(unsigned long long)));
  CUC(cudaMemset(h->d_hdc, 0, sizeof(int)));
  CUC(cudaMallocHost(&h->h_res, sizeof(*h->h_res)));
  std::memset(h->h_res, 0, sizeof(*h->h_res));
#undef CUC
  *out = h;
  return SIAFD_B200_OK;
}

void siafd_b200_destroy(siafd_b200_handle *h) {
  if (!h) {
    return;
  }
  cudaSetDevice(h->device);
  if (h->own_stream) {
    cudaStreamSynchronize(h->own_stream);
  }
  for (int f = 0; f < SIAFD_B200_F_COUNT; ++f) {
    if (h->buf[f] && h->owned[f]) {
      cudaFree(h->buf[f]);
    }
  }
  cudaFree(h->d_z);
  cudaFree(h->d_err);
  cudaFree(h->d_dmax);
  cudaFree(h->d_hdc);
  cudaFree(h->d_segw);
  cudaFree(h->d_pieces);
  cudaFree(h->d_segdone);
  cudaFree(h->d_cfl);
  if (h->h_cfl) cudaFreeHost(h->h_cfl);
  cudaFree(h->d_global_bed);
  comm_release(h);
  for (void *p : h->ipc_mapped) cudaIpcCloseMemHandle(p);
  cudaFree(h->d_pad);
  if (h->h_res) {
    cudaFreeHost(h->h_res);
  }
  for (cudaEvent_t e : h->ev_pipe) cudaEventDestroy(e);
  if (h->s_up) cudaStreamDestroy(h->s_up);
  if (h->s_dn) cudaStreamDestroy(h->s_dn);
  for (auto &e : h->ev_sec) cudaEventDestroy(e);
  for (size_t q = 0; q < h->ev_start.size(); ++q) {
    cudaEventDestroy(h->ev_start[q]);
    cudaEventDestroy(h->ev_stop[q]);
  }
  if (h->own_stream) {
    cudaStreamDestroy(h->own_stream);
  }
  delete h;
}

const char *siafd_b200_last_error(const siafd_b200_handle *h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int64_t siafd_b200_field_size(const siafd_b200_handle *h, int f) {
  if (!h) return -1;
  const FieldMeta m = meta(h->cfg, f);
  return m.width < 0 ? -1 : field_cells(h->cfg, m.width) * m.dof;
}
int siafd_b200_field_width(const siafd_b200_handle *h, int f) {
  if (!h) return -1; return meta(h->cfg, f).width; }
int siafd_b200_field_dof(const siafd_b200_handle *h, int f) {
  if (!h) return -1; return meta(h->cfg, f).dof; }

int siafd_b200_bind(siafd_b200_handle *h, int f, void *device_ptr) {
  if (!h) return null_handle();
  h->cfl3_fresh = false; // the fields the fused CFL maxima were taken on are about to change
  if (f < 0 || f >= SIAFD_B200_F_COUNT) {
    return fail(h, SIAFD_B200_ERR_BAD_ARGUMENT, "bad field id %d", f);
  }
  if (device_ptr && ((uintptr_t)device_ptr & 15u)) {
    return fail(h, SIAFD_B200_ERR_BAD_ARGUMENT, "device pointer of field %d must be 16-byte aligned", f);
  }
  if (h->buf[f] == device_ptr) {
    return SIAFD_B200_OK;
  }
  if (h->buf[f] && h->owned[f]) {
    CU(h, cudaSetDevice(h->device));
    CU(h, cudaStreamSynchronize(h->stream));
    CU(h, cudaFree(h->buf[f]));
  }
  h->buf[f] = device_ptr;
  h->owned[f] = false;
  invalidate_graphs(h);
  return SIAFD_B200_OK;
}

void *siafd_b200_device_ptr(siafd_b200_handle *h, int f) {
  if (!h) return nullptr;
  if (cudaSetDevice(h->device) != cudaSuccess || ensure(h, f) != SIAFD_B200_OK) {
    return nullptr;
  }
  return h->buf[f];
}

int siafd_b200_set_stream(siafd_b200_handle *h, void *cuda_stream) {
  if (!h) return null_handle();
  h->stream = cuda_stream ? (cudaStream_t)cuda_stream : h->own_stream;
  return SIAFD_B200_OK;
}

int siafd_b200_upload(siafd_b200_handle *h, int f, const double *host) {
  if (!h) return null_handle();
  h->cfl3_fresh = false; // the fields the fused CFL maxima were taken on are about to change
  CU(h, cudaSetDevice(h->device));
  int st = ensure(h, f);
  if (st) return st;
  const size_t bytes = (size_t)siafd_b200_field_size(h, f) * sizeof(double);
  CU(h, cudaMemcpyAsync(h->buf[f], host, bytes, cudaMemcpyHostToDevice, h->stream));
  h->bytes_h2d += (int64_t)bytes;
  return SIAFD_B200_OK;
}

int siafd_b200_download(siafd_b200_handle *h, int f, double *host) {
  if (!h) return null_handle();
  CU(h, cudaSetDevice(h->device));
  int st = ensure(h, f);
  if (st) return st;
  const size_t bytes = (size_t)siafd_b200_field_size(h, f) * sizeof(double);
  const void *src = h->buf[f];
  if (f == SIAFD_B200_F_TOPGSMOOTH && !(h->cfg.smoother_range > 0.0) && !h->smoother_set && h->buf[SIAFD_B200_F_BED]) {
    src = h->buf[SIAFD_B200_F_BED]; // smoother off: topgsmooth IS the bed (sia/BedSmoother.cc:101-109)
  }
  CU(h, cudaMemcpyAsync(host, src, bytes, cudaMemcpyDeviceToHost, h->stream));
  CU(h, cudaStreamSynchronize(h->stream));
  h->bytes_d2h += (int64_t)bytes;
  return SIAFD_B200_OK;
}

int siafd_b200_wrap_ghosts(siafd_b200_handle *h, int f) {
  if (!h) return null_handle();
  CU(h, cudaSetDevice(h->device));
  int st = ensure(h, f);
  if (st) return st;
  st = wrap_dir(h, f, 0);
  if (st) return st;
  return wrap_dir(h, f, 1);
}

int siafd_b200_wrap_ghosts_dir(siafd_b200_handle *h, int f, int dir) {
  if (!h) return null_handle();
  CU(h, cudaSetDevice(h->device));
  int st = ensure(h, f);
  if (st) return st;
  if (dir != 0 && dir != 1) {
    return fail(h, SIAFD_B200_ERR_BAD_ARGUMENT, "dir must be 0 (x) or 1 (y)");
  }
  return wrap_dir(h, f, dir);
}

// Strip geometry of the two-stage exchange.  Stage x (dir_x != 0): w columns, owned rows.
// Stage y (dir_y != 0): w rows, all columns of the local array (x ghosts included).
static bool halo_rect(const siafd_b200_handle *h, int f, int dir_x, int dir_y, int width, bool ghost_side, int *i0,
                      int *j0, int *wc, int *hc) {
  const FieldMeta m = meta(h->cfg, f);
  if (m.width < 0 || width < 1 || width > m.width || ((dir_x != 0) == (dir_y != 0))) {
    return false;
  }
  const int W = m.width, xm = h->cfg.xm, ym = h->cfg.ym;
  if (dir_x != 0) {
    *wc = width, *hc = ym, *j0 = W;
    if (ghost_side) {
      *i0 = dir_x < 0 ? W - width : W + xm;
    } else {
      *i0 = dir_x < 0 ? W : W + xm - width;
    }
  } else {
    *wc = xm + 2 * W, *hc = width, *i0 = 0;
    if (ghost_side) {
      *j0 = dir_y < 0 ? W - width : W + ym;
    } else {
      *j0 = dir_y < 0 ? W : W + ym - width;
    }
  }
  return true;
}

int64_t siafd_b200_halo_count(const siafd_b200_handle *h, int f, int dir_x, int dir_y, int width) {
  if (!h) return -1;
  int i0, j0, wc, hc;
  if (!halo_rect(h, f, dir_x, dir_y, width, false, &i0, &j0, &wc, &hc)) {
    return -1;
  }
  return (int64_t)wc * hc * meta(h->cfg, f).dof;
}

int siafd_b200_halo_pack(siafd_b200_handle *h, int f, int dir_x, int dir_y, int width, double *device_buf) {
  if (!h) return null_handle();
  CU(h, cudaSetDevice(h->device));
  int st = ensure(h, f);
  if (st) return st;
  int i0, j0, wc, hc;
  if (!halo_rect(h, f, dir_x, dir_y, width, false, &i0, &j0, &wc, &hc)) {
    return fail(h, SIAFD_B200_ERR_BAD_ARGUMENT, "bad halo description");
  }
  const FieldMeta m = meta(h->cfg, f);
  h->launches += launch_copy_region(device_buf, wc, 0, 0, (const double *)h->buf[f], h->cfg.xm + 2 * m.width, i0, j0, wc,
                                    hc, m.dof, h->stream);
  CU(h, cudaGetLastError());
  return SIAFD_B200_OK;
}

int siafd_b200_halo_unpack(siafd_b200_handle *h, int f, int dir_x, int dir_y, int width, const double *device_buf) {
  if (!h) return null_handle();
  CU(h, cudaSetDevice(h->device));
  int st = ensure(h, f);
  if (st) return st;
  int i0, j0, wc, hc;
  if (!halo_rect(h, f, dir_x, dir_y, width, true, &i0, &j0, &wc, &hc)) {
    return fail(h, SIAFD_B200_ERR_BAD_ARGUMENT, "bad halo description");
  }
  const FieldMeta m = meta(h->cfg, f);
  h->launches += launch_copy_region((double *)h->buf[f], h->cfg.xm + 2 * m.width, i0, j0, device_buf, wc, 0, 0, wc, hc,
                                    m.dof, h->stream);
  CU(h, cudaGetLastError());
  return SIAFD_B200_OK;
}

// ---- peer halo exchange -------------------------------------------------------------------------------
static const int HALO_DX[8] = {-1, 0, 1, -1, 1, -1, 0, 1}, HALO_DY[8] = {-1, -1, -1, 0, 0, 1, 1, 1};

static int ensure_pad(siafd_b200_handle *h) {
  if (!h->d_pad) {
    CU(h, cudaMalloc(&h->d_pad, 32 * sizeof(unsigned long long)));
    CU(h, cudaMemset(h->d_pad, 0, 32 * sizeof(unsigned long long)));
  }
  return SIAFD_B200_OK;
}

int siafd_b200_ipc_export(siafd_b200_handle *h, int field, void *handle64) {
  if (!h) return null_handle();
  CU(h, cudaSetDevice(h->device));
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
  void *p = nullptr;
  if (field < 0) {
    int st = ensure_pad(h);
    if (st) return st;
    p = h->d_pad;
  } else {
    int st = ensure(h, field);
    if (st) return st;
    if (!h->owned[field]) {
      return fail(h, SIAFD_B200_ERR_BAD_ARGUMENT, "field %d is bound to caller memory: only handle-owned storage can be exported",
                  field);
    }
    p = h->buf[field];
  }
  CU(h, cudaStreamSynchronize(h->stream)); // the zero-fill of a fresh buffer
  CU(h, cudaIpcGetMemHandle(reinterpret_cast<cudaIpcMemHandle_t *>(handle64), p));
  return SIAFD_B200_OK;
}

int siafd_b200_ipc_open(siafd_b200_handle *h, const void *handle64, void **peer_ptr) {
  if (!h) return null_handle();
  CU(h, cudaSetDevice(h->device));
  cudaIpcMemHandle_t mh;
  std::memcpy(&mh, handle64, sizeof(mh));
  CU(h, cudaIpcOpenMemHandle(peer_ptr, mh, cudaIpcMemLazyEnablePeerAccess));
  h->ipc_mapped.push_back(*peer_ptr);
  return SIAFD_B200_OK;
}

int siafd_b200_halo_attach(siafd_b200_handle *h, int field, int dir, void *peer_base, int peer_xm, int peer_ym) {
  if (!h) return null_handle();
  if (dir < 0 || dir > 7 || field >= SIAFD_B200_F_COUNT) {
    return fail(h, SIAFD_B200_ERR_BAD_ARGUMENT, "bad halo attachment (field %d, dir %d)", field, dir);
  }
  if (field < 0) {
    int st = ensure_pad(h);
    if (st) return st;
    h->peer_pad[dir] = peer_base ? (unsigned long long *)peer_base : h->d_pad;
    h->pad_attached[dir] = true;
    return SIAFD_B200_OK;
  }
  h->peers[field][dir].base = (double *)peer_base;
  h->peers[field][dir].xm = peer_xm, h->peers[field][dir].ym = peer_ym;
  h->peers[field][dir].attached = true;
  return SIAFD_B200_OK;
}

// descriptors of the eight strips (4 edges, 4 corners) of field f going to the neighbours (self if !to_peers)
static int halo_descriptors(siafd_b200_handle *h, int f, int w, bool to_peers, HaloBatch &B) {
  const siafd_b200_config &c = h->cfg;
  const FieldMeta m = meta(c, f);
  if (w < 1 || w > m.width) return fail(h, SIAFD_B200_ERR_BAD_ARGUMENT, "ghost update: bad width %d for field %d", w, f);
  const int W = m.width;
  for (int d = 0; d < 8; ++d) {
    const siafd_b200_handle::Peer &P = h->peers[f][d];
    if (to_peers && (!P.attached || !h->pad_attached[d])) {
      return fail(h, SIAFD_B200_ERR_BAD_ARGUMENT, "halo_push: field %d / direction %d not attached", f, d);
    }
    const bool remote = to_peers && P.base != nullptr;
    const int dx = HALO_DX[d], dy = HALO_DY[d];
    HaloDesc &D = B.d[B.n++];
    D.src = (const double *)h->buf[f];
    D.dst = remote ? P.base : (double *)h->buf[f];
    const int pxm = remote ? P.xm : c.xm, pym = remote ? P.ym : c.ym;
    D.src_row_cells = c.xm + 2 * W, D.dst_row_cells = pxm + 2 * W;
    D.dof = m.dof, D.pad = 0;
    // my owned strip facing the neighbour -> the neighbour's ghost cells facing me (local array indices)
    D.wc = dx == 0 ? c.xm : w, D.hc = dy == 0 ? c.ym : w;
    D.src_i0 = W + (dx > 0 ? c.xm - w : 0), D.src_j0 = W + (dy > 0 ? c.ym - w : 0);
    D.dst_i0 = dx > 0 ? W - w : (dx < 0 ? W + pxm : W);
    D.dst_j0 = dy > 0 ? W - w : (dy < 0 ? W + pym : W);
  }
  return SIAFD_B200_OK;
}

int siafd_b200_wrap_ghosts_many(siafd_b200_handle *h, int n, const int *fields) {
  if (!h) return null_handle();
  CU(h, cudaSetDevice(h->device));
  const siafd_b200_config &c = h->cfg;
  if (c.xm != c.Mx || c.ym != c.My) {
    return fail(h, SIAFD_B200_ERR_BAD_ARGUMENT, "periodic self-wrap needs a patch spanning the whole domain");
  }
  if (n < 1 || n > 6) return fail(h, SIAFD_B200_ERR_BAD_ARGUMENT, "wrap_ghosts_many: 1..6 fields");
  HaloBatch B;
  B.n = 0;
  for (int q = 0; q < n; ++q) {
    int st = ensure(h, fields[q]);
    if (st) return st;
    if ((st = halo_descriptors(h, fields[q], meta(c, fields[q]).width, false, B))) return st;
  }
  h->launches += launch_halo_push(B, h->stream);
  CU(h, cudaGetLastError());
  return SIAFD_B200_OK;
}

int siafd_b200_halo_push(siafd_b200_handle *h, int n, const int *fields, const int *widths, int phase) {
  if (!h) return null_handle();
  CU(h, cudaSetDevice(h->device));
  if (phase < 0 || phase > 3 || n < 1 || n > 6) {
    return fail(h, SIAFD_B200_ERR_BAD_ARGUMENT, "halo_push: phase in 0..3 and 1..6 fields");
  }
  HaloBatch B;
  B.n = 0;
  for (int q = 0; q < n; ++q) {
    int st = ensure(h, fields[q]);
    if (st) return st;
    if ((st = halo_descriptors(h, fields[q], widths[q], true, B))) return st;
  }
  h->launches += launch_halo_push(B, h->stream);
  HaloSignal S;
  h->halo_step[phase] += 1;
  S.value = h->halo_step[phase];
  for (int d = 0; d < 8; ++d) S.slot[d] = h->peer_pad[d] + phase * 8 + (7 - d); // the neighbour sees me in direction 7 - d
  h->launches += launch_halo_signal(S, h->stream);
  CU(h, cudaGetLastError());
  return SIAFD_B200_OK;
}

int siafd_b200_halo_wait(siafd_b200_handle *h, int phase) {
  if (!h) return null_handle();
  CU(h, cudaSetDevice(h->device));
  if (phase < 0 || phase > 3 || !h->d_pad) return fail(h, SIAFD_B200_ERR_BAD_ARGUMENT, "halo_wait: nothing was pushed");
  h->launches += launch_halo_wait(h->d_pad + phase * 8, h->halo_step[phase], h->stream);
  CU(h, cudaGetLastError());
  return SIAFD_B200_OK;
}

int siafd_b200_preprocess_bed(siafd_b200_handle *h, const double *global_bed_host) {
  if (!h) return null_handle();
  CU(h, cudaSetDevice(h->device));
  const siafd_b200_config &c = h->cfg;
  const int five[5] = {SIAFD_B200_F_TOPGSMOOTH, SIAFD_B200_F_MAXTL, SIAFD_B200_F_C2, SIAFD_B200_F_C3, SIAFD_B200_F_C4};
  for (int q = 0; q < 5; ++q) {
    int st = ensure(h, five[q]);
    if (st) return st;
  }
  if (c.smoother_range <= 0.0) {
    // sia/BedSmoother.cc:101-109: topgsmooth is a ghosted copy of topg (done per update from the
    // uploaded bed), theta() returns 1
    h->bedNx = h->bedNy = -1;
    h->P.smoother_active = 0;
    h->smoother_set = false;
    invalidate_graphs(h);
    return SIAFD_B200_OK;
  }
  // sia/BedSmoother.cc:111-137
  int Nx = (int)ceil(c.smoother_range / c.dx), Ny = (int)ceil(c.smoother_range / c.dy);
  if (Nx < 1) Nx = 1;
  if (Ny < 1) Ny = 1;
  if (Nx >= c.Mx || Ny >= c.My) {
    return fail(h, SIAFD_B200_ERR_BAD_CONFIG,
                "input Nx, Ny in bed smoother is too large because domain of smoothing exceeds IceGrid domain");
  }
  const size_t bytes = (size_t)c.Mx * c.My * sizeof(double);
  if (!h->d_global_bed) {
    CU(h, cudaMalloc(&h->d_global_bed, bytes));
  }
  CU(h, cudaMemcpyAsync(h->d_global_bed, global_bed_host, bytes, cudaMemcpyHostToDevice, h->stream));
  h->launches += launch_preprocess_bed(h->P, h->d_global_bed, Nx, Ny, (double *)h->buf[SIAFD_B200_F_TOPGSMOOTH],
                                       (double *)h->buf[SIAFD_B200_F_MAXTL], (double *)h->buf[SIAFD_B200_F_C2],
                                       (double *)h->buf[SIAFD_B200_F_C3], (double *)h->buf[SIAFD_B200_F_C4], h->stream);
  CU(h, cudaGetLastError());
  CU(h, cudaStreamSynchronize(h->stream));
  h->bedNx = Nx, h->bedNy = Ny;
  h->P.smoother_active = 1;
  h->smoother_set = true;
  invalidate_graphs(h);
  return SIAFD_B200_OK;
}

int siafd_b200_set_smoothed_bed(siafd_b200_handle *h, const double *topgsmooth, const double *maxtl, const double *C2,
                                const double *C3, const double *C4, int smoother_active) {
  if (!h) return null_handle();
  int st;
  if ((st = siafd_b200_upload(h, SIAFD_B200_F_TOPGSMOOTH, topgsmooth))) return st;
  if ((st = siafd_b200_upload(h, SIAFD_B200_F_MAXTL, maxtl))) return st;
  if ((st = siafd_b200_upload(h, SIAFD_B200_F_C2, C2))) return st;
  if ((st = siafd_b200_upload(h, SIAFD_B200_F_C3, C3))) return st;
  if ((st = siafd_b200_upload(h, SIAFD_B200_F_C4, C4))) return st;
  CU(h, cudaStreamSynchronize(h->stream));
  h->P.smoother_active = smoother_active ? 1 : 0;
  h->smoother_set = true;
  invalidate_graphs(h);
  return SIAFD_B200_OK;
}

int siafd_b200_compute_gradient(siafd_b200_handle *h) {
  if (!h) return null_handle();
  CU(h, cudaSetDevice(h->device));
  const int need[] = {SIAFD_B200_F_SURFACE, SIAFD_B200_F_THICKNESS, SIAFD_B200_F_MASK, SIAFD_B200_F_BED,
                      SIAFD_B200_F_H_X,     SIAFD_B200_F_H_Y,       SIAFD_B200_F_W_I,  SIAFD_B200_F_W_J};
  for (int f : need) {
    int st = ensure(h, f);
    if (st) return st;
  }
  const Fields F = fields_of(h);
  h->launches += launch_gradient(h->P, F, h->stream);
  CU(h, cudaGetLastError());
  return SIAFD_B200_OK;
}

// checks, scratch fields and the 2D preparation of SIAFD::compute_diffusivity (SIAFD.cc:555-582)
extern "C++" {
int siafd_host::flux_velocity_prepare(siafd_b200_handle *h, int full_update, double current_time, bool prep2d_done) {
  CU(h, cudaSetDevice(h->device));
  const siafd_b200_config &c = h->cfg;
  const int need[] = {SIAFD_B200_F_SURFACE,    SIAFD_B200_F_THICKNESS, SIAFD_B200_F_MASK,  SIAFD_B200_F_BED,
                      SIAFD_B200_F_ENTHALPY,   SIAFD_B200_F_TOPGSMOOTH, SIAFD_B200_F_MAXTL, SIAFD_B200_F_C2,
                      SIAFD_B200_F_C3,         SIAFD_B200_F_C4,        SIAFD_B200_F_H_X,   SIAFD_B200_F_H_Y,
                      SIAFD_B200_F_THK_SMOOTH, SIAFD_B200_F_THETA,     SIAFD_B200_F_D,     SIAFD_B200_F_FLUX};
  for (int f : need) {
    int st = ensure(h, f);
    if (st) return st;
  }
  if (full_update) {
    int st;
    if ((st = ensure(h, SIAFD_B200_F_U))) return st;
    if ((st = ensure(h, SIAFD_B200_F_V))) return st;
    if ((st = ensure(h, SIAFD_B200_F_SLIDING))) return st; // zero-filled = ZeroSliding
  }
  if (h->P.use_age && !h->buf[SIAFD_B200_F_AGE]) {
    // sia/SIAFD.cc:78-86
    return fail(h, SIAFD_B200_ERR_BAD_CONFIG, "SIAFD: age is needed (age coupling is on) but no age field was provided");
  }
  if (c.smoother_range > 0.0 && !h->smoother_set) {
    return fail(h, SIAFD_B200_ERR_BAD_CONFIG,
                "bed smoother is on (range %.1f m): call siafd_b200_preprocess_bed or _set_smoothed_bed first",
                c.smoother_range);
  }
  // (smoother off: sia/BedSmoother.cc:101-109, topgsmooth = ghosted copy of the bed: fields_of() hands the bed to the
  // kernels in its place; maxtl, C2..C4 stay zero)
  h->P.current_time = current_time;
  const Fields F = fields_of(h);
  CU(h, cudaMemsetAsync(h->d_dmax, 0, sizeof(unsigned long long), h->stream));
  CU(h, cudaMemsetAsync(h->d_hdc, 0, sizeof(int), h->stream));
  if (!prep2d_done) h->launches += launch_prep2d(h->P, F, h->stream); // sia/SIAFD.cc:580-582
  CU(h, cudaGetLastError());
  return SIAFD_B200_OK;
}
} // extern "C++"

// the fused kernel on the row segments [seg0, seg0 + nseg) (nseg < 0: all)
extern "C++" {
int siafd_host::flux_velocity_launch(siafd_b200_handle *h, int full_update, int seg0, int nseg, const PeerPush *push) {
  const Fields F = fields_of(h);
  const Tuning T = h->tuning;
  const bool timed = h->timing && h->ev_count < (int)h->ev_start.size();
  if (timed) {
    CU(h, cudaEventRecord(h->ev_start[h->ev_count], h->stream));
  }
  const int n = launch_slab(h->P, F, full_update != 0, T, (long)siafd_b200_field_size(h, SIAFD_B200_F_ENTHALPY),
                            (long)siafd_b200_field_size(h, SIAFD_B200_F_THK_SMOOTH), h->inv_dz, seg0, nseg, h->stream, push);
  if (timed) {
    CU(h, cudaEventRecord(h->ev_stop[h->ev_count], h->stream));
    h->ev_count += 1;
  }
  if (n < 0) {
    return fail(h, SIAFD_B200_ERR_CUDA, "could not configure the fused kernel (shared memory %zu bytes)",
                slab_smem_need(h->P, full_update != 0, (h->P.Mz & 1) != 0));
  }
  h->launches += n;
  CU(h, cudaGetLastError());
  h->result_pending = true;
  return SIAFD_B200_OK;
}
} // extern "C++"

int siafd_b200_compute_flux_velocity(siafd_b200_handle *h, int full_update, double current_time) {
  if (!h) return null_handle();
  h->cfl3_fresh = false; // the fields the fused CFL maxima were taken on are about to change
  int st = flux_velocity_prepare(h, full_update, current_time);
  if (st) return st;
  return flux_velocity_launch(h, full_update, 0, -1);
}

static int ensure_cfl(siafd_b200_handle *h) {
  if (!h->d_cfl) {
    CU(h, cudaMalloc(&h->d_cfl, 8 * sizeof(unsigned long long)));
    CU(h, cudaMallocHost(&h->h_cfl, 8 * sizeof(unsigned long long)));
    CU(h, cudaMemsetAsync(h->d_cfl, 0, 8 * sizeof(unsigned long long), h->stream));
  }
  return SIAFD_B200_OK;
}

int siafd_b200_compute_vertical_velocity(siafd_b200_handle *h, int upstream, int use_basal_melt) {
  if (!h) return null_handle();
  CU(h, cudaSetDevice(h->device));
  const int need[] = {SIAFD_B200_F_MASK, SIAFD_B200_F_THICKNESS, SIAFD_B200_F_U, SIAFD_B200_F_V, SIAFD_B200_F_W};
  for (int f : need) {
    int st = ensure(h, f);
    if (st) return st;
  }
  if (use_basal_melt) {
    int st = ensure(h, SIAFD_B200_F_BASAL_MELT);
    if (st) return st;
  }
  int st = ensure_cfl(h);
  if (st) return st;
  auto D = [&](int f) { return (double *)h->buf[f]; };
  const double *bmr = use_basal_melt ? D(SIAFD_B200_F_BASAL_MELT) : nullptr;
  // the marching kernel also takes the 3D CFL maxima (StressBalance.cc:186-200 evaluates them right after w)
  CU(h, cudaMemsetAsync(h->d_cfl, 0, 4 * sizeof(unsigned long long), h->stream));
  int n = 0;
  if (h->vvel_kind == 0) { // shared-memory kernel (z sweep in registers); 0 launches = does not apply
    n = launch_vvel_slab(h->P, D(SIAFD_B200_F_MASK), D(SIAFD_B200_F_THICKNESS), D(SIAFD_B200_F_U), D(SIAFD_B200_F_V), bmr,
                         upstream, h->d_z, D(SIAFD_B200_F_W), h->d_cfl, h->d_err, h->vvel_rows, h->vvel_wz,
                         (long)siafd_b200_field_size(h, SIAFD_B200_F_U), h->inv_dz, h->stream);
  }
  if (n == 0) n = launch_vvel_march(h->P, D(SIAFD_B200_F_MASK), D(SIAFD_B200_F_THICKNESS), D(SIAFD_B200_F_U), D(SIAFD_B200_F_V),
                            bmr, upstream, h->d_z, D(SIAFD_B200_F_W), h->d_cfl, h->d_err, h->vvel_rows, h->stream);
  h->cfl3_fresh = n > 0;
  if (n == 0) { // more than 256 levels: generic warp-per-column kernel
    n = launch_vertical_velocity(h->P, D(SIAFD_B200_F_MASK), D(SIAFD_B200_F_U), D(SIAFD_B200_F_V), bmr, upstream, h->d_z,
                                 D(SIAFD_B200_F_W), h->stream);
  }
  h->launches += n;
  h->result_pending = true;
  CU(h, cudaGetLastError());
  return SIAFD_B200_OK;
}

// ---- SIAFD_Regional::compute_surface_gradient -------------------------------------------------
int siafd_b200_compute_gradient_no_model(siafd_b200_handle *h) {
  if (!h) return null_handle();
  CU(h, cudaSetDevice(h->device));
  const int need[] = {SIAFD_B200_F_NO_MODEL_SURFACE, SIAFD_B200_F_MASK,  SIAFD_B200_F_H_X_NO_MODEL,
                      SIAFD_B200_F_H_Y_NO_MODEL,     SIAFD_B200_F_W_I,   SIAFD_B200_F_W_J,
                      SIAFD_B200_F_THICKNESS,        SIAFD_B200_F_BED};
  for (int f : need) {
    int st = ensure(h, f);
    if (st) return st;
  }
  Fields F = fields_of(h);
  F.h = (const double *)h->buf[SIAFD_B200_F_NO_MODEL_SURFACE];
  F.h_x = (double *)h->buf[SIAFD_B200_F_H_X_NO_MODEL];
  F.h_y = (double *)h->buf[SIAFD_B200_F_H_Y_NO_MODEL];
  DP P = h->P;
  P.grad = SIAFD_B200_GRAD_HASELOFF; // regional/SIAFD_Regional.cc:52-55 calls surface_gradient_haseloff directly
  h->launches += launch_gradient(P, F, h->stream);
  CU(h, cudaGetLastError());
  return SIAFD_B200_OK;
}

int siafd_b200_apply_no_model_gradient(siafd_b200_handle *h) {
  if (!h) return null_handle();
  CU(h, cudaSetDevice(h->device));
  const int need[] = {SIAFD_B200_F_NO_MODEL_MASK, SIAFD_B200_F_H_X_NO_MODEL, SIAFD_B200_F_H_Y_NO_MODEL,
                      SIAFD_B200_F_H_X, SIAFD_B200_F_H_Y};
  for (int f : need) {
    int st = ensure(h, f);
    if (st) return st;
  }
  auto D = [&](int f) { return (double *)h->buf[f]; };
  h->launches += launch_regional_override(h->P, D(SIAFD_B200_F_NO_MODEL_MASK), D(SIAFD_B200_F_H_X_NO_MODEL),
                                          D(SIAFD_B200_F_H_Y_NO_MODEL), D(SIAFD_B200_F_H_X), D(SIAFD_B200_F_H_Y),
                                          h->stream);
  CU(h, cudaGetLastError());
  return SIAFD_B200_OK;
}

// ---- IceModelVec3::getSurfaceValues / getHorSlice (util/iceModelVec3.cc:209-240) ----------------------------------
static int value_at_height(siafd_b200_handle *h, int field3d, bool at_surface, double z, double *out_dev) {
  CU(h, cudaSetDevice(h->device));
  switch (field3d) {
  case SIAFD_B200_F_ENTHALPY:
  case SIAFD_B200_F_AGE:
  case SIAFD_B200_F_U:
  case SIAFD_B200_F_V:
  case SIAFD_B200_F_W:
  case SIAFD_B200_F_STRAIN_HEATING:
    break;
  default:
    return fail(h, SIAFD_B200_ERR_BAD_ARGUMENT, "field %d is not a 3D field", field3d);
  }
  if (!out_dev) return fail(h, SIAFD_B200_ERR_BAD_ARGUMENT, "out_dev is NULL");
  if (!h->buf[field3d]) {
    return fail(h, SIAFD_B200_ERR_BAD_ARGUMENT, "field %d has not been computed, uploaded or bound yet", field3d);
  }
  if (at_surface) {
    int st = ensure(h, SIAFD_B200_F_THICKNESS);
    if (st) return st;
  }
  h->launches += launch_value_at_height(h->P, (const double *)h->buf[field3d], meta(h->cfg, field3d).width,
                                        at_surface ? (const double *)h->buf[SIAFD_B200_F_THICKNESS] : nullptr,
                                        h->cfg.w_geom, z, h->d_z, out_dev, h->stream);
  CU(h, cudaGetLastError());
  return SIAFD_B200_OK;
}

int siafd_b200_surface_values(siafd_b200_handle *h, int field3d, double *out_dev) {
  if (!h) return null_handle();
  return value_at_height(h, field3d, true, 0.0, out_dev);
}

int siafd_b200_hor_slice(siafd_b200_handle *h, int field3d, double z, double *out_dev) {
  if (!h) return null_handle();
  return value_at_height(h, field3d, false, z, out_dev);
}

// ---- volumetric strain heating -------------------------------------------------------------
int siafd_b200_compute_strain_heating(siafd_b200_handle *h, int flow_law, double glen_exponent,
                                      double enhancement_factor) {
  if (!h) return null_handle();
  CU(h, cudaSetDevice(h->device));
  if (!(glen_exponent > 0.0) || !(enhancement_factor > 0.0)) {
    return fail(h, SIAFD_B200_ERR_BAD_ARGUMENT, "Glen exponent and enhancement factor must be positive");
  }
  const int need[] = {SIAFD_B200_F_MASK, SIAFD_B200_F_THICKNESS, SIAFD_B200_F_ENTHALPY, SIAFD_B200_F_U, SIAFD_B200_F_V,
                      SIAFD_B200_F_STRAIN_HEATING};
  for (int f : need) {
    int st = ensure(h, f);
    if (st) return st;
  }
  auto D = [&](int f) { return (double *)h->buf[f]; };
  const int n = launch_strain_heating(h->P, flow_law, glen_exponent, enhancement_factor, D(SIAFD_B200_F_MASK),
                                      D(SIAFD_B200_F_THICKNESS), D(SIAFD_B200_F_ENTHALPY), D(SIAFD_B200_F_U),
                                      D(SIAFD_B200_F_V), h->d_z, D(SIAFD_B200_F_STRAIN_HEATING), h->d_err, h->stream);
  if (n < 0) {
    return fail(h, SIAFD_B200_ERR_BAD_CONFIG,
                "strain heating: flow law %d has no softness (gk: GoldsbyKohlstedt.cc:102-108) or Mz > 256", flow_law);
  }
  h->launches += n;
  h->result_pending = true;
  CU(h, cudaGetLastError());
  return SIAFD_B200_OK;
}

// ---- GeometryEvolution on device ----------------------------------------------------------
int siafd_b200_mass_flow_step(siafd_b200_handle *h, double dt) {
  if (!h) return null_handle();
  h->cfl3_fresh = false; // the fields the fused CFL maxima were taken on are about to change
  CU(h, cudaSetDevice(h->device));
  const int need[] = {SIAFD_B200_F_THICKNESS,  SIAFD_B200_F_BED,      SIAFD_B200_F_FLUX,
                      SIAFD_B200_F_THK_CHANGE, SIAFD_B200_F_FLUX_DIV, SIAFD_B200_F_CONS_ERR};
  for (int f : need) {
    int st = ensure(h, f);
    if (st) return st;
  }
  auto D = [&](int f) { return (double *)h->buf[f]; };
  // the advective velocity needs its neighbours: without ghosts only ZeroSliding (no field / all zero) is possible
  const double *vel = (h->cfg.w_sliding >= 1) ? D(SIAFD_B200_F_SLIDING) : nullptr;
  h->launches += launch_mass_flow(h->P, dt, D(SIAFD_B200_F_THICKNESS), D(SIAFD_B200_F_BED), D(SIAFD_B200_F_SEA_LEVEL),
                                  vel, D(SIAFD_B200_F_VEL_BC_MASK), D(SIAFD_B200_F_THK_BC_MASK), D(SIAFD_B200_F_FLUX),
                                  D(SIAFD_B200_F_FLUX_DIV), D(SIAFD_B200_F_THK_CHANGE), D(SIAFD_B200_F_CONS_ERR),
                                  h->stream);
  h->launches += launch_mass_apply(h->P, D(SIAFD_B200_F_THICKNESS), D(SIAFD_B200_F_THK_CHANGE), h->stream);
  CU(h, cudaGetLastError());
  return SIAFD_B200_OK;
}

int siafd_b200_mass_source_step(siafd_b200_handle *h, double dt, double ice_density, int use_basal_melt) {
  if (!h) return null_handle();
  h->cfl3_fresh = false; // the fields the fused CFL maxima were taken on are about to change
  CU(h, cudaSetDevice(h->device));
  const int need[] = {SIAFD_B200_F_THICKNESS, SIAFD_B200_F_MASK, SIAFD_B200_F_SMB, SIAFD_B200_F_EFF_SMB,
                      SIAFD_B200_F_EFF_BMB};
  for (int f : need) {
    int st = ensure(h, f);
    if (st) return st;
  }
  if (use_basal_melt) {
    int st = ensure(h, SIAFD_B200_F_BASAL_MELT);
    if (st) return st;
  }
  if (!(ice_density > 0.0)) {
    return fail(h, SIAFD_B200_ERR_BAD_ARGUMENT, "ice_density must be positive");
  }
  auto D = [&](int f) { return (double *)h->buf[f]; };
  h->launches += launch_mass_source(h->P, dt, ice_density, use_basal_melt, D(SIAFD_B200_F_THICKNESS),
                                    D(SIAFD_B200_F_MASK), D(SIAFD_B200_F_THK_BC_MASK), D(SIAFD_B200_F_SMB),
                                    D(SIAFD_B200_F_BASAL_MELT), D(SIAFD_B200_F_EFF_SMB), D(SIAFD_B200_F_EFF_BMB),
                                    h->stream);
  CU(h, cudaGetLastError());
  return SIAFD_B200_OK;
}

int siafd_b200_ensure_consistency(siafd_b200_handle *h, int wrap_thickness) {
  if (!h) return null_handle();
  h->cfl3_fresh = false; // the fields the fused CFL maxima were taken on are about to change
  CU(h, cudaSetDevice(h->device));
  const int need[] = {SIAFD_B200_F_THICKNESS, SIAFD_B200_F_BED, SIAFD_B200_F_MASK, SIAFD_B200_F_SURFACE};
  for (int f : need) {
    int st = ensure(h, f);
    if (st) return st;
  }
  if (wrap_thickness) {
    const int f = SIAFD_B200_F_THICKNESS;
    int st = siafd_b200_wrap_ghosts_many(h, 1, &f);
    if (st) return st;
  }
  auto D = [&](int f) { return (double *)h->buf[f]; };
  h->launches += launch_consistency(h->P, (long)siafd_b200_field_size(h, SIAFD_B200_F_THICKNESS),
                                    D(SIAFD_B200_F_SEA_LEVEL), D(SIAFD_B200_F_BED), D(SIAFD_B200_F_THICKNESS),
                                    D(SIAFD_B200_F_MASK), D(SIAFD_B200_F_SURFACE), h->d_err, h->stream);
  CU(h, cudaGetLastError());
  h->result_pending = true;
  return SIAFD_B200_OK;
}

// ---- SURVEY.md 8(f) N3 (CFL part) ----------------------------------------------------------------------------
int siafd_b200_cfl(siafd_b200_handle *h, double max_dt_seconds, int do_3d, double *out8) {
  if (!h) return null_handle();
  CU(h, cudaSetDevice(h->device));
  if (!out8) return fail(h, SIAFD_B200_ERR_BAD_ARGUMENT, "out8 is NULL");
  int st;
  if ((st = ensure(h, SIAFD_B200_F_THICKNESS)) || (st = ensure(h, SIAFD_B200_F_MASK))) return st;
  if (do_3d) {
    const int need[] = {SIAFD_B200_F_U, SIAFD_B200_F_V, SIAFD_B200_F_W};
    for (int f : need) {
      if ((st = ensure(h, f))) return st;
    }
  }
  if ((st = ensure_cfl(h))) return st;
  // the 3D maxima may already be there: the vertical-velocity kernel takes them on the fly
  const bool run_3d = do_3d && !h->cfl3_fresh;
  CU(h, cudaMemsetAsync(h->d_cfl + (run_3d ? 0 : 4), 0, (run_3d ? 8 : 4) * sizeof(unsigned long long), h->stream));
  auto D = [&](int f) { return (double *)h->buf[f]; };
  h->launches += launch_cfl(h->P, run_3d, D(SIAFD_B200_F_THICKNESS), D(SIAFD_B200_F_MASK), D(SIAFD_B200_F_U),
                            D(SIAFD_B200_F_V), D(SIAFD_B200_F_W), h->d_z, D(SIAFD_B200_F_SLIDING), h->d_cfl, h->d_err,
                            h->stream);
  CU(h, cudaGetLastError());
  h->result_pending = true;
  CU(h, cudaMemcpyAsync(h->h_cfl, h->d_cfl, 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, h->stream));
  CU(h, cudaStreamSynchronize(h->stream));
  double m[8];
  std::memcpy(m, h->h_cfl, sizeof(m));
  // dt_max = min(max_dt, min_k 1 / denom_k) = min(max_dt, 1 / max_k denom_k): timestepping.cc:79-83, :139-142
  out8[0] = (m[0] > 0.0) ? std::min(max_dt_seconds, 1.0 / m[0]) : max_dt_seconds;
  out8[1] = m[1], out8[2] = m[2], out8[3] = m[3];
  out8[4] = (m[4] > 0.0) ? std::min(max_dt_seconds, 1.0 / m[4]) : max_dt_seconds;
  out8[5] = m[5], out8[6] = m[6], out8[7] = 0.0;
  return SIAFD_B200_OK;
}

int siafd_b200_finish(siafd_b200_handle *h) {
  if (!h) return null_handle();
  CU(h, cudaSetDevice(h->device));
  int st = fetch_result(h);
  if (st) return st;
  st = status_from_bits(h->h_res->err);
  if (st) {
    return fail(h, st, "%s", siafd_b200_status_string(st));
  }
  double dmax;
  std::memcpy(&dmax, &h->h_res->dmax, sizeof(double));
  if (dmax > h->cfg.D_limit) {
    // sia/SIAFD.cc:752-760
    return fail(h, SIAFD_B200_ERR_DIFFUSIVITY,
                "Maximum diffusivity of SIA flow (%f m2/s) is too high.\n"
                "This probably means that the bed elevation or the ice thickness is too rough.\n"
                "Increase stress_balance.sia.max_diffusivity to suppress this message.",
                dmax);
  }
  return SIAFD_B200_OK;
}

double siafd_b200_max_diffusivity(siafd_b200_handle *h) {
  if (!h) return NAN;
  if (cudaSetDevice(h->device) != cudaSuccess || fetch_result(h) != SIAFD_B200_OK) {
    return NAN;
  }
  double dmax;
  std::memcpy(&dmax, &h->h_res->dmax, sizeof(double));
  return dmax;
}

int siafd_b200_high_diffusivity_count(siafd_b200_handle *h) {
  if (!h) return -1;
  if (cudaSetDevice(h->device) != cudaSuccess || fetch_result(h) != SIAFD_B200_OK) {
    return -1;
  }
  return h->h_res->hdc;
}

// The plan of the call -- which pieces of the 3D arrays cross PCIe, what the host fills in itself -- is plain C++ in
// siafd_hostplan.hh (checked without a GPU through siafd_b200_host_plan_emulate); this file executes it.
#include "siafd_hostplan.hh"
using namespace siafd_hostplan;

int siafd_b200_host_levels_needed(const double *z, int Mz, double max_thickness) {
  return host_levels_needed(z, Mz, max_thickness);
}

int siafd_b200_host_plan_emulate(const siafd_b200_config *cfg, int rows_per_segment, int band, int sparse, int level_cut,
                                 int cut_cols, int patch, const double *thickness, const double *surface, const double *bed,
                                 const double *mask, const double *sliding, const double *enthalpy, double *enthalpy_dev,
                                 const double *u_dev, const double *v_dev, double *u, double *v, int64_t *h2d, int64_t *d2h) {
  if (!cfg || !cfg->z || rows_per_segment < 1 || !thickness || !surface || !mask || !enthalpy || !enthalpy_dev || !u_dev ||
      !v_dev || !u || !v || (level_cut && !bed)) {
    return SIAFD_B200_ERR_BAD_ARGUMENT;
  }
  const siafd_b200_config &c = *cfg;
  HostPlan P;
  plan_host_update(c, rows_per_segment, band, sparse != 0, level_cut != 0, patch != 0, cut_cols, level_cut > 1 ? level_cut - 1 : 0, 2, thickness, surface, bed, mask,
                   P);
  const int we = c.w_3d_in, wuv = c.w_uv, Mz = c.Mz;
  const long cellsE = c.xm + 2 * we, cellsUV = c.xm + 2 * wuv, rowUV = cellsUV * Mz;
  int64_t up = 0, dn = 0;
  for (const Piece &p : P.up_pieces) copy_piece_host(enthalpy_dev, enthalpy, cellsE, Mz, p), up += piece_bytes(p);
  if (!P.fills.empty()) fill_rows(c, sliding, u, v, P.fills.data(), P.fills.size(), 0, 1);
  for (const Piece &p : P.down_pieces) {
    copy_piece_host(u, u_dev, cellsUV, Mz, p), copy_piece_host(v, v_dev, cellsUV, Mz, p);
    dn += 2 * piece_bytes(p);
  }
  for (const ReplTask &T : P.repl) replicate_piece(T.p, u, v, cellsUV, Mz);
  // the end of update_host_pipelined: ghost rows, and the ghost columns of a patch's owned rows
  const double *src[2] = {u_dev, v_dev};
  double *dst[2] = {u, v};
  for (int q = 0; q < 2; ++q) {
    if (patch) {
      for (int j = wuv; j < c.ym + wuv; ++j) {
        for (int side = 0; side < 2; ++side) {
          const long off = (long)j * rowUV + (side ? (long)(c.xm + wuv) * Mz : 0L);
          std::memcpy(dst[q] + off, src[q] + off, (size_t)wuv * Mz * sizeof(double));
        }
      }
      dn += 2 * (int64_t)wuv * Mz * 8 * c.ym;
    }
    std::memcpy(dst[q], src[q], (size_t)wuv * rowUV * sizeof(double));
    const long off = (long)(wuv + c.ym) * rowUV;
    std::memcpy(dst[q] + off, src[q] + off, (size_t)wuv * rowUV * sizeof(double));
    dn += 2 * (int64_t)wuv * rowUV * 8;
  }
  if (h2d) *h2d = up;
  if (d2h) *d2h = dn;
  return SIAFD_B200_OK;
}

// one piece between a host and a device array of rows of `row_cells` columns (same local layout on both sides)
static cudaError_t copy_piece(double *dst, const double *src, long row_cells, int Mz, const Piece &p, cudaMemcpyKind kind,
                              cudaStream_t s) {
  const size_t colB = (size_t)Mz * sizeof(double);
  if (p.n >= Mz) {
    const long off = (long)p.r0 * row_cells * Mz + (long)p.c0 * Mz;
    if (p.c1 - p.c0 == row_cells) {
      return cudaMemcpyAsync(dst + off, src + off, (size_t)(p.r1 - p.r0) * row_cells * colB, kind, s);
    }
    return cudaMemcpy2DAsync(dst + off, (size_t)row_cells * colB, src + off, (size_t)row_cells * colB,
                             (size_t)(p.c1 - p.c0) * colB, (size_t)(p.r1 - p.r0), kind, s);
  }
  // a cut piece: lines of n levels at the pitch of a column.  Measured at 4096^2 (profiles/e2e_sweep_r02.json): one 3D
  // copy per piece of 64 rows x 256 / 1024 / 4096 columns: 350 / 392 / 425 ms per step; one 2D copy per row of 512 /
  // 1024 / 2048 columns: 465 / 440 / 418 ms -- although plain 2D copies of 65 k such lines run at 46 GB/s on their own
  // (tools/memcpy2d_bw.cu).  Small 3D pieces are the best of these.
  if (p.r1 - p.r0 == 1) {
    const long off = ((long)p.r0 * row_cells + p.c0) * Mz;
    return cudaMemcpy2DAsync(dst + off, colB, src + off, colB, (size_t)p.n * sizeof(double), (size_t)(p.c1 - p.c0), kind, s);
  }
  cudaMemcpy3DParms q;
  memset(&q, 0, sizeof(q));
  q.srcPtr = make_cudaPitchedPtr(const_cast<double *>(src), colB, colB, (size_t)row_cells);
  q.dstPtr = make_cudaPitchedPtr(dst, colB, colB, (size_t)row_cells);
  q.srcPos = make_cudaPos(0, (size_t)p.c0, (size_t)p.r0);
  q.dstPos = q.srcPos;
  q.extent = make_cudaExtent((size_t)p.n * sizeof(double), (size_t)(p.c1 - p.c0), (size_t)(p.r1 - p.r0));
  q.kind = kind;
  return cudaMemcpy3DAsync(&q, s);
}

// The host threads that replicate above the cut: the tasks of a band may start once its copies have finished (the
// band's event), which the issuing thread announces through `recorded` after it has recorded that event.
struct ReplShared {
  std::atomic<int> recorded{0}; // bands whose download event has been recorded
  std::atomic<bool> abort{false};
  std::atomic<int> failed{0};
};

static void replicate_pieces(int device, const cudaEvent_t *done, ReplShared *S, const ReplTask *tasks, size_t n, size_t first,
                             size_t stride, double *u, double *v, long row_cells, int Mz) {
  if (cudaSetDevice(device) != cudaSuccess) {
    S->failed = 1;
    return;
  }
  int synced = -1;
  for (size_t t = first; t < n; t += stride) {
    const ReplTask &T = tasks[t];
    if (T.band != synced) {
      while (S->recorded.load(std::memory_order_acquire) <= T.band) {
        if (S->abort.load(std::memory_order_acquire)) return;
        std::this_thread::sleep_for(std::chrono::microseconds(20));
      }
      if (cudaEventSynchronize(done[T.band]) != cudaSuccess) {
        S->failed = 1;
        return;
      }
      synced = T.band;
    }
    replicate_piece(T.p, u, v, row_cells, Mz);
  }
#if defined(__x86_64__) || defined(_M_X64)
  _mm_sfence();
#endif
}

static double wall_ms() {
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// SIAFD_B200_TRACE=1: host-clock timeline of one call on stderr (when each band's enthalpy was up, when its u, v were
// down, when the host threads were done) -- tools/e2e_sweep.py --trace
struct PipeTrace {
  bool on = false;
  double t0 = 0;
  std::vector<double> up, down, workers;
  std::atomic<int> up_recorded{0};
};

static void trace_watch(int device, const cudaEvent_t *ev, std::atomic<int> *recorded, std::atomic<bool> *abort, int n, double t0,
                        double *out) {
  if (cudaSetDevice(device) != cudaSuccess) return;
  for (int b = 0; b < n; ++b) {
    while (recorded->load(std::memory_order_acquire) <= b) {
      if (abort->load(std::memory_order_acquire)) return;
      std::this_thread::sleep_for(std::chrono::microseconds(20));
    }
    if (cudaEventSynchronize(ev[b]) != cudaSuccess) return;
    out[b] = wall_ms() - t0;
  }
}

static int update_host_pipelined(siafd_b200_handle *h, const siafd_b200_inputs *in, siafd_b200_outputs *out) {
  const siafd_b200_config &c = h->cfg;
  int st;
  PipeTrace tr;
  tr.on = getenv("SIAFD_B200_TRACE") != nullptr;
  tr.t0 = wall_ms();
  if (!h->s_up) {
    CU(h, cudaStreamCreateWithFlags(&h->s_up, cudaStreamNonBlocking));
    int lo_prio = 0, hi_prio = 0; // (the store kernel of the zero-copy path shares the SMs with the fused kernel)
    CU(h, cudaDeviceGetStreamPriorityRange(&lo_prio, &hi_prio));
    CU(h, cudaStreamCreateWithPriority(&h->s_dn, cudaStreamNonBlocking, hi_prio));
  }
  // With a communicator (siafd_b200_comm_init*, one rank or many) the ghost updates of h_x, h_y and u, v are stores by
  // the producing kernels into the neighbours' arrays (this rank's own where it is its own periodic neighbour), and the
  // status / D_max are reduced over all ranks.  patch: the neighbours in x are other ranks, so the ghost columns of
  // u, v arrive from them and are downloaded at the end, with the ghost rows.
  const bool comm = h->comm.active;
  const bool patch = comm && (c.xm != c.Mx || c.ym != c.My);
  // which parts of the 3D arrays have to move at all (sparse = 0 moves everything); the level cut: bed smoother off,
  // the bed given with this call; several ranks per host only when asked for (level_cut = 2) -- each link then carries
  // a fraction of the bytes and PCIe is not what bounds the call; not measured there
  const bool sparse = h->tuning.sparse_host != 0;
  const bool cut = sparse && in->bed && !(c.smoother_range > 0.0) && !h->smoother_set &&
                   (h->tuning.level_cut >= 2 || (h->tuning.level_cut == 1 && (!comm || h->comm.size == 1)));
  const int nseg = slab_segments(h->P, h->tuning), band = std::max(1, h->tuning.pipeline_band), NB = (nseg + band - 1) / band;
  while ((int)h->ev_pipe.size() < 3 * NB + 4) {
    cudaEvent_t e;
    CU(h, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    h->ev_pipe.push_back(e);
  }
  const int outs_f[] = {SIAFD_B200_F_H_X, SIAFD_B200_F_H_Y, SIAFD_B200_F_D, SIAFD_B200_F_FLUX, SIAFD_B200_F_U, SIAFD_B200_F_V,
                        SIAFD_B200_F_ENTHALPY};
  for (int f : outs_f) {
    if ((st = ensure(h, f))) return st;
  }
  cudaEvent_t ev_start = h->ev_pipe[2 * NB];
  CU(h, cudaEventRecord(ev_start, h->stream)); // earlier work on the handle's stream (fresh buffers' zero-fill)
  CU(h, cudaStreamWaitEvent(h->s_up, ev_start, 0));
  CU(h, cudaStreamWaitEvent(h->s_dn, ev_start, 0));
  // 2D inputs on the main stream (the gradient needs them first); they travel while the host makes its plan
  struct {
    int f;
    const double *p;
  } small[] = {{SIAFD_B200_F_SURFACE, in->surface}, {SIAFD_B200_F_THICKNESS, in->thickness}, {SIAFD_B200_F_MASK, in->mask},
               {SIAFD_B200_F_BED, in->bed}, {SIAFD_B200_F_SLIDING, in->sliding}};
  for (auto &q : small) {
    if (q.p && (st = siafd_b200_upload(h, q.f, q.p))) return st;
  }
  HostPlan plan;
  // (the scans of the plan run before the fill / replication threads start: up to 8 threads for a few milliseconds)
  const int plan_threads = std::max(h->fill_threads, (int)std::min(8u, std::max(1u, std::thread::hardware_concurrency())));
  plan_host_update(c, slab_rows_per_segment(h->tuning), band, sparse, cut, patch, h->tuning.cut_cols, h->tuning.cut_rows, plan_threads,
                   in->thickness, in->surface, in->bed, in->mask, plan);
  const double t_plan = wall_ms() - tr.t0;
  if (plan.nseg != nseg || plan.NB != NB) return fail(h, SIAFD_B200_ERR_BAD_ARGUMENT, "host plan and kernel disagree on the row segments");
  const int we = c.w_3d_in, wuv = c.w_uv;
  const long rowUV = (long)(c.xm + 2 * wuv) * c.Mz;
  // zero copy: the caller's u, v are pinned and mapped into the device's address space (cudaHostAlloc / cudaHostRegister
  // under unified addressing): a kernel stores the pieces there itself, whole aligned lines over PCIe, instead of the
  // copy engine's strided lines
  double *zc_uv[2] = {nullptr, nullptr};
  if (h->zero_copy && !plan.down_pieces.empty()) {
    double *hp[2] = {out->u, out->v};
    for (int q = 0; q < 2; ++q) {
      cudaPointerAttributes a;
      if (cudaPointerGetAttributes(&a, hp[q]) == cudaSuccess && a.type == cudaMemoryTypeHost && a.devicePointer != nullptr) {
        zc_uv[q] = (double *)a.devicePointer;
      } else {
        cudaGetLastError(); // (pageable memory: not an error of this call)
      }
    }
    if (!zc_uv[0] || !zc_uv[1]) zc_uv[0] = zc_uv[1] = nullptr;
  }
  static_assert(sizeof(StorePiece) == sizeof(Piece), "the device's view of a piece is the plan's");
  if (zc_uv[0]) {
    const size_t need = plan.down_pieces.size() * sizeof(Piece);
    if (need > h->d_pieces_bytes) {
      cudaFree(h->d_pieces);
      h->d_pieces = nullptr, h->d_pieces_bytes = 0;
      CU(h, cudaMalloc(&h->d_pieces, need * 2));
      h->d_pieces_bytes = need * 2;
    }
    CU(h, cudaMemcpyAsync(h->d_pieces, plan.down_pieces.data(), need, cudaMemcpyHostToDevice, h->s_dn));
  }
  // the host's own work is known up front, so its threads start before the copies: the fills at once, the tasks above
  // the cut as their bands come down
  std::vector<std::thread> workers;
  ReplShared repl_state;
  const std::vector<FillTask> &fills = plan.fills;
  const std::vector<ReplTask> &repl = plan.repl;
  const std::vector<Piece> &pieces = plan.down_pieces;
  const std::vector<size_t> &piece0 = plan.down0;
  if (!fills.empty()) {
    const size_t nt = std::min<size_t>(std::max(1u, std::min((unsigned)h->fill_threads, std::thread::hardware_concurrency())), fills.size());
    if (tr.on) tr.workers.assign(nt + 64, 0.0);
    for (size_t t = 0; t < nt; ++t) {
      if (tr.on) {
        workers.emplace_back([&, t, nt] {
          fill_rows(c, in->sliding, out->u, out->v, fills.data(), fills.size(), t, nt);
          tr.workers[t] = wall_ms() - tr.t0;
        });
      } else {
        workers.emplace_back(fill_rows, std::cref(c), in->sliding, out->u, out->v, fills.data(), fills.size(), t, nt);
      }
    }
  }
  const cudaEvent_t *ev_down = h->ev_pipe.data() + 2 * NB + 4; // band b's copies of u, v have finished
  if (!repl.empty()) {
    const size_t nt = std::min<size_t>(std::max(1u, std::min((unsigned)h->repl_threads, std::thread::hardware_concurrency())), repl.size());
    for (size_t t = 0; t < nt; ++t) {
      workers.emplace_back(replicate_pieces, h->device, ev_down, &repl_state, repl.data(), repl.size(), t, nt, out->u, out->v,
                           (long)(c.xm + 2 * wuv), c.Mz);
    }
  }
  if (tr.on) { // two watchers: when was band b's enthalpy up, when were its u, v down
    tr.up.assign(NB, 0.0), tr.down.assign(NB, 0.0);
    workers.emplace_back(trace_watch, h->device, h->ev_pipe.data(), &tr.up_recorded, &repl_state.abort, NB, tr.t0, tr.up.data());
    workers.emplace_back(trace_watch, h->device, ev_down, &repl_state.recorded, &repl_state.abort, NB, tr.t0, tr.down.data());
  }
  struct Joiner { // the workers are joined on every way out of this function
    std::vector<std::thread> &w;
    ReplShared &S;
    bool released = false;
    void join() {
      for (auto &t : w)
        if (t.joinable()) t.join();
    }
    ~Joiner() {
      if (!released) S.abort = true; // an early return: tasks that still wait for their band give up
      join();
    }
  } joiner{workers, repl_state};

  // enthalpy bands on the upload stream, issued a few bands ahead of the kernel that reads them (with the level cut a
  // band is hundreds of copies: issuing them all up front would hold the first kernel back)
  double *E_dev = (double *)h->buf[SIAFD_B200_F_ENTHALPY];
  int up_issued = 0;
  auto issue_uploads = [&](int upto) -> int { // bands [up_issued, upto)
    for (; up_issued < std::min(upto, NB); ++up_issued) {
      const int b = up_issued;
      for (size_t t = plan.up0[b]; t < plan.up0[b + 1]; ++t) {
        CU(h, copy_piece(E_dev, in->enthalpy, c.xm + 2 * we, c.Mz, plan.up_pieces[t], cudaMemcpyHostToDevice, h->s_up));
        h->bytes_h2d += piece_bytes(plan.up_pieces[t]);
      }
      CU(h, cudaEventRecord(h->ev_pipe[b], h->s_up));
      tr.up_recorded.store(b + 1, std::memory_order_release);
    }
    return SIAFD_B200_OK;
  };
  if ((st = issue_uploads(plan.up_pieces.size() > 4 * (size_t)NB ? 3 : NB))) return st;
  const double t_up_issued = wall_ms() - tr.t0;
  // gradient and 2D preparation while the first band is in flight
  PeerPush PPu = PeerPush();
  bool prep_done = false;
  if (comm) {
    const int more[] = {SIAFD_B200_F_W_I, SIAFD_B200_F_W_J};
    for (int f : more) {
      if ((st = ensure(h, f))) return st;
    }
    PeerPush PPg;
    comm_make_push(h, SIAFD_B200_F_H_X, SIAFD_B200_F_H_Y, c.w_stag, 1, PPg);
    comm_make_push(h, SIAFD_B200_F_U, SIAFD_B200_F_V, c.w_uv, 1, PPu);
    const bool haseloff = c.gradient_method == SIAFD_B200_GRAD_HASELOFF;
    // SIAFD.cc:137, :498-499; haseloff: thk_smooth / theta in the same pass (no row-segment weights: the bands keep
    // the plain order)
    h->P.seg_n = 0;
    h->launches += launch_gradient(h->P, fields_of(h), h->stream, haseloff ? &PPg : nullptr, haseloff);
    CU(h, cudaGetLastError());
    if (haseloff && h->comm.size > 1 && !gradient_ring_is_local(h->P)) {
      h->launches += launch_comm_sync(h->comm.d_peers, 2, h->stream);
    }
    prep_done = haseloff;
  } else {
    if ((st = siafd_b200_compute_gradient(h))) return st;
    if (c.gradient_method == SIAFD_B200_GRAD_HASELOFF) { // sia/SIAFD.cc:498-499
      const int hxy[2] = {SIAFD_B200_F_H_X, SIAFD_B200_F_H_Y};
      if ((st = siafd_b200_wrap_ghosts_many(h, 2, hxy))) return st;
    }
  }
  if ((st = flux_velocity_prepare(h, 1, in->current_time, prep_done))) return st;
  const int uvf[2] = {SIAFD_B200_F_U, SIAFD_B200_F_V};
  double *uvh[2] = {out->u, out->v};
  for (int b = 0; b < NB; ++b) {
    const int s0 = b * band, s1 = std::min(nseg, (b + 1) * band);
    if ((st = issue_uploads(b + 4))) return st;
    CU(h, cudaStreamWaitEvent(h->stream, h->ev_pipe[b], 0));
    if ((st = flux_velocity_launch(h, 1, s0, s1 - s0, comm ? &PPu : nullptr))) return st;
    // owned rows of this band (extended row e = ys - 1 + s RS ... ; owned rows are ys .. ys + ym - 1)
    const HostPlan::Rect &R = plan.down[b];
    const int o0 = R.o0, o1 = R.o1;
    if (o1 > o0) {
      for (int q = 0; q < 2 && !comm; ++q) { // periodic wrap in x of the band's rows (SIAFD.cc:946-947), then download
        double *a = (double *)h->buf[uvf[q]];
        h->launches += launch_copy_region(a, c.xm + 2 * wuv, 0, wuv + o0, a, c.xm + 2 * wuv, c.xm, wuv + o0, wuv, o1 - o0, c.Mz,
                                          h->stream);
        h->launches += launch_copy_region(a, c.xm + 2 * wuv, c.xm + wuv, wuv + o0, a, c.xm + 2 * wuv, wuv, wuv + o0, wuv, o1 - o0,
                                          c.Mz, h->stream);
      }
      CU(h, cudaEventRecord(h->ev_pipe[NB + b], h->stream));
      CU(h, cudaStreamWaitEvent(h->s_dn, h->ev_pipe[NB + b], 0));
      if (zc_uv[0]) {
        h->launches += launch_store_pieces((const double *)h->buf[uvf[0]], (const double *)h->buf[uvf[1]], zc_uv[0], zc_uv[1],
                                           (const StorePiece *)h->d_pieces, (int)piece0[b], (int)piece0[b + 1], c.xm + 2 * wuv, c.Mz,
                                           h->s_dn);
        CU(h, cudaGetLastError());
        for (size_t t = piece0[b]; t < piece0[b + 1]; ++t) h->bytes_d2h += 2 * piece_bytes(pieces[t]);
      } else {
        for (size_t t = piece0[b]; t < piece0[b + 1]; ++t) {
          const Piece &p = pieces[t];
          for (int q = 0; q < 2; ++q) {
            CU(h, copy_piece(uvh[q], (const double *)h->buf[uvf[q]], c.xm + 2 * wuv, c.Mz, p, cudaMemcpyDeviceToHost, h->s_dn));
            h->bytes_d2h += piece_bytes(p);
          }
        }
      }
    }
    if (!repl.empty() || tr.on) { // (also for a band without copies: the counter is the number of bands issued)
      CU(h, cudaEventRecord(ev_down[b], h->s_dn));
      repl_state.recorded.store(b + 1, std::memory_order_release);
    }
  }
  // ghost rows of u, v (periodic wrap in y / the neighbours' stores), the 2D outputs, D_max and the error flags
  if (comm) {
    // wait for the neighbours' u, v; {D_max, error bits, counter} over all ranks (SIAFD.cc:748-750)
    h->launches += launch_comm_final(h->comm.d_peers, h->comm.size > 1 ? 3 : -1, h->d_dmax, h->d_err, h->d_hdc, h->comm.d_res,
                                     h->comm.h_res, h->stream);
    CU(h, cudaGetLastError());
    h->result_pending = false;
    h->comm.result_from_comm = true;
  }
  for (int q = 0; q < 2; ++q) {
    if (!comm && (st = wrap_dir(h, uvf[q], 1))) return st;
    double *a = (double *)h->buf[uvf[q]];
    if (patch) { // ghost columns of the owned rows: they arrived from the x neighbours
      for (int side = 0; side < 2; ++side) {
        const long off = (long)wuv * rowUV + (side ? (long)(c.xm + wuv) * c.Mz : 0L);
        CU(h, cudaMemcpy2DAsync(uvh[q] + off, (size_t)rowUV * sizeof(double), a + off, (size_t)rowUV * sizeof(double),
                                (size_t)wuv * c.Mz * sizeof(double), (size_t)c.ym, cudaMemcpyDeviceToHost, h->stream));
      }
      h->bytes_d2h += 2 * (int64_t)wuv * c.Mz * 8 * c.ym;
    }
    CU(h, cudaMemcpyAsync(uvh[q], a, (size_t)wuv * rowUV * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    h->bytes_d2h += 2 * (int64_t)wuv * rowUV * 8;
    const long off = (long)(wuv + c.ym) * rowUV;
    CU(h, cudaMemcpyAsync(uvh[q] + off, a + off, (size_t)wuv * rowUV * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  }
  struct {
    int f;
    double *p;
  } outs2[] = {{SIAFD_B200_F_H_X, out->h_x}, {SIAFD_B200_F_H_Y, out->h_y}, {SIAFD_B200_F_D, out->D}, {SIAFD_B200_F_FLUX, out->flux}};
  for (auto &q : outs2) {
    if (!q.p) continue;
    CU(h, cudaMemcpyAsync(q.p, h->buf[q.f], (size_t)siafd_b200_field_size(h, q.f) * sizeof(double), cudaMemcpyDeviceToHost,
                          h->stream));
    h->bytes_d2h += (int64_t)siafd_b200_field_size(h, q.f) * 8;
  }
  const double t_issued = wall_ms() - tr.t0;
  st = siafd_b200_finish(h);
  const double t_finish = wall_ms() - tr.t0;
  CU(h, cudaStreamSynchronize(h->s_dn));
  CU(h, cudaStreamSynchronize(h->s_up));
  const double t_streams = wall_ms() - tr.t0;
  joiner.released = true;
  joiner.join();
  if (tr.on) {
    double wmax = 0;
    for (double w : tr.workers) wmax = std::max(wmax, w);
    fprintf(stderr, "siafd_b200 trace [ms]: plan %.1f, uploads issued %.1f, all issued %.1f, main stream done %.1f, copy streams done %.1f, "
                    "host threads joined %.1f (fill threads done %.1f); bands %d, pieces up %zu down %zu\n",
            t_plan, t_up_issued, t_issued, t_finish, t_streams, wall_ms() - tr.t0, wmax, NB, plan.up_pieces.size(), plan.down_pieces.size());
    fprintf(stderr, "  band:       ");
    for (int b = 0; b < NB; b += std::max(1, NB / 16)) fprintf(stderr, "%7d", b);
    fprintf(stderr, "\n  E up at:    ");
    for (int b = 0; b < NB; b += std::max(1, NB / 16)) fprintf(stderr, "%7.1f", tr.up[b]);
    fprintf(stderr, "\n  u, v down at:");
    for (int b = 0; b < NB; b += std::max(1, NB / 16)) fprintf(stderr, "%7.1f", tr.down[b]);
    fprintf(stderr, "  last %.1f\n", tr.down[NB - 1]);
  }
  if (repl_state.failed.load()) return fail(h, SIAFD_B200_ERR_CUDA, "a host thread of the level cut could not wait for its band");
  return st;
}

int siafd_b200_update(siafd_b200_handle *h, const siafd_b200_inputs *in, siafd_b200_outputs *out, int full_update) {
  if (!h) return null_handle();
  h->cfl3_fresh = false; // the fields the fused CFL maxima were taken on are about to change
  if (!h || !in || !out) {
    return fail(h, SIAFD_B200_ERR_BAD_ARGUMENT, "null argument");
  }
  CU(h, cudaSetDevice(h->device));
  const siafd_b200_config &c = h->cfg;
  const bool whole = (c.xm == c.Mx && c.ym == c.My);
  if (!whole && !h->comm.active) {
    return fail(h, SIAFD_B200_ERR_BAD_ARGUMENT,
                "siafd_b200_update on a patch of a decomposed domain needs a communicator (siafd_b200_comm_init), or "
                "the split calls with ghost exchanges at SIAFD.cc:498-499 and :946-947");
  }
  if (!whole && !in->ghosts_valid && in->memory_space == 0) {
    return fail(h, SIAFD_B200_ERR_BAD_ARGUMENT, "host arrays of a patch must carry valid ghosts (PISM's always do)");
  }
  if (in->memory_space == 0 && out->memory_space == 0 && full_update && in->ghosts_valid && in->enthalpy && !in->age &&
      out->u && out->v && in->surface && in->thickness && in->mask && h->tuning.pipeline_host) {
    return update_host_pipelined(h, in, out);
  }
  struct {
    int f;
    const double *p;
  } ins[] = {{SIAFD_B200_F_SURFACE, in->surface},   {SIAFD_B200_F_THICKNESS, in->thickness},
             {SIAFD_B200_F_MASK, in->mask},         {SIAFD_B200_F_BED, in->bed},
             {SIAFD_B200_F_ENTHALPY, in->enthalpy}, {SIAFD_B200_F_AGE, in->age},
             {SIAFD_B200_F_SLIDING, in->sliding}};
  int st;
  for (auto &q : ins) {
    if (!q.p) {
      continue;
    }
    if (in->memory_space == 0) {
      if ((st = siafd_b200_upload(h, q.f, q.p))) return st;
    } else {
      if ((st = siafd_b200_bind(h, q.f, (void *)q.p))) return st;
    }
    if (!in->ghosts_valid && !h->comm.active) {
      if ((st = siafd_b200_wrap_ghosts(h, q.f))) return st;
    }
  }
  if (!in->surface || !in->thickness || !in->mask || !in->enthalpy) {
    if (!h->buf[SIAFD_B200_F_SURFACE] || !h->buf[SIAFD_B200_F_THICKNESS] || !h->buf[SIAFD_B200_F_MASK] ||
        !h->buf[SIAFD_B200_F_ENTHALPY]) {
      return fail(h, SIAFD_B200_ERR_BAD_ARGUMENT, "surface, thickness, mask and enthalpy are required");
    }
  }
  struct {
    int f;
    double *p;
  } outs[] = {{SIAFD_B200_F_H_X, out->h_x}, {SIAFD_B200_F_H_Y, out->h_y}, {SIAFD_B200_F_D, out->D},
              {SIAFD_B200_F_FLUX, out->flux}, {SIAFD_B200_F_U, out->u},   {SIAFD_B200_F_V, out->v}};
  if (out->memory_space != 0) {
    for (auto &q : outs) {
      if (q.p && (st = siafd_b200_bind(h, q.f, q.p))) return st;
    }
  }
  if (h->comm.active) {
    // one rank of a decomposed run (or a whole-domain handle with a one-rank communicator): every ghost update inside
    if (out->memory_space != 0 || in->memory_space != 0) {
      return fail(h, SIAFD_B200_ERR_BAD_ARGUMENT,
                  "with a communicator the fields must live in the handle's own storage: use host pointers here, or "
                  "siafd_b200_device_ptr + siafd_b200_update_decomposed for device-resident fields");
    }
    if ((st = siafd_b200_update_decomposed(h, full_update, in->current_time, in->ghosts_valid ? 0 : 1))) return st;
  } else {
    if ((st = siafd_b200_compute_gradient(h))) return st;
    if (c.gradient_method == SIAFD_B200_GRAD_HASELOFF) { // sia/SIAFD.cc:498-499
      const int hxy[2] = {SIAFD_B200_F_H_X, SIAFD_B200_F_H_Y};
      if ((st = siafd_b200_wrap_ghosts_many(h, 2, hxy))) return st;
    }
    if ((st = siafd_b200_compute_flux_velocity(h, full_update, in->current_time))) return st;
    if (full_update) { // sia/SIAFD.cc:946-947
      const int uv[2] = {SIAFD_B200_F_U, SIAFD_B200_F_V};
      if ((st = siafd_b200_wrap_ghosts_many(h, 2, uv))) return st;
    }
  }
  if (out->memory_space == 0) {
    for (auto &q : outs) {
      if (!q.p) continue;
      if (!full_update && (q.f == SIAFD_B200_F_U || q.f == SIAFD_B200_F_V)) continue; // G10: untouched
      const size_t bytes = (size_t)siafd_b200_field_size(h, q.f) * sizeof(double);
      CU(h, cudaMemcpyAsync(q.p, h->buf[q.f], bytes, cudaMemcpyDeviceToHost, h->stream));
    }
  }
  return siafd_b200_finish(h);
}

int siafd_b200_geometry_compute(siafd_b200_handle *h, int64_t n, const double *sea_level_dev, const double *bed_dev,
                                const double *thickness_dev, double *mask_out_dev, double *surface_out_dev) {
  if (!h) return null_handle();
  CU(h, cudaSetDevice(h->device));
  h->launches += launch_geometry(h->P, n, sea_level_dev, bed_dev, thickness_dev, mask_out_dev, surface_out_dev, h->stream);
  CU(h, cudaGetLastError());
  CU(h, cudaStreamSynchronize(h->stream));
  return SIAFD_B200_OK;
}

int siafd_b200_flow_n(siafd_b200_handle *h, int64_t n, const double *stress_dev, const double *enthalpy_dev,
                      const double *pressure_dev, const double *grainsize_dev, double *result_dev) {
  if (!h) return null_handle();
  CU(h, cudaSetDevice(h->device));
  const int k = launch_flow_n(h->P, n, stress_dev, enthalpy_dev, pressure_dev, grainsize_dev, result_dev, h->stream);
  if (k < 0) {
    return fail(h, SIAFD_B200_ERR_BAD_CONFIG, "unknown flow law");
  }
  h->launches += k;
  CU(h, cudaGetLastError());
  CU(h, cudaStreamSynchronize(h->stream));
  return SIAFD_B200_OK;
}

int siafd_b200_flow_host(siafd_b200_handle *h, int64_t n, const double *stress, const double *enthalpy,
                         const double *pressure, const double *grainsize, double *result) {
  if (!h) return null_handle();
  if (n < 1 || !stress || !enthalpy || !pressure || !result) return fail(h, SIAFD_B200_ERR_BAD_ARGUMENT, "flow_host: null argument");
  CU(h, cudaSetDevice(h->device));
  double *d = nullptr;
  CU(h, cudaMalloc(&d, (size_t)n * 5 * sizeof(double)));
  std::vector<double> gs((size_t)n, h->cfg.grain_size);
  if (grainsize) gs.assign(grainsize, grainsize + n);
  const double *src[4] = {stress, enthalpy, pressure, gs.data()};
  cudaError_t e = cudaSuccess;
  for (int q = 0; q < 4 && e == cudaSuccess; ++q) {
    e = cudaMemcpyAsync(d + (size_t)q * n, src[q], (size_t)n * sizeof(double), cudaMemcpyHostToDevice, h->stream);
  }
  if (e == cudaSuccess) {
    h->launches += launch_flow_n(h->P, (long)n, d, d + n, d + 2 * n, d + 3 * n, d + 4 * n, h->stream);
    e = cudaMemcpyAsync(result, d + 4 * n, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, h->stream);
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  cudaFree(d);
  if (e != cudaSuccess) return fail(h, SIAFD_B200_ERR_CUDA, "flow_host: %s", cudaGetErrorString(e));
  return SIAFD_B200_OK;
}

int siafd_b200_set_tuning(siafd_b200_handle *h, int rows_per_cta, int use_bulk_copy, int skip_ice_free_rows) {
  if (!h) return null_handle();
  if (rows_per_cta > 0) h->tuning.rows_per_cta = rows_per_cta;
  if (use_bulk_copy >= 0) h->tuning.use_bulk_copy = use_bulk_copy ? 1 : 0;
  if (skip_ice_free_rows >= 0) h->tuning.skip_ice_free = skip_ice_free_rows ? 1 : 0;
  invalidate_graphs(h);
  return SIAFD_B200_OK;
}

int64_t siafd_b200_launch_count(const siafd_b200_handle *h) {
  if (!h) return -1; return h->launches; }

int siafd_b200_transfer_bytes(const siafd_b200_handle *h, int64_t *h2d, int64_t *d2h) {
  if (!h) return null_handle();
  if (h2d) *h2d = h->bytes_h2d;
  if (d2h) *d2h = h->bytes_d2h;
  return SIAFD_B200_OK;
}

int siafd_b200_kernel_timing(siafd_b200_handle *h, int enable) {
  if (!h) return null_handle();
  CU(h, cudaSetDevice(h->device));
  if (enable && h->ev_start.empty()) {
    h->ev_start.resize(256);
    h->ev_stop.resize(256);
    for (size_t q = 0; q < h->ev_start.size(); ++q) {
      CU(h, cudaEventCreate(&h->ev_start[q]));
      CU(h, cudaEventCreate(&h->ev_stop[q]));
    }
  }
  if (enable && h->ev_sec.empty()) {
    h->ev_sec.resize(6 * 64);
    for (auto &e : h->ev_sec) CU(h, cudaEventCreate(&e));
  }
  h->timing = enable != 0;
  h->ev_count = 0;
  h->sec_count = 0;
  return SIAFD_B200_OK;
}

int siafd_b200_step_breakdown_ms(siafd_b200_handle *h, double *out5, int *steps_out) {
  if (!h) return null_handle();
  if (!out5) return fail(h, SIAFD_B200_ERR_BAD_ARGUMENT, "null argument");
  CU(h, cudaSetDevice(h->device));
  CU(h, cudaStreamSynchronize(h->stream));
  for (int q = 0; q < 5; ++q) out5[q] = 0.0;
  for (int n = 0; n < h->sec_count; ++n) {
    for (int q = 0; q < 5; ++q) {
      float ms = 0.f;
      CU(h, cudaEventElapsedTime(&ms, h->ev_sec[6 * n + q], h->ev_sec[6 * n + q + 1]));
      out5[q] += ms;
    }
  }
  for (int q = 0; q < 5 && h->sec_count > 0; ++q) out5[q] /= h->sec_count;
  if (steps_out) *steps_out = h->sec_count;
  h->sec_count = 0;
  return SIAFD_B200_OK;
}

double siafd_b200_kernel_time_ms(siafd_b200_handle *h, int *launches_out) {
  if (!h) return NAN;
  if (cudaSetDevice(h->device) != cudaSuccess || cudaStreamSynchronize(h->stream) != cudaSuccess) {
    return -1.0;
  }
  double total = 0.0;
  for (int q = 0; q < h->ev_count; ++q) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, h->ev_start[q], h->ev_stop[q]) != cudaSuccess) {
      return -1.0;
    }
    total += ms;
  }
  if (launches_out) {
    *launches_out = h->ev_count;
  }
  h->ev_count = 0;
  return total;
}

} // extern "C"
