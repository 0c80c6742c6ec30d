// siafd_mass.cu -- the consumers of SIAFD's outputs (SURVEY.md 8(f) N1 and the CFL part of N3):
//   k_mass_flow      GeometryEvolution::flow_step  geometry/GeometryEvolution.cc:241-324 (part_grid off):
//                    gc.compute -> compute_interface_fluxes :535-654 (limit_diffusive_flux :462-525,
//                    limit_advective_velocity :395-457) -> compute_flux_divergence :660-688 -> update_in_place
//                    :716-771 -> thickness_change :295-300 -> ensure_nonnegativity :960-1000
//   k_mass_apply     GeometryEvolution::apply_flux_divergence :347-350
//   k_mass_source    source_term_step :327-343 + apply_mass_fluxes :360-390 (effective_change :1005-1011)
//   k_consistency    Geometry::ensure_consistency geometry/Geometry.cc:121-187 (mask + surface; H < 0 is an error)
//   k_cfl            max_timestep_cfl_3d / _2d  stressbalance/timestepping.cc:42-101, :113-153
// All are 2D, a few tens of bytes per column, HBM-bound and tiny next to k_sia_slab (2.5 KB per column) -- except
// k_cfl's 3D part, which re-reads u, v, w of the icy columns below the surface (<= 24 Mz bytes per column).
// Every expression whose rounding the reference fixes is written with explicit __dadd_rn / __dmul_rn so that nvcc
// cannot contract it into an FMA: thickness, masks and the CFL scalars are BIT-EXACT against the oracle.
#include "siafd_kernels.cuh"

#define FULLMASK 0xffffffffu

namespace siafd {

namespace {

__host__ __device__ inline unsigned nblk_(long n, int b) { return (unsigned)((n + b - 1) / b); }

// util/Mask.hh:96-133 (mask only)
__device__ __forceinline__ int gc_mask(const DP &P, double sea_level, double bed, double thk) {
  const double hgrounded = __dadd_rn(bed, thk);
  const double hfloating = __dadd_rn(sea_level, __dmul_rn(P.gc_alpha, thk));
  const bool is_floating = (hfloating > hgrounded), ice_free = (thk <= P.gc_icefree);
  if (is_floating && !P.gc_dry) {
    return ice_free ? 4 : 3;
  }
  return ice_free ? 0 : 2;
}

__device__ __forceinline__ bool grounded_ice(int M) { return m_icy(M) && m_grounded(M); }
__device__ __forceinline__ bool ice_free_land(int M) { return m_grounded(M) && m_ice_free(M); }

// GeometryEvolution.cc:462-525: the sixteen cases reduce to "a grounded icy cell on either side"
__device__ __forceinline__ double limit_diffusive_flux(int a, int b, double flux) {
  return (grounded_ice(a) || grounded_ice(b)) ? flux : 0.0;
}

// GeometryEvolution.cc:395-457: zero between floating ice and ice-free land and between two ice-free cells
__device__ __forceinline__ double limit_advective_velocity(int a, int b, double v) {
  if (grounded_ice(a) || grounded_ice(b)) return v;                                  // cases 1-7
  if (m_floating_ice(a) && m_floating_ice(b)) return v;                              // case 8
  if ((m_floating_ice(a) && ice_free_land(b)) || (ice_free_land(a) && m_floating_ice(b))) return 0.0; // 9, 10
  if ((m_floating_ice(a) && m_ice_free_ocean(b)) || (m_ice_free_ocean(a) && m_floating_ice(b))) return v; // 11, 12
  return 0.0;                                                                        // 13-16
}

// one interface of compute_interface_fluxes (:553-645): cell (M, H, V, BC) and its neighbour in +x or +y
__device__ __forceinline__ double interface_flux(int M, int M_n, double H, double H_n, double V, double V_n, int BC,
                                                 int BC_n, double Q_sia) {
  double v = 0.0;
  if (m_icy(M) && m_icy(M_n)) {
    v = __dmul_rn(0.5, __dadd_rn(V, V_n));
  } else if (m_icy(M) && m_ice_free(M_n)) {
    v = V;
  } else if (m_ice_free(M) && m_icy(M_n)) {
    v = V_n;
  }
  if (BC == 1 && BC_n == 1) {
    v = __dmul_rn(0.5, __dadd_rn(V, V_n));
  } else if (BC == 1 && BC_n == 0) {
    v = V;
  } else if (BC == 0 && BC_n == 1) {
    v = V_n;
  }
  v = limit_advective_velocity(M, M_n, v);
  const double Q_advective = __dmul_rn(v, (v > 0.0 ? H : H_n)); // first order upwinding
  const double Q_diffusive = limit_diffusive_flux(M, M_n, Q_sia);
  return __dadd_rn(Q_diffusive, Q_advective);
}

struct MassArgs {
  const double *H, *bed, *sea, *vel, *vel_bc, *thk_bc, *Q;
  double *flux_div, *dH, *cons_err;
  double dt;
};

__global__ void k_mass_flow(const __grid_constant__ DP P, const MassArgs A) {
  const long q = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= (long)P.xm * P.ym) return;
  const int i = P.xs + (int)(q % P.xm), j = P.ys + (int)(q / P.xm);
  // the five cells of the star stencil: centre, east, west, north, south
  const int di[5] = {0, 1, -1, 0, 0}, dj[5] = {0, 0, 0, 1, -1};
  int M[5], BC[5];
  double H[5], U[5], V[5];
#pragma unroll
  for (int s = 0; s < 5; ++s) {
    const long g = idx2(P, i + di[s], j + dj[s], P.wg);
    H[s] = A.H[g];
    M[s] = gc_mask(P, A.sea ? A.sea[g] : 0.0, A.bed[g], H[s]); // :262-266
    BC[s] = A.vel_bc ? mask_int(A.vel_bc[g]) : 0;
    if (A.vel) {
      const long gv = idx2(P, i + di[s], j + dj[s], P.wsl) * 2;
      U[s] = A.vel[gv], V[s] = A.vel[gv + 1];
    } else {
      U[s] = 0.0, V[s] = 0.0;
    }
  }
  const long sc = idx2(P, i, j, P.wst) * 2, sw = idx2(P, i - 1, j, P.wst) * 2, ss = idx2(P, i, j - 1, P.wst) * 2;
  // interfaces: east = (i,j) o=0; west = (i-1,j) o=0; north = (i,j) o=1; south = (i,j-1) o=1
  const double Qe = interface_flux(M[0], M[1], H[0], H[1], U[0], U[1], BC[0], BC[1], A.Q[sc]);
  const double Qw = interface_flux(M[2], M[0], H[2], H[0], U[2], U[0], BC[2], BC[0], A.Q[sw]);
  const double Qn = interface_flux(M[0], M[3], H[0], H[3], V[0], V[3], BC[0], BC[3], A.Q[sc + 1]);
  const double Qs = interface_flux(M[4], M[0], H[4], H[0], V[4], V[0], BC[4], BC[0], A.Q[ss + 1]);
  double divQ;
  if (A.thk_bc && A.thk_bc[idx2(P, i, j, P.wg)] > 0.5) {
    divQ = 0.0;
  } else { // :680
    divQ = __dadd_rn(__ddiv_rn(__dsub_rn(Qe, Qw), P.dx), __ddiv_rn(__dsub_rn(Qn, Qs), P.dy));
  }
  A.flux_div[q] = divQ;
  const double H_old = H[0];
  const double H_new = __dadd_rn(H_old, __dmul_rn(-A.dt, divQ)); // :764
  double dH = __dsub_rn(H_new, H_old);                          // :297
  double ce = 0.0;
  if (__dadd_rn(H_old, dH) < 0.0) { // :980-983 (sic: the reference assigns H, not -H)
    ce = -__dadd_rn(H_old, dH);
    dH = H_old;
  }
  A.dH[q] = dH;
  A.cons_err[q] = ce;
}

__global__ void k_mass_apply(const __grid_constant__ DP P, double *H, const double *dH) {
  const long q = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= (long)P.xm * P.ym) return;
  const int i = P.xs + (int)(q % P.xm), j = P.ys + (int)(q / P.xm);
  const long g = idx2(P, i, j, P.wg);
  H[g] = __dadd_rn(H[g], dH[q]);
}

// GeometryEvolution.cc:1005-1011
__device__ __forceinline__ double effective_change(double H, double dH) { return (__dadd_rn(H, dH) <= 0) ? -H : dH; }

__global__ void k_mass_source(const __grid_constant__ DP P, double dt, double ice_density, int use_bmr, double *H,
                              const double *mask, const double *thk_bc, const double *smb, const double *bmr,
                              double *eff_smb, double *eff_bmb) {
  const long q = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= (long)P.xm * P.ym) return;
  const int i = P.xs + (int)(q % P.xm), j = P.ys + (int)(q / P.xm);
  const long g = idx2(P, i, j, P.wg);
  const int M = mask_int(mask[g]);
  const int bc = thk_bc ? mask_int(thk_bc[g]) : 0;
  double dS = 0.0, dB = 0.0;
  if (!(bc == 1 || m_ice_free_ocean(M))) {
    const double Hc = H[g];
    dS = effective_change(Hc, __ddiv_rn(__dmul_rn(dt, smb[q]), ice_density)); // :1058
    dB = effective_change(__dadd_rn(Hc, dS), __dmul_rn(dt, (use_bmr ? -bmr[q] : 0.0)));
    H[g] = __dadd_rn(__dadd_rn(Hc, dS), dB); // :375
  }
  eff_smb[q] = dS;
  eff_bmb[q] = dB;
}

// Geometry.cc:121-187 on every local point (ghosts included: pointwise in fields whose ghosts are valid)
__global__ void k_consistency(const __grid_constant__ DP P, long n, const double *sea, const double *bed,
                              const double *thk, double *mask_out, double *surf_out, unsigned *err) {
  const long q = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n) return;
  const double H = thk[q], s = sea ? sea[q] : 0.0, b = bed[q];
  if (H < 0.0) atomicOr(err, EB_NEG_THK); // check_minimum_ice_thickness, Geometry.cc:126
  const double hgrounded = __dadd_rn(b, H);
  const double hfloating = __dadd_rn(s, __dmul_rn(P.gc_alpha, H));
  const bool is_floating = (hfloating > hgrounded), ice_free = (H <= P.gc_icefree);
  double m, surf;
  if (is_floating && !P.gc_dry) {
    surf = hfloating;
    m = ice_free ? 4.0 : 3.0;
  } else {
    surf = hgrounded;
    m = ice_free ? 0.0 : 2.0;
  }
  mask_out[q] = m;
  surf_out[q] = surf;
}

// IceGrid::kBelowHeight (util/IceGrid.cc:427-440) by bisection on the levels in global memory
__device__ __forceinline__ int k_below_height_g(const double *z, int Mz, double height, unsigned *err) {
  if (height < 0.0 - 1.0e-6) {
    atomicOr(err, EB_BELOW);
    return 0;
  }
  if (height > z[Mz - 1] + 1.0e-6) {
    atomicOr(err, EB_ABOVE);
    return 0;
  }
  int ilo = 0, ihi = Mz - 1;
  while (ihi > ilo + 1) {
    const int m = (ihi + ilo) >> 1;
    if (z[m] > height) {
      ihi = m;
    } else {
      ilo = m;
    }
  }
  return ilo;
}

__device__ __forceinline__ double warp_max(double x) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) x = fmax(x, __shfl_xor_sync(FULLMASK, x, d));
  return x;
}

// One warp per column, lanes across z.  out[0..3] = bit patterns of max denom, max |u|, max |v|, max |w| (all >= 0,
// so the unsigned order is the double order).  min_k 1/denom_k = 1 / max_k denom_k exactly: correctly rounded
// division is monotone, so the host takes one reciprocal of the maximum (timestepping.cc:79-83).
__global__ void k_cfl_3d(const __grid_constant__ DP P, const double *__restrict__ thk, const double *__restrict__ mask,
                         const double *__restrict__ u, const double *__restrict__ v, const double *__restrict__ w,
                         const double *__restrict__ z, unsigned long long *out, unsigned *err) {
  const long col = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  double dmax = 0.0, umax = 0.0, vmax = 0.0, wmax = 0.0;
  if (col < (long)P.xm * P.ym) {
    const int i = P.xs + (int)(col % P.xm), j = P.ys + (int)(col / P.xm);
    const long g = idx2(P, i, j, P.wg);
    if (m_icy(mask_int(mask[g]))) {
      const int ks = k_below_height_g(z, P.Mz, thk[g], err);
      const double *uc = u + idx2(P, i, j, P.wuv) * P.Mz, *vc = v + idx2(P, i, j, P.wuv) * P.Mz;
      const double *wc = w + col * P.Mz;
      for (int k = lane; k <= ks; k += 32) {
        const double ua = fabs(uc[k]), va = fabs(vc[k]);
        umax = fmax(umax, ua);
        vmax = fmax(vmax, va);
        dmax = fmax(dmax, __dadd_rn(fabs(__dmul_rn(ua, P.inv_dx)), fabs(__dmul_rn(va, P.inv_dy))));
        wmax = fmax(wmax, fabs(wc[k]));
      }
    }
  }
  dmax = warp_max(dmax), umax = warp_max(umax), vmax = warp_max(vmax), wmax = warp_max(wmax);
  __shared__ double sm[4][8];
  const int wid = threadIdx.x >> 5;
  if (lane == 0) sm[0][wid] = dmax, sm[1][wid] = umax, sm[2][wid] = vmax, sm[3][wid] = wmax;
  __syncthreads();
  if (threadIdx.x < 4) {
    double m = 0.0;
    for (int q = 0; q < (int)(blockDim.x >> 5); ++q) m = fmax(m, sm[threadIdx.x][q]);
    if (m > 0.0) atomicMax(out + threadIdx.x, (unsigned long long)__double_as_longlong(m));
  }
}

// timestepping.cc:113-153.  out[4..6] = max denom, max |u|, max |v| of the 2D (sliding) velocity in icy cells.
// denom = u_abs / dx + v_abs / dy (true divisions here, unlike the 3D version).
__global__ void k_cfl_2d(const __grid_constant__ DP P, const double *__restrict__ mask, const double *__restrict__ vel,
                         unsigned long long *out) {
  const long q = (long)blockIdx.x * blockDim.x + threadIdx.x;
  double dmax = 0.0, umax = 0.0, vmax = 0.0;
  if (q < (long)P.xm * P.ym) {
    const int i = P.xs + (int)(q % P.xm), j = P.ys + (int)(q / P.xm);
    if (m_icy(mask_int(mask[idx2(P, i, j, P.wg)]))) {
      const long gv = idx2(P, i, j, P.wsl) * 2;
      umax = fabs(vel[gv]), vmax = fabs(vel[gv + 1]);
      dmax = __dadd_rn(__ddiv_rn(umax, P.dx), __ddiv_rn(vmax, P.dy));
    }
  }
  dmax = warp_max(dmax), umax = warp_max(umax), vmax = warp_max(vmax);
  if ((threadIdx.x & 31) == 0) {
    if (dmax > 0.0) atomicMax(out + 4, (unsigned long long)__double_as_longlong(dmax));
    if (umax > 0.0) atomicMax(out + 5, (unsigned long long)__double_as_longlong(umax));
    if (vmax > 0.0) atomicMax(out + 6, (unsigned long long)__double_as_longlong(vmax));
  }
}

} // namespace

int launch_mass_flow(const DP &P, double dt, const double *H, const double *bed, const double *sea, const double *vel,
                     const double *vel_bc, const double *thk_bc, const double *Q, double *flux_div, double *dH,
                     double *cons_err, cudaStream_t s) {
  const long n = (long)P.xm * P.ym;
  MassArgs A{H, bed, sea, vel, vel_bc, thk_bc, Q, flux_div, dH, cons_err, dt};
  k_mass_flow<<<nblk_(n, 256), 256, 0, s>>>(P, A);
  return 1;
}

int launch_mass_apply(const DP &P, double *H, const double *dH, cudaStream_t s) {
  k_mass_apply<<<nblk_((long)P.xm * P.ym, 256), 256, 0, s>>>(P, H, dH);
  return 1;
}

int launch_mass_source(const DP &P, double dt, double ice_density, int use_bmr, double *H, const double *mask,
                       const double *thk_bc, const double *smb, const double *bmr, double *eff_smb, double *eff_bmb,
                       cudaStream_t s) {
  k_mass_source<<<nblk_((long)P.xm * P.ym, 256), 256, 0, s>>>(P, dt, ice_density, use_bmr, H, mask, thk_bc, smb, bmr,
                                                               eff_smb, eff_bmb);
  return 1;
}

int launch_consistency(const DP &P, long n, const double *sea, const double *bed, const double *thk, double *mask_out,
                       double *surf_out, unsigned *err, cudaStream_t s) {
  k_consistency<<<nblk_(n, 256), 256, 0, s>>>(P, n, sea, bed, thk, mask_out, surf_out, err);
  return 1;
}

int launch_cfl(const DP &P, bool do3d, const double *thk, const double *mask, const double *u, const double *v,
               const double *w, const double *z, const double *vel, unsigned long long *out, unsigned *err,
               cudaStream_t s) {
  int n = 0;
  const long cols = (long)P.xm * P.ym;
  if (do3d) {
    k_cfl_3d<<<nblk_(cols * 32, 256), 256, 0, s>>>(P, thk, mask, u, v, w, z, out, err);
    n += 1;
  }
  if (vel) {
    k_cfl_2d<<<nblk_(cols, 256), 256, 0, s>>>(P, mask, vel, out);
    n += 1;
  }
  return n;
}

} // namespace siafd
