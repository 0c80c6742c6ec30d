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
#include "siafd_math.cuh"

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


// ---------------------------------------------------------------------------------------------
// StressBalance::compute_vertical_velocity (stressbalance/StressBalance.cc:283-424) fused with max_timestep_cfl_3d
// (stressbalance/timestepping.cc:42-101), which the reference evaluates right after it on the same fields
// (StressBalance.cc:186-200).  SURVEY.md 8(f) N2 + N3-CFL.
//
// One warp per column, lanes across z (NCH chunks of 32 levels held in registers), marching up the rows of a row
// segment: the v columns of rows j-1 and j stay in registers while row j+1 is loaded, so v is read once per row
// (plus two rows per segment) instead of three times; the three u columns of a row are adjacent in memory and the
// eight warps of a CTA take eight adjacent columns, so the neighbours' u comes from L1.  Per row a warp has 4 NCH
// independent loads in flight.  The running integral w[k] = w[k-1] - dz/2 (s[k] + s[k-1]) is a warp scan with a
// carry between chunks.  The CFL maxima (max denom, |u|, |v|, |w| over icy columns below the surface) are taken on
// the values already in registers: no second pass over u, v, w.
// HBM-bound: 24 Mz bytes per column (u, v read, w written).
// ---------------------------------------------------------------------------------------------
// inclusive warp scan step: t += (value of lane - d), only where that lane exists.  shfl.up returns the in-range
// predicate itself, so a step is two SHFL and one predicated DADD (no compare / select).
__device__ __forceinline__ double scan_step(double t, int d) {
  int lo = __double2loint(t), hi = __double2hiint(t);
  double r;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      ".reg .b32 ylo, yhi;\n"
      ".reg .f64 y, tt;\n"
      "shfl.sync.up.b32 ylo|p, %1, %3, 0, 0xffffffff;\n"
      "shfl.sync.up.b32 yhi, %2, %3, 0, 0xffffffff;\n"
      "mov.b64 y, {ylo, yhi};\n"
      "mov.b64 tt, {%1, %2};\n"
      "@p add.f64 tt, tt, y;\n"
      "mov.f64 %0, tt;\n"
      "}\n"
      : "=d"(r)
      : "r"(lo), "r"(hi), "r"(d));
  return r;
}

template <int NCH>
struct VvelRow { // everything a warp keeps while it marches up its column
  double hdz[NCH], zk[NCH]; // 0.5 (z[k] - z[k-1]) and z[k] of this lane's levels (k = 32 c + lane)
  int koff[NCH];            // min(k, Mz - 1): loads past the top re-read the top level instead of branching
  double ztop;              // z[Mz - 1]
};

// one row of one column; (vs, vc, vn) = v of rows j - 1, j, j + 1 at this lane's levels (vn is loaded here)
template <int NCH>
__device__ __forceinline__ void vvel_row(const DP &P, const VvelRow<NCH> &R, int lane, int Mz, int upstream, bool cfl,
                                         const double *__restrict__ mask, const double *__restrict__ thk,
                                         const double *__restrict__ bmr, long g, long o2, long rowuv,
                                         const double *__restrict__ uc_p, const double *__restrict__ vc_p,
                                         double *__restrict__ w_p, const double (&vs)[NCH], const double (&vc)[NCH],
                                         double (&vn)[NCH], int Ms, int M0, int Mn, double &dmax, double &umax,
                                         double &vmax, double &wmax, unsigned *err) {
  double uw[NCH], uc[NCH], ue[NCH];
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    const int k = R.koff[c];
    vn[c] = vc_p[k + rowuv];
    uw[c] = uc_p[k - Mz];
    uc[c] = uc_p[k];
    ue[c] = uc_p[k + Mz];
  }
  const int Me = mask_int(mask[g + 1]), Mw = mask_int(mask[g - 1]);
  double west = 1.0, east = 1.0, south = 1.0, north = 1.0;
  if (upstream) { // :336-350, :372-386 (basal velocities decide the direction)
    const double uw0 = __shfl_sync(FULLMASK, uw[0], 0), uc0 = __shfl_sync(FULLMASK, uc[0], 0),
                 ue0 = __shfl_sync(FULLMASK, ue[0], 0), vs0 = __shfl_sync(FULLMASK, vs[0], 0),
                 vc0 = __shfl_sync(FULLMASK, vc[0], 0), vn0 = __shfl_sync(FULLMASK, vn[0], 0);
    const double uwf = 0.5 * (uw0 + uc0), uef = 0.5 * (uc0 + ue0);
    if (uwf > 0.0 && uef >= 0.0) {
      west = 1.0, east = 0.0;
    } else if (uwf <= 0.0 && uef < 0.0) {
      west = 0.0, east = 1.0;
    }
    const double vsf = 0.5 * (vs0 + vc0), vnf = 0.5 * (vc0 + vn0);
    if (vsf > 0.0 && vnf >= 0.0) {
      south = 1.0, north = 0.0;
    } else if (vsf <= 0.0 && vnf < 0.0) {
      south = 0.0, north = 1.0;
    }
  }
  // one-sided differences at ice margins (:352-357, :388-393)
  const bool icy0 = m_icy(M0);
  if (icy0 != m_icy(Me)) east = 0;
  if (icy0 != m_icy(Mw)) west = 0;
  if (icy0 != m_icy(Mn)) north = 0;
  if (icy0 != m_icy(Ms)) south = 0;
  // 1 / (dx (east + west)) with east + west in {0, 1, 2}: RN(1 / dx) and exactly half of it, no division
  const double D_x = (east + west > 1.5) ? 0.5 * P.inv_dx : ((east + west > 0) ? P.inv_dx : 0.0);
  const double D_y = (north + south > 1.5) ? 0.5 * P.inv_dy : ((north + south > 0) ? P.inv_dy : 0.0);
  // fold the weights into four coefficients: u_x = a_e (u_e - u) + a_w (u - u_w), likewise v_y
  const double a_e = D_x * east, a_w = D_x * west, a_n = D_y * north, a_s = D_y * south;
  double wacc = (bmr != nullptr) ? -bmr[o2] : 0.0; // w at the base (:409-413)
  double carry = 0.0;                              // s at the level below the chunk
  double wk[NCH];
#pragma unroll
  for (int c = 0; c < NCH; ++c) {
    const double u_x = a_w * (uc[c] - uw[c]) + a_e * (ue[c] - uc[c]);
    const double v_y = a_s * (vc[c] - vs[c]) + a_n * (vn[c] - vc[c]);
    const double sk = u_x + v_y;
    double sprev = __shfl_up_sync(FULLMASK, sk, 1);
    if (lane == 0) sprev = carry;
    double t = -R.hdz[c] * (sk + sprev); // w[k] - w[k-1] (:418-422); hdz = 0 at k = 0 and past the top
    t = scan_step(t, 1);
    t = scan_step(t, 2);
    t = scan_step(t, 4);
    t = scan_step(t, 8);
    t = scan_step(t, 16);
    wk[c] = wacc + t;
    if (c * 32 + lane < Mz) w_p[c * 32 + lane] = wk[c];
    if (c + 1 < NCH) {
      wacc = __shfl_sync(FULLMASK, wk[c], 31);
      carry = __shfl_sync(FULLMASK, sk, 31);
    }
  }
  if (cfl && icy0) { // max_timestep_cfl_3d, timestepping.cc:62-87: icy columns, levels 0 .. ks (warp-uniform branch)
    // IceGrid::kBelowHeight (util/IceGrid.cc:427-440): the levels are sorted, so the largest k with z[k] <= H is a
    // count; clamped to [0, Mz - 2] like GSL's bsearch
    const double H = thk[g];
    if (H < 0.0 - 1.0e-6) {
      if (lane == 0) atomicOr(err, EB_BELOW);
    } else if (H > R.ztop + 1.0e-6) {
      if (lane == 0) atomicOr(err, EB_ABOVE);
    } else {
      int cnt = 0;
#pragma unroll
      for (int c = 0; c < NCH; ++c) cnt += __popc(__ballot_sync(FULLMASK, (c * 32 + lane < Mz) && R.zk[c] <= H));
      const int ks = min(max(cnt - 1, 0), Mz - 2);
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        if (c * 32 + lane <= ks) {
          const double ua = fabs(uc[c]), va = fabs(vc[c]);
          umax = fmax(umax, ua);
          vmax = fmax(vmax, va);
          dmax = fmax(dmax, __dadd_rn(fabs(__dmul_rn(ua, P.inv_dx)), fabs(__dmul_rn(va, P.inv_dy))));
          wmax = fmax(wmax, fabs(wk[c]));
        }
      }
    }
  }
}

template <int NCH>
__global__ void __launch_bounds__(256, (NCH <= 4) ? 2 : 1)
    k_vvel_march(const __grid_constant__ DP P, const double *__restrict__ mask, const double *__restrict__ thk,
                 const double *__restrict__ u, const double *__restrict__ v, const double *__restrict__ bmr,
                 int upstream, const double *__restrict__ z, double *__restrict__ w, int RS, unsigned long long *cfl,
                 unsigned *err) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int i = P.xs + blockIdx.x * 8 + wid;
  const int j0 = P.ys + blockIdx.y * RS, j1 = min(j0 + RS, P.ys + P.ym);
  double dmax = 0.0, umax = 0.0, vmax = 0.0, wmax = 0.0;
  if (i < P.xs + P.xm) {
    const int Mz = P.Mz;
    VvelRow<NCH> R;
    R.ztop = z[Mz - 1];
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      const int k = c * 32 + lane;
      R.koff[c] = min(k, Mz - 1);
      R.zk[c] = z[R.koff[c]];
      R.hdz[c] = (k > 0 && k < Mz) ? 0.5 * (z[k] - z[k - 1]) : 0.0;
    }
    const long rowuv = (long)(P.xm + 2 * P.wuv) * Mz, roww = (long)P.xm * Mz, rowg = P.xm + 2 * P.wg;
    const double *uc_p = u + idx2(P, i, j0, P.wuv) * Mz, *vc_p = v + idx2(P, i, j0, P.wuv) * Mz;
    long o2 = (long)(j0 - P.ys) * P.xm + (i - P.xs);
    double *w_p = w + o2 * Mz;
    double va[NCH], vb[NCH], vcc[NCH];
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      va[c] = vc_p[R.koff[c] - rowuv];
      vb[c] = vc_p[R.koff[c]];
    }
    long g = idx2(P, i, j0, P.wg);
    int Ms = mask_int(mask[g - rowg]), M0 = mask_int(mask[g]);
    const bool docfl = cfl != nullptr;
    // the three v rows rotate through (va, vb, vcc) by unrolling the row loop three times: no register moves
#define VVEL_ROW(S_, C_, N_)                                                                                           \
  {                                                                                                                    \
    const int Mn = mask_int(mask[g + rowg]);                                                                           \
    vvel_row<NCH>(P, R, lane, Mz, upstream, docfl, mask, thk, bmr, g, o2, rowuv, uc_p, vc_p, w_p, S_, C_, N_, Ms, M0,  \
                  Mn, dmax, umax, vmax, wmax, err);                                                                    \
    Ms = M0, M0 = Mn;                                                                                                  \
    uc_p += rowuv, vc_p += rowuv, w_p += roww, g += rowg, o2 += P.xm;                                                  \
  }
    int j = j0;
    while (j < j1) {
      VVEL_ROW(va, vb, vcc);
      if (++j >= j1) break;
      VVEL_ROW(vb, vcc, va);
      if (++j >= j1) break;
      VVEL_ROW(vcc, va, vb);
      ++j;
    }
#undef VVEL_ROW
  }
  if (cfl != nullptr) {
    dmax = warp_max(dmax), umax = warp_max(umax), vmax = warp_max(vmax), wmax = warp_max(wmax);
    __shared__ double sm[4][8];
    if (lane == 0) sm[0][wid] = dmax, sm[1][wid] = umax, sm[2][wid] = vmax, sm[3][wid] = wmax;
    __syncthreads();
    if (threadIdx.x < 4) {
      double m = 0.0;
#pragma unroll
      for (int q = 0; q < 8; ++q) m = fmax(m, sm[threadIdx.x][q]);
      // most CTAs find the running maximum already above theirs: a plain load spares the atomic
      const unsigned long long bits = (unsigned long long)__double_as_longlong(m);
      if (m > 0.0 && bits > *(volatile unsigned long long *)(cfl + threadIdx.x)) atomicMax(cfl + threadIdx.x, bits);
    }
  }
}


// ---------------------------------------------------------------------------------------------
// k_vvel_slab: the same computation (w + fused 3D CFL maxima) with the z sweep in registers.
//
// k_vvel_march keeps lanes across z; every access is coalesced but the running integral becomes warp scans and the
// kernel is issue-bound: 560 warp instructions per column, 45 % of the HBM roofline
// (profiles/ncu_r01_vvel_march_2048_summary.txt).  Here a CTA takes a strip of NC = 16 columns and marches up the
// rows of a row segment, transposing through shared memory like k_sia_slab:
//   * loads: per row one mbarrier-tracked group of two cp.async.bulk copies (the u row, NC + 2 columns, and the v row
//     two rows ahead, NC columns: both contiguous in PISM's layout), issued one row ahead by one thread into a
//     2-slot (u) / 4-slot (v) ring; the per-column scalars (5 masks, thickness, basal melt) of the next row are
//     prefetched into registers by the q = 0 threads during the sweep;
//   * sweep: thread = (column c, z range q of Lq = ceil(Mz / WZ) levels).  Shared-memory columns are Mz doubles
//     apart (odd), so the 16 columns of a half-warp fall into 16 different 8-byte banks.  The thread forms
//     s = u_x + v_y level by level and the running integral of its range (one add per level), leaves the range-local
//     w and the range total in shared memory, and takes the |u|, |v|, denominator maxima of the CFL criterion;
//   * fix-up and store: the thread adds w(0) and the totals of the ranges below to its levels (and takes max |w|
//     below the surface); the finished row (NC columns, contiguous in w) leaves by ONE bulk store from shared memory
//     (cp.async.bulk.global.shared::cta), double-buffered so that it overlaps the next row's sweep.
// Two barriers per row.  Shared memory 112 KB at Mz = 101: 2 CTAs per SM.  HBM-bound: 24 Mz bytes per column.
// Needs an odd Mz (bank layout) -- otherwise the caller falls back to k_vvel_march.
// ---------------------------------------------------------------------------------------------
#ifndef VVEL_NC
#define VVEL_NC 16 // columns per strip of k_vvel_slab
#endif
#ifndef VVEL_MINB
#define VVEL_MINB 2 // CTAs per SM its register allocation aims at
#endif
constexpr int VNC = VVEL_NC;

struct VvelArgs {
  const double *mask, *thk, *u, *v, *bmr, *z;
  double *w;
  unsigned long long *cfl;
  unsigned *err;
  int upstream, RS, WZ, Lq;
  long nUV; // doubles in the u / v arrays (bulk copies are clamped to it)
  double inv_dz;
};

__host__ __device__ inline int vvel_su(int Mz) { return ((VNC + 2) * Mz + 2 + 1) & ~1; }
__host__ __device__ inline int vvel_sv(int Mz) { return (VNC * Mz + 2 + 1) & ~1; }
__host__ __device__ inline int vvel_sw(int Mz) { return (VNC * Mz + 2 + 1) & ~1; }
__host__ inline size_t vvel_smem_bytes(int Mz) {
  const size_t d = 2 * (size_t)vvel_su(Mz) + 4 * (size_t)vvel_sv(Mz) + 2 * (size_t)vvel_sw(Mz) +
                   2 * (size_t)((Mz + 1) & ~1) + VNC * 16 + 4 * VNC;
  return d * 8 + 2 * 8 + (2 * VNC + 2 * VNC * 5) * 4 + 16;
}

// IceGrid::kBelowHeight (util/IceGrid.cc:427-440) on the levels in shared memory
__device__ __forceinline__ int k_below_height_s(const double *zz, int Mz, double height, double inv_dz, unsigned *err) {
  if (height < 0.0 - 1.0e-6) {
    atomicOr(err, EB_BELOW);
    return 0;
  }
  if (height > zz[Mz - 1] + 1.0e-6) {
    atomicOr(err, EB_ABOVE);
    return 0;
  }
  if (inv_dz > 0.0) { // equally spaced levels: guess, then correct against the table
    int k = min(max((int)(height * inv_dz), 0), Mz - 2);
    while (k < Mz - 2 && zz[k + 1] <= height) ++k;
    while (k > 0 && zz[k] > height) --k;
    return k;
  }
  int ilo = 0, ihi = Mz - 1;
  while (ihi > ilo + 1) {
    const int m = (ihi + ilo) >> 1;
    if (zz[m] > height) {
      ihi = m;
    } else {
      ilo = m;
    }
  }
  return ilo;
}

__global__ void __launch_bounds__(VNC * 16, VVEL_MINB) k_vvel_slab(const __grid_constant__ DP P, const VvelArgs A) {
  extern __shared__ __align__(16) double smem[];
  const int Mz = P.Mz, WZ = A.WZ, Lq = A.Lq, T = blockDim.x, tid = threadIdx.x;
  const int SU = vvel_su(Mz), SV = vvel_sv(Mz), SW = vvel_sw(Mz), MzE = (Mz + 1) & ~1;
  double *su = smem;                 // [2][SU]   u rows (slot = r & 1); data starts 0 or 1 double into the slot
  double *sv = su + 2 * SU;          // [4][SV]   v rows (slot = (rv + 1) & 3, rv = row - j0 = -1 ..)
  double *sw = sv + 4 * SV;          // [2][SW]   w of the row (slot = r & 1), same 0 / 1 shift as its place in w
  double *shz = sw + 2 * SW;         // [Mz]      0.5 (z[k] - z[k-1]), 0 at k = 0
  double *szz = shz + MzE;           // [Mz]      z[k]
  double *stot = szz + MzE;          // [16][VNC] range totals
  double *sthk = stot + VNC * 16;    // [2][VNC]  thickness (slot = r & 1)
  double *sbmr = sthk + 2 * VNC;     // [2][VNC]  basal melt rate
  unsigned long long *bars = (unsigned long long *)(sbmr + 2 * VNC); // [2]
  int *sks = (int *)(bars + 2);      // [2][VNC]  ks of icy columns, -1 otherwise
  int *smk = sks + 2 * VNC;          // [2][VNC][5] masks: centre, east, west, north, south

  const int i0 = P.xs + blockIdx.x * VNC, ncol = min(VNC, P.xs + P.xm - i0);
  const int j0 = P.ys + blockIdx.y * A.RS, nrows = min(A.RS, P.ys + P.ym - j0);
  const int c = tid % VNC, q = tid / VNC; // q < WZ
  // The chores of a row -- issuing the bulk copies and stores, fetching the next row's per-column scalars -- go to the
  // threads of the LAST z range: it holds the fewest levels (none when WZ Lq >= Mz + Lq), whereas the first range's warp
  // also takes the CFL maxima; with the chores on warp 0 the other warps waited for it at every barrier (31 % of the
  // stall samples, profiles/ncu_r02_k_vvel_slab_2048_summary.txt)
  const int q_io = WZ - 1;
  const bool is_io = tid == q_io * VNC;
  const long rowuv = (long)(P.xm + 2 * P.wuv) * Mz, rowg = P.xm + 2 * P.wg, roww = (long)P.xm * Mz;
  const long gu0 = idx2(P, i0 - 1, j0, P.wuv) * Mz, gv0 = idx2(P, i0, j0, P.wuv) * Mz;
  const long gw0 = ((long)(j0 - P.ys) * P.xm + (i0 - P.xs)) * Mz;
  const int nu = (ncol + 2) * Mz, nv = ncol * Mz;
  const bool docfl = A.cfl != nullptr;

  for (int k = tid; k < Mz; k += T) {
    shz[k] = (k > 0) ? 0.5 * (A.z[k] - A.z[k - 1]) : 0.0;
    szz[k] = A.z[k];
  }
  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    fence_mbar_init();
  }
  // per-column scalars of a row: masks (centre, east, west, north, south), thickness, basal melt
  auto load_scalars = [&](int j, int (&M)[5], double &H, double &B) {
    const long g = idx2(P, i0 + c, j, P.wg);
    M[0] = mask_int(A.mask[g]), M[1] = mask_int(A.mask[g + 1]), M[2] = mask_int(A.mask[g - 1]);
    M[3] = mask_int(A.mask[g + rowg]), M[4] = mask_int(A.mask[g - rowg]);
    H = A.thk[g];
    B = (A.bmr != nullptr) ? A.bmr[(long)(j - P.ys) * P.xm + (i0 + c - P.xs)] : 0.0;
  };
  auto store_scalars = [&](int slot, const int (&M)[5], double H, double B) {
    int *mk = smk + (slot * VNC + c) * 5;
#pragma unroll
    for (int e = 0; e < 5; ++e) mk[e] = M[e];
    sthk[slot * VNC + c] = H;
    sbmr[slot * VNC + c] = B;
    sks[slot * VNC + c] = (docfl && m_icy(M[0])) ? k_below_height_s(szz, Mz, H, A.inv_dz, A.err) : -1;
  };
  __syncthreads(); // tables and mbarrier init visible
  // one bulk copy of n doubles starting at base[goff] into a slot; the copy runs on 16-byte boundaries, so the data
  // lands (goff & 1) doubles into the slot
  auto copy_bytes = [&](long goff, int n) {
    const long a0 = goff & ~1L;
    long a1 = (goff + n + 1) & ~1L;
    if (a1 > A.nUV) a1 -= 2;
    return (unsigned)((a1 - a0) * 8);
  };
  auto copy_issue = [&](const double *base, long goff, int n, double *slot, unsigned long long *bar) {
    const long a0 = goff & ~1L;
    long a1 = (goff + n + 1) & ~1L;
    if (a1 > A.nUV) { // last row of the array and an odd end: fetch the last double separately
      a1 -= 2;
      slot[(goff - a0) + n - 1] = base[goff + n - 1];
    }
    bulk_g2s(slot, base + a0, (unsigned)((a1 - a0) * 8), bar);
  };
  if (is_io) { // rows j0 - 1, j0, j0 + 1 of v and row j0 of u: the group of row r = 0
    unsigned bytes = copy_bytes(gu0, nu);
    for (int rv = -1; rv <= 1; ++rv) bytes += copy_bytes(gv0 + rv * rowuv, nv);
    mbar_expect_tx(&bars[0], bytes);
    copy_issue(A.u, gu0, nu, su, &bars[0]);
    for (int rv = -1; rv <= 1; ++rv) copy_issue(A.v, gv0 + rv * rowuv, nv, sv + ((rv + 1) & 3) * SV, &bars[0]);
  }
  if (q == q_io && c < ncol) {
    int M[5];
    double H, B;
    load_scalars(j0, M, H, B);
    store_scalars(0, M, H, B);
  }
  __syncthreads();

  double dmax = 0.0, umax = 0.0, vmax = 0.0, wmax = 0.0;
  for (int r = 0; r < nrows; ++r) {
    const int j = j0 + r, sl = r & 1;
    const bool more = r + 1 < nrows;
    // ---- issue the loads of row r + 1: u row r + 1 and v row r + 2 -------------------------------------------------
    if (is_io && more) {
      const long gu = gu0 + (long)(r + 1) * rowuv, gv = gv0 + (long)(r + 2) * rowuv;
      unsigned long long *bar = &bars[(r + 1) & 1];
      mbar_expect_tx(bar, copy_bytes(gu, nu) + copy_bytes(gv, nv));
      copy_issue(A.u, gu, nu, su + ((r + 1) & 1) * SU, bar);
      copy_issue(A.v, gv, nv, sv + ((r + 3) & 3) * SV, bar);
    }
    int Mnext[5];
    double Hnext = 0.0, Bnext = 0.0;
    const bool pre = more && q == q_io && c < ncol;
    if (pre) load_scalars(j + 1, Mnext, Hnext, Bnext);
    mbar_wait(&bars[sl], (r >> 1) & 1);
    // ---- sweep: range-local running integral -----------------------------------------------------------------------
    const long gw = gw0 + (long)r * roww;             // where this row's NC columns start in w
    double *wc = sw + sl * SW + (gw & 1) + c * Mz;    // same parity in shared memory: aligned bulk store
    const int k0 = q * Lq, k1 = min(k0 + Lq, Mz);
    int ks = -1;
    if (c < ncol) {
      const int *mk = smk + (sl * VNC + c) * 5;
      const int M0 = mk[0], Me = mk[1], Mw = mk[2], Mn = mk[3], Ms = mk[4];
      const double *uw = su + sl * SU + ((gu0 + (long)r * rowuv) & 1) + c * Mz, *uc = uw + Mz, *ue = uc + Mz;
      const double *vs = sv + ((r + 0) & 3) * SV + ((gv0 + (long)(r - 1) * rowuv) & 1) + c * Mz;
      const double *vc = sv + ((r + 1) & 3) * SV + ((gv0 + (long)r * rowuv) & 1) + c * Mz;
      const double *vn = sv + ((r + 2) & 3) * SV + ((gv0 + (long)(r + 1) * rowuv) & 1) + c * Mz;
      double west = 1.0, east = 1.0, south = 1.0, north = 1.0;
      if (A.upstream) { // :336-350, :372-386 (basal velocities decide the direction)
        const double uwf = 0.5 * (uw[0] + uc[0]), uef = 0.5 * (uc[0] + ue[0]);
        if (uwf > 0.0 && uef >= 0.0) {
          west = 1.0, east = 0.0;
        } else if (uwf <= 0.0 && uef < 0.0) {
          west = 0.0, east = 1.0;
        }
        const double vsf = 0.5 * (vs[0] + vc[0]), vnf = 0.5 * (vc[0] + vn[0]);
        if (vsf > 0.0 && vnf >= 0.0) {
          south = 1.0, north = 0.0;
        } else if (vsf <= 0.0 && vnf < 0.0) {
          south = 0.0, north = 1.0;
        }
      }
      // one-sided differences at ice margins (:352-357, :388-393)
      const bool icy0 = m_icy(M0);
      if (icy0 != m_icy(Me)) east = 0;
      if (icy0 != m_icy(Mw)) west = 0;
      if (icy0 != m_icy(Mn)) north = 0;
      if (icy0 != m_icy(Ms)) south = 0;
      // 1 / (dx (east + west)) with east + west in {0, 1, 2}: RN(1 / dx) and exactly half of it, no division
      const double D_x = (east + west > 1.5) ? 0.5 * P.inv_dx : ((east + west > 0) ? P.inv_dx : 0.0);
      const double D_y = (north + south > 1.5) ? 0.5 * P.inv_dy : ((north + south > 0) ? P.inv_dy : 0.0);
      const double a_e = D_x * east, a_w = D_x * west, a_n = D_y * north, a_s = D_y * south;
      ks = sks[sl * VNC + c];
      auto s_at = [&](int k) {
        const double u_x = a_w * (uc[k] - uw[k]) + a_e * (ue[k] - uc[k]);
        const double v_y = a_s * (vc[k] - vs[k]) + a_n * (vn[k] - vc[k]);
        return u_x + v_y;
      };
      double sprev = (k0 > 0 && k0 < Mz) ? s_at(k0 - 1) : 0.0, run = 0.0;
#pragma unroll 2
      for (int k = k0; k < k1; ++k) {
        const double sk = s_at(k);
        run -= shz[k] * (sk + sprev); // :418-422
        sprev = sk;
        wc[k] = run;
      }
      if (ks >= k0) { // timestepping.cc:68-83, levels k0 .. min(ks, k1 - 1) of this range
        const int ke = min(ks + 1, k1);
        for (int k = k0; k < ke; ++k) {
          const double ua = fabs(uc[k]), va = fabs(vc[k]);
          umax = fmax(umax, ua);
          vmax = fmax(vmax, va);
          dmax = fmax(dmax, __dadd_rn(fabs(__dmul_rn(ua, P.inv_dx)), fabs(__dmul_rn(va, P.inv_dy))));
        }
      }
      stot[q * VNC + c] = run;
    }
    if (pre) store_scalars(sl ^ 1, Mnext, Hnext, Bnext);
    __syncthreads();
    // ---- add w(0) and the totals of the ranges below; the row then leaves by one bulk store ---------------------------
    if (c < ncol) {
      double off = -sbmr[sl * VNC + c]; // :409-413 (0 without a basal melt rate)
      for (int e = 0; e < q; ++e) off += stot[e * VNC + c];
      const int ke = min(ks + 1, k1);
      for (int k = k0; k < k1; ++k) {
        const double wk = off + wc[k];
        wc[k] = wk;
        if (k < ke) wmax = fmax(wmax, fabs(wk));
      }
    }
    fence_proxy_async();                // this thread's shared-memory writes, for the bulk-copy engine
    if (is_io) bulk_wait_read0();       // the store of row r - 1 has read its slot: the next sweep may overwrite it
    __syncthreads();
    if (is_io) {
      const int n = ncol * Mz;
      const long a0 = (gw + 1) & ~1L, a1 = (gw + n) & ~1L; // the 16-byte aligned middle
      const double *src = sw + sl * SW + (gw & 1);
      if (gw & 1) A.w[gw] = src[0];
      if ((gw + n) & 1) A.w[gw + n - 1] = src[n - 1];
      if (a1 > a0) {
        bulk_s2g(A.w + a0, src + (a0 - gw), (unsigned)((a1 - a0) * 8));
        bulk_commit();
      }
    }
  }
  if (is_io) bulk_wait_read0();
  if (docfl) {
    dmax = warp_max(dmax), umax = warp_max(umax), vmax = warp_max(vmax), wmax = warp_max(wmax);
    __syncthreads();
    double *red = stot;
    const int lane = tid & 31, wid = tid >> 5, nw = T >> 5;
    if (lane == 0) red[wid] = dmax, red[8 + wid] = umax, red[16 + wid] = vmax, red[24 + wid] = wmax;
    __syncthreads();
    if (tid < 4) {
      double m = 0.0;
      for (int e = 0; e < nw; ++e) m = fmax(m, red[tid * 8 + e]);
      // most CTAs find the running maximum already above theirs: a plain load spares the atomic
      const unsigned long long bits = (unsigned long long)__double_as_longlong(m);
      if (m > 0.0 && bits > *(volatile unsigned long long *)(A.cfl + tid)) atomicMax(A.cfl + tid, bits);
    }
  }
}


// ---------------------------------------------------------------------------------------------
// StressBalance::compute_volumetric_strain_heating (stressbalance/StressBalance.cc:426-642), SURVEY.md 8(f) N3.
// Sigma = 2 e^(-1/n) B(E, p) D2^((1/n + 1)/2) on levels 0 .. ks, zero above; B = softness^(-1/n) (FlowLaw.cc:142-144)
// of the SHALLOW stress balance's flow law (law id, n, e are arguments, not the handle's SIA law).
// A half-warp per column, lanes across z, the eight warps of a CTA on sixteen adjacent columns, marching up the rows
// of a row segment: every level is independent, the stencil neighbours (x: adjacent in memory; y: the rows the walk has
// just read / reads next; z: levels k +- 1) come from L1 / L2.  For n = 3 the two powers are cube roots (x^(-1/3) = 1 / cbrt x,
// x^(2/3) = cbrt(x)^2: two ~40-instruction calls instead of two ~150-instruction pow); other n use pow.
// FP64-bound where there is ice (one exp, two cbrt per level), write-bound (8 Mz bytes) where there is none.
// ---------------------------------------------------------------------------------------------
template <int LAW> __device__ __forceinline__ double softness_eval(const DP &P, double E, double p) {
  // literal restatements of softness_impl: the pressure-adjusted temperature is T - T_m + T_melting
  // (util/EnthalpyConverter.cc:196-198), not the T + beta p of flow_from_temp
  if (LAW == LAW_ISO) {
    return P.iso_A; // rheology/IsothermalGlen.cc:41-43
  }
  const double T_m = ec_melting_temperature(P, p);
  const double T_pa = ec_temperature(P, E, p) - T_m + P.T_melting;
  if (LAW == LAW_PB) return softness_paterson_budd(P, T_pa);           // PatersonBudd.cc:41-45
  if (LAW == LAW_ARR) return P.A_cold * exp(-P.Q_cold / (P.R * T_pa)); // PatersonBuddCold.cc:43-45
  if (LAW == LAW_ARRWARM) return P.A_warm * exp(-P.Q_warm / (P.R * T_pa));
  if (LAW == LAW_HOOKE) return P.hk_A * exp(-P.hk_Q / (P.R * T_pa) + 3.0 * P.hk_C * pow(P.hk_Tr - T_pa, -P.hk_K));
  if (LAW == LAW_GPBLD) { // GPBLD.cc:49-61
    const double E_s = P.c_i * (T_m - P.T_0);
    if (E < E_s) return softness_paterson_budd(P, T_pa);
    double omega = (E - E_s) / ec_L(P, T_m);
    omega = fmin(omega, P.gp_limit);
    return P.gp_softness_T0 * (1.0 + P.gp_coeff * omega);
  }
  return __longlong_as_double(0x7ff8000000000000LL); // gk: the reference throws (GoldsbyKohlstedt.cc:102-108)
}

__device__ __forceinline__ void prefetch_l1(const void *p) { asm volatile("prefetch.global.L1 [%0];\n" ::"l"(p)); }

constexpr int HEAT_MAX_MZ = 512; // levels the strain-heating kernel's table of 1 / dz holds
struct HeatArgs {
  const double *mask, *thk, *E, *u, *v, *z;
  double *sigma;
  unsigned *err;
  double n, two_e_pow;         // Glen exponent; 2 e^(-1/n)
  double inv_n, iso_hardness;  // 1 / n; iso_A^(-1/n)
  int RS;
};

// hardness B = softness^(-1/n) (FlowLaw.cc:142-144).  Where the softness is an Arrhenius factor A exp(-Q / (R T)) the
// power is taken inside the exponential, B = exp((Q / (R T) - ln A) / n): one exp instead of an exp and a pow
// (relative difference to pow(softness, -1/n): a few 1e-15).  The temperate branch of gpbld and hooke take the
// literal softness and one cbrt (n = 3) or pow.
template <int LAW>
__device__ __forceinline__ double hardness_eval(const DP &P, double E, double p, double n, double inv_n, double iso_hardness,
                                                const double *tab16) {
  if (LAW == LAW_ISO) {
    return iso_hardness;
  }
  if (LAW == LAW_PB || LAW == LAW_ARR || LAW == LAW_ARRWARM || LAW == LAW_GPBLD) {
    const double T_m = ec_melting_temperature(P, p);
    if (LAW != LAW_GPBLD || E < P.c_i * (T_m - P.T_0)) {
      const double T_pa = ec_temperature(P, E, p) - T_m + P.T_melting; // EnthalpyConverter.cc:196-198
      const bool cold = (LAW == LAW_ARR) || (LAW != LAW_ARRWARM && T_pa < P.T_crit);
      // (ln A and Q / R in units of ln2 / 16, the table-based exp2 and the cubic reciprocal of the fused kernel,
      // siafd_math.cuh: 14 FP64 operations instead of 23; both within ~2 ulp)
      const double lnA2 = cold ? P.lnA2_cold : P.lnA2_warm, QoR2 = cold ? P.QoR2_cold : P.QoR2_warm;
      return exp2_tab16((QoR2 * rcp_cubic(T_pa) - lnA2) * inv_n, tab16);
    }
  }
  const double soft = softness_eval<LAW>(P, E, p);
  return (n == 3.0) ? 1.0 / cbrt(soft) : pow(soft, -inv_n);
}

template <int LAW>
__global__ void __launch_bounds__(256, 4) k_strain_heating(const __grid_constant__ DP P, const HeatArgs A) {
  // a half-warp per column (16 lanes across z: 128-byte pieces; Mz = 101 fills 7 x 16 lanes to 90 %, 4 x 32 to 79 %,
  // and the part of a column above the ice wastes at most 15 lanes), two adjacent columns per warp
  const int lane = threadIdx.x & 15, half = (threadIdx.x >> 4) & 1, wid = threadIdx.x >> 5;
  const int i_raw = P.xs + blockIdx.x * 16 + wid * 2 + half;
  const bool live = i_raw < P.xs + P.xm; // (an idle half still takes part in the warp's ballots)
  const int i = live ? i_raw : P.xs + P.xm - 1;
  const int j0 = P.ys + blockIdx.y * A.RS, j1 = min(j0 + A.RS, P.ys + P.ym);
  const int Mz = P.Mz, nch = (Mz + 15) >> 4;
  const bool n3 = (A.n == 3.0);
  const double exponent = 0.5 * (1.0 / A.n + 1.0);
  const double *__restrict__ z = A.z;
  const double ztop = z[Mz - 1];
  // 1 / (z[k+1] - z[max(k-1, 0)]) of the centred / one-sided vertical differences (StressBalance.cc:593-603), once per
  // CTA instead of two IEEE divisions (~50 FP64 instructions: 14.5 -> 12.3 ms at 4096^2 with the table-based exp2 below) per level and column; u_z then differs from the quotient
  // by <= 1 ulp
  __shared__ double s_idz[HEAT_MAX_MZ];
  __shared__ double s_tab16[16]; // 2^(j/16), for exp2_tab16
  for (int k = threadIdx.x; k < Mz - 1; k += blockDim.x) s_idz[k] = 1.0 / (z[k + 1] - z[max(k - 1, 0)]);
  if (threadIdx.x < 16) s_tab16[threadIdx.x] = EXPT[threadIdx.x];
  __syncthreads();
  const long rowuv = (long)(P.xm + 2 * P.wuv) * Mz, rowe = (long)(P.xm + 2 * P.we) * Mz, rowg = P.xm + 2 * P.wg;
  const double *uc_p = A.u + idx2(P, i, j0, P.wuv) * Mz, *vc_p = A.v + idx2(P, i, j0, P.wuv) * Mz;
  const double *e_p = A.E + idx2(P, i, j0, P.we) * Mz;
  double *s_p = A.sigma + ((long)(j0 - P.ys) * P.xm + (i - P.xs)) * Mz;
  long g = idx2(P, i, j0, P.wg);
  for (int j = j0; j < j1; ++j) {
    // the rows the next iterations miss in cache (u, v two rows up; enthalpy one row up) are requested now, so that
    // their DRAM latency overlaps this row's arithmetic -- but only the levels that will be read: those at or below
    // the ice surface of that row's column, plus one (u_z, v_z look one level up)
    {
      const bool uv_ok = j + 2 < P.ys + P.ym + P.wuv, e_ok = j + 1 < j1; // rows inside the arrays
      const double H1 = e_ok ? A.thk[g + rowg] : -1.0, H2 = (uv_ok && j + 2 <= P.ys + P.ym + P.wg - 1) ? A.thk[g + 2 * rowg] : -1.0;
      for (int c = 0; c < nch; ++c) {
        const int k = min(c * 16 + lane, Mz - 1);
        const double zb = z[max(k - 1, 0)];
        if (H2 > 0.0 && zb <= H2) {
          prefetch_l1(uc_p + k + 2 * rowuv);
          prefetch_l1(vc_p + k + 2 * rowuv);
        }
        if (H1 > 0.0 && zb <= H1) prefetch_l1(e_p + k + rowe);
      }
    }
    const double H = A.thk[g];
    // IceGrid::kBelowHeight (util/IceGrid.cc:427-440), for EVERY column, icy or not (StressBalance.cc:540): the levels
    // are sorted, so the largest k with z[k] <= H is a count; clamped to [0, Mz - 2] like GSL's bsearch
    int ks = 0;
    {
      int cnt = 0; // (the ballots are taken by the whole warp: the two halves hold different columns)
      for (int c = 0; c < nch; ++c) {
        const int k = c * 16 + lane;
        const unsigned b = __ballot_sync(FULLMASK, (k < Mz) && z[min(k, Mz - 1)] <= H);
        cnt += __popc((b >> (half * 16)) & 0xffffu);
      }
      if (H < 0.0 - 1.0e-6) {
        if (lane == 0) atomicOr(A.err, EB_BELOW);
      } else if (H > ztop + 1.0e-6) {
        if (lane == 0) atomicOr(A.err, EB_ABOVE);
      } else {
        ks = min(max(cnt - 1, 0), Mz - 2);
      }
    }
    const int M0 = mask_int(A.mask[g]), Me = mask_int(A.mask[g + 1]), Mw = mask_int(A.mask[g - 1]),
              Mn = mask_int(A.mask[g + rowg]), Ms = mask_int(A.mask[g - rowg]);
    const bool icy0 = m_icy(M0);
    double west = 1.0, east = 1.0, south = 1.0, north = 1.0;
    if (icy0 != m_icy(Me)) east = 0;
    if (icy0 != m_icy(Mw)) west = 0;
    if (icy0 != m_icy(Mn)) north = 0;
    if (icy0 != m_icy(Ms)) south = 0;
    // 1 / (dx (east + west)) with east + west in {0, 1, 2}: RN(1 / dx) and exactly half of it, no division
    const double D_x = (east + west > 1.5) ? 0.5 * P.inv_dx : ((east + west > 0) ? P.inv_dx : 0.0);
    const double D_y = (north + south > 1.5) ? 0.5 * P.inv_dy : ((north + south > 0) ? P.inv_dy : 0.0);
    const double a_e = D_x * east, a_w = D_x * west, a_n = D_y * north, a_s = D_y * south;
#pragma unroll 2
    for (int c = 0; c < nch; ++c) {
      const int k = c * 16 + lane;
      double sig = 0.0;
      if (k <= ks) { // (chunks wholly above the ice skip the arithmetic warp-uniformly)
        const double uc = uc_p[k], vc = vc_p[k];
        const double u_x = a_w * (uc - uc_p[k - Mz]) + a_e * (uc_p[k + Mz] - uc);
        const double v_x = a_w * (vc - vc_p[k - Mz]) + a_e * (vc_p[k + Mz] - vc);
        const double u_y = a_s * (uc - uc_p[k - rowuv]) + a_n * (uc_p[k + rowuv] - uc);
        const double v_y = a_s * (vc - vc_p[k - rowuv]) + a_n * (vc_p[k + rowuv] - vc);
        const int kp = k + 1, km = max(k - 1, 0); // k <= ks <= Mz - 2: level k + 1 exists; one-sided at the base
        const double idz = s_idz[k];
        const double u_z = (uc_p[kp] - uc_p[km]) * idz, v_z = (vc_p[kp] - vc_p[km]) * idz;
        const double d2 = 0.5 * ((u_x + v_y) * (u_x + v_y) + u_x * u_x + v_y * v_y +
                                 0.5 * ((u_y + v_x) * (u_y + v_x) + u_z * u_z + v_z * v_z));
        const double pr = P.p_air + P.rg * (H - z[k]); // EnthalpyConverter.cc:137-152 (no depth clamp)
        const double hard = hardness_eval<LAW>(P, e_p[k], pr, A.n, A.inv_n, A.iso_hardness, s_tab16);
        double dpow;
        if (n3) {
          const double cr = cbrt(d2); // (a single-precision seed + two Newton steps instead: no faster, 12.3 ms either way)
          dpow = cr * cr;
        } else {
          dpow = pow(d2, exponent);
        }
        sig = A.two_e_pow * hard * dpow;
      }
      if (live && k < Mz) s_p[k] = sig;
    }
    uc_p += rowuv, vc_p += rowuv, e_p += rowe, s_p += (long)P.xm * Mz, g += rowg;
  }
}


// SIAFD_Regional::compute_surface_gradient, the override loop (regional/SIAFD_Regional.cc:63-116), SURVEY.md 8(f) N4:
// next to no_model cells the gradient of the stored surface (haseloff, computed beforehand into hx_nm / hy_nm with
// valid ghosts) replaces the regular one; zero where the stencil leaves the domain.  On owned + 1.
__global__ void k_regional_override(const __grid_constant__ DP P, const double *__restrict__ no_model,
                                    const double *__restrict__ hx_nm, const double *__restrict__ hy_nm,
                                    double *__restrict__ h_x, double *__restrict__ h_y) {
  const long q = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const int nx = P.xm + 2;
  if (q >= (long)nx * (P.ym + 2)) return;
  const int i = P.xs - 1 + (int)(q % nx), j = P.ys - 1 + (int)(q / nx);
  auto NM = [&](int ii, int jj) { return mask_int(no_model[idx2(P, ii, jj, P.wg)]) >= 1; };
  const bool ij = NM(i, j), n = NM(i, j + 1), nw = NM(i - 1, j + 1), w = NM(i - 1, j), s = NM(i, j - 1),
             se = NM(i + 1, j - 1), e = NM(i + 1, j), ne = NM(i + 1, j + 1);
  const long o = idx2(P, i, j, P.wst) * 2;
  const int Mx = P.Mx, My = P.My;
  if (ij || e) h_x[o] = (i < 0 || i + 1 > Mx - 1) ? 0.0 : hx_nm[o];
  if (nw || ne || w || e) h_x[o + 1] = (i - 1 < 0 || j + 1 > My - 1 || i + 1 > Mx - 1) ? 0.0 : hx_nm[o + 1];
  if (n || ne || s || se) h_y[o] = (i < 0 || j + 1 > My - 1 || i + 1 > Mx - 1 || j - 1 < 0) ? 0.0 : hy_nm[o];
  if (ij || n) h_y[o + 1] = (j < 0 || j + 1 > My - 1) ? 0.0 : hy_nm[o + 1];
}

// IceModelVec3D::getValZ (util/iceModelVec3.cc:153-182) for every owned column: the value of a 3D field at height
// zq[i,j] (IceModelVec3::getSurfaceValues, :226-240, zq = ice thickness) or at one height z0 (getHorSlice, :209-223).
// Linear interpolation between the two levels around the height, the end levels outside [z_0, z_Mz-1]; expression
// order and roundings of the reference (no contraction), so the result is bit-identical to the CPU's.  One thread
// per column: two 8-byte reads of a column that is 8 Mz bytes long, so the field is touched at 64 of its 8 Mz
// bytes per column (sector granularity) and the launch is bound by those sectors, not by the 3D field's size.
__global__ void k_value_at_height(const __grid_constant__ DP P, const double *__restrict__ a, int wa,
                                  const double *__restrict__ zq, int wz, double z0, const double *__restrict__ z,
                                  double *__restrict__ out) {
  const long q = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= (long)P.xm * P.ym) return;
  const int i = P.xs + (int)(q % P.xm), j = P.ys + (int)(q / P.xm);
  const double height = zq ? zq[idx2(P, i, j, wz)] : z0;
  const double *col = a + idx2(P, i, j, wa) * P.Mz;
  const int Mz = P.Mz;
  double r;
  if (height >= z[Mz - 1]) {
    r = col[Mz - 1];
  } else if (height <= z[0]) {
    r = col[0];
  } else {
    int ilo = 0, ihi = Mz - 1; // gsl_interp_accel_find: largest m in [0, Mz - 2] with z[m] <= height
    while (ihi > ilo + 1) {
      const int m = (ihi + ilo) >> 1;
      if (z[m] > height) {
        ihi = m;
      } else {
        ilo = m;
      }
    }
    const double incr = __ddiv_rn(__dsub_rn(height, z[ilo]), __dsub_rn(z[ilo + 1], z[ilo]));
    const double valm = col[ilo];
    r = __dadd_rn(valm, __dmul_rn(incr, __dsub_rn(col[ilo + 1], valm)));
  }
  out[q] = r;
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

namespace siafd {
// the shared-memory kernel; returns 0 when it does not apply (even Mz, or rows too long for shared memory)
int launch_vvel_slab(const DP &P, const double *mask, const double *thk, const double *u, const double *v,
                     const double *bmr, int upstream, const double *z, double *w, unsigned long long *cfl,
                     unsigned *err, int rows_per_cta, int wz, long nUV, double inv_dz, cudaStream_t s) {
  if (P.xm <= 0 || P.ym <= 0) return 0;
  const int WZ = (wz == 2 || wz == 4 || wz == 8 || wz == 16) ? wz : 16;
  const size_t smem = vvel_smem_bytes(P.Mz);
  if ((P.Mz & 1) == 0 || smem > 113 * 1024 || P.Mz < 3 || P.wuv < 1) return 0;
  // (set on every launch: the attribute is per device and a process may hold handles on several)
  if (cudaFuncSetAttribute(k_vvel_slab, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  const int RS = rows_per_cta > 0 ? rows_per_cta : 32;
  VvelArgs A{mask, thk, u, v, bmr, z, w, cfl, err, upstream, RS, WZ, (P.Mz + WZ - 1) / WZ, nUV, inv_dz};
  const dim3 grid((unsigned)((P.xm + VNC - 1) / VNC), (unsigned)((P.ym + RS - 1) / RS));
  k_vvel_slab<<<grid, VNC * WZ, smem, s>>>(P, A);
  return 1;
}

// returns the number of launches, or 0 when Mz has more than 8 chunks of 32 levels (the caller then uses the
// generic kernel of siafd_kernels.cu and the stand-alone CFL kernel)
int launch_vvel_march(const DP &P, const double *mask, const double *thk, const double *u, const double *v,
                      const double *bmr, int upstream, const double *z, double *w, unsigned long long *cfl,
                      unsigned *err, int rows_per_cta, cudaStream_t s) {
  if (P.xm <= 0 || P.ym <= 0) return 0;
  const int nch = (P.Mz + 31) / 32;
  const int RS = rows_per_cta > 0 ? rows_per_cta : 32;
  const dim3 grid((unsigned)((P.xm + 7) / 8), (unsigned)((P.ym + RS - 1) / RS));
  if (nch <= 1) {
    k_vvel_march<1><<<grid, 256, 0, s>>>(P, mask, thk, u, v, bmr, upstream, z, w, RS, cfl, err);
  } else if (nch <= 2) {
    k_vvel_march<2><<<grid, 256, 0, s>>>(P, mask, thk, u, v, bmr, upstream, z, w, RS, cfl, err);
  } else if (nch <= 4) {
    k_vvel_march<4><<<grid, 256, 0, s>>>(P, mask, thk, u, v, bmr, upstream, z, w, RS, cfl, err);
  } else if (nch <= 8) {
    k_vvel_march<8><<<grid, 256, 0, s>>>(P, mask, thk, u, v, bmr, upstream, z, w, RS, cfl, err);
  } else {
    return 0;
  }
  return 1;
}
} // namespace siafd

namespace siafd {
template <int LAW>
static int launch_heat_law(const DP &P, const HeatArgs &A, cudaStream_t s) {
  const dim3 grid((unsigned)((P.xm + 15) / 16), (unsigned)((P.ym + A.RS - 1) / A.RS));
  k_strain_heating<LAW><<<grid, 256, 0, s>>>(P, A);
  return 1;
}

// returns launches, or -1: law not supported here (gk has no softness, GoldsbyKohlstedt.cc:102-108)
int launch_strain_heating(const DP &P, int law, double n, double e, const double *mask, const double *thk,
                          const double *E, const double *u, const double *v, const double *z, double *sigma,
                          unsigned *err, cudaStream_t s) {
  if (P.xm <= 0 || P.ym <= 0) return 0;
  if (P.Mz > HEAT_MAX_MZ) return -1;
  HeatArgs A{mask, thk, E, u, v, z, sigma, err, n, 2.0 * pow(e, -1.0 / n), 1.0 / n, pow(P.iso_A, -1.0 / n), 32};
  switch (law) {
  case LAW_ISO:
    return launch_heat_law<LAW_ISO>(P, A, s);
  case LAW_PB:
    return launch_heat_law<LAW_PB>(P, A, s);
  case LAW_GPBLD:
    return launch_heat_law<LAW_GPBLD>(P, A, s);
  case LAW_HOOKE:
    return launch_heat_law<LAW_HOOKE>(P, A, s);
  case LAW_ARR:
    return launch_heat_law<LAW_ARR>(P, A, s);
  case LAW_ARRWARM:
    return launch_heat_law<LAW_ARRWARM>(P, A, s);
  default:
    return -1;
  }
}
} // namespace siafd

namespace siafd {
int launch_regional_override(const DP &P, const double *no_model, const double *hx_nm, const double *hy_nm, double *h_x,
                             double *h_y, cudaStream_t s) {
  const long n = (long)(P.xm + 2) * (P.ym + 2);
  k_regional_override<<<nblk_(n, 256), 256, 0, s>>>(P, no_model, hx_nm, hy_nm, h_x, h_y);
  return 1;
}

int launch_value_at_height(const DP &P, const double *a, int wa, const double *zq, int wz, double z0, const double *z,
                           double *out, cudaStream_t s) {
  k_value_at_height<<<nblk_((long)P.xm * P.ym, 256), 256, 0, s>>>(P, a, wa, zq, wz, z0, z, out);
  return 1;
}
} // namespace siafd
