// siafd_kernels.cu -- sm_100a kernels of the SIAFD hot path.
//
// Reference path: stressbalance::SIAFD::update() of juliusgarbe/pism v1.2.1
// (src/stressbalance/sia/SIAFD.cc:122-155).  Kernel map:
//
//   k_prep2d        BedSmoother::smoothed_thk + ::theta      sia/BedSmoother.cc:284-327, :351-404
//   k_eta           eta = H^((2n+2)/n)                        sia/SIAFD.cc:241-245
//   k_grad_*        surface_gradient_{mahaffy,eta,haseloff}   sia/SIAFD.cc:224-496
//   k_sia_fused     compute_diffusivity + compute_diffusive_flux + compute_I +
//                   compute_3d_horizontal_velocity, fused     sia/SIAFD.cc:543-948
//   k_copy_region   ghost wrap / halo pack (DMLocalToLocal)   util/iceModelVec.cc:630-643
//   k_geometry      GeometryCalculator::compute               util/Mask.hh:96-133
//   k_flow_n        FlowLaw::flow_n                           rheology/FlowLaw.cc:107-119
//   k_bed_*         BedSmoother::preprocess_bed               sia/BedSmoother.cc:157-267
//
// The fused kernel never writes delta or I to HBM (the reference round-trips four 3D scratch
// fields): a CTA marches over rows of a 16-column strip, keeps the enthalpy rows it needs in
// shared memory (async copies two rows ahead), integrates every staggered column with 16
// lanes across z (half-warp prefix sums) and hands I(z) to the velocity stage through shared
// memory.  See DESIGN.md.
#include "siafd_kernels.cuh"

#include <cstdio>

namespace siafd {

#define FULLMASK 0xffffffffu

// ---------------------------------------------------------------------------------------------
// small PTX helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void cp_async8(void *smem_dst, const void *gmem_src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }

// mbarrier + 1-D bulk copy (TMA engine, no tensor map: rows are contiguous runs of doubles)
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
  asm volatile("{\n"
               ".reg .pred p;\n"
               "WAIT_LOOP:\n"
               "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
               "@p bra.uni WAIT_DONE;\n"
               "bra.uni WAIT_LOOP;\n"
               "WAIT_DONE:\n"
               "}\n" ::"r"(smem_u32(bar)),
               "r"(parity)
               : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gmem_src, unsigned bytes, unsigned long long *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }

// ---------------------------------------------------------------------------------------------
// k_prep2d: thk_smooth and theta on owned + wg ghosts (no communication, like the reference)
// ---------------------------------------------------------------------------------------------
__global__ void k_prep2d(const __grid_constant__ DP P, const Fields F) {
  const long n = (long)(P.xm + 2 * P.wg) * (P.ym + 2 * P.wg);
  const long q = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n) {
    return;
  }
  // BedSmoother::smoothed_thk, sia/BedSmoother.cc:300-322
  const double thk = F.H[q];
  double ts;
  if (thk < 0.0) {
    atomicOr(F.err, EB_NEG_THK);
    ts = 0.0;
  } else if (thk == 0.0) {
    ts = 0.0;
  } else if (F.maxtl[q] >= thk) {
    ts = thk;
  } else if (m_grounded(mask_int(F.mask[q]))) {
    const double thks_try = F.h[q] - F.topgsmooth[q];
    ts = (thks_try > 0.0) ? thks_try : 0.0;
  } else {
    ts = thk;
  }
  F.thk_smooth[q] = ts;

  // BedSmoother::theta, sia/BedSmoother.cc:353-397
  double th;
  if (!P.smoother_active) {
    th = 1.0;
  } else {
    const double H = F.h[q] - F.topgsmooth[q];
    if (H > F.maxtl[q]) {
      const double Hinv = 1.0 / fmax(H, 1.0);
      // explicit rn ops: no FMA contraction, so omega rounds exactly like the reference expression
      double omega = __dadd_rn(
          1.0, __dmul_rn(__dmul_rn(Hinv, Hinv),
                         __dadd_rn(F.C2[q], __dmul_rn(Hinv, __dadd_rn(F.C3[q], __dmul_rn(Hinv, F.C4[q]))))));
      if (omega <= 0) {
        atomicOr(F.err, EB_OMEGA);
      }
      if (omega < 0.001) {
        omega = 0.001;
      }
      th = pow(omega, -P.n);
    } else {
      th = 0.0;
    }
    th = fmin(fmax(P.theta_min, th), 1.0); // clip(), util/pism_utilities.hh:91-93
  }
  F.theta[q] = th;
}

// ---------------------------------------------------------------------------------------------
// surface gradients
// ---------------------------------------------------------------------------------------------
// SIAFD::surface_gradient_mahaffy, sia/SIAFD.cc:312-323, on owned + 1
__global__ void k_grad_mahaffy(const __grid_constant__ DP P, const Fields F) {
  const int nx = P.xm + 2;
  const long q = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= (long)nx * (P.ym + 2)) {
    return;
  }
  const int i = P.xs - 1 + (int)(q % nx), j = P.ys - 1 + (int)(q / nx);
  const double *h = F.h;
#define Hh(a, b) h[idx2(P, (a), (b), P.wg)]
  const long s = idx2(P, i, j, P.wst) * 2;
  F.h_x[s + 0] = (Hh(i + 1, j) - Hh(i, j)) / P.dx;
  F.h_y[s + 0] = (+Hh(i + 1, j + 1) + Hh(i, j + 1) - Hh(i + 1, j - 1) - Hh(i, j - 1)) / (4.0 * P.dy);
  F.h_y[s + 1] = (Hh(i, j + 1) - Hh(i, j)) / P.dy;
  F.h_x[s + 1] = (+Hh(i + 1, j + 1) + Hh(i + 1, j) - Hh(i - 1, j + 1) - Hh(i - 1, j)) / (4.0 * P.dx);
#undef Hh
}

// eta = H^((2n+2)/n) on owned + wg (sia/SIAFD.cc:241-245); stored in the w_i scratch field
__global__ void k_eta(const __grid_constant__ DP P, const Fields F) {
  const long n = (long)(P.xm + 2 * P.wg) * (P.ym + 2 * P.wg);
  const long q = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n) {
    return;
  }
  const double etapow = (2.0 * P.n + 2.0) / P.n;
  F.w_i[q] = pow(F.H[q], etapow);
}

// SIAFD::surface_gradient_eta, sia/SIAFD.cc:255-292, on owned + 1
__global__ void k_grad_eta(const __grid_constant__ DP P, const Fields F) {
  const int nx = P.xm + 2;
  const long q = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= (long)nx * (P.ym + 2)) {
    return;
  }
  const int i = P.xs - 1 + (int)(q % nx), j = P.ys - 1 + (int)(q / nx);
  const double n = P.n, etapow = (2.0 * n + 2.0) / n, invpow = 1.0 / etapow, dinvpow = (-n - 2.0) / (2.0 * n + 2.0);
  const double dx = P.dx, dy = P.dy;
#define Et(a, b) F.w_i[idx2(P, (a), (b), P.wg)]
#define Bd(a, b) F.bed[idx2(P, (a), (b), P.wg)]
  const double e_ij = Et(i, j), e_e = Et(i + 1, j), e_w = Et(i - 1, j), e_n = Et(i, j + 1), e_s = Et(i, j - 1),
               e_ne = Et(i + 1, j + 1), e_nw = Et(i - 1, j + 1), e_se = Et(i + 1, j - 1);
  const double b_ij = Bd(i, j), b_e = Bd(i + 1, j), b_w = Bd(i - 1, j), b_n = Bd(i, j + 1), b_s = Bd(i, j - 1),
               b_ne = Bd(i + 1, j + 1), b_nw = Bd(i - 1, j + 1), b_se = Bd(i + 1, j - 1);
#undef Et
#undef Bd
  const long s = idx2(P, i, j, P.wst) * 2;
  double hx, hy;
  {
    const double mean_eta = 0.5 * (e_e + e_ij);
    if (mean_eta > 0.0) {
      const double factor = invpow * pow(mean_eta, dinvpow);
      hx = __dmul_rn(factor, (e_e - e_ij)) / dx;
      hy = __dmul_rn(factor, (e_ne + e_n - e_se - e_s)) / (4.0 * dy);
    } else {
      hx = 0.0;
      hy = 0.0;
    }
    hx = __dadd_rn(hx, (b_e - b_ij) / dx);
    hy = __dadd_rn(hy, (b_ne + b_n - b_se - b_s) / (4.0 * dy));
    F.h_x[s + 0] = hx;
    F.h_y[s + 0] = hy;
  }
  {
    const double mean_eta = 0.5 * (e_n + e_ij);
    if (mean_eta > 0.0) {
      const double factor = invpow * pow(mean_eta, dinvpow);
      hx = __dmul_rn(factor, (e_ne + e_e - e_nw - e_w)) / (4.0 * dx);
      hy = __dmul_rn(factor, (e_n - e_ij)) / dy;
    } else {
      hx = 0.0;
      hy = 0.0;
    }
    hx = __dadd_rn(hx, (b_ne + b_e - b_nw - b_w) / (4.0 * dx));
    hy = __dadd_rn(hy, (b_n - b_ij) / dy);
    F.h_x[s + 1] = hx;
    F.h_y[s + 1] = hy;
  }
}

// SIAFD::surface_gradient_haseloff, first loop (sia/SIAFD.cc:396-436), on owned + 1
__global__ void k_grad_haseloff_a(const __grid_constant__ DP P, const Fields F) {
  const int nx = P.xm + 2;
  const long q = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= (long)nx * (P.ym + 2)) {
    return;
  }
  const int i = P.xs - 1 + (int)(q % nx), j = P.ys - 1 + (int)(q / nx);
  const long g0 = idx2(P, i, j, P.wg), ge = idx2(P, i + 1, j, P.wg), gn = idx2(P, i, j + 1, P.wg);
  const double h0 = F.h[g0], he = F.h[ge], hn = F.h[gn];
  const int M0 = mask_int(F.mask[g0]), Me = mask_int(F.mask[ge]), Mn = mask_int(F.mask[gn]);
  const long s = idx2(P, i, j, P.wst) * 2;
  // x-derivative, i-offset
  if ((m_floating_ice(M0) && m_ice_free_ocean(Me)) || (m_ice_free_ocean(M0) && m_floating_ice(Me))) {
    F.h_x[s + 0] = 0.0;
    F.w_i[g0] = 0.0;
  } else if ((m_icy(M0) && m_ice_free(Me) && he > h0) || (m_ice_free(M0) && m_icy(Me) && h0 > he)) {
    F.h_x[s + 0] = 0.0;
    F.w_i[g0] = 0.0;
  } else {
    F.h_x[s + 0] = (he - h0) / P.dx;
    F.w_i[g0] = 1.0;
  }
  // y-derivative, j-offset
  if ((m_floating_ice(M0) && m_ice_free_ocean(Mn)) || (m_ice_free_ocean(M0) && m_floating_ice(Mn))) {
    F.h_y[s + 1] = 0.0;
    F.w_j[g0] = 0.0;
  } else if ((m_icy(M0) && m_ice_free(Mn) && hn > h0) || (m_ice_free(M0) && m_icy(Mn) && h0 > hn)) {
    F.h_y[s + 1] = 0.0;
    F.w_j[g0] = 0.0;
  } else {
    F.h_y[s + 1] = (hn - h0) / P.dy;
    F.w_j[g0] = 1.0;
  }
}

// second loop (sia/SIAFD.cc:438-496), on owned points only
__global__ void k_grad_haseloff_b(const __grid_constant__ DP P, const Fields F) {
  const long q = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= (long)P.xm * P.ym) {
    return;
  }
  const int i = P.xs + (int)(q % P.xm), j = P.ys + (int)(q / P.xm);
#define WI(a, b) F.w_i[idx2(P, (a), (b), P.wg)]
#define WJ(a, b) F.w_j[idx2(P, (a), (b), P.wg)]
#define HX0(a, b) F.h_x[idx2(P, (a), (b), P.wst) * 2 + 0]
#define HY1(a, b) F.h_y[idx2(P, (a), (b), P.wst) * 2 + 1]
  const bool icy = m_icy(mask_int(F.mask[idx2(P, i, j, P.wg)]));
  const long s = idx2(P, i, j, P.wst) * 2;
  double r;
  // x-derivative, j-offset
  if (WJ(i, j) > 0) {
    const double W = WI(i, j) + WI(i - 1, j) + WI(i - 1, j + 1) + WI(i, j + 1);
    r = (W > 0) ? __dmul_rn(1.0 / W, (HX0(i, j) + HX0(i - 1, j) + HX0(i - 1, j + 1) + HX0(i, j + 1))) : 0.0;
  } else if (icy) {
    const double W = WI(i, j) + WI(i - 1, j);
    r = (W > 0) ? __dmul_rn(1.0 / W, (HX0(i, j) + HX0(i - 1, j))) : 0.0;
  } else {
    const double W = WI(i, j + 1) + WI(i - 1, j + 1);
    r = (W > 0) ? __dmul_rn(1.0 / W, (HX0(i - 1, j + 1) + HX0(i, j + 1))) : 0.0;
  }
  F.h_x[s + 1] = r;
  // y-derivative, i-offset
  if (WI(i, j) > 0) {
    const double W = WJ(i, j) + WJ(i, j - 1) + WJ(i + 1, j - 1) + WJ(i + 1, j);
    r = (W > 0) ? __dmul_rn(1.0 / W, (HY1(i, j) + HY1(i, j - 1) + HY1(i + 1, j - 1) + HY1(i + 1, j))) : 0.0;
  } else if (icy) {
    const double W = WJ(i, j) + WJ(i, j - 1);
    r = (W > 0) ? __dmul_rn(1.0 / W, (HY1(i, j) + HY1(i, j - 1))) : 0.0;
  } else {
    const double W = WJ(i + 1, j - 1) + WJ(i + 1, j);
    r = (W > 0) ? __dmul_rn(1.0 / W, (HY1(i + 1, j - 1) + HY1(i + 1, j))) : 0.0;
  }
  F.h_y[s + 0] = r;
#undef WI
#undef WJ
#undef HX0
#undef HY1
}

// ---------------------------------------------------------------------------------------------
// The fused diffusivity / flux / I / velocity kernel
// ---------------------------------------------------------------------------------------------
//
// CTA = strip of TX "extended" columns [ca, cb) x row segment [ra, rb) of the extended patch
// (owned + 1 ghost ring = the reference's PointsWithGhosts(1) iteration space).
// blockDim.x = 16 * (TX + 2): half-warp "groups" g = 0 .. TX+1; lane l = z level within a chunk
// of 16 levels.  Per row r, group g integrates two staggered columns together:
//   A: o = 0 (i-offset) at column c0 = ca - 1 + g   (g = 0 is the west halo column: its I feeds
//                                                     the u,v of column ca, its D belongs to the
//                                                     neighbouring strip)
//   B: o = 1 (j-offset) at column c1 = ca + g
// then (full update) group g writes u,v of the regular column c1 in row r from
//   I_e = I0[g+1], I_w = I0[g], I_n = I1[cur][g], I_s = I1[prev][g].
// Shared memory: z[Mz] | E rows: 3 slots x (TX+2) columns x Mz | [age rows likewise] |
//                I0: (TX+1) x Mz | I1: 2 x TX x Mz.
// Enthalpy row r+2 streams in (cp.async or one cp.async.bulk) while row r is integrated.

// lane l of group g fetches one scalar of row r:
//  0,1: thk_smooth at the two ends of A      2,3: theta at the two ends of A     4,5: h_x,h_y of A
//  6,7: thk_smooth at the two ends of B      8,9: theta at the two ends of B   10,11: h_x,h_y of B
// 12,13: sliding u,v at (c1, r)             14,15: h_x,h_y of the east staggered point (c1, r, 0)
__device__ __forceinline__ double fetch_scalar(const DP &P, const Fields &F, int r, int g, int l, int ca, int ncol,
                                               bool has_west, int ra, bool full) {
  const int c0 = ca - 1 + g, c1 = ca + g;
  const bool validA = (g <= ncol) && (g > 0 || has_west) && (r >= ra);
  const bool validB = (g < ncol);
  const double *ptr = nullptr;
  switch (l) {
  case 0:
    if (validA) ptr = F.thk_smooth + idx2(P, c0, r, P.wg);
    break;
  case 1:
    if (validA) ptr = F.thk_smooth + idx2(P, c0 + 1, r, P.wg);
    break;
  case 2:
    if (validA) ptr = F.theta + idx2(P, c0, r, P.wg);
    break;
  case 3:
    if (validA) ptr = F.theta + idx2(P, c0 + 1, r, P.wg);
    break;
  case 4:
    if (validA) ptr = F.h_x + idx2(P, c0, r, P.wst) * 2;
    break;
  case 5:
    if (validA) ptr = F.h_y + idx2(P, c0, r, P.wst) * 2;
    break;
  case 6:
    if (validB) ptr = F.thk_smooth + idx2(P, c1, r, P.wg);
    break;
  case 7:
    if (validB) ptr = F.thk_smooth + idx2(P, c1, r + 1, P.wg);
    break;
  case 8:
    if (validB) ptr = F.theta + idx2(P, c1, r, P.wg);
    break;
  case 9:
    if (validB) ptr = F.theta + idx2(P, c1, r + 1, P.wg);
    break;
  case 10:
    if (validB) ptr = F.h_x + idx2(P, c1, r, P.wst) * 2 + 1;
    break;
  case 11:
    if (validB) ptr = F.h_y + idx2(P, c1, r, P.wst) * 2 + 1;
    break;
  case 12:
  case 13: {
    const bool owned = validB && c1 >= P.xs && c1 < P.xs + P.xm && r >= P.ys && r < P.ys + P.ym;
    if (full && owned && F.sliding != nullptr) ptr = F.sliding + idx2(P, c1, r, P.wsl) * 2 + (l - 12);
    break;
  }
  case 14:
    if (full && validB && r >= ra) ptr = F.h_x + idx2(P, c1, r, P.wst) * 2;
    break;
  default:
    if (full && validB && r >= ra) ptr = F.h_y + idx2(P, c1, r, P.wst) * 2;
    break;
  }
  return ptr ? __ldg(ptr) : 0.0;
}

// IceGrid::kBelowHeight (util/IceGrid.cc:427-440; GSL bsearch: largest k in [0, Mz-2] with z[k] <= height)
__device__ __forceinline__ int k_below_height(const double *z_s, int Mz, double height, unsigned *err) {
  if (height < 0.0 - 1.0e-6) {
    atomicOr(err, EB_BELOW);
    return 0;
  }
  if (height > z_s[Mz - 1] + 1.0e-6) {
    atomicOr(err, EB_ABOVE);
    return 0;
  }
  int ilo = 0, ihi = Mz - 1;
  while (ihi > ilo + 1) {
    const int m = (ihi + ilo) >> 1;
    if (z_s[m] > height) {
      ihi = m;
    } else {
      ilo = m;
    }
  }
  return ilo;
}

__device__ __forceinline__ double scan16(double x, int l) {
#pragma unroll
  for (int d = 1; d < 16; d <<= 1) {
    const double y = __shfl_up_sync(FULLMASK, x, d, 16);
    if (l >= d) {
      x += y;
    }
  }
  return x;
}

__device__ __forceinline__ double sum16(double x) {
#pragma unroll
  for (int d = 8; d >= 1; d >>= 1) {
    x += __shfl_xor_sync(FULLMASK, x, d, 16);
  }
  return x;
}

template <int LAW, bool FULL>
__global__ void __launch_bounds__(288, 2)
    k_sia_fused(const __grid_constant__ DP P, const Fields F, const int TX, const int RS, const int use_bulk) {
  extern __shared__ __align__(16) double sm[];
  const int tid = threadIdx.x, NT = blockDim.x;
  const int l = tid & 15, g = tid >> 4;
  const int Mz = P.Mz;
  const int NCE = TX + 2; // enthalpy columns per row slot

  // ---- shared memory carve-up (all offsets in doubles; every region starts 16B-aligned) ----
  const int Mz2 = (Mz + 1) & ~1;
  const long slotE = ((long)NCE * Mz + 2 + 1) & ~1L; // +2: bulk copies may start one double early
  double *z_s = sm;
  double *E_s = z_s + Mz2;
  double *A_s = E_s + 3 * slotE; // age rows (only when P.use_age)
  double *I0_s = A_s + (P.use_age ? 3 * slotE : 0);
  double *I1_s = I0_s + (FULL ? (((long)(TX + 1) * Mz + 1) & ~1L) : 0);
  unsigned long long *bars = (unsigned long long *)(I1_s + (FULL ? (((long)2 * TX * Mz + 1) & ~1L) : 0));

  const int ca = (P.xs - 1) + blockIdx.x * TX;
  const int cb = min(ca + TX, P.xs + P.xm + 1);
  const int ncol = cb - ca;
  const int ra = (P.ys - 1) + blockIdx.y * RS;
  const int rb = min(ra + RS, P.ys + P.ym + 1);
  const bool has_west = blockIdx.x > 0;
  const int r0 = (FULL && blockIdx.y > 0) ? ra - 1 : ra; // warm-up row: I1 of the row below the segment

  for (int k = tid; k < Mz; k += NT) {
    z_s[k] = F.z[k];
  }
  if (use_bulk && tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_init(&bars[2], 1);
    fence_mbar_init();
  }
  __syncthreads();

  // ---- enthalpy (and age) row loader: columns [ca-1, cb] of row r -> slot ----
  const int rowcount = (ncol + 2) * Mz;
  const long NXe = P.xm + 2 * P.we;
  auto row_goff = [&](int r) -> long { return ((long)(r - (P.ys - P.we)) * NXe + (ca - 1 - (P.xs - P.we))) * Mz; };
  // with bulk copies the row lands shifted by (goff & 1) doubles so that source and destination are 16B-aligned
  auto issue_row = [&](int r, int slot) {
    const long goff = row_goff(r);
    if (use_bulk) {
      if (tid == 0) {
        const long a0 = goff & ~1L;
        const long a1 = (goff + rowcount + 1) & ~1L;
        unsigned bytes = (unsigned)((a1 - a0) * 8);
        if (P.use_age) {
          mbar_expect_tx(&bars[slot], 2 * bytes);
          bulk_g2s(E_s + slot * slotE, F.E + a0, bytes, &bars[slot]);
          bulk_g2s(A_s + slot * slotE, F.age + a0, bytes, &bars[slot]);
        } else {
          mbar_expect_tx(&bars[slot], bytes);
          bulk_g2s(E_s + slot * slotE, F.E + a0, bytes, &bars[slot]);
        }
      }
    } else {
      double *dst = E_s + slot * slotE;
      const double *src = F.E + goff;
      for (int e = tid; e < rowcount; e += NT) {
        cp_async8(dst + e, src + e);
      }
      if (P.use_age) {
        double *dstA = A_s + slot * slotE;
        const double *srcA = F.age + goff;
        for (int e = tid; e < rowcount; e += NT) {
          cp_async8(dstA + e, srcA + e);
        }
      }
      cp_async_commit();
    }
  };

  unsigned bar_phase = 0; // bit s = parity to wait for on bars[s]
  auto wait_row = [&](int slot) {
    if (use_bulk) {
      mbar_wait(&bars[slot], (bar_phase >> slot) & 1u);
      bar_phase ^= (1u << slot);
    }
  };

  // prologue: rows r0 and r0 + 1 (r0 + 1 <= rb always exists in the ghosted array)
  issue_row(r0, 0);
  issue_row(r0 + 1, 1);
  if (!use_bulk) {
    cp_async_wait_all();
  } else {
    wait_row(0);
    wait_row(1);
  }
  __syncthreads();

  double sc_next = fetch_scalar(P, F, r0, g, l, ca, ncol, has_west, ra, FULL);
  double prev_hxB = 0.0, prev_hyB = 0.0; // h_x, h_y of the j-offset point one row below (stage B "south")
  double dmax_local = 0.0;
  int hdc_local = 0;

  for (int r = r0; r < rb; ++r) {
    const int it = r - r0;
    const int s_cur = it % 3, s_nxt = (it + 1) % 3, s_pre = (it + 2) % 3;
    const bool prefetch = (r + 2 <= rb);
    if (prefetch) {
      issue_row(r + 2, s_pre);
    }
    const double sc = sc_next;
    if (r + 1 < rb) {
      sc_next = fetch_scalar(P, F, r + 1, g, l, ca, ncol, has_west, ra, FULL);
    }

    // ---------------- stage A: integrate the two staggered columns of this group ----------------
    const bool validA = (g <= ncol) && (g > 0 || has_west) && (r >= ra);
    const bool validB = (g < ncol);
    const double tsA0 = __shfl_sync(FULLMASK, sc, 0, 16), tsA1 = __shfl_sync(FULLMASK, sc, 1, 16);
    const double thA0 = __shfl_sync(FULLMASK, sc, 2, 16), thA1 = __shfl_sync(FULLMASK, sc, 3, 16);
    const double hxA = __shfl_sync(FULLMASK, sc, 4, 16), hyA = __shfl_sync(FULLMASK, sc, 5, 16);
    const double tsB0 = __shfl_sync(FULLMASK, sc, 6, 16), tsB1 = __shfl_sync(FULLMASK, sc, 7, 16);
    const double thB0 = __shfl_sync(FULLMASK, sc, 8, 16), thB1 = __shfl_sync(FULLMASK, sc, 9, 16);
    const double hxB = __shfl_sync(FULLMASK, sc, 10, 16), hyB = __shfl_sync(FULLMASK, sc, 11, 16);
    const double ub = __shfl_sync(FULLMASK, sc, 12, 16), vb = __shfl_sync(FULLMASK, sc, 13, 16);
    const double hxe = __shfl_sync(FULLMASK, sc, 14, 16), hye = __shfl_sync(FULLMASK, sc, 15, 16);

    // sia/SIAFD.cc:627-639
    const double thkA = 0.5 * (tsA0 + tsA1), thkB = 0.5 * (tsB0 + tsB1);
    const bool actA = validA && (thkA != 0.0), actB = validB && (thkB != 0.0);
    const int ksA = actA ? k_below_height(z_s, Mz, thkA, F.err) : -1;
    const int ksB = actB ? k_below_height(z_s, Mz, thkB, F.err) : -1;
    // sia/SIAFD.cc:686, :693-696
    const double alphaA = sqrt(hxA * hxA + hyA * hyA), alphaB = sqrt(hxB * hxB + hyB * hyB);
    const double thetaA = 0.5 * (thA0 + thA1), thetaB = 0.5 * (thB0 + thB1);
    const double c2A = P.e * thetaA * 2.0, c2B = P.e * thetaB * 2.0; // e_factor * theta_local * 2.0 (no age coupling)

    const double *Ecur = E_s + s_cur * slotE + (use_bulk ? (row_goff(r) & 1) : 0);
    const double *Enxt = E_s + s_nxt * slotE + (use_bulk ? (row_goff(r + 1) & 1) : 0);
    const double *EaA = Ecur + (long)g * Mz, *EbA = Ecur + (long)(g + 1) * Mz;
    const double *EaB = Ecur + (long)(g + 1) * Mz, *EbB = Enxt + (long)(g + 1) * Mz;
    const double *Acur = A_s + s_cur * slotE + (use_bulk ? (row_goff(r) & 1) : 0);
    const double *Anxt = A_s + s_nxt * slotE + (use_bulk ? (row_goff(r + 1) & 1) : 0);

    double *I0row = I0_s + (long)g * Mz;               // o = 0 point of this group (index g <-> column ca-1+g)
    double *I1row = I1_s + ((long)(it & 1) * TX + g) * Mz; // o = 1 point, slot by row parity

    int nch = max(actA ? (ksA >> 4) + 1 : 0, actB ? (ksB >> 4) + 1 : 0);
    nch = max(nch, __shfl_xor_sync(FULLMASK, nch, 16));

    double carryA = 0.0, carryB = 0.0, dpA = 0.0, dpB = 0.0, lastA = 0.0, lastB = 0.0;
    for (int c = 0; c < nch; ++c) {
      const int k = (c << 4) + l;
      const int kk = min(k, Mz - 1);
      const double zk = z_s[kk];
      const double dz = zk - z_s[max(kk - 1, 0)];
      const bool inA = (k <= ksA), inB = (k <= ksB);
      double dA = 0.0, dB = 0.0, depA = 0.0, depB = 0.0;
      if (inA) {
        depA = thkA - zk;                              // :641-643
        const double p = P.p_air + P.rg * depA;        // EnthalpyConverter.cc:146-152
        const double Eavg = 0.5 * (EaA[k] + EbA[k]);   // :677-684
        double c2 = c2A, gs = P.grain_size;
        if (P.use_age) {                               // :649-675
          const double age = 0.5 * (Acur[(long)g * Mz + k] + Acur[(long)(g + 1) * Mz + k]);
          if (P.gs_age) gs = grain_size_vostok(age * P.years_per_second);
          if (P.e_age) c2 = (interglacial(P, P.current_time - age) ? P.e_inter : P.e) * thetaA * 2.0;
        }
        const double fl = flow_eval<LAW>(P, alphaA * p, Eavg, p, gs); // :688-691
        dA = c2 * p * fl;                              // :696
      }
      if (inB) {
        depB = thkB - zk;
        const double p = P.p_air + P.rg * depB;
        const double Eavg = 0.5 * (EaB[k] + EbB[k]);
        double c2 = c2B, gs = P.grain_size;
        if (P.use_age) {
          const double age = 0.5 * (Acur[(long)(g + 1) * Mz + k] + Anxt[(long)(g + 1) * Mz + k]);
          if (P.gs_age) gs = grain_size_vostok(age * P.years_per_second);
          if (P.e_age) c2 = (interglacial(P, P.current_time - age) ? P.e_inter : P.e) * thetaB * 2.0;
        }
        const double fl = flow_eval<LAW>(P, alphaB * p, Eavg, p, gs);
        dB = c2 * p * fl;
      }
      // delta[k-1]: from the lane below, or the last lane of the previous chunk
      double pA = __shfl_up_sync(FULLMASK, dA, 1, 16), pB = __shfl_up_sync(FULLMASK, dB, 1, 16);
      if (l == 0) {
        pA = lastA;
        pB = lastB;
      }
      lastA = __shfl_sync(FULLMASK, dA, 15, 16);
      lastB = __shfl_sync(FULLMASK, dB, 15, 16);
      double tA = 0.0, tB = 0.0;
      if (inA && k >= 1) {
        tA = 0.5 * dz * (pA + dA);                                // compute_I, :855-858
        dpA += 0.5 * dz * ((depA + dz) * pA + depA * dA);         // D trapezoid, :701-705
      }
      if (inB && k >= 1) {
        tB = 0.5 * dz * (pB + dB);
        dpB += 0.5 * dz * ((depB + dz) * pB + depB * dB);
      }
      if (k == ksA) dpA += 0.5 * depA * depA * dA;                // :707-708 (dz = thk - z[ks] = depth[ks])
      if (k == ksB) dpB += 0.5 * depB * depB * dB;
      if (FULL) {
        const double IA = scan16(tA, l) + carryA, IB = scan16(tB, l) + carryB;
        carryA = __shfl_sync(FULLMASK, IA, 15, 16);
        carryB = __shfl_sync(FULLMASK, IB, 15, 16);
        if (k < Mz) {
          if (g <= ncol) I0row[k] = IA;
          if (validB) I1row[k] = IB;
        }
      }
    }
    if (FULL) {
      // above the ice (and ice-free / absent points): I stays at its last value (:861-863), 0 if no ice
      for (int k = (nch << 4) + l; k < Mz; k += 16) {
        if (g <= ncol) I0row[k] = carryA;
        if (validB) I1row[k] = carryB;
      }
    }

    // D, flux, D_max (lane 0 of the group), sia/SIAFD.cc:711-731, :772-793
    dpA = sum16(dpA);
    dpB = sum16(dpB);
    if (l == 0 && r >= ra) {
      const bool edge_r = (r < 0 || r >= P.My - 1);
      if (validA && g >= 1) { // own cell (c0 >= ca)
        const int c0 = ca - 1 + g;
        double D = actA ? dpA : 0.0;
        if (actA) {
          if (c0 < 0 || c0 >= P.Mx - 1 || edge_r) D = 0.0;
          if (P.limit_diffusivity && D >= P.D_limit) {
            D = P.D_limit;
            hdc_local += 1;
          }
          dmax_local = fmax(dmax_local, D);
        }
        const long s = idx2(P, c0, r, P.wst) * 2;
        F.D[s] = D;
        F.Q[s] = -D * hxA;
      }
      if (validB) {
        const int c1 = ca + g;
        double D = actB ? dpB : 0.0;
        if (actB) {
          if (c1 < 0 || c1 >= P.Mx - 1 || edge_r) D = 0.0;
          if (P.limit_diffusivity && D >= P.D_limit) {
            D = P.D_limit;
            hdc_local += 1;
          }
          dmax_local = fmax(dmax_local, D);
        }
        const long s = idx2(P, c1, r, P.wst) * 2 + 1;
        F.D[s] = D;
        F.Q[s] = -D * hyB;
      }
    }

    if (FULL) {
      __syncthreads();
      // ---------------- stage B: u, v of the regular column (c1, r), sia/SIAFD.cc:904-943 ----------------
      const int c1 = ca + g;
      if (validB && r >= ra && c1 >= P.xs && c1 < P.xs + P.xm && r >= P.ys && r < P.ys + P.ym) {
        const double *Ie = I0_s + (long)(g + 1) * Mz, *Iw = I0_s + (long)g * Mz;
        const double *In = I1_s + ((long)(it & 1) * TX + g) * Mz, *Is = I1_s + ((long)((it + 1) & 1) * TX + g) * Mz;
        const long o = idx2(P, c1, r, P.wuv) * Mz;
        for (int k = l; k < Mz; k += 16) {
          const double ie = Ie[k], iw = Iw[k], in = In[k], is = Is[k];
          F.u[o + k] = ub - 0.25 * (ie * hxe + iw * hxA + in * hxB + is * prev_hxB);
          F.v[o + k] = vb - 0.25 * (ie * hye + iw * hyA + in * hyB + is * prev_hyB);
        }
      }
      prev_hxB = hxB;
      prev_hyB = hyB;
    }

    if (!use_bulk) {
      cp_async_wait_all();
    } else if (prefetch) {
      wait_row(s_pre);
    }
    __syncthreads();
  }

  // ---- D_max / counter reduction: warp shuffle -> shared -> one atomic per CTA ----
  {
    unsigned long long m = (unsigned long long)__double_as_longlong(dmax_local);
    int cnt = hdc_local;
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
      const unsigned long long o = __shfl_xor_sync(FULLMASK, m, d);
      m = (o > m) ? o : m;
      cnt += __shfl_xor_sync(FULLMASK, cnt, d);
    }
    __shared__ unsigned long long wm[32];
    __shared__ int wc[32];
    if ((tid & 31) == 0) {
      wm[tid >> 5] = m;
      wc[tid >> 5] = cnt;
    }
    __syncthreads();
    if (tid == 0) {
      const int nw = (NT + 31) >> 5;
      for (int w = 1; w < nw; ++w) {
        m = (wm[w] > m) ? wm[w] : m;
        cnt += wc[w];
      }
      if (m != 0ull) atomicMax(F.dmax, m);
      if (cnt != 0) atomicAdd(F.hdc, cnt);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// rectangle copy (ghost wrap, halo pack / unpack)
// ---------------------------------------------------------------------------------------------
__global__ void k_copy_region(double *__restrict__ dst, long dst_row_cells, int dst_i0, int dst_j0,
                              const double *__restrict__ src, long src_row_cells, int src_i0, int src_j0, int wc, int hc,
                              int dof) {
  const long rowlen = (long)wc * dof;
  const long n = rowlen * hc;
  for (long q = (long)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (long)gridDim.x * blockDim.x) {
    const long jj = q / rowlen, e = q - jj * rowlen;
    dst[((dst_j0 + jj) * dst_row_cells + dst_i0) * dof + e] = src[((src_j0 + jj) * src_row_cells + src_i0) * dof + e];
  }
}

// ---------------------------------------------------------------------------------------------
// GeometryCalculator::compute, util/Mask.hh:96-133
// ---------------------------------------------------------------------------------------------
__global__ void k_geometry(const __grid_constant__ DP P, long n, const double *sea_level, const double *bed,
                           const double *thk, double *mask_out, double *surf_out) {
  const long q = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n) {
    return;
  }
  const double hgrounded = bed[q] + thk[q];
  const double hfloating = __dadd_rn(sea_level[q], __dmul_rn(P.gc_alpha, thk[q])); // no FMA: bit-exact surface
  const bool is_floating = (hfloating > hgrounded), ice_free = (thk[q] <= P.gc_icefree);
  double m, s;
  if (is_floating && !P.gc_dry) {
    s = hfloating;
    m = ice_free ? 4.0 : 3.0;
  } else {
    s = hgrounded;
    m = ice_free ? 0.0 : 2.0;
  }
  if (mask_out) mask_out[q] = m;
  if (surf_out) surf_out[q] = s;
}

// ---------------------------------------------------------------------------------------------
// FlowLaw::flow_n, rheology/FlowLaw.cc:107-119
// ---------------------------------------------------------------------------------------------
template <int LAW>
__global__ void k_flow_n(const __grid_constant__ DP P, long n, const double *stress, const double *E, const double *p,
                         const double *gs, double *out) {
  const long q = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (q < n) {
    out[q] = flow_eval<LAW>(P, stress[q], E[q], p[q], gs ? gs[q] : P.grain_size);
  }
}

// ---------------------------------------------------------------------------------------------
// BedSmoother::preprocess_bed, sia/BedSmoother.cc:157-267: one thread per local point
// (owned + wg ghosts; ghosts are the periodic images, as get_from_proc0 delivers them)
// ---------------------------------------------------------------------------------------------
__global__ void k_bed_smooth(const __grid_constant__ DP P, const double *__restrict__ b0, int Nx, int Ny,
                             double *topgsmooth, double *maxtl, double *C2, double *C3, double *C4) {
  const int nx = P.xm + 2 * P.wg;
  const long q = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= (long)nx * (P.ym + 2 * P.wg)) {
    return;
  }
  int i = P.xs - P.wg + (int)(q % nx), j = P.ys - P.wg + (int)(q / nx);
  i = ((i % P.Mx) + P.Mx) % P.Mx;
  j = ((j % P.My) + P.My) % P.My;
  const int Mx = P.Mx, My = P.My;
  // smooth_the_bed_on_proc0 (:157-191): plain sum in (r, s) order, in-domain points only
  double sum = 0.0, count = 0.0;
  for (int r = -Nx; r <= Nx; r++) {
    for (int s = -Ny; s <= Ny; s++) {
      if ((i + r >= 0) && (i + r < Mx) && (j + s >= 0) && (j + s < My)) {
        sum = __dadd_rn(sum, b0[(long)(j + s) * Mx + (i + r)]);
        count += 1.0;
      }
    }
  }
  const double topgs = sum / count;
  // compute_coefficients_on_proc0 (:194-267)
  double maxtltemp = 0.0, sum2 = 0.0, sum3 = 0.0, sum4 = 0.0;
  for (int r = -Nx; r <= Nx; r++) {
    for (int s = -Ny; s <= Ny; s++) {
      if ((i + r >= 0) && (i + r < Mx) && (j + s >= 0) && (j + s < My)) {
        const double tl = b0[(long)(j + s) * Mx + (i + r)] - topgs;
        maxtltemp = fmax(maxtltemp, tl);
        const double tl2 = __dmul_rn(tl, tl);
        sum2 = __dadd_rn(sum2, tl2);
        sum3 = __dadd_rn(sum3, __dmul_rn(tl2, tl));
        sum4 = __dadd_rn(sum4, __dmul_rn(tl2, tl2));
      }
    }
  }
  const double n = P.n, k = (n + 2) / n, s2 = k * (2 * n + 2) / (2 * n), s3 = s2 * (3 * n + 2) / (3 * n),
               s4 = s3 * (4 * n + 2) / (4 * n);
  topgsmooth[q] = topgs;
  maxtl[q] = maxtltemp;
  C2[q] = __dmul_rn(s2, sum2 / count);
  C3[q] = __dmul_rn(s3, sum3 / count);
  C4[q] = __dmul_rn(s4, sum4 / count);
}

// =============================================================================================
// launchers
// =============================================================================================
static inline unsigned nblk(long n, int t) { return (unsigned)((n + t - 1) / t); }

int launch_prep2d(const DP &P, const Fields &F, cudaStream_t s) {
  const long n = (long)(P.xm + 2 * P.wg) * (P.ym + 2 * P.wg);
  k_prep2d<<<nblk(n, 256), 256, 0, s>>>(P, F);
  return 1;
}

int launch_gradient(const DP &P, const Fields &F, cudaStream_t s) {
  const long n1 = (long)(P.xm + 2) * (P.ym + 2);
  const long n2 = (long)(P.xm + 2 * P.wg) * (P.ym + 2 * P.wg);
  switch (P.grad) { // sia/SIAFD.cc:197-220
  case GRAD_MAHAFFY:
    k_grad_mahaffy<<<nblk(n1, 256), 256, 0, s>>>(P, F);
    return 1;
  case GRAD_ETA:
    k_eta<<<nblk(n2, 256), 256, 0, s>>>(P, F);
    k_grad_eta<<<nblk(n1, 256), 256, 0, s>>>(P, F);
    return 2;
  default:
    k_grad_haseloff_a<<<nblk(n1, 256), 256, 0, s>>>(P, F);
    k_grad_haseloff_b<<<nblk((long)P.xm * P.ym, 256), 256, 0, s>>>(P, F);
    return 2;
  }
}

size_t fused_smem_bytes(const DP &P, bool full, int TX) {
  const long Mz = P.Mz, Mz2 = (Mz + 1) & ~1L;
  const long slotE = ((long)(TX + 2) * Mz + 2 + 1) & ~1L;
  long d = Mz2 + 3 * slotE + (P.use_age ? 3 * slotE : 0);
  if (full) {
    d += (((long)(TX + 1) * Mz + 1) & ~1L) + (((long)2 * TX * Mz + 1) & ~1L);
  }
  return (size_t)d * 8 + 3 * 8 /* mbarriers */ + 16;
}

// largest strip width whose shared memory allows two CTAs per SM (else one), 227 KB usable per SM
int pick_tile_x(const DP &P, bool full) {
  const int cand[3] = {16, 8, 4};
  for (int q = 0; q < 3; ++q) {
    if (fused_smem_bytes(P, full, cand[q]) + 1024 <= (size_t)(227 * 1024) / 2) return cand[q];
  }
  for (int q = 0; q < 3; ++q) {
    if (fused_smem_bytes(P, full, cand[q]) <= (size_t)227 * 1024) return cand[q];
  }
  return 0; // does not fit: Mz too large
}

template <int LAW, bool FULL>
static int launch_fused_t(const DP &P, const Fields &F, const Tuning &T, cudaStream_t s) {
  const int TX = T.tile_x;
  const int RS = T.rows_per_cta;
  const size_t smem = fused_smem_bytes(P, FULL, TX);
  static size_t configured = 0; // per instantiation
  if (smem > configured) {
    if (cudaFuncSetAttribute(k_sia_fused<LAW, FULL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) !=
        cudaSuccess) {
      return -1;
    }
    configured = smem;
  }
  dim3 grid((unsigned)((P.xm + 2 + TX - 1) / TX), (unsigned)((P.ym + 2 + RS - 1) / RS));
  k_sia_fused<LAW, FULL><<<grid, 16 * (TX + 2), smem, s>>>(P, F, TX, RS, T.use_bulk_copy);
  return 1;
}

template <bool FULL> static int launch_fused_f(const DP &P, const Fields &F, const Tuning &T, cudaStream_t s) {
  switch (P.law) {
  case LAW_ISO:
    return launch_fused_t<LAW_ISO, FULL>(P, F, T, s);
  case LAW_PB:
    return launch_fused_t<LAW_PB, FULL>(P, F, T, s);
  case LAW_GPBLD:
    return launch_fused_t<LAW_GPBLD, FULL>(P, F, T, s);
  case LAW_HOOKE:
    return launch_fused_t<LAW_HOOKE, FULL>(P, F, T, s);
  case LAW_ARR:
    return launch_fused_t<LAW_ARR, FULL>(P, F, T, s);
  case LAW_ARRWARM:
    return launch_fused_t<LAW_ARRWARM, FULL>(P, F, T, s);
  case LAW_GK:
    return launch_fused_t<LAW_GK, FULL>(P, F, T, s);
  default:
    return -1;
  }
}

int launch_fused(const DP &P, const Fields &F, bool full, const Tuning &T, cudaStream_t s) {
  return full ? launch_fused_f<true>(P, F, T, s) : launch_fused_f<false>(P, F, T, s);
}

int launch_copy_region(double *dst, long dst_row_cells, int dst_i0, int dst_j0, const double *src, long src_row_cells,
                       int src_i0, int src_j0, int wc, int hc, int dof, cudaStream_t s) {
  const long n = (long)wc * hc * dof;
  if (n <= 0) return 0;
  unsigned blocks = nblk(n, 256);
  if (blocks > 148u * 32u) blocks = 148u * 32u; // grid-stride beyond that
  k_copy_region<<<blocks, 256, 0, s>>>(dst, dst_row_cells, dst_i0, dst_j0, src, src_row_cells, src_i0, src_j0, wc, hc,
                                       dof);
  return 1;
}

int launch_geometry(const DP &P, long n, const double *sea_level, const double *bed, const double *thk, double *mask_out,
                    double *surf_out, cudaStream_t s) {
  if (n <= 0) return 0;
  k_geometry<<<nblk(n, 256), 256, 0, s>>>(P, n, sea_level, bed, thk, mask_out, surf_out);
  return 1;
}

int launch_flow_n(const DP &P, long n, const double *stress, const double *E, const double *p, const double *gs,
                  double *out, cudaStream_t s) {
  if (n <= 0) return 0;
  const unsigned b = nblk(n, 256);
  switch (P.law) {
  case LAW_ISO:
    k_flow_n<LAW_ISO><<<b, 256, 0, s>>>(P, n, stress, E, p, gs, out);
    break;
  case LAW_PB:
    k_flow_n<LAW_PB><<<b, 256, 0, s>>>(P, n, stress, E, p, gs, out);
    break;
  case LAW_GPBLD:
    k_flow_n<LAW_GPBLD><<<b, 256, 0, s>>>(P, n, stress, E, p, gs, out);
    break;
  case LAW_HOOKE:
    k_flow_n<LAW_HOOKE><<<b, 256, 0, s>>>(P, n, stress, E, p, gs, out);
    break;
  case LAW_ARR:
    k_flow_n<LAW_ARR><<<b, 256, 0, s>>>(P, n, stress, E, p, gs, out);
    break;
  case LAW_ARRWARM:
    k_flow_n<LAW_ARRWARM><<<b, 256, 0, s>>>(P, n, stress, E, p, gs, out);
    break;
  case LAW_GK:
    k_flow_n<LAW_GK><<<b, 256, 0, s>>>(P, n, stress, E, p, gs, out);
    break;
  default:
    return -1;
  }
  return 1;
}

int launch_preprocess_bed(const DP &P, const double *global_bed, int Nx, int Ny, double *topgsmooth, double *maxtl,
                          double *C2, double *C3, double *C4, cudaStream_t s) {
  const long n = (long)(P.xm + 2 * P.wg) * (P.ym + 2 * P.wg);
  k_bed_smooth<<<nblk(n, 128), 128, 0, s>>>(P, global_bed, Nx, Ny, topgsmooth, maxtl, C2, C3, C4);
  return 1;
}

} // namespace siafd
