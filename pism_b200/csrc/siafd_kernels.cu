// siafd_kernels.cu -- sm_100a kernels of the SIAFD hot path.
//
// Reference path: stressbalance::SIAFD::update() of juliusgarbe/pism v1.2.1
// (src/stressbalance/sia/SIAFD.cc:122-155).  Kernel map:
//
//   k_prep2d        BedSmoother::smoothed_thk + ::theta      sia/BedSmoother.cc:284-327, :351-404
//   k_eta           eta = H^((2n+2)/n)                        sia/SIAFD.cc:241-245
//   k_grad_*        surface_gradient_{mahaffy,eta,haseloff}   sia/SIAFD.cc:224-496 (haseloff: both loops in one pass)
//   k_sia_slab      (siafd_slab.cu) diffusivity + flux + I + 3D velocity, fused    sia/SIAFD.cc:543-948
//   k_copy_region   ghost wrap / halo pack (DMLocalToLocal)   util/iceModelVec.cc:630-643
//   k_geometry      GeometryCalculator::compute               util/Mask.hh:96-133
//   k_flow_n        FlowLaw::flow_n                           rheology/FlowLaw.cc:107-119
//   k_bed_*         BedSmoother::preprocess_bed               sia/BedSmoother.cc:157-267
//
// The fused kernel never writes delta or I to HBM (the reference round-trips four 3D scratch
// fields); see siafd_slab.cu and DESIGN.md.
#include "siafd_kernels.cuh"

#include <algorithm>
#include <cstdio>

namespace siafd {

#define FULLMASK 0xffffffffu

// ---------------------------------------------------------------------------------------------
// k_prep2d: thk_smooth and theta on owned + wg ghosts (no communication, like the reference)
// ---------------------------------------------------------------------------------------------
// returns the row segment of the fused kernel the point counts for (it holds ice), or -1
__device__ __forceinline__ int prep2d_point(const DP &P, const Fields &F, const long q) {
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
  int seg = -1;
  if (P.seg_n > 0 && ts > 0.0) {
    const int j = (int)(q / (P.xm + 2 * P.wg)) - P.wg; // row relative to the first owned row
    seg = min(max(j + 1, 0) / P.seg_rows, P.seg_n - 1);
  }

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
  return seg;
}

// weight of the fused kernel's row segments = icy points per segment: counted per CTA in shared memory (a CTA's points
// lie in a few consecutive rows), one global atomic per CTA and segment.  Called by every thread of the CTA.
// j0: the row (relative to the first owned row) of the CTA's first point; n: icy points this thread adds to `seg`
__device__ __forceinline__ void seg_weight_cta(const DP &P, const Fields &F, const int seg, const int j0, const int n = 1) {
  if (P.seg_n <= 0) return;
  __shared__ int cnt[8];
  const int seg_lo = min(max(j0 + 1, 0) / P.seg_rows, P.seg_n - 1);
  if (threadIdx.x < 8) cnt[threadIdx.x] = 0;
  __syncthreads();
  if (seg >= 0 && n > 0) {
    const int d = seg - seg_lo;
    if (d < 8) {
      atomicAdd(&cnt[d], n);
    } else {
      atomicAdd(F.segw + seg, n);
    }
  }
  __syncthreads();
  if (threadIdx.x < 8 && cnt[threadIdx.x] != 0) atomicAdd(F.segw + seg_lo + threadIdx.x, cnt[threadIdx.x]);
  __syncthreads(); // (the counters are reused by the CTA's next batch of points)
}

// The last CTA of the 2D pass sorts the row segments of the fused kernel by weight, heaviest first (ties by index): the
// expensive (icy) segments start first and the cheap ones fill the tail of the launch.
__device__ __forceinline__ void seg_order_epilogue(const DP &P, const Fields &F) {
  if (P.seg_n <= 0) return;
  __shared__ int last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    last = (atomicAdd(F.segdone, 1u) == gridDim.x - 1);
    if (last) *F.segdone = 0u;
  }
  __syncthreads();
  if (!last) return;
  __threadfence();
  for (int s = threadIdx.x; s < P.seg_n; s += blockDim.x) {
    const int ws = *(volatile int *)(F.segw + s);
    int rank = 0;
    for (int t = 0; t < P.seg_n; ++t) {
      const int wt = *(volatile int *)(F.segw + t);
      rank += (wt > ws) || (wt == ws && t < s);
    }
    F.segw[128 + rank] = s;
  }
}

__global__ void k_prep2d(const __grid_constant__ DP P, const Fields F) {
  const long n = (long)(P.xm + 2 * P.wg) * (P.ym + 2 * P.wg);
  const long q = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const int seg = (q < n) ? prep2d_point(P, F, q) : -1;
  seg_weight_cta(P, F, seg, (int)(((long)blockIdx.x * blockDim.x) / (P.xm + 2 * P.wg)) - P.wg);
  seg_order_epilogue(P, F);
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

// SIAFD::surface_gradient_haseloff in ONE pass.  The reference runs two loops (sia/SIAFD.cc:396-436 on owned + 1:
// the direct components h_x(.,0), h_y(.,1) and the weights w_i, w_j; :438-496 on owned: the cross components
// as weighted means of up to four direct neighbours).  A direct component and its weight are a function of two
// adjacent cells only (surface + mask), so the second loop's neighbours are re-evaluated here from the cells
// around (i, j) -- the same expressions, hence the same bits -- instead of being round-tripped through w_i,
// w_j and the half-written staggered arrays.
struct HasDirect {
  double g; // the direct gradient component
  double w; // its weight (0 or 1)
};
// i-offset point between (i, j) and (i + 1, j): sia/SIAFD.cc:399-416
__device__ __forceinline__ HasDirect haseloff_direct(double h0, double h1, int M0, int M1, double d, double inv_d) {
  HasDirect r;
  if ((m_floating_ice(M0) && m_ice_free_ocean(M1)) || (m_ice_free_ocean(M0) && m_floating_ice(M1))) {
    r.g = 0.0, r.w = 0.0;
  } else if ((m_icy(M0) && m_ice_free(M1) && h1 > h0) || (m_ice_free(M0) && m_icy(M1) && h0 > h1)) {
    r.g = 0.0, r.w = 0.0;
  } else {
    r.g = div_rn_by(h1 - h0, d, inv_d), r.w = 1.0;
  }
  return r;
}
// 1.0 / W for a sum W of up to four 0/1 weights (W in {1, 2, 3, 4}): the correctly rounded quotients
__device__ __forceinline__ double inv_count(double W) {
  return W == 1.0 ? 1.0 : (W == 2.0 ? 0.5 : (W == 3.0 ? (1.0 / 3.0) : 0.25));
}
// PUSH: the ghost update of sia/SIAFD.cc:498-499 fused in -- an owned point on the rim of the patch is also stored into
// the neighbours' ghost cells (PeerPush, siafd_kernels.cuh).  One point per thread, on owned + 1: the kernel of the
// split calls (siafd_b200_compute_gradient); the whole-step entry points take k_grad_haseloff_quad below.
template <bool PUSH>
__device__ __forceinline__ void grad_haseloff_point(const DP &P, const Fields &F, const PeerPush &PP, const long q) {
  const int nx = P.xm + 2;
  if (q >= (long)nx * (P.ym + 2)) {
    return;
  }
  const int i = P.xs - 1 + (int)(q % nx), j = P.ys - 1 + (int)(q / nx);
  // the 3 x 3 cells around (i, j): c[b][a] = cell (i - 1 + a, j - 1 + b)
  double h[3][3];
  int M[3][3];
#pragma unroll
  for (int b = 0; b < 3; ++b) {
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const long g = idx2(P, i - 1 + a, j - 1 + b, P.wg);
      h[b][a] = F.h[g];
      M[b][a] = mask_int(F.mask[g]);
    }
  }
  const long s = idx2(P, i, j, P.wst) * 2;
  // direct components of this point (first loop)
  const HasDirect x00 = haseloff_direct(h[1][1], h[1][2], M[1][1], M[1][2], P.dx, P.inv_dx); // h_x(i, j, 0), w_i(i, j)
  const HasDirect y00 = haseloff_direct(h[1][1], h[2][1], M[1][1], M[2][1], P.dy, P.inv_dy); // h_y(i, j, 1), w_j(i, j)
  F.h_x[s + 0] = x00.g;
  F.h_y[s + 1] = y00.g;
  // (the weights w_i, w_j of the reference's work vectors are not stored: nothing reads them after this kernel)
  if (i < P.xs || i >= P.xs + P.xm || j < P.ys || j >= P.ys + P.ym) {
    return; // the second loop runs over owned points only; its ghosts come from the exchange (:498-499)
  }
  const bool icy = m_icy(M[1][1]);
  // neighbours of the second loop, each evaluated as the first loop evaluates it at its own point
  const HasDirect xm0 = haseloff_direct(h[1][0], h[1][1], M[1][0], M[1][1], P.dx, P.inv_dx); // (i-1, j)
  const HasDirect xm1 = haseloff_direct(h[2][0], h[2][1], M[2][0], M[2][1], P.dx, P.inv_dx); // (i-1, j+1)
  const HasDirect x01 = haseloff_direct(h[2][1], h[2][2], M[2][1], M[2][2], P.dx, P.inv_dx); // (i, j+1)
  const HasDirect y0m = haseloff_direct(h[0][1], h[1][1], M[0][1], M[1][1], P.dy, P.inv_dy); // (i, j-1)
  const HasDirect y1m = haseloff_direct(h[0][2], h[1][2], M[0][2], M[1][2], P.dy, P.inv_dy); // (i+1, j-1)
  const HasDirect y10 = haseloff_direct(h[1][2], h[2][2], M[1][2], M[2][2], P.dy, P.inv_dy); // (i+1, j)
  double r;
  // x-derivative, j-offset (:441-467)
  if (y00.w > 0) {
    const double W = x00.w + xm0.w + xm1.w + x01.w;
    r = (W > 0) ? __dmul_rn(inv_count(W), (x00.g + xm0.g + xm1.g + x01.g)) : 0.0;
  } else if (icy) {
    const double W = x00.w + xm0.w;
    r = (W > 0) ? __dmul_rn(inv_count(W), (x00.g + xm0.g)) : 0.0;
  } else {
    const double W = x01.w + xm1.w;
    r = (W > 0) ? __dmul_rn(inv_count(W), (xm1.g + x01.g)) : 0.0;
  }
  F.h_x[s + 1] = r;
  const double hx_cross = r;
  // y-derivative, i-offset (:469-495)
  if (x00.w > 0) {
    const double W = y00.w + y0m.w + y1m.w + y10.w;
    r = (W > 0) ? __dmul_rn(inv_count(W), (y00.g + y0m.g + y1m.g + y10.g)) : 0.0;
  } else if (icy) {
    const double W = y00.w + y0m.w;
    r = (W > 0) ? __dmul_rn(inv_count(W), (y00.g + y0m.g)) : 0.0;
  } else {
    const double W = y1m.w + y10.w;
    r = (W > 0) ? __dmul_rn(inv_count(W), (y1m.g + y10.g)) : 0.0;
  }
  F.h_y[s + 0] = r;
  if (PUSH) {
    const int a = i - P.xs, b = j - P.ys;
    const bool W_ = a < PP.w, E_ = a >= P.xm - PP.w, S_ = b < PP.w, N_ = b >= P.ym - PP.w;
    if (W_ || E_ || S_ || N_) {
      const double2 hx2 = make_double2(x00.g, hx_cross), hy2 = make_double2(r, y00.g);
#pragma unroll
      for (int d = 0; d < 8; ++d) {
        if (PP.a[d] != nullptr && peer_strip_member(d, W_, E_, S_, N_)) {
          const long t = ((long)(b + P.wst + PP.dj[d]) * PP.rowc[d] + (a + P.wst + PP.di[d])) * 2;
          *reinterpret_cast<double2 *>(PP.a[d] + t) = hx2;
          *reinterpret_cast<double2 *>(PP.b[d] + t) = hy2;
        }
      }
    }
  }
  return;
}

template <bool PUSH>
__global__ void __launch_bounds__(256, 6)
    k_grad_haseloff(const __grid_constant__ DP P, const Fields F, const __grid_constant__ PeerPush PP) {
  grad_haseloff_point<PUSH>(P, F, PP, (long)blockIdx.x * blockDim.x + threadIdx.x);
}

// Gradient AND thk_smooth / theta (k_prep2d) in one pass, GQ_ROWS points per thread: a column of rows, the lanes of a
// warp across x; the threads cover owned + wg ghosts, those on owned + 1 go on to the gradient.  The
// 3 x 3 neighbourhoods of vertically adjacent points overlap: with two rows a thread loads 3 x 4 cells and evaluates 12
// direct components where two single-point threads load 18 cells and evaluate 16, the index arithmetic is shared, and
// every load and store of a warp is a contiguous run of 8- or 16-byte items.  Every value is produced by the same
// expression as in grad_haseloff_point (the tests pin the gradients bit for bit).  The pass is bound by the latency of
// its chains of predicated selects, not by DRAM (profiles/ncu_r02_k_grad_haseloff_4096_summary.txt: DRAM at 30 %):
// what counts is warps per SM times work shared per thread, hence two rows at 64 registers (8 CTAs per SM).
#ifndef GQ_ROWS
#define GQ_ROWS 2 // points (rows) per thread (measured at 4096^2, B200: 1 / 2 / 4 rows at 8 / 8 / 4 CTAs per SM: 0.54 / 0.47 / 0.65 ms)
#endif
#ifndef GQ_MINB
#define GQ_MINB 8 // CTAs per SM the register allocation aims at (the pass is bound by latency: warps count)
#endif
constexpr int GQ_T = 128;    // threads per CTA = columns per tile; a tile is GQ_T columns x GQ_R rows
constexpr int GQ_R = GQ_ROWS;
template <bool PUSH, bool LOCAL_RING>
__global__ void __launch_bounds__(GQ_T, GQ_MINB)
    k_grad_haseloff_quad(const __grid_constant__ DP P, const Fields F, const __grid_constant__ PeerPush PP) {
  const int wg = P.wg;
  const int nx = P.xm + 2 * wg, ny = P.ym + 2 * wg;
  const int ntx = (nx + GQ_T - 1) / GQ_T, nty = (ny + GQ_R - 1) / GQ_R;
  const long nt = (long)ntx * nty;
  __shared__ int cnt[8];
  for (long t = blockIdx.x; t < nt; t += gridDim.x) {
    const int ty = (int)(t / ntx), tx = (int)(t - (long)ty * ntx);
    const int c0 = tx * GQ_T + threadIdx.x, jy0 = GQ_R * ty; // column / first row in the geometry-width array
    // weights of the fused kernel's row segments: icy points per segment, counted per CTA in shared memory
    const int seg_lo = (P.seg_n > 0) ? min(max(jy0 - wg + 1, 0) / P.seg_rows, P.seg_n - 1) : 0;
    if (P.seg_n > 0) {
      if (threadIdx.x < 8) cnt[threadIdx.x] = 0;
      __syncthreads();
    }
    if (c0 < nx) {
      // thk_smooth, theta (k_prep2d) of the four points
#pragma unroll
      for (int p = 0; p < GQ_R; ++p) {
        if (jy0 + p < ny) {
          const int sp = prep2d_point(P, F, (long)(jy0 + p) * nx + c0);
          if (sp >= 0) {
            const int d = sp - seg_lo;
            if (d < 8) {
              atomicAdd(&cnt[d], 1);
            } else {
              atomicAdd(F.segw + sp, 1);
            }
          }
        }
      }
      // gradient on owned + 1: rows ys - 1 .. ys + ym, columns xs - 1 .. xs + xm
      const int i = P.xs - wg + c0, ja = P.ys - wg + jy0; // column, row of point 0
      if (i >= P.xs - 1 && i <= P.xs + P.xm && ja + GQ_R - 1 >= P.ys - 1 && ja <= P.ys + P.ym) {
        // cells: column index a = 0..2 <-> column i - 1 + a, row index b = 0..5 <-> row ja - 1 + b (clamped to the array:
        // a clamped cell is only ever read by a point outside owned + 1, which is not stored)
        double h[GQ_R + 2][3];
        int M[GQ_R + 2][3];
#pragma unroll
        for (int b = 0; b < GQ_R + 2; ++b) {
          const int rb = min(max(jy0 - 1 + b, 0), ny - 1);
#pragma unroll
          for (int a = 0; a < 3; ++a) {
            const int ca = min(max(c0 - 1 + a, 0), nx - 1);
            const long g = (long)rb * nx + ca;
            h[b][a] = F.h[g];
            M[b][a] = mask_int(F.mask[g]);
          }
        }
        // direct components: X[a][b] at the i-offset point of cell (a, b + 1), Y[a][b] at the j-offset point of cell (a + 1, b)
        HasDirect X[2][GQ_R + 1], Y[2][GQ_R + 1];
#pragma unroll
        for (int a = 0; a < 2; ++a) {
#pragma unroll
          for (int b = 0; b < GQ_R + 1; ++b) {
            X[a][b] = haseloff_direct(h[b + 1][a], h[b + 1][a + 1], M[b + 1][a], M[b + 1][a + 1], P.dx, P.inv_dx);
            Y[a][b] = haseloff_direct(h[b][a + 1], h[b + 1][a + 1], M[b][a + 1], M[b + 1][a + 1], P.dy, P.inv_dy);
          }
        }
        // The reference's second loop runs over the owned points and a ghost update (sia/SIAFD.cc:498-499) brings the
        // cross components of the ring around them.  With geometry ghosts two cells wide (PISM's WIDE_STENCIL; the
        // kernel's precondition) a point of that ring has its whole 3 x 3 neighbourhood in this rank's arrays, and the
        // owner's value is the same expression of the same cells: it is evaluated here, bit for bit, and the exchange --
        // a synchronisation of all ranks in the middle of the step -- is not needed.
        const bool col_owned = i >= P.xs && i < P.xs + P.xm;
#pragma unroll
        for (int p = 0; p < GQ_R; ++p) {
          const int j = ja + p;
          if (j < P.ys - 1 || j > P.ys + P.ym) continue;
          const HasDirect x00 = X[1][p], y00 = Y[0][p + 1]; // h_x(i, j, 0), w_i(i, j); h_y(i, j, 1), w_j(i, j)
          const long s2 = idx2(P, i, j, P.wst) * 2;
          if (!LOCAL_RING && (!col_owned || j < P.ys || j >= P.ys + P.ym)) {
            F.h_x[s2 + 0] = x00.g; // (geometry ghosts one cell wide: the ring's cross components come from an exchange)
            F.h_y[s2 + 1] = y00.g;
            continue;
          }
          const bool icy = m_icy(M[p + 1][1]);
          const HasDirect xm0 = X[0][p], xm1 = X[0][p + 1], x01 = X[1][p + 1]; // (i-1, j), (i-1, j+1), (i, j+1)
          const HasDirect y0m = Y[0][p], y1m = Y[1][p], y10 = Y[1][p + 1];     // (i, j-1), (i+1, j-1), (i+1, j)
          double r;
          if (y00.w > 0) { // x-derivative, j-offset (:441-467)
            const double W = x00.w + xm0.w + xm1.w + x01.w;
            r = (W > 0) ? __dmul_rn(inv_count(W), (x00.g + xm0.g + xm1.g + x01.g)) : 0.0;
          } else if (icy) {
            const double W = x00.w + xm0.w;
            r = (W > 0) ? __dmul_rn(inv_count(W), (x00.g + xm0.g)) : 0.0;
          } else {
            const double W = x01.w + xm1.w;
            r = (W > 0) ? __dmul_rn(inv_count(W), (xm1.g + x01.g)) : 0.0;
          }
          const double hx_cross = r;
          if (x00.w > 0) { // y-derivative, i-offset (:469-495)
            const double W = y00.w + y0m.w + y1m.w + y10.w;
            r = (W > 0) ? __dmul_rn(inv_count(W), (y00.g + y0m.g + y1m.g + y10.g)) : 0.0;
          } else if (icy) {
            const double W = y00.w + y0m.w;
            r = (W > 0) ? __dmul_rn(inv_count(W), (y00.g + y0m.g)) : 0.0;
          } else {
            const double W = y1m.w + y10.w;
            r = (W > 0) ? __dmul_rn(inv_count(W), (y1m.g + y10.g)) : 0.0;
          }
          const double2 hx2 = make_double2(x00.g, hx_cross), hy2 = make_double2(r, y00.g);
          *reinterpret_cast<double2 *>(F.h_x + s2) = hx2; // (a staggered pair starts on a 16-byte boundary)
          *reinterpret_cast<double2 *>(F.h_y + s2) = hy2;
          if (PUSH && !LOCAL_RING) {
            const int a = i - P.xs, b = j - P.ys;
            const bool W_ = a < PP.w, E_ = a >= P.xm - PP.w, S_ = b < PP.w, N_ = b >= P.ym - PP.w;
            if (W_ || E_ || S_ || N_) {
#pragma unroll
              for (int d = 0; d < 8; ++d) {
                if (PP.a[d] != nullptr && peer_strip_member(d, W_, E_, S_, N_)) {
                  const long tt = ((long)(b + P.wst + PP.dj[d]) * PP.rowc[d] + (a + P.wst + PP.di[d])) * 2;
                  *reinterpret_cast<double2 *>(PP.a[d] + tt) = hx2;
                  *reinterpret_cast<double2 *>(PP.b[d] + tt) = hy2;
                }
              }
            }
          }
        }
      }
    }
    if (P.seg_n > 0) {
      __syncthreads();
      if (threadIdx.x < 8 && cnt[threadIdx.x] != 0) atomicAdd(F.segw + seg_lo + threadIdx.x, cnt[threadIdx.x]);
      __syncthreads(); // (the counters are reused by the CTA's next tile)
    }
  }
  seg_order_epilogue(P, F);
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
// peer halo exchange: the ghost update of util/iceModelVec.cc:630-643 as direct stores into the neighbours'
// arrays (peer memory over NVLink).  blockIdx.y = strip, blockIdx.x strides over its elements.
// ---------------------------------------------------------------------------------------------
__global__ void k_halo_push(const __grid_constant__ HaloBatch B) {
  const HaloDesc &D = B.d[blockIdx.y];
  const long rowlen = (long)D.wc * D.dof;
  const long n = rowlen * D.hc;
  for (long q = (long)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (long)gridDim.x * blockDim.x) {
    const long jj = q / rowlen, e = q - jj * rowlen;
    D.dst[((D.dst_j0 + jj) * D.dst_row_cells + D.dst_i0) * D.dof + e] =
        D.src[((D.src_j0 + jj) * D.src_row_cells + D.src_i0) * D.dof + e];
  }
}
// runs after k_halo_push in stream order: every strip has landed when the counters move
__global__ void k_halo_signal(const __grid_constant__ HaloSignal S) {
  __threadfence_system();
  if (threadIdx.x < 8 && S.slot[threadIdx.x] != nullptr) {
    *reinterpret_cast<volatile unsigned long long *>(S.slot[threadIdx.x]) = S.value;
  }
  __threadfence_system();
}
// spins until all eight neighbours have delivered the strips of this phase (they never wait for this rank
// before signalling, so the wait cannot deadlock)
__global__ void k_halo_wait(const unsigned long long *slots, unsigned long long value) {
  if (threadIdx.x < 8) {
    const volatile unsigned long long *p = slots + threadIdx.x;
    while (*p < value) {
      __nanosleep(100);
    }
  }
  __threadfence_system();
}

// ---------------------------------------------------------------------------------------------
// StressBalance::compute_vertical_velocity, stressbalance/StressBalance.cc:283-424.
// One warp per column, lanes across z: the six columns it reads (u west / centre / east, v south / centre /
// north) and the one it writes are contiguous in z, so every access is coalesced; the running integral
// w[k] = w[k-1] - dz/2 (s[k] + s[k-1]), s = u_x + v_y, is a warp scan with a carry between 32-level chunks.
// HBM-bound: 24 Mz bytes per column (u, v read once -- the neighbours come from L2 -- and w written).
// ---------------------------------------------------------------------------------------------
__global__ void k_vertical_velocity(const __grid_constant__ DP P, const double *__restrict__ mask,
                                    const double *__restrict__ u, const double *__restrict__ v,
                                    const double *__restrict__ bmr, int upstream, const double *__restrict__ z,
                                    double *__restrict__ w) {
  const long col = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (col >= (long)P.xm * P.ym) {
    return; // whole warps leave together
  }
  const int i = P.xs + (int)(col % P.xm), j = P.ys + (int)(col / P.xm);
  const int Mz = P.Mz;
  const int M0 = mask_int(mask[idx2(P, i, j, P.wg)]), Me = mask_int(mask[idx2(P, i + 1, j, P.wg)]),
            Mw = mask_int(mask[idx2(P, i - 1, j, P.wg)]), Mn = mask_int(mask[idx2(P, i, j + 1, P.wg)]),
            Ms = mask_int(mask[idx2(P, i, j - 1, P.wg)]);
  const double *u_ij = u + idx2(P, i, j, P.wuv) * Mz, *u_w = u_ij - Mz, *u_e = u_ij + Mz;
  const double *v_ij = v + idx2(P, i, j, P.wuv) * Mz;
  const long rowuv = (long)(P.xm + 2 * P.wuv) * Mz;
  const double *v_s = v_ij - rowuv, *v_n = v_ij + rowuv;
  double west = 1.0, east = 1.0, south = 1.0, north = 1.0;
  if (upstream) { // :336-350, :372-386 (basal velocities decide the direction)
    const double uw = 0.5 * (u_w[0] + u_ij[0]), ue = 0.5 * (u_ij[0] + u_e[0]);
    if (uw > 0.0 && ue >= 0.0) {
      west = 1.0, east = 0.0;
    } else if (uw <= 0.0 && ue < 0.0) {
      west = 0.0, east = 1.0;
    }
    const double vs = 0.5 * (v_s[0] + v_ij[0]), vn = 0.5 * (v_ij[0] + v_n[0]);
    if (vs > 0.0 && vn >= 0.0) {
      south = 1.0, north = 0.0;
    } else if (vs <= 0.0 && vn < 0.0) {
      south = 0.0, north = 1.0;
    }
  }
  // one-sided differences at ice margins (:352-357, :388-393)
  if ((m_icy(M0) && m_ice_free(Me)) || (m_ice_free(M0) && m_icy(Me))) east = 0;
  if ((m_icy(M0) && m_ice_free(Mw)) || (m_ice_free(M0) && m_icy(Mw))) west = 0;
  if ((m_icy(M0) && m_ice_free(Mn)) || (m_ice_free(M0) && m_icy(Mn))) north = 0;
  if ((m_icy(M0) && m_ice_free(Ms)) || (m_ice_free(M0) && m_icy(Ms))) south = 0;
  const double D_x = (east + west > 0) ? 1.0 / (P.dx * (east + west)) : 0.0;
  const double D_y = (north + south > 0) ? 1.0 / (P.dy * (north + south)) : 0.0;
  double *w_ij = w + col * Mz;
  double wacc = (bmr != nullptr) ? -bmr[col] : 0.0; // w at the base (:409-413)
  double carry = 0.0;                              // s at the last level of the previous chunk
  for (int k0 = 0; k0 < Mz; k0 += 32) {
    const int k = k0 + lane;
    const bool valid = k < Mz;
    double sk = 0.0, hdz = 0.0;
    if (valid) {
      const double u_x = D_x * (west * (u_ij[k] - u_w[k]) + east * (u_e[k] - u_ij[k]));
      const double v_y = D_y * (south * (v_ij[k] - v_s[k]) + north * (v_n[k] - v_ij[k]));
      sk = u_x + v_y;
      hdz = (k > 0) ? 0.5 * (z[k] - z[k - 1]) : 0.0;
    }
    double sprev = __shfl_up_sync(FULLMASK, sk, 1);
    if (lane == 0) sprev = carry;
    double t = -hdz * (sk + sprev); // w[k] - w[k-1] (:418-422); 0 at k = 0 and past the top
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const double y = __shfl_up_sync(FULLMASK, t, d);
      t += (lane >= d) ? y : 0.0;
    }
    const double wk = wacc + t;
    if (valid) w_ij[k] = wk;
    wacc = __shfl_sync(FULLMASK, wk, 31);
    carry = __shfl_sync(FULLMASK, sk, 31);
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

int launch_gradient(const DP &P, const Fields &F, cudaStream_t s, const PeerPush *push, bool with_prep2d) {
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
    if (with_prep2d) {
      // tiles of GQ_T columns x GQ_R rows, one thread per column of a tile
      const long ntiles = (long)((P.xm + 2 * P.wg + GQ_T - 1) / GQ_T) * ((P.ym + 2 * P.wg + GQ_R - 1) / GQ_R);
      const unsigned nb = (unsigned)std::min<long>(ntiles, 148L * 4L * GQ_MINB);
      if (gradient_ring_is_local(P)) {
        k_grad_haseloff_quad<false, true><<<nb, GQ_T, 0, s>>>(P, F, PeerPush());
      } else if (push != nullptr && push->on) {
        k_grad_haseloff_quad<true, false><<<nb, GQ_T, 0, s>>>(P, F, *push);
      } else {
        k_grad_haseloff_quad<false, false><<<nb, GQ_T, 0, s>>>(P, F, PeerPush());
      }
    } else if (push != nullptr && push->on) {
      k_grad_haseloff<true><<<nblk(n1, 256), 256, 0, s>>>(P, F, *push);
    } else {
      k_grad_haseloff<false><<<nblk(n1, 256), 256, 0, s>>>(P, F, PeerPush());
    }
    return 1;
  }
}

// ---------------------------------------------------------------------------------------------
// u, v of the pieces [p0, p1) into mapped host memory.  A row of a piece is one run of elements [base, base + len) of
// which the levels k < n of every column are wanted: the threads sweep the run in windows that start on a 256-byte
// boundary of the array, element e by the thread e mod window, so that a warp's stores are whole aligned lines except
// where a column's cut begins or ends (PCIe write packets as large as the hardware makes them).  blockIdx.y strides
// over the pieces, the blockIdx.x CTAs of a piece share each of its rows.
constexpr int SP_T = 128, SP_X = 4, SP_Y = 32;
__global__ void __launch_bounds__(SP_T) k_store_pieces(const double *__restrict__ u, const double *__restrict__ v, double *__restrict__ hu,
                                                       double *__restrict__ hv, const StorePiece *__restrict__ pieces, int p0, int p1,
                                                       long row_cells, int Mz) {
  for (int p = p0 + blockIdx.y; p < p1; p += gridDim.y) {
    const StorePiece P = pieces[p];
    for (int r = P.r0; r < P.r1; ++r) {
      const long base = ((long)r * row_cells + P.c0) * Mz, end = base + (long)(P.c1 - P.c0) * Mz;
      for (long e = (base & ~31L) + (long)blockIdx.x * SP_T + threadIdx.x; e < end; e += (long)gridDim.x * SP_T) {
        if (e >= base && (int)(e - base) % Mz < P.n) { // (a row of a piece is far below 2^31 elements)
          hu[e] = u[e];
          hv[e] = v[e];
        }
      }
    }
  }
}

int launch_store_pieces(const double *u, const double *v, double *host_u, double *host_v, const StorePiece *pieces_dev, int p0,
                        int p1, long row_cells, int Mz, cudaStream_t s) {
  if (p1 <= p0) return 0;
  k_store_pieces<<<dim3(SP_X, std::min(SP_Y, p1 - p0)), SP_T, 0, s>>>(u, v, host_u, host_v, pieces_dev, p0, p1, row_cells, Mz);
  return 1;
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

int launch_halo_push(const HaloBatch &B, cudaStream_t s) {
  if (B.n <= 0) return 0;
  long nmax = 1;
  for (int q = 0; q < B.n; ++q) nmax = std::max(nmax, (long)B.d[q].wc * B.d[q].dof * B.d[q].hc);
  const unsigned bx = (unsigned)std::min<long>((nmax + 255) / 256, 64);
  k_halo_push<<<dim3(bx, (unsigned)B.n), 256, 0, s>>>(B);
  return 1;
}
int launch_halo_signal(const HaloSignal &S, cudaStream_t s) {
  k_halo_signal<<<1, 32, 0, s>>>(S);
  return 1;
}
int launch_halo_wait(const unsigned long long *slots8, unsigned long long value, cudaStream_t s) {
  k_halo_wait<<<1, 32, 0, s>>>(slots8, value);
  return 1;
}

int launch_vertical_velocity(const DP &P, const double *mask, const double *u, const double *v, const double *bmr,
                             int upstream, const double *z, double *w, cudaStream_t s) {
  const long threads = (long)P.xm * P.ym * 32;
  if (threads <= 0) return 0;
  k_vertical_velocity<<<nblk(threads, 256), 256, 0, s>>>(P, mask, u, v, bmr, upstream, z, w);
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
