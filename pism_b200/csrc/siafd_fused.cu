// siafd_fused.cu -- the fused diffusivity / flux / I / velocity kernel (sm_100a).
//
// Fuses, for one SIAFD::update (reference: juliusgarbe/pism v1.2.1, src/stressbalance/sia/SIAFD.cc):
//   compute_diffusivity            :543-770   (delta column, D, D_max, diffusivity cap, edge override)
//   compute_diffusive_flux         :772-793
//   compute_I                      :807-870
//   compute_3d_horizontal_velocity :890-948   (the ghost exchange at :946-947 stays outside)
// The reference round-trips four 3D scratch fields (delta_0/1, I_0/1) through memory and sweeps z three
// times; here delta never leaves registers and I only lives in shared memory, so HBM traffic is the
// compulsory read of the enthalpy and write of u, v (measured: profiles/).
//
// Work decomposition
//   CTA    = strip of TX "extended" columns [ca, cb) x a segment of RS extended rows; the extended patch is
//            owned + 1 ghost ring = the reference's PointsWithGhosts(1) iteration space.  The CTA marches
//            over its rows; row r + 2 of the enthalpy streams into shared memory (one cp.async.bulk per row,
//            mbarrier-tracked; 8-byte cp.async as an alternative) while row r is integrated.
//   group  = half-warp (16 lanes) working on the two staggered columns of one regular column together (ILP 2):
//              A: o = 0 (i-offset) at column c0 = ca - 1 + g     (g = 0: west halo, feeds u,v of column ca)
//              B: o = 1 (j-offset) at column c1 = ca + g
//            Lane l integrates a run of Lc = ceil((ks+1)/16) consecutive z levels serially in registers
//            (delta, the D and I trapezoids); one half-warp scan per column per row stitches the runs.
//   stage B: group g writes u, v of regular column c1 from I_e = I0[g+1], I_w = I0[g], I_n = I1[cur][g],
//            I_s = I1[prev][g] with 16 lanes across z: 128-byte coalesced stores.
//   Ice-free staggered points (thk == 0, SIAFD.cc:631-637) are flagged, never integrated and never stored;
//   enthalpy rows no active staggered point touches are not even loaded.
//
// Arithmetic deviations from the reference's glibc build, all far inside the 1e-10 bar:
// FMA contraction; exp() and the division inside the Arrhenius factor use an inlined 1-ulp exp and a
// Newton reciprocal; sums over z are taken in scan order.
#include "siafd_math.cuh"

namespace siafd {

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

// inclusive prefix sum over the 16 lanes of a half-warp
__device__ __forceinline__ double scan16(double x, int l) {
#pragma unroll
  for (int d = 1; d < 16; d <<= 1) {
    const double y = __shfl_up_sync(FULLMASK, x, d, 16);
    x += (l >= d) ? y : 0.0;
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
__global__ void __launch_bounds__(256, 2)
    k_sia_fused(const __grid_constant__ DP P, const Fields F, const int TX, const int RS, const int use_bulk,
                const int skip_rows) {
  extern __shared__ __align__(16) double sm[];
  const int tid = threadIdx.x, NT = blockDim.x;
  const int l = tid & 15, g = tid >> 4, lane = tid & 31;
  const int Mz = P.Mz;

  // ---- shared memory carve-up (offsets in doubles; every region 16-byte aligned) ----
  const int Mz2 = (Mz + 1) & ~1;
  const int slotE = ((TX + 2) * Mz + 2 + 1) & ~1; // +2: a bulk copy may start one double early / end one late
  double *z_s = sm;                               // z[k]
  double *hz_s = z_s + Mz2;                       // 0.5 * (z[k] - z[k-1])
  double *E_s = hz_s + Mz2;
  double *A_s = E_s + 3 * slotE; // age rows (only with age coupling)
  double *I0_s = A_s + (P.use_age ? 3 * slotE : 0);
  double *I1_s = I0_s + (FULL ? (((TX + 1) * Mz + 1) & ~1) : 0);
  unsigned long long *bars = (unsigned long long *)(I1_s + (FULL ? ((2 * TX * Mz + 1) & ~1) : 0));
  int *flag_s = (int *)(bars + 4); // flag_s[g]: staggered o = 0 point of group g has ice this row

  const int ca = (P.xs - 1) + blockIdx.x * TX;
  const int cb = min(ca + TX, P.xs + P.xm + 1);
  const int ncol = cb - ca;
  const int ra = (P.ys - 1) + blockIdx.y * RS;
  const int rb = min(ra + RS, P.ys + P.ym + 1);
  const bool has_west = blockIdx.x > 0;
  const int r0 = (FULL && blockIdx.y > 0) ? ra - 1 : ra; // warm-up row: I1 of the row below the segment

  for (int k = tid; k < Mz; k += NT) {
    const double zk = F.z[k];
    z_s[k] = zk;
    hz_s[k] = (k > 0) ? 0.5 * (zk - F.z[k - 1]) : 0.0;
  }
  if (use_bulk && tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_init(&bars[2], 1);
    fence_mbar_init();
  }

  // ---- which enthalpy rows are needed at all: rowflag(rho) = any thk_smooth > 0 in columns [ca-1, cb] ----
  // need(rho) = rowflag(rho-1) | rowflag(rho) | rowflag(rho+1).  Every warp evaluates the flags itself
  // (lanes = columns, one ballot), so no broadcast is needed.  Bit q of `rf` = rowflag(r + q - 1).
  const int wgx = P.xm + 2 * P.wg;
  const bool flag_lane = lane < ncol + 2;
  const double *ts_col = F.thk_smooth + (ca - 1 - (P.xs - P.wg)) + lane; // column of this lane, local row 0
  const int row_lo = P.ys - P.wg, row_hi = P.ys + P.ym + P.wg;           // valid rows [row_lo, row_hi)
  auto rowflag_load = [&](int rho) -> double {
    return (flag_lane && rho >= row_lo && rho < row_hi) ? __ldg(ts_col + (long)(rho - row_lo) * wgx) : 0.0;
  };
  unsigned rf = 0;
  if (skip_rows) {
#pragma unroll
    for (int q = 0; q < 5; ++q) {
      const double v = rowflag_load(r0 + q - 1);
      if (__ballot_sync(FULLMASK, v > 0.0)) rf |= (1u << q);
    }
  } else {
    rf = 0xffffffffu;
  }
  double ts_pref = skip_rows ? rowflag_load(r0 + 4) : 0.0; // rowflag(r + 4), consumed next iteration

  // ---- enthalpy (and age) row loader: columns [ca-1, cb] of row r -> slot ----
  const int rowcount = (ncol + 2) * Mz;
  const long NXe = P.xm + 2 * P.we;
  const long goff0 = ((long)(r0 - (P.ys - P.we)) * NXe + (ca - 1 - (P.xs - P.we))) * Mz; // row r0
  const long gstride = NXe * Mz;
  auto issue_row = [&](int it_row, int slot) { // it_row = r - r0
    const long goff = goff0 + (long)it_row * gstride;
    if (use_bulk) {
      if (tid == 0) {
        const long a0 = goff & ~1L;
        const long a1 = (goff + rowcount + 1) & ~1L;
        const unsigned bytes = (unsigned)((a1 - a0) * 8);
        mbar_expect_tx(&bars[slot], P.use_age ? 2 * bytes : bytes);
        bulk_g2s(E_s + slot * slotE, F.E + a0, bytes, &bars[slot]);
        if (P.use_age) bulk_g2s(A_s + slot * slotE, F.age + a0, bytes, &bars[slot]);
      }
    } else {
      const unsigned dst = smem_u32(E_s + slot * slotE);
      const double *src = F.E + goff;
      for (int e = tid; e < rowcount; e += NT) {
        cp_async8(dst + 8u * e, src + e);
      }
      if (P.use_age) {
        const unsigned dstA = smem_u32(A_s + slot * slotE);
        const double *srcA = F.age + goff;
        for (int e = tid; e < rowcount; e += NT) {
          cp_async8(dstA + 8u * e, srcA + e);
        }
      }
    }
  };
  // slot bookkeeping (uniform over the CTA): parity to wait for, and whether the slot was armed at all
  unsigned bar_phase = 0, slot_loaded = 0;
  __syncthreads(); // z_s, mbarrier init visible

  // prologue: rows r0 and r0 + 1
  {
    const bool n0 = (rf & 7u) != 0, n1 = (rf & 14u) != 0;
    if (n0) issue_row(0, 0);
    if (n1) issue_row(1, 1);
    if (n0) slot_loaded |= 1u;
    if (n1) slot_loaded |= 2u;
    if (!use_bulk) {
      cp_async_commit();
      cp_async_wait_all();
      __syncthreads();
    }
  }

  // ---- per-lane scalar fetch, resolved ONCE: lane l of group g reads one 2D value per row ----
  //  0,1: thk_smooth at the two ends of A      2,3: theta at the two ends of A     4,5: h_x, h_y of A
  //  6,7: thk_smooth at the two ends of B      8,9: theta at the two ends of B   10,11: h_x, h_y of B
  // 12,13: sliding u, v at (c1, r)            14,15: h_x, h_y of the east staggered point (c1, r, 0)
  const double *sp = nullptr; // address for row r0
  long sstride = 0;           // doubles per row
  int smin = ra, smax = rb;   // valid rows [smin, smax)
  {
    const int c0 = ca - 1 + g, c1 = ca + g;
    const bool vA = (g <= ncol) && (g > 0 || has_west);
    const bool vB = (g < ncol);
    const long sg = P.xm + 2 * P.wg, sst = 2L * (P.xm + 2 * P.wst), ssl = 2L * (P.xm + 2 * P.wsl);
    switch (l) {
    case 0: if (vA) { sp = F.thk_smooth + idx2(P, c0, r0, P.wg); sstride = sg; } break;
    case 1: if (vA) { sp = F.thk_smooth + idx2(P, c0 + 1, r0, P.wg); sstride = sg; } break;
    case 2: if (vA) { sp = F.theta + idx2(P, c0, r0, P.wg); sstride = sg; } break;
    case 3: if (vA) { sp = F.theta + idx2(P, c0 + 1, r0, P.wg); sstride = sg; } break;
    case 4: if (vA) { sp = F.h_x + idx2(P, c0, r0, P.wst) * 2; sstride = sst; } break;
    case 5: if (vA) { sp = F.h_y + idx2(P, c0, r0, P.wst) * 2; sstride = sst; } break;
    case 6: if (vB) { sp = F.thk_smooth + idx2(P, c1, r0, P.wg); sstride = sg; smin = r0; } break;
    case 7: if (vB) { sp = F.thk_smooth + idx2(P, c1, r0 + 1, P.wg); sstride = sg; smin = r0; } break;
    case 8: if (vB) { sp = F.theta + idx2(P, c1, r0, P.wg); sstride = sg; smin = r0; } break;
    case 9: if (vB) { sp = F.theta + idx2(P, c1, r0 + 1, P.wg); sstride = sg; smin = r0; } break;
    case 10: if (vB) { sp = F.h_x + idx2(P, c1, r0, P.wst) * 2 + 1; sstride = sst; smin = r0; } break;
    case 11: if (vB) { sp = F.h_y + idx2(P, c1, r0, P.wst) * 2 + 1; sstride = sst; smin = r0; } break;
    case 12:
    case 13:
      if (FULL && vB && c1 >= P.xs && c1 < P.xs + P.xm && F.sliding != nullptr) {
        sp = F.sliding + idx2(P, c1, r0, P.wsl) * 2 + (l - 12);
        sstride = ssl;
        smin = max(ra, P.ys);
        smax = min(rb, P.ys + P.ym);
      }
      break;
    case 14: if (FULL && vB) { sp = F.h_x + idx2(P, c1, r0, P.wst) * 2; sstride = sst; } break;
    default: if (FULL && vB) { sp = F.h_y + idx2(P, c1, r0, P.wst) * 2; sstride = sst; } break;
    }
  }
  auto fetch_scalar = [&](int r) -> double {
    return (sp != nullptr && r >= smin && r < smax) ? __ldg(sp + (long)(r - r0) * sstride) : 0.0;
  };

  double sc_next = fetch_scalar(r0);
  double prev_hxB = 0.0, prev_hyB = 0.0; // h_x, h_y of the j-offset point one row below (stage B "south")
  bool prev_actB = false;
  double dmax_local = 0.0;
  int hdc_local = 0;

  const bool own_col = (g < ncol) && (ca + g >= P.xs) && (ca + g < P.xs + P.xm); // regular column gets u, v
  const int oA = g * Mz, oE = (g + 1) * Mz; // smem offsets of E columns c0 and c0 + 1 (= c1) within a row slot
  const long uv_row = (long)(P.xm + 2 * P.wuv) * Mz;
  const long uv0 = FULL ? idx2(P, min(max(ca + g, P.xs - P.wuv), P.xs + P.xm + P.wuv - 1), P.ys - P.wuv, P.wuv) * Mz : 0;

  if (use_bulk) {
    if (slot_loaded & 1u) { mbar_wait(&bars[0], 0); bar_phase ^= 1u; }
    if (slot_loaded & 2u) { mbar_wait(&bars[1], 0); bar_phase ^= 2u; }
  }

  for (int r = r0; r < rb; ++r) {
    const int it = r - r0;
    const int s_cur = it % 3, s_nxt = (it + 1) % 3, s_pre = (it + 2) % 3;
    // prefetch row r + 2 if any staggered point will need it: rowflag(r+1) | (r+2) | (r+3) = bits 2,3,4
    const bool prefetch = (r + 2 <= rb) && ((rf & 28u) != 0);
    const bool row_active = (rf & 6u) != 0; // rowflag(r) | rowflag(r + 1): can any staggered point of row r have ice?
    if (prefetch) {
      issue_row(it + 2, s_pre);
      slot_loaded |= (1u << s_pre);
    } else {
      slot_loaded &= ~(1u << s_pre);
    }
    if (!use_bulk) cp_async_commit();
    const double sc = sc_next;
    if (r + 1 < rb) sc_next = fetch_scalar(r + 1);
    if (skip_rows) { // shift the row-flag window by one row; rowflag(r + 4) arrives, rowflag(r + 5) is requested
      const unsigned newbit = __ballot_sync(FULLMASK, ts_pref > 0.0) ? 1u : 0u;
      rf = (rf >> 1) | (newbit << 4);
      ts_pref = rowflag_load(r + 5);
    }

    // ---------------- stage A: integrate the two staggered columns of this group ----------------
    const bool validA = (g <= ncol) && (g > 0 || has_west) && (r >= ra);
    const bool validB = (g < ncol);
    const double hxA = __shfl_sync(FULLMASK, sc, 4, 16), hyA = __shfl_sync(FULLMASK, sc, 5, 16);
    const double hxB = __shfl_sync(FULLMASK, sc, 10, 16), hyB = __shfl_sync(FULLMASK, sc, 11, 16);

    bool actA = false, actB = false;
    double thkA = 0.0, thkB = 0.0;
    if (row_active) { // CTA-uniform: some thk_smooth > 0 in rows r, r + 1 of this strip
      const double tsA0 = __shfl_sync(FULLMASK, sc, 0, 16), tsA1 = __shfl_sync(FULLMASK, sc, 1, 16);
      const double tsB0 = __shfl_sync(FULLMASK, sc, 6, 16), tsB1 = __shfl_sync(FULLMASK, sc, 7, 16);
      thkA = 0.5 * (tsA0 + tsA1), thkB = 0.5 * (tsB0 + tsB1); // sia/SIAFD.cc:627-628
      actA = validA && (thkA != 0.0), actB = validB && (thkB != 0.0); // :631-637
    }
    const bool any_act = row_active && __any_sync(FULLMASK, actA || actB);

    double DA = 0.0, DB = 0.0;
    double *I0row = I0_s + oA;                       // o = 0 point of this group (index g)
    double *I1row = I1_s + ((it & 1) * TX + g) * Mz; // o = 1 point, slot by row parity

    if (any_act) { // warp-uniform
      const double thA0 = __shfl_sync(FULLMASK, sc, 2, 16), thA1 = __shfl_sync(FULLMASK, sc, 3, 16);
      const double thB0 = __shfl_sync(FULLMASK, sc, 8, 16), thB1 = __shfl_sync(FULLMASK, sc, 9, 16);
      const int ksA = actA ? k_below_height(z_s, Mz, thkA, F.err) : -1; // :639
      const int ksB = actB ? k_below_height(z_s, Mz, thkB, F.err) : -1;
      // sia/SIAFD.cc:686, :693-696
      const double alphaA = sqrt(hxA * hxA + hyA * hyA), alphaB = sqrt(hxB * hxB + hyB * hyB);
      const double thetaA = 0.5 * (thA0 + thA1), thetaB = 0.5 * (thB0 + thB1);
      const double c2A = P.e * thetaA * 2.0, c2B = P.e * thetaB * 2.0; // e_factor * theta_local * 2.0

      // Each lane integrates a RUN of Lc consecutive levels [l Lc, (l+1) Lc) of both columns, serially and
      // in registers; 16 Lc >= ks + 1.  The trapezoid that straddles two runs and the offsets of the runs
      // are added afterwards (one half-warp scan per column per row).
      int Lc = max(ksA, ksB);
      Lc = max(Lc, __shfl_xor_sync(FULLMASK, Lc, 16));
      Lc = (Lc >> 4) + 1;
      const int k0 = l * Lc;

      const int shc = use_bulk ? (int)((goff0 + (long)it * gstride) & 1) : 0;
      const int shn = use_bulk ? (int)((goff0 + (long)(it + 1) * gstride) & 1) : 0;
      const double *Ecur = E_s + s_cur * slotE + shc;
      const double *Enxt = E_s + s_nxt * slotE + shn;
      // (groups past the strip's last column read a clamped column: loads are unconditional, results masked)
      const int eA = min(g, ncol + 1) * Mz + k0, eE = min(g + 1, ncol + 1) * Mz + k0;
      const double *EaA = Ecur + eA, *Emd = Ecur + eE; // columns c0 and c0 + 1 (= c1) of row r
      const double *EbB = Enxt + eE;                   // column c1 of row r + 1
      const double *zl = z_s + k0, *hzl = hz_s + k0;
      double *I0p = I0row + k0, *I1p = I1row + k0;

      double prevA = 0.0, prevB = 0.0, firstA = 0.0, firstB = 0.0, depFA = 0.0, depFB = 0.0;
      double runA = 0.0, runB = 0.0, dpA = 0.0, dpB = 0.0;
      const double *Ac = A_s + s_cur * slotE + shc, *An = A_s + s_nxt * slotE + shn;
      // Two levels (k, k + 1) of both columns per trip: four independent flow-law chains in flight.
      // j = 0: A at k, 1: B at k, 2: A at k + 1, 3: B at k + 1.  Masked-out slots compute on clamped loads.
      for (int m = 0; m < Lc; m += 2) {
        const int k = k0 + m;
        const bool two = (m + 1 < Lc);
        const bool in0A = (k <= ksA), in0B = (k <= ksB), in1A = two && (k + 1 <= ksA), in1B = two && (k + 1 <= ksB);
        const int q0 = min(k, Mz - 1) - k0, q1 = min(k + 1, Mz - 1) - k0; // in-bounds offsets from the run start
        const double z0 = zl[q0], z1 = zl[q1], hz0 = hzl[q0], hz1 = hzl[q1];
        const double Em0 = Emd[q0], Em1 = Emd[q1]; // far end of A, near end of B
        double dep[4], pr[4], Ea[4], st[4], gsz[4], fl[4], c2[4];
        dep[0] = thkA - z0, dep[1] = thkB - z0, dep[2] = thkA - z1, dep[3] = thkB - z1;      // :641-643
        Ea[0] = 0.5 * (EaA[q0] + Em0), Ea[1] = 0.5 * (Em0 + EbB[q0]);                        // :677-684
        Ea[2] = 0.5 * (EaA[q1] + Em1), Ea[3] = 0.5 * (Em1 + EbB[q1]);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          pr[j] = fma(P.rg, dep[j], P.p_air);              // EnthalpyConverter.cc:146-152
          st[j] = ((j & 1) ? alphaB : alphaA) * pr[j];     // :688
          gsz[j] = P.grain_size;
          c2[j] = (j & 1) ? c2B : c2A;
        }
        if (P.use_age) { // :649-675 (uniform branch; off by default)
          double age[4];
          age[0] = 0.5 * (Ac[eA + q0] + Ac[eE + q0]), age[1] = 0.5 * (Ac[eE + q0] + An[eE + q0]);
          age[2] = 0.5 * (Ac[eA + q1] + Ac[eE + q1]), age[3] = 0.5 * (Ac[eE + q1] + An[eE + q1]);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (P.gs_age) gsz[j] = grain_size_vostok(age[j] * P.years_per_second);
            if (P.e_age) {
              c2[j] = (interglacial(P, P.current_time - age[j]) ? P.e_inter : P.e) * ((j & 1) ? thetaB : thetaA) * 2.0;
            }
          }
        }
        flow_lean_v<LAW, 4>(P, st, Ea, pr, gsz, fl); // :691
        const double d0A = in0A ? c2[0] * pr[0] * fl[0] : 0.0, d0B = in0B ? c2[1] * pr[1] * fl[1] : 0.0; // :696
        const double d1A = in1A ? c2[2] * pr[2] * fl[2] : 0.0, d1B = in1B ? c2[3] * pr[3] * fl[3] : 0.0;
        // trapezoids   I: 0.5 dz (delta[k-1] + delta[k])                          compute_I, :855-858
        //              D: 0.5 dz ((depth[k] + dz) delta[k-1] + depth[k] delta[k])  :701-705
        if (m == 0) { // the trapezoid ending at the first level of the run needs the lane below: deferred
          firstA = d0A, firstB = d0B, depFA = dep[0], depFB = dep[1];
        } else {
          const double dz = hz0 + hz0;
          if (in0A) {
            runA = fma(hz0, prevA + d0A, runA);
            dpA = fma(hz0, fma(dep[0] + dz, prevA, dep[0] * d0A), dpA);
          }
          if (in0B) {
            runB = fma(hz0, prevB + d0B, runB);
            dpB = fma(hz0, fma(dep[1] + dz, prevB, dep[1] * d0B), dpB);
          }
        }
        if (k == ksA) dpA = fma(0.5 * dep[0] * dep[0], d0A, dpA); // :707-708 (dz = thk - z[ks] = depth[ks])
        if (k == ksB) dpB = fma(0.5 * dep[1] * dep[1], d0B, dpB);
        if (FULL && k < Mz) { // run-local prefix; offset added below
          if (actA) I0p[m] = runA;
          if (actB) I1p[m] = runB;
        }
        prevA = d0A, prevB = d0B;
        if (two) {
          const double dz = hz1 + hz1;
          if (in1A) {
            runA = fma(hz1, d0A + d1A, runA);
            dpA = fma(hz1, fma(dep[2] + dz, d0A, dep[2] * d1A), dpA);
          }
          if (in1B) {
            runB = fma(hz1, d0B + d1B, runB);
            dpB = fma(hz1, fma(dep[3] + dz, d0B, dep[3] * d1B), dpB);
          }
          if (k + 1 == ksA) dpA = fma(0.5 * dep[2] * dep[2], d1A, dpA);
          if (k + 1 == ksB) dpB = fma(0.5 * dep[3] * dep[3], d1B, dpB);
          if (FULL && k + 1 < Mz) {
            if (actA) I0p[m + 1] = runA;
            if (actB) I1p[m + 1] = runB;
          }
          prevA = d1A, prevB = d1B;
        }
      }
      // trapezoid across the run boundary: delta of the last level of the lane below (0 above the ice)
      double leftA = __shfl_up_sync(FULLMASK, prevA, 1, 16), leftB = __shfl_up_sync(FULLMASK, prevB, 1, 16);
      double tFA = 0.0, tFB = 0.0;
      if (l > 0 && k0 < Mz) {
        const double hz0 = hz_s[k0], dz0 = hz0 + hz0;
        if (k0 <= ksA) {
          tFA = hz0 * (leftA + firstA);
          dpA = fma(hz0, fma(depFA + dz0, leftA, depFA * firstA), dpA);
        }
        if (k0 <= ksB) {
          tFB = hz0 * (leftB + firstB);
          dpB = fma(hz0, fma(depFB + dz0, leftB, depFB * firstB), dpB);
        }
      }
      if (FULL) {
        const double inclA = scan16(runA + tFA, l), inclB = scan16(runB + tFB, l);
        double offA = __shfl_up_sync(FULLMASK, inclA, 1, 16), offB = __shfl_up_sync(FULLMASK, inclB, 1, 16);
        if (l == 0) offA = offB = 0.0;
        const double addA = offA + tFA, addB = offB + tFB;
        const double topA = __shfl_sync(FULLMASK, inclA, 15, 16), topB = __shfl_sync(FULLMASK, inclB, 15, 16);
        const int mend = min(Lc, Mz - k0);
        if (l > 0) {
          for (int m = 0; m < mend; ++m) {
            if (actA) I0p[m] += addA;
            if (actB) I1p[m] += addB;
          }
        }
        // above the ice I keeps its last value (:861-863)
        for (int k = (Lc << 4) + l; k < Mz; k += 16) {
          if (actA) I0row[k] = topA;
          if (actB) I1row[k] = topB;
        }
      }
      DA = sum16(dpA);
      DB = sum16(dpB);
    }
    if (FULL && l == 0) flag_s[g] = actA ? 1 : 0;

    // D, flux, D_max (lane 0 of the group), sia/SIAFD.cc:711-731, :772-793
    if (l == 0 && r >= ra) {
      const bool edge_r = (r < 0 || r >= P.My - 1);
      if (validA && g >= 1) { // own cell (c0 >= ca)
        const int c0 = ca - 1 + g;
        double D = 0.0;
        if (actA) {
          D = (c0 < 0 || c0 >= P.Mx - 1 || edge_r) ? 0.0 : DA;
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
        double D = 0.0;
        if (actB) {
          D = (c1 < 0 || c1 >= P.Mx - 1 || edge_r) ? 0.0 : DB;
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
      if (own_col && r >= ra && r >= P.ys && r < P.ys + P.ym) {
        const double ub = __shfl_sync(0xffffu << (tid & 16), sc, 12, 16);
        const double vb = __shfl_sync(0xffffu << (tid & 16), sc, 13, 16);
        const bool actE = flag_s[g + 1] != 0;
        double *up = F.u + uv0 + (long)(r - (P.ys - P.wuv)) * uv_row + l;
        double *vp = F.v + uv0 + (long)(r - (P.ys - P.wuv)) * uv_row + l;
        if (!(actE || actA || actB || prev_actB)) {
          // no ice at any of the four staggered neighbours: I == 0, u = sliding velocity for all z (G9)
          for (int k = l; k < Mz; k += 16, up += 16, vp += 16) {
            *up = ub;
            *vp = vb;
          }
        } else {
          const double hxe = __shfl_sync(0xffffu << (tid & 16), sc, 14, 16);
          const double hye = __shfl_sync(0xffffu << (tid & 16), sc, 15, 16);
          const double *Ie = I0_s + oE + l, *Iw = I0_s + oA + l;
          const double *In = I1_s + ((it & 1) * TX + g) * Mz + l, *Is = I1_s + (((it + 1) & 1) * TX + g) * Mz + l;
          for (int k = l; k < Mz; k += 16, up += 16, vp += 16, Ie += 16, Iw += 16, In += 16, Is += 16) {
            const double ie = actE ? *Ie : 0.0, iw = actA ? *Iw : 0.0, in = actB ? *In : 0.0,
                         is = prev_actB ? *Is : 0.0;
            *up = ub - 0.25 * (ie * hxe + iw * hxA + in * hxB + is * prev_hxB);
            *vp = vb - 0.25 * (ie * hye + iw * hyA + in * hyB + is * prev_hyB);
          }
        }
      }
      prev_hxB = hxB;
      prev_hyB = hyB;
      prev_actB = actB;
    }

    if (!use_bulk) {
      cp_async_wait_all();
    } else if (prefetch) {
      mbar_wait(&bars[s_pre], (bar_phase >> s_pre) & 1u);
      bar_phase ^= (1u << s_pre);
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

// =============================================================================================
// launch
// =============================================================================================
size_t fused_smem_bytes(const DP &P, bool full, int TX) {
  const long Mz = P.Mz, Mz2 = (Mz + 1) & ~1L;
  const long slotE = ((long)(TX + 2) * Mz + 2 + 1) & ~1L;
  long d = 2 * Mz2 + 3 * slotE + (P.use_age ? 3 * slotE : 0);
  if (full) {
    d += (((long)(TX + 1) * Mz + 1) & ~1L) + (((long)2 * TX * Mz + 1) & ~1L);
  }
  return (size_t)d * 8 + 4 * 8 /* mbarriers */ + (TX + 2) * 4 /* flags */ + 16;
}

// widest strip (<= 14 columns = 256 threads) whose shared memory allows two CTAs per SM, else one
int pick_tile_x(const DP &P, bool full) {
  const int cand[4] = {14, 10, 6, 2};
  for (int q = 0; q < 4; ++q) {
    if (fused_smem_bytes(P, full, cand[q]) + 1024 <= (size_t)(227 * 1024) / 2) return cand[q];
  }
  for (int q = 0; q < 4; ++q) {
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
  k_sia_fused<LAW, FULL><<<grid, 16 * (TX + 2), smem, s>>>(P, F, TX, RS, T.use_bulk_copy, T.skip_ice_free);
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

} // namespace siafd
