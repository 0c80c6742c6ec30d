// siafd_slab.cu -- the fused diffusivity / flux / I / velocity kernel, "row slab" form (sm_100a).
//
// Fuses, for one SIAFD::update (reference: juliusgarbe/pism v1.2.1, src/stressbalance/sia/SIAFD.cc):
//   compute_diffusivity            :543-770   (delta column, D, D_max, diffusivity cap, edge override)
//   compute_diffusive_flux         :772-793
//   compute_I                      :807-870
//   compute_3d_horizontal_velocity :890-948   (with a communicator the ghost update of :946-947 too: rim rows store
//                                              into the neighbours' ghost cells, stage_b_rim / push_rim_row)
// delta never leaves registers and I only lives in shared memory: HBM sees one read of the enthalpy
// and one write of u, v.
//
// Work decomposition
//   CTA     = strip of NC "lane columns" [ca, ca + NC) of the extended patch (owned + 1 ghost ring, the
//             reference's PointsWithGhosts(1)) x a segment of RS extended rows, marched row by row.
//             Strips overlap by one column: lane column 0 of every strip but the first only provides the
//             west face (I_w) of lane column 1, so a strip owns NC - 1 columns.
//   stage A : thread (c, pt, w): lane column c, staggered point pt (0: i-offset, 1: j-offset), z range w.
//             A warp is one z range of both points of NC = 16 adjacent columns.  Shared-memory columns are Mz
//             (odd) doubles apart, so the column-strided reads of the enthalpy and writes of I are bank-
//             conflict free, and every per-column scalar (thickness, k_s, slope, theta) is per lane -- no
//             scans.  The thread integrates levels [w Lc, (w+1) Lc) of its point serially in registers;
//             Lc = ceil((ks + 1) / WZ) per column (rounded up to whole trips of the level loop).  The WZ ranges of a
//             column exchange {last delta, first delta, I sum, D sum} through shared memory once (one barrier); every
//             thread then evaluates the trapezoids across the range boundaries of its column itself.
//   stage B : thread (q, li): column q, 2 WZ lanes li across z: u, v of the regular column from
//             I_e, I_w (I0 of lane columns q, q - 1), I_n, I_s (I1 of this and the previous row).
//   One mbarrier-tracked group of cp.async.bulk copies per row (six copies, issued by six lanes of warp 0 from
//   descriptors in shared memory) brings the enthalpy row AND the row's 2D scalars (thk_smooth, theta, h_x, h_y) into
//   shared memory, one row ahead of use into the slot the previous row just vacated; rows no staggered point needs are
//   never loaded, and rows of a strip without ice skip stage A altogether (u, v = sliding velocity, SIAFD.cc:631-637).
//   Three barriers per row where there is ice; none of them in rows without.
//
// Arithmetic deviations from the reference's glibc build, all far inside the 1e-10 bar:
// FMA contraction; exp() and the division inside the Arrhenius factor use an inlined table-based 2^(y/16) (~2 ulp) and
// a cubic-step reciprocal (2^-60); sums over z are taken per z range and then added in a fixed order.
#include "siafd_math.cuh"

#include <atomic>
#include <type_traits>

#ifndef SLAB_LZ
#define SLAB_LZ 16 // lanes across z in stage B
#endif
#ifndef SLAB_MINB
#define SLAB_MINB 3 // CTAs per SM the register allocation aims at (128-thread CTAs)
#endif
#ifndef SLAB_NL
#define SLAB_NL 3 // levels per trip of the Arrhenius loop
#endif

namespace siafd {

// IceGrid::kBelowHeight (util/IceGrid.cc:427-440; GSL bsearch: largest k in [0, Mz-2] with z[k] <= height).
// zz[k] = {z[k], 0.5 (z[k] - z[k-1])}.  inv_dz > 0 (equally spaced levels): first guess, then corrected
// against the table, so the result is the reference's for any rounding of the guess.
__device__ __forceinline__ int k_below_height2(const double2 *zz, int Mz, double height, double inv_dz, unsigned *err) {
  if (height < 0.0 - 1.0e-6) {
    atomicOr(err, EB_BELOW);
    return 0;
  }
  if (height > zz[Mz - 1].x + 1.0e-6) {
    atomicOr(err, EB_ABOVE);
    return 0;
  }
  if (inv_dz > 0.0) {
    // the guess is nearly always the answer: both neighbours are read at once (one shared-memory latency instead of
    // two); otherwise the exact search from there
    int k = min(max((int)(height * inv_dz), 0), Mz - 2);
    const double za = zz[k].x, zb = zz[k + 1].x;
    if (za <= height && height < zb) return k;
    while (k < Mz - 2 && zz[k + 1].x <= height) ++k;
    while (k > 0 && zz[k].x > height) --k;
    return k;
  }
  int ilo = 0, ihi = Mz - 1;
  while (ihi > ilo + 1) {
    const int m = (ihi + ilo) >> 1;
    if (zz[m].x > height) {
      ihi = m;
    } else {
      ilo = m;
    }
  }
  return ilo;
}

struct SlabArgs {
  int RS;        // extended rows per CTA
  int use_bulk;  // 1: cp.async.bulk rows (column stride Mz); 0: 8-byte cp.async (column stride Mz | 1)
  int skip_rows; // 1: do not load enthalpy rows no staggered point needs
  long nE;       // doubles in the enthalpy array (bulk copies are clamped to it)
  long n2;       // doubles in a geometry-width 2D array
  double inv_dz; // (Mz - 1) / Lz if the levels are equally spaced, else 0
  int seg0;      // first row segment of this launch (a launch may cover a band of segments)
  int nseg;      // segments in this launch
  const int *order; // segment to take for each blockIdx.y (heaviest first), or NULL: seg0 + blockIdx.y
};

// one bulk copy of the row loader: source array, offset of the CTA's first row (doubles), row stride, array size,
// doubles per row; destination in slot 0 (shared-memory address) and the distance between the two slots (bytes)
struct RowCopy {
  const double *base;
  long off0, stride, lim;
  int cnt, dslot;
  unsigned dst, pad;
};

// per-slot staging area of a row's 2D scalars (offsets in doubles)
enum : int { AUX_TS = 0, AUX_TH = 18, AUX_HX = 36, AUX_HY = 68, AUX_N = 100 };

// u, v of one regular column on every level, u = u_b - 0.25 (I_e h_x_e + I_w h_x_w + I_n h_x_n + I_s h_x_s) and the same
// for v (sia/SIAFD.cc:904-943): LZ lanes across z, this lane takes levels lz, lz + LZ, ...  c? = {-0.25 h_x, -0.25 h_y} of
// the four staggered points around the column, I? their integrals at this lane's first level (the west point is one
// shared-memory column, S doubles, before the east one).  DUAL: the values also go to up2 / vp2 (a neighbour's ghost
// column).  Four levels per trip, then ONE predicated block for the up to three levels that are left (a loop over single
// levels cost 2.3 times as many instructions per level, profiles/).
template <int LZ, bool DUAL>
__device__ __forceinline__ void uv_column(const int Mz, const int lz, const int S, const double ub, const double vb,
                                          const double2 ce, const double2 cw, const double2 cn, const double2 cs,
                                          const double *Ie, const double *In, const double *Is, double *up, double *vp,
                                          double *up2, double *vp2) {
  const double hxe = ce.x, hye = ce.y, hxw = cw.x, hyw = cw.y, hxn = cn.x, hyn = cn.y, hxs = cs.x, hys = cs.y;
  int k = lz;
  for (; k + 3 * LZ < Mz; k += 4 * LZ, up += 4 * LZ, vp += 4 * LZ, Ie += 4 * LZ, In += 4 * LZ, Is += 4 * LZ) {
    double ie[4], iw[4], in[4], is[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) ie[j] = Ie[j * LZ], iw[j] = (Ie - S)[j * LZ], in[j] = In[j * LZ], is[j] = Is[j * LZ];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const double uu = fma(is[j], hxs, fma(in[j], hxn, fma(iw[j], hxw, fma(ie[j], hxe, ub))));
      const double vv = fma(is[j], hys, fma(in[j], hyn, fma(iw[j], hyw, fma(ie[j], hye, vb))));
      up[j * LZ] = uu;
      vp[j * LZ] = vv;
      if (DUAL) up2[j * LZ] = uu, vp2[j * LZ] = vv;
    }
    if (DUAL) up2 += 4 * LZ, vp2 += 4 * LZ;
  }
  if (k < Mz) {
    double ie[3], iw[3], in[3], is[3];
    bool ok[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      ok[j] = k + j * LZ < Mz;
      const int o = ok[j] ? j * LZ : 0; // (a level past the top re-reads the first one; its result is not stored)
      ie[j] = Ie[o], iw[j] = (Ie - S)[o], in[j] = In[o], is[j] = Is[o];
    }
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const double uu = fma(is[j], hxs, fma(in[j], hxn, fma(iw[j], hxw, fma(ie[j], hxe, ub))));
      const double vv = fma(is[j], hys, fma(in[j], hyn, fma(iw[j], hyw, fma(ie[j], hye, vb))));
      if (ok[j]) {
        up[j * LZ] = uu;
        vp[j * LZ] = vv;
        if (DUAL) up2[j * LZ] = uu, vp2[j * LZ] = vv;
      }
    }
  }
}

// Stage B for one row of a strip on the rim of the patch: u, v of the regular columns (sia/SIAFD.cc:904-943) stored into
// this rank's array AND, from the registers, into the ghost cell of the neighbour that faces the column -- the fused
// ghost update of SIAFD.cc:946-947 (peer memory over NVLink, or this rank's own array: the periodic wrap).  The few
// cells in the corners of the patch face three neighbours; push_rim_row serves the other two.
struct RimTargets { // per pass: where the column's values also go (nullptr: nowhere), offset by the lane's first level
  double *u2[4], *v2[4];
};
template <int NC, int NPASS, int LB, int LZ>
__device__ __noinline__ void stage_b_rim(const int Mz, const int i_lo, const int i_hi, const long uv_row0, const RimTargets T,
                                         double *u, double *v, const double *I0_s, const double *I1_s, const double *cf, int S,
                                         int s_cur, bool south, bool any_valid, int ca, int qg, int lz, int lane, double svx,
                                         double svy) {
  // i_lo .. i_hi: the owned columns; uv_row0: offset (in doubles) of column i_lo of this row in the u / v arrays
  const int s_nxt = s_cur ^ 1;
  for (int p = 0; p < NPASS; ++p) {
    const int qp = qg + p, i_p = ca + qp;
    const int srcl = (lane & ~(LZ - 1)) + p * LB; // a lane whose own column (tid / LB) is qp
    const double ub = __shfl_sync(FULLMASK, svx, srcl), vb = __shfl_sync(FULLMASK, svy, srcl);
    if (qp >= 1 && i_p >= i_lo && i_p <= i_hi) {
      const long uvo = uv_row0 + (long)(i_p - i_lo) * Mz + lz;
      double *up = u + uvo, *vp = v + uvo;
      double *up2 = T.u2[p], *vp2 = T.v2[p];
      const bool dual = up2 != nullptr;
      if (!any_valid) {
        for (int k = lz; k < Mz; k += LZ, up += LZ, vp += LZ) {
          *up = ub;
          *vp = vb;
          if (dual) *up2 = ub, *vp2 = vb, up2 += LZ, vp2 += LZ;
        }
      } else {
        // (the same routine as the in-line stage B of the kernel: the values are bit-identical)
        const double2 *c2 = reinterpret_cast<const double2 *>(cf);
        const double2 ce = c2[(s_cur * NC + qp) * 2], cw = c2[(s_cur * NC + qp - 1) * 2], cn = c2[(s_cur * NC + qp) * 2 + 1];
        const double2 cs = south ? c2[(s_nxt * NC + qp) * 2 + 1] : make_double2(0.0, 0.0);
        const double *Ie = I0_s + qp * S + lz;
        const double *In = I1_s + (s_cur * NC + qp) * S + lz, *Is = I1_s + (s_nxt * NC + qp) * S + lz;
        if (dual) {
          uv_column<LZ, true>(Mz, lz, S, ub, vb, ce, cw, cn, cs, Ie, In, Is, up, vp, up2, vp2);
        } else {
          uv_column<LZ, false>(Mz, lz, S, ub, vb, ce, cw, cn, cs, Ie, In, Is, up, vp, nullptr, nullptr);
        }
      }
    }
  }
}

// Fused ghost update of u, v (SIAFD.cc:946-947), the part stage B does not do itself: a cell in a corner of the patch
// faces three neighbours; stage B stores it into the first one's ghost cell, and here every thread copies the values
// it has just stored (same column / level mapping as stage B) into the other two (peer memory over NVLink, or this
// rank's own array: the periodic wrap).  Four cells per patch.
template <int NPASS, int LB, int LZ>
__device__ __noinline__ void push_rim_row(const DP &P, const PeerPush &PP, const double *u, const double *v, int r, int ca,
                                          int qg, int lz) {
  const int Mz = P.Mz;
  const int b = r - P.ys;
  const bool S_ = b < PP.w, N_ = b >= P.ym - PP.w;
  for (int p = 0; p < NPASS; ++p) {
    const int qp = qg + p, i_p = ca + qp;
    if (qp >= 1 && i_p >= P.xs && i_p < P.xs + P.xm) {
      const int a = i_p - P.xs;
      const bool W_ = a < PP.w, E_ = a >= P.xm - PP.w;
      if (W_ || E_ || S_ || N_) {
        const long own = ((long)(b + P.wuv) * (P.xm + 2 * P.wuv) + (a + P.wuv)) * Mz;
        // all loads first (eight levels per lane and trip), then the stores: the targets may alias the sources as far
        // as the compiler knows, and one load per store would pay the L2 latency once per level
        for (int kb = lz; kb < Mz; kb += 8 * LZ) {
          double uu[8], vv[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int k = min(kb + j * LZ, Mz - 1);
            uu[j] = __ldcg(u + own + k), vv[j] = __ldcg(v + own + k);
          }
          bool first = true; // (stage B itself served the first neighbour that faces the cell)
          for (int d = 0; d < 8; ++d) {
            if (PP.a[d] != nullptr && peer_strip_member(d, W_, E_, S_, N_)) {
              if (first) {
                first = false;
                continue;
              }
              const long t = ((long)(b + P.wuv + PP.dj[d]) * PP.rowc[d] + (a + P.wuv + PP.di[d])) * Mz;
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const int k = kb + j * LZ;
                if (k < Mz) PP.a[d][t + k] = uu[j], PP.b[d][t + k] = vv[j];
              }
            }
          }
        }
      }
    }
  }
}

template <int LAW, bool FULL, int NC, int WZ, bool BULK>
__global__ void __launch_bounds__(2 * NC *WZ, (2 * NC * WZ <= 256) ? SLAB_MINB : 1)
    k_sia_slab(const __grid_constant__ DP P, const Fields F, const SlabArgs A, const __grid_constant__ PeerPush PP) {
  static_assert(NC <= 16, "AUX_* offsets and the partner shuffle assume at most 16 lane columns");
  extern __shared__ __align__(16) double sm[];
  constexpr int NT = 2 * NC * WZ;
  constexpr int OWN = NC - 1;
  constexpr int LB = NT / NC; // stage B lanes per column
  const int tid = threadIdx.x;
  const int c = tid % NC, pt = (tid / NC) & 1, w = tid / (2 * NC); // stage A role
  const int lane = tid & 31;
  const int Mz = P.Mz;
  const int S = BULK ? Mz : (Mz | 1); // shared-memory column stride

  // ---- shared memory carve-up (offsets in doubles; every region 16-byte aligned) ----
  const int slotE = ((NC + 1) * S + 2 + 1) & ~1; // +2: a bulk copy may start one double early / end one late
  const int colI = (NC * S + 1) & ~1;
  double2 *zz = (double2 *)sm;
  double *E_s = sm + 2 * Mz;
  double *A_s = E_s + 2 * slotE; // age rows (only with age coupling)
  double *aux = A_s + (P.use_age ? 2 * slotE : 0);
  double *I0_s = aux + 2 * AUX_N;
  double *I1_s = I0_s + (FULL ? colI : 0);
  double *sX = I1_s + (FULL ? 2 * colI : 0); // [2][WZ][NC][4] per range: last delta, first delta, I sum, D sum
  double *cf = sX + 2 * WZ * NC * 4;             // [2][NC][4]  -0.25 h_x, -0.25 h_y of the o = 0 and o = 1 points, by row parity
  double *tab16 = cf + 2 * NC * 4;           // 2^(j/16), for exp_tab
  double2 *selAQ = (double2 *)(tab16 + 16);  // {ln A, Q / R} 16 / ln2 of the cold [0] and the warm [1] Paterson-Budd branch
  unsigned long long *bars = (unsigned long long *)(tab16 + 20);
  RowCopy *cpt = (RowCopy *)(tab16 + 22); // [6]

  // strips in the order last, 0, 1, ...: the two strips on the west / east rim of the patch (which also store into the
  // neighbours' ghost cells) are the first of their grid row to start, never its tail
  const int strip = (blockIdx.x == 0) ? (int)gridDim.x - 1 : (int)blockIdx.x - 1;
  const int ca = (P.xs - 1) + strip * OWN;
  const int ilast = P.xs + P.xm; // last extended column
  const int i_c = ca + c;
  const bool col_ok = i_c <= ilast;
  const int iv = min(i_c, ilast);
  const bool own_c = col_ok && (c >= 1 || strip == 0); // this strip writes D, Q of the column
  const int ncolE = min(NC + 1, ilast + 2 - ca);            // enthalpy / thk_smooth columns ca .. ca + ncolE - 1
  const int ncolS = ncolE - 1;                              // valid lane columns
  const int seg = (A.order != nullptr) ? A.order[blockIdx.y] : (int)blockIdx.y + A.seg0;
  const int ra = (P.ys - 1) + seg * A.RS;
  const int rb = min(ra + A.RS, P.ys + P.ym + 1);
  const int r0 = (FULL && seg > 0) ? ra - 1 : ra; // warm-up row: I1 of the row below the segment

  for (int k = tid; k < Mz; k += NT) {
    const double zk = F.z[k];
    zz[k] = make_double2(zk, (k > 0) ? 0.5 * (zk - F.z[k - 1]) : 0.0);
  }
  if (tid < 16) tab16[tid] = EXPT[tid];
  if (tid == 16) selAQ[0] = make_double2(P.lnA2_cold, P.QoR2_cold), selAQ[1] = make_double2(P.lnA2_warm, P.QoR2_warm);
  if (BULK && tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    fence_mbar_init();
  }

  // ---- which rows are needed at all: rowts(rho) = any thk_smooth > 0 in columns [ca, ca + NC] ----
  // need(rho) = rowts(rho-1) | rowts(rho) | rowts(rho+1).  All flags of the CTA's rows are computed up front, by
  // every warp for itself (lane = row, one ballot per 32 rows), so that the march over rows without ice never
  // waits on a load: bit j of the 96-bit word fw2:fw1:fw0 = rowts(r0 - 1 + j).
  const int wgx = P.xm + 2 * P.wg;
  const long cb2 = ca - (P.xs - P.wg);                         // local column of lane column 0 in a geometry array
  const int row_lo = P.ys - P.wg, row_hi = P.ys + P.ym + P.wg; // valid rows [row_lo, row_hi)
  unsigned fw0 = 0xffffffffu, fw1 = 0xffffffffu, fw2 = 0xffffffffu;
  if (A.skip_rows) {
    unsigned fw[3];
#pragma unroll
    for (int b = 0; b < 3; ++b) {
      const int rho = r0 - 1 + 32 * b + lane;
      double m = 0.0;
      if (rho >= row_lo && rho < row_hi && rho <= rb + 3) {
        const double *tsr = F.thk_smooth + (long)(rho - row_lo) * wgx + cb2;
#ifndef SLAB_DIAG_NOFLAGS // (diagnostic builds: no ice anywhere, nothing read)
        for (int e = 0; e < ncolE; ++e) m = fmax(m, __ldg(tsr + e));
#endif
      }
      fw[b] = __ballot_sync(FULLMASK, m > 0.0);
    }
    fw0 = fw[0], fw1 = fw[1], fw2 = fw[2];
  }
  // bits 0..4 of the result = rowts(r - 1 .. r + 3) for row r = r0 + it
  auto row_flags = [&](int it) -> unsigned {
    const unsigned lo = it < 32 ? fw0 : (it < 64 ? fw1 : fw2), hi = it < 32 ? fw1 : (it < 64 ? fw2 : 0u);
    return __funnelshift_r(lo, hi, it & 31) & 31u;
  };
  unsigned rf = row_flags(0); // bits 0..4 = rowts(r - 1 .. r + 3) of the current row (prologue: row r0)

  // ---- row loader: enthalpy (and age) columns [ca, ca + ncolE) and the 2D scalars of row r -> slot ----
  const int rowcount = ncolE * Mz;
  const long NXe = P.xm + 2 * P.we;
  const long goff0 = ((long)(r0 - (P.ys - P.we)) * NXe + (ca - (P.xs - P.we))) * Mz; // row r0
  const long gstride = NXe * Mz;
  const long g2off0 = (long)(r0 - row_lo) * wgx + cb2; // thk_smooth / theta, row r0
  const long sst = 2L * (P.xm + 2 * P.wst);
  const long gsoff0 = idx2(P, ca, r0, P.wst) * 2; // h_x / h_y, row r0 (even: always 16-byte aligned)
  // Bulk mode: one copy per lane of warp 0 (lane 0 enthalpy, 1 age, 2 thk_smooth, 3 theta, 4 h_x, 5 h_y), all issued by
  // one instruction stream: source offset of row r0 (in doubles), row stride, doubles per row, array size (a copy is
  // clamped to it), destination in slot 0 and the distance between the two slots.  A row starts on any 8-byte boundary
  // and a bulk copy on a 16-byte one: the copy starts one double early / ends one late where it has to, and the
  // readers add the parity of the row's offset (shc, sh2c below).
  // (the six descriptors live in shared memory: registers are what limits the CTAs per SM)
  if (BULK && tid < 6) {
    RowCopy d;
    d.base = nullptr, d.dst = 0, d.off0 = 0, d.stride = 0, d.lim = 0, d.cnt = 0, d.dslot = 0;
    switch (tid) {
    case 0: d.base = F.E, d.dst = smem_u32(E_s), d.off0 = goff0, d.stride = gstride, d.lim = A.nE, d.cnt = rowcount, d.dslot = slotE * 8; break;
    case 1: if (P.use_age) d.base = F.age, d.dst = smem_u32(A_s), d.off0 = goff0, d.stride = gstride, d.lim = A.nE, d.cnt = rowcount, d.dslot = slotE * 8; break;
    case 2: d.base = F.thk_smooth, d.dst = smem_u32(aux + AUX_TS), d.off0 = g2off0, d.stride = wgx, d.lim = A.n2, d.cnt = ncolE, d.dslot = AUX_N * 8; break;
    case 3: d.base = F.theta, d.dst = smem_u32(aux + AUX_TH), d.off0 = g2off0, d.stride = wgx, d.lim = A.n2, d.cnt = ncolE, d.dslot = AUX_N * 8; break;
    case 4: d.base = F.h_x, d.dst = smem_u32(aux + AUX_HX), d.off0 = gsoff0, d.stride = sst, d.lim = 1L << 60, d.cnt = 2 * ncolS, d.dslot = AUX_N * 8; break;
    default: d.base = F.h_y, d.dst = smem_u32(aux + AUX_HY), d.off0 = gsoff0, d.stride = sst, d.lim = 1L << 60, d.cnt = 2 * ncolS, d.dslot = AUX_N * 8; break;
    }
    cpt[tid] = d;
  }
  auto issue_row = [&](int it_row, int slot) {    // it_row = r - r0
    const long goff = goff0 + (long)it_row * gstride;
    const long g2 = g2off0 + (long)it_row * wgx;
    const long gs = gsoff0 + (long)it_row * sst;
    const bool stag_row = (r0 + it_row) <= P.ys + P.ym; // row rb is only ever a "north" row
    double *ax = aux + slot * AUX_N;
    if (BULK) {
      if (tid < 32) { // (warp 0, converged)
        unsigned bytes = 0, dst = 0;
        const double *src = nullptr;
        if (lane < 6 && (lane < 4 || stag_row)) {
          const RowCopy d = cpt[lane];
          if (d.cnt > 0) {
            const long g = d.off0 + (long)it_row * d.stride;
            const long a0 = g & ~1L;
            long a1 = (g + d.cnt + 1) & ~1L;
            dst = d.dst + (unsigned)(slot * d.dslot);
            if (a1 > d.lim) { // last row of the array and an odd end: fetch the last double separately
              a1 -= 2;
              const double last = d.base[g + d.cnt - 1];
              asm volatile("st.shared.f64 [%0], %1;" ::"r"(dst + 8u * (unsigned)((g - a0) + d.cnt - 1)), "d"(last) : "memory");
            }
            bytes = (unsigned)((a1 - a0) * 8);
            src = d.base + a0;
          }
        }
        const unsigned total = __reduce_add_sync(FULLMASK, bytes);
        if (lane == 0) mbar_expect_tx(&bars[slot], total);
        __syncwarp();
        if (bytes) bulk_g2s_u32(dst, src, bytes, &bars[slot]);
      }
    } else {
      const unsigned dst = smem_u32(E_s + slot * slotE);
      const double *src = F.E + goff;
      for (int e = tid; e < rowcount; e += NT) {
        cp_async8(dst + 8u * ((e / Mz) * S + e % Mz), src + e);
      }
      if (P.use_age) {
        const unsigned dstA = smem_u32(A_s + slot * slotE);
        const double *srcA = F.age + goff;
        for (int e = tid; e < rowcount; e += NT) {
          cp_async8(dstA + 8u * ((e / Mz) * S + e % Mz), srcA + e);
        }
      }
      const unsigned dx = smem_u32(ax);
      if (tid < ncolE) {
        cp_async8(dx + 8u * (AUX_TS + tid), F.thk_smooth + g2 + tid);
        cp_async8(dx + 8u * (AUX_TH + tid), F.theta + g2 + tid);
      }
      if (stag_row && tid < 2 * ncolS) {
        cp_async8(dx + 8u * (AUX_HX + tid), F.h_x + gs + tid);
        cp_async8(dx + 8u * (AUX_HY + tid), F.h_y + gs + tid);
      }
      cp_async_commit();
    }
  };
  unsigned bar_phase = 0, pending = 0; // CTA-uniform: parity to wait for / slot has a copy in flight
  __syncthreads();                     // zz, mbarrier init visible

  // prologue: rows r0 and r0 + 1
  if (rf & 6u) { // row r0 is only read by stage A of row r0
    issue_row(0, 0);
    pending |= 1u;
  }
  if (rf & 14u) {
    issue_row(1, 1);
    pending |= 2u;
  }

  double *DQ_D = F.D + idx2(P, iv, r0, P.wst) * 2 + pt;
  double *DQ_Q = F.Q + idx2(P, iv, r0, P.wst) * 2 + pt;

  // stage B role: column q (lane column index), li across z
  const int q = tid / LB, li = tid % LB;
  const int i_q = ca + q;
  const bool uv_col = FULL && q >= 1 && i_q >= P.xs && i_q < P.xs + P.xm;
  const long uv_row = (long)(P.xm + 2 * P.wuv) * Mz;
  // Stage B stores run 16 lanes across z (128-byte pieces of a column): a 16-lane group takes NPASS = 16 / LB
  // adjacent columns one after the other (LB >= 16: one column per LB lanes)
  constexpr int LZW = SLAB_LZ, LZ = (LB >= LZW) ? LB : LZW, NPASS = (LB >= LZW) ? 1 : LZW / LB;
  const int qg = (tid / LZ) * NPASS, lz = tid % LZ;
  // stage B's stores: offset of (row, column ca + qg, level lz) in u / v, advanced row by row; bit p of pass_ok: column
  // ca + qg + p is an owned column of this strip
  long uvo_run = ((long)(max(ra, P.ys) - (P.ys - P.wuv)) * (P.xm + 2 * P.wuv) + (ca + qg - P.xs + P.wuv)) * Mz + lz;
  unsigned pass_ok = 0;
#pragma unroll
  for (int p = 0; p < NPASS; ++p) {
    const int qp = qg + p, i_p = ca + qp;
    if (qp >= 1 && i_p >= P.xs && i_p < P.xs + P.xm) pass_ok |= 1u << p;
  }
  const double *sl_p =
      (FULL && uv_col && F.sliding != nullptr) ? F.sliding + idx2(P, i_q, P.ys - P.wsl, P.wsl) * 2 : nullptr;
  const long ssl = 2L * (P.xm + 2 * P.wsl);

  // Sliding velocity of stage B's column.  Which rows of the strip have a sliding velocity other than zero at all is
  // found up front, by every warp for itself like the ice flags (lane = row, one ballot per 32 rows; bit j of
  // sw2:sw1:sw0 = row r0 - 1 + j): the march over rows without ice and without sliding -- the write-only regime, where
  // the memory system is saturated with stores and a load takes many microseconds -- then never waits on one, and a
  // row that does slide fetches its values at the top of the row, long before stage B uses them.
  const int rbase = max(ra, P.ys), rend = min(rb, P.ys + P.ym); // stage B rows [rbase, rend)
  unsigned sw0 = 0u, sw1 = 0u, sw2 = 0u;
  if (FULL && F.sliding != nullptr) {
    const int i_first = max(ca + 1, P.xs), i_last = min(ca + NC - 1, P.xs + P.xm - 1);
    unsigned sw[3];
#pragma unroll
    for (int b = 0; b < 3; ++b) {
      const int rho = r0 - 1 + 32 * b + lane;
      bool nz = false;
      if (rho >= rbase && rho < rend) {
        const double2 *sr = reinterpret_cast<const double2 *>(F.sliding + idx2(P, i_first, rho, P.wsl) * 2);
        for (int e = 0; e <= i_last - i_first; ++e) {
          const double2 sl = __ldg(sr + e);
          nz |= !(sl.x == 0.0 && sl.y == 0.0);
        }
      }
      sw[b] = __ballot_sync(FULLMASK, nz);
    }
    sw0 = sw[0], sw1 = sw[1], sw2 = sw[2];
  }

  // fused ghost update of u, v (SIAFD.cc:946-947): does this strip hold owned columns within PP.w of the west / east
  // edge of the patch (CTA-uniform)
  bool strip_rim = false;
  if (FULL && PP.on) {
    const int i_first = max(ca + 1, P.xs), i_last = min(ca + NC - 1, P.xs + P.xm - 1);
    strip_rim = (i_first < P.xs + PP.w) || (i_last >= P.xs + P.xm - PP.w);
  }

  double dmax_local = 0.0;
  int hdc_local = 0;
  // logical state of the I arrays (CTA-uniform): bit set = materialised in shared memory, else "all zero".
  // bit 0, 1: I1 slots; bit 2: I0
  unsigned ivalid = 0;
  // the write-only regime (no ice in the strip, zero sliding velocity): I0_s holds zeros that the bulk-copy engine
  // stores as whole rows of u and v; zero_inflight = such a store may still be reading I0_s
  bool zero_ready = false, zero_inflight = false;

  for (int r = r0; r < rb; ++r, DQ_D += sst, DQ_Q += sst) {
    const int it = r - r0;
    rf = row_flags(it);
    const int s_cur = it & 1, s_nxt = s_cur ^ 1;
    const bool row_active = (rf & 6u) != 0;                    // rowts(r) | rowts(r + 1)
    const bool prefetch = (r + 2 <= rb) && ((rf & 28u) != 0); // rowts(r+1) | (r+2) | (r+3)
    const long ro = (long)it; // row offset from r0
    double2 sv = make_double2(0.0, 0.0);
    bool slide = false; // CTA-uniform: some column of the strip slides in this row
    if (FULL && r >= rbase && r < rend) {
      const unsigned wsel = (it + 1 < 32) ? sw0 : ((it + 1 < 64) ? sw1 : sw2);
      slide = ((wsel >> ((it + 1) & 31)) & 1u) != 0u;
      if (slide && sl_p != nullptr) sv = __ldg(reinterpret_cast<const double2 *>(sl_p + (long)(r - (P.ys - P.wsl)) * ssl));
    }

    double Dsum = 0.0, hx = 0.0, hy = 0.0;
    bool act = false;
    if (row_active) {
      if (FULL && zero_inflight) { // CTA-uniform: stage A is about to overwrite the zeros a bulk store may still read
        if (tid == 0) bulk_wait_read0();
        __syncthreads();
        zero_inflight = false;
      }
      zero_ready = false;
      // ---------------- stage A: integrate z range w of staggered point pt of lane column c ------------
      if (pending) { // CTA-uniform
        if (BULK) {
          if (pending & 1u) { mbar_wait(&bars[0], bar_phase & 1u); bar_phase ^= 1u; }
          if (pending & 2u) { mbar_wait(&bars[1], (bar_phase >> 1) & 1u); bar_phase ^= 2u; }
        } else {
          cp_async_wait_all();
          __syncthreads();
        }
        pending = 0;
      }
      const int cc = min(c, ncolS - 1); // (lanes past the patch read a clamped column; results masked)
      const double *axc = aux + s_cur * AUX_N, *axn = aux + s_nxt * AUX_N;
      const int sh2c = BULK ? (int)((g2off0 + ro * wgx) & 1) : 0;
      const int sh2n = BULK ? (int)((g2off0 + (ro + 1) * wgx) & 1) : 0;
      const double tsC = axc[AUX_TS + sh2c + cc], thC = axc[AUX_TH + sh2c + cc];
      const double ts2 = pt ? axn[AUX_TS + sh2n + cc] : axc[AUX_TS + sh2c + cc + 1];
      const double th2 = pt ? axn[AUX_TH + sh2n + cc] : axc[AUX_TH + sh2c + cc + 1];
      hx = axc[AUX_HX + 2 * cc + pt], hy = axc[AUX_HY + 2 * cc + pt];
      const double thk = 0.5 * (tsC + ts2); // sia/SIAFD.cc:627-628
      const bool ice = col_ok && (thk != 0.0); // :631-637
      act = ice && (pt == 1 || r >= ra);       // (the warm-up row only needs the j-offset point)
      const int ks0 = ice ? k_below_height2(zz, Mz, thk, A.inv_dz, F.err) : -1; // :639
      const int ks = act ? ks0 : -1;
      // levels per z range of this column: a function of the column alone, never of the CTA tiling, so that
      // the order of every sum -- and with it every bit of the result -- is independent of the decomposition
      // (a multiple of the levels per trip of the Arrhenius loop, so that only the range the surface cuts has a short trip)
      const int Lc = SLAB_NL * ((ks0 + WZ * SLAB_NL) / (WZ * SLAB_NL));
      const int k0 = w * Lc;
      const int ke = min(k0 + Lc - 1, ks); // last level of this thread (empty range: ke < k0)
      // sia/SIAFD.cc:686, :693-696
      const double alpha = sqrt(hx * hx + hy * hy);
      const double theta = 0.5 * (thC + th2);
      const double c2c = P.e * theta * 2.0; // e_factor * theta_local * 2.0

      const int shc = BULK ? (int)((goff0 + ro * gstride) & 1) : 0;
      const int shn = BULK ? (int)((goff0 + (ro + 1) * gstride) & 1) : 0;
      const int o1 = s_cur * slotE + shc + cc * S;                                        // column (i, r)
      const int o2 = pt ? (s_nxt * slotE + shn + cc * S) : (s_cur * slotE + shc + (cc + 1) * S); // (i, r+1) or (i+1, r)
      const double *E1 = E_s + o1, *E2 = E_s + o2;
      double *Ic = pt ? (I1_s + (s_cur * NC + c) * S) : (I0_s + c * S);

      double prev = 0.0, first = 0.0, run = 0.0, dp = 0.0;
      constexpr bool ARRH = (LAW == LAW_ARR || LAW == LAW_ARRWARM || LAW == LAW_PB || LAW == LAW_GPBLD);
      if (ARRH && P.n_is_3 && !P.use_age) {
        // Arrhenius-type laws with Glen exponent 3 (the default path).  Same quantities as the generic loop
        // below, regrouped (every regrouping is exact in real arithmetic and moves results by a few ulp):
        //   E = (E1 + E2) / 2 is kept as the sum s; T = s (0.5 / c_i) + T_0; the cold-ice test E < E_cts(p) is
        //   s < cts2_a - cts2_b p; T_pa = T - T_m + T_melting = T + beta p; A exp(-Q / (R T)) = exp(ln A - (Q/R) / T);
        //   delta = (e theta 2 alpha^2) p^3 softness;  (depth[k] + dz) delta[k-1] = depth[k-1] delta[k-1].
        // NL levels per trip (NL independent Arrhenius chains in flight per thread).  The range is walked with
        // stepped pointers and compile-time offsets: whole trips carry no predication at all, and only the trip
        // that the ice surface cuts short (at most one per column: Lc is a multiple of NL) is predicated.
        const double K = c2c * (hx * hx + hy * hy);
        double gprev = 0.0; // depth[k-1] * delta[k-1]
        constexpr int NL = SLAB_NL;
        const double *e1 = E1 + k0, *e2 = E2 + k0;
        const double2 *zp = zz + k0;
        double *ic = Ic + k0;
        int left = ke - k0 + 1; // levels of this thread still to do
        bool is_first = true;
        const double2 *selAQw = selAQ + 1;
        auto trip = [&](auto tail_tag) {
          constexpr bool TAIL = decltype(tail_tag)::value; // fewer than NL levels left: slots >= left repeat the last one
          double2 zh[NL];
          double s[NL], dep[NL], pr[NL], T[NL], arg[NL], soft[NL], d[NL], g[NL];
#pragma unroll
          for (int j = 0; j < NL; ++j) {
            const int jj = TAIL ? min(j, left - 1) : j;
            zh[j] = zp[jj];
            s[j] = e1[jj] + e2[jj];
          }
#pragma unroll
          for (int j = 0; j < NL; ++j) {
            dep[j] = thk - zh[j].x;
            pr[j] = fma(P.rg, dep[j], P.p_air);
            const double Tc = fma(s[j], P.hic, P.T_0); // E / c_i + T_0
            if (LAW == LAW_GPBLD) {
              T[j] = fma(P.ec_beta, pr[j], Tc); // EnthalpyConverter.cc:196-198
            } else {
              const double T_m = fma(-P.ec_beta, pr[j], P.T_melting);
              T[j] = fmin(Tc, T_m); // EnthalpyConverter::temperature, :180-188
              if (LAW == LAW_PB) T[j] = fma(P.beta_ratio, pr[j], T[j]); // rheology/PatersonBudd.cc:57
            }
          }
#pragma unroll
          for (int j = 0; j < NL; ++j) {
            // ln A - (Q / R) / T in units of ln2 / 16
            const double rT = rcp_cubic(T[j]);
            if (LAW == LAW_ARR) {
              arg[j] = fma(-P.QoR2_cold, rT, P.lnA2_cold);
            } else if (LAW == LAW_ARRWARM) {
              arg[j] = fma(-P.QoR2_warm, rT, P.lnA2_warm);
            } else {
              // rheology/FlowLaw.cc:89-94: {ln A, Q / R} of the cold or the warm branch, one 16-byte load
              const double2 aq = *((T[j] < P.T_crit) ? selAQ : selAQw);
              arg[j] = fma(-aq.y, rT, aq.x);
            }
          }
#pragma unroll
          for (int j = 0; j < NL; ++j) soft[j] = exp2_tab16(arg[j], tab16);
          if (LAW == LAW_GPBLD) {
            // temperate ice: E >= E_cts(p), i.e. E / c_i + T_0 + beta p >= T_melting (the flow law is continuous there,
            // so a point within rounding of the CTS may take either branch)
            bool any_temperate = false;
#pragma unroll
            for (int j = 0; j < NL; ++j) any_temperate |= !(T[j] < P.T_melting);
            if (any_temperate) { // rheology/GPBLD.cc:55-60
#pragma unroll
              for (int j = 0; j < NL; ++j) {
                if (!(T[j] < P.T_melting)) {
                  const double cts2 = fma(-P.cts2_b, pr[j], P.cts2_a);
                  const double T_m = fma(-P.ec_beta, pr[j], P.T_melting);
                  const double Lm = fma(P.c_w - P.c_i, T_m - 273.15, P.L0); // EnthalpyConverter::L, :365-367
                  const double omega = fmin(fmax(0.5 * (s[j] - cts2), 0.0) * rcp_fast(Lm), P.gp_limit);
                  soft[j] = P.gp_softness_T0 * fma(P.gp_coeff, omega, 1.0);
                }
              }
            }
          }
#pragma unroll
          for (int j = 0; j < NL; ++j) {
            d[j] = (K * (pr[j] * pr[j] * pr[j])) * soft[j];
            g[j] = dep[j] * d[j];
          }
          first = is_first ? d[0] : first;
#pragma unroll
          for (int j = 0; j < NL; ++j) {
            // (the trapezoid ending at the first level of the range needs the range below: added after the loop;
            // repeated slots of a short trip carry weight 0 and re-store the value of the last level)
            double hz = zh[j].y;
            if (j == 0) hz = is_first ? 0.0 : hz;
            if (TAIL && j > 0) hz = (j < left) ? hz : 0.0;
            run = fma(hz, prev + d[j], run);
            dp = fma(hz, gprev + g[j], dp);
            if (FULL) ic[TAIL ? min(j, left - 1) : j] = run;
            prev = d[j], gprev = g[j];
          }
        };
        while (left >= NL) {
          trip(std::false_type{});
          e1 += NL, e2 += NL, zp += NL, ic += NL, left -= NL;
          is_first = false;
        }
        if (left > 0) trip(std::true_type{});
      } else
      for (int k = k0; k <= ke; ++k) {
        const double2 zh = zz[k];
        const double dep = thk - zh.x;                 // :641-643
        double Ea[1], pr[1], st[1], gsz[1], fl[1];
        Ea[0] = 0.5 * (E1[k] + E2[k]);                 // :677-684
        pr[0] = fma(P.rg, dep, P.p_air);               // EnthalpyConverter.cc:146-152
        st[0] = alpha * pr[0];                         // :688
        gsz[0] = P.grain_size;
        double c2 = c2c;
        if (P.use_age) { // :649-675 (uniform branch; off by default)
          const double ag = 0.5 * (A_s[o1 + k] + A_s[o2 + k]);
          if (P.gs_age) gsz[0] = grain_size_vostok(ag * P.years_per_second);
          if (P.e_age) c2 = (interglacial(P, P.current_time - ag) ? P.e_inter : P.e) * theta * 2.0;
        }
        flow_lean_v<LAW, 1>(P, st, Ea, pr, gsz, fl); // :691
        const double d = c2 * pr[0] * fl[0];         // :696
        // trapezoids   I: 0.5 dz (delta[k-1] + delta[k])                          compute_I, :855-858
        //              D: 0.5 dz ((depth[k] + dz) delta[k-1] + depth[k] delta[k])  :701-705
        // (the trapezoid ending at the first level of the range needs the range below: added after the loop)
        const bool is_first = (k == k0);
        const double hz = is_first ? 0.0 : zh.y;
        run = fma(hz, prev + d, run);
        dp = fma(hz, fma(dep + (zh.y + zh.y), prev, dep * d), dp);
        if (FULL) Ic[k] = run; // range-local prefix; the offset of the range is added below
        first = is_first ? d : first;
        prev = d;
      }
      if (ke >= k0 && ke == ks) { // :707-708 (dz = thk - z[ks] = depth[ks])
        const double depk = thk - zz[ks].x;
        dp = fma(0.5 * depk * depk, prev, dp);
      }
      // what the ranges of a column exchange: {last delta, first delta} and {I sum, D sum} of each range
      {
        double2 *px = reinterpret_cast<double2 *>(sX) + (pt * WZ + w) * NC + c; // (two planes: 16 lanes, 16 contiguous entries)
        px[0] = make_double2(prev, first);
        px[2 * WZ * NC] = make_double2(run, dp);
      }
      __syncthreads(); // #1: the ranges' sums visible; every read of row r's slot is done
      if (prefetch) {
        issue_row(it + 2, s_cur);
        pending |= (1u << s_cur);
      }
      // The trapezoid across the lower boundary of a range (the range below ends at its first level - 1) belongs to that
      // range.  Every thread of a column evaluates the boundaries of all its ranges itself -- a few operations -- instead
      // of a second exchange through shared memory with its own barrier.  Same expressions, same order of the sums as
      // ever: I offset = sum over the ranges below in ascending order; D = ranges 1, 2, ..., WZ - 1, then range 0.
      if (FULL || w == 0) {
        const double2 *px = reinterpret_cast<const double2 *>(sX) + pt * WZ * NC + c;
        double off = 0.0, tot = 0.0, bI = 0.0, D0 = 0.0, dl = 0.0;
#pragma unroll
        for (int ww = 0; ww < WZ; ++ww) {
          const double2 pf = px[ww * NC], rd = px[ww * NC + 2 * WZ * NC]; // {last, first delta}, {I sum, D sum} of range ww
          double bIw = 0.0, bDw = 0.0;
          if (ww > 0 && ks >= ww * Lc) { // the range is not empty
            const double2 zh = zz[ww * Lc];
            const double depF = thk - zh.x;
            bIw = zh.y * (dl + pf.y);
            bDw = zh.y * fma(depF + (zh.y + zh.y), dl, depF * pf.y);
          }
          const double t = rd.x + bIw; // empty range: 0
          off += (ww < w) ? t : 0.0;
          tot += t;
          if (ww == 0) {
            D0 = rd.y + bDw;
          } else {
            Dsum += rd.y + bDw;
          }
          bI = (ww == w) ? bIw : bI;
          dl = pf.x;
        }
        Dsum += D0;
        if (FULL) {
          if (w > 0 && ke >= k0) {
            const double add = off + bI;
            double *icp = Ic + k0;
            int n = ke - k0 + 1;
            for (; n >= 4; n -= 4, icp += 4) {
              const double a0 = icp[0], a1 = icp[1], a2 = icp[2], a3 = icp[3];
              icp[0] = a0 + add, icp[1] = a1 + add, icp[2] = a2 + add, icp[3] = a3 + add;
            }
            for (; n > 0; --n, ++icp) *icp += add;
          }
          // above the ice I keeps its last value (:861-863); ice-free points (ks = -1): I = 0 everywhere
          for (int k = ks + 1 + w; k < Mz; k += WZ) Ic[k] = tot;
        }
      }
      if (FULL) ivalid |= 4u | (1u << s_cur);
    } else {
      if (prefetch) { // nothing reads the slot of row r: it can be refilled at once
        issue_row(it + 2, s_cur);
        pending |= (1u << s_cur);
      }
      if (FULL) ivalid &= ~(4u | (1u << s_cur));
    }

    // D, flux, D_max (thread w = 0 of the point), sia/SIAFD.cc:711-731, :772-793
    if (w == 0) {
      if (FULL) {
        double *cfr = cf + (s_cur * NC + c) * 4 + 2 * pt;
        cfr[0] = -0.25 * hx, cfr[1] = -0.25 * hy; // the factor of sia/SIAFD.cc:928-931, exact (rows without ice: 0)
      }
      if (own_c && r >= ra) {
        const bool edge = (i_c < 0 || i_c >= P.Mx - 1 || r < 0 || r >= P.My - 1);
        const double hq = pt ? hy : hx; // (rows without ice: D = 0 and the flux is written as -0 * 0)
        double D = 0.0;
        if (act) {
          D = edge ? 0.0 : Dsum;
          if (P.limit_diffusivity && D >= P.D_limit) {
            D = P.D_limit;
            hdc_local += 1;
          }
          dmax_local = fmax(dmax_local, D);
        }
#ifdef SLAB_DIAG_NODQ // (diagnostic builds: D, Q of rows without ice are not written)
        if (row_active)
#endif
        {
          *DQ_D = D;
          *DQ_Q = -D * hq;
        }
      }
    }

    if (FULL) {
      const bool any_valid = ivalid != 0;
      if (any_valid) {
        // materialise logically-zero arrays (only at the edge of the ice)
        if (!(ivalid & 4u)) {
          for (int e = tid; e < NC * S; e += NT) I0_s[e] = 0.0;
        }
#pragma unroll
        for (int s = 0; s < 2; ++s) {
          if (!(ivalid & (1u << s))) {
            for (int e = tid; e < NC * S; e += NT) I1_s[s * NC * S + e] = 0.0;
          }
        }
        __syncthreads(); // #3: I0, I1, cf final
      }
      // ---------------- stage B: u, v of the regular columns, sia/SIAFD.cc:904-943 ----------------
      // Write-only rows (no ice at any staggered point of this and the previous row, G9) whose sliding velocity is zero
      // everywhere in the strip: u = v = 0 on every level.  The owned columns of a row are contiguous in memory, so
      // each field's row leaves as ONE bulk store from a block of zeros (measured: 5.2 -> 5.9 TB/s in this regime).
      // (Tried in round 2: thread t storing row r + t of a whole run of such rows at once.  Slower, 4.75 -> 4.90 ms in the
      // ice-free regime at 4096^2: the DRAM takes few concurrent store streams best -- tools/store_pattern.cu, 6.0 TB/s
      // with 16 CTAs per SM writing against 7.0 with 3 -- and sixty rows per CTA in flight scatter the writes.)
      bool row_done = false;
      // rows / strips on the rim of the patch (with a communicator): their u, v also go into the neighbours' ghost cells
      const bool sn_row = (r - P.ys) < PP.w || (r - P.ys) >= P.ym - PP.w;
      const bool rim_row = PP.on && r >= rbase && r < rend && (strip_rim || sn_row); // CTA-uniform
      const bool corner_row = rim_row && strip_rim && sn_row; // holds a cell that faces three neighbours (push_rim_row)
      // per pass of stage B: the ghost cell of the first neighbour that faces this thread's column (the neighbour table
      // is read here, where it is a kernel parameter in the constant bank)
      auto rim_targets = [&]() -> RimTargets {
        RimTargets T;
        const int b = r - P.ys;
        const bool S_ = b < PP.w, N_ = b >= P.ym - PP.w;
#pragma unroll
        for (int p = 0; p < NPASS; ++p) {
          const int qp = qg + p, a = ca + qp - P.xs;
          const bool W_ = a < PP.w, E_ = a >= P.xm - PP.w;
          T.u2[p] = nullptr, T.v2[p] = nullptr;
          if (qp >= 1 && a >= 0 && a < P.xm) {
            for (int d = 0; d < 8; ++d) {
              if (PP.a[d] != nullptr && peer_strip_member(d, W_, E_, S_, N_)) {
                const long t = ((long)(b + P.wuv + PP.dj[d]) * PP.rowc[d] + (a + P.wuv + PP.di[d])) * Mz + lz;
                T.u2[p] = PP.a[d] + t, T.v2[p] = PP.b[d] + t;
                break;
              }
            }
          }
        }
        return T;
      };
      if (BULK && !any_valid && !corner_row && r >= rbase && r < rend) { // CTA-uniform
        if (!slide) {
          if (!zero_ready) {
            for (int e = tid; e < NC * S; e += NT) I0_s[e] = 0.0;
            fence_proxy_async();
            __syncthreads();
            zero_ready = true;
          }
          if (tid == 0) {
            const int i_first = max(ca + 1, P.xs), i_last = min(ca + NC - 1, P.xs + P.xm - 1);
            const long n = (long)(i_last - i_first + 1) * Mz;
            if (n > 0) {
              const long g0 = ((long)(r - (P.ys - P.wuv)) * (P.xm + 2 * P.wuv) + (i_first - P.xs + P.wuv)) * Mz;
              const long a0 = (g0 + 1) & ~1L, a1 = (g0 + n) & ~1L; // the 16-byte aligned middle
              if (g0 & 1) F.u[g0] = 0.0, F.v[g0] = 0.0;
              if ((g0 + n) & 1) F.u[g0 + n - 1] = 0.0, F.v[g0 + n - 1] = 0.0;
              if (a1 > a0) {
                bulk_s2g(F.u + a0, I0_s, (unsigned)((a1 - a0) * 8));
                bulk_s2g(F.v + a0, I0_s, (unsigned)((a1 - a0) * 8));
                bulk_commit();
              }
            }
          }
          if (rim_row) { // u = v = 0 on every level: the neighbours' ghost cells get the same
            const RimTargets T = rim_targets();
#pragma unroll
            for (int p = 0; p < NPASS; ++p) {
              if (T.u2[p] != nullptr) {
                for (int k = 0; lz + k < Mz; k += LZ) T.u2[p][k] = 0.0, T.v2[p][k] = 0.0;
              }
            }
          }
          zero_inflight = true;
          row_done = true;
        }
      }
      if (!row_done && r >= rbase && r < rend) { // CTA-uniform
        if (rim_row) {
          // rows / strips on the rim of the patch (with a communicator): out of line, so that the march stays lean
          const RimTargets T = rim_targets();
          stage_b_rim<NC, NPASS, LB, LZ>(Mz, P.xs, P.xs + P.xm - 1,
                                         ((long)(r - (P.ys - P.wuv)) * (P.xm + 2 * P.wuv) + P.wuv) * Mz, T, F.u, F.v, I0_s, I1_s,
                                         cf, S, s_cur, (ivalid & (1u << s_nxt)) != 0, any_valid, ca, qg, lz, lane, sv.x, sv.y);
        } else {
          const bool south = (ivalid & (1u << s_nxt)) != 0;
#pragma unroll
          for (int p = 0; p < NPASS; ++p) {
            const int qp = qg + p;
            const int srcl = (lane & ~(LZ - 1)) + p * LB; // a lane whose own column (tid / LB) is qp
            const double ub = __shfl_sync(FULLMASK, sv.x, srcl), vb = __shfl_sync(FULLMASK, sv.y, srcl);
            if (pass_ok & (1u << p)) {
              const long uvo = uvo_run + p * Mz;
              double *up = F.u + uvo, *vp = F.v + uvo;
              if (!any_valid) {
                // no ice at any staggered point of this and the previous row: I == 0, u = sliding velocity (G9)
                for (int k = lz; k < Mz; k += LZ, up += LZ, vp += LZ) {
                  *up = ub;
                  *vp = vb;
                }
              } else {
                const double2 *c2 = reinterpret_cast<const double2 *>(cf);
                const double2 ce = c2[(s_cur * NC + qp) * 2], cw = c2[(s_cur * NC + qp - 1) * 2], cn = c2[(s_cur * NC + qp) * 2 + 1];
                const double2 cs = south ? c2[(s_nxt * NC + qp) * 2 + 1] : make_double2(0.0, 0.0);
                const double *Ie = I0_s + qp * S + lz;
                const double *In = I1_s + (s_cur * NC + qp) * S + lz, *Is = I1_s + (s_nxt * NC + qp) * S + lz;
                uv_column<LZ, false>(Mz, lz, S, ub, vb, ce, cw, cn, cs, Ie, In, Is, up, vp, nullptr, nullptr);
              }
            }
          }
        }
      }
      // (only where this strip holds a corner cell of the patch; out of line: keeps the march lean)
      if (corner_row) push_rim_row<NPASS, LB, LZ>(P, PP, F.u, F.v, r, ca, qg, lz);
      if (r >= rbase && r < rend) uvo_run += uv_row;
      if (any_valid) {
        // the zero-filled arrays now ARE valid zeros for the next row's "previous row" role
        ivalid |= (1u << s_cur);
        __syncthreads(); // #4: stage B reads done before the next row's stage A overwrites I0 / I1[s_nxt]
      }
    }
  }

  if (FULL && zero_inflight && tid == 0) bulk_wait_read0(); // shared memory must outlive the stores that read it
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
    __syncthreads(); // sX is free
    unsigned long long *wm = (unsigned long long *)sX;
    int *wc = (int *)(sX + 32);
    if ((tid & 31) == 0) {
      wm[tid >> 5] = m;
      wc[tid >> 5] = cnt;
    }
    __syncthreads();
    if (tid == 0) {
      const int nw = (NT + 31) >> 5;
      for (int ww = 1; ww < nw; ++ww) {
        m = (wm[ww] > m) ? wm[ww] : m;
        cnt += wc[ww];
      }
      if (m != 0ull) atomicMax(F.dmax, m);
      if (cnt != 0) atomicAdd(F.hdc, cnt);
    }
  }
}

// =============================================================================================
// launch
// =============================================================================================
static size_t slab_smem_bytes(const DP &P, bool full, int NC, int WZ, bool bulk) {
  const long Mz = P.Mz, S = bulk ? Mz : (Mz | 1);
  const long slotE = ((NC + 1) * S + 2 + 1) & ~1L;
  const long colI = (NC * S + 1) & ~1L;
  long d = 2 * Mz + 2 * slotE + (P.use_age ? 2 * slotE : 0) + 2 * AUX_N + (full ? 3 * colI : 0) + 2 * WZ * NC * 4 +
           2 * NC * 4 + 16 + 4;
  return (size_t)d * 8 + 2 * 8 /* mbarriers */ + 6 * sizeof(RowCopy) + 16;
}

template <int LAW, bool FULL, int NC, int WZ, bool BULK>
static int launch_slab_t(const DP &P, const Fields &F, const SlabArgs &A, cudaStream_t s, const PeerPush &PP) {
  const size_t smem = slab_smem_bytes(P, FULL, NC, WZ, BULK);
  if (smem > (size_t)227 * 1024) return -1;
  // per instantiation and per device (the attribute is a per-device setting); handles on different host threads may
  // arrive here together
  static std::atomic<size_t> configured[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return -1;
  const int slot = dev & 63;
  if (smem > configured[slot].load(std::memory_order_acquire)) {
    if (cudaFuncSetAttribute(k_sia_slab<LAW, FULL, NC, WZ, BULK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) !=
        cudaSuccess) {
      return -1;
    }
    size_t seen = configured[slot].load(std::memory_order_relaxed);
    while (seen < smem && !configured[slot].compare_exchange_weak(seen, smem, std::memory_order_release)) {
    }
  }
  dim3 grid((unsigned)((P.xm + 1 + (NC - 2)) / (NC - 1)), (unsigned)A.nseg);
  k_sia_slab<LAW, FULL, NC, WZ, BULK><<<grid, 2 * NC * WZ, smem, s>>>(P, F, A, PP);
  return 1;
}

#ifndef SLAB_DEV
template <int LAW, bool FULL> static int launch_slab_l(const DP &P, const Fields &F, const Tuning &T, const SlabArgs &A,
                                                       cudaStream_t s, const PeerPush &PP) {
  // 16 lane columns unless shared memory cannot hold them (very tall grids): then 8.  Rows arrive by bulk copy (odd Mz:
  // columns are an odd number of doubles apart as they lie in memory) or by 8-byte cp.async into padded columns.
  const bool bulk = A.use_bulk != 0;
#ifdef SLAB_WZ_VARIANTS // z ranges per column other than 4 (tuning experiments; SIAFD_B200_WZ)
  if (slab_smem_bytes(P, FULL, 16, 8, bulk) <= (size_t)227 * 1024 && T.wz != 4 && bulk) {
    if (T.wz == 2) return launch_slab_t<LAW, FULL, 16, 2, true>(P, F, A, s, PP);
    return launch_slab_t<LAW, FULL, 16, 8, true>(P, F, A, s, PP);
  }
#endif
  (void)T;
  if (slab_smem_bytes(P, FULL, 16, 4, bulk) <= (size_t)227 * 1024) {
    return bulk ? launch_slab_t<LAW, FULL, 16, 4, true>(P, F, A, s, PP) : launch_slab_t<LAW, FULL, 16, 4, false>(P, F, A, s, PP);
  }
  return bulk ? launch_slab_t<LAW, FULL, 8, 4, true>(P, F, A, s, PP) : launch_slab_t<LAW, FULL, 8, 4, false>(P, F, A, s, PP);
}

#endif

template <bool FULL> static int launch_slab_f(const DP &P, const Fields &F, const Tuning &T, const SlabArgs &A,
                                              cudaStream_t s, const PeerPush &PP) {
#ifdef SLAB_DEV // development builds: one instantiation (gpbld, 16 lane columns, 4 z ranges) compiles in seconds
  (void)T;
#ifdef SLAB_DEV_WZ8 // (experiment: 8 z ranges per column, 256 threads per CTA; build with -DSLAB_MINB=2)
  return (P.law == LAW_GPBLD && A.use_bulk) ? launch_slab_t<LAW_GPBLD, FULL, 16, 8, true>(P, F, A, s, PP) : -1;
#else
  return (P.law == LAW_GPBLD && A.use_bulk) ? launch_slab_t<LAW_GPBLD, FULL, 16, 4, true>(P, F, A, s, PP) : -1;
#endif
#else
  switch (P.law) {
  case LAW_ISO:
    return launch_slab_l<LAW_ISO, FULL>(P, F, T, A, s, PP);
  case LAW_PB:
    return launch_slab_l<LAW_PB, FULL>(P, F, T, A, s, PP);
  case LAW_GPBLD:
    return launch_slab_l<LAW_GPBLD, FULL>(P, F, T, A, s, PP);
  case LAW_HOOKE:
    return launch_slab_l<LAW_HOOKE, FULL>(P, F, T, A, s, PP);
  case LAW_ARR:
    return launch_slab_l<LAW_ARR, FULL>(P, F, T, A, s, PP);
  case LAW_ARRWARM:
    return launch_slab_l<LAW_ARRWARM, FULL>(P, F, T, A, s, PP);
  case LAW_GK:
    return launch_slab_l<LAW_GK, FULL>(P, F, T, A, s, PP);
  default:
    return -1;
  }
#endif
}

size_t slab_smem_need(const DP &P, bool full, bool bulk) { return slab_smem_bytes(P, full, 8, 4, bulk); }

int slab_rows_per_segment(const Tuning &T) { return T.rows_per_cta < 88 ? T.rows_per_cta : 88; }
int slab_segments(const DP &P, const Tuning &T) {
  const int RS = slab_rows_per_segment(T);
  return (P.ym + 2 + RS - 1) / RS;
}

int launch_slab(const DP &P, const Fields &F, bool full, const Tuning &T, long nE, long n2, double inv_dz, int seg0,
                int nseg, cudaStream_t s, const PeerPush *push, const int *seg_order) {
  SlabArgs A;
  A.RS = slab_rows_per_segment(T); // the row flags of a CTA live in a 96-bit word
  A.seg0 = seg0;
  A.nseg = nseg < 0 ? slab_segments(P, T) - seg0 : nseg;
  A.order = (seg0 == 0 && A.nseg == slab_segments(P, T)) ? seg_order : nullptr;
  A.use_bulk = (T.use_bulk_copy && (P.Mz & 1)) ? 1 : 0; // even Mz: padded columns, 8-byte cp.async
  A.skip_rows = T.skip_ice_free;
  A.nE = nE;
  A.n2 = n2;
  A.inv_dz = inv_dz;
  PeerPush PP = PeerPush();
  if (push != nullptr && push->on && full) PP = *push;
  return full ? launch_slab_f<true>(P, F, T, A, s, PP) : launch_slab_f<false>(P, F, T, A, s, PP);
}

} // namespace siafd
