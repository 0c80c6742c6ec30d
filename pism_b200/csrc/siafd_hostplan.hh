// siafd_hostplan.hh -- what the host-array form of siafd_b200_update moves over PCIe and what the host fills in
// itself: plain C++ (no CUDA), so that the plan can be checked without a GPU (siafd_b200_host_plan_emulate,
// tests/test_host_plan.py).  Included by siafd_capi.cu only.  Private to the library.
#pragma once
#include "../../include/siafd_b200.h"

#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>
#if defined(__x86_64__) || defined(_M_X64)
#include <emmintrin.h>
#include <xmmintrin.h>
#endif

namespace siafd_hostplan {

// siafd_b200_update with host arrays, full update: the three legs of the drop-in call -- host->device copy of the
// enthalpy, the fused kernel, device->host copy of u and v -- run as a pipeline over bands of rows on three streams
// (PCIe is full duplex), instead of one after the other.  The bands are whole row segments of the fused kernel, and
// every array is contiguous in rows, so each leg of a band is one cudaMemcpyAsync / one launch.
// Where there is no ice the fused kernel reads no enthalpy and writes u = v = sliding velocity on every level
// (SIAFD.cc:631-637, :935-942), so those parts of the 3D arrays need not cross PCIe: per band of rows only the
// rectangle of columns within 3 cells of ice goes up (enthalpy) and comes down (u, v); the rest of u, v is filled in
// place on the host by a few threads while the copies run.  The host arrays end up bit-identical to a full transfer.
struct IceExtent {
  std::vector<int> lo, hi; // per owned row: columns [lo, hi] to transfer (lo > hi: none; lo < 0: the whole row)
};

// per row of `rows` rows of `pitch` cells starting at H: the first and last of the cells [0, n) that hold ice (first >
// last: none); rows shared between `threads` threads (134 MB of thickness at 4096^2: 17 ms on one thread)
static void row_ice_bounds(const double *H, int rows, long pitch, int n, int *l0, int *h0, int threads) {
  auto work = [=](int first, int stride) {
    for (int r = first; r < rows; r += stride) {
      const double *row = H + (long)r * pitch;
      int a = 0, b = n - 1;
      while (a < n && row[a] == 0.0) ++a;
      while (b >= a && row[b] == 0.0) --b;
      l0[r] = a, h0[r] = b;
    }
  };
  const int nt = std::max(1, std::min(threads, rows / 64));
  std::vector<std::thread> w;
  for (int t = 1; t < nt; ++t) w.emplace_back(work, t, nt);
  work(0, nt);
  for (auto &t : w) t.join();
}

static void ice_extent(const siafd_b200_config &c, const double *H, IceExtent &E, bool patch_mode, int threads = 1) {
  const int xm = c.xm, ym = c.ym, wg = c.w_geom;
  const long pitch = xm + 2 * wg;
  if (patch_mode) {
    // one patch of a decomposed domain: the ghost cells (width wg >= 2) are the neighbours' thickness, no wrap.  A
    // column matters when there is ice within one cell of it (the staggered points around it); two cells are taken.
    const int margin = 2, R = ym + 2 * wg;
    std::vector<int> l0(R), h0(R); // per local row (ghost rows included): local columns [l0, h0] with ice, ghosts included
    row_ice_bounds(H, R, pitch, (int)pitch, l0.data(), h0.data(), threads);
    E.lo.assign(ym, xm), E.hi.assign(ym, -1);
    for (int j = 0; j < ym; ++j) {
      int lo = 1 << 30, hi = -1;
      for (int d = -margin; d <= margin; ++d) {
        const int r = j + wg + d; // wg >= margin: always a row of the array
        if (l0[r] <= h0[r]) lo = std::min(lo, l0[r]), hi = std::max(hi, h0[r]);
      }
      if (hi >= lo) {
        // owned-column indices, clipped to the patch (the enthalpy upload widens a range that touches an edge of the
        // patch to the ghost columns beyond it; the ghost columns of u, v come down separately)
        lo = std::max(lo - wg - margin, 0), hi = std::min(hi - wg + margin, xm - 1);
        E.lo[j] = lo, E.hi[j] = hi;
      }
    }
    return;
  }
  const int margin = 3;
  std::vector<int> l0(ym), h0(ym); // (first > last: no ice in this row)
  row_ice_bounds(H + (long)wg * pitch + wg, ym, pitch, xm, l0.data(), h0.data(), threads);
  E.lo.assign(ym, xm), E.hi.assign(ym, -1);
  for (int j = 0; j < ym; ++j) {
    int lo = xm, hi = -1;
    for (int d = -margin; d <= margin; ++d) { // rows wrap periodically, like the ghosts
      const int jj = ((j + d) % ym + ym) % ym;
      if (l0[jj] <= h0[jj]) lo = std::min(lo, l0[jj]), hi = std::max(hi, h0[jj]);
    }
    if (hi >= lo) {
      lo -= margin, hi += margin;
      if (lo < margin || hi > xm - 1 - margin) lo = -1, hi = xm; // ice near the edge of the domain: whole rows
    }
    E.lo[j] = lo, E.hi[j] = hi;
  }
}

// rectangle of the owned rows [j0, j1) (wrapped into [0, ym)): returns false when there is nothing to transfer
static bool band_extent(const IceExtent &E, int ym, int j0, int j1, int *lo, int *hi, bool *whole, bool patch_mode = false) {
  int a = 1 << 30, b = -1;
  *whole = false;
  for (int j = j0; j < j1; ++j) {
    // (a patch: the ghost rows take the extent of the nearest owned row, which already looks two rows past the edge)
    const int jj = patch_mode ? std::min(std::max(j, 0), ym - 1) : ((j % ym) + ym) % ym;
    if (E.hi[jj] < E.lo[jj]) continue;
    if (E.lo[jj] < 0) *whole = true;
    a = std::min(a, E.lo[jj]), b = std::max(b, E.hi[jj]);
  }
  *lo = a, *hi = b;
  return b >= a || *whole;
}

// n doubles starting at p set to `value` with non-temporal stores where the ISA has them: the filled parts of u, v
// (gigabytes) are not read again by this call, and a regular store would first read every cache line it overwrites,
// doubling the DRAM traffic that competes with the PCIe copies landing in the same arrays
static void fill_stream(double *p, size_t n, double value) {
#if defined(__x86_64__) || defined(_M_X64)
  size_t k = 0;
  while (k < n && (reinterpret_cast<uintptr_t>(p + k) & 15u)) p[k++] = value;
  const __m128d v = _mm_set1_pd(value);
  for (; k + 8 <= n; k += 8) {
    _mm_stream_pd(p + k, v);
    _mm_stream_pd(p + k + 2, v);
    _mm_stream_pd(p + k + 4, v);
    _mm_stream_pd(p + k + 6, v);
  }
  for (; k + 2 <= n; k += 2) _mm_stream_pd(p + k, v);
  for (; k < n; ++k) p[k] = value;
#else
  std::fill(p, p + n, value);
#endif
}

struct FillTask {
  int j, c0, c1; // owned row, local columns [c0, c1) of the u / v arrays (ghost columns included)
};

// u, v of ice-free columns: the sliding velocity on every level (zero when there is no sliding field)
static void fill_rows(const siafd_b200_config &c, const double *sliding, double *u, double *v, const FillTask *tasks,
                      size_t n, size_t first, size_t stride) {
  const int wuv = c.w_uv, wsl = c.w_sliding, Mz = c.Mz, xm = c.xm;
  const long rowUV = (long)(xm + 2 * wuv) * Mz, pitchS = (long)(xm + 2 * wsl) * 2;
  for (size_t t = first; t < n; t += stride) {
    const FillTask &T = tasks[t];
    double *ur = u + (long)(T.j + wuv) * rowUV, *vr = v + (long)(T.j + wuv) * rowUV;
    bool zero = true;
    if (sliding) {
      const double *sr = sliding + (long)(T.j + wsl) * pitchS + 2L * wsl;
      for (int cc = T.c0; cc < T.c1 && zero; ++cc) {
        const int i = ((cc - wuv) % xm + xm) % xm; // ghost columns wrap periodically
        zero = (sr[2L * i] == 0.0 && sr[2L * i + 1] == 0.0);
      }
      if (!zero) {
        for (int cc = T.c0; cc < T.c1; ++cc) {
          const int i = ((cc - wuv) % xm + xm) % xm;
          fill_stream(ur + (long)cc * Mz, (size_t)Mz, sr[2L * i]);
          fill_stream(vr + (long)cc * Mz, (size_t)Mz, sr[2L * i + 1]);
        }
      }
    }
    if (zero) {
      fill_stream(ur + (long)T.c0 * Mz, (size_t)(T.c1 - T.c0) * Mz, 0.0);
      fill_stream(vr + (long)T.c0 * Mz, (size_t)(T.c1 - T.c0) * Mz, 0.0);
    }
  }
#if defined(__x86_64__) || defined(_M_X64)
  _mm_sfence(); // the streamed stores are globally visible before the thread is joined
#endif
}

// ---- the level cut of the sparse host path --------------------------------------------------------------------
// Of a column that is within reach of ice only the levels up to the thickest ice next to it matter: a staggered point
// reads the enthalpy of its two columns on the levels k <= ks = kBelowHeight(thk) only (SIAFD.cc:613-627, :676-689),
// I is constant above ks (SIAFD.cc:857-859), and so are u and v above the largest ks of the four staggered points
// around a column (SIAFD.cc:935-942).  The staggered thickness is the mean of two thk_smooth values, and with the bed
// smoother off thk_smooth is 0 (no ice), max(usurf - topg, 0) (grounded) or H (floating) (BedSmoother.cc:306-320), so
// the largest of those over the column and its neighbours bounds every thickness the column takes part in.
static inline int host_levels_needed(const double *z, int Mz, double max_thickness) {
  if (!z || Mz < 2) return Mz;
  if (!(max_thickness == max_thickness)) return Mz; // NaN: no claim
  // first level above the thickness = ks + 1 levels are read / differ; one more level of slack
  const int kub = (int)(std::upper_bound(z, z + Mz, max_thickness) - z);
  const int n = kub + 1;
  return n >= Mz - 2 ? Mz : std::max(n, 2); // (a cut that saves two or three levels is not worth a strided copy)
}

struct ThkMap {
  static constexpr int kBlock = 32; // columns per block
  int rows = 0, pitch = 0, blocks = 0;
  std::vector<double> mx; // [rows][blocks]: max(H, surface - bed) over the block's local columns
  double range_max(int r0, int r1, int c0, int c1) const { // local rows [r0, r1], columns [c0, c1], clamped
    r0 = std::max(r0, 0), r1 = std::min(r1, rows - 1), c0 = std::max(c0, 0), c1 = std::min(c1, pitch - 1);
    double m = 0.0;
    for (int r = r0; r <= r1; ++r) {
      const double *q = mx.data() + (size_t)r * blocks;
      for (int b = c0 / kBlock; b <= c1 / kBlock; ++b) m = std::max(m, q[b]);
    }
    return m;
  }
};

static void thk_map_rows(const double *H, const double *surface, const double *bed, const double *mask, ThkMap *M, int first,
                         int stride) {
  for (int r = first; r < M->rows; r += stride) {
    const long o = (long)r * M->pitch;
    const double *hr = H + o, *sr = surface + o, *br = bed + o, *mr = mask + o;
    double *q = M->mx.data() + (size_t)r * M->blocks;
    for (int b = 0; b < M->blocks; ++b) {
      const int c1 = std::min(M->pitch, (b + 1) * ThkMap::kBlock);
      double m = 0.0;
      bool bad = false, any = false;
      for (int cc = b * ThkMap::kBlock; cc < c1; ++cc) any |= (hr[cc] != 0.0); // (a NaN counts as ice)
      if (!any) { // no ice in the block: thk_smooth = 0 whatever the other fields say, and they need not be read
        q[b] = 0.0;
        continue;
      }
      for (int cc = b * ThkMap::kBlock; cc < c1; ++cc) {
        // thk_smooth with the smoother off (BedSmoother.cc:306-320; maxtl = 0): 0 where H = 0, usurf - topg where the
        // mask says grounded (Mask.hh:37-66: not ocean, i.e. below 3 after rounding), H where it floats
        double t = hr[cc];
        if (t != 0.0 && mr[cc] < 2.5) t = std::max(t, sr[cc] - br[cc]);
        bad |= !(t == t) || !(mr[cc] == mr[cc]);
        m = std::max(m, t);
      }
      q[b] = bad ? HUGE_VAL : m; // a NaN in the inputs: every level moves
    }
  }
}

static void thk_map_build(const siafd_b200_config &c, const double *H, const double *surface, const double *bed,
                          const double *mask, ThkMap &M, int threads) {
  M.rows = c.ym + 2 * c.w_geom, M.pitch = c.xm + 2 * c.w_geom;
  M.blocks = (M.pitch + ThkMap::kBlock - 1) / ThkMap::kBlock;
  M.mx.assign((size_t)M.rows * M.blocks, 0.0);
  const int nt = std::max(1, std::min(threads, M.rows / 64));
  std::vector<std::thread> w;
  for (int t = 1; t < nt; ++t) w.emplace_back(thk_map_rows, H, surface, bed, mask, &M, t, nt);
  thk_map_rows(H, surface, bed, mask, &M, 0, nt);
  for (auto &t : w) t.join();
}

// one strided piece of a 3D array: local rows [r0, r1), local columns [c0, c1), levels [0, n) of Mz
struct Piece {
  int r0, r1, c0, c1, n;
};

// rows [r0, r1) x columns [c0, c1) of an array with ghost width w, cut into chunks of `cut_cols` columns that share a
// level count (local indices of THAT array; the thickness map is indexed with the geometry's ghost width)
static void cut_pieces(const siafd_b200_config &c, const ThkMap *M, int cut_cols, int w, int r0, int r1, int c0, int c1,
                       std::vector<Piece> &out) {
  if (r1 <= r0 || c1 <= c0) return;
  if (!M) {
    out.push_back({r0, r1, c0, c1, c.Mz});
    return;
  }
  const int CW = std::max(8, cut_cols), sh = c.w_geom - w; // array index + sh = geometry index
  for (int a = c0; a < c1;) {
    const int b = std::min(c1, (a / CW + 1) * CW);
    const double m = M->range_max(r0 + sh - 1, r1 - 1 + sh + 1, a + sh - 1, b - 1 + sh + 1);
    const int n = host_levels_needed(c.z, c.Mz, m);
    if (!out.empty() && out.back().r0 == r0 && out.back().r1 == r1 && out.back().c1 == a && out.back().n == n) {
      out.back().c1 = b; // same level count as the chunk before: one copy
    } else {
      out.push_back({r0, r1, a, b, n});
    }
    a = b;
  }
}

// what is left of a downloaded piece of u or v: every level above the cut takes the value of the last level that came
// down
struct ReplTask {
  int band;
  Piece p;
};

// One column: levels [n, Mz) = the value of level n - 1.  The run starts and ends inside cache lines whose other bytes
// the copy engine has just written (the column's own lower levels, the next column's first ones): those two partial
// lines take ordinary stores (a partial non-temporal store leaves the write-combining buffer as a handful of small
// uncached writes -- measured: 1.2 GB/s per thread with streaming stores throughout), the whole lines between them
// are streamed.
static inline void replicate_column(double *col, int n, int Mz) {
  const double v = col[n - 1];
  double *p = col + n, *e = col + Mz;
#if defined(__x86_64__) || defined(_M_X64)
  while (p < e && (reinterpret_cast<uintptr_t>(p) & 63u)) *p++ = v;
  const __m128d vv = _mm_set1_pd(v);
  for (; p + 8 <= e; p += 8) {
    _mm_stream_pd(p, vv);
    _mm_stream_pd(p + 2, vv);
    _mm_stream_pd(p + 4, vv);
    _mm_stream_pd(p + 6, vv);
  }
#endif
  while (p < e) *p++ = v;
}

static void replicate_piece(const Piece &p, double *u, double *v, long row_cells, int Mz) {
  double *arr[2] = {u, v};
  const int ahead = 6; // columns: the two partial lines of a column are asked for while earlier columns are written
  for (int r = p.r0; r < p.r1; ++r) {
    for (int q = 0; q < 2; ++q) {
      double *col = arr[q] + ((long)r * row_cells + p.c0) * Mz;
      for (int cc = p.c0; cc < p.c1; ++cc, col += Mz) {
#if defined(__x86_64__) || defined(_M_X64)
        if (cc + ahead < p.c1) {
          _mm_prefetch(reinterpret_cast<const char *>(col + (long)ahead * Mz + p.n - 1), _MM_HINT_T0);
          _mm_prefetch(reinterpret_cast<const char *>(col + (long)(ahead + 1) * Mz - 1), _MM_HINT_T0);
        }
#endif
        replicate_column(col, p.n, Mz);
      }
    }
  }
}

// The plan of one call.  Band b = the row segments [b band, (b + 1) band) of the fused kernel (RS rows of the extended
// patch each; extended row e = owned row e - 1).
struct HostPlan {
  int RS = 0, nseg = 0, band = 1, NB = 0;
  bool sparse = false, cut = false, patch = false;
  struct Rect {
    int o0, o1, c0, c1; // owned rows [o0, o1), local columns [c0, c1) of u / v to download (c0 >= c1: none)
  };
  std::vector<Rect> down;
  std::vector<Piece> down_pieces; // what comes down of u, v, band after band (local rows / columns of u, v)
  std::vector<size_t> down0;      // pieces of band b: [down0[b], down0[b + 1])
  std::vector<Piece> up_pieces;   // what goes up of the enthalpy before band b may start (local rows / columns of E)
  std::vector<size_t> up0;
  std::vector<FillTask> fills; // u = v = sliding velocity where nothing comes down
  std::vector<ReplTask> repl;  // levels above the cut of what did come down
  ThkMap tmap;
};

static void plan_host_update(const siafd_b200_config &c, int RS, int band, bool sparse, bool cut, bool patch, int cut_cols,
                             int cut_rows, int threads, const double *H, const double *surface, const double *bed, const double *mask,
                             HostPlan &P) {
  P.RS = RS, P.nseg = (c.ym + 2 + RS - 1) / RS, P.band = std::max(1, band), P.NB = (P.nseg + P.band - 1) / P.band;
  P.sparse = sparse, P.cut = sparse && cut, P.patch = patch;
  const int NB = P.NB, nseg = P.nseg, we = c.w_3d_in, wuv = c.w_uv;
  IceExtent ext;
  if (sparse) ice_extent(c, H, ext, patch, threads);
  if (P.cut) thk_map_build(c, H, surface, bed, mask, P.tmap, threads);
  const ThkMap *M = P.cut ? &P.tmap : nullptr;
  // ---- u, v: per band the rectangle that comes down, the rest of the band's rows is the host's to fill ----
  P.down.resize(NB);
  P.down0.assign(NB + 1, 0);
  for (int b = 0; b < NB; ++b) {
    const int s0 = b * P.band, s1 = std::min(nseg, (b + 1) * P.band);
    // (a patch: owned columns only -- the ghost columns are the neighbours' to fill and come down at the end)
    HostPlan::Rect R{std::max(0, s0 * RS - 1), std::min(c.ym, s1 * RS - 1), patch ? wuv : 0, patch ? c.xm + wuv : c.xm + 2 * wuv};
    if (sparse && R.o1 > R.o0) {
      int lo, hi;
      bool whole;
      const bool any = band_extent(ext, c.ym, R.o0, R.o1, &lo, &hi, &whole, patch);
      if (!any) {
        R.c0 = R.c1 = 0;
      } else if (!whole) {
        R.c0 = lo + wuv, R.c1 = hi + 1 + wuv;
      }
      // (a patch: the host fills owned columns only; the ghost columns come down from the device at the end)
      const int f0 = patch ? wuv : 0, f1 = patch ? c.xm + wuv : c.xm + 2 * wuv;
      for (int j = R.o0; j < R.o1; ++j) {
        if (R.c0 >= R.c1) {
          P.fills.push_back({j, f0, f1});
        } else {
          if (R.c0 > f0) P.fills.push_back({j, f0, R.c0});
          if (R.c1 < f1) P.fills.push_back({j, R.c1, f1});
        }
      }
    }
    P.down[b] = R;
    const size_t first = P.down_pieces.size();
    // (cut_rows > 0: chunks of that many rows instead of the band's -- a row's own maximum cuts lower, but the copies
    // get smaller; one row = a plain 2D copy)
    const int step = (M && cut_rows > 0) ? cut_rows : std::max(1, R.o1 - R.o0);
    for (int r = R.o0; r < R.o1; r += step) {
      cut_pieces(c, M, cut_cols, wuv, wuv + r, wuv + std::min(R.o1, r + step), R.c0, R.c1, P.down_pieces);
    }
    P.down0[b + 1] = P.down_pieces.size();
    for (size_t q = first; q < P.down_pieces.size(); ++q) {
      const Piece &D = P.down_pieces[q];
      if (D.n >= c.Mz) continue;
      // (tasks of at most 16 rows, so that the threads share a band evenly and its last task is short)
      for (int r = D.r0; r < D.r1; r += 16) {
        Piece t = D;
        t.r0 = r, t.r1 = std::min(D.r1, r + 16);
        P.repl.push_back({b, t});
      }
    }
  }
  // ---- enthalpy: band b reads the local rows below (b + 1) band RS + w_3d_in (+1 of slack) ----
  const long rowsE = c.ym + 2 * we;
  P.up0.assign(NB + 1, 0);
  long up0 = 0;
  for (int b = 0; b < NB; ++b) {
    const long up1 = (b == NB - 1) ? rowsE : std::min<long>(rowsE, (long)std::min(nseg, (b + 1) * P.band) * RS + we + 1);
    if (up1 > up0) {
      int lo = 0, hi = 0;
      bool whole = true;
      const bool any = !sparse || band_extent(ext, c.ym, (int)up0 - we, (int)up1 - we, &lo, &hi, &whole, patch);
      if (any) {
        // local columns [c0, c1) of the enthalpy array; a patch: a range that touches an edge takes the ghost columns too
        int c0 = 0, c1 = c.xm + 2 * we;
        if (sparse && !whole) c0 = (patch && lo <= 0) ? 0 : lo + we, c1 = (patch && hi >= c.xm - 1) ? c.xm + 2 * we : hi + 1 + we;
        const int step = (M && cut_rows > 0) ? cut_rows : (int)(up1 - up0);
        for (int r = (int)up0; r < (int)up1; r += step) {
          cut_pieces(c, M, cut_cols, we, r, std::min((int)up1, r + step), c0, c1, P.up_pieces);
        }
      }
    }
    up0 = std::max(up0, up1);
    P.up0[b + 1] = P.up_pieces.size();
  }
}

static inline int64_t piece_bytes(const Piece &p) { return (int64_t)(p.c1 - p.c0) * p.n * 8 * (p.r1 - p.r0); }

// one piece between two arrays of rows of `row_cells` columns with the same local layout: what copy_piece
// (siafd_capi.cu) asks the copy engine for, done with memcpy
static void copy_piece_host(double *dst, const double *src, long row_cells, int Mz, const Piece &p) {
  for (int r = p.r0; r < p.r1; ++r) {
    for (int cc = p.c0; cc < p.c1; ++cc) {
      const long off = ((long)r * row_cells + cc) * Mz;
      std::memcpy(dst + off, src + off, (size_t)p.n * sizeof(double));
    }
  }
}

} // namespace siafd_hostplan
