// siafd_kernels.cuh -- launch interface between the C ABI (siafd_capi.cu) and the kernels.
#pragma once
#include "siafd_device.cuh"

namespace siafd {

struct PeerPush;

// Device pointers of one handle (PISM local ghosted layout, see include/siafd_b200.h).
struct Fields {
  const double *h, *H, *mask, *bed, *E, *age, *sliding;
  const double *topgsmooth, *maxtl, *C2, *C3, *C4;
  double *thk_smooth, *theta, *w_i, *w_j;
  double *h_x, *h_y, *D, *Q, *u, *v;
  const double *z;           // Mz levels
  unsigned *err;             // error bits (EB_*)
  unsigned long long *dmax;  // bit pattern of max D (D >= 0, so unsigned order == double order)
  int *hdc;                  // high_diffusivity_counter
  int *segw;                 // [0, 128): icy points per row segment; [128, 256): segments, heaviest first
  unsigned *segdone;         // CTAs of the 2D pass that have finished (the last one sorts)
};

struct Tuning {
  int rows_per_cta;  // rows of the extended patch one CTA marches over
  int use_bulk_copy; // 1: cp.async.bulk + mbarrier row loads; 0: 8-byte cp.async
  int skip_ice_free; // 1: do not load enthalpy rows no staggered point needs
  int wz;            // z ranges per column (2, 4 or 8)
  int pipeline_host; // 1: siafd_b200_update with host arrays overlaps upload, kernel and download over row bands
  int pipeline_band; // row segments per band of that pipeline
  int sparse_host;   // 1: that pipeline moves only the parts of the 3D arrays that are within 3 cells of ice
  int level_cut;     // 1: ... and of those columns only the levels up to the thickest ice nearby (u, v are constant
                     // above the surface, the enthalpy is not read there); the host replicates the top value
  int cut_cols;      // columns per chunk of a band that share one cut level
  int cut_rows;      // rows per chunk (0: the rows of the band)
  int graph_step;    // 1: siafd_b200_update_decomposed replays a captured CUDA graph of the step
  int order_segments; // 1: the fused kernel takes its row segments heaviest (most icy points) first: short tail
};

// number of kernel launches each call makes is returned (for gpu_launches accounting)
int launch_prep2d(const DP &P, const Fields &F, cudaStream_t s);
// with_prep2d (haseloff only): thk_smooth / theta of launch_prep2d in the same pass; returns the launches made.
// gradient_ring_is_local: that pass also evaluates the cross components on the ring of ghost points around the patch
// (geometry ghosts two cells wide), so that the ghost update of sia/SIAFD.cc:498-499 has nothing left to bring
inline bool gradient_ring_is_local(const DP &P) { return P.grad == GRAD_HASELOFF && P.wg >= 2 && P.wst == 1; }
int launch_gradient(const DP &P, const Fields &F, cudaStream_t s, const PeerPush *push = nullptr, bool with_prep2d = false);
// one launch covers the row segments [seg0, seg0 + nseg) of the extended patch (nseg < 0: all from seg0)
// seg_order: device array of nseg segment indices to take in blockIdx.y order (whole-patch launches only), or NULL
int launch_slab(const DP &P, const Fields &F, bool full, const Tuning &T, long nE, long n2, double inv_dz, int seg0,
                int nseg, cudaStream_t s, const PeerPush *push = nullptr, const int *seg_order = nullptr);
int slab_rows_per_segment(const Tuning &T);
int slab_segments(const DP &P, const Tuning &T);
size_t slab_smem_need(const DP &P, bool full, bool bulk); // shared memory of the smallest configuration

// copy a rectangle of cells between two [rows][cells][dof] arrays (ghost wrap, halo pack/unpack)
int launch_copy_region(double *dst, long dst_row_cells, int dst_i0, int dst_j0, const double *src, long src_row_cells,
                       int src_i0, int src_j0, int width_cells, int height_cells, int dof, cudaStream_t s);

// Pieces of u and v stored straight into the caller's mapped pinned HOST arrays (same local layout as the device's):
// rows [r0, r1) x columns [c0, c1) x levels [0, n) of each piece.  The host-array update's alternative to asking the
// copy engine for strided lines (siafd_capi.cu::copy_piece).
struct StorePiece {
  int r0, r1, c0, c1, n;
};
int launch_store_pieces(const double *u, const double *v, double *host_u, double *host_v, const StorePiece *pieces_dev, int p0,
                        int p1, long row_cells, int Mz, cudaStream_t s);

// Peer halo exchange (CUDA IPC mapped neighbour arrays): one kernel copies every strip of a phase straight into
// the neighbours' ghost cells over NVLink, a second one raises the neighbours' arrival counters, a third waits
// for this rank's own counters.
struct HaloDesc {
  const double *src;
  double *dst;
  long src_row_cells, dst_row_cells;
  int src_i0, src_j0, dst_i0, dst_j0, wc, hc, dof, pad;
};
struct HaloBatch {
  HaloDesc d[48];
  int n;
};
struct HaloSignal {
  unsigned long long *slot[8]; // where to write, one per neighbour direction
  unsigned long long value;
};
int launch_halo_push(const HaloBatch &B, cudaStream_t s);
int launch_halo_signal(const HaloSignal &S, cudaStream_t s);
int launch_halo_wait(const unsigned long long *slots8, unsigned long long value, cudaStream_t s);

// ---- communicator of a decomposed run (siafd_comm.cu): one process (or handle) per GPU of one node ---------------
// Every rank owns a pad in device memory that all ranks map.  Ghost updates are direct stores into the neighbours'
// arrays; the pad orders them (arrival counters, one row per phase) and carries the small all-rank reductions
// (D_max, error bits, counters: SIAFD.cc:748-750; CFL: timestepping.cc).  The expected counter values live in the
// pad itself, so a captured CUDA graph of a step replays unchanged.
constexpr int COMM_MAXR = 32;   // ranks of one communicator
constexpr int COMM_PHASES = 8;  // independent arrival-counter rows
struct CommPad {
  unsigned long long arrive[COMM_PHASES][8];   // [phase][direction the sender is seen in], written by the neighbours
  unsigned long long red_flag[4][COMM_MAXR];   // [buffer][rank]
  double red_val[4][COMM_MAXR][8];             // [buffer][rank][value]
  // local bookkeeping, never written by a peer
  unsigned long long step[COMM_PHASES], red_step[2];
  unsigned int done[COMM_PHASES];
  unsigned int timed_out; // a wait gave up (EB_COMM)
};
struct CommPeers { // device-resident, read-only after setup
  CommPad *self;
  CommPad *nb[8];          // the eight neighbours' pads (== self where the neighbour is this rank)
  CommPad *all[COMM_MAXR]; // every rank's pad
  int rank, size;
};
// Ghost update fused into the producing kernel: the owned cells within `w` of the patch edge are also stored into the
// neighbours' ghost cells (peer memory; the own array where the neighbour is this rank: the periodic self-wrap).
// Neighbour-local index of my local cell (il, jl): (jl + dj[d]) * rowc[d] + (il + di[d]).
struct PeerPush {
  double *a[8], *b[8];
  long rowc[8];
  int di[8], dj[8];
  int w, on;
};
__host__ __device__ inline bool peer_strip_member(int d, bool W_, bool E_, bool S_, bool N_) {
  // dir = 0..7 <-> (dx,dy) = (-1,-1),(0,-1),(1,-1),(-1,0),(1,0),(-1,1),(0,1),(1,1)
  const int dx = (d == 0 || d == 3 || d == 5) ? -1 : ((d == 2 || d == 4 || d == 7) ? 1 : 0);
  const int dy = (d < 3) ? -1 : (d > 4 ? 1 : 0);
  return (dx == 0 || (dx < 0 ? W_ : E_)) && (dy == 0 || (dy < 0 ? S_ : N_));
}
// copy the strips of a phase, then (last CTA) signal the neighbours and wait for theirs
int launch_halo_xchg(const HaloBatch &B, const CommPeers *cp, int phase, int signal, cudaStream_t s);
// after a kernel with fused pushes: signal + wait of one phase
int launch_comm_sync(const CommPeers *cp, int phase, cudaStream_t s);
// end of an update: signal + wait of `phase` (u, v ghosts) and the all-rank reduction of {D_max, error bits,
// high-diffusivity counter}; result {max, or, sum} to res_dev[0..2] (as 64-bit patterns) and, if given, res_host
int launch_comm_final(const CommPeers *cp, int phase, unsigned long long *dmax, unsigned *err, int *hdc,
                      unsigned long long *res_dev, unsigned long long *res_host, cudaStream_t s);
// all-rank reduction of n <= 8 doubles (op 0: max, 1: min, 2: sum in rank order); vals_dev in/out
int launch_comm_allreduce(const CommPeers *cp, double *vals_dev, double *vals_host, int n, int op, cudaStream_t s);

// StressBalance::compute_vertical_velocity (stressbalance/StressBalance.cc:283-424)
int launch_vertical_velocity(const DP &P, const double *mask, const double *u, const double *v, const double *bmr,
                             int upstream, const double *z, double *w, cudaStream_t s);
// the marching kernel with the 3D CFL maxima fused in (siafd_mass.cu); 0 = not applicable (Mz > 256)
int launch_vvel_march(const DP &P, const double *mask, const double *thk, const double *u, const double *v,
                      const double *bmr, int upstream, const double *z, double *w, unsigned long long *cfl,
                      unsigned *err, int rows_per_cta, cudaStream_t s);
int launch_vvel_slab(const DP &P, const double *mask, const double *thk, const double *u, const double *v,
                     const double *bmr, int upstream, const double *z, double *w, unsigned long long *cfl,
                     unsigned *err, int rows_per_cta, int wz, long nUV, double inv_dz, cudaStream_t s);
// StressBalance::compute_volumetric_strain_heating; -1 = flow law without a softness (gk) or Mz > 256
int launch_strain_heating(const DP &P, int law, double n, double e, const double *mask, const double *thk,
                          const double *E, const double *u, const double *v, const double *z, double *sigma,
                          unsigned *err, cudaStream_t s);
int launch_regional_override(const DP &P, const double *no_model, const double *hx_nm, const double *hy_nm, double *h_x,
                             double *h_y, cudaStream_t s);
// IceModelVec3::getSurfaceValues / getHorSlice (util/iceModelVec3.cc:153-240): a 3D field (ghost width wa) at the
// height zq[i,j] (2D, ghost width wz) or, with zq == NULL, at the height z0; out is [ym][xm] without ghosts
int launch_value_at_height(const DP &P, const double *a, int wa, const double *zq, int wz, double z0, const double *z,
                           double *out, cudaStream_t s);
// SURVEY.md 8(f) N1 / N3-CFL (siafd_mass.cu): GeometryEvolution flow and source steps, Geometry::ensure_consistency,
// max_timestep_cfl_3d / _2d.  NULL for an optional field means "all zero".
int launch_mass_flow(const DP &P, double dt, const double *H, const double *bed, const double *sea, const double *vel,
                     const double *vel_bc, const double *thk_bc, const double *Q, double *flux_div, double *dH,
                     double *cons_err, cudaStream_t s);
int launch_mass_apply(const DP &P, double *H, const double *dH, cudaStream_t s);
int launch_mass_source(const DP &P, double dt, double ice_density, int use_bmr, double *H, const double *mask,
                       const double *thk_bc, const double *smb, const double *bmr, double *eff_smb, double *eff_bmb,
                       cudaStream_t s);
int launch_consistency(const DP &P, long n, const double *sea, const double *bed, const double *thk, double *mask_out,
                       double *surf_out, unsigned *err, cudaStream_t s);
int launch_cfl(const DP &P, bool do3d, const double *thk, const double *mask, const double *u, const double *v,
               const double *w, const double *z, const double *vel, unsigned long long *out, unsigned *err,
               cudaStream_t s);
int launch_geometry(const DP &P, long n, const double *sea_level, const double *bed, const double *thk, double *mask_out,
                    double *surf_out, cudaStream_t s);
int launch_flow_n(const DP &P, long n, const double *stress, const double *E, const double *p, const double *gs,
                  double *out, cudaStream_t s);
// BedSmoother::preprocess_bed for this patch (+wg ghosts) from the global bed on device
int launch_preprocess_bed(const DP &P, const double *global_bed, int Nx, int Ny, double *topgsmooth, double *maxtl,
                          double *C2, double *C3, double *C4, cudaStream_t s);

} // namespace siafd
