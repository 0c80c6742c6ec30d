// siafd_comm.cu -- decomposed runs without any host-side communication library: the communicator of
// include/siafd_b200.h ("siafd_b200_comm_*") and the whole-step entry point siafd_b200_update_decomposed.
//
// Reference: PISM's ghost updates are DMLocalToLocalBegin/End on a periodic BOX-stencil DMDA
// (util/iceModelVec.cc:630-643, util/IceGrid.cc:863-885); SIAFD::update communicates at SIAFD.cc:498-499 (h_x, h_y),
// :946-947 (u, v) and reduces D_max / the high-diffusivity counter at :748-750; ParallelSection
// (util/error_handling.cc:189-214) makes every rank fail together.
//
// Here: one process (or one handle) per GPU of one NVLink node.  Every rank maps its neighbours' field arrays and
// every rank's pad (CUDA IPC between processes; plain pointers inside one process).  A ghost update is a set of
// direct stores into the neighbours' ghost cells -- issued by the kernel that PRODUCES the values where there is one
// (gradient kernel: h_x, h_y; fused kernel: u, v), by a strip-copy kernel for inputs -- followed by arrival counters
// in the pad.  D_max, the error bits and the counter travel through the same pad: every rank stores its 24 bytes into
// every rank's pad and reduces the `size` entries it received, so that the status is collective by construction.
// A step is 4-6 launches, captured once into a CUDA graph and replayed.
#include "siafd_handle.cuh"

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <thread>
#include <unistd.h>

using namespace siafd;
using namespace siafd_host;

namespace siafd {

__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
// bounded spin: a peer that never arrives raises EB_COMM instead of hanging the GPU.  The bound is on elapsed time
// (about half a minute of the global nanosecond timer), not on iterations: ranks legitimately reach a ghost update
// seconds apart -- e.g. when the hosts of some are still pinning gigabytes of memory (an iteration bound of ~2.5 s made
// the first host-buffer call of an 8-rank run fail on a busy host).
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void wait_at_least(const unsigned long long *p, unsigned long long v, CommPad *self) {
  unsigned long long t0 = 0ull;
  for (long it = 0; ld_acquire_sys(p) < v; ++it) {
    if (it >= 128) __nanosleep(it < 1024 ? 32 : 512); // (the first polls back to back: a neighbour is rarely far behind)
    if ((it & 0xffff) == 0xffff) { // every ~30 ms: look at the clock
      const unsigned long long t = globaltimer_ns();
      if (t0 == 0ull) {
        t0 = t;
      } else if (t - t0 > 30ull * 1000000000ull) {
        self->timed_out = 1u;
        return;
      }
    }
  }
}

// lanes 0..7 of the calling warp: raise the neighbours' arrival counters of `phase` to this rank's next step, then
// wait until all eight neighbours have raised this rank's.  Everything this rank stored before (the caller fences)
// has landed when a neighbour sees the counter move.
__device__ __forceinline__ void signal_and_wait(const CommPeers *cp, int phase) {
  const int lane = threadIdx.x & 31;
  CommPad *self = cp->self;
  unsigned long long v = 0;
  if (lane == 0) {
    v = self->step[phase] + 1;
    self->step[phase] = v;
  }
  v = __shfl_sync(0xffffffffu, v, 0);
  if (lane < 8) {
    st_release_sys(&cp->nb[lane]->arrive[phase][7 - lane], v); // the neighbour sees me in direction 7 - d
    wait_at_least(&self->arrive[phase][lane], v, self);
  }
  __syncwarp();
  __threadfence_system();
}

// strips of a phase -> the neighbours' ghost cells (blockIdx.y = strip); the last CTA to finish signals and waits
__global__ void k_halo_xchg(const __grid_constant__ HaloBatch B, const CommPeers *cp, int phase, int do_signal) {
  const HaloDesc &D = B.d[blockIdx.y];
  const long rowlen = (long)D.wc * D.dof;
  const long n = rowlen * D.hc;
  for (long q = (long)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (long)gridDim.x * blockDim.x) {
    const long jj = q / rowlen, e = q - jj * rowlen;
    D.dst[((D.dst_j0 + jj) * D.dst_row_cells + D.dst_i0) * D.dof + e] =
        D.src[((D.src_j0 + jj) * D.src_row_cells + D.src_i0) * D.dof + e];
  }
  if (!do_signal) return;
  __threadfence_system();
  __syncthreads();
  __shared__ int last;
  if (threadIdx.x == 0) {
    const unsigned total = gridDim.x * gridDim.y;
    const unsigned prev = atomicAdd(&cp->self->done[phase], 1u);
    last = (prev == total - 1);
    if (last) cp->self->done[phase] = 0;
  }
  __syncthreads();
  if (last && threadIdx.x < 32) {
    __threadfence_system();
    signal_and_wait(cp, phase);
  }
}

__global__ void k_comm_sync(const CommPeers *cp, int phase) {
  __threadfence_system();
  signal_and_wait(cp, phase);
}

// all-rank exchange of up to 8 doubles per rank through the pads: double-buffered by step parity, so that a rank
// that runs ahead never overwrites values a slower rank has yet to read (it cannot be two reductions ahead: the
// next one needs everybody's contribution to this one)
__device__ __forceinline__ void publish_and_collect(const CommPeers *cp, int channel, const double (&mine)[8], int n,
                                                    double (&got)[8], bool &have) {
  const int lane = threadIdx.x & 31;
  CommPad *self = cp->self;
  unsigned long long v = 0;
  if (lane == 0) {
    v = self->red_step[channel] + 1;
    self->red_step[channel] = v;
  }
  v = __shfl_sync(0xffffffffu, v, 0);
  const int buf = 2 * channel + (int)(v & 1ull);
  have = lane < cp->size;
  if (have) {
    CommPad *dst = cp->all[lane];
    for (int q = 0; q < n; ++q) dst->red_val[buf][cp->rank][q] = mine[q];
    __threadfence_system();
    st_release_sys(&dst->red_flag[buf][cp->rank], v);
    wait_at_least(&self->red_flag[buf][lane], v, self);
    for (int q = 0; q < n; ++q) got[q] = *(volatile double *)&self->red_val[buf][lane][q];
  }
  __syncwarp();
}

__global__ void k_comm_final(const CommPeers *cp, int phase, unsigned long long *dmax, unsigned *err, int *hdc,
                             unsigned long long *res_dev, unsigned long long *res_host) {
  const int lane = threadIdx.x;
  __threadfence_system();
  if (phase >= 0) signal_and_wait(cp, phase); // u, v ghosts (SIAFD.cc:946-947)
  double mine[8] = {0, 0, 0, 0, 0, 0, 0, 0}, got[8];
  unsigned long long e = 0;
  if (lane == 0) { // sticky until read (see siafd_b200_finish)
    e = (unsigned long long)atomicExch(err, 0u) | (atomicExch(&cp->self->timed_out, 0u) ? (unsigned long long)EB_COMM : 0ull);
  }
  e = __shfl_sync(0xffffffffu, e, 0);
  mine[0] = __longlong_as_double((long long)*dmax); // D >= 0
  mine[1] = __longlong_as_double((long long)e);
  mine[2] = (double)*hdc;
  bool have;
  if (cp->size == 1) { // one rank: nothing to exchange
    have = lane == 0;
    got[0] = mine[0], got[1] = mine[1], got[2] = mine[2];
  } else {
    publish_and_collect(cp, 0, mine, 3, got, have);
  }
  double m = have ? got[0] : 0.0;
  unsigned long long bits = have ? (unsigned long long)__double_as_longlong(got[1]) : 0ull;
  double cnt = have ? got[2] : 0.0;
  for (int d = 16; d >= 1; d >>= 1) {
    m = fmax(m, __shfl_xor_sync(0xffffffffu, m, d));
    bits |= __shfl_xor_sync(0xffffffffu, bits, d);
    cnt += __shfl_xor_sync(0xffffffffu, cnt, d); // (integers far below 2^53: exact in any order)
  }
  if (lane == 0) {
    const unsigned long long r0 = (unsigned long long)__double_as_longlong(m), r2 = (unsigned long long)(long long)cnt;
    res_dev[0] = r0, res_dev[1] = bits, res_dev[2] = r2;
    if (res_host != nullptr) {
      res_host[0] = r0, res_host[1] = bits, res_host[2] = r2;
      __threadfence_system();
    }
  }
}

__global__ void k_comm_allreduce(const CommPeers *cp, double *vals, double *vals_host, int n, int op) {
  const int lane = threadIdx.x;
  double mine[8], got[8];
  for (int q = 0; q < 8; ++q) mine[q] = q < n ? vals[q] : 0.0;
  bool have;
  publish_and_collect(cp, 1, mine, n, got, have);
  // every lane holds one rank's values: reduce in rank order (sum) / any order (max, min)
  for (int q = 0; q < n; ++q) {
    double r = 0.0;
    if (op == 2) {
      for (int k = 0; k < cp->size; ++k) r += __shfl_sync(0xffffffffu, got[q], k);
    } else {
      r = have ? got[q] : (op == 0 ? -INFINITY : INFINITY);
      for (int d = 16; d >= 1; d >>= 1) {
        const double o = __shfl_xor_sync(0xffffffffu, r, d);
        r = (op == 0) ? fmax(r, o) : fmin(r, o);
      }
    }
    if (lane == 0) {
      vals[q] = r;
      if (vals_host != nullptr) vals_host[q] = r;
    }
  }
  if (lane == 0 && vals_host != nullptr) __threadfence_system();
}

int launch_halo_xchg(const HaloBatch &B, const CommPeers *cp, int phase, int signal, cudaStream_t s) {
  if (B.n <= 0) return 0;
  long nmax = 1;
  for (int q = 0; q < B.n; ++q) nmax = std::max(nmax, (long)B.d[q].wc * B.d[q].dof * B.d[q].hc);
  const unsigned bx = (unsigned)std::min<long>((nmax + 255) / 256, 64);
  k_halo_xchg<<<dim3(bx, (unsigned)B.n), 256, 0, s>>>(B, cp, phase, signal);
  return 1;
}
int launch_comm_sync(const CommPeers *cp, int phase, cudaStream_t s) {
  k_comm_sync<<<1, 32, 0, s>>>(cp, phase);
  return 1;
}
int launch_comm_final(const CommPeers *cp, int phase, unsigned long long *dmax, unsigned *err, int *hdc,
                      unsigned long long *res_dev, unsigned long long *res_host, cudaStream_t s) {
  k_comm_final<<<1, 32, 0, s>>>(cp, phase, dmax, err, hdc, res_dev, res_host);
  return 1;
}
int launch_comm_allreduce(const CommPeers *cp, double *vals_dev, double *vals_host, int n, int op, cudaStream_t s) {
  k_comm_allreduce<<<1, 32, 0, s>>>(cp, vals_dev, vals_host, n, op);
  return 1;
}

} // namespace siafd

// =====================================================================================================
// host side
// =====================================================================================================
namespace {

const int HDX[8] = {-1, 0, 1, -1, 1, -1, 0, 1}, HDY[8] = {-1, -1, -1, 0, 0, 1, 1, 1};

// fields whose ghosts a decomposed run updates (allocated and mapped at comm_init)
const int COMM_FIELDS[] = {SIAFD_B200_F_SURFACE, SIAFD_B200_F_THICKNESS, SIAFD_B200_F_MASK,    SIAFD_B200_F_BED,
                           SIAFD_B200_F_ENTHALPY, SIAFD_B200_F_AGE,      SIAFD_B200_F_SLIDING, SIAFD_B200_F_H_X,
                           SIAFD_B200_F_H_Y,     SIAFD_B200_F_U,         SIAFD_B200_F_V,       SIAFD_B200_F_SEA_LEVEL,
                           SIAFD_B200_F_H_X_NO_MODEL, SIAFD_B200_F_H_Y_NO_MODEL, SIAFD_B200_F_VEL_BC_MASK,
                           SIAFD_B200_F_THK_BC_MASK,  SIAFD_B200_F_NO_MODEL_MASK, SIAFD_B200_F_NO_MODEL_SURFACE};
const int N_COMM_FIELDS = (int)(sizeof(COMM_FIELDS) / sizeof(COMM_FIELDS[0]));

struct RankInfo { // what a rank tells the others
  int32_t xs, xm, ys, ym;
  int32_t have[SIAFD_B200_F_COUNT];                     // field is exported
  unsigned char field_handle[SIAFD_B200_F_COUNT][64];   // cudaIpcMemHandle_t (multi-process)
  unsigned char pad_handle[64];
  uint64_t field_ptr[SIAFD_B200_F_COUNT], pad_ptr;      // plain pointers (same process)
  int32_t device;
};

bool comm_field_wanted(const siafd_b200_handle *h, int f) {
  if (f == SIAFD_B200_F_AGE) return h->P.use_age != 0;
  // the optional fields take part once the caller has touched them (upload / device_ptr) before comm_init
  const bool optional = f == SIAFD_B200_F_SEA_LEVEL || f == SIAFD_B200_F_H_X_NO_MODEL || f == SIAFD_B200_F_H_Y_NO_MODEL ||
                        f == SIAFD_B200_F_VEL_BC_MASK || f == SIAFD_B200_F_THK_BC_MASK ||
                        f == SIAFD_B200_F_NO_MODEL_MASK || f == SIAFD_B200_F_NO_MODEL_SURFACE;
  return !optional || h->buf[f] != nullptr;
}

int wrap(int v, int n) { return ((v % n) + n) % n; }

// the rank that owns global cell (i, j)
int owner_of(const std::vector<RankInfo> &R, int i, int j) {
  for (size_t q = 0; q < R.size(); ++q) {
    if (i >= R[q].xs && i < R[q].xs + R[q].xm && j >= R[q].ys && j < R[q].ys + R[q].ym) return (int)q;
  }
  return -1;
}

int fill_my_info(siafd_b200_handle *h, RankInfo &I, bool ipc) {
  std::memset(&I, 0, sizeof(I));
  const siafd_b200_config &c = h->cfg;
  I.xs = c.xs, I.xm = c.xm, I.ys = c.ys, I.ym = c.ym, I.device = h->device;
  for (int q = 0; q < N_COMM_FIELDS; ++q) {
    const int f = COMM_FIELDS[q];
    if (!comm_field_wanted(h, f)) continue;
    int st = ensure(h, f);
    if (st) return st;
    if (!h->owned[f]) {
      return fail(h, SIAFD_B200_ERR_BAD_ARGUMENT,
                  "field %d is bound to caller memory: a decomposed run needs handle-owned storage for ghosted fields", f);
    }
    I.have[f] = 1;
    I.field_ptr[f] = (uint64_t)(uintptr_t)h->buf[f];
  }
  {
    // everything siafd_b200_update_decomposed touches exists before the first step (no allocation between the launches of
    // ranks that wait for each other)
    const int need[] = {SIAFD_B200_F_W_I,   SIAFD_B200_F_W_J, SIAFD_B200_F_TOPGSMOOTH, SIAFD_B200_F_MAXTL,
                        SIAFD_B200_F_C2,    SIAFD_B200_F_C3,  SIAFD_B200_F_C4,         SIAFD_B200_F_THK_SMOOTH,
                        SIAFD_B200_F_THETA, SIAFD_B200_F_D,   SIAFD_B200_F_FLUX};
    for (int f : need) {
      int st = ensure(h, f);
      if (st) return st;
    }
  }
  if (!h->comm.pad) {
    // (2 MiB: its own allocation block, so that the IPC handle names nothing else)
    CU(h, cudaMalloc(&h->comm.pad, (size_t)2 << 20));
    CU(h, cudaMemset(h->comm.pad, 0, (size_t)2 << 20));
  }
  I.pad_ptr = (uint64_t)(uintptr_t)h->comm.pad;
  CU(h, cudaStreamSynchronize(h->stream)); // zero-fills of fresh buffers
  if (ipc) {
    for (int f = 0; f < SIAFD_B200_F_COUNT; ++f) {
      if (I.have[f]) CU(h, cudaIpcGetMemHandle(reinterpret_cast<cudaIpcMemHandle_t *>(I.field_handle[f]), h->buf[f]));
    }
    CU(h, cudaIpcGetMemHandle(reinterpret_cast<cudaIpcMemHandle_t *>(I.pad_handle), h->comm.pad));
  }
  return SIAFD_B200_OK;
}

// with everybody's info and the mapped pointers at hand: neighbours, descriptors, device-side tables
int finish_setup(siafd_b200_handle *h, int rank, const std::vector<RankInfo> &R,
                 const std::vector<std::vector<void *>> &field_of_rank, const std::vector<void *> &pad_of_rank) {
  const siafd_b200_config &c = h->cfg;
  siafd_b200_handle::Comm &C = h->comm;
  const int size = (int)R.size();
  C.rank = rank, C.size = size;
  // neighbours by coordinates (periodic in both directions, util/IceGrid.cc:870-872)
  for (int d = 0; d < 8; ++d) {
    const int ci = wrap(HDX[d] < 0 ? c.xs - 1 : (HDX[d] > 0 ? c.xs + c.xm : c.xs), c.Mx);
    const int cj = wrap(HDY[d] < 0 ? c.ys - 1 : (HDY[d] > 0 ? c.ys + c.ym : c.ys), c.My);
    const int nb = owner_of(R, ci, cj);
    if (nb < 0) return fail(h, SIAFD_B200_ERR_BAD_ARGUMENT, "comm_init: no rank owns cell (%d, %d): the patches do not tile the grid", ci, cj);
    if ((HDX[d] == 0 && (R[nb].xs != c.xs || R[nb].xm != c.xm)) || (HDY[d] == 0 && (R[nb].ys != c.ys || R[nb].ym != c.ym))) {
      return fail(h, SIAFD_B200_ERR_BAD_ARGUMENT, "comm_init: the decomposition is not a tensor product of x and y ranges");
    }
    C.nb[d] = nb;
    for (int f = 0; f < SIAFD_B200_F_COUNT; ++f) {
      siafd_b200_handle::Peer &P = h->peers[f][d];
      P.attached = R[rank].have[f] && R[nb].have[f];
      P.base = P.attached ? (double *)field_of_rank[nb][f] : nullptr;
      P.xm = R[nb].xm, P.ym = R[nb].ym;
    }
  }
  CommPeers hp;
  std::memset(&hp, 0, sizeof(hp));
  hp.self = (CommPad *)C.pad;
  hp.rank = rank, hp.size = size;
  for (int d = 0; d < 8; ++d) hp.nb[d] = (CommPad *)pad_of_rank[C.nb[d]];
  for (int q = 0; q < size; ++q) hp.all[q] = (CommPad *)pad_of_rank[q];
  if (!C.d_peers) CU(h, cudaMalloc(&C.d_peers, sizeof(CommPeers)));
  CU(h, cudaMemcpy(C.d_peers, &hp, sizeof(hp), cudaMemcpyHostToDevice));
  if (!C.d_res) {
    CU(h, cudaMalloc(&C.d_res, 4 * sizeof(unsigned long long)));
    CU(h, cudaMallocHost(&C.h_res, 4 * sizeof(unsigned long long)));
    CU(h, cudaMalloc(&C.d_red, 8 * sizeof(double)));
    CU(h, cudaMallocHost(&C.h_red, 8 * sizeof(double)));
  }
  // everything the host-buffer pipeline (siafd_b200_update with host arrays) creates lazily: now, so that no rank makes
  // a (possibly device-synchronising) allocation while its neighbours already wait for it inside a kernel
  if (!h->s_up) {
    CU(h, cudaStreamCreateWithFlags(&h->s_up, cudaStreamNonBlocking));
    CU(h, cudaStreamCreateWithFlags(&h->s_dn, cudaStreamNonBlocking));
  }
  {
    const int nseg = slab_segments(h->P, h->tuning);
    while ((int)h->ev_pipe.size() < 2 * nseg + 4) {
      cudaEvent_t e;
      CU(h, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      h->ev_pipe.push_back(e);
    }
  }
  if (!C.s_aux) {
    CU(h, cudaStreamCreateWithFlags(&C.s_aux, cudaStreamNonBlocking));
    CU(h, cudaEventCreateWithFlags(&C.ev_fork, cudaEventDisableTiming));
    CU(h, cudaEventCreateWithFlags(&C.ev_join, cudaEventDisableTiming));
  }
  C.active = true;
  C.graph_valid[0] = C.graph_valid[1] = C.graph_valid[2] = C.graph_valid[3] = false;
  return SIAFD_B200_OK;
}

} // namespace

// fused-push table of two same-shaped fields (h_x / h_y, u / v) with ghost width W, strips of width w
void siafd_host::comm_make_push(const siafd_b200_handle *h, int fa, int fb, int W, int w, PeerPush &PP) {
  std::memset(&PP, 0, sizeof(PP));
  const siafd_b200_config &c = h->cfg;
  for (int d = 0; d < 8; ++d) {
    const siafd_b200_handle::Peer &A = h->peers[fa][d], &B = h->peers[fb][d];
    PP.a[d] = A.base, PP.b[d] = B.base;
    PP.rowc[d] = A.xm + 2 * W;
    PP.di[d] = HDX[d] > 0 ? -c.xm : (HDX[d] < 0 ? A.xm : 0);
    PP.dj[d] = HDY[d] > 0 ? -c.ym : (HDY[d] < 0 ? A.ym : 0);
  }
  PP.w = w, PP.on = 1;
}

namespace {

// descriptors of the eight strips of field f going to the neighbours
int strips_to_peers(siafd_b200_handle *h, int f, int w, HaloBatch &B) {
  const siafd_b200_config &c = h->cfg;
  const FieldMeta m = meta(c, f);
  if (m.width < 1 || w < 1 || w > m.width) return fail(h, SIAFD_B200_ERR_BAD_ARGUMENT, "ghost update: bad width %d for field %d", w, f);
  if (B.n + 8 > 48) return fail(h, SIAFD_B200_ERR_BAD_ARGUMENT, "ghost update: at most 6 fields per call");
  const int W = m.width;
  for (int d = 0; d < 8; ++d) {
    const siafd_b200_handle::Peer &P = h->peers[f][d];
    if (!P.attached) return fail(h, SIAFD_B200_ERR_BAD_ARGUMENT, "ghost update: field %d was not part of the communicator (allocate it before comm_init)", f);
    const int dx = HDX[d], dy = HDY[d];
    HaloDesc &D = B.d[B.n++];
    D.src = (const double *)h->buf[f];
    D.dst = P.base;
    D.src_row_cells = c.xm + 2 * W, D.dst_row_cells = P.xm + 2 * W;
    D.dof = m.dof, D.pad = 0;
    D.wc = dx == 0 ? c.xm : w, D.hc = dy == 0 ? c.ym : w;
    D.src_i0 = W + (dx > 0 ? c.xm - w : 0), D.src_j0 = W + (dy > 0 ? c.ym - w : 0);
    D.dst_i0 = dx > 0 ? W - w : (dx < 0 ? W + P.xm : W);
    D.dst_j0 = dy > 0 ? W - w : (dy < 0 ? W + P.ym : W);
  }
  return SIAFD_B200_OK;
}

bool read_file(const std::string &path, void *dst, size_t n) {
  std::ifstream f(path, std::ios::binary);
  if (!f) return false;
  f.read((char *)dst, (std::streamsize)n);
  return (size_t)f.gcount() == n;
}
bool write_file_atomic(const std::string &path, const void *src, size_t n) {
  const std::string tmp = path + ".tmp";
  {
    std::ofstream f(tmp, std::ios::binary | std::ios::trunc);
    if (!f) return false;
    f.write((const char *)src, (std::streamsize)n);
    if (!f) return false;
  }
  return std::rename(tmp.c_str(), path.c_str()) == 0;
}

} // namespace

void siafd_host::comm_release(siafd_b200_handle *h) {
  siafd_b200_handle::Comm &C = h->comm;
  for (int q = 0; q < 4; ++q) {
    if (C.graph_exec[q]) cudaGraphExecDestroy(C.graph_exec[q]);
    C.graph_exec[q] = nullptr, C.graph_valid[q] = false;
  }
  for (void *p : C.mapped) cudaIpcCloseMemHandle(p);
  C.mapped.clear();
  for (const std::string &f : C.files) unlink(f.c_str());
  C.files.clear();
  if (C.s_aux) cudaStreamDestroy(C.s_aux);
  if (C.ev_fork) cudaEventDestroy(C.ev_fork);
  if (C.ev_join) cudaEventDestroy(C.ev_join);
  cudaFree(C.d_peers), cudaFree(C.d_res), cudaFree(C.d_red), cudaFree(C.pad);
  if (C.h_res) cudaFreeHost(C.h_res);
  if (C.h_red) cudaFreeHost(C.h_red);
  C = siafd_b200_handle::Comm();
}

extern "C" {

int siafd_b200_comm_init(siafd_b200_handle *h, int rank, int size, const char *rendezvous_prefix, double timeout_seconds) {
  if (!h) return null_handle();
  if (size < 1 || size > COMM_MAXR || rank < 0 || rank >= size || !rendezvous_prefix) {
    return fail(h, SIAFD_B200_ERR_BAD_ARGUMENT, "comm_init: rank %d of %d (at most %d ranks), prefix %s", rank, size,
                COMM_MAXR, rendezvous_prefix ? rendezvous_prefix : "NULL");
  }
  CU(h, cudaSetDevice(h->device));
  std::vector<RankInfo> R(size);
  int st = fill_my_info(h, R[rank], true);
  if (st) return st;
  const std::string prefix(rendezvous_prefix);
  auto name = [&](const char *kind, int r) { return prefix + "." + kind + "." + std::to_string(r); };
  if (!write_file_atomic(name("rank", rank), &R[rank], sizeof(RankInfo))) {
    return fail(h, SIAFD_B200_ERR_BAD_ARGUMENT, "comm_init: cannot write %s", name("rank", rank).c_str());
  }
  h->comm.files.push_back(name("rank", rank));
  const auto t0 = std::chrono::steady_clock::now();
  auto timed_out = [&]() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > timeout_seconds;
  };
  for (int q = 0; q < size; ++q) {
    if (q == rank) continue;
    while (!read_file(name("rank", q), &R[q], sizeof(RankInfo))) {
      if (timed_out()) return fail(h, SIAFD_B200_ERR_BAD_ARGUMENT, "comm_init: rank %d did not show up at %s", q, prefix.c_str());
      std::this_thread::sleep_for(std::chrono::milliseconds(2));
    }
  }
  // map every rank's pad, and the field arrays of the ranks that are neighbours
  std::vector<std::vector<void *>> fld(size, std::vector<void *>(SIAFD_B200_F_COUNT, nullptr));
  std::vector<void *> pad(size, nullptr);
  std::vector<bool> is_nb(size, false);
  const siafd_b200_config &c = h->cfg;
  for (int d = 0; d < 8; ++d) {
    const int ci = wrap(HDX[d] < 0 ? c.xs - 1 : (HDX[d] > 0 ? c.xs + c.xm : c.xs), c.Mx);
    const int cj = wrap(HDY[d] < 0 ? c.ys - 1 : (HDY[d] > 0 ? c.ys + c.ym : c.ys), c.My);
    const int nb = owner_of(R, ci, cj);
    if (nb >= 0) is_nb[nb] = true;
  }
  for (int q = 0; q < size; ++q) {
    if (q == rank) {
      pad[q] = h->comm.pad;
      for (int f = 0; f < SIAFD_B200_F_COUNT; ++f) fld[q][f] = R[q].have[f] ? h->buf[f] : nullptr;
      continue;
    }
    cudaIpcMemHandle_t mh;
    std::memcpy(&mh, R[q].pad_handle, sizeof(mh));
    CU(h, cudaIpcOpenMemHandle(&pad[q], mh, cudaIpcMemLazyEnablePeerAccess));
    h->comm.mapped.push_back(pad[q]);
    if (!is_nb[q]) continue;
    for (int f = 0; f < SIAFD_B200_F_COUNT; ++f) {
      if (!R[q].have[f] || !R[rank].have[f]) continue;
      std::memcpy(&mh, R[q].field_handle[f], sizeof(mh));
      CU(h, cudaIpcOpenMemHandle(&fld[q][f], mh, cudaIpcMemLazyEnablePeerAccess));
      h->comm.mapped.push_back(fld[q][f]);
    }
  }
  st = finish_setup(h, rank, R, fld, pad);
  if (st) return st;
  // nobody stores into a peer before everybody has mapped everything
  const char one = 1;
  if (!write_file_atomic(name("ready", rank), &one, 1)) return fail(h, SIAFD_B200_ERR_BAD_ARGUMENT, "comm_init: cannot write the ready file");
  h->comm.files.push_back(name("ready", rank));
  for (int q = 0; q < size; ++q) {
    char b;
    while (!read_file(name("ready", q), &b, 1)) {
      if (timed_out()) return fail(h, SIAFD_B200_ERR_BAD_ARGUMENT, "comm_init: rank %d never became ready", q);
      std::this_thread::sleep_for(std::chrono::milliseconds(2));
    }
  }
  return SIAFD_B200_OK;
}

int siafd_b200_comm_init_local(siafd_b200_handle **hs, int size) {
  if (!hs || size < 1 || size > COMM_MAXR) return fail(nullptr, SIAFD_B200_ERR_BAD_ARGUMENT, "comm_init_local: 1..%d handles", COMM_MAXR);
  std::vector<RankInfo> R(size);
  for (int q = 0; q < size; ++q) {
    if (!hs[q]) return null_handle();
    if (cudaSetDevice(hs[q]->device) != cudaSuccess) return fail(hs[q], SIAFD_B200_ERR_CUDA, "cudaSetDevice failed");
    int st = fill_my_info(hs[q], R[q], false);
    if (st) return st;
  }
  std::vector<std::vector<void *>> fld(size, std::vector<void *>(SIAFD_B200_F_COUNT, nullptr));
  std::vector<void *> pad(size, nullptr);
  for (int q = 0; q < size; ++q) {
    pad[q] = (void *)(uintptr_t)R[q].pad_ptr;
    for (int f = 0; f < SIAFD_B200_F_COUNT; ++f) fld[q][f] = R[q].have[f] ? (void *)(uintptr_t)R[q].field_ptr[f] : nullptr;
  }
  for (int q = 0; q < size; ++q) {
    for (int p = 0; p < size; ++p) {
      if (R[p].device != R[q].device) {
        cudaSetDevice(R[q].device);
        const cudaError_t e = cudaDeviceEnablePeerAccess(R[p].device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return fail(hs[q], SIAFD_B200_ERR_CUDA, "no peer access between devices %d and %d", R[q].device, R[p].device);
        cudaGetLastError();
      }
    }
    if (cudaSetDevice(hs[q]->device) != cudaSuccess) return fail(hs[q], SIAFD_B200_ERR_CUDA, "cudaSetDevice failed");
    int st = finish_setup(hs[q], q, R, fld, pad);
    if (st) return st;
  }
  return SIAFD_B200_OK;
}

int siafd_b200_comm_rank(const siafd_b200_handle *h) { return (h && h->comm.active) ? h->comm.rank : -1; }
int siafd_b200_comm_size(const siafd_b200_handle *h) { return (h && h->comm.active) ? h->comm.size : (h ? 1 : -1); }

int siafd_b200_comm_exchange(siafd_b200_handle *h, int n, const int *fields, const int *widths) {
  if (!h) return null_handle();
  CU(h, cudaSetDevice(h->device));
  if (!h->comm.active) return fail(h, SIAFD_B200_ERR_BAD_ARGUMENT, "comm_exchange: no communicator (siafd_b200_comm_init)");
  if (n < 1 || n > 6) return fail(h, SIAFD_B200_ERR_BAD_ARGUMENT, "comm_exchange: 1..6 fields");
  HaloBatch B;
  B.n = 0;
  for (int q = 0; q < n; ++q) {
    int st = strips_to_peers(h, fields[q], widths[q], B);
    if (st) return st;
  }
  // Two rounds.  (1) "ready to receive": a neighbour may only store into this rank's ghost cells once everything this
  // rank enqueued before this call has run (e.g. an upload that rewrote the whole array, ghost cells included): phase 7.
  // (2) the strips, then "delivered": phases 4..6 in turn; every rank makes the same sequence of calls, so the rows of
  // counters match up.
  if (h->comm.size > 1) h->launches += launch_comm_sync(h->comm.d_peers, 7, h->stream);
  const int phase = 4 + (int)(h->comm.xchg_calls++ % 3);
  h->launches += launch_halo_xchg(B, h->comm.d_peers, phase, h->comm.size > 1 ? 1 : 0, h->stream);
  h->cfl3_fresh = false;
  CU(h, cudaGetLastError());
  return SIAFD_B200_OK;
}

int siafd_b200_comm_allreduce(siafd_b200_handle *h, int op, int n, double *values) {
  if (!h) return null_handle();
  CU(h, cudaSetDevice(h->device));
  if (n < 1 || n > 8 || op < 0 || op > 2 || !values) return fail(h, SIAFD_B200_ERR_BAD_ARGUMENT, "comm_allreduce: op 0..2, 1..8 values");
  if (!h->comm.active || h->comm.size == 1) return SIAFD_B200_OK; // one rank: the values are the result
  std::memcpy(h->comm.h_red, values, sizeof(double) * n);
  CU(h, cudaMemcpyAsync(h->comm.d_red, h->comm.h_red, sizeof(double) * n, cudaMemcpyHostToDevice, h->stream));
  h->launches += launch_comm_allreduce(h->comm.d_peers, h->comm.d_red, nullptr, n, op, h->stream);
  CU(h, cudaMemcpyAsync(h->comm.h_red, h->comm.d_red, sizeof(double) * n, cudaMemcpyDeviceToHost, h->stream));
  CU(h, cudaStreamSynchronize(h->stream));
  std::memcpy(values, h->comm.h_red, sizeof(double) * n);
  return SIAFD_B200_OK;
}

// SIAFD::update (SIAFD.cc:122-155) of one rank of a decomposed run, device-resident, with every ghost update and the
// reductions inside.  Asynchronous: siafd_b200_finish waits and returns the (collective) status.
int siafd_b200_update_decomposed(siafd_b200_handle *h, int full_update, double current_time, int exchange_inputs) {
  if (!h) return null_handle();
  CU(h, cudaSetDevice(h->device));
  siafd_b200_handle::Comm &C = h->comm;
  if (!C.active) return fail(h, SIAFD_B200_ERR_BAD_ARGUMENT, "update_decomposed: no communicator (siafd_b200_comm_init)");
  h->cfl3_fresh = false;
  // allocations and checks happen outside any capture
  {
    const int need[] = {SIAFD_B200_F_SURFACE, SIAFD_B200_F_THICKNESS, SIAFD_B200_F_MASK, SIAFD_B200_F_BED,
                        SIAFD_B200_F_H_X,     SIAFD_B200_F_H_Y,       SIAFD_B200_F_W_I,  SIAFD_B200_F_W_J,
                        SIAFD_B200_F_ENTHALPY, SIAFD_B200_F_TOPGSMOOTH, SIAFD_B200_F_MAXTL, SIAFD_B200_F_C2,
                        SIAFD_B200_F_C3,      SIAFD_B200_F_C4,        SIAFD_B200_F_THK_SMOOTH, SIAFD_B200_F_THETA,
                        SIAFD_B200_F_D,       SIAFD_B200_F_FLUX,      SIAFD_B200_F_U,    SIAFD_B200_F_V,
                        SIAFD_B200_F_SLIDING};
    for (int f : need) {
      int st = ensure(h, f);
      if (st) return st;
    }
  }
  const bool multi = C.size > 1;
  const bool haseloff = h->P.grad == GRAD_HASELOFF;
  const int key = (full_update ? 1 : 0) | (exchange_inputs ? 2 : 0);
  const bool use_graph = h->tuning.graph_step && !h->timing;
  if (use_graph && C.graph_valid[key] && C.graph_time[key] == current_time && C.graph_stream[key] == h->stream) {
    CU(h, cudaGraphLaunch(C.graph_exec[key], h->stream));
    h->launches += C.graph_launches[key];
    h->result_pending = false;
    C.result_from_comm = true;
    return SIAFD_B200_OK;
  }

  auto enqueue = [&]() -> int {
    const int64_t l0 = h->launches;
    int st;
    bool forked = false;
    // (timing mode, ungraphed: an event after every section of the step, see siafd_b200_step_breakdown_ms)
    const bool sec = h->timing && !h->ev_sec.empty() && h->sec_count < (int)h->ev_sec.size() / 6;
    auto mark = [&](int q) {
      if (sec) cudaEventRecord(h->ev_sec[6 * h->sec_count + q], h->stream);
    };
    mark(0);
    if (exchange_inputs) {
      // the inputs' ghosts (device-resident callers; under PISM the host arrays already carry them): the 2D fields on
      // the main stream, the enthalpy beside them on a second one, joined before the fused kernel
      HaloBatch B;
      B.n = 0;
      const int f2[] = {SIAFD_B200_F_SURFACE, SIAFD_B200_F_THICKNESS, SIAFD_B200_F_MASK, SIAFD_B200_F_BED};
      for (int f : f2) {
        if ((st = strips_to_peers(h, f, meta(h->cfg, f).width, B))) return st;
      }
      h->launches += launch_halo_xchg(B, C.d_peers, 0, multi ? 1 : 0, h->stream);
      HaloBatch B3;
      B3.n = 0;
      if ((st = strips_to_peers(h, SIAFD_B200_F_ENTHALPY, h->cfg.w_3d_in, B3))) return st;
      if (h->P.use_age && (st = strips_to_peers(h, SIAFD_B200_F_AGE, h->cfg.w_3d_in, B3))) return st;
      CU(h, cudaEventRecord(C.ev_fork, h->stream));
      CU(h, cudaStreamWaitEvent(C.s_aux, C.ev_fork, 0));
      h->launches += launch_halo_xchg(B3, C.d_peers, 1, multi ? 1 : 0, C.s_aux);
      CU(h, cudaEventRecord(C.ev_join, C.s_aux));
      forked = true;
    }
    mark(1);
    const Fields F = fields_of(h);
    PeerPush PPg;
    comm_make_push(h, SIAFD_B200_F_H_X, SIAFD_B200_F_H_Y, h->cfg.w_stag, 1, PPg);
    // gradient (SIAFD.cc:137, ghost update :498-499 fused) and, for haseloff, thk_smooth / theta (:580-582) in one pass;
    // the pass also weighs the fused kernel's row segments, and its last CTA sorts them, heaviest first
    h->P.current_time = current_time;
    {
      const int nseg_all = slab_segments(h->P, h->tuning);
      const bool order = h->tuning.order_segments && nseg_all <= 128 && nseg_all > 1;
      h->P.seg_rows = slab_rows_per_segment(h->tuning), h->P.seg_n = order ? nseg_all : 0;
      if (order) CU(h, cudaMemsetAsync(h->d_segw, 0, 128 * sizeof(int), h->stream));
    }
    h->launches += launch_gradient(h->P, F, h->stream, haseloff ? &PPg : nullptr, haseloff);
    CU(h, cudaGetLastError());
    if ((st = flux_velocity_prepare(h, full_update, current_time, haseloff))) return st;
    mark(2);
    // (haseloff with geometry ghosts two cells wide: the pass evaluates the ring of ghost points itself, bit for bit
    // what the neighbours hold -- no exchange, no synchronisation of the ranks in the middle of the step)
    if (haseloff && multi && !gradient_ring_is_local(h->P)) h->launches += launch_comm_sync(C.d_peers, 2, h->stream);
    if (forked) CU(h, cudaStreamWaitEvent(h->stream, C.ev_join, 0));
    mark(3);
    PeerPush PPu;
    comm_make_push(h, SIAFD_B200_F_U, SIAFD_B200_F_V, h->cfg.w_uv, 1, PPu);
    {
      const Tuning T = h->tuning;
      const bool timed = h->timing && h->ev_count < (int)h->ev_start.size();
      if (timed) CU(h, cudaEventRecord(h->ev_start[h->ev_count], h->stream));
      const int n = launch_slab(h->P, F, full_update != 0, T, (long)siafd_b200_field_size(h, SIAFD_B200_F_ENTHALPY),
                                (long)siafd_b200_field_size(h, SIAFD_B200_F_THK_SMOOTH), h->inv_dz, 0, -1, h->stream,
                                (full_update && !getenv("SIAFD_B200_NOPUSH")) ? &PPu : nullptr,
                                h->P.seg_n > 0 ? h->d_segw + 128 : nullptr);
      h->P.seg_n = 0; // (only this entry point sorts: the split calls and the banded host pipeline keep the plain order)
      if (timed) {
        CU(h, cudaEventRecord(h->ev_stop[h->ev_count], h->stream));
        h->ev_count += 1;
      }
      if (n < 0) return fail(h, SIAFD_B200_ERR_CUDA, "could not configure the fused kernel");
      h->launches += n;
    }
    mark(4);
    // u, v arrival + {D_max, error bits, counter} over all ranks (SIAFD.cc:748-750), straight into pinned host memory
    h->launches += launch_comm_final(C.d_peers, (full_update && multi) ? 3 : -1, h->d_dmax, h->d_err, h->d_hdc, C.d_res,
                                     C.h_res, h->stream);
    CU(h, cudaGetLastError());
    mark(5);
    if (sec) h->sec_count += 1;
    C.graph_launches[key] = (int)(h->launches - l0);
    return SIAFD_B200_OK;
  };

  if (!use_graph) {
    int st = enqueue();
    if (st) return st;
  } else {
    if (C.graph_exec[key]) {
      cudaGraphExecDestroy(C.graph_exec[key]);
      C.graph_exec[key] = nullptr, C.graph_valid[key] = false;
    }
    CU(h, cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
    const int64_t l0 = h->launches;
    int st = enqueue();
    cudaGraph_t g = nullptr;
    const cudaError_t e = cudaStreamEndCapture(h->stream, &g);
    h->launches = l0;
    if (st) {
      if (g) cudaGraphDestroy(g);
      return st;
    }
    if (e != cudaSuccess) return fail(h, SIAFD_B200_ERR_CUDA, "graph capture of the step failed: %s", cudaGetErrorString(e));
    const cudaError_t e2 = cudaGraphInstantiate(&C.graph_exec[key], g, 0);
    cudaGraphDestroy(g);
    if (e2 != cudaSuccess) return fail(h, SIAFD_B200_ERR_CUDA, "graph instantiation failed: %s", cudaGetErrorString(e2));
    C.graph_valid[key] = true, C.graph_time[key] = current_time, C.graph_stream[key] = h->stream;
    CU(h, cudaGraphLaunch(C.graph_exec[key], h->stream));
    h->launches += C.graph_launches[key];
  }
  h->result_pending = false;
  C.result_from_comm = true;
  return SIAFD_B200_OK;
}

} // extern "C"
