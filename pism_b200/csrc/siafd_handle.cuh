// siafd_handle.cuh -- the handle behind the C ABI (include/siafd_b200.h) and the host helpers shared by
// siafd_capi.cu and siafd_comm.cu.  Private to the library.
#pragma once
#include "../../include/siafd_b200.h"
#include "siafd_kernels.cuh"

#include <cstdint>
#include <string>
#include <vector>

struct siafd_b200_handle {
  siafd_b200_config cfg;
  std::vector<double> z;
  siafd::DP P;
  int device = 0;
  cudaStream_t own_stream = nullptr, stream = nullptr;
  void *buf[SIAFD_B200_F_COUNT];
  bool owned[SIAFD_B200_F_COUNT];
  double *d_z = nullptr;
  unsigned *d_err = nullptr;
  unsigned long long *d_dmax = nullptr;
  unsigned long long *d_cfl = nullptr, *h_cfl = nullptr; // 8 maxima of siafd_b200_cfl and their pinned mirror
  bool cfl3_fresh = false; // slots 0..3 hold the maxima the last vertical-velocity launch took on the current fields
  int fill_threads = 4;    // host threads that fill the ice-free parts of u, v in the sparse host path (more of them
                           // do not help: 4096^2, 4 / 8 threads: 362 / 368 ms; while they run the copy engine's
                           // device-to-host rate drops from 48 to 27 GB/s, profiles/e2e_trace_r02.txt)
  int repl_threads = 8;    // host threads that replicate the top value of u, v above the cut level (level_cut); a
                           // band of 64 rows is four tasks of 16 rows per chunk: 4 or 8 threads keep each thread in
                           // its own rows (6 threads measured 424 ms against 357 ms for 4 or 8)
  int64_t bytes_h2d = 0, bytes_d2h = 0; // bytes the host-path calls moved over PCIe since create
  int vvel_rows = 64;      // rows one CTA of the marching vertical-velocity kernels takes
  int vvel_kind = 0;       // 0: k_vvel_slab (shared memory, z sweep in registers); 1: k_vvel_march (lanes across z)
  int vvel_wz = 16;        // z ranges per column of k_vvel_slab
  int *d_hdc = nullptr;
  int *d_segw = nullptr;          // [0, 128) weights of the fused kernel's row segments, [128, 256) their order
  unsigned *d_segdone = nullptr;  // CTA counter of the 2D pass that sorts them
  // pinned host mirror of {err, hdc, dmax}
  struct Result {
    unsigned long long dmax;
    unsigned err;
    int hdc;
  } *h_res = nullptr;
  bool result_pending = false;
  bool smoother_set = false;
  int bedNx = -1, bedNy = -1;
  double *d_global_bed = nullptr;
  cudaStream_t s_up = nullptr, s_dn = nullptr; // upload / download legs of the pipelined host update
  void *d_pieces = nullptr;                    // the call's download pieces on the device (zero_copy)
  size_t d_pieces_bytes = 0;
  int zero_copy = 0; // 1: u, v go into mapped pinned host arrays by stores of a kernel instead of strided copies
  std::vector<cudaEvent_t> ev_pipe;
  // peer halo exchange: per field and neighbour direction the mapped base of the neighbour's array (nullptr =
  // this rank) and its patch size; the arrival-counter pad [4 phases][8 dirs] and the neighbours' pads
  struct Peer {
    double *base = nullptr;
    int xm = 0, ym = 0;
    bool attached = false;
  } peers[SIAFD_B200_F_COUNT][8];
  // communicator of a decomposed run (siafd_comm.cu)
  struct Comm {
    bool active = false;
    int rank = 0, size = 1;
    int nb[8] = {};
    void *pad = nullptr;               // this rank's siafd::CommPad (device)
    siafd::CommPeers *d_peers = nullptr;
    unsigned long long *d_res = nullptr, *h_res = nullptr; // {D_max bits, error bits, counter} over all ranks
    double *d_red = nullptr, *h_red = nullptr;             // staging of comm_allreduce
    cudaStream_t s_aux = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    cudaGraphExec_t graph_exec[4] = {}; // by (full_update, exchange_inputs)
    bool graph_valid[4] = {};
    double graph_time[4] = {};
    cudaStream_t graph_stream[4] = {};
    int graph_launches[4] = {};
    std::vector<void *> mapped;     // cudaIpcOpenMemHandle results
    std::vector<std::string> files; // rendezvous files this rank wrote
    unsigned long long xchg_calls = 0;
    bool result_from_comm = false; // the last update left its (global) result in h_res
  } comm;
  unsigned long long *d_pad = nullptr, *peer_pad[8] = {};
  bool pad_attached[8] = {};
  unsigned long long halo_step[4] = {0, 0, 0, 0};
  std::vector<void *> ipc_mapped;
  siafd::Tuning tuning;
  double inv_dz = 0.0; // (Mz - 1) / Lz when the levels are equally spaced, else 0
  int64_t launches = 0;
  // CUDA-event pairs around the fused kernel (bench.py's roofline timing), a ring of 256
  std::vector<cudaEvent_t> ev_start, ev_stop;
  int ev_count = 0;
  bool timing = false;
  // ... and, in the same mode, six events per step of siafd_b200_update_decomposed: start | inputs' ghosts (2D) |
  // gradient pass | synchronisation before the fused kernel (+ the 3D ghosts' side stream) | fused kernel | final
  // signal / wait / reduction (siafd_b200_step_breakdown_ms)
  std::vector<cudaEvent_t> ev_sec;
  int sec_count = 0;
  std::string err;
};


namespace siafd_host {

struct FieldMeta {
  int width;
  int dof;
};

int fail(siafd_b200_handle *h, int code, const char *fmt, ...);
int null_handle();
FieldMeta meta(const siafd_b200_config &c, int f);
int64_t field_cells(const siafd_b200_config &c, int w);
int ensure(siafd_b200_handle *h, int f);
siafd::Fields fields_of(siafd_b200_handle *h);
int status_from_bits(unsigned bits);
int fetch_result(siafd_b200_handle *h);
// checks, scratch fields and the 2D preparation of SIAFD::compute_diffusivity (SIAFD.cc:555-582); the fused kernel
int flux_velocity_prepare(siafd_b200_handle *h, int full_update, double current_time, bool prep2d_done = false);
int flux_velocity_launch(siafd_b200_handle *h, int full_update, int seg0, int nseg, const siafd::PeerPush *push = nullptr);
// siafd_comm.cu: fused-push table of two same-shaped fields (h_x / h_y, u / v) with ghost width W, strips of width w
void comm_make_push(const siafd_b200_handle *h, int fa, int fb, int W, int w, siafd::PeerPush &PP);
void comm_release(siafd_b200_handle *h); // siafd_comm.cu: unmaps peers, frees the pad

} // namespace siafd_host

#define CU(h, call)                                                                                                    \
  do {                                                                                                                 \
    cudaError_t e_ = (call);                                                                                           \
    if (e_ != cudaSuccess) {                                                                                           \
      return siafd_host::fail((h), SIAFD_B200_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_),        \
                              __FILE__, __LINE__);                                                                     \
    }                                                                                                                  \
  } while (0)
