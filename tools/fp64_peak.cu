// fp64_peak.cu -- FP64 pipe microbenchmark for B200 (sm_100a): throughput and dependent-issue latency of DFMA,
// and how many warps x independent chains per SM sub-partition it takes to fill the pipe.  DESIGN.md's secondary
// roofline of k_sia_slab (the Arrhenius loop) rests on these numbers.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peak tools/fp64_peak.cu && ./fp64_peak > fp64_peak.json
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x)                                                                                                          \
  do {                                                                                                                 \
    cudaError_t e_ = (x);                                                                                              \
    if (e_ != cudaSuccess) {                                                                                           \
      fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_));                                                         \
      exit(1);                                                                                                         \
    }                                                                                                                  \
  } while (0)

// CH independent DFMA chains per thread, ITER trips; per-thread cycle count through clock64 of warp 0
template <int CH>
__global__ void k_dfma(double *out, long long *cyc, int iters, double a, double b) {
  double x[CH];
#pragma unroll
  for (int c = 0; c < CH; ++c) x[c] = 1.0 + 1e-3 * (threadIdx.x + c);
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
#pragma unroll
      for (int c = 0; c < CH; ++c) x[c] = fma(x[c], a, b);
    }
  }
  const long long t1 = clock64();
  double s = 0.0;
#pragma unroll
  for (int c = 0; c < CH; ++c) s += x[c];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// MUFU.RCP64H + the two Newton steps of rcp_fast (siafd_math.cuh), CH independent chains
template <int CH>
__global__ void k_rcp(double *out, long long *cyc, int iters) {
  double x[CH];
#pragma unroll
  for (int c = 0; c < CH; ++c) x[c] = 250.0 + threadIdx.x + c;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      double r;
      asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x[c]));
      double e = fma(-x[c], r, 1.0);
      r = fma(r, e, r);
      e = fma(-x[c], r, 1.0);
      r = fma(r, e, r);
      x[c] = 250.0 + r; // next trip depends on this one
    }
  }
  const long long t1 = clock64();
  double s = 0.0;
#pragma unroll
  for (int c = 0; c < CH; ++c) s += x[c];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int CH> static void run(int warps_per_smsp, int nsm, double clock_ghz, bool first) {
  const int iters = 4096;
  const int threads = warps_per_smsp * 4 * 32; // one CTA per SM, 4 sub-partitions
  double *out;
  long long *cyc, hc[1024];
  CK(cudaMalloc(&out, sizeof(double) * threads * nsm));
  CK(cudaMalloc(&cyc, sizeof(long long) * nsm));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  k_dfma<CH><<<nsm, threads>>>(out, cyc, 64, 1.0000001, 1e-9);
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  k_dfma<CH><<<nsm, threads>>>(out, cyc, iters, 1.0000001, 1e-9);
  CK(cudaEventRecord(e1));
  CK(cudaDeviceSynchronize());
  float ms;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  CK(cudaMemcpy(hc, cyc, sizeof(long long) * nsm, cudaMemcpyDeviceToHost));
  double cmean = 0;
  for (int q = 0; q < nsm; ++q) cmean += (double)hc[q] / nsm;
  const double dfma_per_warp = (double)iters * 8 * CH;
  const double flops = 2.0 * dfma_per_warp * 32 * warps_per_smsp * 4 * nsm / (ms * 1e-3);
  // cycles the sub-partition spends per DFMA warp-instruction = cycles / (instructions issued by its warps)
  const double cyc_per_inst = cmean / (dfma_per_warp * warps_per_smsp);
  printf("%s    {\"warps_per_smsp\": %d, \"chains\": %d, \"tflops\": %.3f, \"cycles_per_dfma_per_smsp\": %.3f, "
         "\"cycles_per_dependent_dfma\": %.3f, \"ms\": %.4f}",
         first ? "" : ",\n", warps_per_smsp, CH, flops / 1e12, cyc_per_inst, cmean / ((double)iters * 8), ms);
  CK(cudaFree(out));
  CK(cudaFree(cyc));
}

int main() {
  cudaDeviceProp p;
  CK(cudaGetDeviceProperties(&p, 0));
  int clk_khz = 0;
  CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0));
  const int nsm = p.multiProcessorCount;
  printf("{\n  \"device\": \"%s\", \"sms\": %d, \"clock_mhz_nominal\": %d,\n  \"dfma\": [\n", p.name, nsm, clk_khz / 1000);
  bool first = true;
  const int ws[] = {1, 2, 3, 4, 6, 8};
  for (int w : ws) {
    run<1>(w, nsm, 0, first), first = false;
    run<2>(w, nsm, 0, false);
    run<3>(w, nsm, 0, false);
    run<4>(w, nsm, 0, false);
    run<8>(w, nsm, 0, false);
  }
  printf("\n  ],\n");
  // rcp_fast chain: latency of MUFU.RCP64H + 4 dependent DFMA + 1 DADD
  {
    double *out;
    long long *cyc, hc;
    CK(cudaMalloc(&out, sizeof(double) * 32));
    CK(cudaMalloc(&cyc, sizeof(long long)));
    k_rcp<1><<<1, 32>>>(out, cyc, 4096);
    CK(cudaDeviceSynchronize());
    k_rcp<1><<<1, 32>>>(out, cyc, 4096);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(&hc, cyc, sizeof(long long), cudaMemcpyDeviceToHost));
    printf("  \"rcp_fast_chain_cycles\": %.2f,\n", (double)hc / 4096);
  }
  // peak: best tflops row is the FP64 pipe peak at the clock the run held
  printf("  \"note\": \"one CTA per SM, warps_per_smsp x 4 warps; cycles_per_dfma_per_smsp = issue cost of one DFMA warp "
         "instruction on a sub-partition when the pipe is full; cycles_per_dependent_dfma with 1 warp, 1 chain = "
         "dependent-issue latency\"\n}\n");
  return 0;
}
