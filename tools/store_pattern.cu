// store_pattern.cu -- how fast can the write-only regime of the fused kernel go?  Writes zeros into two arrays shaped like
// u and v of the 4096^2 x 101 dome (rows of (4096 + 2) columns x 101 doubles) in several patterns and prints TB/s:
//   memset     cudaMemsetAsync of both arrays (the ceiling)
//   strips     the fused kernel's pattern: CTA = strip of 15 columns x 64 rows, per row one 12 120-byte bulk store per array
//   strips2/4  the same with strips of 30 / 60 columns (24 / 48 KB per store)
//   rows       CTA = one row, marching along x in 48 KB bulk stores (both arrays)
//   plain      the strips pattern with 16-byte st.global instead of bulk stores
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o store_pattern store_pattern.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bulk_s2g(void *g, const void *s, unsigned bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n" ::"l"(g), "r"(smem_u32(s)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory"); }
__device__ __forceinline__ void fence_async() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }

// CTA (strip, segment): columns [1 + strip * W, ...) of W columns, rows [seg * RS, ...): per row one bulk store per array
__global__ void k_strips(double *u, double *v, int xm, int ym, int Mz, int W, int RS, int plain) {
  extern __shared__ __align__(16) double z[];
  const int ncol = min(W, xm - (int)blockIdx.x * W);
  const long n = (long)ncol * Mz;
  for (int e = threadIdx.x; e < W * Mz + 2; e += blockDim.x) z[e] = 0.0;
  fence_async();
  __syncthreads();
  const long rowlen = (long)(xm + 2) * Mz;
  const int r0 = blockIdx.y * RS, r1 = min(r0 + RS, ym);
  for (int r = r0; r < r1; ++r) {
    const long g0 = (long)(r + 1) * rowlen + (long)(1 + blockIdx.x * W) * Mz;
    if (plain) {
      const long a0 = (g0 + 1) & ~1L, a1 = (g0 + n) & ~1L;
      for (long e = a0 + 2 * threadIdx.x; e < a1; e += 2 * blockDim.x) {
        *reinterpret_cast<double2 *>(u + e) = make_double2(0.0, 0.0);
        *reinterpret_cast<double2 *>(v + e) = make_double2(0.0, 0.0);
      }
      if (threadIdx.x == 0) {
        if (g0 & 1) u[g0] = 0.0, v[g0] = 0.0;
        if ((g0 + n) & 1) u[g0 + n - 1] = 0.0, v[g0 + n - 1] = 0.0;
      }
    } else if (threadIdx.x == 0) {
      const long a0 = (g0 + 1) & ~1L, a1 = (g0 + n) & ~1L;
      if (g0 & 1) u[g0] = 0.0, v[g0] = 0.0;
      if ((g0 + n) & 1) u[g0 + n - 1] = 0.0, v[g0 + n - 1] = 0.0;
      bulk_s2g(u + a0, z, (unsigned)((a1 - a0) * 8));
      bulk_s2g(v + a0, z, (unsigned)((a1 - a0) * 8));
      bulk_commit();
    }
    __syncthreads(); // (the fused kernel has a barrier per row too)
  }
  if (threadIdx.x == 0) bulk_wait0();
}
// CTA = rows blockIdx.x, blockIdx.x + gridDim.x, ...: the owned part of a row in pieces of `piece` doubles
__global__ void k_rows(double *u, double *v, int xm, int ym, int Mz, int piece) {
  extern __shared__ __align__(16) double z[];
  for (int e = threadIdx.x; e < piece; e += blockDim.x) z[e] = 0.0;
  fence_async();
  __syncthreads();
  const long rowlen = (long)(xm + 2) * Mz, n = (long)xm * Mz;
  if (threadIdx.x == 0) {
    for (int r = blockIdx.x; r < ym; r += gridDim.x) {
      const long g0 = (long)(r + 1) * rowlen + Mz;
      const long a0 = (g0 + 1) & ~1L, a1 = (g0 + n) & ~1L;
      if (g0 & 1) u[g0] = 0.0, v[g0] = 0.0;
      if ((g0 + n) & 1) u[g0 + n - 1] = 0.0, v[g0 + n - 1] = 0.0;
      for (long a = a0; a < a1; a += piece) {
        const unsigned b = (unsigned)(min((long)piece, a1 - a) * 8);
        bulk_s2g(u + a, z, b);
        bulk_s2g(v + a, z, b);
        bulk_commit();
      }
    }
    bulk_wait0();
  }
}
int main(int argc, char **argv) {
  const int M = argc > 1 ? atoi(argv[1]) : 4096, Mz = 101;
  const long n = (long)(M + 2) * (M + 2) * Mz;
  double *u, *v;
  CK(cudaMalloc(&u, n * 8));
  CK(cudaMalloc(&v, n * 8));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  const double gb = 2.0 * M * M * Mz * 8 / 1e9;
  auto timeit = [&](const char *name, auto fn) {
    for (int w = 0; w < 2; ++w) fn();
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    const int reps = 5;
    for (int w = 0; w < reps; ++w) fn();
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    CK(cudaGetLastError());
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("{\"pattern\": \"%s\", \"ms\": %.3f, \"TBps\": %.3f}\n", name, ms / reps, gb / (ms / reps) );
  };
  timeit("memset", [&] { cudaMemsetAsync(u, 0, n * 8); cudaMemsetAsync(v, 0, n * 8); });
  for (int W : {15, 30, 60}) {
    const size_t sm = ((size_t)W * Mz + 2) * 8;
    CK(cudaFuncSetAttribute(k_strips, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    char name[64];
    for (int RS : {64, 16}) {
      snprintf(name, 64, "strips W=%d RS=%d bulk", W, RS);
      dim3 g((M + W - 1) / W, (M + RS - 1) / RS);
      timeit(name, [&] { k_strips<<<g, 128, sm>>>(u, v, M, M, Mz, W, RS, 0); });
      if (W == 15) {
        snprintf(name, 64, "strips W=%d RS=%d plain", W, RS);
        timeit(name, [&] { k_strips<<<g, 128, sm>>>(u, v, M, M, Mz, W, RS, 1); });
      }
    }
  }
  // the fused kernel's footprint: 75 KB of shared memory per CTA (3 CTAs per SM) with the 12 KB pattern
  {
    const size_t sm = 75 * 1024;
    CK(cudaFuncSetAttribute(k_strips, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    dim3 g((M + 14) / 15, (M + 63) / 64);
    timeit("strips W=15 RS=64 bulk, 75 KB smem (3 CTAs/SM)", [&] { k_strips<<<g, 128, sm>>>(u, v, M, M, Mz, 15, 64, 0); });
  }
  for (int piece : {6144, 12288, 24576}) {
    const size_t sm = (size_t)piece * 8;
    CK(cudaFuncSetAttribute(k_rows, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    char name[64];
    for (int nb : {148, 296, 592}) {
      snprintf(name, 64, "rows piece=%dKB ctas=%d", piece * 8 / 1024, nb);
      timeit(name, [&] { k_rows<<<nb, 64, sm>>>(u, v, M, M, Mz, piece); });
    }
  }
  return 0;
}
