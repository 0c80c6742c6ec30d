// memcpy2d_bw.cu -- device -> pinned host bandwidth of strided copies shaped like a level cut of u / v: columns of Mz = 101
// doubles (pitch 808 B) of which only the lowest L levels (width 8 L bytes) are copied.  Prints GB/s of the bytes moved.
// nvcc -O3 -o memcpy2d_bw memcpy2d_bw.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)
int main() {
  const int Mz = 101;
  const size_t cols = 4u << 20; // 4 M columns = 3.4 GB
  const size_t bytes = cols * Mz * 8;
  char *d, *h;
  CK(cudaMalloc(&d, bytes));
  CK(cudaMallocHost(&h, bytes));
  CK(cudaMemset(d, 1, bytes));
  cudaStream_t s;
  CK(cudaStreamCreate(&s));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  for (int dir = 0; dir < 2; ++dir) {
    for (int L : {101, 90, 75, 60, 45, 30, 16}) {
      const size_t width = (size_t)L * 8, pitch = (size_t)Mz * 8;
      for (int rep = 0; rep < 2; ++rep) {
        CK(cudaEventRecord(e0, s));
        // bands like the host pipeline's: 64 copies of cols / 64 columns each
        for (int b = 0; b < 64; ++b) {
          const size_t off = (size_t)b * (cols / 64) * pitch;
          if (dir == 0) {
            CK(cudaMemcpy2DAsync(h + off, pitch, d + off, pitch, width, cols / 64, cudaMemcpyDeviceToHost, s));
          } else {
            CK(cudaMemcpy2DAsync(d + off, pitch, h + off, pitch, width, cols / 64, cudaMemcpyHostToDevice, s));
          }
        }
        CK(cudaEventRecord(e1, s));
        CK(cudaStreamSynchronize(s));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (rep == 1)
          printf("{\"dir\": \"%s\", \"levels\": %d, \"width_B\": %zu, \"ms\": %.2f, \"GBps_moved\": %.1f, \"ms_full_equiv\": %.2f}\n",
                 dir ? "h2d" : "d2h", L, width, ms, width * cols / 1e6 / ms, ms);
      }
    }
  }
  return 0;
}
