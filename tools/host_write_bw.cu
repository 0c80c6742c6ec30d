// What bounds siafd_b200_update with host arrays on one rank: the rate at which the host's memory takes writes.
// Measures, on pinned host memory, (a) a device-to-host copy alone, (b) non-temporal fills by T host threads alone,
// (c) both at once (the copy into one half of the buffer, the threads into the other) -- the mix of the sparse path.
// nvcc -O2 -std=c++17 -gencode arch=compute_100a,code=sm_100a tools/host_write_bw.cu -o tools/host_write_bw
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <emmintrin.h>
#include <thread>
#include <vector>
#include <cuda_runtime.h>

static void fill(double *p, size_t n) {
  const __m128d v = _mm_set1_pd(0.0);
  for (size_t k = 0; k + 8 <= n; k += 8) {
    _mm_stream_pd(p + k, v), _mm_stream_pd(p + k + 2, v), _mm_stream_pd(p + k + 4, v), _mm_stream_pd(p + k + 6, v);
  }
  _mm_sfence();
}
static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

int main(int argc, char **argv) {
  const size_t GB = argc > 1 ? atoi(argv[1]) : 8, half = GB * (size_t(1) << 30) / 2, n = half / 8;
  double *h, *d;
  if (cudaHostAlloc(&h, 2 * half, cudaHostAllocDefault) != cudaSuccess || cudaMalloc(&d, half) != cudaSuccess) return 1;
  cudaMemset(d, 0, half);
  fill(h, 2 * n);
  cudaStream_t s;
  cudaStreamCreate(&s);
  auto copy = [&](int reps) {
    for (int r = 0; r < reps; ++r) cudaMemcpyAsync(h, d, half, cudaMemcpyDeviceToHost, s);
  };
  auto threads = [&](int T, int reps) {
    std::vector<std::thread> w;
    for (int t = 0; t < T; ++t)
      w.emplace_back([=] {
        for (int r = 0; r < reps; ++r) fill(h + n + (n / T) * t, n / T);
      });
    for (auto &x : w) x.join();
  };
  copy(1), cudaStreamSynchronize(s);
  double t0 = now();
  copy(3), cudaStreamSynchronize(s);
  double dt = now() - t0;
  printf("copy alone: %.1f GB/s\n", 3.0 * half / dt / 1e9);
  for (int T : {1, 2, 4, 8, 12, 16}) {
    t0 = now();
    threads(T, 3);
    dt = now() - t0;
    const double alone = 3.0 * half / dt / 1e9;
    t0 = now();
    copy(3);
    threads(T, 3);
    const double tf = now() - t0;
    cudaStreamSynchronize(s);
    const double tc = now() - t0;
    printf("%2d threads: fill alone %.1f GB/s; together: fill %.1f GB/s (%.3f s), copy %.1f GB/s (%.3f s), sum over the longer %.1f GB/s\n", T, alone,
           3.0 * half / tf / 1e9, tf, 3.0 * half / tc / 1e9, tc, 6.0 * half / (tf > tc ? tf : tc) / 1e9);
  }
  return 0;
}
