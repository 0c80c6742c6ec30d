// Host microbenchmark of the level cut's replication step (siafd_hostplan.hh::replicate_piece): threads x columns of
// Mz = 101 doubles, levels [n, Mz) of every column set to the value of level n - 1, arrays far larger than the caches.
// g++ -O2 -std=c++17 -pthread tools/replicate_bw.cc -o /tmp/replicate_bw && /tmp/replicate_bw [threads] [n] [GiB]
#include "../pism_b200/csrc/siafd_hostplan.hh"

#include <chrono>
#include <cstdio>
#include <cstdlib>

using namespace siafd_hostplan;

static void old_piece(const Piece &p, double *u, double *v, long row_cells, int Mz) { // streaming stores throughout
  double *arr[2] = {u, v};
  for (int r = p.r0; r < p.r1; ++r)
    for (int q = 0; q < 2; ++q) {
      double *col = arr[q] + ((long)r * row_cells + p.c0) * Mz;
      for (int cc = p.c0; cc < p.c1; ++cc, col += Mz) fill_stream(col + p.n, (size_t)(Mz - p.n), col[p.n - 1]);
    }
}

int main(int argc, char **argv) {
  const int T = argc > 1 ? atoi(argv[1]) : 4, n = argc > 2 ? atoi(argv[2]) : 60;
  const double gib = argc > 3 ? atof(argv[3]) : 2.0;
  const int Mz = 101, cells = 4098;
  const int rows = (int)(gib * (1 << 30) / 2 / ((double)cells * Mz * 8));
  const size_t N = (size_t)rows * cells * Mz;
  double *u = (double *)aligned_alloc(4096, (N * 8 + 4095) / 4096 * 4096), *v = (double *)aligned_alloc(4096, (N * 8 + 4095) / 4096 * 4096);
  for (size_t k = 0; k < N; ++k) u[k] = (double)(k % 977), v[k] = (double)(k % 499);
  for (int variant = 0; variant < 2; ++variant) {
    for (int rep = 0; rep < 2; ++rep) {
      const auto t0 = std::chrono::steady_clock::now();
      std::vector<std::thread> w;
      for (int t = 0; t < T; ++t)
        w.emplace_back([=] {
          for (int r = t * 16; r < rows; r += T * 16) {
            Piece p{r, std::min(rows, r + 16), 500, 3600, n};
            variant ? replicate_piece(p, u, v, cells, Mz) : old_piece(p, u, v, cells, Mz);
          }
        });
      for (auto &x : w) x.join();
      const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
      const double cols = 2.0 * rows * 3100;
      printf("%s threads %d n %d: %.3f s, %.1f ns per column per thread, %.2f GB/s of stores\n", variant ? "new" : "old", T, n, s,
             s * T / cols * 1e9, cols * (Mz - n) * 8 / s / 1e9);
    }
  }
  double chk = 0;
  for (size_t k = 0; k < N; k += 4097) chk += u[k] + v[k];
  printf("checksum %.1f\n", chk);
  return 0;
}
