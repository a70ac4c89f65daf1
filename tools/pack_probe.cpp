// Host-side packing probe: gather the node rows of a column-major G x n matrix into a dense column-major block with T
// threads (what the engine does for scattered node rows / pageable matrices).  g++ -O3 -pthread tools/pack_probe.cpp
#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
int main(int argc, char** argv) {
  const int G = 20000, n = 1000, nn = 8405;
  std::vector<double> src((size_t)G * n);
  for (size_t i = 0; i < src.size(); ++i) src[i] = (double)(i % 977);
  std::vector<int> rows;
  std::vector<char> mark(G, 0);
  unsigned s = 12345;
  int cnt = 0;
  while (cnt < nn) {
    s = s * 1664525u + 1013904223u;
    int r = (s >> 8) % G;
    if (!mark[r]) { mark[r] = 1; ++cnt; }
  }
  for (int r = 0; r < G; ++r) if (mark[r]) rows.push_back(r);
  double* dst = nullptr;
  posix_memalign((void**)&dst, 4096, (size_t)nn * n * 8);
  memset(dst, 0, (size_t)nn * n * 8);
  for (int T : {1, 2, 4, 8, 12, 16, 24, 32}) {
    double best = 1e9;
    for (int rep = 0; rep < 5; ++rep) {
      std::atomic<int> next{0};
      auto t0 = std::chrono::steady_clock::now();
      std::vector<std::thread> th;
      for (int t = 0; t < T; ++t)
        th.emplace_back([&]() {
          for (;;) {
            int c0 = next.fetch_add(8);
            if (c0 >= n) break;
            for (int c = c0; c < c0 + 8 && c < n; ++c) {
              const double* sc = src.data() + (size_t)c * G;
              double* dc = dst + (size_t)c * nn;
              for (int i = 0; i < nn; ++i) dc[i] = sc[rows[i]];
            }
          }
        });
      for (auto& x : th) x.join();
      double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
      if (ms < best) best = ms;
    }
    printf("threads %2d: %.3f ms  (%.1f GB/s read of the full matrix, %.1f GB/s packed)\n", T, best, G * (double)n * 8 / best / 1e6,
           nn * (double)n * 8 / best / 1e6);
  }
  printf("hardware_concurrency %u\n", std::thread::hardware_concurrency());
  return 0;
}
