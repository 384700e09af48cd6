// Memory-bandwidth ceiling of one NUMA node for the staging pass: T threads pinned to the node's CPUs,
// (a) AVX2 read of one array, (b) read 3 arrays + write 1 (the pass's mix), regular and NT stores.
// Build: nvcc -O3 -std=c++17 -Xcompiler -pthread,-mavx2 scripts/host_probe2.cu -o build/host_probe2
#include <immintrin.h>
#include <sched.h>
#include <unistd.h>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <functional>
#include <string>
#include <thread>
#include <vector>
static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
static std::vector<int> node_cpus(int node) {
  std::vector<int> out; std::ifstream f("/sys/devices/system/node/node" + std::to_string(node) + "/cpulist"); std::string s; std::getline(f, s);
  size_t p = 0; while (p < s.size()) { int a = atoi(s.c_str() + p), b = a; size_t q = s.find_first_of(",-", p);
    if (q != std::string::npos && s[q] == '-') { b = atoi(s.c_str() + q + 1); q = s.find(',', q); }
    for (int c = a; c <= b; ++c) out.push_back(c); if (q == std::string::npos) break; p = q + 1; }
  return out;
}
static void par(int T, const std::vector<int> &cpus, const std::function<void(int)> &f) {
  std::vector<std::thread> th;
  for (int t = 0; t < T; ++t) th.emplace_back([&, t] { cpu_set_t s; CPU_ZERO(&s); for (int c : cpus) CPU_SET(c, &s); sched_setaffinity(0, sizeof(s), &s); f(t); });
  for (auto &x : th) x.join();
}
int main(int argc, char **argv) {
  const long N = 10000000; const int node = argc > 1 ? atoi(argv[1]) : 0; const int mem_node = argc > 2 ? atoi(argv[2]) : node;
  auto cpus = node_cpus(node), mcpus = node_cpus(mem_node);
  printf("threads on node %d (%zu cpus), memory first-touched on node %d\n", node, cpus.size(), mem_node);
  double *a, *b, *c, *d;
  a = (double *)aligned_alloc(4096, N * 24); b = (double *)aligned_alloc(4096, N * 24); c = (double *)aligned_alloc(4096, N * 24); d = (double *)aligned_alloc(4096, N * 24);
  par(16, mcpus, [&](int t) { long lo = 3 * N * t / 16, hi = 3 * N * (t + 1) / 16; for (long i = lo; i < hi; ++i) { a[i] = i; b[i] = i; c[i] = i * 0.5; d[i] = 0; } });
  for (int T : {4, 8, 12, 15, 20, 24, 32}) {
    double best_r = 0, best_m = 0, best_nt = 0;
    for (int rep = 0; rep < 3; ++rep) {
      double t0 = now();
      par(T, cpus, [&](int t) { long lo = (3 * N * t / T) & ~3L, hi = (3 * N * (t + 1) / T) & ~3L; __m256d s = _mm256_setzero_pd();
        for (long i = lo; i < hi; i += 4) s = _mm256_xor_pd(s, _mm256_loadu_pd(a + i)); if (_mm256_cvtsd_f64(s) == 1.234) printf("x"); });
      best_r = std::max(best_r, N * 24 / (now() - t0) / 1e9);
      t0 = now();
      par(T, cpus, [&](int t) { long lo = (3 * N * t / T) & ~3L, hi = (3 * N * (t + 1) / T) & ~3L; int m = 0;
        for (long i = lo; i < hi; i += 4) { __m256i x = _mm256_xor_si256(_mm256_loadu_si256((__m256i *)(a + i)), _mm256_loadu_si256((__m256i *)(b + i))); m |= !_mm256_testz_si256(x, x);
          _mm256_storeu_si256((__m256i *)(b + i), _mm256_loadu_si256((__m256i *)(c + i))); } if (m == 7) printf("x"); });
      best_m = std::max(best_m, N * 96 / (now() - t0) / 1e9);
      par(T, cpus, [&](int t) { long lo = 3 * N * t / T, hi = 3 * N * (t + 1) / T; memcpy(b + lo, a + lo, (hi - lo) * 8); });
      t0 = now();
      par(T, cpus, [&](int t) { long lo = (3 * N * t / T) & ~3L, hi = (3 * N * (t + 1) / T) & ~3L;
        for (long i = lo; i < hi; i += 4) _mm256_stream_si256((__m256i *)(d + i), _mm256_loadu_si256((__m256i *)(c + i))); });
      best_nt = std::max(best_nt, N * 48 / (now() - t0) / 1e9);
    }
    printf("T=%2d  read %.0f GB/s   compare+refill (3 reads + 1 writeback) %.0f GB/s   nt copy (r+w) %.0f GB/s   [thread spawn included]\n", T, best_r, best_m, best_nt);
  }
}
