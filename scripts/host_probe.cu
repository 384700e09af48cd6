// Host-side probe for the GPU box: cores, NUMA layout, host memory bandwidth with T threads
// (read / copy / non-temporal copy), PCIe H2D bandwidth from pinned and pageable memory,
// cudaHostRegister cost, and a prototype of the "prepare" pass of the host-pointer path
// (compare origin with the previous destinations held in the pinned bounce buffer, refill it).
// Build: nvcc -O3 -std=c++17 -Xcompiler -pthread,-march=x86-64-v3 scripts/host_probe.cu -o gpurun_out/host_probe
#include <cuda_runtime.h>
#include <immintrin.h>
#include <sched.h>
#include <unistd.h>

#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <thread>
#include <vector>

static double now() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

static void par(int T, const std::function<void(int)> &f) {
  std::vector<std::thread> th;
  for (int t = 0; t < T; ++t) th.emplace_back(f, t);
  for (auto &x : th) x.join();
}

static void nt_copy(void *dst, const void *src, size_t bytes) {
  // 32-byte aligned dst assumed
  const char *s = (const char *)src;
  char *d = (char *)dst;
  size_t i = 0;
  for (; i + 64 <= bytes; i += 64) {
    __m256i a = _mm256_loadu_si256((const __m256i *)(s + i));
    __m256i b = _mm256_loadu_si256((const __m256i *)(s + i + 32));
    _mm256_stream_si256((__m256i *)(d + i), a);
    _mm256_stream_si256((__m256i *)(d + i + 32), b);
  }
  if (i < bytes) memcpy(d + i, s + i, bytes - i);
}

int main(int argc, char **argv) {
  const long N = argc > 1 ? atol(argv[1]) : 10000000;
  cpu_set_t set;
  sched_getaffinity(0, sizeof(set), &set);
  printf("online cpus %ld, affinity %d, hw_concurrency %u\n", sysconf(_SC_NPROCESSORS_ONLN), CPU_COUNT(&set),
         std::thread::hardware_concurrency());
  if (system("lscpu | egrep 'Model name|Socket|NUMA|Thread|Core|L3' ; cat /sys/devices/system/node/node*/meminfo 2>/dev/null | egrep 'MemTotal|MemFree'; "
             "cat /sys/fs/cgroup/cpu.max 2>/dev/null; nvidia-smi topo -m 2>/dev/null | head -12")) {}
  const size_t B = size_t(N) * 57;
  // caller-like pageable arrays, first-touched by this (single) thread like numpy would
  double *origin = (double *)malloc(N * 24), *dest = (double *)malloc(N * 24), *w = (double *)malloc(N * 8);
  int8_t *fly = (int8_t *)malloc(N);
  for (long i = 0; i < 3 * N; ++i) { origin[i] = i * 0.5; dest[i] = i * 0.25; }
  for (long i = 0; i < N; ++i) { w[i] = 1.0; fly[i] = (i % 20) != 0; }
  double *b_dest, *b_w;
  double t0 = now();
  cudaHostAlloc((void **)&b_dest, N * 24, cudaHostAllocDefault);
  cudaHostAlloc((void **)&b_w, N * 8, cudaHostAllocDefault);
  printf("cudaHostAlloc of %.0f MB: %.1f ms\n", N * 32 / 1e6, 1e3 * (now() - t0));
  memcpy(b_dest, origin, N * 24);  // so that 94% compare equal below
  for (long i = 0; i < N; i += 16) b_dest[3 * i] = -1.0;
  double *d_buf;
  cudaMalloc((void **)&d_buf, B + 64);
  cudaStream_t st;
  cudaStreamCreate(&st);

  for (int T : {1, 2, 4, 8, 16, 32, 64, 128}) {
    if (T > 2 * (int)std::thread::hardware_concurrency()) break;
    // (a) read
    std::atomic<long> sink{0};
    double best_r = 0, best_c = 0, best_nt = 0, best_prep = 0;
    for (int rep = 0; rep < 3; ++rep) {
      t0 = now();
      par(T, [&](int t) {
        const long lo = 3 * N * t / T, hi = 3 * N * (t + 1) / T;
        double s = 0;
        for (long i = lo; i < hi; ++i) s += origin[i];
        sink += (long)s;
      });
      best_r = std::max(best_r, N * 24 / (now() - t0) / 1e9);
      t0 = now();
      par(T, [&](int t) {
        const long lo = 3 * N * t / T, hi = 3 * N * (t + 1) / T;
        memcpy(b_dest + lo, dest + lo, (hi - lo) * 8);
      });
      best_c = std::max(best_c, N * 48 / (now() - t0) / 1e9);
      t0 = now();
      par(T, [&](int t) {
        const long lo = (3 * N * t / T) & ~3L, hi = t == T - 1 ? 3 * N : (3 * N * (t + 1) / T) & ~3L;
        nt_copy(b_dest + lo, dest + lo, (hi - lo) * 8);
      });
      best_nt = std::max(best_nt, N * 48 / (now() - t0) / 1e9);
      // (d) prototype prepare pass: per particle compare origin with bounce dest (previous), count
      // the changed ones, then overwrite the bounce with this move's dest and weight
      memcpy(b_dest, origin, N * 24);
      t0 = now();
      std::atomic<long> changed{0};
      par(T, [&](int t) {
        const long lo = (N * t / T) & ~3L, hi = t == T - 1 ? N : (N * (t + 1) / T) & ~3L;
        long c = 0;
        for (long i = lo; i < hi; ++i) {
          uint64_t a[3], m[3];
          memcpy(a, origin + 3 * i, 24);
          memcpy(m, b_dest + 3 * i, 24);
          c += (((a[0] ^ m[0]) | (a[1] ^ m[1]) | (a[2] ^ m[2])) != 0) & (fly[i] == 1);
        }
        nt_copy(b_dest + 3 * lo, dest + 3 * lo, (hi - lo) * 24);
        nt_copy(b_w + lo, w + lo, (hi - lo) * 8);
        changed += c;
      });
      best_prep = std::max(best_prep, 1.0 / (now() - t0));
    }
    printf("T=%3d  read %.1f GB/s  memcpy(pageable->pinned) %.1f GB/s (r+w)  nt-copy %.1f GB/s (r+w)  prepare pass %.2f ms / %ld particles\n",
           T, best_r, best_c, best_nt, 1e3 / best_prep, N);
  }
  // PCIe
  for (int rep = 0; rep < 2; ++rep) {
    t0 = now();
    cudaMemcpyAsync(d_buf, b_dest, N * 24, cudaMemcpyHostToDevice, st);
    cudaStreamSynchronize(st);
    double t1 = now() - t0;
    t0 = now();
    cudaMemcpyAsync(d_buf, dest, N * 24, cudaMemcpyHostToDevice, st);
    cudaStreamSynchronize(st);
    double t2 = now() - t0;
    printf("H2D %.0f MB: pinned %.2f ms (%.1f GB/s), pageable %.2f ms (%.1f GB/s)\n", N * 24 / 1e6, 1e3 * t1,
           N * 24 / t1 / 1e9, 1e3 * t2, N * 24 / t2 / 1e9);
  }
  t0 = now();
  cudaError_t e = cudaHostRegister(dest, N * 24, cudaHostRegisterDefault);
  printf("cudaHostRegister %.0f MB: %.1f ms (%s)\n", N * 24 / 1e6, 1e3 * (now() - t0), cudaGetErrorString(e));
  t0 = now();
  cudaMemcpyAsync(d_buf, dest, N * 24, cudaMemcpyHostToDevice, st);
  cudaStreamSynchronize(st);
  printf("H2D registered: %.2f ms (%.1f GB/s)\n", 1e3 * (now() - t0), N * 24 / (now() - t0) / 1e9);
  t0 = now();
  cudaHostUnregister(dest);
  printf("cudaHostUnregister: %.1f ms\n", 1e3 * (now() - t0));
  // pipelined: 10 chunks, prepare (T=16) chunk k+1 while chunk k is on the wire
  for (int T : {8, 16, 32}) {
    const int K = 10;
    t0 = now();
    for (int k = 0; k < K; ++k) {
      const long b = N * k / K, e2 = N * (k + 1) / K;
      par(T, [&](int t) {
        const long lo = b + (((e2 - b) * t / T) & ~3L), hi = t == T - 1 ? e2 : b + (((e2 - b) * (t + 1) / T) & ~3L);
        long c = 0;
        for (long i = lo; i < hi; ++i) {
          uint64_t a[3], m[3];
          memcpy(a, origin + 3 * i, 24);
          memcpy(m, b_dest + 3 * i, 24);
          c += (((a[0] ^ m[0]) | (a[1] ^ m[1]) | (a[2] ^ m[2])) != 0) & (fly[i] == 1);
        }
        nt_copy(b_dest + 3 * lo, dest + 3 * lo, (hi - lo) * 24);
        nt_copy(b_w + lo, w + lo, (hi - lo) * 8);
        if (c < 0) printf("x");
      });
      cudaMemcpyAsync(d_buf + 3 * b, b_dest + 3 * b, (e2 - b) * 24, cudaMemcpyHostToDevice, st);
      cudaMemcpyAsync(d_buf + 3 * N + b, b_w + b, (e2 - b) * 8, cudaMemcpyHostToDevice, st);
    }
    cudaStreamSynchronize(st);
    printf("pipelined prepare(T=%d, thread spawn per chunk)+H2D of 32 B/particle: %.2f ms\n", T, 1e3 * (now() - t0));
  }
  return 0;
}
