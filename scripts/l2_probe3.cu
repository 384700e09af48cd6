// Which SMs share an L2 partition?  scripts/l2_probe.cu: a table read at random by all 148 SMs stays resident only up
// to ~63 MB = half of the L2, i.e. each die's partition keeps its own copy of what its SMs touch.  Then a line that
// only SMs of one die have read is a hit for that die and a miss for the other.  (1) One SM reads 148 disjoint
// 256 KB regions; then every SM chases a pointer chain through "its" region for the first time: SMs of the warming
// SM's die see L2-hit latency, the others DRAM latency.  Repeated with several warming SMs.  (2) With that map:
// a record table read either whole by every SM, or one half per die.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a scripts/l2_probe3.cu -o build/l2_probe3
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

constexpr int kSM = 148, kLines = 2048;             // lines of 128 B per region
constexpr size_t kRegionWords = size_t(kLines) * 32;  // uint32 words per region

__device__ __forceinline__ unsigned smid() { unsigned s; asm("mov.u32 %0, %%smid;" : "=r"(s)); return s; }

// every region: a random cyclic permutation over its lines (word 0 of line i holds the next line's index)
__global__ void build_chains(uint32_t *mem, int nregions) {
  const int r = blockIdx.x;
  if (r >= nregions || threadIdx.x) return;
  uint32_t *base = mem + size_t(r) * kRegionWords;
  uint32_t x = 12345u + 977u * r, cur = 0;
  // visit lines in the order of a full-period LCG over 2048
  for (int i = 0; i < kLines; ++i) {
    x = (x * 1664525u + 1013904223u);
    const uint32_t nxt = (cur * 5u + 1u) & (kLines - 1);  // full period mod 2^k (a=5, c=1)
    base[size_t(cur) * 32] = nxt;
    cur = nxt;
  }
}

__global__ void warm(const uint32_t *mem, size_t words, unsigned warm_sm, int *claimed) {
  if (smid() != warm_sm) return;
  __shared__ int mine;
  if (threadIdx.x == 0) mine = atomicCAS(claimed, 0, 1) == 0;
  __syncthreads();
  if (!mine) return;
  uint32_t acc = 0;
  for (int rep = 0; rep < 2; ++rep)
    for (size_t w = size_t(threadIdx.x) * 8; w < words; w += size_t(blockDim.x) * 8) {
      uint32_t v;
      asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(v) : "l"(mem + w));  // one load per 32-byte sector
      acc += v;
    }
  if (acc == 0xdeadbeefu) claimed[1] = 1;
}

__global__ void first_touch(const uint32_t *mem, float *lat, int *claimed) {
  const unsigned s = smid();
  if (threadIdx.x || s >= kSM) return;
  if (atomicCAS(claimed + s, 0, 1) != 0) return;  // one block per SM does the chase
  const uint32_t *base = mem + size_t(s) * kRegionWords;
  uint32_t i = 0;
  const long long t0 = clock64();
  for (int k = 0; k < kLines; ++k) asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(i) : "l"(base + size_t(i) * 32));
  const long long t1 = clock64();
  lat[s] = float(t1 - t0) / kLines + (i == 0xffffffffu);
}

__global__ void __launch_bounds__(128, 7) chase(const double *table, uint32_t nrec, int steps, const int *die_of_sm,
                                                int split, double *sink) {
  uint32_t lo = 0, n = nrec;
  if (split) {
    n = nrec / 2;
    lo = die_of_sm[smid()] ? n : 0;
  }
  uint32_t x = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
  double acc = 0.0;
  for (int s = 0; s < steps; ++s) {
    x = x * 1664525u + 1013904223u;
    const uint32_t r = lo + (uint32_t)(((uint64_t)x * n) >> 32);
    const double *rec = table + (size_t)r * 16;
    const int skip = x & 3;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const int f = k + (k >= skip ? 1 : 0);
      double a, b, c, d;
      asm volatile("ld.global.nc.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(rec + 4 * f));
      acc += a + b + c + d;
    }
  }
  if (acc == 1.2345) *sink = acc;
}

int main() {
  uint32_t *mem;
  const size_t words = kRegionWords * kSM;
  cudaMalloc(&mem, words * 4);
  cudaMemset(mem, 0, words * 4);
  build_chains<<<kSM, 32>>>(mem, kSM);
  float *lat;
  int *claimed;
  cudaMalloc(&lat, kSM * 4);
  cudaMalloc(&claimed, (kSM + 2) * 4);
  char *flush;
  const size_t flush_bytes = size_t(512) << 20;
  cudaMalloc(&flush, flush_bytes);
  std::vector<int> die(kSM, 0), votes(kSM, 0);
  int nruns = 0;
  std::vector<int> ref;
  for (unsigned warm_sm : {0u, 1u, 37u, 74u, 111u, 147u}) {
    cudaMemset(flush, 1, flush_bytes);  // evict everything
    cudaMemset(claimed, 0, (kSM + 2) * 4);
    warm<<<kSM * 8, 1024>>>(mem, words, warm_sm, claimed);
    cudaMemset(claimed, 0, (kSM + 2) * 4);
    cudaMemset(lat, 0, kSM * 4);
    first_touch<<<kSM * 16, 32>>>(mem, lat, claimed);
    std::vector<float> h(kSM);
    cudaMemcpy(h.data(), lat, kSM * 4, cudaMemcpyDeviceToHost);
    std::vector<float> s(h);
    std::sort(s.begin(), s.end());
    double gap = 0, cut = 0;
    for (int i = 8; i + 9 < kSM; ++i)
      if (s[i + 1] - s[i] > gap) { gap = s[i + 1] - s[i]; cut = 0.5 * (s[i] + s[i + 1]); }
    std::vector<int> near(kSM);
    int nnear = 0;
    for (int i = 0; i < kSM; ++i) { near[i] = h[i] > 0 && h[i] < cut; nnear += near[i]; }
    printf("warmed by SM %3u: first-touch latency min %.0f, median %.0f, max %.0f cycles; widest gap %.0f at %.0f; %d SMs hit (SM %u itself: %.0f)\n",
           warm_sm, s[0], s[kSM / 2], s[kSM - 1], gap, cut, nnear, warm_sm, h[warm_sm]);
    printf("  hit map: ");
    for (int i = 0; i < kSM; ++i) printf("%d", near[i]);
    printf("\n");
    // orient every run like the first one (die 0 = the die of SM 0) and vote
    if (ref.empty()) ref = near;
    int agree = 0;
    for (int i = 0; i < kSM; ++i) agree += near[i] == ref[i];
    const bool flip = agree < kSM / 2;
    for (int i = 0; i < kSM; ++i) votes[i] += (flip ? !near[i] : near[i]);
    ++nruns;
  }
  int n0 = 0;
  for (int i = 0; i < kSM; ++i) { die[i] = votes[i] * 2 > nruns ? 0 : 1; n0 += die[i] == 0; }
  printf("die of SM 0..147 (majority of %d runs; %d / %d): ", nruns, n0, kSM - n0);
  for (int i = 0; i < kSM; ++i) printf("%d", die[i]);
  printf("\n");
  int *d_die;
  cudaMalloc(&d_die, kSM * 4);
  cudaMemcpy(d_die, die.data(), kSM * 4, cudaMemcpyHostToDevice);
  double *sink;
  cudaMalloc(&sink, 8);
  const int blocks = 148 * 7, threads = 128, steps = 400;
  for (int mb : {64, 96, 128, 160, 192, 256}) {
    const uint32_t nrec = uint32_t(size_t(mb) * 1024 * 1024 / 128);
    double *table;
    cudaMalloc(&table, size_t(nrec) * 128);
    cudaMemset(table, 0, size_t(nrec) * 128);
    for (int split = 0; split < 2; ++split) {
      cudaEvent_t a, b;
      cudaEventCreate(&a); cudaEventCreate(&b);
      for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(a);
        chase<<<blocks, threads>>>(table, nrec, steps, d_die, split, sink);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
      }
      float ms;
      cudaEventElapsedTime(&ms, a, b);
      printf("table %3d MB  %-22s : %.3f ms\n", mb, split ? "one half per die" : "whole table, every SM", ms);
    }
    cudaFree(table);
  }
  return 0;
}
