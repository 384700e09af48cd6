// Does the B200's L2 hold one copy of a line or one per die?  scripts/l2_probe.cu found that a table read
// at random by all 148 SMs stays resident only up to ~63 MB, half of the 126 MB L2.  Here (1) every SM
// measures its latency to one 2 KB region (homed on one die): the SMs fall into a near and a far group =
// the two dies; (2) the record table is cut in two halves and either every SM reads the whole table, or
// the SMs of die 0 read only the first half and the SMs of die 1 only the second.  If the second layout
// keeps a 128 MB table resident, a die-aware split of the walk would double the usable L2.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a scripts/l2_probe2.cu -o build/l2_probe2
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

__global__ void latency_probe(const uint32_t *chain, int steps, unsigned long long *lat, unsigned *smid_out) {
  unsigned smid;
  asm("mov.u32 %0, %%smid;" : "=r"(smid));
  if (threadIdx.x != 0) return;
  uint32_t i = 0;
  for (int s = 0; s < 64; ++s) asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(i) : "l"(chain + i));  // warm
  const long long t0 = clock64();
  for (int s = 0; s < steps; ++s) asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(i) : "l"(chain + i));
  const long long t1 = clock64();
  lat[blockIdx.x] = (unsigned long long)(t1 - t0) + (i == 0xffffffffu);
  smid_out[blockIdx.x] = smid;
}

__global__ void __launch_bounds__(128, 7) chase(const double *table, uint32_t nrec, int steps, const int *die_of_sm,
                                                int split, double *sink) {
  unsigned smid;
  asm("mov.u32 %0, %%smid;" : "=r"(smid));
  uint32_t lo = 0, n = nrec;
  if (split) {  // this SM's die reads only its half of the table
    n = nrec / 2;
    lo = die_of_sm[smid] ? n : 0;
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
  // ---- (1) SM -> die
  const int nsm = 148;
  std::vector<uint32_t> h_chain(512);
  for (int i = 0; i < 512; ++i) h_chain[i] = (i * 37 + 11) % 512;  // a permutation walk inside one 2 KB region
  uint32_t *chain;
  cudaMalloc(&chain, 1 << 20);
  unsigned long long *lat;
  unsigned *smid;
  cudaMalloc(&lat, 1184 * 8);
  cudaMalloc(&smid, 1184 * 4);
  std::vector<double> best(nsm, 1e30);
  // several regions: each is homed on one die or the other; use the first whose latencies are clearly bimodal
  std::vector<int> die(nsm, 0);
  double gap_found = 0;
  for (int region = 0; region < 8 && gap_found < 8.0; ++region) {
    cudaMemcpy(chain + region * 4096, h_chain.data(), 2048, cudaMemcpyHostToDevice);
    std::fill(best.begin(), best.end(), 1e30);
    for (int rep = 0; rep < 5; ++rep) {
      latency_probe<<<1184, 32>>>(chain + region * 4096, 2000, lat, smid);
      std::vector<unsigned long long> h_lat(1184);
      std::vector<unsigned> h_sm(1184);
      cudaMemcpy(h_lat.data(), lat, 1184 * 8, cudaMemcpyDeviceToHost);
      cudaMemcpy(h_sm.data(), smid, 1184 * 4, cudaMemcpyDeviceToHost);
      for (int b = 0; b < 1184; ++b)
        if (h_sm[b] < (unsigned)nsm) best[h_sm[b]] = std::min(best[h_sm[b]], double(h_lat[b]) / 2000.0);
    }
    std::vector<double> s(best);
    std::sort(s.begin(), s.end());
    // largest gap between consecutive sorted latencies in the middle half
    double gap = 0, cut = 0;
    for (int i = nsm / 4; i < 3 * nsm / 4; ++i)
      if (s[i + 1] - s[i] > gap) { gap = s[i + 1] - s[i]; cut = 0.5 * (s[i] + s[i + 1]); }
    int n1 = 0;
    for (int i = 0; i < nsm; ++i) { die[i] = best[i] > cut; n1 += die[i]; }
    printf("region %d: latency min %.0f median %.0f max %.0f cycles, widest gap %.1f at %.0f -> %d near / %d far SMs\n", region,
           s[0], s[nsm / 2], s[nsm - 1], gap, cut, nsm - n1, n1);
    gap_found = gap;
  }
  printf("die of SM 0..147: ");
  for (int i = 0; i < nsm; ++i) printf("%d", die[i]);
  printf("\n");
  int *d_die;
  cudaMalloc(&d_die, nsm * 4);
  cudaMemcpy(d_die, die.data(), nsm * 4, cudaMemcpyHostToDevice);
  // ---- (2) whole table for everybody vs one half per die
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
      for (int rep = 0; rep < 2; ++rep) {
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
