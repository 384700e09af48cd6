// How much of a randomly accessed record table stays in the B200's L2?  Every thread walks a private
// pseudo-random chain over a table of 128-byte records and reads three of the four 32-byte sectors of each
// (what the walk kernel does per crossing), with the L2 policy of the walk kernel (evict_last, no L1
// allocation) or the default one; optionally a second, streaming read (evict_first) runs alongside, as the
// particle arrays do.  Prints time per access and the implied hit rate is read with ncu:
//   ncu --metrics lts__t_sector_hit_rate.pct,dram__bytes_read.sum ./build/l2_probe
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a scripts/l2_probe.cu -o build/l2_probe
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ uint64_t policy(bool keep) {
  uint64_t p;
  if (keep) asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  else asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}

template <int MODE>  // 0 = default policy, 1 = evict_last + L1::no_allocate
__global__ void __launch_bounds__(128, 7) chase(const double *table, uint32_t nrec, int steps, const double *stream,
                                                size_t stream_doubles, double *sink) {
  uint32_t x = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
  const uint64_t keep = policy(true), first = policy(false);
  double acc = 0.0;
  size_t sp = (size_t)(blockIdx.x * blockDim.x + threadIdx.x) * 4;
  for (int s = 0; s < steps; ++s) {
    x = x * 1664525u + 1013904223u;
    const uint32_t r = (uint32_t)(((uint64_t)x * nrec) >> 32);
    const double *rec = table + (size_t)r * 16;
    const int skip = x & 3;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const int f = k + (k >= skip ? 1 : 0);
      double a, b, c, d;
      if (MODE == 1)
        asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f64 {%0,%1,%2,%3}, [%4], %5;"
                     : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(rec + 4 * f), "l"(keep));
      else
        asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(rec + 4 * f));
      acc += a + b + c + d;
    }
    if (stream && (s & 7) == 0) {  // ~1 streamed sector per 8 record accesses... scaled by the caller through stream_doubles
      double a, b, c, d;
      asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f64 {%0,%1,%2,%3}, [%4], %5;"
                   : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(stream + (sp % stream_doubles)), "l"(first));
      acc += a;
      sp += (size_t)gridDim.x * blockDim.x * 4;
    }
  }
  if (acc == 1.2345) *sink = acc;
}

int main() {
  const int blocks = 148 * 7, threads = 128, steps = 400;
  double *sink, *stream;
  cudaMalloc(&sink, 8);
  const size_t stream_bytes = size_t(1) << 30;
  cudaMalloc(&stream, stream_bytes);
  cudaMemset(stream, 0, stream_bytes);
  for (int mb : {16, 32, 48, 64, 80, 96, 112, 128, 160, 192, 256}) {
    const uint32_t nrec = uint32_t(size_t(mb) * 1024 * 1024 / 128);
    double *table;
    cudaMalloc(&table, size_t(nrec) * 128);
    cudaMemset(table, 0, size_t(nrec) * 128);
    for (int mode = 0; mode < 2; ++mode)
      for (int with_stream = 0; with_stream < 2; ++with_stream) {
        cudaEvent_t a, b;
        cudaEventCreate(&a); cudaEventCreate(&b);
        for (int rep = 0; rep < 2; ++rep) {  // the second launch finds the table as warm as it gets
          cudaEventRecord(a);
          if (mode) chase<1><<<blocks, threads>>>(table, nrec, steps, with_stream ? stream : nullptr, stream_bytes / 8, sink);
          else chase<0><<<blocks, threads>>>(table, nrec, steps, with_stream ? stream : nullptr, stream_bytes / 8, sink);
          cudaEventRecord(b);
          cudaEventSynchronize(b);
        }
        float ms;
        cudaEventElapsedTime(&ms, a, b);
        const double acc = double(blocks) * threads * steps;
        printf("table %3d MB  policy %-10s stream %d : %.3f ms, %.2f G record reads/s, %.1f TB/s of sectors\n", mb,
               mode ? "evict_last" : "default", with_stream, ms, acc / ms / 1e6, acc * 96 / ms / 1e9);
      }
    cudaFree(table);
  }
  return 0;
}
