// sm_100a kernels of the track-length tally engine.
//
// One fused kernel per particle range does what the reference spreads over
// K1..K12 of SURVEY.md section 2b: the "set dest" kernels
// (PumiTallyImpl.cpp:88-109, 127-142, 186-192), both SearchAndRebuild loops
// (PumiTallyImpl.cpp:433-459 -> external tracer) and the per-iteration functor
// (PumiTallyImpl.cpp:297-316).  Each particle is walked to completion in
// registers; the only global traffic per crossing is one 128-byte tet record
// and one fp64 reduction into flux[elem].
//
// Three record-fetch strategies are compiled (WalkVariant); all share the same
// per-ray state machine below so they produce identical results.
#include "walk_kernels.hpp"

#include <cstdint>

#include "walk_core.cuh"

namespace ptb {
namespace {

__device__ __forceinline__ void flush_counters(const WalkParams &P, const Counters &c) {
  const unsigned segs = __reduce_add_sync(0xffffffffu, c.segs);
  const unsigned tracks = __reduce_add_sync(0xffffffffu, c.tracks);
  const unsigned relocs = __reduce_add_sync(0xffffffffu, c.relocs);
  const unsigned lost = __reduce_add_sync(0xffffffffu, c.lost);
  if ((threadIdx.x & 31) == 0) {
    if (segs) atomicAdd(&P.stats->segments, (unsigned long long)segs);
    if (tracks) atomicAdd(&P.stats->tracks, (unsigned long long)tracks);
    if (relocs) atomicAdd(&P.stats->relocations, (unsigned long long)relocs);
    if (lost) atomicAdd(&P.stats->lost, (unsigned long long)lost);
  }
}

// ---------------------------------------------------------------- variant 0
// Thread per particle; the 128-byte record arrives as four 256-bit loads
// (LDG.E.ENL2.256), one per face.

__device__ __forceinline__ void load_face_256(const double *p, double &a, double &b, double &c,
                                              double &d) {
  asm("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p));
}

__global__ void __launch_bounds__(256) walk_ldg_kernel(const WalkParams P) {
  const int i = P.begin + blockIdx.x * blockDim.x + threadIdx.x;
  Counters c;
  Ray r;
  r.stage = kStageDone;
  if (i < P.end) begin_particle(P, i, r, c, true);
  while (r.stage != kStageDone) {
    const double *rec = P.tets[r.e].d;
    double raw[16];
#pragma unroll
    for (int f = 0; f < 4; ++f)
      load_face_256(rec + 4 * f, raw[4 * f], raw[4 * f + 1], raw[4 * f + 2], raw[4 * f + 3]);
    TetPlanes t;
    decode_record(raw, t);
    double texit;
    int32_t next;
    exit_face(t, r.ox, r.oy, r.oz, r.ux, r.uy, r.uz, texit, next);
    advance(P, i, r, texit, next, c, true);
  }
  flush_counters(P, c);
}

// ---------------------------------------------------------------- variant 1
// Thread per particle; each lane's record is staged into its own shared-memory
// row by one cp.async.bulk (TMA unit, bypasses L1/LSU), completion signalled
// on a per-warp mbarrier.  Rows are 144 bytes apart so the eight 16-byte reads
// of a quarter-warp hit distinct bank groups.

constexpr int kRowBytes = 144;

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes,
                                         uint32_t bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
          "r"(dst),
      "l"(src), "r"(bytes), "r"(bar)
      : "memory");
}

template <int BLOCK>
__global__ void __launch_bounds__(BLOCK) walk_bulk_kernel(const WalkParams P) {
  constexpr int WARPS = BLOCK / 32;
  __shared__ __align__(128) unsigned char rows[WARPS][32 * kRowBytes];
  __shared__ __align__(8) unsigned long long bars[WARPS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t bar = smem_u32(&bars[warp]);
  const uint32_t row = smem_u32(&rows[warp][lane * kRowBytes]);
  if (lane == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncwarp();

  const int i = P.begin + blockIdx.x * BLOCK + threadIdx.x;
  Counters c;
  Ray r;
  r.stage = kStageDone;
  if (i < P.end) begin_particle(P, i, r, c, true);
  uint32_t parity = 0;
  for (;;) {
    // the ballot is also the point where every lane has left the previous wait,
    // so the leader may re-arm the barrier
    const unsigned act = __ballot_sync(0xffffffffu, r.stage != kStageDone);
    if (!act) break;
    if (lane == __ffs(act) - 1) mbar_expect_tx(bar, 128u * __popc(act));
    if (r.stage != kStageDone) bulk_g2s(row, P.tets + r.e, 128u, bar);
    mbar_wait(bar, parity);
    parity ^= 1u;
    if (r.stage != kStageDone) {
      double raw[16];
#pragma unroll
      for (int j = 0; j < 8; ++j)
        asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];"
                     : "=d"(raw[2 * j]), "=d"(raw[2 * j + 1])
                     : "r"(row + 16 * j));
      TetPlanes t;
      decode_record(raw, t);
      double texit;
      int32_t next;
      exit_face(t, r.ox, r.oy, r.oz, r.ux, r.uy, r.uz, texit, next);
      advance(P, i, r, texit, next, c, true);
    }
  }
  flush_counters(P, c);
}

// ---------------------------------------------------------------- variant 2
// Four lanes per particle, lane f owns face f: each lane loads its 32-byte
// plane (the quad's four loads cover the record's 128-byte line exactly),
// evaluates one num/den, and the quad agrees on the exit with two shuffles.

__global__ void __launch_bounds__(256) walk_quad_kernel(const WalkParams P) {
  const int lane = threadIdx.x & 31;
  const int f = lane & 3;
  const unsigned qmask = 0xfu << (lane & ~3);
  const long long gt = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long ii = (long long)P.begin + (gt >> 2);
  const int i = (int)ii;
  const bool writer = (f == 0);
  Counters c;
  Ray r;
  r.stage = kStageDone;
  if (ii < (long long)P.end) begin_particle(P, i, r, c, writer);
  while (r.stage != kStageDone) {
    double a, b, cc, d;
    load_face_256(P.tets[r.e].d + 4 * f, a, b, cc, d);
    const uint32_t nbu = (dlo(a) & 0xffu) | ((dlo(b) & 0xffu) << 8) | ((dlo(cc) & 0xffu) << 16) |
                         ((dlo(d) & 0xffu) << 24);
    const double nx = dmask(a), ny = dmask(b), nz = dmask(cc), pc = dmask(d);
    const double den = nx * r.ux + ny * r.uy + nz * r.uz;
    const double num = pc - (nx * r.ox + ny * r.oy + nz * r.oz);
    const bool out = den > 0.0;
    double tb = out ? num / den : __builtin_huge_val();
    int32_t nb = out ? (int32_t)nbu : -2;
#pragma unroll
    for (int m = 1; m <= 2; m <<= 1) {
      const double to = __shfl_xor_sync(qmask, tb, m);
      const int32_t no = __shfl_xor_sync(qmask, nb, m);
      // ties go to the lower face index, as in the sequential scan of exit_face()
      const bool take = (to < tb) || (to == tb && (lane & m));
      tb = take ? to : tb;
      nb = take ? no : nb;
    }
    advance(P, i, r, tb, nb, c, writer);
  }
  flush_counters(P, c);
}

// ------------------------------------------------------------ small kernels

// K14-K16 of SURVEY.md 2b (PumiTallyImpl.cpp:492-528): every particle starts
// at the centroid of element 0.
__global__ void init_particles_kernel(double *px, double *py, double *pz, int32_t *elem, int32_t n,
                                      double cx, double cy, double cz) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    px[i] = cx; py[i] = cy; pz[i] = cz;
    elem[i] = 0;
  }
}

// K13 (PumiTallyImpl.cpp:393-405) with volumes precomputed at mesh build.
__global__ void normalize_kernel(const double *flux, const double *volume, double *out, int64_t n) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e < n) out[e] = flux[e] / volume[e];
}

}  // namespace

cudaError_t launch_walk(const WalkParams &p, int variant, int block, cudaStream_t stream) {
  const long long n = (long long)p.end - p.begin;
  if (n <= 0) return cudaSuccess;
  if (block != 64 && block != 128 && block != 256) block = 128;
  switch (variant) {
    case kVariantLdg: {
      const unsigned grid = (unsigned)((n + block - 1) / block);
      walk_ldg_kernel<<<grid, block, 0, stream>>>(p);
      break;
    }
    case kVariantBulk: {
      const unsigned grid = (unsigned)((n + block - 1) / block);
      if (block == 64) walk_bulk_kernel<64><<<grid, 64, 0, stream>>>(p);
      else if (block == 128) walk_bulk_kernel<128><<<grid, 128, 0, stream>>>(p);
      else walk_bulk_kernel<256><<<grid, 256, 0, stream>>>(p);
      break;
    }
    case kVariantQuad: {
      const unsigned grid = (unsigned)((4 * n + block - 1) / block);
      walk_quad_kernel<<<grid, block, 0, stream>>>(p);
      break;
    }
    default:
      return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

cudaError_t launch_init_particles(double *px, double *py, double *pz, int32_t *elem, int32_t n,
                                  double cx, double cy, double cz, cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  init_particles_kernel<<<(n + 255) / 256, 256, 0, stream>>>(px, py, pz, elem, n, cx, cy, cz);
  return cudaGetLastError();
}

cudaError_t launch_normalize(const double *flux, const double *volume, double *out, int64_t n,
                             cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  normalize_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(flux, volume, out, n);
  return cudaGetLastError();
}

}  // namespace ptb
