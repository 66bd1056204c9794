// Shared device-side building blocks of the dcb200 kernels (sm_100a only).
//
//  * mbarrier + 1-D bulk TMA (cp.async.bulk, SASS: UBLKCP) wrappers for the column-tile pipeline
//  * dist2_exact: the squared distance in the rounding order of the reference's CPU build
//    (reference loops density_clustering.cpp:171-176, :263-268, :315-318)
//  * tile geometry constants shared by the population / neighbour / screening kernels
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace dcb {

// ---------------------------------------------------------------- tile geometry ----------------
constexpr int RI = 4;                         // rows per thread (register block)
constexpr int CJ = 4;                         // columns per inner iteration (one LDS.128 per dim)
constexpr int N_CONSUMER_WARPS = 8;
constexpr int N_CONSUMERS = N_CONSUMER_WARPS * 32;
constexpr int CTA_THREADS = N_CONSUMERS + 32; // + one producer warp (TMA issue)
constexpr int ROWS_PER_CTA = N_CONSUMERS * RI;   // 1024
constexpr int SUPER_FRAMES = 4096;            // frames per super-tile: the coarse level of the producers' tile pruning
constexpr int LD_ALIGN = 256;                 // frame arrays are padded to a multiple of this (covers every tile width)
// depth of the tile ring: deep for the register kernels (consumer warps that skip or finish a tile early run ahead
// instead of idling), shallow for the run-time-D kernels whose tiles are large
#ifndef DCB_STAGES
#define DCB_STAGES 6
#endif
template <int D> struct StagesOf { static constexpr int n = D == 0 ? 3 : DCB_STAGES; };
constexpr int MAX_BINS = 31;                  // distinct radii per population pass (table of 32 incl. +inf)
constexpr int MAX_TEMPLATE_D = 16;            // dims held in registers by the specialised kernels

// ---------------------------------------------------------------- PTX wrappers -----------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// try_wait with a suspend-time hint: the thread sleeps in hardware until the phase completes (or the hint expires)
// instead of polling, so a waiting warp -- above all the producer, which is usually far ahead -- takes no issue slots
// from the warps that share its scheduler, and is woken by the arrival itself
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity), "r"(0x989680u)
      : "memory");
}
__device__ __forceinline__ void mbar_wait_backoff(uint64_t* bar, uint32_t parity) { mbar_wait(bar, parity); }
// the same with a real pause between polls, for a warp that is usually far ahead of the ones it waits for (a producer whose
// ring is full): ncu counted 5.6e9 poll iterations of the suspend-hinted loop per 258 ms launch, ~10 % of all instructions
// issued, in the very scheduler slots the consumer warps of that SM sub-partition need
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait_sleep(uint64_t* bar, uint32_t parity) {
  while (!mbar_test(bar, parity)) __nanosleep(400);
}
// 1-D bulk copy global -> shared, completion signalled on an mbarrier (bytes % 16 == 0, 16 B aligned)
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// ---------------------------------------------------------------- exact distance ---------------
// Squared Euclidean distance of frames i and j in the rounding order of the reference's CPU build
// (g++ -O3 -ffast-math, SSE2, no FMA): four lane accumulators over the first 4*floor(D/4) columns
// (multiply, then add), S = (a0+a2) + (a1+a3); a remainder of >= 2 columns goes into the two lane
// sums before the final add (lane1 + lane0), a last single column is added to S.
// xT is the dim-major coordinate array [D][ld].  Intrinsics keep nvcc from contracting to FMA.
static __device__ __noinline__ float dist2_exact(const float* __restrict__ xT, size_t ld, int D, uint32_t i, uint32_t j) {
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  int k = 0;
  for (; k + 4 <= D; k += 4) {
    const float* p = xT + (size_t) k * ld;
    float c0 = __fsub_rn(__ldg(p + i), __ldg(p + j));
    float c1 = __fsub_rn(__ldg(p + ld + i), __ldg(p + ld + j));
    float c2 = __fsub_rn(__ldg(p + 2 * ld + i), __ldg(p + 2 * ld + j));
    float c3 = __fsub_rn(__ldg(p + 3 * ld + i), __ldg(p + 3 * ld + j));
    a0 = __fadd_rn(a0, __fmul_rn(c0, c0));
    a1 = __fadd_rn(a1, __fmul_rn(c1, c1));
    a2 = __fadd_rn(a2, __fmul_rn(c2, c2));
    a3 = __fadd_rn(a3, __fmul_rn(c3, c3));
  }
  float l0 = __fadd_rn(a0, a2);
  float l1 = __fadd_rn(a1, a3);
  float s;
  if (D - k >= 2) {
    const float* p = xT + (size_t) k * ld;
    float c0 = __fsub_rn(__ldg(p + i), __ldg(p + j));
    float c1 = __fsub_rn(__ldg(p + ld + i), __ldg(p + ld + j));
    l0 = __fadd_rn(l0, __fmul_rn(c0, c0));
    l1 = __fadd_rn(l1, __fmul_rn(c1, c1));
    s = __fadd_rn(l1, l0);
    k += 2;
  } else {
    s = __fadd_rn(l0, l1);
  }
  if (k < D) {
    const float* p = xT + (size_t) k * ld;
    float c = __fsub_rn(__ldg(p + i), __ldg(p + j));
    s = __fadd_rn(s, __fmul_rn(c, c));
  }
  return s;
}

// The smallest float strictly greater than x (x finite, any sign); used to round thresholds up.
__device__ __forceinline__ float next_up(float x) {
  if (!(x == x) || x == __int_as_float(0x7f800000)) return x;
  if (x == 0.f) return __int_as_float(1);
  int b = __float_as_int(x);
  return __int_as_float(x > 0.f ? b + 1 : b - 1);
}

// ---------------------------------------------------------------- tile stream ------------------
// The producer warp streams column tiles through a StagesOf<D>::n-deep ring; each stage carries a small
// header telling the consumers which work item the tile belongs to.
struct TileMeta {
  int32_t row_block;     // -1: end of stream
  uint32_t col0;         // first column of the tile
  uint32_t flags;        // bit0: first tile of the item, bit1: last tile of the item
  uint32_t aux;          // kernel specific (e.g. neighbour search: tile class)
};

template <int NSTAGES>
struct Pipe {
  uint32_t stage = 0, phase = 0;
  __device__ __forceinline__ void advance() {
    if (++stage == NSTAGES) { stage = 0; phase ^= 1; }
  }
};

// Hand-back of ring stages from the consumer warps to the producer warp through the named barriers
// STAGE_BARRIER0 + stage: every consumer warp arrives (bar.arrive, does not wait) when it is done with a stage, the
// producer warp syncs on the barrier before it refills the stage.  A producer whose ring is full is parked by the
// hardware and takes no issue slots: with mbarrier polling (try_wait with a suspend hint, or test_wait + nanosleep --
// neither really sleeps) the waiting producers issued 14 % of all instructions of the C3 population scan
// (profiles/ncu_summary_r02_c3.txt), in the scheduler slots of the consumer warps they wait for.
// Whole warps execute these, converged.  Fill k (k = 0, 1, ...) uses stage k % NSTAGES and is handed back once; the
// producer syncs before fill k >= NSTAGES and drains what is still out at the end, so every barrier completes.
constexpr int STAGE_BARRIER0 = 2;             // 0: __syncthreads, 1: consumer_barrier
__device__ __forceinline__ void stage_release(uint32_t stage) {
  asm volatile("bar.arrive %0, %1;" ::"r"(STAGE_BARRIER0 + stage), "n"(CTA_THREADS) : "memory");
}
template <int NSTAGES>
struct FillPipe {
  static_assert(STAGE_BARRIER0 + NSTAGES <= 16, "one named barrier per stage");
  uint32_t stage = 0, fills = 0;
  // wait until the consumers are done with the previous content of the current stage
  __device__ __forceinline__ void acquire() const {
    if (fills >= (uint32_t) NSTAGES) asm volatile("bar.sync %0, %1;" ::"r"(STAGE_BARRIER0 + stage), "n"(CTA_THREADS) : "memory");
  }
  __device__ __forceinline__ void advance() {
    ++fills;
    if (++stage == NSTAGES) stage = 0;
  }
  // after the end-of-stream marker (which the consumers do not hand back): collect the hand-backs still out
  __device__ __forceinline__ void drain() {
    for (int q = 0; q < NSTAGES - 1; ++q) {
      advance();
      acquire();
    }
  }
};

}  // namespace dcb
