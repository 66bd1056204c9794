// GEMM-form pair scans for high-dimensional inputs (17 <= n_cols <= 256), written for sm_100a tensor cores:
//
//   gscan_pops_kernel<NB>   multi-radius neighbourhood population count   (density_clustering.cpp:126-195)
//   gscan_nn_kernel         nearest neighbour + nearest neighbour with lower free energy (density_clustering.cpp:230-288)
//
// Where the path really is a dense contraction the squared distance is evaluated as
//     F = |x~|^2 + |y~|^2 - 2 x~.y~ ,        x~ = tf32(x - centre)
// with the dot products on the 5th-generation tensor cores: tcgen05.mma kind::tf32, 128 x 128 x 8 per instruction,
// operands in shared memory (K-major, 128-byte swizzle), accumulators in tensor memory (two 128-column stages), read back
// by the epilogue warps with tcgen05.ld (one thread = one row of the accumulator).  F is the exact squared distance of
// the ROUNDED points up to FP32 accumulation error, so its distance to the true value is ~ 2 d (|dx| + |dy|) with
// |dx| <= 2^-11 |x - centre|: proportional to d, small for close pairs.  As in the FFMA kernels the fast value never
// decides alone: a pair whose F lies within the proven band of a decision boundary is re-evaluated with dist2_exact_coop,
// the reference's own arithmetic, so that populations and neighbours stay bit-identical.
//
// Structure of a CTA (one per SM, persistent):
//   warp 0      producer: claims work items (row tile x range of column tiles), drops the column tiles whose precomputed
//               lower bound (tile_lb_kernel: boxes and spheres in all n_cols dims) is out of reach, streams the others with
//               1-D bulk TMA: the row tile's operand image once per item, the column tiles' images in 16 KB K-chunks
//               through a ring, plus a small per-tile side record (|y~|^2, free-energy ranks)
//   warp 1      MMA issuer: one elected lane issues tcgen05.mma for every chunk and commits to the mbarriers that
//               free the ring slot / publish the accumulator stage
//   warps 2..9  two epilogue warpgroups, one per accumulator stage: tcgen05.ld 32 columns at a time, branch-free
//               count / filter logic per pair, compact exact handler for the rare band pairs
// The operand images are written by gpack_kernel exactly as they must sit in shared memory, so that no tensor map is
// needed: a chunk is one contiguous 16 KB block = one cp.async.bulk.
#pragma once
#include "common.cuh"
#include <float.h>

namespace dcb {

constexpr int GT = 128;                      // rows / columns per tile (= UMMA M = UMMA N)
constexpr int GK = 32;                       // floats per K-chunk row = 128 bytes = the swizzle width
constexpr int G_CHUNK_FLOATS = GT * GK;      // 4096 floats = 16 KB
constexpr int G_CHUNK_BYTES = G_CHUNK_FLOATS * 4;
constexpr int G_MIN_D = 17, G_MAX_D = 256;      // below: the register kernels (n_cols <= 16) are faster
constexpr int G_SIDE_SLOTS = 8;              // per-tile side records in flight
constexpr int G_ACC = 4;                     // accumulator stages in tensor memory (4 x 128 columns = all of it)
constexpr int G_SIDE_FLOATS = 4 + 2 * GT;    // meta (16 B), |y~|^2 [128], rank floats [128]
constexpr int G_EPI_WARPS = 16;
constexpr int G_EPI_THREADS = G_EPI_WARPS * 32;
constexpr int G_THREADS = 64 + G_EPI_THREADS + 32;  // producer warp, MMA warp, four epilogue warpgroups, second MMA warp (ra == 2)
constexpr int G_MMA2_WARP = 2 + G_EPI_WARPS;        // warp index of the second MMA issuer
constexpr int G_LD = 16;                     // accumulator columns per TMEM load
constexpr int G_HALF = 8;                    // columns per band / hit test

// ------------------------------------------------------------------------------------------------
// launch arguments (shared with the host code in api.cu)
// ------------------------------------------------------------------------------------------------
struct GemmGeom {
  const float* gT;            // operand images [tiles][kc][4096]
  const float* gnorm;         // [ld]
  const float* xR;            // [ld][d]
  const float* lb;            // [row tiles of this launch][n_tiles]
  int d, kc, k8;              // dims, K-chunks per tile, total K=8 MMA steps (ceil(d/8))
  int n_stages;               // ring depth (chunks), a multiple of cb
  int cb;                     // chunks per commit group (a power of two): the MMA warp frees ring slots cb at a time (tcgen05.commit is not free)
  int cb_log2;
  uint32_t n;                 // real frame count
  uint32_t n_tiles;           // column tiles (= ld / 128)
  uint32_t row_begin, row_end;   // positions of this shard, row_begin % 128 == 0
  uint32_t n_row_tiles;       // row BLOCKS of this launch: ra row tiles (128 ra rows) each
  uint32_t n_row_tiles128;    // 128-row tiles of this launch that hold at least one row (rows of g.lb)
  int ra;                     // row tiles per work item: 2 when two operand images fit (every streamed column chunk then feeds
                              // 256 rows: the L2 -> SM stream, ~50 B/clk per SM, is what limits the scan otherwise), else 1
  uint32_t tiles_per_item, n_col_items;
  unsigned int* work_counter;
  unsigned long long* stats;  // [0] pairs handed to the handler, [1] exact evaluations, [2] tiles streamed, [3] tile scans
  float rho_c;                // rho = rho_c (sqrt(|x~|^2) + sqrt(max |y~|^2)) bounds |dx| + |dy|
  float nymax;                // max |y~|^2 over all frames (rounded up)
  float c_acc;                // eps = c_acc (|x~|^2 + nymax): accumulation / FP32 rounding part of the band
  float e_rel;                // relative error of the reference's own arithmetic against the real-number distance
  float prune_slack;          // absolute slack of the tile lower bounds
  float prune_thr;            // static pruning threshold (d2 units); +inf: none
  unsigned long long* prof;   // optional [16] cycle counters per role (diagnostics, DCB200_GEMM_PROF=1), see g_prof_names in api.cu
  float* check;               // optional [2]: max observed |F - d2e| / band (float bits, atomicMax), CHECK builds only
};

struct GPopsArgs {
  GemmGeom g;
  int n_bins;
  float rad2[8];             // squared radii of this pass; unused slots -1 (nothing is ever inside)
  float rad[8];              // sqrt(rad2), rounded up; unused 0
  uint32_t* cnt;             // [n_bins][ld_cnt]  #{j : d2(i,j) < rad2[b]} including the frame itself when rad2[b] > 0
  size_t ld_cnt;
};

struct GNnArgs {
  GemmGeom g;
  const uint32_t* perm;         // [n] position -> frame
  const uint32_t* lo;           // [n] number of frames with a strictly lower free energy, by position
  const float* lof;             // [ld] (float)(lo >> shift), +inf padded: streamed with every tile
  const float* lomin;           // [n_tiles] min of lof over the tile
  float lo_bias;
  uint32_t window;              // > 0: only column tiles within `window` tiles of the row tile (first pass)
  unsigned long long* key_nn;   // [row_end - row_begin] (d2 bits << 32 | frame), atomicMin'ed
  unsigned long long* key_hd;
  float* thr_nn;                // [n_row_tiles][4] per 32-row quarter of a row tile: bound (d2 units, margins included) of what
                                // its rows still accept as nearest neighbour; lowered by atomicMin as items finish
  float* thr_hd;                // the same for the lower-free-energy neighbour; 0 where no row can have one
  float* lormax;                // [n_row_tiles][4] largest filter rank (+ bias) among the rows that can have such a neighbour
};

// everything below is device code, compiled in gemm_inst.cu only
#ifdef DCB_GEMM_KERNELS

// ------------------------------------------------------------------------------------------------
// tcgen05 / TMEM wrappers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
// arrives on the mbarrier once every tcgen05.mma issued so far by this thread has completed
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[smem] . B[smem]^T, 128 x 128 x 8, TF32 inputs, FP32 accumulation
__device__ __forceinline__ void tc_mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// 16 consecutive accumulator columns of this thread's TMEM lane: asynchronous load, then a wait that every later use of
// the registers depends on (the "+r" operands), so that the next load can be in flight while a chunk is processed
#define G_R16(X, r) X(r[0]), X(r[1]), X(r[2]), X(r[3]), X(r[4]), X(r[5]), X(r[6]), X(r[7]), X(r[8]), X(r[9]), X(r[10]), X(r[11]), X(r[12]), \
    X(r[13]), X(r[14]), X(r[15])
#define G_OUT(x) "=r"(x)
#define G_INOUT(x) "+r"(x)
__device__ __forceinline__ void tmem_ld16_issue(uint32_t taddr, uint32_t (&r)[G_LD]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : G_R16(G_OUT, r)
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16_wait(uint32_t (&r)[G_LD]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;" : G_R16(G_INOUT, r)::"memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "elect.sync _|p, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}
// warp index as a value the compiler knows to be warp-uniform (role dispatch, uniform-datapath operands of tcgen05.mma)
__device__ __forceinline__ int uniform_warp() { return __shfl_sync(0xffffffffu, (int) (threadIdx.x >> 5), 0); }

// Shared-memory matrix descriptor of a [128 rows][32 floats] K-major operand chunk with 128-byte swizzle:
// start address >> 4 in bits [0,14), leading byte offset (unused for swizzled K-major) 1 in [16,30), stride byte offset
// (8 rows x 128 B = 1024 B) >> 4 in [32,46), descriptor version 1 in [46,48), layout type SWIZZLE_128B = 2 in [61,64).
// A K step of 8 floats (32 bytes) inside the swizzle atom advances the start address by 32 bytes.
__device__ __forceinline__ uint64_t g_smem_desc(uint32_t smem_addr) {
  uint64_t d = (uint64_t) ((smem_addr >> 4) & 0x3fffu);
  d |= (uint64_t) 1 << 16;
  d |= (uint64_t) (1024 >> 4) << 32;
  d |= (uint64_t) 1 << 46;
  d |= (uint64_t) 2 << 61;
  return d;
}
// instruction descriptor: D format F32 (1 << 4), A and B format TF32 (2 << 7, 2 << 10), both K-major, N >> 3 at bit 17, M >> 4 at bit 24
constexpr uint32_t G_IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t) (GT >> 3) << 17) | ((uint32_t) (GT >> 4) << 24);

__device__ __forceinline__ float to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

// Squared distance of positions i and j in the reference's rounding order (see dist2_exact, common.cuh) on the row-major copy
// xR [ld][D] of the original coordinates in context order, computed by a whole warp for ONE pair (i, j): lane l takes columns 4l .. 4l+3 (coalesced 16-byte loads of
// the two rows instead of 2 D scattered 4-byte loads by one thread), and the four lane accumulators of the reference's
// loop are then summed in column order through shuffles, every lane running the identical chain: the reference's rounding
// order, the same bits in all lanes.  Candidates are rare (usually one lane of a warp at a time), so one pair per call costs
// ~1/3 of the instructions and 1/8 of the memory sectors of a per-thread loop (which lost even with many lanes active).
static __device__ __noinline__ float dist2_exact_coop(const float* __restrict__ xR, int D, uint32_t i, uint32_t j, int lane) {
  const float* __restrict__ a = xR + (size_t) i * D;
  const float* __restrict__ b = xR + (size_t) j * D;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  const int n4 = D >> 2;
  const bool vec = (D & 3) == 0;                      // rows are 16-byte aligned
  for (int base = 0; base < n4; base += 32) {
    const int gq = base + lane;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    if (gq < n4) {
      float4 xa, xb;
      if (vec) {
        xa = __ldg(reinterpret_cast<const float4*>(a) + gq);
        xb = __ldg(reinterpret_cast<const float4*>(b) + gq);
      } else {
        xa = make_float4(__ldg(a + 4 * gq), __ldg(a + 4 * gq + 1), __ldg(a + 4 * gq + 2), __ldg(a + 4 * gq + 3));
        xb = make_float4(__ldg(b + 4 * gq), __ldg(b + 4 * gq + 1), __ldg(b + 4 * gq + 2), __ldg(b + 4 * gq + 3));
      }
      const float c0 = __fsub_rn(xa.x, xb.x), c1 = __fsub_rn(xa.y, xb.y), c2 = __fsub_rn(xa.z, xb.z), c3 = __fsub_rn(xa.w, xb.w);
      s0 = __fmul_rn(c0, c0); s1 = __fmul_rn(c1, c1); s2 = __fmul_rn(c2, c2); s3 = __fmul_rn(c3, c3);
    }
    const int cnt = min(32, n4 - base);
    if (cnt == 32) {
      // straight-line: the 128 shuffles do not depend on the four add chains and pipeline ahead of them (a rolled loop
      // exposed one shuffle latency per column group: ~1100 cycles per pair, which stalls the whole accumulator hand-off)
#pragma unroll
      for (int l = 0; l < 32; ++l) {
        a0 = __fadd_rn(a0, __shfl_sync(0xffffffffu, s0, l));
        a1 = __fadd_rn(a1, __shfl_sync(0xffffffffu, s1, l));
        a2 = __fadd_rn(a2, __shfl_sync(0xffffffffu, s2, l));
        a3 = __fadd_rn(a3, __shfl_sync(0xffffffffu, s3, l));
      }
    } else {
#pragma unroll 4
      for (int l = 0; l < cnt; ++l) {
        a0 = __fadd_rn(a0, __shfl_sync(0xffffffffu, s0, l));
        a1 = __fadd_rn(a1, __shfl_sync(0xffffffffu, s1, l));
        a2 = __fadd_rn(a2, __shfl_sync(0xffffffffu, s2, l));
        a3 = __fadd_rn(a3, __shfl_sync(0xffffffffu, s3, l));
      }
    }
  }
  int k = 4 * n4;
  float l0 = __fadd_rn(a0, a2);
  float l1 = __fadd_rn(a1, a3);
  float s;
  if (D - k >= 2) {
    const float c0 = __fsub_rn(__ldg(a + k), __ldg(b + k));
    const float c1 = __fsub_rn(__ldg(a + k + 1), __ldg(b + k + 1));
    l0 = __fadd_rn(l0, __fmul_rn(c0, c0));
    l1 = __fadd_rn(l1, __fmul_rn(c1, c1));
    s = __fadd_rn(l1, l0);
    k += 2;
  } else {
    s = __fadd_rn(l0, l1);
  }
  if (k < D) {
    const float c = __fsub_rn(__ldg(a + k), __ldg(b + k));
    s = __fadd_rn(s, __fmul_rn(c, c));
  }
  return s;
}

// ------------------------------------------------------------------------------------------------
// layout: operand images, norms, row-major exact copy, tile geometry
// ------------------------------------------------------------------------------------------------
// One block (128 threads) per tile of 128 consecutive positions.
//   gT    [tiles][kc][128][32]  tf32(x - centre), K-major rows of 128 bytes, 16-byte pieces XOR-swizzled with (row & 7):
//                               the canonical SWIZZLE_128B layout; K padded with zeros; padded positions: zeros
//   gnorm [ld]                  |x~|^2 of the rounded values (FP32 FMA chain), +inf for padded positions
//   xR    [ld][d]               original coordinates in context order (padding: NaN)
//   tcen / tlo / thi [d][tiles] mean and bounding box of the tile in centred coordinates (x - centre), trad [tiles] the radius
//                               of the tile's frames around tcen (rounded up); empty tiles: box (+inf,-inf), radius 0
__global__ void gpack_kernel(const float* __restrict__ coords, size_t n, int d, int kc, size_t n_tiles, const float* __restrict__ centre,
                             const uint32_t* __restrict__ perm, float* __restrict__ gT, float* __restrict__ gnorm, float* __restrict__ xR,
                             float* __restrict__ tcen, float* __restrict__ tlo, float* __restrict__ thi, float* __restrict__ trad) {
  extern __shared__ float sm[];                      // [128][pitch] centred coordinates, then [d] tile centre
  const int pitch = d + 1 + (d & 1);                 // odd: conflict-free both along rows and along dims
  float* cen = sm + (size_t) GT * pitch;
  __shared__ float red[4];
  const size_t tile = blockIdx.x;
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const size_t p0 = tile * GT;
  const int real_rows = p0 < n ? (int) (n - p0 < (size_t) GT ? n - p0 : (size_t) GT) : 0;
  const float nan = __int_as_float(0x7fc00000);
  // rows: one warp per row, coalesced along the dims
  for (int r = warp; r < GT; r += 4) {
    const size_t p = p0 + r;
    const bool real = r < real_rows;
    const size_t src = real ? (perm ? (size_t) perm[p] : p) : 0;
    for (int k = lane; k < d; k += 32) {
      const float x = real ? coords[src * d + k] : nan;
      xR[p * d + k] = x;
      sm[(size_t) r * pitch + k] = real ? x - centre[k] : 0.f;
    }
  }
  __syncthreads();
  // per dim: mean and box
  for (int k = t; k < d; k += GT) {
    float s = 0.f, lo = INFINITY, hi = -INFINITY;
    for (int r = 0; r < real_rows; ++r) {
      const float v = sm[(size_t) r * pitch + k];
      s += v;
      lo = fminf(lo, v);
      hi = fmaxf(hi, v);
    }
    float c = real_rows ? s / (float) real_rows : 0.f;
    if (!(fabsf(c) < FLT_MAX)) c = 0.f;
    cen[k] = c;
    tcen[(size_t) k * n_tiles + tile] = c;
    tlo[(size_t) k * n_tiles + tile] = lo;
    thi[(size_t) k * n_tiles + tile] = hi;
  }
  __syncthreads();
  // per row: norm of the rounded values, distance to the tile centre
  {
    float nrm = 0.f, dc = 0.f;
    for (int k = 0; k < d; ++k) {
      const float v = sm[(size_t) t * pitch + k];
      const float vr = to_tf32(v);
      nrm = fmaf(vr, vr, nrm);
      const float e = v - cen[k];
      dc = fmaf(e, e, dc);
    }
    const bool real = t < real_rows;
    gnorm[p0 + t] = real ? nrm : INFINITY;
    float rad = real ? sqrtf(dc) * 1.00001f + 1e-30f : 0.f;
    for (int o = 16; o > 0; o >>= 1) rad = fmaxf(rad, __shfl_xor_sync(0xffffffffu, rad, o));
    if (lane == 0) red[warp] = rad;
  }
  __syncthreads();
  if (t == 0) trad[tile] = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
  // operand image: 8 consecutive threads write the 8 sixteen-byte pieces of one 128-byte row
  float* __restrict__ img = gT + tile * (size_t) kc * G_CHUNK_FLOATS;
  for (int q = 0; q < kc; ++q) {
    for (int pass = 0; pass < GT / 16; ++pass) {
      const int r = pass * 16 + (t >> 3);
      const int pc = t & 7;                       // physical piece
      const int lc = pc ^ (r & 7);                // logical piece: columns lc*4 .. lc*4+3 of the chunk
      float4 v;
      const int k0 = q * GK + lc * 4;
      v.x = k0 + 0 < d ? to_tf32(sm[(size_t) r * pitch + k0 + 0]) : 0.f;
      v.y = k0 + 1 < d ? to_tf32(sm[(size_t) r * pitch + k0 + 1]) : 0.f;
      v.z = k0 + 2 < d ? to_tf32(sm[(size_t) r * pitch + k0 + 2]) : 0.f;
      v.w = k0 + 3 < d ? to_tf32(sm[(size_t) r * pitch + k0 + 3]) : 0.f;
      *reinterpret_cast<float4*>(img + (size_t) q * G_CHUNK_FLOATS + r * GK + pc * 4) = v;
    }
  }
}

// lb[(s - s0)][t] = 0.999 x a lower bound of the squared distance between any frame of tile s and any frame of tile t:
// max(box-to-box gap, (centre distance - radius_s - radius_t)^2); rounding slack is added by the caller's thresholds.
// Block: 128 column tiles x LB_ROWS row tiles.
constexpr int LB_ROWS = 8;
__global__ void tile_lb_kernel(const float* __restrict__ tcen, const float* __restrict__ tlo, const float* __restrict__ thi,
                               const float* __restrict__ trad, int d, size_t n_tiles, uint32_t s0, uint32_t s1, float* __restrict__ lb) {
  extern __shared__ float sh[];                      // [LB_ROWS][3][d]
  const uint32_t sb = s0 + blockIdx.y * LB_ROWS;
  const size_t t = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
  for (int q = threadIdx.x; q < LB_ROWS * d; q += blockDim.x) {
    const int r = q / d, k = q % d;
    const uint32_t s = min(sb + (uint32_t) r, s1 - 1);
    sh[(r * 3 + 0) * d + k] = tcen[(size_t) k * n_tiles + s];
    sh[(r * 3 + 1) * d + k] = tlo[(size_t) k * n_tiles + s];
    sh[(r * 3 + 2) * d + k] = thi[(size_t) k * n_tiles + s];
  }
  __syncthreads();
  if (t >= n_tiles) return;
  float gap[LB_ROWS], cd[LB_ROWS];
#pragma unroll
  for (int r = 0; r < LB_ROWS; ++r) gap[r] = cd[r] = 0.f;
  for (int k = 0; k < d; ++k) {
    const float c = tcen[(size_t) k * n_tiles + t], lo = tlo[(size_t) k * n_tiles + t], hi = thi[(size_t) k * n_tiles + t];
#pragma unroll
    for (int r = 0; r < LB_ROWS; ++r) {
      const float g = fmaxf(fmaxf(sh[(r * 3 + 1) * d + k] - hi, lo - sh[(r * 3 + 2) * d + k]), 0.f);
      gap[r] = fmaf(g, g, gap[r]);
      const float e = sh[(r * 3 + 0) * d + k] - c;
      cd[r] = fmaf(e, e, cd[r]);
    }
  }
  const float rt = trad[t];
#pragma unroll
  for (int r = 0; r < LB_ROWS; ++r) {
    const uint32_t s = sb + (uint32_t) r;
    if (s >= s1) break;
    const float sph = fmaxf(sqrtf(cd[r]) * 0.9999f - trad[s] - rt, 0.f);
    float v = fmaxf(gap[r], sph * sph) * 0.999f;
    if (!(v == v)) v = 0.f;                        // empty tiles (inf - inf): never pruned, they hold nothing anyway
    lb[(size_t) (s - s0) * n_tiles + t] = v;
  }
}

// per-tile minimum of the float free-energy ranks (neighbour search: "the tile holds no frame of lower free energy")
__global__ void tile_min_kernel(const float* __restrict__ lof, size_t n_tiles, float* __restrict__ lomin) {
  const size_t tile = (size_t) blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (tile >= n_tiles) return;
  const float4 q = *reinterpret_cast<const float4*>(lof + tile * GT + 4 * lane);
  float m = fminf(fminf(q.x, q.y), fminf(q.z, q.w));
  for (int o = 16; o > 0; o >>= 1) m = fminf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if (lane == 0) lomin[tile] = m;
}

// ------------------------------------------------------------------------------------------------
// scan geometry and shared memory
// ------------------------------------------------------------------------------------------------

struct GSmem {
  float* a;                   // ra x kc chunks
  float* ring;                // n_stages chunks
  float* side;                // G_SIDE_SLOTS records
  float* scratch;             // G_HALF x G_EPI_THREADS
  uint64_t *full, *empty;     // ring: one full / empty barrier per group of cb slots
  uint64_t *side_full, *side_empty;
  uint64_t *tmem_full, *tmem_empty;      // G_ACC each
  uint64_t *a_full, *a_empty;
  uint32_t* tmem_addr;
  unsigned long long* wthr;   // [16 epilogue warps][2]: (item << 32 | float bits) bounds published for the producer (nn, hd)
  __host__ __device__ static size_t bytes(int akc, int n_stages) {        // akc = ra * kc
    return 1024 + (size_t) (akc + n_stages) * G_CHUNK_BYTES + (size_t) G_SIDE_SLOTS * G_SIDE_FLOATS * 4 +
           (size_t) G_EPI_THREADS * G_HALF * 4 + (size_t) (2 * 8 + 2 * G_SIDE_SLOTS + 2 * G_ACC + 2) * 8 + 16 + 2 * G_EPI_WARPS * 8;
  }
  __device__ GSmem(unsigned char* raw, int kc, int n_stages) {            // kc here = ra * kc
    const uint32_t a0 = smem_u32(raw);
    unsigned char* base = raw + ((1024u - (a0 & 1023u)) & 1023u);       // SWIZZLE_128B atoms need 1024-byte alignment
    a = reinterpret_cast<float*>(base);
    ring = a + (size_t) kc * G_CHUNK_FLOATS;
    side = ring + (size_t) n_stages * G_CHUNK_FLOATS;
    scratch = side + G_SIDE_SLOTS * G_SIDE_FLOATS;
    full = reinterpret_cast<uint64_t*>(scratch + G_EPI_THREADS * G_HALF);
    empty = full + 8;
    side_full = empty + 8;
    side_empty = side_full + G_SIDE_SLOTS;
    tmem_full = side_empty + G_SIDE_SLOTS;
    tmem_empty = tmem_full + G_ACC;
    a_full = tmem_empty + G_ACC;
    a_empty = a_full + 1;
    wthr = reinterpret_cast<unsigned long long*>(a_empty + 1);
    tmem_addr = reinterpret_cast<uint32_t*>(wthr + 2 * G_EPI_WARPS);
  }
};

// diagnostics: time a wait when the library is built with -DDCB_GEMM_PROF (the counters cost issue slots in the
// single-warp producer / MMA loops, so they are compiled out by default)
#ifdef DCB_GEMM_PROF
#define G_TIMED(slot, stmt)                                   \
  do {                                                        \
    if (g.prof) {                                             \
      const long long t0__ = clock64();                       \
      stmt;                                                   \
      pacc[slot] += (unsigned long long) (clock64() - t0__);  \
    } else {                                                  \
      stmt;                                                   \
    }                                                         \
  } while (0)
#else
#define G_TIMED(slot, stmt) \
  do {                      \
    (void) pacc;            \
    stmt;                   \
  } while (0)
#endif

// side record meta
struct GMeta {
  uint32_t row_tile;       // row block of the item (ra row tiles), relative to row_begin
  uint32_t col0;           // first column (position) of the tile
  uint32_t flags;          // 0: tile, 1: end of item, 2: end of stream
  uint32_t item;
};
constexpr uint32_t G_END = 1u, G_EXIT = 2u;

// work item -> (row tile, range of column tiles): column-step-major like item_coords (kernels.cuh), own neighbourhood first
__device__ __forceinline__ void g_item_coords(const GemmGeom& g, uint32_t item, uint32_t* rb, uint32_t* ci) {
  const uint32_t step = item / g.n_row_tiles;
  *rb = item % g.n_row_tiles;
  const uint32_t diag = min((g.row_begin / GT + *rb * (uint32_t) g.ra) / g.tiles_per_item, g.n_col_items - 1);
  const uint32_t k = (step + 1) >> 1;
  const uint32_t n = g.n_col_items;
  *ci = (step & 1) ? (diag + k) % n : (diag + n - (k % n)) % n;
}

// Producer warp.  keep(rb, t, lbv, tile_extra[t], item, lane): warp-uniform decision taken right before a tile is streamed (dynamic bounds);
// thr0(rb): the bound an item starts with; range(rb, lim0, lim1): restricts the column tiles (window pass).
template <class Thr0, class Keep, class Range>
__device__ __forceinline__ void g_produce(const GemmGeom& g, GSmem& S, const float* __restrict__ side_extra,
                                          const float* __restrict__ tile_extra, Thr0&& thr0, Keep&& keep, Range&& range) {
  const int lane = threadIdx.x & 31;
  const uint32_t total = g.n_row_tiles * g.n_col_items;
  uint32_t stage = 0, phase = 0;          // chunk ring
  uint32_t sseq = 0;                      // side records issued
  uint32_t a_uses = 0;                    // items that loaded a row tile
  unsigned long long streamed = 0;
  unsigned long long pacc[4] = {0, 0, 0, 0};
  const long long t_begin = clock64();
  const uint32_t side_bytes = (side_extra ? 2u : 1u) * GT * 4u;
  auto side_slot = [&](uint32_t flags, uint32_t rb, uint32_t col0, uint32_t item, uint32_t tile, bool data) {
    // lane 0 only
    const uint32_t s = sseq % G_SIDE_SLOTS;
    G_TIMED(0, mbar_wait(&S.side_empty[s], ((sseq / G_SIDE_SLOTS) & 1u) ^ 1u));
    float* rec = S.side + s * G_SIDE_FLOATS;
    GMeta m;
    m.row_tile = rb; m.col0 = col0; m.flags = flags; m.item = item;
    *reinterpret_cast<GMeta*>(rec) = m;
    if (data) {
      mbar_arrive_expect_tx(&S.side_full[s], side_bytes);
      tma_load_1d(rec + 4, g.gnorm + (size_t) tile * GT, GT * 4, &S.side_full[s]);
      if (side_extra) tma_load_1d(rec + 4 + GT, side_extra + (size_t) tile * GT, GT * 4, &S.side_full[s]);
    } else {
      mbar_arrive(&S.side_full[s]);
    }
    ++sseq;
  };
  for (;;) {
    uint32_t item = 0;
    if (lane == 0) item = atomicAdd(g.work_counter, 1u);
    item = __shfl_sync(0xffffffffu, item, 0);
    if (item >= total) break;
    uint32_t rb, ci;
    g_item_coords(g, item, &rb, &ci);
    uint32_t t0 = ci * g.tiles_per_item;
    uint32_t t1 = min(t0 + g.tiles_per_item, g.n_tiles);
    uint32_t lim0 = 0, lim1 = g.n_tiles;
    range(rb, lim0, lim1);
    t0 = max(t0, lim0);
    t1 = min(t1, lim1);
    if (t0 >= t1) continue;
    const float th0 = thr0(rb);
    const uint32_t rt0 = rb * (uint32_t) g.ra;                       // first 128-row tile of the block
    const float* __restrict__ lbrow = g.lb + (size_t) rt0 * g.n_tiles;
    const bool two = g.ra == 2 && rt0 + 1 < g.n_row_tiles128;        // the block's second row tile holds rows
    bool first = true;
    for (uint32_t base = t0; base < t1; base += 32) {
      const uint32_t t = base + lane;
      float lbv = t < t1 ? __ldg(lbrow + t) : INFINITY;
      if (two && t < t1) lbv = fminf(lbv, __ldg(lbrow + g.n_tiles + t));
      // per-tile value the keep rule wants (neighbour search: the tile's smallest free-energy rank), fetched 32 tiles at a
      // time here instead of one dependent global load per streamed tile on the producer's critical path
      const float exv = (tile_extra && t < t1) ? __ldg(tile_extra + t) : 0.f;
      uint32_t mask = __ballot_sync(0xffffffffu, t < t1 && !(lbv > th0));
      while (mask) {
        const int src = __ffs(mask) - 1;
        mask &= mask - 1;
        const uint32_t tt = base + (uint32_t) src;
        if (!keep(rb, tt, __shfl_sync(0xffffffffu, lbv, src), __shfl_sync(0xffffffffu, exv, src), item, lane)) continue;
        if (lane == 0) {
          if (first) {
            // the row tiles' operand images (consecutive tiles = one contiguous block): resident for the whole item
            G_TIMED(2, mbar_wait(&S.a_empty[0], (a_uses & 1u) ^ 1u));
            mbar_arrive_expect_tx(&S.a_full[0], (uint32_t) (g.ra * g.kc) * G_CHUNK_BYTES);
            const float* src_a = g.gT + (size_t) (g.row_begin / GT + rt0) * g.kc * G_CHUNK_FLOATS;
            for (int q = 0; q < g.ra * g.kc; ++q)
              tma_load_1d(S.a + (size_t) q * G_CHUNK_FLOATS, src_a + (size_t) q * G_CHUNK_FLOATS, G_CHUNK_BYTES, &S.a_full[0]);
          }
          side_slot(0u, rb, tt * GT, item, tt, true);
          const float* src_b = g.gT + (size_t) tt * g.kc * G_CHUNK_FLOATS;
          for (int q = 0; q < g.kc; ++q) {
            // ring slots are filled and freed in groups of cb chunks (cb divides kc: a group never straddles two tiles)
            const uint32_t grp = stage >> g.cb_log2;
            if ((stage & (uint32_t) (g.cb - 1)) == 0) {
              G_TIMED(1, mbar_wait(&S.empty[grp], phase ^ 1u));
              mbar_arrive_expect_tx(&S.full[grp], (uint32_t) g.cb * G_CHUNK_BYTES);
            }
            tma_load_1d(S.ring + (size_t) stage * G_CHUNK_FLOATS, src_b + (size_t) q * G_CHUNK_FLOATS, G_CHUNK_BYTES, &S.full[grp]);
            if (++stage == (uint32_t) g.n_stages) { stage = 0; phase ^= 1u; }
          }
        }
        if (first) ++a_uses;
        first = false;
        ++streamed;
        __syncwarp();
      }
    }
    if (!first && lane == 0) {
      side_slot(G_END, rb, 0, item, 0, false);     // end of the item
    }
  }
  if (lane == 0) {
    side_slot(G_EXIT, 0, 0, 0xffffffffu, 0, false);
    if (g.stats && streamed) atomicAdd(g.stats + 2, streamed);
#ifdef DCB_GEMM_PROF
    if (g.prof) {
      pacc[3] = (unsigned long long) (clock64() - t_begin);
      for (int q = 0; q < 4; ++q) atomicAdd(g.prof + q, pacc[q]);
    }
#else
    (void) t_begin;
#endif
  }
}

// MMA warp: follows the side records; per column tile kc chunks x RA row tiles x (up to 4) K=8 steps into accumulator
// stage p = tile count % (G_ACC / RA), TMEM columns (p RA + r) 128.  The whole warp runs the loop with warp-uniform
// values (descriptors live in uniform registers) and one elected lane issues.  The per-chunk code is kept minimal: one
// warp executes it serially, so every extra instruction between two tcgen05.mma is exposed latency (the first version
// spent ~300 SASS instructions per chunk and reached 1/3 of the MMA rate).  Ring slots are handed back cb at a time and
// the accumulators are published once per column tile (tcgen05.commit is not free either).
template <int RA>
__device__ __forceinline__ void g_mma(const GemmGeom& g, GSmem& S, uint32_t tmem_base, uint32_t r_mine) {
  uint32_t stage = 0, phase = 0, seq = 0, a_uses = 0, tiles = 0;
  uint32_t acc_uses = 0;                    // bit field: parity of the uses of each accumulator
  bool need_a = true;
  unsigned long long pacc[5] = {0, 0, 0, 0, 0};
  const long long t_begin = clock64();
  const uint32_t cb_mask = (uint32_t) g.cb - 1u, cb_log2 = (uint32_t) g.cb_log2;
  constexpr uint32_t p_mask = (uint32_t) (G_ACC / RA) - 1u;      // 4 or 2 accumulator stages
  const int kc = g.kc;
  const int kc_full = g.k8 >> 2;                                 // chunks with all four K steps
  const int tail_steps = g.k8 & 3;                               // K steps of the last chunk if it is partial
  // this warp's row tile: its operand image follows the first one in shared memory
  const uint64_t a_desc0 = g_smem_desc(smem_u32(S.a)) + (uint64_t) r_mine * (uint64_t) kc * (G_CHUNK_BYTES >> 4);
  const uint64_t b_desc0 = g_smem_desc(smem_u32(S.ring));
  const uint32_t n_stages = (uint32_t) g.n_stages;
  const bool leader = elect_one();
  for (;;) {
    const uint32_t s = seq % G_SIDE_SLOTS;
    G_TIMED(0, mbar_wait(&S.side_full[s], (seq / G_SIDE_SLOTS) & 1u));
    const uint32_t flags = reinterpret_cast<const GMeta*>(S.side + s * G_SIDE_FLOATS)->flags;
    if (flags & G_EXIT) break;
    if (flags & G_END) {
      if (leader) tc_commit(&S.a_empty[0]);          // the row tiles may be replaced once every MMA of the item has read them
      need_a = true;
    } else {
      if (need_a) {
        G_TIMED(1, mbar_wait(&S.a_full[0], a_uses & 1u));
        ++a_uses;
        need_a = false;
      }
      const uint32_t acc = (tiles & p_mask) * RA + r_mine;       // accumulator = TMEM columns acc * 128 ..
      ++tiles;
      G_TIMED(2, mbar_wait(&S.tmem_empty[acc], ((acc_uses >> acc) & 1u) ^ 1u));
      acc_uses ^= 1u << acc;
      const uint32_t dt = tmem_base + acc * (uint32_t) GT;
      uint64_t ad = a_desc0;
      for (int q = 0; q < kc; ++q, ad += (G_CHUNK_BYTES >> 4)) {
        if ((stage & cb_mask) == 0) {
          G_TIMED(3, mbar_wait(&S.full[stage >> cb_log2], phase));
          tc_fence_after();
        }
        // chunk q of the row tile against ring slot `stage`: +1024 per 16 KB chunk and +2 per K step in the address field
        const uint64_t bd = b_desc0 + (uint64_t) stage * (G_CHUNK_BYTES >> 4);
        if (leader) {
          tc_mma_tf32(dt, ad, bd, G_IDESC, q ? 1u : 0u);
          if (q < kc_full) {
            tc_mma_tf32(dt, ad + 2, bd + 2, G_IDESC, 1u);
            tc_mma_tf32(dt, ad + 4, bd + 4, G_IDESC, 1u);
            tc_mma_tf32(dt, ad + 6, bd + 6, G_IDESC, 1u);
          } else {
            if (tail_steps > 1) tc_mma_tf32(dt, ad + 2, bd + 2, G_IDESC, 1u);
            if (tail_steps > 2) tc_mma_tf32(dt, ad + 4, bd + 4, G_IDESC, 1u);
          }
          if (((stage + 1) & cb_mask) == 0) tc_commit(&S.empty[stage >> cb_log2]);
        }
        if (++stage == n_stages) { stage = 0; phase ^= 1u; }
      }
      if (leader) tc_commit(&S.tmem_full[acc]);
    }
    if (leader) mbar_arrive(&S.side_empty[s]);
    ++seq;
  }
#ifdef DCB_GEMM_PROF
  if (g.prof && leader && r_mine == 0) {
    pacc[4] = (unsigned long long) (clock64() - t_begin);
    for (int q = 0; q < 5; ++q) atomicAdd(g.prof + 4 + q, pacc[q]);
  }
#else
  (void) t_begin;
#endif
}

__device__ __forceinline__ void g_init(const GemmGeom& g, GSmem& S) {
  if (threadIdx.x == 0) {
    const uint32_t n_mma = (uint32_t) g.ra;          // MMA-issuing warps: one per row tile of the block
    for (int s = 0; s < 8; ++s) {
      mbar_init(&S.full[s], 1);
      mbar_init(&S.empty[s], n_mma);
    }
    for (int s = 0; s < G_SIDE_SLOTS; ++s) {
      mbar_init(&S.side_full[s], 1);
      mbar_init(&S.side_empty[s], n_mma + G_EPI_WARPS);      // MMA warps + the sixteen epilogue warps
    }
    for (int s = 0; s < G_ACC; ++s) {
      mbar_init(&S.tmem_full[s], 1);
      mbar_init(&S.tmem_empty[s], G_EPI_WARPS / g.ra);       // the epilogue warps that read this accumulator
    }
    mbar_init(&S.a_full[0], 1);
    mbar_init(&S.a_empty[0], n_mma);
    for (int q = 0; q < 2 * G_EPI_WARPS; ++q) S.wthr[q] = ~0ull;
    fence_mbar_init();
  }
}

// Epilogue of one warpgroup's share of an accumulator stage: n_loads (2 or 4) 16-column loads starting at column c_lo of
// accumulator `acc_col`, each in flight while the previous chunk is processed; the stage is handed back to the MMA warp
// as soon as the last load has landed in registers.  proc(regs, first column within the tile).
template <class Proc>
__device__ __forceinline__ void g_epilogue_tile(GSmem& S, uint32_t tmem_base, uint32_t quarter, uint32_t acc_col, uint32_t acc, int c_lo,
                                                int n_loads, int lane, Proc&& proc) {
  uint32_t ra[G_LD], rb[G_LD];
  const uint32_t t0 = tmem_base + ((quarter * 32u) << 16) + acc_col + (uint32_t) c_lo;
  tmem_ld16_issue(t0, ra);
#pragma unroll 1
  for (int it = 0; 2 * it < n_loads; ++it) {
    tmem_ld16_wait(ra);
    tmem_ld16_issue(t0 + (uint32_t) (it * 2 * G_LD + G_LD), rb);
    proc(ra, c_lo + it * 2 * G_LD);
    tmem_ld16_wait(rb);
    if (2 * it + 2 < n_loads) {
      tmem_ld16_issue(t0 + (uint32_t) (it * 2 * G_LD + 2 * G_LD), ra);
    } else {
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&S.tmem_empty[acc]);
    }
    proc(rb, c_lo + it * 2 * G_LD + G_LD);
  }
}

// Variant that hands the accumulator back first: all four loads of the warpgroup's 64 columns are issued at once, the
// stage is released as soon as they have landed (~ one TMEM latency after the MMAs finished, instead of after three of
// the four chunks have been processed), and the math runs from registers while the next MMAs already refill the stage.
// Costs 64 live registers: used where the per-pair state is small (one or two radii, neighbour search).
template <class Proc>
__device__ __forceinline__ void g_epilogue_tile_early(GSmem& S, uint32_t tmem_base, uint32_t quarter, uint32_t acc_col, uint32_t acc, int c_lo,
                                                      int lane, Proc&& proc) {
  uint32_t r0[G_LD], r1[G_LD], r2[G_LD], r3[G_LD];
  const uint32_t t0 = tmem_base + ((quarter * 32u) << 16) + acc_col + (uint32_t) c_lo;
  tmem_ld16_issue(t0, r0);
  tmem_ld16_issue(t0 + G_LD, r1);
  tmem_ld16_issue(t0 + 2 * G_LD, r2);
  tmem_ld16_issue(t0 + 3 * G_LD, r3);
  tmem_ld16_wait(r0);                       // tcgen05.wait::ld covers every outstanding load; the other buffers are tied below
  tmem_ld16_wait(r1);
  tmem_ld16_wait(r2);
  tmem_ld16_wait(r3);
  tc_fence_before();
  __syncwarp();
  if (lane == 0) mbar_arrive(&S.tmem_empty[acc]);
  proc(r0, c_lo);
  proc(r1, c_lo + G_LD);
  proc(r2, c_lo + 2 * G_LD);
  proc(r3, c_lo + 3 * G_LD);
}

// band half width around a decision value at distance r (d2 = r^2):  2 r rho + rho^2 + eps + e_rel r^2, rounded up
__device__ __forceinline__ float g_band(float r, float r2, float rho, float eps, float e_rel) {
  return fmaf(e_rel, r2, fmaf(2.0f * r, rho, fmaf(rho, rho, eps))) * 1.0001f;
}

// ================================================================================================
// populations (count mode, up to 8 radii per pass)
// ================================================================================================

template <int NB, bool CHECK>
__global__ void __launch_bounds__(G_THREADS, 1) gscan_pops_kernel(const __grid_constant__ GPopsArgs a) {
  extern __shared__ unsigned char g_smem_raw[];
  const GemmGeom& g = a.g;
  GSmem S(g_smem_raw, g.ra * g.kc, g.n_stages);
  const int warp = uniform_warp(), lane = threadIdx.x & 31;
  g_init(g, S);
  __syncthreads();
  if (warp == 1) tmem_alloc(S.tmem_addr, G_ACC * GT);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(S.tmem_addr);

  if (warp == 0) {
    g_produce(g, S, nullptr, nullptr, [&](uint32_t) { return g.prune_thr; }, [](uint32_t, uint32_t, float, float, uint32_t, int) { return true; },
              [](uint32_t, uint32_t&, uint32_t&) {});
  } else if (warp == 1) {
    if (g.ra == 2) g_mma<2>(g, S, tmem_base, 0u); else g_mma<1>(g, S, tmem_base, 0u);
  } else if (warp == G_MMA2_WARP) {
    if (g.ra == 2) g_mma<2>(g, S, tmem_base, 1u);       // the block's second row tile has its own issuing warp
  } else {
    // Four epilogue warpgroups, thread = one accumulator row.  ra == 2: warpgroup wg serves row tile (wg & 1) of the block,
    // columns 64 (wg >> 1) .. + 63 of its accumulator; ra == 1: all four hold the same 128 rows, 32 columns each.
    const uint32_t wg = (uint32_t) (warp - 2) >> 2;
    const uint32_t my_tile = g.ra == 2 ? (wg & 1u) : 0u;            // row tile of the block this thread's row belongs to
    const uint32_t quarter = (uint32_t) warp & 3u;                  // TMEM lanes 32 quarter .. 32 quarter + 31
    const uint32_t row_in_tile = quarter * 32u + (uint32_t) lane;
    const int et = (warp - 2) * 32 + lane;
    float* scratch = S.scratch + et;
    uint32_t seq = 0, tiles = 0, uses = 0, cur_item = 0xfffffffeu, row = 0;
    bool valid = false;
    float q[NB], E[NB], xn = INFINITY;
    uint32_t cnt[NB];
    uint32_t n_slow = 0, n_exact = 0, n_tiles = 0;
    unsigned long long pacc[3] = {0, 0, 0};
    const long long t_begin = clock64();
#pragma unroll
    for (int b = 0; b < NB; ++b) { cnt[b] = 0; q[b] = INFINITY; E[b] = 0.f; }
    for (;;) {
      const uint32_t s = seq % G_SIDE_SLOTS;
      G_TIMED(0, mbar_wait(&S.side_full[s], (seq / G_SIDE_SLOTS) & 1u));
      const float* rec = S.side + s * G_SIDE_FLOATS;
      const GMeta m = *reinterpret_cast<const GMeta*>(rec);
      if (m.flags & G_EXIT) break;
      if (m.flags & G_END) {
        if (cur_item == m.item && valid) {
#pragma unroll
          for (int b = 0; b < NB; ++b)
            if (cnt[b]) atomicAdd(a.cnt + (size_t) b * a.ld_cnt + (row - g.row_begin), cnt[b]);
        }
        cur_item = 0xfffffffeu;
      } else {
        if (cur_item != m.item) {
          cur_item = m.item;
          row = g.row_begin + (m.row_tile * (uint32_t) g.ra + my_tile) * GT + row_in_tile;
          valid = row < g.row_end;
          xn = valid ? __ldg(g.gnorm + row) : INFINITY;
          const float rho = g.rho_c * (sqrtf(xn) + sqrtf(g.nymax));
          const float eps = g.c_acc * (xn + g.nymax);
#pragma unroll
          for (int b = 0; b < NB; ++b) {
            cnt[b] = 0;
            q[b] = xn - a.rad2[b];
            // the sign of F - r^2 decides outside the band; monotonicity of d^2 -+ band(d) needs r > rho
            E[b] = (a.rad[b] > 1.01f * rho) ? g_band(a.rad[b], a.rad2[b], rho, eps, g.e_rel) + 4.8e-7f * fabsf(q[b]) : INFINITY;
            if (a.rad2[b] < 0.f) E[b] = -1.f;                         // unused slot
          }
        }
        const uint32_t acc = (tiles & ((uint32_t) (G_ACC / g.ra) - 1u)) * (uint32_t) g.ra + my_tile;
        ++tiles;
        G_TIMED(1, mbar_wait(&S.tmem_full[acc], (uses >> acc) & 1u));
        uses ^= 1u << acc;
        tc_fence_after();
        if (wg < (uint32_t) g.ra) ++n_tiles;                 // 128 x 128 units: one count per row tile of the record
        const float* ny = rec + 4;
        auto proc = [&](const uint32_t (&v)[G_LD], int c0) {
#pragma unroll
          for (int h = 0; h < G_LD; h += G_HALF) {
            float mn[NB];
#pragma unroll
            for (int b = 0; b < NB; ++b) mn[b] = INFINITY;
#pragma unroll
            for (int c4 = 0; c4 < G_HALF; c4 += 4) {
              const float4 n4 = *reinterpret_cast<const float4*>(ny + c0 + h + c4);
              const float nn[4] = {n4.x, n4.y, n4.z, n4.w};
#pragma unroll
              for (int c = 0; c < 4; ++c) {
                const float w = fmaf(__uint_as_float(v[h + c4 + c]), -2.0f, nn[c]);
#pragma unroll
                for (int b = 0; b < NB; ++b) {
                  const float vv = w + q[b];
                  cnt[b] += __float_as_uint(vv) >> 31;
                  mn[b] = fminf(mn[b], fabsf(vv));
                }
              }
            }
            bool band = false;
#pragma unroll
            for (int b = 0; b < NB; ++b) band |= mn[b] <= E[b];
            if (CHECK || __any_sync(0xffffffffu, band)) {
              // rare: some lane has a pair of these 8 columns within the band of a radius; its sign decision is replaced by the
              // exact one.  Warp-uniform region: the exact distance of a wanted pair is computed by the whole warp.
#pragma unroll
              for (int c = 0; c < G_HALF; ++c) scratch[c * G_EPI_THREADS] = __uint_as_float(v[h + c]);
#pragma unroll 1
              for (int c = 0; c < G_HALF; ++c) {
                const float w = fmaf(scratch[c * G_EPI_THREADS], -2.0f, ny[c0 + h + c]);
                const uint32_t col = m.col0 + (uint32_t) (c0 + h + c);          // the same column for every lane
                const bool chk = CHECK && valid && col < g.n && g.check;
                bool want = chk;
#pragma unroll
                for (int b = 0; b < NB; ++b) want |= fabsf(w + q[b]) <= E[b] && valid;
                uint32_t mk = __ballot_sync(0xffffffffu, want);
                float d2 = 0.f;
                while (mk) {
                  const int src = __ffs(mk) - 1;
                  mk &= mk - 1;
                  const float t = dist2_exact_coop(g.xR, g.d, __shfl_sync(0xffffffffu, row, src), col, lane);
                  if (lane == src) d2 = t;
                }
                if (want) {
                  ++n_exact;
                  if (chk) {
                    const float rho = g.rho_c * (sqrtf(xn) + sqrtf(g.nymax));
                    const float bnd = g_band(sqrtf(d2), d2, rho, g.c_acc * (xn + g.nymax), g.e_rel);
                    const float ratio = fabsf((w + xn) - d2) / bnd;
                    atomicMax(reinterpret_cast<unsigned int*>(g.check), __float_as_uint(ratio));
                  }
#pragma unroll
                  for (int b = 0; b < NB; ++b) {
                    const float vv = w + q[b];
                    if (fabsf(vv) <= E[b] && valid) {
                      ++n_slow;
                      const uint32_t inside = d2 < a.rad2[b] ? 1u : 0u;      // NaN (padding) -> outside
                      cnt[b] += inside - (__float_as_uint(vv) >> 31);
                    }
                  }
                }
              }
            }
          }
        };
        if (g.ra == 2) {
          if (NB <= 2) g_epilogue_tile_early(S, tmem_base, quarter, acc * (uint32_t) GT, acc, (int) (wg >> 1) * 64, lane, proc);
          else g_epilogue_tile(S, tmem_base, quarter, acc * (uint32_t) GT, acc, (int) (wg >> 1) * 64, 4, lane, proc);
        } else {
          g_epilogue_tile(S, tmem_base, quarter, acc * (uint32_t) GT, acc, (int) wg * 32, 2, lane, proc);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&S.side_empty[s]);
      ++seq;
    }
#ifdef DCB_GEMM_PROF
    if (g.prof && lane == 0 && quarter == 0) {
      pacc[2] = (unsigned long long) (clock64() - t_begin);
      for (int q = 0; q < 3; ++q) atomicAdd(g.prof + 9 + q, pacc[q]);
    }
#else
    (void) t_begin;
#endif
    // statistics
    for (int o = 16; o > 0; o >>= 1) {
      n_slow += __shfl_xor_sync(0xffffffffu, n_slow, o);
      n_exact += __shfl_xor_sync(0xffffffffu, n_exact, o);
    }
    if (lane == 0 && g.stats) {
      if (n_slow) atomicAdd(g.stats, (unsigned long long) n_slow);
      if (n_exact) atomicAdd(g.stats + 1, (unsigned long long) n_exact);
      if (n_tiles && quarter == 0) atomicAdd(g.stats + 3, (unsigned long long) n_tiles);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, G_ACC * GT);
}

// ================================================================================================
// nearest neighbours
// ================================================================================================

__device__ __forceinline__ float g_key_d2(unsigned long long k) { return __uint_as_float((uint32_t) (k >> 32)); }

__global__ void __launch_bounds__(G_THREADS, 1) gscan_nn_kernel(const __grid_constant__ GNnArgs a) {
  extern __shared__ unsigned char g_smem_raw[];
  const GemmGeom& g = a.g;
  GSmem S(g_smem_raw, g.ra * g.kc, g.n_stages);
  const int warp = uniform_warp(), lane = threadIdx.x & 31;
  g_init(g, S);
  __syncthreads();
  if (warp == 1) tmem_alloc(S.tmem_addr, G_ACC * GT);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(S.tmem_addr);

  if (warp == 0) {
    // A tile is needed if it can hold a nearest neighbour of some row (lb <= thr_nn), or a lower-free-energy neighbour
    // (lb <= thr_hd) and it holds a frame of lower rank than the rows' largest one at all.
    float bound_nn = 0.f, bound_hd = 0.f, lmax = 0.f;
    g_produce(g, S, a.lof, a.lomin,
              [&](uint32_t rb) {
                // max over the 4 ra quarters of the row block
                const int ql = lane % (4 * g.ra);
                float bn = *reinterpret_cast<volatile float*>(a.thr_nn + (size_t) rb * 4 * g.ra + ql);
                float bh = *reinterpret_cast<volatile float*>(a.thr_hd + (size_t) rb * 4 * g.ra + ql);
                float lm = __ldg(a.lormax + (size_t) rb * 4 * g.ra + ql);
                if (!(bn < INFINITY)) bn = INFINITY;
                if (!(bh < INFINITY)) bh = INFINITY;
                bound_nn = __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(fmaxf(bn, 0.f))));
                bound_hd = __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(fmaxf(bh, 0.f))));
                lmax = __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(fmaxf(lm, 0.f))));
                return fmaxf(bound_nn, bound_hd);
              },
              [&](uint32_t, uint32_t, float lbv, float lomin_t, uint32_t item, int ln) {
                // What the epilogue warps found since the item started.  Lane l reads slot l: epilogue warp l >> 1 (warpgroup
                // l >> 3), l & 1: nn / hd.  A warpgroup's bound is the max over its four warps (a warp that has not published
                // for this item yet: +inf).  Warpgroups that hold the same rows both give valid bounds (the tighter counts);
                // the block needs the max over its row tiles.
                uint32_t bits = 0x7f800000u;
                {
                  const unsigned long long v = *reinterpret_cast<volatile unsigned long long*>(S.wthr + ln);
                  if ((uint32_t) (v >> 32) == item) bits = (uint32_t) v;
                }
                bits = max(bits, __shfl_xor_sync(0xffffffffu, bits, 2));
                bits = max(bits, __shfl_xor_sync(0xffffffffu, bits, 4));
                const int kind = ln & 1;
                const float w0 = __uint_as_float(__shfl_sync(0xffffffffu, bits, kind)), w1 = __uint_as_float(__shfl_sync(0xffffffffu, bits, 8 + kind));
                const float w2 = __uint_as_float(__shfl_sync(0xffffffffu, bits, 16 + kind)), w3 = __uint_as_float(__shfl_sync(0xffffffffu, bits, 24 + kind));
                // ra == 2: warpgroups 0, 2 hold row tile 0 and 1, 3 row tile 1; ra == 1: all four hold the same rows
                const float mine = g.ra == 2 ? fmaxf(fminf(w0, w2), fminf(w1, w3)) : fminf(fminf(w0, w1), fminf(w2, w3));
                const float bn = fminf(bound_nn, __shfl_sync(0xffffffffu, mine, 0));
                const float bh = fminf(bound_hd, __shfl_sync(0xffffffffu, mine, 1));
                if (!(lbv > bn)) return true;
                return !(lbv > bh) && lomin_t < lmax;
              },
              [&](uint32_t rb, uint32_t& lim0, uint32_t& lim1) {
                if (a.window) {
                  const uint32_t t = g.row_begin / GT + rb * (uint32_t) g.ra;
                  lim0 = t > a.window ? t - a.window : 0u;
                  lim1 = min(lim1, t + (uint32_t) g.ra + a.window);
                }
              });
  } else if (warp == 1) {
    if (g.ra == 2) g_mma<2>(g, S, tmem_base, 0u); else g_mma<1>(g, S, tmem_base, 0u);
  } else if (warp == G_MMA2_WARP) {
    if (g.ra == 2) g_mma<2>(g, S, tmem_base, 1u);       // the block's second row tile has its own issuing warp
  } else {
    const uint32_t wg = (uint32_t) (warp - 2) >> 2;                 // see gscan_pops_kernel
    const uint32_t my_tile = g.ra == 2 ? (wg & 1u) : 0u;
    const uint32_t quarter = (uint32_t) warp & 3u;
    const uint32_t row_in_tile = quarter * 32u + (uint32_t) lane;
    const int et = (warp - 2) * 32 + lane;
    float* scratch = S.scratch + et;
    uint32_t seq = 0, tiles = 0, uses = 0, cur_item = 0xfffffffeu, row = 0, lo_i = 0;
    bool valid = false;
    float xn = INFINITY, rho = 0.f, eps = 0.f, lor = 0.f;
    float t_nn = -INFINITY, t_hd = -INFINITY, dl = 0.f;
    unsigned long long best_nn = ~0ull, best_hd = ~0ull;
    uint32_t n_slow = 0, n_exact = 0, n_tiles = 0;
    // every column whose exact d2 is <= b satisfies  fma(acc, -2, |y~|^2) < thr(b)
    auto thr = [&](float b) {
      if (!(b < FLT_MAX)) return INFINITY;
      const float b1 = fmaf(1.01f * g.e_rel, b, b);
      const float T = b1 + g_band(sqrtf(b1) * 1.000001f, b1, rho, eps, g.e_rel);
      return next_up(next_up(fmaf(T, 1.000001f, -xn) + 2.4e-7f * (T + xn)));
    };
    auto set_dl = [&]() {
      float v = 0.f;
      if (t_nn < INFINITY) v = t_hd < INFINITY ? fminf(next_up((t_hd - t_nn) * 1.000001f), 1e37f) : 1e37f;
      dl = fmaxf(v, 0.f);
    };
    // bound (d2 units, pruning margins included) of what this row still accepts
    auto row_bound = [&](unsigned long long key) {
      const float b = g_key_d2(key);
      float v = (fmaf(g.e_rel, b, b) + g.prune_slack) * 1.00001f;
      if (!(v < INFINITY)) v = INFINITY;
      return v;
    };
    auto publish = [&]() {
      float vn = valid ? row_bound(best_nn) : 0.f;
      float vh = (valid && lo_i != 0) ? row_bound(best_hd) : 0.f;
      const uint32_t bn = __reduce_max_sync(0xffffffffu, __float_as_uint(fmaxf(vn, 0.f)));
      const uint32_t bh = __reduce_max_sync(0xffffffffu, __float_as_uint(fmaxf(vh, 0.f)));
      return make_uint2(bn, bh);
    };
    for (;;) {
      const uint32_t s = seq % G_SIDE_SLOTS;
      mbar_wait(&S.side_full[s], (seq / G_SIDE_SLOTS) & 1u);
      const float* rec = S.side + s * G_SIDE_FLOATS;
      const GMeta m = *reinterpret_cast<const GMeta*>(rec);
      if (m.flags & G_EXIT) break;
      if (m.flags & G_END) {
        if (cur_item == m.item) {
          if (valid) {
            atomicMin(a.key_nn + (row - g.row_begin), best_nn);
            atomicMin(a.key_hd + (row - g.row_begin), best_hd);
          }
          // what this warp's rows still accept bounds every later item of the row tile (the other warpgroup's warp of the
          // same quarter holds the same rows: both bounds are valid, atomicMin keeps the tighter)
          const uint2 b = publish();
          if (lane == 0) {
            const size_t rt = (size_t) m.row_tile * g.ra + my_tile;
            atomicMin(reinterpret_cast<unsigned int*>(a.thr_nn) + rt * 4 + quarter, b.x);
            atomicMin(reinterpret_cast<unsigned int*>(a.thr_hd) + rt * 4 + quarter, b.y);
          }
        }
        cur_item = 0xfffffffeu;
      } else {
        const bool first_of_item = cur_item != m.item;
        if (first_of_item) {
          cur_item = m.item;
          row = g.row_begin + (m.row_tile * (uint32_t) g.ra + my_tile) * GT + row_in_tile;
          valid = row < g.row_end;
          xn = valid ? __ldg(g.gnorm + row) : INFINITY;
          rho = g.rho_c * (sqrtf(xn) + sqrtf(g.nymax));
          eps = g.c_acc * (xn + g.nymax);
          best_nn = valid ? a.key_nn[row - g.row_begin] : 0ull;
          best_hd = valid ? a.key_hd[row - g.row_begin] : 0ull;
          lo_i = valid ? __ldg(a.lo + row) : 0u;
          lor = valid ? __ldg(a.lof + row) + a.lo_bias : 0.f;
          t_nn = valid ? thr(g_key_d2(best_nn)) : -INFINITY;
          // a frame nobody has a lower free energy than has no such neighbour: do not let it hold the filter open
          t_hd = (valid && lo_i != 0) ? thr(g_key_d2(best_hd)) : t_nn;
          set_dl();
        }
        const uint32_t acc = (tiles & ((uint32_t) (G_ACC / g.ra) - 1u)) * (uint32_t) g.ra + my_tile;
        ++tiles;
        mbar_wait(&S.tmem_full[acc], (uses >> acc) & 1u);
        uses ^= 1u << acc;
        tc_fence_after();
        if (wg < (uint32_t) g.ra) ++n_tiles;                 // 128 x 128 units: one count per row tile of the record
        const float* ny = rec + 4;
        const float* lc = rec + 4 + GT;
        bool improved = false;
        auto proc = [&](const uint32_t (&v)[G_LD], int c0) {
#pragma unroll
          for (int h = 0; h < G_LD; h += G_HALF) {
            // two-level filter: most tile pairs are far (only the lower-free-energy search reaches them) and hold nothing below
            // even the looser threshold t_hd >= t_nn: one FFMA + one min per pair settles those; the per-pair free-energy-aware
            // threshold is evaluated only where the group's minimum gets under t_hd
            float w[G_HALF];
            float wmin = INFINITY;
#pragma unroll
            for (int c4 = 0; c4 < G_HALF; c4 += 4) {
              const float4 n4 = *reinterpret_cast<const float4*>(ny + c0 + h + c4);
              const float nn[4] = {n4.x, n4.y, n4.z, n4.w};
#pragma unroll
              for (int c = 0; c < 4; ++c) {
                w[c4 + c] = fmaf(__uint_as_float(v[h + c4 + c]), -2.0f, nn[c]);
                wmin = fminf(wmin, w[c4 + c]);
              }
            }
            bool any = false;
            if (wmin < t_hd) {
#pragma unroll
              for (int c4 = 0; c4 < G_HALF; c4 += 4) {
                const float4 l4 = *reinterpret_cast<const float4*>(lc + c0 + h + c4);
                const float ll[4] = {l4.x, l4.y, l4.z, l4.w};
#pragma unroll
                for (int c = 0; c < 4; ++c) any |= w[c4 + c] < fmaf(__saturatef(lor - ll[c]), dl, t_nn);
              }
            }
            if (__any_sync(0xffffffffu, any)) {
              // warp-uniform region: the exact distance of a candidate is computed by the whole warp (dist2_exact_coop)
#pragma unroll
              for (int c = 0; c < G_HALF; ++c) scratch[c * G_EPI_THREADS] = __uint_as_float(v[h + c]);
#pragma unroll 1
              for (int c = 0; c < G_HALF; ++c) {
                const float wc = fmaf(scratch[c * G_EPI_THREADS], -2.0f, ny[c0 + h + c]);
                const uint32_t j = m.col0 + (uint32_t) (c0 + h + c);            // the same column for every lane
                const bool pass = wc < fmaf(__saturatef(lor - lc[c0 + h + c]), dl, t_nn);
                if (pass) ++n_slow;
                const bool hd_cand = j < g.n && __ldg(a.lo + min(j, g.n - 1u)) < lo_i;
                const bool want = pass && j != row && j < g.n && valid && (wc < t_nn || (hd_cand && wc < t_hd));
                uint32_t mk = __ballot_sync(0xffffffffu, want);
                float d2 = 0.f;
                while (mk) {
                  const int src = __ffs(mk) - 1;
                  mk &= mk - 1;
                  const float t = dist2_exact_coop(g.xR, g.d, __shfl_sync(0xffffffffu, row, src), j, lane);
                  if (lane == src) d2 = t;
                }
                if (!want) continue;
                ++n_exact;
                if (!(d2 < FLT_MAX)) continue;
                const unsigned long long key = ((unsigned long long) __float_as_uint(d2) << 32) | __ldg(a.perm + j);
                bool changed = false;
                if (key < best_nn) {
                  best_nn = key;
                  t_nn = thr(d2);
                  if (lo_i == 0) t_hd = t_nn;
                  changed = true;
                }
                if (hd_cand && key < best_hd) {
                  best_hd = key;
                  t_hd = thr(d2);
                  changed = true;
                }
                if (changed) { set_dl(); improved = true; }
              }
            }
          }
        };
        // (the early-release variant spills here: the neighbour state needs the registers)
        if (g.ra == 2) g_epilogue_tile(S, tmem_base, quarter, acc * (uint32_t) GT, acc, (int) (wg >> 1) * 64, 4, lane, proc);
        else g_epilogue_tile(S, tmem_base, quarter, acc * (uint32_t) GT, acc, (int) wg * 32, 2, lane, proc);
        // bounds for the producer's dynamic pruning of this item: one slot pair per epilogue warp
        if (first_of_item || __any_sync(0xffffffffu, improved)) {
          const uint2 b = publish();
          if (lane == 0) {
            S.wthr[2 * (warp - 2)] = ((unsigned long long) m.item << 32) | b.x;
            S.wthr[2 * (warp - 2) + 1] = ((unsigned long long) m.item << 32) | b.y;
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&S.side_empty[s]);
      ++seq;
    }
    for (int o = 16; o > 0; o >>= 1) {
      n_slow += __shfl_xor_sync(0xffffffffu, n_slow, o);
      n_exact += __shfl_xor_sync(0xffffffffu, n_exact, o);
    }
    if (lane == 0 && g.stats) {
      if (n_slow) atomicAdd(g.stats, (unsigned long long) n_slow);
      if (n_exact) atomicAdd(g.stats + 1, (unsigned long long) n_exact);
      if (n_tiles && quarter == 0) atomicAdd(g.stats + 3, (unsigned long long) n_tiles);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, G_ACC * GT);
}

// per 32-row quarter of every row tile: bounds after seeding (max over the rows' seeded d2, margins included) and the
// largest filter rank among the rows that can have a lower-free-energy neighbour
__global__ void gnn_tile_thr_kernel(const unsigned long long* __restrict__ key_nn, const unsigned long long* __restrict__ key_hd,
                                    const uint32_t* __restrict__ lo, const float* __restrict__ lof, float lo_bias, uint32_t row_begin,
                                    uint32_t row_end, float e_rel, float slack, float* __restrict__ thr_nn, float* __restrict__ thr_hd,
                                    float* __restrict__ lormax) {
  const uint32_t rb = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t i = row_begin + rb * GT + threadIdx.x;
  float vn = 0.f, vh = 0.f, lm = 0.f;
  if (i < row_end) {
    const float dn = __uint_as_float((uint32_t) (key_nn[i - row_begin] >> 32));
    vn = (fmaf(e_rel, dn, dn) + slack) * 1.00001f;
    if (!(vn < INFINITY)) vn = INFINITY;
    if (lo[i] != 0) {
      const float dh = __uint_as_float((uint32_t) (key_hd[i - row_begin] >> 32));
      vh = (fmaf(e_rel, dh, dh) + slack) * 1.00001f;
      if (!(vh < INFINITY)) vh = INFINITY;
      lm = lof[i] + lo_bias;
    }
  }
  for (int o = 16; o > 0; o >>= 1) {
    vn = fmaxf(vn, __shfl_xor_sync(0xffffffffu, vn, o));
    vh = fmaxf(vh, __shfl_xor_sync(0xffffffffu, vh, o));
    lm = fmaxf(lm, __shfl_xor_sync(0xffffffffu, lm, o));
  }
  if (lane == 0) {
    thr_nn[(size_t) rb * 4 + warp] = vn;
    thr_hd[(size_t) rb * 4 + warp] = vh;
    lormax[(size_t) rb * 4 + warp] = lm;
  }
}

// tcgen05.mma kind::tf32 128 x 128 x 8 issued back to back on resident operands: the denominator of the tensor roofline
// (out[blockIdx.x] = cycles for `iters` MMAs)
__global__ void __launch_bounds__(64, 1) tf32_peak_kernel(int iters, long long* out) {
  extern __shared__ unsigned char g_smem_raw[];
  __shared__ uint64_t fin;
  __shared__ uint32_t taddr;
  unsigned char* base = g_smem_raw + ((1024u - (smem_u32(g_smem_raw) & 1023u)) & 1023u);
  float* a = reinterpret_cast<float*>(base);
  for (int i = threadIdx.x; i < G_CHUNK_FLOATS * 8; i += blockDim.x) {
    uint32_t h = (uint32_t) i * 2654435761u + blockIdx.x * 40503u;
    h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
    a[i] = to_tf32(((float) (h & 0xffffffu) / 16777216.0f - 0.5f) * 2.0f);         // random operands (power draw like real data)
  }
  if (threadIdx.x == 0) { mbar_init(&fin, 1); fence_mbar_init(); }
  __syncthreads();
  const int warp = uniform_warp();
  if (warp == 1) tmem_alloc(&taddr, 2 * GT);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = *reinterpret_cast<volatile uint32_t*>(&taddr);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 1) {
    const bool leader = elect_one();
    const uint64_t a0 = g_smem_desc(smem_u32(a)), b0 = g_smem_desc(smem_u32(a + 4 * G_CHUNK_FLOATS));
    const long long t0 = clock64();
    if (leader) {
      for (uint32_t i = 0; i < (uint32_t) iters; ++i) {
        const uint64_t off = (uint64_t) ((i >> 2) & 3) * (G_CHUNK_BYTES >> 4) + 2 * (i & 3);
        tc_mma_tf32(tb + ((i >> 4) & 1) * (uint32_t) GT, a0 + off, b0 + off, G_IDESC, (i & 15) ? 1u : 0u);
      }
      tc_commit(&fin);
      mbar_wait(&fin, 0);
      out[blockIdx.x] = clock64() - t0;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tb, 2 * GT);
}

#endif  // DCB_GEMM_KERNELS

}  // namespace dcb
