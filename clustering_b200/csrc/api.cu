// libdcb200.so: the C ABI of include/dcb200.h -- host orchestration of the sm_100a kernels in kernels.cuh.
// No CPU fallback: every compute entry point needs a CUDA device and fails loudly otherwise.
#include "../../include/dcb200.h"

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cuda_runtime.h>
#include <float.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "kernels.cuh"
#include "gemm_kernels.cuh"
#include "launch.h"

using namespace dcb;

// ------------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------------
static thread_local std::string g_err;
static int fail(const std::string& msg) {
  g_err = msg;
  return 1;
}
extern "C" int dcb200_internal_fail(const char* msg) { return fail(msg); }
#define CK(call)                                                                                     \
  do {                                                                                               \
    cudaError_t e__ = (call);                                                                        \
    if (e__ != cudaSuccess)                                                                          \
      return fail(std::string(#call) + ": " + cudaGetErrorString(e__) + " (" + __FILE__ + ":" +      \
                  std::to_string(__LINE__) + ")");                                                   \
  } while (0)
#define CKI(call)                     \
  do {                                \
    int r__ = (call);                 \
    if (r__) return r__;              \
  } while (0)

// DCB200_TRACE=1: wall-clock phases of the host-pointer entry points on stderr (diagnostics only)
#include <chrono>
namespace {
struct Trace {
  bool on;
  const char* what;
  std::chrono::steady_clock::time_point t0, last;
  explicit Trace(const char* w) : what(w) {
    static const bool enabled = getenv("DCB200_TRACE") != nullptr;
    on = enabled;
    t0 = last = std::chrono::steady_clock::now();
  }
  void lap(const char* phase) {
    if (!on) return;
    const auto now = std::chrono::steady_clock::now();
    fprintf(stderr, "[dcb200] %s: %-18s %8.3f ms\n", what, phase, std::chrono::duration<double, std::milli>(now - last).count());
    last = now;
  }
  ~Trace() {
    if (on) fprintf(stderr, "[dcb200] %s: total %8.3f ms\n", what, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
  }
};
}  // namespace

// ------------------------------------------------------------------------------------------------
// small device kernels (layout, free energies, ordering, finalisation)
// ------------------------------------------------------------------------------------------------
namespace {

constexpr int CENTRE_SAMPLES = 65536;
constexpr int ORDER_MAX_DIMS = 12;      // dims that can enter the spatial ordering key (60 key bits are shared between them)

// centre[k] / spread[k] = mean and standard deviation of a strided sample of the frames (any centre is
// valid: it only tightens the error band of the fast path; the spread only scales the ordering key).
// One block per dim, fixed-order tree => deterministic.
__global__ void centre_kernel(const float* __restrict__ coords, size_t n, int d, float* __restrict__ centre,
                              float* __restrict__ spread) {
  __shared__ double part[256], part2[256];
  const int k = blockIdx.x;
  const size_t stride = n > CENTRE_SAMPLES ? n / CENTRE_SAMPLES : 1;
  const size_t m = (n + stride - 1) / stride;
  double s = 0.0, s2 = 0.0;
  for (size_t q = threadIdx.x; q < m; q += blockDim.x) {
    const double v = (double) coords[q * stride * d + k];
    s += v;
    s2 += v * v;
  }
  part[threadIdx.x] = s;
  part2[threadIdx.x] = s2;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int) threadIdx.x < o) {
      part[threadIdx.x] += part[threadIdx.x + o];
      part2[threadIdx.x] += part2[threadIdx.x + o];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const double mean = part[0] / (double) m;
    const double var = fmax(part2[0] / (double) m - mean * mean, 0.0);
    const float c = (float) mean, sd = (float) sqrt(var);
    centre[k] = (c == c && fabsf(c) < FLT_MAX) ? c : 0.f;
    spread[k] = (sd == sd && sd > 0.f && sd < FLT_MAX) ? sd : 1.f;
  }
}

// Spatial ordering key: Hilbert-curve index of the first m = min(d, 6) dims, each quantised to `bits` bits over
// centre +- 4 spread (Skilling's transpose algorithm: Gray-code the cell coordinates level by level, then interleave
// the bits).  Frames that are close in space end up close in the array, so that the column tiles (and the row
// groups) have small bounding boxes and whole tiles can be skipped; unlike a Z-order curve the Hilbert curve never
// jumps, so a run of 128 consecutive frames never straddles two far-apart cells.
__global__ void spatial_keys_kernel(const float* __restrict__ coords, size_t n, int d, const float* __restrict__ centre,
                                    const float* __restrict__ spread, int m, int bits, unsigned long long* __restrict__ keys,
                                    uint32_t* __restrict__ iota) {
  const size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t q[ORDER_MAX_DIMS];
  for (int k = 0; k < ORDER_MAX_DIMS; ++k) q[k] = 0;
  const float cells = (float) (1u << bits);
  for (int k = 0; k < m; ++k) {
    const float u = (coords[i * d + k] - centre[k]) / (8.0f * spread[k]) + 0.5f;
    const float v = fminf(fmaxf(u * cells, 0.0f), cells - 1.0f);
    q[k] = (v == v) ? (uint32_t) v : 0u;
  }
  if (m > 1) {
    const uint32_t top = 1u << (bits - 1);
    for (uint32_t Q = top; Q > 1; Q >>= 1) {           // inverse undo of the excess work
      const uint32_t P = Q - 1;
      for (int k = 0; k < m; ++k) {
        if (q[k] & Q) {
          q[0] ^= P;
        } else {
          const uint32_t t = (q[0] ^ q[k]) & P;
          q[0] ^= t;
          q[k] ^= t;
        }
      }
    }
    for (int k = 1; k < m; ++k) q[k] ^= q[k - 1];      // Gray encode
    uint32_t t = 0;
    for (uint32_t Q = top; Q > 1; Q >>= 1)
      if (q[m - 1] & Q) t ^= Q - 1;
    for (int k = 0; k < m; ++k) q[k] ^= t;
  }
  unsigned long long key = 0;
  for (int b = bits - 1; b >= 0; --b)
    for (int k = 0; k < m; ++k) key = (key << 1) | ((q[k] >> b) & 1u);
  keys[i] = key;
  iota[i] = (uint32_t) i;
}

// One block per column tile (tj = 128 or 64 frames, one thread per frame).  Row-major coords ->
//   xT   [d][ld]     original values (padding: NaN, so that no exact '<' ever passes)
//   cT   [tiles][(d+1)*tj + dp]  tile-major records (one bulk copy each).  Pack rows: y' = x - c_t with c_t = mean of the
//                    tile's frames (any point near the tile works; the mean keeps |y'| smallest): rows 0..d-1 = -2 y',
//                    row d = |y'|^2 (padding: 0, ..., 0, +inf: the accumulator of a padded column is +inf for every row).
//                    Header [dp]: c_t[0..d-1], max |y'|^2 over the tile's real frames, then the tile's bounding box
//                    lo[d], hi[d] in globally centred coordinates
//   bbox [ld/64][2d] bounding boxes of 64-frame groups in globally centred coordinates (tile pruning)
// in the order given by perm (position -> frame; nullptr = frame order).  Fixed-order reductions => the same
// bits on every GPU.
__global__ void pack_tiles_kernel(const float* __restrict__ coords, size_t n, int d, size_t ld, int dp,
                                  const float* __restrict__ centre, const uint32_t* __restrict__ perm, float* __restrict__ xT,
                                  float* __restrict__ cT, float* __restrict__ bbox, unsigned int* __restrict__ maxnorm_bits) {
  __shared__ float sh_sum[4], sh_lo[4], sh_hi[4];
  const int tj = blockDim.x;                      // 128 or 64
  const int nw = tj / 32;
  const size_t tile = blockIdx.x;
  float* __restrict__ rec = cT + tile * ((size_t) (d + 1) * tj + dp);       // this tile's record: pack rows, then header
  float* __restrict__ tcen = rec + (size_t) (d + 1) * tj;
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const size_t p = tile * tj + t;
  const bool real = p < n;
  const size_t src = real ? (perm ? (size_t) perm[p] : p) : 0;
  const size_t first = tile * tj;
  const float cnt = first < n ? (float) (n - first < (size_t) tj ? n - first : (size_t) tj) : 0.f;
  const float nan = __int_as_float(0x7fc00000);
  float nrm_local = 0.f, nrm_global = 0.f;
  for (int k = 0; k < d; ++k) {
    const float x = real ? coords[src * d + k] : 0.f;
    const float xg = x - centre[k];
    float sum = real ? x : 0.f, lo = real ? xg : INFINITY, hi = real ? xg : -INFINITY;
    for (int o = 16; o > 0; o >>= 1) {
      sum += __shfl_xor_sync(0xffffffffu, sum, o);
      lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
      hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if (lane == 0) { sh_sum[warp] = sum; sh_lo[warp] = lo; sh_hi[warp] = hi; }
    __syncthreads();
    float tot = 0.f;
    for (int q = 0; q < nw; ++q) tot += sh_sum[q];
    float c = cnt > 0.f ? tot / cnt : 0.f;
    if (!(fabsf(c) < FLT_MAX)) c = 0.f;
    if (lane == 0 && (warp & 1) == 0) {           // 64-frame group = two consecutive warps
      const size_t grp = p / 64;
      bbox[grp * 2 * d + k] = fminf(sh_lo[warp], sh_lo[warp + 1]);
      bbox[grp * 2 * d + d + k] = fmaxf(sh_hi[warp], sh_hi[warp + 1]);
    }
    if (t == 0) {                                 // box of the whole tile, for the consumers' warp-level pruning
      float tlo = INFINITY, thi = -INFINITY;
      for (int q = 0; q < nw; ++q) { tlo = fminf(tlo, sh_lo[q]); thi = fmaxf(thi, sh_hi[q]); }
      tcen[d + 1 + k] = tlo;
      tcen[2 * d + 1 + k] = thi;
    }
    __syncthreads();
    const float yl = x - c;
    xT[(size_t) k * ld + p] = real ? x : nan;
    rec[(size_t) k * tj + t] = real ? -2.0f * yl : 0.f;
    nrm_local = fmaf(yl, yl, nrm_local);
    nrm_global = fmaf(xg, xg, nrm_global);
    if (t == 0) tcen[k] = c;
  }
  rec[(size_t) d * tj + t] = real ? nrm_local : INFINITY;
  float mx = real ? nrm_local : 0.f;
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if (lane == 0) sh_sum[warp] = mx;
  __syncthreads();
  if (t == 0) {
    float m = 0.f;
    for (int q = 0; q < nw; ++q) m = fmaxf(m, sh_sum[q]);
    tcen[d] = m;
    for (int k = 3 * d + 1; k < dp; ++k) tcen[k] = 0.f;
  }
  if (real) atomicMax(maxnorm_bits, __float_as_uint(nrm_global));      // >= 0: the bit pattern orders like the value
}

// dst[a][perm[p]] = src[a][p]: position order -> frame order
__global__ void to_frame_order_kernel(const uint32_t* __restrict__ src, uint32_t* __restrict__ dst, const uint32_t* __restrict__ perm,
                                      size_t n, size_t n_arrays) {
  const size_t p = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const size_t o = perm[p];
  for (size_t a = 0; a < n_arrays; ++a) dst[a * n + o] = src[a * n + p];
}

// lo_pos[p] = lo_frame[perm[p]] after lo_frame[fe_perm[q]] = lo_sorted[q]
__global__ void scatter_by_kernel(const uint32_t* __restrict__ src, const uint32_t* __restrict__ index, size_t n,
                                  uint32_t* __restrict__ dst) {
  const size_t q = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (q < n) dst[index[q]] = src[q];
}
__global__ void gather_by_kernel(const uint32_t* __restrict__ src, const uint32_t* __restrict__ index, size_t n,
                                 uint32_t* __restrict__ dst) {
  const size_t p = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (p < n) dst[p] = src[index[p]];
}
// lo by position (gathered through perm) and the same ranks as floats for the neighbour filter: (float)(lo >> shift),
// exact for shift == 0 (n <= 2^24); +inf in the padding so that a padded column is never a candidate
__global__ void lo_by_position_kernel(const uint32_t* __restrict__ lo_frame, const uint32_t* __restrict__ perm, size_t n, size_t ld,
                                      int shift, uint32_t* __restrict__ lo, float* __restrict__ lof) {
  const size_t p = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= ld) return;
  if (p >= n) { lof[p] = INFINITY; return; }
  const uint32_t v = lo_frame[perm[p]];
  lo[p] = v;
  lof[p] = (float) (v >> shift);
}
__global__ void iota_kernel(uint32_t* p, size_t n) {
  const size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = (uint32_t) i;
}

// pops[r][i] = 1 + mult[r] * sum_{b <= bin[r]} cnt[b][i]    (self counted by the initial 1, density_clustering.cpp:133;
// a radius listed m times is one map entry incremented m times per hit, :131-134,:180)
// cumulative (count mode): cnt[b] already holds #{d2 < rad2[b]}; self[q] = 1 where the kernel counted the frame itself
// (count and bin mode count it through d2(i,i) = 0 < r^2; the histogram kernel skips j == i)
struct FinalizeArgs {
  int n_out;
  int cumulative;
  uint32_t self[MAX_BINS * 4];
  int out_row[MAX_BINS * 4];
  int bin[MAX_BINS * 4];
  uint32_t mult[MAX_BINS * 4];
};
__global__ void pops_finalize_kernel(const uint32_t* __restrict__ cnt, size_t ld_cnt, size_t rows, size_t ld_out,
                                     uint32_t* __restrict__ pops, const FinalizeArgs f) {
  const size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows) return;
  for (int q = 0; q < f.n_out; ++q) {
    uint32_t c = 0;
    if (f.cumulative) {
      c = cnt[(size_t) f.bin[q] * ld_cnt + i] - f.self[q];
    } else {
      for (int b = 0; b <= f.bin[q]; ++b) c += cnt[(size_t) b * ld_cnt + i];     // d2 < rad2[bin] = all bins up to it
      c -= f.self[q];
    }
    pops[(size_t) f.out_row[q] * ld_out + i] = 1u + f.mult[q] * c;
  }
}

__global__ void max_u32_kernel(const uint32_t* __restrict__ v, size_t n, unsigned int* __restrict__ out) {
  unsigned int m = 0;
  for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x) m = max(m, v[i]);
  for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMax(out, m);
}

// calculate_free_energies (density_clustering.cpp:197-212) in its compiled form:
//   inv = 1.0f / max_pop;  t = (float)pops[i] * inv;  fe = (float)(-log((double) t))
__global__ void fe_kernel(const uint32_t* __restrict__ pops, size_t n, const unsigned int* __restrict__ max_dev,
                          uint32_t max_host, float* __restrict__ fe) {
  const size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t mx = max_host ? max_host : *max_dev;
  const float inv = __fdiv_rn(1.0f, __uint2float_rn(mx));
  const float t = __fmul_rn(__uint2float_rn(pops[i]), inv);
  fe[i] = (float) (-log((double) t));
}

// order-preserving key of a float (-0.0 and +0.0 compare equal, like the reference's '<' on floats)
__global__ void fe_keys_kernel(const float* __restrict__ fe, size_t n, uint32_t* __restrict__ keys, uint32_t* __restrict__ iota) {
  const size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t u = __float_as_uint(fe[i] + 0.0f);
  keys[i] = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
  iota[i] = (uint32_t) i;
}

// lo[p] = number of frames with a strictly lower free energy = first position of p's run of equal keys
__global__ void lower_bound_kernel(const uint32_t* __restrict__ keys, size_t n, uint32_t* __restrict__ lo) {
  const size_t p = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const uint32_t k = keys[p];
  size_t a = 0, b = p;
  while (a < b) {
    const size_t mid = (a + b) >> 1;
    if (keys[mid] < k) a = mid + 1; else b = mid;
  }
  lo[p] = (uint32_t) a;
}

// Seeds the neighbour keys of positions [row_begin,row_end) with the best of the 2*W frames next to them in the
// context's spatial order (exact arithmetic, real candidates): the pair scan then starts from tight thresholds,
// prunes far tiles from its first item on and re-evaluates only the few columns that really improve on the seed.
__global__ void nn_seed_kernel(const float* __restrict__ xT, size_t ld, int d, uint32_t n, const uint32_t* __restrict__ perm,
                               const uint32_t* __restrict__ lo, uint32_t row_begin, uint32_t row_end, uint32_t rb_stride, uint32_t n_out,
                               int W, unsigned long long none, unsigned long long* __restrict__ key_nn,
                               unsigned long long* __restrict__ key_hd) {
  const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;            // index in the launch's output arrays (kernels.cuh: out_index)
  if (q >= n_out) return;
  const uint32_t p = row_begin + (q / ROWS_PER_CTA) * rb_stride * ROWS_PER_CTA + q % ROWS_PER_CTA;
  if (p >= row_end) {
    key_nn[q] = none;
    key_hd[q] = none;
    return;
  }
  unsigned long long knn = none, khd = none;
  const uint32_t lo_p = lo[p];
  for (int o = 1; o <= W; ++o) {
    for (int side = 0; side < 2; ++side) {
      const long long jj = side ? (long long) p + o : (long long) p - o;
      if (jj < 0 || jj >= (long long) n) continue;
      const uint32_t j = (uint32_t) jj;
      const float d2 = dist2_exact(xT, ld, d, p, j);
      if (!(d2 < FLT_MAX)) continue;
      const unsigned long long key = ((unsigned long long) __float_as_uint(d2) << 32) | perm[j];
      knn = min(knn, key);
      if (lo[j] < lo_p) khd = min(khd, key);
    }
  }
  key_nn[q] = knn;
  key_hd[q] = khd;
}

// bounding boxes of the row blocks of one launch = union of the 64-frame group boxes they touch (one warp per block)
__global__ void row_bbox_kernel(const float* __restrict__ bbox, int d, uint32_t row_begin, uint32_t row_end, uint32_t rb_stride,
                                uint32_t n_row_blocks, float* __restrict__ rbbox) {
  const uint32_t rb = blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (rb >= n_row_blocks) return;
  const uint32_t r0 = row_begin + rb * rb_stride * ROWS_PER_CTA;
  const uint32_t r1 = min(r0 + (uint32_t) ROWS_PER_CTA, row_end);
  const uint32_t g0 = r0 / 64, g1 = (r1 - 1) / 64;
  for (int k = lane; k < d; k += 32) {
    float lo = INFINITY, hi = -INFINITY;
    for (uint32_t q = g0; q <= g1; ++q) {
      lo = fminf(lo, bbox[(size_t) q * 2 * d + k]);
      hi = fmaxf(hi, bbox[(size_t) q * 2 * d + d + k]);
    }
    rbbox[(size_t) rb * 2 * d + k] = lo;
    rbbox[(size_t) rb * 2 * d + d + k] = hi;
  }
}

// sbbox[s] = union of the boxes of the 64-frame groups of super-tile s (SUPER_FRAMES consecutive positions): the coarse level
// of the producers' tile pruning (one warp per super-tile)
__global__ void super_bbox_kernel(const float* __restrict__ bbox, int d, uint32_t n_groups, uint32_t n_super, float* __restrict__ sbbox) {
  const uint32_t sp = blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (sp >= n_super) return;
  const uint32_t g0 = sp * (SUPER_FRAMES / 64), g1 = min(g0 + SUPER_FRAMES / 64, n_groups);
  for (int k = lane; k < d; k += 32) {
    float lo = INFINITY, hi = -INFINITY;
    for (uint32_t q = g0; q < g1; ++q) {
      lo = fminf(lo, bbox[(size_t) q * 2 * d + k]);          // fminf / fmaxf drop the NaN boxes of all-padding groups
      hi = fmaxf(hi, bbox[(size_t) q * 2 * d + d + k]);
    }
    sbbox[(size_t) sp * 2 * d + k] = lo;
    sbbox[(size_t) sp * 2 * d + d + k] = hi;
  }
}

// blk_thr[rb][w] = bound (d2 units, with the pruning margins) of what the rows of block rb owned by consumer warp w
// still accept after seeding: max over its rows of the seeded d2 (nearest / nearest with lower free energy)
__global__ void nn_block_thr_kernel(const unsigned long long* __restrict__ key_nn, const unsigned long long* __restrict__ key_hd,
                                    const uint32_t* __restrict__ lo, uint32_t row_begin, uint32_t row_end, uint32_t rb_stride,
                                    uint32_t n_row_blocks, float e_rel, float slack, float* __restrict__ blk_thr) {
  const uint32_t rb = blockIdx.x;
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;      // blockDim = N_CONSUMERS: thread = consumer slot
  float v = 0.f;
  for (int r = 0; r < RI; ++r) {
    const uint32_t slot = (uint32_t) w * (32u * RI) + (uint32_t) r * 32u + (uint32_t) lane;
    const uint32_t i = row_begin + rb * rb_stride * ROWS_PER_CTA + slot;   // Rows::row
    if (i >= row_end) continue;
    const uint32_t q = rb * ROWS_PER_CTA + slot;                           // out_index
    const float dn = __uint_as_float((uint32_t) (key_nn[q] >> 32));
    const float dh = lo[i] == 0 ? dn : __uint_as_float((uint32_t) (key_hd[q] >> 32));
    v = fmaxf(v, fmaxf(dn, dh));
  }
  v = (fmaf(e_rel, v, v) + slack) * 1.00001f;
  if (!(v < INFINITY)) v = INFINITY;
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  if (lane == 0) blk_thr[(size_t) rb * N_CONSUMER_WARPS + w] = v;
}

__global__ void nn_finish_kernel(const unsigned long long* __restrict__ knn, const unsigned long long* __restrict__ khd,
                                 const uint32_t* __restrict__ perm, size_t n, uint32_t* __restrict__ nn_idx,
                                 float* __restrict__ nn_d2, uint32_t* __restrict__ hd_idx, float* __restrict__ hd_d2) {
  const size_t p = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const size_t o = perm[p];
  const unsigned long long a = knn[p], b = khd[p];
  nn_idx[o] = (uint32_t) a;
  nn_d2[o] = __uint_as_float((uint32_t) (a >> 32));
  hd_idx[o] = (uint32_t) b;
  hd_d2[o] = __uint_as_float((uint32_t) (b >> 32));
}

__global__ void uf_flatten_kernel(uint32_t* parent, size_t m) {
  const size_t p = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= m) return;
  const uint32_t r = uf_find(parent, (uint32_t) p);
  parent[p] = r;                  // any interleaving leaves a valid forest with the same roots
}

// merges another forest over the same positions into `parent`
__global__ void uf_merge_kernel(uint32_t* parent, const uint32_t* __restrict__ other, size_t m) {
  const size_t p = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= m) return;
  const uint32_t q = other[p];
  if (q != p) uf_union(parent, (uint32_t) p, q);
}

// FFMA-only loop: the denominator of the FP32 roofline (16 independent chains per thread)
__global__ void ffma_peak_kernel(float* out, int iters, float a, float b) {
  float v[16];
#pragma unroll
  for (int q = 0; q < 16; ++q) v[q] = (float) (threadIdx.x + q);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int q = 0; q < 16; ++q) v[q] = fmaf(v[q], a, b);
  }
  float s = 0.f;
#pragma unroll
  for (int q = 0; q < 16; ++q) s += v[q];
  if (s == 123.456f) out[0] = s;
}

inline unsigned int blocks_for(size_t n, int bs) { return (unsigned int) ((n + bs - 1) / bs); }

struct Layout {
  float* xT = nullptr;
  float* cT = nullptr;
};

template <class T>
struct DevBuf {
  T* p = nullptr;
  size_t cap = 0;
  cudaError_t reserve(size_t n) {
    if (n <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    cudaError_t e = cudaMalloc(&p, n * sizeof(T));
    if (e == cudaSuccess) cap = n;
    return e;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
};

}  // namespace

// ------------------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------------------
struct dcb200_ctx {
  int device = 0;
  int sm_count = 0;
  cudaStream_t stream = nullptr;
  size_t n = 0, d = 0, ld = 0;
  bool spatial = false;             // frames are held in spatial (Hilbert curve) order; perm maps position -> frame
  DevBuf<float> xT, cT;             // [d][ld] original coords; tile-major column pack records (context order)
  DevBuf<float> bbox;               // [ld/64][2d]
  DevBuf<float> sbbox;              // [ceil(ld/SUPER_FRAMES)][2d] boxes of the super-tiles (coarse pruning level)
  DevBuf<float> rbbox;              // [row blocks of the current launch][2d]
  DevBuf<float> blk_thr;            // [row blocks][N_CONSUMER_WARPS] neighbour search pruning bounds
  DevBuf<uint32_t> perm;            // [n] position -> frame (identity when !spatial)
  DevBuf<uint32_t> lo;              // [n] per position: number of frames with strictly lower free energy
  DevBuf<float> lof;                // [ld] the same ranks as floats (>> lo_shift), +inf padded
  int lo_shift = 0;
  DevBuf<uint32_t> keys_a, keys_b, iota, tmp_u32, tmp2_u32;
  DevBuf<unsigned long long> skeys_a, skeys_b;     // spatial (Hilbert) keys, up to 60 bits
  DevBuf<unsigned char> cub_tmp;
  DevBuf<float> stage;              // row-major staging for host uploads
  DevBuf<float> centre;             // [2d] centre, spread
  DevBuf<uint32_t> cnt;
  DevBuf<float> lut;                // cell table of the bin-mode population kernel (one pass at a time)
  DevBuf<unsigned long long> knn, khd;
  DevBuf<uint32_t> shard_tmp;       // compact rows of a GEMM-form shard before they are spread into the padded shard layout
  DevBuf<uint32_t> io_u32;          // host-pointer entry points: device-side staging of inputs / outputs
  DevBuf<float> io_f32;
  unsigned int* scalars = nullptr;  // [0] work counter, [1] max norm bits, [2] max pop
  unsigned long long* stats = nullptr;   // [0] slow pairs, [1] exact pairs, [2] tiles streamed
  float maxnorm2 = 0.f;
  bool nn_ready = false;
  uint64_t layout_gen = 0;          // counts the layouts built on this context (sessions notice when theirs was replaced)
  uint64_t launches = 0;
  uint64_t pairs_scheduled = 0;     // pairs of the full row x column ranges of the scans since the last reset
  // GEMM-form (tcgen05) path for 17 <= d <= 256, spatial order only (gemm_kernels.cuh)
  bool gemm = false;
  int g_kc = 0, g_k8 = 0;
  size_t g_tiles = 0;               // ld / 128
  float g_nymax = 0.f;              // max |x~|^2
  DevBuf<float> gT, gnorm, xR;      // operand images, norms of the rounded values, row-major exact copy (context order)
  DevBuf<float> thdr;               // tile geometry: tcen [d][T], tlo [d][T], thi [d][T], trad [T]
  DevBuf<float> lbmat;              // lower bounds of the tile pairs for row tiles [lb_s0, lb_s1)
  uint32_t lb_s0 = 0, lb_s1 = 0;
  DevBuf<float> lomin, gthr;        // per-tile min rank; per row tile and quarter: thr_nn, thr_hd, lormax
  float* gcheck = nullptr;          // [2] diagnostics of the CHECK kernel
  unsigned long long* gprof = nullptr;   // [16] cycle counters of the GEMM-form kernels (DCB200_GEMM_PROF=1)
};

// tuning knobs read from the environment (diagnostics)
static int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  const int v = e ? atoi(e) : 0;
  return v > 0 ? v : dflt;
}

static float up(double v) {          // smallest float >= v
  float f = (float) v;
  if ((double) f < v) f = nextafterf(f, INFINITY);
  return f;
}

// Rounding-error bounds of the fast path (u = 2^-24).  Per tile the columns are y' = fl(y - c_t) and the rows
// x' = fl(x - c_t); X = |x'|^2, Y = |y'|^2, T = true squared distance, S = acc + xn in real arithmetic with
//   acc = FFMA chain started at the stored |y'|^2 over x'_k * (-2 y'_k),  xn = FFMA chain of x'_k^2:
//   |S - T''| <= D u (2X + 3Y)            (T'' = X + Y - 2 x'.y' in real arithmetic; D roundings per chain,
//                                          partial sums bounded by X + 2Y resp. X resp. Y)
//   |T'' - T| <= u (T + 2 (X + Y))         (roundings of the two subtractions of c_t)
//   |d2e - T| <= (D/4 + 8) u T             (reference order: sub, mul, <= ceil(D/4)+4 adds of non-negative terms)
// => |S - d2e| <= c_loc (X + Y) + e_rel T with the constants below (safety factor 1.5 on both).
// prune_slack bounds the rounding of the bounding-box arithmetic, done in globally centred coordinates
// (M = max |x - centre|^2).
static void error_bounds(size_t d, float maxnorm2, float* c_loc, float* e_rel, float* prune_slack) {
  const double u = ldexp(1.0, -24);
  *c_loc = up(1.5 * (3.0 * (double) d + 2.0) * u);
  *e_rel = up(1.5 * ((double) d / 4.0 + 9.0) * u);
  *prune_slack = up(1.5 * (5.0 * (double) d + 8.0) * u * (double) maxnorm2 + 1e-37);
}

// rows of the launch: blocks of ROWS_PER_CTA rows starting at row_begin + rb * rb_stride * ROWS_PER_CTA, below row_end
// (rb_stride = 1: the contiguous range; rb_stride = G with row_begin = g * ROWS_PER_CTA, row_end = n: block-cyclic shard g of G)
static size_t launch_row_blocks(size_t row_begin, size_t row_end, size_t rb_stride) {
  if (row_begin >= row_end) return 0;
  const size_t step = rb_stride * ROWS_PER_CTA;
  return (row_end - row_begin + step - 1) / step;
}
static size_t launch_rows(size_t row_begin, size_t row_end, size_t rb_stride) {      // valid rows = size of the compact outputs
  const size_t nb = launch_row_blocks(row_begin, row_end, rb_stride);
  if (nb == 0) return 0;
  const size_t last0 = row_begin + (nb - 1) * rb_stride * ROWS_PER_CTA;
  return (nb - 1) * ROWS_PER_CTA + std::min<size_t>(ROWS_PER_CTA, row_end - last0);
}

static int fill_geom(dcb200_ctx* c, size_t row_begin, size_t row_end, size_t rb_stride, int tj, int occupancy, uint32_t tiles_per_item,
                     ScanGeom* g, int* grid, uint32_t max_tiles_per_item = 0xffffffffu) {
  memset(g, 0, sizeof(*g));
  g->xT = c->xT.p;
  g->cT = c->cT.p;
  g->bbox = c->bbox.p;
  g->sbbox = env_int("DCB200_SUPER_PRUNE", 1) == 1 ? c->sbbox.p : nullptr;
  g->dp = (int) ((3 * c->d + 1 + 3) / 4 * 4);
  g->centre = c->centre.p;
  g->prune_thr = INFINITY;
  g->axis_prune = env_int("DCB200_AXIS_PRUNE", 1) == 1 ? 1 : 0;
  g->ld = c->ld;
  g->d = (int) c->d;
  g->n = (uint32_t) c->n;
  g->row_begin = (uint32_t) row_begin;
  g->row_end = (uint32_t) row_end;
  g->rb_stride = (uint32_t) rb_stride;
  g->n_row_blocks = (uint32_t) launch_row_blocks(row_begin, row_end, rb_stride);
  g->n_col_tiles = (uint32_t) (c->ld / tj);
  if (occupancy < 1) return fail("kernel does not fit on an SM (shared memory / registers)");
  *grid = c->sm_count * occupancy;
  // about 16 column items per row block: enough for load balance and for the column-step-major order
  // (own neighbourhood first), few enough that the producers' per-item work stays negligible.  A shard of a multi-GPU
  // run has 1/G of the row blocks: its column ranges shrink accordingly (down to the caller's minimum), so that the
  // heavy items -- the few column ranges around a block's own position -- still come to several waves over the CTAs
  // (C3 on 8 GPUs: 123 row blocks x 16 ranges left ~1.2 heavy items per CTA and the scan at 60 % of its one-GPU rate)
  const uint32_t want_items = (uint32_t) *grid * (uint32_t) std::max(1, env_int("DCB200_ITEMS_PER_CTA", 192));
  const uint32_t col_items = std::max(16u, (want_items + g->n_row_blocks - 1) / std::max(1u, g->n_row_blocks));
  tiles_per_item = std::min(std::max(tiles_per_item, (g->n_col_tiles + col_items - 1) / col_items), max_tiles_per_item);
  g->tiles_per_item = std::max(1u, std::min(tiles_per_item, g->n_col_tiles));
  g->n_col_items = (g->n_col_tiles + g->tiles_per_item - 1) / g->tiles_per_item;
  if ((uint64_t) g->n_row_blocks * g->n_col_items >= 0x7fffffffull) return fail("too many work items");
  *grid = (int) std::min<uint64_t>((uint64_t) *grid, (uint64_t) g->n_row_blocks * g->n_col_items);
  g->work_counter = c->scalars;
  g->stats = c->stats;
  error_bounds(c->d, c->maxnorm2, &g->c_loc, &g->e_rel, &g->prune_slack);
  CK(c->rbbox.reserve((size_t) g->n_row_blocks * 2 * c->d));
  g->rbbox = c->rbbox.p;
  row_bbox_kernel<<<blocks_for(g->n_row_blocks, 8), 256, 0, c->stream>>>(c->bbox.p, (int) c->d, g->row_begin, g->row_end, g->rb_stride,
                                                                         g->n_row_blocks, c->rbbox.p);
  c->launches += 1;
  c->pairs_scheduled += (uint64_t) launch_rows(row_begin, row_end, rb_stride) * (uint64_t) c->n;
  return 0;
}

#define DCB_DISPATCH(D) case D: return FN(D);
static cudaError_t launch_pops(int d, const PopsArgs& a, int grid, cudaStream_t st) {
  switch (d <= MAX_TEMPLATE_D ? d : 0) {
#define FN(D) launch_pops_d##D(a, grid, st)
    DCB_FOR_EACH_D(DCB_DISPATCH)
#undef FN
  }
  return cudaErrorInvalidValue;
}
static cudaError_t launch_nn(int d, const NnArgs& a, int grid, cudaStream_t st) {
  switch (d <= MAX_TEMPLATE_D ? d : 0) {
#define FN(D) launch_nn_d##D(a, grid, st)
    DCB_FOR_EACH_D(DCB_DISPATCH)
#undef FN
  }
  return cudaErrorInvalidValue;
}
static cudaError_t launch_screen(int d, const ScreenArgs& a, int grid, cudaStream_t st) {
  switch (d <= MAX_TEMPLATE_D ? d : 0) {
#define FN(D) launch_screen_d##D(a, grid, st)
    DCB_FOR_EACH_D(DCB_DISPATCH)
#undef FN
  }
  return cudaErrorInvalidValue;
}
static int occ_pops(int d, int nb) {
  switch (d <= MAX_TEMPLATE_D ? d : 0) {
#define FN(D) occupancy_pops_d##D(nb, d)
    DCB_FOR_EACH_D(DCB_DISPATCH)
#undef FN
  }
  return 0;
}
static cudaError_t launch_pops_count(int d, const PopsArgs& a, int grid, cudaStream_t st) {
  switch (d <= MAX_TEMPLATE_D ? d : 0) {
#define FN(D) launch_pops_count_d##D(a, grid, st)
    DCB_FOR_EACH_D(DCB_DISPATCH)
#undef FN
  }
  return cudaErrorInvalidValue;
}
static int occ_pops_count(int d, int nb) {
  switch (d <= MAX_TEMPLATE_D ? d : 0) {
#define FN(D) occupancy_pops_count_d##D(nb, d)
    DCB_FOR_EACH_D(DCB_DISPATCH)
#undef FN
  }
  return 0;
}
static cudaError_t launch_pops_bin(int d, const PopsArgs& a, int grid, cudaStream_t st) {
  switch (d <= MAX_TEMPLATE_D ? d : 0) {
#define FN(D) launch_pops_bin_d##D(a, grid, st)
    DCB_FOR_EACH_D(DCB_DISPATCH)
#undef FN
  }
  return cudaErrorInvalidValue;
}
static int occ_pops_bin(int d, int nb, int lut_k) {
  switch (d <= MAX_TEMPLATE_D ? d : 0) {
#define FN(D) occupancy_pops_bin_d##D(nb, lut_k, d)
    DCB_FOR_EACH_D(DCB_DISPATCH)
#undef FN
  }
  return 0;
}
static int occ_nn(int d) {
  switch (d <= MAX_TEMPLATE_D ? d : 0) {
#define FN(D) occupancy_nn_d##D(d)
    DCB_FOR_EACH_D(DCB_DISPATCH)
#undef FN
  }
  return 0;
}
static int occ_screen(int d) {
  switch (d <= MAX_TEMPLATE_D ? d : 0) {
#define FN(D) occupancy_screen_d##D(d)
    DCB_FOR_EACH_D(DCB_DISPATCH)
#undef FN
  }
  return 0;
}
static int tile_width(size_t d) { return d <= (size_t) MAX_TEMPLATE_D ? TileW<1>::tj : TileW<0>::tj; }

// ------------------------------------------------------------------------------------------------
// process-wide
// ------------------------------------------------------------------------------------------------
static int g_n_gpus = 0;

extern "C" int dcb200_version(void) { return DCB200_VERSION; }
extern "C" const char* dcb200_last_error(void) { return g_err.c_str(); }

extern "C" int dcb200_device_count(int* n) {
  if (!n) return fail("dcb200_device_count: null argument");
  *n = 0;
  cudaError_t e = cudaGetDeviceCount(n);
  if (e != cudaSuccess) {
    *n = 0;
    return fail(std::string("cudaGetDeviceCount: ") + cudaGetErrorString(e));
  }
  return 0;
}

extern "C" int dcb200_set_gpus(int n) {
  if (n < 0) return fail("dcb200_set_gpus: negative count");
  g_n_gpus = n;
  return 0;
}

// ------------------------------------------------------------------------------------------------
// session
// ------------------------------------------------------------------------------------------------
extern "C" int dcb200_ctx_create(int device, dcb200_ctx** out) {
  if (!out) return fail("dcb200_ctx_create: null argument");
  *out = nullptr;
  int n = 0;
  CKI(dcb200_device_count(&n));
  if (n == 0) return fail("dcb200: no CUDA device available (this library has no CPU fallback)");
  if (device < 0 || device >= n) return fail("dcb200_ctx_create: device index out of range");
  CK(cudaSetDevice(device));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10) return fail(std::string("dcb200 is built for sm_100a (B200); found ") + prop.name);
  dcb200_ctx* c = new dcb200_ctx();
  c->device = device;
  c->sm_count = prop.multiProcessorCount;
  auto init = [&]() -> int {
    CK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    CK(cudaMalloc(&c->scalars, 4 * sizeof(unsigned int)));
    CK(cudaMalloc(&c->stats, 4 * sizeof(unsigned long long)));
    CK(cudaMemsetAsync(c->scalars, 0, 4 * sizeof(unsigned int), c->stream));
    CK(cudaMemsetAsync(c->stats, 0, 4 * sizeof(unsigned long long), c->stream));
    return 0;
  };
  if (const int rc = init()) {              // nothing of a half-built context survives a failure
    if (c->scalars) cudaFree(c->scalars);
    if (c->stats) cudaFree(c->stats);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
    return rc;
  }
  *out = c;
  return 0;
}

extern "C" int dcb200_ctx_destroy(dcb200_ctx* c) {
  if (!c) return 0;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  c->xT.release(); c->cT.release(); c->bbox.release(); c->sbbox.release(); c->rbbox.release(); c->blk_thr.release();
  c->perm.release(); c->lo.release(); c->lof.release(); c->keys_a.release(); c->keys_b.release(); c->iota.release(); c->skeys_a.release(); c->skeys_b.release();
  c->tmp_u32.release(); c->tmp2_u32.release();
  c->cub_tmp.release(); c->stage.release(); c->centre.release(); c->cnt.release(); c->lut.release(); c->knn.release(); c->khd.release();
  c->io_u32.release(); c->io_f32.release(); c->shard_tmp.release();
  c->gT.release(); c->gnorm.release(); c->xR.release(); c->thdr.release(); c->lbmat.release(); c->lomin.release(); c->gthr.release();
  if (c->gcheck) cudaFree(c->gcheck);
  if (c->gprof) cudaFree(c->gprof);
  cudaFree(c->scalars);
  cudaFree(c->stats);
  cudaStreamDestroy(c->stream);
  delete c;
  return 0;
}

extern "C" void* dcb200_ctx_stream(dcb200_ctx* c) { return c ? (void*) c->stream : nullptr; }

extern "C" int dcb200_ctx_sync(dcb200_ctx* c) {
  if (!c) return fail("null context");
  CK(cudaSetDevice(c->device));
  CK(cudaStreamSynchronize(c->stream));
  return 0;
}

extern "C" int dcb200_ctx_stats(dcb200_ctx* c, uint64_t stats[6], int reset) {
  if (!c || !stats) return fail("null argument");
  CK(cudaSetDevice(c->device));
  unsigned long long h[4];
  CK(cudaMemcpyAsync(h, c->stats, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  stats[0] = c->launches;
  stats[1] = h[0];
  stats[2] = h[1];
  stats[3] = h[3];                                       // (warp, tile) scans: a warp owns 32*RI rows
  stats[4] = c->pairs_scheduled;
  stats[5] = c->gemm ? (uint64_t) GT * GT : (uint64_t) tile_width(c->d) * 32 * RI;
  if (reset) {
    CK(cudaMemsetAsync(c->stats, 0, 4 * sizeof(unsigned long long), c->stream));
    c->pairs_scheduled = 0;
  }
  return 0;
}

static int sort_pairs_u64(dcb200_ctx* c, const unsigned long long* keys_in, unsigned long long* keys_out, const uint32_t* vals_in,
                          uint32_t* vals_out, size_t n, int end_bit) {
  size_t tmp_bytes = 0;
  CK(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys_in, keys_out, vals_in, vals_out, (int) n, 0, end_bit, c->stream));
  CK(c->cub_tmp.reserve(tmp_bytes));
  CK(cub::DeviceRadixSort::SortPairs(c->cub_tmp.p, tmp_bytes, keys_in, keys_out, vals_in, vals_out, (int) n, 0, end_bit, c->stream));
  c->launches += 8;
  return 0;
}

static int sort_pairs_u32(dcb200_ctx* c, const uint32_t* keys_in, uint32_t* keys_out, const uint32_t* vals_in, uint32_t* vals_out,
                          size_t n, int end_bit) {
  size_t tmp_bytes = 0;
  CK(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys_in, keys_out, vals_in, vals_out, (int) n, 0, end_bit, c->stream));
  CK(c->cub_tmp.reserve(tmp_bytes));
  CK(cub::DeviceRadixSort::SortPairs(c->cub_tmp.p, tmp_bytes, keys_in, keys_out, vals_in, vals_out, (int) n, 0, end_bit, c->stream));
  c->launches += 4;
  return 0;
}

// The GEMM-form path serves the high-dimensional inputs where the pair scan really is a dense contraction (BASELINE config 5).
// DCB200_GEMM=0 keeps such inputs on the run-time-D FFMA kernels (A/B comparisons, tests).
static bool gemm_eligible(size_t n, size_t d, bool spatial) {
  const char* e = getenv("DCB200_GEMM");
  if (e && e[0] == '0') return false;
  if (!spatial || d < (size_t) G_MIN_D || d > (size_t) G_MAX_D) return false;
  const size_t tiles = (n + LD_ALIGN - 1) / LD_ALIGN * LD_ALIGN / GT;
  return tiles * tiles * sizeof(float) <= (size_t(4) << 30);        // the tile-pair lower bounds must fit comfortably
}

// lower bounds of the tile pairs for the row tiles of [row_begin, row_end), cached until the layout changes
static int ensure_tile_lb(dcb200_ctx* c, size_t row_begin, size_t row_end) {
  const uint32_t s0 = (uint32_t) (row_begin / GT), s1 = (uint32_t) ((row_end + GT - 1) / GT);
  if (s0 >= c->lb_s0 && s1 <= c->lb_s1) return 0;
  CK(c->lbmat.reserve((size_t) (s1 - s0) * c->g_tiles));
  const size_t d = c->d;
  float* tcen = c->thdr.p;
  float* tlo = tcen + d * c->g_tiles;
  float* thi = tlo + d * c->g_tiles;
  float* trad = thi + d * c->g_tiles;
  CK(launch_tile_lb(tcen, tlo, thi, trad, (int) d, c->g_tiles, s0, s1, c->lbmat.p, c->stream));
  c->launches += 1;
  c->lb_s0 = s0;
  c->lb_s1 = s1;
  return 0;
}

static int fill_ggeom(dcb200_ctx* c, size_t row_begin, size_t row_end, GemmGeom* g, int* grid) {
  memset(g, 0, sizeof(*g));
  if (row_begin % GT) return fail("dcb200: the GEMM-form path needs position ranges that start at a multiple of 128");
  CKI(ensure_tile_lb(c, row_begin, row_end));
  g->gT = c->gT.p;
  g->gnorm = c->gnorm.p;
  g->xR = c->xR.p;
  g->lb = c->lbmat.p + (size_t) (row_begin / GT - c->lb_s0) * c->g_tiles;
  g->d = (int) c->d;
  g->kc = c->g_kc;
  g->k8 = c->g_k8;
  int dev_smem = 0;
  CK(cudaDeviceGetAttribute(&dev_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, c->device));
  // two row tiles per work item when both operand images fit next to a ring of at least four chunks (n_cols <= 128) and the
  // range starts at a multiple of 256: every streamed column chunk then feeds 256 rows
  const char* ra_env = getenv("DCB200_GEMM_RA");
  g->ra = (row_begin % (2 * GT) == 0 && gemm_smem_bytes(2 * g->kc, 4) <= (size_t) dev_smem && !(ra_env && ra_env[0] == '1')) ? 2 : 1;
  int avail = 8;
  while (avail > 2 && gemm_smem_bytes(g->ra * g->kc, avail) > (size_t) dev_smem) --avail;
  if (gemm_smem_bytes(g->ra * g->kc, avail) > (size_t) dev_smem) return fail("dcb200: GEMM-form kernel does not fit in shared memory");
  // ring slots are filled and freed in groups of cb chunks (one mbarrier wait, one fence and one tcgen05.commit per group in the
  // single MMA-issuing warp, whose serial instruction stream is what limits the scan); cb divides kc so that a group never
  // straddles two tiles, and the ring holds at least two groups
  g->cb = 1;
  for (int cand = 4; cand >= 2; cand >>= 1)
    if (g->kc % cand == 0 && avail / cand >= 2) { g->cb = cand; break; }
  { const char* e = getenv("DCB200_GEMM_CB"); const int v = e ? atoi(e) : 0; if ((v == 1 || v == 2 || v == 4) && g->kc % v == 0 && avail / v >= 2) g->cb = v; }
  g->cb_log2 = g->cb == 4 ? 2 : g->cb == 2 ? 1 : 0;
  g->n_stages = avail / g->cb * g->cb;
  g->n = (uint32_t) c->n;
  g->n_tiles = (uint32_t) c->g_tiles;
  g->row_begin = (uint32_t) row_begin;
  g->row_end = (uint32_t) row_end;
  g->n_row_tiles128 = (uint32_t) ((row_end - row_begin + GT - 1) / GT);
  g->n_row_tiles = (g->n_row_tiles128 + g->ra - 1) / g->ra;
  // about 16 column items per row block (load balance); the row tiles' operand images are loaded once per item
  g->tiles_per_item = std::max(1u, std::min(g->n_tiles, std::max(32u, (g->n_tiles + 15) / 16)));
  g->n_col_items = (g->n_tiles + g->tiles_per_item - 1) / g->tiles_per_item;
  if ((uint64_t) g->n_row_tiles * g->n_col_items >= 0x7fffffffull) return fail("too many work items");
  *grid = (int) std::min<uint64_t>((uint64_t) c->sm_count, (uint64_t) g->n_row_tiles * g->n_col_items);
  g->work_counter = c->scalars;
  g->stats = c->stats;
  // error model of the fast value (gemm_kernels.cuh): |dx| <= 2^-11 |x - centre| per operand (round to nearest TF32; the
  // subtraction of the centre adds 2^-24), accumulation: norms are FMA chains over K terms, the tensor core adds 9 terms per
  // K=8 step with at worst truncation to 2^-23 (factor 2 on top), a few FP32 roundings at the end; safety factor 1.5
  float c_loc, prune_slack;
  error_bounds(c->d, c->maxnorm2, &c_loc, &g->e_rel, &prune_slack);
  const double u = ldexp(1.0, -24);
  g->rho_c = up(ldexp(1.0, -11) * (1.0 + ldexp(1.0, -8)));
  g->c_acc = up(1.5 * (((double) c->g_kc * GK + 32.0) * u + (double) c->g_k8 * 9.0 * ldexp(1.0, -22)));
  g->nymax = c->g_nymax;
  g->prune_slack = prune_slack;
  g->prune_thr = INFINITY;
  {
    const char* e = getenv("DCB200_GEMM_PROF");
    if (e && e[0] == '1') {
      if (!c->gprof) CK(cudaMalloc(&c->gprof, 16 * sizeof(unsigned long long)));
      CK(cudaMemsetAsync(c->gprof, 0, 16 * sizeof(unsigned long long), c->stream));
      g->prof = c->gprof;
    }
  }
  c->pairs_scheduled += (uint64_t) (row_end - row_begin) * (uint64_t) c->n;
  return 0;
}

static int build_layout(dcb200_ctx* c, const float* dev_coords, size_t n, size_t d, bool keep_order) {
  if (n == 0 || d == 0) return fail("dcb200: empty coordinate array");
  if (n >= 0x7ffffff0ull) return fail("dcb200: more than 2^31-16 frames are not supported");
  const size_t ld = (n + LD_ALIGN - 1) / LD_ALIGN * LD_ALIGN;
  c->n = n; c->d = d; c->ld = ld;
  c->nn_ready = false;
  c->layout_gen += 1;
  c->spatial = !keep_order;
  CK(c->xT.reserve(d * ld));
  CK(c->cT.reserve((d + 1) * ld + ld / (d <= (size_t) MAX_TEMPLATE_D ? TileW<1>::tj : TileW<0>::tj) * ((3 * d + 1 + 3) / 4 * 4)));
  CK(c->bbox.reserve(ld / 64 * 2 * d));
  const int tj = d <= (size_t) MAX_TEMPLATE_D ? TileW<1>::tj : TileW<0>::tj;
  const int dp = (int) ((3 * d + 1 + 3) / 4 * 4);
  CK(c->centre.reserve(2 * d));
  CK(c->perm.reserve(n));
  CK(cudaMemsetAsync(c->scalars + 1, 0, sizeof(unsigned int), c->stream));
  centre_kernel<<<(unsigned int) d, 256, 0, c->stream>>>(dev_coords, n, (int) d, c->centre.p, c->centre.p + d);
  c->launches += 1;
  if (c->spatial) {
    // dims that enter the key: all of them up to 8 (60 key bits: 7 bits each at 8 dims).  Measured at 1M x 10: ordering by
    // 8 dims instead of 6 cuts the evaluated share of the pair matrix from 16.1 % to 13.5 % (populations 212 -> 189 ms);
    // 10 dims gain nothing more and cost the neighbour search 8 %
    const int m = (int) std::min<size_t>(d, (size_t) std::min(ORDER_MAX_DIMS, env_int("DCB200_ORDER_DIMS", 8)));
    const int bits = std::min(16, 60 / m);        // cells far smaller than a 128-frame tile even in the densest regions
    CK(c->skeys_a.reserve(n)); CK(c->skeys_b.reserve(n)); CK(c->iota.reserve(n));
    spatial_keys_kernel<<<blocks_for(n, 256), 256, 0, c->stream>>>(dev_coords, n, (int) d, c->centre.p, c->centre.p + d, m, bits,
                                                                   c->skeys_a.p, c->iota.p);
    c->launches += 1;
    CKI(sort_pairs_u64(c, c->skeys_a.p, c->skeys_b.p, c->iota.p, c->perm.p, n, m * bits));
  } else {
    iota_kernel<<<blocks_for(n, 256), 256, 0, c->stream>>>(c->perm.p, n);
    c->launches += 1;
  }
  pack_tiles_kernel<<<(unsigned int) (ld / tj), tj, 0, c->stream>>>(dev_coords, n, (int) d, ld, dp, c->centre.p,
                                                                    c->spatial ? c->perm.p : nullptr, c->xT.p, c->cT.p, c->bbox.p,
                                                                    c->scalars + 1);
  {
    const uint32_t n_super = (uint32_t) ((ld + SUPER_FRAMES - 1) / SUPER_FRAMES);
    CK(c->sbbox.reserve((size_t) n_super * 2 * d));
    super_bbox_kernel<<<blocks_for(n_super, 8), 256, 0, c->stream>>>(c->bbox.p, (int) d, (uint32_t) (ld / 64), n_super, c->sbbox.p);
  }
  c->launches += 2;
  CK(cudaGetLastError());
  c->gemm = gemm_eligible(n, d, c->spatial);
  c->lb_s0 = c->lb_s1 = 0;
  if (c->gemm) {
    // operand images for the tensor-core path, norms of the rounded values, row-major exact copy, tile geometry
    c->g_kc = (int) ((d + GK - 1) / GK);
    c->g_k8 = (int) ((d + 7) / 8);
    c->g_tiles = ld / GT;
    CK(c->gT.reserve(c->g_tiles * (size_t) c->g_kc * G_CHUNK_FLOATS));
    CK(c->gnorm.reserve(ld));
    CK(c->xR.reserve(ld * d));
    CK(c->thdr.reserve((3 * d + 1) * c->g_tiles));
    float* tcen = c->thdr.p;
    float* tlo = tcen + d * c->g_tiles;
    float* thi = tlo + d * c->g_tiles;
    float* trad = thi + d * c->g_tiles;
    CK(launch_gpack(dev_coords, n, (int) d, c->g_kc, c->g_tiles, c->centre.p, c->perm.p, c->gT.p, c->gnorm.p, c->xR.p, tcen, tlo, thi, trad,
                    c->stream));
    CK(cudaMemsetAsync(c->scalars + 3, 0, sizeof(unsigned int), c->stream));
    max_u32_kernel<<<std::min(blocks_for(n, 256), 1024u), 256, 0, c->stream>>>(reinterpret_cast<const uint32_t*>(c->gnorm.p), n,
                                                                            c->scalars + 3);
    c->launches += 2;
    CK(cudaGetLastError());
  }
  unsigned int bits[3] = {0, 0, 0};
  CK(cudaMemcpyAsync(bits, c->scalars + 1, sizeof(bits), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  memcpy(&c->maxnorm2, &bits[0], 4);
  if (!(c->maxnorm2 == c->maxnorm2) || c->maxnorm2 > FLT_MAX) return fail("dcb200: coordinates contain NaN or infinite values");
  if (c->gemm) {
    memcpy(&c->g_nymax, &bits[2], 4);
    c->g_nymax = nextafterf(c->g_nymax, INFINITY);
  }
  return 0;
}

extern "C" int dcb200_ctx_set_coords_ex(dcb200_ctx* c, const float* coords, size_t n, size_t d, int on_device, int keep_order) {
  if (!c || !coords) return fail("null argument");
  CK(cudaSetDevice(c->device));
  if (on_device) return build_layout(c, coords, n, d, keep_order != 0);
  CK(c->stage.reserve(n * d));
  CK(cudaMemcpyAsync(c->stage.p, coords, n * d * sizeof(float), cudaMemcpyHostToDevice, c->stream));
  return build_layout(c, c->stage.p, n, d, keep_order != 0);
}
extern "C" int dcb200_ctx_set_coords_device(dcb200_ctx* c, const float* dev_coords, size_t n, size_t d) {
  return dcb200_ctx_set_coords_ex(c, dev_coords, n, d, 1, 0);
}
extern "C" int dcb200_ctx_set_coords(dcb200_ctx* c, const float* host_coords, size_t n, size_t d) {
  return dcb200_ctx_set_coords_ex(c, host_coords, n, d, 0, 0);
}

extern "C" int dcb200_ctx_order(dcb200_ctx* c, uint32_t* dev_perm) {
  if (!c || !dev_perm) return fail("null argument");
  if (c->n == 0) return fail("dcb200_ctx_order: no coordinates set");
  CK(cudaSetDevice(c->device));
  CK(cudaMemcpyAsync(dev_perm, c->perm.p, c->n * sizeof(uint32_t), cudaMemcpyDeviceToDevice, c->stream));
  return 0;
}

extern "C" int dcb200_ctx_to_frame_order(dcb200_ctx* c, const uint32_t* dev_src, size_t n_arrays, uint32_t* dev_dst) {
  if (!c || !dev_src || !dev_dst) return fail("null argument");
  if (c->n == 0) return fail("dcb200_ctx_to_frame_order: no coordinates set");
  if (n_arrays == 0) return 0;
  CK(cudaSetDevice(c->device));
  to_frame_order_kernel<<<blocks_for(c->n, 256), 256, 0, c->stream>>>(dev_src, dev_dst, c->perm.p, c->n, n_arrays);
  c->launches += 1;
  CK(cudaGetLastError());
  return 0;
}

static int dump_gprof(dcb200_ctx* c, const char* what, int grid) {
  if (!c->gprof || !getenv("DCB200_GEMM_PROF")) return 0;
  unsigned long long h[16];
  CK(cudaMemcpyAsync(h, c->gprof, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  static const char* names[12] = {"prod side_empty wait", "prod ring empty wait", "prod a_empty wait", "prod total", "mma side_full wait",
                                  "mma a_full wait", "mma tmem_empty wait", "mma ring full wait", "mma total", "epi side_full wait",
                                  "epi tmem_full wait", "epi total"};
  fprintf(stderr, "[dcb200] %s: cycles per CTA (grid %d)\n", what, grid);
  for (int q = 0; q < 12; ++q) fprintf(stderr, "[dcb200]   %-22s %12.0f\n", names[q], (double) h[q] / grid / (q >= 9 ? 2 : 1));
  return 0;
}


// Cell table of pops_bin_kernel (kernels.cuh) for the ascending, distinct squared radii b2[0..nb): K cells over
// s in [0, span], cell(s) = floor(sat(s / span) (K - 1)) up to a quarter-cell rounding.  A pair in cell i has
// s in [i - 0.15, i + 1.15] w (w = span / (K - 1)) and is decided by the table only if its error band is narrower than
// margin = w / 4, so the only radii it can be near lie in [i - 0.4, i + 1.4] w: the table exists iff no such range
// holds two radii.  Entry: that radius with its index k in the low 5 mantissa bits (bin = k + (s >= entry)), else a
// sentinel (+big | 0 if no radius lies below the cell, -big | (k - 1) if k do).  Returns K, 0 if there is no table
// (radii too close together for 4096 cells, or r_max = 0): the histogram kernel serves such lists.
static int build_bin_table(const float* b2, int nb, std::vector<float>* table, float* scale, float* margin) {
  if (nb < 1 || nb > MAX_BINS || !(b2[nb - 1] > 1e-30f) || !(b2[nb - 1] < 1e30f)) return 0;
  const double span = (double) b2[nb - 1] * 1.02;
  for (int K = 256; K <= 4096; K *= 2) {
    const double w = span / (double) (K - 1);
    table->assign((size_t) K, 0.f);
    bool ok = true;
    int below = 0;                                   // radii below the current cell's range
    for (int i = 0; i < K && ok; ++i) {
      const double lo = ((double) i - 0.4) * w, hi = ((double) i + 1.4) * w;
      while (below < nb && (double) b2[below] < lo) ++below;
      int inside = 0;
      while (below + inside < nb && (double) b2[below + inside] <= hi) ++inside;
      uint32_t bits;
      if (inside > 1) {
        ok = false;
        break;
      } else if (inside == 1) {
        memcpy(&bits, &b2[below], 4);
        bits = (bits & ~31u) | (uint32_t) below;
      } else if (below == 0) {
        bits = 0x7f7fffe0u;                          // +big: s >= entry never holds, bin 0
      } else {
        bits = 0xff7fffe0u | (uint32_t) (below - 1);  // -big: s >= entry always holds, bin (below - 1) + 1
      }
      memcpy(&(*table)[i], &bits, 4);
    }
    if (ok) {
      *scale = (float) (1.0 / span);
      *margin = (float) (0.25 * w);
      return K;
    }
  }
  return 0;
}

// ---- populations ----------------------------------------------------------------------------
// GEMM-form passes (tensor cores): up to eight distinct radii per pass, largest radii first, each pass pruned by its own r_max
static int gemm_populations(dcb200_ctx* c, const float* radii, size_t n_radii, const std::vector<float>& rad2, const std::vector<float>& uniq,
                            size_t row_begin, size_t row_end, size_t ld_cnt, uint32_t* dev_pops) {
  const size_t rows = row_end - row_begin, ld_out = rows;
  const char* chk = getenv("DCB200_GEMM_CHECK");
  const bool check = chk && chk[0] == '1' && uniq.size() == 1;
  if (check && !c->gcheck) {
    CK(cudaMalloc(&c->gcheck, 2 * sizeof(float)));
    CK(cudaMemsetAsync(c->gcheck, 0, 2 * sizeof(float), c->stream));
  }
  size_t hi = uniq.size();
  while (hi > 0) {
    const size_t n_pass = std::min<size_t>(8, hi);
    const int nb = n_pass <= 1 ? 1 : n_pass <= 2 ? 2 : n_pass <= 4 ? 4 : 8;
    const size_t b0 = hi - n_pass;
    GPopsArgs a;
    int grid = 0;
    CKI(fill_ggeom(c, row_begin, row_end, &a.g, &grid));
    a.n_bins = nb;
    for (int q = 0; q < 8; ++q) {
      a.rad2[q] = q < (int) n_pass ? uniq[b0 + q] : -1.f;
      a.rad[q] = q < (int) n_pass ? up(sqrt((double) uniq[b0 + q]) * (1.0 + 1e-7)) : 0.f;
    }
    // a tile pair whose lower bound exceeds this cannot contain a pair with exact d2 < r_max^2
    const double rmax2 = (double) uniq[hi - 1];
    a.g.prune_thr = up(rmax2 * (1.0 + 1.01 * (double) a.g.e_rel) * 1.0001 + 2.0 * (double) a.g.prune_slack);
    a.g.check = check ? c->gcheck : nullptr;
    CK(c->cnt.reserve((size_t) nb * ld_cnt));
    a.cnt = c->cnt.p;
    a.ld_cnt = ld_cnt;
    CK(cudaMemsetAsync(c->cnt.p, 0, (size_t) nb * ld_cnt * sizeof(uint32_t), c->stream));
    CK(cudaMemsetAsync(c->scalars, 0, sizeof(unsigned int), c->stream));
    CK(launch_gpops(a, grid, check, c->stream));
    c->launches += 1;
    CKI(dump_gprof(c, "gscan_pops", grid));
    FinalizeArgs f;
    f.n_out = 0;
    f.cumulative = 1;
    auto flush = [&]() -> int {
      if (f.n_out == 0) return 0;
      pops_finalize_kernel<<<blocks_for(rows, 256), 256, 0, c->stream>>>(c->cnt.p, ld_cnt, rows, ld_out, dev_pops, f);
      c->launches += 1;
      f.n_out = 0;
      CK(cudaGetLastError());
      return 0;
    };
    for (size_t r = 0; r < n_radii; ++r) {
      const size_t bin = std::lower_bound(uniq.begin(), uniq.end(), rad2[r]) - uniq.begin();
      if (bin < b0 || bin >= hi) continue;
      uint32_t mult = 0;
      for (size_t q = 0; q < n_radii; ++q) mult += (radii[q] == radii[r]);
      f.out_row[f.n_out] = (int) r;
      f.bin[f.n_out] = (int) (bin - b0);
      f.mult[f.n_out] = mult;
      f.self[f.n_out] = rad2[r] > 0.f ? 1u : 0u;      // d2(i,i) = 0 < r^2
      if (++f.n_out == MAX_BINS * 4) CKI(flush());
    }
    CKI(flush());
    hi = b0;
  }
  return 0;
}

// rows: blocks from row_begin with stride rb_stride (fill_geom); dev_pops: [n_radii][ld_out], the launch's rows in out_index order
static int populations_impl(dcb200_ctx* c, const float* radii, size_t n_radii, size_t row_begin, size_t row_end, size_t rb_stride,
                            size_t ld_out, uint32_t* dev_pops) {
  if (!c || (!radii && n_radii) || !dev_pops) return fail("null argument");
  if (c->n == 0) return fail("dcb200_ctx_populations: no coordinates set");
  if (row_begin > row_end || row_end > c->n) return fail("dcb200_ctx_populations: bad position range");
  if (n_radii == 0 || row_begin == row_end) return 0;
  CK(cudaSetDevice(c->device));
  const size_t rows = launch_rows(row_begin, row_end, rb_stride);
  // squared radii exactly as the reference forms them (float multiply, density_clustering.cpp:139)
  std::vector<float> rad2(n_radii);
  for (size_t r = 0; r < n_radii; ++r) rad2[r] = radii[r] * radii[r];
  std::vector<float> uniq(rad2);
  std::sort(uniq.begin(), uniq.end());
  uniq.erase(std::unique(uniq.begin(), uniq.end()), uniq.end());
  for (float v : uniq)
    if (!(v == v)) return fail("dcb200_ctx_populations: NaN radius");
  const size_t ld_cnt = launch_row_blocks(row_begin, row_end, rb_stride) * ROWS_PER_CTA;
  if (c->gemm && row_begin % GT == 0 && rb_stride == 1) {
    if (ld_out != rows) return fail("dcb200: the GEMM-form path writes compact outputs");
    return gemm_populations(c, radii, n_radii, rad2, uniq, row_begin, row_end, ld_cnt, dev_pops);
  }
  const int tj = tile_width(c->d);
  // Three kernels for the specialised dims (kernels.cuh); any other shape runs the histogram kernel:
  //   count  up to 8 (D <= 6) or 4 distinct radii per pass, branch-free sign-bit counting, ~3 instructions per pair and radius;
  //   bin    up to 31 radii per pass through a cell table, ~14 instructions per pair of a step that holds candidates,
  //          whatever the number of radii (needs row groups that coincide with tiles and radii a table can separate);
  //   hist   up to 31 radii per pass, per-hit handler with a binary search (the general fallback).
  // Passes take the largest radii first so that each is pruned by its own r_max.  DCB200_POPS_MODE=count|bin|hist forces one.
  enum Mode { COUNT, BIN, HIST };
  const size_t count_pass = c->d <= 6 ? 8 : 4;
  const bool specialised = c->d <= (size_t) MAX_TEMPLATE_D;
  Mode mode = HIST;
  if (specialised && uniq.size() <= 2 * count_pass) mode = COUNT;
  if (specialised && uniq.size() >= (size_t) env_int("DCB200_BIN_MIN_RADII", 4) && row_begin % (32 * RI) == 0) mode = BIN;
  if (const char* e = getenv("DCB200_POPS_MODE")) {
    if (!strcmp(e, "count") && specialised && uniq.size() <= 2 * count_pass) mode = COUNT;
    if (!strcmp(e, "bin") && specialised && row_begin % (32 * RI) == 0) mode = BIN;
    if (!strcmp(e, "hist")) mode = HIST;
  }
  size_t hi = uniq.size();                        // radii [b0, hi) of uniq go into the next pass
  while (hi > 0) {
    Mode pm = mode;
    size_t n_pass = std::min(pm == COUNT ? count_pass : (size_t) MAX_BINS, hi);
    std::vector<float> table;
    float lut_scale = 0.f, lut_margin = 0.f;
    int lut_k = 0;
    if (pm == BIN) {
      lut_k = build_bin_table(uniq.data() + (hi - n_pass), (int) n_pass, &table, &lut_scale, &lut_margin);
      if (lut_k == 0 || occ_pops_bin((int) c->d, (int) n_pass, lut_k) < 1) pm = HIST;
    }
    const bool count_mode = pm == COUNT;
    int nb = (int) n_pass;                        // kernel variant: the smallest instantiated count >= n_pass
    if (count_mode) {
      if (n_pass == 5) nb = 6;
      if (n_pass == 7) nb = 8;
    }
    const size_t b0 = hi - n_pass;
    PopsArgs a;
    int grid = 0;
    const int occ = pm == COUNT ? occ_pops_count((int) c->d, nb) : pm == BIN ? occ_pops_bin((int) c->d, nb, lut_k) : occ_pops((int) c->d, nb);
    CKI(fill_geom(c, row_begin, row_end, rb_stride, tj, occ, nb > 4 ? 64u : 32u, &a.g, &grid,
                  count_mode ? 0xffffffffu : (uint32_t) (57344 / tj)));
    a.n_bins = nb;
    a.lut = nullptr;
    a.lut_k = 0;
    a.lut_scale = a.lut_margin = a.band_max = 0.f;
    a.dense_lanes = 0;
    a.steal = 0;
    a.proj_prune = 0;
    for (int q = 0; q < 8; ++q) a.band[q] = 0.f;
    if (count_mode) {
      // radius-dependent part of the band: e_rel * r^2 (fast path) + roundings of s = acc + |x'|^2 and of s - r^2
      const double u = ldexp(1.0, -24);
      for (size_t q = 0; q < n_pass; ++q) {
        const double r2 = (double) uniq[b0 + q];
        a.band[q] = up((1.02 * (double) a.g.e_rel + 4.0 * u) * r2 + 1e-37);
      }
    }
    // unused slots: count mode -1 (nothing is ever inside, far outside every band), histogram / bin mode +inf
    for (int q = 0; q < 32; ++q) a.rad2[q] = q < (int) n_pass ? uniq[b0 + q] : (count_mode ? -1.f : INFINITY);
    const double rmax2 = (double) uniq[hi - 1];
    a.thr_fast = up(rmax2 * (1.0 + 1.01 * (double) a.g.e_rel));
    // a tile whose bounding-box distance exceeds this cannot contain a pair with exact d2 < r_max^2
    a.g.prune_thr = up((double) a.thr_fast * 1.0001 + 2.0 * (double) a.g.prune_slack);
    if (pm == BIN) {
      CK(c->lut.reserve(4096));
      CK(cudaMemcpyAsync(c->lut.p, table.data(), (size_t) lut_k * sizeof(float), cudaMemcpyHostToDevice, c->stream));
      a.lut = c->lut.p;
      a.lut_k = lut_k;
      a.lut_scale = lut_scale;
      a.lut_margin = lut_margin;
      // radius part of the band for r_max (e_rel r^2 + roundings of s and of s - e, as in count mode) + the 31 ulp by which
      // the table's entries may differ from the radii they stand for
      a.band_max = up((1.02 * (double) a.g.e_rel + 4.0 * ldexp(1.0, -24) + 32.0 * ldexp(1.0, -23)) * rmax2 + 1e-37);
      // a step is binned branch-free when this many lanes hold a candidate (C3: 4 / 8 / 16 lanes 126.5 / 127.5 / 131.5 ms with the
      // 2 x 8 step shape of D >= 9; flat between 4 and 16 with the 4 x 4 shape)
      a.dense_lanes = env_int("DCB200_BIN_DENSE_LANES", c->d >= 9 ? 4 : 8);
      a.steal = env_int("DCB200_BIN_STEAL", 1) == 1 ? 1 : 0;
      a.proj_prune = env_int("DCB200_BIN_PROJ", 1) == 1 ? 1 : 0;
    }
    const bool counts_self = pm != HIST;           // count and bin mode count the frame itself through d2 = 0 < r^2
    CK(c->cnt.reserve((size_t) nb * ld_cnt));
    a.cnt = c->cnt.p;
    a.ld_cnt = ld_cnt;
    CK(cudaMemsetAsync(c->cnt.p, 0, (size_t) nb * ld_cnt * sizeof(uint32_t), c->stream));
    CK(cudaMemsetAsync(c->scalars, 0, sizeof(unsigned int), c->stream));
    CK(pm == COUNT ? launch_pops_count((int) c->d, a, grid, c->stream)
                   : pm == BIN ? launch_pops_bin((int) c->d, a, grid, c->stream) : launch_pops((int) c->d, a, grid, c->stream));
    c->launches += 1;
    // input radii served by this pass, in groups the finalize kernel's argument block can hold
    FinalizeArgs f;
    f.n_out = 0;
    f.cumulative = count_mode ? 1 : 0;
    auto flush = [&]() -> int {
      if (f.n_out == 0) return 0;
      pops_finalize_kernel<<<blocks_for(rows, 256), 256, 0, c->stream>>>(c->cnt.p, ld_cnt, rows, ld_out, dev_pops, f);
      c->launches += 1;
      f.n_out = 0;
      CK(cudaGetLastError());
      return 0;
    };
    for (size_t r = 0; r < n_radii; ++r) {
      const size_t bin = std::lower_bound(uniq.begin(), uniq.end(), rad2[r]) - uniq.begin();
      if (bin < b0 || bin >= hi) continue;
      uint32_t mult = 0;
      for (size_t q = 0; q < n_radii; ++q) mult += (radii[q] == radii[r]);
      f.out_row[f.n_out] = (int) r;
      f.bin[f.n_out] = (int) (bin - b0);
      f.mult[f.n_out] = mult;
      f.self[f.n_out] = (counts_self && rad2[r] > 0.f) ? 1u : 0u;      // d2(i,i) = 0 < r^2
      if (++f.n_out == MAX_BINS * 4) CKI(flush());
    }
    CKI(flush());
    hi = b0;
  }
  return 0;
}

extern "C" int dcb200_ctx_populations(dcb200_ctx* c, const float* radii, size_t n_radii, size_t row_begin, size_t row_end,
                                      uint32_t* dev_pops) {
  if (row_begin > row_end) return fail("dcb200_ctx_populations: bad position range");
  return populations_impl(c, radii, n_radii, row_begin, row_end, 1, row_end - row_begin, dev_pops);
}

// ---- shards -------------------------------------------------------------------------------------
// The positions are dealt to n_shards shards in blocks of ROWS_PER_CTA: block b belongs to shard b % n_shards and a shard
// keeps its blocks in order.  Block-cyclic instead of contiguous: the spatial order puts dense and sparse regions into long
// runs, so contiguous shards of equal length cost very different amounts (SCALE_r01: efficiency 0.43 at 8 GPUs).
// Contexts on the GEMM-form path (17 <= n_cols <= 256) keep contiguous shards of `capacity` positions (their work items are
// 256-row pairs of tiles); dcb200_ctx_shards_to_frame_order / dcb200_ctx_nn_finish_shards undo whichever dealing was used.
extern "C" size_t dcb200_shard_capacity(size_t n_rows, int n_shards) {
  if (n_shards < 1) return 0;
  const size_t blocks = (n_rows + ROWS_PER_CTA - 1) / ROWS_PER_CTA;
  return (blocks + n_shards - 1) / n_shards * ROWS_PER_CTA;
}
static bool shard_ok(const dcb200_ctx* c, int shard, int n_shards) { return c && n_shards >= 1 && shard >= 0 && shard < n_shards; }
// first row, end row and block stride of a shard's launch
static void shard_range(const dcb200_ctx* c, int shard, int n_shards, size_t* b, size_t* e, size_t* stride) {
  if (c->gemm) {
    const size_t cap = dcb200_shard_capacity(c->n, n_shards);
    *b = std::min(c->n, cap * shard);
    *e = std::min(c->n, cap * (shard + 1));
    *stride = 1;
  } else {
    *b = std::min(c->n, (size_t) shard * ROWS_PER_CTA);
    *e = *b < c->n ? c->n : *b;
    *stride = (size_t) n_shards;
  }
}
extern "C" int dcb200_ctx_shard_rows(dcb200_ctx* c, int shard, int n_shards, size_t* rows) {
  if (!shard_ok(c, shard, n_shards) || !rows) return fail("dcb200_ctx_shard_rows: bad arguments");
  size_t b, e, stride;
  shard_range(c, shard, n_shards, &b, &e, &stride);
  *rows = launch_rows(b, e, stride);
  return 0;
}
extern "C" int dcb200_ctx_populations_shard(dcb200_ctx* c, const float* radii, size_t n_radii, int shard, int n_shards,
                                            uint32_t* dev_pops) {
  if (!shard_ok(c, shard, n_shards)) return fail("dcb200_ctx_populations_shard: bad shard");
  if (c->n == 0) return fail("dcb200_ctx_populations_shard: no coordinates set");
  size_t b, e, stride;
  shard_range(c, shard, n_shards, &b, &e, &stride);
  const size_t cap = dcb200_shard_capacity(c->n, n_shards);
  if (stride == 1 && c->gemm) {
    // the GEMM-form kernels write compact rows: run into scratch, then copy row by row into the padded layout
    const size_t rows = e - b;
    if (rows == 0 || n_radii == 0) return 0;
    CK(cudaSetDevice(c->device));
    // (a buffer of its own: the caller's dev_pops may live in the context's io buffers)
    CK(c->shard_tmp.reserve(n_radii * rows));
    CKI(populations_impl(c, radii, n_radii, b, e, 1, rows, c->shard_tmp.p));
    CK(cudaMemcpy2DAsync(dev_pops, cap * sizeof(uint32_t), c->shard_tmp.p, rows * sizeof(uint32_t), rows * sizeof(uint32_t), n_radii,
                         cudaMemcpyDeviceToDevice, c->stream));
    return 0;
  }
  return populations_impl(c, radii, n_radii, b, e, stride, cap, dev_pops);
}

// dst[a][frame] = src[shard][a][local] for the gathered shard outputs src [n_shards][n_arrays][capacity]
__global__ void shards_to_frame_order_kernel(const uint32_t* __restrict__ src, uint32_t* __restrict__ dst, const uint32_t* __restrict__ perm,
                                             size_t n, size_t n_arrays, size_t cap, uint32_t n_shards, int cyclic) {
  const size_t p = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  size_t shard, local;
  if (cyclic) {
    const size_t blk = p / ROWS_PER_CTA;
    shard = blk % n_shards;
    local = blk / n_shards * ROWS_PER_CTA + p % ROWS_PER_CTA;
  } else {
    shard = p / cap;
    local = p % cap;
  }
  const size_t o = perm[p];
  for (size_t a = 0; a < n_arrays; ++a) dst[a * n + o] = src[(shard * n_arrays + a) * cap + local];
}
extern "C" int dcb200_ctx_shards_to_frame_order(dcb200_ctx* c, const uint32_t* dev_src, size_t n_arrays, int n_shards, uint32_t* dev_dst) {
  if (!c || !dev_src || !dev_dst || n_shards < 1) return fail("dcb200_ctx_shards_to_frame_order: bad arguments");
  if (c->n == 0) return fail("dcb200_ctx_shards_to_frame_order: no coordinates set");
  if (n_arrays == 0) return 0;
  CK(cudaSetDevice(c->device));
  shards_to_frame_order_kernel<<<blocks_for(c->n, 256), 256, 0, c->stream>>>(dev_src, dev_dst, c->perm.p, c->n, n_arrays,
                                                                             dcb200_shard_capacity(c->n, n_shards), (uint32_t) n_shards,
                                                                             c->gemm ? 0 : 1);
  c->launches += 1;
  CK(cudaGetLastError());
  return 0;
}

// ---- free energies ----------------------------------------------------------------------------
extern "C" int dcb200_ctx_free_energies(dcb200_ctx* c, const uint32_t* dev_pops, size_t n, uint32_t max_pop, float* dev_fe) {
  if (!c || !dev_pops || !dev_fe) return fail("null argument");
  if (n == 0) return 0;
  CK(cudaSetDevice(c->device));
  if (max_pop == 0) {
    CK(cudaMemsetAsync(c->scalars + 2, 0, sizeof(unsigned int), c->stream));
    max_u32_kernel<<<std::min(blocks_for(n, 256), 1024u), 256, 0, c->stream>>>(dev_pops, n, c->scalars + 2);
    c->launches += 1;
  }
  fe_kernel<<<blocks_for(n, 256), 256, 0, c->stream>>>(dev_pops, n, c->scalars + 2, max_pop, dev_fe);
  c->launches += 1;
  CK(cudaGetLastError());
  return 0;
}

// ---- nearest neighbours -----------------------------------------------------------------------
// dev_fe: free energies in FRAME order.  Builds lo[p] = #{frames with fe < fe[frame at position p]}.
extern "C" int dcb200_ctx_nn_prepare(dcb200_ctx* c, const float* dev_fe) {
  if (!c || !dev_fe) return fail("null argument");
  if (c->n == 0) return fail("dcb200_ctx_nn_prepare: no coordinates set");
  CK(cudaSetDevice(c->device));
  const size_t n = c->n;
  CK(c->keys_a.reserve(n)); CK(c->keys_b.reserve(n)); CK(c->iota.reserve(n)); CK(c->tmp_u32.reserve(n));
  CK(c->tmp2_u32.reserve(n)); CK(c->lo.reserve(n)); CK(c->lof.reserve(c->ld));
  c->lo_shift = 0;
  while ((n >> c->lo_shift) > (size_t(1) << 24)) ++c->lo_shift;
  fe_keys_kernel<<<blocks_for(n, 256), 256, 0, c->stream>>>(dev_fe, n, c->keys_a.p, c->iota.p);
  CKI(sort_pairs_u32(c, c->keys_a.p, c->keys_b.p, c->iota.p, c->tmp_u32.p, n, 32));              // tmp_u32: fe position -> frame
  lower_bound_kernel<<<blocks_for(n, 256), 256, 0, c->stream>>>(c->keys_b.p, n, c->keys_a.p);    // keys_a: lo by fe position
  scatter_by_kernel<<<blocks_for(n, 256), 256, 0, c->stream>>>(c->keys_a.p, c->tmp_u32.p, n, c->tmp2_u32.p);   // lo by frame
  lo_by_position_kernel<<<blocks_for(c->ld, 256), 256, 0, c->stream>>>(c->tmp2_u32.p, c->perm.p, n, c->ld, c->lo_shift, c->lo.p,
                                                                       c->lof.p);                                    // lo by position
  c->launches += 4;
  CK(cudaGetLastError());
  c->nn_ready = true;
  return 0;
}

// rows: blocks from pos_begin with stride rb_stride (fill_geom); keys in out_index order
static int nn_scan_impl(dcb200_ctx* c, size_t pos_begin, size_t pos_end, size_t rb_stride, uint64_t* dev_keys_nn, uint64_t* dev_keys_hd) {
  if (!c || !dev_keys_nn || !dev_keys_hd) return fail("null argument");
  if (!c->nn_ready) return fail("dcb200_ctx_nn_scan: call dcb200_ctx_nn_prepare first");
  if (pos_begin > pos_end || pos_end > c->n) return fail("dcb200_ctx_nn_scan: bad position range");
  if (pos_begin == pos_end) return 0;
  CK(cudaSetDevice(c->device));
  const size_t rows = launch_rows(pos_begin, pos_end, rb_stride);
  // "no neighbour" = (n_rows + 1, FLT_MAX)  (density_clustering.cpp:257-260)
  const float fmax = FLT_MAX;
  uint32_t fbits;
  memcpy(&fbits, &fmax, 4);
  const unsigned long long none = ((unsigned long long) fbits << 32) | (unsigned long long) (uint32_t) (c->n + 1);
  nn_seed_kernel<<<blocks_for(rows, 256), 256, 0, c->stream>>>(c->xT.p, c->ld, (int) c->d, (uint32_t) c->n, c->perm.p, c->lo.p,
                                                               (uint32_t) pos_begin, (uint32_t) pos_end, (uint32_t) rb_stride, (uint32_t) rows,
                                                               env_int("DCB200_NN_SEED_W", 8), none, (unsigned long long*) dev_keys_nn,
                                                               (unsigned long long*) dev_keys_hd);
  if (c->gemm && pos_begin % GT == 0 && rb_stride == 1) {
    // GEMM-form scan (tensor cores): window pass over every row tile's own neighbourhood, then all tiles
    GNnArgs ga;
    int ggrid = 0;
    CKI(fill_ggeom(c, pos_begin, pos_end, &ga.g, &ggrid));
    ga.perm = c->perm.p;
    ga.lo = c->lo.p;
    ga.lof = c->lof.p;
    ga.lo_bias = c->lo_shift ? 1.f : 0.f;
    CK(c->lomin.reserve(c->g_tiles));
    CK(launch_tile_min(c->lof.p, c->g_tiles, c->lomin.p, c->stream));
    ga.lomin = c->lomin.p;
    const uint32_t thr_tiles = ga.g.n_row_tiles * (uint32_t) ga.g.ra;      // 128-row tiles incl. the empty second tile of a last block
    CK(c->gthr.reserve((size_t) thr_tiles * 12));
    ga.thr_nn = c->gthr.p;
    ga.thr_hd = ga.thr_nn + (size_t) thr_tiles * 4;
    ga.lormax = ga.thr_hd + (size_t) thr_tiles * 4;
    ga.key_nn = (unsigned long long*) dev_keys_nn;
    ga.key_hd = (unsigned long long*) dev_keys_hd;
    CK(launch_gnn_tile_thr(ga.key_nn, ga.key_hd, c->lo.p, c->lof.p, ga.lo_bias, (uint32_t) pos_begin, (uint32_t) pos_end, thr_tiles,
                           ga.g.e_rel, ga.g.prune_slack, ga.thr_nn, ga.thr_hd, ga.lormax, c->stream));
    const uint32_t full_tpi = ga.g.tiles_per_item, full_items = ga.g.n_col_items;
    if (ga.g.n_tiles > 96) {
      ga.window = (uint32_t) env_int("DCB200_NN_WINDOW", 16);
      ga.g.tiles_per_item = ga.g.n_tiles;
      ga.g.n_col_items = 1;
      CK(cudaMemsetAsync(c->scalars, 0, sizeof(unsigned int), c->stream));
      CK(launch_gnn(ga, (int) std::min<uint64_t>((uint64_t) ggrid, ga.g.n_row_tiles), c->stream));
      c->launches += 1;
    }
    ga.window = 0;
    ga.g.tiles_per_item = full_tpi;
    ga.g.n_col_items = full_items;
    CK(cudaMemsetAsync(c->scalars, 0, sizeof(unsigned int), c->stream));
    if (ga.g.prof) CK(cudaMemsetAsync(c->gprof, 0, 16 * sizeof(unsigned long long), c->stream));      // counters of the full pass only
    CK(launch_gnn(ga, ggrid, c->stream));
    c->launches += 4;
    CKI(dump_gprof(c, "gscan_nn (full pass)", ggrid));
    return 0;
  }
  NnArgs a;
  int grid = 0;
  CKI(fill_geom(c, pos_begin, pos_end, rb_stride, tile_width(c->d), occ_nn((int) c->d), 32u, &a.g, &grid));
  a.perm = c->perm.p;
  a.lo = c->lo.p;
  a.lof = c->lof.p;
  a.lo_bias = c->lo_shift ? 1.f : 0.f;
  a.g.xrow = c->lof.p;
  CK(c->blk_thr.reserve((size_t) a.g.n_row_blocks * N_CONSUMER_WARPS));
  a.g.blk_thr = c->blk_thr.p;
  nn_block_thr_kernel<<<a.g.n_row_blocks, N_CONSUMERS, 0, c->stream>>>((const unsigned long long*) dev_keys_nn,
                                                                       (const unsigned long long*) dev_keys_hd, c->lo.p, (uint32_t) pos_begin,
                                                                       (uint32_t) pos_end, a.g.rb_stride, a.g.n_row_blocks, a.g.e_rel,
                                                                       a.g.prune_slack, c->blk_thr.p);
  a.key_nn = (unsigned long long*) dev_keys_nn;
  a.key_hd = (unsigned long long*) dev_keys_hd;
  // pass 1: every row block against its own neighbourhood in the spatial order (one item per block): the nearest
  // neighbours are almost always there, so pass 2 (all tiles, large items) starts from final-quality thresholds
  const uint32_t full_tpi = a.g.tiles_per_item, full_items = a.g.n_col_items;
  const int full_grid = grid;
  if (c->spatial && a.g.n_col_tiles > 96) {
    // 8 tiles to either side: with ~190 work items per CTA the full pass itself starts with each block's own column range, so
    // the first pass only has to settle the closest neighbours (C3: 16 -> 8 tiles 62.4 -> 61.0 ms, C2 12.7 -> 12.1; 32: 64.7 / 13.6)
    a.window = (uint32_t) std::max(1, env_int("DCB200_NN_WINDOW", 8));
    a.g.tiles_per_item = a.g.n_col_tiles;
    a.g.n_col_items = 1;
    grid = (int) std::min<uint64_t>((uint64_t) full_grid, a.g.n_row_blocks);
    CK(cudaMemsetAsync(c->scalars, 0, sizeof(unsigned int), c->stream));
    CK(launch_nn((int) c->d, a, grid, c->stream));
    c->launches += 1;
  }
  a.window = 0;
  a.g.tiles_per_item = full_tpi;
  a.g.n_col_items = full_items;
  CK(cudaMemsetAsync(c->scalars, 0, sizeof(unsigned int), c->stream));
  CK(launch_nn((int) c->d, a, grid = full_grid, c->stream));
  c->launches += 3;
  return 0;
}

extern "C" int dcb200_ctx_nn_scan(dcb200_ctx* c, size_t pos_begin, size_t pos_end, uint64_t* dev_keys_nn, uint64_t* dev_keys_hd) {
  return nn_scan_impl(c, pos_begin, pos_end, 1, dev_keys_nn, dev_keys_hd);
}

// keys of one shard (dcb200_ctx_populations_shard's dealing), [capacity] each; entries past the shard's rows are not written
extern "C" int dcb200_ctx_nn_scan_shard(dcb200_ctx* c, int shard, int n_shards, uint64_t* dev_keys_nn, uint64_t* dev_keys_hd) {
  if (!shard_ok(c, shard, n_shards)) return fail("dcb200_ctx_nn_scan_shard: bad shard");
  if (c->n == 0) return fail("dcb200_ctx_nn_scan_shard: no coordinates set");
  size_t b, e, stride;
  shard_range(c, shard, n_shards, &b, &e, &stride);
  return nn_scan_impl(c, b, e, stride, dev_keys_nn, dev_keys_hd);
}

__global__ void nn_finish_shards_kernel(const unsigned long long* __restrict__ knn, const unsigned long long* __restrict__ khd,
                                        const uint32_t* __restrict__ perm, size_t n, size_t cap, size_t shard_stride, uint32_t n_shards,
                                        int cyclic, uint32_t* __restrict__ nn_idx, float* __restrict__ nn_d2,
                                        uint32_t* __restrict__ hd_idx, float* __restrict__ hd_d2) {
  const size_t p = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  size_t shard, local;
  if (cyclic) {
    const size_t blk = p / ROWS_PER_CTA;
    shard = blk % n_shards;
    local = blk / n_shards * ROWS_PER_CTA + p % ROWS_PER_CTA;
  } else {
    shard = p / cap;
    local = p % cap;
  }
  const size_t o = perm[p];
  const unsigned long long a = knn[shard * shard_stride + local], b = khd[shard * shard_stride + local];
  nn_idx[o] = (uint32_t) a;
  nn_d2[o] = __uint_as_float((uint32_t) (a >> 32));
  hd_idx[o] = (uint32_t) b;
  hd_d2[o] = __uint_as_float((uint32_t) (b >> 32));
}
// gathered shard keys -> outputs in frame order; shard s's keys start at dev_keys_*[s * shard_stride] (capacity for two
// separately gathered arrays, 2 * capacity for one gathered [n_shards][2][capacity] block)
extern "C" int dcb200_ctx_nn_finish_shards(dcb200_ctx* c, const uint64_t* dev_keys_nn, const uint64_t* dev_keys_hd, int n_shards,
                                           size_t shard_stride, uint32_t* dev_nn_idx, float* dev_nn_d2, uint32_t* dev_hd_idx,
                                           float* dev_hd_d2) {
  if (!c || !dev_keys_nn || !dev_keys_hd || !dev_nn_idx || !dev_nn_d2 || !dev_hd_idx || !dev_hd_d2 || n_shards < 1 ||
      shard_stride < dcb200_shard_capacity(c ? c->n : 0, n_shards))
    return fail("dcb200_ctx_nn_finish_shards: bad arguments");
  if (c->n == 0) return fail("dcb200_ctx_nn_finish_shards: no coordinates set");
  CK(cudaSetDevice(c->device));
  nn_finish_shards_kernel<<<blocks_for(c->n, 256), 256, 0, c->stream>>>(
      (const unsigned long long*) dev_keys_nn, (const unsigned long long*) dev_keys_hd, c->perm.p, c->n,
      dcb200_shard_capacity(c->n, n_shards), shard_stride, (uint32_t) n_shards, c->gemm ? 0 : 1, dev_nn_idx, dev_nn_d2, dev_hd_idx,
      dev_hd_d2);
  c->launches += 1;
  CK(cudaGetLastError());
  return 0;
}

extern "C" int dcb200_ctx_nn_finish(dcb200_ctx* c, const uint64_t* dev_keys_nn, const uint64_t* dev_keys_hd, uint32_t* dev_nn_idx,
                                    float* dev_nn_d2, uint32_t* dev_hd_idx, float* dev_hd_d2) {
  if (!c || !dev_keys_nn || !dev_keys_hd || !dev_nn_idx || !dev_nn_d2 || !dev_hd_idx || !dev_hd_d2) return fail("null argument");
  if (c->n == 0) return fail("dcb200_ctx_nn_finish: no coordinates set");
  CK(cudaSetDevice(c->device));
  nn_finish_kernel<<<blocks_for(c->n, 256), 256, 0, c->stream>>>((const unsigned long long*) dev_keys_nn,
                                                                 (const unsigned long long*) dev_keys_hd, c->perm.p, c->n,
                                                                 dev_nn_idx, dev_nn_d2, dev_hd_idx, dev_hd_d2);
  c->launches += 1;
  CK(cudaGetLastError());
  return 0;
}

// ---- screening ----------------------------------------------------------------------------------
extern "C" int dcb200_ctx_screening_scan(dcb200_ctx* c, size_t m_prev, size_t m_new, size_t row_begin, size_t row_end,
                                         float max_dist2, uint32_t* dev_comp) {
  if (!c || !dev_comp) return fail("null argument");
  if (c->n == 0) return fail("dcb200_ctx_screening_scan: no coordinates set");
  if (c->spatial) return fail("dcb200_ctx_screening_scan: the coordinates must be set with keep_order (free-energy-sorted frames)");
  if (m_prev > m_new || m_new > c->n) return fail("dcb200_ctx_screening_scan: bad threshold positions");
  row_begin = std::max(row_begin, m_prev);
  row_end = std::min(row_end, m_new);
  if (row_begin >= row_end) return 0;
  CK(cudaSetDevice(c->device));
  ScreenArgs a;
  int grid = 0;
  CKI(fill_geom(c, row_begin, row_end, 1, tile_width(c->d), occ_screen((int) c->d), 32u, &a.g, &grid));
  a.cut = max_dist2;
  a.thr_fast = up((double) max_dist2 * (1.0 + 1.01 * (double) a.g.e_rel));
  a.g.prune_thr = up((double) a.thr_fast * 1.0001 + 2.0 * (double) a.g.prune_slack);
  a.parent = dev_comp;
  CK(cudaMemsetAsync(c->scalars, 0, sizeof(unsigned int), c->stream));
  CK(launch_screen((int) c->d, a, grid, c->stream));
  c->launches += 1;
  return 0;
}

extern "C" int dcb200_ctx_screening_flatten(dcb200_ctx* c, size_t m_new, uint32_t* dev_comp) {
  if (!c || !dev_comp) return fail("null argument");
  if (m_new == 0) return 0;
  CK(cudaSetDevice(c->device));
  uf_flatten_kernel<<<blocks_for(m_new, 256), 256, 0, c->stream>>>(dev_comp, m_new);
  c->launches += 1;
  CK(cudaGetLastError());
  return 0;
}

extern "C" int dcb200_ctx_screening_merge(dcb200_ctx* c, size_t m_new, uint32_t* dev_comp, const uint32_t* dev_other) {
  if (!c || !dev_comp || !dev_other) return fail("null argument");
  if (m_new == 0) return 0;
  CK(cudaSetDevice(c->device));
  uf_merge_kernel<<<blocks_for(m_new, 256), 256, 0, c->stream>>>(dev_comp, dev_other, m_new);
  c->launches += 1;
  CK(cudaGetLastError());
  return 0;
}

// Diagnostics of the GEMM-form path: active = 1 if the current coordinates are served by the tensor-core kernels;
// check_ratio = max observed |fast - exact| / (proven band) over all pairs of the populations calls made with
// DCB200_GEMM_CHECK=1 (single radius) since the context was created; must stay below 1.
extern "C" int dcb200_ctx_gemm_info(dcb200_ctx* c, int* active, float* check_ratio) {
  if (!c) return fail("null context");
  if (active) *active = c->gemm ? 1 : 0;
  if (check_ratio) {
    *check_ratio = 0.f;
    if (c->gcheck) {
      CK(cudaSetDevice(c->device));
      CK(cudaMemcpyAsync(check_ratio, c->gcheck, sizeof(float), cudaMemcpyDeviceToHost, c->stream));
      CK(cudaStreamSynchronize(c->stream));
    }
  }
  return 0;
}

// Diagnostics: throughput of tcgen05.mma kind::tf32 (128 x 128 x 8, operands resident in shared memory, random data) on all
// SMs, in TFLOP/s (2 flop per multiply-add), timed with CUDA events: the denominator of the tensor roofline of the GEMM-form scans.
extern "C" int dcb200_ctx_tf32_peak(dcb200_ctx* c, double ms_target, double* tflops) {
  if (!c || !tflops) return fail("null argument");
  CK(cudaSetDevice(c->device));
  long long* out = nullptr;
  CK(cudaMalloc(&out, (size_t) c->sm_count * sizeof(long long)));
  const int iters = 1 << 16;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  CK(launch_tf32_peak(c->sm_count, iters, out, c->stream));        // warm-up
  CK(cudaStreamSynchronize(c->stream));
  double best = 0.0, spent = 0.0;
  int launches = 1;
  while (spent < ms_target) {
    CK(cudaEventRecord(e0, c->stream));
    CK(launch_tf32_peak(c->sm_count, iters, out, c->stream));
    CK(cudaEventRecord(e1, c->stream));
    CK(cudaEventSynchronize(e1));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    spent += ms;
    launches += 1;
    const double fl = 2.0 * GT * GT * 8.0 * (double) iters * (double) c->sm_count;
    best = std::max(best, fl / (ms * 1e-3) / 1e12);
  }
  c->launches += launches;
  *tflops = best;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(out);
  return 0;
}

// Diagnostics: sustained FFMA throughput of the device (TFLOP/s, 2 flop per FFMA), measured with CUDA
// events over `ms_target` milliseconds of back-to-back launches.  Used by bench.py as the FP32 roofline peak.
extern "C" int dcb200_ctx_ffma_peak(dcb200_ctx* c, double ms_target, double* tflops) {
  if (!c || !tflops) return fail("null argument");
  CK(cudaSetDevice(c->device));
  float* out = nullptr;
  CK(cudaMalloc(&out, 4));
  const int iters = 8192, bs = 256, grid = c->sm_count * 8;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  ffma_peak_kernel<<<grid, bs, 0, c->stream>>>(out, iters, 1.0001f, 0.5f);     // warm-up
  CK(cudaStreamSynchronize(c->stream));
  double best = 0.0, spent = 0.0;
  int launches = 1;
  while (spent < ms_target) {
    CK(cudaEventRecord(e0, c->stream));
    for (int q = 0; q < 4; ++q) ffma_peak_kernel<<<grid, bs, 0, c->stream>>>(out, iters, 1.0001f, 0.5f);
    CK(cudaEventRecord(e1, c->stream));
    CK(cudaEventSynchronize(e1));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    spent += ms;
    launches += 4;
    const double fl = 4.0 * 2.0 * 16.0 * (double) iters * (double) grid * bs;
    best = std::max(best, fl / (ms * 1e-3) / 1e12);
  }
  c->launches += launches;
  *tflops = best;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(out);
  return 0;
}

// ------------------------------------------------------------------------------------------------
// (1) host-pointer entry points.  One GPU: everything stays on the device between upload and download.
//     Several GPUs (replaces the per-GPU upload / download / host merge of density_clustering_cuda.cu:152-181, :286-328,
//     :505-571): one worker thread per GPU, ONE host-to-device copy (to GPU 0) and an NCCL broadcast of the coordinates
//     over NVLink, block-cyclic row shards, ncclAllGather of the shard results, frame-order assembly on the device, ONE
//     download from GPU 0.  NCCL is bound at run time (nccl_dl.hpp); without it the shards are assembled through the host.
// ------------------------------------------------------------------------------------------------
#include <condition_variable>

#include "nccl_dl.hpp"

namespace {

std::mutex g_pool_mutex;
std::map<int, dcb200_ctx*> g_pool;     // one cached context per device (buffers are reused across calls)
// The pooled contexts (buffers, stream, work counters) are shared by all callers of the host-pointer entry points: one
// call at a time per process.  (The reference's path is not re-entrant either: SURVEY.md section 8b "Threading".)
std::mutex g_call_mutex;

int pooled_ctx(int device, dcb200_ctx** out) {
  std::lock_guard<std::mutex> lock(g_pool_mutex);
  auto it = g_pool.find(device);
  if (it != g_pool.end()) {
    *out = it->second;
    return 0;
  }
  CKI(dcb200_ctx_create(device, out));
  g_pool[device] = *out;
  return 0;
}

int gpus_to_use(int* n) {
  CKI(dcb200_device_count(n));
  if (*n == 0) return fail("dcb200: no CUDA device available (this library has no CPU fallback)");
  if (g_n_gpus > 0) *n = std::min(*n, g_n_gpus);
  return 0;
}

// runs fn(gpu, n_gpus) on one thread per GPU; returns the first failure
template <class Fn>
int on_gpus(int n_gpus, Fn&& fn) {
  if (n_gpus == 1) return fn(0, 1);
  std::vector<int> rc(n_gpus, 0);
  std::vector<std::string> msg(n_gpus);
  std::vector<std::thread> th;
  for (int g = 0; g < n_gpus; ++g)
    th.emplace_back([&, g]() {
      rc[g] = fn(g, n_gpus);
      if (rc[g]) msg[g] = g_err;
    });
  for (auto& t : th) t.join();
  for (int g = 0; g < n_gpus; ++g)
    if (rc[g]) return fail("GPU " + std::to_string(g) + ": " + msg[g]);
  return 0;
}

// The worker threads agree on a status before every collective: a thread that failed must not leave the others waiting
// inside NCCL.
struct Rendezvous {
  std::mutex m;
  std::condition_variable cv;
  int n, arrived = 0, gen = 0;
  bool ok = true, result = true;
  explicit Rendezvous(int n_) : n(n_) {}
  bool all_ok(bool mine) {
    std::unique_lock<std::mutex> lk(m);
    ok = ok && mine;
    if (++arrived == n) {
      result = ok;
      ok = true;
      arrived = 0;
      ++gen;
      cv.notify_all();
      return result;
    }
    const int my_gen = gen;
    cv.wait(lk, [&] { return gen != my_gen; });
    return result;
  }
};

// the devices the host-pointer entry points shard over, with one NCCL communicator each (created once per device count)
struct Gang {
  int n = 0;
  std::vector<dcb200_ctx*> ctx;
  std::vector<ncclComm_t> comm;
  bool nccl = false;
  std::string why;                     // why NCCL is not in use
};
std::mutex g_gang_mutex;
std::map<int, Gang*> g_gangs;

int get_gang(int n_gpus, Gang** out) {
  std::lock_guard<std::mutex> lock(g_gang_mutex);
  auto it = g_gangs.find(n_gpus);
  if (it != g_gangs.end()) {
    *out = it->second;
    return 0;
  }
  Gang* G = new Gang();
  G->n = n_gpus;
  G->ctx.resize(n_gpus, nullptr);
  for (int g = 0; g < n_gpus; ++g) {
    const int rc = pooled_ctx(g, &G->ctx[g]);
    if (rc) {
      delete G;
      return rc;
    }
  }
  const char* off = getenv("DCB200_NO_NCCL");
  if (n_gpus > 1 && !(off && off[0] == '1')) {
    NcclApi& api = nccl_api();
    if (!api.ok) {
      G->why = api.why;
    } else {
      std::vector<int> devs(n_gpus);
      for (int g = 0; g < n_gpus; ++g) devs[g] = g;
      G->comm.resize(n_gpus);
      const ncclResult_t r = api.CommInitAll(G->comm.data(), n_gpus, devs.data());
      if (r == ncclSuccess) G->nccl = true;
      else G->why = std::string("ncclCommInitAll: ") + api.GetErrorString(r);
    }
    if (!G->nccl && getenv("DCB200_TRACE")) fprintf(stderr, "[dcb200] NCCL not in use (%s): shards are assembled through the host\n", G->why.c_str());
  }
  g_gangs[n_gpus] = G;
  *out = G;
  return 0;
}

#define NCK(call)                                                                                                   \
  do {                                                                                                              \
    ncclResult_t r__ = (call);                                                                                      \
    if (r__ != ncclSuccess) return fail(std::string(#call) + ": " + nccl_api().GetErrorString(r__));                \
  } while (0)

// contiguous shards of the host-assembled fallback: multiples of the row block so that no CTA straddles two devices
void shard(size_t n, int g, int n_gpus, size_t* b, size_t* e) {
  const size_t blocks = (n + ROWS_PER_CTA - 1) / ROWS_PER_CTA;
  const size_t per = (blocks + n_gpus - 1) / n_gpus * ROWS_PER_CTA;
  *b = std::min(n, per * g);
  *e = std::min(n, per * (g + 1));
}

int gpus_for_rows(size_t n_rows, int* n_gpus) {
  CKI(gpus_to_use(n_gpus));
  *n_gpus = (int) std::max<size_t>(1, std::min<size_t>((size_t) *n_gpus, (n_rows + ROWS_PER_CTA - 1) / ROWS_PER_CTA));
  return 0;
}

// One density run on the gang: any of the three stages, results in the caller's host arrays (frame order).
//   radii / pops        populations for n_radii radii (pops: [n_radii][n_rows]) -- skipped when n_radii == 0
//   fe_all              optional: free energies for every radius, [n_radii][n_rows]
//   fe_in               free energies to run the neighbour search on when there is no population stage (else NULL:
//                       the free energies of radius fe_radius, also copied to fe_out when that is not NULL)
//   nn_*                neighbour outputs; nn_idx == NULL skips the stage
struct RunArgs {
  const float* coords;
  size_t n, d;
  const float* radii;
  size_t n_radii, fe_radius;
  uint32_t* pops;
  float* fe_all;
  const float* fe_in;
  float* fe_out;
  uint32_t *nn_idx, *hd_idx;
  float *nn_d2, *hd_d2;
};

int run_single(dcb200_ctx* c, const RunArgs& a, Trace& tr) {
  const size_t n = a.n, R = a.n_radii;
  const bool want_nn = a.nn_idx != nullptr;
  CKI(dcb200_ctx_set_coords(c, a.coords, n, a.d));
  tr.lap("upload+layout");
  CK(c->io_u32.reserve(2 * std::max<size_t>(R, 1) * n + 2 * n));
  CK(c->io_f32.reserve((R + 3) * n));
  uint32_t *pos = c->io_u32.p, *frame = pos + std::max<size_t>(R, 1) * n, *d_ni = frame + std::max<size_t>(R, 1) * n, *d_hi = d_ni + n;
  float *dfe = c->io_f32.p, *d_nd = dfe + n, *d_hd = d_nd + n, *dfe_all = d_hd + n;
  if (R) {
    CKI(dcb200_ctx_populations(c, a.radii, R, 0, n, pos));
    CKI(dcb200_ctx_to_frame_order(c, pos, R, frame));
    if (a.pops) CK(cudaMemcpyAsync(a.pops, frame, R * n * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
    if (a.fe_all) {
      for (size_t r = 0; r < R; ++r) CKI(dcb200_ctx_free_energies(c, frame + r * n, n, 0, dfe_all + r * n));
      CK(cudaMemcpyAsync(a.fe_all, dfe_all, R * n * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    }
    if (want_nn || a.fe_out) {
      CKI(dcb200_ctx_free_energies(c, frame + a.fe_radius * n, n, 0, dfe));
      if (a.fe_out) CK(cudaMemcpyAsync(a.fe_out, dfe, n * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    }
    if (tr.on) { cudaStreamSynchronize(c->stream); tr.lap("populations"); }
  } else if (want_nn) {
    CK(cudaMemcpyAsync(dfe, a.fe_in, n * sizeof(float), cudaMemcpyHostToDevice, c->stream));
  }
  if (want_nn) {
    CK(c->knn.reserve(n));
    CK(c->khd.reserve(n));
    CKI(dcb200_ctx_nn_prepare(c, dfe));
    CKI(dcb200_ctx_nn_scan(c, 0, n, (uint64_t*) c->knn.p, (uint64_t*) c->khd.p));
    CKI(dcb200_ctx_nn_finish(c, (const uint64_t*) c->knn.p, (const uint64_t*) c->khd.p, d_ni, d_nd, d_hi, d_hd));
    CK(cudaMemcpyAsync(a.nn_idx, d_ni, n * 4, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(a.nn_d2, d_nd, n * 4, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(a.hd_idx, d_hi, n * 4, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(a.hd_d2, d_hd, n * 4, cudaMemcpyDeviceToHost, c->stream));
  }
  CK(cudaStreamSynchronize(c->stream));
  tr.lap("scan+download");
  return 0;
}

// several GPUs with NCCL: see the header of this section
int run_gang_nccl(Gang* G, const RunArgs& a) {
  const size_t n = a.n, R = a.n_radii;
  const bool want_nn = a.nn_idx != nullptr;
  const size_t cap = dcb200_shard_capacity(n, G->n);
  Rendezvous rv(G->n);
  NcclApi& api = nccl_api();
  return on_gpus(G->n, [&](int g, int W) -> int {
    dcb200_ctx* c = G->ctx[g];
    ncclComm_t comm = G->comm[g];
    auto step = [&](int rc) -> int {          // agree before the next collective
      if (!rv.all_ok(rc == 0)) return rc ? rc : fail("another GPU of the gang failed");
      return 0;
    };
    auto prep = [&]() -> int {
      CK(cudaSetDevice(c->device));
      CK(c->stage.reserve(n * a.d));
      if (g == 0) CK(cudaMemcpyAsync(c->stage.p, a.coords, n * a.d * sizeof(float), cudaMemcpyHostToDevice, c->stream));
      return 0;
    };
    CKI(step(prep()));
    NCK(api.Broadcast(c->stage.p, c->stage.p, n * a.d, ncclFloat32, 0, comm, c->stream));
    const size_t Rm = std::max<size_t>(R, 1);
    auto layout = [&]() -> int {
      CKI(build_layout(c, c->stage.p, n, a.d, false));
      CK(c->io_u32.reserve(Rm * cap + (size_t) W * Rm * cap + Rm * n + 2 * n));
      CK(c->io_f32.reserve((R + 3) * n));
      return 0;
    };
    int rc = layout();
    uint32_t *loc = c->io_u32.p, *all = loc + Rm * cap, *frame = all + (size_t) W * Rm * cap, *d_ni = frame + Rm * n, *d_hi = d_ni + n;
    float *dfe = c->io_f32.p, *d_nd = dfe + n, *d_hd = d_nd + n, *dfe_all = d_hd + n;
    if (R) {
      if (!rc) rc = dcb200_ctx_populations_shard(c, a.radii, R, g, W, loc);
      CKI(step(rc));
      NCK(api.AllGather(loc, all, R * cap, ncclUint32, comm, c->stream));
      auto after = [&]() -> int {
        // every GPU needs the free energies of all frames for its shard of the neighbour search; only GPU 0 downloads
        if (g == 0 || want_nn) CKI(dcb200_ctx_shards_to_frame_order(c, all, R, W, frame));
        if (g == 0) {
          if (a.pops) CK(cudaMemcpyAsync(a.pops, frame, R * n * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
          if (a.fe_all) {
            for (size_t r = 0; r < R; ++r) CKI(dcb200_ctx_free_energies(c, frame + r * n, n, 0, dfe_all + r * n));
            CK(cudaMemcpyAsync(a.fe_all, dfe_all, R * n * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
          }
        }
        if (want_nn || (g == 0 && a.fe_out)) CKI(dcb200_ctx_free_energies(c, frame + a.fe_radius * n, n, 0, dfe));
        if (g == 0 && a.fe_out) CK(cudaMemcpyAsync(a.fe_out, dfe, n * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
        return 0;
      };
      rc = after();
    } else if (want_nn) {
      if (!rc && g == 0) {
        cudaError_t ce = cudaMemcpyAsync(dfe, a.fe_in, n * sizeof(float), cudaMemcpyHostToDevice, c->stream);
        if (ce != cudaSuccess) rc = fail(std::string("free energy upload: ") + cudaGetErrorString(ce));
      }
      CKI(step(rc));
      NCK(api.Broadcast(dfe, dfe, n, ncclFloat32, 0, comm, c->stream));
      rc = 0;
    }
    if (want_nn) {
      auto scan = [&]() -> int {
        CK(c->knn.reserve(cap + (size_t) W * cap));
        CK(c->khd.reserve(cap + (size_t) W * cap));
        CKI(dcb200_ctx_nn_prepare(c, dfe));
        CKI(dcb200_ctx_nn_scan_shard(c, g, W, (uint64_t*) c->knn.p, (uint64_t*) c->khd.p));
        return 0;
      };
      if (!rc) rc = scan();
      CKI(step(rc));
      NCK(api.GroupStart());
      NCK(api.AllGather(c->knn.p, c->knn.p + cap, cap, ncclUint64, comm, c->stream));
      NCK(api.AllGather(c->khd.p, c->khd.p + cap, cap, ncclUint64, comm, c->stream));
      NCK(api.GroupEnd());
      if (g == 0) {
        CKI(dcb200_ctx_nn_finish_shards(c, (const uint64_t*) (c->knn.p + cap), (const uint64_t*) (c->khd.p + cap), W, cap, d_ni, d_nd, d_hi, d_hd));
        CK(cudaMemcpyAsync(a.nn_idx, d_ni, n * 4, cudaMemcpyDeviceToHost, c->stream));
        CK(cudaMemcpyAsync(a.nn_d2, d_nd, n * 4, cudaMemcpyDeviceToHost, c->stream));
        CK(cudaMemcpyAsync(a.hd_idx, d_hi, n * 4, cudaMemcpyDeviceToHost, c->stream));
        CK(cudaMemcpyAsync(a.hd_d2, d_hd, n * 4, cudaMemcpyDeviceToHost, c->stream));
      }
    } else if (rc) {
      return rc;
    }
    CK(cudaStreamSynchronize(c->stream));
    return 0;
  });
}

int free_energies_host(const uint32_t* pops, size_t n, float* fe);   // below; the caller holds g_call_mutex

// several GPUs without NCCL: every GPU uploads the coordinates itself, contiguous shards, assembly on the host
int run_gang_host(Gang* G, const RunArgs& a) {
  const size_t n = a.n, R = a.n_radii;
  const bool want_nn = a.nn_idx != nullptr;
  std::vector<uint32_t> perm(n), tmp(R * n);
  std::vector<float> fe_host;
  const float* fe_nn = a.fe_in;
  if (R) {
    CKI(on_gpus(G->n, [&](int g, int W) -> int {
      dcb200_ctx* c = G->ctx[g];
      CK(cudaSetDevice(c->device));                        // a fresh worker thread starts on device 0
      size_t b, e;
      shard(n, g, W, &b, &e);
      CKI(dcb200_ctx_set_coords(c, a.coords, n, a.d));     // same deterministic order on every device
      if (e > b) {
        const size_t rows = e - b;
        CK(c->io_u32.reserve(R * rows));
        CKI(dcb200_ctx_populations(c, a.radii, R, b, e, c->io_u32.p));
        CK(cudaMemcpy2DAsync(tmp.data() + b, n * sizeof(uint32_t), c->io_u32.p, rows * sizeof(uint32_t), rows * sizeof(uint32_t), R,
                             cudaMemcpyDeviceToHost, c->stream));
      }
      if (g == 0) CK(cudaMemcpyAsync(perm.data(), c->perm.p, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
      CK(cudaStreamSynchronize(c->stream));
      return 0;
    }));
    std::vector<uint32_t> own;
    uint32_t* pops = a.pops;
    if (!pops) {
      own.resize(R * n);
      pops = own.data();
    }
    for (size_t r = 0; r < R; ++r)
      for (size_t p = 0; p < n; ++p) pops[r * n + perm[p]] = tmp[r * n + p];
    if (a.fe_all)
      for (size_t r = 0; r < R; ++r) CKI(free_energies_host(pops + r * n, n, a.fe_all + r * n));
    if (want_nn || a.fe_out) {
      fe_host.resize(n);
      CKI(free_energies_host(pops + a.fe_radius * n, n, fe_host.data()));
      if (a.fe_out) memcpy(a.fe_out, fe_host.data(), n * sizeof(float));
      fe_nn = fe_host.data();
    }
  }
  if (!want_nn) return 0;
  std::vector<unsigned long long> knn(n), khd(n);
  CKI(on_gpus(G->n, [&](int g, int W) -> int {
    dcb200_ctx* c = G->ctx[g];
    CK(cudaSetDevice(c->device));                          // this worker thread is not the one of the population stage
    size_t b, e;
    shard(n, g, W, &b, &e);
    if (!R) CKI(dcb200_ctx_set_coords(c, a.coords, n, a.d));
    CK(c->io_f32.reserve(n));
    CK(cudaMemcpyAsync(c->io_f32.p, fe_nn, n * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    CKI(dcb200_ctx_nn_prepare(c, c->io_f32.p));
    if (e > b) {
      CK(c->knn.reserve(e - b));
      CK(c->khd.reserve(e - b));
      CKI(dcb200_ctx_nn_scan(c, b, e, (uint64_t*) c->knn.p, (uint64_t*) c->khd.p));
      CK(cudaMemcpyAsync(knn.data() + b, c->knn.p, (e - b) * 8, cudaMemcpyDeviceToHost, c->stream));
      CK(cudaMemcpyAsync(khd.data() + b, c->khd.p, (e - b) * 8, cudaMemcpyDeviceToHost, c->stream));
    }
    if (g == 0) CK(cudaMemcpyAsync(perm.data(), c->perm.p, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return 0;
  }));
  for (size_t p = 0; p < n; ++p) {
    const size_t o = perm[p];
    const uint32_t x = (uint32_t) (knn[p] >> 32), y = (uint32_t) (khd[p] >> 32);
    a.nn_idx[o] = (uint32_t) knn[p];
    memcpy(&a.nn_d2[o], &x, 4);
    a.hd_idx[o] = (uint32_t) khd[p];
    memcpy(&a.hd_d2[o], &y, 4);
  }
  return 0;
}

int run_density(const RunArgs& a, const char* what) {
  if (a.n == 0 || a.d == 0) return fail(std::string(what) + ": empty coordinate array");
  std::lock_guard<std::mutex> call(g_call_mutex);
  int n_gpus = 0;
  CKI(gpus_for_rows(a.n, &n_gpus));
  Trace tr(what);
  if (n_gpus == 1) {
    dcb200_ctx* c = nullptr;
    CKI(pooled_ctx(0, &c));
    return run_single(c, a, tr);
  }
  Gang* G = nullptr;
  CKI(get_gang(n_gpus, &G));
  return G->nccl ? run_gang_nccl(G, a) : run_gang_host(G, a);
}

}  // namespace

extern "C" int dcb200_populations(const float* coords, size_t n_rows, size_t n_cols, const float* radii, size_t n_radii,
                                  uint32_t* pops) {
  if (!coords || !pops || (!radii && n_radii)) return fail("dcb200_populations: null argument");
  if (n_radii == 0) return 0;
  RunArgs a = {coords, n_rows, n_cols, radii, n_radii, 0, pops, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  return run_density(a, "populations");
}

extern "C" int dcb200_nearest_neighbors(const float* coords, size_t n_rows, size_t n_cols, const float* fe, uint32_t* nn_idx,
                                        float* nn_d2, uint32_t* hd_idx, float* hd_d2) {
  if (!coords || !fe || !nn_idx || !nn_d2 || !hd_idx || !hd_d2) return fail("dcb200_nearest_neighbors: null argument");
  RunArgs a = {coords, n_rows, n_cols, nullptr, 0, 0, nullptr, nullptr, fe, nullptr, nn_idx, hd_idx, nn_d2, hd_d2};
  return run_density(a, "nearest_neighbors");
}

// One density run without leaving the device(s) between the stages: ONE upload and ONE layout build serve the populations
// of all radii, the free energies and the neighbour search (the separate entry points above upload and lay out the
// coordinates once each, like the reference's CUDA functions do).
extern "C" int dcb200_density_run(const float* coords, size_t n_rows, size_t n_cols, const float* radii, size_t n_radii,
                                  size_t fe_radius, uint32_t* pops, float* fe_all, float* fe, uint32_t* nn_idx, float* nn_d2,
                                  uint32_t* hd_idx, float* hd_d2) {
  if (!coords || !radii || n_radii == 0) return fail("dcb200_density_run: needs coordinates and at least one radius");
  if (fe_radius >= n_radii) return fail("dcb200_density_run: fe_radius out of range");
  if (nn_idx && (!nn_d2 || !hd_idx || !hd_d2)) return fail("dcb200_density_run: the four neighbour outputs go together");
  RunArgs a = {coords, n_rows, n_cols, radii, n_radii, fe_radius, pops, fe_all, nullptr, fe, nn_idx, hd_idx, nn_d2, hd_d2};
  return run_density(a, "density_run");
}

extern "C" int dcb200_free_energies(const uint32_t* pops, size_t n, float* fe) {
  if (!pops || !fe) return fail("dcb200_free_energies: null argument");
  if (n == 0) return 0;
  std::lock_guard<std::mutex> call(g_call_mutex);
  return free_energies_host(pops, n, fe);
}

namespace {
// caller holds g_call_mutex
int free_energies_host(const uint32_t* pops, size_t n, float* fe) {
  int n_gpus = 0;
  CKI(gpus_to_use(&n_gpus));
  dcb200_ctx* c = nullptr;
  CKI(pooled_ctx(0, &c));
  CK(cudaSetDevice(c->device));
  CK(c->io_u32.reserve(n));
  CK(c->io_f32.reserve(n));
  uint32_t* dp = c->io_u32.p;
  float* df = c->io_f32.p;
  CK(cudaMemcpyAsync(dp, pops, n * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream));
  CKI(dcb200_ctx_free_energies(c, dp, n, 0, df));
  CK(cudaMemcpyAsync(fe, df, n * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return 0;
}
}  // namespace

// ---- screening ----------------------------------------------------------------------------------
// A screening session works on ALL free-energy-sorted frames at once.  At its first step it scans the frames -- laid out in
// spatial order, so that tiles out of reach of the cut radius are never touched -- for every pair with d2 < cut (edge_kernel,
// kernels.cuh; rows dealt block-cyclically to the GPUs of the gang) and sorts the edges by the sorted index of their later
// frame.  A threshold then only unions the edges that became active since the previous one (lock-free union-find on GPU 0)
// and downloads the representatives.  The reference scans (new frames) x (all lower frames) at every threshold
// (density_clustering_common.cpp:98-123, density_clustering_cuda.cu:505-571): N^2/2 pairs per run; the edge list costs
// about one population scan with r^2 = cut.
namespace {

__global__ void edge_union_kernel(const unsigned long long* __restrict__ edges, size_t e0, size_t e1, uint32_t* parent) {
  const size_t q = e0 + (size_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= e1) return;
  const unsigned long long e = edges[q];
  uf_union(parent, (uint32_t) (e >> 32), (uint32_t) e);
}
// out[0] = number of edges (sorted by level) whose level (larger sorted index, high word) is below `limit`
__global__ void edge_count_below_kernel(const unsigned long long* __restrict__ edges, size_t n_edges, uint32_t limit,
                                        unsigned long long* __restrict__ out) {
  size_t lo = 0, hi = n_edges;
  while (lo < hi) {
    const size_t mid = (lo + hi) >> 1;
    if ((uint32_t) (edges[mid] >> 32) < limit) lo = mid + 1; else hi = mid;
  }
  out[0] = lo;
}

// cluster numbers on the device: clusters are numbered 1..K by ascending representative (= first sorted member)
__global__ void root_flag_kernel(const uint32_t* __restrict__ forest, size_t m, uint32_t* __restrict__ flag) {
  const size_t p = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (p < m) flag[p] = forest[p] == (uint32_t) p ? 1u : 0u;
}
// labels[order[p]] = number of p's cluster for p < m, 0 for the frames above the threshold (rank = exclusive sum of the flags)
__global__ void label_scatter_kernel(const uint32_t* __restrict__ forest, const uint32_t* __restrict__ rank, const uint32_t* __restrict__ order,
                                     size_t m, size_t n, uint32_t* __restrict__ labels) {
  const size_t p = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  labels[order[p]] = p < m ? rank[forest[p]] + 1u : 0u;
}

static cudaError_t launch_edge(int d, const EdgeArgs& a, int grid, cudaStream_t st) {
  switch (d <= MAX_TEMPLATE_D ? d : 0) {
#define FN(D) launch_edge_d##D(a, grid, st)
    DCB_FOR_EACH_D(DCB_DISPATCH)
#undef FN
  }
  return cudaErrorInvalidValue;
}
static int occ_edge(int d) {
  switch (d <= MAX_TEMPLATE_D ? d : 0) {
#define FN(D) occupancy_edge_d##D(d)
    DCB_FOR_EACH_D(DCB_DISPATCH)
#undef FN
  }
  return 0;
}

// buf: [0] = counter, [1 ..] = edges.  Scans shard `shard` of `n_shards` (block-cyclic rows, upper triangle of the position
// matrix) of the context's spatially ordered frames; grows the buffer and scans again if the edges did not fit.
int edge_scan(dcb200_ctx* c, int shard, int n_shards, float cut, uint32_t level_min, DevBuf<unsigned long long>& buf,
              unsigned long long* found) {
  if (!c->spatial) return fail("dcb200: the edge scan needs the frames in spatial order");
  CK(cudaSetDevice(c->device));
  const size_t b = std::min(c->n, (size_t) shard * ROWS_PER_CTA);
  *found = 0;
  if (b >= c->n) return 0;
  const size_t rows = launch_rows(b, c->n, (size_t) n_shards);
  if (buf.cap < 2) CK(buf.reserve(std::max<size_t>(size_t(1) << 20, 24 * rows) + 1));
  for (int attempt = 0; attempt < 3; ++attempt) {
    EdgeArgs a;
    int grid = 0;
    CKI(fill_geom(c, b, c->n, (size_t) n_shards, tile_width(c->d), occ_edge((int) c->d), 32u, &a.g, &grid));
    a.cut = cut;
    a.thr_fast = up((double) cut * (1.0 + 1.01 * (double) a.g.e_rel));
    a.g.prune_thr = up((double) a.thr_fast * 1.0001 + 2.0 * (double) a.g.prune_slack);
    a.rank = c->perm.p;                       // the coordinates came in free-energy order: position -> sorted index
    a.level_min = level_min;
    a.count = buf.p;
    a.edges = buf.p + 1;
    a.cap = buf.cap - 1;
    CK(cudaMemsetAsync(buf.p, 0, sizeof(unsigned long long), c->stream));
    CK(cudaMemsetAsync(c->scalars, 0, sizeof(unsigned int), c->stream));
    CK(launch_edge((int) c->d, a, grid, c->stream));
    c->launches += 1;
    unsigned long long n_found = 0;
    CK(cudaMemcpyAsync(&n_found, buf.p, sizeof(n_found), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    *found = n_found;
    if (n_found <= a.cap) return 0;
    CK(buf.reserve((size_t) n_found + (size_t) (n_found / 16) + 2));          // now the count is known: scan once more
  }
  return fail("dcb200: the edge list kept growing between scans");
}

}  // namespace

struct dcb200_screen {
  size_t n = 0, d = 0;
  int n_gpus = 1;
  Gang* gang = nullptr;
  size_t m_done = 0;                      // sorted positions [0, m_done) are in the forest
  std::vector<float> sorted;              // host copy of the sorted coordinates: the per-GPU contexts are shared with the other
  std::vector<uint64_t> gen;              // entry points; if one of them replaced the layout (gen), the session restores it
  // GPU 0: the forest and the edge list, sorted by level (= the larger sorted index of the pair)
  DevBuf<uint32_t> forest;
  DevBuf<unsigned long long> edges, edges_alt;
  std::vector<DevBuf<unsigned long long>> part;      // per GPU: counter + the edges its rows found
  size_t n_edges = 0, e_done = 0;
  float cut = 0.f;
  bool have_edges = false;
  DevBuf<uint32_t> order, flag, rank, labels;        // naming on the device (dcb200_screen_labels)
  bool have_order = false;
};

// (re)builds the session's layout on every GPU of its gang: one upload (+ NCCL broadcast), spatial order
static int screen_upload(dcb200_screen* s) {
  Gang* G = s->gang;
  Rendezvous rv(G->n);
  const size_t count = s->n * s->d;
  return on_gpus(G->n, [&](int g, int W) -> int {
    dcb200_ctx* c = G->ctx[g];
    auto prep = [&]() -> int {
      CK(cudaSetDevice(c->device));
      CK(c->stage.reserve(count));
      if (g == 0 || !G->nccl) CK(cudaMemcpyAsync(c->stage.p, s->sorted.data(), count * sizeof(float), cudaMemcpyHostToDevice, c->stream));
      return 0;
    };
    const int r0 = prep();
    if (W > 1 && G->nccl) {
      if (!rv.all_ok(r0 == 0)) return r0 ? r0 : fail("another GPU of the gang failed");
      NCK(nccl_api().Broadcast(c->stage.p, c->stage.p, count, ncclFloat32, 0, G->comm[g], c->stream));
    } else if (r0) {
      return r0;
    }
    CKI(build_layout(c, c->stage.p, s->n, s->d, false));
    s->gen[g] = c->layout_gen;
    return 0;
  });
}

// all edges with level >= level_min of the session's frames, gathered on GPU 0 and sorted by level
static int screen_build_edges(dcb200_screen* s, float cut, uint32_t level_min) {
  Gang* G = s->gang;
  bool replaced = false;                  // another entry point used the shared contexts in between: restore the layout
  for (int g = 0; g < G->n; ++g) replaced |= G->ctx[g]->layout_gen != s->gen[g];
  if (replaced) CKI(screen_upload(s));
  std::vector<unsigned long long> found(G->n, 0);
  CKI(on_gpus(G->n, [&](int g, int W) -> int { return edge_scan(G->ctx[g], g, W, cut, level_min, s->part[g], &found[g]); }));
  size_t total = 0;
  for (int g = 0; g < G->n; ++g) total += (size_t) found[g];
  dcb200_ctx* c0 = G->ctx[0];
  CK(cudaSetDevice(c0->device));
  CK(s->edges.reserve(std::max<size_t>(total, 1)));
  CK(s->edges_alt.reserve(std::max<size_t>(total, 1)));
  size_t off = 0;
  for (int g = 0; g < G->n; ++g) {
    if (found[g] == 0) continue;
    // peer copy: direct over NVLink where peer access is possible, staged by the driver otherwise
    CK(cudaMemcpyPeerAsync(s->edges_alt.p + off, c0->device, s->part[g].p + 1, G->ctx[g]->device, (size_t) found[g] * sizeof(unsigned long long),
                           c0->stream));
    off += (size_t) found[g];
  }
  if (total) {
    size_t tmp_bytes = 0;
    CK(cub::DeviceRadixSort::SortKeys(nullptr, tmp_bytes, s->edges_alt.p, s->edges.p, (int) total, 32, 64, c0->stream));
    CK(c0->cub_tmp.reserve(tmp_bytes));
    CK(cub::DeviceRadixSort::SortKeys(c0->cub_tmp.p, tmp_bytes, s->edges_alt.p, s->edges.p, (int) total, 32, 64, c0->stream));
    c0->launches += 4;
  }
  CK(cudaStreamSynchronize(c0->stream));
  s->n_edges = total;
  s->e_done = 0;
  s->cut = cut;
  s->have_edges = true;
  if (getenv("DCB200_TRACE")) fprintf(stderr, "[dcb200] screening: %zu edges within the cut (%.1f per frame)\n", total, (double) total / (double) s->n);
  return 0;
}

static void screen_free(dcb200_screen* s);   // caller holds g_call_mutex

extern "C" int dcb200_screen_begin(const float* sorted_coords, size_t n_sorted, size_t n_cols, dcb200_screen** out) {
  if (!sorted_coords || !out) return fail("dcb200_screen_begin: null argument");
  *out = nullptr;
  if (n_sorted == 0 || n_cols == 0) return fail("dcb200_screen_begin: empty coordinate array");
  if (n_sorted >= 0x7fffffffull) return fail("dcb200_screen_begin: too many frames");
  std::lock_guard<std::mutex> call(g_call_mutex);
  dcb200_screen* s = new dcb200_screen();
  s->n = n_sorted;
  s->d = n_cols;
  int rc = gpus_for_rows(n_sorted, &s->n_gpus);
  if (!rc) rc = get_gang(s->n_gpus, &s->gang);
  if (rc) {
    delete s;
    return rc;
  }
  Gang* G = s->gang;
  s->part.resize(G->n);
  s->gen.assign(G->n, 0);
  s->sorted.assign(sorted_coords, sorted_coords + n_sorted * n_cols);
  rc = screen_upload(s);
  if (!rc) {
    dcb200_ctx* c = G->ctx[0];
    auto forest = [&]() -> int {
      CK(cudaSetDevice(c->device));
      CK(s->forest.reserve(n_sorted));
      iota_kernel<<<blocks_for(n_sorted, 256), 256, 0, c->stream>>>(s->forest.p, n_sorted);
      CK(cudaGetLastError());
      CK(cudaStreamSynchronize(c->stream));
      return 0;
    };
    rc = forest();
  }
  if (rc) {
    screen_free(s);
    return rc;
  }
  *out = s;
  return 0;
}

// Extends the forest to the sorted positions [0, m_new) (edges: d2 < max_dist2 between a new frame and any lower position);
// seed (optional, uint32 [m_done]) replaces the forest of the positions done so far.  The flattened forest stays on GPU 0.
static int screen_advance(dcb200_screen* s, size_t m_new, float max_dist2, const uint32_t* seed) {
  if (m_new > s->n) return fail("dcb200_screen_step: m_new exceeds the session's frames");
  if (m_new < s->m_done) return fail("dcb200_screen_step: thresholds must not decrease within a session");
  const size_t m_prev = s->m_done;
  if (seed)
    for (size_t p = 0; p < m_prev; ++p)
      if (seed[p] > p) return fail("dcb200_screen_step: seed[p] must be <= p");
  // the edge list is built at the first step (the cut is known only now); a different cut means a new list.  Edges between
  // two frames that are settled already (both below m_prev) are never needed.
  if (!s->have_edges || max_dist2 != s->cut) CKI(screen_build_edges(s, max_dist2, (uint32_t) m_prev));
  dcb200_ctx* c = s->gang->ctx[0];
  CK(cudaSetDevice(c->device));
  if (seed && m_prev) CK(cudaMemcpyAsync(s->forest.p, seed, m_prev * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream));
  // edges that become active: level in [m_prev, m_new)
  unsigned long long e_new = 0;
  {
    unsigned long long* dcount = reinterpret_cast<unsigned long long*>(c->scalars + 2);   // transient scratch ([2], [3])
    edge_count_below_kernel<<<1, 1, 0, c->stream>>>(s->edges.p, s->n_edges, (uint32_t) m_new, dcount);
    CK(cudaMemcpyAsync(&e_new, dcount, sizeof(e_new), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
  }
  if (e_new > s->e_done) {
    const size_t cnt = (size_t) e_new - s->e_done;
    edge_union_kernel<<<blocks_for(cnt, 256), 256, 0, c->stream>>>(s->edges.p, s->e_done, (size_t) e_new, s->forest.p);
    c->launches += 1;
  }
  if (m_new) CKI(dcb200_ctx_screening_flatten(c, m_new, s->forest.p));
  CK(cudaGetLastError());
  s->e_done = std::max(s->e_done, (size_t) e_new);
  s->m_done = m_new;
  return 0;
}

// ... and writes every position's representative (smallest sorted position of its cluster) to comp[0, m_new)
extern "C" int dcb200_screen_step(dcb200_screen* s, size_t m_new, float max_dist2, const uint32_t* seed, uint32_t* comp) {
  if (!s || !comp) return fail("dcb200_screen_step: null argument");
  std::lock_guard<std::mutex> call(g_call_mutex);
  CKI(screen_advance(s, m_new, max_dist2, seed));
  dcb200_ctx* c = s->gang->ctx[0];
  if (m_new) CK(cudaMemcpyAsync(comp, s->forest.p, m_new * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return 0;
}

// order[p] = frame at sorted position p (all n_sorted of them): lets the session name the clusters on the device
extern "C" int dcb200_screen_set_order(dcb200_screen* s, const uint32_t* order) {
  if (!s || !order) return fail("dcb200_screen_set_order: null argument");
  std::lock_guard<std::mutex> call(g_call_mutex);
  dcb200_ctx* c = s->gang->ctx[0];
  CK(cudaSetDevice(c->device));
  CK(s->order.reserve(s->n));
  CK(s->flag.reserve(s->n));
  CK(s->rank.reserve(s->n));
  CK(s->labels.reserve(s->n));
  CK(cudaMemcpyAsync(s->order.p, order, s->n * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  s->have_order = true;
  return 0;
}

// One threshold with the naming on the device: labels [n_sorted] in FRAME order (through the order given to
// dcb200_screen_set_order), clusters numbered 1..K by ascending representative, 0 above the threshold.
extern "C" int dcb200_screen_labels(dcb200_screen* s, size_t m_new, float max_dist2, uint32_t* labels, uint32_t* n_clusters) {
  if (!s || !labels) return fail("dcb200_screen_labels: null argument");
  std::lock_guard<std::mutex> call(g_call_mutex);
  if (!s->have_order) return fail("dcb200_screen_labels: call dcb200_screen_set_order first");
  CKI(screen_advance(s, m_new, max_dist2, nullptr));
  dcb200_ctx* c = s->gang->ctx[0];
  uint32_t last[2] = {0, 0};
  if (m_new) {
    root_flag_kernel<<<blocks_for(m_new, 256), 256, 0, c->stream>>>(s->forest.p, m_new, s->flag.p);
    size_t tmp_bytes = 0;
    CK(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, s->flag.p, s->rank.p, (int) m_new, c->stream));
    CK(c->cub_tmp.reserve(tmp_bytes));
    CK(cub::DeviceScan::ExclusiveSum(c->cub_tmp.p, tmp_bytes, s->flag.p, s->rank.p, (int) m_new, c->stream));
    CK(cudaMemcpyAsync(&last[0], s->rank.p + (m_new - 1), sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(&last[1], s->flag.p + (m_new - 1), sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
  }
  label_scatter_kernel<<<blocks_for(s->n, 256), 256, 0, c->stream>>>(s->forest.p, s->rank.p, s->order.p, m_new, s->n, s->labels.p);
  c->launches += 4;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(labels, s->labels.p, s->n * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  if (n_clusters) *n_clusters = last[0] + last[1];
  return 0;
}

extern "C" int dcb200_screen_end(dcb200_screen* s) {
  if (!s) return 0;
  std::lock_guard<std::mutex> call(g_call_mutex);
  screen_free(s);
  return 0;
}

// caller holds g_call_mutex
static void screen_free(dcb200_screen* s) {
  if (s->gang) {
    for (size_t g = 0; g < s->part.size(); ++g) {
      cudaSetDevice(s->gang->ctx[g]->device);
      s->part[g].release();
    }
    cudaSetDevice(s->gang->ctx[0]->device);
    s->forest.release();
    s->edges.release();
    s->edges_alt.release();
    s->order.release();
    s->flag.release();
    s->rank.release();
    s->labels.release();
  }
  delete s;
}

// stateless form: one session per call
extern "C" int dcb200_screening_step(const float* sorted_coords, size_t n_cols, size_t m_prev, size_t m_new, float max_dist2,
                                     uint32_t* comp) {
  if (!sorted_coords || !comp) return fail("dcb200_screening_step: null argument");
  if (m_prev > m_new) return fail("dcb200_screening_step: m_prev > m_new");
  if (m_new == m_prev) return 0;
  for (size_t p = 0; p < m_prev; ++p)
    if (comp[p] > p) return fail("dcb200_screening_step: comp[p] must be <= p");
  dcb200_screen* s = nullptr;
  CKI(dcb200_screen_begin(sorted_coords, m_new, n_cols, &s));
  s->m_done = m_prev;
  std::vector<uint32_t> seed(comp, comp + m_prev);
  const int rc = dcb200_screen_step(s, m_new, max_dist2, m_prev ? seed.data() : nullptr, comp);
  dcb200_screen_end(s);
  return rc;
}
