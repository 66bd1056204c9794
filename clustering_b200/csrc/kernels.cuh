// The pair-scan kernels of the `clustering density` hot path, written for sm_100a:
//
//   pops_kernel<D>    multi-radius neighbourhood population count
//                     (replaces reference density_clustering.cpp:126-195 / density_clustering_cuda_kernels.cu:9-56)
//   nn_kernel<D>      nearest neighbour + nearest neighbour with lower free energy, fused
//                     (replaces density_clustering.cpp:230-288 / density_clustering_cuda_kernels.cu:58-130)
//   screen_kernel<D>  edge discovery + union-find of the free-energy screening
//                     (replaces density_clustering_common.cpp:37-134 / density_clustering_cuda_kernels.cu:132-192)
//
// Common structure (not derived from the reference's kernels):
//   * persistent CTAs, 8 consumer warps + 1 producer warp; work items (row block x column range) are
//     handed out by an atomic counter, so dense and sparse regions balance dynamically;
//   * the producer streams column tiles [D+1][TJ] (dim-major, 16 B aligned) into a 3-stage shared
//     memory ring with 1-D bulk TMA (cp.async.bulk + mbarrier complete_tx); no __syncthreads in steady state;
//   * every consumer thread keeps RI=4 rows in registers and sweeps CJ=4 columns per step with one
//     broadcast LDS.128 per dim and RI*CJ FFMAs: acc = |y|^2 - 2 x.y (the column pack holds -2y and |y|^2
//     of the centred coordinates), so the common path costs D FFMA + ~1 compare per pair;
//   * that fast value only *filters*: a pair is handed to the slow path when acc < t_row, where t_row
//     carries a proven rounding-error margin.  The slow path decides with the squared distance
//     evaluated in the exact rounding order of the reference's CPU build (dist2_exact) whenever the fast
//     value is within the error band of a decision boundary, so results are bit-identical.
//   D = 0 selects the run-time-D variant (any n_cols): 4 rows x 16 columns per step, row operands
//   streamed from L1/L2 instead of registers.
#pragma once
#include "common.cuh"
#include <float.h>

namespace dcb {

struct ScanGeom {
  const float* xT;          // [D][ld]   original coordinates, dim-major, NaN padded
  const float* cT;          // [D+1][ld] column pack of the centred coordinates x' = x - centre:
                            //           rows 0..D-1 = -2*x', row D = |x'|^2 (NaN padded)
  size_t ld;                // padded frame count (multiple of 256)
  int d;                    // n_cols (== D for the specialised kernels)
  uint32_t n;               // real frame count
  uint32_t row_begin, row_end;      // rows of this shard
  uint32_t n_row_blocks;            // ceil((row_end-row_begin)/ROWS_PER_CTA)
  uint32_t n_col_tiles;             // column tiles of this launch
  uint32_t tiles_per_item;          // column tiles per work item
  uint32_t n_col_items;             // ceil(n_col_tiles / tiles_per_item)
  unsigned int* work_counter;       // zero-initialised before the launch
  float e_abs, e_rel;               // |fast - exact| <= e_abs + e_rel * value   (api.cu: error_bounds)
  unsigned long long* stats;        // [0] pairs handed to the slow path, [1] pairs re-evaluated exactly
};

// per-thread slow-path counters, added to g.stats once per kernel
struct SlowStats {
  uint32_t slow = 0, exact = 0;
  __device__ __forceinline__ void flush(const ScanGeom& g) {
    for (int o = 16; o > 0; o >>= 1) {
      slow += __shfl_xor_sync(0xffffffffu, slow, o);
      exact += __shfl_xor_sync(0xffffffffu, exact, o);
    }
    if ((threadIdx.x & 31) == 0 && g.stats) {
      if (slow) atomicAdd(g.stats, (unsigned long long) slow);
      if (exact) atomicAdd(g.stats + 1, (unsigned long long) exact);
    }
  }
};

// tile width: 128 columns for the register kernels, 64 for the run-time-D kernel
template <int D> struct TileW { static constexpr int tj = 128; static constexpr int cj = CJ; };
template <> struct TileW<0> { static constexpr int tj = 64; static constexpr int cj = 16; };

// ------------------------------------------------------------------------------------------------
// shared memory ring common to the kernels
// ------------------------------------------------------------------------------------------------
template <int D>
struct SmemRing {
  static constexpr int TJ = TileW<D>::tj;
  float* tiles;            // STAGES * (d+1) * TJ
  uint64_t* full;          // STAGES
  uint64_t* empty;         // STAGES
  TileMeta* meta;          // STAGES
  size_t tile_floats;
  __host__ __device__ static size_t bytes(int d) {
    return (size_t) STAGES * (d + 1) * TJ * 4 + 2 * STAGES * 8 + STAGES * sizeof(TileMeta);
  }
  __device__ SmemRing(unsigned char* base, int d) {
    tile_floats = (size_t) (d + 1) * TJ;
    tiles = reinterpret_cast<float*>(base);
    full = reinterpret_cast<uint64_t*>(base + STAGES * tile_floats * 4);
    empty = full + STAGES;
    meta = reinterpret_cast<TileMeta*>(empty + STAGES);
  }
  __device__ void init() {
    if (threadIdx.x == 0) {
      for (int s = 0; s < STAGES; ++s) {
        mbar_init(&full[s], 1);
        mbar_init(&empty[s], N_CONSUMER_WARPS);
      }
      fence_mbar_init();
    }
  }
};

// Producer: one elected lane walks the work items and streams their column tiles.
// range(rb, lim0, lim1) lets a kernel restrict the column tiles per row block (screening only
// needs the columns below its rows).
template <int D, class TileRange>
__device__ __forceinline__ void produce(const ScanGeom& g, SmemRing<D>& ring, TileRange&& range) {
  constexpr int TJ = TileW<D>::tj;
  Pipe pp;
  const int d = D ? D : g.d;
  const uint32_t total = g.n_row_blocks * g.n_col_items;
  for (;;) {
    const uint32_t item = atomicAdd(g.work_counter, 1u);
    if (item >= total) break;
    const uint32_t rb = item / g.n_col_items, ci = item % g.n_col_items;
    uint32_t t0 = ci * g.tiles_per_item;
    uint32_t t1 = min(t0 + g.tiles_per_item, g.n_col_tiles);
    uint32_t lim0 = 0, lim1 = g.n_col_tiles;
    range(rb, lim0, lim1);
    t0 = max(t0, lim0);
    t1 = min(t1, lim1);
    for (uint32_t t = t0; t < t1; ++t) {
      mbar_wait(&ring.empty[pp.stage], pp.phase ^ 1);
      TileMeta m;
      m.row_block = (int32_t) rb;
      m.col0 = t * TJ;
      m.flags = (t == t0 ? 1u : 0u) | (t + 1 == t1 ? 2u : 0u);
      m.aux = 0;
      ring.meta[pp.stage] = m;
      mbar_arrive_expect_tx(&ring.full[pp.stage], (uint32_t) (ring.tile_floats * 4));
      float* dst = ring.tiles + pp.stage * ring.tile_floats;
      const float* src = g.cT + m.col0;
#pragma unroll 1
      for (int k = 0; k <= d; ++k) tma_load_1d(dst + k * TJ, src + (size_t) k * g.ld, TJ * 4, &ring.full[pp.stage]);
      pp.advance();
    }
  }
  mbar_wait(&ring.empty[pp.stage], pp.phase ^ 1);
  ring.meta[pp.stage].row_block = -1;
  mbar_arrive(&ring.full[pp.stage]);
}

// Row operands of a consumer thread.
template <int D>
struct Rows {
  float x[RI][D];           // centred coordinates x'
  float xn[RI];             // |x'|^2
  uint32_t row[RI];
  __device__ __forceinline__ void load(const ScanGeom& g, uint32_t rb, int tid) {
#pragma unroll
    for (int r = 0; r < RI; ++r) {
      row[r] = g.row_begin + rb * ROWS_PER_CTA + r * N_CONSUMERS + tid;
      const size_t p = min((size_t) row[r], g.ld - 1);     // rows past the end: clamped, results discarded
#pragma unroll
      for (int k = 0; k < D; ++k) x[r][k] = -0.5f * __ldg(g.cT + (size_t) k * g.ld + p);
      xn[r] = __ldg(g.cT + (size_t) D * g.ld + p);
    }
  }
};
template <>
struct Rows<0> {
  float xn[RI];
  uint32_t row[RI];
  size_t p[RI];
  __device__ __forceinline__ void load(const ScanGeom& g, uint32_t rb, int tid) {
#pragma unroll
    for (int r = 0; r < RI; ++r) {
      row[r] = g.row_begin + rb * ROWS_PER_CTA + r * N_CONSUMERS + tid;
      p[r] = min((size_t) row[r], g.ld - 1);
      xn[r] = __ldg(g.cT + (size_t) g.d * g.ld + p[r]);
    }
  }
};

// Consumer inner loop over one tile: acc[r][c] = |y_c|^2 - 2 x_r.y_c for an RI x CJ block per step;
// hit(r, j_in_tile, acc) is called for every pair with acc < t[r].
template <int D, class Hit>
__device__ __forceinline__ void scan_tile(const ScanGeom&, const float* __restrict__ tl, const Rows<D>& R, float (&t)[RI],
                                          Hit&& hit) {
  constexpr int TJ = TileW<D>::tj;
#pragma unroll 2
  for (int g = 0; g < TJ; g += CJ) {
    float acc[RI][CJ];
    {
      const float4 n4 = *reinterpret_cast<const float4*>(tl + D * TJ + g);
      const float4 y4 = *reinterpret_cast<const float4*>(tl + g);
#pragma unroll
      for (int r = 0; r < RI; ++r) {
        acc[r][0] = fmaf(R.x[r][0], y4.x, n4.x);
        acc[r][1] = fmaf(R.x[r][0], y4.y, n4.y);
        acc[r][2] = fmaf(R.x[r][0], y4.z, n4.z);
        acc[r][3] = fmaf(R.x[r][0], y4.w, n4.w);
      }
    }
#pragma unroll
    for (int k = 1; k < D; ++k) {
      const float4 y4 = *reinterpret_cast<const float4*>(tl + k * TJ + g);
#pragma unroll
      for (int r = 0; r < RI; ++r) {
        acc[r][0] = fmaf(R.x[r][k], y4.x, acc[r][0]);
        acc[r][1] = fmaf(R.x[r][k], y4.y, acc[r][1]);
        acc[r][2] = fmaf(R.x[r][k], y4.z, acc[r][2]);
        acc[r][3] = fmaf(R.x[r][k], y4.w, acc[r][3]);
      }
    }
    bool any = false;
#pragma unroll
    for (int r = 0; r < RI; ++r) {
      const float m = fminf(fminf(acc[r][0], acc[r][1]), fminf(acc[r][2], acc[r][3]));
      any |= (m < t[r]);
    }
    if (any) {
#pragma unroll
      for (int r = 0; r < RI; ++r)
#pragma unroll
        for (int c = 0; c < CJ; ++c)
          if (acc[r][c] < t[r]) hit(r, g + c, acc[r][c]);
    }
  }
}

// run-time-D variant: 4 rows x 16 columns per step, row operands from L1/L2
template <class Hit>
__device__ __forceinline__ void scan_tile(const ScanGeom& gm, const float* __restrict__ tl, const Rows<0>& R, float (&t)[RI],
                                          Hit&& hit) {
  constexpr int TJ = TileW<0>::tj, CG = TileW<0>::cj;
  const int d = gm.d;
#pragma unroll 1
  for (int g = 0; g < TJ; g += CG) {
    float acc[RI][CG];
#pragma unroll
    for (int c = 0; c < CG; ++c) {
      const float nrm = tl[d * TJ + g + c];
#pragma unroll
      for (int r = 0; r < RI; ++r) acc[r][c] = nrm;
    }
#pragma unroll 2
    for (int k = 0; k < d; ++k) {
      float xr[RI];
#pragma unroll
      for (int r = 0; r < RI; ++r) xr[r] = -0.5f * __ldg(gm.cT + (size_t) k * gm.ld + R.p[r]);
#pragma unroll
      for (int c4 = 0; c4 < CG; c4 += 4) {
        const float4 y4 = *reinterpret_cast<const float4*>(tl + k * TJ + g + c4);
#pragma unroll
        for (int r = 0; r < RI; ++r) {
          acc[r][c4 + 0] = fmaf(xr[r], y4.x, acc[r][c4 + 0]);
          acc[r][c4 + 1] = fmaf(xr[r], y4.y, acc[r][c4 + 1]);
          acc[r][c4 + 2] = fmaf(xr[r], y4.z, acc[r][c4 + 2]);
          acc[r][c4 + 3] = fmaf(xr[r], y4.w, acc[r][c4 + 3]);
        }
      }
    }
    bool any = false;
#pragma unroll
    for (int r = 0; r < RI; ++r) {
      float m = acc[r][0];
#pragma unroll
      for (int c = 1; c < CG; ++c) m = fminf(m, acc[r][c]);
      any |= (m < t[r]);
    }
    if (any) {
#pragma unroll
      for (int r = 0; r < RI; ++r)
#pragma unroll
        for (int c = 0; c < CG; ++c)
          if (acc[r][c] < t[r]) hit(r, g + c, acc[r][c]);
    }
  }
}

#define DCB_LAUNCH_BOUNDS(D) __launch_bounds__(CTA_THREADS, ((D) >= 1 && (D) <= 10) ? 2 : 1)

// ================================================================================================
// populations
// ================================================================================================
struct PopsArgs {
  ScanGeom g;
  int n_bins;               // distinct radii in this pass (<= MAX_BINS)
  float rad2[32];           // ascending squared radii, padded with +inf
  float thr_fast;           // rad2[n_bins-1] + error margin
  uint32_t* cnt;            // [n_bins][ld_cnt]: #{j != i : d2(i,j) < rad2[b]}, rows relative to row_begin
  size_t ld_cnt;
};

__host__ __device__ inline size_t pops_smem_bytes(size_t ring_bytes, int n_bins) {
  return ((ring_bytes + 15) & ~size_t(15)) + 32 * 4 + (size_t) n_bins * ROWS_PER_CTA * 4;
}

template <int D>
__global__ void DCB_LAUNCH_BOUNDS(D) pops_kernel(const __grid_constant__ PopsArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  const ScanGeom& g = a.g;
  const int d = D ? D : g.d;
  SmemRing<D> ring(smem, d);
  float* rad2s = reinterpret_cast<float*>(smem + ((SmemRing<D>::bytes(d) + 15) & ~size_t(15)));
  uint32_t* hist = reinterpret_cast<uint32_t*>(rad2s + 32);       // [n_bins][ROWS_PER_CTA], slot-private counters
  ring.init();
  if (threadIdx.x < 32) rad2s[threadIdx.x] = a.rad2[threadIdx.x];
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == N_CONSUMER_WARPS) {
    if (lane == 0) produce<D>(g, ring, [](uint32_t, uint32_t&, uint32_t&) {});
    return;
  }
  const int tid = threadIdx.x;
  const int nb = a.n_bins;
  Rows<D> R;
  float t[RI];
  Pipe cp;
  SlowStats st;
  for (;;) {
    mbar_wait(&ring.full[cp.stage], cp.phase);
    const TileMeta m = ring.meta[cp.stage];
    if (m.row_block < 0) break;
    if (m.flags & 1u) {
      R.load(g, (uint32_t) m.row_block, tid);
#pragma unroll
      for (int r = 0; r < RI; ++r) t[r] = next_up(a.thr_fast - R.xn[r]);
      for (int b = 0; b < nb; ++b)
#pragma unroll
        for (int r = 0; r < RI; ++r) hist[b * ROWS_PER_CTA + r * N_CONSUMERS + tid] = 0;
    }
    const float* tl = ring.tiles + cp.stage * ring.tile_floats;
    scan_tile(g, tl, R, t, [&](int r, int jt, float accv) {
      const uint32_t j = m.col0 + jt;
      float s = accv + R.xn[r];
      ++st.slow;
      int b = 0;
#pragma unroll
      for (int step = 16; step >= 1; step >>= 1)
        if (rad2s[b + step - 1] <= s) b += step;
      const float lo = b > 0 ? rad2s[b - 1] : -INFINITY;
      const float hi = rad2s[b];
      const float e = fmaf(g.e_rel, fabsf(s), g.e_abs);
      if ((s - lo < e) || (hi - s <= e)) {       // within the error band of a radius: decide exactly
        s = dist2_exact(g.xT, g.ld, d, R.row[r], j);
        ++st.exact;
        b = 0;
#pragma unroll
        for (int step = 16; step >= 1; step >>= 1)
          if (rad2s[b + step - 1] <= s) b += step;
      }
      if (j != R.row[r] && b < nb) hist[b * ROWS_PER_CTA + r * N_CONSUMERS + tid] += 1;
    });
    __syncwarp();
    if (lane == 0) mbar_arrive(&ring.empty[cp.stage]);
    if (m.flags & 2u) {
#pragma unroll
      for (int r = 0; r < RI; ++r) {
        if (R.row[r] < g.row_end) {
          uint32_t run = 0;
          for (int b = 0; b < nb; ++b) {
            run += hist[b * ROWS_PER_CTA + r * N_CONSUMERS + tid];
            if (run) atomicAdd(a.cnt + (size_t) b * a.ld_cnt + (R.row[r] - g.row_begin), run);
          }
        }
      }
    }
    cp.advance();
  }
  st.flush(g);
}

// ================================================================================================
// nearest neighbours (rows and columns in free-energy-sorted order)
// ================================================================================================
struct NnArgs {
  ScanGeom g;
  const uint32_t* perm;         // [n] sorted position -> original frame
  const uint32_t* lo;           // [n] number of frames with strictly lower free energy than position p
  unsigned long long* key_nn;   // [row_end-row_begin] (d2 bits << 32 | original index), atomicMin'ed
  unsigned long long* key_hd;
};

__host__ __device__ inline size_t screen_smem_bytes(size_t ring_bytes) { return (ring_bytes + 15) & ~size_t(15); }

__host__ __device__ inline size_t nn_smem_bytes(size_t ring_bytes) {
  return ((ring_bytes + 15) & ~size_t(15)) + (size_t) 2 * ROWS_PER_CTA * 8;
}

__device__ __forceinline__ float key_d2(unsigned long long k) { return __uint_as_float((uint32_t) (k >> 32)); }

template <int D>
__global__ void DCB_LAUNCH_BOUNDS(D) nn_kernel(const __grid_constant__ NnArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  const ScanGeom& g = a.g;
  const int d = D ? D : g.d;
  SmemRing<D> ring(smem, d);
  // best[0][slot] = nearest-neighbour key, best[1][slot] = nearest neighbour with lower free energy
  unsigned long long* best = reinterpret_cast<unsigned long long*>(smem + ((SmemRing<D>::bytes(d) + 15) & ~size_t(15)));
  ring.init();
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == N_CONSUMER_WARPS) {
    if (lane == 0) produce<D>(g, ring, [](uint32_t, uint32_t&, uint32_t&) {});
    return;
  }
  const int tid = threadIdx.x;
  Rows<D> R;
  float t[RI], t_nn[RI], t_hd[RI];
  uint32_t lo_max = 0;
  Pipe cp;
  SlowStats st;
  // every column whose exact d2 is <= `d2` satisfies acc < thr(d2) (api.cu: error_bounds)
  auto thr = [&](float d2, float xnr) { return next_up(next_up(fmaf(g.e_rel, d2, d2) + g.e_abs - xnr)); };
  for (;;) {
    mbar_wait(&ring.full[cp.stage], cp.phase);
    const TileMeta m = ring.meta[cp.stage];
    if (m.row_block < 0) break;
    if (m.flags & 1u) {
      R.load(g, (uint32_t) m.row_block, tid);
#pragma unroll
      for (int r = 0; r < RI; ++r) {
        unsigned long long k0 = ~0ull, k1 = ~0ull;
        if (R.row[r] < g.row_end) {                     // warm start from what other items already found
          k0 = a.key_nn[R.row[r] - g.row_begin];
          k1 = a.key_hd[R.row[r] - g.row_begin];
        }
        best[r * N_CONSUMERS + tid] = k0;
        best[ROWS_PER_CTA + r * N_CONSUMERS + tid] = k1;
        t_nn[r] = thr(key_d2(k0), R.xn[r]);
        t_hd[r] = thr(key_d2(k1), R.xn[r]);
      }
      const uint32_t rb0 = g.row_begin + (uint32_t) m.row_block * ROWS_PER_CTA;
      lo_max = __ldg(a.lo + min(rb0 + ROWS_PER_CTA - 1, min(g.row_end, g.n) - 1));
    }
    // Tile class (free energies ascend with the position): no column of the tile has a lower free
    // energy than any row of the block -> the plain nearest-neighbour threshold filters; otherwise
    // the weaker of the two thresholds does, and the slow path sorts the pair out.
    const bool none_hd = m.col0 >= lo_max;
#pragma unroll
    for (int r = 0; r < RI; ++r) t[r] = none_hd ? t_nn[r] : fmaxf(t_nn[r], t_hd[r]);
    const float* tl = ring.tiles + cp.stage * ring.tile_floats;
    scan_tile(g, tl, R, t, [&](int r, int jt, float) {
      const uint32_t j = m.col0 + jt;
      ++st.slow;
      if (j == R.row[r] || j >= g.n || R.row[r] >= g.row_end) return;
      const float d2 = dist2_exact(g.xT, g.ld, d, R.row[r], j);
      ++st.exact;
      if (!(d2 < FLT_MAX)) return;
      const unsigned long long key = ((unsigned long long) __float_as_uint(d2) << 32) | __ldg(a.perm + j);
      unsigned long long* b0 = best + r * N_CONSUMERS + tid;
      unsigned long long* b1 = b0 + ROWS_PER_CTA;
      if (key < *b0) { *b0 = key; t_nn[r] = thr(d2, R.xn[r]); }
      if (j < __ldg(a.lo + R.row[r]) && key < *b1) { *b1 = key; t_hd[r] = thr(d2, R.xn[r]); }
      t[r] = none_hd ? t_nn[r] : fmaxf(t_nn[r], t_hd[r]);
    });
    __syncwarp();
    if (lane == 0) mbar_arrive(&ring.empty[cp.stage]);
    if (m.flags & 2u) {
#pragma unroll
      for (int r = 0; r < RI; ++r) {
        if (R.row[r] < g.row_end) {
          atomicMin(a.key_nn + (R.row[r] - g.row_begin), best[r * N_CONSUMERS + tid]);
          atomicMin(a.key_hd + (R.row[r] - g.row_begin), best[ROWS_PER_CTA + r * N_CONSUMERS + tid]);
        }
      }
    }
    cp.advance();
  }
  st.flush(g);
}

// ================================================================================================
// screening: edges {i new, j < i : d2(i,j) < cut} of the free-energy-sorted frames -> union-find
// ================================================================================================
struct ScreenArgs {
  ScanGeom g;               // rows [row_begin,row_end) = the new sorted positions of this shard
  float cut;                // (float)(4*sigma2); an edge needs d2 < cut (density_clustering.cpp:319)
  float thr_fast;           // cut + error margin
  uint32_t* parent;         // [m_new] union-find forest, parent[p] <= p, roots are the smallest position
};

__device__ __forceinline__ uint32_t uf_find(uint32_t* parent, uint32_t x) {
  for (;;) {
    const uint32_t p = *reinterpret_cast<volatile uint32_t*>(parent + x);
    if (p == x) return x;
    x = p;
  }
}
// lock-free union keeping the smaller position as the root
__device__ __forceinline__ void uf_union(uint32_t* parent, uint32_t a, uint32_t b) {
  for (;;) {
    a = uf_find(parent, a);
    b = uf_find(parent, b);
    if (a == b) return;
    if (a < b) { const uint32_t tmp = a; a = b; b = tmp; }
    const uint32_t old = atomicCAS(parent + a, a, b);      // a > b: hang a below b if a is still a root
    if (old == a) return;
  }
}

template <int D>
__global__ void DCB_LAUNCH_BOUNDS(D) screen_kernel(const __grid_constant__ ScreenArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  constexpr int TJ = TileW<D>::tj;
  const ScanGeom& g = a.g;
  const int d = D ? D : g.d;
  SmemRing<D> ring(smem, d);
  ring.init();
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == N_CONSUMER_WARPS) {
    if (lane == 0)
      produce<D>(g, ring, [&](uint32_t rb, uint32_t&, uint32_t& lim1) {
        // only columns below the last row of the block can form an edge (j < i)
        const uint32_t last_row = min(g.row_begin + (rb + 1) * ROWS_PER_CTA, g.row_end) - 1;
        lim1 = min(lim1, last_row / TJ + 1);
      });
    return;
  }
  const int tid = threadIdx.x;
  Rows<D> R;
  float t[RI];
  Pipe cp;
  SlowStats st;
  for (;;) {
    mbar_wait(&ring.full[cp.stage], cp.phase);
    const TileMeta m = ring.meta[cp.stage];
    if (m.row_block < 0) break;
    if (m.flags & 1u) {
      R.load(g, (uint32_t) m.row_block, tid);
#pragma unroll
      for (int r = 0; r < RI; ++r) t[r] = next_up(a.thr_fast - R.xn[r]);
    }
    const float* tl = ring.tiles + cp.stage * ring.tile_floats;
    scan_tile(g, tl, R, t, [&](int r, int jt, float accv) {
      const uint32_t j = m.col0 + jt;
      const uint32_t i = R.row[r];
      if (j >= i || i >= g.row_end) return;
      ++st.slow;
      float s = accv + R.xn[r];
      const float e = fmaf(g.e_rel, fabsf(s), g.e_abs);
      if (fabsf(s - a.cut) <= e) {
        s = dist2_exact(g.xT, g.ld, d, i, j);
        ++st.exact;
      }
      if (s < a.cut) uf_union(a.parent, i, j);
    });
    __syncwarp();
    if (lane == 0) mbar_arrive(&ring.empty[cp.stage]);
    cp.advance();
  }
  st.flush(g);
}

}  // namespace dcb
