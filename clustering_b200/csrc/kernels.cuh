// The pair-scan kernels of the `clustering density` hot path, written for sm_100a:
//
//   pops_count_kernel<D,NB>  neighbourhood populations for up to 3 distinct radii (branch-free sign-bit counting)
//   pops_bin_kernel<D>       ... for 4 to 31 radii per pass: table-driven binning of the fast squared distance
//   pops_kernel<D>           ... histogram fallback (radius lists no table can serve, shards off the 128-row grid)
//                            (replace reference density_clustering.cpp:126-195 / density_clustering_cuda_kernels.cu:9-56)
//   nn_kernel<D>             nearest neighbour + nearest neighbour with lower free energy, fused
//                            (replaces density_clustering.cpp:230-288 / density_clustering_cuda_kernels.cu:58-130)
//   edge_kernel<D>           all pairs within the screening cut of the spatially ordered frames: the neighbour graph from
//                            which api.cu derives the clusters of EVERY threshold (union-find over the edges by level)
//   screen_kernel<D>         edge discovery + union-find of ONE threshold on free-energy-sorted frames (session API)
//                            (replace density_clustering_common.cpp:37-134 / density_clustering_cuda_kernels.cu:132-192)
//
// Common structure (not derived from the reference's kernels):
//   * persistent CTAs, 8 consumer warps + 1 producer warp; work items (row block x column range) are
//     handed out by an atomic counter, so dense and sparse regions balance dynamically;
//   * the producer prunes (super-tiles, then tiles, against the row block's / row groups' boxes and spheres) and streams
//     the surviving column tiles [D+1][TJ] (dim-major, one contiguous record per tile) into a 6-stage shared memory ring
//     with 1-D bulk TMA (cp.async.bulk + mbarrier complete_tx); stages come back through named barriers (bar.arrive /
//     bar.sync), so a producer whose ring is full is parked; no __syncthreads in steady state;
//   * every consumer thread keeps RI=4 rows in registers and sweeps CJ=4 columns per step with one broadcast LDS.128 per
//     dim and RI*CJ/2 packed FMAs (FFMA2: rows r, r+1 x one broadcast column) -- for D >= 9 the table-driven and the
//     neighbour kernel use 2 rows x 8 columns instead: acc = |y|^2 - 2 x.y (the column pack holds -2y and |y|^2
//     of coordinates taken relative to the TILE's own centre; the rows are re-centred on that point at
//     every tile, so operands are as small as the tile's neighbourhood and the rounding error of the
//     expanded form stays ~1e-6 relative at the decision boundary);
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
  const float* cT;          // [n_col_tiles][(D+1)*TJ + dp] tile-major column pack, one contiguous record per tile (= one bulk
                            //           copy): with c_t the centre of the tile and y' = y - c_t: rows 0..D-1 = -2*y' [TJ],
                            //           row D = |y'|^2 [TJ] (padding: 0, ..., 0, +inf), then the tile's header (tcen below)
  const float* xrow;        // optional extra per-column row [ld] streamed with every tile (neighbour search: free-energy
                            // ranks as floats), nullptr = none; the kernel's SmemRing must be built with xrows = 1
                            // header [dp]: centre c_t[0..D-1], max |y'|^2 over the tile, then the tile's bounding
                            // box lo[0..D-1], hi[0..D-1] in globally centred coordinates (x - centre)
  int dp;                   // floats per header (multiple of 4, >= 3D+1)
  const float* centre;      // [D] global centre the bounding boxes refer to
  size_t ld;                // padded frame count (multiple of 256)
  int d;                    // n_cols (== D for the specialised kernels)
  uint32_t n;               // real frame count
  uint32_t row_begin, row_end;      // rows of this shard
  uint32_t rb_stride;               // row block rb of the launch starts at row_begin + rb * rb_stride * ROWS_PER_CTA: 1 = the contiguous
                                    // range [row_begin,row_end); G = block-cyclic shard of G (rank g: row_begin = g * ROWS_PER_CTA,
                                    // row_end = n), which spreads dense and sparse regions of the spatial order over all ranks
  uint32_t n_row_blocks;            // row blocks of the launch (blocks that start below row_end)
  uint32_t n_col_tiles;             // column tiles of this launch
  uint32_t tiles_per_item;          // column tiles per work item
  uint32_t n_col_items;             // ceil(n_col_tiles / tiles_per_item)
  unsigned int* work_counter;       // zero-initialised before the launch
  float c_loc, e_rel;               // |fast - exact| <= c_loc * (|x'|^2 + max|y'|^2) + e_rel * value   (api.cu: error_bounds)
  float prune_slack;                // absolute slack of the bounding-box arithmetic (globally centred coordinates)
  unsigned long long* stats;        // [0] pairs handed to the slow path, [1] pairs re-evaluated exactly,
                                    // [2] column tiles streamed, [3] (warp, tile) scans (x 32*RI rows x tile width = pairs evaluated)
  const float* bbox;                // [ld/64][2*d] bounding boxes (lo[d], hi[d]) of 64-frame groups, centred coords
  const float* sbbox;               // [ceil(ld/SUPER_FRAMES)][2*d] boxes of the super-tiles, nullptr = no coarse level
  const float* rbbox;               // [n_row_blocks][2*d] bounding boxes of this launch's row blocks (api.cu: row_bbox_kernel)
  float* blk_thr;                   // neighbour search only: [n_row_blocks][N_CONSUMER_WARPS] upper bounds (d2 units) of what the
                                    // rows of a block owned by one consumer warp still accept; lowered by atomicMin as items finish
  float prune_thr;                  // static pruning threshold (fast-value units); +inf disables pruning
  int axis_prune;                   // (row group, tile) units of the count / neighbour kernels also pass a separating-axis test
};

// first row of row block rb of the launch, and the index of a row of that block in the launch's (compact) output arrays
__device__ __forceinline__ uint32_t block_row0(const ScanGeom& g, uint32_t rb) { return g.row_begin + rb * g.rb_stride * (uint32_t) ROWS_PER_CTA; }
__device__ __forceinline__ uint32_t out_index(const ScanGeom& g, uint32_t rb, uint32_t row) {
  return rb * (uint32_t) ROWS_PER_CTA + (row - block_row0(g, rb));
}

// per-thread slow-path counters, added to g.stats once per kernel
struct SlowStats {
  uint32_t slow = 0, exact = 0;
  uint32_t wtiles = 0;          // tiles this warp really scanned (warp-uniform)
  __device__ __forceinline__ void flush(const ScanGeom& g) {
    for (int o = 16; o > 0; o >>= 1) {
      slow += __shfl_xor_sync(0xffffffffu, slow, o);
      exact += __shfl_xor_sync(0xffffffffu, exact, o);
    }
    if ((threadIdx.x & 31) == 0 && g.stats) {
      if (slow) atomicAdd(g.stats, (unsigned long long) slow);
      if (exact) atomicAdd(g.stats + 1, (unsigned long long) exact);
      if (wtiles) atomicAdd(g.stats + 3, (unsigned long long) wtiles);
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
  static constexpr int STAGES = StagesOf<D>::n;
  float* tiles;            // STAGES * ((d+1) * TJ + dp + xrows * TJ): column pack of the tile, its centre entry (tcen),
                           // then the optional extra row
  uint64_t* full;          // STAGES
  uint64_t* empty;         // STAGES (unused: stages are handed back through named barriers, common.cuh: stage_release)
  TileMeta* meta;          // STAGES
  unsigned long long* wthr;   // N_CONSUMER_WARPS: (item << 32 | float bits) pruning threshold published per row group
  uint32_t* unext;         // STAGES: next (row group, tile) unit of the stage to hand out (dynamic kernels)
  uint32_t* umask;         // STAGES: the stage's unit mask as computed by the first warp that got there (0xffffffff = not yet)
  float* rbb;              // 2*d: bounding box of the current row block (producer scratch)
  size_t tile_floats;
  __host__ __device__ static int dp_of(int d) { return (3 * d + 1 + 3) / 4 * 4; }
  __host__ __device__ static size_t bytes(int d, int xrows = 0) {
    return (size_t) STAGES * ((d + 1 + xrows) * TJ + dp_of(d)) * 4 + 2 * STAGES * 8 + STAGES * sizeof(TileMeta) + N_CONSUMER_WARPS * 8 +
           (size_t) STAGES * 8 + (size_t) 2 * d * 4 + 16;
  }
  __device__ SmemRing(unsigned char* base, int d, int xrows = 0) {
    tile_floats = (size_t) (d + 1 + xrows) * TJ + dp_of(d);
    tiles = reinterpret_cast<float*>(base);
    full = reinterpret_cast<uint64_t*>(base + STAGES * tile_floats * 4);
    empty = full + STAGES;
    meta = reinterpret_cast<TileMeta*>(empty + STAGES);
    wthr = reinterpret_cast<unsigned long long*>(meta + STAGES);
    unext = reinterpret_cast<uint32_t*>(wthr + N_CONSUMER_WARPS);
    umask = unext + STAGES;
    rbb = reinterpret_cast<float*>(umask + STAGES);
  }
  __device__ void init() {
    if (threadIdx.x == 0) {
      for (int s = 0; s < STAGES; ++s) {
        mbar_init(&full[s], 1);
      }
      for (int w = 0; w < N_CONSUMER_WARPS; ++w) wthr[w] = ~0ull;
      fence_mbar_init();
    }
  }
};

// Work item -> (row block, column range).  Items are ordered column-step-major: at step 0 every row
// block scans the column range that contains its own frames (spatial order: its near neighbourhood),
// then ranges at growing distance, alternating sides.  Concurrently running items therefore belong
// to different row blocks, and a row block's later items start from what its earlier ones found.
__device__ __forceinline__ void item_coords(const ScanGeom& g, uint32_t item, uint32_t TJ, uint32_t* rb, uint32_t* ci) {
  const uint32_t step = item / g.n_row_blocks;
  *rb = item % g.n_row_blocks;
  const uint32_t diag = min(block_row0(g, *rb) / (g.tiles_per_item * TJ), g.n_col_items - 1);
  // offsets 0, +1, -1, +2, -2, ... folded into [0, n_col_items)
  const uint32_t k = (step + 1) >> 1;
  const uint32_t n = g.n_col_items;
  uint32_t c;
  if (step & 1) c = (diag + k) % n; else c = (diag + n - (k % n)) % n;
  // the alternating walk visits every range exactly once only if it is made a permutation:
  // walk positions p = 0..n-1 are mapped through the cyclic order starting at diag
  *ci = c;
}

// Coarse level of the producers' pruning.  The tile tests cost one round trip to L2 per 32 tiles, and a scan makes
// N^2 / (1024 x 128) of them: at 5M frames that is 1.9e8 tests, ~54 000 cycles per work item in which the consumers of an
// item with nothing in reach just wait.  Super-tiles (SUPER_FRAMES consecutive positions of the spatial order, boxes in
// g.sbbox) are tested first, 32 per round trip: lane l of `super_mask` looks at super-tile s0 + l against the row block's
// box (ring.rbb).  Bit l of the result: the super-tile may hold a column within thr of a row of the block.
template <int D>
__device__ __forceinline__ uint32_t super_mask(const ScanGeom& g, const float* __restrict__ rbb, uint32_t s0, uint32_t s_end, float thr, int lane) {
  const int d = D ? D : g.d;
  const uint32_t sp = s0 + (uint32_t) lane;
  bool keep = false;
  if (sp < s_end) {
    const float* sb = g.sbbox + (size_t) sp * 2 * d;
    float acc = 0.f;
    for (int k = 0; k < d; ++k) {
      const float gap = fmaxf(fmaxf(rbb[k] - __ldg(sb + d + k), __ldg(sb + k) - rbb[d + k]), 0.f);
      acc = fmaf(gap, gap, acc);
    }
    keep = !(acc * 0.999f > thr);                   // NaN keeps
  }
  return __ballot_sync(0xffffffffu, keep);
}

// Producer warp: walks the work items, drops the column tiles whose bounding box is provably out of
// reach of every row of the block (dynamic threshold published by the consumers, or the static one),
// and streams the others with 1-D bulk TMA.  An item that streamed at least one tile is closed by a
// data-less end marker (flags 2|4) on which the consumers write their results back.
// item_thr(rb, lane): warp-collective, returns the (uniform) pruning threshold the item starts with.
template <int D, class TileRange, class ItemThr>
__device__ __forceinline__ void produce(const ScanGeom& g, SmemRing<D>& ring, bool dynamic_thr, TileRange&& range, ItemThr&& item_thr) {
  constexpr int TJ = TileW<D>::tj;
  constexpr int GPT = TJ / 64;                      // 64-frame bounding-box groups per tile
  const int lane = threadIdx.x & 31;
  FillPipe<StagesOf<D>::n> pp;
  const int d = D ? D : g.d;
  const uint32_t total = g.n_row_blocks * g.n_col_items;
  unsigned long long streamed = 0;
  for (;;) {
    uint32_t item = 0;
    if (lane == 0) item = atomicAdd(g.work_counter, 1u);
    item = __shfl_sync(0xffffffffu, item, 0);
    if (item >= total) break;
    uint32_t rb, ci;
    item_coords(g, item, TJ, &rb, &ci);
    uint32_t t0 = ci * g.tiles_per_item;
    uint32_t t1 = min(t0 + g.tiles_per_item, g.n_col_tiles);
    uint32_t lim0 = 0, lim1 = g.n_col_tiles;
    range(rb, lim0, lim1);
    t0 = max(t0, lim0);
    t1 = min(t1, lim1);
    if (t0 >= t1) continue;
    // bounding box of the row block (precomputed per launch) and the pruning threshold the item starts with
    for (int k = lane; k < 2 * d; k += 32) ring.rbb[k] = __ldg(g.rbbox + (size_t) rb * 2 * d + k);
    const float thr0 = item_thr(rb, lane);
    __syncwarp();
    bool first = true;
    constexpr uint32_t TPS = SUPER_FRAMES / TJ;     // tiles per super-tile
    for (uint32_t s0 = t0 / TPS; s0 * TPS < t1; s0 += 32) {
     const uint32_t s_end = (t1 + TPS - 1) / TPS;
     uint32_t smask = g.sbbox ? super_mask<D>(g, ring.rbb, s0, s_end, thr0, lane) : 0xffffffffu;
     while (smask) {
      const uint32_t sp = s0 + (uint32_t) __ffs(smask) - 1u;
      smask &= smask - 1;
      if (sp >= s_end) break;
      const uint32_t seg0 = max(t0, sp * TPS), seg1 = min(t1, (sp + 1) * TPS);
    for (uint32_t base = seg0; base < seg1; base += 32) {
      const uint32_t t = base + lane;
      float lb = INFINITY;                          // lower bound of the fast value over (row block) x (tile t)
      if (t < seg1) {
        for (int q = 0; q < GPT; ++q) {
          const float* bb = g.bbox + (size_t) (t * GPT + q) * 2 * d;
          float s = 0.f;
          for (int k = 0; k < d; ++k) {
            const float gap = fmaxf(fmaxf(ring.rbb[k] - __ldg(bb + d + k), __ldg(bb + k) - ring.rbb[d + k]), 0.f);
            s = fmaf(gap, gap, s);
          }
          lb = q == 0 ? s : fminf(lb, s);
        }
        lb *= 0.999f;
      }
      uint32_t mask = __ballot_sync(0xffffffffu, t < seg1 && !(lb > thr0));     // NaN keeps
      while (mask) {
        const int src = __ffs(mask) - 1;
        const uint32_t tt = base + (uint32_t) src;
        mask &= mask - 1;
        if (dynamic_thr) {
          // the consumers tighten the threshold while they work: look again right before streaming
          float thr = 0.f;
          for (int w = 0; w < N_CONSUMER_WARPS; ++w) {
            const unsigned long long v = *reinterpret_cast<volatile unsigned long long*>(ring.wthr + w);
            // a value published for another item says nothing about this one
            thr = fmaxf(thr, (uint32_t) (v >> 32) == item ? __uint_as_float((uint32_t) v) : INFINITY);
          }
          // one lane decides for the warp: the stage hand-back below is a warp-wide barrier instruction
          if (__shfl_sync(0xffffffffu, (int) (__shfl_sync(0xffffffffu, lb, src) > thr), 0)) continue;
        }
        pp.acquire();
        if (lane == 0) {
          TileMeta m;
          m.row_block = (int32_t) rb;
          m.col0 = tt * TJ;
          m.flags = first ? 1u : 0u;
          m.aux = item;
          ring.meta[pp.stage] = m;
          ring.unext[pp.stage] = 0;
          ring.umask[pp.stage] = 0xffffffffu;
          mbar_arrive_expect_tx(&ring.full[pp.stage], (uint32_t) (ring.tile_floats * 4));
        }
        __syncwarp();
        {
          // one bulk copy for the tile's record (pack + header), one more for the optional extra row
          float* dst = ring.tiles + pp.stage * ring.tile_floats;
          const uint32_t rec = (uint32_t) ((d + 1) * TJ + g.dp);
          if (lane == 0) tma_load_1d(dst, g.cT + (size_t) tt * rec, rec * 4, &ring.full[pp.stage]);
          if (lane == 1 && g.xrow) tma_load_1d(dst + rec, g.xrow + (size_t) tt * TJ, TJ * 4, &ring.full[pp.stage]);
        }
        first = false;
        ++streamed;
        pp.advance();
      }
    }
     }
    }
    if (!first) {                                 // end-of-item marker
      __syncwarp();
      pp.acquire();
      if (lane == 0) {
        TileMeta m;
        m.row_block = (int32_t) rb;
        m.col0 = 0;
        m.flags = 2u | 4u;
        m.aux = item;
        ring.meta[pp.stage] = m;
        mbar_arrive(&ring.full[pp.stage]);
      }
      pp.advance();
    }
  }
  __syncwarp();
  pp.acquire();
  if (lane == 0) {
    ring.meta[pp.stage].row_block = -1;
    mbar_arrive(&ring.full[pp.stage]);
    if (g.stats && streamed) atomicAdd(g.stats + 2, streamed);
  }
  __syncwarp();
  pp.drain();
}

// Row operands of a consumer thread: rows row0 + 32 r, r = 0..RI-1, so that a warp owns 32*RI = 128 CONSECUTIVE rows of the
// block (a compact piece of the spatial order: WarpBox below skips whole tiles for it).  retarget() re-centres them
// on the centre of the tile about to be scanned (cen: the tile's tcen entry in shared memory):
//   x[r][k] = x_original - c_t[k],  xn[r] = |x'|^2,  eabs[r] = c_loc * (xn[r] + max|y'|^2 of the tile),
// the absolute part of the fast path's rounding-error bound for this row against this tile.
// The original coordinates are re-read from xT (coalesced, L1 resident after the first tile of an item).
// register arrays must not be indexed dynamically (that would spill them to local memory)
__device__ __forceinline__ float sel4(const float (&v)[RI], int r) {
  return r == 0 ? v[0] : r == 1 ? v[1] : r == 2 ? v[2] : v[3];
}
__device__ __forceinline__ void put4(float (&v)[RI], int r, float x) {
  if (r == 0) v[0] = x; else if (r == 1) v[1] = x; else if (r == 2) v[2] = x; else v[3] = x;
}

// Registers are the scarce resource of the specialised kernels (96 per thread at two CTAs per SM, 4 D of them row
// operands): whatever can be recomputed in a few instructions outside the inner loop -- the clamped positions, the absolute
// error part -- is a function, not a member, so that nothing of the inner loop's operands is spilled.
template <int D>
struct Rows {
  float x[RI][D];           // tile-local coordinates x'
  float xn[RI];             // |x'|^2
  float ymax;               // max |y'|^2 of the tile the rows are centred on
  uint32_t row0;
  uint32_t out0;            // index of row0 in the launch's output arrays (out_index)
  uint32_t stride;          // 32: the warp owns 128 consecutive rows; N_CONSUMERS: rows interleaved over the whole block
  __device__ __forceinline__ uint32_t row(int r) const { return row0 + (uint32_t) r * stride; }
  __device__ __forceinline__ uint32_t out(int r) const { return out0 + (uint32_t) r * stride; }
  // clamped position: rows past the end read the last padded frame, their results are discarded
  __device__ __forceinline__ uint32_t pos(const ScanGeom& g, int r) const { return min(row(r), (uint32_t) g.ld - 1u); }
  // absolute part of the fast path's error bound of row r against the current tile
  __device__ __forceinline__ float ea(const ScanGeom& g, int r) const { return g.c_loc * (sel4(xn, r) + ymax); }
  __device__ __forceinline__ void load(const ScanGeom& g, uint32_t rb, int tid, bool coherent = true) {
    stride = coherent ? 32u : (uint32_t) N_CONSUMERS;
    row0 = block_row0(g, rb) + (coherent ? (uint32_t) (tid >> 5) * (32u * RI) + (uint32_t) (tid & 31) : (uint32_t) tid);
    out0 = rb * (uint32_t) ROWS_PER_CTA + (row0 - block_row0(g, rb));
  }
  // rows of row group gi (128 consecutive rows of the block), whichever warp works on them
  __device__ __forceinline__ void load_group(const ScanGeom& g, uint32_t rb, uint32_t gi, int lane) {
    stride = 32u;
    row0 = block_row0(g, rb) + gi * (32u * RI) + (uint32_t) lane;
    out0 = rb * (uint32_t) ROWS_PER_CTA + gi * (32u * RI) + (uint32_t) lane;
  }
  __device__ __forceinline__ void retarget(const ScanGeom& g, const float* __restrict__ cen) {
    float c[D];
#pragma unroll
    for (int k = 0; k < D; ++k) c[k] = cen[k];
    ymax = cen[D];
#pragma unroll
    for (int r = 0; r < RI; ++r) {
      const uint32_t pr = pos(g, r);
      float s = 0.f;
#pragma unroll
      for (int k = 0; k < D; ++k) {
        x[r][k] = __ldg(g.xT + (size_t) k * g.ld + pr) - c[k];
        s = fmaf(x[r][k], x[r][k], s);
      }
      xn[r] = s;
    }
  }
};
template <>
struct Rows<0> {
  float xn[RI];
  float ymax;
  uint32_t row0;
  uint32_t out0;
  uint32_t p[RI];
  uint32_t stride;
  __device__ __forceinline__ uint32_t row(int r) const { return row0 + (uint32_t) r * stride; }
  __device__ __forceinline__ uint32_t out(int r) const { return out0 + (uint32_t) r * stride; }
  __device__ __forceinline__ uint32_t pos(const ScanGeom&, int r) const { return p[r]; }
  __device__ __forceinline__ float ea(const ScanGeom& g, int r) const { return g.c_loc * (sel4(xn, r) + ymax); }
  __device__ __forceinline__ void load(const ScanGeom& g, uint32_t rb, int tid, bool coherent = true) {
    stride = coherent ? 32u : (uint32_t) N_CONSUMERS;
    row0 = block_row0(g, rb) + (coherent ? (uint32_t) (tid >> 5) * (32u * RI) + (uint32_t) (tid & 31) : (uint32_t) tid);
    out0 = rb * (uint32_t) ROWS_PER_CTA + (row0 - block_row0(g, rb));
#pragma unroll
    for (int r = 0; r < RI; ++r) p[r] = (uint32_t) min((size_t) row(r), g.ld - 1);
  }
  __device__ __forceinline__ void load_group(const ScanGeom& g, uint32_t rb, uint32_t gi, int lane) {
    stride = 32u;
    row0 = block_row0(g, rb) + gi * (32u * RI) + (uint32_t) lane;
    out0 = rb * (uint32_t) ROWS_PER_CTA + gi * (32u * RI) + (uint32_t) lane;
#pragma unroll
    for (int r = 0; r < RI; ++r) p[r] = (uint32_t) min((size_t) row(r), g.ld - 1);
  }
  __device__ __forceinline__ void retarget(const ScanGeom& g, const float* __restrict__ cen) {
    float s[RI];
#pragma unroll
    for (int r = 0; r < RI; ++r) s[r] = 0.f;
    for (int k = 0; k < g.d; ++k) {
      const float c = cen[k];
#pragma unroll
      for (int r = 0; r < RI; ++r) {
        const float v = __ldg(g.xT + (size_t) k * g.ld + p[r]) - c;
        s[r] = fmaf(v, v, s[r]);
      }
    }
    ymax = cen[g.d];
#pragma unroll
    for (int r = 0; r < RI; ++r) xn[r] = s[r];
  }
};

// The fast value of one RI x CJ block: acc[r][c] = |y'_c|^2 - 2 x'_r . y'_c for columns g .. g+CJ-1 of the tile at tl
// (tlg = tl + g; dim-major rows of TJ floats, row D = |y'|^2): one broadcast LDS.128 per dim.
// sm_100 has a packed FP32 FMA (fma.rn.f32x2, SASS FFMA2: two IEEE FMAs -- the same bits as two FFMAs -- whose second
// source may be one scalar broadcast to both halves): rows r, r+1 share an instruction, the accumulators of the two rows
// sit in an aligned register pair.  Same FMA-pipe time, half the issue slots (scripts/micro/ffma2.cu).
#ifndef DCB_FFMA2
#define DCB_FFMA2 1
#endif
__device__ __forceinline__ unsigned long long pack2(float lo, float hi) {
  unsigned long long v;
  asm("mov.b64 %0, {%1, %2};" : "=l"(v) : "f"(lo), "f"(hi));
  return v;
}
__device__ __forceinline__ void unpack2(unsigned long long v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
template <int D>
__device__ __forceinline__ void fast_block(const float* __restrict__ tlg, const Rows<D>& R, float (&acc)[RI][CJ]) {
  constexpr int TJ = TileW<D>::tj;
#if DCB_FFMA2
  static_assert(RI == 4 && CJ == 4, "row pairs (0,1) and (2,3)");
  unsigned long long a2[RI / 2][CJ];
  {
    const float4 n4 = *reinterpret_cast<const float4*>(tlg + D * TJ);
    const float4 y4 = *reinterpret_cast<const float4*>(tlg);
    const float yc[CJ] = {y4.x, y4.y, y4.z, y4.w}, nc[CJ] = {n4.x, n4.y, n4.z, n4.w};
#pragma unroll
    for (int h = 0; h < RI / 2; ++h)
#pragma unroll
      for (int c = 0; c < CJ; ++c) a2[h][c] = fma2(pack2(R.x[2 * h][0], R.x[2 * h + 1][0]), pack2(yc[c], yc[c]), pack2(nc[c], nc[c]));
  }
#pragma unroll
  for (int k = 1; k < D; ++k) {
    const float4 y4 = *reinterpret_cast<const float4*>(tlg + k * TJ);
    const float yc[CJ] = {y4.x, y4.y, y4.z, y4.w};
#pragma unroll
    for (int h = 0; h < RI / 2; ++h)
#pragma unroll
      for (int c = 0; c < CJ; ++c) a2[h][c] = fma2(pack2(R.x[2 * h][k], R.x[2 * h + 1][k]), pack2(yc[c], yc[c]), a2[h][c]);
  }
#pragma unroll
  for (int h = 0; h < RI / 2; ++h)
#pragma unroll
    for (int c = 0; c < CJ; ++c) unpack2(a2[h][c], acc[2 * h][c], acc[2 * h + 1][c]);
#else
  {
    const float4 n4 = *reinterpret_cast<const float4*>(tlg + D * TJ);
    const float4 y4 = *reinterpret_cast<const float4*>(tlg);
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
    const float4 y4 = *reinterpret_cast<const float4*>(tlg + k * TJ);
#pragma unroll
    for (int r = 0; r < RI; ++r) {
      acc[r][0] = fmaf(R.x[r][k], y4.x, acc[r][0]);
      acc[r][1] = fmaf(R.x[r][k], y4.y, acc[r][1]);
      acc[r][2] = fmaf(R.x[r][k], y4.z, acc[r][2]);
      acc[r][3] = fmaf(R.x[r][k], y4.w, acc[r][3]);
    }
  }
#endif
}

// Bounding box of the 128 rows a consumer warp owns (globally centred coordinates, the same arithmetic as the tile
// boxes of api.cu: pack_tiles_kernel); lane k holds dimension k.  reach() is the warp-level version of the producer's
// tile pruning: a streamed tile is scanned by a warp only if its box comes within the warp's own threshold.
// Specialised kernels only (D <= 16 <= 32 lanes); the run-time-D kernels scan every streamed tile.
template <int D>
struct WarpBox {
  float wlo, whi;
  __device__ __forceinline__ void compute(const ScanGeom& g, const Rows<D>& R, int lane) {
    wlo = INFINITY;
    whi = -INFINITY;
#pragma unroll
    for (int k = 0; k < D; ++k) {
      const float ck = __ldg(g.centre + k);
      float lo = INFINITY, hi = -INFINITY;
#pragma unroll
      for (int r = 0; r < RI; ++r) {
        if (R.row(r) < g.row_end) {
          const float v = __ldg(g.xT + (size_t) k * g.ld + R.pos(g, r)) - ck;
          lo = fminf(lo, v);
          hi = fmaxf(hi, v);
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
      }
      if (lane == k) { wlo = lo; whi = hi; }
    }
  }
  // cen: the tile's header in shared memory.  false => no pair of (this warp's rows) x (tile) is closer than thr (d2 units)
  __device__ __forceinline__ bool reach(const float* __restrict__ cen, int lane, float thr) const {
    float s = 0.f;
    if (lane < D) {
      const float gap = fmaxf(fmaxf(wlo - cen[2 * D + 1 + lane], cen[D + 1 + lane] - whi), 0.f);
      s = gap * gap;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    return !(s * 0.999f > thr);          // NaN keeps
  }
};
template <>
struct WarpBox<0> {
  float wlo = 0.f, whi = 0.f;
  __device__ __forceinline__ void compute(const ScanGeom&, const Rows<0>&, int) {}
  __device__ __forceinline__ bool reach(const float*, int, float) const { return true; }
};

// Dynamic kernels (populations count mode, neighbour search): the eight row groups of a block are NOT tied to the
// eight consumer warps.  For every streamed tile each warp works out which groups can reach it (their boxes live in
// shared memory) and the warps then claim (group, tile) units from a per-stage counter until none is left; a warp
// whose nearby groups have nothing to do with a tile helps with another group's unit instead of idling, and may run
// up to a ring depth ahead.  GBOX_DIMS floats per bound keep the rows 64 B apart.
constexpr int GBOX_DIMS = 16;
__host__ __device__ inline size_t gbox_bytes() { return (size_t) N_CONSUMER_WARPS * 2 * GBOX_DIMS * 4; }

__device__ __forceinline__ void consumer_barrier() { asm volatile("bar.sync 1, %0;" ::"n"(N_CONSUMERS) : "memory"); }

// bit 4*gi of the result: row group gi comes within thr[gi] (d2 units) of the tile whose header is cen
template <int D>
__device__ __forceinline__ uint32_t groups_in_reach(const float* __restrict__ gbox, const float* __restrict__ cen, int lane, float thr_of_group) {
  if (D == 0) return 0x11111111u;                 // run-time-D kernels keep no boxes: every group scans every streamed tile
  const int gi = lane >> 2, q = lane & 3;
  const float* lo = gbox + gi * 2 * GBOX_DIMS;
  const float* hi = lo + GBOX_DIMS;
  float s = 0.f;
#pragma unroll
  for (int k = q; k < D; k += 4) {
    const float gap = fmaxf(fmaxf(lo[k] - cen[2 * D + 1 + k], cen[D + 1 + k] - hi[k]), 0.f);
    s = fmaf(gap, gap, s);
  }
  s += __shfl_xor_sync(0xffffffffu, s, 1);
  s += __shfl_xor_sync(0xffffffffu, s, 2);
  return __ballot_sync(0xffffffffu, !(s * 0.999f > thr_of_group)) & 0x11111111u;     // NaN keeps
}

// Separating-axis test of a (row group, tile) unit: a lower bound of the distance of every pair of the unit from the
// extents of the rows and of the columns along ONE direction u (any direction gives a valid bound; the line between the
// two centres is the one along which two separate blobs overlap least).  Rows: the thread's RI rows in tile-local
// coordinates x' (Rows::retarget); columns: the tile's -2y' in shared memory (lane l looks at columns 4l .. 4l+3).
// In 10 dimensions the boxes and spheres of two neighbouring clusters overlap while their extents along the line between
// them do not (projection of a Gaussian blob: ~3 sigma, its radius: ~(sqrt(D) + 2) sigma).  ~130 instructions per unit.
// Returns the bound on the DISTANCE (<= 0: none).  slack_len: rounding of the centres (globally centred coordinates).
template <int D>
__device__ __forceinline__ float axis_lower_bound(const ScanGeom& g, const Rows<D>& R, const float* __restrict__ tl, const float (&u)[D],
                                                  float ymax, float slack_len, int lane) {
  constexpr int TJ = TileW<D>::tj;
  static_assert(TJ == 128, "one float4 of columns per lane");
  float un = 0.f;
#pragma unroll
  for (int k = 0; k < D; ++k) un = fmaf(u[k], u[k], un);
  if (!(un > 0.f)) return 0.f;
  const float inv = rsqrtf(un);
  const float ulen = un * inv;
  // rounding of a D-term FFMA chain over x'_k u_k: D ulp of |x'||u|, folded into the projections themselves
  const float eps = (float) D * 2e-7f * ulen;
  float pmin = INFINITY;
#pragma unroll
  for (int r = 0; r < RI; ++r) {
    float pr = 0.f;
#pragma unroll
    for (int k = 0; k < D; ++k) pr = fmaf(R.x[r][k], u[k], pr);
    if (R.row(r) < g.row_end) pmin = fminf(pmin, pr - eps * sqrtf(R.xn[r]));
  }
  float q[CJ] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int k = 0; k < D; ++k) {
    const float4 y4 = *reinterpret_cast<const float4*>(tl + k * TJ + 4 * lane);      // -2 y'
    q[0] = fmaf(y4.x, u[k], q[0]);
    q[1] = fmaf(y4.y, u[k], q[1]);
    q[2] = fmaf(y4.z, u[k], q[2]);
    q[3] = fmaf(y4.w, u[k], q[3]);
  }
  // padded columns hold y' = 0 (the tile's centre = the mean of its frames), inside the span of the real ones: harmless
  float qmax = -0.5f * fminf(fminf(q[0], q[1]), fminf(q[2], q[3])) + eps * sqrtf(ymax);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    pmin = fminf(pmin, __shfl_xor_sync(0xffffffffu, pmin, o));
    qmax = fmaxf(qmax, __shfl_xor_sync(0xffffffffu, qmax, o));
  }
  return (pmin - qmax) * inv * 0.99999f - slack_len;
}


constexpr size_t SCRATCH_BYTES = (size_t) RI * CJ * N_CONSUMERS * 4;     // per-thread spill of one 4x4 block

// Slow path entry: the 4x4 block of one step has at least one pair below its row threshold somewhere
// in the warp.  The block is parked in shared memory (slot-private column => no bank conflicts) and
// the hits of this lane are walked with ONE compact copy of the handler, so the instruction footprint
// of the kernel stays small (the hot loop must live in the instruction cache).
template <class Hit>
__device__ __forceinline__ void walk_hits(float* __restrict__ scratch, const float (&acc)[RI][CJ], float (&t)[RI], int jt0, Hit& hit) {
  uint32_t mask = 0;
#pragma unroll
  for (int r = 0; r < RI; ++r)
#pragma unroll
    for (int c = 0; c < CJ; ++c) {
      scratch[(r * CJ + c) * N_CONSUMERS] = acc[r][c];
      mask |= (acc[r][c] < t[r]) ? (1u << (r * CJ + c)) : 0u;
    }
#pragma unroll 1
  while (mask) {
    const int p = __ffs(mask) - 1;
    mask &= mask - 1;
    const int r = p / CJ;
    const float a = scratch[p * N_CONSUMERS];
    if (a < sel4(t, r)) hit(r, jt0 + (p % CJ), a);        // re-checked: the handler may have tightened t[r]
  }
}

// Consumer inner loop over one tile: acc[r][c] = |y_c|^2 - 2 x_r.y_c for an RI x CJ block per step;
// hit(r, j_in_tile, acc) is called for every pair with acc < t[r].
template <int D, class Hit>
__device__ __forceinline__ void scan_tile(const ScanGeom&, const float* __restrict__ tl, const Rows<D>& R, float (&t)[RI],
                                          float* __restrict__ scratch, Hit& hit) {
  constexpr int TJ = TileW<D>::tj;
#pragma unroll 1
  for (int g = 0; g < TJ; g += CJ) {
    float acc[RI][CJ];
    fast_block<D>(tl + g, R, acc);
    bool any = false;
#pragma unroll
    for (int r = 0; r < RI; ++r) {
      const float m = fminf(fminf(acc[r][0], acc[r][1]), fminf(acc[r][2], acc[r][3]));
      any |= (m < t[r]);
    }
    if (any) walk_hits(scratch, acc, t, g, hit);
  }
}

// run-time-D variant: 4 rows x 16 columns per step, row operands from L1/L2
template <class Hit>
__device__ __forceinline__ void scan_tile(const ScanGeom& gm, const float* __restrict__ tl, const Rows<0>& R, float (&t)[RI],
                                          float* __restrict__ scratch, Hit& hit) {
  constexpr int TJ = TileW<0>::tj, CG = TileW<0>::cj;
  const int d = gm.d;
  const float* __restrict__ cen = tl + (d + 1) * TJ;
#pragma unroll 1
  for (int g = 0; g < TJ; g += CG) {
    float acc[CG / CJ][RI][CJ];
#pragma unroll
    for (int c4 = 0; c4 < CG / CJ; ++c4) {
      const float4 n4 = *reinterpret_cast<const float4*>(tl + d * TJ + g + c4 * CJ);
#pragma unroll
      for (int r = 0; r < RI; ++r) {
        acc[c4][r][0] = n4.x; acc[c4][r][1] = n4.y; acc[c4][r][2] = n4.z; acc[c4][r][3] = n4.w;
      }
    }
#pragma unroll 2
    for (int k = 0; k < d; ++k) {
      float xr[RI];
      const float ck = cen[k];
#pragma unroll
      for (int r = 0; r < RI; ++r) xr[r] = __ldg(gm.xT + (size_t) k * gm.ld + R.p[r]) - ck;
#pragma unroll
      for (int c4 = 0; c4 < CG / CJ; ++c4) {
        const float4 y4 = *reinterpret_cast<const float4*>(tl + k * TJ + g + c4 * CJ);
#pragma unroll
        for (int r = 0; r < RI; ++r) {
          acc[c4][r][0] = fmaf(xr[r], y4.x, acc[c4][r][0]);
          acc[c4][r][1] = fmaf(xr[r], y4.y, acc[c4][r][1]);
          acc[c4][r][2] = fmaf(xr[r], y4.z, acc[c4][r][2]);
          acc[c4][r][3] = fmaf(xr[r], y4.w, acc[c4][r][3]);
        }
      }
    }
#pragma unroll
    for (int c4 = 0; c4 < CG / CJ; ++c4) {
      bool any = false;
#pragma unroll
      for (int r = 0; r < RI; ++r) {
        const float m = fminf(fminf(acc[c4][r][0], acc[c4][r][1]), fminf(acc[c4][r][2], acc[c4][r][3]));
        any |= (m < t[r]);
      }
      if (any) walk_hits(scratch, acc[c4], t, g + c4 * CJ, hit);
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
  float thr_fast;           // rad2[n_bins-1] (1 + e_rel): relative part of the filter margin (absolute part: Rows::eabs)
  float band[8];            // count mode: the part of the error band around rad2[b] that depends on the radius only
  uint32_t* cnt;            // [n_bins][ld_cnt]: #{j != i : rad2[b-1] <= d2(i,j) < rad2[b]}, rows in out_index order
  size_t ld_cnt;
  // bin mode (pops_bin_kernel): cell table over the fast squared distance s, built by api.cu (build_bin_table)
  const float* lut;         // [lut_k] device: per cell the one squared radius B_k it can see with k in its low 5 mantissa bits
                            //         (bin = k + (s >= entry)); cells that see none: +big | 0 when no radius lies below them,
                            //         else -big | (k - 1)
  int lut_k;                // cells, a power of two: cell = floor(sat(s * lut_scale) * (lut_k - 1)) (quarter-cell rounding)
  float lut_scale;          // 1 / (span of the table in s units, a little more than r_max^2)
  float lut_margin;         // the table decides only pairs whose error band is narrower than this (s units); others: handler
  float band_max;           // radius part of the band half width for r_max (covers every radius of the pass) + the table's
                            // own perturbation of the boundaries (32 ulp)
  int dense_lanes;          // a 4x4 step is binned branch-free when at least this many lanes hold a candidate pair,
                            // else candidate by candidate
  int steal;                // a warp whose own row group has nothing to do with a streamed tile takes over another group's unit
  int proj_prune;           // separating-axis test of (row group, tile) units along the line between their centres
};

__host__ __device__ inline size_t pops_smem_bytes(size_t ring_bytes, int n_bins) {
  return ((ring_bytes + 15) & ~size_t(15)) + SCRATCH_BYTES + 32 * 4 + (size_t) n_bins * ROWS_PER_CTA * 2;
}

// number of table entries <= s (table ascending, padded with +inf up to 32 entries)
__device__ __forceinline__ int bin_of(const float* __restrict__ rad2s, int nb, float s) {
  int b = 0;
  if (nb <= 4) {
#pragma unroll
    for (int q = 0; q < 4; ++q) b += (rad2s[q] <= s) ? 1 : 0;
  } else {
#pragma unroll
    for (int step = 16; step >= 1; step >>= 1)
      if (rad2s[b + step - 1] <= s) b += step;
  }
  return b;
}

template <int D>
__global__ void DCB_LAUNCH_BOUNDS(D) pops_kernel(const __grid_constant__ PopsArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  const ScanGeom& g = a.g;
  const int d = D ? D : g.d;
  SmemRing<D> ring(smem, d);
  unsigned char* extra = smem + ((SmemRing<D>::bytes(d) + 15) & ~size_t(15));
  float* scratch = reinterpret_cast<float*>(extra) + threadIdx.x;
  float* rad2s = reinterpret_cast<float*>(extra + SCRATCH_BYTES);
  // [n_bins][ROWS_PER_CTA] slot-private 16-bit counters, flushed per work item (api.cu keeps an item below 65536 columns)
  uint16_t* hist = reinterpret_cast<uint16_t*>(rad2s + 32) + threadIdx.x;
  ring.init();
  if (threadIdx.x < 32) rad2s[threadIdx.x] = a.rad2[threadIdx.x];
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == N_CONSUMER_WARPS) {
    produce<D>(g, ring, false, [](uint32_t, uint32_t&, uint32_t&) {}, [&](uint32_t, int) { return g.prune_thr; });
    return;
  }
  const int tid = threadIdx.x;
  const int nb = a.n_bins;
  Rows<D> R;
  float t[RI];
  Pipe<StagesOf<D>::n> cp;
  SlowStats st;
  uint32_t col0 = 0;
  auto hit = [&](int r, int jt, float accv) {
    const uint32_t j = col0 + jt;
    const uint32_t i = R.row(r);
    float s = accv + sel4(R.xn, r);
    if (i >= g.row_end) return;                // clamped rows past the shard: results are discarded anyway
    ++st.slow;
    int b = bin_of(rad2s, nb, s);
    const float lo = b > 0 ? rad2s[b - 1] : -INFINITY;
    const float hi = rad2s[b];
    const float e = fmaf(g.e_rel, fabsf(s), R.ea(g, r)) * 1.0001f;
    if ((s - lo < e) || (hi - s <= e)) {       // within the error band of a radius: decide exactly
      s = dist2_exact(g.xT, g.ld, d, i, j);
      ++st.exact;
      b = bin_of(rad2s, nb, s);
    }
    if (j != i && b < nb) hist[b * ROWS_PER_CTA + r * N_CONSUMERS] += 1;
  };
  for (;;) {
    mbar_wait(&ring.full[cp.stage], cp.phase);
    const TileMeta m = ring.meta[cp.stage];
    if (m.row_block < 0) break;
    if (m.flags & 1u) {
      // rows interleaved over the block and no warp-level tile skipping here: this kernel's cost is the per-hit handler,
      // and near tiles must spread their hits over all eight warps instead of piling them onto the one warp next to them
      R.load(g, (uint32_t) m.row_block, tid, false);
      for (int b = 0; b < nb; ++b)
#pragma unroll
        for (int r = 0; r < RI; ++r) hist[b * ROWS_PER_CTA + r * N_CONSUMERS] = 0;
    }
    col0 = m.col0;
    if (!(m.flags & 4u)) {
      const float* tl = ring.tiles + cp.stage * ring.tile_floats;
      ++st.wtiles;
      R.retarget(g, tl + (d + 1) * TileW<D>::tj);
      // every pair with exact d2 < r_max^2 has acc < t[r]  (thr_fast = r_max^2 (1 + e_rel), eabs: absolute error part)
#pragma unroll
      for (int r = 0; r < RI; ++r) t[r] = next_up(next_up(a.thr_fast + R.ea(g, r) - R.xn[r]));
      scan_tile(g, tl, R, t, scratch, hit);
    }
    __syncwarp();
    stage_release(cp.stage);
    if (m.flags & 2u) {
#pragma unroll
      for (int r = 0; r < RI; ++r) {
        if (R.row(r) < g.row_end) {
          for (int b = 0; b < nb; ++b) {
            const uint32_t h = hist[b * ROWS_PER_CTA + r * N_CONSUMERS];
            if (h) atomicAdd(a.cnt + (size_t) b * a.ld_cnt + R.out(r), h);
          }
        }
      }
    }
    cp.advance();
  }
  st.flush(g);
}

// ------------------------------------------------------------------------------------------------
// populations, count mode (up to 8 distinct radii per pass, D <= MAX_TEMPLATE_D): no filter and no slow path for
// ordinary hits.  Per pair: s = acc + |x'|^2 (the fast squared distance, one FADD); per pair and radius:
// v = s - r_b^2 (one FADD against a constant), the count takes the sign bit of v (one integer op), and a running min
// of |v| per row tells whether any pair of the step lies inside the rounding-error band of the radius; only those
// (rare) pairs are re-decided with dist2_exact.  Cost per pair: D FFMA + 1 + ~3 per radius, independent of the hit
// rate -- the tiles that survive the bounding-box pruning have hit rates of 5-25 %, where a per-hit handler costs far
// more.  More radii than a pass holds: api.cu runs several passes, largest radii first, each pruned by its own r_max.
// cnt[b][row] receives #{j : d2(i,j) < rad2[b]} INCLUDING the frame itself when rad2[b] > 0
// (pops_finalize removes it again).  Unused slots of a pass carry rad2 = -1 (nothing is ever inside).
// ------------------------------------------------------------------------------------------------
__host__ __device__ inline size_t pops_count_smem_bytes(size_t ring_bytes, int n_bins) {
  return ((ring_bytes + 15) & ~size_t(15)) + SCRATCH_BYTES + gbox_bytes() + (size_t) n_bins * ROWS_PER_CTA * 4;
}

__device__ __forceinline__ uint32_t sel4u(const uint32_t (&v)[RI], int r) {
  return r == 0 ? v[0] : r == 1 ? v[1] : r == 2 ? v[2] : v[3];
}
__device__ __forceinline__ void add4u(uint32_t (&v)[RI], int r, uint32_t x) {
  if (r == 0) v[0] += x; else if (r == 1) v[1] += x; else if (r == 2) v[2] += x; else v[3] += x;
}

template <int D, int NB>
__global__ void DCB_LAUNCH_BOUNDS(D) pops_count_kernel(const __grid_constant__ PopsArgs a) {
  static_assert(D >= 1 && NB >= 1 && NB <= 8, "count mode: specialised dims, at most eight radii per pass");
  extern __shared__ __align__(128) unsigned char smem[];
  constexpr int TJ = TileW<D>::tj;
  const ScanGeom& g = a.g;
  SmemRing<D> ring(smem, D);
  unsigned char* extra = smem + ((SmemRing<D>::bytes(D) + 15) & ~size_t(15));
  float* scratch = reinterpret_cast<float*>(extra) + threadIdx.x;
  float* gbox = reinterpret_cast<float*>(extra + SCRATCH_BYTES);
  uint32_t* cnt_s = reinterpret_cast<uint32_t*>(extra + SCRATCH_BYTES + gbox_bytes());      // [NB][ROWS_PER_CTA], per item
  ring.init();
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == N_CONSUMER_WARPS) {
    produce<D>(g, ring, false, [](uint32_t, uint32_t&, uint32_t&) {}, [&](uint32_t, int) { return g.prune_thr; });
    return;
  }
  Rows<D> R;
  float wrow[RI];             // row part of the band half width: fast-path error bound + roundings of s and v
  uint32_t cnt[NB][RI];
  const float slack_len = sqrtf(g.prune_slack);
  Pipe<StagesOf<D>::n> cp;
  SlowStats st;
  for (;;) {
    mbar_wait(&ring.full[cp.stage], cp.phase);
    const TileMeta m = ring.meta[cp.stage];
    if (m.row_block < 0) break;
    if (m.flags & 1u) {
      // first tile of an item: this warp prepares row group `warp` (box, zeroed counters) for everybody
      WarpBox<D> wb;
      R.load_group(g, (uint32_t) m.row_block, (uint32_t) warp, lane);
      wb.compute(g, R, lane);
      if (lane < D) {
        gbox[warp * 2 * GBOX_DIMS + lane] = wb.wlo;
        gbox[warp * 2 * GBOX_DIMS + GBOX_DIMS + lane] = wb.whi;
      }
#pragma unroll
      for (int b = 0; b < NB; ++b)
#pragma unroll
        for (int r = 0; r < RI; ++r) cnt_s[b * ROWS_PER_CTA + warp * (32 * RI) + r * 32 + lane] = 0;
      consumer_barrier();
    }
    if (!(m.flags & 4u)) {
      const float* tl = ring.tiles + cp.stage * ring.tile_floats;
      const uint32_t reach = groups_in_reach<D>(gbox, tl + (D + 1) * TJ, lane, g.prune_thr);
      const uint32_t n_units = __popc(reach);
      for (;;) {
        uint32_t u = 0;
        if (lane == 0) u = atomicAdd(&ring.unext[cp.stage], 1u);
        u = __shfl_sync(0xffffffffu, u, 0);
        if (u >= n_units) break;
        const uint32_t gi = __fns(reach, 0, (int) u + 1) >> 2;
        R.load_group(g, (uint32_t) m.row_block, gi, lane);
        R.retarget(g, tl + (D + 1) * TJ);
        if (g.axis_prune) {
          // separating-axis test along the line between the centre of the group's box and the tile's centre
          const float* cen = tl + (D + 1) * TJ;
          const float* blo = gbox + gi * 2 * GBOX_DIMS;
          float ax[D];
#pragma unroll
          for (int k = 0; k < D; ++k) ax[k] = 0.5f * (blo[k] + blo[GBOX_DIMS + k]) - (cen[k] - __ldg(g.centre + k));
          const float lbp = axis_lower_bound<D>(g, R, tl, ax, cen[D], slack_len, lane);
          if (lbp > 0.f && lbp * lbp * 0.999f > g.prune_thr) continue;
        }
        ++st.wtiles;
        // |v - (d2_exact - r_b^2)| < wrow[r] + band[b]: fast-path error (eabs + e_rel r^2) + roundings of s and of v
#pragma unroll
        for (int r = 0; r < RI; ++r) wrow[r] = fmaf(1.01f, R.ea(g, r), 2.4e-7f * R.xn[r]);
#pragma unroll
        for (int b = 0; b < NB; ++b)
#pragma unroll
          for (int r = 0; r < RI; ++r) cnt[b][r] = 0;
        // one or two radii: v = acc + (|x'|^2 - r^2) with the row constant formed once per unit, one FADD per pair and radius
        // (the two roundings -- of the constant, <= u max(|x'|^2, r^2), and of v -- are inside the band like those of s and v)
        constexpr bool FOLD = NB <= 2;
        float qrow[FOLD ? NB : 1][RI];
        if (FOLD) {
#pragma unroll
          for (int b = 0; b < (FOLD ? NB : 1); ++b)
#pragma unroll
            for (int r = 0; r < RI; ++r) qrow[b][r] = R.xn[r] - a.rad2[b];
        }
#pragma unroll 1
        for (int gcol = 0; gcol < TJ; gcol += CJ) {
          float acc[RI][CJ];
          fast_block<D>(tl + gcol, R, acc);
          bool band = false;
#pragma unroll
          for (int r = 0; r < RI; ++r) {
            const float s0 = FOLD ? acc[r][0] : acc[r][0] + R.xn[r], s1 = FOLD ? acc[r][1] : acc[r][1] + R.xn[r];
            const float s2 = FOLD ? acc[r][2] : acc[r][2] + R.xn[r], s3 = FOLD ? acc[r][3] : acc[r][3] + R.xn[r];
#pragma unroll
            for (int b = 0; b < NB; ++b) {
              const float sub = FOLD ? qrow[FOLD ? b : 0][r] : -a.rad2[b];
              const float v0 = s0 + sub, v1 = s1 + sub, v2 = s2 + sub, v3 = s3 + sub;
              cnt[b][r] += (__float_as_uint(v0) >> 31) + (__float_as_uint(v1) >> 31);
              cnt[b][r] += (__float_as_uint(v2) >> 31) + (__float_as_uint(v3) >> 31);
              const float mn = fminf(fminf(fabsf(v0), fabsf(v1)), fminf(fabsf(v2), fabsf(v3)));
              band |= mn < wrow[r] + a.band[b];
            }
          }
          if (band) {
            // rare: some pair of this 4x4 block is within the error band of a radius.  One compact loop
            // (block parked in shared memory) replaces the sign-bit decision of those pairs by the exact one.
#pragma unroll
            for (int r = 0; r < RI; ++r)
#pragma unroll
              for (int c = 0; c < CJ; ++c) scratch[(r * CJ + c) * N_CONSUMERS] = FOLD ? acc[r][c] : acc[r][c] + R.xn[r];
#pragma unroll 1
            for (int p = 0; p < RI * CJ; ++p) {
              const int r = p / CJ;
              const float sv = scratch[p * N_CONSUMERS];
              const float wr = sel4(wrow, r);
              float d2 = 0.f;
              bool have = false;
#pragma unroll
              for (int b = 0; b < NB; ++b) {
                const float v = FOLD ? sv + sel4(qrow[FOLD ? b : 0], r) : sv - a.rad2[b];      // the very value the sign-bit count used
                if (fabsf(v) < wr + a.band[b] && R.row(r) < g.row_end) {
                  ++st.slow;
                  if (!have) {
                    d2 = dist2_exact(g.xT, g.ld, D, R.row(r), m.col0 + gcol + (p % CJ));
                    have = true;
                    ++st.exact;
                  }
                  const uint32_t inside = d2 < a.rad2[b] ? 1u : 0u;      // NaN (padding) -> outside
                  add4u(cnt[b], r, inside - (__float_as_uint(v) >> 31));
                }
              }
            }
          }
        }
        // the unit's counts join the group's (another warp may be adding to the same rows from another tile)
#pragma unroll
        for (int b = 0; b < NB; ++b)
#pragma unroll
          for (int r = 0; r < RI; ++r)
            if (cnt[b][r]) atomicAdd(&cnt_s[b * ROWS_PER_CTA + gi * (32 * RI) + r * 32 + lane], cnt[b][r]);
      }
    }
    __syncwarp();
    stage_release(cp.stage);
    if (m.flags & 2u) {
      // end of the item: every unit is done once all consumer warps are here; this warp writes group `warp` back
      consumer_barrier();
#pragma unroll
      for (int b = 0; b < NB; ++b)
#pragma unroll
        for (int r = 0; r < RI; ++r) {
          const uint32_t slot = (uint32_t) warp * (32 * RI) + (uint32_t) r * 32 + (uint32_t) lane;
          const uint32_t row = block_row0(g, (uint32_t) m.row_block) + slot;
          const uint32_t c = cnt_s[b * ROWS_PER_CTA + slot];
          if (row < g.row_end && c) atomicAdd(a.cnt + (size_t) b * a.ld_cnt + ((uint32_t) m.row_block * ROWS_PER_CTA + slot), c);
        }
    }
    cp.advance();
  }
  st.flush(g);
}


// ------------------------------------------------------------------------------------------------
// populations, bin mode (long radius lists, D <= MAX_TEMPLATE_D): every pair of a step that holds candidates is binned
// BRANCH-FREE through a cell table over the fast squared distance s = acc + |x'|^2:
//     cell = floor(sat(s / span) * (K - 1))                      FMUL.SAT + FFMA (magic 2^21: the cell's byte offset
//                                                                       sits in the mantissa) + 1 LOP3
//     e    = table[cell]                                         1 LDS: the only radius boundary B_k the cell can see,
//                                                                       k in its low 5 mantissa bits (a sentinel if none)
//     bin  = k + (s >= e)                                        FSETP + select + shift-add
//     hist[bin][row] += 1                                        LDS.U16 / IADD / STS.U16, lane-private bank
//     band |= |s - e| < row band                                 FADD + FSETP
// ~14 instructions per pair next to the D FFMAs, whatever the hit rate and the number of radii (the per-hit handler of
// pops_kernel costs ~30 per HIT and the warp pays the maximum over its lanes: 12.9 % of the FFMA peak at 1M x 10, 20 radii).
// Steps with few candidates (boundary tiles) are walked candidate by candidate with the same table.  Pairs inside the
// rounding-error band of a boundary are re-decided with dist2_exact and the histogram is corrected.
// Rows: the warp owns one 128-row group of the block = one tile of the layout, so its bounding box, centre and radius
// are the tile's own header; tiles are pruned against the eight groups by the producer and against the warp's group
// by the consumer with max(box gap, sphere gap) -- in 10 dimensions the boxes of two clusters overlap in most dims
// while their spheres are far apart.
// hist[b][row] counts rad2[b-1] <= d2 < rad2[b], the frame itself included (pops_finalize removes it).
// ------------------------------------------------------------------------------------------------
constexpr int BIN_STRIDE = N_CONSUMERS * RI * 2;      // bytes between the histogram rows of two bins (2048)
#ifndef DCB_BIN_WIDE_FROM
#define DCB_BIN_WIDE_FROM 9                           // n_cols from which the unit is scanned 2 rows x 8 columns per step (99: never)
#endif
#ifndef DCB_BIN_WIDE_COLS
#define DCB_BIN_WIDE_COLS 4                           // wide shape: columns whose lookups are in flight together (2, 4 or 8)
#endif
#ifndef DCB_BIN_COLS
#define DCB_BIN_COLS 2                                // columns whose table lookups are in flight together (1, 2 or 4)
#endif
constexpr int PGEO = 3 * MAX_TEMPLATE_D + 4;          // floats per group in the producer's geometry scratch

// the cell table is aligned to its own size (2 * lut_k * 4 bytes reserved), so that the address of an entry is
// (cell bits & mask) | base: ONE LOP3 instead of an AND and an ADD per pair
__host__ __device__ inline size_t pops_bin_smem_bytes(size_t ring_bytes, int n_bins, int lut_k) {
  return ((ring_bytes + 15) & ~size_t(15)) + SCRATCH_BYTES + (size_t) N_CONSUMER_WARPS * PGEO * 4 + (size_t) 2 * lut_k * 4 + 32 * 4 +
         (size_t) (n_bins + 1) * BIN_STRIDE;
}

// lower bound (fast-value units) of the squared distance between a 128-row group and a tile from their boxes
// (globally centred coordinates) and their spheres (centre + radius); geometry of the group: lo[D], hi[D], c[D], rad
template <int D>
__device__ __forceinline__ float group_tile_lb(const float* __restrict__ gg, const float (&tlo)[D], const float (&thi)[D],
                                               const float (&tc)[D], float trad, float slack_len) {
  float sb = 0.f, sc = 0.f;
#pragma unroll
  for (int k = 0; k < D; ++k) {
    const float gap = fmaxf(fmaxf(gg[k] - thi[k], tlo[k] - gg[D + k]), 0.f);
    sb = fmaf(gap, gap, sb);
    const float dc = gg[2 * D + k] - tc[k];
    sc = fmaf(dc, dc, sc);
  }
  const float gap = sqrtf(sc) * 0.99999f - (gg[3 * D] + trad) * 1.00001f - slack_len;
  const float ss = gap > 0.f ? gap * gap : 0.f;
  return fmaxf(sb, ss) * 0.999f;
}

// Producer of the bin kernel: like produce(), but a tile is streamed iff it comes within thr of at least one of the
// (up to eight) 128-row groups of the block, by box and by sphere.  The groups' geometry is read from their tiles'
// headers (the shard starts at a multiple of 128 rows: api.cu) into pgeo, private to the producer warp.
template <int D>
__device__ __forceinline__ void produce_groups(const ScanGeom& g, SmemRing<D>& ring, float* __restrict__ pgeo) {
  constexpr int TJ = TileW<D>::tj;
  const int lane = threadIdx.x & 31;
  FillPipe<StagesOf<D>::n> pp;
  const uint32_t total = g.n_row_blocks * g.n_col_items;
  const uint32_t rec = (uint32_t) ((D + 1) * TJ + g.dp);
  const float slack_len = sqrtf(g.prune_slack);
  unsigned long long streamed = 0;
  for (;;) {
    uint32_t item = 0;
    if (lane == 0) item = atomicAdd(g.work_counter, 1u);
    item = __shfl_sync(0xffffffffu, item, 0);
    if (item >= total) break;
    uint32_t rb, ci;
    item_coords(g, item, TJ, &rb, &ci);
    const uint32_t t0 = ci * g.tiles_per_item;
    const uint32_t t1 = min(t0 + g.tiles_per_item, g.n_col_tiles);
    if (t0 >= t1) continue;
    const uint32_t blk0 = block_row0(g, rb);
    const uint32_t n_groups = (min(blk0 + (uint32_t) ROWS_PER_CTA, g.row_end) - blk0 + 32u * RI - 1) / (32u * RI);
    __syncwarp();
    for (uint32_t q = lane; q < n_groups * (3 * D + 1); q += 32) {
      const uint32_t gi = q / (3 * D + 1), f = q % (3 * D + 1);
      const float* hdr = g.cT + (size_t) (blk0 / TJ + gi) * rec + (size_t) (D + 1) * TJ;
      float v;
      if (f < (uint32_t) D) v = __ldg(hdr + D + 1 + f);                       // lo
      else if (f < 2u * D) v = __ldg(hdr + D + 1 + f);                      // hi (header: lo[D] then hi[D])
      else if (f < 3u * D) v = __ldg(hdr + (f - 2 * D)) - __ldg(g.centre + (f - 2 * D));     // centre, globally centred
      else v = sqrtf(__ldg(hdr + D));                                       // radius about the centre
      pgeo[gi * PGEO + f] = v;
    }
    for (int k = lane; k < 2 * D; k += 32) ring.rbb[k] = __ldg(g.rbbox + (size_t) rb * 2 * D + k);      // the block's box: coarse level
    __syncwarp();
    bool first = true;
    constexpr uint32_t TPS = SUPER_FRAMES / TJ;     // tiles per super-tile
    for (uint32_t s0 = t0 / TPS; s0 * TPS < t1; s0 += 32) {
     const uint32_t s_end = (t1 + TPS - 1) / TPS;
     uint32_t smask = g.sbbox ? super_mask<D>(g, ring.rbb, s0, s_end, g.prune_thr, lane) : 0xffffffffu;
     while (smask) {
      const uint32_t sp = s0 + (uint32_t) __ffs(smask) - 1u;
      smask &= smask - 1;
      if (sp >= s_end) break;
      const uint32_t seg0 = max(t0, sp * TPS), seg1 = min(t1, (sp + 1) * TPS);
    for (uint32_t base = seg0; base < seg1; base += 32) {
      const uint32_t t = base + lane;
      uint32_t units = 0;                           // bit gi: group gi of the block comes within r_max of tile t
      if (t < seg1) {
        const float* hdr = g.cT + (size_t) t * rec + (size_t) (D + 1) * TJ;
        float tlo[D], thi[D], tc[D];
#pragma unroll
        for (int k = 0; k < D; ++k) {
          tc[k] = __ldg(hdr + k) - __ldg(g.centre + k);
          tlo[k] = __ldg(hdr + D + 1 + k);
          thi[k] = __ldg(hdr + 2 * D + 1 + k);
        }
        const float trad = sqrtf(__ldg(hdr + D));
        for (uint32_t gi = 0; gi < n_groups; ++gi)
          if (!(group_tile_lb<D>(pgeo + gi * PGEO, tlo, thi, tc, trad, slack_len) > g.prune_thr)) units |= 1u << gi;     // NaN keeps
      }
      uint32_t mask = __ballot_sync(0xffffffffu, units != 0u);
      while (mask) {
        const int src = __ffs(mask) - 1;
        const uint32_t tt = base + (uint32_t) src;
        mask &= mask - 1;
        const uint32_t units_tt = __shfl_sync(0xffffffffu, units, src);
        pp.acquire();
        if (lane == 0) {
          TileMeta m;
          m.row_block = (int32_t) rb;
          m.col0 = tt * TJ;
          m.flags = (first ? 1u : 0u) | (units_tt << 8);      // bits 8..15: the (row group, tile) units of the stage
          m.aux = item;
          ring.meta[pp.stage] = m;
          ring.umask[pp.stage] = 0u;                          // units claimed so far
          mbar_arrive_expect_tx(&ring.full[pp.stage], rec * 4);
          tma_load_1d(ring.tiles + pp.stage * ring.tile_floats, g.cT + (size_t) tt * rec, rec * 4, &ring.full[pp.stage]);
        }
        first = false;
        ++streamed;
        pp.advance();
      }
    }
     }
    }
    if (!first) {
      __syncwarp();
      pp.acquire();
      if (lane == 0) {                              // end-of-item marker
        TileMeta m;
        m.row_block = (int32_t) rb;
        m.col0 = 0;
        m.flags = 2u | 4u;
        m.aux = item;
        ring.meta[pp.stage] = m;
        mbar_arrive(&ring.full[pp.stage]);
      }
      pp.advance();
    }
  }
  __syncwarp();
  pp.acquire();
  if (lane == 0) {
    ring.meta[pp.stage].row_block = -1;
    mbar_arrive(&ring.full[pp.stage]);
    if (g.stats && streamed) atomicAdd(g.stats + 2, streamed);
  }
  __syncwarp();
  pp.drain();
}

template <int D>
__global__ void DCB_LAUNCH_BOUNDS(D) pops_bin_kernel(const __grid_constant__ PopsArgs a) {
  static_assert(D >= 1 && D <= MAX_TEMPLATE_D, "bin mode: specialised dims");
  extern __shared__ __align__(128) unsigned char smem[];
  constexpr int TJ = TileW<D>::tj;
  const ScanGeom& g = a.g;
  SmemRing<D> ring(smem, D);
  unsigned char* extra = smem + ((SmemRing<D>::bytes(D) + 15) & ~size_t(15));
  float* scratch = reinterpret_cast<float*>(extra) + threadIdx.x;
  float* pgeo = reinterpret_cast<float*>(extra + SCRATCH_BYTES);
  float* lut_region = pgeo + N_CONSUMER_WARPS * PGEO;                      // 2 * lut_k floats: room to align the table to its size
  const uint32_t lut_bytes = (uint32_t) a.lut_k * 4u;
  const uint32_t lut_sa = (smem_u32(lut_region) + lut_bytes - 1u) & ~(lut_bytes - 1u);      // shared-window address of the table
  const float* lut = lut_region + (lut_sa - smem_u32(lut_region)) / 4u;
  float* rad2s = lut_region + 2 * a.lut_k;
  unsigned char* hist = reinterpret_cast<unsigned char*>(rad2s + 32);       // [n_bins + 1][BIN_STRIDE]
  ring.init();
  for (int q = threadIdx.x; q < a.lut_k; q += CTA_THREADS) const_cast<float*>(lut)[q] = __ldg(a.lut + q);
  if (threadIdx.x < 32) rad2s[threadIdx.x] = a.rad2[threadIdx.x];
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == N_CONSUMER_WARPS) {
    produce_groups<D>(g, ring, pgeo);
    return;
  }
  const int nb = a.n_bins;
  // this thread's counters: 16-bit, rows r and r^1 share a word, word index = bin * 512 + (r >> 1) * 256 + thread:
  // bank = lane for every bin, so the 32 read-modify-writes of a warp never conflict
  unsigned char* hb = hist + threadIdx.x * 4;
  const uint32_t kmask4 = ((uint32_t) a.lut_k - 1u) << 2;
  const float kscale = (float) (a.lut_k - 1);
  // table entry of the cell of s: 2^21 + sat(s / span) (K - 1) has the cell number in mantissa bits 2.. (ulp 1/4), so
  // masking the bit pattern yields the byte offset of the entry, OR-ing the (size-aligned) base its address: one LOP3
  auto entry = [&](float sv) -> float {
    const float v = fmaf(__saturatef(sv * a.lut_scale), kscale, 2097152.f);
    float e;
    asm("ld.shared.f32 %0, [%1];" : "=f"(e) : "r"((__float_as_uint(v) & kmask4) | lut_sa));
    return e;
  };
  Rows<D> R;
  float thr_s = 0.f;          // every pair with exact d2 < r_max^2 has s = fl(acc + |x'|^2) < thr_s (one value per thread, see bwm)
  float bwm = 0.f;            // band half width (s units): ONE value per thread, the widest of its rows (a wider band only
                              // sends a few more pairs to the exact recheck), to keep the inner loop's registers for operands
  bool slow_unit = false;
  const float slack_len = sqrtf(g.prune_slack);
  Pipe<StagesOf<D>::n> cp;
  SlowStats st;
  uint32_t col0 = 0;
  auto hrow = [&](int r) -> unsigned char* { return hb + (r >> 1) * (N_CONSUMERS * 4) + (r & 1) * 2; };
  auto bump = [&](int r, int b, int delta) {
    uint16_t* p = reinterpret_cast<uint16_t*>(hrow(r) + b * BIN_STRIDE);
    *p = (uint16_t) (*p + delta);
  };
  // candidate-by-candidate path (sparse steps, and every step of a unit the table cannot serve)
  auto hit = [&](int r, int jt, float s) {
    const uint32_t j = col0 + jt;
    const uint32_t i = R.row(r);
    if (i >= g.row_end) return;
    ++st.slow;
    int b;
    bool inband;
    if (!slow_unit) {
      const float e = entry(s);
      b = (int) (__float_as_uint(e) & 31u) + (s >= e ? 1 : 0);
      inband = fabsf(s - e) < bwm;
    } else {
      b = bin_of(rad2s, nb, s);
      const float lo = b > 0 ? rad2s[b - 1] : -INFINITY;
      const float hi = rad2s[b];
      const float e = fmaf(g.e_rel, fabsf(s), R.ea(g, r)) * 1.0001f;
      inband = (s - lo < e) || (hi - s <= e);
    }
    if (inband) {
      s = dist2_exact(g.xT, g.ld, D, i, j);
      ++st.exact;
      b = bin_of(rad2s, nb, s);
    }
    if (b < nb) bump(r, b, 1);
  };
  // The histogram belongs to the WARP, not to a row group: it holds the counts of group `cur` of row block `cur_rb` and is
  // added to the global counters (and cleared) when the warp turns to another group and at the end of every item.  So any
  // warp may work on any group's (group, tile) unit: a warp whose own group has nothing to do with a streamed tile takes
  // over a unit of another group instead of running into the end of the ring and waiting there -- a row block that
  // straddles two clusters streams first the tiles only one half of its groups can reach, then those of the other half.
  constexpr uint32_t NONE = 0xffffffffu;
  uint32_t cur = NONE, cur_rb = 0;
  auto flush_hist = [&]() {
    if (cur == NONE) return;
    const uint32_t row_a = block_row0(g, cur_rb) + cur * (32u * RI) + (uint32_t) lane;        // row r of the thread: row_a + 32 r
    const size_t out_a = (size_t) cur_rb * ROWS_PER_CTA + cur * (32u * RI) + (uint32_t) lane;
    for (int b = 0; b <= nb; ++b) {
#pragma unroll
      for (int h2 = 0; h2 < 2; ++h2) {
        uint32_t* w = reinterpret_cast<uint32_t*>(hb + b * BIN_STRIDE + h2 * (N_CONSUMERS * 4));
        const uint32_t v = *w;
        if (v) {
          *w = 0u;
          if (b < nb) {
            const uint32_t c0 = v & 0xffffu, c1 = v >> 16, r0 = 2u * h2;
            if (c0 && row_a + 32u * r0 < g.row_end) atomicAdd(a.cnt + (size_t) b * a.ld_cnt + out_a + 32u * r0, c0);
            if (c1 && row_a + 32u * (r0 + 1u) < g.row_end) atomicAdd(a.cnt + (size_t) b * a.ld_cnt + out_a + 32u * (r0 + 1u), c1);
          }
        }
      }
    }
    cur = NONE;
  };
  for (int b = 0; b <= nb; ++b) {
    *reinterpret_cast<uint32_t*>(hb + b * BIN_STRIDE) = 0u;
    *reinterpret_cast<uint32_t*>(hb + b * BIN_STRIDE + N_CONSUMERS * 4) = 0u;
  }
  for (;;) {
    mbar_wait(&ring.full[cp.stage], cp.phase);
    const TileMeta m = ring.meta[cp.stage];
    if (m.row_block < 0) break;
    col0 = m.col0;
    if (!(m.flags & 4u)) {
      const float* tl = ring.tiles + cp.stage * ring.tile_floats;
      const float* cen = tl + (D + 1) * TJ;
      // which unit of the stage is this warp's?  Its own group's, if that group reaches the tile (nobody else is left with
      // it then); else ONE unit nobody has claimed yet, preferably of the group whose counts the histogram already holds
      const uint32_t units = (m.flags >> 8) & 0xffu;
      uint32_t pick = NONE;
      if ((units >> warp) & 1u) {
        uint32_t old = 0;
        if (lane == 0) old = atomicOr(&ring.umask[cp.stage], 1u << warp);
        old = __shfl_sync(0xffffffffu, old, 0);
        if (!((old >> warp) & 1u)) pick = (uint32_t) warp;
      } else if (a.steal) {
        uint32_t cand = units & ~*reinterpret_cast<volatile uint32_t*>(&ring.umask[cp.stage]);
        cand = __shfl_sync(0xffffffffu, cand, 0);
        while (cand) {
          const uint32_t c = (cur != NONE && ((cand >> cur) & 1u)) ? cur : (uint32_t) __ffs(cand) - 1u;
          uint32_t old = 0;
          if (lane == 0) old = atomicOr(&ring.umask[cp.stage], 1u << c);
          old = __shfl_sync(0xffffffffu, old, 0);
          if (!((old >> c) & 1u)) {
            pick = c;
            break;
          }
          cand &= ~(old | (1u << c));
        }
      }
      bool reach = pick != NONE;
      if (reach) {
        if (pick != cur || cur_rb != (uint32_t) m.row_block) {
          flush_hist();
          cur = pick;
          cur_rb = (uint32_t) m.row_block;
        }
        R.load_group(g, (uint32_t) m.row_block, pick, lane);
        R.retarget(g, cen);
        // Separating-axis test along the line from the tile's centre to the group's centre: the rows' smallest and the
        // columns' largest projection onto it bound the distance of every pair from below.  In 10 dimensions the boxes and
        // spheres of two neighbouring clusters overlap while their extents along the line between them do not
        // (projection of a Gaussian blob: ~3 sigma, its radius: ~(sqrt(D) + 2) sigma).  ~130 instructions per unit.
        if (a.proj_prune) {
          // the group's centre: header of the group's own tile (a group is one tile of the spatial order), globally centred
          const float* hdr = g.cT + (size_t) ((block_row0(g, (uint32_t) m.row_block) + pick * (32u * RI)) / TJ) * ((D + 1) * TJ + g.dp) +
                             (size_t) (D + 1) * TJ;
          const float gck = lane < D ? __ldg(hdr + lane) - __ldg(g.centre + lane) : 0.f;
          float u[D];
#pragma unroll
          for (int k = 0; k < D; ++k) u[k] = __shfl_sync(0xffffffffu, gck, k) - (cen[k] - __ldg(g.centre + k));
          const float lbp = axis_lower_bound<D>(g, R, tl, u, cen[D], slack_len, lane);
          if (lbp > 0.f && lbp * lbp * 0.999f > g.prune_thr) reach = false;
        }
      }
      if (reach) {
        ++st.wtiles;
        bwm = 0.f;
        float eam = 0.f;
#pragma unroll
        for (int r = 0; r < RI; ++r) {
          eam = fmaxf(eam, R.ea(g, r));
          // |s - e - (d2_exact - B_k)| < bw: fast-path error (ea + e_rel r^2), roundings of s, the table's perturbation of B_k
          bwm = fmaxf(bwm, fmaf(1.01f, R.ea(g, r), 2.4e-7f * R.xn[r]) + a.band_max);
        }
        // exact d2 < r_max^2  =>  acc + |x'|^2 < thr_fast + ea in real arithmetic (thr_fast = r_max^2 (1 + e_rel)); the factor
        // covers the rounding of the sum s and of the threshold itself
        thr_s = next_up((a.thr_fast + eam) * 1.000001f);
        slow_unit = __any_sync(0xffffffffu, !(bwm < a.lut_margin));
        if constexpr (D >= DCB_BIN_WIDE_FROM) {
          // Wide shape for the dims whose row operands crowd the register file (4 D of the 96 registers): the unit is scanned in
          // two halves of TWO rows x EIGHT columns per step -- 2 D operand registers, the same 8 FFMA2 per dim and step (rows
          // r, r+1 x one column each) with two broadcast LDS.128 -- so that the lookup chains of a dense step have registers
          // to be in flight together.  Rows 2h, 2h+1 of the thread in half h: the same rows, histogram slots and results as
          // the 4 x 4 shape.
#pragma unroll 1
          for (int h = 0; h < 2; ++h) {
            float xw[2][D];
            {
              const uint32_t p0 = R.pos(g, 2 * h), p1 = R.pos(g, 2 * h + 1);
#pragma unroll
              for (int k = 0; k < D; ++k) {
                xw[0][k] = __ldg(g.xT + (size_t) k * g.ld + p0) - cen[k];          // the arithmetic of Rows::retarget
                xw[1][k] = __ldg(g.xT + (size_t) k * g.ld + p1) - cen[k];
              }
            }
            const float xnw[2] = {h ? R.xn[2] : R.xn[0], h ? R.xn[3] : R.xn[1]};
            unsigned char* hbh = hb + h * (N_CONSUMERS * 4);                        // rows 2h (low half word), 2h+1 (high)
#pragma unroll 1
            for (int gcol = 0; gcol < TJ; gcol += 8) {
              float acc[2][8];
              {
                unsigned long long a2[8];
                const float* tlg = tl + gcol;
                {
                  const float4 na = *reinterpret_cast<const float4*>(tlg + D * TJ), nb4 = *reinterpret_cast<const float4*>(tlg + D * TJ + 4);
                  const float4 ya = *reinterpret_cast<const float4*>(tlg), yb = *reinterpret_cast<const float4*>(tlg + 4);
                  const float yc[8] = {ya.x, ya.y, ya.z, ya.w, yb.x, yb.y, yb.z, yb.w};
                  const float nc[8] = {na.x, na.y, na.z, na.w, nb4.x, nb4.y, nb4.z, nb4.w};
                  const unsigned long long x2 = pack2(xw[0][0], xw[1][0]);
#pragma unroll
                  for (int c = 0; c < 8; ++c) a2[c] = fma2(x2, pack2(yc[c], yc[c]), pack2(nc[c], nc[c]));
                }
#pragma unroll
                for (int k = 1; k < D; ++k) {
                  const float4 ya = *reinterpret_cast<const float4*>(tlg + k * TJ), yb = *reinterpret_cast<const float4*>(tlg + k * TJ + 4);
                  const float yc[8] = {ya.x, ya.y, ya.z, ya.w, yb.x, yb.y, yb.z, yb.w};
                  const unsigned long long x2 = pack2(xw[0][k], xw[1][k]);
#pragma unroll
                  for (int c = 0; c < 8; ++c) a2[c] = fma2(x2, pack2(yc[c], yc[c]), a2[c]);
                }
#pragma unroll
                for (int c = 0; c < 8; ++c) unpack2(a2[c], acc[0][c], acc[1][c]);
              }
              bool any = false;
#pragma unroll
              for (int r2 = 0; r2 < 2; ++r2) {
                const float mn = fminf(fminf(fminf(acc[r2][0], acc[r2][1]), fminf(acc[r2][2], acc[r2][3])),
                                       fminf(fminf(acc[r2][4], acc[r2][5]), fminf(acc[r2][6], acc[r2][7])));
                any |= (mn + xnw[r2] < thr_s);
              }
              const uint32_t act = __ballot_sync(0xffffffffu, any);
              if (act == 0u) continue;
              if (__popc(act) >= a.dense_lanes && !slow_unit) {
                bool band = false;
#pragma unroll
                for (int c0 = 0; c0 < 8; c0 += DCB_BIN_WIDE_COLS) {
                  uint16_t* hp[DCB_BIN_WIDE_COLS][2];
#pragma unroll
                  for (int cc = 0; cc < DCB_BIN_WIDE_COLS; ++cc)
#pragma unroll
                    for (int r2 = 0; r2 < 2; ++r2) {
                      const float sv = acc[r2][c0 + cc] + xnw[r2];
                      const float e = entry(sv);
                      const float dlt = sv - e;
                      const uint32_t b = (__float_as_uint(e) & 31u) + 1u - (__float_as_uint(dlt) >> 31);
                      band |= fabsf(dlt) < bwm;
                      hp[cc][r2] = reinterpret_cast<uint16_t*>(hbh + r2 * 2 + b * BIN_STRIDE);
                    }
                  // the two counters of a column belong to two different rows; columns of the same row stay in order
#pragma unroll
                  for (int cc = 0; cc < DCB_BIN_WIDE_COLS; ++cc) {
                    const uint16_t h0 = *hp[cc][0], h1 = *hp[cc][1];
                    *hp[cc][0] = (uint16_t) (h0 + 1);
                    *hp[cc][1] = (uint16_t) (h1 + 1);
                  }
                }
                if (band) {
#pragma unroll
                  for (int r2 = 0; r2 < 2; ++r2)
#pragma unroll
                    for (int c = 0; c < 8; ++c) scratch[(r2 * 8 + c) * N_CONSUMERS] = acc[r2][c];
#pragma unroll 1
                  for (int p = 0; p < 16; ++p) {
                    const int r2 = p >> 3, r = 2 * h + r2;
                    const float sv = scratch[p * N_CONSUMERS] + (r2 ? xnw[1] : xnw[0]);
                    const float e = entry(sv);
                    const float dlt = sv - e;
                    if (fabsf(dlt) < bwm) {
                      const int bf = (int) ((__float_as_uint(e) & 31u) + 1u - (__float_as_uint(dlt) >> 31));     // as counted above
                      ++st.slow;
                      ++st.exact;
                      const float d2 = dist2_exact(g.xT, g.ld, D, R.row(r), m.col0 + gcol + (p & 7));
                      const int be = (d2 == d2) ? bin_of(rad2s, nb, d2) : nb;        // NaN (padding) -> outside
                      if (be != bf) {
                        bump(r, bf, -1);
                        bump(r, be, 1);
                      }
                    }
                  }
                }
              } else if (any) {
                uint32_t mask = 0;
#pragma unroll
                for (int r2 = 0; r2 < 2; ++r2)
#pragma unroll
                  for (int c = 0; c < 8; ++c) {
                    const float sv = acc[r2][c] + xnw[r2];
                    scratch[(r2 * 8 + c) * N_CONSUMERS] = sv;
                    mask |= (sv < thr_s) ? (1u << (r2 * 8 + c)) : 0u;
                  }
#pragma unroll 1
                while (mask) {
                  const int p = __ffs(mask) - 1;
                  mask &= mask - 1;
                  hit(2 * h + (p >> 3), gcol + (p & 7), scratch[p * N_CONSUMERS]);
                }
              }
            }
          }
        } else {
#pragma unroll 1
        for (int gcol = 0; gcol < TJ; gcol += CJ) {
          float acc[RI][CJ];
          fast_block<D>(tl + gcol, R, acc);
          bool any = false;
#pragma unroll
          for (int r = 0; r < RI; ++r) {
            const float mn = fminf(fminf(acc[r][0], acc[r][1]), fminf(acc[r][2], acc[r][3]));
            any |= (mn + R.xn[r] < thr_s);
          }
          const uint32_t act = __ballot_sync(0xffffffffu, any);
          if (act == 0u) continue;
          if (__popc(act) >= a.dense_lanes && !slow_unit) {
            // dense step: all 16 pairs of every lane through the table, no branches; misses land in the dump row nb
            bool band = false;
            // two columns (eight independent lookup chains) at a time, then their read-modify-writes column by column:
            // the four counters of a column belong to four different rows (distinct addresses: load all, then store all),
            // while two columns of the same row may hit the same counter and must stay in order
#pragma unroll
            for (int c0 = 0; c0 < CJ; c0 += DCB_BIN_COLS) {
              uint16_t* hp[DCB_BIN_COLS][RI];
#pragma unroll
              for (int cc = 0; cc < DCB_BIN_COLS; ++cc)
#pragma unroll
                for (int r = 0; r < RI; ++r) {
                  const float s = acc[r][c0 + cc] + R.xn[r];
                  const float e = entry(s);
                  const float dlt = s - e;
                  // bin = k + (s >= e): k from the entry's low bits, the comparison from the sign of s - e
                  const uint32_t b = (__float_as_uint(e) & 31u) + 1u - (__float_as_uint(dlt) >> 31);
                  band |= fabsf(dlt) < bwm;
                  hp[cc][r] = reinterpret_cast<uint16_t*>(hrow(r) + b * BIN_STRIDE);
                }
#pragma unroll
              for (int cc = 0; cc < DCB_BIN_COLS; ++cc) {
                uint16_t h[RI];
#pragma unroll
                for (int r = 0; r < RI; ++r) h[r] = *hp[cc][r];
#pragma unroll
                for (int r = 0; r < RI; ++r) *hp[cc][r] = (uint16_t) (h[r] + 1);
              }
            }
            if (band) {
              // rare: some pair of this block is inside the error band of a boundary: re-decide those exactly and move
              // their count to the right bin (one compact loop over the block parked in shared memory)
#pragma unroll
              for (int r = 0; r < RI; ++r)
#pragma unroll
                for (int c = 0; c < CJ; ++c) scratch[(r * CJ + c) * N_CONSUMERS] = acc[r][c];
#pragma unroll 1
              for (int p = 0; p < RI * CJ; ++p) {
                const int r = p / CJ;
                const float s = scratch[p * N_CONSUMERS] + sel4(R.xn, r);
                const float e = entry(s);
                const float dlt = s - e;
                if (fabsf(dlt) < bwm) {
                  const int bf = (int) ((__float_as_uint(e) & 31u) + 1u - (__float_as_uint(dlt) >> 31));     // as counted above
                  ++st.slow;
                  ++st.exact;
                  const float d2 = dist2_exact(g.xT, g.ld, D, R.row(r), m.col0 + gcol + (p % CJ));
                  const int be = (d2 == d2) ? bin_of(rad2s, nb, d2) : nb;        // NaN (padding) -> outside
                  if (be != bf) {
                    bump(r, bf, -1);
                    bump(r, be, 1);
                  }
                }
              }
            }
          } else if (any) {
            // sparse step: the block is parked in shared memory and this lane's candidates are walked one by one
            uint32_t mask = 0;
#pragma unroll
            for (int r = 0; r < RI; ++r)
#pragma unroll
              for (int c = 0; c < CJ; ++c) {
                const float sv = acc[r][c] + R.xn[r];
                scratch[(r * CJ + c) * N_CONSUMERS] = sv;
                mask |= (sv < thr_s) ? (1u << (r * CJ + c)) : 0u;
              }
#pragma unroll 1
            while (mask) {
              const int p = __ffs(mask) - 1;
              mask &= mask - 1;
              hit(p / CJ, gcol + (p % CJ), scratch[p * N_CONSUMERS]);
            }
          }
        }
        }
      }
    }
    __syncwarp();
    stage_release(cp.stage);
    if (m.flags & 2u) flush_hist();               // end of the item
    cp.advance();
  }
  st.flush(g);
}

// ================================================================================================
// nearest neighbours (rows and columns in the context's spatial order)
// ================================================================================================
struct NnArgs {
  ScanGeom g;                   // g.xrow = lof
  const uint32_t* perm;         // [n] position -> original frame
  const uint32_t* lo;           // [n] per position: number of frames with a strictly lower free energy
                                //     (fe[j] < fe[i]  <=>  lo[j] < lo[i])
  const float* lof;             // [ld] (float)(lo >> lo_shift), +inf in the padding: the same ranks for the fast filter
  float lo_bias;                // 0 when the float ranks are exact (lo_shift == 0), else 1 (equal coarse ranks stay candidates)
  uint32_t window;              // > 0: scan only the column tiles within `window` tiles of the row block's own position
                                // (first pass: settles tight thresholds before the full scan); 0: all tiles
  unsigned long long* key_nn;   // [rows of the launch, out_index order] (d2 bits << 32 | original index), atomicMin'ed
  unsigned long long* key_hd;
};

__host__ __device__ inline size_t screen_smem_bytes(size_t ring_bytes) { return ((ring_bytes + 15) & ~size_t(15)) + SCRATCH_BYTES; }

constexpr size_t NN_PARK_BYTES = (size_t) 5 * RI * N_CONSUMERS * 4;      // per thread: the parked values of NnFilter
__host__ __device__ inline size_t nn_smem_bytes(size_t ring_bytes) {
  return ((ring_bytes + 15) & ~size_t(15)) + SCRATCH_BYTES + NN_PARK_BYTES + (size_t) 2 * ROWS_PER_CTA * 8 + (size_t) 2 * ROWS_PER_CTA * 4 +
         gbox_bytes() + (size_t) 3 * N_CONSUMER_WARPS * 4;
}

__device__ __forceinline__ float key_d2(unsigned long long k) { return __uint_as_float((uint32_t) (k >> 32)); }

// Filter state of the neighbour search.  A pair (row r, column c) is worth the exact evaluation iff
//   acc < t_nn[r]                      (could beat or tie the nearest neighbour found so far), or
//   acc < t_hd[r] and lo[c] < lo[r]    (could beat the nearest neighbour with lower free energy).
// Two levels.  The inner loop only asks whether ANY pair of a row's CJ columns lies below tc[r] = max(t_nn[r], t_hd[r])
// (a min tree over the accumulators: one instruction per pair on the ALU pipe, nothing on the FMA pipe the packed FMAs
// saturate).  The few blocks that pass get the exact per-pair threshold  te = t_nn[r] + cand * dl[r],
// cand = sat(lor[r] - lo_c) in {0,1}: frames that are close but have no lower free energy never reach the candidate
// handler, however far the lower-free-energy neighbour of a density peak is.
// Only tc lives in registers; everything the second level, the handler and the unit's prologue / epilogue need is parked in
// shared memory ([value][thread]: conflict-free), so that at D = 9, 10 (96 registers, 40 of them row operands) none of the
// loop's operands is spilled.
struct NnFilter {
  float tc[RI];                 // max(t_nn, t_hd)
  float* park;                  // this thread's slots, value v of row r at park[(v * RI + r) * N_CONSUMERS]
  __device__ __forceinline__ float& t_nn(int r) { return park[r * N_CONSUMERS]; }
  __device__ __forceinline__ float& t_hd(int r) { return park[(RI + r) * N_CONSUMERS]; }
  __device__ __forceinline__ float& dl(int r) { return park[(2 * RI + r) * N_CONSUMERS]; }     // min(t_hd - t_nn, 1e37), rounded up; 0 where t_nn is +inf
  __device__ __forceinline__ float& lor(int r) { return park[(3 * RI + r) * N_CONSUMERS]; }    // (float rank of the row) + lo_bias
  __device__ __forceinline__ float& xn(int r) { return park[(4 * RI + r) * N_CONSUMERS]; }     // |x'|^2 of the row
  // after t_nn(r) or t_hd(r) changed
  __device__ __forceinline__ void update(int r) {
    const float tn = t_nn(r), th = t_hd(r);
    float v = 0.f;
    if (tn < INFINITY) v = th < INFINITY ? fminf(next_up((th - tn) * 1.000001f), 1e37f) : 1e37f;
    dl(r) = fmaxf(v, 0.f);
    put4(tc, r, fmaxf(tn, th));
  }
};
constexpr int NN_PARK_VALUES = 5;

// second level of the filter for one RI x CJ block that passed the first: the pairs below their exact per-pair threshold
// go to the handler; one compact copy of it (see walk_hits).  lrow4: the free-energy ranks of the block's CJ columns.
template <class Hit>
__device__ __forceinline__ void walk_hits_nn(float* __restrict__ scratch, const float (&acc)[RI][CJ], const float* __restrict__ lrow4, NnFilter& F,
                                             int jt0, Hit& hit) {
  uint32_t mask = 0;
  const float4 l4 = *reinterpret_cast<const float4*>(lrow4);
  const float lc[CJ] = {l4.x, l4.y, l4.z, l4.w};
#pragma unroll
  for (int r = 0; r < RI; ++r) {
    const float lor = F.lor(r), dl = F.dl(r), tn = F.t_nn(r);
#pragma unroll
    for (int c = 0; c < CJ; ++c) {
      scratch[(r * CJ + c) * N_CONSUMERS] = acc[r][c];
      const float te = fmaf(__saturatef(lor - lc[c]), dl, tn);
      mask |= (acc[r][c] < te) ? (1u << (r * CJ + c)) : 0u;
    }
  }
#pragma unroll 1
  while (mask) {
    const int p = __ffs(mask) - 1;
    mask &= mask - 1;
    hit(p / CJ, jt0 + (p % CJ), scratch[p * N_CONSUMERS]);     // hit() re-checks against the thresholds of the moment
  }
}

template <int D, class Hit>
__device__ __forceinline__ void scan_tile_nn(const ScanGeom&, const float* __restrict__ tl, const float* __restrict__ lrow, const Rows<D>& R,
                                             NnFilter& F, float* __restrict__ scratch, Hit& hit) {
  constexpr int TJ = TileW<D>::tj;
#pragma unroll 1
  for (int g = 0; g < TJ; g += CJ) {
    float acc[RI][CJ];
    fast_block<D>(tl + g, R, acc);
    bool any = false;
#pragma unroll
    for (int r = 0; r < RI; ++r) {
      const float mn = fminf(fminf(acc[r][0], acc[r][1]), fminf(acc[r][2], acc[r][3]));
      any |= mn < F.tc[r];
    }
    if (any) walk_hits_nn(scratch, acc, lrow + g, F, g, hit);
  }
}

// Wide shape of the neighbour scan for the dims whose row operands crowd the register file (see pops_bin_kernel): the unit
// is scanned in two halves of TWO rows x EIGHT columns per step -- 2 D operand registers, the same 8 FFMA2 per dim and step --
// which leaves registers to keep the broadcast loads of the next dims in flight.  Rows 2h, 2h+1 of the thread in half h.
#ifndef DCB_NN_WIDE_FROM
#define DCB_NN_WIDE_FROM 9
#endif
template <int D, class Hit>
__device__ __forceinline__ void scan_tile_nn_wide(const ScanGeom& g, const float* __restrict__ tl, const float* __restrict__ lrow, const Rows<D>& R,
                                                  NnFilter& F, float* __restrict__ scratch, Hit& hit) {
  constexpr int TJ = TileW<D>::tj;
  const float* __restrict__ cen = tl + (D + 1) * TJ;
#pragma unroll 1
  for (int h = 0; h < 2; ++h) {
    float xw[2][D];
    {
      const uint32_t p0 = R.pos(g, 2 * h), p1 = R.pos(g, 2 * h + 1);
#pragma unroll
      for (int k = 0; k < D; ++k) {
        xw[0][k] = __ldg(g.xT + (size_t) k * g.ld + p0) - cen[k];          // the arithmetic of Rows::retarget
        xw[1][k] = __ldg(g.xT + (size_t) k * g.ld + p1) - cen[k];
      }
    }
#pragma unroll 1
    for (int gc = 0; gc < TJ; gc += 8) {
      float acc[2][8];
      {
        unsigned long long a2[8];
        const float* tlg = tl + gc;
        {
          const float4 na = *reinterpret_cast<const float4*>(tlg + D * TJ), nb = *reinterpret_cast<const float4*>(tlg + D * TJ + 4);
          const float4 ya = *reinterpret_cast<const float4*>(tlg), yb = *reinterpret_cast<const float4*>(tlg + 4);
          const float yc[8] = {ya.x, ya.y, ya.z, ya.w, yb.x, yb.y, yb.z, yb.w};
          const float nc[8] = {na.x, na.y, na.z, na.w, nb.x, nb.y, nb.z, nb.w};
          const unsigned long long x2 = pack2(xw[0][0], xw[1][0]);
#pragma unroll
          for (int c = 0; c < 8; ++c) a2[c] = fma2(x2, pack2(yc[c], yc[c]), pack2(nc[c], nc[c]));
        }
#pragma unroll
        for (int k = 1; k < D; ++k) {
          const float4 ya = *reinterpret_cast<const float4*>(tlg + k * TJ), yb = *reinterpret_cast<const float4*>(tlg + k * TJ + 4);
          const float yc[8] = {ya.x, ya.y, ya.z, ya.w, yb.x, yb.y, yb.z, yb.w};
          const unsigned long long x2 = pack2(xw[0][k], xw[1][k]);
#pragma unroll
          for (int c = 0; c < 8; ++c) a2[c] = fma2(x2, pack2(yc[c], yc[c]), a2[c]);
        }
#pragma unroll
        for (int c = 0; c < 8; ++c) unpack2(a2[c], acc[0][c], acc[1][c]);
      }
      // first level: any of a row's eight columns below tc?
      const float tc0 = h ? F.tc[2] : F.tc[0], tc1 = h ? F.tc[3] : F.tc[1];
      const float m0 = fminf(fminf(fminf(acc[0][0], acc[0][1]), fminf(acc[0][2], acc[0][3])), fminf(fminf(acc[0][4], acc[0][5]), fminf(acc[0][6], acc[0][7])));
      const float m1 = fminf(fminf(fminf(acc[1][0], acc[1][1]), fminf(acc[1][2], acc[1][3])), fminf(fminf(acc[1][4], acc[1][5]), fminf(acc[1][6], acc[1][7])));
      if (m0 < tc0 || m1 < tc1) {
        // second level: exact per-pair thresholds, candidates to the handler (one compact loop)
        uint32_t mask = 0;
        const float4 la = *reinterpret_cast<const float4*>(lrow + gc), lb = *reinterpret_cast<const float4*>(lrow + gc + 4);
        const float lc[8] = {la.x, la.y, la.z, la.w, lb.x, lb.y, lb.z, lb.w};
#pragma unroll
        for (int r2 = 0; r2 < 2; ++r2) {
          const int r = 2 * h + r2;
          const float lor = F.lor(r), dl = F.dl(r), tn = F.t_nn(r);
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            scratch[(r2 * 8 + c) * N_CONSUMERS] = acc[r2][c];
            const float te = fmaf(__saturatef(lor - lc[c]), dl, tn);
            mask |= (acc[r2][c] < te) ? (1u << (r2 * 8 + c)) : 0u;
          }
        }
#pragma unroll 1
        while (mask) {
          const int p = __ffs(mask) - 1;
          mask &= mask - 1;
          hit(2 * h + (p >> 3), gc + (p & 7), scratch[p * N_CONSUMERS]);     // hit() re-checks against the thresholds of the moment
        }
      }
    }
  }
}

// run-time-D variant
template <class Hit>
__device__ __forceinline__ void scan_tile_nn(const ScanGeom& gm, const float* __restrict__ tl, const float* __restrict__ lrow, const Rows<0>& R,
                                             NnFilter& F, float* __restrict__ scratch, Hit& hit) {
  constexpr int TJ = TileW<0>::tj, CG = TileW<0>::cj;
  const int d = gm.d;
  const float* __restrict__ cen = tl + (d + 1) * TJ;
#pragma unroll 1
  for (int g = 0; g < TJ; g += CG) {
    float acc[CG / CJ][RI][CJ];
#pragma unroll
    for (int c4 = 0; c4 < CG / CJ; ++c4) {
      const float4 n4 = *reinterpret_cast<const float4*>(tl + d * TJ + g + c4 * CJ);
#pragma unroll
      for (int r = 0; r < RI; ++r) {
        acc[c4][r][0] = n4.x; acc[c4][r][1] = n4.y; acc[c4][r][2] = n4.z; acc[c4][r][3] = n4.w;
      }
    }
#pragma unroll 2
    for (int k = 0; k < d; ++k) {
      float xr[RI];
      const float ck = cen[k];
#pragma unroll
      for (int r = 0; r < RI; ++r) xr[r] = __ldg(gm.xT + (size_t) k * gm.ld + R.p[r]) - ck;
#pragma unroll
      for (int c4 = 0; c4 < CG / CJ; ++c4) {
        const float4 y4 = *reinterpret_cast<const float4*>(tl + k * TJ + g + c4 * CJ);
#pragma unroll
        for (int r = 0; r < RI; ++r) {
          acc[c4][r][0] = fmaf(xr[r], y4.x, acc[c4][r][0]);
          acc[c4][r][1] = fmaf(xr[r], y4.y, acc[c4][r][1]);
          acc[c4][r][2] = fmaf(xr[r], y4.z, acc[c4][r][2]);
          acc[c4][r][3] = fmaf(xr[r], y4.w, acc[c4][r][3]);
        }
      }
    }
#pragma unroll
    for (int c4 = 0; c4 < CG / CJ; ++c4) {
      bool any = false;
#pragma unroll
      for (int r = 0; r < RI; ++r) {
        const float mn = fminf(fminf(acc[c4][r][0], acc[c4][r][1]), fminf(acc[c4][r][2], acc[c4][r][3]));
        any |= mn < F.tc[r];
      }
      if (any) walk_hits_nn(scratch, acc[c4], lrow + g + c4 * CJ, F, g + c4 * CJ, hit);
    }
  }
}

template <int D>
__global__ void DCB_LAUNCH_BOUNDS(D) nn_kernel(const __grid_constant__ NnArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  constexpr int TJ = TileW<D>::tj;
  const ScanGeom& g = a.g;
  const int d = D ? D : g.d;
  SmemRing<D> ring(smem, d, 1);
  unsigned char* extra = smem + ((SmemRing<D>::bytes(d, 1) + 15) & ~size_t(15));
  float* scratch = reinterpret_cast<float*>(extra) + threadIdx.x;
  // per row of the block (slot = group * 128 + r * 32 + lane): best[0][slot] = nearest-neighbour key, best[1][slot] =
  // nearest neighbour with lower free energy, lo_s = free-energy rank, lor_s = the same rank as the filter's float
  unsigned long long* best = reinterpret_cast<unsigned long long*>(extra + SCRATCH_BYTES + NN_PARK_BYTES);
  uint32_t* lo_s = reinterpret_cast<uint32_t*>(extra + SCRATCH_BYTES + NN_PARK_BYTES + (size_t) 2 * ROWS_PER_CTA * 8);
  float* lor_s = reinterpret_cast<float*>(lo_s + ROWS_PER_CTA);
  float* gbox = lor_s + ROWS_PER_CTA;
  // per row group (float bits, d2 units, pruning margins included): gthr_nn = what its rows still accept as nearest neighbour,
  // gthr_hd = the same for the lower-free-energy neighbour (rows that can have one), glor = its largest filter rank
  unsigned int* gthr_nn = reinterpret_cast<unsigned int*>(gbox + N_CONSUMER_WARPS * 2 * GBOX_DIMS);
  unsigned int* gthr_hd = gthr_nn + N_CONSUMER_WARPS;
  float* glor = reinterpret_cast<float*>(gthr_hd + N_CONSUMER_WARPS);
  ring.init();
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == N_CONSUMER_WARPS) {
    // initial pruning threshold of an item: what the seeds and earlier items already found for the rows of the block
    produce<D>(g, ring, true, [&](uint32_t rb, uint32_t& lim0, uint32_t& lim1) {
      if (a.window) {
        const uint32_t t_first = block_row0(g, rb) / TJ;
        const uint32_t t_last = (min(block_row0(g, rb) + (uint32_t) ROWS_PER_CTA, g.row_end) - 1) / TJ;
        lim0 = t_first > a.window ? t_first - a.window : 0u;
        lim1 = min(lim1, t_last + a.window + 1);
      }
    }, [&](uint32_t rb, int ln) {
      float v = ln < N_CONSUMER_WARPS ? *reinterpret_cast<volatile float*>(g.blk_thr + (size_t) rb * N_CONSUMER_WARPS + ln) : 0.f;
      if (!(v < INFINITY)) v = INFINITY;
      return __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(fmaxf(v, 0.f))));
    });
    return;
  }
  Rows<D> R;
  NnFilter F;
  F.park = reinterpret_cast<float*>(extra + SCRATCH_BYTES) + threadIdx.x;
  const float slack_len = sqrtf(g.prune_slack);
  uint32_t col0 = 0, slot0 = 0;          // slot0: first slot of the group being worked on + lane
  Pipe<StagesOf<D>::n> cp;
  SlowStats st;
  // every column whose exact d2 is <= `d2` satisfies acc < thr(d2) (api.cu: error_bounds)
  auto thr = [&](float d2, float eabs, float xnr) { return next_up(next_up(fmaf(g.e_rel, d2, d2) + eabs - xnr)); };
  auto hit = [&](int r, int jt, float accv) {
    const uint32_t j = col0 + jt;
    const uint32_t i = R.row(r);
    ++st.slow;
    if (j == i || j >= g.n || i >= g.row_end) return;
    const float tn = F.t_nn(r), th = F.t_hd(r);
    const uint32_t slot = slot0 + (uint32_t) r * 32u;
    const bool hd_cand = __ldg(a.lo + j) < lo_s[slot];
    if (!(accv < tn) && !(hd_cand && accv < th)) return;      // thresholds may have tightened since the block was filtered
    const float d2 = dist2_exact(g.xT, g.ld, d, i, j);
    ++st.exact;
    if (!(d2 < FLT_MAX)) return;
    const unsigned long long key = ((unsigned long long) __float_as_uint(d2) << 32) | __ldg(a.perm + j);
    // another warp may be working on the same rows with another tile: the keys are only ever lowered, atomically
    const float xnr = F.xn(r), ear = g.c_loc * (xnr + R.ymax);
    bool changed = false;
    if (key < atomicMin(best + slot, key)) {
      const float v = thr(d2, ear, xnr);
      F.t_nn(r) = v;
      if (lo_s[slot] == 0) F.t_hd(r) = v;
      changed = true;
    }
    if (hd_cand && key < atomicMin(best + ROWS_PER_CTA + slot, key)) { F.t_hd(r) = thr(d2, ear, xnr); changed = true; }
    if (changed) F.update(r);
  };
  for (;;) {
    mbar_wait(&ring.full[cp.stage], cp.phase);
    const TileMeta m = ring.meta[cp.stage];
    if (m.row_block < 0) break;
    if (m.flags & 1u) {
      // first tile of an item: this warp prepares row group `warp` for everybody: keys warm-started from what the seeds
      // and earlier items found, ranks, box, and the group's bound
      R.load_group(g, (uint32_t) m.row_block, (uint32_t) warp, lane);
      float v = 0.f, vh = 0.f, lmax = 0.f;
#pragma unroll
      for (int r = 0; r < RI; ++r) {
        unsigned long long k0 = ~0ull, k1 = ~0ull;
        uint32_t lo_i = 0;
        float lf = 0.f;
        if (R.row(r) < g.row_end) {
          k0 = a.key_nn[R.out(r)];
          k1 = a.key_hd[R.out(r)];
          lo_i = __ldg(a.lo + R.row(r));
          lf = __ldg(a.lof + R.row(r));
          v = fmaxf(v, key_d2(k0));
          if (lo_i != 0) {
            vh = fmaxf(vh, key_d2(k1));
            lmax = fmaxf(lmax, lf + a.lo_bias);
          }
        }
        const uint32_t slot = (uint32_t) warp * (32u * RI) + (uint32_t) r * 32u + (uint32_t) lane;
        best[slot] = k0;
        best[ROWS_PER_CTA + slot] = k1;
        lo_s[slot] = lo_i;
        lor_s[slot] = lf + a.lo_bias;
      }
      if constexpr (D > 0) {
        WarpBox<D> wb;
        wb.compute(g, R, lane);
        if (lane < D) {
          gbox[warp * 2 * GBOX_DIMS + lane] = wb.wlo;
          gbox[warp * 2 * GBOX_DIMS + GBOX_DIMS + lane] = wb.whi;
        }
      }
      v = (fmaf(g.e_rel, v, v) + g.prune_slack) * 1.00001f;
      vh = (fmaf(g.e_rel, vh, vh) + g.prune_slack) * 1.00001f;
      if (!(v < INFINITY)) v = INFINITY;
      if (!(vh < INFINITY)) vh = INFINITY;
      const uint32_t vb = __reduce_max_sync(0xffffffffu, __float_as_uint(fmaxf(v, 0.f)));
      const uint32_t vhb = __reduce_max_sync(0xffffffffu, __float_as_uint(fmaxf(vh, 0.f)));
      const uint32_t lb = __reduce_max_sync(0xffffffffu, __float_as_uint(fmaxf(lmax, 0.f)));
      if (lane == 0) {
        gthr_nn[warp] = vb;
        gthr_hd[warp] = vhb;
        glor[warp] = __uint_as_float(lb);
        ring.wthr[warp] = ((unsigned long long) m.aux << 32) | max(vb, vhb);
      }
      consumer_barrier();
    }
    col0 = m.col0;
    if (!(m.flags & 4u)) {
      const float* tl = ring.tiles + cp.stage * ring.tile_floats;
      const float* cen = tl + (d + 1) * TJ;
      // Which groups need the tile?  A group accepts columns within its nearest-neighbour bound, and within its (for density
      // peaks inter-cluster sized) lower-free-energy bound only if the tile holds a frame of lower rank than its largest
      // one at all.  The first warp to get here decides for everybody, so that unit numbers mean the same to all.
      uint32_t reach;
      float lomin;
      {
        const float* lrow = cen + g.dp;
        if (TJ == 128) {
          const float4 q = *reinterpret_cast<const float4*>(lrow + 4 * lane);
          lomin = fminf(fminf(q.x, q.y), fminf(q.z, q.w));
        } else {
          const float2 q = *reinterpret_cast<const float2*>(lrow + 2 * lane);
          lomin = fminf(q.x, q.y);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) lomin = fminf(lomin, __shfl_xor_sync(0xffffffffu, lomin, o));
        const int gq = lane >> 2;
        float bound = __uint_as_float(*reinterpret_cast<volatile unsigned int*>(gthr_nn + gq));
        if (lomin < glor[gq]) bound = fmaxf(bound, __uint_as_float(*reinterpret_cast<volatile unsigned int*>(gthr_hd + gq)));
        const uint32_t mine = groups_in_reach<D>(gbox, cen, lane, bound);
        uint32_t seen = 0xffffffffu;
        if (lane == 0) seen = atomicCAS(&ring.umask[cp.stage], 0xffffffffu, mine);
        seen = __shfl_sync(0xffffffffu, seen, 0);
        reach = seen == 0xffffffffu ? mine : seen;
      }
      const uint32_t n_units = __popc(reach);
      for (;;) {
        uint32_t u = 0;
        if (lane == 0) u = atomicAdd(&ring.unext[cp.stage], 1u);
        u = __shfl_sync(0xffffffffu, u, 0);
        if (u >= n_units) break;
        const uint32_t gi = __fns(reach, 0, (int) u + 1) >> 2;
        slot0 = gi * (32u * RI) + (uint32_t) lane;
        R.load_group(g, (uint32_t) m.row_block, gi, lane);
        R.retarget(g, cen);
        if constexpr (D > 0) if (g.axis_prune) {
          // separating-axis test along the line between the centre of the group's box and the tile's centre, against what
          // the group's rows still accept (the bound of the moment: it may have tightened since the unit was listed)
          float bound = __uint_as_float(*reinterpret_cast<volatile unsigned int*>(gthr_nn + gi));
          if (lomin < glor[gi]) bound = fmaxf(bound, __uint_as_float(*reinterpret_cast<volatile unsigned int*>(gthr_hd + gi)));
          const float* blo = gbox + gi * 2 * GBOX_DIMS;
          float ax[D ? D : 1];
#pragma unroll
          for (int k = 0; k < D; ++k) ax[k] = 0.5f * (blo[k] + blo[GBOX_DIMS + k]) - (cen[k] - __ldg(g.centre + k));
          const float lbp = axis_lower_bound<D>(g, R, tl, ax, cen[D], slack_len, lane);
          if (lbp > 0.f && lbp * lbp * 0.999f > bound) continue;
        }
        ++st.wtiles;
        // filter thresholds of this unit from the group's best keys so far (the error margin depends on the tile)
#pragma unroll
        for (int r = 0; r < RI; ++r) {
          const uint32_t slot = slot0 + (uint32_t) r * 32u;
          F.lor(r) = lor_s[slot];
          F.xn(r) = R.xn[r];
          const float tn = thr(key_d2(*reinterpret_cast<volatile unsigned long long*>(best + slot)), R.ea(g, r), R.xn[r]);
          F.t_nn(r) = tn;
          // a frame nobody has a lower free energy than has no such neighbour: do not let it hold the filter open
          F.t_hd(r) = lo_s[slot] == 0 ? tn
                                      : thr(key_d2(*reinterpret_cast<volatile unsigned long long*>(best + ROWS_PER_CTA + slot)), R.ea(g, r), R.xn[r]);
          F.update(r);
        }
        if constexpr (D >= DCB_NN_WIDE_FROM) scan_tile_nn_wide<D>(g, tl, cen + g.dp, R, F, scratch, hit);
        else scan_tile_nn(g, tl, cen + g.dp, R, F, scratch, hit);
        // what the group's rows still accept (d2 units, pruning margins included): for the warps' reach tests, for the
        // producer's dynamic pruning of this item, and (at the end of the item) for the later items of the row block
        float v = 0.f, vh = 0.f;
#pragma unroll
        for (int r = 0; r < RI; ++r)
          if (R.row(r) < g.row_end) {
            v = fmaxf(v, (F.t_nn(r) + F.xn(r)) * 1.000001f + g.prune_slack);     // fmaxf drops NaN
            if (lo_s[slot0 + (uint32_t) r * 32u] != 0) vh = fmaxf(vh, (F.t_hd(r) + F.xn(r)) * 1.000001f + g.prune_slack);
          }
        if (!(v < INFINITY)) v = INFINITY;
        if (!(vh < INFINITY)) vh = INFINITY;
        const uint32_t vb = __reduce_max_sync(0xffffffffu, __float_as_uint(fmaxf(v, 0.f)));   // v >= 0: bits order like values
        const uint32_t vhb = __reduce_max_sync(0xffffffffu, __float_as_uint(fmaxf(vh, 0.f)));
        if (lane == 0) {
          const uint32_t now = min(atomicMin(gthr_nn + gi, vb), vb);
          const uint32_t nowh = min(atomicMin(gthr_hd + gi, vhb), vhb);
          ring.wthr[gi] = ((unsigned long long) m.aux << 32) | max(now, nowh);
        }
      }
    }
    __syncwarp();
    stage_release(cp.stage);
    if (m.flags & 2u) {
      // end of the item: every unit is done once all consumer warps are here; this warp writes group `warp` back
      consumer_barrier();
#pragma unroll
      for (int r = 0; r < RI; ++r) {
        const uint32_t slot = (uint32_t) warp * (32u * RI) + (uint32_t) r * 32u + (uint32_t) lane;
        const uint32_t row = block_row0(g, (uint32_t) m.row_block) + slot;
        if (row < g.row_end) {
          atomicMin(a.key_nn + ((uint32_t) m.row_block * ROWS_PER_CTA + slot), best[slot]);
          atomicMin(a.key_hd + ((uint32_t) m.row_block * ROWS_PER_CTA + slot), best[ROWS_PER_CTA + slot]);
        }
      }
      // what this group's rows still accept bounds every later item of the row block (positive floats order like their bits)
      if (lane == 0)
        atomicMin(reinterpret_cast<unsigned int*>(g.blk_thr) + (size_t) m.row_block * N_CONSUMER_WARPS + warp, max(gthr_nn[warp], gthr_hd[warp]));
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
  float thr_fast;           // cut (1 + e_rel)
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
    produce<D>(g, ring, false, [&](uint32_t rb, uint32_t&, uint32_t& lim1) {
      // only columns below the last row of the block can form an edge (j < i)
      const uint32_t last_row = min(block_row0(g, rb) + (uint32_t) ROWS_PER_CTA, g.row_end) - 1;
      lim1 = min(lim1, last_row / TJ + 1);
    }, [&](uint32_t, int) { return g.prune_thr; });
    return;
  }
  const int tid = threadIdx.x;
  float* scratch = reinterpret_cast<float*>(smem + ((SmemRing<D>::bytes(d) + 15) & ~size_t(15))) + threadIdx.x;
  Rows<D> R;
  WarpBox<D> wb;
  float t[RI];
  Pipe<StagesOf<D>::n> cp;
  SlowStats st;
  uint32_t col0 = 0;
  auto hit = [&](int r, int jt, float accv) {
    const uint32_t j = col0 + jt;
    const uint32_t i = R.row(r);
    if (j >= i || i >= g.row_end) return;
    ++st.slow;
    float s = accv + sel4(R.xn, r);
    const float e = fmaf(g.e_rel, fabsf(s), R.ea(g, r)) * 1.0001f;
    if (fabsf(s - a.cut) <= e) {
      s = dist2_exact(g.xT, g.ld, d, i, j);
      ++st.exact;
    }
    if (s < a.cut) uf_union(a.parent, i, j);
  };
  for (;;) {
    mbar_wait(&ring.full[cp.stage], cp.phase);
    const TileMeta m = ring.meta[cp.stage];
    if (m.row_block < 0) break;
    if (m.flags & 1u) {
      R.load(g, (uint32_t) m.row_block, tid);
      wb.compute(g, R, lane);
    }
    col0 = m.col0;
    if (!(m.flags & 4u) && wb.reach(ring.tiles + cp.stage * ring.tile_floats + (d + 1) * TJ, lane, g.prune_thr)) {
      const float* tl = ring.tiles + cp.stage * ring.tile_floats;
      ++st.wtiles;
      R.retarget(g, tl + (d + 1) * TJ);
#pragma unroll
      for (int r = 0; r < RI; ++r) t[r] = next_up(next_up(a.thr_fast + R.ea(g, r) - R.xn[r]));
      scan_tile(g, tl, R, t, scratch, hit);
    }
    __syncwarp();
    stage_release(cp.stage);
    cp.advance();
  }
  st.flush(g);
}


// ================================================================================================
// screening as ONE neighbour-graph scan: all pairs {i, j} with d2 < cut of frames held in spatial order
// ================================================================================================
// The incremental scan above (screen_kernel) needs the frames in free-energy order, where tiles have no spatial
// coherence: nothing is pruned and every threshold costs (new rows) x (all lower frames) -- N^2 / 2 pairs over a whole
// screening run (ncu at 5M x 3: 50 ms per threshold at 49 % of the FFMA pipe, 3.1 s for 106 thresholds).  The clusters of
// EVERY threshold follow from one list of edges instead: the pair {a, b} (free-energy-sorted indices, a > b) with
// d2 < cut joins the graph exactly when frame a does, i.e. at the first threshold with more than a frames.  This kernel
// scans the frames in the context's SPATIAL order (tiles pruned by the cut radius like a population scan, upper
// triangle of the position matrix only) and emits (a << 32 | b) for every such pair; api.cu sorts the list by a and feeds
// the union-find threshold by threshold.  rank[p] = free-energy-sorted index of the frame at position p.
struct EdgeArgs {
  ScanGeom g;
  float cut;                      // (float)(4*sigma2); an edge needs d2 < cut (density_clustering.cpp:319)
  float thr_fast;                 // cut (1 + e_rel)
  const uint32_t* rank;           // [n] by position
  uint32_t level_min;             // edges whose larger index is below this are not wanted (both ends settled already)
  unsigned long long* edges;      // [cap]
  unsigned long long cap;
  unsigned long long* count;      // edges found; keeps counting past cap (api.cu then scans again with room for all)
};

template <int D>
__global__ void DCB_LAUNCH_BOUNDS(D) edge_kernel(const __grid_constant__ EdgeArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  constexpr int TJ = TileW<D>::tj;
  const ScanGeom& g = a.g;
  const int d = D ? D : g.d;
  SmemRing<D> ring(smem, d);
  ring.init();
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == N_CONSUMER_WARPS) {
    produce<D>(g, ring, false, [&](uint32_t rb, uint32_t& lim0, uint32_t&) {
      // upper triangle of the position matrix: a pair is found from its lower position, so only tiles from the block's own on
      lim0 = block_row0(g, rb) / TJ;
    }, [&](uint32_t, int) { return g.prune_thr; });
    return;
  }
  const int tid = threadIdx.x;
  float* scratch = reinterpret_cast<float*>(smem + ((SmemRing<D>::bytes(d) + 15) & ~size_t(15))) + threadIdx.x;
  Rows<D> R;
  WarpBox<D> wb;
  float t[RI];
  Pipe<StagesOf<D>::n> cp;
  SlowStats st;
  uint32_t col0 = 0;
  auto hit = [&](int r, int jt, float accv) {
    const uint32_t j = col0 + jt;
    const uint32_t i = R.row(r);
    if (j <= i || j >= g.n || i >= g.row_end) return;
    ++st.slow;
    float s = accv + sel4(R.xn, r);
    const float e = fmaf(g.e_rel, fabsf(s), R.ea(g, r)) * 1.0001f;
    if (fabsf(s - a.cut) <= e) {
      s = dist2_exact(g.xT, g.ld, d, i, j);
      ++st.exact;
    }
    if (!(s < a.cut)) return;
    const uint32_t ri = __ldg(a.rank + i), rj = __ldg(a.rank + j);
    const uint32_t hi = max(ri, rj), lo = min(ri, rj);
    if (hi < a.level_min) return;
    // the lanes that reach this point together reserve their slots with one atomic
    const uint32_t peers = __activemask();
    const int leader = __ffs(peers) - 1;
    unsigned long long base = 0;
    if (lane == leader) base = atomicAdd(a.count, (unsigned long long) __popc(peers));
    base = __shfl_sync(peers, base, leader);
    const unsigned long long slot = base + (unsigned long long) __popc(peers & ((1u << lane) - 1u));
    if (slot < a.cap) a.edges[slot] = ((unsigned long long) hi << 32) | lo;
  };
  for (;;) {
    mbar_wait(&ring.full[cp.stage], cp.phase);
    const TileMeta m = ring.meta[cp.stage];
    if (m.row_block < 0) break;
    if (m.flags & 1u) {
      R.load(g, (uint32_t) m.row_block, tid);
      wb.compute(g, R, lane);
    }
    col0 = m.col0;
    // tiles below the warp's own rows hold only lower positions: the other orientation finds those pairs
    const bool ahead = m.col0 + (uint32_t) TJ > R.row0 - (uint32_t) lane;
    if (!(m.flags & 4u) && ahead && wb.reach(ring.tiles + cp.stage * ring.tile_floats + (d + 1) * TJ, lane, g.prune_thr)) {
      const float* tl = ring.tiles + cp.stage * ring.tile_floats;
      ++st.wtiles;
      R.retarget(g, tl + (d + 1) * TJ);
#pragma unroll
      for (int r = 0; r < RI; ++r) t[r] = next_up(next_up(a.thr_fast + R.ea(g, r) - R.xn[r]));
      scan_tile(g, tl, R, t, scratch, hit);
    }
    __syncwarp();
    stage_release(cp.stage);
    cp.advance();
  }
  st.flush(g);
}

}  // namespace dcb
