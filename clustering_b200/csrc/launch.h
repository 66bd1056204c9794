// Host-callable launchers of the per-dimension kernel instantiations (kern_inst.cu).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace dcb {
struct PopsArgs;
struct NnArgs;
struct ScreenArgs;
struct EdgeArgs;

#define DCB_FOR_EACH_D(X) X(0) X(1) X(2) X(3) X(4) X(5) X(6) X(7) X(8) X(9) X(10) X(11) X(12) X(13) X(14) X(15) X(16)

#define DCB_DECL(D)                                                          \
  cudaError_t launch_pops_d##D(const PopsArgs&, int grid, cudaStream_t st);  \
  cudaError_t launch_nn_d##D(const NnArgs&, int grid, cudaStream_t st);      \
  int occupancy_pops_d##D(int n_bins, int d);                                \
  cudaError_t launch_pops_count_d##D(const PopsArgs&, int grid, cudaStream_t st);  \
  int occupancy_pops_count_d##D(int n_bins, int d);                          \
  cudaError_t launch_pops_bin_d##D(const PopsArgs&, int grid, cudaStream_t st);    \
  int occupancy_pops_bin_d##D(int n_bins, int lut_k, int d);                 \
  int occupancy_nn_d##D(int d);                                              \
  cudaError_t launch_screen_d##D(const ScreenArgs&, int grid, cudaStream_t st); \
  int occupancy_screen_d##D(int d);                                          \
  cudaError_t launch_edge_d##D(const EdgeArgs&, int grid, cudaStream_t st);  \
  int occupancy_edge_d##D(int d);
DCB_FOR_EACH_D(DCB_DECL)
#undef DCB_DECL

// GEMM-form (tcgen05) path, gemm_inst.cu
struct GPopsArgs;
struct GNnArgs;
size_t gemm_smem_bytes(int kc, int n_stages);
cudaError_t launch_gpack(const float* coords, size_t n, int d, int kc, size_t n_tiles, const float* centre, const uint32_t* perm, float* gT,
                         float* gnorm, float* xR, float* tcen, float* tlo, float* thi, float* trad, cudaStream_t st);
cudaError_t launch_tile_lb(const float* tcen, const float* tlo, const float* thi, const float* trad, int d, size_t n_tiles, uint32_t s0,
                           uint32_t s1, float* lb, cudaStream_t st);
cudaError_t launch_tile_min(const float* lof, size_t n_tiles, float* lomin, cudaStream_t st);
cudaError_t launch_tf32_peak(int grid, int iters, long long* out, cudaStream_t st);
cudaError_t launch_gpops(const GPopsArgs& a, int grid, bool check, cudaStream_t st);
cudaError_t launch_gnn(const GNnArgs& a, int grid, cudaStream_t st);
cudaError_t launch_gnn_tile_thr(const unsigned long long* key_nn, const unsigned long long* key_hd, const uint32_t* lo, const float* lof,
                                float lo_bias, uint32_t row_begin, uint32_t row_end, uint32_t n_row_tiles, float e_rel, float slack,
                                float* thr_nn, float* thr_hd, float* lormax, cudaStream_t st);
}  // namespace dcb
