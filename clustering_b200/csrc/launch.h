// Host-callable launchers of the per-dimension kernel instantiations (kern_inst.cu).
#pragma once
#include <cuda_runtime.h>

namespace dcb {
struct PopsArgs;
struct NnArgs;
struct ScreenArgs;

#define DCB_FOR_EACH_D(X) X(0) X(1) X(2) X(3) X(4) X(5) X(6) X(7) X(8) X(9) X(10) X(11) X(12) X(13) X(14) X(15) X(16)

#define DCB_DECL(D)                                                          \
  cudaError_t launch_pops_d##D(const PopsArgs&, int grid, cudaStream_t st);  \
  cudaError_t launch_nn_d##D(const NnArgs&, int grid, cudaStream_t st);      \
  int occupancy_pops_d##D(int n_bins, int d);                                \
  cudaError_t launch_pops_count_d##D(const PopsArgs&, int grid, cudaStream_t st);  \
  int occupancy_pops_count_d##D(int n_bins, int d);                          \
  int occupancy_nn_d##D(int d);                                              \
  cudaError_t launch_screen_d##D(const ScreenArgs&, int grid, cudaStream_t st); \
  int occupancy_screen_d##D(int d);
DCB_FOR_EACH_D(DCB_DECL)
#undef DCB_DECL
}  // namespace dcb
