// One translation unit per specialised dimension: compiled with -DDCB_D=<n_cols> (0 = run-time D).
#include "kernels.cuh"
#include "launch.h"

#ifndef DCB_D
#error "compile with -DDCB_D=<n>"
#endif
#define DCB_CAT2(a, b) a##b
#define DCB_CAT(a, b) DCB_CAT2(a, b)

namespace dcb {

cudaError_t DCB_CAT(launch_pops_d, DCB_D)(const PopsArgs& a, int grid, cudaStream_t st) {
  const size_t smem = pops_smem_bytes(SmemRing<DCB_D>::bytes(a.g.d), a.n_bins);
  cudaError_t e = cudaFuncSetAttribute(pops_kernel<DCB_D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
  if (e != cudaSuccess) return e;
  pops_kernel<DCB_D><<<grid, CTA_THREADS, smem, st>>>(a);
  return cudaGetLastError();
}

#if DCB_D >= 1
// count mode: 1, 2, 3, 4 (and for D <= 6 also 6 or 8) distinct radii per pass
#if DCB_D <= 6
#define DCB_COUNT_NB(X) X(1) X(2) X(3) X(4) X(6) X(8)
#else
#define DCB_COUNT_NB(X) X(1) X(2) X(3) X(4)
#endif
cudaError_t DCB_CAT(launch_pops_count_d, DCB_D)(const PopsArgs& a, int grid, cudaStream_t st) {
  const size_t smem = pops_count_smem_bytes(SmemRing<DCB_D>::bytes(a.g.d), a.n_bins);
  cudaError_t e = cudaErrorInvalidValue;
  switch (a.n_bins) {
#define DCB_CASE(NB)                                                                                                  \
  case NB:                                                                                                            \
    e = cudaFuncSetAttribute(pops_count_kernel<DCB_D, NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);  \
    if (e != cudaSuccess) return e;                                                                                   \
    pops_count_kernel<DCB_D, NB><<<grid, CTA_THREADS, smem, st>>>(a);                                                 \
    break;
    DCB_COUNT_NB(DCB_CASE)
#undef DCB_CASE
    default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}
int DCB_CAT(occupancy_pops_count_d, DCB_D)(int n_bins, int d) {
  int nb = 0;
  const size_t smem = pops_count_smem_bytes(SmemRing<DCB_D>::bytes(d), n_bins);
  switch (n_bins) {
#define DCB_CASE(NB)                                                                                              \
  case NB:                                                                                                        \
    cudaFuncSetAttribute(pops_count_kernel<DCB_D, NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);  \
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, pops_count_kernel<DCB_D, NB>, CTA_THREADS, smem);          \
    break;
    DCB_COUNT_NB(DCB_CASE)
#undef DCB_CASE
    default: break;
  }
  return nb;
}
// bin mode: long radius lists through the cell table
cudaError_t DCB_CAT(launch_pops_bin_d, DCB_D)(const PopsArgs& a, int grid, cudaStream_t st) {
  const size_t smem = pops_bin_smem_bytes(SmemRing<DCB_D>::bytes(a.g.d), a.n_bins, a.lut_k);
  cudaError_t e = cudaFuncSetAttribute(pops_bin_kernel<DCB_D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
  if (e != cudaSuccess) return e;
  pops_bin_kernel<DCB_D><<<grid, CTA_THREADS, smem, st>>>(a);
  return cudaGetLastError();
}
int DCB_CAT(occupancy_pops_bin_d, DCB_D)(int n_bins, int lut_k, int d) {
  int nb = 0;
  const size_t smem = pops_bin_smem_bytes(SmemRing<DCB_D>::bytes(d), n_bins, lut_k);
  if (cudaFuncSetAttribute(pops_bin_kernel<DCB_D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, pops_bin_kernel<DCB_D>, CTA_THREADS, smem);
  return nb;
}
#else
cudaError_t DCB_CAT(launch_pops_count_d, DCB_D)(const PopsArgs&, int, cudaStream_t) { return cudaErrorInvalidValue; }
int DCB_CAT(occupancy_pops_count_d, DCB_D)(int, int) { return 0; }
cudaError_t DCB_CAT(launch_pops_bin_d, DCB_D)(const PopsArgs&, int, cudaStream_t) { return cudaErrorInvalidValue; }
int DCB_CAT(occupancy_pops_bin_d, DCB_D)(int, int, int) { return 0; }
#endif

cudaError_t DCB_CAT(launch_nn_d, DCB_D)(const NnArgs& a, int grid, cudaStream_t st) {
  const size_t smem = nn_smem_bytes(SmemRing<DCB_D>::bytes(a.g.d, 1));
  cudaError_t e = cudaFuncSetAttribute(nn_kernel<DCB_D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
  if (e != cudaSuccess) return e;
  nn_kernel<DCB_D><<<grid, CTA_THREADS, smem, st>>>(a);
  return cudaGetLastError();
}

cudaError_t DCB_CAT(launch_screen_d, DCB_D)(const ScreenArgs& a, int grid, cudaStream_t st) {
  const size_t smem = screen_smem_bytes(SmemRing<DCB_D>::bytes(a.g.d));
  cudaError_t e = cudaFuncSetAttribute(screen_kernel<DCB_D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
  if (e != cudaSuccess) return e;
  screen_kernel<DCB_D><<<grid, CTA_THREADS, smem, st>>>(a);
  return cudaGetLastError();
}

cudaError_t DCB_CAT(launch_edge_d, DCB_D)(const EdgeArgs& a, int grid, cudaStream_t st) {
  const size_t smem = screen_smem_bytes(SmemRing<DCB_D>::bytes(a.g.d));
  cudaError_t e = cudaFuncSetAttribute(edge_kernel<DCB_D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
  if (e != cudaSuccess) return e;
  edge_kernel<DCB_D><<<grid, CTA_THREADS, smem, st>>>(a);
  return cudaGetLastError();
}
int DCB_CAT(occupancy_edge_d, DCB_D)(int d) {
  int nb = 0;
  const size_t smem = screen_smem_bytes(SmemRing<DCB_D>::bytes(d));
  cudaFuncSetAttribute(edge_kernel<DCB_D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, edge_kernel<DCB_D>, CTA_THREADS, smem);
  return nb;
}

int DCB_CAT(occupancy_pops_d, DCB_D)(int n_bins, int d) {
  int nb = 0;
  const size_t smem = pops_smem_bytes(SmemRing<DCB_D>::bytes(d), n_bins);
  cudaFuncSetAttribute(pops_kernel<DCB_D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, pops_kernel<DCB_D>, CTA_THREADS, smem);
  return nb;
}
int DCB_CAT(occupancy_nn_d, DCB_D)(int d) {
  int nb = 0;
  const size_t smem = nn_smem_bytes(SmemRing<DCB_D>::bytes(d, 1));
  cudaFuncSetAttribute(nn_kernel<DCB_D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, nn_kernel<DCB_D>, CTA_THREADS, smem);
  return nb;
}

int DCB_CAT(occupancy_screen_d, DCB_D)(int d) {
  int nb = 0;
  const size_t smem = screen_smem_bytes(SmemRing<DCB_D>::bytes(d));
  cudaFuncSetAttribute(screen_kernel<DCB_D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, screen_kernel<DCB_D>, CTA_THREADS, smem);
  return nb;
}

}  // namespace dcb
