// Host-callable launchers of the GEMM-form (tcgen05) kernels of gemm_kernels.cuh.
#define DCB_GEMM_KERNELS
#include "gemm_kernels.cuh"
#include "launch.h"

namespace dcb {

size_t gemm_smem_bytes(int kc, int n_stages) { return GSmem::bytes(kc, n_stages); }

cudaError_t launch_gpack(const float* coords, size_t n, int d, int kc, size_t n_tiles, const float* centre, const uint32_t* perm, float* gT,
                         float* gnorm, float* xR, float* tcen, float* tlo, float* thi, float* trad, cudaStream_t st) {
  const int pitch = d + 1 + (d & 1);
  const size_t smem = ((size_t) GT * pitch + d) * sizeof(float);
  cudaError_t e = cudaFuncSetAttribute(gpack_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
  if (e != cudaSuccess) return e;
  gpack_kernel<<<(unsigned int) n_tiles, GT, smem, st>>>(coords, n, d, kc, n_tiles, centre, perm, gT, gnorm, xR, tcen, tlo, thi, trad);
  return cudaGetLastError();
}

cudaError_t launch_tile_lb(const float* tcen, const float* tlo, const float* thi, const float* trad, int d, size_t n_tiles, uint32_t s0,
                           uint32_t s1, float* lb, cudaStream_t st) {
  if (s1 <= s0) return cudaSuccess;
  const dim3 grid((unsigned int) ((n_tiles + 127) / 128), (s1 - s0 + LB_ROWS - 1) / LB_ROWS);
  tile_lb_kernel<<<grid, 128, (size_t) LB_ROWS * 3 * d * sizeof(float), st>>>(tcen, tlo, thi, trad, d, n_tiles, s0, s1, lb);
  return cudaGetLastError();
}

cudaError_t launch_tile_min(const float* lof, size_t n_tiles, float* lomin, cudaStream_t st) {
  tile_min_kernel<<<(unsigned int) ((n_tiles + 7) / 8), 256, 0, st>>>(lof, n_tiles, lomin);
  return cudaGetLastError();
}

template <int NB, bool CHECK>
static cudaError_t launch_gpops_nb(const GPopsArgs& a, int grid, cudaStream_t st) {
  const size_t smem = GSmem::bytes(a.g.ra * a.g.kc, a.g.n_stages);
  cudaError_t e = cudaFuncSetAttribute(gscan_pops_kernel<NB, CHECK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
  if (e != cudaSuccess) return e;
  gscan_pops_kernel<NB, CHECK><<<grid, G_THREADS, smem, st>>>(a);
  return cudaGetLastError();
}

cudaError_t launch_gpops(const GPopsArgs& a, int grid, bool check, cudaStream_t st) {
  if (check) return a.n_bins == 1 ? launch_gpops_nb<1, true>(a, grid, st) : cudaErrorInvalidValue;
  switch (a.n_bins) {
    case 1: return launch_gpops_nb<1, false>(a, grid, st);
    case 2: return launch_gpops_nb<2, false>(a, grid, st);
    case 4: return launch_gpops_nb<4, false>(a, grid, st);
    case 8: return launch_gpops_nb<8, false>(a, grid, st);
  }
  return cudaErrorInvalidValue;
}

cudaError_t launch_gnn(const GNnArgs& a, int grid, cudaStream_t st) {
  const size_t smem = GSmem::bytes(a.g.ra * a.g.kc, a.g.n_stages);
  cudaError_t e = cudaFuncSetAttribute(gscan_nn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
  if (e != cudaSuccess) return e;
  gscan_nn_kernel<<<grid, G_THREADS, smem, st>>>(a);
  return cudaGetLastError();
}

cudaError_t launch_gnn_tile_thr(const unsigned long long* key_nn, const unsigned long long* key_hd, const uint32_t* lo, const float* lof,
                                float lo_bias, uint32_t row_begin, uint32_t row_end, uint32_t n_row_tiles, float e_rel, float slack,
                                float* thr_nn, float* thr_hd, float* lormax, cudaStream_t st) {
  gnn_tile_thr_kernel<<<n_row_tiles, GT, 0, st>>>(key_nn, key_hd, lo, lof, lo_bias, row_begin, row_end, e_rel, slack, thr_nn, thr_hd, lormax);
  return cudaGetLastError();
}

cudaError_t launch_tf32_peak(int grid, int iters, long long* out, cudaStream_t st) {
  const size_t smem = 1024 + (size_t) G_CHUNK_BYTES * 8;
  cudaError_t e = cudaFuncSetAttribute(tf32_peak_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
  if (e != cudaSuccess) return e;
  tf32_peak_kernel<<<grid, 64, smem, st>>>(iters, out);
  return cudaGetLastError();
}

}  // namespace dcb
