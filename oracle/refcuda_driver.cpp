// TEST / BENCHMARK INFRASTRUCTURE ONLY -- never linked into the product library.
//
// extern "C" driver around the *unmodified* reference CUDA implementation of the `clustering density` hot path
// (density_clustering_cuda.cu + density_clustering_cuda_kernels.cu, compiled where they lie under $REF_SRC by
// oracle/Makefile target `refcuda` with the architecture flag replaced: the shipped -arch=compute_30 is rejected by
// nvcc 12.9, the sources themselves build unchanged for sm_100a).  It is the "kernel to beat" comparator that
// BASELINE.json's north_star asks to report next to the OpenMP host path.  Note the reference's CUDA path has its own
// semantics (d2 <= r2, self counted through d2 = 0, duplicates excluded from the neighbour search):
// it is timed here, not used as a parity oracle.
//
// Wrapped reference entry points (file:line under /root/reference/src):
//   CUDA::get_num_gpus            density_clustering_cuda.cu:32-43
//   CUDA::calculate_populations   density_clustering_cuda.cu:139-182
//   CUDA::nearest_neighbors       density_clustering_cuda.cu:286-328
#include <cstdint>
#include <map>
#include <tuple>
#include <vector>

#include "density_clustering_cuda.hpp"

extern "C" {

int dcrefcuda_num_gpus() { return Clustering::Density::CUDA::get_num_gpus(); }

// pops_out: [n_radii][n_rows] in the order of the input radii
void dcrefcuda_populations(const float* coords, uint64_t n_rows, uint64_t n_cols, const float* radii, uint64_t n_radii, uint64_t* pops_out) {
  std::vector<float> r(radii, radii + n_radii);
  Clustering::Density::Pops pops = Clustering::Density::CUDA::calculate_populations(coords, n_rows, n_cols, r);
  for (uint64_t q = 0; q < n_radii; ++q) {
    const std::vector<std::size_t>& p = pops[radii[q]];
    for (uint64_t i = 0; i < n_rows; ++i) pops_out[q * n_rows + i] = p[i];
  }
}

void dcrefcuda_nearest_neighbors(const float* coords, uint64_t n_rows, uint64_t n_cols, const float* fe, uint64_t* nn_idx, float* nn_d2,
                                 uint64_t* hd_idx, float* hd_d2) {
  std::vector<float> f(fe, fe + n_rows);
  auto res = Clustering::Density::CUDA::nearest_neighbors(coords, n_rows, n_cols, f);
  const auto& nh = std::get<0>(res);
  const auto& hd = std::get<1>(res);
  for (uint64_t i = 0; i < n_rows; ++i) {
    const auto a = nh.find(i);
    const auto b = hd.find(i);
    nn_idx[i] = a != nh.end() ? a->second.first : n_rows + 1;
    nn_d2[i] = a != nh.end() ? a->second.second : 0.f;
    hd_idx[i] = b != hd.end() ? b->second.first : n_rows + 1;
    hd_d2[i] = b != hd.end() ? b->second.second : 0.f;
  }
}

}  // extern "C"
