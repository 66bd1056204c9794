// TEST INFRASTRUCTURE ONLY -- never linked into the product library.
//
// extern "C" driver around the *unmodified* reference CPU/OpenMP implementation of the
// `clustering density` hot path.  It is compiled together with the reference's own translation
// units where they lie under $REF_SRC (default /root/reference/src) by oracle/Makefile into
// oracle/_ref/libdcref.so.  Nothing from the reference is copied into this repository; this file
// only converts between plain arrays and the reference's std containers.
//
// Wrapped reference entry points (file:line under /root/reference/src):
//   calculate_populations (multi radius)   density_clustering.cpp:126-195
//   calculate_free_energies                density_clustering.cpp:197-212
//   sorted_free_energies                   density_clustering.cpp:214-228
//   nearest_neighbors                      density_clustering.cpp:230-288
//   compute_sigma2                         density_clustering.cpp:334-343
//   assign_low_density_frames              density_clustering.cpp:345-360
//   sorted_cluster_names                   density_clustering.cpp:458-493
//   screening                              density_clustering_common.cpp:37-134
//   Tools::write_pops/write_fes/write_neighborhood/write_clustered_trajectory  tools.cpp:42-56,144-174
#include <cstdint>
#include <cstring>
#include <map>
#include <tuple>
#include <vector>
#include <string>

#include "density_clustering.hpp"
#include "density_clustering_common.hpp"
#include "logger.hpp"
#include "tools.hpp"

#include <omp.h>

namespace {
using Clustering::Tools::Neighbor;
using Clustering::Tools::Neighborhood;

// the reference promises the compiler 32-byte alignment (ASSUME_ALIGNED); honour it.
struct AlignedCoords {
  float* p;
  AlignedCoords(const float* src, std::size_t n) {
    p = (float*) _mm_malloc((n ? n : 1) * sizeof(float), DC_MEM_ALIGNMENT);
    std::memcpy(p, src, n * sizeof(float));
  }
  ~AlignedCoords() { _mm_free(p); }
};

Neighborhood make_nh(const uint64_t* idx, const float* d2, std::size_t n) {
  Neighborhood nh;
  for (std::size_t i = 0; i < n; ++i) nh.emplace_hint(nh.end(), i, Neighbor(idx[i], d2[i]));
  return nh;
}
}  // namespace

extern "C" {

void dcref_set_threads(int n) { omp_set_num_threads(n); }
int dcref_max_threads() { return omp_get_max_threads(); }

// pops_out: [n_radii][n_rows], row r = input radii[r] (map lookup by float key, as the driver does)
void dcref_populations(const float* coords, uint64_t n_rows, uint64_t n_cols,
                       const float* radii, uint64_t n_radii, uint64_t* pops_out) {
  AlignedCoords c(coords, n_rows * n_cols);
  std::vector<float> rv(radii, radii + n_radii);
  auto pops = Clustering::Density::calculate_populations(c.p, n_rows, n_cols, rv);
  for (uint64_t r = 0; r < n_radii; ++r) {
    const auto& v = pops[radii[r]];
    for (uint64_t i = 0; i < n_rows; ++i) pops_out[r * n_rows + i] = v[i];
  }
}

void dcref_free_energies(const uint64_t* pops, uint64_t n, float* fe_out) {
  std::vector<std::size_t> p(pops, pops + n);
  auto fe = Clustering::Density::calculate_free_energies(p);
  std::memcpy(fe_out, fe.data(), n * sizeof(float));
}

// order_out[k] = original frame index at sorted position k (libstdc++ std::sort tie order)
void dcref_sorted_free_energies(const float* fe, uint64_t n, uint64_t* order_out) {
  std::vector<float> f(fe, fe + n);
  auto s = Clustering::Density::sorted_free_energies(f);
  for (uint64_t k = 0; k < n; ++k) order_out[k] = s[k].first;
}

void dcref_nearest_neighbors(const float* coords, uint64_t n_rows, uint64_t n_cols, const float* fe,
                             uint64_t* nn_idx, float* nn_d2, uint64_t* hd_idx, float* hd_d2) {
  AlignedCoords c(coords, n_rows * n_cols);
  std::vector<float> f(fe, fe + n_rows);
  auto t = Clustering::Density::nearest_neighbors(c.p, n_rows, n_cols, f);
  const Neighborhood& nh = std::get<0>(t);
  const Neighborhood& hd = std::get<1>(t);
  for (uint64_t i = 0; i < n_rows; ++i) {
    const auto& a = nh.at(i);
    const auto& b = hd.at(i);
    nn_idx[i] = a.first; nn_d2[i] = a.second;
    hd_idx[i] = b.first; hd_d2[i] = b.second;
  }
}

double dcref_sigma2(const uint64_t* nn_idx, const float* nn_d2, uint64_t n) {
  return Clustering::Density::compute_sigma2(make_nh(nn_idx, nn_d2, n));
}

// initial: NULL (first threshold) or n_rows labels of the previous threshold.
void dcref_screening(const float* fe, const uint64_t* nn_idx, const float* nn_d2, float threshold,
                     const float* coords, uint64_t n_rows, uint64_t n_cols,
                     const uint64_t* initial, uint64_t* labels_out) {
  AlignedCoords c(coords, n_rows * n_cols);
  std::vector<float> f(fe, fe + n_rows);
  Neighborhood nh = make_nh(nn_idx, nn_d2, n_rows);
  std::vector<std::size_t> init;
  if (initial) init.assign(initial, initial + n_rows);
  auto lab = Clustering::Density::screening(f, nh, threshold, c.p, n_rows, n_cols, init);
  for (uint64_t i = 0; i < n_rows; ++i) labels_out[i] = lab[i];
}

void dcref_assign_low_density_frames(const uint64_t* initial, const uint64_t* hd_idx, const float* hd_d2,
                                     const float* fe, uint64_t n, uint64_t* out) {
  std::vector<std::size_t> init(initial, initial + n);
  std::vector<float> f(fe, fe + n);
  auto r = Clustering::Density::assign_low_density_frames(init, make_nh(hd_idx, hd_d2, n), f);
  for (uint64_t i = 0; i < n; ++i) out[i] = r[i];
}

void dcref_sorted_cluster_names(const uint64_t* clustering, uint64_t n, uint64_t* out) {
  std::vector<std::size_t> c(clustering, clustering + n);
  auto r = Clustering::Density::sorted_cluster_names(c);
  for (uint64_t i = 0; i < n; ++i) out[i] = r[i];
}

// ---- writers (byte formats of the output files; used by the file-format parity tests) ----
static std::map<std::string, float> make_comments(const char** keys, const float* vals, int n) {
  std::map<std::string, float> m;
  for (int i = 0; i < n; ++i) m[keys[i]] = vals[i];
  return m;
}

void dcref_write_pops(const char* fname, const uint64_t* pops, uint64_t n, const char* header,
                      const char** keys, const float* vals, int n_comments) {
  std::vector<std::size_t> p(pops, pops + n);
  Clustering::Tools::write_pops(fname, p, header, make_comments(keys, vals, n_comments));
}

void dcref_write_fes(const char* fname, const float* fe, uint64_t n, const char* header,
                     const char** keys, const float* vals, int n_comments) {
  std::vector<float> f(fe, fe + n);
  Clustering::Tools::write_fes(fname, f, header, make_comments(keys, vals, n_comments));
}

void dcref_write_neighborhood(const char* fname, const uint64_t* nn_idx, const float* nn_d2,
                              const uint64_t* hd_idx, const float* hd_d2, uint64_t n, const char* header,
                              const char** keys, const float* vals, int n_comments) {
  Clustering::Tools::write_neighborhood(fname, make_nh(nn_idx, nn_d2, n), make_nh(hd_idx, hd_d2, n), header,
                                        make_comments(keys, vals, n_comments));
}

void dcref_write_clustered_trajectory(const char* fname, const uint64_t* traj, uint64_t n, const char* header,
                                      const char** keys, const float* vals, int n_comments) {
  std::vector<std::size_t> t(traj, traj + n);
  Clustering::Tools::write_clustered_trajectory(fname, t, header, make_comments(keys, vals, n_comments));
}

// reads coords with the reference's reader; returns rows/cols, copies up to cap floats.
int dcref_read_coords(const char* fname, float* out, uint64_t cap, uint64_t* n_rows, uint64_t* n_cols) {
  float* c; std::size_t r, k;
  std::tie(c, r, k) = Clustering::Tools::read_coords<float>(fname);
  *n_rows = r; *n_cols = k;
  int ok = (r * k <= cap);
  if (ok) std::memcpy(out, c, r * k * sizeof(float));
  Clustering::Tools::free_coords(c);
  return ok ? 0 : 1;
}

}  // extern "C"
