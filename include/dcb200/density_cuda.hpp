// Drop-in replacement of the reference's density_clustering_cuda.hpp (C++11, header-only over libdcb200.so).
//
// Same namespace, names, argument meaning and error behaviour as Clustering::Density::CUDA in
// moldyn/Clustering (reference: src/density_clustering_cuda.hpp:13-54, implementation
// src/density_clustering_cuda.cu); the call sites in src/density_clustering.cpp (:113-118, :616-621,
// :659-663, :716-720, :746-750, :808-814) and src/clustering.cpp:110-113 compile unchanged against it.
// Results follow the reference's CPU semantics bit for bit (see include/dcb200.h).
//
// Inside the reference tree: add this repository's include/ to the include path, include <dcb200/density_cuda.hpp>
// instead of density_clustering_cuda.hpp (the types Pops / Neighborhood then come from the reference's own headers)
// and link libdcb200.so -- cmake/dcb200.cmake does exactly that.
// Stand-alone (this repository's CLI and tests): define DCB200_STANDALONE_TYPES before including.
#pragma once

#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <limits>
#include <map>
#include <set>
#include <string>
#include <tuple>
#include <utility>
#include <vector>

#include "../dcb200.h"

#ifdef DCB200_STANDALONE_TYPES
namespace Clustering {
namespace Tools {
  // reference: src/tools.hpp:64-66
  using Neighbor = std::pair<std::size_t, float>;
  using Neighborhood = std::map<std::size_t, Clustering::Tools::Neighbor>;
}  // namespace Tools
namespace Density {
  // reference: src/density_clustering_common.hpp:39, src/density_clustering.hpp:54
  typedef std::map<float, std::vector<std::size_t>> Pops;
  using FreeEnergy = std::pair<std::size_t, float>;
  using Neighborhood = Clustering::Tools::Neighborhood;
}  // namespace Density
}  // namespace Clustering
#endif

namespace Clustering {
namespace Density {
namespace CUDA {

  using Neighborhood = Clustering::Tools::Neighborhood;

  // reference: check_error, density_clustering_cuda.cu:21-30 -- print and exit, no exceptions, no codes
  inline void
  check_error(std::string msg="") {
    const char* err = dcb200_last_error();
    if (err && err[0] != '\0') {
      std::cerr << "CUDA error: " << msg << "\n" << err << std::endl;
      exit(EXIT_FAILURE);
    }
  }

  inline void
  dcb200_or_die(int rc, const char* what) {
    if (rc != 0) {
      std::cerr << "error: " << what << ": " << dcb200_last_error() << std::endl;
      exit(EXIT_FAILURE);
    }
  }

  // reference: density_clustering_cuda.cu:32-43
  inline int
  get_num_gpus() {
    int n_gpus = 0;
    int rc = dcb200_device_count(&n_gpus);
    if (rc != 0 || n_gpus == 0) {
      std::cerr << "error: no CUDA-compatible GPUs found" << std::endl;
      exit(EXIT_FAILURE);
    }
    return n_gpus;
  }

  // reference: density_clustering_cuda.cu:139-182 (CPU semantics: density_clustering.cpp:126-195)
  inline Pops
  calculate_populations(const float* coords
                      , const std::size_t n_rows
                      , const std::size_t n_cols
                      , std::vector<float> radii) {
    std::vector<std::uint32_t> buf(radii.size() * n_rows);
    dcb200_or_die(dcb200_populations(coords, n_rows, n_cols, radii.data(), radii.size(), buf.data()),
                  "calculate_populations");
    Pops pops;
    for (std::size_t r=0; r < radii.size(); ++r) {
      std::vector<std::size_t>& p = pops[radii[r]];
      p.assign(buf.begin() + r*n_rows, buf.begin() + (r+1)*n_rows);
    }
    return pops;
  }

  // single-radius convenience, mirrors density_clustering.cpp:107-124
  inline std::vector<std::size_t>
  calculate_populations(const float* coords
                      , const std::size_t n_rows
                      , const std::size_t n_cols
                      , const float radius) {
    std::vector<float> radii = {radius};
    return calculate_populations(coords, n_rows, n_cols, radii)[radius];
  }

  // reference: calculate_free_energies, density_clustering.cpp:197-212 (host function in the reference;
  // offered here as well so that a whole density run can stay on the device path)
  inline std::vector<float>
  calculate_free_energies(const std::vector<std::size_t>& pops) {
    std::vector<std::uint32_t> p(pops.begin(), pops.end());
    std::vector<float> fe(pops.size());
    dcb200_or_die(dcb200_free_energies(p.data(), p.size(), fe.data()), "calculate_free_energies");
    return fe;
  }

  // reference: density_clustering_cuda.cu:286-328 (CPU semantics: density_clustering.cpp:230-288):
  // (nearest neighbour, nearest neighbour with lower free energy); values are SQUARED distances;
  // "none" = (n_rows+1, FLT_MAX).  The maps are built once with hinted insertion (frames arrive in order).
  inline std::tuple<Neighborhood, Neighborhood>
  nearest_neighbors(const float* coords,
                    const std::size_t n_rows,
                    const std::size_t n_cols,
                    const std::vector<float>& free_energy) {
    if (free_energy.size() != n_rows) {
      std::cerr << "error: nearest_neighbors: free energies and coordinates differ in length" << std::endl;
      exit(EXIT_FAILURE);
    }
    std::vector<std::uint32_t> ni(n_rows), hi(n_rows);
    std::vector<float> nd(n_rows), hd(n_rows);
    dcb200_or_die(dcb200_nearest_neighbors(coords, n_rows, n_cols, free_energy.data(),
                                           ni.data(), nd.data(), hi.data(), hd.data()),
                  "nearest_neighbors");
    std::tuple<Neighborhood, Neighborhood> result;
    Neighborhood& nh = std::get<0>(result);
    Neighborhood& nh_hd = std::get<1>(result);
    for (std::size_t i=0; i < n_rows; ++i) {
      nh.emplace_hint(nh.end(), i, Clustering::Tools::Neighbor(ni[i], nd[i]));
      nh_hd.emplace_hint(nh_hd.end(), i, Clustering::Tools::Neighbor(hi[i], hd[i]));
    }
    return result;
  }

  namespace detail {
    // the screening run a sequence of screening() calls shares (see screening below)
    struct ScreeningContinuation {
      dcb200_screening_run* run;
      std::size_t n_rows, n_cols;
      std::vector<float> fe, nn_d2, coords;
      std::vector<std::uint32_t> last_labels;
      float last_threshold;
      ScreeningContinuation() : run(NULL), n_rows(0), n_cols(0), last_threshold(0.f) {}
      void reset() {
        if (run) dcb200_screening_end(run);
        run = NULL;
        std::vector<float>().swap(fe);
        std::vector<float>().swap(nn_d2);
        std::vector<float>().swap(coords);
        std::vector<std::uint32_t>().swap(last_labels);
      }
      ~ScreeningContinuation() { reset(); }
    };
  }  // namespace detail

  // reference: density_clustering_cuda.cu:396-594 (CPU semantics: density_clustering_common.cpp:37-134)
  inline std::vector<std::size_t>
  screening(const std::vector<float>& free_energy
          , const Neighborhood& nh
          , const float free_energy_threshold
          , const float* coords
          , const std::size_t n_rows
          , const std::size_t n_cols
          , const std::vector<std::size_t> initial_clusters) {
    if (free_energy.size() != n_rows || nh.size() != n_rows) {
      std::cerr << "error: screening: free energies / neighborhood and coordinates differ in length" << std::endl;
      exit(EXIT_FAILURE);
    }
    std::vector<float> nn_d2(n_rows);
    {
      std::size_t i = 0;
      for (Neighborhood::const_iterator it = nh.begin(); it != nh.end(); ++it, ++i) nn_d2[i] = it->second.second;
    }
    // like the reference, initial clusters only count when they cover all frames (density_clustering.cpp:390-395)
    std::vector<std::uint32_t> init;
    if (initial_clusters.size() == n_rows) init.assign(initial_clusters.begin(), initial_clusters.end());
    std::vector<std::uint32_t> labels(n_rows);
    // The reference's driver calls this function once per threshold with its previous result (density_clustering.cpp:806-816).
    // A call that CONTINUES the previous one -- same free energies, neighbourhood and coordinates (compared value by value),
    // initial_clusters equal to the labels returned last time, threshold not lower -- is served by the screening run kept
    // from that call: no second sort of the free energies, no second upload of the coordinates.  A call without initial
    // clusters starts a new run; anything else goes through the stateless dcb200_screening, which accepts any labelling.
    static detail::ScreeningContinuation k;
    const std::size_t coord_bytes = n_rows * n_cols * sizeof(float);
    const bool continues = k.run != NULL && !init.empty() && k.n_rows == n_rows && k.n_cols == n_cols &&
                           !(free_energy_threshold < k.last_threshold) && init == k.last_labels && free_energy == k.fe &&
                           nn_d2 == k.nn_d2 && std::memcmp(coords, k.coords.data(), coord_bytes) == 0;
    if (!continues) {
      k.reset();
      if (init.empty()) {
        dcb200_or_die(dcb200_screening_begin(free_energy.data(), nn_d2.data(), coords, n_rows, n_cols, &k.run), "screening");
        k.n_rows = n_rows;
        k.n_cols = n_cols;
        k.fe = free_energy;
        k.nn_d2 = nn_d2;
        k.coords.assign(coords, coords + n_rows * n_cols);
      }
    }
    if (k.run != NULL) {
      dcb200_or_die(dcb200_screening_next(k.run, free_energy_threshold, labels.data()), "screening");
      k.last_labels = labels;
      k.last_threshold = free_energy_threshold;
    } else {
      dcb200_or_die(dcb200_screening(free_energy.data(), nn_d2.data(), free_energy_threshold, coords, n_rows, n_cols,
                                     init.empty() ? NULL : init.data(), labels.data()),
                    "screening");
    }
    return std::vector<std::size_t>(labels.begin(), labels.end());
  }

}}} // end Clustering::Density::CUDA
