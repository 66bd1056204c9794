// Host-side pieces of the density path that stay on the CPU in the reference as well (they are
// O(N log N) bookkeeping around the pair scans): the free-energy ordering, sigma^2, and the driver of
// one screening threshold (which frames are new, which cluster names survive).  The pair scan itself
// is dcb200_screening_step (CUDA).  All citations are file:line under the reference's src/.
#include "../../include/dcb200.h"

#include <math.h>
#include <string.h>

#include <algorithm>
#include <mutex>
#include <string>
#include <utility>
#include <vector>

extern "C" int dcb200_internal_fail(const char* msg);

// sorted_free_energies (density_clustering.cpp:214-228): ascending free energy.  The reference uses
// libstdc++'s unstable std::sort on (frame, fe) pairs and free-energy ties are the norm, so the very
// same library call on the same element type is made here -- the tie order decides cluster numbers.
extern "C" int dcb200_sorted_free_energies(const float* fe, size_t n, uint32_t* order) {
  if (!fe || !order) return dcb200_internal_fail("dcb200_sorted_free_energies: null argument");
  typedef std::pair<std::size_t, float> FeEntry;
  std::vector<FeEntry> v(n);
  for (std::size_t i = 0; i < n; ++i) v[i] = FeEntry(i, fe[i]);
  std::sort(v.begin(), v.end(), [](const FeEntry& a, const FeEntry& b) -> bool { return a.second < b.second; });
  for (std::size_t k = 0; k < n; ++k) order[k] = (uint32_t) v[k].first;
  return 0;
}

// compute_sigma2 (density_clustering.cpp:334-343): mean squared nearest-neighbour distance,
// accumulated in double in frame order.
extern "C" int dcb200_sigma2(const float* nn_d2, size_t n, double* sigma2) {
  if (!nn_d2 || !sigma2) return dcb200_internal_fail("dcb200_sigma2: null argument");
  double s = 0.0;
  for (size_t i = 0; i < n; ++i) s += nn_d2[i];
  *sigma2 = s / (double) n;
  return 0;
}

// assign_low_density_frames (density_clustering.cpp:345-360): in ascending free-energy order every frame without a
// state takes the state of its nearest neighbour with lower free energy (which, coming earlier in that order, is
// settled already).  A frame whose neighbour entry is the "none" sentinel keeps state 0 (the reference would index
// out of bounds there).
extern "C" int dcb200_assign_low_density_frames(const uint32_t* initial, const uint32_t* hd_idx, const float* fe, size_t n,
                                                uint32_t* states) {
  if (!initial || !hd_idx || !fe || !states) return dcb200_internal_fail("dcb200_assign_low_density_frames: null argument");
  std::vector<uint32_t> order(n);
  const int rc = dcb200_sorted_free_energies(fe, n, order.data());
  if (rc) return rc;
  memmove(states, initial, n * sizeof(uint32_t));
  for (size_t k = 0; k < n; ++k) {
    const size_t id = order[k];
    if (states[id] == 0 && hd_idx[id] < n) states[id] = states[hd_idx[id]];
  }
  return 0;
}

// sorted_cluster_names (density_clustering.cpp:458-493): states renamed 1..K by DEcreasing population.  Populations
// tie often; the order of equal counts is whatever libstdc++'s unstable std::sort makes of the (state, count) vector,
// so the same call on the same element type is made here.
extern "C" int dcb200_sorted_cluster_names(const uint32_t* states, size_t n, uint32_t* renamed) {
  if (!states || !renamed) return dcb200_internal_fail("dcb200_sorted_cluster_names: null argument");
  uint32_t mx = 0;
  for (size_t i = 0; i < n; ++i) mx = std::max(mx, states[i]);
  std::vector<std::size_t> count((size_t) mx + 1, 0);
  for (size_t i = 0; i < n; ++i) ++count[states[i]];
  typedef std::pair<std::size_t, std::size_t> Entry;                 // (state, population), ascending state like the map
  std::vector<Entry> v;
  for (size_t s = 0; s <= mx; ++s)
    if (count[s]) v.push_back(Entry(s, count[s]));
  std::sort(v.begin(), v.end(), [](const Entry& a, const Entry& b) -> bool { return a.second < b.second; });
  std::vector<uint32_t> name((size_t) mx + 1, 0);
  for (size_t i = 0; i < v.size(); ++i) name[v[i].first] = (uint32_t) (v.size() - i);
  for (size_t i = 0; i < n; ++i) renamed[i] = name[states[i]];
  return 0;
}

// One screening threshold = reference screening() (density_clustering_common.cpp:37-134) with
// prepare_initial_clustering (:382-435), high_density_neighborhood (:292-332), lump_initial_clusters
// (:506-555) and normalized_cluster_names (:437-456) of density_clustering.cpp.
//
// Closed form used here: the clusters are the connected components of
//   { sorted positions a, b < M : d2(a, b) < (float)(4 sigma2) },   M = #{fe <= threshold},
// grown from the clusters of `initial` (frames that already carry a name are not expanded again, :98-99);
// a component keeps the smallest initial name among its members, components without named members
// are named max(initial)+1, +2, ... in the order of their first sorted member (:527-535), and finally
// the names are renumbered 1..K in ascending order (:437-456).
#include <chrono>
#include <stdio.h>
#include <stdlib.h>
namespace {
struct Lap {          // DCB200_TRACE=1: wall-clock phases on stderr (diagnostics only)
  bool on;
  std::chrono::steady_clock::time_point last;
  Lap() : on(getenv("DCB200_TRACE") != nullptr), last(std::chrono::steady_clock::now()) {}
  void operator()(const char* what) {
    if (!on) return;
    const auto now = std::chrono::steady_clock::now();
    fprintf(stderr, "[dcb200] screening: %-16s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(now - last).count());
    last = now;
  }
};
}  // namespace

namespace {
// The reference's driver calls screening() once per threshold with the same free energies (density_clustering.cpp:806-816)
// and sorts them again every time (:401); at 5M frames that sort is 0.3 s per threshold, far more than the pair scan.
// The order is therefore kept between calls, keyed by the CONTENT of the array (a 64-bit hash, ~2 ms): a changed input
// simply misses.
struct OrderCache {
  std::mutex mu;
  size_t n = 0;
  uint64_t hash = 0;
  std::vector<uint32_t> order;
};
OrderCache g_order_cache;

uint64_t hash_floats(const float* v, size_t n) {
  // 4 independent multiply-xor lanes over 64-bit words, folded at the end
  uint64_t h[4] = {0x9e3779b97f4a7c15ull, 0xc2b2ae3d27d4eb4full, 0x165667b19e3779f9ull, 0x27d4eb2f165667c5ull};
  const size_t words = n / 2;
  const uint64_t* w = reinterpret_cast<const uint64_t*>(v);
  size_t i = 0;
  for (; i + 4 <= words; i += 4)
    for (int q = 0; q < 4; ++q) {
      uint64_t x;
      memcpy(&x, w + i + q, 8);
      h[q] = (h[q] ^ x) * 0x100000001b3ull;
      h[q] ^= h[q] >> 29;
    }
  uint64_t r = h[0] ^ (h[1] * 3) ^ (h[2] * 5) ^ (h[3] * 7) ^ (uint64_t) n;
  for (size_t k = i * 2; k < n; ++k) {
    uint32_t x;
    memcpy(&x, v + k, 4);
    r = (r ^ x) * 0x100000001b3ull;
  }
  return r;
}

int cached_order(const float* fe, size_t n, std::vector<uint32_t>& out) {
  const uint64_t h = hash_floats(fe, n);
  std::lock_guard<std::mutex> lock(g_order_cache.mu);
  if (g_order_cache.n != n || g_order_cache.hash != h || g_order_cache.order.size() != n) {
    g_order_cache.order.resize(n);
    const int rc = dcb200_sorted_free_energies(fe, n, g_order_cache.order.data());
    if (rc) {
      g_order_cache.n = 0;
      return rc;
    }
    g_order_cache.n = n;
    g_order_cache.hash = h;
  }
  out = g_order_cache.order;
  return 0;
}
}  // namespace

extern "C" int dcb200_screening(const float* fe, const float* nn_d2, float threshold, const float* coords, size_t n_rows,
                                size_t n_cols, const uint32_t* initial, uint32_t* labels) {
  if (!fe || !nn_d2 || !coords || !labels) return dcb200_internal_fail("dcb200_screening: null argument");
  if (n_rows == 0) return 0;
  Lap lap;
  std::vector<uint32_t> order;
  int rc = cached_order(fe, n_rows, order);
  if (rc) return rc;
  lap("order");
  // first_frame_above_threshold = upper_bound over the sorted free energies (:403-410)
  size_t lo = 0, hi = n_rows;
  while (lo < hi) {
    const size_t mid = (lo + hi) / 2;
    if (threshold < fe[order[mid]]) hi = mid; else lo = mid + 1;
  }
  const size_t M = lo;
  double sigma2 = 0.0;
  dcb200_sigma2(nn_d2, n_rows, &sigma2);
  const float max_dist2 = (float) (4 * sigma2);

  // frames that already carry a name are "visited" (:417-427); they must be the first m_prev sorted frames
  size_t m_prev = 0;
  uint32_t max_name = 0;
  if (initial) {
    for (size_t i = 0; i < n_rows; ++i) max_name = std::max(max_name, initial[i]);
    while (m_prev < M && initial[order[m_prev]] != 0) ++m_prev;
    for (size_t p = m_prev; p < M; ++p)
      if (initial[order[p]] != 0)
        return dcb200_internal_fail(
            "dcb200_screening: initial clusters are not a prefix of the free-energy order "
            "(they must come from a screening at a lower threshold of the same data)");
  }
  std::vector<uint32_t> comp(M);
  {
    // representative of an initial cluster = its first sorted member
    std::vector<uint32_t> first(max_name + 1, 0xffffffffu);
    for (size_t p = 0; p < m_prev; ++p) {
      uint32_t& f = first[initial[order[p]]];
      if (f == 0xffffffffu) f = (uint32_t) p;
      comp[p] = f;
    }
  }
  if (M > m_prev) {
    std::vector<float> sorted((size_t) M * n_cols);
#pragma omp parallel for schedule(static)
    for (long long p = 0; p < (long long) M; ++p)
      memcpy(&sorted[(size_t) p * n_cols], coords + (size_t) order[p] * n_cols, n_cols * sizeof(float));
    lap("prepare+gather");
    rc = dcb200_screening_step(sorted.data(), n_cols, m_prev, M, max_dist2, comp.data());
    if (rc) return rc;
    lap("pair scan (GPU)");
  }
  // name of a component: smallest initial name among its members, else a fresh name by first member
  const uint32_t none = 0xffffffffu;
  std::vector<uint32_t> name_of_rep(M, none);
  for (size_t p = 0; p < m_prev; ++p) {
    uint32_t& nm = name_of_rep[comp[p]];
    nm = std::min(nm, initial[order[p]]);
  }
  // ascending order of names: named components by name; fresh ones (all larger) by representative
  std::vector<std::pair<uint64_t, uint32_t>> keys;
  for (size_t p = 0; p < M; ++p)
    if (comp[p] == p) {
      const uint64_t k = name_of_rep[p] != none ? (uint64_t) name_of_rep[p] : ((uint64_t) 1 << 32) + p;
      keys.push_back(std::make_pair(k, (uint32_t) p));
    }
  std::sort(keys.begin(), keys.end());
  std::vector<uint32_t> final_name(M, 0);
  for (size_t q = 0; q < keys.size(); ++q) final_name[keys[q].second] = (uint32_t) (q + 1);
  memset(labels, 0, n_rows * sizeof(uint32_t));
  for (size_t p = 0; p < M; ++p) labels[order[p]] = final_name[comp[p]];
  lap("names");
  return 0;
}
