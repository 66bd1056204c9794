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

// Screening = reference screening() (density_clustering_common.cpp:37-134) with prepare_initial_clustering (:382-435),
// high_density_neighborhood (:292-332), lump_initial_clusters (:506-555) and normalized_cluster_names (:437-456) of
// density_clustering.cpp.
//
// Closed form.  With S = the free-energy-sorted frames, M = #{fe <= threshold}, cut = (float)(4 sigma2), "visited" = the sorted
// positions below M that carry a name in `initial` (the reference never expands those, :98-99, :417-427):
//   * the clusters are the connected components of { positions < M } under
//       - same initial name  (frames that share a name are one cluster from the start), and
//       - d2(a, b) < cut for an UNVISITED a and any b < M  (the neighbourhoods the reference scans);
//   * a component keeps the smallest initial name among its members; components without named members get fresh names
//     max(initial)+1, +2, ... in the order of their first sorted member (:527-535);
//   * the names in use below the threshold are renumbered 1..K in ascending order (:437-456); frames above the threshold
//     keep their (renumbered) initial name if it is still in use and become 0 otherwise.
// The pair scan runs on the GPU(s) over the frames in the order [visited..., unvisited...] (both ascending): every edge has
// an unvisited row and a column before it, which is what dcb200_screen_step scans.  When `initial` is the labelling of a
// lower threshold of the same data (the reference's driver loop, density_clustering.cpp:806-816) the visited frames are a
// prefix of S and this order is S itself.
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

size_t frames_below(const float* fe, const std::vector<uint32_t>& order, float threshold) {
  // first_frame_above_threshold = upper_bound over the sorted free energies (:403-410)
  size_t lo = 0, hi = order.size();
  while (lo < hi) {
    const size_t mid = (lo + hi) / 2;
    if (threshold < fe[order[mid]]) hi = mid; else lo = mid + 1;
  }
  return lo;
}

void gather_rows(const float* coords, size_t n_cols, const uint32_t* frames, size_t m, float* out) {
#pragma omp parallel for schedule(static)
  for (long long p = 0; p < (long long) m; ++p)
    memcpy(out + (size_t) p * n_cols, coords + (size_t) frames[p] * n_cols, n_cols * sizeof(float));
}
}  // namespace

extern "C" int dcb200_screening(const float* fe, const float* nn_d2, float threshold, const float* coords, size_t n_rows,
                                size_t n_cols, const uint32_t* initial, uint32_t* labels) {
  if (!fe || !nn_d2 || !coords || !labels) return dcb200_internal_fail("dcb200_screening: null argument");
  if (n_rows == 0) return 0;
  Lap lap;
  std::vector<uint32_t> order(n_rows);
  int rc = dcb200_sorted_free_energies(fe, n_rows, order.data());
  if (rc) return rc;
  lap("order");
  const size_t M = frames_below(fe, order, threshold);
  double sigma2 = 0.0;
  dcb200_sigma2(nn_d2, n_rows, &sigma2);
  const float max_dist2 = (float) (4 * sigma2);

  // scan order: visited positions first, then the unvisited ones (both ascending)
  std::vector<uint32_t> pos;                     // scan position -> sorted position
  pos.reserve(M);
  uint32_t max_name = 0;
  size_t m_vis = 0;
  if (initial) {
    for (size_t i = 0; i < n_rows; ++i) max_name = std::max(max_name, initial[i]);
    for (size_t p = 0; p < M; ++p)
      if (initial[order[p]] != 0) pos.push_back((uint32_t) p);
    m_vis = pos.size();
    for (size_t p = 0; p < M; ++p)
      if (initial[order[p]] == 0) pos.push_back((uint32_t) p);
  } else {
    for (size_t p = 0; p < M; ++p) pos.push_back((uint32_t) p);
  }
  std::vector<uint32_t> comp(M);
  {
    // representative of an initial cluster = its first member in scan order
    std::vector<uint32_t> first((size_t) max_name + 1, 0xffffffffu);
    for (size_t q = 0; q < m_vis; ++q) {
      uint32_t& f = first[initial[order[pos[q]]]];
      if (f == 0xffffffffu) f = (uint32_t) q;
      comp[q] = f;
    }
  }
  if (M > m_vis) {
    std::vector<uint32_t> frames(M);
    for (size_t q = 0; q < M; ++q) frames[q] = order[pos[q]];
    std::vector<float> sorted((size_t) M * n_cols);
    gather_rows(coords, n_cols, frames.data(), M, sorted.data());
    lap("prepare+gather");
    rc = dcb200_screening_step(sorted.data(), n_cols, m_vis, M, max_dist2, comp.data());
    if (rc) return rc;
    lap("pair scan (GPU)");
  }
  // name of a component: smallest initial name among its members, else a fresh name ordered by its first sorted member
  // (its unvisited members are in ascending sorted order in the scan, so that member is the representative)
  const uint32_t none = 0xffffffffu;
  std::vector<uint32_t> name_of_rep(M, none);
  for (size_t q = 0; q < m_vis; ++q) {
    uint32_t& nm = name_of_rep[comp[q]];
    nm = std::min(nm, initial[order[pos[q]]]);
  }
  std::vector<std::pair<uint64_t, uint32_t>> keys;          // (old name, representative), ascending = the reference's renumbering
  for (size_t q = 0; q < M; ++q)
    if (comp[q] == q) {
      const uint64_t k = name_of_rep[q] != none ? (uint64_t) name_of_rep[q] : ((uint64_t) 1 << 32) + pos[q];
      keys.push_back(std::make_pair(k, (uint32_t) q));
    }
  std::sort(keys.begin(), keys.end());
  std::vector<uint32_t> final_name(M, 0);
  for (size_t k = 0; k < keys.size(); ++k) final_name[keys[k].second] = (uint32_t) (k + 1);
  if (initial) {
    // frames above the threshold: their initial name, renumbered, if a cluster below the threshold still carries it
    std::vector<uint32_t> renamed((size_t) max_name + 1, 0);
    for (size_t k = 0; k < keys.size(); ++k)
      if (keys[k].first <= max_name) renamed[keys[k].first] = (uint32_t) (k + 1);
    for (size_t i = 0; i < n_rows; ++i) labels[i] = renamed[initial[i]];
  } else {
    memset(labels, 0, n_rows * sizeof(uint32_t));
  }
  for (size_t q = 0; q < M; ++q) labels[order[pos[q]]] = final_name[comp[q]];
  lap("names");
  return 0;
}

// ---- a whole screening run: the thresholds of one `clustering density -T` call --------------------------------------
// The reference's driver (density_clustering.cpp:806-816) calls screening() once per threshold with the previous labels;
// every call sorts the free energies again (:401) and, on its CUDA path, uploads all coordinates again (cuda.cu:505-571).
// A run does both ONCE: the free-energy order is kept, the sorted coordinates and the union-find forest stay on the
// device(s) (dcb200_screen_*).  Labels are identical to the call-per-threshold form: with the previous labels as initial
// clusters, named clusters keep their order and fresh ones follow, i.e. clusters are numbered by their first sorted member.
struct dcb200_screening_run {
  size_t n = 0, d = 0;
  std::vector<float> fe;
  std::vector<uint32_t> order;
  float max_dist2 = 0.f;
  dcb200_screen* scan = nullptr;
  size_t m_done = 0;
  float last_threshold = 0.f;
  bool any = false;
};

extern "C" int dcb200_screening_begin(const float* fe, const float* nn_d2, const float* coords, size_t n_rows, size_t n_cols,
                                      dcb200_screening_run** out) {
  if (!fe || !nn_d2 || !coords || !out) return dcb200_internal_fail("dcb200_screening_begin: null argument");
  *out = nullptr;
  if (n_rows == 0 || n_cols == 0) return dcb200_internal_fail("dcb200_screening_begin: empty input");
  Lap lap;
  dcb200_screening_run* r = new dcb200_screening_run();
  r->n = n_rows;
  r->d = n_cols;
  r->fe.assign(fe, fe + n_rows);
  r->order.resize(n_rows);
  int rc = dcb200_sorted_free_energies(fe, n_rows, r->order.data());
  double sigma2 = 0.0;
  if (!rc) rc = dcb200_sigma2(nn_d2, n_rows, &sigma2);
  r->max_dist2 = (float) (4 * sigma2);
  lap("order");
  if (!rc) {
    std::vector<float> sorted(n_rows * n_cols);
    gather_rows(coords, n_cols, r->order.data(), n_rows, sorted.data());
    lap("gather");
    rc = dcb200_screen_begin(sorted.data(), n_rows, n_cols, &r->scan);
    if (!rc) rc = dcb200_screen_set_order(r->scan, r->order.data());
    lap("upload+layout");
  }
  if (rc) {
    delete r;
    return rc;
  }
  *out = r;
  return 0;
}

extern "C" int dcb200_screening_next(dcb200_screening_run* r, float threshold, uint32_t* labels) {
  if (!r || !labels) return dcb200_internal_fail("dcb200_screening_next: null argument");
  if (r->any && threshold < r->last_threshold) return dcb200_internal_fail("dcb200_screening_next: thresholds must not decrease within a run");
  Lap lap;
  const size_t M = frames_below(r->fe.data(), r->order, threshold);
  // the pair work, the union-find and the naming (clusters numbered 1..K by ascending representative = first sorted
  // member) all happen on the device; one download brings the labels in frame order
  const int rc = dcb200_screen_labels(r->scan, M, r->max_dist2, labels, nullptr);
  if (rc) return rc;
  lap("threshold (GPU)");
  r->m_done = M;
  r->last_threshold = threshold;
  r->any = true;
  return 0;
}

extern "C" int dcb200_screening_end(dcb200_screening_run* r) {
  if (!r) return 0;
  dcb200_screen_end(r->scan);
  delete r;
  return 0;
}
