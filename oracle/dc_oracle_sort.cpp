// TEST INFRASTRUCTURE ONLY (see dc_oracle.c).  The one piece of the oracle that cannot be plain C:
// sorted_free_energies (density_clustering.cpp:214-228) orders (frame, fe) pairs with libstdc++'s
// *unstable* std::sort and the comparator `a.second < b.second`.  Free-energy ties are the norm
// (integer populations), and the tie order decides the cluster numbering of the screening step, so
// the oracle makes the very same library call on the very same element type.
#include <algorithm>
#include <cstdint>
#include <utility>
#include <vector>

extern "C" void dco_sorted_free_energies(const float* fe, uint64_t n, uint64_t* order_out) {
  typedef std::pair<std::size_t, float> FreeEnergy;
  std::vector<FreeEnergy> v;
  for (std::size_t i = 0; i < n; ++i) v.push_back(FreeEnergy(i, fe[i]));
  std::sort(v.begin(), v.end(), [](const FreeEnergy& a, const FreeEnergy& b) -> bool { return a.second < b.second; });
  for (std::size_t k = 0; k < n; ++k) order_out[k] = v[k].first;
}
