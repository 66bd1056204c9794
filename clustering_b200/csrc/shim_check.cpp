// Test driver of the C++ boundary (include/dcb200/density_cuda.hpp): runs the reference-signature entry points on a
// raw float32 coordinate file and dumps the results as raw arrays for tests/test_gpu_shim.py.
//   shim_check <coords.f32> <n_rows> <n_cols> <out_prefix> <threshold_step> <radius> [radius ...]
#define DCB200_STANDALONE_TYPES
#include "../../include/dcb200/density_cuda.hpp"

#include <cstdio>
#include <fstream>

template <class T>
static void dump(const std::string& name, const std::vector<T>& v) {
  std::ofstream f(name.c_str(), std::ios::binary);
  f.write(reinterpret_cast<const char*>(v.data()), v.size() * sizeof(T));
}

int main(int argc, char** argv) {
  if (argc < 7) {
    std::cerr << "usage: shim_check coords.f32 n_rows n_cols out_prefix threshold_step radius..." << std::endl;
    return 2;
  }
  const std::size_t n = std::strtoull(argv[2], NULL, 10), d = std::strtoull(argv[3], NULL, 10);
  const std::string prefix = argv[4];
  const float step = std::strtof(argv[5], NULL);
  std::vector<float> radii;
  for (int a = 6; a < argc; ++a) radii.push_back(std::strtof(argv[a], NULL));
  std::vector<float> coords(n * d);
  {
    std::ifstream f(argv[1], std::ios::binary);
    f.read(reinterpret_cast<char*>(coords.data()), coords.size() * sizeof(float));
    if (!f) { std::cerr << "cannot read " << argv[1] << std::endl; return 2; }
  }
  namespace CU = Clustering::Density::CUDA;
  std::cout << "gpus " << CU::get_num_gpus() << std::endl;
  Clustering::Density::Pops pops = CU::calculate_populations(coords.data(), n, d, radii);
  std::vector<std::uint32_t> pops_out;
  for (std::size_t r = 0; r < radii.size(); ++r)
    for (std::size_t i = 0; i < n; ++i) pops_out.push_back((std::uint32_t) pops[radii[r]][i]);
  dump(prefix + ".pops.u32", pops_out);
  std::vector<float> fe = CU::calculate_free_energies(pops[radii[0]]);
  dump(prefix + ".fe.f32", fe);
  std::tuple<CU::Neighborhood, CU::Neighborhood> nh = CU::nearest_neighbors(coords.data(), n, d, fe);
  std::vector<std::uint32_t> idx;
  std::vector<float> d2;
  for (int w = 0; w < 2; ++w) {
    const CU::Neighborhood& m = w == 0 ? std::get<0>(nh) : std::get<1>(nh);
    for (std::size_t i = 0; i < n; ++i) { idx.push_back((std::uint32_t) m.at(i).first); d2.push_back(m.at(i).second); }
  }
  dump(prefix + ".nn.u32", idx);
  dump(prefix + ".nnd.f32", d2);
  // screening thresholds accumulated in float like the reference's driver loop (density_clustering.cpp:804-806)
  float t_to = 0.f;
  for (std::size_t i = 0; i < n; ++i) t_to = fe[i] > t_to ? fe[i] : t_to;
  std::vector<std::uint32_t> labels;
  std::vector<float> thresholds;
  std::vector<std::size_t> clustering;
  for (float t = step; (t < t_to - step / 10 + step) && !(t_to + step / 10 + step < t); t += step) {
    clustering = CU::screening(fe, std::get<0>(nh), t, coords.data(), n, d, clustering);
    thresholds.push_back(t);
    for (std::size_t i = 0; i < n; ++i) labels.push_back((std::uint32_t) clustering[i]);
  }
  dump(prefix + ".thr.f32", thresholds);
  dump(prefix + ".lab.u32", labels);
  // ARBITRARY initial clusters (density_clustering.cpp:394-427 accepts any labelling): the labels of the middle threshold
  // with every third named frame unassigned again and all names shifted, screened at the last threshold; this call does
  // not continue the loop above, so it must not be served from its state
  if (thresholds.size() >= 2) {
    const std::size_t mid = thresholds.size() / 2;
    std::vector<std::size_t> init(n);
    std::vector<std::uint32_t> init_out(n), arb(n);
    std::size_t named = 0;
    for (std::size_t i = 0; i < n; ++i) {
      const std::uint32_t l = labels[mid * n + i];
      init[i] = l == 0 ? 0 : (++named % 3 == 0 ? 0 : l + 5);
      init_out[i] = (std::uint32_t) init[i];
    }
    const std::vector<std::size_t> res = CU::screening(fe, std::get<0>(nh), thresholds.back(), coords.data(), n, d, init);
    for (std::size_t i = 0; i < n; ++i) arb[i] = (std::uint32_t) res[i];
    dump(prefix + ".arbinit.u32", init_out);
    dump(prefix + ".arb.u32", arb);
  }
  std::cout << "ok " << thresholds.size() << " thresholds" << std::endl;
  return 0;
}
