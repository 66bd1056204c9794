// TEST INFRASTRUCTURE: a stand-in for Boost.Program_options (absent from this image) with just enough of
// variables_map for the reference's Clustering::Density::main (src/density_clustering.cpp:559-825) to compile:
// args.count("key") and args["key"].as<T>().  Nothing here runs; tests/test_reference_callsites.py compiles the
// reference's translation unit with -DUSE_CUDA against include/dcb200 to prove that its call sites bind unchanged.
#pragma once
#include <cstddef>
#include <map>
#include <string>
#include <vector>
namespace boost { namespace program_options {
class variable_value {
 public:
  template <class T> const T& as() const { static T v; return v; }
};
class variables_map {
 public:
  // opaque to the optimiser, so that no branch of Density::main (and none of its CUDA call sites) is folded away
  __attribute__((noinline)) std::size_t count(const std::string& k) const { static volatile std::size_t n = 0; return n + (k.size() & 0); }
  const variable_value& operator[](const std::string&) const { static variable_value v; return v; }
};
}}  // namespace boost::program_options
