// `clustering density` -- command-line driver of the B200-native density path.
//
// Same flags, argument-combination errors, output file names and file formats as the reference's
// `clustering density` (option table: src/clustering.cpp:152-193; driver: Clustering::Density::main,
// src/density_clustering.cpp:559-825), so that scripts written for the reference run unchanged.  The compute
// calls go to libdcb200.so (include/dcb200.h); nothing here falls back to a CPU implementation of the pair scans.
// Boost.Program_options is not available in this build environment, so the option table is parsed by hand with
// the same rules the reference relies on (long/short forms, --opt=value, -ovalue, multitoken values that stop at
// the next *known* option -- which is what makes "-T -1" work).
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <iomanip>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

#include "../../../include/dcb200.h"
#include "density_io.hpp"

using namespace dcb_cli;

namespace {

const char* VERSION = "v1.3.2";          // file headers name the reference version whose formats are produced

struct OptSpec {
  const char* lname;
  char sname;
  int kind;          // 0 switch, 1 single value, 2 multitoken
  const char* help;
};
const OptSpec OPTS[] = {
    {"help", 'h', 0, "show this help."},
    {"file", 'f', 1, "input (required): phase space coordinates (space separated ASCII)."},
    {"radius", 'r', 1, "parameter: hypersphere radius. If not used, the lumping radius will be used instead."},
    {"threshold-screening", 'T', 2,
     "parameters: screening of free energy landscape. format: FROM STEP TO; e.g.: '-T 0.1 0.1 11.1'.\n"
     "set -T -1 for default values: FROM=0.1, STEP=0.1, TO=MAX_FE.\n"
     "parameters may be given partially, e.g.: -T 0.2 0.4 to start at 0.2 and go to MAX_FE at steps 0.4.\n"
     "for threshold-screening, --output denotes the basename only. output files will have the current threshold "
     "limit appended to the given filename."},
    {"output", 'o', 1, "output (optional): clustering information."},
    {"input", 'i', 1, "input (optional): initial state definition."},
    {"radii", 'R', 2,
     "parameter: list of radii for population/free energy calculations (i.e. compute populations/free energies for "
     "several radii in one go)."},
    {"population", 'p', 1, "output (optional): population per frame (if -R is set: this defines only the basename)."},
    {"free-energy", 'd', 1, "output (optional): free energies per frame (if -R is set: this defines only the basename)."},
    {"free-energy-input", 'D', 1, "input (optional): reuse free energy info."},
    {"nearest-neighbors", 'b', 1, "output (optional): nearest neighbor info."},
    {"nearest-neighbors-input", 'B', 1, "input (optional): reuse nearest neighbor info."},
    {"nthreads", 'n', 1, "number of OpenMP threads. default: 0; i.e. use OMP_NUM_THREADS env-variable."},
    {"verbose", 'v', 0, "verbose mode: print runtime information to STDOUT."},
};
const int N_OPTS = (int) (sizeof(OPTS) / sizeof(OPTS[0]));

struct Args {
  std::vector<std::vector<std::string>> values;     // per option: the tokens given
  std::vector<int> seen;
  Args() : values(N_OPTS), seen(N_OPTS, 0) {}
  int index(const char* lname) const {
    for (int i = 0; i < N_OPTS; ++i)
      if (!strcmp(OPTS[i].lname, lname)) return i;
    return -1;
  }
  bool count(const char* lname) const { return seen[index(lname)] != 0; }
  const std::string& str(const char* lname) const { return values[index(lname)][0]; }
};

struct ParseError {
  std::string what;
};

// which option does a command-line token name?  -1: none (a value)
int match_option(const std::string& tok, std::string* attached, bool* has_attached) {
  *has_attached = false;
  if (tok.size() >= 3 && tok[0] == '-' && tok[1] == '-') {
    std::string name = tok.substr(2);
    const size_t eq = name.find('=');
    if (eq != std::string::npos) {
      *attached = name.substr(eq + 1);
      *has_attached = true;
      name = name.substr(0, eq);
    }
    int exact = -1, prefix = -1, n_prefix = 0;
    for (int i = 0; i < N_OPTS; ++i) {
      if (name == OPTS[i].lname) exact = i;
      if (!strncmp(OPTS[i].lname, name.c_str(), name.size())) { prefix = i; ++n_prefix; }
    }
    if (exact >= 0) return exact;
    if (n_prefix == 1) return prefix;              // unambiguous abbreviation, as Boost allows
    if (n_prefix > 1) throw ParseError{"option '--" + name + "' is ambiguous"};
    return -2;                                     // looks like an option, but is not a known one
  }
  if (tok.size() >= 2 && tok[0] == '-' && tok[1] != '-') {
    for (int i = 0; i < N_OPTS; ++i)
      if (OPTS[i].sname == tok[1]) {
        if (tok.size() > 2) {
          *attached = tok.substr(2);
          *has_attached = true;
        }
        return i;
      }
    return -2;
  }
  return -1;
}

Args parse(int argc, char** argv) {
  Args a;
  for (int k = 2; k < argc; ++k) {
    const std::string tok = argv[k];
    std::string attached;
    bool has_attached = false;
    const int o = match_option(tok, &attached, &has_attached);
    if (o == -1) throw ParseError{"too many positional options have been specified on the command line"};
    if (o == -2) throw ParseError{"unrecognised option '" + tok + "'"};
    const OptSpec& s = OPTS[o];
    const std::string shown = std::string("--") + s.lname;
    if (s.kind == 0) {
      if (has_attached && tok[1] == '-') throw ParseError{"option '" + shown + "' does not take any arguments"};
      a.seen[o] = 1;
      if (has_attached) {                           // "-vh": sticky short switches
        std::string rest = "-" + attached;
        std::vector<char*> sub = {argv[0], argv[1], const_cast<char*>(rest.c_str())};
        Args b = parse(3, sub.data());
        for (int i = 0; i < N_OPTS; ++i)
          if (b.seen[i]) { a.seen[i] = 1; a.values[i].insert(a.values[i].end(), b.values[i].begin(), b.values[i].end()); }
      }
      continue;
    }
    if (s.kind == 1 && a.seen[o]) throw ParseError{"option '" + shown + "' cannot be specified more than once"};
    a.seen[o] = 1;
    if (has_attached) {
      a.values[o].push_back(attached);
    } else {
      if (k + 1 >= argc) throw ParseError{"the required argument for option '" + shown + "' is missing"};
      a.values[o].push_back(argv[++k]);
    }
    if (s.kind == 2) {
      // further tokens belong to this option until one names a KNOWN option ("-1" does not)
      while (k + 1 < argc) {
        std::string dummy;
        bool dummy2;
        int next;
        try {
          next = match_option(argv[k + 1], &dummy, &dummy2);
        } catch (const ParseError&) {
          break;
        }
        if (next >= 0) break;
        a.values[o].push_back(argv[++k]);
      }
    }
  }
  return a;
}

float to_float(const std::string& s, const char* lname) {
  char* end = nullptr;
  const float v = strtof(s.c_str(), &end);
  if (end == s.c_str() || *end != '\0')
    throw ParseError{std::string("the argument ('") + s + "') for option '--" + lname + "' is invalid"};
  return v;
}

std::string usage(const std::string& copyright, const char* mode) {
  std::ostringstream o;
  o << copyright << mode
    << ": \n"
       "perform clustering of MD data based on phase space densities.\n"
       "densities are approximated by counting neighboring frames inside\n"
       "a n-dimensional hypersphere of specified radius.\n"
       "distances are measured with n-dim P2-norm.\n"
       "\n"
       "options:\n";
  for (int i = 0; i < N_OPTS; ++i) {
    std::string left = std::string("  -") + OPTS[i].sname + " [ --" + OPTS[i].lname + " ]";
    if (OPTS[i].kind) left += " arg";
    o << std::left << std::setw(38) << left << " ";
    std::string h = OPTS[i].help;
    size_t pos = 0;
    bool first = true;
    while (pos <= h.size()) {
      const size_t nl = h.find('\n', pos);
      const std::string line = h.substr(pos, nl == std::string::npos ? std::string::npos : nl - pos);
      if (!first) o << std::string(39, ' ');
      o << line << "\n";
      first = false;
      if (nl == std::string::npos) break;
      pos = nl + 1;
    }
  }
  return o.str();
}

[[noreturn]] void die_cuda(const char* what) {
  std::cerr << "CUDA error: " << what << "\n" << dcb200_last_error() << std::endl;     // reference: check_error, cuda.cu:21-30
  exit(EXIT_FAILURE);
}

bool has2digits(float val) {                      // density_clustering.cpp:500-504
  const float val_2digits = (int) (val * 100) / 100.0;
  return val_2digits == val;
}

struct Logger {
  template <class T>
  Logger& operator<<(const T& v) {
    if (verbose) std::cout << v;
    return *this;
  }
  Logger& operator<<(std::ostream& (*m)(std::ostream&) ) {
    if (verbose) std::cout << m;
    return *this;
  }
};

std::vector<uint32_t> populations_single(const Coords& c, float radius) {
  std::vector<uint32_t> pops(c.n_rows);
  if (dcb200_populations(c.data.data(), c.n_rows, c.n_cols, &radius, 1, pops.data())) die_cuda("populations");
  return pops;
}
std::vector<float> free_energies_of(const std::vector<uint32_t>& pops) {
  std::vector<float> fe(pops.size());
  if (dcb200_free_energies(pops.data(), pops.size(), fe.data())) die_cuda("free energies");
  return fe;
}
struct Neighbours {
  std::vector<uint32_t> nn_idx, hd_idx;
  std::vector<float> nn_d2, hd_d2;
};
Neighbours nearest_neighbours(const Coords& c, const std::vector<float>& fe) {
  if (fe.size() != c.n_rows) {
    std::cerr << "error: free energies (" << fe.size() << " values) do not match the coordinates (" << c.n_rows << " frames)." << std::endl;
    exit(EXIT_FAILURE);
  }
  Neighbours nb;
  nb.nn_idx.resize(c.n_rows); nb.hd_idx.resize(c.n_rows); nb.nn_d2.resize(c.n_rows); nb.hd_d2.resize(c.n_rows);
  if (dcb200_nearest_neighbors(c.data.data(), c.n_rows, c.n_cols, fe.data(), nb.nn_idx.data(), nb.nn_d2.data(), nb.hd_idx.data(),
                               nb.hd_d2.data()))
    die_cuda("nearest neighbours");
  return nb;
}
double sigma2_of(const std::vector<float>& nn_d2) {
  double s2 = 0.0;
  dcb200_sigma2(nn_d2.data(), nn_d2.size(), &s2);
  return s2;
}

int density_main(const Args& args, const std::string& header_comment) {
  Logger log;
  const Coords coords = read_coords(args.str("file"));
  CommentsMap comments = default_comments();
  std::vector<float> free_energies;
  Neighbours fused_nb;                 // neighbours computed together with the populations (single-radius runs)
  bool have_fused_nb = false;

  if (args.count("input") && (args.count("free-energy") || args.count("nearest-neighbors"))) {
    std::cerr << "error: for input (-i) -D/-B should be used." << std::endl;
    exit(EXIT_FAILURE);
  }
  log << "~~~ free energy and population" << std::endl;
  if (args.count("free-energy-input")) {
    log << "    re-using free energy: " << args.str("free-energy-input") << std::endl;
    if (args.count("radii") || args.count("radius")) log << "warning: radius (-r/-R) is ignored" << std::endl;
    if (args.count("free-energy") || args.count("population")) log << "warning: -p/-d flags are ignored" << std::endl;
    free_energies = read_single_column_float(args.str("free-energy-input"));
    read_comments(args.str("free-energy-input"), comments);
  } else if (args.count("free-energy") || args.count("population") || args.count("output")) {
    if (args.count("radii")) {
      log << "    calculating free energy and population" << std::endl;
      if (args.count("output")) {
        std::cerr << "error: clustering cannot be done with several radii (-R is set)." << std::endl;
        exit(EXIT_FAILURE);
      }
      if (!(args.count("population") || args.count("free-energy"))) {
        std::cerr << "error: no output defined for populations or free energies.\n"
                  << "       why did you define -R ?" << std::endl;
        exit(EXIT_FAILURE);
      }
      std::vector<float> radii;
      for (const std::string& s : args.values[args.index("radii")]) radii.push_back(to_float(s, "radii"));
      log << "    using radii: ";
      for (float r : radii) log << r << ", ";
      log << "\b\b  " << std::endl;
      log << "    using CUDA" << std::endl;
      // one run on the device(s): the populations of all radii and, when wanted, their free energies (one upload, one layout)
      std::vector<uint32_t> pops(radii.size() * coords.n_rows);
      std::vector<float> fes(args.count("free-energy") ? radii.size() * coords.n_rows : 0);
      if (dcb200_density_run(coords.data.data(), coords.n_rows, coords.n_cols, radii.data(), radii.size(), 0, pops.data(),
                             fes.empty() ? nullptr : fes.data(), nullptr, nullptr, nullptr, nullptr, nullptr))
        die_cuda("populations");
      log << "    storing results" << std::endl;
      // the reference keeps the results in a map keyed by radius: one file per DISTINCT radius, ascending
      std::vector<size_t> order(radii.size());
      for (size_t i = 0; i < order.size(); ++i) order[i] = i;
      std::stable_sort(order.begin(), order.end(), [&](size_t x, size_t y) { return radii[x] < radii[y]; });
      for (size_t q = 0; q < order.size(); ++q) {
        if (q > 0 && radii[order[q]] == radii[order[q - 1]]) continue;
        const size_t r = order[q];
        const uint32_t* p = pops.data() + r * coords.n_rows;
        if (args.count("population"))
          write_pops(stringprintf((args.str("population") + "_%f").c_str(), radii[r]), p, coords.n_rows, header_comment, comments);
        if (args.count("free-energy"))
          write_fes(stringprintf((args.str("free-energy") + "_%f").c_str(), radii[r]), fes.data() + r * coords.n_rows, coords.n_rows,
                    header_comment, comments);
      }
    } else {
      float radius_lump = 1.0;
      if (!args.count("radius")) {
        // no radius given: the clustering radius is the lumping radius sqrt(4 sigma^2) of a first pass with radius 1
        log << "    computing lumping radius" << std::endl;
        // The reference runs populations(r = 1) + free energies + neighbours here (density_clustering.cpp:646-673, with a TODO
        // that only sigma is needed).  sigma^2 is the mean squared NEAREST-neighbour distance, which does not depend on the
        // free energies: one neighbour scan with a constant free energy (no frame has a lower one, so the lower-free-energy
        // search is empty) gives the same bits and saves the most expensive population pass of the run.
        const std::vector<float> flat(coords.n_rows, 0.f);
        const Neighbours nb = nearest_neighbours(coords, flat);
        const double sigma2 = sigma2_of(nb.nn_d2);
        radius_lump = sqrt(4 * sigma2);
        log << "        d_lump=" << radius_lump << std::endl;
        comments["lumping_radius"] = radius_lump;
      }
      log << "    calculating free energy and population" << std::endl;
      const float radius = args.count("radius") ? to_float(args.str("radius"), "radius") : radius_lump;
      log << "    using radius: " << radius << std::endl;
      comments["clustering_radius"] = radius;
      // when the neighbours are computed in this run as well, one density run serves all three stages (one upload, one
      // layout build); the neighbour section below then finds them ready
      const bool fuse_nn = !args.count("nearest-neighbors-input") && (args.count("nearest-neighbors") || args.count("output"));
      std::vector<uint32_t> pops;
      if (fuse_nn) {
        pops.resize(coords.n_rows);
        free_energies.resize(coords.n_rows);
        fused_nb.nn_idx.resize(coords.n_rows); fused_nb.hd_idx.resize(coords.n_rows);
        fused_nb.nn_d2.resize(coords.n_rows); fused_nb.hd_d2.resize(coords.n_rows);
        if (dcb200_density_run(coords.data.data(), coords.n_rows, coords.n_cols, &radius, 1, 0, pops.data(), nullptr, free_energies.data(),
                               fused_nb.nn_idx.data(), fused_nb.nn_d2.data(), fused_nb.hd_idx.data(), fused_nb.hd_d2.data()))
          die_cuda("density run");
        have_fused_nb = true;
      } else {
        pops = populations_single(coords, radius);
      }
      if (args.count("population")) {
        log << "    storing population in: " << args.str("population") << std::endl;
        write_pops(args.str("population"), pops.data(), pops.size(), header_comment, comments);
      }
      if (!fuse_nn) free_energies = free_energies_of(pops);
      if (args.count("free-energy")) {
        log << "    storing free energy in: " << args.str("free-energy") << std::endl;
        write_fes(args.str("free-energy"), free_energies.data(), free_energies.size(), header_comment, comments);
      }
    }
  }
  //// nearest neighbours
  Neighbours nb;
  log << "\n~~~ nearest neighbors" << std::endl;
  if (args.count("nearest-neighbors-input")) {
    log << "    re-using nearest neighbor: " << args.str("nearest-neighbors-input") << std::endl;
    read_neighborhood(args.str("nearest-neighbors-input"), nb.nn_idx, nb.nn_d2, nb.hd_idx, nb.hd_d2);
    read_comments(args.str("nearest-neighbors-input"), comments);
  } else if (args.count("nearest-neighbors") || args.count("output")) {
    if (args.count("radii")) {
      std::cerr << "error: nearest neighbor calculation cannot be done with\n"
                << "       several radii (-R is set)." << std::endl;
      exit(EXIT_FAILURE);
    }
    log << "    calculating nearest neighbors" << std::endl;
    if (have_fused_nb) nb = std::move(fused_nb);
    else nb = nearest_neighbours(coords, free_energies);
    if (comments["lumping_radius"] == 0.) {
      const double sigma2 = sigma2_of(nb.nn_d2);
      const float radius_lump = sqrt(4 * sigma2);
      log << "    lumping radius: " << radius_lump << std::endl;
      comments["lumping_radius"] = radius_lump;
    }
    if (args.count("nearest-neighbors")) {
      log << "    storing nearest neighbors in: " << args.str("nearest-neighbors") << std::endl;
      write_neighborhood(args.str("nearest-neighbors"), nb.nn_idx.data(), nb.nn_d2.data(), nb.hd_idx.data(), nb.hd_d2.data(),
                         coords.n_rows, header_comment, comments);
    }
  }
  //// clustering
  if (args.count("output")) {
    if (args.count("radii")) {
      std::cerr << "error: output needs to depend on single radius\n"
                << "       but several radii (-R) are set." << std::endl;
      exit(EXIT_FAILURE);
    }
    const std::string output_file = args.str("output");
    if (args.count("input")) {
      log << "~~~ generating microstates" << std::endl;
      if (args.count("threshold-screening")) log << "warning: screening (-T) is ignored" << std::endl;
      log << "    reading initial states: " << args.str("input") << std::endl;
      const std::vector<std::size_t> initial = read_single_column_size(args.str("input"));
      read_comments(args.str("input"), comments);
      if (initial.size() != coords.n_rows || free_energies.size() != coords.n_rows || nb.hd_idx.size() != coords.n_rows) {
        std::cerr << "error: initial states, free energies and neighbourhood must have one entry per frame." << std::endl;
        exit(EXIT_FAILURE);
      }
      log << "    assigning low density states to initial states" << std::endl;
      std::vector<uint32_t> init32(initial.begin(), initial.end()), assigned(coords.n_rows), named(coords.n_rows);
      if (dcb200_assign_low_density_frames(init32.data(), nb.hd_idx.data(), free_energies.data(), coords.n_rows, assigned.data()))
        die_cuda("assign_low_density_frames");
      log << "    sorting and renaming states by decreasing population" << std::endl;
      if (dcb200_sorted_cluster_names(assigned.data(), coords.n_rows, named.data())) die_cuda("sorted_cluster_names");
      log << "    storing states in: " << output_file << std::endl;
      write_clustered_trajectory(output_file, named.data(), named.size(), header_comment, comments);
    } else if (args.count("threshold-screening")) {
      log << "\n~~~ free energy screening" << std::endl;
      std::vector<float> tp;
      for (const std::string& s : args.values[args.index("threshold-screening")]) tp.push_back(to_float(s, "threshold-screening"));
      if (tp.size() > 3) {
        std::cerr << "error: option -T expects at most three floating point arguments: FROM STEP TO." << std::endl;
        exit(EXIT_FAILURE);
      }
      if (free_energies.size() != coords.n_rows || nb.nn_d2.size() != coords.n_rows) {
        std::cerr << "error: screening needs free energies and nearest neighbours of every frame." << std::endl;
        exit(EXIT_FAILURE);
      }
      float t_from = 0.1, t_step = 0.1;
      float t_to = *std::max_element(free_energies.begin(), free_energies.end());
      if (tp.size() >= 1 && tp[0] >= 0.0f) t_from = tp[0];
      if (tp.size() >= 2) t_step = tp[1];
      if (tp.size() == 3) t_to = tp[2];
      if (!(has2digits(t_from) && has2digits(t_step))) {
        std::cerr << "error: -T can handle at maximum two digits." << std::endl;
        exit(EXIT_FAILURE);
      }
      comments["screening_to"] = t_to;
      comments["screening_from"] = t_from;
      comments["screening_step"] = t_step;
      log << "\n        fe    frames" << std::endl;
      // the loop variable is a float that accumulates (density_clustering.cpp:804-806): the accumulated value is both
      // the threshold compared with the free energies and the number printed into the file name
      const float t_to_low = t_to - t_step / 10.0f + t_step;
      const float t_to_high = t_to + t_step / 10.0f + t_step;
      // one screening run for all thresholds: free energies sorted once, sorted coordinates resident on the device(s);
      // the labels are those of the reference's call-per-threshold loop (density_clustering.cpp:806-816)
      std::vector<uint32_t> clustering(coords.n_rows);
      const char* lb = getenv("DCB200_LABELS_BIN");
      const bool labels_bin = lb && lb[0] == '1';
      bool first_label_file = true;
      dcb200_screening_run* run = nullptr;
      if (dcb200_screening_begin(free_energies.data(), nb.nn_d2.data(), coords.data.data(), coords.n_rows, coords.n_cols, &run))
        die_cuda("screening");
      for (float t = t_from; (t < t_to_low) && !(t_to_high < t); t += t_step) {
        if (dcb200_screening_next(run, t, clustering.data())) die_cuda("screening");
        if (verbose) {
          size_t below = 0;
          for (float f : free_energies) below += f <= t;
          std::cout << "    " << std::setw(6) << stringprintf("%.2f", t) << " " << std::setw(9) << below << std::endl;
        }
        const std::string label_file = stringprintf((output_file + ".%0.2f").c_str(), t);
        write_clustered_trajectory(label_file, clustering.data(), clustering.size(), header_comment, comments);
        // DCB200_LABELS_BIN=1: also keep the labels as raw uint32 in <out>.dcb200labels for `clustering network` (8f-4)
        if (labels_bin) {
          write_labels_record(label_file, clustering.data(), clustering.size(), first_label_file);
          first_label_file = false;
        }
      }
      dcb200_screening_end(run);
    } else {
      std::cerr << "error: one of -T/-i is needed to generate output." << std::endl;
      exit(EXIT_FAILURE);
    }
  }
  log << "~~~ freeing memory" << std::endl;
  return EXIT_SUCCESS;
}

}  // namespace

int main(int argc, char* argv[]) {
  const std::string version_number = VERSION;
  std::ostringstream head;
  head << "\n" << std::string(25 - (19 + version_number.size()) / 2, ' ') << "~~~ clustering " + version_number + " ~~~\n";
  if (argc > 2) head << std::string(25 - (4 + strlen(argv[1])) / 2, ' ') << "~ " << argv[1] << " ~\n";
  const std::string copyright = head.str() +
                                "\nclustering " + version_number + " file formats: density mode of moldyn/Clustering on libdcb200 (sm_100a CUDA)\n\n";
  const std::string general_help = copyright +
                                   "modes:\n"
                                   "  density: run density clustering\n"
                                   "\n"
                                   "usage:\n"
                                   "  clustering MODE --option1 --option2 ...\n"
                                   "\n"
                                   "for a list of available options per mode, run with '-h' option, e.g.\n"
                                   "  clustering density -h\n\n"
                                   "this binary is parallized with cuda\n\n";
  if (argc <= 2) {
    std::cerr << general_help;
    return EXIT_FAILURE;
  }
  if (strcmp(argv[1], "density") != 0) {
    std::cerr << "\nerror: unrecognized mode '" << argv[1] << "'\n\n"
              << "(this build provides the density mode only)\n\n" << general_help;
    return EXIT_FAILURE;
  }
  Args args;
  try {
    args = parse(argc, argv);
    if (args.count("help")) {
      std::cout << usage(copyright, argv[1]) << std::endl;
      return EXIT_SUCCESS;
    }
    if (!args.count("file")) throw ParseError{"the option '--file' is required but missing"};
    if (args.count("radius")) to_float(args.str("radius"), "radius");
    if (args.count("nthreads")) {
      char* end = nullptr;
      strtol(args.str("nthreads").c_str(), &end, 10);
      if (*end != '\0') throw ParseError{"the argument ('" + args.str("nthreads") + "') for option '--nthreads' is invalid"};
    }
    for (const char* o : {"radii", "threshold-screening"})
      if (args.count(o))
        for (const std::string& s : args.values[args.index(o)]) to_float(s, o);
  } catch (const ParseError& e) {
    std::cerr << "\nerror parsing arguments:\n\n" << e.what << "\n\n" << std::endl;
    std::cerr << usage(copyright, argv[1]) << std::endl;
    return EXIT_FAILURE;
  }
  // like the reference with USE_CUDA (clustering.cpp:110-113): no usable GPU is fatal, there is no CPU path
  int n_gpus = 0;
  if (dcb200_device_count(&n_gpus) || n_gpus == 0) {
    std::cerr << "error: no CUDA-compatible GPUs found" << std::endl;
    return EXIT_FAILURE;
  }
  verbose = args.count("verbose");
  if (verbose) {
    std::cout << "\n" << head.str() << std::endl;
    std::cout << "~~~ using for parallization: CUDA" << std::endl;
  }
  // header comment of every output file (clustering.cpp:467-482)
  std::ostringstream header;
  time_t rawtime;
  time(&rawtime);
  struct tm* timeinfo = localtime(&rawtime);
  header << "# clustering " + version_number + " - " << argv[1] << "\n"
         << "#\n"
         << "# Created " << asctime(timeinfo) << "# by following command:\n#\n# ";
  for (int i = 0; i < argc; ++i) header << argv[i] << " ";
  header << "\n#\n# Copyright (c) 2015-2019 Florian Sittel and Daniel Nagel\n"
         << "# please cite the corresponding paper, "
         << "see https://github.com/moldyn/clustering\n";
  try {
    return density_main(args, header.str());
  } catch (const dcb_cli::IoError& e) {
    // the reference's tools print the message and exit (tools.hxx:44-47, :236-243, tools.cpp:106-110)
    std::cerr << e.what() << std::endl;
    return EXIT_FAILURE;
  }
}
