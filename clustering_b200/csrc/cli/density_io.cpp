#include "density_io.hpp"

#include <algorithm>
#include <cctype>
#include <cerrno>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>

namespace dcb_cli {

bool verbose = false;

CommentsMap default_comments() {
  CommentsMap m;
  const char* keys[] = {"clustering_radius", "lumping_radius", "screening_from", "screening_to", "screening_step",
                        "minimal_population", "cmin", "single_coring_time", "limits"};
  for (const char* k : keys) m[k] = 0.f;
  return m;
}

static std::string slurp(const std::string& filename, const char* what_for) {
  FILE* f = fopen(filename.c_str(), "rb");
  if (!f) throw IoError("error: cannot open file '" + filename + "'" + what_for);
  std::string buf;
  fseek(f, 0, SEEK_END);
  const long sz = ftell(f);
  fseek(f, 0, SEEK_SET);
  if (sz > 0) {
    buf.resize((size_t) sz);
    const size_t got = fread(&buf[0], 1, (size_t) sz, f);
    buf.resize(got);
  }
  fclose(f);
  return buf;
}

static inline bool is_space(char c) { return c == ' ' || c == '\t' || c == '\n' || c == '\r' || c == '\v' || c == '\f'; }

Coords read_coords(const std::string& filename) {
  if (verbose) std::cout << "~~~ reading coordinates" << std::endl;
  const std::string buf = slurp(filename, "");
  if (verbose) std::cout << "    from file: " << filename << std::endl;
  Coords c;
  // columns: tokens of the first non-empty line; rows: non-empty lines (a line of blanks counts, as in the reference)
  const char* p = buf.data();
  const char* end = p + buf.size();
  bool first_seen = false;
  while (p < end) {
    const char* nl = (const char*) memchr(p, '\n', (size_t) (end - p));
    const char* le = nl ? nl : end;
    if (le > p) {
      ++c.n_rows;
      if (!first_seen) {
        first_seen = true;
        const char* q = p;
        while (q < le) {
          while (q < le && is_space(*q)) ++q;
          if (q < le) ++c.n_cols;
          while (q < le && !is_space(*q)) ++q;
        }
      }
    }
    p = nl ? nl + 1 : end;
  }
  if (verbose) std::cout << "    with dimensions: " << c.n_rows << "x" << c.n_cols << "\n" << std::endl;
  c.data.assign(c.n_rows * c.n_cols, 0.f);
  // values: the token stream, whatever the line structure (the reference extracts n_rows*n_cols floats in a row)
  const char* q = buf.c_str();
  for (std::size_t k = 0; k < c.data.size(); ++k) {
    char* stop = nullptr;
    const float v = strtof(q, &stop);
    if (stop == q) break;                     // not a number: the reference's stream fails here and leaves the rest untouched
    c.data[k] = v;
    q = stop;
  }
  return c;
}

// Token-wise reader shared by the single-column and neighbourhood files: `fields` numbers in a row form a record;
// when a conversion fails the rest of the line is dropped (comment lines), exactly the recovery of the reference.
template <class OnRecord>
static void scan_records(const std::string& buf, const int* is_float, int fields, OnRecord&& on_record) {
  const char* q = buf.c_str();
  const char* end = q + buf.size();
  double vals[4];
  while (q < end) {
    int got = 0;
    const char* r = q;
    for (; got < fields; ++got) {
      char* stop = nullptr;
      if (is_float[got]) {
        vals[got] = (double) strtof(r, &stop);
      } else {
        const char* s = r;
        while (s < end && is_space(*s)) ++s;
        if (s >= end || !(isdigit((unsigned char) *s) || *s == '+')) break;     // unsigned extraction
        vals[got] = (double) strtoull(s, &stop, 10);
      }
      if (stop == r || stop == nullptr) break;
      r = stop;
    }
    if (got == fields) {
      on_record(vals);
      q = r;
    } else {
      // skip whitespace; at end of input stop, otherwise drop the rest of this line
      const char* s = r;
      while (s < end && is_space(*s)) ++s;
      if (s >= end) break;
      const char* nl = (const char*) memchr(s, '\n', (size_t) (end - s));
      q = nl ? nl + 1 : end;
    }
  }
}

std::vector<float> read_single_column_float(const std::string& filename) {
  const std::string buf = slurp(filename, "");
  std::vector<float> out;
  const int kinds[1] = {1};
  scan_records(buf, kinds, 1, [&](const double* v) { out.push_back((float) v[0]); });
  if (out.empty()) throw IoError("error: opened empty file '" + filename + "'");
  return out;
}

std::vector<std::size_t> read_single_column_size(const std::string& filename) {
  const std::string buf = slurp(filename, "");
  std::vector<std::size_t> out;
  const int kinds[1] = {0};
  scan_records(buf, kinds, 1, [&](const double* v) { out.push_back((std::size_t) v[0]); });
  if (out.empty()) throw IoError("error: opened empty file '" + filename + "'");
  return out;
}

void read_neighborhood(const std::string& filename, std::vector<uint32_t>& nn_idx, std::vector<float>& nn_d2,
                       std::vector<uint32_t>& hd_idx, std::vector<float>& hd_d2) {
  const std::string buf = slurp(filename, " for reading.");
  const int kinds[4] = {0, 1, 0, 1};
  scan_records(buf, kinds, 4, [&](const double* v) {
    nn_idx.push_back((uint32_t) v[0]);
    nn_d2.push_back((float) v[1]);
    hd_idx.push_back((uint32_t) v[2]);
    hd_d2.push_back((float) v[3]);
  });
}

void read_comments(const std::string& filename, CommentsMap& comments) {
  const std::string buf = slurp(filename, "");
  const char* p = buf.data();
  const char* end = p + buf.size();
  while (p < end) {
    const char* nl = (const char*) memchr(p, '\n', (size_t) (end - p));
    const char* le = nl ? nl : end;
    const char* q = p;
    while (q < le && is_space(*q)) ++q;
    if (le - q >= 2 && q[0] == '#' && q[1] == '@' && (q + 2 == le || is_space(q[2]))) {
      q += 2;
      while (q < le && is_space(*q)) ++q;
      const char* k0 = q;
      while (q < le && !is_space(*q)) ++q;
      const std::string key(k0, q);
      auto it = comments.find(key);
      if (it != comments.end()) {
        // the value: first token after the key that starts with a digit ("= 0.30000"); -1 if the line ends first
        float val = -1.f;
        while (q < le) {
          while (q < le && is_space(*q)) ++q;
          if (q < le && isdigit((unsigned char) *q)) {
            val = strtof(std::string(q, le).c_str(), nullptr);
            break;
          }
          while (q < le && !is_space(*q)) ++q;
        }
        if (it->second != 0 && std::abs(it->second - val) > 0.001 && verbose) {
          std::cout << "warning: the values of " << key << " are not in agreement\n"
                    << "        " << val << " vs. " << it->second << std::endl;
        }
        it->second = val;
      }
    }
    p = nl ? nl + 1 : end;
  }
}

std::string stringprintf(const char* fmt, double v) {
  char small[256];
  const int n = snprintf(small, sizeof(small), fmt, v);
  if (n < (int) sizeof(small)) return std::string(small);
  std::string big((size_t) n + 1, '\0');
  snprintf(&big[0], big.size(), fmt, v);
  big.resize((size_t) n);
  return big;
}

std::string comments_block(const CommentsMap& comments) {
  std::string s = "#\n# The following comments are reused for identifying\n# user-based mistakes and should not be modified.\n";
  for (const auto& kv : comments) {
    if (kv.second != 0.) {
      char line[512];
      snprintf(line, sizeof(line), "#@   %s = %.5f\n", kv.first.c_str(), (double) kv.second);
      s += line;
    }
  }
  return s;
}

namespace {
struct OutFile {
  FILE* f;
  std::string buf;
  explicit OutFile(const std::string& filename) {
    f = fopen(filename.c_str(), "wb");
    if (!f) throw IoError("error: cannot open file '" + filename + "' for writing.");
    buf.reserve(1 << 22);
  }
  void flush_if_big() {
    if (buf.size() > (1 << 22) - 256) {
      fwrite(buf.data(), 1, buf.size(), f);
      buf.clear();
    }
  }
  ~OutFile() {
    fwrite(buf.data(), 1, buf.size(), f);
    fclose(f);
  }
};

inline void put_uint(std::string& s, unsigned long long v) {
  char tmp[24];
  int n = 0;
  do { tmp[n++] = (char) ('0' + v % 10); v /= 10; } while (v);
  while (n) s.push_back(tmp[--n]);
}
}  // namespace

static void write_uint_column(const std::string& filename, const uint32_t* v, std::size_t n, const std::string& head) {
  OutFile o(filename);
  o.buf += head;
  for (std::size_t i = 0; i < n; ++i) {
    put_uint(o.buf, v[i]);
    o.buf.push_back('\n');
    o.flush_if_big();
  }
}

void write_pops(const std::string& filename, const uint32_t* pops, std::size_t n, const std::string& header, const CommentsMap& comments) {
  write_uint_column(filename, pops, n, header + comments_block(comments) + "#\n# point density of each frame\n");
}

void write_clustered_trajectory(const std::string& filename, const uint32_t* traj, std::size_t n, const std::string& header,
                                const CommentsMap& comments) {
  write_uint_column(filename, traj, n, header + comments_block(comments) + "#\n# state/cluster id frames are assigned to\n");
}

void write_fes(const std::string& filename, const float* fe, std::size_t n, const std::string& header, const CommentsMap& comments) {
  OutFile o(filename);
  o.buf += header + comments_block(comments) + "#\n# free energy of each frame\n";
  char tmp[48];
  for (std::size_t i = 0; i < n; ++i) {
    const int k = snprintf(tmp, sizeof(tmp), "%e\n", (double) fe[i]);       // std::scientific, precision 6
    o.buf.append(tmp, (size_t) k);
    o.flush_if_big();
  }
}

void write_neighborhood(const std::string& filename, const uint32_t* nn_idx, const float* nn_d2, const uint32_t* hd_idx,
                        const float* hd_d2, std::size_t n, const std::string& header, const CommentsMap& comments) {
  OutFile o(filename);
  o.buf += header + comments_block(comments) +
           "#\n# column definitions:\n"
           "#        nn = nearest neighbor\n"
           "#     nn_hd = nearest neighbor with higher density\n"
           "#     id(i) = id/line number of i\n"
           "#   dsqr(i) = squared euclidean distance to i\n#\n"
           "# id(nn)  dsqr(nn) id(nn_hd) dsqr(nn_hd)\n";
  char tmp[48];
  for (std::size_t i = 0; i < n; ++i) {
    put_uint(o.buf, nn_idx[i]);
    int k = snprintf(tmp, sizeof(tmp), " %g ", (double) nn_d2[i]);            // default ostream float format: %g, 6 digits
    o.buf.append(tmp, (size_t) k);
    put_uint(o.buf, hd_idx[i]);
    k = snprintf(tmp, sizeof(tmp), " %g\n", (double) hd_d2[i]);
    o.buf.append(tmp, (size_t) k);
    o.flush_if_big();
  }
}

}  // namespace dcb_cli

// ---- binary side channel of the per-threshold label files ------------------------------------------
// `clustering density -T` writes one N-line ASCII file per threshold (<out>.0.10, <out>.0.20, ...: 100+ files at 5M frames)
// and `clustering network` reads them all back, one read_clustered_trajectory per file (network_builder.cpp:411-437): the
// ASCII round trip is most of that mode's time.  The container <out>.dcb200labels keeps the same labels as raw uint32,
// one record per file; a reader that finds a record for the file it was asked for -- and whose recorded ASCII size
// still matches the file on disk -- takes the binary copy instead of parsing.  The ASCII files stay the format of record.
namespace dcb_cli {
namespace {
const char LABELS_MAGIC[8] = {'D', 'C', 'B', '2', 'L', 'B', 'L', '1'};

// "<base>.<int>.<2 digits>" (the reference's "%0.2f" suffix) -> "<base>.dcb200labels"; empty if the name has no such suffix
std::string container_of(const std::string& filename) {
  const size_t n = filename.size();
  if (n < 5 || !isdigit((unsigned char) filename[n - 1]) || !isdigit((unsigned char) filename[n - 2]) || filename[n - 3] != '.') return "";
  size_t p = n - 3;
  size_t q = p;
  while (q > 0 && isdigit((unsigned char) filename[q - 1])) --q;
  if (q == p || q == 0 || filename[q - 1] != '.') return "";
  return filename.substr(0, q - 1) + ".dcb200labels";
}
std::string leaf_of(const std::string& filename) {
  const size_t s = filename.find_last_of('/');
  return s == std::string::npos ? filename : filename.substr(s + 1);
}
long long file_size(const std::string& filename) {
  FILE* f = fopen(filename.c_str(), "rb");
  if (!f) return -1;
  fseek(f, 0, SEEK_END);
  const long long sz = ftell(f);
  fclose(f);
  return sz;
}
}  // namespace

void write_labels_record(const std::string& text_file, const uint32_t* states, std::size_t n, bool truncate) {
  const std::string container = container_of(text_file);
  if (container.empty()) throw IoError("error: '" + text_file + "' does not end in a threshold suffix (.%0.2f): no label container for it");
  const long long text_bytes = file_size(text_file);
  if (text_bytes < 0) throw IoError("error: cannot open file '" + text_file + "'");
  FILE* f = fopen(container.c_str(), truncate ? "wb" : "ab");
  if (!f) throw IoError("error: cannot open file '" + container + "' for writing.");
  const std::string leaf = leaf_of(text_file);
  const uint64_t hdr[3] = {(uint64_t) leaf.size(), (uint64_t) n, (uint64_t) text_bytes};
  bool ok = fwrite(LABELS_MAGIC, 1, 8, f) == 8 && fwrite(hdr, sizeof(uint64_t), 3, f) == 3 &&
            fwrite(leaf.data(), 1, leaf.size(), f) == leaf.size() && fwrite(states, sizeof(uint32_t), n, f) == n;
  ok = (fclose(f) == 0) && ok;
  if (!ok) throw IoError("error: short write to '" + container + "'");
}

bool read_labels_record(const std::string& text_file, std::vector<uint32_t>& out) {
  const std::string container = container_of(text_file);
  if (container.empty()) return false;
  FILE* f = fopen(container.c_str(), "rb");
  if (!f) return false;
  const std::string leaf = leaf_of(text_file);
  const long long text_bytes = file_size(text_file);
  bool found = false;
  for (;;) {                                   // the last record of a name wins (a re-run appends)
    char magic[8];
    uint64_t hdr[3];
    if (fread(magic, 1, 8, f) != 8 || memcmp(magic, LABELS_MAGIC, 8) != 0 || fread(hdr, sizeof(uint64_t), 3, f) != 3) break;
    if (hdr[0] > 4096) break;
    std::string name((size_t) hdr[0], '\0');
    if (hdr[0] && fread(&name[0], 1, (size_t) hdr[0], f) != hdr[0]) break;
    if (name == leaf && (long long) hdr[2] == text_bytes) {
      std::vector<uint32_t> v((size_t) hdr[1]);
      if (fread(v.data(), sizeof(uint32_t), v.size(), f) != v.size()) break;
      out.swap(v);
      found = true;
    } else if (fseek(f, (long) (hdr[1] * sizeof(uint32_t)), SEEK_CUR) != 0) {
      break;
    }
  }
  fclose(f);
  return found;
}
}  // namespace dcb_cli

// ---- C ABI of the file formats (include/dcb200.h, section "file formats") -------------------------------------
#include "../../../include/dcb200.h"

extern "C" int dcb200_internal_fail(const char* msg);

namespace {
dcb_cli::CommentsMap map_of(const char* const* keys, const float* vals, size_t n) {
  dcb_cli::CommentsMap m;
  for (size_t i = 0; i < n; ++i) m[keys[i]] = vals[i];
  return m;
}
// runs fn; an I/O failure becomes an error status (message through dcb200_last_error), never an exit
template <class Fn>
int guarded(const char* what, Fn&& fn) {
  try {
    fn();
    return 0;
  } catch (const std::exception& e) {
    return dcb200_internal_fail((std::string(what) + ": " + e.what()).c_str());
  }
}
}  // namespace

#define DCB_IO_NEED(cond, what) \
  if (!(cond)) return dcb200_internal_fail(what ": null argument")

extern "C" int dcb200_io_write_pops(const char* filename, const uint32_t* pops, size_t n, const char* header, const char* const* keys,
                                    const float* vals, size_t n_comments) {
  DCB_IO_NEED(filename && (pops || !n) && ((keys && vals) || !n_comments), "dcb200_io_write_pops");
  return guarded("dcb200_io_write_pops", [&] { dcb_cli::write_pops(filename, pops, n, header ? header : "", map_of(keys, vals, n_comments)); });
}
extern "C" int dcb200_io_write_fes(const char* filename, const float* fe, size_t n, const char* header, const char* const* keys,
                                   const float* vals, size_t n_comments) {
  DCB_IO_NEED(filename && (fe || !n) && ((keys && vals) || !n_comments), "dcb200_io_write_fes");
  return guarded("dcb200_io_write_fes", [&] { dcb_cli::write_fes(filename, fe, n, header ? header : "", map_of(keys, vals, n_comments)); });
}
extern "C" int dcb200_io_write_states(const char* filename, const uint32_t* states, size_t n, const char* header,
                                      const char* const* keys, const float* vals, size_t n_comments) {
  DCB_IO_NEED(filename && (states || !n) && ((keys && vals) || !n_comments), "dcb200_io_write_states");
  return guarded("dcb200_io_write_states",
                 [&] { dcb_cli::write_clustered_trajectory(filename, states, n, header ? header : "", map_of(keys, vals, n_comments)); });
}
extern "C" int dcb200_io_write_neighborhood(const char* filename, const uint32_t* nn_idx, const float* nn_d2, const uint32_t* hd_idx,
                                            const float* hd_d2, size_t n, const char* header, const char* const* keys,
                                            const float* vals, size_t n_comments) {
  DCB_IO_NEED(filename && ((nn_idx && nn_d2 && hd_idx && hd_d2) || !n) && ((keys && vals) || !n_comments), "dcb200_io_write_neighborhood");
  return guarded("dcb200_io_write_neighborhood", [&] {
    dcb_cli::write_neighborhood(filename, nn_idx, nn_d2, hd_idx, hd_d2, n, header ? header : "", map_of(keys, vals, n_comments));
  });
}
extern "C" int dcb200_io_read_coords(const char* filename, float* out, size_t capacity, size_t* n_rows, size_t* n_cols) {
  DCB_IO_NEED(filename && n_rows && n_cols, "dcb200_io_read_coords");
  return guarded("dcb200_io_read_coords", [&] {
    const dcb_cli::Coords c = dcb_cli::read_coords(filename);
    *n_rows = c.n_rows;
    *n_cols = c.n_cols;
    if (out) memcpy(out, c.data.data(), std::min(capacity, c.data.size()) * sizeof(float));
  });
}
extern "C" int dcb200_io_read_column_float(const char* filename, float* out, size_t capacity, size_t* n) {
  DCB_IO_NEED(filename && n, "dcb200_io_read_column_float");
  return guarded("dcb200_io_read_column_float", [&] {
    const std::vector<float> v = dcb_cli::read_single_column_float(filename);
    *n = v.size();
    if (out) memcpy(out, v.data(), std::min(capacity, v.size()) * sizeof(float));
  });
}
extern "C" int dcb200_io_read_column_uint(const char* filename, uint32_t* out, size_t capacity, size_t* n) {
  DCB_IO_NEED(filename && n, "dcb200_io_read_column_uint");
  return guarded("dcb200_io_read_column_uint", [&] {
    const std::vector<std::size_t> v = dcb_cli::read_single_column_size(filename);
    *n = v.size();
    if (out)
      for (size_t i = 0; i < std::min(capacity, v.size()); ++i) out[i] = (uint32_t) v[i];
  });
}
extern "C" int dcb200_io_read_neighborhood(const char* filename, uint32_t* nn_idx, float* nn_d2, uint32_t* hd_idx, float* hd_d2,
                                           size_t capacity, size_t* n) {
  DCB_IO_NEED(filename && n, "dcb200_io_read_neighborhood");
  return guarded("dcb200_io_read_neighborhood", [&] {
    std::vector<uint32_t> a, c;
    std::vector<float> b, d;
    dcb_cli::read_neighborhood(filename, a, b, c, d);
    *n = a.size();
    const size_t m = std::min(capacity, a.size());
    if (nn_idx) memcpy(nn_idx, a.data(), m * 4);
    if (nn_d2) memcpy(nn_d2, b.data(), m * 4);
    if (hd_idx) memcpy(hd_idx, c.data(), m * 4);
    if (hd_d2) memcpy(hd_d2, d.data(), m * 4);
  });
}
extern "C" int dcb200_io_read_comment(const char* filename, const char* key, float current, float* value) {
  DCB_IO_NEED(filename && key && value, "dcb200_io_read_comment");
  return guarded("dcb200_io_read_comment", [&] {
    dcb_cli::CommentsMap m;
    m[key] = current;
    dcb_cli::read_comments(filename, m);
    *value = m[key];
  });
}

// binary side channel of the per-threshold label files (see write_labels_record)
extern "C" int dcb200_io_write_states_record(const char* text_filename, const uint32_t* states, size_t n, int truncate) {
  DCB_IO_NEED(text_filename && (states || !n), "dcb200_io_write_states_record");
  return guarded("dcb200_io_write_states_record", [&] { dcb_cli::write_labels_record(text_filename, states, n, truncate != 0); });
}
extern "C" int dcb200_io_read_states(const char* filename, uint32_t* out, size_t capacity, size_t* n, int* from_binary) {
  DCB_IO_NEED(filename && n, "dcb200_io_read_states");
  return guarded("dcb200_io_read_states", [&] {
    std::vector<uint32_t> v;
    const bool bin = dcb_cli::read_labels_record(filename, v);
    if (!bin) {
      const std::vector<std::size_t> t = dcb_cli::read_single_column_size(filename);
      v.assign(t.begin(), t.end());
    }
    if (from_binary) *from_binary = bin ? 1 : 0;
    *n = v.size();
    if (out) memcpy(out, v.data(), std::min(capacity, v.size()) * sizeof(uint32_t));
  });
}
