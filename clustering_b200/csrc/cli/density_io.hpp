// File formats of `clustering density` (reference: src/tools.hxx, src/tools.cpp, src/clustering.cpp:467-493).
// Written from the formats' behaviour, byte-compatible with the reference's writers and
// tolerant in the same way as its readers; plain C stdio / strto* instead of iostreams, so that multi-million
// line files cost tenths of a second instead of seconds.
#pragma once
#include <cstddef>
#include <cstdint>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

namespace dcb_cli {

// An unreadable / unwritable / empty file.  The `clustering` driver prints the message and exits like the reference's
// tools do; the C ABI (dcb200_io_*) turns it into an error status + dcb200_last_error() -- a library must not end the
// process that embeds it.
struct IoError : std::runtime_error {
  explicit IoError(const std::string& msg) : std::runtime_error(msg) {}
};

// "#@ key = value" parameters carried from file to file (reference: commentsMap, clustering.cpp:483-493)
typedef std::map<std::string, float> CommentsMap;
CommentsMap default_comments();

// row-major float[n_rows][n_cols]; rows = non-empty lines, columns = tokens of the first non-empty line,
// values = the first n_rows*n_cols whitespace separated tokens (reference: read_coords, tools.hxx:39-111)
struct Coords {
  std::vector<float> data;
  std::size_t n_rows = 0, n_cols = 0;
};
Coords read_coords(const std::string& filename);

// one number per line, lines that do not start with a number are skipped (tools.hxx:229-253)
std::vector<float> read_single_column_float(const std::string& filename);
std::vector<std::size_t> read_single_column_size(const std::string& filename);
// four columns: id(nn) dsqr(nn) id(nn_hd) dsqr(nn_hd)  (tools.cpp:103-133)
void read_neighborhood(const std::string& filename, std::vector<uint32_t>& nn_idx, std::vector<float>& nn_d2,
                       std::vector<uint32_t>& hd_idx, std::vector<float>& hd_d2);
// re-reads the "#@ key = value" lines of a file into the map; warns (verbose) when a non-zero value differs by
// more than 1e-3 (tools.cpp:229-265)
void read_comments(const std::string& filename, CommentsMap& comments);

// header + "#@" block + one value per line (tools.cpp:42-56, :64-70, :144-174, :267-277)
void write_pops(const std::string& filename, const uint32_t* pops, std::size_t n, const std::string& header, const CommentsMap& comments);
void write_fes(const std::string& filename, const float* fe, std::size_t n, const std::string& header, const CommentsMap& comments);
void write_clustered_trajectory(const std::string& filename, const uint32_t* traj, std::size_t n, const std::string& header,
                                const CommentsMap& comments);
void write_neighborhood(const std::string& filename, const uint32_t* nn_idx, const float* nn_d2, const uint32_t* hd_idx,
                        const float* hd_d2, std::size_t n, const std::string& header, const CommentsMap& comments);

// binary side channel of the per-threshold label files: record of `text_file`'s labels in the container
// <out>.dcb200labels next to it; read_labels_record is true when a record for the file exists AND the ASCII file on
// disk still has the size it had when the record was written
void write_labels_record(const std::string& text_file, const uint32_t* states, std::size_t n, bool truncate);
bool read_labels_record(const std::string& text_file, std::vector<uint32_t>& out);

std::string comments_block(const CommentsMap& comments);      // append_commentsMap
std::string stringprintf(const char* fmt, double v);           // one float argument is all the driver needs

extern bool verbose;                                           // Clustering::verbose (logger.cpp:28-38)
}  // namespace dcb_cli
