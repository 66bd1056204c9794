/* dcb200 -- C ABI of the B200-native `clustering density` hot path (libdcb200.so).
 *
 * Drop-in boundary for moldyn/Clustering (citations are file:line under the reference's src/):
 * these are the entry points the reference's Clustering::Density::CUDA functions
 * (density_clustering_cuda.hpp:13-54, called from density_clustering.cpp:113-118, :616-621, :659-663,
 * :716-720, :746-750, :808-814 and clustering.cpp:110-113) bind to.  The C++ shim with the reference's own
 * signatures lives in clustering_b200/csrc/density_cuda.hpp.
 *
 * Conventions
 *   - plain pointers and sizes only; every function returns 0 on success, non-zero on failure
 *     (dcb200_last_error() gives the message of the calling thread's last failure);
 *   - coords is the reference's row-major float[n_rows][n_cols] array (tools.hxx:39-111);
 *   - results are bit-identical to the reference's CPU/OpenMP build (CPU semantics,
 *     CPU rounding order): strict '<' radius test, self counted once, nearest-neighbour ties -> smallest
 *     frame index, "no neighbour" = (n_rows+1, FLT_MAX), squared distances stored;
 *   - there is no CPU fallback: without a CUDA device every compute entry point fails.
 *
 * Two layers:
 *   (1) host-pointer entry points (dcb200_populations, ...): allocate/upload/compute/download per call on
 *       all selected GPUs (rows sharded across devices), exactly like the reference's CUDA path;
 *   (2) a device-resident session (dcb200_ctx_*): coordinates stay in HBM, outputs are device pointers,
 *       row shards are explicit -- used by the benchmark, by one-process-per-GPU drivers
 *       (torch.distributed / NCCL) and by the host-pointer layer itself.
 */
#ifndef DCB200_H
#define DCB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DCB200_VERSION 200

/* ---- process-wide ------------------------------------------------------------------------- */
/* replaces Clustering::Density::CUDA::get_num_gpus (density_clustering_cuda.cu:32-43) */
int dcb200_device_count(int* n_devices);
/* number of GPUs the host-pointer entry points shard over (default: all visible); 0 restores the default */
int dcb200_set_gpus(int n_gpus);
const char* dcb200_last_error(void);
int dcb200_version(void);

/* ---- (1) host-pointer entry points -------------------------------------------------------- */
/* replaces CUDA::calculate_populations (density_clustering_cuda.cu:139-182; CPU: density_clustering.cpp:126-195)
 * pops: uint32 [n_radii][n_rows]; row r belongs to radii[r] (input order, duplicates allowed). */
int dcb200_populations(const float* coords, size_t n_rows, size_t n_cols, const float* radii, size_t n_radii,
                       uint32_t* pops);
/* replaces calculate_free_energies (density_clustering.cpp:197-212): fe[i] = -log(pops[i] / max(pops)) */
int dcb200_free_energies(const uint32_t* pops, size_t n_rows, float* fe);
/* replaces CUDA::nearest_neighbors (density_clustering_cuda.cu:286-328; CPU: density_clustering.cpp:230-288)
 * nn_*: nearest neighbour; hd_*: nearest neighbour with strictly lower free energy; *_d2 are SQUARED distances */
int dcb200_nearest_neighbors(const float* coords, size_t n_rows, size_t n_cols, const float* fe, uint32_t* nn_idx,
                             float* nn_d2, uint32_t* hd_idx, float* hd_d2);
/* replaces the pair scan + cluster merging of CUDA::screening (density_clustering_cuda.cu:396-594;
 * CPU: density_clustering_common.cpp:37-134, density_clustering.cpp:292-332, :506-555).
 *   sorted_coords: frames in ascending free-energy order (the order of sorted_free_energies, :214-228),
 *                  row-major [n_sorted][n_cols], at least m_new rows
 *   m_prev / m_new: number of sorted frames below the previous / the current free-energy threshold
 *   max_dist2:     (float)(4 * sigma2)
 *   comp:          uint32 [m_new], in/out.  In: for p < m_prev the component representative of the previous
 *                  threshold (smallest sorted position of p's cluster; pass p itself when m_prev == 0).
 *                  Out: for every p < m_new the smallest sorted position of the cluster p belongs to.
 * The caller maps representatives to the reference's cluster numbers (1..K by ascending representative). */
int dcb200_screening_step(const float* sorted_coords, size_t n_cols, size_t m_prev, size_t m_new, float max_dist2,
                          uint32_t* comp);
/* The same result as a session over ALL free-energy-sorted frames, by ONE neighbour-graph scan instead of one pair scan
 * per threshold: at its first step the session scans the frames -- uploaded once, laid out in spatial order on every
 * selected GPU, rows dealt block-cyclically -- for every pair with d2 < max_dist2 and sorts these edges by the sorted
 * index of their later frame; a threshold then only unions the edges that became active since the previous one.  (The
 * reference scans (new frames) x (all lower frames) at every threshold: N^2/2 pairs per run; the edge list costs about
 * one population scan with r^2 = max_dist2.)  dcb200_screen_step extends the forest from the positions done so far to
 * [0, m_new) (m_new must not decrease) and writes every position's representative to comp[0, m_new); seed (or NULL)
 * replaces the session's forest for the positions done so far (uint32 [positions done], seed[p] <= p).
 * Sessions share the library's per-GPU contexts with the other host-pointer entry points: a call in between is allowed
 * (the session notices that its layout was replaced and restores it from its host copy), it just costs an upload. */
typedef struct dcb200_screen dcb200_screen;
int dcb200_screen_begin(const float* sorted_coords, size_t n_sorted, size_t n_cols, dcb200_screen** session);
int dcb200_screen_step(dcb200_screen* session, size_t m_new, float max_dist2, const uint32_t* seed, uint32_t* comp);
/* the same step with the cluster numbers made on the device: after dcb200_screen_set_order (order[p] = frame at sorted
 * position p, uint32 [n_sorted]) dcb200_screen_labels writes the labels of ALL frames in FRAME order -- clusters numbered
 * 1..K by ascending representative, 0 above the threshold -- and the number of clusters */
int dcb200_screen_set_order(dcb200_screen* session, const uint32_t* order);
int dcb200_screen_labels(dcb200_screen* session, size_t m_new, float max_dist2, uint32_t* labels, uint32_t* n_clusters);
int dcb200_screen_end(dcb200_screen* session);

/* One density run without leaving the device(s) between its stages -- what `clustering density -r/-R ... -p -d -b` computes
 * (density_clustering.cpp:596-735): populations for all radii, free energies, and the neighbour search on the free
 * energies of radii[fe_radius], with ONE upload and ONE layout build (the separate entry points above upload and lay out
 * the coordinates once each, like the reference's CUDA functions).  Outputs in frame order; optional ones may be NULL:
 *   pops [n_radii][n_rows]; fe_all [n_radii][n_rows] (free energies of every radius); fe [n_rows] (of radii[fe_radius]);
 *   nn_idx/nn_d2/hd_idx/hd_d2 [n_rows] (all four or none). */
int dcb200_density_run(const float* coords, size_t n_rows, size_t n_cols, const float* radii, size_t n_radii, size_t fe_radius,
                       uint32_t* pops, float* fe_all, float* fe, uint32_t* nn_idx, float* nn_d2, uint32_t* hd_idx,
                       float* hd_d2);

/* host-side bookkeeping of the screening (no pair work; kept bit-compatible with the reference's libstdc++ calls)
 *   dcb200_sorted_free_energies: order[k] = frame at sorted position k   (sorted_free_energies, density_clustering.cpp:214-228)
 *   dcb200_sigma2:               mean squared nearest-neighbour distance  (compute_sigma2, :334-343)
 *   dcb200_screening:            one free-energy threshold, = reference screening() (density_clustering_common.cpp:37-134);
 *                                initial: NULL or ANY labelling of the frames (uint32 [n_rows], 0 = no cluster; typically the
 *                                labels of a lower threshold, density_clustering.cpp:394-427); labels: uint32 [n_rows],
 *                                0 = unassigned.  The pair scan inside runs on the GPU(s) through dcb200_screening_step.
 *   dcb200_assign_low_density_frames / dcb200_sorted_cluster_names: the microstate step after the screening
 *                                (`clustering density -i`): assign_low_density_frames (density_clustering.cpp:345-360) and
 *                                sorted_cluster_names (:458-493); uint32 [n_rows] in and out, 0 = no state */
int dcb200_sorted_free_energies(const float* fe, size_t n_rows, uint32_t* order);
int dcb200_assign_low_density_frames(const uint32_t* initial, const uint32_t* hd_idx, const float* fe, size_t n_rows,
                                     uint32_t* states);
int dcb200_sorted_cluster_names(const uint32_t* states, size_t n_rows, uint32_t* renamed);
int dcb200_sigma2(const float* nn_d2, size_t n_rows, double* sigma2);
int dcb200_screening(const float* fe, const float* nn_d2, float threshold, const float* coords, size_t n_rows,
                     size_t n_cols, const uint32_t* initial, uint32_t* labels);
/* All thresholds of one `clustering density -T` run (the driver loop density_clustering.cpp:806-816, which calls
 * screening() with the previous labels as initial clusters): the free energies are sorted ONCE and the sorted coordinates
 * stay on the device(s).  Thresholds must not decrease; labels [n_rows] are identical to the call-per-threshold form. */
typedef struct dcb200_screening_run dcb200_screening_run;
int dcb200_screening_begin(const float* fe, const float* nn_d2, const float* coords, size_t n_rows, size_t n_cols,
                           dcb200_screening_run** run);
int dcb200_screening_next(dcb200_screening_run* run, float threshold, uint32_t* labels);
int dcb200_screening_end(dcb200_screening_run* run);

/* ---- file formats of `clustering density` ---------------------------------------------------------------
 * Byte-compatible with the reference's writers (src/tools.cpp:42-56, :64-70, :144-174, :267-277, tools.hxx:256-272) and
 * tolerant like its readers (tools.hxx:39-111, :229-253, tools.cpp:103-133, :229-265).  header: the "# ..." block every
 * file starts with (clustering.cpp:467-482); keys/vals: the "#@ key = value" parameters (zero values are not written).
 * Readers: pass out = NULL (or a small capacity) to query the size first.  A file that cannot be opened (or holds no
 * value) is an error status with the reference's message in dcb200_last_error(); only the `clustering` binary turns it
 * into the reference's print-and-exit. */
int dcb200_io_write_pops(const char* filename, const uint32_t* pops, size_t n, const char* header, const char* const* keys,
                         const float* vals, size_t n_comments);
int dcb200_io_write_fes(const char* filename, const float* fe, size_t n, const char* header, const char* const* keys,
                        const float* vals, size_t n_comments);
int dcb200_io_write_states(const char* filename, const uint32_t* states, size_t n, const char* header, const char* const* keys,
                           const float* vals, size_t n_comments);
int dcb200_io_write_neighborhood(const char* filename, const uint32_t* nn_idx, const float* nn_d2, const uint32_t* hd_idx,
                                 const float* hd_d2, size_t n, const char* header, const char* const* keys, const float* vals,
                                 size_t n_comments);
int dcb200_io_read_coords(const char* filename, float* out, size_t capacity, size_t* n_rows, size_t* n_cols);
int dcb200_io_read_column_float(const char* filename, float* out, size_t capacity, size_t* n);
int dcb200_io_read_column_uint(const char* filename, uint32_t* out, size_t capacity, size_t* n);
int dcb200_io_read_neighborhood(const char* filename, uint32_t* nn_idx, float* nn_d2, uint32_t* hd_idx, float* hd_d2,
                                size_t capacity, size_t* n);
/* Binary side channel of the per-threshold label files.  `clustering density -T` writes one N-line
 * ASCII file per threshold and `clustering network` reads them all back (network_builder.cpp:411-437, one
 * read_clustered_trajectory per file): the ASCII round trip dominates that mode.
 *   dcb200_io_write_states_record: after the ASCII file <out>.<threshold> was written, stores its labels as raw uint32 in
 *       the container <out>.dcb200labels (truncate != 0: start a new container);
 *   dcb200_io_read_states: what read_clustered_trajectory returns for `filename`, taken from the container when it holds a
 *       record for that file whose recorded ASCII size still matches the file on disk (*from_binary = 1), else parsed from
 *       the ASCII file itself.  The ASCII files remain the format of record. */
int dcb200_io_write_states_record(const char* text_filename, const uint32_t* states, size_t n, int truncate);
int dcb200_io_read_states(const char* filename, uint32_t* out, size_t capacity, size_t* n, int* from_binary);
/* value of "#@ key = ..." in the file (current: the value known so far, returned unchanged if the key is absent) */
int dcb200_io_read_comment(const char* filename, const char* key, float current, float* value);

/* ---- (2) device-resident session ---------------------------------------------------------- */
/* A context holds the frames in HBM in its own order ("positions"): by default a spatial (Hilbert curve) order, so
 * that column tiles have small bounding boxes and tiles out of reach of a row block are never scanned.
 * Scans take ranges of POSITIONS (any partition of [0, n_rows) may be spread over devices -- every device
 * builds the same deterministic order) and return results in position order; dcb200_ctx_to_frame_order /
 * dcb200_ctx_nn_finish bring complete arrays back to the frame order of the input. */
typedef struct dcb200_ctx dcb200_ctx;

int dcb200_ctx_create(int device, dcb200_ctx** ctx);
int dcb200_ctx_destroy(dcb200_ctx* ctx);
/* the CUDA stream (cudaStream_t) all work of this context is enqueued on */
void* dcb200_ctx_stream(dcb200_ctx* ctx);
int dcb200_ctx_sync(dcb200_ctx* ctx);

/* upload (host pointer) or adopt (device pointer, row-major) the coordinates; builds the dim-major
 * tile layout in HBM.  The source array is not retained.
 * _ex: on_device: coords is a device pointer; keep_order: positions = frame order (needed by the screening,
 * whose input is already sorted by free energy). */
int dcb200_ctx_set_coords(dcb200_ctx* ctx, const float* host_coords, size_t n_rows, size_t n_cols);
int dcb200_ctx_set_coords_device(dcb200_ctx* ctx, const float* dev_coords, size_t n_rows, size_t n_cols);
int dcb200_ctx_set_coords_ex(dcb200_ctx* ctx, const float* coords, size_t n_rows, size_t n_cols, int on_device, int keep_order);
/* dev_perm: device uint32 [n_rows], frame index at every position */
int dcb200_ctx_order(dcb200_ctx* ctx, uint32_t* dev_perm);
/* dev_dst[a][frame] = dev_src[a][position] for n_arrays arrays of n_rows 32-bit values (device pointers) */
int dcb200_ctx_to_frame_order(dcb200_ctx* ctx, const uint32_t* dev_src, size_t n_arrays, uint32_t* dev_dst);

/* populations of positions [pos_begin,pos_end) against all frames.
 * dev_pops: device uint32 [n_radii][pos_end-pos_begin], position order.  Asynchronous on the context stream. */
int dcb200_ctx_populations(dcb200_ctx* ctx, const float* radii, size_t n_radii, size_t pos_begin, size_t pos_end,
                           uint32_t* dev_pops);
/* free energies from device populations (all n_rows, any order); max_pop == 0: computed on the device. */
int dcb200_ctx_free_energies(dcb200_ctx* ctx, const uint32_t* dev_pops, size_t n_rows, uint32_t max_pop,
                             float* dev_fe);

/* neighbour search, three steps so that shards can be exchanged between devices in between:
 *   prepare : ranks the frames by free energy on the device (dev_fe: device float [n_rows], FRAME order)
 *   scan    : positions [pos_begin,pos_end) against all frames; dev_keys_*: device uint64
 *             [pos_end-pos_begin] = (d2 bits << 32 | frame index), position order, ready for concatenation
 *   finish  : full key arrays [n_rows] in position order -> outputs in frame order (device arrays) */
int dcb200_ctx_nn_prepare(dcb200_ctx* ctx, const float* dev_fe);
int dcb200_ctx_nn_scan(dcb200_ctx* ctx, size_t pos_begin, size_t pos_end, uint64_t* dev_keys_nn,
                       uint64_t* dev_keys_hd);
int dcb200_ctx_nn_finish(dcb200_ctx* ctx, const uint64_t* dev_keys_nn, const uint64_t* dev_keys_hd,
                         uint32_t* dev_nn_idx, float* dev_nn_d2, uint32_t* dev_hd_idx, float* dev_hd_d2);

/* ---- shards: the row split of the multi-GPU drivers (replaces the per-GPU row ranges of
 * density_clustering_cuda.cu:152-181 and :286-328) ------------------------------------------------------------------
 * The positions are dealt to n_shards shards in blocks of 1024 positions, block b going to shard b % n_shards
 * (block-cyclic: the spatial order puts dense and sparse regions into long runs, so contiguous shards of equal length
 * cost very different amounts).  A shard keeps its blocks in order; dcb200_shard_capacity is the common padded number
 * of rows per shard, so that shard outputs [capacity]-strided concatenate with ONE all-gather:
 *   dcb200_ctx_populations_shard   dev_pops: uint32 [n_radii][capacity]
 *   dcb200_ctx_nn_scan_shard       dev_keys_*: uint64 [capacity]
 *   dcb200_ctx_shards_to_frame_order   gathered uint32 [n_shards][n_arrays][capacity] -> [n_arrays][n_rows], frame order
 *   dcb200_ctx_nn_finish_shards        gathered keys -> neighbour outputs in frame order; shard s's keys start at
 *                                      dev_keys_*[s * shard_stride]: shard_stride = capacity for two separately gathered
 *                                      arrays, 2 * capacity for ONE gathered [n_shards][2][capacity] block (nn, then hd)
 * Every context (one per GPU, same coordinates) builds the same deterministic order, so shard s means the same rows
 * everywhere.  Contexts on the GEMM-form path (17 <= n_cols <= 256) deal contiguous runs of `capacity` positions instead;
 * the two assembling functions undo whichever dealing the context uses. */
size_t dcb200_shard_capacity(size_t n_rows, int n_shards);
int dcb200_ctx_shard_rows(dcb200_ctx* ctx, int shard, int n_shards, size_t* rows);
int dcb200_ctx_populations_shard(dcb200_ctx* ctx, const float* radii, size_t n_radii, int shard, int n_shards,
                                 uint32_t* dev_pops);
int dcb200_ctx_nn_scan_shard(dcb200_ctx* ctx, int shard, int n_shards, uint64_t* dev_keys_nn, uint64_t* dev_keys_hd);
int dcb200_ctx_shards_to_frame_order(dcb200_ctx* ctx, const uint32_t* dev_src, size_t n_arrays, int n_shards,
                                     uint32_t* dev_dst);
int dcb200_ctx_nn_finish_shards(dcb200_ctx* ctx, const uint64_t* dev_keys_nn, const uint64_t* dev_keys_hd, int n_shards,
                                size_t shard_stride, uint32_t* dev_nn_idx, float* dev_nn_d2, uint32_t* dev_hd_idx,
                                float* dev_hd_d2);

/* screening on the device: the context's coordinates must be the free-energy-sorted frames, set with keep_order.
 * New rows [m_prev,m_new) restricted to [row_begin,row_end) are scanned against all lower positions;
 * dev_comp: device uint32 [m_new] union-find parents (in/out, see dcb200_screening_step).
 * dcb200_ctx_screening_flatten replaces every entry by its representative. */
int dcb200_ctx_screening_scan(dcb200_ctx* ctx, size_t m_prev, size_t m_new, size_t row_begin, size_t row_end,
                              float max_dist2, uint32_t* dev_comp);
int dcb200_ctx_screening_flatten(dcb200_ctx* ctx, size_t m_new, uint32_t* dev_comp);
/* unions another forest over the same positions (e.g. a peer GPU's result) into dev_comp */
int dcb200_ctx_screening_merge(dcb200_ctx* ctx, size_t m_new, uint32_t* dev_comp, const uint32_t* dev_other);

/* counters of the scans on this context since the last reset (for the benchmark / tests):
 * [0] kernels launched, [1] pairs handed to the slow path, [2] pairs re-evaluated in exact arithmetic,
 * [3] (warp, tile) scans (a consumer warp owns 128 rows), [4] pairs of the full row x column ranges requested, [5] pairs per (warp, tile) scan */
int dcb200_ctx_stats(dcb200_ctx* ctx, uint64_t stats[6], int reset);

/* diagnostics of the GEMM-form (tcgen05 tensor core) path used for 17 <= n_cols <= 256: *active = 1 if the context's current
 * coordinates are served by it; *check_ratio = max observed |fast value - exact d2| / proven error band over the pairs of the
 * populations calls made with DCB200_GEMM_CHECK=1 in the environment (0 if none); either pointer may be NULL */
int dcb200_ctx_gemm_info(dcb200_ctx* ctx, int* active, float* check_ratio);

/* diagnostics: tcgen05.mma kind::tf32 throughput of the device in TFLOP/s (operands resident in shared memory), the tensor
 * roofline denominator of the GEMM-form scans */
int dcb200_ctx_tf32_peak(dcb200_ctx* ctx, double ms_target, double* tflops);

/* diagnostics: FFMA-only throughput of the device in TFLOP/s (2 flop per FFMA), the FP32 roofline denominator */
int dcb200_ctx_ffma_peak(dcb200_ctx* ctx, double ms_target, double* tflops);

#ifdef __cplusplus
}
#endif
#endif /* DCB200_H */
