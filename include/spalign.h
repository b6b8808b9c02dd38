/*
 * spalign.h -- C ABI of libspalign_b200.so: the B200 (sm_100a) hot path of
 * pfnet-research/superpixel-align.
 *
 * The reference has no FFI: its seam is a set of module-level Python functions
 * (batch_spalign_kmeans.py:316-358, direct_clustering.py:204-208; imported by name at
 * utils/apply_spalign_kmeans.py:17-21).  The Python drop-ins in superpixel_align_b200/
 * keep those names and array contracts and call the entry points below through ctypes.
 * Each entry point cites the reference code it replaces (paths relative to the reference
 * repository root).
 *
 * Conventions
 *   - Every pointer is a DEVICE pointer owned by the caller unless marked "host".
 *   - The library allocates nothing persistent, never synchronises the device and is
 *     thread-safe per stream.  All work is enqueued on `stream` (a cudaStream_t).
 *   - Return value: 0 = enqueued OK, otherwise a spalign_status; spalign_last_error()
 *     gives a thread-local message.  Data-dependent conditions (label out of range, CSR
 *     capacity overflow, k-means stop reason) are reported in device-side words that the
 *     caller reads when it next synchronises.
 *   - Images are row-major [n_img, H, W]; feature cells are cell = cy * fw + cx with
 *     cy = min(y*fh/H, fh-1), cx = min(x*fw/W, fw-1)  (== cv2 INTER_NEAREST, the
 *     upsampling used at superpixel_overlaps.py:360-362).
 *   - "Rows" are superpixels of a batch, numbered sp_off[img] + label.
 */
#ifndef SPALIGN_H_
#define SPALIGN_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SPALIGN_ABI_VERSION 1

typedef void* spalign_stream_t; /* cudaStream_t */

enum spalign_status {
  SPALIGN_OK = 0,
  SPALIGN_E_INVALID = 1,     /* bad argument (shape, alignment, NULL) */
  SPALIGN_E_CUDA = 2,        /* CUDA runtime error while enqueueing */
  SPALIGN_E_WORKSPACE = 3,   /* workspace too small */
  SPALIGN_E_UNSUPPORTED = 4  /* valid but not implemented (e.g. K > 8) */
};

/* label / output element types */
#define SPALIGN_I32 0
#define SPALIGN_I64 1
#define SPALIGN_U8 2
/* k-means row element types */
#define SPALIGN_F32 0
#define SPALIGN_F64 1

/* bits of the device-side `flags` word written by spalign_overlap_csr */
#define SPALIGN_F_LABEL_RANGE 1  /* a pixel label was < 0 or >= n_sp of its image */
#define SPALIGN_F_NNZ_OVERFLOW 2 /* more (superpixel, cell) pairs than nnz_cap */
#define SPALIGN_F_EMPTY_ROW 4    /* a superpixel id owns no pixel (ids not contiguous) */

/* k-means stop reasons (status words) */
#define SPALIGN_KM_RUNNING (-1)
#define SPALIGN_KM_CONVERGED 0     /* all(new_assign == assign), batch_spalign_kmeans.py:158 */
#define SPALIGN_KM_EMPTY_CLUSTER 1 /* "Terminate KMeans iteration due to ...", :173-181 */
#define SPALIGN_KM_ITER_CAP 2      /* n_iter exhausted, :153 */
#define SPALIGN_KM_COMM_TIMEOUT 3  /* multi-GPU only: a peer never delivered its partial sums */

int spalign_abi_version(void);
const char* spalign_last_error(void);

/* ---- K0: per-image maximum label ------------------------------------------------------
 * Replaces len(np.unique(superpixel)) at batch_spalign_kmeans.py:321 for contiguous ids
 * (n_superpixels = max + 1).  max_out[n_img] int32, initialised by the call. */
int spalign_label_max(const void* labels, int label_dtype, int n_img, int H, int W,
                      int32_t* max_out, spalign_stream_t stream);

/* ---- K1: label map -> CSR overlap (pixel-count) matrix + per-superpixel statistics ----
 * Replaces the S full-image boolean masks of batch_spalign_kmeans.py:226-233 (membership),
 * :229 (center_of_mass), create_prior :111-129 (mean prior per superpixel) and the overlap
 * loop superpixel_overlaps.py:365-369.  The count matrix itself has no in-tree source
 * (notebooks/Efficient_Superpixel_Align.ipynb is a missing blob); contract: SURVEY 8 a1.
 *
 *   labels   [n_img,H,W] int32 or int64 (label_dtype)
 *   sp_off   [n_img+1] int64 device, row offset of each image; n_rows = sp_off[n_img] (host)
 *   gy, gx   [H], [W] float64 separable prior factors (NULL -> sum_prior not computed)
 *   indptr   [n_rows+1] int32, indices/counts [nnz_cap] int32: columns ascending per row
 *   area     [n_rows] int32; sum_y,sum_x [n_rows] int64 (exact); sum_prior [n_rows] float64
 *   nnz_flags  int64[4] device: [0] = nnz, [1] = OR of SPALIGN_F_* bits, [2] = largest number of
 *              pairs any one image produced (lower bound when pathological cells bypass the
 *              staging buffer), [3] reserved.  Each image may hold nnz_cap / n_img pairs.
 * Results are bit-reproducible run to run (no floating-point atomics). */
size_t spalign_overlap_workspace_bytes(int n_img, int H, int W, int fh, int fw,
                                       int64_t n_rows, int64_t nnz_cap);
int spalign_overlap_csr(const void* labels, int label_dtype, int n_img, int H, int W, int fh,
                        int fw, const int64_t* sp_off, int64_t n_rows, const double* gy,
                        const double* gx, int64_t nnz_cap, int32_t* indptr, int32_t* indices,
                        int32_t* counts, int32_t* area, int64_t* sum_y, int64_t* sum_x,
                        double* sum_prior, int64_t* nnz_flags, void* workspace,
                        size_t ws_bytes, spalign_stream_t stream);

/* ---- K2: superpixel-align pooling (CSR SpMM) ------------------------------------------
 * Replaces superpixel_align() batch_spalign_kmeans.py:210-276 / batch_superpixel_align
 * :316-330 under the count-pooling contract:
 *   out[r, :C] = sum_j counts[j] * feat[img(r), indices[j], :] / area[r]
 *   out[r, C:C+2] = (sum_y/area, sum_x/area) in image pixels when append_pos
 *   feat  [n_img, fh*fw, C] float32, CELL-MAJOR (PyTorch channels_last of [n,C,fh,fw])
 *   out   [n_rows, ld_out] float32, ld_out >= C + 2*append_pos, ld_out % 4 == 0
 * Summation order is ascending cell id (fixed), accumulation in fp32. */
int spalign_pool(const float* feat, int n_img, int C, int fh, int fw, const int64_t* sp_off,
                 int64_t n_rows, int max_rows_per_image, const int32_t* indptr,
                 const int32_t* indices, const int32_t* counts, const int32_t* area,
                 const int64_t* sum_y, const int64_t* sum_x, int append_pos, float* out,
                 int64_t ld_out, spalign_stream_t stream);

/* Same SpMM with float64 weights (bilinear overlap matrix, below). */
int spalign_pool_weighted(const float* feat, int n_img, int C, int fh, int fw,
                          const int64_t* sp_off, int64_t n_rows, int max_rows_per_image,
                          const int32_t* indptr, const int32_t* indices, const double* wvals,
                          const int32_t* area, const int64_t* sum_y, const int64_t* sum_x,
                          int append_pos, float* out, int64_t ld_out, spalign_stream_t stream);

/* ---- K1b: bilinear-weight overlap matrix ----------------------------------
 * Dense superpixel pooling of notebooks/Superpixel_Align.ipynb cell 4 (resize the feature
 * map to the image size with chainer.functions.resize_images, mean over each superpixel) as a
 * sparse matrix: W[r, c] = sum over the pixels p of superpixel r of the bilinear weight of
 * cell c at p (corner-aligned sampling u = linspace(0, n_in-1, n_out), lower neighbour
 * clip(floor(u), 0, n_in-2)).  The per-axis tables are computed by the caller (host, float64):
 *   iy0[H] lower neighbour row, wy0[H]/wy1[H] its weight and the upper neighbour's,
 *   ystart[fh+1] first y with iy0[y] >= cy; ix0, wx0, wx1, xstart likewise.
 * Output: CSR with ascending columns, float64 weights wvals[nnz_cap]; row_weight[n_rows]
 * (optional) = sum of a row's weights (= superpixel area up to rounding).  Bit-reproducible. */
size_t spalign_overlap_bilinear_workspace_bytes(int n_img, int H, int W, int fh, int fw,
                                                int64_t n_rows, int64_t nnz_cap);
int spalign_overlap_bilinear_csr(const void* labels, int label_dtype, int n_img, int H, int W,
                                 int fh, int fw, const int64_t* sp_off, int64_t n_rows,
                                 const int32_t* iy0, const double* wy0, const double* wy1,
                                 const int32_t* ystart, const int32_t* ix0, const double* wx0,
                                 const double* wx1, const int32_t* xstart, int64_t nnz_cap,
                                 int32_t* indptr, int32_t* indices, double* wvals,
                                 double* row_weight, int64_t* nnz_flags, void* workspace,
                                 size_t ws_bytes, spalign_stream_t stream);

/* ---- f2: anchor-sampled superpixel align (the reference's own pooling) -------------------
 * superpixel_align() batch_spalign_kmeans.py:226-274: n_select member pixels per superpixel
 * ("anchors"), each sampled bilinearly from the 4 nearest cell centres, averaged.
 * spalign_sample_anchors draws the anchors on the device (uniform over the members, distinct,
 * counter-based hash of (seed, row, draw); the reference's random.shuffle stream (:232) cannot
 * be replayed -- pass anchors computed elsewhere to reproduce it):
 *   anchors [n_rows, n_select, 2] int32 (y, x) pixels, -1 padded; n_valid [n_rows] int32
 * spalign_anchor_weights turns anchors into a CSR of 4 * n_select (cell, weight) entries per row
 * -- feature coordinates pixel * (fh / H) + 0.5 clipped to [0, f - 0.5] (:215, :235-240), 4
 * nearest centres with ties to the lower cell index (the reference's argsort is unstable there),
 * bounding-box corners and bilinear weights (:247-266), divided by n_valid -- for
 * spalign_pool_weighted (pass an all-ones `area`). */
int spalign_sample_anchors(const void* labels, int label_dtype, int n_img, int H, int W, int fh,
                           int fw, const int64_t* sp_off, int64_t n_rows, const int32_t* indptr,
                           const int32_t* indices, const int32_t* counts, const int32_t* area,
                           int n_select, uint64_t seed, int32_t* anchors, int32_t* n_valid,
                           spalign_stream_t stream);
int spalign_anchor_weights(const int32_t* anchors, const int32_t* n_valid, int64_t n_rows,
                           int n_select, int H, int fh, int fw, int32_t* indptr, int32_t* indices,
                           double* wvals, spalign_stream_t stream);

/* Layout helper: [n_img, C, ncell] (NCHW, what F.concat yields at batch_spalign_kmeans.py:435)
 * -> [n_img, ncell, C] cell-major.  direct_clustering.py:302 does the same transpose. */
int spalign_nchw_to_cellmajor(const float* src, float* dst, int n_img, int C, int ncell,
                              spalign_stream_t stream);

/* ---- K3: prior-weighted k-means --------------------------------------------------------
 * Replaces kmeans() batch_spalign_kmeans.py:136-183 (== direct_clustering.py:115-165,
 * superpixel_overlaps.py:121-171).  Semantics kept: unweighted init means (:150-151),
 * L2 distance + first-minimum argmin with NumPy NaN rules (:155-157), stop when the
 * assignment is unchanged (:158), cluster 0 averaged with w and the others with 1-w
 * (:163-171), stop on an empty cluster after adopting the new assignment (:173-181).
 * Distances and centroid sums are computed in float64.
 *
 *   X        [N, ldx] rows of float32/float64 (x_dtype); row stride ldx elements, rows
 *            16-byte aligned.  pos_mode 1 appends two virtual columns (x, y) = cell
 *            indices of row n (direct_clustering.py:297-303): x = (n % pos_period) % pos_w,
 *            y = (n % pos_period) / pos_w; D counts them.
 *   w        [N] float64 prior weights
 *   assign   [N] int32 in/out: initial assignment in, final assignment out
 *   group_off[G+1] int64 device: G independent problems over contiguous row ranges
 *   centers  [G, K, D] float64 out (may be NULL for spalign_kmeans_groups)
 *   iters / status [G] int32 out
 */
size_t spalign_kmeans_groups_workspace_bytes(int D, int K, int G);
/* one persistent CTA per group; the whole iteration loop runs on the device */
int spalign_kmeans_groups(const void* X, int x_dtype, int64_t ldx, int pos_mode, int pos_w,
                          int64_t pos_period, const double* w, int D, int K, int n_iter,
                          const int64_t* group_off, int G, int32_t* assign, double* centers,
                          int32_t* iters, int32_t* status, void* workspace, size_t ws_bytes,
                          spalign_stream_t stream);

/* multi-CTA building blocks for large groups and for the multi-GPU global clustering:
 *   chunks [n_chunks,4] int64 device, in launch order: (group, row_begin, row_end, slot);
 *            slot = position of the chunk in `partials`; the slots of a group are contiguous
 *            and ordered by row_begin
 *   group_chunk_off [G+1] int32 device: slot range of every group
 *   partials [n_chunks, K*(D+2)+1] float64: per chunk and cluster k the D sums of omega*x,
 *            then sum(omega) and the member count ([K][D+2]); last element = #rows changed
 *   totals   [G, K*(D+2)+1] float64: fixed-order sum over the group's chunks
 *   mode 0: accumulate init means for the given assignment (omega = 1, no reassignment)
 *   mode 1: reassign rows against `centers`, count changes, accumulate with prior weights
 * A multi-GPU caller all-reduces `totals` (NCCL, float64 sum) between reduce and update. */
int spalign_kmeans_sweep(const void* X, int x_dtype, int64_t ldx, int pos_mode, int pos_w,
                         int64_t pos_period, int64_t pos_row0, const double* w, int D, int K,
                         const int64_t* chunks, int n_chunks, const double* centers, int mode,
                         int32_t* assign, const int32_t* status, double* partials,
                         spalign_stream_t stream);
/* sweep + reduce + update in ONE launch (single-GPU): the chunk of a group that finishes last
 * (arrival ticket in counters[G], zero on entry and on exit) sums the group's partials in
 * chunk order and applies the update, so the result is bit-identical to the three-call form.
 * mode 2 (after one mode-1 call): `totals` is kept as RUNNING sums and only rows whose
 * assignment changed are subtracted from their old cluster and added to the new one (float64,
 * row order within a chunk, chunks in fixed order) -- the same sums up to float64 rounding,
 * without re-reading unchanged rows in phase 2.
 * ub, lb (optional, float[N]) and cdelta (double[G,K]): Hamerly bounds.  Mode 1 stores, per row,
 * an upper bound on the distance to its centre and a lower bound on the distance to any other
 * centre; every update records how far each centre moved; a mode-2 sweep shifts the bounds by
 * that drift and only gathers, screens and re-bounds the rows whose bounds no longer prove that
 * the assignment is unchanged (fp32 rows, chunks of <= 1024 rows).  Results are identical. */
/* counters: int32[n_chunks + G], zero on entry and on exit (tickets of the first-level reducers,
 * indexed by chunk slot, then one per group).  The partials of a group are summed along a fixed
 * tree -- runs of 32 consecutive slots in slot order, then the run sums in order -- by the chunks
 * that finish last, so the serial tail of a group of n chunks reads n/32 + 32 vectors, not n. */
int spalign_kmeans_iterate(const void* X, int x_dtype, int64_t ldx, int pos_mode, int pos_w,
                           int64_t pos_period, int64_t pos_row0, const double* w, int D, int K,
                           const int64_t* chunks, int n_chunks, const int32_t* group_chunk_off,
                           int mode, int n_iter, int32_t* assign, double* partials,
                           double* totals, double* centers, int32_t* iters, int32_t* status,
                           int32_t* counters, float* ub, float* lb,
                           double* cdelta, spalign_stream_t stream);
/* Runs every group that is still SPALIGN_KM_RUNNING to its stop condition in ONE launch: one
 * persistent CTA per group repeats mode-2 iterations (bounds pass, gather + screen the rows the
 * bounds cannot prove stable, move the changed rows between the running sums, new centres,
 * drift) without leaving the SM.  Preconditions: fp32 rows; spalign_kmeans_iterate has run
 * mode 0 and at least one mode-1 iteration with ub/lb/cdelta, so totals/centers/cdelta/ub/lb
 * are current.  group_off: device int64[G+1] row ranges.  Same stop rules and results as
 * repeating spalign_kmeans_iterate (batch_spalign_kmeans.py:152-179); meant for many groups
 * of at most a few thousand rows each (per-image clustering), where it replaces one launch
 * and one host poll per iteration.
 * slice_iters > 0: every group runs at most that many iterations in this launch and is left
 * SPALIGN_KM_RUNNING if it has not stopped (all state is in the global arrays; a later call
 * continues).  rows_per_set: rows each warp of the sparse sweep takes at a time -- 2: two CTAs
 * per SM (all of up to 2 x #SM groups resident, each slower), 4 (= 0, default): one CTA per SM,
 * ~1.5x faster per group (24 vs 37 us per iteration of a slow image; the single 4-row launch
 * is also the faster schedule at 300 groups on 148 SMs). */
int spalign_kmeans_finish(const void* X, int x_dtype, int64_t ldx, int pos_mode, int pos_w,
                          int64_t pos_period, int64_t pos_row0, const double* w, int D, int K,
                          const int64_t* group_off, int G, int n_iter, int32_t* assign,
                          double* totals, double* centers, int32_t* iters, int32_t* status,
                          float* ub, float* lb, double* cdelta, int slice_iters, int rows_per_set,
                          spalign_stream_t stream);
int spalign_kmeans_reduce(const double* partials, const int32_t* group_chunk_off, int G, int D,
                          int K, double* totals, spalign_stream_t stream);
int spalign_kmeans_update(const double* totals, int G, int D, int K, int mode, int n_iter,
                          double* centers, int32_t* iters, int32_t* status,
                          spalign_stream_t stream);

/* ---- multi-GPU dataset-wide clustering (BASELINE configs[4]; build-defined: the reference has
 * no collective on this path, direct_clustering.py:297-317 clusters one batch) -----------------
 * Semantics: kmeans() of batch_spalign_kmeans.py:136-183 on the concatenation of all ranks'
 * rows, rank r holding a contiguous slice.  One process per GPU.  A communicator owns one
 * exchange buffer per rank (cudaMalloc), mapped into every peer process through CUDA IPC:
 *   spalign_comm_create   allocate this rank's buffer for vectors of up to pv_cap doubles
 *                         (pv = K*(D+2)+1); synchronises
 *   spalign_comm_handle   64-byte IPC handle of the buffer (host memory); the caller gathers the
 *                         handles of all ranks (torch.distributed / MPI / files)
 *   spalign_comm_connect  map the peers' buffers; handles = world * 64 bytes, rank order (host)
 * spalign_kmeans_iterate_dist = spalign_kmeans_iterate for G = 1 with the exchange inside the
 * kernel: the CTA that finishes the local reduction stores this rank's K*(D+2)+1 sums into every
 * rank's inbox over NVLink (plain stores + release flags at system scope), waits for the peers'
 * flags and adds the world's vectors in rank order, so totals, centres and stop flags are
 * bit-identical on all ranks and an iteration is ONE launch per GPU with no NCCL call and no host
 * round trip.  Every rank must issue the same sequence of calls; launches after the stop
 * condition exit at once (ranks may over-enqueue).  A peer that never answers (3 s) stops the
 * group with SPALIGN_KM_COMM_TIMEOUT; the communicator is then unusable. */
#define SPALIGN_COMM_HANDLE_BYTES 64
typedef struct spalign_comm spalign_comm_t;
int spalign_comm_create(int world, int rank, int64_t pv_cap, spalign_comm_t** out);
int spalign_comm_handle(spalign_comm_t* comm, void* handle_out /* host, 64 bytes */);
int spalign_comm_connect(spalign_comm_t* comm, const void* handles /* host, world*64 bytes */);
int spalign_comm_destroy(spalign_comm_t* comm);
int spalign_kmeans_iterate_dist(const void* X, int x_dtype, int64_t ldx, int pos_mode, int pos_w,
                                int64_t pos_period, int64_t pos_row0, const double* w, int D,
                                int K, const int64_t* chunks, int n_chunks,
                                const int32_t* group_chunk_off, int mode, int n_iter,
                                int32_t* assign, double* partials, double* totals,
                                double* centers, int32_t* iters, int32_t* status,
                                int32_t* counters, float* ub, float* lb, double* cdelta,
                                spalign_comm_t* comm, spalign_stream_t stream);

/* Diagnostics, synchronises the device: out_host[0] = rows screened in fp32 since the last
 * reset, out_host[1] = rows that needed the exact float64 pass (host int64[2]). */
int spalign_kmeans_debug_stats(int64_t* out_host, int reset);

/* Device-side seeded init for many small groups (batch_spalign_kmeans.py:141-149):
 * thr = sort(w_g)[N_g/2]; rows with w > thr -> 0; the others take shuffled[g_shuf_off + i]
 * in row order.  `shuffled` holds, per group, arange(m) % (K-1) + 1 already shuffled by the
 * host with the NumPy legacy stream for the expected m = N_g/2 + 1; status_m[g] receives
 * the actual m so the host can detect a tie-induced mismatch.  Groups of up to 4096 rows sort
 * their weights in shared memory; larger ones (joint clustering of 30 images, cell clustering)
 * find the median with an 8-pass radix select over global memory -- any group size. */
int spalign_kmeans_init(const double* w, const int64_t* group_off, int G,
                        const int32_t* shuffled, const int64_t* shuf_off, int32_t* assign,
                        int32_t* m_out, spalign_stream_t stream);

/* ---- f3: SLIC superpixels ---------------------------------------------------------------
 * Replaces batch_superpixel() batch_spalign_kmeans.py:299-313 for --superpixel_method slic
 * (skimage.segmentation.slic(img.transpose(1, 2, 0), n_segments), one image at a time on the CPU).
 * scikit-image 0.13.1 is not in the reference tree: the contract is the published algorithm of
 * that version as restated in oracle/spalign_oracle.py:slic (regular seed grid, k-means in
 * (y, x, L, a, b) over 2*step windows, lower id wins ties, max_iter, 4-connectivity enforcement
 * with min_size = min_size_factor * H*W / n_segments); colours quantised to 2^-12 and integer
 * centre sums make the result bit-reproducible.  PARITY UNPINNED by the reference.
 *   images  [n_img, 3, H, W] float32 (CHW, the layout the reference holds; values in 0..1)
 *   labels  [n_img, H, W] int32 out: contiguous ids 0..S-1 in raster order of first pixels
 *   n_labels[n_img] int32 out: S per image
 * spalign_slic_segments = number of seeds of the grid (the cluster count before connectivity). */
int spalign_slic_segments(int H, int W, int n_segments);
size_t spalign_slic_workspace_bytes(int n_img, int H, int W, int n_segments);
int spalign_slic(const float* images, int n_img, int H, int W, int n_segments, double compactness,
                 int max_iter, int convert2lab, int enforce_connectivity, double min_size_factor,
                 int32_t* labels, int32_t* n_labels, void* workspace, size_t ws_bytes,
                 spalign_stream_t stream);

/* ---- f3, second branch: Felzenszwalb-Huttenlocher superpixels ---------------------------------
 * Replaces batch_superpixel() batch_spalign_kmeans.py:301-307, the reference's DEFAULT
 * --superpixel_method (skimage.segmentation.felzenszwalb(img.transpose(1, 2, 0) / 255.,
 * scale=300, sigma=0.8, min_size=20), one image at a time on the CPU).  Contract: the algorithm of
 * scikit-image 0.13 as restated in oracle/spalign_oracle.py:felzenszwalb -- Gaussian smoothing
 * (scipy.ndimage.gaussian_filter, reflect), 8-connectivity edge costs, edges in ascending cost
 * (equal costs in edge order: the one point skimage leaves to np.argsort), greedy merges while
 * cost < min(Int(a) + k/|a|, Int(b) + k/|b|) with k = scale / 255, then the min_size pass, labels
 * numbered by ascending union-find root.  Float64 in the oracle's operation order: bit-identical
 * labels.  PARITY UNPINNED by the reference (scikit-image is not in its tree).
 *   images  [n_img, 3, H, W] float32 (CHW; values in 0..1)
 *   labels  [n_img, H, W] int32 out, ids 0..S-1; n_labels[n_img] int32 out
 * The merge pass is sequential by definition; images are processed side by side (one CTA each). */
size_t spalign_felzenszwalb_workspace_bytes(int n_img, int H, int W);
int spalign_felzenszwalb(const float* images, int n_img, int H, int W, double scale, double sigma,
                         int min_size, int32_t* labels, int32_t* n_labels, void* workspace,
                         size_t ws_bytes, spalign_stream_t stream);

/* ---- K4: paint-back -------------------------------------------------------------------
 * Replaces the double loop of weighted_kmeans() batch_spalign_kmeans.py:193-199 and the
 * `== 0` road mask (:207): cluster_map[p] = table[sp_off[img] + label[p]] (0 when the label
 * is outside [0, n_sp)), road_mask[p] = (cluster_map[p] == road_value).
 *   cluster_map: element type out_dtype (SPALIGN_U8 / I32 / I64), may be NULL
 *   road_mask:   uint8 0/1, may be NULL */
int spalign_paint(const void* labels, int label_dtype, int n_img, int H, int W,
                  const int64_t* sp_off, const int32_t* table, void* cluster_map, int out_dtype,
                  uint8_t* road_mask, int road_value, spalign_stream_t stream);

/* ---- K5: overlap refine ---------------------------------------------------------------
 * Replaces superpixel_overlaps.py:359-369: overlap[r] = sum_j counts[j]*road_cell[indices[j]],
 * road_px[img] = sum of overlap over the image's rows, keep[r] = road_px > 0 and
 * overlap/road_px > thr (float64 compare).  keep is int32 so it can be painted with K4. */
int spalign_refine(const int64_t* sp_off, int n_img, int64_t n_rows, int ncell,
                   int max_rows_per_image, const int32_t* indptr, const int32_t* indices, const int32_t* counts,
                   const uint8_t* road_cell, double thr, int64_t* overlap, int64_t* road_px,
                   int32_t* keep, spalign_stream_t stream);

/* Nearest-neighbour resize of uint8 maps [n_img, h, w] -> [n_img, H, W]: the
 * cv.resize(road_mask / clustering_result, (w, h), interpolation=cv.INTER_NEAREST) that brings the
 * estimates to the label shape before evaluation (batch_spalign_kmeans.py:470-477,
 * direct_clustering.py:329-332, superpixel_overlaps.py:360-362). */
int spalign_resize_nearest_u8(const uint8_t* src, int n_img, int h, int w, uint8_t* dst, int H,
                              int W, spalign_stream_t stream);

/* ---- evaluation -----------------------------------------------------------------------
 * Replaces chainercv calc_semantic_segmentation_confusion as used at
 * batch_spalign_kmeans.py:398-402 for 2 classes: conf[img, gt*2+pred] over pixels gt >= 0. */
int spalign_confusion2(const uint8_t* pred, const int32_t* gt, int n_img, int64_t n_pix,
                       int64_t* conf, spalign_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* SPALIGN_H_ */
