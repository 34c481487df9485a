/* qb200.h - C ABI of the B200-native LBG codebook-training / index-assignment path.
 *
 * This is the drop-in boundary for the hot path of coodie/quant: everything the reference does
 * between "RGB bytes of an image" and "(codebook, per-block indices, distortion)" - i.e.
 *   getBlocksAsVectorsFromImage        /root/reference/src/Compressor.cpp:31-62
 *   LBGQuantizer::quantize             /root/reference/src/Quantizer.cpp:121-143
 *     Solution::assignCodeVectors      /root/reference/src/Quantizer.cpp:24-32  (+ src/KDTree.cpp)
 *     Solution::updateDistortion       /root/reference/src/Quantizer.cpp:9-22
 *     Solution::fixCodeVectors         /root/reference/src/Quantizer.cpp:72-87
 *     split step                       /root/reference/src/Quantizer.cpp:134-138
 * runs behind these entry points on one B200 (sm_100a).  Plain pointers and sizes only; no C++
 * or torch types; no exceptions cross this boundary.  The caller owns every host buffer; the
 * library owns device memory inside an opaque context.  A context is not re-entrant: use one per
 * host thread / per GPU.  There is NO CPU fallback: without a usable CUDA device qb200_create
 * fails with QB200_ERR_NODEV.
 *
 * Multi-GPU: the vectors are sharded by contiguous bands of block rows; the only exchange is one
 * sum all-reduce of K*(dim+2) 64-bit integers per split level, done by the library itself over NVLink peer memory
 * (qb200_comm.cu: every rank reads every rank's words directly - no host round trip, no third-party library):
 *   - ONE process driving several GPUs: qb200_create_multi gives a context that splits qb200_set_image over its
 *     devices and runs qb200_train on all of them (one host thread and one stream per device);
 *   - one process per GPU: every process creates its context, qb200_comm_export / qb200_comm_attach join them
 *     through CUDA IPC handles (exchanged by whatever launcher the caller has), then each sets its own shard
 *     (qb200_set_image_shard / _band) and calls qb200_train with allreduce == NULL;
 *   - or the caller supplies the all-reduce as a callback (qb200_allreduce_fn), e.g. ncclAllReduce.
 */
#ifndef QB200_H
#define QB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QB200_VERSION 200 /* major*100 + minor */

/* Status codes (0 = success, negative = failure; qb200_last_error gives the text). */
#define QB200_OK 0
#define QB200_ERR_ARG (-1)    /* bad argument (null pointer, empty training set, nbits out of range...) */
#define QB200_ERR_NODEV (-2)  /* no CUDA device / device is not sm_100 */
#define QB200_ERR_CUDA (-3)   /* a CUDA runtime call or kernel failed */
#define QB200_ERR_OOM (-4)    /* device or pinned-host allocation failed */
#define QB200_ERR_STATE (-5)  /* call sequence error (e.g. train before set_image) */
#define QB200_ERR_COMM (-6)   /* the all-reduce callback reported failure */

/* Colour spaces: same integer values as the reference's `enum class ColorSpaces`
 * (/root/reference/include/ColorSpace.hpp:6).  NORMAL and SCALED vectors live on the byte lattice (integer
 * statistics, tensor-core filter, multi-GPU); CIE1931 vectors do not and take the FP64-vector path
 * (qb200_set_vectors_f64 below). */
#define QB200_CS_NORMAL 0 /* value = (double)(int8)byte        src/ColorSpace.cpp:4-6   */
#define QB200_CS_SCALED 1 /* value = ((int8)byte + 128.0)/255  src/ColorSpace.cpp:16-21 */
#define QB200_CS_CIE1931 2 /* 3x3 matrix / 0.17697 per pixel     src/ColorSpace.cpp:31-48; FP64 vectors, see below */

/* Training schedule. */
#define QB200_MODE_PARITY 0 /* the reference's HEAD schedule: ONE assignment per split level,
                               no empty-cell repair (src/Quantizer.cpp:98-108). */
/* Extensions - what README.md:29-31 describes but HEAD does not do; there is no reference behaviour
 * to match beyond that text, so PARITY IS UNPINNED for both:          */
#define QB200_MODE_FULL 1        /* per split level re-assign and re-fix until the relative change of the
                                    distortion is <= eps (at most 100 iterations, src/Quantizer.cpp:101) */
#define QB200_MODE_FULL_REPAIR 2 /* FULL + empty-cell repair: an empty cell takes a pseudo-random member of
                                    the cell with the largest distortion (qb200_set_seed)                 */

typedef struct qb200_ctx qb200_ctx;

/* Per-level report filled by qb200_train (all optional diagnostics). */
typedef struct qb200_level_report {
  uint32_t K;             /* codebook size of this level */
  uint32_t flagged;       /* queries whose FP32 top-2 gap was inside the error margin (this rank): re-solved in FP64 */
  uint32_t changed;       /* of those, how many the exact FP64 resolver moved to another index */
  uint32_t ties;          /* of those, how many had (near-)exact FP64 ties and took the KD-tree walk */
  uint32_t dead_cells;    /* cells with no member after the (global) reduction */
  uint32_t kd_depth;      /* depth of the nanoflann-order KD tree built for the resolver */
  uint32_t iterations;    /* assignment passes of this level (1 in QB200_MODE_PARITY) */
  uint32_t repaired;      /* empty cells re-seeded at this level (QB200_MODE_FULL_REPAIR) */
  float ms_assign;        /* device time of the filter (codebook staging + tensor-core or CUDA-core kernel [+ finalise]) */
  float ms_resolve;       /* device time of the exact resolver kernels (brute force + tree walk) */
  float ms_accumulate;    /* device time of the separate per-cell statistics kernel (~0 when the filter accumulates) */
  uint32_t refiltered;    /* tensor-core levels: queries inside the tensor-core margin, re-ranked in FP32 (of which
                             `flagged` were still undecided and went to the FP64 resolver); 0 on the other levels */
  double distortion_pre;  /* updateDistortion() before fixCodeVectors (src/Quantizer.cpp:100) */
  double distortion_post; /* updateDistortion() after fixCodeVectors  (src/Quantizer.cpp:104) */
  uint32_t sensitive;     /* of `flagged`: decisions with a second codevector within 2^-44 (relative to the largest
                             distance) of the minimum - exact ties, tree-order ties: the only ones the last bits of
                             the codebook can change */
  uint32_t reserved;
} qb200_level_report;

/* In-place sum all-reduce over `count` unsigned 64-bit integers at device address `dev_u64`.
 * STREAM-ORDERED: when it is called the work that produces the data has been enqueued on
 * `cuda_stream` but has not necessarily run; the callback must enqueue the reduction on that stream
 * (ncclAllReduce(..., (cudaStream_t)cuda_stream) does exactly this) - or synchronise the stream, reduce
 * by other means and return - so that later work on `cuda_stream` sees the reduced values.  The library
 * does not synchronise around the call (the train has no host round trip between split levels).
 * Two's-complement wrap-around makes the same call correct for the signed sums.  Return 0 on success. */
typedef int (*qb200_allreduce_fn)(void *dev_u64, size_t count, void *cuda_stream, void *user);

/* ---- lifetime ---------------------------------------------------------------------------- */
int qb200_version(void);
/* Creates a context on CUDA device `device`.  Replaces nothing in the reference (it has no
 * device state); it is the analogue of constructing `Solution` (src/Quantizer.cpp:89-96). */
int qb200_create(int device, qb200_ctx **out);
/* One context over `ndev` devices of this process (dev_ids == NULL: devices 0..ndev-1; ndev <= 0: every visible
 * device; at most 16; all pairs need peer access).  It accepts qb200_set_image (ONE host image: device r takes block rows
 * [wB*r/ndev, wB*(r+1)/ndev)), qb200_set_vectors_u8/_f64 (contiguous ranges), qb200_train (allreduce must be NULL:
 * the devices reduce among themselves), qb200_get_assign[_u64], qb200_decode, qb200_num_vectors/_dim, the qb200_set_*
 * switches and qb200_destroy; every other call returns QB200_ERR_STATE.  Results are bit-identical to a single-device
 * context's (integer sums; the exact-centroid chains continue from device to device in vector order).  This is what
 * CompressedImage::compress uses when QB200_DEVICES is set (quant_b200/host). */
int qb200_create_multi(int ndev, const int *dev_ids, qb200_ctx **out);
void qb200_destroy(qb200_ctx *ctx);
/* Last error text of this context (or of the failed qb200_create when ctx == NULL). */
const char *qb200_last_error(const qb200_ctx *ctx);
/* Run all device work on an existing CUDA stream (cudaStream_t); NULL restores the context's own. */
int qb200_set_stream(qb200_ctx *ctx, void *cuda_stream);
/* Filter engine for codebooks of 128 or more entries (QB200_TC_MIN_K): 1 (default) = tcgen05 tensor-core kernel,
 * 0 = FP32 CUDA-core kernel.  Results are identical either way (both only filter; flagged queries
 * are re-solved exactly); the switch exists for measurement and for the parity tests.
 * The environment variable QB200_DISABLE_TC=1 forces 0 process-wide. */
int qb200_set_tensor_cores(qb200_ctx *ctx, int enable);
int qb200_device_info(const qb200_ctx *ctx, int *sm_count, int *cc_major, int *cc_minor,
                      size_t *total_mem);

/* Centroid arithmetic of SCALED lattice vectors (NORMAL sums are integers and always exact; FP64 vectors always use
 * the compensated sums).  The reference divides a COMPENSATED FP64 sum of the members, taken in ascending vector
 * order (Solution::sumInArea / trainingSetSum, src/Quantizer.cpp:46-70), by their count.  Modes:
 *   0  integer sums: centroid = ((double)S_t / 255) / n, within 4e-16 relative of the reference's but not always
 *      equal in the last bit.  Every assignment pass is still bit-identical given the same codebook, but on inputs
 *      with duplicated vectors (palettes, flat areas) that last bit can decide exact ties of the next split level.
 *   1  the reference's sums themselves, evaluated in parallel (qb200_exact_fast.cuh: exact integer model of the
 *      compensated loop, anchored segments, speculative residue classes, chained summaries): codebooks bit-identical
 *      to the reference's on any input.
 *   2  the same sums as one literal sequential chain per (cell, dimension) - same bits, ~2 N dependent steps per
 *      train; kept for comparison.
 *   3  AUTO (default): train with mode 0 while counting the decisions that hinged on (near-)ties of several
 *      codevectors (qb200_level_report.sensitive).  If there were none on any rank the result is index-identical to
 *      the reference's and is returned; otherwise the train is repeated with mode 1.  End-to-end identical to the
 *      reference on every input, at the fast path's cost whenever the input allows it.  (With a caller-supplied
 *      all-reduce callback and no qb200_set_rank the chains cannot be continued across ranks: mode 0 is used.)
 * The environment variable QB200_EXACT_CENTROIDS=0|1|2|auto sets the mode of every new context.  Sharded runs
 * continue the chains from rank to rank: qb200_set_rank (or an attached group) is required and rank order must be
 * vector order (rank r owns lower vector indices than rank r+1). */
int qb200_set_exact_centroids(qb200_ctx *ctx, int mode);
/* 1 when the last qb200_train on this context used (or, in auto mode, had to repeat itself with) the compensated sums. */
int qb200_last_train_exact(const qb200_ctx *ctx);
/* Seed of the empty-cell repair's member choice (default 0x5eed).  QB200_MODE_FULL_REPAIR only. */
int qb200_set_seed(qb200_ctx *ctx, uint64_t seed);
/* This context's rank among `world` contexts that train one sharded set together.  Needed by
 * QB200_MODE_FULL_REPAIR (the ranks agree on the chosen members through the sum all-reduce) and by
 * qb200_set_exact_centroids. */
int qb200_set_rank(qb200_ctx *ctx, int rank, int world);

/* ---- one process per GPU: joining the contexts into an all-reduce group --------------------------
 * qb200_comm_export allocates this context's exchange block (room for max_words 64-bit words per round, 0 = default
 * 2^18; larger payloads take several rounds) and writes its CUDA IPC handle (QB200_COMM_HANDLE_BYTES bytes) to
 * handle_out.  After the caller has gathered all ranks' handles (rank order), qb200_comm_attach maps the peers' blocks
 * and makes this context rank `rank` of `world` (it implies qb200_set_rank).  From then on qb200_train with
 * allreduce == NULL reduces over the group.  Every rank must make the same sequence of training calls. */
#define QB200_COMM_HANDLE_BYTES 64
int qb200_comm_export(qb200_ctx *ctx, size_t max_words, void *handle_out);
int qb200_comm_attach(qb200_ctx *ctx, int world, int rank, const void *handles /* world * QB200_COMM_HANDLE_BYTES */);
/* The group's sum all-reduce itself, in place on `count` 64-bit words in device memory, stream-ordered on the
 * context's stream (diagnostics, tests, callers with their own reductions). */
int qb200_allreduce_u64(qb200_ctx *ctx, void *dev_u64, size_t count);

/* ---- training set ------------------------------------------------------------------------
 * Replaces getBlocksAsVectorsFromImage (src/Compressor.cpp:31-62).  The N x dim double vectors
 * are never materialised: kernels gather the raw bytes with the reference's layout rule (pixel
 * index x*ySize + y, vector index i*hBlocks + j, y-overflow wraps, past-the-end elements are 0.0).
 * `rgb` holds n_images consecutive images of xSize*ySize*3 bytes each (n_images >= 1); the
 * training set is the concatenation of their block vectors.  rgb_is_device != 0 means `rgb` is a
 * device pointer that stays valid until the next set_image/destroy (no copy is made). */
int qb200_set_image(qb200_ctx *ctx, const uint8_t *rgb, int xSize, int ySize, int blockWidth,
                    int blockHeight, int colorspace, int n_images, int rgb_is_device);
/* Same for one image, but this context only owns block rows [row_begin, row_end) of the
 * wBlocks = ceil(xSize/blockWidth) rows (multi-GPU sharding).  `rgb` is the WHOLE
 * host image; only the byte range the shard needs is copied to the device. */
int qb200_set_image_shard(qb200_ctx *ctx, const uint8_t *rgb, int xSize, int ySize, int blockWidth,
                          int blockHeight, int colorspace, size_t row_begin, size_t row_end);
/* Same shard, but the caller passes ONLY the shard's bytes: band[0] is image byte
 * row_begin*blockWidth*ySize*3 and band_len covers the shard (a rank that loaded just its own band
 * of a large image).  band_is_device != 0: `band` is a device pointer that is borrowed, not copied. */
int qb200_set_image_band(qb200_ctx *ctx, const uint8_t *band, size_t band_len, int band_is_device,
                         int xSize, int ySize, int blockWidth, int blockHeight, int colorspace,
                         size_t row_begin, size_t row_end);
/* Training set given as an N x dim matrix of lattice bytes (row-major): element value is
 * decoded with `colorspace` exactly like an image byte.  Used by the generic
 * AbstractQuantizer::quantize(vector<Vector>) entry (include/Quantizer.hpp:12-14). */
int qb200_set_vectors_u8(qb200_ctx *ctx, const uint8_t *bytes, size_t n_vectors, int dim,
                         int colorspace, int bytes_is_device);
/* Training set given as an N x dim matrix of doubles (row-major): the fully general form of
 * AbstractQuantizer::quantize(vector<Vector>) (include/Quantizer.hpp:12-14).  x_is_device != 0: a device pointer
 * that stays valid until the next set call (no copy).
 * General FP64 vectors - this call, and images in QB200_CS_CIE1931 (qb200_set_image converts them to doubles on
 * the device) - have no integer statistics: the assignment keeps the filter + exact re-check (the filter works on
 * the values rounded to FP32, its margin covers that), centroids are the reference's compensated FP64 sums executed
 * in vector order (as with qb200_set_exact_centroids, so codebooks are bit-identical to the reference's) and the
 * distortions are FP64 sums (reference: OpenMP reduction, order unspecified; 1e-6 relative).  Modes
 * QB200_MODE_PARITY and QB200_MODE_FULL.  Sharded over several contexts (each holding a contiguous range of the
 * vectors, or qb200_set_image_shard / _band for CIE1931) it needs qb200_set_rank with rank order = vector order:
 * the ranks continue each other's sums in turn and exchange their distortion partials as one slot per rank
 * through the same u64 sum all-reduce, so every rank gets the single-GPU bits. */
int qb200_set_vectors_f64(qb200_ctx *ctx, const double *x, size_t n_vectors, int dim, int x_is_device);
/* Number of vectors this context holds / their dimension. */
size_t qb200_num_vectors(const qb200_ctx *ctx);
int qb200_dim(const qb200_ctx *ctx);

/* ---- the hot path -------------------------------------------------------------------------
 * qb200_train replaces LBGQuantizer::quantize (src/Quantizer.cpp:121-143): initial codevector =
 * mean, then nbits levels of {split by (1 + 0.2) / (1 - 0.2), assign, accumulate, fix}.
 *   n_total       total number of vectors over all ranks (0 = this context's own count)
 *   allreduce     NULL for a single GPU
 *   codebook_out  K*dim doubles (K = 1 << nbits), colour-space domain, as quantize() returns
 *   distortion_out  last updateDistortion() value (src/Quantizer.cpp:142)
 *   reports       nbits entries or NULL
 * The assignment of the LAST level (w.r.t. that level's pre-fix codebook, src/Quantizer.cpp:142)
 * stays on the device; fetch it with qb200_get_assign. */
int qb200_train(qb200_ctx *ctx, int nbits, double eps, int mode, uint64_t n_total,
                qb200_allreduce_fn allreduce, void *allreduce_user, double *codebook_out,
                double *distortion_out, qb200_level_report *reports);
/* Copies this context's assignment (one uint32 per vector) to host memory. */
int qb200_get_assign(qb200_ctx *ctx, uint32_t *assign_out);
/* Same, widened to the reference's std::vector<size_t> element type. */
int qb200_get_assign_u64(qb200_ctx *ctx, uint64_t *assign_out);
/* The assignment as a bit stream, `bits` bits per index (1..32; every index must fit), index i in stream bits
 * [i*bits, (i+1)*bits), bit p of the stream = bit p%8 of byte p/8 (LSB first): the packed form the reference's
 * sizeInBits() counts (src/Compressor.cpp:174-182) but its file never stores (README "possible improvements").
 * Packed on the device, so only ceil(n*bits/8) bytes cross PCIe.  out_bytes must be at least that.  Extension:
 * the packed .quant container of quant_b200/host (CompressedImage::saveToFilePacked) uses it. */
int qb200_get_assign_packed(qb200_ctx *ctx, int bits, uint8_t *out, size_t out_bytes);
/* Device address of the assignment array (uint32[num_vectors]); valid until the next call that
 * changes the training set. */
int qb200_assign_device_ptr(qb200_ctx *ctx, void **dev_ptr);

/* One level's worth of work against a caller-supplied codebook (K x dim doubles, colour-space
 * domain): Solution::assignCodeVectors (src/Quantizer.cpp:24-32, i.e. KDTree(dim, cb) +
 * nearestNeighbour per vector, src/KDTree.cpp:16-29) followed by the integer statistics
 * fixCodeVectors/updateDistortion are derived from.  Any output pointer may be NULL.
 *   assign_out  num_vectors uint32
 *   count_out   K uint64          members per cell
 *   sum_out     K*dim int64       per-cell sums of the lattice value L = (int8)byte
 *   sqsum_out   K uint64          per-cell sums over members and dims of L*L
 *   flagged_out queries that went through the exact FP64 resolver
 * This is also the encode-only entry (BASELINE config 5): pass only assign_out. */
int qb200_assign_accumulate(qb200_ctx *ctx, const double *codebook, uint32_t K, uint32_t *assign_out,
                            uint64_t *count_out, int64_t *sum_out, uint64_t *sqsum_out,
                            uint32_t *flagged_out);
/* Assignment only, result left on the device (timed by bench.py without D2H). */
int qb200_assign_only(qb200_ctx *ctx, const double *codebook, uint32_t K, uint32_t *flagged_out,
                      float *ms_assign_out, float *ms_resolve_out);

/* fixCodeVectors + the two updateDistortion values from integer statistics (O(K*dim) FP64 on the
 * host; SURVEY.md 8a "key simplification"): centroid = (S_t/255.0)/n for SCALED, S/n for NORMAL,
 * zero vector for an empty cell.  codebook_pre may be NULL (then *dist_pre is not written). */
int qb200_finalize_level(int colorspace, uint32_t K, int dim, uint64_t n_total,
                         const uint64_t *count, const int64_t *sum, const uint64_t *sqsum,
                         const double *codebook_pre, double *codebook_post, double *dist_pre,
                         double *dist_post);

/* Codebook doubles -> bytes exactly as vectorsToCharVectorsColorSpaced + colorSpaceToRGB do
 * (src/Compressor.cpp:12-29, src/ColorSpace.cpp:8-11,23-28). Host-side, O(K*dim). */
int qb200_codebook_to_bytes(const double *codebook, size_t K, int dim, int colorspace,
                            uint8_t *bytes_out);

/* CompressedImage::decompress + getImageFromVectors (src/Compressor.cpp:64-92,156-165) on the
 * device: rebuilds the RGB image from codebook bytes and indices.  Also returns the pixel-domain
 * mean squared error against the context's current image when mse_out != NULL (the report's
 * "Distortion", src/Compressor.cpp:137-146).  rgb_out: n_images*xSize*ySize*3 host bytes or NULL. */
int qb200_decode(qb200_ctx *ctx, const uint8_t *codebook_bytes, uint32_t K, uint8_t *rgb_out,
                 double *mse_out);

/* Measured FP32 FMA throughput of this device (TFLOP/s, FMA = 2 flop): the in-run roofline
 * denominator for the assignment kernel. */
int qb200_measure_fp32_peak(qb200_ctx *ctx, double *tflops_out);

/* Host-only diagnostic: builds the nanoflann-order KD tree (leaf size 10, src/KDTree.cpp:4;
 * nanoflann.hpp:1046-1186) the resolver walks and returns its point order (K entries), node count
 * and depth, so that the host logic can be checked without a GPU. */
int qb200_debug_kd_build(const double *points, size_t K, int dim, uint32_t *order_out,
                         int *n_nodes_out, int *depth_out);

/* Host-only diagnostic: the robustness margin of that tree (kd_host.hpp, KdHostTree::min_margin) - the smallest
 * relative distance from flipping of any comparison that shaped it; exact_flags (K bytes or NULL) marks the points
 * whose coordinates are the same numbers with either centroid arithmetic.  The auto centroid mode trusts the tree's
 * visiting order only above 1e-9. */
int qb200_debug_kd_margin(const double *points, size_t K, int dim, const uint8_t *exact_flags, double *margin_out);
/* Diagnostics: the whole host KD tree of a codebook - nodes (32 bytes each: int child1, child2, a, b; double divlow,
 * divhigh; inner nodes: a = cut dimension | census bits 16, 17; leaves: children -1, [a, b) = positions in order_out),
 * nanoflann's point order and the robustness margin.  nodes_out may be NULL (count only). */
int qb200_debug_kd_tree(const double *points, size_t K, int dim, const uint8_t *exact_flags, void *nodes_out, size_t nodes_cap,
                        int *n_nodes_out, uint32_t *order_out, double *margin_out);
/* Diagnostics: the codebook split level `level` (0-based: 2^(level+1) codevectors, before their centroid update)
 * of the last HEAD-schedule train started from, as kept on the device for the auto mode's restart. */
int qb200_debug_level_codebook(qb200_ctx *ctx, int level, double *codebook_out);

/* Diagnostic: the per-query records {best score, second best score, chunk index (bits), 0} the tensor-core
 * filter left behind in its last pass (num_vectors x 4 floats, lattice units: score = |C|^2 - 2<X,C>), so that
 * its rounding error can be audited against the bound its margin is derived from.  Valid right after
 * qb200_assign_only / qb200_assign_accumulate on a codebook of 128 or more entries; set the environment variable
 * QB200_DEBUG_RECORDS=1 first (the fused tensor-core kernel keeps no records otherwise). */
int qb200_debug_filter_records(qb200_ctx *ctx, float *records_out);

/* Number of kernels this library has launched in this process since the last reset
 * (bench.py's "gpu_launches"). */
int qb200_launch_count(int reset);

#ifdef __cplusplus
}
#endif
#endif /* QB200_H */
