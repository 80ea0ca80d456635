/* libssw -- C ABI of the B200-native embed / extract / similarity hot path of
 * iwanders/spread_spectrum_watermarking (Cox et al. 1997).
 *
 * This is the drop-in boundary: every entry point mirrors one item of the reference crate's public
 * Rust API (paths below are relative to the reference repository).  Plain pointers and sizes only.
 * All arithmetic runs in hand-written sm_100a CUDA kernels; there is NO CPU fallback -- without a
 * CUDA device `ssw_ctx_create` fails and nothing else can be called.
 *
 * Conventions
 *   - every function returns SSW_OK (0) or a negative ssw_status; `ssw_last_error()` gives the text
 *     (thread-local).  The reference panics where these return an error (INTEGRATION.md shows the
 *     Rust shim turning non-zero into panic!).
 *   - images are row-major, interleaved RGB, 8-bit or f32 in [0,1] (image::DynamicImage::into_rgb8 /
 *     into_rgb32f layout).  "host" pointers are ordinary (ideally pinned) host memory, "_dev"
 *     variants take device pointers on the context's device and never synchronise the host.
 *   - a context is bound to one device and one CUDA stream; it is not thread-safe (the reference's
 *     Writer/Reader are !Send + !Sync: src/algorithm.rs:286-291,441-445).
 */
#ifndef SSW_H_
#define SSW_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum ssw_status {
    SSW_OK = 0,
    SSW_ERR_INVALID = -1,     /* bad argument (the reference would panic / fail an assert)      */
    SSW_ERR_CUDA = -2,        /* CUDA runtime error, or no CUDA device                          */
    SSW_ERR_UNSUPPORTED = -3, /* Custom closures, line lengths the FFT planner cannot handle    */
    SSW_ERR_STATE = -4        /* e.g. extract called on a derived reader (reference: unwrap())  */
} ssw_status;

/* Insertion::Option1/2/3 and Extraction::Option1/2/3 -- src/algorithm.rs:68-78,115-125 */
enum { SSW_METHOD_OPTION1 = 1, SSW_METHOD_OPTION2 = 2, SSW_METHOD_OPTION3 = 3 };
/* OrderingMethod::{Energy, EnergyOrthogonal, Legacy} -- src/algorithm.rs:143-152 */
enum { SSW_ORDER_ENERGY = 0, SSW_ORDER_ENERGY_ORTHOGONAL = 1, SSW_ORDER_LEGACY = 2 };
/* dct2d::Type -- src/dct2d.rs:71-79 */
enum { SSW_DCT2 = 0, SSW_DCT2_ORTHOGONAL = 1, SSW_DCT3 = 2 };

/* WriteConfig / ReadConfig -- src/algorithm.rs:99-112,127-140.  Defaults: {2, 0.1f, 0}.
 * Insertion::Custom / Extraction::Custom / OrderingMethod::Custom (host closures) are rejected. */
typedef struct ssw_config {
    int32_t method;   /* SSW_METHOD_OPTION{1,2,3}                  */
    float alpha;      /* the scaling passed to OptionN(alpha)      */
    int32_t ordering; /* SSW_ORDER_*                               */
} ssw_config;

typedef struct ssw_ctx ssw_ctx;
typedef struct ssw_writer ssw_writer;
typedef struct ssw_reader ssw_reader;
typedef struct ssw_bank ssw_bank;

const char* ssw_last_error(void);
const char* ssw_version(void);

/* ---- context: replaces rustdct::DctPlanner (src/algorithm.rs:309,477): plan + twiddle cache,
 *      workspace, stream.  `stream` is a cudaStream_t (NULL = create an own non-blocking stream). */
int ssw_ctx_create(int device, ssw_ctx** out);
int ssw_ctx_create_on_stream(int device, void* stream, ssw_ctx** out);
/* Waits for the context's stream and drops the caller's reference.  Writers, readers, banks and sharded frames created
 * from the context keep it alive until they are destroyed themselves, so the order of the destroy calls is free. */
int ssw_ctx_destroy(ssw_ctx* ctx);
int ssw_ctx_synchronize(ssw_ctx* ctx);
void* ssw_ctx_stream(ssw_ctx* ctx);
/* diagnostics: per-CTA time line of the persistent row / column pipelines (load issued, tile landed, compute done, store
 * issued ... as clock64 values, layout in csrc/dct_pipe.cuh).  dev_buf: 16 x 1024 x 64 int64 (8 MiB) of device memory;
 * launch i of a pipeline kernel writes block i % 16.  NULL switches it off (default).  tools/pipe_trace.py prints it. */
int ssw_ctx_set_trace(ssw_ctx* ctx, void* dev_buf);
/* number of kernels this context has launched so far (bench.py reports it as gpu_launches) */
uint64_t ssw_ctx_launch_count(ssw_ctx* ctx);
/* per-kernel timing: between _begin and _end every kernel launch of this context is bracketed by
 * CUDA events on the context's stream; _end writes {"kernel": {"launches": n, "ms": total}, ...}
 * as JSON text into `json_out` (bench.py's roofline attribution). */
int ssw_ctx_profile_begin(ssw_ctx* ctx);
int ssw_ctx_profile_end(ssw_ctx* ctx, char* json_out, size_t cap);
/* tuning knobs: line pairs per CTA tile for the row / column passes (0 = automatic) */
int ssw_ctx_set_tiling(ssw_ctx* ctx, int row_pairs, int col_pairs);

/* pinned host memory helpers (cudaHostAlloc / cudaFreeHost) */
int ssw_host_alloc(size_t bytes, void** out);
int ssw_host_free(void* p);

/* ---- dct2d::dct2_2d(planner, type, width, height, data) -- src/dct2d.rs:83-219.
 *      `data` is [height][width] f32, transformed in place (host or device pointer). */
int ssw_dct2_2d(ssw_ctx* ctx, int type, uint32_t width, uint32_t height, float* data_host);
int ssw_dct2_2d_dev(ssw_ctx* ctx, int type, uint32_t width, uint32_t height, float* data_dev);

/* ---- yiq: From<&Rgb32FImage> for YIQ32FImage / From<&YIQ32FImage> for Rgb32FImage
 *      -- src/yiq.rs:177-197 (planes are separate [h][w] f32 buffers, src/yiq.rs:58-62). */
int ssw_rgb32f_to_yiq(ssw_ctx* ctx, const float* rgb_host, uint32_t width, uint32_t height,
                      float* y_host, float* i_host, float* q_host);
int ssw_yiq_to_rgb32f(ssw_ctx* ctx, const float* y_host, const float* i_host, const float* q_host,
                      uint32_t width, uint32_t height, float* rgb_host);

/* ---- Writer -- src/algorithm.rs:286-433 */
/* Writer::new(image, config): upload, RGB->Y, forward 2-D DCT (:295-316).  The coefficient
 * ordering (:324-327) is evaluated lazily at embed time for the mark length actually needed. */
int ssw_writer_new_rgb8(ssw_ctx* ctx, const uint8_t* rgb_host, uint32_t width, uint32_t height,
                        const ssw_config* cfg, ssw_writer** out);
int ssw_writer_new_rgb32f(ssw_ctx* ctx, const float* rgb_host, uint32_t width, uint32_t height,
                          const ssw_config* cfg, ssw_writer** out);
int ssw_writer_new_rgb8_dev(ssw_ctx* ctx, const uint8_t* rgb_dev, uint32_t width, uint32_t height,
                            const ssw_config* cfg, ssw_writer** out);
/* Writer::embed(&mut self, marks) (:348-352, embed_watermark :382-410).  Marks longer than
 * width*height-1 are truncated like the reference's zip (:396).  Several marks: deltas against the
 * original coefficients are summed (:399-408).  The reference fixes the ordering in Writer::new (:324-327); here it is
 * computed for the longest mark of the first embed (or an earlier ssw_writer_indices call): a later embed that needs
 * MORE ordered coefficients returns SSW_ERR_STATE instead of ordering already modified coefficients. */
int ssw_writer_embed(ssw_writer* w, const float* const* marks_host, const size_t* lens, size_t n_marks);
/* Writer::coefficient_image() (:319-321): copy of the [h][w] coefficient plane. */
int ssw_writer_coefficients(ssw_writer* w, float* out_host);
/* the ordered coefficient indices (row-major r*w+c, DC excluded) the embedding used / would use:
 * first n entries of obtain_indices_by_function (:200-210). */
int ssw_writer_indices(ssw_writer* w, uint64_t* out_host, size_t n);
/* Writer::result(self) (:361-379) [+ DynamicImage::into_rgb8()]: inverse DCT, YIQ->RGB, download.
 * Like the reference this consumes the coefficients; call once, then destroy the writer. */
int ssw_writer_result_rgb8(ssw_writer* w, uint8_t* out_host);
int ssw_writer_result_rgb32f(ssw_writer* w, float* out_host);
int ssw_writer_result_rgb8_dev(ssw_writer* w, uint8_t* out_dev);
int ssw_writer_destroy(ssw_writer* w);

/* ---- Reader / ReaderDerived -- src/algorithm.rs:435-594 */
int ssw_reader_base_rgb8(ssw_ctx* ctx, const uint8_t* rgb_host, uint32_t width, uint32_t height,
                         const ssw_config* cfg, ssw_reader** out);          /* Reader::base    :462 */
int ssw_reader_base_rgb32f(ssw_ctx* ctx, const float* rgb_host, uint32_t width, uint32_t height,
                           const ssw_config* cfg, ssw_reader** out);
int ssw_reader_derived_rgb8(ssw_ctx* ctx, const uint8_t* rgb_host, uint32_t width, uint32_t height,
                            ssw_reader** out);                              /* Reader::derived :469 */
int ssw_reader_derived_rgb32f(ssw_ctx* ctx, const float* rgb_host, uint32_t width, uint32_t height,
                              ssw_reader** out);
int ssw_reader_base_rgb8_dev(ssw_ctx* ctx, const uint8_t* rgb_dev, uint32_t width, uint32_t height,
                             const ssw_config* cfg, ssw_reader** out);
int ssw_reader_derived_rgb8_dev(ssw_ctx* ctx, const uint8_t* rgb_dev, uint32_t width, uint32_t height,
                                ssw_reader** out);
/* Reader::extract(&self, &derived, extracted) (:529-562).  Errors (reference panics): `base` is a
 * derived reader (:530 unwrap), sizes differ (:550-552), n >= width*height (:553-555). */
int ssw_reader_extract(ssw_reader* base, ssw_reader* derived, float* out_host, size_t n);
int ssw_reader_extract_dev(ssw_reader* base, ssw_reader* derived, float* out_dev, size_t n);
int ssw_reader_coefficients(ssw_reader* r, float* out_host);                /* :502-504 */
/* Reader::indices() (:506-508) -- first n ordered indices (the reference returns all w*h-1). */
int ssw_reader_indices(ssw_reader* base, uint64_t* out_host, size_t n);
int ssw_reader_destroy(ssw_reader* r);

/* ---- Tester::similarity -- src/algorithm.rs:696-714.  One kernel thread walks one mark in the
 *      reference's sequential f32 order, so the result is bit-identical to the reference loop.
 *      Bank searches and the 1:1 scores of the fused pipelines default to a fixed-shape tree reduction
 *      (deterministic, a few ulp from the sequential loop; north_star asks for 1e-3 relative) that runs at HBM
 *      speed; SSW_SIM_EXACT=1 in the environment of ssw_ctx_create selects the sequential order there too. */
int ssw_similarity(ssw_ctx* ctx, const float* extracted_host, const float* mark_host, size_t n, float* out);
/* bank of stored marks [n_marks][n] resident on the device (README.md:62 "test against any number
 * of marks"); out is [n_extracted][n_marks]. */
int ssw_bank_create(ssw_ctx* ctx, const float* marks_host, size_t n_marks, size_t n, ssw_bank** out);
int ssw_bank_create_normal(ssw_ctx* ctx, uint64_t seed, size_t n_marks, size_t n, ssw_bank** out);
int ssw_bank_row(ssw_bank* bank, size_t index, float* out_host);
int ssw_bank_similarity(ssw_bank* bank, const float* extracted_host, size_t n_extracted, float* out_host);
int ssw_bank_similarity_dev(ssw_bank* bank, const float* extracted_dev, size_t n_extracted, float* out_dev);
int ssw_bank_destroy(ssw_bank* bank);

/* ---- MarkBuf::generate_normal(length) -- src/algorithm.rs:619-626 (unseeded thread_rng there;
 *      seed == 0 draws the seed from the OS, anything else is reproducible). */
int ssw_mark_generate_normal(ssw_ctx* ctx, uint64_t seed, size_t n, float* out_host);

/* ---- fused device-resident pipelines (bench configs 2/3; no host synchronisation inside) ----
 * embed : Writer::new(img,cfg).mark(&[mark]).into_rgb8() for `batch` images of equal size.
 * extract: Reader::base + Reader::derived + extract (+ Tester::similarity if sim_dev != NULL).
 * rgb buffers are [batch][h][w][3] u8, marks [batch][n] f32, extracted [batch][n], sim [batch]. */
int ssw_embed_batch_rgb8_dev(ssw_ctx* ctx, const uint8_t* rgb_dev, uint32_t width, uint32_t height,
                             uint32_t batch, const ssw_config* cfg, const float* marks_dev, size_t n,
                             uint8_t* out_rgb_dev);
int ssw_extract_batch_rgb8_dev(ssw_ctx* ctx, const uint8_t* base_rgb_dev, const uint8_t* derived_rgb_dev,
                               uint32_t width, uint32_t height, uint32_t batch, const ssw_config* cfg,
                               size_t n, float* extracted_dev, const float* marks_dev, float* sim_dev);
/* host-buffer versions of the same (pinned memory recommended): the end-to-end API path */
int ssw_embed_batch_rgb8(ssw_ctx* ctx, const uint8_t* rgb_host, uint32_t width, uint32_t height,
                         uint32_t batch, const ssw_config* cfg, const float* marks_host, size_t n,
                         uint8_t* out_rgb_host);
int ssw_extract_batch_rgb8(ssw_ctx* ctx, const uint8_t* base_rgb_host, const uint8_t* derived_rgb_host,
                           uint32_t width, uint32_t height, uint32_t batch, const ssw_config* cfg,
                           size_t n, float* extracted_host, const float* marks_host, float* sim_host);
/* asynchronous forms of the host-buffer calls: enqueue and return; results are valid after ssw_ctx_synchronize().
 * Consecutive calls overlap (upload of call i+1 beside the kernels and the download of call i: PCIe is full duplex).
 * Host buffers may be reused between calls without synchronising: a copy touching a host range with an earlier copy
 * still in flight is ordered behind it (embed -> extract on the same buffers works).  The caller must not touch the
 * buffers itself before ssw_ctx_synchronize().  Overflowing frames are handled as by the _dev entry points. */
int ssw_embed_batch_rgb8_async(ssw_ctx* ctx, const uint8_t* rgb_host, uint32_t width, uint32_t height,
                               uint32_t batch, const ssw_config* cfg, const float* marks_host, size_t n,
                               uint8_t* out_rgb_host);
int ssw_extract_batch_rgb8_async(ssw_ctx* ctx, const uint8_t* base_rgb_host, const uint8_t* derived_rgb_host,
                                 uint32_t width, uint32_t height, uint32_t batch, const ssw_config* cfg,
                                 size_t n, float* extracted_host, const float* marks_host, float* sim_host);
/* completion markers: ssw_ctx_marker hands out a ticket for "everything enqueued so far", ssw_ctx_wait_marker blocks
 * the host until then (results of those calls are in host memory) without draining later calls. */
int ssw_ctx_marker(ssw_ctx* ctx, uint64_t* marker);
int ssw_ctx_wait_marker(ssw_ctx* ctx, uint64_t marker);
/* Number of frames of the fused calls since the last query whose ordered top-k could not be served by the
 * candidate list (degenerate, noise-like spectrum: more than SSW_TOPK_CAP near-equal keys).  Synchronises the
 * stream.  What the calls did with such frames:
 *   _dev entry points (no host synchronisation inside, hence no repair): the frame is left UNMARKED (embed:
 *     the output is the input up to the transform round trip; extract: the vector is zero, the score NaN) --
 *     never modified through a wrong index list.  Callers poll this counter and re-run those frames through
 *     the Writer / Reader API, which repairs (full-plane histogram, then the exact general sort).
 *   host-buffer entry points: re-run the whole batch once with the full-plane histogram; if frames still
 *     overflow they return SSW_ERR_UNSUPPORTED (outputs of those frames as above). */
int ssw_ctx_last_topk_fallbacks(ssw_ctx* ctx);

/* device self-test: the packed f32 -> RGB8 output conversion of the inverse row passes against
 * round(clamp(v, 0, 1) * 255) (image::into_rgb8, tests/single_simple.rs:28) over all 2^32 float bit patterns */
int ssw_selftest_pack_u8(ssw_ctx* ctx, uint64_t* mismatches);

/* ---- synthetic frames for the benchmark (SURVEY.md section 8(d) generator, integer only) */
int ssw_synth_frame_rgb8_dev(ssw_ctx* ctx, uint32_t width, uint32_t height, uint64_t seed,
                             uint32_t first_image, uint32_t n_images, uint8_t* out_dev);

/* rows [row0, row0+n_rows) of synthetic frame `image` (one rank's shard of a sharded frame) */
int ssw_synth_rows_rgb8_dev(ssw_ctx* ctx, uint32_t width, uint64_t seed, uint32_t image, uint32_t row0, uint32_t n_rows,
                            uint8_t* out_dev);

/* ---- per-stage timing hooks for bench.py / profiling: run one stage of the pipeline on
 *      device-resident data (used to attribute time to kernels with CUDA events). */
int ssw_stage_forward_rgb8_dev(ssw_ctx* ctx, const uint8_t* rgb_dev, uint32_t width, uint32_t height,
                               uint32_t batch, float* plane_dev);
int ssw_stage_topk_dev(ssw_ctx* ctx, const float* plane_dev, uint32_t width, uint32_t height,
                       uint32_t batch, int ordering, size_t k, uint32_t* idx_dev);
int ssw_stage_inverse_rgb8_dev(ssw_ctx* ctx, float* plane_dev, const uint8_t* rgb_src_dev,
                               uint32_t width, uint32_t height, uint32_t batch, uint8_t* out_rgb_dev);

/* ---- sharded single frames (BASELINE config 4; nothing in the reference corresponds -- it holds a whole
 *      frame on one core).  Rank g owns rows [g*H/G, (g+1)*H/G) of the pixels and, after the row pass and an
 *      all-to-all transpose, columns [col0, col0+ncols) of the coefficients, stored TRANSPOSED as a local
 *      plane [ncols][height] so that the column pass is again a pass over contiguous lines.  These entry
 *      points are the per-rank steps; the exchanges between them (all-to-all, max, all-gather, sum) are
 *      issued by the host layer over NCCL.  Index lists always hold the reference's flat indices
 *      p = r*width + c (src/algorithm.rs:204), so results are comparable with the unsharded path. */
typedef struct ssw_shard {
    uint32_t width, height; /* whole frame                    */
    uint32_t col0, ncols;   /* coefficient columns of this rank */
} ssw_shard;
#define SSW_TOPK_CAP 8192   /* entries of one candidate list */

/* 1-D DCT-II (x2, reference scaling) of n_lines contiguous lines of length n; src_type 0 RGB8 (luma is
 * computed, src/yiq.rs:157), 1 RGB32F, 2 f32 lines.  The row pass of dct2_2d (src/dct2d.rs:129-170). */
int ssw_lines_forward_dev(ssw_ctx* ctx, int src_type, const void* src_dev, uint32_t n, uint32_t n_lines, float* plane_dev);
/* the same over lines assembled from all-to-all blocks, read in place (no interleaving copy): src layout
 * [chunks][ranks][n_lines][seg_len]; SSW_ERR_UNSUPPORTED unless seg_len, chunks are powers of two and n is planned */
int ssw_lines_forward_seg_dev(ssw_ctx* ctx, const float* src_dev, uint32_t n, uint32_t n_lines, uint32_t seg_len,
                              uint32_t chunks, uint32_t ranks, float* plane_dev);
/* 1-D DCT-III (x0.5) of the lines, then x scale; dst_type 2: f32 lines, 0 / 1: RGB8 / RGB32F with the chroma
 * of the original pixels `src_dev` (src/yiq.rs:187-197). */
int ssw_lines_inverse_dev(ssw_ctx* ctx, float* plane_dev, uint32_t n, uint32_t n_lines, float scale, int dst_type,
                          void* dst_dev, int src_type, const void* src_dev);
/* dst[b][c][r] = src[b][r][c]: batched f32 transpose (packs / unpacks the all-to-all blocks) */
int ssw_transpose_dev(ssw_ctx* ctx, const float* src_dev, uint32_t rows, uint32_t cols, int64_t src_ld, int64_t src_bstride,
                      float* dst_dev, int64_t dst_ld, int64_t dst_bstride, uint32_t batch);
/* the same with the coefficient lines held as all-to-all blocks [chunks][ranks][n_lines][seg_len], read in place */
int ssw_lines_inverse_seg_dev(ssw_ctx* ctx, const float* src_dev, uint32_t n, uint32_t n_lines, uint32_t seg_len,
                              uint32_t chunks, uint32_t ranks, float scale, int dst_type, void* dst_dev, int pix_type,
                              const void* pixels_dev);
/* distributed ordered top-k (obtain_indices_by_function, src/algorithm.rs:200-221): local bound -> [max over
 * ranks] -> local candidates -> [all-gather] -> merge */
int ssw_shard_topk_bin_dev(ssw_ctx* ctx, const float* plane_dev, const ssw_shard* sh, int ordering, size_t k, uint32_t* bin_dev);
int ssw_shard_topk_collect_dev(ssw_ctx* ctx, const float* plane_dev, const ssw_shard* sh, int ordering,
                               const uint32_t* bin_dev, uint64_t* cand_dev, uint32_t* count_dev);
int ssw_shard_topk_merge_dev(ssw_ctx* ctx, const uint64_t* lists_dev, const uint32_t* counts_dev, uint32_t n_lists,
                             size_t k, uint32_t* idx_dev, uint32_t* overflow_dev);
/* embed_watermark / extract_watermark on the coefficients this rank owns (src/algorithm.rs:382-410,543-562);
 * extract writes 0 for coefficients owned by other ranks (the per-rank vectors are summed) */
int ssw_shard_embed_dev(ssw_ctx* ctx, float* plane_dev, const ssw_shard* sh, const uint32_t* idx_dev, size_t k,
                        const float* marks_dev, size_t mark_stride, size_t n_marks, const uint32_t* lens_dev,
                        const ssw_config* cfg);
int ssw_shard_extract_dev(ssw_ctx* ctx, const float* base_plane_dev, const float* derived_plane_dev, const ssw_shard* sh,
                          const uint32_t* idx_dev, size_t n, const ssw_config* cfg, float* out_dev);

/* ---- the same behind one object per rank: orchestration in the library, exchange through peer-mapped memory.
 *      One process per GPU of one node.  ssw_sharded_create maps the coefficient planes of all ranks into each
 *      other (CUDA IPC over NVLink / NVSwitch); the block transposes between the two passes of the transform STORE
 *      their tiles straight into the owner's plane, slice by slice beside the line kernels of the next slice -- there is
 *      no all-to-all collective.  NCCL (resolved with dlopen at the first use; no link-time dependency) carries the IPC
 *      handles, the barrier that closes an exchange and the few hundred bytes of the distributed top-k.
 *      Bootstrap: rank 0 calls ssw_sharded_unique_id and hands the 128 bytes to the other ranks by any means (the
 *      host application's own channel: MPI, a socket, torch.distributed in the tests), then every rank calls
 *      ssw_sharded_create (collective).  embed / extract are collective calls, stream-ordered on the context's stream,
 *      without host synchronisation.  Mirrors Writer::new(img,cfg).mark(&[mark]).into_rgb8() and Reader::base +
 *      Reader::derived + extract (src/algorithm.rs:295-379, 462-562) for frames of up to 2^32-2 pixels. */
typedef struct ssw_sharded ssw_sharded;
#define SSW_SHARDED_ID_BYTES 128
int ssw_sharded_unique_id(void* id_out /* SSW_SHARDED_ID_BYTES */);
int ssw_sharded_create(ssw_ctx* ctx, const void* id, int rank, int world, uint32_t width, uint32_t height, ssw_sharded** out);
int ssw_sharded_destroy(ssw_sharded* s);
/* rows_dev / out_rows_dev: this rank's rows [rank*H/G, (rank+1)*H/G) as [H/G][W][3] u8; mark_dev: n floats (all ranks) */
int ssw_sharded_embed_rgb8_dev(ssw_sharded* s, const uint8_t* rows_dev, const ssw_config* cfg, const float* mark_dev, size_t n,
                               uint8_t* out_rows_dev);
/* extracted_dev: n floats, the whole vector on every rank */
int ssw_sharded_extract_rgb8_dev(ssw_sharded* s, const uint8_t* base_rows_dev, const uint8_t* derived_rows_dev,
                                 const ssw_config* cfg, size_t n, float* extracted_dev);
/* first n ordered indices (flat r*W + c) of the last embed / extract; this rank's coefficient columns of the base (0) /
 * derived (1) frame, transposed [W/G][H]; the sticky candidate-overflow flag (read and clear).  All three synchronise. */
int ssw_sharded_indices(ssw_sharded* s, uint32_t* out_host, size_t n);
int ssw_sharded_coefficients(ssw_sharded* s, int which, float* out_host);
int ssw_sharded_overflow(ssw_sharded* s, int* overflowed);

#ifdef __cplusplus
}
#endif
#endif /* SSW_H_ */
