/* lars_b200 -- C ABI of the B200-native RGNir per-pixel analysis path.
 *
 * The reference (lars-uav/lars-image-processing) has no FFI: its boundary for this path is
 * a set of module-level Python helpers over NumPy.  Each entry
 * point below names the reference interface whose arithmetic it replaces (file:line
 * relative to the reference repository); lars_image_processing_b200/*.py re-exposes them
 * under the reference's own function names, and INTEGRATION.md shows the ctypes stub a
 * maintainer would add.
 *
 * Conventions
 *  - Every data pointer is a DEVICE pointer (sm_100a, B200).  The library never allocates,
 *    never copies host<->device on the data path and never synchronises the stream; the
 *    caller (PyTorch host code: tensor.data_ptr(), torch.cuda.current_stream()) owns memory
 *    and ordering.  `stream` is a cudaStream_t passed as void*.
 *  - Return value: 0 (LARS_OK) or a negative LARS_ERR_* code; lars_last_error() returns a
 *    thread-local description.  The library never calls abort()/exit().
 *  - Thread safety: all entry points are re-entrant; lars_init() is idempotent and
 *    internally locked.  There is no CPU fallback: without a B200 every compute entry
 *    point fails with LARS_ERR_CUDA / LARS_ERR_UNSUPPORTED.
 *  - Frames are interleaved HWC, treated as flat arrays of n_pixels pixels.  Frame f of a
 *    batch starts at base + f * frame_stride; every frame start must be 16-byte aligned
 *    and each frame slot must be padded to a whole number of 16-pixel groups (the kernels
 *    move 16-byte vectors and TMA bulk copies; bytes past n_pixels in a slot are scratch).
 */
#ifndef LARS_B200_H_
#define LARS_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LARS_OK 0
#define LARS_ERR_INVALID (-1)     /* bad argument (NULL, misaligned, out of range) */
#define LARS_ERR_CUDA (-2)        /* CUDA runtime / launch failure */
#define LARS_ERR_UNSUPPORTED (-3) /* not an sm_100 device, or an unsupported variant */
#define LARS_ERR_NOT_INIT (-4)    /* lars_init() has not succeeded for the current device */

#define LARS_NDVI 0  /* (NIR - R) / (NIR + R + eps)   process-images.py:466-470 */
#define LARS_GNDVI 1 /* (NIR - G) / (NIR + G + eps)   process-images.py:472-476 */
#define LARS_NDWI 2  /* (G - NIR) / (G + NIR + eps)   process-images.py:478-482 */
#define LARS_NUM_INDICES 3

#define LARS_MAX_BINS 64      /* histogram bins supported by the fused pass (reference: 50) */
#define LARS_PIXEL_GROUP 16   /* frame slots are padded to this many pixels */

#define LARS_CMAP_RDYLGN 0 /* process-images.py:692 */
#define LARS_CMAP_RDYLBU 1 /* process-images.py:690 */
#define LARS_CMAP_BWR 2    /* process-images.py:956 */

#define LARS_DTYPE_U8 0
#define LARS_DTYPE_U16 1
#define LARS_DTYPE_F32 2
#define LARS_DTYPE_F64 3

/* Statistics of one index map of one frame (replaces analyze_index, process-images.py:492-513,
 * the inline statistics at :647-658 / :830-832, analyze_ndvi_statistics, process-ndvi.py:50-73,
 * and the 50-bin histogram at process-ndvi.py:97). 576 bytes, all fields naturally aligned. */
typedef struct lars_index_stats {
  uint64_t count;       /* pixels                                                        */
  uint64_t count_above; /* pixels with x > threshold, compared in float32 (0.2f / 0.0f)   */
  double sum;           /* sum of x      (float64 accumulation of the float32 values)     */
  double sumsq;         /* sum of x * x                                                    */
  double mean;          /* sum / count                                                     */
  double std;           /* population standard deviation, sqrt(sumsq/count - mean^2)       */
  float min;
  float max;
  float threshold;
  uint32_t bins;
  uint64_t hist[LARS_MAX_BINS]; /* np.histogram(x, bins, range=(-1, 1)); entries >= bins are 0 */
} lars_index_stats;

/* White-balance stretch of one channel of a uint16 frame: the monotone map
 * v -> uint8(clip((v - p_lo) / (p_hi - p_lo) * 255, 0, 255)) (process-images.py:438-441) stored as
 * its 256 step positions thr[k] = smallest v with LUT(v) >= k (thr[0] = 0, unreachable k = 65536),
 * laid out as pairs (thr[k], thr[k + 1]) so one 8-byte load brackets a guess, plus the parameters
 * of that guess: t = ((v - lo_int) - lo_frac) * scale.  2064 bytes. */
typedef struct lars_stretch_u16 {
  int32_t lo_int;    /* floor(p_lo)                                         */
  float lo_frac;     /* p_lo - floor(p_lo)                                  */
  float scale;       /* 255 / (p_hi - p_lo); 2^20 around the single step if p_hi == p_lo */
  uint32_t reserved;
  uint32_t pairs[256][2];
} lars_stretch_u16;

/* Arguments of the fused Pass 2 (one struct so the ABI can grow without breaking callers). */
typedef struct lars_fused_args {
  uint32_t struct_bytes;      /* = sizeof(lars_fused_args), checked                         */
  int32_t n_frames;
  int32_t channels;           /* 3 (RGN) or 4 (RGNA; alpha is ignored and written as 0)     */
  int32_t bins;               /* 1..LARS_MAX_BINS                                            */
  int64_t n_pixels;           /* pixels per frame                                            */
  const uint8_t* src;         /* raw frames (uint8 samples, or little-endian uint16 for *_u16) */
  int64_t src_frame_stride;   /* bytes                                                       */
  const uint8_t* wb_lut;      /* u8: [frame][3][256] stretch LUTs, NULL = identity (no WB);
                                 u16: [frame][3] lars_stretch_u16 (required)                 */
  int64_t lut_frame_stride;   /* bytes between frames' LUTs; 0 = one LUT set for all frames  */
  uint8_t* wb_out;            /* white-balanced frames, same channel count as src, or NULL   */
  int64_t wb_frame_stride;    /* bytes                                                       */
  float* maps[LARS_NUM_INDICES];   /* float32 index maps (NDVI, GNDVI, NDWI), each or all NULL */
  int64_t map_frame_stride;   /* elements                                                    */
  uint8_t* rgb[LARS_NUM_INDICES];  /* colormapped HWC RGB images, each or all NULL            */
  int64_t rgb_frame_stride;   /* bytes                                                       */
  int32_t cmap[LARS_NUM_INDICES];  /* LARS_CMAP_* per index (reference: RdYlGn, RdYlGn, RdYlBu) */
  float thresholds[LARS_NUM_INDICES]; /* coverage thresholds (reference: 0.2, 0.2, 0.0)      */
  lars_index_stats* stats;    /* [frame][3] or NULL                                          */
  void* workspace;            /* lars_fused_workspace_bytes(n_frames) bytes, 16-byte aligned */
  size_t workspace_bytes;
} lars_fused_args;

/* ---- lifecycle ---------------------------------------------------------------------- */
int lars_init(int device_ordinal); /* checks sm_100, sets kernel attributes, uploads colormaps */
int lars_shutdown(void);
const char* lars_last_error(void);
int lars_abi_version(void);
int lars_sm_count(void);

/* ---- host-only helpers (no GPU needed) ----------------------------------------------- */
/* matplotlib 'RdYlGn' / 'RdYlBu' / 'bwr' as 256 RGB byte triplets
 * (process-images.py:689-695, :956; backend-process.py:42; process-ndvi.py:38). */
int lars_colormap_table(int cmap_id, uint8_t* rgb_out /* [256][3] host */);
/* np.linspace(-1, 1, bins + 1) rounded to float32 (numpy/lib/_histograms_impl.py:440-447). */
int lars_histogram_edges_f32(int bins, float* edges_out /* [bins + 1] host */);

/* ---- Pass 1: white-balance statistics -------------------------------------------------
 * Replaces the three np.percentile(channel, (2, 98)) calls of fix_white_balance
 * (process-images.py:435-437; backend-process.py:21-23; process-rgn.py:28).
 * hist is [n_frames][3][256] uint64 and is ZEROED by the call, then filled.  With shared_hist != 0
 * all frames accumulate into ONE [3][256] set: the frames are tiles of a single image whose
 * percentiles are global (tile-sharded orthomosaic). */
int lars_wb_hist_u8(const uint8_t* src, int32_t n_frames, int64_t n_pixels, int32_t channels,
                    int64_t src_frame_stride, uint64_t* hist, int32_t shared_hist, void* stream);

/* Percentiles (NumPy "linear" method, float64) and the stretch
 * clip((v - p_lo) / (p_hi - p_lo) * 255, 0, 255) -> float32 -> uint8 for every v in 0..255
 * (process-images.py:437-441).  n_sets histogram sets in, n_sets LUT sets out:
 * lut [n_sets][3][256] uint8, pct [n_sets][3][2] float64 (may be NULL).
 * q_lo / q_hi are fractions (reference: 0.02, 0.98). */
int lars_wb_lut_build_u8(const uint64_t* hist, int32_t n_sets, double q_lo, double q_hi,
                         uint8_t* lut, double* pct, void* stream);
/* The same with the reference expression named: LARS_WB_CHAIN_IMAGES is the call above
 * (process-images.py:437-441, backend-process.py:21-26: float64 stretch, float32 store, truncation);
 * LARS_WB_CHAIN_RGN is fix_white_balance_rgnir (process-rgn.py:25-33, :44): samples clipped to
 * [p_lo, p_hi] first, float64 all the way, the float64 value truncated.  The two differ by one step
 * wherever the float64 stretch lands a hair below an integer (fractional percentiles). */
#define LARS_WB_CHAIN_IMAGES 0
#define LARS_WB_CHAIN_RGN 1
int lars_wb_lut_build_u8_chain(const uint64_t* hist, int32_t n_sets, double q_lo, double q_hi, int32_t chain,
                               uint8_t* lut, double* pct, void* stream);

/* K1b fused with the exchange of a tile-sharded image (BASELINE config 4; the percentiles of
 * process-images.py:435-438 are global to the image, so the ranks' local histograms must be summed between Pass 1 and
 * the LUT build).  Instead of an NCCL all-reduce of 3 x 256 counters, ONE kernel pushes this rank's counts into every
 * peer's symmetric buffer over NVLink, raises / awaits per-rank flags there, sums the world contributions in rank
 * order and builds the table: `hist` ([1][3][256], device) holds the local counts on entry and the image-wide counts
 * on return.  peer_bufs is a DEVICE array of `world` pointers: every rank's buffer of lars_wb_peer_buffer_bytes(world)
 * bytes (zero-initialised once) as mapped into this process (torch.distributed._symmetric_memory: `buffer_ptrs_dev`;
 * any CUDA IPC / VMM mapping works).  epoch = 1, 2, 3, ... must advance by one per call on every rank.  *status
 * (device) is set to 1 if a peer did not arrive within ~10 s (the call never hangs). */
size_t lars_wb_peer_buffer_bytes(int32_t world);
int lars_wb_lut_build_u8_peers(uint64_t* hist, uint64_t* const* peer_bufs, int32_t rank, int32_t world,
                               uint32_t epoch, double q_lo, double q_hi, int32_t chain, uint8_t* lut, double* pct,
                               uint32_t* status, void* stream);

/* ---- Pass 2: fused WB + NDVI/GNDVI/NDWI + statistics + histogram + colormap -----------
 * Replaces, in one read of the raw frame: the stretch application (process-images.py:438-441),
 * calculate_index x3 (:449-490), analyze_index x3 (:492-513) minus the median, np.std
 * (process-ndvi.py:65), the 50-bin histogram (process-ndvi.py:97) and the colormap lookup
 * of create_index_visualization (:689-695). */
size_t lars_fused_workspace_bytes(int32_t n_frames);
int lars_fused_index_u8(const lars_fused_args* args, void* stream);

/* ---- uint16 frames (16-bit TIFF batches, BASELINE config 3) ----------------------------------
 * fix_white_balance accepts uint16 arrays and still returns uint8 0..255.
 * Pass 1 = two-level radix histogram (high byte, then the low byte of the buckets that hold the
 * four percentile ranks) -> exact NumPy percentiles -> lars_stretch_u16 per channel.  src strides
 * are in BYTES; workspace from lars_wb_u16_workspace_bytes(n_sets), n_sets = shared_hist ? 1 :
 * n_frames.  stretch is [n_sets][3], pct [n_sets][3][2] (may be NULL). */
size_t lars_wb_u16_workspace_bytes(int32_t n_sets);
int lars_wb_stretch_build_u16(const uint16_t* src, int32_t n_frames, int64_t n_pixels, int32_t channels,
                              int64_t src_frame_stride, double q_lo, double q_hi, lars_stretch_u16* stretch,
                              double* pct, void* workspace, size_t workspace_bytes, int32_t shared_hist,
                              void* stream);
/* The same in stages, for an image whose tiles live on several GPUs: the
 * caller SUM-all-reduces the counters between the stages so that every rank derives identical
 * thresholds.  Workspace layout: [n_sets][3][256] uint64 high-byte histogram, then
 * [n_sets][3][4][256] uint64 low-byte histograms, then private selection records.
 *   LARS_U16_STAGE_HIST_HI  zero the workspace, level A          -> all-reduce the first block
 *   LARS_U16_STAGE_HIST_LO  bucket selection + level B           -> all-reduce the second block
 *   LARS_U16_STAGE_BUILD    percentiles + thresholds into stretch / pct
 *   LARS_U16_STAGE_ALL      everything (= lars_wb_stretch_build_u16).  With one set per frame this is the guided
 *                           single pass: a 6 % sample of the frame guesses the buckets, ONE full read counts the
 *                           high bytes and the low bytes of the guessed buckets, and level B only runs for frames
 *                           whose guess missed (results are exact either way)
 *   LARS_U16_STAGE_ALL_TWO_LEVEL  everything, always as level A + level B (the round-1 form; tests and A/B timing) */
#define LARS_U16_STAGE_ALL 0
#define LARS_U16_STAGE_HIST_HI 1
#define LARS_U16_STAGE_HIST_LO 2
#define LARS_U16_STAGE_BUILD 3
#define LARS_U16_STAGE_ALL_TWO_LEVEL 4
int lars_wb_stretch_build_u16_staged(const uint16_t* src, int32_t n_frames, int64_t n_pixels, int32_t channels,
                                     int64_t src_frame_stride, double q_lo, double q_hi, lars_stretch_u16* stretch,
                                     double* pct, void* workspace, size_t workspace_bytes, int32_t shared_hist,
                                     int32_t stage, void* stream);
/* Pass 2 on uint16 frames: same products as lars_fused_index_u8 (WB output is uint8);
 * lut_frame_stride = 3 * sizeof(lars_stretch_u16) or 0 for one shared set. */
int lars_fused_index_u16(const lars_fused_args* args, void* stream);

/* ---- float-map operations next to the fused pass ---------------------------------------- */
/* Statistics + np.histogram(x, bins, range=(-1, 1)) of arbitrary float32 maps: replaces
 * analyze_index on a caller-supplied array (process-images.py:492-513), the inline statistics
 * at :647-658 / :830-832, analyze_ndvi_statistics (process-ndvi.py:50-73) and plt.hist (:97).
 * data is [n_maps][stride] float32 (16-byte aligned rows), stats is [n_maps]. */
size_t lars_map_stats_workspace_bytes(int32_t n_maps);
int lars_map_stats_f32(const float* data, int32_t n_maps, int64_t n, int64_t stride, int32_t bins,
                       float threshold, lars_index_stats* stats, void* workspace,
                       size_t workspace_bytes, void* stream);

/* Exact order statistics of a float32 map (np.median, process-images.py:508, :654;
 * process-ndvi.py:62): out3 (device) = { x[rank_lo], x[rank_hi], float32 mean of the two }
 * of the sorted data.  3-pass (11 + 11 + 10 bit) radix select, no sort, no host round trip. */
size_t lars_select_workspace_bytes(void);
int lars_select_f32(const float* data, int64_t n, uint64_t rank_lo, uint64_t rank_hi, float* out3,
                    void* workspace, size_t workspace_bytes, void* stream);

/* The float64 flavour of the two calls above, for the arrays process-ndvi.py works on: calculate_ndvi returns
 * float64 (process-ndvi.py:18-31) and analyze_ndvi_statistics / generate_ndvi_report reduce it as it is -- min,
 * max and median are float64 values, `ndvi > 0.2` compares in float64 (:60-71), plt.hist bins against float64
 * linspace edges (:97).  592 bytes, all fields naturally aligned. */
typedef struct lars_map_record_f64 {
  uint64_t count;
  uint64_t count_above; /* elements with x > threshold, compared in float64 */
  double sum;
  double sumsq;
  double mean;
  double std;           /* population standard deviation */
  double min;
  double max;
  double threshold;
  uint32_t bins;
  uint32_t has_nan;
  uint64_t hist[LARS_MAX_BINS]; /* np.histogram(x, bins, range=(-1, 1)) with float64 edges */
} lars_map_record_f64;
size_t lars_map_stats_f64_workspace_bytes(void);
int lars_map_stats_f64(const double* data, int64_t n, int32_t bins, double threshold, lars_map_record_f64* stats,
                       void* workspace, size_t workspace_bytes, void* stream);
/* out3 (device) = { x[rank_lo], x[rank_hi], float64 mean of the two }: 6-pass radix select on 64-bit keys. */
size_t lars_select_f64_workspace_bytes(void);
int lars_select_f64(const double* data, int64_t n, uint64_t rank_lo, uint64_t rank_hi, double* out3,
                    void* workspace, size_t workspace_bytes, void* stream);

/* Normalize(vmin, vmax) + colormap lookup of a float32 map -> [n][3] uint8 (the per-pixel part of
 * create_index_visualization, process-images.py:689-695; 'bwr' +-0.5 for change detection, :956). */
int lars_colormap_f32(const float* data, int64_t n, int32_t cmap_id, float vmin, float vmax,
                      uint8_t* rgb, void* stream);

/* float64 NDVI of a raw uint8 frame, no white balance (calculate_ndvi, process-ndvi.py:18-31). */
int lars_ndvi_f64_u8(const uint8_t* src, int64_t n_pixels, int32_t channels, double* out, void* stream);

/* clip((hi - lo) / (hi + lo + 1e-10), -1, 1) on separate float32 planes
 * (calculate_index(red, green, nir, index_type), backend-process.py:28-38). */
int lars_index_planes_f32(const float* hi, const float* lo, int64_t n, float* out, void* stream);

/* Merge n_sets x 3 statistics records (per frame, or per rank after the NCCL all-gather) into
 * 3 dataset-wide records, in index order of the input -- the exchange step of SURVEY.md 8(e).
 * `out` may alias one of the input sets (a running fold: in = {fold, new}, out = fold). */
int lars_stats_merge(const lars_index_stats* in, int32_t n_sets, lars_index_stats* out, void* stream);

/* Index map of an interleaved frame whose samples are not uint8 (calculate_index applies
 * astype(float32) to whatever it is given, process-images.py:456-490).  dtype: LARS_DTYPE_U16,
 * LARS_DTYPE_F32 or LARS_DTYPE_F64; index: LARS_NDVI / LARS_GNDVI / LARS_NDWI. */
int lars_index_hwc(const void* src, int32_t dtype, int64_t n_pixels, int32_t channels, int32_t index,
                   float* out, void* stream);

/* Change detection between two white-balanced uint8 frames (create_change_detection_visualization,
 * process-images.py:908-923 and :956): per-pixel index of both, diff = late - early (float32) and
 * the 'bwr' colormap over [vmin, vmax] (reference: -0.5, 0.5).  early_map / late_map / rgb may be NULL. */
int lars_index_change_u8(const uint8_t* early, const uint8_t* late, int64_t n_pixels, int32_t channels,
                         int32_t index, float vmin, float vmax, float* early_map, float* late_map,
                         float* diff, uint8_t* rgb, void* stream);

/* ---- resize in front of the path -----------------------------
 * preprocess_large_image(img_array, max_dimension=1024), process-images.py:398-422:
 * Image.fromarray(img).resize((new_w, new_h), Image.Resampling.LANCZOS).  Bit-identical to Pillow's
 * src/libImaging/Resample.c for 8-bit channels: horizontal pass into a uint8 intermediate image,
 * then the vertical pass, 22-bit fixed-point coefficients.  1, 3 or 4 interleaved channels, every
 * channel filtered independently (Pillow modes L / RGB; RGBA differs: Pillow premultiplies alpha). */
typedef struct lars_resize_plan {
  int32_t in_h, in_w, out_h, out_w, channels;
  int32_t need_h, need_v;         /* which passes run (Resample.c need_horizontal / need_vertical)   */
  int32_t ksize_h, ksize_v;       /* taps per output column / row                                    */
  int32_t row_first, row_count;   /* source rows the horizontal pass produces (ybox_first .. last)   */
  int32_t xo_tile, plane_words, out_pitch; /* launch geometry of the horizontal pass                 */
  int32_t groups_h, groups_v;     /* 4-tap coefficient groups per output column / row                */
  int32_t mma_ksteps;             /* > 0: tensor-core horizontal pass available (k-steps of 32 pixels) */
  int32_t mma_plane_words;        /* its shared words per row per channel plane                      */
  uint64_t mma_table_offset;      /* byte offset of its block-start / fragment tables in the block   */
  uint64_t table_bytes;           /* size of the coefficient block (host and device copies)          */
  uint64_t temp_frame_bytes;      /* intermediate image bytes per frame (multiple of 16); 0 if unused */
} lars_resize_plan;
/* Host only: geometry of a resize.  LARS_ERR_UNSUPPORTED if a single filter window does not fit
 * in shared memory (down-scaling by more than ~250x). */
int lars_resize_plan_lanczos(int32_t in_h, int32_t in_w, int32_t out_h, int32_t out_w, int32_t channels,
                             lars_resize_plan* plan);
/* Host only: fills plan->table_bytes bytes of HOST memory with the coefficient block
 * (Resample.c precompute_coeffs + normalize_coeffs_8bpc, double precision + libm sin);
 * the caller copies it to the device once per geometry. */
int lars_resize_tables_lanczos(const lars_resize_plan* plan, void* tables_host);
/* Device: n_frames frames [in_h][in_w][channels] -> [out_h][out_w][channels].  tables_dev is the
 * device copy of the coefficient block (4-byte aligned); temp holds n_frames * temp_frame_bytes
 * bytes (may be NULL when temp_frame_bytes == 0).  Frame strides in bytes. */
int lars_resize_lanczos_u8(const lars_resize_plan* plan, const void* tables_dev, const uint8_t* src,
                           int64_t src_frame_stride, int32_t n_frames, uint8_t* dst, int64_t dst_frame_stride,
                           void* temp, size_t temp_bytes, void* stream);
/* RGBA frames are resized with premultiplied alpha, as Pillow does (Image.resize: convert("RGBa"), resize,
 * convert("RGBA"); libImaging/Convert.c rgbA2rgba / rgba2rgbA): this call converts a batch of RGBA frames in place,
 * premultiply != 0 before lars_resize_lanczos_u8 and premultiply == 0 on its output. */
int lars_rgba_alpha_u8(uint8_t* data, int32_t n_frames, int64_t n_pixels, int64_t frame_stride, int32_t premultiply,
                       void* stream);

/* ---- ingest ---------------------------------------------------
 * Host-only TIFF 6.0 / BigTIFF reader: 8- or 16-bit unsigned (or 32-bit float) samples, 1 / 3 / 4 samples per pixel, chunky
 * or planar (band-interleaved, read back interleaved), either byte order; strips or tiles; uncompressed, LZW (5), Deflate (8 / 32946; needs the
 * system zlib at run time) or PackBits (32773), predictor 1 or 2.  Replaces PIL.Image.open +
 * np.array for such files (process-images.py:183-193; backend-process.py:52; process-ndvi.py:18;
 * process-rgn.py:18) and adds the true 16-bit path Pillow cannot deliver (it opens 16-bit RGB as
 * 8-bit): the strips / tiles of a (memory-mapped) file go straight into the caller's pinned HWC
 * buffer.  lars_tiff_read_region reads only the chunks that touch a rectangle -- the tiles or row
 * bands one rank owns of an orthomosaic (BASELINE config 4) -- with n_threads host threads decoding
 * independent chunks side by side.  Anything else (JPEG-in-TIFF, 1-bit, WhiteIsZero, YCbCr, float predictor)
 * returns LARS_ERR_UNSUPPORTED and the caller decodes with Pillow. */
typedef struct lars_tiff_info {
  int32_t width, height, samples_per_pixel, bits_per_sample;
  int32_t big_endian, compression, planar_config, photometric, sample_format;
  int32_t rows_per_strip, n_strips;              /* tiled files: tile_length, number of tiles      */
  int32_t strip_offsets_type, strip_counts_type; /* TIFF field types (3 SHORT, 4 LONG, 16 LONG8)   */
  int32_t predictor;                             /* 1 = none, 2 = horizontal differencing          */
  uint64_t strip_offsets_pos, strip_counts_pos;  /* file positions of the strip / tile arrays      */
  uint64_t frame_bytes;                          /* height * width * samples * bytes per sample    */
  int32_t tile_width, tile_length;               /* 0 = the image is stored in strips              */
  int32_t tiles_across, tiles_down;
  int32_t bigtiff, reserved;
} lars_tiff_info;
int lars_tiff_probe(const void* file, size_t file_bytes, lars_tiff_info* info);
/* dst receives height x width x samples, little-endian samples, rows contiguous (one thread). */
int lars_tiff_read(const void* file, size_t file_bytes, const lars_tiff_info* info, void* dst, size_t dst_bytes);
/* dst receives rows [row0, row1) x columns [col0, col1) as one contiguous block. */
int lars_tiff_read_region(const void* file, size_t file_bytes, const lars_tiff_info* info, int32_t row0,
                          int32_t row1, int32_t col0, int32_t col1, void* dst, size_t dst_bytes,
                          int32_t n_threads);

/* Device-side decode of LZW TIFF strips: the compressed file bytes go to the GPU as they are and one warp
 * decodes each strip straight into the frame slot (thousands of independent streams per batch).
 * lars_tiff_lzw_chunks (host) fills one descriptor per strip of a file lars_tiff_probe accepted -- offsets
 * relative to the start of the file and of the frame; the caller adds the file's / frame's position in its
 * batch buffers -- and fails with LARS_ERR_UNSUPPORTED unless the file is LZW-compressed, stored in strips of
 * at most 1 MB decoded.  lars_lzw_decode_device launches the decode; `counters` is two zeroed uint32 on the
 * device: [0] receives the number of corrupt / short strips, [1] is the work counter.
 * lars_tiff_post_device undoes the differencing predictor and big-endian 16-bit samples in place. */
typedef struct lars_lzw_chunk {
  uint64_t src_offset;   /* compressed bytes: offset from `src`                       */
  uint64_t dst_offset;   /* decoded bytes: offset from `dst`                           */
  uint32_t src_bytes;
  uint32_t dst_bytes;    /* bytes this strip must produce (rows x width x samples)     */
} lars_lzw_chunk;
int lars_tiff_lzw_chunks(const void* file, size_t file_bytes, const lars_tiff_info* info, lars_lzw_chunk* chunks,
                         int32_t max_chunks);
int lars_lzw_decode_device(const uint8_t* src, const lars_lzw_chunk* chunks, int32_t n_chunks, uint8_t* dst,
                           uint32_t* counters, void* stream);
/* The same for Deflate-compressed strips (pinned against zlib on the CPU, against the host reader on B200) --
 * lars_tiff_deflate_chunks fills the table, lars_inflate_decode_device runs one warp per zlib stream.  The Adler-32
 * trailers are verified on the device (checksum accumulated as the output leaves shared memory). */
int lars_tiff_deflate_chunks(const void* file, size_t file_bytes, const lars_tiff_info* info, lars_lzw_chunk* chunks,
                             int32_t max_chunks);
int lars_inflate_decode_device(const uint8_t* src, const lars_lzw_chunk* chunks, int32_t n_chunks, uint8_t* dst,
                               uint32_t* counters, void* stream);
/* Strips / tiles that were decoded into scratch slots of slot_bytes each (rows of slot_row_bytes) are moved to their
 * place inside one frame, e.g. the row band of a tiled mosaic a rank owns.  table (device, int64 [n_chunks][5]):
 * first slot row, number of rows, first frame row, byte column in the frame, bytes per row. */
int lars_untile_device(const uint8_t* scratch, int64_t slot_bytes, int64_t slot_row_bytes, const int64_t* table,
                       int32_t n_chunks, uint8_t* dst, int64_t dst_row_bytes, void* stream);
int lars_tiff_post_device(uint8_t* dst, int32_t n_frames, int64_t frame_stride, int32_t rows, int32_t width,
                          int32_t samples_per_pixel, int32_t sample_bytes, int32_t predictor, int32_t swap16,
                          void* stream);

/* Host-only PNG reader: non-interlaced grayscale / RGB / RGBA, 8- or 16-bit samples (BASELINE config 1
 * is a uint8 RGNir PNG).  One call per frame -- chunk walk, one zlib inflate (bound at run time), the
 * five row filters -- into the caller's pinned HWC buffer; 16-bit samples arrive as little-endian
 * uint16 (Pillow reduces 16-bit RGB to 8 bits).  Palette, gray + alpha, sub-byte depths, tRNS and Adam7
 * return LARS_ERR_UNSUPPORTED and the caller decodes with Pillow.  Same reference call sites as the
 * TIFF reader above. */
typedef struct lars_png_info {
  int32_t width, height, channels, bit_depth;
  int32_t color_type, interlace, n_idat, reserved;
  uint64_t idat_bytes;                           /* compressed bytes over all IDAT chunks           */
  uint64_t frame_bytes;                          /* height * width * channels * bytes per sample    */
} lars_png_info;
int lars_png_probe(const void* file, size_t file_bytes, lars_png_info* info);
/* dst receives height x width x channels, little-endian samples, rows contiguous. */
int lars_png_read(const void* file, size_t file_bytes, const lars_png_info* info, void* dst, size_t dst_bytes);

#ifdef __cplusplus
}
#endif
#endif /* LARS_B200_H_ */
