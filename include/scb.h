/* scb.h -- C ABI of the B200-native seamlessClone(NORMAL_CLONE) hot path.
 *
 * Plain C, POD only: opaque handles, raw pointers + (rows, cols, channels, stride), int status
 * codes.  No OpenCV, torch or C++ types cross this boundary, so any host language can bind it
 * (ctypes / cgo / JNI / N-API); see INTEGRATION.md for the stubs.
 *
 * What each entry point replaces in the reference (paths under /root/reference/seamlessClone-CUDA/):
 *
 *   scb_create / scb_destroy / scb_sync
 *       seamlessClone_imp_create_instance / _destroy / _sync      seamlessClone_imp.cu:239-263, 354-370
 *       (one context = one device + one stream + grow-only workspace + table cache;
 *        the reference never calls cudaSetDevice -- every entry point here does)
 *   scb_plan_create / scb_plan_destroy
 *       SeamlessClone::init_resize -> initMask (ring-zero, bbox, 3x erode, leftTop)
 *                                                                  seamlessClone_imp.cpp:978-1116
 *       plus the DST tables the reference rebuilds on every call (initDSTMatrix_kernel :569-603)
 *   scb_plan_execute
 *       SeamlessClone::seamlessCloneGPU -> run()                   seamlessClone_imp.cpp:430-486, 2105-2135
 *   scb_seamless_clone
 *       cv::seamlessClone(src, dst, mask, p, blend, NORMAL_CLONE)  call sites seamlessClone-OpenCV/seamlessClone_OpenCV.cpp:104,110
 *       == seamlessClone_imp_run without its double execution      seamlessClone_imp.cu:265-352
 *   my_seamlessclone_api_imp_*
 *       the four extern "C" names of seamlessclone_cuda.h:4-63 (which return cv::Mat by value and
 *       so are not a C ABI); same names and argument meaning, POD image views instead of cv::Mat*
 *   scb_plan_execute_graph
 *       the same run() captured once and replayed as a single CUDA graph launch (no reference counterpart;
 *       the reference re-launches ~47 kernels and syncs twice per call)
 *   scb_clone_batch
 *       a loop over seamlessClone_imp_run (seamlessClone_imp.cu:265-352), pipelined over streams
 *   scb_plan_get_intermediate
 *       the SCDEBUG YAML dumps compared by compare/vs.py:12-34 (g*.yml vs OpenCV's mod_diff*.yml)
 */
#ifndef SCB_H_
#define SCB_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SCB_VERSION 100

/* status codes */
enum {
    SCB_OK = 0,
    SCB_ERR_INVALID_ARGUMENT = 1,  /* null pointers, wrong channel counts, sizes that disagree   */
    SCB_ERR_ROI_OUT_OF_BOUNDS = 2, /* OpenCV: (-215) assertion on roi inside dst                 */
    SCB_ERR_UNSUPPORTED = 3,       /* unknown clone flags, ROI side > 8194, bbox smaller than 3  */
    SCB_ERR_CUDA = 4,              /* a CUDA call failed; scb_last_error has the text            */
    SCB_ERR_NO_DEVICE = 5,
    SCB_ERR_OUT_OF_MEMORY = 6
};

/* cv::seamlessClone flags (photo.hpp).  NORMAL_CLONE is the hot path (vectorised integer stencil); MIXED_CLONE and
 * MONOCHROME_TRANSFER run the same solver behind a per-pixel float stencil with OpenCV's gradient selection; the _WIDE
 * variants (OpenCV >= 4.11) place the centre of src, instead of the centre of the mask bounding box, at p. */
enum {
    SCB_NORMAL_CLONE = 1, SCB_MIXED_CLONE = 2, SCB_MONOCHROME_TRANSFER = 3,
    SCB_NORMAL_CLONE_WIDE = 9, SCB_MIXED_CLONE_WIDE = 10, SCB_MONOCHROME_TRANSFER_WIDE = 11
};

/* where the pixel buffers of a call live */
enum { SCB_MEM_HOST = 0, SCB_MEM_DEVICE = 1 };

/* scb_plan_execute flags */
enum {
    SCB_EXEC_DEFAULT = 0,
    SCB_EXEC_BLEND_PREFILLED = 1 /* blend already holds a copy of dst (or aliases it): write the ROI interior only */
};

/* transform engines.  TRI (default; AUTO resolves to it): Bluestein FFT passes along x, and along y the
 * mathematically identical tridiagonal (Thomas) solve of every spectral column, with the lowest 32 x 32
 * frequencies in float64 against OpenCV's float32 denominators -- half the transform work of FFT.
 * FFT: the Bluestein shared-memory FFT engine on both axes (also what the sharded entry points run).  TC: dense
 * sine-basis contraction on the tensor cores (tcgen05, 3xTF32, even/odd fold) for line lengths 16..4096 -- opt-in:
 * its FP32 accumulation error (~1e-5 relative at K ~ 900) is inside the 1e-4 bar for float intermediates but
 * costs exactly-matching bytes at some shapes.  The environment variable SCB_ENGINE=tri|tc|fft|i8 sets the default.
 * I8: the tridiagonal solve along y with, along x, the DST as an EXACT integer contraction on the INT8 tensor cores
 * (tcgen05.mma.kind::i8 over balanced base-256 digit planes, scb_i8.h): no rounding before the final float, the exact
 * float64 low-frequency row sums fall out of the same accumulators.  AUTO resolves to I8 for line lengths 64..8192
 * and to TRI otherwise. */
enum { SCB_ENGINE_AUTO = 0, SCB_ENGINE_FFT = 1, SCB_ENGINE_TC = 2, SCB_ENGINE_TRI = 3, SCB_ENGINE_I8 = 4 };

/* scb_plan_get_intermediate selectors; all float32, planar [3][rows][cols] */
enum {
    SCB_INT_GRADIENT_X = 0, /* [3][h][w]    blended forward-difference gradient                  */
    SCB_INT_GRADIENT_Y = 1, /* [3][h][w]                                                         */
    SCB_INT_RHS = 2,        /* [3][ny][nx]  divergence minus Dirichlet boundary (OpenCV mod_diff) */
    SCB_INT_SPECTRUM = 3,   /* [3][nx][ny]  forward 2-D DST, TRANSPOSED, before the division (FFT / TC engines only) */
    SCB_INT_SOLVED = 4,     /* [3][ny][nx]  solved field before clamp / truncation                */
    SCB_INT_ERODED_MASK = 5 /* [1][h][w]    eroded mask as float 0..255                           */
};

typedef struct scb_context scb_context;
typedef struct scb_plan scb_plan;

/* borrowed view of an 8-bit image; stride in bytes; channels 1 (mask) or 3 (BGR interleaved) */
typedef struct scb_image {
    void* data;
    int32_t rows, cols, channels;
    int64_t stride;
} scb_image;

typedef struct scb_geometry {
    int32_t x, y, w, h;  /* bounding box of the ring-zeroed mask, in src/mask coordinates */
    int32_t rx, ry;      /* ROI origin in dst: (px - w/2, py - h/2)                       */
    int32_t nx, ny;      /* unknowns per channel: w-2, h-2                                */
    int32_t empty;       /* mask had no interior non-zero pixel: blend = dst              */
    int32_t log2m_x, log2m_y; /* convolution lengths chosen for the two DST axes           */
} scb_geometry;

/* ---- context ---- */
/* external_stream: a cudaStream_t/CUstream to adopt (not owned), or NULL to create one. */
int scb_create(int device, void* external_stream, scb_context** out);
int scb_destroy(scb_context* ctx);
int scb_sync(scb_context* ctx);
void* scb_stream(scb_context* ctx);
const char* scb_last_error(const scb_context* ctx); /* ctx may be NULL: error of the last failed scb_create on this thread */
const char* scb_status_string(int status);
uint64_t scb_kernel_launches(const scb_context* ctx); /* kernels this library has launched on ctx so far */
int scb_device_count(void);
/* 16 hex digits: hash of the sources the library was built from (the test fixture compares it with the tree and rebuilds a stale library) */
const char* scb_source_hash(void);
/* "name=value ..." of the kernel generations in force (defaults, or the SCB_* environment overrides used for A/B measurements);
 * bench.py prints it with every line so that a number always names the kernels that produced it. */
const char* scb_kernel_variants(void);
/* Chooses the DST engine of plans created afterwards (SCB_ENGINE_*).  The reference makes the same choice at
 * compile time: SC_FFT_ENABLE, seamlessClone_imp.h:15 (cuFFT solver vs cuBLAS sine-basis solver). */
int scb_set_engine(scb_context* ctx, int engine);
/* Tridiagonal engine: which axis carries the FFT passes (the other one is solved as tridiagonal systems).
 * -1 = chosen per plan by transform cost (default), 0 = along x (rows), 1 = along y (columns).  Results agree to
 * rounding; the choice only moves time.  (No reference counterpart: its solver is fixed at compile time.) */
int scb_set_orientation(scb_context* ctx, int orientation);
/* Unit check of one tensor-core pass: random lines of length n against float64 direct sums; max error relative
 * to the largest output of the line.  (No reference counterpart: SC_Test, seamlessClone_imp.cpp:532-554, is dead code.) */
int scb_tc_selftest(scb_context* ctx, int n, int lines, int transposed, double* max_rel_err);

/* pinned host memory helpers (cudaMallocHost / cudaFreeHost) */
int scb_host_alloc(void** out, size_t bytes);
int scb_host_free(void* p);

/* ---- plan: everything that depends on (mask, sizes, p) only ---- */
int scb_plan_create(scb_context* ctx, const scb_image* mask, int mask_mem_kind, int src_rows, int src_cols,
                    int dst_rows, int dst_cols, int px, int py, scb_plan** out);
/* The same with cv::seamlessClone's `flags` (SCB_*_CLONE*); scb_plan_create == NORMAL_CLONE. */
int scb_plan_create_ex(scb_context* ctx, const scb_image* mask, int mask_mem_kind, int src_rows, int src_cols,
                       int dst_rows, int dst_cols, int px, int py, int clone_flags, scb_plan** out);
int scb_plan_destroy(scb_plan* plan);
int scb_plan_geometry(const scb_plan* plan, scb_geometry* out);
int scb_plan_engine(const scb_plan* plan); /* SCB_ENGINE_TRI, SCB_ENGINE_FFT or SCB_ENGINE_TC */
/* Asynchronous on the context stream for SCB_MEM_DEVICE; for SCB_MEM_HOST returns when blend is complete. */
int scb_plan_execute(scb_plan* plan, const scb_image* src, const scb_image* dst, scb_image* blend, int mem_kind, int exec_flags);
/* Same call with CUDA events between the stages (on the context stream); returns after a stream sync.
 * stage_ms[7] = { input copies, RHS stencil, low-frequency refinement, rows forward, columns, rows inverse, output copy }.
 * (The reference times its whole run() with one event pair: seamlessClone_imp.cu:281-349.) */
int scb_plan_execute_timed(scb_plan* plan, const scb_image* src, const scb_image* dst, scb_image* blend, int mem_kind, int exec_flags, float* stage_ms);
/* The same, plus the kernels of the INT8 tensor-core passes on their own event pairs:
 * i8_ms[5] = { digitise forward, GEMM forward, digitise inverse, GEMM inverse, compose } (zeros on another engine). */
int scb_plan_execute_timed_i8(scb_plan* plan, const scb_image* src, const scb_image* dst, scb_image* blend, int mem_kind, int exec_flags, float* stage_ms,
                              float* i8_ms);
/* DEVICE-resident execute replayed as one CUDA graph launch (captured on first use; re-captured when a
 * pointer, stride or the workspace changes): the per-frame call of a fixed-mask stream (BASELINE cfg5).
 * Replaces the reference's ~47 launches and 2 host syncs per frame (seamlessClone_imp.cpp:2105-2135). */
int scb_plan_execute_graph(scb_plan* plan, const scb_image* src, const scb_image* dst, scb_image* blend, int exec_flags);
int scb_plan_set_debug(scb_plan* plan, int on);
int scb_plan_get_intermediate(scb_plan* plan, int which, float* out_host, size_t capacity_floats, size_t* written);

/* ---- one-shot, OpenCV-shaped ---- */
int scb_seamless_clone(scb_context* ctx, const scb_image* src, const scb_image* dst, const scb_image* mask,
                       int px, int py, scb_image* blend, int clone_flags, int mem_kind);

/* scb_seamless_clone keeps the plans of its last few (mask, sizes, p, flags) in the context, keyed by a 64-bit hash of the HOST
 * mask's bytes: a caller that passes the same mask again (video: fixed mask, new frames) skips the mask upload, the bounding-box
 * round trip, the erosion and the table lookups.  SCB_PLAN_CACHE=n sets the capacity (default 4, 0 disables).
 * (Replaces the reference's per-call initMask with its blocking D2H, seamlessClone_imp.cpp:978-1071.) */
int scb_plan_cache_stats(const scb_context* ctx, uint64_t* hits, uint64_t* misses);

/* ---- batch of independent jobs (all HOST or all DEVICE) ---- */
typedef struct scb_job {
    scb_image src, dst, mask, blend;
    int32_t px, py;
    int32_t status; /* out */
    int32_t flags;  /* cv::seamlessClone flags of this job; 0 = SCB_NORMAL_CLONE */
} scb_job;
int scb_clone_batch(scb_context* ctx, scb_job* jobs, int n_jobs, int mem_kind);

/* ---- row/column-sharded single solve (one rank per GPU; the caller exchanges At / Ct between the passes) ---- */
/* Runs pass A (stencil + row DST) for interior rows [y0, y1) into `at` ([3][nx][ny] device buffer). */
int scb_plan_rows_forward(scb_plan* plan, const scb_image* src, const scb_image* dst, int mem_kind, int y0, int y1, float* at_dev, double* lowrows_dev);
/* Runs pass B for columns [x0, x1): reads at_dev ([3][nx][ny]), writes ct_dev ([3][ny][nx]). lowspec_dev may be NULL. */
int scb_plan_cols(scb_plan* plan, int x0, int x1, const float* at_dev, float* ct_dev, const float* lowspec_dev);
/* Reduces the low-frequency row sums ([3][lowkx][ny], complete over all rows) into lowspec ([3][lowkx][lowky]). */
int scb_plan_lowfreq_finish(scb_plan* plan, const double* lowrows_dev, float* lowspec_dev);
/* Runs pass C for interior rows [y0, y1): reads ct_dev, writes the blend interior rows. */
int scb_plan_rows_inverse(scb_plan* plan, const float* ct_dev, scb_image* blend, int mem_kind, int y0, int y1);
int scb_plan_lowk(const scb_plan* plan, int* lowkx, int* lowky);

/* ---- the same on the tridiagonal engine: NO transpose exchange.  The column solve is a partitioned (SPIKE) Thomas solve whose
 * segments follow the row shards, so a rank keeps its rows from the stencil to the composed bytes; the ranks only combine
 *   ends32 / ends64 : the two end values of every segment's local solution (3 x 16 x 2 x nx floats, ~1.5 MB at 8K), and
 *   w               : the 32 x 32 low-frequency projections (3 x 32 x 32 doubles),
 * each filled (zero elsewhere) by scb_plan_tri_forward so that one all-reduce(SUM) per buffer completes them.
 * Rank r owns segments [seg0, seg1) = interior rows [seg0 * seg_len, min(ny, seg1 * seg_len)).  The plan's workspace carries
 * the field between the two calls: run nothing else on the plan in between.  (The reference has no multi-GPU path.) */
int scb_plan_tri_layout(const scb_plan* plan, int* seg_len, int* n_segs, size_t* ends32_floats, size_t* ends64_doubles, size_t* w_doubles);
int scb_plan_tri_forward(scb_plan* plan, const scb_image* src, const scb_image* dst, int mem_kind, int seg0, int seg1, float* ends32_dev, double* ends64_dev, double* w_dev);
int scb_plan_tri_finish(scb_plan* plan, scb_image* blend, int mem_kind, int seg0, int seg1, const float* ends32_dev, const double* ends64_dev, const double* w_dev);
/* The same with the low-frequency projections left as `w_slots` consecutive per-rank partials ([w_slots][w_doubles], rank r's
 * scb_plan_tri_forward writes slot r and the others stay zero): every buffer of the exchange then has DISJOINT supports across
 * the ranks, so ends32 | ends64 | w can be packed and combined with ONE integer all-reduce (x + 0 is exact for any bit pattern). */
int scb_plan_tri_finish_slots(scb_plan* plan, scb_image* blend, int mem_kind, int seg0, int seg1, const float* ends32_dev, const double* ends64_dev,
                              const double* w_dev, int w_slots);

/* ---- the reference's four entry points (seamlessclone_cuda.h:4-63), POD views instead of cv::Mat* ---- */
void* my_seamlessclone_api_imp_create_instance(int gpu_id);
/* face = patch (src), body = dst; blend_out must be a caller-allocated image of dst's size. Returns a status code. */
int my_seamlessclone_api_imp_run(void* instance_ptr, const scb_image* face, const scb_image* body, const scb_image* mask,
                                 int centerX, int centerY, int gpu_id, int bSync, scb_image* blend_out);
void my_seamlessclone_api_imp_destroy(void* instance_ptr);
void my_seamlessclone_api_imp_sync(void* instance_ptr);

#ifdef __cplusplus
}
#endif
#endif /* SCB_H_ */
