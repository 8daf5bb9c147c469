// scb_api.cu -- host driver + C ABI (include/scb.h) of the B200-native NORMAL_CLONE hot path.
//
// Mirrors, and replaces, the reference's instance layer:
//   /root/reference/seamlessClone-CUDA/seamlessClone_imp.cu:239-370   create / run / destroy / sync
//   /root/reference/seamlessClone-CUDA/seamlessClone_imp.cpp:430-486  seamlessCloneGPU (upload, run, download, scatter)
//   /root/reference/seamlessClone-CUDA/seamlessClone_imp.cpp:978-1116 initMask / init_resize
// with the differences SURVEY.md 8b asks for: POD-only ABI, status codes instead of assert/exit,
// cudaSetDevice on every entry, ROI-only transfers, no host sync inside the hot path other than the
// one that reports a HOST-resident result, dst never modified, mask never modified.
#include <chrono>
#include <climits>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <atomic>
#include <condition_variable>
#include <map>
#include <mutex>
#include <functional>
#include <string>
#include <thread>
#include <vector>

#if defined(__SSE2__) && !defined(SCB_EMU)
#include <emmintrin.h>
#endif
#if defined(__linux__) && !defined(SCB_EMU)
#include <pthread.h>
#include <sched.h>
#endif

#if !defined(SCB_EMU) && !defined(SCB_NO_NVTX)
#include <nvtx3/nvToolsExt.h>  // header-only; ranges cost nothing unless a profiler is attached
#define SCB_NVTX_PUSH(name) nvtxRangePushA(name)
#define SCB_NVTX_POP() nvtxRangePop()
#else
#define SCB_NVTX_PUSH(name) ((void)0)
#define SCB_NVTX_POP() ((void)0)
#endif

#include "../../include/scb.h"
#include "scb_kernels.cuh"
#include "scb_kernels3.cuh"
#include "scb_platform.h"
#include "scb_tables.h"
#include "scb_i8.h"
#include "scb_tc.cuh"
#include "scb_tri.cuh"

using namespace scb;

// One NVTX range per stage of a clone (the reference has CUDA-event timing only, imp.cu:281-349).  The ranges
// bracket the ENQUEUE of a stage on the host; Nsight Systems projects them onto the kernels they launched.
struct NvtxRange {
    explicit NvtxRange(const char* name) { SCB_NVTX_PUSH(name); }
    ~NvtxRange() { SCB_NVTX_POP(); }
    NvtxRange(const NvtxRange&) = delete;
    NvtxRange& operator=(const NvtxRange&) = delete;
};

#ifdef SCB_EMU
static inline cudaError_t scbMallocAsync(void** p, size_t n, cudaStream_t) { return emu_malloc(p, n); }
typedef void* cudaMemPool_t;
static inline cudaError_t scbMallocPoolAsync(void** p, size_t n, cudaMemPool_t, cudaStream_t) { return emu_malloc(p, n); }
static inline cudaError_t scbFreeAsync(void* p, cudaStream_t) { std::free(p); return cudaSuccess; }
#else
static inline cudaError_t scbMallocAsync(void** p, size_t n, cudaStream_t s) { return cudaMallocAsync(p, n, s); }
static inline cudaError_t scbMallocPoolAsync(void** p, size_t n, cudaMemPool_t pool, cudaStream_t s) {
    return pool ? cudaMallocFromPoolAsync(p, n, pool, s) : cudaMallocAsync(p, n, s);
}
static inline cudaError_t scbFreeAsync(void* p, cudaStream_t s) { return cudaFreeAsync(p, s); }
#endif

// ------------------------------------------------------------------------------------------------
// Bookkeeping of one cached table: the context's table caches are bounded (byte budget, LRU) and an entry is only evicted
// while no live plan references it.
struct CacheMeta {
    size_t bytes = 0;
    int refs = 0;
    uint64_t last_use = 0;
};

struct DevLenTab {
    LenTabDev dev{};
    void* block = nullptr;
    CacheMeta meta;
};
struct DevFilter {
    float* d = nullptr;
    CacheMeta meta;
};

// A lane is one in-order pipeline of the context: a stream, the side stream of the low-frequency
// refinement, and a grow-only workspace arena.  Lane 0 is the context's main stream (the only one the
// single-job entry points use); scb_clone_batch spreads independent jobs over all lanes so that one
// job's transfers overlap another's kernels and small jobs share the 148 SMs.
static const int kMaxBands = 4;

struct Lane {
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    cudaStream_t side = nullptr;              // low-frequency refinement runs here, beside pass A
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    cudaStream_t copy = nullptr;              // HOST calls: banded H2D / D2H pipelined against the row passes
    cudaEvent_t ev_copy = nullptr;            // DEVICE calls: dst -> blend copy done (runs on `side`, beside the passes)
    cudaEvent_t ev_band[kMaxBands] = {}, ev_out[kMaxBands] = {};
    char* ws = nullptr;
    size_t ws_cap = 0;
    uint64_t ws_epoch = 0;                    // bumps when the arena moves (captured graphs hold its addresses)
};
static const int kMaxLanes = 8;
static const int kDefaultLanes = 8;  // SCB_LANES=1..8 overrides (tuning; measured with 4 submit threads: profiles/r2_batch_submit_sweep.txt)

struct DevTcTab {
    TcTabDev dev{};
    void* block = nullptr;
    CacheMeta meta;
#ifndef SCB_EMU
    CUtensorMap map;
#endif
};

struct DevTriTab {
    TriTabDev dev{};
    void* block = nullptr;
    CacheMeta meta;
};

struct DevI8Tab {  // digit planes of the folded sine basis of one line length (scb_i8.h)
    I8Geom g{};
    signed char* basis = nullptr;
    CacheMeta meta;
};

class HostPool;
struct scb_context;
static void host_pool_free(scb_context* c);  // defined next to the class

struct scb_context {
    int device = 0;
    HostPool* pool = nullptr;           // helper threads of the host-side dst -> blend copy; one pool per context (created on first use)
    cudaMemPool_t mem_pool = nullptr;   // private stream-ordered pool of the plans' eroded masks / staged masks (never the device's default pool)
    std::map<std::pair<int, int>, DevTriTab> tritabs;  // keyed by ROI (w, h): LU factors of the tridiagonal engine
    int engine = SCB_ENGINE_AUTO;
    int orientation = -1;               // tridiagonal engine: -1 cost model, 0 FFT passes along x, 1 along y (scb_set_orientation)
    std::map<int, DevTcTab> tctabs;     // keyed by n (tensor-core engine: split sine bases + tensor maps)
    std::map<int, DevI8Tab> i8tabs;     // keyed by n (exact INT8 tensor-core engine: digit planes of the basis)
    Lane lanes[kMaxLanes];
    int n_lanes = 0;
    cudaStream_t prep = nullptr;  // scb_clone_batch: mask uploads + bounding boxes of the next chunk
    std::string err;
    std::atomic<uint64_t> launches{0};  // scb_clone_batch plans the next chunk on a helper thread
    std::mutex err_mu;
    std::map<int, DevLenTab> lentabs;   // keyed by n
    std::map<int, DevFilter> filters;   // keyed by ROI extent
    std::mutex table_mu;                // the table caches: scb_clone_batch plans on a helper thread while the caller's thread destroys plans
    size_t table_bytes = 0, table_budget = (size_t)1 << 30;  // SCB_TABLE_BUDGET_MB overrides (a 4K ROI size costs ~10 MB of tables)
    uint64_t use_clock = 0;
    // scb_seamless_clone's plan cache: (hash of the mask bytes, sizes, p, flags, engine) -> plan.  A video-style caller that passes
    // the same mask every frame skips the mask upload, the bounding-box round trip, the erosion and the table lookups.
    struct CachedPlan {
        uint64_t hash = 0;
        int key[11] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
        scb_plan* plan = nullptr;
        uint64_t last_use = 0;
    };
    std::vector<CachedPlan> plan_cache;
    int plan_cache_cap = 4;  // SCB_PLAN_CACHE=n overrides, 0 disables
    uint64_t plan_hits = 0, plan_misses = 0;
    int* bbox_dev = nullptr;     // [slots][kBboxInts]: min x, min y, max x, max y, "mask has grey values", 3 unused
    int* bbox_pinned = nullptr;  // [0..kBboxInts) init pattern, then [slots][kBboxInts] results
    int bbox_slots = 0;
    int sm_count = 148;
    int max_smem = 232448;
};

struct GraphKey {
    const void *src = nullptr, *dst = nullptr, *blend = nullptr;
    int64_t s_stride = 0, d_stride = 0, b_stride = 0;
    int flags = 0;
    uint64_t ws_epoch = 0;
    bool operator==(const GraphKey& o) const {
        return src == o.src && dst == o.dst && blend == o.blend && s_stride == o.s_stride && d_stride == o.d_stride && b_stride == o.b_stride &&
               flags == o.flags && ws_epoch == o.ws_epoch;
    }
};

struct scb_plan {
    scb_context* ctx = nullptr;
    Lane* lane = nullptr;
#ifndef SCB_EMU
    cudaGraphExec_t graph_exec = nullptr;
#endif
    GraphKey graph_key;
    uint64_t graph_kernels = 0;
    unsigned char* mask_stage = nullptr;  // batch: staged host mask between plan_begin and plan_finish
    scb_geometry g{};
    int src_rows = 0, src_cols = 0, dst_rows = 0, dst_cols = 0;
    unsigned char* E = nullptr;
    long long e_pitch = 0;
    LenTabDev tx{}, ty{};
    bool use_tc = false;                // tensor-core dense engine (scb_tc.cuh) instead of the FFT engine
    bool use_tri = false;               // tridiagonal column solve (scb_tri.cuh) instead of the column FFT pass
    bool use_i8 = false;                // with use_tri: the passes along x as exact INT8 tensor-core contractions (scb_i8.h) instead of FFTs
    const DevI8Tab* i8x = nullptr;
    std::vector<CacheMeta*> table_refs; // cached tables this plan uses (released in scb_plan_destroy)
    bool grey_mask = false;             // the mask holds values other than 0 / 255 inside its ring: the right-hand side is not integer valued
    bool swap = false;                  // tridiagonal engine: FFT passes along y and the tridiagonal solve along x (choose_swap)
    TriTabDev tri{};                    // LU factors for the chosen orientation
    const DevTcTab *ttx = nullptr, *tty = nullptr;
    const float* fx = nullptr;
    const float* fy = nullptr;
    int lowkx = 0, lowky = 0;
    int mode = SCB_NORMAL_CLONE;        // gradient selection: NORMAL_CLONE, MIXED_CLONE or MONOCHROME_TRANSFER
    bool wide = false;                  // *_WIDE flags: src (not the mask bounding box) is centred at p
    bool debug = false;
    float *dbg_vx = nullptr, *dbg_vy = nullptr, *dbg_rhs = nullptr, *dbg_spec = nullptr, *dbg_u = nullptr;
};

static thread_local std::string g_create_error;

static int fail(scb_context* c, int code, const std::string& msg) {
    if (c) {
        std::lock_guard<std::mutex> lk(c->err_mu);
        c->err = msg;
    } else {
        g_create_error = msg;
    }
    return code;
}
#define SCB_CUDA(c, expr)                                                                               \
    do {                                                                                                \
        cudaError_t e__ = (expr);                                                                       \
        if (e__ != cudaSuccess)                                                                         \
            return fail(c, e__ == cudaErrorMemoryAllocation ? SCB_ERR_OUT_OF_MEMORY : SCB_ERR_CUDA,     \
                        std::string(#expr) + ": " + cudaGetErrorString(e__));                           \
    } while (0)

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

static int ensure_ws(scb_context* c, Lane* l, size_t bytes) {
    if (bytes <= l->ws_cap) return SCB_OK;
    SCB_CUDA(c, cudaStreamSynchronize(l->stream));
    if (l->ws) SCB_CUDA(c, cudaFree(l->ws));
    l->ws = nullptr;
    l->ws_cap = 0;
    l->ws_epoch++;
    size_t want = align_up(bytes + bytes / 4, 1 << 20);
    SCB_CUDA(c, cudaMalloc(&l->ws, want));
    l->ws_cap = want;
    return SCB_OK;
}

static cudaError_t lane_create(Lane* l, cudaStream_t adopt) {
    cudaError_t e;
    if (adopt) {
        l->stream = adopt;
    } else {
        if ((e = cudaStreamCreateWithFlags(&l->stream, cudaStreamNonBlocking)) != cudaSuccess) return e;
        l->own_stream = true;
    }
    if ((e = cudaStreamCreateWithFlags(&l->side, cudaStreamNonBlocking)) != cudaSuccess) return e;
    if ((e = cudaEventCreateWithFlags(&l->ev_fork, cudaEventDisableTiming)) != cudaSuccess) return e;
    if ((e = cudaEventCreateWithFlags(&l->ev_join, cudaEventDisableTiming)) != cudaSuccess) return e;
    if ((e = cudaStreamCreateWithFlags(&l->copy, cudaStreamNonBlocking)) != cudaSuccess) return e;
    if ((e = cudaEventCreateWithFlags(&l->ev_copy, cudaEventDisableTiming)) != cudaSuccess) return e;
    for (int i = 0; i < kMaxBands; ++i) {
        if ((e = cudaEventCreateWithFlags(&l->ev_band[i], cudaEventDisableTiming)) != cudaSuccess) return e;
        if ((e = cudaEventCreateWithFlags(&l->ev_out[i], cudaEventDisableTiming)) != cudaSuccess) return e;
    }
    return cudaSuccess;
}

static void lane_destroy(Lane* l) {
    if (l->stream) cudaStreamSynchronize(l->stream);
    if (l->side) {
        cudaStreamSynchronize(l->side);
        cudaStreamDestroy(l->side);
    }
    if (l->ws) cudaFree(l->ws);
    if (l->copy) {
        cudaStreamSynchronize(l->copy);
        cudaStreamDestroy(l->copy);
    }
    if (l->ev_fork) cudaEventDestroy(l->ev_fork);
    if (l->ev_join) cudaEventDestroy(l->ev_join);
    if (l->ev_copy) cudaEventDestroy(l->ev_copy);
    for (int i = 0; i < kMaxBands; ++i) {
        if (l->ev_band[i]) cudaEventDestroy(l->ev_band[i]);
        if (l->ev_out[i]) cudaEventDestroy(l->ev_out[i]);
    }
    if (l->own_stream && l->stream) cudaStreamDestroy(l->stream);
    *l = Lane();
}

// lanes 1.. are created on first use (scb_clone_batch)
static int ensure_lanes(scb_context* c, int n) {
    if (n > kMaxLanes) n = kMaxLanes;
    while (c->n_lanes < n) {
        SCB_CUDA(c, lane_create(&c->lanes[c->n_lanes], nullptr));
        c->n_lanes++;
    }
    return SCB_OK;
}

static const int kBboxInts = 8;
static int ensure_bbox_slots(scb_context* c, int slots) {
    if (slots <= c->bbox_slots) return SCB_OK;
    for (int i = 0; i < c->n_lanes; ++i) SCB_CUDA(c, cudaStreamSynchronize(c->lanes[i].stream));
    if (c->bbox_dev) cudaFree(c->bbox_dev);
    if (c->bbox_pinned) cudaFreeHost(c->bbox_pinned);
    c->bbox_dev = nullptr;
    c->bbox_pinned = nullptr;
    c->bbox_slots = 0;
    SCB_CUDA(c, cudaMalloc(&c->bbox_dev, (size_t)slots * kBboxInts * sizeof(int)));
    SCB_CUDA(c, cudaMallocHost(&c->bbox_pinned, (size_t)(slots + 1) * kBboxInts * sizeof(int)));
    for (int i = 0; i < kBboxInts; ++i) c->bbox_pinned[i] = 0;
    c->bbox_pinned[0] = INT_MAX;
    c->bbox_pinned[1] = INT_MAX;
    c->bbox_pinned[2] = -1;
    c->bbox_pinned[3] = -1;
    c->bbox_slots = slots;
    return SCB_OK;
}

// ------------------------------------------------------------------------------------------------
// kernel dispatch over the convolution length
// ------------------------------------------------------------------------------------------------
// Engine selection per convolution length:
//   group engine  (scb_kernels3.cuh): M <= 8192; CTA = 2 lines x 3 channel groups (1 group at M = 8192)
//   scalar engine (scb_kernels.cuh) : M = 16384 (one 139 KB sequence per CTA), or everything when
//                                     SCB_ENGINE=scalar is set in the environment (A/B checks)
template <int LOG2M>
struct Nch {
    static constexpr int value = (LOG2M <= 13) ? 3 : 1;  // scalar engine: 3 x 16384-point lines do not fit in 227 KB
};
template <int LOG2M>
static constexpr size_t smem_bytes() { return (size_t)Nch<LOG2M>::value * FftCfg<LOG2M>::PADDED * sizeof(float2); }

static bool use_scalar_engine(int log2m) {
    static const bool forced = [] {
        const char* e = std::getenv("SCB_ENGINE");
        return e && std::strcmp(e, "scalar") == 0;
    }();
    return forced || log2m >= 14;
}

template <class K>
static cudaError_t set_smem(K kernel, size_t bytes) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
}

template <int LOG2M>
static cudaError_t configure_one() {
    cudaError_t e;
    if ((e = set_smem(rows_fwd_kernel<LOG2M, Nch<LOG2M>::value>, smem_bytes<LOG2M>())) != cudaSuccess) return e;
    if ((e = set_smem(cols_kernel<LOG2M, Nch<LOG2M>::value>, smem_bytes<LOG2M>())) != cudaSuccess) return e;
    if ((e = set_smem(rows_inv_kernel<LOG2M, Nch<LOG2M>::value>, smem_bytes<LOG2M>())) != cudaSuccess) return e;
    if constexpr (LOG2M <= 13) {
        if ((e = set_smem(rows_fwd3_kernel<LOG2M>, GCfg<LOG2M>::SMEM)) != cudaSuccess) return e;
        if ((e = set_smem(cols3_kernel<LOG2M>, GCfg<LOG2M>::SMEM)) != cudaSuccess) return e;
        if ((e = set_smem(rows_inv3_kernel<LOG2M>, GCfg<LOG2M>::SMEM)) != cudaSuccess) return e;
        if ((e = set_smem(rows_fwd4_kernel<LOG2M>, GCfg<LOG2M>::SMEM)) != cudaSuccess) return e;
        if ((e = set_smem(cols4_kernel<LOG2M>, GCfg<LOG2M>::SMEM)) != cudaSuccess) return e;
        if ((e = set_smem(rows_inv4_kernel<LOG2M>, GCfg<LOG2M>::SMEM)) != cudaSuccess) return e;
    }
    return cudaSuccess;
}

#define SCB_FOR_LOG2M(X) X(5) X(6) X(7) X(8) X(9) X(10) X(11) X(12) X(13) X(14)

// ---- kernel switches -------------------------------------------------------------------------------------------------------------
// Each hot kernel that exists in two generations is chosen by a default below; the environment variable of the same name overrides it
// for A/B runs (tools/ab_select.py measures every switch on a B200 and requires the variant to reproduce the default's bytes).
// scb_kernel_variants() reports what is in force.
static constexpr int kRhsFoldDefault = 2;       // SCB_RHS_FOLD      1: rhs_fold_kernel<2>, 2: rhs_fold2_kernel (packed 16-bit lanes), binary masks
static constexpr int kTriSmemDefault = 1;       // SCB_TRI_SMEM      1: tri_solve_smem_kernel for whole solves whose column tile fits in shared memory, 2: tri_solve_smem2_kernel
static constexpr int kLowProjDefault = 1;       // SCB_LOWPROJ       1: tri_lowproj_kernel, 2: tri_lowproj2_kernel (no shared-memory atomics)
static int sw_env(const char* name, int dflt) {
    const char* e = std::getenv(name);
    return e ? std::atoi(e) : dflt;
}
static int sw_rhs_fold() { return sw_env("SCB_RHS_FOLD", kRhsFoldDefault); }
static int sw_tri_smem() { return sw_env("SCB_TRI_SMEM", kTriSmemDefault); }
static int sw_lowproj() { return sw_env("SCB_LOWPROJ", kLowProjDefault); }
extern "C" const char* scb_kernel_variants(void) {
    static const std::string v = "rhs_fold=" + std::to_string(sw_rhs_fold()) + " tri_smem=" + std::to_string(sw_tri_smem()) + " tri_unroll=" + std::to_string(kTriUnroll) +
                                 " lowproj=" + std::to_string(sw_lowproj()) + " " + i8_variant_string();
    return v.c_str();
}

static cudaError_t configure_all() {
    cudaError_t e = cudaSuccess;
#define X(L) if (e == cudaSuccess) e = configure_one<L>();
    SCB_FOR_LOG2M(X)
#undef X
    return e;
}

template <int LOG2M>
static dim3 group_grid(int nlines) { return dim3((nlines + 1) / 2, GCfg<LOG2M>::NG == 1 ? 3 : 1); }
template <int LOG2M>
static dim3 quad_grid(int nlines) { return dim3((nlines + 3) / 4, GCfg<LOG2M>::NG == 1 ? 3 : 1); }
// quad mode (two real lines per complex sequence) whenever the length's table carries the extended chirp spectrum;
// SCB_QUAD=0 disables it (A/B checks)
// Measured on B200 (profiles/r1_quad_sweep.txt): the row passes gain 1.0-1.7x at every length; the column pass, whose
// in-register bridge between its two convolutions becomes two extra shared-memory round trips, gains 1.3-1.5x at
// M >= 4096 but loses ~10 % at M <= 2048 -- so columns use quad mode from 4096 up.
static bool use_quad(const LenTabDev& t, bool cols = false) {
    static const bool off = [] {
        const char* e = std::getenv("SCB_QUAD");
        return e && std::strcmp(e, "0") == 0;
    }();
    return t.bhat_q != nullptr && !off && (!cols || t.log2m >= 12);
}

template <int LOG2M>
static void launch_rows_fwd_t(cudaStream_t stream, int nlines, const RowsFwdParams& p) {
    if constexpr (LOG2M <= 13) {
        if (!use_scalar_engine(LOG2M)) {
            RowsFwd3Params pp{p, p.tx.gtw, p.y0 + nlines};
            if (use_quad(p.tx) && p.rhs_in && !p.rhs_dump)
                SCB_LAUNCH(rows_fwd4_kernel<LOG2M>, quad_grid<LOG2M>(nlines), dim3(GCfg<LOG2M>::T), GCfg<LOG2M>::SMEM, stream, pp);
            else
                SCB_LAUNCH(rows_fwd3_kernel<LOG2M>, group_grid<LOG2M>(nlines), dim3(GCfg<LOG2M>::T), GCfg<LOG2M>::SMEM, stream, pp);
            return;
        }
    }
    auto k = rows_fwd_kernel<LOG2M, Nch<LOG2M>::value>;
    SCB_LAUNCH(k, dim3(nlines), dim3(FftCfg<LOG2M>::T), smem_bytes<LOG2M>(), stream, p);
}
template <int LOG2M>
static void launch_cols_t(cudaStream_t stream, int nlines, const ColsParams& p) {
    if constexpr (LOG2M <= 13) {
        if (!use_scalar_engine(LOG2M)) {
            Cols3Params pp{p, p.ty.gtw, p.x0 + nlines};
            if (use_quad(p.ty, true))
                SCB_LAUNCH(cols4_kernel<LOG2M>, quad_grid<LOG2M>(nlines), dim3(GCfg<LOG2M>::T), GCfg<LOG2M>::SMEM, stream, pp);
            else
                SCB_LAUNCH(cols3_kernel<LOG2M>, group_grid<LOG2M>(nlines), dim3(GCfg<LOG2M>::T), GCfg<LOG2M>::SMEM, stream, pp);
            return;
        }
    }
    auto k = cols_kernel<LOG2M, Nch<LOG2M>::value>;
    SCB_LAUNCH(k, dim3(nlines), dim3(FftCfg<LOG2M>::T), smem_bytes<LOG2M>(), stream, p);
}
template <int LOG2M>
static void launch_rows_inv_t(cudaStream_t stream, int nlines, const RowsInvParams& p) {
    if constexpr (LOG2M <= 13) {
        if (!use_scalar_engine(LOG2M)) {
            RowsInv3Params pp{p, p.tx.gtw, p.y0 + nlines};
            if (use_quad(p.tx))
                SCB_LAUNCH(rows_inv4_kernel<LOG2M>, quad_grid<LOG2M>(nlines), dim3(GCfg<LOG2M>::T), GCfg<LOG2M>::SMEM, stream, pp);
            else
                SCB_LAUNCH(rows_inv3_kernel<LOG2M>, group_grid<LOG2M>(nlines), dim3(GCfg<LOG2M>::T), GCfg<LOG2M>::SMEM, stream, pp);
            return;
        }
    }
    auto k = rows_inv_kernel<LOG2M, Nch<LOG2M>::value>;
    SCB_LAUNCH(k, dim3(nlines), dim3(FftCfg<LOG2M>::T), smem_bytes<LOG2M>(), stream, p);
}

static void launch_rows_fwd(scb_context* c, cudaStream_t stream, int log2m, int nlines, const RowsFwdParams& p) {
    if (nlines <= 0) return;
    switch (log2m) {
#define X(L) case L: launch_rows_fwd_t<L>(stream, nlines, p); break;
        SCB_FOR_LOG2M(X)
#undef X
    }
    c->launches++;
}
static void launch_cols(scb_context* c, cudaStream_t stream, int log2m, int nlines, const ColsParams& p) {
    if (nlines <= 0) return;
    switch (log2m) {
#define X(L) case L: launch_cols_t<L>(stream, nlines, p); break;
        SCB_FOR_LOG2M(X)
#undef X
    }
    c->launches++;
}
static void launch_rows_inv(scb_context* c, cudaStream_t stream, int log2m, int nlines, const RowsInvParams& p) {
    if (nlines <= 0) return;
    switch (log2m) {
#define X(L) case L: launch_rows_inv_t<L>(stream, nlines, p); break;
        SCB_FOR_LOG2M(X)
#undef X
    }
    c->launches++;
}

// ------------------------------------------------------------------------------------------------
// table caches
// ------------------------------------------------------------------------------------------------
// A plan takes a reference on every table it uses; callers hold c->table_mu.
static void table_ref(scb_context* c, scb_plan* p, CacheMeta* m) {
    m->last_use = ++c->use_clock;
    if (p) {  // (a table fetched without a plan -- the self-test -- is only touched, not pinned)
        m->refs++;
        p->table_refs.push_back(m);
    }
}
static void table_added(scb_context* c, CacheMeta* m, size_t bytes) {
    m->bytes = bytes;
    c->table_bytes += bytes;
}
// Evicts least-recently-used tables that no live plan references until the caches fit the budget again (cudaFree waits for
// the device, so nothing can still be reading them).  Called with c->table_mu held, after an insertion.
static void evict_tables(scb_context* c) {
    while (c->table_bytes > c->table_budget) {
        CacheMeta* best = nullptr;
        std::function<void()> drop;
        auto consider = [&](CacheMeta& m, std::function<void()> d) {
            if (m.refs == 0 && m.bytes > 0 && (!best || m.last_use < best->last_use)) {
                best = &m;
                drop = std::move(d);
            }
        };
        for (auto it = c->lentabs.begin(); it != c->lentabs.end(); ++it) consider(it->second.meta, [c, it] { cudaFree(it->second.block); c->lentabs.erase(it); });
        for (auto it = c->filters.begin(); it != c->filters.end(); ++it) consider(it->second.meta, [c, it] { cudaFree(it->second.d); c->filters.erase(it); });
        for (auto it = c->tctabs.begin(); it != c->tctabs.end(); ++it) consider(it->second.meta, [c, it] { cudaFree(it->second.block); c->tctabs.erase(it); });
        for (auto it = c->tritabs.begin(); it != c->tritabs.end(); ++it) consider(it->second.meta, [c, it] { cudaFree(it->second.block); c->tritabs.erase(it); });
        for (auto it = c->i8tabs.begin(); it != c->i8tabs.end(); ++it) consider(it->second.meta, [c, it] { cudaFree(it->second.basis); c->i8tabs.erase(it); });
        if (!best) return;  // everything left is in use
        c->table_bytes -= best->bytes;
        drop();
    }
}

static int get_lentab(scb_context* c, scb_plan* p, int n, LenTabDev* out) {
    std::lock_guard<std::mutex> lk(c->table_mu);
    auto it = c->lentabs.find(n);
    if (it != c->lentabs.end()) {
        *out = it->second.dev;
        table_ref(c, p, &it->second.meta);
        return SCB_OK;
    }
    HostLenTab h = build_len_tab(n);
    const size_t M = (size_t)1 << h.log2m;
    const size_t off_chirp = 0;
    const size_t off_bhat = align_up(off_chirp + (n + 1) * sizeof(float2), 256);
    const size_t off_bhq = align_up(off_bhat + M * sizeof(float2), 256);
    const size_t off_tw = align_up(off_bhq + h.bhat_q.size() * sizeof(float2), 256);
    const size_t off_ptw = align_up(off_tw + M * sizeof(float2), 256);
    const size_t off_sin = align_up(off_ptw + h.gtw.size() * sizeof(float), 256);
    const size_t total = align_up(off_sin + h.sinlow.size() * sizeof(double), 256);
    std::vector<char> host(total, 0);
    std::memcpy(host.data() + off_chirp, h.chirp.data(), h.chirp.size() * sizeof(HostF2));
    std::memcpy(host.data() + off_bhat, h.bhat_t.data(), h.bhat_t.size() * sizeof(HostF2));
    if (!h.bhat_q.empty()) std::memcpy(host.data() + off_bhq, h.bhat_q.data(), h.bhat_q.size() * sizeof(HostF2));
    std::memcpy(host.data() + off_tw, h.tw.data(), h.tw.size() * sizeof(HostF2));
    std::memcpy(host.data() + off_ptw, h.gtw.data(), h.gtw.size() * sizeof(float));
    std::memcpy(host.data() + off_sin, h.sinlow.data(), h.sinlow.size() * sizeof(double));
    DevLenTab d;
    SCB_CUDA(c, cudaMalloc(&d.block, total));
    {
        cudaError_t e = cudaMemcpyAsync(d.block, host.data(), total, cudaMemcpyHostToDevice, c->lanes[0].stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->lanes[0].stream);  // `host` dies at scope exit; a blocking sync also publishes the table to every lane
        if (e != cudaSuccess) {
            cudaFree(d.block);
            return fail(c, SCB_ERR_CUDA, std::string("get_lentab: ") + cudaGetErrorString(e));
        }
    }
    char* b = (char*)d.block;
    d.dev.n = n;
    d.dev.log2m = h.log2m;
    d.dev.lowk = h.lowk;
    d.dev.chirp = (const float2*)(b + off_chirp);
    d.dev.bhat_t = (const float2*)(b + off_bhat);
    d.dev.bhat_q = h.bhat_q.empty() ? nullptr : (const float2*)(b + off_bhq);
    d.dev.tw = (const float2*)(b + off_tw);
    d.dev.gtw = (const float4*)(b + off_ptw);
    d.dev.sinlow = (const double*)(b + off_sin);
    auto ins = c->lentabs.emplace(n, d);
    table_added(c, &ins.first->second.meta, total);
    table_ref(c, p, &ins.first->second.meta);
    *out = d.dev;
    evict_tables(c);
    return SCB_OK;
}

#ifndef SCB_EMU
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = [] {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) f = nullptr;
        return (EncodeTiledFn)f;
    }();
    return fn;
}
#endif

// split sine basis of one line length for the tensor-core engine, built on the device at plan time
static int get_tctab(scb_context* c, scb_plan* p, int n, const DevTcTab** out) {
    std::lock_guard<std::mutex> lk(c->table_mu);
    auto it = c->tctabs.find(n);
    if (it != c->tctabs.end()) {
        *out = &it->second;
        table_ref(c, p, &it->second.meta);
        return SCB_OK;
    }
    DevTcTab d;
    tc_geometry(n, &d.dev);
    const size_t floats = (size_t)4 * d.dev.rows * d.dev.kpad;
    SCB_CUDA(c, cudaMalloc(&d.block, floats * sizeof(float)));
    d.dev.basis = (const float*)d.block;
    {
        const long long total = 2LL * d.dev.rows * d.dev.kpad;
        long long blocks = (total + 255) / 256;
        if (blocks > (long long)c->sm_count * 16) blocks = (long long)c->sm_count * 16;
        SCB_LAUNCH(tc_basis_kernel, dim3((unsigned)blocks), dim3(256), 0, c->lanes[0].stream, d.dev, (float*)d.block);
        c->launches++;
    }
    {
        cudaError_t e = cudaStreamSynchronize(c->lanes[0].stream);  // publishes the table to every lane
        if (e == cudaSuccess) e = cudaGetLastError();
        if (e != cudaSuccess) {
            cudaFree(d.block);
            return fail(c, SCB_ERR_CUDA, std::string("get_tctab: ") + cudaGetErrorString(e));
        }
    }
#ifndef SCB_EMU
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) {
        cudaFree(d.block);
        return fail(c, SCB_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
    }
    const cuuint64_t gdim[2] = {(cuuint64_t)d.dev.kpad, (cuuint64_t)4 * d.dev.rows};
    const cuuint64_t gstride[1] = {(cuuint64_t)d.dev.kpad * sizeof(float)};
    const cuuint32_t box[2] = {(cuuint32_t)kTcKB, (cuuint32_t)d.dev.nt};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&d.map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d.block, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        cudaFree(d.block);
        return fail(c, SCB_ERR_CUDA, "cuTensorMapEncodeTiled failed with code " + std::to_string((int)r));
    }
#endif
    auto ins = c->tctabs.emplace(n, d);
    table_added(c, &ins.first->second.meta, floats * sizeof(float));
    table_ref(c, p, &ins.first->second.meta);
    *out = &ins.first->second;
    evict_tables(c);
    return SCB_OK;
}

static int wanted_engine(const scb_context* c) {
    static const int env_engine = [] {
        const char* e = std::getenv("SCB_ENGINE");
        if (!e) return (int)SCB_ENGINE_AUTO;
        if (std::strcmp(e, "tc") == 0) return (int)SCB_ENGINE_TC;
        if (std::strcmp(e, "tri") == 0) return (int)SCB_ENGINE_TRI;
        if (std::strcmp(e, "i8") == 0) return (int)SCB_ENGINE_I8;
        if (std::strcmp(e, "fft") == 0 || std::strcmp(e, "scalar") == 0) return (int)SCB_ENGINE_FFT;
        return (int)SCB_ENGINE_AUTO;
    }();
    return c->engine != SCB_ENGINE_AUTO ? c->engine : env_engine;
}

// AUTO resolves to the tridiagonal solve along y (scb_tri.cuh) with, along x, the exact INT8 tensor-core contraction
// (scb_i8.h) for line lengths in [kI8MinN, kI8MaxN] and the Bluestein FFT passes otherwise.  TRI forces the FFT passes.
static bool tri_eligible(const scb_context* c) {
    const int want = wanted_engine(c);
    return want == SCB_ENGINE_AUTO || want == SCB_ENGINE_TRI || want == SCB_ENGINE_I8;
}
static bool i8_eligible(const scb_context* c, int nx, int mode) {
    const int want = wanted_engine(c);
    if (want != SCB_ENGINE_AUTO && want != SCB_ENGINE_I8) return false;
    static const bool auto_on = [] {  // SCB_I8_AUTO=0 keeps AUTO on the FFT passes (A/B checks)
        const char* e = std::getenv("SCB_I8_AUTO");
        return !(e && std::strcmp(e, "0") == 0);
    }();
    if (want == SCB_ENGINE_AUTO && !auto_on) return false;
    (void)mode;
    const int lo = want == SCB_ENGINE_I8 ? 8 : kI8MinN;  // forced: every length the digit planes can hold (tests); AUTO: where it pays
    return nx >= lo && nx <= kI8MaxN;
}

// digit planes of the folded sine basis of one line length, built on the device at plan time, cached in the context
static int get_i8tab(scb_context* c, scb_plan* p, int n, const DevI8Tab** out) {
    std::lock_guard<std::mutex> lk(c->table_mu);
    auto it = c->i8tabs.find(n);
    if (it != c->i8tabs.end()) {
        *out = &it->second;
        table_ref(c, p, &it->second.meta);
        return SCB_OK;
    }
    DevI8Tab d;
    d.g = i8_geometry(n);
    void* b = nullptr;
    SCB_CUDA(c, cudaMalloc(&b, i8_basis_bytes(d.g)));
    d.basis = (signed char*)b;
    cudaStream_t s = c->lanes[0].stream;
    if (i8_launch_basis((void*)s, d.g, d.basis) != 0) {
        cudaFree(b);
        return fail(c, SCB_ERR_CUDA, "i8_basis_kernel launch failed");
    }
    c->launches++;
    {
        cudaError_t e = cudaStreamSynchronize(s);  // publishes the table to every lane
        if (e == cudaSuccess) e = cudaGetLastError();
        if (e != cudaSuccess) {
            cudaFree(b);
            return fail(c, SCB_ERR_CUDA, std::string("get_i8tab: ") + cudaGetErrorString(e));
        }
    }
    auto ins = c->i8tabs.emplace(n, d);
    table_added(c, &ins.first->second.meta, i8_basis_bytes(d.g));
    table_ref(c, p, &ins.first->second.meta);
    *out = &ins.first->second;
    evict_tables(c);
    return SCB_OK;
}

// LU factors m[d][k] of tridiag(-1, 4 - fx[k], -1), built on the device at plan time, cached per ROI (w, h)
static int get_tritab(scb_context* c, scb_plan* p, int w, int h, TriTabDev* out) {
    std::lock_guard<std::mutex> lk(c->table_mu);
    const auto key = std::make_pair(w, h);
    auto it = c->tritabs.find(key);
    if (it != c->tritabs.end()) {
        *out = it->second.dev;
        table_ref(c, p, &it->second.meta);
        return SCB_OK;
    }
    const int nx = w - 2, ny = h - 2;
    const std::vector<double> th = build_theta(build_filter(w));
    const int pm = (int)align_up((size_t)nx, 4);
    const int rows = tri_seg_len(ny) + 1;
    const size_t off_m32 = 0;
    const size_t off_p32 = align_up(off_m32 + (size_t)rows * pm * sizeof(float), 256);
    const size_t off_m64 = align_up(off_p32 + (size_t)rows * pm * sizeof(float), 256);
    const size_t off_p64 = align_up(off_m64 + (size_t)rows * kTriLowK * sizeof(double), 256);
    const size_t off_th = align_up(off_p64 + (size_t)rows * kTriLowK * sizeof(double), 256);
    const size_t total = align_up(off_th + (size_t)nx * sizeof(double), 256);
    DevTriTab d;
    SCB_CUDA(c, cudaMalloc(&d.block, total));
    char* b = (char*)d.block;
    cudaStream_t s = c->lanes[0].stream;
    auto drop = [&](cudaError_t e) {  // every failure path frees the block
        cudaFree(d.block);
        return fail(c, e == cudaErrorMemoryAllocation ? SCB_ERR_OUT_OF_MEMORY : SCB_ERR_CUDA, std::string("get_tritab: ") + cudaGetErrorString(e));
    };
    {
        cudaError_t e = cudaMemsetAsync(b + off_m64, 0, off_th - off_m64, s);
        if (e == cudaSuccess) e = cudaMemcpyAsync(b + off_th, th.data(), (size_t)nx * sizeof(double), cudaMemcpyHostToDevice, s);
        if (e != cudaSuccess) return drop(e);
    }
    TriTableParams tp;
    tp.theta = (const double*)(b + off_th);
    tp.nx = nx;
    tp.rows = rows;
    tp.pm = pm;
    tp.m32 = (float*)(b + off_m32);
    tp.p32 = (float*)(b + off_p32);
    tp.m64 = (double*)(b + off_m64);
    tp.p64 = (double*)(b + off_p64);
    {
        const long long total_e = (long long)rows * pm;
        long long blocks = (total_e + 255) / 256;
        if (blocks > (long long)c->sm_count * 16) blocks = (long long)c->sm_count * 16;
        SCB_LAUNCH(tri_table_kernel, dim3((unsigned)blocks), dim3(256), 0, s, tp);
        c->launches++;
    }
    {
        cudaError_t e = cudaStreamSynchronize(s);  // `th` dies at scope exit; also publishes the table to every lane
        if (e == cudaSuccess) e = cudaGetLastError();
        if (e != cudaSuccess) return drop(e);
    }
    d.dev.m32 = tp.m32;
    d.dev.p32 = tp.p32;
    d.dev.pm = pm;
    d.dev.rows = rows;
    d.dev.m64 = tp.m64;
    d.dev.p64 = tp.p64;
    d.dev.theta = tp.theta;
    auto ins = c->tritabs.emplace(key, d);
    table_added(c, &ins.first->second.meta, total);
    table_ref(c, p, &ins.first->second.meta);
    *out = d.dev;
    evict_tables(c);
    return SCB_OK;
}

static bool tc_eligible(const scb_context* c, int nx, int ny) {
    // The TF32 tensor-core engine is opt-in (scb_set_engine / SCB_ENGINE=tc): FP32 accumulation inside the tensor core truncates at
    // every MMA step, which leaves ~1e-5 relative error after ~340 steps (K ~ 900) -- inside the 1e-4 bar for the float intermediates
    // but enough to cost 0.2 % of exactly matching bytes at some shapes.  The default engine contracts in exact INT8 digit planes
    // instead (scb_i8.h).
    const int want = wanted_engine(c);
    if (want != SCB_ENGINE_TC) return false;
    return nx >= kTcMinN && ny >= kTcMinN && nx <= kTcMaxN && ny <= kTcMaxN;
}

static int get_filter(scb_context* c, scb_plan* p, int extent, const float** out) {
    std::lock_guard<std::mutex> lk(c->table_mu);
    auto it = c->filters.find(extent);
    if (it != c->filters.end()) {
        *out = it->second.d;
        table_ref(c, p, &it->second.meta);
        return SCB_OK;
    }
    std::vector<float> f = build_filter(extent);
    float* d = nullptr;
    SCB_CUDA(c, cudaMalloc(&d, f.size() * sizeof(float)));
    {
        cudaError_t e = cudaMemcpyAsync(d, f.data(), f.size() * sizeof(float), cudaMemcpyHostToDevice, c->lanes[0].stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->lanes[0].stream);
        if (e != cudaSuccess) {
            cudaFree(d);
            return fail(c, SCB_ERR_CUDA, std::string("get_filter: ") + cudaGetErrorString(e));
        }
    }
    DevFilter df;
    df.d = d;
    auto ins = c->filters.emplace(extent, df);
    table_added(c, &ins.first->second.meta, f.size() * sizeof(float));
    table_ref(c, p, &ins.first->second.meta);
    *out = d;
    evict_tables(c);
    return SCB_OK;
}

// ------------------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------------------
extern "C" int scb_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

extern "C" int scb_create(int device, void* external_stream, scb_context** out) {
    if (!out) return fail(nullptr, SCB_ERR_INVALID_ARGUMENT, "scb_create: out is null");
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0)
        return fail(nullptr, SCB_ERR_NO_DEVICE, "scb_create: no CUDA device visible (this library has no CPU fallback)");
    if (device < 0 || device >= n) return fail(nullptr, SCB_ERR_INVALID_ARGUMENT, "scb_create: device index out of range");
    scb_context* c = new scb_context();
    c->device = device;
    auto bail = [&](const char* what, cudaError_t e) {
        g_create_error = std::string(what) + ": " + cudaGetErrorString(e);
        delete c;
        return (int)SCB_ERR_CUDA;
    };
    cudaError_t e;
    if ((e = cudaSetDevice(device)) != cudaSuccess) return bail("cudaSetDevice", e);
#ifndef SCB_EMU
    {   // Plans allocate their eroded mask stream-ordered.  A PRIVATE pool: its release threshold (keep freed blocks across
        // synchronisations instead of handing them back to the OS every time) is nobody else's business, and scb_destroy
        // returns the memory.  (Setting the threshold on the device's default pool would change the host application.)
        cudaMemPoolProps props = {};
        props.allocType = cudaMemAllocationTypePinned;
        props.handleTypes = cudaMemHandleTypeNone;
        props.location.type = cudaMemLocationTypeDevice;
        props.location.id = device;
        if (cudaMemPoolCreate(&c->mem_pool, &props) == cudaSuccess && c->mem_pool) {
            unsigned long long keep = 256ull << 20;
            cudaMemPoolSetAttribute(c->mem_pool, cudaMemPoolAttrReleaseThreshold, &keep);
        } else {
            c->mem_pool = nullptr;  // fall back to the default pool, untouched
            cudaGetLastError();
        }
    }
#endif
    if ((e = lane_create(&c->lanes[0], (cudaStream_t)external_stream)) != cudaSuccess) return bail("stream/event creation", e);
    c->n_lanes = 1;
    if (const char* e = std::getenv("SCB_PLAN_CACHE")) c->plan_cache_cap = std::atoi(e) > 0 ? std::atoi(e) : 0;
    if (const char* e = std::getenv("SCB_TABLE_BUDGET_MB")) c->table_budget = (size_t)(std::atoll(e) > 0 ? std::atoll(e) : 1) << 20;
    cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device);
    cudaDeviceGetAttribute(&c->max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
    if ((e = configure_all()) != cudaSuccess) return bail("cudaFuncSetAttribute (is this an sm_100a device?)", e);
    if ((e = (cudaError_t)i8_configure()) != cudaSuccess) return bail("cudaFuncSetAttribute of the INT8 tensor-core kernels", e);
#ifndef SCB_EMU
    if ((e = cudaFuncSetAttribute(tc_pass_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTcSmemBytes)) != cudaSuccess) return bail("cudaFuncSetAttribute(tc_pass_kernel)", e);
    if ((e = cudaFuncSetAttribute(tri_solve_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTriSmemLimit)) != cudaSuccess) return bail("cudaFuncSetAttribute(tri_solve_smem_kernel)", e);
    if ((e = cudaFuncSetAttribute(tri_solve_smem2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTriSmemLimit)) != cudaSuccess) return bail("cudaFuncSetAttribute(tri_solve_smem2_kernel)", e);
#endif
    if (ensure_bbox_slots(c, 1) != SCB_OK) {
        g_create_error = c->err;
        for (int i = 0; i < c->n_lanes; ++i) lane_destroy(&c->lanes[i]);
        delete c;
        return SCB_ERR_CUDA;
    }
    *out = c;
    return SCB_OK;
}

extern "C" int scb_destroy(scb_context* c) {
    if (!c) return SCB_OK;
    cudaSetDevice(c->device);
    for (auto& cp : c->plan_cache) scb_plan_destroy(cp.plan);
    c->plan_cache.clear();
    for (int i = 0; i < c->n_lanes; ++i) lane_destroy(&c->lanes[i]);
    if (c->prep) {
        cudaStreamSynchronize(c->prep);
        cudaStreamDestroy(c->prep);
    }
    for (auto& kv : c->lentabs) cudaFree(kv.second.block);
    for (auto& kv : c->filters) cudaFree(kv.second.d);
    for (auto& kv : c->tctabs) cudaFree(kv.second.block);
    for (auto& kv : c->tritabs) cudaFree(kv.second.block);
    for (auto& kv : c->i8tabs) cudaFree(kv.second.basis);
    if (c->bbox_dev) cudaFree(c->bbox_dev);
    if (c->bbox_pinned) cudaFreeHost(c->bbox_pinned);
    host_pool_free(c);
#ifndef SCB_EMU
    if (c->mem_pool) cudaMemPoolDestroy(c->mem_pool);
#endif
    delete c;
    return SCB_OK;
}

extern "C" int scb_sync(scb_context* c) {
    if (!c) return SCB_ERR_INVALID_ARGUMENT;
    SCB_CUDA(c, cudaSetDevice(c->device));
    for (int i = 0; i < c->n_lanes; ++i) SCB_CUDA(c, cudaStreamSynchronize(c->lanes[i].stream));
    SCB_CUDA(c, cudaGetLastError());
    return SCB_OK;
}

extern "C" int scb_set_engine(scb_context* c, int engine) {
    if (!c) return SCB_ERR_INVALID_ARGUMENT;
    if (engine != SCB_ENGINE_AUTO && engine != SCB_ENGINE_FFT && engine != SCB_ENGINE_TC && engine != SCB_ENGINE_TRI && engine != SCB_ENGINE_I8)
        return fail(c, SCB_ERR_INVALID_ARGUMENT, "scb_set_engine: unknown engine");
    c->engine = engine;
    return SCB_OK;
}

extern "C" int scb_set_orientation(scb_context* c, int orientation) {
    if (!c) return SCB_ERR_INVALID_ARGUMENT;
    if (orientation < -1 || orientation > 1) return fail(c, SCB_ERR_INVALID_ARGUMENT, "scb_set_orientation: -1 (auto), 0 (FFT along x) or 1 (FFT along y)");
    c->orientation = orientation;
    return SCB_OK;
}
extern "C" void* scb_stream(scb_context* c) { return c ? (void*)c->lanes[0].stream : nullptr; }
extern "C" const char* scb_last_error(const scb_context* c) { return c ? c->err.c_str() : g_create_error.c_str(); }
extern "C" uint64_t scb_kernel_launches(const scb_context* c) { return c ? c->launches.load() : 0; }
extern "C" const char* scb_status_string(int s) {
    switch (s) {
        case SCB_OK: return "ok";
        case SCB_ERR_INVALID_ARGUMENT: return "invalid argument";
        case SCB_ERR_ROI_OUT_OF_BOUNDS: return "ROI outside dst";
        case SCB_ERR_UNSUPPORTED: return "unsupported";
        case SCB_ERR_CUDA: return "CUDA error";
        case SCB_ERR_NO_DEVICE: return "no CUDA device";
        case SCB_ERR_OUT_OF_MEMORY: return "out of device memory";
    }
    return "unknown status";
}
extern "C" int scb_host_alloc(void** out, size_t bytes) {
    if (!out) return SCB_ERR_INVALID_ARGUMENT;
    return cudaMallocHost(out, bytes ? bytes : 1) == cudaSuccess ? SCB_OK : SCB_ERR_OUT_OF_MEMORY;
}
extern "C" int scb_host_free(void* p) { return cudaFreeHost(p) == cudaSuccess ? SCB_OK : SCB_ERR_CUDA; }

// ------------------------------------------------------------------------------------------------
// plan
// ------------------------------------------------------------------------------------------------
static void plan_free_debug(scb_plan* p) {
    cudaFree(p->dbg_vx);
    cudaFree(p->dbg_vy);
    cudaFree(p->dbg_rhs);
    cudaFree(p->dbg_spec);
    cudaFree(p->dbg_u);
    p->dbg_vx = p->dbg_vy = p->dbg_rhs = p->dbg_spec = p->dbg_u = nullptr;
}

static void plan_drop_graph(scb_plan* p) {
#ifndef SCB_EMU
    if (p->graph_exec) cudaGraphExecDestroy(p->graph_exec);
    p->graph_exec = nullptr;
#endif
    p->graph_key = GraphKey();
}

extern "C" int scb_plan_destroy(scb_plan* p) {
    if (!p) return SCB_OK;
    cudaSetDevice(p->ctx->device);
    if (p->E) scbFreeAsync(p->E, p->lane->stream);
    if (p->mask_stage) scbFreeAsync(p->mask_stage, p->lane->stream);
    {
        std::lock_guard<std::mutex> lk(p->ctx->table_mu);
        for (CacheMeta* m : p->table_refs) m->refs--;
        p->table_refs.clear();
    }
    plan_drop_graph(p);
    plan_free_debug(p);
    delete p;
    return SCB_OK;
}

// Plan creation in two halves so that a batch can put every job's bounding-box reduction in flight
// before the first host sync (the reference's initMask blocks on a D2H per call: imp.cpp:1012).
//   plan_begin : validate, stage the mask (HOST), launch the bbox reduction into slot `slot`
//   plan_finish: (after a sync of the lane) geometry, erosion, tables
static bool choose_swap(const scb_plan* p);

struct PlanInput {
    MaskView mv;
    int px, py;
    int slot;
    bool staged;
    bool host_bbox;  // HOST masks: the bounding box was taken on the host while the mask uploads -- no device reduction, no round trip
    int hb[5];       // min x, min y, max x, max y, grey
};

// One pass over a HOST mask on the context's helper threads: the bounding box of the non-zero pixels inside the 1-pixel ring
// (OpenCV: copyMakeBorder(mask(1..-1), 0) + boundingRect), whether a pixel inside the ring is neither 0 nor 255, and (for the
// plan cache of scb_seamless_clone) a 64-bit hash of every byte.  8 bytes at a time; four independent multiply-xor chains keep
// the multiplier pipelined.  Replaces the reference's device reduction + blocking D2H per call (initMask, imp.cpp:927-963, 1008-1012).
struct MaskScan {
    uint64_t hash = 0;
    int minx = INT_MAX, miny = INT_MAX, maxx = -1, maxy = -1, grey = 0;
};
static MaskScan scan_rows(const scb_image* m, int y0, int y1) {
    const uint64_t K = 0x9E3779B97F4A7C15ull, M01 = 0x0101010101010101ull;
    uint64_t h0 = 0x243F6A8885A308D3ull, h1 = 0x13198A2E03707344ull, h2 = 0xA4093822299F31D0ull, h3 = 0x082EFA98EC4E6C89ull;
    MaskScan r;
    const int cols = m->cols, nw = (cols + 7) / 8, nfull = cols / 8;
    const int wl = (cols - 1) / 8;                                    // the word that holds column cols-1 (ring)
    const uint64_t last_mask = ~(0xFFull << (8 * ((cols - 1) & 7)));
    for (int y = y0; y < y1; ++y) {
        const unsigned char* row = (const unsigned char*)m->data + (size_t)y * m->stride;
        const bool inner = y > 0 && y < m->rows - 1;
        int first = -1, last_i = -1;
        uint64_t last_w = 0, grey = 0;
        auto visit = [&](int i, uint64_t w) {  // bounding box / grey flag of one word of an inner row
            if (i == 0) w &= ~0xFFull;         // column 0 belongs to the ring
            if (i == wl) w &= last_mask;       // so does column cols-1
            if (!w) return;
            if (first < 0) first = 8 * i + (__builtin_ctzll(w) >> 3);
            last_i = i;
            last_w = w;
            grey |= w ^ (((w >> 7) & M01) * 0xFFull);  // a byte is 0 or 255 iff it equals its top bit spread over the byte
        };
        int i = 0;
        for (; i + 4 <= nfull; i += 4) {  // 32 bytes per step, four independent multiply-xor chains
            uint64_t w[4];
            std::memcpy(w, row + 8 * i, 32);
            h0 = ((h0 ^ w[0]) * K) ^ (h0 >> 29);
            h1 = ((h1 ^ w[1]) * K) ^ (h1 >> 29);
            h2 = ((h2 ^ w[2]) * K) ^ (h2 >> 29);
            h3 = ((h3 ^ w[3]) * K) ^ (h3 >> 29);
            if (inner && (w[0] | w[1] | w[2] | w[3])) {
                if (i == 0 || i + 3 >= wl) {  // the blocks that touch the ring columns: word by word
                    visit(i, w[0]);
                    visit(i + 1, w[1]);
                    visit(i + 2, w[2]);
                    visit(i + 3, w[3]);
                } else {
                    if (first < 0) {
                        const int k = w[0] ? 0 : (w[1] ? 1 : (w[2] ? 2 : 3));
                        first = 8 * (i + k) + (__builtin_ctzll(w[k]) >> 3);
                    }
                    const int k = w[3] ? 3 : (w[2] ? 2 : (w[1] ? 1 : 0));
                    last_i = i + k;
                    last_w = w[k];
                    grey |= (w[0] ^ (((w[0] >> 7) & M01) * 0xFFull)) | (w[1] ^ (((w[1] >> 7) & M01) * 0xFFull)) | (w[2] ^ (((w[2] >> 7) & M01) * 0xFFull)) |
                            (w[3] ^ (((w[3] >> 7) & M01) * 0xFFull));
                }
            }
        }
        for (; i < nw; ++i) {  // the last words of the row (the very last one may be partial)
            uint64_t w = 0;
            const int left = cols - 8 * i;
            std::memcpy(&w, row + 8 * i, left >= 8 ? 8 : (size_t)left);
            h0 = ((h0 ^ w) * K) ^ (h0 >> 29);
            if (inner && w) visit(i, w);
        }
        h1 = ((h1 ^ (uint64_t)y) * K) ^ (h1 >> 31);
        if (first >= 0) {
            const int last = 8 * last_i + 7 - (__builtin_clzll(last_w) >> 3);
            if (first < r.minx) r.minx = first;
            if (last > r.maxx) r.maxx = last;
            if (y < r.miny) r.miny = y;
            if (y > r.maxy) r.maxy = y;
            if (grey) r.grey = 1;
        }
    }
    r.hash = ((h0 * K) ^ h1) * K ^ ((h2 * K) ^ h3);
    return r;
}
static int host_threads();
class HostPool;
static HostPool& pool_of(scb_context* c);
static void pool_run(scb_context* c, int n, const std::function<void(int)>& fn);  // defined next to the pool
static MaskScan scan_mask(scb_context* c, const scb_image* m, bool parallel) {
    const size_t bytes = (size_t)m->rows * m->cols;
    int T = parallel ? (int)(bytes >> 18) : 1;  // one thread per 256 KiB
    if (T > host_threads()) T = host_threads();
    if (T > 16) T = 16;
    if (T <= 1) return scan_rows(m, 0, m->rows);
    MaskScan part[16];
    const int per = (m->rows + T - 1) / T;
    pool_run(c, T, [&](int t) {
        const int a = t * per, b = (a + per < m->rows) ? a + per : m->rows;
        if (a < b) part[t] = scan_rows(m, a, b);
    });
    MaskScan r;
    uint64_t h = 0x452821E638D01377ull;
    for (int t = 0; t < T; ++t) {
        h = ((h ^ part[t].hash) * 0x9E3779B97F4A7C15ull) ^ (h >> 32);
        if (part[t].maxx < 0) continue;
        if (part[t].minx < r.minx) r.minx = part[t].minx;
        if (part[t].maxx > r.maxx) r.maxx = part[t].maxx;
        if (part[t].miny < r.miny) r.miny = part[t].miny;
        if (part[t].maxy > r.maxy) r.maxy = part[t].maxy;
        r.grey |= part[t].grey;
    }
    r.hash = h;
    return r;
}

static int plan_begin(scb_context* c, Lane* lane, cudaStream_t prep, const scb_image* mask, int mask_mem_kind, int src_rows, int src_cols,
                      int dst_rows, int dst_cols, int px, int py, int slot, scb_plan** out, PlanInput* in, int clone_flags = SCB_NORMAL_CLONE,
                      const MaskScan* scanned = nullptr, bool parallel_scan = true) {
    NvtxRange nvtx_("scb:plan_begin");
    *out = nullptr;
    const bool wide = clone_flags >= SCB_NORMAL_CLONE_WIDE;
    const int mode = wide ? clone_flags - (SCB_NORMAL_CLONE_WIDE - SCB_NORMAL_CLONE) : clone_flags;
    if (mode != SCB_NORMAL_CLONE && mode != SCB_MIXED_CLONE && mode != SCB_MONOCHROME_TRANSFER)
        return fail(c, SCB_ERR_UNSUPPORTED, "seamlessClone: flags must be NORMAL_CLONE, MIXED_CLONE, MONOCHROME_TRANSFER or their _WIDE variants");
    if (!mask || !mask->data) return fail(c, SCB_ERR_INVALID_ARGUMENT, "scb_plan_create: mask is null (pass an all-255 mask for 'no mask')");
    if (mask->channels != 1) return fail(c, SCB_ERR_INVALID_ARGUMENT, "scb_plan_create: mask must be single channel 8-bit (convert colour masks to grey first)");
    if (mask->rows != src_rows || mask->cols != src_cols) return fail(c, SCB_ERR_INVALID_ARGUMENT, "scb_plan_create: mask and src sizes differ");
    if (src_rows < 3 || src_cols < 3 || dst_rows < 3 || dst_cols < 3) return fail(c, SCB_ERR_INVALID_ARGUMENT, "scb_plan_create: images must be at least 3x3");
    if (mask->stride < mask->cols) return fail(c, SCB_ERR_INVALID_ARGUMENT, "scb_plan_create: mask stride smaller than a row");
    if (mask_mem_kind != SCB_MEM_HOST && mask_mem_kind != SCB_MEM_DEVICE) return fail(c, SCB_ERR_INVALID_ARGUMENT, "scb_plan_create: bad mem kind");
    scb_plan* p = new scb_plan();
    p->ctx = c;
    p->lane = lane;
    p->src_rows = src_rows;
    p->src_cols = src_cols;
    p->dst_rows = dst_rows;
    p->dst_cols = dst_cols;
    p->mode = mode;
    p->wide = wide;
    MaskView& mv = in->mv;
    mv.rows = mask->rows;
    mv.cols = mask->cols;
    in->px = px;
    in->py = py;
    in->slot = slot;
    in->staged = false;
    in->host_bbox = false;
    auto bad = [&](int code, const std::string& msg) {
        scb_plan_destroy(p);
        return fail(c, code, msg);
    };
    if (mask_mem_kind == SCB_MEM_HOST) {
        const size_t pitch = align_up((size_t)mask->cols, 16);
        void* m = nullptr;
        if (scbMallocPoolAsync(&m, pitch * mask->rows, c->mem_pool, prep) != cudaSuccess) return bad(SCB_ERR_OUT_OF_MEMORY, "scb_plan_create: cudaMallocAsync failed");
        p->mask_stage = (unsigned char*)m;
        cudaError_t e = cudaMemcpy2DAsync(m, pitch, mask->data, (size_t)mask->stride, (size_t)mask->cols, (size_t)mask->rows, cudaMemcpyHostToDevice, prep);
        if (e != cudaSuccess) return bad(SCB_ERR_CUDA, std::string("mask upload: ") + cudaGetErrorString(e));
        mv.data = p->mask_stage;
        mv.pitch = (long long)pitch;
        in->staged = true;
        // the bounding box on the host, while the upload is in flight: plan creation makes no device round trip for HOST masks
        const MaskScan sc = scanned ? *scanned : scan_mask(c, mask, parallel_scan);
        in->host_bbox = true;
        in->hb[0] = sc.minx;
        in->hb[1] = sc.miny;
        in->hb[2] = sc.maxx;
        in->hb[3] = sc.maxy;
        in->hb[4] = sc.grey;
        *out = p;
        return SCB_OK;
    } else {
        mv.data = (const unsigned char*)mask->data;
        mv.pitch = mask->stride;
    }
    // bbox of the ring-zeroed mask (OpenCV: copyMakeBorder + boundingRect)
    int* slot_dev = c->bbox_dev + kBboxInts * slot;
    cudaMemcpyAsync(slot_dev, c->bbox_pinned, kBboxInts * sizeof(int), cudaMemcpyHostToDevice, prep);
    {
        long long blocks = ((long long)mv.rows + 7) / 8;  // one warp per row, 8 warps per CTA
        if (blocks > (long long)c->sm_count * 8) blocks = (long long)c->sm_count * 8;
        SCB_LAUNCH(mask_bbox_kernel, dim3((unsigned)blocks), dim3(256), 0, prep, mv, slot_dev);
        c->launches++;
    }
    cudaError_t e = cudaMemcpyAsync(c->bbox_pinned + kBboxInts * (slot + 1), slot_dev, kBboxInts * sizeof(int), cudaMemcpyDeviceToHost, prep);
    if (e != cudaSuccess) return bad(SCB_ERR_CUDA, std::string("bbox download: ") + cudaGetErrorString(e));
    *out = p;
    return SCB_OK;
}

// Call after the prep stream has been synchronised past plan_begin.  On failure the plan is destroyed.
static int plan_finish(scb_plan* p, const PlanInput& in) {
    NvtxRange nvtx_("scb:plan_finish");
    scb_context* c = p->ctx;
    Lane* lane = p->lane;
    const int* r = in.host_bbox ? in.hb : c->bbox_pinned + kBboxInts * (in.slot + 1);
    const int minx = r[0], miny = r[1], maxx = r[2], maxy = r[3];
    p->grey_mask = r[4] != 0;
    auto release_stage = [&]() {
        if (p->mask_stage) scbFreeAsync(p->mask_stage, lane->stream);
        p->mask_stage = nullptr;
    };
    scb_geometry& g = p->g;
    if (maxx < 0) {  // nothing inside the ring: OpenCV returns dst unchanged
        g.empty = 1;
        release_stage();
        return SCB_OK;
    }
    g.x = minx;
    g.y = miny;
    g.w = maxx - minx + 1;
    g.h = maxy - miny + 1;
    if (p->wide) {  // *_WIDE: p is where the centre of src goes; the ROI keeps its place inside src
        g.rx = in.px - p->src_cols / 2 + g.x;
        g.ry = in.py - p->src_rows / 2 + g.y;
    } else {
        g.rx = in.px - g.w / 2;  // truncating division on the BBOX size, like cv::seamlessClone
        g.ry = in.py - g.h / 2;
    }
    g.nx = g.w - 2;
    g.ny = g.h - 2;
    auto bad = [&](int code, const char* msg) {
        scb_plan_destroy(p);
        return fail(c, code, msg);
    };
    if (g.rx < 0 || g.ry < 0 || g.rx + g.w > p->dst_cols || g.ry + g.h > p->dst_rows)
        return bad(SCB_ERR_ROI_OUT_OF_BOUNDS, "seamlessClone: ROI (mask bounding box centred at p) does not fit inside dst");
    if (g.w < 3 || g.h < 3) return bad(SCB_ERR_UNSUPPORTED, "seamlessClone: mask bounding box must be at least 3x3 (OpenCV itself crashes on such masks)");
    g.log2m_x = choose_log2m(g.nx);
    g.log2m_y = choose_log2m(g.ny);
    if (g.log2m_x > kMaxLog2M || g.log2m_y > kMaxLog2M) return bad(SCB_ERR_UNSUPPORTED, "seamlessClone: ROI side larger than 8194 pixels is not supported");

    // eroded mask, ROI sized
    p->e_pitch = (long long)align_up((size_t)g.w, 16);
    {
        void* e = nullptr;
        cudaError_t ce = scbMallocPoolAsync(&e, (size_t)p->e_pitch * g.h, c->mem_pool, lane->stream);
        if (ce != cudaSuccess) return bad(SCB_ERR_OUT_OF_MEMORY, "scb_plan_create: cudaMallocAsync failed");
        p->E = (unsigned char*)e;
    }
    SCB_LAUNCH(mask_erode_kernel, dim3((g.w + kErodeTW - 1) / kErodeTW, (g.h + kErodeTH - 1) / kErodeTH), dim3(256), 0, lane->stream, in.mv, g.x, g.y, g.w, g.h, p->E, p->e_pitch);
    c->launches++;
    release_stage();  // stream-ordered: freed after the erosion has read it
    int rc;
    if ((rc = get_lentab(c, p, g.nx, &p->tx)) || (rc = get_lentab(c, p, g.ny, &p->ty)) || (rc = get_filter(c, p, g.w, &p->fx)) || (rc = get_filter(c, p, g.h, &p->fy))) {
        scb_plan_destroy(p);
        return rc;
    }
    p->lowkx = p->tx.lowk;
    p->lowky = p->ty.lowk;
    p->use_tc = tc_eligible(c, g.nx, g.ny);
    if (p->use_tc && ((rc = get_tctab(c, p, g.nx, &p->ttx)) || (rc = get_tctab(c, p, g.ny, &p->tty)))) {
        scb_plan_destroy(p);
        return rc;
    }
    p->use_tri = !p->use_tc && tri_eligible(c);
    p->swap = p->use_tri && choose_swap(p);
    if (p->use_tri && (rc = p->swap ? get_tritab(c, p, g.h, g.w, &p->tri) : get_tritab(c, p, g.w, g.h, &p->tri))) {
        scb_plan_destroy(p);
        return rc;
    }
    p->use_i8 = p->use_tri && !p->swap && i8_eligible(c, g.nx, p->mode);
    if (p->use_i8 && (rc = get_i8tab(c, p, g.nx, &p->i8x))) {
        scb_plan_destroy(p);
        return rc;
    }
    return SCB_OK;
}

extern "C" int scb_plan_create(scb_context* c, const scb_image* mask, int mask_mem_kind, int src_rows, int src_cols,
                               int dst_rows, int dst_cols, int px, int py, scb_plan** out) {
    return scb_plan_create_ex(c, mask, mask_mem_kind, src_rows, src_cols, dst_rows, dst_cols, px, py, SCB_NORMAL_CLONE, out);
}

// sync_mask = false: the caller (scb_seamless_clone) runs a HOST execute on the same stream next, whose trailing sync also covers the
// mask upload -- a HOST plan then costs no synchronisation at all.  `scanned`: the mask scan the caller already did (hash + bbox).
static int plan_create_impl(scb_context* c, const scb_image* mask, int mask_mem_kind, int src_rows, int src_cols, int dst_rows, int dst_cols, int px, int py,
                            int clone_flags, scb_plan** out, bool sync_mask, const MaskScan* scanned) {
    if (!c) return SCB_ERR_INVALID_ARGUMENT;
    if (!out) return fail(c, SCB_ERR_INVALID_ARGUMENT, "scb_plan_create: out is null");
    *out = nullptr;
    SCB_CUDA(c, cudaSetDevice(c->device));
    Lane* lane = &c->lanes[0];
    scb_plan* p = nullptr;
    PlanInput in;
    int rc = plan_begin(c, lane, lane->stream, mask, mask_mem_kind, src_rows, src_cols, dst_rows, dst_cols, px, py, 0, &p, &in, clone_flags, scanned);
    if (rc) return rc;
    if (!in.host_bbox) {  // DEVICE masks: the bounding box comes back from the device
        cudaError_t e = cudaStreamSynchronize(lane->stream);
        if (e == cudaSuccess) e = cudaGetLastError();
        if (e != cudaSuccess) {
            scb_plan_destroy(p);
            return fail(c, SCB_ERR_CUDA, std::string("scb_plan_create: ") + cudaGetErrorString(e));
        }
    }
    if ((rc = plan_finish(p, in))) return rc;
    if (in.host_bbox && sync_mask) {  // the caller's mask buffer is free again on return (the erosion has consumed the staged copy)
        cudaError_t e = cudaStreamSynchronize(lane->stream);
        if (e == cudaSuccess) e = cudaGetLastError();
        if (e != cudaSuccess) {
            scb_plan_destroy(p);
            return fail(c, SCB_ERR_CUDA, std::string("scb_plan_create: ") + cudaGetErrorString(e));
        }
    }
    *out = p;
    return SCB_OK;
}

extern "C" int scb_plan_create_ex(scb_context* c, const scb_image* mask, int mask_mem_kind, int src_rows, int src_cols,
                                  int dst_rows, int dst_cols, int px, int py, int clone_flags, scb_plan** out) {
    return plan_create_impl(c, mask, mask_mem_kind, src_rows, src_cols, dst_rows, dst_cols, px, py, clone_flags, out, /*sync_mask=*/true, nullptr);
}

extern "C" int scb_plan_geometry(const scb_plan* p, scb_geometry* out) {
    if (!p || !out) return SCB_ERR_INVALID_ARGUMENT;
    *out = p->g;
    return SCB_OK;
}
extern "C" int scb_plan_engine(const scb_plan* p) {
    if (!p) return -1;
    return p->use_tc ? SCB_ENGINE_TC : (p->use_tri ? (p->use_i8 ? SCB_ENGINE_I8 : SCB_ENGINE_TRI) : SCB_ENGINE_FFT);
}
extern "C" int scb_plan_lowk(const scb_plan* p, int* lowkx, int* lowky) {
    if (!p) return SCB_ERR_INVALID_ARGUMENT;
    if (lowkx) *lowkx = p->lowkx;
    if (lowky) *lowky = p->lowky;
    return SCB_OK;
}

extern "C" int scb_plan_set_debug(scb_plan* p, int on) {
    if (!p) return SCB_ERR_INVALID_ARGUMENT;
    scb_context* c = p->ctx;
    SCB_CUDA(c, cudaSetDevice(c->device));
    SCB_CUDA(c, cudaStreamSynchronize(p->lane->stream));
    plan_free_debug(p);
    p->debug = false;
    if (on && !p->g.empty && p->swap) {  // the intermediates are defined in the natural orientation
        p->swap = false;
        int rc = get_tritab(c, p, p->g.w, p->g.h, &p->tri);
        if (rc) return rc;
    }
    if (on && !p->g.empty) {
        const size_t roi = (size_t)3 * p->g.w * p->g.h * sizeof(float), in = (size_t)3 * p->g.nx * p->g.ny * sizeof(float);
        SCB_CUDA(c, cudaMalloc(&p->dbg_vx, roi));
        SCB_CUDA(c, cudaMalloc(&p->dbg_vy, roi));
        SCB_CUDA(c, cudaMalloc(&p->dbg_rhs, in));
        SCB_CUDA(c, cudaMalloc(&p->dbg_spec, in));
        SCB_CUDA(c, cudaMalloc(&p->dbg_u, in));
        p->debug = true;
    }
    return SCB_OK;
}

extern "C" int scb_plan_get_intermediate(scb_plan* p, int which, float* out_host, size_t capacity, size_t* written) {
    if (!p || !out_host) return SCB_ERR_INVALID_ARGUMENT;
    scb_context* c = p->ctx;
    if (p->g.empty) return fail(c, SCB_ERR_INVALID_ARGUMENT, "scb_plan_get_intermediate: empty plan");
    SCB_CUDA(c, cudaSetDevice(c->device));
    const size_t roi = (size_t)3 * p->g.w * p->g.h, in = (size_t)3 * p->g.nx * p->g.ny;
    const float* src = nullptr;
    size_t n = 0;
    switch (which) {
        case SCB_INT_GRADIENT_X: src = p->dbg_vx; n = roi; break;
        case SCB_INT_GRADIENT_Y: src = p->dbg_vy; n = roi; break;
        case SCB_INT_RHS: src = p->dbg_rhs; n = in; break;
        case SCB_INT_SPECTRUM:
            if (p->use_tri) return fail(c, SCB_ERR_UNSUPPORTED, "scb_plan_get_intermediate: the tridiagonal engine never forms the 2-D spectrum (use SCB_ENGINE_FFT)");
            src = p->dbg_spec;
            n = in;
            break;
        case SCB_INT_SOLVED: src = p->dbg_u; n = in; break;
        case SCB_INT_ERODED_MASK: {
            n = (size_t)p->g.w * p->g.h;
            if (capacity < n) return fail(c, SCB_ERR_INVALID_ARGUMENT, "scb_plan_get_intermediate: buffer too small");
            std::vector<unsigned char> tmp((size_t)p->e_pitch * p->g.h);
            SCB_CUDA(c, cudaMemcpyAsync(tmp.data(), p->E, tmp.size(), cudaMemcpyDeviceToHost, p->lane->stream));
            SCB_CUDA(c, cudaStreamSynchronize(p->lane->stream));
            for (int y = 0; y < p->g.h; ++y)
                for (int x = 0; x < p->g.w; ++x) out_host[(size_t)y * p->g.w + x] = (float)tmp[(size_t)y * p->e_pitch + x];
            if (written) *written = n;
            return SCB_OK;
        }
        default: return fail(c, SCB_ERR_INVALID_ARGUMENT, "scb_plan_get_intermediate: unknown selector");
    }
    if (!p->debug || !src) return fail(c, SCB_ERR_INVALID_ARGUMENT, "scb_plan_get_intermediate: call scb_plan_set_debug(plan, 1) and execute first");
    if (capacity < n) return fail(c, SCB_ERR_INVALID_ARGUMENT, "scb_plan_get_intermediate: buffer too small");
    SCB_CUDA(c, cudaMemcpyAsync(out_host, src, n * sizeof(float), cudaMemcpyDeviceToHost, p->lane->stream));
    SCB_CUDA(c, cudaStreamSynchronize(p->lane->stream));
    if (written) *written = n;
    return SCB_OK;
}

// ------------------------------------------------------------------------------------------------
// execute
// ------------------------------------------------------------------------------------------------
static int check_image(scb_context* c, const scb_image* im, int rows, int cols, const char* name) {
    if (!im || !im->data) return fail(c, SCB_ERR_INVALID_ARGUMENT, std::string(name) + " is null");
    if (im->channels != 3) return fail(c, SCB_ERR_INVALID_ARGUMENT, std::string(name) + " must be 8-bit 3-channel (BGR interleaved)");
    if (im->rows != rows || im->cols != cols) return fail(c, SCB_ERR_INVALID_ARGUMENT, std::string(name) + " has a different size than the plan was made for");
    if (im->stride < (int64_t)3 * cols) return fail(c, SCB_ERR_INVALID_ARGUMENT, std::string(name) + " stride smaller than a row");
    return SCB_OK;
}

// Streaming copy for the host-side dst -> blend copy: non-temporal stores, so the destination lines are not read into the
// cache first (a plain memcpy of 11 KB image rows costs a read-for-ownership per line: 3 bytes of memory traffic per byte
// copied instead of 2 -- which is what made eight ranks sharing one host collapse, SCALE_r01).
static void stream_copy(char* d, const char* s, size_t n) {
#if defined(__SSE2__) && !defined(SCB_EMU)
    if (n < 2048) {
        std::memcpy(d, s, n);
        return;
    }
    const size_t head = (64 - ((uintptr_t)d & 63)) & 63;
    if (head) {
        std::memcpy(d, s, head);
        d += head;
        s += head;
        n -= head;
    }
    const size_t blocks = n / 64;
    for (size_t i = 0; i < blocks; ++i) {
        const __m128i a0 = _mm_loadu_si128((const __m128i*)(s + 64 * i)), a1 = _mm_loadu_si128((const __m128i*)(s + 64 * i + 16));
        const __m128i a2 = _mm_loadu_si128((const __m128i*)(s + 64 * i + 32)), a3 = _mm_loadu_si128((const __m128i*)(s + 64 * i + 48));
        _mm_stream_si128((__m128i*)(d + 64 * i), a0);
        _mm_stream_si128((__m128i*)(d + 64 * i + 16), a1);
        _mm_stream_si128((__m128i*)(d + 64 * i + 32), a2);
        _mm_stream_si128((__m128i*)(d + 64 * i + 48), a3);
    }
    if (n - 64 * blocks) std::memcpy(d + 64 * blocks, s + 64 * blocks, n - 64 * blocks);
#else
    std::memcpy(d, s, n);
#endif
}
static void stream_fence() {
#if defined(__SSE2__) && !defined(SCB_EMU)
    _mm_sfence();
#endif
}

// blend = dst everywhere EXCEPT the ROI interior (which the device result overwrites): disjoint from the
// D2H target, so it needs no ordering against the download and can run on several host threads while the
// GPU works.  (OpenCV: dst.copyTo(blend) of the whole frame, then the ROI is overwritten.)
static void host_copy_rows(const scb_image* dst, scb_image* blend, const scb_geometry& g, int y0, int y1) {
    const size_t row_bytes = (size_t)3 * dst->cols;
    const int iy0 = g.empty ? dst->rows : g.ry + 1, iy1 = g.empty ? dst->rows : g.ry + g.h - 1;
    const size_t left = (size_t)3 * (g.rx + 1), right0 = (size_t)3 * (g.rx + g.w - 1);
    const bool dense = (size_t)dst->stride == row_bytes && (size_t)blend->stride == row_bytes;
    int y = y0;
    while (y < y1) {
        char* o = (char*)blend->data + (size_t)y * blend->stride;
        const char* i = (const char*)dst->data + (size_t)y * dst->stride;
        if (y < iy0 || y >= iy1) {
            int run = 1;  // dense images: the whole run of rows above / below the ROI interior is one streaming copy
            if (dense) run = (y < iy0 ? (iy0 < y1 ? iy0 : y1) : y1) - y;
            stream_copy(o, i, row_bytes * (size_t)run);
            y += run;
        } else {
            stream_copy(o, i, left);
            stream_copy(o + right0, i + right0, row_bytes - right0);
            ++y;
        }
    }
    stream_fence();
}

// How many helper threads a context may use for host-side copies, and which cores they sit on.  One process per GPU is the
// deployment (torchrun sets LOCAL_RANK / LOCAL_WORLD_SIZE): the ranks of a node share its cores, so a rank takes
// cores / LOCAL_WORLD_SIZE of them (at most 8: the copy is memory bound long before that) and pins its helpers inside its
// own slice, instead of every rank starting 8 threads on the same cores.  SCB_HOST_THREADS overrides the count, SCB_PIN=0 the pinning.
struct HostCpus {
    std::vector<int> cpus;  // this rank's slice of the cores the process may run on
    int threads = 1;
    bool pin = false;
};
static const HostCpus& host_cpus() {
    static const HostCpus hc = [] {
        HostCpus h;
        std::vector<int> allowed;
#if defined(__linux__) && !defined(SCB_EMU)
        cpu_set_t set;
        CPU_ZERO(&set);
        if (sched_getaffinity(0, sizeof(set), &set) == 0)
            for (int i = 0; i < CPU_SETSIZE; ++i)
                if (CPU_ISSET(i, &set)) allowed.push_back(i);
#endif
        if (allowed.empty()) {
            unsigned n = std::thread::hardware_concurrency();
            for (unsigned i = 0; i < (n ? n : 1); ++i) allowed.push_back((int)i);
        }
        int world = 1, rank = 0;
        if (const char* e = std::getenv("LOCAL_WORLD_SIZE")) world = std::atoi(e) > 0 ? std::atoi(e) : 1;
        if (const char* e = std::getenv("LOCAL_RANK")) rank = std::atoi(e) >= 0 ? std::atoi(e) : 0;
        const int per = (int)allowed.size() / world > 0 ? (int)allowed.size() / world : 1;
        const int first = (rank % world) * per < (int)allowed.size() ? (rank % world) * per : 0;
        for (int i = 0; i < per && first + i < (int)allowed.size(); ++i) h.cpus.push_back(allowed[first + i]);
        h.threads = per < 8 ? per : 8;
        if (const char* e = std::getenv("SCB_HOST_THREADS")) h.threads = std::atoi(e) > 0 ? std::atoi(e) : 1;
        h.pin = world > 1;
        if (const char* e = std::getenv("SCB_PIN")) h.pin = std::atoi(e) != 0;
        return h;
    }();
    return hc;
}
static int host_threads() { return host_cpus().threads; }

// Persistent helper threads for the host-side copy (creating and joining std::threads per call costs ~50 us, a visible part
// of a 0.8 ms end-to-end clone).  One pool per CONTEXT: contexts are single-threaded per handle, so calls of
// different contexts on different threads never wait for one another.
class HostPool {
  public:
    HostPool() = default;
    ~HostPool() {
        {
            std::lock_guard<std::mutex> lk(mu_);
            stop_ = true;
        }
        cv_.notify_all();
        for (auto& t : threads_) t.join();
    }
    // runs fn(t) for t = 0 .. n-1: t = 0 on the caller, the rest on the pool; returns when all are done
    void run(int n, const std::function<void(int)>& fn) {
        if (n <= 1) {
            fn(0);
            return;
        }
        std::lock_guard<std::mutex> call(call_mu_);  // one caller at a time (scb_clone_batch plans on a second thread)
        ensure(n - 1);
        {
            std::lock_guard<std::mutex> lk(mu_);
            fn_ = &fn;
            next_ = 1;
            total_ = n;
            pending_ = n - 1;
            ++epoch_;
        }
        cv_.notify_all();
        fn(0);
        std::unique_lock<std::mutex> lk(mu_);
        done_.wait(lk, [&] { return pending_ == 0; });
        fn_ = nullptr;
    }

  private:
    void ensure(int n) {
        while ((int)threads_.size() < n) {
            threads_.emplace_back([this] { loop(); });
#if defined(__linux__) && !defined(SCB_EMU)
            const HostCpus& hc = host_cpus();
            if (hc.pin && !hc.cpus.empty()) {
                cpu_set_t set;
                CPU_ZERO(&set);
                CPU_SET(hc.cpus[threads_.size() % hc.cpus.size()], &set);
                pthread_setaffinity_np(threads_.back().native_handle(), sizeof(set), &set);
            }
#endif
        }
    }
    void loop() {
        uint64_t seen = 0;
        for (;;) {
            std::unique_lock<std::mutex> lk(mu_);
            cv_.wait(lk, [&] { return stop_ || (epoch_ != seen && next_ < total_); });
            if (stop_) return;
            while (next_ < total_) {
                const int t = next_++;
                const std::function<void(int)>* fn = fn_;
                lk.unlock();
                (*fn)(t);
                lk.lock();
                if (--pending_ == 0) done_.notify_all();
            }
            seen = epoch_;
        }
    }
    std::mutex mu_, call_mu_;
    std::condition_variable cv_, done_;
    std::vector<std::thread> threads_;
    const std::function<void(int)>* fn_ = nullptr;
    int next_ = 0, total_ = 0, pending_ = 0;
    uint64_t epoch_ = 0;
    bool stop_ = false;
};
static HostPool& pool_of(scb_context* c) {
    if (!c->pool) c->pool = new HostPool();
    return *c->pool;
}
static void pool_run(scb_context* c, int n, const std::function<void(int)>& fn) { pool_of(c).run(n, fn); }
static void host_pool_free(scb_context* c) {
    delete c->pool;
    c->pool = nullptr;
}

static void host_copy_outside(scb_context* c, const scb_image* dst, scb_image* blend, const scb_geometry& g, int max_threads) {
    NvtxRange nvtx_("scb:host_copy");
    const size_t bytes = (size_t)3 * dst->cols * dst->rows;
    int T = (int)(bytes >> 19);  // one thread per 512 KiB (the pool's threads wake in ~10 us; a 1080p frame is 6 MB)
    if (T > max_threads) T = max_threads;
    if (T <= 1) {
        host_copy_rows(dst, blend, g, 0, dst->rows);
        return;
    }
    const int per = (dst->rows + T - 1) / T;
    pool_of(c).run(T, [&](int t) {
        const int a = t * per, b = (a + per < dst->rows) ? a + per : dst->rows;
        if (a < b) host_copy_rows(dst, blend, g, a, b);
    });
}

struct Workspace {
    unsigned char *stD = nullptr, *stS = nullptr, *stO = nullptr;
    long long pD = 0, pS = 0, pO = 0;
    float *At = nullptr, *Ct = nullptr, *lowspec = nullptr, *G = nullptr;
    int gp = 0;
    int tpx = 0, tpy = 0;  // tensor-core engine: line pitches of the [..][nx] and [..][ny] orientations
    double* R = nullptr;
    double* Y64 = nullptr;  // tridiagonal engine: float64 columns k < kTriLowK
    double* W = nullptr;    // tridiagonal engine: low-frequency block coefficients [3][kTriLowL][kTriLowK]
    signed char* Adig = nullptr;  // INT8 engine: folded digit planes of the lines (forward: right-hand side, inverse: column solution)
    float* lscale = nullptr;      // INT8 engine: per-line scale of the digit planes
};

static int carve(scb_plan* p, bool host, Workspace* w) {
    scb_context* c = p->ctx;
    const scb_geometry& g = p->g;
    size_t off = 0;
    auto take = [&](size_t bytes) {
        size_t o = off;
        off = align_up(off + bytes, 256);
        return o;
    };
    const size_t in = (size_t)3 * g.nx * g.ny;
    w->pD = w->pS = (long long)align_up((size_t)3 * g.w, 128);
    w->pO = (long long)align_up((size_t)3 * g.nx, 128);
    w->gp = (int)align_up((size_t)g.nx, 4);
    size_t buf = in;
    if (p->use_tc) {  // tensor-core engine: lines padded to 16 bytes, each buffer holds either orientation
        w->tpx = (int)align_up((size_t)g.nx, 4);
        w->tpy = (int)align_up((size_t)g.ny, 4);
        const size_t a = (size_t)3 * g.nx * w->tpy, b = (size_t)3 * g.ny * w->tpx;
        buf = a > b ? a : b;
    }
    const size_t oAt = take(buf * sizeof(float)), oCt = take(buf * sizeof(float));
    const size_t gpt = align_up((size_t)g.ny, 4);
    const size_t gN = (size_t)3 * g.ny * w->gp, gT = (size_t)3 * g.nx * gpt;
    const size_t oG = take((p->swap && gT > gN ? gT : gN) * sizeof(float));
    const size_t rN = (size_t)3 * p->lowkx * g.ny, rT = (size_t)3 * p->lowky * g.nx;
    const size_t oR = take((p->swap && rT > rN ? rT : rN) * sizeof(double));
    const size_t oLow = take((size_t)3 * p->lowkx * p->lowky * sizeof(float));
    const size_t oY64 = take(p->use_tri ? (size_t)3 * (p->swap ? g.nx : g.ny) * kTriLowK * sizeof(double) : 0);
    const size_t oW = take(p->use_tri ? (size_t)3 * kTriLowL * kTriLowK * sizeof(double) : 0);
    const int i8_lines = 3 * g.ny;
    const size_t oAdig = take(p->use_i8 ? i8_adig_bytes(p->i8x->g, i8_lines, 4) : 0);
    const size_t oLs = take(p->use_i8 ? (size_t)i8_m_rows(i8_lines) * sizeof(float) : 0);
    size_t oD = 0, oS = 0, oO = 0;
    if (host) {
        oD = take((size_t)w->pD * g.h);
        oS = take((size_t)w->pS * g.h);
        oO = take((size_t)w->pO * g.ny);
    }
    int rc = ensure_ws(c, p->lane, off);
    if (rc) return rc;
    w->At = (float*)(p->lane->ws + oAt);
    w->Ct = (float*)(p->lane->ws + oCt);
    w->G = (float*)(p->lane->ws + oG);
    w->R = (double*)(p->lane->ws + oR);
    w->lowspec = (float*)(p->lane->ws + oLow);
    w->Y64 = (double*)(p->lane->ws + oY64);
    w->W = (double*)(p->lane->ws + oW);
    w->Adig = (signed char*)(p->lane->ws + oAdig);
    w->lscale = (float*)(p->lane->ws + oLs);
    if (host) {
        w->stD = (unsigned char*)(p->lane->ws + oD);
        w->stS = (unsigned char*)(p->lane->ws + oS);
        w->stO = (unsigned char*)(p->lane->ws + oO);
    }
    return SCB_OK;
}

static StencilSrc make_stencil(const scb_plan* p, const unsigned char* D, long long dp, const unsigned char* S, long long sp) {
    StencilSrc s;
    s.D = D;
    s.d_pitch = dp;
    s.S = S;
    s.s_pitch = sp;
    s.E = p->E;
    s.e_pitch = p->e_pitch;
    s.w = p->g.w;
    s.h = p->g.h;
    return s;
}

// The tridiagonal engine transforms along ONE axis and solves tridiagonal systems along the other; which is which is
// free.  A frame names the two roles: lines (FFT direction, length len, cnt of them per channel) and the cross direction.
struct Frame {
    LenTabDev tl;
    int len, cnt, log2m, lowk;
    const float *fl, *fc;  // OpenCV's filter along the lines / across them
};
static Frame frame_of(const scb_plan* p, bool swap) {
    const scb_geometry& g = p->g;
    if (swap) return Frame{p->ty, g.ny, g.nx, g.log2m_y, p->lowky, p->fy, p->fx};
    return Frame{p->tx, g.nx, g.ny, g.log2m_x, p->lowkx, p->fx, p->fy};
}
// FFT work of the two row passes: sequences x length x log2(length); quad mode packs 4 lines per sequence, else 2.
static double fft_cost(const LenTabDev& t, int lines) {
    const double seqs = 3.0 * lines / (use_quad(t) ? 4.0 : 2.0);
    return seqs * (double)(1 << t.log2m) * t.log2m;
}
// Lines along y when that is clearly cheaper (4K irregular mask: 1337-point lines run in quad mode, 1808-point ones
// do not).  The transposed stencil stores and the byte-scattered compose cost a little, hence the margin.  SCB_SWAP=0|1 forces.
static bool choose_swap(const scb_plan* p) {
    static const int forced = [] {
        const char* e = std::getenv("SCB_SWAP");
        return e ? std::atoi(e) : -1;
    }();
    if (p->mode != SCB_NORMAL_CLONE) return false;  // the mode stencils store the natural layout only
    if (p->ctx->orientation >= 0) return p->ctx->orientation != 0;
    if (forced >= 0) return forced != 0;
    // Measured on B200 (profiles/r1_s3_ab_swap.txt): at 4K the row passes drop from 221 to 186 us with the lines along y, but the
    // transposed stencil stores (+35 us), the longer tridiagonal chains (+25 us) and the byte-scattered compose cost more than
    // that, so the cost model is off until those stores go through shared-memory transposes.
    (void)fft_cost;
    return false;
}

static void run_rhs(scb_plan* p, const StencilSrc& st, float* G, int gp, int y0, int y1, bool swap = false) {
    NvtxRange nvtx_("scb:rhs");
    scb_context* c = p->ctx;
    if (y1 <= y0) return;
    RhsParams rp;
    rp.st = st;
    rp.nx = p->g.nx;
    rp.ny = p->g.ny;
    rp.g = G;
    rp.gp = gp;
    rp.y0 = y0;
    rp.y1 = y1;
    rp.transposed = swap ? 1 : 0;
    rp.gpt = (int)align_up((size_t)p->g.ny, 4);
    if (p->mode != SCB_NORMAL_CLONE) {  // MIXED_CLONE / MONOCHROME_TRANSFER: per-pixel float stencil with the mode's gradient selection
        SCB_LAUNCH(rhs_mode_kernel, dim3((gp + 31) / 32, (y1 - y0 + 7) / 8), dim3(256), 0, p->lane->stream, rp, p->mode);
    } else {
        const int chunks = (gp / 4 + kRhsThreads - 1) / kRhsThreads;
        if (swap)
            SCB_LAUNCH(rhs_kernel<true>, dim3(chunks, y1 - y0), dim3(kRhsThreads), 0, p->lane->stream, rp);
        else
            SCB_LAUNCH(rhs_kernel<false>, dim3(chunks, y1 - y0), dim3(kRhsThreads), 0, p->lane->stream, rp);
    }
    c->launches++;
    if (p->debug && !swap)  // dense [3][ny][nx] copy for scb_plan_get_intermediate (debug plans keep the natural orientation)
        for (int ch = 0; ch < 3; ++ch)
            cudaMemcpy2DAsync(p->dbg_rhs + ((size_t)ch * p->g.ny + y0) * p->g.nx, (size_t)p->g.nx * sizeof(float), G + ((size_t)ch * p->g.ny + y0) * gp,
                              (size_t)gp * sizeof(float), (size_t)p->g.nx * sizeof(float), (size_t)(y1 - y0), cudaMemcpyDeviceToDevice, p->lane->stream);
}

static void run_lowfreq_rows(scb_plan* p, const StencilSrc& st, const float* G, int gp, double* R, int y0, int y1, cudaStream_t stream, bool swap = false) {
    NvtxRange nvtx_("scb:lowfreq_rows");
    scb_context* c = p->ctx;
    if (y1 <= y0) return;
    const Frame f = frame_of(p, swap);
    LowRowsParams lp;
    lp.st = st;
    lp.sinx = f.tl.sinlow;
    lp.nx = f.len;
    lp.ny = f.cnt;
    lp.lowkx = f.lowk;
    lp.R = R;
    lp.rhs_in = G;
    lp.rhs_pitch = gp;
    lp.y0 = y0;
    SCB_LAUNCH(lowfreq_rows_kernel, dim3(y1 - y0), dim3(kLowThreads), 0, stream, lp);
    c->launches++;
}
static void run_lowfreq_cols(scb_plan* p, const double* R, float* lowspec, cudaStream_t stream) {
    scb_context* c = p->ctx;
    LowColsParams lc;
    lc.R = R;
    lc.siny = p->ty.sinlow;
    lc.ny = p->g.ny;
    lc.lowkx = p->lowkx;
    lc.lowky = p->lowky;
    lc.lowspec = lowspec;
    SCB_LAUNCH(lowfreq_cols_kernel, dim3(3 * p->lowkx), dim3(kLowThreads), 0, stream, lc);
    c->launches++;
}
static void run_rows_fwd(scb_plan* p, const StencilSrc& st, const float* G, int gp, float* At, int y0, int y1, bool natural = false, bool swap = false) {
    NvtxRange nvtx_("scb:rows_fwd");
    const Frame f = frame_of(p, swap);
    RowsFwdParams a;
    a.st = st;
    a.tx = f.tl;
    a.nx = f.len;
    a.ny = f.cnt;
    a.At = At;
    a.rhs_dump = nullptr;
    a.rhs_in = G;
    a.rhs_pitch = gp;
    a.y0 = y0;
    a.natural = natural ? 1 : 0;
    launch_rows_fwd(p->ctx, p->lane->stream, f.log2m, y1 - y0, a);
}
// SCB_REFINE=0: the FFT engine divides its own float32 spectrum everywhere (no exact low-frequency corner) -- the "plain
// float32 FFT" variant of SURVEY.md Appendix A, kept as a measurement switch (tools/parity_report.py).
static bool refine_off() {
    static const bool off = [] {
        const char* e = std::getenv("SCB_REFINE");
        return e && std::strcmp(e, "0") == 0;
    }();
    return off;
}

static void run_cols(scb_plan* p, const float* At, float* Ct, const float* lowspec, int x0, int x1) {
    NvtxRange nvtx_("scb:cols");
    if (refine_off()) lowspec = nullptr;
    ColsParams b;
    b.ty = p->ty;
    b.nx = p->g.nx;
    b.ny = p->g.ny;
    b.At = At;
    b.Ct = Ct;
    b.fx = p->fx;
    b.fy = p->fy;
    b.lowspec = lowspec;
    b.lowkx = p->lowkx;
    b.lowky = p->lowky;
    b.spec_dump = p->debug ? p->dbg_spec : nullptr;
    b.inv_scale = (float)(1.0 / (double)(p->g.ny + 1));
    b.x0 = x0;
    launch_cols(p->ctx, p->lane->stream, p->g.log2m_y, x1 - x0, b);
}
// Tridiagonal engine, pass B (scb_tri.cuh): partitioned Thomas solve of every spectral column (A [3][ny][nx] -> Ct [3][ny][nx]);
// the projections of the low-frequency block need only pass A and run on `proj_stream` beside the solve.
static TriLowParams tri_low_params(scb_plan* p, const Frame& f, const float* A, float* Ct, const double* R, double* Y64, double* W, int y0, int y1) {
    TriLowParams l;
    l.nx = f.len;
    l.ny = f.cnt;
    l.A = A;
    l.R = R;
    l.lowkx = f.lowk;
    l.Y64 = Y64;
    l.fx = f.fl;
    l.fy = f.fc;
    l.W = W;
    l.w_slots = 1;
    l.Ct = Ct;
    l.y0 = y0;
    l.y1 = y1;
    return l;
}
static TriSolveParams tri_solve_params(scb_plan* p, const Frame& f, const float* A, float* Ct, double* Y64) {
    TriSolveParams t;
    t.tab = p->tri;
    t.nx = f.len;
    t.ny = f.cnt;
    t.A = A;
    t.Ct = Ct;
    t.Y64 = Y64;
    t.x0 = 0;
    t.x1 = f.len;
    t.seg_len = tri_seg_len(f.cnt);
    t.phase = 0;
    t.seg0 = 0;
    t.seg1 = kTriSegs;
    t.ends32 = nullptr;
    t.ends64 = nullptr;
    return t;
}
// The whole solve in one launch (phase 0) runs out of shared memory when the column tile of a CTA fits (tri_solve_smem_kernel:
// ny <= ~3000 with 16 columns); longer columns and the row-sharded phases walk global memory (tri_solve_kernel).
// SCB_TRI_SMEM=0/1 overrides the default (A/B checks).
static void launch_tri_solve(scb_plan* p, const TriSolveParams& t) {
    static const int smem_variant = sw_tri_smem();
    static const bool smem_on = smem_variant != 0;
    if (smem_on && t.phase == 0 && t.x0 == 0 && t.x1 == t.nx) {
        int tables = 1;
        size_t bytes = tri_smem_bytes(t.ny, true, kTriSmemLimit);
        if (!bytes) {
            tables = 0;
            bytes = tri_smem_bytes(t.ny, false, kTriSmemLimit);
        }
        if (bytes) {
            const int nfloat = t.nx > kTriLowK ? (t.nx - kTriLowK + kTriCols - 1) / kTriCols : 0;
            if (smem_variant == 2)
                SCB_LAUNCH(tri_solve_smem2_kernel, dim3(kTriLowK / kTriCols64 + nfloat, 3), dim3(kTriCols * kTriSegs), bytes, p->lane->stream, t, tables);
            else
                SCB_LAUNCH(tri_solve_smem_kernel, dim3(kTriLowK / kTriCols64 + nfloat, 3), dim3(kTriCols * kTriSegs), bytes, p->lane->stream, t, tables);
            p->ctx->launches++;
            return;
        }
    }
    SCB_LAUNCH(tri_solve_kernel, dim3((t.nx + kTriCols - 1) / kTriCols, 3), dim3(kTriCols * kTriSegs), 0, p->lane->stream, t);
    p->ctx->launches++;
}
static void launch_tri_low(scb_plan* p, bool apply, const TriLowParams& l, cudaStream_t stream) {
    const dim3 grid((l.y1 - l.y0 + kTriLowRows - 1) / kTriLowRows, 3), block(32 * kTriLowWarps);
    if (l.y1 <= l.y0) return;
    static const int proj_variant = sw_lowproj();
    if (apply)
        SCB_LAUNCH(tri_lowapply_kernel, grid, block, 0, stream, l);
    else if (proj_variant == 2)
        SCB_LAUNCH(tri_lowproj2_kernel, grid, block, 0, stream, l);
    else
        SCB_LAUNCH(tri_lowproj_kernel, grid, block, 0, stream, l);
    p->ctx->launches++;
}

// Tridiagonal engine, pass B (scb_tri.cuh): partitioned Thomas solve of every spectral column (A [3][cnt][len] -> Ct [3][cnt][len]);
// the projections of the low-frequency block need only pass A and run on `proj_stream` beside the solve.
static int run_tri(scb_plan* p, const float* A, float* Ct, const double* R, double* Y64, double* W, cudaStream_t proj_stream, bool swap) {
    NvtxRange nvtx_("scb:tri_solve");
    scb_context* c = p->ctx;
    const Frame f = frame_of(p, swap);  // "columns" of the solve = the f.len frequencies of a line, solved across the f.cnt lines
    Lane* L = p->lane;
    cudaStream_t ms = L->stream;
    const TriLowParams l = tri_low_params(p, f, A, Ct, R, Y64, W, 0, f.cnt);
    if (proj_stream != ms) {
        SCB_CUDA(c, cudaEventRecord(L->ev_fork, ms));
        SCB_CUDA(c, cudaStreamWaitEvent(proj_stream, L->ev_fork, 0));
    }
    SCB_CUDA(c, cudaMemsetAsync(W, 0, (size_t)3 * kTriLowL * kTriLowK * sizeof(double), proj_stream));
    launch_tri_low(p, false, l, proj_stream);
    if (proj_stream != ms) SCB_CUDA(c, cudaEventRecord(L->ev_join, proj_stream));
    launch_tri_solve(p, tri_solve_params(p, f, A, Ct, Y64));
    if (proj_stream != ms) SCB_CUDA(c, cudaStreamWaitEvent(ms, L->ev_join, 0));
    launch_tri_low(p, true, l, ms);
    return SCB_OK;
}

static void run_rows_inv(scb_plan* p, const float* Ct, unsigned char* out, long long out_pitch, int y0, int y1, bool swap = false) {
    NvtxRange nvtx_("scb:rows_inv");
    const Frame f = frame_of(p, swap);
    RowsInvParams r;
    r.tx = f.tl;
    r.nx = f.len;
    r.ny = f.cnt;
    r.Ct = Ct;
    r.out = out;
    r.out_pitch = out_pitch;
    r.u_dump = p->debug ? p->dbg_u : nullptr;
    r.inv_scale = (float)(1.0 / (double)(f.len + 1));
    r.y0 = y0;
    r.transposed = swap ? 1 : 0;
    launch_rows_inv(p->ctx, p->lane->stream, f.log2m, y1 - y0, r);
}

// stage boundaries recorded by scb_plan_execute_timed
enum { ST_BEGIN = 0, ST_IN, ST_RHS, ST_LOW, ST_ROWS_FWD, ST_COLS, ST_ROWS_INV, ST_OUT, ST_X_DIGF, ST_X_DIGI, ST_X_GEMMI, ST_COUNT };
static const int kStages = ST_OUT;  // stage_ms has ST_OUT entries; the ST_X_* events split the INT8 passes into their kernels

struct StageTimer {
    cudaEvent_t ev[ST_COUNT];
    bool on = false;
    cudaStream_t stream = nullptr;
    void mark(int i) {
        if (on) cudaEventRecord(ev[i], stream);
    }
};

// ---- exact INT8 tensor-core passes along x (scb_i8.h): digit planes -> tcgen05.mma.kind::i8 -> class sums -> float ----
// forward: G [3][ny][gp] -> A [3][ny][nx] (= -2 sum g sin, what rows_fwd produces) and the exact float64 row sums R [3][lowkx][ny]
// NORMAL_CLONE on the INT8 engine: the stencil writes the folded digit planes itself (rhs_fold_kernel) -- no float right-hand side,
// no digitise launch.  SCB_I8_FUSE=0 keeps the two-kernel path (A/B checks); debug plans need G and keep it too.
static bool i8_fused_rhs(const scb_plan* p) {
    static const bool off = [] {
        const char* e = std::getenv("SCB_I8_FUSE");
        return e && std::strcmp(e, "0") == 0;
    }();
    return p->use_i8 && p->mode == SCB_NORMAL_CLONE && !p->debug && !off;
}
static void run_rhs_fold(scb_plan* p, const StencilSrc& st, const Workspace& w, int y0, int y1) {
    NvtxRange nvtx_("scb:rhs_fold");
    const scb_geometry& g = p->g;
    RhsFoldParams f;
    f.st = st;
    f.nx = g.nx;
    f.ny = g.ny;
    f.kpar0 = p->i8x->g.kpar[0];
    f.kpad = p->i8x->g.kpad;
    f.lines = 3 * g.ny;
    f.m_rows = i8_m_rows(f.lines);
    f.planes = w.Adig;
    f.lscale = w.lscale;
    f.scale = p->grey_mask ? 65536.0f : 1.0f;
    f.y0 = y0;
    const int rows = (y1 >= g.ny ? (f.m_rows + 2) / 3 : y1) - y0;  // the last range also writes the zero pad lines up to the last whole tile
    if (rows <= 0) return;
    const dim3 grid((f.kpad / 4 + kRhsThreads - 1) / kRhsThreads, rows);
    static const int variant = sw_rhs_fold();
    if (p->grey_mask)
        SCB_LAUNCH(rhs_fold_kernel<4>, grid, dim3(kRhsThreads), 0, p->lane->stream, f);
    else if (variant == 2)
        SCB_LAUNCH(rhs_fold2_kernel, grid, dim3(kRhsThreads), 0, p->lane->stream, f);
    else
        SCB_LAUNCH(rhs_fold_kernel<2>, grid, dim3(kRhsThreads), 0, p->lane->stream, f);
    p->ctx->launches++;
}

// Rows [y0, y1) of the ROI interior = lines [3 y0, 3 y1) (channel-interleaved).  The tiles that straddle the ends of the range are
// computed whole (their foreign lines hold whatever the digit planes hold) but only the lines of the range are stored.
static int run_i8_forward(scb_plan* p, const Workspace& w, const float* G, int gp, float* A, double* R, int y0, int y1, StageTimer* tm = nullptr) {
    NvtxRange nvtx_("scb:rows_fwd_i8");
    scb_context* c = p->ctx;
    const scb_geometry& g = p->g;
    const int lines = 3 * g.ny, da = p->grey_mask ? 4 : 2;
    I8DigitizeParams d{};
    d.g = p->i8x->g;
    d.in = G;
    d.in_plane = (long long)g.ny * gp;
    d.in_pitch = gp;
    d.lpc = g.ny;
    d.lines = lines;
    d.m_rows = i8_m_rows(lines);
    d.a = w.Adig;
    d.lscale = w.lscale;
    d.fixed_scale = p->grey_mask ? 65536.0f : 1.0f;  // binary mask: the right-hand side is integer valued (|g| <= 1530), two digits hold its fold exactly
    d.per_line = 0;
    d.line0 = 3 * y0;
    d.line1 = y1 >= g.ny ? d.m_rows : 3 * y1;  // the pad lines up to the last whole tile are written as zeros
    if (!i8_fused_rhs(p)) {  // (fused: rhs_fold_kernel has written the digit planes already)
        if (i8_launch_digitize((void*)p->lane->stream, d, da) != 0) return fail(c, SCB_ERR_CUDA, "i8_digitize_kernel launch failed");
        c->launches++;
    }
    if (tm) tm->mark(ST_X_DIGF);
    I8GemmParams m{};
    m.g = p->i8x->g;
    m.lines = lines;
    m.lpc = g.ny;
    m.m_rows = d.m_rows;
    m.a = w.Adig;
    m.basis = p->i8x->basis;
    m.lscale = w.lscale;
    m.mt0 = 3 * y0 / kI8M;
    m.mt1 = (3 * y1 + kI8M - 1) / kI8M;
    m.line0 = 3 * y0;
    m.line1 = 3 * y1 < lines ? 3 * y1 : lines;
    const double unit = std::ldexp(1.0, 8 * (da - 1) - kI8BasisBits);
    m.scale = (float)(-2.0 * unit);  // OpenCV: Im of the odd-extension FFT = -2 sum x sin
    m.out = A;
    m.out_plane = (long long)g.ny * g.nx;
    m.out_pitch = g.nx;
    m.R = R;
    m.lowk = p->lowkx;
    m.rscale = unit;
    if (i8_launch_gemm((void*)p->lane->stream, m, da, 4) != 0) return fail(c, SCB_ERR_CUDA, "i8_gemm_kernel (forward) launch failed");
    c->launches++;
    return SCB_OK;
}
// inverse: Ct [3][ny][nx] -> U [3][ny][nx] (= sum Ct sin / (nx+1))
// With `out8` the pass is also the compose: its epilogue clamps, truncates and stores the interleaved bytes (no float field, no compose
// launch).  Debug plans (which dump the float field) and SCB_I8_FUSE=0 keep the separate compose kernel.
static bool i8_fused_compose(const scb_plan* p) {
    static const bool off = [] {
        const char* e = std::getenv("SCB_I8_FUSE");
        return e && std::strcmp(e, "0") == 0;
    }();
    return p->use_i8 && !p->debug && !off;
}
static int run_i8_inverse(scb_plan* p, const Workspace& w, const float* Ct, float* U, int y0, int y1, StageTimer* tm = nullptr, unsigned char* out8 = nullptr,
                          long long out8_pitch = 0) {
    NvtxRange nvtx_("scb:rows_inv_i8");
    scb_context* c = p->ctx;
    const scb_geometry& g = p->g;
    const int lines = 3 * g.ny;
    I8DigitizeParams d{};
    d.g = p->i8x->g;
    d.in = Ct;
    d.in_plane = (long long)g.ny * g.nx;
    d.in_pitch = g.nx;
    d.lpc = g.ny;
    d.lines = lines;
    d.m_rows = i8_m_rows(lines);
    d.a = w.Adig;
    d.lscale = w.lscale;
    d.fixed_scale = 1.0f;
    d.per_line = 1;  // 30-bit fixed point relative to the line's largest magnitude
    d.line0 = 3 * y0;
    d.line1 = y1 >= g.ny ? d.m_rows : 3 * y1;
    if (i8_launch_digitize((void*)p->lane->stream, d, 4) != 0) return fail(c, SCB_ERR_CUDA, "i8_digitize_kernel launch failed");
    c->launches++;
    if (tm) tm->mark(ST_X_DIGI);
    I8GemmParams m{};
    m.g = p->i8x->g;
    m.lines = lines;
    m.lpc = g.ny;
    m.m_rows = d.m_rows;
    m.a = w.Adig;
    m.basis = p->i8x->basis;
    m.lscale = w.lscale;
    m.mt0 = 3 * y0 / kI8M;
    m.mt1 = (3 * y1 + kI8M - 1) / kI8M;
    m.line0 = 3 * y0;
    m.line1 = 3 * y1 < lines ? 3 * y1 : lines;
    m.scale = (float)(std::ldexp(1.0, 8 * 3 - kI8BasisBits) / (double)(g.nx + 1));
    m.out = U;
    m.out_plane = (long long)g.ny * g.nx;
    m.out_pitch = g.nx;
    m.out_u8 = out8;
    m.out_u8_pitch = out8_pitch;
    m.R = nullptr;
    if (i8_launch_gemm((void*)p->lane->stream, m, 4, 3) != 0) return fail(c, SCB_ERR_CUDA, "i8_gemm_kernel (inverse) launch failed");
    c->launches++;
    if (tm) tm->mark(ST_X_GEMMI);
    if (p->debug && y1 >= g.ny) cudaMemcpyAsync(p->dbg_u, U, (size_t)3 * g.nx * g.ny * sizeof(float), cudaMemcpyDeviceToDevice, p->lane->stream);
    return SCB_OK;
}
// compose rows [y0, y1): planar float solved field -> interleaved u8 (clamp, truncate)
static void run_compose(scb_plan* p, const float* U, unsigned char* out, long long out_pitch, int y0, int y1) {
    NvtxRange nvtx_("scb:compose");
    if (y1 <= y0) return;
    const scb_geometry& g = p->g;
    I8ComposeParams cp;
    cp.u = U;
    cp.plane = (long long)g.ny * g.nx;
    cp.pitch = g.nx;
    cp.nx = g.nx;
    cp.out = out;
    cp.out_pitch = out_pitch;
    cp.y0 = y0;
    i8_launch_compose((void*)p->lane->stream, cp, y1 - y0);
    p->ctx->launches++;
}

// ---- tensor-core engine: the four passes + compose (scb_tc.cuh) ----
// (the dynamic shared memory attribute of tc_pass_kernel is per DEVICE: it is set in scb_create, after cudaSetDevice)
static int tc_configure_once(scb_context* c) {
    (void)c;
    return SCB_OK;
}

static void launch_tc_pass(scb_plan* p, const DevTcTab* tab, const TcPassParams& pp) {
    const int lines = pp.line_end - pp.line0;
    if (lines <= 0) return;
    const dim3 grid((lines + kTcM - 1) / kTcM, 2 * tab->dev.ntiles);
#ifdef SCB_EMU
    SCB_LAUNCH(tc_pass_kernel, grid, dim3(kTcM), 0, p->lane->stream, pp);
#else
    tc_pass_kernel<<<grid, kTcThreads, kTcSmemBytes, p->lane->stream>>>(tab->map, pp);
#endif
    p->ctx->launches++;
}

static int tc_solve(scb_plan* p, const Workspace& w, unsigned char* out, long long out_pitch, StageTimer& tm);

// defer_host: batch mode -- leave the trailing stream sync AND the host-side dst->blend copy to the caller
static int tc_solve(scb_plan* p, const Workspace& w, unsigned char* out, long long out_pitch, StageTimer& tm) {
    scb_context* c = p->ctx;
    const scb_geometry& g = p->g;
    int rc = tc_configure_once(c);
    if (rc) return rc;
    float* bufA = w.At;
    float* bufB = w.Ct;
    TcPassParams a{};
    // rows forward: G [c][y][x] -> At [c][kx][y]
    a.tab = p->ttx->dev;
    a.in = w.G;
    a.in_plane = (long long)g.ny * w.gp;
    a.in_pitch = w.gp;
    a.lpc = g.ny;
    a.line0 = 0;
    a.line_end = 3 * g.ny;
    a.out = bufA;
    a.out_plane = (long long)g.nx * w.tpy;
    a.out_pitch = w.tpy;
    a.transposed = 1;
    a.scale = -2.0f;  // OpenCV: Im of the odd-extension FFT = -2 sum x sin
    launch_tc_pass(p, p->ttx, a);
    if (!tm.on) SCB_CUDA(c, cudaStreamWaitEvent(p->lane->stream, p->lane->ev_join, 0));  // the refinement corner is needed from here on
    tm.mark(ST_ROWS_FWD);
    // columns forward + refinement + eigenvalue division: At [c][kx][y] -> Q [c][kx][ky]
    TcPassParams b{};
    b.tab = p->tty->dev;
    b.in = bufA;
    b.in_plane = (long long)g.nx * w.tpy;
    b.in_pitch = w.tpy;
    b.lpc = g.nx;
    b.line0 = 0;
    b.line_end = 3 * g.nx;
    b.out = bufB;
    b.out_plane = (long long)g.nx * w.tpy;
    b.out_pitch = w.tpy;
    b.transposed = 0;
    b.scale = -2.0f;
    b.f_line = p->fx;
    b.f_out = p->fy;
    b.lowspec = w.lowspec;
    b.lowk_line = p->lowkx;
    b.lowk_out = p->lowky;
    b.spec_dump = p->debug ? p->dbg_spec : nullptr;
    launch_tc_pass(p, p->tty, b);
    // columns inverse: Q [c][kx][ky] -> Ct [c][y][kx]
    TcPassParams d{};
    d.tab = p->tty->dev;
    d.in = bufB;
    d.in_plane = (long long)g.nx * w.tpy;
    d.in_pitch = w.tpy;
    d.lpc = g.nx;
    d.line0 = 0;
    d.line_end = 3 * g.nx;
    d.out = bufA;
    d.out_plane = (long long)g.ny * w.tpx;
    d.out_pitch = w.tpx;
    d.transposed = 1;
    d.scale = (float)(1.0 / (double)(g.ny + 1));
    launch_tc_pass(p, p->tty, d);
    tm.mark(ST_COLS);
    // rows inverse: Ct [c][y][kx] -> U [c][y][x]
    TcPassParams e{};
    e.tab = p->ttx->dev;
    e.in = bufA;
    e.in_plane = (long long)g.ny * w.tpx;
    e.in_pitch = w.tpx;
    e.lpc = g.ny;
    e.line0 = 0;
    e.line_end = 3 * g.ny;
    e.out = bufB;
    e.out_plane = (long long)g.ny * w.tpx;
    e.out_pitch = w.tpx;
    e.transposed = 0;
    e.scale = (float)(1.0 / (double)(g.nx + 1));
    launch_tc_pass(p, p->ttx, e);
    // compose: U planar float -> interleaved u8
    TcComposeParams cp;
    cp.u = bufB;
    cp.plane = (long long)g.ny * w.tpx;
    cp.pitch = w.tpx;
    cp.nx = g.nx;
    cp.ny = g.ny;
    cp.out = out;
    cp.out_pitch = out_pitch;
    cp.u_dump = p->debug ? p->dbg_u : nullptr;
    SCB_LAUNCH(tc_compose_kernel, dim3((g.nx + 511) / 512, g.ny), dim3(128), 0, p->lane->stream, cp);
    c->launches++;
    tm.mark(ST_ROWS_INV);
    return SCB_OK;
}

static int execute_impl(scb_plan* p, const scb_image* src, const scb_image* dst, scb_image* blend, int mem_kind, int exec_flags, StageTimer& tm, bool defer_host = false) {
    NvtxRange nvtx_("scb:execute");
    if (!p) return SCB_ERR_INVALID_ARGUMENT;
    scb_context* c = p->ctx;
    int rc;
    if ((rc = check_image(c, src, p->src_rows, p->src_cols, "src"))) return rc;
    if ((rc = check_image(c, dst, p->dst_rows, p->dst_cols, "dst"))) return rc;
    if ((rc = check_image(c, blend, p->dst_rows, p->dst_cols, "blend"))) return rc;
    if (mem_kind != SCB_MEM_HOST && mem_kind != SCB_MEM_DEVICE) return fail(c, SCB_ERR_INVALID_ARGUMENT, "scb_plan_execute: bad mem kind");
    SCB_CUDA(c, cudaSetDevice(c->device));
    const bool host = (mem_kind == SCB_MEM_HOST);
    const bool copy_dst = (blend->data != dst->data) && !(exec_flags & SCB_EXEC_BLEND_PREFILLED);
    const scb_geometry& g = p->g;
    const size_t row_bytes = (size_t)3 * p->dst_cols;

    if (g.empty) {  // OpenCV: blend = dst
        if (copy_dst) {
            if (host) {
                if (!defer_host) host_copy_outside(c, dst, blend, g, host_threads());
            } else {
                SCB_CUDA(c, cudaMemcpy2DAsync(blend->data, (size_t)blend->stride, dst->data, (size_t)dst->stride, row_bytes, (size_t)p->dst_rows, cudaMemcpyDeviceToDevice, p->lane->stream));
            }
        }
        return SCB_OK;
    }

    Workspace w;
    if ((rc = carve(p, host, &w))) return rc;
    const unsigned char* dROI = (const unsigned char*)dst->data + (size_t)g.ry * dst->stride + (size_t)3 * g.rx;
    const unsigned char* sROI = (const unsigned char*)src->data + (size_t)g.y * src->stride + (size_t)3 * g.x;
    unsigned char* bInt = (unsigned char*)blend->data + (size_t)(g.ry + 1) * blend->stride + (size_t)3 * (g.rx + 1);
    StencilSrc st;
    unsigned char* out;
    long long out_pitch;
    tm.mark(ST_BEGIN);
    Lane* L = p->lane;
    cudaStream_t ms = L->stream;
    // HOST calls on the FFT engine move the ROI in row bands on a copy stream so that the upload of band b+1 runs
    // under the stencil + row transform of band b, and the download of band b under the inverse rows of band b+1.
    int nb = 1;
    if (host && !tm.on && !p->use_tc && !p->debug) {
        const size_t roi_bytes = (size_t)3 * g.w * g.h;
        nb = (int)(roi_bytes >> 21);  // ~2 MiB of ROI per band at least
        if (const char* e = std::getenv("SCB_BANDS")) nb = std::atoi(e);  // tests force banding on small ROIs
        if (nb > kMaxBands) nb = kMaxBands;
        if (nb > g.ny / 4) nb = g.ny / 4;
        if (nb < 1) nb = 1;
    }
    const bool swap = p->use_tri && p->swap;  // lines along y: the stencil still runs band by band, the passes need the whole ROI
    const Frame fr = frame_of(p, swap);
    const int nb_out = swap ? 1 : nb;
    int yb[kMaxBands + 1];
    if (p->use_i8 && nb > g.ny / 128) nb = g.ny / 128 > 0 ? g.ny / 128 : 1;  // INT8 passes: bands of whole 128-row blocks (= 3 x 128-line tiles)
    const long long bq = p->use_i8 ? 127LL : 3LL;
    for (int b = 0; b <= nb; ++b) yb[b] = (b == nb) ? g.ny : (int)(((long long)g.ny * b / nb) & ~bq);  // multiples of 4: whole quads
    // Small jobs of a batch keep everything on the lane's own stream: the other lanes already fill the GPU, and the event
    // forks / joins of the side stream cost more host time than such a job's kernels take (the batch path is host-bound).
    const bool serial = tm.on || (defer_host && !p->use_tc && (size_t)g.nx * g.ny < ((size_t)1 << 18));
    const bool side_copy = !host && copy_dst && !serial;
    if (host) {
        st = make_stencil(p, w.stD, w.pD, w.stS, w.pS);
        out = w.stO;
        out_pitch = w.pO;
        if (nb == 1) {
            // ROI-only transfers (the reference uploads the whole dst every call: seamlessClone_imp.cpp:419-421)
            SCB_CUDA(c, cudaMemcpy2DAsync(w.stD, (size_t)w.pD, dROI, (size_t)dst->stride, (size_t)3 * g.w, (size_t)g.h, cudaMemcpyHostToDevice, ms));
            SCB_CUDA(c, cudaMemcpy2DAsync(w.stS, (size_t)w.pS, sROI, (size_t)src->stride, (size_t)3 * g.w, (size_t)g.h, cudaMemcpyHostToDevice, ms));
        } else {
            SCB_CUDA(c, cudaEventRecord(L->ev_fork, ms));  // the staging buffers are free once earlier work on this lane is done
            SCB_CUDA(c, cudaStreamWaitEvent(L->copy, L->ev_fork, 0));
            int r0 = 0;
            for (int b = 0; b < nb; ++b) {
                const int r1 = (b == nb - 1) ? g.h : yb[b + 1] + 2;  // interior rows [yb, yb1) read ROI rows [yb, yb1 + 2)
                SCB_CUDA(c, cudaMemcpy2DAsync(w.stD + (size_t)r0 * w.pD, (size_t)w.pD, dROI + (size_t)r0 * dst->stride, (size_t)dst->stride, (size_t)3 * g.w, (size_t)(r1 - r0),
                                              cudaMemcpyHostToDevice, L->copy));
                SCB_CUDA(c, cudaMemcpy2DAsync(w.stS + (size_t)r0 * w.pS, (size_t)w.pS, sROI + (size_t)r0 * src->stride, (size_t)src->stride, (size_t)3 * g.w, (size_t)(r1 - r0),
                                              cudaMemcpyHostToDevice, L->copy));
                SCB_CUDA(c, cudaEventRecord(L->ev_band[b], L->copy));
                r0 = r1;
            }
        }
    } else {
        if (copy_dst && !side_copy) SCB_CUDA(c, cudaMemcpy2DAsync(blend->data, (size_t)blend->stride, dst->data, (size_t)dst->stride, row_bytes, (size_t)p->dst_rows, cudaMemcpyDeviceToDevice, ms));
        st = make_stencil(p, dROI, dst->stride, sROI, src->stride);
        out = bInt;
        out_pitch = blend->stride;
    }
    if (p->debug) {
        SCB_LAUNCH(gradients_dump_kernel, dim3((g.w + 31) / 32, (g.h + 7) / 8), dim3(256), 0, ms, st, p->dbg_vx, p->dbg_vy, p->mode);
        c->launches++;
    }
    tm.mark(ST_IN);
    const int gpl = swap ? (int)align_up((size_t)g.ny, 4) : w.gp;  // pitch of a line of G in the solve's frame
    const bool fused = i8_fused_rhs(p);
    if (nb == 1) {
        if (fused)
            run_rhs_fold(p, st, w, 0, g.ny);
        else
            run_rhs(p, st, w.G, w.gp, 0, g.ny, swap);
    } else {
        for (int b = 0; b < nb; ++b) {
            SCB_CUDA(c, cudaStreamWaitEvent(ms, L->ev_band[b], 0));
            if (fused)
                run_rhs_fold(p, st, w, yb[b], yb[b + 1]);
            else
                run_rhs(p, st, w.G, w.gp, yb[b], yb[b + 1], swap);
            if (b + 1 < nb && !swap) {  // the last band's rows follow the refinement fork
                if (p->use_i8) {
                    if ((rc = run_i8_forward(p, w, w.G, w.gp, w.At, w.R, yb[b], yb[b + 1]))) return rc;
                } else {
                    run_rows_fwd(p, st, w.G, w.gp, w.At, yb[b], yb[b + 1], p->use_tri);
                }
            }
        }
    }
    tm.mark(ST_RHS);
    if (serial) {  // stage timing serialises the refinement so that every stage has its own event pair
        if (!p->use_i8) run_lowfreq_rows(p, st, w.G, gpl, w.R, 0, fr.cnt, ms, swap);  // the INT8 pass delivers the exact row sums itself
        if (!p->use_tri) run_lowfreq_cols(p, w.R, w.lowspec, ms);
    } else {      // production: the refinement (small CTAs, no smem) co-runs with pass A (1 big CTA per SM)
        SCB_CUDA(c, cudaEventRecord(L->ev_fork, ms));
        SCB_CUDA(c, cudaStreamWaitEvent(L->side, L->ev_fork, 0));
        if (!p->use_i8) run_lowfreq_rows(p, st, w.G, gpl, w.R, 0, fr.cnt, L->side, swap);
        if (!p->use_tri) run_lowfreq_cols(p, w.R, w.lowspec, L->side);
        SCB_CUDA(c, cudaEventRecord(L->ev_join, L->side));
        if (side_copy) {  // blend = dst.copy() rides along on the side stream; only the compose pass has to wait for it
            if ((size_t)blend->stride == row_bytes && (size_t)dst->stride == row_bytes)
                SCB_CUDA(c, cudaMemcpyAsync(blend->data, dst->data, row_bytes * (size_t)p->dst_rows, cudaMemcpyDeviceToDevice, L->side));
            else
                SCB_CUDA(c, cudaMemcpy2DAsync(blend->data, (size_t)blend->stride, dst->data, (size_t)dst->stride, row_bytes, (size_t)p->dst_rows, cudaMemcpyDeviceToDevice, L->side));
            SCB_CUDA(c, cudaEventRecord(L->ev_copy, L->side));
        }
    }
    tm.mark(ST_LOW);
    if (p->use_tc) {
        if (side_copy) SCB_CUDA(c, cudaStreamWaitEvent(ms, L->ev_copy, 0));
        if ((rc = tc_solve(p, w, out, out_pitch, tm))) return rc;
    } else {
        if (p->use_i8) {
            if ((rc = run_i8_forward(p, w, w.G, w.gp, w.At, w.R, yb[nb - 1], g.ny, &tm))) return rc;
        } else {
            run_rows_fwd(p, st, w.G, gpl, w.At, swap ? 0 : yb[nb - 1], fr.cnt, p->use_tri, swap);
        }
        if (!serial) SCB_CUDA(c, cudaStreamWaitEvent(ms, L->ev_join, 0));
        tm.mark(ST_ROWS_FWD);
        if (p->use_tri) {
            if ((rc = run_tri(p, w.At, w.Ct, w.R, w.Y64, w.W, serial ? ms : L->side, swap))) return rc;
        } else
            run_cols(p, w.At, w.Ct, w.lowspec, 0, g.nx);
        tm.mark(ST_COLS);
        if (side_copy) SCB_CUDA(c, cudaStreamWaitEvent(ms, L->ev_copy, 0));
        if (nb_out == 1) {
            if (p->use_i8) {
                const bool fc = i8_fused_compose(p);
                if ((rc = run_i8_inverse(p, w, w.Ct, w.At, 0, g.ny, &tm, fc ? out : nullptr, out_pitch))) return rc;  // the row-transformed right-hand side is dead: U overwrites it
                if (!fc) run_compose(p, w.At, out, out_pitch, 0, g.ny);
            } else {
                run_rows_inv(p, w.Ct, out, out_pitch, 0, fr.cnt, swap);
            }
        } else {
            for (int b = 0; b < nb; ++b) {
                if (p->use_i8) {
                    const bool fc = i8_fused_compose(p);
                    if ((rc = run_i8_inverse(p, w, w.Ct, w.At, yb[b], yb[b + 1], nullptr, fc ? out : nullptr, out_pitch))) return rc;
                    if (!fc) run_compose(p, w.At, out, out_pitch, yb[b], yb[b + 1]);
                } else {
                    run_rows_inv(p, w.Ct, out, out_pitch, yb[b], yb[b + 1]);
                }
                SCB_CUDA(c, cudaEventRecord(L->ev_out[b], ms));
                SCB_CUDA(c, cudaStreamWaitEvent(L->copy, L->ev_out[b], 0));
                SCB_CUDA(c, cudaMemcpy2DAsync(bInt + (size_t)yb[b] * blend->stride, (size_t)blend->stride, w.stO + (size_t)yb[b] * w.pO, (size_t)w.pO, (size_t)3 * g.nx,
                                              (size_t)(yb[b + 1] - yb[b]), cudaMemcpyDeviceToHost, L->copy));
            }
            SCB_CUDA(c, cudaEventRecord(L->ev_copy, L->copy));
            SCB_CUDA(c, cudaStreamWaitEvent(ms, L->ev_copy, 0));  // a sync of the lane's stream covers the downloads
        }
        tm.mark(ST_ROWS_INV);
    }
    SCB_CUDA(c, cudaGetLastError());
    if (host) {
        // only the ROI interior comes back; everything else of blend is copied from dst on the host while the GPU works
        if (nb_out == 1) SCB_CUDA(c, cudaMemcpy2DAsync(bInt, (size_t)blend->stride, w.stO, (size_t)w.pO, (size_t)3 * g.nx, (size_t)g.ny, cudaMemcpyDeviceToHost, ms));
        tm.mark(ST_OUT);
        if (!defer_host) {
            if (copy_dst) host_copy_outside(c, dst, blend, g, host_threads());
            SCB_CUDA(c, cudaStreamSynchronize(ms));
            SCB_CUDA(c, cudaGetLastError());
        }
    } else {
        tm.mark(ST_OUT);
    }
    return SCB_OK;
}

extern "C" int scb_plan_execute(scb_plan* p, const scb_image* src, const scb_image* dst, scb_image* blend, int mem_kind, int exec_flags) {
    StageTimer tm;
    return execute_impl(p, src, dst, blend, mem_kind, exec_flags, tm);
}

// Same as scb_plan_execute, with CUDA events between the stages on the context stream; returns after a
// stream sync.  stage_ms[7] = { input copies, RHS stencil, low-frequency refinement, rows forward, columns, rows inverse, output copy }.
extern "C" int scb_plan_execute_timed(scb_plan* p, const scb_image* src, const scb_image* dst, scb_image* blend, int mem_kind, int exec_flags, float* stage_ms) {
    return scb_plan_execute_timed_i8(p, src, dst, blend, mem_kind, exec_flags, stage_ms, nullptr);
}

// The same, plus the kernels of the INT8 passes on their own: i8_ms[5] = { digitise forward, GEMM forward, digitise inverse, GEMM inverse,
// compose } (zeros for plans on another engine).  rows forward = i8_ms[0] + i8_ms[1], rows inverse = i8_ms[2] + i8_ms[3] + i8_ms[4].
extern "C" int scb_plan_execute_timed_i8(scb_plan* p, const scb_image* src, const scb_image* dst, scb_image* blend, int mem_kind, int exec_flags, float* stage_ms,
                                         float* i8_ms) {
    if (!p || !stage_ms) return SCB_ERR_INVALID_ARGUMENT;
    scb_context* c = p->ctx;
    SCB_CUDA(c, cudaSetDevice(c->device));
    StageTimer tm;
    tm.on = true;
    tm.stream = p->lane->stream;
    for (int i = 0; i < ST_COUNT; ++i) SCB_CUDA(c, cudaEventCreate(&tm.ev[i]));
    for (int i = 0; i < kStages; ++i) stage_ms[i] = 0.f;
    if (i8_ms)
        for (int i = 0; i < 5; ++i) i8_ms[i] = 0.f;
    int rc = execute_impl(p, src, dst, blend, mem_kind, exec_flags, tm);
    if (rc == SCB_OK && !p->g.empty) {
        cudaStreamSynchronize(p->lane->stream);
        for (int i = 0; i < kStages; ++i) cudaEventElapsedTime(&stage_ms[i], tm.ev[i], tm.ev[i + 1]);
        if (i8_ms && p->use_i8) {
            cudaEventElapsedTime(&i8_ms[0], tm.ev[ST_LOW], tm.ev[ST_X_DIGF]);
            cudaEventElapsedTime(&i8_ms[1], tm.ev[ST_X_DIGF], tm.ev[ST_ROWS_FWD]);
            cudaEventElapsedTime(&i8_ms[2], tm.ev[ST_COLS], tm.ev[ST_X_DIGI]);
            cudaEventElapsedTime(&i8_ms[3], tm.ev[ST_X_DIGI], tm.ev[ST_X_GEMMI]);
            cudaEventElapsedTime(&i8_ms[4], tm.ev[ST_X_GEMMI], tm.ev[ST_ROWS_INV]);
        }
    }
    for (int i = 0; i < ST_COUNT; ++i) cudaEventDestroy(tm.ev[i]);
    return rc;
}

extern "C" int scb_seamless_clone(scb_context* c, const scb_image* src, const scb_image* dst, const scb_image* mask,
                                  int px, int py, scb_image* blend, int clone_flags, int mem_kind) {
    if (!c) return SCB_ERR_INVALID_ARGUMENT;
    if (!src || !dst || !blend) return fail(c, SCB_ERR_INVALID_ARGUMENT, "seamlessClone: null image");
    const bool cacheable = mem_kind == SCB_MEM_HOST && c->plan_cache_cap > 0 && mask && mask->data && mask->channels == 1 && mask->rows > 0 && mask->cols > 0 &&
                           mask->stride >= mask->cols;
    if (!cacheable) {
        scb_plan* p = nullptr;
        int rc = scb_plan_create_ex(c, mask, mem_kind, src->rows, src->cols, dst->rows, dst->cols, px, py, clone_flags, &p);
        if (rc) return rc;
        rc = scb_plan_execute(p, src, dst, blend, mem_kind, SCB_EXEC_DEFAULT);
        scb_plan_destroy(p);
        return rc;
    }
    // plan cache: everything a plan depends on is in the key (the mask by its hash)
    scb_context::CachedPlan want;
    MaskScan scan;
    {
        NvtxRange nvtx_("scb:mask_scan");
        scan = scan_mask(c, mask, true);  // hash (the cache key) and bounding box (a miss needs it) in one pass
        want.hash = scan.hash;
    }
    const int key[11] = {mask->rows, mask->cols, src->rows, src->cols, dst->rows, dst->cols, px, py, clone_flags, wanted_engine(c), c->orientation};
    std::memcpy(want.key, key, sizeof(key));
    scb_plan* p = nullptr;
    for (auto& cp : c->plan_cache)
        if (cp.hash == want.hash && std::memcmp(cp.key, want.key, sizeof(key)) == 0) {
            cp.last_use = ++c->use_clock;
            p = cp.plan;
            c->plan_hits++;
            break;
        }
    if (!p) {
        c->plan_misses++;
        int rc = plan_create_impl(c, mask, mem_kind, src->rows, src->cols, dst->rows, dst->cols, px, py, clone_flags, &p, /*sync_mask=*/false, &scan);
        if (rc) return rc;
        if ((int)c->plan_cache.size() >= c->plan_cache_cap) {  // evict the least recently used plan
            size_t lru = 0;
            for (size_t i = 1; i < c->plan_cache.size(); ++i)
                if (c->plan_cache[i].last_use < c->plan_cache[lru].last_use) lru = i;
            scb_plan_destroy(c->plan_cache[lru].plan);
            c->plan_cache.erase(c->plan_cache.begin() + (long)lru);
        }
        want.plan = p;
        want.last_use = ++c->use_clock;
        c->plan_cache.push_back(want);
        const int rc2 = scb_plan_execute(p, src, dst, blend, mem_kind, SCB_EXEC_DEFAULT);
        // the plan was created without waiting for its mask upload: a HOST execute ends with a stream sync that covers it, except
        // for an empty plan or a failed execute -- wait here then, so that the caller's mask buffer is always free on return
        if (rc2 != SCB_OK || p->g.empty) cudaStreamSynchronize(c->lanes[0].stream);
        return rc2;
    }
    return scb_plan_execute(p, src, dst, blend, mem_kind, SCB_EXEC_DEFAULT);
}
extern "C" int scb_plan_cache_stats(const scb_context* c, uint64_t* hits, uint64_t* misses) {
    if (!c) return SCB_ERR_INVALID_ARGUMENT;
    if (hits) *hits = c->plan_hits;
    if (misses) *misses = c->plan_misses;
    return SCB_OK;
}

// Replays the plan's device-resident launches (D2D copy of dst, stencil, refinement on the side stream,
// three transform passes) as ONE CUDA graph launch -- the per-frame call of a fixed-mask video stream
// (BASELINE cfg5).  The graph is captured on first use and re-captured when a pointer, stride or the
// workspace arena changes.  The reference has no counterpart (47 launches + 2 host syncs per frame).
extern "C" int scb_plan_execute_graph(scb_plan* p, const scb_image* src, const scb_image* dst, scb_image* blend, int exec_flags) {
    if (!p) return SCB_ERR_INVALID_ARGUMENT;
    StageTimer tm;
#ifdef SCB_EMU
    return execute_impl(p, src, dst, blend, SCB_MEM_DEVICE, exec_flags, tm);
#else
    scb_context* c = p->ctx;
    int rc;
    if ((rc = check_image(c, src, p->src_rows, p->src_cols, "src"))) return rc;
    if ((rc = check_image(c, dst, p->dst_rows, p->dst_cols, "dst"))) return rc;
    if ((rc = check_image(c, blend, p->dst_rows, p->dst_cols, "blend"))) return rc;
    if (p->g.empty) return execute_impl(p, src, dst, blend, SCB_MEM_DEVICE, exec_flags, tm);
    SCB_CUDA(c, cudaSetDevice(c->device));
    Workspace w;
    if ((rc = carve(p, false, &w))) return rc;  // grows the arena (sync + cudaMalloc) outside the capture
    GraphKey k;
    k.src = src->data;
    k.dst = dst->data;
    k.blend = blend->data;
    k.s_stride = src->stride;
    k.d_stride = dst->stride;
    k.b_stride = blend->stride;
    k.flags = exec_flags;
    k.ws_epoch = p->lane->ws_epoch;
    if (!p->graph_exec || !(k == p->graph_key)) {
        plan_drop_graph(p);
        const uint64_t before = c->launches;
        SCB_CUDA(c, cudaStreamBeginCapture(p->lane->stream, cudaStreamCaptureModeThreadLocal));
        rc = execute_impl(p, src, dst, blend, SCB_MEM_DEVICE, exec_flags, tm);
        cudaGraph_t graph = nullptr;
        cudaError_t e = cudaStreamEndCapture(p->lane->stream, &graph);
        p->graph_kernels = c->launches - before;
        c->launches = before;  // captured, not run
        if (rc != SCB_OK || e != cudaSuccess) {
            if (graph) cudaGraphDestroy(graph);
            return rc != SCB_OK ? rc : fail(c, SCB_ERR_CUDA, std::string("cudaStreamEndCapture: ") + cudaGetErrorString(e));
        }
        e = cudaGraphInstantiate(&p->graph_exec, graph, 0);
        cudaGraphDestroy(graph);
        if (e != cudaSuccess) {
            p->graph_exec = nullptr;
            return fail(c, SCB_ERR_CUDA, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(e));
        }
        p->graph_key = k;
    }
    SCB_CUDA(c, cudaGraphLaunch(p->graph_exec, p->lane->stream));
    c->launches += p->graph_kernels;
    return SCB_OK;
#endif
}

// Independent jobs (BASELINE cfg3).  Jobs are taken in chunks: the mask uploads and bounding-box
// reductions of a whole chunk go onto a dedicated prep stream while the lanes still run the previous
// chunk, ONE host sync fetches every bounding box, then each job (erosion, ROI upload, solve, ROI
// download) is queued on lane i mod L without further syncs.  HOST results are complete on return;
// DEVICE results after the trailing sync (also done here).
extern "C" int scb_clone_batch(scb_context* c, scb_job* jobs, int n_jobs, int mem_kind) {
    if (!c || (!jobs && n_jobs > 0)) return SCB_ERR_INVALID_ARGUMENT;
    if (n_jobs <= 0) return SCB_OK;
    if (mem_kind != SCB_MEM_HOST && mem_kind != SCB_MEM_DEVICE) return fail(c, SCB_ERR_INVALID_ARGUMENT, "scb_clone_batch: bad mem kind");
    SCB_CUDA(c, cudaSetDevice(c->device));
    // jobs are planned in chunks (one sync of the prep stream per chunk); ~8 chunks per call keep the lanes fed while the
    // next chunk's bounding boxes come back, also when a rank of an 8-GPU batch only holds a few dozen jobs
    static const int kMaxChunk = 64;
    int kChunk = (n_jobs + 7) / 8;
    if (kChunk < 8) kChunk = 8;
    if (kChunk > kMaxChunk) kChunk = kMaxChunk;
    if (const char* e = std::getenv("SCB_CHUNK")) kChunk = std::atoi(e) > 0 && std::atoi(e) <= kMaxChunk ? std::atoi(e) : kChunk;
    int rc;
    int want_lanes = kDefaultLanes;
    if (const char* e = std::getenv("SCB_LANES")) want_lanes = std::atoi(e) > 0 ? std::atoi(e) : kDefaultLanes;
    if (want_lanes > n_jobs) want_lanes = n_jobs;
    if (want_lanes < c->n_lanes) want_lanes = c->n_lanes;
    if ((rc = ensure_lanes(c, want_lanes))) return rc;
    if ((rc = ensure_bbox_slots(c, 2 * kMaxChunk))) return rc;
    const int L = c->n_lanes;
    if (!c->prep) {  // high priority: the next chunk's bounding boxes must not queue behind the lanes' transform kernels
#ifdef SCB_EMU
        SCB_CUDA(c, cudaStreamCreateWithFlags(&c->prep, cudaStreamNonBlocking));
#else
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);
        SCB_CUDA(c, cudaStreamCreateWithPriority(&c->prep, cudaStreamNonBlocking, hi));
#endif
    }
    int worst = SCB_OK;
    std::string first_error;
    std::mutex note_mu;
    auto note = [&](scb_job& j, int status) {
        j.status = status;
        if (status != SCB_OK) {
            std::lock_guard<std::mutex> lk(note_mu);
            if (worst == SCB_OK) {
                worst = status;
                std::lock_guard<std::mutex> lk2(c->err_mu);
                first_error = c->err;
            }
        }
    };
    // Two chunk states: while this thread queues the solves of chunk k on the lanes, a helper thread plans chunk k+1
    // (mask uploads, bounding boxes, the one sync of the prep stream, geometry, erosion, tables) -- the batch path is
    // host-bound (~3 us per CUDA call, ~20 calls per job), so the two phases are worth overlapping.
    struct Chunk {
        std::vector<scb_plan*> plans;
        std::vector<PlanInput> inputs;
        std::vector<int> planned;  // plan_finish succeeded
        int base = 0, m = 0, rc = SCB_OK;
    } chunks[2];
    for (auto& ch : chunks) {
        ch.plans.resize(kChunk);
        ch.inputs.resize(kChunk);
        ch.planned.resize(kChunk);
    }
    std::vector<scb_geometry> geoms(kChunk);
    std::vector<char> copy_ok(kChunk);
    // SCB_BATCH_TRACE=1: host wall time of each phase of the batch (stderr), for tuning
    static const bool trace = std::getenv("SCB_BATCH_TRACE") != nullptr;
    double t_phase[5] = {0, 0, 0, 0, 0};
    auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    auto plan_chunk = [&](Chunk& ch, int base, int parity) {
        ch.base = base;
        ch.m = (n_jobs - base < kChunk) ? n_jobs - base : kChunk;
        ch.rc = SCB_OK;
        if (cudaSetDevice(c->device) != cudaSuccess) {
            ch.rc = fail(c, SCB_ERR_CUDA, "scb_clone_batch: cudaSetDevice failed on the planning thread");
            ch.m = 0;
            return;
        }
        double t0 = now();
        for (int i = 0; i < ch.m; ++i) {
            scb_job& j = jobs[base + i];
            ch.plans[i] = nullptr;
            ch.planned[i] = 0;
            if (!j.src.data || !j.dst.data || !j.blend.data) {
                note(j, fail(c, SCB_ERR_INVALID_ARGUMENT, "scb_clone_batch: null image"));
                continue;
            }
            // (the mask scan of a HOST job runs on THIS thread: the context's helper pool is busy submitting the previous chunk)
            note(j, plan_begin(c, &c->lanes[(base + i) % L], c->prep, &j.mask, mem_kind, j.src.rows, j.src.cols, j.dst.rows, j.dst.cols, j.px, j.py, parity * kMaxChunk + i,
                               &ch.plans[i], &ch.inputs[i], j.flags ? j.flags : SCB_NORMAL_CLONE, nullptr, /*parallel_scan=*/false));
        }
        double t1 = now();
        cudaError_t e = cudaStreamSynchronize(c->prep);
        if (e == cudaSuccess) e = cudaGetLastError();
        if (e != cudaSuccess) {
            ch.rc = fail(c, SCB_ERR_CUDA, std::string("scb_clone_batch: ") + cudaGetErrorString(e));
            for (int i = 0; i < ch.m; ++i)
                if (ch.plans[i]) {
                    scb_plan_destroy(ch.plans[i]);
                    ch.plans[i] = nullptr;
                }
            return;
        }
        double t2 = now();
        for (int i = 0; i < ch.m; ++i) {
            if (!ch.plans[i]) continue;
            const int st = plan_finish(ch.plans[i], ch.inputs[i]);  // destroys the plan on failure
            if (st != SCB_OK) {
                ch.plans[i] = nullptr;
                note(jobs[base + i], st);
            } else {
                ch.planned[i] = 1;
            }
        }
        double t3 = now();
        t_phase[0] += t1 - t0;
        t_phase[1] += t2 - t1;
        t_phase[2] += t3 - t2;
    };
    const int n_chunks = (n_jobs + kChunk - 1) / kChunk;
    plan_chunk(chunks[0], 0, 0);
    for (int k = 0; k < n_chunks; ++k) {
        Chunk& ch = chunks[k & 1];
        if (ch.rc != SCB_OK) return ch.rc;
        std::thread planner;
        const bool more = k + 1 < n_chunks;
#ifdef SCB_EMU
        if (more) plan_chunk(chunks[(k + 1) & 1], (k + 1) * kChunk, (k + 1) & 1);  // the interpreter runs kernels on the calling thread: no helper
#else
        if (more) planner = std::thread(plan_chunk, std::ref(chunks[(k + 1) & 1]), (k + 1) * kChunk, (k + 1) & 1);
#endif
        const int base = ch.base, m = ch.m;
        double b0 = now();
        // The lanes are fed from several host threads: a job costs ~25 CUDA calls of ~3 us each on one thread, more than its kernels
        // take on the GPU.  Thread t owns the lanes l with l % T == t, so the jobs of a lane stay in order.  (SCB_SUBMIT_THREADS)
        auto submit = [&](int t, int T) {
            if (T > 1 && cudaSetDevice(c->device) != cudaSuccess) return;
            for (int i = 0; i < m; ++i) {
                if (((base + i) % L) % T != t) continue;
                copy_ok[i] = 0;
                geoms[i] = scb_geometry();
                if (!ch.plans[i] || !ch.planned[i]) continue;
                scb_job& j = jobs[base + i];
                StageTimer tm;
                geoms[i] = ch.plans[i]->g;
                const int st = execute_impl(ch.plans[i], &j.src, &j.dst, &j.blend, mem_kind, SCB_EXEC_DEFAULT, tm, /*defer_host=*/true);
                scb_plan_destroy(ch.plans[i]);  // stream-ordered frees: the queued kernels finish first
                ch.plans[i] = nullptr;
                copy_ok[i] = (st == SCB_OK);
                note(j, st);
            }
        };
#ifdef SCB_EMU
        submit(0, 1);  // the interpreter runs kernels on the calling thread
#else
        {
            static const int want = [] {
                const char* e = std::getenv("SCB_SUBMIT_THREADS");
                return e ? std::atoi(e) : 4;
            }();
            int T = want < 1 ? 1 : want;
            if (T > L) T = L;
            if (T > host_threads()) T = host_threads();
            if (T <= 1)
                submit(0, 1);
            else
                pool_of(c).run(T, [&](int t) { submit(t, T); });
        }
#endif
        if (mem_kind == SCB_MEM_HOST) {  // blend = dst outside each ROI interior: a parallel-for over the chunk's jobs
            const int T = host_threads();
            pool_of(c).run(T, [&](int t) {
                for (int i = t; i < m; i += T) {
                    scb_job& j = jobs[base + i];
                    if (copy_ok[i] && j.blend.data != j.dst.data) host_copy_rows(&j.dst, &j.blend, geoms[i], 0, j.dst.rows);
                }
            });
        }
        t_phase[3] += now() - b0;
        if (planner.joinable()) planner.join();
    }
    {
        double a = now();
        if ((rc = scb_sync(c))) return rc;
        t_phase[4] = now() - a;
    }
    if (trace)
        std::fprintf(stderr, "scb_clone_batch: %d jobs, %d lanes: plan_begin %.2f ms, prep syncs %.2f ms, plan_finish %.2f ms, execute+destroy %.2f ms, final sync %.2f ms\n",
                     n_jobs, L, t_phase[0], t_phase[1], t_phase[2], t_phase[3], t_phase[4]);
    if (worst != SCB_OK) c->err = first_error;
    return worst;
}

// ------------------------------------------------------------------------------------------------
// sharded single solve: the caller (one rank per GPU) owns At / Ct and exchanges them between passes
// ------------------------------------------------------------------------------------------------
static int check_range(scb_context* c, int a, int b, int n, const char* what) {
    if (a < 0 || b < a || b > n) return fail(c, SCB_ERR_INVALID_ARGUMENT, std::string(what) + ": bad range");
    return SCB_OK;
}

extern "C" int scb_plan_rows_forward(scb_plan* p, const scb_image* src, const scb_image* dst, int mem_kind, int y0, int y1, float* at_dev, double* lowrows_dev) {
    if (!p || !at_dev) return SCB_ERR_INVALID_ARGUMENT;
    scb_context* c = p->ctx;
    if (p->g.empty) return fail(c, SCB_ERR_INVALID_ARGUMENT, "sharded solve: empty plan");
    if (mem_kind != SCB_MEM_DEVICE) return fail(c, SCB_ERR_UNSUPPORTED, "sharded solve: images must be device resident");
    int rc;
    if ((rc = check_image(c, src, p->src_rows, p->src_cols, "src"))) return rc;
    if ((rc = check_image(c, dst, p->dst_rows, p->dst_cols, "dst"))) return rc;
    if ((rc = check_range(c, y0, y1, p->g.ny, "scb_plan_rows_forward"))) return rc;
    SCB_CUDA(c, cudaSetDevice(c->device));
    const scb_geometry& g = p->g;
    const unsigned char* dROI = (const unsigned char*)dst->data + (size_t)g.ry * dst->stride + (size_t)3 * g.rx;
    const unsigned char* sROI = (const unsigned char*)src->data + (size_t)g.y * src->stride + (size_t)3 * g.x;
    StencilSrc st = make_stencil(p, dROI, dst->stride, sROI, src->stride);
    Workspace w;
    if ((rc = carve(p, false, &w))) return rc;
    run_rhs(p, st, w.G, w.gp, y0, y1);
    if (lowrows_dev) run_lowfreq_rows(p, st, w.G, w.gp, lowrows_dev, y0, y1, p->lane->stream);
    run_rows_fwd(p, st, w.G, w.gp, at_dev, y0, y1);
    SCB_CUDA(c, cudaGetLastError());
    return SCB_OK;
}

extern "C" int scb_plan_lowfreq_finish(scb_plan* p, const double* lowrows_dev, float* lowspec_dev) {
    if (!p || !lowrows_dev || !lowspec_dev) return SCB_ERR_INVALID_ARGUMENT;
    scb_context* c = p->ctx;
    if (p->g.empty) return fail(c, SCB_ERR_INVALID_ARGUMENT, "sharded solve: empty plan");
    SCB_CUDA(c, cudaSetDevice(c->device));
    run_lowfreq_cols(p, lowrows_dev, lowspec_dev, p->lane->stream);
    SCB_CUDA(c, cudaGetLastError());
    return SCB_OK;
}

extern "C" int scb_plan_cols(scb_plan* p, int x0, int x1, const float* at_dev, float* ct_dev, const float* lowspec_dev) {
    if (!p || !at_dev || !ct_dev) return SCB_ERR_INVALID_ARGUMENT;
    scb_context* c = p->ctx;
    if (p->g.empty) return fail(c, SCB_ERR_INVALID_ARGUMENT, "sharded solve: empty plan");
    int rc;
    if ((rc = check_range(c, x0, x1, p->g.nx, "scb_plan_cols"))) return rc;
    SCB_CUDA(c, cudaSetDevice(c->device));
    run_cols(p, at_dev, ct_dev, lowspec_dev, x0, x1);
    SCB_CUDA(c, cudaGetLastError());
    return SCB_OK;
}

extern "C" int scb_plan_rows_inverse(scb_plan* p, const float* ct_dev, scb_image* blend, int mem_kind, int y0, int y1) {
    if (!p || !ct_dev) return SCB_ERR_INVALID_ARGUMENT;
    scb_context* c = p->ctx;
    if (p->g.empty) return fail(c, SCB_ERR_INVALID_ARGUMENT, "sharded solve: empty plan");
    if (mem_kind != SCB_MEM_DEVICE) return fail(c, SCB_ERR_UNSUPPORTED, "sharded solve: images must be device resident");
    int rc;
    if ((rc = check_image(c, blend, p->dst_rows, p->dst_cols, "blend"))) return rc;
    if ((rc = check_range(c, y0, y1, p->g.ny, "scb_plan_rows_inverse"))) return rc;
    SCB_CUDA(c, cudaSetDevice(c->device));
    const scb_geometry& g = p->g;
    unsigned char* bInt = (unsigned char*)blend->data + (size_t)(g.ry + 1) * blend->stride + (size_t)3 * (g.rx + 1);
    run_rows_inv(p, ct_dev, bInt, blend->stride, y0, y1);
    SCB_CUDA(c, cudaGetLastError());
    return SCB_OK;
}

// ---- row-sharded solve on the tridiagonal engine -------------------------------------------------
// A rank owns the segments [seg0, seg1) of every spectral column = the interior rows [seg0 L, min(ny, seg1 L)).  Nothing of
// the field ever moves between the ranks: the partitioned Thomas solve (scb_tri.cuh) couples the ranks only through the two
// end values of every segment's local solution and through the 32 x 32 low-frequency projections.
static int tri_shard_check(scb_plan* p, int seg0, int seg1, int* y0, int* y1) {
    scb_context* c = p->ctx;
    if (p->g.empty) return fail(c, SCB_ERR_INVALID_ARGUMENT, "sharded solve: empty plan");
    if (!p->use_tri || p->swap) return fail(c, SCB_ERR_UNSUPPORTED, "sharded tridiagonal solve: the plan must use SCB_ENGINE_TRI in its natural orientation");
    const int L = tri_seg_len(p->g.ny), nseg = (p->g.ny + L - 1) / L;
    if (seg0 < 0 || seg1 < seg0 || seg1 > nseg) return fail(c, SCB_ERR_INVALID_ARGUMENT, "sharded tridiagonal solve: segment range outside [0, n_segs]");
    *y0 = seg0 * L;
    *y1 = seg1 * L < p->g.ny ? seg1 * L : p->g.ny;
    return SCB_OK;
}

extern "C" int scb_plan_tri_layout(const scb_plan* p, int* seg_len, int* n_segs, size_t* ends32_floats, size_t* ends64_doubles, size_t* w_doubles) {
    if (!p || p->g.empty) return SCB_ERR_INVALID_ARGUMENT;
    const int L = tri_seg_len(p->g.ny);
    if (seg_len) *seg_len = L;
    if (n_segs) *n_segs = (p->g.ny + L - 1) / L;
    if (ends32_floats) *ends32_floats = (size_t)3 * kTriSegs * 2 * align_up((size_t)p->g.nx, 4);
    if (ends64_doubles) *ends64_doubles = (size_t)3 * kTriSegs * 2 * kTriLowK;
    if (w_doubles) *w_doubles = (size_t)3 * kTriLowL * kTriLowK;
    return SCB_OK;
}

extern "C" int scb_plan_tri_forward(scb_plan* p, const scb_image* src, const scb_image* dst, int mem_kind, int seg0, int seg1, float* ends32_dev, double* ends64_dev,
                                    double* w_dev) {
    if (!p || !ends32_dev || !ends64_dev || !w_dev) return SCB_ERR_INVALID_ARGUMENT;
    scb_context* c = p->ctx;
    if (mem_kind != SCB_MEM_DEVICE) return fail(c, SCB_ERR_UNSUPPORTED, "sharded solve: images must be device resident");
    int rc, y0, y1;
    if ((rc = tri_shard_check(p, seg0, seg1, &y0, &y1))) return rc;
    if ((rc = check_image(c, src, p->src_rows, p->src_cols, "src"))) return rc;
    if ((rc = check_image(c, dst, p->dst_rows, p->dst_cols, "dst"))) return rc;
    SCB_CUDA(c, cudaSetDevice(c->device));
    const scb_geometry& g = p->g;
    const unsigned char* dROI = (const unsigned char*)dst->data + (size_t)g.ry * dst->stride + (size_t)3 * g.rx;
    const unsigned char* sROI = (const unsigned char*)src->data + (size_t)g.y * src->stride + (size_t)3 * g.x;
    StencilSrc st = make_stencil(p, dROI, dst->stride, sROI, src->stride);
    Workspace w;
    if ((rc = carve(p, false, &w))) return rc;
    cudaStream_t ms = p->lane->stream;
    size_t n32, n64, nw;
    scb_plan_tri_layout(p, nullptr, nullptr, &n32, &n64, &nw);
    // zero-filled so that the ranks can combine their parts with a plain sum
    SCB_CUDA(c, cudaMemsetAsync(ends32_dev, 0, n32 * sizeof(float), ms));
    SCB_CUDA(c, cudaMemsetAsync(ends64_dev, 0, n64 * sizeof(double), ms));
    SCB_CUDA(c, cudaMemsetAsync(w_dev, 0, nw * sizeof(double), ms));
    if (i8_fused_rhs(p))
        run_rhs_fold(p, st, w, y0, y1);
    else
        run_rhs(p, st, w.G, w.gp, y0, y1);
    if (p->use_i8) {  // the INT8 pass delivers the exact low-frequency row sums itself
        if ((rc = run_i8_forward(p, w, w.G, w.gp, w.At, w.R, y0, y1))) return rc;
    } else {
        run_lowfreq_rows(p, st, w.G, w.gp, w.R, y0, y1, ms);
        run_rows_fwd(p, st, w.G, w.gp, w.At, y0, y1, /*natural=*/true);
    }
    const Frame f = frame_of(p, false);
    launch_tri_low(p, false, tri_low_params(p, f, w.At, w.Ct, w.R, w.Y64, w_dev, y0, y1), ms);
    TriSolveParams t = tri_solve_params(p, f, w.At, w.Ct, w.Y64);
    t.phase = 1;
    t.seg0 = seg0;
    t.seg1 = seg1;
    t.ends32 = ends32_dev;
    t.ends64 = ends64_dev;
    if (seg1 > seg0) launch_tri_solve(p, t);
    SCB_CUDA(c, cudaGetLastError());
    return SCB_OK;
}

extern "C" int scb_plan_tri_finish(scb_plan* p, scb_image* blend, int mem_kind, int seg0, int seg1, const float* ends32_dev, const double* ends64_dev, const double* w_dev) {
    return scb_plan_tri_finish_slots(p, blend, mem_kind, seg0, seg1, ends32_dev, ends64_dev, w_dev, 1);
}

extern "C" int scb_plan_tri_finish_slots(scb_plan* p, scb_image* blend, int mem_kind, int seg0, int seg1, const float* ends32_dev, const double* ends64_dev,
                                         const double* w_dev, int w_slots) {
    if (!p || !ends32_dev || !ends64_dev || !w_dev || w_slots < 1) return SCB_ERR_INVALID_ARGUMENT;
    scb_context* c = p->ctx;
    if (mem_kind != SCB_MEM_DEVICE) return fail(c, SCB_ERR_UNSUPPORTED, "sharded solve: images must be device resident");
    int rc, y0, y1;
    if ((rc = tri_shard_check(p, seg0, seg1, &y0, &y1))) return rc;
    if ((rc = check_image(c, blend, p->dst_rows, p->dst_cols, "blend"))) return rc;
    SCB_CUDA(c, cudaSetDevice(c->device));
    const scb_geometry& g = p->g;
    Workspace w;
    if ((rc = carve(p, false, &w))) return rc;  // same arena, same offsets as in scb_plan_tri_forward
    const Frame f = frame_of(p, false);
    TriSolveParams t = tri_solve_params(p, f, w.At, w.Ct, w.Y64);
    t.phase = 2;
    t.seg0 = seg0;
    t.seg1 = seg1;
    t.ends32 = const_cast<float*>(ends32_dev);
    t.ends64 = const_cast<double*>(ends64_dev);
    if (seg1 > seg0) launch_tri_solve(p, t);
    TriLowParams la = tri_low_params(p, f, w.At, w.Ct, w.R, w.Y64, const_cast<double*>(w_dev), y0, y1);
    la.w_slots = w_slots;
    launch_tri_low(p, true, la, p->lane->stream);
    unsigned char* bInt = (unsigned char*)blend->data + (size_t)(g.ry + 1) * blend->stride + (size_t)3 * (g.rx + 1);
    if (p->use_i8) {
        const bool fc = i8_fused_compose(p);
        if ((rc = run_i8_inverse(p, w, w.Ct, w.At, y0, y1, nullptr, fc ? bInt : nullptr, blend->stride))) return rc;
        if (!fc) run_compose(p, w.At, bInt, blend->stride, y0, y1);
    } else {
        run_rows_inv(p, w.Ct, bInt, blend->stride, y0, y1);
    }
    SCB_CUDA(c, cudaGetLastError());
    return SCB_OK;
}

// ------------------------------------------------------------------------------------------------
// self-test of one tensor-core pass against float64 direct sums (unit check of the tcgen05/TMA plumbing)
// ------------------------------------------------------------------------------------------------
extern "C" int scb_tc_selftest(scb_context* c, int n, int lines, int transposed, double* max_rel_err) {
    if (!c || !max_rel_err) return SCB_ERR_INVALID_ARGUMENT;
    if (n < kTcMinN || n > kTcMaxN || lines < 1) return fail(c, SCB_ERR_INVALID_ARGUMENT, "scb_tc_selftest: n or lines out of range");
    SCB_CUDA(c, cudaSetDevice(c->device));
    int rc = tc_configure_once(c);
    if (rc) return rc;
    const DevTcTab* tab = nullptr;
    if ((rc = get_tctab(c, nullptr, n, &tab))) return rc;
    const int pin = (int)align_up((size_t)n, 4), pl = (int)align_up((size_t)lines, 4);
    std::vector<float> hin((size_t)lines * pin, 0.f);
    unsigned long long seed = 0x9E3779B97F4A7C15ull ^ ((unsigned long long)n << 20) ^ (unsigned long long)lines;
    auto rnd = [&]() {
        seed = seed * 6364136223846793005ull + 1442695040888963407ull;
        return (double)(seed >> 11) / (double)(1ull << 53);
    };
    for (int l = 0; l < lines; ++l)
        for (int j = 0; j < n; ++j) hin[(size_t)l * pin + j] = (float)((rnd() - 0.5) * 2000.0);
    const size_t out_floats = transposed ? (size_t)n * pl : (size_t)lines * pin;
    float *din = nullptr, *dout = nullptr;
    SCB_CUDA(c, cudaMalloc(&din, hin.size() * sizeof(float)));
    SCB_CUDA(c, cudaMalloc(&dout, out_floats * sizeof(float)));
    cudaStream_t st = c->lanes[0].stream;
    SCB_CUDA(c, cudaMemcpyAsync(din, hin.data(), hin.size() * sizeof(float), cudaMemcpyHostToDevice, st));
    SCB_CUDA(c, cudaMemsetAsync(dout, 0xff, out_floats * sizeof(float), st));  // NaN pattern: unwritten outputs show up
    TcPassParams a{};
    a.tab = tab->dev;
    a.in = din;
    a.in_plane = 0;
    a.in_pitch = pin;
    a.lpc = lines;
    a.line0 = 0;
    a.line_end = lines;
    a.out = dout;
    a.out_plane = 0;
    a.out_pitch = transposed ? pl : pin;
    a.transposed = transposed;
    a.scale = 1.0f;
    {
        scb_plan fake;
        fake.ctx = c;
        fake.lane = &c->lanes[0];
        launch_tc_pass(&fake, tab, a);
    }
    std::vector<float> hout(out_floats);
    SCB_CUDA(c, cudaMemcpyAsync(hout.data(), dout, out_floats * sizeof(float), cudaMemcpyDeviceToHost, st));
    cudaError_t e = cudaStreamSynchronize(st);
    if (e == cudaSuccess) e = cudaGetLastError();
    cudaFree(din);
    cudaFree(dout);
    if (e != cudaSuccess) return fail(c, SCB_ERR_CUDA, std::string("scb_tc_selftest: ") + cudaGetErrorString(e));
    const double PI = 3.14159265358979323846;
    double worst = 0.0;
    const int step = lines > 24 ? lines / 24 : 1;
    std::vector<double> ref(n);
    for (int l = 0; l < lines; l += step) {
        double mx = 0.0;
        for (int k = 0; k < n; ++k) {
            double sum = 0.0;
            for (int j = 0; j < n; ++j) {
                const long long ph = ((long long)(j + 1) * (k + 1)) % (2LL * (n + 1));
                sum += (double)hin[(size_t)l * pin + j] * std::sin(PI * (double)ph / (double)(n + 1));
            }
            ref[k] = sum;
            if (std::fabs(sum) > mx) mx = std::fabs(sum);
        }
        for (int k = 0; k < n; ++k) {
            const float got = transposed ? hout[(size_t)k * pl + l] : hout[(size_t)l * pin + k];
            double err = std::fabs((double)got - ref[k]) / (mx > 0 ? mx : 1.0);
            if (!(err == err)) err = 1e30;  // NaN
            if (err > worst) worst = err;
        }
    }
    *max_rel_err = worst;
    return SCB_OK;
}

// ------------------------------------------------------------------------------------------------
// the reference's four names (seamlessclone_cuda.h:4-63)
// ------------------------------------------------------------------------------------------------
extern "C" void* my_seamlessclone_api_imp_create_instance(int gpu_id) {
    scb_context* c = nullptr;
    if (scb_create(gpu_id, nullptr, &c) != SCB_OK) {
        std::fprintf(stderr, "my_seamlessclone_api_imp_create_instance: %s\n", scb_last_error(nullptr));
        return nullptr;
    }
    return c;
}
extern "C" int my_seamlessclone_api_imp_run(void* inst, const scb_image* face, const scb_image* body, const scb_image* mask,
                                            int centerX, int centerY, int /*gpu_id*/, int bSync, scb_image* blend_out) {
    scb_context* c = (scb_context*)inst;
    if (!c) return SCB_ERR_INVALID_ARGUMENT;
    int rc = scb_seamless_clone(c, face, body, mask, centerX, centerY, blend_out, SCB_NORMAL_CLONE, SCB_MEM_HOST);
    if (rc == SCB_OK && bSync) rc = scb_sync(c);
    return rc;
}
extern "C" void my_seamlessclone_api_imp_destroy(void* inst) { scb_destroy((scb_context*)inst); }
extern "C" void my_seamlessclone_api_imp_sync(void* inst) {
    if (inst) scb_sync((scb_context*)inst);
}
