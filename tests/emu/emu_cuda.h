// emu_cuda.h -- a tiny SIMT interpreter for CI without a GPU.  TEST INFRASTRUCTURE ONLY.
//
// The product is hand-written CUDA for sm_100a and the package loads nothing else.  The
// authoring container has no GPU and a B200 round trip costs minutes of a small budget, so the
// same kernel sources are ALSO compiled by g++ with -DSCB_EMU against this header: every CUDA
// thread of a block becomes a ucontext fiber, __syncthreads()/__syncwarp()/shuffles become fiber
// barriers, blocks are spread over OS threads, and the handful of runtime calls the host driver
// uses are stubbed on top of malloc/memcpy.  The resulting library lives in tests/emu/_build/, is
// loaded explicitly by tests/ (never by the package) and is orders of magnitude slower than the
// oracle -- it checks indexing and host logic, it is not a fallback.
#pragma once

#include <ucontext.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <thread>
#include <vector>

// ------------------------------------------------------------------ qualifiers
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __launch_bounds__(...)
#define __shared__ static thread_local
#define __align__(n) __attribute__((aligned(n)))
#define __noinline__ __attribute__((noinline))

// ------------------------------------------------------------------ vector types
struct uint3 { unsigned x, y, z; };
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct float2 { float x, y; };
struct float4 { float x, y, z, w; };
struct double2 { double x, y; };
struct int2 { int x, y; };
struct int4 { int x, y, z, w; };
struct uint2 { unsigned x, y; };
struct uint4 { unsigned x, y, z, w; };
struct uchar3 { unsigned char x, y, z; };
struct uchar4 { unsigned char x, y, z, w; };
inline float2 make_float2(float x, float y) { return {x, y}; }
inline float4 make_float4(float x, float y, float z, float w) { return {x, y, z, w}; }
inline double2 make_double2(double x, double y) { return {x, y}; }
inline int2 make_int2(int x, int y) { return {x, y}; }
inline int4 make_int4(int x, int y, int z, int w) { return {x, y, z, w}; }
inline uint2 make_uint2(unsigned x, unsigned y) { return {x, y}; }
inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return {x, y, z, w}; }
inline uchar4 make_uchar4(unsigned char x, unsigned char y, unsigned char z, unsigned char w) { return {x, y, z, w}; }

inline thread_local uint3 threadIdx, blockIdx;
inline thread_local dim3 blockDim, gridDim;
static const int warpSize = 32;

using std::max;
using std::min;

// ------------------------------------------------------------------ scheduler
namespace emu {

struct Fiber {
    ucontext_t ctx;
    std::unique_ptr<char[]> stack;
    bool done = false;
    uint3 tid{};
    unsigned warp = 0, lane = 0;
};
struct Warp {
    unsigned live = 0, count = 0;
    unsigned long gen = 0;
    uint64_t slot[32];
};
struct Block {
    std::vector<Fiber> fibers;
    std::vector<Warp> warps;
    ucontext_t sched;
    Fiber* cur = nullptr;
    unsigned live = 0, bar_count = 0;
    unsigned long bar_gen = 0;
    unsigned named_count[16] = {0};
    unsigned long named_gen[16] = {0};
    std::function<void()> body;
    void* dyn = nullptr;
};
inline thread_local Block* g_blk = nullptr;
static const size_t kStack = 128 * 1024;

inline void yield() { swapcontext(&g_blk->cur->ctx, &g_blk->sched); }
inline void* dyn_smem() { return g_blk->dyn; }

inline void syncthreads() {
    Block* b = g_blk;
    unsigned long gen = b->bar_gen;
    if (++b->bar_count >= b->live) {
        b->bar_count = 0;
        b->bar_gen++;
        return;
    }
    while (b->bar_gen == gen) yield();
}
// bar.sync id, nthreads : named barrier among a fixed number of threads
inline void bar_sync(int id, unsigned nthreads) {
    Block* b = g_blk;
    unsigned long gen = b->named_gen[id];
    if (++b->named_count[id] >= nthreads) {
        b->named_count[id] = 0;
        b->named_gen[id]++;
        return;
    }
    while (b->named_gen[id] == gen) yield();
}
inline void syncwarp() {
    Block* b = g_blk;
    Warp& w = b->warps[b->cur->warp];
    unsigned long gen = w.gen;
    if (++w.count >= w.live) {
        w.count = 0;
        w.gen++;
        return;
    }
    while (w.gen == gen) yield();
}
template <class T>
inline T shfl(T v, int src) {
    static_assert(sizeof(T) <= 8, "shuffle payload");
    Block* b = g_blk;
    Warp& w = b->warps[b->cur->warp];
    uint64_t raw = 0;
    std::memcpy(&raw, &v, sizeof(T));
    w.slot[b->cur->lane] = raw;
    syncwarp();
    unsigned base = b->cur->warp * 32u;
    if (src < 0 || src > 31 || base + (unsigned)src >= b->fibers.size() || b->fibers[base + src].done) src = (int)b->cur->lane;
    uint64_t got = w.slot[src];
    syncwarp();
    T r;
    std::memcpy(&r, &got, sizeof(T));
    return r;
}

inline void trampoline() {
    Block* b = g_blk;
    b->body();
    b->cur->done = true;
}

inline void run_block(Block& b, unsigned nthreads, dim3 block) {
    g_blk = &b;
    if (b.fibers.size() != nthreads) {
        b.fibers.clear();
        b.fibers.resize(nthreads);
        for (auto& f : b.fibers) f.stack.reset(new char[kStack]);
    }
    unsigned nw = (nthreads + 31) / 32;
    b.warps.assign(nw, Warp());
    b.live = nthreads;
    b.bar_count = 0;
    b.bar_gen = 0;
    for (int k = 0; k < 16; ++k) b.named_count[k] = 0;
    for (unsigned t = 0; t < nthreads; ++t) {
        Fiber& f = b.fibers[t];
        f.done = false;
        f.tid.x = t % block.x;
        f.tid.y = (t / block.x) % block.y;
        f.tid.z = t / (block.x * block.y);
        f.warp = t / 32;
        f.lane = t % 32;
        b.warps[f.warp].live++;
        getcontext(&f.ctx);
        f.ctx.uc_stack.ss_sp = f.stack.get();
        f.ctx.uc_stack.ss_size = kStack;
        f.ctx.uc_link = &b.sched;
        makecontext(&f.ctx, (void (*)())trampoline, 0);
    }
    unsigned remaining = nthreads;
    unsigned long spins = 0;
    while (remaining) {
        unsigned progressed = 0;
        for (unsigned t = 0; t < nthreads; ++t) {
            Fiber& f = b.fibers[t];
            if (f.done) continue;
            b.cur = &f;
            ::threadIdx = f.tid;
            swapcontext(&b.sched, &f.ctx);
            if (f.done) {
                --remaining;
                ++progressed;
                --b.live;
                Warp& w = b.warps[f.warp];
                --w.live;
                if (w.live && w.count >= w.live) { w.count = 0; w.gen++; }
                if (b.live && b.bar_count >= b.live) { b.bar_count = 0; b.bar_gen++; }
            }
        }
        if (!progressed && ++spins > 1000000ul) {
            std::fprintf(stderr, "emu: deadlock in block (%u,%u,%u)\n", ::blockIdx.x, ::blockIdx.y, ::blockIdx.z);
            std::abort();
        }
        if (progressed) spins = 0;
    }
    g_blk = nullptr;
}

inline unsigned& worker_limit() {
    static unsigned n = [] {
        const char* e = std::getenv("SCB_EMU_THREADS");
        unsigned v = e ? (unsigned)std::atoi(e) : std::thread::hardware_concurrency();
        return v ? v : 1u;
    }();
    return n;
}

template <class... KA, class... A>
void launch(void (*kernel)(KA...), dim3 grid, dim3 block, size_t smem, A... args) {
    const unsigned nblocks = grid.x * grid.y * grid.z;
    const unsigned nthreads = block.x * block.y * block.z;
    if (!nblocks || !nthreads) return;
    unsigned nworkers = std::min(worker_limit(), nblocks);
    auto worker = [&](unsigned wid) {
        Block b;
        std::unique_ptr<char[]> dyn(new char[smem + 64]);
        b.dyn = (void*)(((uintptr_t)dyn.get() + 63) & ~(uintptr_t)63);
        b.body = [&]() { kernel(args...); };
        for (unsigned bid = wid; bid < nblocks; bid += nworkers) {
            ::gridDim = grid;
            ::blockDim = block;
            ::blockIdx.x = bid % grid.x;
            ::blockIdx.y = (bid / grid.x) % grid.y;
            ::blockIdx.z = bid / (grid.x * grid.y);
            run_block(b, nthreads, block);
        }
    };
    if (nworkers == 1) {
        worker(0);
    } else {
        std::vector<std::thread> th;
        for (unsigned w = 0; w < nworkers; ++w) th.emplace_back(worker, w);
        for (auto& t : th) t.join();
    }
}
}  // namespace emu

// ------------------------------------------------------------------ device intrinsics
inline void __syncthreads() { emu::syncthreads(); }
inline void __syncwarp(unsigned = 0xffffffffu) { emu::syncwarp(); }
template <class T> inline T __shfl_sync(unsigned, T v, int src, int = 32) { return emu::shfl(v, src); }
template <class T> inline T __shfl_xor_sync(unsigned, T v, int m, int = 32) { return emu::shfl(v, (int)(emu::g_blk->cur->lane ^ (unsigned)m)); }
template <class T> inline T __shfl_down_sync(unsigned, T v, unsigned d, int = 32) { return emu::shfl(v, (int)(emu::g_blk->cur->lane + d)); }
template <class T> inline T __shfl_up_sync(unsigned, T v, unsigned d, int = 32) { return emu::shfl(v, (int)emu::g_blk->cur->lane - (int)d); }
template <class T> inline T __ldg(const T* p) { return *p; }
inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }
inline float __fsub_rn(float a, float b) { volatile float r = a - b; return r; }
inline float __fdiv_rn(float a, float b) { volatile float r = a / b; return r; }
inline float __fmaf_rn(float a, float b, float c) { return std::fmaf(a, b, c); }
inline double __dmul_rn(double a, double b) { volatile double r = a * b; return r; }
inline double __dadd_rn(double a, double b) { volatile double r = a + b; return r; }
inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned shift) { return (unsigned)((((uint64_t)hi << 32) | lo) >> (shift & 31)); }
inline unsigned __byte_perm(unsigned x, unsigned y, unsigned s) {  // PRMT, default mode (selector nibbles 0..7, no sign replication)
    const uint64_t v = ((uint64_t)y << 32) | x;
    unsigned r = 0;
    for (int i = 0; i < 4; ++i) r |= (unsigned)((v >> (8 * ((s >> (4 * i)) & 7))) & 0xff) << (8 * i);
    return r;
}
inline float __uint_as_float(unsigned u) { float f; std::memcpy(&f, &u, 4); return f; }
inline unsigned __float_as_uint(float f) { unsigned u; std::memcpy(&u, &f, 4); return u; }
inline int __float2int_rz(float f) { return (int)f; }
inline int __float2int_rn(float f) { return (int)std::nearbyintf(f); }
inline unsigned __float2uint_rz(float f) { return (unsigned)f; }
inline float __int2float_rn(int i) { return (float)i; }
inline double sinpi(double x) { return std::sin(M_PI * x); }
inline double cospi(double x) { return std::cos(M_PI * x); }
inline void __threadfence() { std::atomic_thread_fence(std::memory_order_seq_cst); }

inline int atomicMin(int* a, int v) { int o = __atomic_load_n(a, __ATOMIC_RELAXED); while (v < o && !__atomic_compare_exchange_n(a, &o, v, true, __ATOMIC_SEQ_CST, __ATOMIC_RELAXED)) {} return o; }
inline int atomicMax(int* a, int v) { int o = __atomic_load_n(a, __ATOMIC_RELAXED); while (v > o && !__atomic_compare_exchange_n(a, &o, v, true, __ATOMIC_SEQ_CST, __ATOMIC_RELAXED)) {} return o; }
inline int atomicAdd(int* a, int v) { return __atomic_fetch_add(a, v, __ATOMIC_SEQ_CST); }
inline unsigned atomicAdd(unsigned* a, unsigned v) { return __atomic_fetch_add(a, v, __ATOMIC_SEQ_CST); }
inline unsigned long long atomicAdd(unsigned long long* a, unsigned long long v) { return __atomic_fetch_add(a, v, __ATOMIC_SEQ_CST); }
template <class F> inline F emu_atomic_add_fp(F* a, F v) {
    using U = typename std::conditional<sizeof(F) == 4, uint32_t, uint64_t>::type;
    U* ua = reinterpret_cast<U*>(a);
    U o = __atomic_load_n(ua, __ATOMIC_RELAXED), n;
    F of;
    do { std::memcpy(&of, &o, sizeof(F)); F nf = of + v; std::memcpy(&n, &nf, sizeof(F)); } while (!__atomic_compare_exchange_n(ua, &o, n, true, __ATOMIC_SEQ_CST, __ATOMIC_RELAXED));
    return of;
}
inline float atomicAdd(float* a, float v) { return emu_atomic_add_fp(a, v); }
inline double atomicAdd(double* a, double v) { return emu_atomic_add_fp(a, v); }

// ------------------------------------------------------------------ runtime stubs
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorInvalidValue = 1, cudaErrorMemoryAllocation = 2, cudaErrorNotSupported = 801 };
struct emu_stream_t { int dummy; };
typedef emu_stream_t* cudaStream_t;
struct emu_event_t { int dummy; };
typedef emu_event_t* cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost = 0, cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3, cudaMemcpyDefault = 4 };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2, cudaHostAllocDefault = 0 };
enum cudaDeviceAttr { cudaDevAttrMultiProcessorCount = 16, cudaDevAttrMaxSharedMemoryPerBlockOptin = 97, cudaDevAttrComputeCapabilityMajor = 75, cudaDevAttrComputeCapabilityMinor = 76 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8, cudaFuncAttributePreferredSharedMemoryCarveout = 9 };

inline cudaError_t emu_malloc(void** p, size_t n) {
    size_t r = (n + 255) & ~(size_t)255;
    if (!r) r = 256;
    void* q = std::aligned_alloc(256, r);
    if (!q) return cudaErrorMemoryAllocation;
    std::memset(q, 0xFF, r);  // NaN-poison: reading unwritten floats shows up in the tests
    *p = q;
    return cudaSuccess;
}
template <class T> inline cudaError_t cudaMalloc(T** p, size_t n) { return emu_malloc((void**)p, n); }
template <class T> inline cudaError_t cudaMallocHost(T** p, size_t n) { return emu_malloc((void**)p, n); }
inline cudaError_t cudaFree(void* p) { std::free(p); return cudaSuccess; }
inline cudaError_t cudaFreeHost(void* p) { std::free(p); return cudaSuccess; }
inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { std::memcpy(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t = nullptr) { std::memcpy(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemcpy2DAsync(void* d, size_t dp, const void* s, size_t sp, size_t wbytes, size_t rows, cudaMemcpyKind, cudaStream_t = nullptr) {
    for (size_t r = 0; r < rows; ++r) std::memcpy((char*)d + r * dp, (const char*)s + r * sp, wbytes);
    return cudaSuccess;
}
inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t = nullptr) { std::memset(d, v, n); return cudaSuccess; }
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = new emu_stream_t(); return cudaSuccess; }
inline cudaError_t cudaStreamDestroy(cudaStream_t s) { delete s; return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = new emu_event_t(); return cudaSuccess; }
inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = new emu_event_t(); return cudaSuccess; }
inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t = nullptr) { return cudaSuccess; }
inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return cudaSuccess; }
inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) { *ms = 0.f; return cudaSuccess; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline cudaError_t cudaPeekAtLastError() { return cudaSuccess; }
inline const char* cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "emulated CUDA error"; }
inline cudaError_t cudaSetDevice(int d) { return d == 0 ? cudaSuccess : cudaErrorInvalidValue; }
inline cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
inline cudaError_t cudaDeviceGetAttribute(int* v, cudaDeviceAttr a, int) {
    switch (a) {
        case cudaDevAttrMultiProcessorCount: *v = 148; break;
        case cudaDevAttrMaxSharedMemoryPerBlockOptin: *v = 232448; break;
        case cudaDevAttrComputeCapabilityMajor: *v = 10; break;
        case cudaDevAttrComputeCapabilityMinor: *v = 0; break;
        default: *v = 0;
    }
    return cudaSuccess;
}
template <class F> inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }
