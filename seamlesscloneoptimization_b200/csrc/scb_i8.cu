// scb_i8.cu -- kernels of the exact INT8 tensor-core DST engine (see scb_i8.h for the arithmetic).
//
// Replaces, along x, the reference's per-row transform  dft_kernel_0 -> cufftExecC2C -> dft_kernel_1
// (/root/reference/seamlessClone-CUDA/seamlessClone_imp.cpp:1694-1811) and its GEMM flavour (cublasSgemmBatched,
// imp.cpp:1266-1334), with the same mathematical operator and results that are exact up to 2^-31 of full scale.
#include <string>  // (before the platform header: the emulator build defines CUDA keywords as macros)

#include "scb_i8.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#ifndef SCB_EMU
#include <cuda.h>  // CUtensorMap (types only; the encoder is fetched with cudaGetDriverEntryPoint)
#endif

namespace scb {

// ---------------------------------------------------------------------------------------------
// digits
// ---------------------------------------------------------------------------------------------
// balanced base-256 digits, most significant first: v = sum_i d[i] 256^(ND-1-i), d[i] in [-128, 127] (the top digit takes what is left)
template <int ND>
SCB_D void i8_digits(int v, int (&d)[ND]) {
    SCB_UNROLL
    for (int i = ND - 1; i > 0; --i) {
        const int lo = ((v + 128) & 255) - 128;
        d[i] = lo;
        v = (v - lo) >> 8;  // exact: v - lo is a multiple of 256
    }
    d[0] = v;
}

// The four class sums of one output -> float.  W0 256^3 + W1 256^2 + W2 256 + W3 reaches 2^56: combined in FP32 with two
// roundings (the rounding of the high pair, the final one), identical on the device and in the emulator build.
SCB_D float i8_combine(int w0, int w1, int w2, int w3) {
    const float hi = fmaf((float)w0, 256.0f, (float)w1);
    const float lo = fmaf((float)w2, 256.0f, (float)w3);
    return fmaf(hi, 65536.0f, lo);
}
SCB_D double i8_combine_exact(int w0, int w1, int w2, int w3) {
    const long long t = (((long long)w0 * 256 + (long long)w1) * 256 + (long long)w2) * 256 + (long long)w3;
    return (double)t;
}

// ---------------------------------------------------------------------------------------------
// basis digit planes (plan time, cached per line length in the context)
//   row ((par * nsb + sb) * 4 + d) * 64 + r  holds digit d of  rint(2^30 sin(pi (j+1)(k0+1) / N)),  k0 = 2 (64 sb + r) + par,
//   for j = 0 .. kpar[par]-1; zero elsewhere.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) i8_basis_kernel(I8Geom g, signed char* basis) {
    const long long per_par = (long long)g.nsb * kI8P * (g.kpad / 4);  // (ki, j4) pairs; one thread = 4 consecutive j
    const long long total = 2 * per_par;
    const double invN = 1.0 / (double)(g.n + 1);
    const long long N2 = 2LL * (g.n + 1);
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int par = (int)(t / per_par);
        const long long rem = t - (long long)par * per_par;
        const int ki = (int)(rem / (g.kpad / 4)), j0 = 4 * (int)(rem - (long long)ki * (g.kpad / 4));
        unsigned word[kI8BasisDigits] = {0u, 0u, 0u, 0u};
        for (int e = 0; e < 4; ++e) {
            const int j = j0 + e;
            int d[kI8BasisDigits] = {0, 0, 0, 0};
            if (ki < g.nout[par] && j < g.kpar[par]) {
                const long long ph = ((long long)(j + 1) * (2 * ki + par + 1)) % N2;  // exact argument reduction
                const double s = sinpi((double)ph * invN);
                i8_digits<kI8BasisDigits>((int)llrint(s * (double)(1 << kI8BasisBits)), d);
            }
            for (int q = 0; q < kI8BasisDigits; ++q) word[q] |= (unsigned)(d[q] & 255) << (8 * e);
        }
        const int sb = ki / kI8P, r = ki % kI8P;
        for (int q = 0; q < kI8BasisDigits; ++q) {
            const size_t row = ((size_t)(par * g.nsb + sb) * kI8BasisDigits + q) * kI8P + r;
            *reinterpret_cast<unsigned*>(basis + row * g.kpad + j0) = word[q];
        }
    }
}

// ---------------------------------------------------------------------------------------------
// lines -> folded digit planes.  Block = one line, 256 threads; a thread converts 4 consecutive folded elements at a time
// (coalesced: a warp reads 512 contiguous bytes forwards and 512 backwards, and writes 128 contiguous bytes per plane).
//   v[j] = rint(x[j] * s),  f0 = v[j] + v[n-1-j],  f1 = v[j] - v[n-1-j]  (j < n/2;  the middle element of an odd line goes to f0)
//   a[(par * DA + i) * m_rows + line][j] = digit i of f_par[j]
// s = fixed_scale, or (per_line) 2^(29 - e) with 2^e > max |x| over the line, so |v| <= 2^29 and |f| <= 2^30.
// ---------------------------------------------------------------------------------------------
static constexpr int kI8DigThreads = 256;

template <int DA>
__global__ void __launch_bounds__(kI8DigThreads) i8_digitize_kernel(I8DigitizeParams p) {
    __shared__ float red[kI8DigThreads / 32];
    const int t = threadIdx.x;
    const int line = p.line0 + blockIdx.x;
    const int n = p.g.n, h = n >> 1;
    const bool real = line < p.lines;
    const float* x = nullptr;
    if (real) {
        const int r = line / 3, c = line - 3 * r;  // lines are channel-interleaved: line = 3 row + channel
        x = p.in + (size_t)c * p.in_plane + (size_t)r * p.in_pitch;
    }
    // The line is read ONCE: thread t keeps the elements j0 .. j0+3 and their mirror images for its first chunk (lines up to
    // 2 * 4 * 256 = 2048 points, i.e. every line of a 4K clone); longer lines re-read the later chunks.
    float fa[4] = {0.f, 0.f, 0.f, 0.f}, fb[4] = {0.f, 0.f, 0.f, 0.f};
    {
        const int j0 = 4 * t;
        if (real && j0 < p.g.kpar[0]) {
            SCB_UNROLL
            for (int e = 0; e < 4; ++e) {
                const int j = j0 + e;
                if (j < h) {
                    fa[e] = __ldg(x + j);
                    fb[e] = __ldg(x + (n - 1 - j));
                } else if (j == h && (n & 1)) {
                    fa[e] = __ldg(x + h);
                }
            }
        }
    }
    float s = p.fixed_scale;
    if (p.per_line) {
        float m = 0.f;
        SCB_UNROLL
        for (int e = 0; e < 4; ++e) m = fmaxf(m, fmaxf(fabsf(fa[e]), fabsf(fb[e])));
        if (real)
            for (int j = 4 * kI8DigThreads + t; j < n - 4 * kI8DigThreads; j += kI8DigThreads) m = fmaxf(m, fabsf(__ldg(x + j)));  // the part no first chunk holds
        for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        if ((t & 31) == 0) red[t >> 5] = m;
        __syncthreads();
        m = red[0];
        SCB_UNROLL
        for (int w = 1; w < kI8DigThreads / 32; ++w) m = fmaxf(m, red[w]);
        int e = 0;
        if (m > 0.f) (void)frexpf(m, &e);  // m = f 2^e, 0.5 <= f < 1
        if (e < -60) e = -60;
        if (e > 90) e = 90;
        s = ldexpf(1.0f, 29 - e);
        if (t == 0) p.lscale[line] = ldexpf(1.0f, e - 29);
    } else if (t == 0) {
        p.lscale[line] = 1.0f / p.fixed_scale;
    }
    for (int j0 = 4 * t; j0 < p.g.kpad; j0 += 4 * kI8DigThreads) {
        unsigned w[2][DA];
        int f0[4] = {0, 0, 0, 0}, f1[4] = {0, 0, 0, 0};
        if (real && j0 < p.g.kpar[0]) {
            const bool first = j0 == 4 * t;
            SCB_UNROLL
            for (int e = 0; e < 4; ++e) {
                const int j = j0 + e;
                if (j < h) {
                    const float xa = first ? fa[e] : __ldg(x + j), xb = first ? fb[e] : __ldg(x + (n - 1 - j));
                    const int va = __float2int_rn(xa * s), vb = __float2int_rn(xb * s);
                    f0[e] = va + vb;
                    f1[e] = va - vb;
                } else if (j == h && (n & 1)) {
                    f0[e] = __float2int_rn((first ? fa[e] : __ldg(x + h)) * s);
                }
            }
        }
        balanced_digits4<DA>(f0, w[0]);
        balanced_digits4<DA>(f1, w[1]);
        SCB_UNROLL
        for (int q = 0; q < 2; ++q)
            SCB_UNROLL
            for (int i = 0; i < DA; ++i)
                *reinterpret_cast<unsigned*>(p.a + ((size_t)(q * DA + i) * p.m_rows + line) * p.g.kpad + j0) = w[q][i];
    }
}

// ---------------------------------------------------------------------------------------------
// compose: planar float solved field -> interleaved u8 (OpenCV solve() epilogue + merge; reference post_processing,
// imp.cpp:2078-2103): v < 0 -> 0, v > 255 -> 255, else truncate toward zero.  A thread owns one 4-byte-aligned word group
// (12 bytes) of an output row: full words are stored as words whatever the pixel phase of the row, the two ragged ends bytewise.
// grid = (ceil((3 nx + 3) / 12 / 128), rows), block = 128
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) i8_compose_kernel(I8ComposeParams p) {
    const int y = p.y0 + blockIdx.y;
    unsigned char* row = p.out + (long long)y * p.out_pitch;
    const int off = (int)((uintptr_t)row & 3);           // the row's first byte sits `off` bytes into an aligned word
    const int nb = 3 * p.nx;                              // bytes of the row
    const int b0 = 12 * (blockIdx.x * 128 + threadIdx.x) - off;  // first byte (row-relative) of this thread's three aligned words
    if (b0 >= nb) return;
    const float* u = p.u + (size_t)y * p.pitch;
    unsigned wv[3] = {0u, 0u, 0u};
    SCB_UNROLL
    for (int k = 0; k < 12; ++k) {
        const int b = b0 + k;
        if (b >= 0 && b < nb) {
            const int x = b / 3, c = b - 3 * x;
            const float v = __ldg(u + (size_t)c * p.plane + x);
            const unsigned q = v < 0.f ? 0u : (v > 255.f ? 255u : (unsigned)__float2int_rz(v));
            wv[k >> 2] |= q << (8 * (k & 3));
        }
    }
    SCB_UNROLL
    for (int wi = 0; wi < 3; ++wi) {
        const int bw = b0 + 4 * wi;
        if (bw >= nb || bw + 4 <= 0) continue;
        if (bw >= 0 && bw + 4 <= nb) {
            *reinterpret_cast<unsigned*>(row + bw) = wv[wi];
        } else {
            SCB_UNROLL
            for (int k = 0; k < 4; ++k)
                if (bw + k >= 0 && bw + k < nb) row[bw + k] = (unsigned char)(wv[wi] >> (8 * k));
        }
    }
}

// v < 0 -> 0, v > 255 -> 255, else truncate toward zero (OpenCV's static_cast<uchar> after the clamp; reference post_processing, imp.cpp:2078-2103)
SCB_D unsigned char i8_to_u8(float v) { return v < 0.f ? (unsigned char)0 : (v > 255.f ? (unsigned char)255 : (unsigned char)__float2int_rz(v)); }

// which basis planes digit i of the lines multiplies: planes 0 .. cnt-1, landing in classes i .. i+cnt-1
SCB_HD int i8_plane_count(int i, int db) {
    const int left = kI8Classes - i;
    return left < db ? left : db;
}

// epilogue of one output element, shared by the tensor-core kernel and the emulator kernel
SCB_D void i8_store(const I8GemmParams& p, int line, int par, int ki, int w0, int w1, int w2, int w3, float ls) {
    if (line < p.line0 || line >= p.line1 || ki >= p.g.nout[par]) return;
    const int r = line / 3, c = line - 3 * r;  // lines are channel-interleaved: line = 3 row + channel
    const int k0 = 2 * ki + par;
    const float v = i8_combine(w0, w1, w2, w3) * (p.scale * ls);
    if (p.out_u8)
        p.out_u8[(long long)r * p.out_u8_pitch + 3 * k0 + c] = i8_to_u8(v);
    else
        p.out[(size_t)c * p.out_plane + (size_t)r * p.out_pitch + k0] = v;
    if (p.R && k0 < p.lowk) p.R[((size_t)c * p.lowk + k0) * p.lpc + r] = i8_combine_exact(w0, w1, w2, w3) * (p.rscale * (double)ls);
}

#ifdef SCB_EMU
// ---------------------------------------------------------------------------------------------
// CI stand-in for the tensor-core kernel: the same digit planes, the same class sums (exact integers, so the SAME
// numbers the GPU accumulates in TMEM), the same epilogue.  Checks the host logic, the tables and the digit/class
// bookkeeping in the GPU-less container; the tcgen05/TMA plumbing itself is checked on a B200 (tests -m gpu).
// ---------------------------------------------------------------------------------------------
template <int DA, int DB>
__global__ void i8_gemm_kernel(I8GemmParams p) {
    const int line = (p.mt0 + blockIdx.x) * kI8M + threadIdx.x;
    if (line >= p.lines) return;
    const int par = blockIdx.y / p.g.nsb, sb = blockIdx.y % p.g.nsb;
    const float ls = p.lscale[line];
    for (int r = 0; r < kI8P; ++r) {
        const int ki = sb * kI8P + r;
        if (ki >= p.g.nout[par]) break;
        long long W[kI8Classes] = {0, 0, 0, 0};
        for (int i = 0; i < DA; ++i) {
            const signed char* a = p.a + ((size_t)(par * DA + i) * p.m_rows + line) * p.g.kpad;
            const int cnt = i8_plane_count(i, DB);
            for (int d = 0; d < cnt; ++d) {
                const signed char* b = p.basis + (((size_t)(par * p.g.nsb + sb) * kI8BasisDigits + d) * kI8P + r) * p.g.kpad;
                long long acc = 0;
                for (int j = 0; j < p.g.kpar[par]; ++j) acc += (int)a[j] * (int)b[j];
                W[i + d] += acc;
            }
        }
        i8_store(p, line, par, ki, (int)W[0], (int)W[1], (int)W[2], (int)W[3], ls);
    }
}

template <int DA, int DB>
static int i8_launch_gemm_t(void* stream, const I8GemmParams& p) {
    (void)stream;
    if (p.mt1 <= p.mt0) return 0;
    SCB_LAUNCH((i8_gemm_kernel<DA, DB>), dim3(p.mt1 - p.mt0, 2 * p.g.nsb), dim3(kI8M), 0, stream, p);
    return 0;
}
int i8_configure() { return 0; }

#else  // ------------------------------ the real thing: sm_100a ------------------------------------

SCB_D unsigned i8_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
SCB_D void i8_mbar_init(unsigned bar, unsigned count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
SCB_D void i8_mbar_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
SCB_D void i8_mbar_wait(unsigned bar, unsigned parity) {
    unsigned done = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
    }
}
SCB_D void i8_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
SCB_D void i8_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
SCB_D void i8_tma_2d(unsigned dst, const CUtensorMap* map, int x, int y, unsigned bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst), "l"(map), "r"(x),
                 "r"(y), "r"(bar)
                 : "memory");
}
// K-major operand tile with KB-byte rows and a KB-byte swizzle (KB = 128 or 64): 8-row atoms of 8 KB bytes, SBO = 8 KB, LBO unused.
template <int KB>
SCB_D unsigned long long i8_desc(unsigned smem_addr) {
    unsigned long long d = 0;
    d |= (unsigned long long)((smem_addr >> 4) & 0x3fff);
    d |= (unsigned long long)((8 * KB) >> 4) << 32;
    d |= (unsigned long long)1 << 46;                      // descriptor version (sm_100)
    d |= (unsigned long long)(KB == 128 ? 2 : 4) << 61;    // SWIZZLE_128B / SWIZZLE_64B
    return d;
}
// kind::i8: s8 x s8 -> s32, A and B K-major, M = 128, N = n
SCB_D unsigned i8_idesc(int n) {
    unsigned d = 0;
    d |= 2u << 4;                      // c_format = S32
    d |= 1u << 7;                      // a_format = signed 8 bit
    d |= 1u << 10;                     // b_format = signed 8 bit
    d |= (unsigned)(n >> 3) << 17;     // n_dim
    d |= (unsigned)(kI8M >> 4) << 24;  // m_dim
    return d;
}
SCB_D void i8_mma(unsigned tmem_d, unsigned long long adesc, unsigned long long bdesc, unsigned idesc, unsigned accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// the same load delivered to the same CTA-relative offset (and mbarrier) of every CTA in `mask` of the cluster
SCB_D void i8_tma_2d_mc(unsigned dst, const CUtensorMap* map, int x, int y, unsigned bar, unsigned short mask) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%2, %3}], [%4], %5;" ::"r"(dst),
                 "l"(map), "r"(x), "r"(y), "r"(bar), "h"(mask)
                 : "memory");
}
SCB_D void i8_commit_mc(unsigned bar, unsigned short mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(mask) : "memory");
}
SCB_D void i8_cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
SCB_D unsigned i8_cluster_ctarank() {
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
SCB_D void i8_commit(unsigned bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
SCB_D void i8_tmem_ld16(unsigned taddr, int (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
          "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
SCB_D void i8_tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

static constexpr int kI8EpiWarps = 8;                        // two per TMEM lane quarter, each half of the CTA's output columns
static constexpr int kI8Threads = 32 * (2 + kI8EpiWarps);    // warp 0: TMA producer + TMEM allocator, warp 1: MMA issuer, warps 2-9: epilogue

// NSUB = 64-output sub-blocks per CTA tile, KB = K bytes per pipeline stage row (= the swizzle width).
//   NSUB = 2, KB = 128: 128 x 128 outputs per CTA, all 512 TMEM columns, 2 stages of 96-112 KB: ONE CTA per SM -- its prologue, first
//                       stage and epilogue (as long as the main loop, measured) are exposed.
//   NSUB = 1, KB = 64 : 128 x 64 outputs, 256 TMEM columns, 2-3 stages of 32-44 KB: TWO CTAs per SM, so one CTA's epilogue runs under
//                       the other's main loop; the line digits are re-read per 64 instead of 128 outputs (more L2 traffic).
template <int DA, int DB, int NSUB, int KB>
struct I8Cfg {
    static constexpr size_t A_DIGIT = (size_t)kI8M * KB;           // one digit tile of the lines
    static constexpr size_t A_BYTES = (size_t)DA * A_DIGIT;
    static constexpr size_t B_PLANE = (size_t)kI8P * KB;           // one digit plane of the basis
    static constexpr size_t B_SUB = (size_t)DB * B_PLANE;
    static constexpr size_t STAGE = A_BYTES + NSUB * B_SUB;
    static constexpr int CTAS_PER_SM = NSUB == 1 ? 2 : 1;
    static constexpr size_t BUDGET = (size_t)(CTAS_PER_SM == 2 ? 113 : 226) * 1024 - 2048;
    static constexpr int STAGES = (int)(BUDGET / STAGE) > 4 ? 4 : (int)(BUDGET / STAGE);
    static constexpr int TMEM_COLS = NSUB * kI8Classes * kI8P;      // 256 or 512
    static constexpr int OUT_PITCH = NSUB * kI8P + 1;               // floats per row of the epilogue transpose buffer (reuses the stages)
    static constexpr size_t OUT_STAGE = (size_t)kI8M * OUT_PITCH * sizeof(float);
    static constexpr size_t SMEM = STAGES * STAGE + 1024 /* alignment slack */ + 256 /* barriers */;
    static_assert(STAGES >= 2, "at least two pipeline stages");
    static_assert(OUT_STAGE <= STAGES * STAGE, "epilogue buffer must fit in the pipeline stages");
    static_assert(SMEM <= 232448, "shared memory budget of one CTA");
};

// grid = (line tiles, 2 * nsb / NSUB): blockIdx.y / (nsb / NSUB) = parity, the rest = first sub-block / NSUB
// CL = 2: clusters of two CTAs with consecutive line tiles share their basis planes -- each CTA loads HALF the rows of every sub-block
// and multicasts them into both shared memories, which halves the L2 traffic of the larger operand.  A stage is released to the
// producers only when BOTH CTAs have consumed it (tcgen05.commit multicast to both empty barriers).
template <int DA, int DB, int NSUB, int KB, int CL>
__global__ void __launch_bounds__(kI8Threads, (NSUB == 1 ? 2 : 1))
i8_gemm_kernel(const __grid_constant__ CUtensorMap amap, const __grid_constant__ CUtensorMap bmap, I8GemmParams p) {
    static_assert(CL == 1 || CL == 2, "cluster of one or two CTAs");
    constexpr int kI8TmemCols = I8Cfg<DA, DB, NSUB, KB>::TMEM_COLS;
    using Cfg = I8Cfg<DA, DB, NSUB, KB>;
    constexpr int kI8Stages = Cfg::STAGES;
    extern __shared__ unsigned char i8_smem_raw[];
    unsigned char* base = reinterpret_cast<unsigned char*>(((size_t)i8_smem_raw + 1023) & ~(size_t)1023);
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(base + (size_t)kI8Stages * Cfg::STAGE);
    const unsigned bar0 = i8_smem_u32(bars);
    auto full = [&](int s) { return bar0 + 8u * s; };
    auto empty = [&](int s) { return bar0 + 8u * (kI8Stages + s); };
    const unsigned acc_bar = bar0 + 8u * (2 * kI8Stages);
    unsigned* tmem_slot = reinterpret_cast<unsigned*>(bars + 2 * kI8Stages + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    long long* trace = p.trace ? p.trace + 8 * ((size_t)blockIdx.y * gridDim.x + blockIdx.x) : nullptr;
    auto stamp = [&](int i) {
        if (trace) trace[i] = clock64();
    };
    if (threadIdx.x == 0) stamp(0);
    const int groups = p.g.nsb / NSUB;
    const int par = blockIdx.y / groups, sb0 = (blockIdx.y % groups) * NSUB;
    const int kpar = p.g.kpar[par];
    const int num_kb = (kpar + KB - 1) / KB;
    const int m0 = (p.mt0 + (int)blockIdx.x) * kI8M;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&amap) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&bmap) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < kI8Stages; ++s) {
            i8_mbar_init(full(s), 1);    // the TMA thread's arrive.expect_tx
            i8_mbar_init(empty(s), CL);  // tcgen05.commit of every CTA of the cluster
        }
        i8_mbar_init(acc_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {  // TMEM: all 512 columns (NSUB x 4 classes x 64 outputs of int32 accumulators)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(i8_smem_u32(tmem_slot)), "r"(kI8TmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    i8_fence_before();
    __syncthreads();
    if (CL > 1) i8_cluster_sync();  // the peer's barriers exist before anything is multicast to them
    i8_fence_after();
    const unsigned tmem_base = *tmem_slot;
    const unsigned cta_rank = CL > 1 ? i8_cluster_ctarank() : 0u;
    const unsigned short cl_mask = (unsigned short)((1u << CL) - 1u);
    if (threadIdx.x == 0) stamp(1);

    if (warp == 0) {
        // ===== TMA producer: DA digit tiles of the lines + NSUB x DB digit planes of the basis per k-block =====
        if (lane == 0) {
            const unsigned bytes = (unsigned)Cfg::STAGE;
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % kI8Stages;
                i8_mbar_wait(empty(s), ((kb / kI8Stages) & 1) ^ 1);
                i8_mbar_expect_tx(full(s), bytes);
                unsigned char* st = base + (size_t)s * Cfg::STAGE;
                SCB_UNROLL
                for (int i = 0; i < DA; ++i) i8_tma_2d(i8_smem_u32(st + (size_t)i * Cfg::A_DIGIT), &amap, kb * KB, (par * DA + i) * p.m_rows + m0, full(s));
                SCB_UNROLL
                for (int u = 0; u < NSUB; ++u) {
                    const unsigned dstb = i8_smem_u32(st + Cfg::A_BYTES + (size_t)u * Cfg::B_SUB);
                    const int rowb = ((par * p.g.nsb + sb0 + u) * kI8BasisDigits) * kI8P;
                    if (CL == 1) {
                        i8_tma_2d(dstb, &bmap, kb * KB, rowb, full(s));
                    } else {  // this CTA's half of the rows, delivered to both CTAs; the other half arrives from the peer
                        constexpr int HR = DB * kI8P / 2;
                        i8_tma_2d_mc(dstb + cta_rank * (unsigned)(HR * KB), &bmap, kb * KB, rowb + (int)cta_rank * HR, full(s), cl_mask);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: per 32-element k-step and sub-block, one MMA per digit of the lines against its planes =====
        if (lane == 0) {
            unsigned idesc[kI8Classes + 1];
            SCB_UNROLL
            for (int c = 1; c <= kI8Classes; ++c) idesc[c] = i8_idesc(kI8P * c);
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % kI8Stages;
                i8_mbar_wait(full(s), (kb / kI8Stages) & 1);
                i8_fence_after();
                if (kb == 0) stamp(2);
                const unsigned st = i8_smem_u32(base + (size_t)s * Cfg::STAGE);
                const int left = kpar - kb * KB;
                const int nks = left >= KB ? KB / 32 : (left + 31) / 32;
                for (int ks = 0; ks < nks; ++ks) {
                    const unsigned long long o = (unsigned long long)(ks * 2);  // 32 bytes of K = 2 x 16-byte units of the start address
                    const bool first = (kb | ks) == 0;
                    SCB_UNROLL
                    for (int u = 0; u < NSUB; ++u) {
                        const unsigned tm = tmem_base + (unsigned)(u * kI8Classes * kI8P);
                        const unsigned long long bd = i8_desc<KB>(st + (unsigned)(Cfg::A_BYTES + (size_t)u * Cfg::B_SUB)) + o;
                        auto digit = [&](int i, unsigned acc) {
                            const int cnt = i8_plane_count(i, DB);
                            i8_mma(tm + (unsigned)(i * kI8P), i8_desc<KB>(st + (unsigned)((size_t)i * Cfg::A_DIGIT)) + o, bd, idesc[cnt], acc);
                        };
                        // Every class must be zero-initialised by the first MMA that touches it.  With DB >= 4 digit 0 covers all four
                        // classes; with DB = 3 it leaves class 3, whose first writer that covers nothing else is the last digit.
                        if (DB < kI8Classes && first) digit(DA - 1, 0u);
                        digit(0, first ? 0u : 1u);
                        SCB_UNROLL
                        for (int i = 1; i < DA; ++i)
                            if (!(DB < kI8Classes && first && i == DA - 1)) digit(i, 1u);
                    }
                }
                if (CL == 1)
                    i8_commit(empty(s));  // frees the stage when these MMAs have read it
                else
                    i8_commit_mc(empty(s), cl_mask);
            }
            i8_commit(acc_bar);
            stamp(3);
        }
    } else {
        // ===== epilogue: 8 warps; warp w may read TMEM lanes 32 (w % 4) .. +31 (= lines), the two warps of a lane quarter take half of
        // the CTA's output columns each.  class sums -> float -> shared-memory transpose -> rows of the output =====
        i8_mbar_wait(acc_bar, 0);
        i8_fence_after();
        if (threadIdx.x == 64) stamp(4);
        constexpr int OP = Cfg::OUT_PITCH;
        constexpr int COLS = NSUB * kI8P / 2;        // output columns of this warp: 64 (NSUB = 2) or 32 (NSUB = 1)
        const int q = warp & 3, half = (warp - 2) >> 2;
        const int col0 = half * COLS;
        const int line = m0 + 32 * q + lane;
        const float ls = line < p.m_rows ? __ldg(p.lscale + line) : 0.f;
        float* tile = reinterpret_cast<float*>(base) + (size_t)(32 * q) * OP;  // this quarter's 32 rows of the transpose buffer
        const float sc = p.scale * ls;
        const int nout = p.g.nout[par];
        SCB_UNROLL
        for (int cc = 0; cc < COLS; cc += 16) {
            const int col = col0 + cc;               // CTA-relative output column
            const int u = col / kI8P, within = col % kI8P;
            if ((sb0 + u) * kI8P + within >= nout) break;
            const unsigned tm = tmem_base + ((unsigned)(32 * q) << 16) + (unsigned)(u * kI8Classes * kI8P + within);
            int w0[16], w1[16], w2[16], w3[16];
            i8_tmem_ld16(tm + (unsigned)(0 * kI8P), w0);
            i8_tmem_ld16(tm + (unsigned)(1 * kI8P), w1);
            i8_tmem_ld16(tm + (unsigned)(2 * kI8P), w2);
            i8_tmem_ld16(tm + (unsigned)(3 * kI8P), w3);
            i8_tmem_wait_ld();
            SCB_UNROLL
            for (int i = 0; i < 16; ++i) tile[(size_t)lane * OP + col + i] = i8_combine(w0[i], w1[i], w2[i], w3[i]) * sc;
            if (p.R && col == 0 && sb0 == 0 && line >= p.line0 && line < p.line1) {  // exact float64 row sums of the lowest frequencies
                const int r = line / 3, c = line - 3 * r;  // lines are channel-interleaved: line = 3 row + channel
                SCB_UNROLL
                for (int i = 0; i < (kI8LowK + 1) / 2; ++i) {
                    const int k0 = 2 * i + par;
                    if (k0 < p.lowk && i < nout) p.R[((size_t)c * p.lowk + k0) * p.lpc + r] = i8_combine_exact(w0[i], w1[i], w2[i], w3[i]) * (p.rscale * (double)ls);
                }
            }
        }
        __syncwarp();
        // rows of the tile -> rows of the output: a lane writes one output column of a line per step (stride-2 floats: the other
        // parity's CTA fills the gaps)
        {
            int ln = m0 + 32 * q;
            int rr = ln / 3, c = ln - 3 * rr;
            for (int r = 0; r < 32 && ln < p.line1; ++r, ++ln) {
                float* o = p.out + (size_t)c * p.out_plane + (size_t)rr * p.out_pitch;
                const float* trow = tile + (size_t)r * OP + col0;
                SCB_UNROLL
                for (int j = lane; j < COLS; j += 32) {
                    const int ki = sb0 * kI8P + col0 + j;
                    if (ki < nout && ln >= p.line0) {
                        if (p.out_u8)
                            p.out_u8[(long long)rr * p.out_u8_pitch + 3 * (2 * ki + par) + c] = i8_to_u8(trow[j]);
                        else
                            o[2 * ki + par] = trow[j];
                    }
                }
                if (++c == 3) {
                    c = 0;
                    ++rr;
                }
            }
        }
        i8_fence_before();
        if (threadIdx.x == 64) stamp(5);
    }
    __syncthreads();
    if (threadIdx.x == 0) stamp(6);
    if (CL > 1) i8_cluster_sync();  // no CTA leaves while its peer may still signal its barriers
    if (warp == 0) {
        i8_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kI8TmemCols) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------
// The persistent variant (the default): one CTA per SM walks a static list of 128-line x 64-output tiles.  The TMEM holds TWO
// accumulator sets of 256 columns, so the MMA issuer starts tile t+1 while the epilogue warps drain tile t, and the TMA producer
// keeps the stage ring full across tile boundaries -- the prologue, first-stage latency and epilogue that the one-tile-per-CTA
// kernel above exposes (as long as its main loop: profiles/r2_i8_variants.txt) all run under the MMAs of another tile.
//   tiles : t -> (group g = t / mt, line tile m = t % mt);  g < nsbr[0]: parity 0, sub-block g;  else parity 1, sub-block g - nsbr[0]
//           (only the sub-blocks that hold outputs: nsbr[p] = ceil(nout[p] / 64)); CTA b takes t = b, b + grid, ...
//           CL = 2: a cluster takes PAIRS of line tiles (2 j, 2 j + 1) of one group and shares the basis planes by multicast.
// ---------------------------------------------------------------------------------------------
template <int DA, int DB, int KB>
struct I8PCfg {
    static constexpr size_t A_DIGIT = (size_t)kI8M * KB;
    static constexpr size_t A_BYTES = (size_t)DA * A_DIGIT;
    static constexpr size_t B_PLANE = (size_t)kI8P * KB;
    static constexpr size_t B_SUB = (size_t)DB * B_PLANE;
    static constexpr size_t STAGE = A_BYTES + B_SUB;
    static constexpr int OUT_PITCH = kI8P + 1;
    static constexpr size_t OUT_BYTES = ((size_t)kI8M * OUT_PITCH * sizeof(float) + 1023) / 1024 * 1024;
    static constexpr size_t BUDGET = (size_t)226 * 1024 - OUT_BYTES - 2048;
    static constexpr int STAGES = (int)(BUDGET / STAGE) > 8 ? 8 : (int)(BUDGET / STAGE);
    static constexpr size_t SMEM = STAGES * STAGE + OUT_BYTES + 1024 /* alignment slack */ + 256 /* barriers */;
    static_assert(STAGES >= 2, "at least two pipeline stages");
    static_assert(SMEM <= 232448, "shared memory budget of one CTA");
};

template <int DA, int DB, int KB, int CL>
__global__ void __launch_bounds__(kI8Threads, 1)
i8_gemm_pkernel(const __grid_constant__ CUtensorMap amap, const __grid_constant__ CUtensorMap bmap, I8GemmParams p) {
    using Cfg = I8PCfg<DA, DB, KB>;
    constexpr int S = Cfg::STAGES;
    extern __shared__ unsigned char i8_smem_raw[];
    unsigned char* base = reinterpret_cast<unsigned char*>(((size_t)i8_smem_raw + 1023) & ~(size_t)1023);
    float* outbuf = reinterpret_cast<float*>(base + (size_t)S * Cfg::STAGE);
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(base + (size_t)S * Cfg::STAGE + Cfg::OUT_BYTES);
    const unsigned bar0 = i8_smem_u32(bars);
    auto full = [&](int s) { return bar0 + 8u * s; };
    auto empty = [&](int s) { return bar0 + 8u * (S + s); };
    auto acc_full = [&](int b) { return bar0 + 8u * (2 * S + b); };
    auto acc_empty = [&](int b) { return bar0 + 8u * (2 * S + 2 + b); };
    unsigned* tmem_slot = reinterpret_cast<unsigned*>(bars + 2 * S + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int mt = p.mt1 - p.mt0;                               // line tiles of this launch
    const int mtu = (mt + CL - 1) / CL;                         // scheduling units per group (pairs of line tiles when CL = 2)
    const int nsbr0 = (p.g.nout[0] + kI8P - 1) / kI8P, nsbr1 = (p.g.nout[1] + kI8P - 1) / kI8P;
    const int units = (nsbr0 + nsbr1) * mtu;
    const unsigned cta_rank = CL > 1 ? i8_cluster_ctarank() : 0u;
    const int first_unit = (int)(blockIdx.x / CL), unit_step = (int)(gridDim.x / CL);
    const unsigned short cl_mask = (unsigned short)((1u << CL) - 1u);

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&amap) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&bmap) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < S; ++s) {
            i8_mbar_init(full(s), 1);
            i8_mbar_init(empty(s), CL);
        }
        for (int b = 0; b < 2; ++b) {
            i8_mbar_init(acc_full(b), 1);             // tcgen05.commit of the tile's last MMA
            i8_mbar_init(acc_empty(b), kI8EpiWarps);  // one arrival per epilogue warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(i8_smem_u32(tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    i8_fence_before();
    __syncthreads();
    if (CL > 1) i8_cluster_sync();
    i8_fence_after();
    const unsigned tmem_base = *tmem_slot;

    auto decode = [&](int unit, int& par, int& sb, int& m0) {
        const int g = unit / mtu, mu = unit - g * mtu;
        par = g >= nsbr0 ? 1 : 0;
        sb = par ? g - nsbr0 : g;
        m0 = (p.mt0 + mu * CL + (int)cta_rank) * kI8M;
    };

    if (warp == 0) {
        // ===== TMA producer: runs ahead over tile boundaries, bounded by the stage ring =====
        if (lane == 0) {
            unsigned it = 0;
            for (int unit = first_unit; unit < units; unit += unit_step) {
                int par, sb, m0;
                decode(unit, par, sb, m0);
                const int num_kb = (p.g.kpar[par] + KB - 1) / KB;
                const int rowb = ((par * p.g.nsb + sb) * kI8BasisDigits) * kI8P;
                for (int kb = 0; kb < num_kb; ++kb, ++it) {
                    const int s = (int)(it % S);
                    i8_mbar_wait(empty(s), ((it / S) & 1) ^ 1);
                    i8_mbar_expect_tx(full(s), (unsigned)Cfg::STAGE);
                    unsigned char* st = base + (size_t)s * Cfg::STAGE;
                    SCB_UNROLL
                    for (int i = 0; i < DA; ++i) i8_tma_2d(i8_smem_u32(st + (size_t)i * Cfg::A_DIGIT), &amap, kb * KB, (par * DA + i) * p.m_rows + m0, full(s));
                    const unsigned dstb = i8_smem_u32(st + Cfg::A_BYTES);
                    if (CL == 1) {
                        i8_tma_2d(dstb, &bmap, kb * KB, rowb, full(s));
                    } else {  // this CTA's 1/CL of the rows, delivered to every CTA of the cluster
                        constexpr int HR = DB * kI8P / CL;
                        static_assert(HR % 8 == 0, "whole swizzle atoms per CTA");
                        i8_tma_2d_mc(dstb + cta_rank * (unsigned)(HR * KB), &bmap, kb * KB, rowb + (int)cta_rank * HR, full(s), cl_mask);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            unsigned idesc[kI8Classes + 1];
            SCB_UNROLL
            for (int c = 1; c <= kI8Classes; ++c) idesc[c] = i8_idesc(kI8P * c);
            unsigned it = 0;
            int lt = 0;
            for (int unit = first_unit; unit < units; unit += unit_step, ++lt) {
                int par, sb, m0;
                decode(unit, par, sb, m0);
                const int kpar = p.g.kpar[par];
                const int num_kb = (kpar + KB - 1) / KB;
                const int b = lt & 1;
                i8_mbar_wait(acc_empty(b), ((lt >> 1) & 1) ^ 1);  // the epilogue has drained this accumulator set (passes at once the first two times)
                i8_fence_after();
                const unsigned tm = tmem_base + (unsigned)(b * kI8Classes * kI8P);
                for (int kb = 0; kb < num_kb; ++kb, ++it) {
                    const int s = (int)(it % S);
                    i8_mbar_wait(full(s), (it / S) & 1);
                    i8_fence_after();
                    const unsigned st = i8_smem_u32(base + (size_t)s * Cfg::STAGE);
                    const int left = kpar - kb * KB;
                    const int nks = left >= KB ? KB / 32 : (left + 31) / 32;
                    for (int ks = 0; ks < nks; ++ks) {
                        const unsigned long long o = (unsigned long long)(ks * 2);
                        const bool first = (kb | ks) == 0;
                        const unsigned long long bd = i8_desc<KB>(st + (unsigned)Cfg::A_BYTES) + o;
                        auto digit = [&](int i, unsigned acc) {
                            const int cnt = i8_plane_count(i, DB);
                            i8_mma(tm + (unsigned)(i * kI8P), i8_desc<KB>(st + (unsigned)((size_t)i * Cfg::A_DIGIT)) + o, bd, idesc[cnt], acc);
                        };
                        if (DB < kI8Classes && first) digit(DA - 1, 0u);
                        digit(0, first ? 0u : 1u);
                        SCB_UNROLL
                        for (int i = 1; i < DA; ++i)
                            if (!(DB < kI8Classes && first && i == DA - 1)) digit(i, 1u);
                    }
                    if (CL == 1)
                        i8_commit(empty(s));
                    else
                        i8_commit_mc(empty(s), cl_mask);
                }
                i8_commit(acc_full(b));
            }
        }
    } else {
        // ===== epilogue: 8 warps, 32 lines x 32 output columns each =====
        constexpr int OP = Cfg::OUT_PITCH;
        constexpr int COLS = kI8P / 2;
        const int q = warp & 3, half = (warp - 2) >> 2;
        const int col0 = half * COLS;
        float* tile = outbuf + (size_t)(32 * q) * OP;
        int lt = 0;
        for (int unit = first_unit; unit < units; unit += unit_step, ++lt) {
            int par, sb, m0;
            decode(unit, par, sb, m0);
            const int b = lt & 1;
            const int nout = p.g.nout[par];
            const int line = m0 + 32 * q + lane;
            const float ls = line < p.m_rows ? __ldg(p.lscale + line) : 0.f;
            const float sc = p.scale * ls;
            i8_mbar_wait(acc_full(b), (lt >> 1) & 1);
            i8_fence_after();
            const unsigned tmq = tmem_base + ((unsigned)(32 * q) << 16) + (unsigned)(b * kI8Classes * kI8P);
            SCB_UNROLL
            for (int cc = 0; cc < COLS; cc += 16) {
                const int col = col0 + cc;
                int w0[16], w1[16], w2[16], w3[16];
                i8_tmem_ld16(tmq + (unsigned)(0 * kI8P + col), w0);
                i8_tmem_ld16(tmq + (unsigned)(1 * kI8P + col), w1);
                i8_tmem_ld16(tmq + (unsigned)(2 * kI8P + col), w2);
                i8_tmem_ld16(tmq + (unsigned)(3 * kI8P + col), w3);
                i8_tmem_wait_ld();
                SCB_UNROLL
                for (int i = 0; i < 16; ++i) tile[(size_t)lane * OP + col + i] = i8_combine(w0[i], w1[i], w2[i], w3[i]) * sc;
                if (p.R && col == 0 && sb == 0 && line >= p.line0 && line < p.line1) {
                    const int r = line / 3, c = line - 3 * r;  // lines are channel-interleaved: line = 3 row + channel
                    SCB_UNROLL
                    for (int i = 0; i < (kI8LowK + 1) / 2; ++i) {
                        const int k0 = 2 * i + par;
                        if (k0 < p.lowk && i < nout) p.R[((size_t)c * p.lowk + k0) * p.lpc + r] = i8_combine_exact(w0[i], w1[i], w2[i], w3[i]) * (p.rscale * (double)ls);
                    }
                }
            }
            // the accumulator set is in registers / shared memory now: hand it back to the MMA issuer before the global stores
            i8_fence_before();
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(acc_empty(b)) : "memory");
            {
                int ln = m0 + 32 * q;
                int rr = ln / 3, c = ln - 3 * rr;
                const int ki = sb * kI8P + col0 + lane;
                for (int r = 0; r < 32 && ln < p.line1; ++r, ++ln) {
                    if (ki < nout && ln >= p.line0) {
                        const float v = tile[(size_t)r * OP + col0 + lane];
                        if (p.out_u8)
                            p.out_u8[(long long)rr * p.out_u8_pitch + 3 * (2 * ki + par) + c] = i8_to_u8(v);
                        else
                            p.out[(size_t)c * p.out_plane + (size_t)rr * p.out_pitch + 2 * ki + par] = v;
                    }
                    if (++c == 3) {
                        c = 0;
                        ++rr;
                    }
                }
            }
            __syncwarp();  // the tile rows are free for the next tile
        }
    }
    __syncthreads();
    if (CL > 1) i8_cluster_sync();
    if (warp == 0) {
        i8_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------
// The persistent kernel on 128-line x 128-output tiles (SCB_I8_PERSISTENT=2).  What bounds i8_gemm_pkernel is the operand stream
// into shared memory (~33 B/clk/SM = 9 TB/s chip-wide out of L2, profiles/r2_i8_variants.txt), not the tensor pipe: per 64 outputs it
// re-reads the digit tiles of its 128 lines.  Two 64-output sub-blocks per tile share those tiles: 0.75 x (forward) / 0.64 x (inverse)
// of the bytes per MAC.  The price: the accumulators of one tile (2 sub-blocks x 4 classes x 64 columns) fill all 512 TMEM columns,
// so there is no second set -- the epilogue of tile t is NOT hidden under the MMAs of tile t+1 (the TMA producer still runs ahead
// through the stage ring, and the other SMs' main loops keep the L2 busy meanwhile).
//   tiles : t -> (group g = t / mt, line tile m = t % mt);  group = (parity, PAIR of sub-blocks); the second sub-block of the last
//           pair of a parity may hold no outputs: its MMAs and its epilogue are skipped.
//   epilogue warp (q, u): TMEM lane quarter q, sub-block u -- 32 lines x 64 outputs: all 64 class-sum quadruples -> 64 floats in
//           registers, hand the accumulators back, then two 32 x 32 transposes through the warp's shared-memory tile and the stores.
// ---------------------------------------------------------------------------------------------
template <int DA, int DB, int KB>
struct I8P2Cfg {
    static constexpr size_t A_DIGIT = (size_t)kI8M * KB;
    static constexpr size_t A_BYTES = (size_t)DA * A_DIGIT;
    static constexpr size_t B_PLANE = (size_t)kI8P * KB;
    static constexpr size_t B_SUB = (size_t)DB * B_PLANE;
    static constexpr size_t STAGE = A_BYTES + 2 * B_SUB;
    static constexpr int OUT_PITCH = 33;                                                        // floats per row of a warp's 32 x 32 transpose tile
    static constexpr size_t OUT_BYTES = (size_t)kI8EpiWarps * 32 * OUT_PITCH * sizeof(float);   // 33 KB
    static constexpr size_t BUDGET = (size_t)232448 - OUT_BYTES - 1024 - 256;
    static constexpr int STAGES = (int)(BUDGET / STAGE) > 6 ? 6 : (int)(BUDGET / STAGE);
    static constexpr size_t SMEM = STAGES * STAGE + OUT_BYTES + 1024 /* alignment slack */ + 256 /* barriers */;
    static_assert(STAGE % 1024 == 0 && OUT_BYTES % 1024 == 0, "stages stay 1024-byte aligned");
    static_assert(STAGES >= 2, "at least two pipeline stages");
    static_assert(SMEM <= 232448, "shared memory budget of one CTA");
};

template <int DA, int DB, int KB>
__global__ void __launch_bounds__(kI8Threads, 1)
i8_gemm_p2kernel(const __grid_constant__ CUtensorMap amap, const __grid_constant__ CUtensorMap bmap, I8GemmParams p) {
    using Cfg = I8P2Cfg<DA, DB, KB>;
    constexpr int S = Cfg::STAGES;
    constexpr unsigned kSubCols = (unsigned)(kI8Classes * kI8P);  // TMEM columns of one sub-block's accumulators
    extern __shared__ unsigned char i8_smem_raw[];
    unsigned char* base = reinterpret_cast<unsigned char*>(((size_t)i8_smem_raw + 1023) & ~(size_t)1023);
    float* outbuf = reinterpret_cast<float*>(base + (size_t)S * Cfg::STAGE);
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(base + (size_t)S * Cfg::STAGE + Cfg::OUT_BYTES);
    const unsigned bar0 = i8_smem_u32(bars);
    auto full = [&](int s) { return bar0 + 8u * s; };
    auto empty = [&](int s) { return bar0 + 8u * (S + s); };
    const unsigned acc_full = bar0 + 8u * (2 * S), acc_empty = bar0 + 8u * (2 * S + 1);
    unsigned* tmem_slot = reinterpret_cast<unsigned*>(bars + 2 * S + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int mt = p.mt1 - p.mt0;
    const int nsbr0 = (p.g.nout[0] + kI8P - 1) / kI8P, nsbr1 = (p.g.nout[1] + kI8P - 1) / kI8P;
    const int np0 = (nsbr0 + 1) / 2, np1 = (nsbr1 + 1) / 2;  // pairs of sub-blocks per parity (the table holds an even number of sub-blocks)
    const int units = (np0 + np1) * mt;
    const int first_unit = (int)blockIdx.x, unit_step = (int)gridDim.x;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&amap) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&bmap) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < S; ++s) {
            i8_mbar_init(full(s), 1);
            i8_mbar_init(empty(s), 1);
        }
        i8_mbar_init(acc_full, 1);              // tcgen05.commit of the tile's last MMA
        i8_mbar_init(acc_empty, kI8EpiWarps);   // one arrival per epilogue warp
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(i8_smem_u32(tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    i8_fence_before();
    __syncthreads();
    i8_fence_after();
    const unsigned tmem_base = *tmem_slot;

    // unit -> parity, first sub-block of the pair, how many of the pair's sub-blocks hold outputs, first line of the tile
    auto decode = [&](int unit, int& par, int& sb0, int& nu, int& m0) {
        const int g = unit / mt, mu = unit - g * mt;
        par = g >= np0 ? 1 : 0;
        sb0 = 2 * (par ? g - np0 : g);
        nu = (sb0 + 1 < (par ? nsbr1 : nsbr0)) ? 2 : 1;
        m0 = (p.mt0 + mu) * kI8M;
    };

    if (warp == 0) {
        // ===== TMA producer: DA digit tiles of the lines + the DB planes of both sub-blocks per k-block; runs ahead over tile boundaries =====
        if (lane == 0) {
            unsigned it = 0;
            for (int unit = first_unit; unit < units; unit += unit_step) {
                int par, sb0, nu, m0;
                decode(unit, par, sb0, nu, m0);
                const int num_kb = (p.g.kpar[par] + KB - 1) / KB;
                const int rowb = ((par * p.g.nsb + sb0) * kI8BasisDigits) * kI8P;
                for (int kb = 0; kb < num_kb; ++kb, ++it) {
                    const int s = (int)(it % S);
                    i8_mbar_wait(empty(s), ((it / S) & 1) ^ 1);
                    i8_mbar_expect_tx(full(s), (unsigned)Cfg::STAGE);  // the empty second sub-block of a last pair is loaded too (zero planes): one byte count
                    unsigned char* st = base + (size_t)s * Cfg::STAGE;
                    SCB_UNROLL
                    for (int i = 0; i < DA; ++i) i8_tma_2d(i8_smem_u32(st + (size_t)i * Cfg::A_DIGIT), &amap, kb * KB, (par * DA + i) * p.m_rows + m0, full(s));
                    SCB_UNROLL
                    for (int u = 0; u < 2; ++u)
                        i8_tma_2d(i8_smem_u32(st + Cfg::A_BYTES + (size_t)u * Cfg::B_SUB), &bmap, kb * KB, rowb + u * kI8BasisDigits * kI8P, full(s));
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: per 32-element k-step and sub-block, one MMA per digit of the lines against its planes =====
        if (lane == 0) {
            unsigned idesc[kI8Classes + 1];
            SCB_UNROLL
            for (int c = 1; c <= kI8Classes; ++c) idesc[c] = i8_idesc(kI8P * c);
            unsigned it = 0;
            int lt = 0;
            for (int unit = first_unit; unit < units; unit += unit_step, ++lt) {
                int par, sb0, nu, m0;
                decode(unit, par, sb0, nu, m0);
                const int kpar = p.g.kpar[par];
                const int num_kb = (kpar + KB - 1) / KB;
                i8_mbar_wait(acc_empty, (lt & 1) ^ 1);  // the epilogue has drained the accumulators of the previous tile (passes at once the first time)
                i8_fence_after();
                for (int kb = 0; kb < num_kb; ++kb, ++it) {
                    const int s = (int)(it % S);
                    i8_mbar_wait(full(s), (it / S) & 1);
                    i8_fence_after();
                    const unsigned st = i8_smem_u32(base + (size_t)s * Cfg::STAGE);
                    const int left = kpar - kb * KB;
                    const int nks = left >= KB ? KB / 32 : (left + 31) / 32;
                    for (int ks = 0; ks < nks; ++ks) {
                        const unsigned long long o = (unsigned long long)(ks * 2);  // 32 bytes of K = 2 x 16-byte units of the start address
                        const bool first = (kb | ks) == 0;
                        for (int u = 0; u < nu; ++u) {
                            const unsigned tm = tmem_base + (unsigned)u * kSubCols;
                            const unsigned long long bd = i8_desc<KB>(st + (unsigned)(Cfg::A_BYTES + (size_t)u * Cfg::B_SUB)) + o;
                            auto digit = [&](int i, unsigned acc) {
                                const int cnt = i8_plane_count(i, DB);
                                i8_mma(tm + (unsigned)(i * kI8P), i8_desc<KB>(st + (unsigned)((size_t)i * Cfg::A_DIGIT)) + o, bd, idesc[cnt], acc);
                            };
                            if (DB < kI8Classes && first) digit(DA - 1, 0u);  // see i8_gemm_kernel: class 3 is first written by the last digit
                            digit(0, first ? 0u : 1u);
                            SCB_UNROLL
                            for (int i = 1; i < DA; ++i)
                                if (!(DB < kI8Classes && first && i == DA - 1)) digit(i, 1u);
                        }
                    }
                    i8_commit(empty(s));
                }
                i8_commit(acc_full);
            }
        }
    } else {
        // ===== epilogue: 8 warps, warp = (TMEM lane quarter q, sub-block u) =====
        constexpr int OP = Cfg::OUT_PITCH;
        const int q = warp & 3, u = (warp - 2) >> 2;
        float* tile = outbuf + (size_t)(warp - 2) * 32 * OP;
        int lt = 0;
        for (int unit = first_unit; unit < units; unit += unit_step, ++lt) {
            int par, sb0, nu, m0;
            decode(unit, par, sb0, nu, m0);
            const int sb = sb0 + u;
            const int nout = p.g.nout[par];
            const int line = m0 + 32 * q + lane;
            const float ls = line < p.m_rows ? __ldg(p.lscale + line) : 0.f;
            const float sc = p.scale * ls;
            const bool live = u < nu;
            float vals[kI8P];
            i8_mbar_wait(acc_full, lt & 1);
            i8_fence_after();
            if (live) {
                const unsigned tmq = tmem_base + ((unsigned)(32 * q) << 16) + (unsigned)u * kSubCols;
                SCB_UNROLL
                for (int cc = 0; cc < kI8P; cc += 16) {
                    int w0[16], w1[16], w2[16], w3[16];
                    i8_tmem_ld16(tmq + (unsigned)(0 * kI8P + cc), w0);
                    i8_tmem_ld16(tmq + (unsigned)(1 * kI8P + cc), w1);
                    i8_tmem_ld16(tmq + (unsigned)(2 * kI8P + cc), w2);
                    i8_tmem_ld16(tmq + (unsigned)(3 * kI8P + cc), w3);
                    i8_tmem_wait_ld();
                    SCB_UNROLL
                    for (int i = 0; i < 16; ++i) vals[cc + i] = i8_combine(w0[i], w1[i], w2[i], w3[i]) * sc;
                    if (cc == 0 && p.R && sb == 0 && line >= p.line0 && line < p.line1) {  // exact float64 row sums of the lowest frequencies
                        const int r = line / 3, c = line - 3 * r;  // lines are channel-interleaved: line = 3 row + channel
                        SCB_UNROLL
                        for (int i = 0; i < (kI8LowK + 1) / 2; ++i) {
                            const int k0 = 2 * i + par;
                            if (k0 < p.lowk && i < nout) p.R[((size_t)c * p.lowk + k0) * p.lpc + r] = i8_combine_exact(w0[i], w1[i], w2[i], w3[i]) * (p.rscale * (double)ls);
                        }
                    }
                }
            }
            // the accumulators are in registers now: hand them back to the MMA issuer before the stores
            i8_fence_before();
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(acc_empty) : "memory");
            if (live) {
                SCB_UNROLL
                for (int h = 0; h < 2; ++h) {  // two 32-column halves through the warp's transpose tile
                    SCB_UNROLL
                    for (int i = 0; i < 32; ++i) tile[(size_t)lane * OP + i] = vals[32 * h + i];
                    __syncwarp();
                    int ln = m0 + 32 * q;
                    int rr = ln / 3, c = ln - 3 * rr;
                    const int ki = sb * kI8P + 32 * h + lane;
                    for (int r = 0; r < 32 && ln < p.line1; ++r, ++ln) {
                        if (ki < nout && ln >= p.line0) {
                            const float v = tile[(size_t)r * OP + lane];
                            if (p.out_u8)
                                p.out_u8[(long long)rr * p.out_u8_pitch + 3 * (2 * ki + par) + c] = i8_to_u8(v);
                            else
                                p.out[(size_t)c * p.out_plane + (size_t)rr * p.out_pitch + 2 * ki + par] = v;
                        }
                        if (++c == 3) {
                            c = 0;
                            ++rr;
                        }
                    }
                    __syncwarp();  // the tile rows are free for the next half / the next tile
                }
            }
        }
    }
    __syncthreads();
    if (warp == 0) {
        i8_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

// ---- tensor maps -------------------------------------------------------------------------------
typedef CUresult (*I8EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                    CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static I8EncodeTiledFn i8_encode_fn() {
    static I8EncodeTiledFn fn = [] {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) f = nullptr;
        return (I8EncodeTiledFn)f;
    }();
    return fn;
}
// 2-D byte tensor [rows][kpad], box = kb bytes x box_rows, kb-byte swizzle
static int i8_make_map(CUtensorMap* map, const void* ptr, size_t rows, int kpad, int box_rows, int kb) {
    I8EncodeTiledFn enc = i8_encode_fn();
    if (!enc) return (int)cudaErrorNotSupported;
    const cuuint64_t gdim[2] = {(cuuint64_t)kpad, (cuuint64_t)rows};
    const cuuint64_t gstride[1] = {(cuuint64_t)kpad};
    const cuuint32_t box[2] = {(cuuint32_t)kb, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(ptr), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     kb == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : (int)cudaErrorInvalidValue;
}

static int i8_env_int(const char* name, int dflt) {
    const char* e = std::getenv(name);
    return e ? std::atoi(e) : dflt;
}
static constexpr int kI8PersistentDefault = 1;  // 0: one tile per CTA, 1: i8_gemm_pkernel (128 x 64 tiles, TMEM ping-pong), 2: i8_gemm_p2kernel (128 x 128 tiles), 3: p2 forward / p inverse
static int i8_sw_persistent() { return i8_env_int("SCB_I8_PERSISTENT", kI8PersistentDefault); }
static int i8_sw_kb() { return i8_env_int("SCB_I8_KB", 128) == 64 ? 64 : 128; }
template <int DA, int DB, int NSUB, int KB, int CL>
static int i8_launch_gemm_t5(void* stream, const I8GemmParams& p) {
    using Cfg = I8Cfg<DA, DB, NSUB, KB>;
    CUtensorMap amap, bmap;
    int rc;
    if ((rc = i8_make_map(&amap, p.a, (size_t)2 * DA * p.m_rows, p.g.kpad, kI8M, KB))) return rc;
    if ((rc = i8_make_map(&bmap, p.basis, i8_basis_rows(p.g), p.g.kpad, DB * kI8P / CL, KB))) return rc;  // CL = 2: each CTA loads half the rows
    const int mt = p.mt1 - p.mt0;
    if (mt <= 0) return 0;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((mt + CL - 1) / CL * CL, 2 * (p.g.nsb / NSUB));
    cfg.blockDim = dim3(kI8Threads);
    cfg.dynamicSmemBytes = Cfg::SMEM;
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    static const bool tracing = i8_env_int("SCB_I8_TRACE", 0) != 0;
    if (!tracing) return (int)cudaLaunchKernelEx(&cfg, i8_gemm_kernel<DA, DB, NSUB, KB, CL>, amap, bmap, p);
    // tuning aid: per-CTA phase stamps (clock64), averaged and printed; serialises the stream
    const size_t nct = (size_t)cfg.gridDim.x * cfg.gridDim.y;
    long long* d = nullptr;
    cudaMalloc(&d, nct * 8 * sizeof(long long));
    cudaMemset(d, 0, nct * 8 * sizeof(long long));
    I8GemmParams q = p;
    q.trace = d;
    cudaError_t e = cudaLaunchKernelEx(&cfg, i8_gemm_kernel<DA, DB, NSUB, KB, CL>, amap, bmap, q);
    cudaStreamSynchronize((cudaStream_t)stream);
    std::vector<long long> h(nct * 8);
    cudaMemcpy(h.data(), d, h.size() * sizeof(long long), cudaMemcpyDeviceToHost);
    cudaFree(d);
    double acc[7] = {0, 0, 0, 0, 0, 0, 0};
    for (size_t i = 0; i < nct; ++i)
        for (int k = 1; k <= 6; ++k) acc[k] += (double)(h[8 * i + k] - h[8 * i + k - 1]);
    std::fprintf(stderr, "i8_gemm<%d,%d,%d,%d,%d> %zu CTAs, mean cycles: setup %.0f | first stage %.0f | mainloop issue %.0f | mma drain %.0f | epilogue %.0f | join %.0f\n", DA, DB,
                 NSUB, KB, CL, nct, acc[1] / nct, acc[2] / nct, acc[3] / nct, acc[4] / nct, acc[5] / nct, acc[6] / nct);
    return (int)e;
}
template <int DA, int DB, int KB, int CL>
static int i8_launch_gemm_p(void* stream, const I8GemmParams& p) {
    using Cfg = I8PCfg<DA, DB, KB>;
    CUtensorMap amap, bmap;
    int rc;
    if ((rc = i8_make_map(&amap, p.a, (size_t)2 * DA * p.m_rows, p.g.kpad, kI8M, KB))) return rc;
    if ((rc = i8_make_map(&bmap, p.basis, i8_basis_rows(p.g), p.g.kpad, DB * kI8P / CL, KB))) return rc;
    static const int sms = [] {
        int dev = 0, n = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        return n;
    }();
    const int mt = p.mt1 - p.mt0;
    if (mt <= 0) return 0;
    const int units = ((p.g.nout[0] + kI8P - 1) / kI8P + (p.g.nout[1] + kI8P - 1) / kI8P) * ((mt + CL - 1) / CL);
    cudaLaunchConfig_t cfg = {};
    cfg.blockDim = dim3(kI8Threads);
    cfg.dynamicSmemBytes = Cfg::SMEM;
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    // one resident wave: as many clusters as the device holds at once (clusters must fit inside a GPC, so fewer than sms / CL may)
    static const int max_clusters = [&] {
        int n = sms / CL;
        if (CL > 1) {
            cudaLaunchConfig_t q = cfg;
            q.gridDim = dim3(sms / CL * CL);
            int c = 0;
            if (cudaOccupancyMaxActiveClusters(&c, i8_gemm_pkernel<DA, DB, KB, CL>, &q) == cudaSuccess && c > 0 && c < n) n = c;
        }
        return n;
    }();
    int clusters = max_clusters;
    if (clusters > units) clusters = units;
    cfg.gridDim = dim3(clusters * CL);
    return (int)cudaLaunchKernelEx(&cfg, i8_gemm_pkernel<DA, DB, KB, CL>, amap, bmap, p);
}
// 128 x 128-output tiles (i8_gemm_p2kernel): one CTA per SM, as many as there are tiles
template <int DA, int DB, int KB>
static int i8_launch_gemm_p2(void* stream, const I8GemmParams& p) {
    using Cfg = I8P2Cfg<DA, DB, KB>;
    CUtensorMap amap, bmap;
    int rc;
    if ((rc = i8_make_map(&amap, p.a, (size_t)2 * DA * p.m_rows, p.g.kpad, kI8M, KB))) return rc;
    if ((rc = i8_make_map(&bmap, p.basis, i8_basis_rows(p.g), p.g.kpad, DB * kI8P, KB))) return rc;
    static const int sms = [] {
        int dev = 0, n = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        return n;
    }();
    const int mt = p.mt1 - p.mt0;
    if (mt <= 0) return 0;
    const int nsbr0 = (p.g.nout[0] + kI8P - 1) / kI8P, nsbr1 = (p.g.nout[1] + kI8P - 1) / kI8P;
    const int units = ((nsbr0 + 1) / 2 + (nsbr1 + 1) / 2) * mt;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(units < sms ? units : sms);
    cfg.blockDim = dim3(kI8Threads);
    cfg.dynamicSmemBytes = Cfg::SMEM;
    cfg.stream = (cudaStream_t)stream;
    return (int)cudaLaunchKernelEx(&cfg, i8_gemm_p2kernel<DA, DB, KB>, amap, bmap, p);
}
template <int DA, int DB, int KB>
static cudaError_t i8_set_smem_p2() {
    return cudaFuncSetAttribute(i8_gemm_p2kernel<DA, DB, KB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)I8P2Cfg<DA, DB, KB>::SMEM);
}
template <int DA, int DB, int KB, int CL>
static cudaError_t i8_set_smem_p() {
    return cudaFuncSetAttribute(i8_gemm_pkernel<DA, DB, KB, CL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)I8PCfg<DA, DB, KB>::SMEM);
}

// Tile shape: SCB_I8_NSUB=2 -> 128 x 128 outputs, 128-byte rows, one CTA per SM; 1 (default) -> 128 x 64 outputs, 64-byte rows, two CTAs
// per SM.  SCB_I8_CLUSTER=2 adds the basis multicast over pairs of CTAs.  (Environment switches are A/B aids; the defaults are the
// measured best, profiles/r2_i8_variants.txt.)
template <int DA, int DB>
static int i8_launch_gemm_t(void* stream, const I8GemmParams& p) {
    static const int nsub = i8_env_int("SCB_I8_NSUB", 1) == 2 ? 2 : 1;
    static const int clp = i8_env_int("SCB_I8_CLUSTER", 1);
    static const int cl = clp == 2 ? 2 : 1;
    static const int persistent = i8_sw_persistent();
    static const int kb = i8_sw_kb();
    if (persistent == 2 || persistent == 3) {  // 128 x 128-output tiles; 3: for the forward pass only (the inverse pass, whose 4 + 3
        // digits only fit with 64-byte stage rows and three stages, measured slower on them: profiles/r2b_ab_variants_call1.log)
        if constexpr (DA == 2 && DB == 4) return kb == 64 ? i8_launch_gemm_p2<2, 4, 64>(stream, p) : i8_launch_gemm_p2<2, 4, 128>(stream, p);
        if constexpr (DA == 4 && DB == 3)
            if (persistent == 2) return i8_launch_gemm_p2<4, 3, 64>(stream, p);
    }
    if (persistent) {
        if constexpr (!(DA == 4 && DB == 4)) {  // 4 + 4 digits: 96 KB per 128-byte-row stage, only the 64-byte rows leave two stages
            if (kb == 128) {
                if (clp == 8) return i8_launch_gemm_p<DA, DB, 128, 8>(stream, p);
                if (clp == 4) return i8_launch_gemm_p<DA, DB, 128, 4>(stream, p);
                return cl == 2 ? i8_launch_gemm_p<DA, DB, 128, 2>(stream, p) : i8_launch_gemm_p<DA, DB, 128, 1>(stream, p);
            }
        }
        return cl == 2 ? i8_launch_gemm_p<DA, DB, 64, 2>(stream, p) : i8_launch_gemm_p<DA, DB, 64, 1>(stream, p);
    }
    if constexpr (!(DA == 4 && DB == 4)) {
        if (nsub == 2) return cl == 2 ? i8_launch_gemm_t5<DA, DB, 2, 128, 2>(stream, p) : i8_launch_gemm_t5<DA, DB, 2, 128, 1>(stream, p);
    }
    return cl == 2 ? i8_launch_gemm_t5<DA, DB, 1, 64, 2>(stream, p) : i8_launch_gemm_t5<DA, DB, 1, 64, 1>(stream, p);
}
template <int DA, int DB, int NSUB, int KB, int CL>
static cudaError_t i8_set_smem() {
    return cudaFuncSetAttribute(i8_gemm_kernel<DA, DB, NSUB, KB, CL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)I8Cfg<DA, DB, NSUB, KB>::SMEM);
}
template <int DA, int DB>
static cudaError_t i8_set_smem_all() {
    cudaError_t e;
    if constexpr (!(DA == 4 && DB == 4)) {
        if ((e = i8_set_smem_p<DA, DB, 128, 1>()) != cudaSuccess || (e = i8_set_smem_p<DA, DB, 128, 2>()) != cudaSuccess) return e;
        if ((e = i8_set_smem_p<DA, DB, 128, 4>()) != cudaSuccess || (e = i8_set_smem_p<DA, DB, 128, 8>()) != cudaSuccess) return e;
    }
    if ((e = i8_set_smem_p<DA, DB, 64, 1>()) != cudaSuccess || (e = i8_set_smem_p<DA, DB, 64, 2>()) != cudaSuccess) return e;
    if ((e = i8_set_smem<DA, DB, 1, 64, 1>()) != cudaSuccess || (e = i8_set_smem<DA, DB, 1, 64, 2>()) != cudaSuccess) return e;
    if constexpr (!(DA == 4 && DB == 4)) {
        if ((e = i8_set_smem<DA, DB, 2, 128, 1>()) != cudaSuccess || (e = i8_set_smem<DA, DB, 2, 128, 2>()) != cudaSuccess) return e;
    }
    return cudaSuccess;
}
int i8_configure() {
    cudaError_t e;
    if ((e = i8_set_smem_p2<2, 4, 128>()) != cudaSuccess || (e = i8_set_smem_p2<2, 4, 64>()) != cudaSuccess || (e = i8_set_smem_p2<4, 3, 64>()) != cudaSuccess) return (int)e;
    if ((e = i8_set_smem_all<2, 4>()) != cudaSuccess || (e = i8_set_smem_all<4, 4>()) != cudaSuccess || (e = i8_set_smem_all<4, 3>()) != cudaSuccess) return (int)e;
    return 0;
}
#endif  // SCB_EMU

// ---------------------------------------------------------------------------------------------
// host launchers
// ---------------------------------------------------------------------------------------------
int i8_launch_basis(void* stream, const I8Geom& g, signed char* basis) {
    const long long total = 2LL * g.nsb * kI8P * (g.kpad / 4);
    long long blocks = (total + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    SCB_LAUNCH(i8_basis_kernel, dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)stream, g, basis);
#ifdef SCB_EMU
    return 0;
#else
    return (int)cudaGetLastError();
#endif
}

int i8_launch_compose(void* stream, const I8ComposeParams& p, int rows) {
    if (rows <= 0) return 0;
    const int groups = (3 * p.nx + 3 + 11) / 12;
    SCB_LAUNCH(i8_compose_kernel, dim3((groups + 127) / 128, rows), dim3(128), 0, (cudaStream_t)stream, p);
#ifdef SCB_EMU
    return 0;
#else
    return (int)cudaGetLastError();
#endif
}

const char* i8_variant_string() {
#ifdef SCB_EMU
    static const std::string v = "i8_gemm=emulator";
#else
    static const std::string v = "i8_persistent=" + std::to_string(i8_sw_persistent()) + " i8_kb=" + std::to_string(i8_sw_kb());
#endif
    return v.c_str();
}
int i8_launch_digitize(void* stream, const I8DigitizeParams& p, int da) {
    if (p.line1 <= p.line0) return 0;
    const dim3 grid(p.line1 - p.line0), block(kI8DigThreads);
    if (da == 2)
        SCB_LAUNCH(i8_digitize_kernel<2>, grid, block, 0, (cudaStream_t)stream, p);
    else
        SCB_LAUNCH(i8_digitize_kernel<4>, grid, block, 0, (cudaStream_t)stream, p);
#ifdef SCB_EMU
    return 0;
#else
    return (int)cudaGetLastError();
#endif
}

int i8_launch_gemm(void* stream, const I8GemmParams& p, int da, int db) {
    if (da == 2 && db == 4) return i8_launch_gemm_t<2, 4>(stream, p);
    if (da == 4 && db == 4) return i8_launch_gemm_t<4, 4>(stream, p);
    if (da == 4 && db == 3) return i8_launch_gemm_t<4, 3>(stream, p);
    return 1;  // cudaErrorInvalidValue
}

}  // namespace scb
