// scb_i8.cu -- kernels of the exact INT8 tensor-core DST engine (see scb_i8.h for the arithmetic).
//
// Replaces, along x, the reference's per-row transform  dft_kernel_0 -> cufftExecC2C -> dft_kernel_1
// (/root/reference/seamlessClone-CUDA/seamlessClone_imp.cpp:1694-1811) and its GEMM flavour (cublasSgemmBatched,
// imp.cpp:1266-1334), with the same mathematical operator and results that are exact up to 2^-31 of full scale.
#include "scb_i8.h"

#include <cmath>
#include <cstring>

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
// lines -> folded digit planes.  Block = 4 lines x 64 threads; a thread converts 16 consecutive folded elements at a time.
//   v[j] = rint(x[j] * s),  f0 = v[j] + v[n-1-j],  f1 = v[j] - v[n-1-j]  (j < n/2;  the middle element of an odd line goes to f0)
//   a[(par * DA + i) * m_rows + line][j] = digit i of f_par[j]
// s = fixed_scale, or (per_line) 2^(29 - e) with 2^e > max |x| over the line, so |v| <= 2^29 and |f| <= 2^30.
// ---------------------------------------------------------------------------------------------
static constexpr int kI8DigLines = 4;
static constexpr int kI8DigThreads = 64;

template <int DA>
__global__ void __launch_bounds__(kI8DigLines* kI8DigThreads) i8_digitize_kernel(I8DigitizeParams p) {
    __shared__ float red[kI8DigLines][2];
    const int lid = threadIdx.x / kI8DigThreads, t = threadIdx.x % kI8DigThreads;
    const int line = blockIdx.x * kI8DigLines + lid;
    const int n = p.g.n, h = n >> 1;
    const bool real = line < p.lines;
    const float* x = nullptr;
    if (real) {
        const int c = line / p.lpc, r = line - c * p.lpc;
        x = p.in + (size_t)c * p.in_plane + (size_t)r * p.in_pitch;
    }
    float s = p.fixed_scale;
    if (p.per_line) {
        float m = 0.f;
        if (real)
            for (int j = t; j < n; j += kI8DigThreads) m = fmaxf(m, fabsf(x[j]));
        for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        if ((t & 31) == 0) red[lid][t >> 5] = m;
        __syncthreads();
        m = fmaxf(red[lid][0], red[lid][1]);
        int e = 0;
        if (m > 0.f) (void)frexpf(m, &e);  // m = f 2^e, 0.5 <= f < 1
        if (e < -60) e = -60;
        if (e > 90) e = 90;
        s = ldexpf(1.0f, 29 - e);
        if (t == 0 && line < p.m_rows) p.lscale[line] = ldexpf(1.0f, e - 29);
    } else if (t == 0 && line < p.m_rows) {
        p.lscale[line] = 1.0f / p.fixed_scale;
    }
    if (line >= p.m_rows) return;
    for (int j0 = 16 * t; j0 < p.g.kpad; j0 += 16 * kI8DigThreads) {
        unsigned w[2][DA][4];
        SCB_UNROLL
        for (int q = 0; q < 2; ++q)
            SCB_UNROLL
            for (int i = 0; i < DA; ++i)
                SCB_UNROLL
                for (int e = 0; e < 4; ++e) w[q][i][e] = 0u;
        if (real && j0 < p.g.kpar[0]) {
            SCB_UNROLL
            for (int e = 0; e < 16; ++e) {
                const int j = j0 + e;
                int f0 = 0, f1 = 0;
                if (j < h) {
                    const int va = __float2int_rn(x[j] * s), vb = __float2int_rn(x[n - 1 - j] * s);
                    f0 = va + vb;
                    f1 = va - vb;
                } else if (j == h && (n & 1)) {
                    f0 = __float2int_rn(x[h] * s);
                }
                int d0[DA], d1[DA];
                i8_digits<DA>(f0, d0);
                i8_digits<DA>(f1, d1);
                SCB_UNROLL
                for (int i = 0; i < DA; ++i) {
                    w[0][i][e >> 2] |= (unsigned)(d0[i] & 255) << (8 * (e & 3));
                    w[1][i][e >> 2] |= (unsigned)(d1[i] & 255) << (8 * (e & 3));
                }
            }
        }
        SCB_UNROLL
        for (int q = 0; q < 2; ++q)
            SCB_UNROLL
            for (int i = 0; i < DA; ++i) {
                signed char* dst = p.a + ((size_t)(q * DA + i) * p.m_rows + line) * p.g.kpad + j0;
                *reinterpret_cast<uint4*>(dst) = make_uint4(w[q][i][0], w[q][i][1], w[q][i][2], w[q][i][3]);
            }
    }
}

// which basis planes digit i of the lines multiplies: planes 0 .. cnt-1, landing in classes i .. i+cnt-1
SCB_HD int i8_plane_count(int i, int db) {
    const int left = kI8Classes - i;
    return left < db ? left : db;
}

// epilogue of one output element, shared by the tensor-core kernel and the emulator kernel
SCB_D void i8_store(const I8GemmParams& p, int line, int par, int ki, int w0, int w1, int w2, int w3, float ls) {
    if (line >= p.lines || ki >= p.g.nout[par]) return;
    const int c = line / p.lpc, r = line - c * p.lpc;
    const int k0 = 2 * ki + par;
    p.out[(size_t)c * p.out_plane + (size_t)r * p.out_pitch + k0] = i8_combine(w0, w1, w2, w3) * (p.scale * ls);
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
    const int line = blockIdx.x * kI8M + threadIdx.x;
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
    SCB_LAUNCH((i8_gemm_kernel<DA, DB>), dim3((p.lines + kI8M - 1) / kI8M, 2 * p.g.nsb), dim3(kI8M), 0, stream, p);
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
// K-major operand tile, 128-byte rows, 128-byte swizzle: 8-row atoms of 1024 bytes, SBO = 1024, LBO unused.
SCB_D unsigned long long i8_desc_sw128(unsigned smem_addr) {
    unsigned long long d = 0;
    d |= (unsigned long long)((smem_addr >> 4) & 0x3fff);
    d |= (unsigned long long)(1024 >> 4) << 32;
    d |= (unsigned long long)1 << 46;  // descriptor version (sm_100)
    d |= (unsigned long long)2 << 61;  // SWIZZLE_128B
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

static constexpr int kI8Threads = 192;  // warp 0: TMA producer + TMEM allocator, warp 1: MMA issuer, warps 2-5: epilogue
static constexpr int kI8Stages = 2;
static constexpr int kI8TmemCols = 512;

template <int DA, int DB, int NSUB>
struct I8Cfg {
    static constexpr size_t A_BYTES = (size_t)DA * kI8M * kI8KB;   // 16 KB per digit of the lines
    static constexpr size_t B_PLANE = (size_t)kI8P * kI8KB;        // 8 KB per digit plane of the basis
    static constexpr size_t B_SUB = (size_t)DB * B_PLANE;
    static constexpr size_t STAGE = A_BYTES + NSUB * B_SUB;
    static constexpr size_t OUT_STAGE = (size_t)kI8M * (kI8P + 1) * sizeof(float);  // epilogue transpose buffer (reuses the stages)
    static constexpr size_t SMEM = kI8Stages * STAGE + 1024 /* alignment slack */ + 256 /* barriers */;
    static_assert(OUT_STAGE <= kI8Stages * STAGE, "epilogue buffer must fit in the pipeline stages");
    static_assert(SMEM <= 232448, "shared memory budget of one CTA");
};

// grid = (line tiles, 2 * nsb / NSUB): blockIdx.y / (nsb / NSUB) = parity, the rest = first sub-block / NSUB
template <int DA, int DB, int NSUB>
__global__ void __launch_bounds__(kI8Threads, 1)
i8_gemm_kernel(const __grid_constant__ CUtensorMap amap, const __grid_constant__ CUtensorMap bmap, I8GemmParams p) {
    using Cfg = I8Cfg<DA, DB, NSUB>;
    extern __shared__ unsigned char i8_smem_raw[];
    unsigned char* base = reinterpret_cast<unsigned char*>(((size_t)i8_smem_raw + 1023) & ~(size_t)1023);
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(base + (size_t)kI8Stages * Cfg::STAGE);
    const unsigned bar0 = i8_smem_u32(bars);
    auto full = [&](int s) { return bar0 + 8u * s; };
    auto empty = [&](int s) { return bar0 + 8u * (kI8Stages + s); };
    const unsigned acc_bar = bar0 + 8u * (2 * kI8Stages);
    unsigned* tmem_slot = reinterpret_cast<unsigned*>(bars + 2 * kI8Stages + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int groups = p.g.nsb / NSUB;
    const int par = blockIdx.y / groups, sb0 = (blockIdx.y % groups) * NSUB;
    const int kpar = p.g.kpar[par];
    const int num_kb = (kpar + kI8KB - 1) / kI8KB;
    const int m0 = blockIdx.x * kI8M;

    if (warp == 1 && lane == 0) {
        for (int s = 0; s < kI8Stages; ++s) {
            i8_mbar_init(full(s), 1);   // the TMA thread's arrive.expect_tx
            i8_mbar_init(empty(s), 1);  // tcgen05.commit
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
    i8_fence_after();
    const unsigned tmem_base = *tmem_slot;

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
                for (int i = 0; i < DA; ++i) i8_tma_2d(i8_smem_u32(st + (size_t)i * kI8M * kI8KB), &amap, kb * kI8KB, (par * DA + i) * p.m_rows + m0, full(s));
                SCB_UNROLL
                for (int u = 0; u < NSUB; ++u)
                    i8_tma_2d(i8_smem_u32(st + Cfg::A_BYTES + (size_t)u * Cfg::B_SUB), &bmap, kb * kI8KB, ((par * p.g.nsb + sb0 + u) * kI8BasisDigits) * kI8P, full(s));
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
                const unsigned st = i8_smem_u32(base + (size_t)s * Cfg::STAGE);
                const int left = kpar - kb * kI8KB;
                const int nks = left >= kI8KB ? kI8KB / 32 : (left + 31) / 32;
                for (int ks = 0; ks < nks; ++ks) {
                    const unsigned long long o = (unsigned long long)(ks * 2);  // 32 bytes of K = 2 x 16-byte units of the start address
                    const bool first = (kb | ks) == 0;
                    SCB_UNROLL
                    for (int u = 0; u < NSUB; ++u) {
                        const unsigned tm = tmem_base + (unsigned)(u * kI8Classes * kI8P);
                        const unsigned long long bd = i8_desc_sw128(st + (unsigned)(Cfg::A_BYTES + (size_t)u * Cfg::B_SUB)) + o;
                        auto digit = [&](int i, unsigned acc) {
                            const int cnt = i8_plane_count(i, DB);
                            i8_mma(tm + (unsigned)(i * kI8P), i8_desc_sw128(st + (unsigned)((size_t)i * kI8M * kI8KB)) + o, bd, idesc[cnt], acc);
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
                i8_commit(empty(s));  // frees the stage when these MMAs have read it
            }
            i8_commit(acc_bar);
        }
    } else {
        // ===== epilogue: warp w owns TMEM lanes 32 (w % 4) .. +31 = lines; class sums -> float -> smem transpose -> coalesced rows =====
        i8_mbar_wait(acc_bar, 0);
        i8_fence_after();
        const int q = warp & 3;
        const int line = m0 + 32 * q + lane;
        const float ls = line < p.m_rows ? __ldg(p.lscale + line) : 0.f;
        float* tile = reinterpret_cast<float*>(base) + (size_t)(32 * q) * (kI8P + 1);  // this warp's 32 rows of the transpose buffer
        const float sc = p.scale;
        for (int u = 0; u < NSUB; ++u) {
            const int sb = sb0 + u;
            if (sb * kI8P >= p.g.nout[par]) break;
            const unsigned tm = tmem_base + ((unsigned)(32 * q) << 16) + (unsigned)(u * kI8Classes * kI8P);
            SCB_UNROLL
            for (int cc = 0; cc < kI8P; cc += 16) {
                int w0[16], w1[16], w2[16], w3[16];
                i8_tmem_ld16(tm + (unsigned)(0 * kI8P + cc), w0);
                i8_tmem_ld16(tm + (unsigned)(1 * kI8P + cc), w1);
                i8_tmem_ld16(tm + (unsigned)(2 * kI8P + cc), w2);
                i8_tmem_ld16(tm + (unsigned)(3 * kI8P + cc), w3);
                i8_tmem_wait_ld();
                SCB_UNROLL
                for (int i = 0; i < 16; ++i) tile[(size_t)lane * (kI8P + 1) + cc + i] = i8_combine(w0[i], w1[i], w2[i], w3[i]) * (sc * ls);
                if (p.R && cc == 0 && sb == 0 && line < p.lines) {  // exact float64 row sums of the lowest frequencies
                    const int c = line / p.lpc, r = line - c * p.lpc;
                    SCB_UNROLL
                    for (int i = 0; i < (kI8LowK + 1) / 2; ++i) {
                        const int k0 = 2 * i + par;
                        if (k0 < p.lowk && i < p.g.nout[par]) p.R[((size_t)c * p.lowk + k0) * p.lpc + r] = i8_combine_exact(w0[i], w1[i], w2[i], w3[i]) * (p.rscale * (double)ls);
                    }
                }
            }
            __syncwarp();
            // rows of the tile -> rows of the output: lane l writes outputs ki = 64 sb + l and + 32 of one line (stride-2 floats: the
            // other parity's CTA fills the gaps)
            const int nout = p.g.nout[par];
            for (int r = 0; r < 32; ++r) {
                const int ln = m0 + 32 * q + r;
                if (ln >= p.lines) break;
                const int c = ln / p.lpc, rr = ln - c * p.lpc;
                float* o = p.out + (size_t)c * p.out_plane + (size_t)rr * p.out_pitch;
                const float* trow = tile + (size_t)r * (kI8P + 1);
                SCB_UNROLL
                for (int hh = 0; hh < 2; ++hh) {
                    const int ki = sb * kI8P + 32 * hh + lane;
                    if (ki < nout) o[2 * ki + par] = trow[32 * hh + lane];
                }
            }
            __syncwarp();
        }
        i8_fence_before();
    }
    __syncthreads();
    if (warp == 0) {
        i8_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kI8TmemCols) : "memory");
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
// 2-D byte tensor [rows][kpad], box = 128 bytes x box_rows, 128-byte swizzle
static int i8_make_map(CUtensorMap* map, const void* ptr, size_t rows, int kpad, int box_rows) {
    I8EncodeTiledFn enc = i8_encode_fn();
    if (!enc) return (int)cudaErrorNotSupported;
    const cuuint64_t gdim[2] = {(cuuint64_t)kpad, (cuuint64_t)rows};
    const cuuint64_t gstride[1] = {(cuuint64_t)kpad};
    const cuuint32_t box[2] = {(cuuint32_t)kI8KB, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(ptr), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : (int)cudaErrorInvalidValue;
}

template <int DA, int DB, int NSUB>
static int i8_launch_gemm_t3(void* stream, const I8GemmParams& p) {
    using Cfg = I8Cfg<DA, DB, NSUB>;
    CUtensorMap amap, bmap;
    int rc;
    if ((rc = i8_make_map(&amap, p.a, (size_t)2 * DA * p.m_rows, p.g.kpad, kI8M))) return rc;
    if ((rc = i8_make_map(&bmap, p.basis, i8_basis_rows(p.g), p.g.kpad, DB * kI8P))) return rc;
    const dim3 grid((p.lines + kI8M - 1) / kI8M, 2 * (p.g.nsb / NSUB));
    i8_gemm_kernel<DA, DB, NSUB><<<grid, kI8Threads, Cfg::SMEM, (cudaStream_t)stream>>>(amap, bmap, p);
    return (int)cudaGetLastError();
}
template <int DA, int DB>
static int i8_launch_gemm_t(void* stream, const I8GemmParams& p) {
    if constexpr (DA == 4 && DB == 4) return i8_launch_gemm_t3<DA, DB, 1>(stream, p);  // 64 + 32 KB per stage: one sub-block per CTA
    else return i8_launch_gemm_t3<DA, DB, 2>(stream, p);
}
int i8_configure() {
    cudaError_t e;
    if ((e = cudaFuncSetAttribute(i8_gemm_kernel<2, 4, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)I8Cfg<2, 4, 2>::SMEM)) != cudaSuccess) return (int)e;
    if ((e = cudaFuncSetAttribute(i8_gemm_kernel<4, 4, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)I8Cfg<4, 4, 1>::SMEM)) != cudaSuccess) return (int)e;
    if ((e = cudaFuncSetAttribute(i8_gemm_kernel<4, 3, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)I8Cfg<4, 3, 2>::SMEM)) != cudaSuccess) return (int)e;
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

int i8_launch_digitize(void* stream, const I8DigitizeParams& p, int da) {
    const dim3 grid((p.m_rows + kI8DigLines - 1) / kI8DigLines), block(kI8DigLines * kI8DigThreads);
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
