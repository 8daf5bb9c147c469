// scb_tc.cuh -- the DST-I passes as dense contractions on the 5th-generation tensor cores (tcgen05).
//
// OpenCV's Cloning::dst is, mathematically, a multiplication by the sine matrix
//     S[j][k] = sin(pi (j+1)(k+1) / N),  N = n+1                      (reference: the GEMM flavour of the
//     solver, /root/reference/seamlessClone-CUDA/seamlessClone_imp.cpp:488-665, 1118-1334, cublasSgemmBatched)
// For the ROI sides of BASELINE's configs (n ~ 400 .. 4000) that contraction, run on tensor cores, is
// cheaper than the Bluestein FFT engine (scb_gfft.cuh), whose convolution length is 2-3x the line length
// and which runs on the FP32 pipes.  Three devices make it cheap and accurate enough:
//
//  (1) Even/odd folding.  sin(pi (N-j) k / N) = (-1)^(k+1) sin(pi j k / N): outputs with odd k see only
//      x[j] + x[N-j], outputs with even k only x[j] - x[N-j].  The n x n product becomes two products of
//      half the size on folded inputs: 2x fewer flops.  The fold is done by the producer warps on the way
//      from global to shared memory.
//  (2) 3xTF32.  Operands are split a = a_hi + a_lo, s = s_hi + s_lo into TF32-representable parts
//      (cvt.rna.tf32.f32); D += a_hi s_hi + a_lo s_hi + a_hi s_lo with FP32 accumulation in TMEM.  The
//      dropped a_lo s_lo term is 2^-22 relative.  Measured end to end (profiles/, tests): the final image
//      matches cv2.seamlessClone exactly as often as a float64 solve does.
//  (3) The exact low-frequency refinement of the FFT engine (lowfreq_rows / lowfreq_cols kernels) is kept unchanged.
//
// Superseded as the tensor-core path by the exact INT8 engine (scb_i8.h): FP32 accumulation in TMEM truncates at every MMA step,
// ~1e-5 relative at K ~ 900, which costs 0.2-0.3 % of exactly matching bytes.  Kept selectable (SCB_ENGINE_TC) for comparison.
//
// One pass = one kernel:  out = epilogue( fold(in) x basis ).
//   CTA tile : 128 lines (UMMA M) x NT output bins of one parity (UMMA N <= 256), K loop over the folded line.
//   warp 0   : TMA producer of the basis tiles (cp.async.bulk.tensor, 128-byte swizzle), TMEM allocator
//   warp 1   : MMA issuer (one elected lane): 4 k-steps x 3 tcgen05.mma.kind::tf32 per 32-element k-block
//   warps 2-5: A producers (global -> fold -> hi/lo split -> swizzled shared memory), then the epilogue
//              (tcgen05.ld of their 32 TMEM lanes, scale / eigenvalue division, transposed or plain store)
//   shared   : 2 stages x (A_hi 16 KB, A_lo 16 KB, B_hi 32 KB, B_lo 32 KB), mbarriers full_A/full_B/empty
// The four passes of a solve:
//   rows fwd  G [c][y][x]     -> At [c][kx][y]   (transposed store)         x(-2)
//   cols fwd  At[c][kx][y]    -> Q  [c][kx][ky]  (plain store)              x(-2), refinement, / (fx+fy-4)
//   cols inv  Q [c][kx][ky]   -> Ct [c][y][kx]   (transposed store)         / (ny+1)
//   rows inv  Ct[c][y][kx]    -> U  [c][y][x]    (plain store)              / (nx+1)   -> compose kernel -> u8
#pragma once

#include "scb_platform.h"

#ifndef SCB_EMU
#include <cuda.h>  // CUtensorMap (types only; the encoder is fetched with cudaGetDriverEntryPoint)
#endif

namespace scb {

static constexpr int kTcM = 128;        // lines per CTA tile (UMMA M)
static constexpr int kTcKB = 32;        // K elements per k-block: one 128-byte swizzle atom of fp32
static constexpr int kTcMaxNT = 256;    // widest UMMA N
static constexpr int kTcStages = 2;
static constexpr int kTcThreads = 192;  // 6 warps
static constexpr int kTcMinN = 16;      // shorter lines stay on the FFT engine
static constexpr int kTcMaxN = 4096;    // longer lines stay on the FFT engine (dense flops grow as n^2 per line)

static constexpr size_t kTcStageA = (size_t)kTcM * 128;          // one of A_hi / A_lo
static constexpr size_t kTcStageB = (size_t)kTcMaxNT * 128;      // one of B_hi / B_lo
static constexpr size_t kTcStageBytes = 2 * kTcStageA + 2 * kTcStageB;
static constexpr size_t kTcSmemBytes = kTcStages * kTcStageBytes + 1024 /* alignment slack */ + 256 /* barriers */;

// round to nearest (ties away) TF32: what cvt.rna.tf32.f32 returns
SCB_HD float tf32_round(float x) {
#if defined(__CUDA_ARCH__)
    unsigned u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
    return __uint_as_float(u);
#else
    unsigned u;
    std::memcpy(&u, &x, 4);
    if ((u & 0x7f800000u) != 0x7f800000u) u = (u + 0x1000u) & 0xffffe000u;
    float r;
    std::memcpy(&r, &u, 4);
    return r;
#endif
}

// ---------------------------------------------------------------------------------------------
// per-length basis table (plan time).  For parity p (0: k0 even <-> odd k, folded input x[j] + x[n-1-j];
// 1: k0 odd, x[j] - x[n-1-j]) and plane q (hi, lo):  B[p][q][ki][j] = split_q( sin(pi (j+1)(2 ki + p + 1) / N) ),
// K-major rows of kpad floats, zero outside [nout[p]) x [kpar[p]).  Rows are stacked [p0 hi | p0 lo | p1 hi | p1 lo],
// `rows` rows each, so that one 2-D tensor map covers the table.
// ---------------------------------------------------------------------------------------------
struct TcTabDev {
    int n = 0;
    int kpad = 0;      // row pitch in floats, multiple of 32
    int rows = 0;      // rows per (parity, plane) block, multiple of nt
    int nt = 0;        // UMMA N for this length (multiple of 16, <= 256)
    int ntiles = 0;    // N tiles per parity
    int kpar[2] = {0, 0};
    int nout[2] = {0, 0};
    const float* basis = nullptr;
};

SCB_HD void tc_geometry(int n, TcTabDev* t) {
    const int h = n / 2;
    t->n = n;
    t->kpar[0] = h + (n & 1);
    t->kpar[1] = h;
    t->nout[0] = (n + 1) / 2;
    t->nout[1] = n / 2;
    t->kpad = (t->kpar[0] + kTcKB - 1) / kTcKB * kTcKB;
    t->ntiles = (t->nout[0] + kTcMaxNT - 1) / kTcMaxNT;
    const int per = (t->nout[0] + t->ntiles - 1) / t->ntiles;
    t->nt = (per + 15) / 16 * 16;
    t->rows = t->nt * t->ntiles;
}

__global__ void __launch_bounds__(256) tc_basis_kernel(TcTabDev t, float* basis) {
    const long long per_block = (long long)t.rows * t.kpad;
    const long long total = 2 * per_block;  // (parity, ki, j); hi and lo written together
    const double invN = 1.0 / (double)(t.n + 1);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int p = (int)(i / per_block);
        const long long rem = i - (long long)p * per_block;
        const int ki = (int)(rem / t.kpad), j = (int)(rem - (long long)ki * t.kpad);
        float hi = 0.f, lo = 0.f;
        if (ki < t.nout[p] && j < t.kpar[p]) {
            const long long N2 = 2LL * (t.n + 1);
            const long long e = ((long long)(j + 1) * (2 * ki + p + 1)) % N2;  // exact argument reduction
            const double s = sinpi((double)e * invN);
            hi = tf32_round((float)s);
            lo = tf32_round((float)(s - (double)hi));
        }
        basis[((long long)(2 * p + 0) * t.rows + ki) * t.kpad + j] = hi;
        basis[((long long)(2 * p + 1) * t.rows + ki) * t.kpad + j] = lo;
    }
}

// ---------------------------------------------------------------------------------------------
// one pass
// ---------------------------------------------------------------------------------------------
struct TcPassParams {
    TcTabDev tab;
    const float* in;         // line (c, r) starts at in + c*in_plane + r*in_pitch; n contiguous floats; pitches multiples of 4
    long long in_plane;
    int in_pitch;
    int lpc;                 // lines per channel; 3*lpc lines in all
    int line0, line_end;     // global line range of this launch (sharding hook; [0, 3*lpc) normally)
    float* out;
    long long out_plane;
    int out_pitch;
    int transposed;          // 1: out[c*plane + k0*pitch + r]     0: out[c*plane + r*pitch + k0]
    float scale;             // multiplies the contraction
    // cols-forward extras (null / 0 otherwise): eigenvalue division and the exact low-frequency corner
    const float* f_line;     // filter indexed by the line's r   (OpenCV filter_X: the line is a column kx)
    const float* f_out;      // filter indexed by k0             (OpenCV filter_Y)
    const float* lowspec;    // [3][lowk_line][lowk_out]
    int lowk_line, lowk_out;
    float* spec_dump;        // [3][lpc][n] forward spectrum before the division (debug), or null
};

// epilogue arithmetic shared by the tensor-core kernel and the emulator kernel
SCB_D float tc_epilogue_value(const TcPassParams& p, float acc, int c, int r, int k0) {
    float v = acc * p.scale;
    if (p.f_line) {
        if (p.lowspec && r < p.lowk_line && k0 < p.lowk_out) v = __ldg(p.lowspec + ((size_t)c * p.lowk_line + r) * p.lowk_out + k0);
        if (p.spec_dump) p.spec_dump[((size_t)c * p.lpc + r) * p.tab.n + k0] = v;
        // OpenCV: res /= (filter_X[i] + filter_Y[j] - 4), left to right in float32
        v = __fdiv_rn(v, __fsub_rn(__fadd_rn(__ldg(p.f_line + r), __ldg(p.f_out + k0)), 4.0f));
    }
    return v;
}

SCB_D void tc_store(const TcPassParams& p, int c, int r, int k0, float v) {
    if (p.transposed)
        p.out[(size_t)c * p.out_plane + (size_t)k0 * p.out_pitch + r] = v;
    else
        p.out[(size_t)c * p.out_plane + (size_t)r * p.out_pitch + k0] = v;
}

// folded input element j of a line for parity par
SCB_D float tc_fold(const float* x, int n, int par, int j) {
    const int h = n >> 1;
    if (j < h) {
        const float a = x[j], b = x[n - 1 - j];
        return par ? a - b : a + b;
    }
    if (j == h && (n & 1) && par == 0) return x[h];
    return 0.f;
}

#ifdef SCB_EMU
// CI stand-in for the tensor-core kernel: same parameters, same tables, same epilogue, the contraction as a
// plain loop with the 3xTF32 split.  Checks the host logic and the table/epilogue indexing in the GPU-less
// container; the tcgen05/TMA plumbing itself can only be checked on a B200 (tests -m gpu).
__global__ void tc_pass_kernel(TcPassParams p) {
    const int line = p.line0 + blockIdx.x * kTcM + (threadIdx.x % kTcM);
    if (line >= p.line_end || threadIdx.x >= kTcM) return;
    const int par = blockIdx.y / p.tab.ntiles, tile = blockIdx.y % p.tab.ntiles;
    const int c = line / p.lpc, r = line % p.lpc;
    const float* x = p.in + (size_t)c * p.in_plane + (size_t)r * p.in_pitch;
    const float* bh = p.tab.basis + (size_t)(2 * par + 0) * p.tab.rows * p.tab.kpad;
    const float* bl = p.tab.basis + (size_t)(2 * par + 1) * p.tab.rows * p.tab.kpad;
    for (int i = 0; i < p.tab.nt; ++i) {
        const int ki = tile * p.tab.nt + i;
        if (ki >= p.tab.nout[par]) break;
        float acc = 0.f;
        for (int j = 0; j < p.tab.kpar[par]; ++j) {
            const float f = tc_fold(x, p.tab.n, par, j);
            const float fh = tf32_round(f), fl = tf32_round(f - fh);
            const float sh = bh[(size_t)ki * p.tab.kpad + j], sl = bl[(size_t)ki * p.tab.kpad + j];
            acc += fh * sh + (fl * sh + fh * sl);
        }
        const int k0 = 2 * ki + par;
        tc_store(p, c, r, k0, tc_epilogue_value(p, acc, c, r, k0));
    }
}
#else  // ------------------------------ the real thing: sm_100a ------------------------------------

SCB_D unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

SCB_D void mbar_init(unsigned bar, unsigned count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
SCB_D void mbar_arrive(unsigned bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
SCB_D void mbar_arrive_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
SCB_D void mbar_wait(unsigned bar, unsigned parity) {
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
SCB_D void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
SCB_D void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
SCB_D void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

SCB_D void tma_load_2d(unsigned dst, const CUtensorMap* map, int x, int y, unsigned bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst), "l"(map), "r"(x),
                 "r"(y), "r"(bar)
                 : "memory");
}

// K-major operand tile, 128-byte rows, 128-byte swizzle: 8-row atoms of 1024 bytes, SBO = 1024, LBO unused.
SCB_D unsigned long long umma_desc_sw128(unsigned smem_addr) {
    unsigned long long d = 0;
    d |= (unsigned long long)((smem_addr >> 4) & 0x3fff);  // start address
    d |= (unsigned long long)0 << 16;                      // leading byte offset (ignored for swizzled K-major)
    d |= (unsigned long long)(1024 >> 4) << 32;            // stride byte offset between 8-row atoms
    d |= (unsigned long long)1 << 46;                      // descriptor version (sm_100)
    d |= (unsigned long long)2 << 61;                      // SWIZZLE_128B
    return d;
}
// kind::tf32, FP32 accumulate, A and B K-major, M = 128, N = nt
SCB_D unsigned umma_idesc_tf32(int nt) {
    unsigned d = 0;
    d |= 1u << 4;                    // c_format = F32
    d |= 2u << 7;                    // a_format = TF32
    d |= 2u << 10;                   // b_format = TF32
    d |= (unsigned)(nt >> 3) << 17;  // n_dim
    d |= (unsigned)(kTcM >> 4) << 24;  // m_dim
    return d;
}
SCB_D void umma_tf32(unsigned tmem_d, unsigned long long adesc, unsigned long long bdesc, unsigned idesc, unsigned accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
SCB_D void umma_commit(unsigned bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
SCB_D void tmem_ld16(unsigned taddr, float (&v)[16]) {
    unsigned r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
          "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    SCB_UNROLL
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

struct TcSmem {
    unsigned char* base;  // 1024-aligned
    SCB_D unsigned char* a_hi(int s) const { return base + (size_t)s * kTcStageBytes; }
    SCB_D unsigned char* a_lo(int s) const { return a_hi(s) + kTcStageA; }
    SCB_D unsigned char* b_hi(int s) const { return a_lo(s) + kTcStageA; }
    SCB_D unsigned char* b_lo(int s) const { return b_hi(s) + kTcStageB; }
    SCB_D unsigned long long* bars() const { return reinterpret_cast<unsigned long long*>(base + (size_t)kTcStages * kTcStageBytes); }
};

// grid = (line tiles, 2 * ntiles): blockIdx.y / ntiles = parity, % ntiles = N tile
__global__ void __launch_bounds__(kTcThreads, 1) tc_pass_kernel(const __grid_constant__ CUtensorMap bmap, TcPassParams p) {
    extern __shared__ unsigned char tc_smem_raw[];
    TcSmem sm;
    sm.base = reinterpret_cast<unsigned char*>(((size_t)tc_smem_raw + 1023) & ~(size_t)1023);
    unsigned long long* bars = sm.bars();
    // barrier slots: [0,S) full_A, [S,2S) full_B, [2S,3S) empty, [3S] accumulator ready; then the TMEM base address
    const unsigned bar0 = smem_u32(bars);
    auto full_a = [&](int s) { return bar0 + 8u * s; };
    auto full_b = [&](int s) { return bar0 + 8u * (kTcStages + s); };
    auto empty = [&](int s) { return bar0 + 8u * (2 * kTcStages + s); };
    const unsigned acc_bar = bar0 + 8u * (3 * kTcStages);
    unsigned* tmem_slot = reinterpret_cast<unsigned*>(bars + 3 * kTcStages + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int par = blockIdx.y / p.tab.ntiles, tile = blockIdx.y % p.tab.ntiles;
    const int nt = p.tab.nt;
    const int n0 = tile * nt;
    const int kpar = p.tab.kpar[par];
    const int num_kb = (kpar + kTcKB - 1) / kTcKB;
    const int l0 = p.line0 + blockIdx.x * kTcM;

    if (warp == 1 && lane == 0) {
        for (int s = 0; s < kTcStages; ++s) {
            mbar_init(full_a(s), 4 * 32);  // every producer thread arrives
            mbar_init(full_b(s), 1);       // the TMA thread's arrive.expect_tx
            mbar_init(empty(s), 1);        // tcgen05.commit
        }
        mbar_init(acc_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {  // TMEM: 256 columns x 128 lanes of FP32 accumulators
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(256) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const unsigned tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer: the two basis tiles (hi, lo) of every k-block =====
        if (lane == 0) {
            const int row_hi = (2 * par + 0) * p.tab.rows + n0, row_lo = (2 * par + 1) * p.tab.rows + n0;
            const unsigned bytes = 2u * (unsigned)nt * 128u;
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % kTcStages;
                mbar_wait(empty(s), ((kb / kTcStages) & 1) ^ 1);
                mbar_arrive_expect_tx(full_b(s), bytes);
                tma_load_2d(smem_u32(sm.b_hi(s)), &bmap, kb * kTcKB, row_hi, full_b(s));
                tma_load_2d(smem_u32(sm.b_lo(s)), &bmap, kb * kTcKB, row_lo, full_b(s));
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            const unsigned idesc = umma_idesc_tf32(nt);
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % kTcStages;
                const unsigned ph = (kb / kTcStages) & 1;
                mbar_wait(full_a(s), ph);
                mbar_wait(full_b(s), ph);
                tc_fence_after();
                const unsigned long long ah = umma_desc_sw128(smem_u32(sm.a_hi(s))), al = umma_desc_sw128(smem_u32(sm.a_lo(s)));
                const unsigned long long bh = umma_desc_sw128(smem_u32(sm.b_hi(s))), bl = umma_desc_sw128(smem_u32(sm.b_lo(s)));
                SCB_UNROLL
                for (int ks = 0; ks < kTcKB / 8; ++ks) {  // UMMA K = 8 tf32 = 32 bytes: advance the start address by 2 (16-byte units)
                    const unsigned long long o = (unsigned long long)(ks * 2);
                    umma_tf32(tmem_base, al + o, bh + o, idesc, (kb | ks) != 0);  // small terms first
                    umma_tf32(tmem_base, ah + o, bl + o, idesc, 1);
                    umma_tf32(tmem_base, ah + o, bh + o, idesc, 1);
                }
                umma_commit(empty(s));  // frees the stage when these MMAs have read it
            }
            umma_commit(acc_bar);
        }
    } else {
        // ===== A producers (4 warps): global -> fold -> hi/lo -> swizzled shared memory =====
        const int t = threadIdx.x - 64;       // 0..127
        const int chunk = t & 7;              // 16-byte chunk of the 128-byte row
        const int n = p.tab.n, h = n >> 1;
        const float* lp[8];
        SCB_UNROLL
        for (int it = 0; it < 8; ++it) {
            const int line = l0 + (t >> 3) + 16 * it;
            lp[it] = nullptr;
            if (line < p.line_end) {
                const int c = line / p.lpc, r = line - c * p.lpc;
                lp[it] = p.in + (size_t)c * p.in_plane + (size_t)r * p.in_pitch;
            }
        }
        auto fetch = [&](int kb, float (&f)[8][4]) {
            const int j0 = kb * kTcKB + 4 * chunk;
            SCB_UNROLL
            for (int it = 0; it < 8; ++it) {
                float a[4] = {0.f, 0.f, 0.f, 0.f}, b[4] = {0.f, 0.f, 0.f, 0.f};
                const float* x = lp[it];
                if (x && j0 < kpar) {
                    if (j0 + 3 < h) {  // four full pairs: one aligned 16-byte load + four reversed scalars
                        const float4 v = __ldg(reinterpret_cast<const float4*>(x + j0));
                        a[0] = v.x; a[1] = v.y; a[2] = v.z; a[3] = v.w;
                        SCB_UNROLL
                        for (int e = 0; e < 4; ++e) b[e] = __ldg(x + (n - 1 - j0 - e));
                    } else {
                        SCB_UNROLL
                        for (int e = 0; e < 4; ++e) {
                            const int j = j0 + e;
                            if (j < h) {
                                a[e] = __ldg(x + j);
                                b[e] = __ldg(x + (n - 1 - j));
                            } else if (j == h && (n & 1) && par == 0) {
                                a[e] = __ldg(x + h);
                            }
                        }
                    }
                }
                SCB_UNROLL
                for (int e = 0; e < 4; ++e) f[it][e] = par ? a[e] - b[e] : a[e] + b[e];
            }
        };
        float cur[8][4], nxt[8][4];
        fetch(0, cur);
        for (int kb = 0; kb < num_kb; ++kb) {
            const int s = kb % kTcStages;
            if (kb + 1 < num_kb) fetch(kb + 1, nxt);  // in flight while we wait for the slot
            mbar_wait(empty(s), ((kb / kTcStages) & 1) ^ 1);
            unsigned char* ahi = sm.a_hi(s);
            unsigned char* alo = sm.a_lo(s);
            SCB_UNROLL
            for (int it = 0; it < 8; ++it) {
                const int row = (t >> 3) + 16 * it;
                const unsigned off = (unsigned)row * 128u + (unsigned)((chunk ^ (row & 7)) << 4);
                float hi[4], lo[4];
                SCB_UNROLL
                for (int e = 0; e < 4; ++e) {
                    hi[e] = tf32_round(cur[it][e]);
                    lo[e] = tf32_round(cur[it][e] - hi[e]);
                }
                *reinterpret_cast<float4*>(ahi + off) = make_float4(hi[0], hi[1], hi[2], hi[3]);
                *reinterpret_cast<float4*>(alo + off) = make_float4(lo[0], lo[1], lo[2], lo[3]);
            }
            fence_proxy_async();  // generic-proxy stores -> visible to the tensor core's async-proxy reads
            mbar_arrive(full_a(s));
            if (kb + 1 < num_kb) {
                SCB_UNROLL
                for (int it = 0; it < 8; ++it) {
                    SCB_UNROLL
                    for (int e = 0; e < 4; ++e) cur[it][e] = nxt[it][e];
                }
            }
        }
        // ===== epilogue: warp w owns TMEM lanes 32*(w%4) .. +31 =====
        mbar_wait(acc_bar, 0);
        tc_fence_after();
        const int q = warp & 3;
        const int line = l0 + 32 * q + lane;
        const bool line_ok = line < p.line_end;
        const int c = line_ok ? line / p.lpc : 0, r = line_ok ? line - c * p.lpc : 0;
        const int nout = p.tab.nout[par];
        for (int cc = 0; cc < nt; cc += 16) {
            float v[16];
            tmem_ld16(tmem_base + ((unsigned)(32 * q) << 16) + (unsigned)cc, v);
            if (line_ok) {
                SCB_UNROLL
                for (int i = 0; i < 16; ++i) {
                    const int ki = n0 + cc + i;
                    if (ki < nout) {
                        const int k0 = 2 * ki + par;
                        tc_store(p, c, r, k0, tc_epilogue_value(p, v[i], c, r, k0));
                    }
                }
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256) : "memory");
    }
}
#endif  // SCB_EMU

// ---------------------------------------------------------------------------------------------
// compose: planar float solved field -> interleaved u8 (OpenCV solve() epilogue + merge; reference
// post_processing, imp.cpp:2078-2103).  A thread packs 4 pixels = 12 bytes.
// ---------------------------------------------------------------------------------------------
struct TcComposeParams {
    const float* u;  // [3][ny][pitch]
    long long plane;
    int pitch, nx, ny;
    unsigned char* out;  // interior origin pixel
    long long out_pitch;
    float* u_dump;  // [3][ny][nx] or null
};

__global__ void __launch_bounds__(128) tc_compose_kernel(TcComposeParams p) {
    const int y = blockIdx.y, x0 = 4 * (blockIdx.x * 128 + threadIdx.x);
    if (x0 >= p.nx) return;
    unsigned char* o = p.out + (long long)y * p.out_pitch + 3 * x0;
    for (int k = 0; k < 4 && x0 + k < p.nx; ++k) {
        for (int c = 0; c < 3; ++c) {
            const float v = __ldg(p.u + (size_t)c * p.plane + (size_t)y * p.pitch + x0 + k);
            if (p.u_dump) p.u_dump[((size_t)c * p.ny + y) * p.nx + x0 + k] = v;
            // v < 0 -> 0, v > 255 -> 255, else truncate toward zero
            o[3 * k + c] = v < 0.f ? (unsigned char)0 : (v > 255.f ? (unsigned char)255 : (unsigned char)__float2int_rz(v));
        }
    }
}

}  // namespace scb
