// scb_i8.h -- the DST-I along x as an EXACT integer contraction on the INT8 tensor cores (tcgen05.mma.kind::i8).
//
// OpenCV's Cloning::dst is a multiplication by the sine matrix S[j][k] = sin(pi (j+1)(k+1) / N) (the reference's GEMM
// flavour of the solver: /root/reference/seamlessClone-CUDA/seamlessClone_imp.cpp:488-665, 1118-1334, cublasSgemmBatched
// in FP32).  A floating-point tensor-core contraction cannot replace it here: tcgen05 accumulates in FP32 with truncation
// at every MMA step, ~1e-5 relative after K ~ 900, which costs 0.2-0.3 % of exactly matching bytes (scb_tc.cuh, measured).
// Integers do not round.  Both operands are written in balanced base-256 digits (int8, [-128, 127]):
//
//     x[j] * 2^F  =  sum_i a_i[j] 256^(DA-1-i)          (DA digits: 2 for the integer right-hand side, 4 otherwise)
//     S[j][k] * 2^30 ~ sum_d s_d[j][k] 256^(3-d)        (DB digit planes of the basis, built once per length in float64)
//
// and every digit-plane product a_i x s_d is one s8 x s8 -> s32 MMA whose accumulation is exact.  Products of equal
// weight 256^(..-c), c = i + d ("class" c), share one TMEM accumulator; classes 0..3 are kept (the dropped ones are below
// 2^-31 of full scale).  The digit planes of the basis sit side by side along the MMA N dimension, so ONE instruction
// multiplies a digit of the lines by several planes at once and the class shift is a TMEM column offset:
//
//     D[:, 64 i : 64 (i + cnt_i)]  +=  A_i (128 lines x 32 k)  x  [ s_0 | s_1 | .. | s_(cnt_i - 1) ] (64 outputs each)
//
// The epilogue recombines the four int32 class sums per output into one float (or, for the lowest frequencies, an
// exact float64 row sum -- which replaces the FP64 lowfreq_rows_kernel of the FFT engine).
//
// Also used: even/odd folding (sin(pi (N-j) k / N) = (-1)^(k+1) sin(pi j k / N)), which halves K and N (scb_tc.cuh).
//
// Kernels (scb_i8.cu):
//   i8_basis_kernel      plan time, per line length n: the digit planes of the folded basis
//   i8_digitize_kernel   float lines -> folded digit planes (per-line power-of-two scale for the inverse pass)
//   i8_gemm_kernel       TMA-fed, warp-specialised tcgen05 kernel: TMEM accumulators 512 columns, epilogue through smem
#pragma once

#include <cstddef>
#include <cstdint>

#include "scb_platform.h"

namespace scb {

static constexpr int kI8M = 128;        // lines per CTA tile (UMMA M)
static constexpr int kI8P = 64;         // outputs per digit plane (TMEM columns per class)
static constexpr int kI8KB = 128;       // K elements (= bytes) per stage row: one 128-byte swizzle atom row
static constexpr int kI8Classes = 4;    // weight classes kept
static constexpr int kI8BasisDigits = 4;
static constexpr int kI8BasisBits = 30; // basis scale 2^30: |digits| <= 64 at the top, no overflow at sin = +-1
static constexpr int kI8MinN = 64;      // shorter lines stay on the FFT engine
static constexpr int kI8MaxN = 8192;
static constexpr int kI8LowK = 4;       // exact float64 row sums for the lowest frequencies (= kLowK of the FFT engine)

// Folded geometry of one line length (same convention as scb_tc.cuh): parity p = 0 pairs x[j] + x[n-1-j] with the outputs
// k0 = 2 ki, p = 1 pairs x[j] - x[n-1-j] with k0 = 2 ki + 1.
struct I8Geom {
    int n = 0;
    int kpar[2] = {0, 0};   // folded K per parity
    int nout[2] = {0, 0};   // outputs per parity
    int kpad = 0;           // row pitch of the digit planes in bytes, multiple of 128
    int nsb = 0;            // 64-output sub-blocks per parity (even)
};

SCB_HD I8Geom i8_geometry(int n) {
    I8Geom g;
    const int h = n / 2;
    g.n = n;
    g.kpar[0] = h + (n & 1);
    g.kpar[1] = h;
    g.nout[0] = (n + 1) / 2;
    g.nout[1] = n / 2;
    g.kpad = (g.kpar[0] + kI8KB - 1) / kI8KB * kI8KB;
    g.nsb = (g.nout[0] + kI8P - 1) / kI8P;
    g.nsb += g.nsb & 1;
    return g;
}
// basis table: rows [parity][sub-block][digit 0..3][64], kpad bytes each
SCB_HD size_t i8_basis_rows(const I8Geom& g) { return (size_t)2 * g.nsb * kI8BasisDigits * kI8P; }
SCB_HD size_t i8_basis_bytes(const I8Geom& g) { return i8_basis_rows(g) * g.kpad; }
// digit planes of the lines: rows [parity][digit][m_rows], kpad bytes each
SCB_HD int i8_m_rows(int lines) { return (lines + kI8M - 1) / kI8M * kI8M; }
SCB_HD size_t i8_adig_bytes(const I8Geom& g, int lines, int da) { return (size_t)2 * da * i8_m_rows(lines) * g.kpad; }

struct I8DigitizeParams {
    I8Geom g;
    const float* in;        // line (c, r) at in + c*in_plane + r*in_pitch, n floats; pitch a multiple of 4
    long long in_plane;
    int in_pitch;
    int lpc;                // lines per channel
    int lines;              // 3 * lpc; line = 3 * row + channel (channel-interleaved, so a band of image rows is a contiguous line range)
    int m_rows;             // i8_m_rows(lines)
    int line0, line1;       // lines of this launch (HOST calls digitise band by band); [lines, m_rows) are written as zeros
    signed char* a;         // digit planes
    float* lscale;          // [m_rows] per-line scale the epilogue multiplies by
    float fixed_scale;      // per_line == 0: x is multiplied by this before rounding (1: integer input, 65536: 16 fractional bits)
    int per_line;           // 1: per-line power-of-two scale from the line's largest magnitude (inverse pass)
};

struct I8GemmParams {
    I8Geom g;
    int lines, lpc, m_rows;
    int mt0, mt1;              // 128-line tiles of this launch
    int line0, line1;          // only lines in [line0, line1) are stored (a row shard's first / last tile straddles its neighbours' lines)
    const signed char* a;      // digit planes of the lines          (emulator build reads them directly; the GPU kernel goes through TMA)
    const signed char* basis;  // digit planes of the basis
    const float* lscale;       // [m_rows]
    float scale;               // out = t * scale * lscale[line],  t = sum of the kept classes in units of 256^(DA-1) / 2^30 ... folded into scale by the host
    float* out;                // out[c*out_plane + r*out_pitch + k0]
    long long out_plane;
    int out_pitch;
    double* R;                 // [3][lowk][lpc] exact row sums sum_j x[j] sin(..) for k0 < lowk (forward pass), or null
    int lowk;
    double rscale;             // R = t * rscale * lscale[line]
    unsigned char* out_u8;     // non-null: the pass is the LAST one -- its epilogue clamps, truncates and stores the byte of channel c at
    long long out_u8_pitch;    //   out_u8[r * out_u8_pitch + 3 * k0 + c] (OpenCV solve() epilogue + merge) instead of the float into `out`
    long long* trace;          // tuning aid (SCB_I8_TRACE=1): 8 clock64() stamps per CTA, or null
};

struct I8ComposeParams {
    const float* u;       // [3][ny][pitch] solved field
    long long plane;
    int pitch, nx;
    unsigned char* out;   // interior origin pixel of the output rows (any alignment)
    long long out_pitch;
    int y0;               // first row of this launch
};

// host launchers (scb_i8.cu).  `da` = digits of the lines: 2 (integer right-hand side, |x| <= 4095 after folding) or 4.
// `db` = basis digits used: 4 (forward) or 3 (inverse).  Supported (da, db): (2,4), (4,4), (4,3).
// Return 0 or a cudaError_t value.
int i8_configure();  // cudaFuncSetAttribute of the kernels, once per device context
int i8_launch_basis(void* stream, const I8Geom& g, signed char* basis);
int i8_launch_digitize(void* stream, const I8DigitizeParams& p, int da);
const char* i8_variant_string();  // the kernel switches in force ("i8_persistent=.. i8_kb=..")
int i8_launch_gemm(void* stream, const I8GemmParams& p, int da, int db);
int i8_launch_compose(void* stream, const I8ComposeParams& p, int rows);

}  // namespace scb
