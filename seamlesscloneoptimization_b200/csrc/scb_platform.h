// Platform glue: the kernels and the host driver are written once and compile two ways.
//
//   nvcc (product)      : real CUDA for sm_100a.  This is the only thing the package ever loads.
//   g++ -DSCB_EMU (CI)  : tests/emu/emu_cuda.h supplies a fiber-based SIMT interpreter (one fiber
//                         per CUDA thread, __syncthreads/shuffles as fiber barriers) and a stub of
//                         the few runtime calls the driver uses.  It exists so that the indexing
//                         and host logic can be checked in the GPU-less authoring container; it is
//                         built into tests/emu/_build/ only and is never importable from the package.
#pragma once

#include <cstddef>
#include <cstdint>

#ifdef SCB_EMU
#include "emu_cuda.h"
#define SCB_LAUNCH(kernel, grid, block, smem, stream, ...) \
    ::emu::launch(kernel, grid, block, smem, __VA_ARGS__)
#define SCB_DYN_SMEM(type, name) type* name = reinterpret_cast<type*>(::emu::dyn_smem())
#define SCB_UNROLL
#else
#include <cuda_runtime.h>
#define SCB_LAUNCH(kernel, grid, block, smem, stream, ...) \
    kernel<<<grid, block, smem, stream>>>(__VA_ARGS__)
#define SCB_DYN_SMEM(type, name)                                      \
    extern __shared__ __align__(16) unsigned char name##_raw_[];     \
    type* name = reinterpret_cast<type*>(name##_raw_)
#define SCB_UNROLL _Pragma("unroll")
#endif

#define SCB_D __device__ __forceinline__
#define SCB_HD __host__ __device__ __forceinline__

// Packed fp32x2 arithmetic: FADD2 / FMUL2 / FFMA2 on sm_100a (one instruction, two lanes, operand
// negation and broadcast immediates are free).  The emulator spells them out per lane.
namespace scb {
#ifdef SCB_EMU
SCB_D float2 f2add(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
SCB_D float2 f2sub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
SCB_D float2 f2mul(float2 a, float2 b) { return make_float2(a.x * b.x, a.y * b.y); }
SCB_D float2 f2fma(float2 a, float2 b, float2 c) { return make_float2(a.x * b.x + c.x, a.y * b.y + c.y); }
#else
SCB_D float2 f2add(float2 a, float2 b) { return __fadd2_rn(a, b); }
SCB_D float2 f2sub(float2 a, float2 b) { return __fadd2_rn(a, make_float2(-b.x, -b.y)); }
SCB_D float2 f2mul(float2 a, float2 b) { return __fmul2_rn(a, b); }
SCB_D float2 f2fma(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
#endif
// Balanced base-256 digits of four integers at once, transposed: word i holds digit i (most significant first, DA digits) of
// v[0..3] in its four bytes -- the layout of a digit plane.  v + 0x80808080 has the plain unsigned bytes d_k + 128 (the per-byte
// bias absorbs the borrows of the balanced representation), so XOR 0x80 per byte gives the int8 digits; six byte permutes transpose.
// Requires |v| < 2^(8 DA - 1) (the top digit takes what is left).
template <int DA>
SCB_D void balanced_digits4(const int (&v)[4], unsigned (&w)[DA]) {
    unsigned u[4];
    SCB_UNROLL
    for (int e = 0; e < 4; ++e) u[e] = ((unsigned)v[e] + 0x80808080u) ^ 0x80808080u;  // byte k of u[e] = digit of weight 256^k
    const unsigned lo01 = __byte_perm(u[0], u[1], 0x5140u), lo23 = __byte_perm(u[2], u[3], 0x5140u);
    const unsigned k0 = __byte_perm(lo01, lo23, 0x5410u), k1 = __byte_perm(lo01, lo23, 0x7632u);
    if (DA == 2) {
        w[1] = k0;
        w[0] = k1;
    } else {
        const unsigned hi01 = __byte_perm(u[0], u[1], 0x7362u), hi23 = __byte_perm(u[2], u[3], 0x7362u);
        w[DA - 1] = k0;
        w[DA - 2] = k1;
        w[DA - 3] = __byte_perm(hi01, hi23, 0x5410u);
        w[DA - 4] = __byte_perm(hi01, hi23, 0x7632u);
    }
}

SCB_D float2 f2neg(float2 a) { return make_float2(-a.x, -a.y); }
SCB_D float2 f2dup(float s) { return make_float2(s, s); }
}  // namespace scb
