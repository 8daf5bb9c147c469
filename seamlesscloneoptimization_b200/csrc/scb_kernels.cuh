// scb_kernels.cuh -- device code of the NORMAL_CLONE hot path.
//
// Data layout in HBM (one clone job; ROI = w x h, unknowns nx = w-2, ny = h-2):
//   D, S      u8 interleaved BGR rows (caller's dst at the ROI origin / src at the mask-bbox origin,
//             or pitched staging copies of exactly those rectangles when the caller's images are on the host)
//   E         u8 h x w, mask after ring-zero + 3x erosion, cropped to the ROI        (plan-time)
//   At        f32 [3][nx][ny]   row-transformed RHS, stored TRANSPOSED so that the column pass reads lines
//   Ct        f32 [3][ny][nx]   after the column pass (forward, / eigenvalues, inverse), transposed back
//   out       u8 interleaved, the ROI interior of `blend`
// Three line passes, each one CTA per line and all three channels of the line together:
//   rows_fwd : stencil (gradients (+) mask blend (+) divergence (+) Dirichlet injection) -> DST-I along x
//   cols     : DST-I along y -> / (fx + fy - 4) -> inverse DST-I along y
//   rows_inv : inverse DST-I along x -> clamp + truncate -> u8
// so a pixel makes three HBM round trips (7+12 | 12+12 | 12+3 = 58 B per solved RGB pixel).
// The default engine (SCB_ENGINE_TRI, scb_tri.cuh) keeps rows_fwd / rows_inv and replaces `cols` by a tridiagonal solve of every
// spectral column on the NATURAL layout A [3][ny][nx] -> Ct [3][ny][nx]; the kernels below then serve the x axis only, and
// cols_kernel / lowfreq_cols_kernel the FFT-on-both-axes engine and its row/column-sharded entry points.
//
// What each piece replaces in the reference (/root/reference/seamlessClone-CUDA/seamlessClone_imp.cpp):
//   rhs_pixel            pre_process_kernel_gradient :1920-1964 + pre_process_kernel_lapXY :1966-2018
//   rows_fwd/cols/rows_inv   dst() :1694-1811 (dft_kernel_0/1, cufftExecC2C x2, dft_copy_kernel_transpose),
//                        updateUij_kernel_fft :1642-1669, scale_kernel_fft_transpose :1672-1691
//   compose in rows_inv  post_processing :2078-2103
//   mask_* kernels       setMaskBoundaryToConstant :967-976, calBoundingBox :927-963, myErode :892-925
// Arithmetic follows OpenCV (the parity target), not the reference where the two differ.
#pragma once

#include "scb_fft.cuh"
#include "scb_platform.h"

namespace scb {

struct LenTabDev {
    int n, log2m, lowk;
    const float2* chirp;
    const float2* bhat_t;
    const float2* bhat_q;  // quad mode (two real lines per sequence): spectrum of conj chirp over [-2n, n-1]; null when 3n > M
    const float2* tw;
    const float4* gtw;  // per-pass twiddle tables of the group engine (scb_gfft.cuh)
    const double* sinlow;
};

struct StencilSrc {
    const unsigned char* D;  // dst pixel (0,0) of the ROI
    long long d_pitch;
    const unsigned char* S;  // src pixel at the mask-bbox origin
    long long s_pitch;
    const unsigned char* E;  // eroded mask, ROI sized
    long long e_pitch;
    int w, h;
};

// RHS of the Poisson system at interior pixel (x, y), 0 <= x < w-2, 0 <= y < h-2, all 3 channels.
//   v   = grad(D) * (255-E)/255 + grad(S) * E/255      two rounded products, one rounded sum -- never an FMA
//   lap = (vx[X] - vx[X-1]) + (vy[Y] - vy[Y-1])
//   g   = lap - (dst pixels just outside the interior)  == OpenCV's  lap - Laplacian(bound)
SCB_D void rhs_pixel(const StencilSrc& s, int x, int y, float g[3]) {
    const int X = x + 1, Y = y + 1;
    const unsigned char* d1 = s.D + (long long)Y * s.d_pitch + 3 * (X - 1);
    const unsigned char* d0 = d1 - s.d_pitch + 3;
    const unsigned char* d2 = d1 + s.d_pitch + 3;
    const unsigned char* s1 = s.S + (long long)Y * s.s_pitch + 3 * (X - 1);
    const unsigned char* s0 = s1 - s.s_pitch + 3;
    const unsigned char* s2 = s1 + s.s_pitch + 3;
    const unsigned char* e1 = s.E + (long long)Y * s.e_pitch + X;
    const float inv255 = 1.0f / 255.0f;
    const int ec = __ldg(e1), el = __ldg(e1 - 1), eu = __ldg(e1 - s.e_pitch);
    const float mc = __fmul_rn((float)ec, inv255), mic = __fmul_rn((float)(255 - ec), inv255);
    const float ml = __fmul_rn((float)el, inv255), mil = __fmul_rn((float)(255 - el), inv255);
    const float mu = __fmul_rn((float)eu, inv255), miu = __fmul_rn((float)(255 - eu), inv255);
    SCB_UNROLL
    for (int c = 0; c < 3; ++c) {
        const float Dl = (float)__ldg(d1 + c), Dc = (float)__ldg(d1 + 3 + c), Dr = (float)__ldg(d1 + 6 + c);
        const float Du = (float)__ldg(d0 + c), Dd = (float)__ldg(d2 + c);
        const float Sl = (float)__ldg(s1 + c), Sc = (float)__ldg(s1 + 3 + c), Sr = (float)__ldg(s1 + 6 + c);
        const float Su = (float)__ldg(s0 + c), Sd = (float)__ldg(s2 + c);
        const float vxc = __fadd_rn(__fmul_rn(Dr - Dc, mic), __fmul_rn(Sr - Sc, mc));
        const float vxl = __fadd_rn(__fmul_rn(Dc - Dl, mil), __fmul_rn(Sc - Sl, ml));
        const float vyc = __fadd_rn(__fmul_rn(Dd - Dc, mic), __fmul_rn(Sd - Sc, mc));
        const float vyu = __fadd_rn(__fmul_rn(Dc - Du, miu), __fmul_rn(Sc - Su, mu));
        const float lap = __fadd_rn(__fsub_rn(vxc, vxl), __fsub_rn(vyc, vyu));
        float bnd = 0.f;
        if (X == 1) bnd += Dl;
        if (X == s.w - 2) bnd += Dr;
        if (Y == 1) bnd += Du;
        if (Y == s.h - 2) bnd += Dd;
        g[c] = __fsub_rn(lap, bnd);
    }
}

// ---------------------------------------------------------------------------------------------
// Stand-alone fused stencil: u8 dst/src/mask in, float RHS out, one HBM round trip per pixel
// (7 B read + 12 B written).  Replaces pre_process_kernel_gradient + pre_process_kernel_lapXY
// (imp.cpp:1920-2018), which move 24 + 24 B per pixel of float gradients through HBM in between.
//
// A thread owns 4 consecutive interior pixels of one row: it fetches the 3 x 20 bytes of dst, the
// 3 x 20 bytes of src and the 2 x 8 bytes of mask it needs as aligned 32-bit words (funnel-shifted to
// the pixel boundary -- rows of a caller's image start at any byte), and writes three float4.
// Where all three mask taps are 0 or 255 -- everywhere except on grey masks -- the blend is a select
// and the whole stencil runs in integer arithmetic (exact, hence bit-identical to the float path).
// ---------------------------------------------------------------------------------------------
struct RhsParams {
    StencilSrc st;
    int nx, ny;
    float* g;  // [3][ny][gp]
    int gp;    // row pitch of g in floats, multiple of 4
    int y0;    // first interior row of this launch
    int y1;          // one past the last interior row of this launch (rhs_mode_kernel)
    int transposed;  // 1: store g as [3][nx][gpt] (lines along y: the solve runs its FFT passes along y, scb_api.cu choose_swap)
    int gpt;         // line pitch of the transposed layout in floats
};

// four consecutive pixels x0..x0+3 of row y, channel c
template <bool TRANSPOSED>
SCB_D void rhs_store4(const RhsParams& p, int c, int x0, int y, float a, float b, float d, float e) {
    if (!TRANSPOSED) {
        *reinterpret_cast<float4*>(p.g + ((size_t)c * p.ny + y) * p.gp + x0) = make_float4(a, b, d, e);
    } else {
        float* o = p.g + ((size_t)c * p.nx + x0) * p.gpt + y;
        o[0] = a;
        o[(size_t)p.gpt] = b;
        o[2 * (size_t)p.gpt] = d;
        o[3 * (size_t)p.gpt] = e;
    }
}

template <int N>
SCB_D void load_unaligned_words(const unsigned char* p, unsigned (&out)[N]) {
    const unsigned a = (unsigned)((size_t)p & 3);
    const unsigned* q = reinterpret_cast<const unsigned*>(p - a);
    unsigned w[N + 1];
    SCB_UNROLL
    for (int k = 0; k < N; ++k) w[k] = __ldg(q + k);
    w[N] = a ? __ldg(q + N) : 0u;  // an unaligned span of 4N bytes touches N+1 words; never reads a word without a needed byte
    SCB_UNROLL
    for (int k = 0; k < N; ++k) out[k] = __funnelshift_r(w[k], w[k + 1], 8 * a);
}
template <int N>
SCB_D int byte_of(const unsigned (&w)[N], int i) { return (int)((w[i >> 2] >> (8 * (i & 3))) & 255u); }

static constexpr int kRhsThreads = 128;

template <bool TRANSPOSED>
__global__ void __launch_bounds__(kRhsThreads) rhs_kernel(RhsParams p) {
    const int y = p.y0 + blockIdx.y;
    const int x0 = 4 * (blockIdx.x * kRhsThreads + threadIdx.x);
    if (x0 >= p.nx) return;
    const StencilSrc& s = p.st;
    float* g0 = p.g + ((size_t)0 * p.ny + y) * p.gp + x0;
    const size_t plane = (size_t)p.ny * p.gp;
    if (x0 + 8 > p.nx) {  // row tail (and the pad columns up to gp): per-pixel path with byte loads
        for (int k = 0; k < 4; ++k) {
            float g[3] = {0.f, 0.f, 0.f};
            if (x0 + k < p.nx) rhs_pixel(s, x0 + k, y, g);
            if (TRANSPOSED) {
                if (x0 + k < p.nx)
                    for (int c = 0; c < 3; ++c) p.g[((size_t)c * p.nx + x0 + k) * p.gpt + y] = g[c];
            } else if (x0 + k < p.gp) {
                g0[k] = g[0];
                g0[plane + k] = g[1];
                g0[2 * plane + k] = g[2];
            }
        }
        return;
    }
    const int X = x0 + 1, Y = y + 1;
    // row Y from column X-1 (20 bytes cover X-1 .. X+4), rows Y-1 / Y+1 from column X (12 bytes cover X .. X+3)
    unsigned dm[5], du[3], dd[3], sm[5], su[3], sd[3], em[2], eu[1];
    load_unaligned_words<2>(s.E + (long long)Y * s.e_pitch + (X - 1), em);
    load_unaligned_words<1>(s.E + (long long)(Y - 1) * s.e_pitch + X, eu);
    {
        // Fast path (almost every thread): the nine mask taps of the four pixels are all 0 or all 255 and no
        // pixel touches the ROI border, so g is the 5-point Laplacian of ONE image, a 1-D stencil over its
        // interleaved bytes (taps at -3, +3, -pitch, +pitch bytes).  Only that image is read, and the twelve
        // values are formed two at a time in packed 16-bit lanes (biased by 2048 so that no borrow crosses a
        // lane) and converted with the 2^23 trick: ~35 instructions per pixel instead of ~140, which is what
        // lets the kernel run at HBM speed.  The results are small exact integers, hence bit-identical to
        // the float path below.
        const unsigned m_and = em[0] & eu[0] & (em[1] | 0xffffff00u), m_or = em[0] | eu[0] | (em[1] & 0xffu);
        const bool all_src = (m_and == 0xffffffffu), all_dst = (m_or == 0u);
        if ((all_src || all_dst) && x0 > 0 && Y > 1 && Y < s.h - 2) {
            const unsigned char* img = all_src ? s.S : s.D;
            const long long pitch = all_src ? s.s_pitch : s.d_pitch;
            unsigned m[5], u[3], d[3];
            load_unaligned_words<5>(img + (long long)Y * pitch + 3 * (X - 1), m);
            load_unaligned_words<3>(img + (long long)(Y - 1) * pitch + 3 * X, u);
            load_unaligned_words<3>(img + (long long)(Y + 1) * pitch + 3 * X, d);
            float o[12];
            SCB_UNROLL
            for (int j = 0; j < 3; ++j) {
                const unsigned l = m[j];                                        // bytes e-3
                const unsigned c = __funnelshift_r(m[j], m[j + 1], 24);         // bytes e
                const unsigned r = __funnelshift_r(m[j + 1], m[j + 2], 16);     // bytes e+3
                SCB_UNROLL
                for (int half = 0; half < 2; ++half) {
                    const unsigned sel = half ? 0x4342u : 0x4140u;  // two bytes -> two 16-bit lanes
                    unsigned t = 0x08000800u + __byte_perm(l, 0u, sel) + __byte_perm(r, 0u, sel) + __byte_perm(u[j], 0u, sel) + __byte_perm(d[j], 0u, sel);
                    t -= 4u * __byte_perm(c, 0u, sel);
                    o[4 * j + 2 * half + 0] = __uint_as_float(__byte_perm(t, 0x4b000000u, 0x7610u)) - 8390656.0f;  // (2^23 + 2048 + v) - (2^23 + 2048)
                    o[4 * j + 2 * half + 1] = __uint_as_float(__byte_perm(t, 0x4b000000u, 0x7632u)) - 8390656.0f;
                }
            }
            SCB_UNROLL
            for (int c = 0; c < 3; ++c) rhs_store4<TRANSPOSED>(p, c, x0, y, o[c], o[3 + c], o[6 + c], o[9 + c]);
            return;
        }
    }
    load_unaligned_words<5>(s.D + (long long)Y * s.d_pitch + 3 * (X - 1), dm);
    load_unaligned_words<3>(s.D + (long long)(Y - 1) * s.d_pitch + 3 * X, du);
    load_unaligned_words<3>(s.D + (long long)(Y + 1) * s.d_pitch + 3 * X, dd);
    load_unaligned_words<5>(s.S + (long long)Y * s.s_pitch + 3 * (X - 1), sm);
    load_unaligned_words<3>(s.S + (long long)(Y - 1) * s.s_pitch + 3 * X, su);
    load_unaligned_words<3>(s.S + (long long)(Y + 1) * s.s_pitch + 3 * X, sd);
    float out[3][4];
    SCB_UNROLL
    for (int k = 0; k < 4; ++k) {
        const int el = byte_of(em, k), ec = byte_of(em, k + 1), eup = byte_of(eu, k);
        const bool binary = ((el == 0) | (el == 255)) & ((ec == 0) | (ec == 255)) & ((eup == 0) | (eup == 255));
        const int Xk = X + k;
        if (binary) {
            SCB_UNROLL
            for (int c = 0; c < 3; ++c) {
                const int Dl = byte_of(dm, 3 * k + c), Dc = byte_of(dm, 3 * k + 3 + c), Dr = byte_of(dm, 3 * k + 6 + c);
                const int Du = byte_of(du, 3 * k + c), Dd = byte_of(dd, 3 * k + c);
                const int Sl = byte_of(sm, 3 * k + c), Sc = byte_of(sm, 3 * k + 3 + c), Sr = byte_of(sm, 3 * k + 6 + c);
                const int Su = byte_of(su, 3 * k + c), Sd = byte_of(sd, 3 * k + c);
                int lap = (ec ? (Sr - Sc) + (Sd - Sc) : (Dr - Dc) + (Dd - Dc)) - (el ? (Sc - Sl) : (Dc - Dl)) - (eup ? (Sc - Su) : (Dc - Du));
                if (Xk == 1) lap -= Dl;
                if (Xk == s.w - 2) lap -= Dr;
                if (Y == 1) lap -= Du;
                if (Y == s.h - 2) lap -= Dd;
                out[c][k] = (float)lap;
            }
        } else {  // grey mask values: OpenCV's float arithmetic, operation by operation
            const float inv255 = 1.0f / 255.0f;
            const float mc = __fmul_rn((float)ec, inv255), mic = __fmul_rn((float)(255 - ec), inv255);
            const float ml = __fmul_rn((float)el, inv255), mil = __fmul_rn((float)(255 - el), inv255);
            const float mu = __fmul_rn((float)eup, inv255), miu = __fmul_rn((float)(255 - eup), inv255);
            SCB_UNROLL
            for (int c = 0; c < 3; ++c) {
                const float Dl = (float)byte_of(dm, 3 * k + c), Dc = (float)byte_of(dm, 3 * k + 3 + c), Dr = (float)byte_of(dm, 3 * k + 6 + c);
                const float Du = (float)byte_of(du, 3 * k + c), Dd = (float)byte_of(dd, 3 * k + c);
                const float Sl = (float)byte_of(sm, 3 * k + c), Sc = (float)byte_of(sm, 3 * k + 3 + c), Sr = (float)byte_of(sm, 3 * k + 6 + c);
                const float Su = (float)byte_of(su, 3 * k + c), Sd = (float)byte_of(sd, 3 * k + c);
                const float vxc = __fadd_rn(__fmul_rn(Dr - Dc, mic), __fmul_rn(Sr - Sc, mc));
                const float vxl = __fadd_rn(__fmul_rn(Dc - Dl, mil), __fmul_rn(Sc - Sl, ml));
                const float vyc = __fadd_rn(__fmul_rn(Dd - Dc, mic), __fmul_rn(Sd - Sc, mc));
                const float vyu = __fadd_rn(__fmul_rn(Dc - Du, miu), __fmul_rn(Sc - Su, mu));
                const float lap = __fadd_rn(__fsub_rn(vxc, vxl), __fsub_rn(vyc, vyu));
                float bnd = 0.f;
                if (Xk == 1) bnd += Dl;
                if (Xk == s.w - 2) bnd += Dr;
                if (Y == 1) bnd += Du;
                if (Y == s.h - 2) bnd += Dd;
                out[c][k] = __fsub_rn(lap, bnd);
            }
        }
    }
    SCB_UNROLL
    for (int c = 0; c < 3; ++c) rhs_store4<TRANSPOSED>(p, c, x0, y, out[c][0], out[c][1], out[c][2], out[c][3]);
}

// The same four-pixel stencil as rhs_kernel's vector path, as a function: the right-hand side of pixels x0 .. x0+3 of row y, all three
// channels.  Requires x0 + 8 <= nx (the word loads cover pixels x0 .. x0+5).  Used by rhs_fold_kernel below.
SCB_D void rhs_quad(const StencilSrc& s, int x0, int y, float (&out)[3][4]) {
    const int X = x0 + 1, Y = y + 1;
    // row Y from column X-1 (20 bytes cover X-1 .. X+4), rows Y-1 / Y+1 from column X (12 bytes cover X .. X+3)
    unsigned dm[5], du[3], dd[3], sm[5], su[3], sd[3], em[2], eu[1];
    load_unaligned_words<2>(s.E + (long long)Y * s.e_pitch + (X - 1), em);
    load_unaligned_words<1>(s.E + (long long)(Y - 1) * s.e_pitch + X, eu);
    {
        // Fast path (almost every thread): the nine mask taps of the four pixels are all 0 or all 255 and no
        // pixel touches the ROI border, so g is the 5-point Laplacian of ONE image, a 1-D stencil over its
        // interleaved bytes (taps at -3, +3, -pitch, +pitch bytes).  Only that image is read, and the twelve
        // values are formed two at a time in packed 16-bit lanes (biased by 2048 so that no borrow crosses a
        // lane) and converted with the 2^23 trick: ~35 instructions per pixel instead of ~140, which is what
        // lets the kernel run at HBM speed.  The results are small exact integers, hence bit-identical to
        // the float path below.
        const unsigned m_and = em[0] & eu[0] & (em[1] | 0xffffff00u), m_or = em[0] | eu[0] | (em[1] & 0xffu);
        const bool all_src = (m_and == 0xffffffffu), all_dst = (m_or == 0u);
        if ((all_src || all_dst) && x0 > 0 && Y > 1 && Y < s.h - 2) {
            const unsigned char* img = all_src ? s.S : s.D;
            const long long pitch = all_src ? s.s_pitch : s.d_pitch;
            unsigned m[5], u[3], d[3];
            load_unaligned_words<5>(img + (long long)Y * pitch + 3 * (X - 1), m);
            load_unaligned_words<3>(img + (long long)(Y - 1) * pitch + 3 * X, u);
            load_unaligned_words<3>(img + (long long)(Y + 1) * pitch + 3 * X, d);
            float o[12];
            SCB_UNROLL
            for (int j = 0; j < 3; ++j) {
                const unsigned l = m[j];                                        // bytes e-3
                const unsigned c = __funnelshift_r(m[j], m[j + 1], 24);         // bytes e
                const unsigned r = __funnelshift_r(m[j + 1], m[j + 2], 16);     // bytes e+3
                SCB_UNROLL
                for (int half = 0; half < 2; ++half) {
                    const unsigned sel = half ? 0x4342u : 0x4140u;  // two bytes -> two 16-bit lanes
                    unsigned t = 0x08000800u + __byte_perm(l, 0u, sel) + __byte_perm(r, 0u, sel) + __byte_perm(u[j], 0u, sel) + __byte_perm(d[j], 0u, sel);
                    t -= 4u * __byte_perm(c, 0u, sel);
                    o[4 * j + 2 * half + 0] = __uint_as_float(__byte_perm(t, 0x4b000000u, 0x7610u)) - 8390656.0f;  // (2^23 + 2048 + v) - (2^23 + 2048)
                    o[4 * j + 2 * half + 1] = __uint_as_float(__byte_perm(t, 0x4b000000u, 0x7632u)) - 8390656.0f;
                }
            }
            SCB_UNROLL
            for (int c = 0; c < 3; ++c) {
                out[c][0] = o[c];
                out[c][1] = o[3 + c];
                out[c][2] = o[6 + c];
                out[c][3] = o[9 + c];
            }
            return;
        }
    }
    load_unaligned_words<5>(s.D + (long long)Y * s.d_pitch + 3 * (X - 1), dm);
    load_unaligned_words<3>(s.D + (long long)(Y - 1) * s.d_pitch + 3 * X, du);
    load_unaligned_words<3>(s.D + (long long)(Y + 1) * s.d_pitch + 3 * X, dd);
    load_unaligned_words<5>(s.S + (long long)Y * s.s_pitch + 3 * (X - 1), sm);
    load_unaligned_words<3>(s.S + (long long)(Y - 1) * s.s_pitch + 3 * X, su);
    load_unaligned_words<3>(s.S + (long long)(Y + 1) * s.s_pitch + 3 * X, sd);
    SCB_UNROLL
    for (int k = 0; k < 4; ++k) {
        const int el = byte_of(em, k), ec = byte_of(em, k + 1), eup = byte_of(eu, k);
        const bool binary = ((el == 0) | (el == 255)) & ((ec == 0) | (ec == 255)) & ((eup == 0) | (eup == 255));
        const int Xk = X + k;
        if (binary) {
            SCB_UNROLL
            for (int c = 0; c < 3; ++c) {
                const int Dl = byte_of(dm, 3 * k + c), Dc = byte_of(dm, 3 * k + 3 + c), Dr = byte_of(dm, 3 * k + 6 + c);
                const int Du = byte_of(du, 3 * k + c), Dd = byte_of(dd, 3 * k + c);
                const int Sl = byte_of(sm, 3 * k + c), Sc = byte_of(sm, 3 * k + 3 + c), Sr = byte_of(sm, 3 * k + 6 + c);
                const int Su = byte_of(su, 3 * k + c), Sd = byte_of(sd, 3 * k + c);
                int lap = (ec ? (Sr - Sc) + (Sd - Sc) : (Dr - Dc) + (Dd - Dc)) - (el ? (Sc - Sl) : (Dc - Dl)) - (eup ? (Sc - Su) : (Dc - Du));
                if (Xk == 1) lap -= Dl;
                if (Xk == s.w - 2) lap -= Dr;
                if (Y == 1) lap -= Du;
                if (Y == s.h - 2) lap -= Dd;
                out[c][k] = (float)lap;
            }
        } else {  // grey mask values: OpenCV's float arithmetic, operation by operation
            const float inv255 = 1.0f / 255.0f;
            const float mc = __fmul_rn((float)ec, inv255), mic = __fmul_rn((float)(255 - ec), inv255);
            const float ml = __fmul_rn((float)el, inv255), mil = __fmul_rn((float)(255 - el), inv255);
            const float mu = __fmul_rn((float)eup, inv255), miu = __fmul_rn((float)(255 - eup), inv255);
            SCB_UNROLL
            for (int c = 0; c < 3; ++c) {
                const float Dl = (float)byte_of(dm, 3 * k + c), Dc = (float)byte_of(dm, 3 * k + 3 + c), Dr = (float)byte_of(dm, 3 * k + 6 + c);
                const float Du = (float)byte_of(du, 3 * k + c), Dd = (float)byte_of(dd, 3 * k + c);
                const float Sl = (float)byte_of(sm, 3 * k + c), Sc = (float)byte_of(sm, 3 * k + 3 + c), Sr = (float)byte_of(sm, 3 * k + 6 + c);
                const float Su = (float)byte_of(su, 3 * k + c), Sd = (float)byte_of(sd, 3 * k + c);
                const float vxc = __fadd_rn(__fmul_rn(Dr - Dc, mic), __fmul_rn(Sr - Sc, mc));
                const float vxl = __fadd_rn(__fmul_rn(Dc - Dl, mil), __fmul_rn(Sc - Sl, ml));
                const float vyc = __fadd_rn(__fmul_rn(Dd - Dc, mic), __fmul_rn(Sd - Sc, mc));
                const float vyu = __fadd_rn(__fmul_rn(Dc - Du, miu), __fmul_rn(Sc - Su, mu));
                const float lap = __fadd_rn(__fsub_rn(vxc, vxl), __fsub_rn(vyc, vyu));
                float bnd = 0.f;
                if (Xk == 1) bnd += Dl;
                if (Xk == s.w - 2) bnd += Dr;
                if (Y == 1) bnd += Du;
                if (Y == s.h - 2) bnd += Dd;
                out[c][k] = __fsub_rn(lap, bnd);
            }
        }
    }
}


// ---------------------------------------------------------------------------------------------
// Stencil fused with the fold + digit split of the INT8 DST engine (scb_i8.h): instead of the float right-hand side G, the kernel
// writes the balanced base-256 digit planes of  f0 = g[j] + g[n-1-j],  f1 = g[j] - g[n-1-j]  directly -- what i8_digitize_kernel would
// produce from G, bit for bit -- so G (12 B per pixel written, then read again) never exists and one launch disappears.
// A thread owns the folded elements j0 .. j0+3 of one row: the four pixels at x = j0 and their mirror images at x = n-4-j0.
//   planes[(par * DA + i) * m_rows + line][j],  line = 3 y + channel  (the engine's channel-interleaved line order)
// DA = 2: integer right-hand side (binary mask); DA = 4: 16 fractional bits (grey mask values).  Rows y >= ny are the zero pad lines.
// grid = (ceil(kpad / 4 / 128), rows), block = 128
// ---------------------------------------------------------------------------------------------
struct RhsFoldParams {
    StencilSrc st;
    int nx, ny;
    int kpar0, kpad;     // folded length (parity 0) and row pitch of the digit planes
    int m_rows, lines;   // padded / real line count (lines = 3 ny)
    signed char* planes;
    float* lscale;       // [m_rows] per-line scale of the digit planes (1 or 2^-16)
    float scale;         // 1 (DA = 2) or 65536 (DA = 4)
    int y0;
};

template <int DA>
__global__ void __launch_bounds__(kRhsThreads, 8) rhs_fold_kernel(RhsFoldParams p) {
    const int y = p.y0 + blockIdx.y;
    const int j0 = 4 * (blockIdx.x * kRhsThreads + threadIdx.x);
    if (j0 >= p.kpad) return;
    const int n = p.nx, h = n >> 1;
    float a[3][4], b[3][4];  // g at j0+e, and at its mirror n-1-(j0+e)
    SCB_UNROLL
    for (int c = 0; c < 3; ++c)
        SCB_UNROLL
        for (int e = 0; e < 4; ++e) a[c][e] = b[c][e] = 0.f;
    bool mid[4] = {false, false, false, false};  // the middle element of an odd line: f0 = g[h], f1 = 0
    if (y < p.ny && j0 < p.kpar0) {
        if (j0 >= 4 && j0 + 3 < h && j0 + 8 <= n) {
            float m[3][4];
            rhs_quad(p.st, j0, y, a);
            rhs_quad(p.st, n - 4 - j0, y, m);
            SCB_UNROLL
            for (int c = 0; c < 3; ++c)
                SCB_UNROLL
                for (int e = 0; e < 4; ++e) b[c][e] = m[c][3 - e];
        } else {  // the ends of the row and the thread that straddles the middle: pixel by pixel
            for (int e = 0; e < 4; ++e) {
                const int j = j0 + e;
                float g[3] = {0.f, 0.f, 0.f}, gm[3] = {0.f, 0.f, 0.f};
                if (j < h) {
                    rhs_pixel(p.st, j, y, g);
                    rhs_pixel(p.st, n - 1 - j, y, gm);
                } else if (j == h && (n & 1)) {
                    rhs_pixel(p.st, h, y, g);
                    mid[e] = true;
                }
                for (int c = 0; c < 3; ++c) {
                    a[c][e] = g[c];
                    b[c][e] = gm[c];
                }
            }
        }
    }
    SCB_UNROLL
    for (int c = 0; c < 3; ++c) {
        const int line = 3 * y + c;
        if (line >= p.m_rows) continue;
        if (j0 == 0) p.lscale[line] = 1.0f / p.scale;
        unsigned w[2][DA];
        int f0[4], f1[4];
        SCB_UNROLL
        for (int e = 0; e < 4; ++e) {
            const int va = __float2int_rn(a[c][e] * p.scale), vb = __float2int_rn(b[c][e] * p.scale);
            f0[e] = va + vb;
            f1[e] = mid[e] ? 0 : va - vb;
        }
        balanced_digits4<DA>(f0, w[0]);
        balanced_digits4<DA>(f1, w[1]);
        SCB_UNROLL
        for (int q = 0; q < 2; ++q)
            SCB_UNROLL
            for (int i = 0; i < DA; ++i)
                *reinterpret_cast<unsigned*>(p.planes + ((size_t)(q * DA + i) * p.m_rows + line) * p.kpad + j0) = w[q][i];
    }
}

// ---------------------------------------------------------------------------------------------
// rhs_fold2_kernel: the binary-mask (DA = 2) stencil + fold + digit split entirely in packed 16-bit integer lanes (SCB_RHS_FOLD=2).
// Same thread mapping, same bytes out as rhs_fold_kernel<2>; what changes is the instruction stream:
//   * the twelve right-hand-side values of a quad stay in six words of two 16-bit lanes (value + 2048, byte order 3 pixel + channel)
//     from the stencil to the digits -- no float conversion, no per-value int arithmetic;
//   * the fold is one add and one subtract per word against the mirror quad (whose lanes are reversed pixel-wise by six
//     16-bit funnel shifts), biased so that the lane holds f + 0x8080: XOR 0x8080 then leaves the two balanced base-256 digits of f in
//     the lane's two bytes (balanced_digits4 in 16 bits), and four byte permutes per line and parity gather the digit words;
//   * the two ends of the row and the thread that straddles the middle run the same quad code (border terms in the general path,
//     lane masks for the elements past the middle) instead of a pixel-by-pixel path that made their whole warp ~4 x slower;
//   * the mask words of both quads are fetched before either image, the image words of both quads together: two dependent
//     memory round trips per thread instead of four.
// Loads: the mid-row words of the quad at the right end of the row cover 2 bytes past the ROI row (pixel w); rows 1 .. h-2 are never
// the last row of the image, so those bytes lie inside the caller's buffer (next pixel of the row, or the start of the next row).
// ---------------------------------------------------------------------------------------------
struct QuadTaps {
    unsigned em[2], eu[1];
    bool fast, src;
};

// mask taps of the quad at x0 (pixels x0 .. x0+3 of interior row y) and its classification
SCB_D void quad_taps(const StencilSrc& s, int x0, int y, int n, QuadTaps& q) {
    const int X = x0 + 1, Y = y + 1;
    load_unaligned_words<2>(s.E + (long long)Y * s.e_pitch + (X - 1), q.em);
    load_unaligned_words<1>(s.E + (long long)(Y - 1) * s.e_pitch + X, q.eu);
    const unsigned m_and = q.em[0] & q.eu[0] & (q.em[1] | 0xffffff00u), m_or = q.em[0] | q.eu[0] | (q.em[1] & 0xffu);
    q.src = (m_and == 0xffffffffu);
    // one image only and no pixel of the quad on the ROI border: the 5-point Laplacian of that image
    q.fast = (q.src || m_or == 0u) && x0 > 0 && x0 + 4 < n && Y > 1 && Y < s.h - 2;
}

struct QuadWords {
    unsigned m[5], u[3], d[3];
};
SCB_D void quad_load(const unsigned char* img, long long pitch, int x0, int y, QuadWords& w) {
    const int X = x0 + 1, Y = y + 1;
    load_unaligned_words<5>(img + (long long)Y * pitch + 3 * (X - 1), w.m);
    load_unaligned_words<3>(img + (long long)(Y - 1) * pitch + 3 * X, w.u);
    load_unaligned_words<3>(img + (long long)(Y + 1) * pitch + 3 * X, w.d);
}
// 2048 + (l + r + u + d - 4 c) for the twelve bytes of the quad, two per word
SCB_D void quad_fast_lanes(const QuadWords& w, unsigned (&t)[6]) {
    SCB_UNROLL
    for (int j = 0; j < 3; ++j) {
        const unsigned l = w.m[j];                                          // bytes e-3
        const unsigned c = __funnelshift_r(w.m[j], w.m[j + 1], 24);         // bytes e
        const unsigned r = __funnelshift_r(w.m[j + 1], w.m[j + 2], 16);     // bytes e+3
        SCB_UNROLL
        for (int half = 0; half < 2; ++half) {
            const unsigned sel = half ? 0x4342u : 0x4140u;  // two bytes -> two 16-bit lanes
            unsigned v = 0x08000800u + __byte_perm(l, 0u, sel) + __byte_perm(r, 0u, sel) + __byte_perm(w.u[j], 0u, sel) + __byte_perm(w.d[j], 0u, sel);
            v -= 4u * __byte_perm(c, 0u, sel);
            t[2 * j + half] = v;
        }
    }
}
// any binary mask pattern and the ROI borders: integer arithmetic per pixel (rhs_quad's binary branch), packed into the same lanes
SCB_D void quad_general_lanes(const StencilSrc& s, int x0, int y, const QuadTaps& q, unsigned (&t)[6]) {
    const int X = x0 + 1, Y = y + 1;
    QuadWords D, S;
    quad_load(s.D, s.d_pitch, x0, y, D);
    quad_load(s.S, s.s_pitch, x0, y, S);
    SCB_UNROLL
    for (int i = 0; i < 6; ++i) t[i] = 0u;
    SCB_UNROLL
    for (int k = 0; k < 4; ++k) {
        const int el = byte_of(q.em, k), ec = byte_of(q.em, k + 1), eup = byte_of(q.eu, k);
        const int Xk = X + k;
        SCB_UNROLL
        for (int c = 0; c < 3; ++c) {
            const int Dl = byte_of(D.m, 3 * k + c), Dc = byte_of(D.m, 3 * k + 3 + c), Dr = byte_of(D.m, 3 * k + 6 + c);
            const int Du = byte_of(D.u, 3 * k + c), Dd = byte_of(D.d, 3 * k + c);
            const int Sl = byte_of(S.m, 3 * k + c), Sc = byte_of(S.m, 3 * k + 3 + c), Sr = byte_of(S.m, 3 * k + 6 + c);
            const int Su = byte_of(S.u, 3 * k + c), Sd = byte_of(S.d, 3 * k + c);
            int lap = (ec ? (Sr - Sc) + (Sd - Sc) : (Dr - Dc) + (Dd - Dc)) - (el ? (Sc - Sl) : (Dc - Dl)) - (eup ? (Sc - Su) : (Dc - Du));
            if (Xk == 1) lap -= Dl;
            if (Xk == s.w - 2) lap -= Dr;
            if (Y == 1) lap -= Du;
            if (Y == s.h - 2) lap -= Dd;
            const int b = 3 * k + c;
            t[b >> 1] += (unsigned)(2048 + lap) << (16 * (b & 1));  // |lap| <= 1530
        }
    }
}

#ifndef SCB_RHS2_MIN_CTAS
#define SCB_RHS2_MIN_CTAS 6  // 80 registers; 8 would cap them at 64
#endif
__global__ void __launch_bounds__(kRhsThreads, SCB_RHS2_MIN_CTAS) rhs_fold2_kernel(RhsFoldParams p) {
    constexpr int DA = 2;
    const int y = p.y0 + blockIdx.y;
    const int j0 = 4 * (blockIdx.x * kRhsThreads + threadIdx.x);
    if (j0 >= p.kpad) return;
    const int n = p.nx, h = n >> 1;
    unsigned lo[2][3], hi[2][3];  // [parity][channel]: the four low / high digits of the folded elements j0 .. j0+3
    SCB_UNROLL
    for (int q = 0; q < 2; ++q)
        SCB_UNROLL
        for (int c = 0; c < 3; ++c) lo[q][c] = hi[q][c] = 0u;
    if (y < p.ny && j0 < p.kpar0) {
        const StencilSrc& s = p.st;
        const int xa = j0, xb = n - 4 - j0;  // the quad and its mirror image (for the straddling thread they overlap: masked below)
        QuadTaps qa, qb;
        quad_taps(s, xa, y, n, qa);
        quad_taps(s, xb, y, n, qb);
        unsigned ta[6], tm[6];
        if (qa.fast && qb.fast) {
            QuadWords wa, wb;
            quad_load(qa.src ? s.S : s.D, qa.src ? s.s_pitch : s.d_pitch, xa, y, wa);
            quad_load(qb.src ? s.S : s.D, qb.src ? s.s_pitch : s.d_pitch, xb, y, wb);
            quad_fast_lanes(wa, ta);
            quad_fast_lanes(wb, tm);
        } else {
            if (qa.fast) {
                QuadWords wa;
                quad_load(qa.src ? s.S : s.D, qa.src ? s.s_pitch : s.d_pitch, xa, y, wa);
                quad_fast_lanes(wa, ta);
            } else {
                quad_general_lanes(s, xa, y, qa, ta);
            }
            if (qb.fast) {
                QuadWords wb;
                quad_load(qb.src ? s.S : s.D, qb.src ? s.s_pitch : s.d_pitch, xb, y, wb);
                quad_fast_lanes(wb, tm);
            } else {
                quad_general_lanes(s, xb, y, qb, tm);
            }
        }
        // mirror quad, pixels reversed: lane 3 e + c  <-  lane 3 (3 - e) + c
        unsigned tr[6];
        tr[0] = __funnelshift_r(tm[4], tm[5], 16);  // lanes  9, 10
        tr[1] = __funnelshift_r(tm[5], tm[3], 16);  // lanes 11,  6
        tr[2] = __funnelshift_r(tm[3], tm[4], 16);  // lanes  7,  8
        tr[3] = __funnelshift_r(tm[1], tm[2], 16);  // lanes  3,  4
        tr[4] = __funnelshift_r(tm[2], tm[0], 16);  // lanes  5,  0
        tr[5] = __funnelshift_r(tm[0], tm[1], 16);  // lanes  1,  2
        unsigned mid_mask[6] = {0u, 0u, 0u, 0u, 0u, 0u};
        if (j0 + 3 >= h) {  // the thread that straddles the middle: element j >= h is zero, except the middle of an odd line (f0 = g[h], f1 = 0)
            SCB_UNROLL
            for (int L = 0; L < 12; ++L) {
                const int j = j0 + L / 3;
                const unsigned lane = 0xffffu << (16 * (L & 1));
                const bool is_mid = (j == h) && (n & 1);
                if (j >= h) {
                    tr[L >> 1] = (tr[L >> 1] & ~lane) | (0x08000800u & lane);
                    if (is_mid)
                        mid_mask[L >> 1] |= lane;
                    else
                        ta[L >> 1] = (ta[L >> 1] & ~lane) | (0x08000800u & lane);
                }
            }
        }
        // lanes hold value + 2048:  f0 + 0x8080 = ta + tr + 0x7080,  f1 + 0x8080 = ta - tr + 0x8080;  XOR 0x8080 -> the two balanced digits
        unsigned x0w[6], x1w[6];
        SCB_UNROLL
        for (int i = 0; i < 6; ++i) {
            x0w[i] = (ta[i] + tr[i] + 0x70807080u) ^ 0x80808080u;
            x1w[i] = (((ta[i] + 0x80808080u) - tr[i]) ^ 0x80808080u) & ~mid_mask[i];  // digits of zero are zero bytes
        }
        SCB_UNROLL
        for (int c = 0; c < 3; ++c) {
            // lanes c, 3+c | 6+c, 9+c: elements 0, 1 | 2, 3 of channel c
            const unsigned sel = (c == 1) ? 0x5432u : 0x7610u;
            const int wA = c >> 1, wB = (3 + c) >> 1, wC = (6 + c) >> 1, wD = (9 + c) >> 1;
            const unsigned p0 = __byte_perm(x0w[wA], x0w[wB], sel), q0 = __byte_perm(x0w[wC], x0w[wD], sel);
            const unsigned p1 = __byte_perm(x1w[wA], x1w[wB], sel), q1 = __byte_perm(x1w[wC], x1w[wD], sel);
            lo[0][c] = __byte_perm(p0, q0, 0x6420u);
            hi[0][c] = __byte_perm(p0, q0, 0x7531u);
            lo[1][c] = __byte_perm(p1, q1, 0x6420u);
            hi[1][c] = __byte_perm(p1, q1, 0x7531u);
        }
    }
    SCB_UNROLL
    for (int c = 0; c < 3; ++c) {
        const int line = 3 * y + c;
        if (line >= p.m_rows) continue;
        if (j0 == 0) p.lscale[line] = 1.0f / p.scale;
        SCB_UNROLL
        for (int q = 0; q < 2; ++q) {
            *reinterpret_cast<unsigned*>(p.planes + ((size_t)(q * DA + 0) * p.m_rows + line) * p.kpad + j0) = hi[q][c];
            *reinterpret_cast<unsigned*>(p.planes + ((size_t)(q * DA + 1) * p.m_rows + line) * p.kpad + j0) = lo[q][c];
        }
    }
}

// ---------------------------------------------------------------------------------------------
// The other cv::seamlessClone flags: the same solver behind a different gradient selection
// (Cloning::normalClone in OpenCV's seamless_cloning_impl.cpp).
//   MIXED_CLONE          per pixel and channel, the patch gradient PAIR is replaced by dst's unless
//                        |gxS - gyS| > |gxD - gyD|
//   MONOCHROME_TRANSFER  the patch gradients are those of cvtColor(patch, BGR2GRAY) (15-bit fixed point), all 3 channels
// One thread per interior pixel, OpenCV's float operations in OpenCV's order.  The selection at (X-1, Y) and (X, Y-1)
// needs the taps (X-1, Y+1) and (X+1, Y-1) on top of the 5-point cross.
// ---------------------------------------------------------------------------------------------
enum { kModeNormal = 1, kModeMixed = 2, kModeMono = 3 };

SCB_D float gray_of(const unsigned char* px) {  // cv::cvtColor(BGR2GRAY), 8-bit: (B*3735 + G*19235 + R*9798 + 2^14) >> 15
    return (float)((__ldg(px) * 3735 + __ldg(px + 1) * 19235 + __ldg(px + 2) * 9798 + (1 << 14)) >> 15);
}

// forward differences of both images at (X, Y) towards (Xn, Y) and (X, Yn), after the mode's selection
SCB_D void mode_gradients(const StencilSrc& s, int mode, int c, int X, int Y, int Xn, int Yn, float& gxD, float& gyD, float& gxS, float& gyS) {
    const unsigned char* d = s.D + (long long)Y * s.d_pitch + 3 * X + c;
    const float Dc = (float)__ldg(d);
    gxD = (float)__ldg(s.D + (long long)Y * s.d_pitch + 3 * Xn + c) - Dc;
    gyD = (float)__ldg(s.D + (long long)Yn * s.d_pitch + 3 * X + c) - Dc;
    if (mode == kModeMono) {
        const float Sc = gray_of(s.S + (long long)Y * s.s_pitch + 3 * X);
        gxS = gray_of(s.S + (long long)Y * s.s_pitch + 3 * Xn) - Sc;
        gyS = gray_of(s.S + (long long)Yn * s.s_pitch + 3 * X) - Sc;
    } else {
        const float Sc = (float)__ldg(s.S + (long long)Y * s.s_pitch + 3 * X + c);
        gxS = (float)__ldg(s.S + (long long)Y * s.s_pitch + 3 * Xn + c) - Sc;
        gyS = (float)__ldg(s.S + (long long)Yn * s.s_pitch + 3 * X + c) - Sc;
    }
    if (mode == kModeMixed && !(fabsf(gxS - gyS) > fabsf(gxD - gyD))) {
        gxS = gxD;
        gyS = gyD;
    }
}

__global__ void __launch_bounds__(256) rhs_mode_kernel(RhsParams p, int mode) {
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = p.y0 + blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= p.gp || y >= p.y1) return;
    const StencilSrc& s = p.st;
    float g[3] = {0.f, 0.f, 0.f};
    if (x < p.nx) {
        const int X = x + 1, Y = y + 1;
        const float inv255 = 1.0f / 255.0f;
        const int ec = __ldg(s.E + (long long)Y * s.e_pitch + X), el = __ldg(s.E + (long long)Y * s.e_pitch + X - 1), eu = __ldg(s.E + (long long)(Y - 1) * s.e_pitch + X);
        const float mc = __fmul_rn((float)ec, inv255), mic = __fmul_rn((float)(255 - ec), inv255);
        const float ml = __fmul_rn((float)el, inv255), mil = __fmul_rn((float)(255 - el), inv255);
        const float mu = __fmul_rn((float)eu, inv255), miu = __fmul_rn((float)(255 - eu), inv255);
        for (int c = 0; c < 3; ++c) {
            float gxD, gyD, gxS, gyS;
            mode_gradients(s, mode, c, X, Y, X + 1, Y + 1, gxD, gyD, gxS, gyS);
            const float vxc = __fadd_rn(__fmul_rn(gxD, mic), __fmul_rn(gxS, mc));
            const float vyc = __fadd_rn(__fmul_rn(gyD, mic), __fmul_rn(gyS, mc));
            mode_gradients(s, mode, c, X - 1, Y, X, Y + 1, gxD, gyD, gxS, gyS);
            const float vxl = __fadd_rn(__fmul_rn(gxD, mil), __fmul_rn(gxS, ml));
            mode_gradients(s, mode, c, X, Y - 1, X + 1, Y, gxD, gyD, gxS, gyS);
            const float vyu = __fadd_rn(__fmul_rn(gyD, miu), __fmul_rn(gyS, mu));
            const float lap = __fadd_rn(__fsub_rn(vxc, vxl), __fsub_rn(vyc, vyu));
            float bnd = 0.f;
            if (X == 1) bnd += (float)__ldg(s.D + (long long)Y * s.d_pitch + 3 * (X - 1) + c);
            if (X == s.w - 2) bnd += (float)__ldg(s.D + (long long)Y * s.d_pitch + 3 * (X + 1) + c);
            if (Y == 1) bnd += (float)__ldg(s.D + (long long)(Y - 1) * s.d_pitch + 3 * X + c);
            if (Y == s.h - 2) bnd += (float)__ldg(s.D + (long long)(Y + 1) * s.d_pitch + 3 * X + c);
            g[c] = __fsub_rn(lap, bnd);
        }
    }
    for (int c = 0; c < 3; ++c) p.g[((size_t)c * p.ny + y) * p.gp + x] = g[c];
}

// ---------------------------------------------------------------------------------------------
// pass A: stencil -> forward DST-I along x.   grid = ny CTAs, block = FftCfg::T
// ---------------------------------------------------------------------------------------------
struct RowsFwdParams {
    StencilSrc st;
    LenTabDev tx;
    int nx, ny;
    float* At;            // [3][nx][ny]
    float* rhs_dump;      // [3][ny][nx] or null
    const float* rhs_in;  // [3][ny][rhs_pitch] from rhs_kernel, or null: evaluate the stencil in place
    int rhs_pitch;
    int y0;               // first interior row handled by this launch (row-sharded solves)
    int natural;          // 1: store At as [3][ny][nx] (tridiagonal engine: the column solve reads rows of 32 columns)
};

template <int LOG2M, int NCH>
__global__ void __launch_bounds__(FftCfg<LOG2M>::T) rows_fwd_kernel(RowsFwdParams p) {
    using C = FftCfg<LOG2M>;
    SCB_DYN_SMEM(float2, buf);
    const int tid = threadIdx.x, y = p.y0 + blockIdx.x, n = p.nx;
    for (int c0 = 0; c0 < 3; c0 += NCH) {
        for (int j = tid; j < C::M; j += C::T) {
            float g[3] = {0.f, 0.f, 0.f};
            float2 ch = make_float2(0.f, 0.f);
            if (j >= 1 && j <= n) {
                if (p.rhs_in) {
                    SCB_UNROLL
                    for (int c = 0; c < 3; ++c) g[c] = p.rhs_in[((size_t)c * p.ny + y) * p.rhs_pitch + (j - 1)];
                } else {
                    rhs_pixel(p.st, j - 1, y, g);
                }
                ch = __ldg(p.tx.chirp + j);
                if (p.rhs_dump) {
                    SCB_UNROLL
                    for (int c = 0; c < NCH; ++c) p.rhs_dump[((size_t)(c0 + c) * p.ny + y) * p.nx + (j - 1)] = g[c0 + c];
                }
            }
            SCB_UNROLL
            for (int c = 0; c < NCH; ++c) buf[c * C::PADDED + padi(j)] = make_float2(g[c0 + c] * ch.x, g[c0 + c] * ch.y);
        }
        __syncthreads();
        fft_convolve<LOG2M, NCH>(buf, p.tx.tw, p.tx.bhat_t, tid);
        for (int k = tid + 1; k <= n; k += C::T) {
            const float2 ch = __ldg(p.tx.chirp + k);
            SCB_UNROLL
            for (int c = 0; c < NCH; ++c) {
                const float2 v = buf[c * C::PADDED + padi(k)];
                const float s = ch.x * v.y + ch.y * v.x;  // Im(c[k] * conv[k]) = sum_j x[j] sin(pi j k / N)
                const size_t o = p.natural ? ((size_t)(c0 + c) * p.ny + y) * p.nx + (k - 1) : ((size_t)(c0 + c) * p.nx + (k - 1)) * p.ny + y;
                p.At[o] = -2.0f * s;  // OpenCV: Im of the odd-extension FFT
            }
        }
        if (NCH < 3) __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// pass B: forward DST-I along y, eigenvalue division, inverse DST-I along y.  grid = nx CTAs
// ---------------------------------------------------------------------------------------------
struct ColsParams {
    LenTabDev ty;
    int nx, ny;
    const float* At;       // [3][nx][ny]
    float* Ct;             // [3][ny][nx]
    const float* fx;       // nx   OpenCV filter_X
    const float* fy;       // ny   OpenCV filter_Y
    const float* lowspec;  // [3][lowkx][lowky] exact low-frequency corner of the forward spectrum, or null
    int lowkx, lowky;
    float* spec_dump;      // [3][nx][ny] forward 2-D spectrum before the division, or null
    float inv_scale;       // 1 / (ny + 1)
    int x0;                // first column handled by this launch (column-sharded solves)
};

template <int LOG2M, int NCH>
__global__ void __launch_bounds__(FftCfg<LOG2M>::T) cols_kernel(ColsParams p) {
    using C = FftCfg<LOG2M>;
    SCB_DYN_SMEM(float2, buf);
    const int tid = threadIdx.x, kx = p.x0 + blockIdx.x, n = p.ny;
    const float fxv = __ldg(p.fx + kx);
    for (int c0 = 0; c0 < 3; c0 += NCH) {
        for (int j = tid; j < C::M; j += C::T) {
            const bool in = (j >= 1 && j <= n);
            const float2 ch = in ? __ldg(p.ty.chirp + j) : make_float2(0.f, 0.f);
            SCB_UNROLL
            for (int c = 0; c < NCH; ++c) {
                const float v = in ? __ldg(p.At + ((size_t)(c0 + c) * p.nx + kx) * p.ny + (j - 1)) : 0.f;
                buf[c * C::PADDED + padi(j)] = make_float2(v * ch.x, v * ch.y);
            }
        }
        __syncthreads();
        fft_convolve<LOG2M, NCH>(buf, p.ty.tw, p.ty.bhat_t, tid);
        for (int j = tid; j < C::M; j += C::T) {
            const bool in = (j >= 1 && j <= n);
            float2 ch = make_float2(0.f, 0.f);
            float den = 1.f;
            if (in) {
                ch = __ldg(p.ty.chirp + j);
                // OpenCV: res /= (filter_X[i] + filter_Y[j] - 4), evaluated left to right in float32
                den = __fsub_rn(__fadd_rn(fxv, __ldg(p.fy + (j - 1))), 4.0f);
            }
            SCB_UNROLL
            for (int c = 0; c < NCH; ++c) {
                float2 o = make_float2(0.f, 0.f);
                if (in) {
                    const float2 v = buf[c * C::PADDED + padi(j)];
                    float s = -2.0f * (ch.x * v.y + ch.y * v.x);
                    if (p.lowspec && kx < p.lowkx && (j - 1) < p.lowky) s = __ldg(p.lowspec + ((size_t)(c0 + c) * p.lowkx + kx) * p.lowky + (j - 1));
                    if (p.spec_dump) p.spec_dump[((size_t)(c0 + c) * p.nx + kx) * p.ny + (j - 1)] = s;
                    const float q = __fdiv_rn(s, den);
                    o = make_float2(q * ch.x, q * ch.y);
                }
                buf[c * C::PADDED + padi(j)] = o;
            }
        }
        __syncthreads();
        fft_convolve<LOG2M, NCH>(buf, p.ty.tw, p.ty.bhat_t, tid);
        for (int k = tid + 1; k <= n; k += C::T) {
            const float2 ch = __ldg(p.ty.chirp + k);
            SCB_UNROLL
            for (int c = 0; c < NCH; ++c) {
                const float2 v = buf[c * C::PADDED + padi(k)];
                p.Ct[((size_t)(c0 + c) * p.ny + (k - 1)) * p.nx + kx] = (ch.x * v.y + ch.y * v.x) * p.inv_scale;
            }
        }
        if (NCH < 3) __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// pass C: inverse DST-I along x, clamp, truncate, interleaved u8 store.   grid = ny CTAs
// ---------------------------------------------------------------------------------------------
struct RowsInvParams {
    LenTabDev tx;
    int nx, ny;
    const float* Ct;     // [3][ny][nx]
    unsigned char* out;  // interior origin pixel of blend (or of the staging image)
    long long out_pitch;
    float* u_dump;       // [3][ny][nx] solved field before clamp/truncate, or null
    float inv_scale;     // 1 / (nx + 1)
    int y0;
    int transposed;      // 1: the lines run along y of the image: line L, position k -> pixel (row k, column L)
};

SCB_D unsigned char compose_u8(float v) {
    // OpenCV solve(): v < 0 -> 0, v > 255 -> 255, else static_cast<uchar>(v)  (truncation toward zero)
    return v < 0.f ? (unsigned char)0 : (v > 255.f ? (unsigned char)255 : (unsigned char)__float2int_rz(v));
}

template <int LOG2M, int NCH>
__global__ void __launch_bounds__(FftCfg<LOG2M>::T) rows_inv_kernel(RowsInvParams p) {
    using C = FftCfg<LOG2M>;
    SCB_DYN_SMEM(float2, buf);
    const int tid = threadIdx.x, y = p.y0 + blockIdx.x, n = p.nx;
    for (int c0 = 0; c0 < 3; c0 += NCH) {
        for (int j = tid; j < C::M; j += C::T) {
            const bool in = (j >= 1 && j <= n);
            const float2 ch = in ? __ldg(p.tx.chirp + j) : make_float2(0.f, 0.f);
            SCB_UNROLL
            for (int c = 0; c < NCH; ++c) {
                const float v = in ? __ldg(p.Ct + ((size_t)(c0 + c) * p.ny + y) * p.nx + (j - 1)) : 0.f;
                buf[c * C::PADDED + padi(j)] = make_float2(v * ch.x, v * ch.y);
            }
        }
        __syncthreads();
        fft_convolve<LOG2M, NCH>(buf, p.tx.tw, p.tx.bhat_t, tid);
        for (int k = tid + 1; k <= n; k += C::T) {
            const float2 ch = __ldg(p.tx.chirp + k);
            unsigned char* px = p.transposed ? p.out + (long long)(k - 1) * p.out_pitch + 3 * y : p.out + (long long)y * p.out_pitch + 3 * (k - 1);
            SCB_UNROLL
            for (int c = 0; c < NCH; ++c) {
                const float2 v = buf[c * C::PADDED + padi(k)];
                const float u = (ch.x * v.y + ch.y * v.x) * p.inv_scale;
                if (p.u_dump) p.u_dump[p.transposed ? ((size_t)(c0 + c) * p.nx + (k - 1)) * p.ny + y : ((size_t)(c0 + c) * p.ny + y) * p.nx + (k - 1)] = u;
                px[c0 + c] = compose_u8(u);
            }
        }
        if (NCH < 3) __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// exact low-frequency corner of the forward spectrum (float64 direct sums)
//   The Bluestein passes have a white float32 error floor; dividing by eigenvalues ~ (pi k / N)^2
//   amplifies it at the lowest bins only.  Recomputing the kLowK x kLowK corner exactly removes
//   > 95 % of the solve's error energy for ~1 % extra work (DESIGN.md "low-frequency refinement").
// ---------------------------------------------------------------------------------------------
#ifndef SCB_LOWK
#define SCB_LOWK 4
#endif
static constexpr int kLowKDev = SCB_LOWK;  // == kLowK of scb_tables.h
static constexpr int kLowThreads = 128;

template <int NACC>
SCB_D void block_reduce_store(double (&acc)[NACC], double* red /* [kLowThreads/32][NACC] */, int tid) {
    SCB_UNROLL
    for (int i = 0; i < NACC; ++i) {
        double v = acc[i];
        for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
        acc[i] = v;
    }
    if ((tid & 31) == 0) {
        SCB_UNROLL
        for (int i = 0; i < NACC; ++i) red[(tid >> 5) * NACC + i] = acc[i];
    }
    __syncthreads();
}

struct LowRowsParams {
    StencilSrc st;
    const double* sinx;  // [lowkx][nx]
    int nx, ny, lowkx;
    double* R;           // [3][lowkx][ny]   R = sum_x g[y][x] sin(pi (x+1)(k+1) / Nx)
    const float* rhs_in;
    int rhs_pitch;
    int y0;
};

__global__ void __launch_bounds__(kLowThreads) lowfreq_rows_kernel(LowRowsParams p) {
    __shared__ double red[(kLowThreads / 32) * 3 * kLowKDev];
    const int tid = threadIdx.x, y = p.y0 + blockIdx.x;
    double acc[3 * kLowKDev];
    SCB_UNROLL
    for (int i = 0; i < 3 * kLowKDev; ++i) acc[i] = 0.0;
    for (int x = tid; x < p.nx; x += kLowThreads) {
        float g[3];
        if (p.rhs_in) {
            SCB_UNROLL
            for (int c = 0; c < 3; ++c) g[c] = p.rhs_in[((size_t)c * p.ny + y) * p.rhs_pitch + x];
        } else {
            rhs_pixel(p.st, x, y, g);
        }
        SCB_UNROLL
        for (int k = 0; k < kLowKDev; ++k) {
            if (k < p.lowkx) {
                const double s = __ldg(p.sinx + (size_t)k * p.nx + x);
                SCB_UNROLL
                for (int c = 0; c < 3; ++c) acc[c * kLowKDev + k] += (double)g[c] * s;
            }
        }
    }
    block_reduce_store<3 * kLowKDev>(acc, red, tid);
    if (tid < 3 * kLowKDev) {
        double s = 0.0;
        for (int w = 0; w < kLowThreads / 32; ++w) s += red[w * 3 * kLowKDev + tid];
        const int c = tid / kLowKDev, k = tid % kLowKDev;
        if (k < p.lowkx) p.R[((size_t)c * p.lowkx + k) * p.ny + y] = s;
    }
}

struct LowColsParams {
    const double* R;     // [3][lowkx][ny]
    const double* siny;  // [lowky][ny]
    int ny, lowkx, lowky;
    float* lowspec;      // [3][lowkx][lowky]  = 4 * sum_y sum_x g sin sin   (OpenCV's unnormalised forward)
};

__global__ void __launch_bounds__(kLowThreads) lowfreq_cols_kernel(LowColsParams p) {
    __shared__ double red[(kLowThreads / 32) * kLowKDev];
    const int tid = threadIdx.x;
    const int c = blockIdx.x / p.lowkx, kx = blockIdx.x % p.lowkx;
    double acc[kLowKDev];
    SCB_UNROLL
    for (int i = 0; i < kLowKDev; ++i) acc[i] = 0.0;
    const double* r = p.R + ((size_t)c * p.lowkx + kx) * p.ny;
    for (int y = tid; y < p.ny; y += kLowThreads) {
        const double rv = r[y];
        SCB_UNROLL
        for (int k = 0; k < kLowKDev; ++k)
            if (k < p.lowky) acc[k] += rv * __ldg(p.siny + (size_t)k * p.ny + y);
    }
    block_reduce_store<kLowKDev>(acc, red, tid);
    if (tid < p.lowky) {
        double s = 0.0;
        for (int w = 0; w < kLowThreads / 32; ++w) s += red[w * kLowKDev + tid];
        p.lowspec[((size_t)c * p.lowkx + kx) * p.lowky + tid] = (float)(4.0 * s);
    }
}

// ---------------------------------------------------------------------------------------------
// mask preparation (plan time)
// ---------------------------------------------------------------------------------------------
struct MaskView {
    const unsigned char* data;
    long long pitch;
    int rows, cols;
};

// bbox of non-zero pixels of the ring-zeroed mask.  bbox = {minx, miny, maxx, maxy, grey}, pre-set to
// {INT_MAX, INT_MAX, -1, -1, 0}; grey = 1 when a pixel inside the ring is neither 0 nor 255 (then the blended right-hand
// side is not integer valued: the INT8 engine digitises it with 16 fractional bits).  OpenCV: copyMakeBorder(mask(1..-1), 0) + boundingRect.
// A warp owns whole rows (lanes stride over the columns: coalesced byte loads, no index division).
__global__ void __launch_bounds__(256) mask_bbox_kernel(MaskView m, int* bbox) {
    int minx = 0x7fffffff, miny = 0x7fffffff, maxx = -1, maxy = -1, grey = 0;
    const int lane = threadIdx.x & 31;
    const int warp = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5), nwarps = (int)((gridDim.x * blockDim.x) >> 5);
    for (int y = 1 + warp; y < m.rows - 1; y += nwarps) {
        const unsigned char* row = m.data + (long long)y * m.pitch;
        int lo = 0x7fffffff, hi = -1;
        for (int x = 1 + lane; x < m.cols - 1; x += 32) {
            const unsigned char v = __ldg(row + x);
            if (v) {
                lo = min(lo, x);
                hi = x;  // x grows within the lane
                grey |= (v != 255);
            }
        }
        if (hi >= 0) {
            minx = min(minx, lo);
            maxx = max(maxx, hi);
            miny = min(miny, y);
            maxy = max(maxy, y);
        }
    }
    for (int off = 16; off > 0; off >>= 1) {
        minx = min(minx, __shfl_xor_sync(0xffffffffu, minx, off));
        miny = min(miny, __shfl_xor_sync(0xffffffffu, miny, off));
        maxx = max(maxx, __shfl_xor_sync(0xffffffffu, maxx, off));
        maxy = max(maxy, __shfl_xor_sync(0xffffffffu, maxy, off));
        grey |= __shfl_xor_sync(0xffffffffu, grey, off);
    }
    if (lane == 0 && grey) atomicMax(bbox + 4, 1);
    if (lane == 0 && maxx >= 0) {
        atomicMin(bbox + 0, minx);
        atomicMin(bbox + 1, miny);
        atomicMax(bbox + 2, maxx);
        atomicMax(bbox + 3, maxy);
    }
}

// E = erode(ring-zeroed mask, 3x3 ones, iterations = 3) cropped to the ROI: a 7x7 minimum over the
// FULL mask (pixels outside the ROI take part; outside the image the border does not lower the min).
// Separable through shared memory: a 64 x 16 output tile stages its (64+6) x (16+6) inputs once, takes the
// horizontal 7-minimum, then the vertical one.   grid = (ceil(w/64), ceil(h/16)), block = 256
static constexpr int kErodeTW = 64, kErodeTH = 16;

__global__ void __launch_bounds__(256) mask_erode_kernel(MaskView m, int x0, int y0, int w, int h, unsigned char* E, long long e_pitch) {
    __shared__ unsigned char raw[kErodeTH + 6][kErodeTW + 8];
    __shared__ unsigned char hm[kErodeTH + 6][kErodeTW];
    const int tid = threadIdx.x, bx = blockIdx.x * kErodeTW, by = blockIdx.y * kErodeTH;
    for (int i = tid; i < (kErodeTH + 6) * (kErodeTW + 6); i += 256) {
        const int ry = i / (kErodeTW + 6), rx = i - ry * (kErodeTW + 6);
        const int yy = y0 + by + ry - 3, xx = x0 + bx + rx - 3;
        int v = 255;
        if (yy >= 0 && yy < m.rows && xx >= 0 && xx < m.cols) {
            const bool ring = (xx == 0 || yy == 0 || xx == m.cols - 1 || yy == m.rows - 1);
            v = ring ? 0 : (int)__ldg(m.data + (long long)yy * m.pitch + xx);
        }
        raw[ry][rx] = (unsigned char)v;
    }
    __syncthreads();
    for (int i = tid; i < (kErodeTH + 6) * kErodeTW; i += 256) {
        const int ry = i / kErodeTW, rx = i - ry * kErodeTW;
        int v = 255;
        SCB_UNROLL
        for (int d = 0; d < 7; ++d) v = min(v, (int)raw[ry][rx + d]);
        hm[ry][rx] = (unsigned char)v;
    }
    __syncthreads();
    for (int i = tid; i < kErodeTH * kErodeTW; i += 256) {
        const int ry = i / kErodeTW, rx = i - ry * kErodeTW;
        const int X = bx + rx, Y = by + ry;
        if (X < w && Y < h) {
            int v = 255;
            SCB_UNROLL
            for (int d = 0; d < 7; ++d) v = min(v, (int)hm[ry + d][rx]);
            E[(long long)Y * e_pitch + X] = (unsigned char)v;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// debug: blended gradients over the whole ROI (test hook for the 1e-4 intermediate checks)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gradients_dump_kernel(StencilSrc s, float* vx, float* vy /* [3][h][w] */, int mode) {
    const int X = blockIdx.x * 32 + (threadIdx.x & 31), Y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (X >= s.w || Y >= s.h) return;
    const int Xn = (X < s.w - 1) ? X + 1 : X - 1, Yn = (Y < s.h - 1) ? Y + 1 : Y - 1;  // BORDER_REFLECT_101
    const int e = __ldg(s.E + (long long)Y * s.e_pitch + X);
    const float inv255 = 1.0f / 255.0f;
    const float m = __fmul_rn((float)e, inv255), mi = __fmul_rn((float)(255 - e), inv255);
    for (int c = 0; c < 3; ++c) {
        float gxD, gyD, gxS, gyS;
        mode_gradients(s, mode, c, X, Y, Xn, Yn, gxD, gyD, gxS, gyS);
        vx[((size_t)c * s.h + Y) * s.w + X] = __fadd_rn(__fmul_rn(gxD, mi), __fmul_rn(gxS, m));
        vy[((size_t)c * s.h + Y) * s.w + X] = __fadd_rn(__fmul_rn(gyD, mi), __fmul_rn(gyS, m));
    }
}

}  // namespace scb
