// scb_fft.cuh -- shared-memory complex FFT of length M = 2^LOG2M used as a circular convolution
// engine for Bluestein's chirp-z form of the DST-I.
//
// Why chirp-z: OpenCV's Cloning::dst takes the DST-I of length n through a complex FFT of length
// 2(n+1) (reference dst(): /root/reference/seamlessClone-CUDA/seamlessClone_imp.cpp:1694-1811,
// there via cufftExecC2C).  n+1 is prime for the headline shapes (509, 4093), so a radix FFT of
// that length does not exist.  With  jk = (j^2 + k^2 - (k-j)^2)/2  and  N = n+1:
//     S[k] = sum_{j=1..n} x[j] sin(pi j k / N) = Im( c[k] * sum_j (x[j] c[j]) * conj(c[k-j]) ),
//     c[j] = exp(i pi j^2 / (2N)),
// a linear convolution, done as a circular one of any length M >= 2n-1: here the next power of two.
//
// Layout of one transform in shared memory: M float2, natural order, padded by one element every 16
// (index p lives at p + p/16) so that the radix-16 passes whose 16 operands are contiguous hit 16
// distinct bank pairs per half-warp.
//
// Pass structure (decimation in frequency forward, its exact mirror backward, so no bit reversal
// is ever materialised; the spectrum of the chirp is stored in the same permuted order):
//   forward :  radix R0 (L = M)  ->  radix 16 (L = M/R0, M/(16 R0), ... > 16)
//   middle  :  radix 16 at L = 16, multiply by Bhat, inverse radix 16 at L = 16  -- all in registers
//   inverse :  radix 16 (L = 256 ... M/R0)  ->  radix R0 (L = M)
// R0 = 2^(LOG2M mod 4) (16 when that is 0).  One thread owns one radix-16 butterfly (T = M/16
// threads per line) and runs it for every channel of the line, so each twiddle is fetched once per
// NCH butterflies.
#pragma once

#include "scb_platform.h"

namespace scb {

SCB_HD int padi(int p) { return p + (p >> 4); }

template <int LOG2M>
struct FftCfg {
    static_assert(LOG2M >= 5 && LOG2M <= 14, "supported transform lengths: 32 .. 16384");
    static constexpr int M = 1 << LOG2M;
    static constexpr int R0_LOG = (LOG2M % 4 == 0) ? 4 : (LOG2M % 4);
    static constexpr int R0 = 1 << R0_LOG;
    static constexpr int T = (M / 16 < 32) ? 32 : (M / 16);  // threads per line
    static constexpr int PADDED = M + (M >> 4);
};

SCB_HD float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
// a * conj(b)
SCB_HD float2 cmulc(float2 a, float2 b) { return make_float2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y); }
SCB_HD float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
SCB_HD float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }

// a * W16^K  (forward W = exp(-2 pi i /16); INV conjugates).  K in 0..7.
template <int K, bool INV>
SCB_HD float2 mul_w16(float2 a) {
    constexpr float C8 = 0.92387953251128674f;  // cos(pi/8)
    constexpr float S8 = 0.38268343236508977f;  // sin(pi/8)
    constexpr float H = 0.70710678118654752f;
    if constexpr (K == 0) {
        return a;
    } else if constexpr (K == 4) {
        return INV ? make_float2(-a.y, a.x) : make_float2(a.y, -a.x);
    } else if constexpr (K == 2) {
        return INV ? make_float2(H * (a.x - a.y), H * (a.x + a.y)) : make_float2(H * (a.x + a.y), H * (a.y - a.x));
    } else if constexpr (K == 6) {
        return INV ? make_float2(-H * (a.x + a.y), H * (a.x - a.y)) : make_float2(H * (a.y - a.x), -H * (a.x + a.y));
    } else {
        constexpr float wr = (K == 1) ? C8 : (K == 3) ? S8 : (K == 5) ? -S8 : -C8;
        constexpr float wi0 = (K == 1) ? S8 : (K == 3) ? C8 : (K == 5) ? C8 : S8;  // sin(2 pi K/16) > 0
        constexpr float wi = INV ? wi0 : -wi0;
        return make_float2(a.x * wr - a.y * wi, a.x * wi + a.y * wr);
    }
}

// In-register DFT of R points, natural order in and out:  v[q] <- sum_r v[r] W_R^{rq}.
template <int R, bool INV>
struct DftR;

template <bool INV>
struct DftR<1, INV> {
    SCB_HD static void run(float2*) {}
};

template <int R, bool INV, int K>
struct DftCombine {
    SCB_HD static void run(float2* v, const float2* e, const float2* o) {
        const float2 t = mul_w16<K*(16 / R), INV>(o[K]);
        v[K] = cadd(e[K], t);
        v[K + R / 2] = csub(e[K], t);
        if constexpr (K + 1 < R / 2) DftCombine<R, INV, K + 1>::run(v, e, o);
    }
};

template <int R, bool INV>
struct DftR {
    SCB_HD static void run(float2* v) {
        float2 e[R / 2], o[R / 2];
        SCB_UNROLL
        for (int k = 0; k < R / 2; ++k) {
            e[k] = v[2 * k];
            o[k] = v[2 * k + 1];
        }
        DftR<R / 2, INV>::run(e);
        DftR<R / 2, INV>::run(o);
        DftCombine<R, INV, 0>::run(v, e, o);
    }
};

// One radix-R pass over sub-transforms of length L, for NCH lines laid out PADDED apart.
//   forward: v <- DFT_R(v), then v[q] *= W_L^{iq};   inverse: v[q] *= conj(W_L^{iq}), then IDFT_R.
// tw[t] = exp(-2 pi i t / M).
template <int LOG2M, int R, int L, bool INV, int NCH>
SCB_D void radix_pass(float2* buf, const float2* __restrict__ tw, int tid) {
    using C = FftCfg<LOG2M>;
    constexpr int S = L / R;         // operand stride
    constexpr int TWS = C::M / L;    // twiddle table stride
    static_assert(S >= 2, "L == R passes are fused into fft_middle");
    for (int b = tid; b < C::M / R; b += C::T) {
        const int i = b & (S - 1);
        const int base = (b / S) * L + i;
        float2 w[R];
        SCB_UNROLL
        for (int q = 1; q < R; ++q) w[q] = __ldg(tw + (i * q) * TWS);
        SCB_UNROLL
        for (int c = 0; c < NCH; ++c) {
            float2* line = buf + c * C::PADDED;
            float2 v[R];
            SCB_UNROLL
            for (int r = 0; r < R; ++r) v[r] = line[padi(base + r * S)];
            if (!INV) {
                DftR<R, false>::run(v);
                SCB_UNROLL
                for (int q = 1; q < R; ++q) v[q] = cmul(v[q], w[q]);
            } else {
                SCB_UNROLL
                for (int q = 1; q < R; ++q) v[q] = cmulc(v[q], w[q]);
                DftR<R, true>::run(v);
            }
            SCB_UNROLL
            for (int r = 0; r < R; ++r) line[padi(base + r * S)] = v[r];
        }
    }
}

// Last forward pass (L = 16), pointwise product with the chirp spectrum, first inverse pass.
// bhat_t is stored operand-major: bhat_t[q * (M/16) + b] = Bhat[16 b + q] / M  (coalesced).
template <int LOG2M, int NCH>
SCB_D void fft_middle(float2* buf, const float2* __restrict__ bhat_t, int tid) {
    using C = FftCfg<LOG2M>;
    for (int b = tid; b < C::M / 16; b += C::T) {
        float2 hq[16];
        SCB_UNROLL
        for (int q = 0; q < 16; ++q) hq[q] = __ldg(bhat_t + q * (C::M / 16) + b);
        SCB_UNROLL
        for (int c = 0; c < NCH; ++c) {
            float2* line = buf + c * C::PADDED + padi(16 * b);  // 16 contiguous slots (16 b is a multiple of 16)
            float2 v[16];
            SCB_UNROLL
            for (int r = 0; r < 16; ++r) v[r] = line[r];
            DftR<16, false>::run(v);
            SCB_UNROLL
            for (int q = 0; q < 16; ++q) v[q] = cmul(v[q], hq[q]);
            DftR<16, true>::run(v);
            SCB_UNROLL
            for (int r = 0; r < 16; ++r) line[r] = v[r];
        }
    }
}

template <int LOG2M, int L, int NCH>
struct Fwd16 {
    SCB_D static void run(float2* buf, const float2* __restrict__ tw, int tid) {
        if constexpr (L > 16) {
            radix_pass<LOG2M, 16, L, false, NCH>(buf, tw, tid);
            __syncthreads();
            Fwd16<LOG2M, L / 16, NCH>::run(buf, tw, tid);
        }
    }
};
template <int LOG2M, int L, int NCH>
struct Inv16 {
    SCB_D static void run(float2* buf, const float2* __restrict__ tw, int tid) {
        if constexpr (L > 16) {
            Inv16<LOG2M, L / 16, NCH>::run(buf, tw, tid);
            radix_pass<LOG2M, 16, L, true, NCH>(buf, tw, tid);
            __syncthreads();
        }
    }
};

// Circular convolution of NCH lines (already in shared memory, barrier already passed) with the
// chirp whose permuted, 1/M-scaled spectrum is bhat_t.  Ends with a barrier.
template <int LOG2M, int NCH>
SCB_D void fft_convolve(float2* buf, const float2* __restrict__ tw, const float2* __restrict__ bhat_t, int tid) {
    using C = FftCfg<LOG2M>;
    radix_pass<LOG2M, C::R0, C::M, false, NCH>(buf, tw, tid);
    __syncthreads();
    Fwd16<LOG2M, C::M / C::R0, NCH>::run(buf, tw, tid);
    fft_middle<LOG2M, NCH>(buf, bhat_t, tid);
    __syncthreads();
    Inv16<LOG2M, C::M / C::R0, NCH>::run(buf, tw, tid);
    radix_pass<LOG2M, C::R0, C::M, true, NCH>(buf, tw, tid);
    __syncthreads();
}

}  // namespace scb
