// scb_pfft.cuh -- pair-packed shared-memory FFT (the production engine).
//
// Same pass structure as scb_fft.cuh (DIF forward, mirrored DIT inverse, chirp spectrum applied in
// registers between the two L = 16 passes), but every shared-memory element holds the same
// frequency sample of TWO independent sequences:
//        float4 { re_a, re_b, im_a, im_b }
// so that each complex add is one FADD2 pair and each complex multiply 2 FMUL2 + 2 FFMA2 -- Blackwell's
// packed fp32x2 pipe -- and each operand moves with one LDS.128/STS.128.  Versus the scalar engine this
// halves both the FP and the LSU instruction counts per sequence (ncu, profiles/r1: the scalar
// passes are issue-bound at 156 M warp instructions for the column pass of the 4K config).
//
// Twiddles: per-pass tables laid out [q][i] so a warp reads consecutive entries (coalesced), only
// q = 1..8 are stored; W^{iq} for q = 9..15 is fl(W^{i(q-8)} * W^{8i}) (max error 1.2e-7, measured).
// One thread owns one butterfly index and runs it for all NP pairs of its CTA, so twiddles are
// fetched once per NP pair-butterflies.
#pragma once

#include "scb_fft.cuh"
#include "scb_platform.h"

namespace scb {

struct P4 {
    float2 re, im;  // (re_a, re_b), (im_a, im_b)
};

SCB_D P4 p4_load(const float4* p) {
    const float4 t = *p;
    P4 r;
    r.re = make_float2(t.x, t.y);
    r.im = make_float2(t.z, t.w);
    return r;
}
SCB_D void p4_store(float4* p, const P4& v) { *p = make_float4(v.re.x, v.re.y, v.im.x, v.im.y); }
SCB_D P4 padd(const P4& a, const P4& b) { return P4{f2add(a.re, b.re), f2add(a.im, b.im)}; }
SCB_D P4 psub(const P4& a, const P4& b) { return P4{f2sub(a.re, b.re), f2sub(a.im, b.im)}; }
// a * w   (w = wr + i wi, the same for both lanes)
SCB_D P4 pmul(const P4& a, float2 w) {
    const float2 wr = f2dup(w.x), wi = f2dup(w.y);
    return P4{f2fma(f2neg(a.im), wi, f2mul(a.re, wr)), f2fma(a.re, wi, f2mul(a.im, wr))};
}
// a * conj(w)
SCB_D P4 pmulc(const P4& a, float2 w) {
    const float2 wr = f2dup(w.x), wi = f2dup(w.y);
    return P4{f2fma(a.im, wi, f2mul(a.re, wr)), f2fma(f2neg(a.re), wi, f2mul(a.im, wr))};
}

// a * W16^K (forward W = exp(-2 pi i/16); INV conjugates), K in 0..7
template <int K, bool INV>
SCB_D P4 pmul_w16(const P4& a) {
    constexpr float C8 = 0.92387953251128674f, S8 = 0.38268343236508977f, H = 0.70710678118654752f;
    if constexpr (K == 0) {
        return a;
    } else if constexpr (K == 4) {  // -i (fwd) / +i (inv)
        return INV ? P4{f2neg(a.im), a.re} : P4{a.im, f2neg(a.re)};
    } else if constexpr (K == 2) {
        const float2 h = f2dup(H);
        return INV ? P4{f2mul(f2sub(a.re, a.im), h), f2mul(f2add(a.re, a.im), h)} : P4{f2mul(f2add(a.re, a.im), h), f2mul(f2sub(a.im, a.re), h)};
    } else if constexpr (K == 6) {
        const float2 h = f2dup(H), nh = f2dup(-H);
        return INV ? P4{f2mul(f2add(a.re, a.im), nh), f2mul(f2sub(a.re, a.im), h)} : P4{f2mul(f2sub(a.im, a.re), h), f2mul(f2add(a.re, a.im), nh)};
    } else {
        constexpr float wr = (K == 1) ? C8 : (K == 3) ? S8 : (K == 5) ? -S8 : -C8;
        constexpr float wi0 = (K == 1) ? S8 : (K == 3) ? C8 : (K == 5) ? C8 : S8;  // sin(2 pi K / 16)
        constexpr float wi = INV ? wi0 : -wi0;                                     // twiddle = wr + i wi
        const float2 r2 = f2dup(wr), i2 = f2dup(wi);
        return P4{f2fma(f2neg(a.im), i2, f2mul(a.re, r2)), f2fma(a.re, i2, f2mul(a.im, r2))};
    }
}

template <int R, bool INV>
struct PDft;
template <bool INV>
struct PDft<1, INV> {
    SCB_D static void run(P4*) {}
};
template <int R, bool INV, int K>
struct PDftCombine {
    SCB_D static void run(P4* v, const P4* e, const P4* o) {
        const P4 t = pmul_w16<K*(16 / R), INV>(o[K]);
        v[K] = padd(e[K], t);
        v[K + R / 2] = psub(e[K], t);
        if constexpr (K + 1 < R / 2) PDftCombine<R, INV, K + 1>::run(v, e, o);
    }
};
template <int R, bool INV>
struct PDft {
    SCB_D static void run(P4* v) {
        P4 e[R / 2], o[R / 2];
        SCB_UNROLL
        for (int k = 0; k < R / 2; ++k) {
            e[k] = v[2 * k];
            o[k] = v[2 * k + 1];
        }
        PDft<R / 2, INV>::run(e);
        PDft<R / 2, INV>::run(o);
        PDftCombine<R, INV, 0>::run(v, e, o);
    }
};

// ---- per-pass twiddle tables: [q-1][i], q = 1..min(R-1, 8), i < L/R -------------------------------
SCB_HD constexpr int ptw_rows(int R) { return (R - 1 < 8) ? (R - 1) : 8; }
SCB_HD constexpr int ptw_first_radix(int log2m) { return (log2m % 4 == 0) ? 16 : (1 << (log2m % 4)); }
// offset (in float2) of the table of the radix-16 pass with sub-length L (L <= M / R0)
SCB_HD constexpr int ptw_offset16(int log2m, int L) {
    const int M = 1 << log2m, R0 = ptw_first_radix(log2m);
    int off = ptw_rows(R0) * (M / R0);
    for (int l = M / R0; l > L; l /= 16) off += 8 * (l / 16);
    return off;
}
SCB_HD constexpr int ptw_total(int log2m) { return ptw_offset16(log2m, 16); }

template <int LOG2M, int R, int L, bool INV, int NP>
SCB_D void pradix_pass(float4* buf, const float2* __restrict__ tws, int tid) {
    using C = FftCfg<LOG2M>;
    constexpr int S = L / R;
    constexpr int NL = ptw_rows(R);
    static_assert(S >= 2, "L == R passes are fused into pfft_middle");
    for (int b = tid; b < C::M / R; b += C::T) {
        const int i = b & (S - 1);
        const int base = (b / S) * L + i;
        float2 w[R];
        SCB_UNROLL
        for (int q = 1; q <= NL; ++q) w[q] = __ldg(tws + (q - 1) * S + i);
        if constexpr (R == 16) {
            SCB_UNROLL
            for (int q = 9; q < 16; ++q) w[q] = cmul(w[q - 8], w[8]);
        }
#ifndef SCB_EMU
#pragma unroll 1
#endif
        for (int p = 0; p < NP; ++p) {
            float4* line = buf + p * C::PADDED;
            P4 v[R];
            SCB_UNROLL
            for (int r = 0; r < R; ++r) v[r] = p4_load(line + padi(base + r * S));
            if (!INV) {
                PDft<R, false>::run(v);
                SCB_UNROLL
                for (int q = 1; q < R; ++q) v[q] = pmul(v[q], w[q]);
            } else {
                SCB_UNROLL
                for (int q = 1; q < R; ++q) v[q] = pmulc(v[q], w[q]);
                PDft<R, true>::run(v);
            }
            SCB_UNROLL
            for (int r = 0; r < R; ++r) p4_store(line + padi(base + r * S), v[r]);
        }
    }
}

template <int LOG2M, int NP>
SCB_D void pfft_middle(float4* buf, const float2* __restrict__ bhat_t, int tid) {
    using C = FftCfg<LOG2M>;
    for (int b = tid; b < C::M / 16; b += C::T) {
        float2 hq[16];
        SCB_UNROLL
        for (int q = 0; q < 16; ++q) hq[q] = __ldg(bhat_t + q * (C::M / 16) + b);
#ifndef SCB_EMU
#pragma unroll 1
#endif
        for (int p = 0; p < NP; ++p) {
            float4* line = buf + p * C::PADDED + padi(16 * b);
            P4 v[16];
            SCB_UNROLL
            for (int r = 0; r < 16; ++r) v[r] = p4_load(line + r);
            PDft<16, false>::run(v);
            SCB_UNROLL
            for (int q = 0; q < 16; ++q) v[q] = pmul(v[q], hq[q]);
            PDft<16, true>::run(v);
            SCB_UNROLL
            for (int r = 0; r < 16; ++r) p4_store(line + r, v[r]);
        }
    }
}

template <int LOG2M, int L, int NP>
struct PFwd16 {
    SCB_D static void run(float4* buf, const float2* __restrict__ ptw, int tid) {
        if constexpr (L > 16) {
            pradix_pass<LOG2M, 16, L, false, NP>(buf, ptw + ptw_offset16(LOG2M, L), tid);
            __syncthreads();
            PFwd16<LOG2M, L / 16, NP>::run(buf, ptw, tid);
        }
    }
};
template <int LOG2M, int L, int NP>
struct PInv16 {
    SCB_D static void run(float4* buf, const float2* __restrict__ ptw, int tid) {
        if constexpr (L > 16) {
            PInv16<LOG2M, L / 16, NP>::run(buf, ptw, tid);
            pradix_pass<LOG2M, 16, L, true, NP>(buf, ptw + ptw_offset16(LOG2M, L), tid);
            __syncthreads();
        }
    }
};

// Circular convolution of 2*NP sequences with the chirp.  Called after a barrier, ends with one.
// Not inlined: the column pass calls it twice and the instruction cache is the scarce resource.
template <int LOG2M, int NP>
__device__ __noinline__ void pfft_convolve(float4* buf, const float2* __restrict__ ptw, const float2* __restrict__ bhat_t, int tid) {
    using C = FftCfg<LOG2M>;
    pradix_pass<LOG2M, C::R0, C::M, false, NP>(buf, ptw, tid);
    __syncthreads();
    PFwd16<LOG2M, C::M / C::R0, NP>::run(buf, ptw, tid);
    pfft_middle<LOG2M, NP>(buf, bhat_t, tid);
    __syncthreads();
    PInv16<LOG2M, C::M / C::R0, NP>::run(buf, ptw, tid);
    pradix_pass<LOG2M, C::R0, C::M, true, NP>(buf, ptw, tid);
    __syncthreads();
}

}  // namespace scb
