// scb_gfft.cuh -- the production transform engine: pair-packed, group-synchronised shared-memory FFT.
//
// (1) Pair packing.  Every shared-memory element carries the same sample of TWO sequences
//     (the same channel of two adjacent lines): re-plane float2 (re_a, re_b), im-plane float2
//     (im_a, im_b).  A complex add is two FADD2, a complex multiply 2 FMUL2 + 2 FFMA2 -- Blackwell's
//     packed fp32x2 pipe, which is also the only way to reach the full FP32 rate on sm_100 -- and
//     every access is a conflict-free 64-bit LDS/STS (an AoS float4 layout costs 4 MOVs per STS.128).
// (2) Groups.  A CTA holds NG independent pairs (one per colour channel; NG = 1 from 512-point sequences up,
//     where blockIdx.y selects the channel -- measured on B200: three independent 1-group CTAs per SM finish
//     15-25 % sooner than one 3-group CTA, because the CTAs desynchronise and one CTA's exposed first-pass
//     load latency hides under the others' butterflies); pair g is owned by its own
//     group of G threads that synchronises on its own named barrier (bar.sync g+1, G).  The groups
//     drift apart, so one group's shared-memory bursts overlap another's FMA bursts.  With a single
//     CTA-wide barrier all warps load, compute and store in lockstep and the two pipes serialise
//     (measured: profiles/r1_packed_v2_ncu.txt).
// (3) Fusion.  The first forward pass takes its operands from a functor (stencil / global loads x chirp)
//     and the last inverse pass hands its results to a functor (chirp, scale, store / compose), so a
//     line makes 4 shared-memory round trips per convolution instead of 6; the column pass bridges
//     its two convolutions (inverse pass -> divide -> forward pass) in registers.
//
// Pass structure is that of scb_fft.cuh: radix R0 = 2^(LOG2M mod 4) (16 if 0) at L = M, radix 16 below,
// the two L = 16 passes fused around the product with the chirp spectrum; decimation in frequency
// forward and its mirror backward, so no reordering pass exists.
// Twiddles: per-pass tables of float4 (re_q, re_q+1, im_q, im_q+1) for q = 1,3,5,7, laid out [row][i]
// (coalesced 16-byte loads); q = 9..15 are fl(W^{i(q-8)} * W^{8i}) computed with packed arithmetic.
#pragma once

#include "scb_fft.cuh"
#include "scb_platform.h"

namespace scb {

struct P4 {
    float2 re, im;  // (re_a, re_b), (im_a, im_b)
};

SCB_D P4 p4_zero() { return P4{make_float2(0.f, 0.f), make_float2(0.f, 0.f)}; }
SCB_D P4 padd(const P4& a, const P4& b) { return P4{f2add(a.re, b.re), f2add(a.im, b.im)}; }
SCB_D P4 psub(const P4& a, const P4& b) { return P4{f2sub(a.re, b.re), f2sub(a.im, b.im)}; }
// a * (wr + i wi), same twiddle for both lanes
SCB_D P4 pmul(const P4& a, float wr, float wi) {
    const float2 r2 = f2dup(wr), i2 = f2dup(wi);
    return P4{f2fma(f2neg(a.im), i2, f2mul(a.re, r2)), f2fma(a.re, i2, f2mul(a.im, r2))};
}
// a * conj(wr + i wi)
SCB_D P4 pmulc(const P4& a, float wr, float wi) {
    const float2 r2 = f2dup(wr), i2 = f2dup(wi);
    return P4{f2fma(a.im, i2, f2mul(a.re, r2)), f2fma(f2neg(a.re), i2, f2mul(a.im, r2))};
}

// a * W16^K (forward W = exp(-2 pi i/16); INV conjugates), K in 0..7
template <int K, bool INV>
SCB_D P4 pmul_w16(const P4& a) {
    constexpr float C8 = 0.92387953251128674f, S8 = 0.38268343236508977f, H = 0.70710678118654752f;
    if constexpr (K == 0) {
        return a;
    } else if constexpr (K == 4) {  // -i (fwd) / +i (inv): free, the negation folds into the next add
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
        return pmul(a, wr, INV ? wi0 : -wi0);
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

// per-pass twiddle tables: rows of float4 (re_q, re_q+1, im_q, im_q+1), q = 1, 3, 5, 7 (as many as R needs)
SCB_HD constexpr int gtw_rows(int R) { return R >= 16 ? 4 : R / 2; }  // R=2 -> 1, 4 -> 2, 8 -> 4, 16 -> 4
SCB_HD constexpr int gtw_first_radix(int log2m) { return (log2m % 4 == 0) ? 16 : (1 << (log2m % 4)); }
// offset (in float4) of the table of the radix-16 pass with sub-length L
SCB_HD constexpr int gtw_offset16(int log2m, int L) {
    const int M = 1 << log2m, R0 = gtw_first_radix(log2m);
    int off = gtw_rows(R0) * (M / R0);
    for (int l = M / R0; l > L; l /= 16) off += 4 * (l / 16);
    return off;
}

SCB_HD constexpr int gtw_total_c(int log2m) { return gtw_offset16(log2m, 16); }

// ---- configuration ------------------------------------------------------------------------------
#ifndef SCB_NG1_FROM
#define SCB_NG1_FROM 9
#endif
#ifndef SCB_TW_SMEM_LOG2M
#define SCB_TW_SMEM_LOG2M 0
#endif
#ifndef SCB_G1_FROM
#define SCB_G1_FROM 12
#endif
template <int LOG2M>
struct GCfg {
    static_assert(LOG2M >= 5 && LOG2M <= 13, "group engine: convolution lengths 32 .. 8192");
    static constexpr int M = 1 << LOG2M;
    static constexpr int R0 = (LOG2M % 4 == 0) ? 16 : (1 << (LOG2M % 4));
    // threads per group: two radix-16 butterflies each; from 2^SCB_G1_FROM points up ONE butterfly each (twice the warps)
    static constexpr int G = (LOG2M >= SCB_G1_FROM) ? (M / 16) : ((M / 32 < 32) ? 32 : (M / 32));
    static constexpr int PADDED = M + (M >> 4);              // float2 elements per plane
    static constexpr int NG = (LOG2M < SCB_NG1_FROM) ? 3 : 1;  // groups (channels) per CTA; from 2^SCB_NG1_FROM points up: one pair per CTA, blockIdx.y = channel
    static constexpr int T = NG * G;
    static constexpr size_t DATA_BYTES = (size_t)NG * 2 * PADDED * sizeof(float2);
    // Twiddle tables live in shared memory when the CTA owns the SM anyway (M >= 4096: 17 KB / 81 KB
    // next to 209 KB / 139 KB of data).  At max carveout L1 is only ~24 KB and every table read would
    // otherwise pay L2 latency (ncu: long-scoreboard was the top stall of the fused passes).
    static constexpr bool TW_SMEM = (LOG2M == SCB_TW_SMEM_LOG2M);  // experiment switch (round 1, profiles/README.md)
    static constexpr int TW_F4 = gtw_total_c(LOG2M);
    static constexpr size_t SMEM = DATA_BYTES + (TW_SMEM ? (size_t)TW_F4 * sizeof(float4) : 0);
};

// named barrier of one group (id 1..15); NG == 1 uses the CTA barrier
template <int NG>
SCB_D void group_sync(int group, int nthreads) {
    if constexpr (NG == 1) {
        __syncthreads();
    } else {
#ifdef SCB_EMU
        ::emu::bar_sync(group + 1, nthreads);
#else
        asm volatile("bar.sync %0, %1;" ::"r"(group + 1), "r"(nthreads) : "memory");
#endif
    }
}

// twiddles of one butterfly: wre[q] + i wim[q] = W_L^{iq}, q = 1..R-1
template <int R, bool SMEM_TABLE = false>
SCB_D void load_twiddles(const float4* __restrict__ tws, int S, int i, float (&wre)[16], float (&wim)[16]) {
    constexpr int ROWS = gtw_rows(R);
    float4 t[ROWS];
    SCB_UNROLL
    for (int j = 0; j < ROWS; ++j) t[j] = SMEM_TABLE ? tws[j * S + i] : __ldg(tws + j * S + i);
    SCB_UNROLL
    for (int j = 0; j < ROWS; ++j) {
        wre[2 * j + 1] = t[j].x;
        wim[2 * j + 1] = t[j].z;
        if (2 * j + 2 < 16) {
            wre[2 * j + 2] = t[j].y;
            wim[2 * j + 2] = t[j].w;
        }
    }
    if constexpr (R == 16) {
        const float2 r8 = f2dup(wre[8]), i8 = f2dup(wim[8]);
        SCB_UNROLL
        for (int j = 0; j < 4; ++j) {  // (W^{2j+1}, W^{2j+2}) * W^8 -> (W^{2j+9}, W^{2j+10})
            const float2 re = make_float2(t[j].x, t[j].y), im = make_float2(t[j].z, t[j].w);
            const float2 pr = f2fma(f2neg(im), i8, f2mul(re, r8));
            const float2 pi = f2fma(re, i8, f2mul(im, r8));
            wre[2 * j + 9] = pr.x;
            wim[2 * j + 9] = pi.x;
            if (2 * j + 10 < 16) {
                wre[2 * j + 10] = pr.y;
                wim[2 * j + 10] = pi.y;
            }
        }
    }
}

struct Planes {
    float2* re;
    float2* im;
};
struct SmemIn {
    Planes pl;
    SCB_D P4 operator()(int j) const {
        const int p = padi(j);
        return P4{pl.re[p], pl.im[p]};
    }
};
struct SmemOut {
    Planes pl;
    SCB_D void operator()(int j, const P4& v) const {
        const int p = padi(j);
        pl.re[p] = v.re;
        pl.im[p] = v.im;
    }
};

// One radix-R pass of one group over sub-transforms of length L.  In/Out are functors indexed by the
// natural element position.
template <int LOG2M, int R, int L, bool INV, class In, class Out>
SCB_D void gpass(const float4* __restrict__ tws, int gtid, const In& in, const Out& out) {
    using C = GCfg<LOG2M>;
    constexpr int S = L / R;
    static_assert(S >= 2, "L == R passes are fused into gmiddle");
    for (int b = gtid; b < C::M / R; b += C::G) {
        const int i = b & (S - 1);
        const int base = (b / S) * L + i;
        float wre[16], wim[16];
        load_twiddles<R, GCfg<LOG2M>::TW_SMEM>(tws, S, i, wre, wim);
        P4 v[R];
        SCB_UNROLL
        for (int r = 0; r < R; ++r) v[r] = in(base + r * S);
        if (!INV) {
            PDft<R, false>::run(v);
            SCB_UNROLL
            for (int q = 1; q < R; ++q) v[q] = pmul(v[q], wre[q], wim[q]);
        } else {
            SCB_UNROLL
            for (int q = 1; q < R; ++q) v[q] = pmulc(v[q], wre[q], wim[q]);
            PDft<R, true>::run(v);
        }
        SCB_UNROLL
        for (int r = 0; r < R; ++r) out(base + r * S, v[r]);
    }
}

// Last inverse pass of one convolution, an element-wise bridge, first forward pass of the next
// convolution -- same radix, same positions, same thread: no shared-memory round trip, no barrier.
template <int LOG2M, class Bridge>
SCB_D void gpass_bridge(const float4* __restrict__ tws, int gtid, const Planes& pl, const Bridge& bridge) {
    using C = GCfg<LOG2M>;
    constexpr int R = C::R0, S = C::M / R;
    const SmemIn in{pl};
    const SmemOut out{pl};
    for (int b = gtid; b < C::M / R; b += C::G) {
        const int i = b;  // L == M: a single group of sub-transforms
        float wre[16], wim[16];
        load_twiddles<R, GCfg<LOG2M>::TW_SMEM>(tws, S, i, wre, wim);
        P4 v[R];
        SCB_UNROLL
        for (int r = 0; r < R; ++r) v[r] = in(i + r * S);
        SCB_UNROLL
        for (int q = 1; q < R; ++q) v[q] = pmulc(v[q], wre[q], wim[q]);
        PDft<R, true>::run(v);
        SCB_UNROLL
        for (int r = 0; r < R; ++r) v[r] = bridge(i + r * S, v[r]);
        PDft<R, false>::run(v);
        SCB_UNROLL
        for (int q = 1; q < R; ++q) v[q] = pmul(v[q], wre[q], wim[q]);
        SCB_UNROLL
        for (int r = 0; r < R; ++r) out(i + r * S, v[r]);
    }
}

// Last forward pass (L = 16), product with the chirp spectrum, first inverse pass: in registers.
template <int LOG2M>
SCB_D void gmiddle(const float2* __restrict__ bhat_t, int gtid, const Planes& pl) {
    using C = GCfg<LOG2M>;
    for (int b = gtid; b < C::M / 16; b += C::G) {
        float2 hq[16];
        SCB_UNROLL
        for (int q = 0; q < 16; ++q) hq[q] = __ldg(bhat_t + q * (C::M / 16) + b);
        const int p0 = padi(16 * b);  // 16 contiguous slots
        P4 v[16];
        SCB_UNROLL
        for (int r = 0; r < 16; ++r) v[r] = P4{pl.re[p0 + r], pl.im[p0 + r]};
        PDft<16, false>::run(v);
        SCB_UNROLL
        for (int q = 0; q < 16; ++q) v[q] = pmul(v[q], hq[q].x, hq[q].y);
        PDft<16, true>::run(v);
        SCB_UNROLL
        for (int r = 0; r < 16; ++r) {
            pl.re[p0 + r] = v[r].re;
            pl.im[p0 + r] = v[r].im;
        }
    }
}

template <int LOG2M, int L>
struct GFwd16 {
    SCB_D static void run(const float4* __restrict__ gtw, int gtid, int group, const Planes& pl) {
        if constexpr (L > 16) {
            gpass<LOG2M, 16, L, false>(gtw + gtw_offset16(LOG2M, L), gtid, SmemIn{pl}, SmemOut{pl});
            group_sync<GCfg<LOG2M>::NG>(group, GCfg<LOG2M>::G);
            GFwd16<LOG2M, L / 16>::run(gtw, gtid, group, pl);
        }
    }
};
template <int LOG2M, int L>
struct GInv16 {
    SCB_D static void run(const float4* __restrict__ gtw, int gtid, int group, const Planes& pl) {
        if constexpr (L > 16) {
            GInv16<LOG2M, L / 16>::run(gtw, gtid, group, pl);
            gpass<LOG2M, 16, L, true>(gtw + gtw_offset16(LOG2M, L), gtid, SmemIn{pl}, SmemOut{pl});
            group_sync<GCfg<LOG2M>::NG>(group, GCfg<LOG2M>::G);
        }
    }
};

// Everything between the first forward pass and the last inverse pass of one convolution.
// Entered with the first pass's stores issued (no barrier yet); leaves with the data of the last
// radix-16 inverse pass in shared memory and the group barrier passed.
template <int LOG2M>
__device__ __noinline__ void gconv_core(const float4* __restrict__ gtw, const float2* __restrict__ bhat_t, int gtid, int group, Planes pl) {
    using C = GCfg<LOG2M>;
    group_sync<C::NG>(group, C::G);
    GFwd16<LOG2M, C::M / C::R0>::run(gtw, gtid, group, pl);
    gmiddle<LOG2M>(bhat_t, gtid, pl);
    group_sync<C::NG>(group, C::G);
    GInv16<LOG2M, C::M / C::R0>::run(gtw, gtid, group, pl);
}

// First forward pass of a convolution whose operand is already in shared memory (written by other
// threads of the group: hence the barrier first).  In place: a butterfly rewrites exactly what it read.
template <int LOG2M>
SCB_D void gpass_first_from_smem(const float4* __restrict__ gtw, int gtid, int group, const Planes& pl) {
    using C = GCfg<LOG2M>;
    group_sync<C::NG>(group, C::G);
    gpass<LOG2M, C::R0, C::M, false>(gtw, gtid, SmemIn{pl}, SmemOut{pl});
}

}  // namespace scb
