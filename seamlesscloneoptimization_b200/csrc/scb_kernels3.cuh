// scb_kernels3.cuh -- the three line passes on the group engine (scb_gfft.cuh).  Production path for
// every ROI side up to 4098 (convolution length <= 8192); longer lines use scb_kernels.cuh.
//
// Work decomposition: a CTA owns TWO adjacent lines; group c of the CTA owns colour channel c of both
// lines as one packed pair (lane a = first line, lane b = second line).  For 8192-point lines a CTA
// holds a single group and blockIdx.y selects the channel.
// HBM layout, arithmetic and what each pass replaces in the reference: see scb_kernels.cuh.
#pragma once

#include "scb_gfft.cuh"
#include "scb_kernels.cuh"

namespace scb {

// single-channel variant of rhs_pixel (scb_kernels.cuh): same operations in the same order
SCB_D float rhs_pixel_c(const StencilSrc& s, int x, int y, int c) {
    const int X = x + 1, Y = y + 1;
    const unsigned char* d1 = s.D + (long long)Y * s.d_pitch + 3 * (X - 1) + c;
    const unsigned char* s1 = s.S + (long long)Y * s.s_pitch + 3 * (X - 1) + c;
    const unsigned char* e1 = s.E + (long long)Y * s.e_pitch + X;
    const float inv255 = 1.0f / 255.0f;
    const int ec = __ldg(e1), el = __ldg(e1 - 1), eu = __ldg(e1 - s.e_pitch);
    const float mc = __fmul_rn((float)ec, inv255), mic = __fmul_rn((float)(255 - ec), inv255);
    const float ml = __fmul_rn((float)el, inv255), mil = __fmul_rn((float)(255 - el), inv255);
    const float mu = __fmul_rn((float)eu, inv255), miu = __fmul_rn((float)(255 - eu), inv255);
    const float Dl = (float)__ldg(d1), Dc = (float)__ldg(d1 + 3), Dr = (float)__ldg(d1 + 6);
    const float Du = (float)__ldg(d1 + 3 - s.d_pitch), Dd = (float)__ldg(d1 + 3 + s.d_pitch);
    const float Sl = (float)__ldg(s1), Sc = (float)__ldg(s1 + 3), Sr = (float)__ldg(s1 + 6);
    const float Su = (float)__ldg(s1 + 3 - s.s_pitch), Sd = (float)__ldg(s1 + 3 + s.s_pitch);
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
    return __fsub_rn(lap, bnd);
}

// Im(c[k] * conv[k]) for both lanes
SCB_D float2 chirp_imag2(float2 ch, const P4& v) { return make_float2(ch.x * v.im.x + ch.y * v.re.x, ch.x * v.im.y + ch.y * v.re.y); }
// (a, b) real pair times the chirp -> packed complex element
SCB_D P4 chirp_times(float2 ch, float a, float b) { return P4{make_float2(a * ch.x, b * ch.x), make_float2(a * ch.y, b * ch.y)}; }

template <int LOG2M>
struct GroupIds {
    int gtid, group, ch, l0;
    bool has1;
    Planes pl;
    const float4* gtw;  // twiddle tables: shared-memory copy when GCfg::TW_SMEM, else the global tables
};

template <int LOG2M>
SCB_D GroupIds<LOG2M> group_ids(float2* smem, int first_line, int line_end, const float4* __restrict__ gtw_global) {
    using C = GCfg<LOG2M>;
    GroupIds<LOG2M> g;
    const int tid = threadIdx.x;
    g.group = (C::NG == 1) ? 0 : tid / C::G;
    g.gtid = tid - g.group * C::G;
    g.ch = (C::NG == 1) ? (int)blockIdx.y : g.group;
    g.l0 = first_line + 2 * (int)blockIdx.x;
    g.has1 = g.l0 + 1 < line_end;
    g.pl.re = smem + (size_t)g.group * 2 * C::PADDED;
    g.pl.im = g.pl.re + C::PADDED;
    g.gtw = gtw_global;
    if constexpr (C::TW_SMEM) {
        float4* tw = reinterpret_cast<float4*>(reinterpret_cast<unsigned char*>(smem) + C::DATA_BYTES);
        for (int i = tid; i < C::TW_F4; i += C::T) tw[i] = __ldg(gtw_global + i);
        __syncthreads();
        g.gtw = tw;
    }
    return g;
}

// ---------------------------------------------------------------------------------------------
// pass A: stencil -> forward DST-I along x.   grid = (ceil(rows/2), NG==1 ? 3 : 1)
// ---------------------------------------------------------------------------------------------
struct RowsFwd3Params {
    RowsFwdParams base;
    const float4* gtw;
    int y_end;
};

template <int LOG2M>
__global__ void __launch_bounds__(GCfg<LOG2M>::T) rows_fwd3_kernel(RowsFwd3Params pp) {
    using C = GCfg<LOG2M>;
    SCB_DYN_SMEM(float2, smem);
    const RowsFwdParams& p = pp.base;
    const GroupIds<LOG2M> g = group_ids<LOG2M>(smem, p.y0, pp.y_end, pp.gtw);
    const int n = p.nx, c = g.ch, y0 = g.l0, y1 = g.l0 + 1;
    const bool has1 = g.has1;
    auto load = [&](int j) -> P4 {
        if (j < 1 || j > n) return p4_zero();
        const float2 ch = __ldg(p.tx.chirp + j);
        float a, b = 0.f;
        if (p.rhs_in) {
            a = __ldg(p.rhs_in + ((size_t)c * p.ny + y0) * p.rhs_pitch + (j - 1));
            if (has1) b = __ldg(p.rhs_in + ((size_t)c * p.ny + y1) * p.rhs_pitch + (j - 1));
        } else {
            a = rhs_pixel_c(p.st, j - 1, y0, c);
            if (has1) b = rhs_pixel_c(p.st, j - 1, y1, c);
        }
        if (p.rhs_dump) {
            p.rhs_dump[((size_t)c * p.ny + y0) * p.nx + (j - 1)] = a;
            if (has1) p.rhs_dump[((size_t)c * p.ny + y1) * p.nx + (j - 1)] = b;
        }
        return chirp_times(ch, a, b);
    };
    auto store = [&](int k, const P4& v) {
        if (k < 1 || k > n) return;
        const float2 s = chirp_imag2(__ldg(p.tx.chirp + k), v);
        if (p.natural) {
            float* o = p.At + ((size_t)c * p.ny + y0) * p.nx + (k - 1);
            o[0] = -2.0f * s.x;
            if (has1) o[p.nx] = -2.0f * s.y;
            return;
        }
        float* o = p.At + ((size_t)c * p.nx + (k - 1)) * p.ny + y0;
        o[0] = -2.0f * s.x;  // OpenCV: Im of the odd-extension FFT = -2 sum x sin
        if (has1) o[1] = -2.0f * s.y;
    };
    gpass<LOG2M, C::R0, C::M, false>(g.gtw, g.gtid, load, SmemOut{g.pl});
    gconv_core<LOG2M>(g.gtw, p.tx.bhat_t, g.gtid, g.group, g.pl);
    gpass<LOG2M, C::R0, C::M, true>(g.gtw, g.gtid, SmemIn{g.pl}, store);
}

// ---------------------------------------------------------------------------------------------
// pass B: forward DST-I along y, eigenvalue division, inverse DST-I along y.   grid = (ceil(cols/2), ...)
// ---------------------------------------------------------------------------------------------
struct Cols3Params {
    ColsParams base;
    const float4* gtw;
    int x_end;
};

template <int LOG2M>
__global__ void __launch_bounds__(GCfg<LOG2M>::T) cols3_kernel(Cols3Params pp) {
    using C = GCfg<LOG2M>;
    SCB_DYN_SMEM(float2, smem);
    const ColsParams& p = pp.base;
    const GroupIds<LOG2M> g = group_ids<LOG2M>(smem, p.x0, pp.x_end, pp.gtw);
    const int n = p.ny, c = g.ch, k0 = g.l0, k1 = g.l0 + 1;
    const bool has1 = g.has1;
    const float* in0 = p.At + ((size_t)c * p.nx + k0) * p.ny;
    const float* in1 = in0 + p.ny;
    const float fx0 = __ldg(p.fx + k0), fx1 = has1 ? __ldg(p.fx + k1) : 0.f;
    auto load = [&](int j) -> P4 {
        if (j < 1 || j > n) return p4_zero();
        const float a = __ldg(in0 + (j - 1));
        const float b = has1 ? __ldg(in1 + (j - 1)) : 0.f;
        return chirp_times(__ldg(p.ty.chirp + j), a, b);
    };
    auto bridge = [&](int j, const P4& v) -> P4 {
        if (j < 1 || j > n) return p4_zero();
        const float2 ch = __ldg(p.ty.chirp + j);
        const float2 im = chirp_imag2(ch, v);
        float sa = -2.0f * im.x, sb = -2.0f * im.y;
        if (p.lowspec && (j - 1) < p.lowky) {  // exact low-frequency corner (float64 direct sums)
            if (k0 < p.lowkx) sa = __ldg(p.lowspec + ((size_t)c * p.lowkx + k0) * p.lowky + (j - 1));
            if (has1 && k1 < p.lowkx) sb = __ldg(p.lowspec + ((size_t)c * p.lowkx + k1) * p.lowky + (j - 1));
        }
        if (p.spec_dump) {
            p.spec_dump[((size_t)c * p.nx + k0) * p.ny + (j - 1)] = sa;
            if (has1) p.spec_dump[((size_t)c * p.nx + k1) * p.ny + (j - 1)] = sb;
        }
        // OpenCV: res /= (filter_X[i] + filter_Y[j] - 4), left to right in float32
        const float fy = __ldg(p.fy + (j - 1));
        const float qa = __fdiv_rn(sa, __fsub_rn(__fadd_rn(fx0, fy), 4.0f));
        const float qb = has1 ? __fdiv_rn(sb, __fsub_rn(__fadd_rn(fx1, fy), 4.0f)) : 0.f;
        return chirp_times(ch, qa, qb);
    };
    auto store = [&](int k, const P4& v) {
        if (k < 1 || k > n) return;
        const float2 s = chirp_imag2(__ldg(p.ty.chirp + k), v);
        float* o = p.Ct + ((size_t)c * p.ny + (k - 1)) * p.nx + k0;
        o[0] = s.x * p.inv_scale;
        if (has1) o[1] = s.y * p.inv_scale;
    };
    gpass<LOG2M, C::R0, C::M, false>(g.gtw, g.gtid, load, SmemOut{g.pl});
    gconv_core<LOG2M>(g.gtw, p.ty.bhat_t, g.gtid, g.group, g.pl);
    gpass_bridge<LOG2M>(g.gtw, g.gtid, g.pl, bridge);
    gconv_core<LOG2M>(g.gtw, p.ty.bhat_t, g.gtid, g.group, g.pl);
    gpass<LOG2M, C::R0, C::M, true>(g.gtw, g.gtid, SmemIn{g.pl}, store);
}

// ---------------------------------------------------------------------------------------------
// pass C: inverse DST-I along x, clamp, truncate, interleaved u8 store
// ---------------------------------------------------------------------------------------------
struct RowsInv3Params {
    RowsInvParams base;
    const float4* gtw;
    int y_end;
};

template <int LOG2M>
__global__ void __launch_bounds__(GCfg<LOG2M>::T) rows_inv3_kernel(RowsInv3Params pp) {
    using C = GCfg<LOG2M>;
    SCB_DYN_SMEM(float2, smem);
    const RowsInvParams& p = pp.base;
    const GroupIds<LOG2M> g = group_ids<LOG2M>(smem, p.y0, pp.y_end, pp.gtw);
    const int n = p.nx, c = g.ch, y0 = g.l0, y1 = g.l0 + 1;
    const bool has1 = g.has1;
    const float* in0 = p.Ct + ((size_t)c * p.ny + y0) * p.nx;
    const float* in1 = in0 + p.nx;
    auto load = [&](int j) -> P4 {
        if (j < 1 || j > n) return p4_zero();
        const float a = __ldg(in0 + (j - 1));
        const float b = has1 ? __ldg(in1 + (j - 1)) : 0.f;
        return chirp_times(__ldg(p.tx.chirp + j), a, b);
    };
    auto store = [&](int k, const P4& v) {
        if (k < 1 || k > n) return;
        const float2 s = chirp_imag2(__ldg(p.tx.chirp + k), v);
        const float ua = s.x * p.inv_scale, ub = s.y * p.inv_scale;
        if (p.u_dump) {
            const size_t d0 = p.transposed ? ((size_t)c * p.nx + (k - 1)) * p.ny + y0 : ((size_t)c * p.ny + y0) * p.nx + (k - 1);
            p.u_dump[d0] = ua;
            if (has1) p.u_dump[d0 + (p.transposed ? 1 : p.nx)] = ub;
        }
        unsigned char* o = p.transposed ? p.out + (long long)(k - 1) * p.out_pitch + 3 * y0 + c : p.out + (long long)y0 * p.out_pitch + 3 * (k - 1) + c;
        o[0] = compose_u8(ua);
        if (has1) o[p.transposed ? 3 : p.out_pitch] = compose_u8(ub);
    };
    gpass<LOG2M, C::R0, C::M, false>(g.gtw, g.gtid, load, SmemOut{g.pl});
    gconv_core<LOG2M>(g.gtw, p.tx.bhat_t, g.gtid, g.group, g.pl);
    gpass<LOG2M, C::R0, C::M, true>(g.gtw, g.gtid, SmemIn{g.pl}, store);
}


// =============================================================================================
// Quad variants: TWO real lines per complex sequence (so four lines per packed pair, twelve per CTA).
//
// With z = x_a + i x_b the chirp-z sum  T[k] = sum_j z[j] e^{i pi j k / N}  holds both sine transforms:
//     S_a[k] = (Im T[k] - Im T[-k]) / 2,      S_b[k] = (Re T[-k] - Re T[k]) / 2,
// i.e. with d = y[k] - y[M-k] (y = the circular convolution, T[+-k] = c[k] y[+-k]):
//     S_a[k] = Im(c[k] d) / 2,                S_b[k] = -Re(c[k] d) / 2.
// Outputs are needed for k in [-n, n] and the kernel h[m] = conj(c[m]) for m in [-2n, n-1], so the circular
// length must be >= 3n instead of 2n-1.  Whenever the power of two chosen for 2n-1 also covers 3n (41 % of all
// line lengths, e.g. n = 1339 or 1199 or 659 -> M = 4096 / 4096 / 2048) the same FFTs carry twice the lines:
// half the convolutions per image for one extra shared-memory round trip (y[k] and y[M-k] live in different
// butterflies of the last pass, so that pass writes shared memory and a short combine loop follows).
// =============================================================================================
template <int LOG2M>
struct QuadIds {
    int gtid, group, ch, l0, nl;  // nl = valid lines of this quad (1..4)
    Planes pl;
    const float4* gtw;
};

template <int LOG2M>
SCB_D QuadIds<LOG2M> quad_ids(float2* smem, int first_line, int line_end, const float4* __restrict__ gtw_global) {
    using C = GCfg<LOG2M>;
    QuadIds<LOG2M> g;
    const int tid = threadIdx.x;
    g.group = (C::NG == 1) ? 0 : tid / C::G;
    g.gtid = tid - g.group * C::G;
    g.ch = (C::NG == 1) ? (int)blockIdx.y : g.group;
    g.l0 = first_line + 4 * (int)blockIdx.x;
    g.nl = min(4, line_end - g.l0);
    g.pl.re = smem + (size_t)g.group * 2 * C::PADDED;
    g.pl.im = g.pl.re + C::PADDED;
    g.gtw = gtw_global;
    return g;
}

// (a0 + i a1) c[j] in lane a, (a2 + i a3) c[j] in lane b
SCB_D P4 quad_pack(float2 ch, float a0, float a1, float a2, float a3) { return pmul(P4{make_float2(a0, a2), make_float2(a1, a3)}, ch.x, ch.y); }
// c[k] (y[k] - y[M-k]):  .im = 2 S of lines 0 / 2,  -.re = 2 S of lines 1 / 3
template <int LOG2M>
SCB_D P4 quad_unpack(const Planes& pl, float2 ch, int k) {
    const int pk = padi(k), pm = padi(GCfg<LOG2M>::M - k);
    return pmul(psub(P4{pl.re[pk], pl.im[pk]}, P4{pl.re[pm], pl.im[pm]}), ch.x, ch.y);
}

// ---- pass A, four rows per group ----
template <int LOG2M>
__global__ void __launch_bounds__(GCfg<LOG2M>::T) rows_fwd4_kernel(RowsFwd3Params pp) {
    using C = GCfg<LOG2M>;
    SCB_DYN_SMEM(float2, smem);
    const RowsFwdParams& p = pp.base;
    const QuadIds<LOG2M> g = quad_ids<LOG2M>(smem, p.y0, pp.y_end, pp.gtw);
    const int n = p.nx, c = g.ch, nl = g.nl;
    const float* r0 = p.rhs_in + ((size_t)c * p.ny + g.l0) * p.rhs_pitch;
    auto load = [&](int j) -> P4 {
        if (j < 1 || j > n) return p4_zero();
        float a[4];
        SCB_UNROLL
        for (int i = 0; i < 4; ++i) a[i] = (i < nl) ? __ldg(r0 + (size_t)i * p.rhs_pitch + (j - 1)) : 0.f;
        return quad_pack(__ldg(p.tx.chirp + j), a[0], a[1], a[2], a[3]);
    };
    gpass<LOG2M, C::R0, C::M, false>(g.gtw, g.gtid, load, SmemOut{g.pl});
    gconv_core<LOG2M>(g.gtw, p.tx.bhat_q, g.gtid, g.group, g.pl);
    gpass<LOG2M, C::R0, C::M, true>(g.gtw, g.gtid, SmemIn{g.pl}, SmemOut{g.pl});
    group_sync<C::NG>(g.group, C::G);
    for (int k = 1 + g.gtid; k <= n; k += C::G) {
        const P4 t = quad_unpack<LOG2M>(g.pl, __ldg(p.tx.chirp + k), k);
        // OpenCV: Im of the odd-extension FFT = -2 S
        const size_t st = p.natural ? (size_t)p.nx : 1;
        float* o = p.natural ? p.At + ((size_t)c * p.ny + g.l0) * p.nx + (k - 1) : p.At + ((size_t)c * p.nx + (k - 1)) * p.ny + g.l0;
        o[0] = -t.im.x;
        if (nl > 1) o[st] = t.re.x;
        if (nl > 2) o[2 * st] = -t.im.y;
        if (nl > 3) o[3 * st] = t.re.y;
    }
}

// ---- pass B, four columns per group ----
template <int LOG2M>
__global__ void __launch_bounds__(GCfg<LOG2M>::T) cols4_kernel(Cols3Params pp) {
    using C = GCfg<LOG2M>;
    SCB_DYN_SMEM(float2, smem);
    const ColsParams& p = pp.base;
    const QuadIds<LOG2M> g = quad_ids<LOG2M>(smem, p.x0, pp.x_end, pp.gtw);
    const int n = p.ny, c = g.ch, nl = g.nl, k0 = g.l0;
    const float* in0 = p.At + ((size_t)c * p.nx + k0) * p.ny;
    float fxv[4];
    SCB_UNROLL
    for (int i = 0; i < 4; ++i) fxv[i] = (i < nl) ? __ldg(p.fx + k0 + i) : 0.f;
    auto load = [&](int j) -> P4 {
        if (j < 1 || j > n) return p4_zero();
        float a[4];
        SCB_UNROLL
        for (int i = 0; i < 4; ++i) a[i] = (i < nl) ? __ldg(in0 + (size_t)i * p.ny + (j - 1)) : 0.f;
        return quad_pack(__ldg(p.ty.chirp + j), a[0], a[1], a[2], a[3]);
    };
    gpass<LOG2M, C::R0, C::M, false>(g.gtw, g.gtid, load, SmemOut{g.pl});
    gconv_core<LOG2M>(g.gtw, p.ty.bhat_q, g.gtid, g.group, g.pl);
    gpass<LOG2M, C::R0, C::M, true>(g.gtw, g.gtid, SmemIn{g.pl}, SmemOut{g.pl});
    group_sync<C::NG>(g.group, C::G);
    // bridge: spectrum of the four columns -> refinement -> / eigenvalues -> operand of the inverse transform, in place.
    // Position k is read and rewritten, position M-k read and zeroed, by the same thread: no hazard.
    for (int pos = g.gtid; pos < C::M; pos += C::G) {
        if (pos >= 1 && pos <= n) {
            const int k = pos;
            const float2 ch = __ldg(p.ty.chirp + k);
            const P4 t = quad_unpack<LOG2M>(g.pl, ch, k);
            float s[4] = {-t.im.x, t.re.x, -t.im.y, t.re.y};  // -2 S: OpenCV's unnormalised forward transform
            const float fy = __ldg(p.fy + (k - 1));
            float q[4];
            SCB_UNROLL
            for (int i = 0; i < 4; ++i) {
                q[i] = 0.f;
                if (i < nl) {
                    if (p.lowspec && (k - 1) < p.lowky && k0 + i < p.lowkx) s[i] = __ldg(p.lowspec + ((size_t)c * p.lowkx + k0 + i) * p.lowky + (k - 1));
                    if (p.spec_dump) p.spec_dump[((size_t)c * p.nx + k0 + i) * p.ny + (k - 1)] = s[i];
                    // OpenCV: res /= (filter_X[i] + filter_Y[j] - 4), left to right in float32
                    q[i] = __fdiv_rn(s[i], __fsub_rn(__fadd_rn(fxv[i], fy), 4.0f));
                }
            }
            const P4 z = quad_pack(ch, q[0], q[1], q[2], q[3]);
            const int pk = padi(k), pm = padi(C::M - k);
            g.pl.re[pk] = z.re;
            g.pl.im[pk] = z.im;
            g.pl.re[pm] = make_float2(0.f, 0.f);
            g.pl.im[pm] = make_float2(0.f, 0.f);
        } else if (pos == 0 || (pos > n && pos < C::M - n)) {
            const int pz = padi(pos);
            g.pl.re[pz] = make_float2(0.f, 0.f);
            g.pl.im[pz] = make_float2(0.f, 0.f);
        }
    }
    // (gconv_core opens with the group barrier)
    gpass_first_from_smem<LOG2M>(g.gtw, g.gtid, g.group, g.pl);
    gconv_core<LOG2M>(g.gtw, p.ty.bhat_q, g.gtid, g.group, g.pl);
    gpass<LOG2M, C::R0, C::M, true>(g.gtw, g.gtid, SmemIn{g.pl}, SmemOut{g.pl});
    group_sync<C::NG>(g.group, C::G);
    const float hs = 0.5f * p.inv_scale;
    for (int k = 1 + g.gtid; k <= n; k += C::G) {
        const P4 t = quad_unpack<LOG2M>(g.pl, __ldg(p.ty.chirp + k), k);
        float* o = p.Ct + ((size_t)c * p.ny + (k - 1)) * p.nx + k0;
        o[0] = t.im.x * hs;
        if (nl > 1) o[1] = -t.re.x * hs;
        if (nl > 2) o[2] = t.im.y * hs;
        if (nl > 3) o[3] = -t.re.y * hs;
    }
}

// ---- pass C, four rows per group ----
template <int LOG2M>
__global__ void __launch_bounds__(GCfg<LOG2M>::T) rows_inv4_kernel(RowsInv3Params pp) {
    using C = GCfg<LOG2M>;
    SCB_DYN_SMEM(float2, smem);
    const RowsInvParams& p = pp.base;
    const QuadIds<LOG2M> g = quad_ids<LOG2M>(smem, p.y0, pp.y_end, pp.gtw);
    const int n = p.nx, c = g.ch, nl = g.nl, y0 = g.l0;
    const float* in0 = p.Ct + ((size_t)c * p.ny + y0) * p.nx;
    auto load = [&](int j) -> P4 {
        if (j < 1 || j > n) return p4_zero();
        float a[4];
        SCB_UNROLL
        for (int i = 0; i < 4; ++i) a[i] = (i < nl) ? __ldg(in0 + (size_t)i * p.nx + (j - 1)) : 0.f;
        return quad_pack(__ldg(p.tx.chirp + j), a[0], a[1], a[2], a[3]);
    };
    gpass<LOG2M, C::R0, C::M, false>(g.gtw, g.gtid, load, SmemOut{g.pl});
    gconv_core<LOG2M>(g.gtw, p.tx.bhat_q, g.gtid, g.group, g.pl);
    gpass<LOG2M, C::R0, C::M, true>(g.gtw, g.gtid, SmemIn{g.pl}, SmemOut{g.pl});
    group_sync<C::NG>(g.group, C::G);
    const float hs = 0.5f * p.inv_scale;
    for (int k = 1 + g.gtid; k <= n; k += C::G) {
        const P4 t = quad_unpack<LOG2M>(g.pl, __ldg(p.tx.chirp + k), k);
        const float u[4] = {t.im.x * hs, -t.re.x * hs, t.im.y * hs, -t.re.y * hs};
        unsigned char* o = p.transposed ? p.out + (long long)(k - 1) * p.out_pitch + 3 * y0 + c : p.out + (long long)y0 * p.out_pitch + 3 * (k - 1) + c;
        const long long ostep = p.transposed ? 3 : p.out_pitch;
        SCB_UNROLL
        for (int i = 0; i < 4; ++i) {
            if (i < nl) {
                if (p.u_dump) p.u_dump[p.transposed ? ((size_t)c * p.nx + (k - 1)) * p.ny + y0 + i : ((size_t)c * p.ny + y0 + i) * p.nx + (k - 1)] = u[i];
                o[(long long)i * ostep] = compose_u8(u[i]);
            }
        }
    }
}

}  // namespace scb
