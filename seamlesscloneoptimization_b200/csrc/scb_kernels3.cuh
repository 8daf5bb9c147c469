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
            p.u_dump[((size_t)c * p.ny + y0) * p.nx + (k - 1)] = ua;
            if (has1) p.u_dump[((size_t)c * p.ny + y1) * p.nx + (k - 1)] = ub;
        }
        unsigned char* o = p.out + (long long)y0 * p.out_pitch + 3 * (k - 1) + c;
        o[0] = compose_u8(ua);
        if (has1) o[p.out_pitch] = compose_u8(ub);
    };
    gpass<LOG2M, C::R0, C::M, false>(g.gtw, g.gtid, load, SmemOut{g.pl});
    gconv_core<LOG2M>(g.gtw, p.tx.bhat_t, g.gtid, g.group, g.pl);
    gpass<LOG2M, C::R0, C::M, true>(g.gtw, g.gtid, SmemIn{g.pl}, store);
}

}  // namespace scb
