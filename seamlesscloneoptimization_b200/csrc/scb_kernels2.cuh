// scb_kernels2.cuh -- the three line passes on the pair-packed FFT engine (scb_pfft.cuh).
//
// Work decomposition: a pass over `nlines` lines has 3 * nlines sequences (line-major, channel
// minor).  Consecutive sequences are paired, a CTA owns NP consecutive pairs:
//     NP = 3  (M <= 4096): exactly two whole lines (6 sequences) per CTA, 209 KB of shared memory
//     NP = 1  (M  = 8192): two sequences per CTA (may straddle two lines), 139 KB
// M = 16384 lines (ROI side > 4098) stay on the scalar engine (scb_kernels.cuh), one sequence per CTA.
// HBM layout and arithmetic are identical to scb_kernels.cuh; see the header there for what each
// pass replaces in the reference.
#pragma once

#include "scb_kernels.cuh"
#include "scb_pfft.cuh"

namespace scb {

template <int NP>
struct SlotMap {
    int line[2 * NP];
    int ch[2 * NP];
    bool ok[2 * NP];
};

// sequences [first_line*3 ...) relative to this launch; line_end is exclusive
template <int NP>
SCB_D SlotMap<NP> make_slots(int first_line, int line_end) {
    SlotMap<NP> m;
    const int seq0 = 2 * NP * (int)blockIdx.x;
    SCB_UNROLL
    for (int q = 0; q < 2 * NP; ++q) {
        const int s = seq0 + q;
        m.line[q] = first_line + s / 3;
        m.ch[q] = s % 3;
        m.ok[q] = m.line[q] < line_end;
    }
    return m;
}

SCB_D float pick3(const float g[3], int c) { return c == 0 ? g[0] : (c == 1 ? g[1] : g[2]); }

// Im(c[k] * conv[k]) for both lanes of a pair element
SCB_D float2 chirp_imag(float2 ch, const float4& v) { return make_float2(ch.x * v.z + ch.y * v.x, ch.x * v.w + ch.y * v.y); }

// ---------------------------------------------------------------------------------------------
// pass A
// ---------------------------------------------------------------------------------------------
struct RowsFwd2Params {
    RowsFwdParams base;
    const float2* ptw;
    int y_end;
};

template <int LOG2M, int NP>
__global__ void __launch_bounds__(FftCfg<LOG2M>::T) rows_fwd2_kernel(RowsFwd2Params pp) {
    using C = FftCfg<LOG2M>;
    SCB_DYN_SMEM(float4, buf);
    const RowsFwdParams& p = pp.base;
    const int tid = threadIdx.x, n = p.nx;
    const SlotMap<NP> sm = make_slots<NP>(p.y0, pp.y_end);
    for (int j = tid; j < C::M; j += C::T) {
        float val[2 * NP];
        SCB_UNROLL
        for (int q = 0; q < 2 * NP; ++q) val[q] = 0.f;
        float2 ch = make_float2(0.f, 0.f);
        if (j >= 1 && j <= n) {
            ch = __ldg(p.tx.chirp + j);
            if constexpr (NP == 3) {  // two whole lines: one stencil evaluation per line and pixel
                SCB_UNROLL
                for (int l = 0; l < 2; ++l) {
                    if (sm.ok[3 * l]) {
                        float g[3];
                        const int y = sm.line[3 * l];
                        if (p.rhs_in) {
                            SCB_UNROLL
                            for (int c = 0; c < 3; ++c) g[c] = p.rhs_in[((size_t)c * p.ny + y) * p.nx + (j - 1)];
                        } else {
                            rhs_pixel(p.st, j - 1, y, g);
                        }
                        SCB_UNROLL
                        for (int c = 0; c < 3; ++c) {
                            val[3 * l + c] = g[c];
                            if (p.rhs_dump) p.rhs_dump[((size_t)c * p.ny + y) * p.nx + (j - 1)] = g[c];
                        }
                    }
                }
            } else {
                SCB_UNROLL
                for (int q = 0; q < 2 * NP; ++q) {
                    if (sm.ok[q]) {
                        float g[3];
                        const int y = sm.line[q];
                        if (p.rhs_in) {
                            SCB_UNROLL
                            for (int c = 0; c < 3; ++c) g[c] = p.rhs_in[((size_t)c * p.ny + y) * p.nx + (j - 1)];
                        } else {
                            rhs_pixel(p.st, j - 1, y, g);
                        }
                        val[q] = pick3(g, sm.ch[q]);
                        if (p.rhs_dump) p.rhs_dump[((size_t)sm.ch[q] * p.ny + y) * p.nx + (j - 1)] = val[q];
                    }
                }
            }
        }
        SCB_UNROLL
        for (int pr = 0; pr < NP; ++pr)
            buf[pr * C::PADDED + padi(j)] = make_float4(val[2 * pr] * ch.x, val[2 * pr + 1] * ch.x, val[2 * pr] * ch.y, val[2 * pr + 1] * ch.y);
    }
    __syncthreads();
    pfft_convolve<LOG2M, NP>(buf, pp.ptw, p.tx.bhat_t, tid);
    for (int k = tid + 1; k <= n; k += C::T) {
        const float2 ch = __ldg(p.tx.chirp + k);
        SCB_UNROLL
        for (int pr = 0; pr < NP; ++pr) {
            const float2 s = chirp_imag(ch, buf[pr * C::PADDED + padi(k)]);
            if (sm.ok[2 * pr]) p.At[((size_t)sm.ch[2 * pr] * p.nx + (k - 1)) * p.ny + sm.line[2 * pr]] = -2.0f * s.x;
            if (sm.ok[2 * pr + 1]) p.At[((size_t)sm.ch[2 * pr + 1] * p.nx + (k - 1)) * p.ny + sm.line[2 * pr + 1]] = -2.0f * s.y;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// pass B
// ---------------------------------------------------------------------------------------------
struct Cols2Params {
    ColsParams base;
    const float2* ptw;
    int x_end;
};

template <int LOG2M, int NP>
__global__ void __launch_bounds__(FftCfg<LOG2M>::T) cols2_kernel(Cols2Params pp) {
    using C = FftCfg<LOG2M>;
    SCB_DYN_SMEM(float4, buf);
    const ColsParams& p = pp.base;
    const int tid = threadIdx.x, n = p.ny;
    const SlotMap<NP> sm = make_slots<NP>(p.x0, pp.x_end);
    const float* in[2 * NP];
    float fxv[2 * NP];
    SCB_UNROLL
    for (int q = 0; q < 2 * NP; ++q) {
        in[q] = p.At + ((size_t)sm.ch[q] * p.nx + (sm.ok[q] ? sm.line[q] : 0)) * p.ny;
        fxv[q] = sm.ok[q] ? __ldg(p.fx + sm.line[q]) : 0.f;
    }
    for (int j = tid; j < C::M; j += C::T) {
        const bool inr = (j >= 1 && j <= n);
        const float2 ch = inr ? __ldg(p.ty.chirp + j) : make_float2(0.f, 0.f);
        SCB_UNROLL
        for (int pr = 0; pr < NP; ++pr) {
            const float a = (inr && sm.ok[2 * pr]) ? __ldg(in[2 * pr] + (j - 1)) : 0.f;
            const float b = (inr && sm.ok[2 * pr + 1]) ? __ldg(in[2 * pr + 1] + (j - 1)) : 0.f;
            buf[pr * C::PADDED + padi(j)] = make_float4(a * ch.x, b * ch.x, a * ch.y, b * ch.y);
        }
    }
    __syncthreads();
    pfft_convolve<LOG2M, NP>(buf, pp.ptw, p.ty.bhat_t, tid);
    for (int j = tid; j < C::M; j += C::T) {
        const bool inr = (j >= 1 && j <= n);
        float2 ch = make_float2(0.f, 0.f);
        float fyv = 0.f;
        if (inr) {
            ch = __ldg(p.ty.chirp + j);
            fyv = __ldg(p.fy + (j - 1));
        }
        SCB_UNROLL
        for (int pr = 0; pr < NP; ++pr) {
            float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
            if (inr) {
                const float2 im = chirp_imag(ch, buf[pr * C::PADDED + padi(j)]);
                float s[2] = {-2.0f * im.x, -2.0f * im.y};
                float qv[2];
                SCB_UNROLL
                for (int e = 0; e < 2; ++e) {
                    const int q = 2 * pr + e;
                    if (sm.ok[q]) {
                        const int kx = sm.line[q], c = sm.ch[q];
                        if (p.lowspec && kx < p.lowkx && (j - 1) < p.lowky) s[e] = __ldg(p.lowspec + ((size_t)c * p.lowkx + kx) * p.lowky + (j - 1));
                        if (p.spec_dump) p.spec_dump[((size_t)c * p.nx + kx) * p.ny + (j - 1)] = s[e];
                        // OpenCV: res /= (filter_X[i] + filter_Y[j] - 4), left to right in float32
                        qv[e] = __fdiv_rn(s[e], __fsub_rn(__fadd_rn(fxv[q], fyv), 4.0f));
                    } else {
                        qv[e] = 0.f;
                    }
                }
                o = make_float4(qv[0] * ch.x, qv[1] * ch.x, qv[0] * ch.y, qv[1] * ch.y);
            }
            buf[pr * C::PADDED + padi(j)] = o;
        }
    }
    __syncthreads();
    pfft_convolve<LOG2M, NP>(buf, pp.ptw, p.ty.bhat_t, tid);
    for (int k = tid + 1; k <= n; k += C::T) {
        const float2 ch = __ldg(p.ty.chirp + k);
        SCB_UNROLL
        for (int pr = 0; pr < NP; ++pr) {
            const float2 s = chirp_imag(ch, buf[pr * C::PADDED + padi(k)]);
            if (sm.ok[2 * pr]) p.Ct[((size_t)sm.ch[2 * pr] * p.ny + (k - 1)) * p.nx + sm.line[2 * pr]] = s.x * p.inv_scale;
            if (sm.ok[2 * pr + 1]) p.Ct[((size_t)sm.ch[2 * pr + 1] * p.ny + (k - 1)) * p.nx + sm.line[2 * pr + 1]] = s.y * p.inv_scale;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// pass C
// ---------------------------------------------------------------------------------------------
struct RowsInv2Params {
    RowsInvParams base;
    const float2* ptw;
    int y_end;
};

template <int LOG2M, int NP>
__global__ void __launch_bounds__(FftCfg<LOG2M>::T) rows_inv2_kernel(RowsInv2Params pp) {
    using C = FftCfg<LOG2M>;
    SCB_DYN_SMEM(float4, buf);
    const RowsInvParams& p = pp.base;
    const int tid = threadIdx.x, n = p.nx;
    const SlotMap<NP> sm = make_slots<NP>(p.y0, pp.y_end);
    const float* in[2 * NP];
    SCB_UNROLL
    for (int q = 0; q < 2 * NP; ++q) in[q] = p.Ct + ((size_t)sm.ch[q] * p.ny + (sm.ok[q] ? sm.line[q] : 0)) * p.nx;
    for (int j = tid; j < C::M; j += C::T) {
        const bool inr = (j >= 1 && j <= n);
        const float2 ch = inr ? __ldg(p.tx.chirp + j) : make_float2(0.f, 0.f);
        SCB_UNROLL
        for (int pr = 0; pr < NP; ++pr) {
            const float a = (inr && sm.ok[2 * pr]) ? __ldg(in[2 * pr] + (j - 1)) : 0.f;
            const float b = (inr && sm.ok[2 * pr + 1]) ? __ldg(in[2 * pr + 1] + (j - 1)) : 0.f;
            buf[pr * C::PADDED + padi(j)] = make_float4(a * ch.x, b * ch.x, a * ch.y, b * ch.y);
        }
    }
    __syncthreads();
    pfft_convolve<LOG2M, NP>(buf, pp.ptw, p.tx.bhat_t, tid);
    for (int k = tid + 1; k <= n; k += C::T) {
        const float2 ch = __ldg(p.tx.chirp + k);
        SCB_UNROLL
        for (int pr = 0; pr < NP; ++pr) {
            const float2 s = chirp_imag(ch, buf[pr * C::PADDED + padi(k)]);
            const float u[2] = {s.x * p.inv_scale, s.y * p.inv_scale};
            SCB_UNROLL
            for (int e = 0; e < 2; ++e) {
                const int q = 2 * pr + e;
                if (sm.ok[q]) {
                    const int y = sm.line[q], c = sm.ch[q];
                    if (p.u_dump) p.u_dump[((size_t)c * p.ny + y) * p.nx + (k - 1)] = u[e];
                    p.out[(long long)y * p.out_pitch + 3 * (k - 1) + c] = compose_u8(u[e]);
                }
            }
        }
    }
}

}  // namespace scb
