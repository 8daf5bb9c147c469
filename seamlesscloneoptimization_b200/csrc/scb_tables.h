// scb_tables.h -- host-side (plan-time) tables, all evaluated in double and rounded once.
//
//  * chirp / chirp spectrum / twiddles for the Bluestein DST-I engine (scb_fft.cuh)
//  * sin rows for the exact low-frequency refinement
//  * OpenCV's float32 eigenvalue filters, reproduced operation by operation:
//      scale = CV_PI / (w - 1)  (double);  filter_X[i] = 2.0f * (float)cos(scale * (i + 1))
//    (the reference's initDSTMatrix_kernel, /root/reference/seamlessClone-CUDA/seamlessClone_imp.cpp:581-599,
//     uses a float PI and deviates from OpenCV; OpenCV's recipe is the parity target)
#pragma once

#include <cmath>
#include <complex>
#include <cstdint>
#include <vector>

namespace scb {

#ifndef SCB_LOWK
#define SCB_LOWK 4
#endif
static const int kLowK = SCB_LOWK;   // lowest frequencies per axis whose row sums are taken exactly (float64)
static const int kMaxLog2M = 14;     // longest convolution: 16384  ->  n <= 8192 unknowns per line
static const int kMinLog2M = 5;

struct HostF2 { float x, y; };

struct HostLenTab {
    int n = 0, log2m = 0, lowk = 0;
    std::vector<HostF2> chirp;   // n+1 : exp(+i pi j^2 / 2N)
    std::vector<HostF2> bhat_t;  // M   : permuted spectrum of conj chirp, / M, operand-major
    std::vector<HostF2> bhat_q;  // M   : the same over the support [-2n, n-1] (quad mode, scb_kernels3.cuh); empty when 3n > M
    std::vector<HostF2> tw;      // M   : exp(-2 pi i t / M)                       (scalar engine)
    std::vector<float> gtw;      // per-pass [row][i] float4 (re_q, re_q+1, im_q, im_q+1), q = 1,3,5,7  (group engine, scb_gfft.cuh)
    std::vector<double> sinlow;  // lowk x n : sin(pi (j+1)(k+1) / N)
};

inline int choose_log2m(int n) {
    int need = 2 * n - 1;
    int l = kMinLog2M;
    while ((1 << l) < need) ++l;
    return l;  // caller checks against kMaxLog2M
}

// quad mode packs two real lines into one complex sequence; it needs outputs k in [-n, n], hence M >= 3n
inline bool quad_ok(int n) { return n >= 2 && 3LL * n <= (1LL << choose_log2m(n)); }

inline int first_radix(int log2m) { return (log2m % 4 == 0) ? 16 : (1 << (log2m % 4)); }

typedef std::complex<double> cd;

// The exact forward pass sequence of scb_fft.cuh (radix R0 at L = M, then radix 16 down to L = 16),
// in double with directly evaluated DFT kernels.  Leaves the spectrum in the device's permuted order.
inline void host_dif_forward(std::vector<cd>& a, int log2m) {
    const int M = 1 << log2m;
    const double PI = 3.14159265358979323846;
    auto pass = [&](int R, int L) {
        const int S = L / R;
        std::vector<cd> wr(R * R), v(R), out(R);
        for (int r = 0; r < R; ++r)
            for (int q = 0; q < R; ++q) {
                int e = (r * q) % R;
                wr[r * R + q] = cd(std::cos(2 * PI * e / R), -std::sin(2 * PI * e / R));
            }
        for (int b = 0; b < M / R; ++b) {
            const int i = b % S, base = (b / S) * L + i;
            for (int r = 0; r < R; ++r) v[r] = a[base + r * S];
            for (int q = 0; q < R; ++q) {
                cd s(0, 0);
                for (int r = 0; r < R; ++r) s += v[r] * wr[r * R + q];
                long long e = (long long)i * q;  // < L
                s *= cd(std::cos(2 * PI * e / L), -std::sin(2 * PI * e / L));
                out[q] = s;
            }
            for (int q = 0; q < R; ++q) a[base + q * S] = out[q];
        }
    };
    const int R0 = first_radix(log2m);
    pass(R0, M);
    for (int L = M / R0; L >= 16; L /= 16) pass(16, L);
}

inline HostLenTab build_len_tab(int n) {
    HostLenTab t;
    const double PI = 3.14159265358979323846;
    t.n = n;
    t.log2m = choose_log2m(n);
    t.lowk = n < kLowK ? n : kLowK;
    const int M = 1 << t.log2m;
    const long long N = n + 1;
    std::vector<cd> c(n + 1);
    t.chirp.resize(n + 1);
    for (long long j = 0; j <= n; ++j) {
        long long ph = (j * j) % (4 * N);
        double ang = PI * (double)ph / (2.0 * (double)N);
        c[j] = cd(std::cos(ang), std::sin(ang));
        t.chirp[j] = HostF2{(float)c[j].real(), (float)c[j].imag()};
    }
    std::vector<cd> b(M, cd(0, 0));
    for (int m = 0; m <= n - 1; ++m) {
        b[m] = std::conj(c[m]);
        if (m) b[M - m] = std::conj(c[m]);
    }
    host_dif_forward(b, t.log2m);
    auto permute = [&](const std::vector<cd>& spec, std::vector<HostF2>& out) {
        out.resize(M);
        for (int blk = 0; blk < M / 16; ++blk)
            for (int q = 0; q < 16; ++q) {
                cd v = spec[16 * blk + q] / (double)M;
                out[(size_t)q * (M / 16) + blk] = HostF2{(float)v.real(), (float)v.imag()};
            }
    };
    permute(b, t.bhat_t);
    if (quad_ok(n)) {
        std::vector<cd> bq(M, cd(0, 0));
        auto chirp_at = [&](long long m) {  // c[m] = exp(+i pi m^2 / 2N), even in m, for |m| up to 2n
            long long ph = (m * m) % (4 * N);
            double ang = PI * (double)ph / (2.0 * (double)N);
            return cd(std::cos(ang), std::sin(ang));
        };
        for (long long m = 0; m <= n - 1; ++m) bq[m] = std::conj(chirp_at(m));
        for (long long m = 1; m <= 2LL * n; ++m) bq[M - m] = std::conj(chirp_at(m));
        host_dif_forward(bq, t.log2m);
        permute(bq, t.bhat_q);
    }
    t.tw.resize(M);
    for (int k = 0; k < M; ++k) {
        double ang = 2.0 * PI * (double)k / (double)M;
        t.tw[k] = HostF2{(float)std::cos(ang), (float)(-std::sin(ang))};
    }
    {   // per-pass coalesced tables, same order as gtw_offset16() in scb_gfft.cuh
        auto emit = [&](int R, int L) {
            const int S = L / R, rows = (R >= 16) ? 4 : R / 2;
            for (int j = 0; j < rows; ++j)
                for (int i = 0; i < S; ++i) {
                    float re[2] = {1.f, 1.f}, im[2] = {0.f, 0.f};
                    for (int e = 0; e < 2; ++e) {
                        const int q = 2 * j + 1 + e;
                        if (q >= R && !(R == 16 && q == 8)) continue;
                        if (q > 8) continue;
                        const double ang = 2.0 * PI * (double)((long long)i * q) / (double)L;
                        re[e] = (float)std::cos(ang);
                        im[e] = (float)(-std::sin(ang));
                    }
                    t.gtw.push_back(re[0]);
                    t.gtw.push_back(re[1]);
                    t.gtw.push_back(im[0]);
                    t.gtw.push_back(im[1]);
                }
        };
        const int R0 = first_radix(t.log2m);
        emit(R0, M);
        for (int L = M / R0; L > 16; L /= 16) emit(16, L);
    }
    t.sinlow.resize((size_t)t.lowk * n);
    for (int k = 0; k < t.lowk; ++k)
        for (int j = 0; j < n; ++j) {
            long long e = ((long long)(j + 1) * (k + 1)) % (2 * N);
            t.sinlow[(size_t)k * n + j] = std::sin(PI * (double)e / (double)N);
        }
    return t;
}

// filter_X / filter_Y of OpenCV's Cloning::initVariables; `extent` is the ROI width (or height).
inline std::vector<float> build_filter(int extent) {
    const double CV_PI_ = 3.1415926535897932384626433832795;
    std::vector<float> f(extent - 2);
    const double scale = CV_PI_ / (extent - 1);
    for (int i = 0; i < extent - 2; ++i) f[i] = 2.0f * (float)std::cos(scale * (i + 1));
    return f;
}

// theta[k] = acosh((4 - filter[k]) / 2): the tridiagonal engine's pivots are sinh ratios in theta (scb_tri.cuh).
// filter[k] is OpenCV's float32 value, taken as exact.
inline std::vector<double> build_theta(const std::vector<float>& filter) {
    std::vector<double> th(filter.size());
    for (size_t k = 0; k < filter.size(); ++k) {
        const double half_beta_m1 = (2.0 - (double)filter[k]) * 0.5;  // beta/2 - 1 >= 0, exact in double
        th[k] = half_beta_m1 > 0.0 ? std::log1p(half_beta_m1 + std::sqrt(half_beta_m1 * (half_beta_m1 + 2.0))) : 0.0;  // acosh(1 + e)
    }
    return th;
}

}  // namespace scb
