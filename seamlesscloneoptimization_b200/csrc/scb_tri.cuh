// scb_tri.cuh -- the column half of the Poisson solve as a tridiagonal solve (production engine, SCB_ENGINE_TRI).
//
// OpenCV solves  u = IDST2( DST2(g) / (fx[k] + fy[l] - 4) )  (Cloning::solve; the reference mirrors it in
// /root/reference/seamlessClone-CUDA/seamlessClone_imp.cpp:1814-1896: dst(), updateUij_kernel_fft :1642-1669, dst()).
// Along y, "DST -> divide by (fx[k] + fy[l] - 4) -> inverse DST" of one spectral column k is, in exact arithmetic,
//     Ct_k = -(T + (fx[k] - 4) I)^-1 A_k ,      T = tridiag(1, 0, 1)   (eigenvalues fy[l] = 2 cos(pi (l+1) / Ny))
// i.e. the solution of  M Ct_k = A_k  with the symmetric Toeplitz M = tridiag(-1, beta_k, -1), beta_k = 4 - fx[k] > 2:
// strictly diagonally dominant, so plain LU without pivoting (Thomas) is stable, costs ~4 flops per unknown and
// is HBM-bound -- against two Bluestein convolutions (4 FFTs of length >= 2 ny) per column for the spectral route.
// The row transforms stay FFTs (scb_kernels3.cuh): one axis has to be diagonalised.
//
// Parity with OpenCV.  OpenCV's denominators are float32: fl(fl(fx[k] + fy[l]) - 4), with fx, fy themselves rounded
// to float32.  M uses beta_k = 4 - fx[k] with OpenCV's float32 fx[k] (exact), and the exact fy[l].  The two differ by
// ~2.4e-7 absolute, which matters only where the denominator is tiny: at the lowest frequencies (measured against
// cv2.seamlessClone, /tmp-free restatement in tests/test_tri_model.py: exact tridiagonal solve for k >= 32 leaves the
// byte-exact fraction at the float64 floor; for all k it drops to 41 % at 4K).  So:
//   * columns k < kTriLowK (one warp per channel) are solved in float64 (their systems have condition ~ (N/pi k)^2),
//     from the exact float64 row sums where those exist (k < 8: lowfreq_rows_kernel), and
//   * tri_lowcorr_kernel adds, for k < kTriLowK and l < kTriLowL, the difference between OpenCV's float32 denominator
//     and the exact one:  Ct_k += -(2/Ny) sum_l sin_l (1/den32[k][l] - 1/den_exact[k][l]) <sin_l, A_k>   (float64 sums).
// That block subsumes the 8 x 8 exact low-frequency corner of the FFT engine.
//
// LU factors in closed form.  With beta = 2 cosh(theta), rho = exp(-theta), the pivots of M are
//     p_j = sinh((j+1) theta) / sinh(j theta),   m_j = 1 / p_j = rho (1 - rho^2j) / (1 - rho^(2j+2)),   j = 1..n,
// evaluated in float64 at plan time (tri_table_kernel; cached per (w, h) in the context) and rounded once.
// The pivot sequence is the same from either end of the column, so each column is eliminated from BOTH ends towards
// the middle (two warps per 32 columns: twice the parallelism, half the dependent chain), the two halves meet in a
// 2 x 2 system and substitute back outwards.
#pragma once

#include <cstring>

#include "scb_kernels.cuh"
#include "scb_platform.h"

namespace scb {

static constexpr int kTriLowK = 32;  // spectral columns solved in float64 and corrected to OpenCV's float32 denominators
static constexpr int kTriLowL = 32;  // ... for this many lowest frequencies along y
static constexpr int kTriCols = 32;  // columns per CTA (one warp per column end)
static constexpr int kTriU = 16;     // rows per cp.async stage

struct TriTabDev {
    const float* m32;    // [ny][pm]   m_(d+1) of column k, d = distance from the column end
    int pm;
    const double* m64;   // [ny][kTriLowK]
    const double* theta; // [nx]
};

struct TriTableParams {
    const double* theta;  // [nx]  acosh((4 - fx[k]) / 2)
    int nx, ny, pm;
    float* m32;
    double* m64;
};

// m_(d+1) = rho (1 - rho^(2d+2)) / (1 - rho^(2d+4))
__global__ void __launch_bounds__(256) tri_table_kernel(TriTableParams p) {
    const long long total = (long long)p.ny * p.pm;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int d = (int)(i / p.pm), k = (int)(i - (long long)d * p.pm);
        double m = 0.0;
        if (k < p.nx) {
            const double th = p.theta[k];
            if (th < 1e-10)
                m = (double)(d + 1) / (double)(d + 2);
            else
                m = exp(-th) * (expm1(-2.0 * (d + 1) * th) / expm1(-2.0 * (d + 2) * th));
        }
        p.m32[i] = (float)m;
        if (k < kTriLowK) p.m64[(size_t)d * kTriLowK + k] = m;
    }
}

struct TriSolveParams {
    TriTabDev tab;
    int nx, ny;
    const float* A;    // [3][ny][nx]  row-transformed RHS (OpenCV scale: -2 sum g sin)
    float* Ct;         // [3][ny][nx]
    const double* R;   // [3][lowkx][ny] exact row sums (A = -2 R) for k < lowkx, or null
    int lowkx;
    double* Y64;       // [3][ny][kTriLowK] float64 work / result columns k < kTriLowK
    int x0, x1;        // columns of this launch (multiples of kTriCols except the end)
};

// cp.async (LDGSTS): global -> shared without a register stop, so a warp keeps D x kTriU rows in flight
// while its dependent chain works on the oldest stage.  Each lane copies, and later reads, only its own column:
// no cross-lane synchronisation, cp.async.wait_group is enough.
#ifdef SCB_EMU
template <int B>
SCB_D void cp_async(void* smem, const void* g) { std::memcpy(smem, g, B); }
SCB_D void cp_async_commit() {}
template <int N>
SCB_D void cp_async_wait() {}
#else
template <int B>
SCB_D void cp_async(void* smem, const void* g) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(s), "l"(g), "n"(B) : "memory");
}
SCB_D void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
SCB_D void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
#endif

static constexpr int kTriRingBytesPerWarp = 32768;  // D stages x {operand, factor} x kTriU rows x 32 lanes x sizeof(T)
static constexpr size_t kTriSmemBytes = 2 * kTriRingBytesPerWarp;

// One column end: eliminates towards the middle, meets the other end, substitutes back.  T = float or double.
//   issue_a(slot, y) / read_a(slot)  RHS element of row y          issue_m(slot, d)  factor m_d
//   issue_v(slot, y)                 eliminated RHS of row y (written by store_v)
template <class T, class IssueA, class ReadA, class IssueM, class IssueV, class StoreV, class StoreU>
SCB_D void tri_column(int n, bool top, int lane, T* ring, T* xch /* smem [2][kTriCols] */, const IssueA& issue_a, const ReadA& read_a, const IssueM& issue_m,
                      const IssueV& issue_v, const StoreV& store_v, const StoreU& store_u, bool active) {
    constexpr int U = kTriU, D = kTriRingBytesPerWarp / (2 * U * 32 * (int)sizeof(T));
    static_assert(D >= 2, "ring too small");
    const int h = (n + 1) / 2;
    const int len = top ? h : n - h;
    auto row = [&](int d) { return top ? d : n - 1 - d; };
    auto slot = [&](int stage, int arr, int i) { return ring + ((size_t)((stage % D) * 2 + arr) * U + i) * 32 + lane; };
    // ---- forward elimination: v_d = a_row(d) + m_(d-1) v_(d-1) ----
    {
        const int nb = (len + U - 1) / U;
        auto issue = [&](int b) {
            if (active && b < nb) {
                SCB_UNROLL
                for (int i = 0; i < U; ++i) {
                    const int d = b * U + i;
                    if (d < len) {
                        issue_a(slot(b, 0, i), row(d));
                        if (d > 0) issue_m(slot(b, 1, i), d - 1);
                    }
                }
            }
            cp_async_commit();
        };
        for (int s = 0; s < D - 1; ++s) issue(s);
        T v = T(0);
        for (int b = 0; b < nb; ++b) {
            issue(b + D - 1);
            cp_async_wait<D - 1>();
            if (active) {
                SCB_UNROLL
                for (int i = 0; i < U; ++i) {
                    const int d = b * U + i;
                    if (d < len) {
                        const T a = read_a(slot(b, 0, i));
                        const T m = d > 0 ? *slot(b, 1, i) : T(0);
                        v = m * v + a;
                        store_v(row(d), v);
                    }
                }
            }
        }
        cp_async_wait<0>();
        // ---- the two halves meet ----
        if (active) xch[(top ? 0 : 1) * kTriCols + lane] = v;
    }
    __syncthreads();
    // pivots of rows h-1 (from the top) and h (from the bottom): two plain loads
    T u = T(0);
    if (active) {
        T mP, mQ = T(1);
        issue_m(slot(0, 1, 0), h - 1);
        if (n - h > 0) issue_m(slot(0, 1, 1), n - h - 1);
        cp_async_commit();
        cp_async_wait<0>();
        mP = *slot(0, 1, 0);
        if (n - h > 0) mQ = *slot(0, 1, 1);
        const T t = xch[lane];
        if (n - h == 0) {  // a single row
            u = t * mP;
        } else {
            const T b = xch[kTriCols + lane];
            const T P = T(1) / mP, Q = T(1) / mQ;
            const T det = P * Q - T(1);
            u = top ? (Q * t + b) / det : (t + P * b) / det;
        }
        if (len > 0) store_u(row(len - 1), u);
    }
    // ---- back substitution outwards: u_row(d) = m_d (v_d + u_row(d+1)),  e = 0.. <-> d = len-2-e ----
    {
        const int cnt = len - 1;
        const int nb = cnt > 0 ? (cnt + U - 1) / U : 0;
        auto issue = [&](int b) {
            if (active && b < nb) {
                SCB_UNROLL
                for (int i = 0; i < U; ++i) {
                    const int e = b * U + i;
                    if (e < cnt) {
                        const int d = len - 2 - e;
                        issue_v(slot(b, 0, i), row(d));
                        issue_m(slot(b, 1, i), d);
                    }
                }
            }
            cp_async_commit();
        };
        for (int s = 0; s < D - 1; ++s) issue(s);
        for (int b = 0; b < nb; ++b) {
            issue(b + D - 1);
            cp_async_wait<D - 1>();
            if (active) {
                SCB_UNROLL
                for (int i = 0; i < U; ++i) {
                    const int e = b * U + i;
                    if (e < cnt) {
                        const T vv = *slot(b, 0, i), m = *slot(b, 1, i);
                        u = m * u + m * vv;
                        store_u(row(len - 2 - e), u);
                    }
                }
            }
        }
        cp_async_wait<0>();
    }
}

// grid = (ceil((x1 - x0) / 32), 3), block = 64: warp 0 eliminates from the top, warp 1 from the bottom.
__global__ void __launch_bounds__(2 * kTriCols) tri_solve_kernel(TriSolveParams p) {
    SCB_DYN_SMEM(unsigned char, ring_raw);
    __shared__ double xch_raw[2 * kTriCols];
    const int lane = threadIdx.x & 31;
    const bool top = threadIdx.x < kTriCols;
    const int c = blockIdx.y;
    const int kb = p.x0 + (int)blockIdx.x * kTriCols;
    const int k = kb + lane;
    const bool active = k < p.x1;
    const int n = p.ny;
    const float* A = p.A + (size_t)c * n * p.nx + k;
    float* Ct = p.Ct + (size_t)c * n * p.nx + k;
    unsigned char* ring = ring_raw + (top ? 0 : kTriRingBytesPerWarp);
    if (kb < kTriLowK) {  // float64 columns (whole warp: kTriLowK == kTriCols)
        double* Y = p.Y64 + (size_t)c * n * kTriLowK + k;
        const double* m = p.tab.m64 + k;
        const double* R = (p.R && k < p.lowkx) ? p.R + ((size_t)c * p.lowkx + k) * n : nullptr;
        tri_column<double>(
            n, top, lane, reinterpret_cast<double*>(ring), xch_raw,
            [&](double* s, int y) {
                if (R)
                    cp_async<8>(s, R + y);
                else
                    cp_async<4>(s, A + (size_t)y * p.nx);
            },
            [&](const double* s) { return R ? -2.0 * *s : (double)*reinterpret_cast<const float*>(s); },
            [&](double* s, int d) { cp_async<8>(s, m + (size_t)d * kTriLowK); },
            [&](double* s, int y) { cp_async<8>(s, Y + (size_t)y * kTriLowK); },
            [&](int y, double v) { Y[(size_t)y * kTriLowK] = v; },
            [&](int y, double u) { Y[(size_t)y * kTriLowK] = u; }, active);
    } else {
        const float* m = p.tab.m32 + k;
        const int pm = p.tab.pm;
        tri_column<float>(
            n, top, lane, reinterpret_cast<float*>(ring), reinterpret_cast<float*>(xch_raw),
            [&](float* s, int y) { cp_async<4>(s, A + (size_t)y * p.nx); },
            [&](const float* s) { return *s; },
            [&](float* s, int d) { cp_async<4>(s, m + (size_t)d * pm); },
            [&](float* s, int y) { cp_async<4>(s, Ct + (size_t)y * p.nx); },
            [&](int y, float v) { Ct[(size_t)y * p.nx] = v; },
            [&](int y, float u) { Ct[(size_t)y * p.nx] = u; }, active);
    }
}

// ---------------------------------------------------------------------------------------------
// Low-frequency block: float64 solution of columns k < kTriLowK + OpenCV's float32 denominators for l < kTriLowL.
//   grid = (min(kTriLowK, nx), 3), block = 128
// ---------------------------------------------------------------------------------------------
struct TriLowParams {
    int nx, ny;
    const float* A;        // [3][ny][nx]
    const double* R;       // [3][lowkx][ny] or null
    int lowkx;
    const double* Y64;     // [3][ny][kTriLowK] float64 tridiagonal solution of the low columns
    const double* sinfull; // [2 (ny+1)]  sin(pi i / (ny+1))
    const float* fx;       // OpenCV filter_X (nx)
    const float* fy;       // OpenCV filter_Y (ny)
    float* Ct;             // [3][ny][nx]
};

static constexpr int kTriLowThreads = 128;

__global__ void __launch_bounds__(kTriLowThreads) tri_lowcorr_kernel(TriLowParams p) {
    __shared__ double red[(kTriLowThreads / 32) * kTriLowL];
    __shared__ double wl[kTriLowL];
    const int tid = threadIdx.x, k = blockIdx.x, c = blockIdx.y, n = p.ny;
    const int N2 = 2 * (n + 1);
    const int L = n < kTriLowL ? n : kTriLowL;
    const double* R = (p.R && k < p.lowkx) ? p.R + ((size_t)c * p.lowkx + k) * n : nullptr;
    const float* A = p.A + (size_t)c * n * p.nx + k;
    double acc[kTriLowL];
    SCB_UNROLL
    for (int l = 0; l < kTriLowL; ++l) acc[l] = 0.0;
    for (int y = tid; y < n; y += kTriLowThreads) {
        const double a = R ? -2.0 * R[y] : (double)__ldg(A + (size_t)y * p.nx);
        int idx = 0;  // ((y+1)(l+1)) mod 2N, incrementally
        SCB_UNROLL
        for (int l = 0; l < kTriLowL; ++l) {
            idx += y + 1;
            if (idx >= N2) idx -= N2;
            if (l < L) acc[l] += a * __ldg(p.sinfull + idx);
        }
    }
    block_reduce_store<kTriLowL>(acc, red, tid);
    if (tid < kTriLowL) {
        double w = 0.0;
        if (tid < L) {
            double t = 0.0;
            for (int wi = 0; wi < kTriLowThreads / 32; ++wi) t += red[wi * kTriLowL + tid];
            const float fxk = __ldg(p.fx + k), fyl = __ldg(p.fy + tid);
            const double den32 = (double)__fsub_rn(__fadd_rn(fxk, fyl), 4.0f);             // OpenCV: (filter_X + filter_Y) - 4 in float32
            const double den = (double)fxk + 2.0 * cospi((double)(tid + 1) / (double)(n + 1)) - 4.0;  // what the tridiagonal solve divides by
            w = -(2.0 / (double)(n + 1)) * t * (1.0 / den32 - 1.0 / den);
        }
        wl[tid] = w;
    }
    __syncthreads();
    const double* Y = p.Y64 + (size_t)c * n * kTriLowK + k;
    float* Ct = p.Ct + (size_t)c * n * p.nx + k;
    for (int y = tid; y < n; y += kTriLowThreads) {
        double s = Y[(size_t)y * kTriLowK];
        int idx = 0;
        SCB_UNROLL
        for (int l = 0; l < kTriLowL; ++l) {
            idx += y + 1;
            if (idx >= N2) idx -= N2;
            if (l < L) s += wl[l] * __ldg(p.sinfull + idx);
        }
        Ct[(size_t)y * p.nx] = (float)s;
    }
}

}  // namespace scb
