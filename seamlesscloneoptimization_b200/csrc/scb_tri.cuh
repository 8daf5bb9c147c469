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
// cv2.seamlessClone; numpy model in tests/test_tri_model.py: an exact tridiagonal solve for k >= 32 leaves the byte-exact
// fraction at the float64 floor; for all k it drops to 41 % at 4K).  So:
//   * columns k < kTriLowK are solved in float64 (their systems have condition ~ (N / pi k)^2), and
//   * tri_lowproj_kernel / tri_lowapply_kernel add, for k < kTriLowK and l < kTriLowL, the difference between OpenCV's float32
//     denominator and the exact one:  Ct_k += -(2/Ny) sum_l sin_l (<sin_l, a_exact_k> / den32[k][l] - <sin_l, a_fft_k> / den[k][l])
//     in float64, a_exact = the exact float64 row sums where those exist (k < kLowK: lowfreq_rows_kernel).
// That block subsumes the 8 x 8 exact low-frequency corner of the FFT engine.
//
// LU factors in closed form.  With beta = 2 cosh(theta), rho = exp(-theta), the pivots of M are
//     p_j = sinh((j+1) theta) / sinh(j theta),   m_j = 1 / p_j = rho (1 - rho^2j) / (1 - rho^(2j+2)),   j = 1..n,
// evaluated in float64 at plan time (tri_table_kernel; cached per (w, h) in the context) and rounded once.
// A column is cut into up to 16 segments solved concurrently (SPIKE partitioning, see tri_segments): a plain Thomas
// sweep is a dependent chain of ny steps per column with only 3 nx columns to spread over 148 SMs.
#pragma once

#include <cstring>

#include "scb_kernels.cuh"
#include "scb_platform.h"

namespace scb {

static constexpr int kTriLowK = 32;  // spectral columns solved in float64 and corrected to OpenCV's float32 denominators
static constexpr int kTriLowL = 32;  // ... for this many lowest frequencies along y
#ifndef SCB_TRI_COLS
#define SCB_TRI_COLS 16
#endif
static constexpr int kTriCols = SCB_TRI_COLS;  // columns per CTA: 16 -> 64-byte rows per segment, twice the CTAs (171 CTAs of 32 columns left 23 of the 148 SMs with two CTAs and the rest with one)
static_assert(32 % kTriCols == 0 && kTriLowK % kTriCols == 0, "column tiles must pack into warps and split the float64 block evenly");
#ifndef SCB_TRI_SEGS
#define SCB_TRI_SEGS 16
#endif
static constexpr int kTriSegs = SCB_TRI_SEGS;  // segments per column
#ifndef SCB_TRI_UNROLL
#define SCB_TRI_UNROLL 8
#endif
static constexpr int kTriUnroll = SCB_TRI_UNROLL;  // rows whose loads are in flight together

// Segment length for ny unknowns: at most kTriSegs segments, none shorter than 4 rows unless the column is.
SCB_HD int tri_seg_len(int ny) {
    int s = ny / 4;
    if (s < 1) s = 1;
    if (s > kTriSegs) s = kTriSegs;
    return (ny + s - 1) / s;
}

struct TriTabDev {
    const float* m32;   // [rows][pm]  m_d = sinh((d+1) theta) / sinh((d+2) theta): reciprocal pivot at distance d from a segment end
    const float* p32;   // [rows][pm]  P_d = sinh(theta) / sinh((d+1) theta)     : prod_{i<d} m_i
    int pm, rows;       // rows = tri_seg_len(ny) + 1
    const double* m64;  // [rows][kTriLowK]
    const double* p64;  // [rows][kTriLowK]
    const double* theta;  // [nx]
};

struct TriTableParams {
    const double* theta;  // [nx]  acosh((4 - fx[k]) / 2)
    int nx, rows, pm;
    float *m32, *p32;
    double *m64, *p64;
};

// rho = exp(-theta):  m_d = rho (1 - rho^(2d+2)) / (1 - rho^(2d+4)),   P_d = rho^d (1 - rho^2) / (1 - rho^(2d+2))
__global__ void __launch_bounds__(256) tri_table_kernel(TriTableParams p) {
    const long long total = (long long)p.rows * p.pm;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int d = (int)(i / p.pm), k = (int)(i - (long long)d * p.pm);
        double m = 0.0, P = 0.0;
        if (k < p.nx) {
            const double th = p.theta[k];
            if (th < 1e-10) {
                m = (double)(d + 1) / (double)(d + 2);
                P = 1.0 / (double)(d + 1);
            } else {
                m = exp(-th) * (expm1(-2.0 * (d + 1) * th) / expm1(-2.0 * (d + 2) * th));
                P = exp(-(double)d * th) * (expm1(-2.0 * th) / expm1(-2.0 * (d + 1) * th));
            }
        }
        p.m32[i] = (float)m;
        p.p32[i] = (float)P;
        if (k < kTriLowK) {
            p.m64[(size_t)d * kTriLowK + k] = m;
            p.p64[(size_t)d * kTriLowK + k] = P;
        }
    }
}

struct TriSolveParams {
    TriTabDev tab;
    int nx, ny;
    const float* A;    // [3][ny][nx]  row-transformed RHS (OpenCV scale: -2 sum g sin)
    float* Ct;         // [3][ny][nx]
    double* Y64;       // [3][ny][kTriLowK] float64 work / result columns k < kTriLowK
    int x0, x1;        // columns of this launch (x0 a multiple of kTriCols)
    int seg_len;       // tri_seg_len(ny)
    // Row-sharded solves (one rank per GPU owns the segments [seg0, seg1) of every column):
    //   phase 0  everything in one launch (seg0 = 0, seg1 = all)
    //   phase 1  pass 1 of the own segments; their local-solution ends go to `ends32` / `ends64`
    //   phase 2  (after the ranks have exchanged the ends) reduced systems, then pass 2 of the own segments
    int phase, seg0, seg1;
    float* ends32;     // [3][kTriSegs][2][pm]        w_last, w_first of every segment, columns >= kTriLowK
    double* ends64;    // [3][kTriSegs][2][kTriLowK]  the same for the float64 columns
};

// Partitioned (SPIKE) Thomas solve of M u = a, M = tridiag(-1, beta, -1), one thread per (column, segment):
//   pass 1  every segment eliminates its own rows downwards (v, stored) and upwards (kept in a register): the two ends
//           w_first, w_last of its LOCAL solution (neighbours taken as zero) fall out without a back substitution;
//   reduce  per column, the 2 (S-1) values next to the segment interfaces solve a small banded system whose coefficients
//           are the spike ends  e = sinh(theta)/sinh((L+1) theta),  f = sinh(L theta)/sinh((L+1) theta)  (table entries);
//   pass 2  back substitution with the TRUE neighbours: u_d = m_d (v_d + alpha P_d + u_(d+1)),  u_L = beta
//           (alpha = u just above the segment, beta = u just below).
// The pivots restart in every segment, so only seg_len + 1 table rows exist and they stay in L1/L2.
template <class T>
struct TriIo;  // global-memory views of one column: a(y), v/u(y) and the tables

// The reduced system of one column (run by the thread of its first segment, between pass 1 and pass 2): from the ends of the
// segments' local solutions (sm rows 0 = w_last, 1 = w_first) to the true values just above / below every segment
// (alpha[s] = sm row 2, beta[s] = sm row 3; rows 4, 5 are scratch).
template <class T, class LoadM, class LoadP>
SCB_D void tri_reduced_system(int n, int L, int nseg, int lane, T* sm /* [6][kTriSegs][kTriCols] */, const LoadM& load_m, const LoadP& load_p) {
    T* al = sm + (2 * kTriSegs) * kTriCols + lane;  // alpha[s] at al[s * kTriCols]
    T* be = sm + (3 * kTriSegs) * kTriCols + lane;  // beta[s]
    T* Js = sm + (4 * kTriSegs) * kTriCols + lane;
    T* Ks = sm + (5 * kTriSegs) * kTriCols + lane;
    const int lastlen = n - (nseg - 1) * L;
    auto ecoef = [&](int s) { return load_p((s == nseg - 1) ? lastlen : L); };      // sinh(theta) / sinh((len+1) theta)
    auto fcoef = [&](int s) { return load_m(((s == nseg - 1) ? lastlen : L) - 1); };  // sinh(len theta) / sinh((len+1) theta)
    const T* WL = sm + lane;
    const T* WF = sm + kTriSegs * kTriCols + lane;
    // p_s = G + H q_s  (p_s: last row of segment s, q_s: first row of segment s+1);  q_s = J_s + K_s q_(s+1)
    T G = WL[0], H = fcoef(0);
    T* Gs = al;  // alpha/beta slots double as G/H storage until the back substitution
    T* Hs = be;
    for (int s = 0; s + 1 < nseg; ++s) {
        const T e1 = ecoef(s + 1), f1 = fcoef(s + 1);
        const T den = T(1) - f1 * H;
        const T J = (WF[(s + 1) * kTriCols] + f1 * G) / den, K = e1 / den;
        Js[s * kTriCols] = J;
        Ks[s * kTriCols] = K;
        Gs[s * kTriCols] = G;
        Hs[s * kTriCols] = H;
        const T Gn = WL[(s + 1) * kTriCols] + e1 * (G + H * J);
        H = f1 + e1 * H * K;
        G = Gn;
    }
    // back substitution: q_(nseg-1) = 0 (Dirichlet boundary below the last segment);  alpha_s = p_(s-1),  beta_s = q_s
    T qn = T(0);
    be[(nseg - 1) * kTriCols] = T(0);
    for (int s = nseg - 2; s >= 0; --s) {
        const T q = Js[s * kTriCols] + Ks[s * kTriCols] * qn;
        const T pp = Gs[s * kTriCols] + Hs[s * kTriCols] * q;
        be[s * kTriCols] = q;         // H_s consumed
        al[(s + 1) * kTriCols] = pp;  // G_(s+1) consumed
        qn = q;
    }
    al[0] = T(0);
}

// INPLACE: a and v share their storage (tri_solve_smem_kernel's shared-memory tile): the upward sweep, which needs the original a,
// runs to its end before the downward sweep overwrites a with v.  Each chain is the same sequence of operations either way, so
// the results are bit-identical to the fused loop.  store_u receives the final solution of pass 2 (store_v: pass 1's v).
template <class T, bool INPLACE, class LoadA, class LoadM, class LoadP, class LoadV, class StoreV, class StoreU>
SCB_D void tri_segments(int n, int L, int seg, int nseg, int lane, T* sm /* [6][kTriSegs][kTriCols] */, const LoadA& load_a, const LoadM& load_m, const LoadP& load_p,
                        const LoadV& load_v, const StoreV& store_v, const StoreU& store_u, bool active, int phase, int seg0, int seg1,
                        T* ends /* global [kTriSegs][2][estride] + column, or null */, size_t estride) {
    const int r0 = seg * L;
    const bool mine = seg >= seg0 && seg < seg1 && seg < nseg;
    const int len = mine ? ((n - r0 < L) ? n - r0 : L) : 0;
    T* wl = sm + (0 * kTriSegs + seg) * kTriCols + lane;
    T* wf = sm + (1 * kTriSegs + seg) * kTriCols + lane;
    // ---- pass 1 ----
    if (INPLACE && active && len > 0 && phase != 2) {
        T b = load_a(r0 + len - 1);
        int d = 1;
        for (; d + kTriUnroll <= len; d += kTriUnroll) {
            T ab[kTriUnroll], mm[kTriUnroll];
            SCB_UNROLL
            for (int i = 0; i < kTriUnroll; ++i) {
                ab[i] = load_a(r0 + len - 1 - d - i);
                mm[i] = load_m(d + i - 1);
            }
            SCB_UNROLL
            for (int i = 0; i < kTriUnroll; ++i) b = mm[i] * b + ab[i];
        }
        for (; d < len; ++d) b = load_m(d - 1) * b + load_a(r0 + len - 1 - d);
        T v = load_a(r0);
        store_v(r0, v);
        d = 1;
        for (; d + kTriUnroll <= len; d += kTriUnroll) {
            T av[kTriUnroll], mm[kTriUnroll];
            SCB_UNROLL
            for (int i = 0; i < kTriUnroll; ++i) {
                av[i] = load_a(r0 + d + i);
                mm[i] = load_m(d + i - 1);
            }
            SCB_UNROLL
            for (int i = 0; i < kTriUnroll; ++i) {
                v = mm[i] * v + av[i];
                store_v(r0 + d + i, v);
            }
        }
        for (; d < len; ++d) {
            v = load_m(d - 1) * v + load_a(r0 + d);
            store_v(r0 + d, v);
        }
        const T ml = load_m(len - 1);
        if (phase == 1) {
            ends[((size_t)seg * 2 + 0) * estride] = ml * v;
            ends[((size_t)seg * 2 + 1) * estride] = ml * b;
        } else {
            *wl = ml * v;
            *wf = ml * b;
        }
    }
    if (!INPLACE && active && len > 0 && phase != 2) {
        T v = load_a(r0), b = load_a(r0 + len - 1);
        store_v(r0, v);
        int d = 1;
        for (; d + kTriUnroll <= len; d += kTriUnroll) {
            T av[kTriUnroll], ab[kTriUnroll], mm[kTriUnroll];
            SCB_UNROLL
            for (int i = 0; i < kTriUnroll; ++i) {
                av[i] = load_a(r0 + d + i);
                ab[i] = load_a(r0 + len - 1 - d - i);
                mm[i] = load_m(d + i - 1);
            }
            SCB_UNROLL
            for (int i = 0; i < kTriUnroll; ++i) {
                v = mm[i] * v + av[i];
                b = mm[i] * b + ab[i];
                store_v(r0 + d + i, v);
            }
        }
        for (; d < len; ++d) {
            const T m = load_m(d - 1);
            v = m * v + load_a(r0 + d);
            b = m * b + load_a(r0 + len - 1 - d);
            store_v(r0 + d, v);
        }
        const T ml = load_m(len - 1);
        if (phase == 1) {
            ends[((size_t)seg * 2 + 0) * estride] = ml * v;
            ends[((size_t)seg * 2 + 1) * estride] = ml * b;
        } else {
            *wl = ml * v;
            *wf = ml * b;
        }
    }
    if (phase == 1) return;
    if (phase == 2) {  // every segment's ends, gathered from all ranks
        if (active && seg < nseg) {
            *wl = ends[((size_t)seg * 2 + 0) * estride];
            *wf = ends[((size_t)seg * 2 + 1) * estride];
        }
    }
    __syncthreads();
    // ---- reduced system: thread (lane, seg 0) of every column ----
    if (active && seg == 0) tri_reduced_system<T>(n, L, nseg, lane, sm, load_m, load_p);
    __syncthreads();
    // ---- pass 2 ----
    const T* al = sm + (2 * kTriSegs) * kTriCols + lane;  // alpha[s] at al[s * kTriCols]
    const T* be = sm + (3 * kTriSegs) * kTriCols + lane;  // beta[s]
    if (active && len > 0) {
        const T alpha = al[seg * kTriCols];
        T u = be[seg * kTriCols];
        int d = len - 1;
        for (; d >= kTriUnroll - 1; d -= kTriUnroll) {
            T vv[kTriUnroll], mm[kTriUnroll], pp[kTriUnroll];
            SCB_UNROLL
            for (int i = 0; i < kTriUnroll; ++i) {
                vv[i] = load_v(r0 + d - i);
                mm[i] = load_m(d - i);
                pp[i] = load_p(d - i);
            }
            SCB_UNROLL
            for (int i = 0; i < kTriUnroll; ++i) {
                u = mm[i] * (u + (pp[i] * alpha + vv[i]));
                store_u(r0 + d - i, u);
            }
        }
        for (; d >= 0; --d) {
            u = load_m(d) * (u + (load_p(d) * alpha + load_v(r0 + d)));
            store_u(r0 + d, u);
        }
    }
}

// grid = (ceil((x1 - x0) / kTriCols), 3), block = kTriCols x kTriSegs
__global__ void __launch_bounds__(kTriCols * kTriSegs) tri_solve_kernel(TriSolveParams p) {
    __shared__ double sm_raw[6 * kTriSegs * kTriCols];
    const int lane = threadIdx.x % kTriCols, seg = threadIdx.x / kTriCols;  // a warp holds 32 / kTriCols segments of the same columns
    const int c = blockIdx.y;
    const int kb = p.x0 + (int)blockIdx.x * kTriCols;
    const int k = kb + lane;
    const bool active = k < p.x1;
    const int n = p.ny, L = p.seg_len;
    const int nseg = (n + L - 1) / L;
    const float* A = p.A + (size_t)c * n * p.nx + k;
    float* Ct = p.Ct + (size_t)c * n * p.nx + k;
    if (kb < kTriLowK) {  // float64 columns (whole CTA: kTriLowK is a multiple of kTriCols)
        double* Y = p.Y64 + (size_t)c * n * kTriLowK + k;
        const double* m = p.tab.m64 + k;
        const double* P = p.tab.p64 + k;
        const auto store = [&](int y, double v) { Y[(size_t)y * kTriLowK] = v; };
        tri_segments<double, false>(
            n, L, seg, nseg, lane, sm_raw,
            [&](int y) { return (double)__ldg(A + (size_t)y * p.nx); },
            [&](int d) { return __ldg(m + (size_t)d * kTriLowK); },
            [&](int d) { return __ldg(P + (size_t)d * kTriLowK); },
            [&](int y) { return Y[(size_t)y * kTriLowK]; }, store, store, active, p.phase, p.seg0, p.seg1,
            p.ends64 ? p.ends64 + (size_t)c * kTriSegs * 2 * kTriLowK + k : nullptr, (size_t)kTriLowK);
    } else {
        const float* m = p.tab.m32 + k;
        const float* P = p.tab.p32 + k;
        const int pm = p.tab.pm;
        const auto store = [&](int y, float v) { Ct[(size_t)y * p.nx] = v; };
        tri_segments<float, false>(
            n, L, seg, nseg, lane, reinterpret_cast<float*>(sm_raw),
            [&](int y) { return __ldg(A + (size_t)y * p.nx); },
            [&](int d) { return __ldg(m + (size_t)d * pm); },
            [&](int d) { return __ldg(P + (size_t)d * pm); },
            [&](int y) { return Ct[(size_t)y * p.nx]; }, store, store, active, p.phase, p.seg0, p.seg1,
            p.ends32 ? p.ends32 + (size_t)c * kTriSegs * 2 * pm + k : nullptr, (size_t)pm);
    }
}

// ---------------------------------------------------------------------------------------------
// tri_solve_smem_kernel (SCB_TRI_SMEM=1): the same partitioned solve with the CTA's column tile held in SHARED MEMORY.
// tri_solve_kernel walks global memory row by row -- 3 dependent-latency loads per row and pass, v written out and read back: a chain
// of ~2 ny / kTriSegs L2 round trips per thread with ~18 warps per SM (latency bound, profiles/r2_final_ncu_full_cfg2.txt).  Here
// every thread first copies the rows of its own segment into the tile (independent loads, all in flight together), both sweeps of
// pass 1 and the reduced system then run out of shared memory (29-cycle loads), and pass 2 streams the solution straight to global
// memory: A is read once and Ct written once (8 B per unknown instead of 24), and the dependent chain never leaves the SM.
//   CTA tile : float columns k >= kTriLowK: kTriCols columns;  float64 columns: 8 (so that both kinds need ~the same shared memory
//              and two CTAs fit an SM);  rows of a segment are contiguous, consecutive segments sit an odd multiple of 16 words apart
//              (the two segments of a warp hit disjoint banks)
//   tables   : the tile's columns of m / P, (seg_len + 1) rows, in shared memory too when they fit (tab_smem)
// Same arithmetic, operation for operation, as tri_solve_kernel (tri_segments<T, INPLACE = true>): bit-identical results
// (tests/test_tri_smem.py).  One launch = the whole solve (phase 0, x0 = 0); the row-sharded phases stay on tri_solve_kernel.
// grid = (kTriLowK / 8 + ceil((nx - kTriLowK) / kTriCols), 3), block = kTriCols x kTriSegs, dynamic smem = tri_smem_bytes()
// ---------------------------------------------------------------------------------------------
static constexpr int kTriCols64 = 8;  // columns per CTA of the float64 block
static constexpr size_t kTriSmemLimit = 232448;  // opt-in dynamic shared memory of one CTA on sm_100

SCB_HD int tri_smem_seg_stride(int L, int ncol, int elem_bytes) {  // elements between the tiles of consecutive segments
    const int e = L * ncol;
    return (elem_bytes == 4 && ncol == 16 && (e % 32) == 0) ? e + 16 : e;
}
// dynamic shared memory of one CTA: reduced-system scratch + (tables) + tile, for the larger of the two CTA kinds; 0 = does not fit
SCB_HD size_t tri_smem_bytes(int ny, bool tables, size_t limit = kTriSmemLimit) {
    const int L = tri_seg_len(ny), nseg = (ny + L - 1) / L;
    size_t need = 0;
    for (int f64 = 0; f64 < 2; ++f64) {
        const int ncol = f64 ? kTriCols64 : kTriCols, eb = f64 ? 8 : 4;
        size_t elems = (size_t)6 * kTriSegs * kTriCols + (size_t)nseg * tri_smem_seg_stride(L, ncol, eb);
        if (tables) elems += (size_t)2 * (L + 1) * ncol;
        if (elems * eb > need) need = elems * eb;
    }
    return need <= limit ? need : 0;
}

template <class T, int NCOL>
SCB_D void tri_smem_body(const TriSolveParams& p, T* sm, int kb, int c, int tab_smem) {
    const int lane = threadIdx.x % kTriCols, seg = threadIdx.x / kTriCols;
    const int k = kb + lane;
    const bool active = lane < NCOL && k < p.x1;
    const int n = p.ny, L = p.seg_len;
    const int nseg = (n + L - 1) / L;
    const int rows = L + 1;
    T* red = sm;
    T* tabm = red + 6 * kTriSegs * kTriCols;
    T* tabp = tabm + (tab_smem ? rows * NCOL : 0);
    T* tile = tabp + (tab_smem ? rows * NCOL : 0);
    const int ss = tri_smem_seg_stride(L, NCOL, (int)sizeof(T));
    constexpr bool kF64 = sizeof(T) == 8;
    const T* gm;
    const T* gp;
    size_t gstride;
    if constexpr (kF64) {
        gm = reinterpret_cast<const T*>(p.tab.m64);
        gp = reinterpret_cast<const T*>(p.tab.p64);
        gstride = kTriLowK;
    } else {
        gm = reinterpret_cast<const T*>(p.tab.m32);
        gp = reinterpret_cast<const T*>(p.tab.p32);
        gstride = (size_t)p.tab.pm;
    }
    if (tab_smem) {
        for (int i = threadIdx.x; i < rows * NCOL; i += kTriCols * kTriSegs) {
            const int d = i / NCOL, col = i - d * NCOL;
            const bool in = kb + col < p.x1;
            tabm[i] = in ? gm[(size_t)d * gstride + kb + col] : T(0);
            tabp[i] = in ? gp[(size_t)d * gstride + kb + col] : T(0);
        }
    }
    // own segment's rows of A -> tile (independent loads; a warp reads 32 / kTriCols rows x kTriCols consecutive columns per step)
    const int r0 = seg * L;
    const int len = (seg < nseg) ? ((n - r0 < L) ? n - r0 : L) : 0;
    T* mine = tile + (size_t)seg * ss + lane;  // row d of the segment at mine[d * NCOL]
    if (active && len > 0) {
        const float* A = p.A + ((size_t)c * n + r0) * p.nx + k;
        int d = 0;
        for (; d + kTriUnroll <= len; d += kTriUnroll) {
            float a[kTriUnroll];
            SCB_UNROLL
            for (int i = 0; i < kTriUnroll; ++i) a[i] = __ldg(A + (size_t)(d + i) * p.nx);
            SCB_UNROLL
            for (int i = 0; i < kTriUnroll; ++i) mine[(d + i) * NCOL] = (T)a[i];
        }
        for (; d < len; ++d) mine[d * NCOL] = (T)__ldg(A + (size_t)d * p.nx);
    }
    if (tab_smem) __syncthreads();  // the tables are read by every segment
    const T* mrow = tab_smem ? tabm + lane : gm + k;
    const T* prow = tab_smem ? tabp + lane : gp + k;
    const size_t tstride = tab_smem ? (size_t)NCOL : gstride;
    T* base = mine - (size_t)r0 * NCOL;  // row y of the column at base[y * NCOL], valid for the rows of the own segment
    const auto tile_load = [&](int y) { return base[(size_t)y * NCOL]; };
    const auto tile_store = [&](int y, T v) { base[(size_t)y * NCOL] = v; };
    if constexpr (kF64) {
        double* Y = p.Y64 + (size_t)c * n * kTriLowK + k;
        tri_segments<T, true>(
            n, L, seg, nseg, lane, red, tile_load, [&](int d) { return mrow[(size_t)d * tstride]; }, [&](int d) { return prow[(size_t)d * tstride]; }, tile_load, tile_store,
            [&](int y, T u) { Y[(size_t)y * kTriLowK] = u; }, active, 0, 0, kTriSegs, (T*)nullptr, (size_t)0);
    } else {
        float* Ct = p.Ct + (size_t)c * n * p.nx + k;
        tri_segments<T, true>(
            n, L, seg, nseg, lane, red, tile_load, [&](int d) { return mrow[(size_t)d * tstride]; }, [&](int d) { return prow[(size_t)d * tstride]; }, tile_load, tile_store,
            [&](int y, T u) { Ct[(size_t)y * p.nx] = u; }, active, 0, 0, kTriSegs, (T*)nullptr, (size_t)0);
    }
}

// tri_solve_smem2_kernel (SCB_TRI_SMEM=2): the same kernel written out by hand -- 32-bit shared-memory offsets instead of 64-bit
// pointer arithmetic through the load / store functors of tri_segments, the table pointers in ONE address space per instantiation
// (tri_solve_smem_kernel selects shared or global tables at run time, which turns every table access into a generic load with 64-bit
// address arithmetic: ~56 instructions per row, profiles/r2b_*), the ends and the reduced system unchanged.  Same operations in the
// same order: bit-identical to both other kernels (tests/emu/test_tri_smem.cpp).
template <class T, int NCOL, bool TAB>
SCB_D void tri_smem2_body(const TriSolveParams& p, T* sm, int kb, int c) {
    const int lane = threadIdx.x % kTriCols, seg = threadIdx.x / kTriCols;
    const int k = kb + lane;
    const bool active = lane < NCOL && k < p.x1;
    const int n = p.ny, L = p.seg_len;
    const int nseg = (n + L - 1) / L;
    const int rows = L + 1;
    T* red = sm;
    T* tabm = red + 6 * kTriSegs * kTriCols;
    T* tabp = tabm + (TAB ? rows * NCOL : 0);
    T* tile = tabp + (TAB ? rows * NCOL : 0);
    const int ss = tri_smem_seg_stride(L, NCOL, (int)sizeof(T));
    constexpr bool kF64 = sizeof(T) == 8;
    const T* gm = kF64 ? reinterpret_cast<const T*>(p.tab.m64) : reinterpret_cast<const T*>(p.tab.m32);
    const T* gp = kF64 ? reinterpret_cast<const T*>(p.tab.p64) : reinterpret_cast<const T*>(p.tab.p32);
    const int gstride = kF64 ? kTriLowK : p.tab.pm;
    if (TAB) {
        for (int i = threadIdx.x; i < rows * NCOL; i += kTriCols * kTriSegs) {
            const int d = i / NCOL, col = i - d * NCOL;
            const bool in = kb + col < p.x1;
            tabm[i] = in ? __ldg(gm + (size_t)d * gstride + kb + col) : T(0);
            tabp[i] = in ? __ldg(gp + (size_t)d * gstride + kb + col) : T(0);
        }
    }
    const int r0 = seg * L;
    const int len = (seg < nseg) ? ((n - r0 < L) ? n - r0 : L) : 0;
    T* col = tile + seg * ss + lane;  // row d of the own segment at col[d * NCOL]
    const T* gmk = gm + k;
    const T* gpk = gp + k;
    // table entry d of the own column: shared (TAB) or global, decided at compile time
    auto M = [&](int d) { return TAB ? tabm[d * NCOL + lane] : __ldg(gmk + (size_t)d * gstride); };
    auto P = [&](int d) { return TAB ? tabp[d * NCOL + lane] : __ldg(gpk + (size_t)d * gstride); };
    const bool work = active && len > 0;
    if (work) {  // own rows of A -> tile
        const float* A = p.A + ((size_t)c * n + r0) * p.nx + k;
        const size_t nx = (size_t)p.nx;
        int d = 0;
        for (; d + kTriUnroll <= len; d += kTriUnroll) {
            float a[kTriUnroll];
            SCB_UNROLL
            for (int i = 0; i < kTriUnroll; ++i) a[i] = __ldg(A + (d + i) * nx);
            SCB_UNROLL
            for (int i = 0; i < kTriUnroll; ++i) col[(d + i) * NCOL] = (T)a[i];
        }
        for (; d < len; ++d) col[d * NCOL] = (T)__ldg(A + d * nx);
    }
    if (TAB) __syncthreads();  // the tables are read by every segment
    if (work) {
        // ---- pass 1: the upward sweep to its end (it needs the original a), then the downward sweep in place ----
        T b = col[(len - 1) * NCOL];
        int d = 1;
        for (; d + kTriUnroll <= len; d += kTriUnroll) {
            T ab[kTriUnroll], mm[kTriUnroll];
            SCB_UNROLL
            for (int i = 0; i < kTriUnroll; ++i) {
                ab[i] = col[(len - 1 - d - i) * NCOL];
                mm[i] = M(d + i - 1);
            }
            SCB_UNROLL
            for (int i = 0; i < kTriUnroll; ++i) b = mm[i] * b + ab[i];
        }
        for (; d < len; ++d) b = M(d - 1) * b + col[(len - 1 - d) * NCOL];
        T v = col[0];
        d = 1;
        for (; d + kTriUnroll <= len; d += kTriUnroll) {
            T av[kTriUnroll], mm[kTriUnroll];
            SCB_UNROLL
            for (int i = 0; i < kTriUnroll; ++i) {
                av[i] = col[(d + i) * NCOL];
                mm[i] = M(d + i - 1);
            }
            SCB_UNROLL
            for (int i = 0; i < kTriUnroll; ++i) {
                v = mm[i] * v + av[i];
                col[(d + i) * NCOL] = v;
            }
        }
        for (; d < len; ++d) {
            v = M(d - 1) * v + col[d * NCOL];
            col[d * NCOL] = v;
        }
        const T ml = M(len - 1);
        red[(0 * kTriSegs + seg) * kTriCols + lane] = ml * v;
        red[(1 * kTriSegs + seg) * kTriCols + lane] = ml * b;
    }
    __syncthreads();
    if (active && seg == 0) tri_reduced_system<T>(n, L, nseg, lane, red, M, P);
    __syncthreads();
    if (work) {
        // ---- pass 2: back substitution with the true neighbours, straight to global memory ----
        const T alpha = red[(2 * kTriSegs + seg) * kTriCols + lane];
        T u = red[(3 * kTriSegs + seg) * kTriCols + lane];
        const size_t ostride = kF64 ? (size_t)kTriLowK : (size_t)p.nx;
        auto out = [&](int dd, T val) {
            if constexpr (kF64)
                p.Y64[((size_t)c * n + r0 + dd) * kTriLowK + k] = val;
            else
                p.Ct[((size_t)c * n + r0 + dd) * ostride + k] = val;
        };
        int d = len - 1;
        for (; d >= kTriUnroll - 1; d -= kTriUnroll) {
            T vv[kTriUnroll], mm[kTriUnroll], pp[kTriUnroll];
            SCB_UNROLL
            for (int i = 0; i < kTriUnroll; ++i) {
                vv[i] = col[(d - i) * NCOL];
                mm[i] = M(d - i);
                pp[i] = P(d - i);
            }
            SCB_UNROLL
            for (int i = 0; i < kTriUnroll; ++i) {
                u = mm[i] * (u + (pp[i] * alpha + vv[i]));
                out(d - i, u);
            }
        }
        for (; d >= 0; --d) {
            u = M(d) * (u + (P(d) * alpha + col[d * NCOL]));
            out(d, u);
        }
    }
}

__global__ void __launch_bounds__(kTriCols * kTriSegs) tri_solve_smem2_kernel(TriSolveParams p, int tab_smem) {
    SCB_DYN_SMEM(double, sm_dyn);
    constexpr int nb64 = kTriLowK / kTriCols64;
    const int c = blockIdx.y, bx = blockIdx.x;
    if (bx < nb64) {
        if (tab_smem)
            tri_smem2_body<double, kTriCols64, true>(p, sm_dyn, bx * kTriCols64, c);
        else
            tri_smem2_body<double, kTriCols64, false>(p, sm_dyn, bx * kTriCols64, c);
    } else {
        if (tab_smem)
            tri_smem2_body<float, kTriCols, true>(p, reinterpret_cast<float*>(sm_dyn), kTriLowK + (bx - nb64) * kTriCols, c);
        else
            tri_smem2_body<float, kTriCols, false>(p, reinterpret_cast<float*>(sm_dyn), kTriLowK + (bx - nb64) * kTriCols, c);
    }
}

__global__ void __launch_bounds__(kTriCols * kTriSegs) tri_solve_smem_kernel(TriSolveParams p, int tab_smem) {
    SCB_DYN_SMEM(double, sm_dyn);
    constexpr int nb64 = kTriLowK / kTriCols64;
    const int c = blockIdx.y, bx = blockIdx.x;
    if (bx < nb64)
        tri_smem_body<double, kTriCols64>(p, sm_dyn, bx * kTriCols64, c, tab_smem);
    else
        tri_smem_body<float, kTriCols>(p, reinterpret_cast<float*>(sm_dyn), kTriLowK + (bx - nb64) * kTriCols, c, tab_smem);
}

// ---------------------------------------------------------------------------------------------
// Low-frequency block: OpenCV's float32 denominators for k < kTriLowK, l < kTriLowL, in float64.
//   W[c][l][k] = -(2/Ny) ( <sin_l, a_exact_k> / den32[k][l]  -  <sin_l, a_fft_k> / den[k][l] )
//     a_fft   = the float32 row-transform output the tridiagonal solve consumed,
//     a_exact = -2 R (float64 row sums) where they exist (k < lowkx), else a_fft,
//     den32   = fl(fl(fx[k] + fy[l]) - 4)  (OpenCV),   den = fx[k] + 2 cos(pi (l+1) / (ny+1)) - 4  (what M inverts)
//   Ct[c][y][k] = Y64[c][y][k] + sum_l W[c][l][k] sin_l[y]
// tri_lowproj_kernel (sums over y, needs only pass A: runs on the side stream beside the column solve) and
// tri_lowapply_kernel.  Lanes run over k (coalesced rows of A / Y64 / Ct), warps over y.
// ---------------------------------------------------------------------------------------------
struct TriLowParams {
    int nx, ny;
    const float* A;        // [3][ny][nx]
    const double* R;       // [3][lowkx][ny] exact float64 row sums for k < lowkx, or null
    int lowkx;
    const double* Y64;     // [3][ny][kTriLowK] float64 tridiagonal solution of the low columns
    const float* fx;       // OpenCV filter_X (nx)
    const float* fy;       // OpenCV filter_Y (ny)
    double* W;             // [3][kTriLowL][kTriLowK], zeroed before tri_lowproj_kernel
    int w_slots;           // tri_lowapply_kernel: W is the SUM of this many consecutive [3][kTriLowL][kTriLowK] blocks (row-sharded solves: one
                           // partial per rank, exchanged with a single integer all-reduce of disjoint slots); 0 or 1 = just W
    float* Ct;             // [3][ny][nx]
    int y0, y1;            // rows of this launch (row-sharded solves: the own rows; W then holds a partial sum)
};

static constexpr int kTriLowWarps = 8;
static constexpr int kTriLowRows = 32;  // rows of y per CTA

// grid = (ceil(ny / kTriLowRows), 3), block = 32 x kTriLowWarps
__global__ void __launch_bounds__(32 * kTriLowWarps) tri_lowproj_kernel(TriLowParams p) {
    __shared__ double red[2][kTriLowL][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, c = blockIdx.y, n = p.ny;
    for (int i = threadIdx.x; i < 2 * kTriLowL * 32; i += 32 * kTriLowWarps) (&red[0][0][0])[i] = 0.0;
    __syncthreads();
    const int k = lane;
    const bool kin = k < p.nx;
    const int L = n < kTriLowL ? n : kTriLowL;
    const bool has_r = p.R && k < p.lowkx;
    double tf[kTriLowL], te[kTriLowL];  // <sin_l, a_fft>, <sin_l, a_exact - a_fft>
    SCB_UNROLL
    for (int l = 0; l < kTriLowL; ++l) tf[l] = te[l] = 0.0;
    const int yb = p.y0 + blockIdx.x * kTriLowRows;
    for (int y = yb + warp; y < yb + kTriLowRows && y < p.y1; y += kTriLowWarps) {
        const double af = kin ? (double)__ldg(p.A + ((size_t)c * n + y) * p.nx + k) : 0.0;
        const double dx = has_r ? -2.0 * __ldg(p.R + ((size_t)c * p.lowkx + k) * n + y) - af : 0.0;
        // sin(pi (y+1)(l+1) / N), l = 0.., by the three-term recurrence s_(l+1) = 2 cos(phi) s_l - s_(l-1) in float64
        // (32 steps from phi >= pi / 8193: error < 1e-11, against 1e-9 needed)
        const double ph = (double)(y + 1) / (double)(n + 1);
        const double twoc = 2.0 * cospi(ph);
        double s0 = 0.0, s1 = sinpi(ph);
        SCB_UNROLL
        for (int l = 0; l < kTriLowL; ++l) {
            tf[l] += s1 * af;
            te[l] += s1 * dx;
            const double s2 = twoc * s1 - s0;
            s0 = s1;
            s1 = s2;
        }
    }
    SCB_UNROLL
    for (int l = 0; l < kTriLowL; ++l) {
        atomicAdd(&red[0][l][lane], tf[l]);
        atomicAdd(&red[1][l][lane], te[l]);
    }
    __syncthreads();
    // thread (warp, lane) finishes frequencies l = warp, warp + 8, ... of column k = lane
    if (kin) {
        const float fxk = __ldg(p.fx + k);
        for (int l = warp; l < L; l += kTriLowWarps) {
            const double sf = red[0][l][lane], se = red[1][l][lane];
            const double i32 = 1.0 / (double)__fsub_rn(__fadd_rn(fxk, __ldg(p.fy + l)), 4.0f);             // OpenCV: (filter_X + filter_Y) - 4 in float32
            const double iex = 1.0 / ((double)fxk + 2.0 * cospi((double)(l + 1) / (double)(n + 1)) - 4.0);  // what the tridiagonal solve divides by
            const double w = -(2.0 / (double)(n + 1)) * ((sf + se) * i32 - sf * iex);
            atomicAdd(p.W + ((size_t)c * kTriLowL + l) * kTriLowK + k, w);
        }
    }
}

// tri_lowproj2_kernel (SCB_LOWPROJ=2): the same projections without the 64 float64 accumulators per thread (170 registers, one CTA
// per SM) and without the shared-memory atomics that fold eight warps' partial sums (8-way contention on every address -- the bulk of
// tri_lowproj_kernel's 13 us).  The CTA's 32 rows of sines (the same three-term recurrence, one row per lane of warp 0) and of
// a_fft / a_exact - a_fft go to shared memory once; thread (warp w, lane k) then owns the frequencies l = w, w + 8, w + 16, w + 24 of
// column k outright and sums them over the CTA's rows in registers -- no reduction inside the CTA -- and finishes them as before
// (one float64 atomicAdd per (l, k) and CTA into W).  Summation order differs from tri_lowproj_kernel's (which is itself not
// deterministic: atomics), values agree to ~1e-15 relative.
// grid = (ceil(ny / kTriLowRows), 3), block = 32 x kTriLowWarps
__global__ void __launch_bounds__(32 * kTriLowWarps) tri_lowproj2_kernel(TriLowParams p) {
    static_assert(kTriLowRows == 32 && kTriLowL == 32 && kTriLowK == 32, "one row / frequency / column per lane");
    __shared__ double S[kTriLowRows][kTriLowL + 1];  // sin(pi (y+1)(l+1) / N); +1: the recurrence writes a column per step, conflict free
    __shared__ double Af[kTriLowRows][32];           // a_fft
    __shared__ double Dx[kTriLowRows][32];           // a_exact - a_fft
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, c = blockIdx.y, n = p.ny;
    const int yb = p.y0 + blockIdx.x * kTriLowRows;
    const int k = lane;
    const bool kin = k < p.nx;
    const bool has_r = p.R && k < p.lowkx;
    const int L = n < kTriLowL ? n : kTriLowL;
    for (int row = warp; row < kTriLowRows; row += kTriLowWarps) {
        const int y = yb + row;
        const bool yin = y < p.y1;
        const double af = (kin && yin) ? (double)__ldg(p.A + ((size_t)c * n + y) * p.nx + k) : 0.0;
        Af[row][lane] = af;
        Dx[row][lane] = (has_r && yin) ? -2.0 * __ldg(p.R + ((size_t)c * p.lowkx + k) * n + y) - af : 0.0;
    }
    if (warp == 0) {  // lane = row
        const int y = yb + lane;
        const double ph = (double)(y + 1) / (double)(n + 1);
        const double twoc = 2.0 * cospi(ph);
        double s0 = 0.0, s1 = (y < p.y1) ? sinpi(ph) : 0.0;  // rows past the end contribute nothing (their a are zero too)
        SCB_UNROLL
        for (int l = 0; l < kTriLowL; ++l) {
            S[lane][l] = s1;
            const double s2 = twoc * s1 - s0;
            s0 = s1;
            s1 = s2;
        }
    }
    __syncthreads();
    constexpr int NL = kTriLowL / kTriLowWarps;  // frequencies per thread
    double tf[NL], te[NL];
    SCB_UNROLL
    for (int j = 0; j < NL; ++j) tf[j] = te[j] = 0.0;
    for (int row = 0; row < kTriLowRows; ++row) {
        const double af = Af[row][lane], dx = Dx[row][lane];
        SCB_UNROLL
        for (int j = 0; j < NL; ++j) {
            const double s = S[row][warp + kTriLowWarps * j];
            tf[j] += s * af;
            te[j] += s * dx;
        }
    }
    if (kin) {
        const float fxk = __ldg(p.fx + k);
        SCB_UNROLL
        for (int j = 0; j < NL; ++j) {
            const int l = warp + kTriLowWarps * j;
            if (l >= L) continue;
            const double sf = tf[j], se = te[j];
            const double i32 = 1.0 / (double)__fsub_rn(__fadd_rn(fxk, __ldg(p.fy + l)), 4.0f);             // OpenCV: (filter_X + filter_Y) - 4 in float32
            const double iex = 1.0 / ((double)fxk + 2.0 * cospi((double)(l + 1) / (double)(n + 1)) - 4.0);  // what the tridiagonal solve divides by
            const double w = -(2.0 / (double)(n + 1)) * ((sf + se) * i32 - sf * iex);
            atomicAdd(p.W + ((size_t)c * kTriLowL + l) * kTriLowK + k, w);
        }
    }
}

// grid = (ceil(ny / kTriLowRows), 3), block = 32 x kTriLowWarps
__global__ void __launch_bounds__(32 * kTriLowWarps) tri_lowapply_kernel(TriLowParams p) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, c = blockIdx.y, n = p.ny;
    const int k = lane;
    if (k >= p.nx) return;
    const int L = n < kTriLowL ? n : kTriLowL;
    double w[kTriLowL];
    SCB_UNROLL
    for (int l = 0; l < kTriLowL; ++l) w[l] = (l < L) ? p.W[((size_t)c * kTriLowL + l) * kTriLowK + k] : 0.0;
    for (int sl = 1; sl < p.w_slots; ++sl) {  // row-sharded solves: the other ranks' partials (the static inner loop keeps w[] in registers)
        const double* ws = p.W + (size_t)sl * 3 * kTriLowL * kTriLowK;
        SCB_UNROLL
        for (int l = 0; l < kTriLowL; ++l)
            if (l < L) w[l] += ws[((size_t)c * kTriLowL + l) * kTriLowK + k];
    }
    const int yb = p.y0 + blockIdx.x * kTriLowRows;
    for (int y = yb + warp; y < yb + kTriLowRows && y < p.y1; y += kTriLowWarps) {
        double sum = p.Y64[((size_t)c * n + y) * kTriLowK + k];
        const double ph = (double)(y + 1) / (double)(n + 1);
        const double twoc = 2.0 * cospi(ph);
        double s0 = 0.0, s1 = sinpi(ph);
        SCB_UNROLL
        for (int l = 0; l < kTriLowL; ++l) {
            sum += w[l] * s1;
            const double s2 = twoc * s1 - s0;
            s0 = s1;
            s1 = s2;
        }
        p.Ct[((size_t)c * n + y) * p.nx + k] = (float)sum;
    }
}

}  // namespace scb
