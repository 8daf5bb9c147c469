// CI check (g++ -DSCB_EMU): tri_solve_smem_kernel and tri_solve_smem2_kernel produce the SAME bits as tri_solve_kernel -- Ct for the float columns, Y64 for the
// float64 block -- on random right-hand sides, for column counts with ragged last tiles, row counts with short last segments,
// nx below / across the float64 block, tables in shared memory and in global memory.
//   build + run: tests/test_kernel_variants.py
#include <cmath>
#include <cstdio>
#include <cstring>
#include <random>
#include <vector>

#include "scb_tri.cuh"

using namespace scb;

static int run_case(int nx, int ny, unsigned seed, int tables, int variant) {
    std::mt19937 rng(seed);
    std::uniform_real_distribution<float> dist(-1000.f, 1000.f);
    const int L = tri_seg_len(ny), rows = L + 1, pm = (nx + 3) / 4 * 4;
    std::vector<double> theta(nx);
    for (int k = 0; k < nx; ++k) {
        const float fx = 2.0f * (float)std::cos(3.14159265358979323846 / (double)(nx + 1) * (double)(k + 1));  // OpenCV's filter_X
        theta[k] = std::acosh((4.0 - (double)fx) / 2.0);
    }
    std::vector<float> m32((size_t)rows * pm), p32((size_t)rows * pm);
    std::vector<double> m64((size_t)rows * kTriLowK), p64((size_t)rows * kTriLowK);
    TriTableParams tp;
    tp.theta = theta.data();
    tp.nx = nx;
    tp.rows = rows;
    tp.pm = pm;
    tp.m32 = m32.data();
    tp.p32 = p32.data();
    tp.m64 = m64.data();
    tp.p64 = p64.data();
    SCB_LAUNCH(tri_table_kernel, dim3(8), dim3(256), 0, 0, tp);
    std::vector<float> A((size_t)3 * ny * nx);
    for (auto& v : A) v = dist(rng);
    std::vector<float> C1((size_t)3 * ny * nx, -7.f), C2((size_t)3 * ny * nx, -7.f);
    std::vector<double> Y1((size_t)3 * ny * kTriLowK, -7.0), Y2((size_t)3 * ny * kTriLowK, -7.0);
    TriSolveParams t;
    t.tab.m32 = m32.data();
    t.tab.p32 = p32.data();
    t.tab.pm = pm;
    t.tab.rows = rows;
    t.tab.m64 = m64.data();
    t.tab.p64 = p64.data();
    t.tab.theta = theta.data();
    t.nx = nx;
    t.ny = ny;
    t.A = A.data();
    t.x0 = 0;
    t.x1 = nx;
    t.seg_len = L;
    t.phase = 0;
    t.seg0 = 0;
    t.seg1 = kTriSegs;
    t.ends32 = nullptr;
    t.ends64 = nullptr;
    t.Ct = C1.data();
    t.Y64 = Y1.data();
    SCB_LAUNCH(tri_solve_kernel, dim3((nx + kTriCols - 1) / kTriCols, 3), dim3(kTriCols * kTriSegs), 0, 0, t);
    t.Ct = C2.data();
    t.Y64 = Y2.data();
    const size_t bytes = tri_smem_bytes(ny, tables != 0);
    if (!bytes) {
        std::printf("SKIP nx=%d ny=%d tables=%d: tile does not fit\n", nx, ny, tables);
        return 0;
    }
    const int nfloat = nx > kTriLowK ? (nx - kTriLowK + kTriCols - 1) / kTriCols : 0;
    if (variant == 2)
        SCB_LAUNCH(tri_solve_smem2_kernel, dim3(kTriLowK / kTriCols64 + nfloat, 3), dim3(kTriCols * kTriSegs), bytes, 0, t, tables);
    else
        SCB_LAUNCH(tri_solve_smem_kernel, dim3(kTriLowK / kTriCols64 + nfloat, 3), dim3(kTriCols * kTriSegs), bytes, 0, t, tables);
    long long bad = 0;
    for (int c = 0; c < 3; ++c)
        for (int y = 0; y < ny; ++y) {
            for (int k = kTriLowK; k < nx; ++k) {
                const size_t i = ((size_t)c * ny + y) * nx + k;
                bad += std::memcmp(&C1[i], &C2[i], 4) != 0;
            }
            for (int k = 0; k < kTriLowK && k < nx; ++k) {
                const size_t i = ((size_t)c * ny + y) * kTriLowK + k;
                bad += std::memcmp(&Y1[i], &Y2[i], 8) != 0;
            }
        }
    // nothing outside the solved entries may be touched
    for (int c = 0; c < 3; ++c)
        for (int y = 0; y < ny; ++y)
            for (int k = 0; k < kTriLowK && k < nx; ++k) bad += C2[((size_t)c * ny + y) * nx + k] != -7.f;
    if (bad) {
        std::printf("FAIL nx=%d ny=%d seed=%u tables=%d kernel %d: %lld differing values\n", nx, ny, seed, tables, variant, bad);
        return 1;
    }
    return 0;
}

int main() {
    int fails = 0, cases = 0;
    const int nxs[] = {5, 31, 32, 33, 47, 48, 70, 200};
    const int nys[] = {1, 3, 7, 16, 63, 64, 65, 130, 333};
    unsigned seed = 1;
    for (int nx : nxs)
        for (int ny : nys)
            for (int tables = 0; tables < 2; ++tables)
                for (int variant = 1; variant <= 2; ++variant) {
                    fails += run_case(nx, ny, seed++, tables, variant);
                    ++cases;
                }
    for (int variant = 1; variant <= 2; ++variant) {
        fails += run_case(100, 1337, seed++, 1, variant);  // the 4K clone's column length
        fails += run_case(40, 3070, seed++, 0, variant);   // the 8K clone's: tables stay in global memory
        cases += 2;
    }
    std::printf("%d cases, %d failed\n", cases, fails);
    return fails ? 1 : 0;
}
