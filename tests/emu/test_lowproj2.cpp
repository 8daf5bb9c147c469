// CI check (g++ -DSCB_EMU): tri_lowproj2_kernel accumulates the same low-frequency block coefficients W as tri_lowproj_kernel
// (float64 sums in a different order: agreement to 2e-9 of the largest coefficient, below float32 resolution; an indexing error would show at O(1)), for several shapes and row ranges.
//   build + run: tests/test_kernel_variants.py
#include <cmath>
#include <cstdio>
#include <random>
#include <vector>

#include "scb_tri.cuh"

using namespace scb;

static int run_case(int nx, int ny, int lowkx, int y0, int y1, unsigned seed) {
    std::mt19937 rng(seed);
    std::uniform_real_distribution<double> dist(-1.0, 1.0);
    std::vector<float> A((size_t)3 * ny * nx), fx(nx), fy(ny);
    for (auto& v : A) v = (float)(1000.0 * dist(rng));
    for (int k = 0; k < nx; ++k) fx[k] = 2.0f * (float)std::cos(M_PI / (double)(nx + 1) * (double)(k + 1));
    for (int l = 0; l < ny; ++l) fy[l] = 2.0f * (float)std::cos(M_PI / (double)(ny + 1) * (double)(l + 1));
    std::vector<double> R((size_t)3 * (lowkx ? lowkx : 1) * ny);
    for (auto& v : R) v = 500.0 * dist(rng);
    std::vector<double> W1((size_t)3 * kTriLowL * kTriLowK, 0.0), W2(W1);
    TriLowParams l;
    l.nx = nx;
    l.ny = ny;
    l.A = A.data();
    l.R = lowkx ? R.data() : nullptr;
    l.lowkx = lowkx;
    l.Y64 = nullptr;
    l.fx = fx.data();
    l.fy = fy.data();
    l.w_slots = 1;
    l.Ct = nullptr;
    l.y0 = y0;
    l.y1 = y1;
    const dim3 grid((y1 - y0 + kTriLowRows - 1) / kTriLowRows, 3), block(32 * kTriLowWarps);
    l.W = W1.data();
    SCB_LAUNCH(tri_lowproj_kernel, grid, block, 0, 0, l);
    l.W = W2.data();
    SCB_LAUNCH(tri_lowproj2_kernel, grid, block, 0, 0, l);
    double big = 0.0, err = 0.0;
    for (size_t i = 0; i < W1.size(); ++i) {
        big = std::fmax(big, std::fabs(W1[i]));
        err = std::fmax(err, std::fabs(W1[i] - W2[i]));
    }
    if (!(err <= 2e-9 * big) || big == 0.0) {  // (W is a difference of nearly equal terms: reordering the float64 sums shows at up to ~2e-10 of the largest coefficient)
        std::printf("FAIL nx=%d ny=%d lowkx=%d rows [%d, %d): max |W| %.3e, max difference %.3e\n", nx, ny, lowkx, y0, y1, big, err);
        return 1;
    }
    return 0;
}

int main() {
    int fails = 0, cases = 0;
    unsigned seed = 1;
    for (int nx : {5, 31, 32, 70, 300})
        for (int ny : {3, 31, 32, 33, 100, 333})
            for (int lowkx : {0, 4}) {
                fails += run_case(nx, ny, lowkx < nx ? lowkx : 0, 0, ny, seed++);
                ++cases;
            }
    fails += run_case(200, 333, 4, 64, 200, seed++);  // a row shard
    fails += run_case(200, 333, 4, 70, 333, seed++);
    cases += 2;
    std::printf("%d cases, %d failed\n", cases, fails);
    return fails ? 1 : 0;
}
