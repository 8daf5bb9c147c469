// CI check (g++ -DSCB_EMU): rhs_fold2_kernel writes the SAME digit planes as rhs_fold_kernel<2>, byte for byte, on random images and
// binary masks of several shapes -- even / odd line lengths, lengths that leave a straddling thread, masks that touch the ROI
// border region, unaligned image origins.  rhs_fold_kernel<2> itself is pinned against the oracle's right-hand side through the
// parity tests (tests/test_pipeline.py); this harness transfers that pin to the packed-lane kernel.
//   build + run: tests/test_kernel_variants.py
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>

#include "scb_i8.h"
#include "scb_kernels.cuh"

using namespace scb;

static int run_case(int w, int h, int mask_kind, unsigned seed, int d_off, int s_off) {
    const int nx = w - 2, ny = h - 2;
    std::mt19937 rng(seed);
    const long long d_pitch = 3LL * (w + 7) + 5, s_pitch = 3LL * (w + 2) + 1, e_pitch = (w + 15) / 16 * 16;
    std::vector<unsigned char> D((size_t)d_pitch * (h + 1) + 64), S((size_t)s_pitch * (h + 1) + 64), E((size_t)e_pitch * h + 64, 0);
    for (auto& v : D) v = (unsigned char)(rng() & 255);
    for (auto& v : S) v = (unsigned char)(rng() & 255);
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            bool in = false;
            switch (mask_kind) {
                case 0: {  // ellipse well inside
                    const double dx = (x - w / 2.0) / (w * 0.35), dy = (y - h / 2.0) / (h * 0.35);
                    in = dx * dx + dy * dy < 1.0;
                    break;
                }
                case 1: in = x >= 3 && x < w - 3 && y >= 3 && y < h - 3; break;  // as close to the border as an eroded mask gets
                case 2: in = ((rng() >> 7) & 3) != 0; break;                       // noise: every tap pattern
                case 3: in = false; break;                                         // all dst
                case 4: in = x >= 1 && x < w - 1 && y >= 1 && y < h - 1; break;    // all src up to the ring (exercises the border terms with src taps)
            }
            E[(size_t)y * e_pitch + x] = in ? 255 : 0;
        }
    const I8Geom g = i8_geometry(nx);
    const int lines = 3 * ny, m_rows = i8_m_rows(lines);
    const size_t plane_bytes = (size_t)2 * 2 * m_rows * g.kpad;
    std::vector<signed char> P1(plane_bytes, 0x55), P2(plane_bytes, 0x2a);
    std::vector<float> L1(m_rows, -1.f), L2(m_rows, -2.f);
    RhsFoldParams f;
    f.st.D = D.data() + d_off;
    f.st.d_pitch = d_pitch;
    f.st.S = S.data() + s_off;
    f.st.s_pitch = s_pitch;
    f.st.E = E.data();
    f.st.e_pitch = e_pitch;
    f.st.w = w;
    f.st.h = h;
    f.nx = nx;
    f.ny = ny;
    f.kpar0 = g.kpar[0];
    f.kpad = g.kpad;
    f.lines = lines;
    f.m_rows = m_rows;
    f.scale = 1.0f;
    f.y0 = 0;
    const int rows = (m_rows + 2) / 3;
    const dim3 grid((g.kpad / 4 + kRhsThreads - 1) / kRhsThreads, rows);
    f.planes = P1.data();
    f.lscale = L1.data();
    SCB_LAUNCH(rhs_fold_kernel<2>, grid, dim3(kRhsThreads), 0, 0, f);
    f.planes = P2.data();
    f.lscale = L2.data();
    SCB_LAUNCH(rhs_fold2_kernel, grid, dim3(kRhsThreads), 0, 0, f);
    long long bad = 0, first = -1;
    for (size_t i = 0; i < plane_bytes; ++i)
        if (P1[i] != P2[i]) {
            if (first < 0) first = (long long)i;
            ++bad;
        }
    int badl = 0;
    for (int i = 0; i < m_rows; ++i) badl += L1[i] != L2[i];
    if (bad || badl) {
        const long long row = first / g.kpad, col = first % g.kpad;
        std::printf("FAIL w=%d h=%d mask=%d seed=%u: %lld differing bytes (first at plane row %lld [plane %lld, line %lld], j=%lld: %d vs %d), %d lscale\n", w, h, mask_kind,
                    seed, bad, row, row / m_rows, row % m_rows, col, (int)P1[first], (int)P2[first], badl);
        return 1;
    }
    return 0;
}

int main() {
    int fails = 0, cases = 0;
    const int widths[] = {66, 67, 68, 69, 70, 71, 72, 73, 131, 258, 515, 600};
    const int heights[] = {5, 9, 20};
    unsigned seed = 1;
    for (int w : widths)
        for (int h : heights)
            for (int mk = 0; mk < 5; ++mk) {
                fails += run_case(w, h, mk, seed, (int)(seed % 4), (int)((seed / 4) % 4));
                ++seed;
                ++cases;
            }
    std::printf("%d cases, %d failed\n", cases, fails);
    return fails ? 1 : 0;
}
