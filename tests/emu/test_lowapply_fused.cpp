// CI check (g++ -DSCB_EMU): i8_digitize2_kernel with the fused low-frequency block == tri_lowapply_kernel followed by the plain
// i8_digitize2_kernel, byte for byte (digit planes and line scales), for even / odd line lengths and few / many rows.
//   build + run: tests/test_kernel_variants.py
#include <cstdio>
#include <cstring>
#include <random>
#include <vector>

#include "scb_i8.cu"
#include "scb_tri.cuh"

using namespace scb;

static int run_case(int nx, int ny, unsigned seed) {
    std::mt19937 rng(seed);
    std::uniform_real_distribution<double> dist(-1.0, 1.0);
    const I8Geom g = i8_geometry(nx);
    const int lines = 3 * ny, m_rows = i8_m_rows(lines);
    std::vector<float> Ct((size_t)3 * ny * nx);
    for (auto& v : Ct) v = (float)(1000.0 * dist(rng));
    std::vector<double> Y64((size_t)3 * ny * kTriLowK), W((size_t)3 * kTriLowL * kTriLowK);
    for (auto& v : Y64) v = 5000.0 * dist(rng);
    for (auto& v : W) v = 300.0 * dist(rng);
    std::vector<float> fx(nx, 0.f), fy(ny, 0.f);
    std::vector<float> CtA = Ct;  // lowapply writes its columns k < 32 into this copy
    TriLowParams l;
    l.nx = nx;
    l.ny = ny;
    l.A = nullptr;
    l.R = nullptr;
    l.lowkx = 0;
    l.Y64 = Y64.data();
    l.fx = fx.data();
    l.fy = fy.data();
    l.W = W.data();
    l.w_slots = 1;
    l.Ct = CtA.data();
    l.y0 = 0;
    l.y1 = ny;
    SCB_LAUNCH(tri_lowapply_kernel, dim3((ny + kTriLowRows - 1) / kTriLowRows, 3), dim3(32 * kTriLowWarps), 0, 0, l);
    const size_t bytes = i8_adig_bytes(g, lines, 4);
    std::vector<signed char> P1(bytes, 0x55), P2(bytes, 0x2a);
    std::vector<float> L1(m_rows, -1.f), L2(m_rows, -2.f);
    I8DigitizeParams d{};
    d.g = g;
    d.in_plane = (long long)ny * nx;
    d.in_pitch = nx;
    d.lpc = ny;
    d.lines = lines;
    d.m_rows = m_rows;
    d.line0 = 0;
    d.line1 = m_rows;
    d.fixed_scale = 1.0f;
    d.per_line = 1;
    d.in = CtA.data();
    d.a = P1.data();
    d.lscale = L1.data();
    SCB_LAUNCH(i8_digitize2_kernel<4>, dim3(m_rows), dim3(kI8DigThreads), 0, 0, d);
    d.in = Ct.data();  // the columns k < 32 of this one were never written by lowapply
    d.a = P2.data();
    d.lscale = L2.data();
    d.low_w = W.data();
    d.low_y64 = Y64.data();
    d.low_k = kTriLowK;
    d.low_l = kTriLowL;
    d.low_nk = nx < kTriLowK ? nx : kTriLowK;
    d.low_nl = ny < kTriLowL ? ny : kTriLowL;
    SCB_LAUNCH(i8_digitize2_kernel<4>, dim3(m_rows), dim3(kI8DigThreads), 0, 0, d);
    const bool bad = std::memcmp(P1.data(), P2.data(), bytes) != 0, badl = std::memcmp(L1.data(), L2.data(), m_rows * sizeof(float)) != 0;
    if (bad || badl) {
        std::printf("FAIL nx=%d ny=%d: planes differ %d, scales differ %d\n", nx, ny, (int)bad, (int)badl);
        return 1;
    }
    return 0;
}

int main() {
    int fails = 0, cases = 0;
    unsigned seed = 1;
    for (int nx : {64, 65, 66, 97, 128, 500, 1808, 2048})
        for (int ny : {1, 7, 31, 32, 33, 100}) {
            fails += run_case(nx, ny, seed++);
            ++cases;
        }
    std::printf("%d cases, %d failed\n", cases, fails);
    return fails ? 1 : 0;
}
