// CI check (g++ -DSCB_EMU): i8_digitize2_kernel writes the SAME digit planes and line scales as i8_digitize_kernel, byte for byte:
// even / odd line lengths, fixed scale (forward pass, 2 and 4 digits) and per-line scale (inverse pass), pad lines, line ranges.
//   build + run: tests/test_kernel_variants.py
#include <cstdio>
#include <cstring>
#include <random>
#include <vector>

#include "scb_i8.cu"

using namespace scb;

template <int DA>
static int run_case(int n, int lpc, int per_line, unsigned seed) {
    std::mt19937 rng(seed);
    const I8Geom g = i8_geometry(n);
    const int lines = 3 * lpc, m_rows = i8_m_rows(lines), pitch = (n + 3) / 4 * 4 + 4;
    std::vector<float> in((size_t)3 * lpc * pitch);
    std::uniform_real_distribution<float> mag(-12.f, 12.f);
    for (int r = 0; r < 3 * lpc; ++r) {
        const float sc = per_line ? std::ldexp(1.0f, (int)mag(rng)) : 1.0f;
        for (int j = 0; j < pitch; ++j)
            in[(size_t)r * pitch + j] = per_line ? sc * (float)((int)(rng() % 20001) - 10000) / 7.0f : (DA == 2 ? (float)((int)(rng() % 3061) - 1530) : (float)((int)(rng() % 200001) - 100000) / 65536.0f);
    }
    if (per_line) std::fill(in.begin() + pitch, in.begin() + 2 * pitch, 0.f);  // an all-zero line
    const size_t bytes = i8_adig_bytes(g, lines, DA);
    std::vector<signed char> P1(bytes, 0x55), P2(bytes, 0x2a);
    std::vector<float> L1(m_rows, -1.f), L2(m_rows, -2.f);
    I8DigitizeParams d{};
    d.g = g;
    d.in = in.data();
    d.in_plane = (long long)lpc * pitch;
    d.in_pitch = pitch;
    d.lpc = lpc;
    d.lines = lines;
    d.m_rows = m_rows;
    d.line0 = 0;
    d.line1 = m_rows;
    d.fixed_scale = per_line ? 1.0f : (DA == 2 ? 1.0f : 65536.0f);
    d.per_line = per_line;
    d.a = P1.data();
    d.lscale = L1.data();
    SCB_LAUNCH(i8_digitize_kernel<DA>, dim3(m_rows), dim3(kI8DigThreads), 0, 0, d);
    d.a = P2.data();
    d.lscale = L2.data();
    SCB_LAUNCH(i8_digitize2_kernel<DA>, dim3(m_rows), dim3(kI8DigThreads), 0, 0, d);
    const long long bad = std::memcmp(P1.data(), P2.data(), bytes) != 0, badl = std::memcmp(L1.data(), L2.data(), m_rows * sizeof(float)) != 0;
    if (bad || badl) {
        std::printf("FAIL DA=%d n=%d lpc=%d per_line=%d: planes differ %lld, scales differ %lld\n", DA, n, lpc, per_line, bad, badl);
        return 1;
    }
    return 0;
}

int main() {
    int fails = 0, cases = 0;
    unsigned seed = 1;
    const int ns[] = {64, 65, 66, 67, 127, 128, 129, 255, 257, 900, 1023, 1808, 2047, 2048};
    for (int n : ns)
        for (int lpc : {1, 5, 43}) {
            fails += run_case<2>(n, lpc, 0, seed++);
            fails += run_case<4>(n, lpc, 0, seed++);
            fails += run_case<4>(n, lpc, 1, seed++);
            cases += 3;
        }
    std::printf("%d cases, %d failed\n", cases, fails);
    return fails ? 1 : 0;
}
