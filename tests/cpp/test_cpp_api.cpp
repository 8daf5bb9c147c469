// C++ host API (include/seamless_clone.hpp) driven the way a cv::seamlessClone call site would be.
//   test_cpp_api src.bin hs ws dst.bin H W mask.bin px py out.bin
// reads raw interleaved u8 images, clones, writes the blend; also checks the error behaviour.
// tests/test_cpp_api.py compares out.bin with the oracle.
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <vector>

#include "seamless_clone.hpp"

static std::vector<unsigned char> slurp(const char* path, size_t n) {
    std::vector<unsigned char> v(n);
    std::ifstream f(path, std::ios::binary);
    f.read((char*)v.data(), (std::streamsize)n);
    if ((size_t)f.gcount() != n) {
        std::fprintf(stderr, "short read on %s\n", path);
        std::exit(2);
    }
    return v;
}

int main(int argc, char** argv) {
    if (argc != 11) {
        std::fprintf(stderr, "usage: %s src.bin hs ws dst.bin H W mask.bin px py out.bin\n", argv[0]);
        return 2;
    }
    const int hs = std::atoi(argv[2]), ws = std::atoi(argv[3]), H = std::atoi(argv[5]), W = std::atoi(argv[6]);
    const int px = std::atoi(argv[8]), py = std::atoi(argv[9]);
    auto sb = slurp(argv[1], (size_t)hs * ws * 3), db = slurp(argv[4], (size_t)H * W * 3), mb = slurp(argv[7], (size_t)hs * ws);
    const auto db_before = db;
    const auto mb_before = mb;
    scb::Mat src(hs, ws, 3, sb.data()), dst(H, W, 3, db.data()), mask(hs, ws, 1, mb.data()), blend;
    try {
        scb::seamlessClone(src, dst, mask, scb::Point(px, py), blend, scb::NORMAL_CLONE);
    } catch (const scb::Exception& e) {
        std::fprintf(stderr, "unexpected: %s\n", e.what());
        return 1;
    }
    if (db != db_before || mb != mb_before) {
        std::fprintf(stderr, "dst or mask was modified\n");
        return 1;
    }
    if (blend.rows != H || blend.cols != W || blend.channels() != 3 || blend.data == dst.data) {
        std::fprintf(stderr, "blend has the wrong shape or aliases dst\n");
        return 1;
    }
    // error behaviour: ROI outside dst -> code 2 (OpenCV: -215 assertion); other flags -> code 3 (no CPU fallback)
    int seen = 0;
    try {
        scb::Mat b2;
        scb::seamlessClone(src, dst, mask, scb::Point(1, 1), b2, scb::NORMAL_CLONE);
    } catch (const scb::Exception& e) {
        seen += (e.code == SCB_ERR_ROI_OUT_OF_BOUNDS);
    }
    try {
        scb::Mat b2;
        scb::seamlessClone(src, dst, mask, scb::Point(px, py), b2, 4 /* not a cv::seamlessClone flag */);
    } catch (const scb::Exception& e) {
        seen += (e.code == SCB_ERR_UNSUPPORTED);
    }
    if (seen != 2) {
        std::fprintf(stderr, "error behaviour differs (%d of 2)\n", seen);
        return 1;
    }
    std::ofstream(argv[10], std::ios::binary).write((const char*)blend.data, (std::streamsize)((size_t)H * W * 3));
    std::puts("ok");
    return 0;
}
