// The cv::Mat overload of include/seamless_clone.hpp (SCB_WITH_OPENCV) -- the call a cv::seamlessClone call site
// (/root/reference/seamlessClone-OpenCV/seamlessClone_OpenCV.cpp:104,110) is re-pointed at -- compiled against the MOCK
// <opencv2/core.hpp> of tests/cpp/mock_opencv (no OpenCV headers exist in this image).
//   test_cvmat_overload src.bin hs ws dst.bin H W mask.bin px py out.bin
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <vector>

#define SCB_WITH_OPENCV 1
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
    if (argc != 11) return 2;
    const int hs = std::atoi(argv[2]), ws = std::atoi(argv[3]), H = std::atoi(argv[5]), W = std::atoi(argv[6]);
    const int px = std::atoi(argv[8]), py = std::atoi(argv[9]);
    auto sb = slurp(argv[1], (size_t)hs * ws * 3), db = slurp(argv[4], (size_t)H * W * 3), mb = slurp(argv[7], (size_t)hs * ws);
    const auto db_before = db;
    cv::Mat src(hs, ws, CV_8UC3, sb.data()), dst(H, W, CV_8UC3, db.data()), mask(hs, ws, CV_8UC1, mb.data()), blend;
    try {
        scb::seamlessClone(src, dst, mask, cv::Point(px, py), blend, scb::NORMAL_CLONE);
    } catch (const std::exception& e) {
        std::fprintf(stderr, "unexpected: %s\n", e.what());
        return 1;
    }
    if (db != db_before || blend.rows != H || blend.cols != W || blend.type() != CV_8UC3 || blend.data == dst.data) {
        std::fprintf(stderr, "dst modified, or blend has the wrong shape / aliases dst\n");
        return 1;
    }
    std::ofstream o(argv[10], std::ios::binary);
    o.write((const char*)blend.data, (std::streamsize)((size_t)H * W * 3));
    return 0;
}
