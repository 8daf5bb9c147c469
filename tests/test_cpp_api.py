"""The C++ host API (include/seamless_clone.hpp, OpenCV's seamlessClone signature) against the oracle."""
import os
import subprocess

import numpy as np
import pytest

from oracle import seamless_oracle as so


@pytest.mark.gpu
@pytest.mark.parametrize("which", ["scb_mat", "cv_mat_overload"])
def test_cpp_seamless_clone(tmp_path, cuda_lib, which):
    """scb::seamlessClone on scb::Mat, and the cv::Mat overload (compiled against the mock <opencv2/core.hpp>: no OpenCV headers here)."""
    import __graft_entry__ as ge

    exe = ge.build_cpp_test() if which == "scb_mat" else ge.build_cvmat_test()
    assert exe and os.path.exists(exe)
    src, dst, mask, p = so.make_config("small", 21)
    ref = so.restate(src, dst, mask, p, transform="f64")
    for name, a in (("src", src), ("dst", dst), ("mask", mask)):
        a.tofile(tmp_path / f"{name}.bin")
    out = tmp_path / "out.bin"
    r = subprocess.run([exe, str(tmp_path / "src.bin"), str(src.shape[0]), str(src.shape[1]), str(tmp_path / "dst.bin"), str(dst.shape[0]), str(dst.shape[1]),
                        str(tmp_path / "mask.bin"), str(p[0]), str(p[1]), str(out)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    blend = np.fromfile(out, np.uint8).reshape(dst.shape)
    cmp = so.compare_u8(blend, ref.blend)
    assert cmp["max_abs"] <= 1 and cmp["n_diff"] <= 4, cmp


def test_cpp_api_compiles_against_the_header(cuda_lib):
    """CPU: the header-only C++ API builds and links against libscb.so (no GPU needed to link)."""
    import __graft_entry__ as ge

    exe = ge.build_cpp_test(force=True)
    assert exe and os.path.exists(exe)
    exe = ge.build_cvmat_test(force=True)  # the cv::Mat overload (SCB_WITH_OPENCV) against the mock <opencv2/core.hpp>
    assert exe and os.path.exists(exe)
