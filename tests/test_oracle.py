"""CPU: the oracle against the golden vectors (made by cv2.seamlessClone, tests/golden/make_golden.py)
and, when cv2 is importable, against cv2.seamlessClone live."""
import numpy as np
import pytest

from oracle import seamless_oracle as so
from tests import common

try:
    import cv2  # noqa: F401

    HAVE_CV = True
except Exception:  # pragma: no cover
    HAVE_CV = False


@pytest.mark.parametrize("name", common.golden_names())
def test_restatement_bit_exact_vs_golden(name):
    if not HAVE_CV:
        pytest.skip("cv2.dft is needed for the bit-exact transform back end")
    z = common.load_golden(name)
    tr = so.restate(z["src"], z["dst"], z["mask"], tuple(z["p"]), flags=z["flags"], transform="cv")
    g = tr.geom
    assert [g.x, g.y, g.w, g.h, g.rx, g.ry] == list(z["geom"])
    assert np.array_equal(tr.blend[g.ry : g.ry + g.h, g.rx : g.rx + g.w], z["blend_roi"])
    assert np.array_equal(tr.rhs, z["rhs"])
    assert np.array_equal(tr.eroded, z["eroded"])


@pytest.mark.parametrize("name", common.golden_names())
def test_float64_back_end_within_tolerance(name):
    z = common.load_golden(name)
    tr = so.restate(z["src"], z["dst"], z["mask"], tuple(z["p"]), flags=z["flags"], transform="f64")
    g = tr.geom
    assert np.array_equal(tr.rhs, z["rhs"])  # integer stencil needs no cv2
    assert so.rel_linf(tr.spectrum, z["spectrum"]) < 1e-5
    assert so.rel_linf(tr.solved, z["solved"]) < 1e-4
    cmp = so.compare_u8(tr.blend[g.ry : g.ry + g.h, g.rx : g.rx + g.w], z["blend_roi"])
    assert cmp["max_abs"] <= 1 and cmp["n_diff"] <= common.allowed_mismatches(z["rhs"].size)


@pytest.mark.skipif(not HAVE_CV, reason="needs cv2")
@pytest.mark.parametrize("cfg,seed", [("small", 1), ("small", 2), ("cfg1", 0)])
def test_restatement_bit_exact_vs_cv2_live(cfg, seed):
    src, dst, mask, p = so.make_config(cfg, seed)
    m0 = mask.copy()
    ref = so.cv_reference(src, dst, mask, p)
    assert np.array_equal(mask, m0)
    tr = so.restate(src, dst, mask, p, transform="cv")
    assert np.array_equal(tr.blend, ref)


@pytest.mark.skipif(not HAVE_CV, reason="needs cv2")
def test_opencv_semantics_probed():
    """Facts of cv::seamlessClone the boundary relies on."""
    import cv2

    src, dst, mask, p = so.make_config("small", 7)
    # p is the centre of the mask BOUNDING BOX; odd sizes use truncating division
    tr = so.restate(src, dst, mask, p)
    assert tr.geom.rx == p[0] - tr.geom.w // 2 and tr.geom.ry == p[1] - tr.geom.h // 2
    # ROI outside dst raises
    with pytest.raises(cv2.error):
        cv2.seamlessClone(src, dst, mask.copy(), (2, 2), cv2.NORMAL_CLONE)
    with pytest.raises(so.OracleError):
        so.restate(src, dst, mask, (2, 2))
    # all-zero mask: blend == dst
    z = np.zeros_like(mask)
    assert np.array_equal(so.restate(src, dst, z, p).blend, dst)
    # 3-channel mask == its grey conversion
    m3 = np.repeat(mask[:, :, None], 3, axis=2)
    assert np.array_equal(so.cv_reference(src, dst, m3, p), so.cv_reference(src, dst, mask, p))


def test_eigen_filters_recipe():
    fx, fy = so.filters(510, 382)
    assert fx.dtype == np.float32 and fx.shape == (508,) and fy.shape == (380,)
    den = so.denominator(510, 382)
    assert den.dtype == np.float32 and (den < 0).all()


def test_generators_are_deterministic():
    a = so.make_config("small", 3)
    b = so.make_config("small", 3)
    for x, y in zip(a[:3], b[:3]):
        assert np.array_equal(x, y)
    jobs = so.make_batch_jobs(5, seed=1)
    assert jobs == so.make_batch_jobs(5, seed=1)


@pytest.mark.skipif(not HAVE_CV, reason="needs cv2")
@pytest.mark.parametrize("flags", [so.MIXED_CLONE, so.MONOCHROME_TRANSFER, so.NORMAL_CLONE_WIDE, so.MIXED_CLONE_WIDE, so.MONOCHROME_TRANSFER_WIDE])
@pytest.mark.parametrize("grey", [False, True])
def test_other_clone_flags_bit_exact_vs_cv2_live(flags, grey):
    """MIXED_CLONE / MONOCHROME_TRANSFER (gradient selection) and the _WIDE placement, pinned against cv2.seamlessClone."""
    rng = np.random.default_rng(31 + flags)
    src = so.smooth_rand(rng, 61, 83, 2.0)
    dst = so.smooth_rand(rng, 140, 170, 2.0)
    mask = so.ellipse_mask(61, 83, 36.0, 27.0, 30.0, 20.0, 0.3)
    if grey:
        mask = ((mask > 0) * rng.integers(1, 256, size=mask.shape)).astype(np.uint8)
    p = (88, 71)
    ref = so.cv_reference(src, dst, mask, p, flags)
    tr = so.restate(src, dst, mask, p, flags=flags, transform="cv")
    assert np.array_equal(tr.blend, ref)


@pytest.mark.skipif(not HAVE_CV, reason="needs cv2")
def test_gray_conversion_bit_exact_vs_cv2():
    import cv2

    img = np.random.default_rng(3).integers(0, 256, size=(97, 131, 3), dtype=np.uint8)
    assert np.array_equal(so.bgr2gray_u8(img), cv2.cvtColor(img, cv2.COLOR_BGR2GRAY))
