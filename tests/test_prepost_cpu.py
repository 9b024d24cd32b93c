"""The oracle of the device-side pre-processing (oracle/warp_ref.py) pinned against OpenCV itself."""
import numpy as np
import pytest

from oracle import warp_ref


def test_warp_affine_oracle_equals_opencv():
    cv2 = pytest.importorskip("cv2")
    from codenet_b200.compat.detector import get_affine_transform
    rng = np.random.default_rng(0)
    for (h, w, ih, iw) in [(300, 400, 256, 256), (375, 500, 512, 512), (500, 375, 512, 512), (512, 512, 512, 512), (97, 203, 128, 160)]:
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        c, s = np.array([w / 2., h / 2.], np.float32), max(h, w) * 1.0
        M = get_affine_transform(c, s, 0, [iw, ih])                  # the matrix pre_process hands to cv2 (base_detector.py:61)
        np.testing.assert_array_equal(warp_ref.warp_affine(img, M, (iw, ih)), cv2.warpAffine(img, M, (iw, ih), flags=cv2.INTER_LINEAR))
    # a general (rotated, sheared) matrix as well
    M = np.array([[0.83, -0.21, 14.3], [0.17, 0.91, -7.9]])
    img = rng.integers(0, 256, (120, 90, 3), dtype=np.uint8)
    np.testing.assert_array_equal(warp_ref.warp_affine(img, M, (100, 110)), cv2.warpAffine(img, M, (100, 110), flags=cv2.INTER_LINEAR))
