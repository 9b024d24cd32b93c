"""TEST INFRASTRUCTURE ONLY.  numpy restatement of cv2.warpAffine(src, M, dsize, flags=cv2.INTER_LINEAR) with the default
constant border 0 for uint8 HWC images -- the transform of BaseDetector.pre_process (lib/detectors/base_detector.py:61-65).

OpenCV (imgwarp.cpp, third-party; the reference pins no version, this container has cv2 4.13) inverts the 2x3 matrix in
double, forms the source coordinate of every destination pixel in 1/1024 pixel with the column term and the row term rounded
SEPARATELY, adds 16 and drops to 1/32 pixel; the bilinear weights are the products of the 1/32 fractions scaled to 2^15 and
the result is (sum + 2^14) >> 15.  Pinned against cv2 itself by tests/test_prepost_cpu.py."""
import numpy as np


def warp_affine(src, M, dsize):
    dw, dh = dsize
    M = np.asarray(M, np.float64).copy().reshape(2, 3)
    D = M[0, 0] * M[1, 1] - M[0, 1] * M[1, 0]
    D = 1.0 / D if D != 0 else 0.0
    A11, A22 = M[1, 1] * D, M[0, 0] * D
    M[0, 0] = A11; M[0, 1] *= -D; M[1, 0] *= -D; M[1, 1] = A22
    b1 = -M[0, 0] * M[0, 2] - M[0, 1] * M[1, 2]
    b2 = -M[1, 0] * M[0, 2] - M[1, 1] * M[1, 2]
    M[0, 2] = b1; M[1, 2] = b2
    H, W = src.shape[:2]
    xs = np.arange(dw)
    adelta = np.rint(M[0, 0] * xs * 1024.0).astype(np.int64)
    bdelta = np.rint(M[1, 0] * xs * 1024.0).astype(np.int64)
    dst = np.zeros((dh, dw, src.shape[2]), np.uint8)
    for y in range(dh):
        X0 = int(np.rint((M[0, 1] * y + M[0, 2]) * 1024.0)) + 16
        Y0 = int(np.rint((M[1, 1] * y + M[1, 2]) * 1024.0)) + 16
        X, Y = (X0 + adelta) >> 5, (Y0 + bdelta) >> 5
        sx, sy = np.clip(X >> 5, -32768, 32767), np.clip(Y >> 5, -32768, 32767)
        fx, fy = X & 31, Y & 31

        def px(yy, xx):
            ok = (yy >= 0) & (yy < H) & (xx >= 0) & (xx < W)
            return src[np.clip(yy, 0, H - 1), np.clip(xx, 0, W - 1)].astype(np.int64) * ok[:, None]

        v = (px(sy, sx) * ((32 - fy) * (32 - fx) * 32)[:, None] + px(sy, sx + 1) * ((32 - fy) * fx * 32)[:, None] +
             px(sy + 1, sx) * (fy * (32 - fx) * 32)[:, None] + px(sy + 1, sx + 1) * (fy * fx * 32)[:, None])
        dst[y] = np.clip((v + (1 << 14)) >> 15, 0, 255).astype(np.uint8)
    return dst
