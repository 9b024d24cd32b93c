"""Host-side mirror of the reference's `soft_nms` (lib/models/external/nms.pyx:77-170), used by
`CtdetDetector.merge_outputs` for multi-scale testing and `--nms` (lib/detectors/ctdet.py:59-74).

The reference runs this on the CPU as well (compiled Cython); it is not part of the accelerated path.  The restatement
keeps the reference's in-place semantics and its single-precision arithmetic (`cdef float` temporaries; `np.exp` evaluated
in double and rounded to float), so results are bit-identical to the compiled reference (tests/golden/soft_nms_kat.npz):

  * `boxes` [N, 5] float32 (x1, y1, x2, y2, score) is modified IN PLACE: selection-sorts by score, decays the scores of
    overlapping boxes, and swaps boxes whose score falls below `threshold` to the tail;
  * the outer loop runs over the ORIGINAL N (Cython evaluates `range(N)` once) while the inner loops use the shrinking N;
  * returns `list(range(N_final))`, which `merge_outputs` ignores -- it keeps every row of the modified array.
"""
import numpy as np

F = np.float32
D = np.float64


def soft_nms(boxes, sigma=0.5, Nt=0.3, threshold=0.001, method=0):
    assert boxes.ndim == 2 and boxes.dtype == np.float32 and boxes.shape[1] >= 5
    n0 = N = boxes.shape[0]
    sigma, Nt, threshold, one = F(sigma), F(Nt), F(threshold), F(1)
    for i in range(n0):
        maxscore, maxpos = boxes[i, 4], i
        tx1, ty1, tx2, ty2, ts = (F(v) for v in boxes[i, :5])
        pos = i + 1
        while pos < N:                                   # get max box
            if maxscore < boxes[pos, 4]:
                maxscore, maxpos = boxes[pos, 4], pos
            pos += 1
        boxes[i, :5] = boxes[maxpos, :5]                 # add max box as a detection
        boxes[maxpos, :5] = (tx1, ty1, tx2, ty2, ts)     # swap ith box with position of max box
        tx1, ty1, tx2, ty2, ts = (F(v) for v in boxes[i, :5])
        pos = i + 1
        while pos < N:                                   # N shrinks when a box falls below the threshold
            x1, y1, x2, y2 = (F(v) for v in boxes[pos, :4])
            # Differences of two floats are single precision; the `+ 1` is a double addition (Cython types the literal as
            # 1.0), so the products / sums that follow run in double and are rounded once on assignment to the
            # `cdef float` (checked against the compiled reference, Cython 3.x: tests/golden/soft_nms_kat.npz).
            area = F((D(F(x2 - x1)) + 1.0) * (D(F(y2 - y1)) + 1.0))
            iw = F(D(F(min(tx2, x2) - max(tx1, x1))) + 1.0)
            if iw > 0:
                ih = F(D(F(min(ty2, y2) - max(ty1, y1))) + 1.0)
                if ih > 0:
                    ua = F((D(F(tx2 - tx1)) + 1.0) * (D(F(ty2 - ty1)) + 1.0) + D(area) - D(F(iw * ih)))
                    ov = F(F(iw * ih) / ua)              # iou between max box and detection box
                    if method == 1:                      # linear
                        weight = F(1.0 - D(ov)) if ov > Nt else one
                    elif method == 2:                    # gaussian
                        weight = F(np.exp(D(F(F(-F(ov * ov)) / sigma))))
                    else:                                # original NMS
                        weight = F(0) if ov > Nt else one
                    boxes[pos, 4] = F(weight * boxes[pos, 4])
                    if boxes[pos, 4] < threshold:        # discard the box by swapping with the last box
                        boxes[pos, :5] = boxes[N - 1, :5]
                        N -= 1
                        pos -= 1
            pos += 1
    return list(range(N))
