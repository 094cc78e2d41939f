"""Seeded random sweeps over parameters the fixed cases do not enumerate: image sizes, block sizes, quality levels,
min-distances, corner limits, masks for the detector; window sizes, pyramid depths, termination criteria, points near
and beyond the border for the tracker. Same contracts as tests/test_gpu_vision.py."""
import numpy as np
import pytest

from oracle import image_oracle as io
import synth
from test_gpu_vision import as_list, check_features

pytestmark = pytest.mark.gpu


def test_random_detector_cases(ctx):
    import ofb200
    rng = np.random.default_rng(20260101)
    for case in range(28):
        h, w = int(rng.integers(40, 420)), int(rng.integers(40, 700))
        img = synth.texture(h, w, 300 + case)
        kind = case % 4
        if kind == 1:                                     # low-contrast half: many values near the threshold
            img[:, : w // 2] = (img[:, : w // 2] // 8 + 100).astype(np.uint8)
        elif kind == 2:                                   # flat regions and a saturated block: plateaus
            img[: h // 3] = 30; img[h // 2: h // 2 + 7, w // 3: w // 3 + 40] = 255
        bs = int(rng.choice([3, 5, 7, 8, 12, 15, 32]))
        if min(h, w) < bs + 4 and case % 2:
            bs = 3
        q = float(rng.choice([0.001, 0.01, 0.05, 0.3, 0.7]))
        md = float(rng.choice([0.0, 0.5, 1.0, 2.5, 7.0, 10.0, 23.0]))
        mc = int(rng.choice([0, 1, 7, 100, 300, 700, 3000]))
        mask = (rng.random((h, w)) > 0.4).astype(np.uint8) if case % 5 == 0 else None
        check_features(ofb200, ctx, img, mc, q, md, bs, mask=mask)


def test_random_tracker_cases(ctx):
    import ofb200
    rng = np.random.default_rng(77)
    worst = 0.0
    for case in range(14):
        h, w = int(rng.integers(60, 300)), int(rng.integers(80, 420))
        a, b = synth.affine_pair(h, w, 400 + case, shift=(float(rng.uniform(-6, 6)), float(rng.uniform(-6, 6))),
                                 rot=float(rng.uniform(-0.02, 0.02)), scale=float(rng.uniform(0.99, 1.01)))
        n = int(rng.integers(5, 120))
        pts = np.stack([rng.uniform(-3, w + 3, n), rng.uniform(-3, h + 3, n)], 1).astype(np.float32)
        pts[: n // 3] = np.round(pts[: n // 3])            # integer positions (zero fractional weights)
        win = [(15, 15), (9, 13), (16, 15), (21, 21), (5, 7), (31, 11)][case % 6]
        ml = int(rng.integers(0, 5))
        crit = [(3, 20, 0.03), (3, 10, 0.5), (3, 5, 0.001), (3, 30, 0.01)][case % 4]
        n1, s1, e1 = ofb200.calcOpticalFlowPyrLK(a, b, pts.reshape(-1, 1, 2), None, winSize=win, maxLevel=ml, criteria=crit,
                                                 ctx=ctx)
        on, os_, oe = io.pyrlk(a, b, pts.reshape(-1, 1, 2), win, ml, crit)
        assert np.array_equal(s1, os_), (case, win, ml, crit)
        ok = s1.ravel() == 1
        if ok.any():
            d = float(np.abs(n1 - on)[ok].max())
            worst = max(worst, d)
            assert d <= 5e-3, (case, win, ml, crit, d)
            assert np.abs(e1 - oe)[ok].max() <= 5e-3
    assert worst <= 5e-3
