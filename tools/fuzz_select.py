#!/usr/bin/env python
"""One-off fuzz of the corner selection (run on the GPU box): random images / parameters as in tests/test_gpu_random.py, every
case under a forced cluster size and block size of the selection kernel (OFB_SELECT_CLUSTER / OFB_SELECT_THREADS), checked
with tests/test_gpu_vision.py::check_features (OpenCV's selection rule applied by the oracle to the GPU's own map).
    python tools/fuzz_select.py [n_cases] [seed]"""
import os
import sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import ofb200            # noqa: E402
import synth             # noqa: E402
from test_gpu_vision import check_features   # noqa: E402

n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 200
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
ctx = ofb200.Context(0)
ties = 0
for case in range(n_cases):
    h, w = int(rng.integers(40, 520)), int(rng.integers(40, 900))
    if case % 10 == 9:                                    # a large image now and then: many chunks per image
        h, w = int(rng.integers(700, 1300)), int(rng.integers(1000, 2200))
    img = synth.texture(h, w, 5000 + case)
    kind = case % 5
    if kind == 1:
        img[:, : w // 2] = (img[:, : w // 2] // 8 + 100).astype(np.uint8)
    elif kind == 2:
        img[: h // 3] = 30; img[h // 2: h // 2 + 7, w // 3: w // 3 + 40] = 255
    elif kind == 3:                                       # tiled: plateaus of exactly equal lambda_min
        t = synth.texture(24, 24, 7000 + case)
        img = np.ascontiguousarray(np.tile(t, (h // 24 + 1, w // 24 + 1))[:h, :w])
    bs = int(rng.choice([3, 5, 7, 12]))
    if min(h, w) < bs + 4:
        bs = 3
    q = float(rng.choice([0.0005, 0.001, 0.01, 0.05, 0.3]))
    md = float(rng.choice([0.0, 1.0, 2.0, 3.0, 5.0, 10.0, 23.0]))
    mc = int(rng.choice([0, 1, 50, 200, 400, 600, 1500, 5000]))
    os.environ["OFB_SELECT_CLUSTER"] = str(rng.choice([1, 2, 4, 8, 16]))
    os.environ["OFB_SELECT_THREADS"] = str(rng.choice([256, 512, 1024]))
    try:
        ties += check_features(ofb200, ctx, img, mc, q, md, bs) or 0
    except Exception as e:
        print("FAILED case", case, (h, w), dict(bs=bs, q=q, md=md, mc=mc), os.environ["OFB_SELECT_CLUSTER"], os.environ["OFB_SELECT_THREADS"])
        raise
print("fuzz_select: %d cases passed (tie groups explained: %d)" % (n_cases, ties))
