"""GPU parity at BASELINE.json's full sizes, through size-independent properties and direct oracle/cv2 checks
that stay within seconds: C2 (1080p/1000/maxLevel 4), C4 (3840x2160/5000/maxLevel 5), C5 (720p fleet), C3 (1e8
Monte-Carlo trials: sharding invariance and sampling consistency)."""
import numpy as np
import pytest

from oracle import image_oracle as io
from oracle import velocity_oracle as vo
import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def frame4k():
    return synth.make_pair(2160, 3840, 0, 7, max_disp=20.0)


def test_c4_pyramid_and_features_at_4k(ctx, frame4k):
    import ofb200
    a, b, mo = frame4k
    lv = ofb200.buildPyramid(a, 5, ctx=ctx)
    ref = io.build_pyramid(a, 5)
    assert [x.shape for x in lv] == [(2160, 3840), (1080, 1920), (540, 960), (270, 480), (135, 240), (68, 120)]
    for x, y in zip(lv, ref):
        assert np.array_equal(x, y)
    pts = ofb200.goodFeaturesToTrack(a, 5000, 0.01, 10, blockSize=7, ctx=ctx)
    assert len(pts) == 5000
    eig = ofb200.cornerMinEigenVal(a, 7, ctx=ctx)
    expect = io.select_features(eig, 5000, 0.01, 10)
    assert np.array_equal(pts, expect)
    # min-distance property of the accepted set and descending quality
    p = pts.reshape(-1, 2)
    q = eig[p[:, 1].astype(int), p[:, 0].astype(int)]
    assert np.all(np.diff(q) <= 0)
    from scipy.spatial import cKDTree
    d, _ = cKDTree(p).query(p, k=2)
    assert d[:, 1].min() >= 10.0


def test_c4_lk_and_velocity_at_4k(ctx, frame4k):
    import ofb200
    cv2 = pytest.importorskip("cv2")
    a, b, mo = frame4k
    pts = ofb200.goodFeaturesToTrack(a, 5000, 0.01, 10, blockSize=7, ctx=ctx)
    kw = dict(winSize=(15, 15), maxLevel=5, criteria=(3, 20, 0.03))
    n, s, e = ofb200.calcOpticalFlowPyrLK(a, b, pts, None, ctx=ctx, **kw)
    rn, rs, re_ = cv2.calcOpticalFlowPyrLK(a, b, pts, None, **kw)
    assert np.array_equal(s, rs)
    ok = rs.ravel() == 1
    assert ok.sum() > 4800
    assert np.abs(n - rn)[ok].max() <= 0.05
    cfg = ofb200.make_pair_cfg(3840, 2160, 5000, 0.01, 10, 7, (15, 15), 5, (3, 20, 0.03), variant="node",
                               principal=(mo["cx"], mo["cy"]), pos_scale=1.0 / mo["f"], flow_scale=1.0 / (mo["f"] * mo["dt"]))
    imu = np.zeros(1, ofb200._lib.IMU_DTYPE)
    imu["d"], imu["n"], imu["w"] = mo["d"], mo["n"], mo["w"]
    res, pp, pn, st = ofb200.frame_pairs(a[None], b[None], imu, cfg, want_tracks=True, ctx=ctx)
    assert np.array_equal(pp[0], pts.reshape(-1, 2)) and np.array_equal(st[0], s.ravel()) and np.array_equal(pn[0], n.reshape(-1, 2))
    okg = st[0] == 1
    x = (pn[0][okg].astype(np.float64) - np.array([mo["cx"], mo["cy"]])) / mo["f"]
    u = (pn[0][okg] - pp[0][okg]).astype(np.float64) / (mo["f"] * mo["dt"])
    vref = vo.solve_lgs(x, u, mo["d"], mo["n"], mo["w"], variant="node")[0]
    assert np.abs(res["v"][0] - vref).max() <= 1e-4 * np.abs(vref).max()
    assert np.abs(res["v"][0] - mo["v"]).max() <= 0.03 * max(1.0, np.abs(mo["v"]).max())


def test_c2_and_c5_batches_equal_single_pairs(ctx):
    """A batch is processed exactly like its pairs one by one (resident and host paths)."""
    import ofb200
    for (h, w, K, ml, nb) in [(1080, 1920, 1000, 4, 3), (720, 1280, 500, 3, 11)]:
        frames = [synth.make_pair(h, w, s % 3, 50 + s) for s in range(3)]
        frames = [frames[i % 3] for i in range(nb)]
        a = np.stack([f[0] for f in frames]); b = np.stack([f[1] for f in frames])
        mo0 = frames[0][2]
        cfg = ofb200.make_pair_cfg(w, h, K, 0.01, 10, 7, (15, 15), ml, (3, 20, 0.03), variant="node",
                                   principal=(mo0["cx"], mo0["cy"]), pos_scale=1.0 / mo0["f"], flow_scale=1.0 / (mo0["f"] * mo0["dt"]))
        imu = np.zeros(nb, ofb200._lib.IMU_DTYPE)
        for i, f in enumerate(frames):
            imu["d"][i], imu["n"][i], imu["w"][i] = f[2]["d"], f[2]["n"], f[2]["w"]
        res, pp, pn, st = ofb200.frame_pairs(a, b, imu, cfg, want_tracks=True, ctx=ctx)
        for i in range(3):
            r1, p1, n1, s1 = ofb200.frame_pairs(a[i:i + 1], b[i:i + 1], imu[i:i + 1], cfg, want_tracks=True, ctx=ctx)
            assert np.array_equal(r1["v"][0], res["v"][i]) and np.array_equal(n1[0], pn[i]) and np.array_equal(s1[0], st[i])
            # identical pairs inside the batch give identical results
            for j in range(i, nb, 3):
                assert np.array_equal(res["v"][j], res["v"][i]) and np.array_equal(pn[j], pn[i])
            assert np.abs(res["v"][i] - frames[i][2]["v"]).max() <= 0.05 * max(1.0, np.abs(frames[i][2]["v"]).max())
        assert (res["n_tracked"] >= 0.95 * K).all()


def test_c3_full_size_monte_carlo_properties(ctx, points200):
    """1e8 trials x 50 points in total (100 steps x 1e6): counts exact, two disjoint halves agree with the whole to
    fp64 rounding, and the per-step statistics are consistent with a 1e5-trial run at sampling tolerance."""
    import ofb200
    sim = ofb200.simulation
    steps, pos, flow = sim.build_sweep("flow_errors", points200[:50])
    assert len(steps) == 100
    per_step = 1_000_000
    whole = sim.run_steps(steps, pos, flow, per_step, seed=11, ctx=ctx)
    assert np.all(whole["n"] == per_step) and whole["n"].sum() == 1e8
    h1 = sim.run_steps(steps, pos, flow, per_step // 2, seed=11, trial_begin=0, ctx=ctx)
    h2 = sim.run_steps(steps, pos, flow, per_step // 2, seed=11, trial_begin=per_step // 2, ctx=ctx)
    acc = h1.view(np.float64).reshape(100, 8) + h2.view(np.float64).reshape(100, 8)
    np.testing.assert_allclose(acc, whole.view(np.float64).reshape(100, 8), rtol=1e-9, atol=1e-7)
    mean, std, mR, n = sim.stats_from_sums(whole, steps)
    small = sim.run_steps(steps, pos, flow, 100_000, seed=12, ctx=ctx)
    m2, s2, _, _ = sim.stats_from_sums(small, steps)
    assert np.all(np.abs(m2 - mean) <= 5 * std / np.sqrt(1e5) + 1e-9)
    ok = std > 1e-6
    assert np.all(np.abs(s2[ok] / std[ok] - 1) <= 5 / np.sqrt(2e5))
    # no noise at step 0 (flow_sig = position_sig = 0): the only scatter left is gyro/height/lever-arm noise
    assert std[0].max() < 0.03 and std[99].max() > std[0].max()


def test_c4_lifecycle_at_4k(ctx, frame4k):
    """Feature lifecycle (ofb_tracker_step) at the C4 shape: 3840x2160, 5000 features, maxLevel 5. Step 1 detects
    (cluster-mode selection), step 2 tracks and -- with min_features just below the maximum -- runs a masked top-up.
    Tracks equal the pair path bit for bit; the appended corners equal OpenCV's selection rule on the GPU's own
    lambda_min map under the cv2.circle exclusion mask of the surviving points."""
    import ofb200
    from oracle import tracker_oracle
    a, b, mo = frame4k
    K, R = 5000, 12
    trk = ofb200.StreamTracker(3840, 2160, max_features=K, min_features=K - 1, topup="node", mask_radius=R, variant="node",
                               principal=(mo["cx"], mo["cy"]), scaling=1.0 / mo["f"], flow_scaling=1.0 / (mo["f"] * mo["dt"]),
                               lk_params=dict(winSize=(15, 15), maxLevel=5, criteria=(3, 20, 0.03)), ctx=ctx)
    imu = np.zeros(1, ofb200._lib.IMU_DTYPE)
    imu["d"], imu["n"], imu["w"] = mo["d"], mo["n"], mo["w"]
    try:
        r0, p0 = trk.step(a, imu, want_points=True)
        r1, p1, kp, kn = trk.step(b, imu, want_points=True, want_kept=True)
    finally:
        trk.close()
    pts = ofb200.goodFeaturesToTrack(a, K, 0.01, 10, blockSize=7, ctx=ctx)
    assert int(r0["n_added"][0]) == K and np.array_equal(p0[0], pts)
    n, s, e = ofb200.calcOpticalFlowPyrLK(a, b, pts, None, winSize=(15, 15), maxLevel=5, criteria=(3, 20, 0.03), ctx=ctx)
    ok = s.ravel() == 1
    assert int(r1["n_tracked"][0]) == int(ok.sum()) == int(r1["n_kept"][0]) and ok.sum() > 4800
    assert np.array_equal(kp[0], pts.reshape(-1, 2)[ok]) and np.array_equal(kn[0], n.reshape(-1, 2)[ok])
    cfg = ofb200.make_pair_cfg(3840, 2160, K, 0.01, 10, 7, (15, 15), 5, (3, 20, 0.03), variant="node",
                               principal=(mo["cx"], mo["cy"]), pos_scale=1.0 / mo["f"], flow_scale=1.0 / (mo["f"] * mo["dt"]))
    res = ofb200.frame_pairs(a[None], b[None], imu, cfg, ctx=ctx)
    assert np.abs(r1["v"][0] - res["v"][0]).max() <= 1e-12 * max(1.0, np.abs(res["v"][0]).max())
    # masked top-up on frame 2
    na = int(r1["n_added"][0])
    assert na == K - int(ok.sum()) or na < K - int(ok.sum())          # fewer when the unmasked area runs out of corners
    mask = tracker_oracle.exclusion_mask(kn[0], R, 3840, 2160)
    assert np.array_equal(mask, ofb200.exclusion_mask(kn[0], R, 3840, 2160, ctx=ctx))
    eig = ofb200.cornerMinEigenVal(b, 7, ctx=ctx)
    expect = io.select_features(eig, K - int(ok.sum()), 0.01, 10, mask)
    expect = np.zeros((0, 2), np.float32) if expect is None else expect.reshape(-1, 2)
    assert na == len(expect) and np.array_equal(p1[0].reshape(-1, 2)[int(ok.sum()):], expect)
