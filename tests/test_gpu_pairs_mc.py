"""GPU parity: the fused frame-pair path (stages 1-4 in one call) and the Monte-Carlo stage, through
the C ABI, against the oracle, the reference's golden outputs and sampling-tolerance statistics."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import image_oracle as io
from oracle import velocity_oracle as vo
import synth

pytestmark = pytest.mark.gpu


def oracle_pair(a, b, mo, K, maxlevel, variant="node", t=None):
    pts = io.good_features(a, K, 0.01, 10, block_size=7)
    nxt, st, err = io.pyrlk(a, b, pts, (15, 15), maxlevel, (3, 20, 0.03))
    ok = st.ravel() == 1
    p0 = pts.reshape(-1, 2)[ok]; p1 = nxt.reshape(-1, 2)[ok]
    x = (p1.astype(np.float64) - np.array([mo["cx"], mo["cy"]])) / mo["f"]
    u = (p1 - p0).astype(np.float64) / (mo["f"] * mo["dt"])
    v = vo.solve_lgs(x, u, mo["d"], mo["n"], mo["w"], t, variant=variant)
    return pts, nxt, st, v


def run_pairs(ofb200, ctx, frames, K, maxlevel, variant="node", t=(0, 0, 0), device_resident=False):
    n = len(frames)
    a = np.stack([f[0] for f in frames]); b = np.stack([f[1] for f in frames])
    h, w = a.shape[1:]
    mo0 = frames[0][2]
    cfg = ofb200.make_pair_cfg(w, h, K, 0.01, 10, 7, (15, 15), maxlevel, (3, 20, 0.03), variant=variant,
                               principal=(mo0["cx"], mo0["cy"]), pos_scale=1.0 / mo0["f"],
                               flow_scale=1.0 / (mo0["f"] * mo0["dt"]))
    imu = np.zeros(n, ofb200._lib.IMU_DTYPE)
    for i, f in enumerate(frames):
        imu["d"][i], imu["n"][i], imu["w"][i], imu["t"][i] = f[2]["d"], f[2]["n"], f[2]["w"], t
    if device_resident:
        import torch
        ta = torch.from_numpy(a).cuda(); tb = torch.from_numpy(b).cuda()
        torch.cuda.synchronize()
        out = ofb200.frame_pairs(ta, tb, imu, cfg, want_tracks=True, ctx=ctx)
        return out
    return ofb200.frame_pairs(a, b, imu, cfg, want_tracks=True, ctx=ctx)


@pytest.mark.parametrize("shape,K,ml,resident", [((240, 320), 100, 3, False), ((480, 640), 200, 3, False),
                                                 ((480, 640), 200, 3, True)])
def test_frame_pairs_vs_oracle_pipeline(ctx, shape, K, ml, resident):
    import ofb200
    frames = [synth.make_pair(shape[0], shape[1], s, s, max_disp=8.0) for s in range(3)]
    res, pp, pn, st = run_pairs(ofb200, ctx, frames, K, ml, device_resident=resident)
    for i, (a, b, mo) in enumerate(frames):
        pts, nxt, ost, (v, _, rank, s) = oracle_pair(a, b, mo, K, ml)
        n = int(res["n_features"][i])
        # features: selection rule exact on the GPU's own map (ties documented in test_gpu_vision)
        eig = ofb200.cornerMinEigenVal(a, 7, ctx=ctx)
        expect = io.select_features(eig, K, 0.01, 10).reshape(-1, 2)
        assert n == len(expect) and np.array_equal(pp[i, :n], expect)
        if np.array_equal(expect, pts.reshape(-1, 2)):
            assert np.array_equal(st[i, :n], ost.ravel())
            ok = ost.ravel() == 1
            assert np.abs(pn[i, :n] - nxt.reshape(-1, 2))[ok].max() <= 0.05
            assert int(res["n_tracked"][i]) == int(ok.sum())
            rel = np.abs(res["v"][i] - v).max() / np.abs(v).max()
            # LK differs from the oracle by <=5e-3 px, which moves v by O(1e-3) relative; the SOLVE itself
            # is checked to 1e-4 below on identical tracks
            assert rel <= 2e-2, rel
        # stage 4 on IDENTICAL tracks: feed the GPU's own tracks to the fp64 reference solve
        ok = st[i, :n] == 1
        x = (pn[i, :n][ok].astype(np.float64) - np.array([mo["cx"], mo["cy"]])) / mo["f"]
        u = (pn[i, :n][ok] - pp[i, :n][ok]).astype(np.float64) / (mo["f"] * mo["dt"])
        vref, _, rk, sref = vo.solve_lgs(x, u, mo["d"], mo["n"], mo["w"], variant="node")
        assert np.abs(res["v"][i] - vref).max() <= 1e-4 * np.abs(vref).max()
        np.testing.assert_allclose(res["s"][i], sref, rtol=1e-8)
        assert res["rank"][i] == 3
        # end-to-end sanity (not a parity bar): recovered vs generated velocity
        assert np.abs(res["v"][i] - mo["v"]).max() <= 0.05 * max(1.0, np.abs(mo["v"]).max())


def test_frame_pairs_track_only_and_variants(ctx):
    import ofb200
    a, b, mo = synth.make_pair(240, 320, 4, 4, max_disp=6.0)
    K = 64
    pts = io.good_features(a, K, 0.01, 10, block_size=7).reshape(-1, 2)
    cfg = ofb200.make_pair_cfg(320, 240, K, variant="exp", principal=(mo["cx"], mo["cy"]), pos_scale=1.0 / mo["f"],
                               flow_scale=1.0 / (mo["f"] * mo["dt"]), detect=False)
    imu = np.zeros(1, ofb200._lib.IMU_DTYPE)
    t = np.array([0.02, 0.0, 0.205])
    imu["d"], imu["n"], imu["w"], imu["t"] = mo["d"], mo["n"], mo["w"], t
    pin = np.zeros((1, K, 2), np.float32); pin[0, :len(pts)] = pts
    res, pp, pn, st = ofb200.frame_pairs(a[None], b[None], imu, cfg, pts_in=pin, n_in=[len(pts)], want_tracks=True, ctx=ctx)
    n = len(pts)
    assert int(res["n_features"][0]) == n
    ok = st[0, :n] == 1
    x = (pn[0, :n][ok].astype(np.float64) - np.array([mo["cx"], mo["cy"]])) / mo["f"]
    u = (pn[0, :n][ok] - pp[0, :n][ok]).astype(np.float64) / (mo["f"] * mo["dt"])
    vref = vo.solve_lgs(x, u, mo["d"], mo["n"], mo["w"], t, variant="exp")[0]
    assert np.abs(res["v"][0] - vref).max() <= 1e-4 * np.abs(vref).max()
    # the same pair five times as one batch: host frames, device-resident frames (twin-context chunks) and the
    # sequence layout (a, b, a, b, ...: even pairs are this pair) must all reproduce the single-pair result
    import torch
    B = 5
    imu5 = np.repeat(imu, B); pin5 = np.repeat(pin, B, axis=0); nin5 = [len(pts)] * B
    a5, b5 = np.stack([a] * B), np.stack([b] * B)
    r_h = ofb200.frame_pairs(a5, b5, imu5, cfg, pts_in=pin5, n_in=nin5, want_tracks=True, ctx=ctx)
    r_d = ofb200.frame_pairs(torch.from_numpy(a5).cuda(), torch.from_numpy(b5).cuda(), imu5, cfg, pts_in=pin5, n_in=nin5,
                             want_tracks=True, ctx=ctx)
    seq = torch.from_numpy(np.stack([a, b] * 3)).cuda()                  # 6 frames -> 5 pairs
    r_s = ofb200.frame_pairs(seq[:-1], seq[1:], imu5, cfg, pts_in=pin5, n_in=nin5, want_tracks=True, ctx=ctx)
    for r, idx in ((r_h, range(B)), (r_d, range(B)), (r_s, (0, 2, 4))):
        for i in idx:
            assert np.array_equal(r[0]["v"][i], res["v"][0]) and np.array_equal(r[2][i], pn[0]) and np.array_equal(r[3][i], st[0])


# ---- Monte-Carlo -----------------------------------------------------------------------------------
V, W, N3, T3 = np.ones(3), np.ones(3), np.array([0.0, 0, 1]), np.array([0.02, 0, 0.205])
SIG = dict(ang_vel_sig=0.00071, translation_sig=0.005, height_sig=0.01, flow_sig=0.056 * np.sqrt(2) * 1.23,
           position_sig=0.056 * 1.23, normal_sig=0.00065)


@pytest.fixture(scope="module")
def pos50(points200):
    import ofb200
    return ofb200.simulation.centred_points(points200)[:50]


@pytest.mark.parametrize("precision,tol", [("fp64", 1e-8), ("fp32", 5e-4)])
def test_mc_per_trial_parity_on_replayed_stream(ctx, pos50, precision, tol):
    """Every trial's v_obs and R equal the oracle's restatement of simulation.py:39-64 when the
    oracle replays the same Philox draws."""
    import ofb200
    sim = ofb200.simulation
    n = np.array([0.1, -0.05, 1.0])
    tf = vo.generate_test_data(pos50, V, W, 1.3, n, T3)
    step = sim.make_step(V, W, 1.3, n, T3, 50, 0, **SIG)
    seed, base, ntr, t0 = 2024, 5, 24, 1_000_000_007
    sums, vd, Rd = sim.run_steps([step], pos50, tf, ntr, seed=seed, step_id_base=base, trial_begin=t0,
                                 precision=precision, dump=True, ctx=ctx)
    z = vo.mc_normals(seed, base, t0 + np.arange(ntr), 50)
    for k in range(ntr):
        v_ref, R_ref = vo.of_trial(V, W, 1.3, n, T3, pos50, tf, SIG["ang_vel_sig"], SIG["translation_sig"],
                                   SIG["ang_vel_sig"] * z[k, 0, :3], SIG["translation_sig"] * z[k, 1, :3],
                                   SIG["height_sig"] * z[k, 0, 3], SIG["flow_sig"] * z[k, 3:, 0:2],
                                   SIG["position_sig"] * z[k, 3:, 2:4])
        assert np.abs(vd[0, k] - v_ref).max() <= tol * max(1.0, np.abs(v_ref).max()), (k, vd[0, k], v_ref)
        assert abs(Rd[0, k] - R_ref) <= max(tol, 1e-6) * R_ref
    # the sums are the sums of the dumped trials
    np.testing.assert_allclose(sums["n"][0], ntr)
    np.testing.assert_allclose(sums["sum_dv"][0], (vd[0] - V).sum(0), rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(sums["sum_dv2"][0], ((vd[0] - V) ** 2).sum(0), rtol=1e-9)
    np.testing.assert_allclose(sums["sum_R"][0], Rd[0].sum(), rtol=1e-9)


def test_mc_statistics_vs_reference_run(ctx, pos50):
    """SURVEY 8d tolerance: |mean_gpu - mean_ref| <= 4 sigma_ref/sqrt(n_ref), |sigma_gpu/sigma_ref - 1| <= 4/sqrt(2 n_ref)
    against the 2000-trial run of the reference's own of_simulation (golden), GPU with 2e5 trials."""
    import ofb200
    g = np.load(os.path.join(GOLDEN, "velocity_golden.npz"))
    sim = ofb200.simulation
    tf = g["ofsim_flow"]
    np.testing.assert_allclose(pos50, g["ofsim_pos"], atol=1e-15)
    step = sim.make_step(V, W, 1.0, N3, T3, 50, 0, **SIG)
    n_ref = 2000
    for precision in ("fp32", "fp64"):
        mean, std, mR, n = sim.run_sweep([step], pos50, tf, 200_000, seed=9, precision=precision, ctx=ctx)
        assert n[0] == 200_000
        assert np.all(np.abs(mean[0] - g["ofsim2000_mean"]) <= 4 * g["ofsim2000_std"] / np.sqrt(n_ref))
        assert np.all(np.abs(std[0] / g["ofsim2000_std"] - 1) <= 4 / np.sqrt(2 * n_ref))
        assert abs(mR[0] / float(g["ofsim2000_R"]) - 1) <= 0.05
    # errors-in-variables bias of v_z is reproduced, not "fixed" 
    assert mean[0, 2] < 0.99


def test_mc_sharding_is_exact_union(ctx, pos50):
    """Counter-based RNG: any split of the trial range gives the same sums up to fp64 summation order."""
    import ofb200
    sim = ofb200.simulation
    tf = vo.generate_test_data(pos50, V, W, 1.0, N3, T3)
    steps = [sim.make_step(V, W, 1.0, N3, T3, 50, 0, **dict(SIG, flow_sig=0.001 * i)) for i in range(5)]
    pos = np.tile(pos50, (1, 1)); total = 50_000
    whole = sim.run_steps(steps, pos, tf, total, seed=3, ctx=ctx)
    for world in (2, 3, 8):
        acc = np.zeros((5, 8))
        for r in range(world):
            b, c = sim.shard_range(total, r, world)
            part = sim.run_steps(steps, pos, tf, c, seed=3, trial_begin=b, ctx=ctx)
            acc += part.view(np.float64).reshape(5, 8)
        np.testing.assert_allclose(acc, whole.view(np.float64).reshape(5, 8), rtol=1e-10, atol=1e-9)
    # bit-reproducible for a fixed launch shape
    again = sim.run_steps(steps, pos, tf, total, seed=3, ctx=ctx)
    assert np.array_equal(again.view(np.float64), whole.view(np.float64))


@pytest.mark.parametrize("name,golden,trials_ref", [("flow_errors", "effect_of_flow_errors", 100),
                                                    ("distance_error", "effect_of_distance_error", 100),
                                                    ("ang_vel_error", "effect_o_ang_vel_error", 100),
                                                    ("translation_error", "effect_of_translation_error", 100),
                                                    ("normal_error", "effect_of_normal_error", 10),
                                                    ("orientation", "effect_of_orientation", 10),
                                                    ("point_position", "effect_of_point_position", 100)])
def test_sweeps_consistent_with_reference_npy(ctx, points200, name, golden, trials_ref):
    """The seven saved sweeps the committed reference code reproduces: the GPU sweep (20000 trials/step)
    must explain the reference's 10-100-trial outputs at their own sampling noise."""
    import ofb200
    ref = np.load(os.path.join(GOLDEN, "sweep_%s.npy" % golden))
    flat, mR = ofb200.run_named_sweep(name, points200, trials=20000, seed=1, ctx=ctx)
    k = len(ref) // 6
    assert flat.shape == ref.shape
    gm, gs = flat[:3 * k].reshape(k, 3), flat[3 * k:].reshape(k, 3)
    rm, rs = ref[:3 * k].reshape(k, 3), ref[3 * k:].reshape(k, 3)
    # z-scores of the reference means under the GPU's (much better resolved) distribution
    sig = np.maximum(gs, 1e-12) / np.sqrt(trials_ref)
    z = (rm - gm) / sig
    finite = np.isfinite(z) & (gs > 1e-9) & (gs < 5)          # skip the singular step (orientation at 90 deg)
    if name in ("ang_vel_error", "translation_error"):
        # these two saved files carry a different errors-in-variables bias of v_z (0.98-0.99) than the
        # committed reference code produces (0.958, checked with the reference's own functions): like
        # effect_of_height.npy they were saved under an earlier parameterisation. x/y means and all three
        # sigmas do reproduce and are gated; the z mean is not ("golden sweeps").
        finite[:, 2] = False
    frac_bad = np.mean(np.abs(z[finite]) > 5)
    assert frac_bad <= 0.03, (name, frac_bad, np.abs(z[finite]).max())
    finite = np.isfinite(z) & (gs > 1e-9) & (gs < 5)
    ratio = rs[finite] / gs[finite]
    lo, hi = (0.2, 3.0) if trials_ref == 10 else (0.6, 1.5)
    assert np.mean((ratio < lo) | (ratio > hi)) <= 0.05, (name, ratio.min(), ratio.max())


def test_of_simulation_signature_and_feas(ctx, points200, pos50):
    import ofb200
    sim = ofb200.simulation
    g = np.load(os.path.join(GOLDEN, "velocity_golden.npz"))
    tf = g["ofsim_flow"]
    v_obs, feasible, R = ofb200.of_simulation(V, W, 1.0, N3, T3, pos50, SIG["ang_vel_sig"], SIG["translation_sig"],
                                              SIG["height_sig"], SIG["flow_sig"], SIG["position_sig"], SIG["normal_sig"],
                                              iterations=100, true_flow=tf, seed=5, step_id=0, ctx=ctx)
    assert v_obs.shape == (100, 3) and feasible.shape == (2, 50) and R.shape == (100,)
    # `feasible` is the last trial's: rebuild it with the oracle from the replayed stream
    z = vo.mc_normals(5, 0, np.array([99]), 50)[0]
    f_ref = vo.feasibility(pos50 + SIG["position_sig"] * z[3:, 2:4], V, tf + SIG["flow_sig"] * z[3:, 0:2],
                           W + SIG["ang_vel_sig"] * z[0, :3], T3 + SIG["translation_sig"] * z[1, :3], N3)
    np.testing.assert_allclose(feasible, f_ref, rtol=1e-9, atol=1e-12)
    with pytest.raises(ValueError):
        ofb200.of_simulation(V, W, 1.0, N3, T3, pos50, 0, 0, 0, 0, 0, 0, iterations=0, true_flow=tf, ctx=ctx)
    # module-global style, as the reference script drives it
    sim.iterations = 10; sim.true_flow = tf
    v_obs, _, _ = sim.of_simulation(V, W, 1.0, N3, T3, pos50, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0)
    np.testing.assert_allclose(v_obs, np.ones((10, 3)), atol=1e-5)       # no noise -> exact round trip (fp32 path)
    sim.true_flow = None
    # feas_simulation: per-trial parity of the six per-point quantities on the replayed stream
    pos = g["feas_pos"]; tf2 = g["feas_tf"]
    seed, sid, ntr = 77, 2, 8
    sums = sim.feas_simulation(V, W, 2.0, N3, T3, pos, SIG["ang_vel_sig"], SIG["translation_sig"], SIG["height_sig"],
                               SIG["flow_sig"], SIG["position_sig"], 0.0, V, iterations=ntr, true_flow=tf2, seed=seed,
                               step_id=sid, normal_sig=0.3, velocity_sig=0.01, ctx=ctx, return_sums=True)
    z = vo.mc_normals(seed, sid, np.arange(ntr), 200)
    acc = np.zeros((6, 200))
    for k in range(ntr):
        out = vo.feas_trial(W, 2.0, N3, T3, pos, tf2, V, SIG["ang_vel_sig"] * z[k, 0, :3], SIG["translation_sig"] * z[k, 1, :3],
                            SIG["height_sig"] * z[k, 0, 3], SIG["flow_sig"] * z[k, 3:, 0:2], SIG["position_sig"] * z[k, 3:, 2:4],
                            0.01 * z[k, 2, :3], 0.3 * 0.3 * z[k, 1, 3], 0.3 * 0.3 * z[k, 2, 3])
        acc += np.array(out)
    np.testing.assert_allclose(sums, acc, rtol=1e-7, atol=1e-9)
    # statistical agreement with the reference's own 300-trial feas_simulation (population means)
    means = sim.feas_simulation(V, W, 2.0, N3, T3, pos, SIG["ang_vel_sig"], SIG["translation_sig"], SIG["height_sig"],
                                SIG["flow_sig"], SIG["position_sig"], 0.0, V, iterations=20000, true_flow=tf2, seed=1,
                                step_id=9, ctx=ctx)
    for i in range(6):
        ref = g["feas_mean%d" % i]
        assert abs(np.mean(means[i]) - np.mean(ref)) <= 0.03 * max(abs(np.mean(ref)), 0.05), (i, np.mean(means[i]), np.mean(ref))


def test_overlap(ctx):
    import ofb200
    g = np.load(os.path.join(GOLDEN, "velocity_golden.npz"))
    assert ofb200.overlap(g["ov_a"], g["ov_b"], ctx=ctx) == int(g["ov_out"])
    rng = np.random.default_rng(0)
    for _ in range(3):
        a = rng.normal(size=5000); b = rng.normal(0.3, 2.0, size=3000)
        assert ofb200.overlap(a, b, ctx=ctx) == vo.overlap(a, b)
    a = np.array([1.0, 2.0, 2.0, 3.0]); b = np.array([3.0, 3.0, 1.0])
    assert ofb200.overlap(a, b, ctx=ctx) == vo.overlap(a, b)


@pytest.mark.parametrize("width", [160, 161])
def test_frame_pairs_pipelined_host_path_equals_resident(ctx, width):
    """Host frames in batches > 8 go through the copy/compute pipeline (device staging buffer, sub-batches on the
    context and its twin); results must be identical to the same pairs processed one small batch at a time and to
    the resident path. Width 161: the staging pitch (176) differs from the host pitch -> per-frame 2-D copies."""
    import ofb200
    import torch
    frames = [synth.make_pair(120, width, s % 5, s, max_disp=4.0) for s in range(19)]
    a = np.stack([f[0] for f in frames]); b = np.stack([f[1] for f in frames])
    mo0 = frames[0][2]
    K = 48
    cfg = ofb200.make_pair_cfg(width, 120, K, 0.01, 6, 5, (15, 15), 2, (3, 20, 0.03), variant="node",
                               principal=(mo0["cx"], mo0["cy"]), pos_scale=1.0 / mo0["f"], flow_scale=1.0 / (mo0["f"] * mo0["dt"]))
    imu = np.zeros(len(frames), ofb200._lib.IMU_DTYPE)
    for i, f in enumerate(frames):
        imu["d"][i], imu["n"][i], imu["w"][i] = f[2]["d"], f[2]["n"], f[2]["w"]
    for rep in range(2):          # twice: the second call reuses the workspace slots
        res, pp, pn, st = ofb200.frame_pairs(a, b, imu, cfg, want_tracks=True, ctx=ctx)
    ta, tb = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
    torch.cuda.synchronize()
    res_d, pp_d, pn_d, st_d = ofb200.frame_pairs(ta, tb, imu, cfg, want_tracks=True, ctx=ctx)
    for name in ("v", "s", "res", "rank", "n_features", "n_tracked"):
        assert np.array_equal(res[name], res_d[name]), name
    assert np.array_equal(pp, pp_d) and np.array_equal(pn, pn_d) and np.array_equal(st, st_d)
    for i0 in (0, 8, 16):
        sl = slice(i0, min(i0 + 5, len(frames)))
        r1, p1, n1, s1 = ofb200.frame_pairs(a[sl], b[sl], imu[sl], cfg, want_tracks=True, ctx=ctx)
        assert np.array_equal(r1["v"], res["v"][sl]) and np.array_equal(n1, pn[sl]) and np.array_equal(s1, st[sl])
    assert (res["n_tracked"] > 10).all()


def test_mc_without_error_bound_gives_identical_velocities(ctx, pos50):
    import ofb200
    sim = ofb200.simulation
    tf = vo.generate_test_data(pos50, V, W, 1.0, N3, T3)
    step = sim.make_step(V, W, 1.0, N3, T3, 50, 0, **SIG)
    for prec in ("fp32", "fp64"):
        s1, v1, R1 = sim.run_steps([step], pos50, tf, 300, seed=4, precision=prec, dump=True, ctx=ctx)
        s2, v2, R2 = sim.run_steps([step], pos50, tf, 300, seed=4, precision=prec, dump=True, want_R=False, ctx=ctx)
        # the two template instantiations schedule/contract the fp arithmetic differently: equal to rounding
        tol = 1e-5 if prec == "fp32" else 1e-11
        np.testing.assert_allclose(v1, v2, rtol=tol, atol=tol)
        assert np.all(R2 == 0) and np.all(R1 > 0)
        np.testing.assert_allclose(s1["sum_dv"], s2["sum_dv"], rtol=1e-3, atol=300 * tol)
        assert s2["sum_R"][0] == 0


def test_frame_sequence_equals_independent_pairs(ctx):
    """Sequence layout (next == prev + image_stride: consecutive frames of one stream in one buffer) uploads every
    frame and builds its pyramid once; the results must be bit-identical to the same pairs passed as two buffers,
    for host frames (pipelined sub-batches of 16 pairs + 1 frame) and for device-resident frames (twin chunks)."""
    import ofb200
    import torch
    a0, b0, mo = synth.make_pair(120, 160, 3, 7, max_disp=4.0)
    a1, b1, _ = synth.make_pair(120, 160, 3, 9, max_disp=3.0)
    # ping-pong chains: every consecutive pair is a small known motion
    frames = np.stack(([a0, b0] * 6 + [a1, b1] * 6)[:22] + [a1])       # 23 frames -> 22 pairs (12th pair jumps)
    n = len(frames) - 1
    K = 48
    cfg = ofb200.make_pair_cfg(160, 120, K, 0.01, 6, 5, (15, 15), 2, (3, 20, 0.03), variant="node",
                               principal=(mo["cx"], mo["cy"]), pos_scale=1.0 / mo["f"], flow_scale=1.0 / (mo["f"] * mo["dt"]))
    imu = np.zeros(n, ofb200._lib.IMU_DTYPE)
    imu["d"][:], imu["n"][:], imu["w"][:] = mo["d"], mo["n"], mo["w"]
    prev_c, next_c = frames[:-1].copy(), frames[1:].copy()           # separate buffers: independent-pair path
    ref = ofb200.frame_pairs(prev_c, next_c, imu, cfg, want_tracks=True, ctx=ctx)
    for rep in range(2):
        seq = ofb200.frame_sequence(frames, imu, cfg, want_tracks=True, ctx=ctx)
    tf = torch.from_numpy(frames).cuda()
    torch.cuda.synchronize()
    seq_d = ofb200.frame_sequence(tf, imu, cfg, want_tracks=True, ctx=ctx)
    short = ofb200.frame_sequence(frames[:4], imu[:3], cfg, want_tracks=True, ctx=ctx)     # single-pass path
    for got, sl in ((seq, slice(None)), (seq_d, slice(None)), (short, slice(0, 3))):
        for name in ("v", "s", "res", "rank", "n_features", "n_tracked"):
            assert np.array_equal(got[0][name], ref[0][name][sl]), name
        for k in (1, 2, 3):
            assert np.array_equal(got[k], ref[k][sl])
    assert (ref[0]["n_tracked"][:10] > 10).all()
