#!/usr/bin/env python
"""Short, fixed workload for ncu: a few frame-pair steps (BASELINE config 2 geometry) and one
Monte-Carlo sweep. Prints nothing that is used as a bench value.

    python tools/profile_pairs.py [--batch 8] [--steps 3] [--mc-trials 2000000]
"""
import argparse
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--mc-trials", type=int, default=2_000_000)
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--features", type=int, default=1000)
    ap.add_argument("--max-level", type=int, default=4)
    args = ap.parse_args()
    import torch
    import ofb200
    import synth
    ctx = ofb200.Context(0)
    W, H, B = args.width, args.height, args.batch
    pairs = [synth.make_pair(H, W, 0, i) for i in range(min(B, 4))]
    mo0 = pairs[0][2]
    cfg = ofb200.make_pair_cfg(W, H, args.features, 0.01, 10.0, 7, (15, 15), args.max_level, (3, 20, 0.03), variant="node",
                               principal=(mo0["cx"], mo0["cy"]), pos_scale=1.0 / mo0["f"],
                               flow_scale=1.0 / (mo0["f"] * mo0["dt"]))
    imu = np.zeros(B, ofb200._lib.IMU_DTYPE)
    a = np.stack([pairs[i % len(pairs)][0] for i in range(B)])
    b = np.stack([pairs[i % len(pairs)][1] for i in range(B)])
    for i in range(B):
        mo = pairs[i % len(pairs)][2]
        imu["d"][i], imu["n"][i], imu["w"][i] = mo["d"], mo["n"], mo["w"]
    da, db = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
    dimu = torch.from_numpy(imu.view(np.uint8).reshape(-1).copy()).cuda()
    dres = torch.zeros(B * ofb200._lib.RESULT_DTYPE.itemsize, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    P = ofb200._lib.ptr
    for _ in range(args.steps):
        ofb200._lib.check(ctx.lib.ofb_frame_pairs(ctx.h, C.byref(cfg), B, P(da), P(db), W, W * H, P(dimu), None, None, P(dres),
                                                  None, None, None))
    ctx.sync()
    res = np.zeros(B, ofb200._lib.RESULT_DTYPE)
    ctx.memcpy(res, dres, res.nbytes)
    print("pairs ok: tracked", res["n_tracked"].tolist()[:4], "v0", np.round(res["v"][0], 3), "truth", np.round(mo0["v"], 3))
    sim = ofb200.simulation
    pts = np.load(os.path.join(ROOT, "tests", "golden", "points.npy"))[:50]
    steps, pos, flow = sim.build_sweep("flow_errors", pts)
    for prec in ("fp32", "fp64"):
        mean, std, mR, n = sim.run_sweep(steps, pos, flow, max(args.mc_trials // len(steps), 1), seed=1, precision=prec, ctx=ctx)
    print("mc ok: step 50 mean", np.round(mean[50], 3), "std", np.round(std[50], 3))
    ctx.close()


if __name__ == "__main__":
    main()
