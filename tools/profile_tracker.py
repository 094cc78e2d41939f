#!/usr/bin/env python
"""Short, fixed workload for ncu: a fleet of camera streams through the device-resident feature lifecycle
(ofb_tracker_step): one detecting step, steady-state tracking steps and one step that forces a masked top-up.
Prints nothing that is used as a bench value.

    python tools/profile_tracker.py [--streams 64] [--steps 4]
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--streams", type=int, default=64)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--width", type=int, default=1280)
    ap.add_argument("--height", type=int, default=720)
    ap.add_argument("--features", type=int, default=500)
    args = ap.parse_args()
    import torch
    import ofb200
    import synth
    ctx = ofb200.Context(0)
    W, H, S = args.width, args.height, args.streams
    pairs = [synth.make_pair(H, W, 0, i) for i in range(4)]
    mo0 = pairs[0][2]
    a = torch.from_numpy(np.stack([pairs[i % 4][0] for i in range(S)])).cuda()
    b = torch.from_numpy(np.stack([pairs[i % 4][1] for i in range(S)])).cuda()
    imu = np.zeros(S, ofb200._lib.IMU_DTYPE)
    for i in range(S):
        mo = pairs[i % 4][2]
        imu["d"][i], imu["n"][i], imu["w"][i] = mo["d"], mo["n"], mo["w"]
    trk = ofb200.StreamTracker(W, H, max_features=args.features, min_features=args.features // 2, n_streams=S, topup="node",
                               mask_radius=30, variant="node", principal=(mo0["cx"], mo0["cy"]), scaling=1.0 / mo0["f"],
                               flow_scaling=1.0 / (mo0["f"] * mo0["dt"]), ctx=ctx)
    for k in range(args.steps):
        res = trk.step(b if k & 1 else a, imu)
        print("step", k, "tracked", int(res["n_tracked"].min()), "added", int(res["n_added"].max()), "points", int(res["n_points"].min()))
    # force a masked top-up in half of the streams: keep only 100 points there
    res, pts = trk.step(a if args.steps & 1 else b, imu, want_points=True)
    trk.set_points([p[:100] if s % 2 == 0 else p for s, p in enumerate(pts)])
    res = trk.step(b if args.steps & 1 else a, imu)
    print("top-up step: added", res["n_added"].tolist()[:4], "points", res["n_points"].tolist()[:4])
    trk.close()


if __name__ == "__main__":
    main()
