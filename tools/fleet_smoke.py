#!/usr/bin/env python
"""torchrun --nproc-per-node N tools/fleet_smoke.py : a small fleet through ofb200.FleetTracker over NCCL; every rank
checks that the gathered velocity table equals what one process computes for all streams."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist
    import ofb200
    import synth
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    S, w, h, T = 6, 320, 240, 3
    base = [synth.texture(h + 16, w + 16, 60 + s) for s in range(S)]
    frames = [[np.ascontiguousarray(base[s][2 * k:2 * k + h, (s % 3 + 1) * k:(s % 3 + 1) * k + w]) for k in range(T)] for s in range(S)]
    imu = np.zeros(S, ofb200._lib.IMU_DTYPE)
    imu["d"], imu["n"] = 1.0 + 0.1 * np.arange(S), [0.0, 0.0, 1.0]
    kw = dict(max_features=150, min_features=50, topup="node", variant="node", scaling=1.0 / 256.0)
    ctx = ofb200.Context(local)
    fleet = ofb200.FleetTracker(S, w, h, ctx=ctx, **kw)
    for k in range(T):
        fleet.step(np.stack(fleet.select([frames[s][k] for s in range(S)])), imu[fleet.streams])
    table = fleet.gather_velocities()
    fleet.close()
    ref = ofb200.StreamTracker(w, h, n_streams=S, ctx=ctx, **kw)
    for k in range(T):
        r = ref.step(np.stack([frames[s][k] for s in range(S)]), imu)
    ref.close()
    assert np.array_equal(table, r["v"]), (table, r["v"])
    print("rank %d of %d: fleet table equals the single-process result, v[0] = %s" % (dist.get_rank(), dist.get_world_size(), table[0]))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
