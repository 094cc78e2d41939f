"""Error behaviour of the C ABI: invalid arguments come back as error codes with a message (ValueError at the Python
layer), never as a crash, and the context stays usable. Mirrors the reference's habit of raising ValueError for
non-positive sizes (of_library.py:56-57, 240-243)."""
import ctypes as C

import numpy as np
import pytest

import synth

pytestmark = pytest.mark.gpu


def test_invalid_arguments_are_reported_and_context_survives(ctx):
    import ofb200
    L, lib, P = ofb200._lib, ctx.lib, ofb200._lib.ptr
    a, b, mo = synth.make_pair(120, 160, 1, 1, max_disp=3.0)
    cfg = ofb200.make_pair_cfg(160, 120, 32, 0.01, 6, 7, (15, 15), 2, (3, 20, 0.03), variant="node",
                               principal=(mo["cx"], mo["cy"]), pos_scale=1.0 / mo["f"], flow_scale=1.0 / (mo["f"] * mo["dt"]))
    imu = np.zeros(1, L.IMU_DTYPE); imu["d"], imu["n"], imu["w"] = mo["d"], mo["n"], mo["w"]
    res = np.zeros(1, L.RESULT_DTYPE)

    def pairs(cfg_, n=1, prev=a, nxt=b, pitch=160, imu_=imu, pts=None, nin=None):
        return lib.ofb_frame_pairs(ctx.h, C.byref(cfg_), n, P(prev), P(nxt), pitch, 160 * 120, P(imu_), P(pts), P(nin), P(res),
                                   None, None, None)

    def bad_cfg(**kw):
        c = ofb200.make_pair_cfg(160, 120, 32, 0.01, 6, 7, (15, 15), 2, (3, 20, 0.03), variant="node")
        for k, v in kw.items():
            setattr(c, k, v)
        return c

    cases = [lambda: pairs(cfg, n=0), lambda: pairs(cfg, pitch=100), lambda: pairs(bad_cfg(max_corners=0)),
             lambda: pairs(bad_cfg(variant=7)), lambda: pairs(bad_cfg(max_level=-1)), lambda: pairs(bad_cfg(detect=0)),
             lambda: pairs(cfg, prev=None), lambda: pairs(bad_cfg(width=0)),
             lambda: lib.ofb_good_features(ctx.h, P(a), 160, 120, 100, None, 0, 10, C.c_double(0.01), C.c_double(5.0), 7,
                                           P(np.zeros((10, 2), np.float32)), 10, C.byref(C.c_int())),
             lambda: lib.ofb_solve_velocity(ctx.h, 9, P(np.zeros((4, 2))), P(np.zeros((4, 2))), C.c_int(4), C.c_double(1.0),
                                            P(np.array([0.0, 0, 1])), P(np.zeros(3)), None, P(np.zeros(3)),
                                            C.byref(C.c_double()), C.byref(C.c_int()), P(np.zeros(3)))]
    for i, f in enumerate(cases):
        rc = f()
        assert rc != 0, i
        with pytest.raises((ValueError, ofb200.OfbError, MemoryError)):
            L.check(rc)
        assert len(lib.ofb_last_error()) > 0
        # the context still works after every rejected call
        assert pairs(cfg) == 0
        ctx.sync()
        assert res["n_tracked"][0] > 5
    # Python layer: shape mismatches and unknown variants are ValueErrors
    with pytest.raises(ValueError):
        ofb200.frame_pairs(a[None], b[None, :100], imu, cfg, ctx=ctx)
    with pytest.raises(ValueError):
        ofb200.frame_pairs(a[None], b[None], imu[:0], cfg, ctx=ctx)
    with pytest.raises((ValueError, KeyError)):
        ofb200.make_pair_cfg(160, 120, 32, variant="nope")
    with pytest.raises(ValueError):
        ofb200.goodFeaturesToTrack(a, 10, 0.01, 5, useHarrisDetector=True, ctx=ctx)
