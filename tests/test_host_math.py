"""Runs the product's __host__ __device__ arithmetic (Philox, Box-Muller, one Monte-Carlo trial, the
3x3 eigen helpers -- csrc/math3.cuh, csrc/montecarlo.cu) on the CPU through tools/host_check.cu and
compares it with the oracle. No GPU needed; needs nvcc only to compile the harness."""
import os
import shutil
import struct
import subprocess

import numpy as np
import pytest

from conftest import ROOT, GOLDEN
from oracle import velocity_oracle as vo

NVCC = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    if not os.path.exists(NVCC):
        pytest.skip("nvcc not available")
    exe = str(tmp_path_factory.mktemp("hc") / "host_check")
    subprocess.check_call([NVCC, "-O1", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a",
                           "--expt-relaxed-constexpr", "-diag-suppress", "177", "-o", exe,
                           os.path.join(ROOT, "tools", "host_check.cu")])
    return exe


def test_trial_arithmetic_matches_oracle(harness, tmp_path):
    import ofb200
    from ofb200 import simulation as sim
    pts = np.load(os.path.join(GOLDEN, "points.npy"))
    pos = sim.centred_points(pts)[:50]
    v, w, n, t = np.ones(3), np.ones(3), np.array([0.1, -0.05, 1.0]), np.array([0.02, 0, 0.205])
    h = 1.3
    flow = vo.generate_test_data(pos, v, w, h, n, t)
    sig = dict(ang_vel_sig=0.00071, translation_sig=0.005, height_sig=0.01, flow_sig=0.0974, position_sig=0.0689, normal_sig=0.00065)
    step = sim.make_step(v, w, h, n, t, len(pos), 0, **sig)
    seed, step_id, ntr = 0x1234567890ABCDEF, 7, 16
    fin, fout = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    with open(fin, "wb") as f:
        f.write(bytes(step)); f.write(struct.pack("<QII", seed, step_id, ntr))
        f.write(np.ascontiguousarray(pos).tobytes()); f.write(np.ascontiguousarray(flow).tobytes())
    subprocess.check_call([harness, fin, fout])
    out = np.fromfile(fout, dtype=np.float64)
    r32 = out[:4 * ntr].reshape(ntr, 4); r64 = out[4 * ntr:8 * ntr].reshape(ntr, 4)
    zz = out[8 * ntr:12 * ntr].reshape(ntr, 4)
    rest = out[12 * ntr:]
    z = vo.mc_normals(seed, step_id, np.arange(ntr), len(pos))
    np.testing.assert_allclose(zz, z[:, 3, :], atol=1e-12)           # RNG contract, draw block 3
    for k in range(ntr):
        v_ref, R_ref = vo.of_trial(v, w, h, n, t, pos, flow, sig["ang_vel_sig"], sig["translation_sig"],
                                   sig["ang_vel_sig"] * z[k, 0, :3], sig["translation_sig"] * z[k, 1, :3],
                                   sig["height_sig"] * z[k, 0, 3], sig["flow_sig"] * z[k, 3:, 0:2],
                                   sig["position_sig"] * z[k, 3:, 2:4])
        np.testing.assert_allclose(r64[k, :3], v_ref, rtol=1e-9, atol=1e-11)
        np.testing.assert_allclose(r64[k, 3], R_ref, rtol=1e-9)
        np.testing.assert_allclose(r32[k, :3], v_ref, rtol=2e-4, atol=2e-4)     # fp32 per-point path
        np.testing.assert_allclose(r32[k, 3], R_ref, rtol=2e-4)
    M6, ev, q, lmin = rest[:6], rest[6:9], rest[9:18].reshape(3, 3), rest[18]
    M = np.array([[M6[0], M6[1], M6[2]], [M6[1], M6[3], M6[4]], [M6[2], M6[4], M6[5]]])
    w_ref = np.linalg.eigvalsh(M)[::-1]
    np.testing.assert_allclose(ev, w_ref, rtol=1e-12)
    np.testing.assert_allclose(lmin, w_ref[-1], rtol=1e-12)
    for i in range(3):
        np.testing.assert_allclose(M @ q[i], ev[i] * q[i], atol=1e-10)
