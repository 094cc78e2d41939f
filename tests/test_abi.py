"""The C-ABI library loads on a CPU-only box and exports exactly what include/ofb200.h declares;
the Python mirror fails loudly (no CPU fallback) when no device is present."""
import ctypes
import os
import re

import pytest

from conftest import ROOT, _has_gpu


def _declared():
    src = open(os.path.join(ROOT, "include", "ofb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ofb_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as ge
    ge.build()
    from ofb200 import _lib
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), "missing export: " + n
    assert set(names) == set(_lib._SIGNATURES), set(names) ^ set(_lib._SIGNATURES)
    assert lib.ofb_version() == 100


def test_struct_layouts_match_header(tmp_path):
    """sizeof of every struct as the C compiler lays it out == the ctypes / numpy mirrors."""
    import subprocess
    from ofb200 import _lib
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include "ofb200.h"\nint main(void){printf("%zu %zu %zu %zu %zu\\n",'
                   'sizeof(ofb_imu_sample),sizeof(ofb_pair_result),sizeof(ofb_mc_step),sizeof(ofb_mc_sums),'
                   'sizeof(ofb_pair_cfg));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)])
    sizes = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    assert sizes == [ctypes.sizeof(_lib.ImuSample), ctypes.sizeof(_lib.PairResult), ctypes.sizeof(_lib.McStep),
                     ctypes.sizeof(_lib.McSums), ctypes.sizeof(_lib.PairCfg)]
    assert sizes[0] == _lib.IMU_DTYPE.itemsize and sizes[1] == _lib.RESULT_DTYPE.itemsize
    assert sizes[3] == _lib.MCSUMS_DTYPE.itemsize
    # tracker structs: sizes and the offsets of the fields that follow padding
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "ofb200.h"\nint main(void){printf("%zu %zu %zu %zu %zu %zu %zu\\n",'
                   'sizeof(ofb_tracker_cfg),sizeof(ofb_track_result),offsetof(ofb_tracker_cfg,n_streams),'
                   'offsetof(ofb_tracker_cfg,max_speed),offsetof(ofb_tracker_cfg,gate_T),offsetof(ofb_tracker_cfg,borrow_frames),'
                   'offsetof(ofb_track_result,n_points));return 0;}\n')
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)])
    t = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    T, R = _lib.TrackerCfg, _lib.TrackResult
    assert t == [ctypes.sizeof(T), ctypes.sizeof(R), T.n_streams.offset, T.max_speed.offset, T.gate_T.offset,
                 T.borrow_frames.offset, R.n_points.offset]
    assert t[1] == _lib.TRACK_RESULT_DTYPE.itemsize and t[6] == _lib.TRACK_RESULT_DTYPE.fields["n_points"][1]


@pytest.mark.skipif(_has_gpu(), reason="checks the no-device failure mode")
def test_no_cpu_fallback():
    import ofb200
    with pytest.raises(ofb200.OfbError):
        ofb200.solve_lgs([[0.1, 0.2], [0.3, 0.1], [0.0, 0.4]], [[0, 0]] * 3, 1.0, [0, 0, 1], [0, 0, 0])
    with pytest.raises(ofb200.OfbError):
        ofb200.goodFeaturesToTrack(__import__("numpy").zeros((32, 32), "uint8"), 10, 0.01, 5)


def test_library_path_override(tmp_path):
    """OFB200_LIB selects another build of the library (the debug build with device-side index assertions); a path that
    does not exist must fail loudly -- there is no fallback to the default build or to the CPU."""
    import subprocess
    import sys
    code = ("import sys; sys.path.insert(0, %r); from ofb200 import _lib; print(_lib.LIB_PATH); "
            "lib = _lib.load(); print(lib.ofb_version())" % ROOT)
    env = dict(os.environ)
    from ofb200 import _lib
    env["OFB200_LIB"] = _lib.LIB_PATH if not os.environ.get("OFB200_LIB") else os.environ["OFB200_LIB"]
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.split()[-1] == "100", out.stderr[-400:]
    env["OFB200_LIB"] = str(tmp_path / "no_such_build.so")
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
    assert out.returncode != 0 and "not built" in (out.stderr + out.stdout)
