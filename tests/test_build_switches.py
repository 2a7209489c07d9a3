"""The measurement switches of the kernels must keep compiling: the evidence under profiles/ depends on builds with
-DFFB_RNN_PROFILE (phase counters, timeline stamps, group events: tools/rnn_phase_profile.py, tools/step_timeline.py) and
-DFFB_RNN_ABLATE=<mask> (timing-only ablations: tools/ablate_timing.py).  Compile-only (nvcc cross-compiles for sm_100a
without a GPU); the product build itself is exercised by __graft_entry__.build()."""
import os
import shutil
import subprocess

import pytest

from flappie_b200.build import CSRC, NVCC_FLAGS, _nvcc

CASES = [("rnn_tc.cu", ["-DFFB_RNN_PROFILE", "-DFFB_RNN_ABLATE=8191"]),
         ("gemm_tc.cu", ["-DFFB_RNN_PROFILE", "-DFFB_GEMM_TICKET_BATCH=1"])]


@pytest.mark.parametrize("src,flags", CASES)
def test_measurement_switches_compile(tmp_path, src, flags):
    if shutil.which("nvcc") is None and not os.path.exists("/usr/local/cuda/bin/nvcc"):
        pytest.skip("no nvcc")
    obj = tmp_path / (src + ".o")
    r = subprocess.run([_nvcc()] + NVCC_FLAGS + flags + ["-c", os.path.join(CSRC, src), "-o", str(obj)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    assert obj.stat().st_size > 0
