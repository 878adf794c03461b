"""Pins the register-FFT templates and the warp-level 512-point index arithmetic
(csrc/fft_regs.cuh, used by the fast kernel) on the CPU: the same header is compiled with
g++ and the lane-by-lane emulation is compared with numpy's rfft."""
import os
import shutil
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")
def test_warp_fft512_emulation(tmp_path):
    exe = tmp_path / "fft_check"
    subprocess.run(["g++", "-O2", "-std=c++17", "-o", str(exe), os.path.join(ROOT, "tests", "host", "fft_check.cpp")],
                   check=True)
    rng = np.random.default_rng(0)
    x = rng.uniform(-1, 1, (4, 512)).astype(np.float32)
    x[:, 400:] = 0
    out = subprocess.run([str(exe)], input=x.tobytes(), stdout=subprocess.PIPE, check=True).stdout
    got = np.frombuffer(out, dtype=np.float32).reshape(4, 256)
    ref = np.abs(np.fft.rfft(x.astype(np.float64), axis=1))[:, :256] ** 2
    rel = np.abs(got - ref) / (ref.mean())
    assert rel.max() < 5e-6, rel.max()
    np.testing.assert_allclose(got, ref, rtol=2e-3, atol=1e-4 * ref.mean())
