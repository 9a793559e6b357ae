"""The register-resident DFT-16 / DFT-25 and the two-step 256- and 400-point FFTs of the fbank kernels
(asr_streaming_b200/csrc/fft_regs.cuh), compiled for the host with g++ and checked against numpy.fft — the lane / exchange-buffer
index maps are emulated lane by lane exactly as the kernel runs them."""
import os
import shutil
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CUDA_INC = "/usr/local/cuda/include"


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    if shutil.which("g++") is None or not os.path.isdir(CUDA_INC):
        pytest.skip("g++ or the CUDA headers are not available")
    exe = str(tmp_path_factory.mktemp("fft") / "fft_regs_host")
    subprocess.run(["g++", "-O1", "-std=c++17", "-I", CUDA_INC, "-I", os.path.join(ROOT, "asr_streaming_b200", "csrc"),
                    os.path.join(ROOT, "tests", "fft_regs_host.cpp"), "-o", exe], check=True)
    return exe


def _run(exe, nc):
    rows = np.array([[float(v) for v in ln.split()] for ln in subprocess.run([exe, str(nc)], check=True, capture_output=True, text=True).stdout.splitlines()])
    return rows[:, 0] + 1j * rows[:, 1], rows[:, 2] + 1j * rows[:, 3]


@pytest.mark.parametrize("n", [16, 25])
def test_small_dfts(harness, n):
    x, y = _run(harness, n)
    assert np.abs(y - np.fft.fft(x)).max() < 5e-6


@pytest.mark.parametrize("nc", [256, 400])
def test_two_step_fft(harness, nc):
    x, y = _run(harness, nc)
    ref = np.fft.fft(x.reshape(2, nc), axis=1).reshape(-1)
    assert np.abs(y - ref).max() < 2e-5 * np.abs(ref).max()
