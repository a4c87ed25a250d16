"""CPU check of the inverse-FFT kernels' algebra: the device code of csrc/fft_kernels.cu compiled as
host C++ with host threads as the CTA (tests/host_emulation/fft_harness.cpp) against a long-double DFT.
Covers the generic radix-16/4/2 chirp-z kernel and the M = 2*16^p "halves" kernel (M = 32, 512,
8192), odd and even n_fft, the extremes of each M window and numpy's DC / Nyquist convention."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "popkinmocks_b200", "csrc")
CUDA_INC = os.environ.get("CUDA_HOME", "/usr/local/cuda") + "/include"


@pytest.mark.skipif(shutil.which("g++") is None or not os.path.isdir(CUDA_INC), reason="needs g++ and the CUDA headers")
def test_fft_kernels_host_emulation(tmp_path):
    src = open(os.path.join(CSRC, "fft_kernels.cu")).read()
    a = src.index("namespace {") + len("namespace {")
    b = src.index("// ----------------------------------------------------------------------------\n// four-step variant")
    body = re.sub(r"__launch_bounds__\(\w+\)", "", src[a:b]).replace("extern __shared__ double2 sbuf[];", "")
    (tmp_path / "body.inc").write_text(body)
    exe = tmp_path / "fft_harness"
    subprocess.run(["g++", "-O1", "-std=c++17", "-w", "-I", CUDA_INC, "-I", os.path.join(ROOT, "include"), "-I", CSRC,
                    "-I", str(tmp_path), os.path.join(ROOT, "tests", "host_emulation", "fft_harness.cpp"),
                    "-x", "c++", os.path.join(CSRC, "tables.cpp"), "-o", str(exe), "-lpthread"], check=True)
    sizes = [17, 20, 12, 21, 288, 300, 171, 341, 1021, 1992, 2731, 4096, 4907, 5461]
    out = subprocess.run([str(exe)] + [str(n) for n in sizes], check=True, capture_output=True, text=True).stdout
    lines = [l for l in out.splitlines() if l.startswith("n=")]
    assert len(lines) == len(sizes), out
    halves = 0
    for line in lines:
        m = re.match(r"n=(\d+) M=(\d+) generic (\S+)(.*)", line)
        assert m, line
        assert float(m.group(3)) < 1e-13, line
        errs = re.findall(r"halves (\S+) (\S+)", m.group(4))
        if errs:
            halves += 1
            assert int(m.group(2)) in (32, 512, 8192), line
            assert all(float(e) < 1e-13 for e in errs[0]), line   # first pixel, second (prefetched) pixel
    assert halves >= 9, out
