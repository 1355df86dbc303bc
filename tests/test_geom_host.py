"""The product's geometry header and scene lowering, instantiated for the host, agree bit for bit
with the oracle on random inputs (f32 and f64).  CPU only — catches transcription errors between
oracle/ORACLE.md's two independent implementations before any GPU time is spent."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NVCC = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


@pytest.mark.skipif(not os.path.exists(NVCC), reason="nvcc not available")
def test_device_geometry_matches_oracle_on_host(tmp_path):
    exe = str(tmp_path / "host_geom_check")
    ccbin = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    cmd = [NVCC, "-ccbin", ccbin, "-O2", "-std=c++17", "-fmad=false", "-Xcompiler", "-ffp-contract=off,-mfma",
           "-o", exe, os.path.join(ROOT, "tests", "host_geom_check.cu")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe, "100000"], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.startswith("OK"), r.stdout[-2000:]
