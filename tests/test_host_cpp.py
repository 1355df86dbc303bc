"""The C++ host mirror of the reference's Tracer API (light_garden_b200/host/lg_tracer.hpp) compiles against the
C ABI (CPU) and, on a GPU, produces the same trace as the Python host layer."""
import os
import re
import subprocess

import numpy as np
import pytest

from util import have_cuda

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "light_garden_b200", "_lib")
CXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"


def build(tmp_path, product_lib):
    exe = str(tmp_path / "host_cpp_smoke")
    cmd = [CXX, "-std=c++17", "-O1", "-Wall", "-o", exe, os.path.join(ROOT, "tests", "host_cpp_smoke.cpp"),
           "-L" + LIBDIR, "-llight_garden_b200", "-Wl,-rpath," + LIBDIR]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_cpp_host_mirror_compiles_and_links(tmp_path, product_lib):
    build(tmp_path, product_lib)


@pytest.mark.gpu
@pytest.mark.skipif(not have_cuda(), reason="no CUDA device")
def test_cpp_host_mirror_matches_python_host(tmp_path, product_lib):
    from light_garden_b200 import scenes
    from light_garden_b200.tracer import Tracer
    exe = build(tmp_path, product_lib)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
    m = re.search(r"vertices (\d+) segments (\d+) ray_steps (\d+)", r.stdout)
    cs = [float(x) for x in re.search(r"checksum (\S+) (\S+) (\S+)", r.stdout).groups()]
    pe = re.search(r"polygon\+ellipse vertices (\d+) \(tile map\) (\d+) \(all objects\)", r.stdout)
    assert pe and int(pe.group(1)) == int(pe.group(2)) > 4000
    assert "error -4" in r.stdout                      # LG_ERR_UNSUPPORTED crossed the boundary as a status
    spec = scenes.c1_default(total_rays=6000, width=480, height=270)
    t = spec.apply(Tracer(spec.canvas_bounds))
    seg = t.trace_all()
    assert int(m.group(1)) == 2 * len(seg) and int(m.group(2)) == len(seg) - 3
    assert int(m.group(3)) == t.last_stats.ray_steps
    sx = float(seg["a"][:, 0].astype(np.float64).sum() + seg["b"][:, 0].astype(np.float64).sum())
    sy = float(seg["a"][:, 1].astype(np.float64).sum() + seg["b"][:, 1].astype(np.float64).sum())
    sc = 2 * float(seg["color"][:, :3].astype(np.float64).sum())
    np.testing.assert_allclose(cs, [sx, sy, sc], rtol=1e-7)   # the digest is printed with 10 significant digits
