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


def seg_pairs(seg):
    """LgSegment records (one colour) as host vertex pairs."""
    from light_garden_b200 import abi
    p = np.zeros(len(seg), dtype=abi.VERTEX_PAIR_DTYPE)
    p["a"], p["b"], p["color_a"], p["color_b"] = seg["a"], seg["b"], seg["color"], seg["color"]
    return p


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
    # boundary B2 through the C++ Renderer: host LineList == fused device frame (worst deviation within the stated
    # 1e-5 tolerance), and the same fragments / image / 8-bit frames as the Python host layer
    from light_garden_b200.tracer import Renderer
    im = re.search(r"image fragments (\d+) sum (\S+) worst_tol_ratio (\S+) screenshot (\d+) surface (\d+)", r.stdout)
    assert im and float(im.group(3)) <= 1.0
    rend = Renderer(t.ctx, 480, 270)
    st = rend.render_lines(seg_pairs(seg))
    assert int(im.group(1)) == st.pixel_updates
    img = rend.read_rgba32f()
    np.testing.assert_allclose(float(im.group(2)), float(img.astype(np.float64).sum()), rtol=1e-5)
    assert abs(int(im.group(4)) - int(rend.make_screenshot().astype(np.int64).sum())) <= 1e-4 * int(im.group(4))
    assert abs(int(im.group(5)) - int(rend.read_surface_bgra8().astype(np.int64).sum())) <= 1e-4 * int(im.group(5))
    # Mode::StringMod through the C++ StringMod / Renderer::render_string_mod
    from light_garden_b200.scene import ModRemColor, StringMod, StringModMode
    smm = re.search(r"string_mod fragments (\d+) sum (\S+)", r.stdout)
    srend = Renderer(t.ctx, 256, 256)
    c = 1e-2
    st = srend.render_string_mod(StringMod(modulo=3000, num=2, mode=StringModMode.Mul, turns=1, color=[c] * 4,
                                           modulo_colors=[ModRemColor(3, 0, [c, 0, 0, c]), ModRemColor(3, 1, [0, c, 0, c]),
                                                          ModRemColor(3, 2, [0, 0, c, c])]))
    assert smm and int(smm.group(1)) == st.pixel_updates
    np.testing.assert_allclose(float(smm.group(2)), float(srend.read_rgba32f().astype(np.float64).sum()), rtol=1e-5)
