"""Golden fixtures (tests/golden/*.npz, made by tests/golden/make_golden.py from the f64/f32 oracle).

CPU: the oracle still reproduces them bit for bit (guards the checker against drift).
GPU: the device reproduces them bit for bit through the C ABI (no oracle code involved on that side).
"""
import os
import sys

import numpy as np
import pytest

from light_garden_b200 import abi
from util import have_cuda

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
from make_golden import golden_specs  # noqa: E402

NAMES = ["c1", "c2", "c3", "c5", "ell", "poly"]


def load(name):
    return np.load(os.path.join(HERE, "golden", f"{name}.npz"))


def same_bits(a, b):
    return a.dtype == b.dtype and a.shape == b.shape and a.tobytes() == b.tobytes()


@pytest.mark.parametrize("name", NAMES)
def test_oracle_reproduces_golden(oracle, name):
    g = load(name)
    spec = golden_specs()[name]
    osc = oracle.OracleScene.from_spec(spec)
    rays = np.ascontiguousarray(g["rays"]).view(abi.RAY_DTYPE).reshape(-1)
    for tag, prec in (("f64", abi.LG_PRECISION_F64), ("f32", abi.LG_PRECISION_F32)):
        res = osc.trace_rays(rays, prec)
        assert same_bits(res.seg, g[f"seg_{tag}"].view(abi.SEGMENT_DTYPE).reshape(-1))
        assert same_bits(res.tags, g[f"tags_{tag}"].view(abi.SEGMENT_TAG_DTYPE).reshape(-1))
    img = oracle.new_image(spec.width, spec.height)
    seg32 = np.ascontiguousarray(g["seg_f32"]).view(abi.SEGMENT_DTYPE).reshape(-1)
    assert oracle.accumulate_segments(img, seg32) == int(g["fragments"])
    assert np.array_equal(img.sum(axis=(0, 1), dtype=np.float64), g["image_sum"])


@pytest.mark.gpu
@pytest.mark.skipif(not have_cuda(), reason="no CUDA device")
@pytest.mark.parametrize("name", NAMES)
def test_device_reproduces_golden(name):
    from light_garden_b200.tracer import Context, Renderer, Tracer
    g = load(name)
    spec = golden_specs()[name]
    rays = np.ascontiguousarray(g["rays"]).view(abi.RAY_DTYPE).reshape(-1)
    for tag, prec in (("f64", abi.LG_PRECISION_F64), ("f32", abi.LG_PRECISION_F32)):
        ctx = Context(0, prec)
        try:
            t = spec.apply(Tracer(spec.canvas_bounds, ctx=ctx))
            seg, tags, f64 = t.trace(rays)
            assert same_bits(seg, g[f"seg_{tag}"].view(abi.SEGMENT_DTYPE).reshape(-1))
            assert same_bits(tags, g[f"tags_{tag}"].view(abi.SEGMENT_TAG_DTYPE).reshape(-1))
            if tag == "f64":
                assert same_bits(f64, g["end_f64"].view(abi.SEGMENT_F64_DTYPE).reshape(-1))
            else:
                r = Renderer(ctx, spec.width, spec.height)
                st = r.render_traced()
                assert st.pixel_updates == int(g["fragments"])
                s = r.read_rgba32f().sum(axis=(0, 1), dtype=np.float64)
                np.testing.assert_allclose(s, g["image_sum"], rtol=1e-6)
        finally:
            ctx.close()
