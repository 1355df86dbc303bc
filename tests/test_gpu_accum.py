"""Parity of the CUDA accumulation path (K3 string mod, K4 accumulate, K5 finalize) with the oracle.

Coverage (which pixels, how many fragments) is integer work: bit-exact.  Sums are fp32 reductions whose
order is not defined on the device (red.global.add): with power-of-two colours every partial sum is exact,
so those images are compared bit for bit; with arbitrary colours the tolerance is stated in the test.
"""
import ctypes as C

import numpy as np
import pytest

from light_garden_b200 import abi, scenes
from light_garden_b200.scene import Curve, ModRemColor, StringMod, StringModMode
from util import have_cuda, primary_rays, small_specs

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not have_cuda(), reason="no CUDA device")]


@pytest.fixture(scope="module")
def ctx():
    from light_garden_b200.tracer import Context
    c = Context(0, abi.LG_PRECISION_F32)
    yield c
    c.close()


DIRECT, TILED = 1, 2   # lg_accumulate_mode_set: one L2 reduction per fragment / shared-memory tile bins
MODES = [pytest.param(DIRECT, id="direct"), pytest.param(TILED, id="tiled")]


@pytest.fixture(autouse=True)
def _auto_mode_after(ctx):
    yield
    ctx.call("lg_accumulate_mode_set", 0)


def random_pairs(n, seed, span=2.2, pow2=True):
    rng = np.random.default_rng(seed)
    p = np.zeros(n, dtype=abi.VERTEX_PAIR_DTYPE)
    p["a"] = rng.uniform(-span, span, (n, 2))
    p["b"] = rng.uniform(-span, span, (n, 2))
    if pow2:
        p["color_a"] = 2.0 ** -rng.integers(4, 9, (n, 4))
        p["color_b"] = p["color_a"]
    else:
        p["color_a"] = rng.uniform(0, 0.02, (n, 4))
        p["color_b"] = rng.uniform(0, 0.02, (n, 4))
    return p


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("size", [(96, 54), (257, 131), (64, 64), (33, 200), (1, 1), (1, 300), (300, 1), (4097, 3)])
def test_pairs_exact_coverage_and_sums(oracle, ctx, size, mode):
    from light_garden_b200.tracer import Renderer
    ctx.call("lg_accumulate_mode_set", mode)
    W, H = size
    r = Renderer(ctx, W, H)
    p = random_pairs(3000, seed=W * 1000 + H)
    # degenerate and hostile inputs: zero length, fully outside, huge, axis-aligned on pixel centres
    p["b"][0] = p["a"][0]
    p["a"][1], p["b"][1] = (10.0, 10.0), (12.0, 11.0)
    p["a"][2], p["b"][2] = (-1e6, -1e6), (1e6, 1e6)
    p["a"][3], p["b"][3] = (0.0, 0.0), (0.5, 0.0)
    p["a"][4], p["b"][4] = (0.25, -3.0), (0.25, 3.0)
    # non-finite and out-of-range end points draw nothing (ORACLE.md 8.2), whichever end they are on
    nan, inf = float("nan"), float("inf")
    p["a"][5] = (nan, 0.0)
    p["b"][6] = (0.0, nan)
    p["a"][7] = (inf, 0.1)
    p["b"][8] = (0.2, -inf)
    p["a"][9], p["b"][9] = (-inf, -inf), (inf, inf)
    p["a"][10], p["b"][10] = (1e300, 0.0), (0.0, 0.0)          # overflows the `as f32` cast
    p["a"][11], p["b"][11] = (0.1, 0.1), (0.1 + 1e-12, 0.1 - 1e-12)   # shorter than a pixel: between two centres or on one
    st = r.render_lines(p)
    got = r.read_rgba32f()
    exp = oracle.new_image(W, H)
    n = oracle.accumulate_pairs(exp, p)
    assert st.pixel_updates == n
    assert np.array_equal(got, exp)


@pytest.mark.parametrize("mode", MODES)
def test_two_colour_lerp_matches_oracle(oracle, ctx, mode):
    """Arbitrary per-endpoint colours: same fragments; sums within fp32 reordering error."""
    from light_garden_b200.tracer import Renderer
    ctx.call("lg_accumulate_mode_set", mode)
    W, H = 160, 90
    r = Renderer(ctx, W, H)
    p = random_pairs(5000, seed=11, pow2=False)
    st = r.render_lines(p)
    got = r.read_rgba32f()
    exp = oracle.new_image(W, H)
    assert st.pixel_updates == oracle.accumulate_pairs(exp, p)
    # <= ~60 adds of O(0.02) per pixel: reordering error far below 1e-6 absolute; stated tolerance 2e-6
    assert np.abs(got - exp).max() < 2e-6
    assert np.array_equal(got == (0, 0, 0, 1), exp == (0, 0, 0, 1))          # identical coverage


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("name", ["C1", "C5-16"])
def test_traced_segments_image(oracle, ctx, name, mode):
    """trace -> accumulate on the device vs the oracle accumulating the device's own segments, and
    lg_render (waves through a small segment buffer) vs the one-shot path."""
    from light_garden_b200.tracer import Renderer, Tracer
    ctx.call("lg_accumulate_mode_set", mode)
    spec = small_specs()[name]
    t = spec.apply(Tracer(spec.canvas_bounds, ctx=ctx))
    r = Renderer(ctx, spec.width, spec.height)
    seg = t.trace_all(ordered=False, control_lines=False)
    st = r.render_traced()
    got = r.read_rgba32f()
    exp = oracle.new_image(spec.width, spec.height)
    n = oracle.accumulate_segments(exp, seg)
    assert st.pixel_updates == n
    assert np.array_equal(got[..., 3] > 1, exp[..., 3] > 1)                    # same covered pixels
    err = np.abs(got - exp)
    # hot pixels next to a light sum thousands of equal fragments.  The direct resolve adds them one by one like
    # the oracle's fp32 loop (same rounding pattern: 1e-5 of the pixel value); the tiled resolve adds per-tile
    # partial sums, a different (more accurate) summation tree: it is bounded against the f64 sum instead.
    tol = 1e-5 if mode == DIRECT else 3e-4
    assert (err <= tol * np.maximum(1.0, np.abs(exp))).all(), err.max()
    exact = np.zeros((spec.height, spec.width, 4), dtype=np.float64)
    exact[..., 3] = 1.0
    oracle.accumulate_segments_f64(exact, seg)
    rel64 = np.abs(got - exact) / np.maximum(1.0, np.abs(exact))
    assert rel64.max() < (3e-4 if mode == DIRECT else 2e-5), rel64.max()
    mse = float(np.mean((got - exp) ** 2))
    psnr = 10 * np.log10(float(exp.max()) ** 2 / mse) if mse > 0 else np.inf
    assert psnr > 100
    # waves: capacity for only a fraction of the segments
    ctx.call("lg_segment_capacity_set", max(2048, len(seg) // 5))
    try:
        r.clear()
        st2 = r.render(t)
        assert st2.segments == len(seg) and st2.pixel_updates == n and st2.trace_launches >= 4
        got2 = r.read_rgba32f()
        assert (np.abs(got2 - exp) <= tol * np.maximum(1.0, np.abs(exp))).all()
    finally:
        ctx.call("lg_segment_capacity_set", 64 << 20)


@pytest.mark.parametrize("curve", [Curve.ComplexExp(complex(0.9995, 0.01)), Curve.Hypotrochoid(3, 7, 2),
                                   Curve.Lissajous(3, 2, 0.5)], ids=["complex_exp", "hypotrochoid", "lissajous"])
def test_string_mod_other_curves(oracle, ctx, curve):
    """the non-circle init curves of string_mod.rs:36-84."""
    from light_garden_b200.tracer import Renderer
    W = H = 256
    k = 2.0 ** -8
    sm = StringMod(modulo=1500, num=7, mode=StringModMode.Mul, color=(k, k, k, k), init_curve=curve)
    r = Renderer(ctx, W, H)
    st = r.render_string_mod(sm)
    got = r.read_rgba32f()
    exp = oracle.new_image(W, H)
    n = oracle.accumulate_pairs(exp, oracle.string_mod(sm))
    assert n > 20000
    # end points pass through device sin/cos (or 10 complex squarings) vs libm: a handful of fragments may move
    assert abs(int(st.pixel_updates) - int(n)) <= 8
    assert len(np.nonzero((got != exp).any(axis=2))[0]) <= 32


@pytest.mark.parametrize("mode", MODES)
def test_string_mod_matches_oracle(oracle, ctx, mode):
    from light_garden_b200.tracer import Renderer
    ctx.call("lg_accumulate_mode_set", mode)
    W = H = 256
    k = 2.0 ** -8
    rules = [ModRemColor(3, 0, (k, 0, 0, k)), ModRemColor(3, 1, (0, k, 0, k)), ModRemColor(3, 2, (0, 0, k, k))]
    for mode, num, m in ((StringModMode.Mul, 2, 3001), (StringModMode.Add, 977, 2048), (StringModMode.Pow, 3, 1000),
                         (StringModMode.Base, 3, 500), (StringModMode.Mul, 7919, 4099)):
        sm = StringMod(modulo=m, num=num, mode=mode, color=(k, k, k, k), modulo_colors=rules)
        r = Renderer(ctx, W, H)
        st = r.render_string_mod(sm)
        got = r.read_rgba32f()
        exp = oracle.new_image(W, H)
        n = oracle.accumulate_pairs(exp, oracle.string_mod(sm))
        # chord end points come from sincos on both sides (device vs libm: last-place differences in f64
        # that survive the cast to f32 only rarely) -> allow a handful of fragments to move; the two end
        # points of a chord have different colours, so the lerped sums differ by fp32 reordering only
        assert abs(int(st.pixel_updates) - int(n)) <= 8, (mode, st.pixel_updates, n)
        diff = np.nonzero((np.abs(got - exp) > 1e-6 * np.maximum(1.0, np.abs(exp))).any(axis=2))
        assert len(diff[0]) <= 16, (mode, len(diff[0]))
        assert st.segments == m
    # sub-range + shard semantics: two halves add up to the whole (one colour, a power of two: exact sums)
    sm = StringMod(modulo=3001, num=2, mode=StringModMode.Mul, color=(k, k, k, k))
    r = Renderer(ctx, W, H)
    r.render_string_mod(sm)
    whole = r.read_rgba32f()
    r.clear()
    r.render_string_mod(sm, first=0, count=1500)
    r.render_string_mod(sm, first=1500, count=1501)
    assert np.array_equal(r.read_rgba32f(), whole)
    # modulo = 0 draws nothing (string_mod.rs:106-108)
    r.clear()
    assert r.render_string_mod(StringMod(modulo=0)).segments == 0


@pytest.mark.parametrize("mode", MODES)
def test_nested_string_mod(oracle, ctx, mode):
    """StringMod with nested = Some(inner) (string_mod.rs:87-101,152-158).  The crossing
    points of the device's own outer chords must equal the oracle's on the same chords bit for bit, in the
    reference's order; the inner pattern drawn between them must match the oracle's line pass."""
    from light_garden_b200.tracer import Renderer
    ctx.call("lg_accumulate_mode_set", mode)
    W = H = 256
    k = 2.0 ** -8
    inner = StringMod(modulo=5000, num=3, mode=StringModMode.Mul, color=(k, k, k, k),
                      modulo_colors=[ModRemColor(3, 0, (k, 0, 0, k)), ModRemColor(5, 1, (0, k, 0, k))])
    for outer in (StringMod(modulo=120, num=2, mode=StringModMode.Mul, nested=inner),
                  StringMod(modulo=257, num=100, mode=StringModMode.Add, nested=inner,
                            init_curve=Curve.Lissajous(3, 2, 0.5))):
        r = Renderer(ctx, W, H)
        st = r.render_string_mod(outer)
        got = r.read_rgba32f()
        lines, pts = r.nested_crossings()
        # the outer chords themselves: device sin/cos vs libm, a few ulps of f64
        ref_lines = oracle.string_mod(StringMod(modulo=outer.modulo, num=outer.num, mode=outer.mode, init_curve=outer.init_curve))
        assert len(lines) == outer.modulo
        np.testing.assert_allclose(lines["a"], ref_lines["a"], atol=1e-15)
        np.testing.assert_allclose(lines["b"], ref_lines["b"], atol=1e-15)
        # crossings of THOSE chords: same points, same order, same bits
        exp_pts = oracle.line_crossings(lines)
        assert len(exp_pts) > 1000
        assert pts.shape == exp_pts.shape and pts.tobytes() == exp_pts.tobytes()
        # the inner chords between them through the line pass
        exp = oracle.new_image(W, H)
        n = oracle.accumulate_pairs(exp, oracle.nested_chords(inner, pts))
        assert st.segments == inner.modulo and int(st.pixel_updates) == int(n) > 100000
        assert (np.abs(got - exp) <= 1e-6 * np.maximum(1.0, np.abs(exp))).all()
    # an outer pattern without crossings draws nothing
    r = Renderer(ctx, W, H)
    assert r.render_string_mod(StringMod(modulo=1, nested=inner)).segments == 0
    assert np.array_equal(r.read_rgba32f(), oracle.new_image(W, H))


BLENDS = {
    "max": ((abi.LG_BF_ONE, abi.LG_BF_ONE, abi.LG_BO_MAX), (abi.LG_BF_ONE, abi.LG_BF_ONE, abi.LG_BO_MAX)),
    "min_alpha": ((abi.LG_BF_ONE, abi.LG_BF_ONE, abi.LG_BO_ADD), (abi.LG_BF_ONE, abi.LG_BF_ONE, abi.LG_BO_MIN)),
    "src_alpha": ((abi.LG_BF_SRC_ALPHA, abi.LG_BF_ONE, abi.LG_BO_ADD), (abi.LG_BF_ONE, abi.LG_BF_ONE, abi.LG_BO_ADD)),
    "rev_sub_constant": ((abi.LG_BF_CONSTANT, abi.LG_BF_ONE, abi.LG_BO_REVERSE_SUBTRACT),
                         (abi.LG_BF_ONE_MINUS_SRC_ALPHA, abi.LG_BF_ONE, abi.LG_BO_ADD)),
}


@pytest.mark.parametrize("name", list(BLENDS))
def test_blend_states_match_oracle(oracle, ctx, name):
    """the order-independent blend states of gui/settings.rs:59-127 (lg_blend_set).  Power-of-two
    colours make every Add exact, Min / Max are exact anyway: the image must equal the oracle's bit for bit."""
    from light_garden_b200.tracer import Renderer
    color, alpha = BLENDS[name]
    const = (0.5, 0.25, 2.0, 1.0)
    ctx.call("lg_accumulate_mode_set", 0)
    r = Renderer(ctx, 320, 200)
    p = random_pairs(4000, seed=33, pow2=True)
    try:
        r.set_blend(color, alpha, const)
        st = r.render_lines(p)
        got = r.read_rgba32f()
        exp = oracle.new_image(320, 200)
        n = oracle.accumulate_pairs_blend(exp, p, color, alpha, const)
        assert int(st.pixel_updates) == int(n) > 50000
        assert np.array_equal(got, exp)
        # string mod chords and traced segments go through the same blend
        k = 2.0 ** -6
        sm = StringMod(modulo=700, num=3, mode=StringModMode.Mul, color=(k, 2 * k, 4 * k, 8 * k))
        r.clear()
        r.render_string_mod(sm)
        exp = oracle.new_image(320, 200)
        oracle.accumulate_pairs_blend(exp, oracle.string_mod(sm), color, alpha, const)
        diff = np.nonzero((r.read_rgba32f() != exp).any(axis=2))
        assert len(diff[0]) <= 16          # chord end points: device sincos vs libm
    finally:
        r.set_blend()
    # back on the default state the dedicated kernels run again
    r.clear()
    r.render_lines(p)
    exp = oracle.new_image(320, 200)
    oracle.accumulate_pairs(exp, p)
    assert np.array_equal(r.read_rgba32f(), exp)


def test_blend_states_that_depend_on_the_fragment_order_are_refused(ctx):
    from light_garden_b200._lib import LightGardenError
    from light_garden_b200.tracer import Renderer
    r = Renderer(ctx, 64, 64)
    ONE, ADD = abi.LG_BF_ONE, abi.LG_BO_ADD
    for color in ((ONE, abi.LG_BF_ZERO, ADD),                         # dst factor != One: later fragments erase earlier ones
                  (abi.LG_BF_DST, ONE, ADD),                          # source factor reads the image
                  (ONE, ONE, abi.LG_BO_SUBTRACT),                     # src - dst
                  (abi.LG_BF_SRC_ALPHA_SATURATED, ONE, ADD)):
        with pytest.raises(LightGardenError) as e:
            r.set_blend(color, None)
        assert e.value.code == abi.LG_ERR_UNSUPPORTED
    with pytest.raises(LightGardenError):
        r.set_blend((99, ONE, ADD), None)
    # the tile-binned resolve implements the default state only
    r.set_blend((ONE, ONE, abi.LG_BO_MAX), None)
    try:
        with pytest.raises(LightGardenError):
            ctx.call("lg_accumulate_mode_set", 2)
    finally:
        r.set_blend()
        ctx.call("lg_accumulate_mode_set", 0)


def test_finalize_rgba16f(oracle, ctx):
    """K5: the Rgba16Float image is the fp32 image rounded to nearest even, bit for bit."""
    from light_garden_b200.tracer import Renderer
    W, H = 128, 72
    r = Renderer(ctx, W, H)
    p = random_pairs(4000, seed=5, pow2=False)
    p["color_a"] *= 50       # push some pixels past fp16's integer range
    p["color_b"] *= 50
    r.render_lines(p)
    f32 = r.read_rgba32f()
    f16 = r.read_rgba16f()
    assert np.array_equal(f16.view(np.uint16), oracle.to_f16(f32).view(np.uint16))
    assert np.array_equal(f16.view(np.uint16), f32.astype(np.float16).view(np.uint16))


def test_screenshot_bgra8(oracle, ctx):
    """Renderer::make_screenshot's fp16 -> gamma u8 conversion (renderer.rs:313-328) and the
    256-byte padded row layout of its readback (renderer.rs:250-255)."""
    from light_garden_b200.tracer import Renderer
    W, H = 100, 40          # 400-byte rows: padded to 512
    r = Renderer(ctx, W, H)
    p = random_pairs(3000, seed=21, pow2=False)
    p["color_a"] *= 20
    p["color_b"] *= 20
    r.render_lines(p)
    f32 = r.read_rgba32f()
    exp = oracle.to_bgra8(f32)
    got = r.make_screenshot()
    # powf in f64-then-rounded (device) vs libm powf (oracle): identical except when f^(1/2.2)*255 sits within an
    # ulp of an integer
    diff = np.abs(got.astype(np.int32) - exp.astype(np.int32))
    assert diff.max() <= 1 and (diff != 0).mean() < 1e-3
    assert got[..., 3].min() == 255                  # alpha >= 1 saturates
    padded = r.make_screenshot(pitch=512)
    assert np.array_equal(padded, got)


def test_surface_bgra8_srgb(oracle, ctx):
    """the 8-bit surface target of the path with render_to_texture off
    (sub_render_pass.rs:59-63; Bgra8UnormSrgb, renderer.rs:207-209), ORACLE.md 8.7 -- bit for bit, padded rows too."""
    from light_garden_b200.tracer import Renderer
    W, H = 100, 40
    r = Renderer(ctx, W, H)
    p = random_pairs(3000, seed=22, pow2=False)
    p["color_a"] *= 6        # part of the frame saturates, most of it does not
    p["color_b"] *= 6
    r.render_lines(p)
    f32 = r.read_rgba32f()
    got = r.read_surface_bgra8()
    assert np.array_equal(got, oracle.to_bgra8_srgb(f32))
    assert got[..., 3].min() == 255 and 0 < (got[..., :3] == 255).mean() < 0.9 and (got[..., :3] == 0).any()
    assert np.array_equal(r.read_surface_bgra8(pitch=512), got)
    r.clear(0.0)
    assert not r.read_surface_bgra8().any()


def test_exported_frame_is_importable_device_memory(oracle, ctx):
    """SURVEY.md §8f rank 3 (display hand-off without the host bounce, renderer.rs:356-417): the frame as device memory
    behind a POSIX file descriptor.  A consumer that imports the descriptor (here cuMemImportFromShareableHandle through
    lg_import_fd_read; Vulkan's OPAQUE_FD import takes the same handle) sees exactly what lg_image_read returns."""
    import os
    from light_garden_b200 import abi
    from light_garden_b200._lib import LightGardenError, check, load
    from light_garden_b200.tracer import Renderer
    W, H = 200, 120
    r = Renderer(ctx, W, H)
    r.render_lines(random_pairs(3000, seed=23, pow2=False))
    try:
        fd, nbytes = r.export_fd(abi.LG_RGBA16F)
    except LightGardenError as e:                      # a driver that cannot export says so instead of faking it
        assert e.code == abi.LG_ERR_UNSUPPORTED
        pytest.skip("driver cannot export device memory: " + e.message)
    try:
        assert fd >= 0 and nbytes >= W * H * 8
        r.export_refresh(abi.LG_RGBA16F)
        got = np.zeros((H, W, 4), dtype=np.float16)
        check(None, load().lg_import_fd_read(0, fd, nbytes, abi.array_ptr(got), got.nbytes))
        assert np.array_equal(got.view(np.uint16), r.read_rgba16f().view(np.uint16))
        # the same memory after more lines and another refresh: importers keep their mapping
        r.render_lines(random_pairs(500, seed=24, pow2=False))
        r.export_refresh(abi.LG_RGBA16F)
        check(None, load().lg_import_fd_read(0, fd, nbytes, abi.array_ptr(got), got.nbytes))
        assert np.array_equal(got.view(np.uint16), r.read_rgba16f().view(np.uint16))
    finally:
        os.close(fd)
    # the 8-bit surface frame through a second descriptor
    fd8, n8 = r.export_fd(abi.LG_BGRA8_SRGB)
    try:
        r.export_refresh(abi.LG_BGRA8_SRGB)
        got8 = np.zeros((H, W, 4), dtype=np.uint8)
        check(None, load().lg_import_fd_read(0, fd8, n8, abi.array_ptr(got8), got8.nbytes))
        assert np.array_equal(got8, r.read_surface_bgra8())
    finally:
        os.close(fd8)
    with pytest.raises(LightGardenError):
        r.export_refresh(abi.LG_RGBA32F)               # never exported in this format: a state error, not a silent no-op


def test_clear_value_and_partial_alpha(ctx):
    from light_garden_b200.tracer import Renderer
    r = Renderer(ctx, 40, 30)
    img = r.read_rgba32f()
    assert np.all(img[..., :3] == 0) and np.all(img[..., 3] == 1)       # LoadOp::Clear(BLACK)
    r.clear(0.0)                                                          # non-owning ranks of a multi-GPU frame
    assert np.all(r.read_rgba32f() == 0)


@pytest.mark.parametrize("wh", [(4096, 4096)])
def test_full_size_string_mod_properties(ctx, wh):
    """C4 at BASELINE size (10 M chords, 4096 x 4096) through size-independent properties: with a
    power-of-two colour every pixel holds count * colour exactly, so the image sums are a checksum of the
    fragment counter; the two halves of the chord range add up to the whole."""
    from light_garden_b200.tracer import Renderer
    W, H = wh
    k = 2.0 ** -12
    sm = StringMod(modulo=10_000_000, num=2, mode=StringModMode.Mul, color=(k, k, k, k))
    r = Renderer(ctx, W, H)
    r.clear(0.0)        # alpha starts at 0 so that count * 2^-24 stays exact (1 + 2^-24 is not an fp32 number)
    st = r.render_string_mod(sm)
    img = r.read_rgba32f()
    n = int(st.pixel_updates)
    assert st.segments == 10_000_000
    # mean chord of the unit circle is 4/pi world units = 4/pi * 2048 px; DDA steps ~ x (2 sqrt2 / pi)
    assert 2.2e10 < n < 2.5e10
    for ch in range(3):
        assert int(round(float(img[..., ch].astype(np.float64).sum()) / k)) == n
    assert int(round(float(img[..., 3].astype(np.float64).sum()) / (k * k))) == n
    assert float(img[..., 0].max()) / k < 2 ** 24                       # counts stayed exactly representable
    # the i -> 2i pattern (a cardioid envelope) is mirror symmetric about the x axis: top and bottom halves balance
    top, bottom = img[: H // 2, :, 0].sum(dtype=np.float64), img[H // 2:, :, 0].sum(dtype=np.float64)
    assert abs(top - bottom) / (top + bottom) < 0.01
    r.clear(0.0)
    r.render_string_mod(sm, first=0, count=5_000_000)
    r.render_string_mod(sm, first=5_000_000, count=5_000_000)
    assert np.array_equal(r.read_rgba32f(), img)
    # the direct and the tile-binned resolve agree bit for bit (every partial sum is exact with this colour)
    for mode in (DIRECT, TILED):
        ctx.call("lg_accumulate_mode_set", mode)
        r.clear(0.0)
        st2 = r.render_string_mod(sm)
        assert int(st2.pixel_updates) == n
        assert np.array_equal(r.read_rgba32f(), img), mode


def test_wide_fill_pass_gives_the_same_frame(oracle):
    """tile_fill has a form for pair lists of 2^32 entries and more (C2 at full size: 5.7e9 pairs; 64-bit list starts
    staged in shared memory).  Forced here on a small frame: same fragments, same sums as the 32-bit form."""
    import os
    from light_garden_b200.tracer import Context, Renderer, Tracer
    spec = small_specs()["C2"]
    frames = []
    for wide in ("0", "1"):
        os.environ["LG_FILL_WIDE"] = wide
        try:
            c = Context(0, abi.LG_PRECISION_F32)
        finally:
            del os.environ["LG_FILL_WIDE"]
        try:
            c.call("lg_accumulate_mode_set", TILED)
            t = spec.apply(Tracer(spec.canvas_bounds, ctx=c))
            r = Renderer(c, spec.width, spec.height)
            r.clear()
            st = r.render(t)
            frames.append((st.pixel_updates, st.segments, r.read_rgba32f()))
        finally:
            c.close()
    assert frames[0][0] == frames[1][0] > 0 and frames[0][1] == frames[1][1]
    # entries land in a list in a different order (atomics): fp32 association within a tile only
    assert (np.abs(frames[0][2] - frames[1][2]) <= 2e-5 * np.maximum(1.0, np.abs(frames[0][2]))).all()
    # a 4096 x 4096 frame has 32768 lists: their 64-bit starts no longer fit in shared memory beside the cursors and are
    # read from global memory instead (a round-2 bench run failed right here: C4 after C2 had grown the list past 2^32)
    k = 2.0 ** -8
    sm = StringMod(modulo=20000, num=7, mode=StringModMode.Mul, color=(k, k, k, k))
    counts = []
    for wide in ("0", "1"):
        os.environ["LG_FILL_WIDE"] = wide
        try:
            c = Context(0, abi.LG_PRECISION_F32)
        finally:
            del os.environ["LG_FILL_WIDE"]
        try:
            c.call("lg_accumulate_mode_set", TILED)
            r = Renderer(c, 4096, 4096)
            r.clear()
            counts.append(r.render_string_mod(sm).pixel_updates)
        finally:
            c.close()
    assert counts[0] == counts[1] > 20000 * 1000


@pytest.mark.parametrize("mode", MODES)
def test_colours_that_are_not_numbers_poison_their_own_pixels_only(oracle, ctx, mode):
    """NaN, infinite, negative and huge colours in the vertex pairs: the pixels those lines cover hold what IEEE addition
    makes of them (NaN, +-inf, the oracle's sums), every other pixel is untouched, and the 8- and 16-bit read-backs follow
    ORACLE.md 8.5-8.7 (NaN -> 0)."""
    from light_garden_b200.tracer import Renderer
    ctx.call("lg_accumulate_mode_set", mode)
    W, H = 160, 90
    r = Renderer(ctx, W, H)
    p = random_pairs(400, seed=77)
    nan, inf = float("nan"), float("inf")
    p["color_a"][0], p["color_b"][0] = (nan, 0.1, 0.1, 0.1), (nan, 0.1, 0.1, 0.1)
    p["color_a"][1], p["color_b"][1] = (inf, 0.1, 0.1, 0.1), (inf, 0.1, 0.1, 0.1)
    p["color_a"][2], p["color_b"][2] = (-inf, 0.1, 0.1, 0.1), (inf, 0.1, 0.1, 0.1)          # lerp from -inf to +inf
    p["color_a"][3], p["color_b"][3] = (-0.5, -0.5, -0.5, -0.5), (-0.5, -0.5, -0.5, -0.5)
    p["color_a"][4], p["color_b"][4] = (3e38, 3e38, 3e38, 1e19), (3e38, 3e38, 3e38, 1e19)   # sums and alpha^2 overflow
    st = r.render_lines(p)
    got = r.read_rgba32f()
    exp = oracle.new_image(W, H)
    assert st.pixel_updates == oracle.accumulate_pairs(exp, p)
    assert np.array_equal(np.isnan(got), np.isnan(exp)) and np.array_equal(np.isinf(got), np.isinf(exp))
    fin = np.isfinite(exp)
    assert (np.abs(got[fin] - exp[fin]) <= 2e-6 * np.maximum(1.0, np.abs(exp[fin]))).all()
    assert np.array_equal(r.read_rgba16f().view(np.uint16) & 0x7fff > 0x7c00, np.isnan(got))   # NaN stays NaN in fp16
    # the 8-bit conversions of the device's own fp32 frame: NaN and negative channels give 0, +inf saturates
    d8 = np.abs(r.make_screenshot().astype(np.int32) - oracle.to_bgra8(got).astype(np.int32))
    assert d8.max() <= 1 and (d8 != 0).mean() < 1e-3          # powf: see test_screenshot_bgra8
    assert np.array_equal(r.read_surface_bgra8(), oracle.to_bgra8_srgb(got))
