"""Known-answer tests that pin the oracle (oracle/lg_oracle.hpp) to analytic results.

The reference ships no tests and no golden vectors and its geometry crate
collision2d is not available offline, so parity at that boundary is UNPINNED; these KATs are the
closed-form checks SURVEY.md §4 lists, authored here.  CPU only.
"""
import math

import numpy as np
import pytest

from light_garden_b200 import abi, scenes
from light_garden_b200.scene import (AND, AND_NOT, OR, Circle, CubicBezier, DirectionalLight, LineSegment, Logic,
                                     Material, ModRemColor, Object, PointLight, Rect, SpotLight, StringMod,
                                     StringModMode, rot2, rot2_identity)
from light_garden_b200.scene import Curve, Ellipse

CANVAS = scenes.canvas(16.0 / 9.0)
CUT = [0.001] * 4


def scene(oracle, objects, max_bounce=5, canvas=CANVAS):
    return oracle.OracleScene(objects, max_bounce, CUT, canvas)


def ray(o, d, color=(0.5, 0.5, 0.5, 0.5), n=1.0):
    r = np.zeros(1, dtype=abi.RAY_DTYPE)
    ln = math.hypot(*d)
    r["origin"], r["direction"], r["color"], r["refractive_index"] = o, (d[0] / ln, d[1] / ln), color, n
    return r


# ---- primitives -----------------------------------------------------------------------------------
def test_ray_circle_hand_computed(oracle):
    sc = scene(oracle, [Object.new_circle((2.0, 0.0), 0.5)])
    h = sc.intersect(0, (0, 0), (1, 0))
    assert h.shape[0] == 2
    np.testing.assert_allclose(h[0], [1.5, 0, -1, 0, 1.5], atol=1e-15)
    np.testing.assert_allclose(h[1], [2.5, 0, 1, 0, 2.5], atol=1e-15)
    # from inside: only the exit point
    h = sc.intersect(0, (2.0, 0.0), (0, 1))
    assert h.shape[0] == 1
    np.testing.assert_allclose(h[0], [2.0, 0.5, 0, 1, 0.5], atol=1e-15)
    # offset chord: y = 0.3 -> x = 2 -/+ 0.4
    h = sc.intersect(0, (0, 0.3), (1, 0))
    np.testing.assert_allclose(h[:, 0], [1.6, 2.4], atol=1e-14)
    np.testing.assert_allclose(h[0, 2:4], [-0.8, 0.6], atol=1e-14)
    # miss, and behind the origin
    assert sc.intersect(0, (0, 0.6), (1, 0)).shape[0] == 0
    assert sc.intersect(0, (3, 0), (1, 0)).shape[0] == 0


def test_ray_circle_self_hit_rejected(oracle):
    """A bounce origin lies on the surface: the t~0 root is dropped (t > T_MIN = 1e-5)."""
    sc = scene(oracle, [Object.new_circle((0.0, 0.0), 1.0)])
    h = sc.intersect(0, (1.0, 0.0), (-1, 0))
    assert h.shape[0] == 1 and abs(h[0, 4] - 2.0) < 1e-14
    assert sc.intersect(0, (1.0, 0.0), (1, 0)).shape[0] == 0


def test_ray_segment(oracle):
    sc = scene(oracle, [Object.new_mirror((1.0, -1.0), (1.0, 1.0))])
    h = sc.intersect(0, (0, 0.25), (1, 0))
    assert h.shape[0] == 1
    np.testing.assert_allclose(h[0, [0, 1, 4]], [1.0, 0.25, 1.0], atol=1e-15)
    assert abs(abs(h[0, 2]) - 1.0) < 1e-15 and abs(h[0, 3]) < 1e-15
    assert sc.intersect(0, (0, 1.5), (1, 0)).shape[0] == 0      # beyond the end point
    assert sc.intersect(0, (0, 0), (0, 1)).shape[0] == 0        # parallel
    assert sc.intersect(0, (2, 0), (1, 0)).shape[0] == 0        # behind
    # end points are inclusive (u in [0, 1])
    assert sc.intersect(0, (0, 1.0), (1, 0)).shape[0] == 1


def test_ray_oriented_rect(oracle):
    # 2 x 1 rect rotated by 90 degrees about (3, 0): spans x in [2.5, 3.5], y in [-1, 1]
    r = Object(Rect((3.0, 0.0), rot2(math.pi / 2), 2.0, 1.0), Material(1.5))
    sc = scene(oracle, [r])
    h = sc.intersect(0, (0, 0.2), (1, 0))
    assert h.shape[0] == 2
    xs = sorted(h[:, 0])
    np.testing.assert_allclose(xs, [2.5, 3.5], atol=1e-12)
    for row in h:
        assert abs(abs(row[2]) - 1) < 1e-12 and abs(row[3]) < 1e-12
    assert sc.intersect(0, (0, 1.2), (1, 0)).shape[0] == 0
    assert sc.contains(0, (3.0, 0.9)) and not sc.contains(0, (3.6, 0.0)) and not sc.contains(0, (3.0, 1.01))


def test_rect_edge_order_is_right_bottom_left_top(oracle):
    """Rect::line_segments() order [right, bottom, left, top] (grid.rs:31) is the order hits are listed in."""
    sc = scene(oracle, [Object.new_rect((0.0, 0.0), 2.0, 2.0)])
    h = sc.intersect(0, (-3, 0.1), (1, 0))   # crosses left then right; listed right first
    np.testing.assert_allclose(h[:, 0], [1.0, -1.0], atol=1e-15)
    h = sc.intersect(0, (0.1, -3), (0, 1))   # crosses bottom then top; listed bottom first
    np.testing.assert_allclose(h[:, 1], [-1.0, 1.0], atol=1e-15)


def test_bezier_degenerate_to_line(oracle):
    # control points on the line x = 1, evenly spaced: the curve is the segment (1,-1)-(1,1)
    pts = ((1.0, -1.0), (1.0, -1.0 / 3), (1.0, 1.0 / 3), (1.0, 1.0))
    sc = scene(oracle, [Object.new_curved_mirror(CubicBezier(pts))])
    for y in (-0.9, -0.3, 0.0, 0.5, 0.99):
        h = sc.intersect(0, (0, y), (1, 0))
        assert h.shape[0] == 1
        np.testing.assert_allclose(h[0, [0, 1]], [1.0, y], atol=1e-12)
        assert abs(abs(h[0, 2]) - 1) < 1e-12
    assert sc.intersect(0, (0, 1.1), (1, 0)).shape[0] == 0


def test_bezier_multiple_roots(oracle):
    # an S-shaped cubic crossed three times by the x axis
    pts = ((0.0, -0.5), (1.0, 2.0), (2.0, -2.0), (3.0, 0.5))
    sc = scene(oracle, [Object.new_curved_mirror(CubicBezier(pts))])
    h = sc.intersect(0, (-1, 0), (1, 0))
    assert h.shape[0] == 3
    assert np.all(np.diff(h[:, 0]) > 0)          # listed in curve-parameter order = increasing x here
    for row in h:                                  # each point is on the curve: solve B_y(t) = 0 via numpy
        assert abs(row[1]) < 1e-12
    ys = np.array([p[1] for p in pts])
    coef = [ys[3] - ys[0] + 3 * (ys[1] - ys[2]), 3 * (ys[0] - 2 * ys[1] + ys[2]), 3 * (ys[1] - ys[0]), ys[0]]
    ts = sorted(t.real for t in np.roots(coef) if abs(t.imag) < 1e-12 and 0 <= t.real <= 1)
    xs = [(1 - t) ** 3 * 0 + 3 * (1 - t) ** 2 * t * 1 + 3 * (1 - t) * t * t * 2 + t ** 3 * 3 for t in ts]
    np.testing.assert_allclose(h[:, 0], xs, atol=1e-12)


# ---- Snell / Fresnel ---------------------------------------------------------------------------------
@pytest.mark.parametrize("deg", [0.0, 30.0, 45.0, 80.0, 89.9])
def test_snell_angles(oracle, deg):
    th = math.radians(deg)
    d = (math.sin(th), -math.cos(th))           # travelling down onto the surface y = 0, normal (0, 1)
    refl, refr, R = oracle.refract(d, (0, 1), 1.0, 1.5)
    np.testing.assert_allclose(refl, (d[0], -d[1]), atol=1e-15)
    tt = math.asin(math.sin(th) / 1.5)
    np.testing.assert_allclose(refr, (math.sin(tt), -math.cos(tt)), atol=1e-15)
    rs = ((math.cos(th) - 1.5 * math.cos(tt)) / (math.cos(th) + 1.5 * math.cos(tt))) ** 2
    rp = ((math.cos(tt) - 1.5 * math.cos(th)) / (math.cos(tt) + 1.5 * math.cos(th))) ** 2
    assert abs(R - 0.5 * (rs + rp)) < 1e-15
    assert 0.0 <= R <= 1.0


def test_fresnel_normal_incidence_and_tir(oracle):
    for n1, n2 in [(1.0, 1.5), (1.5, 1.0), (1.0, 1.73), (1.33, 2.4)]:
        _, _, R = oracle.refract((0, -1), (0, 1), n1, n2)
        assert abs(R - ((n1 - n2) / (n1 + n2)) ** 2) < 1e-16
    crit = math.asin(1.0 / 1.5)
    for th, tir in [(crit - 1e-3, False), (crit + 1e-3, True), (math.radians(80), True)]:
        d = (math.sin(th), -math.cos(th))
        refl, refr, R = oracle.refract(d, (0, 1), 1.5, 1.0)
        assert (refr is None) == tir
        if tir:
            assert R == 1.0
    # the normal may point either way: it is oriented against the ray
    a = oracle.refract((0.6, -0.8), (0, 1), 1.0, 1.5)
    b = oracle.refract((0.6, -0.8), (0, -1), 1.0, 1.5)
    assert a == b


def test_reflect(oracle):
    np.testing.assert_allclose(oracle.reflect((0.6, -0.8), (0, 1)), (0.6, 0.8), atol=1e-16)
    np.testing.assert_allclose(oracle.reflect((0.6, -0.8), (0, -1)), (0.6, 0.8), atol=1e-16)


# ---- Tracer::trace control flow ------------------------------------------------------------------------
def test_two_parallel_mirrors_64_bounces(oracle):
    """Closed form: between mirrors x = -1 and x = +1 a ray with direction (cos a, sin a) from (0, 0) hits
    x = +-1 alternately, hit k at y = (2k - 1) tan a."""
    big = Rect.from_tlbr(100, -100, -100, 100)
    objs = [Object.new_mirror((-1, -50), (-1, 50)), Object.new_mirror((1, -50), (1, 50))]
    sc = scene(oracle, objs, max_bounce=64, canvas=big)
    a = 0.05
    res = sc.trace_rays(ray((0, 0), (math.cos(a), math.sin(a))))
    assert res.segments_emitted == 64 and res.ray_steps == 64
    for k in range(1, 65):
        end = res.f64["b"][k - 1]
        assert abs(end[0] - (1 if k % 2 else -1)) < 1e-12
        assert abs(end[1] - (2 * k - 1) * math.tan(a)) < 1e-10
        assert res.tags["hit_object"][k - 1] == (1 if k % 2 else 0)
        assert res.tags["generation"][k - 1] == k - 1
    assert np.all(res.seg["color"] == np.float32(0.5))          # mirrors never attenuate (tracer.rs:473-481)


def test_max_bounce_counts_generations(oracle):
    """max_bounce = n means the primary ray + n-1 bounces (tracer.rs:373)."""
    objs = [Object.new_mirror((-1, -50), (-1, 50)), Object.new_mirror((1, -50), (1, 50))]
    for mb in (0, 1, 2, 5):
        sc = scene(oracle, objs, max_bounce=mb, canvas=Rect.from_tlbr(100, -100, -100, 100))
        res = sc.trace_rays(ray((0, 0), (1, 0.01)))
        assert res.segments_emitted == mb


def test_cutoff_is_checked_at_pop(oracle):
    sc = scene(oracle, [Object.new_circle((2, 0), 0.5)])
    assert sc.trace_rays(ray((0, 0), (1, 0), color=(0.0009, 0.0009, 0.0009, 0.5))).segments_emitted == 0
    assert sc.trace_rays(ray((0, 0), (1, 0), color=(0.0009, 0.0011, 0.0009, 0.5))).segments_emitted > 0
    assert sc.trace_rays(ray((0, 0), (1, 0), color=(0.5, 0.5, 0.5, 0.0009))).segments_emitted == 0


def test_refractive_split_colours_and_media(oracle):
    """Normal incidence on a slab n = 1.5: reflected gets R, refracted 1-R, alpha untouched (tracer.rs:454-472)."""
    sc = scene(oracle, [Object.new_rect((2.0, 0.0), 1.0, 1.0).with_index(1.5)], max_bounce=3)
    res = sc.trace_rays(ray((0, 0), (1, 0), color=(0.5, 0.25, 0.125, 0.75)))
    R = np.float32(((1 - 1.5) / (1 + 1.5)) ** 2)
    seg, tag = res.seg, res.tags
    # generation 0: origin -> (1.5, 0)
    np.testing.assert_allclose(seg[0]["b"], (1.5, 0), atol=1e-7)
    # generation 1 in queue order: reflected first (path 0), then refracted (path 1)
    assert list(tag["generation"][:3]) == [0, 1, 1] and list(tag["path"][:3]) == [0, 0, 1]
    np.testing.assert_array_equal(seg[1]["color"], np.float32([0.5, 0.25, 0.125, 0.75]) * [R, R, R, 1])
    np.testing.assert_array_equal(seg[2]["color"][:3], np.float32([0.5, 0.25, 0.125]) * (np.float32(1) - R))
    assert tag["hit_object"][1] == -1                      # reflected ray leaves through the canvas
    assert tag["hit_object"][2] == 0                       # refracted ray hits the far face from inside
    np.testing.assert_allclose(seg[2]["b"], (2.5, 0), atol=1e-7)


def test_leaving_into_overlapping_object_uses_its_index(oracle):
    """tracer.rs:430-442: leaving object A at a point inside object B -> the refracted ray's medium is B's."""
    a = Object.new_circle((0.0, 0.0), 1.0).with_index(1.5)
    b = Object.new_circle((1.2, 0.0), 1.0).with_index(2.0)   # overlaps A around x in [0.2, 1]
    sc = scene(oracle, [a, b], max_bounce=2)
    res = sc.trace_rays(ray((0, 0), (1, 0), n=1.5))
    # nearest hit is B's near boundary at x = 0.2 (entering B): n2 = B's index
    assert res.tags["hit_object"][0] == 1 and abs(res.f64["b"][0][0] - 0.2) < 1e-12
    _, _, R = oracle.refract((1, 0), (-1, 0), 1.5, 2.0)
    refr = res.seg[res.tags["path"] == 1]
    np.testing.assert_allclose(refr["color"][0][0], np.float32(0.5) * (np.float32(1) - np.float32(R)), rtol=1e-7)
    # a ray that starts inside both and leaves A at (1, 0), which is inside B: medium becomes B's (2.0), so
    # with n1 = 2.0 as the current medium there is no index change: reflectance 0
    res = sc.trace_rays(ray((0.5, 0), (1, 0), n=2.0))
    assert res.tags["hit_object"][0] == 0
    kids = res.seg[res.tags["generation"] == 1]
    assert len(kids) == 1 and res.tags["path"][res.tags["generation"] == 1][0] == 1   # reflected child culled (R = 0)


def test_start_medium_last_match_wins(oracle):
    a = Object.new_circle((0.0, 0.0), 1.0).with_index(1.5)
    b = Object.new_circle((0.0, 0.0), 0.5).with_index(2.0)
    m = Object.new_mirror((-1, 0), (1, 0))
    sc = scene(oracle, [a, b, m])
    assert sc.start_medium(PointLight((0.1, 0.1), 1, (1, 1, 1, 1))) == 2.0     # tracer.rs:280-287, no break
    assert sc.start_medium(PointLight((0.7, 0.0), 1, (1, 1, 1, 1))) == 1.5
    assert sc.start_medium(PointLight((3.0, 0.0), 1, (1, 1, 1, 1))) == 1.0
    sc = scene(oracle, [b, a])
    assert sc.start_medium(PointLight((0.1, 0.1), 1, (1, 1, 1, 1))) == 1.5


def test_canvas_exit(oracle):
    sc = scene(oracle, [])
    res = sc.trace_rays(ray((0, 0), (1, 0)))
    assert res.segments_emitted == 1 and res.tags["hit_object"][0] == -1
    np.testing.assert_allclose(res.f64["b"][0], (16.0 / 9.0, 0), atol=1e-12)
    res = sc.trace_rays(ray((0.5, 0.5), (-1, -1)))
    np.testing.assert_allclose(res.f64["b"][0], (-1.0, -1.0), atol=1e-12)
    # origin outside, pointing in: the ray is drawn up to the FIRST (nearest) crossing and ends
    res = sc.trace_rays(ray((-3, 0), (1, 0)))
    np.testing.assert_allclose(res.f64["b"][0], (-16.0 / 9.0, 0), atol=1e-12)
    # origin outside, pointing away: nothing
    assert sc.trace_rays(ray((-3, 0), (-1, 0))).segments_emitted == 0


def test_lens_is_symmetric_about_its_axis(oracle):
    """Lens = Logic(And, circle(+d/2), circle(-d/2)) (object.rs:393-410): mirror-image rays give mirror-image paths."""
    lens = Object.new_lens((0.0, 0.0), 2.0, 3.8).with_index(1.5)
    sc = scene(oracle, [lens], max_bounce=4)
    up = sc.trace_rays(ray((-1, 0.3), (1, 0), color=(0.5, 0.5, 0.5, 0.5)))
    dn = sc.trace_rays(ray((-1, -0.3), (1, 0), color=(0.5, 0.5, 0.5, 0.5)))
    assert up.segments_emitted == dn.segments_emitted > 2
    np.testing.assert_allclose(up.f64["b"][:, 0], dn.f64["b"][:, 0], atol=1e-12)
    np.testing.assert_allclose(up.f64["b"][:, 1], -dn.f64["b"][:, 1], atol=1e-12)
    # a convex lens bends an off-axis ray towards the axis
    first_exit = up.f64["b"][up.tags["generation"] == 1]
    assert any(abs(e[1]) < 0.3 for e in first_exit)


# ---- CSG -----------------------------------------------------------------------------------------------------
def _rays(n, seed=1):
    rng = np.random.default_rng(seed)
    o = rng.uniform(-2, 2, (n, 2))
    a = rng.uniform(0, 2 * math.pi, n)
    return o, np.stack([np.cos(a), np.sin(a)], 1)


def test_csg_identities(oracle):
    """A∧A = A, A∨A = A, A∧¬A = ∅ as point sets (contains)."""
    A = Circle((0.2, 0.1), 0.7)
    plain = scene(oracle, [Object.new_geo(A)])
    a_and_a = scene(oracle, [Object.new_geo(Logic(AND, A, A))])
    a_or_a = scene(oracle, [Object.new_geo(Logic(OR, A, A))])
    a_not_a = scene(oracle, [Object.new_geo(Logic(AND_NOT, A, A))])
    pts, _ = _rays(600)
    inside = 0
    for p in pts:
        c = plain.contains(0, p)
        inside += c
        assert a_and_a.contains(0, p) == c
        assert a_or_a.contains(0, p) == c
        assert not a_not_a.contains(0, p)
    assert 20 < inside < 580


def test_csg_hit_sets(oracle):
    A, B = Circle((-0.3, 0.0), 0.6), Circle((0.3, 0.0), 0.6)
    sa, sb = scene(oracle, [Object.new_geo(A)]), scene(oracle, [Object.new_geo(B)])
    s_and = scene(oracle, [Object.new_geo(Logic(AND, A, B))])
    s_or = scene(oracle, [Object.new_geo(Logic(OR, A, B))])
    s_not = scene(oracle, [Object.new_geo(Logic(AND_NOT, A, B))])
    o, d = _rays(500, seed=2)
    for i in range(len(o)):
        ha, hb = sa.intersect(0, o[i], d[i]), sb.intersect(0, o[i], d[i])
        exp_and = [h for h in ha if sb.contains(0, h[:2])] + [h for h in hb if sa.contains(0, h[:2])]
        exp_or = [h for h in ha if not sb.contains(0, h[:2])] + [h for h in hb if not sa.contains(0, h[:2])]
        exp_not = [h for h in ha if not sb.contains(0, h[:2])] + [h for h in hb if sa.contains(0, h[:2])]
        for sc, exp in ((s_and, exp_and), (s_or, exp_or), (s_not, exp_not)):
            got = sc.intersect(0, o[i], d[i])
            assert got.shape[0] == len(exp)
            if len(exp):
                np.testing.assert_array_equal(got[:, :2], np.array(exp)[:, :2])
        # AndNot flips the normals of B's surviving hits (the carved-out boundary faces the other way)
        got = s_not.intersect(0, o[i], d[i])
        nb = [h for h in hb if sa.contains(0, h[:2])]
        if nb:
            np.testing.assert_array_equal(got[-len(nb):, 2:4], -np.array(nb)[:, 2:4])


def _contains_recursive(geo, p):
    """Independent evaluation in LOCAL frames (the way a Rust library would recurse)."""
    if isinstance(geo, Circle):
        return (p[0] - geo.origin[0]) ** 2 + (p[1] - geo.origin[1]) ** 2 < geo.radius ** 2
    if isinstance(geo, Rect):
        c, s = geo.rotation[0], geo.rotation[1]
        x, y = p[0] - geo.origin[0], p[1] - geo.origin[1]
        lx, ly = c * x + s * y, -s * x + c * y
        return abs(lx) < geo.width / 2 and abs(ly) < geo.height / 2
    if isinstance(geo, Logic):
        c, s = geo.rotation[0], geo.rotation[1]
        x, y = p[0] - geo.origin[0], p[1] - geo.origin[1]
        q = (c * x + s * y, -s * x + c * y)
        a, b = _contains_recursive(geo.a, q), _contains_recursive(geo.b, q)
        return (a and b) if geo.op == AND else (a or b) if geo.op == OR else (a and not b)
    return False


def test_lowering_to_world_space_equals_local_frame_recursion(oracle):
    inner = Logic(AND_NOT, Rect((0.1, 0.0), rot2(0.3), 0.9, 0.5), Circle((0.2, 0.1), 0.25), (0.1, -0.2), rot2(-0.8))
    outer = Logic(OR, inner, Logic(AND, Circle((0.3, 0), 0.5), Circle((-0.3, 0), 0.5), (-0.4, 0.3), rot2(1.1)),
                  (0.25, 0.15), rot2(2.0))
    sc = scene(oracle, [Object.new_geo(outer)])
    rng = np.random.default_rng(3)
    pts = rng.uniform(-1.5, 1.5, (4000, 2))
    got = np.array([sc.contains(0, p) for p in pts])
    exp = np.array([_contains_recursive(outer, p) for p in pts])
    assert got.sum() > 200 and (~got).sum() > 200
    assert np.array_equal(got, exp)
    # every reported hit lies on the boundary of the solid: points just either side differ in containment
    o, d = _rays(300, seed=4)
    n_hits = 0
    for i in range(len(o)):
        for h in sc.intersect(0, o[i] * 0.75, d[i]):
            n_hits += 1
            inside = _contains_recursive(outer, h[:2] - 1e-6 * h[2:4])
            outside = _contains_recursive(outer, h[:2] + 1e-6 * h[2:4])
            assert inside != outside
    assert n_hits > 100


# ---- lights --------------------------------------------------------------------------------------------------------
def test_point_light_directions(oracle):
    n = 1000
    rays = oracle.emit_rays(PointLight((0.3, -0.2), n, (1, 1, 1, 1)))
    i = np.arange(n)
    f = i * math.pi * 2.0 / n
    np.testing.assert_allclose(rays["direction"][:, 0], np.cos(f), atol=2e-16)
    np.testing.assert_allclose(rays["direction"][:, 1], np.sin(f), atol=2e-16)
    assert np.all(rays["origin"] == (0.3, -0.2))


def test_spot_light_cone(oracle):
    n = 100
    l = SpotLight((1.0, 0.0), math.radians(10), (-1.0, 0.0), n, (1, 1, 1, 1))
    rays = oracle.emit_rays(l)
    ang = np.arctan2(rays["direction"][:, 1], rays["direction"][:, 0])
    ang = np.where(ang < 0, ang + 2 * math.pi, ang)
    # centred on pi (pointing -x), total aperture 10 degrees, `for step in 1..=num_rays` (light.rs:241)
    assert abs(ang.max() - (math.pi + math.radians(5))) < 1e-12
    assert abs(ang.min() - (math.pi - math.radians(5) + math.radians(10) / n)) < 1e-12
    up = SpotLight((0, 0), 0.2, (0.0, 1.0), 3, (1, 1, 1, 1))   # |dx| < EPSILON branch (light.rs:230-236)
    r = oracle.emit_rays(up)
    assert np.all(r["direction"][:, 1] > 0.99)


def test_directional_light(oracle):
    l = DirectionalLight((1, 1, 1, 1), 4, LineSegment((0.0, 0.0), (2.0, 0.0)))
    r = oracle.emit_rays(l)
    np.testing.assert_allclose(r["origin"][:, 0], [0, 0.5, 1.0, 1.5])
    np.testing.assert_allclose(r["direction"], [[0, 1]] * 4, atol=1e-16)


# ---- string mod -------------------------------------------------------------------------------------------------------
def test_string_mod_5_2_mul_exact(oracle):
    sm = StringMod(modulo=5, num=2, mode=StringModMode.Mul)
    ch = oracle.string_mod(sm)
    tgt = [0, 2, 4, 1, 3]
    for i in range(5):
        np.testing.assert_allclose(ch["a"][i], (math.cos(i * math.tau / 5), math.sin(i * math.tau / 5)), atol=1e-15)
        j = tgt[i]
        np.testing.assert_allclose(ch["b"][i], (math.cos(j * math.tau / 5), math.sin(j * math.tau / 5)), atol=1e-15)
    assert np.all(ch["color_a"] == 1.0) and np.all(ch["color_b"] == 1.0)


def test_string_mod_modes_and_colours(oracle):
    m = 97
    pts = lambda k: (math.cos(k * math.tau / m), math.sin(k * math.tau / m))
    for mode, f in ((StringModMode.Add, lambda i: (i + 7) % m), (StringModMode.Mul, lambda i: (i * 7) % m),
                    (StringModMode.Pow, lambda i: pow(i, 7, 1 << 64) % m),
                    (StringModMode.Base, lambda i: pow(7, i, 1 << 64) % m)):
        ch = oracle.string_mod(StringMod(modulo=m, num=7, mode=mode))
        for i in range(m):
            np.testing.assert_allclose(ch["b"][i], pts(f(i)), atol=1e-15)
    rules = [ModRemColor(3, 0, (1, 0, 0, 1)), ModRemColor(2, 0, (0, 0, 1, 1))]
    ch = oracle.string_mod(StringMod(modulo=12, num=1, mode=StringModMode.Add, color=(0.5, 0.5, 0.5, 0.5),
                                     modulo_colors=rules))
    np.testing.assert_array_equal(ch["color_a"][0], (0.5, 0, 0.5, 1))     # 0 matches both rules: average
    np.testing.assert_array_equal(ch["color_a"][3], (1, 0, 0, 1))
    np.testing.assert_array_equal(ch["color_a"][2], (0, 0, 1, 1))
    np.testing.assert_array_equal(ch["color_a"][1], (0.5, 0.5, 0.5, 0.5))  # no rule: base colour
    np.testing.assert_array_equal(ch["color_b"][0], ch["color_a"][1])


# ---- accumulate --------------------------------------------------------------------------------------------------------
def _pair(a, b, ca=(1, 1, 1, 1), cb=None):
    p = np.zeros(1, dtype=abi.VERTEX_PAIR_DTYPE)
    p["a"], p["b"], p["color_a"], p["color_b"] = a, b, ca, (cb or ca)
    return p


def _world(px, py, W, H):
    a = W / H
    return ((px / (W / 2) - 1) * a, 1 - py / (H / 2))


def test_dda_coverage_counts(oracle):
    W, H = 64, 32
    img = oracle.new_image(W, H)
    assert np.all(img[..., 3] == 1) and np.all(img[..., :3] == 0)          # LoadOp::Clear(BLACK)
    # horizontal: pixel centres 10.5 .. 19.5 inside [10.2, 20.2) -> 10 fragments on row 5
    n = oracle.accumulate_pairs(img, _pair(_world(10.2, 5.5, W, H), _world(20.2, 5.5, W, H), (0.25, 0.5, 1, 0.5)))
    assert n == 10
    assert np.all(img[5, 10:20, 0] == 0.25) and img[5, 9, 0] == 0 and img[5, 20, 0] == 0
    assert np.all(img[5, 10:20, 3] == 1 + 0.25)                             # a += src.a * src.a
    # vertical
    img = oracle.new_image(W, H)
    assert oracle.accumulate_pairs(img, _pair(_world(7.5, 3.0, W, H), _world(7.5, 13.0, W, H))) == 10
    assert np.all(img[3:13, 7, 1] == 1)
    # 45 degrees: one fragment per column, stepping one row each
    img = oracle.new_image(W, H)
    assert oracle.accumulate_pairs(img, _pair(_world(2.0, 2.0, W, H), _world(12.0, 12.0, W, H))) == 10
    assert all(img[2 + k, 2 + k, 0] == 1 for k in range(10))
    # clipped: only the on-canvas part, and a fully off-canvas segment draws nothing
    img = oracle.new_image(W, H)
    assert oracle.accumulate_pairs(img, _pair(_world(-20.0, 8.5, W, H), _world(5.0, 8.5, W, H))) == 5
    assert oracle.accumulate_pairs(img, _pair(_world(-20.0, 8.5, W, H), _world(-5.0, 8.5, W, H))) == 0
    assert oracle.accumulate_pairs(img, _pair((0, 0), (0, 0))) == 0          # zero length


def test_colour_is_lerped_between_endpoints(oracle):
    W, H = 64, 32
    img = oracle.new_image(W, H)
    oracle.accumulate_pairs(img, _pair(_world(0.0, 4.5, W, H), _world(64.0, 4.5, W, H), (0, 0, 0, 0), (1, 0.5, 0, 0)))
    np.testing.assert_allclose(img[4, :, 0], (np.arange(64) + 0.5) / 64, rtol=1e-5)
    np.testing.assert_allclose(img[4, :, 1], (np.arange(64) + 0.5) / 128, rtol=1e-5)


def test_accumulate_is_additive_and_order_free(oracle):
    W, H = 96, 54
    rng = np.random.default_rng(5)
    n = 400
    p = np.zeros(n, dtype=abi.VERTEX_PAIR_DTYPE)
    p["a"] = rng.uniform(-1.9, 1.9, (n, 2))
    p["b"] = rng.uniform(-1.9, 1.9, (n, 2))
    # power-of-two colours: every partial sum is exact in fp32, so any order gives the same bits
    p["color_a"] = p["color_b"] = 2.0 ** -rng.integers(4, 9, (n, 4))
    a, b = oracle.new_image(W, H), oracle.new_image(W, H)
    na = oracle.accumulate_pairs(a, p)
    nb = oracle.accumulate_pairs(b, p[rng.permutation(n)])
    assert na == nb and np.array_equal(a, b)
    c = oracle.new_image(W, H)
    oracle.accumulate_pairs(c, p[: n // 2])
    oracle.accumulate_pairs(c, p[n // 2:])
    assert np.array_equal(a, c)
    # banded multi-thread accumulation == single thread
    d = oracle.new_image(W, H)
    assert oracle.accumulate_pairs(d, p, threads=1) == na and np.array_equal(a, d)


def test_f16_conversion_matches_numpy(oracle):
    rng = np.random.default_rng(6)
    x = np.concatenate([rng.uniform(-70000, 70000, 5000), rng.uniform(-1e-4, 1e-4, 5000), [0, 1, 65504, 65520, 1e-8,
                        2.0 ** -24, 2.0 ** -25, 3 * 2.0 ** -25]]).astype(np.float32).reshape(-1, 4)
    with np.errstate(over="ignore"):
        assert np.array_equal(oracle.to_f16(x).view(np.uint16), x.astype(np.float16).view(np.uint16))


def test_screenshot_conversion_known_values(oracle):
    """renderer.rs:325-328: u8 = (f^(1/2.2) * 255) as u8 on the fp16 value; pixel order [b, g, r, a]."""
    img = np.zeros((1, 6, 4), dtype=np.float32)
    img[0, :, 0] = [0.0, 1.0, 0.5, 2.0, -1.0, 0.25]
    img[0, :, 1] = 1.0
    img[0, :, 3] = 0.0
    out = oracle.to_bgra8(img)
    exp_r = [0, 255, int(0.5 ** (1 / 2.2) * 255), 255, 0, int(0.25 ** (1 / 2.2) * 255)]   # saturates; NaN -> 0
    assert out[0, :, 2].tolist() == exp_r          # red lands in byte 2
    assert np.all(out[0, :, 1] == 255) and np.all(out[0, :, 0] == 0) and np.all(out[0, :, 3] == 0)


def test_surface_srgb_conversion_known_values(oracle):
    """ORACLE.md 8.7 (sub_render_pass.rs:59-63, renderer.rs:207-209): saturate, sRGB-encode, round to nearest;
    alpha linear; pixel order [b, g, r, a].  Checked against the closed-form sRGB encoding."""
    img = np.zeros((1, 8, 4), dtype=np.float32)
    img[0, :, 0] = [0.0, 1.0, 0.5, 7.0, -1.0, 0.0031308, 0.21404114, np.nan]
    img[0, :, 1] = 1.0
    img[0, :, 3] = [0.0, 1.0, 0.5, 3.0, -2.0, 0.25, 0.001, np.nan]
    out = oracle.to_bgra8_srgb(img)
    #        0    1    .5: 1.055*.5^(1/2.4)-.055 = .7354 -> 187.5.. -> 188; >1 saturates; <0 -> 0; 12.92*.0031308*255 = 10.3
    assert out[0, :, 2].tolist() == [0, 255, 188, 255, 0, 10, 128, 0]          # red lands in byte 2
    assert np.all(out[0, :, 1] == 255) and np.all(out[0, :, 0] == 0)
    assert out[0, :, 3].tolist() == [0, 255, 128, 255, 0, 64, 0, 0]            # alpha: linear, floor(a*255 + .5)
    rng = np.random.default_rng(3)
    x = np.concatenate([rng.uniform(0, 1.2, 20000), rng.uniform(0, 0.01, 4000)]).astype(np.float32)
    im = np.zeros((1, x.size, 4), dtype=np.float32)
    im[0, :, 0] = im[0, :, 1] = im[0, :, 2] = x
    got = oracle.to_bgra8_srgb(im)[0, :, 0].astype(np.int64)
    c = np.clip(x.astype(np.float64), 0, 1)
    enc = np.where(c <= 0.0031308, 12.92 * c, 1.055 * c ** (1 / 2.4) - 0.055)
    exp = np.floor(enc * 255 + 0.5).astype(np.int64)
    bad = got != exp
    # the two forms can only disagree within rounding of a threshold (0.0031308 is itself a rounded constant)
    frac = np.abs(enc * 255 + 0.5 - np.round(enc * 255 + 0.5))
    assert bad.mean() < 1e-3 and np.all(np.abs(got - exp)[bad] == 1) and np.all(frac[bad] < 1e-3)


def test_string_mod_curves(oracle):
    """string_mod.rs:36-84: ComplexExp, Hypotrochoid and Lissajous point sets against direct formulas."""
    m = 60
    c = complex(0.995, 0.03)
    ch = oracle.string_mod(StringMod(modulo=m, num=1, mode=StringModMode.Add, turns=2, init_curve=Curve.ComplexExp(c)))
    for n in range(m):
        z = c ** (2 * n)
        np.testing.assert_allclose(ch["a"][n], (z.real, z.imag), rtol=1e-12, atol=1e-14)
    r, s_, d = 3, 7, 2
    ch = oracle.string_mod(StringMod(modulo=m, num=1, mode=StringModMode.Add, init_curve=Curve.Hypotrochoid(r, s_, d)))
    for n in range(m):
        ang = n * math.tau / m
        smr = s_ - r
        x = smr * math.cos(ang) + d * math.cos(ang * smr / r)
        y = smr * math.sin(ang) - d * math.sin(ang * smr / r)
        np.testing.assert_allclose(ch["a"][n], (x / (smr + d), y / (smr + d)), atol=1e-15)
    ch = oracle.string_mod(StringMod(modulo=m, num=1, mode=StringModMode.Add, init_curve=Curve.Lissajous(3, 2, 0.5)))
    for n in range(m):
        ang = n * math.tau / m
        np.testing.assert_allclose(ch["a"][n], (math.sin(3 * ang + 0.5), math.sin(2 * ang)), atol=1e-15)
    assert np.array_equal(ch["b"][:-1], ch["a"][1:])          # Add 1: chord i -> i + 1


def test_ray_ellipse(oracle):
    """Ellipse{origin, a, b, rot} (object.rs:38-45): x^2/a^2 + y^2/b^2 = 1 in the rotated frame."""
    sc = scene(oracle, [Object.new_ellipse((1.0, 0.5), 2.0, 1.0)])
    h = sc.intersect(0, (-5, 0.5), (1, 0))
    np.testing.assert_allclose(h[:, 0], [-1.0, 3.0], atol=1e-14)
    np.testing.assert_allclose(h[:, 2:4], [[-1, 0], [1, 0]], atol=1e-14)
    h = sc.intersect(0, (1.0, -5), (0, 1))
    np.testing.assert_allclose(h[:, 1], [-0.5, 1.5], atol=1e-14)
    # a generic point of the ellipse: (a cos t, b sin t); the normal is (cos t / a, sin t / b) normalised
    t = 0.7
    p = (1.0 + 2.0 * math.cos(t), 0.5 + 1.0 * math.sin(t))
    h = sc.intersect(0, (1.0, 0.5), (p[0] - 1.0, p[1] - 0.5))
    dirv = np.array([p[0] - 1.0, p[1] - 0.5])
    dirv /= np.linalg.norm(dirv)
    h = sc.intersect(0, (1.0, 0.5), dirv)
    assert h.shape[0] == 1
    np.testing.assert_allclose(h[0, :2], p, atol=1e-14)
    nrm = np.array([math.cos(t) / 2.0, math.sin(t) / 1.0])
    np.testing.assert_allclose(h[0, 2:4], nrm / np.linalg.norm(nrm), atol=1e-14)
    assert sc.contains(0, (2.9, 0.5)) and not sc.contains(0, (3.1, 0.5)) and not sc.contains(0, (1.0, 1.6))
    # rotated by 90 degrees the axes swap
    rot = Object(Ellipse((0.0, 0.0), 2.0, 1.0, rot2(math.pi / 2)), Material(1.5), "Ellipse")
    sc = scene(oracle, [rot])
    h = sc.intersect(0, (-5, 0), (1, 0))
    np.testing.assert_allclose(h[:, 0], [-1.0, 1.0], atol=1e-12)
    h = sc.intersect(0, (0, -5), (0, 1))
    np.testing.assert_allclose(h[:, 1], [-2.0, 2.0], atol=1e-12)
    # a circle is the ellipse with a = b: same hit points as the circle primitive
    e = scene(oracle, [Object.new_ellipse((0.3, -0.2), 0.7, 0.7)])
    c = scene(oracle, [Object.new_circle((0.3, -0.2), 0.7)])
    o, d = _rays(200, seed=9)
    for i in range(len(o)):
        he, hc = e.intersect(0, o[i], d[i]), c.intersect(0, o[i], d[i])
        assert he.shape == hc.shape
        if len(he):
            np.testing.assert_allclose(he[:, :4], hc[:, :4], atol=1e-12)


def test_convex_hull_and_polygon(oracle):
    """ConvexPolygon::new_convex_hull (object.rs:34-36) and ORACLE.md §3.8: edges in hull order as §3.2 segments,
    contains = strictly inside every edge."""
    from light_garden_b200.scene import ConvexPolygon, convex_hull
    # hull: interior and collinear points dropped, counter-clockwise from the lowest (x, y) point
    pts = [(1, 1), (0, 0), (2, 0), (1, 0), (2, 2), (0, 2), (1, 0.5), (0, 1)]
    assert convex_hull(pts) == [(0.0, 0.0), (2.0, 0.0), (2.0, 2.0), (0.0, 2.0)]
    with pytest.raises(ValueError):
        convex_hull([(0, 0), (1, 1), (2, 2)])
    sq = Object.new_convex_polygon(pts)
    assert sq.material_opt is not None and sq.kind == "ConvexPolygon"
    sc = scene(oracle, [sq])
    # hits are listed in edge order: bottom (0), right (1), top (2), left (3)
    h = sc.intersect(0, (-3, 0.5), (1, 0))
    np.testing.assert_allclose(h[:, :2], [[2.0, 0.5], [0.0, 0.5]], atol=1e-15)
    np.testing.assert_allclose(np.abs(h[:, 2:4]), [[1, 0], [1, 0]], atol=1e-15)
    h = sc.intersect(0, (0.5, -3), (0, 1))
    np.testing.assert_allclose(h[:, :2], [[0.5, 0.0], [0.5, 2.0]], atol=1e-15)
    assert sc.intersect(0, (-3, 2.5), (1, 0)).shape[0] == 0
    assert sc.contains(0, (1.0, 1.0)) and sc.contains(0, (1.99, 0.01))
    assert not sc.contains(0, (2.0, 1.0)) and not sc.contains(0, (1.0, -0.01))   # the boundary is outside
    # the same square as a Rect: identical hit points and containment on random rays
    rc = scene(oracle, [Object.new_rect((1.0, 1.0), 2.0, 2.0)])
    o, d = _rays(300, seed=5)
    for i in range(len(o)):
        hp, hr = sc.intersect(0, o[i] * 2, d[i]), rc.intersect(0, o[i] * 2, d[i])
        assert hp.shape == hr.shape
        if len(hp):
            key = lambda a: a[np.lexsort((a[:, 1], a[:, 0]))]
            np.testing.assert_allclose(key(hp)[:, :2], key(hr)[:, :2], atol=1e-12)
        assert sc.contains(0, o[i]) == rc.contains(0, o[i])
    # clockwise input order, own frame: a triangle rotated by 90 degrees about its origin (1, 0)
    tri = Object(ConvexPolygon(((0.0, 0.0), (0.0, 1.0), (2.0, 0.0)), (1.0, 0.0), rot2(math.pi / 2)), Material(1.5))
    sc = scene(oracle, [tri])          # world vertices (1, 0), (0, 0), (1, 2)
    assert sc.contains(0, (0.8, 0.3)) and not sc.contains(0, (0.2, 1.0)) and not sc.contains(0, (1.1, 0.5))
    h = sc.intersect(0, (-1, 0.5), (1, 0))
    np.testing.assert_allclose(sorted(h[:, 0]), [0.25, 1.0], atol=1e-12)
    # 32 vertices approximate a circle: hit distances within the sagitta of the polygon
    ring = [(math.cos(2 * math.pi * k / 32), math.sin(2 * math.pi * k / 32)) for k in range(32)]
    sc = scene(oracle, [Object.new_convex_polygon(ring)])
    for ang in np.linspace(0.05, 6.2, 23):
        h = sc.intersect(0, (0, 0), (math.cos(ang), math.sin(ang)))
        assert h.shape[0] == 1 and math.cos(math.pi / 32) - 1e-12 <= h[0, 4] <= 1.0 + 1e-12
    with pytest.raises(ValueError):
        scene(oracle, [Object.new_convex_polygon([(math.cos(k), math.sin(k)) for k in range(40)])])


def test_prism_deviates_a_ray_by_the_textbook_angle(oracle):
    """An equilateral n = 1.5 prism at minimum deviation: delta = 2 asin(n sin(A / 2)) - A with A = 60 degrees."""
    A, n = math.radians(60.0), 1.5
    s = 1.0
    prism = Object.new_convex_polygon([(-s / 2, 0.0), (s / 2, 0.0), (0.0, s * math.sin(A))]).with_index(n)
    sc = scene(oracle, [prism])
    # minimum deviation: the ray inside runs parallel to the base; outside angle of incidence i = asin(n sin(A/2))
    i = math.asin(n * math.sin(A / 2))
    # left face goes from (-0.5, 0) to (0, 0.866): outward normal points up-left at 150 degrees
    nl = (math.cos(math.radians(150)), math.sin(math.radians(150)))
    hit = (-0.25, 0.5 * s * math.sin(A))                      # midpoint of the left face
    # incoming direction = -normal rotated by i (towards the base side)
    a = math.atan2(-nl[1], -nl[0]) + i
    d_in = (math.cos(a), math.sin(a))
    o = (hit[0] - 0.5 * d_in[0], hit[1] - 0.5 * d_in[1])
    res = sc.trace_rays(ray(o, d_in, color=(1, 1, 1, 1)), abi.LG_PRECISION_F64)
    seg, tag = res.seg, res.tags
    # follow the refracted path: generation 1 segment that is inside (path bit 1), then generation 2 leaving
    inside = [k for k in range(len(seg)) if tag["generation"][k] == 1 and tag["path"][k] == 1]
    assert len(inside) == 1
    k = inside[0]
    v = np.array(seg["b"][k], dtype=np.float64) - np.array(seg["a"][k], dtype=np.float64)
    assert abs(v[1]) < 1e-6 * abs(v[0]) and v[0] > 0               # parallel to the base
    out = [k for k in range(len(seg)) if tag["generation"][k] == 2 and tag["path"][k] == 3]
    assert len(out) == 1
    w = np.array(seg["b"][out[0]], dtype=np.float64) - np.array(seg["a"][out[0]], dtype=np.float64)
    dev = math.atan2(d_in[1], d_in[0]) - math.atan2(w[1], w[0])
    assert abs(dev - (2 * i - A)) < 1e-6


def test_nested_string_mod_crossings(oracle):
    """StringMod with nested = Some(inner) (string_mod.rs:87-101,152-158): crossings of the outer chords in the
    reference's loop order (diff = 1..L-1, ixa = 0..L-1), then the inner pattern's chords between those points."""
    # a pentagram (5 points, i -> i + 2): its 5 inner crossings lie on the circle of radius 1 / phi^2
    outer = StringMod(modulo=5, num=2, mode=StringModMode.Add)
    lines = oracle.string_mod(outer)
    pts = oracle.line_crossings(lines)
    rad = np.hypot(pts[:, 0], pts[:, 1])
    inner_r = 1.0 / ((1 + math.sqrt(5)) / 2) ** 2
    assert (np.abs(rad - inner_r) < 1e-12).sum() == 10       # every unordered pair is visited twice (diff and L - diff)
    assert ((np.abs(rad - inner_r) < 1e-12) | (np.abs(rad - 1.0) < 1e-12)).all()   # the rest are shared end points
    # order and values against a plain numpy restatement of the loop nest on random segments
    rng = np.random.default_rng(11)
    L = 40
    segs = np.zeros(L, dtype=abi.VERTEX_PAIR_DTYPE)
    segs["a"], segs["b"] = rng.uniform(-1, 1, (L, 2)), rng.uniform(-1, 1, (L, 2))
    exp = []
    for diff in range(1, L):
        for i in range(L):
            j = (i + diff) % L
            a1, e1 = segs["a"][i], segs["b"][i] - segs["a"][i]
            a2, e2 = segs["a"][j], segs["b"][j] - segs["a"][j]
            den = e1[0] * e2[1] - e1[1] * e2[0]
            w = a2 - a1
            t, u = (w[0] * e2[1] - w[1] * e2[0]) / den, (w[0] * e1[1] - w[1] * e1[0]) / den
            if 0 <= t <= 1 and 0 <= u <= 1:
                exp.append(a1 + t * e1)
    got = oracle.line_crossings(segs)
    assert len(got) == len(exp) > 100
    np.testing.assert_allclose(got, np.array(exp), atol=1e-13)
    # the inner pattern indexes the crossing points modulo their number and colours by its own rules
    k = 2.0 ** -6
    inner = StringMod(modulo=7, num=3, mode=StringModMode.Mul, color=(k, k, k, k), modulo_colors=[ModRemColor(2, 0, (k, 0, 0, k))])
    outer.nested = inner
    ch = oracle.string_mod_draw(outer)
    assert len(ch) == 7
    P = len(pts)
    for i in range(7):
        j = (i * 3) % 7
        np.testing.assert_array_equal(ch["a"][i], pts[i % P])
        np.testing.assert_array_equal(ch["b"][i], pts[j % P])
        np.testing.assert_array_equal(ch["color_a"][i], (k, 0, 0, k) if i % 2 == 0 else (k, k, k, k))
        np.testing.assert_array_equal(ch["color_b"][i], (k, 0, 0, k) if j % 2 == 0 else (k, k, k, k))
    # no crossings -> nothing is drawn (string_mod.rs:106-108)
    assert len(oracle.string_mod_draw(StringMod(modulo=1, nested=inner))) == 0


def test_blend_states(oracle):
    """gui/settings.rs:59-127: the blend equation per component, op(src * src_factor, dst * dst_factor), on two
    crossing lines whose values are known in every pixel."""
    W = H = 16
    c1, c2 = (0.5, 0.25, 0.125, 0.5), (0.25, 0.5, 1.0, 0.25)
    horiz = _pair(_world(0.0, 8.5, W, H), _world(16.0, 8.5, W, H), c1)     # row 8
    vert = _pair(_world(8.5, 0.0, W, H), _world(8.5, 16.0, W, H), c2)      # column 8
    pairs = np.concatenate([horiz, vert])
    ONE, ADD = abi.LG_BF_ONE, abi.LG_BO_ADD

    def run(color, alpha, constant=(0, 0, 0, 0)):
        img = oracle.new_image(W, H)
        assert oracle.accumulate_pairs_blend(img, pairs, color, alpha, constant) == 32
        return img

    # the default state through the generic path equals the dedicated one
    img = run((ONE, ONE, ADD), (abi.LG_BF_SRC_ALPHA, ONE, ADD))
    ref = oracle.new_image(W, H)
    oracle.accumulate_pairs(ref, pairs)
    assert np.array_equal(img, ref)
    np.testing.assert_array_equal(img[8, 8], (0.75, 0.75, 1.125, 1 + 0.25 + 0.0625))
    # Max: the brighter of the two lines where they cross; the clear alpha 1 stays
    img = run((ONE, ONE, abi.LG_BO_MAX), (ONE, ONE, abi.LG_BO_MAX))
    np.testing.assert_array_equal(img[8, 8], (0.5, 0.5, 1.0, 1.0))
    np.testing.assert_array_equal(img[8, 3], (0.5, 0.25, 0.125, 1.0))
    # Min against the cleared image: black stays black, alpha drops to the smallest source alpha
    img = run((ONE, ONE, abi.LG_BO_MIN), (ONE, ONE, abi.LG_BO_MIN))
    np.testing.assert_array_equal(img[8, 8], (0, 0, 0, 0.25))
    np.testing.assert_array_equal(img[0, 0], (0, 0, 0, 1))
    # Add with SrcAlpha on the colour and a constant on alpha
    img = run((abi.LG_BF_SRC_ALPHA, ONE, ADD), (abi.LG_BF_CONSTANT, ONE, ADD), constant=(0, 0, 0, 2.0))
    np.testing.assert_array_equal(img[8, 8], (0.5 * 0.5 + 0.25 * 0.25, 0.25 * 0.5 + 0.5 * 0.25, 0.125 * 0.5 + 1.0 * 0.25,
                                             1 + 0.5 * 2 + 0.25 * 2))
    # ReverseSubtract: the lines darken the image; OneMinusSrc as the factor
    img = run((abi.LG_BF_ONE_MINUS_SRC, ONE, abi.LG_BO_REVERSE_SUBTRACT), (abi.LG_BF_ZERO, ONE, ADD))
    np.testing.assert_array_equal(img[8, 3], (-(0.5 * 0.5), -(0.25 * 0.75), -(0.125 * 0.875), 1.0))


def test_directional_light_negative_parameter_switch(oracle):
    """light.rs:111 passes -(i as f64)/n to LineSegment::eval_at_r, whose definition is in collision2d (not vendored).
    neg_r = False (default): origins a + (i/n)(b - a), on the drawn segment; neg_r = True (LG_LIGHT_DIRECTIONAL_NEG_R):
    the call taken literally on eval_at_r(r) = a + r (b - a).  Same direction (the left normal of b - a) either way."""
    seg = LineSegment((1.0, -1.0), (3.0, -1.0))
    fwd = oracle.emit_rays(DirectionalLight((1, 1, 1, 1), 4, seg))
    neg = oracle.emit_rays(DirectionalLight((1, 1, 1, 1), 4, seg, neg_r=True))
    np.testing.assert_array_equal(fwd["origin"][:, 0], [1.0, 1.5, 2.0, 2.5])
    np.testing.assert_array_equal(neg["origin"][:, 0], [1.0, 0.5, 0.0, -0.5])
    np.testing.assert_array_equal(fwd["origin"][:, 1], neg["origin"][:, 1])
    np.testing.assert_array_equal(fwd["direction"], neg["direction"])
