"""Host-side scene model: the names and constructors of the reference's
src/light_garden/object.rs, light.rs and string_mod.rs (and of the collision2d
types they wrap), flattened to the PODs of include/light_garden_b200.h.

Only data lives here.  Nothing in this module intersects, traces or
rasterises anything: that is the CUDA library's job.
"""
import ctypes
import math
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import abi

P2 = Tuple[float, float]
Color = Tuple[float, float, float, float]  # light.rs:7  pub type Color = [f32; 4]


def rot2_identity():
    return (1.0, 0.0, 0.0, 1.0)


def rot2(angle: float):
    """nalgebra Rotation2 as serde writes it, column-major [m11, m21, m12, m22] (default.ron:24-29)."""
    c, s = math.cos(angle), math.sin(angle)
    return (c, s, -s, c)


# ---- collision2d Geo variants used by the BASELINE configs ---------------------------------
@dataclass
class Circle:
    origin: P2
    radius: float


@dataclass
class Rect:
    origin: P2
    rotation: Tuple[float, float, float, float]
    width: float
    height: float

    @staticmethod
    def new(origin, rotation, width, height):
        return Rect(tuple(origin), tuple(rotation), width, height)

    @staticmethod
    def from_tlbr(top, left, bottom, right):
        """Rect::from_tlbr(top, left, bottom, right) — argument order of sub_render_pass.rs:156."""
        return Rect(((left + right) * 0.5, (top + bottom) * 0.5), rot2_identity(), right - left, top - bottom)

    def tlbr(self):
        hw, hh = self.width * 0.5, self.height * 0.5
        return (self.origin[1] + hh, self.origin[0] - hw, self.origin[1] - hh, self.origin[0] + hw)


@dataclass
class Ellipse:
    """collision2d Ellipse{origin, a, b, rot} (object.rs:38-45): x^2/a^2 + y^2/b^2 = 1 in the rotated local frame."""
    origin: P2
    a: float
    b: float
    rot: Tuple[float, float, float, float] = (1.0, 0.0, 0.0, 1.0)


def convex_hull(points: Sequence[P2]):
    """ConvexPolygon::new_convex_hull (object.rs:34-36).  collision2d's own hull routine is not available offline;
    ORACLE.md §3.8 fixes it as Andrew's monotone chain: counter-clockwise, starting at the lowest (x, then y) point,
    collinear points dropped."""
    pts = sorted(set((float(p[0]), float(p[1])) for p in points))
    if len(pts) < 3:
        raise ValueError("a convex polygon needs three points that are not collinear")

    def turn(o, a, b):
        return (a[0] - o[0]) * (b[1] - o[1]) - (a[1] - o[1]) * (b[0] - o[0])

    lower, upper = [], []
    for p in pts:
        while len(lower) >= 2 and turn(lower[-2], lower[-1], p) <= 0.0:
            lower.pop()
        lower.append(p)
    for p in reversed(pts):
        while len(upper) >= 2 and turn(upper[-2], upper[-1], p) <= 0.0:
            upper.pop()
        upper.append(p)
    hull = lower[:-1] + upper[:-1]
    if len(hull) < 3:
        raise ValueError("a convex polygon needs three points that are not collinear")
    return hull


@dataclass
class ConvexPolygon:
    """collision2d ConvexPolygon: hull vertices in the local frame (world = origin + rot * local)."""
    points: Tuple[P2, ...]
    origin: P2 = (0.0, 0.0)
    rotation: Tuple[float, float, float, float] = (1.0, 0.0, 0.0, 1.0)

    @staticmethod
    def new_convex_hull(points):
        return ConvexPolygon(tuple(convex_hull(points)))


@dataclass
class LineSegment:
    a: P2
    b: P2

    @staticmethod
    def from_ab(a, b):
        return LineSegment(tuple(a), tuple(b))


@dataclass
class CubicBezier:
    points: Tuple[P2, P2, P2, P2]


@dataclass
class Logic:
    """collision2d Logic{op, a, b, origin, rotation}; children live in the local frame (object.rs:393-410)."""
    op: int
    a: object
    b: object
    origin: P2 = (0.0, 0.0)
    rotation: Tuple[float, float, float, float] = (1.0, 0.0, 0.0, 1.0)


AND, OR, AND_NOT = abi.LG_OP_AND, abi.LG_OP_OR, abi.LG_OP_ANDNOT


@dataclass
class Material:
    refractive_index: float = 1.2  # object.rs:441-446


@dataclass
class Object:
    """Object{object_enum, material_opt, moved} (object.rs:51-55)."""
    geo: object
    material_opt: Optional[Material]
    kind: str = "Geo"
    moved: bool = True

    # constructors mirror object.rs:58-113 (which shapes get a default material)
    @staticmethod
    def new_mirror(a, b):
        return Object(LineSegment.from_ab(a, b), None, "StraightMirror")

    @staticmethod
    def new_curved_mirror(cubic: CubicBezier):
        return Object(cubic, None, "CurvedMirror")

    @staticmethod
    def new_circle(origin, radius):
        return Object(Circle(tuple(origin), radius), Material(), "Circle")

    @staticmethod
    def new_rect(origin, width, height):
        return Object(Rect(tuple(origin), rot2_identity(), width, height), Material(), "Rect")

    @staticmethod
    def new_lens(origin, radius, distance):
        # Lens::new, object.rs:393-410
        return Object(Logic(AND, Circle((distance * 0.5, 0.0), radius), Circle((-distance * 0.5, 0.0), radius),
                            tuple(origin), rot2_identity()), Material(), "Lens")

    @staticmethod
    def new_ellipse(origin, a, b):
        # object.rs:38-45: rot = Rotation2::new(0.0)
        return Object(Ellipse(tuple(origin), a, b, rot2_identity()), Material(), "Ellipse")

    @staticmethod
    def new_convex_polygon(points):
        # object.rs:34-36
        return Object(ConvexPolygon.new_convex_hull(points), Material(), "ConvexPolygon")

    @staticmethod
    def new_geo(geo):
        return Object(geo, Material(), "Geo")

    def get_material(self):
        return self.material_opt

    def with_index(self, n):
        self.material_opt = Material(n)
        return self

    def get_control_lines(self):
        """CurvedMirror::get_control_lines (object.rs:355-365): three red segments."""
        if self.kind != "CurvedMirror":
            return []
        p = self.geo.points
        red = (1.0, 0.0, 0.0, 1.0)
        return [(p[0], p[1], red), (p[1], p[2], red), (p[2], p[3], red)]


# ---- lights (light.rs) --------------------------------------------------------------------
@dataclass
class PointLight:
    position: P2
    num_rays: int
    color: Color

    @staticmethod
    def new(position, num_rays, color):  # argument order of light.rs:153
        return PointLight(tuple(position), int(num_rays), tuple(color))


@dataclass
class SpotLight:
    position: P2
    spot_angle: float
    spot_direction: P2
    num_rays: int
    color: Color

    @staticmethod
    def new(position, spot_angle, spot_direction, num_rays, color):  # light.rs:205-211
        return SpotLight(tuple(position), spot_angle, tuple(spot_direction), int(num_rays), tuple(color))


@dataclass
class DirectionalLight:
    """light.rs:77-115.  `neg_r`: how `start.eval_at_r(-(i as f64) / n)` (light.rs:111) is read -- collision2d's
    LineSegment::eval_at_r is not in the reference.  False (default): ray origins walk the drawn segment,
    a + (i/n)(b - a) (ORACLE.md 6.3).  True: the call taken literally on eval_at_r(r) = a + r (b - a), origins
    a - (i/n)(b - a).  Directional-light parity is UNVERIFIED either way."""
    color: Color
    num_rays: int
    start: LineSegment
    neg_r: bool = False

    @staticmethod
    def new(color, num_rays, start):  # light.rs:91
        return DirectionalLight(tuple(color), int(num_rays), start)


def light_to_pod(l) -> abi.LgLight:
    o = abi.LgLight()
    o.num_rays = int(l.num_rays)
    o.color[:] = [float(np.float32(c)) for c in l.color]
    if isinstance(l, PointLight):
        o.kind = abi.LG_LIGHT_POINT
        o.position[:] = l.position
    elif isinstance(l, SpotLight):
        o.kind = abi.LG_LIGHT_SPOT
        o.position[:] = l.position
        o.spot_angle = l.spot_angle
        o.spot_direction[:] = l.spot_direction
    elif isinstance(l, DirectionalLight):
        o.kind = abi.LG_LIGHT_DIRECTIONAL
        o.position[:] = l.start.a
        o.b[:] = l.start.b
        if l.neg_r:
            o.flags |= abi.LG_LIGHT_DIRECTIONAL_NEG_R
    else:
        raise TypeError(f"not a light: {l!r}")
    # tracer.rs:279-287 chains `drawing_object` behind the scene's objects: when the object being dragged out contains
    # the light and has a material it wins the start-medium scan (Tracer.add_drawing_object sets this attribute)
    sm = getattr(l, "start_medium", None)
    if sm is not None:
        o.flags |= abi.LG_LIGHT_START_MEDIUM
        o.start_medium = float(sm)
    return o


def lights_to_array(lights: Sequence):
    arr = (abi.LgLight * max(1, len(lights)))()
    for i, l in enumerate(lights):
        arr[i] = light_to_pod(l)
    return arr


# ---- flattening of Geo trees ----------------------------------------------------------------
def _push_geo(geo, nodes: list) -> int:
    n = abi.LgGeoNode()
    n.child_a = n.child_b = -1
    n.rot[:] = rot2_identity()
    if isinstance(geo, Circle):
        n.kind = abi.LG_GEO_CIRCLE
        n.p[0], n.p[1], n.p[2] = geo.origin[0], geo.origin[1], geo.radius
    elif isinstance(geo, Rect):
        n.kind = abi.LG_GEO_RECT
        n.p[0], n.p[1], n.p[2], n.p[3] = geo.origin[0], geo.origin[1], geo.width, geo.height
        n.rot[:] = geo.rotation
    elif isinstance(geo, Ellipse):
        n.kind = abi.LG_GEO_ELLIPSE
        n.p[0], n.p[1], n.p[2], n.p[3] = geo.origin[0], geo.origin[1], geo.a, geo.b
        n.rot[:] = geo.rot
    elif isinstance(geo, LineSegment):
        n.kind = abi.LG_GEO_SEGMENT
        n.p[0], n.p[1], n.p[2], n.p[3] = geo.a[0], geo.a[1], geo.b[0], geo.b[1]
    elif isinstance(geo, CubicBezier):
        n.kind = abi.LG_GEO_BEZIER
        for k, pt in enumerate(geo.points):
            n.p[2 * k], n.p[2 * k + 1] = pt[0], pt[1]
    elif isinstance(geo, ConvexPolygon):
        k = len(geo.points)
        if not 3 <= k <= abi.LG_POLYGON_MAX_VERTICES:
            raise ValueError(f"convex polygon with {k} vertices (3..{abi.LG_POLYGON_MAX_VERTICES} supported)")
        n.kind = abi.LG_GEO_POLYGON
        n.op = k
        n.p[0], n.p[1] = geo.origin
        n.rot[:] = geo.rotation
        ix = len(nodes)
        nodes.append(n)
        prev = n
        for v in range(0, k, 4):   # continuation nodes, four vertices each
            c = abi.LgGeoNode()
            c.kind, c.child_a, c.child_b = abi.LG_GEO_POINTS, -1, -1
            c.rot[:] = rot2_identity()
            chunk = geo.points[v:v + 4]
            c.op = len(chunk)
            for q, pt in enumerate(chunk):
                c.p[2 * q], c.p[2 * q + 1] = pt[0], pt[1]
            prev.child_a = len(nodes)
            nodes.append(c)
            prev = c
        return ix
    elif isinstance(geo, Logic):
        n.kind = abi.LG_GEO_LOGIC
        n.op = geo.op
        n.p[0], n.p[1] = geo.origin
        n.rot[:] = geo.rotation
        ix = len(nodes)
        nodes.append(n)
        n.child_a = _push_geo(geo.a, nodes)
        n.child_b = _push_geo(geo.b, nodes)
        return ix
    else:
        raise TypeError(f"unsupported Geo: {geo!r}")
    nodes.append(n)
    return len(nodes) - 1


def flatten_objects(objects: Sequence[Object]):
    """Vec<Object> -> (LgObject[], n, LgGeoNode[], n)."""
    nodes: list = []
    objs = (abi.LgObject * max(1, len(objects)))()
    for i, ob in enumerate(objects):
        objs[i].root = _push_geo(ob.geo, nodes)
        objs[i].has_material = 1 if ob.material_opt is not None else 0
        objs[i].refractive_index = ob.material_opt.refractive_index if ob.material_opt is not None else 0.0
    arr = (abi.LgGeoNode * max(1, len(nodes)))()
    for i, n in enumerate(nodes):
        arr[i] = n
    return objs, len(objects), arr, len(nodes)


def trace_params(max_bounce: int, cutoff_color: Sequence[float], canvas_bounds: Rect) -> abi.LgTraceParams:
    p = abi.LgTraceParams()
    p.max_bounce = int(max_bounce)
    p.cutoff_color[:] = [float(np.float32(c)) for c in cutoff_color]
    p.canvas_tlbr[:] = canvas_bounds.tlbr()
    return p


# ---- string mod (string_mod.rs) ---------------------------------------------------------------
class Curve:
    """string_mod.rs:182-188.  Circle | ComplexExp{c} | Hypotrochoid{r, s, d} | Lissajous{a, b, delta}."""

    def __init__(self, kind, params=()):
        self.kind, self.params = kind, tuple(float(p) for p in params)

    def __eq__(self, other):
        return isinstance(other, Curve) and (self.kind, self.params) == (other.kind, other.params)

    def __hash__(self):
        return hash((self.kind, self.params))

    @staticmethod
    def ComplexExp(c: complex):
        return Curve(abi.LG_CURVE_COMPLEX_EXP, (c.real, c.imag))

    @staticmethod
    def Hypotrochoid(r: int, s: int, d: int):
        return Curve(abi.LG_CURVE_HYPOTROCHOID, (int(r), int(s), int(d)))

    @staticmethod
    def Lissajous(a: int, b: int, delta: float):
        return Curve(abi.LG_CURVE_LISSAJOUS, (int(a), int(b), delta))


Curve.Circle = Curve(abi.LG_CURVE_CIRCLE)


class StringModMode:
    Add, Mul, Pow, Base = abi.LG_SM_ADD, abi.LG_SM_MUL, abi.LG_SM_POW, abi.LG_SM_BASE


@dataclass
class ModRemColor:
    modulo: int
    rem: int
    color: Color


@dataclass
class StringMod:
    """Defaults of StringMod::new (string_mod.rs:18-31)."""
    modulo: int = 5
    num: int = 1
    pow: int = 0
    color: Color = (1.0, 1.0, 1.0, 1.0)
    turns: int = 1
    init_curve: Curve = field(default_factory=lambda: Curve.Circle)
    mode: int = StringModMode.Mul
    modulo_colors: List[ModRemColor] = field(default_factory=list)
    modulo_color_index: int = 0
    nested: Optional["StringMod"] = None   # string_mod.rs:14: drawn between the crossings of this pattern's chords

    def to_pod(self):
        s = abi.LgStringMod()
        s.modulo, s.num, s.turns = int(self.modulo), int(self.num), int(self.turns)
        s.mode, s.curve = int(self.mode), int(self.init_curve.kind)
        for i, v in enumerate(self.init_curve.params):
            s.curve_p[i] = v
        s.color[:] = [float(np.float32(c)) for c in self.color]
        rules = (abi.LgModRemColor * max(1, len(self.modulo_colors)))()
        for i, r in enumerate(self.modulo_colors):
            rules[i].modulo, rules[i].rem = int(r.modulo), int(r.rem)
            rules[i].color[:] = [float(np.float32(c)) for c in r.color]
        return s, rules, len(self.modulo_colors)
