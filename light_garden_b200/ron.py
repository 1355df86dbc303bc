"""Minimal RON reader / writer for scene files shaped like the reference's default.ron:
`(Vec<Object>, Vec<Light>)` as written by Tracer::serialize (tracer.rs:183-188)
and read by Tracer::load / LightGarden::load_from_file (tracer.rs:190-204, mod.rs:708-725).

Supports what serde emits for the types in that tuple: structs `( name: v, )`,
tuples `( v, v )`, sequences `[ v, v ]`, enum variants `Name(..)`, `Some(..)`,
`None`, numbers, booleans.  Every ObjectE and Light variant is covered.  The
field names of `Rect`, `Circle`, `Logic`, `CubicBezier`, `LineSegment` are pinned
by default.ron and those of `Ellipse` by object.rs:38-43; collision2d's
`ConvexPolygon` layout is in neither (the crate is not vendored): it is read and
written as `(points: [[x, y], ..], origin: [x, y], rotation: [m11, m21, m12, m22])`,
the names a derived Serialize gives a struct with the fields this package's
ConvexPolygon has, with `origin` / `rotation` (or `rot`) optional on input.
"""
import re

import numpy as np

from .scene import (AND, AND_NOT, OR, Circle, ConvexPolygon, CubicBezier, DirectionalLight, Ellipse, LineSegment, Logic,
                    Material, Object, PointLight, Rect, SpotLight)

_TOKEN = re.compile(r"\s*(?:(//[^\n]*)|([A-Za-z_][A-Za-z_0-9]*)|([-+]?(?:\d+\.?\d*(?:[eE][-+]?\d+)?|\.\d+(?:[eE][-+]?\d+)?|inf|NaN))|(.))")


def _tokens(text):
    pos = 0
    out = []
    while pos < len(text):
        m = _TOKEN.match(text, pos)
        if not m:
            break
        pos = m.end()
        if m.group(1):
            continue
        if m.group(2):
            out.append(("id", m.group(2)))
        elif m.group(3):
            out.append(("num", float(m.group(3))))
        elif m.group(4) and not m.group(4).isspace():
            out.append(("sym", m.group(4)))
    return out


class _Parser:
    def __init__(self, text):
        self.t = _tokens(text)
        self.i = 0

    def peek(self):
        return self.t[self.i] if self.i < len(self.t) else ("eof", None)

    def take(self, kind=None, val=None):
        k, v = self.peek()
        if (kind and k != kind) or (val is not None and v != val):
            raise ValueError(f"RON: expected {kind} {val}, got {k} {v} at token {self.i}")
        self.i += 1
        return v

    def value(self):
        k, v = self.peek()
        if k == "num":
            self.i += 1
            return v
        if k == "sym" and v == "[":
            return self.seq("[", "]")
        if k == "sym" and v == "(":
            return self.paren()
        if k == "id":
            self.i += 1
            if v in ("true", "false"):
                return v == "true"
            if self.peek() == ("sym", "("):
                return (v, self.paren())  # enum variant / Some(..)
            return (v, None)  # unit variant / None
        raise ValueError(f"RON: unexpected token {k} {v}")

    def seq(self, open_, close):
        self.take("sym", open_)
        items = []
        while self.peek() != ("sym", close):
            items.append(self.value())
            if self.peek() == ("sym", ","):
                self.i += 1
        self.take("sym", close)
        return items

    def paren(self):
        """`( name: v, .. )` -> dict, `( v, .. )` -> list."""
        self.take("sym", "(")
        named, items = {}, []
        while self.peek() != ("sym", ")"):
            k, v = self.peek()
            nxt = self.t[self.i + 1] if self.i + 1 < len(self.t) else ("eof", None)
            if k == "id" and nxt == ("sym", ":"):
                self.i += 2
                named[v] = self.value()
            else:
                items.append(self.value())
            if self.peek() == ("sym", ","):
                self.i += 1
        self.take("sym", ")")
        return named if named and not items else items


def parse(text):
    return _Parser(text).value()


def _single(x):
    return x[0] if isinstance(x, list) and len(x) == 1 else x


_OPS = {"And": AND, "Or": OR, "AndNot": AND_NOT}


def _geo(v):
    name, inner = v
    d = _single(inner)
    if name == "GeoCircle":
        return Circle(tuple(d["origin"]), d["radius"])
    if name == "GeoRect":
        return Rect(tuple(d["origin"]), tuple(d["rotation"]), d["width"], d["height"])
    if name == "GeoLogic":
        return _logic(d)
    if name == "GeoCubicBezier":
        return CubicBezier(tuple(tuple(p) for p in d["points"]))
    if name == "GeoEllipse":
        return Ellipse(tuple(d["origin"]), d["a"], d["b"], tuple(d["rot"]))
    if name == "GeoLineSegment" and "a" in d and "b" in d:
        return LineSegment(tuple(d["a"]), tuple(d["b"]))
    if name == "GeoConvexPolygon":
        return _polygon(d)
    raise ValueError(f"RON: unsupported Geo variant {name}")


def _polygon(d):
    pts = d["points"] if isinstance(d, dict) else d
    rot = (d.get("rotation") or d.get("rot")) if isinstance(d, dict) else None
    org = d.get("origin") if isinstance(d, dict) else None
    return ConvexPolygon(tuple(tuple(p) for p in pts), tuple(org) if org else (0.0, 0.0),
                         tuple(rot) if rot else (1.0, 0.0, 0.0, 1.0))


def _logic(d):
    return Logic(_OPS[d["op"][0]], _geo(d["a"]), _geo(d["b"]), tuple(d["origin"]), tuple(d["rotation"]))


def _object(d):
    name, inner = d["object_enum"]
    body = _single(inner)
    mat = d["material_opt"]
    material = Material(_single(mat[1])["refractive_index"]) if mat[0] == "Some" else None
    if name == "Lens":
        return Object(_logic(_single(body["l"]) if isinstance(body["l"], list) else body["l"]), material, "Lens", bool(d.get("moved", False)))
    if name == "CurvedMirror":
        cubic = body["cubic"]
        cubic = _single(cubic)
        return Object(CubicBezier(tuple(tuple(p) for p in cubic["points"])), material, "CurvedMirror", bool(d.get("moved", False)))
    if name == "Rect":
        return Object(Rect(tuple(body["origin"]), tuple(body["rotation"]), body["width"], body["height"]), material, "Rect", bool(d.get("moved", False)))
    if name == "Circle":
        return Object(Circle(tuple(body["origin"]), body["radius"]), material, "Circle", bool(d.get("moved", False)))
    if name == "Ellipse":
        return Object(Ellipse(tuple(body["origin"]), body["a"], body["b"], tuple(body["rot"])), material, "Ellipse",
                      bool(d.get("moved", False)))
    if name == "ConvexPolygon":
        return Object(_polygon(body), material, "ConvexPolygon", bool(d.get("moved", False)))
    if name == "Geo":
        return Object(_geo(_single(inner)), material, "Geo", bool(d.get("moved", False)))
    if name == "StraightMirror":
        ls = _single(body["line_segment"])
        if "a" in ls and "b" in ls:
            return Object(LineSegment(tuple(ls["a"]), tuple(ls["b"])), material, "StraightMirror", bool(d.get("moved", False)))
    raise ValueError(f"RON: unsupported ObjectE variant {name}")


def _light(v):
    name, inner = v
    d = _single(inner)
    color = tuple(d["color"])
    if name == "PointLight":
        return PointLight(tuple(d["position"]), int(d["num_rays"]), color)
    if name == "SpotLight":
        return SpotLight(tuple(d["position"]), d["spot_angle"], tuple(d["spot_direction"]), int(d["num_rays"]), color)
    if name == "DirectionalLight":
        st = _single(d["start"])
        return DirectionalLight(color, int(d["num_rays"]), LineSegment(tuple(st["a"]), tuple(st["b"])))
    raise ValueError(f"RON: unsupported Light variant {name}")


def load_scene(text):
    objects, lights = parse(text)
    return [_object(o) for o in objects], [_light(l) for l in lights]


# ---- writer: Tracer::serialize (tracer.rs:183-188) ------------------------------------------------
def _f(x):
    """Shortest round-trip decimal without an exponent, integral values without a fraction: what ron writes
    (default.ron: `radius: 2`, `0.00000000000000006123233995736766`)."""
    return np.format_float_positional(float(x), trim="-")


def _f32(x):
    return np.format_float_positional(np.float32(x), trim="-")


def _seq(vals):
    return "[" + ", ".join(_f(v) for v in vals) + "]"


_OP_NAMES = {AND: "And", OR: "Or", AND_NOT: "AndNot"}


def _logic_body(l: Logic):
    return (f"(op: {_OP_NAMES[l.op]}, a: {_geo_text(l.a)}, b: {_geo_text(l.b)}, origin: {_seq(l.origin)}, "
            f"rotation: {_seq(l.rotation)})")


def _rect_body(r: Rect):
    return f"(origin: {_seq(r.origin)}, rotation: {_seq(r.rotation)}, width: {_f(r.width)}, height: {_f(r.height)})"


def _circle_body(c: Circle):
    return f"(origin: {_seq(c.origin)}, radius: {_f(c.radius)})"


def _ellipse_body(e: Ellipse):
    return f"(origin: {_seq(e.origin)}, a: {_f(e.a)}, b: {_f(e.b)}, rot: {_seq(e.rot)})"


def _bezier_body(c: CubicBezier):
    return "(points: (" + ", ".join(_seq(p) for p in c.points) + "))"


def _segment_body(s: LineSegment):
    return f"(a: {_seq(s.a)}, b: {_seq(s.b)})"


def _polygon_body(p: ConvexPolygon):
    return ("(points: [" + ", ".join(_seq(q) for q in p.points) + f"], origin: {_seq(p.origin)}, "
            f"rotation: {_seq(p.rotation)})")


def _geo_text(g):
    if isinstance(g, Circle):
        return f"GeoCircle({_circle_body(g)})"
    if isinstance(g, Rect):
        return f"GeoRect({_rect_body(g)})"
    if isinstance(g, Logic):
        return f"GeoLogic({_logic_body(g)})"
    if isinstance(g, CubicBezier):
        return f"GeoCubicBezier({_bezier_body(g)})"
    if isinstance(g, Ellipse):
        return f"GeoEllipse({_ellipse_body(g)})"
    if isinstance(g, LineSegment):
        return f"GeoLineSegment({_segment_body(g)})"
    if isinstance(g, ConvexPolygon):
        return f"GeoConvexPolygon({_polygon_body(g)})"
    raise TypeError(f"not a Geo: {g!r}")


def _object_text(o: Object):
    g, k = o.geo, o.kind
    if k == "Lens" and isinstance(g, Logic):
        e = f"Lens((l: {_logic_body(g)}))"
    elif k == "CurvedMirror" and isinstance(g, CubicBezier):
        e = f"CurvedMirror((cubic: {_bezier_body(g)}))"
    elif k == "StraightMirror" and isinstance(g, LineSegment):
        e = f"StraightMirror((line_segment: {_segment_body(g)}))"
    elif k == "Rect" and isinstance(g, Rect):
        e = f"Rect({_rect_body(g)})"
    elif k == "Circle" and isinstance(g, Circle):
        e = f"Circle({_circle_body(g)})"
    elif k == "Ellipse" and isinstance(g, Ellipse):
        e = f"Ellipse({_ellipse_body(g)})"
    elif k == "ConvexPolygon" and isinstance(g, ConvexPolygon):
        e = f"ConvexPolygon({_polygon_body(g)})"
    else:
        e = f"Geo({_geo_text(g)})"
    m = f"Some((refractive_index: {_f(o.material_opt.refractive_index)}))" if o.material_opt is not None else "None"
    return f"(object_enum: {e}, material_opt: {m}, moved: {'true' if o.moved else 'false'})"


def _light_text(l):
    col = "(" + ", ".join(_f32(c) for c in l.color) + ")"
    if isinstance(l, PointLight):
        return f"PointLight((position: {_seq(l.position)}, color: {col}, num_rays: {int(l.num_rays)}))"
    if isinstance(l, SpotLight):
        return (f"SpotLight((position: {_seq(l.position)}, color: {col}, num_rays: {int(l.num_rays)}, "
                f"spot_angle: {_f(l.spot_angle)}, spot_direction: {_seq(l.spot_direction)}))")
    if isinstance(l, DirectionalLight):
        return f"DirectionalLight((color: {col}, num_rays: {int(l.num_rays)}, start: {_segment_body(l.start)}))"
    raise TypeError(f"not a light: {l!r}")


def serialize_scene(objects, lights) -> str:
    """RON text of `(Vec<Object>, Vec<Light>)` (Tracer::serialize, tracer.rs:183-188); load_scene reads it back to
    equal objects.  `rays` of the lights are skipped like the reference's #[serde(skip)] fields."""
    objs = ",\n    ".join(_object_text(o) for o in objects)
    lts = ",\n    ".join(_light_text(l) for l in lights)
    return f"([\n    {objs}\n], [\n    {lts}\n])\n"
