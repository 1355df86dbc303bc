"""Minimal RON reader for scene files shaped like the reference's default.ron:
`(Vec<Object>, Vec<Light>)` as written by Tracer::serialize (tracer.rs:183-188)
and read by LightGarden::load_from_file (mod.rs:708-725).

Supports what serde emits for the types in that tuple: structs `( name: v, )`,
tuples `( v, v )`, sequences `[ v, v ]`, enum variants `Name(..)`, `Some(..)`,
`None`, numbers, booleans.
"""
import re

from .scene import (AND, AND_NOT, OR, Circle, CubicBezier, DirectionalLight, Ellipse, LineSegment, Logic, Material, Object,
                    PointLight, Rect, SpotLight)

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
    raise ValueError(f"RON: unsupported Geo variant {name}")


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
