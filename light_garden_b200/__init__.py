"""light_garden_b200 — B200-native (sm_100a) trace + line accumulation for Light Garden.

The package is a thin host layer over light_garden_b200/_lib/liblight_garden_b200.so
(C ABI: include/light_garden_b200.h).  There is no CPU or PyTorch fallback.
"""
from . import abi
from .scene import (AND, AND_NOT, OR, Circle, ConvexPolygon, CubicBezier, Curve, DirectionalLight, Ellipse, LineSegment, Logic, Material, ModRemColor,
                    Object, PointLight, Rect, SpotLight, StringMod, StringModMode, rot2, rot2_identity)

__all__ = ["abi", "AND", "AND_NOT", "OR", "Circle", "ConvexPolygon", "CubicBezier", "Curve", "DirectionalLight", "Ellipse", "LineSegment", "Logic",
           "Material", "ModRemColor", "Object", "PointLight", "Rect", "SpotLight", "StringMod", "StringModMode",
           "rot2", "rot2_identity", "Context", "Tracer", "Renderer"]


def __getattr__(name):
    # the device classes load the CUDA library on first use
    if name in ("Context", "Tracer", "Renderer", "sort_segments"):
        from . import tracer
        return getattr(tracer, name)
    raise AttributeError(name)
