"""default.ron: the embedded C1 fixture equals the reference's file (when the checkout is present),
and the RON reader round-trips the shapes serde writes."""
import os

import pytest

from light_garden_b200 import scenes
from light_garden_b200.ron import load_scene, parse

REF = "/root/reference/default.ron"


@pytest.mark.skipif(not os.path.exists(REF), reason="reference checkout not present (GPU box)")
def test_embedded_default_scene_matches_reference_file():
    objects, lights = load_scene(open(REF).read())
    assert objects == scenes.default_objects()
    assert lights == scenes.default_lights()


def test_parse_shapes():
    v = parse("( a: 1, b: [2, 3.5e-1], c: Some(( x: -1 )), d: None, e: Foo((1, 2)), )")
    assert v["a"] == 1 and v["b"] == [2, 0.35]
    assert v["c"] == ("Some", [{"x": -1}]) and v["d"] == ("None", None) and v["e"] == ("Foo", [[1, 2]])


def test_load_minimal_scene():
    text = """([ ( object_enum: Circle(( origin: [0.5, 0.25], radius: 0.1, )), material_opt: Some(( refractive_index: 1.5, )), moved: false, ),
                ( object_enum: Geo(GeoLogic(( op: AndNot, a: GeoRect(( origin: [0,0], rotation: [1,0,0,1], width: 1, height: 2, )),
                    b: GeoCircle(( origin: [0, 0], radius: 0.3, )), origin: [0.1, 0.2], rotation: [1, 0, 0, 1], ))), material_opt: None, moved: true, ), ],
               [ PointLight(( position: [0, 0], color: (0.1, 0.2, 0.3, 0.4), num_rays: 7, )) ])"""
    objects, lights = load_scene(text)
    assert objects[0].geo.radius == 0.1 and objects[0].material_opt.refractive_index == 1.5
    assert objects[1].geo.op == scenes.AND_NOT and objects[1].material_opt is None
    assert lights[0].num_rays == 7
