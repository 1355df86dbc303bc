"""The oracle's restatement of the reference's TileMap (src/light_garden/tile_map.rs; enabled by default, line 61;
used at tracer.rs:385-411): indexing formulas as written in the reference, and -- the property the reference relies
on -- the same segments as the all-objects loop.  ORACLE.md 5.6.  CPU only."""
import math

import numpy as np
import pytest

from light_garden_b200 import abi, scenes
from light_garden_b200.scene import Object, PointLight
from util import small_specs


def test_tile_and_slab_indexing(oracle):
    """TileMap::get_tile (tile_map.rs:134-146) and Tile::get_index (229-235) on a 16:9 window, 100 x 100 x 8."""
    spec = scenes.c1_default(total_rays=100, width=480, height=270)
    osc = oracle.OracleScene.from_spec(spec)
    osc.enable_tile_map(True, 100, 100, 8)
    w = 2 * 480 / 270
    assert osc.tile_of(-w / 2 + 1e-9, -1 + 1e-9) == 0                      # bottom-left tile
    assert osc.tile_of(w / 2 - 1e-9, -1 + 1e-9) == 99
    assert osc.tile_of(-w / 2 + 1e-9, 1 - 1e-9) == 9900
    assert osc.tile_of(0.0, 0.0) == 50 + 50 * 100
    assert osc.tile_of(w, 0.0) == -1 and osc.tile_of(0.0, 1.5) == -1       # outside: no tile
    assert osc.tile_of(-w, 0.0) == 50 * 100                                 # `as usize` saturates a negative index to 0
    # clockwise angle from +y: index = (8 * angle / tau - EPSILON) as usize with EPSILON = f64::EPSILON (ORACLE.md 5.6:
    # collision2d's constant is not in the repository).  2 - eps is representable, 4 - eps and 6 - eps round back
    for (dx, dy), k in (((0, 1), 0), ((1, 0), 1), ((0, -1), 4), ((-1, 0), 6)):
        assert osc.slab_of(dx, dy) == k
    for deg in range(0, 360, 7):
        a = math.radians(deg) + 1e-3
        assert osc.slab_of(math.sin(a), math.cos(a)) == int(8 * (a % math.tau) / math.tau)


def test_candidates_of_a_tile(oracle):
    """One circle straight above the window centre: it is a candidate of the upward sectors of the centre tile, of
    every sector of the tiles it overlaps, and of no downward sector below it."""
    spec = scenes.SceneSpec("one", [Object.new_circle((0.0, 0.6), 0.05)], [PointLight.new((0.0, 0.0), 10, (0.01,) * 4)],
                            max_bounce=2, width=160, height=90)
    osc = oracle.OracleScene.from_spec(spec)
    osc.enable_tile_map(True, 100, 100, 8)
    centre = osc.tile_of(0.0, 0.0)
    up = {osc.slab_of(0.0, 1.0), osc.slab_of(-1e-3, math.sqrt(1 - 1e-6))}           # sectors 0 and 7 meet at +y
    for k in range(8):
        assert (0 in osc.tile_candidates(centre, k)) == (k in up)
    inside = osc.tile_of(0.0, 0.6)
    assert all(0 in osc.tile_candidates(inside, k) for k in range(8))                # overlap: every sector
    above = osc.tile_of(0.0, 0.9)
    assert 0 in osc.tile_candidates(above, osc.slab_of(0.0, -1.0)) and 0 not in osc.tile_candidates(above, osc.slab_of(0.0, 1.0))


@pytest.mark.parametrize("name", sorted(small_specs()))
def test_tile_map_gives_the_all_objects_result(oracle, name):
    """tile_map.rs is a culling structure: with it the trace returns the segments of the all-objects loop (both
    precisions), with fewer Ray::intersect calls on the many-object scenes."""
    spec = small_specs()[name]
    osc = oracle.OracleScene.from_spec(spec)
    for prec in (abi.LG_PRECISION_F64, abi.LG_PRECISION_F32):
        osc.enable_tile_map(False)
        a = osc.trace_all(spec.lights, prec)
        assert a.object_tests == a.ray_steps * osc.n_obj
        assert osc.enable_tile_map(True) > 0
        b = osc.trace_all(spec.lights, prec)
        assert b.ray_steps == a.ray_steps and b.segments_emitted == a.segments_emitted
        assert b.seg.tobytes() == a.seg.tobytes() and b.tags.tobytes() == a.tags.tobytes()
        assert b.f64.tobytes() == a.f64.tobytes()
        assert b.object_tests <= a.object_tests
        if osc.n_obj >= 256:
            assert b.object_tests < a.object_tests / 4       # 8 sectors of 45 degrees: about a sixth to an eighth
