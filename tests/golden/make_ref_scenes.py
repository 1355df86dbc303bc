#!/usr/bin/env python
"""Writes the scene files rust/dump_golden feeds to the UNMODIFIED reference: the small parity specs of tests/util.py
(C1, C2, C3, C5-16, ellipses, polygons) as RON text `(Vec<Object>, Vec<Light>)`, the format Tracer::load reads
(tracer.rs:190-204), into tests/golden/ref_scenes/.  A leading `// max_bounce = N` comment carries the one Tracer field
the file format has no place for.  Ray counts are cut down (the probe walk records every intersect result)."""
import copy
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

from light_garden_b200.ron import serialize_scene  # noqa: E402
from util import small_specs  # noqa: E402

RAYS_PER_LIGHT = 96


def main():
    out = os.path.join(HERE, "ref_scenes")
    os.makedirs(out, exist_ok=True)
    for name, spec in small_specs().items():
        lights = []
        for l in spec.lights:
            l = copy.deepcopy(l)
            l.num_rays = min(int(l.num_rays), RAYS_PER_LIGHT)
            lights.append(l)
        text = f"// max_bounce = {spec.max_bounce}\n" + serialize_scene(spec.objects, lights)
        path = os.path.join(out, name.lower().replace("-", "_") + ".ron")
        open(path, "w").write(text)
        print(path)


if __name__ == "__main__":
    main()
