#!/usr/bin/env python
"""The one output of the real reference that exists offline: /root/reference/screenshot.jpg (1800 x 1000), a frame of the
app's start-up scene (Tracer::new, tracer.rs:19-26: point light at (-0.1, 0.1), curved mirror CubicBezier::new_sample2()
scaled by 0.5 -- its red control polygon is in the picture: (0, 0.5) (0.7, 0) (0.3, 1) (1, 0.3) to the pixel).  Two
edges in it depend only on the light, the mirror's end point and end tangent, the reflection law and the projection:
the shadow edge of the mirror's end on the top border, and the last reflected ray on the left border.  This script
measures both in the JPEG and computes them with the oracle.  Not a test (the picture cannot travel and the rest of that
scene is not reproducible from the repository); the numbers are quoted in DESIGN.md section 2.
    python tests/screenshot_landmarks.py [/root/reference/screenshot.jpg]
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def main():
    path = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/screenshot.jpg"
    from PIL import Image
    import lg_oracle as oracle
    from light_garden_b200 import abi, scenes
    from light_garden_b200.scene import CubicBezier, Object, PointLight
    im = np.asarray(Image.open(path).convert("L")).astype(float)
    H, W = im.shape
    half_w, half_h = W / 2.0, H / 2.0          # world y in [-1, 1], x in [-W/H, W/H]: 500 px per unit
    scale = half_h

    def top_edge(row):
        k = np.convolve(im[row, 850:1100], np.ones(5) / 5, "same")
        return 850 + int(np.argmax(k < 12))

    col = np.convolve(im[:, 2:12].mean(axis=1), np.ones(9) / 9, "same")
    left_edge = 700 + int(np.argmax(np.abs(np.gradient(col)[700:950])))

    oracle.build()
    mirror = Object.new_curved_mirror(CubicBezier([(0, 0.5), (0.7, 0.0), (0.3, 1.0), (1.0, 0.3)]))
    spec = scenes.SceneSpec("startup", [mirror], [PointLight.new((-0.1, 0.1), 200000, (0.01,) * 4)], max_bounce=5,
                            width=W, height=H)
    osc = oracle.OracleScene.from_spec(spec)
    res = osc.trace_all(spec.lights, abi.LG_PRECISION_F64)
    gen, seg = res.tags["generation"], res.f64
    refl = (gen == 1) & (np.abs(seg["b"][:, 0] + spec.aspect) < 1e-9)
    y_last = seg["b"][refl, 1].max()

    def shadow_x(row):                         # direct rays that end on the line y(row): the rightmost one left of the mirror
        y = (half_h - row) / scale
        d = seg["b"][gen == 0] - seg["a"][gen == 0]
        a = seg["a"][gen == 0]
        ok = (d[:, 1] > 0) & (a[:, 1] + d[:, 1] >= y - 1e-12)
        x = a[ok, 0] + d[ok, 0] * (y - a[ok, 1]) / d[ok, 1]
        x = x[(x > -0.5) & (x < 0.5)]          # the mirror's shadow lies to the right of these
        return x.max() if len(x) else np.nan

    print(f"{'landmark':46s} {'screenshot.jpg':>15s} {'oracle':>10s}")
    for row in (3, 100):
        print(f"{'shadow edge of the mirror end, x at pixel row ' + str(row):46s} {top_edge(row):15d} {half_w + scale * shadow_x(row):10.1f}")
    print(f"{'last reflected ray on the left border, y pixel':46s} {left_edge:15d} {half_h - scale * y_last:10.1f}")


if __name__ == "__main__":
    main()
