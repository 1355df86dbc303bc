"""Regenerates tests/golden/*.npz from the oracle (f64).  These are NOT reference outputs — the reference cannot
be built or run offline (oracle/ORACLE.md, "PARITY UNPINNED") — they freeze the oracle's own answers on small
samples of the BASELINE configs so that (a) the oracle cannot drift silently and (b) the GPU box, which has no
/root/reference and may have a different libm, checks the device against committed bits.

    python tests/golden/make_golden.py [name ...]      (default: all; existing files are rewritten with the same bits)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))


def golden_specs():
    from light_garden_b200 import scenes
    from util import ellipse_spec, polygon_spec
    ell, poly = ellipse_spec(total_rays=300), polygon_spec(total_rays=300)
    ell.width = poly.width = 240
    ell.height = poly.height = 135
    return {
        "ell": ell,     # SURVEY.md 8f rank 2 geometry: ellipses and convex polygons, alone and in CSG trees
        "poly": poly,
        "c1": scenes.c1_default(total_rays=360, width=240, height=135),
        "c2": scenes.c2_cavity(total_rays=48, max_bounce=64, width=240, height=135),
        "c3": scenes.c3_refraction(total_rays=400, grid=16, width=240, height=135),
        "c5": scenes.c5_large(n_lights=2, rays_per_light=150, grid=16, width=240, height=135),
    }


def main():
    import lg_oracle as oracle
    from light_garden_b200 import abi
    from util import primary_rays
    only = set(sys.argv[1:])
    for name, spec in golden_specs().items():
        if only and name not in only:
            continue
        osc = oracle.OracleScene.from_spec(spec)
        rays = primary_rays(oracle, spec, osc)
        out = {"rays": rays}
        for tag, prec in (("f64", abi.LG_PRECISION_F64), ("f32", abi.LG_PRECISION_F32)):
            res = osc.trace_rays(rays, prec)
            out[f"seg_{tag}"] = res.seg
            out[f"tags_{tag}"] = res.tags
            if tag == "f64":
                out["end_f64"] = res.f64
        img = oracle.new_image(spec.width, spec.height)
        out["fragments"] = np.int64(oracle.accumulate_segments(img, out["seg_f32"]))
        out["image_sum"] = img.sum(axis=(0, 1), dtype=np.float64)
        np.savez_compressed(os.path.join(HERE, f"{name}.npz"), **out)
        print(name, len(rays), "rays", len(out["seg_f64"]), "segments")


if __name__ == "__main__":
    main()
