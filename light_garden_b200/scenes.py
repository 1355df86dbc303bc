"""The BASELINE.json workloads as scene builders.

C1 is the reference's stock scene (default.ron, decoded in SURVEY.md Appendix C;
tests/test_ron.py checks this embedded copy against the file when the reference
checkout is present).  C2..C5 are synthetic scenes of the named shapes, seeded
with SplitMix64 so every run — device, oracle, any rank — builds the same bits.
"""
import math

from .scene import (AND, AND_NOT, OR, Circle, CubicBezier, DirectionalLight, LineSegment, Logic, Material,
                    ModRemColor, Object, PointLight, Rect, SpotLight, StringMod, StringModMode, rot2, rot2_identity)

_M64 = (1 << 64) - 1


class SplitMix64:
    def __init__(self, seed):
        self.s = seed & _M64

    def next(self):
        self.s = (self.s + 0x9E3779B97F4A7C15) & _M64
        z = self.s
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _M64
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _M64
        return z ^ (z >> 31)

    def uniform(self, lo=0.0, hi=1.0):
        return lo + (hi - lo) * ((self.next() >> 11) * (1.0 / (1 << 53)))

    def below(self, n):
        return self.next() % n


def canvas(aspect):
    """Tracer.canvas_bounds as the app sets it: Rect::from_tlbr(1, -aspect, -1, aspect) (sub_render_pass.rs:156)."""
    return Rect.from_tlbr(1.0, -aspect, -1.0, aspect)


class SceneSpec:
    def __init__(self, name, objects, lights, max_bounce, width, height, cutoff=(0.001,) * 4):
        self.name = name
        self.objects = objects
        self.lights = lights
        self.max_bounce = max_bounce
        self.cutoff_color = list(cutoff)
        self.width, self.height = width, height
        # aspect exactly as sub_render_pass.rs:146,156: f32 division, then `as f64`
        import numpy as np
        self.aspect = float(np.float32(width) / np.float32(height))
        self.canvas_bounds = canvas(self.aspect)

    def total_rays(self):
        return sum(l.num_rays for l in self.lights)

    def apply(self, tracer):
        tracer.clear()
        for o in self.objects:
            tracer.push_object(o)
        for l in self.lights:
            tracer.push_light(l)
        tracer.max_bounce = self.max_bounce
        tracer.cutoff_color = list(self.cutoff_color)
        tracer.resize(self.canvas_bounds)
        return tracer


# ---- C1: default.ron --------------------------------------------------------------------------
def default_objects():
    lens = Object(Logic(AND, Circle((1.9, 0.0), 2.0), Circle((-1.9, 0.0), 2.0),
                        (-0.022772240638732733, -0.09999999999999998),
                        (0.00000000000000006123233995736766, -1.0, 1.0, 0.00000000000000006123233995736766)),
                  Material(1.05), "Lens", False)
    mirror = Object(CubicBezier(((-0.622772240638733, 0.40000000000000013), (-0.3227722406387328, 0.8),
                                 (0.2772277593612673, 0.8), (0.5772277593612682, 0.40000000000000013))),
                    None, "CurvedMirror", False)
    rect = Object(Rect((-0.022772240638732733, -0.5), (1.0, 0.0, 0.0, 1.0), 0.40000000000000036, 0.3999999999999999),
                  Material(1.73), "Rect", False)
    return [lens, mirror, rect]


def default_lights(point_rays=10000, spot_rays=2000):
    return [
        PointLight((-0.022772240638732733, -0.5), point_rays, (0.009721218, 0.009721218, 0.009721218, 0.011764706)),
        SpotLight((1.0772277593612674, -0.09999999999999998), 0.17453292519943295,
                  (-0.9999922358557027, 0.003940587305547067), spot_rays,
                  (0.0036765062, 0.020288562, 0.016807375, 0.03529412)),
    ]


def c1_default(total_rays=1_000_000, width=1920, height=1080):
    """default.ron verbatim, ray counts scaled 10000:2000 (1 M -> 833 334 + 166 666)."""
    spot = total_rays * 2000 // 12000
    return SceneSpec("C1 default.ron", default_objects(), default_lights(total_rays - spot, spot), 5, width, height)


# ---- C2: mirror + curved-mirror cavity ----------------------------------------------------------
def c2_cavity(total_rays=4_000_000, max_bounce=64, width=1920, height=1080, seed=0x4C470002):
    rng = SplitMix64(seed)
    a = 16.0 / 9.0
    bx, by = a - 0.05, 0.95
    objs = [Object.new_mirror((-bx, -by), (bx, -by)), Object.new_mirror((bx, -by), (bx, by)),
            Object.new_mirror((bx, by), (-bx, by)), Object.new_mirror((-bx, by), (-bx, -by))]
    for k in range(4):
        cx = (-0.9 + 0.6 * k) + rng.uniform(-0.1, 0.1)
        cy = rng.uniform(-0.45, 0.45)
        ang = rng.uniform(0.0, math.tau)
        ln = rng.uniform(0.25, 0.45)
        bulge = rng.uniform(0.08, 0.25)
        ux, uy = math.cos(ang), math.sin(ang)
        nx, ny = -uy, ux
        p0 = (cx - ln * ux, cy - ln * uy)
        p3 = (cx + ln * ux, cy + ln * uy)
        p1 = (cx - 0.4 * ln * ux + bulge * nx, cy - 0.4 * ln * uy + bulge * ny)
        p2 = (cx + 0.4 * ln * ux + bulge * nx, cy + 0.4 * ln * uy + bulge * ny)
        objs.append(Object.new_curved_mirror(CubicBezier((p0, p1, p2, p3))))
    lights = [PointLight((0.013, 0.007), total_rays, (0.002, 0.002, 0.002, 0.01))]
    return SceneSpec("C2 cavity", objs, lights, max_bounce, width, height)


# ---- C3: 256 refractive CSG objects ----------------------------------------------------------------
_INDICES = (1.05, 1.2, 1.33, 1.5, 1.73, 2.4)


def c3_refraction(total_rays=16_000_000, grid=16, width=1920, height=1080, seed=0x4C470003):
    rng = SplitMix64(seed)
    a = 16.0 / 9.0
    cw, ch = 2 * a / grid, 2.0 / grid
    objs = []
    for gy in range(grid):
        for gx in range(grid):
            s = min(cw, ch)
            cx = -a + (gx + 0.5) * cw + rng.uniform(-0.12, 0.12) * cw
            cy = -1.0 + (gy + 0.5) * ch + rng.uniform(-0.12, 0.12) * ch
            kind = rng.below(5)
            n = _INDICES[rng.below(len(_INDICES))]
            r = rng.uniform(0.22, 0.32) * s
            ang = rng.uniform(0.0, math.tau)
            if kind == 0:
                ob = Object.new_circle((cx, cy), r)
            elif kind == 1:
                ob = Object(Rect((cx, cy), rot2(ang), 2.0 * r, 1.4 * r), Material(), "Rect")
            elif kind == 2:  # lens: And of two circles, rotated local frame
                ob = Object(Logic(AND, Circle((0.6 * r, 0.0), r), Circle((-0.6 * r, 0.0), r), (cx, cy), rot2(ang)),
                            Material(), "Lens")
            elif kind == 3:  # circle ∪ rect
                ob = Object(Logic(OR, Circle((0.0, 0.0), 0.8 * r), Rect((0.5 * r, 0.0), rot2_identity(), 1.6 * r, 0.8 * r),
                                  (cx, cy), rot2(ang)), Material(), "Geo")
            else:  # rect \ circle
                ob = Object(Logic(AND_NOT, Rect((0.0, 0.0), rot2_identity(), 2.0 * r, 1.6 * r), Circle((0.7 * r, 0.0), 0.7 * r),
                                  (cx, cy), rot2(ang)), Material(), "Geo")
            objs.append(ob.with_index(n))
    q = total_rays // 4
    # lights sit on cell corners (objects stay inside their cells, so no light starts inside one)
    lights = [
        PointLight((-a + 4 * cw, -1.0 + 4 * ch), q, (0.012, 0.004, 0.003, 0.02)),
        PointLight((-a + 12 * cw, -1.0 + 11 * ch), q, (0.003, 0.004, 0.012, 0.02)),
        SpotLight((-a + 0.5 * cw, -1.0 + 8 * ch), 0.6, (1.0, 0.05), q, (0.004, 0.012, 0.004, 0.02)),
        DirectionalLight((0.008, 0.008, 0.003, 0.02), total_rays - 3 * q,
                         LineSegment((-a + 2 * cw, 1.0 - 0.02), (a - 2 * cw, 1.0 - 0.02))),
    ]
    return SceneSpec("C3 refraction", objs, lights, 5, width, height)


# ---- C4: string mod ------------------------------------------------------------------------------------
def c4_string_mod(modulo=10_000_000, num=2):
    k = 1e-3
    return StringMod(modulo=modulo, num=num, turns=1, mode=StringModMode.Mul, color=(k, k, k, k),
                     modulo_colors=[ModRemColor(3, 0, (k, 0.0, 0.0, k)), ModRemColor(3, 1, (0.0, k, 0.0, k)),
                                    ModRemColor(3, 2, (0.0, 0.0, k, k))])


# ---- C5: 4096-object scene ---------------------------------------------------------------------------------
def c5_objects(grid=64, seed=0x4C470005):
    """2048 circles, 1024 straight mirrors, 1024 rects on a jittered grid x grid lattice."""
    rng = SplitMix64(seed)
    a = 16.0 / 9.0
    cw, ch = 2 * a / grid, 2.0 / grid
    objs = []
    for gy in range(grid):
        for gx in range(grid):
            size = rng.uniform(0.008, 0.012) * (64.0 / grid)
            jx = max(0.0, 0.5 * cw - size * 1.05)
            jy = max(0.0, 0.5 * ch - size * 1.05)
            cx = -a + (gx + 0.5) * cw + rng.uniform(-jx, jx)
            cy = -1.0 + (gy + 0.5) * ch + rng.uniform(-jy, jy)
            sel = (gx + 2 * gy) % 4  # 2:1:1 circles : mirrors : rects, interleaved
            n = rng.uniform(1.1, 1.8)
            ang = rng.uniform(0.0, math.tau)
            if sel in (0, 2):
                objs.append(Object.new_circle((cx, cy), size).with_index(n))
            elif sel == 1:
                ux, uy = math.cos(ang) * size, math.sin(ang) * size
                objs.append(Object.new_mirror((cx - ux, cy - uy), (cx + ux, cy + uy)))
            else:
                # rotation limited so the rotated square stays inside its cell
                half = size / math.sqrt(2.0)
                objs.append(Object(Rect((cx, cy), rot2(ang), 2.0 * half, 2.0 * half), Material(n), "Rect"))
    return objs


def c5_lights(n_lights=8, rays_per_light=32_000_000, grid=64):
    a = 16.0 / 9.0
    cw, ch = 2 * a / grid, 2.0 / grid
    cols = [(0.010, 0.006, 0.004), (0.004, 0.010, 0.006), (0.006, 0.004, 0.010), (0.010, 0.010, 0.004),
            (0.004, 0.010, 0.010), (0.010, 0.004, 0.010), (0.008, 0.008, 0.008), (0.012, 0.005, 0.003)]
    lights = []
    for g in range(n_lights):
        gx = (8 + 7 * g) * grid // 64
        gy = (8 + 11 * (g % 5)) * grid // 64
        pos = (-a + gx * cw, -1.0 + gy * ch)  # a lattice corner: outside every object
        c = cols[g % len(cols)]
        lights.append(PointLight(pos, rays_per_light, (c[0], c[1], c[2], 0.02)))
    return lights


def c5_large(n_lights=8, rays_per_light=32_000_000, grid=64, width=3840, height=2160):
    return SceneSpec("C5 4096 objects", c5_objects(grid), c5_lights(n_lights, rays_per_light, grid), 5, width, height)
