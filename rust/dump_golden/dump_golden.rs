//! dump_golden -- golden vectors of the UNMODIFIED reference for the two hot paths of light_garden_b200.
//!
//! This file is copied into a checkout of sphereflow/light_garden as `src/bin/dump_golden.rs` (rust/dump_golden/run.sh
//! does that) and built by the reference's own Cargo.toml, so it links the reference's `light_garden` modules and the
//! pinned `collision2d` (Cargo.lock: a7b471b54a940622f658a16177d63925536f16c3) exactly as the app does.  Nothing of the
//! reference is modified: the modules are included by path, and everything below goes through their public API.
//!
//!   cargo run --release --bin dump_golden -- <out_dir> <scene.ron> [<scene.ron> ...]
//!
//! For every scene file `(Vec<Object>, Vec<Light>)` (default.ron, and the RON dumps of the small C1/C2/C3/C5/ellipse/
//! polygon specs that tests/golden/make_ref_scenes.py writes) it writes `<out_dir>/ref_<stem>.json` with
//!   * the primary rays of every light as `Light::get_rays()` returns them (light.rs:103-115,163-174,225-249: this pins
//!     `LineSegment::eval_at_r`, `get_normal` and `Ray::from_origin`),
//!   * the vertex pairs `Tracer::trace` appends for every primary ray (tracer.rs:360-493), tile map off,
//!   * a "probe walk" of every primary ray: this program follows the ray through the scene with the SAME public calls
//!     the tracer makes at tracer.rs:414 (`ray.intersect(&obj.get_geometry())`), 431/433 (`obj.contains`), 444-449
//!     (`ray.refract`), 477 (`ray.reflect`) and 484-486 (`ray.intersect(&canvas_bounds)` + `get_first`) and records the
//!     raw return values at every step together with the hit-object index.  The walk's own vertex pairs are compared
//!     with `Tracer::trace`'s before anything is written (the program aborts if they differ), so the recorded
//!     hit-object sequences ARE the reference's.
//! tests/test_reference_golden.py holds oracle/ to these files when they exist.
//!
//! Written for this repository (it is not part of the reference); the control flow it follows is the one SURVEY.md 3.2
//! describes, re-stated here only as far as needed to know which call to probe next.
#![allow(dead_code)]
extern crate nalgebra as na;

#[path = "../light_garden/mod.rs"]
pub mod light_garden;

use collision2d::geo::*;
use light_garden::*;
use na::distance_squared;
use std::fmt::Write as _;

fn num(v: f64) -> String {
    if v.is_finite() { format!("{:?}", v) } else { format!("\"{:?}\"", v) }
}
fn num32(v: f32) -> String {
    if v.is_finite() { format!("{:?}", v) } else { format!("\"{:?}\"", v) }
}
fn p2(p: &P2) -> String { format!("[{},{}]", num(p.x), num(p.y)) }
fn v2(x: f64, y: f64) -> String { format!("[{},{}]", num(x), num(y)) }
fn col(c: &Color) -> String { format!("[{},{},{},{}]", num32(c[0]), num32(c[1]), num32(c[2]), num32(c[3])) }

struct Walk<'a> {
    objects: Vec<&'a Object>,
    canvas: Rect,
    cutoff: Color,
    max_bounce: u32,
    steps: String,      // JSON array body
    n_steps: usize,
    lines: Vec<(P2, Color)>,
}

impl<'a> Walk<'a> {
    /// One primary ray, generation by generation in the tracer's queue order (reflected child first).
    fn primary(&mut self, ray_id: usize, ray: &Ray, color: Color, n0: Float) {
        let mut cur: Vec<(Ray, Color, Float, u64)> = vec![(*ray, color, n0, 0)];
        for generation in 0..self.max_bounce {
            if cur.is_empty() { return; }
            let mut next = Vec::new();
            for (ray, color, n, path) in cur.iter() {
                let cut = self.cutoff;
                if (color[0] < cut[0] && color[1] < cut[1] && color[2] < cut[2]) || color[3] < cut[3] { continue; }
                let o = ray.get_origin();
                let d = ray.get_direction();
                let mut rec = String::new();
                write!(rec, "{{\"ray\":{},\"generation\":{},\"path\":{},\"origin\":{},\"direction\":{},\"color\":{},\"n\":{},\"intersect\":[",
                       ray_id, generation, path, p2(&o), v2(d.x, d.y), col(color), num(*n)).unwrap();
                // tracer.rs:414 -- every object, every returned (point, normal), in the library's order
                let mut nearest = Float::MAX;
                let mut target: Option<(P2, Normal, usize)> = None;
                let mut first = true;
                for (index, obj) in self.objects.iter().enumerate() {
                    if let Some(hits) = ray.intersect(&obj.get_geometry()) {
                        for (point, normal) in hits {
                            if !first { rec.push(','); }
                            first = false;
                            write!(rec, "[{},{},{},{},{}]", index, num(point.x), num(point.y), num(normal.x), num(normal.y)).unwrap();
                            let dist_sq = distance_squared(&o, &point);
                            if dist_sq < nearest { nearest = dist_sq; target = Some((point, normal, index)); }
                        }
                    }
                }
                rec.push_str("],");
                match target {
                    Some((point, normal, index)) => {
                        let obj = self.objects[index];
                        write!(rec, "\"hit_object\":{},\"hit_point\":{},\"hit_normal\":{},", index, p2(&point), v2(normal.x, normal.y)).unwrap();
                        if let Some(material) = obj.material_opt {
                            let inside = obj.contains(&o);                                   // tracer.rs:431
                            let mut n2 = 1.;
                            let mut other: i64 = -1;
                            if inside {
                                for (ix, oo) in self.objects.iter().enumerate() {           // tracer.rs:432-439
                                    if ix != index && oo.contains(&point) {
                                        if let Some(m) = oo.get_material() { n2 = m.refractive_index; other = ix as i64; break; }
                                    }
                                }
                            } else {
                                n2 = material.refractive_index;
                            }
                            let (reflected, orefracted, reflectance) = ray.refract(&point, &normal, *n, n2); // tracer.rs:444-449
                            let ro = reflected.get_origin();
                            let rd = reflected.get_direction();
                            write!(rec, "\"contains_origin\":{},\"other_object\":{},\"n2\":{},\"reflectance\":{},\"reflected\":{{\"origin\":{},\"direction\":{}}},",
                                   inside, other, num(n2), num(reflectance), p2(&ro), v2(rd.x, rd.y)).unwrap();
                            self.lines.push((o, *color));
                            self.lines.push((ro, *color));
                            let refl = reflectance as f32;
                            let omrefl = 1. - refl;
                            next.push((reflected, [color[0] * refl, color[1] * refl, color[2] * refl, color[3]], *n, path << 1));
                            match orefracted {
                                Some(refracted) => {
                                    let fo = refracted.get_origin();
                                    let fd = refracted.get_direction();
                                    write!(rec, "\"refracted\":{{\"origin\":{},\"direction\":{}}}", p2(&fo), v2(fd.x, fd.y)).unwrap();
                                    next.push((refracted, [color[0] * omrefl, color[1] * omrefl, color[2] * omrefl, color[3]], n2, (path << 1) | 1));
                                }
                                None => rec.push_str("\"refracted\":null"),
                            }
                        } else {
                            let reflected = ray.reflect(&point, &normal);                   // tracer.rs:477
                            let ro = reflected.get_origin();
                            let rd = reflected.get_direction();
                            write!(rec, "\"mirror\":true,\"reflected\":{{\"origin\":{},\"direction\":{}}}", p2(&ro), v2(rd.x, rd.y)).unwrap();
                            self.lines.push((o, *color));
                            self.lines.push((point, *color));
                            next.push((reflected, *color, *n, path << 1));
                        }
                    }
                    None => {
                        match ray.intersect(&self.canvas) {                                  // tracer.rs:484-486
                            Some(hit) => {
                                let f = hit.get_first().0;
                                write!(rec, "\"hit_object\":-1,\"canvas_first\":{}", p2(&f)).unwrap();
                                self.lines.push((o, *color));
                                self.lines.push((f, *color));
                            }
                            None => rec.push_str("\"hit_object\":-1,\"canvas_first\":null"),
                        }
                    }
                }
                rec.push('}');
                if self.n_steps > 0 { self.steps.push_str(",\n"); }
                self.steps.push_str(&rec);
                self.n_steps += 1;
            }
            cur = next;
        }
    }
}

fn dump(scene_path: &str, out_dir: &str) {
    let text = std::fs::read_to_string(scene_path).expect("cannot read the scene file");
    let aspect = 16.0 / 9.0;
    let canvas = Rect::from_tlbr(1., -aspect, -1., aspect);           // sub_render_pass.rs:156
    let mut tracer = Tracer::new(&canvas);
    tracer.load(&text);                                               // tracer.rs:190-204
    tracer.enable_tile_map(false);                                    // the all-objects loop, tracer.rs:412-424
    tracer.chunk_size = 100;
    // optional header comment of the scene file: `// max_bounce = 64`
    for line in text.lines() {
        if let Some(rest) = line.trim().strip_prefix("// max_bounce =") { tracer.max_bounce = rest.trim().parse().unwrap(); }
    }
    let objects: Vec<&Object> = tracer.object_iterator().collect();
    let mut out = String::new();
    write!(out, "{{\"scene\":{:?},\"collision2d\":\"a7b471b54a940622f658a16177d63925536f16c3\",\"max_bounce\":{},\"cutoff_color\":{},\"canvas_tlbr\":[1.0,{},-1.0,{}],\n\"lights\":[",
           scene_path, tracer.max_bounce, col(&tracer.cutoff_color), num(-aspect), num(aspect)).unwrap();
    let mut walk = Walk { objects: objects.clone(), canvas, cutoff: tracer.cutoff_color, max_bounce: tracer.max_bounce,
                          steps: String::new(), n_steps: 0, lines: Vec::new() };
    let mut ref_lines: Vec<(P2, Color)> = Vec::new();
    let mut ray_id = 0usize;
    let mut rays_json = String::new();
    for (li, light) in tracer.light_iterator().enumerate() {
        // tracer.rs:280-287: last object that contains the light and has a material
        let mut n0 = 1.;
        for obj in objects.iter() {
            if obj.contains(&light.get_origin()) { if let Some(m) = obj.material_opt { n0 = m.refractive_index; } }
        }
        if li > 0 { out.push(','); }
        write!(out, "{{\"origin\":{},\"color\":{},\"num_rays\":{},\"start_medium\":{}}}",
               p2(&light.get_origin()), col(&light.get_color()), light.get_rays().len(), num(n0)).unwrap();
        for ray in light.get_rays() {
            let (o, d) = (ray.get_origin(), ray.get_direction());
            if ray_id > 0 { rays_json.push_str(",\n"); }
            write!(rays_json, "[{},{},{},{},{}]", li, num(o.x), num(o.y), num(d.x), num(d.y)).unwrap();
            tracer.trace(&mut ref_lines, ray, light.get_color(), n0, tracer.max_bounce);   // the reference, unmodified
            walk.primary(ray_id, ray, light.get_color(), n0);
            ray_id += 1;
        }
    }
    // the walk IS the reference's trace: same vertex pairs, same order
    assert_eq!(walk.lines.len(), ref_lines.len(), "probe walk and Tracer::trace disagree on the number of vertices");
    for (a, b) in walk.lines.iter().zip(ref_lines.iter()) {
        assert!(a.0 == b.0 && a.1 == b.1, "probe walk and Tracer::trace disagree: {:?} vs {:?}", a, b);
    }
    write!(out, "],\n\"rays\":[\n{}\n],\n\"steps\":[\n{}\n],\n\"segments\":[\n", rays_json, walk.steps).unwrap();
    for (k, pair) in ref_lines.chunks(2).enumerate() {
        if k > 0 { out.push_str(",\n"); }
        write!(out, "[{},{},{},{},{},{},{},{}]", num(pair[0].0.x), num(pair[0].0.y), num(pair[1].0.x), num(pair[1].0.y),
               num32(pair[0].1[0]), num32(pair[0].1[1]), num32(pair[0].1[2]), num32(pair[0].1[3])).unwrap();
    }
    out.push_str("\n]}\n");
    let stem = std::path::Path::new(scene_path).file_stem().unwrap().to_string_lossy().to_string();
    let path = format!("{}/ref_{}.json", out_dir, stem);
    std::fs::write(&path, out).expect("cannot write the golden file");
    println!("{}: {} primary rays, {} steps, {} segments", path, ray_id, walk.n_steps, ref_lines.len() / 2);
}

fn main() {
    let args: Vec<String> = std::env::args().collect();
    if args.len() < 3 { eprintln!("usage: dump_golden <out_dir> <scene.ron> [...]"); std::process::exit(2); }
    for scene in &args[2..] { dump(scene, &args[1]); }
}
