//! src/light_garden/cuda.rs -- the CUDA path of Tracer::trace_all (tracer.rs:276-358) and of the LineList pass of
//! Renderer::render (renderer.rs:164-188,431; sub_render_pass.rs:188-212), through lightgarden-cuda-sys
//! (include/light_garden_b200.h).  Added to the app by rust/patches/0001..0003; nothing else of the app changes.
//!
//! The Tracer keeps objects / lights / max_bounce / cutoff_color / canvas_bounds and every scene-edit method exactly as
//! they are.  `Object.moved` (object.rs:54) is already the dirty bit that tells when the scene has to be sent again.
//!
//! NOT COMPILED in the repository this file comes from (no Rust toolchain there): accessor names of collision2d types
//! (`get_a`, `get_b`, `points`, `get_rotation`) follow their use in the reference (drawer.rs:23-99, object.rs) and may
//! need a touch-up against collision2d@a7b471b5.  The PODs and entry points are generated from the C header and are
//! held to it by tests/test_rust_sys.py.
use crate::light_garden::*;
use lightgarden_cuda_sys as cu;
use std::ffi::CStr;

pub struct CudaPath {
    ctx: *mut cu::lg_ctx,
    size: (u32, u32),
    pub frame16: Vec<u16>, // the Rgba16Float frame, ready for queue.write_texture
}
// the context is only touched from the winit event-loop thread, like the Tracer itself (framework.rs:180-258)
unsafe impl Send for CudaPath {}

fn rot(m: &Rot2) -> [f64; 4] {
    let m = m.matrix(); // column-major [m11, m21, m12, m22], what serde writes (default.ron:24-29)
    [m[(0, 0)], m[(1, 0)], m[(0, 1)], m[(1, 1)]]
}
fn node(kind: i32) -> cu::LgGeoNode {
    cu::LgGeoNode { kind, op: 0, child_a: -1, child_b: -1, p: [0.; 8], rot: [1., 0., 0., 1.] }
}

/// Geo tree -> LgGeoNode array (mirrors light_garden_b200/scene.py::_push_geo); returns the root's index.
fn push_geo(geo: &Geo, nodes: &mut Vec<cu::LgGeoNode>) -> i32 {
    let mut n;
    match geo {
        Geo::GeoCircle(c) => { n = node(cu::LG_GEO_CIRCLE); n.p[..3].copy_from_slice(&[c.origin.x, c.origin.y, c.radius]); }
        Geo::GeoRect(r) => { n = node(cu::LG_GEO_RECT); n.p[..4].copy_from_slice(&[r.origin.x, r.origin.y, r.width, r.height]);
                             n.rot = rot(&r.rotation); }
        Geo::GeoLineSegment(s) => { n = node(cu::LG_GEO_SEGMENT); let (a, b) = (s.get_a(), s.get_b());
                                    n.p[..4].copy_from_slice(&[a.x, a.y, b.x, b.y]); }
        Geo::GeoCubicBezier(cb) => { n = node(cu::LG_GEO_BEZIER);
                                     for (k, q) in cb.points.iter().enumerate() { n.p[2 * k] = q.x; n.p[2 * k + 1] = q.y; } }
        Geo::GeoEllipse(e) => { n = node(cu::LG_GEO_ELLIPSE); n.p[..4].copy_from_slice(&[e.origin.x, e.origin.y, e.a, e.b]);
                                n.rot = rot(&e.rot); }
        Geo::GeoLogic(l) => {
            n = node(cu::LG_GEO_LOGIC);
            n.op = match l.op { LogicOp::And => cu::LG_OP_AND, LogicOp::Or => cu::LG_OP_OR, LogicOp::AndNot => cu::LG_OP_ANDNOT };
            let o = l.get_origin(); n.p[0] = o.x; n.p[1] = o.y; n.rot = rot(&l.get_rotation());
            let ix = nodes.len(); nodes.push(n);
            let (a, b) = (push_geo(&l.get_a(), nodes), push_geo(&l.get_b(), nodes));
            nodes[ix].child_a = a; nodes[ix].child_b = b;
            return ix as i32;
        }
        Geo::GeoConvexPolygon(cp) => { // header node + continuation nodes of up to four local-frame hull vertices
            n = node(cu::LG_GEO_POLYGON);
            let o = cp.get_origin(); n.p[0] = o.x; n.p[1] = o.y; n.rot = rot(&cp.get_rotation());
            let pts = cp.points(); n.op = pts.len() as i32;
            let ix = nodes.len(); nodes.push(n);
            let mut prev = ix;
            for chunk in pts.chunks(4) {
                let mut c = node(cu::LG_GEO_POINTS); c.op = chunk.len() as i32;
                for (q, v) in chunk.iter().enumerate() { c.p[2 * q] = v.x; c.p[2 * q + 1] = v.y; }
                nodes[prev].child_a = nodes.len() as i32; prev = nodes.len(); nodes.push(c);
            }
            return ix as i32;
        }
        _ => panic!("Geo variant not supported by the CUDA path (MCircle, Ray, Point)"),
    }
    nodes.push(n);
    (nodes.len() - 1) as i32
}

fn object_pod(o: &Object, nodes: &mut Vec<cu::LgGeoNode>) -> cu::LgObject {
    cu::LgObject { root: push_geo(&o.get_geometry(), nodes), has_material: o.material_opt.is_some() as i32,
                   refractive_index: o.material_opt.map(|m| m.refractive_index).unwrap_or(0.) }
}

fn light_pod(l: &Light) -> cu::LgLight {
    let mut p = cu::LgLight { kind: 0, flags: 0, num_rays: l.get_num_rays() as u64, color: l.get_color(), position: [0.; 2],
                              b: [0.; 2], spot_angle: 0., spot_direction: [0.; 2], start_medium: 0. };
    match l {
        Light::PointLight(pl) => { p.kind = cu::LG_LIGHT_POINT; let o = pl.get_origin(); p.position = [o.x, o.y]; }
        Light::SpotLight(sl) => { p.kind = cu::LG_LIGHT_SPOT; let o = sl.get_origin(); p.position = [o.x, o.y];
                                  p.spot_angle = sl.spot_angle; let d = sl.get_spot_direction(); p.spot_direction = [d.x, d.y]; }
        Light::DirectionalLight(dl) => { p.kind = cu::LG_LIGHT_DIRECTIONAL; let s = dl.get_start();
                                         let (a, b) = (s.get_a(), s.get_b()); p.position = [a.x, a.y]; p.b = [b.x, b.y]; }
    }
    p
}

impl CudaPath {
    pub fn new() -> Self {
        let mut ctx = std::ptr::null_mut();
        // F64: the reference's own width (collision2d Float = f64); LG_PRECISION_F32 is the throughput mode
        let rc = unsafe { cu::lg_create(0, cu::LG_PRECISION_F64, &mut ctx) };
        assert!(rc == cu::LG_OK, "light_garden_b200: lg_create failed ({rc}): no CUDA device?  (there is no CPU fallback)");
        let s = CudaPath { ctx, size: (0, 0), frame16: Vec::new() };
        s.check(unsafe { cu::lg_tile_map_enable(ctx, 1) }); // the app starts with its TileMap enabled (tile_map.rs:61)
        s
    }
    fn check(&self, rc: i32) {
        if rc != cu::LG_OK {
            let msg = unsafe { CStr::from_ptr(cu::lg_last_error(self.ctx)) }.to_string_lossy().into_owned();
            panic!("light_garden_b200: {rc}: {msg}"); // the reference panics on its own failures too (tracer.rs:192)
        }
    }
    pub fn enable_tile_map(&mut self, enable: bool) { self.check(unsafe { cu::lg_tile_map_enable(self.ctx, enable as i32) }); }

    /// Everything Tracer::trace_all reads (tracer.rs:4-17), incl. the drawing object / light it chains (tracer.rs:279-281).
    pub fn sync(&mut self, objects: &[Object], drawing_object: Option<&Object>, lights: &[Light], drawing_light: Option<&Light>,
                max_bounce: u32, cutoff_color: Color, canvas: &Rect) {
        let mut nodes = Vec::new();
        let objs: Vec<cu::LgObject> = objects.iter().map(|o| object_pod(o, &mut nodes)).collect();
        let (hw, hh) = (canvas.width / 2., canvas.height / 2.);
        let prm = cu::LgTraceParams { max_bounce, cutoff_color, _pad: 0,
            canvas_tlbr: [canvas.origin.y + hh, canvas.origin.x - hw, canvas.origin.y - hh, canvas.origin.x + hw] };
        let ls: Vec<cu::LgLight> = lights.iter().chain(drawing_light.into_iter()).map(light_pod).collect();
        unsafe {
            self.check(cu::lg_scene_set(self.ctx, objs.as_ptr(), objs.len() as u32, nodes.as_ptr(), nodes.len() as u32, &prm));
            match drawing_object {
                Some(o) => { let mut dn = Vec::new(); let pod = object_pod(o, &mut dn);
                             self.check(cu::lg_drawing_object_set(self.ctx, &pod, dn.as_ptr(), dn.len() as u32)); }
                None => self.check(cu::lg_drawing_object_set(self.ctx, std::ptr::null(), std::ptr::null(), 0)),
            }
            self.check(cu::lg_lights_set(self.ctx, ls.as_ptr(), ls.len() as u32));
        }
    }

    /// Tracer::trace_all's rayon fan-out (tracer.rs:288-330): the segments stay on the device.
    pub fn trace(&mut self) -> cu::LgTraceStats {
        let mut st: cu::LgTraceStats = unsafe { std::mem::zeroed() };
        self.check(unsafe { cu::lg_trace(self.ctx, &mut st) });
        st
    }

    /// The vertex list itself, in the reference's order, for callers that want it (tests, export).
    pub fn segments(&mut self) -> Vec<(P2, Color)> {
        unsafe {
            self.check(cu::lg_tags_enable(self.ctx, 1));
            self.trace();
            let mut n = 0u64;
            self.check(cu::lg_segments_count(self.ctx, &mut n));
            let mut seg: Vec<cu::LgSegment> = Vec::with_capacity(n as usize);
            let mut tag: Vec<cu::LgSegmentTag> = Vec::with_capacity(n as usize);
            let mut got = 0u64;
            self.check(cu::lg_segments_read(self.ctx, seg.as_mut_ptr(), tag.as_mut_ptr(), std::ptr::null_mut(), n, &mut got));
            seg.set_len(got as usize); tag.set_len(got as usize);
            self.check(cu::lg_tags_enable(self.ctx, 0));
            let mut order: Vec<usize> = (0..seg.len()).collect();
            order.sort_by_key(|&i| (tag[i].ray, tag[i].generation, tag[i].path)); // light -> ray -> generation -> queue order
            order.iter().flat_map(|&i| { let s = seg[i];
                [(P2::new(s.a[0] as f64, s.a[1] as f64), s.color), (P2::new(s.b[0] as f64, s.b[1] as f64), s.color)] }).collect()
        }
    }

    /// sub_rpass_lines.update_vertex_buffer + render into the Rgba16Float target (renderer.rs:164-188,431): the traced
    /// segments on the device (or the string mods), then the host lines (control polygons, grid, drawer overlays:
    /// tracer.rs:342-349, mod.rs:692); leaves the frame in self.frame16.
    pub fn line_pass(&mut self, width: u32, height: u32, string_mods: Option<&[StringMod]>, host_lines: &[(P2, Color)]) {
        unsafe {
            if self.size != (width, height) {
                self.check(cu::lg_image_configure(self.ctx, width, height));
                self.size = (width, height);
            }
            self.check(cu::lg_image_clear(self.ctx, 1.0)); // LoadOp::Clear(BLACK), renderer.rs:174-177
            match string_mods {
                Some(sms) => for s in sms {
                    let (pod, rules) = string_mod_pod(s);
                    match s.nested.as_ref() { // StringMod::draw, string_mod.rs:152-158
                        Some(inner) => { let (ipod, irules) = string_mod_pod(inner);
                            self.check(cu::lg_string_mod_nested(self.ctx, &pod, &ipod, irules.as_ptr(), irules.len() as u32, std::ptr::null_mut())); }
                        None => self.check(cu::lg_string_mod(self.ctx, &pod, rules.as_ptr(), rules.len() as u32, 0, 0, std::ptr::null_mut())),
                    }
                },
                None => self.check(cu::lg_accumulate_traced(self.ctx, std::ptr::null_mut())),
            }
            let pairs: Vec<cu::LgVertexPair> = host_lines.chunks_exact(2).map(|v| cu::LgVertexPair {
                a: [v[0].0.x, v[0].0.y], b: [v[1].0.x, v[1].0.y], color_a: v[0].1, color_b: v[1].1 }).collect();
            self.check(cu::lg_accumulate_segments(self.ctx, pairs.as_ptr(), pairs.len() as u64, std::ptr::null_mut()));
            self.frame16.resize((width * height * 4) as usize, 0);
            self.check(cu::lg_image_read(self.ctx, cu::LG_RGBA16F, self.frame16.as_mut_ptr() as *mut _, 0));
        }
    }

    /// LightGarden.color_state_descriptor.blend as the GUI edits it (gui/settings.rs:59-127); false = the parallel pass
    /// cannot reproduce that state (its result depends on the fragment order) and the previous one stays in force.
    pub fn set_blend(&mut self, b: &wgpu::BlendState) -> bool {
        let comp = |c: &wgpu::BlendComponent| cu::LgBlendComponent { src_factor: c.src_factor as i32, dst_factor: c.dst_factor as i32,
                                                                     operation: c.operation as i32 };
        let st = cu::LgBlendState { color: comp(&b.color), alpha: comp(&b.alpha), constant: [0.; 4] };
        unsafe { cu::lg_blend_set(self.ctx, &st) == cu::LG_OK }
    }

    /// Renderer::make_screenshot's read-back + per-pixel conversion (renderer.rs:250-328): same bytes, same row padding.
    pub fn screenshot(&mut self, render_to_texture: bool, padded_bytes_per_row: usize, dst: &mut [u8]) {
        let fmt = if render_to_texture { cu::LG_BGRA8_GAMMA } else { cu::LG_BGRA8_SRGB };
        self.check(unsafe { cu::lg_image_read(self.ctx, fmt, dst.as_mut_ptr() as *mut _, padded_bytes_per_row) });
    }
}

impl Drop for CudaPath {
    fn drop(&mut self) { unsafe { cu::lg_destroy(self.ctx); } }
}

fn string_mod_pod(s: &StringMod) -> (cu::LgStringMod, Vec<cu::LgModRemColor>) {
    let (curve, p) = match s.init_curve { // string_mod.rs:182-188
        Curve::Circle => (cu::LG_CURVE_CIRCLE, [0.; 4]),
        Curve::ComplexExp { c } => (cu::LG_CURVE_COMPLEX_EXP, [c.re, c.im, 0., 0.]),
        Curve::Hypotrochoid { r, s, d } => (cu::LG_CURVE_HYPOTROCHOID, [r as f64, s as f64, d as f64, 0.]),
        Curve::Lissajous { a, b, delta } => (cu::LG_CURVE_LISSAJOUS, [a as f64, b as f64, delta, 0.]),
    };
    let mode = match s.mode { StringModMode::Add => cu::LG_SM_ADD, StringModMode::Mul => cu::LG_SM_MUL,
                              StringModMode::Pow => cu::LG_SM_POW, StringModMode::Base => cu::LG_SM_BASE };
    (cu::LgStringMod { modulo: s.modulo, num: s.num, turns: s.turns, mode, curve, color: s.color, curve_p: p },
     s.modulo_colors.iter().map(|m| cu::LgModRemColor { modulo: m.modulo, rem: m.rem, color: m.color }).collect())
}
