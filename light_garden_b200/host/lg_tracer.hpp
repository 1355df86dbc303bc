// lg_tracer.hpp — C++ host-side mirror of the reference's scene / trace API on top of the C ABI
// (include/light_garden_b200.h).  Header only; link against liblight_garden_b200.so.
//
// Same names, argument order, defaults and semantics as the Rust code it stands in for:
//   Object::{new_mirror,new_curved_mirror,new_circle,new_rect,new_lens,new_geo}   src/light_garden/object.rs:58-113
//   Material{refractive_index = 1.2}                                              src/light_garden/object.rs:437-447
//   PointLight::new / SpotLight::new / DirectionalLight::new                      src/light_garden/light.rs:91,153,205
//   Tracer{max_bounce = 5, cutoff_color = [0.001;4], chunk_size = 100, canvas_bounds}, push_object, push_light,
//   clear, replace_object, remove_object, remove_light, resize, trace_all()       src/light_garden/tracer.rs:4-358
//   Renderer{render, make_screenshot, resize} + SubRenderPass::update_vertex_buffer  src/renderer.rs:164-431, src/sub_render_pass.rs:188-212
//   StringMod{modulo = 5, num = 1, color = [1;4], turns = 1, Circle, Mul}, ModRemColor, Curve    src/light_garden/string_mod.rs:4-31,160-188
// Where the reference panics (tracer.rs:192, framework.rs:44) this throws lg::Error carrying lg_last_error().
// All computation happens in the CUDA library; this file only holds and flattens data.
#pragma once
#include <algorithm>
#include <array>
#include <cstdint>
#include <memory>
#include <numeric>
#include <optional>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../../include/light_garden_b200.h"

namespace lg {

struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string &m) : std::runtime_error(m), code(c) {}
};

struct P2 {
  double x = 0, y = 0;
};
using V2 = P2;
using Color = std::array<float, 4>; // light.rs:7
using Rot2 = std::array<double, 4>; // nalgebra Rotation2 as serde writes it: [m11, m21, m12, m22]
inline Rot2 rot2_identity() { return {1, 0, 0, 1}; }

enum class LogicOp { And = LG_OP_AND, Or = LG_OP_OR, AndNot = LG_OP_ANDNOT };

// collision2d Geo (the variants the BASELINE configs use)
struct Geo {
  int kind = LG_GEO_CIRCLE;
  std::array<double, 8> p{};
  Rot2 rot = rot2_identity();
  LogicOp op = LogicOp::And;
  std::shared_ptr<Geo> a, b;
  std::vector<P2> points; // ConvexPolygon: hull vertices, local frame

  static Geo circle(P2 origin, double radius) {
    Geo g;
    g.kind = LG_GEO_CIRCLE;
    g.p = {origin.x, origin.y, radius};
    return g;
  }
  static Geo rect(P2 origin, Rot2 rotation, double width, double height) {
    Geo g;
    g.kind = LG_GEO_RECT;
    g.p = {origin.x, origin.y, width, height};
    g.rot = rotation;
    return g;
  }
  static Geo line_segment(P2 a, P2 b) { // LineSegment::from_ab
    Geo g;
    g.kind = LG_GEO_SEGMENT;
    g.p = {a.x, a.y, b.x, b.y};
    return g;
  }
  static Geo cubic_bezier(const std::array<P2, 4> &pts) {
    Geo g;
    g.kind = LG_GEO_BEZIER;
    for (int k = 0; k < 4; ++k) g.p[2 * k] = pts[k].x, g.p[2 * k + 1] = pts[k].y;
    return g;
  }
  static Geo ellipse(P2 origin, double a, double b, Rot2 rotation = rot2_identity()) { // object.rs:38-45
    Geo g;
    g.kind = LG_GEO_ELLIPSE;
    g.p = {origin.x, origin.y, a, b};
    g.rot = rotation;
    return g;
  }
  // ConvexPolygon::new_convex_hull (object.rs:34-36): Andrew's monotone chain, counter-clockwise from the lowest
  // (x, then y) point, collinear points dropped (ORACLE.md §3.8)
  static Geo convex_polygon(std::vector<P2> pts, P2 origin = {}, Rot2 rotation = rot2_identity()) {
    std::sort(pts.begin(), pts.end(), [](P2 a, P2 b) { return a.x < b.x || (a.x == b.x && a.y < b.y); });
    pts.erase(std::unique(pts.begin(), pts.end(), [](P2 a, P2 b) { return a.x == b.x && a.y == b.y; }), pts.end());
    auto turn = [](P2 o, P2 a, P2 b) { return (a.x - o.x) * (b.y - o.y) - (a.y - o.y) * (b.x - o.x); };
    std::vector<P2> lower, upper;
    for (const P2 &q : pts) {
      while (lower.size() >= 2 && turn(lower[lower.size() - 2], lower.back(), q) <= 0.0) lower.pop_back();
      lower.push_back(q);
    }
    for (auto it = pts.rbegin(); it != pts.rend(); ++it) {
      while (upper.size() >= 2 && turn(upper[upper.size() - 2], upper.back(), *it) <= 0.0) upper.pop_back();
      upper.push_back(*it);
    }
    Geo g;
    g.kind = LG_GEO_POLYGON;
    g.p = {origin.x, origin.y};
    g.rot = rotation;
    if (!lower.empty()) lower.pop_back();
    if (!upper.empty()) upper.pop_back();
    g.points = lower;
    g.points.insert(g.points.end(), upper.begin(), upper.end());
    if (g.points.size() < 3 || g.points.size() > LG_POLYGON_MAX_VERTICES)
      throw Error(LG_ERR_INVALID, "convex polygon needs 3..32 hull vertices");
    return g;
  }
  static Geo logic(LogicOp op, Geo a, Geo b, P2 origin, Rot2 rotation) { // Logic::new
    Geo g;
    g.kind = LG_GEO_LOGIC;
    g.op = op;
    g.p = {origin.x, origin.y};
    g.rot = rotation;
    g.a = std::make_shared<Geo>(std::move(a));
    g.b = std::make_shared<Geo>(std::move(b));
    return g;
  }
};

struct Rect { // collision2d Rect as the tracer uses it for canvas_bounds
  double top = 1, left = -1, bottom = -1, right = 1;
  static Rect from_tlbr(double t, double l, double b, double r) { return {t, l, b, r}; } // sub_render_pass.rs:156
};

struct Material {
  double refractive_index = 1.2;
};

struct Object {
  Geo geo;
  std::optional<Material> material_opt;
  bool curved_mirror = false;
  bool moved = true;

  static Object new_mirror(P2 a, P2 b) { return {Geo::line_segment(a, b), std::nullopt}; }
  static Object new_curved_mirror(const std::array<P2, 4> &cubic) {
    Object o{Geo::cubic_bezier(cubic), std::nullopt};
    o.curved_mirror = true;
    return o;
  }
  static Object new_circle(P2 origin, double radius) { return {Geo::circle(origin, radius), Material{}}; }
  static Object new_rect(P2 origin, double width, double height) {
    return {Geo::rect(origin, rot2_identity(), width, height), Material{}};
  }
  static Object new_lens(P2 origin, double radius, double distance) { // Lens::new, object.rs:393-410
    return {Geo::logic(LogicOp::And, Geo::circle({distance * 0.5, 0}, radius), Geo::circle({-distance * 0.5, 0}, radius),
                       origin, rot2_identity()),
            Material{}};
  }
  static Object new_ellipse(P2 origin, double a, double b) { return {Geo::ellipse(origin, a, b), Material{}}; }
  static Object new_convex_polygon(const std::vector<P2> &points) { return {Geo::convex_polygon(points), Material{}}; }
  static Object new_geo(Geo g) { return {std::move(g), Material{}}; }
  std::optional<Material> get_material() const { return material_opt; }
};

struct Light {
  LgLight pod{};
  static Light point(P2 position, size_t num_rays, Color color) { // PointLight::new
    Light l;
    l.pod.kind = LG_LIGHT_POINT;
    l.pod.num_rays = num_rays;
    std::copy(color.begin(), color.end(), l.pod.color);
    l.pod.position[0] = position.x, l.pod.position[1] = position.y;
    return l;
  }
  static Light spot(P2 position, double spot_angle, V2 spot_direction, size_t num_rays, Color color) { // SpotLight::new
    Light l = point(position, num_rays, color);
    l.pod.kind = LG_LIGHT_SPOT;
    l.pod.spot_angle = spot_angle;
    l.pod.spot_direction[0] = spot_direction.x, l.pod.spot_direction[1] = spot_direction.y;
    return l;
  }
  static Light directional(Color color, size_t num_rays, P2 a, P2 b) { // DirectionalLight::new(color, n, start)
    Light l = point(a, num_rays, color);
    l.pod.kind = LG_LIGHT_DIRECTIONAL;
    l.pod.b[0] = b.x, l.pod.b[1] = b.y;
    return l;
  }
};

class Tracer {
public:
  uint32_t max_bounce = 5;                               // tracer.rs:37
  Color cutoff_color{0.001f, 0.001f, 0.001f, 0.001f};   // tracer.rs:38
  size_t chunk_size = 100;                               // tracer.rs:39 (rayon chunking; unused on the device)
  Rect canvas_bounds;
  LgTraceStats last_stats{};

  explicit Tracer(const Rect &canvas, int device = 0, int precision = LG_PRECISION_F32) : canvas_bounds(canvas) {
    int rc = lg_create(device, precision, &ctx_);
    if (rc != LG_OK) throw Error(rc, "lg_create failed: no CUDA device (there is no CPU fallback)");
  }
  ~Tracer() {
    if (ctx_) lg_destroy(ctx_);
  }
  Tracer(const Tracer &) = delete;
  Tracer &operator=(const Tracer &) = delete;

  lg_ctx *context() { return ctx_; }
  void clear() { objects_.clear(), lights_.clear(); }
  void clear_objects() { objects_.clear(); }
  void push_object(Object o) { objects_.push_back(std::move(o)); }
  void push_light(Light l) { lights_.push_back(std::move(l)); }
  Object &index_object(size_t ix) { return objects_.at(ix); }
  Light &index_light(size_t ix) { return lights_.at(ix); }
  void replace_object(size_t ix, Object o) { objects_.at(ix) = std::move(o); }
  void remove_object(size_t ix) { objects_.erase(objects_.begin() + (long)ix); }
  void remove_light(size_t ix) { lights_.erase(lights_.begin() + (long)ix); }
  const std::vector<Object> &object_iterator() const { return objects_; }
  const std::vector<Light> &light_iterator() const { return lights_; }
  void resize(const Rect &bounds) { canvas_bounds = bounds; }
  // Tracer::enable_tile_map / tile_map_enabled (tracer.rs:126-146)
  void enable_tile_map(bool enable) {
    check(lg_tile_map_enable(ctx_, enable ? 1 : 0));
    tile_map_enabled_ = enable;
  }
  bool tile_map_enabled() const { return tile_map_enabled_; }

  // Tracer::trace_all -> Vec<(P2, Color)>: two vertices per segment, in the reference's order
  // (light -> ray -> generation -> queue order), followed by the curved mirrors' control lines (tracer.rs:342-346).
  std::vector<std::pair<P2, Color>> trace_all() {
    upload();
    check(lg_tags_enable(ctx_, 1));
    check(lg_trace(ctx_, &last_stats));
    uint64_t n = 0;
    check(lg_segments_count(ctx_, &n));
    std::vector<LgSegment> seg(n);
    std::vector<LgSegmentTag> tag(n);
    uint64_t got = 0;
    check(lg_segments_read(ctx_, seg.data(), tag.data(), nullptr, n, &got));
    std::vector<uint64_t> order(n);
    std::iota(order.begin(), order.end(), 0);
    std::sort(order.begin(), order.end(), [&](uint64_t i, uint64_t j) {
      const LgSegmentTag &a = tag[i], &b = tag[j];
      if (a.ray != b.ray) return a.ray < b.ray;
      if (a.generation != b.generation) return a.generation < b.generation;
      return a.path < b.path;
    });
    std::vector<std::pair<P2, Color>> lines;
    lines.reserve(2 * n + 8);
    for (uint64_t k : order) {
      const LgSegment &s = seg[k];
      Color c{s.color[0], s.color[1], s.color[2], s.color[3]};
      lines.push_back({P2{s.a[0], s.a[1]}, c});
      lines.push_back({P2{s.b[0], s.b[1]}, c});
    }
    const Color red{1.f, 0.f, 0.f, 1.f};
    for (const Object &o : objects_) {
      if (!o.curved_mirror) continue;
      for (int k = 0; k < 3; ++k) {
        lines.push_back({P2{o.geo.p[2 * k], o.geo.p[2 * k + 1]}, red});
        lines.push_back({P2{o.geo.p[2 * k + 2], o.geo.p[2 * k + 3]}, red});
      }
    }
    return lines;
  }

  // scene -> device (lg_scene_set + lg_lights_set)
  void upload() {
    std::vector<LgGeoNode> nodes;
    std::vector<LgObject> objs;
    for (const Object &o : objects_) {
      LgObject po{};
      po.root = push_geo(o.geo, nodes);
      po.has_material = o.material_opt.has_value();
      po.refractive_index = o.material_opt ? o.material_opt->refractive_index : 0.0;
      objs.push_back(po);
    }
    LgTraceParams prm{};
    prm.max_bounce = max_bounce;
    std::copy(cutoff_color.begin(), cutoff_color.end(), prm.cutoff_color);
    prm.canvas_tlbr[0] = canvas_bounds.top, prm.canvas_tlbr[1] = canvas_bounds.left;
    prm.canvas_tlbr[2] = canvas_bounds.bottom, prm.canvas_tlbr[3] = canvas_bounds.right;
    check(lg_scene_set(ctx_, objs.data(), (uint32_t)objs.size(), nodes.data(), (uint32_t)nodes.size(), &prm));
    std::vector<LgLight> ls;
    for (const Light &l : lights_) ls.push_back(l.pod);
    check(lg_lights_set(ctx_, ls.data(), (uint32_t)ls.size()));
  }

  void check(int rc) {
    if (rc != LG_OK) throw Error(rc, lg_last_error(ctx_));
  }

private:
  static int32_t push_geo(const Geo &g, std::vector<LgGeoNode> &nodes) {
    LgGeoNode n{};
    n.kind = g.kind;
    n.child_a = n.child_b = -1;
    std::copy(g.p.begin(), g.p.end(), n.p);
    std::copy(g.rot.begin(), g.rot.end(), n.rot);
    if (g.kind == LG_GEO_POLYGON) { // header node + continuation nodes of four vertices each
      n.op = (int32_t)g.points.size();
      const size_t ix = nodes.size();
      nodes.push_back(n);
      size_t prev = ix;
      for (size_t v = 0; v < g.points.size(); v += 4) {
        LgGeoNode c{};
        c.kind = LG_GEO_POINTS;
        c.child_a = c.child_b = -1;
        c.rot[0] = c.rot[3] = 1.0;
        c.op = (int32_t)std::min<size_t>(4, g.points.size() - v);
        for (int q = 0; q < c.op; ++q) c.p[2 * q] = g.points[v + q].x, c.p[2 * q + 1] = g.points[v + q].y;
        nodes[prev].child_a = (int32_t)nodes.size();
        prev = nodes.size();
        nodes.push_back(c);
      }
      return (int32_t)ix;
    }
    if (g.kind != LG_GEO_LOGIC) {
      nodes.push_back(n);
      return (int32_t)nodes.size() - 1;
    }
    n.op = (int32_t)g.op;
    const size_t ix = nodes.size();
    nodes.push_back(n);
    const int32_t a = push_geo(*g.a, nodes), b = push_geo(*g.b, nodes);
    nodes[ix].child_a = a, nodes[ix].child_b = b;
    return (int32_t)ix;
  }
  lg_ctx *ctx_ = nullptr;
  bool tile_map_enabled_ = false;
  std::vector<Object> objects_;
  std::vector<Light> lights_;
};

// StringMod (src/light_garden/string_mod.rs:4-31,160-188): same fields and defaults as StringMod::new()
struct ModRemColor {
  uint64_t modulo = 1, rem = 0;
  Color color{1.f, 1.f, 1.f, 1.f};
};
enum class StringModMode { Add = LG_SM_ADD, Mul = LG_SM_MUL, Pow = LG_SM_POW, Base = LG_SM_BASE };
struct Curve {
  int kind = LG_CURVE_CIRCLE;
  std::array<double, 4> params{};
  static Curve Circle() { return {}; }
  static Curve ComplexExp(double re, double im) { return {LG_CURVE_COMPLEX_EXP, {re, im, 0, 0}}; }
  static Curve Hypotrochoid(uint64_t r, uint64_t s, uint64_t d) {
    return {LG_CURVE_HYPOTROCHOID, {(double)r, (double)s, (double)d, 0}};
  }
  static Curve Lissajous(uint64_t a, uint64_t b, double delta) { return {LG_CURVE_LISSAJOUS, {(double)a, (double)b, delta, 0}}; }
};
struct StringMod {
  uint64_t modulo = 5, num = 1;
  uint32_t pow = 0;
  Color color{1.f, 1.f, 1.f, 1.f};
  uint64_t turns = 1;
  Curve init_curve = Curve::Circle();
  StringModMode mode = StringModMode::Mul;
  std::vector<ModRemColor> modulo_colors;
  std::shared_ptr<StringMod> nested; // Option<Box<StringMod>>

  LgStringMod pod() const {
    LgStringMod s{};
    s.modulo = modulo, s.num = num, s.turns = turns;
    s.mode = (int32_t)mode, s.curve = init_curve.kind;
    std::copy(color.begin(), color.end(), s.color);
    std::copy(init_curve.params.begin(), init_curve.params.end(), s.curve_p);
    return s;
  }
  std::vector<LgModRemColor> rules() const {
    std::vector<LgModRemColor> r(modulo_colors.size());
    for (size_t k = 0; k < r.size(); ++k) {
      r[k].modulo = modulo_colors[k].modulo, r[k].rem = modulo_colors[k].rem;
      std::copy(modulo_colors[k].color.begin(), modulo_colors[k].color.end(), r[k].color);
    }
    return r;
  }
};

// The line pass (boundary B2): Renderer::render's LineList draw of the traced lines into the Rgba16Float target
// (src/renderer.rs:164-188,431; SubRenderPass::update_vertex_buffer + render, src/sub_render_pass.rs:188-212), the
// screenshot conversion (renderer.rs:190-328) and the frame hand-off.  Shares the tracer's context.
class Renderer {
public:
  uint32_t width, height;
  LgTraceStats last_stats{};

  Renderer(Tracer &tracer, uint32_t w, uint32_t h) : width(w), height(h), t_(tracer) {
    t_.check(lg_image_configure(t_.context(), w, h));
  }
  // SurfaceConfiguration change (renderer.rs:resize): a new target of the new size, cleared
  void resize(uint32_t w, uint32_t h) {
    width = w, height = h;
    t_.check(lg_image_configure(t_.context(), w, h));
  }
  // LoadOp::Clear(BLACK), renderer.rs:174-177
  void clear(float clear_alpha = 1.0f) { t_.check(lg_image_clear(t_.context(), clear_alpha)); }
  // Renderer::render: trace_all + update_vertex_buffer + the line pass, fused on the device
  const LgTraceStats &render() {
    t_.upload();
    t_.check(lg_tags_enable(t_.context(), 0));
    t_.check(lg_render(t_.context(), &last_stats));
    return last_stats;
  }
  // sub_rpass_lines.render for the segments of the last Tracer::trace_all
  const LgTraceStats &render_traced() {
    t_.check(lg_accumulate_traced(t_.context(), &last_stats));
    return last_stats;
  }
  // SubRenderPass::update_vertex_buffer(&lines) + render for a host LineList (vertex pairs): control lines, grid,
  // drawer overlays (tracer.rs:342-349, mod.rs:692)
  const LgTraceStats &render_lines(const std::vector<std::pair<P2, Color>> &lines) {
    std::vector<LgVertexPair> vp(lines.size() / 2);
    for (size_t k = 0; k < vp.size(); ++k) {
      const auto &a = lines[2 * k], &b = lines[2 * k + 1];
      vp[k].a[0] = a.first.x, vp[k].a[1] = a.first.y, vp[k].b[0] = b.first.x, vp[k].b[1] = b.first.y;
      std::copy(a.second.begin(), a.second.end(), vp[k].color_a);
      std::copy(b.second.begin(), b.second.end(), vp[k].color_b);
    }
    t_.check(lg_accumulate_segments(t_.context(), vp.data(), vp.size(), &last_stats));
    return last_stats;
  }
  // LightGarden::draw in Mode::StringMod (mod.rs:681-689: StringMod::draw, string_mod.rs:152-158) + the line pass
  const LgTraceStats &render_string_mod(const StringMod &sm) {
    const LgStringMod pod = sm.pod();
    if (sm.nested) {
      const LgStringMod inner = sm.nested->pod();
      const std::vector<LgModRemColor> irules = sm.nested->rules();
      t_.check(lg_string_mod_nested(t_.context(), &pod, &inner, irules.data(), (uint32_t)irules.size(), &last_stats));
    } else {
      const std::vector<LgModRemColor> r = sm.rules();
      t_.check(lg_string_mod(t_.context(), &pod, r.data(), (uint32_t)r.size(), 0, 0, &last_stats));
    }
    return last_stats;
  }
  std::vector<float> read_rgba32f() {
    std::vector<float> img((size_t)width * height * 4);
    t_.check(lg_image_read(t_.context(), LG_RGBA32F, img.data(), 0));
    return img;
  }
  // the Rgba16Float texture's bits (texture_renderer.rs:5)
  std::vector<uint16_t> read_rgba16f() {
    std::vector<uint16_t> img((size_t)width * height * 4);
    t_.check(lg_image_read(t_.context(), LG_RGBA16F, img.data(), 0));
    return img;
  }
  // Renderer::make_screenshot's pixels: [b, g, r, a] bytes, rows padded to `padded_bytes_per_row` (0 = tight;
  // renderer.rs:250-255 pads to 256); render_to_texture selects the fp16 gamma conversion (renderer.rs:313-328) or
  // the 8-bit sRGB surface (renderer.rs:207-209)
  std::vector<uint8_t> make_screenshot(bool render_to_texture = true, size_t padded_bytes_per_row = 0) {
    const size_t row = padded_bytes_per_row ? padded_bytes_per_row : (size_t)width * 4;
    std::vector<uint8_t> px(row * height);
    t_.check(lg_image_read(t_.context(), render_to_texture ? LG_BGRA8_GAMMA : LG_BGRA8_SRGB, px.data(), row));
    return px;
  }
  // the frame as importable device memory instead of a host copy (include/light_garden_b200.h: lg_image_export_fd)
  std::pair<int, uint64_t> export_fd(int32_t format = LG_RGBA16F) {
    int32_t fd = -1;
    uint64_t bytes = 0;
    t_.check(lg_image_export_fd(t_.context(), format, &fd, &bytes));
    return {fd, bytes};
  }
  void export_refresh(int32_t format = LG_RGBA16F) { t_.check(lg_image_export_refresh(t_.context(), format)); }

private:
  Tracer &t_;
};

} // namespace lg
