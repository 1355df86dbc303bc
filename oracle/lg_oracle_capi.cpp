// oracle/lg_oracle_capi.cpp — C entry points of the CPU oracle for ctypes.
// TEST INFRASTRUCTURE ONLY (tests/, __graft_entry__.smoke(), bench.py's
// cpu_baseline / --impl reference legs).  PARITY UNPINNED — see lg_oracle.hpp.
//
// Threading mirrors the reference: lights are processed sequentially and each
// light's rays are split into chunks of `chunk_size` (default 100) spread over
// the host threads, results concatenated in order
// (src/light_garden/tracer.rs:279-307, rayon par_chunks + collect + concat).
#include "lg_oracle.hpp"

#include <algorithm>
#include <chrono>
#include <limits>
#include <memory>
#ifdef _OPENMP
#include <omp.h>
#endif

using namespace lgo;

struct OracleScene {
  Scene base;
  SceneT<double> d;
  SceneT<float> f;
  TileMap tm; // Tracer::tile_map; built by lgo_tile_map_enable
};

struct OracleResult {
  std::vector<SegOut> segs;
  TraceCounters cnt;
  uint64_t primary_rays = 0;
  double seconds = 0;
};

template <class T> static const SceneT<T> &pick(const OracleScene &s);
template <> const SceneT<double> &pick<double>(const OracleScene &s) { return s.d; }
template <> const SceneT<float> &pick<float>(const OracleScene &s) { return s.f; }

template <class T>
static void trace_block(const OracleScene &sc, const LgRay *rays, uint64_t n, uint64_t id0, int chunk, int threads,
                        bool store, OracleResult &res) {
  const SceneT<T> &s = pick<T>(sc);
  if (chunk < 1) chunk = 100;
  int64_t nchunks = (int64_t)((n + chunk - 1) / chunk);
  std::vector<std::vector<SegOut>> parts(store ? nchunks : 0);
  std::vector<TraceCounters> cnts(nchunks);
#ifdef _OPENMP
  if (threads > 0) omp_set_num_threads(threads);
#endif
#pragma omp parallel for schedule(dynamic, 1)
  for (int64_t c = 0; c < nchunks; ++c) {
    uint64_t lo = (uint64_t)c * chunk, hi = std::min<uint64_t>(n, lo + chunk);
    for (uint64_t i = lo; i < hi; ++i)
      trace_ray<T>(s, rays[i], id0 + i, store ? &parts[c] : nullptr, cnts[c], sc.tm.enabled ? &sc.tm : nullptr);
  }
  for (int64_t c = 0; c < nchunks; ++c) {
    res.cnt.object_tests += cnts[c].object_tests;
    res.cnt.ray_steps += cnts[c].ray_steps;
    res.cnt.segments += cnts[c].segments;
    if (store) res.segs.insert(res.segs.end(), parts[c].begin(), parts[c].end());
  }
  res.primary_rays += n;
}

extern "C" {

int32_t lgo_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

void *lgo_scene_create(const LgObject *objs, uint32_t n_obj, const LgGeoNode *nodes, uint32_t n_nodes,
                       const LgTraceParams *prm) {
  auto s = std::make_unique<OracleScene>();
  s->base = make_scene(objs, n_obj, nodes, n_nodes, prm);
  if (!s->base.ok) return nullptr;
  s->d = cast_scene<double>(s->base);
  s->f = cast_scene<float>(s->base);
  return s.release();
}
void lgo_scene_destroy(void *h) { delete (OracleScene *)h; }

// Tracer::enable_tile_map (tracer.rs:137-146) with TileMap::new(width, height, tiles_x, tiles_y, slabs)
// (tracer.rs:27: 100, 100, 8).  Returns the number of candidate entries (0 when disabled).
uint64_t lgo_tile_map_enable(void *h, int32_t enable, int32_t tiles_x, int32_t tiles_y, int32_t slabs) {
  OracleScene *s = (OracleScene *)h;
  if (!enable) {
    s->tm.enabled = false;
    return 0;
  }
  if (tiles_x < 1 || tiles_y < 1 || slabs < 1) return 0;
  if (s->tm.nx != tiles_x || s->tm.ny != tiles_y || s->tm.ns != slabs || s->tm.cand.empty())
    s->tm = build_tile_map(s->base, tiles_x, tiles_y, slabs);
  s->tm.enabled = true;
  return s->tm.cand.size();
}

// number of lowered tokens and a copy of them (for lowering parity tests):
// each token is written as 12 doubles: kind, op, a_start, b_start, p[0..7]
uint32_t lgo_scene_tokens(void *h, double *dst, uint32_t cap) {
  OracleScene *s = (OracleScene *)h;
  uint32_t n = (uint32_t)s->base.tokens.size();
  for (uint32_t i = 0; i < n && i < cap; ++i) {
    const Token &t = s->base.tokens[i];
    double *o = dst + 12 * i;
    o[0] = t.kind;
    o[1] = t.op;
    o[2] = t.a_start;
    o[3] = t.b_start;
    for (int k = 0; k < 8; ++k) o[4 + k] = t.p[k];
  }
  return n;
}

void lgo_emit_rays(const LgLight *l, uint64_t first, uint64_t count, LgRay *dst) {
  for (uint64_t i = 0; i < count; ++i) emit_ray(*l, first + i, dst[i]);
}

double lgo_start_medium(void *h, const LgLight *l) { return start_medium(((OracleScene *)h)->d, *l); }

int32_t lgo_contains(void *h, int32_t precision, int32_t obj, double x, double y) {
  OracleScene *s = (OracleScene *)h;
  if (precision == LG_PRECISION_F64) return contains_object(s->d, obj, V2<double>{x, y});
  return contains_object(s->f, obj, V2<float>{(float)x, (float)y});
}

// Ray::intersect(&obj.get_geometry()): writes up to cap hits as 5 doubles
// (px, py, nx, ny, t); returns the number of hits.
int32_t lgo_intersect(void *h, int32_t precision, int32_t obj, const double o[2], const double d[2], double *dst,
                      int32_t cap) {
  OracleScene *s = (OracleScene *)h;
  int n = 0;
  if (precision == LG_PRECISION_F64) {
    ObjHits<double> oh;
    intersect_object(s->d, obj, V2<double>{o[0], o[1]}, V2<double>{d[0], d[1]}, oh);
    for (; n < oh.n && n < cap; ++n) {
      double *q = dst + 5 * n;
      q[0] = oh.h[n].p.x, q[1] = oh.h[n].p.y, q[2] = oh.h[n].n.x, q[3] = oh.h[n].n.y, q[4] = oh.h[n].t;
    }
  } else {
    ObjHits<float> oh;
    intersect_object(s->f, obj, V2<float>{(float)o[0], (float)o[1]}, V2<float>{(float)d[0], (float)d[1]}, oh);
    for (; n < oh.n && n < cap; ++n) {
      double *q = dst + 5 * n;
      q[0] = oh.h[n].p.x, q[1] = oh.h[n].p.y, q[2] = oh.h[n].n.x, q[3] = oh.h[n].n.y, q[4] = oh.h[n].t;
    }
  }
  return n;
}

// Ray::refract(&hit,&normal,n1,n2) -> out = {rx, ry, tx, ty, reflectance}; returns has_refracted
int32_t lgo_refract(int32_t precision, const double d[2], const double n[2], double n1, double n2, double out[5]) {
  bool has;
  if (precision == LG_PRECISION_F64) {
    V2<double> r, t{0, 0};
    double R = refract_dir(V2<double>{d[0], d[1]}, V2<double>{n[0], n[1]}, n1, n2, r, t, has);
    out[0] = r.x, out[1] = r.y, out[2] = t.x, out[3] = t.y, out[4] = R;
  } else {
    V2<float> r, t{0, 0};
    float R = refract_dir(V2<float>{(float)d[0], (float)d[1]}, V2<float>{(float)n[0], (float)n[1]}, (float)n1,
                          (float)n2, r, t, has);
    out[0] = r.x, out[1] = r.y, out[2] = t.x, out[3] = t.y, out[4] = R;
  }
  return has ? 1 : 0;
}
void lgo_reflect(const double d[2], const double n[2], double out[2]) {
  V2<double> r = reflect_dir(V2<double>{d[0], d[1]}, V2<double>{n[0], n[1]});
  out[0] = r.x, out[1] = r.y;
}

// Tracer::trace over caller supplied rays
void *lgo_trace_rays(void *h, int32_t precision, const LgRay *rays, uint64_t n, int32_t chunk, int32_t threads,
                     int32_t store) {
  OracleScene *s = (OracleScene *)h;
  auto r = std::make_unique<OracleResult>();
  auto t0 = std::chrono::steady_clock::now();
  if (precision == LG_PRECISION_F64)
    trace_block<double>(*s, rays, n, 0, chunk, threads, store != 0, *r);
  else
    trace_block<float>(*s, rays, n, 0, chunk, threads, store != 0, *r);
  r->seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  return r.release();
}

// Tracer::trace_all (tracer.rs:276-330) for shard `rank` of `world`.
// `stride` > 1 traces only every stride-th ray of the shard (bounded CPU
// baseline samples); ray ids stay those of the full ray set.
void *lgo_trace_all(void *h, int32_t precision, const LgLight *lights, uint32_t n_lights, uint32_t rank,
                    uint32_t world, uint64_t stride, int32_t chunk, int32_t threads, int32_t store) {
  OracleScene *s = (OracleScene *)h;
  auto r = std::make_unique<OracleResult>();
  if (stride < 1) stride = 1;
  auto t0 = std::chrono::steady_clock::now();
  uint64_t id_base = 0;
  for (uint32_t li = 0; li < n_lights; ++li) {
    const LgLight &l = lights[li];
    double n0 = start_medium(s->d, l);
    // local ray k of the shard is ray rank + k*world of the light; `stride` samples every stride-th local ray
    const uint64_t local = shard_count(l.num_rays, rank, world);
    const uint64_t cnt = (local + stride - 1) / stride;
    const uint64_t step = (uint64_t)world * stride;
    // materialise in blocks to bound memory
    const uint64_t BLK = 1u << 20;
    std::vector<LgRay> rays;
    for (uint64_t b = 0; b < cnt; b += BLK) {
      uint64_t m = std::min<uint64_t>(BLK, cnt - b);
      rays.resize(m);
#pragma omp parallel for schedule(static)
      for (int64_t i = 0; i < (int64_t)m; ++i) {
        emit_ray(l, rank + (b + i) * step, rays[i]);
        rays[i].refractive_index = n0;
      }
      // ids are not contiguous: trace the block with local ids and map them back
      size_t before = r->segs.size();
      if (precision == LG_PRECISION_F64)
        trace_block<double>(*s, rays.data(), m, 0, chunk, threads, store != 0, *r);
      else
        trace_block<float>(*s, rays.data(), m, 0, chunk, threads, store != 0, *r);
      for (size_t k = before; k < r->segs.size(); ++k)
        r->segs[k].tag.ray = id_base + rank + (b + r->segs[k].tag.ray) * step;
    }
    id_base += l.num_rays;
  }
  r->seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  return r.release();
}

// TileMap::get_tile (tile_map.rs:134-146), Tile::get_index (229-235) and the candidate list of (tile, slab)
int32_t lgo_tile_map_tile(void *h, double x, double y) { return ((OracleScene *)h)->tm.tile_of(x, y); }
int32_t lgo_tile_map_slab(void *h, double dx, double dy) { return ((OracleScene *)h)->tm.slab_of(dx, dy); }
uint32_t lgo_tile_map_candidates(void *h, int32_t tile, int32_t slab, int32_t *dst, uint32_t cap) {
  const TileMap &tm = ((OracleScene *)h)->tm;
  if (tile < 0 || slab < 0 || slab >= tm.ns || (size_t)tile * tm.ns + slab + 1 >= tm.start.size()) return 0;
  const size_t q = (size_t)tile * tm.ns + slab;
  const uint32_t n = (uint32_t)(tm.start[q + 1] - tm.start[q]);
  for (uint32_t i = 0; i < n && i < cap; ++i) dst[i] = tm.cand[tm.start[q] + i];
  return n;
}

uint64_t lgo_result_count(void *r) { return ((OracleResult *)r)->cnt.segments; }
uint64_t lgo_result_stored(void *r) { return ((OracleResult *)r)->segs.size(); }
uint64_t lgo_result_ray_steps(void *r) { return ((OracleResult *)r)->cnt.ray_steps; }
uint64_t lgo_result_object_tests(void *r) { return ((OracleResult *)r)->cnt.object_tests; }
uint64_t lgo_result_primary_rays(void *r) { return ((OracleResult *)r)->primary_rays; }
double lgo_result_seconds(void *r) { return ((OracleResult *)r)->seconds; }
void lgo_result_copy(void *r, LgSegment *seg, LgSegmentTag *tag, LgSegmentF64 *f64) {
  OracleResult *res = (OracleResult *)r;
  for (size_t i = 0; i < res->segs.size(); ++i) {
    if (seg) seg[i] = res->segs[i].seg;
    if (tag) tag[i] = res->segs[i].tag;
    if (f64) f64[i] = res->segs[i].f64;
  }
}
void lgo_result_free(void *r) { delete (OracleResult *)r; }

// StringMod::draw (string_mod.rs:152-158), chords [first, first+count)
void lgo_string_mod(const LgStringMod *sm, const LgModRemColor *rules, uint32_t n_rules, uint64_t first,
                    uint64_t count, LgVertexPair *dst) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < (int64_t)count; ++i) sm_chord(*sm, rules, n_rules, first + i, dst[i]);
}

// StringMod::line_crossings_as_points (string_mod.rs:87-101) over given chords: returns the number of crossing
// points and copies up to `cap` of them (x, y pairs)
uint64_t lgo_line_crossings(const LgVertexPair *lines, uint64_t n, double *xy, uint64_t cap) {
  std::vector<double> pts = line_crossings(lines, n);
  const uint64_t np = pts.size() / 2;
  if (xy) std::memcpy(xy, pts.data(), std::min(np, cap) * 16);
  return np;
}
// inner.draw_init_points(points) (string_mod.rs:103-122): inner.modulo chords between the given points
void lgo_nested_chords(const LgStringMod *inner, const LgModRemColor *rules, uint32_t n_rules, const double *xy,
                       uint64_t n_points, LgVertexPair *dst) {
  if (n_points == 0) return;
  for (uint64_t i = 0; i < inner->modulo; ++i) nested_chord(*inner, rules, n_rules, xy, n_points, i, dst[i]);
}

} // extern "C"

// SubRenderPass::update_vertex_buffer + render into an fp32 RGBA image
// (row-major, y down).  Returns the number of blended fragments.
template <class Get> static uint64_t accumulate_t(float *img, int W, int H, uint64_t n, int threads, Get get) {
  Proj pr = make_proj(W, H);
  int nt = 1;
#ifdef _OPENMP
  nt = threads > 0 ? threads : omp_get_max_threads();
#endif
  if (nt > H) nt = H;
  uint64_t total = 0;
#pragma omp parallel for num_threads(nt) schedule(static, 1) reduction(+ : total)
  for (int b = 0; b < nt; ++b) {
    int row0 = (int)((int64_t)H * b / nt), row1 = (int)((int64_t)H * (b + 1) / nt);
    for (uint64_t i = 0; i < n; ++i) {
      float a[2], bb[2], ca[4], cb[4];
      get(i, a, bb, ca, cb);
      total += raster_segment(pr, a, bb, ca, cb, row0, row1, [&](int px, int py, const float c[4]) {
        float *p = img + ((size_t)py * W + px) * 4;
        p[0] += c[0]; // rgb: src*1 + dst*1            (mod.rs:66-70)
        p[1] += c[1];
        p[2] += c[2];
        p[3] += c[3] * c[3]; // a: src.a*src.a + dst.a (mod.rs:61-65)
      });
    }
  }
  return total;
}

extern "C" {

void lgo_image_clear(float *img, int32_t W, int32_t H, float clear_alpha) {
  for (size_t i = 0; i < (size_t)W * H; ++i) {
    img[4 * i] = img[4 * i + 1] = img[4 * i + 2] = 0.f; // LoadOp::Clear(BLACK), renderer.rs:174-177
    img[4 * i + 3] = clear_alpha;
  }
}

uint64_t lgo_accumulate_segments(float *img, int32_t W, int32_t H, const LgSegment *seg, uint64_t n,
                                 int32_t threads) {
  return accumulate_t(img, W, H, n, threads, [&](uint64_t i, float a[2], float b[2], float ca[4], float cb[4]) {
    const LgSegment &s = seg[i];
    a[0] = s.a[0], a[1] = s.a[1], b[0] = s.b[0], b[1] = s.b[1];
    std::memcpy(ca, s.color, 16);
    std::memcpy(cb, s.color, 16);
  });
}

// the same fragments summed in f64: the exact value both device resolves are bounded against
uint64_t lgo_accumulate_segments_f64(double *img, int32_t W, int32_t H, const LgSegment *seg, uint64_t n) {
  Proj pr = make_proj(W, H);
  uint64_t total = 0;
  for (uint64_t i = 0; i < n; ++i) {
    const LgSegment &s = seg[i];
    total += raster_segment(pr, s.a, s.b, s.color, s.color, 0, H, [&](int px, int py, const float c[4]) {
      double *p = img + ((size_t)py * W + px) * 4;
      p[0] += c[0], p[1] += c[1], p[2] += c[2];
      p[3] += (double)(c[3] * c[3]);
    });
  }
  return total;
}

// §8.6 the line pass under a user-selected blend state (gui/settings.rs:59-127), fragment by fragment in list order
static float blend_src_factor(int f, float c, float alpha, float constant) {
  switch (f) {
  case LG_BF_ZERO: return 0.f;
  case LG_BF_ONE: return 1.f;
  case LG_BF_SRC: return c;
  case LG_BF_ONE_MINUS_SRC: return 1.f - c;
  case LG_BF_SRC_ALPHA: return alpha;
  case LG_BF_ONE_MINUS_SRC_ALPHA: return 1.f - alpha;
  case LG_BF_CONSTANT: return constant;
  default: return 1.f - constant;
  }
}
static void blend_one(float &dst, const LgBlendComponent &k, float src, float alpha, float constant) {
  switch (k.operation) {
  case LG_BO_MIN: dst = src < dst ? src : dst; break; // wgpu: factors ignored
  case LG_BO_MAX: dst = src > dst ? src : dst; break;
  case LG_BO_ADD: dst = dst + src * blend_src_factor(k.src_factor, src, alpha, constant); break;
  default: dst = dst - src * blend_src_factor(k.src_factor, src, alpha, constant); break; // ReverseSubtract, dst One
  }
}
uint64_t lgo_accumulate_pairs_blend(float *img, int32_t W, int32_t H, const LgVertexPair *vp, uint64_t n,
                                    const LgBlendState *st) {
  Proj pr = make_proj(W, H);
  uint64_t total = 0;
  for (uint64_t i = 0; i < n; ++i) {
    const LgVertexPair &s = vp[i];
    float a[2] = {(float)s.a[0], (float)s.a[1]}, b[2] = {(float)s.b[0], (float)s.b[1]};
    total += raster_segment(pr, a, b, s.color_a, s.color_b, 0, H, [&](int px, int py, const float c[4]) {
      float *p = img + ((size_t)py * W + px) * 4;
      for (int k = 0; k < 3; ++k) blend_one(p[k], st->color, c[k], c[3], st->constant[k]);
      blend_one(p[3], st->alpha, c[3], c[3], st->constant[3]);
    });
  }
  return total;
}

uint64_t lgo_accumulate_pairs(float *img, int32_t W, int32_t H, const LgVertexPair *vp, uint64_t n,
                              int32_t threads) {
  return accumulate_t(img, W, H, n, threads, [&](uint64_t i, float a[2], float b[2], float ca[4], float cb[4]) {
    const LgVertexPair &s = vp[i];
    a[0] = (float)s.a[0], a[1] = (float)s.a[1], b[0] = (float)s.b[0], b[1] = (float)s.b[1]; // `as f32`
    std::memcpy(ca, s.color_a, 16);
    std::memcpy(cb, s.color_b, 16);
  });
}

// fp32 -> IEEE binary16, round to nearest even (what an Rgba16Float store does)
static uint16_t f2h(float f) {
  uint32_t x;
  std::memcpy(&x, &f, 4);
  uint32_t sign = (x >> 16) & 0x8000u;
  uint32_t ex = (x >> 23) & 0xff;
  uint32_t man = x & 0x7fffffu;
  if (ex == 0xff) return (uint16_t)(sign | 0x7c00u | (man ? 0x200u : 0));
  int e = (int)ex - 127 + 15;
  if (e >= 31) return (uint16_t)(sign | 0x7c00u);
  if (e <= 0) {
    if (e < -10) return (uint16_t)sign;
    man |= 0x800000u;
    int shift = 14 - e;
    uint32_t half = man >> shift;
    uint32_t rem = man & ((1u << shift) - 1), mid = 1u << (shift - 1);
    if (rem > mid || (rem == mid && (half & 1))) half++;
    return (uint16_t)(sign | half);
  }
  uint32_t half = ((uint32_t)e << 10) | (man >> 13);
  uint32_t rem = man & 0x1fffu;
  if (rem > 0x1000u || (rem == 0x1000u && (half & 1))) half++;
  return (uint16_t)(sign | half);
}
static float h2f(uint16_t h) {
  uint32_t sign = (uint32_t)(h & 0x8000u) << 16, ex = (h >> 10) & 0x1f, man = h & 0x3ffu, x;
  if (ex == 0) {
    if (man == 0) {
      x = sign;
    } else {
      int e = -1;
      do {
        man <<= 1;
        ++e;
      } while (!(man & 0x400u));
      x = sign | ((uint32_t)(127 - 15 - e) << 23) | ((man & 0x3ffu) << 13);
    }
  } else if (ex == 31) {
    x = sign | 0x7f800000u | (man << 13);
  } else {
    x = sign | ((ex + 112) << 23) | (man << 13);
  }
  float f;
  std::memcpy(&f, &x, 4);
  return f;
}
// Renderer::f16_to_u8 + convert_pixel_rgbaf16_to_bgra8 (src/renderer.rs:313-328)
void lgo_image_to_bgra8(const float *img, uint64_t n_px, uint8_t *dst) {
  for (uint64_t i = 0; i < n_px; ++i) {
    uint8_t c[4];
    for (int k = 0; k < 4; ++k) {
      float f = h2f(f2h(img[4 * i + k]));
      float g = std::pow(f, 1.f / 2.2f) * 255.f; // f32 powf, like Rust's f32::powf
      c[k] = !(g == g) ? 0 : g <= 0.f ? 0 : g >= 255.f ? 255 : (uint8_t)g; // `as u8`: saturating, NaN -> 0
    }
    dst[4 * i] = c[2], dst[4 * i + 1] = c[1], dst[4 * i + 2] = c[0], dst[4 * i + 3] = c[3];
  }
}
// ORACLE.md 8.7: the 8-bit surface target (sub_render_pass.rs:59-63, Bgra8UnormSrgb at renderer.rs:207-209).
// colour byte = number of k in 1..255 whose threshold T[k] = f32(linear value encoding to (k - 0.5) / 255) is <= c;
// alpha byte = floor(min(a, 1) * 255 + 0.5) in f32; anything not > 0 (NaN included) is 0.  Stored [b, g, r, a].
void lgo_image_to_bgra8_srgb(const float *img, uint64_t n_px, uint8_t *dst) {
  float thr[256];
  lgo::surface_thresholds(thr);
  for (uint64_t i = 0; i < n_px; ++i) {
    const float *p = img + 4 * i;
    dst[4 * i] = lgo::surface_colour_byte(thr, p[2]), dst[4 * i + 1] = lgo::surface_colour_byte(thr, p[1]);
    dst[4 * i + 2] = lgo::surface_colour_byte(thr, p[0]), dst[4 * i + 3] = lgo::surface_alpha_byte(p[3]);
  }
}
void lgo_image_to_f16(const float *img, uint64_t n_floats, uint16_t *dst) {
  for (uint64_t i = 0; i < n_floats; ++i) dst[i] = f2h(img[i]);
}

} // extern "C"
