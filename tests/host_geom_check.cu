// tests/host_geom_check.cu — runs the product's geometry header
// (light_garden_b200/csrc/lg_geom.cuh, host instantiation of its
// __host__ __device__ functions) against the oracle (oracle/lg_oracle.hpp) on
// random inputs and demands bit-identical results.  CPU only; built and run by
// tests/test_geom_host.py.  Prints "OK <n>" or "MISMATCH ..." lines.
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <cstdint>
#include <cmath>
#include <vector>
#include <string>

#include "../light_garden_b200/csrc/lg_geom.cuh"
#include "../light_garden_b200/csrc/lg_scene.h"
#include "../light_garden_b200/csrc/lg_nearest.cuh"
#include "../light_garden_b200/csrc/lg_tables.h"
#include "../light_garden_b200/csrc/lg_srgb.h"
#include "../oracle/lg_oracle.hpp"

static uint64_t s_state = 0x4C47BEEFull;
static uint64_t nextu() {
  s_state += 0x9E3779B97F4A7C15ull;
  uint64_t z = s_state;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
static double uni(double lo, double hi) { return lo + (hi - lo) * ((nextu() >> 11) * (1.0 / 9007199254740992.0)); }

template <class T> static bool same(T a, T b) { return std::memcmp(&a, &b, sizeof(T)) == 0 || (a != a && b != b); }

static long g_bad = 0;
#define CHECK(cond, what)                                  \
  do {                                                     \
    if (!(cond)) {                                         \
      if (g_bad < 20) std::printf("MISMATCH %s (%s)\n", what, sizeof(T) == 4 ? "f32" : "f64"); \
      ++g_bad;                                             \
    }                                                      \
  } while (0)

template <class T> static void cmp_lists(const lg::CandList<T> &a, const lgo::HitList<T> &b, const char *what) {
  CHECK(a.n == b.n, what);
  for (int i = 0; i < a.n && i < b.n; ++i) {
    CHECK(same(a.h[i].t, b.h[i].t), what);
    CHECK(same(a.h[i].p.x, b.h[i].p.x) && same(a.h[i].p.y, b.h[i].p.y), what);
  }
}

template <class T> static long run(int iters) {
  long n = 0;
  for (int it = 0; it < iters; ++it) {
    // ray
    double ang = uni(0, 6.283185307179586);
    lg::V2<T> o{(T)uni(-1.5, 1.5), (T)uni(-1, 1)}, d{(T)std::cos(ang), (T)std::sin(ang)};
    lgo::V2<T> oo{o.x, o.y}, od{d.x, d.y};
    // circle
    {
      lg::Tok<T> k{};
      k.kind = lg::TOK_CIRCLE;
      k.p[0] = (T)uni(-1.5, 1.5), k.p[1] = (T)uni(-1, 1), k.p[2] = (T)uni(0.005, 0.6);
      k.p[3] = k.p[2] * k.p[2];
      lg::CandList<T> a;
      a.n = 0;
      lg::hit_circle(k.p, o, d, a);
      lgo::HitList<T> b;
      lgo::hit_circle(k.p, oo, od, b);
      cmp_lists(a, b, "circle");
      for (int i = 0; i < a.n && i < b.n; ++i) {
        lg::V2<T> nn = lg::hit_normal(k, a.h[i].p, a.h[i].aux);
        CHECK(same(nn.x, b.h[i].n.x) && same(nn.y, b.h[i].n.y), "circle normal");
      }
      lgo::LeafT<T> l{};
      l.kind = lgo::TOK_CIRCLE;
      std::memcpy(l.p, k.p, sizeof k.p);
      lg::V2<T> q{(T)uni(-1.5, 1.5), (T)uni(-1, 1)};
      CHECK(lg::contains_leaf(k, q) == lgo::contains_leaf(l, lgo::V2<T>{q.x, q.y}), "circle contains");
      n += 1 + a.n;
    }
    // segment
    {
      lg::Tok<T> k{};
      k.kind = lg::TOK_SEGMENT;
      T ax = (T)uni(-1.5, 1.5), ay = (T)uni(-1, 1), bx = (T)uni(-1.5, 1.5), by = (T)uni(-1, 1);
      k.p[0] = ax, k.p[1] = ay, k.p[2] = bx - ax, k.p[3] = by - ay;
      lg::CandList<T> a;
      a.n = 0;
      lg::hit_segment(k.p, o, d, a);
      lgo::HitList<T> b;
      lgo::hit_segment(k.p, oo, od, b);
      cmp_lists(a, b, "segment");
      for (int i = 0; i < a.n && i < b.n; ++i) {
        lg::V2<T> nn = lg::hit_normal(k, a.h[i].p, a.h[i].aux);
        CHECK(same(nn.x, b.h[i].n.x) && same(nn.y, b.h[i].n.y), "segment normal");
      }
      n += 1 + a.n;
    }
    // rect
    {
      lg::Tok<T> k{};
      k.kind = lg::TOK_RECT;
      double ra = uni(0, 6.283185307179586), hw = uni(0.01, 0.5), hh = uni(0.01, 0.5);
      k.p[0] = (T)uni(-1.5, 1.5), k.p[1] = (T)uni(-1, 1);
      k.p[2] = (T)(std::cos(ra) * hw), k.p[3] = (T)(std::sin(ra) * hw);
      k.p[4] = (T)(-std::sin(ra) * hh), k.p[5] = (T)(std::cos(ra) * hh);
      k.p[6] = lg::dot(lg::V2<T>{k.p[2], k.p[3]}, lg::V2<T>{k.p[2], k.p[3]});
      k.p[7] = lg::dot(lg::V2<T>{k.p[4], k.p[5]}, lg::V2<T>{k.p[4], k.p[5]});
      lg::CandList<T> a;
      a.n = 0;
      lg::hit_rect(k.p, o, d, a);
      lgo::HitList<T> b;
      lgo::hit_rect(k.p, oo, od, b);
      cmp_lists(a, b, "rect");
      for (int i = 0; i < a.n && i < b.n; ++i) {
        lg::V2<T> nn = lg::hit_normal(k, a.h[i].p, a.h[i].aux);
        CHECK(same(nn.x, b.h[i].n.x) && same(nn.y, b.h[i].n.y), "rect normal");
      }
      lgo::LeafT<T> l{};
      l.kind = lgo::TOK_RECT;
      std::memcpy(l.p, k.p, sizeof k.p);
      lg::V2<T> q{(T)(k.p[0] + uni(-0.6, 0.6)), (T)(k.p[1] + uni(-0.6, 0.6))};
      CHECK(lg::contains_leaf(k, q) == lgo::contains_leaf(l, lgo::V2<T>{q.x, q.y}), "rect contains");
      n += 1 + a.n;
    }
    // ellipse
    {
      lg::Tok<T> k{};
      k.kind = lg::TOK_ELLIPSE;
      double ra = uni(0, 6.283185307179586);
      k.p[0] = (T)uni(-1.5, 1.5), k.p[1] = (T)uni(-1, 1);
      k.p[2] = (T)std::cos(ra), k.p[3] = (T)std::sin(ra);
      k.p[4] = (T)uni(0.01, 0.8), k.p[5] = (T)uni(0.01, 0.8);
      k.p[6] = (T)1 / k.p[4], k.p[7] = (T)1 / k.p[5];
      lg::CandList<T> a;
      a.n = 0;
      lg::hit_ellipse(k.p, o, d, a);
      lgo::HitList<T> b;
      lgo::hit_ellipse(k.p, oo, od, b);
      cmp_lists(a, b, "ellipse");
      for (int i = 0; i < a.n && i < b.n; ++i) {
        lg::V2<T> nn = lg::hit_normal(k, a.h[i].p, a.h[i].aux);
        CHECK(same(nn.x, b.h[i].n.x) && same(nn.y, b.h[i].n.y), "ellipse normal");
      }
      lgo::LeafT<T> l{};
      l.kind = lgo::TOK_ELLIPSE;
      std::memcpy(l.p, k.p, sizeof k.p);
      lg::V2<T> q{(T)(k.p[0] + uni(-0.8, 0.8)), (T)(k.p[1] + uni(-0.8, 0.8))};
      CHECK(lg::contains_leaf(k, q) == lgo::contains_leaf(l, lgo::V2<T>{q.x, q.y}), "ellipse contains");
      n += 1 + a.n;
    }
    // bezier
    {
      lg::Tok<T> k{};
      k.kind = lg::TOK_BEZIER;
      double cx = uni(-1, 1), cy = uni(-0.7, 0.7);
      for (int i = 0; i < 4; ++i) k.p[2 * i] = (T)(cx + uni(-0.6, 0.6)), k.p[2 * i + 1] = (T)(cy + uni(-0.6, 0.6));
      lg::CandList<T> a;
      a.n = 0;
      lg::hit_bezier(k.p, o, d, a);
      lgo::HitList<T> b;
      lgo::hit_bezier(k.p, oo, od, b);
      cmp_lists(a, b, "bezier");
      for (int i = 0; i < a.n && i < b.n; ++i) {
        lg::V2<T> nn = lg::hit_normal(k, a.h[i].p, a.h[i].aux);
        CHECK(same(nn.x, b.h[i].n.x) && same(nn.y, b.h[i].n.y), "bezier normal");
      }
      n += 1 + a.n;
    }
    // reflect / refract
    {
      double na = uni(0, 6.283185307179586);
      lg::V2<T> nn{(T)std::cos(na), (T)std::sin(na)};
      lg::V2<T> r = lg::reflect_dir(d, nn);
      lgo::V2<T> r2 = lgo::reflect_dir(od, lgo::V2<T>{nn.x, nn.y});
      CHECK(same(r.x, r2.x) && same(r.y, r2.y), "reflect");
      T n1 = (T)uni(1.0, 2.4), n2 = (T)uni(1.0, 2.4);
      lg::V2<T> fl, fr{0, 0};
      bool has;
      T R = lg::refract_dir(d, nn, n1, n2, fl, fr, has);
      lgo::V2<T> gl, gr{0, 0};
      bool has2;
      T R2 = lgo::refract_dir(od, lgo::V2<T>{nn.x, nn.y}, n1, n2, gl, gr, has2);
      CHECK(has == has2 && same(R, R2), "refract reflectance");
      CHECK(same(fl.x, gl.x) && same(fl.y, gl.y), "refract reflected");
      if (has && has2) CHECK(same(fr.x, gr.x) && same(fr.y, gr.y), "refract refracted");
      n += 2;
    }
  }
  return n;
}

// random Geo trees: product lowering (lg_scene.h) vs oracle lowering, bit for bit
static int rand_geo(std::vector<LgGeoNode> &nodes, int depth) {
  LgGeoNode g{};
  g.child_a = g.child_b = -1;
  g.rot[0] = 1, g.rot[3] = 1;
  int kind = depth >= 3 ? (int)(nextu() % 5) : (int)(nextu() % 6);
  if (depth >= 3 && kind == 4) kind = 5; // no deeper Logic nodes
  if (nextu() % 6 == 0) kind = LG_GEO_POLYGON;
  g.kind = kind;
  double ra = uni(0, 6.283185307179586);
  switch (kind) {
  case LG_GEO_POLYGON: { // vertices on an ellipse at sorted random angles: convex, either winding
    const int k = 3 + (int)(nextu() % 10);
    std::vector<double> ang(k);
    for (double &a : ang) a = uni(0, 6.283185307179586);
    std::sort(ang.begin(), ang.end());
    if (nextu() % 2) std::reverse(ang.begin(), ang.end());
    const double ax = uni(0.05, 0.5), ay = uni(0.05, 0.5);
    g.op = k;
    g.p[0] = uni(-1, 1), g.p[1] = uni(-1, 1);
    g.rot[0] = std::cos(ra), g.rot[1] = std::sin(ra), g.rot[2] = -std::sin(ra), g.rot[3] = std::cos(ra);
    const int ix = (int)nodes.size();
    nodes.push_back(g);
    int prev = ix;
    for (int v = 0; v < k; v += 4) {
      LgGeoNode c{};
      c.kind = LG_GEO_POINTS, c.child_a = c.child_b = -1;
      c.rot[0] = 1, c.rot[3] = 1;
      c.op = std::min(4, k - v);
      for (int q = 0; q < c.op; ++q) c.p[2 * q] = ax * std::cos(ang[v + q]), c.p[2 * q + 1] = ay * std::sin(ang[v + q]);
      nodes[prev].child_a = (int)nodes.size();
      prev = (int)nodes.size();
      nodes.push_back(c);
    }
    return ix;
  }
  case LG_GEO_CIRCLE: g.p[0] = uni(-1, 1), g.p[1] = uni(-1, 1), g.p[2] = uni(0.05, 0.5); break;
  case LG_GEO_RECT:
    g.p[0] = uni(-1, 1), g.p[1] = uni(-1, 1), g.p[2] = uni(0.05, 0.8), g.p[3] = uni(0.05, 0.8);
    g.rot[0] = std::cos(ra), g.rot[1] = std::sin(ra), g.rot[2] = -std::sin(ra), g.rot[3] = std::cos(ra);
    break;
  case LG_GEO_ELLIPSE:
    g.p[0] = uni(-1, 1), g.p[1] = uni(-1, 1), g.p[2] = uni(0.05, 0.6), g.p[3] = uni(0.05, 0.6);
    g.rot[0] = std::cos(ra), g.rot[1] = std::sin(ra), g.rot[2] = -std::sin(ra), g.rot[3] = std::cos(ra);
    break;
  case LG_GEO_SEGMENT: for (int k = 0; k < 4; ++k) g.p[k] = uni(-1, 1); break;
  case LG_GEO_BEZIER: for (int k = 0; k < 8; ++k) g.p[k] = uni(-1, 1); break;
  default: {
    g.op = (int)(nextu() % 3);
    g.p[0] = uni(-0.5, 0.5), g.p[1] = uni(-0.5, 0.5);
    g.rot[0] = std::cos(ra), g.rot[1] = std::sin(ra), g.rot[2] = -std::sin(ra), g.rot[3] = std::cos(ra);
    int ix = (int)nodes.size();
    nodes.push_back(g);
    int a = rand_geo(nodes, depth + 1);
    int b = rand_geo(nodes, depth + 1);
    nodes[ix].child_a = a, nodes[ix].child_b = b;
    return ix;
  }
  }
  nodes.push_back(g);
  return (int)nodes.size() - 1;
}

static long check_lowering(int n_scenes) {
  long n = 0;
  for (int sidx = 0; sidx < n_scenes; ++sidx) {
    std::vector<LgGeoNode> nodes;
    std::vector<LgObject> objs;
    int nobj = 1 + (int)(nextu() % 6);
    for (int i = 0; i < nobj; ++i) {
      LgObject o{};
      o.root = rand_geo(nodes, 0);
      o.has_material = (int)(nextu() % 2);
      o.refractive_index = uni(1.05, 2.4);
      objs.push_back(o);
    }
    LgTraceParams prm{};
    prm.max_bounce = 5;
    prm.canvas_tlbr[0] = 1, prm.canvas_tlbr[1] = -1.7, prm.canvas_tlbr[2] = -1, prm.canvas_tlbr[3] = 1.7;
    lg::HostScene hs;
    std::string err;
    int rc = lg::lower_scene(objs.data(), (uint32_t)objs.size(), nodes.data(), (uint32_t)nodes.size(), prm, hs, err);
    lgo::Scene os = lgo::make_scene(objs.data(), (uint32_t)objs.size(), nodes.data(), (uint32_t)nodes.size(), &prm);
    if ((rc == 0) != os.ok) { std::printf("MISMATCH lowering status\n"); ++g_bad; continue; }
    if (rc) continue;
    if (hs.toks.size() != os.tokens.size() || hs.objs.size() != os.objects.size()) { std::printf("MISMATCH lowering size\n"); ++g_bad; continue; }
    for (size_t i = 0; i < hs.toks.size(); ++i) {
      const lg::HostTok &a = hs.toks[i];
      const lgo::Token &b = os.tokens[i];
      bool ok = a.kind == b.kind && (a.kind != 4 || (a.op == b.op && a.a_start == b.a_start && a.b_start == b.b_start));
      if (a.kind == 6 || a.kind == 7) ok = ok && a.op == b.op;
      int np = a.kind == 0 ? 3 : a.kind == 1 ? 6 : a.kind == 2 ? 4 : a.kind == 3 ? 8 : a.kind == 5 ? 6 : a.kind == 7 ? 2 * a.op : 0;
      for (int k = 0; k < np; ++k) ok = ok && std::memcmp(&a.p[k], &b.p[k], 8) == 0;
      if (!ok) { if (g_bad < 20) std::printf("MISMATCH lowering token %zu kind %d\n", i, a.kind); ++g_bad; }
      ++n;
    }
    for (size_t i = 0; i < hs.objs.size(); ++i)
      if (hs.objs[i].first != os.objects[i].first || hs.objs[i].count != os.objects[i].count) { std::printf("MISMATCH object range\n"); ++g_bad; }
    // start medium + overlap-candidate completeness: any other material object containing a
    // point near object i's boundary must be in i's candidate list
    lgo::SceneT<double> sd = lgo::cast_scene<double>(os);
    for (int q = 0; q < 50; ++q) {
      double x = uni(-1.7, 1.7), y = uni(-1, 1);
      LgLight l{};
      l.position[0] = x, l.position[1] = y;
      double m1 = lg::host_start_medium(hs, x, y), m2 = lgo::start_medium(sd, l);
      if (std::memcmp(&m1, &m2, 8) != 0) { if (g_bad < 20) std::printf("MISMATCH start medium\n"); ++g_bad; }
      ++n;
    }
    for (size_t i = 0; i < hs.objs.size(); ++i) {
      if (!hs.objs[i].has_material) continue;
      for (int q = 0; q < 200; ++q) {
        double ang = uni(0, 6.283185307179586);
        lgo::V2<double> o{uni(-1.7, 1.7), uni(-1, 1)}, d{std::cos(ang), std::sin(ang)};
        lgo::ObjHits<double> oh;
        lgo::intersect_object(sd, (int)i, o, d, oh);
        for (int h = 0; h < oh.n; ++h) {
          for (size_t j = 0; j < hs.objs.size(); ++j) {
            if (j == i || !hs.objs[j].has_material) continue;
            if (!lgo::contains_object(sd, (int)j, oh.h[h].p)) continue;
            bool listed = false;
            for (int e = hs.ovl_start[i]; e < hs.ovl_start[i + 1]; ++e) listed = listed || hs.ovl_list[e] == (int)j;
            if (!listed) { if (g_bad < 20) std::printf("MISMATCH overlap list misses %zu in %zu\n", j, i); ++g_bad; }
            ++n;
          }
        }
      }
    }
  }
  return n;
}

// The uniform-grid walk (lg_nearest.cuh grid_nearest, tables from lg_tables.h) must return the all-objects loop's
// nearest hit bit for bit: object, token, distance, point.  Scenes: random sizes, and a jittered lattice whose pitch
// equals the cell size so that rays graze cell corners; rays: random, axis-parallel, from outside the box, and aimed
// at cell corners.
template <class T> static long check_grid_scene(const std::vector<LgObject> &objs, const std::vector<LgGeoNode> &nodes, int n_rays,
                                                double density) {
  LgTraceParams prm{};
  prm.max_bounce = 5;
  prm.canvas_tlbr[0] = 1, prm.canvas_tlbr[1] = -1.7, prm.canvas_tlbr[2] = -1, prm.canvas_tlbr[3] = 1.7;
  lg::HostScene hs;
  std::string err;
  if (lg::lower_scene(objs.data(), (uint32_t)objs.size(), nodes.data(), (uint32_t)nodes.size(), prm, hs, err)) return 0;
  const std::vector<lg::Tok<T>> toks = lg::device_tokens<T>(hs);
  std::vector<int> first, count;
  for (const lg::HostObj &o : hs.objs) first.push_back(o.first), count.push_back(o.count);
  const std::vector<double> circ = lg::object_circles(hs);
  const size_t n = hs.objs.size();
  const double B = std::fmax(hs.bound, 4.0);
  const lg::BoundsTable<T> bt = lg::build_bounds<T>(circ.data(), n, B);
  const lg::HostGrid g = lg::build_scene_grid<T>(bt, n, density);
  lg::SceneArgs<T> A{};
  A.toks = toks.data(), A.obj_first = first.data(), A.obj_count = count.data(), A.delta = (T)bt.delta;
  A.grid_x0 = (T)g.x0, A.grid_y0 = (T)g.y0, A.grid_x1 = (T)g.x1, A.grid_y1 = (T)g.y1;
  A.grid_cs = (T)g.cs, A.grid_ics = (T)(1.0 / g.cs), A.grid_eta = (T)bt.delta;
  A.grid_nx = g.nx, A.grid_ny = g.ny, A.grid_start = g.start.data(), A.grid_obj = g.obj.data();
  // the oracle's view of the same scene: Ray::intersect per object, strict `<` on the squared distance in object
  // order (tracer.rs:412-424) -- the product's all-objects loop must agree bit for bit
  const lgo::Scene osc = lgo::make_scene(objs.data(), (uint32_t)objs.size(), nodes.data(), (uint32_t)nodes.size(), &prm);
  const lgo::SceneT<T> ost = lgo::cast_scene<T>(osc);
  long done = 0;
  for (int q = 0; q < n_rays; ++q) {
    lg::V2<T> o{(T)uni(-2.2, 2.2), (T)uni(-1.4, 1.4)};
    double ang = uni(0, 6.283185307179586);
    lg::V2<T> d{(T)std::cos(ang), (T)std::sin(ang)};
    const int mode = q % 8;
    if (mode == 1) d = {(T)1, (T)0};
    if (mode == 2) d = {(T)0, (T)-1};
    if (mode == 3 || mode == 4) { // through a cell corner (exactly, or an ulp-scale distance away)
      const double cx = g.x0 + (double)(nextu() % (unsigned)(g.nx + 1)) * g.cs, cy = g.y0 + (double)(nextu() % (unsigned)(g.ny + 1)) * g.cs;
      const double tt = uni(0.01, 1.5), wob = mode == 4 ? uni(-1e-6, 1e-6) : 0.0;
      o = {(T)(cx - tt * std::cos(ang) + wob), (T)(cy - tt * std::sin(ang))};
    }
    if (mode == 5 && n) { // starts on an object's bounding circle
      const size_t i = nextu() % n;
      o = {(T)(circ[3 * i] + circ[3 * i + 2] * std::cos(ang)), (T)(circ[3 * i + 1] + circ[3 * i + 2] * std::sin(ang))};
    }
    lg::Best<T> b0;
    b0.d2 = lg::Real<T>::max_value(), b0.obj = -1, b0.tok = -1, b0.px = b0.py = b0.aux = (T)0;
    const lg::Best<T> a = lg::all_objects_nearest(A, (int)n, b0, o, d), b = lg::grid_nearest(A, b0, o, d);
    if (a.obj != b.obj || a.tok != b.tok || !same(a.d2, b.d2) || !same(a.px, b.px) || !same(a.py, b.py) || !same(a.aux, b.aux)) {
      if (g_bad < 20)
        std::printf("MISMATCH grid walk (%s): all-objects obj %d d2 %.9g, grid obj %d d2 %.9g, ray (%.9g %.9g)+(%.9g %.9g), grid %dx%d\n",
                    sizeof(T) == 4 ? "f32" : "f64", a.obj, (double)a.d2, b.obj, (double)b.d2, (double)o.x, (double)o.y, (double)d.x,
                    (double)d.y, g.nx, g.ny);
      ++g_bad;
    }
    if (q % 4 == 0 && osc.ok) {
      T best = lg::Real<T>::max_value();
      int bobj = -1;
      lgo::V2<T> bp{0, 0};
      for (int j = 0; j < (int)n; ++j) {
        lgo::ObjHits<T> oh;
        lgo::intersect_object(ost, j, lgo::V2<T>{o.x, o.y}, lgo::V2<T>{d.x, d.y}, oh);
        for (int h = 0; h < oh.n; ++h) {
          const T dx = oh.h[h].p.x - o.x, dy = oh.h[h].p.y - o.y, d2 = dx * dx + dy * dy;
          if (d2 < best) best = d2, bobj = j, bp = oh.h[h].p;
        }
      }
      if (bobj != a.obj || (bobj >= 0 && (!same(best, a.d2) || !same(bp.x, a.px) || !same(bp.y, a.py)))) {
        if (g_bad < 20)
          std::printf("MISMATCH nearest hit vs oracle (%s): oracle obj %d d2 %.9g, product obj %d d2 %.9g\n",
                      sizeof(T) == 4 ? "f32" : "f64", bobj, (double)best, a.obj, (double)a.d2);
        ++g_bad;
      }
      ++done;
    }
    ++done;
  }
  return done;
}

static long check_grid(int n_scenes) {
  long n = 0;
  for (int sidx = 0; sidx < n_scenes; ++sidx) {
    std::vector<LgGeoNode> nodes;
    std::vector<LgObject> objs;
    const bool lattice = sidx % 2 == 0;
    const int side = 4 + (int)(nextu() % 20);
    const int nobj = lattice ? side * side : 1 + (int)(nextu() % 60);
    for (int i = 0; i < nobj; ++i) {
      LgObject o{};
      if (lattice) {
        LgGeoNode g{};
        g.child_a = g.child_b = -1;
        g.rot[0] = 1, g.rot[3] = 1;
        const double pitch = 3.0 / side, x = -1.5 + (i % side + 0.5 + uni(-0.3, 0.3)) * pitch, y = -1.0 + (i / side + 0.5 + uni(-0.3, 0.3)) * (2.0 / side);
        const int kind = (int)(nextu() % 3);
        if (kind == 0) {
          g.kind = LG_GEO_CIRCLE, g.p[0] = x, g.p[1] = y, g.p[2] = uni(0.1, 0.45) * pitch;
        } else if (kind == 1) {
          const double ra = uni(0, 6.283185307179586);
          g.kind = LG_GEO_RECT, g.p[0] = x, g.p[1] = y, g.p[2] = uni(0.1, 0.4) * pitch, g.p[3] = uni(0.1, 0.4) * pitch;
          g.rot[0] = std::cos(ra), g.rot[1] = std::sin(ra), g.rot[2] = -std::sin(ra), g.rot[3] = std::cos(ra);
        } else {
          const double ra = uni(0, 6.283185307179586), l = uni(0.1, 0.6) * pitch;
          g.kind = LG_GEO_SEGMENT, g.p[0] = x - l * std::cos(ra), g.p[1] = y - l * std::sin(ra), g.p[2] = x + l * std::cos(ra), g.p[3] = y + l * std::sin(ra);
        }
        nodes.push_back(g);
        o.root = (int)nodes.size() - 1;
      } else {
        o.root = rand_geo(nodes, 0);
      }
      o.has_material = (int)(nextu() % 2);
      o.refractive_index = uni(1.05, 2.4);
      objs.push_back(o);
    }
    const double density = sidx % 3 == 0 ? 1.0 : (sidx % 3 == 1 ? 4.0 : 0.25);
    n += check_grid_scene<float>(objs, nodes, 2000, density);
    n += check_grid_scene<double>(objs, nodes, 2000, density);
  }
  return n;
}

// ORACLE.md 8.7: the product's 8-bit surface encoding (binary search over its threshold table) against the oracle's
// (its own table, linear count): random values, every threshold and its two neighbours, the special values.
static long check_surface(int iters) {
  const lg::SrgbThresholds T = lg::srgb_thresholds();
  float thr[256];
  lgo::surface_thresholds(thr);
  long n = 0;
  auto one = [&](float v) {
    if (lg::srgb_byte(T.t, v) != lgo::surface_colour_byte(thr, v) || lg::unorm_byte(v) != lgo::surface_alpha_byte(v)) {
      if (g_bad < 20) std::printf("MISMATCH surface byte of %.9g\n", (double)v);
      ++g_bad;
    }
    ++n;
  };
  for (int k = 1; k < 256; ++k) {
    one(thr[k]), one(std::nextafterf(thr[k], 0.f)), one(std::nextafterf(thr[k], 2.f));
    one((float)k / 255.f), one(((float)k - 0.5f) / 255.f), one(std::nextafterf(((float)k - 0.5f) / 255.f, 0.f));
  }
  const float special[] = {0.f, -0.f, 1.f, 2.f, -1.f, INFINITY, -INFINITY, NAN, 1e-30f, 1e-45f, 0.0031308f, 0.04045f, 0.5f};
  for (float v : special) one(v);
  for (int i = 0; i < iters; ++i) {
    one((float)uni(0.0, 1.2)), one((float)uni(0.0, 0.01)), one((float)uni(-0.5, 300.0));
  }
  return n;
}

int main(int argc, char **argv) {
  int iters = argc > 1 ? std::atoi(argv[1]) : 200000;
  long n = run<float>(iters) + run<double>(iters);
  n += check_lowering(iters / 500 + 10);
  n += check_grid(iters / 2500 + 4);
  n += check_surface(iters);
  if (g_bad) {
    std::printf("FAILED %ld mismatches\n", g_bad);
    return 1;
  }
  std::printf("OK %ld\n", n);
  return 0;
}
