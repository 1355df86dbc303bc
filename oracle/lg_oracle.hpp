// oracle/lg_oracle.hpp — CPU restatement of Light Garden's trace + accumulate
// path.  TEST INFRASTRUCTURE ONLY: nothing under light_garden_b200/ may
// include, link or execute this file (see oracle/ORACLE.md).
//
// PARITY UNPINNED.  The reference (sphereflow/light_garden) has no tests and no
// golden vectors, and its ray/shape arithmetic lives in the un-vendored git
// dependency collision2d (github.com/sphereflow/collision2d, pinned
// a7b471b54a940622f658a16177d63925536f16c3, Cargo.lock:880-887) whose source is
// not available offline.  Control flow below follows the reference's own
// files line by line (cited at each function); the geometry kernel is the
// builder-specified stand-in written down in oracle/ORACLE.md.
//
// Everything is templated on the real type T (double = the reference's Float,
// float = the device throughput mode) and uses only IEEE-754 correctly rounded
// operations (+ - * / sqrt fma) in a fixed order, so that the CUDA kernels,
// which are written independently to the same ORACLE.md formulas, can be
// compared bit-for-bit.  Compile with -ffp-contract=off.
#pragma once
#include <cmath>
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <limits>
#include <vector>

#include "../include/light_garden_b200.h"

namespace lgo {

// ORACLE.md §1 constants
constexpr double T_MIN = 1e-5;    // ray parameter acceptance threshold
constexpr double PAR_EPS = 1e-12; // |cross(d,e)| <= PAR_EPS -> parallel
constexpr double EPSILON = 1e-10; // stand-in for collision2d::EPSILON (light.rs:230)
constexpr int MAX_TOKENS = 64;    // postfix program length cap per object

template <class T> struct V2 {
  T x, y;
};
template <class T> inline V2<T> sub(V2<T> a, V2<T> b) { return {a.x - b.x, a.y - b.y}; }
template <class T> inline V2<T> add(V2<T> a, V2<T> b) { return {a.x + b.x, a.y + b.y}; }
// ORACLE.md §1: dot = fma(ax,bx, ay*by); cross = fma(ax,by, -(ay*bx))
template <class T> inline T dot(V2<T> a, V2<T> b) { return std::fma(a.x, b.x, a.y * b.y); }
template <class T> inline T cross(V2<T> a, V2<T> b) { return std::fma(a.x, b.y, -(a.y * b.x)); }
template <class T> inline V2<T> normalize(V2<T> v) {
  T len = std::sqrt(dot(v, v));
  return {v.x / len, v.y / len};
}
// o + t*d
template <class T> inline V2<T> along(V2<T> o, T t, V2<T> d) {
  return {std::fma(t, d.x, o.x), std::fma(t, d.y, o.y)};
}

// ---------------------------------------------------------------------------
// Lowered scene (ORACLE.md §2): every Geo tree becomes a postfix program over
// world-space leaves.  Lowering is done in f64 and then cast to T.
// ---------------------------------------------------------------------------
enum TokKind : int32_t {
  TOK_CIRCLE = 0, TOK_RECT = 1, TOK_SEGMENT = 2, TOK_BEZIER = 3, TOK_OP = 4, TOK_ELLIPSE = 5,
  TOK_POLY = 6,  // convex polygon header: op = number of vertices; its vertices follow in TOK_VERTS tokens
  TOK_VERTS = 7  // up to four world-space vertices (op = how many)
};

struct Token {
  int32_t kind;    // TokKind
  int32_t op;      // LG_OP_* when kind == TOK_OP
  int32_t a_start; // TOK_OP: first token of subtree a (relative to the object)
  int32_t b_start; // TOK_OP: first token of subtree b
  double p[8];     // world-space leaf parameters (f64 master copy)
};

struct Object {
  int32_t first; // first token in Scene::tokens
  int32_t count;
  bool has_material;
  double n;
};

struct Mat2 { // column-major like nalgebra: [m11, m21, m12, m22]
  double m11, m21, m12, m22;
};
inline Mat2 mat_mul(const Mat2 &p, const Mat2 &c) {
  return {p.m11 * c.m11 + p.m12 * c.m21, p.m21 * c.m11 + p.m22 * c.m21,
          p.m11 * c.m12 + p.m12 * c.m22, p.m21 * c.m12 + p.m22 * c.m22};
}
inline void mat_apply(const Mat2 &m, const double t[2], double x, double y, double out[2]) {
  out[0] = m.m11 * x + m.m12 * y + t[0];
  out[1] = m.m21 * x + m.m22 * y + t[1];
}

struct Scene {
  std::vector<Token> tokens;
  std::vector<Object> objects;
  uint32_t max_bounce = 5;
  float cutoff[4] = {0.001f, 0.001f, 0.001f, 0.001f};
  double canvas_tlbr[4] = {1, -1, -1, 1};
  std::vector<LgGeoNode> nodes; // original tree (for the recursive cross-check)
  std::vector<LgObject> raw_objects;
  bool ok = true;
};

// Recursive lowering: world = M * local + t  (Logic children live in the
// node's local frame: src/light_garden/object.rs:393-410, default.ron:3-31).
inline bool lower_node(const std::vector<LgGeoNode> &nodes, int32_t ix, const Mat2 &M,
                       const double t[2], std::vector<Token> &out, int32_t base, int depth) {
  if (ix < 0 || (size_t)ix >= nodes.size() || depth > 32) return false;
  const LgGeoNode &g = nodes[ix];
  Token tok{};
  switch (g.kind) {
  case LG_GEO_CIRCLE: {
    tok.kind = TOK_CIRCLE;
    mat_apply(M, t, g.p[0], g.p[1], tok.p);
    tok.p[2] = g.p[2];
    out.push_back(tok);
    return true;
  }
  case LG_GEO_RECT: {
    tok.kind = TOK_RECT;
    mat_apply(M, t, g.p[0], g.p[1], tok.p);
    Mat2 R{g.rot[0], g.rot[1], g.rot[2], g.rot[3]};
    Mat2 W = mat_mul(M, R);
    double hw = g.p[2] * 0.5, hh = g.p[3] * 0.5;
    tok.p[2] = W.m11 * hw; // u = W * (hw, 0)
    tok.p[3] = W.m21 * hw;
    tok.p[4] = W.m12 * hh; // v = W * (0, hh)
    tok.p[5] = W.m22 * hh;
    out.push_back(tok);
    return true;
  }
  case LG_GEO_SEGMENT: {
    tok.kind = TOK_SEGMENT;
    mat_apply(M, t, g.p[0], g.p[1], tok.p);
    mat_apply(M, t, g.p[2], g.p[3], tok.p + 2);
    out.push_back(tok);
    return true;
  }
  case LG_GEO_BEZIER: {
    tok.kind = TOK_BEZIER;
    for (int k = 0; k < 4; ++k) mat_apply(M, t, g.p[2 * k], g.p[2 * k + 1], tok.p + 2 * k);
    out.push_back(tok);
    return true;
  }
  case LG_GEO_ELLIPSE: { // centre, unit x axis of the ellipse in world space, semi axes
    tok.kind = TOK_ELLIPSE;
    mat_apply(M, t, g.p[0], g.p[1], tok.p);
    Mat2 R{g.rot[0], g.rot[1], g.rot[2], g.rot[3]};
    Mat2 W = mat_mul(M, R);
    tok.p[2] = W.m11;
    tok.p[3] = W.m21;
    tok.p[4] = g.p[2];
    tok.p[5] = g.p[3];
    out.push_back(tok);
    return true;
  }
  case LG_GEO_POLYGON: { // ORACLE.md §3.8: vertices given in hull order, local frame (origin, rot)
    const int nv = g.op;
    if (nv < 3 || nv > LG_POLYGON_MAX_VERTICES) return false;
    Mat2 R{g.rot[0], g.rot[1], g.rot[2], g.rot[3]};
    Mat2 W = mat_mul(M, R);
    double tw[2];
    mat_apply(M, t, g.p[0], g.p[1], tw);
    std::vector<double> xy; // world-space vertices, flattened
    for (int32_t pn = g.child_a; (int)xy.size() < 2 * nv;) {
      if (pn < 0 || (size_t)pn >= nodes.size()) return false;
      const LgGeoNode &pts = nodes[pn];
      if (pts.kind != LG_GEO_POINTS || pts.op < 1 || pts.op > 4) return false;
      for (int q = 0; q < pts.op && (int)xy.size() < 2 * nv; ++q) {
        double w[2];
        mat_apply(W, tw, pts.p[2 * q], pts.p[2 * q + 1], w);
        xy.push_back(w[0]);
        xy.push_back(w[1]);
      }
      pn = pts.child_a;
    }
    tok.kind = TOK_POLY;
    tok.op = nv;
    out.push_back(tok);
    for (int v = 0; v < nv; v += 4) {
      Token vt{};
      vt.kind = TOK_VERTS;
      vt.op = std::min(4, nv - v);
      for (int q = 0; q < vt.op; ++q) vt.p[2 * q] = xy[2 * (v + q)], vt.p[2 * q + 1] = xy[2 * (v + q) + 1];
      out.push_back(vt);
    }
    return true;
  }
  case LG_GEO_LOGIC: {
    Mat2 R{g.rot[0], g.rot[1], g.rot[2], g.rot[3]};
    Mat2 W = mat_mul(M, R);
    double tw[2];
    mat_apply(M, t, g.p[0], g.p[1], tw);
    int32_t a_start = (int32_t)out.size() - base;
    if (!lower_node(nodes, g.child_a, W, tw, out, base, depth + 1)) return false;
    int32_t b_start = (int32_t)out.size() - base;
    if (!lower_node(nodes, g.child_b, W, tw, out, base, depth + 1)) return false;
    tok.kind = TOK_OP;
    tok.op = g.op;
    tok.a_start = a_start;
    tok.b_start = b_start;
    if (g.op < LG_OP_AND || g.op > LG_OP_ANDNOT) return false;
    out.push_back(tok);
    return true;
  }
  default:
    return false;
  }
}

inline Scene make_scene(const LgObject *objs, uint32_t n_obj, const LgGeoNode *nodes, uint32_t n_nodes,
                        const LgTraceParams *prm) {
  Scene s;
  s.nodes.assign(nodes, nodes + n_nodes);
  s.raw_objects.assign(objs, objs + n_obj);
  if (prm) {
    s.max_bounce = prm->max_bounce;
    std::memcpy(s.cutoff, prm->cutoff_color, sizeof s.cutoff);
    std::memcpy(s.canvas_tlbr, prm->canvas_tlbr, sizeof s.canvas_tlbr);
  }
  const Mat2 I{1, 0, 0, 1};
  const double z[2] = {0, 0};
  for (uint32_t i = 0; i < n_obj; ++i) {
    Object o;
    o.first = (int32_t)s.tokens.size();
    if (!lower_node(s.nodes, objs[i].root, I, z, s.tokens, o.first, 0)) {
      s.ok = false;
      return s;
    }
    o.count = (int32_t)s.tokens.size() - o.first;
    if (o.count > MAX_TOKENS) {
      s.ok = false;
      return s;
    }
    o.has_material = objs[i].has_material != 0;
    o.n = objs[i].refractive_index;
    s.objects.push_back(o);
  }
  return s;
}

// ---------------------------------------------------------------------------
// Leaf tables in precision T (ORACLE.md §2.2: cast the f64 world parameters to
// T first, then derive r2 / e / uu / vv in T).
// ---------------------------------------------------------------------------
template <class T> struct LeafT {
  int32_t kind;
  int32_t op, a_start, b_start;
  T p[8];
  // CIRCLE : cx cy r r2
  // RECT   : cx cy ux uy vx vy uu vv
  // SEGMENT: ax ay ex ey
  // BEZIER : x0 y0 .. x3 y3
};

template <class T> struct SceneT {
  std::vector<LeafT<T>> tok;
  std::vector<Object> objects;
  std::vector<T> n; // refractive index per object in T
  uint32_t max_bounce;
  float cutoff[4];
  T canvas[8]; // rect form of the canvas: cx cy ux uy vx vy uu vv
};

template <class T> inline SceneT<T> cast_scene(const Scene &s) {
  SceneT<T> r;
  r.objects = s.objects;
  r.max_bounce = s.max_bounce;
  std::memcpy(r.cutoff, s.cutoff, sizeof r.cutoff);
  for (const Token &t : s.tokens) {
    LeafT<T> l{};
    l.kind = t.kind;
    l.op = t.op;
    l.a_start = t.a_start;
    l.b_start = t.b_start;
    switch (t.kind) {
    case TOK_CIRCLE:
      l.p[0] = (T)t.p[0];
      l.p[1] = (T)t.p[1];
      l.p[2] = (T)t.p[2];
      l.p[3] = l.p[2] * l.p[2];
      break;
    case TOK_RECT: {
      for (int k = 0; k < 6; ++k) l.p[k] = (T)t.p[k];
      V2<T> u{l.p[2], l.p[3]}, v{l.p[4], l.p[5]};
      l.p[6] = dot(u, u);
      l.p[7] = dot(v, v);
      break;
    }
    case TOK_SEGMENT:
      l.p[0] = (T)t.p[0];
      l.p[1] = (T)t.p[1];
      l.p[2] = (T)t.p[2] - l.p[0];
      l.p[3] = (T)t.p[3] - l.p[1];
      break;
    case TOK_BEZIER:
      for (int k = 0; k < 8; ++k) l.p[k] = (T)t.p[k];
      break;
    case TOK_ELLIPSE: // cx cy ux uy a b 1/a 1/b
      for (int k = 0; k < 6; ++k) l.p[k] = (T)t.p[k];
      l.p[6] = (T)1 / l.p[4];
      l.p[7] = (T)1 / l.p[5];
      break;
    case TOK_VERTS:
      for (int k = 0; k < 8; ++k) l.p[k] = (T)t.p[k];
      break;
    default:
      break;
    }
    r.tok.push_back(l);
  }
  for (const Object &o : s.objects) r.n.push_back((T)o.n);
  // canvas: Rect::from_tlbr(top, left, bottom, right), sub_render_pass.rs:156
  double top = s.canvas_tlbr[0], left = s.canvas_tlbr[1], bottom = s.canvas_tlbr[2], right = s.canvas_tlbr[3];
  r.canvas[0] = (T)((left + right) * 0.5);
  r.canvas[1] = (T)((top + bottom) * 0.5);
  r.canvas[2] = (T)((right - left) * 0.5);
  r.canvas[3] = (T)0;
  r.canvas[4] = (T)0;
  r.canvas[5] = (T)((top - bottom) * 0.5);
  r.canvas[6] = r.canvas[2] * r.canvas[2];
  r.canvas[7] = r.canvas[5] * r.canvas[5];
  return r;
}

// ---------------------------------------------------------------------------
// Primitive tests (ORACLE.md §3).  A hit is (t, point, unit normal).
// ---------------------------------------------------------------------------
template <class T> struct Hit {
  T t;
  V2<T> p;
  V2<T> n;
};
template <class T> struct HitList {
  Hit<T> h[4];
  int n = 0;
  void push(const Hit<T> &x) {
    if (n < 4) h[n++] = x;
  }
};

// §3.1 circle: cross-product discriminant (no m·m - tca² cancellation)
template <class T> inline void hit_circle(const T *c, V2<T> o, V2<T> d, HitList<T> &out) {
  V2<T> m{c[0] - o.x, c[1] - o.y};
  T cr = cross(m, d);
  T disc = std::fma(-cr, cr, c[3]);
  if (!(disc >= (T)0)) return;
  T tca = dot(m, d);
  T thc = std::sqrt(disc);
  T t0 = tca - thc, t1 = tca + thc;
  V2<T> ctr{c[0], c[1]};
  if (t0 > (T)T_MIN) {
    V2<T> p = along(o, t0, d);
    out.push({t0, p, normalize(sub(p, ctr))});
  }
  if (t1 > (T)T_MIN) {
    V2<T> p = along(o, t1, d);
    out.push({t1, p, normalize(sub(p, ctr))});
  }
}

// §3.2 segment a + u*e, u in [0,1] decided without a division
template <class T> inline bool hit_segment_ae(V2<T> a, V2<T> e, V2<T> o, V2<T> d, Hit<T> &h) {
  T denom = cross(d, e);
  T ad = std::fabs(denom);
  if (!(ad > (T)PAR_EPS)) return false;
  V2<T> w = sub(a, o);
  T s = cross(w, d);
  if (denom < (T)0) s = -s;
  if (!(s >= (T)0) || !(s <= ad)) return false;
  T t = cross(w, e) / denom;
  if (!(t > (T)T_MIN)) return false;
  h.t = t;
  h.p = along(o, t, d);
  h.n = normalize(V2<T>{-e.y, e.x});
  return true;
}
template <class T> inline void hit_segment(const T *s, V2<T> o, V2<T> d, HitList<T> &out) {
  Hit<T> h;
  if (hit_segment_ae(V2<T>{s[0], s[1]}, V2<T>{s[2], s[3]}, o, d, h)) out.push(h);
}

// §3.3 rect (centre c, half axes u, v): separating-axis reject along the ray
// normal, then the four edges in Rect::line_segments() order
// [right, bottom, left, top] (src/light_garden/grid.rs:31).
template <class T> inline void hit_rect(const T *r, V2<T> o, V2<T> d, HitList<T> &out) {
  V2<T> c{r[0], r[1]}, u{r[2], r[3]}, v{r[4], r[5]};
  V2<T> m = sub(c, o);
  T s = cross(d, m);
  T ext = std::fabs(cross(d, u)) + std::fabs(cross(d, v));
  if (!(std::fabs(s) <= ext)) return;
  V2<T> u2{u.x + u.x, u.y + u.y}, v2{v.x + v.x, v.y + v.y};
  V2<T> pmm{c.x - u.x - v.x, c.y - u.y - v.y}; // c - u - v
  V2<T> ppm{c.x + u.x - v.x, c.y + u.y - v.y}; // c + u - v
  V2<T> pmp{c.x - u.x + v.x, c.y - u.y + v.y}; // c - u + v
  Hit<T> h;
  if (hit_segment_ae(ppm, v2, o, d, h)) out.push(h); // right
  if (hit_segment_ae(pmm, u2, o, d, h)) out.push(h); // bottom
  if (hit_segment_ae(pmm, v2, o, d, h)) out.push(h); // left
  if (hit_segment_ae(pmp, u2, o, d, h)) out.push(h); // top
}

// §3.4 cubic Bézier: signed distances of the control points to the ray line,
// monotone intervals from the derivative's roots, fixed-count bisection.
template <class T> struct BezIters {};
template <> struct BezIters<float> { static constexpr int N = 28; };
template <> struct BezIters<double> { static constexpr int N = 56; };

template <class T> inline T horner3(T c3, T c2, T c1, T c0, T t) {
  return std::fma(std::fma(std::fma(c3, t, c2), t, c1), t, c0);
}

template <class T> inline void hit_bezier(const T *b, V2<T> o, V2<T> d, HitList<T> &out) {
  T y[4], x[4];
  for (int i = 0; i < 4; ++i) {
    V2<T> q{b[2 * i] - o.x, b[2 * i + 1] - o.y};
    y[i] = cross(d, q);
    x[i] = dot(d, q);
  }
  bool allpos = y[0] > 0 && y[1] > 0 && y[2] > 0 && y[3] > 0;
  bool allneg = y[0] < 0 && y[1] < 0 && y[2] < 0 && y[3] < 0;
  if (allpos || allneg) return;
  if (!(x[0] > (T)T_MIN) && !(x[1] > (T)T_MIN) && !(x[2] > (T)T_MIN) && !(x[3] > (T)T_MIN)) return;
  // power basis
  T c0 = y[0];
  T c1 = (T)3 * (y[1] - y[0]);
  T c2 = (T)3 * ((y[0] - (y[1] + y[1])) + y[2]);
  T c3 = (y[3] - y[0]) + (T)3 * (y[1] - y[2]);
  T e0 = x[0];
  T e1 = (T)3 * (x[1] - x[0]);
  T e2 = (T)3 * ((x[0] - (x[1] + x[1])) + x[2]);
  T e3 = (x[3] - x[0]) + (T)3 * (x[1] - x[2]);
  // critical points of y(t): A t^2 + B t + C = 0
  T A = (T)3 * c3, B = c2 + c2, C = c1;
  T split[4];
  int ns = 0;
  split[ns++] = (T)0;
  T r1 = (T)-1, r2 = (T)-1;
  if (A != (T)0) {
    T D = std::fma(B, B, -((T)4 * A * C));
    if (D > (T)0) {
      T sq = std::sqrt(D);
      T q = (T)-0.5 * (B + (B < (T)0 ? -sq : sq));
      r1 = q / A;
      if (q != (T)0) r2 = C / q;
    }
  } else if (B != (T)0) {
    r1 = -C / B;
  }
  if (r1 > r2) {
    T tmp = r1;
    r1 = r2;
    r2 = tmp;
  }
  if (r1 > (T)0 && r1 < (T)1) split[ns++] = r1;
  if (r2 > (T)0 && r2 < (T)1 && r2 != r1) split[ns++] = r2;
  split[ns++] = (T)1;
  for (int k = 0; k + 1 < ns; ++k) {
    T lo = split[k], hi = split[k + 1];
    T flo = horner3(c3, c2, c1, c0, lo), fhi = horner3(c3, c2, c1, c0, hi);
    bool nlo = flo < (T)0, nhi = fhi < (T)0;
    if (nlo == nhi) continue;
    for (int it = 0; it < BezIters<T>::N; ++it) {
      T mid = (T)0.5 * (lo + hi);
      T fm = horner3(c3, c2, c1, c0, mid);
      if ((fm < (T)0) == nlo)
        lo = mid;
      else
        hi = mid;
    }
    T tt = (T)0.5 * (lo + hi);
    T s = horner3(e3, e2, e1, e0, tt);
    if (!(s > (T)T_MIN)) continue;
    // tangent B'(t) in world space
    T om = (T)1 - tt;
    T w0 = om * om, w1 = (om + om) * tt, w2 = tt * tt;
    V2<T> d0{b[2] - b[0], b[3] - b[1]}, d1{b[4] - b[2], b[5] - b[3]}, d2{b[6] - b[4], b[7] - b[5]};
    V2<T> tg{std::fma(w0, d0.x, std::fma(w1, d1.x, w2 * d2.x)), std::fma(w0, d0.y, std::fma(w1, d1.y, w2 * d2.y))};
    out.push({s, along(o, s, d), normalize(V2<T>{-tg.y, tg.x})});
  }
}

// §3.7 ellipse (centre c, unit axis u, semi axes a, b): the ray in the frame where the ellipse is the unit circle
template <class T> inline V2<T> ellipse_local(const T *e, V2<T> w) { // world vector -> unit-circle frame
  V2<T> u{e[2], e[3]}, up{-e[3], e[2]};
  return {dot(w, u) * e[6], dot(w, up) * e[7]};
}
template <class T> inline V2<T> ellipse_normal(const T *e, V2<T> p) {
  V2<T> u{e[2], e[3]}, up{-e[3], e[2]};
  V2<T> l = ellipse_local(e, V2<T>{p.x - e[0], p.y - e[1]});
  T gx = l.x * e[6], gy = l.y * e[7]; // gradient of x^2/a^2 + y^2/b^2 in the ellipse frame
  return normalize(V2<T>{std::fma(gx, u.x, gy * up.x), std::fma(gx, u.y, gy * up.y)});
}
template <class T> inline void hit_ellipse(const T *e, V2<T> o, V2<T> d, HitList<T> &out) {
  V2<T> lo = ellipse_local(e, V2<T>{o.x - e[0], o.y - e[1]});
  V2<T> ld = ellipse_local(e, d);
  T A = dot(ld, ld), B = dot(lo, ld), C = dot(lo, lo) - (T)1;
  T disc = std::fma(B, B, -(A * C));
  if (!(disc >= (T)0) || !(A > (T)0)) return;
  T sq = std::sqrt(disc);
  T t0 = (-B - sq) / A, t1 = (-B + sq) / A;
  if (t0 > (T)T_MIN) {
    V2<T> p = along(o, t0, d);
    out.push({t0, p, ellipse_normal(e, p)});
  }
  if (t1 > (T)T_MIN) {
    V2<T> p = along(o, t1, d);
    out.push({t1, p, ellipse_normal(e, p)});
  }
}

// §3.8 convex polygon: vertex i of the polygon whose header token is `hdr` (vertices sit in the tokens after it)
template <class T> inline V2<T> polygon_vertex(const LeafT<T> *hdr, int i) {
  const LeafT<T> &blk = hdr[1 + i / 4];
  return {blk.p[2 * (i % 4)], blk.p[2 * (i % 4) + 1]};
}
// the edges v_i -> v_(i+1) as §3.2 segments, in hull order, closing edge last
template <class T> inline void hit_polygon(const LeafT<T> *hdr, V2<T> o, V2<T> d, HitList<T> &out) {
  const int nv = hdr->op;
  for (int i = 0; i < nv; ++i) {
    V2<T> a = polygon_vertex(hdr, i), b = polygon_vertex(hdr, (i + 1) % nv);
    Hit<T> h;
    if (hit_segment_ae(a, sub(b, a), o, d, h)) out.push(h);
  }
}
// inside = strictly on one side of all edges (either winding)
template <class T> inline bool polygon_contains(const LeafT<T> *hdr, V2<T> p) {
  const int nv = hdr->op;
  bool all_left = true, all_right = true;
  for (int i = 0; i < nv; ++i) {
    V2<T> a = polygon_vertex(hdr, i), b = polygon_vertex(hdr, (i + 1) % nv);
    T side = cross(sub(b, a), sub(p, a));
    all_left = all_left && side > (T)0;
    all_right = all_right && side < (T)0;
  }
  return all_left || all_right;
}

// §3.5 contains
template <class T> inline bool contains_leaf(const LeafT<T> &l, V2<T> p) {
  switch (l.kind) {
  case TOK_POLY:
    return polygon_contains(&l, p);
  case TOK_ELLIPSE: {
    V2<T> q = ellipse_local(l.p, V2<T>{p.x - l.p[0], p.y - l.p[1]});
    return dot(q, q) < (T)1;
  }
  case TOK_CIRCLE: {
    V2<T> q{p.x - l.p[0], p.y - l.p[1]};
    return dot(q, q) < l.p[3];
  }
  case TOK_RECT: {
    V2<T> q{p.x - l.p[0], p.y - l.p[1]};
    T a = dot(q, V2<T>{l.p[2], l.p[3]});
    T b = dot(q, V2<T>{l.p[4], l.p[5]});
    return std::fabs(a) < l.p[6] && std::fabs(b) < l.p[7];
  }
  default:
    return false; // mirrors: src/light_garden/object.rs:243-244
  }
}

// postfix evaluation of tokens [s, e] of one object with a bit stack
template <class T> inline bool contains_range(const LeafT<T> *tok, int s, int e, V2<T> p) {
  uint64_t st = 0;
  for (int i = s; i <= e; ++i) {
    const LeafT<T> &l = tok[i];
    if (l.kind == TOK_OP) {
      bool b = st & 1, a = (st >> 1) & 1;
      st >>= 2;
      bool r = l.op == LG_OP_AND ? (a && b) : l.op == LG_OP_OR ? (a || b) : (a && !b);
      st = (st << 1) | (r ? 1 : 0);
    } else if (l.kind == TOK_VERTS) {
      continue; // data of the polygon before it
    } else {
      st = (st << 1) | (contains_leaf(l, p) ? 1 : 0);
    }
  }
  return st & 1;
}

template <class T> inline bool contains_object(const SceneT<T> &s, int obj, V2<T> p) {
  const Object &o = s.objects[obj];
  return contains_range(&s.tok[o.first], 0, o.count - 1, p);
}

// §3.6 object intersection = leaf hits in program order, filtered up the tree
template <class T> struct ObjHits {
  Hit<T> h[MAX_TOKENS * 2];
  int n = 0;
};

template <class T> inline void intersect_object(const SceneT<T> &s, int obj, V2<T> o, V2<T> d, ObjHits<T> &out) {
  const Object &ob = s.objects[obj];
  const LeafT<T> *tok = &s.tok[ob.first];
  out.n = 0;
  for (int k = 0; k < ob.count; ++k) {
    const LeafT<T> &l = tok[k];
    if (l.kind == TOK_OP || l.kind == TOK_VERTS) continue;
    HitList<T> hl;
    switch (l.kind) {
    case TOK_POLY: hit_polygon(&l, o, d, hl); break;
    case TOK_CIRCLE: hit_circle(l.p, o, d, hl); break;
    case TOK_RECT: hit_rect(l.p, o, d, hl); break;
    case TOK_SEGMENT: hit_segment(l.p, o, d, hl); break;
    case TOK_BEZIER: hit_bezier(l.p, o, d, hl); break;
    case TOK_ELLIPSE: hit_ellipse(l.p, o, d, hl); break;
    }
    for (int j = 0; j < hl.n; ++j) {
      bool keep = true;
      bool flip = false;
      for (int i = k + 1; i < ob.count && keep; ++i) {
        const LeafT<T> &q = tok[i];
        if (q.kind != TOK_OP || q.a_start > k) continue;
        // q is an ancestor of leaf k
        if (k < q.b_start) { // hit comes from subtree a
          bool inb = contains_range(tok, q.b_start, i - 1, hl.h[j].p);
          keep = (q.op == LG_OP_AND) ? inb : !inb;
        } else { // from subtree b
          bool ina = contains_range(tok, q.a_start, q.b_start - 1, hl.h[j].p);
          keep = (q.op == LG_OP_OR) ? !ina : ina;
          if (q.op == LG_OP_ANDNOT) flip = !flip;
        }
      }
      if (keep && out.n < MAX_TOKENS * 2) {
        Hit<T> h = hl.h[j];
        if (flip) h.n = {-h.n.x, -h.n.y};
        out.h[out.n++] = h;
      }
    }
  }
}

// ---------------------------------------------------------------------------
// §4 reflect / refract (collision2d Ray::reflect / Ray::refract, called at
// src/light_garden/tracer.rs:444-449,477)
// ---------------------------------------------------------------------------
template <class T> inline V2<T> orient(V2<T> d, V2<T> n) {
  if (dot(d, n) > (T)0) return {-n.x, -n.y};
  return n;
}
template <class T> inline V2<T> reflect_dir(V2<T> d, V2<T> n_in) {
  V2<T> n = orient(d, n_in);
  T k = dot(d, n);
  T k2 = k + k;
  return normalize(V2<T>{std::fma(-k2, n.x, d.x), std::fma(-k2, n.y, d.y)});
}
// returns reflectance; has_refr false on total internal reflection
template <class T>
inline T refract_dir(V2<T> d, V2<T> n_in, T n1, T n2, V2<T> &refl, V2<T> &refr, bool &has_refr) {
  V2<T> n = orient(d, n_in);
  refl = reflect_dir(d, n_in);
  T eta = n1 / n2;
  T cosi = -dot(d, n);
  T sin2t = (eta * eta) * std::fma(-cosi, cosi, (T)1);
  if (sin2t > (T)1) {
    has_refr = false;
    return (T)1;
  }
  T cost = std::sqrt((T)1 - sin2t);
  T k = std::fma(eta, cosi, -cost);
  refr = normalize(V2<T>{std::fma(eta, d.x, k * n.x), std::fma(eta, d.y, k * n.y)});
  has_refr = true;
  T a = n1 * cosi, b = n2 * cost, c = n1 * cost, e = n2 * cosi;
  T ds = a + b, dp = c + e;
  if (ds == (T)0 || dp == (T)0) return (T)1;
  T rs = (a - b) / ds, rp = (c - e) / dp;
  return (T)0.5 * std::fma(rs, rs, rp * rp);
}

// ---------------------------------------------------------------------------
// §5 Tracer::trace — src/light_garden/tracer.rs:360-493, brute-force branch
// (412-424).  Breadth-first, two work lists, exactly as the reference.
// ---------------------------------------------------------------------------
struct SegOut {
  LgSegment seg;
  LgSegmentTag tag;
  LgSegmentF64 f64;
};

struct TraceCounters {
  uint64_t ray_steps = 0, segments = 0;
  uint64_t object_tests = 0; // Ray::intersect(&obj.get_geometry()) calls of the nearest-hit search
};

// ---------------------------------------------------------------------------
// TileMap (src/light_garden/tile_map.rs), ORACLE.md §5.6.  The reference's CPU-side culling, ENABLED by default
// (tile_map.rs:61): the window is cut into num_tilesx x num_tilesy tiles (Tracer::new: 100 x 100, tracer.rs:27),
// every tile has num_slabs (8) angular sectors; a sector lists the objects whose bounding box can be seen from the
// tile in that range of directions, and the tile lists the objects whose box overlaps it.  A ray tests the sector of
// its direction in the tile of its origin plus the tile's overlaps (tracer.rs:385-411) instead of every object.
// The box-to-box angular range (collision2d Aabb::get_crossover) and the boxes themselves (HasAabb) are not in the
// repository: here a box is the hull of the object's leaves and the range is the hull of the 16 corner-to-corner
// directions, widened by 1e-9 rad -- conservative, so the nearest hit is the one of the all-objects loop.
// ---------------------------------------------------------------------------
struct TileMap {
  double w = 0, h = 0;
  int nx = 0, ny = 0, ns = 0;
  std::vector<uint64_t> start; // (tile * ns + slab) -> range of `cand`
  std::vector<int32_t> cand;   // ascending object indices: the sector's objects and the tile's overlaps, merged
  bool enabled = false;

  static long to_usize(double v) { return !(v > 0.0) ? 0 : (v >= 4e18 ? (long)4e18 : (long)v); } // Rust `as usize`
  // TileMap::get_tile, tile_map.rs:134-146 (the window is centred on the origin)
  int tile_of(double x, double y) const {
    const long ix = to_usize((x + w * 0.5) / (w / nx)), iy = to_usize((y + h * 0.5) / (h / ny));
    return (ix < nx && iy < ny) ? (int)(ix + iy * nx) : -1;
  }
  // clockwise angle from +y, Tile::get_index tile_map.rs:229-235
  static double angle_of(double dx, double dy) {
    const double TAU = 6.283185307179586;
    double a = std::acos(dy > 1.0 ? 1.0 : (dy < -1.0 ? -1.0 : dy));
    if (dx < 0.0) a = TAU - a;
    return a;
  }
  int slab_of_angle(double a) const {
    const double TAU = 6.283185307179586;
    const long k = to_usize((double)ns * a / TAU - 2.220446049250313e-16);
    return (int)(k >= ns ? ns - 1 : k);
  }
  int slab_of(double dx, double dy) const { return slab_of_angle(angle_of(dx, dy)); }
};

struct Box {
  double x0 = 1e300, y0 = 1e300, x1 = -1e300, y1 = -1e300;
  void add(double x, double y) { x0 = std::min(x0, x), y0 = std::min(y0, y), x1 = std::max(x1, x), y1 = std::max(y1, y); }
};
// hull of the object's leaves (a superset of the object for every LogicOp)
inline Box object_box(const Scene &s, int ob) {
  Box b;
  const Object &o = s.objects[ob];
  for (int k = o.first; k < o.first + o.count; ++k) {
    const Token &t = s.tokens[k];
    switch (t.kind) {
    case TOK_CIRCLE:
      b.add(t.p[0] - t.p[2], t.p[1] - t.p[2]), b.add(t.p[0] + t.p[2], t.p[1] + t.p[2]);
      break;
    case TOK_RECT:
      for (int su = -1; su <= 1; su += 2)
        for (int sv = -1; sv <= 1; sv += 2) b.add(t.p[0] + su * t.p[2] + sv * t.p[4], t.p[1] + su * t.p[3] + sv * t.p[5]);
      break;
    case TOK_SEGMENT:
      b.add(t.p[0], t.p[1]), b.add(t.p[2], t.p[3]);
      break;
    case TOK_BEZIER: // the curve lies in the hull of its control points
      for (int q = 0; q < 4; ++q) b.add(t.p[2 * q], t.p[2 * q + 1]);
      break;
    case TOK_ELLIPSE: {
      const double r = std::max(std::fabs(t.p[4]), std::fabs(t.p[5]));
      b.add(t.p[0] - r, t.p[1] - r), b.add(t.p[0] + r, t.p[1] + r);
      break;
    }
    case TOK_VERTS:
      for (int q = 0; q < t.op; ++q) b.add(t.p[2 * q], t.p[2 * q + 1]);
      break;
    default:
      break;
    }
  }
  return b;
}

inline TileMap build_tile_map(const Scene &s, int nx, int ny, int ns) {
  const double TAU = 6.283185307179586, PI = 3.141592653589793;
  TileMap tm;
  tm.nx = nx, tm.ny = ny, tm.ns = ns;
  tm.w = s.canvas_tlbr[3] - s.canvas_tlbr[1]; // canvas_bounds.width, tracer.rs:27
  tm.h = s.canvas_tlbr[0] - s.canvas_tlbr[2];
  const int nobj = (int)s.objects.size(), ntiles = nx * ny;
  std::vector<Box> boxes(nobj);
  for (int i = 0; i < nobj; ++i) boxes[i] = object_box(s, i);
  std::vector<std::vector<int32_t>> per_tile(ntiles);       // candidates of the tile's sectors, concatenated
  std::vector<std::vector<uint32_t>> per_tile_start(ntiles); // ns + 1 offsets into per_tile[t]
  const double stepx = tm.w / nx, stepy = tm.h / ny;
#pragma omp parallel for schedule(dynamic, 16)
  for (int t = 0; t < ntiles; ++t) {
    // TileMap::get_aabb, tile_map.rs:65-78
    const double left = (t % nx) * stepx - tm.w * 0.5, bottom = (t / nx) * stepy - tm.h * 0.5;
    const double tx[2] = {left, left + stepx}, ty[2] = {bottom, bottom + stepy};
    const double tcx = left + 0.5 * stepx, tcy = bottom + 0.5 * stepy;
    std::vector<std::vector<int32_t>> slab(ns);
    std::vector<int32_t> ovl;
    for (int i = 0; i < nobj; ++i) {
      const Box &b = boxes[i];
      // Tile::update_overlap, tile_map.rs:237-251: the boxes touch -> the object is tested from every sector
      if (!(b.x0 > tx[1] || b.x1 < tx[0] || b.y0 > ty[1] || b.y1 < ty[0])) {
        ovl.push_back(i);
        continue;
      }
      // Tile::get_range, tile_map.rs:296-331: the sectors between the two extreme directions
      const double ocx = 0.5 * (b.x0 + b.x1), ocy = 0.5 * (b.y0 + b.y1);
      const double rl = std::hypot(ocx - tcx, ocy - tcy);
      const double ar = TileMap::angle_of((ocx - tcx) / rl, (ocy - tcy) / rl);
      const double ox[2] = {b.x0, b.x1}, oy[2] = {b.y0, b.y1};
      double lo = 0, hi = 0;
      for (int q = 0; q < 16; ++q) {
        const double dx = ox[q & 1] - tx[(q >> 2) & 1], dy = oy[(q >> 1) & 1] - ty[(q >> 3) & 1];
        const double l = std::hypot(dx, dy);
        if (!(l > 0)) continue;
        double da = TileMap::angle_of(dx / l, dy / l) - ar;
        if (da > PI) da -= TAU;
        if (da < -PI) da += TAU;
        lo = std::min(lo, da), hi = std::max(hi, da);
      }
      lo -= 1e-9, hi += 1e-9;
      auto wrap = [&](double a) { return a < 0 ? a + TAU : (a >= TAU ? a - TAU : a); };
      const int k0 = tm.slab_of_angle(wrap(ar + lo)), k1 = tm.slab_of_angle(wrap(ar + hi));
      for (int k = k0;; k = (k + 1) % ns) {
        slab[k].push_back(i);
        if (k == k1) break;
      }
    }
    std::vector<int32_t> &out = per_tile[t];
    std::vector<uint32_t> &st = per_tile_start[t];
    st.assign(ns + 1, 0);
    for (int k = 0; k < ns; ++k) {
      st[k] = (uint32_t)out.size();
      out.resize(out.size() + slab[k].size() + ovl.size());
      std::merge(slab[k].begin(), slab[k].end(), ovl.begin(), ovl.end(), out.begin() + st[k]); // both ascending, disjoint
    }
    st[ns] = (uint32_t)out.size();
  }
  tm.start.assign((size_t)ntiles * ns + 1, 0);
  uint64_t total = 0;
  for (int t = 0; t < ntiles; ++t) {
    for (int k = 0; k < ns; ++k) tm.start[(size_t)t * ns + k] = total + per_tile_start[t][k];
    total += per_tile[t].size();
  }
  tm.start[(size_t)ntiles * ns] = total;
  tm.cand.resize(total);
  total = 0;
  for (int t = 0; t < ntiles; ++t) {
    std::copy(per_tile[t].begin(), per_tile[t].end(), tm.cand.begin() + total);
    total += per_tile[t].size();
    std::vector<int32_t>().swap(per_tile[t]);
  }
  tm.enabled = true;
  return tm;
}

template <class T> struct Item {
  V2<T> o, d;
  float c[4];
  T n;
  uint64_t path;
};

template <class T>
inline void emit(std::vector<SegOut> *out, TraceCounters &cnt, V2<T> a, V2<T> b, const float c[4], uint64_t ray,
                 uint32_t gen, uint64_t path, int32_t hit) {
  cnt.segments++;
  if (!out) return;
  SegOut s;
  s.seg.a[0] = (float)a.x; // `p.x as f32`, sub_render_pass.rs:192
  s.seg.a[1] = (float)a.y;
  s.seg.b[0] = (float)b.x;
  s.seg.b[1] = (float)b.y;
  std::memcpy(s.seg.color, c, 16);
  s.tag = {ray, path, gen, hit};
  s.f64 = {{(double)a.x, (double)a.y}, {(double)b.x, (double)b.y}};
  out->push_back(s);
}

template <class T>
inline void trace_ray(const SceneT<T> &s, const LgRay &ray, uint64_t ray_id, std::vector<SegOut> *out,
                      TraceCounters &cnt, const TileMap *tm = nullptr) {
  std::vector<Item<T>> cur, next; // trace_rays / back_buffer, tracer.rs:368-369
  Item<T> it0;
  it0.o = {(T)ray.origin[0], (T)ray.origin[1]};
  it0.d = {(T)ray.direction[0], (T)ray.direction[1]};
  std::memcpy(it0.c, ray.color, 16);
  it0.n = (T)ray.refractive_index;
  it0.path = 0;
  cur.push_back(it0);
  const int nobj = (int)s.objects.size();
  ObjHits<T> oh;
  for (uint32_t gen = 0; gen < s.max_bounce; ++gen) { // tracer.rs:373
    if (cur.empty()) return;                         // 374-376
    for (const Item<T> &it : cur) {                  // 377
      const float *c = it.c;
      if ((c[0] < s.cutoff[0] && c[1] < s.cutoff[1] && c[2] < s.cutoff[2]) || c[3] < s.cutoff[3]) continue; // 378-384
      cnt.ray_steps++;
      // nearest hit, tracer.rs:412-424: strict `<` on distance_squared
      T nearest = std::numeric_limits<T>::max();
      int best = -1;
      Hit<T> bh{};
      // tracer.rs:385-411: with the TileMap, the sector of the ray's direction in the tile of its origin plus the
      // tile's overlaps (ascending object index here, so ties resolve as in the all-objects loop); a ray that starts
      // outside the window tests every object (the reference would test none: its origins are always inside)
      const int32_t *cl = nullptr;
      int ncl = nobj;
      if (tm && tm->enabled) {
        const int tile = tm->tile_of((double)it.o.x, (double)it.o.y);
        if (tile >= 0) {
          const size_t q = (size_t)tile * tm->ns + tm->slab_of((double)it.d.x, (double)it.d.y);
          cl = tm->cand.data() + tm->start[q];
          ncl = (int)(tm->start[q + 1] - tm->start[q]);
        }
      }
      cnt.object_tests += (uint64_t)ncl;
      for (int q = 0; q < ncl; ++q) {
        const int ix = cl ? cl[q] : q;
        intersect_object(s, ix, it.o, it.d, oh);
        for (int j = 0; j < oh.n; ++j) {
          T dx = oh.h[j].p.x - it.o.x, dy = oh.h[j].p.y - it.o.y;
          T d2 = dx * dx + dy * dy; // nalgebra distance_squared: no fma
          if (d2 < nearest) {
            nearest = d2;
            best = ix;
            bh = oh.h[j];
          }
        }
      }
      if (best >= 0) {
        const Object &ob = s.objects[best];
        if (ob.has_material) { // tracer.rs:428
          T n2 = (T)1;         // air, 430
          // tracer.rs:431 obj.contains(&ray.get_origin()) — evaluated at the
          // midpoint of (origin, hit): ORACLE.md §5.2
          V2<T> mid{(it.o.x + bh.p.x) * (T)0.5, (it.o.y + bh.p.y) * (T)0.5};
          if (contains_object(s, best, mid)) {
            for (int ix = 0; ix < nobj; ++ix) { // 432-439
              if (ix != best && contains_object(s, ix, bh.p)) {
                if (s.objects[ix].has_material) {
                  n2 = s.n[ix];
                  break;
                }
              }
            }
          } else {
            n2 = s.n[best]; // 441
          }
          V2<T> rfl, rfr;
          bool has;
          T R = refract_dir(it.d, bh.n, it.n, n2, rfl, rfr, has); // 444-450
          emit(out, cnt, it.o, bh.p, c, ray_id, gen, it.path, best); // 451-452
          float refl = (float)R;                                     // 454
          float om = 1.f - refl;                                     // 455
          Item<T> a;
          a.o = bh.p;
          a.d = rfl;
          a.c[0] = c[0] * refl;
          a.c[1] = c[1] * refl;
          a.c[2] = c[2] * refl;
          a.c[3] = c[3];
          a.n = it.n;
          a.path = it.path << 1;
          next.push_back(a); // 456-458
          if (has) {         // 460-472
            Item<T> b;
            b.o = bh.p;
            b.d = rfr;
            b.c[0] = c[0] * om;
            b.c[1] = c[1] * om;
            b.c[2] = c[2] * om;
            b.c[3] = c[3];
            b.n = n2;
            b.path = (it.path << 1) | 1;
            next.push_back(b);
          }
        } else { // mirror, 473-481
          emit(out, cnt, it.o, bh.p, c, ray_id, gen, it.path, best);
          Item<T> a = it;
          a.o = bh.p;
          a.d = reflect_dir(it.d, bh.n);
          a.path = it.path << 1;
          next.push_back(a);
        }
      } else { // canvas, 482-488
        HitList<T> hl;
        hit_rect(s.canvas, it.o, it.d, hl);
        if (hl.n > 0) {
          int f = 0;
          for (int j = 1; j < hl.n; ++j)
            if (hl.h[j].t < hl.h[f].t) f = j; // get_first(): nearest, ORACLE.md §5.4
          emit(out, cnt, it.o, hl.h[f].p, c, ray_id, gen, it.path, -1);
        }
      }
    }
    cur.clear(); // 490-491
    std::swap(cur, next);
  }
}

// ---------------------------------------------------------------------------
// §6 ray emission — src/light_garden/light.rs (always f64, like the reference)
// ---------------------------------------------------------------------------
inline void unit_from(double x, double y, double out[2]) { // Ray::from_origin -> Unit::new_normalize
  double n = std::sqrt(x * x + y * y);
  out[0] = x / n;
  out[1] = y / n;
}
inline double signum(double x) { // Rust f64::signum
  if (std::isnan(x)) return x;
  return std::signbit(x) ? -1.0 : 1.0;
}

inline void emit_ray(const LgLight &l, uint64_t i, LgRay &r) {
  const double PI = 3.14159265358979323846;
  std::memcpy(r.color, l.color, 16);
  r.refractive_index = 1.0;
  double n = (double)l.num_rays;
  switch (l.kind) {
  case LG_LIGHT_POINT: { // light.rs:163-174
    double f = (double)i * PI * 2. / n;
    double s = std::sin(f), c = std::cos(f);
    r.origin[0] = l.position[0];
    r.origin[1] = l.position[1];
    unit_from(c, s, r.direction);
    break;
  }
  case LG_LIGHT_SPOT: { // light.rs:225-249
    double dx = l.spot_direction[0], dy = l.spot_direction[1];
    double da = std::fabs(dx) < EPSILON ? (dy >= 0. ? PI * 0.5 : -PI * 0.5) : std::atan(dy / dx);
    double min_angle = da - 0.5 * l.spot_angle;
    uint64_t step = i + 1; // for step in 1..=num_rays
    double angle = min_angle + ((double)step / n) * l.spot_angle;
    double yd = std::sin(angle), xd = std::cos(angle);
    double sg = signum(dx);
    r.origin[0] = l.position[0];
    r.origin[1] = l.position[1];
    unit_from(sg * xd, sg * yd, r.direction);
    break;
  }
  default: { // Directional, light.rs:103-115 + ORACLE.md §6.3
    double ex = l.b[0] - l.position[0], ey = l.b[1] - l.position[1];
    // start.eval_at_r(-(i as f64) / n), light.rs:111: origins walk the segment by default; LG_LIGHT_DIRECTIONAL_NEG_R
    // takes the call literally on eval_at_r(r) = a + r (b - a)
    double rr = ((l.flags & LG_LIGHT_DIRECTIONAL_NEG_R) ? -1.0 : 1.0) * ((double)i / n);
    r.origin[0] = l.position[0] + rr * ex;
    r.origin[1] = l.position[1] + rr * ey;
    unit_from(-ey, ex, r.direction);
    break;
  }
  }
}

// start medium of a light: src/light_garden/tracer.rs:280-287 (last match wins)
inline double start_medium(const SceneT<double> &s, const LgLight &l) {
  double n = 1.;
  // Light::get_origin: DirectionalLight -> start.get_origin() (light.rs:118-121),
  // taken as the segment's point a (ORACLE.md §6.3)
  V2<double> p{l.position[0], l.position[1]};
  for (size_t i = 0; i < s.objects.size(); ++i)
    if (contains_object(s, (int)i, p) && s.objects[i].has_material) n = s.objects[i].n;
  // `.chain(&self.drawing_object)` (tracer.rs:281): the host evaluated that last link of the chain itself
  if (l.flags & LG_LIGHT_START_MEDIUM) n = l.start_medium;
  return n;
}

// shard of n rays for rank r of w: the rays r, r + w, r + 2w, ...
inline uint64_t shard_count(uint64_t n, uint32_t r, uint32_t w) { return n > r ? (n - r + w - 1) / w : 0; }

// ---------------------------------------------------------------------------
// §7 string mod — src/light_garden/string_mod.rs:33-158 (Circle curve)
// ---------------------------------------------------------------------------
inline uint64_t wrapping_pow(uint64_t base, uint32_t exp) { // Rust u64::pow in release
  uint64_t acc = 1;
  while (exp) {
    if (exp & 1) acc *= base;
    base *= base;
    exp >>= 1;
  }
  return acc;
}
inline void sm_point(const LgStringMod &sm, uint64_t n, double out[2]) { // string_mod.rs:33-85
  const double TAU = 6.28318530717958647692;
  const uint64_t tn = sm.turns * n; // u64 product, wraps
  if (sm.curve == LG_CURVE_COMPLEX_EXP) {
    // complex.powu((self.turns * n) as u32) -> num_traits::pow::pow (exponentiation by squaring), products
    // (a+bi)(c+di) = (ac - bd) + (ad + bc)i in plain f64 (num-complex Mul)
    uint32_t e = (uint32_t)tn;
    double br = sm.curve_p[0], bi = sm.curve_p[1];
    if (e == 0) {
      out[0] = 1.0, out[1] = 0.0;
      return;
    }
    auto mul = [](double ar, double ai, double cr, double ci, double &rr, double &ri) {
      rr = ar * cr - ai * ci;
      ri = ar * ci + ai * cr;
    };
    while ((e & 1u) == 0u) {
      double r, i;
      mul(br, bi, br, bi, r, i);
      br = r, bi = i;
      e >>= 1;
    }
    double ar = br, ai = bi;
    while (e > 1u) {
      e >>= 1;
      double r, i;
      mul(br, bi, br, bi, r, i);
      br = r, bi = i;
      if (e & 1u) {
        mul(ar, ai, br, bi, r, i);
        ar = r, ai = i;
      }
    }
    out[0] = ar, out[1] = ai;
    return;
  }
  double angle = (double)tn * TAU / (double)sm.modulo;
  if (sm.curve == LG_CURVE_HYPOTROCHOID) { // string_mod.rs:56-71
    double small_radius = (double)(uint64_t)sm.curve_p[0], big_radius = (double)(uint64_t)sm.curve_p[1];
    double off_center = (double)(uint64_t)sm.curve_p[2];
    double smr = big_radius - small_radius, ratio = smr + off_center;
    double x = smr * std::cos(angle) + off_center * std::cos(angle * smr / small_radius);
    double y = smr * std::sin(angle) - off_center * std::sin(angle * smr / small_radius);
    out[0] = x / ratio, out[1] = y / ratio;
  } else if (sm.curve == LG_CURVE_LISSAJOUS) { // string_mod.rs:73-83
    double a = (double)(uint64_t)sm.curve_p[0], b = (double)(uint64_t)sm.curve_p[1];
    out[0] = std::sin(a * angle + sm.curve_p[2]);
    out[1] = std::sin(b * angle);
  } else { // Circle, string_mod.rs:45-55
    out[0] = std::cos(angle);
    out[1] = std::sin(angle);
  }
}
inline void sm_color(const LgStringMod &sm, const LgModRemColor *rules, uint32_t nr, uint64_t ix, float out[4]) {
  float c[4] = {0, 0, 0, 0}; // string_mod.rs:124-150
  int cnt = 0;
  for (uint32_t k = 0; k < nr; ++k) {
    if (rules[k].modulo != 0 && (ix % rules[k].modulo) == rules[k].rem) {
      for (int j = 0; j < 4; ++j) c[j] += rules[k].color[j];
      cnt++;
    }
  }
  if (cnt == 0) {
    std::memcpy(out, sm.color, 16);
  } else {
    for (int j = 0; j < 4; ++j) out[j] = c[j] / (float)cnt;
  }
}
inline uint64_t sm_target(const LgStringMod &sm, uint64_t iix) { // string_mod.rs:111-116
  uint64_t m = sm.modulo;
  switch (sm.mode) {
  case LG_SM_ADD: return (iix + sm.num) % m;
  case LG_SM_MUL: return (iix * sm.num) % m;
  case LG_SM_POW: return wrapping_pow(iix, (uint32_t)sm.num) % m;
  default: return wrapping_pow(sm.num, (uint32_t)iix) % m;
  }
}
inline void sm_chord(const LgStringMod &sm, const LgModRemColor *rules, uint32_t nr, uint64_t iix, LgVertexPair &vp) {
  uint64_t ix = sm_target(sm, iix);
  sm_point(sm, iix, vp.a);
  sm_point(sm, ix, vp.b);
  sm_color(sm, rules, nr, iix, vp.color_a);
  sm_color(sm, rules, nr, ix, vp.color_b);
}

// §7.2 nested string mod (string_mod.rs:87-101,152-158).  LineSegment::intersect(&LineSegment) of collision2d is
// taken as: both parameters in [0, 1] (end points included), none for parallel segments.
inline bool segment_segment(const LgVertexPair &p, const LgVertexPair &q, double out[2]) {
  V2<double> a1{p.a[0], p.a[1]}, e1{p.b[0] - p.a[0], p.b[1] - p.a[1]};
  V2<double> a2{q.a[0], q.a[1]}, e2{q.b[0] - q.a[0], q.b[1] - q.a[1]};
  double denom = cross(e1, e2);
  if (!(std::fabs(denom) > PAR_EPS)) return false;
  V2<double> w = sub(a2, a1);
  double t = cross(w, e2) / denom, u = cross(w, e1) / denom;
  if (!(t >= 0.0 && t <= 1.0 && u >= 0.0 && u <= 1.0)) return false;
  V2<double> x = along(a1, t, e1);
  out[0] = x.x, out[1] = x.y;
  return true;
}
// StringMod::line_crossings_as_points over the chords `lines` (their order is the reference's loop nest)
inline std::vector<double> line_crossings(const LgVertexPair *lines, uint64_t n) {
  std::vector<double> pts;
  for (uint64_t diff = 1; diff < n; ++diff)
    for (uint64_t ixa = 0; ixa < n; ++ixa) {
      double x[2];
      if (segment_segment(lines[ixa], lines[(ixa + diff) % n], x)) pts.push_back(x[0]), pts.push_back(x[1]);
    }
  return pts;
}
// inner.draw_init_points(points): chord iix joins points[iix % P] and points[f(iix) % P]
inline void nested_chord(const LgStringMod &inner, const LgModRemColor *rules, uint32_t nr, const double *pts, uint64_t P,
                         uint64_t iix, LgVertexPair &vp) {
  uint64_t ix = sm_target(inner, iix);
  vp.a[0] = pts[2 * (iix % P)], vp.a[1] = pts[2 * (iix % P) + 1];
  vp.b[0] = pts[2 * (ix % P)], vp.b[1] = pts[2 * (ix % P) + 1];
  sm_color(inner, rules, nr, iix, vp.color_a);
  sm_color(inner, rules, nr, ix, vp.color_b);
}

// ---------------------------------------------------------------------------
// §8 accumulate — sub_render_pass.rs:188-212 + shader.wgsl + blend mod.rs:57-73
// Non-AA 1-px LineList coverage, colour lerp, rgb += src.rgb, a += src.a².
// Image rows are y-down, 4 floats per pixel.  Only rows [row0, row1) are
// written so that callers can band the image across threads without changing
// the per-pixel order of adds.
// ---------------------------------------------------------------------------
struct Proj {
  float m00, m11, hw, hh;
  int W, H;
};
inline Proj make_proj(int W, int H) {
  Proj p;
  float aspect = (float)W / (float)H;      // sub_render_pass.rs:146
  p.m00 = 2.0f / (aspect - (-aspect));     // cgmath::ortho c0r0, renderer.rs:121
  p.m11 = 2.0f / (1.0f - (-1.0f));
  p.hw = (float)W * 0.5f;
  p.hh = (float)H * 0.5f;
  p.W = W;
  p.H = H;
  return p;
}
inline void to_pixel(const Proj &pr, float x, float y, float &px, float &py) {
  float xn = pr.m00 * x, yn = pr.m11 * y;
  px = std::fmaf(xn, pr.hw, pr.hw);
  py = std::fmaf(-yn, pr.hh, pr.hh);
}

template <class Acc>
inline uint64_t raster_segment(const Proj &pr, const float a[2], const float b[2], const float ca[4],
                               const float cb[4], int row0, int row1, Acc &&acc) {
  float x0, y0, x1, y1;
  to_pixel(pr, a[0], a[1], x0, y0);
  to_pixel(pr, b[0], b[1], x1, y1);
  float dx = x1 - x0, dy = y1 - y0;
  if (!(std::fabs(dx) < 1e30f) || !(std::fabs(dy) < 1e30f)) return 0; // NaN / inf vertices are dropped
  bool xmajor = std::fabs(dx) >= std::fabs(dy);
  float m0 = xmajor ? x0 : y0, m1 = xmajor ? x1 : y1; // major axis
  float n0 = xmajor ? y0 : x0, n1 = xmajor ? y1 : x1; // minor axis
  float dm = m1 - m0, dn = n1 - n0;
  if (dm == 0.f) return 0;
  float lo = m0 < m1 ? m0 : m1, hi = m0 < m1 ? m1 : m0;
  int Nmaj = xmajor ? pr.W : pr.H, Nmin = xmajor ? pr.H : pr.W;
  float flo = std::ceil(lo - 0.5f), fhi = std::ceil(hi - 0.5f);
  if (flo < 0.f) flo = 0.f;
  if (fhi > (float)Nmaj) fhi = (float)Nmaj;
  if (!(flo < fhi)) return 0;
  int i0 = (int)flo, i1 = (int)fhi;
  float inv = 1.0f / dm;
  float dc[4] = {cb[0] - ca[0], cb[1] - ca[1], cb[2] - ca[2], cb[3] - ca[3]};
  uint64_t n = 0;
  for (int i = i0; i < i1; ++i) {
    float mc = (float)i + 0.5f;
    float s = (mc - m0) * inv;
    float nv = std::fmaf(s, dn, n0);
    float fj = std::floor(nv);
    if (!(fj >= 0.f) || !(fj < (float)Nmin)) continue;
    int j = (int)fj;
    int px = xmajor ? i : j, py = xmajor ? j : i;
    if (py < row0 || py >= row1) continue;
    float c[4];
    for (int k = 0; k < 4; ++k) c[k] = std::fmaf(s, dc[k], ca[k]);
    acc(px, py, c);
    n++;
  }
  return n;
}

// ---- ORACLE.md 8.7: the 8-bit surface target (sub_render_pass.rs:59-63; Bgra8UnormSrgb, renderer.rs:207-209) ---------
// colour byte = number of k in 1..255 whose threshold T[k] = f32(linear value encoding to (k - 0.5) / 255) is <= c;
// alpha byte = floor(min(a, 1) * 255 + 0.5) in f32; anything not > 0 (NaN included) is 0.
inline void surface_thresholds(float thr[256]) {
  thr[0] = -INFINITY;
  for (int k = 1; k < 256; ++k) {
    const double enc = ((double)k - 0.5) / 255.0;
    thr[k] = (float)(enc <= 0.04045 ? enc / 12.92 : std::pow((enc + 0.055) / 1.055, 2.4));
  }
}
inline uint8_t surface_colour_byte(const float thr[256], float v) {
  int n = 0;
  for (int k = 1; k < 256; ++k) n += thr[k] <= v ? 1 : 0;
  return (uint8_t)n;
}
inline uint8_t surface_alpha_byte(float a) { return !(a > 0.f) ? 0 : (uint8_t)((a < 1.f ? a : 1.f) * 255.f + 0.5f); }

} // namespace lgo
