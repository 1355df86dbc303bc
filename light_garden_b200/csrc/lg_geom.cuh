// lg_geom.cuh — device-side ray/shape arithmetic for the trace kernel.
//
// Stands in for collision2d's Ray::intersect / Contains::contains /
// Ray::reflect / Ray::refract as they are called from
// src/light_garden/tracer.rs:399-449,477,484-486 (collision2d itself is an
// un-vendored git dependency, Cargo.lock:880-887; the formulas are the
// builder-specified ones of oracle/ORACLE.md §1-§4).
//
// Everything is templated on the real type (float = throughput mode, double =
// the reference's Float) and written with explicit fused multiply-adds only;
// the translation unit is compiled with -fmad=false so that no other
// contraction happens and results are bit-identical to any IEEE-754
// implementation of the same formulas.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

namespace lg {

constexpr double kTMin = 1e-5;    // ORACLE.md §1: accept a hit iff t > T_MIN
constexpr double kParEps = 1e-12; // ORACLE.md §1: |cross(d,e)| <= PAR_EPS is parallel

enum : int32_t { TOK_CIRCLE = 0, TOK_RECT = 1, TOK_SEGMENT = 2, TOK_BEZIER = 3, TOK_OP = 4, TOK_ELLIPSE = 5, TOK_POLY = 6, TOK_POINTS = 7 };
enum : int32_t { OP_AND = 0, OP_OR = 1, OP_ANDNOT = 2 };

// ---- real-type wrappers ----------------------------------------------------
// LG_HD: the geometry is also compilable for the host so that
// tests/test_geom_host.py can check this header against the oracle without a
// GPU.  No product entry point ever runs it on the host.
#define LG_HD __host__ __device__ __forceinline__
template <class T> struct Real;
template <> struct Real<float> {
#ifdef __CUDA_ARCH__
  static LG_HD float fma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
  static LG_HD float sqrt(float a) { return __fsqrt_rn(a); }
  static LG_HD float div(float a, float b) { return __fdiv_rn(a, b); }
#else
  static LG_HD float fma(float a, float b, float c) { return ::fmaf(a, b, c); }
  static LG_HD float sqrt(float a) { return ::sqrtf(a); }
  static LG_HD float div(float a, float b) { return a / b; }
#endif
  static LG_HD float abs(float a) { return ::fabsf(a); }
  static LG_HD float max_value() { return 3.402823466e+38f; }
  static constexpr int kBezIters = 28;
};
template <> struct Real<double> {
#ifdef __CUDA_ARCH__
  static LG_HD double fma(double a, double b, double c) { return __fma_rn(a, b, c); }
  static LG_HD double sqrt(double a) { return __dsqrt_rn(a); }
  static LG_HD double div(double a, double b) { return __ddiv_rn(a, b); }
#else
  static LG_HD double fma(double a, double b, double c) { return ::fma(a, b, c); }
  static LG_HD double sqrt(double a) { return ::sqrt(a); }
  static LG_HD double div(double a, double b) { return a / b; }
#endif
  static LG_HD double abs(double a) { return ::fabs(a); }
  static LG_HD double max_value() { return 1.7976931348623157e+308; }
  static constexpr int kBezIters = 56;
};

template <class T> struct V2 {
  T x, y;
};
template <class T> LG_HD T dot(V2<T> a, V2<T> b) { return Real<T>::fma(a.x, b.x, a.y * b.y); }
template <class T> LG_HD T cross(V2<T> a, V2<T> b) {
  return Real<T>::fma(a.x, b.y, -(a.y * b.x));
}
template <class T> LG_HD V2<T> unit(V2<T> v) {
  T len = Real<T>::sqrt(dot(v, v));
  return {Real<T>::div(v.x, len), Real<T>::div(v.y, len)};
}
template <class T> LG_HD V2<T> ray_at(V2<T> o, T t, V2<T> d) {
  return {Real<T>::fma(t, d.x, o.x), Real<T>::fma(t, d.y, o.y)};
}

// One lowered token of an object's postfix program (lg_scene.h builds them).
template <class T> struct alignas(16) Tok { // 48 / 80 bytes: header and parameters load as 16-byte vectors
  int32_t kind, op, a_start, b_start;
  T p[8];
  // CIRCLE : cx cy r r2          SEGMENT: ax ay ex ey
  // RECT   : cx cy ux uy vx vy uu vv     BEZIER : x0 y0 .. x3 y3
  // ELLIPSE: cx cy ux uy a b 1/a 1/b  (u = unit x axis of the ellipse in world space)
  // POLY   : op = vertex count k; the next ceil(k / 4) tokens are POINTS: x0 y0 .. x3 y3 (world space, hull order)
};

// A candidate hit: ray parameter, point, and what is needed to rebuild the
// normal later (aux = rect edge id or Bézier curve parameter).
template <class T> struct Cand {
  T t;
  V2<T> p;
  T aux;
};
template <class T> struct CandList {
  Cand<T> h[4];
  int n;
};

// ---- ORACLE.md §3.1 circle -------------------------------------------------
// The *_each forms hand every hit to a callable in the library's point order instead of appending it to a CandList:
// a list indexed by a running count lives in local memory (STL / LDL in the trace kernel's narrow phase, ncu r02), a
// callable keeps the hit in registers.  Same arithmetic, same order.
template <class T, class F> LG_HD void hit_circle_each(const T *c, V2<T> o, V2<T> d, F &&emit) {
  V2<T> m{c[0] - o.x, c[1] - o.y};
  T cr = cross(m, d);
  T disc = Real<T>::fma(-cr, cr, c[3]);
  if (!(disc >= (T)0)) return;
  T tca = dot(m, d);
  T thc = Real<T>::sqrt(disc);
  T t0 = tca - thc, t1 = tca + thc;
  if (t0 > (T)kTMin) emit(Cand<T>{t0, ray_at(o, t0, d), (T)0});
  if (t1 > (T)kTMin) emit(Cand<T>{t1, ray_at(o, t1, d), (T)1});
}
template <class T> LG_HD void hit_circle(const T *c, V2<T> o, V2<T> d, CandList<T> &out) {
  hit_circle_each(c, o, d, [&](const Cand<T> &h) { out.h[out.n++] = h; });
}

// ---- ORACLE.md §3.2 segment a + u e ---------------------------------------
template <class T>
LG_HD bool hit_edge(V2<T> a, V2<T> e, V2<T> o, V2<T> d, T aux, Cand<T> &h) {
  T denom = cross(d, e);
  T ad = Real<T>::abs(denom);
  if (!(ad > (T)kParEps)) return false;
  V2<T> w{a.x - o.x, a.y - o.y};
  T s = cross(w, d);
  if (denom < (T)0) s = -s;
  if (!(s >= (T)0) || !(s <= ad)) return false;
  T t = Real<T>::div(cross(w, e), denom);
  if (!(t > (T)kTMin)) return false;
  h.t = t;
  h.p = ray_at(o, t, d);
  h.aux = aux;
  return true;
}
template <class T, class F> LG_HD void hit_segment_each(const T *s, V2<T> o, V2<T> d, F &&emit) {
  Cand<T> h;
  if (hit_edge(V2<T>{s[0], s[1]}, V2<T>{s[2], s[3]}, o, d, (T)0, h)) emit(h);
}
template <class T> LG_HD void hit_segment(const T *s, V2<T> o, V2<T> d, CandList<T> &out) {
  hit_segment_each(s, o, d, [&](const Cand<T> &h) { out.h[out.n++] = h; });
}

// ---- ORACLE.md §3.3 rect (centre, half axes u, v) ---------------------------
// edge order = Rect::line_segments(): [right, bottom, left, top] (grid.rs:31)
template <class T> LG_HD bool rect_sat_reject(const T *r, V2<T> o, V2<T> d) {
  V2<T> m{r[0] - o.x, r[1] - o.y};
  T s = cross(d, m);
  T ext = Real<T>::abs(cross(d, V2<T>{r[2], r[3]})) + Real<T>::abs(cross(d, V2<T>{r[4], r[5]}));
  return !(Real<T>::abs(s) <= ext);
}
template <class T> LG_HD void rect_edge(const T *r, int k, V2<T> &a, V2<T> &e) {
  V2<T> c{r[0], r[1]}, u{r[2], r[3]}, v{r[4], r[5]};
  V2<T> u2{u.x + u.x, u.y + u.y}, v2{v.x + v.x, v.y + v.y};
  if (k == 0) {
    a = {c.x + u.x - v.x, c.y + u.y - v.y};
    e = v2;
  } else if (k == 1) {
    a = {c.x - u.x - v.x, c.y - u.y - v.y};
    e = u2;
  } else if (k == 2) {
    a = {c.x - u.x - v.x, c.y - u.y - v.y};
    e = v2;
  } else {
    a = {c.x - u.x + v.x, c.y - u.y + v.y};
    e = u2;
  }
}
template <class T, class F> LG_HD void hit_rect_each(const T *r, V2<T> o, V2<T> d, F &&emit) {
  if (rect_sat_reject(r, o, d)) return;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    V2<T> a, e;
    rect_edge(r, k, a, e);
    Cand<T> h;
    if (hit_edge(a, e, o, d, (T)k, h)) emit(h);
  }
}
template <class T> LG_HD void hit_rect(const T *r, V2<T> o, V2<T> d, CandList<T> &out) {
  hit_rect_each(r, o, d, [&](const Cand<T> &h) { out.h[out.n++] = h; });
}

// ---- ORACLE.md §3.4 cubic Bézier --------------------------------------------
template <class T> LG_HD T cubic(T c3, T c2, T c1, T c0, T t) {
  return Real<T>::fma(Real<T>::fma(Real<T>::fma(c3, t, c2), t, c1), t, c0);
}
template <class T> LG_HD void hit_bezier(const T *b, V2<T> o, V2<T> d, CandList<T> &out) {
  T y[4], x[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    V2<T> q{b[2 * i] - o.x, b[2 * i + 1] - o.y};
    y[i] = cross(d, q);
    x[i] = dot(d, q);
  }
  bool allpos = y[0] > (T)0 && y[1] > (T)0 && y[2] > (T)0 && y[3] > (T)0;
  bool allneg = y[0] < (T)0 && y[1] < (T)0 && y[2] < (T)0 && y[3] < (T)0;
  if (allpos || allneg) return;
  const T tmin = (T)kTMin;
  if (!(x[0] > tmin) && !(x[1] > tmin) && !(x[2] > tmin) && !(x[3] > tmin)) return;
  T c0 = y[0];
  T c1 = (T)3 * (y[1] - y[0]);
  T c2 = (T)3 * ((y[0] - (y[1] + y[1])) + y[2]);
  T c3 = (y[3] - y[0]) + (T)3 * (y[1] - y[2]);
  T e0 = x[0];
  T e1 = (T)3 * (x[1] - x[0]);
  T e2 = (T)3 * ((x[0] - (x[1] + x[1])) + x[2]);
  T e3 = (x[3] - x[0]) + (T)3 * (x[1] - x[2]);
  T A = (T)3 * c3, B = c2 + c2, C = c1;
  T r1 = (T)-1, r2 = (T)-1;
  if (A != (T)0) {
    T D = Real<T>::fma(B, B, -((T)4 * A * C));
    if (D > (T)0) {
      T sq = Real<T>::sqrt(D);
      T q = (T)-0.5 * (B + (B < (T)0 ? -sq : sq));
      r1 = Real<T>::div(q, A);
      if (q != (T)0) r2 = Real<T>::div(C, q);
    }
  } else if (B != (T)0) {
    r1 = Real<T>::div(-C, B);
  }
  if (r1 > r2) {
    T tmp = r1;
    r1 = r2;
    r2 = tmp;
  }
  T split[4];
  int ns = 0;
  split[ns++] = (T)0;
  if (r1 > (T)0 && r1 < (T)1) split[ns++] = r1;
  if (r2 > (T)0 && r2 < (T)1 && r2 != r1) split[ns++] = r2;
  split[ns++] = (T)1;
  for (int k = 0; k + 1 < ns; ++k) {
    T lo = split[k], hi = split[k + 1];
    T flo = cubic(c3, c2, c1, c0, lo), fhi = cubic(c3, c2, c1, c0, hi);
    bool nlo = flo < (T)0, nhi = fhi < (T)0;
    if (nlo == nhi) continue;
    for (int it = 0; it < Real<T>::kBezIters; ++it) {
      T mid = (T)0.5 * (lo + hi);
      T fm = cubic(c3, c2, c1, c0, mid);
      if ((fm < (T)0) == nlo)
        lo = mid;
      else
        hi = mid;
    }
    T tt = (T)0.5 * (lo + hi);
    T s = cubic(e3, e2, e1, e0, tt);
    if (!(s > tmin)) continue;
    out.h[out.n++] = {s, ray_at(o, s, d), tt};
  }
}
template <class T> LG_HD V2<T> bezier_normal(const T *b, T tt) {
  T om = (T)1 - tt;
  T w0 = om * om, w1 = (om + om) * tt, w2 = tt * tt;
  V2<T> d0{b[2] - b[0], b[3] - b[1]}, d1{b[4] - b[2], b[5] - b[3]}, d2{b[6] - b[4], b[7] - b[5]};
  V2<T> tg{Real<T>::fma(w0, d0.x, Real<T>::fma(w1, d1.x, w2 * d2.x)),
           Real<T>::fma(w0, d0.y, Real<T>::fma(w1, d1.y, w2 * d2.y))};
  return unit(V2<T>{-tg.y, tg.x});
}

// ---- ORACLE.md §3.7 ellipse: the ray in the frame where the ellipse is the unit circle ------------
template <class T> LG_HD V2<T> ellipse_frame(const T *e, V2<T> w) {
  V2<T> u{e[2], e[3]}, up{-e[3], e[2]};
  return {dot(w, u) * e[6], dot(w, up) * e[7]};
}
template <class T> LG_HD V2<T> ellipse_normal(const T *e, V2<T> p) {
  V2<T> u{e[2], e[3]}, up{-e[3], e[2]};
  V2<T> l = ellipse_frame(e, V2<T>{p.x - e[0], p.y - e[1]});
  T gx = l.x * e[6], gy = l.y * e[7];
  return unit(V2<T>{Real<T>::fma(gx, u.x, gy * up.x), Real<T>::fma(gx, u.y, gy * up.y)});
}
template <class T> LG_HD void hit_ellipse(const T *e, V2<T> o, V2<T> d, CandList<T> &out) {
  V2<T> lo = ellipse_frame(e, V2<T>{o.x - e[0], o.y - e[1]});
  V2<T> ld = ellipse_frame(e, d);
  T A = dot(ld, ld), B = dot(lo, ld), C = dot(lo, lo) - (T)1;
  T disc = Real<T>::fma(B, B, -(A * C));
  if (!(disc >= (T)0) || !(A > (T)0)) return;
  T sq = Real<T>::sqrt(disc);
  T t0 = Real<T>::div(-B - sq, A), t1 = Real<T>::div(-B + sq, A);
  if (t0 > (T)kTMin) out.h[out.n++] = {t0, ray_at(o, t0, d), (T)0};
  if (t1 > (T)kTMin) out.h[out.n++] = {t1, ray_at(o, t1, d), (T)1};
}

// ---- ORACLE.md §3.8 convex polygon: the edges v_i -> v_(i+1) as segments, in hull order ------------
// (the vertices live in the POINTS tokens that follow the POLY token)
template <class T> LG_HD V2<T> poly_vertex(const Tok<T> &k, int i) {
  const Tok<T> &q = (&k)[1 + (i >> 2)];
  return {q.p[2 * (i & 3)], q.p[2 * (i & 3) + 1]};
}
template <class T> LG_HD void hit_poly(const Tok<T> &k, V2<T> o, V2<T> d, CandList<T> &out) {
  const int n = k.op;
  V2<T> a = poly_vertex(k, 0);
  for (int i = 0; i < n; ++i) {
    const V2<T> b = poly_vertex(k, i + 1 < n ? i + 1 : 0);
    Cand<T> h;
    if (out.n < 4 && hit_edge(a, V2<T>{b.x - a.x, b.y - a.y}, o, d, (T)i, h)) out.h[out.n++] = h;
    a = b;
  }
}
template <class T> LG_HD V2<T> poly_normal(const Tok<T> &k, int i) {
  const int n = k.op;
  const V2<T> a = poly_vertex(k, i), b = poly_vertex(k, i + 1 < n ? i + 1 : 0);
  return unit(V2<T>{-(b.y - a.y), b.x - a.x});
}
template <class T> LG_HD bool poly_contains(const Tok<T> &k, V2<T> p) {
  const int n = k.op;
  int pos = 0, neg = 0;
  V2<T> a = poly_vertex(k, 0);
  for (int i = 0; i < n; ++i) {
    const V2<T> b = poly_vertex(k, i + 1 < n ? i + 1 : 0);
    const T c = cross(V2<T>{b.x - a.x, b.y - a.y}, V2<T>{p.x - a.x, p.y - a.y});
    pos += c > (T)0 ? 1 : 0, neg += c < (T)0 ? 1 : 0;
    a = b;
  }
  return pos == n || neg == n; // strictly on the same side of every edge, whichever way the hull is wound
}

// unit normal of the hit (token, point, aux); orientation is fixed later
template <class T> LG_HD V2<T> hit_normal(const Tok<T> &k, V2<T> p, T aux) {
  switch (k.kind) {
  case TOK_CIRCLE: return unit(V2<T>{p.x - k.p[0], p.y - k.p[1]});
  case TOK_SEGMENT: return unit(V2<T>{-k.p[3], k.p[2]});
  case TOK_RECT: {
    V2<T> a, e;
    rect_edge(k.p, (int)aux, a, e);
    return unit(V2<T>{-e.y, e.x});
  }
  case TOK_ELLIPSE: return ellipse_normal(k.p, p);
  case TOK_POLY: return poly_normal(k, (int)aux);
  default: return bezier_normal(k.p, aux);
  }
}

// ---- ORACLE.md §3.5 contains -------------------------------------------------
template <class T> LG_HD bool contains_leaf(const Tok<T> &l, V2<T> p) {
  if (l.kind == TOK_CIRCLE) {
    V2<T> q{p.x - l.p[0], p.y - l.p[1]};
    return dot(q, q) < l.p[3];
  }
  if (l.kind == TOK_RECT) {
    V2<T> q{p.x - l.p[0], p.y - l.p[1]};
    T a = dot(q, V2<T>{l.p[2], l.p[3]});
    T b = dot(q, V2<T>{l.p[4], l.p[5]});
    return Real<T>::abs(a) < l.p[6] && Real<T>::abs(b) < l.p[7];
  }
  if (l.kind == TOK_ELLIPSE) {
    V2<T> q = ellipse_frame(l.p, V2<T>{p.x - l.p[0], p.y - l.p[1]});
    return dot(q, q) < (T)1;
  }
  if (l.kind == TOK_POLY) return poly_contains(l, p);
  return false; // mirrors never contain: src/light_garden/object.rs:243-244
}
// postfix evaluation of tokens [s, e] with a bit stack
template <class T> LG_HD bool contains_range(const Tok<T> *tok, int s, int e, V2<T> p) {
  unsigned long long st = 0;
  for (int i = s; i <= e; ++i) {
    const Tok<T> &l = tok[i];
    if (l.kind == TOK_OP) {
      bool b = st & 1, a = (st >> 1) & 1;
      st >>= 2;
      bool r = l.op == OP_AND ? (a && b) : l.op == OP_OR ? (a || b) : (a && !b);
      st = (st << 1) | (r ? 1ull : 0ull);
    } else if (l.kind != TOK_POINTS) {
      st = (st << 1) | (contains_leaf(l, p) ? 1ull : 0ull);
    }
  }
  return st & 1;
}

// ---- ORACLE.md §4 reflect / refract ------------------------------------------
template <class T> LG_HD V2<T> face(V2<T> d, V2<T> n) {
  if (dot(d, n) > (T)0) return {-n.x, -n.y};
  return n;
}
template <class T> LG_HD V2<T> reflect_dir(V2<T> d, V2<T> n_in) {
  V2<T> n = face(d, n_in);
  T k = dot(d, n);
  T k2 = k + k;
  return unit(V2<T>{Real<T>::fma(-k2, n.x, d.x), Real<T>::fma(-k2, n.y, d.y)});
}
template <class T>
LG_HD T refract_dir(V2<T> d, V2<T> n_in, T n1, T n2, V2<T> &refl, V2<T> &refr, bool &has) {
  V2<T> n = face(d, n_in);
  refl = reflect_dir(d, n_in);
  T eta = Real<T>::div(n1, n2);
  T cosi = -dot(d, n);
  T sin2t = (eta * eta) * Real<T>::fma(-cosi, cosi, (T)1);
  if (sin2t > (T)1) {
    has = false;
    return (T)1;
  }
  T cost = Real<T>::sqrt((T)1 - sin2t);
  T k = Real<T>::fma(eta, cosi, -cost);
  refr = unit(V2<T>{Real<T>::fma(eta, d.x, k * n.x), Real<T>::fma(eta, d.y, k * n.y)});
  has = true;
  T a = n1 * cosi, b = n2 * cost, c = n1 * cost, e = n2 * cosi;
  T ds = a + b, dp = c + e;
  if (ds == (T)0 || dp == (T)0) return (T)1;
  T rs = Real<T>::div(a - b, ds), rp = Real<T>::div(c - e, dp);
  return (T)0.5 * Real<T>::fma(rs, rs, rp * rp);
}

} // namespace lg
