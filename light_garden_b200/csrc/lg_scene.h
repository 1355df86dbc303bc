// lg_scene.h — host-side lowering of the reference's scene model to the flat
// device tables the trace kernel reads.
//
// Input is what the Rust app holds in Tracer{objects, lights, ...}
// (src/light_garden/tracer.rs:4-17) flattened to the PODs of
// include/light_garden_b200.h.  Every Geo tree (object.rs:282-295; Lens =
// Logic(And, circle, circle), object.rs:393-410) becomes a postfix program of
// world-space leaves: the Logic nodes' local frames (origin + Rotation2,
// default.ron:20-29) are composed in f64 here, once, instead of per ray.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/light_garden_b200.h"

namespace lg {

struct HostTok {
  int32_t kind, op, a_start, b_start;
  double p[8]; // world-space: CIRCLE cx cy r | RECT cx cy ux uy vx vy | SEGMENT ax ay bx by | BEZIER 8 |
               // ELLIPSE (kind 5) cx cy ux uy a b | POLYGON (kind 6) op = vertex count, followed by
               // POINTS tokens (kind 7) with op = 1..4 world-space vertices x0 y0 .. x3 y3 each
};
struct HostObj {
  int32_t first, count;
  int32_t has_material;
  double n;
  double aabb[4]; // xmin ymin xmax ymax of the leaves that can contain a point
  bool can_contain;
};

struct HostScene {
  std::vector<HostTok> toks;
  std::vector<HostObj> objs;
  // per object: indices (ascending) of the OTHER material objects whose
  // containing region may overlap it — the only candidates of the
  // "which medium does the ray leave into" scan of tracer.rs:432-439
  std::vector<int32_t> ovl_start, ovl_list;
  LgTraceParams params{};
  double bound = 1.0; // max |coordinate| of anything in the scene or canvas
};

struct Affine {
  double m11, m21, m12, m22, tx, ty; // column-major 2x2 like nalgebra + translation
};
inline Affine compose_rot(const Affine &P, const double rot[4]) {
  Affine r = P;
  r.m11 = P.m11 * rot[0] + P.m12 * rot[1];
  r.m21 = P.m21 * rot[0] + P.m22 * rot[1];
  r.m12 = P.m11 * rot[2] + P.m12 * rot[3];
  r.m22 = P.m21 * rot[2] + P.m22 * rot[3];
  return r;
}
inline void xform(const Affine &A, double x, double y, double *o) {
  o[0] = A.m11 * x + A.m12 * y + A.tx;
  o[1] = A.m21 * x + A.m22 * y + A.ty;
}

// What a node may hold.  The exact tests follow the oracle for any IEEE input, but the bounding circles and the grid
// that decide WHICH objects get an exact test are built from these numbers: with a non-finite or absurdly large one, a
// negative extent or a "rotation" that scales, they would prune hits the reference finds
// (tests/test_gpu_fuzz.py::test_hostile_object_parameters...).  Such a node is refused, never traced differently.
constexpr double kMaxCoordinate = 1e12;
inline bool node_ok(const LgGeoNode &g, std::string &err) {
  int np = 0;
  bool rot = false, extents = false;
  switch (g.kind) {
  case LG_GEO_CIRCLE: np = 3; break;
  case LG_GEO_RECT: np = 4, rot = true, extents = true; break;
  case LG_GEO_SEGMENT: np = 4; break;
  case LG_GEO_BEZIER: np = 8; break;
  case LG_GEO_ELLIPSE: np = 4, rot = true; break;
  case LG_GEO_LOGIC: np = 2, rot = true; break;
  case LG_GEO_POLYGON: np = 2, rot = true; break;
  case LG_GEO_POINTS: np = (g.op >= 1 && g.op <= 4) ? 2 * g.op : 0; break;
  default: return true; // refused by kind further down
  }
  for (int k = 0; k < np; ++k)
    if (!(std::fabs(g.p[k]) <= kMaxCoordinate)) {
      err = "geometry parameter is not finite or beyond 1e12";
      return false;
    }
  if (g.kind == LG_GEO_CIRCLE && !(g.p[2] >= 0.0)) {
    err = "negative radius";
    return false;
  }
  if (extents && (!(g.p[2] >= 0.0) || !(g.p[3] >= 0.0))) {
    err = "negative width or height";
    return false;
  }
  if (rot) {
    const double *r = g.rot;
    const double tol = 1e-6;
    if (!(std::fabs(r[0] * r[0] + r[1] * r[1] - 1.0) <= tol) || !(std::fabs(r[2] * r[2] + r[3] * r[3] - 1.0) <= tol) ||
        !(std::fabs(r[0] * r[2] + r[1] * r[3]) <= tol)) {
      err = "rotation matrix is not orthonormal";
      return false;
    }
  }
  return true;
}

inline bool lower_geo(const LgGeoNode *nodes, uint32_t n_nodes, int32_t ix, const Affine &A,
                      std::vector<HostTok> &out, size_t base, int depth, std::string &err) {
  if (ix < 0 || (uint32_t)ix >= n_nodes) {
    err = "geometry node index out of range";
    return false;
  }
  if (depth > 32) {
    err = "geometry tree deeper than 32 (cycle?)";
    return false;
  }
  if (out.size() - base > (size_t)1 << 16) { // nodes shared between branches unfold into a tree: 2^depth tokens
    err = "geometry tree of more than 65536 tokens (shared nodes?)";
    return false;
  }
  const LgGeoNode &g = nodes[ix];
  if (!node_ok(g, err)) return false;
  HostTok t{};
  t.a_start = t.b_start = -1;
  switch (g.kind) {
  case LG_GEO_CIRCLE:
    t.kind = 0;
    xform(A, g.p[0], g.p[1], t.p);
    t.p[2] = g.p[2];
    out.push_back(t);
    return true;
  case LG_GEO_RECT: {
    t.kind = 1;
    xform(A, g.p[0], g.p[1], t.p);
    Affine W = compose_rot(A, g.rot);
    double hw = g.p[2] * 0.5, hh = g.p[3] * 0.5;
    t.p[2] = W.m11 * hw;
    t.p[3] = W.m21 * hw;
    t.p[4] = W.m12 * hh;
    t.p[5] = W.m22 * hh;
    out.push_back(t);
    return true;
  }
  case LG_GEO_SEGMENT:
    t.kind = 2;
    xform(A, g.p[0], g.p[1], t.p);
    xform(A, g.p[2], g.p[3], t.p + 2);
    out.push_back(t);
    return true;
  case LG_GEO_BEZIER:
    t.kind = 3;
    for (int k = 0; k < 4; ++k) xform(A, g.p[2 * k], g.p[2 * k + 1], t.p + 2 * k);
    out.push_back(t);
    return true;
  case LG_GEO_ELLIPSE: { // object.rs:38-45: Ellipse{origin, a, b, rot}
    t.kind = 5;
    xform(A, g.p[0], g.p[1], t.p);
    Affine W = compose_rot(A, g.rot);
    t.p[2] = W.m11;
    t.p[3] = W.m21;
    t.p[4] = g.p[2];
    t.p[5] = g.p[3];
    if (!(g.p[2] > 0.0) || !(g.p[3] > 0.0)) {
      err = "ellipse semi axes must be > 0";
      return false;
    }
    out.push_back(t);
    return true;
  }
  case LG_GEO_LOGIC: {
    if (g.op < LG_OP_AND || g.op > LG_OP_ANDNOT) {
      err = "unknown LogicOp";
      return false;
    }
    Affine W = compose_rot(A, g.rot);
    double tw[2];
    xform(A, g.p[0], g.p[1], tw);
    W.tx = tw[0];
    W.ty = tw[1];
    int32_t a0 = (int32_t)(out.size() - base);
    if (!lower_geo(nodes, n_nodes, g.child_a, W, out, base, depth + 1, err)) return false;
    int32_t b0 = (int32_t)(out.size() - base);
    if (!lower_geo(nodes, n_nodes, g.child_b, W, out, base, depth + 1, err)) return false;
    t.kind = 4;
    t.op = g.op;
    t.a_start = a0;
    t.b_start = b0;
    out.push_back(t);
    return true;
  }
  case LG_GEO_POLYGON: { // object.rs:34-36: ConvexPolygon::new_convex_hull(points); vertices arrive in hull order
    const int k = g.op;
    if (k < 3 || k > LG_POLYGON_MAX_VERTICES) {
      err = "convex polygon needs 3..32 vertices";
      return false;
    }
    const Affine W = [&] {
      Affine w = compose_rot(A, g.rot);
      double tw[2];
      xform(A, g.p[0], g.p[1], tw);
      w.tx = tw[0], w.ty = tw[1];
      return w;
    }();
    t.kind = 6;
    t.op = k;
    out.push_back(t);
    // kind 7 tokens carry the world-space vertices, four each
    int got = 0;
    int32_t pn = g.child_a;
    HostTok dt{};
    dt.kind = 7, dt.a_start = dt.b_start = -1;
    for (int guard = 0; got < k && guard < LG_POLYGON_MAX_VERTICES; ++guard) {
      if (pn < 0 || (uint32_t)pn >= n_nodes || nodes[pn].kind != LG_GEO_POINTS || nodes[pn].op < 1 || nodes[pn].op > 4) {
        err = "convex polygon: bad vertex list";
        return false;
      }
      if (!node_ok(nodes[pn], err)) return false;
      for (int q = 0; q < nodes[pn].op && got < k; ++q, ++got) {
        xform(W, nodes[pn].p[2 * q], nodes[pn].p[2 * q + 1], dt.p + 2 * (got & 3));
        if ((got & 3) == 3 || got == k - 1) {
          dt.op = (got & 3) + 1;
          out.push_back(dt);
          dt = HostTok{};
          dt.kind = 7, dt.a_start = dt.b_start = -1;
        }
      }
      pn = nodes[pn].child_a;
    }
    if (got != k) {
      err = "convex polygon: vertex list shorter than its count";
      return false;
    }
    return true;
  }
  default:
    // MCircle exists in collision2d's Geo (drawer.rs:57-78) but no Object constructor makes one: SURVEY.md §8f.
    err = "unsupported Geo kind";
    return false;
  }
}

constexpr int kMaxTokensPerObject = 64;

inline int32_t lower_scene(const LgObject *objects, uint32_t n_obj, const LgGeoNode *nodes, uint32_t n_nodes,
                           const LgTraceParams &prm, HostScene &hs, std::string &err) {
  hs = HostScene{};
  hs.params = prm;
  const Affine I{1, 0, 0, 1, 0, 0};
  double bound = 0;
  for (int k = 0; k < 4; ++k) {
    if (!(std::fabs(prm.canvas_tlbr[k]) <= kMaxCoordinate)) { // the grid spans the canvas: see node_ok
      err = "canvas bounds are not finite or beyond 1e12";
      return LG_ERR_INVALID;
    }
    bound = std::fmax(bound, std::fabs(prm.canvas_tlbr[k]));
  }
  for (uint32_t i = 0; i < n_obj; ++i) {
    HostObj o{};
    o.first = (int32_t)hs.toks.size();
    if (!lower_geo(nodes, n_nodes, objects[i].root, I, hs.toks, (size_t)o.first, 0, err)) return LG_ERR_INVALID;
    o.count = (int32_t)hs.toks.size() - o.first;
    if (o.count > kMaxTokensPerObject) {
      err = "geometry tree with more than 64 tokens";
      return LG_ERR_UNSUPPORTED;
    }
    o.has_material = objects[i].has_material != 0;
    o.n = objects[i].refractive_index;
    if (o.has_material && !(o.n > 0.0)) {
      // the GUI slider reaches <= 0 (gui/mod.rs:287-294); not a physical medium
      err = "refractive index must be > 0";
      return LG_ERR_UNSUPPORTED;
    }
    o.aabb[0] = o.aabb[1] = 1e300;
    o.aabb[2] = o.aabb[3] = -1e300;
    o.can_contain = false;
    for (int k = 0; k < o.count; ++k) {
      const HostTok &t = hs.toks[o.first + k];
      int np = t.kind == 0 ? 1 : t.kind == 1 ? 1 : t.kind == 2 ? 2 : t.kind == 3 ? 4 : t.kind == 5 ? 1 : t.kind == 7 ? t.op : 0;
      for (int q = 0; q < np; ++q) bound = std::fmax(bound, std::fmax(std::fabs(t.p[2 * q]), std::fabs(t.p[2 * q + 1])));
      if (t.kind == 7) { // vertices of a convex polygon: it contains points
        o.can_contain = true;
        for (int q = 0; q < np; ++q) {
          o.aabb[0] = std::fmin(o.aabb[0], t.p[2 * q]), o.aabb[2] = std::fmax(o.aabb[2], t.p[2 * q]);
          o.aabb[1] = std::fmin(o.aabb[1], t.p[2 * q + 1]), o.aabb[3] = std::fmax(o.aabb[3], t.p[2 * q + 1]);
        }
        continue;
      }
      double ex = 0, ey = 0;
      if (t.kind == 0) {
        ex = ey = std::fabs(t.p[2]);
      } else if (t.kind == 1) {
        ex = std::fabs(t.p[2]) + std::fabs(t.p[4]);
        ey = std::fabs(t.p[3]) + std::fabs(t.p[5]);
      } else if (t.kind == 5) {
        ex = ey = std::fmax(t.p[4], t.p[5]);
      } else {
        continue;
      }
      bound = std::fmax(bound, std::fmax(std::fabs(t.p[0]) + ex, std::fabs(t.p[1]) + ey));
      o.can_contain = true;
      o.aabb[0] = std::fmin(o.aabb[0], t.p[0] - ex);
      o.aabb[1] = std::fmin(o.aabb[1], t.p[1] - ey);
      o.aabb[2] = std::fmax(o.aabb[2], t.p[0] + ex);
      o.aabb[3] = std::fmax(o.aabb[3], t.p[1] + ey);
    }
    hs.objs.push_back(o);
  }
  hs.bound = bound > 0 ? bound : 1.0;
  // overlap candidates: conservative (AABBs inflated by 1e-4 of the scene bound,
  // far above any rounding of a hit point) so the scan result is unchanged.
  // Sweep and prune along x instead of all pairs: O(n log n + overlaps).
  const double pad = 1e-4 * hs.bound;
  struct Box {
    double x0, y0, x1, y1;
  };
  std::vector<Box> hull(n_obj); // everything a hit point on object i can lie on: all leaves
  for (uint32_t i = 0; i < n_obj; ++i) {
    const HostObj &a = hs.objs[i];
    Box bx{1e300, 1e300, -1e300, -1e300};
    for (int k = 0; k < a.count; ++k) {
      const HostTok &t = hs.toks[a.first + k];
      if (t.kind == 4 || t.kind == 6) continue;
      if (t.kind == 0 || t.kind == 5) {
        const double rr = t.kind == 0 ? t.p[2] : std::fmax(t.p[4], t.p[5]);
        bx.x0 = std::fmin(bx.x0, t.p[0] - rr), bx.x1 = std::fmax(bx.x1, t.p[0] + rr);
        bx.y0 = std::fmin(bx.y0, t.p[1] - rr), bx.y1 = std::fmax(bx.y1, t.p[1] + rr);
      } else if (t.kind == 1) {
        double ex = std::fabs(t.p[2]) + std::fabs(t.p[4]), ey = std::fabs(t.p[3]) + std::fabs(t.p[5]);
        bx.x0 = std::fmin(bx.x0, t.p[0] - ex), bx.x1 = std::fmax(bx.x1, t.p[0] + ex);
        bx.y0 = std::fmin(bx.y0, t.p[1] - ey), bx.y1 = std::fmax(bx.y1, t.p[1] + ey);
      } else {
        int np = t.kind == 2 ? 2 : t.kind == 7 ? t.op : 4;
        for (int q = 0; q < np; ++q) {
          bx.x0 = std::fmin(bx.x0, t.p[2 * q]), bx.x1 = std::fmax(bx.x1, t.p[2 * q]);
          bx.y0 = std::fmin(bx.y0, t.p[2 * q + 1]), bx.y1 = std::fmax(bx.y1, t.p[2 * q + 1]);
        }
      }
    }
    hull[i] = bx;
  }
  // i lists j  <=>  i has a material (the scan only runs for material hits, tracer.rs:428), j has a
  // material and can contain a point, and hull(i) meets j's padded containing box
  std::vector<uint32_t> order;
  for (uint32_t i = 0; i < n_obj; ++i)
    if (hs.objs[i].has_material) order.push_back(i);
  std::sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return hull[a].x0 < hull[b].x0; });
  std::vector<std::vector<int32_t>> lists(n_obj);
  auto lists_j = [&](uint32_t i, uint32_t j) {
    const HostObj &b = hs.objs[j];
    if (!b.can_contain) return false;
    const Box &h = hull[i];
    return !(b.aabb[0] - pad > h.x1 || b.aabb[2] + pad < h.x0 || b.aabb[1] - pad > h.y1 || b.aabb[3] + pad < h.y0);
  };
  for (size_t a = 0; a < order.size(); ++a) {
    const uint32_t i = order[a];
    for (size_t b = a + 1; b < order.size(); ++b) {
      const uint32_t j = order[b];
      if (hull[j].x0 - pad > hull[i].x1 + pad) break; // hull(j) contains j's containing box: nothing further overlaps
      if (lists_j(i, j)) lists[i].push_back((int32_t)j);
      if (lists_j(j, i)) lists[j].push_back((int32_t)i);
    }
  }
  hs.ovl_start.assign(n_obj + 1, 0);
  for (uint32_t i = 0; i < n_obj; ++i) {
    hs.ovl_start[i] = (int32_t)hs.ovl_list.size();
    std::sort(lists[i].begin(), lists[i].end());
    hs.ovl_list.insert(hs.ovl_list.end(), lists[i].begin(), lists[i].end());
  }
  hs.ovl_start[n_obj] = (int32_t)hs.ovl_list.size();
  return LG_OK;
}

// postfix containment in f64 on the host: start medium of a light,
// src/light_garden/tracer.rs:280-287 (`last match wins`, no break).
inline bool host_contains_leaf(const HostTok &t, double x, double y) {
  if (t.kind == 0) {
    double qx = x - t.p[0], qy = y - t.p[1];
    return std::fma(qx, qx, qy * qy) < t.p[2] * t.p[2];
  }
  if (t.kind == 1) {
    double qx = x - t.p[0], qy = y - t.p[1];
    double a = std::fma(qx, t.p[2], qy * t.p[3]), b = std::fma(qx, t.p[4], qy * t.p[5]);
    double uu = std::fma(t.p[2], t.p[2], t.p[3] * t.p[3]), vv = std::fma(t.p[4], t.p[4], t.p[5] * t.p[5]);
    return std::fabs(a) < uu && std::fabs(b) < vv;
  }
  if (t.kind == 5) {
    double qx = x - t.p[0], qy = y - t.p[1];
    double lx = std::fma(qx, t.p[2], qy * t.p[3]) * (1.0 / t.p[4]);
    double ly = std::fma(qx, -t.p[3], qy * t.p[2]) * (1.0 / t.p[5]);
    return std::fma(lx, lx, ly * ly) < 1.0;
  }
  if (t.kind == 6) { // convex polygon: strictly on the same side of every edge (ORACLE.md §3.8)
    const int k = t.op;
    int pos = 0, neg = 0;
    for (int i = 0; i < k; ++i) {
      const int j = i + 1 < k ? i + 1 : 0;
      const double ax = (&t)[1 + (i >> 2)].p[2 * (i & 3)], ay = (&t)[1 + (i >> 2)].p[2 * (i & 3) + 1];
      const double bx = (&t)[1 + (j >> 2)].p[2 * (j & 3)], by = (&t)[1 + (j >> 2)].p[2 * (j & 3) + 1];
      const double ex = bx - ax, ey = by - ay, c = std::fma(ex, y - ay, -(ey * (x - ax)));
      pos += c > 0.0, neg += c < 0.0;
    }
    return pos == k || neg == k;
  }
  return false;
}
inline bool host_contains(const HostScene &hs, int obj, double x, double y) {
  const HostObj &o = hs.objs[obj];
  uint64_t st = 0;
  for (int i = 0; i < o.count; ++i) {
    const HostTok &t = hs.toks[o.first + i];
    if (t.kind == 4) {
      bool b = st & 1, a = (st >> 1) & 1;
      st >>= 2;
      bool r = t.op == LG_OP_AND ? (a && b) : t.op == LG_OP_OR ? (a || b) : (a && !b);
      st = (st << 1) | (r ? 1 : 0);
    } else if (t.kind != 7) {
      st = (st << 1) | (host_contains_leaf(t, x, y) ? 1 : 0);
    }
  }
  return st & 1;
}
inline double host_start_medium(const HostScene &hs, double x, double y) {
  double n = 1.0;
  for (size_t i = 0; i < hs.objs.size(); ++i)
    if (hs.objs[i].has_material && host_contains(hs, (int)i, x, y)) n = hs.objs[i].n;
  return n;
}

// ---- uniform grid over the objects' bounding circles (stands in for tile_map.rs) ----------
// Cell (ix, iy) is [x0 + ix cs, x0 + (ix + 1) cs] x [y0 + iy cs, y0 + (iy + 1) cs]; it lists every object whose
// circle (cx, cy, r) comes within `reach` of it.  x0, y0, cs, x1, y1 are representable in the device precision, so
// the device reproduces the cell boundaries with one fma.
struct HostGrid {
  double x0 = 0, y0 = 0, x1 = 0, y1 = 0, cs = 1;
  int nx = 1, ny = 1;
  std::vector<unsigned> start, obj; // CSR
};
inline double grid_round(double v, bool f32) { return f32 ? (double)(float)v : v; }
// circ = n x (cx, cy, r); `reach` is added to every radius; `pad` (> reach) is the empty rim around the box;
// `density` = target number of cells per object
inline HostGrid build_grid(const double *circ, size_t n, double reach, double pad, bool f32, double density = 1.0) {
  HostGrid g;
  if (n == 0) {
    g.start.assign(2, 0u);
    g.x1 = g.y1 = 1;
    return g;
  }
  double lox = 1e300, loy = 1e300, hix = -1e300, hiy = -1e300;
  for (size_t i = 0; i < n; ++i) {
    const double r = circ[3 * i + 2] + reach;
    lox = std::min(lox, circ[3 * i] - r), hix = std::max(hix, circ[3 * i] + r);
    loy = std::min(loy, circ[3 * i + 1] - r), hiy = std::max(hiy, circ[3 * i + 1] + r);
  }
  const double w = hix - lox, h = hiy - loy;
  double cs = std::sqrt(std::max(w * h, 1e-300) / (std::max(density, 1e-3) * (double)n));
  cs = std::max(cs, std::max(w, h) / 1024.0); // at most 1024 cells a side
  if (!(cs > 0)) cs = 1.0;
  for (int attempt = 0;; ++attempt) {
    pad = std::max(pad, 1e-3 * cs);
    g.cs = grid_round(cs, f32);
    g.x0 = grid_round(lox - 2 * pad, f32), g.y0 = grid_round(loy - 2 * pad, f32);
    g.nx = std::max(1, (int)std::ceil((hix + 2 * pad - g.x0) / g.cs));
    g.ny = std::max(1, (int)std::ceil((hiy + 2 * pad - g.y0) / g.cs));
    g.x1 = grid_round(std::fma((double)g.nx, g.cs, g.x0), f32), g.y1 = grid_round(std::fma((double)g.ny, g.cs, g.y0), f32);
    const size_t cells = (size_t)g.nx * g.ny;
    std::vector<unsigned> count(cells + 1, 0u);
    auto for_cells = [&](size_t i, auto f) {
      const double cx = circ[3 * i], cy = circ[3 * i + 1], r = circ[3 * i + 2] + reach;
      const int ia = std::max(0, (int)std::floor((cx - r - g.x0) / g.cs) - 1), ib = std::min(g.nx - 1, (int)std::floor((cx + r - g.x0) / g.cs) + 1);
      const int ja = std::max(0, (int)std::floor((cy - r - g.y0) / g.cs) - 1), jb = std::min(g.ny - 1, (int)std::floor((cy + r - g.y0) / g.cs) + 1);
      for (int j = ja; j <= jb; ++j)
        for (int i2 = ia; i2 <= ib; ++i2) {
          const double bx0 = g.x0 + i2 * g.cs, bx1 = g.x0 + (i2 + 1) * g.cs, by0 = g.y0 + j * g.cs, by1 = g.y0 + (j + 1) * g.cs;
          const double dx = std::max(std::max(bx0 - cx, cx - bx1), 0.0), dy = std::max(std::max(by0 - cy, cy - by1), 0.0);
          if (dx * dx + dy * dy <= r * r) f((size_t)j * g.nx + i2);
        }
    };
    for (size_t i = 0; i < n; ++i) for_cells(i, [&](size_t c) { ++count[c + 1]; });
    size_t total = 0;
    for (size_t c = 1; c <= cells; ++c) total += count[c];
    if (total > 32 * n + 4096 && attempt < 12) { // large objects in a fine grid: coarsen
      cs *= 2;
      continue;
    }
    for (size_t c = 1; c <= cells; ++c) count[c] += count[c - 1];
    g.start = count;
    g.obj.assign(total, 0u);
    std::vector<unsigned> cur(count.begin(), count.end() - 1);
    for (size_t i = 0; i < n; ++i) for_cells(i, [&](size_t c) { g.obj[cur[c]++] = (unsigned)i; }); // ascending object index
    return g;
  }
}

} // namespace lg
