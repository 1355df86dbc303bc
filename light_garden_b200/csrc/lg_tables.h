// lg_tables.h — host-side construction of the device tables from the lowered scene (lg_scene.h): tokens in the
// device precision, one bounding circle per object, the broad-phase table and the uniform grid.  No CUDA calls:
// lg_capi.cu uploads the results, tests/host_geom_check.cu runs the nearest-hit search on them on the CPU.
#pragma once
#include <cmath>
#include <limits>
#include <vector>

#include "lg_geom.cuh"
#include "lg_scene.h"

namespace lg {

// tokens in precision T (ORACLE.md §2.2: cast to T, then derive in T)
template <class T> std::vector<Tok<T>> device_tokens(const HostScene &hs) {
  std::vector<Tok<T>> toks(hs.toks.size());
  for (size_t i = 0; i < hs.toks.size(); ++i) {
    const HostTok &h = hs.toks[i];
    Tok<T> t{};
    t.kind = h.kind, t.op = h.op, t.a_start = h.a_start, t.b_start = h.b_start;
    switch (h.kind) {
    case 0:
      t.p[0] = (T)h.p[0], t.p[1] = (T)h.p[1], t.p[2] = (T)h.p[2];
      t.p[3] = t.p[2] * t.p[2];
      break;
    case 1:
      for (int k = 0; k < 6; ++k) t.p[k] = (T)h.p[k];
      t.p[6] = std::fma(t.p[2], t.p[2], t.p[3] * t.p[3]);
      t.p[7] = std::fma(t.p[4], t.p[4], t.p[5] * t.p[5]);
      break;
    case 2:
      t.p[0] = (T)h.p[0], t.p[1] = (T)h.p[1];
      t.p[2] = (T)h.p[2] - t.p[0];
      t.p[3] = (T)h.p[3] - t.p[1];
      break;
    case 3:
      for (int k = 0; k < 8; ++k) t.p[k] = (T)h.p[k];
      break;
    case 5:
      for (int k = 0; k < 6; ++k) t.p[k] = (T)h.p[k];
      t.p[6] = (T)1 / t.p[4];
      t.p[7] = (T)1 / t.p[5];
      break;
    case 7: // polygon vertices
      for (int k = 0; k < 8; ++k) t.p[k] = (T)h.p[k];
      break;
    default: break;
    }
    toks[i] = t;
  }
  return toks;
}

// One bounding circle (cx, cy, r) per object, in f64: every hit point of an object lies on one of its leaves, so the
// circle around all the leaves' bounding circles bounds every hit.
inline std::vector<double> object_circles(const HostScene &hs) {
  std::vector<double> out(3 * hs.objs.size(), 0.0);
  auto leaf_circle = [](const HostTok &t, double &cx, double &cy, double &rr) {
    if (t.kind == 0) {
      cx = t.p[0], cy = t.p[1], rr = std::fabs(t.p[2]);
    } else if (t.kind == 5) {
      cx = t.p[0], cy = t.p[1], rr = std::fmax(t.p[4], t.p[5]);
    } else if (t.kind == 1) {
      cx = t.p[0], cy = t.p[1], rr = std::hypot(std::hypot(t.p[2], t.p[3]), std::hypot(t.p[4], t.p[5]));
    } else { // segment / Bezier: the control polygon; kind 7: up to four vertices of a convex polygon
      const int np = t.kind == 2 ? 2 : t.kind == 7 ? t.op : 4;
      cx = cy = 0;
      for (int q = 0; q < np; ++q) cx += t.p[2 * q] / np, cy += t.p[2 * q + 1] / np;
      rr = 0;
      for (int q = 0; q < np; ++q) rr = std::fmax(rr, std::hypot(t.p[2 * q] - cx, t.p[2 * q + 1] - cy));
    }
  };
  for (size_t i = 0; i < hs.objs.size(); ++i) {
    const HostObj &o = hs.objs[i];
    double x0 = 1e300, y0 = 1e300, x1 = -1e300, y1 = -1e300;
    for (int k = 0; k < o.count; ++k) {
      const HostTok &t = hs.toks[o.first + k];
      if (t.kind == 4 || t.kind == 6) continue;
      double cx, cy, rr;
      leaf_circle(t, cx, cy, rr);
      x0 = std::fmin(x0, cx - rr), x1 = std::fmax(x1, cx + rr), y0 = std::fmin(y0, cy - rr), y1 = std::fmax(y1, cy + rr);
    }
    const double mx = 0.5 * (x0 + x1), my = 0.5 * (y0 + y1);
    double rad = 0;
    for (int k = 0; k < o.count; ++k) {
      const HostTok &t = hs.toks[o.first + k];
      if (t.kind == 4 || t.kind == 6) continue;
      double cx, cy, rr;
      leaf_circle(t, cx, cy, rr);
      rad = std::fmax(rad, std::hypot(cx - mx, cy - my) + rr);
    }
    out[3 * i] = mx, out[3 * i + 1] = my, out[3 * i + 2] = rad;
  }
  return out;
}

// Broad-phase table for coordinate bound B (scene, canvas, lights, explicit ray origins): SoA of n_pad entries each —
// centre x, centre y, (radius + margin)^2, radius + margin.  The margin delta = 64 eps B covers every rounding
// difference between the 3-FFMA line test and the exact tests; padding entries can never become candidates.
template <class T> struct BoundsTable {
  int n_pad = 0;
  double delta = 0;
  std::vector<T> tab;
};
template <class T> BoundsTable<T> build_bounds(const double *circ, size_t n, double B) {
  BoundsTable<T> bt;
  bt.n_pad = (int)((n + 31) / 32 * 32);
  const double eps = std::numeric_limits<T>::epsilon();
  bt.delta = 64.0 * eps * B;
  bt.tab.assign(4 * (size_t)bt.n_pad, (T)0);
  T *bx = bt.tab.data(), *by = bx + bt.n_pad, *br2 = by + bt.n_pad, *brb = br2 + bt.n_pad;
  for (int i = 0; i < bt.n_pad; ++i) {
    if ((size_t)i < n) {
      bx[i] = (T)circ[3 * i];
      by[i] = (T)circ[3 * i + 1];
      const double rb = (circ[3 * i + 2] * (1.0 + 1e-6) + bt.delta) * (1.0 + 4 * eps);
      brb[i] = std::nextafter((T)rb, std::numeric_limits<T>::max());
      br2[i] = std::nextafter((T)((double)brb[i] * (double)brb[i] * (1.0 + 4 * eps)), std::numeric_limits<T>::max());
    } else {
      bx[i] = by[i] = (T)0;
      br2[i] = (T)-1; // never a candidate
      brb[i] = (T)0;
    }
  }
  return bt;
}

// The uniform grid over the same padded circles (lg_tile_map_enable): a cell lists every object whose padded circle
// comes within 4 delta of it; the box keeps an empty rim of at least 16 delta.
template <class T> HostGrid build_scene_grid(const BoundsTable<T> &bt, size_t n, double density) {
  std::vector<double> circ(3 * n);
  const T *bx = bt.tab.data(), *by = bx + bt.n_pad, *brb = by + 2 * (size_t)bt.n_pad;
  for (size_t i = 0; i < n; ++i) circ[3 * i] = (double)bx[i], circ[3 * i + 1] = (double)by[i], circ[3 * i + 2] = (double)brb[i];
  return build_grid(circ.data(), n, 4 * bt.delta, 16 * bt.delta, sizeof(T) == 4, density);
}

} // namespace lg
