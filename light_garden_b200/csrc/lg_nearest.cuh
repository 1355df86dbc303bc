// lg_nearest.cuh — the nearest-hit search of Tracer::trace (src/light_garden/tracer.rs:385-424) over the lowered
// scene: the exact per-object test (narrow phase), the all-objects loop and the walk through the uniform grid.
// __host__ __device__ like lg_geom.cuh, so tests/host_geom_check.cu runs the very same code on the CPU.
#pragma once
#include "lg_geom.cuh"

#ifdef __CUDA_ARCH__
#define LG_LDG(p) __ldg(p)
#else
#define LG_LDG(p) (*(p))
#endif

namespace lg {

// the part of the device scene the nearest-hit search reads
template <class T> struct SceneArgs {
  const Tok<T> *toks;
  const int *obj_first, *obj_count;
  // every object is ONE leaf token and token i belongs to object i (scenes of plain circles / rects / mirrors, C5):
  // the narrow phase then reads toks[obj] directly instead of waiting for obj_first[obj] first -- one dependent
  // global load less on the path the nearest-hit search stalls on (ncu r02: long_scoreboard, ~10 % of the samples)
  int flat;
  T delta; // rounding margin of the broad phase (64 eps x coordinate bound)
  // uniform grid over the objects' bounding circles (lg_tile_map_enable; SURVEY.md 8f rank 1, the device-side
  // stand-in for tile_map.rs): cell (ix, iy) covers [x0 + ix cs, x0 + (ix + 1) cs) x [y0 + iy cs, ...) and lists
  // every object whose padded bounding circle comes within 4 delta of it (CSR: grid_start[cell], grid_obj[])
  T grid_x0, grid_y0, grid_x1, grid_y1, grid_cs, grid_ics, grid_eta;
  int grid_nx, grid_ny;
  const unsigned int *grid_start, *grid_obj;
};

// ---- best-hit bookkeeping (tracer.rs:412-424) -------------------------------------
template <class T> struct Best {
  T d2;
  T px, py, aux;
  int obj, tok;
};

template <class T>
LG_HD void take(Best<T> &b, V2<T> o, const Cand<T> &c, int obj, int tok) {
  T dx = c.p.x - o.x, dy = c.p.y - o.y;
  T d2 = dx * dx + dy * dy; // nalgebra distance_squared, no fused multiply-add
  // strict `<`; the sweep visits objects grouped by type, so equal distances
  // are resolved towards the lower object index, as the in-order loop would
  if (d2 < b.d2 || (d2 == b.d2 && obj < b.obj)) {
    b.d2 = d2;
    b.px = c.p.x;
    b.py = c.p.y;
    b.aux = c.aux;
    b.obj = obj;
    b.tok = tok;
  }
}

// Ray::intersect(&Geo::GeoLogic): leaf hits in program order, each filtered by
// the sibling subtrees on the way to the root (ORACLE.md §3.6)
// (force-inlined, like narrow_phase below: with two levels of __noinline__ device functions the sm_100a build of
// nvcc 12.9 produced a kernel that lost a live register across the nested call — compute-sanitizer: misaligned
// shared-memory reads in the candidate loop right after CALL narrow_phase -> CALL sweep_csg_object.  The kernel
// therefore contains no user-level calls; tools/gpu_round.sh runs the GPU suite under compute-sanitizer.)
template <class T>
LG_HD void sweep_csg_object(const SceneArgs<T> &A, int obj, V2<T> o, V2<T> d, Best<T> &best) {
  const int first = A.obj_first[obj], count = A.obj_count[obj];
  const Tok<T> *tok = A.toks + first;
  for (int k = 0; k < count; ++k) {
    const Tok<T> &l = tok[k];
    if (l.kind == TOK_OP || l.kind == TOK_POINTS) continue;
    CandList<T> hl;
    hl.n = 0;
    if (l.kind == TOK_CIRCLE)
      hit_circle(l.p, o, d, hl);
    else if (l.kind == TOK_RECT)
      hit_rect(l.p, o, d, hl);
    else if (l.kind == TOK_SEGMENT)
      hit_segment(l.p, o, d, hl);
    else if (l.kind == TOK_ELLIPSE)
      hit_ellipse(l.p, o, d, hl);
    else if (l.kind == TOK_POLY)
      hit_poly(l, o, d, hl);
    else
      hit_bezier(l.p, o, d, hl);
    for (int j = 0; j < hl.n; ++j) {
      bool keep = true;
      for (int i = k + 1; i < count && keep; ++i) {
        const Tok<T> &q = tok[i];
        if (q.kind != TOK_OP || q.a_start > k) continue;
        if (k < q.b_start) {
          bool inb = contains_range(tok, q.b_start, i - 1, hl.h[j].p);
          keep = (q.op == OP_AND) ? inb : !inb;
        } else {
          bool ina = contains_range(tok, q.a_start, q.b_start - 1, hl.h[j].p);
          keep = (q.op == OP_OR) ? !ina : ina;
        }
      }
      if (keep) take(best, o, hl.h[j], obj, first + k);
    }
  }
}

// Narrow phase: the exact Ray::intersect of ORACLE.md §3 for one object.  It runs in the candidate loop that
// follows every 32-object chunk of the broad phase, so the broad-phase loop itself stays a handful of instructions
// per test.
template <class T>
LG_HD Best<T> narrow_phase(const SceneArgs<T> &A, Best<T> b, int obj, V2<T> o, V2<T> d) {
  const int first = A.flat ? obj : A.obj_first[obj];
  if (A.flat || A.obj_count[obj] == 1) {
    const Tok<T> &k = A.toks[first];
    auto emit = [&](const Cand<T> &h) { take(b, o, h, obj, first); }; // hits stay in registers, library order
    if (k.kind == TOK_CIRCLE) {
      hit_circle_each(k.p, o, d, emit);
    } else if (k.kind == TOK_SEGMENT) {
      hit_segment_each(k.p, o, d, emit);
    } else if (k.kind == TOK_RECT) {
      hit_rect_each(k.p, o, d, emit);
    } else {
      CandList<T> hl;
      hl.n = 0;
      if (k.kind == TOK_ELLIPSE)
        hit_ellipse(k.p, o, d, hl);
      else
        hit_bezier(k.p, o, d, hl);
      for (int q = 0; q < hl.n; ++q) take(b, o, hl.h[q], obj, first);
    }
  } else {
    sweep_csg_object(A, obj, o, d, b);
  }
  return b;
}

LG_HD float g_min(float a, float b) { return ::fminf(a, b); }
LG_HD double g_min(double a, double b) { return ::fmin(a, b); }
LG_HD float g_max(float a, float b) { return ::fmaxf(a, b); }
LG_HD double g_max(double a, double b) { return ::fmax(a, b); }
LG_HD float g_floor(float a) { return ::floorf(a); }
LG_HD double g_floor(double a) { return ::floor(a); }

// Nearest hit through the uniform grid: the cells the ray crosses are visited in ray order and only the objects
// listed there run the exact test.  The result is the brute-force loop's, bit for bit: every object with an actual
// hit at parameter t is listed in a cell visited before the walk may stop (cells list objects within grid_eta of
// them, far more than the rounding of the cell-crossing parameters computed here), the exact test does not depend on
// the cell, and take() resolves equal distances towards the lower object index whatever the visiting order.
template <class T>
LG_HD Best<T> grid_nearest(const SceneArgs<T> &A, Best<T> best, V2<T> o, V2<T> d) {
  const T big = Real<T>::max_value();
  const bool hx = d.x != (T)0, hy = d.y != (T)0;
  const T idx = hx ? Real<T>::div((T)1, d.x) : (T)0, idy = hy ? Real<T>::div((T)1, d.y) : (T)0;
  // the stretch of the ray inside the grid's box (the box is padded well beyond every listed circle)
  T t0 = (T)0, t1 = big;
  if (hx) {
    const T ta = (A.grid_x0 - o.x) * idx, tb = (A.grid_x1 - o.x) * idx;
    t0 = g_max(t0, g_min(ta, tb));
    t1 = g_min(t1, g_max(ta, tb));
  } else if (o.x < A.grid_x0 || o.x > A.grid_x1) {
    return best;
  }
  if (hy) {
    const T ta = (A.grid_y0 - o.y) * idy, tb = (A.grid_y1 - o.y) * idy;
    t0 = g_max(t0, g_min(ta, tb));
    t1 = g_min(t1, g_max(ta, tb));
  } else if (o.y < A.grid_y0 || o.y > A.grid_y1) {
    return best;
  }
  if (!(t0 <= t1)) return best;
  const int nx = A.grid_nx, ny = A.grid_ny;
  const T px = Real<T>::fma(t0, d.x, o.x), py = Real<T>::fma(t0, d.y, o.y);
  int ix = (int)g_floor((px - A.grid_x0) * A.grid_ics), iy = (int)g_floor((py - A.grid_y0) * A.grid_ics);
  ix = ix < 0 ? 0 : (ix > nx - 1 ? nx - 1 : ix), iy = iy < 0 ? 0 : (iy > ny - 1 ? ny - 1 : iy);
  const int sx = d.x > (T)0 ? 1 : -1, sy = d.y > (T)0 ? 1 : -1;
  T tbest = big; // upper bound of the best hit's ray parameter
  int last = -1; // the object tested last: neighbouring cells mostly list the same one again
  for (int guard = nx + ny + 4; guard > 0; --guard) {
    const unsigned cell = (unsigned)(iy * nx + ix);
    const unsigned q0 = LG_LDG(A.grid_start + cell), q1 = LG_LDG(A.grid_start + cell + 1);
    for (unsigned q = q0; q < q1; ++q) {
      const int obj = (int)LG_LDG(A.grid_obj + q);
      if (obj == last) continue;
      last = obj;
      const T before = best.d2;
      best = narrow_phase(A, best, obj, o, d);
      if (best.d2 != before) tbest = Real<T>::sqrt(best.d2) * (T)1.000001 + A.delta;
    }
    // where the ray leaves this cell
    const T bxn = Real<T>::fma((T)(ix + (sx > 0 ? 1 : 0)), A.grid_cs, A.grid_x0);
    const T byn = Real<T>::fma((T)(iy + (sy > 0 ? 1 : 0)), A.grid_cs, A.grid_y0);
    const T tx = hx ? (bxn - o.x) * idx : big, ty = hy ? (byn - o.y) * idy : big;
    if (tbest < g_min(tx, ty) - A.grid_eta) break; // nothing in the cells ahead can be nearer
    if (tx < ty) {
      ix += sx;
      if ((unsigned)ix >= (unsigned)nx) break;
    } else {
      iy += sy;
      if ((unsigned)iy >= (unsigned)ny) break;
    }
  }
  return best;
}

// The all-objects loop of tracer.rs:412-424 in object order (what the broad phase of the trace kernel prunes without
// changing the result); the host cross-check compares the grid walk with it.
template <class T> LG_HD Best<T> all_objects_nearest(const SceneArgs<T> &A, int n_obj, Best<T> best, V2<T> o, V2<T> d) {
  for (int j = 0; j < n_obj; ++j) best = narrow_phase(A, best, j, o, d);
  return best;
}

} // namespace lg
