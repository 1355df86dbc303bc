// tests/host_cpp_smoke.cpp — the C++ host mirror (light_garden_b200/host/lg_tracer.hpp) driving the C ABI:
// builds default.ron's scene through the reference's constructor names, traces it, renders it, prints a digest
// that tests/test_host_cpp.py compares with the Python host layer.  Needs a GPU.
#include <algorithm>
#include <cmath>
#include <cstdio>

#include "../light_garden_b200/host/lg_tracer.hpp"

int main() {
  try {
    const double aspect = (double)(480.0f / 270.0f);
    lg::Tracer t(lg::Rect::from_tlbr(1., -aspect, -1., aspect));
    lg::Object lens = lg::Object::new_lens({-0.022772240638732733, -0.09999999999999998}, 2.0, 3.8);
    lens.geo.rot = {0.00000000000000006123233995736766, -1.0, 1.0, 0.00000000000000006123233995736766};
    lens.material_opt = lg::Material{1.05};
    t.push_object(lens);
    t.push_object(lg::Object::new_curved_mirror({lg::P2{-0.622772240638733, 0.40000000000000013}, lg::P2{-0.3227722406387328, 0.8},
                                                  lg::P2{0.2772277593612673, 0.8}, lg::P2{0.5772277593612682, 0.40000000000000013}}));
    lg::Object rect = lg::Object::new_rect({-0.022772240638732733, -0.5}, 0.40000000000000036, 0.3999999999999999);
    rect.material_opt = lg::Material{1.73};
    t.push_object(rect);
    t.push_light(lg::Light::point({-0.022772240638732733, -0.5}, 5000, {0.009721218f, 0.009721218f, 0.009721218f, 0.011764706f}));
    t.push_light(lg::Light::spot({1.0772277593612674, -0.09999999999999998}, 0.17453292519943295,
                                 {-0.9999922358557027, 0.003940587305547067}, 1000,
                                 {0.0036765062f, 0.020288562f, 0.016807375f, 0.03529412f}));
    auto lines = t.trace_all();
    double sx = 0, sy = 0, sc = 0;
    for (auto &v : lines) sx += v.first.x, sy += v.first.y, sc += v.second[0] + v.second[1] + v.second[2];
    std::printf("vertices %zu segments %llu ray_steps %llu\n", lines.size(), (unsigned long long)t.last_stats.segments,
                (unsigned long long)t.last_stats.ray_steps);
    std::printf("checksum %.9e %.9e %.9e\n", sx, sy, sc);
    // boundary B2 through the same mirror: the traced lines as a host LineList vs the fused device frame
    lg::Renderer rend(t, 480, 270);
    rend.render_lines(lines);
    const unsigned long long frag_lines = rend.last_stats.pixel_updates;
    const std::vector<float> img_lines = rend.read_rgba32f();
    rend.clear();
    rend.render();                      // trace + accumulate on the device; the control lines are host lines
    rend.render_lines(std::vector<std::pair<lg::P2, lg::Color>>(lines.end() - 6, lines.end()));
    const std::vector<float> img_dev = rend.read_rgba32f();
    double si = 0, worst = 0;
    for (size_t k = 0; k < img_dev.size(); ++k) {
      si += img_dev[k];
      const double tol = 1e-5 * std::max(1.0, (double)std::fabs(img_lines[k]));
      worst = std::max(worst, std::fabs((double)img_dev[k] - img_lines[k]) / tol);
    }
    const std::vector<uint8_t> shot = rend.make_screenshot(true, 2048), surf = rend.make_screenshot(false);
    unsigned long long sb = 0, ss = 0;
    for (uint8_t v : shot) sb += v;
    for (uint8_t v : surf) ss += v;
    std::printf("image fragments %llu sum %.9e worst_tol_ratio %.3f screenshot %llu surface %llu\n", frag_lines, si, worst, sb, ss);
    // Mode::StringMod through the mirror: 3000 chords i -> 2 i mod 3000 with the three colour rules of BASELINE's C4
    lg::StringMod sm;
    sm.modulo = 3000, sm.num = 2;
    sm.color = {1e-2f, 1e-2f, 1e-2f, 1e-2f};
    sm.modulo_colors = {{3, 0, {1e-2f, 0.f, 0.f, 1e-2f}}, {3, 1, {0.f, 1e-2f, 0.f, 1e-2f}}, {3, 2, {0.f, 0.f, 1e-2f, 1e-2f}}};
    lg::Renderer srend(t, 256, 256);
    srend.render_string_mod(sm);
    double ssum = 0;
    for (float v : srend.read_rgba32f()) ssum += v;
    std::printf("string_mod fragments %llu sum %.9e\n", (unsigned long long)srend.last_stats.pixel_updates, ssum);
    // the remaining Object constructors (object.rs:34-45): a prism and an ellipse in front of a point light
    lg::Tracer extra(lg::Rect::from_tlbr(1., -aspect, -1., aspect));
    extra.push_object(lg::Object::new_convex_polygon({{-0.5, -0.3}, {0.5, -0.3}, {0.0, 0.5}, {0.0, 0.0}}));
    extra.push_object(lg::Object::new_ellipse({0.9, 0.1}, 0.3, 0.15));
    extra.push_light(lg::Light::point({-1.2, 0.0}, 2000, {0.01f, 0.01f, 0.01f, 0.02f}));
    extra.enable_tile_map(true);
    const size_t with_grid = extra.trace_all().size();
    extra.enable_tile_map(false);
    std::printf("polygon+ellipse vertices %zu (tile map) %zu (all objects)\n", with_grid, extra.trace_all().size());
    // error behaviour: the reference panics on a bad scene, here the error crosses the boundary as a status
    lg::Tracer bad(lg::Rect::from_tlbr(1, -1, -1, 1));
    lg::Object neg = lg::Object::new_circle({0, 0}, 0.5);
    neg.material_opt = lg::Material{-1.0};
    bad.push_object(neg);
    try {
      bad.trace_all();
      std::printf("error NOT raised\n");
      return 1;
    } catch (const lg::Error &e) {
      std::printf("error %d %s\n", e.code, e.what());
    }
    return 0;
  } catch (const lg::Error &e) {
    std::printf("FAILED %d %s\n", e.code, e.what());
    return 2;
  }
}
