// ref_render.cc -- ORACLE (test infrastructure, never shipped, never on the product path).
//
// The reference keeps its hot path inline in main() of
// include/blackhole/blackhole_solution_test.cc:161-308, with resolution, nstep and scene as
// literals.  This driver restates that loop with those literals as parameters and runs it on the
// REFERENCE'S OWN classes: it is compiled against /root/reference/include (plus the stand-in
// opencv2/opencv.hpp), so Camera::PixelVector, StaticBlackhole::SolveG/InvSqrtG,
// ObjectManager::FindCollision, Annulus/Rectangle/InfinitePlane::Collide and ::color are the
// reference's code, not a port.  oracle/Makefile checks that at 960x540 / cfg 0 / nstep 20 the
// frame is byte-identical to the one the UNCHANGED blackhole_solution_test binary writes.
//
// Besides the BGR frame it records, per pixel, the hit object's ObjectManager key, its class and
// the number of geodesic updates, and it dumps the scene + camera as the bh8 POD snapshot (JSON)
// so the C port (oracle/bh8_oracle.c) and the CUDA path can be fed exactly the same inputs.
//
// Only the bit-exact arithmetic below matters; the loop order (x outer in the reference,
// :162-163) does not influence results and rows are distributed over OpenMP threads when built
// with -fopenmp (the loop body is re-entrant: FindCollision/Collide/color are const).
#include <chrono>
#include <cinttypes>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

#include "blackhole/ray_tracer.h"
#include "scenes.h"

namespace {

using value_type = double;
using point_type = cv::Vec<value_type, 3>;
using vector_type = cv::Vec<value_type, 3>;
using matrix_type = cv::Matx<value_type, 3, 3>;

struct PixelResult {
  unsigned char bgr[3] = {0, 0, 0};  // frame is cleared to 0 (blackhole_solution_test.cc:140)
  const bh8scenes::object_type* hit = nullptr;
  int steps = 0;
};

// One pixel of blackhole_solution_test.cc:164-298 with nstep as a parameter.
template <typename Manager>
PixelResult TracePixel(const blackhole::Camera<value_type>& camera, const vector_type& fv,
                       const blackhole::StaticBlackhole<value_type>& blackhole, const Manager& manager,
                       int x, int y, int nstep) {
  PixelResult out;

  const auto pv = camera.PixelVector(x, y, fv) - blackhole.position();  // :167
  const auto F = camera.focus() - blackhole.position();                 // :168

  vector_type yv = cv::normalize(F - pv);  // :170
  vector_type zv = cv::normalize(pv.cross(yv));
  if (std::isnan(zv[0])) zv = vector_type(1, 0, 0);
  vector_type xv = cv::normalize(yv.cross(zv));

  auto matrix_reconversion = matrix_type(zv[0], yv[0], xv[0],  //
                                         zv[1], yv[1], xv[1],  //
                                         zv[2], yv[2], xv[2]);  // :176-179
  auto matrix_conversion = matrix_reconversion.inv();           // :180

  auto convertedCameraFocus = matrix_conversion * F;  // :182

  const auto b = convertedCameraFocus[2];  // :185
  value_type sol = -1;
  if (b >= blackhole.b_c()) sol = blackhole.SolveG(b);
  const auto periapsis = b < blackhole.b_c() ? 1.0 / (3 * blackhole.mass()) : sol;  // :190
  auto r0 = std::sqrt(convertedCameraFocus.dot(convertedCameraFocus));

  value_type phi = atan(convertedCameraFocus[2] / convertedCameraFocus[1]);  // :193
  value_type dphi = 0;
  value_type dphi_prev = 0;
  value_type u = 1. / r0;
  value_type r = r0;

  const value_type integrate_begin = u;
  const value_type integrate_end = periapsis;

  const int nstep_safe = nstep - 1;                                          // :203
  value_type du = (integrate_end - integrate_begin) * (1.0 / nstep);         // :204
  const value_type du_h = du / 2.;                                           // :205

  auto light_vector = convertedCameraFocus;
  point_type light_vector_original = camera.focus();
  point_type light_vector_prev_original = camera.focus();  // :210-211
  point_type inter;

  // One geodesic update + segment test; `delta` is +du, +0.9du or -du (:218-235, :241-258, :275-290).
  auto advance = [&](value_type delta) -> bool {
    u += delta;
    dphi = blackhole.InvSqrtG(u, b);
    phi += (dphi_prev + dphi) * du_h;
    r = 1. / u;
    light_vector = {0, r * cos(phi), r * sin(phi)};
    light_vector_original = (matrix_reconversion * light_vector) + blackhole.position();
    dphi_prev = dphi;
    ++out.steps;
    if (const auto& obj = manager.FindCollision(light_vector_prev_original, light_vector_original, &inter);
        obj != nullptr) {
      const auto c = obj->color(inter);
      out.bgr[0] = c[0];
      out.bgr[1] = c[1];
      out.bgr[2] = c[2];
      out.hit = obj;
      return true;
    }
    light_vector_prev_original = light_vector_original;
    return false;
  };

  for (int i = 0; i < nstep_safe; ++i)  // :217
    if (advance(du)) return out;
  if (advance(du * 0.9)) return out;  // :241

  if (sol < 0) {  // :264-272
    if (const auto& obj = manager.FindCollision(light_vector_prev_original, blackhole.center(), &inter);
        obj != nullptr) {
      const auto c = obj->color(inter);
      out.bgr[0] = c[0];
      out.bgr[1] = c[1];
      out.bgr[2] = c[2];
      out.hit = obj;
    }
    return out;
  }

  for (int i = 0; i < nstep_safe; ++i)  // :274
    if (advance(-du)) return out;
  return out;
}

// One pixel of the flat-space driver, ray_tracer_test.cc:143-145: the reference's RayTracer class
// itself (BasicLinearRayRecurrence, Prograde) does the work.  Prograde writes the colour but does
// not say which object was hit or after how many segments, so the same march is repeated with
// FindCollision to recover the key and the count (identical arithmetic, ray_tracer.h:24-31,68-85).
template <typename Manager>
PixelResult TracePixelLinear(const blackhole::Camera<value_type>& camera, const vector_type& fv,
                             const Manager& manager, int x, int y, int steps) {
  PixelResult out;
  blackhole::RayTracer ray_tracer(camera.focus(), camera.PixelVector(x, y, fv));
  Manager& mutable_manager = const_cast<Manager&>(manager);
  ray_tracer.Prograde(mutable_manager, out.bgr, steps);

  point_type old = camera.focus(), present = camera.PixelVector(x, y, fv), inter;
  blackhole::BasicLinearRayRecurrence<point_type> next(old, present);
  for (int s = 0; s < steps; ++s) {
    ++out.steps;
    if (const auto* obj = manager.FindCollision(old, present, &inter); obj != nullptr) {
      out.hit = obj;
      break;
    }
    auto p = next(old, present);
    old = present;
    present = p;
  }
  return out;
}

void PrintVec(FILE* f, const char* name, const double* v, int n, bool comma = true) {
  std::fprintf(f, "\"%s\": [", name);
  for (int i = 0; i < n; ++i) std::fprintf(f, "%s%.17g", i ? ", " : "", v[i]);
  std::fprintf(f, "]%s", comma ? ", " : "");
}

void WriteFile(const std::string& path, const void* data, size_t bytes) {
  FILE* f = std::fopen(path.c_str(), "wb");
  if (!f) {
    std::fprintf(stderr, "cannot write %s\n", path.c_str());
    std::exit(2);
  }
  std::fwrite(data, 1, bytes, f);
  std::fclose(f);
}

}  // namespace

int main(int argc, char** argv) {
  int cfg = 0, width = 960, height = 540, frame = 0, nstep = -1, threads = 1, repeat = 1;
  std::string texdir = "build/textures", out;
  bool snapshot_only = false;
  int row0 = 0, row1 = -1;
  for (int i = 1; i < argc; ++i) {
    const std::string a = argv[i];
    auto next = [&]() -> const char* { return i + 1 < argc ? argv[++i] : ""; };
    if (a == "--cfg") cfg = std::atoi(next());
    else if (a == "--width") width = std::atoi(next());
    else if (a == "--height") height = std::atoi(next());
    else if (a == "--frame") frame = std::atoi(next());
    else if (a == "--nstep") nstep = std::atoi(next());
    else if (a == "--threads") threads = std::atoi(next());
    else if (a == "--repeat") repeat = std::atoi(next());
    else if (a == "--texdir") texdir = next();
    else if (a == "--out") out = next();
    else if (a == "--rows") { row0 = std::atoi(next()); row1 = std::atoi(next()); }
    else if (a == "--snapshot-only") snapshot_only = true;
    else {
      std::fprintf(stderr,
                   "usage: ref_render --cfg N --width W --height H [--frame K] [--nstep S] [--threads T]\n"
                   "                  [--repeat R] [--rows y0 y1] [--texdir DIR] [--snapshot-only] --out PREFIX\n");
      return 2;
    }
  }
  if (row1 < 0 || row1 > height) row1 = height;

  bh8scenes::Scene* scene = bh8scenes::Build(cfg, width, height, frame, texdir);
  if (!scene) {
    std::fprintf(stderr, "unknown cfg %d\n", cfg);
    return 2;
  }
  if (nstep <= 0) nstep = scene->nstep;
  const int linear_steps = scene->linear_steps;
  auto& manager = bh8scenes::manager_type::GetInstance();
  const auto& camera = scene->camera;

  const size_t npix = static_cast<size_t>(width) * height;
  std::vector<unsigned char> bgr(npix * 3, 0), cls(npix, 0);
  std::vector<signed char> key(npix, -1);
  std::vector<uint16_t> steps(npix, 0);

  double best_ms = 0, sum_ms = 0;
  uint64_t total_steps = 0, class_count[4] = {0, 0, 0, 0};
  if (!snapshot_only) {
#ifdef _OPENMP
    omp_set_num_threads(threads > 0 ? threads : 1);
#else
    threads = 1;
#endif
    for (int rep = 0; rep < repeat; ++rep) {
      const auto t1 = std::chrono::steady_clock::now();
      const auto fv = camera.focus_vector();  // :161
#pragma omp parallel for schedule(dynamic, 4)
      for (int y = row0; y < row1; ++y) {
        for (int x = 0; x < width; ++x) {
          const PixelResult p = linear_steps > 0
                                    ? TracePixelLinear(camera, fv, manager, x, y, linear_steps)
                                    : TracePixel(camera, fv, *scene->blackhole, manager, x, y, nstep);
          const size_t i = static_cast<size_t>(y) * width + x;  // :214
          bgr[i * 3 + 0] = p.bgr[0];
          bgr[i * 3 + 1] = p.bgr[1];
          bgr[i * 3 + 2] = p.bgr[2];
          steps[i] = static_cast<uint16_t>(p.steps);
          int k = -1, c = BH8_CLASS_BACKGROUND;
          if (p.hit) {
            for (const auto& e : scene->objects)
              if (e.object == p.hit) {
                k = e.key;
                c = e.kind == BH8_KIND_BLACKHOLE ? BH8_CLASS_HORIZON
                    : e.kind == BH8_KIND_ANNULUS ? BH8_CLASS_DISC
                                                 : BH8_CLASS_OBJECT;
              }
          }
          key[i] = static_cast<signed char>(k);
          cls[i] = static_cast<unsigned char>(c);
        }
      }
      const auto t2 = std::chrono::steady_clock::now();
      const double ms = std::chrono::duration<double, std::milli>(t2 - t1).count();
      sum_ms += ms;
      if (rep == 0 || ms < best_ms) best_ms = ms;
    }
    for (size_t i = 0; i < npix; ++i) {
      total_steps += steps[i];
      ++class_count[cls[i]];
    }
  }

  // ---- snapshot + run summary (JSON on stdout and PREFIX.json) ----
  int bh_index = -1;
  std::vector<std::string> textures;
  const std::vector<bh8_object> objs = bh8scenes::SnapshotObjects(*scene, &bh_index, &textures);
  const bh8_camera cam = bh8scenes::SnapshotCamera(camera);

  std::string json_path = out.empty() ? "" : out + ".json";
  FILE* jf = json_path.empty() ? stdout : std::fopen(json_path.c_str(), "w");
  if (!jf) {
    std::fprintf(stderr, "cannot write %s\n", json_path.c_str());
    return 2;
  }
  std::fprintf(jf, "{\"cfg\": %d, \"frame\": %d, \"nstep\": %d, \"linear_steps\": %d, \"width\": %d, \"height\": %d,\n",
               cfg, frame, nstep, linear_steps, width, height);
  std::fprintf(jf, " \"camera\": {");
  PrintVec(jf, "pos", cam.pos, 3);
  PrintVec(jf, "vx", cam.vx, 3);
  PrintVec(jf, "vy", cam.vy, 3);
  PrintVec(jf, "vz", cam.vz, 3);
  std::fprintf(jf, "\"focus_len\": %.17g, \"width\": %d, \"height\": %d},\n", cam.focus_len, cam.width,
               cam.height);
  std::fprintf(jf, " \"textures\": [");
  for (size_t t = 0; t < textures.size(); ++t) std::fprintf(jf, "%s\"%s\"", t ? ", " : "", textures[t].c_str());
  std::fprintf(jf, "],\n \"bh_index\": %d,\n \"objects\": [\n", bh_index);
  for (size_t k = 0; k < objs.size(); ++k) {
    const bh8_object& o = objs[k];
    std::fprintf(jf, "  {\"kind\": %d, \"key\": %d, \"tex_id\": %d, \"pattern\": %d, ", o.kind, o.key, o.tex_id,
                 o.pattern);
    PrintVec(jf, "v", &o.v[0][0], 15);
    PrintVec(jf, "n", o.n, 3);
    PrintVec(jf, "ex", o.ex, 3);
    PrintVec(jf, "ey", o.ey, 3);
    std::fprintf(jf, "\"r_in\": %.17g, \"r_out\": %.17g, \"mass\": %.17g, \"pattern_size\": %.17g}%s\n", o.r_in,
                 o.r_out, o.mass, o.pattern_size, k + 1 < objs.size() ? "," : "");
  }
  std::fprintf(jf, " ],\n");
  std::fprintf(jf,
               " \"run\": {\"rendered\": %s, \"rows\": [%d, %d], \"threads\": %d, \"repeat\": %d, "
               "\"best_ms\": %.3f, \"mean_ms\": %.3f, \"rays\": %zu, \"steps\": %" PRIu64
               ", \"class_count\": [%" PRIu64 ", %" PRIu64 ", %" PRIu64 ", %" PRIu64 "]}\n}\n",
               snapshot_only ? "false" : "true", row0, row1, threads, repeat, best_ms,
               repeat ? sum_ms / repeat : 0.0, static_cast<size_t>(row1 - row0) * width, total_steps,
               class_count[0], class_count[1], class_count[2], class_count[3]);
  if (jf != stdout) std::fclose(jf);

  if (!out.empty() && !snapshot_only) {
    std::vector<unsigned char> raw(8 + bgr.size());
    const int32_t hdr[2] = {height, width};
    std::memcpy(raw.data(), hdr, 8);
    std::memcpy(raw.data() + 8, bgr.data(), bgr.size());
    WriteFile(out + ".bgr", raw.data(), raw.size());
    WriteFile(out + ".cls", cls.data(), cls.size());
    WriteFile(out + ".key", key.data(), key.size());
    WriteFile(out + ".steps", steps.data(), steps.size() * sizeof(uint16_t));
  }
  return 0;
}
