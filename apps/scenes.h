// scenes.h -- the BASELINE.json configurations, built through the blackhole:: header API.
//
// Scene source of the GPU driver (apps/blackhole_solution_gpu.cc); the oracle's restatement
// (oracle/ref_render.cc) builds the SAME scenes from it through the reference's own headers.  It is not
// part of the oracle and contains no rendering code.  This file only
// uses the public API that lackhole/blackhole_8 ships in include/blackhole/ (Camera,
// ObjectManager, Annulus, Rectangle, InfinitePlane, ChessPattern2D, StaticBlackhole), so it
// compiles unchanged against EITHER the reference's own headers (-I/root/reference/include, used
// by oracle/ref_render.cc -> oracle/_ref/) OR this repository's source-compatible headers
// (-Iinclude).  Scene definitions follow SURVEY.md section 8(d):
//
//   cfg 0 ("0a", as shipped)  blackhole_solution_test.cc:61-62,103-112,148
//   cfg 5 ("0b", background only)  reference idiom blackhole_solution_test.cc:85-92
//   cfg 1  disc + background rectangle
//   cfg 2  multiple_blackhole_solution_test2.cc:60-96,129 scene verbatim
//   cfg 3  fly-through frame k of cfg 1 (disc spin blackhole_solution_test.cc:407, camera script in
//          the reference's key actions :346-389)
//   cfg 4  cfg 1 scene at nstep 200
//   cfg 6-9  stress scenes for the GPU path's rarely taken branches (not BASELINE configs):
//          6 hole moved off the disc's plane (arrow keys, blackhole_solution_test.cc:391-396) and
//            camera tilted; 7 camera turned away from the hole (mirrored-start rays, the atan quirk);
//          8 more planes than the kernel has filter slots; 9 camera inside the photon sphere;
//          11 exactly four filtered planes
//   cfg 10 the flat-space scene of ray_tracer_test.cc:45-98 (800x450, 3 textured rectangles + chess
//          floor, no hole), frame k after k rounds of its object animation (:237-261)
//
// Because ObjectManager is a leaked singleton without enumeration (object_manager.h:30-33,92) the
// builder keeps its own list of what it inserted, in insertion order.
#ifndef BH8_ORACLE_SCENES_H_
#define BH8_ORACLE_SCENES_H_

#include <string>
#include <vector>

#include "blackhole/blackhole_solution.h"
#include "blackhole/camera.h"
#include "blackhole/constants.h"
#include "blackhole/object.h"
#include "blackhole/object/object_manager.h"

#include "bh8.h"

namespace bh8scenes {

using value_type = double;
using point_type = cv::Vec<value_type, 3>;
using manager_type = blackhole::ObjectManager<value_type>;
using object_type = blackhole::DrawableObject<value_type>;

struct Inserted {
  int kind = 0;
  int key = 0;
  const object_type* object = nullptr;
  std::string texture;       // reference resource name, "" = none
  double r_in = 0, r_out = 0;  // Annulus ctor arguments (no getters in the reference)
  double pattern_size = 0;
  int pattern = BH8_PATTERN_BLACK;
};

struct Scene {
  blackhole::Camera<value_type> camera;
  std::vector<Inserted> objects;  // insertion order; key == index
  blackhole::StaticBlackhole<value_type>* blackhole = nullptr;
  blackhole::Annulus<value_type>* disc = nullptr;
  int nstep = 20;
  int linear_steps = 0;  // > 0: flat-space scene, traced with RayTracer::Prograde(.., linear_steps)
  object_type* movers[3] = {nullptr, nullptr, nullptr};  // cfg 10: mooni, karina, winter

  Scene(int w, int h, value_type fov) : camera(w, h, fov) {}
};

inline cv::Mat LoadTexture(const std::string& dir, const std::string& name) {
  return cv::imread(dir + "/" + name, cv::IMREAD_COLOR);
}

inline void AddDisc(Scene* s, manager_type& m, const std::string& texdir) {
  // blackhole_solution_test.cc:103-111, literals verbatim (-0100 is octal: -64).
  auto id = m.InsertObject<blackhole::Annulus>(point_type{1000, 1000, 0}, point_type{-1000, 1000, 0},
                                               point_type{-1000, -0100, 0}, point_type{1000, -1000, 0},
                                               1000, 100);
  id.second->SetTexture(LoadTexture(texdir, "acc_disc.png"));
  Inserted e;
  e.kind = BH8_KIND_ANNULUS;
  e.key = id.first;
  e.object = id.second;
  e.texture = "acc_disc.png";
  e.r_out = 1000;
  e.r_in = 100;
  s->objects.push_back(e);
  s->disc = id.second;
}

inline object_type* AddRectangle(Scene* s, manager_type& m, const std::string& texdir, const std::string& tex,
                                 double x1, double y1, double z1, double x2, double y2, double z2, double x3,
                                 double y3, double z3, double x4, double y4, double z4) {
  auto id = m.InsertObject<blackhole::Rectangle>(x1, y1, z1, x2, y2, z2, x3, y3, z3, x4, y4, z4);
  id.second->SetTexture(LoadTexture(texdir, tex));
  Inserted e;
  e.kind = BH8_KIND_RECTANGLE;
  e.key = id.first;
  e.object = id.second;
  e.texture = tex;
  s->objects.push_back(e);
  return id.second;
}

inline void AddBackground(Scene* s, manager_type& m, const std::string& texdir) {
  // SURVEY 8(d) cfg 0b/1: one textured Rectangle behind the hole, inside r < r0.
  AddRectangle(s, m, texdir, "winter.jpg", 1000, 1800, 1400, 1000, -1800, 1400, 1000, -1800, -1000,
               1000, 1800, -1000);
}

inline void AddChess(Scene* s, manager_type& m, int size) {
  auto id = m.InsertObject<blackhole::InfinitePlane>(point_type(0, 0, 0),
                                                     blackhole::ChessPattern2D<value_type>{size});
  Inserted e;
  e.kind = BH8_KIND_INFINITE_PLANE;
  e.key = id.first;
  e.object = id.second;
  e.pattern = BH8_PATTERN_CHESS;
  e.pattern_size = size;
  s->objects.push_back(e);
}

inline void AddBlackhole(Scene* s, manager_type& m, double mass) {
  auto id = m.InsertObject<blackhole::StaticBlackhole>(point_type(0, 0, 0), mass);
  Inserted e;
  e.kind = BH8_KIND_BLACKHOLE;
  e.key = id.first;
  e.object = id.second;
  s->objects.push_back(e);
  s->blackhole = id.second;
}

// Builds configuration `cfg` at frame `frame` into the ObjectManager singleton (call once per
// process).  Returns nullptr for an unknown cfg.
inline Scene* Build(int cfg, int width, int height, int frame, const std::string& texdir) {
  auto& m = manager_type::GetInstance();
  Scene* s = nullptr;
  switch (cfg) {
    case 0:  // as shipped: disc only
      s = new Scene(width, height, blackhole::pi / 2);
      s->camera.MoveTo(-2000, 0, 400);
      AddDisc(s, m, texdir);
      AddBlackhole(s, m, 10);
      break;
    case 5:  // "0b": background rectangle only
      s = new Scene(width, height, blackhole::pi / 2);
      s->camera.MoveTo(-2000, 0, 400);
      AddBackground(s, m, texdir);
      AddBlackhole(s, m, 10);
      break;
    case 1:
    case 3:
    case 4:
      s = new Scene(width, height, blackhole::pi / 2);
      s->camera.MoveTo(-2000, 0, 400);
      AddDisc(s, m, texdir);
      AddBackground(s, m, texdir);
      AddBlackhole(s, m, 10);
      if (cfg == 4) s->nstep = 200;
      break;
    case 2:  // multiple_blackhole_solution_test2.cc:60-96,129
      s = new Scene(width, height, blackhole::pi / 2);
      s->camera.MoveTo(-2000, 120, 10);
      s->camera.fov(s->camera.fov() / 3);
      AddRectangle(s, m, texdir, "mooni.jpeg", 100, 253, 199, 100, 0, 199, 100, 0, 0, 100, 253, 0);
      AddRectangle(s, m, texdir, "karina.jpeg", 100, 0, 300, 100, -200, 300, 100, -200, 0, 100, 0, 0);
      AddRectangle(s, m, texdir, "winter.jpg", -100, 50, 100, -100, -50, 100, -100, -50, 0, -100, 50, 0);
      AddChess(s, m, 10);
      AddBlackhole(s, m, 20);
      break;
    case 6:  // hole off the disc plane: the disc and the rectangle are both "non-central"
      s = new Scene(width, height, blackhole::pi / 2);
      s->camera.MoveTo(-2000, 0, 400);
      s->camera.RotateY(0.1);
      AddDisc(s, m, texdir);
      AddBackground(s, m, texdir);
      AddBlackhole(s, m, 10);
      for (int k = 0; k < 10; ++k) s->blackhole->MoveZ(5);   // kUp x10
      for (int k = 0; k < 6; ++k) s->blackhole->MoveY(-5);   // kRight x6
      break;
    case 7:  // camera looks mostly away from the hole: F.pv > F.F for many pixels
      s = new Scene(width, height, blackhole::pi / 2);
      s->camera.MoveTo(-300, 40, 120);
      s->camera.RotateZ(2.2);
      AddDisc(s, m, texdir);
      AddBackground(s, m, texdir);
      AddBlackhole(s, m, 10);
      break;
    case 8:  // seven planes that do not contain the hole's centre + the chess floor + the disc
      s = new Scene(width, height, blackhole::pi / 2);
      s->camera.MoveTo(-900, 150, 260);
      AddDisc(s, m, texdir);
      AddBackground(s, m, texdir);
      AddRectangle(s, m, texdir, "mooni.jpeg", 100, 253, 199, 100, 0, 199, 100, 0, 0, 100, 253, 0);
      AddRectangle(s, m, texdir, "karina.jpeg", 100, 0, 300, 100, -200, 300, 100, -200, 0, 100, 0, 0);
      AddRectangle(s, m, texdir, "winter.jpg", -100, 50, 100, -100, -50, 100, -100, -50, 0, -100, 50, 0);
      AddRectangle(s, m, texdir, "mooni.jpeg", -400, 300, 350, -400, 100, 350, -400, 100, 150, -400, 300, 150);
      AddRectangle(s, m, texdir, "winter.jpg", 300, -300, 500, 300, -600, 500, 300, -600, 200, 300, -300, 200);
      AddRectangle(s, m, texdir, "karina.jpeg", -200, -400, 300, 0, -400, 300, 0, -400, 0, -200, -400, 0);
      AddChess(s, m, 25);
      AddBlackhole(s, m, 10);
      break;
    case 11:  // exactly four planes off the hole's centre (the kernel's largest filtered instantiation)
      s = new Scene(width, height, blackhole::pi / 2);
      s->camera.MoveTo(-1200, 200, 300);
      s->camera.RotateZ(0.15);
      AddDisc(s, m, texdir);
      AddBackground(s, m, texdir);
      AddRectangle(s, m, texdir, "mooni.jpeg", -400, 300, 350, -400, 100, 350, -400, 100, 150, -400, 300, 150);
      AddRectangle(s, m, texdir, "winter.jpg", 300, -300, 500, 300, -600, 500, 300, -600, 200, 300, -300, 200);
      AddRectangle(s, m, texdir, "karina.jpeg", -200, -400, 300, 0, -400, 300, 0, -400, 0, -200, -400, 0);
      AddBlackhole(s, m, 10);
      break;
    case 9:  // camera at r = 25.5 < 3M: du < 0, every segment takes the exact path
      s = new Scene(width, height, blackhole::pi / 2);
      s->camera.MoveTo(-25, 0, 5);
      AddDisc(s, m, texdir);
      AddBackground(s, m, texdir);
      AddBlackhole(s, m, 10);
      break;
    case 10: {  // ray_tracer_test.cc:45-98
      s = new Scene(width, height, blackhole::pi / 2);
      s->camera.MoveTo(-400, 120, 100);
      s->camera.RotateY(-blackhole::pi / 24);
      s->camera.RotateZ(blackhole::pi / 12);
      s->movers[0] = AddRectangle(s, m, texdir, "mooni.jpeg", 100, 120, 100, 100, 0, 100, 100, 0, 0, 100, 120, 0);
      s->movers[1] = AddRectangle(s, m, texdir, "karina.jpeg", 100, 0, 150, 100, -100, 150, 100, -100, 0, 100, 0, 0);
      s->movers[2] = AddRectangle(s, m, texdir, "winter.jpg", -100, 50, 100, -100, -50, 100, -100, -50, 0, -100, 50, 0);
      AddChess(s, m, 10);
      s->linear_steps = 10;  // Prograde(manager, dst, 10), ray_tracer_test.cc:145
      int karina_move_direction = 1;
      for (int k = 0; k < frame; ++k) {  // "Movement test", ray_tracer_test.cc:237-261
        s->movers[1]->MoveX(3 * karina_move_direction);
        if (auto x = s->movers[1]->position()[0]; x > 100)
          karina_move_direction = -1;
        else if (x < 0)
          karina_move_direction = 1;
        {
          const auto y = -120.0, z = -100.0;
          s->movers[0]->MoveY(-y / 2);
          s->movers[0]->MoveZ(-z / 2);
          s->movers[0]->RotateX(blackhole::pi / 180);
          s->movers[0]->MoveY(y / 2);
          s->movers[0]->MoveZ(z / 2);
        }
        {
          const auto x = 100.0, z = -100.0;
          s->movers[2]->MoveX(-x);
          s->movers[2]->MoveZ(-z / 2);
          s->movers[2]->RotateY(blackhole::pi / 120);
          s->movers[2]->MoveX(x);
          s->movers[2]->MoveZ(z / 2);
        }
      }
      break;
    }
    default:
      return nullptr;
  }
  // Replay `frame` iterations of the reference's frame loop tail (blackhole_solution_test.cc:346-407):
  // the key action for that frame, then the unconditional disc spin.
  if (cfg == 3 || ((cfg == 0 || cfg == 1) && frame > 0)) {
    const auto move_d = 10;
    const auto rotate_d = blackhole::pi / 1800.0;
    const auto faster = 10;
    for (int k = 0; k < frame; ++k) {
      if (cfg == 3) {
        if (k < 120) {
          s->camera.MoveX(move_d);  // 'w'
        } else {
          s->camera.RotateZ(rotate_d * faster);  // 'L'
          s->camera.MoveY(move_d);               // 'd'
        }
      }
      if (s->disc) s->disc->RotateZ(blackhole::pi / 180);
    }
  }
  return s;
}

// ---- snapshot: header-API state -> bh8 POD ------------------------------------------------

inline void Put3(double dst[3], const point_type& p) {
  dst[0] = p[0];
  dst[1] = p[1];
  dst[2] = p[2];
}

inline bh8_camera SnapshotCamera(const blackhole::Camera<value_type>& c) {
  bh8_camera out{};
  Put3(out.pos, c.focus());
  Put3(out.vx, c.vector_x());
  Put3(out.vy, c.vector_y());
  Put3(out.vz, c.vector_z());
  // focus_len_ is private (camera.h:68); focus_vector() = vector_x() * focus_len_ (camera.h:61-63)
  // and |vector_x()| = 1, so recover it by projection.
  const auto fv = c.focus_vector();
  out.focus_len = fv.dot(c.vector_x()) / c.vector_x().dot(c.vector_x());
  out.width = c.width();
  out.height = c.height();
  return out;
}

// ObjectManager iterates a std::unordered_map<int, ...> (object_manager.h:74,92).  With libstdc++
// and keys 0..n-1 inserted in order, each new key lands in an empty bucket and is linked at the
// list head, so iteration runs n-1..0.  The order only decides exact distance ties.
inline std::vector<bh8_object> SnapshotObjects(const Scene& s, int* bh_index,
                                               std::vector<std::string>* textures) {
  std::vector<bh8_object> out;
  *bh_index = -1;
  for (int i = static_cast<int>(s.objects.size()) - 1; i >= 0; --i) {
    const Inserted& e = s.objects[i];
    bh8_object o{};
    o.kind = e.kind;
    o.key = e.key;
    o.tex_id = -1;
    o.pattern = e.pattern;
    const auto& vtx = e.object->vertex();
    for (size_t k = 0; k < vtx.size() && k < 5; ++k) Put3(o.v[k], vtx[k]);
    if (!e.texture.empty()) {
      size_t t = 0;
      for (; t < textures->size(); ++t)
        if ((*textures)[t] == e.texture) break;
      if (t == textures->size()) textures->push_back(e.texture);
      o.tex_id = static_cast<int>(t);
    }
    switch (e.kind) {
      case BH8_KIND_BLACKHOLE:
        o.mass = static_cast<const blackhole::StaticBlackhole<value_type>*>(e.object)->mass();
        *bh_index = static_cast<int>(out.size());
        break;
      case BH8_KIND_ANNULUS:
        Put3(o.n, static_cast<const blackhole::Annulus<value_type>*>(e.object)->norm());
        o.r_in = e.r_in;
        o.r_out = e.r_out;
        break;
      case BH8_KIND_INFINITE_PLANE:
        Put3(o.n, e.object->vector_z());
        Put3(o.ex, e.object->vector_x());
        Put3(o.ey, e.object->vector_y());
        o.pattern_size = e.pattern_size;
        break;
      default:
        break;
    }
    out.push_back(o);
  }
  return out;
}

}  // namespace bh8scenes

#endif  // BH8_ORACLE_SCENES_H_
