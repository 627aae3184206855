// blackhole/gpu/snapshot.h -- header-API scene state -> the plain-old-data of include/bh8.h.
//
// The GPU path cannot call virtual Collide()/color() per step the way the reference's pixel loop
// does (blackhole_solution_test.cc:229-233); it takes a SNAPSHOT of everything that loop reads:
// the camera (camera.h), every object of the ObjectManager in FindCollision's visiting order
// (object_manager.h), and the textures (Material::texture_).  Unsupported content -- Triangle,
// Sphere, an InfinitePlane with an opaque std::function pattern -- is reported, never skipped.
#ifndef BLACKHOLE_GPU_SNAPSHOT_H_
#define BLACKHOLE_GPU_SNAPSHOT_H_

#include <stdexcept>
#include <string>
#include <vector>

#include "bh8.h"
#include "blackhole/blackhole_solution.h"
#include "blackhole/camera.h"
#include "blackhole/object.h"

namespace blackhole {
namespace gpu {

struct SceneSnapshot {
  std::vector<bh8_object> objects;   // ObjectManager visiting order
  std::vector<cv::Mat> textures;     // CV_8UC3 BGR, indexed by bh8_object::tex_id
  int bh_index = -1;

  bh8_scene view() const { return bh8_scene{static_cast<int32_t>(objects.size()), bh_index, objects.data()}; }
};

namespace detail {
template <typename V>
inline void Put3(double dst[3], const V& v) {
  dst[0] = v[0];
  dst[1] = v[1];
  dst[2] = v[2];
}
}  // namespace detail

template <typename T>
bh8_camera Snapshot(const Camera<T>& camera) {
  bh8_camera out{};
  detail::Put3(out.pos, camera.focus());
  detail::Put3(out.vx, camera.vector_x());
  detail::Put3(out.vy, camera.vector_y());
  detail::Put3(out.vz, camera.vector_z());
  out.focus_len = camera.focus_len();
  out.width = camera.width();
  out.height = camera.height();
  return out;
}

// `blackhole` names the hole the geodesics bend around (the reference's drivers keep that reference
// next to the manager, blackhole_solution_test.cc:148).  Textures are shared, not copied
// (cv::Mat reference counting); slots are assigned per distinct pixel buffer.
template <typename T>
SceneSnapshot Snapshot(const ObjectManager<T>& manager, const StaticBlackhole<T>* blackhole_ptr) {
  SceneSnapshot snap;
  if (manager.size() > BH8_MAX_OBJECTS) throw std::runtime_error("scene has more than BH8_MAX_OBJECTS objects");
  manager.ForEach([&](int key, const DrawableObject<T>& obj) {
    bh8_object o{};
    o.key = key;
    o.tex_id = -1;
    o.pattern = BH8_PATTERN_BLACK;
    const auto& vtx = obj.vertex();
    for (size_t k = 0; k < vtx.size() && k < 5; ++k) detail::Put3(o.v[k], vtx[k]);
    const auto bind_texture = [&]() {
      if (obj.texture_.empty()) return;
      if (obj.texture_.type() != CV_8UC3) throw std::runtime_error("texture is not CV_8UC3");
      size_t slot = 0;
      while (slot < snap.textures.size() && snap.textures[slot].data != obj.texture_.data) ++slot;
      if (slot == snap.textures.size()) {
        if (slot >= BH8_MAX_TEXTURES) throw std::runtime_error("scene uses more than BH8_MAX_TEXTURES textures");
        snap.textures.push_back(obj.texture_);
      }
      o.tex_id = static_cast<int32_t>(slot);
    };
    switch (obj.kind()) {
      case ShapeKind::kBlackhole:
        o.kind = BH8_KIND_BLACKHOLE;
        o.mass = static_cast<const StaticBlackhole<T>&>(obj).mass();
        if (&obj == blackhole_ptr) snap.bh_index = static_cast<int>(snap.objects.size());
        break;
      case ShapeKind::kAnnulus: {
        const auto& disc = static_cast<const Annulus<T>&>(obj);
        o.kind = BH8_KIND_ANNULUS;
        detail::Put3(o.n, disc.norm());
        o.r_in = disc.r_inner();
        o.r_out = disc.r_outer();
        bind_texture();
        break;
      }
      case ShapeKind::kRectangle:
        o.kind = BH8_KIND_RECTANGLE;
        bind_texture();
        break;
      case ShapeKind::kInfinitePlane: {
        const auto& plane = static_cast<const InfinitePlane<T>&>(obj);
        o.kind = BH8_KIND_INFINITE_PLANE;
        detail::Put3(o.n, plane.vector_z());
        detail::Put3(o.ex, plane.vector_x());
        detail::Put3(o.ey, plane.vector_y());
        switch (plane.pattern_kind()) {
          case InfinitePlane<T>::PatternKind::kBlack: o.pattern = BH8_PATTERN_BLACK; break;
          case InfinitePlane<T>::PatternKind::kChess:
            o.pattern = BH8_PATTERN_CHESS;
            o.pattern_size = plane.pattern_size();
            break;
          default:
            throw std::runtime_error("InfinitePlane with an opaque std::function pattern cannot run on the GPU");
        }
        break;
      }
      default:
        throw std::runtime_error("object kind not supported by the GPU path (Triangle / Sphere / unknown)");
    }
    snap.objects.push_back(o);
  });
  if (blackhole_ptr && snap.bh_index < 0)
    throw std::runtime_error("the black hole is not an object of this ObjectManager");
  return snap;
}

template <typename T>
SceneSnapshot Snapshot(const ObjectManager<T>& manager, const StaticBlackhole<T>& blackhole) {
  return Snapshot(manager, &blackhole);
}

// A flat-space scene (ray_tracer_test.cc: no hole the rays bend around; bh_index = -1).
template <typename T>
SceneSnapshot SnapshotFlat(const ObjectManager<T>& manager) {
  return Snapshot(manager, static_cast<const StaticBlackhole<T>*>(nullptr));
}

}  // namespace gpu
}  // namespace blackhole

#endif  // BLACKHOLE_GPU_SNAPSHOT_H_
