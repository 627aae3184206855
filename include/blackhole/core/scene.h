// blackhole/core/scene.h -- ObjectManager.
//
// Implementation header of this repository's blackhole:: API.  The file names the reference uses
// (blackhole/camera.h, blackhole/object/vector_object.h, ...) are thin forwarding headers onto the
// blackhole/core/ set, so code written against the reference's include paths compiles unchanged.
//
// blackhole/object/object_manager.h -- owner of the scene's drawable objects and the nearest-hit
// query every ray step calls.
//
// Source-compatible with the reference's object/object_manager.h:21-95: a leaked singleton, four
// InsertObject overloads (key = number of objects so far), FindCollision(p1, p2, &hit).
// FindCollision tests the segment against EVERY object and returns the hit nearest to p1 (squared
// distance); among exact ties the object visited first wins.  Objects are visited in the iteration
// order of the std::unordered_map that stores them, as in the reference, so tie-breaking agrees.
// Additive API (the reference offers no way to enumerate a scene): size(), ForEach(), Find(), Clear().
#ifndef BLACKHOLE_CORE_SCENE_H_
#define BLACKHOLE_CORE_SCENE_H_

#include <memory>
#include <type_traits>
#include <unordered_map>
#include <utility>

#include "blackhole/core/scene_object.h"

namespace blackhole {

template <typename T>
class ObjectManager {
 public:
  using value_type = T;
  using object_type = DrawableObject<T>;
  using key_type = int;
  using mapped_type = std::unique_ptr<object_type>;
  using point_type = typename object_type::point_type;
  using vector_type = typename object_type::vector_type;

  static ObjectManager& GetInstance() {
    static ObjectManager* instance = new ObjectManager();  // never destroyed, like the reference's
    return *instance;
  }

  std::pair<const key_type, object_type*> InsertObject(std::unique_ptr<object_type> object) {
    return Adopt<object_type>(std::move(object));
  }

  template <typename U, std::enable_if_t<std::is_base_of_v<object_type, U>, int> = 0>
  std::pair<const key_type, U*> InsertObject(std::unique_ptr<U> object) {
    return Adopt<U>(std::move(object));
  }

  template <typename U, typename... Args,
            std::enable_if_t<std::is_base_of_v<object_type, U> && std::is_constructible_v<U, Args&&...>, int> = 0>
  std::pair<const key_type, U*> InsertObject(Args&&... args) {
    return InsertObject(std::make_unique<U>(std::forward<Args>(args)...));
  }

  template <template <typename> class U, typename... Args,
            std::enable_if_t<std::is_base_of_v<object_type, U<value_type>> &&
                                 std::is_constructible_v<U<value_type>, Args&&...>,
                             int> = 0>
  std::pair<const key_type, U<value_type>*> InsertObject(Args&&... args) {
    return InsertObject(std::make_unique<U<value_type>>(std::forward<Args>(args)...));
  }

  // The object hit first along p1 -> p2 (nullptr: none); *intersection receives the hit point.
  const object_type* FindCollision(const point_type& p1, const point_type& p2, point_type* intersection) const {
    const object_type* nearest = nullptr;
    value_type nearest_d2 = 0;
    point_type nearest_point, candidate;
    for (const auto& entry : objects_) {
      if (!entry.second->Collide(p1, p2, &candidate)) continue;
      const auto offset = p1 - candidate;
      const value_type d2 = offset.dot(offset);
      if (nearest == nullptr || d2 < nearest_d2) {
        nearest = entry.second.get();
        nearest_d2 = d2;
        nearest_point = candidate;
      }
    }
    if (nearest != nullptr) *intersection = nearest_point;
    return nearest;
  }

  // ---- additive API ---------------------------------------------------------------------------
  std::size_t size() const { return objects_.size(); }

  // fn(key, const object_type&) for every object, in FindCollision's visiting order.
  template <typename Fn>
  void ForEach(Fn&& fn) const {
    for (const auto& entry : objects_) fn(entry.first, *entry.second);
  }

  object_type* Find(key_type key) {
    const auto it = objects_.find(key);
    return it == objects_.end() ? nullptr : it->second.get();
  }

  void Clear() { objects_.clear(); }

 private:
  ObjectManager() = default;
  ObjectManager(const ObjectManager&) = delete;
  ObjectManager& operator=(const ObjectManager&) = delete;

  template <typename U, typename Ptr>
  std::pair<const key_type, U*> Adopt(Ptr object) {
    const auto placed = objects_.emplace(static_cast<key_type>(objects_.size()), std::move(object));
    return {placed.first->first, static_cast<U*>(placed.first->second.get())};
  }

  std::unordered_map<key_type, mapped_type> objects_;
};

}  // namespace blackhole

#endif  // BLACKHOLE_CORE_SCENE_H_
