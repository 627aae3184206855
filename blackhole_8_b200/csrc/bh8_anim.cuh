// bh8_anim.cuh -- scene / camera animation replayed on the device.
//
// The reference animates between frames on the host: key handlers call Camera::MoveX/RotateZ/...
// and the disc spins by Annulus::RotateZ(pi/180) after every frame (blackhole_solution_test.cc:346-407),
// all of which are Object's members (object/object.h:58-88) over RotationMatrixForAxis
// (matrix.h:342-356) and OpenCV's small-vector arithmetic.  For a scripted fly-through the same
// updates run here, one thread per entity (camera or object), sequentially over the frames, and
// leave one bh8_camera + bh8_object[] snapshot per frame in device memory; bh8_build_frames_kernel
// then derives every frame's constants (Bh8Frame) next to them, so rendering frame k needs nothing
// from the host any more.
//
// Bit-exactness.  Every operation is a single correctly rounded +, -, *, /, sqrt in the reference's
// order (BH8F_* of bh8_frame.h: round-to-nearest intrinsics on the device, which are never contracted
// into FMAs); cos/sin of the rotation angles come from the host's libm with the action (Bh8Action::c,
// ::s), as the reference computes them.  tests/test_script_host.py replays BASELINE configs[3]'s 240
// frames through this code on the CPU and tests/test_gpu_script.py on the GPU: identical to the
// states the reference's own classes produced (tests/golden/states/).
#ifndef BH8_ANIM_CUH_
#define BH8_ANIM_CUH_

#include "bh8_frame.h"

struct Bh8Action {   // device form of bh8_action
  int32_t frame, target, op, reserved;
  double amount;
  double c, s;       // cos(amount), sin(amount): std::cos / std::sin of matrix.h:349,351
  double to[3];
};

struct Bh8Entity {   // an Object (object.h:109-115): vertex_ and the basis vx_, vy_, vz_
  int32_t nv, reserved;
  double v[5][3];
  double vx[3], vy[3], vz[3];
};

// cv::normalize: v * (n ? 1./n : 0.), n = sqrt(sum v_i^2) summed left to right from 0
BH8F_HD void bh8a_normalize(const double* v, double* out) {
  double s = 0.0;
  for (int i = 0; i < 3; ++i) s = BH8F_ADD(s, BH8F_MUL(v[i], v[i]));
  const double n = BH8F_SQRT(s);
  const double k = n ? BH8F_DIV(1., n) : 0.;
  for (int i = 0; i < 3; ++i) out[i] = BH8F_MUL(v[i], k);
}

// cv::Matx33 * cv::Vec3: row-wise, s = 0; s += a(i,k) * b[k]
BH8F_HD void bh8a_matvec(const double* r, const double* v, double* out) {
  for (int i = 0; i < 3; ++i) {
    double s = 0.0;
    for (int k = 0; k < 3; ++k) s = BH8F_ADD(s, BH8F_MUL(r[3 * i + k], v[k]));
    out[i] = s;
  }
}

// RotationMatrixForAxis, matrix.h:342-356 (row-major)
BH8F_HD void bh8a_rotation(const double* axis, double c, double s, double* r) {
  double u[3];
  bh8a_normalize(axis, u);
  const double ux = u[0], uy = u[1], uz = u[2];
  const double c2 = BH8F_SUB(1.0, c);
  r[0] = BH8F_ADD(c, BH8F_MUL(BH8F_MUL(ux, ux), c2));
  r[1] = BH8F_SUB(BH8F_MUL(BH8F_MUL(ux, uy), c2), BH8F_MUL(uz, s));
  r[2] = BH8F_ADD(BH8F_MUL(BH8F_MUL(ux, uz), c2), BH8F_MUL(uy, s));
  r[3] = BH8F_ADD(BH8F_MUL(BH8F_MUL(uy, ux), c2), BH8F_MUL(uz, s));
  r[4] = BH8F_ADD(c, BH8F_MUL(BH8F_MUL(uy, uy), c2));
  r[5] = BH8F_SUB(BH8F_MUL(BH8F_MUL(uy, uz), c2), BH8F_MUL(ux, s));
  r[6] = BH8F_SUB(BH8F_MUL(BH8F_MUL(uz, ux), c2), BH8F_MUL(uy, s));
  r[7] = BH8F_ADD(BH8F_MUL(BH8F_MUL(uz, uy), c2), BH8F_MUL(ux, s));
  r[8] = BH8F_ADD(c, BH8F_MUL(BH8F_MUL(uz, uz), c2));
}

// Object::RotateX/Y/Z, object.h:58-80: the two other basis vectors are rotated and re-normalised,
// every vertex turns about position() = vertex_[0] (read afresh for every vertex, as the loop does).
BH8F_HD void bh8a_rotate(Bh8Entity* e, double* axis, double* b1, double* b2, double c, double s) {
  double r[9], t[3];
  bh8a_rotation(axis, c, s, r);
  bh8a_matvec(r, b1, t);
  bh8a_normalize(t, b1);
  bh8a_matvec(r, b2, t);
  bh8a_normalize(t, b2);
  for (int k = 0; k < e->nv; ++k) {
    double d[3];
    for (int i = 0; i < 3; ++i) d[i] = BH8F_SUB(e->v[k][i], e->v[0][i]);
    bh8a_matvec(r, d, t);
    for (int i = 0; i < 3; ++i) e->v[k][i] = BH8F_ADD(t[i], e->v[0][i]);
  }
}

// Object::MoveX/Y/Z, object.h:86-88: v += basis * distance for every vertex
BH8F_HD void bh8a_move(Bh8Entity* e, const double* dir, double distance) {
  double step[3];
  for (int i = 0; i < 3; ++i) step[i] = BH8F_MUL(dir[i], distance);
  for (int k = 0; k < e->nv; ++k)
    for (int i = 0; i < 3; ++i) e->v[k][i] = BH8F_ADD(e->v[k][i], step[i]);
}

BH8F_HD void bh8a_apply(Bh8Entity* e, const Bh8Action* a) {
  switch (a->op) {
    case BH8_OP_MOVE_X: bh8a_move(e, e->vx, a->amount); break;
    case BH8_OP_MOVE_Y: bh8a_move(e, e->vy, a->amount); break;
    case BH8_OP_MOVE_Z: bh8a_move(e, e->vz, a->amount); break;
    case BH8_OP_ROTATE_X: bh8a_rotate(e, e->vx, e->vy, e->vz, a->c, a->s); break;  // :74-80
    case BH8_OP_ROTATE_Y: bh8a_rotate(e, e->vy, e->vx, e->vz, a->c, a->s); break;  // :66-72
    case BH8_OP_ROTATE_Z: bh8a_rotate(e, e->vz, e->vx, e->vy, a->c, a->s); break;  // :58-64
    case BH8_OP_MOVE_TO:                                                           // :82-83: vertex_[0] only
      for (int i = 0; i < 3; ++i) e->v[0][i] = a->to[i];
      break;
    default: break;
  }
}

// What blackhole::gpu::Snapshot() would read from the live object (include/blackhole/gpu/snapshot.h).
BH8F_HD void bh8a_emit_object(const Bh8Entity* e, const bh8_object* proto, bh8_object* out) {
  *out = *proto;
  for (int k = 0; k < e->nv && k < 5; ++k)
    for (int i = 0; i < 3; ++i) out->v[k][i] = e->v[k][i];
  if (proto->kind == BH8_KIND_INFINITE_PLANE)
    for (int i = 0; i < 3; ++i) {
      out->n[i] = e->vz[i];   // vector_z(), vector_object.h:211
      out->ex[i] = e->vx[i];
      out->ey[i] = e->vy[i];
    }
  // Annulus::norm_ is fixed at construction and NOT rotated (vector_object.h:325): proto's stays.
}

BH8F_HD void bh8a_emit_camera(const Bh8Entity* e, const bh8_camera* proto, bh8_camera* out) {
  *out = *proto;
  for (int i = 0; i < 3; ++i) {
    out->pos[i] = e->v[0][i];
    out->vx[i] = e->vx[i];
    out->vy[i] = e->vy[i];
    out->vz[i] = e->vz[i];
  }
}

// One entity through all frames: snapshot frame k, then apply the actions of frame k that target it
// (the reference draws first and handles keys / spins the disc afterwards).  `actions` are sorted by
// frame.  target: index into the scene's objects, or n_obj for the camera.
BH8F_HD void bh8a_replay_entity(Bh8Entity e, int me, int n_obj, const Bh8Action* actions, int n_actions,
                                int n_frames, const bh8_object* proto_obj, const bh8_camera* proto_cam,
                                bh8_camera* cams, bh8_object* objs) {
  int cursor = 0;
  for (int k = 0; k < n_frames; ++k) {
    if (me == n_obj)
      bh8a_emit_camera(&e, proto_cam, &cams[k]);
    else
      bh8a_emit_object(&e, &proto_obj[me], &objs[(size_t)k * n_obj + me]);
    for (; cursor < n_actions && actions[cursor].frame <= k; ++cursor) {
      const Bh8Action* a = &actions[cursor];
      const int target = a->target == BH8_TARGET_CAMERA ? n_obj : a->target;
      if (a->frame == k && target == me) bh8a_apply(&e, a);
    }
  }
}

#if !defined(__CUDACC__) || defined(BH8_HOST_BUILD)
// ---- host side: bh8_action / scene0 -> the device forms -------------------------------------------
#include <math.h>

// Returns NULL or the reason the script is malformed.  out: n_actions entries.
static inline const char* bh8a_prepare_actions(const bh8_action* actions, int n_actions, int n_obj, Bh8Action* out) {
  for (int i = 0; i < n_actions; ++i) {
    const bh8_action* a = &actions[i];
    if (a->op < BH8_OP_MOVE_X || a->op > BH8_OP_MOVE_TO) return "bh8_action: unknown op";
    if (a->target != BH8_TARGET_CAMERA && (a->target < 0 || a->target >= n_obj))
      return "bh8_action: target is neither the camera nor an object of the scene";
    if (a->frame < 0 || (i > 0 && a->frame < actions[i - 1].frame))
      return "bh8_action: actions must be sorted by frame (>= 0)";
    Bh8Action* d = &out[i];
    d->frame = a->frame;
    d->target = a->target;
    d->op = a->op;
    d->reserved = 0;
    d->amount = a->amount;
    d->c = cos(a->amount);  // matrix.h:349,351: the host's libm, as the reference
    d->s = sin(a->amount);
    for (int k = 0; k < 3; ++k) d->to[k] = a->to[k];
  }
  return 0;
}

// out: n_obj + 1 entities, the camera last.
static inline void bh8a_prepare_entities(const bh8_scene* scene0, const bh8_basis* obj_basis, const bh8_camera* cam0,
                                         Bh8Entity* out) {
  const int n_obj = scene0->n_obj;
  for (int j = 0; j <= n_obj; ++j) {
    Bh8Entity* e = &out[j];
    memset(e, 0, sizeof *e);
    if (j == n_obj) {  // the camera: Object(p, vx, vy, vz) with one vertex, camera.h:30-35
      e->nv = 1;
      for (int k = 0; k < 3; ++k) {
        e->v[0][k] = cam0->pos[k];
        e->vx[k] = cam0->vx[k];
        e->vy[k] = cam0->vy[k];
        e->vz[k] = cam0->vz[k];
      }
      continue;
    }
    const bh8_object* o = &scene0->obj[j];
    // vertex_: the position plus, for the shapes built from corner points, four corners (object.h:38-40,109)
    e->nv = (o->kind == BH8_KIND_RECTANGLE || o->kind == BH8_KIND_ANNULUS) ? 5 : 1;
    for (int v = 0; v < 5; ++v)
      for (int k = 0; k < 3; ++k) e->v[v][k] = o->v[v][k];
    for (int k = 0; k < 3; ++k) {
      if (obj_basis) {
        e->vx[k] = obj_basis[j].vx[k];
        e->vy[k] = obj_basis[j].vy[k];
        e->vz[k] = obj_basis[j].vz[k];
      } else if (o->kind == BH8_KIND_INFINITE_PLANE) {
        e->vx[k] = o->ex[k];
        e->vy[k] = o->ey[k];
        e->vz[k] = o->n[k];
      } else {
        e->vx[k] = k == 0;  // object.h:112-114
        e->vy[k] = k == 1;
        e->vz[k] = k == 2;
      }
    }
  }
}
#endif  // host

#if defined(__CUDACC__)
namespace bh8 {

struct Bh8TexSizes {
  int rows[BH8_MAX_TEXTURES], cols[BH8_MAX_TEXTURES];
};

// <<<1, 32>>>: thread t < n_obj replays object t, thread n_obj the camera (n_obj <= 16).
__global__ void bh8_animate_kernel(const Bh8Entity* entities, const Bh8Action* actions, int n_actions, int n_obj,
                                   int n_frames, const bh8_object* proto_obj, const bh8_camera proto_cam,
                                   bh8_camera* cams, bh8_object* objs) {
  const int me = threadIdx.x;
  if (me > n_obj) return;
  bh8a_replay_entity(entities[me], me, n_obj, actions, n_actions, n_frames, proto_obj, &proto_cam, cams, objs);
}

// One thread per frame: snapshot k -> frames[k] (exactly what the host's bh8_build_frame gives).
__global__ void bh8_build_frames_kernel(const bh8_camera* cams, const bh8_object* objs, int n_obj, int bh_index,
                                        const bh8_params prm, const Bh8TexSizes tex, int resolve_wait,
                                        Bh8Frame* frames, int* status, int n_frames) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n_frames) return;
  bh8_scene scene;
  scene.n_obj = n_obj;
  scene.bh_index = bh_index;
  scene.obj = objs + (size_t)k * n_obj;
  const int reason = bh8_build_frame_core(&scene, &cams[k], &prm, tex.rows, tex.cols, &frames[k]);
  if (reason == BH8F_OK) {
    if (resolve_wait != 0x7fffffff) frames[k].resolve_wait = resolve_wait;
    if (prm.flags & BH8_FLAG_NO_BATCHING) frames[k].resolve_wait = -1;
  }
  status[2 * k] = reason;
  status[2 * k + 1] = reason == BH8F_OK ? frames[k].n_nc : 0;
}

}  // namespace bh8
#endif  // __CUDACC__

#endif  // BH8_ANIM_CUH_
