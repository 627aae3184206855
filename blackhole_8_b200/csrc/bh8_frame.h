// bh8_frame.h -- per-frame constants of the geodesic render kernel and the host code that
// derives them from a bh8_scene / bh8_camera snapshot.
//
// Everything a ray needs that does not depend on the pixel is hoisted here once per frame on the
// host (the reference recomputes all of it per pixel or per step: F at blackhole_solution_test.cc:168,
// Rectangle's normal at vector_object.h:108-111, 1/(b*b) at blackhole_solution.h:28, ...).  The struct
// travels to the GPU as a __grid_constant__ kernel parameter, so its fields are constant-bank
// operands of the FP64 instructions: no loads in the stepping loop.
#ifndef BH8_FRAME_H_
#define BH8_FRAME_H_

#include <math.h>
#include <stdint.h>
#include <string.h>

#include "bh8.h"

#define BH8_BISECT_ITERS 20  /* static const int count = 20, blackhole_solution.h:36 */

struct Bh8Obj {
  int32_t kind, key, cls, tex;  // tex = texture slot or -1
  int32_t pattern, central, tex_rows, tex_cols;
  double n[3];    // unit normal Collide() uses: Annulus norm_; Rectangle normalize(e1 x e3); plane vector_z
  double p0[3];   // point of the plane: Annulus center() / Rectangle vertex()[1] / plane position(); BH centre
  double d;       // n . p0
  double c_bh;    // n . (bh_pos - p0): signed distance of the black hole from the plane (0 => central)
  double nF;      // n . Fhat (Fhat = F/|F|): the plane normal's component along the camera direction
  double e1[3], e3[3], e1e1, e3e3;  // Rectangle edges vertex()[2]-[1], vertex()[4]-[1]
  double r_in, r_out;               // Annulus radii
  double t0[3], s1[3];              // texture frame (Rectangle::color): origin vertex()[1], s1 = [2]-[1]
  double inv_s1s1;                  // 1 / (s1 . s1)
  double kw, kh;                    // cols / (s1 . s1),  rows / |s2|
  double ex0, ex1, ey0, ey1, psize; // InfinitePlane::color coefficients, ChessPattern2D size
};

struct Bh8Frame {
  // camera (camera.h:55-63)
  double cam[3], fv[3], vy[3], vz[3];
  double half_w, half_h;  // width/2.0, height/2.0
  int32_t width, height;
  // black hole (blackhole_solution.h:24-57)
  double bh[3], F[3];     // F = camera.focus() - blackhole.position()
  double Fhat[3];         // F / |F|
  double FF;              // F . F
  double mass, two_m, b_c2 /* b_c^2 */, inv3m, R, R2;
  double r0, u0;          // |F| and 1/|F|: convertedCameraFocus has the length of F
  // SolveG's bisection (blackhole_solution.h:35-53) on [l0, r0] = [cbrt(eps), 1/(3M)]: after test i
  // (0-based) the midpoint moves by bis_h[i] = (r0 - l0) / 2^(i+2); the result lies on the grid
  // l0 + K * bis_grid, bis_grid = (r0 - l0) / 2^20.
  double bis_mid0;
  double bis_h[BH8_BISECT_ITERS + 1];
  double bis_l0, bis_grid, bis_inv_grid;
  double nine_m2;         // (3M)^2: q = nine_m2 / b^2 is the constant term of the normalised cubic
  // integration
  int32_t nstep, n_obj, bh_index, n_central;
  int32_t tracer, linear_steps;  // BH8_TRACER_*; segments per ray of the linear tracer
  // step indices at which a ray's leg changes (lane_event): the 0.9-step (nstep - 1), the first outbound
  // step / captured chord (nstep) and the end of the ray (2 nstep - 1)
  int32_t evt_turn, evt_back, evt_end, evt_pad;
  double inv_nstep;
  // conservative filters (see bh8_ray.cuh)
  double u_gate;        // non-central planes can only be crossed while min(u) <= u_gate
  double u_horizon;     // horizon sphere cannot be reached while u <= u_horizon and |dphi| <= 1
  uint32_t noncentral_mask;  // objects handled by the side-sign filter
  uint32_t central_mask;     // objects handled by the phi-crossing filter
  uint32_t hole_mask;        // BH8_KIND_BLACKHOLE objects (the horizon sphere)
  int32_t central_obj0;      // index of the first central plane (the only one when n_central == 1)
  int32_t first_resolve;     // 1: always resolve the first segment exactly (camera too close / in a plane)
  int32_t n_nc;              // number of non-central planes (<= BH8_MAX_OBJECTS)
  int32_t nc_obj[BH8_MAX_OBJECTS];  // their object indices, in scene order
  float nc_nF[BH8_MAX_OBJECTS];     // float(n . Fhat)
  float nc_c[BH8_MAX_OBJECTS];      // float(c_bh)
  // leases of filter (2), see lane_update: 0.4995 / max|n_j| and 0.24975 / max|c_bh_j| over the
  // non-central planes (the margin is split between the phi and the u movement, 0.1 % kept back)
  float lease_kphi, lease_ku;
  float nc_max_n, nc_max_c;  // max|n_j|, max|c_bh_j| themselves (outbound leases use the whole margin, lane_update_rare)
  uint32_t nc_cam_bits;      // side of the camera w.r.t. non-central plane j: bit j = positive, bit 16+j = negative
  int32_t resolve_wait;      // warp iterations a pending exact test may wait for company (batching window)
  // output
  int32_t pixel_format, stripe_rows, shard_index, shard_count;
  uint32_t flags;
  Bh8Obj obj[BH8_MAX_OBJECTS];
};

// ---- snapshot -> Bh8Frame (host, and device for scripted animations: SURVEY.md 8f-3) ---------
//
// One definition for both sides.  On the device every operation goes through the round-to-nearest
// intrinsics, which nvcc never contracts into FMAs, so a frame built by bh8_build_frames_kernel is
// bit-identical to the one the host builds from the same snapshot (tests/test_gpu_script.py).
#if defined(__CUDA_ARCH__)
#define BH8F_HD __host__ __device__ inline
#define BH8F_MUL(a, b) __dmul_rn((a), (b))
#define BH8F_ADD(a, b) __dadd_rn((a), (b))
#define BH8F_SUB(a, b) __dsub_rn((a), (b))
#define BH8F_DIV(a, b) __ddiv_rn((a), (b))
#define BH8F_SQRT(a) __dsqrt_rn(a)
#else
#if defined(__CUDACC__)
#define BH8F_HD __host__ __device__ inline
#else
#define BH8F_HD static inline
#endif
#define BH8F_MUL(a, b) ((a) * (b))
#define BH8F_ADD(a, b) ((a) + (b))
#define BH8F_SUB(a, b) ((a) - (b))
#define BH8F_DIV(a, b) ((a) / (b))
#define BH8F_SQRT(a) sqrt(a)
#endif

// cbrt(DBL_EPSILON) as glibc returns it (utility.h:17-21 epsilon<double>(); checked by tests/test_abi.py):
// a literal, so that host and device start SolveG's interval from the same bits.
#define BH8_CBRT_DBL_EPSILON 0x1.965fea53d6e3dp-18

BH8F_HD double bh8h_dot(const double* a, const double* b) {
  return BH8F_ADD(BH8F_ADD(BH8F_MUL(a[0], b[0]), BH8F_MUL(a[1], b[1])), BH8F_MUL(a[2], b[2]));
}
BH8F_HD void bh8h_sub(double* r, const double* a, const double* b) {
  for (int i = 0; i < 3; ++i) r[i] = BH8F_SUB(a[i], b[i]);
}
BH8F_HD void bh8h_cross(double* r, const double* a, const double* b) {
  r[0] = BH8F_SUB(BH8F_MUL(a[1], b[2]), BH8F_MUL(a[2], b[1]));
  r[1] = BH8F_SUB(BH8F_MUL(a[2], b[0]), BH8F_MUL(a[0], b[2]));
  r[2] = BH8F_SUB(BH8F_MUL(a[0], b[1]), BH8F_MUL(a[1], b[0]));
}
BH8F_HD void bh8h_normalize(double* v) {  // cv::normalize semantics: zero stays zero
  const double n = BH8F_SQRT(bh8h_dot(v, v));
  const double s = n ? BH8F_DIV(1., n) : 0.;
  for (int i = 0; i < 3; ++i) v[i] = BH8F_MUL(v[i], s);
}

// Failure reasons of bh8_build_frame_core (index into bh8_frame_error_text / code).
enum {
  BH8F_OK = 0,
  BH8F_NULL,
  BH8F_NOBJ,
  BH8F_TRACER,
  BH8F_LINEAR_STEPS,
  BH8F_BH_INDEX,
  BH8F_CAMERA_SIZE,
  BH8F_NSTEP,
  BH8F_PIXEL_FORMAT,
  BH8F_MASS,
  BH8F_SHARDING,
  BH8F_TWO_HOLES,
  BH8F_PATTERN,
  BH8F_CHESS_SIZE,
  BH8F_KIND,
  BH8F_TEXTURE,
  BH8F_ZERO_NORMAL,
  BH8F_N_ERRORS
};

// tex_rows/tex_cols: sizes of the texture slots (0 = slot empty).  Returns BH8F_OK or a BH8F_* reason.
BH8F_HD int bh8_build_frame_core(const bh8_scene* scene, const bh8_camera* cam, const bh8_params* prm,
                                 const int* tex_rows, const int* tex_cols, Bh8Frame* f) {
  if (!scene || !cam || !prm || !scene->obj) return BH8F_NULL;
  if (scene->n_obj < 1 || scene->n_obj > BH8_MAX_OBJECTS) return BH8F_NOBJ;
  const bool linear = prm->tracer == BH8_TRACER_LINEAR;
  if (prm->tracer != BH8_TRACER_GEODESIC && !linear) return BH8F_TRACER;
  if (linear && (prm->linear_steps < 1 || prm->linear_steps > 65535)) return BH8F_LINEAR_STEPS;
  // The geodesic tracer bends rays around obj[bh_index]; the linear one needs no hole (bh_index may
  // be -1), and a StaticBlackhole in a flat-space scene is just a black sphere to FindCollision.
  if (!(linear && scene->bh_index == -1) &&
      (scene->bh_index < 0 || scene->bh_index >= scene->n_obj ||
       scene->obj[scene->bh_index].kind != BH8_KIND_BLACKHOLE))
    return BH8F_BH_INDEX;
  if (cam->width < 1 || cam->height < 1 || cam->width > 65536 || cam->height > 65536) return BH8F_CAMERA_SIZE;
  if (!linear && (prm->nstep < 2 || prm->nstep > 32767)) return BH8F_NSTEP;
  if (prm->pixel_format < 0 || prm->pixel_format > BH8_PIXEL_BGR8) return BH8F_PIXEL_FORMAT;
  // The linear tracer bends nothing: bh_index may be -1.  A StaticBlackhole among the objects is then just
  // a black sphere of radius 2M to FindCollision (one such sphere: the kernels hold one radius); with no
  // hole at all a unit-mass stand-in at the origin feeds the (unused) geodesic constants.
  const double no_hole_pos[3] = {0.0, 0.0, 0.0};
  int hole = scene->bh_index;
  if (linear && hole < 0)
    for (int k = 0; k < scene->n_obj; ++k)
      if (scene->obj[k].kind == BH8_KIND_BLACKHOLE) {
        if (hole >= 0) return BH8F_TWO_HOLES;
        hole = k;
      }
  const bh8_object* bho = hole >= 0 ? &scene->obj[hole] : nullptr;
  const double bh_mass = bho ? bho->mass : 1.0;
  const double* bh_pos = bho ? bho->v[0] : no_hole_pos;
  if (!(bh_mass > 0)) return BH8F_MASS;

  memset(f, 0, sizeof *f);
  for (int i = 0; i < 3; ++i) {
    f->cam[i] = cam->pos[i];
    f->fv[i] = BH8F_MUL(cam->vx[i], cam->focus_len);  // Camera::focus_vector, camera.h:61-63
    f->vy[i] = cam->vy[i];
    f->vz[i] = cam->vz[i];
    f->bh[i] = bh_pos[i];
  }
  f->half_w = BH8F_DIV((double)cam->width, 2.0);
  f->half_h = BH8F_DIV((double)cam->height, 2.0);
  f->width = cam->width;
  f->height = cam->height;
  bh8h_sub(f->F, f->cam, f->bh);
  f->FF = bh8h_dot(f->F, f->F);
  for (int i = 0; i < 3; ++i) f->Fhat[i] = BH8F_DIV(f->F[i], BH8F_SQRT(f->FF));
  f->resolve_wait = 2;
  f->mass = bh_mass;
  f->two_m = BH8F_MUL(2.0, bh_mass);
  const double b_c = BH8F_MUL(BH8F_MUL(3.0, 1.7320508075688772 /* sqrt(3.0) */), bh_mass);  // blackhole_solution.h:25
  f->b_c2 = BH8F_MUL(b_c, b_c);
  f->inv3m = BH8F_DIV(1.0, BH8F_MUL(3.0, bh_mass));
  f->R = BH8F_MUL(2.0, bh_mass);  // radius(), blackhole_solution.h:57
  f->R2 = BH8F_MUL(f->R, f->R);
  f->r0 = BH8F_SQRT(f->FF);
  f->u0 = BH8F_DIV(1., f->r0);
  {  // SolveG's interval, blackhole_solution.h:37-38; utility.h:17-21
    const double l = BH8F_ADD(0.0, BH8_CBRT_DBL_EPSILON);
    const double r = f->inv3m;
    f->bis_mid0 = BH8F_DIV(BH8F_ADD(l, r), 2.0);
    f->bis_l0 = l;
    f->bis_grid = BH8F_DIV(BH8F_SUB(r, l), 1048576.0);
    f->bis_inv_grid = BH8F_DIV(1048576.0, BH8F_SUB(r, l));
    f->nine_m2 = BH8F_MUL(BH8F_MUL(9.0, bh_mass), bh_mass);
    double h = BH8F_DIV(BH8F_SUB(r, l), 2.0);
    for (int i = 0; i <= BH8_BISECT_ITERS; ++i) {
      h = BH8F_DIV(h, 2.0);
      f->bis_h[i] = h;  // bis_h[i] = (r-l)/2^(i+2)
    }
  }
  f->tracer = prm->tracer;
  f->linear_steps = prm->linear_steps;
  f->nstep = linear ? 2 : prm->nstep;
  f->inv_nstep = BH8F_DIV(1.0, (double)f->nstep);
  f->evt_turn = f->nstep - 1;
  f->evt_back = f->nstep;
  f->evt_end = 2 * f->nstep - 1;
  f->n_obj = scene->n_obj;
  f->bh_index = scene->bh_index;
  f->pixel_format = prm->pixel_format;
  f->stripe_rows = prm->stripe_rows;
  f->shard_index = prm->shard_index;
  f->shard_count = prm->shard_count;
  f->flags = prm->flags;
  if (f->shard_count > 1 && (f->shard_index < 0 || f->shard_index >= f->shard_count || f->stripe_rows < 1))
    return BH8F_SHARDING;

  double min_dist = INFINITY, max_n = 0.0, max_c = 0.0;
  for (int k = 0; k < scene->n_obj; ++k) {
    const bh8_object* o = &scene->obj[k];
    Bh8Obj* q = &f->obj[k];
    q->kind = o->kind;
    q->key = o->key;
    q->tex = -1;
    q->pattern = o->pattern;
    switch (o->kind) {
      case BH8_KIND_BLACKHOLE:
        q->cls = BH8_CLASS_HORIZON;
        f->hole_mask |= 1u << k;
        if (k != hole) return BH8F_TWO_HOLES;
        for (int i = 0; i < 3; ++i) q->p0[i] = o->v[0][i];
        continue;
      case BH8_KIND_ANNULUS:
        q->cls = BH8_CLASS_DISC;
        for (int i = 0; i < 3; ++i) {
          q->n[i] = o->n[i];          // norm_, vector_object.h:325 (not re-normalised here: as stored)
          q->p0[i] = o->v[0][i];      // center(), :349
        }
        q->r_in = o->r_in;
        q->r_out = o->r_out;
        break;
      case BH8_KIND_RECTANGLE:
        q->cls = BH8_CLASS_OBJECT;
        for (int i = 0; i < 3; ++i) q->p0[i] = o->v[1][i];  // vector_object.h:108
        bh8h_sub(q->e1, o->v[2], o->v[1]);
        bh8h_sub(q->e3, o->v[4], o->v[1]);
        bh8h_cross(q->n, q->e1, q->e3);
        bh8h_normalize(q->n);  // :111
        q->e1e1 = bh8h_dot(q->e1, q->e1);
        q->e3e3 = bh8h_dot(q->e3, q->e3);
        break;
      case BH8_KIND_INFINITE_PLANE:
        q->cls = BH8_CLASS_OBJECT;
        if (o->pattern != BH8_PATTERN_BLACK && o->pattern != BH8_PATTERN_CHESS) return BH8F_PATTERN;
        for (int i = 0; i < 3; ++i) {
          q->n[i] = o->n[i];      // vector_z(), vector_object.h:211
          q->p0[i] = o->v[0][i];  // position()
        }
        q->ex0 = o->ex[0];
        q->ex1 = o->ex[1];
        q->ey0 = o->ey[0];
        q->ey1 = o->ey[1];
        q->psize = o->pattern_size;
        if (o->pattern == BH8_PATTERN_CHESS && !((int)BH8F_MUL(o->pattern_size, 2.0) != 0)) return BH8F_CHESS_SIZE;
        break;
      default:
        return BH8F_KIND;
    }
    q->d = bh8h_dot(q->n, q->p0);
    {
      double w[3];
      bh8h_sub(w, f->bh, q->p0);
      q->c_bh = bh8h_dot(q->n, w);
    }
    q->central = (q->c_bh == 0.0);
    q->nF = bh8h_dot(q->n, f->Fhat);
    if (o->kind == BH8_KIND_ANNULUS || o->kind == BH8_KIND_RECTANGLE) {
      double s2[3];
      for (int i = 0; i < 3; ++i) q->t0[i] = o->v[1][i];  // vector_object.h:164-166
      bh8h_sub(q->s1, o->v[2], o->v[1]);
      bh8h_sub(s2, o->v[4], o->v[1]);
      const double s1s1 = bh8h_dot(q->s1, q->s1);
      q->inv_s1s1 = BH8F_DIV(1.0, s1s1);
      if (o->tex_id >= 0) {
        if (o->tex_id >= BH8_MAX_TEXTURES || !tex_rows || tex_rows[o->tex_id] <= 0) return BH8F_TEXTURE;
        q->tex = o->tex_id;
        q->tex_rows = tex_rows[o->tex_id];
        q->tex_cols = tex_cols[o->tex_id];
        q->kw = BH8F_DIV((double)q->tex_cols, s1s1);
        q->kh = BH8F_DIV((double)q->tex_rows, BH8F_SQRT(bh8h_dot(s2, s2)));
      }
    }
    double wc[3];
    bh8h_sub(wc, f->cam, q->p0);
    const double side_cam = bh8h_dot(q->n, wc);
    if (side_cam == 0) f->first_resolve = 1;
    if (q->central) {
      if (f->n_central == 0) f->central_obj0 = k;
      f->central_mask |= 1u << k;
      f->n_central++;
    } else {
      const int j = f->n_nc++;
      f->noncentral_mask |= 1u << k;
      f->nc_obj[j] = k;
      f->nc_nF[j] = (float)q->nF;
      f->nc_c[j] = (float)q->c_bh;
      if (side_cam > 0) f->nc_cam_bits |= 1u << j;
      if (side_cam < 0) f->nc_cam_bits |= 1u << (16 + j);
      const double nn = BH8F_SQRT(bh8h_dot(q->n, q->n));  // 1 for the reference's shapes; not relied upon
      if (!(nn > 0)) return BH8F_ZERO_NORMAL;
      if (BH8F_DIV(fabs(q->c_bh), nn) < min_dist) min_dist = BH8F_DIV(fabs(q->c_bh), nn);
      if (nn > max_n) max_n = nn;
      if (fabs(q->c_bh) > max_c) max_c = fabs(q->c_bh);
    }
  }
  f->lease_kphi = f->noncentral_mask ? (float)BH8F_DIV(0.4995, max_n) : 0.0f;
  f->lease_ku = f->noncentral_mask ? (float)BH8F_DIV(0.24975, max_c) : 0.0f;
  f->nc_max_n = f->noncentral_mask ? (float)max_n : 0.0f;
  f->nc_max_c = f->noncentral_mask ? (float)max_c : 0.0f;
  // A chord between two points of the ray stays inside radius max(r1,r2); a plane at distance D
  // from the hole can only be met when max(r1,r2) >= D, i.e. min(u1,u2) <= 1/D (small margin).
  f->u_gate = f->noncentral_mask ? BH8F_DIV(BH8F_ADD(1.0, 1e-9), min_dist) : -1.0;
  // Horizon sphere R = 2M: a chord whose ends are both outside 1.5R and subtend <= 1 rad stays
  // outside R (1.5 cos(0.5) = 1.316 > 1).
  f->u_horizon = BH8F_DIV(1.0, BH8F_MUL(BH8F_MUL(1.5, f->R), BH8F_ADD(1.0, 1e-9)));
  if (!(f->u0 <= f->u_horizon)) f->first_resolve = 1;
  return BH8F_OK;
}

#if !defined(__CUDACC__) || defined(BH8_HOST_BUILD)
// ---- host side: reasons -> error codes and messages ---------------------------------------------
static inline int bh8_frame_error_code(int reason) {
  switch (reason) {
    case BH8F_OK: return BH8_OK;
    case BH8F_TWO_HOLES:
    case BH8F_PATTERN:
    case BH8F_KIND: return BH8_EUNSUPPORTED;
    default: return BH8_EINVAL;
  }
}
static inline const char* bh8_frame_error_text(int reason) {
  static const char* const kText[BH8F_N_ERRORS] = {
      "ok",
      "null scene / camera / params",
      "n_obj out of range",
      "unknown tracer",
      "linear_steps must be in [1, 65535]",
      "bh_index does not name a BH8_KIND_BLACKHOLE object",
      "camera size out of range",
      "nstep must be in [2, 32767]",
      "bad pixel_format",
      "black hole mass must be positive",
      "bad stripe sharding parameters",
      "more than one black hole in a scene is not supported",
      "InfinitePlane pattern is not BLACK or CHESS (opaque std::function)",
      "chess pattern_size too small: (int)(2*size) == 0 divides by zero in the reference",
      "object kind not supported by the GPU path (Triangle/Sphere/Cylinder)",
      "object refers to a texture slot that was never set",
      "plane with a zero normal"};
  return reason >= 0 && reason < BH8F_N_ERRORS ? kText[reason] : "unknown frame error";
}
// Returns BH8_OK or an error code and a message in err (>= 160 bytes).
static inline int bh8_build_frame(const bh8_scene* scene, const bh8_camera* cam, const bh8_params* prm,
                                  const int* tex_rows, const int* tex_cols, Bh8Frame* f, char* err) {
  const int reason = bh8_build_frame_core(scene, cam, prm, tex_rows, tex_cols, f);
  if (reason != BH8F_OK) strcpy(err, bh8_frame_error_text(reason));
  return bh8_frame_error_code(reason);
}
#endif  // host

#endif  // BH8_FRAME_H_
