/* bh8_oracle.c -- ORACLE: plain-C CPU restatement of blackhole_8's geodesic pixel loop.
 *
 * TEST INFRASTRUCTURE ONLY (see bh8_oracle.h).  Parity status: PINNED against the reference's own
 * classes run in the build container (oracle/_ref/ref_render -> tests/golden/*), byte for byte.
 *
 * Every function cites the reference lines it restates (paths relative to the reference's
 * include/blackhole/).  Floating-point expressions keep the reference's association and
 * evaluation order; build WITHOUT -ffast-math and without FMA contraction (plain x86-64 -O3, as
 * oracle/Makefile does) or the byte-for-byte pin is lost.
 *
 * Third-party arithmetic: the reference's vectors are cv::Vec / cv::Matx from OpenCV's
 * opencv2/core/matx.hpp (OpenCV is not vendored and its version is unpinned, CMakeLists.txt:27,60).
 * The helpers below restate the published semantics of that header, which are stable across
 * OpenCV 3.x/4.x: dot and Matx*Vec accumulate left to right from 0, normalize(v) scales by
 * (n ? 1./n : 0.), Vec/s multiplies by (1./s), Matx33::inv() is 1/det times the cofactor matrix
 * and all zeros when det == 0.
 */
#include "bh8_oracle.h"

#include <float.h>
#include <math.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct { double v[3]; } vec3;

/* ---- opencv2/core/matx.hpp semantics -------------------------------------------------- */

static vec3 v3(double a, double b, double c) { vec3 r = {{a, b, c}}; return r; }
static vec3 v3p(const double* p) { return v3(p[0], p[1], p[2]); }
static vec3 add(vec3 a, vec3 b) { return v3(a.v[0] + b.v[0], a.v[1] + b.v[1], a.v[2] + b.v[2]); }
static vec3 sub(vec3 a, vec3 b) { return v3(a.v[0] - b.v[0], a.v[1] - b.v[1], a.v[2] - b.v[2]); }
static vec3 scale(vec3 a, double s) { return v3(a.v[0] * s, a.v[1] * s, a.v[2] * s); }
static vec3 divs(vec3 a, double s) { return scale(a, 1. / s); } /* Vec / alpha = Vec * (1./alpha) */
static double dot(vec3 a, vec3 b) {
  double s = 0;
  s += a.v[0] * b.v[0];
  s += a.v[1] * b.v[1];
  s += a.v[2] * b.v[2];
  return s;
}
static vec3 cross(vec3 a, vec3 b) {
  return v3(a.v[1] * b.v[2] - a.v[2] * b.v[1], a.v[2] * b.v[0] - a.v[0] * b.v[2],
            a.v[0] * b.v[1] - a.v[1] * b.v[0]);
}
static vec3 normalize(vec3 a) {
  double s = 0;
  s += a.v[0] * a.v[0];
  s += a.v[1] * a.v[1];
  s += a.v[2] * a.v[2];
  const double nv = sqrt(s);
  return scale(a, nv ? 1. / nv : 0.);
}

typedef struct { double m[3][3]; } mat3;

static mat3 mat_inv(const mat3* A) {
  const double(*a)[3] = A->m;
  mat3 B;
  memset(&B, 0, sizeof B);
  double d = a[0][0] * (a[1][1] * a[2][2] - a[2][1] * a[1][2]) -
             a[0][1] * (a[1][0] * a[2][2] - a[2][0] * a[1][2]) +
             a[0][2] * (a[1][0] * a[2][1] - a[2][0] * a[1][1]);
  if (d == 0) return B;
  d = 1 / d;
  B.m[0][0] = (a[1][1] * a[2][2] - a[1][2] * a[2][1]) * d;
  B.m[0][1] = (a[0][2] * a[2][1] - a[0][1] * a[2][2]) * d;
  B.m[0][2] = (a[0][1] * a[1][2] - a[0][2] * a[1][1]) * d;
  B.m[1][0] = (a[1][2] * a[2][0] - a[1][0] * a[2][2]) * d;
  B.m[1][1] = (a[0][0] * a[2][2] - a[0][2] * a[2][0]) * d;
  B.m[1][2] = (a[0][2] * a[1][0] - a[0][0] * a[1][2]) * d;
  B.m[2][0] = (a[1][0] * a[2][1] - a[1][1] * a[2][0]) * d;
  B.m[2][1] = (a[0][1] * a[2][0] - a[0][0] * a[2][1]) * d;
  B.m[2][2] = (a[0][0] * a[1][1] - a[0][1] * a[1][0]) * d;
  return B;
}

static vec3 mat_vec(const mat3* A, vec3 x) {
  vec3 r;
  for (int i = 0; i < 3; ++i) {
    double s = 0;
    for (int k = 0; k < 3; ++k) s += A->m[i][k] * x.v[k];
    r.v[i] = s;
  }
  return r;
}

/* ---- blackhole_solution.h -------------------------------------------------------------- */

/* StaticBlackhole::G, blackhole_solution.h:27-29 */
double bh8_oracle_G(double mass, double u, double b) {
  return (u * u * (2.0 * mass * u - 1.0)) + (1.0 / (b * b));
}

/* StaticBlackhole::InvSqrtG, blackhole_solution.h:31-33 */
static double inv_sqrt_g(double mass, double u, double b) { return 1.0 / sqrt(bh8_oracle_G(mass, u, b)); }

/* StaticBlackhole::SolveG, blackhole_solution.h:35-53; epsilon<T>() = cbrt(eps), utility.h:17-21 */
double bh8_oracle_solve_g(double mass, double b) {
  double l = 0.0 + cbrt(DBL_EPSILON);
  double r = 1.0 / (3.0 * mass);
  double mid;
  for (int i = 0; i < 20; ++i) {
    mid = (l + r) / 2.0;
    if (bh8_oracle_G(mass, mid, b) > 0.0)
      l = mid;
    else
      r = mid;
  }
  return l;
}

/* StaticBlackhole::Collide, blackhole_solution.h:65-88 (horizon sphere, entry root only) */
static int collide_blackhole(const bh8_object* o, vec3 q1, vec3 q2, vec3* inter) {
  const vec3 center = v3p(o->v[0]);
  const double radius = 2 * o->mass;
  const vec3 p1 = sub(q1, center);
  const vec3 p2 = sub(q2, center);
  const double d1 = sqrt(dot(p1, p1));
  const double d2 = sqrt(dot(p2, p2));
  if (d1 < radius && d2 < radius) return 0;
  const vec3 Q1 = sub(q1, center);
  const vec3 Q2 = sub(q2, center);
  const vec3 Q = sub(Q2, Q1);
  const double c = pow(dot(Q, Q1), 2) - dot(Q, Q) * (dot(Q1, Q1) - radius * radius);
  if (c <= 0) return 0;
  const double t = (((double)1) / dot(Q, Q)) * (-dot(Q, Q1) - sqrt(c));
  if (t <= 0 || t >= 1) return 0;
  *inter = add(add(Q1, scale(Q, t)), center);
  return 1;
}

/* ---- object/vector_object.h ------------------------------------------------------------ */

/* Rectangle::Collide, vector_object.h:107-127 (corners are vertex()[1..4], object.h:38-40,109) */
static int collide_rectangle(const bh8_object* o, vec3 aa, vec3 bb, vec3* inter) {
  const vec3 p0 = v3p(o->v[1]);
  const vec3 p1 = sub(v3p(o->v[2]), p0);
  const vec3 p3 = sub(v3p(o->v[4]), p0);
  const vec3 n = normalize(cross(p1, p3));
  const vec3 Q1 = sub(aa, p0);
  const vec3 Q2 = sub(bb, p0);
  const double q1 = dot(n, Q1);
  const double q2 = dot(n, Q2);
  if (q1 * q2 >= 0) return 0;
  const vec3 Q = divs(add(scale(Q1, fabs(q2)), scale(Q2, fabs(q1))), fabs(q1) + fabs(q2));
  if (dot(p1, Q) > 0 && dot(p1, p1) > dot(p1, Q) && dot(p3, Q) > 0 && dot(p3, p3) > dot(p3, Q)) {
    *inter = add(Q, p0);
    return 1;
  }
  return 0;
}

/* Annulus::Collide, vector_object.h:328-347 (center() = vertex()[0], norm_ fixed at construction) */
static int collide_annulus(const bh8_object* o, vec3 q1, vec3 q2, vec3* inter) {
  const vec3 center = v3p(o->v[0]);
  const vec3 norm = v3p(o->n);
  const vec3 p1 = sub(q1, center);
  const vec3 p2 = sub(q2, center);
  if (dot(norm, p1) * dot(norm, p2) >= 0) return 0;
  const double t1 = dot(norm, p1);
  const double t2 = dot(norm, p2);
  const vec3 c = divs(add(scale(p1, fabs(t2)), scale(p2, fabs(t1))), fabs(t1) + fabs(t2));
  if (sqrt(dot(c, c)) > o->r_out) return 0;
  if (sqrt(dot(c, c)) < o->r_in) return 0;
  *inter = add(c, center);
  return 1;
}

/* InfinitePlane::Collide, vector_object.h:210-225 (touching counts as a hit: t1*t2 > 0 misses) */
static int collide_plane(const bh8_object* o, vec3 p1, vec3 p2, vec3* inter) {
  const double* vz = o->n;
  const double* pos = o->v[0];
  const double t1 = vz[0] * (p1.v[0] - pos[0]) + vz[1] * (p1.v[1] - pos[1]) + vz[2] * (p1.v[2] - pos[2]);
  const double t2 = vz[0] * (p2.v[0] - pos[0]) + vz[1] * (p2.v[1] - pos[1]) + vz[2] * (p2.v[2] - pos[2]);
  if (t1 * t2 > 0) return 0;
  const double d1 = fabs(t1);
  const double d2 = fabs(t2);
  *inter = divs(add(scale(p1, d2), scale(p2, d1)), d1 + d2);
  return 1;
}

static int collide(const bh8_object* o, vec3 p1, vec3 p2, vec3* inter) {
  switch (o->kind) {
    case BH8_KIND_BLACKHOLE: return collide_blackhole(o, p1, p2, inter);
    case BH8_KIND_ANNULUS: return collide_annulus(o, p1, p2, inter);
    case BH8_KIND_RECTANGLE: return collide_rectangle(o, p1, p2, inter);
    case BH8_KIND_INFINITE_PLANE: return collide_plane(o, p1, p2, inter);
    default: return 0; /* DrawableObject::Collide default, object.h:132 */
  }
}

/* ObjectManager::FindCollision, object_manager.h:69-85: nearest hit by distance^2 from p1; the
 * std::map keeps the FIRST inserted entry among equal keys (emplace does not overwrite). */
static int find_collision(const bh8_scene* s, vec3 p1, vec3 p2, vec3* inter) {
  int best = -1;
  double best_d = 0;
  vec3 best_p = v3(0, 0, 0);
  for (int i = 0; i < s->n_obj; ++i) {
    vec3 tmp;
    if (collide(&s->obj[i], p1, p2, &tmp)) {
      const vec3 d = sub(p1, tmp);
      const double key = dot(d, d);
      if (best < 0 || key < best_d) {
        best = i;
        best_d = key;
        best_p = tmp;
      }
    }
  }
  if (best >= 0) *inter = best_p;
  return best;
}

int bh8_oracle_find_collision(const bh8_scene* scene, const double p1[3], const double p2[3],
                              double inter[3]) {
  vec3 x = v3(0, 0, 0);
  const int k = find_collision(scene, v3p(p1), v3p(p2), &x);
  if (k >= 0) memcpy(inter, x.v, sizeof x.v);
  return k;
}

/* ChessPattern2D::mod / operator(), object/pattern.h:22-47 */
static double chess_mod(double size, double a) {
  const double b = fabs(a);
  return b - ((int)(b) / (int)(size * 2)) * (size * 2);
}
static int chess_white(double size, double x, double y) {
  const double x2 = chess_mod(size, x);
  const double y2 = chess_mod(size, y);
  const int same = (x2 <= size && y2 <= size) || (size <= x2 && size <= y2);
  if (x * y > 0) return same;
  return !same;
}

/* Rectangle::color (also Annulus, which inherits it), vector_object.h:159-179: nearest texel by
 * truncation, no filtering, BGR.  The reference does no bounds check; an index outside the image
 * is clamped here and counted in *oob (expected 0 for every BASELINE scene). */
static void color_textured(const bh8_object* o, const bh8_oracle_texture* tex, vec3 p, uint8_t bgr[3],
                           uint64_t* oob) {
  const vec3 s1 = sub(v3p(o->v[2]), v3p(o->v[1]));
  const vec3 s2 = sub(v3p(o->v[4]), v3p(o->v[1]));
  const vec3 v = sub(p, v3p(o->v[1]));
  const double r = sqrt(dot(v, v));
  double theta = acos(dot(s1, v) / (sqrt(dot(s1, s1)) * r));
  theta = isnan(theta) ? 0 : theta;
  if (!tex || !tex->bgr) {
    bgr[0] = bgr[1] = bgr[2] = 0;
    return;
  }
  const double fw = (r * cos(theta) / sqrt(dot(s1, s1))) * tex->cols;
  const double fh = (r * sin(theta) / sqrt(dot(s2, s2))) * tex->rows;
  long px_w = (fw >= -2147483648.0 && fw < 2147483648.0) ? (long)(int)fw : (long)INT32_MIN;
  long px_h = (fh >= -2147483648.0 && fh < 2147483648.0) ? (long)(int)fh : (long)INT32_MIN;
  long idx = px_h * tex->cols + px_w;
  const long n = (long)tex->rows * tex->cols;
  if (idx < 0 || idx >= n) {
    if (oob) ++*oob;
    idx = idx < 0 ? 0 : n - 1;
  }
  const uint8_t* ptr = tex->bgr + idx * 3;
  bgr[0] = ptr[0];
  bgr[1] = ptr[1];
  bgr[2] = ptr[2];
}

/* InfinitePlane::color, vector_object.h:227-232 (absolute x,y; only x,y components of vx,vy) */
static void color_plane(const bh8_object* o, vec3 p, uint8_t bgr[3]) {
  const double x = p.v[0], y = p.v[1];
  const double b = (o->ex[0] * y - o->ex[1] * x) / (o->ex[0] * o->ey[1] - o->ex[1] * o->ey[0]);
  const double a = (x - b * o->ey[0]) / o->ex[0];
  uint8_t c = 0;
  if (o->pattern == BH8_PATTERN_CHESS) c = chess_white(o->pattern_size, a, b) ? 255 : 0;
  bgr[0] = bgr[1] = bgr[2] = c;
}

static void color_of(const bh8_object* o, const bh8_oracle_texture* textures, int n_textures, vec3 p,
                     uint8_t bgr[3], uint64_t* oob) {
  switch (o->kind) {
    case BH8_KIND_ANNULUS:
    case BH8_KIND_RECTANGLE:
      color_textured(o, (o->tex_id >= 0 && o->tex_id < n_textures) ? &textures[o->tex_id] : 0, p, bgr, oob);
      break;
    case BH8_KIND_INFINITE_PLANE:
      color_plane(o, p, bgr);
      break;
    default: /* StaticBlackhole::color, blackhole_solution.h:61-63 */
      bgr[0] = bgr[1] = bgr[2] = 0;
  }
}

void bh8_oracle_color(const bh8_object* obj, const bh8_oracle_texture* tex, const double p[3],
                      uint8_t bgr[3]) {
  bh8_object o = *obj;
  if (o.tex_id >= 0) o.tex_id = 0;
  color_of(&o, tex, tex ? 1 : 0, v3p(p), bgr, 0);
}

static int class_of(const bh8_object* o) {
  return o->kind == BH8_KIND_BLACKHOLE ? BH8_CLASS_HORIZON
         : o->kind == BH8_KIND_ANNULUS ? BH8_CLASS_DISC
                                       : BH8_CLASS_OBJECT;
}

/* ---- the pixel loop body, blackhole_solution_test.cc:164-298 ---------------------------- */

typedef struct {
  uint8_t bgr[3];
  int hit; /* index in scene->obj or -1 */
  int steps;
} pixel_result;

typedef struct {
  const bh8_scene* scene;
  const bh8_oracle_texture* textures;
  int n_textures;
  const bh8_object* bh;
  vec3 cam_pos, vx, vy, vz, fv, bh_pos;
  double width, height, mass, b_c;
  int nstep;
} frame_ctx;

typedef struct {
  double u, phi, dphi_prev, du_h, b;
  mat3 M;
  vec3 prev;
} ray_state;

/* One geodesic update and its segment test (blackhole_solution_test.cc:218-235, 241-258, 275-290). */
static int advance(const frame_ctx* f, ray_state* s, double delta, pixel_result* out, uint64_t* oob) {
  s->u += delta;
  const double dphi = inv_sqrt_g(f->mass, s->u, s->b);
  s->phi += (s->dphi_prev + dphi) * s->du_h;
  const double r = 1. / s->u;
  const vec3 lv = v3(0, r * cos(s->phi), r * sin(s->phi));
  const vec3 cur = add(mat_vec(&s->M, lv), f->bh_pos);
  s->dphi_prev = dphi;
  ++out->steps;
  vec3 inter;
  const int k = find_collision(f->scene, s->prev, cur, &inter);
  if (k >= 0) {
    color_of(&f->scene->obj[k], f->textures, f->n_textures, inter, out->bgr, oob);
    out->hit = k;
    return 1;
  }
  s->prev = cur;
  return 0;
}

static void trace_pixel(const frame_ctx* f, int x, int y, pixel_result* out, uint64_t* oob) {
  out->bgr[0] = out->bgr[1] = out->bgr[2] = 0; /* frame pre-cleared, :140 */
  out->hit = -1;
  out->steps = 0;

  /* Camera::PixelVector(x, y, fv), camera.h:55-59 */
  const vec3 d = sub(sub(f->fv, scale(f->vy, (double)(f->width / 2.0 - x))),
                     scale(f->vz, (double)(f->height / 2.0 - y)));
  const vec3 pv = sub(d, f->bh_pos);        /* :167 (a direction minus a point, as written) */
  const vec3 F = sub(f->cam_pos, f->bh_pos); /* :168 */

  vec3 yv = normalize(sub(F, pv)); /* :170-174 */
  vec3 zv = normalize(cross(pv, yv));
  if (isnan(zv.v[0])) zv = v3(1, 0, 0);
  vec3 xv = normalize(cross(yv, zv));

  ray_state s;
  for (int i = 0; i < 3; ++i) { /* :176-179, columns zv yv xv */
    s.M.m[i][0] = zv.v[i];
    s.M.m[i][1] = yv.v[i];
    s.M.m[i][2] = xv.v[i];
  }
  const mat3 Minv = mat_inv(&s.M); /* :180 */
  const vec3 cf = mat_vec(&Minv, F); /* :182 */

  const double b = cf.v[2]; /* :185 */
  double sol = -1;
  if (b >= f->b_c) sol = bh8_oracle_solve_g(f->mass, b);
  const double periapsis = b < f->b_c ? 1.0 / (3 * f->mass) : sol; /* :190 */
  const double r0 = sqrt(dot(cf, cf));

  s.phi = atan(cf.v[2] / cf.v[1]); /* :193 */
  s.dphi_prev = 0;
  s.u = 1. / r0;
  s.b = b;
  const int nstep_safe = f->nstep - 1;
  const double du = (periapsis - s.u) * (1.0 / f->nstep); /* :204 */
  s.du_h = du / 2.;
  s.prev = f->cam_pos; /* :211 */

  for (int i = 0; i < nstep_safe; ++i) /* :217 */
    if (advance(f, &s, du, out, oob)) return;
  if (advance(f, &s, du * 0.9, out, oob)) return; /* :241, trapezoid width stays du/2 */

  if (sol < 0) { /* :264-272: captured ray, one straight chord to the centre */
    vec3 inter;
    const int k = find_collision(f->scene, s.prev, f->bh_pos, &inter);
    if (k >= 0) {
      color_of(&f->scene->obj[k], f->textures, f->n_textures, inter, out->bgr, oob);
      out->hit = k;
    }
    return;
  }

  for (int i = 0; i < nstep_safe; ++i) /* :274 */
    if (advance(f, &s, -du, out, oob)) return;
}

static int oracle_render(const bh8_scene* scene, const bh8_camera* cam, int nstep, int linear_steps,
                         const bh8_oracle_texture* textures, int n_textures, int row0, int row1,
                         int threads, uint8_t* out_bgr, uint8_t* out_class, int8_t* out_key,
                         uint16_t* out_steps, bh8_oracle_result* result);

/* One pixel of the flat-space driver, ray_tracer_test.cc:143-145:
 *   RayTracer(old = camera.focus(), present = camera.PixelVector(x, y, fv))   ray_tracer.h:59-61
 *   BasicLinearRayRecurrence: vec = present - old; if (vec.vec < 100) vec *= 100.0 / d2   :21-27
 *   Prograde(manager, dst, steps): FindCollision(old, present) then present + vec          :68-85 */
static void trace_pixel_linear(const frame_ctx* f, int x, int y, int steps, pixel_result* out, uint64_t* oob) {
  out->bgr[0] = out->bgr[1] = out->bgr[2] = 0;
  out->hit = -1;
  out->steps = 0;
  vec3 old = f->cam_pos;
  vec3 present = sub(sub(f->fv, scale(f->vy, (double)(f->width / 2.0 - x))),
                     scale(f->vz, (double)(f->height / 2.0 - y)));
  vec3 step = sub(present, old);
  const double d2 = dot(step, step);
  if (d2 < 100) step = scale(step, 100.0 / d2);
  for (int s = 0; s < steps; ++s) {
    ++out->steps;
    vec3 inter;
    const int k = find_collision(f->scene, old, present, &inter);
    if (k >= 0) {
      color_of(&f->scene->obj[k], f->textures, f->n_textures, inter, out->bgr, oob);
      out->hit = k;
      return;
    }
    const vec3 next = add(present, step);
    old = present;
    present = next;
  }
}

int bh8_oracle_render(const bh8_scene* scene, const bh8_camera* cam, int nstep,
                      const bh8_oracle_texture* textures, int n_textures, int row0, int row1,
                      int threads, uint8_t* out_bgr, uint8_t* out_class, int8_t* out_key,
                      uint16_t* out_steps, bh8_oracle_result* result) {
  return oracle_render(scene, cam, nstep, 0, textures, n_textures, row0, row1, threads, out_bgr, out_class,
                       out_key, out_steps, result);
}

int bh8_oracle_render_linear(const bh8_scene* scene, const bh8_camera* cam, int linear_steps,
                             const bh8_oracle_texture* textures, int n_textures, int row0, int row1,
                             int threads, uint8_t* out_bgr, uint8_t* out_class, int8_t* out_key,
                             uint16_t* out_steps, bh8_oracle_result* result) {
  if (linear_steps < 1) return -1;
  return oracle_render(scene, cam, 2, linear_steps, textures, n_textures, row0, row1, threads, out_bgr,
                       out_class, out_key, out_steps, result);
}

static int oracle_render(const bh8_scene* scene, const bh8_camera* cam, int nstep, int linear_steps,
                         const bh8_oracle_texture* textures, int n_textures, int row0, int row1,
                         int threads, uint8_t* out_bgr, uint8_t* out_class, int8_t* out_key,
                         uint16_t* out_steps, bh8_oracle_result* result) {
  static const bh8_object no_hole = {BH8_KIND_BLACKHOLE, -1, -1, 0, {{0}}, {0}, {0}, {0}, 0, 0, 1.0, 0};
  if (!scene || !cam || !out_bgr || scene->n_obj < 1 || scene->n_obj > BH8_MAX_OBJECTS || nstep < 2) return -1;
  const int has_hole = scene->bh_index >= 0 && scene->bh_index < scene->n_obj &&
                       scene->obj[scene->bh_index].kind == BH8_KIND_BLACKHOLE;
  if (!has_hole && !(linear_steps > 0 && scene->bh_index == -1)) return -1;
  frame_ctx f;
  f.scene = scene;
  f.textures = textures;
  f.n_textures = n_textures;
  f.bh = has_hole ? &scene->obj[scene->bh_index] : &no_hole;
  f.cam_pos = v3p(cam->pos);
  f.vx = v3p(cam->vx);
  f.vy = v3p(cam->vy);
  f.vz = v3p(cam->vz);
  f.fv = scale(f.vx, cam->focus_len); /* Camera::focus_vector, camera.h:61-63 */
  f.bh_pos = v3p(f.bh->v[0]);
  f.width = cam->width;
  f.height = cam->height;
  f.mass = f.bh->mass;
  f.b_c = 3.0 * sqrt(3) * f.mass; /* blackhole_solution.h:25 */
  f.nstep = nstep;
  if (row0 < 0) row0 = 0;
  if (row1 > cam->height) row1 = cam->height;

  uint64_t steps = 0, oob = 0, c0 = 0, c1 = 0, c2 = 0, c3 = 0;
  const int W = cam->width;
#ifdef _OPENMP
  if (threads < 1) threads = 1;
#pragma omp parallel for schedule(dynamic, 4) num_threads(threads) reduction(+ : steps, oob, c0, c1, c2, c3)
#endif
  for (int y = row0; y < row1; ++y) {
    for (int x = 0; x < W; ++x) {
      pixel_result p;
      if (linear_steps > 0)
        trace_pixel_linear(&f, x, y, linear_steps, &p, &oob);
      else
        trace_pixel(&f, x, y, &p, &oob);
      const size_t i = (size_t)y * W + x; /* :214 */
      out_bgr[i * 3 + 0] = p.bgr[0];
      out_bgr[i * 3 + 1] = p.bgr[1];
      out_bgr[i * 3 + 2] = p.bgr[2];
      const int cls = p.hit < 0 ? BH8_CLASS_BACKGROUND : class_of(&scene->obj[p.hit]);
      if (out_class) out_class[i] = (uint8_t)cls;
      if (out_key) out_key[i] = (int8_t)(p.hit < 0 ? -1 : scene->obj[p.hit].key);
      if (out_steps) out_steps[i] = (uint16_t)p.steps;
      steps += (uint64_t)p.steps;
      c0 += cls == 0;
      c1 += cls == 1;
      c2 += cls == 2;
      c3 += cls == 3;
    }
  }
  (void)threads;
  if (result) {
    result->rays = (uint64_t)(row1 > row0 ? row1 - row0 : 0) * (uint64_t)W;
    result->steps = steps;
    result->class_count[0] = c0;
    result->class_count[1] = c1;
    result->class_count[2] = c2;
    result->class_count[3] = c3;
    result->tex_oob = oob;
  }
  return 0;
}
