/* bh8_oracle.h -- ORACLE: plain-C CPU restatement of blackhole_8's geodesic pixel loop.
 *
 * TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this; the product library (libbh8.so) never links or calls it.
 *
 * Parity status: PINNED.  The reference's own tests hold no golden vectors for this path
 * (SURVEY.md 8c), so the pin is the reference itself run in the build container:
 * oracle/_ref/ref_render (the reference's classes, compiled from /root/reference) produced the
 * fixtures in tests/golden/ (tools/make_golden.py), and tests/test_oracle.py requires this port to
 * reproduce them byte for byte -- pixels, hit keys, classes and step counts.
 */
#ifndef BH8_ORACLE_H_
#define BH8_ORACLE_H_

#include <stdint.h>

#include "bh8.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct bh8_oracle_texture {
  const uint8_t* bgr; /* rows*cols*3, row-contiguous (cv::Mat CV_8UC3) */
  int32_t rows, cols;
} bh8_oracle_texture;

typedef struct bh8_oracle_result {
  uint64_t rays, steps, class_count[4], tex_oob;
} bh8_oracle_result;

/* Renders rows [row0,row1) of one frame with `threads` OpenMP threads.
 * out_bgr: H*W*3 (full-frame buffer, only the rows are written); out_class/out_key/out_steps
 * nullable full-frame buffers.  Returns 0, or -1 on a malformed scene. */
int bh8_oracle_render(const bh8_scene* scene, const bh8_camera* cam, int nstep,
                      const bh8_oracle_texture* textures, int n_textures, int row0, int row1,
                      int threads, uint8_t* out_bgr, uint8_t* out_class, int8_t* out_key,
                      uint16_t* out_steps, bh8_oracle_result* result);

/* Flat-space tracer (ray_tracer_test.cc:140-155, ray_tracer.h:17-35,68-85): up to linear_steps
 * segments per ray; the scene needs no black hole (bh_index = -1). */
int bh8_oracle_render_linear(const bh8_scene* scene, const bh8_camera* cam, int linear_steps,
                             const bh8_oracle_texture* textures, int n_textures, int row0, int row1,
                             int threads, uint8_t* out_bgr, uint8_t* out_class, int8_t* out_key,
                             uint16_t* out_steps, bh8_oracle_result* result);

/* Single-function known-answer hooks (tests/test_oracle.py). */
double bh8_oracle_G(double mass, double u, double b);
double bh8_oracle_solve_g(double mass, double b);
int bh8_oracle_find_collision(const bh8_scene* scene, const double p1[3], const double p2[3],
                              double inter[3]); /* index in scene->obj, -1 = none */
void bh8_oracle_color(const bh8_object* obj, const bh8_oracle_texture* tex, const double p[3],
                      uint8_t bgr[3]);

#ifdef __cplusplus
}
#endif
#endif /* BH8_ORACLE_H_ */
