/* bh8.h -- C ABI of the B200 renderer for blackhole_8's per-pixel null-geodesic hot path.
 *
 * The reference (lackhole/blackhole_8) has NO plugin / operator / FFI seam for this path: the
 * pixel loop is written inline in main() of include/blackhole/blackhole_solution_test.cc:161-308
 * and calls per-pixel C++ member functions.  This header is therefore the boundary a maintainer
 * would add: one batched call "scene snapshot + camera(s) -> frame(s)" that replaces the whole
 * loop body, with plain-old-data mirrors of the state that loop reads:
 *
 *   bh8_camera   <- blackhole::Camera<double>          camera.h:30-63  (position, basis, focus_len)
 *   bh8_object   <- blackhole::DrawableObject<double>  object/object.h:33-109 (vertex()[0..4]) plus
 *                   StaticBlackhole  blackhole_solution.h:24-88   (mass)
 *                   Annulus          object/vector_object.h:320-355 (norm_, r_outer_, r_inner_)
 *                   Rectangle        object/vector_object.h:94-179
 *                   InfinitePlane    object/vector_object.h:203-232 (position, vector_x/y/z) with
 *                   ChessPattern2D   object/pattern.h:20-47       (pattern_size_)
 *   bh8_scene    <- blackhole::ObjectManager<double>   object/object_manager.h:69-92 (objects_)
 *   textures     <- Material::texture_ (cv::Mat CV_8UC3, BGR)  object/material.h:29-60
 *   output       <- the cv::Mat frame written at blackhole_solution_test.cc:214,231-233
 *
 * Plain pointers and sizes only; the caller owns every host buffer, the context owns all device
 * memory, streams and texture objects.  Every function returns 0 on success and a negative
 * BH8_E* code on failure; bh8_last_error() gives the message.  There is no CPU fallback: without a
 * usable CUDA device bh8_create() fails.
 */
#ifndef BH8_H_
#define BH8_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BH8_ABI_VERSION 5
#define BH8_MAX_OBJECTS 16
#define BH8_MAX_TEXTURES 16
#define BH8_MAX_DEVICES 8

/* error codes */
#define BH8_OK 0
#define BH8_EINVAL (-1)      /* bad argument / malformed scene */
#define BH8_EUNSUPPORTED (-2) /* object kind or pattern the GPU path does not implement */
#define BH8_ECUDA (-3)       /* CUDA runtime error (message has the cudaError string) */
#define BH8_ENOMEM (-4)
#define BH8_ENODEVICE (-5)   /* no CUDA device: there is no CPU fallback */

/* object kinds (which Collide()/color() pair of the reference applies) */
enum {
  BH8_KIND_BLACKHOLE = 0,      /* StaticBlackhole: horizon sphere R = 2M, colour black */
  BH8_KIND_ANNULUS = 1,        /* thin accretion disc */
  BH8_KIND_RECTANGLE = 2,      /* textured rectangle */
  BH8_KIND_INFINITE_PLANE = 3  /* plane with a procedural pattern */
};

/* per-pixel hit class written to the optional class map */
enum {
  BH8_CLASS_BACKGROUND = 0, /* FindCollision returned nullptr on every segment: pixel stays 0 */
  BH8_CLASS_HORIZON = 1,
  BH8_CLASS_DISC = 2,
  BH8_CLASS_OBJECT = 3
};

/* InfinitePlane pattern_ (an opaque std::function in the reference; only these are recognised) */
enum {
  BH8_PATTERN_BLACK = 0, /* the default lambda: always (0,0,0) */
  BH8_PATTERN_CHESS = 1  /* ChessPattern2D{pattern_size} */
};

/* pixel formats of the output frame */
enum {
  BH8_PIXEL_RGBA8 = 0, /* 4 B/pixel R,G,B,255 */
  BH8_PIXEL_BGRA8 = 1, /* 4 B/pixel B,G,R,255 */
  BH8_PIXEL_BGR8 = 2   /* 3 B/pixel B,G,R: the reference's CV_8UC3 cv::Mat layout */
};

/* which stepper walks the ray (bh8_params.tracer) */
enum {
  BH8_TRACER_GEODESIC = 0, /* Schwarzschild null geodesic, blackhole_solution_test.cc:161-308 */
  BH8_TRACER_LINEAR = 1    /* flat space: RayTracer::Prograde with BasicLinearRayRecurrence,
                              ray_tracer.h:17-35,68-85 as driven by ray_tracer_test.cc:140-155;
                              the scene needs no black hole (bh_index = -1) */
};

/* bh8_params.flags */
#define BH8_FLAG_STATS 1u        /* accumulate bh8_stats counters on the device (tiny cost) */
#define BH8_FLAG_NO_BATCHING 2u  /* diagnostic: run every exact segment test at once instead of batching per warp */

typedef struct bh8_camera {
  double pos[3];    /* Camera::focus() */
  double vx[3];     /* Object::vector_x() (viewing direction) */
  double vy[3];     /* Object::vector_y() */
  double vz[3];     /* Object::vector_z() */
  double focus_len; /* width / (2 tan(fov/2)), camera.h:40 */
  int32_t width;
  int32_t height;
} bh8_camera;

typedef struct bh8_object {
  int32_t kind;        /* BH8_KIND_* */
  int32_t key;         /* ObjectManager key (insertion index), reported back in the key map */
  int32_t tex_id;      /* texture slot for ANNULUS / RECTANGLE, -1 = none */
  int32_t pattern;     /* BH8_PATTERN_* for INFINITE_PLANE */
  double v[5][3];      /* Object::vertex()[0..4]; [0] is position()/center(), corners are [1..4] */
  double n[3];         /* ANNULUS: norm_ (fixed at construction); INFINITE_PLANE: vector_z() */
  double ex[3];        /* INFINITE_PLANE: vector_x() */
  double ey[3];        /* INFINITE_PLANE: vector_y() */
  double r_in;         /* ANNULUS r_inner_ */
  double r_out;        /* ANNULUS r_outer_ */
  double mass;         /* BLACKHOLE mass() */
  double pattern_size; /* CHESS pattern_size_ */
} bh8_object;

typedef struct bh8_scene {
  int32_t n_obj;         /* <= BH8_MAX_OBJECTS, listed in ObjectManager iteration order */
  int32_t bh_index;      /* index in obj[] of the black hole the geodesics bend around */
  const bh8_object* obj;
} bh8_scene;

typedef struct bh8_params {
  int32_t nstep;         /* the literal 20 at blackhole_solution_test.cc:202; legs run nstep-1 steps */
  int32_t pixel_format;  /* BH8_PIXEL_* */
  uint32_t flags;        /* BH8_FLAG_* */
  /* Row-stripe sharding of one frame across devices / ranks: stripes of stripe_rows rows are dealt
   * round-robin, this call renders stripes s with s % shard_count == shard_index into the
   * full-frame buffer.  stripe_rows = 0 or shard_count <= 1 renders the whole frame. */
  int32_t stripe_rows;
  int32_t shard_index;
  int32_t shard_count;
  int32_t tracer;        /* BH8_TRACER_* */
  int32_t linear_steps;  /* BH8_TRACER_LINEAR: segments per ray, the 10 of ray_tracer_test.cc:145 */
} bh8_params;

typedef struct bh8_stats {
  uint64_t rays;         /* rays traced by this call */
  uint64_t steps;        /* geodesic updates executed (reference definition: per-ray count of
                            u/phi updates up to and including the one whose segment hit); LINEAR:
                            segments tested */
  uint64_t class_count[4];
  uint64_t tex_oob;      /* texture fetches whose reference index lay outside the image (clamped) */
  double kernel_ms;      /* device time of the render kernel(s), CUDA events, max over devices */
  double total_ms;       /* host wall time of the call */
  /* The warp schedule of the geodesic kernel (0 for the linear tracer).  A warp (an 8x4-pixel patch) runs
   * the same geodesic update for its 32 lanes until its last ray has ended:
   *   update_slots / warps      updates a warp issued; its rays needed steps / rays each on average, so
   *                             steps / (32 * update_slots) of the issued lane-updates moved a ray
   *   resolve_passes / warps    times a warp ran its parked exact segment tests (FindCollision) together
   *   exact_tests / rays        exact tests per ray */
  uint64_t warps;
  uint64_t update_slots;
  uint64_t resolve_passes;
  uint64_t exact_tests;
} bh8_stats;

typedef struct bh8_ctx bh8_ctx;

/* Create a context driving n_dev CUDA devices (devices[i] = ordinal; NULL = {0..n_dev-1}).
 * With n_dev > 1 peer access to devices[0] is enabled and frames are gathered there.  An ordinal may be
 * listed more than once: each entry is a logical device of its own (streams, buffers, textures). */
int bh8_create(bh8_ctx** out, const int* devices, int n_dev);
/* Sinks and scripts opened on the context use its device, streams and textures: close them first. */
void bh8_destroy(bh8_ctx* ctx);
const char* bh8_last_error(const bh8_ctx* ctx); /* ctx may be NULL: error of the failed bh8_create */
int bh8_abi_version(void);

/* Copy a CV_8UC3 BGR image (Material::texture_) into slot tex_id on every device of the context. */
int bh8_set_texture(bh8_ctx* ctx, int tex_id, const uint8_t* bgr, int rows, int cols,
                    size_t row_stride_bytes);

/* Render n_frames frames.  scenes[f] / cams[f] are the snapshots for frame f (host memory).
 * out_pixels: HOST buffer of n_frames * H * W * bpp bytes; out_class / out_key (nullable): HOST
 * buffers of n_frames * H * W bytes (hit class; ObjectManager key of the hit object, -1 = none);
 * out_steps (nullable): n_frames * H * W uint16 step counts.  Synchronous; includes the
 * host<->device copies.  With several devices whole frames are dealt round-robin when
 * n_frames >= n_dev, else each frame is striped over the devices (params->stripe_rows, default 16).
 * One frame on one device is drawn as five row bands (BH8_RENDER_BANDS, 1..8), each read back while the
 * next is drawn: the call ends one band's copy after the last kernel, not one frame's. */
int bh8_render(bh8_ctx* ctx, const bh8_scene* scenes, const bh8_camera* cams, int n_frames,
               const bh8_params* params, uint8_t* out_pixels, uint8_t* out_class, int8_t* out_key,
               uint16_t* out_steps, bh8_stats* stats);

/* Streaming form of bh8_render for one frame at a time: bh8_submit() launches the frame and queues
 * its read-back into out_pixels (HOST memory, ideally from bh8_host_alloc) and returns at once with
 * a ticket; bh8_wait() blocks until that frame is complete in out_pixels.  Two frames per device
 * may be in flight, each on its own stream (kernel, then read-back): the kernel of frame k+1 takes
 * the SMs frame k's last wave leaves idle and runs under frame k's copy; frames go to the
 * context's devices round-robin.  Tickets must be waited for in order of submission. */
int bh8_submit(bh8_ctx* ctx, const bh8_scene* scene, const bh8_camera* cam, const bh8_params* params,
               uint8_t* out_pixels, uint64_t* ticket);
int bh8_wait(bh8_ctx* ctx, uint64_t ticket);

/* Device-resident path (device 0 of the context).  d_pixels / d_class / d_key / d_steps are DEVICE
 * pointers for ONE frame -- they may be peer / IPC mappings of another GPU's memory, in which case
 * the kernel's stores go over NVLink and no separate gather copy is needed.  Asynchronous on the
 * context's stream; bh8_sync() waits.  The scene/camera snapshot (a few hundred bytes) is passed
 * as kernel parameters, so nothing else has to be resident. */
int bh8_render_device(bh8_ctx* ctx, const bh8_scene* scene, const bh8_camera* cam,
                      const bh8_params* params, void* d_pixels, void* d_class, void* d_key,
                      void* d_steps);
int bh8_sync(bh8_ctx* ctx);
/* Device time in ms between the first and the last kernel of the bh8_render_device() calls issued
 * since the previous bh8_timer_begin(); call after bh8_sync(). */
int bh8_timer_begin(bh8_ctx* ctx);
int bh8_timer_end_ms(bh8_ctx* ctx, double* ms);
/* Read (and reset) the device-side counters accumulated by BH8_FLAG_STATS launches. */
int bh8_read_stats(bh8_ctx* ctx, bh8_stats* stats);
/* Number of kernels this context has launched so far (the bench's gpu_launches claim). */
uint64_t bh8_launch_count(const bh8_ctx* ctx);

/* Frame buffers owned by the library (cudaMalloc on device 0 of the context) and their CUDA IPC
 * handles, for one-process-per-GPU sharding: rank 0 allocates and exports, the other ranks import
 * and pass the mapping to bh8_render_device(). */
int bh8_frame_alloc(bh8_ctx* ctx, size_t bytes, void** d_ptr);
int bh8_frame_free(bh8_ctx* ctx, void* d_ptr);
int bh8_ipc_export(bh8_ctx* ctx, void* d_ptr, uint8_t handle[64]);
int bh8_ipc_import(bh8_ctx* ctx, const uint8_t handle[64], void** d_ptr);
int bh8_ipc_close(bh8_ctx* ctx, void* d_ptr);
int bh8_memcpy_d2h(bh8_ctx* ctx, void* host, const void* d_ptr, size_t bytes);
/* Pinned host memory (cudaHostAlloc) so the frame read-back of bh8_render() is a true async DMA. */
int bh8_host_alloc(void** p, size_t bytes);
#define BH8_HOST_WRITE_COMBINED 1u /* cudaHostAllocWriteCombined: faster DMA target, slow for the CPU to read */
int bh8_host_alloc_flags(void** p, size_t bytes, unsigned flags);
int bh8_host_free(void* p);
/* Pin memory the caller already owns (cudaHostRegister, portable) -- e.g. the cv::Mat frame a driver hands to
 * bh8_render() every frame: a read-back into pageable memory is staged by the driver and takes 0.58 ms per
 * 1080p BGR8 frame instead of 0.21 (profiles/r02bj_pageable.txt).  Registering costs ~19 ms once; unregister
 * before the memory is freed. */
int bh8_host_register(void* p, size_t bytes);
int bh8_host_unregister(void* p);
/* Read-back ceiling of this process' GPU: `reps` device-to-host copies of `bytes` each, alternating
 * between the two pinned buffers exactly as bh8_submit() queues its frames' read-backs (staging slot ->
 * host, one copy per slot in flight) but with no kernel in between; wall time in *seconds. */
int bh8_measure_d2h(bh8_ctx* ctx, uint8_t* host_a, uint8_t* host_b, size_t bytes, int reps, double* seconds);
int bh8_memset_d(bh8_ctx* ctx, void* d_ptr, int value, size_t bytes);

/* FP64 pipe peak: runs a dependent-free DFMA chain on every SM of device 0 and reports the
 * sustained FP64 FLOP/s (FMA = 2) -- the roofline denominator for this path. */
int bh8_measure_fp64_peak(bh8_ctx* ctx, double* flops_per_s, double* seconds_run);

/* Precision study: geodesic updates per second of the bare update chain (no hit logic), in FP64 as
 * the renderer runs it (fp32 = 0) or in FP32 (fp32 = 1).  See DESIGN.md 4.4. */
int bh8_measure_stepping(bh8_ctx* ctx, int fp32, double* updates_per_s);

/* ---- Frame sink ------------------------------------------------------------
 * Replaces cv::VideoWriter(path, cv::VideoWriter::fourcc('M','J','P','G'), fps, size, true) and
 * out_capture.write(frame) of blackhole_solution_test.cc:71-72,334: a Motion-JPEG AVI file whose
 * frames are JPEG-encoded ON THE GPU by this library's own kernels (baseline, 4:2:0, IJG quantisation
 * tables scaled by `quality`, standard Huffman tables, restart markers; csrc/bh8_jpeg.cuh) straight from
 * the device-resident BGR8 frame, so only the bitstream crosses PCIe.  $BH8_SINK_NVJPEG=1 selects
 * nvJPEG instead (kept for comparison).
 *   ctx != NULL: GPU sink on device 0 of the context.  ctx == NULL: host-only container that is fed
 *                ready JPEGs with bh8_sink_append_jpeg() (no GPU needed).
 *   avi_path:    file to write, or NULL to encode only (read each frame with bh8_sink_last_jpeg()).
 * Errors: negative return, text from bh8_sink_last_error() (bh8_last_error(ctx) for bh8_sink_open). */
typedef struct bh8_sink bh8_sink;
int bh8_sink_open(bh8_ctx* ctx, const char* avi_path, int width, int height, double fps, int quality,
                  bh8_sink** out);
/* Render one frame (as bh8_render_device, pixel format forced to BGR8) into the sink's own device
 * buffer, encode it and append it: the GPU form of "trace the frame; out_capture.write(frame)". */
int bh8_sink_render(bh8_sink* sink, const bh8_scene* scene, const bh8_camera* cam, const bh8_params* params);
/* Pipelined form of bh8_sink_render: queues the frame's kernel and its JPEG encode on one of two
 * internal streams and returns; the bitstream is fetched and appended when the slot is needed again
 * (two calls later), by bh8_sink_flush() or by bh8_sink_close().  Frame k+1 is traced while frame k
 * is encoded and read back; frames reach the file in submission order.  bh8_sink_last_jpeg() and
 * bh8_sink_stats() describe the frames appended so far. */
int bh8_sink_submit(bh8_sink* sink, const bh8_scene* scene, const bh8_camera* cam, const bh8_params* params);
int bh8_sink_flush(bh8_sink* sink);
/* HUD text of the reference's frames: blackhole_solution_test.cc:309-326 draws five lines of
 * cv::putText(FONT_HERSHEY_PLAIN, 1, green, 1) into the frame BEFORE out_capture.write (:334).  The lines set
 * here are drawn on the device into every frame the sink renders from now on, before it is encoded
 * (n_lines = 0: none); for text inside the image the pixels equal cv::putText's bit for bit. */
#define BH8_HUD_MAX_LINES 16
#define BH8_HUD_MAX_TEXT 128
typedef struct bh8_hud_line {
  int32_t x, y;       /* bottom-left corner of the text, as in cv::putText */
  uint8_t b, g, r, reserved;
  char text[BH8_HUD_MAX_TEXT]; /* NUL-terminated */
} bh8_hud_line;
int bh8_sink_hud(bh8_sink* sink, const bh8_hud_line* lines, int n_lines);
/* The same blit on any BGR8 frame in device memory (device 0 of the context); synchronous. */
int bh8_hud_draw_device(bh8_ctx* ctx, void* d_bgr_frame, int width, int height, const bh8_hud_line* lines, int n_lines);
/* Encode and append a BGR8 frame that is already in device memory (H * W * 3 bytes, device 0). */
int bh8_sink_write_device(bh8_sink* sink, const void* d_bgr_frame);
int bh8_sink_append_jpeg(bh8_sink* sink, const uint8_t* jpeg, size_t bytes);
/* Bitstream of the most recent frame; valid until the next call on this sink. */
int bh8_sink_last_jpeg(bh8_sink* sink, const uint8_t** data, size_t* bytes);
/* Frames so far, their JPEG bytes, and the device time the encoder took for them (CUDA events, ms). */
int bh8_sink_stats(const bh8_sink* sink, uint64_t* frames, uint64_t* jpeg_bytes, double* encode_ms);
const char* bh8_sink_last_error(const bh8_sink* sink);
/* Finish the AVI (index, sizes), release everything; the handle is invalid afterwards. */
int bh8_sink_close(bh8_sink* sink, uint64_t* file_bytes);

/* One file from the per-GPU files of a frame-sharded job (host only, no GPU needed): frame k of the
 * result is frame k / n_parts of part k % n_parts -- the round-robin ownership of whole frames that
 * sharding.frames_of / bh8_render use -- so the merged video.avi has the frames in the order the
 * reference's single loop would have written them.  The parts must agree in size and rate and their
 * frame counts must be a round-robin split.  Errors: bh8_last_error(NULL). */
int bh8_sink_merge(const char* const* part_paths, int n_parts, const char* out_path, uint64_t* frames,
                   uint64_t* file_bytes);

/* ---- Scene / camera animation on the device -----------------------------------
 * The reference animates on the host between frames: key handlers call Camera::MoveX / RotateZ / ...
 * and the disc spins by Annulus::RotateZ(pi/180) after every frame (blackhole_solution_test.cc:346-407;
 * Object::Move* / Rotate*, object/object.h:58-88; RotationMatrixForAxis, matrix.h:342-356).  A script
 * is that sequence written down: bh8_script_create() replays it ON THE GPU (bit-identical to the
 * reference's classes), leaves one camera + object snapshot per frame in device memory and derives
 * every frame's constants there as well; bh8_script_render() then draws frame k without any per-frame
 * host data (no host replay, no snapshot upload). */
enum {
  BH8_OP_MOVE_X = 0,   /* Object::MoveX(amount): every vertex += vector_x() * amount */
  BH8_OP_MOVE_Y = 1,
  BH8_OP_MOVE_Z = 2,
  BH8_OP_ROTATE_X = 3, /* Object::RotateX(amount): about vector_x() through position() */
  BH8_OP_ROTATE_Y = 4,
  BH8_OP_ROTATE_Z = 5,
  BH8_OP_MOVE_TO = 6   /* Object::MoveTo(to): vertex()[0] = to (the other vertices stay, as in the reference) */
};
#define BH8_TARGET_CAMERA (-1)

typedef struct bh8_action {
  int32_t frame;   /* applied AFTER frame `frame` is drawn, i.e. it shapes frame + 1 (the reference draws,
                      then handles the key and spins the disc); actions must be sorted by frame, actions of
                      one frame apply in list order */
  int32_t target;  /* BH8_TARGET_CAMERA or an index into scene->obj */
  int32_t op;      /* BH8_OP_* */
  int32_t reserved;
  double amount;   /* distance, or angle in radians */
  double to[3];    /* BH8_OP_MOVE_TO */
} bh8_action;

typedef struct bh8_basis { /* Object::vector_x/y/z() of an object at frame 0 */
  double vx[3], vy[3], vz[3];
} bh8_basis;

typedef struct bh8_script bh8_script;

/* scene0 / cam0: the state at frame 0.  obj_basis: n_obj entries, or NULL for the basis every shape
 * the reference's drivers build has -- (1,0,0),(0,1,0),(0,0,1) -- except INFINITE_PLANE objects, whose
 * basis is their (ex, ey, n).  params: as for bh8_render_device (nstep, pixel format, sharding, tracer).
 * Runs on device 0 of the context; synchronous (it reports a frame whose scene is invalid). */
int bh8_script_create(bh8_ctx* ctx, const bh8_scene* scene0, const bh8_basis* obj_basis, const bh8_camera* cam0,
                      const bh8_params* params, const bh8_action* actions, int n_actions, int n_frames,
                      bh8_script** out);
int bh8_script_frames(const bh8_script* script);
/* Draw frame `frame` into device buffers (as bh8_render_device; d_class / d_key / d_steps nullable).
 * Asynchronous on the context's stream; bh8_sync() waits. */
int bh8_script_render(bh8_script* script, int frame, void* d_pixels, void* d_class, void* d_key, void* d_steps);
/* Draw frames first .. first + count - 1 into d_base + i * frame_stride_bytes (device memory; the stride
 * is at least one frame, so frames do not overlap).  The launch sequence is captured once into a CUDA
 * graph (and re-used while the arguments stay the same), so the whole range costs the host ONE launch.
 * Asynchronous on the context's stream. */
int bh8_script_render_range(bh8_script* script, int first, int count, void* d_base, size_t frame_stride_bytes);
/* Read frame `frame`'s snapshot back: cam (nullable) and objs (nullable, n_obj entries). */
int bh8_script_state(bh8_script* script, int frame, bh8_camera* cam, bh8_object* objs);
/* Verification hook: the device-built frame constants of frame `frame` (bytes = bh8_frame_bytes()) and
 * the ones the host derives from a snapshot -- the two must be identical. */
int bh8_script_frame_constants(bh8_script* script, int frame, void* out, size_t bytes);
int bh8_host_frame_constants(const bh8_ctx* ctx, const bh8_scene* scene, const bh8_camera* cam,
                             const bh8_params* params, void* out, size_t bytes);
size_t bh8_frame_bytes(void);
void bh8_script_destroy(bh8_script* script);

/* ---- HUD text (host) -------------------------------------------------------------------------------
 * cv::putText(img, text, {x, y}, cv::FONT_HERSHEY_PLAIN, 1, {b, g, r}, 1) of blackhole_solution_test.cc:313-325
 * on a CV_8UC3 (BGR) frame in HOST memory; (x, y) is the bottom-left corner of the text as in OpenCV.  For
 * text inside the image the pixels equal OpenCV's bit for bit.  No GPU needed. */
int bh8_draw_text(uint8_t* bgr, int rows, int cols, size_t row_stride_bytes, int x, int y, const char* text,
                  int b, int g, int r);

size_t bh8_pixel_bytes(int pixel_format);
/* Bytes that travel host -> device per frame: the frame constants derived from the snapshot, passed as
 * kernel parameters (there is no other per-frame input). */
size_t bh8_launch_param_bytes(void);

#ifdef __cplusplus
}
#endif

#endif /* BH8_H_ */
