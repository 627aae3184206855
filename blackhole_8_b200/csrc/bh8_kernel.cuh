// bh8_kernel.cuh -- the render kernel: one thread per ray, one 32x8-pixel tile per 256-thread CTA.
//
// Replaces the pixel loop of blackhole_solution_test.cc:162-298 (ray setup, inbound steps, 0.9-step,
// captured chord, outbound steps, colour, frame store) for a whole frame or a set of row stripes.
//
// Warp schedule.  A warp is an 8x4-pixel patch, so neighbouring rays end at neighbouring steps.
// Every lane walks its own ray through ONE loop whose body is the lean geodesic update
// (lane_update: 11 FP64 instructions + MUFU.RSQ64H, three integer compares); the step index is per-lane
// state, so lanes may drift apart.  A lane whose segment needs the exact object test (a plane crossing, a
// possible horizon hit) does not run it on the spot -- that would serialise the warp once per distinct
// hit step -- but freezes (kPend) while the other lanes keep stepping: its increments are zero, the update
// leaves it where it is.  One __reduce_or_sync() of the lanes' states per round of UPV updates tells the
// warp who waits and who still runs (the usual round -- everybody runs or has ended -- is one uniform
// compare); the exact tests (lane_resolve: sincos, 1/u, the objects' Collide()) are executed by all 32
// lanes together, on their mailboxes, once no lane is left running or the oldest has waited
// `resolve_wait` rounds.  The pass colours the rays that end and tells whether any ray is still alive:
// the usual pass is the last one and the warp leaves at once.
// Plain launches store each warp's patch directly; launches with maps or counters go through a shared
// RGBA tile that leaves as 16-byte coalesced stores.
#ifndef BH8_KERNEL_CUH_
#define BH8_KERNEL_CUH_

#include <cuda_runtime.h>

#include "bh8_ray.cuh"
#include "bh8_warp.cuh"

namespace bh8 {

constexpr int kTileW = 32;
#ifndef BH8_TILE_H
#define BH8_TILE_H 8  // tile rows per CTA (multiple of 4: a warp is an 8x4 patch)
#endif
constexpr int kTileH = BH8_TILE_H;
constexpr int kThreads = kTileW * kTileH;
#ifndef BH8_PATCH_W
#define BH8_PATCH_W 8  // a warp is a BH8_PATCH_W x (32 / BH8_PATCH_W) pixel patch (measured: 8x4 best of 4x8, 8x4, 16x2)
#endif
constexpr int kPatchW = BH8_PATCH_W, kPatchH = 32 / BH8_PATCH_W, kPatchesAcross = kTileW / kPatchW;
static_assert(kPatchW * kPatchH == 32 && kTileH % kPatchH == 0, "a warp must tile the CTA's pixels");
constexpr int kMaxFilterPlanes = kMaxFilterSlots;
constexpr int kFineNstep = 64;  // from here on the kernel takes 3 updates per round of votes instead of 2
#ifndef BH8_MIN_BLOCKS
#define BH8_MIN_BLOCKS 5  // CTAs per SM the register allocation is sized for (48 registers; measured best of 3..6 on B200)
#endif
// The mailbox is 88 + 4 (10 + 2 NN) bytes per ray, so up to NN = 1 SIX CTAs would fit the SM's shared memory
// at 40 registers (-DBH8_MIN_BLOCKS=6).  Measured: 2 % slower than five at 48 (0.1638 vs 0.1607 ms) -- the
// spill code of the 40-register build costs more instructions than the extra eight warps hide; issue slots
// stay 79 % busy either way.  Four CTAs (dummy shared memory, same registers): 3.7 % slower.
template <int NN>
struct MinBlocks {
  static constexpr int value = (NN >= 2 && BH8_MIN_BLOCKS > 5) ? 5 : BH8_MIN_BLOCKS;
};

struct Bh8Tex {
  unsigned long long obj[BH8_MAX_TEXTURES];
};

struct Bh8Out {
  uint8_t* pixels;
  uint8_t* cls;
  int8_t* key;
  uint16_t* steps;
  unsigned long long* stats;  // [0] rays [1] steps [2..5] class counts [6] tex_oob [7] warps [8] update slots
                              // [9] resolve passes [10] exact tests (see bh8_stats)
  int32_t vec_ok;             // (linear kernel) 16-byte row stores are legal: 4-byte formats width % 4 == 0, BGR8
                              // width % 32 == 0; base 16-byte aligned
  // persistent-warp scheduling of the geodesic kernel (render_warps): ticket counter + leavers, both zero
  // between launches; the frame as tiles of 32x8 pixels, 8 patches each
  unsigned* sched;
  int32_t tiles_x, n_patches;
};
constexpr int kStatSlots = 11;

struct DeviceFetch {
  const Bh8Tex& tex;
  __device__ __forceinline__ uint32_t operator()(int slot, int col, int row) const {
    const uchar4 t = tex2D<uchar4>(tex.obj[slot], col + 0.5f, row + 0.5f);
    return (uint32_t)t.x | ((uint32_t)t.y << 8) | ((uint32_t)t.z << 16);
  }
};

// Pixel store + optional maps + optional counters, shared by both kernels.  RGBA8 / BGRA8 go through
// a shared tile and leave as 16-byte coalesced stores (4 pixels per store, 128 B per tile row).
__device__ __forceinline__ void store_pixel(const Bh8Frame& f, const Bh8Out& out, uint32_t* sh_rgba,
                                            unsigned long long* sh_red, int tid, int lane, int slot, int x0,
                                            int y0, int x, int y, bool inside, uint32_t bgr, uint32_t oob,
                                            int cls, int key, int steps, unsigned sched_slots = 0,
                                            unsigned sched_passes = 0, unsigned sched_tests = 0) {
  const size_t gi = (size_t)y * f.width + x;
  if (f.pixel_format == BH8_PIXEL_BGR8) {
    if (out.vec_ok) {
      // The reference's CV_8UC3 frame: a tile row is 32 x 3 = 96 contiguous bytes = six 16-byte stores
      // (width % 32 == 0, so every tile is whole in x and 96 x0/32 keeps the alignment).
      uint8_t* sb = reinterpret_cast<uint8_t*>(sh_rgba) + slot * 3;
      sb[0] = (uint8_t)bgr;
      sb[1] = (uint8_t)(bgr >> 8);
      sb[2] = (uint8_t)(bgr >> 16);
      __syncthreads();
      if (tid < kTileH * 6) {
        const int row = tid / 6, q = tid - row * 6;
        if (y0 + row < f.height) {
          const uint4 v = *reinterpret_cast<const uint4*>(reinterpret_cast<const uint8_t*>(sh_rgba) + row * 96 + q * 16);
          *reinterpret_cast<uint4*>(out.pixels + ((size_t)(y0 + row) * f.width + x0) * 3 + q * 16) = v;
        }
      }
    } else if (inside) {
      uint8_t* d = out.pixels + gi * 3;
      d[0] = (uint8_t)bgr;
      d[1] = (uint8_t)(bgr >> 8);
      d[2] = (uint8_t)(bgr >> 16);
    }
  } else {
    const uint32_t px4 = (f.pixel_format == BH8_PIXEL_BGRA8)
                             ? (bgr | 0xFF000000u)
                             : (((bgr >> 16) & 0xFFu) | (bgr & 0xFF00u) | ((bgr & 0xFFu) << 16) | 0xFF000000u);
    if (out.vec_ok) {
      sh_rgba[slot] = px4;
      __syncthreads();
      if (tid < kThreads / 4) {
        const int row = tid >> 3, col = (tid & 7) * 4;
        if (x0 + col < f.width && y0 + row < f.height) {
          const uint4 v = *reinterpret_cast<const uint4*>(&sh_rgba[row * kTileW + col]);
          *reinterpret_cast<uint4*>(out.pixels + ((size_t)(y0 + row) * f.width + x0 + col) * 4) = v;
        }
      }
    } else if (inside) {
      reinterpret_cast<uint32_t*>(out.pixels)[gi] = px4;
    }
  }
  if (inside) {
    if (out.cls) out.cls[gi] = (uint8_t)cls;
    if (out.key) out.key[gi] = (int8_t)key;
    if (out.steps) out.steps[gi] = (uint16_t)steps;
  }

  // ---- optional counters ----------------------------------------------------------------------------
  if (f.flags & BH8_FLAG_STATS) {
    if (tid < kStatSlots) sh_red[tid] = 0ull;
    __syncthreads();
    unsigned long long v[kStatSlots];
    v[0] = inside ? 1ull : 0ull;
    v[1] = inside ? (unsigned long long)steps : 0ull;
    for (int k = 0; k < 4; ++k) v[2 + k] = (inside && cls == k) ? 1ull : 0ull;
    v[6] = oob;
    // the warp schedule (bh8_render_kernel's STATS instantiation; 0 elsewhere): per warp, so lane 0 speaks
    v[7] = (lane == 0) ? 1ull : 0ull;
    v[8] = (lane == 0) ? sched_slots : 0ull;
    v[9] = (lane == 0) ? sched_passes : 0ull;
    v[10] = sched_tests;
#pragma unroll
    for (int k = 0; k < kStatSlots; ++k) {
      unsigned long long s = v[k];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (lane == 0 && s) atomicAdd(&sh_red[k], s);
    }
    __syncthreads();
    if (tid < kStatSlots && sh_red[tid]) atomicAdd(&out.stats[tid], sh_red[tid]);
  }
}

// ---- the geodesic kernel ------------------------------------------------------------------------------
// `f` is a __grid_constant__ kernel parameter, i.e. in the constant bank: its fields are direct operands
// of the FP64 instructions.
// STATS: the instantiation BH8_FLAG_STATS launches take -- it also counts the warp schedule (update slots,
// resolve passes, exact tests); the plain one carries no counter through the stepping loop.

struct SchedCount {
  unsigned n_iter = 0, n_pass = 0, n_test = 0;
};

// One 8x4-pixel patch, all 32 lanes of the warp: ray setup, then stepping phases and exact passes in turn
// until every ray of the patch has ended.  The result waits in each lane's mailbox (mail_steps, mail_hit,
// kMwBgr, kMwOob).
template <int NN, bool STATS, int UPV>
__device__ __forceinline__ void trace_patch(const Bh8Frame& f, const DeviceFetch& fetch, const Mail mail,
                                            uint32_t sc_addr, int x, int y, bool inside, SchedCount& sc_n) {
  Lane<NN> L;
  lane_inert(L);
  if (inside) {
    lane_setup(f, x, y, L, mail);
    lane_park_constants(L, mail);
  }
  for (;;) {
    // Lean stepping: the same straight-line update for every lane (frozen lanes are inert, see
    // lane_freeze), a few updates per round of warp votes, until the warp decides to attend to its
    // parked lanes (or every ray has ended).  `waited` and the update's constants live in this phase
    // only: nothing of the warp's bookkeeping is carried across the exact pass.
    int waited = 0, todo;
    const StepConst sc = StepConst::load_shared(sc_addr);
    for (;;) {
#pragma unroll
      for (int k = 0; k < UPV; ++k) lane_update(f, L, mail, sc);
      if (STATS) ++sc_n.n_iter;
      const unsigned present = __reduce_or_sync(0xffffffffu, (unsigned)L.state);
      if (present == (unsigned)kRun) continue;  // the usual round: nobody waits (ended lanes are state 0)
      todo = warp_decide(present, waited, f.resolve_wait);
      if (todo != kWarpStep) break;
    }
    if (todo == kWarpDone) break;  // every ray of the patch has ended
    if (STATS) {
      ++sc_n.n_pass;
      if (L.state & (kPend | kPendChord)) ++sc_n.n_test;
    }
    // All 32 lanes together; it colours the rays that end.  The usual pass is the last one (every ray of
    // the patch met something): no further round then.
    if (!lane_resolve(f, L, mail, fetch)) break;
  }
}

// Mailbox of this thread + the warp's copy of the stepping constants (StepConst) in shared memory.
template <int NN>
__device__ __forceinline__ Mail make_mail(const Bh8Frame& f, int tid, int lane, int warp, uint32_t* sc_addr) {
  __shared__ double sh_md[kMailDoubles * kThreads];
  __shared__ int sh_mi[MailInts<NN>::value * kThreads];
  __shared__ double sh_step_const[kThreads / 32][2];
  Mail mail;
#if defined(__CUDA_ARCH__)
  mail.d = (uint32_t)__cvta_generic_to_shared(sh_md + tid);
  mail.w = (uint32_t)__cvta_generic_to_shared(sh_mi + tid);
  asm volatile("" : "+r"(mail.d), "+r"(mail.w));  // opaque: two registers, never re-derived
#else
  mail.d = sh_md + tid;  // (nvcc's host pass only parses this)
  mail.w = sh_mi + tid;
#endif
  mail.stride = kThreads;
  // The two FP64 constants of the geodesic update (2M, 3/8) go through shared memory once per warp so
  // that the stepping loop holds them in registers: see StepConst.
  if (lane == 0) {
    sh_step_const[warp][0] = f.two_m;
    sh_step_const[warp][1] = 0.375;
  }
  __syncwarp();
  *sc_addr = (uint32_t)__cvta_generic_to_shared(&sh_step_const[warp][0]);
  return mail;
}

#if !defined(BH8_PERSISTENT_WARPS)
// One 32x8 tile of the frame per CTA; a warp is one 8x4 patch of it.  UPV: geodesic updates between two
// rounds of warp votes -- 2 for the reference's nstep 20, 3 for fine steps (nstep >= 64: lanes drift apart
// more slowly relative to the length of a ray; measured +2.4 % at 8K / nstep 200, -1.3 % at nstep 20).
template <int NN, bool STATS, int UPV>
__device__ __forceinline__ void render_tile(const Bh8Frame& f, const Bh8Tex& tex, const Bh8Out& out) {
  __shared__ __align__(16) uint32_t sh_rgba[kThreads];
  __shared__ unsigned long long sh_red[kStatSlots];
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;

  // ---- tile origin; blockIdx.y counts tile rows of THIS shard's stripes -----------------------
  const int x0 = blockIdx.x * kTileW;
  int y0;
  if (f.shard_count > 1) {
    const int tiles_per_stripe = f.stripe_rows / kTileH;
    const int ls = blockIdx.y / tiles_per_stripe;
    const int within = blockIdx.y - ls * tiles_per_stripe;
    y0 = (ls * f.shard_count + f.shard_index) * f.stripe_rows + within * kTileH;
  } else {
    y0 = blockIdx.y * kTileH;
  }
  if (y0 >= f.height) return;

  // warp = kPatchW x kPatchH patch (8x4); slot = row-major position in the 32x8 tile
  const int px = (warp % kPatchesAcross) * kPatchW + (lane % kPatchW);
  const int py = (warp / kPatchesAcross) * kPatchH + (lane / kPatchW);
  const int slot = py * kTileW + px;
  const int x = x0 + px, y = y0 + py;
  const bool inside = x < f.width && y < f.height;

  uint32_t sc_addr;
  const Mail mail = make_mail<NN>(f, tid, lane, warp, &sc_addr);
#if defined(BH8_DUMMY_SMEM)  // occupancy experiment: shared memory nobody uses, to run 4 instead of 5 CTAs per SM
  __shared__ volatile char sh_dummy[BH8_DUMMY_SMEM];
  if (f.width < 0) sh_dummy[tid] = 1;
#endif
  const DeviceFetch fetch{tex};
  SchedCount n;
  trace_patch<NN, STATS, UPV>(f, fetch, mail, sc_addr, x, y, inside, n);

  // ---- colour: lane_exact left it in the mailbox when the ray hit -----------------------------------
  const uint32_t result = inside ? (uint32_t)mail.get_w(kMwFlags) : 0u;  // flags | steps << 8 | (hit + 1) << 24
  const int hit_obj = (int)(result >> 24) - 1;
  const uint32_t bgr = hit_obj >= 0 ? (uint32_t)mail.get_w(kMwBgr) : 0u;
  // Plain launches (no maps, no counters): every warp stores its own patch the moment it is done -- four
  // 32-byte row segments for the 4-byte formats, 24-byte ones for BGR8 -- and leaves; no CTA barrier, no
  // staging tile (+1.5 % over the tile path on B200, which the launches with maps / counters still take).
  if (!STATS && !out.cls && !out.key && !out.steps) {
    if (inside) {
      const size_t gi = (size_t)y * f.width + x;
      if (f.pixel_format == BH8_PIXEL_BGR8) {
        uint8_t* d = out.pixels + gi * 3;
        d[0] = (uint8_t)bgr;
        d[1] = (uint8_t)(bgr >> 8);
        d[2] = (uint8_t)(bgr >> 16);
      } else {
        const uint32_t px4 = (f.pixel_format == BH8_PIXEL_BGRA8)
                                 ? (bgr | 0xFF000000u)
                                 : (((bgr >> 16) & 0xFFu) | (bgr & 0xFF00u) | ((bgr & 0xFFu) << 16) | 0xFF000000u);
        reinterpret_cast<uint32_t*>(out.pixels)[gi] = px4;
      }
    }
    return;
  }
  const int steps = (int)((result >> 8) & 0xFFFFu);
  const uint32_t oob = hit_obj >= 0 ? (uint32_t)mail.get_w(kMwOob) : 0u;
  int cls = BH8_CLASS_BACKGROUND, key = -1;
  if (hit_obj >= 0) {
    cls = f.obj[hit_obj].cls;
    key = f.obj[hit_obj].key;
  }
  store_pixel(f, out, sh_rgba, sh_red, tid, lane, slot, x0, y0, x, y, inside, bgr, oob, cls, key, steps,
              n.n_iter * UPV, n.n_pass, n.n_test);
}

template <int NN, bool STATS, int UPV>
__global__ void __launch_bounds__(kThreads, MinBlocks<NN>::value)
bh8_render_kernel(const __grid_constant__ Bh8Frame f, const __grid_constant__ Bh8Tex tex, const Bh8Out out) {
  render_tile<NN, STATS, UPV>(f, tex, out);
}

#else  // BH8_PERSISTENT_WARPS
// A/B VARIANT, not the shipped configuration (profiles/r02_refill_ab.json): persistent
// warps.  The grid is one wave of CTAs (SMs x BH8_MIN_BLOCKS); every warp draws 8x4-pixel patches from a
// device-wide ticket counter (out.sched[0]) until the frame has none left, so no warp slot idles while
// the slowest warp of a CTA finishes and the tail of the frame is one patch long.  Patches are numbered
// tile by tile.  The warp that leaves last resets the counters (out.sched[1] counts leavers).  Measured
// on B200: 8 % SLOWER at 1080p / nstep 20 (warps of a CTA no longer run the same phase at the same time),
// 1.5 % faster at 8K / nstep 200.
template <int NN, bool STATS, int UPV>
__device__ __forceinline__ void render_warps(const Bh8Frame& f, const Bh8Tex& tex, const Bh8Out& out) {
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  uint32_t sc_addr;
  const Mail mail = make_mail<NN>(f, tid, lane, warp, &sc_addr);
  const DeviceFetch fetch{tex};

  // per-warp totals for the optional counters (added to out.stats once, when the warp leaves)
  unsigned long long st_rays = 0, st_steps = 0, st_oob = 0, st_cls0 = 0, st_cls1 = 0, st_cls2 = 0, st_cls3 = 0;
  unsigned st_patches = 0;
  SchedCount n;

  constexpr int kPatchesPerTile = kPatchesAcross * (kTileH / kPatchH);
  for (;;) {
    unsigned patch = 0;
    if (lane == 0) patch = atomicAdd(out.sched, 1u);
    patch = __shfl_sync(0xffffffffu, patch, 0);
    if (patch >= (unsigned)out.n_patches) break;
    const int tile = (int)(patch / kPatchesPerTile), sub = (int)(patch % kPatchesPerTile);
    const int ty = tile / out.tiles_x, tx = tile - ty * out.tiles_x;
    const int x0 = tx * kTileW;
    int y0;  // ty counts tile rows of THIS shard's stripes
    if (f.shard_count > 1) {
      const int tiles_per_stripe = f.stripe_rows / kTileH;
      const int ls = ty / tiles_per_stripe;
      const int within = ty - ls * tiles_per_stripe;
      y0 = (ls * f.shard_count + f.shard_index) * f.stripe_rows + within * kTileH;
    } else {
      y0 = ty * kTileH;
    }
    const int x = x0 + (sub % kPatchesAcross) * kPatchW + (lane % kPatchW);
    const int y = y0 + (sub / kPatchesAcross) * kPatchH + (lane / kPatchW);
    const bool inside = x < f.width && y < f.height;
    if (STATS) ++st_patches;
    trace_patch<NN, STATS, UPV>(f, fetch, mail, sc_addr, x, y, inside, n);

    if (inside) {  // a warp stores its own patch: four 32-byte row segments (4-byte formats)
      const int steps = mail_steps(mail);
      const int hit_obj = mail_hit(mail);
      const uint32_t bgr = hit_obj >= 0 ? (uint32_t)mail.get_w(kMwBgr) : 0u;
      int cls = BH8_CLASS_BACKGROUND, key = -1;
      if (hit_obj >= 0) {
        cls = f.obj[hit_obj].cls;
        key = f.obj[hit_obj].key;
      }
      const size_t gi = (size_t)y * f.width + x;
      if (f.pixel_format == BH8_PIXEL_BGR8) {
        uint8_t* d = out.pixels + gi * 3;
        d[0] = (uint8_t)bgr;
        d[1] = (uint8_t)(bgr >> 8);
        d[2] = (uint8_t)(bgr >> 16);
      } else {
        const uint32_t px4 = (f.pixel_format == BH8_PIXEL_BGRA8)
                                 ? (bgr | 0xFF000000u)
                                 : (((bgr >> 16) & 0xFFu) | (bgr & 0xFF00u) | ((bgr & 0xFFu) << 16) | 0xFF000000u);
        reinterpret_cast<uint32_t*>(out.pixels)[gi] = px4;
      }
      if (out.cls) out.cls[gi] = (uint8_t)cls;
      if (out.key) out.key[gi] = (int8_t)key;
      if (out.steps) out.steps[gi] = (uint16_t)steps;
      if (STATS) {
        ++st_rays;
        st_steps += (unsigned)steps;
        st_oob += hit_obj >= 0 ? (uint32_t)mail.get_w(kMwOob) : 0u;
        st_cls0 += cls == 0;
        st_cls1 += cls == 1;
        st_cls2 += cls == 2;
        st_cls3 += cls == 3;
      }
    }
  }
  if (STATS) {
    unsigned long long v[kStatSlots] = {st_rays, st_steps, st_cls0, st_cls1, st_cls2, st_cls3, st_oob,
                                        lane == 0 ? st_patches : 0ull,
                                        lane == 0 ? (unsigned long long)n.n_iter * UPV : 0ull,
                                        lane == 0 ? n.n_pass : 0ull, n.n_test};
#pragma unroll
    for (int k = 0; k < kStatSlots; ++k) {
      unsigned long long t = v[k];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
      if (lane == 0 && t) atomicAdd(&out.stats[k], t);
    }
  }
  if (lane == 0) {
    __threadfence();
    const unsigned total = gridDim.x * (kThreads / 32);
    if (atomicAdd(out.sched + 1, 1u) == total - 1u) {  // every warp has made its last (failing) draw
      out.sched[0] = 0u;
      out.sched[1] = 0u;
      __threadfence();
    }
  }
}

template <int NN, bool STATS, int UPV>
__global__ void __launch_bounds__(kThreads, MinBlocks<NN>::value)
bh8_render_kernel(const __grid_constant__ Bh8Frame f, const __grid_constant__ Bh8Tex tex, const Bh8Out out) {
  render_warps<NN, STATS, UPV>(f, tex, out);
}
#endif  // BH8_PERSISTENT_WARPS

// Flat-space tracer (BH8_TRACER_LINEAR): one thread per pixel, at most linear_steps segment tests,
// same tile mapping, colour and store path as the geodesic kernel.
__device__ __forceinline__ void linear_tile(const Bh8Frame& f, const Bh8Tex& tex, const Bh8Out& out) {
  __shared__ __align__(16) uint32_t sh_rgba[kThreads];
  __shared__ unsigned long long sh_red[kStatSlots];
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const int x0 = blockIdx.x * kTileW;
  int y0;
  if (f.shard_count > 1) {
    const int tiles_per_stripe = f.stripe_rows / kTileH;
    const int ls = blockIdx.y / tiles_per_stripe;
    const int within = blockIdx.y - ls * tiles_per_stripe;
    y0 = (ls * f.shard_count + f.shard_index) * f.stripe_rows + within * kTileH;
  } else {
    y0 = blockIdx.y * kTileH;
  }
  if (y0 >= f.height) return;
  const int px = (warp % kPatchesAcross) * kPatchW + (lane % kPatchW);
  const int py = (warp / kPatchesAcross) * kPatchH + (lane / kPatchW);
  const int slot = py * kTileW + px;
  const int x = x0 + px, y = y0 + py;
  const bool inside = x < f.width && y < f.height;

  int hit = -1, steps = 0;
  double hp[3] = {0.0, 0.0, 0.0};
  if (inside) steps = trace_linear(f, x, y, &hit, hp);
  uint32_t bgr = 0, oob = 0;
  int cls = BH8_CLASS_BACKGROUND, key = -1;
  if (hit >= 0) {
    bgr = shade(f, hit, hp, DeviceFetch{tex}, &oob);
    cls = f.obj[hit].cls;
    key = f.obj[hit].key;
  }
  store_pixel(f, out, sh_rgba, sh_red, tid, lane, slot, x0, y0, x, y, inside, bgr, oob, cls, key, steps);
}

__global__ void __launch_bounds__(kThreads, 4)
bh8_linear_kernel(const __grid_constant__ Bh8Frame f, const __grid_constant__ Bh8Tex tex, const Bh8Out out) {
  linear_tile(f, tex, out);
}

// Precision study / stepping yardstick: the bare geodesic update chain for `updates` steps per thread --
// in FP64 the very operations of lane_advance() and lane_update()'s tests (11 FP64 instructions + MUFU.RSQ64H,
// the two triggers compared on high words), or the same chain in FP32 (FFMA + MUFU.RSQ).  Nothing else of
// the renderer is in here: it isolates what the stepping loop costs and what the choice of precision can
// buy for it.  Each thread gets its own impact parameter so the compiler cannot share
// work; the result is folded into a checksum the host ignores.
template <typename Real>
__global__ void __launch_bounds__(256, 5)
bh8_stepping_probe_kernel(double* sink, int updates, double two_m, double u0, double du, double binv2_0) {
  const int gid = blockIdx.x * blockDim.x + threadIdx.x;
  Real u = (Real)u0, phi = 0, dphi_prev = 0;
  const Real delta = (Real)(du * (1.0 + 1e-4 * (gid & 1023))), du_h = (Real)0.5 * delta;
  const Real binv2 = (Real)(binv2_0 * (1.0 + 1e-3 * (gid & 255))), tm = (Real)two_m;
  int parked = 0;
  if (sizeof(Real) == 8) {
    StepConst sc;
    sc.two_m = two_m;
    sc.k375 = 0.375;
    asm volatile("" : "+d"(sc.two_m), "+d"(sc.k375));  // registers, as StepConst::load_shared leaves them
    const uint32_t t_thr = step_turn_threshold((double)du_h), trig_hi = trig_word(1e30);
    double ud = (double)u, phid = 0.0, prev = 0.0;
    for (int i = 0; i < updates; ++i) {
      ud += (double)delta;
      const double dphi = fast_rsqrt(fma(ud * ud, fma(sc.two_m, ud, -1.0), (double)binv2), sc.k375);
      const double s = prev + dphi;
      prev = dphi;
      phid = fma(s, (double)du_h, phid);
      if (hi_word(s) >= t_thr || hi_word(phid) >= trig_hi) ++parked;
    }
    phi = (Real)phid;
  } else {
    const Real trig = (Real)1e30;
    for (int i = 0; i < updates; ++i) {
      u += delta;
      const Real g = fma(u * u, fma(tm, u, (Real)-1), binv2);
      const Real dphi = (Real)rsqrtf((float)g);
      const Real s = dphi_prev + dphi;
      dphi_prev = dphi;
      phi = fma(s, du_h, phi);
      if (!(s * du_h <= (Real)1) || !(phi < trig)) ++parked;
    }
  }
  if (phi == (Real)123.456 || parked == updates + 1) sink[0] = (double)phi + parked;
}

// FP64 pipe peak: 8 independent DFMA chains per thread, enough warps to fill every SM.  The addend is an
// immediate: a DFMA whose three sources are all registers is bound by operand reads, not by the pipe, on B200
// (2.18 SMSP-cycles per warp instruction with the reuse cache, 3.0 with three fresh registers, 2.01 with two:
// tools/exp_fp64_operands.cu, profiles/r02_fp64_operand_mix.txt) -- the yardstick must not be.
__global__ void __launch_bounds__(256, 4) bh8_dfma_peak_kernel(double* sink, int iters, double a, double b) {
  double x0 = threadIdx.x * 1e-9 + b, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6,
         x7 = x0 + 7;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      x0 = fma(x0, a, 0.5);
      x1 = fma(x1, a, 0.5);
      x2 = fma(x2, a, 0.5);
      x3 = fma(x3, a, 0.5);
      x4 = fma(x4, a, 0.5);
      x5 = fma(x5, a, 0.5);
      x6 = fma(x6, a, 0.5);
      x7 = fma(x7, a, 0.5);
    }
  }
  const double s = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
  if (s == 123.456) sink[0] = s;  // never true: keeps the chains alive
}

}  // namespace bh8

#endif  // BH8_KERNEL_CUH_
