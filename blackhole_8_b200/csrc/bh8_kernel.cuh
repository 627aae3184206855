// bh8_kernel.cuh -- the render kernel: one thread per ray, one 32x8-pixel tile per 256-thread CTA.
//
// Replaces the pixel loop of blackhole_solution_test.cc:162-298 (ray setup, inbound steps, 0.9-step,
// captured chord, outbound steps, colour, frame store) for a whole frame or a set of row stripes.
//
// CTA schedule
//   A  every thread sets up its ray and runs the inbound leg (nstep-1 steps of +du, one of +0.9du).
//      A warp is an 8x4-pixel patch (neighbouring rays end at neighbouring steps); the leg loop
//      leaves as soon as __any_sync() reports no live lane, so finished patches cost nothing.
//   B  captured rays (b < b_c) take their straight chord to the centre and finish.
//   C  mid-ray compaction: the survivors' states are packed through shared memory with
//      ballot/popc ranks so the outbound leg runs in dense warps -- rays that ended on the way in
//      (disc hits, the horizon) no longer hold lanes of a half-empty warp.
//   D  outbound leg (nstep-1 steps of -du) on the packed rays; hits go back to the pixel's slot.
//   E  every thread colours its own pixel slot (texture object fetch / chess / black) into a shared
//      RGBA tile, which leaves as 16-byte coalesced stores (4 pixels per store, 128 B per row).
#ifndef BH8_KERNEL_CUH_
#define BH8_KERNEL_CUH_

#include <cuda_runtime.h>

#include "bh8_ray.cuh"

namespace bh8 {

constexpr int kTileW = 32;
constexpr int kTileH = 8;
constexpr int kThreads = kTileW * kTileH;
constexpr int kWarps = kThreads / 32;
constexpr int kStateDoubles = 12;

struct Bh8Tex {
  unsigned long long obj[BH8_MAX_TEXTURES];
};

struct Bh8Out {
  uint8_t* pixels;
  uint8_t* cls;
  int8_t* key;
  uint16_t* steps;
  unsigned long long* stats;  // [0] rays [1] steps [2..5] class counts [6] tex_oob
  int32_t vec_ok;             // 16-byte row stores are legal (width % 4 == 0, base 16-byte aligned)
};

struct TileShared {
  double hp[3][kThreads];              // hit point per pixel slot
  double st[kStateDoubles][kThreads];  // packed ray states (compaction)
  uint32_t st_mask[kThreads];
  uint32_t rgba[kThreads];
  uint16_t st_pix[kThreads];
  uint16_t st_meta[kThreads];  // flags
  uint16_t st_steps[kThreads];
  uint16_t steps[kThreads];
  int8_t hobj[kThreads];
  int32_t warp_cnt[kWarps];
  unsigned long long red[7];
};

struct DeviceFetch {
  const Bh8Tex& tex;
  __device__ __forceinline__ uint32_t operator()(int slot, int col, int row) const {
    const uchar4 t = tex2D<uchar4>(tex.obj[slot], col + 0.5f, row + 0.5f);
    return (uint32_t)t.x | ((uint32_t)t.y << 8) | ((uint32_t)t.z << 16);
  }
};

__device__ __forceinline__ void record_hit(TileShared& sh, int slot, const Hit& h, int steps) {
  sh.hobj[slot] = (int8_t)h.obj;
  sh.steps[slot] = (uint16_t)steps;
  if (h.obj >= 0) {
    sh.hp[0][slot] = h.p[0];
    sh.hp[1][slot] = h.p[1];
    sh.hp[2][slot] = h.p[2];
  }
}

template <bool kCompact>
__global__ void __launch_bounds__(kThreads, 2)
bh8_render_kernel(const __grid_constant__ Bh8Frame f, const __grid_constant__ Bh8Tex tex, const Bh8Out out) {
  __shared__ TileShared sh;
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

  // warp = 8x4 patch; slot = row-major position in the 32x8 tile
  const int px = (warp & 3) * 8 + (lane & 7);
  const int py = (warp >> 2) * 4 + (lane >> 3);
  const int slot = py * kTileW + px;
  const int x = x0 + px, y = y0 + py;
  const bool inside = x < f.width && y < f.height;

  // ---- A: setup + inbound leg ------------------------------------------------------------------
  Ray r;
  Hit h;
  h.obj = -1;
  bool alive = inside;
  r.steps = 0;
  r.flags = 0;
  if (inside) {
    ray_setup(f, x, y, r);
    if (r.flags & kDegenerate) {
      ray_degenerate(f, r, h);
      alive = false;
    }
  }
  const int nsafe = f.nstep - 1;
  for (int i = 0; i < f.nstep; ++i) {
    if (!__any_sync(0xffffffffu, alive)) break;
    if (alive) {
      const double delta = (i < nsafe) ? r.du : r.du * 0.9;  // :218 / :241
      if (ray_step(f, r, delta, h)) alive = false;
    }
  }
  // ---- B: captured rays ------------------------------------------------------------------------
  if (alive && (r.flags & kCaptured)) {
    ray_chord(f, r, h);
    alive = false;
  }

  if (!kCompact) {
    // ---- D (in place) ---------------------------------------------------------------------------
    for (int i = 0; i < nsafe; ++i) {
      if (!__any_sync(0xffffffffu, alive)) break;
      if (alive) {
        if (ray_step(f, r, -r.du, h)) alive = false;  // :275
      }
    }
    record_hit(sh, slot, h, r.steps);
  } else {
    // ---- C: pack survivors -----------------------------------------------------------------------
    if (!alive) record_hit(sh, slot, h, r.steps);
    const unsigned bal = __ballot_sync(0xffffffffu, alive);
    if (lane == 0) sh.warp_cnt[warp] = __popc(bal);
    __syncthreads();
    int base = 0, total = 0;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) {
      const int c = sh.warp_cnt[w];
      if (w < warp) base += c;
      total += c;
    }
    if (alive) {
      const int dst = base + __popc(bal & ((1u << lane) - 1u));
      sh.st[0][dst] = r.u;
      sh.st[1][dst] = r.phi;
      sh.st[2][dst] = r.dphi_prev;
      sh.st[3][dst] = r.du;
      sh.st[4][dst] = r.binv2;
      sh.st[5][dst] = r.phi_next;
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        sh.st[6 + i][dst] = r.yv[i];
        sh.st[9 + i][dst] = r.xv[i];
      }
      sh.st_mask[dst] = r.mask;
      sh.st_pix[dst] = (uint16_t)slot;
      sh.st_meta[dst] = (uint16_t)r.flags;
      sh.st_steps[dst] = (uint16_t)r.steps;
    }
    __syncthreads();
    // ---- D: outbound leg on packed rays ------------------------------------------------------------
    alive = tid < total;
    int pix = 0;
    if (alive) {
      r.u = sh.st[0][tid];
      r.phi = sh.st[1][tid];
      r.dphi_prev = sh.st[2][tid];
      r.du = sh.st[3][tid];
      r.binv2 = sh.st[4][tid];
      r.phi_next = sh.st[5][tid];
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        r.yv[i] = sh.st[6 + i][tid];
        r.xv[i] = sh.st[9 + i][tid];
      }
      r.mask = sh.st_mask[tid];
      pix = sh.st_pix[tid];
      r.flags = sh.st_meta[tid];
      r.steps = sh.st_steps[tid];
      h.obj = -1;
    }
    const bool mine = alive;
    for (int i = 0; i < nsafe; ++i) {
      if (!__any_sync(0xffffffffu, alive)) break;
      if (alive) {
        if (ray_step(f, r, -r.du, h)) alive = false;  // :275
      }
    }
    if (mine) record_hit(sh, pix, h, r.steps);
  }
  __syncthreads();

  // ---- E: colour own slot (slot == tid in row-major tile order) ----------------------------------
  const int cx = x0 + (tid & 31), cy = y0 + (tid >> 5);
  const bool cin = cx < f.width && cy < f.height;
  const int ho = sh.hobj[tid];
  uint32_t bgr = 0, oob = 0;
  int cls = BH8_CLASS_BACKGROUND, key = -1;
  if (cin && ho >= 0) {
    const double p[3] = {sh.hp[0][tid], sh.hp[1][tid], sh.hp[2][tid]};
    bgr = shade(f, ho, p, DeviceFetch{tex}, &oob);
    cls = f.obj[ho].cls;
    key = f.obj[ho].key;
  }
  const size_t gi = (size_t)cy * f.width + cx;
  if (f.pixel_format == BH8_PIXEL_BGR8) {
    if (cin) {
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
      sh.rgba[tid] = px4;
      __syncthreads();
      if (tid < kThreads / 4) {
        const int row = tid >> 3, col = (tid & 7) * 4;
        if (x0 + col < f.width && y0 + row < f.height) {
          const uint4 v = *reinterpret_cast<const uint4*>(&sh.rgba[row * kTileW + col]);
          *reinterpret_cast<uint4*>(out.pixels + ((size_t)(y0 + row) * f.width + x0 + col) * 4) = v;
        }
      }
    } else if (cin) {
      reinterpret_cast<uint32_t*>(out.pixels)[gi] = px4;
    }
  }
  if (cin) {
    if (out.cls) out.cls[gi] = (uint8_t)cls;
    if (out.key) out.key[gi] = (int8_t)key;
    if (out.steps) out.steps[gi] = sh.steps[tid];
  }

  // ---- optional counters ----------------------------------------------------------------------------
  if (f.flags & BH8_FLAG_STATS) {
    if (tid < 7) sh.red[tid] = 0ull;
    __syncthreads();
    unsigned long long v[7];
    v[0] = cin ? 1ull : 0ull;
    v[1] = cin ? (unsigned long long)sh.steps[tid] : 0ull;
    for (int c = 0; c < 4; ++c) v[2 + c] = (cin && cls == c) ? 1ull : 0ull;
    v[6] = oob;
#pragma unroll
    for (int k = 0; k < 7; ++k) {
      unsigned long long s = v[k];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (lane == 0 && s) atomicAdd(&sh.red[k], s);
    }
    __syncthreads();
    if (tid < 7 && sh.red[tid]) atomicAdd(&out.stats[tid], sh.red[tid]);
  }
}

// FP64 pipe peak: 8 independent DFMA chains per thread, enough warps to fill every SM.
__global__ void __launch_bounds__(256, 4) bh8_dfma_peak_kernel(double* sink, int iters, double a, double b) {
  double x0 = threadIdx.x * 1e-9, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6,
         x7 = x0 + 7;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      x0 = fma(x0, a, b);
      x1 = fma(x1, a, b);
      x2 = fma(x2, a, b);
      x3 = fma(x3, a, b);
      x4 = fma(x4, a, b);
      x5 = fma(x5, a, b);
      x6 = fma(x6, a, b);
      x7 = fma(x7, a, b);
    }
  }
  const double s = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
  if (s == 123.456) sink[0] = s;  // never true: keeps the chains alive
}

}  // namespace bh8

#endif  // BH8_KERNEL_CUH_
