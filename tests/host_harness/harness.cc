// TEST-ONLY harness: runs the per-ray code of blackhole_8_b200/csrc/bh8_ray.cuh (the exact source
// the CUDA kernel inlines) on the CPU, pixel by pixel, so the algorithmic restructuring (closed-form
// ray setup, phi-crossing / distance / horizon filters, acos-free texture index) can be checked
// against the oracle without a GPU.  It is compiled only by tests/test_ray_math_host.py into
// tests/host_harness/_build/; the product library never contains or calls it -- there is no CPU
// rendering path in libbh8.so.
#include <cstdint>
#include <cstdio>
#include <cmath>
#include <cstring>

#define BH8_HOST_COUNTERS 1
static unsigned long long bh8_host_filter_evaluations = 0;  // bumped by side_filter()
// Which parts of lane_update_rare() the updates entered (BH8_TRACE): per lane, and per update slot of the
// emulated warp (a slot counts once however many of its lanes went in -- what the warp pays for).
static unsigned long long g_trace_lane[16] = {0}, g_trace_slot[16] = {0};
static unsigned g_trace_mask[8] = {0};
static int g_trace_k = 0;
static inline void bh8_host_trace(int reason) {
  ++g_trace_lane[reason];
  g_trace_mask[g_trace_k] |= 1u << reason;
}
#include "bh8_ray.cuh"
#include "bh8_warp.cuh"

extern "C" void bh8_harness_take_trace(unsigned long long* per_lane, unsigned long long* per_slot) {
  for (int r = 0; r < 16; ++r) {
    per_lane[r] = g_trace_lane[r];
    per_slot[r] = g_trace_slot[r];
    g_trace_lane[r] = g_trace_slot[r] = 0;
  }
}

extern "C" unsigned long long bh8_harness_take_filter_evaluations() {
  const unsigned long long n = bh8_host_filter_evaluations;
  bh8_host_filter_evaluations = 0;
  return n;
}

struct HarnessTexture {
  const uint8_t* bgr;
  int32_t rows, cols;
};

struct HostFetch {
  const HarnessTexture* tex;
  uint32_t operator()(int slot, int col, int row) const {
    const uint8_t* p = tex[slot].bgr + (static_cast<size_t>(row) * tex[slot].cols + col) * 3;
    return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16);
  }
};

static int g_frozen_updates = 0;
extern "C" void bh8_harness_set_frozen_updates(int n) { g_frozen_updates = n; }

template <int NN>
static void trace_frame(const Bh8Frame& f, const HostFetch& fetch, uint8_t* out_bgr, uint8_t* out_class,
                        int8_t* out_key, uint16_t* out_steps, uint64_t* counters) {
  for (int y = 0; y < f.height; ++y) {
    for (int x = 0; x < f.width; ++x) {
      bh8::Lane<NN> L;
      double md[bh8::kMailDoubles];
      int32_t mw[bh8::kMailInts];
      const bh8::Mail mail{md, mw, 1};
      bh8::lane_setup(f, x, y, L, mail);
      bh8::lane_park_constants(L, mail);
      while (L.state != bh8::kDead) {  // the kernel's per-lane sequence, one lane, no batching
        if (L.state == bh8::kRun) {
          counters[0]++;
          bh8::lane_update(f, L, mail, bh8::StepConst::load(f));
        } else {
          // In the kernel a frozen lane keeps executing the warp's straight-line updates until the
          // warp attends to it: they must not change anything.
          for (int k = 0; k < g_frozen_updates; ++k) bh8::lane_update(f, L, mail, bh8::StepConst::load(f));
          counters[1]++;
          if (!bh8::lane_resolve(f, L, mail, fetch)) break;  // ended: the lane is not loaded back
        }
      }
      for (int k = 0; k < g_frozen_updates; ++k) bh8::lane_update(f, L, mail, bh8::StepConst::load(f));  // ended rays too
      const int hit_obj = bh8::mail_hit(mail);
      const uint32_t bgr = hit_obj >= 0 ? (uint32_t)mail.get_w(bh8::kMwBgr) : 0u;
      int cls = BH8_CLASS_BACKGROUND, key = -1;
      if (hit_obj >= 0) {
        cls = f.obj[hit_obj].cls;
        key = f.obj[hit_obj].key;
      }
      const size_t i = static_cast<size_t>(y) * f.width + x;
      out_bgr[3 * i] = bgr & 255;
      out_bgr[3 * i + 1] = (bgr >> 8) & 255;
      out_bgr[3 * i + 2] = (bgr >> 16) & 255;
      out_class[i] = (uint8_t)cls;
      out_key[i] = (int8_t)key;
      out_steps[i] = (uint16_t)bh8::mail_steps(mail);
    }
  }
}

// counters[0] = geodesic updates computed, counters[1] = exact segment tests (ray_resolve calls).
// filter_slots < 0 forces the generic instantiation (exact test on every gated step).
extern "C" int bh8_harness_render(const bh8_scene* scene, const bh8_camera* cam, const bh8_params* prm,
                                  const HarnessTexture* textures, int n_textures, int filter_slots,
                                  uint8_t* out_bgr, uint8_t* out_class, int8_t* out_key, uint16_t* out_steps,
                                  uint64_t* counters, char* err) {
  int rows[BH8_MAX_TEXTURES] = {0}, cols[BH8_MAX_TEXTURES] = {0};
  for (int i = 0; i < n_textures && i < BH8_MAX_TEXTURES; ++i) {
    rows[i] = textures[i].rows;
    cols[i] = textures[i].cols;
  }
  Bh8Frame f;
  const int rc = bh8_build_frame(scene, cam, prm, rows, cols, &f, err);
  if (rc != BH8_OK) return rc;
  const HostFetch fetch{textures};
  counters[0] = counters[1] = 0;
  if (f.tracer == BH8_TRACER_LINEAR) {
    for (int y = 0; y < f.height; ++y)
      for (int x = 0; x < f.width; ++x) {
        int hit = -1;
        double hp[3] = {0, 0, 0};
        const int steps = bh8::trace_linear(f, x, y, &hit, hp);
        uint32_t bgr = 0, oob = 0;
        if (hit >= 0) bgr = bh8::shade(f, hit, hp, fetch, &oob);
        const size_t i = static_cast<size_t>(y) * f.width + x;
        out_bgr[3 * i] = bgr & 255;
        out_bgr[3 * i + 1] = (bgr >> 8) & 255;
        out_bgr[3 * i + 2] = (bgr >> 16) & 255;
        out_class[i] = (uint8_t)(hit >= 0 ? f.obj[hit].cls : BH8_CLASS_BACKGROUND);
        out_key[i] = (int8_t)(hit >= 0 ? f.obj[hit].key : -1);
        out_steps[i] = (uint16_t)steps;
        counters[0] += steps;
      }
    return BH8_OK;
  }
  switch ((filter_slots >= 0 && f.n_nc <= 4) ? f.n_nc : -1) {
    case 0: trace_frame<0>(f, fetch, out_bgr, out_class, out_key, out_steps, counters); break;
    case 1: trace_frame<1>(f, fetch, out_bgr, out_class, out_key, out_steps, counters); break;
    case 2: trace_frame<2>(f, fetch, out_bgr, out_class, out_key, out_steps, counters); break;
    case 3: trace_frame<3>(f, fetch, out_bgr, out_class, out_key, out_steps, counters); break;
    case 4: trace_frame<4>(f, fetch, out_bgr, out_class, out_key, out_steps, counters); break;
    default: trace_frame<-1>(f, fetch, out_bgr, out_class, out_key, out_steps, counters); break;
  }
  return BH8_OK;
}

// Debugging aid: ONE pixel, the lane's registers and the filter-(2) words of its mailbox printed after every
// update and every exact pass (tests/debug_pixel.py).
template <int NN>
static void trace_pixel(const Bh8Frame& f, const HostFetch& fetch, int x, int y) {
  bh8::Lane<NN> L;
  double md[bh8::kMailDoubles];
  int32_t mw[bh8::kMailInts];
  std::memset(md, 0, sizeof md);
  std::memset(mw, 0, sizeof mw);
  const bh8::Mail mail{md, mw, 1};
  bh8::lane_setup(f, x, y, L, mail);
  bh8::lane_park_constants(L, mail);
  std::printf("NN %d n_nc %d nstep %d evt_turn %d u0 %.9g u_gate %.9g lease_ku %.6g lease_kphi %.6g max_n %.6g max_c %.6g\n", NN, f.n_nc,
              f.nstep, f.evt_turn, f.u0, f.u_gate, (double)f.lease_ku, (double)f.lease_kphi, (double)f.nc_max_n, (double)f.nc_max_c);
  auto show = [&](const char* what) {
    std::printf("%-7s idx %3d state %d u %.12g phi %.12g delta %.6g du_h %.6g lo %d span %u trig_hi %08x flags %02x "
                "gate_in %d gate_out %d next %d fbits %08x fstep %d\n", what, L.idx(), (int)L.state, L.u, L.phi, L.delta, L.du_h,
                (int)L.lo, (unsigned)L.span, (unsigned)L.trig_hi, (unsigned)mail.get_w(bh8::kMwFlags) & 0xffu,
                NN != 0 ? mail.get_w(bh8::kMwGateIn) : 0, NN != 0 ? mail.get_w(bh8::kMwGateOut) : 0,
                mail.get_w(bh8::kMwNext), (unsigned)mail.get_w(bh8::kMwFbits), mail.get_w(bh8::kMwFstep));
  };
  show("setup");
  for (int guard = 0; guard < 100000 && L.state != bh8::kDead; ++guard) {
    if (L.state == bh8::kRun) {
      bh8::lane_update(f, L, mail, bh8::StepConst::load(f));
      show("update");
    } else {
      const bool alive = bh8::lane_resolve(f, L, mail, fetch);
      show("resolve");
      if (!alive) break;
    }
  }
  std::printf("result: steps %d hit %d\n", bh8::mail_steps(mail), bh8::mail_hit(mail));
}

extern "C" int bh8_harness_trace_pixel(const bh8_scene* scene, const bh8_camera* cam, const bh8_params* prm,
                                       const HarnessTexture* textures, int n_textures, int filter_slots, int x, int y,
                                       char* err) {
  int rows[BH8_MAX_TEXTURES] = {0}, cols[BH8_MAX_TEXTURES] = {0};
  for (int i = 0; i < n_textures && i < BH8_MAX_TEXTURES; ++i) {
    rows[i] = textures[i].rows;
    cols[i] = textures[i].cols;
  }
  Bh8Frame f;
  const int rc = bh8_build_frame(scene, cam, prm, rows, cols, &f, err);
  if (rc != BH8_OK) return rc;
  const HostFetch fetch{textures};
  switch ((filter_slots >= 0 && f.n_nc <= 4) ? f.n_nc : -1) {
    case 0: trace_pixel<0>(f, fetch, x, y); break;
    case 1: trace_pixel<1>(f, fetch, x, y); break;
    case 2: trace_pixel<2>(f, fetch, x, y); break;
    case 3: trace_pixel<3>(f, fetch, x, y); break;
    case 4: trace_pixel<4>(f, fetch, x, y); break;
    default: trace_pixel<-1>(f, fetch, x, y); break;
  }
  std::fflush(stdout);
  return BH8_OK;
}

// Known-answer hook: turning point by the fast path vs the literal bisection, same frame constants.
extern "C" int bh8_harness_solve(double mass, const double* b, int n, double* fast, double* bisect) {
  Bh8Frame f;
  std::memset(&f, 0, sizeof f);
  f.two_m = 2.0 * mass;
  f.inv3m = 1.0 / (3.0 * mass);
  f.nine_m2 = 9.0 * mass * mass;
  const double l = 0.0 + cbrt(2.220446049250313e-16), r = f.inv3m;
  f.bis_mid0 = (l + r) / 2.0;
  double h = (r - l) / 2.0;
  for (int i = 0; i <= BH8_BISECT_ITERS; ++i) {
    h /= 2.0;
    f.bis_h[i] = h;
  }
  f.bis_l0 = l;
  f.bis_grid = (r - l) / 1048576.0;
  f.bis_inv_grid = 1048576.0 / (r - l);
  for (int i = 0; i < n; ++i) {
    const double binv2 = 1.0 / (b[i] * b[i]);
    fast[i] = bh8::solve_turning_point(f, binv2);
    bisect[i] = bh8::solve_turning_point_bisect(f, binv2);
  }
  return 0;
}

extern "C" void bh8_harness_atan2f(const float* y, const float* x, int n, float* out) {
  for (int i = 0; i < n; ++i) out[i] = bh8::fast_atan2f(y[i], x[i]);
}

extern "C" void bh8_harness_sincos(const double* x, int n, double* s, double* c) {
  for (int i = 0; i < n; ++i) bh8::sincos_(x[i], &s[i], &c[i]);
}

// Scripted animation: the code bh8_animate_kernel runs per entity, on the host.
#include <vector>

#include "bh8_anim.cuh"

extern "C" int bh8_harness_replay(const bh8_scene* scene0, const bh8_basis* basis, const bh8_camera* cam0,
                                  const bh8_action* actions, int n_actions, int n_frames, bh8_camera* cams,
                                  bh8_object* objs, char* err) {
  const int n_obj = scene0->n_obj;
  std::vector<Bh8Action> acts(n_actions + 1);
  if (const char* why = bh8a_prepare_actions(actions, n_actions, n_obj, acts.data())) {
    std::strcpy(err, why);
    return BH8_EINVAL;
  }
  std::vector<Bh8Entity> ent(n_obj + 1);
  bh8a_prepare_entities(scene0, basis, cam0, ent.data());
  for (int me = 0; me <= n_obj; ++me)
    bh8a_replay_entity(ent[me], me, n_obj, acts.data(), n_actions, n_frames, scene0->obj, cam0, cams, objs);
  return BH8_OK;
}

// ---- a warp of the kernel, emulated ---------------------------------------------------------------
// 32 lanes = one 8x4-pixel patch, stepped in lockstep exactly as render_tile (bh8_kernel.cuh) does it:
// `updates_per_vote` straight-line updates for EVERY lane (frozen and dead ones included), the OR of
// the lanes' states, bh8::warp_decide, and bh8::lane_exact for ALL lanes (parked or not) -- the same functions the kernel inlines, with a mailbox laid out as in shared
// memory (component c of lane l at [c * 32 + l]).  What the per-lane loop above cannot show -- lanes
// waiting frozen while others travel, batched tests, the register parking -- is exercised here.
template <int NN>
static void trace_frame_warps(const Bh8Frame& f, const HostFetch& fetch, int updates_per_vote, uint8_t* out_bgr,
                              uint8_t* out_class, int8_t* out_key, uint16_t* out_steps, uint64_t* counters) {
  constexpr int kLanes = 32, kPW = 8, kPH = 4;
  for (int py0 = 0; py0 < f.height; py0 += kPH) {
    for (int px0 = 0; px0 < f.width; px0 += kPW) {
      double md[bh8::kMailDoubles * kLanes];
      int32_t mw[bh8::kMailInts * kLanes];
      bh8::Lane<NN> L[kLanes];
      bh8::Mail mail[kLanes];
      bool inside[kLanes];
      for (int l = 0; l < kLanes; ++l) {
        mail[l] = bh8::Mail{md + l, mw + l, kLanes};
        const int x = px0 + (l % kPW), y = py0 + (l / kPW);
        inside[l] = x < f.width && y < f.height;
        bh8::lane_inert(L[l]);
        if (inside[l]) {
          bh8::lane_setup(f, x, y, L[l], mail[l]);
          bh8::lane_park_constants(L[l], mail[l]);
        }
      }
      int waited = 0;
      for (uint64_t round = 0;; ++round) {
        unsigned present = 0;
        for (int l = 0; l < kLanes; ++l) {
          for (int k = 0; k < updates_per_vote; ++k) {
            g_trace_k = k & 7;
            bh8::lane_update(f, L[l], mail[l], bh8::StepConst::load(f));
          }
          present |= (unsigned)L[l].state;
        }
        for (int k = 0; k < 8; ++k) {
          for (int r = 0; r < 16; ++r) g_trace_slot[r] += (g_trace_mask[k] >> r) & 1u;
          g_trace_mask[k] = 0;
        }
        g_trace_k = 0;
        counters[0] += (uint64_t)updates_per_vote;  // update slots of this warp
        const int todo = bh8::warp_decide(present, waited, f.resolve_wait);
        if (todo == bh8::kWarpStep) continue;
        if (todo == bh8::kWarpDone) break;
        counters[1]++;  // resolve passes
        bool alive = false;
        for (int l = 0; l < kLanes; ++l) {
          alive = bh8::lane_resolve(f, L[l], mail[l], fetch) || alive;  // ALL lanes, as in the kernel
        }
        if (!alive) break;  // the usual pass is the last one: no further round (trace_patch)
        if (round > 100000000ull) return;  // a warp that never ends would hang the GPU: leave zeros, the test fails
      }
      for (int l = 0; l < kLanes; ++l) {
        if (!inside[l]) continue;
        const int x = px0 + (l % kPW), y = py0 + (l / kPW);
        const int hit_obj = bh8::mail_hit(mail[l]);
        const uint32_t bgr = hit_obj >= 0 ? (uint32_t)mail[l].get_w(bh8::kMwBgr) : 0u;
        const size_t i = static_cast<size_t>(y) * f.width + x;
        out_bgr[3 * i] = bgr & 255;
        out_bgr[3 * i + 1] = (bgr >> 8) & 255;
        out_bgr[3 * i + 2] = (bgr >> 16) & 255;
        out_class[i] = (uint8_t)(hit_obj >= 0 ? f.obj[hit_obj].cls : BH8_CLASS_BACKGROUND);
        out_key[i] = (int8_t)(hit_obj >= 0 ? f.obj[hit_obj].key : -1);
        out_steps[i] = (uint16_t)bh8::mail_steps(mail[l]);
      }
    }
  }
}

// counters[0] = update slots summed over warps (each slot = one update of all 32 lanes), counters[1] =
// resolve passes summed over warps.
extern "C" int bh8_harness_render_warps(const bh8_scene* scene, const bh8_camera* cam, const bh8_params* prm,
                                        const HarnessTexture* textures, int n_textures, int filter_slots,
                                        int updates_per_vote, int resolve_wait, uint8_t* out_bgr, uint8_t* out_class,
                                        int8_t* out_key, uint16_t* out_steps, uint64_t* counters, char* err) {
  int rows[BH8_MAX_TEXTURES] = {0}, cols[BH8_MAX_TEXTURES] = {0};
  for (int i = 0; i < n_textures && i < BH8_MAX_TEXTURES; ++i) {
    rows[i] = textures[i].rows;
    cols[i] = textures[i].cols;
  }
  Bh8Frame f;
  const int rc = bh8_build_frame(scene, cam, prm, rows, cols, &f, err);
  if (rc != BH8_OK) return rc;
  if (f.tracer == BH8_TRACER_LINEAR) {
    std::strcpy(err, "the warp emulation is for the geodesic kernel");
    return BH8_EINVAL;
  }
  f.resolve_wait = resolve_wait;
  const HostFetch fetch{textures};
  counters[0] = counters[1] = 0;
  const int u = updates_per_vote;
  switch ((filter_slots >= 0 && f.n_nc <= 4) ? f.n_nc : -1) {
    case 0: trace_frame_warps<0>(f, fetch, u, out_bgr, out_class, out_key, out_steps, counters); break;
    case 1: trace_frame_warps<1>(f, fetch, u, out_bgr, out_class, out_key, out_steps, counters); break;
    case 2: trace_frame_warps<2>(f, fetch, u, out_bgr, out_class, out_key, out_steps, counters); break;
    case 3: trace_frame_warps<3>(f, fetch, u, out_bgr, out_class, out_key, out_steps, counters); break;
    case 4: trace_frame_warps<4>(f, fetch, u, out_bgr, out_class, out_key, out_steps, counters); break;
    default: trace_frame_warps<-1>(f, fetch, u, out_bgr, out_class, out_key, out_steps, counters); break;
  }
  return BH8_OK;
}
