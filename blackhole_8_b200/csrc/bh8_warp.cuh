// bh8_warp.cuh -- what a warp of the render kernel does between two rounds of votes, written so that the
// kernel (bh8_kernel.cuh) and the CPU harness (tests/host_harness) run the SAME code: the decision taken
// from the OR of the lanes' states, and the exact test of a parked lane with its register parking.
// The harness emulates a warp as 32 lanes stepped in lockstep (tests/test_ray_math_host.py), so the
// batching schedule -- lanes frozen while others travel, tests run together, lanes thawed or ended -- is
// checked against the reference's frames without a GPU.
#ifndef BH8_WARP_CUH_
#define BH8_WARP_CUH_

#include "bh8_ray.cuh"

namespace bh8 {

// Register budget.  The stepping loop needs ~40 registers, the exact segment test ~110.  Letting the
// second set the kernel's register count would halve the occupancy, so (a) everything only the rare
// paths use lives in the ray's mailbox in shared memory to begin with (bh8::Mail), and (b) a lane
// that enters the exact test first parks the few stepping values it holds in registers and
// re-loads them afterwards: no stepping value is live inside the test, and the test is entered
// about once per ray.  (volatile: the compiler must not forward the stores to the loads, which
// would keep the values alive in registers.)
// Mailbox, stride kThreads: bh8::Mail's slots, then doubles u, phi, dphi_prev, binv2 and ints i,
// state, lo of a parked lane (binv2 and lo are written once: a frozen lane always has its base lo).
constexpr int kMailDoubles = kMailDoublesRay + 4;
constexpr int kMailInts = kMailIntsRay + 3;
enum : int { kKdU = kMailDoublesRay, kKdPhi, kKdDphi, kKdBinv2 };
enum : int { kKwI = kMailIntsRay, kKwState, kKwLo };

template <int NN>
BH8_HD void lane_park_constants(const Lane<NN>& L, const Mail m) {
  m.set_d(kKdBinv2, L.binv2);
  m.set_w(kKwLo, L.lo);
}

// What a frozen lane still carries in registers and the exact test needs or changes.
template <int NN>
BH8_HD void lane_park(const Lane<NN>& L, const Mail m) {
  m.set_d(kKdU, L.u);
  m.set_d(kKdPhi, L.phi);
  m.set_d(kKdDphi, L.dphi_prev);
  m.set_w(kKwI, L.idx());
  m.set_w(kKwState, L.state);
}

// Re-load a lane from its mailbox: frozen (as lane_freeze leaves it) or, if the exact test cleared
// the segment (state kRun), travelling again with the values the test left in Mail's slots.
template <int NN>
BH8_HD void lane_unpark(Lane<NN>& L, const Mail m) {
  L.u = m.get_d(kKdU);
  L.phi = m.get_d(kKdPhi);
  L.dphi_prev = m.get_d(kKdDphi);
  L.binv2 = m.get_d(kKdBinv2);
  L.state = m.get_w(kKwState);
  L.lo = m.get_w(kKwLo);
  L.set_idx(m.get_w(kKwI));
  L.bgr = 0;
  L.oob = 0;
  if (L.state == kRun) {
    lane_thaw(L, m);
  } else {
    L.delta = 0.0;
    L.du_h = 0.0;
    L.trig_hi = kTrigNever;
    L.t_thr = kTrigNever;
    L.span = 0xffffffffu;
    L.inc = 0;
  }
}

// After `updates per vote` straight-line updates the warp ORs its lanes' states (one-hot, see Lane::state)
// and decides, uniformly for all lanes:
enum : int { kWarpStep = 0, kWarpDone = 1, kWarpResolve = 2 };
// Parked exact tests wait for company: until no lane is stepping any more or the oldest has waited
// resolve_wait rounds; then all of them run together.
BH8_HD int warp_decide(unsigned present, int& waited, int resolve_wait) {
  const unsigned runs = present & kRun, pend = present & (kPend | kPendChord);
  if (pend == 0u) return runs == 0u ? kWarpDone : kWarpStep;
  if (runs != 0u && ++waited <= resolve_wait) return kWarpStep;
  waited = 0;
  return kWarpResolve;
}

// The exact test of a lane in kPend / kPendChord.  The test works on its own copy of the lane, loaded
// from the mailbox, and hands the result back through the mailbox (see "Register budget" above).
template <int NN>
BH8_HD void lane_resolve(const Bh8Frame& f, Lane<NN>& L, const Mail m) {
  lane_park(L, m);
  {
    Lane<NN> T;
    lane_unpark(T, m);
    lane_exact(f, T, m);
    if (T.state == kPendChord) lane_exact(f, T, m);  // event right after a cleared segment
    if (T.state == kRun) lane_freeze(T, m, kRun);    // hand the thawed values over through Mail
    lane_park(T, m);
  }
  lane_unpark(L, m);
}

}  // namespace bh8

#endif  // BH8_WARP_CUH_
