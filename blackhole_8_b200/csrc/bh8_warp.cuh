// bh8_warp.cuh -- what a warp of the render kernel does between two rounds of votes, written so that the
// kernel (bh8_kernel.cuh) and the CPU harness (tests/host_harness) run the SAME code: the decision taken
// from the OR of the lanes' states (the exact test of the parked lanes is bh8::lane_exact, bh8_ray.cuh).
// The harness emulates a warp as 32 lanes stepped in lockstep (tests/test_ray_math_host.py), so the
// batching schedule -- lanes frozen while others travel, tests run together, lanes thawed or ended -- is
// checked against the reference's frames without a GPU.
#ifndef BH8_WARP_CUH_
#define BH8_WARP_CUH_

#include "bh8_ray.cuh"

namespace bh8 {

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

}  // namespace bh8

#endif  // BH8_WARP_CUH_
