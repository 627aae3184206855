// Experiment (not part of libbh8.so): what bounds the bare FP64 geodesic update chain on B200 -- the rsqrt seed
// (MUFU.RSQ64H vs an FP32 seed), latency (two rays per thread, two consecutive updates computed together) or
// the FP64 pipe.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o build/exp/stepping_chain tools/exp_stepping_chain.cu
// Result (profiles/r02_stepping_chain_variants.txt): ~30 SMSP-cycles per warp-update whatever the variant: the chain is
// bound by the FP64 pipe (11 FP64 instructions at ~2.6 cycles each in this operand mix), not by latency or the seed.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t hi_word(double x) { return (uint32_t)__double2hiint(x); }
template <int SEED> __device__ __forceinline__ double seed(double x) {
  if (SEED == 0) { double y; asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x)); return y; }
  if (SEED == 1) { float xf = (float)x, yf; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(yf) : "f"(xf)); return (double)yf; }
  if (SEED == 2) {  // bit tricks: double -> float by hand (normal range), MUFU.RSQ, float -> double by hand
    const uint32_t h = (uint32_t)__double2hiint(x), l = (uint32_t)__double2loint(x);
    const uint32_t fb = ((h - (896u << 20)) << 3) | (l >> 29);
    float yf; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(yf) : "f"(__uint_as_float(fb)));
    const uint32_t yb = __float_as_uint(yf);
    return __hiloint2double((int)((yb >> 3) + (896u << 20)), (int)(yb << 29));
  }
  if (SEED == 3) { return x; }  // no seed at all (lower bound: refinement only)
  return x;
}
template <int SEED>
__global__ void __launch_bounds__(256, 5) probe(double* sink, int updates, double two_m_, double u0, double du, double binv2_0) {
  const int gid = blockIdx.x * blockDim.x + threadIdx.x;
  const double delta = du * (1.0 + 1e-4 * (gid & 1023)), du_h = 0.5 * delta;
  const double binv2 = binv2_0 * (1.0 + 1e-3 * (gid & 255));
  double two_m = two_m_, k375 = 0.375;
  asm volatile("" : "+d"(two_m), "+d"(k375));
  const uint32_t t_thr = 0x7ff00000u, trig_hi = 0x7ff00000u;
  double ud = u0, phid = 0.0, prev = 0.0; int parked = 0;
  for (int i = 0; i < updates; ++i) {
    ud += delta;
    const double x = fma(ud * ud, fma(two_m, ud, -1.0), binv2);
    const double y = seed<SEED>(x);
    const double e = fma(-(x * y), y, 1.0);
    const double dphi = fma(y * e, fma(e, k375, 0.5), y);
    const double s = prev + dphi;
    prev = dphi;
    phid = fma(s, du_h, phid);
    if (hi_word(s) >= t_thr || hi_word(phid) >= trig_hi) ++parked;
  }
  if (phid == 123.456 || parked == updates + 1) sink[0] = phid + parked;
}
// two independent rays per thread, their updates interleaved by hand (ILP 2)
__global__ void __launch_bounds__(256, 5) probe2(double* sink, int updates, double two_m_, double u0, double du, double binv2_0) {
  const int gid = blockIdx.x * blockDim.x + threadIdx.x;
  const double delta = du * (1.0 + 1e-4 * (gid & 1023)), du_h = 0.5 * delta;
  const double binv2 = binv2_0 * (1.0 + 1e-3 * (gid & 255)), binv2b = binv2 * 1.01;
  double two_m = two_m_, k375 = 0.375;
  asm volatile("" : "+d"(two_m), "+d"(k375));
  const uint32_t t_thr = 0x7ff00000u, trig_hi = 0x7ff00000u;
  double ua = u0, pa = 0.0, qa = 0.0, ub = u0 * 1.001, pb = 0.0, qb = 0.0; int parked = 0;
  for (int i = 0; i < updates / 2; ++i) {
    ua += delta; ub += delta;
    const double xa = fma(ua * ua, fma(two_m, ua, -1.0), binv2), xb = fma(ub * ub, fma(two_m, ub, -1.0), binv2b);
    const double ya = seed<0>(xa), yb = seed<0>(xb);
    const double ea = fma(-(xa * ya), ya, 1.0), eb = fma(-(xb * yb), yb, 1.0);
    const double da = fma(ya * ea, fma(ea, k375, 0.5), ya), db = fma(yb * eb, fma(eb, k375, 0.5), yb);
    const double sa = qa + da, sb = qb + db;
    qa = da; qb = db;
    pa = fma(sa, du_h, pa); pb = fma(sb, du_h, pb);
    if (hi_word(sa) >= t_thr || hi_word(pa) >= trig_hi) ++parked;
    if (hi_word(sb) >= t_thr || hi_word(pb) >= trig_hi) ++parked;
  }
  if (pa + pb == 123.456 || parked == updates + 1) sink[0] = pa + pb + parked;
}
// the same ray, two CONSECUTIVE updates computed together (u_{i+2} does not wait for dphi_{i+1})
__global__ void __launch_bounds__(256, 5) probe_pair(double* sink, int updates, double two_m_, double u0, double du, double binv2_0) {
  const int gid = blockIdx.x * blockDim.x + threadIdx.x;
  const double delta = du * (1.0 + 1e-4 * (gid & 1023)), du_h = 0.5 * delta;
  const double binv2 = binv2_0 * (1.0 + 1e-3 * (gid & 255));
  double two_m = two_m_, k375 = 0.375;
  asm volatile("" : "+d"(two_m), "+d"(k375));
  const uint32_t t_thr = 0x7ff00000u, trig_hi = 0x7ff00000u;
  double u = u0, phi = 0.0, prev = 0.0; int parked = 0;
  for (int i = 0; i < updates / 2; ++i) {
    const double ua = u + delta, ub = ua + delta;
    const double xa = fma(ua * ua, fma(two_m, ua, -1.0), binv2), xb = fma(ub * ub, fma(two_m, ub, -1.0), binv2);
    const double ya = seed<0>(xa), yb = seed<0>(xb);
    const double ea = fma(-(xa * ya), ya, 1.0), eb = fma(-(xb * yb), yb, 1.0);
    const double da = fma(ya * ea, fma(ea, k375, 0.5), ya), db = fma(yb * eb, fma(eb, k375, 0.5), yb);
    const double sa = prev + da, sb = da + db;
    const double pa = fma(sa, du_h, phi), pb = fma(sb, du_h, pa);
    asm volatile("" ::: "memory");
    const bool ra = hi_word(sa) >= t_thr || hi_word(pa) >= trig_hi, rb = hi_word(sb) >= t_thr || hi_word(pb) >= trig_hi;
    if (ra || rb) ++parked;
    u = ub; phi = pb; prev = db;
  }
  if (phi == 123.456 || parked == updates + 1) sink[0] = phi + parked;
}

// Operand-mix variants of the update (tools/exp_fp64_operands.cu: a DFMA whose three sources are all fresh registers
// costs more than 2 SMSP-cycles).  FORM bit 0: the rsqrt correction as y * (1 + e p) -- DFMA(reg, reg, imm) + DMUL --
// instead of DFMA(y e, p, y); bit 1: the angle kept as the plain sum of s (DADD), multiplied by du/2 on demand.
template <int FORM>
__global__ void __launch_bounds__(256, 5) probe_form(double* sink, int updates, double two_m_, double u0, double du, double binv2_0) {
  const int gid = blockIdx.x * blockDim.x + threadIdx.x;
  const double delta = du * (1.0 + 1e-4 * (gid & 1023)), du_h = 0.5 * delta;
  const double binv2 = binv2_0 * (1.0 + 1e-3 * (gid & 255));
  double two_m = two_m_, k375 = 0.375;
  asm volatile("" : "+d"(two_m), "+d"(k375));
  const uint32_t t_thr = 0x7ff00000u, trig_hi = 0x7ff00000u;
  double ud = u0, phid = 0.0, prev = 0.0; int parked = 0;
  for (int i = 0; i < updates; ++i) {
    ud += delta;
    const double x = fma(ud * ud, fma(two_m, ud, -1.0), binv2);
    const double y = seed<0>(x);
    const double e = fma(-(x * y), y, 1.0);
    double dphi;
    if (FORM & 1) dphi = y * fma(e, fma(e, k375, 0.5), 1.0);
    else dphi = fma(y * e, fma(e, k375, 0.5), y);
    const double s = prev + dphi;
    prev = dphi;
    if (FORM & 2) phid += s;
    else phid = fma(s, du_h, phid);
    if (hi_word(s) >= t_thr || hi_word(phid) >= trig_hi) ++parked;
  }
  if (phid * du_h == 123.456 || parked == updates + 1) sink[0] = phid + parked;
}
template <typename K> void run_k(const char* name, K kern) {
  cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
  double* sink; cudaMalloc(&sink, 8);
  const int blocks = prop.multiProcessorCount * 5 * 8, threads = 256, updates = 390;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  double best = 0;
  for (int rep = 0; rep < 6; ++rep) {
    cudaEventRecord(e0);
    kern<<<blocks, threads>>>(sink, updates, 20.0, 4.9e-4, 4e-6, 1.1e-5);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double rate = (double)blocks * threads * updates / (ms * 1e-3);
    if (rep > 0 && rate > best) best = rate;
  }
  const double cyc = (double)prop.multiProcessorCount * 4 * 1.965e9 * 32 / best;
  printf("probe %-28s %8.1f Gupdates/s  %.1f SMSP-cycles per warp-update\n", name, best * 1e-9, cyc);
}
template <int SEED> void run(const char* name) {
  cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
  double* sink; cudaMalloc(&sink, 8);
  const int blocks = prop.multiProcessorCount * 5 * 8, threads = 256, updates = 390;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  double best = 0;
  for (int rep = 0; rep < 6; ++rep) {
    cudaEventRecord(e0);
    probe<SEED><<<blocks, threads>>>(sink, updates, 20.0, 4.9e-4, 4e-6, 1.1e-5);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double rate = (double)blocks * threads * updates / (ms * 1e-3);
    if (rep > 0 && rate > best) best = rate;
  }
  const double cyc = (double)prop.multiProcessorCount * 4 * 1.965e9 * 32 / best;
  printf("probe %-28s %8.1f Gupdates/s  %.1f SMSP-cycles per warp-update\n", name, best * 1e-9, cyc);
}
int main() {
  run<0>("MUFU.RSQ64H seed");
  run<1>("F2F + MUFU.RSQ + F2F");
  run<2>("bit tricks + MUFU.RSQ");
  run<3>("no seed (refinement only)");
  run_k("two rays per thread (ILP 2)", probe2);
  run_k("two consecutive updates", probe_pair);
  run_k("form 0 (shipped, as template)", probe_form<0>);
  run_k("form 1: y*(1+e p)", probe_form<1>);
  run_k("form 2: angle as plain sum", probe_form<2>);
  run_k("form 3: both", probe_form<3>);
  return 0;
}
