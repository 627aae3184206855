// Experiment (not part of libbh8.so): SMSP-cycles per warp FP64 instruction on B200 by operand mix -- how many
// of a DFMA's three sources are registers (the others constant-bank words / immediates), and DMUL / DADD.
// The geodesic update's 11 FP64 instructions run at 2.63 cycles each, the all-constant DFMA chain at 2.18
// (profiles/r02_stepping_chain_variants.txt): is the difference the register operands?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o build/exp/fp64_operands tools/exp_fp64_operands.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int V>
__global__ void __launch_bounds__(256, 5) probe(double* sink, int iters, double ca, double cb) {
  double r = 1.0 + 1e-9 * threadIdx.x, q = 1e-12 * (threadIdx.x + 1), r2 = r + 1e-7, q2 = q * 3.0;
  asm volatile("" : "+d"(r), "+d"(q), "+d"(r2), "+d"(q2));
  double x[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) x[k] = 1.0 + 1e-6 * (threadIdx.x + k);
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int rep = 0; rep < 4; ++rep) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        if (V == 0) x[k] = fma(x[k], ca, cb);                 // 1 register source
        if (V == 1) x[k] = fma(x[k], r, cb);                  // 2 register sources
        if (V == 2) x[k] = fma(x[k], r, q);                   // 3 register sources
        if (V == 3) x[k] = x[k] * r;                          // DMUL, 2 registers
        if (V == 4) x[k] = x[k] + q;                          // DADD, 2 registers
        if (V == 5) x[k] = fma(x[k], (k & 1) ? r : r2, (k & 2) ? q : q2);  // 3 registers, varied
        if (V == 6) x[k] = fma(x[k], x[(k + 1) & 7], q);      // 3 registers, two of them chain values
        if (V == 7) x[k] = fma(x[k], 0.999, 0.5);             // immediates
        if (V == 8) x[k] = x[k] * ca;                         // DMUL, 1 register
      }
    }
  }
  double s = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) s += x[k];
  if (s == 123.456) sink[0] = s;
}
// four chains, every source a different ("fresh") register: no operand comes out of the reuse cache
template <int V>
__global__ void __launch_bounds__(256, 5) fresh(double* sink, int iters, double ca, double cb) {
  double x[4], a[4], b[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    x[k] = 1.0 + 1e-6 * (threadIdx.x + k);
    a[k] = 1.0 - 1e-9 * (threadIdx.x + k);
    b[k] = 1e-12 * (threadIdx.x + k + 1);
    asm volatile("" : "+d"(a[k]), "+d"(b[k]));
  }
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int rep = 0; rep < 8; ++rep) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (V == 0) x[k] = fma(x[k], a[k], b[k]);     // 3 fresh registers
        if (V == 1) x[k] = fma(x[k], a[k], 0.5);      // 2 fresh registers + immediate
        if (V == 2) x[k] = x[k] * a[k];               // DMUL, 2 fresh registers
        if (V == 3) x[k] = x[k] + b[k];               // DADD, 2 fresh registers
        if (V == 4) x[k] = fma(x[k], 0.999, 0.5);     // 1 fresh register
        if (V == 5) x[k] = fma(a[k], b[k], x[k]);     // 3 fresh, the chain value as the addend
      }
    }
  }
  double s = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) s += x[k];
  if (s == 123.456) sink[0] = s;
}
template <int V>
void run(const char* name) {
  double* sink; cudaMalloc(&sink, 8);
  int sms = 0, khz = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  const int iters = 4000, grid = sms * 5;
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  probe<V><<<grid, 256>>>(sink, 100, 0.9999, 1e-4);
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int t = 0; t < 5; ++t) {
    cudaEventRecord(a); probe<V><<<grid, 256>>>(sink, iters, 0.9999, 1e-4); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
  }
  const double warp_instr_per_smsp = (double)iters * 32 * (5.0 * 8 / 4);  // 10 warps per SMSP, 32 instr per iteration
  const double cycles = best * 1e-3 * khz * 1e3;
  printf("%-52s %8.3f ms  %.3f SMSP-cycles per warp instruction (at %d MHz)\n", name, best, cycles / warp_instr_per_smsp, khz / 1000);
  cudaFree(sink);
}
template <int V>
void run_fresh(const char* name) {
  double* sink; cudaMalloc(&sink, 8);
  int sms = 0, khz = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  const int iters = 4000, grid = sms * 5;
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  fresh<V><<<grid, 256>>>(sink, 100, 0.9999, 1e-4);
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int t = 0; t < 5; ++t) {
    cudaEventRecord(a); fresh<V><<<grid, 256>>>(sink, iters, 0.9999, 1e-4); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
  }
  const double warp_instr_per_smsp = (double)iters * 32 * (5.0 * 8 / 4);
  const double cycles = best * 1e-3 * khz * 1e3;
  printf("%-52s %8.3f ms  %.3f SMSP-cycles per warp instruction (at %d MHz)\n", name, best, cycles / warp_instr_per_smsp, khz / 1000);
  cudaFree(sink);
}
int main() {
  run_fresh<0>("fresh: DFMA x = x*a + b   (3 fresh registers)");
  run_fresh<5>("fresh: DFMA x = a*b + x   (3 fresh registers)");
  run_fresh<1>("fresh: DFMA x = x*a + imm (2 fresh registers)");
  run_fresh<2>("fresh: DMUL x = x*a       (2 fresh registers)");
  run_fresh<3>("fresh: DADD x = x+b       (2 fresh registers)");
  run_fresh<4>("fresh: DFMA x = x*imm+imm (1 fresh register)");
  run<0>("DFMA x = x*c + c   (1 register source)");
  run<7>("DFMA x = x*imm + imm");
  run<1>("DFMA x = x*r + c   (2 register sources)");
  run<2>("DFMA x = x*r + q   (3 register sources)");
  run<5>("DFMA x = x*r + q   (3 registers, 2x2 different)");
  run<6>("DFMA x = x*x' + q  (3 registers, two chain values)");
  run<3>("DMUL x = x*r       (2 registers)");
  run<8>("DMUL x = x*c       (1 register)");
  run<4>("DADD x = x+q       (2 registers)");
  return 0;
}
