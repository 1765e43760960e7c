// Integer-multiply throughput probe (B200): IMAD (lo), IMAD.HI, IMAD.WIDE, LOP3 -- warp-instructions per cycle per SM.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
template <int MODE>
__global__ void __launch_bounds__(256) probe(uint32_t* out, int iters, uint32_t seed) {
  uint32_t a[8];
  for (int i = 0; i < 8; ++i) a[i] = seed + threadIdx.x * 8 + i;
  uint64_t w[8];
  for (int i = 0; i < 8; ++i) w[i] = a[i];
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) a[i] = a[i] * 0xD2511F53u + 1u;                                   // IMAD lo
      if (MODE == 1) a[i] = __umulhi(a[i], 0xD2511F53u) ^ a[i];                       // IMAD.HI + LOP
      if (MODE == 2) { w[i] = (uint64_t)(uint32_t)w[i] * 0xD2511F53u + (w[i] >> 32); } // IMAD.WIDE
      if (MODE == 3) a[i] = (a[i] ^ 0x9E3779B9u) + (a[i] >> 7);                       // ALU
    }
  }
  uint32_t s = 0;
  for (int i = 0; i < 8; ++i) s += a[i] + (uint32_t)w[i] + (uint32_t)(w[i] >> 32);
  if (s == 0x12345678u) out[0] = s;
}
int main() {
  uint32_t* d; cudaMalloc(&d, 4);
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000; int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  const char* names[4] = {"imad_lo", "imad_hi_xor", "imad_wide", "alu_xor_add_shift"};
  printf("{\"clock_khz\": %d", clk);
  for (int mode = 0; mode < 4; ++mode) {
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
      cudaEventRecord(e0);
      if (mode == 0) probe<0><<<p.multiProcessorCount * 4, 256>>>(d, iters, r);
      if (mode == 1) probe<1><<<p.multiProcessorCount * 4, 256>>>(d, iters, r);
      if (mode == 2) probe<2><<<p.multiProcessorCount * 4, 256>>>(d, iters, r);
      if (mode == 3) probe<3><<<p.multiProcessorCount * 4, 256>>>(d, iters, r);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    // loop bodies per second per SM (8 per thread-iteration); instructions per body differ by mode
    double bodies = 8.0 * iters * 256.0 * 4 / (best * 1e-3);   // per SM per second (thread-level)
    printf(", \"%s_thread_bodies_per_clk_per_sm\": %.1f", names[mode], bodies / (clk * 1e3));
  }
  printf("}\n");
  return 0;
}
