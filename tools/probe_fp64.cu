// FP64 peak probe for B200: DFMA-only, DMMA-only and both in one kernel.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/probe_fp64 tools/probe_fp64.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// mode 0: DFMA only, 1: DMMA only, 2: even warps DFMA / odd warps DMMA
template <int ILP>
__global__ void __launch_bounds__(256) probe(double* out, int iters, int mode, double x) {
  const int warp = threadIdx.x >> 5;
  double a[ILP], c0[ILP], c1[ILP];
  for (int i = 0; i < ILP; ++i) { a[i] = x + i + threadIdx.x; c0[i] = 0.0; c1[i] = 0.0; }
  const bool use_mma = (mode == 1) || (mode == 2 && (warp & 1));
  if (use_mma) {
    for (int it = 0; it < iters; ++it)
#pragma unroll
      for (int i = 0; i < ILP; ++i) dmma(c0[i], c1[i], a[i], x);
  } else {
    for (int it = 0; it < iters; ++it)
#pragma unroll
      for (int i = 0; i < ILP; ++i) { c0[i] = fma(a[i], x, c0[i]); c1[i] = fma(a[i], x, c1[i]); }
  }
  double s = 0;
  for (int i = 0; i < ILP; ++i) s += c0[i] + c1[i];
  if (s == 12345.678) out[0] = s;
}

int main() {
  double* d; cudaMalloc(&d, 8);
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"l2_bytes\": %d, \"persist_l2_max\": %d, \"smem_optin\": %zu", p.name,
         p.multiProcessorCount, p.l2CacheSize, p.persistingL2CacheMaxSize, p.sharedMemPerBlockOptin);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000, ILP = 8;
  const char* names[3] = {"dfma", "dmma", "mixed"};
  for (int blocks_per_sm = 1; blocks_per_sm <= 4; blocks_per_sm *= 2)
    for (int mode = 0; mode < 3; ++mode) {
      const int grid = p.multiProcessorCount * blocks_per_sm;
      probe<ILP><<<grid, 256>>>(d, 100, mode, 1e-9);
      cudaDeviceSynchronize();
      float best = 1e30f;
      for (int r = 0; r < 3; ++r) {
        cudaEventRecord(e0);
        probe<ILP><<<grid, 256>>>(d, iters, mode, 1e-9);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
      }
      // flops per thread-iteration: dfma 2*ILP fma = 4*ILP flop; dmma: ILP * 256 MAC / 32 lanes * 2 = 16*ILP flop
      double fl_dfma = 4.0 * ILP, fl_dmma = 16.0 * ILP;
      double per_thread = mode == 0 ? fl_dfma : mode == 1 ? fl_dmma : 0.5 * (fl_dfma + fl_dmma);
      double tf = per_thread * iters * 256.0 * grid / (best * 1e-3) / 1e12;
      printf(", \"%s_x%d_tflops\": %.2f", names[mode], blocks_per_sm, tf);
    }
  // sustained: the same kernels back to back for ~1.5 s each (the board's power controller settles within ~0.1 s)
  const char* snames[2] = {"dfma_sustained_tflops", "dmma_sustained_tflops"};
  for (int mode = 0; mode < 2; ++mode) {
    const int grid = p.multiProcessorCount * 4;
    const double per_thread = mode == 0 ? 4.0 * ILP : 16.0 * ILP;
    probe<ILP><<<grid, 256>>>(d, iters, mode, 1e-9);
    cudaDeviceSynchronize();
    int n = 0;
    float ms = 0;
    cudaEventRecord(e0);
    do {
      for (int r = 0; r < 20; ++r) probe<ILP><<<grid, 256>>>(d, iters, mode, 1e-9);
      n += 20;
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      cudaEventElapsedTime(&ms, e0, e1);
    } while (ms < 1500.f);
    printf(", \"%s\": %.2f", snames[mode], per_thread * iters * 256.0 * grid * n / (ms * 1e-3) / 1e12);
  }
  printf("}\n");
  return 0;
}
