// Memory-pattern probe for the plane kernel (B200): same grid, same per-CTA traffic as df_plane_panel1_kernel<21,21>
// at 512^2 (one CTA per 64-column strip x plane x component; per step one contiguous 10.75 KB read and 16 rows x
// 64 doubles written), but no arithmetic.  Variants differ only in how the step's 8 KB are written:
//   0: rows of 512 B strided by K*8 (the kernel's real pattern)      1: one contiguous 8 KB block
//   2: real pattern, but strips of 128 columns (1 KB rows)          3: real pattern, full 512-column rows (4 KB)
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/probe_pattern_bw tools/probe_pattern_bw.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(128, 6) pat(const double* __restrict__ src, double* __restrict__ dst, int K, int J,
                                              long comp_stride_in, int panel_cols) {
  constexpr int SJ = 16;
  const int strips = gridDim.x;
  const int TK = K / strips;                          // 64, 128 or 512 columns per CTA
  const long pc = blockIdx.y;
  const int k0 = blockIdx.x * TK;
  const int nsteps = J / SJ;
  const double* in = src + pc * comp_stride_in + (long)k0 * SJ;
  double* out = dst + pc * (long)J * K;
  const int tid = threadIdx.x;
  const int chunk = (TK + 20) * SJ;                   // doubles read per step (with the z halo)
  double acc = 0.0;
  for (int n = 0; n < nsteps; ++n) {
    const double* p = in + (long)n * panel_cols * SJ;
    for (int i = tid; i < chunk; i += 128) acc += __ldg(p + i);
    if (MODE == 1) {
      double* o = out + ((long)n * strips + blockIdx.x) * (TK * SJ);
      for (int i = tid; i < TK * SJ; i += 128) __stcs(o + i, acc);
    } else {
      for (int i = tid; i < TK * SJ; i += 128) {
        const int r = i / TK, c = i % TK;
        __stcs(out + (long)(n * SJ + r) * K + k0 + c, acc);
      }
    }
  }
}

int main() {
  const int J = 512, K = 512, planes = 1024, Qpad = 532, PB = 34;
  const long comp_in = (long)PB * Qpad * 16;
  double *src, *dst;
  cudaMalloc(&src, sizeof(double) * comp_in * 3 * planes);
  cudaMalloc(&dst, sizeof(double) * (long)J * K * 3 * planes);
  cudaMemset(src, 0, sizeof(double) * comp_in * 3 * planes);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const char* names[4] = {"rows_512B", "contiguous_8KB", "rows_1KB", "rows_4KB"};
  printf("{");
  for (int mode = 0; mode < 4; ++mode) {
    const int strips = mode == 2 ? 4 : (mode == 3 ? 1 : 8);
    dim3 grid(strips, 3 * planes);
    float best = 1e30f;
    for (int r = 0; r < 4; ++r) {
      cudaEventRecord(e0);
      if (mode == 0) pat<0><<<grid, 128>>>(src, dst, K, J, comp_in, Qpad);
      if (mode == 1) pat<1><<<grid, 128>>>(src, dst, K, J, comp_in, Qpad);
      if (mode >= 2) pat<0><<<grid, 128>>>(src, dst, K, J, comp_in, Qpad);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (r && ms < best) best = ms;
    }
    const int TK = K / strips;
    const double bytes = 8.0 * 3 * planes * ((double)(J / 16) * strips * ((TK + 20) * 16) + (double)J * K);
    printf("%s\"%s\": {\"ms\": %.3f, \"gbs\": %.1f}", mode ? ", " : "", names[mode], best, bytes / (best * 1e-3) / 1e9);
  }
  printf(", \"err\": \"%s\"}\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
