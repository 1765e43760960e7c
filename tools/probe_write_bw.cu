// HBM write-only / read-only / copy bandwidth probe (streaming stores), B200.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/probe_write_bw tools/probe_write_bw.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void wr(double2* p, size_t n, double v) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x, st = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += st) p[i] = make_double2(v, v + i);
}
__global__ void wr_cs(double2* p, size_t n, double v) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x, st = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += st) __stcs(p + i, make_double2(v, v + i));
}
__global__ void rd(const double2* p, size_t n, double* out) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x, st = (size_t)gridDim.x * blockDim.x;
  double s = 0;
  for (; i < n; i += st) { double2 v = p[i]; s += v.x + v.y; }
  if (s == 1.2345) *out = s;
}
__global__ void cp(const double2* a, double2* b, size_t n) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x, st = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += st) b[i] = a[i];
}
int main() {
  const size_t bytes = 4ull << 30, n = bytes / 16;
  double2 *a, *b; double* o;
  cudaMalloc(&a, bytes); cudaMalloc(&b, bytes); cudaMalloc(&o, 8);
  cudaMemset(a, 0, bytes); cudaMemset(b, 0, bytes);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int grid = 148 * 16, blk = 512;
  const char* names[4] = {"write", "write_cs", "read", "copy"};
  printf("{");
  for (int k = 0; k < 4; ++k) {
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
      cudaEventRecord(e0);
      if (k == 0) wr<<<grid, blk>>>(a, n, 1.0);
      if (k == 1) wr_cs<<<grid, blk>>>(a, n, 1.0);
      if (k == 2) rd<<<grid, blk>>>(a, n, o);
      if (k == 3) cp<<<grid, blk>>>(a, b, n);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (ms < best) best = ms;
    }
    double gb = (k == 3 ? 2.0 : 1.0) * bytes / 1e9;
    printf("%s\"%s_gbs\": %.1f", k ? ", " : "", names[k], gb / (best * 1e-3));
  }
  printf("}\n");
  return 0;
}
