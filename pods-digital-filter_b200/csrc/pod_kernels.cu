// PODFS kernels for sm_100a: weighted Gram X^T W X (FP64 tensor-core MMA), mode projection,
// direct DFT of the temporal modes, row means.
//
// Reference behaviour being replaced (file:line under /root/reference):
//   PODFS.py:1451-1464  calculate_correlation_matrix (np.dot(A.T, A)/ns, or the weighted loop)
//   PODFS.py:1329-1333  spatial modes = A . a[:, :nm] . diag(1/lambda) / ns
//   PODFS.py:1562-1573  O(ns^2) Python DFT of every temporal mode
//   digitalfilters.py:1492-1495 mean over time, subtracted column by column
//
// FP64 tensor path on sm_100a: tcgen05 has no f64 kind, so the FP64 tensor instruction is the
// warp-level mma.sync m8n8k4 (SASS DMMA.8x8x4).  A-fragment (8x4, row-major): lane l holds
// A[l/4][l%4]; B-fragment (4x8, col-major): lane l holds B[l%4][l/4]; C-fragment (8x8): lane l
// holds C[l/4][2*(l%4) + {0,1}].  Shared-memory pitches are = 4 (mod 16) doubles, which makes
// every fragment load of a half-warp hit 16 distinct 8-byte banks.
#include "pod_kernels.cuh"

#include <algorithm>
#include <type_traits>

namespace podsdf {

__device__ __forceinline__ void dmma884(double& c0, double& c1, const double a, const double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}
__device__ __forceinline__ void cp_async_8(void* smem_dst, const void* gsrc, bool valid) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  const int bytes = valid ? 8 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(d), "l"(gsrc), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_16(void* smem_dst, const void* gsrc, bool valid) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  const int bytes = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(d), "l"(gsrc), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

// ------------------------------------------------------------------------------------------------
// Gram: W[split][tile][128][128] partial = sum over this split's rows of  A_i^T diag(w) A_j
//   tile = (ib <= jb) pair of 128-snapshot blocks (upper triangle only); 8 warps, warp tile 64 x 32.
//   LAYOUT 0: A[row][t] (reference, time contiguous): smem tiles [KB rows][128 + 4]
//   LAYOUT 1: a[t][row] (DF-native, space contiguous): smem tiles [128 t][KB + 4]
// ------------------------------------------------------------------------------------------------
constexpr int GT = 128;   // snapshots per tile side
#ifndef PODSDF_GKB
#define PODSDF_GKB 32             // 16 -> 32: half the CTA barriers per MMA, 2 % faster at the bench size (A/B in r02 notes)
#endif
constexpr int GKB = PODSDF_GKB;   // rows (reduction depth) per pipeline stage
constexpr int GST = 3;    // cp.async stages
#ifndef PODSDF_GRAM_WARPS
#define PODSDF_GRAM_WARPS 8
#endif
// warps per CTA: 8 (warp tile 64 x 32) or 16 (32 x 32, -DPODSDF_GRAM_WARPS=16).  Measured at 393 216 x 10^4: 1.244 s with
// 8, 1.266 s with 16 -- two warps per scheduler already keep the tensor pipe at 87 %, four pay more shared-memory loads
// per DMMA (0.5 instead of 0.375) for nothing.  (The rank-2b update of the eigen-solver, whose CTAs spend half their
// time outside the MMA loop, is the opposite case: there a lone warp per scheduler reached 45 %.)
constexpr int GWARPS = PODSDF_GRAM_WARPS;
constexpr int GTHR = 32 * GWARPS;
constexpr int GMT = 128 / (GWARPS / 4) / 8;   // 8-row MMA tiles per warp
static_assert(GWARPS == 8 || GWARPS == 16, "warp layout");
constexpr int GP0 = GT + 4;   // layout-0 pitch (132 = 4 mod 16)
constexpr int GP1 = GKB + 4;  // layout-1 pitch (20 = 4 mod 16)

template <int LAYOUT>
struct GramSmem {
  static constexpr int TILE = (LAYOUT == 0) ? GKB * GP0 : GT * GP1;  // doubles per operand per stage
  static constexpr size_t BYTES = (size_t)GST * (2 * TILE + GKB) * sizeof(double);
};

template <int LAYOUT, bool VEC, bool WEIGHT>
__global__ void __launch_bounds__(GTHR, 1)
gram_dmma_kernel(const double* __restrict__ A, const long M, const long N, const double* __restrict__ w,
                 const long P, double* __restrict__ W, const int ntile_side, const long rows_per_split) {
  extern __shared__ double smem[];
  constexpr int TILE = GramSmem<LAYOUT>::TILE;
  double* sA = smem;                      // [GST][TILE]  i-block operand
  double* sB = smem + GST * TILE;         // [GST][TILE]  j-block operand
  double* sW = smem + 2 * GST * TILE;     // [GST][GKB]   row weights

  // linear tile id -> (ib, jb) with ib <= jb
  int ib = 0, rem = blockIdx.x;
  while (rem >= ntile_side - ib) { rem -= ntile_side - ib; ++ib; }
  const int jb = ib + rem;
  const long i0 = (long)ib * GT, j0 = (long)jb * GT;
  const long r_begin = (long)blockIdx.y * rows_per_split;
  const long r_end = min(M, r_begin + rows_per_split);
  const int nchunks = (int)((r_end - r_begin + GKB - 1) / GKB);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = (warp >> 2) * (GMT * 8), wn = (warp & 3) * 32;  // warp tile origin inside the CTA tile
  const int lr = lane & 3, lc = lane >> 2;

  auto load_stage = [&](int chunk, int stage) {
    const long r0 = r_begin + (long)chunk * GKB;
    double* a = sA + stage * TILE;
    double* b = sB + stage * TILE;
    if (LAYOUT == 0) {
      if (VEC) {  // 16-byte copies: GKB rows x 64 pairs per operand
        for (int e = tid; e < GKB * (GT / 2); e += GTHR) {
          const int r = e / (GT / 2), x = (e % (GT / 2)) * 2;
          const bool rok = r0 + r < r_end;
          cp_async_16(a + r * GP0 + x, A + (r0 + r) * N + i0 + x, rok && (i0 + x < N));
          cp_async_16(b + r * GP0 + x, A + (r0 + r) * N + j0 + x, rok && (j0 + x < N));
        }
      } else {
        for (int e = tid; e < GKB * GT; e += GTHR) {
          const int r = e / GT, x = e % GT;
          const bool rok = r0 + r < r_end;
          cp_async_8(a + r * GP0 + x, A + (r0 + r) * N + i0 + x, rok && (i0 + x < N));
          cp_async_8(b + r * GP0 + x, A + (r0 + r) * N + j0 + x, rok && (j0 + x < N));
        }
      }
    } else {
      if (VEC) {
        for (int e = tid; e < GT * (GKB / 2); e += GTHR) {
          const int t = e / (GKB / 2), r = (e % (GKB / 2)) * 2;
          const bool rok = r0 + r < r_end;  // M even and r0 even under VEC, so pairs never straddle
          cp_async_16(a + t * GP1 + r, A + (i0 + t) * M + r0 + r, rok && (i0 + t < N));
          cp_async_16(b + t * GP1 + r, A + (j0 + t) * M + r0 + r, rok && (j0 + t < N));
        }
      } else {
        for (int e = tid; e < GT * GKB; e += GTHR) {
          const int t = e / GKB, r = e % GKB;
          const bool rok = r0 + r < r_end;
          cp_async_8(a + t * GP1 + r, A + (i0 + t) * M + r0 + r, rok && (i0 + t < N));
          cp_async_8(b + t * GP1 + r, A + (j0 + t) * M + r0 + r, rok && (j0 + t < N));
        }
      }
    }
    if (WEIGHT && tid < GKB) {
      const long r = r0 + tid;
      cp_async_8(sW + stage * GKB + tid, w + (r < r_end ? r % P : 0), r < r_end);
    }
  };

  double acc[GMT][4][2];
#pragma unroll
  for (int a = 0; a < GMT; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;

  for (int s = 0; s < GST - 1; ++s) {
    if (s < nchunks) load_stage(s, s);
    cp_async_commit();
  }
  for (int c = 0; c < nchunks; ++c) {
    cp_async_wait<GST - 2>();
    __syncthreads();
    if (c + GST - 1 < nchunks) load_stage(c + GST - 1, (c + GST - 1) % GST);
    cp_async_commit();
    const double* a = sA + (c % GST) * TILE;
    const double* b = sB + (c % GST) * TILE;
    const double* ws = sW + (c % GST) * GKB;
#pragma unroll
    for (int k4 = 0; k4 < GKB; k4 += 4) {
      double af[GMT], bf[4];
      const int r = k4 + lr;
#pragma unroll
      for (int mt = 0; mt < GMT; ++mt)
        af[mt] = (LAYOUT == 0) ? a[r * GP0 + wm + mt * 8 + lc] : a[(wm + mt * 8 + lc) * GP1 + r];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        bf[nt] = (LAYOUT == 0) ? b[r * GP0 + wn + nt * 8 + lc] : b[(wn + nt * 8 + lc) * GP1 + r];
        if (WEIGHT) bf[nt] *= ws[r];
      }
#pragma unroll
      for (int mt = 0; mt < GMT; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) dmma884(acc[mt][nt][0], acc[mt][nt][1], af[mt], bf[nt]);
    }
  }
  cp_async_wait<0>();

  // partial tile -> workspace [split][tile][128][128]
  double* Wt = W + ((long)blockIdx.y * gridDim.x + blockIdx.x) * (GT * GT);
#pragma unroll
  for (int mt = 0; mt < GMT; ++mt)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      const int row = wm + mt * 8 + lc, col = wn + nt * 8 + 2 * lr;
      *reinterpret_cast<double2*>(Wt + row * GT + col) = make_double2(acc[mt][nt][0], acc[mt][nt][1]);
    }
}

// C[i][j] = C[j][i] = scale * sum_split W[split][tile][i-i0][j-j0]   (deterministic reduction order)
__global__ void __launch_bounds__(256)
gram_reduce_kernel(const double* __restrict__ W, const int nsplit, const int ntiles, const int ntile_side,
                   const long N, const double scale, double* __restrict__ C) {
  int ib = 0, rem = blockIdx.x;
  while (rem >= ntile_side - ib) { rem -= ntile_side - ib; ++ib; }
  const int jb = ib + rem;
  for (int e = threadIdx.x; e < GT * GT; e += 256) {
    const int r = e / GT, c = e % GT;
    const long i = (long)ib * GT + r, j = (long)jb * GT + c;
    if (i >= N || j >= N) continue;
    if (ib == jb && c < r) continue;  // diagonal tile: take the upper half, mirror below
    double s = 0.0;
    for (int sp = 0; sp < nsplit; ++sp) s += W[((long)sp * ntiles + blockIdx.x) * (GT * GT) + e];
    s *= scale;
    C[i * N + j] = s;
    C[j * N + i] = s;
  }
}

// How many row ranges the reduction is cut into (one CTA per tile and range, one CTA resident per SM):
// few tiles -> split until ~2 waves exist; many tiles -> the smallest split (<= 8) that fills >= 99.5 % of its waves
// (N = 10^4: 3160 tiles = 21.35 waves of 148 -> 2.9 % of the machine idles in the tail; split 5 = 106.76 waves -> 0.2 %;
// measured at M = 393 216: 1300 ms -> 1266 ms).
// PODSDF_GRAM_SPLIT overrides (profiling).
static int gram_pick_split(long ntiles, long M) {
  if (const char* v = getenv("PODSDF_GRAM_SPLIT")) {
    const int s = atoi(v);
    if (s >= 1 && M / s >= 4 * GKB) return s;
  }
  int nsplit = 1;
  while (ntiles * nsplit < 2L * 148 && M / (nsplit * 2) >= 4 * GKB) nsplit *= 2;
  if (ntiles >= 2L * 148) {
    double best_eff = 0.0;
    for (int s = 1; s <= 8 && M / s >= 64 * GKB; ++s) {
      const double waves = (double)ntiles * s / 148.0;
      const double eff = waves / ceil(waves);
      if (eff > best_eff + 1e-9) {
        best_eff = eff;
        nsplit = s;
      }
      if (eff >= 0.995) break;
    }
  }
  return nsplit;
}

int launch_gram(const double* A, long M, long N, int layout, const double* w, long P, double scale, double* C,
                double* W, size_t W_bytes, cudaStream_t st) {
  const int side = ceil_div(N, GT);
  const int ntiles = side * (side + 1) / 2;
  int nsplit = gram_pick_split(ntiles, M);
  long rows_per_split = ((M + nsplit - 1) / nsplit + GKB - 1) / GKB * GKB;
  nsplit = ceil_div(M, rows_per_split);
  const size_t need = (size_t)nsplit * ntiles * GT * GT * sizeof(double);
  if (need > W_bytes) return fail(PODSDF_ERR_INVALID, "gram workspace too small");
  const bool aligned = (((uintptr_t)A) & 15) == 0;
  const bool vec = aligned && ((layout == 0) ? (N % 2 == 0) : (M % 2 == 0));
  dim3 grid(ntiles, nsplit);
#define LAUNCH(L, V, WT)                                                                                     \
  do {                                                                                                       \
    auto k = gram_dmma_kernel<L, V, WT>;                                                                     \
    PODSDF_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GramSmem<L>::BYTES)); \
    k<<<grid, GTHR, GramSmem<L>::BYTES, st>>>(A, M, N, w, P, W, side, rows_per_split);                       \
  } while (0)
  if (layout == 0) {
    if (vec) { if (w) LAUNCH(0, true, true); else LAUNCH(0, true, false); }
    else     { if (w) LAUNCH(0, false, true); else LAUNCH(0, false, false); }
  } else {
    if (vec) { if (w) LAUNCH(1, true, true); else LAUNCH(1, true, false); }
    else     { if (w) LAUNCH(1, false, true); else LAUNCH(1, false, false); }
  }
#undef LAUNCH
  count_launch();
  gram_reduce_kernel<<<ntiles, 256, 0, st>>>(W, nsplit, ntiles, side, N, scale, C);
  count_launch();
  PODSDF_CUDA(cudaGetLastError());
  return 0;
}

size_t gram_workspace_bytes(long M, long N) {
  const int side = ceil_div(N, GT);
  const long ntiles = (long)side * (side + 1) / 2;
  int nsplit = gram_pick_split(ntiles, M);
  long rows_per_split = ((M + nsplit - 1) / nsplit + GKB - 1) / GKB * GKB;
  nsplit = ceil_div(M, rows_per_split);
  return (size_t)nsplit * ntiles * GT * GT * sizeof(double);
}

// ------------------------------------------------------------------------------------------------
// Projection: Phi[row][m] = scale * dinv[m] * sum_t A[row][t] T[t][m]     (PODFS.py:1329-1333)
//   CTA = 128 rows x 64 or 128 modes (mode blocks are neighbours in the grid); 8 warps, warp tile 16 rows x all modes.
// ------------------------------------------------------------------------------------------------
#ifndef PODSDF_PKB
#define PODSDF_PKB 32
#endif
constexpr int PR = 128, PKB = PODSDF_PKB, PST = 3;   // K chunk of 32: half as many CTA barriers as 16 (same trade as the Gram kernel's)
constexpr int PPA0 = PKB + 4;  // layout 0: A tile [128 rows][PKB + 4]
constexpr int PPA1 = PR + 4;   // layout 1: A tile [PKB t][128 + 4]

// PM = modes per CTA: 64, or 128 when more than 64 modes are asked for (config 4 keeps 100: one pass over A with
// 13 of 16 column tiles live instead of two passes with 8 + 5 of 8)
template <int LAYOUT, int PM>
struct ProjSmem {
  static constexpr int PPT = PM + 4;    // T tile [PKB t][PM + 4]
  static constexpr int ATILE = (LAYOUT == 0) ? PR * PPA0 : PKB * PPA1;
  static constexpr int TTILE = PKB * PPT;
  static constexpr size_t BYTES = (size_t)PST * (ATILE + TTILE) * sizeof(double);
};

template <int LAYOUT, int PM>
__global__ void __launch_bounds__(256, 1)
project_dmma_kernel(const double* __restrict__ A, const long M, const long N, const double* __restrict__ T,
                    const long ldT, const double* __restrict__ dinv, const int nm, const double scale,
                    double* __restrict__ Phi, const int vec) {
  extern __shared__ double smem[];
  constexpr int ATILE = ProjSmem<LAYOUT, PM>::ATILE, TTILE = ProjSmem<LAYOUT, PM>::TTILE, PPT = ProjSmem<LAYOUT, PM>::PPT;
  constexpr int NTL = PM / 8;
  double* sA = smem;
  double* sT = smem + PST * ATILE;
  // mode blocks of the same rows are neighbours in the grid: the second one finds its A tiles in L2
  const int nmb = (nm + PM - 1) / PM;
  const long row0 = (long)(blockIdx.x / nmb) * PR;
  const int m0 = (int)(blockIdx.x % nmb) * PM;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int lr = lane & 3, lc = lane >> 2;
  const int wr = warp * 16;
  const int nchunks = (int)((N + PKB - 1) / PKB);
  const int ntl = min(NTL, (nm - m0 + 7) / 8);    // 8-mode column tiles that hold real modes (CTA-uniform)

  auto load_stage = [&](int chunk, int stage) {
    const long t0 = (long)chunk * PKB;
    double* a = sA + stage * ATILE;
    double* tt = sT + stage * TTILE;
    if (vec) {          // every row of A and T starts 16-byte aligned and has an even length: two entries per copy
      if (LAYOUT == 0) {
        for (int e = tid; e < PR * PKB / 2; e += 256) {
          const int r = e / (PKB / 2), t = 2 * (e % (PKB / 2));
          cp_async_16(a + r * PPA0 + t, A + (row0 + r) * N + t0 + t, (row0 + r < M) && (t0 + t < N));
        }
      } else {
        for (int e = tid; e < PKB * PR / 2; e += 256) {
          const int t = e / (PR / 2), r = 2 * (e % (PR / 2));
          cp_async_16(a + t * PPA1 + r, A + (t0 + t) * M + row0 + r, (row0 + r < M) && (t0 + t < N));
        }
      }
      for (int e = tid; e < PKB * PM / 2; e += 256) {
        const int t = e / (PM / 2), m = 2 * (e % (PM / 2));
        cp_async_16(tt + t * PPT + m, T + (t0 + t) * ldT + m0 + m, (t0 + t < N) && (m0 + m < nm));
      }
      return;
    }
    if (LAYOUT == 0) {
      for (int e = tid; e < PR * PKB; e += 256) {
        const int r = e / PKB, t = e % PKB;
        cp_async_8(a + r * PPA0 + t, A + (row0 + r) * N + t0 + t, (row0 + r < M) && (t0 + t < N));
      }
    } else {
      for (int e = tid; e < PKB * PR; e += 256) {
        const int t = e / PR, r = e % PR;
        cp_async_8(a + t * PPA1 + r, A + (t0 + t) * M + row0 + r, (row0 + r < M) && (t0 + t < N));
      }
    }
    for (int e = tid; e < PKB * PM; e += 256) {
      const int t = e / PM, m = e % PM;
      cp_async_8(tt + t * PPT + m, T + (t0 + t) * ldT + m0 + m, (t0 + t < N) && (m0 + m < nm));
    }
  };

  // The number of live 8-mode column tiles is CTA-uniform but not a compile-time constant, and predicating the MMAs on it
  // costs the tensor pipe a slot for every predicated-off DMMA (nm = 100: 16 issued for 13 executed per row tile -- the
  // pipe was 85 % busy at 66 % of its useful peak, ncu).  So the whole loop is instantiated per count.
  auto run = [&](auto tag) {
    constexpr int NV = decltype(tag)::value;
    double acc[2][NV][2];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int b = 0; b < NV; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;

    for (int s = 0; s < PST - 1; ++s) {
      if (s < nchunks) load_stage(s, s);
      cp_async_commit();
    }
    for (int c = 0; c < nchunks; ++c) {
      cp_async_wait<PST - 2>();
      __syncthreads();
      if (c + PST - 1 < nchunks) load_stage(c + PST - 1, (c + PST - 1) % PST);
      cp_async_commit();
      const double* a = sA + (c % PST) * ATILE;
      const double* tt = sT + (c % PST) * TTILE;
#pragma unroll
      for (int k4 = 0; k4 < PKB; k4 += 4) {
        double af[2], bf[NV];
        const int t = k4 + lr;
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
          af[mt] = (LAYOUT == 0) ? a[(wr + mt * 8 + lc) * PPA0 + t] : a[t * PPA1 + wr + mt * 8 + lc];
#pragma unroll
        for (int nt = 0; nt < NV; ++nt) bf[nt] = tt[t * PPT + nt * 8 + lc];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
          for (int nt = 0; nt < NV; ++nt) dmma884(acc[mt][nt][0], acc[mt][nt][1], af[mt], bf[nt]);
      }
    }
    cp_async_wait<0>();
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < NV; ++nt) {
        const long row = row0 + wr + mt * 8 + lc;
        const int m = m0 + nt * 8 + 2 * lr;
        if (row < M) {
          if (m < nm) Phi[row * nm + m] = acc[mt][nt][0] * dinv[m] * scale;
          if (m + 1 < nm) Phi[row * nm + m + 1] = acc[mt][nt][1] * dinv[m + 1] * scale;
        }
      }
  };
#define PODSDF_PROJ_CASE(n) \
  case n:                   \
    if constexpr (n <= NTL) run(std::integral_constant<int, n>{}); \
    break;
  switch (ntl) {
    PODSDF_PROJ_CASE(1) PODSDF_PROJ_CASE(2) PODSDF_PROJ_CASE(3) PODSDF_PROJ_CASE(4)
    PODSDF_PROJ_CASE(5) PODSDF_PROJ_CASE(6) PODSDF_PROJ_CASE(7) PODSDF_PROJ_CASE(8)
    PODSDF_PROJ_CASE(9) PODSDF_PROJ_CASE(10) PODSDF_PROJ_CASE(11) PODSDF_PROJ_CASE(12)
    PODSDF_PROJ_CASE(13) PODSDF_PROJ_CASE(14) PODSDF_PROJ_CASE(15) PODSDF_PROJ_CASE(16)
    default: break;
  }
#undef PODSDF_PROJ_CASE
}

template <int LAYOUT, int PM>
static int launch_project_t(const double* A, long M, long N, const double* T, long ldT, const double* dinv, int nm,
                            double scale, double* Phi, cudaStream_t st) {
  const long nblocks = ceil_div(M, (long)PR) * ceil_div(nm, PM);
  if (nblocks > 2147483647L) return fail(PODSDF_ERR_UNSUPPORTED, "too many tiles for one projection launch");
  auto k = project_dmma_kernel<LAYOUT, PM>;
  PODSDF_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ProjSmem<LAYOUT, PM>::BYTES));
  // 16-byte copies: aligned bases, even row lengths (layout 0: rows of length N; layout 1: rows of length M), even nm
  const bool aligned = (((uintptr_t)A | (uintptr_t)T) & 15) == 0;
  const int vec = aligned && ((LAYOUT == 0 ? N : M) % 2 == 0) && (ldT % 2 == 0) && (nm % 2 == 0) ? 1 : 0;
  k<<<dim3((unsigned)nblocks), 256, ProjSmem<LAYOUT, PM>::BYTES, st>>>(A, M, N, T, ldT, dinv, nm, scale, Phi, vec);
  count_launch();
  PODSDF_CUDA(cudaGetLastError());
  return 0;
}

int launch_project(const double* A, long M, long N, int layout, const double* T, long ldT, const double* dinv,
                   int nm, double scale, double* Phi, cudaStream_t st) {
  const bool wide = nm > 64;
  if (layout == 0)
    return wide ? launch_project_t<0, 128>(A, M, N, T, ldT, dinv, nm, scale, Phi, st)
                : launch_project_t<0, 64>(A, M, N, T, ldT, dinv, nm, scale, Phi, st);
  return wide ? launch_project_t<1, 128>(A, M, N, T, ldT, dinv, nm, scale, Phi, st)
              : launch_project_t<1, 64>(A, M, N, T, ldT, dinv, nm, scale, Phi, st);
}

// ------------------------------------------------------------------------------------------------
// DFT of the temporal modes (PODFS.py:1562-1573): c[n][i] = (1/N) sum_j a[j][i] e^{-2 pi i k j / N},
// k = n - N/2.  Twiddles come from an exact table tw[r] = e^{-2 pi i r / N} indexed with (k j) mod N,
// so the phase never loses accuracy to large arguments.  Composite N runs the two-factor transform below;
// the direct kernel (thread = (mode i, 4 frequencies)) remains for prime N.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) dft_twiddle_kernel(double2* __restrict__ tw, const long N) {
  const long r = blockIdx.x * 256L + threadIdx.x;
  if (r >= N) return;
  double s, c;
  sincospi(-2.0 * (double)r / (double)N, &s, &c);
  tw[r] = make_double2(c, s);
}

constexpr int DF_NF = 4;  // frequencies per thread
__global__ void __launch_bounds__(128)
dft_modes_kernel(const double* __restrict__ a, const long N, const long ld, const int nm,
                 const double2* __restrict__ tw, double2* __restrict__ c) {
  const int i = blockIdx.x * 32 + (threadIdx.x & 31);          // mode
  const long nb = ((long)blockIdx.y * 4 + (threadIdx.x >> 5)) * DF_NF;  // first frequency index
  if (nb >= N) return;
  const long half = N / 2;
  long step[DF_NF], idx[DF_NF];
  double re[DF_NF], im[DF_NF];
#pragma unroll
  for (int f = 0; f < DF_NF; ++f) {
    const long k = (nb + f) - half;
    step[f] = ((k % N) + N) % N;
    idx[f] = 0;
    re[f] = im[f] = 0.0;
  }
  const bool ok = i < nm;
  for (long j = 0; j < N; ++j) {
    const double v = ok ? a[j * ld + i] : 0.0;
#pragma unroll
    for (int f = 0; f < DF_NF; ++f) {
      const double2 t = tw[idx[f]];
      re[f] = fma(v, t.x, re[f]);
      im[f] = fma(v, t.y, im[f]);
      idx[f] += step[f];
      if (idx[f] >= N) idx[f] -= N;
    }
  }
  if (!ok) return;
  const double inv = 1.0 / (double)N;
#pragma unroll
  for (int f = 0; f < DF_NF; ++f)
    if (nb + f < N) c[(nb + f) * nm + i] = make_double2(re[f] * inv, im[f] * inv);
}

// Two-factor (Cooley-Tukey, one split) DFT for N = N1 N2: with j = N2 j1 + j2 and k' = k1 + N1 k2,
//   X[k1 + N1 k2] = sum_{j2} W^{N1 k2 j2} [ W^{k1 j2} sum_{j1} a[N2 j1 + j2] W^{N2 k1 j1} ],   W = exp(-2 pi i / N),
// i.e. N (N1 + N2) complex multiply-adds per mode instead of N^2 (10^4: 2 10^6 instead of 10^8), every twiddle taken
// from the same exact table tw[r], r reduced mod N in integers.  N prime degenerates to the direct sum (N1 = 1).
// Step 1: Y[k1][j2][i] (with the middle twiddle applied); thread = (mode i, j2, k1).
__global__ void __launch_bounds__(128)
dft_step1_kernel(const double* __restrict__ a, const long N, const long N1, const long N2, const long ld, const int nm,
                 const double2* __restrict__ tw, double2* __restrict__ Y) {
  const int i = blockIdx.x * 32 + (threadIdx.x & 31);
  const long e = (long)blockIdx.y * 4 + (threadIdx.x >> 5);      // k1 * N2 + j2
  if (e >= N || i >= nm) return;
  const long k1 = e / N2, j2 = e % N2;
  const long step = (N2 * k1) % N;                               // exponent step per j1
  long idx = 0;
  double re = 0.0, im = 0.0;
  for (long j1 = 0; j1 < N1; ++j1) {
    const double v = a[(N2 * j1 + j2) * ld + i];
    const double2 t = tw[idx];
    re = fma(v, t.x, re);
    im = fma(v, t.y, im);
    idx += step;
    if (idx >= N) idx -= N;
  }
  const double2 m = tw[(k1 * j2) % N];                            // W^{k1 j2}
  Y[e * nm + i] = make_double2(re * m.x - im * m.y, re * m.y + im * m.x);
}

// Step 2: c[n][i] = (1/N) sum_{j2} Y[k1][j2][i] W^{N1 k2 j2}, n = (k1 + N1 k2 + N/2) mod N; thread = (mode i, k').
__global__ void __launch_bounds__(128)
dft_step2_kernel(const double2* __restrict__ Y, const long N, const long N1, const long N2, const int nm,
                 const double2* __restrict__ tw, double2* __restrict__ c) {
  const int i = blockIdx.x * 32 + (threadIdx.x & 31);
  const long kp = (long)blockIdx.y * 4 + (threadIdx.x >> 5);     // k' = k1 + N1 k2
  if (kp >= N || i >= nm) return;
  const long k1 = kp % N1, k2 = kp / N1;
  const long step = (N1 * k2) % N;
  long idx = 0;
  double re = 0.0, im = 0.0;
  const double2* y = Y + (k1 * N2) * nm + i;
  for (long j2 = 0; j2 < N2; ++j2) {
    const double2 v = y[j2 * nm];
    const double2 t = tw[idx];
    re = fma(v.x, t.x, fma(-v.y, t.y, re));
    im = fma(v.x, t.y, fma(v.y, t.x, im));
    idx += step;
    if (idx >= N) idx -= N;
  }
  const double inv = 1.0 / (double)N;
  const long n = (kp + N / 2) % N;
  c[n * nm + i] = make_double2(re * inv, im * inv);
}

size_t dft_workspace_doubles(long N, int nm) { return 2 * (size_t)N + 2 * (size_t)N * nm; }

int launch_dft(const double* a, long N, long ld, int nm, double* tw_ws, double* c, cudaStream_t st) {
  dft_twiddle_kernel<<<ceil_div(N, 256), 256, 0, st>>>(reinterpret_cast<double2*>(tw_ws), N);
  count_launch();
  long N1 = 1;                                                   // largest divisor of N not above sqrt(N)
  for (long f = 1; f * f <= N; ++f)
    if (N % f == 0) N1 = f;
  const long N2 = N / N1;
  const double2* tw = reinterpret_cast<const double2*>(tw_ws);
  if (N1 == 1) {                                                 // prime length: the direct sum
    dim3 grid(ceil_div(nm, 32), ceil_div(N, 4 * DF_NF));
    dft_modes_kernel<<<grid, 128, 0, st>>>(a, N, ld, nm, tw, reinterpret_cast<double2*>(c));
    count_launch();
  } else {
    double2* Y = reinterpret_cast<double2*>(tw_ws + 2 * N);
    dim3 grid(ceil_div(nm, 32), ceil_div(N, 4L));
    dft_step1_kernel<<<grid, 128, 0, st>>>(a, N, N1, N2, ld, nm, tw, Y);
    dft_step2_kernel<<<grid, 128, 0, st>>>(Y, N, N1, N2, nm, tw, reinterpret_cast<double2*>(c));
    count_launch(2);
  }
  PODSDF_CUDA(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------------
// Row means over time and in-place subtraction (digitalfilters.py:1492-1495)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
mean_rows_l0_kernel(double* __restrict__ A, const long M, const long N, double* __restrict__ mean, const int sub) {
  const long row = blockIdx.x * 8L + (threadIdx.x >> 5);  // warp per row, time contiguous
  if (row >= M) return;
  const int lane = threadIdx.x & 31;
  double* p = A + row * N;
  double s = 0.0;
  for (long t = lane; t < N; t += 32) s += p[t];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const double m = s / (double)N;
  if (lane == 0) mean[row] = m;
  if (sub)
    for (long t = lane; t < N; t += 32) p[t] -= m;
}

__global__ void __launch_bounds__(256)
mean_rows_l1_kernel(double* __restrict__ A, const long M, const long N, double* __restrict__ mean, const int sub) {
  const long row = blockIdx.x * 256L + threadIdx.x;  // thread per row, rows contiguous
  if (row >= M) return;
  double s = 0.0;
  for (long t = 0; t < N; ++t) s += A[t * M + row];
  const double m = s / (double)N;
  mean[row] = m;
  if (sub)
    for (long t = 0; t < N; ++t) A[t * M + row] -= m;
}

// Layout 0, rows up to MR_MAXPT * 256 entries long: one CTA per row keeps the row in registers between the sum and the
// subtraction, so the matrix is read once and written once (the warp-per-row kernel above reads it twice: 94 GB instead
// of 63 GB at 393 216 x 10^4, 14.2 ms -> see DESIGN.md section 6).
constexpr int MR_MAXPT = 40;
__global__ void __launch_bounds__(256)
mean_rows_l0_reg_kernel(double* __restrict__ A, const long M, const long N, double* __restrict__ mean, const int sub) {
  __shared__ double s_part[8];
  __shared__ double s_mean;
  const long row = blockIdx.x;
  double* p = A + row * N;
  const int tid = threadIdx.x;
  double v[MR_MAXPT];
  double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll
  for (int q = 0; q < MR_MAXPT; ++q) {
    const long t = tid + 256L * q;
    v[q] = (t < N) ? __ldcs(p + t) : 0.0;           // streamed: nothing of this row is read again
  }
#pragma unroll
  for (int q = 0; q < MR_MAXPT; q += 4) {
    s0 += v[q];
    s1 += v[q + 1];
    s2 += v[q + 2];
    s3 += v[q + 3];
  }
  double s = (s0 + s1) + (s2 + s3);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((tid & 31) == 0) s_part[tid >> 5] = s;
  __syncthreads();
  if (tid == 0) {
    double tot = 0.0;
#pragma unroll
    for (int w = 0; w < 8; ++w) tot += s_part[w];
    const double m = tot / (double)N;
    s_mean = m;
    mean[row] = m;
  }
  if (!sub) return;
  __syncthreads();
  const double m = s_mean;
#pragma unroll
  for (int q = 0; q < MR_MAXPT; ++q) {
    const long t = tid + 256L * q;
    if (t < N) p[t] = v[q] - m;
  }
}
static_assert(MR_MAXPT % 4 == 0, "four partial sums");

int launch_mean_rows(double* A, long M, long N, int layout, double* mean, int subtract, cudaStream_t st) {
  if (layout == 0 && N <= 256L * MR_MAXPT && N >= 2048 && M <= 2147483647L)
    mean_rows_l0_reg_kernel<<<(unsigned)M, 256, 0, st>>>(A, M, N, mean, subtract);
  else if (layout == 0)
    mean_rows_l0_kernel<<<ceil_div(M, 8), 256, 0, st>>>(A, M, N, mean, subtract);
  else
    mean_rows_l1_kernel<<<ceil_div(M, 256), 256, 0, st>>>(A, M, N, mean, subtract);
  count_launch();
  PODSDF_CUDA(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------------
// One- and two-point statistics of a three-component field given as a snapshot matrix (a-posteriori
// checks of the PODFS reconstruction, docs/usage/CFDCodeIntegration.rst:85-91; the same moments the Lund transform
// prescribes, digitalfilters.py:143-178).
//   moments: mean[c][p] and the six second central moments uu, vv, ww, uv, uw, vw per point, accumulated about
//            the first sample (no cancellation between U ~ 1 and u' ~ 0.02).
//   lag covariance: <a'(t, j, k) a'(t + dt, j + dj, k + dk)> over all valid (t, j, k) of one component.
// ------------------------------------------------------------------------------------------------
template <int LAYOUT>
__global__ void __launch_bounds__(128)
field_moments_kernel(const double* __restrict__ F, const long P, const long N, double* __restrict__ mean,
                     double* __restrict__ stress) {
  // LAYOUT 1 (f[t][c*P + p]): thread = point, consecutive threads read consecutive doubles
  // LAYOUT 0 (F[c*P + p][t]): warp = point, lanes stride over time
  const int lane = threadIdx.x & 31;
  const long p = (LAYOUT == 1) ? blockIdx.x * 128L + threadIdx.x : blockIdx.x * 4L + (threadIdx.x >> 5);
  if (p >= P) return;
  auto at = [&](int c, long t) { return (LAYOUT == 1) ? F[t * 3 * P + c * P + p] : F[(c * P + p) * N + t]; };
  const double k0 = at(0, 0), k1 = at(1, 0), k2 = at(2, 0);
  double s[3] = {0, 0, 0}, q[6] = {0, 0, 0, 0, 0, 0};
  for (long t = (LAYOUT == 1) ? 0 : lane; t < N; t += (LAYOUT == 1) ? 1 : 32) {
    const double a = at(0, t) - k0, b = at(1, t) - k1, c = at(2, t) - k2;
    s[0] += a; s[1] += b; s[2] += c;
    q[0] = fma(a, a, q[0]); q[1] = fma(b, b, q[1]); q[2] = fma(c, c, q[2]);
    q[3] = fma(a, b, q[3]); q[4] = fma(a, c, q[4]); q[5] = fma(b, c, q[5]);
  }
  if (LAYOUT == 0) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
      for (int i = 0; i < 3; ++i) s[i] += __shfl_xor_sync(0xffffffffu, s[i], o);
#pragma unroll
      for (int i = 0; i < 6; ++i) q[i] += __shfl_xor_sync(0xffffffffu, q[i], o);
    }
    if (lane != 0) return;
  }
  const double inv = 1.0 / (double)N;
  const double m0 = s[0] * inv, m1 = s[1] * inv, m2 = s[2] * inv;
  mean[p] = k0 + m0; mean[P + p] = k1 + m1; mean[2 * P + p] = k2 + m2;
  stress[p] = q[0] * inv - m0 * m0;
  stress[P + p] = q[1] * inv - m1 * m1;
  stress[2 * P + p] = q[2] * inv - m2 * m2;
  stress[3 * P + p] = q[3] * inv - m0 * m1;
  stress[4 * P + p] = q[4] * inv - m0 * m2;
  stress[5 * P + p] = q[5] * inv - m1 * m2;
}

int launch_field_moments(const double* F, long P, long N, int layout, double* mean, double* stress, cudaStream_t st) {
  if (layout == 1) field_moments_kernel<1><<<ceil_div(P, 128L), 128, 0, st>>>(F, P, N, mean, stress);
  else field_moments_kernel<0><<<ceil_div(P, 4L), 128, 0, st>>>(F, P, N, mean, stress);
  count_launch();
  PODSDF_CUDA(cudaGetLastError());
  return 0;
}

constexpr int LAGC_MAX = 16;
struct LagList {
  int dt[LAGC_MAX], dj[LAGC_MAX], dk[LAGC_MAX];
  int n;
};

// partial[block][lag] = sum over this block's (t, j, k) of a'(t,j,k) a'(t+dt, j+dj, k+dk); counts are analytic
template <int LAYOUT>
__global__ void __launch_bounds__(256)
lag_covariance_kernel(const double* __restrict__ F, const double* __restrict__ mean, const long J, const long K,
                      const long N, const int comp, const LagList lags, double* __restrict__ partial) {
  __shared__ double red[256];
  const long P = J * K;
  const long total = N * P;
  double acc[LAGC_MAX];
#pragma unroll
  for (int l = 0; l < LAGC_MAX; ++l) acc[l] = 0.0;
  auto val = [&](long t, long p) {
    const double m = mean ? mean[comp * P + p] : 0.0;
    return ((LAYOUT == 1) ? F[t * 3 * P + comp * P + p] : F[(comp * P + p) * N + t]) - m;
  };
  for (long e = blockIdx.x * 256L + threadIdx.x; e < total; e += (long)gridDim.x * 256) {
    const long t = (LAYOUT == 1) ? e / P : e % N;          // the contiguous index varies fastest across threads
    const long p = (LAYOUT == 1) ? e % P : e / N;
    const long j = p / K, k = p % K;
    const double a = val(t, p);
#pragma unroll
    for (int l = 0; l < LAGC_MAX; ++l) {
      if (l < lags.n) {
        const long t2 = t + lags.dt[l], j2 = j + lags.dj[l], k2 = k + lags.dk[l];
        if (t2 < N && j2 < J && k2 < K) acc[l] = fma(a, val(t2, j2 * K + k2), acc[l]);
      }
    }
  }
  for (int l = 0; l < lags.n; ++l) {
    red[threadIdx.x] = acc[l];
    __syncthreads();
    for (int s2 = 128; s2 > 0; s2 >>= 1) {
      if ((int)threadIdx.x < s2) red[threadIdx.x] += red[threadIdx.x + s2];
      __syncthreads();
    }
    if (threadIdx.x == 0) partial[(long)blockIdx.x * LAGC_MAX + l] = red[0];
    __syncthreads();
  }
}

__global__ void lag_covariance_final_kernel(const double* __restrict__ partial, const int nblocks, const LagList lags,
                                            const long J, const long K, const long N, double* __restrict__ out) {
  const int l = threadIdx.x;
  if (l >= lags.n) return;
  double sacc = 0.0;
  for (int b = 0; b < nblocks; ++b) sacc += partial[(long)b * LAGC_MAX + l];       // fixed order
  const double cnt = (double)(N - lags.dt[l]) * (double)(J - lags.dj[l]) * (double)(K - lags.dk[l]);
  out[l] = cnt > 0.0 ? sacc / cnt : 0.0;
}

int lag_covariance_max() { return LAGC_MAX; }
size_t lag_covariance_workspace_doubles() { return (size_t)1184 * LAGC_MAX + LAGC_MAX; }

int launch_lag_covariance(const double* F, long J, long K, long N, int layout, const double* mean, int comp,
                          const int* lags3, int nl, double* ws, double* out_dev, cudaStream_t st) {
  if (nl < 1 || nl > LAGC_MAX) return fail(PODSDF_ERR_INVALID, "1..16 lags per call");
  LagList ll;
  ll.n = nl;
  for (int l = 0; l < LAGC_MAX; ++l) {
    ll.dt[l] = l < nl ? lags3[3 * l] : 0;
    ll.dj[l] = l < nl ? lags3[3 * l + 1] : 0;
    ll.dk[l] = l < nl ? lags3[3 * l + 2] : 0;
    if (ll.dt[l] < 0 || ll.dj[l] < 0 || ll.dk[l] < 0) return fail(PODSDF_ERR_INVALID, "lags must be non-negative");
  }
  const long total = N * J * K;
  int nblocks = (int)std::min<long>(1184, ceil_div(total, 256L));    // 148 SMs x 8
  if (layout == 1) lag_covariance_kernel<1><<<nblocks, 256, 0, st>>>(F, mean, J, K, N, comp, ll, ws);
  else lag_covariance_kernel<0><<<nblocks, 256, 0, st>>>(F, mean, J, K, N, comp, ll, ws);
  lag_covariance_final_kernel<<<1, 32, 0, st>>>(ws, nblocks, ll, J, K, N, out_dev);
  count_launch(2);
  PODSDF_CUDA(cudaGetLastError());
  return 0;
}

// Packed upper triangle of a symmetric N x N matrix (row-major rows i, columns j >= i): the all-reduce of the
// per-rank Gram matrices moves N (N + 1) / 2 doubles instead of N^2.
__global__ void __launch_bounds__(256)
sym_pack_kernel(const double* __restrict__ C, const long N, double* __restrict__ packed, const int unpack,
                double* __restrict__ Cout) {
  const long i = blockIdx.y;
  const long j = i + blockIdx.x * 256L + threadIdx.x;
  if (j >= N) return;
  const long off = i * N - i * (i - 1) / 2 + (j - i);     // rows 0..i-1 hold N, N-1, ..., N-i+1 entries
  if (!unpack) {
    packed[off] = C[i * N + j];
  } else {
    const double v = packed[off];
    Cout[i * N + j] = v;
    Cout[j * N + i] = v;
  }
}

int launch_sym_pack(const double* C, long N, double* packed, cudaStream_t st) {
  sym_pack_kernel<<<dim3((unsigned)ceil_div(N, 256L), (unsigned)N), 256, 0, st>>>(C, N, packed, 0, nullptr);
  count_launch();
  PODSDF_CUDA(cudaGetLastError());
  return 0;
}
int launch_sym_unpack(const double* packed, long N, double* C, cudaStream_t st) {
  sym_pack_kernel<<<dim3((unsigned)ceil_div(N, 256L), (unsigned)N), 256, 0, st>>>(nullptr, N, const_cast<double*>(packed), 1, C);
  count_launch();
  PODSDF_CUDA(cudaGetLastError());
  return 0;
}

// out[c][r] = in[r][c]  (n x n), used to hand eigenvectors back as columns
__global__ void __launch_bounds__(256) transpose_kernel(const double* __restrict__ in, double* __restrict__ out, const long n) {
  __shared__ double tile[32][33];
  const long bx = blockIdx.x * 32L, by = blockIdx.y * 32L;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int r = ty; r < 32; r += 8)
    if (by + r < n && bx + tx < n) tile[r][tx] = in[(by + r) * n + bx + tx];
  __syncthreads();
  for (int r = ty; r < 32; r += 8)
    if (bx + r < n && by + tx < n) out[(bx + r) * n + by + tx] = tile[tx][r];
}

int launch_transpose(const double* in, double* out, long n, cudaStream_t st) {
  dim3 grid(ceil_div(n, 32), ceil_div(n, 32));
  transpose_kernel<<<grid, 256, 0, st>>>(in, out, n);
  count_launch();
  PODSDF_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace podsdf
