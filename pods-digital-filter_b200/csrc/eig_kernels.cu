// Symmetric tridiagonal eigen-solver pieces for "all eigenvalues + the nm leading eigenvectors" (SURVEY 8(f)-3):
// the dense matrix is reduced to tridiagonal form by cuSOLVER (Dsytrd); here are
//   * bisection on the Sturm sequence for ALL N eigenvalues (one thread per eigenvalue, LAPACK dstebz's recurrence
//     with its pivmin guard),
//   * inverse iteration for the nm largest (one thread per vector: LU of T - lambda I with partial pivoting as in
//     LAPACK dlagtf/dlagts, three solves, residual reported),
//   * Cholesky-QR of the nm back-transformed vectors (restores orthonormality to rounding; the leading vectors keep
//     their direction, like modified Gram-Schmidt in descending order of the eigenvalue).
// PODFS only consumes the nm leading temporal modes (PODFS.py:1329-1333), the full eigenvalue list feeds the energy
// spectrum and the validity count (PODFS.py:1312-1320).
#include <cfloat>

#include "eig_kernels.cuh"

#include <cstdlib>

namespace podsdf {

// bounds[0] = gl, [1] = gu (Gershgorin interval, widened as dstebz does), [2] = pivmin, [3] = |T| (max-norm bound)
__global__ void __launch_bounds__(1024)
tridiag_prepare_kernel(const double* __restrict__ d, const double* __restrict__ e, const long N, double* __restrict__ e2,
                       double* __restrict__ bounds) {
  __shared__ double s_lo[1024], s_hi[1024], s_e2[1024];
  double lo = DBL_MAX, hi = -DBL_MAX, me2 = 0.0;
  for (long i = threadIdx.x; i < N; i += 1024) {
    const double el = (i > 0) ? fabs(e[i - 1]) : 0.0, er = (i + 1 < N) ? fabs(e[i]) : 0.0;
    lo = fmin(lo, d[i] - el - er);
    hi = fmax(hi, d[i] + el + er);
    if (i + 1 < N) {
      const double v = e[i] * e[i];
      e2[i] = v;
      me2 = fmax(me2, v);
    }
  }
  s_lo[threadIdx.x] = lo;
  s_hi[threadIdx.x] = hi;
  s_e2[threadIdx.x] = me2;
  __syncthreads();
  for (int s = 512; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) {
      s_lo[threadIdx.x] = fmin(s_lo[threadIdx.x], s_lo[threadIdx.x + s]);
      s_hi[threadIdx.x] = fmax(s_hi[threadIdx.x], s_hi[threadIdx.x + s]);
      s_e2[threadIdx.x] = fmax(s_e2[threadIdx.x], s_e2[threadIdx.x + s]);
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const double gl = s_lo[0], gu = s_hi[0];
    const double tnorm = fmax(fabs(gl), fabs(gu));
    const double pivmin = DBL_MIN * fmax(1.0, s_e2[0]);
    const double widen = 2.1 * tnorm * DBL_EPSILON * (double)N + 4.2 * pivmin;
    bounds[0] = gl - widen;
    bounds[1] = gu + widen;
    bounds[2] = pivmin;
    bounds[3] = tnorm;
  }
}

// number of eigenvalues of T smaller than x
__device__ __forceinline__ int sturm_count(const double* __restrict__ d, const double* __restrict__ e2, const long N,
                                           const double x, const double pivmin) {
  double q = __ldg(d) - x;
  if (fabs(q) < pivmin) q = -pivmin;
  int cnt = q < 0.0;
  for (long i = 1; i < N; ++i) {
    q = __ldg(d + i) - x - __ldg(e2 + i - 1) / q;
    if (fabs(q) < pivmin) q = -pivmin;
    cnt += q < 0.0;
  }
  return cnt;
}

// Eigenvalues k_first <= k < k_end (ascending index), G lanes each: every sweep over the matrix the G lanes count at
// the G interior points of a (G+1)-section of the bracket, so it shrinks (G+1)x per serial Sturm sweep instead of 2x.
// G = 32 (a warp, 11 sweeps) serves the few leading eigenvalues the inverse iteration waits for; G = 4 (23 sweeps
// instead of ~55, four times the threads) serves all the others -- more lanes per eigenvalue there would saturate
// the SMs with divisions (measured).  Tolerances are dstebz's defaults: relative 2 ulp, absolute ulp * |T|.
template <int G>
__global__ void __launch_bounds__(1024)
tridiag_sect_kernel(const double* __restrict__ d, const double* __restrict__ e2, const double* __restrict__ bounds,
                    const long N, const long k_first, const long k_end, double* __restrict__ lam) {
  // eigenvalue handled by this group of lanes (grid-stride: the launch may use a few fat CTAs, see launch_tridiag_bisect)
  for (long gid = ((long)blockIdx.x * blockDim.x + threadIdx.x) / G;; gid += (long)gridDim.x * blockDim.x / G) {
  const long k = k_first + gid;
  const int lane = threadIdx.x & 31;
  const int gl = lane % G;                                         // position inside the group
  const int base = lane - gl;
  const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << base);
  if (k >= k_end) return;                                          // whole groups leave together
  double lo = bounds[0], hi = bounds[1];
  const double pivmin = bounds[2], tnorm = bounds[3];
  for (int it = 0; it < 64; ++it) {
    const double tol = fmax(2.0 * DBL_EPSILON * fmax(fabs(lo), fabs(hi)), DBL_EPSILON * tnorm) + 2.0 * pivmin;
    if (hi - lo <= tol) break;
    const double w = (hi - lo) / (double)(G + 1);
    const double x = lo + (double)(gl + 1) * w;
    const bool inside = (x > lo) && (x < hi);                      // false once the points collapse onto the ends
    const int c = inside ? sturm_count(d, e2, N, x, pivmin) : (x >= hi ? (int)N : 0);
    const unsigned above = (__ballot_sync(gmask, c > k) >> base) & ((G == 32) ? 0xffffffffu : ((1u << G) - 1u));
    const unsigned any_inside = __ballot_sync(gmask, inside) & gmask;
    if (any_inside == 0u) break;
    double nlo = lo, nhi = hi;
    if (above == 0u) {
      nlo = __shfl_sync(gmask, x, base + G - 1);
    } else {
      const int j = __ffs(above) - 1;
      nhi = __shfl_sync(gmask, x, base + j);
      const double below = __shfl_sync(gmask, x, base + (j > 0 ? j - 1 : 0));
      if (j > 0) nlo = below;
    }
    nlo = fmin(fmax(nlo, lo), hi);
    nhi = fmax(fmin(nhi, hi), nlo);
    if (nlo == lo && nhi == hi) break;                             // no progress: the bracket is at rounding level
    lo = nlo;
    hi = nhi;
  }
  if (gl == 0) lam[k] = 0.5 * (lo + hi);
  }
}

// The leading eigenvalues, one CTA of BS_THREADS lanes per eigenvalue: every Sturm sweep shrinks the bracket
// (BS_THREADS + 1)-fold, so seven sweeps do what eleven did with one warp per eigenvalue.  A sweep is a chain of N
// dependent divisions that a warp issues at IPC ~0.25 whatever it does (four chains per lane took 2.3x longer per
// sweep), only nm + 1 eigenvalues are wanted before the inverse iteration can start, and the SMs the side-stream
// bisection leaves alone are idle: the parallelism goes into more points per sweep.  256 lanes: two CTAs fit an SM, so
// nm + 1 = 101 CTAs still run in one wave on the 100 SMs that bisection does not use.
constexpr int BS_THREADS = 256;
__global__ void __launch_bounds__(BS_THREADS)
tridiag_sect_block_kernel(const double* __restrict__ d, const double* __restrict__ e2, const double* __restrict__ bounds,
                          const long N, const long k_first, const long k_end, double* __restrict__ lam) {
  __shared__ unsigned s_above[BS_THREADS / 32];
  __shared__ int s_any;
  const long k = k_first + blockIdx.x;
  if (k >= k_end) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double lo = bounds[0], hi = bounds[1];
  const double pivmin = bounds[2], tnorm = bounds[3];
  for (int it = 0; it < 64; ++it) {
    const double tol = fmax(2.0 * DBL_EPSILON * fmax(fabs(lo), fabs(hi)), DBL_EPSILON * tnorm) + 2.0 * pivmin;
    if (hi - lo <= tol) break;                                     // (uniform: lo, hi are the same in every thread)
    const double w = (hi - lo) / (double)(BS_THREADS + 1);
    const double x = lo + (double)(tid + 1) * w;
    const bool inside = (x > lo) && (x < hi);                      // false once the points collapse onto the ends
    if (tid == 0) s_any = 0;
    __syncthreads();
    if (inside) s_any = 1;
    __syncthreads();
    if (!s_any) break;
    const int c = inside ? sturm_count(d, e2, N, x, pivmin) : (x >= hi ? (int)N : 0);
    const unsigned above = __ballot_sync(0xffffffffu, c > k);
    if (lane == 0) s_above[warp] = above;
    __syncthreads();
    int jfirst = -1;                                               // first point with more than k eigenvalues below it
#pragma unroll
    for (int q = 0; q < BS_THREADS / 32; ++q)
      if (jfirst < 0 && s_above[q] != 0u) jfirst = q * 32 + __ffs(s_above[q]) - 1;
    double nlo, nhi;
    if (jfirst < 0) {
      nlo = lo + (double)BS_THREADS * w;
      nhi = hi;
    } else {
      nhi = lo + (double)(jfirst + 1) * w;
      nlo = (jfirst > 0) ? lo + (double)jfirst * w : lo;
    }
    nlo = fmin(fmax(nlo, lo), hi);
    nhi = fmax(fmin(nhi, hi), nlo);
    __syncthreads();                                               // s_above / s_any are rewritten in the next sweep
    if (nlo == lo && nhi == hi) break;                             // no progress: the bracket is at rounding level
    lo = nlo;
    hi = nhi;
  }
  if (tid == 0) lam[k] = 0.5 * (lo + hi);
}

// Inverse iteration for the nm largest eigenvalues.  CTA j owns vector j (eigenvalue lam[N-1-j]): LU of T - lambda I
// with partial pivoting as in LAPACK dlagtf / dlagts, three solves, residual reported.  The recurrences are serial in i,
// so ONE thread runs them -- out of shared memory: the CTA streams the operands of 512 steps in (coalesced), thread 0
// walks the chunk at register / shared-memory latency, the CTA streams the results out.  (Round 1 ran one thread per
// vector straight on global memory: 19 ms at N = 10^4, nm = 100; the arithmetic alone is ~1 ms.)
// Work arrays: seven planes [vector][i].  Z[j][i] (row-major nm x N) receives the unit eigenvector of T;
// resid[j] = |T z - lambda z|_inf / |T|.
constexpr int IV_CH = 512, IV_THREADS = 128;

__global__ void __launch_bounds__(IV_THREADS)
tridiag_invit_kernel(const double* __restrict__ d, const double* __restrict__ e, const double* __restrict__ bounds,
                     const double* __restrict__ lam, const long N, const int nm, double* __restrict__ wP,
                     double* __restrict__ wQ, double* __restrict__ wR, double* __restrict__ wM, double* __restrict__ wS,
                     double* __restrict__ wX, double* __restrict__ wY, double* __restrict__ Z, double* __restrict__ resid) {
  __shared__ double si[4][IV_CH + 2];                // chunk inputs
  __shared__ double so[5][IV_CH];                    // chunk outputs
  __shared__ double s_red[IV_THREADS];
  __shared__ double s_carry[4];                      // recurrence state between chunks (thread 0 only)
  const int j = blockIdx.x, tid = threadIdx.x;
  if (j >= nm) return;
  const double tnorm = fmax(bounds[3], DBL_MIN);
  const double floor_p = DBL_EPSILON * tnorm;
  const double l = lam[N - 1 - j];
  const long off = (long)j * N;
  wP += off; wQ += off; wR += off; wM += off; wS += off; wX += off; wY += off;

  // ---- LU of (T - l I): row i of U = (P[i], Q[i], R[i]), multiplier M[i], swap flag S[i]
  if (tid == 0) {
    s_carry[0] = d[0] - l;
    s_carry[1] = (N > 1) ? e[0] : 0.0;
  }
  for (long i0 = 0; i0 + 1 < N; i0 += IV_CH) {
    const int cnt = (int)min((long)IV_CH, N - 1 - i0);            // steps i0 .. i0+cnt-1 (each needs row i+1)
    __syncthreads();
    for (int t = tid; t < cnt + 1; t += IV_THREADS) {
      const long i = i0 + t;
      si[0][t] = (i < N - 1) ? e[i] : 0.0;                        // sub-diagonal entry of row i+1
      si[1][t] = (i + 1 < N) ? d[i + 1] - l : 0.0;
      si[2][t] = (i + 2 < N) ? e[i + 1] : 0.0;
    }
    __syncthreads();
    if (tid == 0) {
      double p = s_carry[0], q = s_carry[1];
#pragma unroll 4
      for (int t = 0; t < cnt; ++t) {
        const double c = si[0][t], a1 = si[1][t], b1 = si[2][t];
        if (fabs(p) >= fabs(c)) {
          if (fabs(p) < floor_p) p = copysign(floor_p, p);
          const double m = c / p;
          so[0][t] = p; so[1][t] = q; so[2][t] = 0.0; so[3][t] = m; so[4][t] = 0.0;
          p = a1 - m * q;
          q = b1;
        } else {
          const double m = p / c;
          so[0][t] = c; so[1][t] = a1; so[2][t] = b1; so[3][t] = m; so[4][t] = 1.0;
          p = q - m * a1;
          q = -m * b1;
        }
      }
      s_carry[0] = p;
      s_carry[1] = q;
    }
    __syncthreads();
    for (int t = tid; t < cnt; t += IV_THREADS) {
      const long i = i0 + t;
      wP[i] = so[0][t]; wQ[i] = so[1][t]; wR[i] = so[2][t]; wM[i] = so[3][t]; wS[i] = so[4][t];
    }
  }
  __syncthreads();
  if (tid == 0) {
    double p = s_carry[0];
    if (fabs(p) < floor_p) p = copysign(floor_p, p);
    wP[N - 1] = p; wQ[N - 1] = 0.0; wR[N - 1] = 0.0;
  }
  // ---- start vector: a fixed pseudo-random pattern per (i, j)
  for (long i = tid; i < N; i += IV_THREADS) {
    unsigned h = (unsigned)(i * 2654435761u) ^ (unsigned)(j * 40503u + 12345u);
    h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
    wX[i] = ((double)(h & 0xFFFFF) + 0.5) / 1048576.0 - 0.5;
  }
  __syncthreads();
  for (int iter = 0; iter < 3; ++iter) {
    // ---- forward sweep X -> Y: apply the row operations to the right-hand side (y carried between chunks)
    if (tid == 0) s_carry[0] = wX[0];
    for (long i0 = 0; i0 + 1 < N; i0 += IV_CH) {
      const int cnt = (int)min((long)IV_CH, N - 1 - i0);
      __syncthreads();
      for (int t = tid; t < cnt; t += IV_THREADS) {
        const long i = i0 + t;
        si[0][t] = wX[i + 1];
        si[1][t] = wS[i];
        si[2][t] = wM[i];
      }
      __syncthreads();
      if (tid == 0) {
        double y = s_carry[0];
#pragma unroll 4
        for (int t = 0; t < cnt; ++t) {
          double y1 = si[0][t];
          if (si[1][t] != 0.0) { const double tmp = y; y = y1; y1 = tmp; }
          y1 -= si[2][t] * y;
          so[0][t] = y;
          y = y1;
        }
        s_carry[0] = y;
      }
      __syncthreads();
      for (int t = tid; t < cnt; t += IV_THREADS) wY[i0 + t] = so[0][t];
    }
    __syncthreads();
    if (tid == 0) {
      wY[N - 1] = s_carry[0];
      s_carry[0] = 0.0;   // x1
      s_carry[1] = 0.0;   // x2
      s_carry[2] = 0.0;   // big
    }
    // ---- back substitution Y -> X with two super-diagonals (descending chunks)
    for (long hi = N; hi > 0; hi -= IV_CH) {
      const long lo = max(0L, hi - IV_CH);
      const int cnt = (int)(hi - lo);
      __syncthreads();
      for (int t = tid; t < cnt; t += IV_THREADS) {
        const long i = lo + t;
        si[0][t] = wY[i];
        si[1][t] = wQ[i];
        si[2][t] = wR[i];
        si[3][t] = wP[i];
      }
      __syncthreads();
      if (tid == 0) {
        double x1 = s_carry[0], x2 = s_carry[1], big = s_carry[2];
#pragma unroll 4
        for (int t = cnt - 1; t >= 0; --t) {
          const double x0 = (si[0][t] - si[1][t] * x1 - si[2][t] * x2) / si[3][t];
          so[0][t] = x0;
          big = fmax(big, fabs(x0));
          x2 = x1;
          x1 = x0;
        }
        s_carry[0] = x1; s_carry[1] = x2; s_carry[2] = big;
      }
      __syncthreads();
      for (int t = tid; t < cnt; t += IV_THREADS) wX[lo + t] = so[0][t];
    }
    __syncthreads();
    const double sc = 1.0 / fmax(s_carry[2], DBL_MIN);
    for (long i = tid; i < N; i += IV_THREADS) wX[i] *= sc;
    __syncthreads();
  }
  // ---- unit 2-norm, residual, write the row of Z
  double ss = 0.0;
  for (long i = tid; i < N; i += IV_THREADS) ss = fma(wX[i], wX[i], ss);
  s_red[tid] = ss;
  __syncthreads();
  for (int st = IV_THREADS / 2; st > 0; st >>= 1) {
    if (tid < st) s_red[tid] += s_red[tid + st];
    __syncthreads();
  }
  const double sn = 1.0 / sqrt(s_red[0]);
  __syncthreads();
  double r = 0.0;
  for (long i = tid; i < N; i += IV_THREADS) {
    const double x0 = wX[i] * sn;
    const double xm = (i > 0) ? wX[i - 1] * sn : 0.0, xp = (i + 1 < N) ? wX[i + 1] * sn : 0.0;
    const double t = ((i > 0) ? __ldg(e + i - 1) * xm : 0.0) + (__ldg(d + i) - l) * x0 + ((i + 1 < N) ? __ldg(e + i) * xp : 0.0);
    r = fmax(r, fabs(t));
    Z[(long)j * N + i] = x0;
  }
  s_red[tid] = r;
  __syncthreads();
  for (int st = IV_THREADS / 2; st > 0; st >>= 1) {
    if (tid < st) s_red[tid] = fmax(s_red[tid], s_red[tid + st]);
    __syncthreads();
  }
  if (tid == 0) resid[j] = s_red[0] / tnorm;
}

// S[a][b] = <Z_a, Z_b>
__global__ void __launch_bounds__(256)
rows_gram_kernel(const double* __restrict__ Z, const long N, const int nm, double* __restrict__ S) {
  const int a = blockIdx.x, b = blockIdx.y;
  if (b > a) return;
  __shared__ double red[256];
  double acc = 0.0;
  for (long i = threadIdx.x; i < N; i += 256) acc = fma(Z[(long)a * N + i], Z[(long)b * N + i], acc);
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) S[a * nm + b] = S[b * nm + a] = red[0];
}

// in-place lower Cholesky factor of S (nm x nm, one block); flag[0] = 1 if a pivot is not positive
__global__ void __launch_bounds__(256)
small_cholesky_kernel(double* __restrict__ S, const int nm, int* __restrict__ flag) {
  for (int k = 0; k < nm; ++k) {
    if (threadIdx.x == 0) {
      const double v = S[k * nm + k];
      if (!(v > 0.0)) flag[0] = 1;
      S[k * nm + k] = sqrt(fmax(v, DBL_MIN));
    }
    __syncthreads();
    const double dk = S[k * nm + k];
    for (int i = k + 1 + threadIdx.x; i < nm; i += 256) S[i * nm + k] /= dk;
    __syncthreads();
    const int rem = nm - k - 1;
    for (int t = threadIdx.x; t < rem * rem; t += 256) {
      const int i = k + 1 + t / rem, jj = k + 1 + t % rem;
      if (jj <= i) S[i * nm + jj] -= S[i * nm + k] * S[jj * nm + k];
    }
    __syncthreads();
  }
}

// Z <- L^{-1} Z, column by column (thread = column i): the rows become orthonormal
__global__ void __launch_bounds__(128)
rows_apply_inverse_kernel(double* __restrict__ Z, const long N, const int nm, const double* __restrict__ L) {
  const long i = blockIdx.x * 128L + threadIdx.x;
  if (i >= N) return;
  for (int a = 0; a < nm; ++a) {
    double v = Z[(long)a * N + i];
    for (int b = 0; b < a; ++b) v = fma(-__ldg(L + a * nm + b), Z[(long)b * N + i], v);
    Z[(long)a * N + i] = v / __ldg(L + a * nm + a);
  }
}

// Z[j][:] = C[N-1-j][:] for j < nm (the full solver returns ascending rows; callers want the leading ones first)
__global__ void __launch_bounds__(256)
rows_gather_descending_kernel(const double* __restrict__ C, const long N, const int nm, double* __restrict__ Z) {
  const long i = blockIdx.x * 256L + threadIdx.x;
  const int j = blockIdx.y;
  if (i < N) Z[(long)j * N + i] = C[(N - 1 - j) * N + i];
}

int launch_rows_gather_descending(const double* C, long N, int nm, double* Z, cudaStream_t st) {
  rows_gather_descending_kernel<<<dim3((unsigned)ceil_div(N, 256L), nm), 256, 0, st>>>(C, N, nm, Z);
  count_launch();
  PODSDF_CUDA(cudaGetLastError());
  return 0;
}

int launch_tridiag_prepare(const double* d, const double* e, long N, double* e2, double* bounds, cudaStream_t st) {
  tridiag_prepare_kernel<<<1, 1024, 0, st>>>(d, e, N, e2, bounds);
  count_launch();
  PODSDF_CUDA(cudaGetLastError());
  return 0;
}

int launch_tridiag_bisect(const double* d, const double* e2, const double* bounds, long N, long k_end, double* lam,
                          cudaStream_t st, int reserve_sms) {
  if (k_end <= 0) return 0;
  constexpr int G = 4;
  // This kernel runs on a side stream beside the inverse iteration and the back-transformation of the leading vectors,
  // whose CTAs need whole SMs (512 threads x 128 registers).  Its own CTAs live for the whole ~20 ms (a chain of
  // dependent divisions), so small CTAs spread over every SM keep those kernels waiting for registers (measured: 11 ms
  // of the 22 ms it should hide).  With `reserve_sms` > 0 it is packed into a few 1024-thread CTAs -- at most one per SM
  // that the others do not need -- and strides over its eigenvalues.
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const long lanes = k_end * G;
  if (reserve_sms > 0 && sms - reserve_sms >= 16 && lanes > 64L * sms) {
    const long fat = std::min<long>(sms - reserve_sms, ceil_div(lanes, 1024L));
    tridiag_sect_kernel<G><<<(unsigned)fat, 1024, 0, st>>>(d, e2, bounds, N, 0, k_end, lam);
  } else {
    tridiag_sect_kernel<G><<<(unsigned)ceil_div(lanes, 64L), 64, 0, st>>>(d, e2, bounds, N, 0, k_end, lam);
  }
  count_launch();
  PODSDF_CUDA(cudaGetLastError());
  return 0;
}

int launch_tridiag_multisect(const double* d, const double* e2, const double* bounds, long N, long k_first, double* lam,
                             cudaStream_t st) {
  if (k_first >= N) return 0;
  static const bool warp_version = getenv("PODSDF_EIG_MULTISECT_WARP") != nullptr;      // A/B: one warp per eigenvalue
  if (warp_version)
    tridiag_sect_kernel<32><<<(unsigned)ceil_div((N - k_first) * 32, 64L), 64, 0, st>>>(d, e2, bounds, N, k_first, N, lam);
  else
    tridiag_sect_block_kernel<<<(unsigned)(N - k_first), BS_THREADS, 0, st>>>(d, e2, bounds, N, k_first, N, lam);
  count_launch();
  PODSDF_CUDA(cudaGetLastError());
  return 0;
}

size_t tridiag_invit_workspace_doubles(long N, int nm) { return (size_t)7 * N * nm; }

int launch_tridiag_invit(const double* d, const double* e, const double* bounds, const double* lam, long N, int nm,
                         double* ws, double* Z, double* resid, cudaStream_t st) {
  const size_t pl = (size_t)N * nm;
  tridiag_invit_kernel<<<(unsigned)nm, IV_THREADS, 0, st>>>(d, e, bounds, lam, N, nm, ws, ws + pl, ws + 2 * pl, ws + 3 * pl,
                                                                  ws + 4 * pl, ws + 5 * pl, ws + 6 * pl, Z, resid);
  count_launch();
  PODSDF_CUDA(cudaGetLastError());
  return 0;
}

int launch_orthonormalize_rows(double* Z, long N, int nm, double* S, int* flag, cudaStream_t st) {
  PODSDF_CUDA(cudaMemsetAsync(flag, 0, sizeof(int), st));
  rows_gram_kernel<<<dim3(nm, nm), 256, 0, st>>>(Z, N, nm, S);
  small_cholesky_kernel<<<1, 256, 0, st>>>(S, nm, flag);
  rows_apply_inverse_kernel<<<(unsigned)ceil_div(N, 128L), 128, 0, st>>>(Z, N, nm, S);
  count_launch(3);
  PODSDF_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace podsdf
