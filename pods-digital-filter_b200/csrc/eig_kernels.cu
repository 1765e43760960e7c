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
__global__ void __launch_bounds__(64)
tridiag_sect_kernel(const double* __restrict__ d, const double* __restrict__ e2, const double* __restrict__ bounds,
                    const long N, const long k_first, const long k_end, double* __restrict__ lam) {
  const long gid = (blockIdx.x * 64L + threadIdx.x) / G;          // eigenvalue handled by this group of lanes
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

// Inverse iteration for the nm largest eigenvalues.  Thread j owns vector j (eigenvalue lam[N-1-j]); its seven
// work arrays are separate [i][j] planes (consecutive threads touch consecutive doubles, and the compiler may
// hoist the loads of an unrolled stretch above the serial recurrence because no plane is read after it is written
// inside one sweep).  Z[j][i] (row-major nm x N) receives the unit eigenvector of T;
// resid[j] = |T z - lambda z|_inf / |T|.
__global__ void __launch_bounds__(32)
tridiag_invit_kernel(const double* __restrict__ d, const double* __restrict__ e, const double* __restrict__ bounds,
                     const double* __restrict__ lam, const long N, const int nm, double* __restrict__ wP,
                     double* __restrict__ wQ, double* __restrict__ wR, double* __restrict__ wM, double* __restrict__ wS,
                     double* __restrict__ wX, double* __restrict__ wY, double* __restrict__ Z, double* __restrict__ resid) {
  const int j = blockIdx.x * 32 + threadIdx.x;
  if (j >= nm) return;
  const double tnorm = fmax(bounds[3], DBL_MIN);
  const double floor_p = DBL_EPSILON * tnorm;
  const double l = lam[N - 1 - j];
  const long st = nm;                                    // stride between consecutive i of one plane
  wP += j; wQ += j; wR += j; wM += j; wS += j; wX += j; wY += j;
  // Every sweep below is a serial recurrence over i whose operands do not depend on the recurrence: they are
  // fetched U steps at a time into registers first, so one memory latency is paid per U steps instead of per step.
  constexpr int U = 8;
  // ---- LU of (T - l I) with partial pivoting: row i of U = (P[i], Q[i], R[i]), multiplier M[i], swap flag S[i]
  double p = __ldg(d) - l, q = (N > 1) ? __ldg(e) : 0.0;
  for (long i0 = 0; i0 + 1 < N; i0 += U) {
    double cc[U], aa[U], bb[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long i = i0 + u;
      const bool ok = i + 1 < N;
      cc[u] = ok ? __ldg(e + i) : 0.0;                   // sub-diagonal entry of row i+1
      aa[u] = ok ? __ldg(d + i + 1) - l : 0.0;
      bb[u] = (i + 2 < N) ? __ldg(e + i + 1) : 0.0;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long i = i0 + u;
      if (i + 1 < N) {
        const double c = cc[u], a1 = aa[u], b1 = bb[u];
        if (fabs(p) >= fabs(c)) {
          if (fabs(p) < floor_p) p = copysign(floor_p, p);
          const double m = c / p;
          wP[i * st] = p; wQ[i * st] = q; wR[i * st] = 0.0; wM[i * st] = m; wS[i * st] = 0.0;
          p = a1 - m * q;
          q = b1;
        } else {
          const double m = p / c;
          wP[i * st] = c; wQ[i * st] = a1; wR[i * st] = b1; wM[i * st] = m; wS[i * st] = 1.0;
          p = q - m * a1;
          q = -m * b1;
        }
      }
    }
  }
  if (fabs(p) < floor_p) p = copysign(floor_p, p);
  wP[(N - 1) * st] = p; wQ[(N - 1) * st] = 0.0; wR[(N - 1) * st] = 0.0;
  // ---- start vector: a fixed pseudo-random pattern per (i, j)
  for (long i = 0; i < N; ++i) {
    unsigned h = (unsigned)(i * 2654435761u) ^ (unsigned)(j * 40503u + 12345u);
    h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
    wX[i * st] = ((double)(h & 0xFFFFF) + 0.5) / 1048576.0 - 0.5;
  }
  for (int iter = 0; iter < 3; ++iter) {
    // forward sweep X -> Y: apply the row operations to the right-hand side (y carried in a register)
    double y = wX[0];
    for (long i0 = 0; i0 + 1 < N; i0 += U) {
      double xs[U], ss[U], ms[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long i = i0 + u;
        const bool ok = i + 1 < N;
        xs[u] = ok ? wX[(i + 1) * st] : 0.0;
        ss[u] = ok ? wS[i * st] : 0.0;
        ms[u] = ok ? wM[i * st] : 0.0;
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long i = i0 + u;
        if (i + 1 < N) {
          double y1 = xs[u];
          if (ss[u] != 0.0) { const double t = y; y = y1; y1 = t; }
          y1 -= ms[u] * y;
          wY[i * st] = y;
          y = y1;
        }
      }
    }
    wY[(N - 1) * st] = y;
    // back substitution Y -> X with two super-diagonals
    double x1 = 0.0, x2 = 0.0, big = 0.0;
    for (long i0 = N - 1; i0 >= 0; i0 -= U) {
      double ys[U], qs[U], rs[U], ps[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long i = i0 - u;
        const bool ok = i >= 0;
        ys[u] = ok ? wY[i * st] : 0.0;
        qs[u] = ok ? wQ[i * st] : 0.0;
        rs[u] = ok ? wR[i * st] : 0.0;
        ps[u] = ok ? wP[i * st] : 1.0;
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long i = i0 - u;
        if (i >= 0) {
          const double x0 = (ys[u] - qs[u] * x1 - rs[u] * x2) / ps[u];
          wX[i * st] = x0;
          big = fmax(big, fabs(x0));
          x2 = x1;
          x1 = x0;
        }
      }
    }
    const double s = 1.0 / fmax(big, DBL_MIN);
    for (long i = 0; i < N; ++i) wX[i * st] *= s;
  }
  // ---- unit 2-norm, residual, write the row of Z
  double ss = 0.0;
  for (long i = 0; i < N; ++i) ss = fma(wX[i * st], wX[i * st], ss);
  const double s = 1.0 / sqrt(ss);
  double r = 0.0, xm = 0.0, x0 = wX[0] * s;
  for (long i = 0; i < N; ++i) {
    const double xp = (i + 1 < N) ? wX[(i + 1) * st] * s : 0.0;
    const double t = ((i > 0) ? __ldg(e + i - 1) * xm : 0.0) + (__ldg(d + i) - l) * x0 + ((i + 1 < N) ? __ldg(e + i) * xp : 0.0);
    r = fmax(r, fabs(t));
    Z[(long)j * N + i] = x0;
    xm = x0;
    x0 = xp;
  }
  resid[j] = r / tnorm;
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
                          cudaStream_t st) {
  if (k_end <= 0) return 0;
  constexpr int G = 4;
  tridiag_sect_kernel<G><<<(unsigned)ceil_div(k_end * G, 64L), 64, 0, st>>>(d, e2, bounds, N, 0, k_end, lam);
  count_launch();
  PODSDF_CUDA(cudaGetLastError());
  return 0;
}

int launch_tridiag_multisect(const double* d, const double* e2, const double* bounds, long N, long k_first, double* lam,
                             cudaStream_t st) {
  if (k_first >= N) return 0;
  tridiag_sect_kernel<32><<<(unsigned)ceil_div((N - k_first) * 32, 64L), 64, 0, st>>>(d, e2, bounds, N, k_first, N, lam);
  count_launch();
  PODSDF_CUDA(cudaGetLastError());
  return 0;
}

size_t tridiag_invit_workspace_doubles(long N, int nm) { return (size_t)7 * N * nm; }

int launch_tridiag_invit(const double* d, const double* e, const double* bounds, const double* lam, long N, int nm,
                         double* ws, double* Z, double* resid, cudaStream_t st) {
  const size_t pl = (size_t)N * nm;
  tridiag_invit_kernel<<<(unsigned)ceil_div(nm, 32), 32, 0, st>>>(d, e, bounds, lam, N, nm, ws, ws + pl, ws + 2 * pl, ws + 3 * pl,
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
