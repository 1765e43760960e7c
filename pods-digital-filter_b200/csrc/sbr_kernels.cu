// Two-stage symmetric tridiagonalisation on sm_100a (successive band reduction) + back-transformation.
//
// Reference behaviour being replaced (file:line under /root/reference): PODFS.py:1309 `linalg.eig(C)` -- LAPACK's
// dense eigen-solver, of which PODFS consumes all eigenvalues (PODFS.py:1312-1320, 1409-1427) and the nm leading
// eigenvectors (PODFS.py:1323-1333).  Round 1 reduced C with cuSOLVER's one-stage Dsytrd (half of its flops are
// matrix-vector products: 0.64 s at N = 10^4); here:
//
//   stage 1  dense -> band of half-width b = 32.  Per 32-column panel:
//              QR of the panel below the band    one cooperative kernel, thread = panel row (registers), one exchange
//                                                of partial sums per column through flagged mailbox entries; the T
//                                                factor of the block reflector by a one-warp kernel behind it
//              X = A22 V                         FP64 tensor-core MMA (mma.sync m8n8k4), lower triangle read in both
//                                                orientations, split over K
//              W = X T - 1/2 V (T^T V^T X T)     two small kernels
//              A22 -= V W^T + W V^T              FP64 tensor-core MMA rank-2b update, lower tiles + full diagonal tiles
//   stage 2  band -> tridiagonal by bulge chasing (Murata-Horikoshi / Lang): sweep i = chain of b x b tasks, one CTA per
//            sweep in a persistent cooperative kernel: two groups of four warps take the even / odd tasks (a task's
//            successor starts when its reflector is known); of the 32 band columns of a task, 31 come through the band
//            one task early and the one it actually waits for through flagged mailbox entries, posted mid-task
//   back     eigenvectors of T -> eigenvectors of C: stage-2 reflectors (reverse order), then the block reflectors of
//            stage 1, one CTA per vector with the vector in shared memory
//
// Storage: A is N x N with element (r, c) at A[c*N + r] (symmetric, so row- and column-major coincide); skinny
// matrices are k-major: V[m*N + r] is entry (r, m).  tools/proto_sbr.py is the numpy model of every step below.
#include "sbr_kernels.cuh"

#include <algorithm>
#include <atomic>
#include <vector>
#include <cstdio>

#include <cfloat>
#include <cmath>
#include <cstdlib>

namespace podsdf {

namespace {

constexpr int B = kSbrB;
constexpr int LDB = kSbrLdb;

__device__ __forceinline__ void dmma884s(double& c0, double& c1, const double a, const double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}
__device__ __forceinline__ void cpa8(void* smem_dst, const void* gsrc, bool valid) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  const int bytes = valid ? 8 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(d), "l"(gsrc), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cpa16(void* smem_dst, const void* gsrc, bool valid) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  const int bytes = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(d), "l"(gsrc), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cpa_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N_>
__device__ __forceinline__ void cpa_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N_) : "memory");
}

// TMA bulk copies (global -> shared, completion on an mbarrier): one instruction moves a whole operand row
__device__ __forceinline__ void mbar_init(unsigned mbar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(mbar), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned mbar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(mbar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned mbar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(mbar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void* src, unsigned bytes, unsigned mbar) {
  asm volatile("cp.async.bulk.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(mbar) : "memory");
}

// ---- acquire / release accesses of the progress counters and the error flag (the cooperative kernels spin on them; all
// their CTAs are co-resident, and every spin is bounded: a lost CTA raises the error flag instead of hanging the GPU)
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_u32(unsigned* p, unsigned v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;\n" ::"l"(p), "r"(v) : "memory");
}

// ---- flagged 16-byte mailbox entries {lo, flag, hi, flag}: value and "it is there" travel in ONE store and ONE load,
// each 8-byte half validated by its own copy of the flag (the low-latency protocol of collective libraries).  Replaces
// "store, fence, count, poll the counter, then load" (three L2 round trips) by one; a flag is never reused: it carries
// the launch's epoch.
#ifdef PODSDF_LL_RELAXED
#define PODSDF_LL_ST "st.relaxed.gpu.global.v4.u32"
#define PODSDF_LL_LD "ld.relaxed.gpu.global.v4.u32"
#else
#define PODSDF_LL_ST "st.volatile.global.v4.u32"
#define PODSDF_LL_LD "ld.volatile.global.v4.u32"
#endif
__device__ __forceinline__ void ll_store(uint4* p, const double v, const unsigned flag) {
  const unsigned lo = (unsigned)__double2loint(v), hi = (unsigned)__double2hiint(v);
  asm volatile(PODSDF_LL_ST " [%0], {%1, %2, %3, %4};\n" ::"l"(p), "r"(lo), "r"(flag), "r"(hi), "r"(flag)
               : "memory");
}
__device__ __forceinline__ bool ll_try(const uint4* p, const unsigned flag, double& v) {
  unsigned a, b, c, d;
  asm volatile(PODSDF_LL_LD " {%0, %1, %2, %3}, [%4];\n" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "l"(p) : "memory");
  v = __hiloint2double((int)c, (int)a);
  return b == flag && d == flag;
}

// ------------------------------------------------------------------------------------------------
// Stage 1a: Householder QR of the panel A[r0:N, c0:c0+B] (n = N - r0 rows), LAPACK dgeqr2 in one pass.
//   Column j needs only ONE global reduction: with S_m = sum_{rows > j} P[r][j] P[r][m] (m >= j) every thread derives
//   beta, tau, v and the update of its entries (v^T P[:, m] = P[j][m] + S_m / (alpha - beta)).  The same reduction
//   carries g_i = V[:, i]^T v_{j-1} (i < j-1) for the T factor (dlarft), which sbr_tfactor_kernel builds off the
//   critical path from Gq = [g | tau].
//   CTA = 256 panel rows x 4 column octets = 1024 threads: thread = (row, 8 columns) in registers, warp = 32 rows of
//   one octet.  The chain of 33 steps is latency-bound, so the work per thread and step is kept short (8 products,
//   9 double shuffles by recursive halving) and spread over 32 warps; a thread-per-row version (32 products, 31
//   shuffles, 8 warps) spent 3.8 k cycles in each of the two compute phases of a step (profiles/r02_sbr_profile.md).
//   The CTAs exchange their partial sums and row j through mailbox entries mail[j & 1][cta | G][B]: every CTA reads
//   all of them and adds them up in the same order -> identical reflectors everywhere, no grid barrier.  A CTA can
//   write parity (j & 1) again only after everybody has published step j + 1, i.e. has finished reading step j.
// Outputs: R (upper triangle) back into A, V rows c0 .. c0+B-1 of Vall (explicit ones / zeros), Gq (B x B + B).
// ------------------------------------------------------------------------------------------------
constexpr int QR_OCT = 8;                    // columns per thread
// panel rows per CTA: 256 (1024 threads) for tall panels; shorter panels take 128 or 64 rows per CTA -- the steps are a
// latency chain, so fewer rows per SM shorten it as long as the number of mailboxes to read stays small
constexpr int QR_ROWS_MAX = 256;
static inline int qr_rows_for(long n) { return n > 8192 ? 256 : (n > 2048 ? 128 : 64); }

// The step's work on one thread's 8 columns depends on where column j lies relative to the octet: D = j - 8 * octet,
// clamped to -2 (octet entirely right of j + 1), -1 .. 8, 9 (entirely left of j - 1).  D is warp-uniform, so the
// kernel switches over 12 compile-time variants instead of predicating every column on j (which cost 670 instructions
// per warp and step, 5 % of them floating point: profiles/r02_sbr_profile.md).
template <int D>
__device__ __forceinline__ void qr_terms(const double (&p)[QR_OCT], const double fa, const double fb, double (&t)[QR_OCT]) {
#pragma unroll
  for (int s = 0; s < QR_OCT; ++s) {
    if (D < 0 || (D <= 8 && s >= D)) t[s] = fa * p[s];             // column >= j: dot with column j
    else if (D == 9 || s < D - 1) t[s] = fb * p[s];               // column < j - 1: V[:, i]^T v_{j-1}
    else t[s] = 0.0;                                              // column j - 1
  }
}
template <int D>
__device__ __forceinline__ void qr_update(double (&p)[QR_OCT], const bool mine, const bool top, const double v,
                                          const double tv, const double beta, const double scale,
                                          const double2* __restrict__ sr, double* __restrict__ next_col) {
  if (mine) {
#pragma unroll
    for (int s = 0; s < QR_OCT; ++s) {
      if (D < 0 || (D <= 8 && s > D)) {
        const double2 x = sr[s];                                  // {S_m, R_m}
        p[s] = fma(-tv, fma(scale, x.x, x.y), p[s]);
      } else if (D >= 0 && s == D) {
        p[s] = top ? beta : v;
      }
    }
  }
  if (D >= -1 && D <= 6) *next_col = p[D + 1];                    // column j + 1 for everybody's next step
}

template <int QR_ROWS>
__global__ void __launch_bounds__(QR_ROWS * (B / QR_OCT), 1)
sbr_panel_qr_kernel(double* __restrict__ A, const long N, const int c0, double* __restrict__ Vall,
                    double* __restrict__ Gq, uint4* __restrict__ mail, const unsigned epoch,
                    unsigned* __restrict__ bar) {
  constexpr int QR_THREADS = QR_ROWS * (B / QR_OCT);
  constexpr int QR_WARPS = QR_THREADS / 32;
  constexpr int IF = QR_ROWS == 256 ? 2 : 4;   // mailbox reads in flight per thread (registers: 64 per thread at 1024)
  __shared__ double s_red[QR_WARPS][QR_OCT];
  __shared__ double2 s_SR[B];                  // {S_m, R_m}
  __shared__ double s_scal[4];
  __shared__ double s_part[QR_THREADS];        // thread (q, m): the partials of CTAs q, q + QR_WARPS, ... for slot m
  __shared__ double s_pj[2][QR_ROWS];          // column j of the panel, one entry per row of the CTA
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int oct = warp & 3, rl = (warp >> 2) * 32 + lane;      // column octet, row within the CTA
  const int G = gridDim.x, cta = blockIdx.x;
  const long r0 = c0 + B;
  const long n = N - r0;
  const long row = (long)cta * QR_ROWS + rl;
  const bool active = row < n;
  unsigned* err = bar + 1;
  const bool up4 = (lane & 4) != 0, up2 = (lane & 2) != 0, up1 = (lane & 1) != 0;

  double p[QR_OCT];
#pragma unroll
  for (int s = 0; s < QR_OCT; ++s) p[s] = active ? A[(long)(c0 + oct * QR_OCT + s) * N + r0 + row] : 0.0;
  if (oct == 0) s_pj[0][rl] = p[0];
  double vprev = 0.0;                            // v_{j-1}[row]
  __syncthreads();

#ifdef PODSDF_CHASE_PROFILE
  long long qc[4] = {0, 0, 0, 0};
  long long qlast = clock64();
#define QCLK(n) { const long long t_ = clock64(); qc[n] += t_ - qlast; qlast = t_; }
#else
#define QCLK(n)
#endif
#pragma unroll 1
  for (int j = 0; j <= B; ++j) {
    QCLK(3)
    const unsigned flag = epoch * 64u + (unsigned)(j + 1);
    uint4* box = mail + (size_t)(j & 1) * (G + 1) * B;
    const double pj = (j < B) ? s_pj[j & 1][rl] : 0.0;
    const int d = j - QR_OCT * oct;
    const int cls = d < -1 ? -2 : (d > 8 ? 9 : d);
    // ---- phase A: this thread's contributions.  slots m >= j: P[r][j] P[r][m] over rows > j (column j's norm and
    // dots); slots i < j-1: V[r][i] v_{j-1}[r] over rows >= j-1 (T factor of the column finished in the previous step)
    const double fa = (j < B && active && row > j) ? pj : 0.0;
    const double fb = vprev;
    if (j < B && row == j) {                       // the owners of row j post it (zeros when the panel is shorter)
#pragma unroll
      for (int s = 0; s < QR_OCT; ++s) ll_store(box + (size_t)G * B + oct * QR_OCT + s, p[s], flag);
    }
    {
      double t[QR_OCT];
      switch (cls) {
        case -2: qr_terms<-2>(p, fa, fb, t); break;
        case -1: qr_terms<-1>(p, fa, fb, t); break;
        case 0: qr_terms<0>(p, fa, fb, t); break;
        case 1: qr_terms<1>(p, fa, fb, t); break;
        case 2: qr_terms<2>(p, fa, fb, t); break;
        case 3: qr_terms<3>(p, fa, fb, t); break;
        case 4: qr_terms<4>(p, fa, fb, t); break;
        case 5: qr_terms<5>(p, fa, fb, t); break;
        case 6: qr_terms<6>(p, fa, fb, t); break;
        case 7: qr_terms<7>(p, fa, fb, t); break;
        case 8: qr_terms<8>(p, fa, fb, t); break;
        default: qr_terms<9>(p, fa, fb, t); break;
      }
      // 8 sums over the 32 lanes (rows) of a warp: three rounds of recursive halving (4 + 2 + 1 shuffles), after
      // which lane l holds slot (l & 7) summed over 8 lanes, then two butterfly rounds over the other lane bits
      double a4[4], a2[2], a1;
#pragma unroll
      for (int q = 0; q < 4; ++q) a4[q] = (up4 ? t[q + 4] : t[q]) + __shfl_xor_sync(0xffffffffu, up4 ? t[q] : t[q + 4], 4);
#pragma unroll
      for (int q = 0; q < 2; ++q) a2[q] = (up2 ? a4[q + 2] : a4[q]) + __shfl_xor_sync(0xffffffffu, up2 ? a4[q] : a4[q + 2], 2);
      a1 = (up1 ? a2[1] : a2[0]) + __shfl_xor_sync(0xffffffffu, up1 ? a2[0] : a2[1], 1);
      a1 += __shfl_xor_sync(0xffffffffu, a1, 8);
      a1 += __shfl_xor_sync(0xffffffffu, a1, 16);
      if (lane < QR_OCT) s_red[warp][lane] = a1;
    }
    __syncthreads();
    if (tid < B) {                                 // slot tid: octet tid / 8 -> warps (tid / 8), (tid / 8) + 4, ...
      double v = 0.0;
#pragma unroll
      for (int g = 0; g < QR_WARPS / 4; ++g) v += s_red[g * 4 + (tid >> 3)][tid & 7];
      ll_store(box + (size_t)cta * B + tid, v, flag);
    }
    QCLK(0)
    // ---- phase B: collect everybody's partial sums; thread (q, m) takes CTAs q, q + QR_WARPS, ... of slot m, IF in
    // flight at a time, and adds them in CTA order
    int fail = 0;
    double rj = 0.0;
    {
      double acc = 0.0;
      for (int g0 = warp; g0 < G; g0 += IF * QR_WARPS) {
        double val[IF];
        bool need[IF];
#pragma unroll
        for (int u = 0; u < IF; ++u) {
          val[u] = 0.0;
          need[u] = g0 + u * QR_WARPS < G;
        }
        const uint4* e0 = box + (size_t)g0 * B + lane;
        unsigned spins = 0;
        bool pending = true;
        while (pending) {
          pending = false;
#pragma unroll
          for (int u = 0; u < IF; ++u)
            if (need[u]) {
              need[u] = !ll_try(e0 + (size_t)u * QR_WARPS * B, flag, val[u]);
              pending |= need[u];
            }
          if (pending && ++spins > (1u << 16)) {
            if ((spins & 0xFFFu) == 0 && ld_acquire_u32(err) != 0) { fail = 1; break; }
            if (spins > (1u << 23)) { atomicExch(err, 2u); fail = 1; break; }
          }
        }
#pragma unroll
        for (int u = 0; u < IF; ++u) acc += val[u];
      }
      s_part[tid] = acc;
      if (tid < B && j < B && !fail) {
        unsigned spins = 0;
        while (!ll_try(box + (size_t)G * B + tid, flag, rj)) {
          if (++spins > (1u << 23)) { atomicExch(err, 2u); fail = 1; break; }
        }
      }
    }
    if (__syncthreads_or(fail)) return;
    QCLK(1)
    if (tid < B) {                                 // warp 0: totals (same order in every CTA), then the reflector's scalars
      double v0 = 0.0, v1 = 0.0, v2 = 0.0, v3 = 0.0;
#pragma unroll
      for (int q = 0; q < QR_WARPS; q += 4) {
        v0 += s_part[q * B + tid];
        v1 += s_part[(q + 1) * B + tid];
        v2 += s_part[(q + 2) * B + tid];
        v3 += s_part[(q + 3) * B + tid];
      }
      const double v = (v0 + v1) + (v2 + v3);
      s_SR[tid] = make_double2(v, rj);
      if (cta == 0 && j >= 1 && tid < j - 1) Gq[(j - 1) * B + tid] = v;   // g of the column finished in step j-1
      if (j < B) {
        const double alpha = __shfl_sync(0xffffffffu, rj, j), sigma = __shfl_sync(0xffffffffu, v, j);
        double beta = alpha, tau = 0.0, scale = 0.0;
        if (sigma > 0.0) {
          beta = -copysign(sqrt(fma(alpha, alpha, sigma)), alpha);
          tau = (beta - alpha) / beta;
          scale = 1.0 / (alpha - beta);
        }
        if (tid == 0) {
          s_scal[0] = beta;
          s_scal[1] = tau;
          s_scal[2] = scale;
          if (cta == 0) Gq[B * B + j] = tau;
        }
      }
    }
    __syncthreads();
    if (j < B) {
      const double beta = s_scal[0], tau = s_scal[1], scale = s_scal[2];
      const bool mine = active && row >= j, top = row == j;
      const double v = mine ? (top ? 1.0 : pj * scale) : 0.0;
      const double tv = tau * v;
      const double2* sr = s_SR + oct * QR_OCT;
      double* nc = &s_pj[(j + 1) & 1][rl];
      switch (cls) {
        case -2: qr_update<-2>(p, mine, top, v, tv, beta, scale, sr, nc); break;
        case -1: qr_update<-1>(p, mine, top, v, tv, beta, scale, sr, nc); break;
        case 0: qr_update<0>(p, mine, top, v, tv, beta, scale, sr, nc); break;
        case 1: qr_update<1>(p, mine, top, v, tv, beta, scale, sr, nc); break;
        case 2: qr_update<2>(p, mine, top, v, tv, beta, scale, sr, nc); break;
        case 3: qr_update<3>(p, mine, top, v, tv, beta, scale, sr, nc); break;
        case 4: qr_update<4>(p, mine, top, v, tv, beta, scale, sr, nc); break;
        case 5: qr_update<5>(p, mine, top, v, tv, beta, scale, sr, nc); break;
        case 6: qr_update<6>(p, mine, top, v, tv, beta, scale, sr, nc); break;
        case 7: qr_update<7>(p, mine, top, v, tv, beta, scale, sr, nc); break;
        case 8: qr_update<8>(p, mine, top, v, tv, beta, scale, sr, nc); break;
        default: break;                            // octet left of column j - 1: finished
      }
      vprev = v;
    }
    __syncthreads();
    QCLK(2)
  }
#ifdef PODSDF_CHASE_PROFILE
  if (tid == 0 && cta == 0 && (c0 == 0 || c0 == 4992))
    printf("qr c0 %d G %d: cycles per column: phaseA %lld, wait %lld, phaseB %lld, loop %lld\n", c0, G, qc[0] / (B + 1),
           qc[1] / (B + 1), qc[2] / (B + 1), qc[3] / (B + 1));
#endif
  // ---- write back: R (rows < B, upper triangle incl. diagonal) into A; V (unit lower trapezoid) into Vall
  if (active) {
#pragma unroll
    for (int s = 0; s < QR_OCT; ++s) {
      const int m = oct * QR_OCT + s;
      if (row <= m) A[(long)(c0 + m) * N + r0 + row] = p[s];
      Vall[(long)(c0 + m) * N + r0 + row] = (row > m) ? p[s] : (row == m ? 1.0 : 0.0);
    }
  }
}

// T factor of the panel's block reflector (dlarft, forward / columnwise) from Gq = [g (B x B) | tau (B)]:
//   T[j][j] = tau_j,  T[:j, j] = -tau_j T[:j, :j] g_j,  g_j[l] = V[:, l]^T v_j (l < j).
// Row i of T only depends on row i: one lane per row, no synchronisation.  Runs behind the QR on its stream, beside
// the symm kernel of the same panel (T is first needed by sbr_s_kernel).
__global__ void __launch_bounds__(B)
sbr_tfactor_kernel(const double* __restrict__ Gq, double* __restrict__ Tout) {
  __shared__ double sT[B][B + 1], sG[B][B + 1], sTau[B];
  const int i = threadIdx.x;
  for (int m = 0; m < B; ++m) {
    sT[i][m] = 0.0;
    sG[m][i] = Gq[m * B + i];                    // coalesced; sG[j][l] = g_j[l]
  }
  sTau[i] = Gq[B * B + i];
  __syncwarp();
  sT[i][i] = sTau[i];
  for (int j = i + 1; j < B; ++j) {
    double acc = 0.0;
    for (int l = i; l < j; ++l) acc = fma(sT[i][l], sG[j][l], acc);
    sT[i][j] = -sTau[j] * acc;
  }
  __syncwarp();
  for (int m = 0; m < B; ++m) Tout[m * B + i] = sT[m][i];
}

// ------------------------------------------------------------------------------------------------
// Stage 1b: Xp[split][m][r] = sum over this split's t of A22[r][t] V[t][m]   (rows r0 .. N-1, t in r0 .. N-1)
//   CTA = 128 rows x 32 columns, 8 warps (16 rows x 32 columns each), K in chunks of 16, 3-stage cp.async.
//   Only the lower triangle of A22 is valid (plus full diagonal tiles): a chunk whose t-block lies left of / on the
//   row block is read as A[t*N + r] ("t-major": tile [16 t][128 + 4]), one to the right of it as A[r*N + t]
//   ("row-major": tile [128 r][16 + 4]) -- the two smem layouts of the projection kernel, chosen per chunk.
// ------------------------------------------------------------------------------------------------
constexpr int XR = 128, XKB = 16, XST = 3;
constexpr int XPA0 = XKB + 4, XPA1 = XR + 4, XPV = XKB + 4;
constexpr int XATILE = (XR * XPA0 > XKB * XPA1) ? XR * XPA0 : XKB * XPA1;
constexpr int XVTILE = B * XPV;
constexpr size_t XSMEM = (size_t)XST * (XATILE + XVTILE) * sizeof(double);

__global__ void __launch_bounds__(256, 2)
sbr_symm_kernel(const double* __restrict__ A, const long N, const long r0, const double* __restrict__ V,
                double* __restrict__ Xp, const int chunks_per_split) {
  extern __shared__ double smem[];
  double* sA = smem;
  double* sV = sA + XST * XATILE;
  // tiles are anchored at absolute multiples of 128 (not at r0), so "diagonal tile" means the same elements for
  // every panel; rows / columns below r0 are masked
  const long tile = r0 / XR + blockIdx.x;                    // absolute row-tile index
  const long row0 = tile * XR;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int lr = lane & 3, lc = lane >> 2;
  const int wr = warp * 16;
  const int chunk_lo = (int)(r0 / XKB), chunk_hi = (int)((N + XKB - 1) / XKB);   // absolute 16-column chunks
  const int c_begin = chunk_lo + blockIdx.y * chunks_per_split;
  const int c_end = min(chunk_hi, c_begin + chunks_per_split);
  const int nchunks = max(0, c_end - c_begin);

  auto tmajor = [&](int chunk) { return ((long)chunk * XKB) / XR <= tile; };   // t-block <= row block
  // even N: every row starts 16-byte aligned -> two entries per copy (the copies' address arithmetic and predicates
  // compete with the DMMAs for issue slots: 8-byte copies cost 15 % of the kernel; bulk copies of 128-byte rows were
  // 2.5x slower than either, profiles/r02_sbr_profile.md)
  const bool vec = (N & 1) == 0;
  auto load_stage = [&](int chunk, int stage) {
    const long t0 = (long)chunk * XKB;
    double* a = sA + stage * XATILE;
    double* vv = sV + stage * XVTILE;
    if (vec) {
      if (tmajor(chunk)) {                                      // A[t*N + row]: rows contiguous
        for (int e = tid; e < XKB * XR / 2; e += 256) {
          const int t = e / (XR / 2), r = 2 * (e % (XR / 2));
          cpa16(a + t * XPA1 + r, A + (t0 + t) * N + row0 + r, (row0 + r >= r0) && (row0 + r < N) && (t0 + t < N));
        }
      } else {                                                  // A[row*N + t]: t contiguous
        for (int e = tid; e < XR * XKB / 2; e += 256) {
          const int r = e / (XKB / 2), t = 2 * (e % (XKB / 2));
          cpa16(a + r * XPA0 + t, A + (row0 + r) * N + t0 + t, (row0 + r >= r0) && (row0 + r < N) && (t0 + t < N));
        }
      }
      {                                                         // V[m][t] -> sV[m][t]: B * XKB / 2 = 256 copies
        const int m = tid / (XKB / 2), t = 2 * (tid % (XKB / 2));
        cpa16(vv + m * XPV + t, V + (long)m * N + t0 + t, t0 + t < N);
      }
      return;
    }
    if (tmajor(chunk)) {                                        // A[t*N + row]: rows contiguous
      for (int e = tid; e < XKB * XR; e += 256) {
        const int t = e / XR, r = e % XR;
        cpa8(a + t * XPA1 + r, A + (t0 + t) * N + row0 + r, (row0 + r >= r0) && (row0 + r < N) && (t0 + t < N));
      }
    } else {                                                    // A[row*N + t]: t contiguous
      for (int e = tid; e < XR * XKB; e += 256) {
        const int r = e / XKB, t = e % XKB;
        cpa8(a + r * XPA0 + t, A + (row0 + r) * N + t0 + t, (row0 + r >= r0) && (row0 + r < N) && (t0 + t < N));
      }
    }
    for (int e = tid; e < B * XKB; e += 256) {                  // V[m][t] -> sV[m][t]
      const int m = e / XKB, t = e % XKB;
      cpa8(vv + m * XPV + t, V + (long)m * N + t0 + t, t0 + t < N);
    }
  };

  double acc[2][4][2];
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;

  for (int s = 0; s < XST - 1; ++s) {
    if (s < nchunks) load_stage(c_begin + s, s);
    cpa_commit();
  }
  for (int c = 0; c < nchunks; ++c) {
    cpa_wait<XST - 2>();
    __syncthreads();
    if (c + XST - 1 < nchunks) load_stage(c_begin + c + XST - 1, (c + XST - 1) % XST);
    cpa_commit();
    const double* a = sA + (c % XST) * XATILE;
    const double* vv = sV + (c % XST) * XVTILE;
    const bool tm = tmajor(c_begin + c);
#pragma unroll
    for (int k4 = 0; k4 < XKB; k4 += 4) {
      double af[2], bf[4];
      const int t = k4 + lr;
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) af[mt] = tm ? a[t * XPA1 + wr + mt * 8 + lc] : a[(wr + mt * 8 + lc) * XPA0 + t];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) bf[nt] = vv[(nt * 8 + lc) * XPV + t];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) dmma884s(acc[mt][nt][0], acc[mt][nt][1], af[mt], bf[nt]);
    }
  }
  cpa_wait<0>();
  double* Xs = Xp + (long)blockIdx.y * B * N;
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      const long row = row0 + wr + mt * 8 + lc;
      const int m = nt * 8 + 2 * lr;
      if (row >= r0 && row < N) {
        Xs[(long)m * N + row] = acc[mt][nt][0];
        Xs[(long)(m + 1) * N + row] = acc[mt][nt][1];
      }
    }
}

// ------------------------------------------------------------------------------------------------
// Stage 1c: X = sum of the split partials; Zp[cta] = V_chunk^T X_chunk (B x B partial of V^T X)
// ------------------------------------------------------------------------------------------------
constexpr int WR = 64;    // rows per CTA of sbr_w1_kernel
constexpr int WR2 = 32;   // rows per CTA of sbr_w2_kernel

__global__ void __launch_bounds__(256)
sbr_w1_kernel(const double* __restrict__ Xp, const int nsplit, const double* __restrict__ V, const long N,
              const long r0, double* __restrict__ X, double* __restrict__ Zp) {
  __shared__ double sV[B][WR + 1], sX[B][WR + 1];
  const long n = N - r0;
  const long row0 = (long)blockIdx.x * WR;
  const int tid = threadIdx.x;
  for (int e = tid; e < B * WR; e += 256) {
    const int m = e / WR, r = e % WR;
    const bool ok = row0 + r < n;
    double x = 0.0;
    if (ok)
      for (int s = 0; s < nsplit; ++s) x += Xp[((long)s * B + m) * N + r0 + row0 + r];
    sX[m][r] = x;
    sV[m][r] = ok ? V[(long)m * N + r0 + row0 + r] : 0.0;
    if (ok) X[(long)m * N + r0 + row0 + r] = x;
  }
  __syncthreads();
  for (int e = tid; e < B * B; e += 256) {
    const int m1 = e / B, m2 = e % B;
    double acc = 0.0;
#pragma unroll 8
    for (int r = 0; r < WR; ++r) acc = fma(sV[m1][r], sX[m2][r], acc);
    Zp[(long)blockIdx.x * B * B + e] = acc;
  }
}

// Stage 1d: Z = sum Zp;  S = T^T Z T  (one CTA: 1024 threads, one entry each; the partials are read coalesced)
__global__ void __launch_bounds__(B * B)
sbr_s_kernel(const double* __restrict__ T, const double* __restrict__ Zp, const int nz, double* __restrict__ S) {
  __shared__ double sT[B][B + 1], sZ[B][B + 1], sU[B][B + 1];
  const int e = threadIdx.x, a = e / B, c = e % B;
  double z0 = 0.0, z1 = 0.0, z2 = 0.0, z3 = 0.0;       // four independent streams of loads, fixed order
  int g2 = 0;
  for (; g2 + 3 < nz; g2 += 4) {
    z0 += Zp[(long)g2 * B * B + e];
    z1 += Zp[(long)(g2 + 1) * B * B + e];
    z2 += Zp[(long)(g2 + 2) * B * B + e];
    z3 += Zp[(long)(g2 + 3) * B * B + e];
  }
  for (; g2 < nz; ++g2) z0 += Zp[(long)g2 * B * B + e];
  sZ[a][c] = (z0 + z1) + (z2 + z3);
  sT[a][c] = T[e];
  __syncthreads();
  double acc = 0.0;                                    // U = Z T (T upper triangular)
  for (int l = 0; l <= c; ++l) acc = fma(sZ[a][l], sT[l][c], acc);
  sU[a][c] = acc;
  __syncthreads();
  acc = 0.0;                                           // S = T^T U
  for (int l = 0; l <= a; ++l) acc = fma(sT[l][a], sU[l][c], acc);
  S[e] = acc;
}

// Stage 1d': W = X T - 1/2 V S
__global__ void __launch_bounds__(256)
sbr_w2_kernel(const double* __restrict__ X, const double* __restrict__ V, const double* __restrict__ T,
              const double* __restrict__ S, const long N, const long r0, double* __restrict__ W) {
  __shared__ double sT[B][B + 1], sS[B][B + 1];
  __shared__ double sV[B][WR2 + 1], sX[B][WR2 + 1];
  const long n = N - r0;
  const long row0 = (long)blockIdx.x * WR2;
  const int tid = threadIdx.x;
  for (int e = tid; e < B * B; e += 256) {
    sS[e / B][e % B] = S[e];
    sT[e / B][e % B] = T[e];
  }
  for (int e = tid; e < B * WR2; e += 256) {
    const int m = e / WR2, r = e % WR2;
    const bool ok = row0 + r < n;
    sX[m][r] = ok ? X[(long)m * N + r0 + row0 + r] : 0.0;
    sV[m][r] = ok ? V[(long)m * N + r0 + row0 + r] : 0.0;
  }
  __syncthreads();
  for (int e = tid; e < B * WR2; e += 256) {
    const int m = e / WR2, r = e % WR2;
    if (row0 + r < n) {
      double acc = 0.0;
      for (int l = 0; l <= m; ++l) acc = fma(sX[l][r], sT[l][m], acc);
      double acc2 = 0.0;
#pragma unroll 8
      for (int l = 0; l < B; ++l) acc2 = fma(sV[l][r], sS[l][m], acc2);
      W[(long)m * N + r0 + row0 + r] = fma(-0.5, acc2, acc);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Stage 1e: A22 -= V W^T + W V^T on the tiles (I >= J) of the trailing matrix (diagonal 128-blocks in full).
//   CTA = 128 (i) x 64 (j) tile, 8 warps x (32 x 32), K = 2B resident in shared memory ([k][pitch], the gram kernel's
//   micro-kernel); two CTAs per SM, so that one CTA's loads (old tile -> accumulators, operands -> shared memory) and
//   stores overlap the other's DMMAs -- at K = 64 they are half of a tile's time (profiles/r02_sbr_profile.md).  Eight
//   warps, not four: a lone warp per scheduler drives the FP64 tensor pipe at ~45 % (its DMMAs wait on each other's issue),
//   two reach what the Gram kernel reaches -- so each CTA brings its own two.
//   operand a (rows i): k < B -> -V[k][i], else -W[k-B][i];  operand b (rows j): k < B -> W[k][j], else V[k-B][j].
//   The accumulators start as the old tile, so the epilogue is a plain store.
//   Tiles are anchored at absolute multiples of 128 rows / 64 columns (see sbr_symm_kernel); with h = 1 when r0 lies in
//   the upper half of its 128-block, row r (from the block of r0) holds the column blocks 0 .. 2r + 1 - h.
// ------------------------------------------------------------------------------------------------
constexpr int UT = 128, UTJ = 64, UPA = UT + 4, UPB = UTJ + 4;
constexpr int U_THREADS = 256;
constexpr size_t USMEM = (size_t)(2 + 2 * B * (UPA + UPB)) * sizeof(double);    // mbarrier + operands

// number of CTAs of one update: part 0 = every tile, 1 = the first column block only (it holds the next panel),
// 2 = the other column blocks -- parts 1 + 2 let the next panel's QR overlap most of the update
static inline long syr2k_ctas(long N, long r0, int part) {
  const long side = (N + UT - 1) / UT - r0 / UT;
  const long h = (r0 % UT) / UTJ;
  if (part == 1) return side;
  if (part == 2) return side * side - h * side;
  return side * (side + 1) - h * side;
}

// At K = 64 a tile's loads and stores are as long as its DMMAs, and with four warps per CTA they were bound by the
// ISSUE of their address arithmetic and predicates (prologue = 49 % of the warp samples, profiles/r02_sbr_profile.md):
//   * even N: the 2 x 64 operand rows arrive by 128 bulk copies issued by one warp.  They are not masked: entries
//     outside [r0, N) are other rows' or stale values, which only reach accumulator rows / columns that are not stored
//     (reads stay inside the workspace: the skinny matrices are followed by other buffers);
//   * tiles entirely inside [r0, N) take a path without predicates, addresses = 8 column pointers + immediates.
// -DPODSDF_SYR2K_RED=1 (A/B): accumulators start at zero and the epilogue adds them at the L2 -- measured equal
// (101.5 vs 102.4 ms for stage 1), so the plain read / store form stays
#ifndef PODSDF_SYR2K_RED
#define PODSDF_SYR2K_RED 0
#endif
#if PODSDF_SYR2K_RED
#define U_OUT(ptr, val) atomicAdd((ptr), (val))     // result unused -> RED.ADD.F64 at the L2: no read of the old tile
#else
#define U_OUT(ptr, val) (*(ptr) = (val))
#endif
__global__ void __launch_bounds__(U_THREADS, 2)
sbr_syr2k_kernel(double* __restrict__ A, const long N, const long r0, const double* __restrict__ V,
                 const double* __restrict__ W, const int part) {
  extern __shared__ double smem[];
  double* sA = smem + 2;                // [2B][UPA]  (smem[0]: the mbarrier)
  double* sB = sA + 2 * B * UPA;        // [2B][UPB]
  const unsigned mbar = (unsigned)__cvta_generic_to_shared(smem);
  const long tile0 = r0 / UT;
  const int h = (int)((r0 % UT) / UTJ);
  int r, rem;                           // linear id -> (row r, column block rem); row q holds 2 q + c blocks
  if (part == 1) {
    r = blockIdx.x;
    rem = 0;
  } else {
    const int c1 = (part == 2 ? 0 : 1) - h, id = blockIdx.x;     // blocks before row q: q^2 + c1 q
    r = (int)((sqrtf((float)(c1 * c1 + 4 * id)) - (float)c1) * 0.5f);
    while (r > 0 && r * r + c1 * r > id) --r;
    while ((r + 1) * (r + 1) + c1 * (r + 1) <= id) ++r;
    rem = id - (r * r + c1 * r) + (part == 2 ? 1 : 0);
  }
  const long i0 = (tile0 + r) * UT, j0 = (2 * tile0 + h + rem) * UTJ;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = (warp >> 1) * 32, wn = (warp & 1) * 32;     // 8 warps: 4 row groups of 32 x 2 column groups of 32
  const int lr = lane & 3, lc = lane >> 2;
  const bool bulk = (N & 1) == 0;       // operand rows start 16-byte aligned

  if (bulk) {
    if (tid == 0) mbar_init(mbar, 1);
    __syncthreads();
    if (warp == 0) {
      if (lane == 0) mbar_expect_tx(mbar, 2 * B * (UT + UTJ) * (unsigned)sizeof(double));
      __syncwarp();
      for (int k = lane; k < 2 * B; k += 32) {
        const double* pa = (k < B) ? V + (long)k * N : W + (long)(k - B) * N;
        const double* pb = (k < B) ? W + (long)k * N : V + (long)(k - B) * N;
        bulk_g2s((unsigned)__cvta_generic_to_shared(sA + k * UPA), pa + i0, UT * sizeof(double), mbar);
        bulk_g2s((unsigned)__cvta_generic_to_shared(sB + k * UPB), pb + j0, UTJ * sizeof(double), mbar);
      }
    }
  } else {
    for (int e = tid; e < 2 * B * UT; e += U_THREADS) {
      const int k = e / UT, x = e % UT;
      const double* pa = (k < B) ? V + (long)k * N : W + (long)(k - B) * N;
      cpa8(sA + k * UPA + x, pa + i0 + x, (i0 + x >= r0) && (i0 + x < N));
    }
    for (int e = tid; e < 2 * B * UTJ; e += U_THREADS) {
      const int k = e / UTJ, x = e % UTJ;
      const double* pb = (k < B) ? W + (long)k * N : V + (long)(k - B) * N;
      cpa8(sB + k * UPB + x, pb + j0 + x, (j0 + x >= r0) && (j0 + x < N));
    }
    cpa_commit();
  }

  const bool full = i0 >= r0 && i0 + UT <= N && j0 >= r0 && j0 + UTJ <= N;
  double* const c00 = A + (j0 + wn + 2 * lr) * N + i0 + wm + lc;     // element (mt, nt, e) at c00[(nt * 8 + e) * N + mt * 8]
  double acc[4][4][2];
#if PODSDF_SYR2K_RED
#pragma unroll
  for (int mt = 0; mt < 4; ++mt)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = 0.0;
#else
  if (full) {
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const double* col = c00 + (long)(nt * 8 + e) * N;
#pragma unroll
        for (int mt = 0; mt < 4; ++mt) acc[mt][nt][e] = col[mt * 8];
      }
  } else {
#pragma unroll
    for (int mt = 0; mt < 4; ++mt)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const long i = i0 + wm + mt * 8 + lc, j = j0 + wn + nt * 8 + 2 * lr;
        const bool iok = i >= r0 && i < N;
        acc[mt][nt][0] = (iok && j >= r0 && j < N) ? A[j * N + i] : 0.0;
        acc[mt][nt][1] = (iok && j + 1 >= r0 && j + 1 < N) ? A[(j + 1) * N + i] : 0.0;
      }
  }
#endif
  if (bulk) {
    mbar_wait(mbar, 0);
  } else {
    cpa_wait<0>();
    __syncthreads();
  }
#pragma unroll 4
  for (int k4 = 0; k4 < 2 * B; k4 += 4) {
    double af[4], bf[4];
    const int rr = k4 + lr;
#pragma unroll
    for (int mt = 0; mt < 4; ++mt) af[mt] = -sA[rr * UPA + wm + mt * 8 + lc];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) bf[nt] = sB[rr * UPB + wn + nt * 8 + lc];
#pragma unroll
    for (int mt = 0; mt < 4; ++mt)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) dmma884s(acc[mt][nt][0], acc[mt][nt][1], af[mt], bf[nt]);
  }
  if (full) {
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        double* col = c00 + (long)(nt * 8 + e) * N;
#pragma unroll
        for (int mt = 0; mt < 4; ++mt) U_OUT(col + mt * 8, acc[mt][nt][e]);
      }
  } else {
#pragma unroll
    for (int mt = 0; mt < 4; ++mt)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const long i = i0 + wm + mt * 8 + lc, j = j0 + wn + nt * 8 + 2 * lr;
        const bool iok = i >= r0 && i < N;
        if (iok && j >= r0 && j < N) U_OUT(A + j * N + i, acc[mt][nt][0]);
        if (iok && j + 1 >= r0 && j + 1 < N) U_OUT(A + (j + 1) * N + i, acc[mt][nt][1]);
      }
  }
}

// Stage 1f: lower band storage Bb[c][d] = A[c][c + d] (d <= B), zero where the bulge will live (B < d < LDB)
__global__ void __launch_bounds__(256)
sbr_extract_band_kernel(const double* __restrict__ A, const long N, double* __restrict__ Bb) {
  const long e = blockIdx.x * 256L + threadIdx.x;
  if (e >= N * LDB) return;
  const long c = e / LDB;
  const int d = (int)(e % LDB);
  Bb[e] = (d <= B && c + d < N) ? A[c * N + c + d] : 0.0;
}


// ------------------------------------------------------------------------------------------------
// Stage 2: band -> tridiagonal by bulge chasing (tools/proto_sbr.py: task()).  Band storage Bb[c*LDB + d] = B[c+d][c].
//   Sweep i annihilates column i below the sub-diagonal and chases the bulge down the band: task k works on the
//   diagonal block D = B[J_k, J_k] (two-sided reflector H_k) and on E = B[J_{k+1}, J_k] (E <- E H_k, new reflector
//   H_{k+1} from its first column, E <- H_{k+1} E), J_k = [i+1+k b, i+(k+1) b].
//   One warp per sweep (lane = block row, blocks staged in shared memory); chain c runs sweeps c, c+G, ...;
//   task (i, k) starts when sweep i-1 has finished task k+1: prog[i-1] >= k+2 (progress counters in global memory,
//   release / acquire).  Reflectors go to V2[i][row] / TAU2[i][k] for the back-transformation.
// ------------------------------------------------------------------------------------------------

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// LAPACK dlarfg on a warp: lane l holds x[l] (l < len, zero beyond); returns v (v[0] = 1), tau, beta
__device__ __forceinline__ void warp_house(const double x, const int lane, double& v, double& tau, double& beta) {
  const double alpha = __shfl_sync(0xffffffffu, x, 0);
  const double sigma = warp_sum(lane >= 1 ? x * x : 0.0);
  if (sigma == 0.0) {
    tau = 0.0;
    beta = alpha;
    v = (lane == 0) ? 1.0 : 0.0;
  } else {
    beta = -copysign(sqrt(fma(alpha, alpha, sigma)), alpha);
    tau = (beta - alpha) / beta;
    const double scale = 1.0 / (alpha - beta);
    v = (lane == 0) ? 1.0 : x * scale;
  }
}

__device__ __forceinline__ void cpa16_cg(void* smem_dst, const void* gsrc, bool valid) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  const int bytes = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(d), "l"(gsrc), "r"(bytes) : "memory");
}

// The blocks of a task are worked on in the band layout itself: sB[c][d] = B[j0 + c + d][j0 + c] is a copy of
// band columns j0 .. j0+L-1 (64 doubles each, fetched with ONE batch of 16-byte L2 copies: a task costs one
// memory round trip in, one out).  D(r, c) = sB[c][r - c] (r >= c, mirrored otherwise), E(a, c) = sB[c][B - c + a].
// ------------------------------------------------------------------------------------------------
// Stage 2, four warps per task (the first versions ran a task on one warp -- IPC ~0.4 on ~2000 instructions, 139 ms --
// and on two, D block / E chain, 105 ms; both are in the history of this file).  With two warps the task's own dependent
// arithmetic (~2.4 us) is what the 2 N-long chain is made of (hand-offs through shared memory or DSMEM instead of L2
// bought 6 %), so the blocks are split once more by columns: warps 0 / 1 take columns 0-15 / 16-31 of the E-block chain, warps 2 / 3
// those of the D block; partial matrix-vector products meet in shared memory behind 64-thread named barriers.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void pair_barrier(const int id) {
  asm volatile("bar.sync %0, 64;\n" ::"r"(id) : "memory");
}

__global__ void __launch_bounds__(128)
sbr_chase4_kernel(double* __restrict__ Bb, const long N, double* __restrict__ V2, double* __restrict__ TAU2,
                  const long ntask_max, unsigned* __restrict__ bar) {
  __shared__ __align__(16) double sB[B][LDB];
  __shared__ double sv[B], sv2[B], sz[B], sqD[2][B], sY[2][B], sP[2][B];
  __shared__ double s_tau, s_tau2, s_beta2;
  __shared__ int s_ok;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int half = warp & 1;                         // which 16 columns of its block this warp owns
  const int cb = half * (B / 2);
  unsigned* err = bar + 1;
  unsigned* prog = bar + 2;
  const long nsweep = N - 2;
  for (long i = blockIdx.x; i < nsweep; i += gridDim.x) {
    const int nt = (int)((N - 1 - i + B - 1) / B);
    double v = 0.0, tau = 0.0;                       // warp 0: reflector of the current task (lane = row inside J_k)
    for (int k = 0; k < nt; ++k) {
      if (threadIdx.x == 0) {                        // dependency: sweep i-1 has finished task k+1
        int ok = 1;
        if (i > 0) {
          const unsigned need = (unsigned)(k + 2);
          unsigned spins = 0;
          while (ld_acquire_u32(prog + i - 1) < need) {
            if (++spins > (1u << 22)) {
              if ((spins & 0xFFFFu) == 0 && ld_acquire_u32(err) != 0) { ok = 0; break; }
              if (spins > (1u << 27)) { atomicExch(err, 2u); ok = 0; break; }
            }
          }
        }
        s_ok = ok;
      }
      __syncthreads();
      if (!s_ok) return;
      const long j0 = i + 1 + (long)k * B;
      const int L = (int)min((long)B, N - j0);
      const int L2 = (int)max(0L, min((long)B, N - j0 - B));
#pragma unroll
      for (int cc = 0; cc < B / 4; ++cc) {           // band columns j0 .. j0+L-1 (zero-filled beyond), a quarter per warp
        const int c = warp * (B / 4) + cc;
        cpa16_cg(&sB[c][2 * lane], Bb + (j0 + c) * LDB + 2 * lane, c < L);
      }
      cpa_commit();
      if (warp == 0) {
        if (k == 0) {                                // reflector of column i (overlaps the copies)
          const double x = (lane < L) ? __ldcg(Bb + i * LDB + 1 + lane) : 0.0;
          double beta;
          warp_house(x, lane, v, tau, beta);
          if (lane < L) __stcg(Bb + i * LDB + 1 + lane, lane == 0 ? beta : 0.0);
        }
        if (lane >= L) v = 0.0;
        if (lane < L) V2[i * N + j0 + lane] = v;
        if (lane == 0) {
          TAU2[i * ntask_max + k] = tau;
          s_tau = tau;
        }
        sv[lane] = v;
      }
      cpa_wait<0>();
      __syncthreads();
      const double tk = s_tau, vl = sv[lane];
      if (warp >= 2) {
        // ---- D <- H D H: p = tau D v (each warp sums its 16 columns), q = p - (tau/2)(p^T v) v, D -= v q^T + q v^T
        double pp[2] = {0.0, 0.0};
#pragma unroll
        for (int cc = 0; cc < B / 2; ++cc) {
          const int c = cb + cc;
          const int idx = (c <= lane) ? c * LDB + lane - c : lane * LDB + c - lane;
          pp[cc & 1] = fma((&sB[0][0])[idx], sv[c], pp[cc & 1]);
        }
        sP[half][lane] = pp[0] + pp[1];
        pair_barrier(2);                             // both halves summed; nobody reads D any more
        const double pl = tk * (sP[0][lane] + sP[1][lane]);
        const double kk = 0.5 * tk * warp_sum(pl * vl);
        const double ql = fma(-kk, vl, pl);
        sqD[half][lane] = ql;                        // each warp keeps its own copy: no second barrier
        __syncwarp();
#pragma unroll
        for (int cc = 0; cc < B / 2; ++cc) {
          const int c = cb + cc;
          if (c <= lane) sB[c][lane - c] -= fma(vl, sqD[half][c], ql * sv[c]);
        }
        __syncwarp();
#pragma unroll
        for (int cc = 0; cc < B / 2; ++cc) {         // write back column c, d in [0, L - c)
          const int c = cb + cc;
          const int len = L - c, d0 = 2 * lane;
          double* dst = Bb + (j0 + c) * LDB + d0;
          if (d0 + 1 < len) __stcg(reinterpret_cast<double2*>(dst), make_double2(sB[c][d0], sB[c][d0 + 1]));
          else if (d0 < len) __stcg(dst, sB[c][d0]);
        }
      } else {
        double v2 = 0.0, tau2 = 0.0;
        if (L2 > 0) {
          // ---- E <- E H (lane = row a of E; each warp its 16 columns)
          double yy[2] = {0.0, 0.0};
#pragma unroll
          for (int cc = 0; cc < B / 2; ++cc) {
            const int c = cb + cc;
            yy[cc & 1] = fma(sB[c][B - c + lane], sv[c], yy[cc & 1]);
          }
          sY[half][lane] = yy[0] + yy[1];
          pair_barrier(1);
          const double yl = tk * (sY[0][lane] + sY[1][lane]);
#pragma unroll
          for (int cc = 0; cc < B / 2; ++cc) {
            const int c = cb + cc;
            sB[c][B - c + lane] = fma(-yl, sv[c], sB[c][B - c + lane]);
          }
          __syncwarp();
          if (warp == 0) {                           // new reflector from the first column of E
            double beta2;
            warp_house(lane < L2 ? sB[0][B + lane] : 0.0, lane, v2, tau2, beta2);
            if (lane >= L2) v2 = 0.0;
            sv2[lane] = v2;
            if (lane == 0) {
              s_tau2 = tau2;
              s_beta2 = beta2;
            }
          }
          pair_barrier(1);
          const double t2 = s_tau2, v2l = sv2[lane];
          // ---- E <- H2 E: z_c = tau2 v2^T E[:, c]; lanes 0-15 / 16-31 sum rows 0-15 / 16-31 of column cb + (lane & 15)
          {
            const int c = cb + (lane & 15), a0 = (lane >> 4) * 16;
            const double* col = &sB[c][B - c + a0];
            double zz[2] = {0.0, 0.0};
#pragma unroll
            for (int a2 = 0; a2 < 16; ++a2) zz[a2 & 1] = fma(sv2[a0 + a2], col[a2], zz[a2 & 1]);
            double zl = zz[0] + zz[1];
            zl += __shfl_xor_sync(0xffffffffu, zl, 16);
            if (lane < 16) sz[c] = t2 * zl;
          }
          __syncwarp();
#pragma unroll
          for (int cc = 0; cc < B / 2; ++cc) {
            const int c = cb + cc;
            sB[c][B - c + lane] = fma(-v2l, sz[c], sB[c][B - c + lane]);
          }
          if (warp == 0) sB[0][B + lane] = (lane == 0) ? s_beta2 : 0.0;     // exactly (beta, 0, ..., 0)
          __syncwarp();
#pragma unroll
          for (int cc = 0; cc < B / 2; ++cc) {       // write back column c, d = B - c + a
            const int c = cb + cc;
            if (c < L && lane < L2) __stcg(Bb + (j0 + c) * LDB + B - c + lane, sB[c][B - c + lane]);
          }
        }
        v = v2;
        tau = tau2;
      }
      // ---- publish progress: the CTA barrier orders every warp's stores before thread 0's release store
      __syncthreads();
      if (threadIdx.x == 0) st_release_u32(prog + i, (k + 1 == nt) ? 0x7fffffffu : (unsigned)(k + 1));
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Stage 2, four warps per task, the critical hand-off through a mailbox.  Of the 32 band columns of task (i, k), 31
// were written by task (i-1, k) -- a whole task time before they are needed -- and ONE, the last, by task (i-1, k+1),
// whose completion is what task (i, k) waits for: it is that task's column 0, final as soon as its D block is updated
// and the new reflector is known (H2 leaves (beta, 0, ..., 0) below).  So a task posts its column 0 in flagged 16-byte
// entries {lo, flag, hi, flag} (flag = sweep + 1) as soon as it is final, mid-task, and never writes it to the band;
// the next sweep polls those 64 entries: data and "it is there" in one L2 round trip, instead of "finish the task,
// store, release, poll, load".  Slots are per (sweep parity, task): sweep i+2 reaches task k only after sweep i+1 has
// finished task k+1, i.e. long after it read slot (i, k).  The 31 other columns and the progress counters go through
// the band as before (waited for with k+1 instead of k+2: never the critical edge).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
sbr_chase5_kernel(double* __restrict__ Bb, const long N, double* __restrict__ V2, double* __restrict__ TAU2,
                  const long ntask_max, unsigned* __restrict__ bar, uint4* __restrict__ mail) {
  __shared__ __align__(16) double sBB[2][B][LDB];   // task k works in sBB[k & 1]; the other half receives task k+1's columns
  __shared__ double sv[B], sv2[B], sz[B], sqD[2][B], sY[2][B], sP[2][B];
  __shared__ double s_tau, s_tau2, s_beta2;
  __shared__ int s_ok, s_pre[2];                    // s_pre[k & 1]: the prefetch of task k's columns succeeded
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int half = warp & 1;                         // which 16 columns of its block this warp owns
  const int cb = half * (B / 2);
  unsigned* err = bar + 1;
  unsigned* prog = bar + 2;
  const long nsweep = N - 2;
  for (long i = blockIdx.x; i < nsweep; i += gridDim.x) {
    const int nt = (int)((N - 1 - i + B - 1) / B);
    const unsigned myflag = (unsigned)(i + 1), inflag = (unsigned)i;       // mailbox flags: sweep index + 1
    uint4* outbox = mail + (size_t)(i & 1) * ntask_max * LDB;              // slot of task k: outbox + k * LDB
    const uint4* inbox = mail + (size_t)((i + 1) & 1) * ntask_max * LDB;   // the previous sweep's
    double v = 0.0, tau = 0.0;                       // warp 0: reflector of the current task (lane = row inside J_k)
    for (int k = 0; k < nt; ++k) {
      double (*sB)[LDB] = sBB[k & 1];
      const long j0 = i + 1 + (long)k * B;
      const int L = (int)min((long)B, N - j0);
      const int L2 = (int)max(0L, min((long)B, N - j0 - B));
      const bool boxed = i > 0 && L == B;            // column 31 arrives in the mailbox of task (i-1, k+1)
      const bool keep0 = k == 0;                     // column 0 of task 0 has no mailbox reader: it goes to the band
      if (k == 0) {
        // first task of the sweep: columns 0 .. 30 come from task 0 of sweep i-1, through the band (later tasks find
        // theirs prefetched by the D warps of the task before)
        if (threadIdx.x == 0) {
          int ok = 1;
          if (i > 0) {
            unsigned spins = 0;
            while (ld_acquire_u32(prog + i - 1) < 1u) {
              if (++spins > (1u << 22)) {
                if ((spins & 0xFFFFu) == 0 && ld_acquire_u32(err) != 0) { ok = 0; break; }
                if (spins > (1u << 27)) { atomicExch(err, 2u); ok = 0; break; }
              }
            }
          }
          s_ok = ok;
          s_pre[0] = 1;
        }
        __syncthreads();
        if (!s_ok) return;
#pragma unroll
        for (int cc = 0; cc < B / 4; ++cc) {         // band columns j0 .. j0+L-1 (zero-filled beyond), a quarter per warp
          const int c = warp * (B / 4) + cc;
          if (c < B - 1 || !boxed) cpa16_cg(&sB[c][2 * lane], Bb + (j0 + c) * LDB + 2 * lane, c < L);
        }
        cpa_commit();
      }
      int fail = 0;
      if (warp == 3 && boxed) {                      // (warp 3 has nothing else to do here) value and flag in one load
        const uint4* e0 = inbox + (size_t)(k + 1) * LDB + lane;
        double x0 = 0.0, x1 = 0.0;
        bool need0 = true, need1 = true;
        unsigned spins = 0;
        while (need0 || need1) {
          if (need0) need0 = !ll_try(e0, inflag, x0);
          if (need1) need1 = !ll_try(e0 + 32, inflag, x1);
          if ((need0 || need1) && ++spins > (1u << 18)) {
            if ((spins & 0xFFFu) == 0 && ld_acquire_u32(err) != 0) { fail = 1; break; }
            if (spins > (1u << 25)) { atomicExch(err, 3u); fail = 1; break; }
          }
        }
        sB[B - 1][lane] = x0;
        sB[B - 1][lane + 32] = x1;
      }
      if (warp == 0) {
        if (k == 0) {                                // reflector of column i (overlaps the copies)
          const double x = (lane < L) ? __ldcg(Bb + i * LDB + 1 + lane) : 0.0;
          double beta;
          warp_house(x, lane, v, tau, beta);
          if (lane < L) __stcg(Bb + i * LDB + 1 + lane, lane == 0 ? beta : 0.0);
        }
        if (lane >= L) v = 0.0;
        if (lane < L) V2[i * N + j0 + lane] = v;
        if (lane == 0) {
          TAU2[i * ntask_max + k] = tau;
          s_tau = tau;
        }
        sv[lane] = v;
      }
      cpa_wait<0>();
      if (__syncthreads_or(fail)) return;
      if (!s_pre[k & 1]) return;                     // the prefetch of this task's columns gave up waiting
      const double tk = s_tau, vl = sv[lane];
      if (warp >= 2) {
        // ---- D <- H D H: p = tau D v (each warp sums its 16 columns), q = p - (tau/2)(p^T v) v, D -= v q^T + q v^T
        double pp[2] = {0.0, 0.0};
#pragma unroll
        for (int cc = 0; cc < B / 2; ++cc) {
          const int c = cb + cc;
          const int idx = (c <= lane) ? c * LDB + lane - c : lane * LDB + c - lane;
          pp[cc & 1] = fma((&sB[0][0])[idx], sv[c], pp[cc & 1]);
        }
        sP[half][lane] = pp[0] + pp[1];
        pair_barrier(2);                             // both halves summed; nobody reads D any more
        const double pl = tk * (sP[0][lane] + sP[1][lane]);
        const double kk = 0.5 * tk * warp_sum(pl * vl);
        const double ql = fma(-kk, vl, pl);
        sqD[half][lane] = ql;                        // each warp keeps its own copy: no second barrier
        __syncwarp();
#pragma unroll
        for (int cc = 0; cc < B / 2; ++cc) {
          const int c = cb + cc;
          if (c <= lane) sB[c][lane - c] -= fma(vl, sqD[half][c], ql * sv[c]);
        }
        __syncwarp();
        if (warp == 2 && !keep0)                     // column 0 is final for this sweep: straight to the next sweep
          ll_store(outbox + (size_t)k * LDB + lane, lane < L ? sB[0][lane] : 0.0, myflag);
#pragma unroll
        for (int cc = 0; cc < B / 2; ++cc) {         // write back column c, d in [0, L - c)
          const int c = cb + cc;
          if (c == 0 && !keep0) continue;
          const int len = L - c, d0 = 2 * lane;
          double* dst = Bb + (j0 + c) * LDB + d0;
          if (d0 + 1 < len) __stcg(reinterpret_cast<double2*>(dst), make_double2(sB[c][d0], sB[c][d0 + 1]));
          else if (d0 < len) __stcg(dst, sB[c][d0]);
        }
        if (k + 1 < nt) {
          // ---- the D warps are done long before the E chain: they fetch the next task's columns 0 .. 30 (written by
          // task k+1 of sweep i-1 -- wait for its release) into the other buffer, off everybody's critical path
          if (threadIdx.x == 64) {
            int ok = 1;
            if (i > 0) {
              const unsigned need = (unsigned)(k + 2);
              unsigned spins = 0;
              while (ld_acquire_u32(prog + i - 1) < need) {
                if (++spins > (1u << 22)) {
                  if ((spins & 0xFFFFu) == 0 && ld_acquire_u32(err) != 0) { ok = 0; break; }
                  if (spins > (1u << 27)) { atomicExch(err, 2u); ok = 0; break; }
                }
              }
            }
            s_pre[(k + 1) & 1] = ok;
          }
          pair_barrier(2);
          if (s_pre[(k + 1) & 1]) {
            double (*sN)[LDB] = sBB[(k + 1) & 1];
            const long jn = j0 + B;
            const int Ln = (int)min((long)B, N - jn);
            const bool boxed_n = i > 0 && Ln == B;
#pragma unroll
            for (int cc = 0; cc < B / 2; ++cc) {
              const int c = cb + cc;
              if (c < B - 1 || !boxed_n) cpa16_cg(&sN[c][2 * lane], Bb + (jn + c) * LDB + 2 * lane, c < Ln);
            }
          }
          cpa_commit();
        }
      } else {
        double v2 = 0.0, tau2 = 0.0;
        if (L2 > 0) {
          // ---- E <- E H (lane = row a of E; each warp its 16 columns)
          double yy[2] = {0.0, 0.0};
#pragma unroll
          for (int cc = 0; cc < B / 2; ++cc) {
            const int c = cb + cc;
            yy[cc & 1] = fma(sB[c][B - c + lane], sv[c], yy[cc & 1]);
          }
          sY[half][lane] = yy[0] + yy[1];
          pair_barrier(1);
          const double yl = tk * (sY[0][lane] + sY[1][lane]);
#pragma unroll
          for (int cc = 0; cc < B / 2; ++cc) {
            const int c = cb + cc;
            sB[c][B - c + lane] = fma(-yl, sv[c], sB[c][B - c + lane]);
          }
          __syncwarp();
          if (warp == 0) {                           // new reflector from the first column of E
            double beta2;
            warp_house(lane < L2 ? sB[0][B + lane] : 0.0, lane, v2, tau2, beta2);
            if (lane >= L2) v2 = 0.0;
            sv2[lane] = v2;
            if (lane == 0) {
              s_tau2 = tau2;
              s_beta2 = beta2;
            }
            // H2 leaves exactly (beta, 0, ..., 0) in the first column of E: the next sweep can have it now
            if (!keep0) ll_store(outbox + (size_t)k * LDB + B + lane, lane == 0 ? beta2 : 0.0, myflag);
          }
          pair_barrier(1);
          const double t2 = s_tau2, v2l = sv2[lane];
          // ---- E <- H2 E: z_c = tau2 v2^T E[:, c]; lanes 0-15 / 16-31 sum rows 0-15 / 16-31 of column cb + (lane & 15)
          {
            const int c = cb + (lane & 15), a0 = (lane >> 4) * 16;
            const double* col = &sB[c][B - c + a0];
            double zz[2] = {0.0, 0.0};
#pragma unroll
            for (int a2 = 0; a2 < 16; ++a2) zz[a2 & 1] = fma(sv2[a0 + a2], col[a2], zz[a2 & 1]);
            double zl = zz[0] + zz[1];
            zl += __shfl_xor_sync(0xffffffffu, zl, 16);
            if (lane < 16) sz[c] = t2 * zl;
          }
          __syncwarp();
#pragma unroll
          for (int cc = 0; cc < B / 2; ++cc) {
            const int c = cb + cc;
            sB[c][B - c + lane] = fma(-v2l, sz[c], sB[c][B - c + lane]);
          }
          if (warp == 0) sB[0][B + lane] = (lane == 0) ? s_beta2 : 0.0;     // exactly (beta, 0, ..., 0)
          __syncwarp();
#pragma unroll
          for (int cc = 0; cc < B / 2; ++cc) {       // write back column c, d = B - c + a
            const int c = cb + cc;
            if (c == 0 && !keep0) continue;
            if (c < L && lane < L2) __stcg(Bb + (j0 + c) * LDB + B - c + lane, sB[c][B - c + lane]);
          }
        } else if (warp == 0 && !keep0) {            // no E block: the lower half of column 0 as it was loaded
          ll_store(outbox + (size_t)k * LDB + B + lane, sB[0][B + lane], myflag);
        }
        if (warp == 1 && boxed) {
          // Column 31 came through the mailbox, so the band never saw the (beta, 0, ..., 0) that closed it in the
          // previous sweep: what the band holds below its first entry is two sweeps old.  Write the whole column
          // (later sweeps read entries of it that no task of this sweep writes).
          double* dst = Bb + (j0 + B - 1) * LDB;
          __stcg(dst + 1 + lane, sB[B - 1][1 + lane]);
          if (lane < B - 1) __stcg(dst + B + 1 + lane, sB[B - 1][B + 1 + lane]);
        }
        v = v2;
        tau = tau2;
      }
      // ---- publish progress: the CTA barrier orders every warp's stores before thread 0's release store
      __syncthreads();
      if (threadIdx.x == 0) st_release_u32(prog + i, (k + 1 == nt) ? 0x7fffffffu : (unsigned)(k + 1));
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Stage 2, two task groups per sweep.  Task k+1 of a sweep needs from task k only the reflector that task k derives
// half-way (from the first column of its E block); the second half of task k (H2 applied to E, write-back) touches
// columns J_k only, which task k+1 never reads.  So a CTA holds TWO groups of four warps that take the even / odd tasks
// of its sweep: group B starts task k+1 as soon as group A has the reflector (named-barrier hand-off through shared
// memory) and the mailbox column of task k+1 has arrived, while group A finishes task k.  Per sweep the start of task
// k then trails the previous sweep's by (half a task + mailbox hop) instead of (a whole task + half a task + hop).
// Everything else as in sbr_chase5_kernel.  Barriers: 1 + g = group g (128 threads), 3 + g / 5 + g = its E / D warp
// pair (64), 7 + g = reflector hand-off INTO group g (warp 0 of each group: 64), 9 + g = "the task before yours has
// released its progress" into group g (64).  No early exits: a timed-out wait raises the error flag and s_dead, every
// other wait in the CTA then falls through, and the (wrong) result is discarded by the host -- barriers stay balanced.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void named_sync(const int id, const int count) {
  asm volatile("bar.sync %0, %1;\n" ::"r"(id), "r"(count) : "memory");
}
__device__ __forceinline__ void named_arrive(const int id, const int count) {
  asm volatile("bar.arrive %0, %1;\n" ::"r"(id), "r"(count) : "memory");
}

template <int NG>
__global__ void __launch_bounds__(128 * NG)
sbr_chase6_kernel(double* __restrict__ Bb, const long N, double* __restrict__ V2, double* __restrict__ TAU2,
                  const long ntask_max, unsigned* __restrict__ bar, uint4* __restrict__ mail) {
  extern __shared__ __align__(16) double smem_chase[];
  double (*sBB)[B][LDB] = reinterpret_cast<double (*)[B][LDB]>(smem_chase);   // group g works in sBB[g]
  __shared__ double sv_[NG][B], sv2_[NG][B], sz_[NG][B], sqD_[NG][2][B], sY_[NG][2][B], sP_[NG][2][B];
  __shared__ double s_tau_[NG], s_tau2_[NG], s_beta2_[NG];
  __shared__ volatile int s_dead;
  const int g = threadIdx.x >> 7;                    // task group: tasks k with k % NG == g
  const int gp = (g + NG - 1) % NG, gn = (g + 1) % NG;   // the groups of tasks k-1 / k+1
  const int lane = threadIdx.x & 31, warp = (threadIdx.x >> 5) & 3;   // warp within the group
  const int half = warp & 1;                         // which 16 columns of its block this warp owns
  const int cb = half * (B / 2);
  const bool leader = (threadIdx.x & 127) == 0;
  double (*sB)[LDB] = sBB[g];
  double* sv = sv_[g];
  double* sv2 = sv2_[g];
  double* sz = sz_[g];
  double (*sqD)[B] = sqD_[g];
  double (*sY)[B] = sY_[g];
  double (*sP)[B] = sP_[g];
  const int GB = 1 + g, EB = 1 + NG + g, DB = 1 + 2 * NG + g, HB_in = 1 + 3 * NG + g, HB_out = 1 + 3 * NG + gn,
            RB_in = 1 + 4 * NG + g, RB_out = 1 + 4 * NG + gn;
  static_assert(5 * NG <= 15, "sixteen named barriers per CTA");
  unsigned* err = bar + 1;
  unsigned* prog = bar + 2;
  const long nsweep = N - 2;
  if (threadIdx.x == 0) s_dead = 0;
  __syncthreads();
  for (long i = blockIdx.x; i < nsweep; i += gridDim.x) {
    const int nt = (int)((N - 1 - i + B - 1) / B);
    const unsigned myflag = (unsigned)(i + 1), inflag = (unsigned)i;       // mailbox flags: sweep index + 1
    uint4* outbox = mail + (size_t)(i & 1) * ntask_max * LDB;              // slot of task k: outbox + k * LDB
    const uint4* inbox = mail + (size_t)((i + 1) & 1) * ntask_max * LDB;   // the previous sweep's
    for (int k = g; k < nt; k += NG) {
      const long j0 = i + 1 + (long)k * B;
      const int L = (int)min((long)B, N - j0);
      const int L2 = (int)max(0L, min((long)B, N - j0 - B));
      const bool boxed = i > 0 && L == B;            // column 31 arrives in the mailbox of task (i-1, k+1)
      const bool keep0 = k == 0;                     // column 0 of task 0 has no mailbox reader: it goes to the band
      // ---- columns 0 .. 30: written by task k of sweep i-1, through the band
      if (leader && i > 0) {
        const unsigned need = (unsigned)(k + 1);
        unsigned spins = 0;
        while (!s_dead && ld_acquire_u32(prog + i - 1) < need) {
          if (++spins > (1u << 22)) {
            if ((spins & 0xFFFFu) == 0 && ld_acquire_u32(err) != 0) { s_dead = 1; break; }
            if (spins > (1u << 27)) { atomicExch(err, 2u); s_dead = 1; break; }
          }
        }
      }
      named_sync(GB, 128);
#pragma unroll
      for (int cc = 0; cc < B / 4; ++cc) {           // band columns j0 .. j0+L-1 (zero-filled beyond), a quarter per warp
        const int c = warp * (B / 4) + cc;
        if (c < B - 1 || !boxed) cpa16_cg(&sB[c][2 * lane], Bb + (j0 + c) * LDB + 2 * lane, c < L);
      }
      cpa_commit();
      if (warp == 3 && boxed) {                      // column 31: value and flag in one load
        const uint4* e0 = inbox + (size_t)(k + 1) * LDB + lane;
        double x0 = 0.0, x1 = 0.0;
        bool need0 = true, need1 = true;
        unsigned spins = 0;
        while ((need0 || need1) && !s_dead) {
          if (need0) need0 = !ll_try(e0, inflag, x0);
          if (need1) need1 = !ll_try(e0 + 32, inflag, x1);
          if ((need0 || need1) && ++spins > (1u << 18)) {
            if ((spins & 0xFFFu) == 0 && ld_acquire_u32(err) != 0) { s_dead = 1; break; }
            if (spins > (1u << 25)) { atomicExch(err, 3u); s_dead = 1; break; }
          }
        }
        sB[B - 1][lane] = (need0 || need1) ? 0.0 : x0;
        sB[B - 1][lane + 32] = (need0 || need1) ? 0.0 : x1;
      }
      if (warp == 0) {
        double v, tau;
        if (k == 0) {                                // reflector of column i (overlaps the copies)
          const double x = (lane < L) ? __ldcg(Bb + i * LDB + 1 + lane) : 0.0;
          double beta;
          warp_house(x, lane, v, tau, beta);
          if (lane < L) __stcg(Bb + i * LDB + 1 + lane, lane == 0 ? beta : 0.0);
        } else {                                     // handed over by the other group, half-way through task k-1
          named_sync(HB_in, 64);
          v = sv2_[gp][lane];
          tau = s_tau2_[gp];
        }
        if (lane >= L) v = 0.0;
        if (lane < L) V2[i * N + j0 + lane] = v;
        if (lane == 0) {
          TAU2[i * ntask_max + k] = tau;
          s_tau_[g] = tau;
        }
        sv[lane] = v;
      }
      cpa_wait<0>();
      named_sync(GB, 128);
      const double tk = s_tau_[g], vl = sv[lane];
      if (warp >= 2) {
        // ---- D <- H D H: p = tau D v (each warp sums its 16 columns), q = p - (tau/2)(p^T v) v, D -= v q^T + q v^T
        double pp[2] = {0.0, 0.0};
#pragma unroll
        for (int cc = 0; cc < B / 2; ++cc) {
          const int c = cb + cc;
          const int idx = (c <= lane) ? c * LDB + lane - c : lane * LDB + c - lane;
          pp[cc & 1] = fma((&sB[0][0])[idx], sv[c], pp[cc & 1]);
        }
        sP[half][lane] = pp[0] + pp[1];
        named_sync(DB, 64);                          // both halves summed; nobody reads D any more
        const double pl = tk * (sP[0][lane] + sP[1][lane]);
        const double kk = 0.5 * tk * warp_sum(pl * vl);
        const double ql = fma(-kk, vl, pl);
        sqD[half][lane] = ql;                        // each warp keeps its own copy: no second barrier
        __syncwarp();
#pragma unroll
        for (int cc = 0; cc < B / 2; ++cc) {
          const int c = cb + cc;
          if (c <= lane) sB[c][lane - c] -= fma(vl, sqD[half][c], ql * sv[c]);
        }
        __syncwarp();
        if (warp == 2 && !keep0)                     // column 0 is final for this sweep: straight to the next sweep
          ll_store(outbox + (size_t)k * LDB + lane, lane < L ? sB[0][lane] : 0.0, myflag);
#pragma unroll
        for (int cc = 0; cc < B / 2; ++cc) {         // write back column c, d in [0, L - c)
          const int c = cb + cc;
          if (c == 0 && !keep0) continue;
          const int len = L - c, d0 = 2 * lane;
          double* dst = Bb + (j0 + c) * LDB + d0;
          if (d0 + 1 < len) __stcg(reinterpret_cast<double2*>(dst), make_double2(sB[c][d0], sB[c][d0 + 1]));
          else if (d0 < len) __stcg(dst, sB[c][d0]);
        }
      } else {
        if (L2 > 0) {
          // ---- E <- E H (lane = row a of E; each warp its 16 columns)
          double yy[2] = {0.0, 0.0};
#pragma unroll
          for (int cc = 0; cc < B / 2; ++cc) {
            const int c = cb + cc;
            yy[cc & 1] = fma(sB[c][B - c + lane], sv[c], yy[cc & 1]);
          }
          sY[half][lane] = yy[0] + yy[1];
          named_sync(EB, 64);
          const double yl = tk * (sY[0][lane] + sY[1][lane]);
#pragma unroll
          for (int cc = 0; cc < B / 2; ++cc) {
            const int c = cb + cc;
            sB[c][B - c + lane] = fma(-yl, sv[c], sB[c][B - c + lane]);
          }
          __syncwarp();
          if (warp == 0) {                           // new reflector from the first column of E
            double v2, tau2, beta2;
            warp_house(lane < L2 ? sB[0][B + lane] : 0.0, lane, v2, tau2, beta2);
            if (lane >= L2) v2 = 0.0;
            sv2[lane] = v2;
            if (lane == 0) {
              s_tau2_[g] = tau2;
              s_beta2_[g] = beta2;
            }
            // H2 leaves exactly (beta, 0, ..., 0) in the first column of E: the next sweep can have it now
            if (!keep0) ll_store(outbox + (size_t)k * LDB + B + lane, lane == 0 ? beta2 : 0.0, myflag);
            if (k + 1 < nt) {                        // ... and the other group the reflector of task k+1
              __threadfence_block();
              named_arrive(HB_out, 64);
            }
          }
          named_sync(EB, 64);
          const double t2 = s_tau2_[g], v2l = sv2[lane];
          // ---- E <- H2 E: z_c = tau2 v2^T E[:, c]; lanes 0-15 / 16-31 sum rows 0-15 / 16-31 of column cb + (lane & 15)
          {
            const int c = cb + (lane & 15), a0 = (lane >> 4) * 16;
            const double* col = &sB[c][B - c + a0];
            double zz[2] = {0.0, 0.0};
#pragma unroll
            for (int a2 = 0; a2 < 16; ++a2) zz[a2 & 1] = fma(sv2[a0 + a2], col[a2], zz[a2 & 1]);
            double zl = zz[0] + zz[1];
            zl += __shfl_xor_sync(0xffffffffu, zl, 16);
            if (lane < 16) sz[c] = t2 * zl;
          }
          __syncwarp();
#pragma unroll
          for (int cc = 0; cc < B / 2; ++cc) {
            const int c = cb + cc;
            sB[c][B - c + lane] = fma(-v2l, sz[c], sB[c][B - c + lane]);
          }
          if (warp == 0) sB[0][B + lane] = (lane == 0) ? s_beta2_[g] : 0.0;     // exactly (beta, 0, ..., 0)
          __syncwarp();
#pragma unroll
          for (int cc = 0; cc < B / 2; ++cc) {       // write back column c, d = B - c + a
            const int c = cb + cc;
            if (c == 0 && !keep0) continue;
            if (c < L && lane < L2) __stcg(Bb + (j0 + c) * LDB + B - c + lane, sB[c][B - c + lane]);
          }
        } else if (warp == 0) {                      // no E block (hence no task k+1): the lower half of column 0 as loaded
          if (!keep0) ll_store(outbox + (size_t)k * LDB + B + lane, sB[0][B + lane], myflag);
        }
        if (warp == 1 && boxed) {
          // Column 31 came through the mailbox, so the band never saw the (beta, 0, ..., 0) that closed it in the
          // previous sweep: write the whole column (later sweeps read entries of it that no task of this sweep writes).
          double* dst = Bb + (j0 + B - 1) * LDB;
          __stcg(dst + 1 + lane, sB[B - 1][1 + lane]);
          if (lane < B - 1) __stcg(dst + B + 1 + lane, sB[B - 1][B + 1 + lane]);
        }
      }
      // ---- publish progress (in task order): the group barrier orders every warp's stores before the release store
      named_sync(GB, 128);
      if (warp == 0) {
        if (k > 0) named_sync(RB_in, 64);            // task k-1 has released
        if (lane == 0) st_release_u32(prog + i, (k + 1 == nt) ? 0x7fffffffu : (unsigned)(k + 1));
        if (k + 1 < nt) named_arrive(RB_out, 64);
      }
    }
    __syncthreads();                                 // both groups done with the sweep (barrier phases realigned)
  }
}

__global__ void __launch_bounds__(256)
sbr_band_to_tridiag_kernel(const double* __restrict__ Bb, const long N, double* __restrict__ d, double* __restrict__ e) {
  const long c = blockIdx.x * 256L + threadIdx.x;
  if (c >= N) return;
  d[c] = Bb[c * LDB];
  if (c + 1 < N) e[c] = Bb[c * LDB + 1];
}

// ------------------------------------------------------------------------------------------------
// Back-transformation: CTA = NV eigenvectors of T held in shared memory (N doubles each).
//   Q2: sweeps in descending order; the reflectors of one sweep act on disjoint index ranges -> one warp each.
//   Q1: panels in descending order: z[r0:] -= V (T (V^T z[r0:])).
// ------------------------------------------------------------------------------------------------
constexpr int BT_THREADS = 512;
constexpr int Q2_THREADS = 1024;

// Q2: one eigenvector per CTA, held in shared memory.  The reflectors of one sweep act on disjoint 32-element
// ranges; EIGHT threads share one reflector (4 entries each, 3 shuffle rounds).  A warp per reflector was bound by
// the shuffle unit (one warp-instruction per clock and SM: 32-36 ms at N = 10^4, 100 vectors), a thread per
// reflector by the latency of 64 dependent shared-memory operations on too few warps (42 ms), and staging each
// reflector row through shared memory by the address arithmetic and the second barrier of the staging (29 ms) --
// profiles/r02_sbr_profile.md.  So each thread fetches exactly the entries it will use, one sweep ahead, straight
// into registers (two register sets swap roles, no copies): one barrier per sweep, and the loop body is 8 LDS + 8
// FMA + 3 shuffles + 4 STS per reflector quarter.  z lives in shared memory with 8 pads per 32 entries, so that the
// two reflectors of a half-warp land in different banks for every sweep offset.
__device__ __forceinline__ long q2_idx(long e) { return e + ((e >> 5) << 3); }
constexpr int Q2F_THREADS = 512;

template <int MAXIT>
__global__ void __launch_bounds__(Q2F_THREADS, 1)
sbr_back_q2_kernel(const double* __restrict__ V2, const double* __restrict__ TAU2, const long ntask_max, const long N,
                   double* __restrict__ Z) {
  extern __shared__ double smem_q2[];                // z[e] at zs[q2_idx(e)], zero tail beyond N
  double* zs = smem_q2;
  const int tid = threadIdx.x;
  const int sub = tid & 7, kt = tid >> 3;
  const unsigned gmask = 0xFFu << (tid & 24);
  double* z = Z + (long)blockIdx.x * N;
  for (long r = tid; r < N + 64; r += Q2F_THREADS) zs[q2_idx(r)] = (r < N) ? z[r] : 0.0;
  // entries sub, sub+8, sub+16, sub+24 of reflector kt + it * 64 of sweep i (zero beyond the matrix), and its tau
  auto fetch = [&](long i, double (&v)[MAXIT][4], double (&t)[MAXIT]) {
    const int nt = (int)((N - 1 - i + B - 1) / B);
    const double* row = V2 + i * N + i + 1;
    const long len = N - 1 - i;
#pragma unroll
    for (int it = 0; it < MAXIT; ++it) {
      const int k = kt + it * (Q2F_THREADS / 8);
      const long o = (long)k * B + sub;
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) v[it][jj] = (o + 8 * jj < len) ? __ldg(row + o + 8 * jj) : 0.0;
      t[it] = (k < nt) ? __ldg(TAU2 + i * ntask_max + k) : 0.0;
    }
  };
  auto apply = [&](long i, const double (&v)[MAXIT][4], const double (&t)[MAXIT]) {
    const int nt = (int)((N - 1 - i + B - 1) / B);
    const int i1 = (int)(i + 1), bo = i1 & 31;
    const int c0 = i1 + ((i1 >> 5) << 3);            // padded index of entry i+1; entry i+1+32k+r sits 40k + r (+8) on
    int off[4];
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      const int r = sub + 8 * jj;
      off[jj] = r + ((bo + r >= 32) ? 8 : 0);
    }
#pragma unroll
    for (int it = 0; it < MAXIT; ++it) {
      const int k = kt + it * (Q2F_THREADS / 8);
      if (k < nt) {                                  // whole groups of 8 threads enter or skip together
        const int cb2 = c0 + 40 * k;
        double zz[4], dot = 0.0;
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          zz[jj] = zs[cb2 + off[jj]];
          dot = fma(v[it][jj], zz[jj], dot);
        }
        dot += __shfl_xor_sync(gmask, dot, 1);
        dot += __shfl_xor_sync(gmask, dot, 2);
        dot += __shfl_xor_sync(gmask, dot, 4);
        const double sfac = -t[it] * dot;
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) zs[cb2 + off[jj]] = fma(sfac, v[it][jj], zz[jj]);
      }
    }
  };
  double va[MAXIT][4], vb[MAXIT][4], ta[MAXIT], tb[MAXIT];
  if (N >= 3) fetch(N - 3, va, ta);
  __syncthreads();
  for (long i = N - 3; i >= 0; i -= 2) {             // two sweeps per trip: the register sets swap roles
    if (i >= 1) fetch(i - 1, vb, tb);
    apply(i, va, ta);
    __syncthreads();
    if (i >= 1) {
      if (i >= 2) fetch(i - 2, va, ta);
      apply(i - 1, vb, tb);
      __syncthreads();
    }
  }
  for (long r = tid; r < N; r += Q2F_THREADS) z[r] = zs[q2_idx(r)];
}

// Q2 for vectors too long for the layout above (two padded arrays): only z in shared memory, one warp per reflector
__global__ void __launch_bounds__(Q2_THREADS, 1)
sbr_back_q2_long_kernel(const double* __restrict__ V2, const double* __restrict__ TAU2, const long ntask_max,
                        const long N, double* __restrict__ Z) {
  extern __shared__ double smem_q2[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int NW = Q2_THREADS / 32;
  double* z = Z + (long)blockIdx.x * N;
  for (long r = tid; r < N; r += Q2_THREADS) smem_q2[r] = z[r];
  __syncthreads();
  for (long i = N - 3; i >= 0; --i) {
    const int nt = (int)((N - 1 - i + B - 1) / B);
    double* zz = smem_q2 + i + 1;
#pragma unroll 2
    for (int k = warp; k < nt; k += NW) {
      const long o = (long)k * B + lane;
      const bool in = i + 1 + o < N;
      const double v = in ? __ldg(V2 + i * N + i + 1 + o) : 0.0;
      const double tau = __ldg(TAU2 + i * ntask_max + k);
      const double x = in ? zz[o] : 0.0;
      const double dot = warp_sum(v * x);
      if (in) zz[o] = fma(-tau * dot, v, x);
    }
    __syncthreads();
  }
  for (long r = tid; r < N; r += Q2_THREADS) z[r] = smem_q2[r];
}

// Q1: NV eigenvectors per CTA share every read of the panel reflectors
template <int NV>
__global__ void __launch_bounds__(BT_THREADS, 1)
sbr_back_q1_kernel(const double* __restrict__ Vall, const double* __restrict__ Tall, const long npanel, const long N,
                   double* __restrict__ Z, const int nvec) {
  extern __shared__ double zs[];                     // [NV][N]
  __shared__ double s_y[NV][B], s_t[NV][B], s_T[B][B + 1], s_part[BT_THREADS / 32][NV][B];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int NW = BT_THREADS / 32;
  const int vec0 = blockIdx.x * NV;
  for (int q = 0; q < NV; ++q)
    for (long r = tid; r < N; r += BT_THREADS) zs[q * N + r] = (vec0 + q < nvec) ? Z[(long)(vec0 + q) * N + r] : 0.0;
  __syncthreads();
  for (long kp = npanel - 1; kp >= 0; --kp) {
    const long c0 = kp * B, r0 = c0 + B;
    const double* Vk = Vall + c0 * N;
    for (int e = tid; e < B * B; e += BT_THREADS) s_T[e / B][e % B] = Tall[kp * B * B + e];
    // y = V^T z: thread strides over rows, partial sums per vector (16 columns at a time) -> warp reduce -> smem
#pragma unroll 1
    for (int mh = 0; mh < B; mh += 16) {
      double acc[NV][16];
#pragma unroll
      for (int q = 0; q < NV; ++q)
#pragma unroll
        for (int m = 0; m < 16; ++m) acc[q][m] = 0.0;
      for (long r = r0 + tid; r < N; r += BT_THREADS) {
        double zz[NV];
#pragma unroll
        for (int q = 0; q < NV; ++q) zz[q] = zs[q * N + r];
#pragma unroll
        for (int m = 0; m < 16; ++m) {
          const double vv = __ldg(Vk + (long)(mh + m) * N + r);
#pragma unroll
          for (int q = 0; q < NV; ++q) acc[q][m] = fma(vv, zz[q], acc[q][m]);
        }
      }
#pragma unroll
      for (int q = 0; q < NV; ++q)
#pragma unroll
        for (int m = 0; m < 16; ++m) {
          const double v = warp_sum(acc[q][m]);
          if (lane == 0) s_part[warp][q][mh + m] = v;
        }
    }
    __syncthreads();
    if (tid < NV * B) {
      const int q = tid / B, m = tid % B;
      double v = 0.0;
#pragma unroll
      for (int w = 0; w < NW; ++w) v += s_part[w][q][m];
      s_y[q][m] = v;
    }
    __syncthreads();
    if (tid < NV * B) {                              // t = T y (T upper triangular)
      const int q = tid / B, m = tid % B;
      double v = 0.0;
      for (int l = m; l < B; ++l) v = fma(s_T[m][l], s_y[q][l], v);
      s_t[q][m] = v;
    }
    __syncthreads();
    for (long r = r0 + tid; r < N; r += BT_THREADS) {
      double dz[NV];
#pragma unroll
      for (int q = 0; q < NV; ++q) dz[q] = 0.0;
#pragma unroll
      for (int m = 0; m < B; ++m) {
        const double vv = __ldg(Vk + (long)m * N + r);
#pragma unroll
        for (int q = 0; q < NV; ++q) dz[q] = fma(vv, s_t[q][m], dz[q]);
      }
#pragma unroll
      for (int q = 0; q < NV; ++q) zs[q * N + r] -= dz[q];
    }
    __syncthreads();
  }
  for (int q = 0; q < NV; ++q)
    if (vec0 + q < nvec)
      for (long r = tid; r < N; r += BT_THREADS) Z[(long)(vec0 + q) * N + r] = zs[q * N + r];
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// workspace layout
// ------------------------------------------------------------------------------------------------
struct SbrWs {
  double *Vall, *Tall, *W, *X, *Xp, *Zp, *S, *mail, *Gq, *Bb, *V2, *TAU2, *mail2;
  unsigned* bar;      // [0] barrier counter, [1] error flag, [2..] sweep progress (N)
  long npanel, ntask_max;
  int max_split, max_qr_ctas;
};

static long sbr_npanel(long N) {
  long k = 0;
  while (N - (k + 1) * B >= 2) ++k;
  return k;
}

static SbrWs sbr_layout(double* ws, long N, size_t* total) {
  SbrWs w;
  w.npanel = sbr_npanel(N);
  w.ntask_max = (N + B - 1) / B + 1;
  w.max_split = 16;
  w.max_qr_ctas = std::max((int)((N + QR_ROWS_MAX - 1) / QR_ROWS_MAX), 64);
  size_t off = 0;
  auto take = [&](size_t cnt) {
    double* q = ws + off;
    off += (cnt + 31) & ~(size_t)31;
    return q;
  };
  w.Vall = take((size_t)N * N);
  w.Tall = take((size_t)(w.npanel + 1) * B * B);
  w.W = take((size_t)B * N);
  w.X = take((size_t)B * N);
  w.Xp = take((size_t)w.max_split * B * N);
  w.Zp = take((size_t)((N + WR - 1) / WR) * B * B);
  w.S = take((size_t)B * B);
  w.mail = take((size_t)2 * (w.max_qr_ctas + 1) * B * 2);      // 16-byte mailbox entries, two parities
  w.Gq = take((size_t)B * B + B);
  w.Bb = take((size_t)N * LDB);
  w.V2 = take((size_t)N * N);
  w.TAU2 = take((size_t)N * w.ntask_max);
  w.mail2 = take((size_t)2 * w.ntask_max * LDB * 2);          // stage-2 mailboxes: [sweep parity][task][64] entries of 16 bytes
  w.bar = reinterpret_cast<unsigned*>(take((size_t)(N + 64) / 2 + 64));
  if (total) *total = off;
  return w;
}

size_t sbr_workspace_doubles(long N) {
  size_t total = 0;
  sbr_layout(nullptr, N, &total);
  return total;
}

// second (high-priority) stream + events per device: the QR of panel k+1 runs beside the bulk of panel k's update
struct SbrStreams {
  cudaStream_t qr = nullptr;
  cudaEvent_t head = nullptr, qr_done = nullptr, t_done = nullptr;
  bool ok = false;
};
static SbrStreams& sbr_streams(int dev) {
  static SbrStreams ss[64];
  SbrStreams& x = ss[dev & 63];
  if (!x.ok) {
    int lo = 0, hi = 0;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);
    x.ok = cudaStreamCreateWithPriority(&x.qr, cudaStreamNonBlocking, hi) == cudaSuccess &&
           cudaEventCreateWithFlags(&x.head, cudaEventDisableTiming) == cudaSuccess &&
           cudaEventCreateWithFlags(&x.qr_done, cudaEventDisableTiming) == cudaSuccess &&
           cudaEventCreateWithFlags(&x.t_done, cudaEventDisableTiming) == cudaSuccess;
  }
  return x;
}

static int sbr_launch_qr(double* A, long N, long k, const SbrWs& w, cudaStream_t st, cudaEvent_t after_qr) {
  static std::atomic<unsigned> epochs{0};
  const int c0 = (int)(k * B);
  const long n = N - c0 - B;
  const int rows = qr_rows_for(n);
  const int G = (int)((n + rows - 1) / rows);
  double* Aarg = A;
  long Narg = N;
  int c0arg = c0;
  double *Varg = w.Vall, *garg = w.Gq;
  uint4* marg = reinterpret_cast<uint4*>(w.mail);
  unsigned earg = (epochs.fetch_add(1u) + 1u) & 0x03FFFFFFu;          // flag = epoch * 64 + step: 26 bits of epoch
  if (earg == 0) earg = (epochs.fetch_add(1u) + 1u) & 0x03FFFFFFu;
  unsigned* barg = w.bar;
  void* args[] = {&Aarg, &Narg, &c0arg, &Varg, &garg, &marg, &earg, &barg};
  const void* fn = rows == 256 ? (const void*)sbr_panel_qr_kernel<256>
                               : (rows == 128 ? (const void*)sbr_panel_qr_kernel<128> : (const void*)sbr_panel_qr_kernel<64>);
  PODSDF_CUDA(cudaLaunchCooperativeKernel(fn, dim3(G), dim3(rows * (B / QR_OCT)), args, 0, st));
  if (after_qr) PODSDF_CUDA(cudaEventRecord(after_qr, st));
  sbr_tfactor_kernel<<<1, B, 0, st>>>(w.Gq, w.Tall + k * B * B);      // off the critical path: T is first used by 1d
  count_launch(2);
  return 0;
}

static int sbr_stage1(double* A, long N, const SbrWs& w, cudaStream_t st) {
  static bool configured[64] = {};
  int dev = 0;
  PODSDF_CUDA(cudaGetDevice(&dev));
  if (!configured[dev & 63]) {
    PODSDF_CUDA(cudaFuncSetAttribute(sbr_symm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)XSMEM));
    PODSDF_CUDA(cudaFuncSetAttribute(sbr_syr2k_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)USMEM));
    configured[dev & 63] = true;
  }
  SbrStreams& ss = sbr_streams(dev);
  if (!ss.ok) return fail(PODSDF_ERR_CUDA, "stream creation failed");
  // mailbox flags of earlier solves (or whatever the allocation held) must not look like this solve's
  PODSDF_CUDA(cudaMemsetAsync(w.mail, 0, sizeof(double) * 4 * (size_t)(w.max_qr_ctas + 1) * B, st));
  if (w.npanel > 0) PODSDF_TRY(sbr_launch_qr(A, N, 0, w, st, nullptr));
  bool t_elsewhere = false;             // the T factor of panel k is being built on the QR stream
  // PODSDF_SBR_TRACE=1 (diagnosis only): events around the phases of every panel, averages printed per quarter
  static const bool trace = getenv("PODSDF_SBR_TRACE") && atoi(getenv("PODSDF_SBR_TRACE")) != 0;
  std::vector<cudaEvent_t> tev;
  auto mark = [&](long k, int slot, cudaStream_t s2) {
    if (!trace) return;
    if (tev.empty()) {
      tev.resize((size_t)w.npanel * 6);
      for (auto& x : tev) cudaEventCreate(&x);
    }
    cudaEventRecord(tev[(size_t)k * 6 + slot], s2);
  };
  for (long k = 0; k < w.npanel; ++k) {
    const int c0 = (int)(k * B);
    const long r0 = c0 + B, n = N - r0;
    double* Vk = w.Vall + (long)c0 * N;
    double* Tk = w.Tall + k * B * B;
    // 1b: X = A22 V, split over K so that the grid fills the machine
    const int row_tiles = (int)((N + XR - 1) / XR - r0 / XR);        // absolute tiles that reach rows >= r0
    const int total_chunks = (int)((N + XKB - 1) / XKB - r0 / XKB);
    // split over K: two CTAs fit an SM, so pick the split whose grid fills whole waves of 2 x 148 CTAs best
    // (at least ~2 waves when the matrix allows it, chunks of whole 128-column blocks)
    int nsplit = 1, cps = ((total_chunks + 7) / 8) * 8;
    {
      double best = -1.0;
      for (int cand = 1; cand <= w.max_split; ++cand) {
        int c2 = (total_chunks + cand - 1) / cand;
        c2 = ((c2 + 7) / 8) * 8;
        const int ns = (total_chunks + c2 - 1) / c2;
        const double ctas = (double)row_tiles * ns, slots = 2.0 * 148.0;
        const double waves = ctas / slots;
        double eff = waves / ceil(waves);
        if (waves < 1.0) eff = waves * 0.9;                  // under-filled grids: more CTAs are better
        if (eff > best + 1e-9) {
          best = eff;
          nsplit = ns;
          cps = c2;
        }
      }
    }
    mark(k, 0, st);
    sbr_symm_kernel<<<dim3(row_tiles, nsplit), 256, XSMEM, st>>>(A, N, r0, Vk, w.Xp, cps);
    // 1c / 1d: W
    const int wct = (int)((n + WR - 1) / WR);
    sbr_w1_kernel<<<wct, 256, 0, st>>>(w.Xp, nsplit, Vk, N, r0, w.X, w.Zp);
    if (t_elsewhere) PODSDF_CUDA(cudaStreamWaitEvent(st, ss.t_done, 0));
    t_elsewhere = false;
    sbr_s_kernel<<<1, B * B, 0, st>>>(Tk, w.Zp, wct, w.S);
    sbr_w2_kernel<<<(int)((n + WR2 - 1) / WR2), 256, 0, st>>>(w.X, Vk, Tk, w.S, N, r0, w.W);
    count_launch(4);
    mark(k, 1, st);
    // 1e: rank-2b update of the lower tiles.  The first tile column (it contains the next panel) goes first; the
    // next panel's QR -- a latency-bound chain of 33 exchanges between a few dozen CTAs -- then runs on the high-priority
    // stream beside the rest of the update.
    const int side = (int)((N + UT - 1) / UT - r0 / UT);
    const bool ahead = k + 1 < w.npanel;
    if (ahead && side > 1) {
      sbr_syr2k_kernel<<<(unsigned)syr2k_ctas(N, r0, 1), U_THREADS, USMEM, st>>>(A, N, r0, Vk, w.W, 1);
      PODSDF_CUDA(cudaEventRecord(ss.head, st));
      PODSDF_CUDA(cudaStreamWaitEvent(ss.qr, ss.head, 0));
      mark(k, 2, st);
      mark(k, 4, ss.qr);
      PODSDF_TRY(sbr_launch_qr(A, N, k + 1, w, ss.qr, ss.qr_done));
      mark(k, 5, ss.qr);
      PODSDF_CUDA(cudaEventRecord(ss.t_done, ss.qr));
      t_elsewhere = true;
      sbr_syr2k_kernel<<<(unsigned)syr2k_ctas(N, r0, 2), U_THREADS, USMEM, st>>>(A, N, r0, Vk, w.W, 2);
      mark(k, 3, st);
      PODSDF_CUDA(cudaStreamWaitEvent(st, ss.qr_done, 0));
      count_launch(2);
    } else {
      sbr_syr2k_kernel<<<(unsigned)syr2k_ctas(N, r0, 0), U_THREADS, USMEM, st>>>(A, N, r0, Vk, w.W, 0);
      count_launch();
      if (ahead) PODSDF_TRY(sbr_launch_qr(A, N, k + 1, w, st, nullptr));
    }
  }
  sbr_extract_band_kernel<<<ceil_div(N * LDB, 256L), 256, 0, st>>>(A, N, w.Bb);
  count_launch();
  PODSDF_CUDA(cudaGetLastError());
  if (trace && !tev.empty()) {
    cudaStreamSynchronize(st);
    cudaStreamSynchronize(ss.qr);
    const long np = w.npanel - 2;       // panels with look-ahead and a successor
    for (int qtr = 0; qtr < 4; ++qtr) {
      double sum[6] = {0, 0, 0, 0, 0, 0};
      long cnt = 0;
      for (long k = qtr * np / 4; k < (qtr + 1) * np / 4; ++k) {
        auto el = [&](long ka, int a, long kb, int b) {
          float ms = 0.f;
          if (cudaEventElapsedTime(&ms, tev[(size_t)ka * 6 + a], tev[(size_t)kb * 6 + b]) != cudaSuccess) { cudaGetLastError(); return 0.0; }
          return (double)ms * 1e3;
        };
        sum[0] += el(k, 0, k, 1);       // symm + W
        sum[1] += el(k, 1, k, 2);       // head column of the update
        sum[2] += el(k, 2, k, 3);       // rest of the update
        sum[3] += el(k, 4, k, 5);       // QR + T factor of panel k + 1 (its stream)
        sum[4] += el(k, 3, k + 1, 0);   // main stream waiting for the QR
        sum[5] += el(k, 0, k + 1, 0);   // whole panel
        ++cnt;
      }
      if (cnt > 0)
        fprintf(stderr, "sbr trace, panels %ld..%ld (us): symm+W %.1f, head %.1f, rest %.1f, QR+T %.1f, wait %.1f, panel %.1f\n",
                qtr * np / 4, (qtr + 1) * np / 4 - 1, sum[0] / cnt, sum[1] / cnt, sum[2] / cnt, sum[3] / cnt, sum[4] / cnt, sum[5] / cnt);
    }
    for (auto& x : tev) cudaEventDestroy(x);
  }
  return 0;
}


static int sbr_stage2(long N, const SbrWs& w, double* d, double* e, cudaStream_t st) {
  // progress counters: one per sweep, zeroed; [0] barrier, [1] error flag keep their values
  PODSDF_CUDA(cudaMemsetAsync(w.bar + 2, 0, sizeof(unsigned) * (size_t)(N + 2), st));
  const long nsweep = N - 2;
  if (nsweep > 0) {
    int dev = 0, sms = 148;
    PODSDF_CUDA(cudaGetDevice(&dev));
    PODSDF_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    double* Bb = w.Bb;
    long Narg = N, ntm = w.ntask_max;
    double *V2 = w.V2, *TAU2 = w.TAU2;
    unsigned* bar = w.bar;
    uint4* mail2 = reinterpret_cast<uint4*>(w.mail2);
    void* args[] = {&Bb, &Narg, &V2, &TAU2, &ntm, &bar, &mail2};
    // PODSDF_CHASE (A/B only): 4 = four warps per task with every hand-off through the band, 5 = four warps, the critical column through mailboxes, 7 = three task groups per sweep (67 ms: twelve
    // latency-bound warps on an SM slow each other down); default: two task groups (47 ms)
    const char* variant = getenv("PODSDF_CHASE");
    const int var = variant ? atoi(variant) : 0;
    const void* kern = (var == 4) ? (const void*)sbr_chase4_kernel
                     : (var == 5) ? (const void*)sbr_chase5_kernel
                     : (var == 7) ? (const void*)sbr_chase6_kernel<3> : (const void*)sbr_chase6_kernel<2>;
    const int groups = (var == 7) ? 3 : 2;
    const int threads = (var == 4 || var == 5) ? 128 : 128 * groups;
    const size_t dyn = (var == 4 || var == 5) ? 0 : (size_t)groups * B * LDB * sizeof(double);
    if (dyn > 0) PODSDF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
    int per_sm = 1;
    PODSDF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, dyn));
    // mailbox flags of an earlier solve must not look like this one's (flag = sweep + 1 > 0)
    PODSDF_CUDA(cudaMemsetAsync(w.mail2, 0, sizeof(double) * 4 * (size_t)w.ntask_max * LDB, st));
    long G = (long)sms * (per_sm < 4 ? per_sm : 4);             // all chains co-resident (cooperative launch checks it)
    const long G_max = G;
    // the sweeps advance in step, so the slowest CTA sets the pace: with the two-group kernel an SM that hosts two CTAs
    // (sixteen latency-bound warps) holds everybody back (N = 10^4: 52 ms with 164 CTAs on 148 SMs, 48 ms with 148)
    if (var != 4 && var != 5 && G > sms) G = sms;
    const long useful = (N + 2 * B - 1) / (2 * B) + 8;          // sweeps that can be in flight at once
    if (G > useful) G = useful;
    if (const char* og = getenv("PODSDF_CHASE_G")) G = std::min<long>(G_max, std::max(1, atoi(og)));   // A/B only
    if (G > nsweep) G = nsweep;
    PODSDF_CUDA(cudaLaunchCooperativeKernel(kern, dim3((unsigned)G), dim3(threads), args, dyn, st));
    count_launch();
  }
  sbr_band_to_tridiag_kernel<<<ceil_div(N, 256L), 256, 0, st>>>(w.Bb, N, d, e);
  count_launch();
  PODSDF_CUDA(cudaGetLastError());
  return 0;
}

int sbr_tridiagonalize(double* A, long N, double* d, double* e, double* ws, cudaStream_t st, float* timings) {
  if (N < 3) return fail(PODSDF_ERR_INVALID, "sbr_tridiagonalize needs N >= 3");
  SbrWs w = sbr_layout(ws, N, nullptr);
  PODSDF_CUDA(cudaMemsetAsync(w.bar, 0, 2 * sizeof(unsigned), st));
  cudaEvent_t ev[3] = {nullptr, nullptr, nullptr};
  if (timings) {
    for (auto& x : ev) PODSDF_CUDA(cudaEventCreate(&x));
    cudaEventRecord(ev[0], st);
  }
  PODSDF_TRY(sbr_stage1(A, N, w, st));
  if (timings) cudaEventRecord(ev[1], st);
  PODSDF_TRY(sbr_stage2(N, w, d, e, st));
  if (timings) cudaEventRecord(ev[2], st);
  unsigned herr = 0;
  PODSDF_CUDA(cudaMemcpyAsync(&herr, w.bar + 1, sizeof(unsigned), cudaMemcpyDeviceToHost, st));
  PODSDF_CUDA(cudaStreamSynchronize(st));
  if (timings) {
    cudaEventElapsedTime(&timings[0], ev[0], ev[1]);
    cudaEventElapsedTime(&timings[1], ev[1], ev[2]);
    for (auto& x : ev) cudaEventDestroy(x);
  }
  if (herr != 0) return fail(PODSDF_ERR_CUDA, "tridiagonalisation: a grid-wide wait timed out (code " + std::to_string(herr) + ")");
  return 0;
}

int sbr_stage1_only(double* A, long N, double* ws, double* band, cudaStream_t st) {
  SbrWs w = sbr_layout(ws, N, nullptr);
  PODSDF_CUDA(cudaMemsetAsync(w.bar, 0, 2 * sizeof(unsigned), st));
  PODSDF_TRY(sbr_stage1(A, N, w, st));
  PODSDF_CUDA(cudaMemcpyAsync(band, w.Bb, sizeof(double) * (size_t)N * LDB, cudaMemcpyDeviceToDevice, st));
  unsigned herr = 0;
  PODSDF_CUDA(cudaMemcpyAsync(&herr, w.bar + 1, sizeof(unsigned), cudaMemcpyDeviceToHost, st));
  PODSDF_CUDA(cudaStreamSynchronize(st));
  if (herr != 0) return fail(PODSDF_ERR_CUDA, "band reduction: a wait between the CTAs of the panel QR timed out");
  return 0;
}

int sbr_back_transform(const double* ws, long N, double* Z, int nvec, cudaStream_t st) {
  SbrWs w = sbr_layout(const_cast<double*>(ws), N, nullptr);
  const size_t one = sizeof(double) * (size_t)N;
  int dev = 0, max_smem = 0;
  PODSDF_CUDA(cudaGetDevice(&dev));
  PODSDF_CUDA(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  const size_t fixed = 20 * 1024;                            // static shared memory of the Q1 kernels (17.6 KB at NV = 2)
  if (one + fixed > (size_t)max_smem)
    return fail(PODSDF_ERR_UNSUPPORTED, "back-transformation: one eigenvector does not fit shared memory");
  {
    const long its = (w.ntask_max + Q2F_THREADS / 8 - 1) / (Q2F_THREADS / 8);   // reflectors per thread group and sweep
    const size_t q2_smem = sizeof(double) * (size_t)(N + 8 * (N / 32) + 16 + 96);
#define PODSDF_Q2(M)                                                                                         \
  do {                                                                                                       \
    auto k = sbr_back_q2_kernel<M>;                                                                          \
    PODSDF_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)q2_smem));         \
    k<<<nvec, Q2F_THREADS, q2_smem, st>>>(w.V2, w.TAU2, w.ntask_max, N, Z);                                  \
  } while (0)
    if (q2_smem + 2048 > (size_t)max_smem || its > 6) {
      if (one + 2048 > (size_t)max_smem)
        return fail(PODSDF_ERR_UNSUPPORTED, "back-transformation: one eigenvector does not fit shared memory");
      PODSDF_CUDA(cudaFuncSetAttribute(sbr_back_q2_long_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)one));
      sbr_back_q2_long_kernel<<<nvec, Q2_THREADS, one, st>>>(w.V2, w.TAU2, w.ntask_max, N, Z);
    } else if (its <= 2) PODSDF_Q2(2);
    else if (its <= 4) PODSDF_Q2(4);
    else if (its <= 5) PODSDF_Q2(5);
    else PODSDF_Q2(6);
#undef PODSDF_Q2
  }
  if (2 * one + fixed <= (size_t)max_smem) {
    auto k = sbr_back_q1_kernel<2>;
    PODSDF_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(2 * one)));
    k<<<(nvec + 1) / 2, BT_THREADS, 2 * one, st>>>(w.Vall, w.Tall, w.npanel, N, Z, nvec);
  } else {
    auto k = sbr_back_q1_kernel<1>;
    PODSDF_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)one));
    k<<<nvec, BT_THREADS, one, st>>>(w.Vall, w.Tall, w.npanel, N, Z, nvec);
  }
  count_launch(2);
  PODSDF_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace podsdf
