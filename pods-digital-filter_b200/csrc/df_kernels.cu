// Digital-filter kernels for sm_100a (Klein 2003 filter + Lund transform).
//
// Reference behaviour being replaced (file:line under /root/reference):
//   digitalfilters.py:100-140  filter3DSciPy1D -- three chained 'valid' direct convolutions x->y->z
//   digitalfilters.py:143-485  adapt1d / adapt2prf / adapt2d -- y = A x + mean per point
//   digitalfilters.py:1119-1131 rotate_velocity -- constant 3x3 rotation per point
//   digitalfilters.py:1454-1467 np.roll + one new random slab per component per step
//
// B200 design: the contraction is linear and separable, so it is split as
//   K-time : T1[t] = sum_m bx[m] X[t+m]      per padded point, sliding window of raw randoms in
//            REGISTERS (one new slab read or Philox-generated per step, each random touched once)
//   K-plane: out[t] = Lund( by (*) bz (*) T1[t] )   marching down the rows of a 64-column strip:
//            z-pass from a cp.async-staged row block into a shared-memory ring of z-filtered rows,
//            y-pass as register-blocked column windows over that ring, fused with the 3x3
//            Lund/rotation transform and the mean add, so every output plane is written once.
// Both are FP64-FMA work at the DFMA pipe rate with HBM traffic = algorithmic bytes; no GEMM shape.
#include "df_kernels.cuh"
#include "philox.cuh"

namespace podsdf {

// ------------------------------------------------------------------------------------------------
// Philox slabs
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) df_random_slabs_kernel(double* __restrict__ X, long Np, long npair,
                                                              uint64_t seed, long s0) {
  const long pair = blockIdx.x * 256L + threadIdx.x;
  if (pair >= npair) return;
  const int c = blockIdx.y;
  const long s = blockIdx.z;
  double a, b;
  philox_pair(seed, (uint64_t)pair, (uint64_t)(s0 + s), c, a, b);
  double* dst = X + (s * 3 + c) * Np + 2 * pair;
  dst[0] = a;
  if (2 * pair + 1 < Np) dst[1] = b;
}

int launch_df_random_slabs(const DfGeom& g, uint64_t seed, long s0, long n, double* X, cudaStream_t st) {
  if (n <= 0) return 0;
  const long npair = (g.Np + 1) / 2;
  for (long done = 0; done < n; done += 32768) {  // gridDim.z limit
    const long cnt = (n - done < 32768) ? n - done : 32768;
    dim3 grid(ceil_div(npair, 256), 3, (unsigned)cnt);
    df_random_slabs_kernel<<<grid, 256, 0, st>>>(X + done * 3 * g.Np, g.Np, npair, seed, s0 + done);
    count_launch();
  }
  PODSDF_CUDA(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------------
// K-time, tuned: window of NT raw randoms per point in registers (static indices via unroll by NT)
// ------------------------------------------------------------------------------------------------
template <int NT, bool GEN>
__global__ void __launch_bounds__(256, (NT <= 21 ? 2 : 1))
df_time_tuned_kernel(const double* __restrict__ X, double* __restrict__ T1, const Taps bx, const long Np,
                     const long npair, const int nsteps, const int tchunk, const uint64_t seed,
                     const long s_base) {
  const long pair = blockIdx.x * 256L + threadIdx.x;
  if (pair >= npair) return;
  const int c = blockIdx.y;
  const int t0 = blockIdx.z * tchunk;
  const int t1 = min(t0 + tchunk, nsteps);
  const bool has2 = 2 * pair + 1 < Np;
  const bool vec = (Np & 1) == 0;  // every [s][c] plane stays 16-byte aligned
  const int last_slab = nsteps + NT - 2;

  double w0[NT], w1[NT];
  auto load = [&](int s, double& a, double& b) {
    if (GEN) {
      philox_pair(seed, (uint64_t)pair, (uint64_t)(s_base + s), c, a, b);
    } else {
      const double* p = X + ((long)s * 3 + c) * Np + 2 * pair;
      if (vec) {
        const double2 v = __ldcs(reinterpret_cast<const double2*>(p));  // read once: streaming
        a = v.x;
        b = v.y;
      } else {
        a = __ldcs(p);
        b = has2 ? __ldcs(p + 1) : 0.0;
      }
    }
  };
#pragma unroll
  for (int i = 0; i < NT - 1; ++i) load(t0 + i, w0[i], w1[i]);

  for (int t = t0; t < t1; t += NT) {
#pragma unroll
    for (int u = 0; u < NT; ++u) {
      const int tt = t + u;
      load(min(tt + NT - 1, last_slab), w0[(u + NT - 1) % NT], w1[(u + NT - 1) % NT]);
      // true convolution (scipy.signal.convolve flips the taps): tap m meets slab tt + NT-1-m
      double a0 = 0.0, a1 = 0.0, b0 = 0.0, b1 = 0.0;
#pragma unroll
      for (int m = 0; m < NT; ++m) {
        const int idx = (u + NT - 1 - m) % NT;
        if (m & 1) {
          a1 = fma(bx.b[m], w0[idx], a1);
          b1 = fma(bx.b[m], w1[idx], b1);
        } else {
          a0 = fma(bx.b[m], w0[idx], a0);
          b0 = fma(bx.b[m], w1[idx], b0);
        }
      }
      if (tt < t1) {
        double* q = T1 + ((long)tt * 3 + c) * Np + 2 * pair;
        if (vec) {
          *reinterpret_cast<double2*>(q) = make_double2(a0 + a1, b0 + b1);
        } else {
          q[0] = a0 + a1;
          if (has2) q[1] = b0 + b1;
        }
      }
    }
  }
}

// K-time, generic: any tap count, window re-read from global (L1/L2 catch the re-reads)
__global__ void __launch_bounds__(256)
df_time_generic_kernel(const double* __restrict__ X, double* __restrict__ T1, const Taps bx, const int nt,
                       const long Np, const int nsteps) {
  const long i = blockIdx.x * 256L + threadIdx.x;
  if (i >= Np) return;
  const int c = blockIdx.y;
  for (int t = blockIdx.z; t < nsteps; t += gridDim.z) {
    double acc = 0.0;
    for (int m = 0; m < nt; ++m) acc = fma(bx.b[m], X[((long)(t + nt - 1 - m) * 3 + c) * Np + i], acc);
    T1[((long)t * 3 + c) * Np + i] = acc;
  }
}

#define PODSDF_TIME_TAPS(F) F(9) F(13) F(17) F(21) F(25) F(33)

bool df_time_has_tuned(int ntx) {
#define F(N) if (ntx == N) return true;
  PODSDF_TIME_TAPS(F)
#undef F
  return false;
}

template <int NT>
static void launch_time_tuned(const Taps& bx, const double* X, double* T1, long Np, int nsteps, bool gen,
                              uint64_t seed, long s_base, cudaStream_t st) {
  const long npair = (Np + 1) / 2;
  // time chunks: enough CTAs to smooth the tail, few enough that the NT-1 slab warm-up stays small
  int nchunk = 1;
  const long ctas_per_chunk = 3L * ceil_div(npair, 256);
  while (nchunk < 64 && ctas_per_chunk * nchunk < 148L * 2 * 8 && nsteps / (nchunk * 2) >= 8 * NT) nchunk *= 2;
  int tchunk = ceil_div(nsteps, nchunk);
  tchunk = ceil_div(tchunk, NT) * NT;
  nchunk = ceil_div(nsteps, tchunk);
  dim3 grid(ceil_div(npair, 256), 3, nchunk);
  if (gen)
    df_time_tuned_kernel<NT, true><<<grid, 256, 0, st>>>(X, T1, bx, Np, npair, nsteps, tchunk, seed, s_base);
  else
    df_time_tuned_kernel<NT, false><<<grid, 256, 0, st>>>(X, T1, bx, Np, npair, nsteps, tchunk, seed, s_base);
  count_launch();
}

int launch_df_time(const DfGeom& g, const Taps& bx, const double* X, double* T1, int nsteps, bool gen,
                   uint64_t seed, long s_base, bool tuned, cudaStream_t st) {
  if (nsteps <= 0) return 0;
  const int nt = 2 * g.nfx + 1;
  if (tuned) {
    switch (nt) {
#define F(N) case N: launch_time_tuned<N>(bx, X, T1, g.Np, nsteps, gen, seed, s_base, st); break;
      PODSDF_TIME_TAPS(F)
#undef F
      default: return fail(PODSDF_ERR_UNSUPPORTED, "no tuned time-filter kernel for this nfx");
    }
  } else {
    if (gen) return fail(PODSDF_ERR_INVALID, "generic time filter needs materialised slabs");
    int zt = nsteps < 1024 ? nsteps : 1024;
    dim3 grid(ceil_div(g.Np, 256), 3, zt);
    df_time_generic_kernel<<<grid, 256, 0, st>>>(X, T1, bx, nt, g.Np, nsteps);
    count_launch();
  }
  PODSDF_CUDA(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------------
// Lund / rotation epilogue: y = M x + m with only the planes named in `mask` present
// ------------------------------------------------------------------------------------------------
template <typename Idx>
__device__ __forceinline__ void lund_point(const double* __restrict__ coef, const unsigned mask, const long P,
                                           const Idx pt, const double x0, const double x1, const double x2,
                                           double (&y)[3]) {
  const double x[3] = {x0, x1, x2};
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    double acc = 0.0;
#pragma unroll
    for (int d = 0; d < 3; ++d)
      if (mask & (1u << (3 * c + d))) acc = fma(__ldg(coef + (long)(3 * c + d) * P + pt), x[d], acc);
    if (mask & (1u << (9 + c))) acc += __ldg(coef + (long)(9 + c) * P + pt);
    y[c] = acc;
  }
}

__global__ void __launch_bounds__(256)
df_lund_kernel(double* __restrict__ planes, const double* __restrict__ coef, const unsigned mask, const long P) {
  const long pt = blockIdx.x * 256L + threadIdx.x;
  if (pt >= P) return;
  double y[3];
  lund_point(coef, mask, P, pt, planes[pt], planes[P + pt], planes[2 * P + pt], y);
  planes[pt] = y[0];
  planes[P + pt] = y[1];
  planes[2 * P + pt] = y[2];
}

int launch_df_lund(const DfGeom& g, double* planes, const double* coef, unsigned mask, cudaStream_t st) {
  df_lund_kernel<<<ceil_div(g.P, 256), 256, 0, st>>>(planes, coef, mask, g.P);
  count_launch();
  PODSDF_CUDA(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------------
// K-plane, tuned: marching strip kernel
//   CTA = (64-column strip, plane), 128 threads, marches down the padded rows SJ = 16 at a time:
//     stage  : cp.async of the 16 new input rows x (64 + NTZ-1) columns x 3 components
//     z-pass : thread = (row, 8-column segment): 8 outputs from a (8+NTZ-1)-value register window,
//              written to a shared ring of z-filtered rows (NB blocks of 16 rows)
//     y-pass : thread = (column, 8-row segment) for all 3 components: register windows down the
//              ring column; block/row offsets are compile-time (only NB block bases rotate)
//     Lund   : y = M x + m with the coefficient planes, streaming stores of the finished rows
//   Odd shared-memory pitches make both lane<->row (z) and lane<->column (y) access conflict free.
//   Taps live in registers (2 x <=33 doubles); every DFMA has register operands only.
// ------------------------------------------------------------------------------------------------
template <int NTY, int NTZ>
struct PlaneCfg {
  static constexpr int TK = 64;    // output columns per strip
  static constexpr int SJ = 16;    // rows per marching step
  static constexpr int R = 8;      // outputs per thread along the filtered axis (register window)
  static constexpr int NTHR = 128;
  static constexpr int TKIN = TK + NTZ - 1;
  static constexpr int PIN = TKIN | 1;   // odd pitch: lanes <-> rows are bank-conflict free
  static constexpr int NB = (SJ + NTY - 1 + SJ - 1) / SJ;  // ring blocks of SJ rows
  static constexpr int RJ = NB * SJ;
  static constexpr int PT2 = TK + 1;
  static constexpr int STAGE = 3 * SJ * PIN;
  static constexpr int RING = 3 * RJ * PT2;
  static constexpr size_t SMEM = (size_t)(STAGE + RING) * sizeof(double);
  static constexpr int QL = (TKIN + 31) / 32;  // cp.async per lane per staged row
  static_assert(SJ * (TK / R) == NTHR, "z-pass: one item per thread per component");
  static_assert(TK * (SJ / R) == NTHR, "y-pass: one item per thread");
  static_assert(SMEM <= 113 * 1024, "two CTAs per SM");
};

__device__ __forceinline__ void cp_async8(unsigned smem_dst, const double* gsrc, bool valid) {
  const int bytes = valid ? 8 : 0;  // src-size 0 => zero fill
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(smem_dst), "l"(gsrc), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory");
}

// Lund / rotation epilogue of one y-pass item: rows o in [olo, ohi) of column k, all 3 components.
// Plane-major so each mask bit is tested once per item; GUARD = item straddles the plane edge.
template <int R, bool GUARD>
__device__ __forceinline__ void plane_epilogue(const double (&f)[3][R], double* __restrict__ dstp,
                                               const double* __restrict__ coef, const unsigned mask,
                                               const int lund, const long P, const int K, const int pt0,
                                               const int olo, const int ohi) {
  int pt[R];
#pragma unroll
  for (int o = 0; o < R; ++o) pt[o] = pt0 + ((GUARD && (o < olo || o >= ohi)) ? olo : o) * K;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    double y[R];
    if (lund) {
#pragma unroll
      for (int o = 0; o < R; ++o) y[o] = 0.0;
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        if (mask & (1u << (3 * c + d))) {
          const double* cp = coef + (long)(3 * c + d) * P;
#pragma unroll
          for (int o = 0; o < R; ++o) y[o] = fma(__ldg(cp + pt[o]), f[d][o], y[o]);
        }
      }
      if (mask & (1u << (9 + c))) {
        const double* cp = coef + (long)(9 + c) * P;
#pragma unroll
        for (int o = 0; o < R; ++o) y[o] += __ldg(cp + pt[o]);
      }
    } else {
#pragma unroll
      for (int o = 0; o < R; ++o) y[o] = f[c][o];
    }
    double* op = dstp + (long)c * P;
#pragma unroll
    for (int o = 0; o < R; ++o)
      if (!GUARD || (o >= olo && o < ohi)) __stcs(op + pt[o], y[o]);
  }
}

// keep a value in a (non-uniform) register: stops ptxas from re-fetching taps from the constant
// bank / uniform registers inside the marching loop
__device__ __forceinline__ void pin_reg(double& v) { asm volatile("" : "+d"(v)); }

// y-pass of one thread: column kk, output rows j0 .. j0+R-1 (j0 = n*SJ - (NTY-1) + SG*R), 3 comps.
// blk[d] = ring base (component 0, column kk) of the block written d steps ago.
template <int NTY, int NTZ, int SG>
__device__ __forceinline__ void plane_ypass(const double* const (&blk)[PlaneCfg<NTY, NTZ>::NB],
                                            const double (&ty)[NTY], double (&f)[3][PlaneCfg<NTY, NTZ>::R]) {
  using C = PlaneCfg<NTY, NTZ>;
  constexpr int R = C::R, SJ = C::SJ, PT2 = C::PT2, RJ = C::RJ;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
#pragma unroll
    for (int i = 0; i < NTY + R - 1; ++i) {
      constexpr int D = NTY - 1;
      const int off = i + SG * R - D;                       // row relative to n*SJ (compile time)
      const int d = (off >= 0) ? 0 : (-off + SJ - 1) / SJ;  // blocks back
      const int row = off + d * SJ;                         // row inside that block
      const double v = blk[d][(c * RJ + row) * PT2];
#pragma unroll
      for (int o = 0; o < R; ++o) {
        const int m = i - o;
        if (m == 0) f[c][o] = ty[NTY - 1] * v;
        else if (m > 0 && m < NTY) f[c][o] = fma(ty[NTY - 1 - m], v, f[c][o]);
      }
    }
  }
}

template <int NTY, int NTZ>
__global__ void __launch_bounds__(128, 2)
df_plane_tuned_kernel(const double* __restrict__ T1, double* __restrict__ out, const Taps by, const Taps bz,
                      const DfGeom g, const double* __restrict__ coef, const unsigned mask, const int lund) {
  using C = PlaneCfg<NTY, NTZ>;
  constexpr int TK = C::TK, SJ = C::SJ, R = C::R, PIN = C::PIN, RJ = C::RJ, PT2 = C::PT2, NB = C::NB;
  extern __shared__ double smem[];
  double* stage = smem;
  double* ring = smem + C::STAGE;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int k0 = blockIdx.x * TK;
  const long plane = blockIdx.y;
  const double* src = T1 + plane * 3 * g.Np;
  double* dstp = out + plane * 3 * g.P;
  const int nsteps = (g.Jp + SJ - 1) / SJ;

  double ty[NTY], tz[NTZ];
#pragma unroll
  for (int m = 0; m < NTY; ++m) { ty[m] = by.b[m]; pin_reg(ty[m]); }
#pragma unroll
  for (int m = 0; m < NTZ; ++m) { tz[m] = bz.b[m]; pin_reg(tz[m]); }

  // staging: warp w copies rows rc = w + 4*it (rc = c*SJ + r  =>  c = it/4, r = w + 4*(it%4));
  // out-of-range rows / columns are zero-filled (cp.async src-size 0, address never dereferenced)
  const unsigned stage_s = (unsigned)__cvta_generic_to_shared(stage) + (unsigned)((warp * PIN + lane) * 8);
  const long row4 = 4L * g.Kp;
  bool colok[C::QL];
#pragma unroll
  for (int jj = 0; jj < C::QL; ++jj) colok[jj] = (lane + 32 * jj < C::TKIN) && (k0 + lane + 32 * jj < g.Kp);
  const double* gstage = src + (long)warp * g.Kp + k0 + lane;  // advanced by SJ rows per step
  auto issue_stage = [&](int n) {
    const int rows_left = g.Jp - n * SJ - warp;  // row r = warp + 4*rr valid iff 4*rr < rows_left
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const double* gr = gstage + (long)c * g.Np;
#pragma unroll
      for (int rr = 0; rr < SJ / 4; ++rr) {
        const bool rowok = 4 * rr < rows_left;
#pragma unroll
        for (int jj = 0; jj < C::QL; ++jj) {
          if (lane + 32 * jj < C::TKIN)
            cp_async8(stage_s + (unsigned)(((4 * (c * (SJ / 4) + rr)) * PIN + 32 * jj) * 8), gr + 32 * jj,
                      rowok && colok[jj]);
        }
        gr += row4;
      }
    }
    gstage += (long)SJ * g.Kp;
  };

  // z-pass role: row r, column segment sgm
  const int zr = tid & (SJ - 1), zs = tid / SJ;
  const double* zsrc = stage + zr * PIN + zs * R;
  double* zdst = ring + zr * PT2 + zs * R;
  // y-pass role: column kk, row segment sg
  const int kk = tid & (TK - 1), sg = tid / TK;
  const int k = k0 + kk;

  issue_stage(0);
  for (int n = 0; n < nsteps; ++n) {
    cp_async_commit_wait_all();
    __syncthreads();  // stage(n) landed; previous y-pass done with the ring block rewritten now
    const int bn = n % NB;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const double* srow = zsrc + c * SJ * PIN;
      double acc[R];
#pragma unroll
      for (int i = 0; i < NTZ + R - 1; ++i) {
        const double v = srow[i];
#pragma unroll
        for (int o = 0; o < R; ++o) {
          const int m = i - o;
          if (m == 0) acc[o] = tz[NTZ - 1] * v;
          else if (m > 0 && m < NTZ) acc[o] = fma(tz[NTZ - 1 - m], v, acc[o]);
        }
      }
      double* dst = zdst + (c * RJ + bn * SJ) * PT2;
#pragma unroll
      for (int o = 0; o < R; ++o) dst[o] = acc[o];
    }
    __syncthreads();  // ring block of this step complete, stage free
    if (n + 1 < nsteps) issue_stage(n + 1);

    const int j0 = n * SJ - (NTY - 1) + sg * R;
    if (j0 + R > 0 && j0 < g.J) {
      const double* blk[NB];
#pragma unroll
      for (int d = 0; d < NB; ++d) blk[d] = ring + ((n + NB - d) % NB) * SJ * PT2 + kk;
      double f[3][R];
      if (sg == 0)
        plane_ypass<NTY, NTZ, 0>(blk, ty, f);
      else
        plane_ypass<NTY, NTZ, 1>(blk, ty, f);
      if (k < g.K) {
        // 32-bit point indices (P < 2^31): one IMAD.WIDE per access instead of 64-bit pointer chains
        const int pt0 = j0 * g.K + k;
        if (j0 >= 0 && j0 + R <= g.J)
          plane_epilogue<R, false>(f, dstp, coef, mask, lund, g.P, g.K, pt0, 0, R);
        else
          plane_epilogue<R, true>(f, dstp, coef, mask, lund, g.P, g.K, pt0, max(0, -j0), min(R, g.J - j0));
      }
    }
  }
}

// K-plane, generic: two plain kernels through a global intermediate (any tap counts)
__global__ void __launch_bounds__(256)
df_zpass_generic_kernel(const double* __restrict__ T1, double* __restrict__ T2, const Taps bz, const int ntz,
                        const int Jp, const int Kp, const int K, const long nrows /* planes*comps*Jp */) {
  const long e = blockIdx.x * 256L + threadIdx.x;
  if (e >= nrows * K) return;
  const long row = e / K;
  const int k = (int)(e % K);
  const double* s = T1 + row * Kp + k;
  double acc = 0.0;
  for (int m = 0; m < ntz; ++m) acc = fma(bz.b[ntz - 1 - m], s[m], acc);
  T2[e] = acc;
}

__global__ void __launch_bounds__(256)
df_ypass_generic_kernel(const double* __restrict__ T2, double* __restrict__ out, const Taps by, const int nty,
                        const DfGeom g, const int ncomp, const double* __restrict__ coef, const unsigned mask,
                        const int lund) {
  const long pt = blockIdx.x * 256L + threadIdx.x;
  if (pt >= g.P) return;
  const long plane = blockIdx.y;
  const int j = (int)(pt / g.K), k = (int)(pt % g.K);
  double f[3] = {0.0, 0.0, 0.0};
  for (int c = 0; c < ncomp; ++c) {
    const double* s = T2 + ((plane * ncomp + c) * g.Jp + j) * (long)g.K + k;
    double acc = 0.0;
    for (int m = 0; m < nty; ++m) acc = fma(by.b[nty - 1 - m], s[(long)m * g.K], acc);
    f[c] = acc;
  }
  if (lund) {
    double y[3];
    lund_point(coef, mask, g.P, pt, f[0], f[1], f[2], y);
    for (int c = 0; c < 3; ++c) out[(plane * 3 + c) * g.P + pt] = y[c];
  } else {
    for (int c = 0; c < ncomp; ++c) out[(plane * ncomp + c) * g.P + pt] = f[c];
  }
}

// (NTY, NTZ) pairs with a tuned kernel: the isotropic sizes of configs 1, 3, 5 and config 2's (8, 4)
#define PODSDF_PLANE_TAPS(F) F(9, 9) F(13, 13) F(17, 17) F(21, 21) F(25, 25) F(33, 33) F(17, 9)

bool df_plane_has_tuned(int nty, int ntz) {
#define F(A, B) if (nty == A && ntz == B) return true;
  PODSDF_PLANE_TAPS(F)
#undef F
  return false;
}

template <int NTY, int NTZ>
static int launch_plane_tuned(const DfGeom& g, const Taps& by, const Taps& bz, const double* T1, double* out,
                              int nplanes, const double* coef, unsigned mask, bool lund, cudaStream_t st) {
  using C = PlaneCfg<NTY, NTZ>;
  auto kern = df_plane_tuned_kernel<NTY, NTZ>;
  static bool configured[64] = {};
  int dev = 0;
  PODSDF_CUDA(cudaGetDevice(&dev));
  if (!configured[dev & 63]) {
    PODSDF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM));
    configured[dev & 63] = true;
  }
  for (int done = 0; done < nplanes; done += 65535) {
    const int cnt = (nplanes - done < 65535) ? nplanes - done : 65535;
    dim3 grid(ceil_div(g.K, C::TK), cnt);
    kern<<<grid, C::NTHR, C::SMEM, st>>>(T1 + (long)done * 3 * g.Np, out + (long)done * 3 * g.P, by, bz, g, coef,
                                         mask, lund ? 1 : 0);
    count_launch();
  }
  PODSDF_CUDA(cudaGetLastError());
  return 0;
}

int launch_df_plane(const DfGeom& g, const Taps& by, const Taps& bz, const double* T1, double* out, int nplanes,
                    int ncomp, const double* coef, unsigned mask, bool lund, double* T2, bool tuned,
                    cudaStream_t st) {
  if (nplanes <= 0) return 0;
  if (lund && ncomp != 3) return fail(PODSDF_ERR_INVALID, "Lund transform needs 3 components");
  const int nty = 2 * g.nfy + 1, ntz = 2 * g.nfz + 1;
  if (tuned) {
    if (ncomp != 3) return fail(PODSDF_ERR_UNSUPPORTED, "tuned plane kernel handles 3 components");
#define F(A, B) \
  if (nty == A && ntz == B) return launch_plane_tuned<A, B>(g, by, bz, T1, out, nplanes, coef, mask, lund, st);
    PODSDF_PLANE_TAPS(F)
#undef F
    return fail(PODSDF_ERR_UNSUPPORTED, "no tuned plane-filter kernel for this (nfy, nfz)");
  }
  if (!T2) return fail(PODSDF_ERR_INVALID, "generic plane filter needs the T2 workspace");
  const long nrows = (long)nplanes * ncomp * g.Jp;
  df_zpass_generic_kernel<<<ceil_div(nrows * g.K, 256), 256, 0, st>>>(T1, T2, bz, ntz, g.Jp, g.Kp, g.K, nrows);
  count_launch();
  for (int done = 0; done < nplanes; done += 65535) {
    const int cnt = (nplanes - done < 65535) ? nplanes - done : 65535;
    dim3 grid(ceil_div(g.P, 256), cnt);
    df_ypass_generic_kernel<<<grid, 256, 0, st>>>(T2 + (long)done * ncomp * g.Jp * g.K,
                                                  out + (long)done * ncomp * g.P, by, nty, g, ncomp, coef, mask,
                                                  lund ? 1 : 0);
    count_launch();
  }
  PODSDF_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace podsdf
