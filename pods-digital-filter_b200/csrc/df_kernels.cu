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
//   K-plane: out[t] = Lund( by (*) bz (*) T1[t] )   marching along the columns of a 64-row strip:
//            bulk-copied (cp.async.bulk + mbarrier) 16-column panels of T1, y-pass from the panel
//            into a shared-memory ring of y-filtered columns, z-pass as register-blocked windows
//            over that ring, fused with the 3x3 Lund/rotation transform and the mean add, so every
//            output plane is written once, in 64-byte runs per thread.
// Both are FP64-FMA work at the DFMA pipe rate with HBM traffic = algorithmic bytes; no GEMM shape.
#include "df_kernels.cuh"
#include "philox.cuh"

#include <type_traits>

namespace podsdf {


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

// y <- R y, written like numpy's R.dot(V): products summed left to right (NaN / inf propagate as 0*NaN = NaN)
__device__ __forceinline__ void rotate_point(const Rot& rot, double& y0, double& y1, double& y2) {
  const double a = y0, b = y1, c = y2;
  y0 = fma(rot.r[2], c, fma(rot.r[1], b, rot.r[0] * a));
  y1 = fma(rot.r[5], c, fma(rot.r[4], b, rot.r[3] * a));
  y2 = fma(rot.r[8], c, fma(rot.r[7], b, rot.r[6] * a));
}

__global__ void __launch_bounds__(256)
df_lund_kernel(double* __restrict__ planes, const double* __restrict__ coef, const unsigned mask, const long P,
               const int lund, const Rot rot) {
  const long pt = blockIdx.x * 256L + threadIdx.x;
  if (pt >= P) return;
  double y[3] = {planes[pt], planes[P + pt], planes[2 * P + pt]};
  if (lund) lund_point(coef, mask, P, pt, planes[pt], planes[P + pt], planes[2 * P + pt], y);
  if (rot.on) rotate_point(rot, y[0], y[1], y[2]);
  planes[pt] = y[0];
  planes[P + pt] = y[1];
  planes[2 * P + pt] = y[2];
}

int launch_df_lund(const DfGeom& g, double* planes, const double* coef, unsigned mask, bool lund, const Rot& rot,
                   cudaStream_t st) {
  df_lund_kernel<<<ceil_div(g.P, 256), 256, 0, st>>>(planes, coef, mask, g.P, lund ? 1 : 0, rot);
  count_launch();
  PODSDF_CUDA(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------------
// K-plane, tuned: row-marching strip kernel over the row-panel layout
//   CTA = (64-column strip, plane): 4 compute warps + 1 producer warp, marches down the padded rows one
//   16-row panel per step:
//     stage  : ONE cp.async.bulk per component ((64 + NTZ-1) columns x 16 rows, contiguous in the panel
//              layout) completing on an mbarrier, issued by the producer warp
//     z-pass : thread = (row, 8-column segment) x 3 components: 8 outputs from a (8+NTZ-1)-value register
//              window along the staged columns (lanes <-> the 16 contiguous rows of a column: conflict
//              free), written to a shared ring of z-filtered rows (NB blocks of 16 rows, odd pitch)
//     y-pass : thread = (column, 8-row segment) x 3 components: register windows down the ring column
//              (lanes <-> consecutive columns: conflict free); block/row offsets are compile-time
//     Lund   : y = M x + m on the thread's 8 output points; lanes <-> consecutive k, so coefficient loads
//              and the streaming stores of the finished rows are whole 128-byte lines; for the common
//              sparsity pattern (PAT 1) the coefficient loads are issued before the y-pass
// ------------------------------------------------------------------------------------------------
template <int NTY, int NTZ>
struct PanelCfg {
  static constexpr int TK = 64;     // output columns per strip
  static constexpr int SJ = 16;     // rows per marching step (= panel height)
  static constexpr int R = 8;       // outputs per thread along the filtered axis (register window)
  static constexpr int NCOMPUTE = 128;
  static constexpr int NTHR = NCOMPUTE + 32;                 // + producer warp
  static constexpr int TKIN = TK + NTZ - 1;                  // staged columns
  static constexpr int NB = (SJ + NTY - 1 + SJ - 1) / SJ;    // ring blocks of SJ rows
  static constexpr int RJ = NB * SJ;
  static constexpr int PK = TK + 1;                          // ring pitch (odd)
  static constexpr int STAGE = 3 * TKIN * SJ;                // doubles
  static constexpr int RING = 3 * RJ * PK;
  static constexpr size_t SMEM = (size_t)(STAGE + RING) * sizeof(double) + 16;
  static constexpr unsigned COPY_BYTES = TKIN * SJ * sizeof(double);
  static_assert(SJ * (TK / R) == NCOMPUTE, "z-pass: one item per thread");
  static_assert(TK * (SJ / R) == NCOMPUTE, "y-pass: one item per thread");
  static constexpr bool kCoupledFits = SMEM <= 113 * 1024;   // two CTAs of the 3-component kernel per SM
  static_assert(COPY_BYTES % 16 == 0, "bulk copy size");
};

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

// y-pass of one thread: column kk, output rows j0 .. j0+R-1 (j0 = n*SJ - (NTY-1) + SG*R), 3 comps.
// blk[d] = ring base (component 0, column kk) of the block written d steps ago.
template <int NTY, int NTZ, int SG>
__device__ __forceinline__ void panel_ypass(const double* const (&blk)[PanelCfg<NTY, NTZ>::NB], const Taps& by,
                                            double (&f)[3][PanelCfg<NTY, NTZ>::R]) {
  using C = PanelCfg<NTY, NTZ>;
  constexpr int R = C::R, SJ = C::SJ, PK = C::PK, RJ = C::RJ;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
#pragma unroll
    for (int i = 0; i < NTY + R - 1; ++i) {
      constexpr int D = NTY - 1;
      const int off = i + SG * R - D;                       // row relative to n*SJ (compile time)
      const int d = (off >= 0) ? 0 : (-off + SJ - 1) / SJ;  // blocks back
      const int row = off + d * SJ;                         // row inside that block
      const double v = blk[d][(c * RJ + row) * PK];
#pragma unroll
      for (int o = 0; o < R; ++o) {
        const int m = i - o;
        if (m == 0) f[c][o] = by.b[NTY - 1] * v;
        else if (m > 0 && m < NTY) f[c][o] = fma(by.b[NTY - 1 - m], v, f[c][o]);
      }
    }
  }
}

// coefficient patterns with a specialised epilogue
//   PAT 0: no transform (identity), PAT 1: mask within {M00, M11, M20, M22, m0} -- adapt1d / adapt2d
//   without a rotation (digitalfilters.py:143-178, 233-485), PAT 2: anything (full 3x3 + means)
constexpr unsigned kPat1Mask = (1u << 0) | (1u << 4) | (1u << 6) | (1u << 8) | (1u << 9);
constexpr unsigned kOffDiagMask = (1u << 1) | (1u << 2) | (1u << 3) | (1u << 5) | (1u << 6) | (1u << 7);

template <int NTY, int NTZ, int PAT>
__global__ void __launch_bounds__(160, 2)
df_plane_panel_kernel(const double* __restrict__ T1, double* __restrict__ out, const Taps by, const Taps bz,
                      const DfGeom g, const PanelGeom panel, const double* __restrict__ coef, const unsigned mask,
                      const Rot rot) {
  using C = PanelCfg<NTY, NTZ>;
  static_assert(C::kCoupledFits, "two CTAs per SM");
  constexpr int TK = C::TK, SJ = C::SJ, R = C::R, TKIN = C::TKIN, RJ = C::RJ, PK = C::PK, NB = C::NB;
  extern __shared__ __align__(128) double smem[];
  double* stage = smem;                      // [3][TKIN][SJ]
  double* ring = smem + C::STAGE;            // [3][RJ][PK]
  const unsigned mbar = (unsigned)__cvta_generic_to_shared(smem + C::STAGE + C::RING);

  const int tid = threadIdx.x;
  const int k0 = blockIdx.x * TK;
  const long plane = blockIdx.y;
  const int nsteps = panel.PB;

  if (tid == 0) mbar_init(mbar, 1);
  __syncthreads();

  if (tid >= C::NCOMPUTE) {
    // ---- producer warp: one lane issues the three bulk copies of every step and follows the barriers
    const unsigned stage_s = (unsigned)__cvta_generic_to_shared(stage);
    const double* src = T1 + plane * 3 * panel.comp_stride + (long)k0 * SJ;   // panel 0, first staged column
    const long panel_stride = (long)panel.Qpad * SJ;
    auto issue = [&](int n) {
      if (tid == C::NCOMPUTE) {
        mbar_expect_tx(mbar, 3 * C::COPY_BYTES);
#pragma unroll
        for (int c = 0; c < 3; ++c)
          bulk_g2s(stage_s + c * C::COPY_BYTES, src + c * panel.comp_stride + n * panel_stride, C::COPY_BYTES, mbar);
      }
    };
    issue(0);
    for (int n = 0; n < nsteps; ++n) {
      __syncthreads();                // (A) the compute warps have consumed stage n
      if (n + 1 < nsteps) {
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");   // generic reads before async writes
        issue(n + 1);
      }
      __syncthreads();                // (B)
    }
    return;
  }

  double* dstp = out + plane * 3 * g.P;
  // z-pass role: panel row zr, column segment zs
  const int zr = tid & (SJ - 1), zs = tid / SJ;
  const double* zsrc = stage + (zs * R) * SJ + zr;
  double* zdst = ring + zr * PK + zs * R;
  // y-pass role: strip column kk, row segment sg
  const int kk = tid & (TK - 1), sg = tid / TK;
  const int k = k0 + kk;
  const bool kok = k < g.K;

  for (int n = 0; n < nsteps; ++n) {
    mbar_wait(mbar, n & 1);           // panel n has landed
    const int bn = n % NB;
    {
      double acc[3][R];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const double* col = zsrc + c * TKIN * SJ;
#pragma unroll
        for (int i = 0; i < NTZ + R - 1; ++i) {
          const double v = col[i * SJ];
#pragma unroll
          for (int o = 0; o < R; ++o) {
            const int m = i - o;
            if (m == 0) acc[c][o] = bz.b[NTZ - 1] * v;
            else if (m > 0 && m < NTZ) acc[c][o] = fma(bz.b[NTZ - 1 - m], v, acc[c][o]);
          }
        }
      }
#pragma unroll
      for (int c = 0; c < 3; ++c) {   // stores after all loads: no load is ordered behind a store
        double* dst = zdst + (c * RJ + bn * SJ) * PK;
#pragma unroll
        for (int o = 0; o < R; ++o) dst[o] = acc[c][o];
      }
    }
    __syncthreads();                  // (A) ring block n complete; the stage is free for panel n+1

    const int j0 = n * SJ - (NTY - 1) + sg * R;
    const bool item = (j0 + R > 0) && (j0 < g.J);
    const bool fast = item && kok && (j0 >= 0) && (j0 + R <= g.J);
    const int pt0 = j0 * g.K + k;
    double cf[5][R];
    if (PAT == 1) {                   // coefficient loads in flight during the y-pass
      constexpr int bit[5] = {0, 4, 6, 8, 9};
#pragma unroll
      for (int q = 0; q < 5; ++q) {
        if (fast && (mask & (1u << bit[q]))) {
          const double* cp = coef + (long)bit[q] * g.P + pt0;
#pragma unroll
          for (int o = 0; o < R; ++o) cf[q][o] = __ldg(cp + (long)o * g.K);
        } else {
#pragma unroll
          for (int o = 0; o < R; ++o) cf[q][o] = 0.0;
        }
      }
    }
    if (item) {
      const double* blk[NB];
#pragma unroll
      for (int d = 0; d < NB; ++d) blk[d] = ring + ((n + NB - d) % NB) * SJ * PK + kk;
      double f[3][R];
      if (sg == 0)
        panel_ypass<NTY, NTZ, 0>(blk, by, f);
      else
        panel_ypass<NTY, NTZ, 1>(blk, by, f);
      if (fast && PAT <= 1) {
#pragma unroll
        for (int o = 0; o < R; ++o) {
          double y0 = f[0][o], y1 = f[1][o], y2 = f[2][o];
          if (PAT == 1) {
            y0 = fma(cf[0][o], f[0][o], cf[4][o]);
            y1 = cf[1][o] * f[1][o];
            y2 = fma(cf[3][o], f[2][o], cf[2][o] * f[0][o]);
          }
          if (rot.on) rotate_point(rot, y0, y1, y2);
          double* op = dstp + pt0 + (long)o * g.K;
          __stcs(op, y0);
          __stcs(op + g.P, y1);
          __stcs(op + 2 * g.P, y2);
        }
      } else if (kok) {               // general pattern, or a strip edge: guarded
#pragma unroll
        for (int o = 0; o < R; ++o) {
          const int j = j0 + o;
          if (j >= 0 && j < g.J) {
            const int pt = pt0 + o * g.K;
            double y[3] = {f[0][o], f[1][o], f[2][o]};
            if (PAT != 0) lund_point(coef, mask, g.P, pt, f[0][o], f[1][o], f[2][o], y);
            if (rot.on) rotate_point(rot, y[0], y[1], y[2]);
#pragma unroll
            for (int c = 0; c < 3; ++c) __stcs(dstp + (long)c * g.P + pt, y[c]);
          }
        }
      }
    }
    __syncthreads();                  // (B) y-pass done with the ring block the next z-pass overwrites
  }
}

// ------------------------------------------------------------------------------------------------
// K-plane, tuned, DIAGONAL transform: when M is diagonal (no shear stress, no rotation -- the default
// `build_profile` inlet, digitalfilters.py:1038-1062 with the default plane normal) the three components
// never meet, so each (strip, plane, component) is its own CTA: a third of the shared memory (36 KB),
// 4-5 resident CTAs per SM instead of 2, i.e. twice the warps to hide LDS latency and barriers.
// Same stage / z-pass / y-pass structure as df_plane_panel_kernel.
// ------------------------------------------------------------------------------------------------
template <int NTY, int NTZ>
struct Panel1Cfg {
  using B = PanelCfg<NTY, NTZ>;
  static constexpr int STAGE = B::TKIN * B::SJ;
  static constexpr int RING = B::RJ * B::PK;
  static constexpr size_t SMEM = (size_t)(STAGE + RING) * sizeof(double) + 16;
};

template <int NTY, int NTZ>
__global__ void __launch_bounds__(128, 6)
df_plane_panel1_kernel(const double* __restrict__ T1, double* __restrict__ out, const Taps by, const Taps bz,
                       const DfGeom g, const PanelGeom panel, const double* __restrict__ coef, const unsigned mask,
                       const int lund) {
  using C = PanelCfg<NTY, NTZ>;
  using C1 = Panel1Cfg<NTY, NTZ>;
  constexpr int TK = C::TK, SJ = C::SJ, R = C::R, RJ = C::RJ, PK = C::PK, NB = C::NB;
  extern __shared__ __align__(128) double smem[];
  double* stage = smem;                       // [TKIN][SJ]
  double* ring = smem + C1::STAGE;            // [RJ][PK]
  const unsigned mbar = (unsigned)__cvta_generic_to_shared(smem + C1::STAGE + C1::RING);

  const int tid = threadIdx.x;
  const int k0 = blockIdx.x * TK;
  const long pc = blockIdx.y;                 // plane * 3 + component
  const int comp = (int)(pc % 3);
  const int nsteps = panel.PB;

  // no producer warp here (it would cost every CTA a fifth of its registers): thread 0 issues the one bulk copy
  const unsigned stage_s = (unsigned)__cvta_generic_to_shared(stage);
  const double* src = T1 + pc * panel.comp_stride + (long)k0 * SJ;
  const long panel_stride = (long)panel.Qpad * SJ;
  if (tid == 0) {
    mbar_init(mbar, 1);
    mbar_expect_tx(mbar, C::COPY_BYTES);
    bulk_g2s(stage_s, src, C::COPY_BYTES, mbar);
  }
  __syncthreads();

  const int zr = tid & (SJ - 1), zs = tid / SJ;
  const double* zsrc = stage + (zs * R) * SJ + zr;
  double* zdst = ring + zr * PK + zs * R;
  const int kk = tid & (TK - 1), sg = tid / TK;
  const int k = k0 + kk;
  const bool kok = k < g.K;
  const bool has_m = lund && (mask & (1u << (4 * comp)));       // M_cc plane present
  const bool has_b = lund && (mask & (1u << (9 + comp)));       // mean plane present
  const double m_default = lund ? 0.0 : 1.0;                    // absent M_cc plane is an exact zero
  // per-thread walking pointers: point (j0, k) of this step, advanced by SJ rows per step
  // 32-bit point indices (J*K < 2^29 is checked at plan creation): one IMAD.WIDE per address
  const int rowK = g.K, stepK = SJ * g.K;
  int pt = (sg * R - (NTY - 1)) * rowK + k;
  const double* cm = coef + (long)(4 * comp) * g.P;
  const double* cb = coef + (long)(9 + comp) * g.P;
  double* dstp = out + pc * g.P;
  const double* ringk = ring + kk;
  int bn = 0;                                                   // ring block written this step = n % NB
  int j0 = sg * R - (NTY - 1);

  for (int n = 0; n < nsteps; ++n, pt += stepK, j0 += SJ, bn = (bn + 1 == NB) ? 0 : bn + 1) {
    mbar_wait(mbar, n & 1);
    {
      double acc[R];
#pragma unroll
      for (int i = 0; i < NTZ + R - 1; ++i) {
        const double v = zsrc[i * SJ];
#pragma unroll
        for (int o = 0; o < R; ++o) {
          const int m = i - o;
          if (m == 0) acc[o] = bz.b[NTZ - 1] * v;
          else if (m > 0 && m < NTZ) acc[o] = fma(bz.b[NTZ - 1 - m], v, acc[o]);
        }
      }
      double* dst = zdst + (bn * SJ) * PK;
#pragma unroll
      for (int o = 0; o < R; ++o) dst[o] = acc[o];
    }
    __syncthreads();
    if (tid == 0 && n + 1 < nsteps) {
      asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");   // generic reads before async writes
      mbar_expect_tx(mbar, C::COPY_BYTES);
      bulk_g2s(stage_s, src + (n + 1) * panel_stride, C::COPY_BYTES, mbar);
    }

    if (kok && (j0 + R > 0) && (j0 < g.J)) {
      const bool fast = (j0 >= 0) && (j0 + R <= g.J);
      double cfm[R], cfb[R];
      if (fast) {                     // coefficient loads in flight during the y-pass
        const double* pm = cm + pt;
        const double* pb = cb + pt;
#pragma unroll
        for (int o = 0; o < R; ++o) {
          cfm[o] = has_m ? __ldg(pm) : m_default;
          cfb[o] = has_b ? __ldg(pb) : 0.0;
          pm += rowK;
          pb += rowK;
        }
      } else {
#pragma unroll
        for (int o = 0; o < R; ++o) {
          const bool ok = (j0 + o >= 0) && (j0 + o < g.J);
          cfm[o] = (ok && has_m) ? __ldg(cm + pt + (long)o * rowK) : m_default;
          cfb[o] = (ok && has_b) ? __ldg(cb + pt + (long)o * rowK) : 0.0;
        }
      }
      double f[R];
      const double* blk[NB];
      int b = bn;
#pragma unroll
      for (int d = 0; d < NB; ++d) {
        blk[d] = ringk + b * (SJ * PK);
        b = (b == 0) ? NB - 1 : b - 1;
      }
      auto ypass = [&](auto SGc) {
        constexpr int SG = decltype(SGc)::value;
#pragma unroll
        for (int i = 0; i < NTY + R - 1; ++i) {
          constexpr int D = NTY - 1;
          const int off = i + SG * R - D;
          const int d = (off >= 0) ? 0 : (-off + SJ - 1) / SJ;
          const int row = off + d * SJ;
          const double v = blk[d][row * PK];
#pragma unroll
          for (int o = 0; o < R; ++o) {
            const int m = i - o;
            if (m == 0) f[o] = by.b[NTY - 1] * v;
            else if (m > 0 && m < NTY) f[o] = fma(by.b[NTY - 1 - m], v, f[o]);
          }
        }
      };
      if (sg == 0) ypass(std::integral_constant<int, 0>());
      else ypass(std::integral_constant<int, 1>());
      double* op = dstp + pt;
      if (fast) {
#pragma unroll
        for (int o = 0; o < R; ++o) {
          __stcs(op, fma(cfm[o], f[o], cfb[o]));
          op += rowK;
        }
      } else {
#pragma unroll
        for (int o = 0; o < R; ++o)
          if ((j0 + o >= 0) && (j0 + o < g.J)) __stcs(op + o * rowK, fma(cfm[o], f[o], cfb[o]));
      }
    }
    __syncthreads();
  }
}

template <int NTY, int NTZ>
static int launch_plane_panel1(const DfGeom& g, const PanelGeom& panel, const Taps& by, const Taps& bz, const double* T1,
                               double* out, int nplanes, const double* coef, unsigned mask, bool lund, cudaStream_t st) {
  using C = PanelCfg<NTY, NTZ>;
  using C1 = Panel1Cfg<NTY, NTZ>;
  auto kern = df_plane_panel1_kernel<NTY, NTZ>;
  static bool configured[64] = {};
  int dev = 0;
  PODSDF_CUDA(cudaGetDevice(&dev));
  if (!configured[dev & 63]) {
    PODSDF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C1::SMEM));
    configured[dev & 63] = true;
  }
  for (int done = 0; done < nplanes; done += 21845) {   // grid.y = 3 * planes <= 65535
    const int cnt = (nplanes - done < 21845) ? nplanes - done : 21845;
    dim3 grid(ceil_div(g.K, C::TK), 3 * cnt);
    kern<<<grid, C::NCOMPUTE, C1::SMEM, st>>>(T1 + (long)done * 3 * panel.comp_stride, out + (long)done * 3 * g.P, by, bz,
                                              g, panel, coef, mask, lund ? 1 : 0);
    count_launch();
  }
  PODSDF_CUDA(cudaGetLastError());
  return 0;
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
                        const int lund, const Rot rot) {
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
  if (lund || rot.on) {
    double y[3] = {f[0], f[1], f[2]};
    if (lund) lund_point(coef, mask, g.P, pt, f[0], f[1], f[2], y);
    if (rot.on) rotate_point(rot, y[0], y[1], y[2]);
    for (int c = 0; c < 3; ++c) out[(plane * 3 + c) * g.P + pt] = y[c];
  } else {
    for (int c = 0; c < ncomp; ++c) out[(plane * ncomp + c) * g.P + pt] = f[c];
  }
}

// (NTY, NTZ) pairs with a tuned kernel: the isotropic sizes of configs 1, 3, 5 and config 2's (8, 4)
#define PODSDF_PLANE_TAPS(F) F(9, 9) F(13, 13) F(17, 17) F(21, 21) F(25, 25) F(33, 33) F(17, 9)
// wider filters (config 5: nf = 24 .. 64) have a tuned kernel for diagonal transforms only: three rings of
// (16 + 2 nfy) rows do not fit one SM's shared memory
#define PODSDF_PLANE_TAPS_WIDE(F) F(49, 49) F(65, 65) F(97, 97) F(129, 129)

bool df_plane_has_tuned(int nty, int ntz, bool diagonal) {
#define F(A, B) if (nty == A && ntz == B) return true;
  PODSDF_PLANE_TAPS(F)
  if (diagonal) { PODSDF_PLANE_TAPS_WIDE(F) }
#undef F
  return false;
}
bool df_transform_is_diagonal(bool lund, unsigned mask, const Rot& rot) {
  return !rot.on && (!lund || (mask & kOffDiagMask) == 0);
}

PanelGeom make_panel_geom(const DfGeom& g) {
  PanelGeom p;
  p.PB = (g.Jp + 15) / 16;
  p.Qpad = ((g.K + 63) / 64) * 64 + 2 * g.nfz;
  p.comp_stride = (long)p.PB * p.Qpad * 16;
  return p;
}

template <int NTY, int NTZ, int PAT>
static int launch_plane_panel(const DfGeom& g, const PanelGeom& panel, const Taps& by, const Taps& bz, const double* T1,
                              double* out, int nplanes, const double* coef, unsigned mask, const Rot& rot,
                              cudaStream_t st) {
  using C = PanelCfg<NTY, NTZ>;
  auto kern = df_plane_panel_kernel<NTY, NTZ, PAT>;
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
    kern<<<grid, C::NTHR, C::SMEM, st>>>(T1 + (long)done * 3 * panel.comp_stride, out + (long)done * 3 * g.P, by, bz, g,
                                         panel, coef, mask, rot);
    count_launch();
  }
  PODSDF_CUDA(cudaGetLastError());
  return 0;
}

int launch_df_plane(const DfGeom& g, const PanelGeom& panel, const Taps& by, const Taps& bz, const double* T1,
                    double* out, int nplanes, int ncomp, const double* coef, unsigned mask, bool lund, const Rot& rot,
                    double* T2, bool tuned, cudaStream_t st) {
  if (nplanes <= 0) return 0;
  if ((lund || rot.on) && ncomp != 3) return fail(PODSDF_ERR_INVALID, "Lund transform / rotation need 3 components");
  const int nty = 2 * g.nfy + 1, ntz = 2 * g.nfz + 1;
  if (tuned) {
    if (ncomp != 3 || panel.PB == 0) return fail(PODSDF_ERR_UNSUPPORTED, "tuned plane kernel: 3 components, panel layout");
    const int pat = !lund ? 0 : ((mask & ~kPat1Mask) == 0 ? 1 : 2);
    // off-diagonal planes of M absent: the components are independent -> one CTA per component
    const bool diagonal = df_transform_is_diagonal(lund, mask, rot);
#define F(A, B) \
  if (diagonal && nty == A && ntz == B) return launch_plane_panel1<A, B>(g, panel, by, bz, T1, out, nplanes, coef, mask, lund, st);
    PODSDF_PLANE_TAPS(F)
    PODSDF_PLANE_TAPS_WIDE(F)
#undef F
#define F(A, B)                                                                                            \
  if (nty == A && ntz == B) {                                                                              \
    if (pat == 0) return launch_plane_panel<A, B, 0>(g, panel, by, bz, T1, out, nplanes, coef, mask, rot, st);  \
    if (pat == 1) return launch_plane_panel<A, B, 1>(g, panel, by, bz, T1, out, nplanes, coef, mask, rot, st);  \
    return launch_plane_panel<A, B, 2>(g, panel, by, bz, T1, out, nplanes, coef, mask, rot, st);                \
  }
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
                                                  lund ? 1 : 0, rot);
    count_launch();
  }
  PODSDF_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace podsdf

