// K-time of the digital filter for sm_100a:  T1[t] = sum_m bx[m] X[t + m]  per padded point, plus the Philox slab
// materialiser.  Reference behaviour being replaced (file:line under /root/reference):
//   digitalfilters.py:124        the x (time) convolution of filter3DSciPy1D
//   digitalfilters.py:1361-1366, 1454-1467  random block, np.roll + one new random slab per component per step
// The reference redoes the whole time contraction every step; here each raw random is touched once: a sliding
// window per point lives in registers (DFMA kernels) or in 8-slab chunks feeding the FP64 tensor instruction.
// SRC selects the random source: 0 = injected slabs X, 1 = Philox U53 stream, 2 = Philox U32 stream (philox.cuh).
#include "df_kernels.cuh"
#include "philox.cuh"

namespace podsdf {

constexpr int kSrcInjected = 0, kSrcU53 = 1, kSrcU32 = 2;

// ------------------------------------------------------------------------------------------------
// Philox slabs
// ------------------------------------------------------------------------------------------------
template <int KIND>
__global__ void __launch_bounds__(256) df_random_slabs_kernel(double* __restrict__ X, long Np, long npair,
                                                              const PhiloxKeys keys, long s0) {
  const long pair = blockIdx.x * 256L + threadIdx.x;
  if (pair >= npair) return;
  const int c = blockIdx.y;
  const long s = blockIdx.z;
  double a, b;
  if (KIND == kRngU32) {       // (the time filters take four slabs per call; this materialiser is not a hot path)
    a = philox_value<kRngU32>(keys, (uint64_t)(2 * pair), (uint64_t)(s0 + s), c);
    b = philox_value<kRngU32>(keys, (uint64_t)(2 * pair + 1), (uint64_t)(s0 + s), c);
  } else {
    philox_pair(keys, (uint64_t)pair, (uint64_t)(s0 + s), c, a, b);
  }
  double* dst = X + (s * 3 + c) * Np + 2 * pair;
  dst[0] = a;
  if (2 * pair + 1 < Np) dst[1] = b;
}

int launch_df_random_slabs(const DfGeom& g, uint64_t seed, int kind, long s0, long n, double* X, cudaStream_t st) {
  if (n <= 0) return 0;
  const long npair = (g.Np + 1) / 2;
  for (long done = 0; done < n; done += 32768) {  // gridDim.z limit
    const long cnt = (n - done < 32768) ? n - done : 32768;
    dim3 grid(ceil_div(npair, 256), 3, (unsigned)cnt);
    if (kind == kRngU32)
      df_random_slabs_kernel<kRngU32><<<grid, 256, 0, st>>>(X + done * 3 * g.Np, g.Np, npair, philox_keys(seed), s0 + done);
    else
      df_random_slabs_kernel<kRngU53><<<grid, 256, 0, st>>>(X + done * 3 * g.Np, g.Np, npair, philox_keys(seed), s0 + done);
    count_launch();
  }
  PODSDF_CUDA(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------------
// K-time, tuned: window of NT raw randoms per point in registers (static indices via unroll by 2 NT)
// ------------------------------------------------------------------------------------------------
// T1 is written either in the canonical plane layout [p][q] (panel.PB == 0) or in the row-panel layout
// [pb][q][16] the tuned plane kernel bulk-copies from (PanelGeom); thread L owns the two points (p, q), (p, q+1)
// either way (one Philox call / one 16-byte read), and a warp's writes are whole 128-byte lines in both.
// Steps are taken in pairs (two Philox calls / four loads issued ahead of the 4 NT DFMA of the two steps).
// Serves the U53 stream and injected slabs in the canonical layout; the U32 stream runs in df_time_ring8_kernel.
template <int NT, int SRC>
__global__ void __launch_bounds__(256, (NT <= 21 ? 2 : 1))
df_time_tuned_kernel(const double* __restrict__ X, double* __restrict__ T1, const Taps bx, const DfGeom g,
                     const PanelGeom panel, const long nthread, const int nsteps, const int tchunk,
                     const PhiloxKeys keys, const long s_base) {
  static_assert(NT % 2 == 1, "2 nf + 1 taps");
  const long L = blockIdx.x * 256L + threadIdx.x;
  if (L >= nthread) return;
  const int c = blockIdx.y;
  const int t0 = blockIdx.z * tchunk;
  const int t1 = min(t0 + tchunk, nsteps);
  const long Np = g.Np;
  long ridx, widx, wstride;    // canonical point index (reads / Philox), write offset, write plane stride
  bool valid, has2, vec;
  const bool paneled = panel.PB > 0;
  if (panel.PB > 0) {          // panel layout: L = (pb * QP + qp) * 16 + r, two columns q, q+1 of plane row p
    const int r = (int)(L & 15);
    const long colp = L >> 4;
    const int QP = panel.Qpad >> 1;
    const int pb = (int)(colp / QP), q = 2 * (int)(colp % QP);
    const int p = pb * 16 + r;
    valid = (p < g.Jp) && (q < g.Kp);   // Kp is even here: both points of the pair are valid together
    has2 = valid;
    vec = true;
    ridx = (long)p * g.Kp + q;
    widx = ((long)pb * panel.Qpad + q) * 16 + r;
    wstride = panel.comp_stride;
  } else {
    valid = true;
    has2 = 2 * L + 1 < Np;
    vec = (Np & 1) == 0;       // every [s][c] plane stays 16-byte aligned
    ridx = widx = 2 * L;
    wstride = Np;
  }
  const int last_slab = nsteps + NT - 2;
  if (!valid) {                // panel padding: zeros, so the plane kernel never sees garbage
    for (int t = t0; t < t1; ++t) {
      double* q = T1 + ((long)t * 3 + c) * wstride + widx;
      q[0] = 0.0;
      q[16] = 0.0;
    }
    return;
  }
  static_assert(SRC != kSrcU32, "the U32 stream runs in df_time_ring8_kernel");
  const int E = t0;

  double w0[NT], w1[NT];
  // slabs s and s + 1 of this thread's two points: (a0, b0) and (a1, b1)
  auto load2 = [&](int s, double& a0, double& b0, double& a1, double& b1) {
    if (SRC == kSrcU53) {
      philox_pair(keys, (uint64_t)(ridx >> 1), (uint64_t)(s_base + s), c, a0, b0);
      philox_pair(keys, (uint64_t)(ridx >> 1), (uint64_t)(s_base + s + 1), c, a1, b1);
    } else {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const double* p = X + ((long)min(s + h, last_slab) * 3 + c) * Np + ridx;
        double a, b;
        if (paneled) {         // lanes <-> rows: the rest of each 128-byte line is read by the next warps -> keep in L1
          const double2 v = __ldg(reinterpret_cast<const double2*>(p));
          a = v.x;
          b = v.y;
        } else if (vec) {
          const double2 v = __ldcs(reinterpret_cast<const double2*>(p));  // read once: streaming
          a = v.x;
          b = v.y;
        } else {
          a = __ldcs(p);
          b = has2 ? __ldcs(p + 1) : 0.0;
        }
        if (h == 0) { a0 = a; b0 = b; } else { a1 = a; b1 = b; }
      }
    }
  };
#pragma unroll
  for (int i = 0; i < NT - 1; i += 2) load2(E + i, w0[i], w1[i], w0[i + 1], w1[i + 1]);

  // one output step: taps against the ring whose newest entry sits in slot (U + NT - 1) % NT
  // (U is a literal after unrolling, so every ring index is a register name)
  auto step = [&](const int U, const int tt) {
    // true convolution (scipy.signal.convolve flips the taps): tap m meets slab tt + NT-1-m
    // three independent partial sums per point: six DFMA chains per thread cover the DFMA latency
    double pa[3], pb[3];
#pragma unroll
    for (int m = 0; m < NT; ++m) {
      const int idx = (U + NT - 1 - m) % NT;
      if (m < 3) {
        pa[m] = bx.b[m] * w0[idx];
        pb[m] = bx.b[m] * w1[idx];
      } else {
        pa[m % 3] = fma(bx.b[m], w0[idx], pa[m % 3]);
        pb[m % 3] = fma(bx.b[m], w1[idx], pb[m % 3]);
      }
    }
    const double a0 = pa[0] + pa[1], a1 = pa[2], b0 = pb[0] + pb[1], b1 = pb[2];
    if (tt >= t0 && tt < t1) {
      double* q = T1 + ((long)tt * 3 + c) * wstride + widx;
      if (paneled) {           // lanes <-> 16 consecutive rows: two coalesced 128-byte lines per half-warp
        q[0] = a0 + a1;
        q[16] = b0 + b1;
      } else if (vec) {
        *reinterpret_cast<double2*>(q) = make_double2(a0 + a1, b0 + b1);
      } else {
        q[0] = a0 + a1;
        if (has2) q[1] = b0 + b1;
      }
    }
  };

  for (int t = E; t < t1; t += 2 * NT) {
#pragma unroll
    for (int u = 0; u < 2 * NT; u += 2) {
      const int tt = t + u;
      if (tt < t1) {                             // CTA-uniform
        // the two new slabs tt+NT-1, tt+NT: the first goes to slot (u+NT-1)%NT, the second would overwrite the
        // oldest slab (slot u%NT) step tt still needs -> parked in registers until step tt is done
        double na, nb;
        load2(tt + NT - 1, w0[(u + NT - 1) % NT], w1[(u + NT - 1) % NT], na, nb);
        step(u % NT, tt);
        w0[u % NT] = na;
        w1[u % NT] = nb;
        step((u + 1) % NT, tt + 1);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// K-time, U32 stream: ONE point per thread, steps in blocks of EIGHT.  A block costs two Philox calls (the U32
// stream yields four consecutive slabs of one point per call) + 8 NT DFMA in eight independent accumulator chains,
// written like the z-pass of the plane kernel: every window value is read once and feeds up to eight outputs.
// The window is a ring of R = 8 ceil((NT+7)/8) registers (the slabs one block touches); its pattern repeats after
// R/8 blocks, so the unrolled body is R steps of one point (about a fifth of the pair kernel's code, which ran
// 33 % of its cycles waiting for instruction fetches -- profiles/r02_ktime_ncu.md).  Lanes <-> consecutive panel
// entries: 256-byte stores per warp and plane.
// ------------------------------------------------------------------------------------------------
#ifndef PODSDF_R8_MINB
#define PODSDF_R8_MINB 2
#endif
template <int NT>
__global__ void __launch_bounds__(256, PODSDF_R8_MINB)
df_time_ring8_kernel(double* __restrict__ T1, const Taps bx, const DfGeom g, const PanelGeom panel, const long npoint,
                     const int nsteps, const int tchunk, const PhiloxKeys keys, const long s_base) {
  static_assert(NT % 4 == 1, "whole Philox quads in the NT - 1 warm-up slabs");
  constexpr int R = ((NT + 7 + 7) / 8) * 8;
  const long L = blockIdx.x * 256L + threadIdx.x;
  if (L >= npoint) return;
  const int c = blockIdx.y;
  const int t0 = blockIdx.z * tchunk;
  const int t1 = min(t0 + tchunk, nsteps);
  long ridx, wstride;
  bool valid = true;
  if (panel.PB > 0) {          // panel layout: L = (pb * Qpad + q) * 16 + r
    const int r = (int)(L & 15);
    const long col = L >> 4;
    const int pb = (int)(col / panel.Qpad), q = (int)(col % panel.Qpad);
    const int p = pb * 16 + r;
    valid = (p < g.Jp) && (q < g.Kp);
    ridx = (long)p * g.Kp + q;
    wstride = panel.comp_stride;
  } else {
    ridx = L;
    wstride = g.Np;
  }
  const long ostride = 3 * wstride;
  if (!valid) {
    double* o = T1 + (long)c * wstride + L + (long)t0 * ostride;
    for (int t = t0; t < t1; ++t, o += ostride) *o = 0.0;
    return;
  }
  // blocks start on a global slab that is a multiple of 4 (NT - 1 is one, so the slab entering at step tt is
  // congruent to s_base + tt): up to three extra leading steps are computed and not written
  const int E = t0 - (int)((s_base + t0) & 3);
  const uint64_t q_base = (uint64_t)((s_base + E) >> 2);       // quad index of slab E
  double* out = T1 + (long)c * wstride + L + (long)E * ostride; // walks one plane per step

  double w[R];
  auto load4 = [&](uint64_t quad, double& a0, double& a1, double& a2, double& a3) {
    uint32_t r[4];
    philox_quad(keys, (uint64_t)ridx, quad, c, r);
    a0 = philox_u32_to_uniform(r[0]);
    a1 = philox_u32_to_uniform(r[1]);
    a2 = philox_u32_to_uniform(r[2]);
    a3 = philox_u32_to_uniform(r[3]);
  };
#pragma unroll
  for (int i = 0; i < NT - 1; i += 4) load4(q_base + i / 4, w[i], w[i + 1], w[i + 2], w[i + 3]);
  uint64_t quad = q_base + (NT - 1) / 4;                       // next quad to draw

  for (int t = E; t < t1; t += R) {
#pragma unroll
    for (int bq = 0; bq < R; bq += 8) {
      const int tt = t + bq;
      if (tt < t1) {                                           // CTA-uniform
        load4(quad, w[(bq + NT - 1) % R], w[(bq + NT) % R], w[(bq + NT + 1) % R], w[(bq + NT + 2) % R]);
        load4(quad + 1, w[(bq + NT + 3) % R], w[(bq + NT + 4) % R], w[(bq + NT + 5) % R], w[(bq + NT + 6) % R]);
        quad += 2;
        // out[tt + o] = sum_m bx[m] X[tt + o + NT-1-m]  (true convolution: scipy.signal.convolve flips the taps)
        double acc[8];
#pragma unroll
        for (int i = 0; i < NT + 7; ++i) {
          const double v = w[(bq + i) % R];
#pragma unroll
          for (int o = 0; o < 8; ++o) {
            const int m = i - o;
            if (m == 0) acc[o] = bx.b[NT - 1] * v;
            else if (m > 0 && m < NT) acc[o] = fma(bx.b[NT - 1 - m], v, acc[o]);
          }
        }
#pragma unroll
        for (int o = 0; o < 8; ++o) {
          if (tt + o >= t0 && tt + o < t1) *out = acc[o];
          out += ostride;
        }
      }
    }
  }
}

template <int NT>
static void launch_time_ring8(const DfGeom& g, const PanelGeom& panel, const Taps& bx, double* T1, int nsteps,
                              uint64_t seed, long s_base, cudaStream_t st) {
  constexpr int R = ((NT + 7 + 7) / 8) * 8;
  const long npoint = panel.PB > 0 ? (long)panel.PB * panel.Qpad * 16 : g.Np;
  int nchunk = 1;
  const long ctas_per_chunk = 3L * ceil_div(npoint, 256);
  while (nchunk < 64 && ctas_per_chunk * nchunk < 148L * 2 * 8 && nsteps / (nchunk * 2) >= 8 * NT) nchunk *= 2;
  int tchunk = ceil_div(nsteps, nchunk);
  tchunk = ceil_div(tchunk, R) * R;              // whole unrolled blocks; a multiple of 8
  nchunk = ceil_div(nsteps, tchunk);
  dim3 grid(ceil_div(npoint, 256), 3, nchunk);
  df_time_ring8_kernel<NT><<<grid, 256, 0, st>>>(T1, bx, g, panel, npoint, nsteps, tchunk, philox_keys(seed), s_base);
  count_launch();
}

// K-time for wide filters (nfx = 24 .. 64): one point per thread (the window alone is NT doubles), and the
// tap vector is cut in passes of <= 65 taps -- T1 = sum of the passes, each the same sliding-window kernel on
// a shifted slab range (`s_off`), the later ones accumulating.  Config 5's wide end runs here.
template <int NT, int SRC>
__global__ void __launch_bounds__(256, 1)
df_time_wide_kernel(const double* __restrict__ X, double* __restrict__ T1, const Taps bx, const DfGeom g,
                    const PanelGeom panel, const long nthread, const int nsteps, const int tchunk,
                    const PhiloxKeys keys, const long s_base, const int s_off, const int last_slab,
                    const int accumulate) {
  const long L = blockIdx.x * 256L + threadIdx.x;
  if (L >= nthread) return;
  const int c = blockIdx.y;
  const int t0 = blockIdx.z * tchunk;
  const int t1 = min(t0 + tchunk, nsteps);
  const long Np = g.Np;
  long ridx, widx, wstride;
  bool valid = true;
  if (panel.PB > 0) {          // panel layout: L = (pb * Qpad + q) * 16 + r
    const int r = (int)(L & 15);
    const long col = L >> 4;
    const int pb = (int)(col / panel.Qpad), q = (int)(col % panel.Qpad);
    const int p = pb * 16 + r;
    valid = (p < g.Jp) && (q < g.Kp);
    ridx = (long)p * g.Kp + q;
    widx = L;
    wstride = panel.comp_stride;
  } else {
    ridx = widx = L;
    wstride = Np;
  }
  if (!valid) {
    if (!accumulate)
      for (int t = t0; t < t1; ++t) T1[((long)t * 3 + c) * wstride + widx] = 0.0;
    return;
  }
  auto load = [&](int s) -> double {
    s = min(s_off + s, last_slab);
    if (SRC == kSrcU32) return philox_value<kRngU32>(keys, (uint64_t)ridx, (uint64_t)(s_base + s), c);
    if (SRC == kSrcU53) return philox_value<kRngU53>(keys, (uint64_t)ridx, (uint64_t)(s_base + s), c);
    return __ldcs(X + ((long)s * 3 + c) * Np + ridx);
  };
  double w[NT];
#pragma unroll
  for (int i = 0; i < NT - 1; ++i) w[i] = load(t0 + i);
  for (int t = t0; t < t1; t += NT) {
#pragma unroll
    for (int u = 0; u < NT; ++u) {
      const int tt = t + u;
      w[(u + NT - 1) % NT] = load(tt + NT - 1);
      double pa[4];
#pragma unroll
      for (int m = 0; m < NT; ++m) {
        const int idx = (u + NT - 1 - m) % NT;
        if (m < 4) pa[m] = bx.b[m] * w[idx];
        else pa[m % 4] = fma(bx.b[m], w[idx], pa[m % 4]);
      }
      if (tt < t1) {
        double* q = T1 + ((long)tt * 3 + c) * wstride + widx;
        const double v = (pa[0] + pa[1]) + (pa[2] + pa[3]);
        *q = accumulate ? *q + v : v;
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
// wide filters: total tap count -> passes of 49 or 65 taps (the second pass front-padded with one zero tap)
static int wide_pass_taps(int ntx) { return ntx == 49 || ntx == 97 ? 49 : (ntx == 65 || ntx == 129 ? 65 : 0); }

bool df_time_has_tuned(int ntx) {
#define F(N) if (ntx == N) return true;
  PODSDF_TIME_TAPS(F)
#undef F
  return wide_pass_taps(ntx) != 0;
}

template <int NT>
static void launch_time_tuned(const DfGeom& g, const PanelGeom& panel, const Taps& bx, const double* X, double* T1,
                              int nsteps, int src, uint64_t seed, long s_base, cudaStream_t st) {
  const long nthread = panel.PB > 0 ? (long)panel.PB * (panel.Qpad / 2) * 16 : (g.Np + 1) / 2;
  // time chunks: enough CTAs to smooth the tail, few enough that the NT-1 slab warm-up stays small
  int nchunk = 1;
  const long ctas_per_chunk = 3L * ceil_div(nthread, 256);
  while (nchunk < 64 && ctas_per_chunk * nchunk < 148L * 2 * 8 && nsteps / (nchunk * 2) >= 8 * NT) nchunk *= 2;
  int tchunk = ceil_div(nsteps, nchunk);
  tchunk = ceil_div(tchunk, 2 * NT) * (2 * NT);     // whole unrolled blocks; even, so chunk starts keep the slab parity
  nchunk = ceil_div(nsteps, tchunk);
  dim3 grid(ceil_div(nthread, 256), 3, nchunk);
  const PhiloxKeys keys = philox_keys(seed);
  if (src == kSrcU32) {
    launch_time_ring8<NT>(g, panel, bx, T1, nsteps, seed, s_base, st);
    return;
  }
  if (src == kSrcU53)
    df_time_tuned_kernel<NT, kSrcU53><<<grid, 256, 0, st>>>(X, T1, bx, g, panel, nthread, nsteps, tchunk, keys, s_base);
  else
    df_time_tuned_kernel<NT, kSrcInjected><<<grid, 256, 0, st>>>(X, T1, bx, g, panel, nthread, nsteps, tchunk, keys, s_base);
  count_launch();
}

template <int NT>
static void launch_time_wide(const DfGeom& g, const PanelGeom& panel, const Taps& bx_total, int nt_total, const double* X,
                             double* T1, int nsteps, int src, uint64_t seed, long s_base, cudaStream_t st) {
  const long nthread = panel.PB > 0 ? (long)panel.PB * panel.Qpad * 16 : g.Np;
  int nchunk = 1;
  const long ctas_per_chunk = 3L * ceil_div(nthread, 256);
  while (nchunk < 64 && ctas_per_chunk * nchunk < 148L * 8 && nsteps / (nchunk * 2) >= 4 * NT) nchunk *= 2;
  int tchunk = ceil_div(nsteps, nchunk);
  tchunk = ceil_div(tchunk, NT) * NT;
  nchunk = ceil_div(nsteps, tchunk);
  dim3 grid(ceil_div(nthread, 256), 3, nchunk);
  const PhiloxKeys keys = philox_keys(seed);
  const int last_slab = nsteps + nt_total - 2;
  const int npass = (nt_total + NT - 1) / NT;      // 1 or 2
  for (int pass = 0; pass < npass; ++pass) {
    Taps b;
    for (int i = 0; i < kMaxTaps; ++i) b.b[i] = 0.0;
    int s_off = 0;
    if (pass == 0) {                               // taps [0, NT): newest slabs
      for (int i = 0; i < NT; ++i) b.b[i] = bx_total.b[i];
      s_off = nt_total - NT;
    } else {                                       // taps [NT, nt_total), front-padded with zeros to NT
      const int rest = nt_total - NT;
      for (int i = 0; i < rest; ++i) b.b[NT - rest + i] = bx_total.b[NT + i];
    }
    if (src == kSrcU32)
      df_time_wide_kernel<NT, kSrcU32><<<grid, 256, 0, st>>>(X, T1, b, g, panel, nthread, nsteps, tchunk, keys, s_base, s_off,
                                                              last_slab, pass);
    else if (src == kSrcU53)
      df_time_wide_kernel<NT, kSrcU53><<<grid, 256, 0, st>>>(X, T1, b, g, panel, nthread, nsteps, tchunk, keys, s_base, s_off,
                                                              last_slab, pass);
    else
      df_time_wide_kernel<NT, kSrcInjected><<<grid, 256, 0, st>>>(X, T1, b, g, panel, nthread, nsteps, tchunk, keys, s_base,
                                                                   s_off, last_slab, pass);
    count_launch();
  }
}

// ------------------------------------------------------------------------------------------------
// K-time on the FP64 tensor path (panel layout only).  The DFMA kernel above is bound by instruction issue
// (134 instructions per 64 outputs, Philox included); here the sliding-window sum is a banded Toeplitz
// matrix product  T1[t0+i][pt] = sum_s A[i][s] X[s][pt],  A[i][s] = bx[NT-1-(s-i)],  done with
// mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4): one instruction per 256 multiply-adds, 8 output steps x 8 points
// per tile.  The band wastes (NT+7 rounded up to 4)/NT of the tensor flops (25 % at NT = 21, 6 % at
// NT = 129) but the instruction count per output drops 1.7x and the window needs 8 doubles per tile and lane.
//   warp   = (16-row panel, column pair q, q+1) = 32 points = NTILE tiles of (4 rows x 2 columns)
//   lane l : B operand holds X[slab 4*chunk + (l&3)][point l>>2]; point n = 2*row + col, so lane l^4 is the
//            other column of the same row: ONE Philox call (two neighbouring columns, one slab) per lane
//            and one 64-bit shuffle yield the lane's values for the two new chunks of every 8-step tile
//   C frag : lane holds time t0 + (l>>2), rows 4*tile + (l&3), both columns -> per store instruction every
//            plane receives one full 32-byte sector, the 4 tiles of the warp complete the 128-byte lines
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void dmma884_t(double& c0, double& c1, const double a, const double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

template <int NT, int SRC>
__global__ void __launch_bounds__(256, (NT <= 33 ? 3 : 1))
df_time_mma_kernel(const double* __restrict__ X, double* __restrict__ T1, const Taps bx, const DfGeom g,
                   const PanelGeom panel, const long nwarp, const int nsteps, const int tchunk,
                   const PhiloxKeys keys, const long s_base) {
  constexpr int NCH = (NT + 7 + 3) / 4;        // 4-slab chunks under an 8-step tile
  constexpr int NSLOT = (NCH + 1) & ~1;        // window slots (even, so it rotates by 2 with a static period)
  constexpr int PERIOD = NSLOT / 2;            // tile steps until the slot pattern repeats
  constexpr int NTILE = (NT <= 65) ? 2 : 1;    // tiles (4 rows x 2 columns) per warp; register budget decides
  constexpr int NPART = 4 / NTILE;             // warps sharing one column pair of the 16-row panel
  const long W2 = blockIdx.x * 8L + (threadIdx.x >> 5);
  if (W2 >= nwarp) return;
  const long W = W2 / NPART;
  const int row0 = (int)(W2 % NPART) * 4 * NTILE;   // first panel row of this warp
  const int lane = threadIdx.x & 31;
  const int k = lane & 3, n = lane >> 2;       // B: slab-in-chunk, point;  A / C: k / column pair, row = time
  const int c = blockIdx.y;
  const int tc0 = blockIdx.z * tchunk;
  const int tc1 = min(tc0 + tchunk, nsteps);
  const int QP = panel.Qpad >> 1;
  const int pb = (int)(W / QP), q0 = 2 * (int)(W % QP);
  const int ro = n >> 1, co = n & 1;
  const bool colok = q0 < g.Kp;                // Kp is even: the pair is valid together
  const int last_slab = nsteps + NT - 2;

  double a[NCH];                               // Toeplitz band: A[i][4*cc + k] = bx[NT-1-(4*cc+k-i)]
#pragma unroll
  for (int cc = 0; cc < NCH; ++cc) {
    const int d = 4 * cc + k - n;              // n doubles as the A row (time) index i
    a[cc] = (d >= 0 && d < NT) ? bx.b[NT - 1 - d] : 0.0;
  }

  // the two values of this lane's point at slabs S1 = base + k and S2 = base + 4 + k
  auto load2 = [&](int tile, int base, double& v1, double& v2) {
    const int p = pb * 16 + row0 + 4 * tile + ro;
    const bool ok = colok && p < g.Jp;
    if (SRC != kSrcInjected) {
      const long pair = ((long)p * g.Kp + q0) >> 1;
      const int smine = base + k + 4 * co;     // even column computes slab S1, odd column S2
      double x0, x1;
      if (SRC == kSrcU32) {                    // one call = 4 slabs of ONE point: this lane's point, its slab
        x0 = x1 = 0.0;
      } else {
        philox_pair(keys, (uint64_t)pair, (uint64_t)(s_base + smine), c, x0, x1);
      }
      if (SRC == kSrcU32) {
        // lanes k = 0..3 of one point need word k of quads (base/4) and (base/4 + 1): even k computes the first,
        // odd k the second, and neighbours swap the word the other one needs (aligned streams); otherwise two calls
        const long sg = s_base + base;
        const uint64_t idx = (uint64_t)((long)p * g.Kp + q0 + co);
        if ((sg & 3) == 0) {
          uint32_t r[4];
          philox_quad(keys, idx, (uint64_t)((sg >> 2) + (k & 1)), c, r);
          const uint32_t mine = philox_pick(r, k), give = philox_pick(r, k ^ 1);
          const uint32_t got = __shfl_xor_sync(0xffffffffu, give, 1);
          v1 = philox_u32_to_uniform((k & 1) ? got : mine);
          v2 = philox_u32_to_uniform((k & 1) ? mine : got);
        } else {
          v1 = philox_value<kRngU32>(keys, idx, (uint64_t)(sg + k), c);
          v2 = philox_value<kRngU32>(keys, idx, (uint64_t)(sg + 4 + k), c);
        }
      } else {
        const double keep = co ? x1 : x0, send = co ? x0 : x1;
        const double recv = __shfl_xor_sync(0xffffffffu, send, 4);
        v1 = co ? recv : keep;
        v2 = co ? keep : recv;
      }
      if (!ok) v1 = v2 = 0.0;
    } else {
      const int s1 = min(max(base + k, 0), last_slab), s2 = min(max(base + 4 + k, 0), last_slab);
      const long off = (long)p * g.Kp + q0 + co;
      v1 = ok ? __ldcs(X + ((long)s1 * 3 + c) * g.Np + off) : 0.0;
      v2 = ok ? __ldcs(X + ((long)s2 * 3 + c) * g.Np + off) : 0.0;
    }
  };

  double wb[NTILE][NSLOT];
  // prefill chunk positions -1 .. NCH-3 (pairs); position cc at tile step u lives in slot (cc + 2u) % NSLOT
#pragma unroll
  for (int tile = 0; tile < NTILE; ++tile)
#pragma unroll
    for (int cc = -1; cc < NCH - 2; cc += 2) {
      double v1, v2;
      load2(tile, tc0 + 4 * cc, v1, v2);
      wb[tile][(cc + NSLOT) % NSLOT] = v1;
      if (cc + 1 < NCH - 2) wb[tile][(cc + 1 + NSLOT) % NSLOT] = v2;
    }

  double* wbase = T1 + (long)c * panel.comp_stride + ((long)pb * panel.Qpad + q0) * 16;
  for (int t = tc0; t < tc1; t += 8 * PERIOD) {
#pragma unroll
    for (int u = 0; u < PERIOD; ++u) {
      const int t0 = t + 8 * u;
      if (t0 < tc1) {                          // warp-uniform
#pragma unroll
        for (int tile = 0; tile < NTILE; ++tile) {
          double v1, v2;
          load2(tile, t0 + 4 * (NCH - 2), v1, v2);      // the two newest chunks of this tile step
          wb[tile][(NCH - 2 + 2 * u) % NSLOT] = v1;
          wb[tile][(NCH - 1 + 2 * u) % NSLOT] = v2;
          // two independent accumulator chains (even / odd chunks): DMMA latency is long
          double c0 = 0.0, c1 = 0.0, d0 = 0.0, d1 = 0.0;
#pragma unroll
          for (int cc = 0; cc < NCH; ++cc) {
            if (cc & 1) dmma884_t(d0, d1, a[cc], wb[tile][(cc + 2 * u) % NSLOT]);
            else dmma884_t(c0, c1, a[cc], wb[tile][(cc + 2 * u) % NSLOT]);
          }
          // C fragment: time t0 + n, row row0 + 4*tile + k, columns q0 and q0 + 1
          const int tt = t0 + n;
          if (tt < tc1) {
            double* q = wbase + (long)tt * 3 * panel.comp_stride + row0 + 4 * tile + k;
            q[0] = c0 + d0;
            q[16] = c1 + d1;
          }
        }
      }
    }
  }
}

template <int NT>
static void launch_time_mma(const DfGeom& g, const PanelGeom& panel, const Taps& bx, const double* X, double* T1,
                            int nsteps, int src, uint64_t seed, long s_base, cudaStream_t st) {
  constexpr int NCH = (NT + 7 + 3) / 4, PERIOD = ((NCH + 1) & ~1) / 2;
  constexpr int NPART = (NT <= 65) ? 2 : 4;                  // warps per column pair (must match the kernel)
  const long nwarp = (long)NPART * panel.PB * (panel.Qpad / 2);
  int nchunk = 1;
  const long ctas_per_chunk = 3L * ceil_div(nwarp, 8);
  while (nchunk < 64 && ctas_per_chunk * nchunk < 148L * 2 * 8 && nsteps / (nchunk * 2) >= 4 * NT) nchunk *= 2;
  int tchunk = ceil_div(nsteps, nchunk);
  tchunk = ceil_div(tchunk, 8 * PERIOD) * (8 * PERIOD);
  nchunk = ceil_div(nsteps, tchunk);
  dim3 grid(ceil_div(nwarp, 8), 3, nchunk);
  const PhiloxKeys keys = philox_keys(seed);
  if (src == kSrcU32)
    df_time_mma_kernel<NT, kSrcU32><<<grid, 256, 0, st>>>(X, T1, bx, g, panel, nwarp, nsteps, tchunk, keys, s_base);
  else if (src == kSrcU53)
    df_time_mma_kernel<NT, kSrcU53><<<grid, 256, 0, st>>>(X, T1, bx, g, panel, nwarp, nsteps, tchunk, keys, s_base);
  else
    df_time_mma_kernel<NT, kSrcInjected><<<grid, 256, 0, st>>>(X, T1, bx, g, panel, nwarp, nsteps, tchunk, keys, s_base);
  count_launch();
}

// 0 = auto (DFMA sliding window up to 25 taps, tensor path from 33 taps), 1 = DFMA kernels only,
// 2 = tensor-path (DMMA) kernel wherever the layout allows
static int g_time_variant = 0;
void df_set_time_variant(int v) { g_time_variant = v; }

int launch_df_time(const DfGeom& g, const PanelGeom& panel, const Taps& bx, const double* X, double* T1, int nsteps,
                   int src, uint64_t seed, long s_base, bool tuned, cudaStream_t st) {
  if (nsteps <= 0) return 0;
  const int nt = 2 * g.nfx + 1;
  if (tuned) {
    if (panel.PB > 0 && g_time_variant != 1) {
#define F(N) case N: launch_time_mma<N>(g, panel, bx, X, T1, nsteps, src, seed, s_base, st); PODSDF_CUDA(cudaGetLastError()); return 0;
      if (g_time_variant == 2 || src == kSrcInjected) {   // injected slabs: the tensor-path kernel's 8-byte loads pattern is faster
        switch (nt) {
          PODSDF_TIME_TAPS(F)
          default: break;
        }
      }
      // from 33 taps on the tensor path wins over the pair kernel; beyond, a DFMA register window no longer fits.
      // The U32 stream's ring-of-8 kernel still holds 33 taps (ring of 40 registers) and beats the tensor path there
      // (1.29 vs 1.94 ms per 512 planes, profiles/r02_config5_sweep.md); PODSDF_TIME_VARIANT=2 forces the tensor path
      if (!(nt == 33 && src == kSrcU32 && g_time_variant != 2)) {
        switch (nt) {
          F(33) F(49) F(65) F(97) F(129)
          default: break;
        }
      }
#undef F
    }
    switch (nt) {
#define F(N) case N: launch_time_tuned<N>(g, panel, bx, X, T1, nsteps, src, seed, s_base, st); break;
      PODSDF_TIME_TAPS(F)
#undef F
      default:
        if (wide_pass_taps(nt) == 49) launch_time_wide<49>(g, panel, bx, nt, X, T1, nsteps, src, seed, s_base, st);
        else if (wide_pass_taps(nt) == 65) launch_time_wide<65>(g, panel, bx, nt, X, T1, nsteps, src, seed, s_base, st);
        else return fail(PODSDF_ERR_UNSUPPORTED, "no tuned time-filter kernel for this nfx");
    }
  } else {
    if (src != kSrcInjected) return fail(PODSDF_ERR_INVALID, "generic time filter needs materialised slabs");
    if (panel.PB > 0) return fail(PODSDF_ERR_INVALID, "generic time filter writes the canonical layout only");
    int zt = nsteps < 1024 ? nsteps : 1024;
    dim3 grid(ceil_div(g.Np, 256), 3, zt);
    df_time_generic_kernel<<<grid, 256, 0, st>>>(X, T1, bx, nt, g.Np, nsteps);
    count_launch();
  }
  PODSDF_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace podsdf
