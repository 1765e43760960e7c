// Launchers of the digital-filter kernels (implemented in df_kernels.cu).
//
// Data layout in HBM (all float64):
//   X   [s][c][Jp*Kp]   random slabs (injected) -- s = slab, c = u,v,w
//   T1  [t][c][Jp*Kp]   time-filtered padded planes (workspace, one chunk of steps)
//   out [t][c][J*K]     finished planes (Lund-transformed), the transposed snapshot matrix
//   coef[12][J*K]       M00..M22, m0..m2 planes of the time-invariant Lund/rotation transform
#pragma once
#include "common.cuh"

namespace podsdf {

struct DfGeom {
  int J, K, nfx, nfy, nfz;
  int Jp, Kp;     // padded plane
  long Np, P;     // Jp*Kp, J*K
};

// Row-panel layout of the workspace T1 between the two tuned kernels: [t][c][pb][q][16], i.e. panels of 16
// plane rows; inside a panel the 16 rows of one column q are contiguous (128 bytes) and columns follow each
// other, so the (64 + 2 nfz) columns x 16 rows a strip needs per marching step are ONE contiguous run.
// PB == 0 means the canonical [p][q] layout.
struct PanelGeom {
  int PB = 0;            // number of 16-row panels covering Jp
  int Qpad = 0;          // columns per panel: 64*ceil(K/64) + 2 nfz (every 64-column strip can read its window)
  long comp_stride = 0;  // doubles per (plane, component) = PB * Qpad * 16
};
PanelGeom make_panel_geom(const DfGeom& g);

// Constant rotation applied after the Lund transform (digitalfilters.py:1119-1131): y <- R y.  Kept as nine
// scalars in the kernel parameters instead of being folded into twelve dense coefficient planes.
struct Rot {
  double r[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  int on = 0;
};

// K-time variant for the panel layout: 0 = auto (DFMA sliding window up to 33 taps, DMMA Toeplitz for wider
// filters), 1 = DFMA only, 2 = DMMA wherever possible
void df_set_time_variant(int v);

// true when a tuned (compile-time tap count) kernel exists for this size
bool df_time_has_tuned(int ntx);
bool df_plane_has_tuned(int nty, int ntz, bool diagonal);
// true when the transform never mixes components (no Lund, or no off-diagonal plane of M present)
bool df_transform_is_diagonal(bool lund, unsigned mask, const Rot& rot);

// T1[t] = sum_m bx[m] X[t + (NTX-1-m)] for t in [0, nsteps);  X has nsteps + NTX - 1 slabs.
// src: 0 = injected slabs X; 1 / 2 = X is ignored and slab s is the Philox U53 / U32 stream (seed, s_base + s)
// (generated mode, tuned kernels only; the two streams are defined in philox.cuh).
int launch_df_time(const DfGeom& g, const PanelGeom& panel, const Taps& bx, const double* X, double* T1, int nsteps,
                   int src, uint64_t seed, long s_base, bool tuned, cudaStream_t st);

// out[t][c][j][k] = sum_c' M[c][c'] * (by (*) bz (*) T1[t][c'])[j][k] + m[c]   for t in [0, nplanes)
// ncomp_in = 3 normally; lund = false writes the filtered field unchanged (coef ignored).
// T2 is the generic path's workspace [nplanes][3][Jp][K] (may be null for the tuned path).
int launch_df_plane(const DfGeom& g, const PanelGeom& panel, const Taps& by, const Taps& bz, const double* T1,
                    double* out, int nplanes, int ncomp, const double* coef, unsigned mask, bool lund, const Rot& rot,
                    double* T2, bool tuned, cudaStream_t st);

// in-place y = M x + m on three planes (backs adapt*): planes[3][P]
int launch_df_lund(const DfGeom& g, double* planes, const double* coef, unsigned mask, bool lund, const Rot& rot,
                   cudaStream_t st);

// materialise Philox slabs [s0, s0+n) of stream `kind` (PODSDF_RNG_*) -> X[n][3][Np]
int launch_df_random_slabs(const DfGeom& g, uint64_t seed, int kind, long s0, long n, double* X, cudaStream_t st);

}  // namespace podsdf
