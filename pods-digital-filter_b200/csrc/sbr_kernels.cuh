// Two-stage reduction of a dense symmetric matrix to tridiagonal form and the back-transformation of a few
// eigenvectors (implemented in sbr_kernels.cu; tools/proto_sbr.py is the numpy model of the same task order).
// Replaces the cuSOLVER Dsytrd / Dormtr calls of round 1 on the "all eigenvalues + nm leading eigenvectors" route
// that stands in for numpy.linalg.eig at PODFS.py:1309.
#pragma once
#include "common.cuh"

namespace podsdf {

constexpr int kSbrB = 32;          // band half-width after stage 1 (= panel width)
constexpr int kSbrLdb = 2 * kSbrB; // leading dimension of the band storage: the bulge reaches 2b-1 below the diagonal

// workspace (doubles) of sbr_tridiagonalize for an N x N matrix
size_t sbr_workspace_doubles(long N);

// Stage 1 + stage 2.  A (N x N, symmetric, both triangles valid on entry, overwritten) -> d[N], e[N-1].
// ws (sbr_workspace_doubles) keeps the reflectors of both stages for sbr_back_transform.
// timings (may be null): ms of [stage 1, stage 2] measured with events on `st` (debug / profiling only: it syncs).
int sbr_tridiagonalize(double* A, long N, double* d, double* e, double* ws, cudaStream_t st, float* timings);

// Z (nvec x N, row j = eigenvector j of the tridiagonal matrix) <- (Q1 Q2 Z^T)^T: eigenvectors of the dense matrix.
int sbr_back_transform(const double* ws, long N, double* Z, int nvec, cudaStream_t st);

// testing: stage 1 only; band (device, N x kSbrLdb) receives band[c * kSbrLdb + d] = B[c + d][c], d <= kSbrB
int sbr_stage1_only(double* A, long N, double* ws, double* band, cudaStream_t st);

}  // namespace podsdf
