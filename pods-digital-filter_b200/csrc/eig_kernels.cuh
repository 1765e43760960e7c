// Launchers of the symmetric tridiagonal eigen-solver pieces (implemented in eig_kernels.cu).
#pragma once
#include "common.cuh"

namespace podsdf {

// Eigenvalues (ascending) of the tridiagonal matrix (d[N], e[N-1]).  prepare: e2 = e^2 (N doubles of scratch) and
// bounds[4] = Gershgorin interval, pivmin, |T|.  bisect: indices [0, k_end), four lanes each (5-section).  multisect:
// indices [k_first, N), one warp each (33-section: a fifth of the serial Sturm sweeps of a plain bisection).
int launch_tridiag_prepare(const double* d, const double* e, long N, double* e2, double* bounds, cudaStream_t st);
int launch_tridiag_bisect(const double* d, const double* e2, const double* bounds, long N, long k_end, double* lam,
                          cudaStream_t st, int reserve_sms = 0);
int launch_tridiag_multisect(const double* d, const double* e2, const double* bounds, long N, long k_first, double* lam,
                             cudaStream_t st);
size_t tridiag_invit_workspace_doubles(long N, int nm);
// unit eigenvectors of T for the nm largest eigenvalues: Z[j][:] <-> lam[N-1-j]; resid[j] = scaled residual
int launch_tridiag_invit(const double* d, const double* e, const double* bounds, const double* lam, long N, int nm,
                         double* ws, double* Z, double* resid, cudaStream_t st);
// make the nm rows of Z (nm x N) orthonormal (Cholesky-QR); S: nm*nm doubles of scratch; flag[0] != 0 on breakdown
int launch_orthonormalize_rows(double* Z, long N, int nm, double* S, int* flag, cudaStream_t st);

// Z[j][:] = C[N-1-j][:], j < nm
int launch_rows_gather_descending(const double* C, long N, int nm, double* Z, cudaStream_t st);

}  // namespace podsdf
