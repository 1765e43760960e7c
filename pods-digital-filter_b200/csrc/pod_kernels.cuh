// Launchers of the PODFS kernels (implemented in pod_kernels.cu).
#pragma once
#include "common.cuh"

namespace podsdf {

// bytes of workspace launch_gram needs for an (M x N) snapshot matrix
size_t gram_workspace_bytes(long M, long N);
// C[N][N] = scale * A^T diag(w) A (exactly symmetric). layout 0: A[row][t]; 1: a[t][row]. w may be null.
int launch_gram(const double* A, long M, long N, int layout, const double* w, long P, double scale, double* C,
                double* W, size_t W_bytes, cudaStream_t st);
// Phi[M][nm] = scale * (A . T[:, :nm]) . diag(dinv)
int launch_project(const double* A, long M, long N, int layout, const double* T, long ldT, const double* dinv,
                   int nm, double scale, double* Phi, cudaStream_t st);
// c[N][nm] complex128 = fftshift-ordered DFT of the first nm columns of a (N x ld); tw_ws: dft_workspace_doubles
// (twiddle table + the intermediate of the two-factor transform)
size_t dft_workspace_doubles(long N, int nm);
int launch_dft(const double* a, long N, long ld, int nm, double* tw_ws, double* c, cudaStream_t st);
// statistics of a 3-component field (layout as above with M = 3 P): mean[3][P], stress[6][P] = uu, vv, ww, uv, uw, vw
int launch_field_moments(const double* F, long P, long N, int layout, double* mean, double* stress, cudaStream_t st);
// <a'(t,j,k) a'(t+dt,j+dj,k+dk)> of component comp for nl <= lag_covariance_max() lags (lags3[nl][3] on the host);
// mean (3 P, may be null for a mean-free field); ws: lag_covariance_workspace_doubles(); out_dev[nl]
int lag_covariance_max();
size_t lag_covariance_workspace_doubles();
int launch_lag_covariance(const double* F, long J, long K, long N, int layout, const double* mean, int comp,
                          const int* lags3, int nl, double* ws, double* out_dev, cudaStream_t st);
int launch_mean_rows(double* A, long M, long N, int layout, double* mean, int subtract, cudaStream_t st);
// packed[i*N - i(i-1)/2 + (j-i)] = C[i][j], j >= i; unpack mirrors
int launch_sym_pack(const double* C, long N, double* packed, cudaStream_t st);
int launch_sym_unpack(const double* packed, long N, double* C, cudaStream_t st);
int launch_transpose(const double* in, double* out, long n, cudaStream_t st);

}  // namespace podsdf
