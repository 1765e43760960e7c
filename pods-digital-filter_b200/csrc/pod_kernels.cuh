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
int launch_mean_rows(double* A, long M, long N, int layout, double* mean, int subtract, cudaStream_t st);
int launch_transpose(const double* in, double* out, long n, cudaStream_t st);

}  // namespace podsdf
