/* libpodsdf -- C ABI of the B200-native digital-filter (DF) + PODFS hot paths.
 *
 * This is the drop-in boundary: plain pointers and sizes, no torch / numpy types.  The Python
 * shims in pods-digital-filter_b200/{digitalfilters,PODFS}.py bind it with ctypes and keep the
 * reference's function names and argument order; INTEGRATION.md shows the stub a maintainer of
 * the reference would add.  "ref:" comments cite the reference interface each entry replaces
 * (DF = digitalfilters.py, POD = PODFS.py under /root/reference).
 *
 * Conventions
 *  - every function returns 0 on success, else a PODSDF_ERR_* code; podsdf_last_error() gives
 *    the text (thread-local).  There is NO CPU fallback: without a usable sm_100 GPU every
 *    compute entry returns PODSDF_ERR_NO_DEVICE.
 *  - `_host` entries take host pointers and do the H2D / D2H copies themselves (pinned staging,
 *    chunked); `_dev` entries take device pointers on the plan's device and a cudaStream_t
 *    (passed as void*, 0 = default stream) and are asynchronous.
 *  - all arrays are float64, C order.  J = jma (span), K = kma (wall-normal, fastest axis),
 *    nf* = HALF filter width (2nf+1 taps), Jp = J+2nfy, Kp = K+2nfz, P = J*K.
 *  - DF planes use the layout out[t][c][j][k]  (t = step, c = u,v,w): one plane is one column
 *    A[:, t] of the reference snapshot matrix (DF:1471-1473), i.e. `out` is A transposed.
 */
#ifndef PODSDF_H
#define PODSDF_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define PODSDF_API __attribute__((visibility("default")))
#else
#define PODSDF_API
#endif

#define PODSDF_OK 0
#define PODSDF_ERR_INVALID 1
#define PODSDF_ERR_CUDA 2
#define PODSDF_ERR_NO_DEVICE 3
#define PODSDF_ERR_UNSUPPORTED 4
#define PODSDF_ERR_NUMERIC 5

/* snapshot-matrix layouts accepted by the POD entries */
#define PODSDF_LAYOUT_ROWS_BY_TIME 0 /* reference layout A[row][t], row = c*P+p, time contiguous (DF:1397) */
#define PODSDF_LAYOUT_TIME_BY_ROWS 1 /* DF-native layout A^T: a[t][row], space contiguous            */

/* kernel selection (testing / profiling): 0 = auto, 1 = generic kernels, 2 = tuned kernels only */
#define PODSDF_IMPL_AUTO 0
#define PODSDF_IMPL_GENERIC 1
#define PODSDF_IMPL_TUNED 2

/* random streams of the seeded DF entries (csrc/philox.cuh; oracle/philox.py restates both):
 * U32 (default): Philox4x32-10, one call = 2 points x 2 slabs, 32 random bits per number, centred levels;
 * U53: round 1's stream, one call = 2 points of one slab, numpy's 53-bit uniform(). */
#define PODSDF_RNG_U53 0
#define PODSDF_RNG_U32 1

PODSDF_API const char* podsdf_last_error(void);
PODSDF_API int podsdf_version(void);
/* number of usable sm_100 devices (0 and PODSDF_ERR_NO_DEVICE if none) */
PODSDF_API int podsdf_device_count(int* count);
/* kernels launched by this library since load (bench.py reports it as gpu_launches) */
PODSDF_API long long podsdf_launch_count(void);
/* pinned host memory for async staging (ref: H5:11-64 consumers read results from here) */
PODSDF_API int podsdf_host_alloc(void** ptr, size_t bytes);
PODSDF_API int podsdf_host_free(void* ptr);

/* ------------------------------------------------------------------ digital filter (DF) */
typedef struct podsdf_df_plan podsdf_df_plan;

/* ref: filter sizes/taps DF:73-89 (calccoeff), DF:100-140 (filter3DSciPy1D), DF:1277-1285.
 * bx,by,bz: host tap vectors of 2nf+1 doubles each (the shim computes them exactly as calccoeff).
 * The filter is the true convolution of DF:124-137 (taps flipped; identical for symmetric taps). */
PODSDF_API int podsdf_df_plan_create(int device, int J, int K, int nfx, int nfy, int nfz, const double* bx,
                          const double* by, const double* bz, podsdf_df_plan** plan);
PODSDF_API int podsdf_df_plan_destroy(podsdf_df_plan* plan);
PODSDF_API int podsdf_df_plan_set_impl(podsdf_df_plan* plan, int impl);
/* planes generated per internal pass (workspace = min(chunk, steps of the call)*3*Jp*Kp doubles, grown on
 * demand); rounded down to a multiple of 8 (minimum 8); default: what fits a 12 GiB workspace, at most 1280.
 * Results do not depend on it bitwise. */
PODSDF_API int podsdf_df_plan_set_chunk(podsdf_df_plan* plan, int chunk_planes);

/* which Philox stream the *_seeded_* and random_slabs entries use (PODSDF_RNG_*; default U32).  The reference
 * draws from the unseeded global MT19937 (DF:1344, 1361-1366), so there is no reference stream to match. */
PODSDF_API int podsdf_df_plan_set_rng(podsdf_df_plan* plan, int kind);

/* ref: adapt1d DF:143-178, adapt2prf DF:180-231, adapt2d DF:233-485, rotate_velocity DF:1119-1131.
 * All four are y = M(j,k) x + m(j,k) with time-invariant M (3x3) and m (3): the shim evaluates
 * the Cholesky / clamps / splines / rotation once on the host and hands over
 * coef[12][J*K] = M00,M01,M02,M10,...,M22,m0,m1,m2 planes.  mask bit i set = plane i present
 * (absent planes are exact zeros and are neither stored nor read).  coef == NULL: identity. */
PODSDF_API int podsdf_df_set_coefficients(podsdf_df_plan* plan, const double* coef_host, unsigned mask);
/* ref: rotate_velocity DF:1119-1131 / prof_rotation_matrix DF:1064-1116.  Constant 3x3 (row-major) applied to
 * every velocity vector after the Lund transform; NULL switches it off. */
PODSDF_API int podsdf_df_set_rotation(podsdf_df_plan* plan, const double* R9);

/* ref: filter3DSciPy1D(x,y,...) DF:100-140.  x_host (2nfx+1, Jp, Kp) -> y_host (J, K), no Lund. */
PODSDF_API int podsdf_df_filter_block_host(podsdf_df_plan* plan, const double* x_host, double* y_host);
/* ref: adapt*(yu,yv,yw,...) in place on three (J,K) host planes, using the plan coefficients. */
PODSDF_API int podsdf_df_lund_apply_host(podsdf_df_plan* plan, double* yu, double* yv, double* yw);

/* ref: the time loop DF:1403-1477 on an injected random stream.
 * slabs[n_slabs][3][Jp][Kp]; step i uses slabs i..i+2nfx (DF:1454-1467), so
 * nsteps = n_slabs - 2nfx planes are written to out[nsteps][3][J][K]. */
PODSDF_API int podsdf_df_generate_injected_host(podsdf_df_plan* plan, const double* slabs_host, long n_slabs,
                                     double* out_host);
PODSDF_API int podsdf_df_generate_injected_dev(podsdf_df_plan* plan, const double* slabs_dev, long n_slabs,
                                    double* out_dev, void* stream);
/* same loop with the random stream generated on the device (Philox4x32-10, see csrc/philox.cuh and
 * podsdf_df_plan_set_rng):
 * planes t0 .. t0+nsteps-1 of the infinite stream `seed` (any shard reproduces its own window). */
PODSDF_API int podsdf_df_generate_seeded_host(podsdf_df_plan* plan, uint64_t seed, long t0, long nsteps,
                                   double* out_host);
PODSDF_API int podsdf_df_generate_seeded_dev(podsdf_df_plan* plan, uint64_t seed, long t0, long nsteps,
                                  double* out_dev, void* stream);
/* materialise slabs [s0, s0+n) of the Philox stream: slabs[n][3][Jp][Kp] (tests, halo checks) */
PODSDF_API int podsdf_df_random_slabs_host(podsdf_df_plan* plan, uint64_t seed, long s0, long n, double* slabs_host);
PODSDF_API int podsdf_df_random_slabs_dev(podsdf_df_plan* plan, uint64_t seed, long s0, long n, double* slabs_dev,
                               void* stream);
/* device time (ms) of the kernels of the last generate_* call on this plan: [0] time filter,
 * [1] plane filter + Lund, measured with CUDA events on the launching stream */
PODSDF_API int podsdf_df_last_kernel_ms(podsdf_df_plan* plan, double* ms2);
/* what the plan will run with its current impl / coefficients / rotation: info[0] planes per pass (chunk),
 * info[1] 1 if a tuned time filter exists for 2*nfx+1 taps, info[2] 1 if a tuned plane filter exists for
 * (2*nfy+1, 2*nfz+1) and the current transform, info[3] 1 if the transform is diagonal (per-component plane
 * kernel) and 0 if it couples components, info[4] passes of the last generate_* call */
PODSDF_API int podsdf_df_plan_info(podsdf_df_plan* plan, int* info5);

/* ------------------------------------------------------------------------------ PODFS */
/* ref: mean removal DF:1492-1495.  mean[M] = mean over time; A -= mean (in place) if subtract. */
PODSDF_API int podsdf_mean_rows_dev(int device, double* A_dev, long M, long N, int layout, double* mean_dev,
                         int subtract, void* stream);

/* ref: calculate_correlation_matrix POD:1451-1464.  C[N][N] = scale * A^T diag(w) A, exactly
 * symmetric (upper triangle computed with FP64 tensor-core MMA, mirrored).  A has M = ncomp*P rows
 * and N snapshots; w (P doubles, shared by all components) may be NULL ('false' branch).
 * The reference's scale is 1/num_snapshots.  `_dev` leaves the partial sum of this rank's rows in
 * C_dev; the caller all-reduces across ranks (NCCL) before the eigen-solve. */
PODSDF_API int podsdf_gram_host(int device, const double* A_host, long M, long N, int layout, const double* w_host,
                     long P, double scale, double* C_host);
PODSDF_API int podsdf_gram_dev(int device, const double* A_dev, long M, long N, int layout, const double* w_dev,
                    long P, double scale, double* C_dev, void* stream);
/* Packed upper triangle of the (exactly symmetric) Gram matrix for the cross-rank all-reduce: N (N + 1) / 2 doubles
 * instead of N^2 (ref: the sum over spatial rows of POD:1455 split across ranks).  packed[i N - i (i - 1) / 2 + j - i] =
 * C[i][j], j >= i; unpack writes both triangles. */
PODSDF_API int podsdf_sym_pack_dev(int device, const double* C_dev, long N, double* packed_dev, void* stream);
PODSDF_API int podsdf_sym_unpack_dev(int device, const double* packed_dev, long N, double* C_dev, void* stream);
/* The PODFS entries keep grow-only per-device scratch between calls (Gram split workspace, eigen-solver workspace:
 * ~3 N^2 doubles after a leading-pairs solve); this frees it (synchronises the device first). */
PODSDF_API int podsdf_release_scratch(int device);
/* device time (ms) of the last Gram kernel on this thread */
PODSDF_API int podsdf_gram_last_ms(double* ms);

/* ref: linalg.eig(C) POD:1309 on the exactly symmetric C.  Ascending eigenvalues lambda[N] and orthonormal
 * eigenvectors.  `_host`: V_host[N][N] row-major with eigenvector j in COLUMN j (numpy's convention).
 * `_dev`: in place -- on exit ROW j of C_dev_inout_V is eigenvector j (no transpose pass).
 * The sort / NaN guard / validity count / scaling of POD:1310-1327, 1430-1447 stay in the Python shim. */
PODSDF_API int podsdf_eig_sym_host(int device, const double* C_host, long N, double* lambda_host, double* V_host);

/* "All eigenvalues + the nm leading eigenvectors" (what PODFS.py:1329-1333 consumes of linalg.eig's result):
 * lambda[N] ascending as above; on exit rows 0..nm-1 of C (dev) / V[nm][N] (host) are the unit eigenvectors of the
 * nm LARGEST eigenvalues in DESCENDING order; the remaining rows of C are scratch.  *path (may be NULL) = 1 when the
 * tridiagonal route ran -- all of it this library's kernels: two-stage reduction (dense -> band -> tridiagonal),
 * multisection / bisection, inverse iteration, back-transformation of the nm vectors, re-orthonormalisation --
 * 0 when the leading eigenvalues were not well separated (relative gap below 1e-7), a residual check failed or the
 * size is unusual (N < 256, N > 26000, 4 nm > N), and the full library solver of podsdf_eig_sym_* ran. */
PODSDF_API int podsdf_eig_sym_top_dev(int device, double* C_dev_inout_V, long N, int nm, double* lambda_dev, int* path,
                           void* stream);
PODSDF_API int podsdf_eig_sym_top_host(int device, const double* C_host, long N, int nm, double* lambda_host,
                            double* V_host, int* path);
PODSDF_API int podsdf_eig_sym_dev(int device, double* C_dev_inout_V, long N, double* lambda_dev, void* stream);
/* The reduction behind podsdf_eig_sym_top_*: two-stage tridiagonalisation of the symmetric C (N >= 3) with this
 * library's own kernels (dense -> band of half-width 32 -> tridiagonal).  d[N], e[N-1]: diagonal and sub-diagonal of
 * the similar tridiagonal matrix.  band (may be NULL): the intermediate band matrix, band[c*33 + k] = B[c+k][c],
 * k = 0..32.  ms2 (may be NULL): device time of the two stages in ms.  Testing / profiling entry. */
PODSDF_API int podsdf_sym_tridiag_host(int device, const double* C_host, long N, double* d_host, double* e_host,
                            double* band_host, double* ms2);

/* ref: projection POD:1329-1333.  Phi[M][nm] = A (M x N) . T (N x nm) . diag(dinv) * scale,
 * T = scaled temporal modes (row-major N x ldT, first nm columns used), dinv[nm] = 1/lambda. */
PODSDF_API int podsdf_project_host(int device, const double* A_host, long M, long N, int layout, const double* T_host,
                        long ldT, const double* dinv_host, int nm, double scale, double* Phi_host);
PODSDF_API int podsdf_project_dev(int device, const double* A_dev, long M, long N, int layout, const double* T_dev,
                       long ldT, const double* dinv_dev, int nm, double scale, double* Phi_dev,
                       void* stream);

/* ref: fourier_coefficients POD:1562-1573.  c[n][i] = (1/N) sum_j a[j][i] exp(-2 pi i (n - N/2) j / N)
 * for the first nm columns of the row-major (N x ld) matrix a.  Output interleaved complex128
 * c[N][nm][2].  The complex64 cast, ranking and energy cut (POD:1577-1593) stay in the shim.
 * The _dev entry is asynchronous: c_dev is complete in the order of `stream` (its workspace is a cached device buffer). */
PODSDF_API int podsdf_dft_modes_host(int device, const double* a_host, long N, long ld, int nm, double* c_host);
PODSDF_API int podsdf_dft_modes_dev(int device, const double* a_dev, long N, long ld, int nm, double* c_dev,
                         void* stream);

/* One- and two-point statistics of a three-component field held as a snapshot matrix (M = 3 P rows, N samples, either
 * layout): what a-posteriori checks of the PODFS reconstruction U~(x,t) of docs/usage/CFDCodeIntegration.rst:85-91
 * (the reference's plot-only y2, POD:1602-1612) and of the generated inlet itself compare -- the mean profile and
 * Reynolds stresses the Lund transform prescribes (DF:143-178) and the two-point correlations the filter imposes
 * (DF:73-89).  mean_dev[3][P]; stress_dev[6][P] = uu, vv, ww, uv, uw, vw (central moments over the N samples). */
PODSDF_API int podsdf_field_moments_dev(int device, const double* F_dev, long P, long N, int layout, double* mean_dev,
                             double* stress_dev, void* stream);
/* cov_host[l] = < a'(t, j, k) a'(t + dt_l, j + dj_l, k + dk_l) > over all valid (t, j, k) of component comp (0..2) of
 * the J x K plane; lags3_host[nl][3] = (dt, dj, dk) >= 0, nl <= 16; mean_dev (3 P doubles) is subtracted first (NULL:
 * the field is already mean-free).  Synchronous (the result lands on the host). */
PODSDF_API int podsdf_lag_covariance_dev(int device, const double* F_dev, long J, long K, long N, int layout,
                              const double* mean_dev, int comp, const int* lags3_host, int nl, double* cov_host,
                              void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PODSDF_H */
