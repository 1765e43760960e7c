// C ABI of the PODFS path (include/podsdf.h).
#include <dlfcn.h>

#include <algorithm>
#include <cmath>
#include <mutex>
#include <vector>

#include "../../include/podsdf.h"
#include "eig_kernels.cuh"
#include "pod_kernels.cuh"
#include "sbr_kernels.cuh"

using namespace podsdf;

namespace {

thread_local double g_gram_ms = 0.0;

// Grow-only per-device scratch (Gram split-K workspace, eigen-solver workspace): cudaMalloc/cudaFree of
// hundreds of MB on every call costs more than the small kernels they serve.  One slot per purpose.
struct Scratch {
  double* p = nullptr;
  size_t bytes = 0;
};
Scratch g_scratch[64][4];      // slots: 0 Gram partials, 1 library solver, 2 leading-pairs solve, 3 DFT workspace
std::mutex g_scratch_mu;
int scratch_get(int device, int slot, size_t bytes, double** out) {
  std::lock_guard<std::mutex> lk(g_scratch_mu);
  Scratch& s = g_scratch[device & 63][slot];
  if (s.bytes < bytes) {
    if (s.p) cudaFree(s.p);
    s.p = nullptr;
    s.bytes = 0;
    PODSDF_CUDA(cudaMalloc(&s.p, bytes));
    s.bytes = bytes;
  }
  *out = s.p;
  return 0;
}

int scratch_release(int device) {
  std::lock_guard<std::mutex> lk(g_scratch_mu);
  for (Scratch& s : g_scratch[device & 63]) {
    if (s.p) cudaFree(s.p);
    s.p = nullptr;
    s.bytes = 0;
  }
  return 0;
}

struct DevBuf {  // RAII device allocation
  double* p = nullptr;
  cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 8); }
  ~DevBuf() {
    if (p) cudaFree(p);
  }
};

// ---- cuSOLVER, bound at run time -------------------------------------------------------------
// The FULL dense symmetric eigen-solve (all N eigenvectors: the drop-in POD() publishes every temporal mode,
// PODFS.py:1309, 1391) and the fallback of the leading-pairs route are the one library call left.  cuSOLVER is
// dlopen()ed so the library has no link-time dependency on a particular cuBLAS/cuSOLVER build: inside
// a PyTorch process the copy PyTorch already loaded is reused, otherwise the toolkit's is taken.
typedef void* cusolverDnHandle_t;
typedef int cusolverStatus_t;
struct Cusolver {
  void* lib = nullptr;
  cusolverStatus_t (*create)(cusolverDnHandle_t*) = nullptr;
  cusolverStatus_t (*destroy)(cusolverDnHandle_t) = nullptr;
  cusolverStatus_t (*set_stream)(cusolverDnHandle_t, cudaStream_t) = nullptr;
  cusolverStatus_t (*bufsize)(cusolverDnHandle_t, int, int, int, const double*, int, const double*, int*) = nullptr;
  cusolverStatus_t (*syevd)(cusolverDnHandle_t, int, int, int, double*, int, double*, double*, int, int*) = nullptr;
  std::string err;
};

Cusolver& cusolver() {
  static Cusolver cs;
  static std::once_flag once;
  std::call_once(once, [] {
    const char* names[] = {"libcusolver.so.11", "/usr/local/cuda/lib64/libcusolver.so.11", "libcusolver.so.12",
                           "libcusolver.so"};
    for (const char* n : names) {
      cs.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (cs.lib) break;
    }
    if (!cs.lib) {
      cs.err = std::string("cannot dlopen libcusolver: ") + (dlerror() ? dlerror() : "?");
      return;
    }
    cs.create = (decltype(cs.create))dlsym(cs.lib, "cusolverDnCreate");
    cs.destroy = (decltype(cs.destroy))dlsym(cs.lib, "cusolverDnDestroy");
    cs.set_stream = (decltype(cs.set_stream))dlsym(cs.lib, "cusolverDnSetStream");
    cs.bufsize = (decltype(cs.bufsize))dlsym(cs.lib, "cusolverDnDsyevd_bufferSize");
    cs.syevd = (decltype(cs.syevd))dlsym(cs.lib, "cusolverDnDsyevd");
    if (!cs.create || !cs.destroy || !cs.set_stream || !cs.bufsize || !cs.syevd) cs.err = "libcusolver lacks Dsyevd";
  });
  return cs;
}

cusolverDnHandle_t g_handles[64] = {};
std::recursive_mutex g_solver_mu;            // one solve at a time per process: the handles are shared

// eigen-decomposition in place: on exit row j of C is eigenvector j (ascending eigenvalues)
int eig_inplace(int device, double* C_dev, long N, double* lambda_dev, cudaStream_t st) {
  if (N > 46340) return fail(PODSDF_ERR_UNSUPPORTED, "N too large for the 32-bit cuSOLVER interface");
  Cusolver& cs = cusolver();
  if (!cs.err.empty()) return fail(PODSDF_ERR_CUDA, cs.err);
  std::lock_guard<std::recursive_mutex> lk(g_solver_mu);
  cusolverDnHandle_t& h = g_handles[device & 63];
  if (!h && cs.create(&h) != 0) return fail(PODSDF_ERR_CUDA, "cusolverDnCreate failed");
  int rc = 0, lwork = 0;
  double* work = nullptr;
  DevBuf info;
  const int n = (int)N;
  // CUSOLVER_EIG_MODE_VECTOR = 1, CUBLAS_FILL_MODE_LOWER = 0 (column-major view of a symmetric matrix)
  if (cs.set_stream(h, st) != 0) rc = fail(PODSDF_ERR_CUDA, "cusolverDnSetStream failed");
  if (!rc && cs.bufsize(h, 1, 0, n, C_dev, n, lambda_dev, &lwork) != 0)
    rc = fail(PODSDF_ERR_CUDA, "cusolverDnDsyevd_bufferSize failed");
  if (!rc) rc = scratch_get(device, 1, sizeof(double) * (size_t)lwork, &work);
  if (!rc && info.alloc(8) != cudaSuccess) rc = fail(PODSDF_ERR_CUDA, "eigen-solver workspace allocation failed");
  if (!rc) {
    const cusolverStatus_t s = cs.syevd(h, 1, 0, n, C_dev, n, lambda_dev, work, lwork, (int*)info.p);
    int hinfo = -1;
    cudaError_t e = cudaMemcpyAsync(&hinfo, info.p, sizeof(int), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) rc = cuda_fail(e, "Dsyevd", __FILE__, __LINE__);
    else if (s != 0) rc = fail(PODSDF_ERR_CUDA, "cusolverDnDsyevd returned status " + std::to_string(s));
    else if (hinfo != 0) rc = fail(PODSDF_ERR_NUMERIC, "Dsyevd did not converge, info = " + std::to_string(hinfo));
  }
  return rc;
}

// second stream + fork/join events per device for the part of the tridiagonal route that is off the critical path
struct SideStream {
  cudaStream_t s = nullptr;
  cudaEvent_t fork = nullptr, join = nullptr;
  bool ok = false;
};
SideStream& side_stream(int device) {
  static SideStream ss[64];
  SideStream& x = ss[device & 63];
  if (!x.ok) {
    x.ok = cudaStreamCreateWithFlags(&x.s, cudaStreamNonBlocking) == cudaSuccess &&
           cudaEventCreateWithFlags(&x.fork, cudaEventDisableTiming) == cudaSuccess &&
           cudaEventCreateWithFlags(&x.join, cudaEventDisableTiming) == cudaSuccess;
  }
  return x;
}

// All eigenvalues (ascending, lambda_dev) and the eigenvectors of the nm largest: on exit rows 0..nm-1 of C are
// those eigenvectors in DESCENDING order of the eigenvalue (the other rows are scratch).
// Route 1 (all of it this library's own kernels): two-stage tridiagonalisation (sbr_kernels.cu: dense -> band with
// FP64 tensor-core rank-2b updates, band -> tridiagonal by bulge chasing) -> multisection / bisection for all
// eigenvalues -> inverse iteration for nm vectors of T -> back-transformation through both stages -> Cholesky-QR.
// Route 0 (taken when the leading eigenvalues are not well separated, when a check of route 1 fails, or for small /
// unusual sizes): cuSOLVER's full Dsyevd on a saved copy of C, then the leading rows are gathered.
// *path reports which one ran.  PODSDF_EIG_ROUTE=0 forces route 0 (A/B timing).
int eig_top_inplace(int device, double* C_dev, long N, int nm, double* lambda_dev, cudaStream_t st, int* path) {
  if (N > 46340) return fail(PODSDF_ERR_UNSUPPORTED, "N too large for the 32-bit cuSOLVER interface");
  if (nm < 1 || nm > N) return fail(PODSDF_ERR_INVALID, "nm must be in 1..N");
  std::lock_guard<std::recursive_mutex> lk(g_solver_mu);
  const char* force = getenv("PODSDF_EIG_ROUTE");
  const bool eligible = N >= 256 && N <= 26000 && nm <= 512 && 4L * nm <= N && !(force && atoi(force) == 0);
  // scratch: [C copy | Z nm*N | d | e | e2 | bounds 8 | resid nm | S nm*nm | flag | invit | sbr workspace]
  const size_t nn = (size_t)N * N, zn = (size_t)nm * N;
  const size_t invit = eligible ? tridiag_invit_workspace_doubles(N, nm) : 0;
  const size_t sbrw = eligible ? sbr_workspace_doubles(N) : 0;
  const size_t total = nn + zn + 3 * (size_t)N + 8 + (size_t)nm + (size_t)nm * nm + 2 + invit + sbrw + 64;
  double* ws = nullptr;
  PODSDF_TRY(scratch_get(device, 2, sizeof(double) * total, &ws));
  double* Ccopy = ws;
  double* Z = Ccopy + nn;
  double* d = Z + zn;
  double* e = d + N;
  double* e2 = e + N;
  double* bounds = e2 + N;
  double* resid = bounds + 8;
  double* S = resid + nm;
  int* flag = reinterpret_cast<int*>(S + (size_t)nm * nm);
  double* winv = S + (size_t)nm * nm + 2;
  double* wsbr = winv + invit;
  wsbr += (32 - ((wsbr - ws) & 31)) & 31;             // keep the 256-byte alignment of the scratch base
  *path = 0;
  int rc = 0;
  bool fast_ok = false, side_pending = false;
  if (eligible) {
    PODSDF_CUDA(cudaMemcpyAsync(Ccopy, C_dev, sizeof(double) * nn, cudaMemcpyDeviceToDevice, st));
    rc = sbr_tridiagonalize(C_dev, N, d, e, wsbr, st, nullptr);
    if (rc) {                                   // e.g. a cooperative launch refused: the library solver still works
      cudaGetLastError();
      PODSDF_CUDA(cudaMemcpyAsync(C_dev, Ccopy, sizeof(double) * nn, cudaMemcpyDeviceToDevice, st));
      rc = eig_inplace(device, C_dev, N, lambda_dev, st);
      if (rc) return rc;
      PODSDF_TRY(launch_rows_gather_descending(C_dev, N, nm, Z, st));
      PODSDF_CUDA(cudaMemcpyAsync(C_dev, Z, sizeof(double) * zn, cudaMemcpyDeviceToDevice, st));
      PODSDF_CUDA(cudaStreamSynchronize(st));
      return 0;
    }
    // leading nm+1 eigenvalues first (warp-wide multisection, on the caller's stream: the inverse iteration waits
    // for them); the other N-nm-1 by plain bisection on a side stream, concurrently with everything below
    const long first = N - 1 - nm;      // ascending index of the (nm+1)-th largest (>= 0 because 4 nm <= N)
    PODSDF_TRY(launch_tridiag_prepare(d, e, N, e2, bounds, st));
    SideStream& side = side_stream(device);
    if (!side.ok) return fail(PODSDF_ERR_CUDA, "side stream creation failed");
    static const int side_mode = getenv("PODSDF_EIG_SIDE") ? atoi(getenv("PODSDF_EIG_SIDE")) : 1;   // A/B: 0 = after the vectors
    static const bool fat_side = !(getenv("PODSDF_EIG_SIDE_FAT") && atoi(getenv("PODSDF_EIG_SIDE_FAT")) == 0);   // A/B
    if (side_mode == 1) {
      PODSDF_CUDA(cudaEventRecord(side.fork, st));
      PODSDF_CUDA(cudaStreamWaitEvent(side.s, side.fork, 0));
      rc = launch_tridiag_bisect(d, e2, bounds, N, first, lambda_dev, side.s, fat_side ? nm : 0);
      if (!rc) {
        PODSDF_CUDA(cudaEventRecord(side.join, side.s));
        side_pending = true;
      }
    }
    auto join_side = [&]() {            // every exit below waits for the side stream: it shares the scratch
      if (side_pending) cudaStreamWaitEvent(st, side.join, 0);
      side_pending = false;
    };
    if (!rc) rc = launch_tridiag_multisect(d, e2, bounds, N, first, lambda_dev, st);
    // the leading nm+1 eigenvalues decide whether independent inverse iterations are safe
    std::vector<double> top((size_t)nm + 1);
    double hb[4] = {0, 0, 0, 0};
    if (!rc) {
      cudaError_t ce = cudaMemcpyAsync(top.data(), lambda_dev + first, sizeof(double) * (nm + 1), cudaMemcpyDeviceToHost, st);
      if (ce == cudaSuccess) ce = cudaMemcpyAsync(hb, bounds, sizeof(double) * 4, cudaMemcpyDeviceToHost, st);
      if (ce == cudaSuccess) ce = cudaStreamSynchronize(st);
      if (ce != cudaSuccess) rc = cuda_fail(ce, "eigenvalue readback", __FILE__, __LINE__);
    }
    if (rc) {
      join_side();
      cudaStreamSynchronize(st);
      return rc;
    }
    const double lmax = std::fabs(top[nm]);
    bool separated = lmax > 0.0;
    for (int j = 0; j < nm && separated; ++j)
      if (!(top[j + 1] - top[j] > 1.0e-7 * lmax)) separated = false;
    if (separated) {
      rc = launch_tridiag_invit(d, e, bounds, lambda_dev, N, nm, winv, Z, resid, st);
      if (!rc) rc = sbr_back_transform(wsbr, N, Z, nm, st);
      if (!rc) rc = launch_orthonormalize_rows(Z, N, nm, S, flag, st);
      std::vector<double> hres((size_t)nm);
      int hflag = 1;
      if (!rc) {
        cudaError_t ce = cudaMemcpyAsync(hres.data(), resid, sizeof(double) * nm, cudaMemcpyDeviceToHost, st);
        if (ce == cudaSuccess) ce = cudaMemcpyAsync(&hflag, flag, sizeof(int), cudaMemcpyDeviceToHost, st);
        if (ce == cudaSuccess) ce = cudaStreamSynchronize(st);
        if (ce != cudaSuccess) rc = cuda_fail(ce, "eigenvector check readback", __FILE__, __LINE__);
      }
      if (rc) {
        join_side();
        cudaStreamSynchronize(st);
        return rc;
      }
      // acceptance: the tridiagonal residual |T z - lambda z|_inf / |T| of every vector at rounding level
      // (inverse iteration with an eigenvalue accurate to eps |T| gives ~eps; with gaps > 1e-7 |T| the vector
      // error is then below ~1e-8 before, and at rounding level after, the re-orthonormalisation)
      fast_ok = hflag == 0;
      for (int j = 0; j < nm && fast_ok; ++j)
        if (!(hres[j] < 1.0e-13)) fast_ok = false;
    }
    if (side_mode != 1 && !rc) rc = launch_tridiag_bisect(d, e2, bounds, N, first, lambda_dev, st);
    join_side();                                             // all of lambda is written
    if (!fast_ok) PODSDF_CUDA(cudaMemcpyAsync(C_dev, Ccopy, sizeof(double) * nn, cudaMemcpyDeviceToDevice, st));
  }
  if (fast_ok) {
    *path = 1;
  } else {
    rc = eig_inplace(device, C_dev, N, lambda_dev, st);
    if (rc) return rc;
    PODSDF_TRY(launch_rows_gather_descending(C_dev, N, nm, Z, st));
  }
  PODSDF_CUDA(cudaMemcpyAsync(C_dev, Z, sizeof(double) * zn, cudaMemcpyDeviceToDevice, st));
  PODSDF_CUDA(cudaStreamSynchronize(st));
  return 0;
}

}  // namespace

extern "C" {

int podsdf_mean_rows_dev(int device, double* A_dev, long M, long N, int layout, double* mean_dev, int subtract,
                         void* stream) {
  if (!A_dev || !mean_dev || M < 1 || N < 1 || (layout != 0 && layout != 1)) return fail(PODSDF_ERR_INVALID, "bad argument");
  PODSDF_TRY(check_device(device));
  DeviceGuard guard(device);
  return launch_mean_rows(A_dev, M, N, layout, mean_dev, subtract, (cudaStream_t)stream);
}

int podsdf_gram_dev(int device, const double* A_dev, long M, long N, int layout, const double* w_dev, long P,
                    double scale, double* C_dev, void* stream) {
  if (!A_dev || !C_dev || M < 1 || N < 1 || (layout != 0 && layout != 1)) return fail(PODSDF_ERR_INVALID, "bad argument");
  if (w_dev && (P < 1 || M % P != 0)) return fail(PODSDF_ERR_INVALID, "weights need M = ncomp * P");
  PODSDF_TRY(check_device(device));
  DeviceGuard guard(device);
  cudaStream_t st = (cudaStream_t)stream;
  const size_t wb = gram_workspace_bytes(M, N);
  double* ws = nullptr;
  PODSDF_TRY(scratch_get(device, 0, wb, &ws));
  cudaEvent_t e0, e1;
  PODSDF_CUDA(cudaEventCreate(&e0));
  PODSDF_CUDA(cudaEventCreate(&e1));
  cudaEventRecord(e0, st);
  int rc = launch_gram(A_dev, M, N, layout, w_dev, P > 0 ? P : 1, scale, C_dev, ws, wb, st);
  cudaEventRecord(e1, st);
  cudaError_t e = cudaEventSynchronize(e1);  // the shared scratch may be reused by the next call
  float ms = 0;
  if (e == cudaSuccess) cudaEventElapsedTime(&ms, e0, e1);
  g_gram_ms = ms;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  if (!rc && e != cudaSuccess) rc = cuda_fail(e, "gram", __FILE__, __LINE__);
  return rc;
}

int podsdf_release_scratch(int device) {
  PODSDF_TRY(check_device(device));
  DeviceGuard guard(device);
  PODSDF_CUDA(cudaDeviceSynchronize());
  return scratch_release(device);
}

int podsdf_gram_last_ms(double* ms) {
  if (!ms) return fail(PODSDF_ERR_INVALID, "null argument");
  *ms = g_gram_ms;
  return 0;
}

int podsdf_gram_host(int device, const double* A_host, long M, long N, int layout, const double* w_host, long P,
                     double scale, double* C_host) {
  if (!A_host || !C_host || M < 1 || N < 1) return fail(PODSDF_ERR_INVALID, "bad argument");
  PODSDF_TRY(check_device(device));
  DeviceGuard guard(device);
  DevBuf A, C, w;
  PODSDF_CUDA(A.alloc(sizeof(double) * (size_t)M * N));
  PODSDF_CUDA(C.alloc(sizeof(double) * (size_t)N * N));
  PODSDF_CUDA(cudaMemcpy(A.p, A_host, sizeof(double) * (size_t)M * N, cudaMemcpyHostToDevice));
  if (w_host) {
    PODSDF_CUDA(w.alloc(sizeof(double) * (size_t)P));
    PODSDF_CUDA(cudaMemcpy(w.p, w_host, sizeof(double) * (size_t)P, cudaMemcpyHostToDevice));
  }
  PODSDF_TRY(podsdf_gram_dev(device, A.p, M, N, layout, w_host ? w.p : nullptr, P, scale, C.p, nullptr));
  PODSDF_CUDA(cudaMemcpy(C_host, C.p, sizeof(double) * (size_t)N * N, cudaMemcpyDeviceToHost));
  return 0;
}

int podsdf_eig_sym_dev(int device, double* C_dev_inout_V, long N, double* lambda_dev, void* stream) {
  if (!C_dev_inout_V || !lambda_dev || N < 1) return fail(PODSDF_ERR_INVALID, "bad argument");
  PODSDF_TRY(check_device(device));
  DeviceGuard guard(device);
  return eig_inplace(device, C_dev_inout_V, N, lambda_dev, (cudaStream_t)stream);
}

int podsdf_eig_sym_top_dev(int device, double* C_dev_inout_V, long N, int nm, double* lambda_dev, int* path,
                           void* stream) {
  if (!C_dev_inout_V || !lambda_dev || N < 1) return fail(PODSDF_ERR_INVALID, "bad argument");
  PODSDF_TRY(check_device(device));
  DeviceGuard guard(device);
  int route = 0;
  const int rc = eig_top_inplace(device, C_dev_inout_V, N, nm, lambda_dev, (cudaStream_t)stream, &route);
  if (path) *path = route;
  return rc;
}

int podsdf_eig_sym_top_host(int device, const double* C_host, long N, int nm, double* lambda_host, double* V_host,
                            int* path) {
  if (!C_host || !lambda_host || !V_host || N < 1 || nm < 1 || nm > N) return fail(PODSDF_ERR_INVALID, "bad argument");
  PODSDF_TRY(check_device(device));
  DeviceGuard guard(device);
  DevBuf C, lam;
  const size_t nn = sizeof(double) * (size_t)N * N;
  PODSDF_CUDA(C.alloc(nn));
  PODSDF_CUDA(lam.alloc(sizeof(double) * N));
  PODSDF_CUDA(cudaMemcpy(C.p, C_host, nn, cudaMemcpyHostToDevice));
  int route = 0;
  PODSDF_TRY(eig_top_inplace(device, C.p, N, nm, lam.p, 0, &route));
  if (path) *path = route;
  PODSDF_CUDA(cudaMemcpy(lambda_host, lam.p, sizeof(double) * N, cudaMemcpyDeviceToHost));
  PODSDF_CUDA(cudaMemcpy(V_host, C.p, sizeof(double) * (size_t)nm * N, cudaMemcpyDeviceToHost));
  return 0;
}

int podsdf_eig_sym_host(int device, const double* C_host, long N, double* lambda_host, double* V_host) {
  if (!C_host || !lambda_host || !V_host || N < 1) return fail(PODSDF_ERR_INVALID, "bad argument");
  PODSDF_TRY(check_device(device));
  DeviceGuard guard(device);
  DevBuf C, V, lam;
  const size_t nn = sizeof(double) * (size_t)N * N;
  PODSDF_CUDA(C.alloc(nn));
  PODSDF_CUDA(V.alloc(nn));
  PODSDF_CUDA(lam.alloc(sizeof(double) * (size_t)N));
  PODSDF_CUDA(cudaMemcpy(C.p, C_host, nn, cudaMemcpyHostToDevice));
  PODSDF_TRY(eig_inplace(device, C.p, N, lam.p, nullptr));
  PODSDF_TRY(launch_transpose(C.p, V.p, N, nullptr));  // rows = vectors  ->  columns = vectors
  PODSDF_CUDA(cudaMemcpy(V_host, V.p, nn, cudaMemcpyDeviceToHost));
  PODSDF_CUDA(cudaMemcpy(lambda_host, lam.p, sizeof(double) * (size_t)N, cudaMemcpyDeviceToHost));
  return 0;
}

int podsdf_sym_tridiag_host(int device, const double* C_host, long N, double* d_host, double* e_host, double* band_host,
                            double* ms2) {
  if (!C_host || !d_host || !e_host || N < 3) return fail(PODSDF_ERR_INVALID, "bad argument");
  PODSDF_TRY(check_device(device));
  DeviceGuard guard(device);
  std::lock_guard<std::recursive_mutex> lk(g_solver_mu);     // the look-ahead stream and its events are per device
  DevBuf C, ws, de, band;
  const size_t nn = sizeof(double) * (size_t)N * N;
  PODSDF_CUDA(C.alloc(nn));
  PODSDF_CUDA(ws.alloc(sizeof(double) * sbr_workspace_doubles(N)));
  PODSDF_CUDA(de.alloc(sizeof(double) * 2 * (size_t)N));
  if (band_host) {      // stage 1 alone on a first copy
    PODSDF_CUDA(band.alloc(sizeof(double) * (size_t)N * kSbrLdb));
    PODSDF_CUDA(cudaMemcpy(C.p, C_host, nn, cudaMemcpyHostToDevice));
    PODSDF_TRY(sbr_stage1_only(C.p, N, ws.p, band.p, nullptr));
    std::vector<double> hb((size_t)N * kSbrLdb);
    PODSDF_CUDA(cudaMemcpy(hb.data(), band.p, sizeof(double) * hb.size(), cudaMemcpyDeviceToHost));
    for (long c = 0; c < N; ++c)
      for (int dd = 0; dd <= kSbrB; ++dd) band_host[c * (kSbrB + 1) + dd] = hb[(size_t)c * kSbrLdb + dd];
  }
  PODSDF_CUDA(cudaMemcpy(C.p, C_host, nn, cudaMemcpyHostToDevice));
  float t[2] = {0, 0};
  PODSDF_TRY(sbr_tridiagonalize(C.p, N, de.p, de.p + N, ws.p, nullptr, t));
  if (ms2) {
    ms2[0] = t[0];
    ms2[1] = t[1];
  }
  PODSDF_CUDA(cudaMemcpy(d_host, de.p, sizeof(double) * (size_t)N, cudaMemcpyDeviceToHost));
  PODSDF_CUDA(cudaMemcpy(e_host, de.p + N, sizeof(double) * (size_t)(N - 1), cudaMemcpyDeviceToHost));
  return 0;
}

int podsdf_sym_pack_dev(int device, const double* C_dev, long N, double* packed_dev, void* stream) {
  if (!C_dev || !packed_dev || N < 1 || N > 65535) return fail(PODSDF_ERR_INVALID, "bad argument");
  PODSDF_TRY(check_device(device));
  DeviceGuard guard(device);
  return launch_sym_pack(C_dev, N, packed_dev, (cudaStream_t)stream);
}

int podsdf_sym_unpack_dev(int device, const double* packed_dev, long N, double* C_dev, void* stream) {
  if (!C_dev || !packed_dev || N < 1 || N > 65535) return fail(PODSDF_ERR_INVALID, "bad argument");
  PODSDF_TRY(check_device(device));
  DeviceGuard guard(device);
  return launch_sym_unpack(packed_dev, N, C_dev, (cudaStream_t)stream);
}

int podsdf_field_moments_dev(int device, const double* F_dev, long P, long N, int layout, double* mean_dev,
                             double* stress_dev, void* stream) {
  if (!F_dev || !mean_dev || !stress_dev || P < 1 || N < 1 || (layout != 0 && layout != 1))
    return fail(PODSDF_ERR_INVALID, "bad argument");
  PODSDF_TRY(check_device(device));
  DeviceGuard guard(device);
  return launch_field_moments(F_dev, P, N, layout, mean_dev, stress_dev, (cudaStream_t)stream);
}

int podsdf_lag_covariance_dev(int device, const double* F_dev, long J, long K, long N, int layout, const double* mean_dev,
                              int comp, const int* lags3_host, int nl, double* cov_host, void* stream) {
  if (!F_dev || !lags3_host || !cov_host || J < 1 || K < 1 || N < 1 || comp < 0 || comp > 2 || (layout != 0 && layout != 1))
    return fail(PODSDF_ERR_INVALID, "bad argument");
  PODSDF_TRY(check_device(device));
  DeviceGuard guard(device);
  DevBuf ws;
  PODSDF_CUDA(ws.alloc(sizeof(double) * lag_covariance_workspace_doubles()));
  double* out = ws.p + lag_covariance_workspace_doubles() - lag_covariance_max();
  cudaStream_t st = (cudaStream_t)stream;
  PODSDF_TRY(launch_lag_covariance(F_dev, J, K, N, layout, mean_dev, comp, lags3_host, nl, ws.p, out, st));
  PODSDF_CUDA(cudaMemcpyAsync(cov_host, out, sizeof(double) * nl, cudaMemcpyDeviceToHost, st));
  PODSDF_CUDA(cudaStreamSynchronize(st));
  return 0;
}

int podsdf_project_dev(int device, const double* A_dev, long M, long N, int layout, const double* T_dev, long ldT,
                       const double* dinv_dev, int nm, double scale, double* Phi_dev, void* stream) {
  if (!A_dev || !T_dev || !dinv_dev || !Phi_dev || M < 1 || N < 1 || nm < 1 || ldT < nm)
    return fail(PODSDF_ERR_INVALID, "bad argument");
  PODSDF_TRY(check_device(device));
  DeviceGuard guard(device);
  return launch_project(A_dev, M, N, layout, T_dev, ldT, dinv_dev, nm, scale, Phi_dev, (cudaStream_t)stream);
}

int podsdf_project_host(int device, const double* A_host, long M, long N, int layout, const double* T_host, long ldT,
                        const double* dinv_host, int nm, double scale, double* Phi_host) {
  if (!A_host || !T_host || !dinv_host || !Phi_host || M < 1 || N < 1 || nm < 1 || ldT < nm)
    return fail(PODSDF_ERR_INVALID, "bad argument");
  PODSDF_TRY(check_device(device));
  DeviceGuard guard(device);
  DevBuf A, T, d, Phi;
  PODSDF_CUDA(A.alloc(sizeof(double) * (size_t)M * N));
  PODSDF_CUDA(T.alloc(sizeof(double) * (size_t)N * ldT));
  PODSDF_CUDA(d.alloc(sizeof(double) * (size_t)nm));
  PODSDF_CUDA(Phi.alloc(sizeof(double) * (size_t)M * nm));
  PODSDF_CUDA(cudaMemcpy(A.p, A_host, sizeof(double) * (size_t)M * N, cudaMemcpyHostToDevice));
  PODSDF_CUDA(cudaMemcpy(T.p, T_host, sizeof(double) * (size_t)N * ldT, cudaMemcpyHostToDevice));
  PODSDF_CUDA(cudaMemcpy(d.p, dinv_host, sizeof(double) * (size_t)nm, cudaMemcpyHostToDevice));
  PODSDF_TRY(launch_project(A.p, M, N, layout, T.p, ldT, d.p, nm, scale, Phi.p, nullptr));
  PODSDF_CUDA(cudaMemcpy(Phi_host, Phi.p, sizeof(double) * (size_t)M * nm, cudaMemcpyDeviceToHost));
  return 0;
}

int podsdf_dft_modes_dev(int device, const double* a_dev, long N, long ld, int nm, double* c_dev, void* stream) {
  if (!a_dev || !c_dev || N < 1 || nm < 1 || ld < nm) return fail(PODSDF_ERR_INVALID, "bad argument");
  PODSDF_TRY(check_device(device));
  DeviceGuard guard(device);
  cudaStream_t st = (cudaStream_t)stream;
  // The workspace (twiddle table + intermediate) lives in the grow-only scratch: no allocation, no free and no
  // host-side wait per call (an allocate / synchronise / free here stalled the host for 0.2 - 0.4 s per PODFS step
  // once the process held tens of GB).  A call that re-uses the table on another stream waits for the previous one.
  static cudaEvent_t dft_done[64] = {};
  static std::mutex dft_mu;
  std::lock_guard<std::mutex> lk(dft_mu);
  cudaEvent_t& ev = dft_done[device & 63];
  if (!ev) PODSDF_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
  else PODSDF_CUDA(cudaStreamWaitEvent(st, ev, 0));
  double* tw = nullptr;
  PODSDF_TRY(scratch_get(device, 3, sizeof(double) * dft_workspace_doubles(N, nm), &tw));
  const int rc = launch_dft(a_dev, N, ld, nm, tw, c_dev, st);
  PODSDF_CUDA(cudaEventRecord(ev, st));
  return rc;
}

int podsdf_dft_modes_host(int device, const double* a_host, long N, long ld, int nm, double* c_host) {
  if (!a_host || !c_host || N < 1 || nm < 1 || ld < nm) return fail(PODSDF_ERR_INVALID, "bad argument");
  PODSDF_TRY(check_device(device));
  DeviceGuard guard(device);
  DevBuf a, c;
  PODSDF_CUDA(a.alloc(sizeof(double) * (size_t)N * ld));
  PODSDF_CUDA(c.alloc(sizeof(double) * 2 * (size_t)N * nm));
  PODSDF_CUDA(cudaMemcpy(a.p, a_host, sizeof(double) * (size_t)N * ld, cudaMemcpyHostToDevice));
  PODSDF_TRY(podsdf_dft_modes_dev(device, a.p, N, ld, nm, c.p, nullptr));
  PODSDF_CUDA(cudaMemcpy(c_host, c.p, sizeof(double) * 2 * (size_t)N * nm, cudaMemcpyDeviceToHost));
  return 0;
}

}  // extern "C"
