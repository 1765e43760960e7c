// C ABI of the digital-filter path (include/podsdf.h): plan object, host/device entry points.
#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#include "../../include/podsdf.h"
#include "df_kernels.cuh"
#include "philox.cuh"

namespace podsdf {

static thread_local std::string g_last_error;
static std::atomic<long long> g_launches{0};

void set_error(const std::string& msg) { g_last_error = msg; }
int fail(int code, const std::string& msg) {
  g_last_error = msg;
  return code;
}
int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
  g_last_error = std::string(cudaGetErrorName(e)) + ": " + cudaGetErrorString(e) + " in " + what + " (" + file +
                 ":" + std::to_string(line) + ")";
  const bool nodev = (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver);
  cudaGetLastError();  // clear sticky-less errors
  return nodev ? PODSDF_ERR_NO_DEVICE : PODSDF_ERR_CUDA;
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// A device is usable only if it is sm_100 (the library carries sm_100a SASS only).
int check_device(int device) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    cudaGetLastError();
    return fail(PODSDF_ERR_NO_DEVICE, "no CUDA device visible: libpodsdf has no CPU fallback");
  }
  if (device < 0 || device >= n) return fail(PODSDF_ERR_INVALID, "device index out of range");
  int major = 0;
  PODSDF_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
  if (major != 10)
    return fail(PODSDF_ERR_NO_DEVICE, "device is not sm_100 (B200); libpodsdf carries sm_100a code only");
  return 0;
}

}  // namespace podsdf

using namespace podsdf;

struct podsdf_df_plan {
  int device = 0;
  DfGeom g{};
  Taps bx{}, by{}, bz{};
  int impl = PODSDF_IMPL_AUTO;
  int chunk = 512;         // replaced at plan creation by default_chunk()
  int rng = PODSDF_RNG_U32;  // stream of the seeded entries (podsdf_df_plan_set_rng)
  // device state
  double* d_coef = nullptr;  // [12][P]
  unsigned mask = 0;
  bool have_coef = false;
  Rot rot;                   // constant post-rotation (off by default)
  double* d_T1 = nullptr;  // [chunk][3][Np]
  double* d_T2 = nullptr;  // generic path: [chunk][3][Jp][K]
  double* d_X = nullptr;   // host-injected / generic-generated staging: [chunk+2nfx][3][Np]
  double* d_out[2] = {nullptr, nullptr};  // host paths: [chunk][3][P]
  int cap_core = 0, cap_t2 = 0, cap_x = 0, cap_out[2] = {0, 0};   // planes each workspace currently holds
  cudaStream_t s_compute = nullptr, s_copy = nullptr;
  cudaEvent_t ev_done[2] = {nullptr, nullptr}, ev_copied[2] = {nullptr, nullptr};
  std::vector<cudaEvent_t> ev_pool;  // timing events of the last call: (t0,t1,t2) per chunk
  size_t ev_used = 0;
  cudaStream_t ev_stream = nullptr;
  double last_ms[2] = {0, 0};
};

namespace {

bool use_tuned_time(const podsdf_df_plan* p) {
  if (p->impl == PODSDF_IMPL_GENERIC) return false;
  return df_time_has_tuned(2 * p->g.nfx + 1);
}
// the tuned plane kernel reads the column-panel layout only the tuned time kernel writes, with
// 16-byte pieces: both kernels must exist for these widths and the padded row length must be even
bool use_tuned_plane(const podsdf_df_plan* p) {
  if (p->impl == PODSDF_IMPL_GENERIC) return false;
  return use_tuned_time(p) && (p->g.Kp % 2 == 0) &&
         df_plane_has_tuned(2 * p->g.nfy + 1, 2 * p->g.nfz + 1, df_transform_is_diagonal(p->have_coef, p->mask, p->rot));
}
PanelGeom panel_of(const podsdf_df_plan* p) { return use_tuned_plane(p) ? make_panel_geom(p->g) : PanelGeom(); }
// sized for either layout: which one is used can change when the coefficients (coupled / diagonal) change
size_t t1_doubles_per_plane(const podsdf_df_plan* p) {
  const PanelGeom pg = make_panel_geom(p->g);
  return 3 * (size_t)std::max<long>(pg.comp_stride, p->g.Np);
}

int free_workspace(podsdf_df_plan* p) {
  for (double** q : {&p->d_T1, &p->d_T2, &p->d_X, &p->d_out[0], &p->d_out[1]}) {
    if (*q) cudaFree(*q);
    *q = nullptr;
  }
  p->cap_core = p->cap_t2 = p->cap_x = p->cap_out[0] = p->cap_out[1] = 0;
  return 0;
}

// Default planes per pass: as many as fit a 12 GiB T1 workspace, at most 1280.  Longer passes amortise the 2*nfx
// warm-up slabs of the time filter (config 3: 14.6 ms per 2500 planes at 512, 14.0 ms from 840 up).
int default_chunk(const podsdf_df_plan* p) {
  const size_t per_plane = sizeof(double) * t1_doubles_per_plane(p);
  const size_t fit = ((size_t)12 << 30) / per_plane;
  const int c = (int)std::min<size_t>(fit, 1280) & ~7;
  return c < 8 ? 8 : c;
}

// workspaces grow on demand to `need` planes (= min(chunk, steps of the call)); cudaFree synchronises the device,
// so a buffer still read by an earlier asynchronous pass is never pulled away
static int grow(double** buf, int* cap, int need, size_t doubles_per_plane, size_t extra_planes = 0) {
  if (*buf && *cap >= need) return 0;
  if (*buf) cudaFree(*buf);
  *buf = nullptr;
  *cap = 0;
  PODSDF_CUDA(cudaMalloc(buf, sizeof(double) * doubles_per_plane * ((size_t)need + extra_planes)));
  *cap = need;
  return 0;
}
// T1 (+T2 for the generic plane path, allocated lazily: it is as large as T1)
int ensure_core_ws(podsdf_df_plan* p, int need) {
  PODSDF_TRY(grow(&p->d_T1, &p->cap_core, need, t1_doubles_per_plane(p)));
  if (!use_tuned_plane(p)) PODSDF_TRY(grow(&p->d_T2, &p->cap_t2, need, 3 * (size_t)p->g.Jp * p->g.K));
  return 0;
}
int ensure_x_ws(podsdf_df_plan* p, int need) {
  return grow(&p->d_X, &p->cap_x, need, 3 * (size_t)p->g.Np, 2 * (size_t)p->g.nfx);
}
int ensure_out_ws(podsdf_df_plan* p, int need) {
  for (int i = 0; i < 2; ++i) PODSDF_TRY(grow(&p->d_out[i], &p->cap_out[i], need, 3 * (size_t)p->g.P));
  return 0;
}

cudaEvent_t next_event(podsdf_df_plan* p) {
  if (p->ev_used == p->ev_pool.size()) {
    cudaEvent_t e;
    cudaEventCreate(&e);
    p->ev_pool.push_back(e);
  }
  return p->ev_pool[p->ev_used++];
}

// one chunk on `st`: X (injected, device) or Philox -> T1 -> out
int run_chunk(podsdf_df_plan* p, const double* X_dev, bool gen, uint64_t seed, long s_base, int nsteps,
              double* out_dev, cudaStream_t st) {
  const bool tt = use_tuned_time(p), tp = use_tuned_plane(p);
  if (p->impl == PODSDF_IMPL_TUNED && (!tt || !tp))
    return fail(PODSDF_ERR_UNSUPPORTED, "PODSDF_IMPL_TUNED requested but no tuned kernel for these filter widths");
  const double* X = X_dev;
  if (gen && !tt) {  // generic time filter needs materialised slabs
    PODSDF_TRY(ensure_x_ws(p, nsteps));
    PODSDF_TRY(launch_df_random_slabs(p->g, seed, p->rng, s_base, nsteps + 2 * p->g.nfx, p->d_X, st));
    X = p->d_X;
    gen = false;
  }
  cudaEvent_t e0 = next_event(p), e1 = next_event(p), e2 = next_event(p);
  PODSDF_CUDA(cudaEventRecord(e0, st));
  const PanelGeom pg = panel_of(p);
  const int src = !gen ? 0 : (p->rng == PODSDF_RNG_U32 ? 2 : 1);
  PODSDF_TRY(launch_df_time(p->g, pg, p->bx, X, p->d_T1, nsteps, src, seed, s_base, tt, st));
  PODSDF_CUDA(cudaEventRecord(e1, st));
  PODSDF_TRY(launch_df_plane(p->g, pg, p->by, p->bz, p->d_T1, out_dev, nsteps, 3, p->d_coef, p->mask, p->have_coef,
                             p->rot, p->d_T2, tp, st));
  PODSDF_CUDA(cudaEventRecord(e2, st));
  return 0;
}

}  // namespace

extern "C" {

const char* podsdf_last_error(void) { return g_last_error.c_str(); }
int podsdf_version(void) { return 200; }  // 0.2.0
long long podsdf_launch_count(void) { return g_launches.load(); }

int podsdf_device_count(int* count) {
  if (!count) return fail(PODSDF_ERR_INVALID, "count is null");
  *count = 0;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    cudaGetLastError();
    return fail(PODSDF_ERR_NO_DEVICE, "no CUDA device visible: libpodsdf has no CPU fallback");
  }
  int ok = 0;
  for (int d = 0; d < n; ++d) {
    int major = 0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, d) == cudaSuccess && major == 10) ++ok;
  }
  *count = ok;
  if (!ok) return fail(PODSDF_ERR_NO_DEVICE, "no sm_100 (B200) device visible");
  return 0;
}

int podsdf_host_alloc(void** ptr, size_t bytes) {
  if (!ptr) return fail(PODSDF_ERR_INVALID, "ptr is null");
  PODSDF_CUDA(cudaHostAlloc(ptr, bytes, cudaHostAllocDefault));
  return 0;
}
int podsdf_host_free(void* ptr) {
  if (ptr) PODSDF_CUDA(cudaFreeHost(ptr));
  return 0;
}

int podsdf_df_plan_create(int device, int J, int K, int nfx, int nfy, int nfz, const double* bx, const double* by,
                          const double* bz, podsdf_df_plan** plan) {
  if (!plan) return fail(PODSDF_ERR_INVALID, "plan is null");
  *plan = nullptr;
  if (J < 1 || K < 1 || nfx < 0 || nfy < 0 || nfz < 0) return fail(PODSDF_ERR_INVALID, "bad plane / filter size");
  if (2 * nfx + 1 > kMaxTaps || 2 * nfy + 1 > kMaxTaps || 2 * nfz + 1 > kMaxTaps)
    return fail(PODSDF_ERR_UNSUPPORTED, "half filter width above 64");
  if (!bx || !by || !bz) return fail(PODSDF_ERR_INVALID, "tap vectors are null");
  if ((long)J * K >= (1L << 31) / 4) return fail(PODSDF_ERR_UNSUPPORTED, "plane too large for 32-bit point indices");
  PODSDF_TRY(check_device(device));
  DeviceGuard guard(device);
  if (const char* v = getenv("PODSDF_TIME_VARIANT")) df_set_time_variant(atoi(v));   // profiling / A-B switch
  auto* p = new podsdf_df_plan();
  p->device = device;
  p->g = DfGeom{J, K, nfx, nfy, nfz, J + 2 * nfy, K + 2 * nfz, (long)(J + 2 * nfy) * (K + 2 * nfz), (long)J * K};
  std::memset(&p->bx, 0, sizeof(Taps));
  std::memset(&p->by, 0, sizeof(Taps));
  std::memset(&p->bz, 0, sizeof(Taps));
  std::copy(bx, bx + 2 * nfx + 1, p->bx.b);
  std::copy(by, by + 2 * nfy + 1, p->by.b);
  std::copy(bz, bz + 2 * nfz + 1, p->bz.b);
  cudaError_t e = cudaStreamCreateWithFlags(&p->s_compute, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&p->s_copy, cudaStreamNonBlocking);
  for (int i = 0; i < 2 && e == cudaSuccess; ++i) {
    e = cudaEventCreateWithFlags(&p->ev_done[i], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&p->ev_copied[i], cudaEventDisableTiming);
  }
  if (e != cudaSuccess) {
    podsdf_df_plan_destroy(p);
    return cuda_fail(e, "plan stream/event creation", __FILE__, __LINE__);
  }
  p->chunk = default_chunk(p);
  *plan = p;
  return 0;
}

int podsdf_df_plan_destroy(podsdf_df_plan* p) {
  if (!p) return 0;
  DeviceGuard guard(p->device);
  cudaDeviceSynchronize();
  free_workspace(p);
  if (p->d_coef) cudaFree(p->d_coef);
  for (cudaEvent_t e : p->ev_pool) cudaEventDestroy(e);
  for (int i = 0; i < 2; ++i) {
    if (p->ev_done[i]) cudaEventDestroy(p->ev_done[i]);
    if (p->ev_copied[i]) cudaEventDestroy(p->ev_copied[i]);
  }
  if (p->s_compute) cudaStreamDestroy(p->s_compute);
  if (p->s_copy) cudaStreamDestroy(p->s_copy);
  delete p;
  return 0;
}

int podsdf_df_plan_set_impl(podsdf_df_plan* p, int impl) {
  if (!p || impl < 0 || impl > 2) return fail(PODSDF_ERR_INVALID, "bad plan / impl");
  DeviceGuard guard(p->device);
  cudaDeviceSynchronize();
  free_workspace(p);
  p->impl = impl;
  return 0;
}

int podsdf_df_plan_set_chunk(podsdf_df_plan* p, int chunk) {
  if (!p || chunk < 1 || chunk > 65535) return fail(PODSDF_ERR_INVALID, "chunk must be in 1..65535");
  // multiples of 8: the tensor-path time filter works in 8-step tiles anchored at the start of every pass, so
  // results are bitwise independent of the chunking (and of any time sharding whose offsets are multiples of 8)
  p->chunk = chunk < 8 ? 8 : (chunk & ~7);
  return 0;
}

int podsdf_df_plan_set_rng(podsdf_df_plan* p, int kind) {
  if (!p || (kind != PODSDF_RNG_U53 && kind != PODSDF_RNG_U32)) return fail(PODSDF_ERR_INVALID, "bad plan / rng kind");
  p->rng = kind;
  return 0;
}

int podsdf_df_set_coefficients(podsdf_df_plan* p, const double* coef_host, unsigned mask) {
  if (!p) return fail(PODSDF_ERR_INVALID, "plan is null");
  DeviceGuard guard(p->device);
  if (!coef_host) {
    p->have_coef = false;
    p->mask = 0;
    return 0;
  }
  if (mask >= (1u << 12)) return fail(PODSDF_ERR_INVALID, "mask has bits above 11");
  const long P = p->g.P;
  if (!p->d_coef) PODSDF_CUDA(cudaMalloc(&p->d_coef, sizeof(double) * 12 * P));
  PODSDF_CUDA(cudaStreamSynchronize(p->s_compute));
  for (int i = 0; i < 12; ++i)
    if (mask & (1u << i))
      PODSDF_CUDA(cudaMemcpy(p->d_coef + i * P, coef_host + i * P, sizeof(double) * P, cudaMemcpyHostToDevice));
  p->mask = mask;
  p->have_coef = true;
  return 0;
}

int podsdf_df_set_rotation(podsdf_df_plan* p, const double* R9) {
  if (!p) return fail(PODSDF_ERR_INVALID, "plan is null");
  p->rot = Rot();
  if (R9) {
    for (int i = 0; i < 9; ++i) p->rot.r[i] = R9[i];
    p->rot.on = 1;
  }
  return 0;
}

int podsdf_df_filter_block_host(podsdf_df_plan* p, const double* x_host, double* y_host) {
  if (!p || !x_host || !y_host) return fail(PODSDF_ERR_INVALID, "null argument");
  DeviceGuard guard(p->device);
  const DfGeom& g = p->g;
  const int nt = 2 * g.nfx + 1;
  // one component, one plane: the generic kernels (X is [s][1][Np] here, so use a 1-comp geometry walk)
  double *dX = nullptr, *dT1 = nullptr, *dT2 = nullptr, *dY = nullptr;
  cudaStream_t st = p->s_compute;
  int rc = 0;
  auto cleanup = [&]() {
    for (double* q : {dX, dT1, dT2, dY})
      if (q) cudaFree(q);
  };
  // layout trick: replicate the block into the 3-component slab layout expected by the time kernel
  PODSDF_CUDA(cudaMalloc(&dX, sizeof(double) * 3 * g.Np * nt));
  cudaError_t e = cudaMalloc(&dT1, sizeof(double) * t1_doubles_per_plane(p));
  if (e == cudaSuccess) e = cudaMalloc(&dT2, sizeof(double) * 3 * (size_t)g.Jp * g.K);
  if (e == cudaSuccess) e = cudaMalloc(&dY, sizeof(double) * 3 * g.P);
  if (e == cudaSuccess) e = cudaMemsetAsync(dX, 0, sizeof(double) * 3 * g.Np * nt, st);
  for (int s = 0; s < nt && e == cudaSuccess; ++s)
    e = cudaMemcpyAsync(dX + (size_t)s * 3 * g.Np, x_host + (size_t)s * g.Np, sizeof(double) * g.Np,
                        cudaMemcpyHostToDevice, st);
  if (e != cudaSuccess) {
    cleanup();
    return cuda_fail(e, "filter_block staging", __FILE__, __LINE__);
  }
  const bool tt = use_tuned_time(p), tp = use_tuned_plane(p);
  const PanelGeom pg = panel_of(p);
  rc = launch_df_time(g, pg, p->bx, dX, dT1, 1, 0, 0, 0, tt, st);
  if (!rc) rc = launch_df_plane(g, pg, p->by, p->bz, dT1, dY, 1, 3, nullptr, 0, false, Rot(), dT2, tp, st);
  if (!rc) {
    e = cudaMemcpyAsync(y_host, dY, sizeof(double) * g.P, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) rc = cuda_fail(e, "filter_block readback", __FILE__, __LINE__);
  }
  cleanup();
  return rc;
}

int podsdf_df_lund_apply_host(podsdf_df_plan* p, double* yu, double* yv, double* yw) {
  if (!p || !yu || !yv || !yw) return fail(PODSDF_ERR_INVALID, "null argument");
  if (!p->have_coef && !p->rot.on) return 0;  // identity
  DeviceGuard guard(p->device);
  const long P = p->g.P;
  double* d = nullptr;
  PODSDF_CUDA(cudaMalloc(&d, sizeof(double) * 3 * P));
  cudaStream_t st = p->s_compute;
  double* h[3] = {yu, yv, yw};
  cudaError_t e = cudaSuccess;
  for (int c = 0; c < 3 && e == cudaSuccess; ++c)
    e = cudaMemcpyAsync(d + c * P, h[c], sizeof(double) * P, cudaMemcpyHostToDevice, st);
  int rc = 0;
  if (e == cudaSuccess) rc = launch_df_lund(p->g, d, p->d_coef, p->mask, p->have_coef, p->rot, st);
  for (int c = 0; c < 3 && e == cudaSuccess && !rc; ++c)
    e = cudaMemcpyAsync(h[c], d + c * P, sizeof(double) * P, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  cudaFree(d);
  if (e != cudaSuccess) return cuda_fail(e, "lund_apply", __FILE__, __LINE__);
  return rc;
}

int podsdf_df_generate_injected_dev(podsdf_df_plan* p, const double* slabs_dev, long n_slabs, double* out_dev,
                                    void* stream) {
  if (!p || !slabs_dev || !out_dev) return fail(PODSDF_ERR_INVALID, "null argument");
  const long nsteps = n_slabs - 2L * p->g.nfx;
  if (nsteps < 1) return fail(PODSDF_ERR_INVALID, "need at least 2*nfx+1 slabs");
  DeviceGuard guard(p->device);
  PODSDF_TRY(ensure_core_ws(p, (int)std::min<long>(p->chunk, nsteps)));
  cudaStream_t st = (cudaStream_t)stream;
  p->ev_used = 0;
  p->ev_stream = st;
  for (long t = 0; t < nsteps; t += p->chunk) {
    const int n = (int)std::min<long>(p->chunk, nsteps - t);
    PODSDF_TRY(run_chunk(p, slabs_dev + t * 3 * p->g.Np, false, 0, 0, n, out_dev + t * 3 * p->g.P, st));
  }
  return 0;
}

int podsdf_df_generate_seeded_dev(podsdf_df_plan* p, uint64_t seed, long t0, long nsteps, double* out_dev,
                                  void* stream) {
  if (!p || !out_dev) return fail(PODSDF_ERR_INVALID, "null argument");
  if (nsteps < 1 || t0 < 0) return fail(PODSDF_ERR_INVALID, "bad step range");
  DeviceGuard guard(p->device);
  PODSDF_TRY(ensure_core_ws(p, (int)std::min<long>(p->chunk, nsteps)));
  cudaStream_t st = (cudaStream_t)stream;
  p->ev_used = 0;
  p->ev_stream = st;
  for (long t = 0; t < nsteps; t += p->chunk) {
    const int n = (int)std::min<long>(p->chunk, nsteps - t);
    PODSDF_TRY(run_chunk(p, nullptr, true, seed, t0 + t, n, out_dev + t * 3 * p->g.P, st));
  }
  return 0;
}

// host paths: chunked, double-buffered output so the D2H of chunk i overlaps the kernels of chunk i+1
// (truly asynchronous when out_host is pinned, e.g. from podsdf_host_alloc)
static int generate_host(podsdf_df_plan* p, const double* slabs_host, bool gen, uint64_t seed, long t0, long nsteps,
                         double* out_host) {
  DeviceGuard guard(p->device);
  // the host link is ~100x slower than the kernels: short passes (<= 256 planes) start the first copy early and
  // keep both staging buffers small; the extra warm-up slabs of the time filter hide behind the copies
  const int hchunk = std::min(p->chunk, 256);
  const int need = (int)std::min<long>(hchunk, nsteps);
  PODSDF_TRY(ensure_core_ws(p, need));
  PODSDF_TRY(ensure_out_ws(p, need));
  if (!gen) PODSDF_TRY(ensure_x_ws(p, need));
  const DfGeom& g = p->g;
  p->ev_used = 0;
  p->ev_stream = p->s_compute;
  int i = 0;
  for (long t = 0; t < nsteps; t += hchunk, ++i) {
    const int n = (int)std::min<long>(hchunk, nsteps - t);
    const int b = i & 1;
    if (i >= 2) PODSDF_CUDA(cudaStreamWaitEvent(p->s_compute, p->ev_copied[b], 0));
    if (!gen)
      PODSDF_CUDA(cudaMemcpyAsync(p->d_X, slabs_host + t * 3 * g.Np, sizeof(double) * 3 * g.Np * (n + 2 * g.nfx),
                                  cudaMemcpyHostToDevice, p->s_compute));
    PODSDF_TRY(run_chunk(p, p->d_X, gen, seed, t0 + t, n, p->d_out[b], p->s_compute));
    PODSDF_CUDA(cudaEventRecord(p->ev_done[b], p->s_compute));
    PODSDF_CUDA(cudaStreamWaitEvent(p->s_copy, p->ev_done[b], 0));
    PODSDF_CUDA(cudaMemcpyAsync(out_host + t * 3 * g.P, p->d_out[b], sizeof(double) * 3 * g.P * n,
                                cudaMemcpyDeviceToHost, p->s_copy));
    PODSDF_CUDA(cudaEventRecord(p->ev_copied[b], p->s_copy));
  }
  PODSDF_CUDA(cudaStreamSynchronize(p->s_copy));
  PODSDF_CUDA(cudaStreamSynchronize(p->s_compute));
  return 0;
}

int podsdf_df_generate_injected_host(podsdf_df_plan* p, const double* slabs_host, long n_slabs, double* out_host) {
  if (!p || !slabs_host || !out_host) return fail(PODSDF_ERR_INVALID, "null argument");
  const long nsteps = n_slabs - 2L * p->g.nfx;
  if (nsteps < 1) return fail(PODSDF_ERR_INVALID, "need at least 2*nfx+1 slabs");
  return generate_host(p, slabs_host, false, 0, 0, nsteps, out_host);
}

int podsdf_df_generate_seeded_host(podsdf_df_plan* p, uint64_t seed, long t0, long nsteps, double* out_host) {
  if (!p || !out_host) return fail(PODSDF_ERR_INVALID, "null argument");
  if (nsteps < 1 || t0 < 0) return fail(PODSDF_ERR_INVALID, "bad step range");
  return generate_host(p, nullptr, true, seed, t0, nsteps, out_host);
}

int podsdf_df_random_slabs_dev(podsdf_df_plan* p, uint64_t seed, long s0, long n, double* slabs_dev, void* stream) {
  if (!p || !slabs_dev || n < 0 || s0 < 0) return fail(PODSDF_ERR_INVALID, "bad argument");
  DeviceGuard guard(p->device);
  return launch_df_random_slabs(p->g, seed, p->rng, s0, n, slabs_dev, (cudaStream_t)stream);
}

int podsdf_df_random_slabs_host(podsdf_df_plan* p, uint64_t seed, long s0, long n, double* slabs_host) {
  if (!p || !slabs_host || n < 0 || s0 < 0) return fail(PODSDF_ERR_INVALID, "bad argument");
  if (n == 0) return 0;
  DeviceGuard guard(p->device);
  const size_t per = sizeof(double) * 3 * p->g.Np;
  const long chunk = std::max<long>(1, std::min<long>(n, (256L << 20) / (long)per));
  double* d = nullptr;
  PODSDF_CUDA(cudaMalloc(&d, per * chunk));
  int rc = 0;
  for (long s = 0; s < n && !rc; s += chunk) {
    const long c = std::min(chunk, n - s);
    rc = launch_df_random_slabs(p->g, seed, p->rng, s0 + s, c, d, p->s_compute);
    if (!rc) {
      cudaError_t e = cudaMemcpyAsync(slabs_host + s * 3 * p->g.Np, d, per * c, cudaMemcpyDeviceToHost, p->s_compute);
      if (e == cudaSuccess) e = cudaStreamSynchronize(p->s_compute);
      if (e != cudaSuccess) rc = cuda_fail(e, "random_slabs readback", __FILE__, __LINE__);
    }
  }
  cudaFree(d);
  return rc;
}

int podsdf_df_last_kernel_ms(podsdf_df_plan* p, double* ms2) {
  if (!p || !ms2) return fail(PODSDF_ERR_INVALID, "null argument");
  DeviceGuard guard(p->device);
  double a = 0, b = 0;
  for (size_t i = 0; i + 3 <= p->ev_used; i += 3) {
    PODSDF_CUDA(cudaEventSynchronize(p->ev_pool[i + 2]));
    float m1 = 0, m2 = 0;
    PODSDF_CUDA(cudaEventElapsedTime(&m1, p->ev_pool[i], p->ev_pool[i + 1]));
    PODSDF_CUDA(cudaEventElapsedTime(&m2, p->ev_pool[i + 1], p->ev_pool[i + 2]));
    a += m1;
    b += m2;
  }
  ms2[0] = a;
  ms2[1] = b;
  return 0;
}

int podsdf_df_plan_info(podsdf_df_plan* p, int* info5) {
  if (!p || !info5) return fail(PODSDF_ERR_INVALID, "null argument");
  info5[0] = p->chunk;
  info5[1] = use_tuned_time(p) ? 1 : 0;
  info5[2] = use_tuned_plane(p) ? 1 : 0;
  info5[3] = df_transform_is_diagonal(p->have_coef, p->mask, p->rot) ? 1 : 0;
  info5[4] = (int)(p->ev_used / 3);
  return 0;
}

}  // extern "C"
