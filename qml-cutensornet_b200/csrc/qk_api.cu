// C ABI of libqkmps.so (declared in include/qkmps.h).  Thin host glue: argument checks, device
// memory, launches.  No CPU compute path: every compute entry point needs a CUDA device.
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <mutex>
#include <string>
#include <vector>
#include "../../include/qkmps.h"
#include "qk_kernels.cuh"
#include "qk_plan.h"
#include "qk_sim_core.h"

static thread_local std::string g_err;
static thread_local long long* g_gram_clk = nullptr;
static thread_local int64_t g_gram_clk_cap = 0, g_gram_clk_used = 0;
static int fail(int code, const std::string& msg) { g_err = msg; return code; }
static int cuda_fail(cudaError_t e, const char* what) {
  g_err = std::string(what) + ": " + cudaGetErrorString(e);
  return QK_ERR_CUDA;
}
#define QK_CUDA(call, what)                         \
  do {                                              \
    cudaError_t _e = (call);                        \
    if (_e != cudaSuccess) return cuda_fail(_e, what); \
  } while (0)

// ---- small caching pool for device blocks: a Gram job allocates the same few large buffers every
// call (state store, scratch); cudaMalloc/cudaFree of hundreds of MB would otherwise dominate the
// host side.  Every API call synchronises its stream before returning, so a released block is idle.
struct PoolBlock { void* p; size_t bytes; int device; };
static std::vector<PoolBlock> g_pool_free;
static std::vector<PoolBlock> g_pool_live;
static size_t g_pool_cached = 0;
static const size_t kPoolMaxCached = (size_t)8 << 30;
static std::mutex g_pool_mu;

static cudaError_t pool_alloc(void** out, size_t bytes) {
  std::lock_guard<std::mutex> lock(g_pool_mu);
  int dev = 0;
  cudaGetDevice(&dev);
  if (bytes == 0) bytes = 16;
  int best = -1;
  for (size_t i = 0; i < g_pool_free.size(); ++i) {
    const PoolBlock& b = g_pool_free[i];
    if (b.device == dev && b.bytes >= bytes && b.bytes <= 2 * bytes + 4096 && (best < 0 || b.bytes < g_pool_free[best].bytes))
      best = (int)i;
  }
  if (best >= 0) {
    PoolBlock b = g_pool_free[best];
    g_pool_free.erase(g_pool_free.begin() + best);
    g_pool_cached -= b.bytes;
    g_pool_live.push_back(b);
    *out = b.p;
    return cudaSuccess;
  }
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, bytes);
  if (e != cudaSuccess) {   // drop the cache and retry once
    for (auto& b : g_pool_free) cudaFree(b.p);
    g_pool_free.clear(); g_pool_cached = 0;
    cudaGetLastError();
    e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) return e;
  }
  g_pool_live.push_back({p, bytes, dev});
  *out = p;
  return cudaSuccess;
}

static void pool_free(void* p) {
  if (!p) return;
  std::lock_guard<std::mutex> lock(g_pool_mu);
  for (size_t i = 0; i < g_pool_live.size(); ++i) {
    if (g_pool_live[i].p == p) {
      PoolBlock b = g_pool_live[i];
      g_pool_live.erase(g_pool_live.begin() + i);
      if (g_pool_cached + b.bytes <= kPoolMaxCached && g_pool_free.size() < 64) {
        g_pool_free.push_back(b);
        g_pool_cached += b.bytes;
      } else {
        cudaFree(p);
      }
      return;
    }
  }
  cudaFree(p);
}
template <typename T>
static cudaError_t pool_alloc_t(T** out, size_t bytes) { return pool_alloc((void**)out, bytes); }

// Stream-ordered release for the asynchronous entry points: the block goes back to the pool only once everything
// queued on `stream` up to now has finished (an event is recorded; pool_reap() polls it).
struct DeferredFree { void* p; cudaEvent_t ev; };
static std::vector<DeferredFree> g_deferred;
static std::mutex g_deferred_mu;

static void pool_reap(bool wait) {
  std::vector<void*> done;
  {
    std::lock_guard<std::mutex> lock(g_deferred_mu);
    for (size_t i = 0; i < g_deferred.size();) {
      cudaError_t q = wait ? cudaEventSynchronize(g_deferred[i].ev) : cudaEventQuery(g_deferred[i].ev);
      if (q == cudaSuccess || q != cudaErrorNotReady) {
        cudaEventDestroy(g_deferred[i].ev);
        done.push_back(g_deferred[i].p);
        g_deferred[i] = g_deferred.back();
        g_deferred.pop_back();
      } else {
        ++i;
      }
    }
    cudaGetLastError();
  }
  for (void* p : done) pool_free(p);
}

static void pool_free_after(void* p, cudaStream_t stream) {
  if (!p) return;
  cudaEvent_t ev = nullptr;
  if (cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess || cudaEventRecord(ev, stream) != cudaSuccess) {
    if (ev) cudaEventDestroy(ev);
    cudaStreamSynchronize(stream);
    pool_free(p);
    return;
  }
  {
    std::lock_guard<std::mutex> lock(g_deferred_mu);
    g_deferred.push_back({p, ev});
  }
  pool_reap(false);
}

struct qk_batch {
  int device = 0;
  int n = 0, N = 0;
  int chi_cap = 1;
  std::vector<int32_t> cap;          // [n+1]
  std::vector<int64_t> site_off;     // [n+1] c128 units
  int64_t state_stride = 0;
  c128* store = nullptr;             // device [N][state_stride]
  int32_t* chi = nullptr;            // device [N][n+1]
  QkStat* stats = nullptr;           // device [N]
  int64_t* site_off_dev = nullptr;   // device [n+1]
  float sim_ms = 0.f;
  int sim_grid = 0;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;   // around the stage-1 kernel (asynchronous launches: read lazily)
  cudaStream_t stream = nullptr;
  long long* unit_clk = nullptr;              // device [N] clock64 ticks per datapoint (per-unit timing)
};

extern "C" {

int qk_version(void) { return QKMPS_VERSION; }
const char* qk_last_error(void) { return g_err.c_str(); }

int qk_device_count(int* count) {
  if (!count) return fail(QK_ERR_ARG, "count is NULL");
  int c = 0;
  cudaError_t e = cudaGetDeviceCount(&c);
  if (e != cudaSuccess) { *count = 0; return cuda_fail(e, "cudaGetDeviceCount"); }
  *count = c;
  return QK_OK;
}

// ---------------------------------------------------------------- plan
int qk_plan_create_gates(int n_qubits, const qk_gate* gates, int n_gates, int trunc_mode, double trunc_error,
                         int chi_cap, int flags, qk_plan** out) {
  if (!out) return fail(QK_ERR_ARG, "out is NULL");
  *out = nullptr;
  qk_plan* p = new qk_plan();
  std::string err;
  int rc = qk_compile_plan(n_qubits, gates, n_gates, trunc_mode, trunc_error, chi_cap, flags, p, &err);
  if (rc != QK_OK) { delete p; return fail(rc, err); }
  *out = p;
  return QK_OK;
}

int qk_plan_create_ansatz(int n_qubits, int reps, double gamma, int hadamard_init, const int32_t* pairs, int n_pairs,
                          int trunc_mode, double trunc_error, int chi_cap, int flags, qk_plan** out) {
  if (!out) return fail(QK_ERR_ARG, "out is NULL");
  *out = nullptr;
  if (n_pairs > 0 && !pairs) return fail(QK_ERR_ARG, "pairs is NULL");
  std::vector<qk_gate> gates;
  std::string err;
  int rc = qk_ansatz_gates(n_qubits, reps, gamma, hadamard_init, pairs, n_pairs, &gates, &err);
  if (rc != QK_OK) return fail(rc, err);
  return qk_plan_create_gates(n_qubits, gates.data(), (int)gates.size(), trunc_mode, trunc_error, chi_cap, flags, out);
}

int qk_plan_info(const qk_plan* plan, qk_plan_info_t* info) {
  if (!plan || !info) return fail(QK_ERR_ARG, "NULL argument");
  info->n_qubits = plan->n; info->n_gates = plan->n_gates; info->n_ops = (int)plan->ops.size();
  info->n_ops_2q = plan->n_2q; info->n_ops_1q = plan->n_1q; info->n_moves = plan->n_moves;
  info->chi_cap = plan->chi_cap; info->threads = plan->threads; info->trunc_mode = plan->trunc_mode;
  info->smem_bytes = (int32_t)plan->smem_bytes; info->state_stride = plan->state_stride;
  info->trunc_error = plan->trunc_error;
  return QK_OK;
}

int qk_plan_ops(const qk_plan* plan, qk_op_view* ops, int max_ops) {
  if (!plan || (max_ops > 0 && !ops)) return fail(QK_ERR_ARG, "NULL argument");
  int k = std::min<int>(max_ops, (int)plan->ops.size());
  for (int i = 0; i < k; ++i) {
    const QkOp& o = plan->ops[i];
    ops[i].kind = o.kind; ops[i].site = o.site; ops[i].fa = o.fa; ops[i].fb = o.fb; ops[i].dir = o.dir; ops[i].coeff = o.coeff;
  }
  return k;
}

void qk_plan_destroy(qk_plan* plan) { delete plan; }

// ---------------------------------------------------------------- batches
void qk_batch_destroy(qk_batch* b) {
  if (!b) return;
  int prev = 0;
  cudaGetDevice(&prev);
  cudaSetDevice(b->device);
  if (b->ev1) cudaEventSynchronize(b->ev1);   // the stage-1 kernel may still be running (asynchronous launch)
  if (b->ev0) cudaEventDestroy(b->ev0);
  if (b->ev1) cudaEventDestroy(b->ev1);
  pool_free(b->store); pool_free(b->chi); pool_free(b->stats); pool_free(b->site_off_dev); pool_free(b->unit_clk);
  cudaSetDevice(prev);
  delete b;
}

static int batch_alloc(qk_batch* b) {
  QK_CUDA(cudaSetDevice(b->device), "cudaSetDevice");
  const size_t nstate = (size_t)std::max(b->N, 1);
  QK_CUDA(pool_alloc_t(&b->store, nstate * b->state_stride * sizeof(c128)), "cudaMalloc(store)");
  QK_CUDA(pool_alloc_t(&b->chi, nstate * (b->n + 1) * sizeof(int32_t)), "cudaMalloc(chi)");
  QK_CUDA(pool_alloc_t(&b->stats, nstate * sizeof(QkStat)), "cudaMalloc(stats)");
  QK_CUDA(pool_alloc_t(&b->site_off_dev, (b->n + 1) * sizeof(int64_t)), "cudaMalloc(site_off)");
  QK_CUDA(cudaMemcpy(b->site_off_dev, b->site_off.data(), (b->n + 1) * sizeof(int64_t), cudaMemcpyHostToDevice),
          "cudaMemcpy(site_off)");
  return QK_OK;
}

static thread_local double* g_trace_dev = nullptr;   // set only inside qk_simulate_trace

// Stage 1 launch.  sync = false: nothing waits for the kernel; the batch carries the events, every scratch block is
// released in stream order (pool_free_after), and results must only be used on `stream` or after qk_batch_sim_ms.
static int simulate_impl(const qk_plan* plan, int device, cudaStream_t stream, const double* X_dev, int N, int ldx,
                         qk_batch** out, bool sync) {
  if (!plan || !out) return fail(QK_ERR_ARG, "NULL argument");
  *out = nullptr;
  if (N < 0 || (N > 0 && !X_dev) || ldx < plan->n) return fail(QK_ERR_ARG, "bad X / N / ldx (ldx must be >= n_qubits)");
  qk_batch* b = new qk_batch();
  b->device = device; b->n = plan->n; b->N = N; b->chi_cap = plan->chi_cap; b->stream = stream;
  b->cap = plan->cap; b->site_off = plan->site_off; b->state_stride = plan->state_stride;
  int rc = batch_alloc(b);
  if (rc != QK_OK) { qk_batch_destroy(b); return rc; }
  if (N == 0) { *out = b; return QK_OK; }
  pool_reap(false);

  std::vector<void*> scratch;
  auto cleanup = [&]() { for (void* p : scratch) pool_free_after(p, stream); scratch.clear(); };
  auto bail = [&](cudaError_t e, const char* what) { cleanup(); cudaStreamSynchronize(stream); qk_batch_destroy(b); return cuda_fail(e, what); };
  auto salloc = [&](void** ptr, size_t bytes) { cudaError_t e = pool_alloc(ptr, bytes); if (e == cudaSuccess) scratch.push_back(*ptr); return e; };
  cudaError_t e;
  QkOp* ops_dev = nullptr; int32_t* cap_dev = nullptr; int* counter = nullptr;
  const size_t nops = plan->ops.size();
  if ((e = salloc((void**)&ops_dev, std::max<size_t>(nops, 1) * sizeof(QkOp))) != cudaSuccess) return bail(e, "cudaMalloc(ops)");
  if ((e = salloc((void**)&cap_dev, (plan->n + 1) * sizeof(int32_t))) != cudaSuccess) return bail(e, "cudaMalloc(cap)");
  if ((e = salloc((void**)&counter, sizeof(int))) != cudaSuccess) return bail(e, "cudaMalloc(counter)");
  if ((e = pool_alloc_t(&b->unit_clk, (size_t)N * sizeof(long long))) != cudaSuccess) return bail(e, "cudaMalloc(unit clocks)");
  if ((e = cudaMemsetAsync(b->unit_clk, 0, (size_t)N * sizeof(long long), stream)) != cudaSuccess) return bail(e, "memset");
  if ((e = cudaMemcpyAsync(ops_dev, plan->ops.data(), nops * sizeof(QkOp), cudaMemcpyHostToDevice, stream)) != cudaSuccess) return bail(e, "copy ops");
  if ((e = cudaMemcpyAsync(cap_dev, plan->cap.data(), (plan->n + 1) * sizeof(int32_t), cudaMemcpyHostToDevice, stream)) != cudaSuccess) return bail(e, "copy cap");

  SimParams P;
  P.n = plan->n; P.n_ops = (int)nops; P.ops = ops_dev; P.cap = cap_dev; P.site_off = b->site_off_dev;
  P.state_stride = plan->state_stride; P.X = X_dev; P.ldx = ldx; P.N = N;
  P.store = b->store; P.chi = b->chi; P.stats = b->stats;
  P.mode = plan->trunc_mode; P.cutoff = plan->trunc_error; P.fidelity_target = 1.0 - plan->trunc_error;
  P.value_of_zero = 1e-16;   // pytket-cutensornet Config default
  P.tol = 1e-15; P.max_sweeps = 60; P.rmax = plan->rmax; P.wr = plan->rmax * plan->rmax;
  P.trace = g_trace_dev;
  P.early_exit = plan->early_exit;
  P.floor_rel = 1e-28;
  P.abs_rel = 0.0;
  P.parallel = plan->parallel; P.lam = nullptr; P.lam_ld = plan->rmax / 2; P.level_start = nullptr; P.n_levels = 0;
  P.big_w = nullptr; P.big_s = nullptr; P.big_w_stride = P.big_s_stride = 0; P.big_flag = nullptr; P.big_jb = plan->jb;
  P.big_wb_entries = (int)qk_big_wb_entries(plan->rmax, plan->jb);
  P.unit_clk = b->unit_clk;

  if ((e = cudaEventCreate(&b->ev0)) != cudaSuccess) return bail(e, "cudaEventCreate");
  if ((e = cudaEventCreate(&b->ev1)) != cudaSuccess) return bail(e, "cudaEventCreate");
  if ((e = cudaEventRecord(b->ev0, stream)) != cudaSuccess) return bail(e, "cudaEventRecord");
  const char* what = "stage-1 kernel";
  if (plan->big) {
    // large-matrix path: theta and the staging area of every resident cluster live in global memory
    what = "stage-1 kernel (large-matrix path)";
    int ncta_req = 0, ncta = 1, n_clusters = 1;
    if (const char* ev = getenv("QK_BIG_CLUSTER")) { const int v = atoi(ev); if (v >= 1 && v <= 16) ncta_req = v; }
    c128 *w_dev = nullptr, *s_dev = nullptr; int* f_dev = nullptr;
    e = qk_sim_big_config(plan->smem_bytes, N, ncta_req, &ncta, &n_clusters);
    const size_t w_stride = (size_t)plan->rmax * plan->rmax, s_stride = w_stride / 2;
    if (e == cudaSuccess) e = salloc((void**)&w_dev, (size_t)n_clusters * w_stride * sizeof(c128));
    if (e == cudaSuccess) e = salloc((void**)&s_dev, (size_t)n_clusters * s_stride * sizeof(c128));
    if (e == cudaSuccess) e = salloc((void**)&f_dev, (size_t)n_clusters * 4 * sizeof(int));
    if (e == cudaSuccess) e = cudaMemsetAsync(f_dev, 0, (size_t)n_clusters * 4 * sizeof(int), stream);
    if (e == cudaSuccess) {
      P.big_w = w_dev; P.big_s = s_dev; P.big_w_stride = (int64_t)w_stride; P.big_s_stride = (int64_t)s_stride; P.big_flag = f_dev;
      e = qk_launch_sim_big(P, plan->smem_bytes, ncta, n_clusters, stream);
      b->sim_grid = n_clusters * ncta;
    }
  } else if (plan->parallel) {
    what = "stage-1 kernel (B form)";
    // CTAs per datapoint: 8 for small batches, 4 for mid-sized ones (measured, profiles/r02_stage1_schedules.txt);
    // QK_SIM_CLUSTER overrides (experiments)
    int ncta = N <= 150 ? 8 : 4;
    if (const char* ev = getenv("QK_SIM_CLUSTER")) { const int v = atoi(ev); if (v >= 1 && v <= 8) ncta = v; }
    double* lam_dev = nullptr; int32_t* lvl_dev = nullptr; QkStat* parts_dev = nullptr;
    e = salloc((void**)&lam_dev, (size_t)N * (plan->n + 1) * P.lam_ld * sizeof(double));
    if (e == cudaSuccess) e = salloc((void**)&lvl_dev, plan->level_start.size() * sizeof(int32_t));
    if (e == cudaSuccess) e = salloc((void**)&parts_dev, (size_t)N * ncta * sizeof(QkStat));
    if (e == cudaSuccess)
      e = cudaMemcpyAsync(lvl_dev, plan->level_start.data(), plan->level_start.size() * sizeof(int32_t),
                          cudaMemcpyHostToDevice, stream);
    if (e == cudaSuccess) {
      P.lam = lam_dev; P.level_start = lvl_dev; P.n_levels = (int)plan->level_start.size() - 1;
      e = qk_launch_sim_b(P, plan->threads, plan->smem_bytes, ncta, parts_dev, stream, &b->sim_grid);
    }
  } else {
    e = qk_launch_sim(P, plan->threads, plan->smem_bytes, counter, stream, &b->sim_grid);
  }
  if (e == cudaSuccess) e = cudaEventRecord(b->ev1, stream);
  if (e != cudaSuccess) return bail(e, what);
  cleanup();
  if (sync) {
    e = cudaEventSynchronize(b->ev1);
    if (e != cudaSuccess) { qk_batch_destroy(b); return cuda_fail(e, what); }
    cudaEventElapsedTime(&b->sim_ms, b->ev0, b->ev1);
  }
  *out = b;
  return QK_OK;
}

int qk_simulate_dev(const qk_plan* plan, int device, void* stream_v, const double* X_dev, int N, int ldx,
                    qk_batch** out) {
  return simulate_impl(plan, device, (cudaStream_t)stream_v, X_dev, N, ldx, out, true);
}

int qk_simulate_async(const qk_plan* plan, int device, void* stream_v, const double* X_dev, int N, int ldx,
                      qk_batch** out) {
  return simulate_impl(plan, device, (cudaStream_t)stream_v, X_dev, N, ldx, out, false);
}

int qk_simulate(const qk_plan* plan, int device, const double* X_host, int N, int ldx, qk_batch** out) {
  if (!plan || !out) return fail(QK_ERR_ARG, "NULL argument");
  *out = nullptr;
  if (N < 0 || (N > 0 && !X_host) || ldx < plan->n) return fail(QK_ERR_ARG, "bad X / N / ldx (ldx must be >= n_qubits)");
  QK_CUDA(cudaSetDevice(device), "cudaSetDevice");
  double* X_dev = nullptr;
  const size_t bytes = (size_t)std::max(N, 1) * ldx * sizeof(double);
  QK_CUDA(pool_alloc_t(&X_dev, bytes), "cudaMalloc(X)");
  cudaError_t e = cudaMemcpy(X_dev, X_host, (size_t)N * ldx * sizeof(double), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) { pool_free(X_dev); return cuda_fail(e, "cudaMemcpy(X)"); }
  int rc = qk_simulate_dev(plan, device, nullptr, X_dev, N, ldx, out);
  pool_free(X_dev);
  return rc;
}

int qk_simulate_trace(const qk_plan* plan, int device, const double* x_host, int ldx, double* bytes_per_op,
                      int max_ops, qk_batch** out) {
  if (!plan || !x_host || !bytes_per_op || !out) return fail(QK_ERR_ARG, "NULL argument");
  const int nops = (int)plan->ops.size();
  if (max_ops < nops) return fail(QK_ERR_ARG, "trace buffer smaller than the number of ops");
  QK_CUDA(cudaSetDevice(device), "cudaSetDevice");
  double* tr = nullptr;
  QK_CUDA(pool_alloc_t(&tr, std::max<size_t>(nops, 1) * sizeof(double)), "cudaMalloc(trace)");
  g_trace_dev = tr;
  int rc = qk_simulate(plan, device, x_host, 1, ldx, out);
  g_trace_dev = nullptr;
  if (rc == QK_OK) {
    cudaError_t e = cudaMemcpy(bytes_per_op, tr, nops * sizeof(double), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) rc = cuda_fail(e, "copy trace");
  }
  pool_free(tr);
  return rc;
}

// Hands the working store (the bulk of a batch's memory) back in stream order once it has been packed / exchanged;
// bond dimensions, statistics and per-datapoint clocks stay readable.
int qk_batch_release_store(qk_batch* b, void* stream_v) {
  if (!b) return fail(QK_ERR_ARG, "NULL batch");
  if (b->store) {
    int prev = 0;
    cudaGetDevice(&prev);
    cudaSetDevice(b->device);
    pool_free_after(b->store, (cudaStream_t)stream_v);
    cudaSetDevice(prev);
    b->store = nullptr;
  }
  return QK_OK;
}

int qk_batch_sim_ms(const qk_batch* b, float* ms) {
  if (!b || !ms) return fail(QK_ERR_ARG, "NULL argument");
  *ms = b->sim_ms;
  if (b->ev1 && b->sim_ms == 0.f && b->N > 0) {   // asynchronous launch: wait for it now
    QK_CUDA(cudaEventSynchronize(b->ev1), "stage-1 kernel");
    float t = 0.f;
    cudaEventElapsedTime(&t, b->ev0, b->ev1);
    const_cast<qk_batch*>(b)->sim_ms = t;
    *ms = t;
  }
  return QK_OK;
}

static int batch_flags_or(const qk_batch* b, int* flags_or);
// OR of the stage-1 flags of every state (QK_FLAG_*): the cheap correctness read-back after an asynchronous launch
int qk_batch_flags(const qk_batch* b, int32_t* flags_or) {
  if (!b || !flags_or) return fail(QK_ERR_ARG, "NULL argument");
  int f = 0;
  int rc = batch_flags_or(b, &f);
  *flags_or = f;
  return rc;
}

int qk_batch_unit_seconds(const qk_batch* b, double* seconds) {
  if (!b || !seconds) return fail(QK_ERR_ARG, "NULL argument");
  if (b->N == 0) return QK_OK;
  QK_CUDA(cudaSetDevice(b->device), "cudaSetDevice");
  if (b->ev1) QK_CUDA(cudaEventSynchronize(b->ev1), "stage-1 kernel");
  std::vector<long long> clk(b->N, 0);
  if (b->unit_clk) QK_CUDA(cudaMemcpy(clk.data(), b->unit_clk, clk.size() * sizeof(long long), cudaMemcpyDeviceToHost), "copy clocks");
  // cudaDevAttrClockRate is a slow query (milliseconds; worse while nvidia-smi polls the device): ask once per device
  static int khz_cache[64] = {0};
  int khz = (b->device >= 0 && b->device < 64) ? khz_cache[b->device] : 0;
  if (khz == 0) {
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, b->device);
    if (b->device >= 0 && b->device < 64) khz_cache[b->device] = khz;
  }
  const double hz = khz > 0 ? 1e3 * khz : 1.965e9;
  for (int i = 0; i < b->N; ++i) seconds[i] = (double)clk[i] / hz;
  return QK_OK;
}

int qk_batch_size(const qk_batch* b, int* N, int* n_qubits) {
  if (!b) return fail(QK_ERR_ARG, "NULL batch");
  if (N) *N = b->N;
  if (n_qubits) *n_qubits = b->n;
  return QK_OK;
}

int qk_batch_info(const qk_batch* b, int32_t* chi, double* fidelity, double* trunc_weight, int64_t* nbytes,
                  int32_t* flags, int32_t* sweeps) {
  if (!b) return fail(QK_ERR_ARG, "NULL batch");
  if (b->N == 0) return QK_OK;
  QK_CUDA(cudaSetDevice(b->device), "cudaSetDevice");
  std::vector<int32_t> h_chi((size_t)b->N * (b->n + 1));
  QK_CUDA(cudaMemcpy(h_chi.data(), b->chi, h_chi.size() * sizeof(int32_t), cudaMemcpyDeviceToHost), "copy chi");
  if (chi) memcpy(chi, h_chi.data(), h_chi.size() * sizeof(int32_t));
  if (nbytes) {
    for (int i = 0; i < b->N; ++i) {
      int64_t t = 0;
      const int32_t* c = &h_chi[(size_t)i * (b->n + 1)];
      for (int s = 0; s < b->n; ++s) t += (int64_t)c[s] * 2 * c[s + 1] * 16;   // sum of tensors[k].nbytes (gpu:295)
      nbytes[i] = t;
    }
  }
  if (fidelity || trunc_weight || flags || sweeps) {
    std::vector<QkStat> st(b->N);
    QK_CUDA(cudaMemcpy(st.data(), b->stats, st.size() * sizeof(QkStat), cudaMemcpyDeviceToHost), "copy stats");
    for (int i = 0; i < b->N; ++i) {
      if (fidelity) fidelity[i] = st[i].fidelity;
      if (trunc_weight) trunc_weight[i] = st[i].trunc_weight;
      if (flags) flags[i] = st[i].flags;
      if (sweeps) sweeps[i] = st[i].sweeps;
    }
  }
  return QK_OK;
}

int qk_batch_export(const qk_batch* b, int i, void* host_buf, int64_t buf_bytes) {
  if (!b || !host_buf) return fail(QK_ERR_ARG, "NULL argument");
  if (i < 0 || i >= b->N) return fail(QK_ERR_ARG, "state index out of range");
  if (!b->store) return fail(QK_ERR_ARG, "the batch's store was released");
  QK_CUDA(cudaSetDevice(b->device), "cudaSetDevice");
  std::vector<int32_t> c(b->n + 1);
  QK_CUDA(cudaMemcpy(c.data(), b->chi + (size_t)i * (b->n + 1), c.size() * sizeof(int32_t), cudaMemcpyDeviceToHost), "copy chi");
  int64_t need = 0;
  for (int s = 0; s < b->n; ++s) need += (int64_t)c[s] * 2 * c[s + 1] * 16;
  if (buf_bytes < need) return fail(QK_ERR_ARG, "export buffer too small");
  char* dst = (char*)host_buf;
  for (int s = 0; s < b->n; ++s) {
    const size_t bytes = (size_t)c[s] * 2 * c[s + 1] * 16;
    QK_CUDA(cudaMemcpy(dst, b->store + (size_t)i * b->state_stride + b->site_off[s], bytes, cudaMemcpyDeviceToHost), "copy site");
    dst += bytes;
  }
  return QK_OK;
}

int qk_batch_import(int device, int n_qubits, int N, const int32_t* chi, const void* host_tensors, int64_t bytes,
                    qk_batch** out) {
  if (!out) return fail(QK_ERR_ARG, "out is NULL");
  *out = nullptr;
  if (n_qubits < 1 || N < 1 || !chi || !host_tensors) return fail(QK_ERR_ARG, "bad arguments");
  qk_batch* b = new qk_batch();
  b->device = device; b->n = n_qubits; b->N = N;
  b->cap.assign(n_qubits + 1, 1);
  for (int i = 0; i < N; ++i)
    for (int s = 0; s <= n_qubits; ++s) {
      const int32_t c = chi[(size_t)i * (n_qubits + 1) + s];
      if (c < 1) { delete b; return fail(QK_ERR_ARG, "bond dimensions must be >= 1"); }
      b->cap[s] = std::max(b->cap[s], c);
    }
  b->chi_cap = *std::max_element(b->cap.begin(), b->cap.end());
  b->site_off.assign(n_qubits + 1, 0);
  for (int s = 0; s < n_qubits; ++s) b->site_off[s + 1] = b->site_off[s] + (int64_t)b->cap[s] * 2 * b->cap[s + 1];
  b->state_stride = b->site_off[n_qubits];
  int rc = batch_alloc(b);
  if (rc != QK_OK) { qk_batch_destroy(b); return rc; }
  std::vector<c128> host((size_t)N * b->state_stride);
  const char* src = (const char*)host_tensors;
  int64_t used = 0;
  for (int i = 0; i < N; ++i)
    for (int s = 0; s < n_qubits; ++s) {
      const int32_t* c = chi + (size_t)i * (n_qubits + 1);
      const int64_t sz = (int64_t)c[s] * 2 * c[s + 1] * 16;
      if (used + sz > bytes) { qk_batch_destroy(b); return fail(QK_ERR_ARG, "tensor buffer too small"); }
      memcpy(&host[(size_t)i * b->state_stride + b->site_off[s]], src + used, sz);
      used += sz;
    }
  cudaError_t e = cudaMemcpy(b->store, host.data(), host.size() * sizeof(c128), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(b->chi, chi, (size_t)N * (n_qubits + 1) * sizeof(int32_t), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemset(b->stats, 0, (size_t)N * sizeof(QkStat));
  if (e != cudaSuccess) { qk_batch_destroy(b); return cuda_fail(e, "upload"); }
  *out = b;
  return QK_OK;
}

int qk_batch_max_chi(const qk_batch* b, int32_t* max_chi) {
  if (!b || !max_chi) return fail(QK_ERR_ARG, "NULL argument");
  for (int s = 0; s <= b->n; ++s) max_chi[s] = 1;
  if (b->N == 0) return QK_OK;
  QK_CUDA(cudaSetDevice(b->device), "cudaSetDevice");
  std::vector<int32_t> h((size_t)b->N * (b->n + 1));
  QK_CUDA(cudaMemcpy(h.data(), b->chi, h.size() * sizeof(int32_t), cudaMemcpyDeviceToHost), "copy chi");
  for (int i = 0; i < b->N; ++i)
    for (int s = 0; s <= b->n; ++s) max_chi[s] = std::max(max_chi[s], h[(size_t)i * (b->n + 1) + s]);
  return QK_OK;
}

// ---------------------------------------------------------------- frag exchange format
static int check_D(int n, const int32_t* D) {
  if (!D) return fail(QK_ERR_ARG, "D is NULL");
  for (int s = 0; s <= n; ++s)
    if (D[s] < 8 || (D[s] & 7)) return fail(QK_ERR_ARG, "padded bond dimensions must be positive multiples of 8");
  return QK_OK;
}

int qk_frag_stride(int n_qubits, const int32_t* D, int64_t* bytes_per_state) {
  if (n_qubits < 1 || !bytes_per_state) return fail(QK_ERR_ARG, "bad arguments");
  int rc = check_D(n_qubits, D);
  if (rc != QK_OK) return rc;
  FragLayout L;
  qk_frag_layout(n_qubits, D, &L, nullptr);
  *bytes_per_state = L.stride_bytes;
  return QK_OK;
}

int qk_batch_pack_scatter(const qk_batch* b, const int32_t* D, void* frag_dev, const int32_t* dst_index,
                          void* stream_v) {
  if (!b || !frag_dev) return fail(QK_ERR_ARG, "NULL argument");
  int rc = check_D(b->n, D);
  if (rc != QK_OK) return rc;
  if (b->N == 0) return QK_OK;
  if (!b->store) return fail(QK_ERR_ARG, "the batch's store was released");
  QK_CUDA(cudaSetDevice(b->device), "cudaSetDevice");
  cudaStream_t stream = (cudaStream_t)stream_v;
  // bond dimensions of the states that are actually packed must fit the padded dims
  std::vector<int32_t> h((size_t)b->N * (b->n + 1));
  QK_CUDA(cudaMemcpy(h.data(), b->chi, h.size() * sizeof(int32_t), cudaMemcpyDeviceToHost), "copy chi");
  for (int i = 0; i < b->N; ++i) {
    if (dst_index && dst_index[i] < 0) continue;
    for (int s = 0; s <= b->n; ++s)
      if (h[(size_t)i * (b->n + 1) + s] > D[s])
        return fail(QK_ERR_ARG, "padded bond dimension smaller than a state's bond dimension");
  }
  FragLayout L;
  std::vector<int64_t> off(b->n + 1);
  qk_frag_layout(b->n, D, &L, off.data());
  int32_t* D_dev = nullptr; int64_t* off_dev = nullptr; int32_t* dst_dev = nullptr;
  QK_CUDA(pool_alloc_t(&D_dev, (b->n + 1) * sizeof(int32_t)), "cudaMalloc(D)");
  cudaError_t e = pool_alloc_t(&off_dev, (b->n + 1) * sizeof(int64_t));
  if (e == cudaSuccess && dst_index) e = pool_alloc_t(&dst_dev, (size_t)b->N * sizeof(int32_t));
  if (e == cudaSuccess) e = cudaMemcpyAsync(D_dev, D, (b->n + 1) * sizeof(int32_t), cudaMemcpyHostToDevice, stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(off_dev, off.data(), (b->n + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, stream);
  if (e == cudaSuccess && dst_index)
    e = cudaMemcpyAsync(dst_dev, dst_index, (size_t)b->N * sizeof(int32_t), cudaMemcpyHostToDevice, stream);
  if (e == cudaSuccess)
    e = qk_launch_pack(b->n, b->N, b->store, b->state_stride, b->site_off_dev, b->chi, D_dev, off_dev, L.stride_bytes,
                       L.data_bytes, frag_dev, dst_dev, 0, stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
  pool_free(D_dev); pool_free(off_dev); pool_free(dst_dev);
  if (e != cudaSuccess) return cuda_fail(e, "pack kernel");
  return QK_OK;
}

// Asynchronous pack: state i of the batch goes to position first_index + i of the frag buffer.  No host
// synchronisation and no check of the bond dimensions against D (the caller passes dims that cover the plan's caps).
int qk_batch_pack_async(const qk_batch* b, const int32_t* D, void* frag_dev, int first_index, void* stream_v) {
  if (!b || !frag_dev || first_index < 0) return fail(QK_ERR_ARG, "bad arguments");
  int rc = check_D(b->n, D);
  if (rc != QK_OK) return rc;
  if (b->N == 0) return QK_OK;
  if (!b->store) return fail(QK_ERR_ARG, "the batch's store was released");
  for (int s = 0; s <= b->n; ++s)
    if (D[s] < b->cap[s]) return fail(QK_ERR_ARG, "asynchronous pack needs padded dimensions >= the batch's bond caps");
  QK_CUDA(cudaSetDevice(b->device), "cudaSetDevice");
  cudaStream_t stream = (cudaStream_t)stream_v;
  FragLayout L;
  std::vector<int64_t> off(b->n + 1);
  qk_frag_layout(b->n, D, &L, off.data());
  const size_t nb = (size_t)b->n + 1;
  const size_t o_off = (nb * sizeof(int32_t) + 15) & ~(size_t)15;
  std::vector<unsigned char> h(o_off + nb * sizeof(int64_t));
  memcpy(h.data(), D, nb * sizeof(int32_t));
  memcpy(h.data() + o_off, off.data(), nb * sizeof(int64_t));
  unsigned char* d = nullptr;
  QK_CUDA(pool_alloc_t(&d, h.size()), "cudaMalloc(pack scratch)");
  cudaError_t e = cudaMemcpyAsync(d, h.data(), h.size(), cudaMemcpyHostToDevice, stream);
  if (e == cudaSuccess)
    e = qk_launch_pack(b->n, b->N, b->store, b->state_stride, b->site_off_dev, b->chi, (const int32_t*)d,
                       (const int64_t*)(d + o_off), L.stride_bytes, L.data_bytes, frag_dev, nullptr, first_index, stream);
  pool_free_after(d, stream);
  if (e != cudaSuccess) return cuda_fail(e, "pack kernel");
  return QK_OK;
}

int qk_batch_pack(const qk_batch* b, const int32_t* D, void* frag_dev, void* stream_v) {
  return qk_batch_pack_scatter(b, D, frag_dev, nullptr, stream_v);
}

// ---------------------------------------------------------------- stage 2
int qk_gram_frags(int device, void* stream_v, int n_qubits, const int32_t* Dx, const void* fragX, int Nx,
                  const int32_t* Dy, const void* fragY, int Ny, const int32_t* tiles, int n_tiles, int symmetric,
                  double* K_dev, int64_t ldk, float* ms_out) {
  if (n_qubits < 1 || !fragX || Nx < 1 || !K_dev || n_tiles < 0 || (n_tiles > 0 && !tiles))
    return fail(QK_ERR_ARG, "bad arguments");
  if (symmetric) { Dy = Dx; fragY = fragX; Ny = Nx; }
  if (!fragY || Ny < 1) return fail(QK_ERR_ARG, "bad Y arguments");
  int rc = check_D(n_qubits, Dx);
  if (rc == QK_OK) rc = check_D(n_qubits, Dy);
  if (rc != QK_OK) return rc;
  if (ldk < Nx) return fail(QK_ERR_ARG, "ldk must be >= Nx");
  int maxD = 8;
  for (int s = 0; s <= n_qubits; ++s) maxD = std::max(maxD, std::max(Dx[s], Dy[s]));
  if (maxD > 32) return fail(QK_ERR_LIMIT, "overlap kernels support padded bond dimensions <= 32");
  const bool generic = maxD > 16;   // above the register-resident tensor-core kernel: CUDA-core kernel, same buffers
  QK_CUDA(cudaSetDevice(device), "cudaSetDevice");
  cudaStream_t stream = (cudaStream_t)stream_v;

  int TI = 0, TJ = 0;
  qk_gram_dmma_tile_shape(maxD, &TI, &TJ);
  // L2 blocking: CTAs are scheduled in list order, so the pair tiles are emitted supertile by supertile,
  // a supertile being S x S states whose packed bras + kets fit in about half of the 126 MB L2; every state
  // is then fetched from HBM once per supertile instead of once per 8-pair tile.
  int S = 8;
  {
    FragLayout Ltmp;
    qk_frag_layout(n_qubits, Dx, &Ltmp, nullptr);
    int64_t per_state = Ltmp.stride_bytes;
    qk_frag_layout(n_qubits, Dy, &Ltmp, nullptr);
    per_state = std::max<int64_t>(per_state, Ltmp.stride_bytes);
    int64_t budget = (int64_t)32 << 20;    // bytes of kets (and as many of bras) per supertile
    if (const char* ev = getenv("QK_GRAM_SUPERTILE_MB")) { const int v = atoi(ev); if (v > 0) budget = (int64_t)v << 20; }
    S = (int)std::max<int64_t>(8, budget / std::max<int64_t>(per_state, 1));
    S = std::max(8, S / 8 * 8);
  }
  std::vector<int4> cta;
  for (int t = 0; t < n_tiles; ++t) {
    const int r0 = tiles[4 * t], r1 = tiles[4 * t + 1], c0 = tiles[4 * t + 2], c1 = tiles[4 * t + 3];
    if (r0 < 0 || c0 < 0 || r1 > Ny || c1 > Nx || r0 > r1 || c0 > c1) return fail(QK_ERR_ARG, "tile out of range");
    for (int sy = r0; sy < r1; sy += S)
      for (int sx = c0; sx < c1; sx += S) {
        const int sye = std::min(sy + S, r1), sxe = std::min(sx + S, c1);
        if (symmetric && sx > sye - 1) continue;
        for (int y0 = sy; y0 < sye; y0 += TJ)
          for (int x0 = sx; x0 < sxe; x0 += TI) {
            const int ye = std::min(y0 + TJ, sye), xe = std::min(x0 + TI, sxe);
            if (symmetric && x0 > ye - 1) continue;   // whole sub-tile above the diagonal
            cta.push_back(make_int4(y0, x0, ye, xe));
          }
      }
  }
  if (cta.empty()) { if (ms_out) *ms_out = 0.f; return QK_OK; }

  FragLayout Lx, Ly;
  std::vector<int64_t> offx(n_qubits + 1), offy(n_qubits + 1);
  qk_frag_layout(n_qubits, Dx, &Lx, offx.data());
  qk_frag_layout(n_qubits, Dy, &Ly, offy.data());
  int slot_x = 0, slot_y = 0;
  for (int s = 0; s < n_qubits; ++s) {
    slot_x = std::max(slot_x, Dx[s] * Dx[s + 1] * 32);
    slot_y = std::max(slot_y, Dy[s] * Dy[s + 1] * 32);
  }

  // one device scratch block: Dx, Dy, offx, offy, tiles
  const size_t nb = (size_t)(n_qubits + 1);
  const size_t bytes_i = 2 * nb * sizeof(int32_t), bytes_o = 2 * nb * sizeof(int64_t), bytes_t = cta.size() * sizeof(int4);
  const size_t o_off = (bytes_i + 15) & ~(size_t)15, t_off = (o_off + bytes_o + 15) & ~(size_t)15;
  std::vector<unsigned char> hbuf(t_off + bytes_t);
  memcpy(hbuf.data(), Dx, nb * sizeof(int32_t));
  memcpy(hbuf.data() + nb * sizeof(int32_t), Dy, nb * sizeof(int32_t));
  memcpy(hbuf.data() + o_off, offx.data(), nb * sizeof(int64_t));
  memcpy(hbuf.data() + o_off + nb * sizeof(int64_t), offy.data(), nb * sizeof(int64_t));
  memcpy(hbuf.data() + t_off, cta.data(), bytes_t);
  std::vector<int2> pair_list;
  if (generic) {
    for (const int4& t4 : cta)
      for (int y = t4.x; y < t4.z; ++y)
        for (int x = t4.y; x < t4.w; ++x)
          if (!symmetric || x <= y) pair_list.push_back(make_int2(y, x));
  }
  int2* pairs_dev = nullptr;
  unsigned char* dbuf = nullptr;
  QK_CUDA(pool_alloc_t(&dbuf, hbuf.size()), "cudaMalloc(gram scratch)");
  if (generic) {
    cudaError_t ep = pool_alloc_t(&pairs_dev, std::max<size_t>(pair_list.size(), 1) * sizeof(int2));
    if (ep == cudaSuccess)
      ep = cudaMemcpyAsync(pairs_dev, pair_list.data(), pair_list.size() * sizeof(int2), cudaMemcpyHostToDevice, stream);
    if (ep != cudaSuccess) { pool_free(dbuf); pool_free(pairs_dev); return cuda_fail(ep, "upload pair list"); }
  }
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  cudaError_t e = cudaMemcpyAsync(dbuf, hbuf.data(), hbuf.size(), cudaMemcpyHostToDevice, stream);
  GramParams P;
  P.n = n_qubits;
  P.Dx = (const int32_t*)dbuf; P.Dy = P.Dx + nb;
  P.offx = (const int64_t*)(dbuf + o_off); P.offy = P.offx + nb;
  P.fragX = (const unsigned char*)fragX; P.fragY = (const unsigned char*)fragY;
  P.strideX = Lx.stride_bytes; P.strideY = Ly.stride_bytes; P.dataX = Lx.data_bytes; P.dataY = Ly.data_bytes;
  P.Nx = Nx; P.Ny = Ny;
  P.tiles = (const int4*)(dbuf + t_off); P.n_cta_tiles = (int)cta.size();
  P.symmetric = symmetric ? 1 : 0;
  P.K = K_dev; P.ldk = ldk; P.slot_x = slot_x; P.slot_y = slot_y;
  P.unit_clk = g_gram_clk;
  if (g_gram_clk_cap < (int64_t)cta.size()) P.unit_clk = nullptr;
  g_gram_clk_used = P.unit_clk ? (int64_t)cta.size() : 0;
  // ms_out == NULL: asynchronous -- nothing waits for the kernel, scratch is released in stream order
  const bool sync = (ms_out != nullptr);
  if (sync) {
    if (e == cudaSuccess) e = cudaEventCreate(&e0);
    if (e == cudaSuccess) e = cudaEventCreate(&e1);
    if (e == cudaSuccess) e = cudaEventRecord(e0, stream);
  }
  if (e == cudaSuccess)
    e = generic ? qk_launch_gram_frag_generic(P, pairs_dev, (int)pair_list.size(), maxD, stream)
                : qk_launch_gram_dmma(P, maxD, stream);
  float ms = 0.f;
  if (sync) {
    if (e == cudaSuccess) e = cudaEventRecord(e1, stream);
    if (e == cudaSuccess) e = cudaEventSynchronize(e1);
    if (e == cudaSuccess) cudaEventElapsedTime(&ms, e0, e1);
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    pool_free(dbuf);
    pool_free(pairs_dev);
  } else {
    pool_free_after(dbuf, stream);
    pool_free_after(pairs_dev, stream);
  }
  if (e != cudaSuccess) return cuda_fail(e, "stage-2 kernel");
  if (ms_out) *ms_out = ms;
  return QK_OK;
}

// Optional per-tile clocks of the next qk_gram_frags call on this thread (per-unit timing of the inner products,
// reference gpu:379-381): clk_dev receives clock64 ticks per CTA tile (8 pairs), at most `capacity` entries.
int qk_gram_set_tile_clocks(long long* clk_dev, int64_t capacity) {
  g_gram_clk = clk_dev;
  g_gram_clk_cap = clk_dev ? capacity : 0;
  g_gram_clk_used = 0;
  return QK_OK;
}
int64_t qk_gram_tile_clocks_used(void) { return g_gram_clk_used; }

static int gram_big_run(int device, cudaStream_t stream, int n, const int64_t* soff_x, const int64_t* soff_y,
                        const int32_t* dims_x, const int32_t* dims_y, const c128* storeX, int64_t strideX,
                        const int32_t* chiX, int Nx, const c128* storeY, int64_t strideY, const int32_t* chiY, int Ny,
                        const int32_t* tiles, int n_tiles, int symmetric, double* K_dev, int64_t ldk, float* ms_out);

int qk_gram_store(const qk_batch* X, const qk_batch* Y, double* K_host, int64_t ldk, float* ms_out) {
  if (!X || !K_host) return fail(QK_ERR_ARG, "NULL argument");
  if (!Y) Y = X;
  if (X->n != Y->n || X->device != Y->device) return fail(QK_ERR_ARG, "batches do not match");
  if ((X->N && !X->store) || (Y->N && !Y->store)) return fail(QK_ERR_ARG, "a batch's store was released");
  if (ldk < X->N) return fail(QK_ERR_ARG, "ldk must be >= Nx");
  if (X->N == 0 || Y->N == 0) return QK_OK;
  QK_CUDA(cudaSetDevice(X->device), "cudaSetDevice");
  double* K_dev = nullptr;
  QK_CUDA(pool_alloc_t(&K_dev, (size_t)Y->N * X->N * sizeof(double)), "cudaMalloc(K)");
  if (std::max(X->chi_cap, Y->chi_cap) > 32) {
    // E and T of a pair no longer fit in shared memory: batched-GEMM sweep on the same stores
    const int32_t tile[4] = {0, Y->N, 0, X->N};
    float ms = 0.f;
    int rc = gram_big_run(X->device, nullptr, X->n, X->site_off.data(), Y->site_off.data(), X->cap.data(), Y->cap.data(),
                          X->store, X->state_stride, X->chi, X->N, Y->store, Y->state_stride, Y->chi, Y->N, tile, 1, 0,
                          K_dev, X->N, &ms);
    if (rc == QK_OK) {
      cudaError_t ec = cudaMemcpy2D(K_host, ldk * sizeof(double), K_dev, X->N * sizeof(double), X->N * sizeof(double), Y->N,
                                    cudaMemcpyDeviceToHost);
      if (ec != cudaSuccess) rc = cuda_fail(ec, "cudaMemcpy2D(K)");
    }
    std::string keep = g_err;
    pool_free(K_dev);
    g_err = keep;
    if (ms_out) *ms_out = ms;
    return rc;
  }
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0, 0);
  cudaError_t e = qk_launch_gram_store(X->n, X->store, X->state_stride, X->site_off_dev, X->chi, X->chi_cap, X->N,
                                       Y->store, Y->state_stride, Y->site_off_dev, Y->chi, Y->chi_cap, Y->N, K_dev,
                                       X->N, 0);
  cudaEventRecord(e1, 0);
  if (e == cudaSuccess) e = cudaEventSynchronize(e1);
  float ms = 0.f;
  if (e == cudaSuccess) cudaEventElapsedTime(&ms, e0, e1);
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  if (e == cudaSuccess)
    e = cudaMemcpy2D(K_host, ldk * sizeof(double), K_dev, X->N * sizeof(double), X->N * sizeof(double), Y->N,
                     cudaMemcpyDeviceToHost);
  pool_free(K_dev);
  if (e != cudaSuccess) return cuda_fail(e, "cross-check Gram kernel");
  if (ms_out) *ms_out = ms;
  return QK_OK;
}

// ---------------------------------------------------------------- low-chi stage 2 on the stores
int qk_batch_store(const qk_batch* b, void** store_dev, int64_t* state_stride, void** chi_dev) {
  if (!b) return fail(QK_ERR_ARG, "NULL batch");
  if (store_dev) *store_dev = b->store;
  if (state_stride) *state_stride = b->state_stride;
  if (chi_dev) *chi_dev = b->chi;
  return QK_OK;
}

int qk_gram_lane(const qk_plan* plan, int device, void* stream_v, int max_chi, const void* storeX, const int32_t* chiX,
                 int Nx, const void* storeY, const int32_t* chiY, int Ny, const int32_t* tiles, int n_tiles,
                 int symmetric, double* K_dev, int64_t ldk, float* ms_out) {
  if (!plan || !storeX || !chiX || Nx < 1 || !K_dev || n_tiles < 0 || (n_tiles > 0 && !tiles))
    return fail(QK_ERR_ARG, "bad arguments");
  if (symmetric) { storeY = storeX; chiY = chiX; Ny = Nx; }
  if (!storeY || !chiY || Ny < 1) return fail(QK_ERR_ARG, "bad Y arguments");
  if (ldk < Nx) return fail(QK_ERR_ARG, "ldk must be >= Nx");
  if (max_chi < 1 || max_chi > 4) return fail(QK_ERR_LIMIT, "the lane-per-pair overlap kernel supports bond dimensions <= 4");
  QK_CUDA(cudaSetDevice(device), "cudaSetDevice");
  cudaStream_t stream = (cudaStream_t)stream_v;
  int TX = 0, TY = 0;
  qk_gram_lane_tile_shape(&TX, &TY);
  std::vector<int4> cta;
  for (int t = 0; t < n_tiles; ++t) {
    const int r0 = tiles[4 * t], r1 = tiles[4 * t + 1], c0 = tiles[4 * t + 2], c1 = tiles[4 * t + 3];
    if (r0 < 0 || c0 < 0 || r1 > Ny || c1 > Nx || r0 > r1 || c0 > c1) return fail(QK_ERR_ARG, "tile out of range");
    for (int y0 = r0; y0 < r1; y0 += TY)
      for (int x0 = c0; x0 < c1; x0 += TX) {
        const int ye = std::min(y0 + TY, r1), xe = std::min(x0 + TX, c1);
        if (symmetric && x0 > ye - 1) continue;
        cta.push_back(make_int4(y0, x0, ye, xe));
      }
  }
  if (cta.empty()) { if (ms_out) *ms_out = 0.f; return QK_OK; }
  const size_t nb = (size_t)(plan->n + 1);
  const size_t o_off = (nb * sizeof(int32_t) + 15) & ~(size_t)15, t_off = (o_off + nb * sizeof(int64_t) + 15) & ~(size_t)15;
  std::vector<unsigned char> hbuf(t_off + cta.size() * sizeof(int4));
  memcpy(hbuf.data(), plan->cap.data(), nb * sizeof(int32_t));
  memcpy(hbuf.data() + o_off, plan->site_off.data(), nb * sizeof(int64_t));
  memcpy(hbuf.data() + t_off, cta.data(), cta.size() * sizeof(int4));
  unsigned char* dbuf = nullptr;
  QK_CUDA(pool_alloc_t(&dbuf, hbuf.size()), "cudaMalloc(gram scratch)");
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  cudaError_t e = cudaMemcpyAsync(dbuf, hbuf.data(), hbuf.size(), cudaMemcpyHostToDevice, stream);
  LaneParams P;
  P.n = plan->n;
  P.cap = (const int32_t*)dbuf; P.site_off = (const int64_t*)(dbuf + o_off); P.state_stride = plan->state_stride;
  P.storeX = (const c128*)storeX; P.chiX = chiX; P.storeY = (const c128*)storeY; P.chiY = chiY;
  P.Nx = Nx; P.Ny = Ny;
  P.tiles = (const int4*)(dbuf + t_off); P.n_cta_tiles = (int)cta.size();
  P.symmetric = symmetric ? 1 : 0; P.K = K_dev; P.ldk = ldk;
  if (e == cudaSuccess) e = cudaEventCreate(&e0);
  if (e == cudaSuccess) e = cudaEventCreate(&e1);
  if (e == cudaSuccess) e = cudaEventRecord(e0, stream);
  if (e == cudaSuccess) e = qk_launch_gram_lane(P, max_chi, stream);
  if (e == cudaSuccess) e = cudaEventRecord(e1, stream);
  if (e == cudaSuccess) e = cudaEventSynchronize(e1);
  float ms = 0.f;
  if (e == cudaSuccess) cudaEventElapsedTime(&ms, e0, e1);
  if (e0) cudaEventDestroy(e0);
  if (e1) cudaEventDestroy(e1);
  pool_free(dbuf);
  if (e != cudaSuccess) return cuda_fail(e, "lane-per-pair Gram kernel");
  if (ms_out) *ms_out = ms;
  return QK_OK;
}


// ---------------------------------------------------------------- stage 2, any bond dimension (batched DMMA GEMMs)
static int gram_big_run(int device, cudaStream_t stream, int n, const int64_t* soff_x, const int64_t* soff_y,
                        const int32_t* dims_x, const int32_t* dims_y, const c128* storeX, int64_t strideX,
                        const int32_t* chiX, int Nx, const c128* storeY, int64_t strideY, const int32_t* chiY, int Ny,
                        const int32_t* tiles, int n_tiles, int symmetric, double* K_dev, int64_t ldk, float* ms_out) {
  QK_CUDA(cudaSetDevice(device), "cudaSetDevice");
  std::vector<int2> pairs;
  for (int t = 0; t < n_tiles; ++t) {
    const int r0 = tiles[4 * t], r1 = tiles[4 * t + 1], c0 = tiles[4 * t + 2], c1 = tiles[4 * t + 3];
    if (r0 < 0 || c0 < 0 || r1 > Ny || c1 > Nx || r0 > r1 || c0 > c1) return fail(QK_ERR_ARG, "tile out of range");
    for (int y = r0; y < r1; ++y)
      for (int x = c0; x < c1; ++x)
        if (!symmetric || x <= y) pairs.push_back(make_int2(y, x));
  }
  if (pairs.empty()) { if (ms_out) *ms_out = 0.f; return QK_OK; }
  int64_t emax = 1, tmax = 2;
  for (int b = 0; b < n; ++b) {
    emax = std::max<int64_t>(emax, (int64_t)dims_y[b] * dims_x[b]);
    emax = std::max<int64_t>(emax, (int64_t)dims_y[b + 1] * dims_x[b + 1]);
    tmax = std::max<int64_t>(tmax, (int64_t)dims_y[b] * 2 * dims_x[b + 1]);
  }
  size_t budget = (size_t)6 << 30;   // bytes of E + T scratch per chunk of pairs
  if (const char* ev = getenv("QK_GRAM_BIG_SCRATCH_MB")) { const long v = atol(ev); if (v > 0) budget = (size_t)v << 20; }
  size_t chunk = budget / ((size_t)(emax + tmax) * sizeof(c128));
  chunk = std::max<size_t>(1, std::min(chunk, pairs.size()));
  chunk = std::min<size_t>(chunk, (size_t)1 << 22);
  c128 *E = nullptr, *T = nullptr; int2* pairs_dev = nullptr;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  cudaError_t e = pool_alloc_t(&E, chunk * emax * sizeof(c128));
  if (e == cudaSuccess) e = pool_alloc_t(&T, chunk * tmax * sizeof(c128));
  if (e == cudaSuccess) e = pool_alloc_t(&pairs_dev, pairs.size() * sizeof(int2));
  if (e == cudaSuccess) e = cudaMemcpyAsync(pairs_dev, pairs.data(), pairs.size() * sizeof(int2), cudaMemcpyHostToDevice, stream);
  if (e == cudaSuccess) e = cudaEventCreate(&e0);
  if (e == cudaSuccess) e = cudaEventCreate(&e1);
  if (e == cudaSuccess) e = cudaEventRecord(e0, stream);
  for (size_t p0 = 0; p0 < pairs.size() && e == cudaSuccess; p0 += chunk) {
    const int cnt = (int)std::min(chunk, pairs.size() - p0);
    e = qk_launch_gram_big(n, soff_x, soff_y, dims_x, dims_y, storeX, strideX, chiX, storeY, strideY, chiY,
                           pairs_dev + p0, cnt, symmetric, E, emax, T, tmax, K_dev, ldk, stream);
  }
  if (e == cudaSuccess) e = cudaEventRecord(e1, stream);
  if (e == cudaSuccess) e = cudaEventSynchronize(e1);
  float ms = 0.f;
  if (e == cudaSuccess) cudaEventElapsedTime(&ms, e0, e1);
  if (e0) cudaEventDestroy(e0);
  if (e1) cudaEventDestroy(e1);
  pool_free(E); pool_free(T); pool_free(pairs_dev);
  if (e != cudaSuccess) return cuda_fail(e, "batched-GEMM Gram kernels");
  if (ms_out) *ms_out = ms;
  return QK_OK;
}

int qk_gram_big(const qk_plan* plan, int device, void* stream_v, const int32_t* dims_x, const int32_t* dims_y,
                const void* storeX, const int32_t* chiX, int Nx, const void* storeY, const int32_t* chiY, int Ny,
                const int32_t* tiles, int n_tiles, int symmetric, double* K_dev, int64_t ldk, float* ms_out) {
  if (!plan || !storeX || !chiX || Nx < 1 || !K_dev || n_tiles < 0 || (n_tiles > 0 && !tiles))
    return fail(QK_ERR_ARG, "bad arguments");
  if (symmetric) { storeY = storeX; chiY = chiX; Ny = Nx; dims_y = dims_x; }
  if (!storeY || !chiY || Ny < 1) return fail(QK_ERR_ARG, "bad Y arguments");
  if (ldk < Nx) return fail(QK_ERR_ARG, "ldk must be >= Nx");
  if (!dims_x) dims_x = plan->cap.data();
  if (!dims_y) dims_y = plan->cap.data();
  for (int b = 0; b <= plan->n; ++b)
    if (dims_x[b] < 1 || dims_y[b] < 1 || dims_x[b] > plan->cap[b] || dims_y[b] > plan->cap[b])
      return fail(QK_ERR_ARG, "per-bond maxima must lie in [1, bond cap of the plan]");
  return gram_big_run(device, (cudaStream_t)stream_v, plan->n, plan->site_off.data(), plan->site_off.data(), dims_x, dims_y,
                      (const c128*)storeX, plan->state_stride, chiX, Nx, (const c128*)storeY, plan->state_stride, chiY, Ny,
                      tiles, n_tiles, symmetric, K_dev, ldk, ms_out);
}

int qk_batch_repack(const qk_batch* b, const qk_plan* plan, void* store_dev, int32_t* chi_dev, const int32_t* dst_index,
                    void* stream_v) {
  if (!b || !plan || !store_dev || !chi_dev) return fail(QK_ERR_ARG, "NULL argument");
  if (b->n != plan->n) return fail(QK_ERR_ARG, "batch and plan differ in the number of qubits");
  if (b->N == 0) return QK_OK;
  if (!b->store) return fail(QK_ERR_ARG, "the batch's store was released");
  QK_CUDA(cudaSetDevice(b->device), "cudaSetDevice");
  cudaStream_t stream = (cudaStream_t)stream_v;
  int64_t* off_dev = nullptr; int32_t* dst_dev = nullptr;
  QK_CUDA(pool_alloc_t(&off_dev, (b->n + 1) * sizeof(int64_t)), "cudaMalloc(offsets)");
  cudaError_t e = cudaMemcpyAsync(off_dev, plan->site_off.data(), (b->n + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, stream);
  if (e == cudaSuccess && dst_index) e = pool_alloc_t(&dst_dev, (size_t)b->N * sizeof(int32_t));
  if (e == cudaSuccess && dst_index)
    e = cudaMemcpyAsync(dst_dev, dst_index, (size_t)b->N * sizeof(int32_t), cudaMemcpyHostToDevice, stream);
  if (e == cudaSuccess)
    e = qk_launch_repack(b->n, b->N, b->store, b->state_stride, b->site_off_dev, b->chi, (c128*)store_dev,
                         plan->state_stride, off_dev, chi_dev, dst_dev, stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
  pool_free(off_dev); pool_free(dst_dev);
  if (e != cudaSuccess) return cuda_fail(e, "repack kernel");
  return QK_OK;
}

// ---------------------------------------------------------------- whole path, host buffers
// OR of the stage-1 flags of every state of a batch
static int batch_flags_or(const qk_batch* b, int* flags_or) {
  *flags_or = 0;
  if (!b || b->N == 0) return QK_OK;
  std::vector<QkStat> st(b->N);
  QK_CUDA(cudaSetDevice(b->device), "cudaSetDevice");
  QK_CUDA(cudaMemcpy(st.data(), b->stats, st.size() * sizeof(QkStat), cudaMemcpyDeviceToHost), "copy stats");
  for (const QkStat& q : st) *flags_or |= q.flags;
  return QK_OK;
}

int qk_gram_host(const qk_plan* plan, int device, const double* X_host, int Nx, const double* Y_host, int Ny, int ldx,
                 double* K_host, int64_t ldk) {
  if (!plan || !X_host || !K_host || Nx < 1) return fail(QK_ERR_ARG, "bad arguments");
  const bool sym = (Y_host == nullptr);
  if (!sym && Ny < 1) return fail(QK_ERR_ARG, "bad Ny");
  if (ldk < Nx) return fail(QK_ERR_ARG, "ldk must be >= Nx");
  qk_batch *bx = nullptr, *by = nullptr;
  int rc = qk_simulate(plan, device, X_host, Nx, ldx, &bx);
  if (rc == QK_OK && !sym) rc = qk_simulate(plan, device, Y_host, Ny, ldx, &by);
  // The plan's bond cap is fixed here (no escalation, unlike the Python engine): a state that wanted more than
  // the cap was hard-truncated, so the Gram matrix would be wrong -- refuse instead of returning it.
  for (qk_batch* b : {bx, by}) {
    int fl = 0;
    if (rc == QK_OK) rc = batch_flags_or(b, &fl);
    if (rc == QK_OK && (fl & QK_FLAG_CAP_HIT))
      rc = fail(QK_ERR_LIMIT, "a state's bond dimension exceeds the plan's chi_cap: re-create the plan with a larger cap");
    if (rc == QK_OK && (fl & QK_FLAG_NO_CONVERGE))
      rc = fail(QK_ERR_CUDA, "Jacobi SVD hit its sweep limit (stage 1 did not converge)");
  }
  void *fx = nullptr, *fy = nullptr;
  double* K_dev = nullptr;
  const int n = plan->n;
  const int rows = sym ? Nx : Ny;
  std::vector<int32_t> Dx(n + 1), Dy(n + 1);
  auto padded = [&](qk_batch* b, std::vector<int32_t>& D) {
    int r = qk_batch_max_chi(b, D.data());
    for (int s = 0; s <= n; ++s) D[s] = (D[s] + 7) & ~7;
    return r;
  };
  if (rc == QK_OK) rc = padded(bx, Dx);
  if (rc == QK_OK && !sym) rc = padded(by, Dy);
  int maxD = 8;
  if (rc == QK_OK) {
    for (int s = 0; s <= n; ++s) maxD = std::max(maxD, std::max(Dx[s], sym ? 8 : Dy[s]));
  }
  if (rc == QK_OK && maxD > 16) {
    // bond dimensions above the tensor-core kernel's register budget: CUDA-core kernel
    rc = qk_gram_store(bx, sym ? nullptr : by, K_host, ldk, nullptr);
  } else if (rc == QK_OK) {
    int64_t sx = 0, sy = 0;
    qk_frag_stride(n, Dx.data(), &sx);
    if (!sym) qk_frag_stride(n, Dy.data(), &sy);
    cudaError_t e = pool_alloc(&fx, (size_t)sx * Nx);
    if (e == cudaSuccess && !sym) e = pool_alloc(&fy, (size_t)sy * Ny);
    if (e == cudaSuccess) e = pool_alloc_t(&K_dev, (size_t)rows * Nx * sizeof(double));
    if (e != cudaSuccess) rc = cuda_fail(e, "cudaMalloc(frag/K)");
    if (rc == QK_OK) rc = qk_batch_pack(bx, Dx.data(), fx, nullptr);
    if (rc == QK_OK && !sym) rc = qk_batch_pack(by, Dy.data(), fy, nullptr);
    const int32_t tile[4] = {0, rows, 0, Nx};
    if (rc == QK_OK)
      rc = qk_gram_frags(device, nullptr, n, Dx.data(), fx, Nx, sym ? nullptr : Dy.data(), fy, Ny, tile, 1, sym ? 1 : 0,
                         K_dev, Nx, nullptr);
    if (rc == QK_OK) {
      e = cudaMemcpy2D(K_host, ldk * sizeof(double), K_dev, Nx * sizeof(double), Nx * sizeof(double), rows,
                       cudaMemcpyDeviceToHost);
      if (e != cudaSuccess) rc = cuda_fail(e, "cudaMemcpy2D(K)");
    }
  }
  std::string keep = g_err;
  pool_free(fx); pool_free(fy); pool_free(K_dev);
  qk_batch_destroy(bx); qk_batch_destroy(by);
  g_err = keep;
  return rc;
}

int qk_dmma_peak(int device, int iters, double* tflops) {
  if (!tflops || iters < 1) return fail(QK_ERR_ARG, "bad arguments");
  QK_CUDA(cudaSetDevice(device), "cudaSetDevice");
  QK_CUDA(qk_run_dmma_peak(iters, tflops), "dmma peak kernel");
  return QK_OK;
}

int qk_pipe_mix(int device, int iters, float* ms) {
  if (!ms || iters < 1) return fail(QK_ERR_ARG, "bad arguments");
  QK_CUDA(cudaSetDevice(device), "cudaSetDevice");
  QK_CUDA(qk_run_pipe_mix(iters, ms), "pipe mix kernel");
  return QK_OK;
}

}  // extern "C"
