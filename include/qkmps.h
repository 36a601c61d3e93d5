/* qkmps.h -- C ABI of libqkmps.so, the B200-native quantum-kernel MPS engine.
 *
 * The reference (mmetcalf14/qml-cutensornet) has no native ABI: its hot path
 * is Python that calls third-party libraries.  Each entry point below states
 * which reference interface it replaces (paths relative to the reference
 * repo).  All functions return 0 on success and a negative qk_status on
 * failure; qk_last_error() returns a thread-local message.  Host pointers are
 * borrowed for the duration of the call.  Opaque handles are owned by the
 * library and released by the matching *_destroy.  There is no CPU fallback:
 * every compute entry point fails with QK_ERR_CUDA when no sm_100 device is
 * usable.
 */
#ifndef QKMPS_H
#define QKMPS_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QKMPS_VERSION 100

typedef enum {
  QK_OK = 0,
  QK_ERR_ARG = -1,       /* bad argument (also: unknown gate -- cpu_backend/kernel_state_ansatz.py:129, KernelPkg.jl:62) */
  QK_ERR_CUDA = -2,      /* CUDA runtime / launch failure, or no device */
  QK_ERR_LIMIT = -3,     /* bond dimension above what the kernels support (chi_cap <= 512), or above a plan's cap */
  QK_ERR_ALLOC = -4
} qk_status;

/* gate kinds accepted by the plan compiler: the gate set of KernelPkg/src/KernelPkg.jl:45-64 */
typedef enum {
  QK_GATE_H = 0,
  QK_GATE_RZ = 1,        /* alpha = coeff * x[fa]            (fa < 0: alpha = coeff) */
  QK_GATE_RX = 2,
  QK_GATE_XX = 3,        /* alpha = coeff * (1-x[fa])*(1-x[fb])  (fa < 0: alpha = coeff) */
  QK_GATE_ZZ = 4,
  QK_GATE_SWAP = 5
} qk_gate_kind;

/* One symbolic gate; alpha is in half-turns (TKET convention, theta = pi*alpha/2). */
typedef struct {
  int32_t kind;          /* qk_gate_kind */
  int32_t q0, q1;        /* q1 ignored for 1-qubit gates; 2-qubit gates need q1 == q0 + 1 */
  int32_t fa, fb;        /* feature indices of the angle expression */
  double coeff;
} qk_gate;

/* truncation rule */
typedef enum {
  QK_TRUNC_ITENSORS = 0, /* relative discarded weight <= cutoff, no renormalisation
                            (cpu_backend/kernel_state_ansatz.py:262 -> KernelPkg.jl:68 apply(..; cutoff)) */
  QK_TRUNC_PYTKET = 1    /* kept weight fraction >= 1 - truncation_error, sigma < 1e-16 dropped,
                            renormalise, track fidelity (gpu_backend/kernel_state_ansatz.py:141-144) */
} qk_trunc_mode;

/* plan flags */
#define QK_PLAN_DEFAULT 0
#define QK_PLAN_LITERAL_ORDER 1   /* keep the gate list's order of the (mutually commuting) XXPhase / ZZPhase
                                     interactions; default: each run of them is applied as one sweep so the
                                     orthogonality centre never has to jump (same unitary, fewer gauge moves) */

#define QK_PLAN_EARLY_EXIT 2      /* stop simulating a datapoint at its first bond-cap hit: its state is then
                                     invalid and QK_FLAG_CAP_HIT (bit 0 of qk_batch_info flags) is set; lets a
                                     caller try a small cap first and re-run only the states that need more */

#define QK_PLAN_NO_FUSION 4       /* keep one SVD per 2-qubit gate (default: gates that follow each other on the
                                     same bond are multiplied into one SVD and SWAP pairs cancel) */
#define QK_PLAN_PARALLEL 8        /* stage 1 in B (Vidal/Hastings) form: no gauge moves; ops levelised by the sites
                                     they touch so that independent bonds of one datapoint are updated concurrently */

#define QK_PLAN_BIG 16            /* force the large-matrix stage-1 kernel (theta in global memory, block Jacobi, one
                                     CTA cluster per datapoint) that plans with chi_cap > 32 use anyway; tests */

typedef struct qk_plan qk_plan;     /* compiled static op schedule of one ansatz (host object) */
typedef struct qk_batch qk_batch;   /* device-resident batch of simulated MPS */

typedef struct {
  int32_t n_qubits, n_gates, n_ops, n_ops_2q, n_ops_1q, n_moves;
  int32_t chi_cap, threads, trunc_mode;
  int32_t smem_bytes;
  int64_t state_stride;             /* c128 elements reserved per state in the working store */
  double trunc_error;
} qk_plan_info_t;

typedef struct {                    /* one op of the compiled schedule (for tests / inspection) */
  int32_t kind;                     /* qk_gate_kind, or 16 = move-right, 17 = move-left */
  int32_t site, fa, fb, dir;
  double coeff;
} qk_op_view;

int qk_version(void);
const char* qk_last_error(void);
int qk_device_count(int* count);

/* ---- plan: replaces KernelStateAnsatz's circuit + routing and the per-gate bookkeeping of the
 *      third-party simulators (gpu_backend/kernel_state_ansatz.py:53-90; main.py:21-45,73). ---- */
int qk_plan_create_gates(int n_qubits, const qk_gate* gates, int n_gates,
                         int trunc_mode, double trunc_error, int chi_cap, int flags, qk_plan** out);
/* builds H / Rz / routed XXPhase from the ansatz parameters, then compiles it */
int qk_plan_create_ansatz(int n_qubits, int reps, double gamma, int hadamard_init,
                          const int32_t* pairs /*[n_pairs][2]*/, int n_pairs,
                          int trunc_mode, double trunc_error, int chi_cap, int flags, qk_plan** out);
int qk_plan_info(const qk_plan* plan, qk_plan_info_t* info);
int qk_plan_ops(const qk_plan* plan, qk_op_view* ops, int max_ops);   /* returns number written */
void qk_plan_destroy(qk_plan* plan);

/* ---- stage 1: replaces simulate(libhandle, circ, MPSxGate, config) per datapoint
 *      (gpu_backend/kernel_state_ansatz.py:213-231,255-273) and build_and_sim_circ
 *      (KernelPkg/src/KernelPkg.jl:45-72).  X is row-major [N][ldx] float64. ---- */
int qk_simulate(const qk_plan* plan, int device, const double* X_host, int N, int ldx, qk_batch** out);
int qk_simulate_dev(const qk_plan* plan, int device, void* stream, const double* X_dev, int N, int ldx,
                    qk_batch** out);
/* Asynchronous variant: returns as soon as the kernel is queued on `stream` (no host synchronisation; scratch is
 * released in stream order).  The batch may be passed to qk_batch_pack_async / qk_gram_* on the same stream at
 * once; qk_batch_info / qk_batch_sim_ms / qk_batch_unit_seconds wait for the kernel.  Lets a caller queue
 * stage 1 -> pack -> exchange -> stage 2 without a host round trip (the reference interleaves its ring exchange
 * with the products, gpu_backend/kernel_state_ansatz.py:330-334,366-419). */
int qk_simulate_async(const qk_plan* plan, int device, void* stream, const double* X_dev, int N, int ldx,
                      qk_batch** out);
/* One datapoint with a memory trace: bytes_per_op[o] = sum of site-tensor bytes after op o of the compiled
 * schedule (qk_plan_ops gives the ops).  Replaces the "MPS size (MiB)=" debug log of pytket-cutensornet that
 * main_track_mem.py:168-172,254-256 captures and runs/mem_evol/plot.py:12-15 parses. */
int qk_simulate_trace(const qk_plan* plan, int device, const double* x_host, int ldx, double* bytes_per_op,
                      int max_ops, qk_batch** out);
/* last stage-1 kernel time in ms (CUDA events on the launching stream) */
int qk_batch_sim_ms(const qk_batch* batch, float* ms);
/* OR of the QK_FLAG_* bits (bit 0: bond cap hit, bit 1: Jacobi sweep limit) of every state of the batch; waits
 * for the stage-1 kernel */
int qk_batch_flags(const qk_batch* batch, int32_t* flags_or);
/* seconds each datapoint's circuit took inside the kernel (clock64 around the datapoint; N doubles): the
 * per-circuit times the reference records around every simulate() call (gpu_backend/kernel_state_ansatz.py:220-222)
 * and reports as median / quartiles (:299-316) */
int qk_batch_unit_seconds(const qk_batch* batch, double* seconds /*[N]*/);

/* ---- MPS handle surface the orchestrator needs (gpu_backend/kernel_state_ansatz.py:223,295-296):
 *      bond dimensions, byte size, accumulated fidelity.  Any output pointer may be NULL. ---- */
int qk_batch_size(const qk_batch* batch, int* N, int* n_qubits);
int qk_batch_info(const qk_batch* batch, int32_t* chi /*[N][n+1]*/, double* fidelity /*[N]*/,
                  double* trunc_weight /*[N]*/, int64_t* nbytes /*[N]*/, int32_t* flags /*[N]*/,
                  int32_t* sweeps /*[N]*/);
/* site tensors of state i, concatenated, each [chi_l][2][chi_r] complex128 row-major */
int qk_batch_export(const qk_batch* batch, int i, void* host_buf, int64_t buf_bytes);
/* test hook: upload MPS made elsewhere (same layout as qk_batch_export, states concatenated) */
int qk_batch_import(int device, int n_qubits, int N, const int32_t* chi /*[N][n+1]*/,
                    const void* host_tensors, int64_t bytes, qk_batch** out);
int qk_batch_max_chi(const qk_batch* batch, int32_t* max_chi /*[n+1]*/);
/* Frees the batch's working store (its site tensors) in stream order once they have been packed / exchanged; bond
 * dimensions, fidelities, flags and per-datapoint times stay readable, export / pack / repack then fail with QK_ERR_ARG */
int qk_batch_release_store(qk_batch* batch, void* stream);
void qk_batch_destroy(qk_batch* batch);

/* ---- exchange format: replaces pickled MPS send/recv/sendrecv
 *      (gpu_backend/kernel_state_ansatz.py:346-352,416-419).  A "frag" buffer holds every state
 *      of a batch in DMMA-fragment order with batch-uniform padded bond dimensions D[n+1]
 *      (multiples of 8, D[0] = D[n] = 8); it is plain device memory the caller may all-gather. ---- */
int qk_frag_stride(int n_qubits, const int32_t* D, int64_t* bytes_per_state);
int qk_batch_pack(const qk_batch* batch, const int32_t* D, void* frag_dev, void* stream);
/* asynchronous: state i -> position first_index + i; no host synchronisation, no bond-dimension check (D must
 * cover the bond caps of the batch's plan) */
int qk_batch_pack_async(const qk_batch* batch, const int32_t* D, void* frag_dev, int first_index, void* stream);
/* same, state i written at position dst_index[i] of the frag buffer (skipped if negative): merges the
 * valid states of several batches (different bond caps) into one exchange buffer */
int qk_batch_pack_scatter(const qk_batch* batch, const int32_t* D, void* frag_dev, const int32_t* dst_index,
                          void* stream);

/* ---- stage 2: replaces x_mps.vdot(y_mps) + |.|^2 (gpu_backend/kernel_state_ansatz.py:380-387) and
 *      abs(inner(y, x))^2 (KernelPkg/src/KernelPkg.jl:103-109).  K[y][x] = |<psi_y|psi_x>|^2 for the
 *      listed tiles (row = y in [r0,r1), col = x in [c0,c1)); K_dev is device float64 with leading
 *      dimension ldk.  symmetric != 0: fragY is ignored (Y = X), only x <= y is computed inside each
 *      tile and mirrored (cpu_backend/kernel_state_ansatz.py:271-274). ---- */
int qk_gram_frags(int device, void* stream, int n_qubits,
                  const int32_t* Dx, const void* fragX, int Nx,
                  const int32_t* Dy, const void* fragY, int Ny,
                  const int32_t* tiles /*[n_tiles][4] = r0,r1,c0,c1*/, int n_tiles, int symmetric,
                  double* K_dev, int64_t ldk, float* ms_out);
/* qk_gram_frags with ms_out == NULL is asynchronous (kernel queued on `stream`, no host synchronisation).
 * Per-tile timing: the next qk_gram_frags call of this thread writes the clock64 ticks every CTA tile (8 inner
 * products) took to clk_dev (device, `capacity` entries; NULL switches it off) -- the per-product times the
 * reference records around every vdot (gpu_backend/kernel_state_ansatz.py:379-381). */
int qk_gram_set_tile_clocks(long long* clk_dev, int64_t capacity);
int64_t qk_gram_tile_clocks_used(void);
/* Same Gram tiles for batches whose bond dimensions are all <= 4 (max_chi; the regime of the reference's published
 * scaling runs, runs/runtime_scaling: chi = 2): one lane per (y, x) pair on the FP64 CUDA cores, reading the
 * unpadded stage-1 stores [N][state_stride] c128 + chi [N][n+1] of states simulated with `plan` (raw device
 * pointers so that shards gathered from other ranks can be passed).  symmetric != 0: storeY / chiY ignored. */
int qk_batch_store(const qk_batch* batch, void** store_dev, int64_t* state_stride_c128, void** chi_dev);
int qk_gram_lane(const qk_plan* plan, int device, void* stream, int max_chi,
                 const void* storeX, const int32_t* chiX, int Nx,
                 const void* storeY, const int32_t* chiY, int Ny,
                 const int32_t* tiles /*[n_tiles][4] = r0,r1,c0,c1*/, int n_tiles, int symmetric,
                 double* K_dev, int64_t ldk, float* ms_out);
/* Same Gram tiles for ANY bond dimension (the path of BASELINE config 4, chi ~ 100): the transfer sweep of every
 * (y, x) pair as two batched complex GEMMs per site on the FP64 tensor cores (DMMA), E and T of the pairs in
 * global memory, site tensors read from the unpadded stores laid out by `plan`.  dims_x / dims_y: per-bond maxima
 * of the bond dimensions over the X / Y states, [n+1] host ints (NULL: the plan's bond caps) -- they size the
 * scratch buffers and the launch grids.  Replaces the same reference calls as qk_gram_frags. */
int qk_gram_big(const qk_plan* plan, int device, void* stream, const int32_t* dims_x, const int32_t* dims_y,
                const void* storeX, const int32_t* chiX, int Nx, const void* storeY, const int32_t* chiY, int Ny,
                const int32_t* tiles /*[n_tiles][4] = r0,r1,c0,c1*/, int n_tiles, int symmetric,
                double* K_dev, int64_t ldk, float* ms_out);
/* Copies the states of `batch` into a store laid out for `plan` (its bond caps define the site slots) with the chi
 * rows alongside: state i goes to position dst_index[i] (negative: skipped; NULL: i).  Merges batches simulated
 * with different bond caps into the one buffer that is exchanged between ranks (replaces pickling MPS objects of
 * different sizes, gpu_backend/kernel_state_ansatz.py:346-352,416-419).  Every copied state must fit the caps. */
int qk_batch_repack(const qk_batch* batch, const qk_plan* plan, void* store_dev, int32_t* chi_dev,
                    const int32_t* dst_index, void* stream);
/* FP64 cross-check on the unpadded stores: CUDA cores for chi <= 32, the batched-GEMM sweep above */
int qk_gram_store(const qk_batch* X, const qk_batch* Y_or_null, double* K_host, int64_t ldk, float* ms_out);

/* ---- whole path with host buffers on one device: what build_kernel_matrix does per process.
 *      K_host is [Ny or Nx][ldk].  The plan's chi_cap is used as is (no escalation): if any state wants a larger
 *      bond dimension the call fails with QK_ERR_LIMIT instead of returning a hard-truncated kernel matrix
 *      (the Python engine re-runs such states with the next cap of its ladder). ---- */
int qk_gram_host(const qk_plan* plan, int device, const double* X_host, int Nx,
                 const double* Y_host_or_null, int Ny, int ldx, double* K_host, int64_t ldk);

/* FP64 tensor-core (DMMA m8n8k4) peak microbenchmark: TFLOP/s over `iters` MMAs per warp */
int qk_dmma_peak(int device, int iters, double* tflops);
/* Do the FP64 tensor pipe (DMMA) and the FP64 FMA pipe (DFMA) run concurrently on one SM?  Four warps of every
 * CTA issue DMMAs, four issue DFMAs; ms[0] = DMMA warps alone, ms[1] = DFMA warps alone, ms[2] = both.
 * ms[2] ~ max(ms[0], ms[1]): independent pipes; ms[2] ~ ms[0] + ms[1]: one shared datapath. */
int qk_pipe_mix(int device, int iters, float* ms /*[3]*/);

#ifdef __cplusplus
}
#endif
#endif /* QKMPS_H */
