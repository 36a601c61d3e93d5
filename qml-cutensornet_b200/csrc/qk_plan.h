// Static schedule compiler: gate list of one ansatz -> op program shared by every datapoint.
#pragma once
#include <string>
#include <vector>
#include "qk_types.h"
#include "../../include/qkmps.h"

struct qk_plan {
  int n = 0;
  int n_gates = 0;
  int trunc_mode = 0;
  double trunc_error = 0.0;
  int chi_cap = 0;
  int threads = 0;                 // cooperative group size G of the stage-1 kernel
  int rmax = 0;                    // 2 * chi_cap
  size_t smem_bytes = 0;
  std::vector<QkOp> ops;
  std::vector<int32_t> cap;        // [n+1]
  std::vector<int64_t> site_off;   // [n+1]
  int64_t state_stride = 0;
  int n_2q = 0, n_1q = 0, n_moves = 0;
  int reorder = 1;                 // commutation-aware reordering of interaction runs (qk_plan.cpp)
  int early_exit = 0;              // QK_PLAN_EARLY_EXIT
  int fuse = 1;                    // fuse consecutive 2-qubit gates on one bond (off: QK_PLAN_NO_FUSION)
  int parallel = 0;                // QK_PLAN_PARALLEL: B form, levelised ops
  std::vector<int32_t> level_start;  // [n_levels + 1] (parallel plans)
  int big = 0;                     // large-matrix path (qk_sim_big.h): chi_cap above QK_CHI_LIMIT, or QK_PLAN_BIG
  int jb = 8;                      // columns per block of its block Jacobi
};

// Returns 0 or a negative qk_status; err receives a message.
int qk_compile_plan(int n_qubits, const qk_gate* gates, int n_gates, int trunc_mode, double trunc_error,
                    int chi_cap, int flags, qk_plan* plan, std::string* err);
// H / Rz / routed XXPhase gate list of the ansatz (gpu_backend/kernel_state_ansatz.py:53-90)
int qk_ansatz_gates(int n_qubits, int reps, double gamma, int hadamard_init, const int32_t* pairs, int n_pairs,
                    std::vector<qk_gate>* out, std::string* err);
int qk_pick_threads(int chi_cap);
#define QK_CHI_LIMIT 32        // shared-memory-resident stage-1 kernels (qk_sim_core.h)
#define QK_CHI_LIMIT_BIG 512   // large-matrix stage-1 kernel (qk_sim_big.h): (2 chi) x 2 jb columns must fit in shared memory
