// TEST INFRASTRUCTURE ONLY: host emulation of the stage-1 device core.
//
// Compiles qk_sim_core.h with QK_HOST_EMU so that every "parallel phase" becomes a loop over the
// thread ids of one cooperative group.  It lets the CPU-only test-suite (-m "not gpu") check the
// *same source* the CUDA kernel is built from against the oracle: schedule compiler, contraction,
// Jacobi SVD, truncation rules, Householder gauge moves.  Nothing in the product package loads
// this library; the product path has no CPU fallback.
#define QK_HOST_EMU 1
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>
#include "../../qml-cutensornet_b200/csrc/qk_plan.h"
#include "../../qml-cutensornet_b200/csrc/qk_sim_big.h"

template <int G>
static void run_group(const SimParams& P, unsigned char* smem, int dp) {
  SimCtx c;
  if (P.big_w) {   // large-matrix path: one CTA stands for the whole cluster
    qk_big_carve(c, &P, smem, G, 0, 0, 1);
    qk_big_datapoint<G>(c, dp);
    return;
  }
  qk_sim_carve(c, &P, smem, G);
  if (P.parallel) {
    QkStat part;
    qk_sim_datapoint_b<G>(c, dp, 0, 1, &part);
  } else {
    qk_sim_datapoint<G>(c, dp);
  }
}

extern "C" {

// Simulate N datapoints.  Outputs: chi [N][n+1], store [N][state_stride] c128 (slots laid out by
// site_off), stats as 4 doubles per datapoint (fidelity, trunc_weight, flags, sweeps).
// Returns state_stride (>0) or a negative status.  If store == NULL only returns the stride.
long long qk_emu_simulate(int n, const qk_gate* gates, int n_gates, int trunc_mode, double trunc_error,
                          int chi_cap, int flags, int force_threads, const double* X, int N, int ldx,
                          int32_t* chi_out, double* store_out, long long* site_off_out, double* stats_out,
                          int* n_ops_out, int* n_moves_out) {
  qk_plan plan;
  std::string err;
  int rc = qk_compile_plan(n, gates, n_gates, trunc_mode, trunc_error, chi_cap, flags, &plan, &err);
  if (rc != 0) return rc;
  if (n_ops_out) *n_ops_out = (int)plan.ops.size();
  if (n_moves_out) *n_moves_out = plan.n_moves;
  if (site_off_out) for (int s = 0; s <= n; ++s) site_off_out[s] = plan.site_off[s];
  if (!store_out) return plan.state_stride;
  int G = force_threads > 0 ? force_threads : plan.threads;
  std::vector<QkStat> stats(N);
  SimParams P;
  P.n = n; P.n_ops = (int)plan.ops.size(); P.ops = plan.ops.data(); P.cap = plan.cap.data();
  P.site_off = plan.site_off.data(); P.state_stride = plan.state_stride;
  P.X = X; P.ldx = ldx; P.N = N;
  P.store = (c128*)store_out; P.chi = chi_out; P.stats = stats.data();
  P.mode = trunc_mode; P.cutoff = trunc_error; P.fidelity_target = 1.0 - trunc_error; P.value_of_zero = 1e-16;
  P.tol = getenv("QK_EMU_TOL") ? atof(getenv("QK_EMU_TOL")) : 1e-15; P.max_sweeps = getenv("QK_EMU_SWEEPS") ? atoi(getenv("QK_EMU_SWEEPS")) : 60; P.rmax = plan.rmax; P.wr = plan.rmax * plan.rmax; P.trace = nullptr; P.early_exit = (flags & 2) ? 1 : 0;
  P.floor_rel = getenv("QK_EMU_FLOOR") ? atof(getenv("QK_EMU_FLOOR")) : 1e-28;
  P.abs_rel = getenv("QK_EMU_ABS") ? atof(getenv("QK_EMU_ABS")) : 0.0;
  std::vector<double> lam;
  P.parallel = plan.parallel; P.lam = nullptr; P.lam_ld = plan.rmax / 2; P.level_start = nullptr; P.n_levels = 0;
  if (plan.parallel) {
    lam.assign((size_t)N * (n + 1) * P.lam_ld, 0.0);
    P.lam = lam.data();
    P.level_start = plan.level_start.data();
    P.n_levels = (int)plan.level_start.size() - 1;
  }
  std::vector<c128> big_w, big_s;
  std::vector<int> big_flag(4, 0);
  P.big_w = nullptr; P.big_s = nullptr; P.big_w_stride = P.big_s_stride = 0; P.big_flag = nullptr; P.big_jb = plan.jb;
  P.big_wb_entries = (int)qk_big_wb_entries(plan.rmax, plan.jb);
  P.unit_clk = nullptr;
  if (plan.big) {
    big_w.resize((size_t)plan.rmax * plan.rmax);
    big_s.resize((size_t)plan.rmax * plan.rmax / 2);
    P.big_w = big_w.data(); P.big_s = big_s.data(); P.big_flag = big_flag.data();
  }
  size_t bytes = plan.big ? qk_big_smem_bytes(n, plan.rmax, plan.jb, G) : qk_sim_smem_bytes(n, plan.rmax, G);
  unsigned char* smem = (unsigned char*)aligned_alloc(64, (bytes + 63) & ~(size_t)63);
  for (int dp = 0; dp < N; ++dp) {
    memset(smem, 0xA5, bytes);   // poison: the core must not depend on stale shared memory
    switch (G) {
      case 32: run_group<32>(P, smem, dp); break;
      case 64: run_group<64>(P, smem, dp); break;
      case 128: run_group<128>(P, smem, dp); break;
      case 256: run_group<256>(P, smem, dp); break;
      default: free(smem); return QK_ERR_ARG;
    }
    if (stats_out) {
      stats_out[4 * dp + 0] = stats[dp].fidelity; stats_out[4 * dp + 1] = stats[dp].trunc_weight;
      stats_out[4 * dp + 2] = stats[dp].flags; stats_out[4 * dp + 3] = stats[dp].sweeps;
    }
  }
  free(smem);
  return plan.state_stride;
}

}  // extern "C"
