// Static schedule compiler (host).  The op sequence -- which bond is updated, where the
// orthogonality centre has to be, which way each SVD absorbs its singular values -- is the same for
// every datapoint of a Gram-matrix job; only the angles differ.  It is compiled once here and then
// interpreted on the device by every cooperative group (qk_sim_core.h).
//
// Reference semantics followed: ITensors apply() orthogonalises the MPS to the gate before every
// gate and leaves the centre on the right site of a 2-site gate.  Gram entries are
// gauge invariant, so we keep the centre *adjacent* to the gate (either of its two sites) and pick
// the absorb direction by looking ahead to the next 2-qubit gate; this removes gauge moves without
// changing any singular value the truncation rule sees.
#include "qk_plan.h"
#include "qk_sim_core.h"
#include <math.h>

int qk_pick_threads(int chi_cap) {
  if (chi_cap <= 4) return 32;
  if (chi_cap <= 8) return 64;
  if (chi_cap <= 16) return 128;
  return 256;
}

int qk_ansatz_gates(int n, int reps, double gamma, int hadamard_init, const int32_t* pairs, int n_pairs,
                    std::vector<qk_gate>* out, std::string* err) {
  out->clear();
  if (n < 1 || reps < 0 || n_pairs < 0) { *err = "bad ansatz parameters"; return QK_ERR_ARG; }
  auto g1 = [&](int kind, int q, int fa, double coeff) {
    qk_gate g; g.kind = kind; g.q0 = q; g.q1 = -1; g.fa = fa; g.fb = -1; g.coeff = coeff; out->push_back(g);
  };
  auto g2 = [&](int kind, int q, int fa, int fb, double coeff) {
    qk_gate g; g.kind = kind; g.q0 = q; g.q1 = q + 1; g.fa = fa; g.fb = fb; g.coeff = coeff; out->push_back(g);
  };
  if (hadamard_init)
    for (int i = 0; i < n; ++i) g1(QK_GATE_H, i, -1, 0.0);                       // gpu:53-55
  for (int r = 0; r < reps; ++r) {
    for (int i = 0; i < n; ++i) g1(QK_GATE_RZ, i, i, (2.0 / M_PI) * gamma);      // gpu:57-60
    for (int e = 0; e < n_pairs; ++e) {                                           // gpu:62-66 + routing gpu:78-88
      int a = pairs[2 * e], b = pairs[2 * e + 1];
      if (a < 0 || b < 0 || a >= n || b >= n || a == b) { *err = "entanglement pair out of range"; return QK_ERR_ARG; }
      const int q0 = a < b ? a : b, q1 = a < b ? b : a;
      for (int q = q0; q < q1 - 1; ++q) g2(QK_GATE_SWAP, q, -1, -1, 0.0);
      g2(QK_GATE_XX, q1 - 1, a, b, gamma * gamma);
      for (int q = q1 - 2; q >= q0; --q) g2(QK_GATE_SWAP, q, -1, -1, 0.0);
    }
  }
  return QK_OK;
}

int qk_compile_plan(int n, const qk_gate* gates, int n_gates, int trunc_mode, double trunc_error, int chi_cap,
                    qk_plan* plan, std::string* err) {
  if (n < 1) { *err = "n_qubits must be >= 1"; return QK_ERR_ARG; }
  if (n_gates < 0 || (n_gates > 0 && !gates)) { *err = "bad gate list"; return QK_ERR_ARG; }
  if (trunc_mode != QK_TRUNC_ITENSORS && trunc_mode != QK_TRUNC_PYTKET) { *err = "bad truncation mode"; return QK_ERR_ARG; }
  if (!(trunc_error >= 0.0) || trunc_error >= 1.0) { *err = "truncation_error must be in [0, 1)"; return QK_ERR_ARG; }
  if (chi_cap < 1) { *err = "chi_cap must be >= 1"; return QK_ERR_ARG; }
  if (chi_cap > QK_CHI_LIMIT) {
    *err = "bond dimension cap above the shared-memory-resident limit (chi <= 32)";
    return QK_ERR_LIMIT;
  }
  plan->n = n; plan->n_gates = n_gates; plan->trunc_mode = trunc_mode; plan->trunc_error = trunc_error;
  plan->ops.clear(); plan->n_2q = plan->n_1q = plan->n_moves = 0;

  // validate + find, for every 2-qubit gate, the bond of the next one
  std::vector<int> next2q(n_gates, -1);
  int nxt = -1;
  for (int i = n_gates - 1; i >= 0; --i) {
    const qk_gate& g = gates[i];
    const bool two = (g.kind == QK_GATE_XX || g.kind == QK_GATE_ZZ || g.kind == QK_GATE_SWAP);
    const bool one = (g.kind == QK_GATE_H || g.kind == QK_GATE_RZ || g.kind == QK_GATE_RX);
    if (!one && !two) { *err = "Unrecognised gate."; return QK_ERR_ARG; }
    if (g.q0 < 0 || g.q0 >= n) { *err = "gate qubit out of range"; return QK_ERR_ARG; }
    if (two) {
      if (g.q1 != g.q0 + 1 || g.q1 >= n) { *err = "two-qubit gates must act on adjacent sites (q, q+1)"; return QK_ERR_ARG; }
    }
    if (g.kind != QK_GATE_H && g.kind != QK_GATE_SWAP && g.fa >= 0) {
      if (g.fa >= n || (two && (g.fb < 0 || g.fb >= n))) { *err = "feature index out of range"; return QK_ERR_ARG; }
    }
    next2q[i] = nxt;
    if (two) nxt = g.q0;
  }

  int centre = -1;   // -1: product state, every site is both left- and right-orthonormal
  for (int i = 0; i < n_gates; ++i) {
    const qk_gate& g = gates[i];
    QkOp op; op.kind = g.kind; op.site = g.q0; op.fa = g.fa; op.fb = g.fb; op.dir = QK_DIR_RIGHT; op.pad = 0; op.coeff = g.coeff;
    const bool two = (g.kind == QK_GATE_XX || g.kind == QK_GATE_ZZ || g.kind == QK_GATE_SWAP);
    if (!two) { plan->ops.push_back(op); plan->n_1q++; continue; }
    const int k = g.q0;
    if (centre >= 0) {
      for (; centre < k; ++centre) {       // centre left of the pair: QR moves to the right
        QkOp mv; mv.kind = QK_OP_MOVE_R; mv.site = centre; mv.fa = mv.fb = -1; mv.dir = 0; mv.pad = 0; mv.coeff = 0.0;
        plan->ops.push_back(mv); plan->n_moves++;
      }
      for (; centre > k + 1; --centre) {   // centre right of the pair: LQ moves to the left
        QkOp mv; mv.kind = QK_OP_MOVE_L; mv.site = centre; mv.fa = mv.fb = -1; mv.dir = 0; mv.pad = 0; mv.coeff = 0.0;
        plan->ops.push_back(mv); plan->n_moves++;
      }
    }
    const int k2 = next2q[i];
    op.dir = (k2 >= 0 && k2 + 1 <= k) ? QK_DIR_LEFT : QK_DIR_RIGHT;
    centre = (op.dir == QK_DIR_LEFT) ? k : k + 1;
    plan->ops.push_back(op); plan->n_2q++;
  }

  // bond caps: user cap, clipped by the chain-edge bound 2^min(b, n-b)
  plan->chi_cap = chi_cap;
  plan->cap.assign(n + 1, 1);
  for (int b = 0; b <= n; ++b) {
    const int e = b < n - b ? b : n - b;
    long long edge = (e >= 30) ? (1LL << 30) : (1LL << e);
    plan->cap[b] = (int32_t)(edge < chi_cap ? edge : chi_cap);
  }
  plan->site_off.assign(n + 1, 0);
  for (int s = 0; s < n; ++s) plan->site_off[s + 1] = plan->site_off[s] + (int64_t)plan->cap[s] * 2 * plan->cap[s + 1];
  plan->state_stride = plan->site_off[n];
  int capmax = 1;
  for (int b = 0; b <= n; ++b) if (plan->cap[b] > capmax) capmax = plan->cap[b];
  plan->rmax = 2 * capmax;
  plan->threads = qk_pick_threads(capmax);
  plan->smem_bytes = qk_sim_smem_bytes(n, plan->rmax, plan->threads);
  return QK_OK;
}
