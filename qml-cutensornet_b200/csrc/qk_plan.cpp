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
#include "qk_sim_big.h"
#include <math.h>
#include <stdlib.h>
#include <algorithm>

int qk_pick_threads(int chi_cap) {
  if (const char* e = getenv("QK_SIM_THREADS")) {   // tuning / experiments only
    const int g = atoi(e);
    if (g == 32 || g == 64 || g == 128 || g == 256) return g;
  }
  if (chi_cap <= 8) return 32;   // matrices <= 16x16: one warp, no block-level barriers to wait on (measured faster than 64)
  if (chi_cap <= 16) return 64;  // 16 column pairs x 4 threads, 8 rows per thread in registers: lower latency than 128
                                 // threads (two shuffle stages instead of three) and twice the datapoints per SM
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

// ------------------------------------------------------------------------------------------------
// Commutation-aware reordering.  A routed interaction  SWAP(q0,q0+1) .. SWAP(q1-2,q1-1), G(q1-1,q1),
// SWAP(q1-2,q1-1) .. SWAP(q0,q0+1)  (gpu_backend/kernel_state_ansatz.py:78-88) acts as G on the logical
// pair (q0, q1) and leaves every qubit in place.  All XXPhase interactions commute with each other
// (functions of X operators only), likewise all ZZPhase ones, so a maximal run of
// such units may be applied in any order.  The reference has no canonical order either (pytket's
// get_commands() is a topological sort and pytket-cutensornet re-sorts gates).  We order each run as
// one sweep, left-to-right or right-to-left depending on where the previous run ended, which keeps
// the orthogonality centre adjacent to every gate and removes nearly all gauge moves.
// ------------------------------------------------------------------------------------------------
struct QkUnit { int first, count; int kind; int lo, hi; };   // gates [first, first+count)

static void qk_reorder_commuting(std::vector<qk_gate>& g) {
  const int ng = (int)g.size();
  std::vector<QkUnit> units;
  for (int i = 0; i < ng;) {
    int k = 0;
    const int q0 = g[i].q0;
    while (i + k < ng && g[i + k].kind == QK_GATE_SWAP && g[i + k].q0 == q0 + k && g[i + k].q1 == q0 + k + 1) ++k;
    bool ok = false;
    const int c = i + k;   // candidate centre gate
    if (c < ng && (g[c].kind == QK_GATE_XX || g[c].kind == QK_GATE_ZZ) && g[c].q1 == g[c].q0 + 1 &&
        (k == 0 || g[c].q0 == q0 + k) && c + k < ng) {
      ok = true;
      for (int j = 0; ok && j < k; ++j) {
        const int idx = c + 1 + j;
        if (idx >= ng || g[idx].kind != QK_GATE_SWAP || g[idx].q0 != q0 + k - 1 - j || g[idx].q1 != q0 + k - j) ok = false;
      }
    }
    if (ok) {
      QkUnit u; u.first = i; u.count = 2 * k + 1; u.kind = g[c].kind; u.lo = (k == 0) ? g[c].q0 : q0; u.hi = g[c].q1;
      units.push_back(u);
      i += u.count;
    } else {
      QkUnit u; u.first = i; u.count = 1; u.kind = -1; u.lo = g[i].q0; u.hi = g[i].q0;
      units.push_back(u);
      i += 1;
    }
  }
  std::vector<qk_gate> out;
  out.reserve(ng);
  int last_pos = 0;   // where the previous run of interactions ended (site index)
  const int n_units = (int)units.size();
  for (int u = 0; u < n_units;) {
    if (units[u].kind < 0) {
      out.push_back(g[units[u].first]);
      ++u;
      continue;
    }
    int v = u;
    while (v < n_units && units[v].kind == units[u].kind) ++v;
    std::vector<QkUnit> run(units.begin() + u, units.begin() + v);
    int lo_min = run[0].lo, hi_max = run[0].hi;
    for (const QkUnit& r : run) { if (r.lo < lo_min) lo_min = r.lo; if (r.hi > hi_max) hi_max = r.hi; }
    const bool ascending = (last_pos - lo_min) <= (hi_max - last_pos);
    std::stable_sort(run.begin(), run.end(), [ascending](const QkUnit& a, const QkUnit& b) {
      if (a.lo != b.lo) return ascending ? a.lo < b.lo : a.lo > b.lo;
      return ascending ? a.hi < b.hi : a.hi > b.hi;
    });
    for (const QkUnit& r : run)
      for (int j = 0; j < r.count; ++j) out.push_back(g[r.first + j]);
    last_pos = run.back().lo;
    u = v;
  }
  g.swap(out);
}

// ------------------------------------------------------------------------------------------------
// Routing of interleaved distance-2 interactions for the levelised (B form) schedule.  The ansatz routes every
// pair (a, a+2) on its own: SWAP(a,a+1) XX(a+1,a+2) SWAP(a,a+1)  (gpu_backend/kernel_state_ansatz.py:78-88), and
// its first distance-2 sub-layer interleaves pairs (a, a+2), (a+1, a+3) on four neighbouring sites: six
// dependent two-qubit ops.  One swap of the two middle qubits brings both pairs next to each other:
//     SWAP(a+1,a+2)  XX(a,a+1) || XX(a+2,a+3)  SWAP(a+1,a+2)
// -- the same unitary (XXPhase is symmetric and all interactions commute), four ops of dependency depth three.
// C3: 386 -> 290 two-qubit ops, depth 28 -> 16.
// ------------------------------------------------------------------------------------------------
static void qk_pair_distance2(std::vector<qk_gate>& g) {
  std::vector<qk_gate> out;
  out.reserve(g.size());
  const size_t ng = g.size();
  auto is_swap = [&](size_t i, int q) { return i < ng && g[i].kind == QK_GATE_SWAP && g[i].q0 == q && g[i].q1 == q + 1; };
  auto is_xx = [&](size_t i, int q) {
    return i < ng && (g[i].kind == QK_GATE_XX || g[i].kind == QK_GATE_ZZ) && g[i].q0 == q && g[i].q1 == q + 1;
  };
  for (size_t i = 0; i < ng;) {
    if (g[i].kind == QK_GATE_SWAP) {
      const int a = g[i].q0;
      if (is_swap(i, a) && is_xx(i + 1, a + 1) && is_swap(i + 2, a) && is_swap(i + 3, a + 1) && is_xx(i + 4, a + 2) &&
          is_swap(i + 5, a + 1) && g[i + 1].kind == g[i + 4].kind) {
        qk_gate sw = g[i + 3];                 // SWAP(a+1, a+2)
        qk_gate x1 = g[i + 1]; x1.q0 = a; x1.q1 = a + 1;          // pair (a, a+2): angle expression unchanged
        qk_gate x2 = g[i + 4];                                    // pair (a+1, a+3) sits on (a+2, a+3) already
        out.push_back(sw); out.push_back(x1); out.push_back(x2); out.push_back(sw);
        i += 6;
        continue;
      }
    }
    out.push_back(g[i]);
    ++i;
  }
  g.swap(out);
}

int qk_compile_plan(int n, const qk_gate* gates_in, int n_gates, int trunc_mode, double trunc_error, int chi_cap,
                    int flags, qk_plan* plan, std::string* err) {
  plan->reorder = (flags & QK_PLAN_LITERAL_ORDER) ? 0 : 1;
  plan->early_exit = (flags & QK_PLAN_EARLY_EXIT) ? 1 : 0;
  plan->fuse = (flags & QK_PLAN_NO_FUSION) ? 0 : 1;
  plan->big = (chi_cap > QK_CHI_LIMIT || (flags & QK_PLAN_BIG)) ? 1 : 0;
  // (the large-matrix kernel already spreads one datapoint over a cluster: no B form there)
  plan->parallel = ((flags & QK_PLAN_PARALLEL) && !plan->big) ? 1 : 0;
  if (plan->parallel) plan->reorder = 0;   // the literal order has the shallow dependency graph (C3: depth 28)
  plan->level_start.clear();
  if (n < 1) { *err = "n_qubits must be >= 1"; return QK_ERR_ARG; }
  if (n_gates < 0 || (n_gates > 0 && !gates_in)) { *err = "bad gate list"; return QK_ERR_ARG; }
  if (trunc_mode != QK_TRUNC_ITENSORS && trunc_mode != QK_TRUNC_PYTKET) { *err = "bad truncation mode"; return QK_ERR_ARG; }
  if (!(trunc_error >= 0.0) || trunc_error >= 1.0) { *err = "truncation_error must be in [0, 1)"; return QK_ERR_ARG; }
  if (chi_cap < 1) { *err = "chi_cap must be >= 1"; return QK_ERR_ARG; }
  if (chi_cap > QK_CHI_LIMIT_BIG) {
    *err = "bond dimension cap above the limit of the stage-1 kernels (chi <= 512)";
    return QK_ERR_LIMIT;
  }
  plan->n = n; plan->n_gates = n_gates; plan->trunc_mode = trunc_mode; plan->trunc_error = trunc_error;
  plan->ops.clear(); plan->n_2q = plan->n_1q = plan->n_moves = 0;
  std::vector<qk_gate> gate_vec(gates_in, gates_in + n_gates);
  for (int i = 0; i < n_gates; ++i) {   // validate before touching the order
    const qk_gate& g = gate_vec[i];
    const bool two = (g.kind == QK_GATE_XX || g.kind == QK_GATE_ZZ || g.kind == QK_GATE_SWAP);
    const bool one = (g.kind == QK_GATE_H || g.kind == QK_GATE_RZ || g.kind == QK_GATE_RX);
    if (!one && !two) { *err = "Unrecognised gate."; return QK_ERR_ARG; }
    if (g.q0 < 0 || g.q0 >= n) { *err = "gate qubit out of range"; return QK_ERR_ARG; }
    if (two && (g.q1 != g.q0 + 1 || g.q1 >= n)) { *err = "two-qubit gates must act on adjacent sites (q, q+1)"; return QK_ERR_ARG; }
  }
  if (plan->reorder) qk_reorder_commuting(gate_vec);
  if (plan->parallel && !(flags & QK_PLAN_LITERAL_ORDER)) {
    qk_pair_distance2(gate_vec);
    n_gates = (int)gate_vec.size();   // plan->n_gates keeps the circuit's own count
  }
  const qk_gate* gates = gate_vec.data();

  for (int i = 0; i < n_gates; ++i) {
    const qk_gate& g = gates[i];
    const bool two = (g.kind == QK_GATE_XX || g.kind == QK_GATE_ZZ || g.kind == QK_GATE_SWAP);
    if (g.kind != QK_GATE_H && g.kind != QK_GATE_SWAP && g.fa >= 0) {
      if (g.fa >= n || (two && (g.fb < 0 || g.fb >= n))) { *err = "feature index out of range"; return QK_ERR_ARG; }
    }
  }

  // Peephole fusion (not in literal mode): 2-qubit gates that follow each other on the same bond become one
  // SVD (product of their 4x4 matrices); two SWAPs in a row cancel.  With the sweep order above this turns
  // the per-interaction swap-in / swap-out routing of the reference into "walk the qubit out once, interact
  // on the way, walk it back once": distance 2: 4 -> 3 SVDs per qubit, distance 4: 16 -> 7.
  struct Item { bool two; int k; std::vector<qk_gate> g; };
  std::vector<Item> items;
  const size_t kMaxFuse = 4;
  for (int i = 0; i < n_gates; ++i) {
    const qk_gate& g = gates[i];
    const bool two = (g.kind == QK_GATE_XX || g.kind == QK_GATE_ZZ || g.kind == QK_GATE_SWAP);
    if (two && (plan->reorder || plan->parallel) && plan->fuse && !items.empty() && items.back().two && items.back().k == g.q0 &&
        items.back().g.size() < kMaxFuse) {
      std::vector<qk_gate>& grp = items.back().g;
      if (g.kind == QK_GATE_SWAP && grp.back().kind == QK_GATE_SWAP) {
        grp.pop_back();
        if (grp.empty()) items.pop_back();
      } else {
        grp.push_back(g);
      }
    } else {
      Item it; it.two = two; it.k = g.q0; it.g.push_back(g);
      items.push_back(it);
    }
  }
  const int n_items = (int)items.size();
  std::vector<int> next2q(n_items, -1);   // bond of the next 2-qubit group after item i
  int nxt = -1;
  for (int i = n_items - 1; i >= 0; --i) {
    next2q[i] = nxt;
    if (items[i].two) nxt = items[i].k;
  }

  if (plan->parallel) {
    // B form needs no orthogonality centre: an item only depends on the earlier items that touch its site(s).
    // Level = longest such chain; items of one level act on disjoint sites and may run concurrently.
    std::vector<int> ready(n, 0), level(n_items, 0);
    int n_levels = 0;
    for (int i = 0; i < n_items; ++i) {
      const Item& it = items[i];
      int lv = ready[it.k];
      if (it.two && ready[it.k + 1] > lv) lv = ready[it.k + 1];
      level[i] = lv;
      ready[it.k] = lv + 1;
      if (it.two) ready[it.k + 1] = lv + 1;
      if (lv + 1 > n_levels) n_levels = lv + 1;
    }
    plan->level_start.assign(n_levels + 1, 0);
    for (int lv = 0; lv < n_levels; ++lv) {
      plan->level_start[lv] = (int32_t)plan->ops.size();
      for (int i = 0; i < n_items; ++i) {
        if (level[i] != lv) continue;
        const Item& it = items[i];
        const int ng = (int)it.g.size();
        for (int j = 0; j < ng; ++j) {
          const qk_gate& g = it.g[j];
          QkOp op; op.kind = g.kind; op.site = it.k; op.fa = g.fa; op.fb = g.fb; op.dir = lv; op.coeff = g.coeff;
          op.pad = it.two ? ((j + 1 < ng ? QK_OPF_CONT : 0) | (j > 0 ? QK_OPF_ACC : 0)) : 0;
          plan->ops.push_back(op);
        }
        if (it.two) plan->n_2q++; else plan->n_1q++;
      }
    }
    plan->level_start[n_levels] = (int32_t)plan->ops.size();
  }
  int centre = -1;   // -1: product state, every site is both left- and right-orthonormal
  for (int i = 0; i < n_items && !plan->parallel; ++i) {
    const Item& it = items[i];
    if (!it.two) {
      const qk_gate& g = it.g[0];
      QkOp op; op.kind = g.kind; op.site = g.q0; op.fa = g.fa; op.fb = g.fb; op.dir = QK_DIR_RIGHT; op.pad = 0; op.coeff = g.coeff;
      plan->ops.push_back(op); plan->n_1q++;
      continue;
    }
    const int k = it.k;
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
    const int dir = (k2 >= 0 && k2 + 1 <= k) ? QK_DIR_LEFT : QK_DIR_RIGHT;
    centre = (dir == QK_DIR_LEFT) ? k : k + 1;
    const int ng = (int)it.g.size();
    for (int j = 0; j < ng; ++j) {
      const qk_gate& g = it.g[j];
      QkOp op; op.kind = g.kind; op.site = k; op.fa = g.fa; op.fb = g.fb; op.dir = dir; op.coeff = g.coeff;
      op.pad = (j + 1 < ng ? QK_OPF_CONT : 0) | (j > 0 ? QK_OPF_ACC : 0);
      plan->ops.push_back(op);
    }
    plan->n_2q++;
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
  if (plan->big) {
    // gauge moves become SVDs of the site pair with an identity gate and no truncation: MOVE_R of site s pushes the
    // centre to s + 1 (bond (s, s+1), singular values to the right), MOVE_L of site s to s - 1 (bond (s-1, s), left)
    for (QkOp& op : plan->ops) {
      if (op.kind == QK_OP_MOVE_R) { op.kind = QK_OP_ID2; op.dir = QK_DIR_RIGHT; op.pad = QK_OPF_NOTRUNC; }
      else if (op.kind == QK_OP_MOVE_L) { op.kind = QK_OP_ID2; op.site -= 1; op.dir = QK_DIR_LEFT; op.pad = QK_OPF_NOTRUNC; }
    }
    plan->threads = 256;
    plan->jb = 8;
    if (const char* e = getenv("QK_BIG_JB")) { const int v = atoi(e); if (v >= 1 && v <= 16) plan->jb = v; }   // tests
    plan->smem_bytes = qk_big_smem_bytes(n, plan->rmax, plan->jb, plan->threads);
  }
  return QK_OK;
}
