// Stage 1 for bond dimensions above the shared-memory-resident limit (32 < chi_cap <= 256): BASELINE config 4
// ("165 qubits, 4 layers, distance 4, high bond dimension").  Replaces the same reference calls as qk_sim_core.h
// (KernelPkg/src/KernelPkg.jl:68 apply(..; cutoff); gpu_backend/kernel_state_ansatz.py:221 simulate(.., MPSxGate, ..)),
// whose published runs reach chi ~ 100-600 (runs/qubit_scaling/results.csv:2-19, runs/crossover/gpu_results.csv:2-7).
//
// One thread-block CLUSTER of `ncta` CTAs works on one datapoint.  theta = gate * (A_k A_{k+1}) is (2 chi)^2 complex
// numbers -- up to 4 MB -- so it lives in global memory (one L2-resident slot per cluster) and the SVD is a BLOCK
// one-sided Jacobi: the columns are cut into blocks of `jb`; a round of the block tournament pairs the blocks up,
// every CTA of the cluster takes block pairs, copies their 2 jb columns into shared memory, runs one full
// tournament among them there (the same register-resident pair step as the small path, qk_jacobi_sweep) and copies
// them back; a cluster barrier ends the round.  L2 traffic per sweep drops by the 2 jb - 1 rounds done per visit.
// Everything else of an op (contraction, column norms, sort, the reference truncation rules, recovery of the other
// factor from W^dag theta) is the algorithm of qk_sim_core.h with the work dealt over all threads of the cluster;
// small decisions (sort, truncation) are recomputed identically by every CTA instead of being broadcast.
// Gauge moves are expressed as an SVD of the site pair with an identity gate and no truncation (QK_OP_ID2).
//
// Data written by one CTA and read by another always goes through L2 (ld.global.cg) and is separated by a
// cluster barrier.  Like qk_sim_core.h the code is a sequence of phases that the host emulation (tests/host_emu,
// one CTA per cluster) runs as loops over thread ids.
#pragma once
#include "qk_sim_core.h"

#if defined(__CUDACC__) && !defined(QK_HOST_EMU)
#define QK_CPAR_BEGIN(gt) { const int gt = c.cta * G + (int)threadIdx.x;
#define QK_CPAR_END } QK_CSYNC(c);
// Loads of data another CTA of the cluster may have written before the last cluster barrier.  QK_CSYNC fences on
// both sides of the barrier, so plain (L1-cached) loads are ordered after those writes by the PTX memory model;
// -DQK_BIG_LDCG=1 switches to L2-only loads (experiments).
#if defined(QK_BIG_LDCG) && QK_BIG_LDCG
QK_DEV c128 qk_ld(const c128* p) { const double2 v = __ldcg((const double2*)p); return cmake(v.x, v.y); }
#else
QK_DEV c128 qk_ld(const c128* p) { return *p; }
#endif
QK_DEV int qk_ld_flag(const int* p) { return __ldcg(p); }
#else
#define QK_CPAR_BEGIN(gt) for (int gt = 0; gt < G; ++gt) {
#define QK_CPAR_END }
#ifndef QK_CSYNC
#define QK_CSYNC(c) ((void)0)
#endif
QK_DEV c128 qk_ld(const c128* p) { return *p; }
QK_DEV int qk_ld_flag(const int* p) { return *p; }
#endif

QK_HD size_t qk_big_smem_bytes(int n, int rmax, int jb, int G) {
  const int capmax = rmax / 2;
  size_t b = qk_big_wb_entries(rmax, jb) * sizeof(c128);              // Wb: one pair of column blocks
  b += (size_t)4 * (G > capmax ? G : capmax) * sizeof(c128);          // ef
  b += (size_t)4 * G * sizeof(double);                                // scr
  b += (size_t)rmax * sizeof(double);                                 // nrm2
  b += 32 * sizeof(c128);                                             // gate + fused-gate accumulator
  b += (size_t)rmax * sizeof(c128);                                   // diag
  b += (size_t)(n + 2) * sizeof(double);                              // x
  b += sizeof(SimShared);
  b += (size_t)2 * rmax * sizeof(int);                                // order, swp
  b += (size_t)(n + 2) * sizeof(int);                                 // chi
  return (b + 15) & ~(size_t)15;
}

QK_DEV void qk_big_carve(SimCtx& c, const SimParams* P, unsigned char* smem, int G, int cluster_slot, int cta, int ncta) {
  const int rmax = P->rmax, capmax = rmax / 2, jb = P->big_jb;
  c.P = P;
  c.Wb = (c128*)smem;
  c.ef = c.Wb + qk_big_wb_entries(rmax, jb);
  c.scr = (double*)(c.ef + 4 * (G > capmax ? G : capmax));
  c.nrm2 = c.scr + 4 * G;
  c.gate = (c128*)(c.nrm2 + rmax + (rmax & 1));
  c.gacc = c.gate + 16;
  c.diag = c.gacc + 16;
  c.x = (double*)(c.diag + rmax);
  c.sh = (SimShared*)(c.x + P->n + (P->n & 1));
  c.order = (int*)(c.sh + 1);
  c.swp = c.order + rmax;
  c.chi = c.swp + rmax;
  c.W = P->big_w + (size_t)cluster_slot * P->big_w_stride;
  c.S = P->big_s + (size_t)cluster_slot * P->big_s_stride;
  c.gflag = P->big_flag + (size_t)cluster_slot * 4;
  c.cta = cta;
  c.ncta = ncta;
}

// ------------------------------------------------------------------------------------------------
// 1-qubit gate, work dealt over the cluster
// ------------------------------------------------------------------------------------------------
template <int G>
QK_DEV void qk_big_op_1q(SimCtx& c, const QkOp& op) {
  const int s = op.site;
  const int cl = c.chi[s], cr = c.chi[s + 1];
  c128 u00, u01, u10, u11;
  if (op.kind == QK_OP_H) {
    const double h = 0.70710678118654752440;
    u00 = cmake(h, 0); u01 = cmake(h, 0); u10 = cmake(h, 0); u11 = cmake(-h, 0);
  } else {
    const double th = qk_angle(op, c.x);
    const double cs = cos(th), sn = sin(th);
    if (op.kind == QK_OP_RZ) {
      u00 = cmake(cs, -sn); u01 = cmake(0, 0); u10 = cmake(0, 0); u11 = cmake(cs, sn);
    } else {  // RX
      u00 = cmake(cs, 0); u01 = cmake(0, -sn); u10 = cmake(0, -sn); u11 = cmake(cs, 0);
    }
  }
  c128* A = qk_site(c, s);
  const int GT = G * c.ncta;
  QK_CPAR_BEGIN(gt)
    for (int idx = gt; idx < cl * cr; idx += GT) {
      const int a = idx / cr, b = idx - a * cr;
      const c128 v0 = qk_ld(&A[(a * 2 + 0) * cr + b]), v1 = qk_ld(&A[(a * 2 + 1) * cr + b]);
      A[(a * 2 + 0) * cr + b] = cadd(cmul(u00, v0), cmul(u01, v1));
      A[(a * 2 + 1) * cr + b] = cadd(cmul(u10, v0), cmul(u11, v1));
    }
  QK_CPAR_END
}

// ------------------------------------------------------------------------------------------------
// block one-sided Jacobi on W (R x C, column-major, global memory)
// ------------------------------------------------------------------------------------------------
template <int G>
QK_DEV void qk_big_jacobi(SimCtx& c, int R, int C) {
  c128* W = c.W;
  c128* Wb = c.Wb;
  const int ldw = R;
  int jb = c.P->big_jb;                                      // columns per block: what fits for this row count
  if ((size_t)2 * jb * R > (size_t)c.P->big_wb_entries) jb = (int)((size_t)c.P->big_wb_entries / ((size_t)2 * R));
  if (jb < 1) jb = 1;
  const double tol2 = c.P->tol * c.P->tol;
  // total weight, computed identically by every CTA
  QK_PAR_BEGIN(tid)
    double s = 0.0;
    for (int i = tid; i < R * C; i += G) { const c128 v = qk_ld(W + i); s += v.x * v.x + v.y * v.y; }
    c.scr[tid] = s;
    if (tid == 0) c.sh->rotated = 0;
  QK_PAR_END
  double total = 0.0;
  for (int t = 0; t < G; ++t) total += c.scr[t];
  QK_BARRIER();
  QK_PAR_BEGIN(tid)
    if (tid == 0) c.sh->total = total;
  QK_PAR_END
  const double floor2 = c.P->floor_rel * total;
  const double abs2 = c.P->abs_rel * c.P->abs_rel * total * total;
  const int nb = (C + jb - 1) / jb;            // column blocks
  const int nbe = (nb + 1) & ~1;
  const int nslots = nb > 1 ? nbe / 2 : 1;
  const int nrounds = nb > 1 ? nbe - 1 : 1;
  int sweep = 0;
  if (C >= 2) {
    for (; sweep < c.P->max_sweeps; ++sweep) {
      int* flag = c.gflag + (sweep % 3);
      for (int r = 0; r < nrounds; ++r) {
        for (int slot = c.cta; slot < nslots; slot += c.ncta) {
          int p = 0, q = -1;
          if (nb > 1 && !qk_rr_pair(slot, r, nbe, nb, p, q)) continue;      // the bye of an odd block count
          const int c0p = p * jb, np = (C - c0p < jb) ? C - c0p : jb;
          const int c0q = q >= 0 ? q * jb : 0;
          const int nq = q >= 0 ? ((C - c0q < jb) ? C - c0q : jb) : 0;
          const int nc = np + nq;
          QK_PAR_BEGIN(tid)
            for (int i = tid; i < R * nc; i += G) {
              const int col = i / R, row = i - col * R;
              const int src = col < np ? c0p + col : c0q + (col - np);
              Wb[i] = qk_ld(W + (size_t)src * ldw + row);
            }
          QK_PAR_END
          qk_jacobi_sweep<G, true>(c, Wb, R, R, nc, tol2, floor2, abs2);
          QK_PAR_BEGIN(tid)
            for (int i = tid; i < R * nc; i += G) {
              const int col = i / R, row = i - col * R;
              const int src = col < np ? c0p + col : c0q + (col - np);
              W[(size_t)src * ldw + row] = Wb[i];
            }
          QK_PAR_END
        }
        QK_PAR_BEGIN(tid)
          if (tid == 0) {
            if (r == nrounds - 1 && c.sh->rotated) *(volatile int*)flag = 1;
          }
        QK_PAR_END
        QK_CSYNC(c);
      }
      const int rot = qk_ld_flag(flag);
      QK_BARRIER();
      if (!rot) { ++sweep; break; }
      QK_PAR_BEGIN(tid)
        if (tid == 0) {
          c.sh->rotated = 0;
          // Three rotating flag slots.  The slot of sweep + 2 was last read after sweep - 1: every CTA is past that
          // read (it has arrived at the final barrier of this sweep), and its next writer is two sweeps away, behind
          // the barriers of sweep + 1.
          if (c.cta == 0) *(volatile int*)(c.gflag + ((sweep + 2) % 3)) = 0;
        }
      QK_PAR_END
    }
  }
  QK_PAR_BEGIN(tid)
    if (tid == 0) {
      c.sh->sweeps += sweep;
      if (C >= 2 && sweep >= c.P->max_sweeps) c.sh->flags |= QK_FLAG_NO_CONVERGE;
    }
  QK_PAR_END
}

// ------------------------------------------------------------------------------------------------
// 2-qubit gate on sites (k, k+1), sequential (orthogonality-centre) form; see qk_op_2q for the algebra
// ------------------------------------------------------------------------------------------------
template <int G>
QK_DEV void qk_big_op_2q(SimCtx& c, const QkOp& op) {
  if (op.pad & QK_OPF_CONT) {   // not the last gate of a fused group: only accumulate its matrix (every CTA its copy)
    QK_PAR_BEGIN(tid)
      if (tid == 0) {
        qk_build_gate_2q_fused(op, c.x, c.gate, c.gacc, (c128*)c.scr);
        for (int i = 0; i < 16; ++i) c.gacc[i] = c.gate[i];
      }
    QK_PAR_END
    return;
  }
  const int k = op.site;
  const int ca = c.chi[k], cb = c.chi[k + 1], cc = c.chi[k + 2];
  const int m = 2 * ca, n2 = 2 * cc;
  const bool transposed = (m < n2);
  const int R = transposed ? n2 : m;
  const int C = transposed ? m : n2;
  const int ldw = R;
  const int GT = G * c.ncta;
  c128* A = qk_site(c, k);       // [a][l][b]
  c128* B = qk_site(c, k + 1);   // [b][r][c]
  c128* W = c.W;
  c128* S = c.S;

  QK_PAR_BEGIN(tid)
    if (tid == 0) qk_build_gate_2q_fused(op, c.x, c.gate, c.gacc, (c128*)c.scr);
  QK_PAR_END

  // theta[(a,L),(R,c)] = sum_{l,r} g[(L,R),(l,r)] sum_b A[a,l,b] B[b,r,c]
  QK_CPAR_BEGIN(gt)
    if (gt == 0) { *(volatile int*)(c.gflag + 0) = 0; *(volatile int*)(c.gflag + 1) = 0; *(volatile int*)(c.gflag + 2) = 0; }
    for (int idx = gt; idx < ca * cc; idx += GT) {
      const int a = idx / cc, cidx = idx - a * cc;
      c128 t[4];
      for (int l = 0; l < 2; ++l)
        for (int r = 0; r < 2; ++r) {
          c128 acc = cmake(0, 0);
          const c128* ap = A + (size_t)(a * 2 + l) * cb;
          const c128* bp = B + (size_t)r * cc + cidx;
          for (int b = 0; b < cb; ++b) cfma(acc, qk_ld(ap + b), qk_ld(bp + (size_t)b * 2 * cc));
          t[l * 2 + r] = acc;
        }
      for (int L = 0; L < 2; ++L)
        for (int Rr = 0; Rr < 2; ++Rr) {
          c128 acc = cmake(0, 0);
          const c128* g = c.gate + (L * 2 + Rr) * 4;
          for (int lr = 0; lr < 4; ++lr) cfma(acc, g[lr], t[lr]);
          const int row = a * 2 + L, col = Rr * cc + cidx;
          if (!transposed) W[row + (size_t)col * ldw] = acc;
          else W[col + (size_t)row * ldw] = cconj(acc);
        }
    }
  QK_CPAR_END

  qk_big_jacobi<G>(c, R, C);

  // squared column norms (every CTA computes all of them: 8 threads per column)
  const int TPC = 8;
  for (int j0 = 0; j0 < C; j0 += G / TPC) {
    QK_PAR_BEGIN(tid)
      const int j = j0 + tid / TPC, sl = tid % TPC;
      double s = 0.0;
      if (j < C) {
        const c128* w = W + (size_t)j * ldw;
        for (int row = sl; row < R; row += TPC) { const c128 v = qk_ld(w + row); s += v.x * v.x + v.y * v.y; }
      }
      c.scr[tid] = s;
    QK_PAR_END
    QK_PAR_BEGIN(tid)
      const int j = j0 + tid / TPC;
      if (tid % TPC == 0 && j < C) {
        double s = 0.0;
        for (int t = 0; t < TPC; ++t) s += c.scr[tid + t];
        c.nrm2[j] = s;
      }
    QK_PAR_END
  }
  QK_PAR_BEGIN(tid)
    const double dead = c.P->floor_rel * c.sh->total;
    for (int j = tid; j < C; j += G) {
      const double v = c.nrm2[j];
      int rk = 0;
      for (int i = 0; i < C; ++i) {
        const double u = c.nrm2[i];
        rk += (u > v) || (u == v && i < j);
      }
      c.order[rk] = j;
      const double sg = sqrt(v);
      c.diag[rk] = cmake(sg, (sg > 0.0 && (v > dead || rk == 0)) ? 1.0 / sg : 0.0);
    }
  QK_PAR_END
  QK_PAR_BEGIN(tid)
    if (tid == 0) {
      if (op.pad & QK_OPF_NOTRUNC) {
        // gauge move (ITensors orthogonalize!: no truncation): keep every column that is not numerically zero
        const double dead = c.P->floor_rel * c.sh->total;
        int kk = 0;
        while (kk < C && c.nrm2[c.order[kk]] > dead) ++kk;
        if (kk < 1) kk = 1;
        if (kk > c.P->cap[k + 1]) kk = c.P->cap[k + 1];
        c.sh->keep = kk;
        c.sh->renorm = 1.0;
      } else {
        qk_truncate(c, C, c.P->cap[k + 1], k + 1);
      }
    }
  QK_PAR_END

  const int keep = c.sh->keep;
  const double renorm = c.sh->renorm;
  const bool right = (op.dir == QK_DIR_RIGHT);
  // other factor from W_out^dag W_in = Sigma^2 V^dag, theta contracted in two halves (qk_op_2q); the tiles of t are
  // dealt to the CTAs of the cluster, the result is staged in S (global)
  int TT = G / cb;
  if (TT < 1) TT = 1;
  for (int t0 = c.cta * TT; t0 < keep; t0 += c.ncta * TT) {
    QK_PAR_BEGIN(tid)
      for (int u = tid; u < TT * cb; u += G) {
        const int tt = u / cb, b = u - tt * cb;
        const int t = t0 + tt;
        if (t < keep) {
          const c128* w = W + (size_t)c.order[t] * ldw;
          c128 e[4];
          e[0] = e[1] = e[2] = e[3] = cmake(0, 0);
          if (!transposed) {
            for (int a = 0; a < ca; ++a) {
              const c128 w0 = qk_ld(w + a * 2), w1 = qk_ld(w + a * 2 + 1);
              const c128 a0 = qk_ld(&A[(size_t)(a * 2) * cb + b]), a1 = qk_ld(&A[(size_t)(a * 2 + 1) * cb + b]);
              cfmac(e[0], w0, a0); cfmac(e[1], w0, a1);     // e[L*2+l] += conj(W[(a,L),t]) A[a,l,b]
              cfmac(e[2], w1, a0); cfmac(e[3], w1, a1);
            }
          } else {
            const c128* b0 = B + (size_t)(b * 2) * cc;
            const c128* b1 = b0 + cc;
            for (int cidx = 0; cidx < cc; ++cidx) {
              const c128 w0 = qk_ld(w + cidx), w1 = qk_ld(w + cc + cidx);
              const c128 v0 = qk_ld(b0 + cidx), v1 = qk_ld(b1 + cidx);
              cfma(e[0], v0, w0); cfma(e[1], v1, w0);       // e[R*2+r] += B[b,r,c] W[(R,c),t]
              cfma(e[2], v0, w1); cfma(e[3], v1, w1);
            }
          }
          c128* f = c.ef + (size_t)u * 4;
          if (!transposed) {
            for (int Rr = 0; Rr < 2; ++Rr)
              for (int r = 0; r < 2; ++r) {
                c128 acc = cmake(0, 0);
                for (int L = 0; L < 2; ++L)
                  for (int l = 0; l < 2; ++l) cfma(acc, c.gate[(L * 2 + Rr) * 4 + (l * 2 + r)], e[L * 2 + l]);
                f[Rr * 2 + r] = acc;
              }
          } else {
            for (int L = 0; L < 2; ++L)
              for (int l = 0; l < 2; ++l) {
                c128 acc = cmake(0, 0);
                for (int Rr = 0; Rr < 2; ++Rr)
                  for (int r = 0; r < 2; ++r) cfma(acc, c.gate[(L * 2 + Rr) * 4 + (l * 2 + r)], e[Rr * 2 + r]);
                f[L * 2 + l] = acc;
              }
          }
        }
      }
    QK_PAR_END
    QK_PAR_BEGIN(tid)
      if (!transposed) {
        for (int idx = tid; idx < TT * n2; idx += G) {
          const int tt = idx / n2, col = idx - tt * n2;
          const int t = t0 + tt;
          if (t < keep) {
            const int Rr = col / cc, cidx = col - Rr * cc;
            const c128* f = c.ef + (size_t)tt * cb * 4 + Rr * 2;
            c128 acc = cmake(0, 0);
            for (int b = 0; b < cb; ++b) {
              cfma(acc, f[(size_t)b * 4], qk_ld(&B[(size_t)(b * 2) * cc + cidx]));
              cfma(acc, f[(size_t)b * 4 + 1], qk_ld(&B[(size_t)(b * 2 + 1) * cc + cidx]));
            }
            const double isg = c.diag[t].y;
            S[(size_t)t * n2 + col] = cscale(acc, right ? renorm * isg : isg * isg);
          }
        }
      } else {
        for (int idx = tid; idx < TT * m; idx += G) {
          const int tt = idx / m, row = idx - tt * m;
          const int t = t0 + tt;
          if (t < keep) {
            const int a = row >> 1, L = row & 1;
            const c128* f = c.ef + (size_t)tt * cb * 4 + L * 2;
            c128 acc = cmake(0, 0);
            for (int b = 0; b < cb; ++b) {
              cfma(acc, qk_ld(&A[(size_t)(a * 2) * cb + b]), f[(size_t)b * 4]);
              cfma(acc, qk_ld(&A[(size_t)(a * 2 + 1) * cb + b]), f[(size_t)b * 4 + 1]);
            }
            const double isg = c.diag[t].y;
            S[(size_t)row * keep + t] = cscale(acc, right ? isg * isg : renorm * isg);
          }
        }
      }
    QK_PAR_END
  }
  QK_CSYNC(c);   // S is complete and nobody reads the old site tensors any more

  QK_CPAR_BEGIN(gt)
    // left site  [a][L][t]  (m x keep),  right site [t][R][c]  (keep x n2)
    for (int idx = gt; idx < m * keep; idx += GT) {
      const int row = idx / keep, t = idx - row * keep;
      const double isg = c.diag[t].y;
      c128 v;
      if (!transposed) v = cscale(qk_ld(&W[row + (size_t)c.order[t] * ldw]), right ? isg : renorm);   // W = U S
      else v = qk_ld(&S[idx]);
      A[idx] = v;
    }
    for (int idx = gt; idx < keep * n2; idx += GT) {
      const int t = idx / n2, col = idx - t * n2;
      const double isg = c.diag[t].y;
      c128 v;
      if (!transposed) v = qk_ld(&S[idx]);
      else v = cscale(cconj(qk_ld(&W[col + (size_t)c.order[t] * ldw])), right ? renorm : isg);   // S V^dag = W^dag
      B[idx] = v;
    }
  QK_CPAR_END
  QK_PAR_BEGIN(tid)
    if (tid == 0) c.chi[k + 1] = keep;
  QK_PAR_END
}

// ------------------------------------------------------------------------------------------------
// whole circuit for datapoint dp (sequential schedule; gauge moves arrive as QK_OP_ID2 ops)
// ------------------------------------------------------------------------------------------------
template <int G>
QK_DEV void qk_big_datapoint(SimCtx& c, int dp) {
  const SimParams* P = c.P;
  const int n = P->n;
  const int GT = G * c.ncta;
  c.state = P->store + (size_t)dp * P->state_stride;
  QK_PAR_BEGIN(tid)
    for (int b = tid; b <= n; b += G) c.chi[b] = 1;
    for (int i = tid; i < n; i += G) c.x[i] = P->X[(size_t)dp * P->ldx + i];
    if (tid == 0) {
      c.sh->flags = 0; c.sh->sweeps = 0; c.sh->max_chi = 1;
      c.sh->fidelity = 1.0; c.sh->trunc_weight = 0.0; c.sh->rotated = 0;
    }
  QK_PAR_END
  QK_CPAR_BEGIN(gt)
    for (int s = gt; s < n; s += GT) {   // |0...0>, KernelPkg.jl:68
      c128* A = qk_site(c, s);
      A[0] = cmake(1, 0);
      A[1] = cmake(0, 0);
    }
  QK_CPAR_END
  for (int o = 0; o < P->n_ops; ++o) {
    const QkOp op = P->ops[o];
    if (op.kind <= QK_OP_RX) qk_big_op_1q<G>(c, op);
    else qk_big_op_2q<G>(c, op);
    if (P->early_exit && op.kind >= QK_OP_XX && (c.sh->flags & QK_FLAG_CAP_HIT)) break;
    if (P->trace) {
      QK_PAR_BEGIN(tid)
        if (tid == 0 && c.cta == 0) {
          double bytes = 0.0;
          for (int s = 0; s < n; ++s) bytes += 32.0 * c.chi[s] * c.chi[s + 1];
          P->trace[(size_t)dp * P->n_ops + o] = bytes;
        }
      QK_PAR_END
    }
  }
  QK_PAR_BEGIN(tid)
    if (c.cta == 0) {
      for (int b = tid; b <= n; b += G) P->chi[(size_t)dp * (n + 1) + b] = c.chi[b];
      if (tid == 0) {
        QkStat st;
        st.fidelity = c.sh->fidelity; st.trunc_weight = c.sh->trunc_weight;
        st.flags = c.sh->flags; st.sweeps = c.sh->sweeps; st.max_chi = c.sh->max_chi; st.pad = 0;
        P->stats[dp] = st;
      }
    }
  QK_PAR_END
  QK_CSYNC(c);
}
