// Stage 1 core: one cooperative group of G threads simulates one datapoint's circuit as an MPS.
//
// Replaces what the reference delegates per gate to ITensors.apply (KernelPkg/src/KernelPkg.jl:68)
// and pytket-cutensornet simulate(.., MPSxGate, ..) (gpu_backend/kernel_state_ansatz.py:221):
//   * 1-qubit gate      -> multiply into the site tensor
//   * gauge move        -> Householder QR / LQ of the centre site, factor pushed into the neighbour
//   * 2-qubit gate      -> contract the site pair with the gate, one-sided Jacobi (Hestenes) SVD in
//                          shared memory, reference truncation rule, write both sites back
//
// The code is written as a sequence of "parallel phases" (QK_PAR_BEGIN .. QK_PAR_END); on the device
// a phase runs once per thread and ends in __syncthreads(); in the host emulation build
// (tests/host_emu, test infrastructure only) the same phase is a loop over the G thread ids.
// Rules that keep both builds equivalent:
//   - per-thread state never lives across a phase boundary (it is recomputed or kept in shared memory)
//   - code outside phases is uniform: it only reads shared scalars written before the last barrier
//   - a shared scalar that is re-written in the very next phase is read, then QK_BARRIER()
#pragma once
#include <math.h>
#include "qk_types.h"

#if defined(__CUDACC__) && !defined(QK_HOST_EMU)
#define QK_DEV __device__ __forceinline__
#define QK_PAR_BEGIN(tid) { const int tid = (int)threadIdx.x;
#define QK_PAR_END } __syncthreads();
#define QK_BARRIER() __syncthreads()
#else
#define QK_DEV inline
#define QK_PAR_BEGIN(tid) for (int tid = 0; tid < G; ++tid) {
#define QK_PAR_END }
#define QK_BARRIER() ((void)0)
#endif

#ifndef QK_PI
#define QK_PI 3.14159265358979323846
#endif
// reduction scratch doubles per thread: the device reduces the Jacobi dot products with warp shuffles
// and only needs 2 (Householder); the host emulation stages 4 partial sums per thread
#if defined(QK_HOST_EMU)
#define QK_SCR_PER_THREAD 4
#else
#define QK_SCR_PER_THREAD 2
#endif

struct SimShared {
  int rotated;
  int keep;
  int hh_skip;
  int flags;
  int sweeps;
  int max_chi;
  double renorm;
  double total;     // Frobenius norm^2 of the matrix the last Jacobi ran on
  double fidelity;
  double trunc_weight;
  double f0;
  c128 z0;
};

struct SimCtx {
  const SimParams* P;
  c128* W;        // working matrix (column-major), rmax^2 entries; its tail doubles as staging
  c128* ef;       // [4*G] tile of the gate-folded half contraction (qk_op_2q)
  double* scr;    // 4*G doubles reduction scratch
  double* nrm2;   // [rmax]
  int* order;     // [rmax]
  int* swp;       // [rmax] column swaps that compact the kept columns
  c128* gate;     // [16]
  c128* gacc;     // [16] product of the gates of a fused group so far
  c128* diag;     // [rmax] diagonal of R (Householder)
  int* chi;       // [n+1] bond dimensions of the current datapoint
  double* x;      // [n] features of the current datapoint
  SimShared* sh;
  c128* state;    // global: this datapoint's site slots
  double* lam;    // global: this datapoint's bond weights [(n+1)][lam_ld] (B form only)
  // large-matrix path (qk_sim_big.h): W above points to global memory
  c128* Wb;       // shared: one pair of column blocks of W (rmax x 2 jb)
  c128* S;        // global: staging of the recovered factor
  int* gflag;     // global: [3] rotating "a rotation happened" flags of the cluster
  int cta, ncta;  // rank / size of the CTA group that shares the datapoint
};

// bytes of shared memory the core needs for group size G
QK_HD size_t qk_sim_smem_bytes(int n, int rmax, int G) {
  size_t wr = (size_t)rmax * rmax;
  size_t b = wr * sizeof(c128);             // W
  b += (size_t)4 * G * sizeof(c128);        // ef
  b += (size_t)QK_SCR_PER_THREAD * G * sizeof(double);   // scr
  b += (size_t)rmax * sizeof(double);       // nrm2
  b += 32 * sizeof(c128);                   // gate + fused-gate accumulator
  b += (size_t)rmax * sizeof(c128);         // diag
  b += (size_t)(n + 1) * sizeof(double);    // x (+pad)
  b += sizeof(SimShared);
  b += (size_t)2 * rmax * sizeof(int);      // order, swp
  b += (size_t)(n + 1) * sizeof(int);       // chi
  return (b + 15) & ~(size_t)15;
}

QK_DEV void qk_sim_carve(SimCtx& c, const SimParams* P, unsigned char* smem, int G) {
  const size_t wr = (size_t)P->rmax * P->rmax;
  c.P = P;
  c.W = (c128*)smem;
  c.ef = c.W + wr;
  c.scr = (double*)(c.ef + 4 * G);
  c.nrm2 = c.scr + QK_SCR_PER_THREAD * G;
  c.gate = (c128*)(c.nrm2 + P->rmax + (P->rmax & 1));
  c.gacc = c.gate + 16;
  c.diag = c.gacc + 16;
  c.x = (double*)(c.diag + P->rmax);
  c.sh = (SimShared*)(c.x + P->n + (P->n & 1));
  c.order = (int*)(c.sh + 1);
  c.swp = c.order + P->rmax;
  c.chi = c.swp + P->rmax;
}

QK_DEV c128* qk_site(const SimCtx& c, int s) { return c.state + c.P->site_off[s]; }

QK_DEV double qk_angle(const QkOp& op, const double* x) {
  // alpha in half-turns (KernelPkg.jl:9,17,25,35: theta = pi*alpha/2)
  double alpha;
  if (op.fa < 0) alpha = op.coeff;
  else if (op.kind == QK_OP_XX || op.kind == QK_OP_ZZ) alpha = op.coeff * (1.0 - x[op.fa]) * (1.0 - x[op.fb]);
  else alpha = op.coeff * x[op.fa];
  return QK_PI * alpha / 2.0;
}

// ------------------------------------------------------------------------------------------------
// 1-qubit gates (KernelPkg.jl:8-22 and ITensors "H")
// ------------------------------------------------------------------------------------------------
template <int G>
QK_DEV void qk_op_1q(SimCtx& c, const QkOp& op) {
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
  QK_PAR_BEGIN(tid)
    for (int idx = tid; idx < cl * cr; idx += G) {
      const int a = idx / cr, b = idx - a * cr;
      const c128 v0 = A[(a * 2 + 0) * cr + b], v1 = A[(a * 2 + 1) * cr + b];
      A[(a * 2 + 0) * cr + b] = cadd(cmul(u00, v0), cmul(u01, v1));
      A[(a * 2 + 1) * cr + b] = cadd(cmul(u10, v0), cmul(u11, v1));
    }
  QK_PAR_END
}

// 4x4 gate [(L,R),(l,r)], first site = more significant bit (KernelPkg.jl:24-42, ITensors "SWAP")
QK_DEV void qk_build_gate_2q(const QkOp& op, const double* x, c128* g) {
  for (int i = 0; i < 16; ++i) g[i] = cmake(0, 0);
  if (op.kind == QK_OP_SWAP) {
    g[0 * 4 + 0] = g[1 * 4 + 2] = g[2 * 4 + 1] = g[3 * 4 + 3] = cmake(1, 0);
    return;
  }
  if (op.kind == QK_OP_ID2) {
    g[0] = g[5] = g[10] = g[15] = cmake(1, 0);
    return;
  }
  const double th = qk_angle(op, x);
  const double cs = cos(th), sn = sin(th);
  if (op.kind == QK_OP_XX) {
    for (int i = 0; i < 4; ++i) { g[i * 4 + i] = cmake(cs, 0); g[i * 4 + (3 - i)] = cmake(0, -sn); }
  } else {  // ZZ
    g[0] = cmake(cs, -sn); g[5] = cmake(cs, sn); g[10] = cmake(cs, sn); g[15] = cmake(cs, -sn);
  }
}

// gate of this op, composed with the accumulated gate of its fused group (applied earlier => on the right);
// `tmp` is 16 c128 of shared scratch (a local array would cost every thread 256 B of stack)
QK_DEV void qk_build_gate_2q_fused(const QkOp& op, const double* x, c128* g, const c128* gacc, c128* tmp) {
  qk_build_gate_2q(op, x, g);
  if (op.pad & QK_OPF_ACC) {
    for (int i = 0; i < 4; ++i)
      for (int j = 0; j < 4; ++j) {
        c128 acc = cmake(0, 0);
        for (int k = 0; k < 4; ++k) cfma(acc, g[i * 4 + k], gacc[k * 4 + j]);
        tmp[i * 4 + j] = acc;
      }
    for (int i = 0; i < 16; ++i) g[i] = tmp[i];
  }
}

// ------------------------------------------------------------------------------------------------
// one-sided Jacobi (Hestenes) on the R x C matrix W (column-major, ld = R):  W_out = W_in * V, V unitary.
// At convergence the columns of W are orthogonal (= U Sigma).  V is NOT accumulated: a second C x C matrix
// would double the shared memory per datapoint (and with it halve the datapoints resident per SM) and the
// rotations applied to it are a third of the work of a round; qk_op_2q recovers the factor it needs from
// W_out^dag W_in = Sigma^2 V^dag through a half contraction with the site tensors instead.
// ------------------------------------------------------------------------------------------------
QK_DEV bool qk_rr_pair(int i, int r, int Ce, int C, int& p, int& q) {
  // round-robin tournament: Ce (even) players, round r in [0, Ce-1), pair slot i in [0, Ce/2)
  const int m = Ce - 1;
  if (i == 0) { p = m; q = r; }
  else {
    p = r + i; if (p >= m) p -= m;
    q = r - i; if (q < 0) q += m;
  }
  if (p > q) { const int t = p; p = q; q = t; }
  return q < C;
}

QK_DEV double qk_rsqrt(double x) {
#if defined(__CUDA_ARCH__)
  return rsqrt(x);
#else
  return 1.0 / sqrt(x);
#endif
}
QK_DEV double qk_rcp(double x) {
#if defined(__CUDA_ARCH__)
  return __drcp_rn(x);
#else
  return 1.0 / x;
#endif
}

// Rotation that orthogonalises columns x, y with a = |x|^2, b = |y|^2, gamma = x^dag y = (gr, gi):
//   x' = cs x - conj(f) y ,  y' = f x + cs y ,  f = sin(theta) e^{i phi} sign(b - a),  cs = cos(theta),
//   tan(2 theta) = 2|gamma| / |b - a|,  |theta| <= pi/4.   Returns false if no rotation is needed.
// Written with two dependent rsqrt only (the parameter chain is the latency floor of a Jacobi round):
//   h = sqrt((b-a)^2 + 4|gamma|^2);  cos(2 theta) = |b-a|/h;  sin(2 theta) e^{i phi} = 2 gamma / h;
//   cos^2(theta) = (1 + cos 2theta)/2;  sin(theta) = sin(2 theta) / (2 cos(theta)).
// Return value: 0 = no rotation, 1 = rotation with relative off-diagonal below QK_QUAD_EPS, 2 = above.
// One-sided Jacobi converges quadratically per sweep, so a sweep whose largest relative off-diagonal
// was below QK_QUAD_EPS = 1e-9 leaves every pair below ~1e-18 after its own rotations: the usual
// "empty" verification sweep (a quarter of the work at ~3 sweeps per SVD) is not needed.
#define QK_QUAD_EPS2 1e-18
QK_DEV int qk_rotation(double a, double b, double gr, double gi, double tol2, double floor2, double abs2, double& cs,
                       c128& f) {
  const double g2 = gr * gr + gi * gi;
  // a column whose weight is < 1e-28 of the total is numerically zero (rank-deficient theta is the
  // common case); rotating it again only chases rounding noise
  const bool live = (a > floor2) && (b > floor2);
  if (!(live && g2 > tol2 * a * b && g2 > abs2)) return 0;
  const double tau = b - a;
  const double rh = qk_rsqrt(fma(tau, tau, 4.0 * g2));
  const double x = fma(0.5 * fabs(tau), rh, 0.5);
  const double rs = qk_rsqrt(x);
  cs = x * rs;
  const double k = (tau >= 0.0 ? rh : -rh) * rs;
  f = cmake(k * gr, k * gi);
  return (g2 > QK_QUAD_EPS2 * a * b) ? 2 : 1;
}

QK_DEV void qk_rot2(c128& xp, c128& xq, double cs, c128 f) {
  const c128 p = xp, q = xq;
  xp.x = fma(-f.y, q.y, fma(-f.x, q.x, cs * p.x));
  xp.y = fma(f.y, q.x, fma(-f.x, q.y, cs * p.y));
  xq.x = fma(-f.y, p.y, fma(f.x, p.x, cs * q.x));
  xq.y = fma(f.y, p.x, fma(f.x, p.y, cs * q.y));
}

QK_DEV void qk_rotate_rows(c128* up, c128* uq, int rows, int sl, int tpp, double cs, c128 f) {
  for (int row = sl; row < rows; row += tpp) {
    c128 xp = up[row], xq = uq[row];
    qk_rot2(xp, xq, cs, f);
    up[row] = xp;
    uq[row] = xq;
  }
}

#if defined(__CUDA_ARCH__) && !defined(QK_HOST_EMU)
// Device fast path of one (pair, round) step: the thread's <= RPT rows of both columns stay in
// registers between the dot products and the rotation; partial sums reduced with warp shuffles.
template <int RPT>
__device__ __forceinline__ void qk_pair_step(c128* __restrict__ wp, c128* __restrict__ wq, int R, int sl, int tpp,
                                             bool valid, double tol2, double floor2, double abs2, int* rotated) {
  c128 xp[RPT], xq[RPT];
  double a = 0, b = 0, gr = 0, gi = 0;
#pragma unroll
  for (int k = 0; k < RPT; ++k) {
    const int row = sl + k * tpp;
    if (valid && row < R) { xp[k] = wp[row]; xq[k] = wq[row]; }
    else { xp[k] = cmake(0, 0); xq[k] = cmake(0, 0); }
    a = fma(xp[k].x, xp[k].x, fma(xp[k].y, xp[k].y, a));
    b = fma(xq[k].x, xq[k].x, fma(xq[k].y, xq[k].y, b));
    gr = fma(xp[k].x, xq[k].x, fma(xp[k].y, xq[k].y, gr));    // gamma = conj(xp) * xq
    gi = fma(xp[k].x, xq[k].y, fma(-xp[k].y, xq[k].x, gi));
  }
  for (int off = tpp >> 1; off > 0; off >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, off);
    b += __shfl_xor_sync(0xffffffffu, b, off);
    gr += __shfl_xor_sync(0xffffffffu, gr, off);
    gi += __shfl_xor_sync(0xffffffffu, gi, off);
  }
  double cs; c128 f;
  const int rot = valid ? qk_rotation(a, b, gr, gi, tol2, floor2, abs2, cs, f) : 0;
  if (rot) {
#pragma unroll
    for (int k = 0; k < RPT; ++k) {
      const int row = sl + k * tpp;
      if (row < R) {
        qk_rot2(xp[k], xq[k], cs, f);
        wp[row] = xp[k];
        wq[row] = xq[k];
      }
    }
    if (sl == 0 && rot == 2) *rotated = 1;
  }
}

// Same step for any number of rows per thread (rows are re-read for the rotation).
__device__ __forceinline__ void qk_pair_step_generic(c128* wp, c128* wq, int R, int sl, int tpp, bool valid,
                                                     double tol2, double floor2, double abs2, int* rotated) {
  double a = 0, b = 0, gr = 0, gi = 0;
  if (valid) {
    for (int row = sl; row < R; row += tpp) {
      const c128 xp = wp[row], xq = wq[row];
      a = fma(xp.x, xp.x, fma(xp.y, xp.y, a));
      b = fma(xq.x, xq.x, fma(xq.y, xq.y, b));
      gr = fma(xp.x, xq.x, fma(xp.y, xq.y, gr));
      gi = fma(xp.x, xq.y, fma(-xp.y, xq.x, gi));
    }
  }
  for (int off = tpp >> 1; off > 0; off >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, off);
    b += __shfl_xor_sync(0xffffffffu, b, off);
    gr += __shfl_xor_sync(0xffffffffu, gr, off);
    gi += __shfl_xor_sync(0xffffffffu, gi, off);
  }
  double cs; c128 f;
  const int rot = valid ? qk_rotation(a, b, gr, gi, tol2, floor2, abs2, cs, f) : 0;
  if (rot) {
    qk_rotate_rows(wp, wq, R, sl, tpp, cs, f);
    if (sl == 0 && rot == 2) *rotated = 1;
  }
}
#endif

// One sweep (every round of the round-robin tournament) of one-sided Jacobi on the R x C matrix W (column-major,
// leading dimension ldw).  Sets c.sh->rotated when a rotation above QK_QUAD_EPS was applied.  Shared by the
// shared-memory-resident path (qk_jacobi) and the column-block visits of the large-matrix path (qk_sim_big.h).
// WIDE: the caller has the register budget for 16 rows of both columns per thread (large-matrix kernel, one CTA per SM).
template <int G, bool WIDE = false>
QK_DEV void qk_jacobi_sweep(SimCtx& c, c128* W, int ldw, int R, int C, double tol2, double floor2, double abs2) {
  const int Ce = (C + 1) & ~1;
  const int npairs = Ce / 2;
  const int nrounds = Ce - 1;
  int tpp = 1;   // threads per column pair: a power of two <= 32 so that a pair's threads share a warp
  while (tpp * 2 * npairs <= G && tpp * 2 <= R && tpp < 32) tpp *= 2;
  const int pp = G / tpp;
  const int npass = (npairs + pp - 1) / pp;
  const int rpt = (R + tpp - 1) / tpp;   // rows per thread
  (void)rpt;
  for (int r = 0; r < nrounds; ++r) {
    for (int pass = 0; pass < npass; ++pass) {
#if defined(__CUDA_ARCH__) && !defined(QK_HOST_EMU)
      // device: rows kept in registers, partial dot products reduced with warp shuffles -> one barrier per round
      QK_PAR_BEGIN(tid)
        const int ps = tid / tpp, sl = tid - ps * tpp;
        const int i = pass * pp + ps;
        int p = 0, q = 0;
        const bool valid = (i < npairs) && qk_rr_pair(i, r, Ce, C, p, q);
        c128* wp = W + (size_t)p * ldw;
        c128* wq = W + (size_t)q * ldw;
        if (rpt <= 1) qk_pair_step<1>(wp, wq, R, sl, tpp, valid, tol2, floor2, abs2, &c.sh->rotated);
        else if (rpt <= 2) qk_pair_step<2>(wp, wq, R, sl, tpp, valid, tol2, floor2, abs2, &c.sh->rotated);
        else if (rpt <= 4) qk_pair_step<4>(wp, wq, R, sl, tpp, valid, tol2, floor2, abs2, &c.sh->rotated);
        else if (rpt <= 8 && G != 128) qk_pair_step<8>(wp, wq, R, sl, tpp, valid, tol2, floor2, abs2, &c.sh->rotated);
        else if (WIDE && rpt <= 16) qk_pair_step<16>(wp, wq, R, sl, tpp, valid, tol2, floor2, abs2, &c.sh->rotated);
        else qk_pair_step_generic(wp, wq, R, sl, tpp, valid, tol2, floor2, abs2, &c.sh->rotated);
      QK_PAR_END
#else
      QK_PAR_BEGIN(tid)
        const int ps = tid / tpp, sl = tid - ps * tpp;
        const int i = pass * pp + ps;
        int p = 0, q = 0;
        const bool valid = (i < npairs) && qk_rr_pair(i, r, Ce, C, p, q);
        double a = 0, b = 0, gr = 0, gi = 0;
        if (valid) {
          const c128* wp = W + (size_t)p * ldw;
          const c128* wq = W + (size_t)q * ldw;
          for (int row = sl; row < R; row += tpp) {
            const c128 xp = wp[row], xq = wq[row];
            a += xp.x * xp.x + xp.y * xp.y;
            b += xq.x * xq.x + xq.y * xq.y;
            gr += xp.x * xq.x + xp.y * xq.y;   // gamma = conj(xp) * xq
            gi += xp.x * xq.y - xp.y * xq.x;
          }
        }
        c.scr[4 * tid + 0] = a; c.scr[4 * tid + 1] = b; c.scr[4 * tid + 2] = gr; c.scr[4 * tid + 3] = gi;
      QK_PAR_END
      QK_PAR_BEGIN(tid)
        const int ps = tid / tpp, sl = tid - ps * tpp;
        const int i = pass * pp + ps;
        int p = 0, q = 0;
        const bool valid = (i < npairs) && qk_rr_pair(i, r, Ce, C, p, q);
        if (valid) {
          double a = 0, b = 0, gr = 0, gi = 0;
          for (int t = 0; t < tpp; ++t) {
            const double* s4 = c.scr + 4 * (ps * tpp + t);
            a += s4[0]; b += s4[1]; gr += s4[2]; gi += s4[3];
          }
          double cs; c128 f;
          const int rot = qk_rotation(a, b, gr, gi, tol2, floor2, abs2, cs, f);
          if (rot) {
            qk_rotate_rows(W + (size_t)p * ldw, W + (size_t)q * ldw, R, sl, tpp, cs, f);
            if (sl == 0 && rot == 2) c.sh->rotated = 1;
          }
        }
      QK_PAR_END
#endif
    }
  }
}

template <int G>
QK_DEV void qk_jacobi(SimCtx& c, int R, int C) {
  c128* W = c.W;
  const double tol2 = c.P->tol * c.P->tol;
  int sweep = 0;
  // total weight (Frobenius norm^2) -- invariant under the rotations
  QK_PAR_BEGIN(tid)
    double s = 0.0;
    for (int i = tid; i < R * C; i += G) s += W[i].x * W[i].x + W[i].y * W[i].y;
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
  // inner products below abs_rel * total are rounding-level relative to the dominant columns: rotating
  // them only polishes directions whose weight is far below any truncation threshold
  const double abs2 = c.P->abs_rel * c.P->abs_rel * total * total;
  if (C >= 2) {
    for (; sweep < c.P->max_sweeps; ++sweep) {
      qk_jacobi_sweep<G>(c, W, R, R, C, tol2, floor2, abs2);
      const int rot = c.sh->rotated;
      QK_BARRIER();
      if (!rot) { ++sweep; break; }
      QK_PAR_BEGIN(tid)
        if (tid == 0) c.sh->rotated = 0;
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
// truncation decision on the sorted squared column norms (thread 0).  SURVEY.md A.4:
//   ITensors  NDTensors truncate!: drop from the tail while discarded + p <= cutoff * sum(p)
//   pytket    drop sigma < value_of_zero, keep the shortest head with numer/denom >= fidelity, renormalise
// ------------------------------------------------------------------------------------------------
QK_DEV void qk_truncate(SimCtx& c, int C, int capb, int bond) {
  const SimParams* P = c.P;
  // a cap equal to the chain-edge rank bound 2^min(b, n-b) is structural, not a limit of the kernel: anything the
  // rule wants to keep beyond it is rounding noise (e.g. cutoff = 0), so QK_FLAG_CAP_HIT is not raised for it
  const int edge_e = bond < P->n - bond ? bond : P->n - bond;
  const bool cap_is_edge = edge_e < 30 && capb >= (1 << edge_e);
  int k;
  double renorm = 1.0;
  if (P->mode == 0) {
    double total = 0.0;
    for (int t = 0; t < C; ++t) total += c.nrm2[c.order[t]];
    double err = 0.0;
    k = C;
    if (!(c.nrm2[c.order[0]] > 0.0)) { k = 1; }
    else {
      const double scale = (total == 0.0) ? 1.0 : total;
      while (k > 1 && err + c.nrm2[c.order[k - 1]] <= P->cutoff * scale) { err += c.nrm2[c.order[k - 1]]; --k; }
      if (k > capb) {
        if (!cap_is_edge) c.sh->flags |= QK_FLAG_CAP_HIT;
        while (k > capb) { err += c.nrm2[c.order[k - 1]]; --k; }
      }
      c.sh->trunc_weight += err / scale;
      c.sh->fidelity *= (1.0 - err / scale);
    }
  } else {
    int m = 0;
    const double zero2 = P->value_of_zero * P->value_of_zero;
    for (int t = 0; t < C; ++t) if (c.nrm2[c.order[t]] >= zero2) ++m;
    if (m < 1) m = 1;
    double denom = 0.0;
    for (int t = 0; t < m; ++t) denom += c.nrm2[c.order[t]];
    if (!(denom > 0.0)) { k = 1; }
    else {
      double numer = 0.0;
      k = 0;
      if (P->fidelity_target < 1.0) {
        while (P->fidelity_target > numer / denom && k < m) { numer += c.nrm2[c.order[k]]; ++k; }
        if (k < 1) { numer = c.nrm2[c.order[0]]; k = 1; }
      } else { k = m; numer = denom; }
      if (k > capb) {
        if (!cap_is_edge) c.sh->flags |= QK_FLAG_CAP_HIT;
        k = capb;
        numer = 0.0;
        for (int t = 0; t < k; ++t) numer += c.nrm2[c.order[t]];
      }
      const double kept = numer / denom;
      renorm = sqrt(1.0 / kept);
      c.sh->fidelity *= kept;
      c.sh->trunc_weight += (1.0 - kept);
    }
  }
  c.sh->keep = k;
  c.sh->renorm = renorm;
  if (k > c.sh->max_chi) c.sh->max_chi = k;
}

// ------------------------------------------------------------------------------------------------
// 2-qubit gate on sites (k, k+1)
// ------------------------------------------------------------------------------------------------
template <int G>
QK_DEV void qk_op_2q(SimCtx& c, const QkOp& op) {
  if (op.pad & QK_OPF_CONT) {   // not the last gate of a fused group: only accumulate its matrix
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
  c128* A = qk_site(c, k);       // [a][l][b]   (global, L1/L2 resident; both sites stay intact until the
  c128* B = qk_site(c, k + 1);   // [b][r][c]    new tensors are complete)
  c128* W = c.W;

  QK_PAR_BEGIN(tid)
    if (tid == 0) qk_build_gate_2q_fused(op, c.x, c.gate, c.gacc, (c128*)c.scr);
  QK_PAR_END

  // theta[(a,L),(R,c)] = sum_{l,r} g[(L,R),(l,r)] sum_b A[a,l,b] B[b,r,c]
  QK_PAR_BEGIN(tid)
    for (int idx = tid; idx < ca * cc; idx += G) {
      const int a = idx / cc, cidx = idx - a * cc;
      c128 t[4];
      for (int l = 0; l < 2; ++l)
        for (int r = 0; r < 2; ++r) {
          c128 acc = cmake(0, 0);
          const c128* ap = A + (size_t)(a * 2 + l) * cb;
          const c128* bp = B + (size_t)r * cc + cidx;
          for (int b = 0; b < cb; ++b) cfma(acc, ap[b], bp[(size_t)b * 2 * cc]);
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
  QK_PAR_END

  qk_jacobi<G>(c, R, C);

  QK_PAR_BEGIN(tid)
    for (int j = tid; j < C; j += G) {
      double s = 0.0;
      const c128* w = W + (size_t)j * ldw;
      for (int row = 0; row < R; ++row) s += w[row].x * w[row].x + w[row].y * w[row].y;
      c.nrm2[j] = s;
    }
  QK_PAR_END
  QK_PAR_BEGIN(tid)
    const double dead = c.P->floor_rel * c.sh->total;
    int* pos = (int*)c.scr;     // int scratch for the swap list below
    int* at = pos + C;
    for (int j = tid; j < C; j += G) {
      const double v = c.nrm2[j];
      int rk = 0;
      for (int i = 0; i < C; ++i) {
        const double u = c.nrm2[i];
        rk += (u > v) || (u == v && i < j);
      }
      c.order[rk] = j;
      // singular value and its inverse, indexed by sorted position.  A column below the Jacobi floor was never
      // orthogonalised against the others, so the projection below is meaningless for it: its inverse is
      // set to 0, which drops the (<= 1e-14 relative) amplitude it carries if the truncation rule keeps it.
      const double sg = sqrt(v);
      c.diag[rk] = cmake(sg, (sg > 0.0 && (v > dead || rk == 0)) ? 1.0 / sg : 0.0);
      pos[j] = j; at[j] = j;
    }
  QK_PAR_END
  QK_PAR_BEGIN(tid)
    if (tid == 0) {
      qk_truncate(c, C, c.P->cap[k + 1], k + 1);
      // The new tensor is staged behind W.  If W (R x C) plus the staging area (keep x C) exceed the
      // region, the kept columns are first compacted to the front of W by column swaps (sorted order);
      // otherwise they are addressed through order[].  pos[j] = where original column j is now,
      // at[p] = which original column sits at p.
      if ((size_t)R * C + (size_t)c.sh->keep * C > (size_t)c.P->rmax * c.P->rmax) {
        int* pos = (int*)c.scr;
        int* at = pos + C;
        for (int t = 0; t < c.sh->keep; ++t) {
          const int j = c.order[t], p = pos[j];
          c.swp[t] = p;
          if (p != t) { const int jj = at[t]; at[p] = jj; pos[jj] = p; at[t] = j; pos[j] = t; }
        }
      }
    }
  QK_PAR_END

  const int keep = c.sh->keep;
  const double renorm = c.sh->renorm;
  const bool right = (op.dir == QK_DIR_RIGHT);
  const bool compact = (size_t)R * C + (size_t)keep * C > (size_t)c.P->rmax * c.P->rmax;
  if (compact) {
    QK_PAR_BEGIN(tid)
      for (int row = tid; row < R; row += G)
        for (int t = 0; t < keep; ++t) {
          const int p = c.swp[t];
          if (p != t) {
            const c128 x = W[row + (size_t)t * ldw];
            W[row + (size_t)t * ldw] = W[row + (size_t)p * ldw];
            W[row + (size_t)p * ldw] = x;
          }
        }
    QK_PAR_END
  }

  // W[:, col(t)] = sigma_t u_t (t < keep), col(t) = t after compaction, order[t] otherwise.
  // The other factor follows from W_out^dag W_in = Sigma^2 V^dag, with
  // W_in = theta (or theta^dag) contracted in two halves so that theta is never rebuilt:
  //   plain      : new right site  B'[t,R,c] = s_t sum_{r,b} F[t,R,r,b] B[b,r,c],
  //                F[t,R,r,b] = sum_{L,l} g[(L,R),(l,r)] sum_a conj(W[(a,L),t]) A[a,l,b]
  //   transposed : new left site   A'[a,L,t] = s_t sum_{l,b} A[a,l,b] F[t,L,l,b],
  //                F[t,L,l,b] = sum_{R,r} g[(L,R),(l,r)] sum_c B[b,r,c] W[(R,c),t]
  // F is produced in tiles of TT values of t (one (t, b) unit per thread) in c.ef; the new tensor is staged in
  // the tail of W (keep * (R + C) <= rmax^2) because both old sites are still being read.
  c128* S = W + (size_t)R * (compact ? keep : C);
  int TT = G / cb;
  if (TT < 1) TT = 1;
  const int upt = (TT * cb + G - 1) / G;   // units per thread (1 unless cb > G)
  (void)upt;
  for (int t0 = 0; t0 < keep; t0 += TT) {
    QK_PAR_BEGIN(tid)
      for (int u = tid; u < TT * cb; u += G) {
        const int tt = u / cb, b = u - tt * cb;
        const int t = t0 + tt;
        if (t < keep) {
          const c128* w = W + (size_t)(compact ? t : c.order[t]) * ldw;
          c128 e[4];
          e[0] = e[1] = e[2] = e[3] = cmake(0, 0);
          if (!transposed) {
            for (int a = 0; a < ca; ++a) {
              const c128 w0 = w[a * 2], w1 = w[a * 2 + 1];
              const c128 a0 = A[(size_t)(a * 2) * cb + b], a1 = A[(size_t)(a * 2 + 1) * cb + b];
              cfmac(e[0], w0, a0); cfmac(e[1], w0, a1);     // e[L*2+l] += conj(W[(a,L),t]) A[a,l,b]
              cfmac(e[2], w1, a0); cfmac(e[3], w1, a1);
            }
          } else {
            const c128* b0 = B + (size_t)(b * 2) * cc;
            const c128* b1 = b0 + cc;
            for (int cidx = 0; cidx < cc; ++cidx) {
              const c128 w0 = w[cidx], w1 = w[cc + cidx];
              const c128 v0 = b0[cidx], v1 = b1[cidx];
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
              cfma(acc, f[(size_t)b * 4], B[(size_t)(b * 2) * cc + cidx]);
              cfma(acc, f[(size_t)b * 4 + 1], B[(size_t)(b * 2 + 1) * cc + cidx]);
            }
            const double sg = c.diag[t].x, isg = c.diag[t].y;
            (void)sg;
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
              cfma(acc, A[(size_t)(a * 2) * cb + b], f[(size_t)b * 4]);
              cfma(acc, A[(size_t)(a * 2 + 1) * cb + b], f[(size_t)b * 4 + 1]);
            }
            const double isg = c.diag[t].y;
            S[(size_t)row * keep + t] = cscale(acc, right ? isg * isg : renorm * isg);
          }
        }
      }
    QK_PAR_END
  }

  QK_PAR_BEGIN(tid)
    // left site  [a][L][t]  (m x keep),  right site [t][R][c]  (keep x n2)
    for (int idx = tid; idx < m * keep; idx += G) {
      const int row = idx / keep, t = idx - row * keep;
      const double sg = c.diag[t].x, isg = c.diag[t].y;
      c128 v;
      if (!transposed) v = cscale(W[row + (size_t)(compact ? t : c.order[t]) * ldw], right ? isg : renorm);   // W = U S
      else v = S[idx];
      (void)sg;
      A[idx] = v;
    }
    for (int idx = tid; idx < keep * n2; idx += G) {
      const int t = idx / n2, col = idx - t * n2;
      const double isg = c.diag[t].y;
      c128 v;
      if (!transposed) v = S[idx];
      else v = cscale(cconj(W[col + (size_t)(compact ? t : c.order[t]) * ldw]), right ? renorm : isg);   // S V^dag = W^dag
      B[idx] = v;
    }
    if (tid == 0) c.chi[k + 1] = keep;
  QK_PAR_END
}

// ------------------------------------------------------------------------------------------------
// B form (QK_PLAN_PARALLEL).  Every site tensor is kept right-orthonormal (B_k = Gamma_k Lambda_k) and the
// Schmidt values Lambda_b of every bond are stored next to the state, so that
//     theta = Lambda_{k-1} * gate * (B_k B_{k+1})
// is the orthogonality-centre tensor of bond k without any gauge move, and updates of bonds that share no
// site are independent (Vidal's TEBD form; the update below is Hastings' division-free variant):
//     theta = U S V^dag ,   B'_{k+1} = V^dag ,   B'_k = (gate * (B_k B_{k+1})) V ,   Lambda'_k = S (renormalised)
// Same truncation rules on the same Schmidt spectrum as the sequential form; results differ at the level
// of the truncated weights (1e-16).
// ------------------------------------------------------------------------------------------------
// out[(a,L)][t] = s_t sum_{l,b} A[a,l,b] F[t,L,l,b],  F[t,L,l,b] = sum_{R,r} g[(L,R),(l,r)] sum_c B[b,r,c] X[(R,c),t]
// with X[(R,c),t] = src[(R*cc + c) + col(t) * sT] (conjugated if conj_src); s_t = 1 or 1/sigma_t.
template <int G>
QK_DEV void qk_half_left(SimCtx& c, const c128* A, const c128* B, int ca, int cb, int cc, int keep, const c128* src,
                         int sT, bool use_order, bool conj_src, bool scale_isg, c128* out) {
  const int m = 2 * ca;
  int TT = G / cb;
  if (TT < 1) TT = 1;
  for (int t0 = 0; t0 < keep; t0 += TT) {
    QK_PAR_BEGIN(tid)
      for (int u = tid; u < TT * cb; u += G) {
        const int tt = u / cb, b = u - tt * cb;
        const int t = t0 + tt;
        if (t < keep) {
          const c128* w = src + (size_t)(use_order ? c.order[t] : t) * sT;
          c128 e[4];
          e[0] = e[1] = e[2] = e[3] = cmake(0, 0);
          const c128* b0 = B + (size_t)(b * 2) * cc;
          const c128* b1 = b0 + cc;
          for (int cidx = 0; cidx < cc; ++cidx) {
            c128 w0 = w[cidx], w1 = w[cc + cidx];
            if (conj_src) { w0 = cconj(w0); w1 = cconj(w1); }
            const c128 v0 = b0[cidx], v1 = b1[cidx];
            cfma(e[0], v0, w0); cfma(e[1], v1, w0);       // e[R*2+r] += B[b,r,c] X[(R,c),t]
            cfma(e[2], v0, w1); cfma(e[3], v1, w1);
          }
          c128* f = c.ef + (size_t)u * 4;
          for (int L = 0; L < 2; ++L)
            for (int l = 0; l < 2; ++l) {
              c128 acc = cmake(0, 0);
              for (int Rr = 0; Rr < 2; ++Rr)
                for (int r = 0; r < 2; ++r) cfma(acc, c.gate[(L * 2 + Rr) * 4 + (l * 2 + r)], e[Rr * 2 + r]);
              f[L * 2 + l] = acc;
            }
        }
      }
    QK_PAR_END
    QK_PAR_BEGIN(tid)
      for (int idx = tid; idx < TT * m; idx += G) {
        const int tt = idx / m, row = idx - tt * m;
        const int t = t0 + tt;
        if (t < keep) {
          const int a = row >> 1, L = row & 1;
          const c128* f = c.ef + (size_t)tt * cb * 4 + L * 2;
          c128 acc = cmake(0, 0);
          for (int b = 0; b < cb; ++b) {
            cfma(acc, A[(size_t)(a * 2) * cb + b], f[(size_t)b * 4]);
            cfma(acc, A[(size_t)(a * 2 + 1) * cb + b], f[(size_t)b * 4 + 1]);
          }
          out[(size_t)row * keep + t] = scale_isg ? cscale(acc, c.diag[t].y) : acc;
        }
      }
    QK_PAR_END
  }
}

// out[t][(R,c)] = isg_t^2 sum_{r,b} F[t,R,r,b] B[b,r,c],  F[t,R,r,b] = sum_{L,l} g[(L,R),(l,r)] sum_a conj(W[(a,L),col(t)]) lam[a] A[a,l,b]
template <int G>
QK_DEV void qk_half_right(SimCtx& c, const c128* A, const double* lam, const c128* B, int ca, int cb, int cc, int keep,
                          const c128* W, int ldw, bool use_order, c128* out) {
  const int n2 = 2 * cc;
  int TT = G / cb;
  if (TT < 1) TT = 1;
  for (int t0 = 0; t0 < keep; t0 += TT) {
    QK_PAR_BEGIN(tid)
      for (int u = tid; u < TT * cb; u += G) {
        const int tt = u / cb, b = u - tt * cb;
        const int t = t0 + tt;
        if (t < keep) {
          const c128* w = W + (size_t)(use_order ? c.order[t] : t) * ldw;
          c128 e[4];
          e[0] = e[1] = e[2] = e[3] = cmake(0, 0);
          for (int a = 0; a < ca; ++a) {
            const double la = lam[a];
            const c128 w0 = cscale(w[a * 2], la), w1 = cscale(w[a * 2 + 1], la);
            const c128 a0 = A[(size_t)(a * 2) * cb + b], a1 = A[(size_t)(a * 2 + 1) * cb + b];
            cfmac(e[0], w0, a0); cfmac(e[1], w0, a1);     // e[L*2+l] += conj(W[(a,L),t]) lam_a A[a,l,b]
            cfmac(e[2], w1, a0); cfmac(e[3], w1, a1);
          }
          c128* f = c.ef + (size_t)u * 4;
          for (int Rr = 0; Rr < 2; ++Rr)
            for (int r = 0; r < 2; ++r) {
              c128 acc = cmake(0, 0);
              for (int L = 0; L < 2; ++L)
                for (int l = 0; l < 2; ++l) cfma(acc, c.gate[(L * 2 + Rr) * 4 + (l * 2 + r)], e[L * 2 + l]);
              f[Rr * 2 + r] = acc;
            }
        }
      }
    QK_PAR_END
    QK_PAR_BEGIN(tid)
      for (int idx = tid; idx < TT * n2; idx += G) {
        const int tt = idx / n2, col = idx - tt * n2;
        const int t = t0 + tt;
        if (t < keep) {
          const int Rr = col / cc, cidx = col - Rr * cc;
          const c128* f = c.ef + (size_t)tt * cb * 4 + Rr * 2;
          c128 acc = cmake(0, 0);
          for (int b = 0; b < cb; ++b) {
            cfma(acc, f[(size_t)b * 4], B[(size_t)(b * 2) * cc + cidx]);
            cfma(acc, f[(size_t)b * 4 + 1], B[(size_t)(b * 2 + 1) * cc + cidx]);
          }
          const double isg = c.diag[t].y;
          out[(size_t)t * n2 + col] = cscale(acc, isg * isg);
        }
      }
    QK_PAR_END
  }
}

template <int G>
QK_DEV void qk_op_2q_b(SimCtx& c, const QkOp& op) {
  if (op.pad & QK_OPF_CONT) {   // not the last gate of a fused group: only accumulate its matrix
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
  const bool transposed = (m <= n2);   // ties go to the orientation that needs one half contraction, not two
  const int R = transposed ? n2 : m;
  const int C = transposed ? m : n2;
  const int ldw = R;
  c128* A = qk_site(c, k);
  c128* B = qk_site(c, k + 1);
  c128* W = c.W;
  const int lam_ld = c.P->lam_ld;
  const double* lamL = c.lam + (size_t)k * lam_ld;          // Schmidt values of the bond left of site k
  double* lamM = c.lam + (size_t)(k + 1) * lam_ld;          // bond (k, k+1): rewritten

  QK_PAR_BEGIN(tid)
    if (tid == 0) qk_build_gate_2q_fused(op, c.x, c.gate, c.gacc, (c128*)c.scr);
  QK_PAR_END

  // theta[(a,L),(R,c)] = lam[a] sum_{l,r} g[(L,R),(l,r)] sum_b A[a,l,b] B[b,r,c]
  QK_PAR_BEGIN(tid)
    for (int idx = tid; idx < ca * cc; idx += G) {
      const int a = idx / cc, cidx = idx - a * cc;
      const double la = lamL[a];
      c128 t[4];
      for (int l = 0; l < 2; ++l)
        for (int r = 0; r < 2; ++r) {
          c128 acc = cmake(0, 0);
          const c128* ap = A + (size_t)(a * 2 + l) * cb;
          const c128* bp = B + (size_t)r * cc + cidx;
          for (int b = 0; b < cb; ++b) cfma(acc, ap[b], bp[(size_t)b * 2 * cc]);
          t[l * 2 + r] = cscale(acc, la);
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
  QK_PAR_END

  qk_jacobi<G>(c, R, C);

  QK_PAR_BEGIN(tid)
    for (int j = tid; j < C; j += G) {
      double s = 0.0;
      const c128* w = W + (size_t)j * ldw;
      for (int row = 0; row < R; ++row) s += w[row].x * w[row].x + w[row].y * w[row].y;
      c.nrm2[j] = s;
    }
  QK_PAR_END
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
    if (tid == 0) qk_truncate(c, C, c.P->cap[k + 1], k + 1);
  QK_PAR_END
  const int keep = c.sh->keep;
  const double renorm = c.sh->renorm;

  // The kept columns stay where they are (addressed through order[]); staging goes behind W when it fits,
  // else the kept columns are first copied out compactly.  keep * (R + C) <= rmax^2 always holds.
  const bool fits = (size_t)R * C + (size_t)keep * C <= (size_t)c.P->rmax * c.P->rmax;
  if (!fits) {
    // compact copy of the kept columns through registers would break the phase rules; use the swap list
    QK_PAR_BEGIN(tid)
      if (tid == 0) {
        int* pos = (int*)c.scr;
        int* at = pos + C;
        for (int j = 0; j < C; ++j) { pos[j] = j; at[j] = j; }
        for (int t = 0; t < keep; ++t) {
          const int j = c.order[t], p = pos[j];
          c.swp[t] = p;
          if (p != t) { const int jj = at[t]; at[p] = jj; pos[jj] = p; at[t] = j; pos[j] = t; }
        }
      }
    QK_PAR_END
    QK_PAR_BEGIN(tid)
      for (int row = tid; row < R; row += G)
        for (int t = 0; t < keep; ++t) {
          const int p = c.swp[t];
          if (p != t) {
            const c128 x = W[row + (size_t)t * ldw];
            W[row + (size_t)t * ldw] = W[row + (size_t)p * ldw];
            W[row + (size_t)p * ldw] = x;
          }
        }
    QK_PAR_END
  }
  const bool use_order = fits;
  c128* S = W + (size_t)R * (fits ? C : keep);    // staging area 1: keep x C entries

  if (transposed) {
    // W[(R,c), t] = sigma_t * (t-th right singular vector of theta): B'_{k+1} = conj(W)^T / sigma; B'_k = C W / sigma
    qk_half_left<G>(c, A, B, ca, cb, cc, keep, W, ldw, use_order, false, true, S);
    QK_PAR_BEGIN(tid)
      for (int idx = tid; idx < keep * n2; idx += G) {
        const int t = idx / n2, col = idx - t * n2;
        B[idx] = cscale(cconj(W[col + (size_t)(use_order ? c.order[t] : t) * ldw]), c.diag[t].y);
      }
      for (int idx = tid; idx < m * keep; idx += G) A[idx] = S[idx];
      for (int t = tid; t < keep; t += G) lamM[t] = c.diag[t].x * renorm;
      if (tid == 0) c.chi[k + 1] = keep;
    QK_PAR_END
  } else {
    // W[(a,L), t] = sigma_t u_t: V^dag = S^-2 W^dag theta (staged in S), then B'_k = C V (staged over W's head)
    qk_half_right<G>(c, A, lamL, B, ca, cb, cc, keep, W, ldw, use_order, S);
    c128* S2 = W;   // m x keep <= R x keep entries: the kept columns of W are dead now
    qk_half_left<G>(c, A, B, ca, cb, cc, keep, S, n2, false, true, false, S2);
    QK_PAR_BEGIN(tid)
      for (int idx = tid; idx < keep * n2; idx += G) B[idx] = S[idx];
      for (int idx = tid; idx < m * keep; idx += G) A[idx] = S2[idx];
      for (int t = tid; t < keep; t += G) lamM[t] = c.diag[t].x * renorm;
      if (tid == 0) c.chi[k + 1] = keep;
    QK_PAR_END
  }
}

// ------------------------------------------------------------------------------------------------
// Householder QR of W (R x C, column-major ld = R): reflectors H_j = I - 2 u_j u_j^dag with u_j stored
// in column j (rows j..R-1), R factor in the strict upper triangle + c.diag.  Q (R x kk) -> Qm.
// ------------------------------------------------------------------------------------------------
template <int G>
QK_DEV void qk_apply_reflector(SimCtx& c, const c128* u, int len, c128* M, int ldm, int row0, int col0, int ncols) {
  // M[row0.., col0..col0+ncols) -= 2 u (u^dag M)
  if (ncols <= 0) return;
  int tpc = 1;
  while (tpc * 2 * ncols <= G && tpc * 2 <= len) tpc *= 2;
  const int cpp = G / tpc;
  const int npass = (ncols + cpp - 1) / cpp;
  for (int pass = 0; pass < npass; ++pass) {
    QK_PAR_BEGIN(tid)
      const int cs = tid / tpc, sl = tid - cs * tpc;
      const int col = pass * cpp + cs;
      double dr = 0, di = 0;
      if (col < ncols) {
        const c128* mc = M + (size_t)(col0 + col) * ldm + row0;
        for (int i = sl; i < len; i += tpc) {
          const c128 uu = u[i], mm = mc[i];
          dr += uu.x * mm.x + uu.y * mm.y;   // conj(u) * m
          di += uu.x * mm.y - uu.y * mm.x;
        }
      }
      c.scr[2 * tid] = dr; c.scr[2 * tid + 1] = di;
    QK_PAR_END
    QK_PAR_BEGIN(tid)
      const int cs = tid / tpc, sl = tid - cs * tpc;
      const int col = pass * cpp + cs;
      if (col < ncols) {
        double dr = 0, di = 0;
        for (int t = 0; t < tpc; ++t) { dr += c.scr[2 * (cs * tpc + t)]; di += c.scr[2 * (cs * tpc + t) + 1]; }
        const c128 d2 = cmake(2.0 * dr, 2.0 * di);
        c128* mc = M + (size_t)(col0 + col) * ldm + row0;
        for (int i = sl; i < len; i += tpc) mc[i] = csub(mc[i], cmul(u[i], d2));
      }
    QK_PAR_END
  }
}

template <int G>
QK_DEV int qk_householder_qr(SimCtx& c, int R, int C, c128* Qm) {
  c128* W = c.W;
  const int ldw = R;
  const int kk = R < C ? R : C;
  for (int j = 0; j < kk; ++j) {
    const int len = R - j;
    c128* col = W + (size_t)j * ldw + j;
    QK_PAR_BEGIN(tid)
      double s = 0.0;
      for (int i = tid; i < len; i += G) s += col[i].x * col[i].x + col[i].y * col[i].y;
      c.scr[tid] = s;
    QK_PAR_END
    QK_PAR_BEGIN(tid)
      if (tid == 0) {
        double n2 = 0.0;
        for (int t = 0; t < G; ++t) n2 += c.scr[t];
        const c128 alpha = col[0];
        const double aabs = sqrt(alpha.x * alpha.x + alpha.y * alpha.y);
        const double nrm = sqrt(n2);
        if (!(nrm > 0.0) || len == 1) {
          // nothing to annihilate: H = I, the diagonal entry stays as it is
          c.sh->hh_skip = 1;
          c.diag[j] = alpha;
          c.order[j] = 1;
        } else {
          const c128 ph = aabs > 0.0 ? cmake(alpha.x / aabs, alpha.y / aabs) : cmake(1.0, 0.0);
          c.diag[j] = cmake(-ph.x * nrm, -ph.y * nrm);
          c.sh->z0 = cmake(alpha.x + ph.x * nrm, alpha.y + ph.y * nrm);
          c.sh->f0 = 1.0 / sqrt(2.0 * nrm * (nrm + aabs));
          c.sh->hh_skip = 0;
          c.order[j] = 0;
        }
      }
    QK_PAR_END
    if (c.sh->hh_skip) continue;
    QK_PAR_BEGIN(tid)
      const double f0 = c.sh->f0;
      for (int i = tid; i < len; i += G) {
        const c128 v = (i == 0) ? c.sh->z0 : col[i];
        col[i] = cscale(v, f0);
      }
    QK_PAR_END
    qk_apply_reflector<G>(c, col, len, W, ldw, j, j + 1, C - 1 - j);
  }
  // Q = H_0 H_1 ... H_{kk-1} applied to the first kk columns of the identity
  QK_PAR_BEGIN(tid)
    for (int i = tid; i < R * kk; i += G) {
      const int cidx = i / R, row = i - cidx * R;
      Qm[i] = cmake(row == cidx ? 1.0 : 0.0, 0.0);
    }
  QK_PAR_END
  for (int j = kk - 1; j >= 0; --j) {
    if (c.order[j]) continue;
    qk_apply_reflector<G>(c, W + (size_t)j * ldw + j, R - j, Qm, R, j, j, kk - j);
  }
  return kk;
}

QK_DEV c128 qk_rfactor(const SimCtx& c, int ldw, int t, int b) {
  if (t < b) return c.W[t + (size_t)b * ldw];
  if (t == b) return c.diag[t];
  return cmake(0, 0);
}

// gauge move: QK_OP_MOVE_R factorises site s as Q R and pushes R into site s+1;
//             QK_OP_MOVE_L factorises site s as L Q and pushes L into site s-1.
// (ITensors orthogonalize!, no truncation -- SURVEY.md A.3)
template <int G>
QK_DEV void qk_op_move(SimCtx& c, const QkOp& op) {
  const int s = op.site;
  const bool to_right = (op.kind == QK_OP_MOVE_R);
  const int cl = c.chi[s], cr = c.chi[s + 1];
  const int R = to_right ? 2 * cl : 2 * cr;
  const int C = to_right ? cr : cl;
  const int ldw = R;
  c128* A = qk_site(c, s);
  c128* W = c.W;
  QK_PAR_BEGIN(tid)
    if (to_right) {
      // W[(a,p), b] = A[a,p,b]
      for (int i = tid; i < R * C; i += G) {
        const int row = i / C, b = i - row * C;
        W[row + (size_t)b * ldw] = A[i];
      }
    } else {
      // W[(p,b), a] = conj(A[a,p,b])
      for (int i = tid; i < R * C; i += G) {
        const int a = i / R, pb = i - a * R;
        W[pb + (size_t)a * ldw] = cconj(A[i]);
      }
    }
  QK_PAR_END
  // W holds R x C <= rmax^2 / 2 entries; Q (R x kk) goes to the other half, which then stages the neighbour
  c128* Qm = c.W + ((size_t)c.P->rmax * c.P->rmax) / 2;
  const int kk = qk_householder_qr<G>(c, R, C, Qm);
  c128* Nb = Qm;
  if (to_right) {
    const int cr2 = c.chi[s + 2];
    c128* Bn = qk_site(c, s + 1);
    QK_PAR_BEGIN(tid)
      for (int i = tid; i < R * kk; i += G) {          // A[(a,p), t] = Q[(a,p), t]
        const int row = i / kk, t = i - row * kk;
        A[i] = Qm[row + (size_t)t * R];
      }
    QK_PAR_END
    QK_PAR_BEGIN(tid)
      for (int i = tid; i < C * 2 * cr2; i += G) Nb[i] = Bn[i];
    QK_PAR_END
    QK_PAR_BEGIN(tid)
      for (int i = tid; i < kk * 2 * cr2; i += G) {    // B'[t, (p,c)] = sum_b R[t,b] B[b,(p,c)]
        const int t = i / (2 * cr2), pc = i - t * 2 * cr2;
        c128 acc = cmake(0, 0);
        for (int b = t; b < C; ++b) cfma(acc, qk_rfactor(c, ldw, t, b), Nb[(size_t)b * 2 * cr2 + pc]);
        Bn[i] = acc;
      }
      if (tid == 0) c.chi[s + 1] = kk;
    QK_PAR_END
  } else {
    const int cl0 = c.chi[s - 1];
    c128* An = qk_site(c, s - 1);
    QK_PAR_BEGIN(tid)
      for (int i = tid; i < kk * R; i += G) {          // A[t,(p,b)] = conj(Q[(p,b), t])
        const int t = i / R, pb = i - t * R;
        A[i] = cconj(Qm[pb + (size_t)t * R]);
      }
    QK_PAR_END
    QK_PAR_BEGIN(tid)
      for (int i = tid; i < cl0 * 2 * C; i += G) Nb[i] = An[i];
    QK_PAR_END
    QK_PAR_BEGIN(tid)
      for (int i = tid; i < cl0 * 2 * kk; i += G) {    // A'[(a',p'), t] = sum_a A[(a',p'), a] conj(R[t,a])
        const int row = i / kk, t = i - row * kk;
        c128 acc = cmake(0, 0);
        for (int a = t; a < C; ++a) cfma(acc, Nb[(size_t)row * C + a], cconj(qk_rfactor(c, ldw, t, a)));
        An[i] = acc;
      }
      if (tid == 0) c.chi[s] = kk;
    QK_PAR_END
  }
}

// ------------------------------------------------------------------------------------------------
// whole circuit for datapoint dp
// ------------------------------------------------------------------------------------------------
template <int G>
QK_DEV void qk_sim_datapoint(SimCtx& c, int dp) {
  const SimParams* P = c.P;
  const int n = P->n;
  c.state = P->store + (size_t)dp * P->state_stride;
  QK_PAR_BEGIN(tid)
    for (int s = tid; s < n; s += G) {   // |0...0>, KernelPkg.jl:68
      c128* A = qk_site(c, s);
      A[0] = cmake(1, 0);
      A[1] = cmake(0, 0);
    }
    for (int b = tid; b <= n; b += G) c.chi[b] = 1;
    for (int i = tid; i < n; i += G) c.x[i] = P->X[(size_t)dp * P->ldx + i];
    if (tid == 0) {
      c.sh->flags = 0; c.sh->sweeps = 0; c.sh->max_chi = 1;
      c.sh->fidelity = 1.0; c.sh->trunc_weight = 0.0; c.sh->rotated = 0;
    }
  QK_PAR_END
  for (int o = 0; o < P->n_ops; ++o) {
    const QkOp op = P->ops[o];
    if (op.kind <= QK_OP_RX) qk_op_1q<G>(c, op);
    else if (op.kind <= QK_OP_SWAP) qk_op_2q<G>(c, op);
    else qk_op_move<G>(c, op);
    if (P->early_exit && op.kind >= QK_OP_XX && op.kind <= QK_OP_SWAP && (c.sh->flags & QK_FLAG_CAP_HIT)) break;
    if (P->trace) {   // memory trace (reference main_track_mem.py logs "MPS size (MiB)=" per gate)
      QK_PAR_BEGIN(tid)
        if (tid == 0) {
          double bytes = 0.0;
          for (int s = 0; s < n; ++s) bytes += 32.0 * c.chi[s] * c.chi[s + 1];
          P->trace[(size_t)dp * P->n_ops + o] = bytes;
        }
      QK_PAR_END
    }
  }
  QK_PAR_BEGIN(tid)
    for (int b = tid; b <= n; b += G) P->chi[(size_t)dp * (n + 1) + b] = c.chi[b];
    if (tid == 0) {
      QkStat st;
      st.fidelity = c.sh->fidelity; st.trunc_weight = c.sh->trunc_weight;
      st.flags = c.sh->flags; st.sweeps = c.sh->sweeps; st.max_chi = c.sh->max_chi; st.pad = 0;
      P->stats[dp] = st;
    }
  QK_PAR_END
}

// ------------------------------------------------------------------------------------------------
// whole circuit for datapoint dp in B form.  `ncta` cooperating groups (the CTAs of a thread-block cluster on
// the device, 1 in the host emulation) share the datapoint: the items of one level are dealt round-robin,
// QK_GROUP_SYNC() separates levels.  chi and the bond weights live in global memory so that every group
// sees the others' updates after the barrier.
// ------------------------------------------------------------------------------------------------
#ifndef QK_GROUP_SYNC
#define QK_GROUP_SYNC() ((void)0)
#endif
template <int G>
QK_DEV void qk_sim_datapoint_b(SimCtx& c, int dp, int cta, int ncta, QkStat* part /* [ncta] partial stats of dp */) {
  const SimParams* P = c.P;
  const int n = P->n;
  c.state = P->store + (size_t)dp * P->state_stride;
  c.lam = P->lam + (size_t)dp * (n + 1) * P->lam_ld;
  c.chi = P->chi + (size_t)dp * (n + 1);
  QK_PAR_BEGIN(tid)
    for (int s = cta * G + tid; s < n; s += G * ncta) {   // |0...0>
      c128* A = qk_site(c, s);
      A[0] = cmake(1, 0);
      A[1] = cmake(0, 0);
    }
    for (int b = cta * G + tid; b <= n; b += G * ncta) { c.chi[b] = 1; c.lam[(size_t)b * P->lam_ld] = 1.0; }
    for (int i = tid; i < n; i += G) c.x[i] = P->X[(size_t)dp * P->ldx + i];
    if (tid == 0) {
      c.sh->flags = 0; c.sh->sweeps = 0; c.sh->max_chi = 1;
      c.sh->fidelity = 1.0; c.sh->trunc_weight = 0.0; c.sh->rotated = 0;
    }
  QK_PAR_END
  QK_GROUP_SYNC();
  for (int lv = 0; lv < P->n_levels; ++lv) {
    int item = -1;
    for (int o = P->level_start[lv]; o < P->level_start[lv + 1]; ++o) {
      const QkOp op = P->ops[o];
      if (!(op.pad & QK_OPF_ACC)) ++item;           // a fused group (CONT ... last) is one item
      if (item % ncta != cta) continue;
      if (op.kind <= QK_OP_RX) qk_op_1q<G>(c, op);
      else qk_op_2q_b<G>(c, op);
    }
    QK_GROUP_SYNC();
  }
  QK_PAR_BEGIN(tid)
    if (tid == 0) {
      QkStat st;
      st.fidelity = c.sh->fidelity; st.trunc_weight = c.sh->trunc_weight;
      st.flags = c.sh->flags; st.sweeps = c.sh->sweeps; st.max_chi = c.sh->max_chi; st.pad = 0;
      part[cta] = st;
    }
  QK_PAR_END
  QK_GROUP_SYNC();
  QK_PAR_BEGIN(tid)
    if (cta == 0 && tid == 0) {
      QkStat st = part[0];
      for (int i = 1; i < ncta; ++i) {
        st.fidelity *= part[i].fidelity; st.trunc_weight += part[i].trunc_weight;
        st.flags |= part[i].flags; st.sweeps += part[i].sweeps;
        if (part[i].max_chi > st.max_chi) st.max_chi = part[i].max_chi;
      }
      P->stats[dp] = st;
    }
  QK_PAR_END
}
