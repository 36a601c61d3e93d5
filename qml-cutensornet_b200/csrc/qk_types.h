// Internal types shared by the plan compiler, the simulation core and the kernels.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define QK_HD __host__ __device__ __forceinline__
#else
#define QK_HD inline
#endif

// complex128 as two doubles; 16-byte aligned so global/shared accesses are 128-bit.
struct alignas(16) c128 {
  double x, y;
};

QK_HD c128 cmake(double x, double y) { c128 r; r.x = x; r.y = y; return r; }
QK_HD c128 cadd(c128 a, c128 b) { return cmake(a.x + b.x, a.y + b.y); }
QK_HD c128 csub(c128 a, c128 b) { return cmake(a.x - b.x, a.y - b.y); }
QK_HD c128 cmul(c128 a, c128 b) { return cmake(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
QK_HD c128 cmulc(c128 a, c128 b) { return cmake(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y); }  // a * conj(b)
QK_HD c128 cconj(c128 a) { return cmake(a.x, -a.y); }
QK_HD c128 cscale(c128 a, double s) { return cmake(a.x * s, a.y * s); }
QK_HD double cabs2(c128 a) { return a.x * a.x + a.y * a.y; }
QK_HD void cfma(c128& acc, c128 a, c128 b) {  // acc += a*b
  acc.x += a.x * b.x - a.y * b.y;
  acc.y += a.x * b.y + a.y * b.x;
}
QK_HD void cfmac(c128& acc, c128 a, c128 b) {  // acc += conj(a)*b
  acc.x += a.x * b.x + a.y * b.y;
  acc.y += a.x * b.y - a.y * b.x;
}

// op kinds of the compiled schedule: gate kinds 0..5 as in qkmps.h, plus gauge moves
enum {
  QK_OP_H = 0,
  QK_OP_RZ = 1,
  QK_OP_RX = 2,
  QK_OP_XX = 3,
  QK_OP_ZZ = 4,
  QK_OP_SWAP = 5,
  QK_OP_ID2 = 6,      // identity on (site, site+1): the large-matrix path expresses a gauge move as an SVD of the
                      // site pair (no truncation, QK_OPF_NOTRUNC) instead of a Householder QR
  QK_OP_MOVE_R = 16,  // QR of site `site`, push R into site+1
  QK_OP_MOVE_L = 17   // LQ of site `site`, push L into site-1
};

enum { QK_DIR_RIGHT = 0, QK_DIR_LEFT = 1 };

// 2-qubit ops on the same bond that follow each other are fused into one SVD: every op but the last of
// such a group only multiplies its 4x4 gate into an accumulator (QK_OPF_CONT), the others start from it
// (QK_OPF_ACC).
enum { QK_OPF_CONT = 1, QK_OPF_ACC = 2, QK_OPF_NOTRUNC = 4 };

struct QkOp {
  int32_t kind;
  int32_t site;   // 1-qubit: the site; 2-qubit: left site k of (k, k+1); move: the site factorised
  int32_t fa, fb; // feature indices of the angle expression (fa < 0: constant angle)
  int32_t dir;    // 2-qubit ops: which factor receives the singular values
  int32_t pad;    // QK_OPF_* flags
  double coeff;
};

enum {
  QK_FLAG_CAP_HIT = 1,      // truncation rule wanted more than the bond cap: extra weight was discarded
  QK_FLAG_NO_CONVERGE = 2   // Jacobi hit the sweep limit
};

struct QkStat {
  double fidelity;      // product of kept weight fractions (pytket mps.fidelity)
  double trunc_weight;  // sum of relative discarded weights
  int32_t flags;
  int32_t sweeps;       // total Jacobi sweeps (profiling)
  int32_t max_chi;
  int32_t pad;
};

struct SimParams {
  int n;                     // qubits / sites
  int n_ops;
  const QkOp* ops;
  const int32_t* cap;        // [n+1] bond caps (cap[0] = cap[n] = 1)
  const int64_t* site_off;   // [n+1] offsets (c128 units) of the site slots inside one state
  int64_t state_stride;      // c128 units per state
  const double* X;           // [N][ldx]
  int ldx;
  int N;
  c128* store;               // [N][state_stride]
  int32_t* chi;              // [N][n+1]
  QkStat* stats;             // [N]
  int mode;                  // qk_trunc_mode
  double cutoff;             // ITensors relative cutoff
  double fidelity_target;    // pytket: 1 - truncation_error
  double value_of_zero;      // pytket: absolute sigma cutoff
  double tol;                // Jacobi rotation threshold
  int max_sweeps;
  int wr;                    // c128 elements per shared-memory matrix region = (2*capmax)^2
  int rmax;                  // 2*capmax
  double* trace;             // optional [N][n_ops]: MPS size in bytes after every op (memory trace), or NULL
  int early_exit;            // stop a datapoint at its first bond-cap hit (state then invalid, flag set)
  double floor_rel;          // Jacobi: columns below floor_rel * total weight are treated as numerically zero
  double abs_rel;            // Jacobi: no rotation when |x^dag y| <= abs_rel * total weight
  // B-form (QK_PLAN_PARALLEL): Schmidt values of every bond, op levels
  int parallel;              // 1: ops are levelised; state is kept right-canonical with explicit bond weights
  double* lam;               // [N][(n+1)][lam_ld] Schmidt values per bond (global)
  int lam_ld;                // = max bond cap
  const int32_t* level_start;  // [n_levels+1] first op of every level
  int n_levels;
  // large-matrix path (bond caps above the shared-memory-resident limit, qk_sim_big.h): theta and the staging
  // area live in global memory (L2), one slot per resident cluster
  c128* big_w;               // [n_clusters][big_w_stride]   theta / U Sigma, column-major, rmax^2 entries
  c128* big_s;               // [n_clusters][big_s_stride]   staging of the recovered factor, capmax * rmax entries
  int64_t big_w_stride, big_s_stride;
  int* big_flag;             // [n_clusters][2] "a rotation happened in this sweep", alternating slots
  int big_jb;                // columns per block of the block Jacobi (upper limit; per SVD: what fits, below)
  int big_wb_entries;        // c128 entries of the shared-memory buffer for one pair of column blocks
  long long* unit_clk;       // optional [N]: clock64 ticks each datapoint took (per-unit timing), or NULL
};

#include <stddef.h>
// c128 entries of the shared-memory buffer that holds one pair of column blocks: room for 2 jb columns of 2 cap rows,
// but never more than 128 KB.  The block size of an SVD is chosen from its ACTUAL row count (qk_big_jacobi), so a
// generous bond cap does not shrink the blocks of the matrices that actually occur.
QK_HD size_t qk_big_wb_entries(int rmax, int jb) {
  size_t e = (size_t)rmax * 2 * jb;
  const size_t lim = (size_t)128 * 1024 / sizeof(c128);
  if (e > lim) e = lim;
  if (e < (size_t)rmax * 2) e = (size_t)rmax * 2;                     // at least one column per block
  return e;
}

