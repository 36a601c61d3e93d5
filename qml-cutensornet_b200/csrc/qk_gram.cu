// Stage 2: all-pairs MPS overlaps  K[y][x] = |<psi_y|psi_x>|^2  (replaces x_mps.vdot(y_mps),
// gpu_backend/kernel_state_ansatz.py:380-387, and abs(inner(y,x))^2, KernelPkg/src/KernelPkg.jl:103-109).
//
// qk_gram_dmma_kernel -- the product kernel.  One CTA owns a tile of TI kets x TJ bras; one warp
// per (bra, ket) pair carries the transfer matrix E (bra bond x ket bond, complex, <= 16x16) in
// FP64 tensor-core accumulator fragments for the whole n-site sweep and never spills it to memory:
//     T_p^T = A_x[p]^T  * E^T          (step 1; E   is consumed directly as the MMA B operand)
//     E'   += A_y[p]^dag * T_p          (step 2; T^T is consumed directly as the MMA B operand)
// using mma.sync.m8n8k4.f64 (DMMA).  The accumulator layout of one MMA is exactly the B-operand
// layout of the next when the contraction index is taken in the order (even columns, odd columns),
// so the site tensors are stored in HBM already permuted into A-fragment order ("frag" layout,
// written by qk_pack_kernel) and arrive in shared memory through an NS-stage (4) cp.async.bulk (TMA)
// + mbarrier pipeline; the warp that is last to finish a site refills its stage with site s + NS.
//   * every bond index is stored permuted inside its group of 8 (logical 8t+4e+j <-> physical
//     8t+2j+e) so that the even / odd k-blocks of a tile hold logical indices 8t..8t+3 / 8t+4..8t+7:
//     a state whose bond dimension is <= 8t+4 skips the odd k-block (contraction granularity 4).
//   * complex products use 3 real MMAs (S1 = Ar Br, S2 = Ai Bi, S3 = (Ar +- Ai)(Br + Bi)) instead of 4.
//
// qk_gram_store_kernel -- CUDA-core FP64 cross-check on the unpadded stores (any chi); used by
// tests and as the path for bond dimensions above 16.
#include <stdint.h>
#include <stdlib.h>
#include "qk_kernels.cuh"

#ifndef QK_TI
#define QK_TI 4
#endif
#ifndef QK_TJ
#define QK_TJ 2
#endif
#ifndef QK_NS
#define QK_NS 4
#endif
#define QK_GRAM_WARPS (QK_TI * QK_TJ)


void qk_frag_layout(int n, const int32_t* D, FragLayout* L, int64_t* site_off_bytes) {
  int64_t off = 0;
  for (int s = 0; s < n; ++s) {
    if (site_off_bytes) site_off_bytes[s] = off;
    off += (int64_t)D[s] * D[s + 1] * 32;
  }
  if (site_off_bytes) site_off_bytes[n] = off;
  L->n = n;
  L->data_bytes = off;
  L->stride_bytes = off + (((int64_t)(n + 1) + 15) & ~(int64_t)15);
}

// ------------------------------------------------------------------------------------------------
// pack: unpadded store -> frag layout
//   block(s) doubles index d:  e = d&1, lane = (d>>1)&31, h = (d>>6)&1, kt, mt, p from d>>7
//   value = Re/Im (h) of A[c][p][c'] with the LOGICAL bond indices
//       c  = 8kt + 4e + (lane&3)                      (physical column 8kt + 2(lane&3) + e)
//       c' = 8mt + 4((lane>>2)&1) + (lane>>3)         (physical row    8mt + (lane>>2))
//   (0 outside chi).  Trailer: n+1 bytes ceil(chi_b / 4) = live k-blocks of bond b.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) qk_pack_kernel(int n, const c128* __restrict__ store, int64_t state_stride,
                                                      const int64_t* __restrict__ site_off,
                                                      const int32_t* __restrict__ chi, const int32_t* __restrict__ D,
                                                      const int64_t* __restrict__ frag_off, int64_t frag_stride,
                                                      int64_t frag_data, unsigned char* __restrict__ frag,
                                                      const int32_t* __restrict__ dst_index, int dst_base) {
  const int s = blockIdx.x;
  const int i = blockIdx.y;
  const int idst = dst_index ? dst_index[i] : dst_base + i;   // position of state i in the frag buffer (< 0: skip)
  if (idst < 0) return;
  const int Dl = D[s], Dr = D[s + 1];
  const int KT = Dl >> 3, MT = Dr >> 3;
  const int cl = chi[(size_t)i * (n + 1) + s], cr = chi[(size_t)i * (n + 1) + s + 1];
  const c128* A = store + (size_t)i * state_stride + site_off[s];
  double* out = (double*)(frag + (size_t)idst * frag_stride + frag_off[s]);
  const int total = Dl * Dr * 4;
  for (int d = threadIdx.x; d < total; d += blockDim.x) {
    const int e = d & 1, lane = (d >> 1) & 31, h = (d >> 6) & 1;
    int rest = d >> 7;
    const int kt = rest % KT; rest /= KT;
    const int mt = rest % MT;
    const int p = rest / MT;
    const int cidx = 8 * kt + 4 * e + (lane & 3);
    const int g = lane >> 2;
    const int cp = 8 * mt + 4 * (g & 1) + (g >> 1);
    double v = 0.0;
    if (cidx < cl && cp < cr) {
      const c128 a = A[(size_t)(cidx * 2 + p) * cr + cp];
      v = h ? a.y : a.x;
    }
    out[d] = v;
  }
  if (s == 0) {
    unsigned char* tc = frag + (size_t)idst * frag_stride + frag_data;
    for (int b = threadIdx.x; b <= n; b += blockDim.x) tc[b] = (unsigned char)((chi[(size_t)i * (n + 1) + b] + 3) >> 2);
  }
}

cudaError_t qk_launch_pack(int n, int N, const c128* store, int64_t state_stride, const int64_t* site_off_dev,
                           const int32_t* chi_dev, const int32_t* D_dev, const int64_t* frag_off_dev,
                           int64_t frag_stride_bytes, int64_t frag_data_bytes, void* frag_dev,
                           const int32_t* dst_index_dev, int dst_base, cudaStream_t stream) {
  if (N <= 0) return cudaSuccess;
  dim3 grid(n, N);
  qk_pack_kernel<<<grid, 256, 0, stream>>>(n, store, state_stride, site_off_dev, chi_dev, D_dev, frag_off_dev,
                                           frag_stride_bytes, frag_data_bytes, (unsigned char*)frag_dev, dst_index_dev, dst_base);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// PTX helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t qk_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void qk_mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void qk_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void qk_mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void qk_mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"   // %3: suspend-time hint (ns)
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity), "r"(100000u)
        : "memory");
  } while (!ok);
}
// TMA bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void qk_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
// FP64 tensor-core MMA, D(8x8) += A(8x4) * B(4x8).  Fragments: A: row = lane/4, col = lane%4;
// B: row = lane%4, col = lane/4; C/D: row = lane/4, cols = 2*(lane%4) + {0,1}.
__device__ __forceinline__ void qk_dmma(double (&acc)[2], double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(acc[0]), "+d"(acc[1]) : "d"(a), "d"(b));
}

// same with a zero accumulator input (first MMA of a chain: no register zero-fill needed)
__device__ __forceinline__ void qk_dmma_z(double (&acc)[2], double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%4,%5};"
      : "=d"(acc[0]), "=d"(acc[1]) : "d"(a), "d"(b), "d"(0.0), "d"(0.0));
}

// ------------------------------------------------------------------------------------------------
// One site of the transfer sweep for one (bra, ket) pair, E kept in accumulator fragments.
// The live 8-tiles of the four bonds involved are compile-time:
//   KX / MX = live tiles of the ket's left / right bond, KY / MY = same for the bra;
// the live k-blocks (of 4) are kbx in {2KX-1, 2KX}, kby in {2KY-1, 2KY}: the last k-block of each
// step sits behind one warp-uniform branch, everything else is straight-line code with no predicate
// around any MMA.  (ptxas paces DMMAs with static stall counts and guards every predicated mma.sync
// with WARPSYNC + NOPs, so a predicated-off MMA costs as much as a real one: a version that predicated
// dead tiles away spent the time of the full 16x16 problem on every site.)
// Layout tile counts (MTx, KTx, MTy, KTy = padded D / 8) only enter the operand addresses.
// ------------------------------------------------------------------------------------------------
template <int NT, int KX, int MX, int KY, int MY>
__device__ __forceinline__ void qk_site_step(double (&Er)[NT][NT][2], double (&Ei)[NT][NT][2], double (&Es)[NT][NT][2],
                                             const double2* __restrict__ bx, const double2* __restrict__ by, int lane,
                                             int MTx, int KTx, int MTy, int KTy, bool last_x, bool last_y) {
  // Both physical indices p are carried together: all step-1 MMAs (p = 0, 1), one combine, all step-2
  // MMAs, one combine -- two MMA -> DADD -> MMA dependency bubbles per site instead of four, and long
  // uninterrupted MMA phases that the sibling warp on the scheduler can interleave with.
  double T1[2][MX][KY][2], T2[2][MX][KY][2], T3[2][MX][KY][2];   // T^T[p][ket-right tile][bra-left tile]
  // step 1: T_p^T[c'][a] += sum_c A_x[c,p,c'] * E[a][c]   (S1 = Ar Er, S2 = Ai Ei, S3 = (Ar+Ai)(Er+Ei));
  // k-block outermost: 6*MX*KY independent accumulators between two MMAs on the same one.
  const double2* px = bx + lane;
  const int sx_mt = KTx * 64, sx_p = MTx * sx_mt;
#pragma unroll
  for (int kb = 0; kb < 2 * KX; ++kb) {
    const int kt = kb >> 1, e = kb & 1;
    if (kb < 2 * KX - 1 || last_x) {
#pragma unroll
      for (int p = 0; p < 2; ++p) {
#pragma unroll
        for (int mt = 0; mt < MX; ++mt) {
          const double2* f = px + p * sx_p + mt * sx_mt + kt * 64;
          const double2 fr = f[0], fm = f[32];
          const double ar = e ? fr.y : fr.x;
          const double ai = e ? fm.y : fm.x;
          const double as = ar + ai;
#pragma unroll
          for (int at = 0; at < KY; ++at) {
            if (kb == 0) {   // first MMA of a chain: zero accumulator input, no register zero-fill
              qk_dmma_z(T1[p][mt][at], ar, Er[at][kt][e]);
              qk_dmma_z(T2[p][mt][at], ai, Ei[at][kt][e]);
              qk_dmma_z(T3[p][mt][at], as, Es[at][kt][e]);
            } else {
              qk_dmma(T1[p][mt][at], ar, Er[at][kt][e]);
              qk_dmma(T2[p][mt][at], ai, Ei[at][kt][e]);
              qk_dmma(T3[p][mt][at], as, Es[at][kt][e]);
            }
          }
        }
      }
    }
  }
  // T = (S1 - S2) + i (S3 - S1 - S2);  reuse T1 = Tr, T2 = Ti, T3 = Tr + Ti
#pragma unroll
  for (int p = 0; p < 2; ++p)
#pragma unroll
    for (int a = 0; a < MX; ++a)
#pragma unroll
      for (int b = 0; b < KY; ++b)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const double u1 = T1[p][a][b][e], u2 = T2[p][a][b][e], u3 = T3[p][a][b][e];
          const double tr = u1 - u2, ts = fma(-2.0, u2, u3);
          T1[p][a][b][e] = tr;
          T2[p][a][b][e] = ts - tr;
          T3[p][a][b][e] = ts;
        }
  // step 2: E'[b'][c'] += sum_{a,p} conj(A_y[a,p,b']) * T_p[a][c']:  S1 = Ar Tr, S2 = Ai Ti, S3 = (Ar-Ai)(Tr+Ti)
  double F1[MY][MX][2], F2[MY][MX][2], F3[MY][MX][2];
  const double2* py = by + lane;
  const int sy_mt = KTy * 64, sy_p = MTy * sy_mt;
#pragma unroll
  for (int kb = 0; kb < 2 * KY; ++kb) {
    const int at = kb >> 1, e = kb & 1;
    if (kb < 2 * KY - 1 || last_y) {
#pragma unroll
      for (int p = 0; p < 2; ++p) {
#pragma unroll
        for (int bt = 0; bt < MY; ++bt) {
          const double2* f = py + p * sy_p + bt * sy_mt + at * 64;
          const double2 fr = f[0], fm = f[32];
          const double ar = e ? fr.y : fr.x;
          const double ai = e ? fm.y : fm.x;
          const double ad = ar - ai;
#pragma unroll
          for (int ct = 0; ct < MX; ++ct) {
            if (kb == 0 && p == 0) {
              qk_dmma_z(F1[bt][ct], ar, T1[p][ct][at][e]);
              qk_dmma_z(F2[bt][ct], ai, T2[p][ct][at][e]);
              qk_dmma_z(F3[bt][ct], ad, T3[p][ct][at][e]);
            } else {
              qk_dmma(F1[bt][ct], ar, T1[p][ct][at][e]);
              qk_dmma(F2[bt][ct], ai, T2[p][ct][at][e]);
              qk_dmma(F3[bt][ct], ad, T3[p][ct][at][e]);
            }
          }
        }
      }
    }
  }
  // E' = (S1 + S2) + i (S3 - S1 + S2); tiles outside MY x MX are not read by the next site
#pragma unroll
  for (int a = 0; a < MY; ++a)
#pragma unroll
    for (int b = 0; b < MX; ++b)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const double u1 = F1[a][b][e], u2 = F2[a][b][e], u3 = F3[a][b][e];
        const double er = u1 + u2, es = fma(2.0, u2, u3);
        Er[a][b][e] = er;
        Ei[a][b][e] = es - er;
        Es[a][b][e] = es;
      }
}

// dispatch on the live tile counts (all warp-uniform)
template <int NT>
__device__ __forceinline__ void qk_site_dispatch(double (&Er)[NT][NT][2], double (&Ei)[NT][NT][2], double (&Es)[NT][NT][2],
                                                 const double2* __restrict__ bx, const double2* __restrict__ by, int lane,
                                                 int MTx, int KTx, int MTy, int KTy, int kbx, int mx, int kby, int my) {
  const bool last_x = !(kbx & 1), last_y = !(kby & 1);
  if constexpr (NT == 1) {
    qk_site_step<1, 1, 1, 1, 1>(Er, Ei, Es, bx, by, lane, MTx, KTx, MTy, KTy, last_x, last_y);
  } else {
    const int kx = (kbx + 1) >> 1, ky = (kby + 1) >> 1;
    const int v = (kx - 1) | ((mx - 1) << 1) | ((ky - 1) << 2) | ((my - 1) << 3);
#define QK_SITE_CASE(V) \
    case V: qk_site_step<2, 1 + ((V) & 1), 1 + (((V) >> 1) & 1), 1 + (((V) >> 2) & 1), 1 + (((V) >> 3) & 1)>( \
                Er, Ei, Es, bx, by, lane, MTx, KTx, MTy, KTy, last_x, last_y); break;
    switch (v) {
      QK_SITE_CASE(0) QK_SITE_CASE(1) QK_SITE_CASE(2) QK_SITE_CASE(3) QK_SITE_CASE(4) QK_SITE_CASE(5) QK_SITE_CASE(6) QK_SITE_CASE(7)
      QK_SITE_CASE(8) QK_SITE_CASE(9) QK_SITE_CASE(10) QK_SITE_CASE(11) QK_SITE_CASE(12) QK_SITE_CASE(13) QK_SITE_CASE(14)
      default: qk_site_step<2, 2, 2, 2, 2>(Er, Ei, Es, bx, by, lane, MTx, KTx, MTy, KTy, last_x, last_y); break;
    }
#undef QK_SITE_CASE
  }
}

// ------------------------------------------------------------------------------------------------
// DMMA Gram kernel.  NT = max 8x8 tiles per bond (1: D <= 8, 2: D <= 16).  PPW = pairs per warp (one bra,
// PPW consecutive kets).  CTA tile = QK_TI * PPW kets x QK_TJ bras, 8 warps.
// ------------------------------------------------------------------------------------------------
// (PPW = 4 at NT = 1 was measured: 40 % slower -- the extra registers halve the resident CTAs and the four
// pairs' MMAs serialise in one warp; kept at 1.)
template <int NT> struct GramCfg { static constexpr int PPW = 1; static constexpr int TI = QK_TI * PPW; };

template <int NT>
__global__ void __launch_bounds__(QK_GRAM_WARPS * 32, NT == 1 ? 4 : (QK_GRAM_WARPS <= 4 ? 2 : 1)) qk_gram_dmma_kernel(const __grid_constant__ GramParams P) {
  constexpr int PPW = GramCfg<NT>::PPW;
  constexpr int TI = GramCfg<NT>::TI;
  extern __shared__ __align__(128) unsigned char gsm[];
  const int n = P.n;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int stage_bytes = TI * P.slot_x + QK_TJ * P.slot_y;
  unsigned char* stage0 = gsm;
  uint64_t* bars = (uint64_t*)(gsm + (size_t)QK_NS * stage_bytes);   // full[NS], empty[NS]
  int* sDx = (int*)(bars + 2 * QK_NS);
  int* sDy = sDx + (n + 1);
  unsigned char* stc = (unsigned char*)(sDy + (n + 1));              // [(TI+TJ)][n+1]

  const int4 tile = P.tiles[blockIdx.x];
  const int y0 = tile.x, x0 = tile.y, y_end = tile.z, x_end = tile.w;
  const long long clk0 = clock64();

  for (int b = threadIdx.x; b <= n; b += blockDim.x) { sDx[b] = P.Dx[b]; sDy[b] = P.Dy[b]; }
  for (int t = warp; t < TI + QK_TJ; t += QK_GRAM_WARPS) {
    const bool is_x = t < TI;
    int idx = is_x ? x0 + t : y0 + (t - TI);
    const int lim = is_x ? P.Nx : P.Ny;
    if (idx >= lim) idx = lim - 1;
    const unsigned char* src = is_x ? P.fragX + (size_t)idx * P.strideX + P.dataX : P.fragY + (size_t)idx * P.strideY + P.dataY;
    for (int b = lane; b <= n; b += 32) stc[t * (n + 1) + b] = src[b];
  }
  int* cnt = (int*)(bars + QK_NS);                                   // consumer counters, one per stage
  __shared__ int s_weight[TI];
  if (threadIdx.x == 0) {
    for (int s = 0; s < QK_NS; ++s) {
      qk_mbar_init(qk_smem_u32(&bars[s]), 1);                       // full: the issuing lane's expect_tx arrive
      cnt[s] = 0;
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  // weight of every ket of the tile (sum of squared live k-block counts) for the warp -> pair assignment below
  if (warp < TI) {
    int w = 0;
    for (int b = lane; b <= n; b += 32) { const int k = stc[warp * (n + 1) + b]; w += k * k; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) w += __shfl_xor_sync(0xffffffffu, w, o);
    if (lane == 0) s_weight[warp] = w;
  }
  __syncthreads();

  // The site blocks of the tile's kets and bras are streamed with cp.async.bulk, NS sites deep.  There is no
  // producer warp (3 warps on one scheduler would cap every thread at 168 registers, which the 3M accumulators
  // exceed) and no empty barrier: the warp that is LAST to finish a site (shared-memory counter) refills that
  // stage with site s + NS at once, so a fast warp is never held back by issuing loads for the slow ones.
  auto issue_site = [&](int sl) {
    const int st = sl % QK_NS;
    const uint32_t bxb = (uint32_t)(sDx[sl] * sDx[sl + 1] * 32);
    const uint32_t byb = (uint32_t)(sDy[sl] * sDy[sl + 1] * 32);
    const uint32_t full = qk_smem_u32(&bars[st]);
    qk_mbar_expect_tx(full, TI * bxb + QK_TJ * byb);
    unsigned char* dst = stage0 + (size_t)st * stage_bytes;
    const int64_t ox = P.offx[sl], oy = P.offy[sl];
#pragma unroll
    for (int t = 0; t < TI; ++t) {
      int xi = x0 + t; if (xi >= P.Nx) xi = P.Nx - 1;
      qk_bulk_g2s(qk_smem_u32(dst + (size_t)t * P.slot_x), P.fragX + (size_t)xi * P.strideX + ox, bxb, full);
    }
#pragma unroll
    for (int t = 0; t < QK_TJ; ++t) {
      int yi = y0 + t; if (yi >= P.Ny) yi = P.Ny - 1;
      qk_bulk_g2s(qk_smem_u32(dst + (size_t)TI * P.slot_x + (size_t)t * P.slot_y),
                  P.fragY + (size_t)yi * P.strideY + oy, byb, full);
    }
  };
  if (threadIdx.x == 0) {
    for (int sl = 0; sl < QK_NS && sl < n; ++sl) issue_site(sl);
  }

  // ===== PPW (bra y, ket x) pairs per warp: one bra, PPW consecutive kets =====
  // Warps q and q + 4 share a scheduler.  Its two pairs are (r-th lightest ket, bra 0) and (r-th heaviest ket,
  // bra 1), so the four schedulers of the SM carry about the same number of MMAs.
  const int tj = warp / QK_TI;
  int tk = (warp % QK_TI) * PPW;
  if (PPW == 1) {
    const int want = (tj & 1) ? QK_TI - 1 - (warp % QK_TI) : (warp % QK_TI);
#pragma unroll
    for (int t = 0; t < QK_TI; ++t) {
      int rank = 0;
#pragma unroll
      for (int u = 0; u < QK_TI; ++u) rank += (s_weight[u] < s_weight[t]) || (s_weight[u] == s_weight[t] && u < t);
      if (rank == want) tk = t;
    }
  }
  const int y = y0 + tj;
  bool active[PPW];
  bool any_active = false;
#pragma unroll
  for (int j = 0; j < PPW; ++j) {
    const int x = x0 + tk + j;
    active[j] = (x < x_end) && (y < y_end) && (!P.symmetric || x <= y);
    any_active = any_active || active[j];
  }
  const unsigned char* tcy = stc + (TI + tj) * (n + 1);

  // E[pair][bra tile][ket tile]: real part, imaginary part and their sum (3M complex products)
  double Er[PPW][NT][NT][2], Ei[PPW][NT][NT][2], Es[PPW][NT][NT][2];
#pragma unroll
  for (int j = 0; j < PPW; ++j) {
#pragma unroll
    for (int a = 0; a < NT; ++a)
#pragma unroll
      for (int b = 0; b < NT; ++b) {
        Er[j][a][b][0] = Er[j][a][b][1] = 0.0; Ei[j][a][b][0] = Ei[j][a][b][1] = 0.0; Es[j][a][b][0] = Es[j][a][b][1] = 0.0;
      }
    if (lane == 0) { Er[j][0][0][0] = 1.0; Es[j][0][0][0] = 1.0; }   // E_0 = [1]
  }

  for (int s = 0; s < n; ++s) {
    const int st = s % QK_NS;
    const uint32_t par = (uint32_t)((s / QK_NS) & 1);
    qk_mbar_wait(qk_smem_u32(&bars[st]), par);
    if (any_active) {
      const int KTx = sDx[s] >> 3, MTx = sDx[s + 1] >> 3;
      const int KTy = sDy[s] >> 3, MTy = sDy[s + 1] >> 3;
      // live k-blocks (of 4) on the contraction side, live 8-tiles on the output side
      const int kby = tcy[s], my = (tcy[s + 1] + 1) >> 1;
      const unsigned char* sb = stage0 + (size_t)st * stage_bytes;
      const double2* by = (const double2*)(sb + (size_t)TI * P.slot_x + (size_t)tj * P.slot_y);
#pragma unroll
      for (int j = 0; j < PPW; ++j) {
        if (!active[j]) continue;
        const unsigned char* tcx = stc + (tk + j) * (n + 1);
        const int kbx = tcx[s], mx = (tcx[s + 1] + 1) >> 1;
        const double2* bx = (const double2*)(sb + (size_t)(tk + j) * P.slot_x);
        qk_site_dispatch<NT>(Er[j], Ei[j], Es[j], bx, by, lane, MTx, KTx, MTy, KTy, kbx, mx, kby, my);
      }
    }
    __syncwarp();
    if (lane == 0) {
      __threadfence_block();                                   // this warp's reads of the stage are done
      if (atomicAdd(&cnt[st], 1) == QK_GRAM_WARPS - 1) {       // last warp out refills the stage
        cnt[st] = 0;
        __threadfence_block();
        if (s + QK_NS < n) issue_site(s + QK_NS);
      }
    }
  }
  if (lane == 0) {
#pragma unroll
    for (int j = 0; j < PPW; ++j) {
      if (!active[j]) continue;
      const int x = x0 + tk + j;
      const double v = Er[j][0][0][0] * Er[j][0][0][0] + Ei[j][0][0][0] * Ei[j][0][0][0];
      P.K[(size_t)y * P.ldk + x] = v;
      if (P.symmetric) P.K[(size_t)x * P.ldk + y] = v;
    }
  }
  // per-tile time (8 inner products; the reference times every vdot, gpu:379-381): last warp out records it
  if (P.unit_clk) {
    __syncthreads();
    if (threadIdx.x == 0) P.unit_clk[blockIdx.x] = clock64() - clk0;
  }
}

// CTA tile shape (kets x bras) of the tensor-core kernel for a given maximum padded bond dimension
void qk_gram_dmma_tile_shape(int maxD, int* ti, int* tj) {
  *ti = (maxD <= 8) ? GramCfg<1>::TI : GramCfg<2>::TI;
  *tj = QK_TJ;
}

template <int NT>
static size_t gram_smem_bytes(const GramParams& P) {
  constexpr int TI = GramCfg<NT>::TI;
  size_t b = (size_t)QK_NS * (TI * (size_t)P.slot_x + QK_TJ * (size_t)P.slot_y);
  b += 2 * QK_NS * sizeof(uint64_t);
  b += 2 * (size_t)(P.n + 1) * sizeof(int);
  b += (size_t)(TI + QK_TJ) * (P.n + 1);
  return (b + 127) & ~(size_t)127;
}

template <int NT>
static cudaError_t launch_gram_nt(const GramParams& P, cudaStream_t stream) {
  const size_t smem = gram_smem_bytes<NT>(P);
  cudaError_t e = cudaFuncSetAttribute(qk_gram_dmma_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  qk_gram_dmma_kernel<NT><<<P.n_cta_tiles, QK_GRAM_WARPS * 32, smem, stream>>>(P);
  return cudaGetLastError();
}

cudaError_t qk_launch_gram_dmma(const GramParams& P, int maxD, cudaStream_t stream) {
  if (P.n_cta_tiles <= 0) return cudaSuccess;
  if (maxD <= 8) return launch_gram_nt<1>(P, stream);
  if (maxD <= 16) return launch_gram_nt<2>(P, stream);
  return cudaErrorInvalidValue;
}

// ------------------------------------------------------------------------------------------------
// Lane-per-pair overlap kernel for bond dimensions <= DM (2 or 4).
// CTA = 16 kets x 8 bras = 128 threads; thread (tx, ty) owns the pair and keeps its transfer matrix E (DM x DM
// complex) in registers for the whole sweep.  Per site the live prefix of the 24 site tensors involved
// (<= DM*2*DM c128 each, straight from the stage-1 store) is double-buffered into shared memory with 16-byte
// cp.async, then re-laid out once per CTA into zero-padded [DM][2][DM] blocks (skewed by 16 B so that the 16
// kets of a warp hit different banks): the per-lane arithmetic is then fully unrolled with no guards.
//   t[a]      = sum_c  E[a][c] A_x[c][p][c']                 (for each p, c')
//   E'[b'][c'] += sum_a conj(A_y[a][p][b']) t[a]
// 8 DM^3 complex MACs per pair and site, all on the FP64 FMA pipe: 64 lanes/clk/SM.
// ------------------------------------------------------------------------------------------------
#define QK_LTX 16
#define QK_LTY 8
void qk_gram_lane_tile_shape(int* tx, int* ty) { *tx = QK_LTX; *ty = QK_LTY; }

template <int DM>
__global__ void __launch_bounds__(QK_LTX * QK_LTY, DM > 2 ? 2 : 4) qk_gram_lane_kernel(const __grid_constant__ LaneParams P) {
  constexpr int NT = QK_LTX + QK_LTY;
  constexpr int RAW = DM * 2 * DM;                    // c128 copied per state and site (live prefix of the slot)
  constexpr int BLK = DM * 2 * DM + 1;                // c128 per zero-padded block (+1: bank skew)
  extern __shared__ __align__(16) unsigned char lsm[];
  const int n = P.n;
  c128* raw = (c128*)lsm;                             // [2][NT][RAW]   as stored: [chi_l][2][chi_r], current dims
  c128* pad = raw + 2 * NT * RAW;                     // [NT][BLK]      re-laid out as [DM][2][DM], zero outside
  int* s_len = (int*)(pad + NT * BLK);                // [n]   c128 copied per state at site s
  int64_t* s_off = (int64_t*)(s_len + n + (n & 1));   // [n]
  unsigned char* s_chi = (unsigned char*)(s_off + n); // [NT][n+1]
  const int tid = threadIdx.x;
  const int tx = tid % QK_LTX, ty = tid / QK_LTX;
  const int4 tile = P.tiles[blockIdx.x];
  const int y0 = tile.x, x0 = tile.y, y_end = tile.z, x_end = tile.w;
  const int x = x0 + tx, y = y0 + ty;
  const bool active = (x < x_end) && (y < y_end) && (!P.symmetric || x <= y);

  for (int s = tid; s < n; s += blockDim.x) {
    const int full = P.cap[s] * 2 * P.cap[s + 1];
    s_len[s] = full < RAW ? full : RAW;
    s_off[s] = P.site_off[s];
  }
  for (int i = tid; i < NT * (n + 1); i += blockDim.x) {
    const int t = i / (n + 1), b = i - t * (n + 1);
    int idx;
    const int32_t* chi;
    if (t < QK_LTX) { idx = x0 + t; if (idx >= P.Nx) idx = P.Nx - 1; chi = P.chiX; }
    else { idx = y0 + (t - QK_LTX); if (idx >= P.Ny) idx = P.Ny - 1; chi = P.chiY; }
    s_chi[i] = (unsigned char)chi[(size_t)idx * (n + 1) + b];
  }
  __syncthreads();

  auto issue = [&](int s, int buf) {
    const int len = s_len[s];
    const int64_t off = s_off[s];
    for (int i = tid; i < NT * len; i += blockDim.x) {
      const int t = i / len, ch = i - t * len;
      const c128* src;
      if (t < QK_LTX) { int idx = x0 + t; if (idx >= P.Nx) idx = P.Nx - 1; src = P.storeX + (size_t)idx * P.state_stride; }
      else { int idx = y0 + (t - QK_LTX); if (idx >= P.Ny) idx = P.Ny - 1; src = P.storeY + (size_t)idx * P.state_stride; }
      const uint32_t dst = qk_smem_u32(raw + ((size_t)buf * NT + t) * RAW + ch);
      asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src + off + ch) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  c128 E[DM][DM];
#pragma unroll
  for (int a = 0; a < DM; ++a)
#pragma unroll
    for (int c = 0; c < DM; ++c) E[a][c] = cmake(0.0, 0.0);
  E[0][0] = cmake(1.0, 0.0);

  issue(0, 0);
  for (int s = 0; s < n; ++s) {
    const int buf = s & 1;
    if (s + 1 < n) issue(s + 1, buf ^ 1);
    else asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 1;" ::: "memory");
    __syncthreads();                                   // site s has landed; everyone is done with `pad`
    // zero-padded fixed-shape copy: the arithmetic below then needs no guard on the per-lane bond dimensions
    for (int i = tid; i < NT * RAW; i += blockDim.x) {
      const int t = i / RAW, e = i - t * RAW;
      const int c = e / (2 * DM), p = (e / DM) & 1, cp = e % DM;
      const int cl = s_chi[t * (n + 1) + s], cr = s_chi[t * (n + 1) + s + 1];
      c128 v = cmake(0.0, 0.0);
      if (c < cl && cp < cr) v = raw[((size_t)buf * NT + t) * RAW + (c * 2 + p) * cr + cp];
      pad[(size_t)t * BLK + e] = v;
    }
    __syncthreads();
    if (active) {
      const c128* Ax = pad + (size_t)tx * BLK;
      const c128* Ay = pad + (size_t)(QK_LTX + ty) * BLK;
      c128 En[DM][DM];
#pragma unroll
      for (int a = 0; a < DM; ++a)
#pragma unroll
        for (int c = 0; c < DM; ++c) En[a][c] = cmake(0.0, 0.0);
#pragma unroll
      for (int p = 0; p < 2; ++p) {
#pragma unroll
        for (int cp = 0; cp < DM; ++cp) {
          c128 t[DM];
#pragma unroll
          for (int a = 0; a < DM; ++a) t[a] = cmake(0.0, 0.0);
#pragma unroll
          for (int c = 0; c < DM; ++c) {
            const c128 ax = Ax[(c * 2 + p) * DM + cp];
#pragma unroll
            for (int a = 0; a < DM; ++a) cfma(t[a], E[a][c], ax);
          }
#pragma unroll
          for (int a = 0; a < DM; ++a) {
#pragma unroll
            for (int bp = 0; bp < DM; ++bp) cfmac(En[bp][cp], Ay[(a * 2 + p) * DM + bp], t[a]);
          }
        }
      }
#pragma unroll
      for (int a = 0; a < DM; ++a)
#pragma unroll
        for (int c = 0; c < DM; ++c) E[a][c] = En[a][c];
    }
  }
  if (active) {
    const double v = E[0][0].x * E[0][0].x + E[0][0].y * E[0][0].y;
    P.K[(size_t)y * P.ldk + x] = v;
    if (P.symmetric) P.K[(size_t)x * P.ldk + y] = v;
  }
}

template <int DM>
static cudaError_t launch_lane(const LaneParams& P, cudaStream_t stream) {
  constexpr int NT = QK_LTX + QK_LTY;
  constexpr int RAW = DM * 2 * DM, BLK = RAW + 1;
  size_t smem = (size_t)(2 * NT * RAW + NT * BLK) * sizeof(c128);
  smem += (size_t)(P.n + (P.n & 1)) * sizeof(int) + (size_t)P.n * sizeof(int64_t) + (size_t)NT * (P.n + 1);
  smem = (smem + 15) & ~(size_t)15;
  cudaError_t e = cudaFuncSetAttribute(qk_gram_lane_kernel<DM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  qk_gram_lane_kernel<DM><<<P.n_cta_tiles, QK_LTX * QK_LTY, smem, stream>>>(P);
  return cudaGetLastError();
}

cudaError_t qk_launch_gram_lane(const LaneParams& P, int dm, cudaStream_t stream) {
  if (P.n_cta_tiles <= 0) return cudaSuccess;
  if (dm <= 2) return launch_lane<2>(P, stream);
  if (dm <= 4) return launch_lane<4>(P, stream);
  return cudaErrorInvalidValue;
}

// ------------------------------------------------------------------------------------------------
// CUDA-core cross-check kernel: one CTA per (y, x) pair, E and T in shared memory, true dims.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(64) qk_gram_store_kernel(int n, const c128* __restrict__ storeX, int64_t strideX,
                                                           const int64_t* __restrict__ offX, const int32_t* __restrict__ chiX,
                                                           int capx, const c128* __restrict__ storeY, int64_t strideY,
                                                           const int64_t* __restrict__ offY, const int32_t* __restrict__ chiY,
                                                           int capy, double* __restrict__ K, int64_t ldk) {
  extern __shared__ __align__(16) unsigned char ssm[];
  c128* E = (c128*)ssm;                       // [capy][capx]
  c128* T = E + (size_t)capy * capx;          // [2*capy][capx]
  const int x = blockIdx.x, y = blockIdx.y;
  const c128* Sx = storeX + (size_t)x * strideX;
  const c128* Sy = storeY + (size_t)y * strideY;
  const int32_t* cx = chiX + (size_t)x * (n + 1);
  const int32_t* cy = chiY + (size_t)y * (n + 1);
  if (threadIdx.x == 0) E[0] = cmake(1.0, 0.0);
  __syncthreads();
  for (int s = 0; s < n; ++s) {
    const int cxl = cx[s], cxr = cx[s + 1], cyl = cy[s], cyr = cy[s + 1];
    const c128* Ax = Sx + offX[s];
    const c128* Ay = Sy + offY[s];
    // T[(a,p), c'] = sum_c E[a,c] Ax[c,p,c']
    for (int idx = threadIdx.x; idx < cyl * 2 * cxr; idx += blockDim.x) {
      const int a = idx / (2 * cxr);
      const int rem = idx - a * 2 * cxr;
      const int p = rem / cxr, cp = rem - p * cxr;
      c128 acc = cmake(0, 0);
      for (int c = 0; c < cxl; ++c) cfma(acc, E[a * cxl + c], Ax[(size_t)(c * 2 + p) * cxr + cp]);
      T[idx] = acc;
    }
    __syncthreads();
    // E'[b', c'] = sum_{a,p} conj(Ay[a,p,b']) T[(a,p), c']
    for (int idx = threadIdx.x; idx < cyr * cxr; idx += blockDim.x) {
      const int bp = idx / cxr, cp = idx - bp * cxr;
      c128 acc = cmake(0, 0);
      for (int ap = 0; ap < 2 * cyl; ++ap) cfmac(acc, Ay[(size_t)ap * cyr + bp], T[ap * cxr + cp]);
      E[idx] = acc;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) K[(size_t)y * ldk + x] = E[0].x * E[0].x + E[0].y * E[0].y;
}

cudaError_t qk_launch_gram_store(int n, const c128* storeX, int64_t strideX, const int64_t* site_off_x,
                                 const int32_t* chiX, int capx, int Nx, const c128* storeY, int64_t strideY,
                                 const int64_t* site_off_y, const int32_t* chiY, int capy, int Ny, double* K,
                                 int64_t ldk, cudaStream_t stream) {
  if (Nx <= 0 || Ny <= 0) return cudaSuccess;
  const size_t smem = (size_t)3 * capx * capy * sizeof(c128);
  cudaError_t e = cudaFuncSetAttribute(qk_gram_store_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  dim3 grid(Nx, Ny);
  qk_gram_store_kernel<<<grid, 64, smem, stream>>>(n, storeX, strideX, site_off_x, chiX, capx, storeY, strideY,
                                                   site_off_y, chiY, capy, K, ldk);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// Generic overlap kernel on the frag exchange format (CUDA cores, any padded D <= 32): one CTA per pair.
// Used when a bond dimension exceeds the register-resident tensor-core kernel's D <= 16; because it reads
// the same packed buffers, the multi-GPU exchange path is unchanged.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ c128 qk_frag_elem(const double* __restrict__ blk, int MT, int KT, int p, int c, int cp) {
  const int kt = c >> 3, e = (c >> 2) & 1, mt = cp >> 3, r = cp & 7;
  const int lane = (c & 3) | ((r >> 2) << 2) | ((r & 3) << 3);
  const size_t base = ((((size_t)(p * MT + mt) * KT + kt) * 2) * 32 + lane) * 2 + e;   // h = 0 (real part)
  return cmake(blk[base], blk[base + 64]);                                              // h = 1: +32 lanes * 2
}

__global__ void __launch_bounds__(64) qk_gram_frag_generic_kernel(const __grid_constant__ GramParams P,
                                                                   const int2* __restrict__ pairs, int dmax) {
  extern __shared__ __align__(16) unsigned char ssm[];
  c128* E = (c128*)ssm;                         // [dmax][dmax]
  c128* T = E + (size_t)dmax * dmax;            // [2*dmax][dmax]
  const int y = pairs[blockIdx.x].x, x = pairs[blockIdx.x].y;
  const unsigned char* bxs = P.fragX + (size_t)x * P.strideX;
  const unsigned char* bys = P.fragY + (size_t)y * P.strideY;
  const unsigned char* tcx = bxs + P.dataX;     // ceil(chi / 4) per bond
  const unsigned char* tcy = bys + P.dataY;
  // extents below are rounded up to 4, so E is read beyond the true bond dimension: it must be zero there
  for (int i = threadIdx.x; i < dmax * dmax; i += blockDim.x) E[i] = cmake(i == 0 ? 1.0 : 0.0, 0.0);
  __syncthreads();
  for (int s = 0; s < P.n; ++s) {
    const int KTx = P.Dx[s] >> 3, MTx = P.Dx[s + 1] >> 3, KTy = P.Dy[s] >> 3, MTy = P.Dy[s + 1] >> 3;
    const int cxl = 4 * tcx[s], cxr = 4 * tcx[s + 1], cyl = 4 * tcy[s], cyr = 4 * tcy[s + 1];   // live extents
    const double* ax = (const double*)(bxs + P.offx[s]);
    const double* ay = (const double*)(bys + P.offy[s]);
    // T[(a,p), c'] = sum_c E[a,c] A_x[c,p,c']        (E stored with row stride = live ket extent of bond s)
    for (int idx = threadIdx.x; idx < cyl * 2 * cxr; idx += blockDim.x) {
      const int a = idx / (2 * cxr);
      const int rem = idx - a * 2 * cxr;
      const int p = rem / cxr, cp = rem - p * cxr;
      c128 acc = cmake(0, 0);
      for (int c = 0; c < cxl; ++c) cfma(acc, E[a * cxl + c], qk_frag_elem(ax, MTx, KTx, p, c, cp));
      T[idx] = acc;
    }
    __syncthreads();
    // E'[b', c'] = sum_{a,p} conj(A_y[a,p,b']) T[(a,p), c']
    for (int idx = threadIdx.x; idx < cyr * cxr; idx += blockDim.x) {
      const int bp = idx / cxr, cp = idx - bp * cxr;
      c128 acc = cmake(0, 0);
      for (int a = 0; a < cyl; ++a)
        for (int p = 0; p < 2; ++p) cfmac(acc, qk_frag_elem(ay, MTy, KTy, p, a, bp), T[(a * 2 + p) * cxr + cp]);
      E[idx] = acc;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const double v = E[0].x * E[0].x + E[0].y * E[0].y;
    P.K[(size_t)y * P.ldk + x] = v;
    if (P.symmetric) P.K[(size_t)x * P.ldk + y] = v;
  }
}

cudaError_t qk_launch_gram_frag_generic(const GramParams& P, const int2* pairs_dev, int n_pairs, int maxD,
                                        cudaStream_t stream) {
  if (n_pairs <= 0) return cudaSuccess;
  const size_t smem = (size_t)3 * maxD * maxD * sizeof(c128);
  cudaError_t e = cudaFuncSetAttribute(qk_gram_frag_generic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  qk_gram_frag_generic_kernel<<<n_pairs, 64, smem, stream>>>(P, pairs_dev, maxD);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// DMMA peak microbenchmark (roofline denominator for stage 2; MEASURED_PEAKS.json has no FP64 figure)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) qk_dmma_peak_kernel(int iters, double* sink) {
  double acc[8][2];
#pragma unroll
  for (int i = 0; i < 8; ++i) { acc[i][0] = 0.0; acc[i][1] = 0.0; }
  const double a = 1.0 + 1e-9 * threadIdx.x, b = 1.0 - 1e-9 * threadIdx.x;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) qk_dmma(acc[i], a, b);
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += acc[i][0] + acc[i][1];
  if (s == 123.456) sink[0] = s;
}

// Variant with distinct A / B operand registers for every MMA (what a real contraction issues); the
// kernel above feeds all eight MMAs of an iteration from the same two registers.
__global__ void __launch_bounds__(256) qk_dmma_peak_kernel_distinct(int iters, double* sink) {
  double acc[8][2], a[8], b[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    acc[i][0] = 0.0; acc[i][1] = 0.0;
    a[i] = 1.0 + 1e-9 * (threadIdx.x + 3 * i);
    b[i] = 1.0 - 1e-9 * (threadIdx.x + 5 * i);
  }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) qk_dmma(acc[i], a[i], b[(i + it) & 7]);
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += acc[i][0] + acc[i][1];
  if (s == 123.456) sink[0] = s;
}

cudaError_t qk_run_dmma_peak(int iters, double* tflops) {
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  double* sink = nullptr;
  cudaError_t e = cudaMalloc(&sink, sizeof(double));
  if (e != cudaSuccess) return e;
  int grid = sms * 4, block = 256;
  if (const char* e = getenv("QK_PEAK_WARPS_PER_SM")) {   // experiments: DMMA rate at a given occupancy
    const int w = atoi(e);
    if (w >= 1 && w <= 8) { grid = sms; block = 32 * w; }
    else if (w > 8 && w <= 64) { grid = sms * (w / 8); block = 256; }
  }
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const bool distinct = getenv("QK_PEAK_DISTINCT") != nullptr;
  qk_dmma_peak_kernel<<<grid, block>>>(iters / 8 + 1, sink);   // warm-up
  float best = 1e30f;
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0);
    if (distinct) qk_dmma_peak_kernel_distinct<<<grid, block>>>(iters, sink);
    else qk_dmma_peak_kernel<<<grid, block>>>(iters, sink);
    cudaEventRecord(e1);
    e = cudaEventSynchronize(e1);
    if (e != cudaSuccess) break;
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(sink);
  if (e != cudaSuccess) return e;
  const double flops = (double)grid * (block / 32) * (double)iters * 8.0 * 512.0;
  *tflops = flops / (best * 1e-3) / 1e12;
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// Pipe co-issue microbenchmark: warps 0-3 of a CTA (one per scheduler) issue DMMAs, warps 4-7 DFMAs.
// Per iteration a DMMA warp occupies its pipe for 8 x 16 cycles and a DFMA warp for 64 x 2 cycles.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) qk_pipe_mix_kernel(int iters, int mode, double* sink) {
  const int warp = threadIdx.x >> 5;
  double s = 0.0;
  if (warp < 4) {
    if (mode & 1) {
      double acc[8][2];
#pragma unroll
      for (int i = 0; i < 8; ++i) { acc[i][0] = 0.0; acc[i][1] = 0.0; }
      const double a = 1.0 + 1e-9 * threadIdx.x, b = 1.0 - 1e-9 * threadIdx.x;
      for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) qk_dmma(acc[i], a, b);
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) s += acc[i][0] + acc[i][1];
    }
  } else if (mode & 2) {
    double acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = 1e-3 * (threadIdx.x + i);
    const double a = 1.0 + 1e-9 * threadIdx.x, b = 1e-9 * threadIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int rep = 0; rep < 4; ++rep)
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] = fma(acc[i], a, b);
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) s += acc[i];
  }
  if (s == 123.456) sink[0] = s;
}

cudaError_t qk_run_pipe_mix(int iters, float* ms3) {
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  double* sink = nullptr;
  cudaError_t e = cudaMalloc(&sink, sizeof(double));
  if (e != cudaSuccess) return e;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  qk_pipe_mix_kernel<<<sms, 256>>>(iters / 8 + 1, 3, sink);   // warm-up
  for (int mode = 1; mode <= 3 && e == cudaSuccess; ++mode) {
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
      cudaEventRecord(e0);
      qk_pipe_mix_kernel<<<sms, 256>>>(iters, mode, sink);
      cudaEventRecord(e1);
      e = cudaEventSynchronize(e1);
      if (e != cudaSuccess) break;
      float ms = 0.f;
      cudaEventElapsedTime(&ms, e0, e1);
      if (ms < best) best = ms;
    }
    ms3[mode - 1] = best;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(sink);
  return e != cudaSuccess ? e : cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// Stage 2 for bond dimensions above the register-resident kernels (chi > 16; BASELINE config 4): the transfer
// sweep of every (bra, ket) pair as two batched complex GEMMs per site on the FP64 tensor cores,
//     step 1   T[a,(p,c')]  = sum_c      E[a,c]              A_x[c,(p,c')]      M = chi_y(l), K = chi_x(l), N = 2 chi_x(r)
//     step 2   E'[b',c']    = sum_{(a,p)} conj(A_y[(a,p),b']) T[(a,p),c']        M = chi_y(r), K = 2 chi_y(l), N = chi_x(r)
// with E and T of every pair of the launch in global memory (compact row-major, true bond dimensions) and the site
// tensors read straight from the unpadded stage-1 stores ([chi_l][2][chi_r] is already the row-major K x N / K x M
// operand).  At chi ~ 100 one site step of one pair is 32 chi^3 = 32 MFLOP against ~1.5 MB of operands: compute
// bound.  CTA = 4 warps, tile 32 x 64, each warp 16 x 32 as 2 x 4 DMMA m8n8k4 tiles; complex product = 4 real MMAs
// (no 3M here: the operand sums would need an extra pass over shared memory and the tiles are full at these sizes).
// ------------------------------------------------------------------------------------------------
struct BigGemmArgs {
  const int2* pairs;          // (y, x) per pair of this launch
  int nb;                     // n + 1
  const int32_t* chiX;        // [Nx][nb]
  const int32_t* chiY;
  const c128* storeX;
  const c128* storeY;
  int64_t strideX, strideY;   // c128 per state
  int64_t off_x, off_y;       // slot offset of this site inside a state, c128
  c128* E;
  c128* T;
  int64_t e_stride, t_stride; // c128 per pair
  int site;
  int tiles_n, tpp;           // tiles along N, tiles per pair (maxima over the batch)
};

template <int STEP>
__global__ void __launch_bounds__(128) qk_big_gemm_kernel(const __grid_constant__ BigGemmArgs a) {
  const int pair = blockIdx.x / a.tpp, tile = blockIdx.x - pair * a.tpp;
  const int tm = tile / a.tiles_n, tn = tile - tm * a.tiles_n;
  const int2 yx = a.pairs[pair];
  const int32_t* cx = a.chiX + (size_t)yx.y * a.nb + a.site;
  const int32_t* cy = a.chiY + (size_t)yx.x * a.nb + a.site;
  const int cxl = cx[0], cxr = cx[1], cyl = cy[0], cyr = cy[1];
  int M, N, K, lda, ldb;
  const c128 *A, *B;
  c128* C;
  if (STEP == 1) {
    M = cyl; K = cxl; N = 2 * cxr; lda = K; ldb = N;
    A = a.E + (size_t)pair * a.e_stride;
    B = a.storeX + (size_t)yx.y * a.strideX + a.off_x;
    C = a.T + (size_t)pair * a.t_stride;
  } else {
    M = cyr; K = 2 * cyl; N = cxr; lda = M; ldb = N;
    A = a.storeY + (size_t)yx.x * a.strideY + a.off_y;
    B = a.T + (size_t)pair * a.t_stride;
    C = a.E + (size_t)pair * a.e_stride;
  }
  const int m0 = tm * 32, n0 = tn * 64;
  if (m0 >= M || n0 >= N) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, j4 = lane & 3;
  const int mw = m0 + (warp >> 1) * 16, nw = n0 + (warp & 1) * 32;
  double cr[2][4][2], ci[2][4][2];
#pragma unroll
  for (int i = 0; i < 2; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) { cr[i][j][0] = cr[i][j][1] = 0.0; ci[i][j][0] = ci[i][j][1] = 0.0; }
  if (mw < M && nw < N) {
#pragma unroll 2
    for (int k0 = 0; k0 < K; k0 += 4) {
      const int kk = k0 + j4;
      double ar[2], ai[2], an[2], br[4], bi[4];
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int row = mw + i * 8 + g;
        double2 v = make_double2(0.0, 0.0);
        if (row < M && kk < K) v = *(const double2*)(STEP == 1 ? A + (size_t)row * lda + kk : A + (size_t)kk * lda + row);
        ar[i] = v.x;
        ai[i] = STEP == 1 ? v.y : -v.y;    // step 2 uses conj(A_y)
        an[i] = -ai[i];
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int col = nw + j * 8 + g;
        double2 v = make_double2(0.0, 0.0);
        if (kk < K && col < N) v = *(const double2*)(B + (size_t)kk * ldb + col);
        br[j] = v.x; bi[j] = v.y;
      }
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          qk_dmma(cr[i][j], ar[i], br[j]);
          qk_dmma(ci[i][j], ar[i], bi[j]);
          qk_dmma(cr[i][j], an[i], bi[j]);
          qk_dmma(ci[i][j], ai[i], br[j]);
        }
    }
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int row = mw + i * 8 + g;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int col = nw + j * 8 + 2 * j4;
      if (row < M) {
        if (col < N) *(double2*)(C + (size_t)row * N + col) = make_double2(cr[i][j][0], ci[i][j][0]);
        if (col + 1 < N) *(double2*)(C + (size_t)row * N + col + 1) = make_double2(cr[i][j][1], ci[i][j][1]);
      }
    }
  }
}

__global__ void qk_big_init_kernel(c128* E, int64_t e_stride, int n_pairs) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p < n_pairs) E[(size_t)p * e_stride] = cmake(1.0, 0.0);     // E_0 = [1]
}

__global__ void qk_big_finish_kernel(const c128* E, int64_t e_stride, const int2* pairs, int n_pairs, int symmetric,
                                     double* K, int64_t ldk) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_pairs) return;
  const c128 v = E[(size_t)p * e_stride];
  const double k = v.x * v.x + v.y * v.y;
  const int2 yx = pairs[p];
  K[(size_t)yx.x * ldk + yx.y] = k;
  if (symmetric) K[(size_t)yx.y * ldk + yx.x] = k;
}

// One chunk of pairs through the whole sweep.  dims_x / dims_y: per-bond maxima of the bond dimensions (host).
cudaError_t qk_launch_gram_big(int n, const int64_t* site_off_x, const int64_t* site_off_y, const int32_t* dims_x,
                               const int32_t* dims_y, const c128* storeX, int64_t strideX, const int32_t* chiX,
                               const c128* storeY, int64_t strideY, const int32_t* chiY, const int2* pairs_dev,
                               int n_pairs, int symmetric, c128* E, int64_t e_stride, c128* T, int64_t t_stride,
                               double* K, int64_t ldk, cudaStream_t stream) {
  if (n_pairs <= 0) return cudaSuccess;
  qk_big_init_kernel<<<(n_pairs + 255) / 256, 256, 0, stream>>>(E, e_stride, n_pairs);
  BigGemmArgs a;
  a.pairs = pairs_dev; a.nb = n + 1; a.chiX = chiX; a.chiY = chiY; a.storeX = storeX; a.storeY = storeY;
  a.strideX = strideX; a.strideY = strideY; a.E = E; a.T = T; a.e_stride = e_stride; a.t_stride = t_stride;
  for (int s = 0; s < n; ++s) {
    a.site = s; a.off_x = site_off_x[s]; a.off_y = site_off_y[s];
    a.tiles_n = (2 * dims_x[s + 1] + 63) / 64;
    a.tpp = ((dims_y[s] + 31) / 32) * a.tiles_n;
    qk_big_gemm_kernel<1><<<(unsigned)((size_t)n_pairs * a.tpp), 128, 0, stream>>>(a);
    a.tiles_n = (dims_x[s + 1] + 63) / 64;
    a.tpp = ((dims_y[s + 1] + 31) / 32) * a.tiles_n;
    qk_big_gemm_kernel<2><<<(unsigned)((size_t)n_pairs * a.tpp), 128, 0, stream>>>(a);
  }
  qk_big_finish_kernel<<<(n_pairs + 255) / 256, 256, 0, stream>>>(E, e_stride, pairs_dev, n_pairs, symmetric, K, ldk);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// repack: states of one store layout -> another (different bond caps => different slot offsets); one CTA per
// (site, state).  Merges batches simulated with different caps into the one buffer that is exchanged.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) qk_repack_kernel(int n, const c128* __restrict__ src, int64_t src_stride,
                                                        const int64_t* __restrict__ src_off, const int32_t* __restrict__ src_chi,
                                                        c128* __restrict__ dst, int64_t dst_stride,
                                                        const int64_t* __restrict__ dst_off, int32_t* __restrict__ dst_chi,
                                                        const int32_t* __restrict__ dst_index) {
  const int s = blockIdx.x, i = blockIdx.y;
  const int idst = dst_index ? dst_index[i] : i;
  if (idst < 0) return;
  const int32_t* chi = src_chi + (size_t)i * (n + 1);
  int64_t len = (int64_t)chi[s] * 2 * chi[s + 1];
  const int64_t capacity = dst_off[s + 1] - dst_off[s];
  if (len > capacity) len = capacity;       // cannot happen when the destination caps cover the states (host contract)
  const c128* a = src + (size_t)i * src_stride + src_off[s];
  c128* b = dst + (size_t)idst * dst_stride + dst_off[s];
  for (int64_t k = threadIdx.x; k < len; k += blockDim.x) b[k] = a[k];
  if (s == 0)
    for (int k = threadIdx.x; k <= n; k += blockDim.x) dst_chi[(size_t)idst * (n + 1) + k] = chi[k];
}

cudaError_t qk_launch_repack(int n, int N, const c128* src, int64_t src_stride, const int64_t* src_off_dev,
                             const int32_t* src_chi, c128* dst, int64_t dst_stride, const int64_t* dst_off_dev,
                             int32_t* dst_chi, const int32_t* dst_index_dev, cudaStream_t stream) {
  if (N <= 0) return cudaSuccess;
  dim3 grid(n, N);
  qk_repack_kernel<<<grid, 128, 0, stream>>>(n, src, src_stride, src_off_dev, src_chi, dst, dst_stride, dst_off_dev,
                                             dst_chi, dst_index_dev);
  return cudaGetLastError();
}
