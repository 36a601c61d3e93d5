// Launch interfaces of the device kernels (implemented in qk_sim.cu / qk_gram.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "qk_types.h"

// ---- stage 1 ----
// Launches the persistent simulation kernel for group size G in {32, 64, 128, 256}.
cudaError_t qk_launch_sim(const SimParams& P, int G, size_t smem_bytes, int* work_counter, cudaStream_t stream,
                          int* grid_out);

// B-form kernel: one cluster of `ncta` CTAs per datapoint; parts = device scratch [N][ncta] QkStat
cudaError_t qk_launch_sim_b(const SimParams& P, int G, size_t smem_bytes, int ncta, QkStat* parts, cudaStream_t stream,
                            int* grid_out);

// Large-matrix path (chi_cap > 32, qk_sim_big.h): ncta_req <= 0 picks the cluster size from the SM count and N.
cudaError_t qk_sim_big_config(size_t smem_bytes, int N, int ncta_req, int* ncta_out, int* n_clusters_out);
cudaError_t qk_launch_sim_big(const SimParams& P, size_t smem_bytes, int ncta, int n_clusters, cudaStream_t stream);

// ---- exchange format ("frag") ----
// Per state: for every site s a block of Dl*Dr*4 doubles in DMMA A-fragment order, then n+1 bytes
// (padded to 16) holding ceil(chi_b / 8) per bond.  See DESIGN.md "Data layout".
struct FragLayout {
  int n;
  int64_t data_bytes;      // bytes of fragment data per state
  int64_t stride_bytes;    // data_bytes + align16(n+1)
};
void qk_frag_layout(int n, const int32_t* D, FragLayout* L, int64_t* site_off_bytes /*[n+1]*/);

cudaError_t qk_launch_pack(int n, int N, const c128* store, int64_t state_stride, const int64_t* site_off_dev,
                           const int32_t* chi_dev, const int32_t* D_dev, const int64_t* frag_off_dev,
                           int64_t frag_stride_bytes, int64_t frag_data_bytes, void* frag_dev, const int32_t* dst_index_dev,
                           int dst_base, cudaStream_t stream);

// ---- stage 2 ----
struct GramParams {
  int n;
  const int32_t* Dx;        // device [n+1]
  const int32_t* Dy;        // device [n+1]
  const int64_t* offx;      // device [n+1] byte offsets of the site blocks
  const int64_t* offy;
  const unsigned char* fragX;
  const unsigned char* fragY;
  int64_t strideX, strideY; // bytes per state
  int64_t dataX, dataY;     // bytes of fragment data per state (tile-count bytes follow)
  int Nx, Ny;
  const int4* tiles;        // device [n_cta_tiles]: (y0, x0, y_end, x_end)
  int n_cta_tiles;
  int symmetric;
  double* K;
  int64_t ldk;
  int slot_x, slot_y;       // bytes reserved per state per pipeline stage
  long long* unit_clk;      // optional device [n_cta_tiles]: clock64 ticks every CTA tile took, or NULL
};
// DMMA + bulk-copy pipeline kernel; requires max(D) <= 16.
cudaError_t qk_launch_gram_dmma(const GramParams& P, int maxD, cudaStream_t stream);
void qk_gram_dmma_tile_shape(int maxD, int* ti, int* tj);
// CUDA-core kernel on the same frag buffers, any padded D <= 32 (one CTA per listed (y, x) pair)
cudaError_t qk_launch_gram_frag_generic(const GramParams& P, const int2* pairs_dev, int n_pairs, int maxD,
                                        cudaStream_t stream);

// Low-bond-dimension overlap kernel (chi <= 4): one LANE per (bra, ket) pair on the FP64 CUDA cores, reading the
// unpadded stage-1 stores.  An 8x8x4 DMMA tile is at most 1/4 (chi = 4) or 1/16 (chi = 2) occupied in this regime.
struct LaneParams {
  int n;
  const int32_t* cap;        // device [n+1] bond caps of the plan (slot capacities)
  const int64_t* site_off;   // device [n+1] slot offsets inside one state, c128 units
  int64_t state_stride;      // c128 units per state
  const c128* storeX;
  const int32_t* chiX;       // device [Nx][n+1]
  const c128* storeY;
  const int32_t* chiY;
  int Nx, Ny;
  const int4* tiles;         // device [n_cta_tiles]: (y0, x0, y_end, x_end)
  int n_cta_tiles;
  int symmetric;
  double* K;
  int64_t ldk;
};
cudaError_t qk_launch_gram_lane(const LaneParams& P, int dm /* 2 or 4 */, cudaStream_t stream);
void qk_gram_lane_tile_shape(int* tx, int* ty);

// Batched-GEMM overlap sweep on the unpadded stores for any bond dimension (qk_gram.cu): one chunk of pairs.
cudaError_t qk_launch_gram_big(int n, const int64_t* site_off_x, const int64_t* site_off_y, const int32_t* dims_x,
                               const int32_t* dims_y, const c128* storeX, int64_t strideX, const int32_t* chiX,
                               const c128* storeY, int64_t strideY, const int32_t* chiY, const int2* pairs_dev,
                               int n_pairs, int symmetric, c128* E, int64_t e_stride, c128* T, int64_t t_stride,
                               double* K, int64_t ldk, cudaStream_t stream);

cudaError_t qk_launch_repack(int n, int N, const c128* src, int64_t src_stride, const int64_t* src_off_dev,
                             const int32_t* src_chi, c128* dst, int64_t dst_stride, const int64_t* dst_off_dev,
                             int32_t* dst_chi, const int32_t* dst_index_dev, cudaStream_t stream);

// CUDA-core cross-check on the unpadded stores
cudaError_t qk_launch_gram_store(int n, const c128* storeX, int64_t strideX, const int64_t* site_off_x,
                                 const int32_t* chiX, int capx, int Nx,
                                 const c128* storeY, int64_t strideY, const int64_t* site_off_y,
                                 const int32_t* chiY, int capy, int Ny,
                                 double* K, int64_t ldk, cudaStream_t stream);

cudaError_t qk_run_dmma_peak(int iters, double* tflops);
cudaError_t qk_run_pipe_mix(int iters, float* ms3);
