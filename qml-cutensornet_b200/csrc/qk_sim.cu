// Stage 1 kernel: persistent grid, one cooperative group (= one CTA of G threads) per datapoint,
// interpreting the static op schedule compiled by qk_plan.cpp.  The algorithm itself is in
// qk_sim_core.h (shared with the test-only host emulation).
//
// B200 mapping: grid = SMs x resident CTAs; each CTA keeps theta / rotation matrices in shared
// memory for the whole SVD and streams the two site tensors of an op from its own state slot, which
// stays L2-resident (working set = resident CTAs x one MPS, far below the 126 MB L2); HBM sees the
// final state once.  Datapoints are handed out through an atomic counter so ragged bond dimensions
// do not leave SMs idle.
#include <cooperative_groups.h>
#include <stdlib.h>
#include "qk_kernels.cuh"
// level barrier of the B-form path: the CTAs of one thread-block cluster share a datapoint
#define QK_GROUP_SYNC() do { __threadfence(); cooperative_groups::this_cluster().sync(); __threadfence(); } while (0)
// Cluster barrier of the large-matrix path.  barrier.cluster.arrive / wait already have release / acquire semantics
// at cluster scope; the gpu-scope fences around them are kept as a belt-and-braces measure for the global-memory
// traffic between the CTAs -- measured cost: none (C4 shape, 8 CTAs per cluster: 5.20 s with, 5.10 s without).
#ifndef QK_CSYNC_FENCE
#define QK_CSYNC_FENCE 1
#endif
#define QK_CSYNC(c) do { if ((c).ncta > 1) { if (QK_CSYNC_FENCE) __threadfence(); cooperative_groups::this_cluster().sync(); \
                                             if (QK_CSYNC_FENCE) __threadfence(); } \
                         else __syncthreads(); } while (0)
#include "qk_sim_big.h"

// resident CTAs per SM the register allocation is sized for (shared memory allows 6 at chi_cap 16)
template <int G> struct SimMinBlocks { static constexpr int value = (G == 128) ? 6 : (G == 64) ? 8 : (G == 32) ? 12 : 2; };

template <int G>
__global__ void __launch_bounds__(G, SimMinBlocks<G>::value) qk_sim_kernel(const __grid_constant__ SimParams P, int* work_counter) {
  extern __shared__ __align__(16) unsigned char qk_smem[];
  __shared__ int next_dp;
  SimCtx c;
  qk_sim_carve(c, &P, qk_smem, G);
  for (;;) {
    if (threadIdx.x == 0) next_dp = atomicAdd(work_counter, 1);
    __syncthreads();
    const int dp = next_dp;
    __syncthreads();
    if (dp >= P.N) break;
    const long long t0 = clock64();
    qk_sim_datapoint<G>(c, dp);
    if (P.unit_clk && threadIdx.x == 0) P.unit_clk[dp] = clock64() - t0;   // per-circuit time (reference gpu:220-222)
  }
}

template <int G>
static cudaError_t launch_g(const SimParams& P, size_t smem, int* counter, cudaStream_t stream, int* grid_out) {
  cudaError_t e = cudaFuncSetAttribute(qk_sim_kernel<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  int dev = 0, sms = 0, occ = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, qk_sim_kernel<G>, G, smem);
  if (e != cudaSuccess) return e;
  if (occ < 1) return cudaErrorInvalidConfiguration;
  long long grid = (long long)sms * occ;
  if (grid > P.N) grid = P.N;
  if (grid < 1) grid = 1;
  if (grid_out) *grid_out = (int)grid;
  e = cudaMemsetAsync(counter, 0, sizeof(int), stream);
  if (e != cudaSuccess) return e;
  qk_sim_kernel<G><<<(unsigned)grid, G, smem, stream>>>(P, counter);
  return cudaGetLastError();
}

cudaError_t qk_launch_sim(const SimParams& P, int G, size_t smem_bytes, int* work_counter, cudaStream_t stream,
                          int* grid_out) {
  switch (G) {
    case 32: return launch_g<32>(P, smem_bytes, work_counter, stream, grid_out);
    case 64: return launch_g<64>(P, smem_bytes, work_counter, stream, grid_out);
    case 128: return launch_g<128>(P, smem_bytes, work_counter, stream, grid_out);
    case 256: return launch_g<256>(P, smem_bytes, work_counter, stream, grid_out);
    default: return cudaErrorInvalidValue;
  }
}

// ------------------------------------------------------------------------------------------------
// B-form kernel (QK_PLAN_PARALLEL): one thread-block CLUSTER per datapoint.  The ops of a level touch
// disjoint sites, so the cluster's CTAs take them round-robin and meet at a cluster barrier between levels;
// the dependency depth of the ansatz (C3: 28 levels for 386 two-qubit ops) replaces the op count as the
// latency of one datapoint.  State, bond dimensions and Schmidt values live in global memory (L2).
// ------------------------------------------------------------------------------------------------
template <int G>
__global__ void __launch_bounds__(G, SimMinBlocks<G>::value) qk_sim_kernel_b(const __grid_constant__ SimParams P, QkStat* parts) {
  extern __shared__ __align__(16) unsigned char qk_smem[];
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  const int ncta = (int)cluster.num_blocks();
  const int cta = (int)cluster.block_rank();
  const int n_clusters = (int)(gridDim.x / ncta);
  SimCtx c;
  qk_sim_carve(c, &P, qk_smem, G);
  for (int dp = (int)(blockIdx.x / ncta); dp < P.N; dp += n_clusters) {
    const long long t0 = clock64();
    qk_sim_datapoint_b<G>(c, dp, cta, ncta, parts + (size_t)dp * ncta);
    QK_GROUP_SYNC();
    if (P.unit_clk && cta == 0 && threadIdx.x == 0) P.unit_clk[dp] = clock64() - t0;
  }
}

template <int G>
static cudaError_t launch_b(const SimParams& P, size_t smem, int ncta, QkStat* parts, cudaStream_t stream, int* grid_out) {
  cudaError_t e = cudaFuncSetAttribute(qk_sim_kernel_b<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)ncta; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.blockDim = dim3(G); cfg.dynamicSmemBytes = smem; cfg.stream = stream; cfg.attrs = attr; cfg.numAttrs = 1;
  cfg.gridDim = dim3((unsigned)ncta);
  int max_clusters = 0;
  e = cudaOccupancyMaxActiveClusters(&max_clusters, qk_sim_kernel_b<G>, &cfg);
  if (e != cudaSuccess) return e;
  if (max_clusters < 1) return cudaErrorInvalidConfiguration;
  long long clusters = max_clusters < P.N ? max_clusters : P.N;
  if (clusters < 1) clusters = 1;
  cfg.gridDim = dim3((unsigned)(clusters * ncta));
  if (grid_out) *grid_out = (int)(clusters * ncta);
  return cudaLaunchKernelEx(&cfg, qk_sim_kernel_b<G>, P, parts);
}

cudaError_t qk_launch_sim_b(const SimParams& P, int G, size_t smem_bytes, int ncta, QkStat* parts, cudaStream_t stream,
                            int* grid_out) {
  switch (G) {
    case 32: return launch_b<32>(P, smem_bytes, ncta, parts, stream, grid_out);
    case 64: return launch_b<64>(P, smem_bytes, ncta, parts, stream, grid_out);
    case 128: return launch_b<128>(P, smem_bytes, ncta, parts, stream, grid_out);
    case 256: return launch_b<256>(P, smem_bytes, ncta, parts, stream, grid_out);
    default: return cudaErrorInvalidValue;
  }
}

// ------------------------------------------------------------------------------------------------
// Large-matrix kernel (qk_sim_big.h): one cluster of `ncta` CTAs x 256 threads per datapoint, persistent clusters.
// ------------------------------------------------------------------------------------------------
template <int G>
__global__ void __launch_bounds__(G, 1) qk_sim_big_kernel(const __grid_constant__ SimParams P) {
  extern __shared__ __align__(16) unsigned char qk_smem[];
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  const int ncta = (int)cluster.num_blocks();
  const int cta = (int)cluster.block_rank();
  const int n_clusters = (int)(gridDim.x / ncta);
  const int slot = (int)(blockIdx.x / ncta);
  SimCtx c;
  qk_big_carve(c, &P, qk_smem, G, slot, cta, ncta);
  for (int dp = slot; dp < P.N; dp += n_clusters) {
    const long long t0 = clock64();
    qk_big_datapoint<G>(c, dp);
    if (P.unit_clk && cta == 0 && threadIdx.x == 0) P.unit_clk[dp] = clock64() - t0;
  }
}

static cudaError_t big_cfg(cudaLaunchConfig_t* cfg, cudaLaunchAttribute* attr, size_t smem, int ncta, cudaStream_t stream) {
  cudaError_t e = cudaFuncSetAttribute(qk_sim_big_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  if (ncta > 8) {
    e = cudaFuncSetAttribute(qk_sim_big_kernel<256>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    if (e != cudaSuccess) return e;
  }
  *cfg = cudaLaunchConfig_t{};
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)ncta; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg->blockDim = dim3(256); cfg->dynamicSmemBytes = smem; cfg->stream = stream; cfg->attrs = attr; cfg->numAttrs = 1;
  cfg->gridDim = dim3((unsigned)ncta);
  return cudaSuccess;
}

cudaError_t qk_sim_big_config(size_t smem_bytes, int N, int ncta_req, int* ncta_out, int* n_clusters_out) {
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int ncta = ncta_req;
  if (ncta <= 0) {
    // as many CTAs per datapoint as the SMs allow: 148 SMs / N datapoints, a power of two <= 8.  (16-CTA clusters are
    // possible with the non-portable attribute, but only ~one fits per GPC: 8 datapoints then need two waves --
    // measured 8.8 s against 5.1 s with 8 CTAs on the C4 shape.)
    ncta = 1;
    while (ncta * 2 <= 8 && ncta * 2 * (N < 1 ? 1 : N) <= sms) ncta *= 2;
  }
  cudaError_t e = cudaSuccess;
  for (; ncta >= 1; ncta /= 2) {
    cudaLaunchConfig_t cfg; cudaLaunchAttribute attr[1];
    e = big_cfg(&cfg, attr, smem_bytes, ncta, nullptr);
    int max_clusters = 0;
    if (e == cudaSuccess) e = cudaOccupancyMaxActiveClusters(&max_clusters, qk_sim_big_kernel<256>, &cfg);
    if (e == cudaSuccess && max_clusters >= 1) {
      *ncta_out = ncta;
      *n_clusters_out = max_clusters < N ? max_clusters : (N < 1 ? 1 : N);
      return cudaSuccess;
    }
    cudaGetLastError();
    if (ncta == 1) break;
  }
  return e != cudaSuccess ? e : cudaErrorInvalidConfiguration;
}

cudaError_t qk_launch_sim_big(const SimParams& P, size_t smem_bytes, int ncta, int n_clusters, cudaStream_t stream) {
  cudaLaunchConfig_t cfg; cudaLaunchAttribute attr[1];
  cudaError_t e = big_cfg(&cfg, attr, smem_bytes, ncta, stream);
  if (e != cudaSuccess) return e;
  cfg.gridDim = dim3((unsigned)(n_clusters * ncta));
  return cudaLaunchKernelEx(&cfg, qk_sim_big_kernel<256>, P);
}
