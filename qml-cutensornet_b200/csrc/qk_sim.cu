// Stage 1 kernel: persistent grid, one cooperative group (= one CTA of G threads) per datapoint,
// interpreting the static op schedule compiled by qk_plan.cpp.  The algorithm itself is in
// qk_sim_core.h (shared with the test-only host emulation).
//
// B200 mapping: grid = SMs x resident CTAs; each CTA keeps theta / rotation matrices in shared
// memory for the whole SVD and streams the two site tensors of an op from its own state slot, which
// stays L2-resident (working set = resident CTAs x one MPS, far below the 126 MB L2); HBM sees the
// final state once.  Datapoints are handed out through an atomic counter so ragged bond dimensions
// do not leave SMs idle.
#include "qk_kernels.cuh"
#include "qk_sim_core.h"

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
    qk_sim_datapoint<G>(c, dp);
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
