// Micro-benchmark: cost of tcgen05.commit as a function of the MMA work between commits (resident operands).
#define DCB_GEMM_KERNELS
#include "gemm_kernels.cuh"
#include <cstdio>
#include <cstdlib>
using namespace dcb;

template <int N>
__global__ void commit_kernel(int n_groups, int per_group, int n_commits, long long* out) {
  extern __shared__ unsigned char raw[];
  __shared__ uint64_t bars[4], fin;
  __shared__ uint32_t taddr;
  unsigned char* base = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
  float* a = reinterpret_cast<float*>(base);
  float* b = a + 4 * G_CHUNK_FLOATS;
  for (int i = threadIdx.x; i < G_CHUNK_FLOATS * (4 + 4 * N / 128); i += blockDim.x) a[i] = (float) ((i * 7) % 13) * 0.125f;
  if (threadIdx.x == 0) {
    for (int q = 0; q < 4; ++q) mbar_init(&bars[q], 1);
    mbar_init(&fin, 1);
    fence_mbar_init();
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 1) tmem_alloc(&taddr, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = taddr;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t) (N >> 3) << 17) | ((uint32_t) (128 >> 4) << 24);
  if (warp == 1 && lane == 0) {
    const long long t0 = clock64();
    int i = 0;
    for (int gidx = 0; gidx < n_groups; ++gidx) {
      for (int k = 0; k < per_group; ++k, ++i) {
        const uint64_t ad = g_smem_desc(smem_u32(a + (size_t) ((i >> 2) & 3) * G_CHUNK_FLOATS));
        const uint64_t bd = g_smem_desc(smem_u32(b + (size_t) ((i >> 2) & 3) * (N / 128) * G_CHUNK_FLOATS));
        tc_mma_tf32(tb + (uint32_t) ((gidx & 1) * N), ad + 2 * (i & 3), bd + 2 * (i & 3), idesc, k ? 1u : 0u);
      }
      for (int q = 0; q < n_commits; ++q) tc_commit(&bars[q]);
    }
    tc_commit(&fin);
    mbar_wait(&fin, 0);
    const long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tb, 512);
}

int main(int argc, char** argv) {
  long long* dout;
  cudaMalloc(&dout, 32);
  const size_t smem = 1024 + (size_t) G_CHUNK_BYTES * 12;
  cudaFuncSetAttribute(commit_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
  cudaFuncSetAttribute(commit_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
  const int total = 16384;
  for (int n : {128, 256})
    for (int per_group : {1, 2, 4, 8, 16, 32})
      for (int n_commits : {1, 2}) {
        if (n == 128) commit_kernel<128><<<148, 64, smem>>>(total / per_group, per_group, n_commits, dout);
        else commit_kernel<256><<<148, 64, smem>>>(total / per_group, per_group, n_commits, dout);
        cudaError_t e = cudaDeviceSynchronize();
        long long h = 0;
        cudaMemcpy(&h, dout, 8, cudaMemcpyDeviceToHost);
        printf("N %3d  %2d MMAs per group, %d commits: %7.1f cycles per MMA, %7.1f per group (%s)\n", n, per_group, n_commits, (double) h / total,
               (double) h / (total / per_group), cudaGetErrorString(e));
        fflush(stdout);
      }
  return 0;
}
