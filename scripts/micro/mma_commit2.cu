// Micro-benchmark: what costs time between groups of tcgen05.mma (uniform fast issue as in the scan kernels):
// nothing / tcgen05.commit / tcgen05.fence::after_thread_sync / a completed-mbarrier wait.
#define DCB_GEMM_KERNELS
#include "gemm_kernels.cuh"
#include <cstdio>
#include <cstdlib>
using namespace dcb;

__global__ void k(int n_groups, int per_group, int mode, long long* out) {
  extern __shared__ unsigned char raw[];
  __shared__ uint64_t bars[4], fin, ready;
  __shared__ uint32_t taddr;
  unsigned char* base = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
  float* a = reinterpret_cast<float*>(base);
  float* b = a + 4 * G_CHUNK_FLOATS;
  for (int i = threadIdx.x; i < G_CHUNK_FLOATS * 8; i += blockDim.x) a[i] = (float) ((i * 7) % 13) * 0.125f;
  if (threadIdx.x == 0) {
    for (int q = 0; q < 4; ++q) mbar_init(&bars[q], 1);
    mbar_init(&fin, 1);
    mbar_init(&ready, 1);
    fence_mbar_init();
  }
  __syncthreads();
  const int warp = uniform_warp();
  if (warp == 1) tmem_alloc(&taddr, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = taddr;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (threadIdx.x == 0) mbar_arrive(&ready);       // phase 0 of `ready` is complete for the whole run
  __syncthreads();
  if (warp == 1) {
    const bool leader = elect_one();
    const uint64_t a0 = g_smem_desc(smem_u32(a)), b0 = g_smem_desc(smem_u32(b));
    const long long t0 = clock64();
    uint32_t i = 0;
    for (int gidx = 0; gidx < n_groups; ++gidx) {
      if (mode & 4) mbar_wait(&ready, 0);
      if (mode & 2) tc_fence_after();
      if (leader) {
        for (int kk = 0; kk < per_group; ++kk, ++i) {
          const uint64_t ad = a0 + (uint64_t) ((i >> 2) & 3) * 1024 + 2 * (i & 3), bd = b0 + (uint64_t) ((i >> 2) & 3) * 1024 + 2 * (i & 3);
          tc_mma_tf32(tb + (uint32_t) ((gidx & 1) * 128), ad, bd, G_IDESC, kk ? 1u : 0u);
        }
        if (mode & 1) tc_commit(&bars[gidx & 3]);
      }
      i += 0;
    }
    if (leader) { tc_commit(&fin); mbar_wait(&fin, 0); }
    const long long t1 = clock64();
    if (blockIdx.x == 0 && leader) out[0] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tb, 512);
}

int main() {
  long long* dout;
  cudaMalloc(&dout, 32);
  const size_t smem = 1024 + (size_t) G_CHUNK_BYTES * 9;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
  const int total = 8192;
  const char* names[8] = {"nothing", "commit", "fence", "commit+fence", "wait", "wait+commit", "wait+fence", "wait+commit+fence"};
  for (int per_group : {4, 8, 16, 32})
    for (int mode = 0; mode < 8; ++mode) {
      k<<<148, 64, smem>>>(total / per_group, per_group, mode, dout);
      cudaError_t e = cudaDeviceSynchronize();
      long long h = 0;
      cudaMemcpy(&h, dout, 8, cudaMemcpyDeviceToHost);
      printf("%2d MMAs per group, between groups: %-18s %7.1f cycles per MMA, %7.1f per group (%s)\n", per_group, names[mode], (double) h / total,
             (double) h / (total / per_group), cudaGetErrorString(e));
      fflush(stdout);
    }
  return 0;
}
