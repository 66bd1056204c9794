// Micro-benchmark: does switching the TMEM accumulator between consecutive tcgen05.mma, or a non-accumulating MMA, cost time?
#define DCB_GEMM_KERNELS
#include "gemm_kernels.cuh"
#include <cstdio>
using namespace dcb;

template <int N>
__global__ void k(int total, int period, int mode, long long* out) {
  extern __shared__ unsigned char raw[];
  __shared__ uint64_t fin;
  __shared__ uint32_t taddr;
  unsigned char* base = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
  float* a = reinterpret_cast<float*>(base);
  float* b = a + 4 * G_CHUNK_FLOATS;
  for (int i = threadIdx.x; i < G_CHUNK_FLOATS * 12; i += blockDim.x) a[i] = (float) ((i * 7) % 13) * 0.125f;
  if (threadIdx.x == 0) { mbar_init(&fin, 1); fence_mbar_init(); }
  __syncthreads();
  const int warp = uniform_warp();
  if (warp == 1) tmem_alloc(&taddr, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = taddr;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t) (N >> 3) << 17) | ((uint32_t) (128 >> 4) << 24);
  if (warp == 1) {
    const bool leader = elect_one();
    const uint64_t a0 = g_smem_desc(smem_u32(a)), b0 = g_smem_desc(smem_u32(b));
    const long long t0 = clock64();
    if (leader) {
      for (uint32_t i = 0; i < (uint32_t) total; ++i) {
        const uint64_t ad = a0 + (uint64_t) ((i >> 2) & 3) * 1024 + 2 * (i & 3), bd = b0 + (uint64_t) ((i >> 2) & 3) * 1024 * (N / 128) + 2 * (i & 3);
        const uint32_t grp = i / (uint32_t) period;
        const uint32_t d = (mode & 1) ? tb + (grp & 1) * (uint32_t) N : tb;          // bit 0: switch the accumulator every `period` MMAs
        const uint32_t acc = (mode & 2) ? ((i % (uint32_t) period) ? 1u : 0u) : 1u;  // bit 1: first MMA of a period overwrites
        tc_mma_tf32(d, ad, bd, idesc, acc);
      }
      tc_commit(&fin);
      mbar_wait(&fin, 0);
    }
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
  const size_t smem = 1024 + (size_t) G_CHUNK_BYTES * 13;
  cudaFuncSetAttribute(k<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
  cudaFuncSetAttribute(k<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
  const int total = 8192;
  const char* names[4] = {"same D, always accumulate", "switch D, always accumulate", "same D, overwrite at period start", "switch D + overwrite"};
  for (int n : {128, 256})
    for (int period : {1, 4, 16})
      for (int mode = 0; mode < 4; ++mode) {
        if (n == 128) k<128><<<148, 64, smem>>>(total, period, mode, dout); else k<256><<<148, 64, smem>>>(total, period, mode, dout);
        cudaError_t e = cudaDeviceSynchronize();
        long long h = 0;
        cudaMemcpy(&h, dout, 8, cudaMemcpyDeviceToHost);
        printf("N %3d period %2d  %-36s %7.1f cycles per MMA (%s)\n", n, period, names[mode], (double) h / total, cudaGetErrorString(e));
        fflush(stdout);
      }
  return 0;
}
