// Micro-benchmark: does the operand data change the tcgen05.mma rate?  (zeros / small integers / random TF32 values)
#define DCB_GEMM_KERNELS
#include "gemm_kernels.cuh"
#include <cstdio>
using namespace dcb;

__global__ void k(int total, int data, long long* out) {
  extern __shared__ unsigned char raw[];
  __shared__ uint64_t fin;
  __shared__ uint32_t taddr;
  unsigned char* base = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
  float* a = reinterpret_cast<float*>(base);
  float* b = a + 4 * G_CHUNK_FLOATS;
  for (int i = threadIdx.x; i < G_CHUNK_FLOATS * 8; i += blockDim.x) {
    uint32_t h = (uint32_t) i * 2654435761u + blockIdx.x * 40503u;
    h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
    float v = 0.f;
    if (data == 1) v = (float) ((i * 7) % 13) * 0.125f;
    if (data == 2) v = to_tf32(((float) (h & 0xffffff) / 16777216.0f - 0.5f) * 2.0f);
    if (data == 3) v = ((float) (h & 0xffffff) / 16777216.0f - 0.5f) * 2.0f;      // full FP32 mantissa (hardware truncates)
    a[i] = v;
  }
  if (threadIdx.x == 0) { mbar_init(&fin, 1); fence_mbar_init(); }
  __syncthreads();
  const int warp = uniform_warp();
  if (warp == 1) tmem_alloc(&taddr, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = taddr;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 1) {
    const bool leader = elect_one();
    const uint64_t a0 = g_smem_desc(smem_u32(a)), b0 = g_smem_desc(smem_u32(b));
    const long long t0 = clock64();
    if (leader) {
      for (uint32_t i = 0; i < (uint32_t) total; ++i) {
        const uint64_t ad = a0 + (uint64_t) ((i >> 2) & 3) * 1024 + 2 * (i & 3), bd = b0 + (uint64_t) ((i >> 2) & 3) * 1024 + 2 * (i & 3);
        tc_mma_tf32(tb + ((i >> 4) & 1) * 128u, ad, bd, G_IDESC, (i & 15) ? 1u : 0u);
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
  const size_t smem = 1024 + (size_t) G_CHUNK_BYTES * 9;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
  const char* names[4] = {"zeros", "small multiples of 1/8", "random TF32", "random FP32"};
  for (int rep = 0; rep < 2; ++rep)
    for (int data = 0; data < 4; ++data) {
      cudaEvent_t e0, e1;
      cudaEventCreate(&e0); cudaEventCreate(&e1);
      const int total = 1 << 17;
      cudaEventRecord(e0);
      k<<<148, 64, smem>>>(total, data, dout);
      cudaEventRecord(e1);
      cudaError_t e = cudaDeviceSynchronize();
      float ms = 0;
      cudaEventElapsedTime(&ms, e0, e1);
      long long h = 0;
      cudaMemcpy(&h, dout, 8, cudaMemcpyDeviceToHost);
      printf("%-24s %7.1f cycles per MMA, %.3f ms -> %.0f TFLOP/s, %.0f MHz (%s)\n", names[data], (double) h / total, ms,
             148.0 * total * 262144.0 / (ms * 1e-3) / 1e12, (double) h / (ms * 1e3), cudaGetErrorString(e));
      fflush(stdout);
    }
  return 0;
}
