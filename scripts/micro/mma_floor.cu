// Micro-benchmark: cycles per tcgen05.mma (kind::tf32, SS mode, M = 128) issued back to back on resident operands.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I clustering_b200/csrc scripts/micro/mma_floor.cu -o gpurun_out/mma_floor
#define DCB_GEMM_KERNELS
#include "gemm_kernels.cuh"
#include <cstdio>
using namespace dcb;

template <int N, int NB>   // NB: distinct B buffers cycled through (1: same operand every time)
__global__ void floor_kernel(int iters, long long* out) {
  extern __shared__ unsigned char raw[];
  __shared__ uint64_t bar;
  __shared__ uint32_t taddr;
  unsigned char* base = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
  float* a = reinterpret_cast<float*>(base);
  float* b = a + G_CHUNK_FLOATS;
  for (int i = threadIdx.x; i < G_CHUNK_FLOATS * (1 + NB * N / 128); i += blockDim.x) a[i] = 0.f;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  __syncthreads();
  if (threadIdx.x < 32) tmem_alloc(&taddr, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = taddr;
  constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t) (N >> 3) << 17) | ((uint32_t) (128 >> 4) << 24);
  if (threadIdx.x == 0) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    const uint64_t ad = g_smem_desc(smem_u32(a));
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      const uint64_t bd = g_smem_desc(smem_u32(b + (size_t) (i % NB) * (N / 128) * G_CHUNK_FLOATS));
      tc_mma_tf32(tb + (uint32_t) ((i & 1) * N), ad + 2 * (i & 3), bd + 2 * (i & 3), idesc, 1u);
    }
    tc_commit(&bar);
    mbar_wait(&bar, 0);
    const long long t1 = clock64();
    if (blockIdx.x == 0) out[0] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tb, 512);
}

template <int N, int NB>
void run(const char* name, long long* dout) {
  const int iters = 4096;
  const size_t smem = 1024 + (size_t) G_CHUNK_BYTES * (1 + NB * N / 128);
  cudaFuncSetAttribute(floor_kernel<N, NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
  for (int grid : {1, 148}) {
    floor_kernel<N, NB><<<grid, 64, smem>>>(iters, dout);
    cudaError_t e = cudaDeviceSynchronize();
    long long h = 0;
    cudaMemcpy(&h, dout, 8, cudaMemcpyDeviceToHost);
    printf("%-28s grid %3d: %7.1f cycles per MMA (%s)\n", name, grid, (double) h / iters, cudaGetErrorString(e));
  }
}

int main() {
  long long* dout;
  cudaMalloc(&dout, 8);
  run<128, 1>("tf32 128x128x8, 1 B buffer", dout);
  run<128, 4>("tf32 128x128x8, 4 B buffers", dout);
  run<256, 1>("tf32 128x256x8, 1 B buffer", dout);
  run<256, 2>("tf32 128x256x8, 2 B buffers", dout);
  run<64, 4>("tf32 128x64x8, 4 B buffers", dout);
  return 0;
}
