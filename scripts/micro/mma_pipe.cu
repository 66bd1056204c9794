// Micro-benchmark: (1) tcgen05.mma groups of 4 with a tcgen05.commit after each group; (2) the same fed by a TMA producer
// through a ring of 16 KB chunks (commit frees the slot), no epilogue.
#define DCB_GEMM_KERNELS
#include "gemm_kernels.cuh"
#include <cstdio>
#include <cstdlib>
using namespace dcb;

__global__ void pipe_kernel(int n_chunks, int mode, int n_stages, const float* src, long long* out) {
  extern __shared__ unsigned char raw[];
  __shared__ uint64_t full[8], empty[8], fin;
  __shared__ uint32_t taddr;
  unsigned char* base = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
  float* a = reinterpret_cast<float*>(base);
  float* ring = a + 4 * G_CHUNK_FLOATS;
  for (int i = threadIdx.x; i < G_CHUNK_FLOATS * (4 + n_stages); i += blockDim.x) a[i] = (float) ((i * 7) % 13) * 0.125f;
  if (threadIdx.x == 0) {
    for (int q = 0; q < 8; ++q) { mbar_init(&full[q], 1); mbar_init(&empty[q], 1); }
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
  if (warp == 0) {
    if (mode >= 1 && lane == 0) {       // TMA producer
      uint32_t stage = 0, phase = 0;
      for (int c = 0; c < n_chunks; ++c) {
        mbar_wait(&empty[stage], phase ^ 1u);
        mbar_arrive_expect_tx(&full[stage], G_CHUNK_BYTES);
        tma_load_1d(ring + (size_t) stage * G_CHUNK_FLOATS, src + ((size_t) (blockIdx.x * 97 + c) % 4096) * G_CHUNK_FLOATS, G_CHUNK_BYTES, &full[stage]);
        if (++stage == (uint32_t) n_stages) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    uint32_t stage = 0, phase = 0;
    const long long t0 = clock64();
    for (int c = 0; c < n_chunks; ++c) {
      if (mode >= 1) { mbar_wait(&full[stage], phase); tc_fence_after(); }
      if (lane == 0) {
        const uint64_t ad = g_smem_desc(smem_u32(a + (size_t) (c & 3) * G_CHUNK_FLOATS));
        const uint64_t bd = g_smem_desc(smem_u32(ring + (size_t) stage * G_CHUNK_FLOATS));
        for (int k = 0; k < 4; ++k) tc_mma_tf32(tb + (uint32_t) (((c >> 2) & 1) * 128), ad + 2 * k, bd + 2 * k, G_IDESC, ((c & 3) | k) ? 1u : 0u);
        if (mode != 3) tc_commit(&empty[stage]);
      }
      __syncwarp();
      if (++stage == (uint32_t) n_stages) { stage = 0; phase ^= 1u; }
    }
    if (lane == 0) { tc_commit(&fin); mbar_wait(&fin, 0); }
    const long long t1 = clock64();
    if (blockIdx.x == 0 && lane == 0) out[0] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tb, 512);
}

int main(int argc, char** argv) {
  long long* dout;
  float* src;
  cudaMalloc(&dout, 32);
  cudaMalloc(&src, (size_t) 4096 * G_CHUNK_BYTES);
  cudaMemset(src, 0, (size_t) 4096 * G_CHUNK_BYTES);
  const int n_chunks = 8192;
  const size_t smem = 1024 + (size_t) G_CHUNK_BYTES * 12;
  cudaFuncSetAttribute(pipe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
  const char* names[4] = {"resident operands, commit per 4 MMAs", "TMA ring, commit frees the slot", "TMA ring (same)", "resident operands, no commit"};
  const int mode = argc > 1 ? atoi(argv[1]) : 0, stages = argc > 2 ? atoi(argv[2]) : 8;
  for (int grid : {1, 148}) {
    pipe_kernel<<<grid, 64, smem>>>(n_chunks, mode, stages, src, dout);
    cudaError_t e = cudaDeviceSynchronize();
    long long h = 0;
    cudaMemcpy(&h, dout, 8, cudaMemcpyDeviceToHost);
    printf("%-40s stages %d grid %3d: %7.1f cycles per chunk of 4 MMAs (%s)\n", names[mode], stages, grid, (double) h / n_chunks, cudaGetErrorString(e));
    fflush(stdout);
  }
  return 0;
}
