// Micro-benchmark: cycles per tcgen05.mma (tf32, 128x128x8, SS) while other warps (a) read TMEM with tcgen05.ld,
// (b) stream 16 KB bulk copies into shared memory.  Isolates what slows the MMA pipe down inside the scan kernels.
#define DCB_GEMM_KERNELS
#include "gemm_kernels.cuh"
#include <cstdio>
using namespace dcb;

__global__ void contend_kernel(int iters, int do_ld, int do_tma, int ld_same_cols, const float* src, long long* out, float* sink) {
  extern __shared__ unsigned char raw[];
  __shared__ uint64_t bar, tbar[4];
  __shared__ uint32_t taddr;
  __shared__ volatile int done;
  unsigned char* base = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
  float* a = reinterpret_cast<float*>(base);
  float* b = a + G_CHUNK_FLOATS;              // 4 chunks
  float* t = b + 4 * G_CHUNK_FLOATS;          // 4 chunks, TMA targets
  for (int i = threadIdx.x; i < G_CHUNK_FLOATS * 9; i += blockDim.x) a[i] = (float) ((i * 7) % 13) * 0.125f;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    for (int q = 0; q < 4; ++q) mbar_init(&tbar[q], 1);
    fence_mbar_init();
    done = 0;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) tmem_alloc(&taddr, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = taddr;
  if (warp == 0) {
    if (lane == 0) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      const uint64_t ad = g_smem_desc(smem_u32(a));
      const long long t0 = clock64();
      for (int i = 0; i < iters; ++i) {
        const uint64_t bd = g_smem_desc(smem_u32(b + (size_t) (i & 3) * G_CHUNK_FLOATS));
        tc_mma_tf32(tb + (uint32_t) (((i >> 4) & 1) * 128), ad + 2 * (i & 3), bd + 2 * (i & 3), G_IDESC, (i & 15) ? 1u : 0u);
      }
      tc_commit(&bar);
      mbar_wait(&bar, 0);
      const long long t1 = clock64();
      if (blockIdx.x == 0) out[0] = t1 - t0;
      done = 1;
    }
  } else if (warp == 1) {
    if (do_tma && lane == 0) {
      uint32_t ph = 0;
      int q = 0;
      long long n = 0;
      // keep 4 bulk copies in flight
      for (int k = 0; k < 4; ++k) { mbar_arrive_expect_tx(&tbar[k], G_CHUNK_BYTES); tma_load_1d(t + k * G_CHUNK_FLOATS, src + ((size_t) (blockIdx.x * 64 + k) % 4096) * G_CHUNK_FLOATS, G_CHUNK_BYTES, &tbar[k]); }
      while (!done) {
        mbar_wait(&tbar[q], ph);
        mbar_arrive_expect_tx(&tbar[q], G_CHUNK_BYTES);
        tma_load_1d(t + q * G_CHUNK_FLOATS, src + ((size_t) (blockIdx.x * 64 + n) % 4096) * G_CHUNK_FLOATS, G_CHUNK_BYTES, &tbar[q]);
        ++n;
        if (++q == 4) { q = 0; ph ^= 1; }
      }
      for (int k = 0; k < 4; ++k) { mbar_wait(&tbar[q], ph); if (++q == 4) { q = 0; ph ^= 1; } }
      if (blockIdx.x == 0) out[1] = n;
    }
  } else if (warp >= 4) {
    if (do_ld) {
      float acc = 0.f;
      long long n = 0;
      const uint32_t quarter = warp & 3;
      while (!done) {
        float v[32];
        tmem_ld32(tb + ((quarter * 32u) << 16) + (ld_same_cols ? 0u : 256u) + (uint32_t) ((n & 3) * 32), v);
#pragma unroll
        for (int q = 0; q < 32; ++q) acc += v[q];
        ++n;
      }
      if (acc == 123.f) sink[0] = acc;
      if (blockIdx.x == 0 && threadIdx.x == 128) out[2] = n;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tb, 512);
}

int main() {
  long long* dout;
  float *src, *sink;
  cudaMalloc(&dout, 32);
  cudaMalloc(&sink, 4);
  cudaMalloc(&src, (size_t) 4096 * G_CHUNK_BYTES);
  cudaMemset(src, 0, (size_t) 4096 * G_CHUNK_BYTES);
  const int iters = 16384;
  const size_t smem = 1024 + (size_t) G_CHUNK_BYTES * 9;
  cudaFuncSetAttribute(contend_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
  for (int nwarps : {8, 12}) {
    for (int mode = 0; mode < 6; ++mode) {
      const int do_ld = (mode == 1 || mode == 3 || mode == 4 || mode == 5), do_tma = (mode == 2 || mode == 3 || mode == 5), same = (mode >= 4);
      cudaMemset(dout, 0, 32);
      contend_kernel<<<148, nwarps * 32, smem>>>(iters, do_ld, do_tma, same, src, dout, sink);
      cudaError_t e = cudaDeviceSynchronize();
      long long h[4];
      cudaMemcpy(h, dout, 32, cudaMemcpyDeviceToHost);
      printf("warps %2d  ld %d (same cols %d)  tma %d : %7.1f cycles per MMA, %6.1f bulk B/cycle, %6.1f tmem-ld B/cycle/warp (%s)\n", nwarps, do_ld, same,
             do_tma, (double) h[0] / iters, (double) h[1] * G_CHUNK_BYTES / (double) h[0], (double) h[2] * 32 * 128 / (double) h[0], cudaGetErrorString(e));
    }
  }
  return 0;
}
