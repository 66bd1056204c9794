// Micro-benchmark: does the packed FP32 FMA of sm_100 (fma.rn.f32x2, SASS FFMA2) free issue slots?
// 16 FMAs per iteration as 16 FFMA or 8 FFMA2, with 0 / 8 / 16 independent integer ALU instructions interleaved.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2 ffma2.cu && ./ffma2
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void fma2(float& d0, float& d1, float a0, float a1, float b0, float b1) {
  unsigned long long d, a, b;
  asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(d0), "f"(d1));
  asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "f"(a0), "f"(a1));
  asm("mov.b64 %0, {%1, %2};" : "=l"(b) : "f"(b0), "f"(b1));
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(a), "l"(b));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(d0), "=f"(d1) : "l"(d));
}

template <int PACKED, int ALU>
__global__ void __launch_bounds__(256, 2) k(float* out, int iters, float seed) {
  float acc[16], x[4], y[4];
  unsigned int z[16];
  for (int i = 0; i < 16; ++i) { acc[i] = seed * i; z[i] = threadIdx.x * 2654435761u + i; }
  for (int i = 0; i < 4; ++i) { x[i] = seed + i + threadIdx.x; y[i] = seed * 0.5f + i; }
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int rep = 0; rep < 4; ++rep) {
      if (PACKED) {
#pragma unroll
        for (int r = 0; r < 4; r += 2)
#pragma unroll
          for (int c = 0; c < 4; ++c) fma2(acc[r * 4 + c], acc[(r + 1) * 4 + c], x[r], x[r + 1], y[c], y[c]);
      } else {
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
          for (int c = 0; c < 4; ++c) acc[r * 4 + c] = fmaf(x[r], y[c], acc[r * 4 + c]);
      }
#pragma unroll
      for (int q = 0; q < ALU; ++q) z[q] = (z[q] ^ (z[(q + 1) & 15] >> 3)) + 0x9e3779b9u;     // LOP3/SHF/IADD mix
    }
  }
  float s = 0.f;
  unsigned int zz = 0;
  for (int i = 0; i < 16; ++i) { s += acc[i]; zz ^= z[i]; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + (float) zz;
}

template <int PACKED, int ALU>
void run(const char* name, float* out) {
  int dev_sms = 148;
  cudaDeviceGetAttribute(&dev_sms, cudaDevAttrMultiProcessorCount, 0);
  const int iters = 20000, grid = dev_sms * 2;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  k<PACKED, ALU><<<grid, 256>>>(out, 100, 1.0f);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  k<PACKED, ALU><<<grid, 256>>>(out, iters, 1.0f);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  const double fma = (double) grid * 256 * iters * 4 * 16;
  printf("%-28s %8.3f ms  %7.2f TFLOP/s  (%s)\n", name, ms, 2.0 * fma / ms / 1e9, cudaGetErrorString(cudaGetLastError()));
}

int main() {
  float* out;
  cudaMalloc(&out, 148 * 2 * 256 * 4 * 4);
  run<0, 0>("16 FFMA", out);
  run<1, 0>("8 FFMA2", out);
  run<0, 4>("16 FFMA + 4 ALU stmts", out);
  run<1, 4>("8 FFMA2 + 4 ALU stmts", out);
  run<0, 8>("16 FFMA + 8 ALU stmts", out);
  run<1, 8>("8 FFMA2 + 8 ALU stmts", out);
  run<0, 16>("16 FFMA + 16 ALU stmts", out);
  run<1, 16>("8 FFMA2 + 16 ALU stmts", out);
  return 0;
}
