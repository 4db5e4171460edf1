// MUFU.EX2 issue rate on this GPU: cycles per warp-level ex2.approx.ftz.f32 with 1 / 2 / 4 warps per SM sub-partition
// (8 independent chains per thread, so latency is hidden and the pipe's throughput is what is measured), and the same
// for an FFMA-only loop as the control.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu_bench mufu_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
template <int MODE>
__global__ void k(float* out, long long* cyc, int iters) {
  float a[8];
  for (int j = 0; j < 8; ++j) a[j] = -0.001f * (threadIdx.x + j);
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (MODE == 0) a[j] = ex2(a[j]);                       // MUFU only
      if (MODE == 1) a[j] = fmaf(a[j], 0.999f, -0.001f);     // FFMA only
      if (MODE == 2) a[j] = ex2(fmaf(a[j], 0.999f, -0.001f)); // FFMA + MUFU per element
    }
  }
  const long long t1 = clock64();
  float s = 0; for (int j = 0; j < 8; ++j) s += a[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main() {
  float* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
  const int iters = 2048;
  const char* names[3] = {"MUFU.EX2", "FFMA", "FFMA+MUFU.EX2"};
  for (int mode = 0; mode < 3; ++mode)
    for (int threads : {128, 256, 512, 1024}) {
      for (int rep = 0; rep < 2; ++rep) {
        if (mode == 0) k<0><<<148, threads>>>(out, cyc, iters);
        if (mode == 1) k<1><<<148, threads>>>(out, cyc, iters);
        if (mode == 2) k<2><<<148, threads>>>(out, cyc, iters);
      }
      cudaDeviceSynchronize();
      long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
      const double per_instr = (double)h[0] / (iters * 8.0);           // cycles per warp-instruction of ONE warp
      const int wps = threads / 128;                                   // warps per sub-partition
      printf("%-14s %d warp(s)/SMSP: %.2f cycles per element-step per warp -> %.2f cycles of the SMSP per warp-instruction group\n",
             names[mode], wps, per_instr, per_instr / wps);
    }
  return 0;
}
