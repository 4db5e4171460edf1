// Tensor-pipe micro-benchmark: cycles per tcgen05.mma (kind::f16, bf16) for the operand patterns the GEMM kernels use.
// Operands sit in shared memory (no TMA traffic), one thread per CTA pair issues R MMAs and waits for the commit.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I multimodal-s2ut_b200/csrc -o /tmp/umma_bench profiles/tools/umma_bench.cu
#include "common.cuh"
using namespace mm;

constexpr int A_BYTES = 128 * 64 * 2, B_BYTES = 128 * 64 * 2;   // per CTA per k-block (64 K)
constexpr int NSTAGE = 4;
constexpr int SMEM = NSTAGE * (A_BYTES + 2 * B_BYTES) + 1024 + 256;

// pattern: 0 = one accumulator, operands cycle over NSTAGE stages (gemm.cu);  1 = two accumulators sharing A
// (fused GEMM+LN main loop);  2 = one accumulator, N = 128;  3 = two accumulators sharing B (different A stage)
template <int PATTERN, int CG>
__global__ void __launch_bounds__(128, 1) umma_bench(long long* out, int reps) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + NSTAGE * (A_BYTES + 2 * B_BYTES));
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = CG == 2 ? cluster_ctarank() : 0;
  for (int i = threadIdx.x; i < NSTAGE * (A_BYTES + 2 * B_BYTES) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_barrier_init();
  }
  fence_proxy_async_smem();
  if (warp == 0) {
    if (CG == 2) tmem_alloc_2sm(slot, 512); else tmem_alloc(slot, 512);
  }
  tc_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  if (warp == 1 && lane == 0 && rank == 0) {
    constexpr int N = PATTERN == 2 ? 128 : 256;
    constexpr uint32_t idesc = umma_idesc(CG == 2 ? 256 : 128, N, 1);
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      const int st = r % NSTAGE;
      uint8_t* base = smem + st * (A_BYTES + 2 * B_BYTES);
      const uint64_t a = umma_desc_sw128(smem_u32(base));
      const uint64_t b0 = umma_desc_sw128(smem_u32(base + A_BYTES));
      const uint64_t b1 = umma_desc_sw128(smem_u32(base + A_BYTES + B_BYTES));
      const uint64_t a2 = umma_desc_sw128(smem_u32(smem + ((st + 1) % NSTAGE) * (A_BYTES + 2 * B_BYTES)));
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        if (CG == 2) {
          umma_f16_2sm(tmem, a + 2 * kk, b0 + 2 * kk, idesc, 1);
          if (PATTERN == 1) umma_f16_2sm(tmem + 256, a + 2 * kk, b1 + 2 * kk, idesc, 1);
          if (PATTERN == 3) umma_f16_2sm(tmem + 256, a2 + 2 * kk, b0 + 2 * kk, idesc, 1);
        } else {
          umma_f16(tmem, a + 2 * kk, b0 + 2 * kk, idesc, 1);
          if (PATTERN == 1) umma_f16(tmem + 256, a + 2 * kk, b1 + 2 * kk, idesc, 1);
          if (PATTERN == 3) umma_f16(tmem + 256, a2 + 2 * kk, b0 + 2 * kk, idesc, 1);
        }
      }
    }
    if (CG == 2) umma_commit_2sm(bar, 1); else umma_commit(bar);
    mbar_wait(bar, 0);
    const long long t1 = clock64();
    out[blockIdx.x / CG] = t1 - t0;
  }
  tc_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    if (CG == 2) tmem_dealloc_2sm(tmem, 512); else tmem_dealloc(tmem, 512);
  }
}

template <int PATTERN, int CG>
void run(const char* name, int ctas, int reps) {
  long long* d;
  cudaMalloc(&d, 148 * sizeof(long long));
  cudaMemset(d, 0, 148 * sizeof(long long));
  auto kern = umma_bench<PATTERN, CG>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(ctas);
  cfg.blockDim = dim3(128);
  cfg.dynamicSmemBytes = SMEM;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  for (int it = 0; it < 2; ++it) {
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, d, reps);
    if (e != cudaSuccess) { printf("%s: launch failed %s\n", name, cudaGetErrorString(e)); return; }
    e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: failed %s\n", name, cudaGetErrorString(e)); return; }
  }
  long long h[148];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  const int per_rep = 4 * ((PATTERN == 1 || PATTERN == 3) ? 2 : 1);
  double sum = 0;
  int n = ctas / CG;
  for (int i = 0; i < n; ++i) sum += (double)h[i];
  const double cyc = sum / n / ((double)reps * per_rep);
  const double flops = 2.0 * (CG == 2 ? 256 : 128) * (PATTERN == 2 ? 128 : 256) * 16;
  printf("%-58s %4d CTAs  %7.1f cycles/MMA  -> %6.0f flop/clk/SM\n", name, ctas, cyc, flops / cyc / CG);
  cudaFree(d);
}

int main() {
  const int reps = 4000;
  for (int ctas : {2, 148}) {
    run<0, 2>("2-CTA M256 N256, one accumulator (gemm.cu)", ctas, reps);
    run<1, 2>("2-CTA M256 N256, two accumulators sharing A (gemm_ln)", ctas, reps);
    run<3, 2>("2-CTA M256 N256, two accumulators sharing B", ctas, reps);
    run<2, 2>("2-CTA M256 N128, one accumulator", ctas, reps);
    run<0, 1>("1-CTA M128 N256, one accumulator", ctas, reps);

  }
  return 0;
}
