// Measured tcgen05.mma issue peaks on this GPU, per kind and N: the tensor-core roofline denominators bench.py uses.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/probe/umma_peak tools/probe/umma_peak.cu
//   tools/probe/umma_peak > gpurun_out/umma_peaks.json
// One CTA per SM; one thread issues `iters` x 4 MMAs (M = 128, N, K = 32 bytes per instruction: 16 bf16 / 8 tf32 / 32 int8),
// SS mode (both operands from shared memory, SWIZZLE_128B K-major tiles filled with a non-trivial pattern), alternating
// between two TMEM accumulators; timed with CUDA events around the launch.  MEASURED_PEAKS.json's bf16 figure is a cuBLAS
// GEMM (cta_group::2, TMA multicast); this is what ONE CTA per SM can issue -- the ceiling of the kernels in paac_b200/csrc.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../paac_b200/csrc/tc_ptx.cuh"

using namespace paacb;

enum { KIND_F16 = 0, KIND_TF32 = 1, KIND_I8 = 2 };

template <int KIND>
__device__ __forceinline__ void umma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  if (KIND == KIND_F16)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
  else if (KIND == KIND_TF32)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
  else
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

template <int KIND>
__global__ void __launch_bounds__(128, 1) peak_kernel(int iters, int n, uint32_t idesc) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* a_sm = smem;                  // 128 rows x 128 B
  uint8_t* b_sm = smem + 16384;          // 256 rows x 128 B
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 16384 + 32768);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  // operand pattern: small finite values of the kind's element type (bf16 ~ +-1, tf32 ~ +-1, int8 +-3)
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) {
    uint32_t w;
    const uint32_t h = (uint32_t)i * 2654435761u;
    if (KIND == KIND_F16) w = 0x3F803F80u ^ ((h & 0x8000u) | ((h << 3) & 0x80000000u)) ^ ((h >> 9) & 0x007F007Fu);
    else if (KIND == KIND_TF32) w = 0x3F800000u ^ (h & 0x80000000u) ^ ((h >> 7) & 0x007FE000u);
    else w = (h & 0x03030303u) | ((h >> 3) & 0x80808080u & 0u);
    reinterpret_cast<uint32_t*>(smem)[i] = w;
  }
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_barrier_init();
  }
  if (threadIdx.x < 32) tmem_alloc(slot, 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  if (threadIdx.x == 0) {
    const uint64_t d0 = make_smem_desc(0, 16, 1024, SWZ_128B);
    const uint32_t a = smem_u32(a_sm), b = smem_u32(b_sm);
    for (int i = 0; i < iters; ++i) {
      const uint32_t d = tmem + (uint32_t)((i & 1) * 256);
#pragma unroll
      for (int ks = 0; ks < 4; ++ks)
        umma<KIND>(d, desc_with_addr(d0, a + ks * 32), desc_with_addr(d0, b + ks * 32), idesc, (i > 1 || ks > 0) ? 1u : 0u);
    }
    umma_commit(bar);
    mbar_wait(bar, 0);
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

static uint32_t idesc_for(int kind, int n) {
  if (kind == KIND_F16) return make_idesc_bf16(n, 0, 0);
  if (kind == KIND_TF32) return make_idesc_tf32(n);
  return (2u << 4) | (0u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);     // D s32, A u8, B s8
}

template <int KIND>
static double run(int sms, int n, int iters) {
  const int smem = 16384 + 32768 + 1024 + 64;
  cudaFuncSetAttribute(peak_kernel<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  peak_kernel<KIND><<<sms, 128, smem>>>(iters / 8, n, idesc_for(KIND, n));      // warm-up
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(e0);
    peak_kernel<KIND><<<sms, 128, smem>>>(iters, n, idesc_for(KIND, n));
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  if (cudaGetLastError() != cudaSuccess) return -1.0;
  const int kelems = KIND == KIND_F16 ? 16 : (KIND == KIND_TF32 ? 8 : 32);
  const double ops = 2.0 * 128.0 * n * kelems * 4.0 * iters * sms;
  return ops / (best * 1e-3) / 1e12;
}

int main() {
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, 0) != cudaSuccess) { printf("{\"error\": \"no device\"}\n"); return 1; }
  const int sms = prop.multiProcessorCount;
  const int ns[] = {256, 128, 96, 64, 32};
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"how\": \"one CTA per SM, one thread issues tcgen05.mma cta_group::1 M=128 SS-mode back to back on two TMEM accumulators; best of 5 launches (CUDA events); T(FL)OP/s = 2*128*N*K per instruction\",\n \"peaks\": {", prop.name, sms);
  const char* names[] = {"f16_bf16", "tf32", "i8"};
  for (int kind = 0; kind < 3; ++kind) {
    printf("%s\"%s\": {", kind ? ", " : "", names[kind]);
    for (int j = 0; j < 5; ++j) {
      const int n = ns[j];
      const int iters = 16000;
      double t = kind == 0 ? run<KIND_F16>(sms, n, iters) : (kind == 1 ? run<KIND_TF32>(sms, n, iters) : run<KIND_I8>(sms, n, iters));
      printf("%s\"N%d\": %.1f", j ? ", " : "", n, t);
    }
    printf("}");
  }
  printf("}}\n");
  return 0;
}
