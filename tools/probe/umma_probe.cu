// Hardware probe (development tool, not product): run ONE tcgen05.mma (kind::f16, M = 128, N = 32, K = 16, fp32
// accumulate) over a caller-supplied shared-memory image with caller-supplied operand / instruction descriptors and
// return the accumulator.  With one operand an identity matrix the accumulator is a direct read-out of WHICH
// shared-memory elements the tensor core fetched for every (row, k) of the other operand, so descriptor semantics
// (swizzle phase of unaligned start addresses, MN-major LBO/SBO, 32-byte swizzle) are measured instead of assumed.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -shared -Xcompiler -fPIC -o libprobe.so umma_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../paac_b200/csrc/tc_ptx.cuh"

using namespace paacb;

__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// image: `bytes` bytes copied to the 1024-byte aligned shared-memory base.  The descriptors carry start addresses
// RELATIVE to that base (in their 14-bit >>4 field); the kernel adds the base.  nmma MMAs are issued: MMA i uses
// adesc + i * a_step, bdesc + i * b_step (steps added to the 64-bit descriptors, i.e. to the >>4 address field).
__global__ void __launch_bounds__(128, 1) probe_kernel(const uint8_t* __restrict__ image, int bytes, uint64_t adesc,
                                                        uint64_t bdesc, uint32_t idesc, int nmma, uint32_t a_step,
                                                        uint32_t b_step, float* __restrict__ out, uint32_t* base_out) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid * 16; i < bytes; i += 128 * 16)
    *reinterpret_cast<uint4*>(smem + i) = *reinterpret_cast<const uint4*>(image + i);
  if (tid == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
    *base_out = smem_u32(smem);
  }
  if (warp == 0) tmem_alloc(&tmem_slot, 32);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  if (warp == 0) {
    if (elect_one_sync()) {
      const uint64_t base16 = (uint64_t)(smem_u32(smem) >> 4);
      for (int i = 0; i < nmma; ++i) {
        const uint64_t a = adesc + base16 + (uint64_t)i * a_step;
        const uint64_t b = bdesc + base16 + (uint64_t)i * b_step;
        umma_f16(tmem_base, a, b, idesc, i > 0 ? 1u : 0u);
      }
      umma_commit(&bar);
    }
    __syncwarp();
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  uint32_t v[32];
  tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16), v);
  tmem_ld_wait();
  for (int j = 0; j < 32; ++j) out[(size_t)tid * 32 + j] = __uint_as_float(v[j]);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 32);
  }
}

extern "C" int probe_run(const void* image_host, int bytes, uint64_t adesc, uint64_t bdesc, uint32_t idesc, int nmma,
                         uint32_t a_step, uint32_t b_step, float* out_host /* [128][32] */) {
  static uint8_t* d_img = nullptr;
  static float* d_out = nullptr;
  static uint32_t* d_base = nullptr;
  const int kMax = 200 * 1024;
  if (bytes > kMax || (bytes & 15)) return -1;
  if (!d_img) {
    if (cudaMalloc(&d_img, kMax) != cudaSuccess) return -2;
    if (cudaMalloc(&d_out, 128 * 32 * 4) != cudaSuccess) return -2;
    if (cudaMalloc(&d_base, 4) != cudaSuccess) return -2;
    if (cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMax + 2048) != cudaSuccess) return -3;
  }
  cudaMemcpy(d_img, image_host, bytes, cudaMemcpyHostToDevice);
  cudaMemset(d_out, 0xff, 128 * 32 * 4);
  probe_kernel<<<1, 128, kMax + 2048>>>(d_img, bytes, adesc, bdesc, idesc, nmma, a_step, b_step, d_out, d_base);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    fprintf(stderr, "probe: %s\n", cudaGetErrorString(e));
    return -4;
  }
  cudaMemcpy(out_host, d_out, 128 * 32 * 4, cudaMemcpyDeviceToHost);
  return 0;
}
