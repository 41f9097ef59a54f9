// tcgen05.mma.cta_group::2 probe: a CTA pair (cluster of 2) issues ONE MMA over both SMs -- M = 256 (128 rows from each
// CTA's shared memory), the B operand split between the two CTAs (each holds N / 2 of its rows), each CTA's accumulator in its
// own tensor memory.  What it answers before any product kernel is touched:
//   (1) functional: D of BOTH CTAs against a host reference for the operand placement assumed in DESIGN.md 3.3
//       (A_r = the 128 rows CTA r holds, B = [rows 0 .. N/2 of CTA 0 | rows N/2 .. N of CTA 1]);
//   (2) timing: SM cycles per K-step of the conv kernels' instruction mix (N = 128 then N = 64) and of the fc mix
//       (N = 256 then N = 128) -- in cta_group::1 they cost 113 and 192 cycles (profiles/r02_umma_mix.json); a CTA of the
//       pair reads only half of B from its shared memory.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/probe/umma_2cta tools/probe/umma_2cta.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <math.h>
#include <algorithm>
#include <vector>
#include "../../paac_b200/csrc/tc_ptx.cuh"

using namespace paacb;

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* slot, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma2_bf16(uint32_t d, uint64_t ad, uint64_t bd, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(ad), "l"(bd), "r"(idesc), "r"(acc)
      : "memory");
}
// arrive on the barrier at this shared-memory offset in every CTA of `mask` once the MMAs issued so far have retired
__device__ __forceinline__ void umma2_commit_mc(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "h"(mask)
               : "memory");
}
// kind::f16, D = f32, A = B = bf16, K-major, cta_group::2: M = 256 over the pair
__host__ __device__ constexpr uint32_t make_idesc_bf16_m256(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
}

__host__ __device__ inline float a_val(int rank, int m, int k) { return (float)(((m + 3 * rank + k) % 5) - 2); }
__host__ __device__ inline float b_val(int n, int k) { return (float)(((n + 2 * k) % 7) - 3); }

// MODE 0: functional (one K-block of 64 = 4 MMAs, N = n1; D of both CTAs to global); MODE 1: timing of the mix (n1, n2)
template <int MODE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) pair_kernel(int iters, int n1, int n2, float* out, long long* cycles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* a_sm = smem;                  // 2 x (128 rows x 128 B)
  uint8_t* b_sm = smem + 32768;          // up to 128 rows x 128 B (this CTA's half of B), twice
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 32768 + 32768);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  const uint32_t rank = cluster_ctarank();
  const int tid = threadIdx.x;
  // operands: K-major, 128-byte rows (64 bf16), SWIZZLE_128B: 16-byte chunk c of row r at chunk c ^ (r & 7)
  for (int i = tid; i < 2 * 128 * 64; i += blockDim.x) {
    const int t = i / (128 * 64), r = (i / 64) % 128, k = i % 64;
    const float va = (MODE == 0) ? a_val((int)rank, r, k) : (float)(((i * 7) % 5) - 2);
    const __nv_bfloat16 h = __float2bfloat16_rn(va);
    *reinterpret_cast<__nv_bfloat16*>(a_sm + t * 16384 + sw128_off((uint32_t)r, (uint32_t)(k / 8)) + (k % 8) * 2) = h;
    const int nglob = r + (int)rank * (n1 / 2);          // this CTA holds B rows [rank * n1/2, (rank + 1) * n1/2)
    const float vb = (MODE == 0) ? b_val(nglob, k) : (float)(((i * 3) % 7) - 3);
    *reinterpret_cast<__nv_bfloat16*>(b_sm + t * 16384 + sw128_off((uint32_t)r, (uint32_t)(k / 8)) + (k % 8) * 2) = __float2bfloat16_rn(vb);
  }
  if (tid == 0) {
    mbar_init(bar, 1);
    fence_barrier_init();
  }
  if (tid < 32) tmem_alloc2(slot, 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                    // both CTAs' operands, barriers and TMEM are ready
  tc_fence_after();
  const uint32_t tmem = *slot;
  long long t0 = 0;
  if (rank == 0 && tid == 0) {
    const uint64_t dk = make_smem_desc(0, 16, 1024, SWZ_128B);
    const uint32_t a = smem_u32(a_sm), b = smem_u32(b_sm);
    const uint32_t i1 = make_idesc_bf16_m256(n1), i2 = make_idesc_bf16_m256(n2 ? n2 : 16);
    t0 = clock64();
    if (MODE == 0) {
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) umma2_bf16(tmem, desc_with_addr(dk, a + ks * 32), desc_with_addr(dk, b + ks * 32), i1, ks ? 1u : 0u);
    } else {
#pragma unroll 1
      for (int i = 0; i < iters; ++i) {
        const uint32_t d = tmem + (uint32_t)(((i >> 4) & 1) * 256);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          umma2_bf16(d, desc_with_addr(dk, a + ks * 32), desc_with_addr(dk, b + ks * 32), i1, ((i & 15) | ks) ? 1u : 0u);
          if (n2) umma2_bf16(d, desc_with_addr(dk, a + 16384 + ks * 32), desc_with_addr(dk, b + 16384 + ks * 32), i2, 1u);
        }
      }
    }
    umma2_commit_mc(bar, 3);
  }
  mbar_wait(bar, 0);                     // both CTAs: the pair's MMAs have retired
  tc_fence_after();
  if (rank == 0 && tid == 0 && cycles != nullptr) cycles[blockIdx.x / 2] = clock64() - t0;
  if (MODE == 0) {
    // every warp reads its 32 TMEM lanes (rows) x n1 columns of THIS CTA's accumulator
    const int warp = tid >> 5, lane = tid & 31;
    for (int c0 = 0; c0 < n1; c0 += 32) {
      uint32_t v[32];
      tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
      tmem_ld_wait();
      for (int j = 0; j < 32; ++j) out[((size_t)(blockIdx.x) * 128 + warp * 32 + lane) * 256 + c0 + j] = __uint_as_float(v[j]);
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (tid < 32) {
    tc_fence_after();
    tmem_dealloc2(tmem, 512);
  }
}

int main() {
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, 0) != cudaSuccess) { printf("{\"error\": \"no device\"}\n"); return 1; }
  const int sms = prop.multiProcessorCount;
  const int smem = 65536 + 1024 + 64;
  cudaFuncSetAttribute(pair_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(pair_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  printf("{\"gpu\": \"%s\", \"sms\": %d,\n \"functional\": [", prop.name, sms);
  // (1) functional check, one CTA pair
  const int ns[] = {128, 64, 256};
  for (int t = 0; t < 3; ++t) {
    const int n = ns[t];
    float* d_out;
    cudaMalloc(&d_out, 2 * 128 * 256 * sizeof(float));
    cudaMemset(d_out, 0, 2 * 128 * 256 * sizeof(float));
    pair_kernel<0><<<2, 128, smem>>>(1, n, 0, d_out, nullptr);
    const cudaError_t e = cudaDeviceSynchronize();
    std::vector<float> h(2 * 128 * 256);
    cudaMemcpy(h.data(), d_out, h.size() * sizeof(float), cudaMemcpyDeviceToHost);
    double maxerr = 0.0;
    int bad = 0;
    for (int r = 0; r < 2; ++r)
      for (int m = 0; m < 128; ++m)
        for (int c = 0; c < n; ++c) {
          double ref = 0.0;
          for (int k = 0; k < 64; ++k) ref += (double)a_val(r, m, k) * (double)b_val(c, k);
          const double err = fabs((double)h[((size_t)r * 128 + m) * 256 + c] - ref);
          if (err > maxerr) maxerr = err;
          if (err > 1e-3) ++bad;
        }
    printf("%s{\"N\": %d, \"cuda\": \"%s\", \"max_abs_err\": %.3g, \"mismatches\": %d}", t ? ", " : "", n, cudaGetErrorString(e), maxerr, bad);
    cudaFree(d_out);
    if (e != cudaSuccess) { printf("]}\n"); return 1; }
  }
  printf("],\n \"timing_unit\": \"SM cycles per K-step (K = 16) of the mix, median over the CTA pairs; cta_group::1 figures from profiles/r02_umma_mix.json\",\n \"timing\": [");
  const int pairs = sms / 2;
  long long* d_cyc;
  cudaMalloc(&d_cyc, pairs * sizeof(long long));
  struct Mix { const char* name; int n1, n2; double one_cta; };
  const Mix mixes[] = {{"N256 alone", 256, 0, 128.0}, {"N128 alone", 128, 0, 64.0}, {"N64 alone", 64, 0, 60.0},
                       {"conv: N128 + N64", 128, 64, 113.3}, {"fc: N256 + N128", 256, 128, 192.0}};
  for (size_t i = 0; i < sizeof(mixes) / sizeof(mixes[0]); ++i) {
    const int iters = 4000;
    double med = -1.0;
    for (int rep = 0; rep < 3; ++rep) {
      pair_kernel<1><<<2 * pairs, 128, smem>>>(iters, mixes[i].n1, mixes[i].n2, nullptr, d_cyc);
      if (cudaDeviceSynchronize() != cudaSuccess) { med = -1.0; break; }
      std::vector<long long> h(pairs);
      cudaMemcpy(h.data(), d_cyc, pairs * sizeof(long long), cudaMemcpyDeviceToHost);
      std::sort(h.begin(), h.end());
      med = (double)h[pairs / 2] / (4.0 * iters);
    }
    printf("%s\n  {\"mix\": \"%s\", \"cycles_per_kstep_pair\": %.1f, \"rows_per_mma\": 256, \"cta_group_1_cycles_for_128_rows\": %.1f}", i ? "," : "",
           mixes[i].name, med, mixes[i].one_cta);
  }
  printf("\n]}\n");
  cudaFree(d_cyc);
  return 0;
}
