// What does the tensor pipe charge for the instruction MIXES the bf16-split kernels issue?  (tools/probe/umma_peak.cu times
// one shape back to back; the product kernels alternate a wide MMA on A_hi with a narrower one on A_lo into the same
// accumulator, and the conv kernels start their A operand at tap offsets that are not 1024-byte aligned.)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/probe/umma_mix tools/probe/umma_mix.cu
//   tools/probe/umma_mix > gpurun_out/umma_mix.json
// One CTA per SM, one thread issues; SM cycles (clock64) from the first issue to the completion of the last commit, per
// "K-step" of the pattern, median over the CTAs.  kind::f16 on bf16, M = 128, K = 16, SS mode, SWIZZLE_128B.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <algorithm>
#include <vector>
#include "../../paac_b200/csrc/tc_ptx.cuh"

using namespace paacb;

// The pattern is a compile-time constant: the issuing thread's loop must be as tight as the product kernels' (a first
// version with run-time pattern fields measured its own scalar overhead: ~56 cycles per MMA).
//   N1, N2, N3  N of the MMAs of one K-step (0 = absent)
//   D2, D3      TMEM column offset of MMA 2 / 3 relative to MMA 1's accumulator (0 = accumulate into the same columns)
//   AOFF2       byte offset of MMA 2 / 3's A operand (another tile = 16384)
//   TAP         1: A start addresses walk over unaligned tap offsets (multiples of 128 B + 32 B), like tc2_conv.cu
//   ORDER       0: interleaved (1,2,1,2,...), 1: all MMA 1 of four K-steps, then all MMA 2
//   MN          1: both operands MN-major (the weight-gradient kernels)
template <int N1, int N2, int N3, int D2, int D3, int AOFF2, int TAP, int ORDER, int MN>
__global__ void __launch_bounds__(128, 1) mix_kernel(int iters, long long* cycles) {
  struct { int n1 = N1, n2 = N2, n3 = N3, d2 = D2, d3 = D3, a_off2 = AOFF2, tap = TAP, order = ORDER, mn = MN; } constexpr pt;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* a_sm = smem;                  // 2 x (128 rows x 128 B) + slack for tap offsets: 64 KB
  uint8_t* b_sm = smem + 65536;          // 256 rows x 128 B (+ second copy): 64 KB
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 131072);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
  for (int i = threadIdx.x; i < 131072 / 4; i += blockDim.x) {
    const uint32_t h = (uint32_t)i * 2654435761u;
    reinterpret_cast<uint32_t*>(smem)[i] = 0x3F803F80u ^ ((h & 0x8000u) | ((h << 3) & 0x80000000u)) ^ ((h >> 9) & 0x007F007Fu);
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
    const uint64_t dk = pt.mn ? make_smem_desc(0, 8192, 1024, SWZ_128B) : make_smem_desc(0, 16, 1024, SWZ_128B);
    const uint32_t kstep = pt.mn ? 2048u : 32u;
    const uint32_t a = smem_u32(a_sm), b = smem_u32(b_sm);
    const uint32_t i1 = make_idesc_bf16(pt.n1, pt.mn, pt.mn), i2 = make_idesc_bf16(pt.n2 ? pt.n2 : 8, pt.mn, pt.mn),
                   i3 = make_idesc_bf16(pt.n3 ? pt.n3 : 8, pt.mn, pt.mn);
    const long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
      constexpr int extent = N1 > D2 + N2 ? (N1 > D3 + N3 ? N1 : D3 + N3) : (D2 + N2 > D3 + N3 ? D2 + N2 : D3 + N3);
      const uint32_t d = tmem + (uint32_t)(extent <= 256 ? ((i >> 4) & 1) * 256 : 0);        // a new accumulator every 16 K-blocks
      const uint32_t acc = (i & 15) ? 1u : 0u;
#pragma unroll
      for (int pass = 0; pass < (pt.order ? 2 : 1); ++pass) {
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const uint32_t ao = pt.tap ? (uint32_t)(((ks * 5 + 2) % 9) * 1152 + (ks & 3) * 32) : ks * kstep;
          const uint64_t ad = desc_with_addr(dk, a + ao), bd = desc_with_addr(dk, b + ks * kstep);
          if (!pt.order || pass == 0) umma_bf16(d, ad, bd, i1, (acc | ks) ? 1u : 0u);
          if (pt.n2 && (!pt.order || pass == 1)) umma_bf16(d + pt.d2, desc_with_addr(dk, a + ao + pt.a_off2), bd, i2, (pt.d2 == 0 || acc || ks) ? 1u : 0u);
          if (pt.n3 && (!pt.order || pass == 1)) umma_bf16(d + pt.d3, desc_with_addr(dk, a + ao + pt.a_off2), bd, i3, (pt.d3 == 0 || acc || ks) ? 1u : 0u);
        }
      }
    }
    umma_commit(bar);
    mbar_wait(bar, 0);
    cycles[blockIdx.x] = clock64() - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

template <int N1, int N2, int N3, int D2, int D3, int AOFF2, int TAP, int ORDER, int MN>
static void run(int sms, const char* name, bool first) {
  const int iters = 4000;
  const int smem = 131072 + 1024 + 64;
  auto kern = mix_kernel<N1, N2, N3, D2, D3, AOFF2, TAP, ORDER, MN>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  long long* d;
  cudaMalloc(&d, sms * sizeof(long long));
  std::vector<double> meds;
  for (int rep = 0; rep < 4; ++rep) {
    kern<<<sms, 128, smem>>>(iters, d);
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("launch failed\n"); return; }
    std::vector<long long> h(sms);
    cudaMemcpy(h.data(), d, sms * sizeof(long long), cudaMemcpyDeviceToHost);
    std::sort(h.begin(), h.end());
    if (rep) meds.push_back((double)h[sms / 2] / (4.0 * iters));
  }
  cudaFree(d);
  std::sort(meds.begin(), meds.end());
  const double c = meds[meds.size() / 2], ideal = (N1 + N2 + N3) / 2.0;
  printf("%s\n  {\"pattern\": \"%s\", \"cycles_per_kstep\": %.1f, \"ideal\": %.0f, \"ratio\": %.2f}", first ? "" : ",", name, c, ideal, c / ideal);
}

int main() {
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, 0) != cudaSuccess) { printf("{\"error\": \"no device\"}\n"); return 1; }
  const int sms = prop.multiProcessorCount;
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"unit\": \"SM cycles per K-step (K = 16) of the pattern, median over CTAs\",\n \"ideal\": \"M=128: N/2 cycles per MMA at the nominal 8192 dense bf16 flop/cycle/SM\",\n \"rows\": [", prop.name, sms);
  run<256, 0, 0, 0, 0, 0, 0, 0, 0>(sms, "N256 alone", true);
  run<128, 0, 0, 0, 0, 0, 0, 0, 0>(sms, "N128 alone", false);
  run<64, 0, 0, 0, 0, 0, 0, 0, 0>(sms, "N64 alone", false);
  run<32, 0, 0, 0, 0, 0, 0, 0, 0>(sms, "N32 alone", false);
  run<256, 128, 0, 0, 0, 16384, 0, 0, 0>(sms, "fc: N256 (A_hi) + N128 (A_lo) into the same columns", false);
  run<256, 128, 0, 256, 0, 16384, 0, 0, 0>(sms, "fc: N256 + N128 into OTHER columns", false);
  run<256, 128, 0, 0, 0, 16384, 0, 1, 0>(sms, "fc: 4 x N256 then 4 x N128 (same columns)", false);
  run<128, 128, 128, 128, 0, 16384, 0, 0, 0>(sms, "fc: three N128 (hi*hi, hi*lo -> other columns, lo*hi)", false);
  run<128, 64, 0, 0, 0, 16384, 0, 0, 0>(sms, "conv: N128 + N64 same columns, aligned A", false);
  run<128, 64, 0, 0, 0, 16384, 1, 0, 0>(sms, "conv: N128 + N64 same columns, tap-offset A", false);
  run<128, 64, 0, 128, 0, 16384, 1, 0, 0>(sms, "conv: N128 + N64 OTHER columns, tap-offset A", false);
  run<128, 64, 0, 0, 0, 16384, 1, 1, 0>(sms, "conv: 4 x N128 then 4 x N64, tap-offset A", false);
  run<128, 0, 0, 0, 0, 0, 1, 0, 0>(sms, "conv: N128 alone, tap-offset A", false);
  run<64, 0, 0, 0, 0, 0, 1, 0, 0>(sms, "conv: N64 alone, tap-offset A", false);
  run<256, 128, 0, 0, 0, 16384, 1, 0, 0>(sms, "conv dgrad2: N256 + N128 same columns, tap-offset A", false);
  run<64, 0, 0, 0, 0, 0, 0, 0, 1>(sms, "wgrad (MN-major): N64 alone", false);
  run<128, 0, 0, 0, 0, 0, 0, 0, 1>(sms, "wgrad (MN-major): N128 alone", false);
  run<128, 64, 0, 0, 0, 16384, 0, 0, 1>(sms, "wgrad (MN-major): N128 + N64 same columns", false);
  run<64, 64, 64, 0, 0, 16384, 0, 0, 1>(sms, "wgrad (MN-major): three N64", false);
  printf("\n]}\n");
  return 0;
}
