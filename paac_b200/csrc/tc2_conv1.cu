// bf16-split tensor-core path, part 4: the first conv layer forward as an INT8 patch-resident implicit GEMM.
//
//   Y[p, co] = relu( (1/255) * sum_{kh, kw, c} X[p; kh, kw, c] * W[kh, kw, c, co] + b[co] )      networks.py:115, 12-21
//
// The layer's input is the uint8 state itself.  Instead of converting 28 KB of pixels per sample to a floating type
// (the first version spent more time on that than on the MMAs), the pixels feed tcgen05.mma.kind::i8 as they are and the
// WEIGHTS are moved to integers: per output channel, w = s_c/63 * (d0 + d1/64 + d2/4096) with three signed digits
// |d0| <= 63, |d1|, |d2| <= 32 -- a fixed-point image of w with error <= 2^-19 of the channel's largest weight (the
// bf16-split weights of the other layers carry 2^-18 of EACH weight).  The three digit images are three groups of N
// columns of ONE MMA (N = 96): int32 accumulation is exact (|acc| <= 255 * 63 * 256 < 2^22), and the epilogue forms
// s_c / (63 * 255) * (acc0 + acc1/64 + acc2/4096) + b in fp32.
//
// A row of the GEMM is an output position, enumerated over the input grid as in tc2_conv.cu: the patch of a tile is
// 9 rows of each of the 4 row-parity planes of the state (input row ih = 4 q + rho), landed by the TMA engine WITHOUT
// swizzle as 16-byte units (4 pixels x 4 channels).  Consecutive output positions are 16 bytes apart and a filter row
// (8 pixels x 4 channels = 32 bytes = one K = 32 MMA) spans two consecutive units, which is exactly the no-swizzle
// K-major canonical layout with LBO = 16 (the second 16-byte K chunk of row r IS the first chunk of row r + 1) and
// SBO = 128: the overlapping windows of the convolution cost nothing.  Filter row kh is plane kh & 3 at a row offset.
#include "tc2.cuh"

namespace paacb {

// NC = output channels: 32 (Nature; bf16-split planes out) or 16 (NIPS; fp32 out, the tf32 pipeline's activation format)
template <int NC>
struct C1 {
  static constexpr int ND = 3 * NC;                          // MMA N: three digit images
  static constexpr int WBYTES = 2 * ND * 128;                // K = 256 bytes per row: two 128-byte K-blocks of ND rows
  static constexpr int EPI_WARPS = 8 * (NC / 16);            // 2 accumulator buffers x (NC / 16) channel groups x 4 TMEM lane quarters
  static constexpr int THREADS = 64 + 32 * EPI_WARPS;
  static_assert(NC == 16 || NC == 32, "conv1 has 16 or 32 output channels");
  static_assert(WBYTES % 1024 == 0, "1024-byte aligned regions");
};
constexpr int kC1_WU = 21, kC1_HQ = 21, kC1_OH = 20, kC1_OW = 20;
// A tile is SIX plane rows (6 x 21 = 126 of the 128 MMA rows): a thread's position inside its plane row is a kernel constant.
// Epilogue history, all measured (profiles/r01_*): (1) 16-byte stores from registers at a 64-byte lane stride: 16 L1
// wavefronts per store instruction, l1tex at 72 % with DRAM at 47 % -- 0.95 ms per step; (2) a swizzled staging tile
// drained by TMA stores, two tiles in flight: 0.89-0.92 ms, half of it the per-tile skeleton (two 256-thread barriers, a
// proxy fence and the wait for the TMA engine); (3) 16 epilogue warps, each thread owning 16 channels of one position =
// exactly ONE 256-bit store (a full 32-byte sector) per plane, no shared memory, no barrier: 0.64 ms.
constexpr int kC1_TROWS = 6;
constexpr int kC1_ROWS = kC1_TROWS + 1;                // + 1 plane row for the filter rows kh >= 4
constexpr int kC1_PLANE = 16 * kC1_WU * kC1_ROWS;      // 2,352 bytes per parity plane patch
constexpr int kC1_BOX = 4 * kC1_PLANE;                 // one TMA box per tile: [4 parities][7 plane rows][336 bytes]
constexpr int kC1_SLOT = 10240;                        // >= 3 planes + (127 + 21 + 2) units: the two spare MMA rows read stale bytes
constexpr int kC1_NSLOTS = 6;                            // six tiles of patches in flight
constexpr int kC1_SMEM_FIXED = kC1_NSLOTS * kC1_SLOT + 256 + 1024 + (2 * kC1_NSLOTS + 5) * 8 + 16;   // + C1<NC>::WBYTES
constexpr int kC1_TMEM = 256;                            // two accumulator buffers of <= 96 columns at 0 and 128
static_assert((kC1_NSLOTS * kC1_SLOT) % 1024 == 0, "1024-byte aligned regions");
static_assert(3 * kC1_PLANE + (127 + kC1_WU + 2) * 16 <= kC1_SLOT, "slot holds every byte an MMA row can address");

struct Conv1Params {
  CUtensorMap tmA;         // uint8 states as (84 words = one 336-byte image row, b * 21 plane rows, 4 row parities)
  CUtensorMap tmW;         // int8 digit image [96 rows = (digit, co)][256 k]
  int num_tiles;
  int batch;
  const float* bias;
  const float* wscale;     // [NC]: s_c / (63 * 255)
  float* out_f32;          // fp32 output [b, 20, 20, NC] (F32OUT variants) instead of the planes
  int dbg;                 // PAACB_DBG ablations (timing experiments only): 1 no stores, 2 no epilogue arithmetic, 4 no MMAs, 8 no A loads
  uint8_t* out_hi;
  uint8_t* out_lo;
};

// weights -> three int8 digit images + per-channel scale.  One block per output channel, one thread per k.
__global__ void __launch_bounds__(256) pack_conv1_i8_kernel(const float* __restrict__ w, int nc, int8_t* __restrict__ wq,
                                                            float* __restrict__ wscale) {
  __shared__ float red[8];
  const int c = blockIdx.x, k = threadIdx.x;
  const float v = __ldg(w + k * nc + c);
  float m = fabsf(v);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((k & 31) == 0) red[k >> 5] = m;
  __syncthreads();
  float s = red[0];
#pragma unroll
  for (int i = 1; i < 8; ++i) s = fmaxf(s, red[i]);
  // double arithmetic: the digits must reproduce x = 63 w / s to 2^-13 absolute
  const double x = (s > 0.f) ? 63.0 * (double)v / (double)s : 0.0;
  const double d0 = rint(x);
  const double r1 = (x - d0) * 64.0;
  const double d1 = rint(r1);
  const double d2 = rint((r1 - d1) * 64.0);
  wq[(0 * nc + c) * 256 + k] = (int8_t)d0;
  wq[(1 * nc + c) * 256 + k] = (int8_t)d1;
  wq[(2 * nc + c) * 256 + k] = (int8_t)d2;
  if (k == 0) wscale[c] = s / (63.0f * 255.0f);
}

int launch_pack_conv1_i8(const paacb_ctx* ctx, const float* params, cudaStream_t st) {
  const LayerGeom& g = ctx->layer[0];
  if (g.K != 256 || (g.N != 16 && g.N != 32) || ctx->wq_i8 == nullptr) return PAACB_EUNSUPPORTED;
  PAACB_LAUNCH_BEGIN(ctx, K_PACK, st);
  pack_conv1_i8_kernel<<<g.N, 256, 0, st>>>(params + g.w_off, g.N, ctx->wq_i8, ctx->wq_scale);
  PAACB_LAUNCH_END(ctx, K_PACK, st);
  return PAACB_OK;
}

// D[tmem] (+)= A[smem] * B[smem], kind::i8 (uint8 x int8 -> int32), K = 32 per instruction
__device__ __forceinline__ void umma_i8(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// D = int32, A = unsigned 8-bit, B = signed 8-bit, both K-major, M = 128
__host__ __device__ constexpr uint32_t make_idesc_i8(int n) {
  return (2u << 4) | (0u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

template <int NC, bool F32OUT>
__global__ void __launch_bounds__(C1<NC>::THREADS, 1) conv1_i8_kernel(const __grid_constant__ Conv1Params p) {
  constexpr int kC1_ND = C1<NC>::ND, kC1_WBYTES = C1<NC>::WBYTES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* ring = smem;
  uint8_t* wsm = smem + kC1_NSLOTS * kC1_SLOT;
  float2* s_sb = reinterpret_cast<float2*>(wsm + kC1_WBYTES);       // per channel: (scale, bias)
  uint64_t* bars = reinterpret_cast<uint64_t*>(wsm + kC1_WBYTES + 256);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kC1_NSLOTS;
  uint64_t* w_bar = bars + 2 * kC1_NSLOTS;
  uint64_t* tfull_bar = w_bar + 1;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    for (int s = 0; s < kC1_NSLOTS; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(w_bar, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], 128 * (NC / 16));
    }
    fence_barrier_init();
    tma_prefetch_desc(&p.tmA);
    tma_prefetch_desc(&p.tmW);
  }
  if (tid >= 64 && tid < 64 + NC) s_sb[tid - 64] = make_float2(__ldg(p.wscale + tid - 64), __ldg(p.bias + tid - 64));
  if (warp == 1) tmem_alloc(tmem_slot, kC1_TMEM);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // =========================== TMA producer ===========================
    if (elect_one_sync()) {
      mbar_arrive_expect_tx(w_bar, kC1_WBYTES);
      tma_load_2d(wsm, &p.tmW, 0, 0, w_bar);
      tma_load_2d(wsm + kC1_ND * 128, &p.tmW, 128, 0, w_bar);
      int slot = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        mbar_wait(&empty_bar[slot], phase ^ 1u);
        if (PAACB_DBGV(p.dbg) & 8) { mbar_arrive(&full_bar[slot]); if (++slot == kC1_NSLOTS) { slot = 0; phase ^= 1u; } continue; }
        mbar_arrive_expect_tx(&full_bar[slot], kC1_BOX);
        tma_load_3d(ring + slot * kC1_SLOT, &p.tmA, 0, tile * kC1_TROWS, 0, &full_bar[slot]);
        if (++slot == kC1_NSLOTS) { slot = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer ===========================
    const bool leader = elect_one_sync();
    constexpr uint32_t idesc = make_idesc_i8(kC1_ND);
    const uint64_t adesc0 = make_smem_desc(0, 16, 128, SWZ_NONE);       // rows 16 B apart, K chunks 16 B apart (overlapping windows)
    const uint64_t bdesc0 = make_smem_desc(0, 16, 1024, SWZ_128B);
    const uint32_t ring_a = smem_u32(ring), w_a = smem_u32(wsm);
    mbar_wait(w_bar, 0);
    int slot = 0;
    uint32_t phase = 0;
    int tl = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++tl) {
      const int ab = tl & 1;
      const uint32_t aph = (uint32_t)((tl >> 1) & 1);
      mbar_wait(&tempty_bar[ab], aph ^ 1u);
      tc_fence_after();
      const uint32_t d0 = tmem_base + (uint32_t)(ab * 128);
      mbar_wait(&full_bar[slot], phase);
      tc_fence_after();
      if (leader) {
#pragma unroll
        for (int kh = 0; kh < ((PAACB_DBGV(p.dbg) & 4) ? 1 : 8); ++kh) {           // filter row kh: parity plane kh & 3, one plane row further down for kh >= 4
          const uint32_t a = ring_a + (uint32_t)(slot * kC1_SLOT + (kh & 3) * kC1_PLANE + (kh >> 2) * kC1_WU * 16);
          const uint64_t bd = desc_with_addr(bdesc0, w_a + (uint32_t)((kh / 4) * (kC1_ND * 128) + (kh % 4) * 32));
          umma_i8(d0, desc_with_addr(adesc0, a), bd, idesc, kh ? 1u : 0u);
        }
        umma_commit(&empty_bar[slot]);
      }
      __syncwarp();
      if (++slot == kC1_NSLOTS) { slot = 0; phase ^= 1u; }
      if (leader) umma_commit(&tfull_bar[ab]);
      __syncwarp();
    }
  } else {
    // =========================== epilogue ===========================
    const int ew = warp & 3;
    const int r = ew * 32 + lane;                   // MMA row = unit r of the tile's 6 x 21 units
    const int grp = ((warp - 2) >> 2) & 1;          // epilogue group = accumulator buffer it drains
    const int half = (warp - 2) >> 3;               // channels [16 half, 16 half + 16): two warps share a tile row, which doubles
                                                    // the warps that hide the TMEM / shared-memory / barrier latencies
    const int pr = r / kC1_WU, ju = r - pr * kC1_WU;
    const bool rowok = (r < kC1_TROWS * kC1_WU) && (ju < kC1_OW);
    const float4* sb4 = reinterpret_cast<const float4*>(s_sb) + half * 8;
    for (int tl = grp, tile = blockIdx.x + grp * (int)gridDim.x; tile < p.num_tiles; tile += 2 * (int)gridDim.x, tl += 2) {
      const uint32_t aph = (uint32_t)((tl >> 1) & 1);
      const uint32_t g = (uint32_t)tile * kC1_TROWS + (uint32_t)pr;           // global plane row
      const uint32_t n = g / (uint32_t)kC1_HQ;
      const bool ok = rowok && ((int)(g - n * (uint32_t)kC1_HQ) < kC1_OH) && ((int)n < p.batch);
      mbar_wait(&tfull_bar[grp], aph);
      tc_fence_after();
      uint32_t v0[16], v1[16], v2[16];
      const uint32_t tcol = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(grp * 128 + half * 16);
      tmem_ld16(tcol, v0);
      tmem_ld16(tcol + (uint32_t)NC, v1);
      tmem_ld16(tcol + (uint32_t)(2 * NC), v2);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(&tempty_bar[grp]);           // the accumulator is in registers: release the buffer before the arithmetic
      uint32_t hw[8], lw[8];
      float of[F32OUT ? 16 : 1];
      if (PAACB_DBGV(p.dbg) & 2) {
#pragma unroll
        for (int j = 0; j < 8; ++j) { hw[j] = v0[2 * j] ^ v1[2 * j + 1]; lw[j] = v2[2 * j] ^ v0[2 * j + 1]; }
      } else
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 sb = sb4[j];              // (scale, bias) of channels 2j, 2j + 1: one broadcast 16-byte load
        float o[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          // |acc| < 2^22: int -> float exactly on the integer + FMA pipes (the conversion pipe issues at a quarter rate)
          const float f0 = __uint_as_float(v0[2 * j + e] + 0x4B400000u) - 12582912.0f;
          const float f1 = __uint_as_float(v1[2 * j + e] + 0x4B400000u) - 12582912.0f;
          const float f2 = __uint_as_float(v2[2 * j + e] + 0x4B400000u) - 12582912.0f;
          const float t = fmaf(f2, 1.0f / 4096.0f, fmaf(f1, 1.0f / 64.0f, f0));
          o[e] = fmaxf(fmaf(t, e ? sb.z : sb.x, e ? sb.w : sb.y), 0.f);
        }
        if constexpr (F32OUT) { of[2 * j] = o[0]; of[2 * j + 1] = o[1]; }
        else split_bf16x2(o[0], o[1], hw[j], lw[j]);
      }
      if (ok && !(PAACB_DBGV(p.dbg) & 1)) {                 // 16 channels = one full 32-byte sector per plane
        const uint32_t oh = g - n * (uint32_t)kC1_HQ;
        const int64_t oe = (((int64_t)n * kC1_OH + oh) * kC1_OW + ju) * NC + half * 16;      // element index
        if constexpr (F32OUT) {
          stg256(p.out_f32 + oe, reinterpret_cast<const uint32_t*>(of));
          stg256(p.out_f32 + oe + 8, reinterpret_cast<const uint32_t*>(of) + 8);
        } else {
          stg256(p.out_hi + oe * 2, hw);
          stg256(p.out_lo + oe * 2, lw);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kC1_TMEM);
  }
}

template <int NC, bool F32OUT>
static int launch_conv1_inst(const paacb_ctx* ctx, const float* params, const uint8_t* states, uint8_t* out_hi, uint8_t* out_lo,
                             float* out_f32, int64_t batch, cudaStream_t st) {
  const LayerGeom& g = ctx->layer[0];
  if (g.C != 4 || g.stride != 4 || g.R != 8 || g.S != 8 || g.N != NC || g.H != 84 || g.W != 84 || g.OH != kC1_OH)
    return PAACB_EUNSUPPORTED;
  const int64_t plane_rows = batch * kC1_HQ;
  if (plane_rows * kC1_WU >= (1LL << 31) - 256) return PAACB_EUNSUPPORTED;
  constexpr int SMEM = kC1_SMEM_FIXED + C1<NC>::WBYTES;
  static DeviceOnce attr_set;   // kernel attributes are per device: one bit per device index
  if (!attr_set.done(ctx->device)) {
    if (cudaFuncSetAttribute(conv1_i8_kernel<NC, F32OUT>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM) != cudaSuccess) {
      cudaGetLastError();
      set_error("conv1_i8: cannot set %d bytes of dynamic shared memory", SMEM);
      return PAACB_ECUDA;
    }
    // one persistent CTA per SM: ask for the largest shared-memory carve-out whatever the kernel's own request (measured
    // on conv1 forward: the same code ran 10 % slower when its request dropped from 160 KB to 86 KB)
    cudaFuncSetAttribute(conv1_i8_kernel<NC, F32OUT>, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
    cudaGetLastError();
    attr_set.mark(ctx->device);
  }
  Conv1Params p;
  memset(&p, 0, sizeof(p));
  // an image row (84 pixels x 4 channels = 336 bytes) is ONE inner box row of 84 32-bit words (box dimensions are limited
  // to 256 elements); plane row q = ih / 4 and row parity rho are separate dimensions (ih = 4 q + rho), parity slowest,
  // so one box lands the four parity planes of a tile one after the other
  const uint64_t adims[3] = {84u, (uint64_t)batch * kC1_HQ, 4u};
  const uint64_t astr[2] = {4u * 336u, 336u};
  const uint32_t abox[3] = {84u, (uint32_t)kC1_ROWS, 4u};
  int rc = encode_tmap(&p.tmA, states, 4, 3, adims, astr, abox, 0);
  const uint64_t wdims[2] = {256u, (uint64_t)C1<NC>::ND};
  const uint64_t wstr[1] = {256u};
  const uint32_t wbox[2] = {128u, (uint32_t)C1<NC>::ND};
  if (rc == PAACB_OK) rc = encode_tmap(&p.tmW, ctx->wq_i8, 1, 2, wdims, wstr, wbox, 128);
  if (rc != PAACB_OK) return rc;
  p.num_tiles = (int)((plane_rows + kC1_TROWS - 1) / kC1_TROWS);
  p.batch = (int)batch;
  p.bias = params + g.b_off;
  p.wscale = ctx->wq_scale;
  p.dbg = ctx->dbg;
  p.out_hi = out_hi;
  p.out_lo = out_lo;
  p.out_f32 = out_f32;
  const unsigned grid = (unsigned)(p.num_tiles < ctx->num_sms ? p.num_tiles : ctx->num_sms);
  PAACB_LAUNCH_BEGIN(ctx, K_FWD0, st);
  conv1_i8_kernel<NC, F32OUT><<<grid, C1<NC>::THREADS, SMEM, st>>>(p);
  PAACB_LAUNCH_END(ctx, K_FWD0, st);
  return PAACB_OK;
}

// bf16-split pipeline (both architectures): planes out
int launch_conv1_fwd_i8(const paacb_ctx* ctx, const float* params, const uint8_t* states, void* fwd_ws, int64_t batch,
                        const WsSlice& slice, cudaStream_t st) {
  const LayerGeom& g = ctx->layer[0];
  const Planes out = layer_planes(fwd_ws, g.out_act_off, (int64_t)g.OH * g.OW * g.N, slice);
  if (g.N == 16)      // NIPS: 16 channels = one 32-byte sector per position and plane
    return launch_conv1_inst<16, false>(ctx, params, states, out.hi, out.lo, nullptr, batch, st);
  return launch_conv1_inst<32, false>(ctx, params, states, out.hi, out.lo, nullptr, batch, st);
}

// NIPS (16 output channels), tf32 pipeline: fp32 activations out.  The int8 digit scheme is exact to 2^-19 of the largest
// weight of a channel, tighter than the tf32 split of the layers that follow.
int launch_conv1_fwd_i8_f32(const paacb_ctx* ctx, const float* params, const uint8_t* states, float* y, int64_t batch,
                            cudaStream_t st) {
  if (ctx->layer[0].N != 16 || ctx->wq_i8 == nullptr) return PAACB_EUNSUPPORTED;
  return launch_conv1_inst<16, true>(ctx, params, states, nullptr, nullptr, y, batch, st);
}

}  // namespace paacb
