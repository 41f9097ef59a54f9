// tcgen05 weight-gradient kernel, sm_100a (TF autodiff: Conv2DBackpropFilter / MatMul grad + bias Sum).
//
//   dW[k, n] += w_scale * sum_m im2col(X)[m, k] * dZ[m, n]        db[n] += sum_m dZ[m, n]
//
// As a tensor-core GEMM the 128-row MMA dimension is a tile of k (filter taps x input channels), the N dimension a
// tile of output channels, and the reduction runs over pixels m.  Both operands have the reduction index slowest in
// memory (rows of im2col(X) and of dZ are pixels), so the producers transpose while staging: a thread owns one pixel
// of the 32-pixel K-block and one 32-wide run of k (or n), loads that run with 16-byte loads (contiguous in NHWC),
// converts to tf32 (hi, and lo in TF32X3 mode) and writes it as a COLUMN of the 128B-swizzled K-major operand tile
// (32 conflict-free 4-byte stores: the lanes of a warp are the 32 pixels = the 128 bytes of one operand row).
//
// A persistent CTA per SM works through (k-tile, n-tile, pixel-split) units; partial sums are combined with fp32
// atomics into the zero-initialised gradient (the order of accumulation across splits is not fixed; TF's own GPU
// kernels make no ordering promise either).  The bias gradient falls out of the B producers for free.
//   warps 0-3   A producers (im2col runs; the uint8 state is exact in tf32, its 1/255 is applied in the epilogue)
//   warps 4-7   B producers (dZ runs) + bias column sums
//   warp  8     MMA issuer (one thread)      warps 9-12  epilogue (tcgen05.ld -> red.global.add.f32)
#include "common.cuh"
#include "tc_ptx.cuh"

namespace paacb {

constexpr int kWgThreads = 256 + 32 + 128;

struct WgParams {
  const void* x;          // layer input (uint8 states or fp32 NHWC)
  const float* dz;        // [M, N]
  float* dw;              // [K, N], zero-initialised by the caller
  float* db;              // [N]
  LayerGeom g;
  int64_t M;              // pixels = batch * OH * OW
  int64_t rows_per_split; // multiple of 32
  int splits;
  int k_tiles;
  int n_tiles;
  float w_scale;
};

template <int BN, bool U8, bool SPLIT>
struct WgCfg {
  static constexpr int A_BYTES = 128 * 128;
  static constexpr bool A_LO = SPLIT && !U8;
  static constexpr int B_BYTES = BN * 128;
  static constexpr int STAGE_BYTES = A_BYTES * (A_LO ? 2 : 1) + B_BYTES * (SPLIT ? 2 : 1);
  static constexpr int STAGES = (200 * 1024 / STAGE_BYTES) > 8 ? 8 : (200 * 1024 / STAGE_BYTES);
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;
  static constexpr int TMEM_COLS = (2 * BN) < 32 ? 32 : (2 * BN);
  static constexpr int NVA = U8 ? 2 : 8;      // 16-byte loads per A thread per K-block (32 k)
  static constexpr int NVB = BN / 16;         // 16-byte loads per B thread per K-block (BN / 4 n-values)
  static constexpr int PFA = U8 ? 6 : 2;      // K-blocks in flight
  static constexpr int PFB = (NVB >= 8) ? 2 : (NVB >= 4 ? 3 : 4);
};

struct WgCursor {   // walks the (unit, k-block) iteration space of one CTA in order
  int64_t unit_i;   // index into this CTA's unit list
  int kb;
  int kbs;          // K-blocks of the current unit
  int ktile, nt, split;
  int64_t mbeg, mend;
};

__device__ __forceinline__ void wg_load_unit(const WgParams& p, WgCursor& c, int64_t total_units) {
  const int64_t u = blockIdx.x + c.unit_i * gridDim.x;
  if (u >= total_units) { c.kbs = 0; return; }
  const int per_split = p.k_tiles * p.n_tiles;
  c.split = (int)(u / per_split);
  const int rem = (int)(u - (int64_t)c.split * per_split);
  c.ktile = rem / p.n_tiles;
  c.nt = rem - c.ktile * p.n_tiles;
  c.mbeg = (int64_t)c.split * p.rows_per_split;
  c.mend = c.mbeg + p.rows_per_split < p.M ? c.mbeg + p.rows_per_split : p.M;
  c.kbs = (int)((c.mend - c.mbeg + 31) / 32);
}
__device__ __forceinline__ void wg_advance(const WgParams& p, WgCursor& c, int64_t total_units) {
  if (++c.kb >= c.kbs) {
    c.kb = 0;
    ++c.unit_i;
    wg_load_unit(p, c, total_units);
  }
}

template <int BN, bool U8, bool SPLIT>
__global__ void __launch_bounds__(kWgThreads, 1) wgrad_tc_kernel(const WgParams p) {
  using Cfg = WgCfg<BN, U8, SPLIT>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
  uint64_t* full_bar = bars;                    // [STAGES]  256 producer threads -> MMA
  uint64_t* empty_bar = bars + STAGES;          // [STAGES]  MMA commit -> producers
  uint64_t* tfull_bar = bars + 2 * STAGES;      // [2]
  uint64_t* tempty_bar = bars + 2 * STAGES + 2; // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  const LayerGeom& g = p.g;
  const int64_t total_units = (int64_t)p.k_tiles * p.n_tiles * p.splits;
  const int64_t my_units = (total_units - blockIdx.x + gridDim.x - 1) / gridDim.x;

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 256);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], 128);
    }
    fence_barrier_init();
  }
  if (warp == 8) tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // total K-block iterations of this CTA (identical in every role)
  int64_t total_it = 0;
  {
    WgCursor c;
    c.unit_i = 0; c.kb = 0;
    for (int64_t i = 0; i < my_units; ++i) {
      c.unit_i = i;
      wg_load_unit(p, c, total_units);
      total_it += c.kbs;
    }
  }

  if (warp < 4) {
    // =========================== A producers: im2col runs, transposed into the operand tile ===========================
    constexpr int NV = Cfg::NVA, PF = Cfg::PFA;
    const int j = warp;                        // k-run of this warp: rows j*32 .. j*32+31 of the tile
    const int SC = g.S * g.C, WC = g.W * g.C, ohw = g.OH * g.OW;
    WgCursor lc;
    lc.unit_i = 0; lc.kb = 0;
    wg_load_unit(p, lc, total_units);
    uint4 buf[PF][NV];

    auto issue = [&](uint4 (&b)[NV]) {
#pragma unroll
      for (int c = 0; c < NV; ++c) b[c] = make_uint4(0u, 0u, 0u, 0u);
      const int64_t m = lc.mbeg + (int64_t)lc.kb * 32 + lane;
      const int k0 = lc.ktile * 128 + j * 32;
      if (m < lc.mend && k0 < g.K) {
        const int64_t sm = m / ohw;
        const int rem = (int)(m - sm * ohw);
        const int oh = rem / g.OW, ow = rem - oh * g.OW;
        const int kh = k0 / SC, off = k0 - kh * SC;
        const int64_t e = ((sm * g.H + (int64_t)oh * g.stride + kh) * g.W + (int64_t)ow * g.stride) * g.C + off;
        if constexpr (U8) {
          const uint4* src = reinterpret_cast<const uint4*>(reinterpret_cast<const uint8_t*>(p.x) + e);
          b[0] = __ldg(src);
          b[1] = __ldg(src + 1);
        } else {
          const uint4* src = reinterpret_cast<const uint4*>(reinterpret_cast<const float*>(p.x) + e);
#pragma unroll
          for (int c = 0; c < 8; ++c) b[c] = __ldg(src + c);
        }
      }
      wg_advance(p, lc, total_units);
    };

    auto process = [&](int64_t it, const uint4 (&b)[NV]) {
      const int stage = (int)(it % STAGES);
      const uint32_t phase = (uint32_t)((it / STAGES) & 1);
      mbar_wait(&empty_bar[stage], phase ^ 1u);
      uint8_t* st = smem + (size_t)stage * Cfg::STAGE_BYTES;
      uint8_t* a_hi = st;
      uint8_t* a_lo = st + Cfg::A_BYTES;
      const uint32_t colb = (uint32_t)(lane & 3) * 4u;
      const uint32_t colc = (uint32_t)(lane >> 2);
      float f[32];
      if constexpr (U8) {
        const uint32_t wds[8] = {b[0].x, b[0].y, b[0].z, b[0].w, b[1].x, b[1].y, b[1].z, b[1].w};
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          f[c * 4 + 0] = u8_to_f32(wds[c], 0);
          f[c * 4 + 1] = u8_to_f32(wds[c], 1);
          f[c * 4 + 2] = u8_to_f32(wds[c], 2);
          f[c * 4 + 3] = u8_to_f32(wds[c], 3);
        }
      } else {
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          f[c * 4 + 0] = __uint_as_float(b[c].x); f[c * 4 + 1] = __uint_as_float(b[c].y);
          f[c * 4 + 2] = __uint_as_float(b[c].z); f[c * 4 + 3] = __uint_as_float(b[c].w);
        }
      }
#pragma unroll
      for (int kk = 0; kk < 32; ++kk) {
        const uint32_t row = (uint32_t)(j * 32 + kk);
        const uint32_t off = row * 128u + ((colc ^ (row & 7u)) << 4) + colb;
        const uint32_t h = U8 ? __float_as_uint(f[kk]) : tf32_rna(f[kk]);
        *reinterpret_cast<uint32_t*>(a_hi + off) = h;
        if constexpr (Cfg::A_LO) *reinterpret_cast<uint32_t*>(a_lo + off) = tf32_rna(f[kk] - __uint_as_float(h));
      }
    };

#pragma unroll
    for (int u = 0; u < PF; ++u)
      if (u < total_it) issue(buf[u]);
    for (int64_t base = 0; base < total_it; base += PF) {
#pragma unroll
      for (int u = 0; u < PF; ++u) {
        const int64_t it = base + u;
        if (it < total_it) {
          process(it, buf[u]);
          if (it + PF < total_it) issue(buf[u]);
          fence_proxy_async();
          mbar_arrive(&full_bar[(int)(it % STAGES)]);
        }
      }
    }
  } else if (warp < 8) {
    // =========================== B producers: dZ runs (+ bias column sums) ===========================
    constexpr int NV = Cfg::NVB, PF = Cfg::PFB;
    const int q = warp - 4;                    // this warp's quarter of the BN columns: n-local q*BN/4 .. +BN/4
    WgCursor lc;
    lc.unit_i = 0; lc.kb = 0;
    wg_load_unit(p, lc, total_units);
    uint4 buf[PF][NV];
    int meta_k0[PF];                           // ktile == 0 and start-of-unit / end-of-unit flags for the bias sums
    float bsum[NV * 4];
#pragma unroll
    for (int i = 0; i < NV * 4; ++i) bsum[i] = 0.f;

    auto issue = [&](uint4 (&b)[NV], int& meta) {
#pragma unroll
      for (int c = 0; c < NV; ++c) b[c] = make_uint4(0u, 0u, 0u, 0u);
      const int64_t m = lc.mbeg + (int64_t)lc.kb * 32 + lane;
      if (m < lc.mend) {
        const uint4* src = reinterpret_cast<const uint4*>(p.dz + m * g.N + lc.nt * BN + q * (BN / 4));
#pragma unroll
        for (int c = 0; c < NV; ++c) b[c] = __ldg(src + c);
      }
      // bit 0: this unit contributes to db (first k-tile); bit 1: last K-block of its unit; bits 8..: n-tile
      meta = (lc.ktile == 0 ? 1 : 0) | ((lc.kb == lc.kbs - 1) ? 2 : 0) | (lc.nt << 8);
      wg_advance(p, lc, total_units);
    };

    auto process = [&](int64_t it, const uint4 (&b)[NV], int meta) {
      const int stage = (int)(it % STAGES);
      const uint32_t phase = (uint32_t)((it / STAGES) & 1);
      mbar_wait(&empty_bar[stage], phase ^ 1u);
      uint8_t* st = smem + (size_t)stage * Cfg::STAGE_BYTES;
      uint8_t* b_hi = st + Cfg::A_BYTES * (Cfg::A_LO ? 2 : 1);
      uint8_t* b_lo = b_hi + Cfg::B_BYTES;
      const uint32_t colb = (uint32_t)(lane & 3) * 4u;
      const uint32_t colc = (uint32_t)(lane >> 2);
#pragma unroll
      for (int c = 0; c < NV; ++c) {
        const float f[4] = {__uint_as_float(b[c].x), __uint_as_float(b[c].y), __uint_as_float(b[c].z), __uint_as_float(b[c].w)};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const uint32_t row = (uint32_t)(q * (BN / 4) + c * 4 + e);
          const uint32_t off = row * 128u + ((colc ^ (row & 7u)) << 4) + colb;
          const uint32_t h = tf32_rna(f[e]);
          *reinterpret_cast<uint32_t*>(b_hi + off) = h;
          if constexpr (SPLIT) *reinterpret_cast<uint32_t*>(b_lo + off) = tf32_rna(f[e] - __uint_as_float(h));
          bsum[c * 4 + e] += f[e];
        }
      }
      if (meta & 2) {                          // unit finished: flush or discard the bias partial sums
        if ((meta & 1) && p.db != nullptr) {
          const int nt = meta >> 8;
#pragma unroll
          for (int i = 0; i < NV * 4; ++i) {
            float s = bsum[i];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if (lane == 0) atomicAdd(p.db + nt * BN + q * (BN / 4) + i, s);
          }
        }
#pragma unroll
        for (int i = 0; i < NV * 4; ++i) bsum[i] = 0.f;
      }
    };

#pragma unroll
    for (int u = 0; u < PF; ++u)
      if (u < total_it) issue(buf[u], meta_k0[u]);
    for (int64_t base = 0; base < total_it; base += PF) {
#pragma unroll
      for (int u = 0; u < PF; ++u) {
        const int64_t it = base + u;
        if (it < total_it) {
          const int meta = meta_k0[u];
          process(it, buf[u], meta);
          if (it + PF < total_it) issue(buf[u], meta_k0[u]);
          fence_proxy_async();
          mbar_arrive(&full_bar[(int)(it % STAGES)]);
        }
      }
    }
  } else if (warp == 8) {
    // =========================== MMA issuer ===========================
    {   // warp-uniform loop, one elected lane issues (see igemm_tc.cu)
      constexpr uint32_t idesc = make_idesc_tf32(BN);
      const bool leader = elect_one_sync();
      WgCursor c;
      c.unit_i = 0; c.kb = 0;
      int64_t it = 0;
      for (int64_t ui = 0; ui < my_units; ++ui) {
        c.unit_i = ui;
        wg_load_unit(p, c, total_units);
        const int acc = (int)(ui & 1);
        const uint32_t acc_phase = (uint32_t)((ui >> 1) & 1);
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        for (int kb = 0; kb < c.kbs; ++kb, ++it) {
          const int stage = (int)(it % STAGES);
          const uint32_t phase = (uint32_t)((it / STAGES) & 1);
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t st = smem_u32(smem + (size_t)stage * Cfg::STAGE_BYTES);
          const uint32_t a_hi = st, a_lo = st + Cfg::A_BYTES;
          const uint32_t b_hi = st + Cfg::A_BYTES * (Cfg::A_LO ? 2 : 1), b_lo = b_hi + Cfg::B_BYTES;
          if (leader) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint32_t ko = (uint32_t)ks * 32u;
            uint32_t accum = (kb > 0 || ks > 0) ? 1u : 0u;
            if constexpr (SPLIT) {
              if constexpr (Cfg::A_LO) {
                umma_tf32(d_tmem, make_sw128_desc(a_lo + ko), make_sw128_desc(b_hi + ko), idesc, accum);
                accum = 1u;
              }
              umma_tf32(d_tmem, make_sw128_desc(a_hi + ko), make_sw128_desc(b_lo + ko), idesc, accum);
              accum = 1u;
            }
            umma_tf32(d_tmem, make_sw128_desc(a_hi + ko), make_sw128_desc(b_hi + ko), idesc, accum);
          }
          umma_commit(&empty_bar[stage]);
          }
          __syncwarp();
        }
        if (leader) umma_commit(&tfull_bar[acc]);
        __syncwarp();
      }
    }
  } else {
    // =========================== epilogue: atomically add the partial tile into dW ===========================
    const int ew = warp & 3;
    WgCursor c;
    c.unit_i = 0; c.kb = 0;
    for (int64_t ui = 0; ui < my_units; ++ui) {
      c.unit_i = ui;
      wg_load_unit(p, c, total_units);
      const int acc = (int)(ui & 1);
      const uint32_t acc_phase = (uint32_t)((ui >> 1) & 1);
      const int k = c.ktile * 128 + ew * 32 + lane;
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(acc * BN);
      if constexpr (BN == 16) {                // the NIPS first layer: 16 output channels
        uint32_t v[16];
        tmem_ld16(taddr, v);
        tmem_ld_wait();
        if (k < g.K && c.kbs > 0) {
          float4* dst = reinterpret_cast<float4*>(p.dw + (int64_t)k * g.N + c.nt * BN);      // RED.ADD.F32x4
#pragma unroll
          for (int jx = 0; jx < 4; ++jx)
            atomicAdd(dst + jx, make_float4(__uint_as_float(v[4 * jx]) * p.w_scale, __uint_as_float(v[4 * jx + 1]) * p.w_scale,
                                            __uint_as_float(v[4 * jx + 2]) * p.w_scale, __uint_as_float(v[4 * jx + 3]) * p.w_scale));
        }
      } else {
#pragma unroll
      for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(taddr + (uint32_t)c0, v);
        tmem_ld_wait();
        if (k < g.K && c.kbs > 0) {
          float4* dst = reinterpret_cast<float4*>(p.dw + (int64_t)k * g.N + c.nt * BN + c0);  // RED.ADD.F32x4
#pragma unroll
          for (int jx = 0; jx < 8; ++jx)
            atomicAdd(dst + jx, make_float4(__uint_as_float(v[4 * jx]) * p.w_scale, __uint_as_float(v[4 * jx + 1]) * p.w_scale,
                                            __uint_as_float(v[4 * jx + 2]) * p.w_scale, __uint_as_float(v[4 * jx + 3]) * p.w_scale));
        }
      }
      }
      tc_fence_before();
      mbar_arrive(&tempty_bar[acc]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

template <int BN, bool U8, bool SPLIT>
static int launch_wg_inst(const paacb_ctx* ctx, const WgParams& p, cudaStream_t st) {
  using Cfg = WgCfg<BN, U8, SPLIT>;
  static DeviceOnce attr_set;   // kernel attributes are per device: one bit per device index
  if (!attr_set.done(ctx->device)) {
    if (cudaFuncSetAttribute(wgrad_tc_kernel<BN, U8, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             Cfg::SMEM_BYTES) != cudaSuccess) {
      set_error("wgrad_tc: cannot set %d bytes of dynamic shared memory", Cfg::SMEM_BYTES);
      return PAACB_ECUDA;
    }
    attr_set.mark(ctx->device);
  }
  const int64_t units = (int64_t)p.k_tiles * p.n_tiles * p.splits;
  const unsigned grid = (unsigned)(units < ctx->num_sms ? units : ctx->num_sms);
  PAACB_LAUNCH_BEGIN(ctx, K_WGRAD0 + p.g.index, st);
  wgrad_tc_kernel<BN, U8, SPLIT><<<grid, kWgThreads, Cfg::SMEM_BYTES, st>>>(p);
  PAACB_LAUNCH_END(ctx, K_WGRAD0 + p.g.index, st);
  return PAACB_OK;
}

int launch_conv_wgrad_tc(const paacb_ctx* ctx, const LayerGeom& g, const void* x, const float* dz, float* dw, float* db,
                         int64_t batch, int split3, cudaStream_t st) {
  const int bn = (g.N % 128 == 0) ? 128 : ((g.N % 64 == 0) ? 64 : ((g.N % 32 == 0) ? 32 : ((g.N % 16 == 0) ? 16 : 0)));
  if (bn == 0 || (g.S * g.C) % 32 != 0 || g.K % 32 != 0) return PAACB_EUNSUPPORTED;
  WgParams p;
  memset(&p, 0, sizeof(p));
  p.x = x; p.dz = dz; p.dw = dw; p.db = db; p.g = g;
  p.M = batch * g.OH * g.OW;
  if (p.M == 0) return PAACB_OK;
  p.k_tiles = (g.K + 127) / 128;
  p.n_tiles = g.N / bn;
  const int tiles = p.k_tiles * p.n_tiles;
  const int64_t row_blocks = (p.M + 31) / 32;
  int64_t splits = (2LL * ctx->num_sms + tiles - 1) / tiles;        // about two units per SM
  if (splits > row_blocks) splits = row_blocks;
  if (splits < 1) splits = 1;
  p.rows_per_split = ((row_blocks + splits - 1) / splits) * 32;
  p.splits = (int)((p.M + p.rows_per_split - 1) / p.rows_per_split);
  p.w_scale = g.in_u8 ? 0.003921568859368563f : 1.0f;
#define WG(BN_, U8_) \
  (split3 ? launch_wg_inst<BN_, U8_, true>(ctx, p, st) : launch_wg_inst<BN_, U8_, false>(ctx, p, st))
  if (g.in_u8) {
    if (bn == 16) return WG(16, true);
    if (bn == 32) return WG(32, true);
    if (bn == 64) return WG(64, true);
    return PAACB_EUNSUPPORTED;
  }
  if (bn == 16) return WG(16, false);
  if (bn == 32) return WG(32, false);
  if (bn == 64) return WG(64, false);
  return WG(128, false);
#undef WG
}

}  // namespace paacb
