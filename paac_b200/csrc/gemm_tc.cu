// tcgen05 (5th-generation tensor core) implicit-GEMM forward for the conv / fc layers, sm_100a.
//
//   Y[m, n] = act( in_scale * sum_k im2col(X)[m, k] W[k, n] + bias[n] )      networks.py:12-21, 49-60, 115
//
// One persistent CTA per SM, warp-specialised:
//   warps 0-7   A producers (two groups of 128 threads alternate K-blocks).  Thread r of a group owns tile
//               row r: it gathers the 32 consecutive k of its im2col row (128 contiguous bytes in NHWC, or
//               32 bytes of the uint8 state), converts to tf32 (round-to-nearest; in TF32X3 mode also the
//               residual lo = rna(a - hi)) and stores 16-byte chunks into the 128B-swizzled K-major operand
//               tile in shared memory.  im2col is never materialised.
//   warp  8     lane 0 issues tcgen05.mma.cta_group::1.kind::tf32 (M = 128, N = BN, K = 8) into a TMEM
//               accumulator; tcgen05.commit releases the smem stage / publishes the accumulator.
//   warps 9-12  epilogue: tcgen05.ld the accumulator rows (32 lanes per warp), scale + bias + ReLU, store.
// B (weights) is prepacked once per forward call by pack_weights_kernel into the exact swizzled shared-memory
// image of every (n-tile, k-block), so a single cp.async.bulk (TMA engine) per stage lands it, completing on
// the stage's mbarrier.  Accumulators are double-buffered in TMEM so the epilogue of tile i overlaps the
// mainloop of tile i+1.
//
// TF32X3: A = Ahi + Alo, B = Bhi + Blo, D += Alo*Bhi + Ahi*Blo + Ahi*Bhi (fp32 accumulate in TMEM): the dropped
// Alo*Blo term is ~2^-22 relative, which is what meets the 1e-4 parity bar that plain TF32 (2^-11 operand
// rounding) does not.  For the uint8 input layer the pixel values 0..255 are exact in tf32 (no lo part) and the
// 1/255 scale is applied to the accumulator.
#include "common.cuh"

namespace paacb {

// ------------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t* slot_in_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot_in_smem)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// D[tmem] (+)= A[smem desc] * B[smem desc], kind::tf32, issued by one thread.
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// mbarrier arrive when all previously issued MMAs of this thread have completed (implies fence::before_thread_sync).
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// 32 lanes x 32 consecutive 32-bit columns: thread t of the warp receives row (lane base + t).
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ uint32_t f32_to_tf32_rna(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}

// Shared-memory matrix descriptor, K-major, SWIZZLE_128B: rows of 128 bytes, 8-row groups 1024 bytes apart.
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t smem_addr) {
  return (uint64_t)((smem_addr >> 4) & 0x3FFFu) | (1ull << 16) /* LBO (unused for swizzled K-major) */ |
         (64ull << 32) /* SBO = 1024 B */ | (1ull << 46) /* descriptor version (sm_100) */ | (2ull << 61) /* SWIZZLE_128B */;
}
// Instruction descriptor: D = f32, A = B = tf32, both K-major, M = 128, N = n.
__host__ __device__ constexpr uint32_t make_idesc_tf32(int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

// ------------------------------------------------------------------------------------------------
// weight prepack: W[K, N] fp32 -> per (n-tile, k-block) swizzled K-major images of tf32 hi and lo
// ------------------------------------------------------------------------------------------------
// image index: ((nt * KB + kb) * BN + row) * 32 + ((chunk ^ (row & 7)) * 4 + e),  k = kb*32 + chunk*4 + e, n = nt*BN + row
__global__ void pack_weights_kernel(const float* __restrict__ w, int K, int N, int BN, uint32_t* __restrict__ hi,
                                    uint32_t* __restrict__ lo) {
  const int KB = K / 32;
  const int64_t total = (int64_t)N * (K / 4);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int n = (int)(i % N);
    const int k4 = (int)(i / N);              // group of 4 consecutive k
    const int kb = k4 >> 3, chunk = k4 & 7;
    const int nt = n / BN, row = n - nt * BN;
    uint32_t h[4], l[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float v = __ldg(w + (int64_t)(k4 * 4 + e) * N + n);
      h[e] = f32_to_tf32_rna(v);
      l[e] = f32_to_tf32_rna(v - __uint_as_float(h[e]));
    }
    const int64_t dst = (((int64_t)nt * KB + kb) * BN + row) * 32 + ((chunk ^ (row & 7)) << 2);
    *reinterpret_cast<uint4*>(hi + dst) = make_uint4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<uint4*>(lo + dst) = make_uint4(l[0], l[1], l[2], l[3]);
  }
}

// ------------------------------------------------------------------------------------------------
// forward kernel
// ------------------------------------------------------------------------------------------------
constexpr int kTcProducerThreads = 256;     // two groups of 128
constexpr int kTcThreads = kTcProducerThreads + 32 + 128;

struct TcFwdParams {
  const void* x;
  const uint32_t* w_hi;
  const uint32_t* w_lo;
  const float* bias;
  float* y;
  LayerGeom g;
  int64_t M;
  int64_t m_tiles;
  int n_tiles;
  int kblocks;
  float in_scale;
  int relu;
};

template <int BN, bool U8, bool SPLIT>
struct TcCfg {
  static constexpr int A_BYTES = 128 * 128;
  static constexpr bool A_LO = SPLIT && !U8;
  static constexpr int B_BYTES = BN * 128;
  static constexpr int STAGE_BYTES = A_BYTES * (A_LO ? 2 : 1) + B_BYTES * (SPLIT ? 2 : 1);
  static constexpr int STAGES = (200 * 1024 / STAGE_BYTES) > 8 ? 8 : (200 * 1024 / STAGE_BYTES);
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /* alignment slack */ + 256 /* barriers */;
  static constexpr int TMEM_COLS = (2 * BN) < 32 ? 32 : (2 * BN);
};

template <int BN, bool U8, bool SPLIT>
__global__ void __launch_bounds__(kTcThreads, 1) conv_fwd_tc_kernel(const TcFwdParams p) {
  using Cfg = TcCfg<BN, U8, SPLIT>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
  uint64_t* full_bar = bars;                    // [STAGES]  producers (128 + 1) -> MMA
  uint64_t* empty_bar = bars + STAGES;          // [STAGES]  MMA commit -> producers
  uint64_t* tfull_bar = bars + 2 * STAGES;      // [2]       MMA commit -> epilogue
  uint64_t* tempty_bar = bars + 2 * STAGES + 2; // [2]       epilogue (128) -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const LayerGeom& g = p.g;
  const int64_t total_tiles = p.m_tiles * p.n_tiles;
  const int KB = p.kblocks;

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 128 + 1);
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

  if (warp < 8) {
    // =========================== A producers (+ B bulk copy) ===========================
    const int grp = warp >> 2;                 // 0 or 1: owns iterations it with (it & 1) == grp
    const int r = tid & 127;                   // tile row
    const int SC = g.S * g.C, WC = g.W * g.C, ohw = g.OH * g.OW;
    const uint32_t sw = (uint32_t)(r & 7);
    int64_t cached_tile = -1;
    int64_t rowbase = 0;
    bool row_ok = false;
    int n_tile = 0;

    const int64_t my_tiles = (total_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x;   // tiles of this CTA
    const int64_t total_it = my_tiles * KB;

    auto locate = [&](int64_t it, int& kb) {
      const int64_t tl = it / KB;
      kb = (int)(it - tl * KB);
      if (tl != cached_tile) {
        cached_tile = tl;
        const int64_t t = blockIdx.x + tl * gridDim.x;
        const int64_t mt = t / p.n_tiles;
        n_tile = (int)(t - mt * p.n_tiles);
        const int64_t m = mt * 128 + r;
        row_ok = m < p.M;
        if (row_ok) {
          const int64_t smp = m / ohw;
          const int rem = (int)(m - smp * ohw);
          const int oh = rem / g.OW, ow = rem - oh * g.OW;
          rowbase = ((smp * g.H + (int64_t)oh * g.stride) * g.W + (int64_t)ow * g.stride) * g.C;
        }
      }
    };

    constexpr int NV = U8 ? 2 : 8;             // 16-byte loads per row per K-block
    uint4 buf[NV];
    auto issue = [&](int64_t it) {
      int kb;
      locate(it, kb);
      const int k0 = kb * 32;
      const int kh = k0 / SC, off = k0 - kh * SC;
#pragma unroll
      for (int c = 0; c < NV; ++c) buf[c] = make_uint4(0u, 0u, 0u, 0u);
      if (row_ok) {
        if constexpr (U8) {
          const uint4* src = reinterpret_cast<const uint4*>(reinterpret_cast<const uint8_t*>(p.x) + rowbase + (int64_t)kh * WC + off);
          buf[0] = __ldg(src);
          buf[1] = __ldg(src + 1);
        } else {
          const uint4* src = reinterpret_cast<const uint4*>(reinterpret_cast<const float*>(p.x) + rowbase + (int64_t)kh * WC + off);
#pragma unroll
          for (int c = 0; c < 8; ++c) buf[c] = __ldg(src + c);
        }
      }
    };

    int64_t it = grp;
    if (it < total_it) issue(it);
    for (; it < total_it; it += 2) {
      const int stage = (int)(it % STAGES);
      const uint32_t phase = (uint32_t)((it / STAGES) & 1);
      // current iteration's coordinates (before the prefetch moves the cache)
      const int64_t tl = it / KB;
      const int kb = (int)(it - tl * KB);
      const int cur_ntile = n_tile;
      mbar_wait(&empty_bar[stage], phase ^ 1u);
      uint8_t* st = smem + (size_t)stage * Cfg::STAGE_BYTES;
      uint8_t* a_hi = st;
      uint8_t* a_lo = st + Cfg::A_BYTES;
      uint8_t* b_hi = st + Cfg::A_BYTES * (Cfg::A_LO ? 2 : 1);
      if (r == 0) {
        mbar_arrive_expect_tx(&full_bar[stage], Cfg::B_BYTES * (SPLIT ? 2 : 1));
        const int64_t img = ((int64_t)cur_ntile * KB + kb) * (BN * 32);
        bulk_g2s(b_hi, p.w_hi + img, Cfg::B_BYTES, &full_bar[stage]);
        if constexpr (SPLIT) bulk_g2s(b_hi + Cfg::B_BYTES, p.w_lo + img, Cfg::B_BYTES, &full_bar[stage]);
      }
      // convert + swizzled store of this row's 32 k
      uint8_t* rowp_hi = a_hi + r * 128;
      uint8_t* rowp_lo = a_lo + r * 128;
      if constexpr (U8) {
        const uint32_t wds[8] = {buf[0].x, buf[0].y, buf[0].z, buf[0].w, buf[1].x, buf[1].y, buf[1].z, buf[1].w};
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const uint32_t wv = wds[c];
          uint4 o;      // small integers are exact in tf32: the fp32 bit pattern is the tf32 operand
          o.x = __float_as_uint((float)(wv & 0xffu));
          o.y = __float_as_uint((float)((wv >> 8) & 0xffu));
          o.z = __float_as_uint((float)((wv >> 16) & 0xffu));
          o.w = __float_as_uint((float)(wv >> 24));
          *reinterpret_cast<uint4*>(rowp_hi + (((uint32_t)c ^ sw) << 4)) = o;
        }
      } else {
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const float f0 = __uint_as_float(buf[c].x), f1 = __uint_as_float(buf[c].y);
          const float f2 = __uint_as_float(buf[c].z), f3 = __uint_as_float(buf[c].w);
          uint4 h;
          h.x = f32_to_tf32_rna(f0); h.y = f32_to_tf32_rna(f1); h.z = f32_to_tf32_rna(f2); h.w = f32_to_tf32_rna(f3);
          *reinterpret_cast<uint4*>(rowp_hi + (((uint32_t)c ^ sw) << 4)) = h;
          if constexpr (Cfg::A_LO) {
            uint4 l;
            l.x = f32_to_tf32_rna(f0 - __uint_as_float(h.x)); l.y = f32_to_tf32_rna(f1 - __uint_as_float(h.y));
            l.z = f32_to_tf32_rna(f2 - __uint_as_float(h.z)); l.w = f32_to_tf32_rna(f3 - __uint_as_float(h.w));
            *reinterpret_cast<uint4*>(rowp_lo + (((uint32_t)c ^ sw) << 4)) = l;
          }
        }
      }
      // prefetch the next owned iteration while the MMA consumes this one
      if (it + 2 < total_it) issue(it + 2);
      fence_proxy_async();                     // generic-proxy stores -> visible to the tensor core (async proxy)
      mbar_arrive(&full_bar[stage]);
    }
  } else if (warp == 8) {
    // =========================== MMA issuer ===========================
    if ((tid & 31) == 0) {
      constexpr uint32_t idesc = make_idesc_tf32(BN);
      const int64_t my_tiles = (total_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x;
      int64_t it = 0;
      for (int64_t tl = 0; tl < my_tiles; ++tl) {
        const int acc = (int)(tl & 1);
        const uint32_t acc_phase = (uint32_t)((tl >> 1) & 1);
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        for (int kb = 0; kb < KB; ++kb, ++it) {
          const int stage = (int)(it % STAGES);
          const uint32_t phase = (uint32_t)((it / STAGES) & 1);
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t st = smem_u32(smem + (size_t)stage * Cfg::STAGE_BYTES);
          const uint32_t a_hi = st, a_lo = st + Cfg::A_BYTES;
          const uint32_t b_hi = st + Cfg::A_BYTES * (Cfg::A_LO ? 2 : 1), b_lo = b_hi + Cfg::B_BYTES;
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {     // K = 8 tf32 = 32 bytes per instruction
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
          umma_commit(&empty_bar[stage]);      // smem stage reusable once these MMAs retire
        }
        umma_commit(&tfull_bar[acc]);          // accumulator complete
      }
    }
  } else {
    // =========================== epilogue ===========================
    const int ew = warp & 3;                   // TMEM lane quarter this warp may access
    const int lane = tid & 31;
    const int64_t my_tiles = (total_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x;
    for (int64_t tl = 0; tl < my_tiles; ++tl) {
      const int acc = (int)(tl & 1);
      const uint32_t acc_phase = (uint32_t)((tl >> 1) & 1);
      const int64_t t = blockIdx.x + tl * gridDim.x;
      const int64_t mt = t / p.n_tiles;
      const int nt = (int)(t - mt * p.n_tiles);
      const int64_t m = mt * 128 + ew * 32 + lane;
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(acc * BN);
#pragma unroll
      for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(taddr + (uint32_t)c0, v);
        tmem_ld_wait();
        if (m < p.M) {
          float* dst = p.y + m * g.N + nt * BN + c0;
          const float* bp = p.bias + nt * BN + c0;
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            float4 o;
            o.x = fmaf(__uint_as_float(v[j + 0]), p.in_scale, __ldg(bp + j + 0));
            o.y = fmaf(__uint_as_float(v[j + 1]), p.in_scale, __ldg(bp + j + 1));
            o.z = fmaf(__uint_as_float(v[j + 2]), p.in_scale, __ldg(bp + j + 2));
            o.w = fmaf(__uint_as_float(v[j + 3]), p.in_scale, __ldg(bp + j + 3));
            if (p.relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
            *reinterpret_cast<float4*>(dst + j) = o;
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

// ------------------------------------------------------------------------------------------------
// launcher
// ------------------------------------------------------------------------------------------------
template <int BN, bool U8, bool SPLIT>
static int launch_tc_inst(const paacb_ctx* ctx, const TcFwdParams& p, cudaStream_t st) {
  using Cfg = TcCfg<BN, U8, SPLIT>;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(conv_fwd_tc_kernel<BN, U8, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             Cfg::SMEM_BYTES) != cudaSuccess) {
      set_error("conv_fwd_tc: cannot set %d bytes of dynamic shared memory", Cfg::SMEM_BYTES);
      return PAACB_ECUDA;
    }
    attr_set = true;
  }
  const int64_t tiles = p.m_tiles * p.n_tiles;
  const unsigned grid = (unsigned)(tiles < ctx->num_sms ? tiles : ctx->num_sms);
  PAACB_LAUNCH_BEGIN(ctx, K_FWD0 + p.g.index, st);
  conv_fwd_tc_kernel<BN, U8, SPLIT><<<grid, kTcThreads, Cfg::SMEM_BYTES, st>>>(p);
  PAACB_LAUNCH_END(ctx, K_FWD0 + p.g.index, st);
  return PAACB_OK;
}

int tc_pack_floats(const LayerGeom& g) { return g.K * g.N; }

int launch_pack_weights(const paacb_ctx* ctx, const LayerGeom& g, const float* w, uint32_t* hi, uint32_t* lo, cudaStream_t st) {
  const int bn = g.N >= 128 ? 128 : g.N;
  const int64_t total = (int64_t)g.N * (g.K / 4);
  const unsigned blocks = (unsigned)((total + 255) / 256);
  PAACB_LAUNCH_BEGIN(ctx, K_PACK, st);
  pack_weights_kernel<<<blocks, 256, 0, st>>>(w, g.K, g.N, bn, hi, lo);
  PAACB_LAUNCH_END(ctx, K_PACK, st);
  return PAACB_OK;
}

int launch_conv_fwd_tc(const paacb_ctx* ctx, const LayerGeom& g, const void* x, const float* w, const float* bias,
                       float* y, int64_t batch, int split3, cudaStream_t st) {
  (void)w;
  if ((g.S * g.C) % 32 != 0 || g.K % 32 != 0 || (g.N != 32 && g.N != 64 && g.N % 128 != 0) || ctx->wpack_hi == nullptr)
    return PAACB_EUNSUPPORTED;
  TcFwdParams p;
  p.x = x;
  p.w_hi = ctx->wpack_hi + g.w_off;
  p.w_lo = ctx->wpack_lo + g.w_off;
  p.bias = bias;
  p.y = y;
  p.g = g;
  p.M = batch * g.OH * g.OW;
  if (p.M == 0) return PAACB_OK;
  p.m_tiles = (p.M + 127) / 128;
  p.kblocks = g.K / 32;
  p.in_scale = g.in_u8 ? 0.003921568859368563f : 1.0f;
  p.relu = 1;
  const int bn = g.N >= 128 ? 128 : g.N;
  p.n_tiles = g.N / bn;
#define TC(BN_, U8_)                                                                  \
  (split3 ? launch_tc_inst<BN_, U8_, true>(ctx, p, st) : launch_tc_inst<BN_, U8_, false>(ctx, p, st))
  if (g.in_u8) {
    if (bn == 32) return TC(32, true);
    if (bn == 64) return TC(64, true);
    return PAACB_EUNSUPPORTED;
  }
  if (bn == 32) return TC(32, false);
  if (bn == 64) return TC(64, false);
  return TC(128, false);
#undef TC
}

}  // namespace paacb
