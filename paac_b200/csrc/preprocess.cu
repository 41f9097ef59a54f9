// K1: Atari observation pipeline as one vectorised uint8 kernel.
//
// Replaces, per env-step (reference file:line):
//   np.amax(frame_pool, axis=0)                atari_emulator.py:72
//   imresize(img, (84, 84), 'nearest')         atari_emulator.py:73  (Pillow NEAREST index tables)
//   ObservationPool.new_observation / get_pooled_observations   environment.py:66-71
//   reset -> fresh 4-plane stack               emulator_runner.py:26-27 + atari_emulator.py:88-96
//
// One CTA per env (grid-stride when the launcher narrows the grid).  Phase 1 streams only the 84 source rows nearest-sampling selects (2 frames x 84 x
// 160 B, 16-byte loads, byte-wise max with __vmaxu4) into shared memory.  Phase 2 gathers the 84
// selected columns from shared memory and merges them into the NHWC uint8 stack: a non-reset step is
// (prev >> 8) | (new << 24) per pixel word (drop the oldest frame, append the newest as channel 3).
// HBM traffic per env-step: 26,880 B frames + 28,224 B previous stack read, 28,224 B written.
#include <stdlib.h>
#include "tc2.cuh"

namespace paacb {

constexpr int kFrameBytes = PAACB_FRAME_H * PAACB_FRAME_W;   // 33,600
constexpr int kRowVec = PAACB_FRAME_W / 16;                   // 10 uint4 per source row
constexpr int kStateVec = PAACB_OBS * PAACB_OBS / 4;          // 1,764 uint4 per state (4 pixels each)
constexpr int kThreads = 256;
constexpr int kVecPerThread = (kStateVec + kThreads - 1) / kThreads;   // 7

__device__ __forceinline__ uint4 shr8(uint4 v) { return make_uint4(v.x >> 8, v.y >> 8, v.z >> 8, v.w >> 8); }

__global__ void __launch_bounds__(kThreads)
preprocess_u8_kernel(const uint8_t* __restrict__ frames, int pairs, const uint8_t* __restrict__ reset,
                     const uint8_t* prev, uint8_t* next, ResizeTables tabs, int64_t n_envs, const StepScalars sc) {
  __shared__ __align__(16) uint8_t plane[PAACB_OBS * PAACB_FRAME_W];
  __shared__ int s_row[PAACB_OBS];
  __shared__ int s_col[PAACB_OBS];
  __shared__ int s_rst;
  const int tid = threadIdx.x;
  if (tid < PAACB_OBS) {
    s_row[tid] = tabs.row[tid] * PAACB_FRAME_W;
    s_col[tid] = tabs.col[tid];
  }
  for (int64_t env = blockIdx.x; env < n_envs; env += gridDim.x) {
  if (sc.over_in != nullptr) {
    // paacb_observe_u8: this step's reward and episode-over flag go into their row of the rollout buffers in the same
    // launch (paac.py:119-123), and the flag is the reset flag of emulator_runner.py:26-27 (one load, shared by the CTA)
    __syncthreads();                       // the previous environment's s_rst has been consumed
    if (tid == 0) {
      const float ov = sc.over_in[env];
      sc.over_out[env] = ov;
      sc.rewards_out[env] = sc.rewards_in[env];
      s_rst = (sc.over_is_reset && ov != 0.f) ? 1 : 0;
    }
    __syncthreads();
  }
  const bool rst = (pairs >= PAACB_STACK) && (((reset != nullptr) && (reset[env] != 0)) || (sc.over_in != nullptr && s_rst != 0));
  const uint4* prev4 = reinterpret_cast<const uint4*>(prev + env * (int64_t)(kStateVec * 16));
  uint4* next4 = reinterpret_cast<uint4*>(next + env * (int64_t)(kStateVec * 16));

  uint4 acc[kVecPerThread];
#pragma unroll
  for (int i = 0; i < kVecPerThread; ++i) {
    const int idx = tid + i * kThreads;
    acc[i] = make_uint4(0u, 0u, 0u, 0u);
    if (!rst && idx < kStateVec) acc[i] = shr8(prev4[idx]);
  }

  const int nplanes = rst ? PAACB_STACK : 1;
  for (int p = 0; p < nplanes; ++p) {
    const uint8_t* f0 = frames + ((env * pairs + p) * 2) * (int64_t)kFrameBytes;
    const uint8_t* f1 = f0 + kFrameBytes;
    __syncthreads();   // tables visible (p == 0) / previous plane fully consumed (p > 0)
    for (int i = tid; i < PAACB_OBS * kRowVec; i += kThreads) {
      const int y = i / kRowVec, c = i - y * kRowVec;
      const int src = s_row[y] + c * 16;
      const uint4 a = __ldg(reinterpret_cast<const uint4*>(f0 + src));
      const uint4 b = __ldg(reinterpret_cast<const uint4*>(f1 + src));
      uint4 m;
      m.x = __vmaxu4(a.x, b.x); m.y = __vmaxu4(a.y, b.y); m.z = __vmaxu4(a.z, b.z); m.w = __vmaxu4(a.w, b.w);
      *reinterpret_cast<uint4*>(plane + y * PAACB_FRAME_W + c * 16) = m;
    }
    __syncthreads();
    const int shift = rst ? 8 * p : 24;
#pragma unroll
    for (int i = 0; i < kVecPerThread; ++i) {
      const int idx = tid + i * kThreads;
      if (idx < kStateVec) {
        const int y = idx / (PAACB_OBS / 4);
        const int x0 = (idx - y * (PAACB_OBS / 4)) * 4;
        const uint8_t* prow = plane + y * PAACB_FRAME_W;
        acc[i].x |= (uint32_t)prow[s_col[x0 + 0]] << shift;
        acc[i].y |= (uint32_t)prow[s_col[x0 + 1]] << shift;
        acc[i].z |= (uint32_t)prow[s_col[x0 + 2]] << shift;
        acc[i].w |= (uint32_t)prow[s_col[x0 + 3]] << shift;
      }
    }
  }
#pragma unroll
  for (int i = 0; i < kVecPerThread; ++i) {
    const int idx = tid + i * kThreads;
    if (idx < kStateVec) next4[idx] = acc[i];
  }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// K1, device-resident frames: the same arithmetic as a persistent, warp-specialised copy pipeline (round 2).
// The kernel above alternates a load phase and a gather phase per CTA and leans on 4-8 resident CTAs per SM to keep
// HBM busy: 0.60 of the copy peak on moved bytes inside a step (4,096 environments = 3.5-7 waves of short CTAs), 0.77 at
// 16,384.  Here ONE CTA per SM keeps kK1Stages environments in flight through the TMA engine:
//   warp 0 (producer)  per environment THREE TMA operations completing on the stage's `full` mbarrier: one 28,224-byte bulk
//                      copy of the previous stack and two tensor boxes of the frame rows.  Pillow's NEAREST table for
//                      210 -> 84 rows is row(2k) = 5k + 1, row(2k + 1) = 5k + 3: the even and the odd
//                      selected rows are each a uniform 800-byte pitch that also runs across frames (33,600 = 42 x 800),
//                      so ONE rank-3 tensor map per parity {160 B, 42 rows, frames} with a box of {160, 42, 2} lands the
//                      42 rows of both frames of a pair.  (First version: 168 bulk copies of 160 B per environment --
//                      2.4x SLOWER than the kernel above, ~60 cycles of TMA issue per copy; measured, tools/gpu_round2_t.sh.)
//   warps 2-9          wait `full`, turn the stage's stack into the next stack IN PLACE ((prev >> 8) | max(f0, f1) << 24 per
//                      pixel word; column table and row offsets live in registers for the whole kernel),
//                      fence.proxy.async, arrive on `done`
//   warp 1 (storer)    wait `done`, ONE 28,224-byte bulk store shared -> global, wait until the engine has read the stage,
//                      arrive on `empty`
// No thread touches global memory on the fast path (resets -- rare, 4 pairs per environment -- are gathered straight from
// global memory by the consumer warps into the same stage).  L2 policy (PAACB_K1_HINTS): bit 0 = the frame rows and the
// old stack are read evict-first (never used again), bit 1 = the new stack is written evict-last.
// The launcher falls back to the kernel above when the row table is not the affine pattern.
// ------------------------------------------------------------------------------------------------------------------
constexpr int kK1Stages = 4;
constexpr int kStackBytes = kStateVec * 16;                            // 28,224
constexpr int kRowsBytes = 2 * PAACB_OBS * PAACB_FRAME_W;              // 26,880 = two boxes of 13,440
constexpr int kBoxBytes = kRowsBytes / 2;                              // one parity: [frame][42][160]
constexpr int kBoxFrame = kBoxBytes / 2;                               // 6,720
constexpr int kStageBytes = (kRowsBytes + kStackBytes + 127) / 128 * 128;   // 55,168: [even box][odd box][stack]
constexpr int kK1Consumers = 256;
constexpr int kK1PipeThreads = 64 + kK1Consumers;
constexpr int kK1PipeSmem = kK1Stages * kStageBytes + 3 * kK1Stages * 8 + 16;
static_assert(PAACB_FRAME_H * PAACB_FRAME_W == 42 * 800 && PAACB_OBS == 84, "row-parity tensor maps assume 210 x 160 -> 84 rows");

struct K1PipeParams {
  CUtensorMap tm_even, tm_odd;
  const uint8_t* frames;
  const uint8_t* reset;
  const uint8_t* prev;
  uint8_t* next;
  int64_t n_envs;
  int pairs, hints;
  StepScalars sc;
  ResizeTables tabs;
};

__device__ __forceinline__ uint64_t l2_policy(int kind) {   // 0 normal, 1 evict-first, 2 evict-last
  uint64_t pol;
  if (kind == 1) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  else if (kind == 2) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  else asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void bulk_g2s_hint(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar, uint64_t pol) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(pol)
               : "memory");
}
__device__ __forceinline__ void tma_load_3d_hint(void* dst, const void* tmap, int c0, int c1, int c2, uint64_t* bar, uint64_t pol) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4, %5}], [%2], %6;" ::"r"(
          smem_u32(dst)),
      "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "l"(pol)
      : "memory");
}
__device__ __forceinline__ void bulk_s2g_hint(void* dst_gmem, const void* src_smem, uint32_t bytes, uint64_t pol) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(dst_gmem),
               "r"(smem_u32(src_smem)), "r"(bytes), "l"(pol)
               : "memory");
}

__global__ void __launch_bounds__(kK1PipeThreads, 1) preprocess_u8_pipe_kernel(const __grid_constant__ K1PipeParams p) {
  extern __shared__ __align__(128) uint8_t k1_smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(k1_smem + kK1Stages * kStageBytes);
  uint64_t* done = full + kK1Stages;
  uint64_t* empty = done + kK1Stages;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int s = 0; s < kK1Stages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&done[s], kK1Consumers);
      mbar_init(&empty[s], 1);
    }
    fence_barrier_init();
  }
  __syncthreads();

  if (warp == 0) {
    // ---- producer ----
    if (lane == 0) {
      tma_prefetch_desc(&p.tm_even);
      tma_prefetch_desc(&p.tm_odd);
      const uint64_t pol = l2_policy((p.hints & 1) ? 1 : 0);
      int k = 0;
      for (int64_t env = blockIdx.x; env < p.n_envs; env += gridDim.x, ++k) {
        const int s = k % kK1Stages;
        const uint32_t ph = (uint32_t)(k / kK1Stages) & 1u;
        if (k >= kK1Stages) mbar_wait(&empty[s], ph ^ 1u);
        uint8_t* stage = k1_smem + s * kStageBytes;
        mbar_arrive_expect_tx(&full[s], kRowsBytes + kStackBytes);
        const int frame0 = (int)(env * p.pairs * 2);
        tma_load_3d_hint(stage, &p.tm_even, 0, 0, frame0, &full[s], pol);
        tma_load_3d_hint(stage + kBoxBytes, &p.tm_odd, 0, 0, frame0, &full[s], pol);
        bulk_g2s_hint(stage + kRowsBytes, p.prev + env * (int64_t)kStackBytes, kStackBytes, &full[s], pol);
      }
    }
  } else if (warp == 1) {
    // ---- storer ----
    if (lane == 0) {
      const uint64_t pol = l2_policy((p.hints & 2) ? 2 : 0);
      int k = 0;
      for (int64_t env = blockIdx.x; env < p.n_envs; env += gridDim.x, ++k) {
        const int s = k % kK1Stages;
        const uint32_t ph = (uint32_t)(k / kK1Stages) & 1u;
        mbar_wait(&done[s], ph);
        bulk_s2g_hint(p.next + env * (int64_t)kStackBytes, k1_smem + s * kStageBytes + kRowsBytes, kStackBytes, pol);
        tma_store_commit();
        tma_store_wait_read();
        mbar_arrive(&empty[s]);
      }
      tma_store_wait_all();
    }
  } else {
    // ---- consumers ----
    const int ctid = tid - 64;
    const StepScalars& sc = p.sc;
    int roff[kVecPerThread];          // byte offset of the thread's plane row inside the stage (frame 0 of its parity box)
    uint32_t cols[kVecPerThread];     // the four source columns of its four pixels, one byte each
#pragma unroll
    for (int i = 0; i < kVecPerThread; ++i) {
      const int idx = ctid + i * kK1Consumers;
      const int y = idx < kStateVec ? idx / (PAACB_OBS / 4) : 0;
      const int x0 = idx < kStateVec ? (idx - y * (PAACB_OBS / 4)) * 4 : 0;
      roff[i] = (y & 1) * kBoxBytes + (y >> 1) * PAACB_FRAME_W;
      cols[i] = (uint32_t)p.tabs.col[x0] | ((uint32_t)p.tabs.col[x0 + 1] << 8) | ((uint32_t)p.tabs.col[x0 + 2] << 16) |
                ((uint32_t)p.tabs.col[x0 + 3] << 24);
    }
    int k = 0;
    for (int64_t env = blockIdx.x; env < p.n_envs; env += gridDim.x, ++k) {
      const int s = k % kK1Stages;
      const uint32_t ph = (uint32_t)(k / kK1Stages) & 1u;
      // the step's scalars and the reset flag: loaded before the wait, consumed after it
      bool rst = false;
      if (p.pairs >= PAACB_STACK) {
        if (p.reset != nullptr) rst = p.reset[env] != 0;
        if (sc.over_in != nullptr && sc.over_is_reset) rst = rst || (sc.over_in[env] != 0.f);
      }
      if (ctid == 0 && sc.over_in != nullptr) {
        sc.over_out[env] = sc.over_in[env];
        sc.rewards_out[env] = sc.rewards_in[env];
      }
      uint8_t* stage = k1_smem + s * kStageBytes;
      uint4* st4 = reinterpret_cast<uint4*>(stage + kRowsBytes);
      mbar_wait(&full[s], ph);
      if (!rst) {
#pragma unroll
        for (int i = 0; i < kVecPerThread; ++i) {
          const int idx = ctid + i * kK1Consumers;
          if (idx < kStateVec) {
            uint4 v = shr8(st4[idx]);
            const uint8_t* a = stage + roff[i];
            const uint8_t* b = a + kBoxFrame;
            const uint32_t c0 = cols[i] & 0xffu, c1 = (cols[i] >> 8) & 0xffu, c2 = (cols[i] >> 16) & 0xffu, c3 = cols[i] >> 24;
            v.x |= (uint32_t)max(a[c0], b[c0]) << 24;
            v.y |= (uint32_t)max(a[c1], b[c1]) << 24;
            v.z |= (uint32_t)max(a[c2], b[c2]) << 24;
            v.w |= (uint32_t)max(a[c3], b[c3]) << 24;
            st4[idx] = v;
          }
        }
      } else {
        // reset: a fresh stack from the environment's four pairs (oldest first), gathered from global memory
#pragma unroll 1
        for (int i = 0; i < kVecPerThread; ++i) {
          const int idx = ctid + i * kK1Consumers;
          if (idx < kStateVec) {
            const int y = idx / (PAACB_OBS / 4);
            const int src_row = (int)p.tabs.row[y] * PAACB_FRAME_W;
            uint32_t w0 = 0u, w1 = 0u, w2 = 0u, w3 = 0u;
            const uint32_t c0 = cols[i] & 0xffu, c1 = (cols[i] >> 8) & 0xffu, c2 = (cols[i] >> 16) & 0xffu, c3 = cols[i] >> 24;
            for (int q = 0; q < PAACB_STACK; ++q) {
              const uint8_t* g0 = p.frames + ((env * p.pairs + q) * 2) * (int64_t)kFrameBytes + src_row;
              const uint8_t* g1 = g0 + kFrameBytes;
              w0 |= (uint32_t)max(__ldg(g0 + c0), __ldg(g1 + c0)) << (8 * q);
              w1 |= (uint32_t)max(__ldg(g0 + c1), __ldg(g1 + c1)) << (8 * q);
              w2 |= (uint32_t)max(__ldg(g0 + c2), __ldg(g1 + c2)) << (8 * q);
              w3 |= (uint32_t)max(__ldg(g0 + c3), __ldg(g1 + c3)) << (8 * q);
            }
            st4[idx] = make_uint4(w0, w1, w2, w3);
          }
        }
      }
      fence_proxy_async();
      mbar_arrive(&done[s]);
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// SURVEY 8(f) rank 1, frame-dedup rollout storage -- the measured half of the A/B.  K1 at its CONTRACT traffic:
// the two frames' 84 selected rows in (26,880 B), ONE new 84 x 84 plane out (7,056 B) into a planar ring
// uint8 [n_envs, ring_slots, 84, 84]; a rollout of T steps then needs T + 3 planes per environment instead of T stacked
// [84, 84, 4] copies (paac.py:92,112 stores every frame four times).  stack_from_planes_kernel rebuilds the NHWC stack the
// conv1 kernels consume from four ring slots (oldest first) -- the cost a planar ring adds back for as long as conv1's
// int8 implicit GEMM needs its 16-byte (4 pixels x 4 frames) units.  tools/microbench_cfg5.py times both against K1.
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
preprocess_planar_u8_kernel(const uint8_t* __restrict__ frames, int pairs, uint8_t* __restrict__ ring, int ring_slots, int slot,
                            ResizeTables tabs, int64_t n_envs) {
  __shared__ __align__(16) uint8_t plane[PAACB_OBS * PAACB_FRAME_W];
  __shared__ int s_row[PAACB_OBS];
  __shared__ int s_col[PAACB_OBS];
  const int tid = threadIdx.x;
  if (tid < PAACB_OBS) {
    s_row[tid] = tabs.row[tid] * PAACB_FRAME_W;
    s_col[tid] = tabs.col[tid];
  }
  for (int64_t env = blockIdx.x; env < n_envs; env += gridDim.x) {
    const uint8_t* f0 = frames + (env * pairs * 2) * (int64_t)kFrameBytes;
    const uint8_t* f1 = f0 + kFrameBytes;
    __syncthreads();
    for (int i = tid; i < PAACB_OBS * kRowVec; i += kThreads) {
      const int y = i / kRowVec, c = i - y * kRowVec;
      const int src = s_row[y] + c * 16;
      const uint4 a = __ldg(reinterpret_cast<const uint4*>(f0 + src));
      const uint4 b = __ldg(reinterpret_cast<const uint4*>(f1 + src));
      uint4 m;
      m.x = __vmaxu4(a.x, b.x); m.y = __vmaxu4(a.y, b.y); m.z = __vmaxu4(a.z, b.z); m.w = __vmaxu4(a.w, b.w);
      *reinterpret_cast<uint4*>(plane + y * PAACB_FRAME_W + c * 16) = m;
    }
    __syncthreads();
    // 84 x 84 bytes = 441 uint4 per plane: 16 output pixels per thread-iteration
    uint4* out = reinterpret_cast<uint4*>(ring + (env * ring_slots + slot) * (int64_t)(PAACB_OBS * PAACB_OBS));
    for (int i = tid; i < PAACB_OBS * PAACB_OBS / 16; i += kThreads) {
      uint32_t w[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint32_t v = 0;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int px = i * 16 + j * 4 + e;
          const int y = px / PAACB_OBS, x = px - y * PAACB_OBS;
          v |= (uint32_t)plane[y * PAACB_FRAME_W + s_col[x]] << (8 * e);
        }
        w[j] = v;
      }
      out[i] = make_uint4(w[0], w[1], w[2], w[3]);
    }
  }
}

__global__ void __launch_bounds__(kThreads)
stack_from_planes_kernel(const uint8_t* __restrict__ ring, int ring_slots, int newest_slot, uint8_t* __restrict__ next, int64_t n_envs) {
  const int tid = threadIdx.x;
  for (int64_t env = blockIdx.x; env < n_envs; env += gridDim.x) {
    const uint8_t* base = ring + env * ring_slots * (int64_t)(PAACB_OBS * PAACB_OBS);
    const uint32_t* pl[PAACB_STACK];
#pragma unroll
    for (int c = 0; c < PAACB_STACK; ++c)        // channel c = the frame 3 - c steps back, oldest first
      pl[c] = reinterpret_cast<const uint32_t*>(base + ((newest_slot - 3 + c + 4 * ring_slots) % ring_slots) * (int64_t)(PAACB_OBS * PAACB_OBS));
    uint4* out = reinterpret_cast<uint4*>(next + env * (int64_t)(kStateVec * 16));
    for (int i = tid; i < kStateVec; i += kThreads) {       // 4 pixels = one 32-bit word of each plane -> one uint4 of the stack
      const uint32_t a = __ldg(pl[0] + i), b = __ldg(pl[1] + i), c = __ldg(pl[2] + i), d = __ldg(pl[3] + i);
      // [a0 b0 a1 b1], [c0 d0 c1 d1] ... then pixel words [a b c d] (channel 0 = oldest frame in the low byte)
      const uint32_t ab01 = __byte_perm(a, b, 0x5140), cd01 = __byte_perm(c, d, 0x5140);
      const uint32_t ab23 = __byte_perm(a, b, 0x7362), cd23 = __byte_perm(c, d, 0x7362);
      uint4 o;
      o.x = __byte_perm(ab01, cd01, 0x5410);
      o.y = __byte_perm(ab01, cd01, 0x7632);
      o.z = __byte_perm(ab23, cd23, 0x5410);
      o.w = __byte_perm(ab23, cd23, 0x7632);
      out[i] = o;
    }
  }
}

int launch_preprocess_planar(const paacb_ctx* ctx, const uint8_t* frames, int pairs, uint8_t* ring, int ring_slots, int slot,
                             int64_t n, cudaStream_t st) {
  if (n == 0) return PAACB_OK;
  PAACB_LAUNCH_BEGIN(ctx, K_PREPROCESS, st);
  preprocess_planar_u8_kernel<<<(unsigned)n, kThreads, 0, st>>>(frames, pairs, ring, ring_slots, slot, ctx->tabs, n);
  PAACB_LAUNCH_END(ctx, K_PREPROCESS, st);
  return PAACB_OK;
}

int launch_stack_from_planes(const paacb_ctx* ctx, const uint8_t* ring, int ring_slots, int newest_slot, uint8_t* next, int64_t n,
                             cudaStream_t st) {
  if (n == 0) return PAACB_OK;
  PAACB_LAUNCH_BEGIN(ctx, K_PREPROCESS, st);
  stack_from_planes_kernel<<<(unsigned)n, kThreads, 0, st>>>(ring, ring_slots, newest_slot, next, n);
  PAACB_LAUNCH_END(ctx, K_PREPROCESS, st);
  return PAACB_OK;
}

int launch_preprocess(const paacb_ctx* ctx, const uint8_t* frames, int pairs, const uint8_t* reset,
                      const uint8_t* prev, uint8_t* next, int64_t n, const StepScalars& sc, cudaStream_t st) {
  if (n == 0) return PAACB_OK;
  // Two kernels.  More than one environment per SM: the persistent TMA copy pipeline (above) -- one CTA per SM for device-resident
  // frames (HBM-bound), PAACB_K1_PIPE_HOST_GRID CTAs for frames in pinned, mapped HOST memory (the runners' buffers, read
  // zero-copy: PCIe-bound, few loads in flight are enough, and the other SMs stay free for the tensor-core kernels of the next
  // environment slice).  Small batches, or a row table that is not the two-pitch pattern: the CTA-per-environment kernel
  // (host frames: as a narrow grid-stride grid, PAACB_K1_HOST_GRID).
  unsigned grid = (unsigned)n;
  // is the frame buffer host memory?  The answer is cached per 2 MB-aligned address range the caller has used (a runner's
  // buffer is registered once and then read every step: no driver query on the hot path)
  int is_host = -1;
  const uintptr_t key = (uintptr_t)frames >> 21;
  for (int i = 0; i < ctx->k1_cache_n; ++i)
    if (ctx->k1_cache_key[i] == key) { is_host = ctx->k1_cache_host[i]; break; }
  if (is_host < 0) {
    cudaPointerAttributes attr;
    is_host = 0;
    if (cudaPointerGetAttributes(&attr, frames) == cudaSuccess) is_host = (attr.type == cudaMemoryTypeHost) ? 1 : 0;
    else cudaGetLastError();
    const int slot = ctx->k1_cache_n < paacb_ctx::kK1Cache ? ctx->k1_cache_n++ : (int)(key % paacb_ctx::kK1Cache);
    ctx->k1_cache_key[slot] = key;
    ctx->k1_cache_host[slot] = is_host;
  }
  if (is_host) {
    unsigned narrow = (unsigned)ctx->k1_host_grid;
    if (n > narrow) grid = narrow;
  }
  bool affine = true;       // row(2k) = row(0) + 5k, row(2k + 1) = row(1) + 5k: what the two row-parity tensor maps can express
  for (int y = 0; y < PAACB_OBS; ++y) affine = affine && ((int)ctx->tabs.row[y] == (int)ctx->tabs.row[y & 1] + 5 * (y >> 1));
  // (a CTA of the pipeline needs at least two environments to overlap anything: batches up to one environment per SM -- the
  // reference's default 32 -- keep the CTA-per-environment kernel)
  // PAACB_K1_PIPE=2 (default): pinned host frames go through the pipeline too, on PAACB_K1_PIPE_HOST_GRID CTAs -- the TMA engine's
  // row reads cross PCIe at 43.1 GB/s against 42.0 for the threads' 16-byte loads (tools/experiments/pcie_probe.py), 32 SMs are
  // held instead of 96, and the end-to-end cycle went 16.37 -> 15.8-15.95 ms (tools/gpu_round2_y.sh)
  const bool pipe_host = is_host && ctx->k1_pipe == 2;
  if ((!is_host || pipe_host) && ctx->k1_pipe && affine && n > (int64_t)ctx->num_sms &&
      (((uintptr_t)frames | (uintptr_t)prev | (uintptr_t)next) & 15) == 0 && n * pairs * 2 < (1LL << 31)) {
    static DeviceOnce attr_once;
    if (!attr_once.done(ctx->device)) {
      cudaError_t e = cudaFuncSetAttribute(preprocess_u8_pipe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kK1PipeSmem);
      if (e != cudaSuccess) {
        set_error("launch_preprocess: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        return PAACB_ECUDA;
      }
      attr_once.mark(ctx->device);
    }
    K1PipeParams p;
    memset(&p, 0, sizeof(p));
    const uint64_t dims[3] = {(uint64_t)PAACB_FRAME_W, 42u, (uint64_t)(n * pairs * 2)};
    const uint64_t strides[2] = {5u * PAACB_FRAME_W, (uint64_t)kFrameBytes};
    const uint32_t box[3] = {(uint32_t)PAACB_FRAME_W, 42u, 2u};
    int rc = encode_tmap(&p.tm_even, frames + (int)ctx->tabs.row[0] * PAACB_FRAME_W, 1, 3, dims, strides, box, 0);
    if (rc == PAACB_OK) rc = encode_tmap(&p.tm_odd, frames + (int)ctx->tabs.row[1] * PAACB_FRAME_W, 1, 3, dims, strides, box, 0);
    if (rc != PAACB_OK) return rc;
    p.frames = frames; p.reset = reset; p.prev = prev; p.next = next;
    p.n_envs = n; p.pairs = pairs; p.hints = ctx->k1_hints; p.sc = sc; p.tabs = ctx->tabs;
    const unsigned pgrid = (unsigned)((pipe_host && ctx->k1_pipe_host_grid < ctx->num_sms) ? ctx->k1_pipe_host_grid : ctx->num_sms);
    PAACB_LAUNCH_BEGIN(ctx, K_PREPROCESS, st);
    preprocess_u8_pipe_kernel<<<pgrid, kK1PipeThreads, kK1PipeSmem, st>>>(p);
    PAACB_LAUNCH_END(ctx, K_PREPROCESS, st);
    return PAACB_OK;
  }
  PAACB_LAUNCH_BEGIN(ctx, K_PREPROCESS, st);
  preprocess_u8_kernel<<<grid, kThreads, 0, st>>>(frames, pairs, reset, prev, next, ctx->tabs, n, sc);
  PAACB_LAUNCH_END(ctx, K_PREPROCESS, st);
  return PAACB_OK;
}

}  // namespace paacb
