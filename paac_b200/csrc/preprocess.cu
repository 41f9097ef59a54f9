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
#include "common.cuh"

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
  // Frames in pinned, mapped HOST memory (the runners' buffers, read zero-copy): the kernel is PCIe-bound and needs only
  // ~100 KB of loads in flight, so it runs as a narrow grid-stride grid (one CTA on a subset of the SMs) that leaves
  // room for the persistent tensor-core kernels of the next environment slice to run beside it.  Device-resident frames:
  // HBM-bound, one CTA per environment.
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
  PAACB_LAUNCH_BEGIN(ctx, K_PREPROCESS, st);
  preprocess_u8_kernel<<<grid, kThreads, 0, st>>>(frames, pairs, reset, prev, next, ctx->tabs, n, sc);
  PAACB_LAUNCH_END(ctx, K_PREPROCESS, st);
  return PAACB_OK;
}

}  // namespace paacb
