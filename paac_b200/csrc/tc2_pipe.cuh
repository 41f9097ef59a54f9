// Layer-pipelined forward (tc2_pipe.cu): hand-off counters between the layers of ONE persistent kernel.
//
// The forward of a batch is four dependent implicit GEMMs.  Launched one after the other, every layer pays its own pipeline
// fill / drain and weight staging (measured: 11-14 us of a 60-70 us launch at 4096 samples), writes its activations to HBM
// and the next launch reads them back.  In the pipelined forward the SMs are PARTITIONED between the layers instead: CTAs
// [0, s1) run conv1, [s1, s1 + s2) conv2, ... each with the persistent loop of its stand-alone kernel, and a consumer's TMA
// producer waits until the samples its next tile reads have been written by the layer above.  The hand-off goes through
// global memory -- which here means the 126 MB L2: the consumer follows the producer by a few tiles, so its operand loads
// are L2 hits.  Progress is tracked by counters in device memory:
//     done[n >> shift] += valid output positions of sample n stored by this warp      (writer: epilogue warps)
//     wait until done[i] >= target for every sample (block) a tile reads                (reader: the TMA producer lane)
// Deadlock freedom: a role only ever waits for roles with LOWER CTA indices, CTAs are dispatched in index order and all of
// them fit on the device (grid <= SM count, one CTA per SM), so whatever a CTA waits for is running or finished.
#pragma once
#include <stdint.h>

namespace paacb {

struct PipeIO {
  const uint32_t* up;        // counters of the layer above (nullptr: no dependency)
  uint32_t up_target;        // positions per complete sample of the layer above
  uint32_t* done;            // this layer's counters
  int done_shift;            // counter index = sample >> done_shift (0: per sample; 7: per 128 samples for the fc layer above)
  uint32_t* err;             // set to 1 when a wait gives up (a bug, never load: the wait is bounded so that nothing can hang the GPU)
};

__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// generic-proxy global writes of other SMs (made visible by the acquire) -> ordered before this thread's async-proxy (TMA) reads
__device__ __forceinline__ void fence_proxy_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }

// one lane: wait until counters [lo, hi] have all reached `target`
__device__ __forceinline__ void pipe_wait(const PipeIO& io, int lo, int hi, uint32_t target) {
  for (int i = lo; i <= hi; ++i) {
    uint32_t spins = 0;
    while (ld_acquire_gpu(io.up + i) < target) {
      __nanosleep(40);
      if (++spins > (1u << 24)) {          // ~ seconds: the layer above is not coming
        *io.err = 1u;
        break;
      }
    }
  }
  fence_proxy_async_global();
}
// same, with a per-counter target (the last block of the batch holds fewer samples)
__device__ __forceinline__ void pipe_wait_one(const PipeIO& io, int i, uint32_t target) { pipe_wait(io, i, i, target); }

// A signal is PUBLISHED one tile late: the release fence has to wait until the warp's stores have reached L2 (~1-2 us after
// they were issued), and an epilogue warp that fenced right after its stores spent more time in the fence than on the tile
// (first version: conv1's role ran 4x slower than its stand-alone kernel).  So a warp notes what it stored (pipe_note),
// goes on with the next tile's TMEM loads and arithmetic, and publishes the note just before the next tile's stores
// (pipe_publish): by then the noted stores are long complete and the fence is cheap.  The last note is published after the loop.
struct PipeNote {
  int idx;              // first counter index (-1: nothing to publish)
  uint32_t ca, cb;      // positions for counter idx and idx + 1
};
__device__ __forceinline__ PipeNote pipe_note_none() { return PipeNote{-1, 0u, 0u}; }
// whole warp, after its global stores of one tile: `ok` = this lane stored a valid output position of sample n
__device__ __forceinline__ PipeNote pipe_note(const PipeIO& io, bool ok, int n) {
  const unsigned okm = __ballot_sync(0xffffffffu, ok);
  if (okm == 0u) return pipe_note_none();
  const int first = __ffs(okm) - 1;
  const int ia = __shfl_sync(0xffffffffu, n >> io.done_shift, first);      // a warp's rows span at most two counters
  const unsigned ma = __ballot_sync(0xffffffffu, ok && (n >> io.done_shift) == ia);
  return PipeNote{ia, (uint32_t)__popc(ma), (uint32_t)__popc(okm & ~ma)};
}
__device__ __forceinline__ void pipe_publish(const PipeIO& io, const PipeNote& nt, int lane) {
  if (nt.idx < 0) return;
  __syncwarp();                                  // every lane's stores are ordered before lane 0's release below
  if (lane == 0) {
#ifndef PAACB_PIPE_NOFENCE_EXPERIMENT
    fence_proxy_async_global();                  // the consumer reads these bytes through the TMA engine
    __threadfence();
#endif
    atomicAdd(io.done + nt.idx, nt.ca);
    if (nt.cb) atomicAdd(io.done + nt.idx + 1, nt.cb);
  }
}

}  // namespace paacb
