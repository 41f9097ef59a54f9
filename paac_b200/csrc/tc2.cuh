// Shared declarations of the bf16-split tensor-core path (PAACB_MATH_BF16X3): tensor-map helper, workspace layout,
// per-layer geometry of the patch-resident implicit GEMMs.
#pragma once
#include <cuda.h>          // CUtensorMap and its enums only; cuTensorMapEncodeTiled is fetched from the driver at run time
#include "common.cuh"
#include "tc_ptx.cuh"

namespace paacb {

// ---- host: tensor maps ------------------------------------------------------------------------------
// rank-`rank` bf16 tensor, dims innermost first, strides in bytes for dims 1..rank-1, zero fill out of bounds.
int encode_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                     const uint32_t* box, int swizzle_bytes /* 0, 32, 64, 128 */);
int encode_tmap(CUtensorMap* out, const void* base, int elem_bytes /* 1: uint8, 2: bf16, 4: uint32 */, int rank, const uint64_t* dims,
                const uint64_t* strides_bytes, const uint32_t* box, int swizzle_bytes);

// ---- workspace layout (bytes) -------------------------------------------------------------------------
// forward / backward workspace, bf16-split mode: for layer l the region [out_act_off(l) * batch * 4, +E_l * batch * 4)
// holds the hi plane (E_l * batch bf16) followed by the lo plane; value = float(hi) + float(lo).
struct Planes {
  uint8_t* hi;
  uint8_t* lo;
};
__host__ __device__ inline Planes layer_planes(void* ws, int64_t off_floats_per_sample, int64_t elems_per_sample, int64_t batch) {
  Planes p;
  p.hi = reinterpret_cast<uint8_t*>(ws) + off_floats_per_sample * batch * 4;
  p.lo = p.hi + elems_per_sample * batch * 2;
  return p;
}
// A forward may fill samples [first, first + batch) of a workspace laid out for `cap` samples (paacb_policy_forward_at):
// the planes of a layer are sized by the capacity, the sample offset moves both plane pointers.
struct WsSlice {
  int64_t cap;      // samples the workspace was laid out for
  int64_t first;    // first sample this call writes
};
__host__ __device__ inline Planes layer_planes(void* ws, int64_t off_floats_per_sample, int64_t elems_per_sample, const WsSlice& s) {
  Planes p = layer_planes(ws, off_floats_per_sample, elems_per_sample, s.cap);
  p.hi += s.first * elems_per_sample * 2;
  p.lo += s.first * elems_per_sample * 2;
  return p;
}
// ReLU bit mask of conv1's output, after all activation planes of the forward workspace: per sample two planes of
// uint16 [OH1 * OW1] (400 words of 32 bits per sample in all) -- plane h holds, for every output position, the 16 flags
// "channel 16 h + c > 0" at bit relu1_bit_pos(c) (NIPS, 16 channels: plane 0 only).  conv1's forward epilogue has the values
// in registers and writes them (a warp's 32 positions are 64 contiguous bytes of one plane); conv2's data gradient reads
// 4 bytes per pixel pair and plane instead of two 64-byte hi-plane rows and can hold the NEXT tile's words in two registers.
// Bit position of channel c (0..15) inside a plane's uint16: the writer gathers the "non-zero" flags of its eight packed
// bf16x2 words with three instructions per word (tc2_conv1.cuh), which leaves channel 2j at bit 7 - j and 2j + 1 at bit 15 - j.
__host__ __device__ constexpr int relu1_bit_pos(int c) { return 8 * (c & 1) + 7 - (c >> 1); }
inline uint32_t* relu1_bits(void* ws, const paacb_ctx* ctx, const WsSlice& s) {
  return reinterpret_cast<uint32_t*>(ws) + ctx->act_floats_per_sample * s.cap + s.first * ctx->relu1_words_per_sample;
}
constexpr int64_t kStateElems = (int64_t)PAACB_OBS * PAACB_OBS * PAACB_STACK;   // 28,224

// ---- launchers (tc2_*.cu) ---------------------------------------------------------------------------
bool bf16x3_supported(const paacb_ctx* ctx);
int launch_pack_bf16_weights(const paacb_ctx* ctx, const float* params, cudaStream_t st);          // forward images
int launch_pack_bf16_dgrad_weights(const paacb_ctx* ctx, const float* params, cudaStream_t st);    // data-gradient images
int launch_conv_fwd_bf16(const paacb_ctx* ctx, int layer, const float* params, const uint8_t* states, void* fwd_ws, int64_t batch,
                         const WsSlice& slice, cudaStream_t st);
int launch_pack_conv1_i8(const paacb_ctx* ctx, const float* params, cudaStream_t st);              // int8 digit image of conv1
int launch_conv1_fwd_i8(const paacb_ctx* ctx, const float* params, const uint8_t* states, void* fwd_ws, int64_t batch,
                        const WsSlice& slice, cudaStream_t st);
int launch_conv1_fwd_i8_f32(const paacb_ctx* ctx, const float* params, const uint8_t* states, float* y, int64_t batch,
                            cudaStream_t st);                                                  // NIPS conv1 in the tf32 pipeline
int launch_fc_fwd_bf16(const paacb_ctx* ctx, int layer, const float* params, void* fwd_ws, int64_t batch, const WsSlice& slice,
                       cudaStream_t st);
int launch_conv_dgrad_bf16(const paacb_ctx* ctx, int layer, const void* fwd_ws, void* bwd_ws, float* grads, int64_t batch,
                           cudaStream_t st);
int launch_fc_dgrad_bf16(const paacb_ctx* ctx, int layer, const void* fwd_ws, void* bwd_ws, float* grads, int64_t batch,
                         cudaStream_t st);
int launch_conv_wgrad_bf16(const paacb_ctx* ctx, int layer, const uint8_t* states, const void* fwd_ws, const void* bwd_ws,
                           float* grads, int64_t batch, cudaStream_t st);
int launch_fc_wgrad_bf16(const paacb_ctx* ctx, int layer, const void* fwd_ws, const void* bwd_ws, float* grads,
                         int64_t batch, cudaStream_t st);

}  // namespace paacb
