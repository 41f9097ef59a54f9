// bf16-split tensor-core path, part 5: the LAYER-PIPELINED forward -- conv1, conv2, (conv3,) fc of one batch as ONE
// persistent kernel whose CTAs are partitioned between the layers (tc2_pipe.cuh has the hand-off protocol).
//
// Why: launched one after the other, each of the four layers pays its own fill / drain and weight staging (11-14 us of a
// 60-70 us launch at 4096 samples, tools/experiments/chunked_forward.py), its CTAs finish in waves, and every activation
// makes a round trip through HBM between two launches.  Here every layer's persistent loop runs for the whole forward on
// its share of the SMs, a consumer reads what its producer wrote a few tiles earlier out of the 126 MB L2, and the only
// fill / drain is the pipeline's own.  The activations are still written to the forward workspace exactly as the separate
// kernels write them (the backward reads them), and every sample is computed by the same instruction sequence: the results
// are bit-identical to the layer-by-layer forward (tests/test_gpu_pipe.py).
//
// Each role is the body of the stand-alone kernel of its layer (tc2_conv1.cuh, tc2_conv.cuh, tc2_stream.cuh) with the
// CTA index remapped into the role; 320 threads per CTA (conv1 runs with 8 epilogue warps here instead of 16: the register
// file has to hold the conv layers' 166-register epilogues).
#include "tc2_conv1.cuh"
#include "tc2_conv.cuh"
#include "tc2_stream.cuh"

namespace paacb {

struct PipeParams {
  Conv1Params c1;
  ConvKParams c2;
  ConvKParams c3;        // Nature only
  int c3_geo;            // G_FWD3 (input-grid tiles) or G_FWD3P (two samples per tile): the geometry the stand-alone launch would use
  StreamParams fc;
  int s1, s2, s3;        // CTAs of conv1, conv2, conv3 (NIPS: s3 = 0); the fc layer takes the rest of the grid
  uint32_t* cnt1;        // [batch]               conv1 output positions stored, per sample        (complete: 400)
  uint32_t* cnt2;        // [batch] / [batch/128]  conv2 ...                                        (complete: 81 per sample)
  uint32_t* cnt3;        // [batch/128]            conv3 ..., per block of 128 samples (Nature)     (complete: 49 per sample)
  uint32_t* err;         // sticky: a bounded wait gave up
};

constexpr int kPipeThreads = 320;

template <bool NATURE>
struct PipeCfg {
  static constexpr int NC1 = NATURE ? 32 : 16;
  static constexpr int G2 = NATURE ? G_FWD2 : G_FWD2N;
  static constexpr int SMEM_C1 = kC1_SMEM_FIXED + C1<NC1>::WBYTES;
  static constexpr int SMEM_C2 = ConvKCfg<G2>::SMEM_BYTES;
  static constexpr int SMEM_C3 = NATURE ? (ConvKCfg<G_FWD3>::SMEM_BYTES > ConvKCfg<G_FWD3P>::SMEM_BYTES ? ConvKCfg<G_FWD3>::SMEM_BYTES
                                                                                                         : ConvKCfg<G_FWD3P>::SMEM_BYTES) : 0;
  static constexpr int SMEM_FC = StreamCfg<kFcBN, ST_FWD>::SMEM_BYTES;
  static constexpr int M1 = SMEM_C1 > SMEM_C2 ? SMEM_C1 : SMEM_C2;
  static constexpr int M2 = SMEM_C3 > SMEM_FC ? SMEM_C3 : SMEM_FC;
  static constexpr int SMEM_BYTES = M1 > M2 ? M1 : M2;
  static_assert(ConvKCfg<G2>::THREADS == kPipeThreads && 64 + 32 * 8 == kPipeThreads && kStreamThreads <= kPipeThreads, "role thread counts");
  static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
};

template <bool NATURE>
__global__ void __launch_bounds__(kPipeThreads, 1) forward_pipe_kernel(const __grid_constant__ PipeParams p) {
  using Cfg = PipeCfg<NATURE>;
  extern __shared__ uint8_t smem_raw[];
  const int b = (int)blockIdx.x, grid = (int)gridDim.x;
  // roles in CTA-index order = dependency order (a role only waits for lower indices: no deadlock, tc2_pipe.cuh)
  if (b < p.s1) {
    const PipeIO io = {nullptr, 0u, p.cnt1, 0, p.err};
    conv1_i8_body<Cfg::NC1, false, 1, true>(p.c1, b, p.s1, smem_raw, io);
  } else if (b < p.s1 + p.s2) {
    const PipeIO io = {p.cnt1, (uint32_t)(kC1_OH * kC1_OW), p.cnt2, NATURE ? 0 : 7, p.err};
    convk_body<Cfg::G2, true>(p.c2, b - p.s1, p.s2, smem_raw, io);
  } else if (NATURE && b < p.s1 + p.s2 + p.s3) {
    if constexpr (NATURE) {
      const PipeIO io = {p.cnt2, (uint32_t)(Geo<G_FWD2>::OH * Geo<G_FWD2>::OW), p.cnt3, 7, p.err};
      if (p.c3_geo == G_FWD3P) convk_body<G_FWD3P, true>(p.c3, b - p.s1 - p.s2, p.s3, smem_raw, io);
      else convk_body<G_FWD3, true>(p.c3, b - p.s1 - p.s2, p.s3, smem_raw, io);
    }
  } else {
    const int first = p.s1 + p.s2 + p.s3;
    const PipeIO io = {NATURE ? p.cnt3 : p.cnt2, (uint32_t)(NATURE ? Geo<G_FWD3>::OH * Geo<G_FWD3>::OW : Geo<G_FWD2N>::OH * Geo<G_FWD2N>::OW),
                       nullptr, 0, p.err};
    stream_gemm_body<kFcBN, ST_FWD, true>(p.fc, b - first, grid - first, smem_raw, io);
  }
}

template <bool NATURE>
static int launch_pipe(const paacb_ctx* ctx, const PipeParams& p, cudaStream_t st) {
  using Cfg = PipeCfg<NATURE>;
  static DeviceOnce attr_set;
  if (!attr_set.done(ctx->device)) {
    if (cudaFuncSetAttribute(forward_pipe_kernel<NATURE>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES) != cudaSuccess) {
      cudaGetLastError();
      set_error("forward_pipe: cannot set %d bytes of dynamic shared memory", Cfg::SMEM_BYTES);
      return PAACB_ECUDA;
    }
    cudaFuncSetAttribute(forward_pipe_kernel<NATURE>, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
    cudaGetLastError();
    attr_set.mark(ctx->device);
  }
  PAACB_LAUNCH_BEGIN(ctx, K_FWD_PIPE, st);
  forward_pipe_kernel<NATURE><<<(unsigned)ctx->num_sms, kPipeThreads, Cfg::SMEM_BYTES, st>>>(p);
  PAACB_LAUNCH_END(ctx, K_FWD_PIPE, st);
  return PAACB_OK;
}

// counters of one in-flight forward: one buffer per stream that has used the pipelined forward (calls on one stream are
// ordered; a second stream gets its own buffer; with every buffer taken the caller falls back to the layer-by-layer forward)
static uint32_t* pipe_buffer_for(const paacb_ctx* ctx, cudaStream_t st) {
  for (int i = 0; i < ctx->pipe_nbuf; ++i)
    if (ctx->pipe_stream[i] == (void*)st) return ctx->pipe_cnt + (size_t)i * ctx->pipe_buf_words;
  if (ctx->pipe_nbuf >= paacb_ctx::kPipeBufs) return nullptr;
  ctx->pipe_stream[ctx->pipe_nbuf] = (void*)st;
  return ctx->pipe_cnt + (size_t)(ctx->pipe_nbuf++) * ctx->pipe_buf_words;
}

bool forward_pipe_usable(const paacb_ctx* ctx, int64_t batch) {
  return ctx->pipe_on && ctx->pipe_cnt != nullptr && ctx->math == PAACB_MATH_BF16X3 && batch >= ctx->pipe_min_batch &&
         batch <= paacb_ctx::kPipeMaxBatch;
}

// conv1 ... fc of `batch` samples into the forward workspace; PAACB_EUNSUPPORTED: use the layer-by-layer launchers
int launch_forward_pipe_bf16(const paacb_ctx* ctx, const float* params, const uint8_t* states, void* fwd_ws, int64_t batch,
                             const WsSlice& slice, cudaStream_t st) {
  if (!forward_pipe_usable(ctx, batch)) return PAACB_EUNSUPPORTED;
  const bool nature = ctx->arch == PAACB_ARCH_NATURE;
  const int L = ctx->n_layers;
  PipeParams p;
  memset(&p, 0, sizeof(p));
  int rc = prepare_conv1_fwd_bf16(ctx, params, states, fwd_ws, batch, slice, &p.c1);
  int geo = 0;
  if (rc == PAACB_OK) rc = prepare_conv_fwd_bf16(ctx, 1, params, fwd_ws, batch, slice, &p.c2, &geo);
  if (rc == PAACB_OK && geo != (nature ? G_FWD2 : G_FWD2N)) rc = PAACB_EUNSUPPORTED;
  if (rc == PAACB_OK && nature) {
    rc = prepare_conv_fwd_bf16(ctx, 2, params, fwd_ws, batch, slice, &p.c3, &geo);
    if (rc == PAACB_OK && geo != G_FWD3 && geo != G_FWD3P) rc = PAACB_EUNSUPPORTED;
    p.c3_geo = geo;
  }
  if (rc == PAACB_OK) rc = prepare_fc_fwd_bf16(ctx, L - 1, params, fwd_ws, batch, slice, &p.fc);
  if (rc != PAACB_OK) return rc;
  uint32_t* buf = pipe_buffer_for(ctx, st);
  if (buf == nullptr) return PAACB_EUNSUPPORTED;
  const size_t blocks = (size_t)((batch + 127) / 128);
  p.cnt1 = buf;
  p.cnt2 = buf + batch;
  p.cnt3 = p.cnt2 + (nature ? (size_t)batch : blocks + 1);
  const size_t words = (size_t)(p.cnt3 - buf) + (nature ? blocks + 1 : 0);
  p.err = ctx->pipe_err;
  // role sizes: the configured split, scaled to the SMs of this device and to roles that have at least one tile
  const int sms = ctx->num_sms;
  int s1 = ctx->pipe_split[0], s2 = ctx->pipe_split[1], s3 = nature ? ctx->pipe_split[2] : 0;
  if (s1 < 1) s1 = 1;
  if (s2 < 1) s2 = 1;
  if (nature && s3 < 1) s3 = 1;
  if (s1 + s2 + s3 > sms - 1) return PAACB_EUNSUPPORTED;
  p.s1 = s1; p.s2 = s2; p.s3 = s3;
  if (cudaMemsetAsync(buf, 0, words * sizeof(uint32_t), st) != cudaSuccess) {
    set_error("forward_pipe: cudaMemsetAsync failed: %s", cudaGetErrorString(cudaGetLastError()));
    return PAACB_ECUDA;
  }
  return nature ? launch_pipe<true>(ctx, p, st) : launch_pipe<false>(ctx, p, st);
}

}  // namespace paacb
