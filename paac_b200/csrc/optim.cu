// K10 + K11: tf.clip_by_global_norm followed by ApplyRMSProp on every variable, as two launches over
// ONE flat fp32 parameter buffer (the reference runs 10 L2Loss + 10 Mul + 10 ApplyRMSProp ops).
//
// Reference (file:line / SURVEY App. B op order):
//   norm  = sqrt(2 * sum_i L2Loss(g_i))                      actor_learner.py:56-57
//   scale = clip * min(1 / norm, 1 / clip)                   (Minimum, then Mul by clip)
//   g_i  <- g_i * scale
//   ms   <- ms + (g*g - ms) * (1 - rho)                      actor_learner.py:33-34,70 (ApplyRMSProp,
//   mom  <- momentum * mom + lr * g / sqrt(ms + eps)           TF-1.0 functor: epsilon inside the sqrt)
//   var  <- var - mom
// `gscale` (1/world_size) is applied to the raw gradient first so that a multi-GPU allreduce-SUM
// followed by this kernel equals the single-learner update on the concatenated batch.
//
// ONE cooperative launch (all CTAs co-resident, one grid barrier).  Every thread loads its slice of the gradient into
// registers and asks for its slice of params / ms / mom to be brought into L2 (prefetch.global.L2), reduces the gradient
// slice, writes one double partial per CTA, crosses the grid barrier, re-reduces the partials in a fixed order
// (deterministic, no atomics on the data path), and applies the update with the gradient still in registers.  The
// gradient is read ONCE (the two-launch version re-read it) and the DRAM latency of params / ms / mom is hidden behind
// the norm reduction and the barrier.  HBM traffic: 16 B read + 12 B written per parameter.  Parameters beyond the resident threads' register slices (only for
// networks far larger than PAAC's) take a strided second pass that re-reads the gradient.
// The two-launch version (sumsq_partials_kernel + clip_rmsprop_kernel) is kept for devices / contexts where the
// cooperative grid does not fit and as the reference the fused kernel is tested against (PAACB_OPT_TWO_PASS=1).
#include "common.cuh"

namespace paacb {

constexpr int kOptThreads = 256;
constexpr int kMaxPartials = 1024;

__global__ void __launch_bounds__(kOptThreads)
sumsq_partials_kernel(const float* __restrict__ g, int64_t P, float gscale, double* __restrict__ partials) {
  const int64_t nvec = P >> 2;
  const float4* g4 = reinterpret_cast<const float4*>(g);
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * kOptThreads + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * kOptThreads) {
    float4 t = __ldg(g4 + i);
    t.x *= gscale; t.y *= gscale; t.z *= gscale; t.w *= gscale;
    s0 = fmaf(t.x, t.x, s0); s1 = fmaf(t.y, t.y, s1); s2 = fmaf(t.z, t.z, s2); s3 = fmaf(t.w, t.w, s3);
  }
  double s = (double)s0 + (double)s1 + (double)s2 + (double)s3;
  if (blockIdx.x == 0) {
    for (int64_t i = (nvec << 2) + threadIdx.x; i < P; i += kOptThreads) {
      const float t = g[i] * gscale;
      s += (double)t * (double)t;
    }
  }
  __shared__ double red[kOptThreads / 32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < kOptThreads / 32; ++i) t += red[i];
    partials[blockIdx.x] = t;
  }
}

__device__ __forceinline__ void rmsprop_one(float& var, float& ms, float& mom, float g, float lr, float one_minus_rho,
                                            float eps, float momentum) {
  ms = ms + (g * g - ms) * one_minus_rho;
  mom = mom * momentum + (g * lr) / sqrtf(ms + eps);
  var = var - mom;
}

__global__ void __launch_bounds__(kOptThreads)
clip_rmsprop_kernel(float* __restrict__ params, float* __restrict__ ms, float* __restrict__ mom,
                    const float* __restrict__ g, int64_t P, float gscale, float lr, float rho, float eps,
                    float momentum, float clip, int clip_type, const double* __restrict__ partials, int npartials,
                    float* __restrict__ norm_out, const float* __restrict__ d_lr) {
  if (d_lr != nullptr) lr = __ldg(d_lr);       // learning rate from device memory (graph replay)
  // ---- deterministic re-reduction of the partial sums (same order in every CTA) ----
  __shared__ double red[kOptThreads / 32];
  __shared__ float s_scale;
  double s = 0.0;
  for (int i = threadIdx.x; i < npartials; i += kOptThreads) s += partials[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < kOptThreads / 32; ++i) t += red[i];
    const float norm = sqrtf((float)t);
    float scale = 1.0f;
    if (clip_type == PAACB_CLIP_GLOBAL) scale = clip * fminf(1.0f / norm, 1.0f / clip);
    s_scale = scale * gscale;
    if (blockIdx.x == 0 && norm_out != nullptr) *norm_out = norm;
  }
  __syncthreads();
  const float scale = s_scale;
  const float omr = 1.0f - rho;

  const int64_t nvec = P >> 2;
  float4* p4 = reinterpret_cast<float4*>(params);
  float4* ms4 = reinterpret_cast<float4*>(ms);
  float4* mom4 = reinterpret_cast<float4*>(mom);
  const float4* g4 = reinterpret_cast<const float4*>(g);
  for (int64_t i = (int64_t)blockIdx.x * kOptThreads + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * kOptThreads) {
    float4 gv = __ldg(g4 + i);
    float4 pv = p4[i], mv = ms4[i], ov = mom4[i];
    rmsprop_one(pv.x, mv.x, ov.x, gv.x * scale, lr, omr, eps, momentum);
    rmsprop_one(pv.y, mv.y, ov.y, gv.y * scale, lr, omr, eps, momentum);
    rmsprop_one(pv.z, mv.z, ov.z, gv.z * scale, lr, omr, eps, momentum);
    rmsprop_one(pv.w, mv.w, ov.w, gv.w * scale, lr, omr, eps, momentum);
    p4[i] = pv; ms4[i] = mv; mom4[i] = ov;
  }
  if (blockIdx.x == 0) {
    for (int64_t i = (nvec << 2) + threadIdx.x; i < P; i += kOptThreads) {
      float pv = params[i], mv = ms[i], ov = mom[i];
      rmsprop_one(pv, mv, ov, g[i] * scale, lr, omr, eps, momentum);
      params[i] = pv; ms[i] = mv; mom[i] = ov;
    }
  }
}

// ---- fused cooperative version ------------------------------------------------------------------------------
constexpr int kFusedVec = 4;      // float4 slices of the gradient per thread held in registers

// Self-resetting grid barrier on two words (arrival counter, generation): the last CTA to arrive zeroes the counter and
// bumps the generation, the others spin on the generation they read BEFORE arriving.  No host-side bookkeeping, no
// wrap-around, any grid size per launch: a captured CUDA graph of the update can be replayed indefinitely.
__device__ __forceinline__ void grid_barrier(unsigned int* bar, unsigned int nblocks) {
  __syncthreads();
  if (threadIdx.x == 0) {
    volatile unsigned int* gen_p = bar + 1;
    const unsigned int gen = *gen_p;
    __threadfence();
    const unsigned int ticket = atomicAdd(bar, 1u);
    if (ticket == nblocks - 1u) {
      *reinterpret_cast<volatile unsigned int*>(bar) = 0u;
      __threadfence();
      atomicAdd(bar + 1, 1u);
    } else {
      while (*gen_p == gen) __nanosleep(20);
    }
    __threadfence();
  }
  __syncthreads();
}

__global__ void __launch_bounds__(kOptThreads)
clip_rmsprop_fused_kernel(float* __restrict__ params, float* __restrict__ ms, float* __restrict__ mom,
                          const float* __restrict__ g, int64_t P, float gscale, float lr, float rho, float eps,
                          float momentum, float clip, int clip_type, double* __restrict__ partials,
                          unsigned int* __restrict__ counter, float* __restrict__ norm_out, const float* __restrict__ d_lr) {
  if (d_lr != nullptr) lr = __ldg(d_lr);       // learning rate from device memory (graph replay)
  const int64_t nvec = P >> 2;
  const int64_t nthreads = (int64_t)gridDim.x * kOptThreads;
  const int64_t t0 = (int64_t)blockIdx.x * kOptThreads + threadIdx.x;
  float4* p4 = reinterpret_cast<float4*>(params);
  float4* ms4 = reinterpret_cast<float4*>(ms);
  float4* mom4 = reinterpret_cast<float4*>(mom);
  const float4* g4 = reinterpret_cast<const float4*>(g);

  float4 gv[kFusedVec];
#pragma unroll
  for (int k = 0; k < kFusedVec; ++k) {
    const int64_t i = t0 + k * nthreads;
    gv[k] = (i < nvec) ? __ldcs(g4 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
#pragma unroll
  for (int k = 0; k < kFusedVec; ++k) {
    const int64_t i = t0 + k * nthreads;
    if (i < nvec && (threadIdx.x & 1) == 0) {        // one prefetch per 32-byte sector
      asm volatile("prefetch.global.L2 [%0];" ::"l"(p4 + i));
      asm volatile("prefetch.global.L2 [%0];" ::"l"(ms4 + i));
      asm volatile("prefetch.global.L2 [%0];" ::"l"(mom4 + i));
    }
  }
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
  for (int k = 0; k < kFusedVec; ++k) {
    const float tx = gv[k].x * gscale, ty = gv[k].y * gscale, tz = gv[k].z * gscale, tw = gv[k].w * gscale;
    s0 = fmaf(tx, tx, s0); s1 = fmaf(ty, ty, s1); s2 = fmaf(tz, tz, s2); s3 = fmaf(tw, tw, s3);
  }
  for (int64_t i = t0 + kFusedVec * nthreads; i < nvec; i += nthreads) {      // beyond the register slices
    float4 t = __ldg(g4 + i);
    t.x *= gscale; t.y *= gscale; t.z *= gscale; t.w *= gscale;
    s0 = fmaf(t.x, t.x, s0); s1 = fmaf(t.y, t.y, s1); s2 = fmaf(t.z, t.z, s2); s3 = fmaf(t.w, t.w, s3);
  }
  double s = (double)s0 + (double)s1 + (double)s2 + (double)s3;
  if (blockIdx.x == 0) {
    for (int64_t i = (nvec << 2) + threadIdx.x; i < P; i += kOptThreads) {
      const float t = g[i] * gscale;
      s += (double)t * (double)t;
    }
  }
  __shared__ double red[kOptThreads / 32];
  __shared__ float s_scale;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < kOptThreads / 32; ++i) t += red[i];
    partials[blockIdx.x] = t;
  }
  grid_barrier(counter, gridDim.x);

  // deterministic re-reduction of the partials (same order in every CTA)
  s = 0.0;
  for (int i = threadIdx.x; i < (int)gridDim.x; i += kOptThreads) s += __ldcg(partials + i);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < kOptThreads / 32; ++i) t += red[i];
    const float norm = sqrtf((float)t);
    float scale = 1.0f;
    if (clip_type == PAACB_CLIP_GLOBAL) scale = clip * fminf(1.0f / norm, 1.0f / clip);
    s_scale = scale;
    if (blockIdx.x == 0 && norm_out != nullptr) *norm_out = norm;
  }
  __syncthreads();
  const float sg = s_scale * gscale;      // one multiplier, exactly as in the two-pass kernel
  const float omr = 1.0f - rho;
#pragma unroll
  for (int k = 0; k < kFusedVec; ++k) {
    const int64_t i = t0 + k * nthreads;
    if (i < nvec) {
      float4 pv = p4[i], mv = ms4[i], ov = mom4[i];
      rmsprop_one(pv.x, mv.x, ov.x, gv[k].x * sg, lr, omr, eps, momentum);
      rmsprop_one(pv.y, mv.y, ov.y, gv[k].y * sg, lr, omr, eps, momentum);
      rmsprop_one(pv.z, mv.z, ov.z, gv[k].z * sg, lr, omr, eps, momentum);
      rmsprop_one(pv.w, mv.w, ov.w, gv[k].w * sg, lr, omr, eps, momentum);
      p4[i] = pv; ms4[i] = mv; mom4[i] = ov;
    }
  }
  for (int64_t i = t0 + kFusedVec * nthreads; i < nvec; i += nthreads) {
    float4 gg = __ldg(g4 + i);
    float4 p = p4[i], m = ms4[i], o = mom4[i];
    rmsprop_one(p.x, m.x, o.x, gg.x * sg, lr, omr, eps, momentum);
    rmsprop_one(p.y, m.y, o.y, gg.y * sg, lr, omr, eps, momentum);
    rmsprop_one(p.z, m.z, o.z, gg.z * sg, lr, omr, eps, momentum);
    rmsprop_one(p.w, m.w, o.w, gg.w * sg, lr, omr, eps, momentum);
    p4[i] = p; ms4[i] = m; mom4[i] = o;
  }
  if (blockIdx.x == 0) {
    for (int64_t i = (nvec << 2) + threadIdx.x; i < P; i += kOptThreads) {
      float p = params[i], m = ms[i], o = mom[i];
      rmsprop_one(p, m, o, g[i] * sg, lr, omr, eps, momentum);
      params[i] = p; ms[i] = m; mom[i] = o;
    }
  }
}

// ---- opt-in summaries (actor_learner.py:85-87 + logger_utils.py:23-33): sum, sum of squares, max and min of the flat
// gradient in one pass.  The clipped gradient is the raw one times a non-negative scalar, so its four statistics follow
// from these without a second pass.  Off the hot path: two small launches, only when the caller asks.
__global__ void __launch_bounds__(kOptThreads)
grad_stats_partials_kernel(const float* __restrict__ g, int64_t P, float gscale, double* __restrict__ partials) {
  double s = 0.0, q = 0.0;
  float mx = -INFINITY, mn = INFINITY;
  for (int64_t i = (int64_t)blockIdx.x * kOptThreads + threadIdx.x; i < P; i += (int64_t)gridDim.x * kOptThreads) {
    const float t = __ldg(g + i) * gscale;
    s += (double)t;
    q += (double)t * (double)t;
    mx = fmaxf(mx, t);
    mn = fminf(mn, t);
  }
  __shared__ double rs[kOptThreads / 32], rq[kOptThreads / 32];
  __shared__ float rmx[kOptThreads / 32], rmn[kOptThreads / 32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    q += __shfl_xor_sync(0xffffffffu, q, o);
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
  }
  if ((threadIdx.x & 31) == 0) { rs[threadIdx.x >> 5] = s; rq[threadIdx.x >> 5] = q; rmx[threadIdx.x >> 5] = mx; rmn[threadIdx.x >> 5] = mn; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < kOptThreads / 32; ++i) { s += rs[i]; q += rq[i]; mx = fmaxf(mx, rmx[i]); mn = fminf(mn, rmn[i]); }
    partials[4 * blockIdx.x] = s; partials[4 * blockIdx.x + 1] = q; partials[4 * blockIdx.x + 2] = mx; partials[4 * blockIdx.x + 3] = mn;
  }
}
__global__ void grad_stats_final_kernel(const double* __restrict__ partials, int n, double* __restrict__ out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double s = 0.0, q = 0.0, mx = -INFINITY, mn = INFINITY;
  for (int i = 0; i < n; ++i) {
    s += partials[4 * i]; q += partials[4 * i + 1];
    mx = fmax(mx, partials[4 * i + 2]); mn = fmin(mn, partials[4 * i + 3]);
  }
  out[0] = s; out[1] = q; out[2] = mx; out[3] = mn;
}

int launch_grad_stats(const paacb_ctx* ctx, const float* grads, float gscale, float* ws, double* out4, cudaStream_t st) {
  const int blocks = kMaxPartials / 4;           // 4 doubles per block inside the optimizer workspace
  double* partials = reinterpret_cast<double*>(ws);
  PAACB_LAUNCH_BEGIN(ctx, K_SUMSQ, st);
  grad_stats_partials_kernel<<<blocks, kOptThreads, 0, st>>>(grads, ctx->param_count, gscale, partials);
  PAACB_LAUNCH_END(ctx, K_SUMSQ, st);
  PAACB_LAUNCH_BEGIN(ctx, K_SUMSQ, st);
  grad_stats_final_kernel<<<1, 32, 0, st>>>(partials, blocks, out4);
  PAACB_LAUNCH_END(ctx, K_SUMSQ, st);
  return PAACB_OK;
}

int64_t optimizer_ws_floats(const paacb_ctx*) { return 2 * kMaxPartials + 4; }   // kMaxPartials doubles (+ spare)

static int opt_blocks(const paacb_ctx* ctx, int64_t P) {
  int64_t want = (P / 4 + kOptThreads * 4 - 1) / (kOptThreads * 4);   // >= 4 float4 per thread
  const int64_t cap = (int64_t)ctx->num_sms * 4 < kMaxPartials ? (int64_t)ctx->num_sms * 4 : kMaxPartials;
  if (want > cap) want = cap;
  if (want < 1) want = 1;
  return (int)want;
}

// cooperative grid of the fused kernel: as many CTAs as can be co-resident on the context's device (capped by the partials
// buffer); queried once per context
static int fused_blocks(const paacb_ctx* ctx) {
  if (ctx->opt_fused_blocks >= 0) return ctx->opt_fused_blocks;
  int per_sm = 0, coop = 0;
  if (cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, ctx->device) != cudaSuccess || !coop ||
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, clip_rmsprop_fused_kernel, kOptThreads, 0) != cudaSuccess) {
    cudaGetLastError();
    per_sm = 0;
  }
  int64_t blocks = (int64_t)per_sm * ctx->num_sms;
  if (blocks > kMaxPartials) blocks = kMaxPartials;
  ctx->opt_fused_blocks = (int)blocks;
  return ctx->opt_fused_blocks;
}

int launch_clip_rmsprop(const paacb_ctx* ctx, float* params, float* ms, float* mom, const float* grads, float gscale,
                        float lr, const float* d_lr, float rho, float eps, float momentum, float clip, int clip_type,
                        float* norm_out, float* ws, cudaStream_t st) {
  int64_t P = ctx->param_count;
  double* partials = reinterpret_cast<double*>(ws);
  const int fused = (ctx->opt_counter != nullptr && !ctx->opt_two_pass) ? fused_blocks(ctx) : 0;
  if (fused > 0) {
    // no more CTAs than the register slices need: do not spin CTAs that hold no data
    int64_t need = ((P >> 2) + kOptThreads * kFusedVec - 1) / (kOptThreads * kFusedVec);
    int blocks = (int)(need < fused ? (need < 1 ? 1 : need) : fused);
    unsigned int* counter = ctx->opt_counter;      // self-resetting barrier: no host-side state, replay-safe
    void* args[] = {&params, &ms, &mom, &grads, &P, &gscale, &lr, &rho, &eps, &momentum, &clip, &clip_type, &partials,
                    &counter, &norm_out, &d_lr};
    PAACB_LAUNCH_BEGIN(ctx, K_RMSPROP, st);
    const cudaError_t e = cudaLaunchCooperativeKernel((const void*)clip_rmsprop_fused_kernel, dim3((unsigned)blocks),
                                                      dim3(kOptThreads), args, 0, st);
    if (e != cudaSuccess) {
      cudaGetLastError();
      set_error("paacb_clip_rmsprop: cooperative launch failed: %s", cudaGetErrorString(e));
      return PAACB_ECUDA;
    }
    PAACB_LAUNCH_END(ctx, K_RMSPROP, st);
    return PAACB_OK;
  }
  const int blocks = opt_blocks(ctx, P);
  PAACB_LAUNCH_BEGIN(ctx, K_SUMSQ, st);
  sumsq_partials_kernel<<<blocks, kOptThreads, 0, st>>>(grads, P, gscale, partials);
  PAACB_LAUNCH_END(ctx, K_SUMSQ, st);
  PAACB_LAUNCH_BEGIN(ctx, K_RMSPROP, st);
  clip_rmsprop_kernel<<<blocks, kOptThreads, 0, st>>>(params, ms, mom, grads, P, gscale, lr, rho, eps, momentum, clip,
                                                      clip_type, partials, blocks, norm_out, d_lr);
  PAACB_LAUNCH_END(ctx, K_RMSPROP, st);
  return PAACB_OK;
}

}  // namespace paacb
