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
// Pass 1 writes one double partial sum of squares per CTA; pass 2 re-reduces the partials in a fixed
// order in every CTA (deterministic, no atomics, no host sync) and applies the update.
// HBM traffic: 4 B/param (pass 1) + 12 B read + 12 B written (pass 2) = 28 B/param.
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
                    float* __restrict__ norm_out) {
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

int64_t optimizer_ws_floats(const paacb_ctx*) { return 2 * kMaxPartials; }   // kMaxPartials doubles

static int opt_blocks(const paacb_ctx* ctx, int64_t P) {
  int64_t want = (P / 4 + kOptThreads * 4 - 1) / (kOptThreads * 4);   // >= 4 float4 per thread
  const int64_t cap = (int64_t)ctx->num_sms * 4 < kMaxPartials ? (int64_t)ctx->num_sms * 4 : kMaxPartials;
  if (want > cap) want = cap;
  if (want < 1) want = 1;
  return (int)want;
}

int launch_clip_rmsprop(const paacb_ctx* ctx, float* params, float* ms, float* mom, const float* grads, float gscale,
                        float lr, float rho, float eps, float momentum, float clip, int clip_type, float* norm_out,
                        float* ws, cudaStream_t st) {
  const int64_t P = ctx->param_count;
  const int blocks = opt_blocks(ctx, P);
  double* partials = reinterpret_cast<double*>(ws);
  PAACB_LAUNCH_BEGIN(ctx, K_SUMSQ, st);
  sumsq_partials_kernel<<<blocks, kOptThreads, 0, st>>>(grads, P, gscale, partials);
  PAACB_LAUNCH_END(ctx, K_SUMSQ, st);
  PAACB_LAUNCH_BEGIN(ctx, K_RMSPROP, st);
  clip_rmsprop_kernel<<<blocks, kOptThreads, 0, st>>>(params, ms, mom, grads, P, gscale, lr, rho, eps, momentum, clip,
                                                      clip_type, partials, blocks, norm_out);
  PAACB_LAUNCH_END(ctx, K_RMSPROP, st);
  return PAACB_OK;
}

}  // namespace paacb
