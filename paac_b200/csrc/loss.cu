// K7 + K8: n-step discounted returns / advantages fused with the entropy-regularised A2C loss and its
// closed-form gradient with respect to logits and value.
//
// Reference (file:line):
//   reward clip to [-1, 1]                         actor_learner.py:95-101, paac.py:123
//   mask_t = 1 - episode_over_t                    paac.py:119
//   R <- V(s_T); R <- r_t + gamma * R * mask_t     paac.py:144-149   (float64 NumPy, fed as float32)
//   y_t = R, adv_t = R - V(s_t)                    paac.py:148-149
//   log pi = log(pi + 1e-30); H = -sum pi log pi   policy_v_network.py:29-35
//   L = 5 * ( mean[-(log pi_a * adv + beta * H)] + mean[0.25 * (y - v)^2] )   policy_v_network.py:40-57
// and TF's autodiff of it (Log, Softmax, Pow gradients), SURVEY App. C:
//   dL/dpi_j = -adv [j == a] / (pi_j + eps) + beta (log(pi_j + eps) + pi_j / (pi_j + eps))   (x -1 folded in)
//   dL/dz    = (5 / B) * pi * (dL/dpi - sum_k dL/dpi_k pi_k)
//   dL/dv    = (5 * 0.25 * 2 / B) * (v - y)
// One CTA owns 32 environments.  The recurrence over t is a dependent chain of T fused multiply-adds per environment;
// everything around it is independent per (t, environment), so the CTA is organised around the memory latency:
//   phase 1  warp w loads r, over, V(s_t) of time steps t = w, w + 8, ... into shared memory AND issues the loads of
//            pi, v, action of its first time step into registers (all DRAM round trips of the kernel overlap)
//   phase 2  warp 0 walks t = T-1 .. 0 out of shared memory carrying R in a double register (the recurrence is float64
//            in the reference and its outputs are rounded once to float32) and writes y, adv
//   phase 3  warp w computes loss and gradient of its time steps (fp32), one lane per environment
// (The first version ran one thread per environment through all three phases: 5 serial DRAM round trips on 32 SMs.)
#include "common.cuh"

namespace paacb {

constexpr int kLossThreads = 256;
constexpr int kLossWarps = kLossThreads / 32;
constexpr int kLossMaxT = 64;            // 5 * 32 floats of shared memory per time step

__global__ void __launch_bounds__(kLossThreads)
returns_loss_grad_kernel(const float* __restrict__ rewards, const float* __restrict__ over,
                         const float* __restrict__ values, const float* __restrict__ boot,
                         const int32_t* __restrict__ actions, const float* __restrict__ pi,
                         const float* __restrict__ v, int T, int64_t N, int A, double gamma, float beta,
                         float* __restrict__ y, float* __restrict__ adv, float* __restrict__ dlogits,
                         float* __restrict__ dv, float* __restrict__ loss) {
  extern __shared__ float sm[];
  float* s_r = sm;                       // [T][32] clipped reward
  float* s_mask = s_r + T * 32;          // [T][32] 1 - episode_over
  float* s_val = s_mask + T * 32;        // [T][32] acting value V(s_t)
  float* s_y = s_val + T * 32;           // [T][32] critic target
  float* s_adv = s_y + T * 32;           // [T][32] advantage
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t n = (int64_t)blockIdx.x * 32 + lane;
  const bool live = n < N;
  const float invB = 1.0f / (float)((int64_t)T * N);
  const float eps = 1e-30f;

  // ---- phase 1 ----
  float pv[PAACB_MAX_ACTIONS];
  float vb0 = 0.f;
  int a0 = 0;
  if (live && warp < T) {               // first time step of this warp: operands of phase 3 into registers
    const int64_t b = (int64_t)warp * N + n;
    const float* p = pi + b * A;
#pragma unroll
    for (int j = 0; j < PAACB_MAX_ACTIONS; ++j) pv[j] = (j < A) ? p[j] : 0.f;
    vb0 = v[b];
    a0 = actions[b];
  }
  for (int t = warp; t < T; t += kLossWarps) {
    if (live) {
      const int64_t b = (int64_t)t * N + n;
      float r = rewards[b];
      r = r > 1.0f ? 1.0f : (r < -1.0f ? -1.0f : r);
      s_r[t * 32 + lane] = r;
      s_mask[t * 32 + lane] = 1.0f - over[b];
      s_val[t * 32 + lane] = values[b];
    }
  }
  __syncthreads();
  // ---- phase 2 ----
  if (warp == 0 && live) {
    double R = (double)boot[n];
    for (int t = T - 1; t >= 0; --t) {
      R = (double)s_r[t * 32 + lane] + gamma * R * (double)s_mask[t * 32 + lane];
      const float yt = (float)R;
      const float at = (float)(R - (double)s_val[t * 32 + lane]);
      s_y[t * 32 + lane] = yt;
      s_adv[t * 32 + lane] = at;
      const int64_t b = (int64_t)t * N + n;
      y[b] = yt;
      adv[b] = at;
    }
  }
  __syncthreads();
  // ---- phase 3 ----
  float lsum = 0.f;
  for (int t = warp; t < T; t += kLossWarps) {
    if (live) {
      const int64_t b = (int64_t)t * N + n;
      float vb = vb0;
      int a = a0;
      if (t != warp) {
        const float* p = pi + b * A;
#pragma unroll
        for (int j = 0; j < PAACB_MAX_ACTIONS; ++j) pv[j] = (j < A) ? p[j] : 0.f;
        vb = v[b];
        a = actions[b];
      }
      const float yt = s_y[t * 32 + lane], at = s_adv[t * 32 + lane];
      float H = 0.f, dot = 0.f, logsel = 0.f;
      float dpi[PAACB_MAX_ACTIONS];
#pragma unroll
      for (int j = 0; j < PAACB_MAX_ACTIONS; ++j) {
        if (j < A) {
          const float pj = pv[j];
          const float lp = logf(pj + eps);
          H -= pj * lp;
          float d = beta * (lp + pj / (pj + eps));
          if (j == a) { d -= at / (pj + eps); logsel = lp; }
          dpi[j] = d;
          dot = fmaf(d, pj, dot);
        }
      }
      const float diff = yt - vb;
      lsum += -(logsel * at + beta * H) + 0.25f * diff * diff;
#pragma unroll
      for (int j = 0; j < PAACB_MAX_ACTIONS; ++j)
        if (j < A) dlogits[b * A + j] = 5.0f * invB * pv[j] * (dpi[j] - dot);
      dv[b] = 2.5f * invB * (vb - yt);
    }
  }
  // block reduction of the loss
  __shared__ float red[kLossWarps];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) lsum += __shfl_xor_sync(0xffffffffu, lsum, o);
  if (lane == 0) red[warp] = lsum;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int i = 0; i < kLossWarps; ++i) s += red[i];
    atomicAdd(loss, 5.0f * invB * s);
  }
}

int launch_returns_loss_grad(const paacb_ctx* ctx, const float* rewards, const float* over, const float* values,
                             const float* boot, const int32_t* actions, const float* pi, const float* v, int T,
                             int64_t N, double gamma, float beta, float* y, float* adv, float* dlogits, float* dv,
                             float* loss, cudaStream_t st) {
  if (N == 0 || T == 0) return PAACB_OK;
  if (T > kLossMaxT) {
    set_error("paacb_returns_loss_grad: t_max %d exceeds the kernel's limit of %d", T, kLossMaxT);
    return PAACB_EUNSUPPORTED;
  }
  if (cudaMemsetAsync(loss, 0, sizeof(float), st) != cudaSuccess) { set_error("memset loss failed"); return PAACB_ECUDA; }
  const unsigned blocks = (unsigned)((N + 31) / 32);
  const size_t smem = (size_t)5 * T * 32 * sizeof(float);
  PAACB_LAUNCH_BEGIN(ctx, K_LOSS, st);
  returns_loss_grad_kernel<<<blocks, kLossThreads, smem, st>>>(rewards, over, values, boot, actions, pi, v, T, N,
                                                                ctx->num_actions, gamma, beta, y, adv, dlogits,
                                                                dv, loss);
  PAACB_LAUNCH_END(ctx, K_LOSS, st);
  return PAACB_OK;
}

}  // namespace paacb
