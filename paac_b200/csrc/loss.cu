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
// One thread per environment walks t = T-1 .. 0 carrying R in a double register (the recurrence is
// float64 in the reference and its outputs are rounded once to float32); everything else is fp32.
#include "common.cuh"

namespace paacb {

constexpr int kLossThreads = 128;

__global__ void __launch_bounds__(kLossThreads)
returns_loss_grad_kernel(const float* __restrict__ rewards, const float* __restrict__ over,
                         const float* __restrict__ values, const float* __restrict__ boot,
                         const int32_t* __restrict__ actions, const float* __restrict__ pi,
                         const float* __restrict__ v, int T, int64_t N, int A, double gamma, float beta,
                         float* __restrict__ y, float* __restrict__ adv, float* __restrict__ dlogits,
                         float* __restrict__ dv, float* __restrict__ loss) {
  const int64_t n = (int64_t)blockIdx.x * kLossThreads + threadIdx.x;
  const float invB = 1.0f / (float)((int64_t)T * N);
  const float eps = 1e-30f;
  float lsum = 0.f;
  if (n < N) {
    double R = (double)boot[n];
    for (int t = T - 1; t >= 0; --t) {
      const int64_t b = (int64_t)t * N + n;
      float r = rewards[b];
      r = r > 1.0f ? 1.0f : (r < -1.0f ? -1.0f : r);
      const double mask = (double)(1.0f - over[b]);
      R = (double)r + gamma * R * mask;
      const float yt = (float)R;
      const float at = (float)(R - (double)values[b]);
      y[b] = yt;
      adv[b] = at;
      // ---- loss and gradient for sample b ----
      const int a = actions[b];
      const float* p = pi + b * A;
      float H = 0.f, dot = 0.f, logsel = 0.f;
      float dpi[PAACB_MAX_ACTIONS];
      float pv[PAACB_MAX_ACTIONS];
#pragma unroll
      for (int j = 0; j < PAACB_MAX_ACTIONS; ++j) {
        if (j < A) {
          const float pj = p[j];
          const float lp = logf(pj + eps);
          H -= pj * lp;
          float d = beta * (lp + pj / (pj + eps));
          if (j == a) { d -= at / (pj + eps); logsel = lp; }
          dpi[j] = d;
          pv[j] = pj;
          dot = fmaf(d, pj, dot);
        }
      }
      const float vb = v[b];
      const float diff = yt - vb;
      lsum += -(logsel * at + beta * H) + 0.25f * diff * diff;
#pragma unroll
      for (int j = 0; j < PAACB_MAX_ACTIONS; ++j)
        if (j < A) dlogits[b * A + j] = 5.0f * invB * pv[j] * (dpi[j] - dot);
      dv[b] = 2.5f * invB * (vb - yt);
    }
  }
  // block reduction of the loss
  __shared__ float red[kLossThreads / 32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) lsum += __shfl_xor_sync(0xffffffffu, lsum, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = lsum;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int i = 0; i < kLossThreads / 32; ++i) s += red[i];
    atomicAdd(loss, 5.0f * invB * s);
  }
}

int launch_returns_loss_grad(const paacb_ctx* ctx, const float* rewards, const float* over, const float* values,
                             const float* boot, const int32_t* actions, const float* pi, const float* v, int T,
                             int64_t N, double gamma, float beta, float* y, float* adv, float* dlogits, float* dv,
                             float* loss, cudaStream_t st) {
  if (N == 0 || T == 0) return PAACB_OK;
  if (cudaMemsetAsync(loss, 0, sizeof(float), st) != cudaSuccess) { set_error("memset loss failed"); return PAACB_ECUDA; }
  const unsigned blocks = (unsigned)((N + kLossThreads - 1) / kLossThreads);
  PAACB_LAUNCH_BEGIN(ctx, K_LOSS, st);
  returns_loss_grad_kernel<<<blocks, kLossThreads, 0, st>>>(rewards, over, values, boot, actions, pi, v, T, N,
                                                             ctx->num_actions, gamma, beta, y, adv, dlogits,
                                                             dv, loss);
  PAACB_LAUNCH_END(ctx, K_LOSS, st);
  return PAACB_OK;
}

}  // namespace paacb
