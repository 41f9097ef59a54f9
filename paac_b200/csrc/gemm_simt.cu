// SIMT fp32 implicit-GEMM kernels for the conv / fc layers: forward, data-gradient, weight-gradient.
// This is the PAACB_MATH_FP32 path: plain FFMA with fp32 accumulation, the parity anchor for the
// tcgen05 path and the arithmetic closest to the reference's fp32 TF graph
// (networks.py:12-21 conv2d, :49-60 fc; gradients via actor_learner.py:44).
//
// All three kernels share one register-tiled core: a CTA of 128 threads multiplies a [32 x BI] tile by
// a [32 x BJ] tile held in shared memory (first index = reduction index), each thread owning TM x TN
// outputs.  They differ in how the tiles are gathered (im2col is never materialised) and in the epilogue.
//   forward : Y[m, n]  = act( sum_k im2col(X)[m, k] W[k, n] + bias[n] )         m = (sample, oh, ow)
//   dgrad   : dX[p, c] = ( sum_{tap, co} dZ[p - tap, co] W[tap, c, co] ) * [X[p, c] > 0]   per stride-parity class
//   wgrad   : dW[k, n] += sum_m im2col(X)[m, k] dZ[m, n];  db[n] += sum_m dZ[m, n]          split over m, fp32 atomics
#include "common.cuh"

namespace paacb {

constexpr int kNT = 128;   // threads per CTA
constexpr int kBR = 32;    // reduction block

template <int TM, int TN, int LDA, int LDB>
__device__ __forceinline__ void tile_fma(const float* __restrict__ As, const float* __restrict__ Bs, int ty, int tx,
                                         float (&acc)[TM][TN]) {
#pragma unroll 8
  for (int r = 0; r < kBR; ++r) {
    float a[TM], b[TN];
    const float* ap = As + r * LDA + ty * TM;
    const float* bp = Bs + r * LDB + tx * TN;
    if constexpr (TM % 4 == 0) {
#pragma unroll
      for (int i = 0; i < TM; i += 4) {
        const float4 t = *reinterpret_cast<const float4*>(ap + i);
        a[i] = t.x; a[i + 1] = t.y; a[i + 2] = t.z; a[i + 3] = t.w;
      }
    } else {
#pragma unroll
      for (int i = 0; i < TM; ++i) a[i] = ap[i];
    }
    if constexpr (TN % 4 == 0) {
#pragma unroll
      for (int j = 0; j < TN; j += 4) {
        const float4 t = *reinterpret_cast<const float4*>(bp + j);
        b[j] = t.x; b[j + 1] = t.y; b[j + 2] = t.z; b[j + 3] = t.w;
      }
    } else if constexpr (TN % 2 == 0) {
#pragma unroll
      for (int j = 0; j < TN; j += 2) {
        const float2 t = *reinterpret_cast<const float2*>(bp + j);
        b[j] = t.x; b[j + 1] = t.y;
      }
    } else {
#pragma unroll
      for (int j = 0; j < TN; ++j) b[j] = bp[j];
    }
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
      for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
  }
}

__device__ __forceinline__ float4 u8x4_to_scaled(uint32_t w) {
  // networks.py:115: tf.scalar_mul(1/255, tf.cast(input, float32)) -- cast then one fp32 multiply
  const float s = 0.003921568859368563f;
  return make_float4((float)(w & 0xffu) * s, (float)((w >> 8) & 0xffu) * s, (float)((w >> 16) & 0xffu) * s,
                     (float)(w >> 24) * s);
}

// ------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------
template <int BN, int TM, int TN, bool U8>
__global__ void __launch_bounds__(kNT)
conv_fwd_simt_kernel(const void* __restrict__ xin, const float* __restrict__ w, const float* __restrict__ bias,
                     float* __restrict__ y, LayerGeom g, int64_t M, int relu) {
  constexpr int BM = kNT;                  // one A row per thread
  constexpr int LDA = BM + 4;
  static_assert((BM / TM) * (BN / TN) == kNT, "thread tiling");
  __shared__ __align__(16) float As[kBR * LDA];
  __shared__ __align__(16) float Bs[kBR * BN];
  const int tid = threadIdx.x;
  const int tx = tid % (BN / TN), ty = tid / (BN / TN);
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;

  // this thread's A row
  const int64_t m = m0 + tid;
  const bool row_ok = m < M;
  int64_t rowbase = 0;
  if (row_ok) {
    const int ohw = g.OH * g.OW;
    const int64_t smp = m / ohw;
    const int rem = (int)(m - smp * ohw);
    const int oh = rem / g.OW, ow = rem - oh * g.OW;
    rowbase = ((smp * g.H + (int64_t)oh * g.stride) * g.W + (int64_t)ow * g.stride) * g.C;
  }
  const int SC = g.S * g.C;
  const int WC = g.W * g.C;

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < g.K; k0 += kBR) {
    const int kh = k0 / SC, off = k0 - kh * SC;
    // ---- A tile: As[k][m] <- 32 consecutive k of row m (contiguous in NHWC) ----
    if constexpr (U8) {
      uint4 v0 = make_uint4(0, 0, 0, 0), v1 = v0;
      if (row_ok) {
        const uint8_t* src = reinterpret_cast<const uint8_t*>(xin) + rowbase + (int64_t)kh * WC + off;
        v0 = __ldg(reinterpret_cast<const uint4*>(src));
        v1 = __ldg(reinterpret_cast<const uint4*>(src) + 1);
      }
      const uint32_t ws[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const float4 f = u8x4_to_scaled(ws[c]);
        As[(c * 4 + 0) * LDA + tid] = f.x;
        As[(c * 4 + 1) * LDA + tid] = f.y;
        As[(c * 4 + 2) * LDA + tid] = f.z;
        As[(c * 4 + 3) * LDA + tid] = f.w;
      }
    } else {
      const float* src = reinterpret_cast<const float*>(xin) + rowbase + (int64_t)kh * WC + off;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        float4 f = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row_ok) f = __ldg(reinterpret_cast<const float4*>(src) + c);
        As[(c * 4 + 0) * LDA + tid] = f.x;
        As[(c * 4 + 1) * LDA + tid] = f.y;
        As[(c * 4 + 2) * LDA + tid] = f.z;
        As[(c * 4 + 3) * LDA + tid] = f.w;
      }
    }
    // ---- B tile: Bs[k][n] <- W[k0 + k][n0 + n] ----
#pragma unroll
    for (int i = 0; i < (kBR * BN / 4) / kNT; ++i) {
      const int idx = tid + i * kNT;
      const int kk = idx / (BN / 4), c4 = idx - kk * (BN / 4);
      float4 f = make_float4(0.f, 0.f, 0.f, 0.f);
      if (n0 + c4 * 4 < g.N) f = __ldg(reinterpret_cast<const float4*>(w + (int64_t)(k0 + kk) * g.N + n0 + c4 * 4));
      *reinterpret_cast<float4*>(Bs + kk * BN + c4 * 4) = f;
    }
    __syncthreads();
    tile_fma<TM, TN, LDA, BN>(As, Bs, ty, tx, acc);
    __syncthreads();
  }

  // ---- epilogue: bias + ReLU, NHWC output is exactly row-major [M, N] ----
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int64_t mm = m0 + ty * TM + i;
    if (mm >= M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int n = n0 + tx * TN + j;
      if (n < g.N) {
        float v = acc[i][j] + __ldg(bias + n);
        if (relu) v = fmaxf(v, 0.f);
        acc[i][j] = v;
      }
    }
    float* dst = y + mm * g.N + n0 + tx * TN;
    if constexpr (TN % 4 == 0) {
#pragma unroll
      for (int j = 0; j < TN; j += 4)
        if (n0 + tx * TN + j < g.N)
          *reinterpret_cast<float4*>(dst + j) = make_float4(acc[i][j], acc[i][j + 1], acc[i][j + 2], acc[i][j + 3]);
    } else {
#pragma unroll
      for (int j = 0; j < TN; ++j)
        if (n0 + tx * TN + j < g.N) dst[j] = acc[i][j];
    }
  }
}

// ------------------------------------------------------------------------------------------------
// dgrad (gather form, one GEMM per stride-parity class; blockIdx.z = class)
// ------------------------------------------------------------------------------------------------
template <int BN, int TM, int TN>
__global__ void __launch_bounds__(kNT)
conv_dgrad_simt_kernel(const float* __restrict__ dz, const float* __restrict__ w, const float* __restrict__ xact,
                       float* __restrict__ dx, LayerGeom g, int64_t batch) {
  constexpr int BM = kNT;
  constexpr int LDA = BM + 4;
  static_assert((BM / TM) * (BN / TN) == kNT, "thread tiling");
  __shared__ __align__(16) float As[kBR * LDA];
  __shared__ __align__(16) float Bs[kBR * BN];
  const int tid = threadIdx.x;
  const int tx = tid % (BN / TN), ty = tid / (BN / TN);
  const int s = g.stride;
  const int ph = blockIdx.z / s, pw = blockIdx.z - ph * s;
  const int Hq = (g.H + s - 1) / s, Wq = (g.W + s - 1) / s;
  const int J = (g.R + s - 1) / s, I = (g.S + s - 1) / s;      // taps per class
  const int64_t Mc = batch * Hq * Wq;
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;                               // input-channel tile

  const int64_t m = m0 + tid;
  const bool row_ok = m < Mc;
  int64_t smp = 0; int hq = 0, wq = 0;
  if (row_ok) {
    smp = m / (Hq * Wq);
    const int rem = (int)(m - smp * (Hq * Wq));
    hq = rem / Wq; wq = rem - hq * Wq;
  }

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  const int Kd = J * I * g.N;                                   // reduction length; g.N = Cout
  for (int k0 = 0; k0 < Kd; k0 += kBR) {
    const int tap = k0 / g.N, co0 = k0 - tap * g.N;
    const int tj = tap / I, ti = tap - tj * I;
    const int kh = ph + s * tj, kw = pw + s * ti;
    const bool tap_ok = (kh < g.R) && (kw < g.S);
    // ---- A tile: dZ[n, hq - tj, wq - ti, co0 .. co0+32) ----
    {
      const int oh = hq - tj, ow = wq - ti;
      const bool ok = row_ok && tap_ok && oh >= 0 && oh < g.OH && ow >= 0 && ow < g.OW;
      const float* src = dz + ((smp * g.OH + oh) * g.OW + ow) * (int64_t)g.N + co0;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        float4 f = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ok) f = __ldg(reinterpret_cast<const float4*>(src) + c);
        As[(c * 4 + 0) * LDA + tid] = f.x;
        As[(c * 4 + 1) * LDA + tid] = f.y;
        As[(c * 4 + 2) * LDA + tid] = f.z;
        As[(c * 4 + 3) * LDA + tid] = f.w;
      }
    }
    // ---- B tile: Bs[co][c] <- W[kh, kw, n0 + c, co0 + co]  (transposing gather, co contiguous in HWIO) ----
#pragma unroll
    for (int i = 0; i < (BN * 8 + kNT - 1) / kNT; ++i) {
      const int idx = tid + i * kNT;
      if (idx < BN * 8) {
        const int c = idx % BN, q = idx / BN;
        float4 f = make_float4(0.f, 0.f, 0.f, 0.f);
        if (tap_ok && n0 + c < g.C)
          f = __ldg(reinterpret_cast<const float4*>(w + ((int64_t)(kh * g.S + kw) * g.C + n0 + c) * g.N + co0 + q * 4));
        Bs[(q * 4 + 0) * BN + c] = f.x;
        Bs[(q * 4 + 1) * BN + c] = f.y;
        Bs[(q * 4 + 2) * BN + c] = f.z;
        Bs[(q * 4 + 3) * BN + c] = f.w;
      }
    }
    __syncthreads();
    tile_fma<TM, TN, LDA, BN>(As, Bs, ty, tx, acc);
    __syncthreads();
  }

  // ---- epilogue: scatter to dX[n, ph + s*hq, pw + s*wq, c], fused ReLU mask of the producing layer ----
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int64_t mm = m0 + ty * TM + i;
    if (mm >= Mc) continue;
    const int64_t sm = mm / (Hq * Wq);
    const int rem = (int)(mm - sm * (Hq * Wq));
    const int qh = rem / Wq, qw = rem - qh * Wq;
    const int h = ph + s * qh, wv = pw + s * qw;
    if (h >= g.H || wv >= g.W) continue;
    const int64_t base = ((sm * g.H + h) * g.W + wv) * (int64_t)g.C + n0 + tx * TN;
#pragma unroll
    for (int j = 0; j < TN; j += 4) {
      if (n0 + tx * TN + j < g.C) {
        float4 o = make_float4(acc[i][j], acc[i][j + 1], acc[i][j + 2], acc[i][j + 3]);
        if (xact != nullptr) {
          const float4 a = __ldg(reinterpret_cast<const float4*>(xact + base + j));
          o.x = a.x > 0.f ? o.x : 0.f; o.y = a.y > 0.f ? o.y : 0.f;
          o.z = a.z > 0.f ? o.z : 0.f; o.w = a.w > 0.f ? o.w : 0.f;
        }
        *reinterpret_cast<float4*>(dx + base + j) = o;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// wgrad (reduction over m split across blockIdx.z; fp32 atomics into zeroed dW / db)
// ------------------------------------------------------------------------------------------------
template <int BJ, int TM, int TN, bool U8>
__global__ void __launch_bounds__(kNT)
conv_wgrad_simt_kernel(const void* __restrict__ xin, const float* __restrict__ dz, float* __restrict__ dw,
                       float* __restrict__ db, LayerGeom g, int64_t M, int64_t rows_per_split) {
  constexpr int BI = 64;                   // k rows of dW per CTA
  static_assert((BI / TM) * (BJ / TN) == kNT, "thread tiling");
  __shared__ __align__(16) float As[kBR * BI];    // As[r][k]
  __shared__ __align__(16) float Bs[kBR * BJ];    // Bs[r][n]
  const int tid = threadIdx.x;
  const int tx = tid % (BJ / TN), ty = tid / (BJ / TN);
  const int kt0 = blockIdx.x * BI;
  const int n0 = blockIdx.y * BJ;
  const int64_t mbeg = (int64_t)blockIdx.z * rows_per_split;
  const int64_t mend = (mbeg + rows_per_split < M) ? mbeg + rows_per_split : M;
  const int SC = g.S * g.C, WC = g.W * g.C, ohw = g.OH * g.OW;

  float acc[TM][TN];
  float bsum[TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
#pragma unroll
  for (int j = 0; j < TN; ++j) bsum[j] = 0.f;
  const bool do_bias = (blockIdx.x == 0) && (ty == 0) && (db != nullptr);

  for (int64_t mb = mbeg; mb < mend; mb += kBR) {
    // ---- A tile: As[r][k] <- im2col(X)[mb + r, kt0 + k]; 16 float4 per row ----
#pragma unroll
    for (int i = 0; i < (kBR * BI / 4) / kNT; ++i) {
      const int idx = tid + i * kNT;
      const int r = idx / (BI / 4), q = idx - r * (BI / 4);
      const int64_t mm = mb + r;
      const int k = kt0 + q * 4;
      float4 f = make_float4(0.f, 0.f, 0.f, 0.f);
      if (mm < mend && k < g.K) {
        const int64_t smp = mm / ohw;
        const int rem = (int)(mm - smp * ohw);
        const int oh = rem / g.OW, ow = rem - oh * g.OW;
        const int kh = k / SC, off = k - kh * SC;
        const int64_t e = ((smp * g.H + (int64_t)oh * g.stride + kh) * g.W + (int64_t)ow * g.stride) * g.C + off;
        if constexpr (U8) {
          f = u8x4_to_scaled(__ldg(reinterpret_cast<const uint32_t*>(reinterpret_cast<const uint8_t*>(xin) + e)));
        } else {
          f = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(xin) + e));
        }
      }
      *reinterpret_cast<float4*>(As + r * BI + q * 4) = f;
    }
    // ---- B tile: Bs[r][n] <- dZ[mb + r, n0 + n] ----
#pragma unroll
    for (int i = 0; i < (kBR * BJ / 4 + kNT - 1) / kNT; ++i) {
      const int idx = tid + i * kNT;
      if (idx < kBR * BJ / 4) {
        const int r = idx / (BJ / 4), q = idx - r * (BJ / 4);
        const int64_t mm = mb + r;
        float4 f = make_float4(0.f, 0.f, 0.f, 0.f);
        if (mm < mend && n0 + q * 4 < g.N) f = __ldg(reinterpret_cast<const float4*>(dz + mm * g.N + n0 + q * 4));
        *reinterpret_cast<float4*>(Bs + r * BJ + q * 4) = f;
      }
    }
    __syncthreads();
    tile_fma<TM, TN, BI, BJ>(As, Bs, ty, tx, acc);
    if (do_bias) {
#pragma unroll 8
      for (int r = 0; r < kBR; ++r)
#pragma unroll
        for (int j = 0; j < TN; ++j) bsum[j] += Bs[r * BJ + tx * TN + j];
    }
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int k = kt0 + ty * TM + i;
    if (k >= g.K) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int n = n0 + tx * TN + j;
      if (n < g.N) atomicAdd(dw + (int64_t)k * g.N + n, acc[i][j]);
    }
  }
  if (do_bias) {
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int n = n0 + tx * TN + j;
      if (n < g.N) atomicAdd(db + n, bsum[j]);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------------
static bool geom_ok(const LayerGeom& g) {
  return (g.S * g.C) % kBR == 0 && g.K % kBR == 0 && g.N % 4 == 0 && g.C % 4 == 0 && g.w_off % 4 == 0;
}

int launch_conv_fwd_simt(const paacb_ctx* ctx, const LayerGeom& g, const void* x, const float* w, const float* bias,
                         float* y, int64_t batch, cudaStream_t st) {
  if (!geom_ok(g)) { set_error("conv_fwd_simt: unsupported geometry"); return PAACB_EUNSUPPORTED; }
  const int64_t M = batch * g.OH * g.OW;
  if (M == 0) return PAACB_OK;
  const unsigned gx = (unsigned)((M + kNT - 1) / kNT);
  PAACB_LAUNCH_BEGIN(ctx, K_FWD0 + g.index, st);
#define FWD(BN, TM, TN, U8) \
  conv_fwd_simt_kernel<BN, TM, TN, U8><<<dim3(gx, (g.N + BN - 1) / BN), kNT, 0, st>>>(x, w, bias, y, g, M, 1)
  if (g.N % 64 == 0) { if (g.in_u8) FWD(64, 8, 8, true); else FWD(64, 8, 8, false); }
  else if (g.N % 32 == 0) { if (g.in_u8) FWD(32, 8, 4, true); else FWD(32, 8, 4, false); }
  else { if (g.in_u8) FWD(16, 4, 4, true); else FWD(16, 4, 4, false); }
#undef FWD
  PAACB_LAUNCH_END(ctx, K_FWD0 + g.index, st);
  return PAACB_OK;
}

int launch_conv_dgrad_simt(const paacb_ctx* ctx, const LayerGeom& g, const float* dz, const float* w, const float* x_act,
                           float* dx, int64_t batch, cudaStream_t st) {
  if (!geom_ok(g) || g.N % kBR != 0 || g.in_u8) { set_error("conv_dgrad_simt: unsupported geometry"); return PAACB_EUNSUPPORTED; }
  if (batch == 0) return PAACB_OK;
  const int s = g.stride;
  const int Hq = (g.H + s - 1) / s, Wq = (g.W + s - 1) / s;
  const int64_t Mc = batch * Hq * Wq;
  const unsigned gx = (unsigned)((Mc + kNT - 1) / kNT);
  PAACB_LAUNCH_BEGIN(ctx, K_DGRAD0 + g.index, st);
#define DG(BN, TM, TN) \
  conv_dgrad_simt_kernel<BN, TM, TN><<<dim3(gx, (g.C + BN - 1) / BN, s * s), kNT, 0, st>>>(dz, w, x_act, dx, g, batch)
  if (g.C >= 64) DG(64, 8, 8);
  else if (g.C >= 32) DG(32, 8, 4);
  else DG(16, 4, 4);
#undef DG
  PAACB_LAUNCH_END(ctx, K_DGRAD0 + g.index, st);
  return PAACB_OK;
}

int launch_conv_wgrad_simt(const paacb_ctx* ctx, const LayerGeom& g, const void* x, const float* dz, float* dw, float* db,
                           int64_t batch, cudaStream_t st) {
  if (!geom_ok(g)) { set_error("conv_wgrad_simt: unsupported geometry"); return PAACB_EUNSUPPORTED; }
  const int64_t M = batch * g.OH * g.OW;
  if (M == 0) return PAACB_OK;
  const int bj = (g.N % 64 == 0) ? 64 : (g.N % 32 == 0 ? 32 : 16);
  const int tiles = ((g.K + 63) / 64) * ((g.N + bj - 1) / bj);
  const int64_t row_blocks = (M + kBR - 1) / kBR;
  int64_t splits = (4LL * ctx->num_sms + tiles - 1) / tiles;
  if (splits > row_blocks) splits = row_blocks;
  if (splits < 1) splits = 1;
  if (splits > 65535) splits = 65535;
  const int64_t rows_per_split = ((row_blocks + splits - 1) / splits) * kBR;
  splits = (M + rows_per_split - 1) / rows_per_split;
  const dim3 grid((g.K + 63) / 64, (g.N + bj - 1) / bj, (unsigned)splits);
  PAACB_LAUNCH_BEGIN(ctx, K_WGRAD0 + g.index, st);
#define WG(BJ, TM, TN, U8) \
  conv_wgrad_simt_kernel<BJ, TM, TN, U8><<<grid, kNT, 0, st>>>(x, dz, dw, db, g, M, rows_per_split)
  if (bj == 64) { if (g.in_u8) WG(64, 4, 8, true); else WG(64, 4, 8, false); }
  else if (bj == 32) { if (g.in_u8) WG(32, 4, 4, true); else WG(32, 4, 4, false); }
  else { if (g.in_u8) WG(16, 4, 2, true); else WG(16, 4, 2, false); }
#undef WG
  PAACB_LAUNCH_END(ctx, K_WGRAD0 + g.index, st);
  return PAACB_OK;
}

}  // namespace paacb
