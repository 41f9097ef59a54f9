// bf16-split tensor-core path, part 1: patch-resident implicit GEMMs for the conv layers (forward and data-gradient),
// plus the small producers of their operands (uint8 -> bf16 states, packed weight images) and the tensor-map helper.
//
// Design (sm_100a):
//  * Activations live in HBM as TWO bf16 planes, value = hi + lo (|x - hi - lo| <= 2^-18 |x|).  A product is evaluated as
//    hi*hi + hi*lo + lo*hi in three kind::f16 MMAs with fp32 accumulation in TMEM: the dropped lo*lo term and the
//    residuals are ~2^-17 relative, an order of magnitude inside the 1e-4 parity bar, at twice the MMA rate and half
//    the operand bytes of the tf32x3 scheme.  uint8 pixels are exact in bf16 (conv1 needs two MMAs per K-step).
//  * No im2col, neither in HBM nor in shared memory.  The activation patch a tile of 128 output positions needs is
//    landed ONCE by the TMA engine, row-parity plane by row-parity plane, as rows of s*C consecutive channels
//    (32 B for conv1, 128 B for conv2/conv3).  Output positions are enumerated over the (padded) INPUT grid, so the
//    operand of filter tap (kh, kw) is the same patch at a byte offset: every tap is just another start address in
//    the MMA's shared-memory descriptor (the absolute-address swizzle that makes this legal was measured, tc_ptx.cuh).
//    Shared-memory fill traffic is 1x the input instead of the 4-9x an im2col-style tile pipeline re-reads.
//  * The layer's weights (both bf16 pieces, <= 144 KB) stay resident in shared memory for the whole kernel.
//  * Persistent CTA per SM: warp 0 = TMA producer, warp 1 = MMA issuer (one elected lane), warps 2-5 = epilogue
//    (tcgen05.ld, bias + ReLU or ReLU-mask, bf16 split, 16-byte stores; bias gradients as warp-transposed column sums).
//    Accumulators are double-buffered in TMEM; patch slots form an mbarrier ring.
#include "tc2.cuh"

namespace paacb {

// ------------------------------------------------------------------------------------------------
// tensor maps
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
    else
      cudaGetLastError();
  }
  return fn;
}

int encode_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                     const uint32_t* box, int swizzle_bytes) {
  return encode_tmap(out, base, 2, rank, dims, strides_bytes, box, swizzle_bytes);
}

int encode_tmap(CUtensorMap* out, const void* base, int elem_bytes, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                const uint32_t* box, int swizzle_bytes) {
  EncodeTiledFn fn = encode_fn();
  if (fn == nullptr) {
    set_error("cuTensorMapEncodeTiled is not available from this driver");
    return PAACB_ECUDA;
  }
  cuuint64_t gd[5];
  cuuint64_t gs[5];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) {
    gd[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
    if (i > 0) gs[i - 1] = strides_bytes[i - 1];
  }
  const CUtensorMapSwizzle sw = swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                : swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE;
  const CUtensorMapDataType dt = elem_bytes == 1 ? CU_TENSOR_MAP_DATA_TYPE_UINT8
                                 : elem_bytes == 4 ? CU_TENSOR_MAP_DATA_TYPE_UINT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  const CUresult r = fn(out, dt, (cuuint32_t)rank, const_cast<void*>(base), gd, gs, bx, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d): rank %d dims %llu,%llu,%llu,%llu box %u,%u,%u,%u", (int)r, rank,
              (unsigned long long)gd[0], (unsigned long long)(rank > 1 ? gd[1] : 0), (unsigned long long)(rank > 2 ? gd[2] : 0),
              (unsigned long long)(rank > 3 ? gd[3] : 0), bx[0], rank > 1 ? bx[1] : 0, rank > 2 ? bx[2] : 0, rank > 3 ? bx[3] : 0);
    return PAACB_ECUDA;
  }
  return PAACB_OK;
}

bool bf16x3_supported(const paacb_ctx* ctx) { return ctx->arch == PAACB_ARCH_NATURE || ctx->arch == PAACB_ARCH_NIPS; }

// ------------------------------------------------------------------------------------------------
// operand producers
// ------------------------------------------------------------------------------------------------
// forward image: Wp[n][k] = W[k][n] as (hi, lo) bf16, row-major [N][K] (the B operand, K-major).  32 x 32 tiles through
// shared memory so that both the fp32 reads and the bf16 writes are coalesced (the fc layer has 1.6 M weights and the
// image is rebuilt on every forward call).
__global__ void __launch_bounds__(256) pack_bf16_transpose_kernel(const float* __restrict__ w, int K, int N,
                                                                  uint16_t* __restrict__ hi, uint16_t* __restrict__ lo) {
  __shared__ float tile[32][33];
  const int k0 = blockIdx.y * 32, n0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;       // 8 rows per pass
#pragma unroll
  for (int r = ty; r < 32; r += 8) {
    const int k = k0 + r, n = n0 + tx;
    tile[r][tx] = (k < K && n < N) ? __ldg(w + (int64_t)k * N + n) : 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int r = ty; r < 32; r += 8) {
    const int n = n0 + r, k = k0 + tx;
    if (n < N && k < K) {
      uint16_t h, l;
      split_bf16(tile[tx][r], h, l);
      hi[(int64_t)n * K + k] = h;
      lo[(int64_t)n * K + k] = l;
    }
  }
}
// same orientation as stored (fc data-gradient: rows = inputs, K = outputs)
__global__ void pack_bf16_copy_kernel(const float* __restrict__ w, int64_t total, uint16_t* __restrict__ hi,
                                      uint16_t* __restrict__ lo) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    uint16_t h, l;
    split_bf16(__ldg(w + i), h, l);
    hi[i] = h;
    lo[i] = l;
  }
}
// conv data-gradient image: Wd[cls][ci][(tj*I + ti)*Cout + co] = W[ph + s*tj][pw + s*ti][ci][co], cls = ph*s + pw
__global__ void pack_bf16_dgrad_kernel(const float* __restrict__ w, LayerGeom g, int J, int I, uint16_t* __restrict__ hi,
                                       uint16_t* __restrict__ lo) {
  const int s = g.stride;
  const int Kd = J * I * g.N;
  const int64_t total = (int64_t)s * s * g.C * Kd;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int kd = (int)(i % Kd);
    const int64_t rest = i / Kd;
    const int ci = (int)(rest % g.C), cls = (int)(rest / g.C);
    const int tap = kd / g.N, co = kd - tap * g.N;
    const int tj = tap / I, ti = tap - tj * I;
    const int kh = cls / s + s * tj, kw = cls % s + s * ti;
    uint16_t h, l;
    split_bf16(__ldg(w + ((int64_t)(kh * g.S + kw) * g.C + ci) * g.N + co), h, l);
    hi[i] = h;
    lo[i] = l;
  }
}

int launch_pack_bf16_weights(const paacb_ctx* ctx, const float* params, cudaStream_t st) {
  const int rc = launch_pack_conv1_i8(ctx, params, st);      // conv1 runs on the int8 pipe (tc2_conv1.cu)
  if (rc != PAACB_OK) return rc;
  for (int l = 1; l < ctx->n_layers; ++l) {
    const LayerGeom& g = ctx->layer[l];
    const dim3 grid((unsigned)((g.N + 31) / 32), (unsigned)((g.K + 31) / 32));
    PAACB_LAUNCH_BEGIN(ctx, K_PACK, st);
    pack_bf16_transpose_kernel<<<grid, 256, 0, st>>>(params + g.w_off, g.K, g.N, ctx->wb_f_hi + g.w_off, ctx->wb_f_lo + g.w_off);
    PAACB_LAUNCH_END(ctx, K_PACK, st);
  }
  return PAACB_OK;
}

int launch_pack_bf16_dgrad_weights(const paacb_ctx* ctx, const float* params, cudaStream_t st) {
  for (int l = 1; l < ctx->n_layers; ++l) {
    const LayerGeom& g = ctx->layer[l];
    const int64_t total = (int64_t)g.K * g.N;
    PAACB_LAUNCH_BEGIN(ctx, K_PACK, st);
    if (g.R == 1 && g.S == 1) {
      pack_bf16_copy_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(params + g.w_off, total, ctx->wb_d_hi + g.w_off,
                                                                             ctx->wb_d_lo + g.w_off);
    } else {
      const int s = g.stride;
      pack_bf16_dgrad_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(params + g.w_off, g, g.R / s, g.S / s,
                                                                              ctx->wb_d_hi + g.w_off, ctx->wb_d_lo + g.w_off);
    }
    PAACB_LAUNCH_END(ctx, K_PACK, st);
  }
  return PAACB_OK;
}

// ------------------------------------------------------------------------------------------------
// geometry of the five patch-resident GEMMs of the Nature network
// ------------------------------------------------------------------------------------------------
enum { G_FWD2 = 1, G_FWD3 = 2, G_DG3 = 3, G_DG2 = 4,        // Nature; conv1 forward has its own int8 kernel (tc2_conv1.cu)
       G_FWD2N = 5, G_DG2N = 6 };                           // NIPS conv2 (networks.py:146): 64-byte units, SWIZZLE_64B

template <int G>
struct Geo;

// conv2 forward: [b,20,20,32] 4x4 stride 2 -> [b,9,9,64].  Unit = 2 pixels x 32 channels = 128 B; 2 row-parity planes.
template <>
struct Geo<G_FWD2> {
  static constexpr bool DGRAD = false, A_LO = true, STAGED = false;
  static constexpr int UB = 128, SWZ = SWZ_128B, PARTS = 2, WU = 10, HQ = 10, BOX_ROWS = 15, SLOT = 19456, NSLOTS = 5;
  static constexpr int NACC = 1, BN = 64, KS = 16, KB = 8, OH = 9, OW = 9;
  static constexpr int CH = 64, PIX = 1;
  __host__ __device__ static constexpr int aoff(int t) { return ((t / 8) * 10 + ((t / 4) % 2)) * 128 + (t % 4) * 32; }
  __host__ __device__ static constexpr int jw(int part, int t) { return ((part + 2 * (t / 8)) * 2 + ((t / 4) % 2)) * 4 + (t % 4); }
};
// conv3 forward: [b,9,9,64] 3x3 stride 1 -> [b,7,7,64].  Unit = 1 pixel x 64 channels = 128 B.
template <>
struct Geo<G_FWD3> {
  static constexpr bool DGRAD = false, A_LO = true, STAGED = false;
  static constexpr int UB = 128, SWZ = SWZ_128B, PARTS = 1, WU = 9, HQ = 9, BOX_ROWS = 18, SLOT = 21504, NSLOTS = 3;
  static constexpr int NACC = 1, BN = 64, KS = 36, KB = 9, OH = 7, OW = 7;
  static constexpr int CH = 64, PIX = 1;
  __host__ __device__ static constexpr int aoff(int t) { return (((t / 4) / 3) * 9 + ((t / 4) % 3)) * 128 + (t % 4) * 32; }
  __host__ __device__ static constexpr int jw(int, int t) { return t; }
};
// conv3 data-gradient: dZ3 [b,7,7,64] -> dX [b,9,9,64]; one sample per tile, positions enumerated 11 wide over the
// zero-padded dZ (TMA out-of-bounds fill), tap (kh, kw) reads position q + (2-kh)*11 + (2-kw).
template <>
struct Geo<G_DG3> {
  static constexpr bool DGRAD = true, A_LO = true, STAGED = false;
  static constexpr int UB = 128, SWZ = SWZ_128B, PARTS = 1, WU = 11, HQ = 9, BOX_ROWS = 11, SLOT = 16384, NSLOTS = 4;
  static constexpr int NACC = 1, BN = 64, KS = 36, KB = 9, OH = 9, OW = 9;       // OH/OW: valid rows/cols of the enumeration
  static constexpr int PAD = 2, S = 1, XH = 9, XW = 9;
  static constexpr int CH = 64, PIX = 1;      // output channels per pixel, pixels per accumulator row
  __host__ __device__ static constexpr int aoff(int t) { return ((2 - (t / 4) / 3) * 11 + (2 - (t / 4) % 3)) * 128 + (t % 4) * 32; }
  __host__ __device__ static constexpr int jw(int, int t) { return t; }
};
// conv2 data-gradient: dZ2 [b,9,9,64] -> dX [b,20,20,32]; the four stride-parity classes are four accumulators over
// the same resident dZ patch (class (ph, pw), tap (tj, ti) reads position q + (1-tj)*11 + (1-ti)).
// STAGED epilogue: a thread's outputs of the two classes (ph, 0), (ph, 1) are one 128-byte row (2 pixels x 32 channels)
// of a [10 x 10 rows] x 128 B tile per output-row parity ph; written straight to HBM every warp store touched 32
// different cache lines (the first version was bound by LSU wavefronts: 34 % tensor-pipe activity).  The rows go to a
// swizzled shared-memory tile instead and the TMA engine stores the tile.
template <>
struct Geo<G_DG2> {
  static constexpr bool DGRAD = true, A_LO = true, STAGED = true;
  static constexpr int UB = 128, SWZ = SWZ_128B, PARTS = 1, WU = 11, HQ = 10, BOX_ROWS = 11, SLOT = 16384, NSLOTS = 4;
  static constexpr int NACC = 4, BN = 32, KS = 16, KB = 4, OH = 10, OW = 10;
  static constexpr int PAD = 1, S = 2, XH = 20, XW = 20;
  static constexpr int CH = 32, PIX = 1;
  __host__ __device__ static constexpr int aoff(int t) { return ((1 - (t / 4) / 2) * 11 + (1 - (t / 4) % 2)) * 128 + (t % 4) * 32; }
  __host__ __device__ static constexpr int jw(int, int t) { return t; }
};

// NIPS conv2 forward: [b,20,20,16] 4x4 stride 2 -> [b,9,9,32].  Unit = 2 pixels x 16 channels = 64 B (SWIZZLE_64B rows); one
// K = 16 MMA is exactly one pixel's 16 channels, so filter tap (kh, kw) = plane kh & 1, row offset kh >> 1, unit offset
// kw >> 1, 32-byte half kw & 1.
template <>
struct Geo<G_FWD2N> {
  static constexpr bool DGRAD = false, A_LO = true, STAGED = false;
  static constexpr int UB = 64, SWZ = SWZ_64B, PARTS = 2, WU = 10, HQ = 10, BOX_ROWS = 15, SLOT = 10240, NSLOTS = 6;
  static constexpr int NACC = 1, BN = 32, KS = 8, KB = 4, OH = 9, OW = 9;
  static constexpr int CH = 32, PIX = 1;
  __host__ __device__ static constexpr int aoff(int t) { return ((t / 4) * 10 + ((t / 2) % 2)) * 64 + (t % 2) * 32; }
  __host__ __device__ static constexpr int jw(int part, int t) { return (part + 2 * (t / 4)) * 4 + 2 * ((t / 2) % 2) + (t % 2); }
};
// NIPS conv2 data-gradient: dZ2 [b,9,9,32] -> dX [b,20,20,16].  The four stride-parity classes (ph, pw) x 16 input channels are
// TWO accumulators of 32 columns: accumulator ph holds (pw, ci) = the two horizontally adjacent output pixels 2 qw, 2 qw + 1,
// which are 32 contiguous elements of the NHWC output -- the packed data-gradient image is row-ordered (ph, pw, ci) already,
// and a thread's 32 columns are one 64-byte run of each output plane (direct 256-bit stores, no staging tile).
template <>
struct Geo<G_DG2N> {
  static constexpr bool DGRAD = true, A_LO = true, STAGED = false;
  static constexpr int UB = 64, SWZ = SWZ_64B, PARTS = 1, WU = 11, HQ = 10, BOX_ROWS = 11, SLOT = 8192, NSLOTS = 6;
  static constexpr int NACC = 2, BN = 32, KS = 8, KB = 2, OH = 10, OW = 10;
  static constexpr int PAD = 1, S = 2, XH = 20, XW = 20;
  static constexpr int CH = 16, PIX = 2;      // output channels per pixel, pixels per accumulator row
  __host__ __device__ static constexpr int aoff(int t) { return ((1 - (t / 2) / 2) * 11 + (1 - (t / 2) % 2)) * 64 + (t % 2) * 32; }
  __host__ __device__ static constexpr int jw(int, int t) { return t; }
};

struct ConvKParams {
  CUtensorMap tmA[2];      // source planes (hi, lo); forward: dims (unit, units/row, row parity, plane rows); dgrad: (C, OW, OH, b)
  CUtensorMap tmW[2];      // packed weights (hi, lo): dims (K, rows)
  CUtensorMap tmOut[2];    // staged epilogue: output planes (hi, lo) as (2 px x C, W/2, row parity, b * H/2)
  int num_tiles;
  int batch;
  const float* bias;       // forward
  float in_scale;          // forward: applied to the accumulator (1/255 for the uint8 layer, networks.py:115)
  uint8_t* out_hi;         // output planes (bf16)
  uint8_t* out_lo;
  const uint8_t* mask_hi;  // dgrad: hi plane of the activation whose ReLU is differentiated (same shape as the output)
  int dbg;                 // PAACB_DBG ablations (timing experiments only): 256 no A_lo MMAs, 512 no stores, 1024 no patch loads, 4096 conv2 dgrad: lo plane through the staging tile, 8192 conv2 dgrad: both planes by direct stores (measured: no gain once the staging tile is gone)
  float* dbias;            // dgrad: += column sums of the output (the bias gradient of the layer that produced that activation)
};

template <int G>
struct ConvKCfg {
  using Ge = Geo<G>;
  static constexpr int RING_BYTES = Ge::NSLOTS * Ge::SLOT;
  // resident weights: per 64-wide K-block one K-major tile of 2 * NT rows: the hi pieces of all accumulators, then the
  // lo pieces.  One MMA of N = 2 * NT evaluates A_hi * [W_hi | W_lo] (the operand A is read from shared memory once:
  // with N <= 64 an SS-mode MMA is bound by its 4 KB A read, not by the tensor pipe), a second of N = NT adds A_lo * W_hi.
  static constexpr int NT = Ge::NACC * Ge::BN;
  static constexpr int KB_BYTES = 2 * NT * 128;
  static constexpr int W_BYTES = Ge::KB * KB_BYTES;
  static constexpr int BOX_BYTES = Ge::UB * Ge::WU * Ge::BOX_ROWS;
  static constexpr int STG_TILE = 13 * 1024;                                // 100 rows x 128 B, 1024-byte aligned
  static constexpr int STG_BYTES = Ge::STAGED ? 2 * STG_TILE : 0;            // one tile per output-row parity: hi plane, then lo plane
  static constexpr int NBARS = 2 * Ge::NSLOTS + 1 + 4 + 1;
  static constexpr int SMEM_BYTES = RING_BYTES + W_BYTES + STG_BYTES + 1024 /* alignment slack */ + NBARS * 8 + 16 + 256 /* forward: bias row */;
  static constexpr int TMEM_COLS = 4 * NT;                                  // two accumulator buffers of 2 * NT columns
  static_assert(BOX_BYTES <= Ge::SLOT && Ge::SLOT % 1024 == 0, "slot too small");
  // TMA warp, MMA warp, two epilogue groups of four warps (one per accumulator buffer; the staged epilogue splits every
  // tile between the groups: output-row parity 0 / 1)
  static constexpr int THREADS = 64 + 256;
  static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
  static_assert(TMEM_COLS == 128 || TMEM_COLS == 256 || TMEM_COLS == 512, "TMEM columns");
  static_assert(2 * NT <= 256, "MMA N");
};


// lane j of the warp ends up with the sum over the warp's 32 lanes of v[j] (31 shuffles)
__device__ __forceinline__ float warp_transpose_sum(float (&v)[32], int lane) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const bool up = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < off; ++i) {
      const float send = up ? v[i] : v[i + off];
      const float keep = up ? v[i + off] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  return v[0];
}

template <int G>
__global__ void __launch_bounds__(ConvKCfg<G>::THREADS, 1) convk_kernel(const __grid_constant__ ConvKParams p) {
  using Ge = Geo<G>;
  using Cfg = ConvKCfg<G>;
  constexpr int NSLOTS = Ge::NSLOTS, BN = Ge::BN, NACC = Ge::NACC;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* ring = smem;
  uint8_t* wsm = smem + Cfg::RING_BYTES;
  uint8_t* stg = wsm + Cfg::W_BYTES;              // staged epilogue tiles
  uint64_t* bars = reinterpret_cast<uint64_t*>(stg + Cfg::STG_BYTES);
  uint64_t* full_bar = bars;                      // [NSLOTS] TMA -> MMA
  uint64_t* empty_bar = bars + NSLOTS;            // [NSLOTS] MMA commit -> TMA
  uint64_t* w_bar = bars + 2 * NSLOTS;            // weights resident
  uint64_t* tfull_bar = w_bar + 1;                // [2] MMA commit -> epilogue
  uint64_t* tempty_bar = tfull_bar + 2;           // [2] epilogue -> MMA
  uint64_t* mask_bar = tempty_bar + 2;            // staged epilogue: the staging tiles have been read out by the TMA stores
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mask_bar + 1);
  float* s_bias = reinterpret_cast<float*>(bars + Cfg::NBARS + 2);      // forward: the layer's bias row (16-byte aligned)

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    for (int s = 0; s < NSLOTS; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(w_bar, 1);
    mbar_init(mask_bar, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], Ge::STAGED ? 256 : 128);     // the epilogue warps that drain buffer s
    }
    fence_barrier_init();
    tma_prefetch_desc(&p.tmA[0]);
    tma_prefetch_desc(&p.tmW[0]);
  }
  if (warp == 1) tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
  if constexpr (!Ge::DGRAD) {
    if (tid >= 64 && tid < 64 + BN) s_bias[tid - 64] = __ldg(p.bias + tid - 64);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // =========================== TMA producer ===========================
    if (elect_one_sync()) {
      mbar_arrive_expect_tx(w_bar, Cfg::W_BYTES);
      for (int piece = 0; piece < 2; ++piece)
        for (int acc = 0; acc < NACC; ++acc)
          for (int kb = 0; kb < Ge::KB; ++kb)
            tma_load_2d(wsm + kb * Cfg::KB_BYTES + (piece * NACC + acc) * (BN * 128), &p.tmW[piece], kb * 64, acc * BN, w_bar);
      int slot = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        // (requesting the boxes of the tile after next into L2 with cp.async.bulk.prefetch.tensor was measured: 3 % slower)
#pragma unroll 1
        for (int part = 0; part < Ge::PARTS; ++part) {
#pragma unroll 1
          for (int piece = 0; piece < (Ge::A_LO ? 2 : 1); ++piece) {
            mbar_wait(&empty_bar[slot], phase ^ 1u);
            if (PAACB_DBGV(p.dbg) & 1024) { mbar_arrive(&full_bar[slot]); if (++slot == NSLOTS) { slot = 0; phase ^= 1u; } continue; }
            mbar_arrive_expect_tx(&full_bar[slot], Cfg::BOX_BYTES);
            uint8_t* dst = ring + slot * Ge::SLOT;
            if constexpr (Ge::DGRAD) {
              tma_load_4d(dst, &p.tmA[piece], 0, -Ge::PAD, -Ge::PAD, tile, &full_bar[slot]);
            } else {
              const int row0 = (int)(((int64_t)tile * 128) / Ge::WU);
              tma_load_4d(dst, &p.tmA[piece], 0, 0, part, row0, &full_bar[slot]);
            }
            if (++slot == NSLOTS) { slot = 0; phase ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer ===========================
    const bool leader = elect_one_sync();
    constexpr uint32_t idesc_full = make_idesc_bf16(2 * Cfg::NT, 0, 0);
    constexpr uint32_t idesc_half = make_idesc_bf16(Cfg::NT, 0, 0);
    const uint64_t adesc0 = make_smem_desc(0, 16, 8 * Ge::UB, Ge::SWZ);
    const uint64_t bdesc0 = make_smem_desc(0, 16, 1024, SWZ_128B);
    const uint32_t ring_a = smem_u32(ring);
    const uint32_t w_a = smem_u32(wsm);
    mbar_wait(w_bar, 0);
    int slot = 0;
    uint32_t phase = 0;
    int tl = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++tl) {
      const int ab = tl & 1;
      const uint32_t aph = (uint32_t)((tl >> 1) & 1);
      mbar_wait(&tempty_bar[ab], aph ^ 1u);
      tc_fence_after();
      uint32_t rel = 0;
      if constexpr (!Ge::DGRAD) rel = (uint32_t)(((int64_t)tile * 128) % Ge::WU) * Ge::UB;
      const uint32_t d0 = tmem_base + (uint32_t)(ab * 2 * Cfg::NT);
#pragma unroll
      for (int part = 0; part < Ge::PARTS; ++part) {
#pragma unroll
        for (int piece = 0; piece < (Ge::A_LO ? 2 : 1); ++piece) {
          mbar_wait(&full_bar[slot], phase);
          tc_fence_after();
          if (leader) {
            const uint32_t a_base = ring_a + (uint32_t)(slot * Ge::SLOT) + rel;
#pragma unroll
            for (int t = 0; t < Ge::KS; ++t) {
              const uint64_t ad = desc_with_addr(adesc0, a_base + (uint32_t)Ge::aoff(t));
              const int jw = Ge::jw(part, t);
              const uint64_t bd = desc_with_addr(bdesc0, w_a + (uint32_t)((jw / 4) * Cfg::KB_BYTES + (jw % 4) * 32));
              if (piece == 0) umma_bf16(d0, ad, bd, idesc_full, (part | t) ? 1u : 0u);    // A_hi * [W_hi | W_lo]
              else if (!(PAACB_DBGV(p.dbg) & 256)) umma_bf16(d0, ad, bd, idesc_half, 1u);              // A_lo * W_hi
            }
            umma_commit(&empty_bar[slot]);
          }
          __syncwarp();
          if (++slot == NSLOTS) { slot = 0; phase ^= 1u; }
        }
      }
      if (leader) umma_commit(&tfull_bar[ab]);
      __syncwarp();
    }
  } else if (Ge::STAGED) {
    // =========================== staged epilogue (conv2 data-gradient) ===========================
    if constexpr (Ge::STAGED) {
      const int ew = warp & 3;
      const int r = ew * 32 + lane;
      const int qh = r / Ge::WU, qw = r - qh * Ge::WU;
      const bool ok = (qh < Ge::OH) && (qw < Ge::OW);
      const uint32_t srow = (uint32_t)(qh * Ge::OW + qw);                  // row of the 100 x 128 B staging tile
      const uint32_t stg_a = smem_u32(stg);
      const int ph = (warp - 2) >> 2;                                       // this warp's output-row parity: classes (ph, 0), (ph, 1)
      const bool io = (tid == 64 + 128 * ph);                               // issues the TMA stores of this group's tile
      float bsum = 0.f;
      // The ReLU-mask words of this thread's two output pixels (2 x 64 B = one 128-byte line) are loaded at the top of the
      // tile, before the wait for the MMAs; the line of the NEXT tile is requested into L2 at the same time, so the load
      // is an L2 hit.  (Version 1 loaded them cold: a DRAM round trip per tile.  Version 2 carried them one tile ahead in
      // registers: 32 registers this kernel does not have -- the spill put a local-memory load on the critical path right
      // after the MMA barrier, 18 % of all stall samples.)
      uint4 mk[2][4];
      auto mask_ptr = [&](int tile, int pw) {
        const int64_t ob = (((int64_t)tile * Ge::XH + Ge::S * qh + ph) * Ge::XW + Ge::S * qw + pw) * BN;
        return reinterpret_cast<const uint4*>(p.mask_hi + ob * 2);
      };
      if (ok && (int)blockIdx.x < p.num_tiles) asm volatile("prefetch.global.L2 [%0];" ::"l"(mask_ptr((int)blockIdx.x, 0)));
      int tl = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++tl) {
        const int ab = tl & 1;
        const uint32_t aph = (uint32_t)((tl >> 1) & 1);
#pragma unroll
        for (int pw = 0; pw < 2; ++pw)
#pragma unroll
          for (int j = 0; j < 4; ++j) mk[pw][j] = make_uint4(0u, 0u, 0u, 0u);
        if (ok) {
#pragma unroll
          for (int pw = 0; pw < 2; ++pw) {
            ldg256(mask_ptr(tile, pw), mk[pw][0], mk[pw][1]);
            ldg256(mask_ptr(tile, pw) + 2, mk[pw][2], mk[pw][3]);
          }
        }
        if (ok && tile + (int)gridDim.x < p.num_tiles) asm volatile("prefetch.global.L2 [%0];" ::"l"(mask_ptr(tile + (int)gridDim.x, 0)));
        mbar_wait(&tfull_bar[ab], aph);
        tc_fence_after();
        float o[2][32];
#pragma unroll
        for (int pw = 0; pw < 2; ++pw) {
          uint32_t v[32], v2[32];
          const uint32_t tcol = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(ab * 2 * Cfg::NT + (ph * 2 + pw) * BN);
          tmem_ld32(tcol, v);
          tmem_ld32(tcol + (uint32_t)Cfg::NT, v2);
          tmem_ld_wait();
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const uint32_t mw[4] = {mk[pw][c].x, mk[pw][c].y, mk[pw][c].z, mk[pw][c].w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const int e = c * 8 + j * 2;
              // bf16 > 0  <=>  the 16 bits, as a signed integer in the top half of a word, are > 0
              const bool p0 = ok && ((int32_t)(mw[j] << 16) > 0), p1 = ok && ((int32_t)(mw[j] & 0xffff0000u) > 0);
              o[pw][e] = p0 ? __uint_as_float(v[e]) + __uint_as_float(v2[e]) : 0.f;
              o[pw][e + 1] = p1 ? __uint_as_float(v[e + 1]) + __uint_as_float(v2[e + 1]) : 0.f;
            }
          }
        }
        tc_fence_before();
        mbar_arrive(&tempty_bar[ab]);                                       // accumulator drained: the MMAs of tile tl + 2 may start
        // This group's 100 x 128 B staging tile carries the hi plane, then the lo plane (shared memory is needed for a
        // 4-deep patch ring: with 2 slots the kernel was bound by the TMA round trip).  io = the group's first thread.
        uint8_t* tile_p = stg + ph * Cfg::STG_TILE;
        const uint32_t t_row = stg_a + (uint32_t)(ph * Cfg::STG_TILE) + srow * 128u;
        uint32_t lw[2][16];
        const bool direct_hi = (PAACB_DBGV(p.dbg) & 8192) != 0;                         // experiment: no staging at all
        if (!direct_hi) {
          if (io) tma_store_wait_read();                                    // previous tile's hi plane has been read out
          named_bar_sync(1 + ph, 128);
        }
#pragma unroll
        for (int pw = 0; pw < 2; ++pw) {
          uint32_t hw[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) split_bf16x2(o[pw][2 * j], o[pw][2 * j + 1], hw[j], lw[pw][j]);
          if (ok) {
            if (direct_hi) {
              uint8_t* dh = p.out_hi + ((((int64_t)tile * Ge::XH + Ge::S * qh + ph) * Ge::XW + Ge::S * qw) * BN) * 2;
#pragma unroll
              for (int c = 0; c < 2; ++c) stg256(dh + pw * 64 + c * 32, &hw[8 * c]);
            } else {
#pragma unroll
              for (int c = 0; c < 4; ++c)                                   // 16-byte chunk pw * 4 + c of the row, swizzled
                sts128(t_row + ((((uint32_t)(pw * 4 + c)) ^ (srow & 7u)) << 4), make_uint4(hw[4 * c], hw[4 * c + 1], hw[4 * c + 2], hw[4 * c + 3]));
            }
          }
        }
        if (!direct_hi) {
          fence_proxy_async();                                              // staging writes -> visible to the TMA engine
          named_bar_sync(1 + ph, 128);
          if (io) {
            tma_store_4d(&p.tmOut[0], tile_p, 0, 0, ph, tile * Ge::OH);
            tma_store_commit();
          }
        }
        // bias gradient: channel = lane for both classes: add the classes first, ONE transposed warp reduction per tile
        // (this kernel has no registers left for per-thread column sums; the shuffles overlap the TMA engine reading the tile)
#pragma unroll
        for (int j = 0; j < 32; ++j) o[0][j] += o[1][j];
        bsum += warp_transpose_sum(o[0], lane);
        if (PAACB_DBGV(p.dbg) & 4096) {
          // (first staged version, kept for A/B timing: the lo plane follows the hi plane through the same staging tile --
          // two more barriers and a second wait for the TMA engine per tile)
          if (io) tma_store_wait_read();
          named_bar_sync(1 + ph, 128);
          if (ok) {
#pragma unroll
            for (int pw = 0; pw < 2; ++pw)
#pragma unroll
              for (int c = 0; c < 4; ++c)
                sts128(t_row + ((((uint32_t)(pw * 4 + c)) ^ (srow & 7u)) << 4), make_uint4(lw[pw][4 * c], lw[pw][4 * c + 1], lw[pw][4 * c + 2], lw[pw][4 * c + 3]));
          }
          fence_proxy_async();
          named_bar_sync(1 + ph, 128);
          if (io) {
            tma_store_4d(&p.tmOut[1], tile_p, 0, 0, ph, tile * Ge::OH);
            tma_store_commit();
          }
        } else if (ok) {
          // lo plane: this thread's two pixels are one 128-byte row of the plane -- four 256-bit stores straight from
          // registers while the TMA engine drains the hi tile (there is no shared memory for a second staging tile)
          uint8_t* dl = p.out_lo + ((((int64_t)tile * Ge::XH + Ge::S * qh + ph) * Ge::XW + Ge::S * qw) * BN) * 2;
#pragma unroll
          for (int pw = 0; pw < 2; ++pw)
#pragma unroll
            for (int c = 0; c < 2; ++c) stg256(dl + pw * 64 + c * 32, &lw[pw][8 * c]);
        }
      }
      if (io) tma_store_wait_all();
      if (p.dbias != nullptr) atomicAdd(p.dbias + lane, bsum);
    }
  } else {
    // =========================== epilogue ===========================
    const int ew = warp & 3;                     // TMEM lane quarter this warp may access
    const int r = ew * 32 + lane;                // tile row of this thread
    const int grp = (warp - 2) >> 2;             // epilogue group = accumulator buffer it drains (tiles tl with tl & 1 == grp)
    // (per-thread column sums reduced once at the end -- what tc2_stream.cu's dgrad does -- cost this kernel 42 registers
    // it does not have: conv3's data gradient got 7 % slower with them)
    float bsum[NACC == 1 ? BN / 32 : 1];
#pragma unroll
    for (int i = 0; i < (NACC == 1 ? BN / 32 : 1); ++i) bsum[i] = 0.f;
    for (int tl = grp, tile = blockIdx.x + grp * (int)gridDim.x; tile < p.num_tiles; tile += 2 * (int)gridDim.x, tl += 2) {
      const int ab = grp;
      const uint32_t aph = (uint32_t)((tl >> 1) & 1);
      // row -> output pixel
      bool ok;
      int64_t pix = 0;          // forward: output pixel index; dgrad: (sample, qh, qw) resolved per class below
      int qh = 0, qw = 0;
      if constexpr (Ge::DGRAD) {
        qh = r / Ge::WU;
        qw = r - qh * Ge::WU;
        ok = (qh < Ge::OH) && (qw < Ge::OW);
      } else {
        const uint32_t q = (uint32_t)tile * 128u + (uint32_t)r;      // < 2^31 (checked by the launcher)
        const uint32_t prow = q / (uint32_t)Ge::WU;
        const int ju = (int)(q - prow * (uint32_t)Ge::WU);
        const uint32_t n = prow / (uint32_t)Ge::HQ;
        const int oh = (int)(prow - n * (uint32_t)Ge::HQ);
        ok = (ju < Ge::OW) && (oh < Ge::OH) && ((int)n < p.batch);
        pix = ((int64_t)n * Ge::OH + oh) * Ge::OW + ju;
      }
      // dgrad: the ReLU-mask words of the whole tile row do not depend on the accumulator: fetch them before waiting
      // for the MMAs so that their DRAM latency overlaps the mainloop (they were the top stall of the first version)
      uint4 mk[Ge::DGRAD ? NACC : 1][Ge::DGRAD ? BN / 32 : 1][4];
      if constexpr (Ge::DGRAD) {
#pragma unroll
        for (int acc = 0; acc < NACC; ++acc) {
          // accumulator -> output pixel: PIX == 1: class (acc / 2, acc % 2) of a stride-2 layer; PIX == 2: row parity acc, the
          // 32 columns are the pixels 2 qw, 2 qw + 1 (CH channels each)
          const int ih = Ge::S * qh + (NACC > 1 ? (Ge::PIX == 2 ? acc : acc / 2) : 0);
          const int iw = Ge::S * qw + ((NACC > 1 && Ge::PIX == 1) ? acc % 2 : 0);
          const int64_t ob = (((int64_t)tile * Ge::XH + ih) * Ge::XW + iw) * Ge::CH;
#pragma unroll
          for (int c = 0; c < BN / 32; ++c) {
#pragma unroll
            for (int j = 0; j < 4; ++j) mk[acc][c][j] = make_uint4(0u, 0u, 0u, 0u);
            if (ok) {
              ldg256(p.mask_hi + (ob + c * 32) * 2, mk[acc][c][0], mk[acc][c][1]);
              ldg256(p.mask_hi + (ob + c * 32) * 2 + 32, mk[acc][c][2], mk[acc][c][3]);
            }
          }
        }
      }
      mbar_wait(&tfull_bar[ab], aph);
      tc_fence_after();
      if constexpr (!Ge::DGRAD) {
        // forward: the whole accumulator row goes to registers first and the buffer is handed back to the MMA warp BEFORE
        // the arithmetic and the stores (the data-gradient variants have no registers for that: they release it at the end)
        uint32_t av[BN / 32][32];
#pragma unroll
        for (int c = 0; c < BN / 32; ++c) {
          uint32_t v2[32];
          const uint32_t tcol = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(ab * 2 * Cfg::NT + c * 32);
          tmem_ld32(tcol, av[c]);                               // A_hi * W_hi + A_lo * W_hi
          tmem_ld32(tcol + (uint32_t)Cfg::NT, v2);              // A_hi * W_lo
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) av[c][j] = __float_as_uint(__uint_as_float(av[c][j]) + __uint_as_float(v2[j]));
        }
        tc_fence_before();
        mbar_arrive(&tempty_bar[ab]);
        const float4* b4 = reinterpret_cast<const float4*>(s_bias);
#pragma unroll
        for (int c = 0; c < BN / 32; ++c) {
          uint32_t hw[16], lw[16];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 bb = b4[c * 8 + j];                    // broadcast 16-byte shared-memory load: 4 channels
            const float o0 = fmaxf(fmaf(__uint_as_float(av[c][4 * j]), p.in_scale, bb.x), 0.f);
            const float o1 = fmaxf(fmaf(__uint_as_float(av[c][4 * j + 1]), p.in_scale, bb.y), 0.f);
            const float o2 = fmaxf(fmaf(__uint_as_float(av[c][4 * j + 2]), p.in_scale, bb.z), 0.f);
            const float o3 = fmaxf(fmaf(__uint_as_float(av[c][4 * j + 3]), p.in_scale, bb.w), 0.f);
            split_bf16x2(o0, o1, hw[2 * j], lw[2 * j]);
            split_bf16x2(o2, o3, hw[2 * j + 1], lw[2 * j + 1]);
          }
          if (ok && !(PAACB_DBGV(p.dbg) & 512)) {
            uint8_t* dh = p.out_hi + (pix * BN + c * 32) * 2;
            uint8_t* dl = p.out_lo + (pix * BN + c * 32) * 2;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              stg256(dh + 32 * j, hw + 8 * j);
              stg256(dl + 32 * j, lw + 8 * j);
            }
          }
        }
        continue;
      }
#pragma unroll
      for (int acc = 0; acc < NACC; ++acc) {
        int64_t obase;          // element index of this row's first output channel
        if constexpr (Ge::DGRAD) {
          const int ih = Ge::S * qh + (NACC > 1 ? (Ge::PIX == 2 ? acc : acc / 2) : 0);
          const int iw = Ge::S * qw + ((NACC > 1 && Ge::PIX == 1) ? acc % 2 : 0);
          obase = (((int64_t)tile * Ge::XH + ih) * Ge::XW + iw) * Ge::CH;
        } else {
          obase = pix * BN;
        }
#pragma unroll
        for (int c0 = 0; c0 < BN; c0 += 32) {
          uint32_t v[32], v2[32];
          const uint32_t tcol = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(ab * 2 * Cfg::NT + acc * BN + c0);
          tmem_ld32(tcol, v);                                   // A_hi * W_hi + A_lo * W_hi
          tmem_ld32(tcol + (uint32_t)Cfg::NT, v2);              // A_hi * W_lo
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(v2[j]));
          float o[32];
          if constexpr (Ge::DGRAD) {
            uint32_t mw[16];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint4 m = mk[acc][c0 / 32][j];
              mw[4 * j] = m.x; mw[4 * j + 1] = m.y; mw[4 * j + 2] = m.z; mw[4 * j + 3] = m.w;
            }
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              // bf16 > 0  <=>  sign clear and magnitude non-zero
              const uint32_t a0 = mw[j] & 0xffffu, a1 = mw[j] >> 16;
              o[2 * j] = (ok && a0 != 0u && a0 < 0x8000u) ? __uint_as_float(v[2 * j]) : 0.f;
              o[2 * j + 1] = (ok && a1 != 0u && a1 < 0x8000u) ? __uint_as_float(v[2 * j + 1]) : 0.f;
            }
          }
          if (ok && !(PAACB_DBGV(p.dbg) & 512)) {
            uint32_t hw[16], lw[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) split_bf16x2(o[2 * j], o[2 * j + 1], hw[j], lw[j]);
            uint8_t* dh = p.out_hi + (obase + c0) * 2;
            uint8_t* dl = p.out_lo + (obase + c0) * 2;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              stg256(dh + 32 * j, hw + 8 * j);
              stg256(dl + 32 * j, lw + 8 * j);
            }
          }
          if constexpr (Ge::DGRAD) {
            const float s = warp_transpose_sum(o, lane);      // column c0 + lane over this warp's 32 rows
            bsum[NACC == 1 ? c0 / 32 : 0] += s;
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&tempty_bar[ab]);
    }
    if constexpr (Ge::DGRAD) {
      if (p.dbias != nullptr) {
        // column -> channel: PIX == 2 packs two pixels of CH channels into one accumulator row
#pragma unroll
        for (int i = 0; i < (NACC == 1 ? BN / 32 : 1); ++i) atomicAdd(p.dbias + (i * 32 + lane) % Ge::CH, bsum[i]);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------------
template <int G>
static int launch_convk(const paacb_ctx* ctx, const ConvKParams& p, int slot, cudaStream_t st) {
  using Cfg = ConvKCfg<G>;
  static DeviceOnce attr_set;   // kernel attributes are per device: one bit per device index
  if (!attr_set.done(ctx->device)) {
    if (cudaFuncSetAttribute(convk_kernel<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES) != cudaSuccess) {
      cudaGetLastError();
      set_error("convk<%d>: cannot set %d bytes of dynamic shared memory", G, Cfg::SMEM_BYTES);
      return PAACB_ECUDA;
    }
    // one persistent CTA per SM: ask for the largest shared-memory carve-out whatever the kernel's own request (measured
    // on conv1 forward: the same code ran 10 % slower when its request dropped from 160 KB to 86 KB)
    cudaFuncSetAttribute(convk_kernel<G>, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
    cudaGetLastError();
    attr_set.mark(ctx->device);
  }
  const unsigned grid = (unsigned)(p.num_tiles < ctx->num_sms ? p.num_tiles : ctx->num_sms);
  PAACB_LAUNCH_BEGIN(ctx, slot, st);
  convk_kernel<G><<<grid, Cfg::THREADS, Cfg::SMEM_BYTES, st>>>(p);
  PAACB_LAUNCH_END(ctx, slot, st);
  return PAACB_OK;
}

static int weight_maps(const paacb_ctx* ctx, const uint16_t* hi, const uint16_t* lo, const LayerGeom& g, uint64_t kdim,
                       uint64_t rows, uint32_t box_rows, CUtensorMap* out) {
  const uint64_t dims[2] = {kdim, rows};
  const uint64_t strides[1] = {kdim * 2};
  const uint32_t box[2] = {64, box_rows};
  int rc = encode_tmap_bf16(&out[0], hi + g.w_off, 2, dims, strides, box, 128);
  if (rc != PAACB_OK) return rc;
  return encode_tmap_bf16(&out[1], lo + g.w_off, 2, dims, strides, box, 128);
}

int launch_conv_fwd_bf16(const paacb_ctx* ctx, int l, const float* params, const uint8_t* states, void* fwd_ws, int64_t batch,
                         const WsSlice& slice, cudaStream_t st) {
  const LayerGeom& g = ctx->layer[l];
  ConvKParams p;
  memset(&p, 0, sizeof(p));
  p.batch = (int)batch;
  p.dbg = ctx->dbg;
  p.bias = params + g.b_off;
  p.in_scale = g.in_u8 ? 0.003921568859368563f : 1.0f;
  const Planes out = layer_planes(fwd_ws, g.out_act_off, (int64_t)g.OH * g.OW * g.N, slice);
  p.out_hi = out.hi;
  p.out_lo = out.lo;
  const uint8_t* in_hi;
  const uint8_t* in_lo;
  if (l == 0) {
    in_hi = in_lo = nullptr;                    // conv1 converts the uint8 states itself
  } else {
    const Planes in = layer_planes(fwd_ws, g.in_act_off, (int64_t)g.H * g.W * g.C, slice);
    in_hi = in.hi;
    in_lo = in.lo;
  }
  const int s = g.stride;
  const uint64_t unit = (uint64_t)s * g.C;                      // elements per unit
  const uint64_t wu = (uint64_t)g.W / s, hq = (uint64_t)g.H / s;
  const uint64_t dims[4] = {unit, wu, (uint64_t)s, (uint64_t)batch * hq};
  const uint64_t strides[3] = {unit * 2, (uint64_t)g.W * g.C * 2, (uint64_t)s * g.W * g.C * 2};
  int rc;
#define PAACB_FWD_CASE(GE)                                                                                         \
  {                                                                                                                \
    using Ge = Geo<GE>;                                                                                            \
    if ((int)unit * 2 != Ge::UB || (int)wu != Ge::WU || (int)hq != Ge::HQ || g.N != Ge::BN || g.K != Ge::KB * 64 ||  \
        s != Ge::PARTS || g.OH != Ge::OH || g.OW != Ge::OW)                                                        \
      break;                                                                                                       \
    const uint32_t box[4] = {(uint32_t)unit, (uint32_t)Ge::WU, 1u, (uint32_t)Ge::BOX_ROWS};                        \
    rc = encode_tmap_bf16(&p.tmA[0], in_hi, 4, dims, strides, box, Ge::UB);                                        \
    if (rc == PAACB_OK) rc = encode_tmap_bf16(&p.tmA[1], in_lo, 4, dims, strides, box, Ge::UB);                    \
    if (rc == PAACB_OK) rc = weight_maps(ctx, ctx->wb_f_hi, ctx->wb_f_lo, g, (uint64_t)g.K, (uint64_t)g.N, Ge::BN, p.tmW); \
    if (rc != PAACB_OK) return rc;                                                                                 \
    const int64_t q_total = batch * Ge::HQ * Ge::WU;                                                               \
    if (q_total >= (1LL << 31) - 256) return PAACB_EUNSUPPORTED;                                                   \
    p.num_tiles = (int)((q_total + 127) / 128);                                                                    \
    return launch_convk<GE>(ctx, p, K_FWD0 + l, st);                                                               \
  }
  if (l == 0) return launch_conv1_fwd_i8(ctx, params, states, fwd_ws, batch, slice, st);
  // (each case leaves its do-while with `break` when the layer's geometry is not the one it was written for)
  if (l == 1) do PAACB_FWD_CASE(G_FWD2) while (0);
  if (l == 1) do PAACB_FWD_CASE(G_FWD2N) while (0);
  if (l == 2) do PAACB_FWD_CASE(G_FWD3) while (0);
#undef PAACB_FWD_CASE
  return PAACB_EUNSUPPORTED;
}

// dX of layer l's input: A = dZ planes of layer l (bwd_ws), output = dZ planes of layer l-1, masked by layer l-1's activation
int launch_conv_dgrad_bf16(const paacb_ctx* ctx, int l, const void* fwd_ws, void* bwd_ws, float* grads, int64_t batch,
                           cudaStream_t st) {
  const LayerGeom& g = ctx->layer[l];
  const LayerGeom& gp = ctx->layer[l - 1];
  ConvKParams p;
  memset(&p, 0, sizeof(p));
  p.batch = (int)batch;
  p.dbg = ctx->dbg;
  p.num_tiles = (int)batch;
  const Planes dz = layer_planes(bwd_ws, g.out_act_off, (int64_t)g.OH * g.OW * g.N, batch);
  const Planes dx = layer_planes(bwd_ws, gp.out_act_off, (int64_t)g.H * g.W * g.C, batch);
  const Planes xa = layer_planes(const_cast<void*>(fwd_ws), gp.out_act_off, (int64_t)g.H * g.W * g.C, batch);
  p.out_hi = dx.hi;
  p.out_lo = dx.lo;
  p.mask_hi = xa.hi;
  p.dbias = grads + gp.b_off;
  const uint64_t dims[4] = {(uint64_t)g.N, (uint64_t)g.OW, (uint64_t)g.OH, (uint64_t)batch};
  const uint64_t strides[3] = {(uint64_t)g.N * 2, (uint64_t)g.OW * g.N * 2, (uint64_t)g.OH * g.OW * g.N * 2};
  int rc;
#define PAACB_DG_CASE(GE)                                                                                          \
  {                                                                                                                \
    using Ge = Geo<GE>;                                                                                            \
    const int s = g.stride;                                                                                        \
    if (g.N * 2 != Ge::UB || g.C != Ge::CH || s != Ge::S || g.H != Ge::XH || g.W != Ge::XW ||                       \
        (g.R / s) * (g.S / s) * g.N != Ge::KB * 64 || s * s * g.C != Ge::NACC * Ge::BN ||                           \
        (g.R - 1) / s != Ge::PAD || g.OH + 2 * Ge::PAD > Ge::BOX_ROWS + 0 || (g.H + s - 1) / s != Ge::OH)          \
      break;                                                                                                       \
    const uint32_t box[4] = {(uint32_t)g.N, (uint32_t)Ge::WU, (uint32_t)Ge::BOX_ROWS, 1u};                          \
    rc = encode_tmap_bf16(&p.tmA[0], dz.hi, 4, dims, strides, box, Ge::UB);                                        \
    if (rc == PAACB_OK) rc = encode_tmap_bf16(&p.tmA[1], dz.lo, 4, dims, strides, box, Ge::UB);                    \
    if (rc == PAACB_OK)                                                                                            \
      rc = weight_maps(ctx, ctx->wb_d_hi, ctx->wb_d_lo, g, (uint64_t)(g.R / s) * (g.S / s) * g.N,                  \
                       (uint64_t)s * s * g.C, Ge::BN, p.tmW);                                                      \
    if (rc == PAACB_OK && Ge::STAGED) {                                                                            \
      const uint64_t od[4] = {(uint64_t)2 * g.C, (uint64_t)g.W / 2, 2u, (uint64_t)batch * (g.H / 2)};              \
      const uint64_t os[3] = {(uint64_t)2 * g.C * 2, (uint64_t)g.W * g.C * 2, (uint64_t)2 * g.W * g.C * 2};        \
      const uint32_t ob[4] = {(uint32_t)(2 * g.C), (uint32_t)(g.W / 2), 1u, (uint32_t)(g.H / 2)};                  \
      rc = encode_tmap_bf16(&p.tmOut[0], dx.hi, 4, od, os, ob, 128);                                               \
      if (rc == PAACB_OK) rc = encode_tmap_bf16(&p.tmOut[1], dx.lo, 4, od, os, ob, 128);                           \
    }                                                                                                              \
    if (rc != PAACB_OK) return rc;                                                                                 \
    return launch_convk<GE>(ctx, p, K_DGRAD0 + l, st);                                                             \
  }
  if (g.stride == 1) do PAACB_DG_CASE(G_DG3) while (0);
  if (g.stride == 2) do PAACB_DG_CASE(G_DG2) while (0);
  if (g.stride == 2) do PAACB_DG_CASE(G_DG2N) while (0);
#undef PAACB_DG_CASE
  return PAACB_EUNSUPPORTED;
}

}  // namespace paacb
