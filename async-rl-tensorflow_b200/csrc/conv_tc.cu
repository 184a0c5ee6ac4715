// K2 on the tensor cores: the two convolutions as implicit GEMMs on tcgen05 (gemm_tc.cuh).
// The im2col matrix is never materialised in HBM: producer warps gather each 8-element K chunk
// (8 consecutive kw bytes of a frame row for conv1, 8 consecutive channels of a pixel for
// conv2) straight from the ring / activation tensor into the UMMA smem layout; the (tiny)
// weight matrix is converted once per CTA and stays resident in smem.
//
//   conv1 fwd : rows = (sample, oy, ox) [400/sample], N = 16, K = 256 ordered (c, kh, kw);
//               u8 pixels are exact in bf16 -> A has no lo part (2 MMAs per k-step);
//               epilogue a1 = relu(acc/255 + b1)
//   conv2 fwd : rows = (sample, oy, ox) [81/sample],  N = 32, K = 256 ordered (kh, kw, c);
//               fp32 activations split hi/lo (3 MMAs); epilogue a2 = relu(acc + b2), which is
//               already the NHWC-flattened fc input
#include "gemm_tc.cuh"

namespace arl {

// ------------------------------------ conv1 forward ------------------------------------------
struct Conv1FwdArgs {
  const float* params;
  const uint8_t* ring;
  float* a1;
  int num_envs, ring_slots, first_slot;
  int64_t rows;                       // 400 * num_samples
};

struct Conv1FwdPolicy {
  using Args = Conv1FwdArgs;
  static constexpr int N_TILE = 16, KB = 32;
  static constexpr bool A_HAS_LO = false, B_RESIDENT = true;
  static constexpr int B_RES_K = 256, B_RES_SETS = 1;
  using TA = tc::OperandTile<tc::kTileM, KB>;
  using TB = tc::OperandTile<N_TILE, B_RES_K>;

  static __device__ __forceinline__ int num_items(const Args& g) {
    return (int)((g.rows + tc::kTileM - 1) / tc::kTileM);
  }
  static __device__ __forceinline__ tc::TileCoord coord(const Args&, int item) {
    tc::TileCoord t;
    t.mt = item; t.nt = 0; t.ks = 0; t.k_begin = 0; t.k_end = 256;
    return t;
  }
  static __device__ __forceinline__ void load_A(const Args& g, const tc::TileCoord& t, int k0,
                                                uint8_t* hi, uint8_t*, int lane) {
    const int kb = k0 >> 5;
    const int c = kb >> 1, kh0 = (kb & 1) * 4;            // k = (c*8 + kh)*8 + kw
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int row = lane + 32 * j;
      const int64_t m = (int64_t)t.mt * tc::kTileM + row;
      uint32_t w[4][2];
      if (m < g.rows) {
        const int n = (int)(m / 400), p = (int)(m - (int64_t)n * 400);
        const int oy = p / 20, ox = p - oy * 20;
        const int tt = n / g.num_envs, b = n - tt * g.num_envs;
        const int slot = (g.first_slot + tt + c) % g.ring_slots;
        const uint8_t* src = g.ring + ((size_t)b * g.ring_slots + slot) * kPlane +
                             (4 * oy + kh0) * ARL_SCREEN + 4 * ox;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const uint32_t* s32 = reinterpret_cast<const uint32_t*>(src + q * ARL_SCREEN);
          w[q][0] = __ldg(s32);
          w[q][1] = __ldg(s32 + 1);
        }
      } else {
#pragma unroll
        for (int q = 0; q < 4; ++q) w[q][0] = w[q][1] = 0u;
      }
#pragma unroll
      for (int q = 0; q < 4; ++q)
        *reinterpret_cast<uint4*>(hi + q * TA::LBO + row * 16) = tc::bytes8_to_bf16(w[q][0], w[q][1]);
    }
  }
  static __device__ __forceinline__ void load_B(const Args&, const tc::TileCoord&, int, uint8_t*,
                                                uint8_t*, int) {}
  static __device__ __forceinline__ void load_B_resident(const Args& g, uint8_t* base, int ptid) {
    // element (co, k=(c,kh,kw)) = W1[((kh*8+kw)*4 + c)*16 + co]
    for (int ch = ptid; ch < 16 * 32; ch += tc::kProdThreads) {
      const int co = ch & 15, kc = ch >> 4;                // kc = c*8 + kh
      const int c = kc >> 3, kh = kc & 7;
      float x[8];
#pragma unroll
      for (int kw = 0; kw < 8; ++kw) x[kw] = g.params[((kh * 8 + kw) * 4 + c) * 16 + co];
      tc::store_chunk_split(base, base + TB::BYTES, kc * TB::LBO + co * 16, x);
    }
  }
  static __device__ __forceinline__ int b_set(const Args&, const tc::TileCoord&) { return 0; }
  static __device__ __forceinline__ int res_k_origin(const Args&, const tc::TileCoord&) { return 0; }
  static __device__ __forceinline__ void store(const Args& g, const tc::TileCoord& t, int row, int,
                                               const float (&v)[16]) {
    const int64_t m = (int64_t)t.mt * tc::kTileM + row;
    if (m >= g.rows) return;
    const float* bias = g.params + 4096;
    float* d = g.a1 + m * 16;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float4 bb = *reinterpret_cast<const float4*>(bias + 4 * q);
      float4 o;
      o.x = fmaxf(fmaf(v[4 * q], 1.0f / 255.0f, bb.x), 0.f);
      o.y = fmaxf(fmaf(v[4 * q + 1], 1.0f / 255.0f, bb.y), 0.f);
      o.z = fmaxf(fmaf(v[4 * q + 2], 1.0f / 255.0f, bb.z), 0.f);
      o.w = fmaxf(fmaf(v[4 * q + 3], 1.0f / 255.0f, bb.w), 0.f);
      *reinterpret_cast<float4*>(d + 4 * q) = o;
    }
  }
};

// ------------------------------------ conv2 forward ------------------------------------------
struct Conv2FwdArgs {
  const float* params;
  const float* a1;
  float* a2;
  int64_t rows;                       // 81 * num_samples
};

struct Conv2FwdPolicy {
  using Args = Conv2FwdArgs;
  static constexpr int N_TILE = 32, KB = 32;
  static constexpr bool A_HAS_LO = true, B_RESIDENT = true;
  static constexpr int B_RES_K = 256, B_RES_SETS = 1;
  using TA = tc::OperandTile<tc::kTileM, KB>;
  using TB = tc::OperandTile<N_TILE, B_RES_K>;

  static __device__ __forceinline__ int num_items(const Args& g) {
    return (int)((g.rows + tc::kTileM - 1) / tc::kTileM);
  }
  static __device__ __forceinline__ tc::TileCoord coord(const Args&, int item) {
    tc::TileCoord t;
    t.mt = item; t.nt = 0; t.ks = 0; t.k_begin = 0; t.k_end = 256;
    return t;
  }
  static __device__ __forceinline__ void load_A(const Args& g, const tc::TileCoord& t, int k0,
                                                uint8_t* hi, uint8_t* lo, int lane) {
    const int q0 = k0 >> 4;                                // (kh*4 + kw) of the first chunk pair
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int row = lane + 32 * j;
      const int64_t m = (int64_t)t.mt * tc::kTileM + row;
      const bool ok = m < g.rows;
      const int n = ok ? (int)(m / 81) : 0, p = ok ? (int)(m - (int64_t)n * 81) : 0;
      const int oy = p / 9, ox = p - oy * 9;
      const float* base = g.a1 + (int64_t)n * ARL_A1_ELEMS;
#pragma unroll
      for (int kc = 0; kc < 4; ++kc) {
        const int q = q0 + (kc >> 1), kh = q >> 2, kw = q & 3;
        float x[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (ok) {
          const float4* s = reinterpret_cast<const float4*>(
              base + ((2 * oy + kh) * 20 + 2 * ox + kw) * 16 + (kc & 1) * 8);
          const float4 x0 = __ldg(s), x1 = __ldg(s + 1);
          x[0] = x0.x; x[1] = x0.y; x[2] = x0.z; x[3] = x0.w;
          x[4] = x1.x; x[5] = x1.y; x[6] = x1.z; x[7] = x1.w;
        }
        tc::store_chunk_split(hi, lo, kc * TA::LBO + row * 16, x);
      }
    }
  }
  static __device__ __forceinline__ void load_B(const Args&, const tc::TileCoord&, int, uint8_t*,
                                                uint8_t*, int) {}
  static __device__ __forceinline__ void load_B_resident(const Args& g, uint8_t* base, int ptid) {
    // element (co, k=(kh,kw,c)) = W2[k*32 + co]: sample-major source, ld = 32
    if (ptid < 32)
      tc::load_tile_f32<N_TILE, B_RES_K, true>(base, base + TB::BYTES, g.params + 4112, 32, 0, 32,
                                               0, 256, ptid);
  }
  static __device__ __forceinline__ int b_set(const Args&, const tc::TileCoord&) { return 0; }
  static __device__ __forceinline__ int res_k_origin(const Args&, const tc::TileCoord&) { return 0; }
  static __device__ __forceinline__ void store(const Args& g, const tc::TileCoord& t, int row,
                                               int c, const float (&v)[16]) {
    const int64_t m = (int64_t)t.mt * tc::kTileM + row;
    if (m >= g.rows) return;
    const float* bias = g.params + 4112 + 8192 + c;
    float* d = g.a2 + m * 32 + c;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float4 bb = *reinterpret_cast<const float4*>(bias + 4 * q);
      float4 o;
      o.x = fmaxf(v[4 * q] + bb.x, 0.f);
      o.y = fmaxf(v[4 * q + 1] + bb.y, 0.f);
      o.z = fmaxf(v[4 * q + 2] + bb.z, 0.f);
      o.w = fmaxf(v[4 * q + 3] + bb.w, 0.f);
      *reinterpret_cast<float4*>(d + 4 * q) = o;
    }
  }
};

}  // namespace arl

using namespace arl;

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

extern "C" int arl_conv1_forward(const float* params, const uint8_t* ring, float* a1, int num_envs,
                                 int ring_slots, int first_slot, int steps, void* stream) {
  ARL_REQUIRE(params && ring && a1, "arl_conv1_forward: null pointer");
  ARL_REQUIRE(num_envs >= 0 && steps >= 0, "arl_conv1_forward: negative size");
  ARL_REQUIRE(ring_slots >= steps + 3 && first_slot >= 0 && first_slot < ring_slots,
              "arl_conv1_forward: ring_slots %d must be >= steps+3 (%d) and first_slot %d inside it",
              ring_slots, steps + 3, first_slot);
  ARL_REQUIRE(aligned16(params) && aligned16(ring) && aligned16(a1),
              "arl_conv1_forward: pointers must be 16-byte aligned");
  const int64_t N = (int64_t)num_envs * steps;
  if (N == 0) return ARL_OK;
  Conv1FwdArgs g{params, ring, a1, num_envs, ring_slots, first_slot, N * 400};
  const int64_t items = (g.rows + tc::kTileM - 1) / tc::kTileM;
  ARL_REQUIRE(items < (1LL << 31), "arl_conv1_forward: too many samples");
  return tc::launch<Conv1FwdPolicy>(g, (int)items, (cudaStream_t)stream);
}

extern "C" int arl_conv2_forward(const float* params, const float* a1, float* a2,
                                 int64_t num_samples, void* stream) {
  ARL_REQUIRE(params && a1 && a2, "arl_conv2_forward: null pointer");
  ARL_REQUIRE(num_samples >= 0, "arl_conv2_forward: negative size");
  ARL_REQUIRE(aligned16(params) && aligned16(a1) && aligned16(a2),
              "arl_conv2_forward: pointers must be 16-byte aligned");
  if (num_samples == 0) return ARL_OK;
  Conv2FwdArgs g{params, a1, a2, num_samples * 81};
  const int64_t items = (g.rows + tc::kTileM - 1) / tc::kTileM;
  ARL_REQUIRE(items < (1LL << 31), "arl_conv2_forward: too many samples");
  return tc::launch<Conv2FwdPolicy>(g, (int)items, (cudaStream_t)stream);
}
