// K2 on the tensor cores: both convolutions, forward and backward, as implicit GEMMs on tcgen05
// (gemm_tc.cuh).  No im2col matrix ever exists in HBM, and none is built in smem either:
//
//   space-to-depth + shifted descriptors.  An 8x8 stride-4 conv over [84,84,4] is a 2x2
//   stride-1 conv over the space-to-depth tensor X[21,21,64]; a 4x4 stride-2 conv over
//   [20,20,16] is a 2x2 stride-1 conv over X2[10,10,64].  If output pixels are numbered on the
//   SAME grid as X (one garbage column/row per sample), output row P needs X rows P + a*GW + b
//   for the four taps (a,b): a constant row shift.  In the UMMA K-major no-swizzle image a row
//   shift is a 16-byte shift of the descriptor start address, so the four taps are four
//   descriptors onto ONE smem image of the tile's X rows (+ a halo).  Every input element is
//   loaded from HBM and converted exactly once per tile.
//
//   conv1 fwd  : X rows from the u8 ring (exact in bf16 -> hi image only), W1 resident
//   conv2 fwd  : X2 rows from a1 (bf16 hi+lo), W2 resident
//   conv2 dgrad: the transposed conv.  Output pixels of one parity class (y&1, x&1) use 4 taps
//                of the zero-padded dy2 grid Z[11,11,32]; the 4 classes share the Z image and
//                differ only in the resident weights -> 4 accumulator sets per stage
//   wgrads     : dW[(tap,ch), co] = sum_P X[P+shift][ch] * dyz[P][co] reduces over ROWS, so the
//                same row-major images serve as MN-major operands (k = row index); tap b is a
//                second copy of the image shifted by one row, tap a a descriptor row offset.
//                Work is split over row ranges; partials are summed in a fixed order.
#include "gemm_tc.cuh"

namespace arl {

int reduce_partials(const float* partials, float* out, int num_partials, int n, cudaStream_t stream);

namespace {
using tc::TileCoord;
using tc::kTileM;

__device__ __forceinline__ void load8(const float* p, float (&x)[8]) {
  const float4 x0 = __ldg(reinterpret_cast<const float4*>(p));
  const float4 x1 = __ldg(reinterpret_cast<const float4*>(p) + 1);
  x[0] = x0.x; x[1] = x0.y; x[2] = x0.z; x[3] = x0.w;
  x[4] = x1.x; x[5] = x1.y; x[6] = x1.z; x[7] = x1.w;
}
__device__ __forceinline__ void zero8(float (&x)[8]) {
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = 0.f;
}
__device__ __forceinline__ TileCoord row_tile(int item) {
  TileCoord t;
  t.mt = item; t.nt = 0; t.ks = 0; t.k_begin = 0; t.k_end = 0;
  return t;
}

// ---- gather of one space-to-depth row ---------------------------------------------------------
// conv1: X row (n, y', x') -> 8 chunks; chunk kc = c*2 + h holds channels ch = (c*4+i)*4+j for
// i in {2h, 2h+1}, j in 0..3 = frame rows 4y'+2h, 4y'+2h+1, columns 4x'..4x'+3 of plane c.
struct RingGeo {
  const uint8_t* ring;
  int num_envs, ring_slots, first_slot;
};
__device__ __forceinline__ void conv1_x_row(const RingGeo& g, int n, int yp, int xp, uint4 (&out)[8]) {
  const int tt = n / g.num_envs, b = n - tt * g.num_envs;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const int slot = (g.first_slot + tt + c) % g.ring_slots;
    const uint8_t* src = g.ring + ((size_t)b * g.ring_slots + slot) * kPlane + (4 * yp) * ARL_SCREEN + 4 * xp;
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) w[i] = __ldg(reinterpret_cast<const uint32_t*>(src + i * ARL_SCREEN));
    out[2 * c] = tc::bytes8_to_bf16(w[0], w[1]);
    out[2 * c + 1] = tc::bytes8_to_bf16(w[2], w[3]);
  }
}
// conv2: X2 row (n, y', x') chunk kc = (i*2+j)*2 + chalf <- a1[n][2y'+i][2x'+j][chalf*8 ..]
__device__ __forceinline__ const float* conv2_x_chunk(const float* a1, int n, int yp, int xp, int kc) {
  const int ij = kc >> 1, i = ij >> 1, j = ij & 1;
  return a1 + (int64_t)n * ARL_A1_ELEMS + ((2 * yp + i) * 20 + 2 * xp + j) * 16 + (kc & 1) * 8;
}

// =================================== conv1 forward ============================================
struct Conv1FwdArgs {
  const float* params;
  RingGeo geo;
  float* a1;
  int64_t rows;          // 441 * num_samples (grid rows)
  int num_samples;
};
struct Conv1Fwd {
  using Args = Conv1FwdArgs;
  static constexpr int GW = 21, GROWS = 441, TROWS = 150;
  static constexpr int PL = (TROWS + 1) * 16;                 // 2416: plane of one k-chunk
  static constexpr int PROD_WARPS = 16, STAGES = 8, STAGE_BYTES = 8 * PL;      // hi only
  static constexpr int PLB = 17 * 16, B_IMG = 32 * PLB, RES_BYTES = 2 * B_IMG;
  static constexpr int ACC_COLS = 16;
  static __device__ __forceinline__ int num_items(const Args& g) { return (int)((g.rows + 127) / 128); }
  static __device__ __forceinline__ TileCoord coord(const Args&, int item) { return row_tile(item); }
  static __device__ __forceinline__ int num_stages(const Args&, const TileCoord&) { return 1; }
  static __device__ __forceinline__ void load_resident(const Args& g, uint8_t* res, int ptid, int nthr) {
    for (int ch = ptid; ch < 16 * 32; ch += nthr) {
      const int co = ch & 15, kc = ch >> 4;                    // kc = tap*8 + c*2 + h
      const int tap = kc >> 3, c = (kc >> 1) & 3, h = kc & 1, a = tap >> 1, b = tap & 1;
      float x[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int i = 2 * h + (e >> 2), j = e & 3;
        x[e] = g.params[(((4 * a + i) * 8 + 4 * b + j) * 4 + c) * 16 + co];
      }
      tc::store_chunk_split(res, res + B_IMG, kc * PLB + co * 16, x);
    }
  }
  static __device__ __forceinline__ void load_stage(const Args& g, const TileCoord& t, int,
                                                    uint8_t* st, int glane, int gsize) {
    for (int r = glane; r < TROWS; r += gsize) {
      const int64_t xr = (int64_t)t.mt * 128 + r;
      const int n = (int)(xr / GROWS);
      uint4 ch[8];
      if (n < g.num_samples) {
        const int q = (int)(xr - (int64_t)n * GROWS), yp = q / GW, xp = q - yp * GW;
        conv1_x_row(g.geo, n, yp, xp, ch);
      } else {
#pragma unroll
        for (int k = 0; k < 8; ++k) ch[k] = make_uint4(0u, 0u, 0u, 0u);
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) *reinterpret_cast<uint4*>(st + k * PL + r * 16) = ch[k];
    }
  }
  static __device__ __forceinline__ void issue(const Args&, const TileCoord&, int, uint32_t st,
                                               uint32_t res, uint32_t d) {
    constexpr uint32_t idesc = tc::make_idesc(16);
#pragma unroll
    for (int tap = 0; tap < 4; ++tap) {
      const uint32_t a0 = st + ((tap >> 1) * GW + (tap & 1)) * 16;
#pragma unroll
      for (int k16 = 0; k16 < 4; ++k16) {
        const uint64_t da = tc::make_sdesc(a0 + 2 * k16 * PL, PL);
        const uint32_t b0 = res + (tap * 8 + 2 * k16) * PLB;
        tc::umma_f16(d, da, tc::make_sdesc(b0, PLB), idesc, (tap | k16) != 0 ? 1u : 0u);
        tc::umma_f16(d, da, tc::make_sdesc(b0 + B_IMG, PLB), idesc, 1u);
      }
    }
  }
  static __device__ __forceinline__ void store(const Args& g, const TileCoord& t, int row, int,
                                               const float (&v)[16]) {
    const int64_t xr = (int64_t)t.mt * 128 + row;
    const int n = (int)(xr / GROWS);
    const int q = (int)(xr - (int64_t)n * GROWS), yp = q / GW, xp = q - yp * GW;
    if (n >= g.num_samples || yp >= 20 || xp >= 20) return;
    const float* bias = g.params + 4096;
    float* d = g.a1 + ((int64_t)n * 400 + yp * 20 + xp) * 16;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float4 bb = *reinterpret_cast<const float4*>(bias + 4 * k);
      float4 o;
      o.x = fmaxf(fmaf(v[4 * k], 1.0f / 255.0f, bb.x), 0.f);
      o.y = fmaxf(fmaf(v[4 * k + 1], 1.0f / 255.0f, bb.y), 0.f);
      o.z = fmaxf(fmaf(v[4 * k + 2], 1.0f / 255.0f, bb.z), 0.f);
      o.w = fmaxf(fmaf(v[4 * k + 3], 1.0f / 255.0f, bb.w), 0.f);
      *reinterpret_cast<float4*>(d + 4 * k) = o;
    }
  }
};

// =================================== conv2 forward ============================================
struct Conv2FwdArgs {
  const float* params;
  const float* a1;
  float* a2;
  int64_t rows;          // 100 * num_samples
  int num_samples;
};
struct Conv2Fwd {
  using Args = Conv2FwdArgs;
  static constexpr int GW = 10, GROWS = 100, TROWS = 140;
  static constexpr int PL = (TROWS + 1) * 16, IMG = 8 * PL;   // 2256, 18048
  static constexpr int PROD_WARPS = 16, STAGES = 4, STAGE_BYTES = 2 * IMG;     // hi + lo
  static constexpr int PLB = 33 * 16, B_IMG = 32 * PLB, RES_BYTES = 2 * B_IMG;
  static constexpr int ACC_COLS = 32;
  static __device__ __forceinline__ int num_items(const Args& g) { return (int)((g.rows + 127) / 128); }
  static __device__ __forceinline__ TileCoord coord(const Args&, int item) { return row_tile(item); }
  static __device__ __forceinline__ int num_stages(const Args&, const TileCoord&) { return 1; }
  static __device__ __forceinline__ void load_resident(const Args& g, uint8_t* res, int ptid, int nthr) {
    const float* w2 = g.params + 4112;
    for (int ch = ptid; ch < 32 * 32; ch += nthr) {
      const int co = ch & 31, kc = ch >> 5;                    // kc = tap*8 + (i*2+j)*2 + chalf
      const int tap = kc >> 3, ij = (kc >> 1) & 3, chalf = kc & 1;
      const int kh = 2 * (tap >> 1) + (ij >> 1), kw = 2 * (tap & 1) + (ij & 1);
      float x[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) x[e] = w2[((kh * 4 + kw) * 16 + chalf * 8 + e) * 32 + co];
      tc::store_chunk_split(res, res + B_IMG, kc * PLB + co * 16, x);
    }
  }
  static __device__ __forceinline__ void load_stage(const Args& g, const TileCoord& t, int,
                                                    uint8_t* st, int glane, int gsize) {
    for (int r = glane; r < TROWS; r += gsize) {
      const int64_t xr = (int64_t)t.mt * 128 + r;
      const int n = (int)(xr / GROWS);
      const bool ok = n < g.num_samples;
      const int q = (int)(xr - (int64_t)n * GROWS), yp = q / GW, xp = q - yp * GW;
#pragma unroll
      for (int kc = 0; kc < 8; ++kc) {
        float x[8];
        if (ok) load8(conv2_x_chunk(g.a1, n, yp, xp, kc), x); else zero8(x);
        tc::store_chunk_split(st, st + IMG, kc * PL + r * 16, x);
      }
    }
  }
  static __device__ __forceinline__ void issue(const Args&, const TileCoord&, int, uint32_t st,
                                               uint32_t res, uint32_t d) {
    constexpr uint32_t idesc = tc::make_idesc(32);
#pragma unroll
    for (int tap = 0; tap < 4; ++tap) {
      const uint32_t a0 = st + ((tap >> 1) * GW + (tap & 1)) * 16;
#pragma unroll
      for (int k16 = 0; k16 < 4; ++k16) {
        const uint64_t da_hi = tc::make_sdesc(a0 + 2 * k16 * PL, PL);
        const uint64_t da_lo = tc::make_sdesc(a0 + IMG + 2 * k16 * PL, PL);
        const uint32_t b0 = res + (tap * 8 + 2 * k16) * PLB;
        const uint64_t db_hi = tc::make_sdesc(b0, PLB), db_lo = tc::make_sdesc(b0 + B_IMG, PLB);
        tc::umma_f16(d, da_hi, db_hi, idesc, (tap | k16) != 0 ? 1u : 0u);
        tc::umma_f16(d, da_hi, db_lo, idesc, 1u);
        tc::umma_f16(d, da_lo, db_hi, idesc, 1u);
      }
    }
  }
  static __device__ __forceinline__ void store(const Args& g, const TileCoord& t, int row, int c,
                                               const float (&v)[16]) {
    const int64_t xr = (int64_t)t.mt * 128 + row;
    const int n = (int)(xr / GROWS);
    const int q = (int)(xr - (int64_t)n * GROWS), yp = q / GW, xp = q - yp * GW;
    if (n >= g.num_samples || yp >= 9 || xp >= 9) return;
    const float* bias = g.params + 4112 + 8192 + c;
    float* d = g.a2 + ((int64_t)n * 81 + yp * 9 + xp) * 32 + c;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float4 bb = *reinterpret_cast<const float4*>(bias + 4 * k);
      float4 o;
      o.x = fmaxf(v[4 * k] + bb.x, 0.f);
      o.y = fmaxf(v[4 * k + 1] + bb.y, 0.f);
      o.z = fmaxf(v[4 * k + 2] + bb.z, 0.f);
      o.w = fmaxf(v[4 * k + 3] + bb.w, 0.f);
      *reinterpret_cast<float4*>(d + 4 * k) = o;
    }
  }
};

// =================================== conv2 input gradient =====================================
struct Conv2DgradArgs {
  const float* params;
  const float* a1;       // relu mask of conv1
  const float* dy2;      // [N, 81, 32]
  float* dy1;            // [N, 400, 16]
  int64_t rows;          // 121 * num_samples
  int num_samples;
};
struct Conv2Dgrad {
  using Args = Conv2DgradArgs;
  static constexpr int GW = 11, GROWS = 121, TROWS = 140;
  static constexpr int PL = (TROWS + 1) * 16, IMG = 4 * PL;   // 32 co = 4 chunks
  static constexpr int PROD_WARPS = 16, STAGES = 8, STAGE_BYTES = 2 * IMG;
  static constexpr int PLB = 17 * 16, B_IMG = 16 * PLB, B_CLS = 2 * B_IMG, RES_BYTES = 4 * B_CLS;
  static constexpr int ACC_COLS = 64;                         // 4 parity classes x 16 channels
  static __device__ __forceinline__ int num_items(const Args& g) { return (int)((g.rows + 127) / 128); }
  static __device__ __forceinline__ TileCoord coord(const Args&, int item) { return row_tile(item); }
  static __device__ __forceinline__ int num_stages(const Args&, const TileCoord&) { return 1; }
  static __device__ __forceinline__ void load_resident(const Args& g, uint8_t* res, int ptid, int nthr) {
    const float* w2 = g.params + 4112;
    for (int ch = ptid; ch < 4 * 16 * 16; ch += nthr) {
      const int c = ch & 15, kc = (ch >> 4) & 15, cls = ch >> 8;   // kc = tap*4 + co8
      const int tap = kc >> 2, co8 = kc & 3;
      const int kh = (cls >> 1) + 2 * (tap >> 1), kw = (cls & 1) + 2 * (tap & 1);
      float x[8];
      load8(w2 + ((kh * 4 + kw) * 16 + c) * 32 + co8 * 8, x);
      uint8_t* base = res + cls * B_CLS;
      tc::store_chunk_split(base, base + B_IMG, kc * PLB + c * 16, x);
    }
  }
  static __device__ __forceinline__ void load_stage(const Args& g, const TileCoord& t, int,
                                                    uint8_t* st, int glane, int gsize) {
    for (int r = glane; r < TROWS; r += gsize) {
      const int64_t zr = (int64_t)t.mt * 128 + r;
      const int n = (int)(zr / GROWS);
      const int q = (int)(zr - (int64_t)n * GROWS), zy = q / GW, zx = q - zy * GW;
      const bool ok = n < g.num_samples && zy >= 1 && zy <= 9 && zx >= 1 && zx <= 9;
      const float* src = g.dy2 + (int64_t)n * ARL_A2_ELEMS + ((zy - 1) * 9 + (zx - 1)) * 32;
#pragma unroll
      for (int kc = 0; kc < 4; ++kc) {
        float x[8];
        if (ok) load8(src + kc * 8, x); else zero8(x);
        tc::store_chunk_split(st, st + IMG, kc * PL + r * 16, x);
      }
    }
  }
  static __device__ __forceinline__ void issue(const Args&, const TileCoord&, int, uint32_t st,
                                               uint32_t res, uint32_t d) {
    constexpr uint32_t idesc = tc::make_idesc(16);
#pragma unroll 1
    for (int cls = 0; cls < 4; ++cls) {
#pragma unroll
      for (int tap = 0; tap < 4; ++tap) {
        // output (yy,xx) of this class reads Z row P + 12 - 11*dkh - dkw
        const uint32_t a0 = st + (12 - 11 * (tap >> 1) - (tap & 1)) * 16;
#pragma unroll
        for (int k16 = 0; k16 < 2; ++k16) {
          const uint64_t da_hi = tc::make_sdesc(a0 + 2 * k16 * PL, PL);
          const uint64_t da_lo = tc::make_sdesc(a0 + IMG + 2 * k16 * PL, PL);
          const uint32_t b0 = res + cls * B_CLS + (tap * 4 + 2 * k16) * PLB;
          const uint64_t db_hi = tc::make_sdesc(b0, PLB), db_lo = tc::make_sdesc(b0 + B_IMG, PLB);
          tc::umma_f16(d + cls * 16, da_hi, db_hi, idesc, (tap | k16) != 0 ? 1u : 0u);
          tc::umma_f16(d + cls * 16, da_hi, db_lo, idesc, 1u);
          tc::umma_f16(d + cls * 16, da_lo, db_hi, idesc, 1u);
        }
      }
    }
  }
  static __device__ __forceinline__ void store(const Args& g, const TileCoord& t, int row, int c,
                                               const float (&v)[16]) {
    const int64_t pr = (int64_t)t.mt * 128 + row;
    const int n = (int)(pr / GROWS);
    const int q = (int)(pr - (int64_t)n * GROWS), yy = q / GW, xx = q - yy * GW;
    if (n >= g.num_samples || yy >= 10 || xx >= 10) return;
    const int cls = c >> 4, y = 2 * yy + (cls >> 1), x = 2 * xx + (cls & 1);
    const int64_t off = (int64_t)n * ARL_A1_ELEMS + (y * 20 + x) * 16;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float4 m = __ldg(reinterpret_cast<const float4*>(g.a1 + off) + k);
      float4 o;
      o.x = m.x > 0.f ? v[4 * k] : 0.f;
      o.y = m.y > 0.f ? v[4 * k + 1] : 0.f;
      o.z = m.z > 0.f ? v[4 * k + 2] : 0.f;
      o.w = m.w > 0.f ? v[4 * k + 3] : 0.f;
      *(reinterpret_cast<float4*>(g.dy1 + off) + k) = o;
    }
  }
};

// =================================== weight gradients =========================================
__device__ __forceinline__ TileCoord range_tile(int item, int64_t rows, int k_chunk) {
  TileCoord t;
  t.mt = item; t.nt = 0; t.ks = item;
  const int64_t b = (int64_t)item * k_chunk;
  t.k_begin = (int)b;
  t.k_end = (int)(b + k_chunk < rows ? b + k_chunk : rows);
  return t;
}

struct Conv2WgradArgs {
  const float* a1;
  const float* dy2;
  float* partials;       // [items][8192]
  int64_t rows;          // 100 * num_samples
  int num_samples, k_chunk, items;
};
struct Conv2Wgrad {
  using Args = Conv2WgradArgs;
  static constexpr int GW = 10, GROWS = 100, TROWS = 140;
  static constexpr int PLA = (TROWS + 1) * 16, A_IMG = 16 * PLA;   // 2 shifted copies x 8 ch groups
  static constexpr int PLB = 129 * 16, B_IMG = 4 * PLB;            // dy2z: 4 co groups
  static constexpr int PROD_WARPS = 16, STAGES = 2, STAGE_BYTES = 2 * A_IMG + 2 * B_IMG, RES_BYTES = 0;
  static constexpr int ACC_COLS = 64;                              // tap a in {0,1} x 32 co
  static __device__ __forceinline__ int num_items(const Args& g) { return g.items; }
  static __device__ __forceinline__ TileCoord coord(const Args& g, int item) {
    return range_tile(item, g.rows, g.k_chunk);
  }
  static __device__ __forceinline__ int num_stages(const Args&, const TileCoord& t) {
    return (t.k_end - t.k_begin + 127) / 128;
  }
  static __device__ __forceinline__ void load_resident(const Args&, uint8_t*, int, int) {}
  static __device__ __forceinline__ void load_stage(const Args& g, const TileCoord& t, int s,
                                                    uint8_t* st, int glane, int gsize) {
    const int64_t p0 = (int64_t)t.k_begin + (int64_t)s * 128;
    uint8_t* a_hi = st, *a_lo = st + A_IMG, *b_hi = st + 2 * A_IMG, *b_lo = b_hi + B_IMG;
    // A: X2 rows p0 .. p0+139; copy 0 holds row r at r*16, copy 1 (ch groups 8..15) holds row r
    // at (r-1)*16, i.e. copy 1 is the image shifted by one row (tap b = 1)
    for (int c = glane; c < TROWS * 8; c += gsize) {
      const int r = c % TROWS, kc = c / TROWS;
      const int64_t xr = p0 + r;
      const int n = (int)(xr / GROWS);
      const int q = (int)(xr - (int64_t)n * GROWS), yp = q / GW, xp = q - yp * GW;
      float x[8];
      if (n < g.num_samples) load8(conv2_x_chunk(g.a1, n, yp, xp, kc), x); else zero8(x);
      uint4 h, l;
      tc::split2(x[0], x[1], h.x, l.x); tc::split2(x[2], x[3], h.y, l.y);
      tc::split2(x[4], x[5], h.z, l.z); tc::split2(x[6], x[7], h.w, l.w);
      *reinterpret_cast<uint4*>(a_hi + kc * PLA + r * 16) = h;
      *reinterpret_cast<uint4*>(a_lo + kc * PLA + r * 16) = l;
      if (r > 0) {
        *reinterpret_cast<uint4*>(a_hi + (8 + kc) * PLA + (r - 1) * 16) = h;
        *reinterpret_cast<uint4*>(a_lo + (8 + kc) * PLA + (r - 1) * 16) = l;
      }
    }
    // B: dy2 on the 10-wide grid, zero at y'=9 / x'=9 and outside [k_begin, k_end)
    for (int c = glane; c < 128 * 4; c += gsize) {
      const int r = c & 127, kc = c >> 7;
      const int64_t pr = p0 + r;
      const int n = (int)(pr / GROWS);
      const int q = (int)(pr - (int64_t)n * GROWS), yp = q / GW, xp = q - yp * GW;
      float x[8];
      if (pr < t.k_end && n < g.num_samples && yp < 9 && xp < 9)
        load8(g.dy2 + (int64_t)n * ARL_A2_ELEMS + (yp * 9 + xp) * 32 + kc * 8, x);
      else
        zero8(x);
      tc::store_chunk_split(b_hi, b_lo, kc * PLB + r * 16, x);
    }
  }
  static __device__ __forceinline__ void issue(const Args&, const TileCoord&, int s, uint32_t st,
                                               uint32_t, uint32_t d) {
    constexpr uint32_t idesc = tc::make_idesc(32, true, true);
    const uint32_t a_hi = st, a_lo = st + A_IMG, b_hi = st + 2 * A_IMG, b_lo = b_hi + B_IMG;
#pragma unroll
    for (int a = 0; a < 2; ++a) {
#pragma unroll
      for (int k16 = 0; k16 < 8; ++k16) {
        const uint32_t ao = (a * GW + k16 * 16) * 16, bo = k16 * 256;
        const uint64_t da_hi = tc::make_sdesc(a_hi + ao, 128, PLA), da_lo = tc::make_sdesc(a_lo + ao, 128, PLA);
        const uint64_t db_hi = tc::make_sdesc(b_hi + bo, 128, PLB), db_lo = tc::make_sdesc(b_lo + bo, 128, PLB);
        tc::umma_f16(d + a * 32, da_hi, db_hi, idesc, (s | k16) != 0 ? 1u : 0u);
        tc::umma_f16(d + a * 32, da_hi, db_lo, idesc, 1u);
        tc::umma_f16(d + a * 32, da_lo, db_hi, idesc, 1u);
      }
    }
  }
  static __device__ __forceinline__ void store(const Args& g, const TileCoord& t, int row, int c,
                                               const float (&v)[16]) {
    // row = b*64 + ch, ch = (i*2+j)*16 + cin ; column c = a*32 + co
    const int b = row >> 6, ch = row & 63, ij = ch >> 4, cin = ch & 15;
    const int a = c >> 5, co = c & 31;
    const int kh = 2 * a + (ij >> 1), kw = 2 * b + (ij & 1);
    float* d = g.partials + (size_t)t.ks * 8192 + ((kh * 4 + kw) * 16 + cin) * 32 + co;
#pragma unroll
    for (int k = 0; k < 4; ++k)
      *reinterpret_cast<float4*>(d + 4 * k) = make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
  }
};

struct Conv1WgradArgs {
  RingGeo geo;
  const float* dy1;
  float* partials;       // [items][4096]
  int64_t rows;          // 441 * num_samples
  int num_samples, k_chunk, items;
};
struct Conv1Wgrad {
  using Args = Conv1WgradArgs;
  static constexpr int GW = 21, GROWS = 441, TROWS = 150;
  static constexpr int PLA = (TROWS + 1) * 16, A_IMG = 16 * PLA;   // exact bf16: hi only
  static constexpr int PLB = 129 * 16, B_IMG = 2 * PLB;            // dy1z: 2 co groups
  static constexpr int PROD_WARPS = 16, STAGES = 4, STAGE_BYTES = A_IMG + 2 * B_IMG, RES_BYTES = 0;
  static constexpr int ACC_COLS = 32;                              // tap a in {0,1} x 16 co
  static __device__ __forceinline__ int num_items(const Args& g) { return g.items; }
  static __device__ __forceinline__ TileCoord coord(const Args& g, int item) {
    return range_tile(item, g.rows, g.k_chunk);
  }
  static __device__ __forceinline__ int num_stages(const Args&, const TileCoord& t) {
    return (t.k_end - t.k_begin + 127) / 128;
  }
  static __device__ __forceinline__ void load_resident(const Args&, uint8_t*, int, int) {}
  static __device__ __forceinline__ void load_stage(const Args& g, const TileCoord& t, int s,
                                                    uint8_t* st, int glane, int gsize) {
    const int64_t p0 = (int64_t)t.k_begin + (int64_t)s * 128;
    uint8_t* a_img = st, *b_hi = st + A_IMG, *b_lo = b_hi + B_IMG;
    for (int r = glane; r < TROWS; r += gsize) {
      const int64_t xr = p0 + r;
      const int n = (int)(xr / GROWS);
      uint4 ch[8];
      if (n < g.num_samples) {
        const int q = (int)(xr - (int64_t)n * GROWS), yp = q / GW, xp = q - yp * GW;
        conv1_x_row(g.geo, n, yp, xp, ch);
      } else {
#pragma unroll
        for (int k = 0; k < 8; ++k) ch[k] = make_uint4(0u, 0u, 0u, 0u);
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        *reinterpret_cast<uint4*>(a_img + k * PLA + r * 16) = ch[k];
        if (r > 0) *reinterpret_cast<uint4*>(a_img + (8 + k) * PLA + (r - 1) * 16) = ch[k];
      }
    }
    for (int c = glane; c < 128 * 2; c += gsize) {
      const int r = c & 127, kc = c >> 7;
      const int64_t pr = p0 + r;
      const int n = (int)(pr / GROWS);
      const int q = (int)(pr - (int64_t)n * GROWS), yp = q / GW, xp = q - yp * GW;
      float x[8];
      if (pr < t.k_end && n < g.num_samples && yp < 20 && xp < 20)
        load8(g.dy1 + ((int64_t)n * 400 + yp * 20 + xp) * 16 + kc * 8, x);
      else
        zero8(x);
      tc::store_chunk_split(b_hi, b_lo, kc * PLB + r * 16, x);
    }
  }
  static __device__ __forceinline__ void issue(const Args&, const TileCoord&, int s, uint32_t st,
                                               uint32_t, uint32_t d) {
    constexpr uint32_t idesc = tc::make_idesc(16, true, true);
    const uint32_t a_img = st, b_hi = st + A_IMG, b_lo = b_hi + B_IMG;
#pragma unroll
    for (int a = 0; a < 2; ++a) {
#pragma unroll
      for (int k16 = 0; k16 < 8; ++k16) {
        const uint64_t da = tc::make_sdesc(a_img + (a * GW + k16 * 16) * 16, 128, PLA);
        tc::umma_f16(d + a * 16, da, tc::make_sdesc(b_hi + k16 * 256, 128, PLB), idesc,
                     (s | k16) != 0 ? 1u : 0u);
        tc::umma_f16(d + a * 16, da, tc::make_sdesc(b_lo + k16 * 256, 128, PLB), idesc, 1u);
      }
    }
  }
  static __device__ __forceinline__ void store(const Args& g, const TileCoord& t, int row, int c,
                                               const float (&v)[16]) {
    // row = b*64 + ch, ch = (cin*4 + i)*4 + j ; column c = a*16 + co
    const int b = row >> 6, ch = row & 63, cin = ch >> 4, i = (ch >> 2) & 3, j = ch & 3;
    const int a = c >> 4;
    const int kh = 4 * a + i, kw = 4 * b + j;
    float* d = g.partials + (size_t)t.ks * 4096 + ((kh * 8 + kw) * 4 + cin) * 16;
    const float s = 1.0f / 255.0f;
#pragma unroll
    for (int k = 0; k < 4; ++k)
      *reinterpret_cast<float4*>(d + 4 * k) =
          make_float4(v[4 * k] * s, v[4 * k + 1] * s, v[4 * k + 2] * s, v[4 * k + 3] * s);
  }
};

// partial column sums of X [rows, COLS] (COLS = 16 or 32): block b owns rows b, b+grid, ...
template <int COLS>
__global__ void colsum_small_kernel(const float* __restrict__ X, float* __restrict__ partials,
                                    int64_t rows) {
  __shared__ float red[256];
  const int col = threadIdx.x % COLS, sub = threadIdx.x / COLS;
  constexpr int SUBS = 256 / COLS;
  float s = 0.f;
  for (int64_t r = (int64_t)blockIdx.x * SUBS + sub; r < rows; r += (int64_t)gridDim.x * SUBS)
    s += X[r * COLS + col];
  red[threadIdx.x] = s;
  __syncthreads();
  if (sub == 0) {
    float v = 0.f;
#pragma unroll
    for (int k = 0; k < SUBS; ++k) v += red[k * COLS + col];
    partials[(size_t)blockIdx.x * COLS + col] = v;
  }
}

int split_rows(int64_t rows, int want, int* k_chunk) {
  int64_t per = (rows + want - 1) / want;
  per = (per + 127) / 128 * 128;
  *k_chunk = (int)per;
  return (int)((rows + per - 1) / per);
}

}  // namespace
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
  ARL_REQUIRE(N * 441 / 128 < (1LL << 31) - 2, "arl_conv1_forward: too many samples");
  Conv1FwdArgs g{params, {ring, num_envs, ring_slots, first_slot}, a1, N * 441, (int)N};
  return tc::launch<Conv1Fwd>(g, (int)((g.rows + 127) / 128), (cudaStream_t)stream);
}

extern "C" int arl_conv2_forward(const float* params, const float* a1, float* a2,
                                 int64_t num_samples, void* stream) {
  ARL_REQUIRE(params && a1 && a2, "arl_conv2_forward: null pointer");
  ARL_REQUIRE(num_samples >= 0 && num_samples < (1LL << 24), "arl_conv2_forward: bad num_samples");
  ARL_REQUIRE(aligned16(params) && aligned16(a1) && aligned16(a2),
              "arl_conv2_forward: pointers must be 16-byte aligned");
  if (num_samples == 0) return ARL_OK;
  Conv2FwdArgs g{params, a1, a2, num_samples * 100, (int)num_samples};
  return tc::launch<Conv2Fwd>(g, (int)((g.rows + 127) / 128), (cudaStream_t)stream);
}

extern "C" int arl_conv1_backward(const uint8_t* ring, const float* d_a1, float* grads,
                                  void* workspace, int num_envs, int ring_slots, int first_slot,
                                  int steps, void* stream) {
  ARL_REQUIRE(ring && d_a1 && grads && workspace, "arl_conv1_backward: null pointer");
  ARL_REQUIRE(num_envs >= 0 && steps >= 0, "arl_conv1_backward: negative size");
  ARL_REQUIRE(ring_slots >= steps + 3 && first_slot >= 0 && first_slot < ring_slots,
              "arl_conv1_backward: ring geometry (slots %d, steps %d, first %d)", ring_slots,
              steps, first_slot);
  ARL_REQUIRE(aligned16(ring) && aligned16(d_a1) && aligned16(grads) && aligned16(workspace),
              "arl_conv1_backward: pointers must be 16-byte aligned");
  const int64_t N = (int64_t)num_envs * steps;
  cudaStream_t st = (cudaStream_t)stream;
  if (N == 0) {
    ARL_CUDA(cudaMemsetAsync(grads, 0, (4096 + 16) * sizeof(float), st));
    return ARL_OK;
  }
  ARL_REQUIRE(N * 441 < (1LL << 31) - 256, "arl_conv1_backward: too many samples");
  Conv1WgradArgs g;
  g.geo = {ring, num_envs, ring_slots, first_slot};
  g.dy1 = d_a1;
  g.partials = (float*)workspace;
  g.rows = N * 441;
  g.num_samples = (int)N;
  g.items = split_rows(g.rows, num_sms(), &g.k_chunk);
  int rc = tc::launch<Conv1Wgrad>(g, g.items, st);
  if (rc) return rc;
  rc = reduce_partials(g.partials, grads, g.items, 4096, st);                    // l1_w
  if (rc) return rc;
  float* part = (float*)workspace + (size_t)g.items * 4096;
  const int grid = num_sms();
  colsum_small_kernel<16><<<grid, 256, 0, st>>>(d_a1, part, N * 400);
  ARL_LAUNCH_CHECK("colsum_small_kernel<16>");
  return reduce_partials(part, grads + 4096, grid, 16, st);                      // l1_b
}

extern "C" int arl_conv2_backward(const float* params, const float* a1, const float* d_a2,
                                  float* d_a1, float* grads, void* workspace, int64_t num_samples,
                                  void* stream) {
  ARL_REQUIRE(params && a1 && d_a2 && d_a1 && grads && workspace,
              "arl_conv2_backward: null pointer");
  ARL_REQUIRE(num_samples >= 0 && num_samples < (1LL << 24), "arl_conv2_backward: bad num_samples");
  ARL_REQUIRE(aligned16(params) && aligned16(a1) && aligned16(d_a2) && aligned16(d_a1) &&
                  aligned16(grads) && aligned16(workspace),
              "arl_conv2_backward: pointers must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  float* g2 = grads + 4096 + 16;                                                 // l2_w | l2_b
  if (num_samples == 0) {
    ARL_CUDA(cudaMemsetAsync(g2, 0, (8192 + 32) * sizeof(float), st));
    return ARL_OK;
  }
  Conv2WgradArgs w;
  w.a1 = a1;
  w.dy2 = d_a2;
  w.partials = (float*)workspace;
  w.rows = num_samples * 100;
  w.num_samples = (int)num_samples;
  w.items = split_rows(w.rows, num_sms(), &w.k_chunk);
  int rc = tc::launch<Conv2Wgrad>(w, w.items, st);
  if (rc) return rc;
  rc = reduce_partials(w.partials, g2, w.items, 8192, st);
  if (rc) return rc;
  float* part = (float*)workspace + (size_t)w.items * 8192;
  const int grid = num_sms();
  colsum_small_kernel<32><<<grid, 256, 0, st>>>(d_a2, part, num_samples * 81);
  ARL_LAUNCH_CHECK("colsum_small_kernel<32>");
  rc = reduce_partials(part, g2 + 8192, grid, 32, st);
  if (rc) return rc;
  Conv2DgradArgs d{params, a1, d_a2, d_a1, num_samples * 121, (int)num_samples};
  return tc::launch<Conv2Dgrad>(d, (int)((d.rows + 127) / 128), st);
}
