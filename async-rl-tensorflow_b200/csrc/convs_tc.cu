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
#include <stdlib.h>

#include "gemm_tc.cuh"

namespace arl {

int reduce_partials(const float* partials, float* out, int num_partials, int n, cudaStream_t stream);
int reduce_partials_scaled(const float* partials, float* out, int num_partials, int n, float alpha,
                           cudaStream_t stream);

namespace {
using tc::TileCoord;
using tc::kTileM;

__device__ __forceinline__ void load8(const float* p, float (&x)[8]) {
  const float4 x0 = __ldg(reinterpret_cast<const float4*>(p));
  const float4 x1 = __ldg(reinterpret_cast<const float4*>(p) + 1);
  x[0] = x0.x; x[1] = x0.y; x[2] = x0.z; x[3] = x0.w;
  x[4] = x1.x; x[5] = x1.y; x[6] = x1.z; x[7] = x1.w;
}
__device__ __forceinline__ TileCoord row_tile(int item) {
  TileCoord t;
  t.mt = item; t.nt = 0; t.ks = 0; t.k_begin = 0; t.k_end = 0;
  return t;
}

// ---- pixel-contiguous fp32 tensors scattered onto a padded row grid --------------------------
// A tensor [samples][PH*PW pixels][4*CF4 floats] (dy1, dy2) is laid on a per-sample grid of
// GROWS = rows of width GW with the pixel (py,px) at grid row (py+OFF)*GW + (px+OFF); the other
// grid rows are zero.  A stage needs the grid rows [z0, z0+ROWS).  The pixels that land there are
// ONE contiguous range of the tensor, so the producers stream that range with warp-contiguous
// float4 loads (lane = 16 consecutive bytes) and scatter each float4 to its image position;
// a second small pass zero-fills the rows no pixel maps to.  All row/pixel indices fit in int32
// (checked by the entry points).
template <int GW, int GROWS, int PH, int PW, int OFF, int CF4>
struct PixelGrid {
  static constexpr int PP = PH * PW, WIDTH = PW;
  // a pixel index that is <= the first pixel whose grid row is >= z
  static __device__ __forceinline__ int lower_pixel(int z) {
    const int n = z / GROWS, rem = z - n * GROWS, gy = rem / GW, gx = rem - gy * GW;
    const int py = min(max(gy - OFF, 0), PH - 1), px = min(max(gx - OFF, 0), PW - 1);
    return n * PP + max(py * PW + px - PW, 0);
  }
  static __device__ __forceinline__ int grid_of_pixel(int pix) {
    const int n = pix / PP, rem = pix - n * PP, py = rem / PW, px = rem - py * PW;
    return n * GROWS + (py + OFF) * GW + (px + OFF);
  }
  static __device__ __forceinline__ bool row_has_pixel(int z, int num_samples) {
    const int n = z / GROWS, rem = z - n * GROWS, gy = rem / GW, gx = rem - gy * GW;
    return n < num_samples && gy >= OFF && gy < PH + OFF && gx >= OFF && gx < PW + OFF;
  }
};

// Streams the pixels of grid rows [z0, z0+ROWS) (rows >= z_end count as zero) into a hi/lo image
// pair whose 16-B vector (row r, chunk kc) sits at kc*PL + r*16; FP16 chooses the element format of
// the pair (fp16 hi + lo next to fp16 activations, bf16 hi + lo next to bf16 weights: the two
// operands of one MMA must share a format).  U float4 loads are in flight per
// lane before the first one is consumed (U * gsize >= the whole window => one latency per stage).
// acc (optional) accumulates the column sums of what was stored: a lane owns floats
// (glane % CF4)*4 .. +3 of every pixel it touches.
template <class G, int CF4, int ROWS, int PL, int U, bool ACC, bool FP16>
__device__ __forceinline__ void stream_pixels(uint8_t* hi, uint8_t* lo, const float* __restrict__ src,
                                              int z0, int z_end, int num_samples, int glane,
                                              int gsize, float (&acc)[4]) {
  const int p0 = G::lower_pixel(z0);
  const int p_total = num_samples * G::PP;
  constexpr int SPAN = ROWS + 2 * G::WIDTH + 2;              // lower_pixel undershoots by <= 2*PW
  constexpr int TOTAL = SPAN * CF4;
  const float4* s4 = reinterpret_cast<const float4*>(src) + (int64_t)p0 * CF4;
  const int q = glane % CF4;                                 // gsize % CF4 == 0
  for (int f0 = glane; f0 < TOTAL; f0 += U * gsize) {
    float4 x[U];
    int rr[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int f = f0 + u * gsize;
      const int pix = p0 + f / CF4;
      x[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      rr[u] = -1;
      if (f < TOTAL && pix < p_total) {
        const int z = G::grid_of_pixel(pix);
        if (z >= z0 && z < z0 + ROWS) {
          rr[u] = z - z0;
          if (z < z_end) x[u] = __ldg(s4 + f);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (rr[u] < 0) continue;
      const int off = (q >> 1) * PL + rr[u] * 16 + (q & 1) * 8;
      if (FP16) tc::store_half_split_h(hi, lo, off, x[u]);
      else tc::store_half_split(hi, lo, off, x[u]);
      if (ACC) { acc[0] += x[u].x; acc[1] += x[u].y; acc[2] += x[u].z; acc[3] += x[u].w; }
    }
  }
  // rows of the window that no pixel maps to (grid padding, beyond the last sample)
  for (int r = glane; r < ROWS; r += gsize) {
    if (!G::row_has_pixel(z0 + r, num_samples)) {
#pragma unroll
      for (int kc = 0; kc < CF4 / 2; ++kc) {
        *reinterpret_cast<uint4*>(hi + kc * PL + r * 16) = make_uint4(0u, 0u, 0u, 0u);
        *reinterpret_cast<uint4*>(lo + kc * PL + r * 16) = make_uint4(0u, 0u, 0u, 0u);
      }
    }
  }
}

using GridDy2 = PixelGrid<10, 100, 9, 9, 0, 8>;       // dy2 [N,81,32] on the conv2 X2 grid
using GridZ   = PixelGrid<11, 121, 9, 9, 1, 8>;       // dy2 zero-padded by one (transposed conv)

// resident operand image: built once per parameter change by arl_prepare_weights (build_image of
// the policy), copied into shared memory by the producer threads in the kernel prologue
template <int BYTES>
__device__ __forceinline__ void copy_image(uint8_t* res, const uint8_t* __restrict__ img, int ptid, int nthr) {
  static_assert(BYTES % 16 == 0, "image size");
  // eight loads in flight per thread: with one at a time the 33-KB conv2 image took ~8 us of the
  // 45-us forward kernel (ncu: 19 % of the samples at the prologue barrier)
  constexpr int N = BYTES / 16, U = 8;
  for (int i0 = ptid; i0 < N; i0 += U * nthr) {
    uint4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (i0 + u * nthr < N) v[u] = __ldg(reinterpret_cast<const uint4*>(img) + i0 + u * nthr);
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (i0 + u * nthr < N) reinterpret_cast<uint4*>(res)[i0 + u * nthr] = v[u];
  }
}

// ---- conv1 A operand: space-to-depth rows of the u8 ring ---------------------------------------
// The ring stores every 84x84 plane in 4x4 blocks (K1 writes it that way): the 16 bytes at
// plane + q*16 are block q = y'*21 + x' = frame rows 4y'..4y'+3, columns 4x'..4x'+3 = the 16
// channels plane c contributes to X row (n, y', x') -> chunks kc = 2c (rows 0,1) and 2c+1 (rows
// 2,3).  Work unit = (row, plane): ONE 16-byte load; lanes run along the rows of one plane, so a
// warp instruction reads 512 contiguous bytes.  U units are in flight per lane.
// SHIFTED: also store the copy shifted by one row into planes 8..15 (tap b = 1 of the wgrad).
struct RingGeo {
  const uint8_t* ring;
  int num_envs, ring_slots, first_slot;
};
// split in two so that the loads of the NEXT stage can be in flight while the producer waits for
// its shared-memory slot (PolicyBase::PREFETCH): U * gsize >= ROWS * 4, one round
template <int ROWS, int U>
__device__ __forceinline__ void x1_request(uint4 (&w)[U], const RingGeo& g, int xr0, int num_samples,
                                           int glane, int gsize) {
  constexpr int GROWS = 441, TOTAL = ROWS * 4;
  const int n0 = xr0 / GROWS, q0 = xr0 - n0 * GROWS;
  const int tt0 = n0 / g.num_envs, b0 = n0 - tt0 * g.num_envs;
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const int un = glane + u * gsize;
    const int c = un / ROWS, r = un - c * ROWS;
    int q = q0 + r, n = n0, tt = tt0, b = b0;
    if (q >= GROWS) {                                      // ROWS < GROWS: at most one wrap
      q -= GROWS; ++n; ++b;
      if (b == g.num_envs) { b = 0; ++tt; }
    }
    w[u] = make_uint4(0u, 0u, 0u, 0u);
    if (un < TOTAL && n < num_samples) {
      int slot = g.first_slot + tt + c;
      slot -= slot >= g.ring_slots ? g.ring_slots : 0;
      slot -= slot >= g.ring_slots ? g.ring_slots : 0;
      w[u] = __ldg(reinterpret_cast<const uint4*>(g.ring + ((size_t)b * g.ring_slots + slot) * kPlane) + q);
    }
  }
}
template <int ROWS, int PL, int U, bool SHIFTED>
__device__ __forceinline__ void x1_store(uint8_t* img, const uint4 (&w)[U], int glane, int gsize) {
  constexpr int TOTAL = ROWS * 4;
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const int un = glane + u * gsize;
    if (un >= TOTAL) break;
    const int c = un / ROWS, r = un - c * ROWS;
    const uint4 lo2 = tc::bytes8_to_f16(w[u].x, w[u].y), hi2 = tc::bytes8_to_f16(w[u].z, w[u].w);   // exact
    uint8_t* d = img + (2 * c) * PL + r * 16;
    *reinterpret_cast<uint4*>(d) = lo2;
    *reinterpret_cast<uint4*>(d + PL) = hi2;
    if (SHIFTED && r > 0) {
      *reinterpret_cast<uint4*>(d + 8 * PL - 16) = lo2;
      *reinterpret_cast<uint4*>(d + 9 * PL - 16) = hi2;
    }
  }
}

// ---- a1 in HBM: fp16, blocked ("a1s") -----------------------------------------------------------
// conv1's output is only ever consumed as a tensor-core operand (conv2 forward / conv2 wgrad)
// and as a relu mask (conv2 dgrad), so it is stored the way those kernels want it -- ONE fp16 per
// value (12 800 bytes per sample, half of fp32 [20,20,16]) in space-to-depth order:
//   a1s[n][kc = (i*2+j)*2 + chalf][q = yp*10 + xp][8 channels]     (16-B vectors)
// with pixel (y,x) = (2yp+i, 2xp+j) and channels chalf*8 .. chalf*8+7.  A run of X2 rows of one
// plane kc is then contiguous: the conv2 kernels fetch their A images with cp.async.bulk and
// convert nothing.  (Round 1 kept a bf16 hi + lo pair, the bytes of fp32; tests/precision_study.py
// is the measurement behind the narrower format.)
constexpr int kA1sSample = 12800, kA1sPlane = 1600;    // bytes
__device__ __forceinline__ int a1s_offset(int y, int x, int chalf) {       // bytes inside a sample
  return (((y & 1) * 2 + (x & 1)) * 2 + chalf) * kA1sPlane + ((y >> 1) * 10 + (x >> 1)) * 16;
}

// Bulk copies of X2 grid rows [xr0, xr0 + ROWS) of plane kc into an image whose vector (row r, kc)
// sits at kc*PL + r*16; SHIFT additionally fills planes 8..15 with the image shifted by one row
// (tap b = 1 of the wgrad).  Called by the lanes that own one plane each; returns the bytes this
// lane has put in flight (it must expect_tx them BEFORE calling).
template <int ROWS, bool SHIFT>
__device__ __forceinline__ uint32_t a1s_bulk_bytes(int xr0, int num_samples) {
  // bytes one plane owner moves: every existing row once (+ once more, minus the first, if SHIFT)
  const int rows_left = num_samples * 100 - xr0;
  const int n = rows_left < ROWS ? (rows_left > 0 ? rows_left : 0) : ROWS;
  return (uint32_t)(SHIFT ? (n > 0 ? 2 * n - 1 : 0) : n) * 16u;
}
template <int ROWS, int PL, bool SHIFT>
__device__ __forceinline__ void a1s_bulk_issue(uint8_t* st, const uint8_t* a1s, int xr0, int num_samples,
                                               int kc, uint64_t* full) {
  int r = 0, n = xr0 / 100, q = xr0 - n * 100;
  while (r < ROWS && n < num_samples) {
    const int cnt = min(100 - q, ROWS - r);
    const uint8_t* src = a1s + (size_t)n * kA1sSample + kc * kA1sPlane + q * 16;
    uint8_t* dst = st + kc * PL + r * 16;
    bulk_g2s(dst, src, (uint32_t)cnt * 16u, full);
    if (SHIFT) {
      if (r > 0) bulk_g2s(dst + 8 * PL - 16, src, (uint32_t)cnt * 16u, full);
      else if (cnt > 1) bulk_g2s(dst + 8 * PL, src + 16, (uint32_t)(cnt - 1) * 16u, full);
    }
    r += cnt; q = 0; ++n;
  }
}
// rows of the window beyond the last sample (last tile only): zero, by all lanes of the group
template <int ROWS, int PL, int PLANES>
__device__ __forceinline__ void a1s_zero_tail(uint8_t* st, int xr0, int num_samples, int glane, int gsize) {
  const int valid = num_samples * 100 - xr0;
  if (valid >= ROWS) return;
  for (int c = glane; c < ROWS * PLANES; c += gsize) {
    const int r = c % ROWS, kc = c / ROWS;
    // planes 8.. (if any) hold the image shifted by one row: their row r is source row r+1
    if (r >= valid - (kc >= 8 ? 1 : 0))
      *reinterpret_cast<uint4*>(st + kc * PL + r * 16) = make_uint4(0u, 0u, 0u, 0u);
  }
}

// =================================== conv1 forward ============================================
struct Conv1FwdArgs {
  const uint8_t* w_img;  // prepared: s8 limb image of l1_w | 3 limb scales | l1_b
  RingGeo geo;
  uint8_t* a1s;          // fp16 blocked output (see a1s above), 12 800 B per sample
  int64_t rows;          // 441 * num_samples (grid rows)
  int num_samples;
};
// The ring planes ARE the tensor-core operand: a plane is 441 blocks of 16 bytes, and 16 bytes
// = 16 u8 K-values of one row is exactly the core-matrix row of the K-major no-swizzle layout for
// 8-bit operands.  So the A image of a tile is 4 x (rows x 16 B) copied verbatim by
// cp.async.bulk (no conversion, no LSU traffic), the MMA is kind::i8 (u8 x s8 -> s32, K = 32)
// and the fp32 weights enter as three s8 limbs, w = s*(L0/64 + L1/2^13 + L2/2^20) (residual
// < 5e-7 * max|w|), concatenated along N (N = 48: one MMA per K block instead of three).  The
// integer accumulation is exact; the epilogue recombines the limb sums in fp32.
struct Conv1Fwd : tc::PolicyBase {
  using Args = Conv1FwdArgs;
  static constexpr int GW = 21, GROWS = 441, TROWS = 150;
  static constexpr int PL = (TROWS + 1) * 16;                 // 2416: plane of one 16-channel chunk
  // two epilogue sets: with bulk-copied operands and 8 short MMAs per tile the accumulator
  // read-out is the longest stage of the pipeline (four sets measured: no further gain)
  static constexpr int EPI_SETS = 2, PROD_WARPS = 8, STAGES = 8, STAGE_BYTES = 4 * PL;   // u8: 4 planes
  // resident W1 image (s8): rows = limb*16 + co (N = 48), 16 k-chunk planes (tap*4 + c); then
  // the three limb scales
  static constexpr int PLB = 49 * 16, B_IMG = 16 * PLB, SCALE_OFF = B_IMG, BIAS_OFF = B_IMG + 16;
  static constexpr int RES_BYTES = B_IMG + 16 + 64;
  static constexpr int ACC_COLS = 64, OUT_COLS = 16, LO_DELTA = 16, SEG = 16;
  static constexpr bool CUSTOM_EPI = true, ACC_LIMBS3 = true;
  static __device__ __forceinline__ int acc_col(int c) { return c; }
  static __device__ __forceinline__ int num_items(const Args& g) { return (int)((g.rows + 127) / 128); }
  static __device__ __forceinline__ TileCoord coord(const Args&, int item) { return row_tile(item); }
  static __device__ __forceinline__ int num_stages(const Args&, const TileCoord&) { return 1; }
  static __device__ __forceinline__ void load_resident(const Args& g, uint8_t* res, int ptid, int nthr) {
    copy_image<RES_BYTES>(res, g.w_img, ptid, nthr);
  }
  // (arl_prepare_weights) params = the flat parameter buffer; all nthr threads of the CTA take part
  static __device__ __forceinline__ void build_image(const float* params, uint8_t* res, int ptid, int nthr) {
    struct { const float* params; } g{params};
    // (1) s = max |w1| over the tensor, by all threads (named barrier 1)
    uint32_t* smax = reinterpret_cast<uint32_t*>(res + SCALE_OFF + 12);
    if (ptid == 0) *smax = 0u;
    asm volatile("bar.sync 1, %0;" ::"r"(nthr) : "memory");
    float m = 0.f;
    for (int i = ptid; i < 4096; i += nthr) m = fmaxf(m, fabsf(g.params[i]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((ptid & 31) == 0) atomicMax(smax, __float_as_uint(m));      // non-negative floats order as uints
    asm volatile("bar.sync 1, %0;" ::"r"(nthr) : "memory");
    const float s = fmaxf(__uint_as_float(*smax), 1e-30f), inv = 1.0f / s;
    if (ptid == 0) {
      float* sc = reinterpret_cast<float*>(res + SCALE_OFF);
      sc[0] = s * (1.0f / 64.0f); sc[1] = s * (1.0f / 8192.0f); sc[2] = s * (1.0f / 1048576.0f);
    }
    // (2) limbs: one thread per (k-chunk plane, co): 16 K-values = block bytes e = i*4 + j
    for (int ch = ptid; ch < 16 * 16; ch += nthr) {
      const int co = ch & 15, kc = ch >> 4, tap = kc >> 2, c = kc & 3, a = tap >> 1, b = tap & 1;
      uint32_t l0[4] = {0, 0, 0, 0}, l1[4] = {0, 0, 0, 0}, l2[4] = {0, 0, 0, 0};
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        const int i = e >> 2, j = e & 3;
        const float v = g.params[(((4 * a + i) * 8 + 4 * b + j) * 4 + c) * 16 + co] * inv * 64.0f;
        const float r0 = rintf(v), f1 = (v - r0) * 128.0f, r1 = rintf(f1), r2 = rintf((f1 - r1) * 128.0f);
        l0[e >> 2] |= ((uint32_t)(int)r0 & 0xFFu) << (8 * (e & 3));
        l1[e >> 2] |= ((uint32_t)(int)r1 & 0xFFu) << (8 * (e & 3));
        l2[e >> 2] |= ((uint32_t)(int)r2 & 0xFFu) << (8 * (e & 3));
      }
      uint8_t* d = res + kc * PLB + co * 16;
      *reinterpret_cast<uint4*>(d) = make_uint4(l0[0], l0[1], l0[2], l0[3]);
      *reinterpret_cast<uint4*>(d + 16 * 16) = make_uint4(l1[0], l1[1], l1[2], l1[3]);
      *reinterpret_cast<uint4*>(d + 32 * 16) = make_uint4(l2[0], l2[1], l2[2], l2[3]);
    }
    if (ptid < 16) reinterpret_cast<float*>(res + BIAS_OFF)[ptid] = g.params[4096 + ptid];
  }
  // rows of the window that belong to no sample (only in the last tile) are zero-filled here
  static __device__ __forceinline__ void load_stage(const Args& g, const TileCoord& t, int,
                                                    uint8_t* st, int glane, int gsize, Prod&) {
    const int64_t valid = (int64_t)g.num_samples * GROWS - (int64_t)t.mt * 128;
    if (valid >= TROWS) return;
    for (int c = glane; c < TROWS * 4; c += gsize) {
      const int r = c % TROWS, kc = c / TROWS;
      if (r >= valid) *reinterpret_cast<uint4*>(st + kc * PL + r * 16) = make_uint4(0u, 0u, 0u, 0u);
    }
  }
  // lanes 0..7 own one copy each: (segment, plane) = (lane / 4, lane % 4).  At most two samples
  // intersect the window: rows [q0, 441) of n0, then rows [0, ..) of n0+1.  (One lane issuing all
  // eight copies kept the stage's turn-around -- free -> requested -- at ~1.3 us of serial address
  // arithmetic; with 0.65 us per tile and 8 stages that was on the critical path.)
  static __device__ __forceinline__ bool bulk_stage(const Args& g, const TileCoord& t, int, uint8_t* st,
                                                    int glane, int, uint64_t* full) {
    if (glane >= 8) return false;
    const int sgm = glane >> 2, c = glane & 3;
    const int xr0 = t.mt * 128;
    const int n0 = xr0 / GROWS, q0 = xr0 - n0 * GROWS;
    const int first = min(GROWS - q0, TROWS);
    const int n = n0 + sgm;
    int cnt = sgm == 0 ? first : TROWS - first;
    if (n >= g.num_samples) cnt = 0;
    mbar_expect_tx(full, (uint32_t)cnt * 16u);
    if (cnt > 0) {
      const int tt = n / g.geo.num_envs, b = n - tt * g.geo.num_envs;
      const int slot = (g.geo.first_slot + tt + c) % g.geo.ring_slots;
      const uint8_t* src = g.geo.ring + ((size_t)b * g.geo.ring_slots + slot) * kPlane + (sgm == 0 ? q0 : 0) * 16;
      bulk_g2s(st + c * PL + (sgm == 0 ? 0 : first) * 16, src, (uint32_t)cnt * 16u, full);
    }
    return true;
  }
  static __device__ __forceinline__ void issue(const Args&, const TileCoord&, int, uint32_t st,
                                               uint32_t res, uint32_t d) {
    constexpr uint32_t idesc = tc::make_idesc_i8(48);              // x(u8) . [L0 | L1 | L2](s8)
    const uint64_t da0 = tc::make_sdesc(st, PL), db0 = tc::make_sdesc(res, PLB);   // one derivation per stage
#pragma unroll
    for (int tap = 0; tap < 4; ++tap) {
#pragma unroll
      for (int k32 = 0; k32 < 2; ++k32) {
        const uint64_t da = tc::sdesc_advance(da0, ((tap >> 1) * GW + (tap & 1)) * 16 + 2 * k32 * PL);
        const uint64_t db = tc::sdesc_advance(db0, (tap * 4 + 2 * k32) * PLB);
        tc::umma_i8(d, da, db, idesc, (tap | k32) != 0 ? 1u : 0u);
      }
    }
  }
  // epilogue: lane = output pixel; limbs -> fp32 -> /255 + bias, relu -> fp16 -> two 16-B vectors
  // straight into the a1s planes (no staging: 16 consecutive lanes write 256 contiguous B)
  static __device__ __forceinline__ void custom_epilogue(const Args& g, const TileCoord& t,
                                                         const uint8_t* res, uint32_t taddr, int row,
                                                         EpiPre&, EpiState&) {
    const float* sc = reinterpret_cast<const float*>(res + SCALE_OFF);
    float v[16];
    {
      float a[8], b[8];
      tc::tmem_ld8_limbs3(taddr, taddr + 16, taddr + 32, sc[0], sc[1], sc[2], a);
      tc::tmem_ld8_limbs3(taddr + 8, taddr + 24, taddr + 40, sc[0], sc[1], sc[2], b);
#pragma unroll
      for (int e = 0; e < 8; ++e) { v[e] = a[e]; v[8 + e] = b[e]; }
    }
    const int xr = t.mt * 128 + row, n = xr / GROWS;
    const int q = xr - n * GROWS, y = q / GW, x = q - y * GW;
    if (n >= g.num_samples || y >= 20 || x >= 20) return;
    const float2* bias = reinterpret_cast<const float2*>(res + BIAS_OFF);
    uint32_t h[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float2 bb = bias[e];
      const float o0 = fmaxf(fmaf(v[2 * e], 1.0f / 255.0f, bb.x), 0.f);
      const float o1 = fmaxf(fmaf(v[2 * e + 1], 1.0f / 255.0f, bb.y), 0.f);
      h[e] = tc::pack_h2(o0, o1);
    }
    uint8_t* d = g.a1s + (size_t)n * kA1sSample + a1s_offset(y, x, 0);
    *reinterpret_cast<uint4*>(d) = make_uint4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<uint4*>(d + kA1sPlane) = make_uint4(h[4], h[5], h[6], h[7]);
  }
};

// =================================== conv2 forward ============================================
struct Conv2FwdArgs {
  const uint8_t* w_img;  // prepared: fp16 [hi | lo] image of l2_w | l2_b
  const uint8_t* a1s;    // fp16 blocked conv1 output
  uint8_t* a2s;          // split-bf16 chunked output, one block of num_samples rows (gemm_tc.cuh SplitMat)
  int64_t rows;          // 100 * num_samples
  int num_samples;
  int reverse;           // walk the row tiles from the last to the first (see serpentine())
};
struct Conv2Fwd : tc::PolicyBase {
  using Args = Conv2FwdArgs;
  static constexpr int GW = 10, GROWS = 100, TROWS = 140;
  static constexpr int PL = (TROWS + 1) * 16, IMG = 8 * PL;
  // operands arrive by cp.async.bulk straight from the a1s planes: one producer warp per stage,
  // lanes 0..7 own one plane each; two epilogue sets.  One fp16 image per stage (18 KB): six
  // stages = six tiles' worth of copies in flight per SM.
  static constexpr int EPI_SETS = 2, PROD_WARPS = 6, STAGES = 6, STAGE_BYTES = IMG;
  // resident W2 image: rows = [32 co hi | 32 co lo] (N = 64), 32 k-chunk planes
  static constexpr int PLB = 65 * 16, B_IMG = 32 * PLB, BIAS_OFF = B_IMG, RES_BYTES = B_IMG + 128;
  static constexpr int ACC_COLS = 64, OUT_COLS = 32, LO_DELTA = 32, SEG = 32;
  static constexpr bool CUSTOM_EPI = true;
  static __device__ __forceinline__ int acc_col(int c) { return c; }
  static __device__ __forceinline__ int num_items(const Args& g) { return (int)((g.rows + 127) / 128); }
  static __device__ __forceinline__ TileCoord coord(const Args& g, int item) {
    return row_tile(g.reverse ? num_items(g) - 1 - item : item);
  }
  static __device__ __forceinline__ int num_stages(const Args&, const TileCoord&) { return 1; }
  static __device__ __forceinline__ void load_resident(const Args& g, uint8_t* res, int ptid, int nthr) {
    copy_image<RES_BYTES>(res, g.w_img, ptid, nthr);
  }
  static __device__ __forceinline__ void build_image(const float* params, uint8_t* res, int ptid, int nthr) {
    const float* w2 = params + 4112;
    if (ptid < 32) reinterpret_cast<float*>(res + BIAS_OFF)[ptid] = params[4112 + 8192 + ptid];
    for (int ch = ptid; ch < 32 * 32; ch += nthr) {
      const int co = ch & 31, kc = ch >> 5;                    // kc = tap*8 + (i*2+j)*2 + chalf
      const int tap = kc >> 3, ij = (kc >> 1) & 3, chalf = kc & 1;
      const int kh = 2 * (tap >> 1) + (ij >> 1), kw = 2 * (tap & 1) + (ij & 1);
      float x[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) x[e] = w2[((kh * 4 + kw) * 16 + chalf * 8 + e) * 32 + co];
      tc::store_chunk_split_h(res, res + 32 * 16, kc * PLB + co * 16, x);
    }
  }
  static __device__ __forceinline__ void load_stage(const Args& g, const TileCoord& t, int,
                                                    uint8_t* st, int glane, int gsize, Prod&) {
    a1s_zero_tail<TROWS, PL, 8>(st, t.mt * 128, g.num_samples, glane, gsize);
  }
  static __device__ __forceinline__ bool bulk_stage(const Args& g, const TileCoord& t, int, uint8_t* st,
                                                    int glane, int, uint64_t* full) {
    if (glane >= 8) return false;
    mbar_expect_tx(full, a1s_bulk_bytes<TROWS, false>(t.mt * 128, g.num_samples));
    a1s_bulk_issue<TROWS, PL, false>(st, g.a1s, t.mt * 128, g.num_samples, glane, full);
    return true;
  }
  static __device__ __forceinline__ void issue(const Args&, const TileCoord&, int, uint32_t st,
                                               uint32_t res, uint32_t d) {
    constexpr uint32_t idesc64 = tc::make_idesc_h(64);
    const uint64_t da0 = tc::make_sdesc(st, PL), db0 = tc::make_sdesc(res, PLB);   // one derivation per stage
#pragma unroll
    for (int tap = 0; tap < 4; ++tap) {
      const uint32_t aoff = ((tap >> 1) * GW + (tap & 1)) * 16;
#pragma unroll
      for (int k16 = 0; k16 < 4; ++k16) {
        const uint64_t da = tc::sdesc_advance(da0, aoff + 2 * k16 * PL);
        const uint64_t db = tc::sdesc_advance(db0, (tap * 8 + 2 * k16) * PLB);
        tc::umma_f16(d, da, db, idesc64, (tap | k16) != 0 ? 1u : 0u);      // x . [w_hi | w_lo]
      }
    }
  }
  // epilogue: lane = output pixel (n, yp, xp); its 32 channels = features (yp*9+xp)*32 .. +31 of the
  // NHWC flatten (agent.py:231-232) = chunks (yp*9+xp)*4 .. +3 of the split-bf16 a2 block
  // [part][324 chunks][num_samples][8] that the fc256 kernels consume with cp.async.bulk
  static __device__ __forceinline__ void custom_epilogue(const Args& g, const TileCoord& t, const uint8_t* res,
                                                         uint32_t taddr, int row, EpiPre&, EpiState&) {
    const int xr = t.mt * 128 + row, n = xr / GROWS;
    const int q = xr - n * GROWS, yp = q / GW, xp = q - yp * GW;
    const bool ok = n < g.num_samples && yp < 9 && xp < 9;
    const float4* bias = reinterpret_cast<const float4*>(res + BIAS_OFF);
    const int64_t plane = (int64_t)g.num_samples * 16, part = 324 * plane;
    uint8_t* d = g.a2s + (int64_t)((yp * 9 + xp) * 4) * plane + (int64_t)n * 16;
#pragma unroll
    for (int c8 = 0; c8 < 4; ++c8) {
      float v[8];
      tc::tmem_ld8_sum(taddr + c8 * 8, taddr + 32 + c8 * 8, v);          // x.w_hi + x.w_lo
      if (!ok) continue;
      const float4 b0 = bias[c8 * 2], b1 = bias[c8 * 2 + 1];
      uint4 h, l;
      tc::split2(fmaxf(v[0] + b0.x, 0.f), fmaxf(v[1] + b0.y, 0.f), h.x, l.x);
      tc::split2(fmaxf(v[2] + b0.z, 0.f), fmaxf(v[3] + b0.w, 0.f), h.y, l.y);
      tc::split2(fmaxf(v[4] + b1.x, 0.f), fmaxf(v[5] + b1.y, 0.f), h.z, l.z);
      tc::split2(fmaxf(v[6] + b1.z, 0.f), fmaxf(v[7] + b1.w, 0.f), h.w, l.w);
      *reinterpret_cast<uint4*>(d + c8 * plane) = h;
      *reinterpret_cast<uint4*>(d + part + c8 * plane) = l;
    }
  }
};

// =================================== conv2 input gradient =====================================
struct Conv2DgradArgs {
  const uint8_t* w_img;  // prepared: transposed bf16 [hi | lo] image of l2_w
  const uint8_t* a1s;    // relu mask of conv1: sign of its fp16 output
  const float* dy2;      // [N, 81, 32]
  uint8_t* dy1s;         // conv1 output gradient, one fp16 per value on the conv1 X grid ("dy1s", see below)
  float* bias_partials;  // [grid * 8 epilogue warps][16]: column sums of dy1 (= the conv1 bias gradient)
  int64_t rows;          // 121 * num_samples
  int num_samples;
};
// dy1s: the gradient w.r.t. conv1's output is only ever the B operand of the conv1 weight-gradient
// MMA (rows of the 21-wide X grid = the reduction index), so conv2 dgrad stores it the way that
// kernel fetches it with cp.async.bulk -- ONE fp16 per value (x tensor_scale, saturating):
//   dy1s[co group (2)][grid row n*441 + y*21 + x][8 channels]      (16-B vectors)
// with the rows y = 20 / x = 20 (no conv1 output there) written as zeros: 14 112 B per sample.
// One term is enough HERE (unlike dy2): its only consumer is l1_w (l1_b comes from the fp32 sums
// of this kernel's epilogue), 2e-4 in tests/precision_study.py's "fp16 a1 + d_a1" row; round 1
// kept a bf16 hi + lo pair (28 224 B).
constexpr int kDy1sRows = 441;
struct Conv2Dgrad : tc::PolicyBase {
  using Args = Conv2DgradArgs;
  static constexpr int GW = 11, GROWS = 121, TROWS = 140;
  static constexpr int PL = 146 * 16, IMG = 4 * PL;           // 32 co = 4 chunks
  // the relu-mask epilogue waits on HBM: two epilogue sets; 8 producer warps (17 warps in all ->
  // 96 registers per thread, 8 float4 mask loads in flight per epilogue lane)
  static constexpr int EPI_SETS = 2, PROD_WARPS = 8, STAGES = 4, STAGE_BYTES = 2 * IMG;
  // resident W2^T image: rows = part*64 + cls*16 + c (N = 128: the 4 parity classes share the A
  // tile of a tap, and lo follows hi), 16 k-chunk planes (tap*4 + co8)
  static constexpr int PLB = 129 * 16, RES_BYTES = 16 * PLB;
  static constexpr int ACC_COLS = 128, OUT_COLS = 64, LO_DELTA = 64, SEG = 32;   // 4 classes x 16 ch
  static constexpr bool CUSTOM_EPI = true;
  static __device__ __forceinline__ int acc_col(int c) { return c; }
  static __device__ __forceinline__ int num_items(const Args& g) { return (int)((g.rows + 127) / 128); }
  static __device__ __forceinline__ TileCoord coord(const Args&, int item) { return row_tile(item); }
  static __device__ __forceinline__ int num_stages(const Args&, const TileCoord&) { return 1; }
  static __device__ __forceinline__ void load_resident(const Args& g, uint8_t* res, int ptid, int nthr) {
    copy_image<RES_BYTES>(res, g.w_img, ptid, nthr);
  }
  static __device__ __forceinline__ void build_image(const float* params, uint8_t* res, int ptid, int nthr) {
    const float* w2 = params + 4112;
    for (int ch = ptid; ch < 4 * 16 * 16; ch += nthr) {
      const int c = ch & 15, kc = (ch >> 4) & 15, cls = ch >> 8;   // kc = tap*4 + co8
      const int tap = kc >> 2, co8 = kc & 3;
      const int kh = (cls >> 1) + 2 * (tap >> 1), kw = (cls & 1) + 2 * (tap & 1);
      float x[8];
      load8(w2 + ((kh * 4 + kw) * 16 + c) * 32 + co8 * 8, x);
      tc::store_chunk_split(res, res + 64 * 16, kc * PLB + (cls * 16 + c) * 16, x);
    }
  }
  static __device__ __forceinline__ void load_stage(const Args& g, const TileCoord& t, int,
                                                    uint8_t* st, int glane, int gsize, Prod&) {
    float unused[4];
    const int z0 = t.mt * 128;
    stream_pixels<GridZ, 8, TROWS, PL, 7, false, false>(st, st + IMG, g.dy2, z0, z0 + TROWS, g.num_samples,
                                                         glane, gsize, unused);
  }
  static __device__ __forceinline__ void issue(const Args&, const TileCoord&, int, uint32_t st,
                                               uint32_t res, uint32_t d) {
    constexpr uint32_t idesc128 = tc::make_idesc(128), idesc64 = tc::make_idesc(64);
#pragma unroll
    const uint64_t da0 = tc::make_sdesc(st, PL), db0 = tc::make_sdesc(res, PLB);   // one derivation per stage
#pragma unroll
    for (int tap = 0; tap < 4; ++tap) {
      // output (yy,xx) of every class reads Z row P + 12 - 11*dkh - dkw
      const uint32_t aoff = (12 - 11 * (tap >> 1) - (tap & 1)) * 16;
#pragma unroll
      for (int k16 = 0; k16 < 2; ++k16) {
        const uint64_t da_hi = tc::sdesc_advance(da0, aoff + 2 * k16 * PL);
        const uint64_t da_lo = tc::sdesc_advance(da0, aoff + IMG + 2 * k16 * PL);
        const uint64_t db = tc::sdesc_advance(db0, (tap * 4 + 2 * k16) * PLB);
        tc::umma_f16(d, da_hi, db, idesc128, (tap | k16) != 0 ? 1u : 0u);  // z_hi . [wT_hi | wT_lo]
        tc::umma_f16(d, da_lo, db, idesc64, 1u);                            // z_lo . wT_hi
      }
    }
  }
  // epilogue: lane = row (yy, xx) of the 11-wide grid = the 2x2 block of dy1 pixels (2yy+dy, 2xx+dx),
  // accumulator columns cls*16 + c with cls = dy*2 + dx.  Per pixel and channel half: relu mask
  // (sign of a1's hi part, prefetched before the accumulator wait), bias-gradient sums, bf16 split,
  // one hi and one lo vector into dy1s.  Rows yy = 10 / xx = 10 write the zero rows y = 20 / x = 20.
  // Lanes of a warp write 16-B vectors 32 B apart, the dx = 0/1 stores interleave: 8 lines each.
  struct EpiPre { uint4 m[8]; };
  struct EpiState { float b[16]; };
  static __device__ __forceinline__ void epi_begin(EpiState& s) {
#pragma unroll
    for (int c = 0; c < 16; ++c) s.b[c] = 0.f;
  }
  static __device__ __forceinline__ void epi_prefetch(const Args& g, const TileCoord& t, int row, EpiPre& pre) {
    const int pr = t.mt * 128 + row, n = pr / GROWS;
    const int q = pr - n * GROWS, yy = q / GW, xx = q - yy * GW;
    const bool ok = n < g.num_samples && yy < 10 && xx < 10;
    const uint8_t* p = g.a1s + (int64_t)n * kA1sSample + (yy * 10 + xx) * 16;     // plane = cls*2 + chalf (fp16 a1)
#pragma unroll
    for (int i = 0; i < 8; ++i)
      pre.m[i] = ok ? __ldg(reinterpret_cast<const uint4*>(p + i * kA1sPlane)) : make_uint4(0u, 0u, 0u, 0u);
  }
  static __device__ __forceinline__ void custom_epilogue(const Args& g, const TileCoord& t, const uint8_t*,
                                                         uint32_t taddr, int row, EpiPre& pre, EpiState& st) {
    const int pr = t.mt * 128 + row, n = pr / GROWS;
    const int q = pr - n * GROWS, yy = q / GW, xx = q - yy * GW;
    const bool live = n < g.num_samples;
    const int64_t R = (int64_t)g.num_samples * kDy1sRows;
#pragma unroll
    for (int cls = 0; cls < 4; ++cls) {
      const int y = 2 * yy + (cls >> 1), x = 2 * xx + (cls & 1);
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        float v[8];
        tc::tmem_ld8_sum(taddr + cls * 16 + half * 8, taddr + 64 + cls * 16 + half * 8, v);
        if (!live || y > 20 || x > 20) continue;
        uint4 hv = make_uint4(0u, 0u, 0u, 0u);
        if (y < 20 && x < 20) {
          const uint4 m = pre.m[cls * 2 + half];
          const uint32_t w[4] = {m.x, m.y, m.z, m.w};
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            if ((int16_t)(w[i] & 0xFFFFu) <= 0) v[2 * i] = 0.f;        // fp16 > 0 <=> its bits as int16 > 0
            if ((int32_t)w[i] < 0x10000) v[2 * i + 1] = 0.f;
          }
#pragma unroll
          for (int e = 0; e < 8; ++e) st.b[half * 8 + e] += v[e];
          hv = make_uint4(tc::pack_h2(v[0], v[1]), tc::pack_h2(v[2], v[3]), tc::pack_h2(v[4], v[5]),
                          tc::pack_h2(v[6], v[7]));
        }
        *reinterpret_cast<uint4*>(g.dy1s + ((int64_t)half * R + (int64_t)n * kDy1sRows + y * 21 + x) * 16) = hv;
      }
    }
  }
  static __device__ __forceinline__ void epi_end(const Args& g, EpiState& st, int warp, int lane) {
#pragma unroll
    for (int c = 0; c < 16; ++c) {
      const float s = warp_sum(st.b[c]);
      if (lane == 0) g.bias_partials[((size_t)blockIdx.x * 8 + warp) * 16 + c] = s;
    }
  }
};

// =================================== weight gradients =========================================
// work item `item` (= its partial slice `ks`) covers the row range number `range`
__device__ __forceinline__ TileCoord range_tile(int item, int64_t rows, int k_chunk, int range) {
  TileCoord t;
  t.mt = item; t.nt = 0; t.ks = item;
  const int64_t b = (int64_t)range * k_chunk;
  t.k_begin = (int)b;
  t.k_end = (int)(b + k_chunk < rows ? b + k_chunk : rows);
  return t;
}
// per-lane bias sums (4 floats of every CF4-th float4) -> per-warp partial row of CF4*4 floats
template <int CF4>
__device__ __forceinline__ void bias_partial_store(float* dst_row, float (&acc)[4], int lane) {
#pragma unroll
  for (int e = 0; e < 4; ++e) {
#pragma unroll
    for (int o = 16; o >= CF4; o >>= 1) acc[e] += __shfl_xor_sync(0xffffffffu, acc[e], o);
  }
  if (lane < CF4)
    *reinterpret_cast<float4*>(dst_row + lane * 4) = make_float4(acc[0], acc[1], acc[2], acc[3]);
}

struct Conv2WgradArgs {
  const uint8_t* a1s;    // fp16 blocked conv1 output
  const float* dy2;
  float* partials;       // [items][8192]
  float* bias_partials;  // [items * PROD_WARPS][32]  column sums of dy2 (= db2), per producer warp
  int64_t rows;          // 100 * num_samples
  int num_samples, k_chunk, items;
  int reverse;           // item i takes row range items-1-i (see serpentine())
};
struct Conv2Wgrad : tc::PolicyBase {
  using Args = Conv2WgradArgs;
  struct Prod { float acc[4]; };    // running column sums of dy2 (bias gradient)
  // dW2[(a,b) tap][ch][co] = sum_P X2[P + a*10 + b][ch] * dy2[P][co] over the rows P of the X2 grid.
  // Both operands are fp16: X2 is bulk-copied from a1s as it is (one fp16 per value), dy2 is
  // converted by the producers into an fp16 hi + lo pair concatenated along N (the gradient keeps
  // its two terms: a single fp16 term left l1_b / l2_w at 1e-3 against the oracle).  Tap b is
  // folded into M: a second bulk copy of the X2 image shifted by one row fills planes 8..15, so
  // rows b*64 + ch of the accumulator are tap b; tap a is a row offset of the A descriptor and its
  // own block of 64 accumulator columns.  One N = 64 MMA per tap a and 16 rows: 16 per stage
  // (round 1: 32, with X2 as a hi/lo pair stacked along M).
  static constexpr int GW = 10, GROWS = 100, TROWS = 140;
  static constexpr int PLA = (TROWS + 1) * 16, A_IMG = 8 * PLA;    // one copy: 8 channel groups
  static constexpr int PLB = 130 * 16, B_IMG = 4 * PLB;            // dy2: 4 co groups per part
  static constexpr int PROD_WARPS = 18, STAGES = 3, STAGE_BYTES = 2 * A_IMG + 2 * B_IMG, RES_BYTES = 0;
  // accumulator columns: a*64 + part*32 + co
  static constexpr int ACC_COLS = 128, OUT_COLS = 64, LO_DELTA = 32, SEG = 32;
  static __device__ __forceinline__ int acc_col(int c) { return (c >> 5) * 64 + (c & 31); }
  static __device__ __forceinline__ int num_items(const Args& g) { return g.items; }
  static __device__ __forceinline__ TileCoord coord(const Args& g, int item) {
    return range_tile(item, g.rows, g.k_chunk, g.reverse ? g.items - 1 - item : item);
  }
  static __device__ __forceinline__ int num_stages(const Args&, const TileCoord& t) {
    return (t.k_end - t.k_begin + 127) / 128;
  }
  static __device__ __forceinline__ void load_resident(const Args&, uint8_t*, int, int) {}
  static __device__ __forceinline__ void prod_begin(Prod& ps) {
    ps.acc[0] = ps.acc[1] = ps.acc[2] = ps.acc[3] = 0.f;
  }
  static __device__ __forceinline__ void load_stage(const Args& g, const TileCoord& t, int s,
                                                    uint8_t* st, int glane, int gsize, Prod& ps) {
    const int p0 = t.k_begin + s * 128;
    // A: X2 rows p0 .. p0+139 (and the copy shifted by one row) arrive by cp.async.bulk
    // (bulk_stage); only the rows beyond the last sample need zeros here
    a1s_zero_tail<TROWS, PLA, 16>(st, p0, g.num_samples, glane, gsize);
    // B: dy2 on the 10-wide grid, zero at y'=9 / x'=9 and outside [k_begin, k_end)
    uint8_t* b_hi = st + 2 * A_IMG;
    stream_pixels<GridDy2, 8, 128, PLB, 7, true, true>(b_hi, b_hi + B_IMG, g.dy2, p0, t.k_end, g.num_samples,
                                                       glane, gsize, ps.acc);
  }
  // 8 lanes of the stage's 192 (every 24th) own one plane each
  static __device__ __forceinline__ bool bulk_stage(const Args& g, const TileCoord& t, int s, uint8_t* st,
                                                    int glane, int, uint64_t* full) {
    if (glane % 24 != 0) return false;
    const int p0 = t.k_begin + s * 128, kc = glane / 24;             // 0..7
    mbar_expect_tx(full, a1s_bulk_bytes<TROWS, true>(p0, g.num_samples));
    a1s_bulk_issue<TROWS, PLA, true>(st, g.a1s, p0, g.num_samples, kc, full);
    return true;
  }
  static __device__ __forceinline__ void prod_end(const Args& g, const TileCoord& t, Prod& ps, int pw,
                                                  int lane) {
    bias_partial_store<8>(g.bias_partials + ((size_t)t.ks * PROD_WARPS + pw) * 32, ps.acc, lane);
  }
  static __device__ __forceinline__ void issue(const Args&, const TileCoord&, int s, uint32_t st,
                                               uint32_t, uint32_t d) {
    constexpr uint32_t idesc = tc::make_idesc_h(64, true, true);     // [x ; x shifted] . [dy_hi | dy_lo]
    // one descriptor derivation per stage
    const uint64_t da0 = tc::make_sdesc(st, 128, PLA), db0 = tc::make_sdesc(st + 2 * A_IMG, 128, PLB);
#pragma unroll
    for (int a = 0; a < 2; ++a) {
#pragma unroll
      for (int k16 = 0; k16 < 8; ++k16) {
        const uint64_t da = tc::sdesc_advance(da0, (a * GW + k16 * 16) * 16);
        const uint64_t db = tc::sdesc_advance(db0, k16 * 256);
        tc::umma_f16(d + a * 64, da, db, idesc, (s | k16) != 0 ? 1u : 0u);
      }
    }
  }
  // row = b*64 + ch, ch = (i*2+j)*16 + cin ; segment a = 32 co of tap (kh = 2a+i, kw = 2b+j)
  static __device__ __forceinline__ float* row_ptr(const Args& g, const TileCoord& t, int row) {
    const int b = row >> 6, ch = row & 63, ij = ch >> 4, cin = ch & 15;
    return g.partials + (size_t)t.ks * 8192 + (((ij >> 1) * 4 + 2 * b + (ij & 1)) * 16 + cin) * 32;
  }
  static __device__ __forceinline__ int64_t seg_offset(const Args&, const TileCoord&, int a) {
    return a * 8 * 16 * 32;
  }
};

struct Conv1WgradArgs {
  RingGeo geo;
  const uint8_t* dy1s;   // fp16 on the X grid, written by conv2 dgrad (see dy1s above)
  float* partials;       // [items][4096]
  int64_t rows;          // 441 * num_samples
  int num_samples, k_chunk, items;
  int reverse;           // item i takes row range items-1-i (see serpentine())
};
struct Conv1Wgrad : tc::PolicyBase {
  using Args = Conv1WgradArgs;
  // dW1[(a,b) tap][ch][co] = sum_P X[P + a*21 + b][ch] * dy1[P][co] over the rows P of the X grid.
  // Both operands fp16: the u8 pixels are exact, dy1 is one fp16 per value.  Tap b is folded into M
  // (a second copy of the X image shifted by one row), tap a into N (a second copy of the dy1 image
  // shifted BACK by 21 rows: sum_Q X[Q + b][ch] * dy1[Q - 21][co]), so one N = 32 MMA per 16 rows
  // covers all four taps and the wide X tile is read from shared memory once.  The dy1 images
  // arrive by cp.async.bulk from dy1s (4 copies of 2 KB per stage).
  static constexpr int GW = 21, TROWS = 129;
  static constexpr int PLA = (TROWS + 1) * 16, A_IMG = 16 * PLA;   // u8 -> fp16, exact
  static constexpr int PLB = 128 * 16, B_IMG = 2 * PLB;            // dy1: 2 co groups
  // stage = [A | dy | dy']  (dy' = shifted copy = tap a = 1); five stages of 41 KB
  static constexpr int PROD_WARPS = 20, STAGES = 5, STAGE_BYTES = A_IMG + 2 * B_IMG, RES_BYTES = 0;
  // accumulator columns: a*16 + co
  static constexpr int ACC_COLS = 32, OUT_COLS = 32, LO_DELTA = 0, SEG = 16;
  static __device__ __forceinline__ int acc_col(int c) { return c; }
  static __device__ __forceinline__ int num_items(const Args& g) { return g.items; }
  static __device__ __forceinline__ TileCoord coord(const Args& g, int item) {
    return range_tile(item, g.rows, g.k_chunk, g.reverse ? g.items - 1 - item : item);
  }
  static __device__ __forceinline__ int num_stages(const Args&, const TileCoord& t) {
    return (t.k_end - t.k_begin + 127) / 128;
  }
  static __device__ __forceinline__ void load_resident(const Args&, uint8_t*, int, int) {}
  // the u8 ring rows of the NEXT stage are requested (5 x 16 B per lane, held in registers) before
  // the producer waits for its slot: the ncu capture of round 1 showed producers 24 % of their time
  // waiting for a free stage and then another 13 % on the first use of these loads
  static constexpr bool PREFETCH = true;
  static constexpr int XU = 5;                                      // 20 warps / 5 stages = 128 lanes x 5 >= 129 x 4
  struct Prod { uint4 w[XU]; };
  static __device__ __forceinline__ void prefetch_stage(const Args& g, const TileCoord& t, int s, int glane,
                                                        int gsize, Prod& ps) {
    x1_request<TROWS, XU>(ps.w, g.geo, t.k_begin + s * 128, g.num_samples, glane, gsize);
  }
  static __device__ __forceinline__ void load_stage(const Args& g, const TileCoord& t, int s,
                                                    uint8_t* st, int glane, int gsize, Prod& ps) {
    const int p0 = t.k_begin + s * 128;
    // A: X rows p0 .. p0+128 (exact in bf16: one image); planes 8..15 = shifted copy (tap b = 1)
    x1_store<TROWS, PLA, XU, true>(st, ps.w, glane, gsize);
    // B: image row r of planes 0..3 = dy1 row p0 + r, of planes 4..7 = dy1 row p0 + r - 21; rows
    // with p0 + r >= k_end (the next work item's) and rows before the first sample are zero
    const int valid = t.k_end - p0, lead = p0 < GW ? GW - p0 : 0;
    if (valid < 128 || lead > 0) {
      for (int i = glane; i < 4 * 128; i += gsize) {
        const int plane = i >> 7, r = i & 127;
        if (r >= valid || (plane >= 2 && r < lead))
          *reinterpret_cast<uint4*>(st + A_IMG + plane * PLB + r * 16) = make_uint4(0u, 0u, 0u, 0u);
      }
    }
  }
  // lanes 0..3 of the stage's group own one dy1 plane each: plane = copy*2 + co group
  static __device__ __forceinline__ bool bulk_stage(const Args& g, const TileCoord& t, int s, uint8_t* st,
                                                    int glane, int, uint64_t* full) {
    if (glane >= 4) return false;
    const int p0 = t.k_begin + s * 128;
    int cnt = t.k_end - p0 < 128 ? t.k_end - p0 : 128, src = p0, dst = 0;
    if (glane >= 2) {
      src = p0 - GW;
      if (src < 0) { dst = -src; cnt -= dst; src = 0; }
    }
    const uint32_t bytes = cnt > 0 ? (uint32_t)cnt * 16u : 0u;
    mbar_expect_tx(full, bytes);
    if (cnt > 0)
      bulk_g2s(st + A_IMG + glane * PLB + dst * 16, g.dy1s + ((int64_t)(glane & 1) * g.rows + src) * 16, bytes, full);
    return true;
  }
  static __device__ __forceinline__ void issue(const Args&, const TileCoord&, int s, uint32_t st,
                                               uint32_t, uint32_t d) {
    constexpr uint32_t idesc = tc::make_idesc_h(32, true, true);   // [x ; x shifted] . [dy | dy']
    // a_img = st, b_hi = st + A_IMG; one descriptor derivation per stage
    const uint64_t da0 = tc::make_sdesc(st, 128, PLA), db0 = tc::make_sdesc(st + A_IMG, 128, PLB);
#pragma unroll
    for (int k16 = 0; k16 < 8; ++k16)
      tc::umma_f16(d, tc::sdesc_advance(da0, k16 * 256), tc::sdesc_advance(db0, k16 * 256), idesc,
                   (s | k16) != 0 ? 1u : 0u);
  }
  // row = b*64 + ch, ch = (cin*4 + i)*4 + j ; segment a = 16 co of tap (kh = 4a+i, kw = 4b+j)
  static __device__ __forceinline__ float* row_ptr(const Args& g, const TileCoord& t, int row) {
    const int b = row >> 6, ch = row & 63, cin = ch >> 4, i = (ch >> 2) & 3, j = ch & 3;
    const int kw = 4 * b + j;
    return g.partials + (size_t)t.ks * 4096 + ((i * 8 + kw) * 4 + cin) * 16;
  }
  static __device__ __forceinline__ int64_t seg_offset(const Args&, const TileCoord&, int a) {
    return a * 4 * 8 * 4 * 16;
  }
  static __device__ __forceinline__ float4 finish(const Args&, float4 v, float4) {
    const float s = 1.0f / 255.0f;
    return make_float4(v.x * s, v.y * s, v.z * s, v.w * s);
  }
};

// arl_prepare_weights: CTA 0 / 1 / 2 builds the resident image of conv1 fwd / conv2 fwd / conv2 dgrad
// in shared memory (the same code the kernels used to run in their prologues) and writes it out
__global__ void __launch_bounds__(256) build_images_kernel(const float* __restrict__ params, uint8_t* prepared) {
  extern __shared__ __align__(128) uint8_t img[];
  const int tid = threadIdx.x;
  int bytes = 0;
  uint8_t* out = prepared;
  if (blockIdx.x == 0) {
    Conv1Fwd::build_image(params, img, tid, 256);
    bytes = Conv1Fwd::RES_BYTES; out += kPrepW1;
  } else if (blockIdx.x == 1) {
    Conv2Fwd::build_image(params, img, tid, 256);
    bytes = Conv2Fwd::RES_BYTES; out += kPrepW2F;
  } else {
    Conv2Dgrad::build_image(params, img, tid, 256);
    bytes = Conv2Dgrad::RES_BYTES; out += kPrepW2D;
  }
  __syncthreads();
  for (int i = tid; i < bytes / 16; i += 256) reinterpret_cast<uint4*>(out)[i] = reinterpret_cast<const uint4*>(img)[i];
}
static_assert(Conv1Fwd::RES_BYTES <= kPrepW1Bytes && Conv2Fwd::RES_BYTES <= kPrepW2FBytes &&
                  Conv2Dgrad::RES_BYTES <= kPrepW2DBytes, "prepared-weight layout");

// Serpentine sweep: consecutive kernels of a chain walk the samples in OPPOSITE directions, so a
// kernel starts with the rows its predecessor touched last -- the ones still in the 126 MB L2 --
// instead of the ones evicted longest ago.  forward: conv1 up, conv2 DOWN; backward: fc dgrad up,
// conv2 wgrad DOWN, conv2 dgrad up, conv1 wgrad DOWN.  ARL_SERPENTINE=0 turns it off (A/B runs).
bool serpentine() {
  static const bool on = [] { const char* e = getenv("ARL_SERPENTINE"); return !(e && e[0] == '0'); }();
  return on;
}

int split_rows(int64_t rows, int want, int* k_chunk) {
  int64_t per = (rows + want - 1) / want;
  per = (per + 127) / 128 * 128;
  *k_chunk = (int)per;
  return (int)((rows + per - 1) / per);
}

}  // namespace

int conv_prepare(const float* params, void* prepared, cudaStream_t st) {
  build_images_kernel<<<3, 256, 33536, st>>>(params, (uint8_t*)prepared);
  ARL_LAUNCH_CHECK("build_images_kernel");
  return ARL_OK;
}

}  // namespace arl

using namespace arl;

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

extern "C" int arl_conv1_forward(const float* prepared, const uint8_t* ring, float* a1, int num_envs,
                                 int ring_slots, int first_slot, int steps, void* stream) {
  const float* params = prepared;
  ARL_REQUIRE(params && ring && a1, "arl_conv1_forward: null pointer");
  ARL_REQUIRE(num_envs >= 0 && steps >= 0, "arl_conv1_forward: negative size");
  ARL_REQUIRE(ring_slots >= steps + 3 && first_slot >= 0 && first_slot < ring_slots,
              "arl_conv1_forward: ring_slots %d must be >= steps+3 (%d) and first_slot %d inside it",
              ring_slots, steps + 3, first_slot);
  ARL_REQUIRE(aligned16(params) && aligned16(ring) && aligned16(a1),
              "arl_conv1_forward: pointers must be 16-byte aligned");
  const int64_t N = (int64_t)num_envs * steps;
  if (N == 0) return ARL_OK;
  ARL_REQUIRE(N * 441 < (1LL << 31) - 256, "arl_conv1_forward: too many samples");
  Conv1FwdArgs g{reinterpret_cast<const uint8_t*>(prepared) + kPrepW1, {ring, num_envs, ring_slots, first_slot},
                 reinterpret_cast<uint8_t*>(a1), N * 441, (int)N};
  return tc::launch<Conv1Fwd>(g, (int)((g.rows + 127) / 128), (cudaStream_t)stream);
}

extern "C" int arl_conv2_forward(const float* prepared, const float* a1, float* a2,
                                 int64_t num_samples, void* stream) {
  const float* params = prepared;
  ARL_REQUIRE(params && a1 && a2, "arl_conv2_forward: null pointer");
  ARL_REQUIRE(num_samples >= 0 && num_samples < (1LL << 24), "arl_conv2_forward: bad num_samples");
  ARL_REQUIRE(aligned16(params) && aligned16(a1) && aligned16(a2),
              "arl_conv2_forward: pointers must be 16-byte aligned");
  if (num_samples == 0) return ARL_OK;
  Conv2FwdArgs g{reinterpret_cast<const uint8_t*>(prepared) + kPrepW2F, reinterpret_cast<const uint8_t*>(a1),
                 reinterpret_cast<uint8_t*>(a2), num_samples * 100, (int)num_samples, serpentine() ? 1 : 0};
  return tc::launch<Conv2Fwd>(g, (int)((g.rows + 127) / 128), (cudaStream_t)stream);
}

extern "C" int arl_conv1_backward(const uint8_t* ring, const float* d_a1, float* grads,
                                  void* workspace, int num_envs, int ring_slots, int first_slot,
                                  int steps, float grad_unscale, void* stream) {
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
    ARL_CUDA(cudaMemsetAsync(grads, 0, 4096 * sizeof(float), st));
    return ARL_OK;
  }
  ARL_REQUIRE(N * 441 < (1LL << 31) - 256, "arl_conv1_backward: too many samples");
  Conv1WgradArgs g;
  g.geo = {ring, num_envs, ring_slots, first_slot};
  g.dy1s = reinterpret_cast<const uint8_t*>(d_a1);
  g.partials = (float*)workspace;
  g.rows = N * 441;
  g.num_samples = (int)N;
  g.items = split_rows(g.rows, num_sms(), &g.k_chunk);
  g.reverse = serpentine() ? 1 : 0;
  int rc = tc::launch<Conv1Wgrad>(g, g.items, st);
  if (rc) return rc;
  return reduce_partials_scaled(g.partials, grads, g.items, 4096, grad_unscale, st);   // l1_w
}

extern "C" int arl_conv2_backward(const float* prepared, const float* a1, const float* d_a2,
                                  float* d_a1, float* grads, void* workspace, int64_t num_samples,
                                  float grad_unscale, void* stream) {
  const float* params = prepared;
  ARL_REQUIRE(params && a1 && d_a2 && d_a1 && grads && workspace,
              "arl_conv2_backward: null pointer");
  ARL_REQUIRE(num_samples >= 0 && num_samples < (1LL << 24), "arl_conv2_backward: bad num_samples");
  ARL_REQUIRE(aligned16(params) && aligned16(a1) && aligned16(d_a2) && aligned16(d_a1) &&
                  aligned16(grads) && aligned16(workspace),
              "arl_conv2_backward: pointers must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  float* g2 = grads + 4096 + 16;                                                 // l2_w | l2_b
  if (num_samples == 0) {
    ARL_CUDA(cudaMemsetAsync(grads + 4096, 0, (16 + 8192 + 32) * sizeof(float), st));   // l1_b | l2_w | l2_b
    return ARL_OK;
  }
  Conv2WgradArgs w;
  w.a1s = reinterpret_cast<const uint8_t*>(a1);
  w.dy2 = d_a2;
  w.partials = (float*)workspace;
  w.rows = num_samples * 100;
  w.num_samples = (int)num_samples;
  w.items = split_rows(w.rows, num_sms(), &w.k_chunk);
  w.reverse = serpentine() ? 1 : 0;
  w.bias_partials = (float*)workspace + (size_t)w.items * 8192;
  int rc = tc::launch<Conv2Wgrad>(w, w.items, st);
  if (rc) return rc;
  rc = reduce_partials_scaled(w.partials, g2, w.items, 8192, grad_unscale, st);
  if (rc) return rc;
  // l2_b = column sums of d_a2, accumulated by the wgrad producers while they stream d_a2
  rc = reduce_partials_scaled(w.bias_partials, g2 + 8192, w.items * Conv2Wgrad::PROD_WARPS, 32, grad_unscale, st);
  if (rc) return rc;
  // dgrad; its epilogue also sums the columns of d_a1 = the conv1 bias gradient l1_b
  float* db1 = (float*)workspace + (size_t)w.items * (8192 + Conv2Wgrad::PROD_WARPS * 32);
  Conv2DgradArgs d{reinterpret_cast<const uint8_t*>(prepared) + kPrepW2D, reinterpret_cast<const uint8_t*>(a1),
                   d_a2, reinterpret_cast<uint8_t*>(d_a1), db1, num_samples * 121, (int)num_samples};
  const int items = (int)((d.rows + 127) / 128);
  rc = tc::launch<Conv2Dgrad>(d, items, st);
  if (rc) return rc;
  return reduce_partials_scaled(db1, grads + 4096, (items < num_sms() ? items : num_sms()) * 8, 16, grad_unscale, st);
}
