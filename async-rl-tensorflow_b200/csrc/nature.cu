// The 'nature' trunk (reference src/network.py:30-42): conv32 8x8 s4 -> conv64 4x4 s2 -> conv64
// 3x3 s1 -> fc512 -> policy / value heads, forward and backward, as a second shape set.
//
// Unlike the 'nips' kernels (convs_tc.cu / fc.cu: one hand-laid-out policy per contraction, operands
// kept in HBM as split-bf16 images) this path is ONE generic tcgen05 contraction,
//     D[m][n] = alpha * sum_k A(m,k) * B(k,n)        (+ bias, relu | relu mask of another tensor)
// whose operands are GATHERED by the producer warps from plain float32 NHWC tensors through affine
// index maps: every row index m and reduction index k decomposes into three digits, each digit
// contributes to an element offset and to a (y, x) position, and the element exists where the
// summed position is inside [0,H) x [0,W).  That one rule covers im2col (conv forward), its
// transpose (conv weight gradient), the strided transposed convolution by output-parity class
// (conv input gradient: zero where the tap falls outside dY) and plain / transposed matrices
// (fc512).  Values enter the tensor cores as bf16 hi + lo (three products, fp32 accumulate) like
// everywhere else in the library.  It is the correct-first shape set of SURVEY §8 f4: parity to the
// same 1e-3 bar, not laid out for the HBM roofline the way the nips kernels are.
#include "gemm_tc.cuh"

namespace arl {

int reduce_partials(const float* partials, float* out, int num_partials, int n, cudaStream_t stream);

namespace {
using tc::TileCoord;
using tc::kTileM;

// index i -> digits (i / (d1*d2), (i / d2) % d1, i % d2) -> element offset and (y, x) contribution
struct Lin3 {
  int d1, d2;
  long long so0, so1, so2, base;
  int sy0, sy1, sy2, sx0, sx1, sx2;
};
struct Pos { long long off; int y, x; };
__host__ __device__ __forceinline__ Pos lin3(const Lin3& L, int i) {
  const int q2 = i % L.d2, t = i / L.d2, q1 = t % L.d1, q0 = t / L.d1;
  Pos p;
  p.off = L.base + q0 * L.so0 + q1 * L.so1 + q2 * L.so2;
  p.y = q0 * L.sy0 + q1 * L.sy1 + q2 * L.sy2;
  p.x = q0 * L.sx0 + q1 * L.sx1 + q2 * L.sx2;
  return p;
}
constexpr int kBig = 1 << 30;
inline Lin3 flat(long long stride, long long base = 0) {     // i -> base + i * stride, no position
  Lin3 L = {1, kBig, 0, 0, stride, base, 0, 0, 0, 0, 0, 0};
  return L;
}

enum { GG_PLAIN = 0, GG_BIAS = 1, GG_MASK = 2 };
struct GatherGemmArgs {
  const float* A; Lin3 a_row, a_k; int H, W;      // A(m,k) = A[a_row(m).off + a_k(k).off] where the summed (y,x) is in [0,H)x[0,W)
  const float* B; Lin3 b_k, b_n;                  // B(k,n) = B[b_k(k).off + b_n(n).off]
  float* D; Lin3 d_row; long long d_slice;        // D(m,n) at D[ks*d_slice + d_row(m).off + n]
  const float* bias;                              // GG_BIAS: bias[n]
  const float* mask;                              // GG_MASK: same offsets as D; output kept where mask > 0
  float alpha; int relu;
  int M, N, K, k_chunk, k_splits, m_tiles, n_tiles;
};

template <int EPI>
struct GatherGemm : tc::PolicyBase {
  using Args = GatherGemmArgs;
  static constexpr int NT = 64, KB = 32, STAGES = 4, PROD_WARPS = 16;
  static constexpr int ACC_COLS = 2 * NT, OUT_COLS = NT, LO_DELTA = NT, SEG = 32;
  static constexpr bool HAS_AUX = EPI != GG_PLAIN, AUX_ROW_INVARIANT = EPI == GG_BIAS;
  static __device__ __forceinline__ int acc_col(int c) { return c; }
  // stage = [A hi | A lo | B^T (rows: hi NT | lo NT)], all K-major: vector (row, k chunk) at chunk*PLANE + row*16
  static constexpr int A_PLANE = kTileM * 16, A_PART = (KB / 8) * A_PLANE;
  static constexpr int B_PLANE = 2 * NT * 16, B_OFF = 2 * A_PART;
  static constexpr int STAGE_BYTES = B_OFF + (KB / 8) * B_PLANE, RES_BYTES = 0;

  static __device__ __forceinline__ int num_items(const Args& g) { return g.m_tiles * g.n_tiles * g.k_splits; }
  static __device__ __forceinline__ TileCoord coord(const Args& g, int item) {
    TileCoord t;
    t.ks = item % g.k_splits;
    t.nt = (item / g.k_splits) % g.n_tiles;
    t.mt = item / (g.k_splits * g.n_tiles);
    t.k_begin = t.ks * g.k_chunk;
    t.k_end = min(g.K, t.k_begin + g.k_chunk);
    return t;
  }
  static __device__ __forceinline__ int num_stages(const Args&, const TileCoord& t) {
    return (t.k_end - t.k_begin + KB - 1) / KB;
  }
  static __device__ __forceinline__ void load_resident(const Args&, uint8_t*, int, int) {}

  static __device__ __forceinline__ void load_stage(const Args& g, const TileCoord& t, int s, uint8_t* st,
                                                    int glane, int gsize, Prod&) {
    const int k0 = t.k_begin + s * KB;
    // A: 128 rows x KB/8 chunks; lanes run along the rows of one chunk
    for (int v = glane; v < kTileM * (KB / 8); v += gsize) {
      const int r = v % kTileM, kc = v / kTileM, m = t.mt * kTileM + r;
      float x[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) x[j] = 0.f;
      if (m < g.M) {
        const Pos pr = lin3(g.a_row, m);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int k = k0 + kc * 8 + j;
          if (k < t.k_end) {
            const Pos pk = lin3(g.a_k, k);
            if ((unsigned)(pr.y + pk.y) < (unsigned)g.H && (unsigned)(pr.x + pk.x) < (unsigned)g.W)
              x[j] = __ldg(g.A + pr.off + pk.off);
          }
        }
      }
      tc::store_chunk_split(st, st + A_PART, kc * A_PLANE + r * 16, x);
    }
    // B^T: NT rows (n) x KB/8 chunks; lanes run along n
    for (int v = glane; v < NT * (KB / 8); v += gsize) {
      const int n = v % NT, kc = v / NT, nn = t.nt * NT + n;
      float x[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) x[j] = 0.f;
      if (nn < g.N) {
        const long long on = lin3(g.b_n, nn).off;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int k = k0 + kc * 8 + j;
          if (k < t.k_end) x[j] = __ldg(g.B + lin3(g.b_k, k).off + on);
        }
      }
      tc::store_chunk_split(st + B_OFF, st + B_OFF + NT * 16, kc * B_PLANE + n * 16, x);
    }
  }
  static __device__ __forceinline__ void issue(const Args&, const TileCoord&, int s, uint32_t st, uint32_t,
                                               uint32_t d_tmem) {
    constexpr uint32_t idesc2 = tc::make_idesc(2 * NT), idesc1 = tc::make_idesc(NT);
    const uint64_t da0 = tc::make_sdesc(st, A_PLANE), db0 = tc::make_sdesc(st + B_OFF, B_PLANE);
#pragma unroll
    for (int k16 = 0; k16 < KB / 16; ++k16) {
      const uint64_t da_hi = tc::sdesc_advance(da0, k16 * 2 * A_PLANE);
      const uint64_t da_lo = tc::sdesc_advance(da0, k16 * 2 * A_PLANE + A_PART);
      const uint64_t db = tc::sdesc_advance(db0, k16 * 2 * B_PLANE);
      tc::umma_f16(d_tmem, da_hi, db, idesc2, (s | k16) != 0 ? 1u : 0u);   // a_hi . [b_hi | b_lo]
      tc::umma_f16(d_tmem, da_lo, db, idesc1, 1u);                          // a_lo . b_hi
    }
  }
  // ---- epilogue: row m of the tile = NT consecutive floats of D
  static __device__ __forceinline__ float* row_ptr(const Args& g, const TileCoord& t, int row) {
    const int m = t.mt * kTileM + row;
    if (m >= g.M) return nullptr;
    return g.D + (long long)t.ks * g.d_slice + lin3(g.d_row, m).off + t.nt * NT;
  }
  static __device__ __forceinline__ bool seg_valid(const Args& g, const TileCoord& t, int sg) {
    return t.nt * NT + sg * SEG < g.N;
  }
  static __device__ __forceinline__ int64_t seg_offset(const Args&, const TileCoord&, int sg) { return sg * SEG; }
  static __device__ __forceinline__ bool col_valid(const Args& g, const TileCoord& t, int col) {
    return t.nt * NT + col < g.N;
  }
  static __device__ __forceinline__ int64_t row_aux(const Args& g, const TileCoord& t, int row) {
    const int m = t.mt * kTileM + row;
    return m < g.M ? lin3(g.d_row, m).off + t.nt * NT : 0;
  }
  static __device__ __forceinline__ float4 aux_load(const Args& g, const TileCoord& t, const float*, int col,
                                                    int64_t rowaux) {
    if (EPI == GG_BIAS) return tc::ldg4(g.bias + t.nt * NT + col);
    return tc::ldg4(g.mask + rowaux + col);
  }
  static __device__ __forceinline__ float4 finish(const Args& g, float4 o, float4 x) {
    o.x *= g.alpha; o.y *= g.alpha; o.z *= g.alpha; o.w *= g.alpha;
    if (EPI == GG_BIAS) {
      o.x += x.x; o.y += x.y; o.z += x.z; o.w += x.w;
      if (g.relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
    } else if (EPI == GG_MASK) {
      o.x = x.x > 0.f ? o.x : 0.f; o.y = x.y > 0.f ? o.y : 0.f;
      o.z = x.z > 0.f ? o.z : 0.f; o.w = x.w > 0.f ? o.w : 0.f;
    }
    return o;
  }
};

// splits the reduction only where the epilogue is a plain store (weight gradients)
int gather_gemm(GatherGemmArgs g, int epi, bool split_k, cudaStream_t st, int* splits_out = nullptr) {
  if (splits_out) *splits_out = 0;
  if (g.M <= 0 || g.N <= 0) return ARL_OK;
  g.m_tiles = (g.M + kTileM - 1) / kTileM;
  g.n_tiles = (g.N + 63) / 64;
  g.k_splits = 1;
  if (split_k) {
    const int tiles = g.m_tiles * g.n_tiles, want = 2 * num_sms();
    g.k_splits = tiles >= want ? 1 : (want + tiles - 1) / tiles;
    const int max_splits = (g.K + 31) / 32;
    if (g.k_splits > max_splits) g.k_splits = max_splits;
    if (g.k_splits < 1) g.k_splits = 1;
  }
  int chunk = (g.K + g.k_splits - 1) / g.k_splits;
  chunk = (chunk + 31) / 32 * 32;
  g.k_chunk = chunk;
  g.k_splits = (g.K + chunk - 1) / chunk;
  if (g.k_splits < 1) g.k_splits = 1;
  if (splits_out) *splits_out = g.k_splits;
  const int items = g.m_tiles * g.n_tiles * g.k_splits;
  if (epi == GG_BIAS) return tc::launch<GatherGemm<GG_BIAS>>(g, items, st);
  if (epi == GG_MASK) return tc::launch<GatherGemm<GG_MASK>>(g, items, st);
  return tc::launch<GatherGemm<GG_PLAIN>>(g, items, st);
}

// ---- layer geometry ----------------------------------------------------------------------------
struct ConvGeo { int H, W, C, KH, KW, S, OH, OW, CO; };
constexpr ConvGeo kConv1 = {84, 84, 4, 8, 8, 4, 20, 20, 32};
constexpr ConvGeo kConv2 = {20, 20, 32, 4, 4, 2, 9, 9, 64};
constexpr ConvGeo kConv3 = {9, 9, 64, 3, 3, 1, 7, 7, 64};
constexpr int kFcIn = 7 * 7 * 64, kHid = ARL_NATURE_FC;

Lin3 conv_rows(const ConvGeo& c) {     // m -> (n, oy, ox) on the INPUT tensor
  Lin3 L = {c.OH, c.OW, (long long)c.H * c.W * c.C, (long long)c.S * c.W * c.C, (long long)c.S * c.C, 0,
            0, c.S, 0, 0, 0, c.S};
  return L;
}
Lin3 conv_taps(const ConvGeo& c) {     // k -> (ky, kx, ch) on the INPUT tensor
  Lin3 L = {c.KW, c.C, (long long)c.W * c.C, c.C, 1, 0, 1, 0, 0, 0, 1, 0};
  return L;
}

// out = relu(alpha * conv(in, w) + b)                                     (ops.py:21-28)
int conv_forward(const ConvGeo& c, const float* in, const float* w, const float* b, float* out, int64_t n,
                 float alpha, cudaStream_t st) {
  GatherGemmArgs g = {};
  g.A = in; g.a_row = conv_rows(c); g.a_k = conv_taps(c); g.H = c.H; g.W = c.W;
  g.B = w; g.b_k = flat(c.CO); g.b_n = flat(1);
  g.D = out; g.d_row = flat(c.CO); g.d_slice = 0;
  g.bias = b; g.alpha = alpha; g.relu = 1;
  g.M = (int)(n * c.OH * c.OW); g.N = c.CO; g.K = c.KH * c.KW * c.C;
  return gather_gemm(g, GG_BIAS, false, st);
}
// dW[(ky,kx,ch), co] = alpha * sum_m in(m, k) * dy[m][co]  -> split-K partials -> grads
int conv_wgrad(const ConvGeo& c, const float* in, const float* dy, float* dw, float* workspace, int64_t n,
               float alpha, cudaStream_t st) {
  GatherGemmArgs g = {};
  g.A = in; g.a_row = conv_taps(c); g.a_k = conv_rows(c); g.H = c.H; g.W = c.W;     // transposed roles
  g.B = dy; g.b_k = flat(c.CO); g.b_n = flat(1);
  g.M = c.KH * c.KW * c.C; g.N = c.CO; g.K = (int)(n * c.OH * c.OW);
  g.D = workspace; g.d_row = flat(c.CO); g.d_slice = (long long)g.M * g.N;
  g.alpha = alpha;
  int splits = 0;
  int rc = gather_gemm(g, GG_PLAIN, true, st, &splits);
  if (rc) return rc;
  return reduce_partials(workspace, dw, splits, g.M * g.N, st);
}
// d_in = relu'(in) * convT(dy, w): one contraction per output-parity class (py, px) of the stride
int conv_dgrad(const ConvGeo& c, const float* dy, const float* w, const float* in_act, float* d_in, int64_t n,
               cudaStream_t st) {
  const int s = c.S, hs = c.H / s, ws = c.W / s, ah = c.KH / s, aw = c.KW / s;
  for (int py = 0; py < s; ++py)
    for (int px = 0; px < s; ++px) {
      GatherGemmArgs g = {};
      g.A = dy; g.H = c.OH; g.W = c.OW;
      Lin3 ar = {hs, ws, (long long)c.OH * c.OW * c.CO, (long long)c.OW * c.CO, c.CO, 0, 0, 1, 0, 0, 0, 1};
      Lin3 ak = {aw, c.CO, -(long long)c.OW * c.CO, -(long long)c.CO, 1, 0, -1, 0, 0, 0, -1, 0};
      g.a_row = ar; g.a_k = ak;
      g.B = w;
      Lin3 bk = {aw, c.CO, (long long)s * c.KW * c.C * c.CO, (long long)s * c.C * c.CO, 1,
                 (long long)(py * c.KW + px) * c.C * c.CO, 0, 0, 0, 0, 0, 0};
      g.b_k = bk; g.b_n = flat(c.CO);
      g.D = d_in;
      Lin3 dr = {hs, ws, (long long)c.H * c.W * c.C, (long long)s * c.W * c.C, (long long)s * c.C,
                 (long long)(py * c.W + px) * c.C, 0, 0, 0, 0, 0, 0};
      g.d_row = dr; g.d_slice = 0;
      g.mask = in_act; g.alpha = 1.f;
      g.M = (int)(n * hs * ws); g.N = c.C; g.K = ah * aw * c.CO;
      int rc = gather_gemm(g, GG_MASK, false, st);
      if (rc) return rc;
    }
  return ARL_OK;
}

// ---- small kernels: heads (512 hidden units), column sums ----------------------------------------
// one warp per sample; logits / value = h . [p_w | q_w] + [p_b | q_b]; probs = softmax(logits)
__global__ void __launch_bounds__(256)
nat_heads_fwd_kernel(const float* __restrict__ pw, const float* __restrict__ pb, const float* __restrict__ qw,
                     const float* __restrict__ qb, const float* __restrict__ h, float* __restrict__ logits,
                     float* __restrict__ probs, float* __restrict__ value, int64_t num_samples, int A) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t n = warp; n < num_samples; n += nwarps) {
    float hv[kHid / 32];
#pragma unroll
    for (int i = 0; i < kHid / 32; ++i) hv[i] = h[n * kHid + i * 32 + lane];
    float z = 0.f;
    for (int j = 0; j <= A; ++j) {
      float a = 0.f;
#pragma unroll
      for (int i = 0; i < kHid / 32; ++i) {
        const int k = i * 32 + lane;
        a = fmaf(hv[i], j < A ? pw[k * A + j] : qw[k], a);
      }
      a = warp_sum(a);
      if (j == lane) z = a + pb[j < A ? j : 0];
      if (j == A && lane == 0) value[n] = a + qb[0];
    }
    float mx = lane < A ? z : -INFINITY;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    const float e = lane < A ? expf(z - mx) : 0.f;
    const float den = warp_sum(e);
    if (lane < A) {
      logits[n * A + lane] = z;
      probs[n * A + lane] = e / den;
    }
  }
}
// d_h[n][k] = (h > 0) * (sum_j dlogits[n][j] p_w[k][j] + dvalue[n] q_w[k]); thread = (n, k)
__global__ void nat_heads_dgrad_kernel(const float* __restrict__ pw, const float* __restrict__ qw,
                                       const float* __restrict__ h, const float* __restrict__ dlogits,
                                       const float* __restrict__ dvalue, float* __restrict__ d_h,
                                       int64_t total, int A) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t n = i / kHid;
    const int k = (int)(i - n * kHid);
    float d = dvalue[n] * qw[k];
    for (int j = 0; j < A; ++j) d = fmaf(dlogits[n * A + j], pw[k * A + j], d);
    d_h[i] = h[i] > 0.f ? d : 0.f;
  }
}
// partial[blk][k][j] = sum over the block's samples of h[n][k] * dz[n][j]; thread = hidden unit k;
// partial layout = the flat-buffer order p_w [512][A] | p_b [A] | q_w [512] | q_b [1]
constexpr int kNatJ = ARL_MAX_ACTIONS + 1;
__global__ void __launch_bounds__(kHid)
nat_heads_wgrad_kernel(const float* __restrict__ h, const float* __restrict__ dlogits,
                       const float* __restrict__ dvalue, float* __restrict__ partials, int64_t num_samples,
                       int64_t per, int A) {
  __shared__ float dz[32][kNatJ];
  const int k = threadIdx.x, J = A + 1;
  float acc[kNatJ];
#pragma unroll
  for (int j = 0; j < kNatJ; ++j) acc[j] = 0.f;
  float bacc = 0.f;
  const int64_t beg = per * blockIdx.x, end = beg + per < num_samples ? beg + per : num_samples;
  for (int64_t c0 = beg; c0 < end; c0 += 32) {
    const int nc = (int)(end - c0 < 32 ? end - c0 : 32);
    __syncthreads();
    for (int i = k; i < nc * J; i += kHid) {
      const int s = i / J, j = i - s * J;
      dz[s][j] = j < A ? dlogits[(c0 + s) * A + j] : dvalue[c0 + s];
    }
    __syncthreads();
    for (int s = 0; s < nc; ++s) {
      const float hv = h[(c0 + s) * kHid + k];
#pragma unroll
      for (int j = 0; j < kNatJ; ++j)
        if (j < J) acc[j] = fmaf(hv, dz[s][j], acc[j]);
    }
    if (k < J)
      for (int s = 0; s < nc; ++s) bacc += dz[s][k];
  }
  float* out = partials + (size_t)blockIdx.x * (kHid * J + J);
#pragma unroll
  for (int j = 0; j < kNatJ; ++j) {
    if (j < A) out[k * A + j] = acc[j];
    else if (j == A) out[kHid * A + A + k] = acc[j];
  }
  if (k < A) out[kHid * A + k] = bacc;
  else if (k == A) out[kHid * A + A + kHid] = bacc;
}
// partial[blk][c] = sum over the block's rows of x[row][c]  (bias gradients).  256 threads =
// (row lane, column): every lane walks its rows with a stride of the lane count, the lanes of a
// column are then added in a fixed order (deterministic, no atomics).
__global__ void __launch_bounds__(256)
nat_colsum_kernel(const float* __restrict__ x, float* __restrict__ partials, int64_t rows, int64_t per, int C) {
  extern __shared__ float cs[];                       // [row lanes][C]
  const int cw = C < 256 ? C : 256, RL = 256 / cw;
  const int rl = threadIdx.x / cw, c0 = threadIdx.x % cw;
  const int64_t beg = per * blockIdx.x, end = beg + per < rows ? beg + per : rows;
  if (rl < RL)
    for (int c = c0; c < C; c += cw) {
      float s = 0.f;
      for (int64_t r = beg + rl; r < end; r += RL) s += x[r * C + c];
      cs[rl * C + c] = s;
    }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += 256) {
    float s = 0.f;
    for (int l = 0; l < RL; ++l) s += cs[l * C + c];
    partials[(size_t)blockIdx.x * C + c] = s;
  }
}
int colsum(const float* x, float* out, float* workspace, int64_t rows, int C, cudaStream_t st) {
  int blocks = (int)(rows < 4LL * num_sms() ? (rows > 0 ? rows : 1) : 4LL * num_sms());
  const int64_t per = (rows + blocks - 1) / blocks > 0 ? (rows + blocks - 1) / blocks : 1;
  blocks = (int)((rows + per - 1) / per);
  if (blocks < 1) blocks = 1;
  const int cw = C < 256 ? C : 256, RL = 256 / cw;
  nat_colsum_kernel<<<blocks, 256, (size_t)RL * C * sizeof(float), st>>>(x, workspace, rows, per, C);
  ARL_LAUNCH_CHECK("nat_colsum_kernel");
  return reduce_partials(workspace, out, blocks, C, st);
}

struct NatLayout { int64_t off[ARL_NATURE_TENSORS + 1]; };
NatLayout nat_layout(int A) {
  const int64_t sz[ARL_NATURE_TENSORS] = {8 * 8 * 4 * 32, 32, 4 * 4 * 32 * 64, 64, 3 * 3 * 64 * 64, 64,
                                          (int64_t)kFcIn * kHid, kHid, (int64_t)kHid * A, A, kHid, 1};
  NatLayout L;
  L.off[0] = 0;
  for (int i = 0; i < ARL_NATURE_TENSORS; ++i) L.off[i + 1] = L.off[i] + sz[i];
  return L;
}
enum { N_L1W = 0, N_L1B, N_L2W, N_L2B, N_L3W, N_L3B, N_L4W, N_L4B, N_PW, N_PB, N_QW, N_QB };

}  // namespace
}  // namespace arl

using namespace arl;

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

extern "C" int arl_nature_param_layout(int action_size, int64_t* offsets) {
  ARL_REQUIRE(offsets, "arl_nature_param_layout: null pointer");
  ARL_REQUIRE(action_size >= 1 && action_size <= ARL_MAX_ACTIONS,
              "arl_nature_param_layout: action_size %d outside [1,%d]", action_size, ARL_MAX_ACTIONS);
  const NatLayout L = nat_layout(action_size);
  for (int i = 0; i <= ARL_NATURE_TENSORS; ++i) offsets[i] = L.off[i];
  return ARL_OK;
}

extern "C" int64_t arl_nature_workspace_bytes(int action_size) {
  // largest user: split-K partials of a weight gradient (<= 2 slices of 3136 x 512, or <= 2*SMs
  // slices of a conv tensor), then the heads / column-sum partials
  (void)action_size;
  return (int64_t)64 << 20;
}

extern "C" int arl_nature_forward(const float* params, int action_size, const float* x, float* a1, float* a2,
                                  float* a3, float* h, float* logits, float* probs, float* value,
                                  int64_t num_samples, void* stream) {
  ARL_REQUIRE(params && x && a1 && a2 && a3 && h && logits && probs && value, "arl_nature_forward: null pointer");
  ARL_REQUIRE(action_size >= 1 && action_size <= ARL_MAX_ACTIONS, "arl_nature_forward: bad action_size %d", action_size);
  ARL_REQUIRE(num_samples >= 0 && num_samples * 400 < (1LL << 30), "arl_nature_forward: bad num_samples");
  ARL_REQUIRE(aligned16(params) && aligned16(x) && aligned16(a1) && aligned16(a2) && aligned16(a3) && aligned16(h),
              "arl_nature_forward: pointers must be 16-byte aligned");
  if (num_samples == 0) return ARL_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const NatLayout L = nat_layout(action_size);
  const float* p = params;
  // network.py:33-40: l0 = s_t / 255 (folded into alpha), three conv2d(..., relu)
  int rc = conv_forward(kConv1, x, p + L.off[N_L1W], p + L.off[N_L1B], a1, num_samples, 1.0f / 255.0f, st);
  if (rc) return rc;
  rc = conv_forward(kConv2, a1, p + L.off[N_L2W], p + L.off[N_L2B], a2, num_samples, 1.0f, st);
  if (rc) return rc;
  rc = conv_forward(kConv3, a2, p + L.off[N_L3W], p + L.off[N_L3B], a3, num_samples, 1.0f, st);
  if (rc) return rc;
  // network.py:41-42: linear(l3, 512, relu) on the NHWC flatten (the reference passes the 4-D
  // tensor to linear, which cannot run; the flatten is agent.py:231-232's)
  GatherGemmArgs g = {};
  g.A = a3; g.a_row = flat(kFcIn); g.a_k = flat(1); g.H = 1; g.W = 1;
  g.B = p + L.off[N_L4W]; g.b_k = flat(kHid); g.b_n = flat(1);
  g.D = h; g.d_row = flat(kHid);
  g.bias = p + L.off[N_L4B]; g.alpha = 1.f; g.relu = 1;
  g.M = (int)num_samples; g.N = kHid; g.K = kFcIn;
  rc = gather_gemm(g, GG_BIAS, false, st);
  if (rc) return rc;
  const int64_t warps = num_samples;
  int grid = (int)((warps + 7) / 8 < 4LL * num_sms() ? (warps + 7) / 8 : 4LL * num_sms());
  nat_heads_fwd_kernel<<<grid, 256, 0, st>>>(p + L.off[N_PW], p + L.off[N_PB], p + L.off[N_QW], p + L.off[N_QB],
                                             h, logits, probs, value, num_samples, action_size);
  ARL_LAUNCH_CHECK("nat_heads_fwd_kernel");
  return ARL_OK;
}

extern "C" int arl_nature_backward(const float* params, int action_size, const float* x, const float* a1,
                                   const float* a2, const float* a3, const float* h, const float* dlogits,
                                   const float* dvalue, float* d_h, float* d_a3, float* d_a2, float* d_a1,
                                   float* grads, void* workspace, int64_t num_samples, void* stream) {
  ARL_REQUIRE(params && x && a1 && a2 && a3 && h && dlogits && dvalue && d_h && d_a3 && d_a2 && d_a1 && grads &&
                  workspace, "arl_nature_backward: null pointer");
  ARL_REQUIRE(action_size >= 1 && action_size <= ARL_MAX_ACTIONS, "arl_nature_backward: bad action_size %d", action_size);
  ARL_REQUIRE(num_samples >= 0 && num_samples * 400 < (1LL << 30), "arl_nature_backward: bad num_samples");
  cudaStream_t st = (cudaStream_t)stream;
  const NatLayout L = nat_layout(action_size);
  const int A = action_size, J = A + 1;
  const int64_t N = num_samples;
  if (N == 0) {
    ARL_CUDA(cudaMemsetAsync(grads, 0, (size_t)L.off[ARL_NATURE_TENSORS] * sizeof(float), st));
    return ARL_OK;
  }
  const float* p = params;
  float* ws = (float*)workspace;
  // heads: weight / bias gradients (deterministic per-block partials), then d_h with h's relu mask
  {
    int blocks = (int)(N < 2LL * num_sms() ? N : 2LL * num_sms());
    const int64_t per = (N + blocks - 1) / blocks;
    blocks = (int)((N + per - 1) / per);
    nat_heads_wgrad_kernel<<<blocks, kHid, 0, st>>>(h, dlogits, dvalue, ws, N, per, A);
    ARL_LAUNCH_CHECK("nat_heads_wgrad_kernel");
    int rc = reduce_partials(ws, grads + L.off[N_PW], blocks, kHid * J + J, st);
    if (rc) return rc;
    const int64_t total = N * kHid;
    const int grid = (int)((total + 255) / 256 < 8LL * num_sms() ? (total + 255) / 256 : 8LL * num_sms());
    nat_heads_dgrad_kernel<<<grid, 256, 0, st>>>(p + L.off[N_PW], p + L.off[N_QW], h, dlogits, dvalue, d_h, total, A);
    ARL_LAUNCH_CHECK("nat_heads_dgrad_kernel");
  }
  // fc512: l4_b = column sums of d_h; l4_w = a3^T . d_h; d_a3 = relu'(a3) * d_h . l4_w^T
  int rc = colsum(d_h, grads + L.off[N_L4B], ws, N, kHid, st);
  if (rc) return rc;
  {
    GatherGemmArgs g = {};
    g.A = a3; g.a_row = flat(1); g.a_k = flat(kFcIn); g.H = 1; g.W = 1;      // A'(kk, n) = a3[n][kk]
    g.B = d_h; g.b_k = flat(kHid); g.b_n = flat(1);
    g.M = kFcIn; g.N = kHid; g.K = (int)N;
    g.D = ws; g.d_row = flat(kHid); g.d_slice = (long long)kFcIn * kHid; g.alpha = 1.f;
    int splits = 0;
    rc = gather_gemm(g, GG_PLAIN, true, st, &splits);
    if (rc) return rc;
    rc = reduce_partials(ws, grads + L.off[N_L4W], splits, kFcIn * kHid, st);
    if (rc) return rc;
  }
  {
    GatherGemmArgs g = {};
    g.A = d_h; g.a_row = flat(kHid); g.a_k = flat(1); g.H = 1; g.W = 1;
    g.B = p + L.off[N_L4W]; g.b_k = flat(1); g.b_n = flat(kHid);               // B(j, kk) = l4_w[kk][j]
    g.D = d_a3; g.d_row = flat(kFcIn); g.mask = a3; g.alpha = 1.f;
    g.M = (int)N; g.N = kFcIn; g.K = kHid;
    rc = gather_gemm(g, GG_MASK, false, st);
    if (rc) return rc;
  }
  // conv3
  rc = colsum(d_a3, grads + L.off[N_L3B], ws, N * 49, 64, st);
  if (rc) return rc;
  rc = conv_wgrad(kConv3, a2, d_a3, grads + L.off[N_L3W], ws, N, 1.f, st);
  if (rc) return rc;
  rc = conv_dgrad(kConv3, d_a3, p + L.off[N_L3W], a2, d_a2, N, st);
  if (rc) return rc;
  // conv2
  rc = colsum(d_a2, grads + L.off[N_L2B], ws, N * 81, 64, st);
  if (rc) return rc;
  rc = conv_wgrad(kConv2, a1, d_a2, grads + L.off[N_L2W], ws, N, 1.f, st);
  if (rc) return rc;
  rc = conv_dgrad(kConv2, d_a2, p + L.off[N_L2W], a1, d_a1, N, st);
  if (rc) return rc;
  // conv1 (no input gradient; the 1/255 of network.py:33 scales the weight gradient)
  rc = colsum(d_a1, grads + L.off[N_L1B], ws, N * 400, 32, st);
  if (rc) return rc;
  return conv_wgrad(kConv1, x, d_a1, grads + L.off[N_L1W], ws, N, 1.0f / 255.0f, st);
}
