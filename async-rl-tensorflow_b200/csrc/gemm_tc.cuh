// tcgen05 contraction framework (sm_100a).  Every dense contraction of the net (fc256 and the
// two convolutions, forward and backward) is an instance of
//     D[128-row tile, N_TILE] (fp32, TMEM) += A[128, K] . B[N_TILE, K]^T
// where the operand tiles are BUILT IN SHARED MEMORY BY PRODUCER WARPS straight from the
// tensors in HBM (fp32 activations / u8 frames / fp32 weights): gather (im2col), split into
// bf16 hi + bf16 lo ("bf16x3": hi*hi + hi*lo + lo*hi with fp32 accumulation, relative error
// ~2^-16 instead of bf16's 2^-8; u8 pixels are exact in bf16 and need no lo part) and store in
// the UMMA canonical K-major no-swizzle layout.  A Policy supplies the gather, the tile
// decomposition and the epilogue; this file owns the pipeline.
//
// One persistent CTA per SM, 416 threads:
//   warps 0-3   epilogue : tcgen05.ld (TMEM lane quadrant = warp id) -> Policy::store -> HBM
//   warps 4-11  producers: 8/STAGES warps own each smem stage and fill it independently of the
//                          other stages (STAGES load->convert->store chains in flight);
//                          fence.proxy.async + mbarrier arrive hand the stage to the MMA warp
//   warp 12     MMA      : one thread issues tcgen05.mma (cta_group::1, kind::f16, M=128,
//                          N=N_TILE, K=16); tcgen05.commit frees the stage / publishes the
//                          accumulator (2 TMEM accumulator stages); owns tcgen05.alloc/dealloc
//
// smem operand image (bf16): element (row r, k) of a [ROWS x KB] tile sits at
//   (k/8)*LBO + r*16 + (k%8)*2,  LBO = (ROWS+1)*16 (one 16-B pad keeps st.shared conflict-free),
//   SBO = 128 (the 8 rows of a core matrix are contiguous).
#pragma once
#include <cuda_bf16.h>

#include "common.cuh"

namespace arl {
namespace tc {

constexpr int kEpiWarps = 4;
constexpr int kTileM = 128;
// per policy: PROD_WARPS producer warps (8 or 16); warp layout = [4 epilogue | producers | MMA]
template <class P> __host__ __device__ constexpr int prod_threads() { return P::PROD_WARPS * 32; }
template <class P> __host__ __device__ constexpr int cta_threads() { return (kEpiWarps + P::PROD_WARPS + 1) * 32; }

// ---- PTX wrappers ---------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   smem_u32(smem_dst)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// D[tmem] (+)= A[smem desc] . B[smem desc]
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                         uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// mbarrier arrives once all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
          smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,"
      "%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// instruction descriptor: bf16 x bf16 -> f32, M=128; operands K-major unless *_mn
__host__ __device__ constexpr uint32_t make_idesc(int n, bool a_mn = false, bool b_mn = false) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) |
         ((uint32_t)(n >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);
}
// shared-memory matrix descriptor, SWIZZLE_NONE.  The operand images in this file place the 16-B
// vector of (index r of the 16-B-strided axis, chunk c of the other axis) at c*PLANE + r*16.
//   K-major operand : r = row (M/N), c = k/8   -> LBO = PLANE (next K chunk), SBO = 128
//   MN-major operand: r = k,         c = row/8 -> LBO = 128 (next 8 k), SBO = PLANE (next 8 rows)
__device__ __forceinline__ uint64_t make_sdesc(uint32_t smem_addr, uint32_t lbo_bytes,
                                               uint32_t sbo_bytes = 128) {
  return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) |
         ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46);
}

template <int ROWS, int KB>
struct OperandTile {
  static constexpr int KC = KB / 8;
  static constexpr int LBO = (ROWS + 1) * 16;
  static constexpr int BYTES = KC * LBO;          // one image (hi or lo)
};

// ---- chunk helpers (a chunk = 8 consecutive k of one row = 16 B of bf16) ---------------------
__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  const float2 hf = __bfloat1622float2(h);
  const __nv_bfloat162 l = __floats2bfloat162_rn(a - hf.x, b - hf.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}
__device__ __forceinline__ void store_chunk_split(uint8_t* hi_img, uint8_t* lo_img, int off,
                                                  const float (&x)[8]) {
  uint4 h, l;
  split2(x[0], x[1], h.x, l.x);
  split2(x[2], x[3], h.y, l.y);
  split2(x[4], x[5], h.z, l.z);
  split2(x[6], x[7], h.w, l.w);
  *reinterpret_cast<uint4*>(hi_img + off) = h;
  *reinterpret_cast<uint4*>(lo_img + off) = l;
}
// 8 bytes (two u32 words) -> 8 exact bf16
__device__ __forceinline__ uint32_t bytes2bf16x2(uint32_t w, int sel_lo, int sel_hi) {
  // 0x4B000000 | byte = 8388608 + byte as fp32; subtracting 8388608 is exact
  const float a = __uint_as_float(__byte_perm(w, 0x4B000000u, sel_lo)) - 8388608.0f;
  const float b = __uint_as_float(__byte_perm(w, 0x4B000000u, sel_hi)) - 8388608.0f;
  const __nv_bfloat162 p = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&p);
}
__device__ __forceinline__ uint4 bytes8_to_bf16(uint32_t w0, uint32_t w1) {
  // selector: result byte0 = src byte i, bytes 1..3 = 0x00,0x00,0x4B of the magic word
  uint4 o;
  o.x = bytes2bf16x2(w0, 0x7440, 0x7441);
  o.y = bytes2bf16x2(w0, 0x7442, 0x7443);
  o.z = bytes2bf16x2(w1, 0x7440, 0x7441);
  o.w = bytes2bf16x2(w1, 0x7442, 0x7443);
  return o;
}

// Generic fp32 operand tile loader used by several policies, one WARP fills the tile.
// TRANS=false: element (r,k) at src[r*ld + k] (k contiguous; r < rmax checked per row,
//              k < kmax per 8-chunk).  TRANS=true: element (r,k) at src[k*ld + r] (r contiguous;
//              r checked per 8-row group, k per element).
template <int ROWS, int KB, bool TRANS>
__device__ __forceinline__ void load_tile_f32(uint8_t* hi_img, uint8_t* lo_img,
                                              const float* __restrict__ src, int64_t ld, int r0,
                                              int rmax, int k0, int kmax, int lane) {
  using T = OperandTile<ROWS, KB>;
  if (!TRANS) {
#pragma unroll 2
    for (int c = lane; c < ROWS * T::KC; c += 32) {
      const int row = c % ROWS, kc = c / ROWS;
      const int r = r0 + row, k = k0 + kc * 8;
      float x[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      if (r < rmax && k < kmax) {
        const float4* p = reinterpret_cast<const float4*>(src + (int64_t)r * ld + k);
        const float4 x0 = p[0], x1 = p[1];
        x[0] = x0.x; x[1] = x0.y; x[2] = x0.z; x[3] = x0.w;
        x[4] = x1.x; x[5] = x1.y; x[6] = x1.z; x[7] = x1.w;
      }
      store_chunk_split(hi_img, lo_img, kc * T::LBO + row * 16, x);
    }
  } else {
    // 8 (rows) x 8 (k) blocks, lanes along the k-chunks (pad in LBO => conflict-free stores)
    for (int b = lane; b < (ROWS / 8) * T::KC; b += 32) {
      const int kc = b % T::KC, rg = b / T::KC;
      const int r = r0 + rg * 8, k = k0 + kc * 8;
      float v[8][8];                                  // [k][row]
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) {
        float4 x0 = make_float4(0.f, 0.f, 0.f, 0.f), x1 = x0;
        if (r < rmax && k + kk < kmax) {
          const float4* p = reinterpret_cast<const float4*>(src + (int64_t)(k + kk) * ld + r);
          x0 = p[0];
          x1 = p[1];
        }
        v[kk][0] = x0.x; v[kk][1] = x0.y; v[kk][2] = x0.z; v[kk][3] = x0.w;
        v[kk][4] = x1.x; v[kk][5] = x1.y; v[kk][6] = x1.z; v[kk][7] = x1.w;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float x[8] = {v[0][i], v[1][i], v[2][i], v[3][i], v[4][i], v[5][i], v[6][i], v[7][i]};
        store_chunk_split(hi_img, lo_img, kc * T::LBO + (rg * 8 + i) * 16, x);
      }
    }
  }
}

struct TileCoord {
  int mt, nt, ks;        // tile indices (meaning is the policy's)
  int k_begin, k_end;    // reduction range of this work item (meaning is the policy's)
};

// A Policy provides:
//   Args; PROD_WARPS (8 or 16); STAGES (2, 4 or 8); STAGE_BYTES; RES_BYTES (resident smem, e.g. converted weights);
//   ACC_COLS (TMEM columns of one accumulator set, multiple of 16)
//   static __device__ int num_items(const Args&)
//   static __device__ TileCoord coord(const Args&, int item)
//   static __device__ int num_stages(const Args&, const TileCoord&)          stages per item
//   static __device__ void load_resident(const Args&, uint8_t* res, int ptid, int nthreads)
//   static __device__ void load_stage(const Args&, const TileCoord&, int s, uint8_t* stage,
//                                     int glane, int gsize)     lanes of the stage's warp group
//   static __device__ void issue(const Args&, const TileCoord&, int s, uint32_t stage_addr,
//                                uint32_t res_addr, uint32_t d_tmem)   all MMAs of stage s
//                                (s == 0 must start with accumulate = 0)
//   static __device__ void store(const Args&, const TileCoord&, int row, int col, const float (&v)[16])
template <class P>
struct Smem {
  static constexpr int BAR_OFF = P::STAGES * P::STAGE_BYTES + P::RES_BYTES;
  static constexpr int TOTAL = BAR_OFF + 256;
};

template <class P>
__global__ void __launch_bounds__(cta_threads<P>(), 1) tc_kernel(typename P::Args g) {
  using S = Smem<P>;
  constexpr int STAGES = P::STAGES, WPS = P::PROD_WARPS / STAGES;  // producer warps per stage
  constexpr int kMmaWarp = kEpiWarps + P::PROD_WARPS;
  static_assert(P::PROD_WARPS % STAGES == 0, "producer warps must divide evenly over the stages");
  static_assert(STAGES == 8 || STAGES == 4 || STAGES == 2, "STAGES");
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* res = smem + STAGES * P::STAGE_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + S::BAR_OFF);
  uint64_t* empty = full + STAGES;
  uint64_t* tfull = empty + STAGES;      // [2]
  uint64_t* tempty = tfull + 2;          // [2]
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr uint32_t kTmemCols = 2 * P::ACC_COLS <= 32 ? 32 : 2 * P::ACC_COLS <= 64 ? 64
                               : 2 * P::ACC_COLS <= 128 ? 128 : 2 * P::ACC_COLS <= 256 ? 256 : 512;
  static_assert(2 * P::ACC_COLS <= 512 && P::ACC_COLS % 16 == 0, "TMEM budget");

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 32 * WPS);
      mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull[a], 1);
      mbar_init(&tempty[a], kEpiWarps * 32);
    }
    fence_mbar_init();
  }
  if (warp == kMmaWarp) tmem_alloc(tmem_base_slot, kTmemCols);
  if (P::RES_BYTES > 0 && warp >= kEpiWarps && warp < kMmaWarp) {
    P::load_resident(g, res, tid - kEpiWarps * 32, prod_threads<P>());
    fence_proxy_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;
  const int items = P::num_items(g);

  if (warp >= kEpiWarps && warp < kMmaWarp) {
    // ===================== producers: warp group pg owns stage pg =====================
    const int pw = warp - kEpiWarps;
    const int pg = pw / WPS, glane = (pw % WPS) * 32 + lane;
    uint8_t* st = smem + pg * P::STAGE_BYTES;
    int i = 0;                                         // running stage index over all items
    for (int item = blockIdx.x; item < items; item += gridDim.x) {
      const TileCoord tc = P::coord(g, item);
      const int ns = P::num_stages(g, tc);
      for (int s = 0; s < ns; ++s, ++i) {
        if (i % STAGES != pg) continue;
        mbar_wait(&empty[pg], ((i / STAGES) & 1) ^ 1);
        P::load_stage(g, tc, s, st, glane, 32 * WPS);
        fence_proxy_async_smem();        // generic-proxy stores -> visible to the MMA (async proxy)
        mbar_arrive(&full[pg]);
      }
    }
  } else if (warp == kMmaWarp) {
    // ===================== MMA issuer (one thread) =====================
    if (lane == 0) {
      int i = 0;
      uint32_t acc = 0, acc_phase = 0;
      const uint32_t res_addr = smem_u32(res);
      for (int item = blockIdx.x; item < items; item += gridDim.x) {
        const TileCoord tc = P::coord(g, item);
        const int ns = P::num_stages(g, tc);
        mbar_wait(&tempty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * P::ACC_COLS;
        for (int s = 0; s < ns; ++s, ++i) {
          const int stage = i % STAGES;
          mbar_wait(&full[stage], (i / STAGES) & 1);
          tc_fence_after();
          P::issue(g, tc, s, smem_u32(smem + stage * P::STAGE_BYTES), res_addr, d_tmem);
          umma_commit(&empty[stage]);        // frees the smem stage when these MMAs retire
        }
        umma_commit(&tfull[acc]);            // accumulator ready for the epilogue
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else {
    // ===================== epilogue =====================
    uint32_t acc = 0, acc_phase = 0;
    for (int item = blockIdx.x; item < items; item += gridDim.x) {
      const TileCoord tc = P::coord(g, item);
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      const int row = warp * 32 + lane;
      const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + acc * P::ACC_COLS;
#pragma unroll 1
      for (int c = 0; c < P::ACC_COLS; c += 16) {
        float v[16];
        tmem_ld16(taddr + c, v);
        P::store(g, tc, row, c, v);
      }
      tc_fence_before();
      mbar_arrive(&tempty[acc]);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

template <class P>
int launch(const typename P::Args& g, int items, cudaStream_t stream) {
  using S = Smem<P>;
  static_assert(S::TOTAL <= 227 * 1024, "smem budget");
  auto kern = tc_kernel<P>;
  static bool attr_set = false;
  if (!attr_set) {
    ARL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL));
    attr_set = true;
  }
  if (items <= 0) return ARL_OK;
  const int grid = items < num_sms() ? items : num_sms();
  kern<<<grid, cta_threads<P>(), S::TOTAL, stream>>>(g);
  ARL_LAUNCH_CHECK("tc_kernel");
  return ARL_OK;
}

// =============================== plain GEMM policy (fc256) ===================================
enum { EPI_PLAIN = 0, EPI_BIAS_RELU = 1, EPI_MASK = 2 };

struct GemmArgs {
  const float* A;      // element (m,k): A_TRANS ? A[k*lda + m] : A[m*lda + k]
  const float* B;      // element (n,k): B_TRANS ? B[k*ldb + n] : B[n*ldb + k]
  float* D;            // D[m*ldd + n]; split-K slice z writes D + z*M*ldd
  const float* extra;  // EPI_BIAS_RELU: bias[n];  EPI_MASK: mask[m*ldd + n] (> 0 keeps)
  int M, N, K;
  int64_t lda, ldb, ldd;
  int k_chunk, k_splits;   // K range per split-K slice (multiple of KB); ceil(K / k_chunk)
  int m_tiles, n_tiles;
};

// MN=false: both operand images K-major (sample-major sources are transposed in registers).
// MN=true : requires A_TRANS && B_TRANS; the images are MN-major (the natural layout of
//           sample-major sources: 8 consecutive rows of one k form the 16-B vector).
template <int N_TILE_, int KB_, bool A_TRANS, bool B_TRANS, int EPI, bool MN = false>
struct GemmPolicy {
  using Args = GemmArgs;
  static constexpr int N_TILE = N_TILE_, KB = KB_, STAGES = 8, ACC_COLS = N_TILE_;
  static constexpr int PROD_WARPS = (MN || (!A_TRANS && !B_TRANS)) ? 16 : 8;   // register budget
  using TA = OperandTile<kTileM, KB>;
  using TB = OperandTile<N_TILE, KB>;
  // MN-major images: plane = 8 rows x KB k  -> (KB+1)*16 bytes per plane, ROWS/8 planes
  static constexpr int PLANE_MN = (KB + 1) * 16;
  static constexpr int A_BYTES = MN ? (kTileM / 8) * PLANE_MN : TA::BYTES;     // one image
  static constexpr int B_BYTES = MN ? (N_TILE / 8) * PLANE_MN : TB::BYTES;
  static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
  static constexpr int RES_BYTES = 0;
  static_assert(!MN || (A_TRANS && B_TRANS), "MN-major images need sample-major sources");

  static __device__ __forceinline__ int num_items(const Args& g) {
    return g.m_tiles * g.n_tiles * g.k_splits;
  }
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

  // MN-major image of a sample-major source: vector (k, row group rg) at rg*PLANE_MN + k*16
  template <int ROWS>
  static __device__ __forceinline__ void load_mn(uint8_t* hi, uint8_t* lo, const float* src,
                                                 int64_t ld, int r0, int rmax, int k0, int kmax,
                                                 int lane) {
    for (int c = lane; c < (ROWS / 8) * KB; c += 32) {
      const int rg = c % (ROWS / 8), kk = c / (ROWS / 8);
      const int r = r0 + rg * 8, k = k0 + kk;
      float x[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      if (r < rmax && k < kmax) {
        const float4* p = reinterpret_cast<const float4*>(src + (int64_t)k * ld + r);
        const float4 x0 = p[0], x1 = p[1];
        x[0] = x0.x; x[1] = x0.y; x[2] = x0.z; x[3] = x0.w;
        x[4] = x1.x; x[5] = x1.y; x[6] = x1.z; x[7] = x1.w;
      }
      store_chunk_split(hi, lo, rg * PLANE_MN + kk * 16, x);
    }
  }
  static __device__ __forceinline__ void load_stage(const Args& g, const TileCoord& t, int s,
                                                    uint8_t* st, int lane, int) {
    const int k0 = t.k_begin + s * KB;
    uint8_t* a_hi = st, *a_lo = st + A_BYTES, *b_hi = st + 2 * A_BYTES, *b_lo = b_hi + B_BYTES;
    if (MN) {
      load_mn<kTileM>(a_hi, a_lo, g.A, g.lda, t.mt * kTileM, g.M, k0, t.k_end, lane);
      load_mn<N_TILE>(b_hi, b_lo, g.B, g.ldb, t.nt * N_TILE, g.N, k0, t.k_end, lane);
    } else {
      load_tile_f32<kTileM, KB, A_TRANS>(a_hi, a_lo, g.A, g.lda, t.mt * kTileM, g.M, k0, t.k_end, lane);
      load_tile_f32<N_TILE, KB, B_TRANS>(b_hi, b_lo, g.B, g.ldb, t.nt * N_TILE, g.N, k0, t.k_end, lane);
    }
  }
  static __device__ __forceinline__ void issue(const Args&, const TileCoord&, int s,
                                               uint32_t st, uint32_t, uint32_t d_tmem) {
    constexpr uint32_t idesc = make_idesc(N_TILE, MN, MN);
    const uint32_t a_hi = st, a_lo = st + A_BYTES, b_hi = st + 2 * A_BYTES, b_lo = b_hi + B_BYTES;
#pragma unroll
    for (int k16 = 0; k16 < KB / 16; ++k16) {
      uint64_t da_hi, da_lo, db_hi, db_lo;
      if (MN) {
        const uint32_t o = k16 * 256;                 // 16 k rows of 16 B
        da_hi = make_sdesc(a_hi + o, 128, PLANE_MN); da_lo = make_sdesc(a_lo + o, 128, PLANE_MN);
        db_hi = make_sdesc(b_hi + o, 128, PLANE_MN); db_lo = make_sdesc(b_lo + o, 128, PLANE_MN);
      } else {
        const uint32_t ao = k16 * 2 * TA::LBO, bo = k16 * 2 * TB::LBO;
        da_hi = make_sdesc(a_hi + ao, TA::LBO); da_lo = make_sdesc(a_lo + ao, TA::LBO);
        db_hi = make_sdesc(b_hi + bo, TB::LBO); db_lo = make_sdesc(b_lo + bo, TB::LBO);
      }
      umma_f16(d_tmem, da_hi, db_hi, idesc, (s | k16) != 0 ? 1u : 0u);
      umma_f16(d_tmem, da_hi, db_lo, idesc, 1u);
      umma_f16(d_tmem, da_lo, db_hi, idesc, 1u);
    }
  }
  static __device__ __forceinline__ void store(const Args& g, const TileCoord& t, int row, int c,
                                               const float (&v)[16]) {
    const int m = t.mt * kTileM + row, n = t.nt * N_TILE + c;
    if (m >= g.M || n >= g.N) return;
    float* d = g.D + (int64_t)t.ks * g.M * g.ldd + (int64_t)m * g.ldd + n;
    const float* x = EPI == EPI_MASK ? g.extra + (int64_t)m * g.ldd + n : g.extra + n;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float4 o = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
      if (EPI == EPI_BIAS_RELU) {
        const float4 bb = *reinterpret_cast<const float4*>(x + 4 * q);
        o.x = fmaxf(o.x + bb.x, 0.f); o.y = fmaxf(o.y + bb.y, 0.f);
        o.z = fmaxf(o.z + bb.z, 0.f); o.w = fmaxf(o.w + bb.w, 0.f);
      } else if (EPI == EPI_MASK) {
        const float4 mm = *reinterpret_cast<const float4*>(x + 4 * q);
        o.x = mm.x > 0.f ? o.x : 0.f; o.y = mm.y > 0.f ? o.y : 0.f;
        o.z = mm.z > 0.f ? o.z : 0.f; o.w = mm.w > 0.f ? o.w : 0.f;
      }
      *reinterpret_cast<float4*>(d + 4 * q) = o;
    }
  }
};

}  // namespace tc
}  // namespace arl
