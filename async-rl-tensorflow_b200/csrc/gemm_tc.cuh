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
//   warps 4-11  producers: producer warp w owns smem stage w (8 stages): it fills k-blocks
//                          w, w+8, ... so 8 independent load->convert->store chains are in
//                          flight; fence.proxy.async + mbarrier arrive hand the stage over
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

constexpr int kEpiWarps = 4, kProdWarps = 8, kStages = kProdWarps;
constexpr int kProdThreads = kProdWarps * 32;
constexpr int kThreads = (kEpiWarps + kProdWarps + 1) * 32;   // 416
constexpr int kMmaWarp = kEpiWarps + kProdWarps;              // 12
constexpr int kTileM = 128;

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

// instruction descriptor: bf16 x bf16 -> f32, both operands K-major, M=128
__host__ __device__ constexpr uint32_t make_idesc(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) |
         ((uint32_t)(kTileM >> 4) << 24);
}
// shared-memory matrix descriptor: SWIZZLE_NONE, K-major; LBO = byte distance between the two
// 8-element K chunks of one MMA, SBO = byte distance between 8-row groups
__device__ __forceinline__ uint64_t make_sdesc(uint32_t smem_addr, uint32_t lbo_bytes) {
  return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) |
         ((uint64_t)(128u >> 4) << 32) | (1ull << 46);
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
  int k_begin, k_end;    // reduction range of this work item
};

// A Policy provides:
//   Args; N_TILE; KB; A_HAS_LO; B_RESIDENT; B_RES_K (K extent of a resident B); B_RES_SETS
//   static __device__ int num_items(const Args&)
//   static __device__ TileCoord coord(const Args&, int item)
//   static __device__ void load_A(const Args&, const TileCoord&, int k0, uint8_t* hi, uint8_t* lo, int lane)
//   static __device__ void load_B(const Args&, const TileCoord&, int k0, uint8_t* hi, uint8_t* lo, int lane)   (staged B)
//   static __device__ void load_B_resident(const Args&, uint8_t* base, int ptid)   (resident B; all 256 producer threads)
//   static __device__ int b_set(const Args&, const TileCoord&)                      (which resident B)
//   static __device__ void store(const Args&, const TileCoord&, int row, int col, const float (&v)[16])
template <class P>
struct Smem {
  using TA = OperandTile<kTileM, P::KB>;
  using TBs = OperandTile<P::N_TILE, P::KB>;
  using TBr = OperandTile<P::N_TILE, P::B_RES_K>;
  static constexpr int A_STAGE = (P::A_HAS_LO ? 2 : 1) * TA::BYTES;
  static constexpr int B_STAGE = P::B_RESIDENT ? 0 : 2 * TBs::BYTES;
  static constexpr int STAGE = A_STAGE + B_STAGE;
  static constexpr int B_RES_ONE = 2 * TBr::BYTES;
  static constexpr int B_RES = P::B_RESIDENT ? B_RES_ONE * P::B_RES_SETS : 0;
  static constexpr int BAR_OFF = kStages * STAGE + B_RES;
  static constexpr int TOTAL = BAR_OFF + 256;
};

template <class P>
__global__ void __launch_bounds__(kThreads, 1) tc_kernel(typename P::Args g) {
  using S = Smem<P>;
  using TA = typename S::TA;
  constexpr int N_TILE = P::N_TILE, KB = P::KB;
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* bres = smem + kStages * S::STAGE;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + S::BAR_OFF);
  uint64_t* empty = full + kStages;
  uint64_t* tfull = empty + kStages;     // [2]
  uint64_t* tempty = tfull + 2;          // [2]
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr uint32_t kTmemCols = 2 * N_TILE <= 32 ? 32 : 2 * N_TILE <= 64 ? 64
                               : 2 * N_TILE <= 128 ? 128 : 2 * N_TILE <= 256 ? 256 : 512;

  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full[s], 32);
      mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull[a], 1);
      mbar_init(&tempty[a], kEpiWarps * 32);
    }
    fence_mbar_init();
  }
  if (warp == kMmaWarp) tmem_alloc(tmem_base_slot, kTmemCols);
  if (P::B_RESIDENT && warp >= kEpiWarps && warp < kMmaWarp) {
    P::load_B_resident(g, bres, tid - kEpiWarps * 32);
    fence_proxy_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;
  const int items = P::num_items(g);

  if (warp >= kEpiWarps && warp < kMmaWarp) {
    // ===================== producers: warp pw owns stage pw =====================
    const int pw = warp - kEpiWarps;
    uint8_t* st = smem + pw * S::STAGE;
    int i = 0;                                         // running k-block index over all items
    for (int item = blockIdx.x; item < items; item += gridDim.x) {
      const TileCoord tc = P::coord(g, item);
      for (int k0 = tc.k_begin; k0 < tc.k_end; k0 += KB, ++i) {
        if ((i & (kStages - 1)) != pw) continue;
        mbar_wait(&empty[pw], ((i >> 3) & 1) ^ 1);
        P::load_A(g, tc, k0, st, st + TA::BYTES, lane);
        if (!P::B_RESIDENT)
          P::load_B(g, tc, k0, st + S::A_STAGE, st + S::A_STAGE + S::TBs::BYTES, lane);
        fence_proxy_async_smem();        // generic-proxy stores -> visible to the MMA (async proxy)
        mbar_arrive(&full[pw]);
      }
    }
  } else if (warp == kMmaWarp) {
    // ===================== MMA issuer (one thread) =====================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(N_TILE);
      int i = 0;
      uint32_t acc = 0, acc_phase = 0;
      for (int item = blockIdx.x; item < items; item += gridDim.x) {
        const TileCoord tc = P::coord(g, item);
        mbar_wait(&tempty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * N_TILE;
        uint32_t accumulate = 0;
        for (int k0 = tc.k_begin; k0 < tc.k_end; k0 += KB, ++i) {
          const int stage = i & (kStages - 1);
          mbar_wait(&full[stage], (i >> 3) & 1);
          tc_fence_after();
          const uint32_t a_hi = smem_u32(smem + stage * S::STAGE), a_lo = a_hi + TA::BYTES;
          uint32_t b_hi, b_lo, b_lbo;
          if (P::B_RESIDENT) {
            b_lbo = S::TBr::LBO;
            b_hi = smem_u32(bres) + P::b_set(g, tc) * S::B_RES_ONE +
                   ((k0 - P::res_k_origin(g, tc)) >> 3) * S::TBr::LBO;
            b_lo = b_hi + S::TBr::BYTES;
          } else {
            b_lbo = S::TBs::LBO;
            b_hi = a_hi + S::A_STAGE;
            b_lo = b_hi + S::TBs::BYTES;
          }
#pragma unroll
          for (int k16 = 0; k16 < KB / 16; ++k16) {
            const uint32_t ao = k16 * 2 * TA::LBO, bo = k16 * 2 * b_lbo;
            const uint64_t da_hi = make_sdesc(a_hi + ao, TA::LBO);
            const uint64_t db_hi = make_sdesc(b_hi + bo, b_lbo);
            const uint64_t db_lo = make_sdesc(b_lo + bo, b_lbo);
            umma_f16(d_tmem, da_hi, db_hi, idesc, accumulate);
            umma_f16(d_tmem, da_hi, db_lo, idesc, 1u);
            if (P::A_HAS_LO) umma_f16(d_tmem, make_sdesc(a_lo + ao, TA::LBO), db_hi, idesc, 1u);
            accumulate = 1u;
          }
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
      const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + acc * N_TILE;
#pragma unroll 1
      for (int c = 0; c < N_TILE; c += 16) {
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
  static_assert(P::N_TILE % 16 == 0 && P::N_TILE >= 16 && P::N_TILE <= 256, "UMMA N");
  static_assert(P::KB % 16 == 0, "KB");
  auto kern = tc_kernel<P>;
  static bool attr_set = false;
  if (!attr_set) {
    ARL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL));
    attr_set = true;
  }
  if (items <= 0) return ARL_OK;
  const int grid = items < num_sms() ? items : num_sms();
  kern<<<grid, kThreads, S::TOTAL, stream>>>(g);
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

template <int N_TILE_, int KB_, bool A_TRANS, bool B_TRANS, int EPI>
struct GemmPolicy {
  using Args = GemmArgs;
  static constexpr int N_TILE = N_TILE_, KB = KB_;
  static constexpr bool A_HAS_LO = true, B_RESIDENT = false;
  static constexpr int B_RES_K = KB_, B_RES_SETS = 1;
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
  static __device__ __forceinline__ void load_A(const Args& g, const TileCoord& t, int k0,
                                                uint8_t* hi, uint8_t* lo, int lane) {
    load_tile_f32<kTileM, KB, A_TRANS>(hi, lo, g.A, g.lda, t.mt * kTileM, g.M, k0, t.k_end, lane);
  }
  static __device__ __forceinline__ void load_B(const Args& g, const TileCoord& t, int k0,
                                                uint8_t* hi, uint8_t* lo, int lane) {
    load_tile_f32<N_TILE, KB, B_TRANS>(hi, lo, g.B, g.ldb, t.nt * N_TILE, g.N, k0, t.k_end, lane);
  }
  static __device__ __forceinline__ void load_B_resident(const Args&, uint8_t*, int) {}
  static __device__ __forceinline__ int b_set(const Args&, const TileCoord&) { return 0; }
  static __device__ __forceinline__ int res_k_origin(const Args&, const TileCoord&) { return 0; }
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
