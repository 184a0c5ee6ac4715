// tcgen05 GEMM building block (sm_100a): D[M,N] (fp32) = A[M,K] . B[N,K]^T with fp32 operands in
// HBM, computed on the 5th-gen tensor cores with fp32-faithful accuracy by splitting every
// operand into bf16 hi + bf16 lo in the producer warps ("bf16x3": hi*hi + hi*lo + lo*hi,
// fp32 accumulation in TMEM; relative error ~2^-16 instead of bf16's 2^-8).
//
// Warp roles in one persistent CTA per SM (416 threads):
//   warps 0-3   epilogue : tcgen05.ld (TMEM lane quadrant = warp id) -> bias/relu/mask -> HBM
//   warps 4-11  producers: LDG fp32 -> split -> st.shared into the UMMA canonical K-major,
//                          no-swizzle layout (8 rows x 16 B core matrices) -> mbarrier arrive
//   warp 12     MMA      : one elected thread issues tcgen05.mma (cta_group::1, kind::f16,
//                          M=128, N=N_TILE, K=16), tcgen05.commit frees the smem stage /
//                          publishes the accumulator; also owns tcgen05.alloc/dealloc
// Pipelines: smem full/empty (producers <-> MMA, STAGES deep) and TMEM full/empty
// (MMA <-> epilogue, 2 accumulator stages of N_TILE columns each).
//
// smem operand image per stage (bf16): element (row r, k) of a [ROWS x KB] tile sits at
//   (k/8) * LBO + r * 16 + (k%8) * 2   with  LBO = (ROWS+1)*16  (one 16-B pad: conflict-free
//   st.shared for both producer mappings), SBO = 128 (8-row groups are contiguous).
#pragma once
#include <cuda_bf16.h>

#include "common.cuh"

namespace arl {
namespace tc {

constexpr int kEpiWarps = 4, kProdWarps = 8;
constexpr int kProdThreads = kProdWarps * 32;
constexpr int kThreads = (kEpiWarps + kProdWarps + 1) * 32;   // 416
constexpr int kMmaWarp = kEpiWarps + kProdWarps;              // 12
constexpr int kTileM = 128;

enum { EPI_PLAIN = 0, EPI_BIAS_RELU = 1, EPI_MASK = 2 };

struct GemmArgs {
  const float* A;      // element (m,k): a_trans ? A[k*lda + m] : A[m*lda + k]
  const float* B;      // element (n,k): b_trans ? B[k*ldb + n] : B[n*ldb + k]
  float* D;            // D[m*ldd + n]; split-K slice z writes D + z*M*ldd
  const float* extra;  // EPI_BIAS_RELU: bias[n];  EPI_MASK: mask[m*ldd + n] (> 0 keeps)
  int M, N, K;
  int64_t lda, ldb, ldd;
  int k_chunk;         // K range per split-K slice (multiple of KB); k_splits = ceil(K/k_chunk)
  int k_splits;
  int desc_swap;       // debug: swap the LBO/SBO fields of the smem descriptors
};

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
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,"
      "%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
        "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
        "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// instruction descriptor: bf16 x bf16 -> f32, both operands K-major, M=128
__host__ __device__ constexpr uint32_t make_idesc(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) |
         ((uint32_t)(kTileM >> 4) << 24);
}
// shared-memory matrix descriptor, SWIZZLE_NONE, K-major
__device__ __forceinline__ uint64_t make_sdesc(uint32_t smem_addr, uint32_t lbo_bytes,
                                               uint32_t sbo_bytes) {
  return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) |
         ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46);
}

// fp32 -> (bf16 hi, bf16 lo) for a pair; returns packed bf16x2 words
__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  const float2 hf = __bfloat1622float2(h);
  const __nv_bfloat162 l = __floats2bfloat162_rn(a - hf.x, b - hf.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}

template <int ROWS, int KB>
struct OperandTile {
  static constexpr int KC = KB / 8;
  static constexpr int LBO = (ROWS + 1) * 16;
  static constexpr int BYTES = KC * LBO;          // one of hi / lo
};

// Fill one operand tile (hi and lo images) for rows [r0, r0+ROWS) and k in [k0, k0+KB).
// TRANS=false: src element (r,k) at src[r*ld + k]   (k contiguous)
// TRANS=true : src element (r,k) at src[k*ld + r]   (r contiguous)
template <int ROWS, int KB, bool TRANS>
__device__ __forceinline__ void produce_tile(uint8_t* hi_img, uint8_t* lo_img,
                                             const float* __restrict__ src, int64_t ld, int r0,
                                             int rmax, int k0, int kmax, int ptid) {
  using T = OperandTile<ROWS, KB>;
  if (!TRANS) {
    for (int c = ptid; c < ROWS * T::KC; c += kProdThreads) {
      const int row = c % ROWS, kc = c / ROWS;
      const int r = r0 + row, k = k0 + kc * 8;
      float4 x0 = make_float4(0.f, 0.f, 0.f, 0.f), x1 = x0;
      if (r < rmax && k < kmax) {
        const float4* p = reinterpret_cast<const float4*>(src + (int64_t)r * ld + k);
        x0 = p[0];
        x1 = p[1];
      }
      uint4 h, l;
      split2(x0.x, x0.y, h.x, l.x);
      split2(x0.z, x0.w, h.y, l.y);
      split2(x1.x, x1.y, h.z, l.z);
      split2(x1.z, x1.w, h.w, l.w);
      const int off = kc * T::LBO + row * 16;
      *reinterpret_cast<uint4*>(hi_img + off) = h;
      *reinterpret_cast<uint4*>(lo_img + off) = l;
    }
  } else {
    // 8 (rows) x 8 (k) blocks; lanes run along k-chunks so that the 8 st.shared of a quarter
    // warp land in different 16-B bank groups (LBO has a one-unit pad)
    for (int b = ptid; b < (ROWS / 8) * T::KC; b += kProdThreads) {
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
        uint4 h, l;
        split2(v[0][i], v[1][i], h.x, l.x);
        split2(v[2][i], v[3][i], h.y, l.y);
        split2(v[4][i], v[5][i], h.z, l.z);
        split2(v[6][i], v[7][i], h.w, l.w);
        const int off = kc * T::LBO + (rg * 8 + i) * 16;
        *reinterpret_cast<uint4*>(hi_img + off) = h;
        *reinterpret_cast<uint4*>(lo_img + off) = l;
      }
    }
  }
}

template <int N_TILE, int KB, int STAGES>
struct GemmSmem {
  using TA = OperandTile<kTileM, KB>;
  using TB = OperandTile<N_TILE, KB>;
  static constexpr int STAGE_BYTES = 2 * TA::BYTES + 2 * TB::BYTES;
  static constexpr int BAR_BYTES = 256;
  static constexpr int TOTAL = STAGES * STAGE_BYTES + BAR_BYTES;
};

template <int N_TILE, int KB, int STAGES, bool A_TRANS, bool B_TRANS, int EPI>
__global__ void __launch_bounds__(kThreads, 1) gemm_bf16x3_kernel(GemmArgs g) {
  using S = GemmSmem<N_TILE, KB, STAGES>;
  using TA = typename S::TA;
  using TB = typename S::TB;
  extern __shared__ __align__(128) uint8_t smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * S::STAGE_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* tfull = empty + STAGES;      // [2]
  uint64_t* tempty = tfull + 2;          // [2]
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr uint32_t kTmemCols = 2 * N_TILE <= 32 ? 32 : 2 * N_TILE <= 64 ? 64
                               : 2 * N_TILE <= 128 ? 128 : 2 * N_TILE <= 256 ? 256 : 512;

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], kProdThreads);
      mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull[a], 1);
      mbar_init(&tempty[a], kEpiWarps * 32);
    }
    fence_mbar_init();
  }
  if (warp == kMmaWarp) tmem_alloc(tmem_base_slot, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;

  const int m_tiles = (g.M + kTileM - 1) / kTileM;
  const int n_tiles = (g.N + N_TILE - 1) / N_TILE;
  const int items = m_tiles * n_tiles * g.k_splits;

  if (warp >= kEpiWarps && warp < kMmaWarp) {
    // ===================== producers =====================
    const int ptid = tid - kEpiWarps * 32;
    int stage = 0;
    uint32_t phase = 0;
    for (int item = blockIdx.x; item < items; item += gridDim.x) {
      const int ks = item % g.k_splits, nt = (item / g.k_splits) % n_tiles,
                mt = item / (g.k_splits * n_tiles);
      const int kbeg = ks * g.k_chunk, kend = min(g.K, kbeg + g.k_chunk);
      for (int k0 = kbeg; k0 < kend; k0 += KB) {
        mbar_wait(&empty[stage], phase ^ 1);
        uint8_t* st = smem + stage * S::STAGE_BYTES;
        produce_tile<kTileM, KB, A_TRANS>(st, st + TA::BYTES, g.A, g.lda, mt * kTileM, g.M, k0,
                                          kend, ptid);
        produce_tile<N_TILE, KB, B_TRANS>(st + 2 * TA::BYTES, st + 2 * TA::BYTES + TB::BYTES, g.B,
                                          g.ldb, nt * N_TILE, g.N, k0, kend, ptid);
        fence_proxy_async_smem();            // generic-proxy writes -> visible to the MMA (async proxy)
        mbar_arrive(&full[stage]);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == kMmaWarp) {
    // ===================== MMA issuer (one thread) =====================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(N_TILE);
      int stage = 0;
      uint32_t phase = 0, acc = 0, acc_phase = 0;
      for (int item = blockIdx.x; item < items; item += gridDim.x) {
        const int ks = item % g.k_splits;
        const int kbeg = ks * g.k_chunk, kend = min(g.K, kbeg + g.k_chunk);
        mbar_wait(&tempty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * N_TILE;
        uint32_t accumulate = 0;
        for (int k0 = kbeg; k0 < kend; k0 += KB) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t base = smem_u32(smem + stage * S::STAGE_BYTES);
          const uint32_t a_hi = base, a_lo = base + TA::BYTES;
          const uint32_t b_hi = base + 2 * TA::BYTES, b_lo = b_hi + TB::BYTES;
#pragma unroll
          for (int k16 = 0; k16 < KB / 16; ++k16) {
            const uint32_t ao = k16 * 2 * TA::LBO, bo = k16 * 2 * TB::LBO;
            uint64_t da_hi, da_lo, db_hi, db_lo;
            if (!g.desc_swap) {
              da_hi = make_sdesc(a_hi + ao, TA::LBO, 128);
              da_lo = make_sdesc(a_lo + ao, TA::LBO, 128);
              db_hi = make_sdesc(b_hi + bo, TB::LBO, 128);
              db_lo = make_sdesc(b_lo + bo, TB::LBO, 128);
            } else {
              da_hi = make_sdesc(a_hi + ao, 128, TA::LBO);
              da_lo = make_sdesc(a_lo + ao, 128, TA::LBO);
              db_hi = make_sdesc(b_hi + bo, 128, TB::LBO);
              db_lo = make_sdesc(b_lo + bo, 128, TB::LBO);
            }
            umma_f16(d_tmem, da_hi, db_hi, idesc, accumulate);
            umma_f16(d_tmem, da_hi, db_lo, idesc, 1u);
            umma_f16(d_tmem, da_lo, db_hi, idesc, 1u);
            accumulate = 1u;
          }
          umma_commit(&empty[stage]);        // frees the smem stage when these MMAs retire
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
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
      const int ks = item % g.k_splits, nt = (item / g.k_splits) % n_tiles,
                mt = item / (g.k_splits * n_tiles);
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      const int m = mt * kTileM + warp * 32 + lane;
      float* drow = g.D + (int64_t)ks * g.M * g.ldd + (int64_t)m * g.ldd + nt * N_TILE;
      const float* xrow = EPI == EPI_MASK ? g.extra + (int64_t)m * g.ldd + nt * N_TILE
                                          : g.extra + nt * N_TILE;
      const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + acc * N_TILE;
#pragma unroll 1
      for (int c = 0; c < N_TILE; c += 16) {
        uint32_t v[16];
        tmem_ld16(taddr + c, v);
        tmem_ld_wait();
        if (m < g.M && nt * N_TILE + c < g.N) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            float4 o = make_float4(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]),
                                   __uint_as_float(v[4 * q + 2]), __uint_as_float(v[4 * q + 3]));
            if (EPI == EPI_BIAS_RELU) {
              const float4 bb = *reinterpret_cast<const float4*>(xrow + c + 4 * q);
              o.x = fmaxf(o.x + bb.x, 0.f); o.y = fmaxf(o.y + bb.y, 0.f);
              o.z = fmaxf(o.z + bb.z, 0.f); o.w = fmaxf(o.w + bb.w, 0.f);
            } else if (EPI == EPI_MASK) {
              const float4 mm = *reinterpret_cast<const float4*>(xrow + c + 4 * q);
              o.x = mm.x > 0.f ? o.x : 0.f; o.y = mm.y > 0.f ? o.y : 0.f;
              o.z = mm.z > 0.f ? o.z : 0.f; o.w = mm.w > 0.f ? o.w : 0.f;
            }
            *reinterpret_cast<float4*>(drow + c + 4 * q) = o;
          }
        }
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

// host-side launcher
template <int N_TILE, int KB, int STAGES, bool A_TRANS, bool B_TRANS, int EPI>
int launch_gemm(const GemmArgs& g, cudaStream_t stream) {
  using S = GemmSmem<N_TILE, KB, STAGES>;
  static_assert(S::TOTAL <= 227 * 1024, "smem budget");
  static_assert(N_TILE % 16 == 0 && N_TILE >= 16 && N_TILE <= 256, "UMMA N");
  auto kern = gemm_bf16x3_kernel<N_TILE, KB, STAGES, A_TRANS, B_TRANS, EPI>;
  static bool attr_set = false;
  if (!attr_set) {
    ARL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL));
    attr_set = true;
  }
  const int m_tiles = (g.M + kTileM - 1) / kTileM, n_tiles = (g.N + N_TILE - 1) / N_TILE;
  const int items = m_tiles * n_tiles * g.k_splits;
  if (items == 0) return ARL_OK;
  const int grid = items < num_sms() ? items : num_sms();
  kern<<<grid, kThreads, S::TOTAL, stream>>>(g);
  ARL_LAUNCH_CHECK("gemm_bf16x3_kernel");
  return ARL_OK;
}

}  // namespace tc
}  // namespace arl
