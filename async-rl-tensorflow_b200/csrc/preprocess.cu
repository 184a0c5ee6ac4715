// K1: Environment.screen (src/environment.py:49-53) fused with History.add
// (src/history.py:13-15).  u8 [B,210,160,3] -> luma (the reference's float64
// expression, truncated) -> cv2 INTER_LINEAR 84x84 (fixed point) -> ring slot.
//
// Layout / data flow per frame (one persistent CTA per SM, 2-stage TMA pipeline):
//   HBM --cp.async.bulk (84 x 960 B: the 168 source rows cv2 actually reads)--> smem raw[stage]
//   raw --luma, 8 px / thread-iteration, dp4a--> smem Y [168][160] u8
//   Y   --2x2 fixed-point taps--> smem out [84*84] u8 --cp.async.bulk--> ring[b][slot]
// Algorithmic HBM bytes per frame: 80 640 read + 7 056 written (x replicate).
//
// Luma: the reference computes (0.2126*R + 0.7152*G) + 0.0722*B in float64 and truncates.
// That equals floor((2126R+7152G+722B)/10000) except on 774 of the 3384 triples whose
// exact value is an integer, where float64 rounding lands one ulp below (white -> 254).
// For a given (R,G) at most one B in 0..255 makes the sum a multiple of 10000
// (722B mod 10000 has period 5000), so a 65536-bit bitmap indexed by (G,R) says "subtract
// one when the remainder is 0".  arl_init derives that bitmap on the device with IEEE
// __dmul_rn/__dadd_rn (never contracted to FMA) -- i.e. from the reference's own expression.
#include "common.cuh"

namespace arl {

constexpr int kH = ARL_FRAME_H, kW = ARL_FRAME_W, kS = ARL_SCREEN;
constexpr int kRowBytes = kW * 3;                 // 480
constexpr int kPairBytes = 2 * kRowBytes;         // 960: source rows sy, sy+1
constexpr int kRawBytes = kS * kPairBytes;        // 80640 per frame
constexpr int kYBytes = 2 * kS * kW;              // 26880
constexpr int kFrameBytes = kH * kW * 3;          // 100800
constexpr int kBitmapWords = 65536 / 32;          // 2048
constexpr int kThreads = 512;

struct TapTables {
  uint32_t x[kS];      // sx | c0 << 8 | c1 << 20
  uint32_t y[kS];      // sy | b0 << 8 | b1 << 20
};
__constant__ TapTables c_taps;
__device__ uint32_t g_luma_fix[kBitmapWords];
static bool g_ready[64] = {};

// ---- init: derive the luma correction bitmap from the reference's float64 expression ----
__global__ void luma_fix_init_kernel() {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;   // idx = G*256 + R
  if (idx >= 65536) return;
  const int R = idx & 255, G = idx >> 8;
  for (int B = 0; B < 256; ++B) {
    const int s = 2126 * R + 7152 * G + 722 * B;
    if (s % 10000 != 0) continue;
    const double y = __dadd_rn(__dadd_rn(__dmul_rn(0.2126, (double)R), __dmul_rn(0.7152, (double)G)),
                               __dmul_rn(0.0722, (double)B));
    const int yt = (int)y;                                   // truncation, as astype(uint8)
    if (yt != s / 10000) atomicOr(&g_luma_fix[idx >> 5], 1u << (idx & 31));
  }
}

// cv2 INTER_LINEAR tap table for one axis (same float32 arithmetic as cv2's resize()).
static void linear_taps(int src, int dst, uint32_t* packed) {
  const double scale = 1.0 / ((double)dst / (double)src);
  for (int d = 0; d < dst; ++d) {
    float fx = (float)((d + 0.5) * scale - 0.5);
    int sx = (int)floorf(fx);
    fx -= (float)sx;
    if (sx < 0) { fx = 0.f; sx = 0; }
    if (sx >= src - 1) { fx = 0.f; sx = src - 1; }
    const int c0 = (int)lrintf((1.f - fx) * 2048.f);   // cvRound: half to even
    const int c1 = (int)lrintf(fx * 2048.f);
    packed[d] = (uint32_t)sx | ((uint32_t)c0 << 8) | ((uint32_t)c1 << 20);
  }
}

// ---- the kernel --------------------------------------------------------------------------
struct __align__(16) K1Smem {
  uint8_t raw[2][kRawBytes];
  uint8_t Y[kYBytes];
  uint8_t out[kPlane];
  uint32_t fix[kBitmapWords];
  uint32_t xtab[kS];
  uint32_t ytab[kS];
  uint64_t full[2];
};

__device__ __forceinline__ uint32_t luma8(uint32_t px /* R | G<<8 | B<<16 */, const uint32_t* fix) {
  const uint32_t s = __dp4a(px, 0x00D2F04Eu, 0u) + (__dp4a(px, 0x00021B08u, 0u) << 8);
  uint32_t q = __umulhi(s, 3518437209u) >> 13;               // s / 10000, exact for s <= 2 550 000
  if (s == q * 10000u) {
    const uint32_t idx = px & 0xFFFFu;                       // G*256 + R
    q -= (fix[idx >> 5] >> (idx & 31)) & 1u;
  }
  return q;
}

__global__ void __launch_bounds__(kThreads, 1)
preprocess_kernel(const uint8_t* __restrict__ frames, uint8_t* __restrict__ ring, int num_envs,
                  int ring_slots, int slot, int replicate) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  K1Smem& sm = *reinterpret_cast<K1Smem*>(smem_raw);
  const int tid = threadIdx.x;

  for (int i = tid; i < kBitmapWords; i += kThreads) sm.fix[i] = g_luma_fix[i];
  if (tid < kS) {
    sm.xtab[tid] = c_taps.x[tid];
    sm.ytab[tid] = c_taps.y[tid];
  }
  if (tid == 0) {
    mbar_init(&sm.full[0], 1);
    mbar_init(&sm.full[1], 1);
    fence_mbar_init();
  }
  __syncthreads();

  // producer: warp 0 issues the 84 row-pair copies of one frame
  auto issue = [&](int env, int stage) {
    if (tid < 32) {
      if (tid == 0) mbar_expect_tx(&sm.full[stage], kRawBytes);
      __syncwarp();
      const uint8_t* src = frames + (size_t)env * kFrameBytes;
      for (int dy = tid; dy < kS; dy += 32) {
        const uint32_t sy = sm.ytab[dy] & 0xFFu;
        bulk_g2s(&sm.raw[stage][dy * kPairBytes], src + sy * kRowBytes, kPairBytes, &sm.full[stage]);
      }
    }
  };

  int it = 0;
  if ((int)blockIdx.x < num_envs) issue(blockIdx.x, 0);
  for (int env = blockIdx.x; env < num_envs; env += gridDim.x, ++it) {
    const int stage = it & 1;
    const int next = env + gridDim.x;
    if (next < num_envs) issue(next, stage ^ 1);             // stage^1 was drained last iteration
    mbar_wait(&sm.full[stage], (it >> 1) & 1);

    // phase 1: luma of the 168 x 160 staged pixels, 8 pixels (24 B) per thread-iteration
    const uint2* raw2 = reinterpret_cast<const uint2*>(sm.raw[stage]);
    for (int g = tid; g < kYBytes / 8; g += kThreads) {
      const uint2 a = raw2[3 * g], b = raw2[3 * g + 1], c = raw2[3 * g + 2];
      const uint32_t w0 = a.x, w1 = a.y, w2 = b.x, w3 = b.y, w4 = c.x, w5 = c.y;
      uint32_t y0, y1;
      y0 = luma8(w0 & 0x00FFFFFFu, sm.fix);
      y0 |= luma8(__byte_perm(w0, w1, 0x4543), sm.fix) << 8;    // bytes 3,4,5
      y0 |= luma8(__byte_perm(w1, w2, 0x4432), sm.fix) << 16;   // bytes 6,7,8   (w1.2,w1.3,w2.0)
      y0 |= luma8(w2 >> 8, sm.fix) << 24;                       // bytes 9,10,11
      y1 = luma8(w3 & 0x00FFFFFFu, sm.fix);                     // bytes 12,13,14
      y1 |= luma8(__byte_perm(w3, w4, 0x4543), sm.fix) << 8;    // bytes 15,16,17
      y1 |= luma8(__byte_perm(w4, w5, 0x4432), sm.fix) << 16;   // bytes 18,19,20
      y1 |= luma8(w5 >> 8, sm.fix) << 24;                       // bytes 21,22,23
      reinterpret_cast<uint2*>(sm.Y)[g] = make_uint2(y0, y1);
    }
    if (tid == 0) bulk_wait_read<0>();                          // previous store has read sm.out
    __syncthreads();

    // phase 2: cv2 fixed-point bilinear, one output pixel per thread-iteration
    for (int idx = tid; idx < kPlane; idx += kThreads) {
      const int dy = idx / kS, dx = idx - dy * kS;
      const uint32_t xt = sm.xtab[dx], yt = sm.ytab[dy];
      const int sx = xt & 0xFF, c0 = (xt >> 8) & 0xFFF, c1 = xt >> 20;
      const int b0 = (yt >> 8) & 0xFFF, b1 = yt >> 20;
      const uint8_t* r0 = sm.Y + (2 * dy) * kW + sx;
      const int h0 = r0[0] * c0 + r0[1] * c1;
      const int h1 = r0[kW] * c0 + r0[kW + 1] * c1;
      const int v = (((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2;
      sm.out[idx] = (uint8_t)v;
    }
    fence_proxy_async_smem();
    __syncthreads();
    if (tid == 0) {
      uint8_t* dst = ring + ((size_t)env * ring_slots) * kPlane;
      for (int r = 0; r < replicate; ++r) {
        int s = slot + r;
        if (s >= ring_slots) s -= ring_slots;
        bulk_s2g(dst + (size_t)s * kPlane, sm.out, kPlane);
      }
      bulk_commit();
    }
  }
  if (tid == 0) bulk_wait<0>();
}

// ---- History.get / reset -----------------------------------------------------------------
template <typename OutT>
__global__ void history_get_kernel(const uint8_t* __restrict__ ring, OutT* __restrict__ out,
                                   int num_envs, int ring_slots, int first_slot) {
  const int64_t total = (int64_t)num_envs * kPlane;            // one thread per (env, pixel)
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int env = (int)(i / kPlane);
    const int px = (int)(i - (int64_t)env * kPlane);
    const uint8_t* base = ring + (size_t)env * ring_slots * kPlane + px;
    OutT v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      int s = first_slot + k;
      if (s >= ring_slots) s -= ring_slots;
      v[k] = (OutT)base[(size_t)s * kPlane];
    }
    OutT* o = out + i * 4;
#pragma unroll
    for (int k = 0; k < 4; ++k) o[k] = v[k];
  }
}

int preprocess_init(int device) {
  if (device < 0 || device >= 64) {
    set_error("arl_init: device %d out of range", device);
    return ARL_ERR_INVALID;
  }
  if (g_ready[device]) return ARL_OK;
  TapTables t;
  linear_taps(kW, kS, t.x);
  linear_taps(kH, kS, t.y);
  for (int d = 0; d < kS; ++d) {
    // the staged-row scheme needs both taps in range and adjacent (true for 210x160 -> 84x84)
    if ((int)(t.y[d] & 0xFF) + 1 >= kH || (int)(t.x[d] & 0xFF) + 1 >= kW) {
      set_error("arl_init: tap table needs edge clamping, unsupported geometry");
      return ARL_ERR_UNSUPPORTED;
    }
  }
  ARL_CUDA(cudaMemcpyToSymbol(c_taps, &t, sizeof(t)));
  void* fix = nullptr;
  ARL_CUDA(cudaGetSymbolAddress(&fix, g_luma_fix));
  ARL_CUDA(cudaMemset(fix, 0, sizeof(uint32_t) * kBitmapWords));
  luma_fix_init_kernel<<<65536 / 256, 256>>>();
  ARL_LAUNCH_CHECK("luma_fix_init_kernel");
  ARL_CUDA(cudaFuncSetAttribute(preprocess_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)sizeof(K1Smem)));
  ARL_CUDA(cudaDeviceSynchronize());
  g_ready[device] = true;
  return ARL_OK;
}

}  // namespace arl

using namespace arl;

extern "C" int arl_preprocess_push(const uint8_t* frames, uint8_t* ring, int num_envs,
                                   int ring_slots, int slot, int replicate, void* stream) {
  ARL_REQUIRE(frames && ring, "arl_preprocess_push: null pointer");
  ARL_REQUIRE(num_envs >= 0, "arl_preprocess_push: num_envs %d < 0", num_envs);
  ARL_REQUIRE(ring_slots >= ARL_HISTORY, "arl_preprocess_push: ring_slots %d < %d", ring_slots,
              ARL_HISTORY);
  ARL_REQUIRE(slot >= 0 && slot < ring_slots, "arl_preprocess_push: slot %d outside [0,%d)", slot,
              ring_slots);
  ARL_REQUIRE(replicate >= 1 && replicate <= ring_slots,
              "arl_preprocess_push: replicate %d outside [1,%d]", replicate, ring_slots);
  ARL_REQUIRE((reinterpret_cast<uintptr_t>(frames) & 15) == 0 &&
                  (reinterpret_cast<uintptr_t>(ring) & 15) == 0,
              "arl_preprocess_push: frames and ring must be 16-byte aligned");
  if (num_envs == 0) return ARL_OK;
  const int grid = num_envs < num_sms() ? num_envs : num_sms();
  preprocess_kernel<<<grid, kThreads, sizeof(K1Smem), (cudaStream_t)stream>>>(
      frames, ring, num_envs, ring_slots, slot, replicate);
  ARL_LAUNCH_CHECK("preprocess_kernel");
  return ARL_OK;
}

extern "C" int arl_history_get(const uint8_t* ring, void* out, int out_is_u8, int num_envs,
                               int ring_slots, int first_slot, void* stream) {
  ARL_REQUIRE(ring && out, "arl_history_get: null pointer");
  ARL_REQUIRE(num_envs >= 0 && ring_slots >= ARL_HISTORY && first_slot >= 0 &&
                  first_slot < ring_slots,
              "arl_history_get: bad geometry (envs %d, slots %d, first %d)", num_envs, ring_slots,
              first_slot);
  if (num_envs == 0) return ARL_OK;
  const int64_t total = (int64_t)num_envs * kPlane;
  const int block = 256;
  int64_t grid64 = (total + block - 1) / block;
  const int grid = (int)(grid64 < (int64_t)num_sms() * 16 ? grid64 : (int64_t)num_sms() * 16);
  if (out_is_u8)
    history_get_kernel<uint8_t><<<grid, block, 0, (cudaStream_t)stream>>>(
        ring, (uint8_t*)out, num_envs, ring_slots, first_slot);
  else
    history_get_kernel<float><<<grid, block, 0, (cudaStream_t)stream>>>(
        ring, (float*)out, num_envs, ring_slots, first_slot);
  ARL_LAUNCH_CHECK("history_get_kernel");
  return ARL_OK;
}

extern "C" int arl_history_reset(uint8_t* ring, int num_envs, int ring_slots, void* stream) {
  ARL_REQUIRE(ring, "arl_history_reset: null pointer");
  ARL_REQUIRE(num_envs >= 0 && ring_slots >= ARL_HISTORY, "arl_history_reset: bad geometry");
  ARL_CUDA(cudaMemsetAsync(ring, 0, (size_t)num_envs * ring_slots * kPlane, (cudaStream_t)stream));
  return ARL_OK;
}
