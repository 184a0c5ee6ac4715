// K2: the two narrow-N convolutions of the trunk, forward and backward, as smem-tiled
// CUDA-core (FFMA) kernels.
//   conv1: u8 stack [84,84,4]/255 -> 8x8 s4 VALID, 16 ch, +b, relu -> a1 [20,20,16]   agent.py:226-227
//   conv2: a1 -> 4x4 s2 VALID, 32 ch, +b, relu -> a2 [9,9,32] (NHWC flatten)          agent.py:228-232
// Weight layouts are TF's [kh,kw,cin,cout] (ops.py:19-22).  Cross-correlation.
//
// Backward (agent.py:317): conv1 needs only the weight gradient (its input is data); conv2
// needs weight + input gradients.  Weight gradients reduce over all samples: every CTA keeps
// its partial sums in registers, writes one partial tensor to the workspace and a second,
// deterministic kernel (reduce_partials) sums the partials in a fixed order.
#include "common.cuh"

namespace arl {

// =============================== conv1 weight gradient =====================================
// dW1[kh][kw][c][co] = sum_{n,oy,ox} x[n][c][4oy+kh][4ox+kw]/255 * dy1[n][oy][ox][co]
// (dy1 = gradient w.r.t. the conv1 pre-activation, i.e. already relu-masked).
// CTA = 8 compute warps + 1 bias warp.  Compute thread = (q 0..7, kh 0..7, c 0..3) holds
// acc[8 kw][16 co]; it walks the pixels p = q, q+8, ... of each sample.  Per pixel:
// 2 x LDS.32 (8 input bytes) + 4 x LDS.128 (16 dy, warp-broadcast) for 128 FFMA.
// Samples are double-buffered with TMA bulk copies (4 planes + 25.6 KB of dy1).
constexpr int kW1Threads = 288;
struct __align__(16) Conv1WgradSmem {
  union {
    struct {
      uint8_t planes[2][4][kPlane];          // 56 448 B
      float dy[2][ARL_A1_ELEMS];             // 51 200 B
    } in;
    float red[8][4096];                      // 131 072 B, reused for the cross-q reduction
  } u;
  uint64_t full[2];
};

__global__ void __launch_bounds__(kW1Threads, 1)
conv1_wgrad_kernel(const uint8_t* __restrict__ ring, const float* __restrict__ dy1,
                   float* __restrict__ partials, int num_envs, int ring_slots, int first_slot,
                   int64_t num_samples) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  Conv1WgradSmem& sm = *reinterpret_cast<Conv1WgradSmem*>(smem_raw);
  const int tid = threadIdx.x;
  if (tid == 0) {
    mbar_init(&sm.full[0], 1);
    mbar_init(&sm.full[1], 1);
    fence_mbar_init();
  }
  __syncthreads();

  auto issue = [&](int64_t n, int stage) {
    const int t = (int)(n / num_envs), b = (int)(n - (int64_t)t * num_envs);
    mbar_expect_tx(&sm.full[stage], 4 * kPlane + ARL_A1_ELEMS * 4);
    for (int k = 0; k < 4; ++k) {
      const int slot = (first_slot + t + k) % ring_slots;
      bulk_g2s(sm.u.in.planes[stage][k], ring + ((size_t)b * ring_slots + slot) * kPlane, kPlane,
               &sm.full[stage]);
    }
    bulk_g2s(sm.u.in.dy[stage], dy1 + n * ARL_A1_ELEMS, ARL_A1_ELEMS * 4, &sm.full[stage]);
  };

  const bool compute = tid < 256;
  const int q = tid >> 5, kh = (tid >> 2) & 7, c = tid & 3;
  float acc[8][16];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 16; ++j) acc[i][j] = 0.f;
  float bsum = 0.f;                               // bias warp: lane -> (co = lane&15, half = lane>>4)

  int it = 0;
  if (tid == 0 && (int64_t)blockIdx.x < num_samples) issue(blockIdx.x, 0);
  for (int64_t n = blockIdx.x; n < num_samples; n += gridDim.x, ++it) {
    const int stage = it & 1;
    const int64_t nn = n + gridDim.x;
    if (tid == 0 && nn < num_samples) issue(nn, stage ^ 1);
    mbar_wait(&sm.full[stage], (it >> 1) & 1);

    if (compute) {
      const uint8_t* plane = sm.u.in.planes[stage][c];
      const float* dyp = sm.u.in.dy[stage];
#pragma unroll 2
      for (int p = q; p < 400; p += 8) {
        const int oy = p / 20, ox = p - oy * 20;
        const uint32_t* xw =
            reinterpret_cast<const uint32_t*>(plane + (4 * oy + kh) * ARL_SCREEN + 4 * ox);
        const uint32_t x0 = xw[0], x1 = xw[1];
        const float4* d4 = reinterpret_cast<const float4*>(dyp + p * 16);
        const float4 d0 = d4[0], d1 = d4[1], d2 = d4[2], d3 = d4[3];
        const float d[16] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w,
                             d2.x, d2.y, d2.z, d2.w, d3.x, d3.y, d3.z, d3.w};
#pragma unroll
        for (int kw = 0; kw < 8; ++kw) {
          const uint32_t wsel = kw < 4 ? x0 : x1;
          const float x = (float)((wsel >> (8 * (kw & 3))) & 0xFFu);
#pragma unroll
          for (int j = 0; j < 16; ++j) acc[kw][j] = fmaf(x, d[j], acc[kw][j]);
        }
      }
    } else {
      const int lane = tid - 256;
      const float* dyp = sm.u.in.dy[stage] + (lane & 15);
      for (int p = lane >> 4; p < 400; p += 2) bsum += dyp[p * 16];
    }
    __syncthreads();   // stage fully consumed before it is refilled two iterations later
  }

  // cross-q reduction through smem (the staging buffers are dead now)
  __syncthreads();
  if (compute) {
#pragma unroll
    for (int kw = 0; kw < 8; ++kw)
#pragma unroll
      for (int j = 0; j < 16; ++j) sm.u.red[q][((kh * 8 + kw) * 4 + c) * 16 + j] = acc[kw][j];
  }
  __syncthreads();
  float* out = partials + (size_t)blockIdx.x * (4096 + 16);
  const float inv255 = 1.0f / 255.0f;
  if (compute) {
    for (int i = tid; i < 4096; i += 256) {
      float v = 0.f;
#pragma unroll
      for (int qq = 0; qq < 8; ++qq) v += sm.u.red[qq][i];
      out[i] = v * inv255;
    }
  } else {
    const int lane = tid - 256;
    bsum += __shfl_xor_sync(0xffffffffu, bsum, 16);
    if (lane < 16) out[4096 + lane] = bsum;
  }
}

// =============================== conv2 weight gradient =====================================
// dW2[kh][kw][c][co] = sum_{n,oy,ox} a1[n][2oy+kh][2ox+kw][c] * dy2[n][oy][ox][co]
// Compute thread = (q 0..3, kh, kw, c-half, co-half): acc[8 c][16 co]; pixels p = q, q+4, ...
// Per pixel 2 + 4 LDS.128 for 128 FFMA.  + 1 bias warp.  Samples double-buffered by TMA.
constexpr int kW2Threads = 288;
struct __align__(16) Conv2WgradSmem {
  union {
    struct {
      float a[2][ARL_A1_ELEMS];              // 51 200 B
      float dy[2][ARL_A2_ELEMS];             // 20 736 B
    } in;
    float red[4][8192];                      // 131 072 B
  } u;
  uint64_t full[2];
};

__global__ void __launch_bounds__(kW2Threads, 1)
conv2_wgrad_kernel(const float* __restrict__ a1, const float* __restrict__ dy2,
                   float* __restrict__ partials, int64_t num_samples) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  Conv2WgradSmem& sm = *reinterpret_cast<Conv2WgradSmem*>(smem_raw);
  const int tid = threadIdx.x;
  if (tid == 0) {
    mbar_init(&sm.full[0], 1);
    mbar_init(&sm.full[1], 1);
    fence_mbar_init();
  }
  __syncthreads();

  auto issue = [&](int64_t n, int stage) {
    mbar_expect_tx(&sm.full[stage], (ARL_A1_ELEMS + ARL_A2_ELEMS) * 4);
    bulk_g2s(sm.u.in.a[stage], a1 + n * ARL_A1_ELEMS, ARL_A1_ELEMS * 4, &sm.full[stage]);
    bulk_g2s(sm.u.in.dy[stage], dy2 + n * ARL_A2_ELEMS, ARL_A2_ELEMS * 4, &sm.full[stage]);
  };

  const bool compute = tid < 256;
  const int q = tid >> 6, kh = (tid >> 4) & 3, kw = (tid >> 2) & 3, ch = (tid >> 1) & 1,
            coh = tid & 1;
  float acc[8][16];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 16; ++j) acc[i][j] = 0.f;
  float bsum = 0.f;                               // bias warp: lane = co

  int it = 0;
  if (tid == 0 && (int64_t)blockIdx.x < num_samples) issue(blockIdx.x, 0);
  for (int64_t n = blockIdx.x; n < num_samples; n += gridDim.x, ++it) {
    const int stage = it & 1;
    const int64_t nn = n + gridDim.x;
    if (tid == 0 && nn < num_samples) issue(nn, stage ^ 1);
    mbar_wait(&sm.full[stage], (it >> 1) & 1);

    if (compute) {
      const float* ap = sm.u.in.a[stage];
      const float* dyp = sm.u.in.dy[stage];
#pragma unroll 1
      for (int p = q; p < 81; p += 4) {
        const int oy = p / 9, ox = p - oy * 9;
        const float4* a4 = reinterpret_cast<const float4*>(
            ap + ((2 * oy + kh) * 20 + (2 * ox + kw)) * 16 + ch * 8);
        const float4 x0 = a4[0], x1 = a4[1];
        const float x[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
        const float4* d4 = reinterpret_cast<const float4*>(dyp + p * 32 + coh * 16);
        const float4 d0 = d4[0], d1 = d4[1], d2 = d4[2], d3 = d4[3];
        const float d[16] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w,
                             d2.x, d2.y, d2.z, d2.w, d3.x, d3.y, d3.z, d3.w};
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < 16; ++j) acc[i][j] = fmaf(x[i], d[j], acc[i][j]);
      }
    } else {
      const int lane = tid - 256;
      const float* dyp = sm.u.in.dy[stage] + lane;
      for (int p = 0; p < 81; ++p) bsum += dyp[p * 32];
    }
    __syncthreads();
  }

  __syncthreads();
  if (compute) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 16; ++j)
        sm.u.red[q][((kh * 4 + kw) * 16 + ch * 8 + i) * 32 + coh * 16 + j] = acc[i][j];
  }
  __syncthreads();
  float* out = partials + (size_t)blockIdx.x * (8192 + 32);
  if (compute) {
    for (int i = tid; i < 8192; i += 256)
      out[i] = (sm.u.red[0][i] + sm.u.red[1][i]) + (sm.u.red[2][i] + sm.u.red[3][i]);
  } else {
    out[8192 + (tid - 256)] = bsum;
  }
}

// =============================== conv2 input gradient ======================================
// d_a1[y][x][c] = sum_{kh,kw,co: y=2oy+kh, x=2ox+kw} dy2[oy][ox][co] * W2[kh][kw][c][co],
// then masked by a1 > 0 (relu of conv1) -> dy1.  Output pixels are split into the 4 parity
// classes (y&1, x&1): all pixels of a class use the same 4 (kh,kw) taps.
// CTA = 256 threads, 3 samples per tile (80 threads each: class x yy 0..9 x half-row),
// thread tile 5 px x 16 c.  smem: W2 transposed to [kh][kw][co][c] + dy2 of the tile.
constexpr int kD2Samples = 3;
constexpr int kD2Threads = 256;
struct __align__(16) Conv2DgradSmem {
  float wt[4 * 4 * 32 * 16];                 // 32 KB [kh][kw][co][c]
  float dy[kD2Samples][ARL_A2_ELEMS + 32];   // zero row appended (index 81) for out-of-range taps
};

__global__ void __launch_bounds__(kD2Threads, 2)
conv2_dgrad_kernel(const float* __restrict__ params, const float* __restrict__ a1,
                   const float* __restrict__ dy2, float* __restrict__ dy1, int64_t num_samples) {
  extern __shared__ __align__(128) uint8_t smem_raw[];
  Conv2DgradSmem& sm = *reinterpret_cast<Conv2DgradSmem*>(smem_raw);
  const int tid = threadIdx.x;
  const float* w2 = params + 4096 + 16;
  for (int i = tid; i < 8192; i += kD2Threads) {
    const int co = i & 31, c = (i >> 5) & 15, k = i >> 9;      // source index [k][c][co]
    sm.wt[(k * 32 + co) * 16 + c] = w2[i];
  }
  if (tid < kD2Samples * 32) sm.dy[tid >> 5][ARL_A2_ELEMS + (tid & 31)] = 0.f;

  const int s = tid / 80, r = tid - s * 80;
  const int cls = r / 20, rr = r - cls * 20;
  const int py = cls >> 1, px = cls & 1, yy = rr >> 1, xh = rr & 1;
  const bool active = tid < 240;
  const int64_t num_tiles = (num_samples + kD2Samples - 1) / kD2Samples;

  for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
    const int64_t n0 = tile * kD2Samples;
    const int ns = (int)((num_samples - n0) < kD2Samples ? (num_samples - n0) : kD2Samples);
    __syncthreads();
    const float4* src = reinterpret_cast<const float4*>(dy2 + n0 * ARL_A2_ELEMS);
    for (int i = tid; i < ns * (ARL_A2_ELEMS / 4); i += kD2Threads) {
      const int j = i / 648, e = (i - j * 648) * 4;
      *reinterpret_cast<float4*>(&sm.dy[j][e]) = src[i];
    }
    __syncthreads();

    if (active && s < ns) {
      float acc[5][16];
#pragma unroll
      for (int j = 0; j < 5; ++j)
#pragma unroll
        for (int c = 0; c < 16; ++c) acc[j][c] = 0.f;

#pragma unroll 1
      for (int tap = 0; tap < 4; ++tap) {
        const int dkh = tap >> 1, dkw = tap & 1;
        const int kh = py + 2 * dkh, kw = px + 2 * dkw;
        const int oy = yy - dkh;
        // pixel j of this thread: xx = 5*xh + j, ox = xx - dkw
        int pidx[5];
#pragma unroll
        for (int j = 0; j < 5; ++j) {
          const int ox = 5 * xh + j - dkw;
          pidx[j] = (oy >= 0 && oy <= 8 && ox >= 0 && ox <= 8) ? (oy * 9 + ox) : 81;
        }
        const float* wp = &sm.wt[(kh * 4 + kw) * 512];
        const float* dyp = sm.dy[s];
#pragma unroll 4
        for (int co = 0; co < 32; ++co) {
          float d[5];
#pragma unroll
          for (int j = 0; j < 5; ++j) d[j] = dyp[pidx[j] * 32 + co];
          const float4* w4 = reinterpret_cast<const float4*>(wp + co * 16);
          const float4 w0 = w4[0], w1 = w4[1], w2v = w4[2], w3 = w4[3];
          const float w[16] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w,
                               w2v.x, w2v.y, w2v.z, w2v.w, w3.x, w3.y, w3.z, w3.w};
#pragma unroll
          for (int j = 0; j < 5; ++j)
#pragma unroll
            for (int c = 0; c < 16; ++c) acc[j][c] = fmaf(d[j], w[c], acc[j][c]);
        }
      }
      const int y = 2 * yy + py;
#pragma unroll
      for (int j = 0; j < 5; ++j) {
        const int x = 2 * (5 * xh + j) + px;
        const int64_t off = (n0 + s) * ARL_A1_ELEMS + (y * 20 + x) * 16;
        const float4* m4 = reinterpret_cast<const float4*>(a1 + off);
        float4* o4 = reinterpret_cast<float4*>(dy1 + off);
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          const float4 m = m4[v];
          o4[v] = make_float4(m.x > 0.f ? acc[j][4 * v] : 0.f, m.y > 0.f ? acc[j][4 * v + 1] : 0.f,
                              m.z > 0.f ? acc[j][4 * v + 2] : 0.f,
                              m.w > 0.f ? acc[j][4 * v + 3] : 0.f);
        }
      }
    }
  }
}

// =============================== deterministic partial reduction ===========================
// out[i] = sum_p partials[p][i], p in fixed order.  One thread per output element.
__global__ void reduce_partials_kernel(const float* __restrict__ partials, float* __restrict__ out,
                                       int num_partials, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  int p = 0;
  for (; p + 4 <= num_partials; p += 4) {
    a0 += partials[(size_t)p * n + i];
    a1 += partials[(size_t)(p + 1) * n + i];
    a2 += partials[(size_t)(p + 2) * n + i];
    a3 += partials[(size_t)(p + 3) * n + i];
  }
  for (; p < num_partials; ++p) a0 += partials[(size_t)p * n + i];
  out[i] = (a0 + a1) + (a2 + a3);
}

int reduce_partials(const float* partials, float* out, int num_partials, int n,
                    cudaStream_t stream) {
  reduce_partials_kernel<<<(n + 255) / 256, 256, 0, stream>>>(partials, out, num_partials, n);
  ARL_LAUNCH_CHECK("reduce_partials_kernel");
  return ARL_OK;
}

int conv_init() {
  ARL_CUDA(cudaFuncSetAttribute(conv1_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)sizeof(Conv1WgradSmem)));
  ARL_CUDA(cudaFuncSetAttribute(conv2_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)sizeof(Conv2WgradSmem)));
  ARL_CUDA(cudaFuncSetAttribute(conv2_dgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)sizeof(Conv2DgradSmem)));
  return ARL_OK;
}

int wgrad_grid() { return num_sms(); }

}  // namespace arl

using namespace arl;

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

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
  const int grid = (int)(N < wgrad_grid() ? N : wgrad_grid());
  conv1_wgrad_kernel<<<grid, kW1Threads, sizeof(Conv1WgradSmem), st>>>(
      ring, d_a1, (float*)workspace, num_envs, ring_slots, first_slot, N);
  ARL_LAUNCH_CHECK("conv1_wgrad_kernel");
  return reduce_partials((const float*)workspace, grads, grid, 4096 + 16, st);   // l1_w | l1_b
}

extern "C" int arl_conv2_backward(const float* params, const float* a1, const float* d_a2,
                                  float* d_a1, float* grads, void* workspace, int64_t num_samples,
                                  void* stream) {
  ARL_REQUIRE(params && a1 && d_a2 && d_a1 && grads && workspace,
              "arl_conv2_backward: null pointer");
  ARL_REQUIRE(num_samples >= 0, "arl_conv2_backward: negative size");
  ARL_REQUIRE(aligned16(params) && aligned16(a1) && aligned16(d_a2) && aligned16(d_a1) &&
                  aligned16(grads) && aligned16(workspace),
              "arl_conv2_backward: pointers must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  float* g2 = grads + 4096 + 16;                                                 // l2_w | l2_b
  if (num_samples == 0) {
    ARL_CUDA(cudaMemsetAsync(g2, 0, (8192 + 32) * sizeof(float), st));
    return ARL_OK;
  }
  const int grid = (int)(num_samples < wgrad_grid() ? num_samples : wgrad_grid());
  conv2_wgrad_kernel<<<grid, kW2Threads, sizeof(Conv2WgradSmem), st>>>(a1, d_a2, (float*)workspace,
                                                                        num_samples);
  ARL_LAUNCH_CHECK("conv2_wgrad_kernel");
  int rc = reduce_partials((const float*)workspace, g2, grid, 8192 + 32, st);
  if (rc) return rc;
  const int64_t tiles = (num_samples + kD2Samples - 1) / kD2Samples;
  const int dgrid = (int)(tiles < 2LL * num_sms() ? tiles : 2LL * num_sms());
  conv2_dgrad_kernel<<<dgrid, kD2Threads, sizeof(Conv2DgradSmem), st>>>(params, a1, d_a2, d_a1,
                                                                         num_samples);
  ARL_LAUNCH_CHECK("conv2_dgrad_kernel");
  return ARL_OK;
}
