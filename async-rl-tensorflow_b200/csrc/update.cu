// K5: per-tensor clip_by_norm (agent.py:316-319) fused with the shared RMSProp apply
// (agent.py:321; optimizer built at main.py:63-65: decay .99, momentum 0, epsilon .1; slot
// `rms` starts at 1.0 in TF).  Two launches over the flat buffers:
//   1. sumsq_kernel   : one CTA per (tensor, 4096-element chunk) -> partial sum of squares
//   2. rmsprop_kernel : same grid; each CTA first adds up its tensor's partials in a fixed
//                       order (deterministic), derives scale = clip / max(norm, clip), then
//                       ms += (g^2 - ms)(1 - decay);  w -= lr * g / sqrt(ms + eps)
// 20 B per parameter per update (read g, ms, w; write ms, w) + 4 B for the norm pass.
#include "common.cuh"

namespace arl {

int num_sms();
constexpr int kChunk = 4096;
constexpr int kUpThreads = 256;

constexpr int kMaxTensors = 16;           // nips: 10 tensors, nature: 12
struct UpdatePlan {
  int n;                                  // tensors
  int64_t off[kMaxTensors + 1];
  int chunk_begin[kMaxTensors + 1];       // first chunk index of each tensor
};

__device__ __forceinline__ int find_tensor(const UpdatePlan& p, int chunk) {
  int t = 0;
#pragma unroll
  for (int i = 1; i < kMaxTensors; ++i) t += (i < p.n && chunk >= p.chunk_begin[i]) ? 1 : 0;
  return t;
}

__global__ void __launch_bounds__(kUpThreads)
sumsq_kernel(const float* __restrict__ grads, float* __restrict__ partial, UpdatePlan plan) {
  __shared__ float red[kUpThreads / 32];
  const int t = find_tensor(plan, blockIdx.x);
  const int64_t beg = plan.off[t] + (int64_t)(blockIdx.x - plan.chunk_begin[t]) * kChunk;
  const int64_t end = beg + kChunk < plan.off[t + 1] ? beg + kChunk : plan.off[t + 1];
  float s = 0.f;
  pdl_wait();
  for (int64_t i = beg + threadIdx.x; i < end; i += kUpThreads) {
    const float g = grads[i];
    s = fmaf(g, g, s);
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float v = 0.f;
#pragma unroll
    for (int i = 0; i < kUpThreads / 32; ++i) v += red[i];
    partial[blockIdx.x] = v;
  }
}

// ---- all-reduce over NVLink peer memory fused into the norm pass (comm.cu: arl_comm_enable_p2p) ----
// publish: copy this rank's gradient into slot (cycle+1)&1 of its own shared buffer; the LAST block to
// finish makes the copy visible system-wide, advances the cycle counter and writes the new cycle
// number into word [rank] of every rank's flag array (remote stores over NVLink).
__global__ void __launch_bounds__(256)
p2p_publish_kernel(const float* __restrict__ grads, P2PView v, long long n) {
  const unsigned long long next = *v.cycle + 1;                 // (only the last block changes it, below)
  float* dst = v.slot[v.rank] + (next & 1) * v.stride;
  const long long n4 = n >> 2;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x)
    reinterpret_cast<float4*>(dst)[i] = reinterpret_cast<const float4*>(grads)[i];
  if (blockIdx.x == 0)
    for (long long i = (n4 << 2) + threadIdx.x; i < n; i += blockDim.x) dst[i] = grads[i];
  __threadfence_system();
  __syncthreads();
  __shared__ bool last;
  if (threadIdx.x == 0) last = atomicAdd(v.done, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!last) return;
  if (threadIdx.x == 0) {
    *v.done = 0u;
    *v.cycle = next;
  }
  __threadfence_system();
  if (threadIdx.x < v.nranks)
    *reinterpret_cast<volatile unsigned long long*>(v.flags[threadIdx.x] + v.rank) = next;
}

// the norm pass of the update with the exchange inside: wait until every rank's flag has reached
// this cycle, then g[i] = sum over ranks (in rank order: the same bits on every rank) of their
// published gradient, written locally, + the partial sum of squares of the chunk
template <int NR>
__global__ void __launch_bounds__(kUpThreads)
p2p_reduce_sumsq_kernel(float* __restrict__ grads, float* __restrict__ partial, UpdatePlan plan, P2PView v) {
  __shared__ float red[kUpThreads / 32];
  const unsigned long long c = *v.cycle;
  if (threadIdx.x < v.nranks) {
    const volatile unsigned long long* f = v.flags[v.rank] + threadIdx.x;
    const long long t0 = clock64();
    while (*f < c) {
      if (clock64() - t0 > 6000000000LL) {                      // ~3 s: a peer is gone; do not hang the GPU
        *v.error = 1;
        break;
      }
      __nanosleep(200);
    }
  }
  __threadfence_system();
  __syncthreads();
  const int t = find_tensor(plan, blockIdx.x);
  const int64_t beg = plan.off[t] + (int64_t)(blockIdx.x - plan.chunk_begin[t]) * kChunk;
  const int64_t end = beg + kChunk < plan.off[t + 1] ? beg + kChunk : plan.off[t + 1];
  const long long so = (long long)(c & 1) * v.stride;
  float s = 0.f;
  // A remote load takes a few microseconds: every thread requests ALL ranks' values of its
  // elements (NR x U vectors in flight) before it adds the first one -- one latency per round, two
  // rounds per 4096-element chunk -- instead of one dependent round per element.
  constexpr int U = 2;
  if (((beg | end) & 3) == 0) {
    for (int64_t i0 = beg + 4 * threadIdx.x; i0 < end; i0 += 4 * kUpThreads * U) {
      float4 x[NR][U];
#pragma unroll
      for (int r = 0; r < NR; ++r)
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int64_t i = i0 + (int64_t)u * 4 * kUpThreads;
          x[r][u] = i < end ? __ldcg(reinterpret_cast<const float4*>(v.slot[r] + so + i)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int64_t i = i0 + (int64_t)u * 4 * kUpThreads;
        if (i >= end) break;
        float4 g = x[0][u];
#pragma unroll
        for (int r = 1; r < NR; ++r) { g.x += x[r][u].x; g.y += x[r][u].y; g.z += x[r][u].z; g.w += x[r][u].w; }
        *reinterpret_cast<float4*>(grads + i) = g;
        s = fmaf(g.x, g.x, fmaf(g.y, g.y, fmaf(g.z, g.z, fmaf(g.w, g.w, s))));
      }
    }
  } else {                                          // small tensors whose offsets are not multiples of 4
    for (int64_t i0 = beg + threadIdx.x; i0 < end; i0 += kUpThreads * 4) {
      float x[NR][4];
#pragma unroll
      for (int r = 0; r < NR; ++r)
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int64_t i = i0 + (int64_t)u * kUpThreads;
          x[r][u] = i < end ? __ldcg(v.slot[r] + so + i) : 0.f;
        }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int64_t i = i0 + (int64_t)u * kUpThreads;
        if (i >= end) break;
        float g = x[0][u];
#pragma unroll
        for (int r = 1; r < NR; ++r) g += x[r][u];
        grads[i] = g;
        s = fmaf(g, g, s);
      }
    }
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < kUpThreads / 32; ++i) acc += red[i];
    partial[blockIdx.x] = acc;
  }
}

__global__ void __launch_bounds__(kUpThreads)
rmsprop_kernel(float* __restrict__ params, float* __restrict__ rms, const float* __restrict__ grads,
               const float* __restrict__ partial, float* __restrict__ norms_out, UpdatePlan plan,
               float lr, float decay, float eps, float clip, const int64_t* __restrict__ step_dev,
               long long step_offset, double base_lr, long long max_step) {
  __shared__ float s_scale;
  __shared__ float s_part[kUpThreads];
  pdl_wait();
  // agent.py:393-395 on the device (the step counter lives in device memory under a CUDA graph):
  // the same double expression the host evaluates, rounded to float once
  if (step_dev != nullptr)
    lr = (float)((double)(max_step - (*step_dev + step_offset) + 1) / (double)max_step * base_lr);
  const int t = find_tensor(plan, blockIdx.x);
  // the tensor's partial sums of squares, added in a FIXED order (thread i takes chunks i, i+256, ..;
  // then a fixed tree): the same bits in every block and on every rank.  (One thread walking the 162
  // chunks of l4_w was most of this kernel's 27 us.)
  {
    float ss = 0.f;
    for (int c = plan.chunk_begin[t] + threadIdx.x; c < plan.chunk_begin[t + 1]; c += kUpThreads) ss += partial[c];
    s_part[threadIdx.x] = ss;
    __syncthreads();
#pragma unroll
    for (int h = kUpThreads / 2; h > 0; h >>= 1) {
      if (threadIdx.x < h) s_part[threadIdx.x] += s_part[threadIdx.x + h];
      __syncthreads();
    }
  }
  if (threadIdx.x == 0) {
    const float norm = sqrtf(s_part[0]);
    s_scale = clip / fmaxf(norm, clip);                      // tf.clip_by_norm
    if (norms_out && blockIdx.x == plan.chunk_begin[t]) norms_out[t] = norm;
  }
  __syncthreads();
  const float scale = s_scale;
  const float omd = 1.0f - decay;
  const int64_t beg = plan.off[t] + (int64_t)(blockIdx.x - plan.chunk_begin[t]) * kChunk;
  const int64_t end = beg + kChunk < plan.off[t + 1] ? beg + kChunk : plan.off[t + 1];
  // 16 x (3 loads, 2 stores) per thread one element at a time left this kernel latency-bound
  // (26.9 us for 13.6 MB, ncu r02): every thread now requests all its vectors -- four float4 of
  // each stream -- before it touches the first one
  if (((beg | end) & 3) == 0) {
    constexpr int U = 4;
    float4 g4[U], m4[U], w4[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t i = beg + 4 * (threadIdx.x + (int64_t)u * kUpThreads);
      if (i < end) {
        g4[u] = *reinterpret_cast<const float4*>(grads + i);
        m4[u] = *reinterpret_cast<const float4*>(rms + i);
        w4[u] = *reinterpret_cast<const float4*>(params + i);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t i = beg + 4 * (threadIdx.x + (int64_t)u * kUpThreads);
      if (i >= end) break;
      float gg[4] = {g4[u].x, g4[u].y, g4[u].z, g4[u].w};
      float mm[4] = {m4[u].x, m4[u].y, m4[u].z, m4[u].w};
      float ww[4] = {w4[u].x, w4[u].y, w4[u].z, w4[u].w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float g = gg[e] * scale;
        mm[e] = fmaf(fmaf(g, g, -mm[e]), omd, mm[e]);          // ms += (g*g - ms) * (1 - decay)
        ww[e] -= lr * g / sqrtf(mm[e] + eps);                   // epsilon inside the sqrt (TF)
      }
      *reinterpret_cast<float4*>(rms + i) = make_float4(mm[0], mm[1], mm[2], mm[3]);
      *reinterpret_cast<float4*>(params + i) = make_float4(ww[0], ww[1], ww[2], ww[3]);
    }
    return;
  }
  for (int64_t i = beg + threadIdx.x; i < end; i += kUpThreads) {
    const float g = grads[i] * scale;
    float ms = rms[i];
    ms = fmaf(fmaf(g, g, -ms), omd, ms);                     // ms += (g*g - ms) * (1 - decay)
    rms[i] = ms;
    params[i] -= lr * g / sqrtf(ms + eps);                   // epsilon inside the sqrt (TF)
  }
}

}  // namespace arl

using namespace arl;

static int clip_rmsprop_offsets(float* params, float* rms, const float* grads, const int64_t* offsets,
                                int num_tensors, float lr, float decay, float eps, float clip_norm,
                                float* norms_out, void* workspace, const int64_t* step_dev,
                                int64_t step_offset, double base_lr, int64_t max_step, void* stream,
                                bool exchange = false) {
  ARL_REQUIRE(params && rms && grads && workspace && offsets, "arl_clip_rmsprop: null pointer");
  ARL_REQUIRE(num_tensors >= 1 && num_tensors <= kMaxTensors, "arl_clip_rmsprop: %d tensors outside [1,%d]",
              num_tensors, kMaxTensors);
  ARL_REQUIRE(clip_norm > 0.f && eps >= 0.f, "arl_clip_rmsprop: clip_norm must be > 0, eps >= 0");
  UpdatePlan plan;
  plan.n = num_tensors;
  int chunks = 0;
  for (int t = 0; t < num_tensors; ++t) {
    ARL_REQUIRE(offsets[t + 1] >= offsets[t], "arl_clip_rmsprop: offsets must not decrease");
    plan.off[t] = offsets[t];
    plan.chunk_begin[t] = chunks;
    chunks += (int)((offsets[t + 1] - offsets[t] + kChunk - 1) / kChunk);
  }
  for (int t = num_tensors; t <= kMaxTensors; ++t) {
    plan.off[t] = offsets[num_tensors];
    plan.chunk_begin[t] = chunks;
  }
  cudaStream_t st = (cudaStream_t)stream;
  float* partial = (float*)workspace;
  if (exchange) {
    // `grads` holds this rank's gradient: publish it, then sum every rank's copy back into it
    const P2PView* v = p2p_view();
    ARL_REQUIRE(v != nullptr, "arl_clip_rmsprop: exchange requested but arl_comm_enable_p2p has not succeeded");
    const int64_t n = offsets[num_tensors];
    ARL_REQUIRE(n <= v->count, "arl_clip_rmsprop: %lld parameters but the exchange buffer holds %lld",
                (long long)n, (long long)v->count);
    ARL_REQUIRE((reinterpret_cast<uintptr_t>(grads) & 15) == 0, "arl_clip_rmsprop: grads must be 16-byte aligned");
    int pgrid = (int)((n / 4 + 255) / 256);
    if (pgrid > 2 * num_sms()) pgrid = 2 * num_sms();
    if (pgrid < 1) pgrid = 1;
    p2p_publish_kernel<<<pgrid, 256, 0, st>>>(grads, *v, (long long)n);
    ARL_LAUNCH_CHECK("p2p_publish_kernel");
    float* gw = const_cast<float*>(grads);
    switch (v->nranks) {
      case 2: p2p_reduce_sumsq_kernel<2><<<chunks, kUpThreads, 0, st>>>(gw, partial, plan, *v); break;
      case 4: p2p_reduce_sumsq_kernel<4><<<chunks, kUpThreads, 0, st>>>(gw, partial, plan, *v); break;
      case 8: p2p_reduce_sumsq_kernel<8><<<chunks, kUpThreads, 0, st>>>(gw, partial, plan, *v); break;
      default:
        set_error("arl_exchange_clip_rmsprop: the peer-memory exchange is built for 2, 4 or 8 ranks (got %d)", v->nranks);
        return ARL_ERR_UNSUPPORTED;
    }
    ARL_LAUNCH_CHECK("p2p_reduce_sumsq_kernel");
  } else {
    ARL_CUDA(launch_pdl(sumsq_kernel, dim3(chunks), dim3(kUpThreads), 0, st, grads, partial, plan));
    ARL_LAUNCH_CHECK("sumsq_kernel");
  }
  ARL_CUDA(launch_pdl(rmsprop_kernel, dim3(chunks), dim3(kUpThreads), 0, st, params, rms, grads, partial, norms_out,
                      plan, lr, decay, eps, clip_norm, step_dev, (long long)step_offset, base_lr,
                      (long long)max_step));
  ARL_LAUNCH_CHECK("rmsprop_kernel");
  return ARL_OK;
}

static int clip_rmsprop(float* params, float* rms, const float* grads, int action_size, float lr,
                        float decay, float eps, float clip_norm, float* norms_out, void* workspace,
                        const int64_t* step_dev, int64_t step_offset, double base_lr, int64_t max_step,
                        void* stream) {
  ARL_REQUIRE(action_size >= 1 && action_size <= ARL_MAX_ACTIONS,
              "arl_clip_rmsprop: action_size %d outside [1,%d]", action_size, ARL_MAX_ACTIONS);
  const ParamLayout L = param_layout(action_size);
  return clip_rmsprop_offsets(params, rms, grads, L.off, ARL_NUM_TENSORS, lr, decay, eps, clip_norm, norms_out,
                              workspace, step_dev, step_offset, base_lr, max_step, stream);
}

extern "C" int arl_clip_rmsprop_layout(float* params, float* rms, const float* grads, const int64_t* offsets,
                                       int num_tensors, float lr, const int64_t* step_dev,
                                       int64_t step_offset, double base_lr, int64_t max_step, float decay,
                                       float eps, float clip_norm, float* norms_out, void* workspace,
                                       void* stream) {
  ARL_REQUIRE(step_dev == nullptr || max_step > 0, "arl_clip_rmsprop_layout: max_step <= 0");
  return clip_rmsprop_offsets(params, rms, grads, offsets, num_tensors, lr, decay, eps, clip_norm, norms_out,
                              workspace, step_dev, step_offset, base_lr, step_dev ? max_step : 1, stream);
}

extern "C" int arl_exchange_clip_rmsprop(float* params, float* rms, float* grads, const int64_t* offsets,
                                         int num_tensors, float lr, const int64_t* step_dev,
                                         int64_t step_offset, double base_lr, int64_t max_step, float decay,
                                         float eps, float clip_norm, float* norms_out, void* workspace,
                                         void* stream) {
  ARL_REQUIRE(step_dev == nullptr || max_step > 0, "arl_exchange_clip_rmsprop: max_step <= 0");
  return clip_rmsprop_offsets(params, rms, grads, offsets, num_tensors, lr, decay, eps, clip_norm, norms_out,
                              workspace, step_dev, step_offset, base_lr, step_dev ? max_step : 1, stream, true);
}

extern "C" int arl_clip_rmsprop(float* params, float* rms, const float* grads, int action_size,
                                float lr, float decay, float eps, float clip_norm, float* norms_out,
                                void* workspace, void* stream) {
  return clip_rmsprop(params, rms, grads, action_size, lr, decay, eps, clip_norm, norms_out, workspace,
                      nullptr, 0, 0.0, 1, stream);
}

extern "C" int arl_clip_rmsprop_sched(float* params, float* rms, const float* grads, int action_size,
                                      const int64_t* step_dev, int64_t step_offset, double base_lr,
                                      int64_t max_step, float decay, float eps, float clip_norm,
                                      float* norms_out, void* workspace, void* stream) {
  ARL_REQUIRE(step_dev && max_step > 0, "arl_clip_rmsprop_sched: null step counter or max_step <= 0");
  return clip_rmsprop(params, rms, grads, action_size, 0.f, decay, eps, clip_norm, norms_out, workspace,
                      step_dev, step_offset, base_lr, max_step, stream);
}
