// K3 (round-1 first cut): the fc256 layer (agent.py:251, ops.py:32-46) forward and backward
// as a register-tiled fp32 SGEMM on CUDA cores.  The tcgen05 version replaces this file's
// kernels behind the same entry points (kernel table).
//   forward : h    = relu(a2 [N,2592] . W [2592,256] + b)
//   dgrad   : d_a2 = (d_h [N,256] . W^T) * (a2 > 0)            (relu of conv2 folded in)
//   wgrad   : dW   = a2^T [2592,N] . d_h [N,256]              (split-K over samples, deterministic)
//   bgrad   : db   = column sums of d_h
#include "common.cuh"

namespace arl {

int reduce_partials(const float* partials, float* out, int num_partials, int n,
                    cudaStream_t stream);

constexpr int BM = 128, BN = 128, BK = 16, PAD = 4;
enum { EPI_PLAIN = 0, EPI_BIAS_RELU = 1, EPI_MASK = 2 };

// C[M,N] = opA[M,K] . opB[K,N].  A_T: A stored [K][M]; else [M][K].  B_T: B stored [N][K];
// else [K][N].  Contiguous dimensions must be multiples of 4 floats and 16-byte aligned.
// gridDim.z > 1 = split-K: slice z covers k in [z*k_chunk, (z+1)*k_chunk) and writes C + z*M*N.
template <bool A_T, bool B_T, int EPI>
__global__ void __launch_bounds__(256)
sgemm_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ C,
             const float* __restrict__ extra, int M, int N, int K, int lda, int ldb, int ldc,
             int k_chunk) {
  __shared__ __align__(16) float As[BK][BM + PAD];
  __shared__ __align__(16) float Bs[BK][BN + PAD];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int kbeg = blockIdx.z * k_chunk;
  const int kend = min(K, kbeg + k_chunk);
  const int ty = tid >> 4, tx = tid & 15;

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  float4 ra[2], rb[2];
  auto load_tiles = [&](int k0) {
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int i = tid + 256 * r;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (A_T) {
        const int k = i >> 5, m4 = (i & 31) * 4;
        if (k0 + k < kend && m0 + m4 < M)
          v = *reinterpret_cast<const float4*>(A + (size_t)(k0 + k) * lda + m0 + m4);
      } else {
        const int m = i >> 2, k4 = (i & 3) * 4;
        if (m0 + m < M && k0 + k4 < kend)
          v = *reinterpret_cast<const float4*>(A + (size_t)(m0 + m) * lda + k0 + k4);
      }
      ra[r] = v;
      v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (!B_T) {
        const int k = i >> 5, n4 = (i & 31) * 4;
        if (k0 + k < kend && n0 + n4 < N)
          v = *reinterpret_cast<const float4*>(B + (size_t)(k0 + k) * ldb + n0 + n4);
      } else {
        const int n = i >> 2, k4 = (i & 3) * 4;
        if (n0 + n < N && k0 + k4 < kend)
          v = *reinterpret_cast<const float4*>(B + (size_t)(n0 + n) * ldb + k0 + k4);
      }
      rb[r] = v;
    }
  };
  auto store_tiles = [&]() {
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int i = tid + 256 * r;
      if (A_T) {
        const int k = i >> 5, m4 = (i & 31) * 4;
        *reinterpret_cast<float4*>(&As[k][m4]) = ra[r];
      } else {
        const int m = i >> 2, k4 = (i & 3) * 4;
        As[k4][m] = ra[r].x; As[k4 + 1][m] = ra[r].y; As[k4 + 2][m] = ra[r].z; As[k4 + 3][m] = ra[r].w;
      }
      if (!B_T) {
        const int k = i >> 5, n4 = (i & 31) * 4;
        *reinterpret_cast<float4*>(&Bs[k][n4]) = rb[r];
      } else {
        const int n = i >> 2, k4 = (i & 3) * 4;
        Bs[k4][n] = rb[r].x; Bs[k4 + 1][n] = rb[r].y; Bs[k4 + 2][n] = rb[r].z; Bs[k4 + 3][n] = rb[r].w;
      }
    }
  };

  if (kbeg < kend) load_tiles(kbeg);
  for (int k0 = kbeg; k0 < kend; k0 += BK) {
    __syncthreads();
    store_tiles();
    __syncthreads();
    if (k0 + BK < kend) load_tiles(k0 + BK);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[k][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[k][64 + tx * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
  }

  float* Cz = C + (size_t)blockIdx.z * M * (size_t)ldc;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (m >= M) continue;
#pragma unroll
    for (int jh = 0; jh < 2; ++jh) {
      const int n = n0 + (jh == 0 ? tx * 4 : 64 + tx * 4);
      if (n >= N) continue;
      float4 v = make_float4(acc[i][jh * 4], acc[i][jh * 4 + 1], acc[i][jh * 4 + 2],
                             acc[i][jh * 4 + 3]);
      if (EPI == EPI_BIAS_RELU) {
        const float4 bb = *reinterpret_cast<const float4*>(extra + n);
        v.x = fmaxf(v.x + bb.x, 0.f); v.y = fmaxf(v.y + bb.y, 0.f);
        v.z = fmaxf(v.z + bb.z, 0.f); v.w = fmaxf(v.w + bb.w, 0.f);
      } else if (EPI == EPI_MASK) {
        const float4 mm = *reinterpret_cast<const float4*>(extra + (size_t)m * ldc + n);
        v.x = mm.x > 0.f ? v.x : 0.f; v.y = mm.y > 0.f ? v.y : 0.f;
        v.z = mm.z > 0.f ? v.z : 0.f; v.w = mm.w > 0.f ? v.w : 0.f;
      }
      *reinterpret_cast<float4*>(Cz + (size_t)m * ldc + n) = v;
    }
  }
}

// partial column sums of X [rows, 256]: block b sums rows b, b+grid, ... -> partials[b][256]
__global__ void colsum256_kernel(const float* __restrict__ X, float* __restrict__ partials,
                                 int64_t rows) {
  const int col = threadIdx.x;
  float s0 = 0.f, s1 = 0.f;
  int64_t r = blockIdx.x;
  for (; r + gridDim.x < rows; r += 2 * (int64_t)gridDim.x) {
    s0 += X[r * 256 + col];
    s1 += X[(r + gridDim.x) * 256 + col];
  }
  if (r < rows) s0 += X[r * 256 + col];
  partials[(size_t)blockIdx.x * 256 + col] = s0 + s1;
}

constexpr int kFcSplitK = 8;

}  // namespace arl

using namespace arl;

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

extern "C" int arl_fc_forward(const float* params, const float* a2, float* h, int64_t num_samples,
                              void* stream) {
  ARL_REQUIRE(params && a2 && h, "arl_fc_forward: null pointer");
  ARL_REQUIRE(num_samples >= 0 && num_samples < (1LL << 31), "arl_fc_forward: bad num_samples");
  ARL_REQUIRE(aligned16(params) && aligned16(a2) && aligned16(h),
              "arl_fc_forward: pointers must be 16-byte aligned");
  if (num_samples == 0) return ARL_OK;
  const ParamLayout L = param_layout(1);
  const float* W = params + L.off[T_L4W];
  const float* b = params + L.off[T_L4B];
  const int M = (int)num_samples;
  dim3 grid(ARL_FC / BN, (M + BM - 1) / BM, 1);
  sgemm_kernel<false, false, EPI_BIAS_RELU><<<grid, 256, 0, (cudaStream_t)stream>>>(
      a2, W, h, b, M, ARL_FC, ARL_A2_ELEMS, ARL_A2_ELEMS, ARL_FC, ARL_FC, ARL_A2_ELEMS);
  ARL_LAUNCH_CHECK("sgemm_kernel<fc forward>");
  return ARL_OK;
}

extern "C" int arl_fc_backward(const float* params, const float* a2, const float* d_h, float* d_a2,
                               float* grads, void* workspace, int64_t num_samples, void* stream) {
  ARL_REQUIRE(params && a2 && d_h && d_a2 && grads && workspace, "arl_fc_backward: null pointer");
  ARL_REQUIRE(num_samples >= 0 && num_samples < (1LL << 31), "arl_fc_backward: bad num_samples");
  ARL_REQUIRE(aligned16(params) && aligned16(a2) && aligned16(d_h) && aligned16(d_a2) &&
                  aligned16(grads) && aligned16(workspace),
              "arl_fc_backward: pointers must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  const ParamLayout L = param_layout(1);
  float* gW = grads + L.off[T_L4W];
  float* gb = grads + L.off[T_L4B];
  if (num_samples == 0) {
    ARL_CUDA(cudaMemsetAsync(gW, 0, (size_t)(ARL_A2_ELEMS + 1) * ARL_FC * sizeof(float), st));
    return ARL_OK;
  }
  const float* W = params + L.off[T_L4W];
  const int M = (int)num_samples;
  // dgrad: d_a2 [M,2592] = d_h [M,256] . W^T, masked by a2 > 0.  W stored [2592][256] = [N][K].
  {
    dim3 grid((ARL_A2_ELEMS + BN - 1) / BN, (M + BM - 1) / BM, 1);
    sgemm_kernel<false, true, EPI_MASK><<<grid, 256, 0, st>>>(
        d_h, W, d_a2, a2, M, ARL_A2_ELEMS, ARL_FC, ARL_FC, ARL_FC, ARL_A2_ELEMS, ARL_FC);
    ARL_LAUNCH_CHECK("sgemm_kernel<fc dgrad>");
  }
  // wgrad: dW [2592,256] = a2^T . d_h, split over samples.
  {
    int splits = kFcSplitK;
    int k_chunk = ((M + splits - 1) / splits + BK - 1) / BK * BK;
    splits = (M + k_chunk - 1) / k_chunk;
    dim3 grid(ARL_FC / BN, (ARL_A2_ELEMS + BM - 1) / BM, splits);
    float* part = (float*)workspace;
    sgemm_kernel<true, false, EPI_PLAIN><<<grid, 256, 0, st>>>(
        a2, d_h, part, nullptr, ARL_A2_ELEMS, ARL_FC, M, ARL_A2_ELEMS, ARL_FC, ARL_FC, k_chunk);
    ARL_LAUNCH_CHECK("sgemm_kernel<fc wgrad>");
    int rc = reduce_partials(part, gW, splits, ARL_A2_ELEMS * ARL_FC, st);
    if (rc) return rc;
  }
  // bias grad
  {
    const int grid = (int)(num_samples < num_sms() ? num_samples : num_sms());
    float* part = (float*)workspace;      // wgrad partials already consumed (stream order)
    colsum256_kernel<<<grid, 256, 0, st>>>(d_h, part, num_samples);
    ARL_LAUNCH_CHECK("colsum256_kernel");
    return reduce_partials(part, gb, grid, ARL_FC, st);
  }
}
