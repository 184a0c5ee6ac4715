// K3: the fc256 layer (agent.py:251, ops.py:32-46) forward and backward on the tcgen05 tensor
// cores (gemm_tc.cuh: TMEM accumulators, bf16x3 split of the fp32 operands, fp32 accumulate).
//   forward : h    = relu(a2 [N,2592] . W [2592,256] + b)        A = a2 (K-major), B = W (k-rows)
//   dgrad   : d_a2 = (d_h [N,256] . W^T) * (a2 > 0)              A = d_h, B = W rows (K-major)
//   wgrad   : dW   = a2^T [2592,N] . d_h [N,256]                 both operands sample-major;
//             split-K over samples, partials summed in a fixed order (deterministic)
//   bgrad   : db   = column sums of d_h
#include "gemm_tc.cuh"

namespace arl {

int reduce_partials(const float* partials, float* out, int num_partials, int n,
                    cudaStream_t stream);

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

// KB per variant: 8 stages must fit in 227 KB (N tile 256 -> KB 16, N tile 64 -> KB 32)
int fc_gemm(int variant, const float* A, const float* B, float* D, const float* extra, int M, int N,
            int K, int64_t lda, int64_t ldb, int64_t ldd, int k_splits, cudaStream_t st) {
  tc::GemmArgs g;
  g.A = A; g.B = B; g.D = D; g.extra = extra;
  g.M = M; g.N = N; g.K = K; g.lda = lda; g.ldb = ldb; g.ldd = ldd;
  const int n_tile = variant == 1 ? 64 : 256, kb = variant == 1 ? 32 : 16;
  if (k_splits < 1) k_splits = 1;
  g.k_chunk = ((K + k_splits - 1) / k_splits + kb - 1) / kb * kb;
  g.k_splits = (K + g.k_chunk - 1) / g.k_chunk;
  g.m_tiles = (M + tc::kTileM - 1) / tc::kTileM;
  g.n_tiles = (N + n_tile - 1) / n_tile;
  const int items = g.m_tiles * g.n_tiles * g.k_splits;
  switch (variant) {
    case 0: return tc::launch<tc::GemmPolicy<256, 16, false, true, tc::EPI_BIAS_RELU>>(g, items, st);
    // few rows (one env step): narrow N tiles so every SM works; KB = 32 keeps the A rows 128-B
    // segments (16 one-warp stages of KB = 16 measured 7 % slower: 64-B segments, same latency)
    case 1: return tc::launch<tc::GemmPolicy<64, 32, false, true, tc::EPI_BIAS_RELU>>(g, items, st);
    case 2: return tc::launch<tc::GemmPolicy<256, 16, false, false, tc::EPI_MASK>>(g, items, st);
    case 3:   // (was: K-major images transposed in registers; now identical to 4)
    case 4: return tc::launch<tc::GemmPolicy<256, 16, true, true, tc::EPI_PLAIN>>(g, items, st);
    default: set_error("fc_gemm: unknown variant %d", variant); return ARL_ERR_INVALID;
  }
}

// number of split-K slices fc_gemm(variant 3) produces for K samples and a request of `want`
int fc_wgrad_splits(int K, int want) {
  const int k_chunk = ((K + want - 1) / want + 15) / 16 * 16;
  return (K + k_chunk - 1) / k_chunk;
}

}  // namespace arl

using namespace arl;

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// Test hook: the four production instantiations of the tcgen05 GEMM on arbitrary shapes.
// variant 0/1: D = relu(A[M,K] . B[K,N] + extra[N])      (N_TILE 256 / 64; B stored [K][N])
// variant 2  : D = (A[M,K] . B[N,K]^T) masked by extra[M,N] > 0
// variant 3  : D[z] = A[K,M]^T . B[K,N] over split-K slice z (k_splits slices, M*N floats each)
extern "C" int arl_debug_gemm(int variant, const float* A, const float* B, float* D,
                              const float* extra, int M, int N, int K, int k_splits, void* stream) {
  ARL_REQUIRE(A && B && D, "arl_debug_gemm: null pointer");
  ARL_REQUIRE(M > 0 && N > 0 && K > 0 && N % 16 == 0 && (variant >= 3 || K % 8 == 0),
              "arl_debug_gemm: need M,N,K > 0, N %% 16 == 0, K %% 8 == 0");
  ARL_REQUIRE(variant < 3 || M % 8 == 0, "arl_debug_gemm: variants 3/4 need M %% 8 == 0");
  const int64_t lda = variant >= 3 ? M : K;
  const int64_t ldb = variant == 2 ? K : N;
  return fc_gemm(variant, A, B, D, extra, M, N, K, lda, ldb, N, variant >= 3 ? k_splits : 1,
                 (cudaStream_t)stream);
}

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
  // few samples (one env step): narrow N tiles so that every SM gets work
  const int variant = ((M + 127) / 128 >= num_sms()) ? 0 : 1;
  return fc_gemm(variant, a2, W, h, b, M, ARL_FC, ARL_A2_ELEMS, ARL_A2_ELEMS, ARL_FC, ARL_FC, 1,
                 (cudaStream_t)stream);
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
  // dgrad: d_a2 [M,2592] = d_h [M,256] . W^T, masked by a2 > 0.  B rows = W rows (K-major).
  int rc = fc_gemm(2, d_h, W, d_a2, a2, M, ARL_A2_ELEMS, ARL_FC, ARL_FC, ARL_FC, ARL_A2_ELEMS, 1, st);
  if (rc) return rc;
  // wgrad: dW [2592,256] = a2^T . d_h, 7 split-K slices x 21 row tiles = 147 work items
  float* part = (float*)workspace;
  rc = fc_gemm(4, a2, d_h, part, nullptr, ARL_A2_ELEMS, ARL_FC, M, ARL_A2_ELEMS, ARL_FC, ARL_FC, 7, st);
  if (rc) return rc;
  const int splits = fc_wgrad_splits(M, 7);
  rc = reduce_partials(part, gW, splits, ARL_A2_ELEMS * ARL_FC, st);
  if (rc) return rc;
  // bias grad
  const int grid = (int)(num_samples < num_sms() ? num_samples : num_sms());
  colsum256_kernel<<<grid, 256, 0, st>>>(d_h, part, num_samples);
  ARL_LAUNCH_CHECK("colsum256_kernel");
  return reduce_partials(part, gb, grid, ARL_FC, st);
}
