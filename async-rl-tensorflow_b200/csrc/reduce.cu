// Deterministic reduction of per-CTA partial tensors: out[i] = sum_p partials[p][i], p in a
// fixed order.  Every weight-gradient kernel writes partials; nothing uses float atomics.
#include "common.cuh"

namespace arl {

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

int conv_init() { return ARL_OK; }

}  // namespace arl
