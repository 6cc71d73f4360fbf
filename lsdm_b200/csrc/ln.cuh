// LayerNorm helpers shared by the transformer-style kernels (clip_text.cu, cf_mha.cu): one warp per row, the row held as
// PER values per lane (row width = 32 * PER), two-pass mean / biased variance in fp32, eps = 1e-5 (torch's default).
#pragma once
#include "common.cuh"

namespace lsdm {

constexpr float LN_EPS = 1e-5f;

template <int PER>
__device__ __forceinline__ void layer_norm_row(float (&v)[PER], const float* __restrict__ g, const float* __restrict__ b, int lane,
                                               float* __restrict__ dst) {
  constexpr int WIDTH = 32 * PER;
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < PER; ++i) s += v[i];
  const float mean = warp_sum(s) * (1.0f / WIDTH);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    const float d = v[i] - mean;
    q += d * d;
  }
  const float rstd = rsqrtf(warp_sum(q) * (1.0f / WIDTH) + LN_EPS);
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    const int c = i * 32 + lane;
    dst[c] = (v[i] - mean) * rstd * g[c] + b[c];
  }
}

// warp per row:  x[r] += y[r] (+ bias already in y);  h[r] = LN(x[r])
template <int PER>
__global__ void __launch_bounds__(256) add_ln_kernel(float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ g,
                                                          const float* __restrict__ b, int rows, float* __restrict__ h) {
  constexpr int WIDTH = 32 * PER;
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (r >= rows) return;
  float v[PER];
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    const int c = i * 32 + lane;
    v[i] = x[(int64_t)r * WIDTH + c] + y[(int64_t)r * WIDTH + c];
    x[(int64_t)r * WIDTH + c] = v[i];
  }
  layer_norm_row<PER>(v, g, b, lane, h + (int64_t)r * WIDTH);
}

}  // namespace lsdm
