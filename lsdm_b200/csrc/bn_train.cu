// Train-mode BatchNorm for the PointNet++ 1x1-conv layers (reference pointnet2_utils.py:192-195,308-311, pointnet2.py:76
// under model.train()): per-channel batch statistics over EVERY row of the layer's pre-activation (all 9B clouds), biased
// variance for the normalisation, running statistics updated with momentum 0.1 and the unbiased variance, ReLU, optional
// Dropout mask (backbone head) and optional 32-row max-pool (set-abstraction output).
#include "kernels.cuh"

namespace lsdm {

namespace {

// sums[0:N] += column sums, sums[N:2N] += column sums of squares (double); y[M,N] row-major
__global__ void __launch_bounds__(256) col_stats_kernel(const float* __restrict__ y, int64_t M, int N, double* __restrict__ sums) {
  // block = 256 threads: tx = column lane (32 or more), ty = row lane
  const int cols_per = N < 256 ? N : 256;
  const int rows_per = 256 / cols_per;
  const int tx = threadIdx.x % cols_per, ty = threadIdx.x / cols_per;
  for (int c0 = 0; c0 < N; c0 += cols_per) {
    const int col = c0 + tx;
    double s = 0.0, q = 0.0;
    for (int64_t r = (int64_t)blockIdx.x * rows_per + ty; r < M; r += (int64_t)gridDim.x * rows_per) {
      float v = y[r * N + col];
      s += v;
      q += (double)v * v;
    }
    atomicAdd(&sums[col], s);
    atomicAdd(&sums[N + col], q);
  }
}

// y <- [mask *] relu((y - mean) * rstd * gamma + beta); optional out[g, col] = max over rows [32g, 32g+32)
// block 0 also folds the batch statistics into the running statistics.
__global__ void __launch_bounds__(256) bn_apply_kernel(float* __restrict__ y, int64_t M, int64_t Ms, int N, const double* __restrict__ sums,
                                                       const float* __restrict__ gamma, const float* __restrict__ beta,
                                                       float* __restrict__ run_mean, float* __restrict__ run_var, float eps,
                                                       float momentum, const float* __restrict__ drop_mask, int mask_points,
                                                       float* __restrict__ pooled, int round_out) {
  extern __shared__ float s_par[];  // scale[N], shift[N]
  for (int c = threadIdx.x; c < N; c += blockDim.x) {
    double mean = sums[c] / (double)Ms;  // Ms: rows behind the statistics (all shards when SyncBN is on)
    double var = sums[N + c] / (double)Ms - mean * mean;
    if (var < 0.0) var = 0.0;
    float sc = (float)(1.0 / sqrt(var + (double)eps)) * gamma[c];
    s_par[c] = sc;
    s_par[N + c] = beta[c] - (float)mean * sc;
    if (blockIdx.x == 0) {
      double unb = Ms > 1 ? var * ((double)Ms / (double)(Ms - 1)) : var;
      run_mean[c] = (1.0f - momentum) * run_mean[c] + momentum * (float)mean;
      run_var[c] = (1.0f - momentum) * run_var[c] + momentum * (float)unb;
    }
  }
  __syncthreads();
  const int cols_per = N < 256 ? N : 256;
  const int rows_per = 256 / cols_per;
  const int tx = threadIdx.x % cols_per, ty = threadIdx.x / cols_per;
  if (pooled == nullptr) {
    for (int c0 = 0; c0 < N; c0 += cols_per) {
      const int col = c0 + tx;
      const float sc = s_par[col], sh = s_par[N + col];
      for (int64_t r = (int64_t)blockIdx.x * rows_per + ty; r < M; r += (int64_t)gridDim.x * rows_per) {
        float v = fmaxf(fmaf(y[r * N + col], sc, sh), 0.0f);
        if (drop_mask) v *= drop_mask[((r / mask_points) * N + col) * mask_points + (r % mask_points)];
        if (round_out) {
          uint32_t rr;
          asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(rr) : "f"(v));
          v = __uint_as_float(rr);
        }
        y[r * N + col] = v;
      }
    }
  } else {
    const int64_t groups = M >> 5;
    for (int c0 = 0; c0 < N; c0 += cols_per) {
      const int col = c0 + tx;
      const float sc = s_par[col], sh = s_par[N + col];
      for (int64_t g = (int64_t)blockIdx.x * rows_per + ty; g < groups; g += (int64_t)gridDim.x * rows_per) {
        float m = 0.0f;
        for (int k = 0; k < 32; ++k) m = fmaxf(m, fmaxf(fmaf(y[(g * 32 + k) * N + col], sc, sh), 0.0f));
        if (round_out) {
          uint32_t rr;
          asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(rr) : "f"(m));
          m = __uint_as_float(rr);
        }
        pooled[g * N + col] = m;
      }
    }
  }
}

}  // namespace

int launch_col_stats(const float* y, int64_t M, int N, double* sums, cudaStream_t st) {
  cudaMemsetAsync(sums, 0, sizeof(double) * 2 * N, st);
  int64_t g = (M + 63) / 64;
  if (g > 148 * 8) g = 148 * 8;
  col_stats_kernel<<<(unsigned)g, 256, 0, st>>>(y, M, N, sums);
  return 1;
}

int launch_bn_apply(float* y, int64_t M, int64_t M_stat, int N, const double* sums, const float* gamma, const float* beta, float* run_mean,
                    float* run_var, const float* drop_mask, int mask_points, float* pooled, int round_out, cudaStream_t st) {
  int64_t g = (M + 63) / 64;
  if (g > 148 * 8) g = 148 * 8;
  bn_apply_kernel<<<(unsigned)g, 256, 2 * N * sizeof(float), st>>>(y, M, M_stat, N, sums, gamma, beta, run_mean, run_var, 1e-5f, 0.1f, drop_mask,
                                                                mask_points, pooled, round_out);
  return 1;
}

}  // namespace lsdm
