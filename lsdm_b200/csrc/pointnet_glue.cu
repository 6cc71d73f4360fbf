// PointNet++ gather / interpolation glue between the selection kernels and the dense layers.
// The first 1x1 conv of every set-abstraction / feature-propagation block is linear in the gathered /
// interpolated features, so it is evaluated ONCE per source point by the GEMM ("P" tensors) and only the
// cheap, geometry-dependent remainder is done per grouped row here:
//   SA (pointnet2_utils.py:120-131,192-195): conv1([xyz_j - c_s || f_j]) = Wx.(xyz_j - c_s) + (Wf.f_j + b)
//   FP (pointnet2_utils.py:297-311):         conv1([f1_n || sum_k w_k f2_k]) = (Wa.f1_n + b) + sum_k w_k (Wb.f2_k)
#include "kernels.cuh"
#include "tc_ptx.cuh"

namespace lsdm {

namespace {

// one warp per grouped row; lanes stride over the C1 output channels (coalesced row writes)
__global__ void __launch_bounds__(256) sa_gather_kernel(const float* __restrict__ P, const float* __restrict__ Wx,
                                                        const float* __restrict__ Wf3, const float* __restrict__ bias,
                                                        const float* __restrict__ xyz, const float* __restrict__ new_xyz,
                                                        const int* __restrict__ group, int64_t rows, int N, int S, int C1,
                                                        float* __restrict__ h1, int round_out, int apply_relu) {
  extern __shared__ __align__(16) float sw[];  // Wx[C1][3] (+ Wf3[C1][3] + bias[C1] when P == nullptr)
  for (int i = threadIdx.x; i < C1 * 3; i += blockDim.x) sw[i] = Wx[i];
  if (P == nullptr) {
    for (int i = threadIdx.x; i < C1 * 3; i += blockDim.x) sw[C1 * 3 + i] = Wf3[i];
    for (int i = threadIdx.x; i < C1; i += blockDim.x) sw[C1 * 6 + i] = bias[i];
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  if (P != nullptr) {
    // Four consecutive rows of a group (same centroid) per warp iteration, every global load of the four issued before the
    // first use: the row -> point -> projected-row chain is three dependent latencies, paid once per four rows.
    // 128-bit accesses: a lane owns 4 consecutive channels per iteration (C1 is a multiple of 128 here).
    for (int64_t row0 = warp0 * 4; row0 < rows; row0 += nwarps * 4) {  // (rows is a multiple of 32)
      const int64_t cs = row0 >> 5;  // (cloud, centroid)
      const int64_t c = cs / S;
      const int4 j4 = *reinterpret_cast<const int4*>(group + row0);
      const int jj[4] = {j4.x, j4.y, j4.z, j4.w};
      const float* pc = new_xyz + cs * 3;
      const float cx = pc[0], cy = pc[1], cz = pc[2];
      float rx[4], ry[4], rz[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float* pj = xyz + (c * N + jj[u]) * 3;
        rx[u] = pj[0] - cx; ry[u] = pj[1] - cy; rz[u] = pj[2] - cz;
      }
      for (int c4 = lane; c4 < (C1 >> 2); c4 += 32) {
        float4 p[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) p[u] = *reinterpret_cast<const float4*>(P + (c * N + jj[u]) * C1 + c4 * 4);
        const float4 wa = *reinterpret_cast<const float4*>(sw + c4 * 12), wb = *reinterpret_cast<const float4*>(sw + c4 * 12 + 4),
                     wc = *reinterpret_cast<const float4*>(sw + c4 * 12 + 8);  // Wx rows of channels 4 c4 .. 4 c4 + 3
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          float r[4] = {fmaf(wa.z, rz[u], fmaf(wa.y, ry[u], fmaf(wa.x, rx[u], p[u].x))), fmaf(wb.y, rz[u], fmaf(wb.x, ry[u], fmaf(wa.w, rx[u], p[u].y))),
                        fmaf(wc.x, rz[u], fmaf(wb.w, ry[u], fmaf(wb.z, rx[u], p[u].z))), fmaf(wc.w, rz[u], fmaf(wc.z, ry[u], fmaf(wc.y, rx[u], p[u].w)))};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            if (apply_relu) r[e] = fmaxf(r[e], 0.0f);
            if (round_out) r[e] = tc::rna_tf32(r[e]);
          }
          *reinterpret_cast<float4*>(h1 + (row0 + u) * C1 + c4 * 4) = make_float4(r[0], r[1], r[2], r[3]);
        }
      }
    }
    return;
  }
  for (int64_t row = warp0; row < rows; row += nwarps) {
    int64_t cs = row >> 5;  // (cloud, centroid)
    int64_t c = cs / S;
    int j = group[row];
    const float* pj = xyz + (c * N + j) * 3;
    const float* pc = new_xyz + cs * 3;
    float jx = pj[0], jy = pj[1], jz = pj[2];
    float rx = jx - pc[0], ry = jy - pc[1], rz = jz - pc[2];
    float* out = h1 + row * C1;
    for (int ch = lane; ch < C1; ch += 32) {
      float v = sw[C1 * 6 + ch];
      v = fmaf(sw[ch * 3 + 0], rx, v);
      v = fmaf(sw[ch * 3 + 1], ry, v);
      v = fmaf(sw[ch * 3 + 2], rz, v);
      v = fmaf(sw[C1 * 3 + ch * 3 + 0], jx, v);
      v = fmaf(sw[C1 * 3 + ch * 3 + 1], jy, v);
      v = fmaf(sw[C1 * 3 + ch * 3 + 2], jz, v);
      if (apply_relu) v = fmaxf(v, 0.0f);
      out[ch] = round_out ? tc::rna_tf32(v) : v;
    }
  }
}

__global__ void __launch_bounds__(256) fp_combine_kernel(const float* __restrict__ Pa, const float* __restrict__ bias,
                                                         const float* __restrict__ Pb, const int* __restrict__ nn_idx,
                                                         const float* __restrict__ nn_w, int64_t rows, int N, int S, int C1,
                                                         float* __restrict__ h, int round_out, int apply_relu) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t row = warp0; row < rows; row += nwarps) {
    int64_t c = row / N;
    int i0 = nn_idx[row * 3], i1 = nn_idx[row * 3 + 1], i2 = nn_idx[row * 3 + 2];
    float w0 = nn_w[row * 3], w1 = nn_w[row * 3 + 1], w2 = nn_w[row * 3 + 2];
    const float* b0 = Pb + (c * S + i0) * C1;
    const float* b1 = Pb + (c * S + i1) * C1;
    const float* b2 = Pb + (c * S + i2) * C1;
    float* out = h + row * C1;
    // 128-bit accesses: a lane owns 4 consecutive channels per iteration (C1 is 128 or 256)
    for (int c4 = lane; c4 < (C1 >> 2); c4 += 32) {
      const float4 v = (Pa != nullptr) ? *reinterpret_cast<const float4*>(Pa + row * C1 + c4 * 4)
                                       : *reinterpret_cast<const float4*>(bias + c4 * 4);
      const float4 x0 = *reinterpret_cast<const float4*>(b0 + c4 * 4);
      const float4 x1 = *reinterpret_cast<const float4*>(b1 + c4 * 4);
      const float4 x2 = *reinterpret_cast<const float4*>(b2 + c4 * 4);
      float r[4] = {v.x + fmaf(x2.x, w2, fmaf(x1.x, w1, x0.x * w0)), v.y + fmaf(x2.y, w2, fmaf(x1.y, w1, x0.y * w0)),
                    v.z + fmaf(x2.z, w2, fmaf(x1.z, w1, x0.z * w0)), v.w + fmaf(x2.w, w2, fmaf(x1.w, w1, x0.w * w0))};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        if (apply_relu) r[e] = fmaxf(r[e], 0.0f);
        if (round_out) r[e] = tc::rna_tf32(r[e]);
      }
      *reinterpret_cast<float4*>(out + c4 * 4) = make_float4(r[0], r[1], r[2], r[3]);
    }
  }
}

__global__ void __launch_bounds__(256) head3_kernel(const float* __restrict__ h, const float* __restrict__ W,
                                                    const float* __restrict__ b, int64_t rows, float* __restrict__ out) {
  __shared__ float sw[3 * 128 + 3];
  for (int i = threadIdx.x; i < 3 * 128; i += blockDim.x) sw[i] = W[i];
  if (threadIdx.x < 3) sw[384 + threadIdx.x] = b[threadIdx.x];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t row = warp0; row < rows; row += nwarps) {
    float4 v = *reinterpret_cast<const float4*>(h + row * 128 + lane * 4);
    float a0 = 0.f, a1 = 0.f, a2 = 0.f;
    const float* w0 = sw + lane * 4;
    a0 = v.x * w0[0] + v.y * w0[1] + v.z * w0[2] + v.w * w0[3];
    a1 = v.x * w0[128] + v.y * w0[129] + v.z * w0[130] + v.w * w0[131];
    a2 = v.x * w0[256] + v.y * w0[257] + v.z * w0[258] + v.w * w0[259];
    a0 = warp_sum(a0);
    a1 = warp_sum(a1);
    a2 = warp_sum(a2);
    if (lane == 0) {
      out[row * 3 + 0] = a0 + sw[384];
      out[row * 3 + 1] = a1 + sw[385];
      out[row * 3 + 2] = a2 + sw[386];
    }
  }
}

__global__ void fold_bn_kernel(const float* __restrict__ W, const float* __restrict__ b, const float* __restrict__ gamma,
                               const float* __restrict__ beta, const float* __restrict__ mean,
                               const float* __restrict__ var, int N, int K, float eps, float* __restrict__ Wf,
                               float* __restrict__ bf) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < (int64_t)N * K) {
    int n = (int)(i / K);
    float s = gamma[n] / sqrtf(var[n] + eps);
    Wf[i] = W[i] * s;
  }
  if (i < N) {
    float s = gamma[i] / sqrtf(var[i] + eps);
    bf[i] = (b[i] - mean[i]) * s + beta[i];
  }
}

__global__ void copy_cols_kernel(const float* __restrict__ src, int ld_src, int col0, int ncols, int rows,
                                 float* __restrict__ dst) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < (int64_t)rows * ncols) {
    int r = (int)(i / ncols), c = (int)(i % ncols);
    dst[i] = src[(int64_t)r * ld_src + col0 + c];
  }
}

inline int grid_for_warps(int64_t rows, int warps_per_cta) {
  int64_t g = (rows + warps_per_cta - 1) / warps_per_cta;
  int64_t cap = 148 * 16;
  return (int)(g < cap ? g : cap);
}

}  // namespace

int launch_sa_gather(const float* P, const float* Wx, const float* Wf3, const float* bias, const float* xyz,
                     const float* new_xyz, const int* group, int n_clouds, int N, int S, int C1, float* h1,
                     int round_out, cudaStream_t st, int apply_relu) {
  int64_t rows = (int64_t)n_clouds * S * 32;
  size_t smem = (size_t)C1 * 7 * sizeof(float);
  sa_gather_kernel<<<grid_for_warps(rows, 8), 256, smem, st>>>(P, Wx, Wf3, bias, xyz, new_xyz, group, rows, N, S, C1, h1, round_out, apply_relu);
  return 1;
}

int launch_fp_combine(const float* Pa, const float* bias, const float* Pb, const int* nn_idx, const float* nn_w,
                      int n_clouds, int N, int S, int C1, float* h, int round_out, cudaStream_t st, int apply_relu) {
  int64_t rows = (int64_t)n_clouds * N;
  fp_combine_kernel<<<grid_for_warps(rows, 8), 256, 0, st>>>(Pa, bias, Pb, nn_idx, nn_w, rows, N, S, C1, h, round_out, apply_relu);
  return 1;
}

int launch_head3(const float* h, const float* W, const float* b, int64_t rows, float* out, cudaStream_t st) {
  head3_kernel<<<grid_for_warps(rows, 8), 256, 0, st>>>(h, W, b, rows, out);
  return 1;
}

int launch_fold_bn(const float* W, const float* b, const float* gamma, const float* beta, const float* mean,
                   const float* var, int N, int K, float eps, float* Wf, float* bf, cudaStream_t st) {
  int64_t n = (int64_t)N * K;
  fold_bn_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(W, b, gamma, beta, mean, var, N, K, eps, Wf, bf);
  return 1;
}

int launch_copy_cols(const float* src, int ld_src, int col0, int ncols, int rows, float* dst, cudaStream_t st) {
  int64_t n = (int64_t)rows * ncols;
  copy_cols_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(src, ld_src, col0, ncols, rows, dst);
  return 1;
}

}  // namespace lsdm
