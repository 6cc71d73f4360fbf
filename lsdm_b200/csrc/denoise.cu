// Pointwise ends of the x0 network and the fused x0 -> posterior -> ancestral-sample epilogue
// (reference model/diffusion_utils.py:51-54,98-103; model/sdm.py:204; diffusion/gaussian_diffusion.py:238-256,
// 266-269,356-357,545-560).  The dense middle layers go through the GEMM interface.
#include "kernels.cuh"

namespace lsdm {

namespace {

// xin = x + add (written back in place: the reference's `x += pcd_out`); h1 = sigmoid(E0 xin + b), 3 -> 64.
// one thread per (row, 4 output channels): 16 threads per row.
__global__ void __launch_bounds__(256) pose_embed0_kernel(float* __restrict__ x, const float* __restrict__ add,
                                                          const float* __restrict__ w, const float* __restrict__ b,
                                                          int64_t rows, float* __restrict__ h1, float* __restrict__ h1_lo) {
  __shared__ float sw[64 * 3], sb[64];
  for (int i = threadIdx.x; i < 192; i += blockDim.x) sw[i] = w[i];
  if (threadIdx.x < 64) sb[threadIdx.x] = b[threadIdx.x];
  __syncthreads();
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t row = i >> 4;
  int q = (int)(i & 15);
  const bool valid = row < rows;
  float vx = 0.f, vy = 0.f, vz = 0.f;
  if (valid) {
    vx = x[row * 3];
    vy = x[row * 3 + 1];
    vz = x[row * 3 + 2];
  }
  __syncwarp();  // all 16 readers of a row have the old x before lane q==0 overwrites it
  if (!valid) return;
  if (add != nullptr) {
    vx += add[row * 3];
    vy += add[row * 3 + 1];
    vz += add[row * 3 + 2];
    if (q == 0) {
      x[row * 3] = vx;
      x[row * 3 + 1] = vy;
      x[row * 3 + 2] = vz;
    }
  }
  float o[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    int ch = q * 4 + j;
    float v = fmaf(sw[ch * 3 + 2], vz, fmaf(sw[ch * 3 + 1], vy, fmaf(sw[ch * 3], vx, sb[ch])));
    o[j] = sigmoidf_(v);
  }
  if (h1_lo != nullptr) {  // consumer is a pre-split 3xTF32 GEMM: store hi and lo planes
    float l[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) split_tf32(o[j], o[j], l[j]);
    *reinterpret_cast<float4*>(h1_lo + row * 64 + q * 4) = make_float4(l[0], l[1], l[2], l[3]);
  }
  *reinterpret_cast<float4*>(h1 + row * 64 + q * 4) = make_float4(o[0], o[1], o[2], o[3]);
}

// one warp per row: x0 = gelu(F2 f1 + b) (64 -> 3), then the posterior mean with the MUTATED x and the noise.
__global__ void __launch_bounds__(256) final3_kernel(const float* __restrict__ f1, const float* __restrict__ w,
                                                     const float* __restrict__ b, int64_t rows, float* __restrict__ x0_out,
                                                     const float* xin /* may alias sample_out */, const int64_t* __restrict__ t,
                                                     const float* __restrict__ c1, const float* __restrict__ c2,
                                                     const float* __restrict__ logvar, const float* __restrict__ noise,
                                                     float* sample_out, int clip) {
  __shared__ float sw[3 * 64 + 3];
  for (int i = threadIdx.x; i < 192; i += blockDim.x) sw[i] = w[i];
  if (threadIdx.x < 3) sw[192 + threadIdx.x] = b[threadIdx.x];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t row = warp0; row < rows; row += nwarps) {
    float2 v = *reinterpret_cast<const float2*>(f1 + row * 64 + lane * 2);
    float a0 = v.x * sw[lane * 2] + v.y * sw[lane * 2 + 1];
    float a1 = v.x * sw[64 + lane * 2] + v.y * sw[64 + lane * 2 + 1];
    float a2 = v.x * sw[128 + lane * 2] + v.y * sw[128 + lane * 2 + 1];
    a0 = warp_sum(a0);
    a1 = warp_sum(a1);
    a2 = warp_sum(a2);
    if (lane < 3) {
      float acc = lane == 0 ? a0 : (lane == 1 ? a1 : a2);
      float x0 = gelu_erf(acc + sw[192 + lane]);
      if (clip) x0 = fminf(fmaxf(x0, -1.0f), 1.0f);
      if (x0_out) x0_out[row * 3 + lane] = x0;
      if (sample_out) {
        int64_t tt = t[row / NPTS];
        float mean = c1[tt] * x0 + c2[tt] * xin[row * 3 + lane];
        float nz = tt != 0 ? 1.0f : 0.0f;
        sample_out[row * 3 + lane] = mean + nz * expf(0.5f * logvar[tt]) * noise[row * 3 + lane];
      }
    }
  }
}

__global__ void q_sample_kernel(const float* __restrict__ x0, const int64_t* __restrict__ t, const float* __restrict__ noise,
                                const float* __restrict__ sa, const float* __restrict__ s1a, int64_t n,
                                float* __restrict__ xt) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int64_t tt = t[i / (NPTS * 3)];
  xt[i] = sa[tt] * x0[i] + s1a[tt] * noise[i];
}

}  // namespace

int launch_pose_embed0(float* x, const float* add, const float* w, const float* b, int64_t rows, float* h1, float* h1_lo,
                       cudaStream_t st) {
  int64_t n = rows * 16;
  pose_embed0_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(x, add, w, b, rows, h1, h1_lo);
  return 1;
}

int launch_final3(const float* f1, const float* w, const float* b, int64_t rows, float* x0_out, const float* xin,
                  const int64_t* t, const float* c1, const float* c2, const float* logvar, const float* noise,
                  float* sample_out, int clip_denoised, cudaStream_t st) {
  int64_t g = (rows + 7) / 8;
  if (g > 148 * 16) g = 148 * 16;
  final3_kernel<<<(unsigned)g, 256, 0, st>>>(f1, w, b, rows, x0_out, xin, t, c1, c2, logvar, noise, sample_out, clip_denoised);
  return 1;
}

int launch_q_sample(const float* x0, const int64_t* t, const float* noise, const float* sa, const float* s1a, int B,
                    float* xt, cudaStream_t st) {
  int64_t n = (int64_t)B * NPTS * 3;
  q_sample_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(x0, t, noise, sa, s1a, n, xt);
  return 1;
}

}  // namespace lsdm
