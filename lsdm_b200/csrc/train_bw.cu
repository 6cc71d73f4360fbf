// Backward pass of GaussianDiffusion.training_losses through SceneDiffusionModel in model.train() mode
// (reference run/train_sdm.py:78-84 `mp_trainer.backward(loss)`; diffusion/gaussian_diffusion.py:1256-1342;
// model/sdm.py:131-218; model/pcd_backbone/pointnet2.py:61-80, pointnet2_utils.py:174-199,273-312;
// posa/posa_models.py:152-187,320-326; model/diffusion_utils.py).
//
// One call = a taped fp32 forward (every pre-activation and activation kept in a caller-provided tape workspace, BatchNorm
// with batch statistics, the caller's Dropout mask and noise) followed by the reverse sweep that accumulates the gradient
// of  g_mse * chamfer + g_cat * lambda_cat * CE  into a flat buffer laid out like the handle's weight arena.
// The selection results (FPS order, ball-query groups, 3-NN indices / weights) are inputs: they are integer / piecewise
// constant functions of the clouds and carry no gradient (the reference's index tensors do not either).
// Arithmetic: fp32 on the CUDA cores (a generic strided SGEMM serves every forward product, dgrad and split-K wgrad).
// Scatter-type backward ops (gather, 3-NN interpolation, chamfer's second direction) use fp32 atomics like PyTorch's own
// index / scatter backward kernels do.
#include <cstdio>
#include <string>
#include <vector>

#include "kernels.cuh"
#include "train_bw.cuh"

namespace lsdm {

namespace {

// ------------------------------------------------------------------------------------------------------------------
// generic strided SGEMM:  C[m,n] (+)= sum_k A(m,k) B(k,n) (+ bias[n]),  A(m,k) = A[m*sam + k*sak], B(k,n) = B[k*sbk + n*sbn]
// 64 x 64 tile, 16-deep k slices, 256 threads x (4 x 4).  gridDim.z > 1: split-K, partial sums added atomically.
// ------------------------------------------------------------------------------------------------------------------
struct SG {
  const float* A; int64_t sam, sak;
  const float* B; int64_t sbk, sbn;
  float* C; int64_t ldc;
  const float* bias;
  int M, N, K;
  int accumulate;  // C += instead of C =
  int kchunk;      // K range per blockIdx.z
};

__global__ void __launch_bounds__(256) sgemm_kernel(SG g) {
  __shared__ float As[16][64 + 4], Bs[16][64 + 4];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.x * 64, n0 = blockIdx.y * 64;
  const int kbeg = blockIdx.z * g.kchunk, kend = min(g.K, kbeg + g.kchunk);
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const bool a_kfast = g.sak == 1, b_kfast = g.sbk == 1;
  for (int k0 = kbeg; k0 < kend; k0 += 16) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = tid + i * 256;
      const int k = a_kfast ? (idx & 15) : (idx >> 6), m = a_kfast ? (idx >> 4) : (idx & 63);
      const int gm = m0 + m, gk = k0 + k;
      As[k][m] = (gm < g.M && gk < kend) ? g.A[(int64_t)gm * g.sam + (int64_t)gk * g.sak] : 0.f;
      const int kb = b_kfast ? (idx & 15) : (idx >> 6), n = b_kfast ? (idx >> 4) : (idx & 63);
      const int gn = n0 + n, gkb = k0 + kb;
      Bs[kb][n] = (gn < g.N && gkb < kend) ? g.B[(int64_t)gkb * g.sbk + (int64_t)gn * g.sbn] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[k][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[k][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int gm = m0 + ty * 4 + i;
    if (gm >= g.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gn = n0 + tx * 4 + j;
      if (gn >= g.N) continue;
      float v = acc[i][j];
      float* c = g.C + (int64_t)gm * g.ldc + gn;
      if (gridDim.z > 1) {
        if (blockIdx.z == 0 && g.bias) v += g.bias[gn];
        atomicAdd(c, v);
      } else {
        if (g.bias) v += g.bias[gn];
        *c = g.accumulate ? *c + v : v;
      }
    }
  }
}

void sgemm(cudaStream_t st, const float* A, int64_t sam, int64_t sak, const float* B, int64_t sbk, int64_t sbn, float* C, int64_t ldc,
           const float* bias, int M, int N, int K, bool accumulate, bool allow_splitk = false) {
  if (M <= 0 || N <= 0 || K <= 0) return;
  SG g{A, sam, sak, B, sbk, sbn, C, ldc, bias, M, N, K, accumulate ? 1 : 0, K};
  dim3 grid((M + 63) / 64, (N + 63) / 64, 1);
  if (allow_splitk && accumulate) {
    const int64_t tiles = (int64_t)grid.x * grid.y;
    int want = (int)((148 * 6 + tiles - 1) / tiles);
    int chunk = (K + want - 1) / want;
    chunk = ((chunk + 15) / 16) * 16;
    if (chunk < 256) chunk = 256;
    grid.z = (K + chunk - 1) / chunk;
    g.kchunk = chunk;
  }
  sgemm_kernel<<<grid, 256, 0, st>>>(g);
}

struct Lin {  // Y = X W^T + b with W[N, K] (row stride ldw); gW / gb: gradient slots (same layout)
  const float* W; const float* b; float* gW; float* gb; int N, K, ldw;
};

// The large forward and dgrad products go through the library's tcgen05 GEMM in 3xTF32 (hi / lo split of both operands inside the
// kernel: fp32-grade accuracy); everything it cannot take (tiny or odd shapes, strided weight slices, wgrad) stays on the SGEMM.
float* g_wt_scratch = nullptr;  // [1024 x 1024] transposed-weight scratch inside the tape (set per call)
void transpose(cudaStream_t st, const float* in, float* out, int batch, int R, int Cn);

bool tc_linear(cudaStream_t st, const float* A, int64_t lda, const float* W, int64_t ldw, float* C, int64_t ldc, const float* bias, int M, int N, int K) {
  if (M < 4096) return false;
  GemmArgs g{};
  g.A = A; g.lda = lda; g.W = W; g.ldw = ldw; g.C = C; g.ldc = ldc;
  g.bias = bias; g.bias_mode = bias ? 1 : 0;
  g.M = M; g.N = N; g.K = K; g.batch = 1; g.act = ACT_NONE; g.precision = 2;
  if (!gemm_tc_eligible(g)) return false;
  return launch_gemm(g, st) > 0;
}

// Z[M,N] = X W^T + b
void lin_fw(cudaStream_t st, const float* X, int64_t ldx, int M, const Lin& L, float* Z, int64_t ldz, bool accumulate = false) {
  if (!accumulate && tc_linear(st, X, ldx, L.W, L.ldw, Z, ldz, L.b, M, L.N, L.K)) return;
  sgemm(st, X, ldx, 1, L.W, 1, L.ldw, Z, ldz, L.b, M, L.N, L.K, accumulate);
}

__global__ void colsum_kernel(const float* __restrict__ z, int64_t ld, int64_t M, int N, float* __restrict__ out) {
  const int col = blockIdx.y * 32 + (threadIdx.x & 31);
  const int rl = threadIdx.x >> 5;  // 8 row lanes
  float s = 0.f;
  if (col < N)
    for (int64_t r = (int64_t)blockIdx.x * 8 + rl; r < M; r += (int64_t)gridDim.x * 8) s += z[r * ld + col];
  __shared__ float red[8][33];
  red[rl][threadIdx.x & 31] = s;
  __syncthreads();
  if (rl == 0 && col < N) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += red[i][threadIdx.x & 31];
    atomicAdd(&out[col], t);
  }
}
void colsum(cudaStream_t st, const float* z, int64_t ld, int64_t M, int N, float* out) {
  int64_t gx = (M + 255) / 256;
  if (gx > 592) gx = 592;
  if (gx < 1) gx = 1;
  colsum_kernel<<<dim3((unsigned)gx, (N + 31) / 32), 256, 0, st>>>(z, ld, M, N, out);
}

// gW += dZ^T X ; gb += colsum(dZ) ; dX (=|+=) dZ W
void lin_bw(cudaStream_t st, const float* X, int64_t ldx, int M, const Lin& L, const float* dZ, int64_t ldz, float* dX, int64_t lddx,
            bool accumulate_dx = false) {
  if (L.gW) sgemm(st, dZ, 1, ldz, X, ldx, 1, L.gW, L.ldw, nullptr, L.N, L.K, M, true, true);
  if (L.gb) colsum(st, dZ, ldz, M, L.N, L.gb);
  if (dX) {
    // dX = dZ W == dZ (W^T)^T: a linear layer with the transposed weight
    if (!accumulate_dx && g_wt_scratch && M >= 4096 && L.ldw == L.K && (int64_t)L.N * L.K <= 1024 * 1024 && (L.N % 32) == 0) {
      transpose(st, L.W, g_wt_scratch, 1, L.N, L.K);
      if (tc_linear(st, dZ, ldz, g_wt_scratch, L.N, dX, lddx, nullptr, M, L.K, L.N)) return;
    }
    sgemm(st, dZ, ldz, 1, L.W, L.ldw, 1, dX, lddx, nullptr, M, L.K, L.N, accumulate_dx);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// elementwise
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float act_grad(float z, int act) {
  switch (act) {
    case ACT_RELU: return z > 0.f ? 1.f : 0.f;
    case ACT_GELU: {
      const float cdf = 0.5f * (1.0f + erff(z * 0.70710678118654752440f));
      const float pdf = 0.3989422804014327f * expf(-0.5f * z * z);
      return cdf + z * pdf;
    }
    case ACT_SIGMOID: {
      const float s = 1.0f / (1.0f + expf(-z));
      return s * (1.0f - s);
    }
    case ACT_SILU: {
      const float s = 1.0f / (1.0f + expf(-z));
      return s + z * s * (1.0f - s);
    }
    default: return 1.f;
  }
}
__global__ void act_fw_kernel(const float* __restrict__ z, int64_t ldz, float* __restrict__ a, int64_t lda, int64_t rows, int cols, int act) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * cols) return;
  const int64_t r = i / cols;
  const int c = (int)(i % cols);
  a[r * lda + c] = apply_act_rt(z[r * ldz + c], act);
}
__global__ void act_bw_kernel(float* __restrict__ d, int64_t ldd, const float* __restrict__ z, int64_t ldz, int64_t rows, int cols, int act) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * cols) return;
  const int64_t r = i / cols;
  const int c = (int)(i % cols);
  d[r * ldd + c] *= act_grad(z[r * ldz + c], act);
}
void act_fw(cudaStream_t st, const float* z, int64_t ldz, float* a, int64_t lda, int64_t rows, int cols, int act) {
  const int64_t n = rows * cols;
  act_fw_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(z, ldz, a, lda, rows, cols, act);
}
void act_bw(cudaStream_t st, float* d, int64_t ldd, const float* z, int64_t ldz, int64_t rows, int cols, int act) {
  const int64_t n = rows * cols;
  act_bw_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d, ldd, z, ldz, rows, cols, act);
}

// ------------------------------------------------------------------------------------------------------------------
// BatchNorm (train): statistics over all M rows (+ the other shards through the all-reduce hook), ReLU, optional mask
// ------------------------------------------------------------------------------------------------------------------
// stat[c] = batch mean, stat[N + c] = 1 / sqrt(biased variance + eps)
__global__ void bn_finalize_kernel(const double* __restrict__ sums, double Mstat, int N, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float eps, float* __restrict__ stat) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= N) return;
  const double mean = sums[c] / Mstat;
  double var = sums[N + c] / Mstat - mean * mean;
  if (var < 0.0) var = 0.0;
  const float rstd = (float)(1.0 / sqrt(var + (double)eps));
  stat[c] = (float)mean;
  stat[N + c] = rstd;
}
// a = [mask *] relu((y - mean) * rstd * gamma + beta), four channels per thread (N is a multiple of 4)
__global__ void bn_fw_kernel(const float* __restrict__ y, int64_t M, int N, const float* __restrict__ stat, const float* __restrict__ gamma,
                             const float* __restrict__ beta, float* __restrict__ a, const float* __restrict__ mask, int mask_points) {
  const int64_t i4 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i4 * 4 >= M * N) return;
  const int64_t r = (i4 * 4) / N;
  const int c = (int)((i4 * 4) % N);
  const float4 v = reinterpret_cast<const float4*>(y)[i4];
  const float in[4] = {v.x, v.y, v.z, v.w};
  float o[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float t = fmaxf((in[j] - stat[c + j]) * stat[N + c + j] * gamma[c + j] + beta[c + j], 0.0f);
    if (mask) t *= mask[((r / mask_points) * N + c + j) * mask_points + (r % mask_points)];
    o[j] = t;
  }
  reinterpret_cast<float4*>(a)[i4] = make_float4(o[0], o[1], o[2], o[3]);
}
// sums2[c] += sum_r dyh, sums2[N + c] += sum_r dyh * xhat, with dyh = dA * mask * [relu > 0]
__global__ void __launch_bounds__(256) bn_bw_reduce_kernel(const float* __restrict__ dA, const float* __restrict__ y, const float* __restrict__ a,
                                                           int64_t M, int N, const float* __restrict__ stat, const float* __restrict__ mask,
                                                           int mask_points, double* __restrict__ sums2) {
  const int cols_per = N < 256 ? N : 256;
  const int rows_per = 256 / cols_per;
  const int tx = threadIdx.x % cols_per, ty = threadIdx.x / cols_per;
  for (int c0 = 0; c0 < N; c0 += cols_per) {
    const int col = c0 + tx;
    const float mean = stat[col], rstd = stat[N + col];
    double s = 0.0, q = 0.0;
    for (int64_t r = (int64_t)blockIdx.x * rows_per + ty; r < M; r += (int64_t)gridDim.x * rows_per) {
      const int64_t i = r * N + col;
      if (a[i] > 0.f) {
        float d = dA[i];
        if (mask) d *= mask[((r / mask_points) * N + col) * mask_points + (r % mask_points)];
        s += d;
        q += (double)d * ((y[i] - mean) * rstd);
      }
    }
    atomicAdd(&sums2[col], s);
    atomicAdd(&sums2[N + col], q);
  }
}
// means[c] = sum dyh / M, means[N + c] = sum dyh * xhat / M (floats, from the double sums)
__global__ void bn_bw_means_kernel(const double* __restrict__ sums2, double Mstat, int N, float* __restrict__ means) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < 2 * N) means[c] = (float)(sums2[c] / Mstat);
}
__global__ void bn_bw_apply_kernel(float* __restrict__ dA, const float* __restrict__ y, const float* __restrict__ a, int64_t M, int N,
                                   const float* __restrict__ stat, const float* __restrict__ gamma, const float* __restrict__ means,
                                   const float* __restrict__ mask, int mask_points) {
  const int64_t i4 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i4 * 4 >= M * N) return;
  const int64_t r = (i4 * 4) / N;
  const int c = (int)((i4 * 4) % N);
  const float4 dv = reinterpret_cast<const float4*>(dA)[i4], yv = reinterpret_cast<const float4*>(y)[i4], av = reinterpret_cast<const float4*>(a)[i4];
  const float din[4] = {dv.x, dv.y, dv.z, dv.w}, yin[4] = {yv.x, yv.y, yv.z, yv.w}, ain[4] = {av.x, av.y, av.z, av.w};
  float o[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float d = 0.f;
    if (ain[j] > 0.f) {
      d = din[j];
      if (mask) d *= mask[((r / mask_points) * N + c + j) * mask_points + (r % mask_points)];
    }
    const float xh = (yin[j] - stat[c + j]) * stat[N + c + j];
    o[j] = gamma[c + j] * stat[N + c + j] * (d - means[c + j] - xh * means[N + c + j]);
  }
  reinterpret_cast<float4*>(dA)[i4] = make_float4(o[0], o[1], o[2], o[3]);
}
__global__ void bn_param_grad_kernel(const double* __restrict__ sums2, int N, float* __restrict__ g_gamma, float* __restrict__ g_beta) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < N) {
    g_beta[c] += (float)sums2[c];
    g_gamma[c] += (float)sums2[N + c];
  }
}

// ------------------------------------------------------------------------------------------------------------------
// PointNet++ pieces
// ------------------------------------------------------------------------------------------------------------------
// Y1[(c,s,k), ch] = (P ? P[(c*N + j)*C1 + ch] : b0[ch] + Wf[ch,:] . x_j) + Wx[ch,:] . (x_j - c_s),  j = grp[c,s,k]
// W0: the first conv's weight [C1, cin], columns 0..2 = Wx, 3.. = Wf
__global__ void sa_gather_fw_kernel(const float* __restrict__ P, const float* __restrict__ W0, int cin, const float* __restrict__ b0,
                                    const float* __restrict__ xyz, const float* __restrict__ new_xyz, const int* __restrict__ grp,
                                    int64_t R, int N, int S, int C1, float* __restrict__ y) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= R * C1) return;
  const int64_t r = i / C1;
  const int ch = (int)(i % C1);
  const int64_t cs = r >> 5, c = cs / S;
  const int j = grp[r];
  const float* pj = xyz + (c * N + j) * 3;
  const float* pc = new_xyz + cs * 3;
  const float* w = W0 + (int64_t)ch * cin;
  float v;
  if (P) v = P[(c * N + j) * C1 + ch];
  else v = b0[ch] + w[3] * pj[0] + w[4] * pj[1] + w[5] * pj[2];
  v += w[0] * (pj[0] - pc[0]) + w[1] * (pj[1] - pc[1]) + w[2] * (pj[2] - pc[2]);
  y[i] = v;
}
// dP[(c*N+j), ch] += dY1[r, ch];  gW0[ch, 0:3] += dY1 . rel;  level 0 (dP == null): gW0[ch, 3:6] += dY1 . x_j, gb0 += dY1
__global__ void __launch_bounds__(256) sa_gather_bw_kernel(const float* __restrict__ dY, const float* __restrict__ xyz,
                                                           const float* __restrict__ new_xyz, const int* __restrict__ grp, int64_t R, int N,
                                                           int S, int C1, int cin, float* __restrict__ dP, float* __restrict__ gW0,
                                                           float* __restrict__ gb0, int rows_per_block) {
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_block;
  for (int ch = threadIdx.x; ch < C1; ch += blockDim.x) {
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, f0 = 0.f, f1 = 0.f, f2 = 0.f, sb = 0.f;
    for (int64_t r = r0; r < r0 + rows_per_block && r < R; ++r) {
      const int64_t cs = r >> 5, c = cs / S;
      const int j = grp[r];
      const float* pj = xyz + (c * N + j) * 3;
      const float* pc = new_xyz + cs * 3;
      const float d = dY[r * C1 + ch];
      a0 = fmaf(d, pj[0] - pc[0], a0);
      a1 = fmaf(d, pj[1] - pc[1], a1);
      a2 = fmaf(d, pj[2] - pc[2], a2);
      if (dP) {
        atomicAdd(&dP[(c * N + j) * C1 + ch], d);
      } else {
        f0 = fmaf(d, pj[0], f0);
        f1 = fmaf(d, pj[1], f1);
        f2 = fmaf(d, pj[2], f2);
        sb += d;
      }
    }
    float* g = gW0 + (int64_t)ch * cin;
    atomicAdd(&g[0], a0);
    atomicAdd(&g[1], a1);
    atomicAdd(&g[2], a2);
    if (!dP) {
      atomicAdd(&g[3], f0);
      atomicAdd(&g[4], f1);
      atomicAdd(&g[5], f2);
      atomicAdd(&gb0[ch], sb);
    }
  }
}
// F[g, ch] = max_k A[(g*32 + k), ch]
__global__ void maxpool_fw_kernel(const float* __restrict__ a, int64_t G, int C, float* __restrict__ f) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= G * C) return;
  const int64_t g = i / C;
  const int ch = (int)(i % C);
  float m = a[(g * 32) * C + ch];
  for (int k = 1; k < 32; ++k) m = fmaxf(m, a[(g * 32 + k) * C + ch]);
  f[i] = m;
}
// dA[(g*32 + k*), ch] = dF[g, ch] at the FIRST maximum k*, 0 elsewhere (torch.max(dim) backward)
__global__ void maxpool_bw_kernel(const float* __restrict__ dF, const float* __restrict__ a, const float* __restrict__ f, int64_t G, int C,
                                  float* __restrict__ dA) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= G * C) return;
  const int64_t g = i / C;
  const int ch = (int)(i % C);
  const float m = f[i], d = dF[i];
  bool done = false;
  for (int k = 0; k < 32; ++k) {
    const int64_t o = (g * 32 + k) * C + ch;
    const bool hit = !done && a[o] == m;
    dA[o] = hit ? d : 0.f;
    done = done || hit;
  }
}
// Y1[(c,n), ch] = (Pa ? Pa[(c,n), ch] : b0[ch]) + sum_k w[c,n,k] * Pb[(c*S + idx[c,n,k]), ch]
__global__ void fp_combine_fw_kernel(const float* __restrict__ Pa, const float* __restrict__ b0, const float* __restrict__ Pb,
                                     const int* __restrict__ idx, const float* __restrict__ w, int64_t rows, int N, int S, int C1,
                                     float* __restrict__ y) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * C1) return;
  const int64_t r = i / C1;
  const int ch = (int)(i % C1);
  const int64_t c = r / N;
  float v = Pa ? Pa[i] : b0[ch];
  for (int k = 0; k < 3; ++k) v = fmaf(w[r * 3 + k], Pb[(c * S + idx[r * 3 + k]) * C1 + ch], v);
  y[i] = v;
}
__global__ void fp_combine_bw_kernel(const float* __restrict__ dY, const int* __restrict__ idx, const float* __restrict__ w, int64_t rows,
                                     int N, int S, int C1, float* __restrict__ dPb) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * C1) return;
  const int64_t r = i / C1;
  const int ch = (int)(i % C1);
  const int64_t c = r / N;
  const float d = dY[i];
  for (int k = 0; k < 3; ++k) atomicAdd(&dPb[(c * S + idx[r * 3 + k]) * C1 + ch], w[r * 3 + k] * d);
}

// ------------------------------------------------------------------------------------------------------------------
// small per-sample pieces
// ------------------------------------------------------------------------------------------------------------------
__global__ void gather_pe_kernel(const float* __restrict__ pe, const int64_t* __restrict__ t, int B, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < B * LAT) out[i] = pe[t[i / LAT] * LAT + (i % LAT)];
}
// s[b, 0:128] = ts[b], s[b, 128:256] = enc[b]
__global__ void concat2_kernel(const float* __restrict__ a, const float* __restrict__ b, int B, float* __restrict__ s) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < B * 256) {
    const int bb = i / 256, j = i % 256;
    s[i] = j < 128 ? a[bb * 128 + j] : b[bb * 128 + j - 128];
  }
}
// zt[(b,o), 0:32] = ec[(b,o)], zt[(b,o), 32:160] = enc[b]
__global__ void concat_trans_kernel(const float* __restrict__ ec, const float* __restrict__ enc, int B, float* __restrict__ zt) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < B * NOBJ * 160) {
    const int bo = i / 160, j = i % 160;
    zt[i] = j < 32 ? ec[bo * 32 + j] : enc[(bo / NOBJ) * 128 + j - 32];
  }
}
// dec[(b,o), :] += dzt[(b,o), 0:32]; denc[b, :] += sum_o dzt[(b,o), 32:160]
__global__ void concat_trans_bw_kernel(const float* __restrict__ dzt, int B, float* __restrict__ dec, float* __restrict__ denc) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < B * NOBJ * 32) dec[i] += dzt[(i / 32) * 160 + (i % 32)];
  if (i < B * 128) {
    const int b = i / 128, j = i % 128;
    float s = 0.f;
    for (int o = 0; o < NOBJ; ++o) s += dzt[(b * NOBJ + o) * 160 + 32 + j];
    denc[i] += s;
  }
}
// out[b, p, s] = in[b, s, p]   ([B,256,1024] -> [B,1024,256]) and the inverse
__global__ void transpose_kernel(const float* __restrict__ in, float* __restrict__ out, int R, int Cn) {  // in[b][R][Cn] -> out[b][Cn][R]
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const float* src = in + (int64_t)b * R * Cn;
  float* dst = out + (int64_t)b * R * Cn;
  int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 32 + threadIdx.y;
  for (int j = 0; j < 32; j += 8)
    if (x < Cn && y + j < R) tile[threadIdx.y + j][threadIdx.x] = src[(int64_t)(y + j) * Cn + x];
  __syncthreads();
  x = blockIdx.y * 32 + threadIdx.x;
  y = blockIdx.x * 32 + threadIdx.y;
  for (int j = 0; j < 32; j += 8)
    if (x < R && y + j < Cn) dst[(int64_t)(y + j) * R + x] = tile[threadIdx.x][threadIdx.y + j];
}
void transpose(cudaStream_t st, const float* in, float* out, int batch, int R, int Cn) {
  transpose_kernel<<<dim3((Cn + 31) / 32, (R + 31) / 32, batch), dim3(32, 8), 0, st>>>(in, out, R, Cn);
}

// attn_layer weights (model/sdm.py:180-182): sm[b,h,o] = softmax_o(q_h . k_oh / 4 + mask'), w[b,o] = mean_h sm
__global__ void attw_fw_kernel(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ mask_global, int B, int Bg,
                               int b_off, float* __restrict__ sm, float* __restrict__ w) {
  const int b = blockIdx.x, h = threadIdx.x;
  __shared__ float s_sm[NHEAD][NOBJ];
  if (h < NHEAD) {
    float lg[NOBJ], mx = -INFINITY;
    const int mrow = (int)((((int64_t)(b + b_off)) * NHEAD + h) % Bg);
    for (int o = 0; o < NOBJ; ++o) {
      float acc = 0.f;
      for (int d = 0; d < 16; ++d) acc = fmaf(q[b * 128 + h * 16 + d], k[(b * NOBJ + o) * 128 + h * 16 + d], acc);
      lg[o] = acc * 0.25f + mask_global[mrow * NOBJ + o];
      mx = fmaxf(mx, lg[o]);
    }
    float sum = 0.f;
    for (int o = 0; o < NOBJ; ++o) {
      lg[o] = expf(lg[o] - mx);
      sum += lg[o];
    }
    for (int o = 0; o < NOBJ; ++o) {
      s_sm[h][o] = lg[o] / sum;
      sm[(b * NHEAD + h) * NOBJ + o] = lg[o] / sum;
    }
  }
  __syncthreads();
  if (h < NOBJ) {
    float acc = 0.f;
    for (int hh = 0; hh < NHEAD; ++hh) acc += s_sm[hh][h];
    w[b * NOBJ + h] = acc / NHEAD;
  }
}
__global__ void attw_bw_kernel(const float* __restrict__ dw, const float* __restrict__ sm, const float* __restrict__ q, const float* __restrict__ k,
                               int B, float* __restrict__ dq, float* __restrict__ dk) {
  const int b = blockIdx.x, h = threadIdx.x;
  if (h >= NHEAD) return;
  float dl[NOBJ], dot = 0.f;
  for (int o = 0; o < NOBJ; ++o) dot += sm[(b * NHEAD + h) * NOBJ + o] * dw[b * NOBJ + o] / NHEAD;
  for (int o = 0; o < NOBJ; ++o) {
    const float p = sm[(b * NHEAD + h) * NOBJ + o];
    dl[o] = p * (dw[b * NOBJ + o] / NHEAD - dot) * 0.25f;
  }
  for (int d = 0; d < 16; ++d) {
    float acc = 0.f;
    for (int o = 0; o < NOBJ; ++o) {
      acc = fmaf(dl[o], k[(b * NOBJ + o) * 128 + h * 16 + d], acc);
      dk[(b * NOBJ + o) * 128 + h * 16 + d] = dl[o] * q[b * 128 + h * 16 + d];
    }
    dq[b * 128 + h * 16 + d] = acc;
  }
}

// scramble #1: p1[b, g] = Fb[b, g % 9, g / 9] * w[b, g % 9], g in [0, 9*3072)
__global__ void scramble1_fw_kernel(const float* __restrict__ Fb, const float* __restrict__ w, int B, float* __restrict__ p1) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int per = NOBJ * NPTS * 3;
  if (i >= (int64_t)B * per) return;
  const int b = (int)(i / per), g = (int)(i % per), o = g % NOBJ, c = g / NOBJ;
  p1[i] = Fb[((int64_t)b * NOBJ + o) * (NPTS * 3) + c] * w[b * NOBJ + o];
}
// dFb[b,o,c] = dp1[b, c*9+o] * w[b,o];  dw[b,o] += sum_c dp1[b, c*9+o] * Fb[b,o,c]   (one CTA per (b,o))
__global__ void __launch_bounds__(256) scramble1_bw_kernel(const float* __restrict__ dp1, const float* __restrict__ Fb, const float* __restrict__ w,
                                                           float* __restrict__ dFb, float* __restrict__ dw) {
  const int bo = blockIdx.x, b = bo / NOBJ, o = bo % NOBJ;
  const float wv = w[bo];
  float s = 0.f;
  for (int c = threadIdx.x; c < NPTS * 3; c += blockDim.x) {
    const float d = dp1[(int64_t)b * NOBJ * NPTS * 3 + (int64_t)c * NOBJ + o];
    const int64_t fi = (int64_t)bo * (NPTS * 3) + c;
    dFb[fi] = d * wv;
    s = fmaf(d, Fb[fi], s);
  }
  __shared__ float red[8];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += red[i];
    dw[bo] = t;
  }
}

struct PaW {  // pcd_attention + point_wise_trans_layer parameters and their gradient slots
  const float *wk, *wv, *inb, *wo, *bo, *wpt, *bpt;
  float *g_wk, *g_wv, *g_inb, *g_wo, *g_bo, *g_wpt, *g_bpt;
};
// forward per (b,o): collapsed 12-head attention over the 1024 points of p1[b,o] + pointwise 15 -> 3 (pre-GELU saved)
__global__ void __launch_bounds__(256) pattn_fw_kernel(PaW w, const float* __restrict__ p1, const float* __restrict__ qq, float* __restrict__ ctx_out,
                                                       float* __restrict__ pa_out, float* __restrict__ pwz, float* __restrict__ pw) {
  __shared__ float s_red[8][TRANS][3];
  __shared__ float s_max[TRANS], s_ctx[TRANS], s_pa[TRANS];
  const int bo = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* P = p1 + (int64_t)bo * NPTS * 3;
  float lmax[TRANS];
  for (int h = 0; h < TRANS; ++h) lmax[h] = -INFINITY;
  for (int p = tid; p < NPTS; p += 256) {
    const float x = P[p * 3], y = P[p * 3 + 1], z = P[p * 3 + 2];
    for (int h = 0; h < TRANS; ++h) {
      const float kk = w.wk[h * 3] * x + w.wk[h * 3 + 1] * y + w.wk[h * 3 + 2] * z + w.inb[TRANS + h];
      lmax[h] = fmaxf(lmax[h], qq[bo * TRANS + h] * kk);
    }
  }
  for (int h = 0; h < TRANS; ++h) {
    const float m = warp_max(lmax[h]);
    if (lane == 0) s_red[warp][h][0] = m;
  }
  __syncthreads();
  if (tid < TRANS) {
    float m = s_red[0][tid][0];
    for (int k = 1; k < 8; ++k) m = fmaxf(m, s_red[k][tid][0]);
    s_max[tid] = m;
  }
  __syncthreads();
  float lsum[TRANS], lacc[TRANS];
  for (int h = 0; h < TRANS; ++h) lsum[h] = 0.f, lacc[h] = 0.f;
  for (int p = tid; p < NPTS; p += 256) {
    const float x = P[p * 3], y = P[p * 3 + 1], z = P[p * 3 + 2];
    for (int h = 0; h < TRANS; ++h) {
      const float kk = w.wk[h * 3] * x + w.wk[h * 3 + 1] * y + w.wk[h * 3 + 2] * z + w.inb[TRANS + h];
      const float vv = w.wv[h * 3] * x + w.wv[h * 3 + 1] * y + w.wv[h * 3 + 2] * z + w.inb[2 * TRANS + h];
      const float e = expf(qq[bo * TRANS + h] * kk - s_max[h]);
      lsum[h] += e;
      lacc[h] = fmaf(e, vv, lacc[h]);
    }
  }
  for (int h = 0; h < TRANS; ++h) {
    const float s = warp_sum(lsum[h]), a = warp_sum(lacc[h]);
    if (lane == 0) {
      s_red[warp][h][1] = s;
      s_red[warp][h][2] = a;
    }
  }
  __syncthreads();
  if (tid < TRANS) {
    float s = 0.f, a = 0.f;
    for (int k = 0; k < 8; ++k) {
      s += s_red[k][tid][1];
      a += s_red[k][tid][2];
    }
    s_ctx[tid] = a / s;
    ctx_out[bo * TRANS + tid] = a / s;
    // softmax statistics for the backward pass: ctx_out[.., 12:24] = max, [24:36] = sum
    ctx_out[(int64_t)gridDim.x * TRANS + bo * TRANS + tid] = s_max[tid];
    ctx_out[(int64_t)2 * gridDim.x * TRANS + bo * TRANS + tid] = s;
  }
  __syncthreads();
  if (tid < TRANS) {
    float acc = w.bo[tid];
    for (int k = 0; k < TRANS; ++k) acc = fmaf(w.wo[tid * TRANS + k], s_ctx[k], acc);
    s_pa[tid] = acc;
    pa_out[bo * TRANS + tid] = acc;
  }
  __syncthreads();
  for (int i = tid; i < NPTS * 3; i += 256) {
    const int p = i / 3, d = i % 3;
    float v = w.bpt[d];
    for (int k = 0; k < 3; ++k) v = fmaf(w.wpt[d * 15 + k], P[p * 3 + k], v);
    for (int k = 0; k < TRANS; ++k) v = fmaf(w.wpt[d * 15 + 3 + k], s_pa[k], v);
    pwz[(int64_t)bo * NPTS * 3 + i] = v;
    pw[(int64_t)bo * NPTS * 3 + i] = gelu_erf(v);
  }
}
// backward per (b,o).  dz = d(pw pre-GELU) [1024,3] (already multiplied by gelu').  Outputs dp1[b,o] (=), dqq[bo] (=), parameter grads (+=).
__global__ void __launch_bounds__(256) pattn_bw_kernel(PaW w, const float* __restrict__ p1, const float* __restrict__ qq, const float* __restrict__ ctx_st,
                                                       const float* __restrict__ pa, const float* __restrict__ dz, float* __restrict__ dp1,
                                                       float* __restrict__ dqq) {
  __shared__ float s_acc[8][64];
  __shared__ float s_dpa[TRANS], s_dctx[TRANS], s_tot[64];
  const int bo = blockIdx.x, nbo = gridDim.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* P = p1 + (int64_t)bo * NPTS * 3;
  const float* DZ = dz + (int64_t)bo * NPTS * 3;
  float* DP = dp1 + (int64_t)bo * NPTS * 3;
  // ---- pointwise layer: dWpt[d, 0:3] += dz_d p1, dWpt[d, 3:15] += dz_d pa, dbpt += dz, dpa += Wpt[:,3:]^T dz, dp1 = Wpt[:, :3]^T dz ----
  float acc[15];  // [d*3 + k] for the p1 part (9), then sum_p dz_d (3) -> indexes 9..11
  for (int i = 0; i < 15; ++i) acc[i] = 0.f;
  for (int p = tid; p < NPTS; p += 256) {
    const float x[3] = {P[p * 3], P[p * 3 + 1], P[p * 3 + 2]};
    const float d3[3] = {DZ[p * 3], DZ[p * 3 + 1], DZ[p * 3 + 2]};
    for (int d = 0; d < 3; ++d) {
      for (int k = 0; k < 3; ++k) acc[d * 3 + k] = fmaf(d3[d], x[k], acc[d * 3 + k]);
      acc[9 + d] += d3[d];
    }
    for (int k = 0; k < 3; ++k) DP[p * 3 + k] = w.wpt[k] * d3[0] + w.wpt[15 + k] * d3[1] + w.wpt[30 + k] * d3[2];
  }
  for (int i = 0; i < 12; ++i) {
    const float s = warp_sum(acc[i]);
    if (lane == 0) s_acc[warp][i] = s;
  }
  __syncthreads();
  if (tid < 12) {
    float s = 0.f;
    for (int k = 0; k < 8; ++k) s += s_acc[k][tid];
    s_tot[tid] = s;
    if (tid < 9) atomicAdd(&w.g_wpt[(tid / 3) * 15 + (tid % 3)], s);
    else atomicAdd(&w.g_bpt[tid - 9], s);
  }
  __syncthreads();
  if (tid < TRANS) {
    float dpa = 0.f;
    for (int d = 0; d < 3; ++d) {
      dpa = fmaf(w.wpt[d * 15 + 3 + tid], s_tot[9 + d], dpa);
      atomicAdd(&w.g_wpt[d * 15 + 3 + tid], s_tot[9 + d] * pa[bo * TRANS + tid]);
    }
    s_dpa[tid] = dpa;
    atomicAdd(&w.g_bo[tid], dpa);
  }
  __syncthreads();
  if (tid < TRANS) {
    float dc = 0.f;
    for (int j = 0; j < TRANS; ++j) {
      dc = fmaf(w.wo[j * TRANS + tid], s_dpa[j], dc);
      atomicAdd(&w.g_wo[j * TRANS + tid], s_dpa[j] * ctx_st[bo * TRANS + tid]);
    }
    s_dctx[tid] = dc;
  }
  __syncthreads();
  // ---- attention: a_hj = exp(qq_h kk_jh - max_h) / sum_h;  dlogit_hj = a_hj dctx_h (vv_jh - ctx_h) ----
  // per head: dqq_h = sum_j dl kk;  dkk_jh = dl qq_h;  dvv_jh = a dctx_h
  // parameter sums per head: gWk[h, :] += sum_j dkk p1_j, gbk[h] += sum_j dkk; gWv[h, :] += sum_j dvv p1_j, gbv[h] += sum_j dvv
  for (int h = 0; h < TRANS; ++h) {
    const float qh = qq[bo * TRANS + h], mx = ctx_st[(int64_t)nbo * TRANS + bo * TRANS + h], sm = ctx_st[(int64_t)2 * nbo * TRANS + bo * TRANS + h];
    const float ch = ctx_st[bo * TRANS + h], dch = s_dctx[h];
    const float wk0 = w.wk[h * 3], wk1 = w.wk[h * 3 + 1], wk2 = w.wk[h * 3 + 2], bk = w.inb[TRANS + h];
    const float wv0 = w.wv[h * 3], wv1 = w.wv[h * 3 + 1], wv2 = w.wv[h * 3 + 2], bv = w.inb[2 * TRANS + h];
    float r[9];  // dqq, gWk(3), gbk, gWv(3), gbv
    for (int i = 0; i < 9; ++i) r[i] = 0.f;
    for (int p = tid; p < NPTS; p += 256) {
      const float x = P[p * 3], y = P[p * 3 + 1], z = P[p * 3 + 2];
      const float kk = wk0 * x + wk1 * y + wk2 * z + bk, vv = wv0 * x + wv1 * y + wv2 * z + bv;
      const float a = expf(qh * kk - mx) / sm;
      const float dl = a * dch * (vv - ch);
      const float dkk = dl * qh, dvv = a * dch;
      r[0] = fmaf(dl, kk, r[0]);
      r[1] = fmaf(dkk, x, r[1]); r[2] = fmaf(dkk, y, r[2]); r[3] = fmaf(dkk, z, r[3]); r[4] += dkk;
      r[5] = fmaf(dvv, x, r[5]); r[6] = fmaf(dvv, y, r[6]); r[7] = fmaf(dvv, z, r[7]); r[8] += dvv;
      DP[p * 3] += dkk * wk0 + dvv * wv0;
      DP[p * 3 + 1] += dkk * wk1 + dvv * wv1;
      DP[p * 3 + 2] += dkk * wk2 + dvv * wv2;
    }
    for (int i = 0; i < 9; ++i) {
      const float s = warp_sum(r[i]);
      if (lane == 0) s_acc[warp][i] = s;
    }
    __syncthreads();
    if (tid < 9) {
      float s = 0.f;
      for (int k = 0; k < 8; ++k) s += s_acc[k][tid];
      if (tid == 0) dqq[bo * TRANS + h] = s;
      else if (tid <= 3) atomicAdd(&w.g_wk[h * 3 + tid - 1], s);
      else if (tid == 4) atomicAdd(&w.g_inb[TRANS + h], s);
      else if (tid <= 7) atomicAdd(&w.g_wv[h * 3 + tid - 5], s);
      else atomicAdd(&w.g_inb[2 * TRANS + h], s);
    }
    __syncthreads();
  }
}
// scene[b,p,d] = sum_o pw[b,o,p,d] m(f);  pcd_out = (scene + hm) / 2     and the backward: dpw = 0.5 dpcd m, dhm = 0.5 dpcd
__global__ void scene_bw_kernel(const float* __restrict__ dpcd, const float* __restrict__ mask_global, int B, int Bg, int b_off,
                                float* __restrict__ dpw, float* __restrict__ dhm) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)B * NPTS * 3) return;
  const int b = (int)(i / (NPTS * 3)), e = (int)(i % (NPTS * 3));
  const float d = 0.5f * dpcd[i];
  dhm[i] = d;
  for (int o = 0; o < NOBJ; ++o) {
    const int64_t f = (((int64_t)(b + b_off) * NOBJ + o) * NPTS * 3) + e;
    dpw[((int64_t)b * NOBJ + o) * NPTS * 3 + e] = d * mask_global[((f / NOBJ) % Bg) * NOBJ + (f % NOBJ)];
  }
}

// ---- GroupNorm(8) over (8 channels x n_pts points) of one sample, ReLU.  y, a: [B, ld_pts, 64] (first n_pts points used) ----
__global__ void __launch_bounds__(256) gn_fw_kernel(const float* __restrict__ y, int n_pts, int ld_pts, const float* __restrict__ gamma,
                                                    const float* __restrict__ beta, float eps, float* __restrict__ stat, float* __restrict__ a) {
  const int b = blockIdx.x / 8, g = blockIdx.x % 8, tid = threadIdx.x;
  const float* Y = y + (int64_t)b * ld_pts * 64;
  double s = 0.0, q = 0.0;
  for (int i = tid; i < n_pts * 8; i += 256) {
    const float v = Y[(i >> 3) * 64 + g * 8 + (i & 7)];
    s += v;
    q += (double)v * v;
  }
  __shared__ double rs[256], rq[256];
  rs[tid] = s;
  rq[tid] = q;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (tid < o) {
      rs[tid] += rs[tid + o];
      rq[tid] += rq[tid + o];
    }
    __syncthreads();
  }
  const double n = (double)n_pts * 8.0;
  const double mean = rs[0] / n;
  double var = rq[0] / n - mean * mean;
  if (var < 0.0) var = 0.0;
  const float rstd = (float)(1.0 / sqrt(var + (double)eps));
  if (tid == 0) {
    stat[blockIdx.x * 2] = (float)mean;
    stat[blockIdx.x * 2 + 1] = rstd;
  }
  float* A = a + (int64_t)b * ld_pts * 64;
  for (int i = tid; i < n_pts * 8; i += 256) {
    const int o = (i >> 3) * 64 + g * 8 + (i & 7), ch = g * 8 + (i & 7);
    A[o] = fmaxf((Y[o] - (float)mean) * rstd * gamma[ch] + beta[ch], 0.0f);
  }
}
// dA (grad wrt the post-ReLU output) -> dY in place; gamma / beta grads accumulated
__global__ void __launch_bounds__(256) gn_bw_kernel(float* __restrict__ dA, const float* __restrict__ y, const float* __restrict__ a, int n_pts,
                                                    int ld_pts, const float* __restrict__ gamma, const float* __restrict__ stat,
                                                    float* __restrict__ g_gamma, float* __restrict__ g_beta) {
  const int b = blockIdx.x / 8, g = blockIdx.x % 8, tid = threadIdx.x;
  const float mean = stat[blockIdx.x * 2], rstd = stat[blockIdx.x * 2 + 1];
  const int64_t base = (int64_t)b * ld_pts * 64;
  double s1 = 0.0, s2 = 0.0;   // sum dxh, sum dxh * xh  (dxh = dyh * gamma)
  float gb = 0.f, gg = 0.f;    // this thread's channel (i & 7 is constant per thread: 256 % 8 == 0)
  const int ch = g * 8 + (tid & 7);
  for (int i = tid; i < n_pts * 8; i += 256) {
    const int64_t o = base + (i >> 3) * 64 + ch;
    const float d = a[o] > 0.f ? dA[o] : 0.f;
    const float xh = (y[o] - mean) * rstd;
    gb += d;
    gg = fmaf(d, xh, gg);
    s1 += (double)d * gamma[ch];
    s2 += (double)d * gamma[ch] * xh;
  }
  __shared__ double rs[256], rq[256];
  __shared__ float rb[256], rg[256];
  rs[tid] = s1; rq[tid] = s2; rb[tid] = gb; rg[tid] = gg;
  __syncthreads();
  for (int o = 128; o >= 8; o >>= 1) {  // keeps the (tid & 7) channel classes separate for rb / rg
    if (tid < o) {
      rb[tid] += rb[tid + o];
      rg[tid] += rg[tid + o];
    }
    __syncthreads();
  }
  for (int o = 128; o > 0; o >>= 1) {
    if (tid < o) {
      rs[tid] += rs[tid + o];
      rq[tid] += rq[tid + o];
    }
    __syncthreads();
  }
  if (tid < 8) {
    atomicAdd(&g_beta[g * 8 + tid], rb[tid]);
    atomicAdd(&g_gamma[g * 8 + tid], rg[tid]);
  }
  const double n = (double)n_pts * 8.0;
  const float m1 = (float)(rs[0] / n), m2 = (float)(rq[0] / n);
  for (int i = tid; i < n_pts * 8; i += 256) {
    const int64_t o = base + (i >> 3) * 64 + ch;
    const float d = a[o] > 0.f ? dA[o] : 0.f;
    const float xh = (y[o] - mean) * rstd;
    dA[o] = rstd * (d * gamma[ch] - m1 - xh * m2);
  }
}
// hm[b,p,:] = h3[b, p/2, :] (p < 1024; h3 has 655 points) and its backward dh3[b,q,:] = sum_{p/2 == q} dhm[b,p,:]
__global__ void upsample2_fw_kernel(const float* __restrict__ h3, int B, float* __restrict__ hm) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * NPTS * 3) return;
  const int b = i / (NPTS * 3), p = (i / 3) % NPTS, d = i % 3;
  hm[i] = h3[(b * 655 + p / 2) * 3 + d];
}
__global__ void upsample2_bw_kernel(const float* __restrict__ dhm, int B, float* __restrict__ dh3) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * 655 * 3) return;
  const int b = i / (655 * 3), q = (i / 3) % 655, d = i % 3;
  dh3[i] = q < 512 ? dhm[(b * NPTS + 2 * q) * 3 + d] + dhm[(b * NPTS + 2 * q + 1) * 3 + d] : 0.f;
}
__global__ void slice_pts_kernel(const float* __restrict__ in, int B, int n_in, int n_out, int C, float* __restrict__ out, int to_padded) {
  // to_padded == 0: out[b, 0:n_out, :] = in[b, 0:n_out, :] (in has n_in points); 1: out[b, 0:n_in.., :] padded copy (zeros beyond n_out.. see caller)
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (!to_padded) {
    if (i >= (int64_t)B * n_out * C) return;
    const int b = (int)(i / ((int64_t)n_out * C));
    const int64_t r = i % ((int64_t)n_out * C);
    out[i] = in[(int64_t)b * n_in * C + r];
  } else {
    if (i >= (int64_t)B * n_in * C) return;
    const int b = (int)(i / ((int64_t)n_in * C));
    const int64_t r = i % ((int64_t)n_in * C);
    out[i] = r < (int64_t)n_out * C ? in[(int64_t)b * n_out * C + r] : 0.f;
  }
}

// ---- losses ----
// q_sample + z = x_t + pcd_out
__global__ void qsample_add_kernel(const float* __restrict__ x0, const int64_t* __restrict__ t, const float* __restrict__ noise,
                                   const float* __restrict__ sa, const float* __restrict__ s1a, const float* __restrict__ pcd, int64_t n,
                                   float* __restrict__ z) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t tt = t[i / (NPTS * 3)];
  z[i] = sa[tt] * x0[i] + s1a[tt] * noise[i] + pcd[i];
}
// chamfer (both directions) value and gradient wrt x: grid (4, B, 2), 256 threads; y = the fixed target
__global__ void __launch_bounds__(256) chamfer_bw_kernel(const float* __restrict__ x, const float* __restrict__ y, int B, float scale,
                                                         float* __restrict__ gx, float* __restrict__ val) {
  __shared__ float so[NPTS * 3];
  const int dir = blockIdx.z, b = blockIdx.y;
  const float* a = (dir == 0 ? x : y) + (int64_t)b * NPTS * 3;
  const float* o = (dir == 0 ? y : x) + (int64_t)b * NPTS * 3;
  for (int i = threadIdx.x; i < NPTS * 3; i += 256) so[i] = o[i];
  __syncthreads();
  const int i = blockIdx.x * 256 + threadIdx.x;
  const float px = a[i * 3], py = a[i * 3 + 1], pz = a[i * 3 + 2];
  float best = INFINITY;
  int bj = 0;
  for (int j = 0; j < NPTS; ++j) {
    const float dx = px - so[j * 3], dy = py - so[j * 3 + 1], dz = pz - so[j * 3 + 2];
    const float d = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
    if (d < best) best = d, bj = j;
  }
  const float s = 2.0f * scale / ((float)B * NPTS);
  float* G = gx + (int64_t)b * NPTS * 3;
  if (dir == 0) {  // x_i -> nearest y: gradient on x_i
    atomicAdd(&G[i * 3], s * (px - so[bj * 3]));
    atomicAdd(&G[i * 3 + 1], s * (py - so[bj * 3 + 1]));
    atomicAdd(&G[i * 3 + 2], s * (pz - so[bj * 3 + 2]));
  } else {         // y_i -> nearest x_bj: gradient on x_bj
    atomicAdd(&G[bj * 3], s * (so[bj * 3] - px));
    atomicAdd(&G[bj * 3 + 1], s * (so[bj * 3 + 1] - py));
    atomicAdd(&G[bj * 3 + 2], s * (so[bj * 3 + 2] - pz));
  }
  const float v = warp_sum(best);
  if ((threadIdx.x & 31) == 0) atomicAdd(val, v / ((float)B * NPTS));
}
// probs = softmax(h); cat = mean_b CE(probs as logits, argmax target); dh = d cat / d h * scale   (one warp per sample)
__global__ void cat_fw_bw_kernel(const float* __restrict__ h, const float* __restrict__ target, int B, int C, float scale, float* __restrict__ dh,
                                 float* __restrict__ val) {
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (b >= B) return;
  const float hv = lane < C ? h[b * C + lane] : -INFINITY;
  const float tv = lane < C ? target[b * C + lane] : -INFINITY;
  const float tmax = warp_max(tv);
  const int cls = __ffs(__ballot_sync(0xffffffffu, tv == tmax && lane < C)) - 1;
  const float mx = warp_max(hv);
  const float e = lane < C ? expf(hv - mx) : 0.f;
  const float p = e / warp_sum(e);                    // probs
  const float pm = warp_max(lane < C ? p : -INFINITY);
  const float e2 = lane < C ? expf(p - pm) : 0.f;
  const float s2 = warp_sum(e2);
  const float sp = e2 / s2;                           // softmax(probs)
  const float pc = __shfl_sync(0xffffffffu, p, cls);
  if (lane == 0) atomicAdd(val, (logf(s2) + pm - pc) / (float)B);
  const float dz = lane < C ? (sp - (lane == cls ? 1.f : 0.f)) * scale / (float)B : 0.f;  // d / d probs
  const float dot = warp_sum(p * dz);
  if (lane < C) dh[b * C + lane] = p * (dz - dot);
}

struct Bump {
  char* base; size_t off;
  float* f(size_t n) {
    off = (off + 255) & ~size_t(255);
    float* p = base ? reinterpret_cast<float*>(base + off) : nullptr;
    off += n * sizeof(float);
    return p;
  }
};

#define KCHECK()                                                                    \
  do {                                                                              \
    cudaError_t e__ = cudaPeekAtLastError();                                        \
    if (e__ != cudaSuccess) return set_error(-3, std::string("train_bw: ") + cudaGetErrorString(e__)); \
  } while (0)

struct SAL { const char* name; int N, S, cin, c[3]; };
struct FPL { const char* name; int nl, N, S, Ca, Cb, c[3]; int fine, coarse; };
const SAL kSAL[4] = {{"sa1", 1024, 1024, 6, {32, 32, 64}}, {"sa2", 1024, 256, 67, {64, 64, 128}}, {"sa3", 256, 64, 131, {128, 128, 256}},
                     {"sa4", 64, 16, 259, {256, 256, 512}}};
const FPL kFPL[4] = {{"fp4", 2, 64, 16, 256, 512, {256, 256, 0}, 3, 4}, {"fp3", 2, 256, 64, 128, 256, {256, 256, 0}, 2, 3},
                     {"fp2", 2, 1024, 256, 64, 256, {256, 128, 0}, 1, 2}, {"fp1", 3, 1024, 1024, 0, 128, {128, 128, 128}, 0, 1}};

}  // namespace

namespace {
__global__ void adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, int64_t n, float lr,
                             float b1, float b2, float eps, float wd, float bc1, float bc2, float gscale) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float gr = g[i] * gscale;
  float pv = p[i] * (1.0f - lr * wd);
  const float mv = b1 * m[i] + (1.0f - b1) * gr;
  const float vv = b2 * v[i] + (1.0f - b2) * gr * gr;
  m[i] = mv;
  v[i] = vv;
  const float denom = sqrtf(vv) / sqrtf(bc2) + eps;
  p[i] = pv - (lr / bc1) * (mv / denom);
}
}  // namespace

// torch.optim.AdamW (no amsgrad): p *= 1 - lr wd; m, v updated; p -= lr / (1 - b1^t) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
int launch_adamw(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2, float eps, float weight_decay,
                 int64_t step, float grad_scale, cudaStream_t st) {
  if (n <= 0) return 0;
  const float bc1 = 1.0f - powf(beta1, (float)step), bc2 = 1.0f - powf(beta2, (float)step);
  adamw_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(p, g, m, v, n, lr, beta1, beta2, eps, weight_decay, bc1, bc2, grad_scale);
  return 1;
}

size_t train_tape_bytes(int B, int n_cats) {
  TrainCtx ctx{};
  ctx.B = B;
  ctx.n_cats = n_cats;
  ctx.tape = nullptr;
  size_t need = 0;
  train_forward_backward(ctx, TrainIO{}, nullptr, &need);
  return need;
}

// size_only != nullptr: only computes the tape size (no launches).
int train_forward_backward(const TrainCtx& ctx, const TrainIO& io, cudaStream_t st, size_t* size_only) {
  const int B = ctx.B, C = B * NOBJ, NC = ctx.n_cats;
  const int64_t rows = (int64_t)B * NPTS;
  const bool dry = size_only != nullptr;
  Bump t{dry ? nullptr : static_cast<char*>(ctx.tape), 0};
  auto W = [&](const std::string& k) { return dry ? (const float*)nullptr : ctx.W(k); };
  auto G = [&](const std::string& k) { return dry ? (float*)nullptr : ctx.G(k); };
  auto mk = [&](const std::string& p, int N, int K) { return Lin{W(p + ".weight"), W(p + ".bias"), G(p + ".weight"), G(p + ".bias"), N, K, K}; };

  // ---------------- tape layout ----------------
  // text / category / translation / time
  float *T0z = t.f(B * 256), *T0a = t.f(B * 256), *T1z = t.f(B * 256), *T1a = t.f(B * 256), *T2z = t.f(B * 128), *enc = t.f(B * 128);
  float *P0z = t.f(B * 64), *P0a = t.f(B * 64), *P1z = t.f(B * 32), *P1a = t.f(B * 32), *P2z = t.f(B * 32), *P2a = t.f(B * 32);
  float *ECz = t.f(C * 32), *ec = t.f(C * 32), *aq = t.f(B * 128), *ak = t.f(C * 128), *asm_ = t.f(B * NHEAD * NOBJ), *attw = t.f(B * NOBJ);
  float *Zt = t.f(C * 160), *TL0z = t.f(C * 128), *TL0a = t.f(C * 128), *TL1z = t.f(C * 16), *tr = t.f(C * 16), *qq = t.f(C * 16);
  float *pet = t.f(B * 128), *TE0z = t.f(B * 128), *TE0a = t.f(B * 128), *ts = t.f(B * 128), *s256 = t.f(B * 256);
  float *U0z = t.f((size_t)B * 256 * 128), *U0a = t.f((size_t)B * 256 * 128), *U1z = t.f((size_t)B * 256 * 512), *U1a = t.f((size_t)B * 256 * 512);
  float *U2z = t.f((size_t)B * 256 * 1024), *U2a = t.f((size_t)B * 256 * 1024), *Ut = t.f((size_t)rows * 256), *CEz = t.f((size_t)rows * 128);
  float* CAT = t.f((size_t)rows * 256);
  // human decoder
  float *H0z = t.f((size_t)rows * 64), *H0a = t.f((size_t)rows * 64), *H1z = t.f((size_t)rows * 64), *H1a = t.f((size_t)rows * 64);
  float *H1s = t.f((size_t)B * 655 * 64), *H2z = t.f((size_t)B * 655 * 64), *H2a = t.f((size_t)B * 655 * 64), *H3 = t.f((size_t)B * 655 * 3);
  float *gnst = t.f(B * 8 * 2 * 3), *hm = t.f((size_t)rows * 3), *pts0 = t.f((size_t)rows * 3);
  // backbone
  float* feat[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  float* dfeat[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  const int fch[5] = {3, 64, 128, 256, 512}, fn[5] = {1024, 1024, 256, 64, 16};
  for (int l = 1; l <= 4; ++l) {
    feat[l] = t.f((size_t)C * fn[l] * fch[l]);
    dfeat[l] = t.f((size_t)C * fn[l] * fch[l]);
  }
  struct SAT { float *P, *Y[3], *A[3], *st[3]; } sat[4];
  for (int l = 0; l < 4; ++l) {
    const SAL& s = kSAL[l];
    const size_t R = (size_t)C * s.S * 32;
    sat[l].P = l > 0 ? t.f((size_t)C * s.N * s.c[0]) : nullptr;
    for (int i = 0; i < 3; ++i) {
      sat[l].Y[i] = t.f(R * s.c[i]);
      sat[l].A[i] = t.f(R * s.c[i]);
      sat[l].st[i] = t.f(2 * s.c[i]);
    }
  }
  struct FPT { float *Pa, *Pb, *Y[3], *A[3], *st[3]; } fpt[4];
  for (int l = 0; l < 4; ++l) {
    const FPL& f = kFPL[l];
    fpt[l].Pa = f.Ca > 0 ? t.f((size_t)C * f.N * f.c[0]) : nullptr;
    fpt[l].Pb = t.f((size_t)C * f.S * f.c[0]);
    for (int i = 0; i < f.nl; ++i) {
      fpt[l].Y[i] = t.f((size_t)C * f.N * f.c[i]);
      fpt[l].A[i] = t.f((size_t)C * f.N * f.c[i]);
      fpt[l].st[i] = t.f(2 * f.c[i]);
    }
  }
  float *Yh = t.f((size_t)C * NPTS * 128), *Ah = t.f((size_t)C * NPTS * 128), *sth = t.f(256), *Fb = t.f((size_t)C * NPTS * 3);
  // scene
  float *p1 = t.f((size_t)C * NPTS * 3), *ctxs = t.f((size_t)3 * C * TRANS), *pa = t.f(C * TRANS), *PWz = t.f((size_t)C * NPTS * 3),
        *pw = t.f((size_t)C * NPTS * 3), *pcd = t.f((size_t)rows * 3);
  // x0 network
  float *z = t.f((size_t)rows * 3), *E0z = t.f((size_t)rows * 64), *E0a = t.f((size_t)rows * 64), *E1z = t.f((size_t)rows * 128);
  float *C0z = t.f((size_t)rows * 192), *C0a = t.f((size_t)rows * 192), *C1z = t.f((size_t)rows * 128), *C1a = t.f((size_t)rows * 128);
  float *F0z = t.f((size_t)rows * 64), *F0a = t.f((size_t)rows * 64), *F1z = t.f((size_t)rows * 3), *x0 = t.f((size_t)rows * 3);
  // gradient scratch: two ping-pong buffers as large as the widest layer (sa1: 32768 x 64 per cloud; upsampler: B*256*1024)
  size_t gmax = (size_t)C * 32768 * 64;
  if ((size_t)B * 256 * 1024 > gmax) gmax = (size_t)B * 256 * 1024;
  if ((size_t)rows * 256 > gmax) gmax = (size_t)rows * 256;
  float *g0 = t.f(gmax), *g1 = t.f(gmax), *g2 = t.f((size_t)C * 1024 * 256);
  float *d_pcd = t.f((size_t)rows * 3), *d_enc = t.f(B * 128), *d_ec = t.f(C * 32), *d_Fb = t.f((size_t)C * NPTS * 3), *d_p1 = t.f((size_t)C * NPTS * 3);
  float *d_P2a = t.f(B * 32), *d_P1 = t.f(B * 32), *d_P0 = t.f(B * 64), *d_s = t.f(B * 256), *d_qq = t.f(C * 16), *d_attw = t.f(B * 16);
  float *d_tr = t.f(C * 16), *d_q = t.f(B * 128), *d_k = t.f(C * 128), *d_t1 = t.f(B * 256), *d_t0 = t.f(B * 256), *d_ts = t.f(B * 128),
        *d_te = t.f(B * 128);
  double* dsum = reinterpret_cast<double*>(t.f(4 * 1024));
  float* vals = t.f(64);
  float* bnmeans = t.f(2 * 1024);
  float* wt_scratch = t.f(1024 * 1024);
  if (dry) {
    *size_only = t.off + 256;
    return 0;
  }
  if (t.off + 256 > ctx.tape_bytes) return set_error(-1, "train tape too small: need " + std::to_string(t.off + 256));

  g_wt_scratch = wt_scratch;
  const int64_t shards = ctx.allreduce ? ctx.Bg / ctx.B : 1;
  int hook_err = 0;
  auto blocks = [](int64_t n) { return (unsigned)((n + 255) / 256); };
  auto zero = [&](float* p, size_t n) { cudaMemsetAsync(p, 0, n * sizeof(float), st); };

  // BatchNorm forward / backward helpers (layer key, rows M, channels N)
  auto bn_fw = [&](const std::string& key, const float* Y, int64_t M, int N, float* stat, float* A, const float* mask) {
    launch_col_stats(Y, M, N, dsum, st);
    if (ctx.allreduce && ctx.allreduce(ctx.allreduce_ctx, dsum, 2 * N) != 0) hook_err = 1;
    bn_finalize_kernel<<<(N + 127) / 128, 128, 0, st>>>(dsum, (double)(M * shards), N, W(key + ".weight"), W(key + ".bias"), 1e-5f, stat);
    bn_fw_kernel<<<blocks(M * N / 4), 256, 0, st>>>(Y, M, N, stat, W(key + ".weight"), W(key + ".bias"), A, mask, NPTS);
  };
  auto bn_bw = [&](const std::string& key, float* dA, const float* Y, const float* A, int64_t M, int N, const float* stat, const float* mask) {
    cudaMemsetAsync(dsum, 0, sizeof(double) * 2 * N, st);
    int64_t g = (M + 63) / 64;
    if (g > 148 * 8) g = 148 * 8;
    bn_bw_reduce_kernel<<<(unsigned)g, 256, 0, st>>>(dA, Y, A, M, N, stat, mask, NPTS, dsum);
    bn_param_grad_kernel<<<(N + 127) / 128, 128, 0, st>>>(dsum, N, G(key + ".weight"), G(key + ".bias"));  // local sums: like DDP + SyncBN
    if (ctx.allreduce && ctx.allreduce(ctx.allreduce_ctx, dsum, 2 * N) != 0) hook_err = 1;
    bn_bw_means_kernel<<<(2 * N + 127) / 128, 128, 0, st>>>(dsum, (double)(M * shards), N, bnmeans);
    bn_bw_apply_kernel<<<blocks(M * N / 4), 256, 0, st>>>(dA, Y, A, M, N, stat, W(key + ".weight"), bnmeans, mask, NPTS);
  };

  // =====================================================================================================================
  // forward (taped)
  // =====================================================================================================================
  const Lin et0 = mk("embed_text.0", 256, CLIP), et2 = mk("embed_text.2", 256, 256), et4 = mk("embed_text.4", 128, 256);
  lin_fw(st, io.text, CLIP, B, et0, T0z, 256); act_fw(st, T0z, 256, T0a, 256, B, 256, ACT_GELU);
  lin_fw(st, T0a, 256, B, et2, T1z, 256);      act_fw(st, T1z, 256, T1a, 256, B, 256, ACT_GELU);
  lin_fw(st, T1a, 256, B, et4, T2z, 128);      act_fw(st, T2z, 128, enc, 128, B, 128, ACT_GELU);
  const Lin pc0 = mk("predict_cat.0", 64, 128), pc2 = mk("predict_cat.2", 32, 64), pc4 = mk("predict_cat.4", NC, 32);
  lin_fw(st, enc, 128, B, pc0, P0z, 64); act_fw(st, P0z, 64, P0a, 64, B, 64, ACT_GELU);
  lin_fw(st, P0a, 64, B, pc2, P1z, 32);  act_fw(st, P1z, 32, P1a, 32, B, 32, ACT_GELU);
  lin_fw(st, P1a, 32, B, pc4, P2z, NC);  act_fw(st, P2z, NC, P2a, NC, B, NC, ACT_GELU);
  const Lin ecl = mk("embed_cat.0", 32, NC);
  lin_fw(st, io.cats, NC, C, ecl, ECz, 32); act_fw(st, ECz, 32, ec, 32, C, 32, ACT_GELU);
  Lin aql{W("attn_layer.q_proj_weight"), W("attn_layer.in_proj_bias"), G("attn_layer.q_proj_weight"), G("attn_layer.in_proj_bias"), 128, 128, 128};
  Lin akl{W("attn_layer.k_proj_weight"), W("attn_layer.in_proj_bias") + 128, G("attn_layer.k_proj_weight"), G("attn_layer.in_proj_bias") + 128, 128, 32, 32};
  lin_fw(st, enc, 128, B, aql, aq, 128);
  lin_fw(st, ec, 32, C, akl, ak, 128);
  attw_fw_kernel<<<B, 32, 0, st>>>(aq, ak, io.mask_global, B, ctx.Bg, ctx.b_off, asm_, attw);
  const Lin tl0 = mk("translation_layer.0", 128, 160), tl2 = mk("translation_layer.2", TRANS, 128);
  concat_trans_kernel<<<blocks(C * 160), 256, 0, st>>>(ec, enc, B, Zt);
  lin_fw(st, Zt, 160, C, tl0, TL0z, 128);   act_fw(st, TL0z, 128, TL0a, 128, C, 128, ACT_GELU);
  lin_fw(st, TL0a, 128, C, tl2, TL1z, TRANS); act_fw(st, TL1z, TRANS, tr, TRANS, C, TRANS, ACT_GELU);
  Lin q12{W("pcd_attention.q_proj_weight"), W("pcd_attention.in_proj_bias"), G("pcd_attention.q_proj_weight"), G("pcd_attention.in_proj_bias"), TRANS, TRANS, TRANS};
  lin_fw(st, tr, TRANS, C, q12, qq, TRANS);
  const Lin te0 = mk("embed_timestep.time_embed.0", 128, 128), te2 = mk("embed_timestep.time_embed.2", 128, 128);
  gather_pe_kernel<<<blocks(B * 128), 256, 0, st>>>(W("embed_timestep.sequence_pos_encoder.pe"), io.t, B, pet);
  lin_fw(st, pet, 128, B, te0, TE0z, 128); act_fw(st, TE0z, 128, TE0a, 128, B, 128, ACT_SILU);
  lin_fw(st, TE0a, 128, B, te2, ts, 128);
  concat2_kernel<<<blocks(B * 256), 256, 0, st>>>(ts, enc, B, s256);
  const Lin u0 = mk("upsampling_layer.0", 128, 1), u2 = mk("upsampling_layer.2", 512, 128), u4 = mk("upsampling_layer.4", NPTS, 512);
  const int SR = B * 256;
  lin_fw(st, s256, 1, SR, u0, U0z, 128);  act_fw(st, U0z, 128, U0a, 128, SR, 128, ACT_GELU);
  lin_fw(st, U0a, 128, SR, u2, U1z, 512); act_fw(st, U1z, 512, U1a, 512, SR, 512, ACT_GELU);
  lin_fw(st, U1a, 512, SR, u4, U2z, 1024); act_fw(st, U2z, 1024, U2a, 1024, SR, 1024, ACT_GELU);
  transpose(st, U2a, Ut, B, 256, 1024);
  const Lin cel = mk("combine_extraction.0", 128, 256);
  lin_fw(st, Ut, 256, (int)rows, cel, CEz, 128);
  act_fw(st, CEz, 128, CAT + 128, 256, rows, 128, ACT_GELU);
  KCHECK();
  // human decoder on the points of slot 0
  slice_pts_kernel<<<blocks(rows * 3), 256, 0, st>>>(io.objs, B, NOBJ * NPTS, NPTS, 3, pts0, 0);
  const Lin h0 = mk("human_backbone.de_spiral.0.conv.layer", 64, 3), h1l = mk("human_backbone.de_spiral.1.conv.layer", 64, 64),
            h2l = mk("human_backbone.de_spiral.2.conv.layer", 64, 64), h3l = mk("human_backbone.de_spiral.3.layer", 3, 64);
  lin_fw(st, pts0, 3, (int)rows, h0, H0z, 64);
  gn_fw_kernel<<<B * 8, 256, 0, st>>>(H0z, NPTS, NPTS, W("human_backbone.de_spiral.0.norm.weight"), W("human_backbone.de_spiral.0.norm.bias"), 1e-5f, gnst, H0a);
  lin_fw(st, H0a, 64, (int)rows, h1l, H1z, 64);
  gn_fw_kernel<<<B * 8, 256, 0, st>>>(H1z, NPTS, NPTS, W("human_backbone.de_spiral.1.norm.weight"), W("human_backbone.de_spiral.1.norm.bias"), 1e-5f, gnst + B * 16, H1a);
  slice_pts_kernel<<<blocks((int64_t)B * 655 * 64), 256, 0, st>>>(H1a, B, NPTS, 655, 64, H1s, 0);
  lin_fw(st, H1s, 64, B * 655, h2l, H2z, 64);
  gn_fw_kernel<<<B * 8, 256, 0, st>>>(H2z, 655, 655, W("human_backbone.de_spiral.2.norm.weight"), W("human_backbone.de_spiral.2.norm.bias"), 1e-5f, gnst + B * 32, H2a);
  lin_fw(st, H2a, 64, B * 655, h3l, H3, 3);
  upsample2_fw_kernel<<<blocks(rows * 3), 256, 0, st>>>(H3, B, hm);
  KCHECK();
  // PointNet++ (selection results are inputs)
  const float* xyzl[5] = {io.objs, io.xyz[1], io.xyz[2], io.xyz[3], io.xyz[4]};
  const float* featl[5] = {io.objs, feat[1], feat[2], feat[3], feat[4]};
  for (int l = 0; l < 4; ++l) {
    const SAL& s = kSAL[l];
    const std::string p = std::string("pcd_backbone.") + s.name;
    const int64_t R = (int64_t)C * s.S * 32;
    const float* W0 = W(p + ".mlp_convs.0.weight");
    if (l > 0) {
      Lin pl{W0 + 3, W(p + ".mlp_convs.0.bias"), nullptr, nullptr, s.c[0], s.cin - 3, s.cin};
      lin_fw(st, featl[l], s.cin - 3, C * s.N, pl, sat[l].P, s.c[0]);
    }
    sa_gather_fw_kernel<<<blocks(R * s.c[0]), 256, 0, st>>>(sat[l].P, W0, s.cin, W(p + ".mlp_convs.0.bias"), xyzl[l], xyzl[l + 1], io.grp[l], R, s.N,
                                                            s.S, s.c[0], sat[l].Y[0]);
    bn_fw(p + ".mlp_bns.0", sat[l].Y[0], R, s.c[0], sat[l].st[0], sat[l].A[0], nullptr);
    for (int i = 1; i < 3; ++i) {
      const std::string ci = p + ".mlp_convs." + std::to_string(i);
      Lin L{W(ci + ".weight"), W(ci + ".bias"), nullptr, nullptr, s.c[i], s.c[i - 1], s.c[i - 1]};
      lin_fw(st, sat[l].A[i - 1], s.c[i - 1], (int)R, L, sat[l].Y[i], s.c[i]);
      bn_fw(p + ".mlp_bns." + std::to_string(i), sat[l].Y[i], R, s.c[i], sat[l].st[i], sat[l].A[i], nullptr);
    }
    maxpool_fw_kernel<<<blocks((R / 32) * s.c[2]), 256, 0, st>>>(sat[l].A[2], R / 32, s.c[2], feat[l + 1]);
    KCHECK();
  }
  const float* coarse = feat[4];
  for (int l = 0; l < 4; ++l) {
    const FPL& f = kFPL[l];
    const std::string p = std::string("pcd_backbone.") + f.name;
    const int64_t M = (int64_t)C * f.N;
    const float* W0 = W(p + ".mlp_convs.0.weight");
    const int cin = f.Ca + f.Cb;
    if (f.Ca > 0) {
      Lin la{W0, W(p + ".mlp_convs.0.bias"), nullptr, nullptr, f.c[0], f.Ca, cin};
      lin_fw(st, featl[f.fine], f.Ca, (int)M, la, fpt[l].Pa, f.c[0]);
    }
    Lin lb{W0 + f.Ca, nullptr, nullptr, nullptr, f.c[0], f.Cb, cin};
    lin_fw(st, coarse, f.Cb, C * f.S, lb, fpt[l].Pb, f.c[0]);
    fp_combine_fw_kernel<<<blocks(M * f.c[0]), 256, 0, st>>>(fpt[l].Pa, W(p + ".mlp_convs.0.bias"), fpt[l].Pb, io.nn_idx[l], io.nn_w[l], M, f.N, f.S,
                                                             f.c[0], fpt[l].Y[0]);
    bn_fw(p + ".mlp_bns.0", fpt[l].Y[0], M, f.c[0], fpt[l].st[0], fpt[l].A[0], nullptr);
    for (int i = 1; i < f.nl; ++i) {
      const std::string ci = p + ".mlp_convs." + std::to_string(i);
      Lin L{W(ci + ".weight"), W(ci + ".bias"), nullptr, nullptr, f.c[i], f.c[i - 1], f.c[i - 1]};
      lin_fw(st, fpt[l].A[i - 1], f.c[i - 1], (int)M, L, fpt[l].Y[i], f.c[i]);
      bn_fw(p + ".mlp_bns." + std::to_string(i), fpt[l].Y[i], M, f.c[i], fpt[l].st[i], fpt[l].A[i], nullptr);
    }
    coarse = fpt[l].A[f.nl - 1];
    KCHECK();
  }
  {
    const int64_t M = (int64_t)C * NPTS;
    Lin c1{W("pcd_backbone.conv1.weight"), W("pcd_backbone.conv1.bias"), nullptr, nullptr, 128, 128, 128};
    lin_fw(st, coarse, 128, (int)M, c1, Yh, 128);
    bn_fw("pcd_backbone.bn1", Yh, M, 128, sth, Ah, io.drop_mask);
    Lin c2{W("pcd_backbone.conv2.weight"), W("pcd_backbone.conv2.bias"), nullptr, nullptr, 3, 128, 128};
    lin_fw(st, Ah, 128, (int)M, c2, Fb, 3);
  }
  // scene branch
  scramble1_fw_kernel<<<blocks((int64_t)C * NPTS * 3), 256, 0, st>>>(Fb, attw, B, p1);
  PaW pw_{W("pcd_attention.k_proj_weight"), W("pcd_attention.v_proj_weight"), W("pcd_attention.in_proj_bias"), W("pcd_attention.out_proj.weight"),
          W("pcd_attention.out_proj.bias"), W("point_wise_trans_layer.0.weight"), W("point_wise_trans_layer.0.bias"),
          G("pcd_attention.k_proj_weight"), G("pcd_attention.v_proj_weight"), G("pcd_attention.in_proj_bias"), G("pcd_attention.out_proj.weight"),
          G("pcd_attention.out_proj.bias"), G("point_wise_trans_layer.0.weight"), G("point_wise_trans_layer.0.bias")};
  pattn_fw_kernel<<<C, 256, 0, st>>>(pw_, p1, qq, ctxs, pa, PWz, pw);
  launch_scene_mix(pw, hm, io.mask_global, B, ctx.Bg, ctx.b_off, pcd, st);
  // x0 network
  qsample_add_kernel<<<blocks(rows * 3), 256, 0, st>>>(io.x_start, io.t, io.noise, ctx.sched_sa, ctx.sched_s1a, pcd, rows * 3, z);
  const Lin e0 = mk("input_process.pose_embedding.0", 64, 3), e2 = mk("input_process.pose_embedding.2", 128, 64),
            c0 = mk("input_process.combination_extraction.0", 192, 256), c2l = mk("input_process.combination_extraction.2", 128, 192),
            f0 = mk("output_process.pose_final.0", 64, 128), f2 = mk("output_process.pose_final.2", 3, 64);
  lin_fw(st, z, 3, (int)rows, e0, E0z, 64);      act_fw(st, E0z, 64, E0a, 64, rows, 64, ACT_SIGMOID);
  lin_fw(st, E0a, 64, (int)rows, e2, E1z, 128);  act_fw(st, E1z, 128, CAT, 256, rows, 128, ACT_SIGMOID);
  lin_fw(st, CAT, 256, (int)rows, c0, C0z, 192); act_fw(st, C0z, 192, C0a, 192, rows, 192, ACT_SIGMOID);
  lin_fw(st, C0a, 192, (int)rows, c2l, C1z, 128); act_fw(st, C1z, 128, C1a, 128, rows, 128, ACT_SIGMOID);
  lin_fw(st, C1a, 128, (int)rows, f0, F0z, 64);  act_fw(st, F0z, 64, F0a, 64, rows, 64, ACT_GELU);
  lin_fw(st, F0a, 64, (int)rows, f2, F1z, 3);    act_fw(st, F1z, 3, x0, 3, rows, 3, ACT_GELU);
  KCHECK();

  // =====================================================================================================================
  // losses + backward
  // =====================================================================================================================
  zero(vals, 64);
  float* d_x0 = g0;
  zero(d_x0, rows * 3);
  chamfer_bw_kernel<<<dim3(NPTS / 256, B, 2), 256, 0, st>>>(x0, io.x_start, B, io.g_mse, d_x0, vals);
  cat_fw_bw_kernel<<<(B + 7) / 8, 256, 0, st>>>(P2a, io.target_cat, B, NC, io.g_cat * io.lambda_cat, d_P2a, vals + 1);
  if (io.losses_out) cudaMemcpyAsync(io.losses_out, vals, 2 * sizeof(float), cudaMemcpyDeviceToDevice, st);
  // predict_cat (its input enc is detached: model/sdm.py:156)
  act_bw(st, d_P2a, NC, P2z, NC, B, NC, ACT_GELU);
  lin_bw(st, P1a, 32, B, pc4, d_P2a, NC, d_P1, 32);
  act_bw(st, d_P1, 32, P1z, 32, B, 32, ACT_GELU);
  lin_bw(st, P0a, 64, B, pc2, d_P1, 32, d_P0, 64);
  act_bw(st, d_P0, 64, P0z, 64, B, 64, ACT_GELU);
  lin_bw(st, enc, 128, B, pc0, d_P0, 64, nullptr, 0);
  // x0 network
  act_bw(st, d_x0, 3, F1z, 3, rows, 3, ACT_GELU);
  lin_bw(st, F0a, 64, (int)rows, f2, d_x0, 3, g1, 64);
  act_bw(st, g1, 64, F0z, 64, rows, 64, ACT_GELU);
  lin_bw(st, C1a, 128, (int)rows, f0, g1, 64, g0, 128);
  act_bw(st, g0, 128, C1z, 128, rows, 128, ACT_SIGMOID);
  lin_bw(st, C0a, 192, (int)rows, c2l, g0, 128, g1, 192);
  act_bw(st, g1, 192, C0z, 192, rows, 192, ACT_SIGMOID);
  lin_bw(st, CAT, 256, (int)rows, c0, g1, 192, g0, 256);   // g0 = dCAT [rows,256]: [:, :128] d pose, [:, 128:] d emb
  act_bw(st, g0, 256, E1z, 128, rows, 128, ACT_SIGMOID);
  lin_bw(st, E0a, 64, (int)rows, e2, g0, 256, g1, 64);
  act_bw(st, g1, 64, E0z, 64, rows, 64, ACT_SIGMOID);
  lin_bw(st, z, 3, (int)rows, e0, g1, 64, d_pcd, 3);       // d z = d pcd_out (x_t carries no gradient)
  // embedding: combine_extraction <- upsampler <- [ts || enc]
  act_bw(st, g0 + 128, 256, CEz, 128, rows, 128, ACT_GELU);
  lin_bw(st, Ut, 256, (int)rows, cel, g0 + 128, 256, g1, 256);      // g1 = dUt [B,1024,256]
  transpose(st, g1, g2, B, 1024, 256);                               // g2 = dU2a [B,256,1024]
  act_bw(st, g2, 1024, U2z, 1024, SR, 1024, ACT_GELU);
  lin_bw(st, U1a, 512, SR, u4, g2, 1024, g1, 512);
  act_bw(st, g1, 512, U1z, 512, SR, 512, ACT_GELU);
  lin_bw(st, U0a, 128, SR, u2, g1, 512, g0, 128);
  act_bw(st, g0, 128, U0z, 128, SR, 128, ACT_GELU);
  lin_bw(st, s256, 1, SR, u0, g0, 128, d_s, 1);   // d_s [B,256]
  // d ts = d_s[:, :128] -> time_embed ; d enc = d_s[:, 128:]
  zero(d_enc, B * 128);
  zero(d_ec, C * 32);
  {
    cudaMemcpy2DAsync(d_ts, 128 * sizeof(float), d_s, 256 * sizeof(float), 128 * sizeof(float), B, cudaMemcpyDeviceToDevice, st);
    cudaMemcpy2DAsync(d_enc, 128 * sizeof(float), d_s + 128, 256 * sizeof(float), 128 * sizeof(float), B, cudaMemcpyDeviceToDevice, st);
    lin_bw(st, TE0a, 128, B, te2, d_ts, 128, d_te, 128);
    act_bw(st, d_te, 128, TE0z, 128, B, 128, ACT_SILU);
    lin_bw(st, pet, 128, B, te0, d_te, 128, nullptr, 0);
  }
  KCHECK();
  // pcd_out = (scene + hm) / 2
  float *d_pw = g0, *d_hm = g1;
  scene_bw_kernel<<<blocks(rows * 3), 256, 0, st>>>(d_pcd, io.mask_global, B, ctx.Bg, ctx.b_off, d_pw, d_hm);
  // human decoder
  {
    float *d_h3 = g2, *d_h2 = g2 + (size_t)B * 655 * 4, *d_h1s = d_h2 + (size_t)B * 655 * 64, *d_h1 = d_h1s + (size_t)B * 655 * 64,
          *d_h0 = d_h1 + (size_t)rows * 64;
    upsample2_bw_kernel<<<blocks((int64_t)B * 655 * 3), 256, 0, st>>>(d_hm, B, d_h3);
    lin_bw(st, H2a, 64, B * 655, h3l, d_h3, 3, d_h2, 64);
    gn_bw_kernel<<<B * 8, 256, 0, st>>>(d_h2, H2z, H2a, 655, 655, W("human_backbone.de_spiral.2.norm.weight"), gnst + B * 32,
                                        G("human_backbone.de_spiral.2.norm.weight"), G("human_backbone.de_spiral.2.norm.bias"));
    lin_bw(st, H1s, 64, B * 655, h2l, d_h2, 64, d_h1s, 64);
    slice_pts_kernel<<<blocks(rows * 64), 256, 0, st>>>(d_h1s, B, NPTS, 655, 64, d_h1, 1);
    gn_bw_kernel<<<B * 8, 256, 0, st>>>(d_h1, H1z, H1a, NPTS, NPTS, W("human_backbone.de_spiral.1.norm.weight"), gnst + B * 16,
                                        G("human_backbone.de_spiral.1.norm.weight"), G("human_backbone.de_spiral.1.norm.bias"));
    lin_bw(st, H0a, 64, (int)rows, h1l, d_h1, 64, d_h0, 64);
    gn_bw_kernel<<<B * 8, 256, 0, st>>>(d_h0, H0z, H0a, NPTS, NPTS, W("human_backbone.de_spiral.0.norm.weight"), gnst,
                                        G("human_backbone.de_spiral.0.norm.weight"), G("human_backbone.de_spiral.0.norm.bias"));
    lin_bw(st, pts0, 3, (int)rows, h0, d_h0, 64, nullptr, 0);
  }
  KCHECK();
  // scene branch: pointwise + attention
  act_bw(st, d_pw, 3, PWz, 3, (int64_t)C * NPTS, 3, ACT_GELU);
  pattn_bw_kernel<<<C, 256, 0, st>>>(pw_, p1, qq, ctxs, pa, d_pw, d_p1, d_qq);
  scramble1_bw_kernel<<<C, 256, 0, st>>>(d_p1, Fb, attw, d_Fb, d_attw);
  // qq <- tr <- translation_layer <- [ec || enc]
  {
    float *d_tl0 = g1, *d_zt = g1 + (size_t)C * 128;
    lin_bw(st, tr, TRANS, C, q12, d_qq, TRANS, d_tr, TRANS);
    act_bw(st, d_tr, TRANS, TL1z, TRANS, C, TRANS, ACT_GELU);
    lin_bw(st, TL0a, 128, C, tl2, d_tr, TRANS, d_tl0, 128);
    act_bw(st, d_tl0, 128, TL0z, 128, C, 128, ACT_GELU);
    lin_bw(st, Zt, 160, C, tl0, d_tl0, 128, d_zt, 160);
    concat_trans_bw_kernel<<<blocks(C * 32 > B * 128 ? C * 32 : B * 128), 256, 0, st>>>(d_zt, B, d_ec, d_enc);
    // attention weights <- q(enc), k(ec)
    attw_bw_kernel<<<B, 32, 0, st>>>(d_attw, asm_, aq, ak, B, d_q, d_k);
    lin_bw(st, enc, 128, B, aql, d_q, 128, d_enc, 128, true);
    lin_bw(st, ec, 32, C, akl, d_k, 128, d_ec, 32, true);
    act_bw(st, d_ec, 32, ECz, 32, C, 32, ACT_GELU);
    lin_bw(st, io.cats, NC, C, ecl, d_ec, 32, nullptr, 0);
    // text MLP
    act_bw(st, d_enc, 128, T2z, 128, B, 128, ACT_GELU);
    lin_bw(st, T1a, 256, B, et4, d_enc, 128, d_t1, 256);
    act_bw(st, d_t1, 256, T1z, 256, B, 256, ACT_GELU);
    lin_bw(st, T0a, 256, B, et2, d_t1, 256, d_t0, 256);
    act_bw(st, d_t0, 256, T0z, 256, B, 256, ACT_GELU);
    lin_bw(st, io.text, CLIP, B, et0, d_t0, 256, nullptr, 0);
  }
  KCHECK();
  // PointNet++ backward
  for (int l = 1; l <= 4; ++l) zero(dfeat[l], (size_t)C * fn[l] * fch[l]);
  {
    const int64_t M = (int64_t)C * NPTS;
    Lin c2{W("pcd_backbone.conv2.weight"), W("pcd_backbone.conv2.bias"), G("pcd_backbone.conv2.weight"), G("pcd_backbone.conv2.bias"), 3, 128, 128};
    lin_bw(st, Ah, 128, (int)M, c2, d_Fb, 3, g0, 128);
    bn_bw("pcd_backbone.bn1", g0, Yh, Ah, M, 128, sth, io.drop_mask);
    Lin c1{W("pcd_backbone.conv1.weight"), W("pcd_backbone.conv1.bias"), G("pcd_backbone.conv1.weight"), G("pcd_backbone.conv1.bias"), 128, 128, 128};
    lin_bw(st, fpt[3].A[2], 128, (int)M, c1, g0, 128, g1, 128);  // g1 = d(fp1 output)
  }
  float* dcur = g1;   // gradient wrt the current FP level's output
  float* dother = g0;
  for (int l = 3; l >= 0; --l) {
    const FPL& f = kFPL[l];
    const std::string p = std::string("pcd_backbone.") + f.name;
    const int64_t M = (int64_t)C * f.N;
    const int cin = f.Ca + f.Cb;
    for (int i = f.nl - 1; i >= 1; --i) {
      bn_bw(p + ".mlp_bns." + std::to_string(i), dcur, fpt[l].Y[i], fpt[l].A[i], M, f.c[i], fpt[l].st[i], nullptr);
      const std::string ci = p + ".mlp_convs." + std::to_string(i);
      Lin L{W(ci + ".weight"), W(ci + ".bias"), G(ci + ".weight"), G(ci + ".bias"), f.c[i], f.c[i - 1], f.c[i - 1]};
      lin_bw(st, fpt[l].A[i - 1], f.c[i - 1], (int)M, L, dcur, f.c[i], dother, f.c[i - 1]);
      std::swap(dcur, dother);
    }
    bn_bw(p + ".mlp_bns.0", dcur, fpt[l].Y[0], fpt[l].A[0], M, f.c[0], fpt[l].st[0], nullptr);   // dcur = dY1 [M, C1]
    // Y1 = Pa (+ b0) + interp(Pb):  d Pa = dY1;  d Pb scattered
    const float* W0 = W(p + ".mlp_convs.0.weight");
    float* gW0 = G(p + ".mlp_convs.0.weight");
    float* dPb = dother;
    zero(dPb, (size_t)C * f.S * f.c[0]);
    fp_combine_bw_kernel<<<blocks(M * f.c[0]), 256, 0, st>>>(dcur, io.nn_idx[l], io.nn_w[l], M, f.N, f.S, f.c[0], dPb);
    if (f.Ca > 0) {
      Lin la{W0, nullptr, gW0, G(p + ".mlp_convs.0.bias"), f.c[0], f.Ca, cin};
      lin_bw(st, featl[f.fine], f.Ca, (int)M, la, dcur, f.c[0], dfeat[f.fine], f.Ca, true);   // skip connection -> d feat[fine]
    } else {
      colsum(st, dcur, f.c[0], M, f.c[0], G(p + ".mlp_convs.0.bias"));
    }
    // coarse input: feat[4] for fp4, else the previous FP level's output
    const float* cin_act = l == 0 ? feat[4] : fpt[l - 1].A[kFPL[l - 1].nl - 1];
    Lin lb{W0 + f.Ca, nullptr, gW0 + f.Ca, nullptr, f.c[0], f.Cb, cin};
    if (l == 0) {
      lin_bw(st, cin_act, f.Cb, C * f.S, lb, dPb, f.c[0], dfeat[4], f.Cb, true);
    } else {
      lin_bw(st, cin_act, f.Cb, C * f.S, lb, dPb, f.c[0], dcur, f.Cb);   // dcur <- gradient wrt the coarser FP output (dY1 is consumed)
    }
    KCHECK();
  }
  for (int l = 3; l >= 0; --l) {
    const SAL& s = kSAL[l];
    const std::string p = std::string("pcd_backbone.") + s.name;
    const int64_t R = (int64_t)C * s.S * 32;
    float *da = g0, *db = g1;
    maxpool_bw_kernel<<<blocks((R / 32) * s.c[2]), 256, 0, st>>>(dfeat[l + 1], sat[l].A[2], feat[l + 1], R / 32, s.c[2], da);
    for (int i = 2; i >= 1; --i) {
      bn_bw(p + ".mlp_bns." + std::to_string(i), da, sat[l].Y[i], sat[l].A[i], R, s.c[i], sat[l].st[i], nullptr);
      const std::string ci = p + ".mlp_convs." + std::to_string(i);
      Lin L{W(ci + ".weight"), W(ci + ".bias"), G(ci + ".weight"), G(ci + ".bias"), s.c[i], s.c[i - 1], s.c[i - 1]};
      lin_bw(st, sat[l].A[i - 1], s.c[i - 1], (int)R, L, da, s.c[i], db, s.c[i - 1]);
      std::swap(da, db);
    }
    bn_bw(p + ".mlp_bns.0", da, sat[l].Y[0], sat[l].A[0], R, s.c[0], sat[l].st[0], nullptr);   // da = dY1 [R, C1]
    float* gW0 = G(p + ".mlp_convs.0.weight");
    float* dP = nullptr;
    if (l > 0) {
      dP = g2;
      zero(dP, (size_t)C * s.N * s.c[0]);
    }
    const int rpb = 256;
    sa_gather_bw_kernel<<<(unsigned)((R + rpb - 1) / rpb), s.c[0] < 256 ? s.c[0] : 256, 0, st>>>(da, xyzl[l], xyzl[l + 1], io.grp[l], R, s.N, s.S, s.c[0],
                                                                                               s.cin, dP, gW0, G(p + ".mlp_convs.0.bias"), rpb);
    if (l > 0) {
      Lin pl{W(p + ".mlp_convs.0.weight") + 3, nullptr, gW0 + 3, G(p + ".mlp_convs.0.bias"), s.c[0], s.cin - 3, s.cin};
      lin_bw(st, featl[l], s.cin - 3, C * s.N, pl, dP, s.c[0], dfeat[l], s.cin - 3, true);
    }
    KCHECK();
  }
  if (hook_err) return set_error(-1, "all-reduce hook failed");
  if (io.x0_out) cudaMemcpyAsync(io.x0_out, x0, sizeof(float) * rows * 3, cudaMemcpyDeviceToDevice, st);
  KCHECK();
  return 0;
}

}  // namespace lsdm
