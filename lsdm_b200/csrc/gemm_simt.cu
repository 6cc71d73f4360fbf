// fp32 CUDA-core implementation of the generic linear-layer GEMM (gemm.cuh).
// 128 x BN output tile per CTA, BK = 16, 256 threads, 8 x (BN/16) register tile per thread,
// register-prefetched global loads, transposed shared-memory staging (conflict-free reads).
// This is the bit-faithful fp32 path; the tcgen05 TF32 path (gemm_tc.cu) implements the same interface.
#include "gemm.cuh"

namespace lsdm {

namespace {

constexpr int BM = 128;
constexpr int BK = 16;
constexpr int NT = 256;
constexpr int TM = 8;

template <int BN>
__global__ void __launch_bounds__(NT, 2) gemm_simt_kernel(GemmArgs g) {
  constexpr int TN = BN / 16;
  constexpr int WF4 = BN * BK / 4 / NT;  // float4 loads of W per thread (1 for BN=64, 2 for BN=128)
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Ws[BK][BN + 4];
  __shared__ float red[16][BN];

  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const float* __restrict__ A = g.A + (int64_t)blockIdx.z * g.strideA;
  const float* __restrict__ W = g.W + (int64_t)blockIdx.z * g.strideW;
  float* __restrict__ C = g.C + (int64_t)blockIdx.z * g.strideC;

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.0f;

  float4 ra[2], rw[WF4];
  auto load_tiles = [&](int k0) {
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      int idx = tid + r * NT;
      int row = idx >> 2, kq = idx & 3;
      int m = m0 + row;
      ra[r] = (m < g.M) ? *reinterpret_cast<const float4*>(A + (int64_t)m * g.lda + k0 + kq * 4)
                        : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int r = 0; r < WF4; ++r) {
      int idx = tid + r * NT;
      int row = idx >> 2, kq = idx & 3;
      int n = n0 + row;
      rw[r] = (n < g.N) ? *reinterpret_cast<const float4*>(W + (int64_t)n * g.ldw + k0 + kq * 4)
                        : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  auto store_tiles = [&]() {
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      int idx = tid + r * NT;
      int row = idx >> 2, kq = idx & 3;
      As[kq * 4 + 0][row] = ra[r].x;
      As[kq * 4 + 1][row] = ra[r].y;
      As[kq * 4 + 2][row] = ra[r].z;
      As[kq * 4 + 3][row] = ra[r].w;
    }
#pragma unroll
    for (int r = 0; r < WF4; ++r) {
      int idx = tid + r * NT;
      int row = idx >> 2, kq = idx & 3;
      Ws[kq * 4 + 0][row] = rw[r].x;
      Ws[kq * 4 + 1][row] = rw[r].y;
      Ws[kq * 4 + 2][row] = rw[r].z;
      Ws[kq * 4 + 3][row] = rw[r].w;
    }
  };

  load_tiles(0);
  for (int k0 = 0; k0 < g.K; k0 += BK) {
    __syncthreads();
    store_tiles();
    __syncthreads();
    if (k0 + BK < g.K) load_tiles(k0 + BK);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[TM], w[TN];
      float4 a0 = *reinterpret_cast<const float4*>(&As[k][ty * TM]);
      float4 a1 = *reinterpret_cast<const float4*>(&As[k][ty * TM + 4]);
      a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w;
      a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
#pragma unroll
      for (int j = 0; j < TN; j += 4) {
        float4 wv = *reinterpret_cast<const float4*>(&Ws[k][tx * TN + j]);
        w[j] = wv.x; w[j + 1] = wv.y; w[j + 2] = wv.z; w[j + 3] = wv.w;
      }
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
    }
  }

  // epilogue: bias + activation
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    int m = m0 + ty * TM + i;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      int n = n0 + tx * TN + j;
      float v = acc[i][j];
      if (g.bias_mode == 1 && n < g.N) v += g.bias[n];
      if (g.bias_mode == 2 && m < g.M) v += g.bias[m];
      acc[i][j] = apply_act_rt(v, g.act);
    }
  }

  if (!g.group_max) {
#pragma unroll
    for (int i = 0; i < TM; ++i) {
      int m = m0 + ty * TM + i;
      if (m >= g.M) continue;
      float* crow = C + (int64_t)m * g.ldc + n0 + tx * TN;
      if (n0 + tx * TN + TN <= g.N && (g.ldc & 3) == 0) {
#pragma unroll
        for (int j = 0; j < TN; j += 4)
          *reinterpret_cast<float4*>(crow + j) = make_float4(acc[i][j], acc[i][j + 1], acc[i][j + 2], acc[i][j + 3]);
      } else {
#pragma unroll
        for (int j = 0; j < TN; ++j)
          if (n0 + tx * TN + j < g.N) crow[j] = acc[i][j];
      }
    }
  } else {
    // max over the 8 rows of this thread (all inside one 32-row group), then over the 4 threads of the group
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      float v = acc[0][j];
#pragma unroll
      for (int i = 1; i < TM; ++i) v = fmaxf(v, acc[i][j]);
      red[ty][tx * TN + j] = v;
    }
    __syncthreads();
    for (int e = tid; e < 4 * BN; e += NT) {
      int grp = e / BN, col = e % BN;
      int gm = (m0 >> 5) + grp;
      if (gm < (g.M >> 5) && n0 + col < g.N) {
        float v = fmaxf(fmaxf(red[grp * 4][col], red[grp * 4 + 1][col]), fmaxf(red[grp * 4 + 2][col], red[grp * 4 + 3][col]));
        C[(int64_t)gm * g.ldc + n0 + col] = v;
      }
    }
  }
}

}  // namespace

int launch_gemm_simt(const GemmArgs& g, cudaStream_t stream) {
  if (g.M <= 0 || g.N <= 0 || g.K <= 0 || (g.K % BK) != 0 || (g.lda & 3) || (g.ldw & 3)) return -1;
  if (g.group_max && (g.M % 32) != 0) return -1;
  int batch = g.batch > 0 ? g.batch : 1;
  if (g.N >= 128 && !g.group_max) {
    dim3 grid((g.M + BM - 1) / BM, (g.N + 127) / 128, batch);
    gemm_simt_kernel<128><<<grid, NT, 0, stream>>>(g);
  } else {
    dim3 grid((g.M + BM - 1) / BM, (g.N + 63) / 64, batch);
    gemm_simt_kernel<64><<<grid, NT, 0, stream>>>(g);
  }
  return 1;
}

// Dispatcher of the generic interface (gemm.cuh).
int g_gemm_ws = 0;
int g_gemm_async = 1;

int launch_gemm(const GemmArgs& g, cudaStream_t stream) {
  if (g.precision >= 1 && gemm_tc_eligible(g)) {
    const bool async_ok = g_gemm_async && g.N <= 1024 &&
                          ((g.precision == 1 && g.a_rounded && g.w_rounded) || (g.precision == 2 && g.A_lo && g.W_lo));
    return (g_gemm_ws || async_ok || g.C_lo != nullptr) ? launch_gemm_ws(g, stream) : launch_gemm_tc(g, stream);
  }
  return launch_gemm_simt(g, stream);
}

}  // namespace lsdm
