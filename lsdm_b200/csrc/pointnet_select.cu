// PointNet++ selection kernels: farthest-point sampling, ball query, 3-nearest-neighbour weights.
// All selection arithmetic is plain fp32 in the reference's formula and operation order (no FMA
// contraction, no re-association), so the integer outputs are bit-identical to the reference's
// (model/pcd_backbone/pointnet2_utils.py:19-38,60-104,290-298); ties break to the lowest index.
#include "kernels.cuh"

namespace lsdm {

namespace {

// ---------------------------------------------------------------------------------------------
// Farthest point sampling, all four levels of one cloud in one CTA (the chain is strictly serial
// per cloud: 1024 + 256 + 64 + 16 dependent argmax rounds).  128 threads, points in shared memory,
// running min-distance in registers, block argmax = 2 x redux.sync per warp + one barrier.
// ---------------------------------------------------------------------------------------------
constexpr int FPS_T = 128;

template <int N, int NP>
__device__ __forceinline__ void fps_level(const float* sx, const float* sy, const float* sz, int start, int* s_idx,
                                          unsigned long long* s_red, int tid) {
  constexpr int PER = (N + FPS_T - 1) / FPS_T;
  constexpr bool PACKED = (N % FPS_T == 0) && (PER % 2 == 0);  // two points per instruction (fp32x2), every slot valid
  constexpr int PER2 = PACKED ? PER / 2 : 1;
  float px[PER], py[PER], pz[PER], dist[PER];
  float2 qx[PER2], qy[PER2], qz[PER2], qd[PER2];
#pragma unroll
  for (int k = 0; k < PER; ++k) {
    int p = tid + k * FPS_T;
    bool ok = p < N;
    px[k] = ok ? sx[p] : 0.f;
    py[k] = ok ? sy[p] : 0.f;
    pz[k] = ok ? sz[p] : 0.f;
    dist[k] = 1e10f;
  }
  if (PACKED) {
#pragma unroll
    for (int j = 0; j < PER2; ++j) {
      qx[j] = make_float2(px[2 * j], px[2 * j + 1]);
      qy[j] = make_float2(py[2 * j], py[2 * j + 1]);
      qz[j] = make_float2(pz[2 * j], pz[2 * j + 1]);
      qd[j] = make_float2(1e10f, 1e10f);
    }
  }
  int far = start;
  const int lane = tid & 31, warp = tid >> 5;
  for (int it = 0; it < NP; ++it) {
    if (tid == 0) s_idx[it] = far;
    float cx = sx[far], cy = sy[far], cz = sz[far];
    unsigned best_bits = 0u;
    unsigned best_idx = PACKED ? (unsigned)tid : 0xffffffffu;
    if (PACKED) {
      // x - c == x + (-c) exactly; every op is an IEEE round-to-nearest add / mul, same as the scalar path (no FMA)
      const float2 ncx = make_float2(-cx, -cx), ncy = make_float2(-cy, -cy), ncz = make_float2(-cz, -cz);
#pragma unroll
      for (int j = 0; j < PER2; ++j) {
        float2 dx = __fadd2_rn(qx[j], ncx), dy = __fadd2_rn(qy[j], ncy), dz = __fadd2_rn(qz[j], ncz);
        float2 d = __fadd2_rn(__fadd2_rn(__fmul2_rn(dx, dx), __fmul2_rn(dy, dy)), __fmul2_rn(dz, dz));
        float n0 = fminf(d.x, qd[j].x), n1 = fminf(d.y, qd[j].y);
        qd[j] = make_float2(n0, n1);
        unsigned b0 = __float_as_uint(n0), b1 = __float_as_uint(n1);  // >= 0: unsigned bit order == float order
        if (b0 > best_bits) { best_bits = b0; best_idx = (unsigned)(tid + (2 * j) * FPS_T); }
        if (b1 > best_bits) { best_bits = b1; best_idx = (unsigned)(tid + (2 * j + 1) * FPS_T); }
      }
    } else {
#pragma unroll
      for (int k = 0; k < PER; ++k) {
        int p = tid + k * FPS_T;
        if (p < N) {
          float dx = __fsub_rn(px[k], cx), dy = __fsub_rn(py[k], cy), dz = __fsub_rn(pz[k], cz);
          float d = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
          float nd = (d < dist[k]) ? d : dist[k];
          dist[k] = nd;
          unsigned b = __float_as_uint(nd);
          if (best_idx == 0xffffffffu || b > best_bits) {
            best_bits = b;
            best_idx = (unsigned)p;
          }
        }
      }
    }
    // warp argmax with lowest-index tie-break
    unsigned wmax = __reduce_max_sync(0xffffffffu, best_bits);
    unsigned cand = (best_bits == wmax && best_idx != 0xffffffffu) ? best_idx : 0xffffffffu;
    unsigned widx = __reduce_min_sync(0xffffffffu, cand);
    // key: high 32 bits distance, low 32 bits inverted index -> max key = max distance, lowest index
    unsigned long long key = ((unsigned long long)wmax << 32) | (unsigned long long)(0xffffffffu - widx);
    if (lane == 0) s_red[(it & 1) * (FPS_T / 32) + warp] = key;
    __syncthreads();
    unsigned long long bk = s_red[(it & 1) * (FPS_T / 32)];
#pragma unroll
    for (int w = 1; w < FPS_T / 32; ++w) {
      unsigned long long o = s_red[(it & 1) * (FPS_T / 32) + w];
      bk = o > bk ? o : bk;
    }
    far = (int)(0xffffffffu - (unsigned)(bk & 0xffffffffull));
  }
  __syncthreads();
}

__global__ void __launch_bounds__(FPS_T) fps4_kernel(const float* __restrict__ xyz0, const int64_t* __restrict__ start,
                                                     int n_clouds, int* __restrict__ idx1, int* __restrict__ idx2,
                                                     int* __restrict__ idx3, int* __restrict__ idx4,
                                                     float* __restrict__ xyz1, float* __restrict__ xyz2,
                                                     float* __restrict__ xyz3, float* __restrict__ xyz4) {
  __shared__ float ax[1024], ay[1024], az[1024];
  __shared__ float bx[1024], by[1024], bz[1024];
  __shared__ int s_idx[1024];
  __shared__ unsigned long long s_red[2 * (FPS_T / 32)];
  const int c = blockIdx.x, tid = threadIdx.x;
  const float* src = xyz0 + (int64_t)c * 1024 * 3;
  for (int p = tid; p < 1024; p += FPS_T) {
    ax[p] = src[p * 3 + 0];
    ay[p] = src[p * 3 + 1];
    az[p] = src[p * 3 + 2];
  }
  __syncthreads();
  // level 1: 1024 of 1024 (an FPS-ordered permutation)
  fps_level<1024, 1024>(ax, ay, az, (int)start[0 * (int64_t)n_clouds + c], s_idx, s_red, tid);
  for (int p = tid; p < 1024; p += FPS_T) {
    int j = s_idx[p];
    idx1[(int64_t)c * 1024 + p] = j;
    float x = ax[j], y = ay[j], z = az[j];
    bx[p] = x; by[p] = y; bz[p] = z;
    float* o = xyz1 + ((int64_t)c * 1024 + p) * 3;
    o[0] = x; o[1] = y; o[2] = z;
  }
  __syncthreads();
  // level 2: 256 of 1024 over l1_xyz
  fps_level<1024, 256>(bx, by, bz, (int)start[1 * (int64_t)n_clouds + c], s_idx, s_red, tid);
  for (int p = tid; p < 256; p += FPS_T) {
    int j = s_idx[p];
    idx2[(int64_t)c * 256 + p] = j;
    float x = bx[j], y = by[j], z = bz[j];
    ax[p] = x; ay[p] = y; az[p] = z;
    float* o = xyz2 + ((int64_t)c * 256 + p) * 3;
    o[0] = x; o[1] = y; o[2] = z;
  }
  __syncthreads();
  // level 3: 64 of 256
  fps_level<256, 64>(ax, ay, az, (int)start[2 * (int64_t)n_clouds + c], s_idx, s_red, tid);
  for (int p = tid; p < 64; p += FPS_T) {
    int j = s_idx[p];
    idx3[(int64_t)c * 64 + p] = j;
    float x = ax[j], y = ay[j], z = az[j];
    bx[p] = x; by[p] = y; bz[p] = z;
    float* o = xyz3 + ((int64_t)c * 64 + p) * 3;
    o[0] = x; o[1] = y; o[2] = z;
  }
  __syncthreads();
  // level 4: 16 of 64
  fps_level<64, 16>(bx, by, bz, (int)start[3 * (int64_t)n_clouds + c], s_idx, s_red, tid);
  for (int p = tid; p < 16; p += FPS_T) {
    int j = s_idx[p];
    idx4[(int64_t)c * 16 + p] = j;
    float* o = xyz4 + ((int64_t)c * 16 + p) * 3;
    o[0] = bx[j]; o[1] = by[j]; o[2] = bz[j];
  }
}

// ---------------------------------------------------------------------------------------------
// Ball query: one warp per centroid scans the N source points in index order, keeps the first 32
// with d <= r^2 (d in the expanded -2ab+a^2+b^2 form), pads with the first hit.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) ball_query_kernel(const float* __restrict__ xyz, const float* __restrict__ new_xyz,
                                                         int n_clouds, int N, int S, float r2, int* __restrict__ group) {
  extern __shared__ float4 sp[];  // (x, y, z, |p|^2) per source point: one 128-bit shared load per pair test
  const int c = blockIdx.y;
  const float* src = xyz + (int64_t)c * N * 3;
  for (int p = threadIdx.x; p < N; p += blockDim.x) {
    float x = src[p * 3], y = src[p * 3 + 1], z = src[p * 3 + 2];
    sp[p] = make_float4(x, y, z, sqnorm3(x, y, z));
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  for (int s = blockIdx.x * nwarp + warp; s < S; s += gridDim.x * nwarp) {
    const float* q = new_xyz + ((int64_t)c * S + s) * 3;
    float qx = q[0], qy = q[1], qz = q[2];
    float q2 = sqnorm3(qx, qy, qz);
    int* out = group + ((int64_t)c * S + s) * 32;
    int cnt = 0, first = N;
    // 64 source points per warp iteration: each lane tests points base+lane and base+32+lane with packed fp32x2 math
    // (same IEEE operations as sqdist_expanded); the two 32-point halves are committed in index order
    const float2 qx2 = make_float2(qx, qx), qy2 = make_float2(qy, qy), qz2 = make_float2(qz, qz), qn2 = make_float2(q2, q2);
    const float2 m2 = make_float2(-2.0f, -2.0f);
    for (int base = 0; base < N && cnt < 32; base += 64) {
      const int p0 = base + lane, p1 = base + 32 + lane;
      const float4 v0 = sp[p0 < N ? p0 : 0], v1 = sp[p1 < N ? p1 : 0];
      float2 dot = __ffma2_rn(qz2, make_float2(v0.z, v1.z), __ffma2_rn(qy2, make_float2(v0.y, v1.y), __fmul2_rn(qx2, make_float2(v0.x, v1.x))));
      float2 d = __fadd2_rn(__fadd2_rn(__fmul2_rn(m2, dot), qn2), make_float2(v0.w, v1.w));
      const bool in0 = p0 < N && !(d.x > r2), in1 = p1 < N && !(d.y > r2);
      unsigned m = __ballot_sync(0xffffffffu, in0);
      if (m) {
        if (first == N) first = base + __ffs(m) - 1;
        int pos = cnt + __popc(m & ((1u << lane) - 1u));
        if (in0 && pos < 32) out[pos] = p0;
        cnt += __popc(m);
      }
      if (cnt < 32) {
        m = __ballot_sync(0xffffffffu, in1);
        if (m) {
          if (first == N) first = base + 32 + __ffs(m) - 1;
          int pos = cnt + __popc(m & ((1u << lane) - 1u));
          if (in1 && pos < 32) out[pos] = p1;
          cnt += __popc(m);
        }
      }
    }
    if (cnt > 32) cnt = 32;
    __syncwarp();
    if (lane >= cnt) out[lane] = first;  // first == N only for an empty ball, which cannot happen (centroid is a source point)
  }
}

// ---------------------------------------------------------------------------------------------
// 3-NN inverse-distance weights for feature propagation: for every fine point the 3 smallest
// expanded-form distances to the S coarse points (ascending, lowest index on ties),
// w_k = (1/(d_k+1e-8)) / sum_k (1/(d_k+1e-8)).  Two fine points per thread share each shared-memory load.
// ---------------------------------------------------------------------------------------------
struct Top3 {
  float d0, d1, d2;
  int i0, i1, i2;
  __device__ __forceinline__ void init() {
    d0 = d1 = d2 = INFINITY;
    i0 = i1 = i2 = 0;
  }
  __device__ __forceinline__ void push(float d, int s) {
    if (d < d2) {
      if (d < d1) {
        d2 = d1; i2 = i1;
        if (d < d0) { d1 = d0; i1 = i0; d0 = d; i0 = s; }
        else { d1 = d; i1 = s; }
      } else { d2 = d; i2 = s; }
    }
  }
  __device__ __forceinline__ void store(int* nn_idx, float* nn_w, int64_t o) const {
    float r0 = 1.0f / (d0 + 1e-8f), r1 = 1.0f / (d1 + 1e-8f), r2 = 1.0f / (d2 + 1e-8f);
    float norm = (r0 + r1) + r2;
    nn_idx[o] = i0; nn_idx[o + 1] = i1; nn_idx[o + 2] = i2;
    nn_w[o] = r0 / norm; nn_w[o + 1] = r1 / norm; nn_w[o + 2] = r2 / norm;
  }
};

__global__ void __launch_bounds__(256) three_nn_kernel(const float* __restrict__ xyz1, const float* __restrict__ xyz2,
                                                       int n_clouds, int N, int S, int* __restrict__ nn_idx,
                                                       float* __restrict__ nn_w) {
  extern __shared__ float4 sp[];
  const int c = blockIdx.y;
  const float* src = xyz2 + (int64_t)c * S * 3;
  for (int p = threadIdx.x; p < S; p += blockDim.x) {
    float x = src[p * 3], y = src[p * 3 + 1], z = src[p * 3 + 2];
    sp[p] = make_float4(x, y, z, sqnorm3(x, y, z));
  }
  __syncthreads();
  const int stride = gridDim.x * blockDim.x;
  for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < N; n += 2 * stride) {
    const int n2 = n + stride;
    const bool two = n2 < N;
    const float* qa = xyz1 + ((int64_t)c * N + n) * 3;
    const float* qb = xyz1 + ((int64_t)c * N + (two ? n2 : n)) * 3;
    const float ax = qa[0], ay = qa[1], az = qa[2], a2 = sqnorm3(ax, ay, az);
    const float bx = qb[0], by = qb[1], bz = qb[2], b2 = sqnorm3(bx, by, bz);
    Top3 ta, tb;
    ta.init();
    tb.init();
    // both queries against the same coarse point in packed fp32x2: identical IEEE operations to sqdist_expanded
    // (fma(z, fma(y, x*x')), then (-2*dot + |q|^2) + |p|^2), two results per instruction
    const float2 qx2 = make_float2(ax, bx), qy2 = make_float2(ay, by), qz2 = make_float2(az, bz), qn2 = make_float2(a2, b2);
    const float2 m2 = make_float2(-2.0f, -2.0f);
    for (int s = 0; s < S; ++s) {
      const float4 v = sp[s];
      float2 dot = __ffma2_rn(qz2, make_float2(v.z, v.z), __ffma2_rn(qy2, make_float2(v.y, v.y), __fmul2_rn(qx2, make_float2(v.x, v.x))));
      float2 d = __fadd2_rn(__fadd2_rn(__fmul2_rn(m2, dot), qn2), make_float2(v.w, v.w));
      ta.push(d.x, s);
      tb.push(d.y, s);
    }
    ta.store(nn_idx, nn_w, ((int64_t)c * N + n) * 3);
    if (two) tb.store(nn_idx, nn_w, ((int64_t)c * N + n2) * 3);
  }
}

}  // namespace

int launch_fps4(const float* xyz0, const int64_t* start, int n_clouds, int* idx1, int* idx2, int* idx3, int* idx4,
                float* xyz1, float* xyz2, float* xyz3, float* xyz4, cudaStream_t st) {
  fps4_kernel<<<n_clouds, FPS_T, 0, st>>>(xyz0, start, n_clouds, idx1, idx2, idx3, idx4, xyz1, xyz2, xyz3, xyz4);
  return 1;
}

int launch_ball_query(const float* xyz, const float* new_xyz, int n_clouds, int N, int S, double radius, int* group,
                      cudaStream_t st) {
  float r2 = (float)(radius * radius);  // python double r**2 compared in fp32 (torch scalar promotion)
  int warps = 8;
  int gx = (S + warps - 1) / warps;
  if (gx > 16) gx = 16;
  dim3 grid(gx, n_clouds);
  ball_query_kernel<<<grid, warps * 32, N * sizeof(float4), st>>>(xyz, new_xyz, n_clouds, N, S, r2, group);
  return 1;
}

int launch_three_nn(const float* xyz1, const float* xyz2, int n_clouds, int N, int S, int* nn_idx, float* nn_w,
                    cudaStream_t st) {
  int threads = N >= 512 ? 256 : 64;
  dim3 grid((N + 2 * threads - 1) / (2 * threads), n_clouds);  // two fine points per thread
  three_nn_kernel<<<grid, threads, S * sizeof(float4), st>>>(xyz1, xyz2, n_clouds, N, S, nn_idx, nn_w);
  return 1;
}

}  // namespace lsdm
