// PointNet++ selection kernels: farthest-point sampling, ball query, 3-nearest-neighbour weights.
// All selection arithmetic is plain fp32 in the reference's formula and operation order (no FMA
// contraction, no re-association), so the integer outputs are bit-identical to the reference's
// (model/pcd_backbone/pointnet2_utils.py:19-38,60-104,290-298); ties break to the lowest index.
#include "kernels.cuh"

namespace lsdm {

namespace {

// ---------------------------------------------------------------------------------------------
// Farthest point sampling, all four levels of a cloud by ONE WARP (the chain is strictly serial per cloud: 1024 + 256 + 64 + 16
// dependent argmax rounds).  Lane l keeps points l, l + 32, ... in registers (N / 32 of them, as fp32x2 pairs) together with
// their running minimum distances; a round is: one broadcast 128-bit shared load of the new centroid (the warp's copy of the
// level's points), the distance update, a per-lane argmax, two redux.sync.  No block barrier, no key exchange through shared
// memory.  Measured on B200 (300 clouds): 0.39 ms at 0.43 x the instruction count of the round-1 form (a 128-thread block per
// cloud: 0.44 ms; 256 threads: 0.72 ms) -- that form spent most of its issue slots on per-warp overhead replicated four
// times, which the kernels of the other streams running beside it paid for.
// ---------------------------------------------------------------------------------------------
template <int N, int NP>
__device__ __forceinline__ void fps_level_warp(const float* __restrict__ src, int start, int* __restrict__ idx_out, float* __restrict__ xyz_out,
                                               int lane, bool uniform, uint32_t s_pts) {
  constexpr int PER = N / 32, PER2 = PER / 2;
  static_assert(N % 64 == 0 && PER <= 32, "pairs of points per lane");
  if (uniform) {
    // Every point of the cloud has the same coordinates (an absent object: the dataset pads it with zeros, reference
    // posa/dataset.py:456).  All squared distances are then (0 + 0) + 0 = 0 in every round, the running minimum is 0 after
    // round 0 and the argmax of an all-equal array is its first index: the selection is {start, 0, 0, ...} -- exactly what
    // the rounds below produce, without running them.
    const float x = src[0], y = src[1], z = src[2];
    for (int it = lane; it < NP; it += 32) {
      idx_out[it] = it == 0 ? start : 0;
      xyz_out[it * 3] = x; xyz_out[it * 3 + 1] = y; xyz_out[it * 3 + 2] = z;
    }
    __syncwarp();
    return;
  }
  float2 qx[PER2], qy[PER2], qz[PER2], qd[PER2];
#pragma unroll
  for (int j = 0; j < PER2; ++j) {
    const int p0 = lane + (2 * j) * 32, p1 = p0 + 32;
    qx[j] = make_float2(src[p0 * 3], src[p1 * 3]);
    qy[j] = make_float2(src[p0 * 3 + 1], src[p1 * 3 + 1]);
    qz[j] = make_float2(src[p0 * 3 + 2], src[p1 * 3 + 2]);
    qd[j] = make_float2(1e10f, 1e10f);
    // the warp's copy of the level's points, (x, y, z, -) per point: where a round's new centroid is looked up
    asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(s_pts + p0 * 16), "f"(qx[j].x), "f"(qy[j].x) : "memory");
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(s_pts + p0 * 16 + 8), "f"(qz[j].x) : "memory");
    asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(s_pts + p1 * 16), "f"(qx[j].y), "f"(qy[j].y) : "memory");
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(s_pts + p1 * 16 + 8), "f"(qz[j].y) : "memory");
  }
  __syncwarp();
  int far = start;
  for (int it = 0; it < NP; ++it) {
    const float4 cc = lds_f4(s_pts + far * 16);  // (same address in every lane: a broadcast)
    const float cx = cc.x, cy = cc.y, cz = cc.z;
    if (lane == 0) {
      idx_out[it] = far;
      xyz_out[it * 3] = cx; xyz_out[it * 3 + 1] = cy; xyz_out[it * 3 + 2] = cz;
    }
    unsigned best_bits = 0u, best_idx = (unsigned)lane;
    const float2 ncx = make_float2(-cx, -cx), ncy = make_float2(-cy, -cy), ncz = make_float2(-cz, -cz);
#pragma unroll
    for (int j = 0; j < PER2; ++j) {
      // x - c == x + (-c) exactly; every op is an IEEE round-to-nearest add / mul (no FMA).  Products packed, sums SCALAR:
      // ptxas (12.9) contracts mul.rn.f32x2 + add.rn.f32x2 into a fused FFMA2 when the product has a single use, despite the
      // explicit rounding modifiers (tools/probe/packed_fma_probe.cu: 21 % of the squared distances then differ from the
      // reference's (dx^2 + dy^2) + dz^2).  Scalar add.rn is never contracted.
      const float2 dx = __fadd2_rn(qx[j], ncx), dy = __fadd2_rn(qy[j], ncy), dz = __fadd2_rn(qz[j], ncz);
      const float2 sxx = __fmul2_rn(dx, dx), syy = __fmul2_rn(dy, dy), szz = __fmul2_rn(dz, dz);
      const float d0 = __fadd_rn(__fadd_rn(sxx.x, syy.x), szz.x), d1 = __fadd_rn(__fadd_rn(sxx.y, syy.y), szz.y);
      const float n0 = fminf(d0, qd[j].x), n1 = fminf(d1, qd[j].y);
      qd[j] = make_float2(n0, n1);
      const unsigned b0 = __float_as_uint(n0), b1 = __float_as_uint(n1);  // >= 0: unsigned bit order == float order
      if (b0 > best_bits) { best_bits = b0; best_idx = (unsigned)(lane + (2 * j) * 32); }
      if (b1 > best_bits) { best_bits = b1; best_idx = (unsigned)(lane + (2 * j + 1) * 32); }
    }
    const unsigned wmax = __reduce_max_sync(0xffffffffu, best_bits);
    far = (int)__reduce_min_sync(0xffffffffu, best_bits == wmax ? best_idx : 0xffffffffu);  // lowest index among the maxima
  }
  __syncwarp();  // the next level reads xyz_out (written by lane 0) from every lane, and rewrites the shared copy
}

// ---------------------------------------------------------------------------------------------
// Level 0 keeps ALL 1024 points (npoint == N): after r rounds at least r points have minimum distance 0 -- the selected ones and
// their exact duplicates -- and a point at distance 0 stays there, so it can only be "selected" again when every point is at 0, in
// which case the argmax of the all-zero array is index 0 whoever is still tracked.  The warp therefore drops the dead points every
// 128 rounds: the live (distance != 0) points are re-dealt over the lanes IN INDEX ORDER (ballot prefix sums through an 8 KB
// shared staging area), so that position p = lane + 32 * slot still enumerates ascending indices and the tie rule (strict '>'
// inside a lane, lowest index across lanes) is unchanged; phase k tracks 32 - 4k points per lane instead of 32.  Same IEEE
// operations on every tracked point, hence the same selection order; 0.56 x the distance updates of level 0.
// ---------------------------------------------------------------------------------------------
template <int PER2>
struct FpsPts {
  float2 x[PER2], y[PER2], z[PER2], d[PER2];
  unsigned i0[PER2], i1[PER2];  // original indices of the pair's two points
};

template <int PER2>
__device__ __forceinline__ int fps_rounds_tracked(FpsPts<PER2>& P, int far, int it0, int n_rounds, int* __restrict__ idx_out,
                                                  float* __restrict__ xyz_out, int lane, uint32_t s_pts) {
  for (int it = it0; it < it0 + n_rounds; ++it) {
    const float4 cc = lds_f4(s_pts + far * 16);
    const float cx = cc.x, cy = cc.y, cz = cc.z;
    if (lane == 0) {
      idx_out[it] = far;
      xyz_out[it * 3] = cx; xyz_out[it * 3 + 1] = cy; xyz_out[it * 3 + 2] = cz;
    }
    unsigned best_bits = 0u, best_idx = (unsigned)lane;
    const float2 ncx = make_float2(-cx, -cx), ncy = make_float2(-cy, -cy), ncz = make_float2(-cz, -cz);
#pragma unroll
    for (int j = 0; j < PER2; ++j) {
      // (products packed, sums scalar: see fps_level_warp)
      const float2 dx = __fadd2_rn(P.x[j], ncx), dy = __fadd2_rn(P.y[j], ncy), dz = __fadd2_rn(P.z[j], ncz);
      const float2 sxx = __fmul2_rn(dx, dx), syy = __fmul2_rn(dy, dy), szz = __fmul2_rn(dz, dz);
      const float d0 = __fadd_rn(__fadd_rn(sxx.x, syy.x), szz.x), d1 = __fadd_rn(__fadd_rn(sxx.y, syy.y), szz.y);
      const float n0 = fminf(d0, P.d[j].x), n1 = fminf(d1, P.d[j].y);
      P.d[j] = make_float2(n0, n1);
      const unsigned b0 = __float_as_uint(n0), b1 = __float_as_uint(n1);
      if (b0 > best_bits) { best_bits = b0; best_idx = P.i0[j]; }
      if (b1 > best_bits) { best_bits = b1; best_idx = P.i1[j]; }
    }
    const unsigned wmax = __reduce_max_sync(0xffffffffu, best_bits);
    far = (int)__reduce_min_sync(0xffffffffu, best_bits == wmax ? best_idx : 0xffffffffu);
  }
  return far;
}

// live points of P (distance != 0), in position order, -> Q (NEW2 pairs per lane; unused slots: a dead dummy at the origin)
template <int OLD2, int NEW2>
__device__ __forceinline__ void fps_compact(const FpsPts<OLD2>& P, FpsPts<NEW2>& Q, int lane, uint32_t s_pts, uint32_t s_stage) {
  const unsigned lt = (1u << lane) - 1u;
  int base = 0;
#pragma unroll
  for (int j = 0; j < OLD2; ++j) {
#pragma unroll
    for (int hf = 0; hf < 2; ++hf) {
      const float d = hf ? P.d[j].y : P.d[j].x;
      const unsigned id = hf ? P.i1[j] : P.i0[j];
      const bool live = !(d == 0.0f);
      const unsigned m = __ballot_sync(0xffffffffu, live);
      if (live) {
        const uint32_t a = s_stage + (uint32_t)(base + __popc(m & lt)) * 8u;
        asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(a), "r"(__float_as_uint(d)), "r"(id) : "memory");
      }
      base += __popc(m);
    }
  }
  __syncwarp();
#pragma unroll
  for (int j = 0; j < NEW2; ++j) {
    float dd[2], xx[2], yy[2], zz[2];
    unsigned ii[2];
#pragma unroll
    for (int hf = 0; hf < 2; ++hf) {
      const int pos = lane + (2 * j + hf) * 32;
      dd[hf] = 0.0f; xx[hf] = 0.0f; yy[hf] = 0.0f; zz[hf] = 0.0f; ii[hf] = 0u;
      if (pos < base) {
        unsigned db, id;
        asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(db), "=r"(id) : "r"(s_stage + (uint32_t)pos * 8u) : "memory");
        const float4 c = lds_f4(s_pts + id * 16);
        dd[hf] = __uint_as_float(db); xx[hf] = c.x; yy[hf] = c.y; zz[hf] = c.z; ii[hf] = id;
      }
    }
    Q.x[j] = make_float2(xx[0], xx[1]); Q.y[j] = make_float2(yy[0], yy[1]); Q.z[j] = make_float2(zz[0], zz[1]);
    Q.d[j] = make_float2(dd[0], dd[1]);
    Q.i0[j] = ii[0]; Q.i1[j] = ii[1];
  }
  __syncwarp();  // the staging area is rewritten by the next compaction
}

template <int K>  // phase K of level 0: rounds [128 K, 128 K + 128) on 32 - 4 K points per lane, then hand the survivors on
__device__ __forceinline__ void fps_level0_phase(FpsPts<16 - 2 * K>& P, int far, int* __restrict__ idx_out, float* __restrict__ xyz_out, int lane,
                                                 uint32_t s_pts, uint32_t s_stage) {
  far = fps_rounds_tracked<16 - 2 * K>(P, far, 128 * K, 128, idx_out, xyz_out, lane, s_pts);
  if constexpr (K < 7) {
    FpsPts<14 - 2 * K> Q;
    fps_compact<16 - 2 * K, 14 - 2 * K>(P, Q, lane, s_pts, s_stage);
    fps_level0_phase<K + 1>(Q, far, idx_out, xyz_out, lane, s_pts, s_stage);
  }
}

__device__ __forceinline__ void fps_level0_compacting(const float* __restrict__ src, int start, int* __restrict__ idx_out, float* __restrict__ xyz_out,
                                                      int lane, uint32_t s_pts, uint32_t s_stage) {
  FpsPts<16> P;
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const int p0 = lane + (2 * j) * 32, p1 = p0 + 32;
    P.x[j] = make_float2(src[p0 * 3], src[p1 * 3]);
    P.y[j] = make_float2(src[p0 * 3 + 1], src[p1 * 3 + 1]);
    P.z[j] = make_float2(src[p0 * 3 + 2], src[p1 * 3 + 2]);
    P.d[j] = make_float2(1e10f, 1e10f);
    P.i0[j] = (unsigned)p0; P.i1[j] = (unsigned)p1;
    asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(s_pts + p0 * 16), "f"(P.x[j].x), "f"(P.y[j].x) : "memory");
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(s_pts + p0 * 16 + 8), "f"(P.z[j].x) : "memory");
    asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(s_pts + p1 * 16), "f"(P.x[j].y), "f"(P.y[j].y) : "memory");
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(s_pts + p1 * 16 + 8), "f"(P.z[j].y) : "memory");
  }
  __syncwarp();
  fps_level0_phase<0>(P, start, idx_out, xyz_out, lane, s_pts, s_stage);
  __syncwarp();
}

__global__ void __launch_bounds__(32) fps4_kernel(const float* __restrict__ xyz0, const int64_t* __restrict__ start, int n_clouds,
                                                       int* __restrict__ idx1, int* __restrict__ idx2, int* __restrict__ idx3,
                                                       int* __restrict__ idx4, float* __restrict__ xyz1, float* __restrict__ xyz2,
                                                       float* __restrict__ xyz3, float* __restrict__ xyz4, int uniform_shortcut,
                                                       int compact_level0) {
  __shared__ float4 s_copy[1024];
  __shared__ uint2 s_live[1024];  // (distance bits, index) of the live points while level 0 re-deals them
  const uint32_t s_pts = smem_addr(s_copy);
  const int c = blockIdx.x, lane = threadIdx.x;
  const float* src = xyz0 + (int64_t)c * 1024 * 3;
  // all 1024 points equal (== compares values: +0 and -0 are the same point, a NaN never is)?  Then every level is uniform too.
  const float x0 = src[0], y0 = src[1], z0 = src[2];
  bool same = fabsf(x0) < INFINITY && fabsf(y0) < INFINITY && fabsf(z0) < INFINITY;  // inf - inf is NaN, not 0: full rounds
  for (int p = lane; p < 1024; p += 32) same = same && src[p * 3] == x0 && src[p * 3 + 1] == y0 && src[p * 3 + 2] == z0;
  const bool uniform = uniform_shortcut != 0 && __all_sync(0xffffffffu, same);
  float* l1 = xyz1 + (int64_t)c * 1024 * 3;
  float* l2 = xyz2 + (int64_t)c * 256 * 3;
  float* l3 = xyz3 + (int64_t)c * 64 * 3;
  float* l4 = xyz4 + (int64_t)c * 16 * 3;
  // start draws are randint(0, N) in the reference (an out-of-range index raises there); clamped here so that a bad caller buffer
  // cannot address outside the cloud
  auto start_of = [&](int level, int n) {
    const int64_t v = start[level * (int64_t)n_clouds + c];
    return (int)(v < 0 ? 0 : (v >= n ? n - 1 : v));
  };
  if (uniform || compact_level0 == 0)
    fps_level_warp<1024, 1024>(src, start_of(0, 1024), idx1 + (int64_t)c * 1024, l1, lane, uniform, s_pts);
  else
    fps_level0_compacting(src, start_of(0, 1024), idx1 + (int64_t)c * 1024, l1, lane, s_pts, smem_addr(s_live));
  fps_level_warp<1024, 256>(l1, start_of(1, 1024), idx2 + (int64_t)c * 256, l2, lane, uniform, s_pts);
  fps_level_warp<256, 64>(l2, start_of(2, 256), idx3 + (int64_t)c * 64, l3, lane, uniform, s_pts);
  fps_level_warp<64, 16>(l3, start_of(3, 64), idx4 + (int64_t)c * 16, l4, lane, uniform, s_pts);
}

// ---------------------------------------------------------------------------------------------
// Shared layout for the two scans below: source points in PAIRS, A[i] = (x0, x1, y0, y1), B[i] = (z0, z1, |p0|^2, |p1|^2),
// so that one broadcast 128-bit shared load yields ready-made fp32x2 operands.  The expanded distance
//   d = ((-2 * (q.p)) + |q|^2) + |p|^2,   q.p = fma(qz, pz, fma(qy, py, qx * px))
// is evaluated with packed fp32x2 instructions; fma(-2, dot, |q|^2) == (-2 * dot) + |q|^2 bit for bit because the product
// by a power of two is exact.  Same IEEE operations and order as sqdist_expanded (common.cuh).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float2 sqdist2(float2 qx, float2 qy, float2 qz, float2 qn, float4 a, float4 b) {
  const float2 dot = __ffma2_rn(qz, make_float2(b.x, b.y), __ffma2_rn(qy, make_float2(a.z, a.w), __fmul2_rn(qx, make_float2(a.x, a.y))));
  return __fadd2_rn(__ffma2_rn(make_float2(-2.0f, -2.0f), dot, qn), make_float2(b.z, b.w));
}

// ---------------------------------------------------------------------------------------------
// Ball query: one THREAD per centroid scans the N source points in index order (every thread of the block reads the same
// point pair: broadcast loads, no cross-lane traffic), keeps the first 32 with d <= r^2, pads with the first hit.
// The per-thread result lists live in a transposed shared tile and leave as coalesced 128-byte rows.
// ---------------------------------------------------------------------------------------------
constexpr int BQ_T = 128;  // centroids per block
__global__ void __launch_bounds__(BQ_T) ball_query_kernel(const float* __restrict__ xyz, const float* __restrict__ new_xyz,
                                                          int n_clouds, int N, int S, float r2, int* __restrict__ group,
                                                          const int* __restrict__ compose) {
  extern __shared__ float4 sp[];                                  // [N/2] A | [N/2] B
  __shared__ int s_out[32][BQ_T + 1];
  const int c = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int NP = N >> 1;
  float4* sA = sp;
  float4* sB = sp + NP;
  const float* src = xyz + (int64_t)c * N * 3;
  for (int i = tid; i < NP; i += BQ_T) {
    const float x0 = src[i * 6], y0 = src[i * 6 + 1], z0 = src[i * 6 + 2];
    const float x1 = src[i * 6 + 3], y1 = src[i * 6 + 4], z1 = src[i * 6 + 5];
    sA[i] = make_float4(x0, x1, y0, y1);
    sB[i] = make_float4(z0, z1, sqnorm3(x0, y0, z0), sqnorm3(x1, y1, z1));
  }
  __syncthreads();
  const int s = blockIdx.x * BQ_T + tid;
  const bool valid = s < S;
  const float* q = new_xyz + ((int64_t)c * S + (valid ? s : 0)) * 3;
  const float qx = q[0], qy = q[1], qz = q[2];
  const float q2 = sqnorm3(qx, qy, qz);
  const float2 qx2 = make_float2(qx, qx), qy2 = make_float2(qy, qy), qz2 = make_float2(qz, qz), qn2 = make_float2(q2, q2);
  int cnt = valid ? 0 : 32;
  for (int i0 = 0; i0 < NP; i0 += 8) {
    if (__all_sync(0xffffffffu, cnt >= 32)) break;  // every centroid of this warp has its 32 samples
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int i = i0 + u;
      const float2 d = sqdist2(qx2, qy2, qz2, qn2, sA[i], sB[i]);
      if (!(d.x > r2) && cnt < 32) s_out[cnt++][tid] = 2 * i;
      if (!(d.y > r2) && cnt < 32) s_out[cnt++][tid] = 2 * i + 1;
    }
  }
  if (valid) {
    const int first = cnt > 0 ? s_out[0][tid] : N;  // an empty ball cannot happen (the centroid is a source point)
    for (int k = cnt; k < 32; ++k) s_out[k][tid] = first;
  }
  __syncwarp();
  // warp w owns centroids [32w, 32w+32) of the block: lane = sample slot, one 128-byte row per centroid
  for (int j = 0; j < 32; ++j) {
    const int sj = blockIdx.x * BQ_T + warp * 32 + j;
    if (sj < S) {
      // compose (optional, [n_clouds, N]): the stored index is compose[c][i] instead of i -- the source points' rows live in
      // another order (lsdm_sample_loop keeps level 1 in cloud order: i is a level-0 FPS position, compose = the FPS indices)
      const int i = s_out[lane][warp * 32 + j];
      group[((int64_t)c * S + sj) * 32 + lane] = compose ? compose[(int64_t)c * N + i] : i;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// 3-NN inverse-distance weights for feature propagation: for every fine point the 3 smallest
// expanded-form distances to the S coarse points (ascending, lowest index on ties),
// w_k = (1/(d_k+1e-8)) / sum_k (1/(d_k+1e-8)).  Two fine points (one fp32x2 pair) per thread share each broadcast load of a
// coarse point.
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
                                                       float* __restrict__ nn_w, int uniform_shortcut) {
  extern __shared__ float4 sp[];  // per coarse point: (x, x, y, y) | (z, z, |p|^2, |p|^2)
  const int c = blockIdx.y;
  const float* src = xyz2 + (int64_t)c * S * 3;
  bool same = true;
  const float x0 = src[0], y0 = src[1], z0 = src[2];
  for (int p = threadIdx.x; p < S; p += blockDim.x) {
    const float x = src[p * 3], y = src[p * 3 + 1], z = src[p * 3 + 2], w = sqnorm3(x, y, z);
    sp[2 * p] = make_float4(x, x, y, y);
    sp[2 * p + 1] = make_float4(z, z, w, w);
    same = same && x == x0 && y == y0 && z == z0;
  }
  // All S coarse points equal (absent object, zero-padded): every candidate has the same distance, so the scan below keeps
  // candidates 0, 1, 2 (insertion is strict '<', ties keep the lowest index) -- scanning only those three is the same result.
  const int all_same = __syncthreads_and(same ? 1 : 0);  // also the barrier that publishes the staged coarse points
  const int S_scan = (uniform_shortcut != 0 && all_same != 0 && S >= 3) ? 3 : S;
  // two fine points per thread (one fp32x2 pair).  Each has its OWN insertion branch: an insertion happens ~3/s of the
  // time at coarse point s, so a warp-level branch per query slot is skipped far more often than a shared one would be.
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
    const float2 qx2 = make_float2(ax, bx), qy2 = make_float2(ay, by), qz2 = make_float2(az, bz), qn2 = make_float2(a2, b2);
    for (int s = 0; s < S_scan; ++s) {
      const float2 d = sqdist2(qx2, qy2, qz2, qn2, sp[2 * s], sp[2 * s + 1]);
      ta.push(d.x, s);
      tb.push(d.y, s);
    }
    ta.store(nn_idx, nn_w, ((int64_t)c * N + n) * 3);
    if (two) tb.store(nn_idx, nn_w, ((int64_t)c * N + n2) * 3);
  }
}

}  // namespace

// 1 (default): clouds whose points all coincide (absent objects) take the closed-form FPS order / 3-candidate 3-NN scan -- the
// same integers as the full scans (tests/test_gpu_parity.py::test_uniform_cloud_shortcuts_are_exact); 0: always the full scans.
int g_select_uniform_shortcut = 1;
// 1 (default): FPS level 0 drops the points at distance 0 every 128 rounds (same selection order); 0: all 1024 points every round.
int g_fps_compact = 1;

int launch_fps4(const float* xyz0, const int64_t* start, int n_clouds, int* idx1, int* idx2, int* idx3, int* idx4,
                float* xyz1, float* xyz2, float* xyz3, float* xyz4, cudaStream_t st) {
  fps4_kernel<<<n_clouds, 32, 0, st>>>(xyz0, start, n_clouds, idx1, idx2, idx3, idx4, xyz1, xyz2, xyz3, xyz4, g_select_uniform_shortcut, g_fps_compact);
  return 1;
}

int launch_ball_query(const float* xyz, const float* new_xyz, int n_clouds, int N, int S, double radius, int* group,
                      cudaStream_t st, const int* compose) {
  float r2 = (float)(radius * radius);  // python double r**2 compared in fp32 (torch scalar promotion)
  if (N & 15) return -1;  // the scan is unrolled over 8 point pairs
  dim3 grid((S + BQ_T - 1) / BQ_T, n_clouds);
  ball_query_kernel<<<grid, BQ_T, N * sizeof(float4) /* N/2 pairs x 2 float4 */, st>>>(xyz, new_xyz, n_clouds, N, S, r2, group, compose);
  return 1;
}

int launch_three_nn(const float* xyz1, const float* xyz2, int n_clouds, int N, int S, int* nn_idx, float* nn_w,
                    cudaStream_t st) {
  int threads = N >= 512 ? 256 : 64;
  dim3 grid((N + 2 * threads - 1) / (2 * threads), n_clouds);  // two fine points per thread
  three_nn_kernel<<<grid, threads, 2 * S * sizeof(float4), st>>>(xyz1, xyz2, n_clouds, N, S, nn_idx, nn_w, g_select_uniform_shortcut);
  return 1;
}

}  // namespace lsdm
