// Cell-grid forms of the two O(N^2) selection scans of PointNet++ (reference pointnet2_utils.py:84-104 query_ball_point,
// :293-295 the 3-NN of PointNetFeaturePropagation), for the levels with 1024 source points.  Results are IDENTICAL to the
// full scans of pointnet_select.cu (and therefore to the reference): every candidate's distance is computed with the same
// IEEE operations in the same order,
//     d = ((-2 * (q.p)) + |q|^2) + |p|^2,   q.p = fma(qz, pz, fma(qy, py, qx * px)),
// and the grid only decides WHICH candidates are looked at, with a margin that covers the rounding of that formula:
//   * |fl(d) - d| <= 40 u (max |p|)^2 (three roundings in the dot product and in each norm, two in the sums; u = 2^-24); E below is
//     1e-5 (max |p|)^2, four times that;
//   * ball query: a point that passes `!(fl(d) > r2)` has true distance <= sqrt(r2 + E).  Cells are 1.001 x that wide (the
//     0.1 % also absorbs the rounding of the cell index itself), so every passing point lies in the 3 x 3 x 3 block of cells
//     around the centroid's cell.  Hits are collected in a per-thread 1024-bit bitmap and read out in index order: "the first 32
//     indices in ascending order, padded with the first" no matter in which order the cells were visited;
//   * 3-NN: rings of cells around the query are visited until the current third-best computed distance is strictly below
//     ((ring - 1e-4) h)^2 - E, a lower bound of the computed distance of every unvisited point; candidates are ranked by
//     (distance, index), which is what the index-order scan with strict '<' insertion produces.
// One block per cloud; the grid (<= 16^3 cells, counting sort in shared memory) is built by the block that uses it.
#include "kernels.cuh"

namespace lsdm {

namespace {

constexpr int GN = 1024;      // source points per cloud handled by these kernels
constexpr int GMAX = 16;      // cells per axis at most
constexpr int GCELLS = GMAX * GMAX * GMAX;

struct Grid {
  float minx, miny, minz, inv_h, h, E;
  int nx, ny, nz;
};

// Builds the grid over the GN points staged in shared memory (sx, sy, sz).  h_min: smallest admissible cell width (0: choose
// from the point density).  Outputs cell_start[ncell + 1] and cell_pts[GN] (point indices grouped by cell, any order inside).
__device__ void build_grid(const float* sx, const float* sy, const float* sz, const float* sn2, int n_pts, float h_req, int target_per_axis,
                           Grid& g, int* cell_start, unsigned short* cell_pts, int* s_tmp /* >= GCELLS + 64 ints */) {
  const int tid = threadIdx.x, nt = blockDim.x;
  // bounding box and largest squared norm (block reduction through shared memory)
  float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY}, nmax = 0.f;
  for (int i = tid; i < n_pts; i += nt) {
    lo[0] = fminf(lo[0], sx[i]); hi[0] = fmaxf(hi[0], sx[i]);
    lo[1] = fminf(lo[1], sy[i]); hi[1] = fmaxf(hi[1], sy[i]);
    lo[2] = fminf(lo[2], sz[i]); hi[2] = fmaxf(hi[2], sz[i]);
    nmax = fmaxf(nmax, sn2[i]);
  }
  float* red = reinterpret_cast<float*>(s_tmp);  // [7][32]
  for (int k = 0; k < 3; ++k) {
    float a = lo[k], b = hi[k];
    for (int o = 16; o > 0; o >>= 1) {
      a = fminf(a, __shfl_xor_sync(0xffffffffu, a, o));
      b = fmaxf(b, __shfl_xor_sync(0xffffffffu, b, o));
    }
    if ((tid & 31) == 0) {
      red[k * 32 + (tid >> 5)] = a;
      red[(3 + k) * 32 + (tid >> 5)] = b;
    }
  }
  nmax = warp_max(nmax);
  if ((tid & 31) == 0) red[6 * 32 + (tid >> 5)] = nmax;
  __syncthreads();
  const int nw = nt >> 5;
  for (int k = 0; k < 3; ++k) {
    float a = red[k * 32], b = red[(3 + k) * 32];
    for (int w = 1; w < nw; ++w) {
      a = fminf(a, red[k * 32 + w]);
      b = fmaxf(b, red[(3 + k) * 32 + w]);
    }
    lo[k] = a;
    hi[k] = b;
  }
  nmax = red[6 * 32];
  for (int w = 1; w < nw; ++w) nmax = fmaxf(nmax, red[6 * 32 + w]);
  __syncthreads();
  const float ext = fmaxf(fmaxf(hi[0] - lo[0], hi[1] - lo[1]), hi[2] - lo[2]);
  g.E = 1e-5f * nmax;
  float h = h_req > 0.f ? h_req : ext / (float)target_per_axis;
  if (h_req > 0.f) h = sqrtf(h_req * h_req + g.E) * 1.001f;          // ball query: covers every point that can pass the fp32 test
  h = fmaxf(h, ext / (float)GMAX * 1.0001f);                          // at most GMAX cells per axis
  if (!(h > 0.f) || !isfinite(h)) h = 1.0f;                           // degenerate cloud (all points equal): a single cell
  g.h = h;
  g.inv_h = 1.0f / h;
  g.minx = lo[0]; g.miny = lo[1]; g.minz = lo[2];
  g.nx = min(GMAX, (int)((hi[0] - lo[0]) * g.inv_h) + 1);
  g.ny = min(GMAX, (int)((hi[1] - lo[1]) * g.inv_h) + 1);
  g.nz = min(GMAX, (int)((hi[2] - lo[2]) * g.inv_h) + 1);
  const int ncell = g.nx * g.ny * g.nz;
  int* cnt = s_tmp;  // [ncell]
  for (int i = tid; i < ncell + 1; i += nt) cnt[i] = 0;
  __syncthreads();
  auto cell_of = [&](int i) {
    const int cx = min(g.nx - 1, max(0, (int)((sx[i] - g.minx) * g.inv_h)));
    const int cy = min(g.ny - 1, max(0, (int)((sy[i] - g.miny) * g.inv_h)));
    const int cz = min(g.nz - 1, max(0, (int)((sz[i] - g.minz) * g.inv_h)));
    return (cz * g.ny + cy) * g.nx + cx;
  };
  for (int i = tid; i < n_pts; i += nt) atomicAdd(&cnt[cell_of(i)], 1);
  __syncthreads();
  // exclusive scan of cnt[0..ncell) -> cell_start (one warp, 32 cells per lane step: ncell <= 4096)
  if (tid < 32) {
    int run = 0;
    for (int b0 = 0; b0 < ncell; b0 += 32) {
      const int i = b0 + tid;
      const int v = i < ncell ? cnt[i] : 0;
      int incl = v;
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (tid >= o) incl += t;
      }
      if (i < ncell) cell_start[i] = run + incl - v;
      run += __shfl_sync(0xffffffffu, incl, 31);
    }
    if (tid == 0) cell_start[ncell] = run;
  }
  __syncthreads();
  for (int i = tid; i < ncell; i += nt) cnt[i] = cell_start[i];  // cursors
  __syncthreads();
  for (int i = tid; i < n_pts; i += nt) cell_pts[atomicAdd(&cnt[cell_of(i)], 1)] = (unsigned short)i;
  __syncthreads();
}

__device__ __forceinline__ float sqdist_ref(float qx, float qy, float qz, float q2, float px, float py, float pz, float p2) {
  const float dot = __fmaf_rn(qz, pz, __fmaf_rn(qy, py, __fmul_rn(qx, px)));
  return __fadd_rn(__fmaf_rn(-2.0f, dot, q2), p2);  // fma(-2, dot, q2) == (-2 * dot) + q2 bit for bit (exact product)
}

// ---- ball query, N = 1024 source points, S centroids (1024 or 256), 32 samples ----
constexpr int BG_T = 512;
__global__ void __launch_bounds__(BG_T) ball_query_grid_kernel(const float* __restrict__ xyz, const float* __restrict__ new_xyz, int S, float radius,
                                                               float r2, int* __restrict__ group) {
  extern __shared__ unsigned char smem[];
  float* sx = reinterpret_cast<float*>(smem);
  float* sy = sx + GN;
  float* sz = sy + GN;
  float* sn2 = sz + GN;
  int* cell_start = reinterpret_cast<int*>(sn2 + GN);          // [GCELLS + 1]
  int* s_tmp = cell_start + GCELLS + 32;                        // [GCELLS + 64]
  unsigned short* cell_pts = reinterpret_cast<unsigned short*>(s_tmp + GCELLS + 64);   // [GN]
  unsigned* bits = reinterpret_cast<unsigned*>(cell_pts + GN);  // [32][BG_T]
  const int c = blockIdx.x, tid = threadIdx.x;
  const float* src = xyz + (int64_t)c * GN * 3;
  for (int i = tid; i < GN; i += BG_T) {
    const float x = src[i * 3], y = src[i * 3 + 1], z = src[i * 3 + 2];
    sx[i] = x; sy[i] = y; sz[i] = z;
    sn2[i] = sqnorm3(x, y, z);
  }
  __syncthreads();
  __shared__ Grid g;
  Grid gl;
  build_grid(sx, sy, sz, sn2, GN, radius, 0, gl, cell_start, cell_pts, s_tmp);
  if (tid == 0) g = gl;
  __syncthreads();
  gl = g;
  for (int s = tid; s < S; s += BG_T) {
    const float* q = new_xyz + ((int64_t)c * S + s) * 3;
    const float qx = q[0], qy = q[1], qz = q[2], q2 = sqnorm3(qx, qy, qz);
#pragma unroll
    for (int w = 0; w < 32; ++w) bits[w * BG_T + tid] = 0u;
    const int cx = min(gl.nx - 1, max(0, (int)((qx - gl.minx) * gl.inv_h)));
    const int cy = min(gl.ny - 1, max(0, (int)((qy - gl.miny) * gl.inv_h)));
    const int cz = min(gl.nz - 1, max(0, (int)((qz - gl.minz) * gl.inv_h)));
    for (int z = max(0, cz - 1); z <= min(gl.nz - 1, cz + 1); ++z)
      for (int y = max(0, cy - 1); y <= min(gl.ny - 1, cy + 1); ++y) {
        const int row = (z * gl.ny + y) * gl.nx;
        const int k0 = cell_start[row + max(0, cx - 1)], k1 = cell_start[row + min(gl.nx - 1, cx + 1) + 1];  // three x-adjacent cells are contiguous
        for (int k = k0; k < k1; ++k) {
          const int j = cell_pts[k];
          const float d = sqdist_ref(qx, qy, qz, q2, sx[j], sy[j], sz[j], sn2[j]);
          if (!(d > r2)) bits[(j >> 5) * BG_T + tid] |= 1u << (j & 31);
        }
      }
    int* out = group + ((int64_t)c * S + s) * 32;
    int cnt = 0, first = GN;
    for (int w = 0; w < 32 && cnt < 32; ++w) {
      unsigned b = bits[w * BG_T + tid];
      while (b && cnt < 32) {
        const int j = w * 32 + __ffs(b) - 1;
        b &= b - 1;
        if (cnt == 0) first = j;
        out[cnt++] = j;
      }
    }
    for (int k = cnt; k < 32; ++k) out[k] = first;
  }
}

// ---- 3-NN: N fine points (1024), S coarse points (1024 or 256) ----
struct Top3L {  // ranked by (distance, index)
  float d0, d1, d2;
  int i0, i1, i2;
  __device__ __forceinline__ void init() {
    d0 = d1 = d2 = INFINITY;
    i0 = i1 = i2 = 0x7fffffff;
  }
  __device__ __forceinline__ static bool lt(float d, int s, float D, int I) { return d < D || (d == D && s < I); }
  __device__ __forceinline__ void push(float d, int s) {
    if (lt(d, s, d2, i2)) {
      if (lt(d, s, d1, i1)) {
        d2 = d1; i2 = i1;
        if (lt(d, s, d0, i0)) { d1 = d0; i1 = i0; d0 = d; i0 = s; }
        else { d1 = d; i1 = s; }
      } else { d2 = d; i2 = s; }
    }
  }
};

constexpr int NG_T = 256;
__global__ void __launch_bounds__(NG_T) three_nn_grid_kernel(const float* __restrict__ xyz1, const float* __restrict__ xyz2, int N, int S,
                                                             int* __restrict__ nn_idx, float* __restrict__ nn_w) {
  extern __shared__ unsigned char smem[];
  float* sx = reinterpret_cast<float*>(smem);
  float* sy = sx + GN;
  float* sz = sy + GN;
  float* sn2 = sz + GN;
  int* cell_start = reinterpret_cast<int*>(sn2 + GN);
  int* s_tmp = cell_start + GCELLS + 32;
  unsigned short* cell_pts = reinterpret_cast<unsigned short*>(s_tmp + GCELLS + 64);
  const int c = blockIdx.x, tid = threadIdx.x;
  const float* src = xyz2 + (int64_t)c * S * 3;
  for (int i = tid; i < S; i += NG_T) {
    const float x = src[i * 3], y = src[i * 3 + 1], z = src[i * 3 + 2];
    sx[i] = x; sy[i] = y; sz[i] = z;
    sn2[i] = sqnorm3(x, y, z);
  }
  __syncthreads();
  __shared__ Grid g;
  Grid gl;
  build_grid(sx, sy, sz, sn2, S, 0.f, S >= 1024 ? 8 : 5, gl, cell_start, cell_pts, s_tmp);
  if (tid == 0) g = gl;
  __syncthreads();
  gl = g;
  const int rmax = max(gl.nx, max(gl.ny, gl.nz));
  for (int n = tid; n < N; n += NG_T) {
    const float* q = xyz1 + ((int64_t)c * N + n) * 3;
    const float qx = q[0], qy = q[1], qz = q[2], q2 = sqnorm3(qx, qy, qz);
    // the query may lie outside the coarse points' bounding box: its (unclamped) cell coordinates define the rings
    const int cx = (int)floorf((qx - gl.minx) * gl.inv_h), cy = (int)floorf((qy - gl.miny) * gl.inv_h), cz = (int)floorf((qz - gl.minz) * gl.inv_h);
    Top3L t;
    t.init();
    for (int ring = 0;; ++ring) {
      const int z0 = max(0, cz - ring), z1 = min(gl.nz - 1, cz + ring), y0 = max(0, cy - ring), y1 = min(gl.ny - 1, cy + ring);
      const int x0 = max(0, cx - ring), x1 = min(gl.nx - 1, cx + ring);
      for (int z = z0; z <= z1; ++z)
        for (int y = y0; y <= y1; ++y) {
          const bool shell_zy = (abs(z - cz) == ring) || (abs(y - cy) == ring);
          const int row = (z * gl.ny + y) * gl.nx;
          if (shell_zy) {  // the whole x range of this row belongs to the ring's shell
            if (x0 <= x1)
              for (int k = cell_start[row + x0]; k < cell_start[row + x1 + 1]; ++k) {
                const int j = cell_pts[k];
                t.push(sqdist_ref(qx, qy, qz, q2, sx[j], sy[j], sz[j], sn2[j]), j);
              }
          } else {         // only the two end cells at x = cx -+ ring
            for (int e = 0; e < 2; ++e) {
              const int x = e == 0 ? cx - ring : cx + ring;
              if (x < 0 || x >= gl.nx || (e == 1 && ring == 0)) continue;
              for (int k = cell_start[row + x]; k < cell_start[row + x + 1]; ++k) {
                const int j = cell_pts[k];
                t.push(sqdist_ref(qx, qy, qz, q2, sx[j], sy[j], sz[j], sn2[j]), j);
              }
            }
          }
        }
      // every unvisited point is more than (ring - 1e-4) cells away along some axis
      const float reach = ((float)ring - 1e-4f) * gl.h;
      const bool covered = (cx - ring <= 0 && cx + ring >= gl.nx - 1) && (cy - ring <= 0 && cy + ring >= gl.ny - 1) && (cz - ring <= 0 && cz + ring >= gl.nz - 1);
      if (covered) break;
      if (ring >= 1 && t.d2 < reach * reach - fmaxf(gl.E, 1e-5f * q2)) break;
      if (ring > 2 * rmax + 64) break;  // (unreachable: `covered` ends the search; guards against a query far outside the box)
    }
    const float r0 = 1.0f / (t.d0 + 1e-8f), r1 = 1.0f / (t.d1 + 1e-8f), r2 = 1.0f / (t.d2 + 1e-8f);
    const float norm = (r0 + r1) + r2;
    const int64_t o = ((int64_t)c * N + n) * 3;
    nn_idx[o] = t.i0; nn_idx[o + 1] = t.i1; nn_idx[o + 2] = t.i2;
    nn_w[o] = r0 / norm; nn_w[o + 1] = r1 / norm; nn_w[o + 2] = r2 / norm;
  }
}

}  // namespace

// N must be 1024.  Same output as launch_ball_query.
int launch_ball_query_grid(const float* xyz, const float* new_xyz, int n_clouds, int N, int S, double radius, int* group, cudaStream_t st) {
  if (N != GN) return -1;
  const float r2 = (float)(radius * radius);
  constexpr int smem = 4 * GN * 4 + (GCELLS + 32) * 4 + (GCELLS + 64) * 4 + GN * 2 + 32 * BG_T * 4;
  static PerDeviceOnce attr_done;
  if (smem_opt_in(attr_done, ball_query_grid_kernel, smem) != cudaSuccess) return -1;
  ball_query_grid_kernel<<<n_clouds, BG_T, smem, st>>>(xyz, new_xyz, S, (float)radius, r2, group);
  return 1;
}

// S (coarse points) must be <= 1024.  Same output as launch_three_nn.
int launch_three_nn_grid(const float* xyz1, const float* xyz2, int n_clouds, int N, int S, int* nn_idx, float* nn_w, cudaStream_t st) {
  if (S > GN || S < 3) return -1;
  constexpr int smem = 4 * GN * 4 + (GCELLS + 32) * 4 + (GCELLS + 64) * 4 + GN * 2;
  static PerDeviceOnce attr_done;
  if (smem_opt_in(attr_done, three_nn_grid_kernel, smem) != cudaSuccess) return -1;
  three_nn_grid_kernel<<<n_clouds, NG_T, smem, st>>>(xyz1, xyz2, N, S, nn_idx, nn_w);
  return 1;
}

}  // namespace lsdm
