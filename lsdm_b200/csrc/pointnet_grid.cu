// Cell-grid forms of the two O(N^2) selection scans of PointNet++ (reference pointnet2_utils.py:84-104 query_ball_point,
// :293-295 the 3-NN of PointNetFeaturePropagation), for the levels with up to 1024 source points.  Results are IDENTICAL to
// the full scans of pointnet_select.cu (and therefore to the reference): every candidate's distance is computed with the
// same IEEE operations in the same order,
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
// One 256-thread block per cloud builds the grid in shared memory (counting sort of the points by cell: one float4
// (x, y, z, |p|^2) + the original index per sorted slot, so a candidate costs ONE 128-bit shared load) and then answers the
// cloud's queries, one query per thread at a time.  The 9 (z, y) rows of a 3 x 3 x 3 block are contiguous runs of the sorted
// array (three x-adjacent cells); a thread walks them in ONE flattened loop (row advance and candidate evaluation are
// iterations of the same loop) so that a warp's trip count is the maximum of its lanes' candidate counts, not the sum over
// rows of per-row maxima.  Queries whose block holds a large share of the cloud (dense clusters, degenerate clouds) take the
// index-order scan with early exit instead -- the worst case costs what the full scan costs.
#include "kernels.cuh"

namespace lsdm {

namespace {

constexpr int GN = 1024;      // source points per cloud at most
constexpr int GMAX = 16;      // cells per axis at most
constexpr int GCELLS = GMAX * GMAX * GMAX;
constexpr int GT = 256;       // threads per block
constexpr int GPER = GN / GT;

struct Grid {
  float minx, miny, minz, inv_h, h, E;
  int nx, ny, nz;
  int uniform;  // every source point has the same coordinates
};

struct GridSmem {
  float4* pts;            // [GN]  sorted by cell: x, y, z, |p|^2
  unsigned short* oidx;   // [GN]  original index of the sorted slot
  unsigned short* cstart; // [GCELLS + 8]
  int* scratch;           // [GCELLS] cell counters during the build (may alias memory the queries use afterwards)
  float* red;             // [7 * 8] block reduction
};

__device__ __forceinline__ float sqdist_ref(float qx, float qy, float qz, float q2, float px, float py, float pz, float p2) {
  const float dot = __fmaf_rn(qz, pz, __fmaf_rn(qy, py, __fmul_rn(qx, px)));
  return __fadd_rn(__fmaf_rn(-2.0f, dot, q2), p2);  // fma(-2, dot, q2) == (-2 * dot) + q2 bit for bit (exact product)
}

// Exclusive scan of the GCELLS cell counters into 16-bit start offsets (start[GCELLS] = total): 16 consecutive cells per
// thread, warp scan of the thread sums, warp totals through shared memory (wtot: GT / 32 ints).  Contains one barrier; the
// caller synchronises before reading `start`.
__device__ __forceinline__ void cell_scan(const int* counters, unsigned short* start, int* wtot) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int c16[16];
  const int4* cp = reinterpret_cast<const int4*>(counters) + tid * 4;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int4 v = cp[q];
    c16[q * 4] = v.x; c16[q * 4 + 1] = v.y; c16[q * 4 + 2] = v.z; c16[q * 4 + 3] = v.w;
  }
  int sum = 0;
#pragma unroll
  for (int j = 0; j < 16; ++j) sum += c16[j];
  int incl = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) wtot[warp] = incl;
  __syncthreads();
  int base = incl - sum;
#pragma unroll
  for (int w = 0; w < GT / 32; ++w) base += (w < warp) ? wtot[w] : 0;
  unsigned short st16[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    st16[j] = (unsigned short)base;
    base += c16[j];
  }
  uint4* op = reinterpret_cast<uint4*>(start + tid * 16);
  op[0] = make_uint4(st16[0] | (st16[1] << 16), st16[2] | (st16[3] << 16), st16[4] | (st16[5] << 16), st16[6] | (st16[7] << 16));
  op[1] = make_uint4(st16[8] | (st16[9] << 16), st16[10] | (st16[11] << 16), st16[12] | (st16[13] << 16), st16[14] | (st16[15] << 16));
  if (tid == GT - 1) start[GCELLS] = (unsigned short)base;
}

// Builds the grid over the n_pts (<= GN) points of `src`.  h_req > 0: cells at least sqrt(h_req^2 + E) * 1.001 wide (ball
// query of radius h_req); h_req == 0: about target_per_axis cells along the longest axis.
__device__ void build_grid(const float* __restrict__ src, int n_pts, float h_req, int target_per_axis, Grid& g, const GridSmem& s) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float px[GPER], py[GPER], pz[GPER];
  float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY}, nmax = 0.f;
#pragma unroll
  for (int k = 0; k < GPER; ++k) {
    const int i = tid + k * GT;
    if (i < n_pts) {
      px[k] = src[i * 3]; py[k] = src[i * 3 + 1]; pz[k] = src[i * 3 + 2];
      lo[0] = fminf(lo[0], px[k]); hi[0] = fmaxf(hi[0], px[k]);
      lo[1] = fminf(lo[1], py[k]); hi[1] = fmaxf(hi[1], py[k]);
      lo[2] = fminf(lo[2], pz[k]); hi[2] = fmaxf(hi[2], pz[k]);
      nmax = fmaxf(nmax, sqnorm3(px[k], py[k], pz[k]));
    } else {
      px[k] = py[k] = pz[k] = 0.f;
    }
  }
#pragma unroll
  for (int k = 0; k < 3; ++k) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      lo[k] = fminf(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], o));
      hi[k] = fmaxf(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], o));
    }
  }
  nmax = warp_max(nmax);
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      s.red[k * 8 + warp] = lo[k];
      s.red[(3 + k) * 8 + warp] = hi[k];
    }
    s.red[6 * 8 + warp] = nmax;
  }
  // zero the cell counters meanwhile
  for (int i = tid; i < GCELLS / 4; i += GT) reinterpret_cast<int4*>(s.scratch)[i] = make_int4(0, 0, 0, 0);
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    float a = s.red[k * 8], b = s.red[(3 + k) * 8];
#pragma unroll
    for (int w = 1; w < GT / 32; ++w) {
      a = fminf(a, s.red[k * 8 + w]);
      b = fmaxf(b, s.red[(3 + k) * 8 + w]);
    }
    lo[k] = a;
    hi[k] = b;
  }
  nmax = s.red[6 * 8];
#pragma unroll
  for (int w = 1; w < GT / 32; ++w) nmax = fmaxf(nmax, s.red[6 * 8 + w]);
  const float ext = fmaxf(fmaxf(hi[0] - lo[0], hi[1] - lo[1]), hi[2] - lo[2]);
  g.uniform = (hi[0] == lo[0] && hi[1] == lo[1] && hi[2] == lo[2]) ? 1 : 0;
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
  // counting sort by cell: the atomic's return value is the point's rank inside its cell
  int cell[GPER], rank[GPER];
#pragma unroll
  for (int k = 0; k < GPER; ++k) {
    const int i = tid + k * GT;
    const int cx = min(g.nx - 1, max(0, (int)((px[k] - g.minx) * g.inv_h)));
    const int cy = min(g.ny - 1, max(0, (int)((py[k] - g.miny) * g.inv_h)));
    const int cz = min(g.nz - 1, max(0, (int)((pz[k] - g.minz) * g.inv_h)));
    cell[k] = (cz * g.ny + cy) * g.nx + cx;
    rank[k] = i < n_pts ? atomicAdd(&s.scratch[cell[k]], 1) : 0;
  }
  __syncthreads();
  cell_scan(s.scratch, s.cstart, reinterpret_cast<int*>(s.red));  // (the float reductions above were consumed before the last barrier)
  __syncthreads();
#pragma unroll
  for (int k = 0; k < GPER; ++k) {
    const int i = tid + k * GT;
    if (i < n_pts) {
      const int pos = s.cstart[cell[k]] + rank[k];
      s.pts[pos] = make_float4(px[k], py[k], pz[k], sqnorm3(px[k], py[k], pz[k]));
      s.oidx[pos] = (unsigned short)i;
    }
  }
  __syncthreads();
}

// Orders the cloud's queries by grid cell (counting sort, clamped cell of the query): consecutive threads then work on
// neighbouring queries, whose candidate runs have similar lengths and the same shared-memory addresses (the walk below is
// one divergent loop per thread).  order[i] = index of the i-th query; `counters` / `qstart` are scratch.
__device__ void sort_queries(const float* __restrict__ qxyz, int n_q, const Grid& g, int* counters, unsigned short* qstart, int* wtot,
                             unsigned short* order) {
  const int tid = threadIdx.x;
  for (int i = tid; i < GCELLS / 4; i += GT) reinterpret_cast<int4*>(counters)[i] = make_int4(0, 0, 0, 0);
  __syncthreads();
  int cell[GPER], rank[GPER];
#pragma unroll
  for (int k = 0; k < GPER; ++k) {
    const int i = tid + k * GT;
    cell[k] = 0;
    rank[k] = 0;
    if (i < n_q) {
      const float x = qxyz[i * 3], y = qxyz[i * 3 + 1], z = qxyz[i * 3 + 2];
      // (int) of a NaN / out-of-range float is clamped by the min / max below: any value is a valid sort key
      const int cx = min(g.nx - 1, max(0, (int)((x - g.minx) * g.inv_h)));
      const int cy = min(g.ny - 1, max(0, (int)((y - g.miny) * g.inv_h)));
      const int cz = min(g.nz - 1, max(0, (int)((z - g.minz) * g.inv_h)));
      cell[k] = (cz * g.ny + cy) * g.nx + cx;
      rank[k] = atomicAdd(&counters[cell[k]], 1);
    }
  }
  __syncthreads();
  cell_scan(counters, qstart, wtot);
  __syncthreads();
#pragma unroll
  for (int k = 0; k < GPER; ++k) {
    const int i = tid + k * GT;
    if (i < n_q) order[qstart[cell[k]] + rank[k]] = (unsigned short)i;
  }
  __syncthreads();
}

// ---- ball query: n_src (<= 1024) source points, S centroids, 32 samples ----
// Output staging: four consecutive slots of a group leave as one 128-bit store.
struct Emit {
  int v0, v1, v2, v3, cnt, first;
  int* out;
  __device__ __forceinline__ void init(int* o, int none) { cnt = 0; first = none; out = o; v0 = v1 = v2 = v3 = 0; }
  __device__ __forceinline__ void push(int j) {
    if (cnt == 0) first = j;
    v0 = v1; v1 = v2; v2 = v3; v3 = j;
    ++cnt;
    if ((cnt & 3) == 0) *reinterpret_cast<int4*>(out + cnt - 4) = make_int4(v0, v1, v2, v3);
  }
  __device__ __forceinline__ void pad() {  // remaining slots repeat the first hit
    while (cnt & 3) push(first);
    const int4 f = make_int4(first, first, first, first);
    for (; cnt < 32; cnt += 4) *reinterpret_cast<int4*>(out + cnt) = f;
  }
};

constexpr int BALL_INORDER_MIN = 480;  // candidates in the 3 x 3 x 3 block from which the index-order scan is expected to be shorter

// Optional plan of sa1's distinct rows (Sa1PlanOut, S == 1024 only): what sa1_plan_kernel (sa_fused.cu) derives from the
// finished groups -- per centroid the number of distinct neighbours, the centroids of each block of 128 packed greedily into
// 64-row half tiles, the (neighbour | centroid << 10 | first-row flag << 20) row words -- written by the block that has just
// produced the groups, so the 39 MB of group lists are not read back by a second kernel.  Same layout, same tiles.
struct Sa1PlanOut {
  int* rows;       // [C, 256 * 128]
  int* tile_used;  // [C, 256]
  int* tiles;      // [C]
};
constexpr int PLAN_MAXT = 256;

__global__ void __launch_bounds__(GT) ball_query_grid_kernel(const float* __restrict__ xyz, const float* __restrict__ new_xyz, int n_src, int S,
                                                             float radius, float r2, int* __restrict__ group, Sa1PlanOut plan) {
  extern __shared__ __align__(16) unsigned char smem[];
  GridSmem s;
  s.pts = reinterpret_cast<float4*>(smem);
  s.oidx = reinterpret_cast<unsigned short*>(s.pts + GN);
  s.cstart = s.oidx + GN;
  unsigned* bits = reinterpret_cast<unsigned*>(s.cstart + GCELLS + 8);  // [32][GT]: bit j of thread t's hit set in word (j >> 5) * GT + t
  s.scratch = reinterpret_cast<int*>(bits);                             // 16 KB of the 32 KB bitmap
  s.red = reinterpret_cast<float*>(bits + 32 * GT);
  unsigned short* order = reinterpret_cast<unsigned short*>(s.red + 64);  // [GN] queries in cell order
  unsigned short* qcnt = order + GN;                                       // [GN] distinct neighbours per centroid (plan)
  const int c = blockIdx.x, tid = threadIdx.x;
  const float* src = xyz + (int64_t)c * n_src * 3;
  Grid g;
  build_grid(src, n_src, radius, 0, g, s);
  sort_queries(new_xyz + (int64_t)c * S * 3, S, g, s.scratch, reinterpret_cast<unsigned short*>(bits + GCELLS), reinterpret_cast<int*>(s.red), order);
  for (int i = tid; i < 32 * GT; i += GT) bits[i] = 0u;  // (the counters and the query offsets lived here)
  __syncthreads();
  const uint32_t a_pts = smem_addr(s.pts), a_oidx = smem_addr(s.oidx), a_cs = smem_addr(s.cstart), a_bits = smem_addr(bits) + tid * 4;
  for (int qi = tid; qi < S; qi += GT) {
    const int q = order[qi];
    const float* qp = new_xyz + ((int64_t)c * S + q) * 3;
    const float qx = qp[0], qy = qp[1], qz = qp[2], q2 = sqnorm3(qx, qy, qz);
    Emit em;
    em.init(group + ((int64_t)c * S + q) * 32, n_src);
    const int cx = min(g.nx - 1, max(0, (int)((qx - g.minx) * g.inv_h)));
    const int cy = min(g.ny - 1, max(0, (int)((qy - g.miny) * g.inv_h)));
    const int cz = min(g.nz - 1, max(0, (int)((qz - g.minz) * g.inv_h)));
    const int x0 = max(0, cx - 1), x1 = min(g.nx - 1, cx + 1) + 1;
    const int y0 = max(0, cy - 1), y1 = min(g.ny - 1, cy + 1);
    const int z0 = max(0, cz - 1), z1 = min(g.nz - 1, cz + 1);
    int total = 0;
    for (int z = z0; z <= z1; ++z)
      for (int y = y0; y <= y1; ++y) {
        const int row = (z * g.ny + y) * g.nx;
        total += (int)lds_u16(a_cs + (row + x1) * 2) - (int)lds_u16(a_cs + (row + x0) * 2);
      }
    if (total >= BALL_INORDER_MIN) {
      // a large share of the cloud is within reach: scanning in index order stops at the 32nd hit
      for (int j = 0; j < n_src && em.cnt < 32; ++j) {
        const float x = src[j * 3], y = src[j * 3 + 1], z = src[j * 3 + 2];
        const float d = sqdist_ref(qx, qy, qz, q2, x, y, z, sqnorm3(x, y, z));
        if (!(d > r2)) em.push(j);
      }
    } else {
      unsigned dirty = 0u;
      int z = z0, y = y0 - 1, k = 0, kend = 0;
      for (;;) {
        if (k >= kend) {  // next (z, y) row: three x-adjacent cells are one contiguous run
          if (++y > y1) {
            y = y0;
            if (++z > z1) break;
          }
          const int row = (z * g.ny + y) * g.nx;
          k = (int)lds_u16(a_cs + (row + x0) * 2);
          kend = (int)lds_u16(a_cs + (row + x1) * 2);
          continue;
        }
        const float4 p = lds_f4(a_pts + k * 16);
        const float d = sqdist_ref(qx, qy, qz, q2, p.x, p.y, p.z, p.w);
        if (!(d > r2)) {
          const int j = (int)lds_u16(a_oidx + k * 2);
          const uint32_t a = a_bits + (j >> 5) * (GT * 4);
          sts_u32(a, lds_u32(a) | (1u << (j & 31)));
          dirty |= 1u << (j >> 5);
        }
        ++k;
      }
      while (dirty) {  // ascending words, ascending bits: index order
        const int w = __ffs(dirty) - 1;
        dirty &= dirty - 1;
        unsigned b = lds_u32(a_bits + w * (GT * 4));
        sts_u32(a_bits + w * (GT * 4), 0u);
        while (b && em.cnt < 32) {
          em.push(w * 32 + __ffs(b) - 1);
          b &= b - 1;
        }
      }
    }
    if (plan.rows) qcnt[q] = (unsigned short)max(em.cnt, 1);
    em.pad();
  }
  if (plan.rows == nullptr) return;
  // ---- sa1 plan (S == 1024): the bitmap memory is free now ----
  __syncthreads();  // every group of the cloud is written (block-visible), every count is in qcnt
  int* s_slot = reinterpret_cast<int*>(bits);      // [1024]
  int* s_gtiles = s_slot + 1024;                   // [8]
  int* s_used = s_gtiles + 8;                      // [8][32]
  int block_rows = 0;  // distinct rows of this warp's block of 128 centroids
  {
    // Rows are packed into 64-row HALF tiles (a centroid never straddles a half: each half of the CTA pools its own 64
    // columns), greedily in centroid order, independently per block of 128 centroids: one warp per block.  Lane l holds the
    // counts of centroids 4l .. 4l + 3 and their inclusive prefix sums P; a half that starts at element `start` with `base`
    // rows before it ends at the first element with P - base > 64 (ballot + ffs): one warp-uniform step per half tile.
    const int gq = tid >> 5, lane = tid & 31;
    const uint2 pk = reinterpret_cast<const uint2*>(qcnt)[gq * 32 + lane];
    const int cn[4] = {(int)(pk.x & 0xffffu), (int)(pk.x >> 16), (int)(pk.y & 0xffffu), (int)(pk.y >> 16)};
    int P[4];
    P[0] = cn[0]; P[1] = P[0] + cn[1]; P[2] = P[1] + cn[2]; P[3] = P[2] + cn[3];
    int incl = P[3];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    block_rows = total;
    const int off = incl - P[3];
#pragma unroll
    for (int e = 0; e < 4; ++e) P[e] += off;
    int slot[4] = {0, 0, 0, 0};
    int half = 0, base = 0, start = 0;
    unsigned usedw = 0u;  // lane t: tile t of the block, rows used in the low half | in the high half << 8
    for (;;) {
      int first = 4, before = 0;  // this lane's first element of the half that does not fit, and the rows in front of it
#pragma unroll
      for (int e = 3; e >= 0; --e)
        if (lane * 4 + e >= start && P[e] - base > 64) { first = e; before = P[e] - cn[e]; }
      const unsigned m = __ballot_sync(0xffffffffu, first < 4);
      const int L = m ? __ffs(m) - 1 : 0;
      const int brk = m ? L * 4 + __shfl_sync(0xffffffffu, first, L) : 128;
      const int next_base = m ? __shfl_sync(0xffffffffu, before, L) : total;
#pragma unroll
      for (int e = 0; e < 4; ++e)
        if (lane * 4 + e >= start && lane * 4 + e < brk) slot[e] = half * 64 + (P[e] - cn[e] - base);  // relative to the block's first tile
      if (lane == (half >> 1)) usedw |= (unsigned)(next_base - base) << ((half & 1) * 8);
      if (!m) break;
      ++half;
      base = next_base;
      start = brk;
    }
    *reinterpret_cast<int4*>(s_slot + gq * 128 + lane * 4) = make_int4(slot[0], slot[1], slot[2], slot[3]);
    s_used[gq * 32 + lane] = (int)usedw;
    if (lane == 0) s_gtiles[gq] = (half >> 1) + 1;
  }
  __syncthreads();
  {
    const int gg = tid >> 5, t = tid & 31;  // 256 threads: (group of 128 centroids, tile of the group)
    int o = 0;
    for (int i = 0; i < gg; ++i) o += s_gtiles[i];
    if (t < s_gtiles[gg]) plan.tile_used[c * PLAN_MAXT + o + t] = s_used[gg * 32 + t];
    if (tid == 0) {
      int tot = 0;
      for (int i = 0; i < 8; ++i) tot += s_gtiles[i];
      plan.tiles[c] = tot;
    }
  }
  // Row words.  Sparse blocks (the usual case, ~5 rows per centroid): one lane per ROW of the block's tiles -- consecutive
  // lanes write consecutive words (a lane per centroid would scatter 4-byte stores over as many sectors); the centroid of a
  // row is the last one whose first slot is not past it.  Dense blocks (balls that hold most of the cloud): one lane per
  // centroid, whole 128-byte groups.
  {
    const int gq = tid >> 5, lane = tid & 31;
    int t0 = 0;
    for (int i = 0; i < gq; ++i) t0 += s_gtiles[i];
    int* dstg = plan.rows + (int64_t)c * (PLAN_MAXT * 128) + t0 * 128;
    const int* slotg = s_slot + gq * 128;
    if (block_rows > 128 * 12) {
#pragma unroll 1
      for (int u = 0; u < 4; ++u) {
        const int sl = lane + 32 * u, sc = gq * 128 + sl, cnt = qcnt[sc], tag = sc << 10;
        const int4* row = reinterpret_cast<const int4*>(group + ((int64_t)c * GN + sc) * 32);
        int4 v[8];
#pragma unroll
        for (int q4 = 0; q4 < 8; ++q4) v[q4] = row[q4];
        int* dst = dstg + slotg[sl];
#pragma unroll
        for (int q4 = 0; q4 < 8; ++q4) {
          const int e[4] = {v[q4].x, v[q4].y, v[q4].z, v[q4].w};
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (q4 * 4 + k < cnt) dst[q4 * 4 + k] = e[k] | tag | (q4 * 4 + k == 0 ? 1 << 20 : 0);
        }
      }
    } else {
      const int npos = s_gtiles[gq] * 128;
      for (int p0 = lane; p0 < npos; p0 += 128) {
        int word[4];
        bool ok[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {  // four independent rows per lane in flight
          const int pos = p0 + 32 * u;
          int lo = 0;
#pragma unroll
          for (int step = 64; step >= 1; step >>= 1)
            if (slotg[lo + step] <= pos) lo += step;
          const int sc = gq * 128 + lo, k = pos - slotg[lo];
          ok[u] = pos < npos && k < (int)qcnt[sc];
          word[u] = ok[u] ? (group[((int64_t)c * GN + sc) * 32 + k] | (sc << 10) | (k == 0 ? 1 << 20 : 0)) : 0;  // neighbour | centroid << 10 | first-row flag << 20
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (ok[u]) dstg[p0 + 32 * u] = word[u];
      }
    }
  }
}

// ---- 3-NN: N fine points, S (<= 1024) coarse points ----
struct Top3L {  // ranked by (distance, index)
  float d0, d1, d2;
  int i0, i1, i2;
  __device__ __forceinline__ void init() {
    d0 = d1 = d2 = INFINITY;
    i0 = i1 = i2 = 0x7fffffff;
  }
  __device__ __forceinline__ static bool lt(float d, int s, float D, int I) { return d < D || (d == D && s < I); }
  __device__ __forceinline__ void push(float d, int s) {
    if (!(d <= d2)) return;  // (the common case once three close candidates are known)
    if (lt(d, s, d2, i2)) {
      if (lt(d, s, d1, i1)) {
        d2 = d1; i2 = i1;
        if (lt(d, s, d0, i0)) { d1 = d0; i1 = i0; d0 = d; i0 = s; }
        else { d1 = d; i1 = s; }
      } else { d2 = d; i2 = s; }
    }
  }
};

__global__ void __launch_bounds__(GT) three_nn_grid_kernel(const float* __restrict__ xyz1, const float* __restrict__ xyz2, int N, int S,
                                                           int target, int* __restrict__ nn_idx, float* __restrict__ nn_w) {
  extern __shared__ __align__(16) unsigned char smem[];
  GridSmem s;
  s.pts = reinterpret_cast<float4*>(smem);
  s.oidx = reinterpret_cast<unsigned short*>(s.pts + GN);
  s.cstart = s.oidx + GN;
  s.scratch = reinterpret_cast<int*>(s.cstart + GCELLS + 8);
  s.red = reinterpret_cast<float*>(s.scratch + GCELLS);
  unsigned short* order = reinterpret_cast<unsigned short*>(s.red + 64);  // [GN] queries in cell order
  unsigned short* qstart = order + GN;                                     // [GCELLS + 8]
  const int c = blockIdx.x, tid = threadIdx.x;
  const float* src = xyz2 + (int64_t)c * S * 3;
  Grid g;
  build_grid(src, S, 0.f, target, g, s);
  sort_queries(xyz1 + (int64_t)c * N * 3, N, g, s.scratch, qstart, reinterpret_cast<int*>(s.red), order);
  const uint32_t a_pts = smem_addr(s.pts), a_oidx = smem_addr(s.oidx), a_cs = smem_addr(s.cstart);
  const int rmax = max(g.nx, max(g.ny, g.nz));
  for (int ni = tid; ni < N; ni += GT) {
    const int n = order[ni];
    const float* q = xyz1 + ((int64_t)c * N + n) * 3;
    const float qx = q[0], qy = q[1], qz = q[2], q2 = sqnorm3(qx, qy, qz);
    Top3L t;
    t.init();
    if (g.uniform) {
      // all coarse points coincide: every distance is the same value, the (distance, index) ranking keeps indices 0, 1, 2
      for (int j = 0; j < 3; ++j) {
        const float x = src[j * 3], y = src[j * 3 + 1], z = src[j * 3 + 2];
        t.push(sqdist_ref(qx, qy, qz, q2, x, y, z, sqnorm3(x, y, z)), j);
      }
    } else {
      // the query may lie outside the coarse points' bounding box: its (unclamped) cell coordinates define the rings
      const int cx = (int)floorf((qx - g.minx) * g.inv_h), cy = (int)floorf((qy - g.miny) * g.inv_h), cz = (int)floorf((qz - g.minz) * g.inv_h);
      const float slack = fmaxf(g.E, 1e-5f * q2);
      {
        // rings 0 and 1 together: the 3 x 3 x 3 block as up to nine contiguous runs, one flattened loop
        const int x0 = max(0, cx - 1), x1 = min(g.nx - 1, cx + 1) + 1;
        const int y0 = max(0, cy - 1), y1 = min(g.ny - 1, cy + 1);
        const int z0 = max(0, cz - 1), z1 = min(g.nz - 1, cz + 1);
        if (x0 < x1 && y0 <= y1 && z0 <= z1) {
          int z = z0, y = y0 - 1, k = 0, kend = 0;
          for (;;) {
            if (k >= kend) {
              if (++y > y1) {
                y = y0;
                if (++z > z1) break;
              }
              const int row = (z * g.ny + y) * g.nx;
              k = (int)lds_u16(a_cs + (row + x0) * 2);
              kend = (int)lds_u16(a_cs + (row + x1) * 2);
              continue;
            }
            const float4 p = lds_f4(a_pts + k * 16);
            const float d = sqdist_ref(qx, qy, qz, q2, p.x, p.y, p.z, p.w);
            if (d <= t.d2) t.push(d, (int)lds_u16(a_oidx + k * 2));
            ++k;
          }
        }
      }
      for (int ring = 1;; ++ring) {
        if (ring >= 2) {  // the shell of cells at Chebyshev distance `ring`
          const int z0 = max(0, cz - ring), z1 = min(g.nz - 1, cz + ring), y0 = max(0, cy - ring), y1 = min(g.ny - 1, cy + ring);
          const int x0 = max(0, cx - ring), x1 = min(g.nx - 1, cx + ring);
          for (int z = z0; z <= z1; ++z)
            for (int y = y0; y <= y1; ++y) {
              const bool shell_zy = (abs(z - cz) == ring) || (abs(y - cy) == ring);
              const int row = (z * g.ny + y) * g.nx;
              if (shell_zy) {  // the whole x range of this row belongs to the shell
                if (x0 <= x1)
                  for (int k = (int)lds_u16(a_cs + (row + x0) * 2), ke = (int)lds_u16(a_cs + (row + x1 + 1) * 2); k < ke; ++k) {
                    const float4 p = lds_f4(a_pts + k * 16);
                    t.push(sqdist_ref(qx, qy, qz, q2, p.x, p.y, p.z, p.w), (int)lds_u16(a_oidx + k * 2));
                  }
              } else {         // only the two end cells at x = cx -+ ring
                for (int e = 0; e < 2; ++e) {
                  const int x = e == 0 ? cx - ring : cx + ring;
                  if (x < 0 || x >= g.nx) continue;
                  for (int k = (int)lds_u16(a_cs + (row + x) * 2), ke = (int)lds_u16(a_cs + (row + x + 1) * 2); k < ke; ++k) {
                    const float4 p = lds_f4(a_pts + k * 16);
                    t.push(sqdist_ref(qx, qy, qz, q2, p.x, p.y, p.z, p.w), (int)lds_u16(a_oidx + k * 2));
                  }
                }
              }
            }
        }
        // every unvisited point is more than (ring - 1e-4) cells away along some axis
        const float reach = ((float)ring - 1e-4f) * g.h;
        const bool covered = (cx - ring <= 0 && cx + ring >= g.nx - 1) && (cy - ring <= 0 && cy + ring >= g.ny - 1) && (cz - ring <= 0 && cz + ring >= g.nz - 1);
        if (covered) break;
        if (t.d2 < reach * reach - slack) break;
        if (ring > 2 * rmax + 64) break;  // (unreachable: `covered` ends the search; guards against a query far outside the box)
      }
    }
    const float r0 = 1.0f / (t.d0 + 1e-8f), r1 = 1.0f / (t.d1 + 1e-8f), r2 = 1.0f / (t.d2 + 1e-8f);
    const float norm = (r0 + r1) + r2;
    const int64_t o = ((int64_t)c * N + n) * 3;
    nn_idx[o] = t.i0; nn_idx[o + 1] = t.i1; nn_idx[o + 2] = t.i2;
    nn_w[o] = r0 / norm; nn_w[o + 1] = r1 / norm; nn_w[o + 2] = r2 / norm;
  }
}

constexpr int GRID_SMEM_BASE = GN * 16 + GN * 2 + (GCELLS + 8) * 2;

}  // namespace

// N (source points) must be <= 1024.  Same output as launch_ball_query.
// plan_rows != nullptr (S == 1024): also writes the plan of sa1's distinct rows (see Sa1PlanOut; the caller still runs the
// prefix sum over the clouds' tile counts, launch_sa1_plan_scan).
int launch_ball_query_grid(const float* xyz, const float* new_xyz, int n_clouds, int N, int S, double radius, int* group, cudaStream_t st,
                           int* plan_rows, int* plan_used, int* plan_tiles) {
  if (N > GN || N < 1) return -1;
  const float r2 = (float)(radius * radius);
  if (S > GN || (plan_rows && S != 1024)) return -1;
  constexpr int smem = GRID_SMEM_BASE + 32 * GT * 4 + 64 * 4 + GN * 2 + GN * 2;
  static PerDeviceOnce attr_done;
  if (smem_opt_in(attr_done, ball_query_grid_kernel, smem) != cudaSuccess) return -1;
  ball_query_grid_kernel<<<n_clouds, GT, smem, st>>>(xyz, new_xyz, N, S, (float)radius, r2, group, Sa1PlanOut{plan_rows, plan_used, plan_tiles});
  return 1;
}

// S (coarse points) must be in [3, 1024].  Same output as launch_three_nn.
int launch_three_nn_grid(const float* xyz1, const float* xyz2, int n_clouds, int N, int S, int* nn_idx, float* nn_w, cudaStream_t st) {
  if (S > GN || S < 3 || N > GN) return -1;
  constexpr int smem = GRID_SMEM_BASE + GCELLS * 4 + 64 * 4 + GN * 2 + (GCELLS + 8) * 2;
  static PerDeviceOnce attr_done;
  if (smem_opt_in(attr_done, three_nn_grid_kernel, smem) != cudaSuccess) return -1;
  const int target = S >= 1024 ? 8 : (S >= 256 ? 5 : 3);  // about two coarse points per cell in a filled cube
  three_nn_grid_kernel<<<n_clouds, GT, smem, st>>>(xyz1, xyz2, N, S, target, nn_idx, nn_w);
  return 1;
}

}  // namespace lsdm
