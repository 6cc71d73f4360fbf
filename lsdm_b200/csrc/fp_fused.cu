// Fused PointNet++ feature-propagation block (reference pointnet2_utils.py:266-316) for the level whose two weight
// matrices fit in shared memory (fp2: 64 + 256 -> 256 -> 128 over the 1024 points of every cloud):
//   h1  = ReLU( X.Wa^T + ba + sum_k w_k . Pb[nn_k] )        X  = fine-level features      [rows, CA]
//   out = ReLU( h1.W1^T + b1 )                              Pb = coarse features already projected by the coarse half of
//                                                                the first conv (linear, so projecting before interpolating
//                                                                is the same map at 1/4 of the rows)
// for 128 rows per tile with NOTHING but `out` leaving the SM: the unfused sequence (GEMM -> fp_combine -> GEMM) wrote and
// re-read two [rows, 256] fp32 intermediates, 2.4 GB of HBM traffic per denoising step at batch 64.
// One persistent CTA of 256 threads per SM; thread pair (t, t+128) owns tile row t&127 == TMEM lane and splits its columns:
//   X half-row (prefetched one tile ahead) -> TF32 -> TMEM A operand -> MMA1 (A from TMEM, Wa resident in smem, N = C1)
//   epilogue 1: tcgen05.ld chunk + bias + interpolated Pb chunk (gathered line-coalesced by all warps, exchanged through a
//               swizzled shared-memory tile) -> ReLU -> TF32 -> tcgen05.st in place
//   MMA2 (A from TMEM, W1 resident, N = C2) -> epilogue 2: bias + ReLU (+ TF32 rounding) -> global.
#include "kernels.cuh"
#include "tc_ptx.cuh"

namespace lsdm {

namespace {

using namespace tc;

struct FpArgs {
  const float* X;       // [rows, CA]
  const float* Wa;      // [C1, CA]
  const float* ba;      // [C1]
  const float* Pb;      // [n_clouds * S, C1]
  const int* nn_idx;    // [rows, 3] indices into the cloud's S coarse points
  const float* nn_w;    // [rows, 3]
  const float* W1;      // [C2, C1]
  const float* b1;      // [C2]
  float* out;           // [rows, C2]
  int n_tiles, S, n_shift;  // n_shift = log2(points per cloud at the fine level)
  int round_out;
  const int* x_perm;    // optional [rows]: row r reads X[(cloud of r) * N + x_perm[r]] (fine features kept in another row order)
};

template <int CA, int C1, int C2>
__global__ void __launch_bounds__(256, 1) fp_fused_kernel(FpArgs a) {
  constexpr int KB1 = CA / 32, KB2 = C1 / 32;
  static_assert(KB1 % 2 == 0 && (C1 / 32) % 2 == 0 && (C2 / 32) % 2 == 0, "column split between the two threads of a row");
  constexpr int WA_BYTES = C1 * CA * 4, W1_BYTES = C2 * C1 * 4;
  constexpr uint32_t COL_A1 = 0, COL_D1 = CA, COL_D2 = CA + C1;
  static_assert(COL_D2 + C2 <= 512, "TMEM budget");
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t s_bar;
  __shared__ uint32_t s_tmem;
  __shared__ float s_ba[C1], s_b1[C2];

  const int tid = threadIdx.x, warp = tid >> 5;
  const int rit = tid & 127, wq = warp & 3, half = tid >> 7;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sWa = base, sW1 = sWa + WA_BYTES;
  const uint32_t bar = smem_u32(&s_bar);

  if (warp == 0) tmem_alloc(smem_u32(&s_tmem), 512);
  if (tid == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  for (int q = tid; q < C1 * CA / 4; q += 256) {
    int n = q / (CA / 4), k4 = q % (CA / 4);
    float4 v = *reinterpret_cast<const float4*>(a.Wa + (int64_t)n * CA + k4 * 4);
    st_shared_v4(sWa + (k4 >> 3) * (C1 * 128) + sw128_off(n, k4 & 7), rna_tf32(v));
  }
  for (int q = tid; q < C2 * C1 / 4; q += 256) {
    int n = q / (C1 / 4), k4 = q % (C1 / 4);
    float4 v = *reinterpret_cast<const float4*>(a.W1 + (int64_t)n * C1 + k4 * 4);
    st_shared_v4(sW1 + (k4 >> 3) * (C2 * 128) + sw128_off(n, k4 & 7), rna_tf32(v));
  }
  for (int i = tid; i < C1; i += 256) s_ba[i] = a.ba[i];
  for (int i = tid; i < C2; i += 256) s_b1[i] = a.b1[i];
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s_tmem;
  const uint32_t tlane = tmem + ((uint32_t)(wq * 32) << 16);
  constexpr uint32_t idesc1 = umma_idesc_tf32(128, C1), idesc2 = umma_idesc_tf32(128, C2);
  uint32_t phase = 0;

  const int per = (a.n_tiles + gridDim.x - 1) / gridDim.x;
  const int t0 = blockIdx.x * per, t1 = (t0 + per < a.n_tiles) ? t0 + per : a.n_tiles;

  // Two roles per thread.  ROW role: thread pair (rit, half) owns tile row rit and half of its columns (TMEM lane == row).
  // LOADER role (the 3-neighbour gather): lanes 8j..8j+7 of a warp read 128 contiguous bytes of ONE coarse row, so every
  // L1 request is a full line (a row-per-lane gather would touch 32 lines per instruction and is L1-tag bound); the
  // interpolated 32-column chunk is handed to the row role through a swizzled shared-memory tile, one per column half.
  constexpr int XQ = CA / 4 / 2;
  float4 xrow[XQ];  // ROW role: this thread's half of the X row, prefetched one tile ahead
  int li[4][3];     // LOADER role: neighbours / weights of tile rows (tid>>3) + 32 i, prefetched one tile ahead
  float lw[4][3];
  const int lr0 = tid >> 3, lc4 = tid & 7;
  auto prefetch = [&](int tile) {
    if (tile < t1) {
      const int64_t row = (int64_t)tile * 128 + rit;
      const int64_t xr = a.x_perm ? (((row >> a.n_shift) << a.n_shift) + a.x_perm[row]) : row;
      const float4* p = reinterpret_cast<const float4*>(a.X + xr * CA) + half * XQ;
#pragma unroll
      for (int q = 0; q < XQ; ++q) xrow[q] = p[q];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int64_t r = (int64_t)tile * 128 + lr0 + 32 * i;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          li[i][j] = a.nn_idx[r * 3 + j];
          lw[i][j] = a.nn_w[r * 3 + j];
        }
      }
    }
  };
  float* stage = reinterpret_cast<float*>(smem_raw + (base - smem_u32(smem_raw)) + WA_BYTES + W1_BYTES);  // 2 x [128][32] floats
  constexpr int CH1 = C1 / 32 / 2;  // 32-column chunks per column half
  // Software pipeline over tiles (round 2b): the A operand of tile i + 1 is staged into tensor memory and its first MMA is issued
  // while tile i's second MMA / output epilogue run (its columns are free by then), and the first gather step of tile i + 1 is
  // issued before tile i's second MMA, so its L2 round trip hides behind both.
  //   [stage A1(i+1) | MMA2(i)] -> MMA1(i+1) issued -> [epilogue 2(i) | MMA1(i+1)] -> epilogue 1(i+1) ...
  const float* gp[4][3];  // loader state of the tile whose epilogue 1 runs / runs next
  float gw[4][3];
  float4 nb[2][4][3];     // gather registers: step s covers chunk s of BOTH column halves (hh = 0, 1): 4 rows x 2 halves x 3 neighbours
  auto gather = [&](int s) {
#pragma unroll
    for (int hh = 0; hh < 2; ++hh)
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) nb[hh][i][j] = *reinterpret_cast<const float4*>(gp[i][j] + (hh * CH1 + s) * 32);
  };
  auto stage_a1 = [&]() {  // own half-row -> TMEM (the +half-ulp of rna_tf32_mma is a no-op on already rounded inputs)
#pragma unroll
    for (int kl = 0; kl < KB1 / 2; ++kl) {
      uint32_t v[32];
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float4 p = xrow[kl * 8 + q];
        v[q * 4 + 0] = rna_tf32_mma(p.x);
        v[q * 4 + 1] = rna_tf32_mma(p.y);
        v[q * 4 + 2] = rna_tf32_mma(p.z);
        v[q * 4 + 3] = rna_tf32_mma(p.w);
      }
      tmem_st32(tlane + COL_A1 + (half * (KB1 / 2) + kl) * 32, v);
    }
    tmem_st_wait();
  };
  auto loader_state = [&](int tile) {  // from li / lw (the prefetch that follows overwrites them with the next tile's)
    const int64_t cbase = (((int64_t)tile * 128) >> a.n_shift) * a.S;  // a tile never straddles two clouds
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        gp[i][j] = a.Pb + (cbase + li[i][j]) * C1 + lc4 * 4;
        gw[i][j] = lw[i][j];
      }
  };
  auto issue_mma1 = [&]() {  // D1 = X . Wa^T
    if (tid == 0) {
      tc_fence_after();
#pragma unroll
      for (int kb = 0; kb < KB1; ++kb) {
        const uint64_t db = umma_desc_sw128(sWa + kb * (C1 * 128));
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
          umma_tf32_ts(tmem + COL_D1, tmem + COL_A1 + kb * 32 + kk * 8, db + (uint64_t)(kk * 2), idesc1, (kb | kk) != 0 ? 1u : 0u);
      }
      umma_commit(bar);
    }
  };
  prefetch(t0);
  if (t0 < t1) {
    stage_a1();
    loader_state(t0);
    prefetch(t0 + 1);
    tc_fence_before();
    __syncthreads();
    issue_mma1();
    gather(0);  // travels while the MMA runs
  }
  for (int tile = t0; tile < t1; ++tile) {
    const int64_t row = (int64_t)tile * 128 + rit;
    const bool has_next = tile + 1 < t1;
    mbar_wait(bar, phase);
    phase ^= 1;
    tc_fence_after();
    // ---- epilogue 1: + bias + interpolated coarse projection, ReLU, TF32 -> A operand of layer 2 (in place) ----
#pragma unroll 1
    for (int s = 0; s < CH1; ++s) {
      // LOADER: interpolate (same operation order as the unfused fp_combine) and publish
#pragma unroll
      for (int hh = 0; hh < 2; ++hh)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 x0 = nb[hh][i][0], x1 = nb[hh][i][1], x2 = nb[hh][i][2];
          const float w0 = gw[i][0], w1 = gw[i][1], w2 = gw[i][2];
          float4 o;
          o.x = fmaf(x2.x, w2, fmaf(x1.x, w1, x0.x * w0));
          o.y = fmaf(x2.y, w2, fmaf(x1.y, w1, x0.y * w0));
          o.z = fmaf(x2.z, w2, fmaf(x1.z, w1, x0.z * w0));
          o.w = fmaf(x2.w, w2, fmaf(x1.w, w1, x0.w * w0));
          const int r = lr0 + 32 * i;
          *reinterpret_cast<float4*>(stage + hh * (128 * 32) + r * 32 + ((lc4 ^ (r & 7)) << 2)) = o;
        }
      __syncthreads();
      if (s + 1 < CH1) {
        gather(s + 1);  // next step's rows are in flight during this step's TMEM round trip
      } else if (has_next) {
        loader_state(tile + 1);  // (li / lw hold tile + 1's neighbours since the last prefetch)
        gather(0);               // the next tile's first step: in flight during this tile's second MMA and output epilogue
      }
      // ROW: this thread's 32 columns of chunk s
      const int c0 = (half * CH1 + s) * 32;
      uint32_t v[32];
      tmem_ld32(tlane + COL_D1 + c0, v);
      tmem_ld_wait();
      const float* srow = stage + half * (128 * 32) + rit * 32;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float4 it = *reinterpret_cast<const float4*>(srow + ((q ^ (rit & 7)) << 2));
        const float iv[4] = {it.x, it.y, it.z, it.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float pa = __uint_as_float(v[q * 4 + e]) + s_ba[c0 + q * 4 + e];  // == the unfused GEMM's epilogue
          v[q * 4 + e] = rna_tf32_mma(fmaxf(pa + iv[e], 0.0f));                    // == fp_combine
        }
      }
      tmem_st32(tlane + COL_D1 + c0, v);
      __syncthreads();  // the staging tiles are rewritten by the next step
    }
    tmem_st_wait();
    tc_fence_before();
    __syncthreads();
    // ---- layer 2: D2 = h1 . W1^T ----
    if (tid == 0) {
      tc_fence_after();
#pragma unroll
      for (int kb = 0; kb < KB2; ++kb) {
        const uint64_t db = umma_desc_sw128(sW1 + kb * (C2 * 128));
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
          umma_tf32_ts(tmem + COL_D2, tmem + COL_D1 + kb * 32 + kk * 8, db + (uint64_t)(kk * 2), idesc2, (kb | kk) != 0 ? 1u : 0u);
      }
      umma_commit(bar);
    }
    // while it runs: the next tile's A operand (its columns were released by this tile's first MMA) and the prefetch of tile + 2
    if (has_next) {
      stage_a1();
      prefetch(tile + 2);
    }
    mbar_wait(bar, phase);
    phase ^= 1;
    tc_fence_after();
    if (has_next) {
      tc_fence_before();
      __syncthreads();  // every thread's A1 stores are complete; D1 is free (this tile's second MMA has read it)
      issue_mma1();
    }
    // ---- epilogue 2: bias + ReLU (+ rounding for the next tensor-core consumer) -> global, 128 B per thread per chunk ----
    float* orow = a.out + row * C2;
#pragma unroll 1
    for (int cl = 0; cl < C2 / 32 / 2; ++cl) {
      const int c0 = (half * (C2 / 32 / 2) + cl) * 32;
      uint32_t v[32];
      tmem_ld32(tlane + COL_D2 + c0, v);
      tmem_ld_wait();
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        float o[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float r = fmaxf(__uint_as_float(v[q * 4 + e]) + s_b1[c0 + q * 4 + e], 0.0f);
          o[e] = a.round_out ? rna_tf32_fin(r) : r;
        }
        *reinterpret_cast<float4*>(orow + c0 + q * 4) = make_float4(o[0], o[1], o[2], o[3]);
      }
    }
    tc_fence_before();  // the next tile's second MMA reuses these columns
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

// ------------------------------------------------------------------------------------------------------------------
// Fused LAST level + head (reference pointnet2_utils.py:290-311, pointnet2.py:71-79): fp1 has no fine-level features, so
//   h0 = ReLU(b1 + sum_k w_k . Pb[nn_k])  ->  W2 -> ReLU -> W3 -> ReLU -> Wh (conv1 + bn1) -> ReLU -> conv2 (128 -> 3)
// for 128 rows per tile; only the [rows, 3] result leaves the SM.  Same thread roles as fp_fused_kernel (coalesced
// 3-neighbour gather by all warps -> swizzled shared tile -> row threads), then three A-from-TMEM MMAs ping-ponging between
// two 128-column TMEM regions with in-place epilogues; the three 128x128 weight matrices stay resident in shared memory.
// ------------------------------------------------------------------------------------------------------------------
struct Fp1Const {
  float b1[128], b2[128], b3[128], bh[128];
  float wc[3 * 128];
  float bc[3];
};
template <int V>
struct IntK { static constexpr int value = V; };

__global__ void __launch_bounds__(256, 1) fp1_fused_kernel(const float* __restrict__ Pb, const int* __restrict__ nn_idx,
                                                           const float* __restrict__ nn_w, const float* __restrict__ W2,
                                                           const float* __restrict__ W3, const float* __restrict__ Wh,
                                                           float* __restrict__ out, int n_tiles, int S, int n_shift,
                                                           const __grid_constant__ Fp1Const k) {
  constexpr int WB = 128 * 128 * 4;
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t s_bar;
  __shared__ uint32_t s_tmem;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int rit = tid & 127, wq = warp & 3, half = tid >> 7;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sW[3] = {base, base + WB, base + 2 * WB};
  float* stage = reinterpret_cast<float*>(smem_raw + (base - smem_u32(smem_raw)) + 3 * WB);  // 2 x [128][32] floats
  const uint32_t bar = smem_u32(&s_bar);
  if (warp == 0) tmem_alloc(smem_u32(&s_tmem), 256);
  if (tid == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  const float* Ws[3] = {W2, W3, Wh};
  for (int m = 0; m < 3; ++m)
    for (int q = tid; q < 128 * 128 / 4; q += 256) {
      int n = q >> 5, k4 = q & 31;
      float4 v = *reinterpret_cast<const float4*>(Ws[m] + (int64_t)n * 128 + k4 * 4);
      st_shared_v4(sW[m] + (k4 >> 3) * (128 * 128) + sw128_off(n, k4 & 7), rna_tf32(v));
    }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s_tmem;
  const uint32_t tlane = tmem + ((uint32_t)(wq * 32) << 16);
  constexpr uint32_t idesc = umma_idesc_tf32(128, 128);

  const int per = (n_tiles + gridDim.x - 1) / gridDim.x;
  const int t0 = blockIdx.x * per, t1 = (t0 + per < n_tiles) ? t0 + per : n_tiles;
  const int lr0 = tid >> 3, lc4 = tid & 7;

  auto run = [&](auto HC) {
    constexpr int H = decltype(HC)::value;  // this thread's column half, a compile-time constant of the code path so that
                                            // the per-channel vectors stay constant-bank operands
    uint32_t phase = 0;
    auto mma_layer = [&](int m, uint32_t col_a, uint32_t col_d) {
      if (tid == 0) {
        tc_fence_after();
#pragma unroll
        for (int kb = 0; kb < 4; ++kb) {
          const uint64_t db = umma_desc_sw128(sW[m] + kb * (128 * 128));
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            umma_tf32_ts(tmem + col_d, tmem + col_a + kb * 32 + kk * 8, db + (uint64_t)(kk * 2), idesc, (kb | kk) != 0 ? 1u : 0u);
        }
        umma_commit(bar);
      }
      mbar_wait(bar, phase);
      phase ^= 1;
      tc_fence_after();
    };
    // loader state: neighbour row pointers / weights of tile rows lr0 + 32 i, and the gathered float4s of one step
    const float* gp[4][3];
    float gw[4][3];
    float4 nb[2][4][3];
    auto load_nbrs = [&](int tile) {
      if (tile < t1) {
        const int64_t cbase = (((int64_t)tile * 128) >> n_shift) * S;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int64_t r = (int64_t)tile * 128 + lr0 + 32 * i;
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            gp[i][j] = Pb + (cbase + nn_idx[r * 3 + j]) * 128 + lc4 * 4;
            gw[i][j] = nn_w[r * 3 + j];
          }
        }
      }
    };
    auto gather = [&](int s) {
#pragma unroll
      for (int hh = 0; hh < 2; ++hh)
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 3; ++j) nb[hh][i][j] = *reinterpret_cast<const float4*>(gp[i][j] + (hh * 2 + s) * 32);
    };
    load_nbrs(t0);
    if (t0 < t1) gather(0);
    for (int tile = t0; tile < t1; ++tile) {
      const int64_t row = (int64_t)tile * 128 + rit;
      // ---- h0 = ReLU(b1 + interpolation) -> TF32 -> TMEM columns [0,128): two steps of 32 columns per column half ----
#pragma unroll
      for (int s = 0; s < 2; ++s) {
#pragma unroll
        for (int hh = 0; hh < 2; ++hh)
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float4 x0 = nb[hh][i][0], x1 = nb[hh][i][1], x2 = nb[hh][i][2];
            const float w0 = gw[i][0], w1 = gw[i][1], w2 = gw[i][2];
            float4 o;
            o.x = fmaf(x2.x, w2, fmaf(x1.x, w1, x0.x * w0));
            o.y = fmaf(x2.y, w2, fmaf(x1.y, w1, x0.y * w0));
            o.z = fmaf(x2.z, w2, fmaf(x1.z, w1, x0.z * w0));
            o.w = fmaf(x2.w, w2, fmaf(x1.w, w1, x0.w * w0));
            const int r = lr0 + 32 * i;
            *reinterpret_cast<float4*>(stage + hh * (128 * 32) + r * 32 + ((lc4 ^ (r & 7)) << 2)) = o;
          }
        __syncthreads();
        if (s == 0) {
          gather(1);
        } else {
          load_nbrs(tile + 1);               // the next tile's first step travels during this tile's three MMAs
          if (tile + 1 < t1) gather(0);
        }
        uint32_t v[32];
        const float* srow = stage + H * (128 * 32) + rit * 32;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 it = *reinterpret_cast<const float4*>(srow + ((q ^ (rit & 7)) << 2));
          const int c = (H * 2 + s) * 32 + q * 4;
          v[q * 4 + 0] = rna_tf32_mma(fmaxf(k.b1[c + 0] + it.x, 0.0f));
          v[q * 4 + 1] = rna_tf32_mma(fmaxf(k.b1[c + 1] + it.y, 0.0f));
          v[q * 4 + 2] = rna_tf32_mma(fmaxf(k.b1[c + 2] + it.z, 0.0f));
          v[q * 4 + 3] = rna_tf32_mma(fmaxf(k.b1[c + 3] + it.w, 0.0f));
        }
        tmem_st32(tlane + (H * 2 + s) * 32, v);
        if (s == 0) __syncthreads();  // the staging tiles are rewritten by step 1 (the barrier below covers step 1)
      }
      tmem_st_wait();
      tc_fence_before();
      __syncthreads();
      mma_layer(0, 0, 128);  // Y = h0 . W2^T
#pragma unroll
      for (int kl = 0; kl < 2; ++kl) {
        constexpr int KB0 = H * 2;
        const int kb = KB0 + kl;
        uint32_t v[32];
        tmem_ld32(tlane + 128 + kb * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 32; ++e) v[e] = rna_tf32_mma(fmaxf(__uint_as_float(v[e]) + k.b2[kb * 32 + e], 0.0f));
        tmem_st32(tlane + 128 + kb * 32, v);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncthreads();
      mma_layer(1, 128, 0);  // X = Y . W3^T
#pragma unroll
      for (int kl = 0; kl < 2; ++kl) {
        constexpr int KB0 = H * 2;
        const int kb = KB0 + kl;
        uint32_t v[32];
        tmem_ld32(tlane + kb * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 32; ++e) v[e] = rna_tf32_mma(fmaxf(__uint_as_float(v[e]) + k.b3[kb * 32 + e], 0.0f));
        tmem_st32(tlane + kb * 32, v);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncthreads();
      mma_layer(2, 0, 128);  // Y = X . Wh^T   (conv1 with bn1 folded)
      // ---- head: relu(Y + bh) . conv2^T + bc (128 -> 3), fp32 CUDA cores; the two column halves meet in shared memory ----
      float o0 = 0.f, o1 = 0.f, o2 = 0.f;
#pragma unroll
      for (int kl = 0; kl < 2; ++kl) {
        constexpr int KB0 = H * 2;
        const int kb = KB0 + kl;
        uint32_t v[32];
        tmem_ld32(tlane + 128 + kb * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          const int ch = kb * 32 + e;
          const float x = fmaxf(__uint_as_float(v[e]) + k.bh[ch], 0.0f);
          o0 = fmaf(k.wc[ch], x, o0);
          o1 = fmaf(k.wc[128 + ch], x, o1);
          o2 = fmaf(k.wc[256 + ch], x, o2);
        }
      }
      if (H == 1) {
        stage[rit * 4 + 0] = o0;
        stage[rit * 4 + 1] = o1;
        stage[rit * 4 + 2] = o2;
      }
      tc_fence_before();
      __syncthreads();
      if (H == 0) {
        out[row * 3 + 0] = (o0 + stage[rit * 4 + 0]) + k.bc[0];
        out[row * 3 + 1] = (o1 + stage[rit * 4 + 1]) + k.bc[1];
        out[row * 3 + 2] = (o2 + stage[rit * 4 + 2]) + k.bc[2];
      }
      __syncthreads();  // the next tile's loaders rewrite the staging tiles
    }
  };
  if (half == 0) {
    run(IntK<0>{});
  } else {
    run(IntK<1>{});
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 256);
}

}  // namespace

// fp2 of the reference backbone: CA = 64 (l1 features), C1 = 256, C2 = 128; N (fine points per cloud) must be a power of
// two and a multiple of 128.  Returns the number of launches (1) or -1 when the shape is not covered.
int launch_fp_fused(const float* X, int CA, const float* Wa, const float* ba, const float* Pb, const int* nn_idx, const float* nn_w,
                    const float* W1, const float* b1, int n_clouds, int N, int S, int C1, int C2, float* out, int round_out,
                    cudaStream_t st, const int* x_perm) {
  if (CA != 64 || C1 != 256 || C2 != 128 || N < 128 || (N & (N - 1)) != 0) return -1;
  FpArgs a{X, Wa, ba, Pb, nn_idx, nn_w, W1, b1, out, n_clouds * (N / 128), S, __builtin_ctz(N), round_out, x_perm};
  constexpr int smem = 256 * 64 * 4 + 128 * 256 * 4 + 2 * 128 * 32 * 4 + 1024;
  static PerDeviceOnce attr_done;
  if (smem_opt_in(attr_done, fp_fused_kernel<64, 256, 128>, smem) != cudaSuccess) return -1;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int grid = sms < a.n_tiles ? sms : a.n_tiles;
  fp_fused_kernel<64, 256, 128><<<grid, 256, smem, st>>>(a);
  return 1;
}

}  // namespace lsdm

namespace lsdm {

// fp1 + head.  h_consts: host copy of [b2(128) | b3(128) | bh(128) | conv2.weight(3x128) | conv2.bias(3)];
// h_b1: host copy of the first conv's folded bias [128].  N (points per cloud) must be a power of two and a multiple of 128.
int launch_fp1_fused(const float* Pb, const int* nn_idx, const float* nn_w, const float* h_b1, const float* W2, const float* W3,
                     const float* Wh, const float* h_consts, int n_clouds, int N, int S, float* out, cudaStream_t st) {
  if (N < 128 || (N & (N - 1)) != 0) return -1;
  Fp1Const k;
  for (int i = 0; i < 128; ++i) {
    k.b1[i] = h_b1[i];
    k.b2[i] = h_consts[i];
    k.b3[i] = h_consts[128 + i];
    k.bh[i] = h_consts[256 + i];
  }
  for (int i = 0; i < 384; ++i) k.wc[i] = h_consts[384 + i];
  for (int i = 0; i < 3; ++i) k.bc[i] = h_consts[768 + i];
  constexpr int smem = 3 * 128 * 128 * 4 + 2 * 128 * 32 * 4 + 1024;
  static PerDeviceOnce attr_done;
  if (smem_opt_in(attr_done, fp1_fused_kernel, smem) != cudaSuccess) return -1;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int n_tiles = n_clouds * (N / 128);
  const int grid = n_tiles < sms ? n_tiles : sms;
  fp1_fused_kernel<<<grid, 256, smem, st>>>(Pb, nn_idx, nn_w, W2, W3, Wh, out, n_tiles, S, __builtin_ctz(N), k);
  return 1;
}

}  // namespace lsdm
