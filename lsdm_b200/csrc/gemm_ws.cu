// Persistent, warp-specialised tcgen05 implementation of the generic linear-layer GEMM (gemm.cuh):
//   C[M,N] = act(A[M,K] . W[N,K]^T + bias),  TF32 (operands rounded to nearest) or 3xTF32, fp32 accumulate in TMEM.
// One CTA per SM loops over 128 x BN output tiles.  Thirteen warps, three roles:
//   warps 0-3  producers : ld.global (next k-block prefetched in registers) -> cvt.rna.tf32 [-> hi/lo split] ->
//                          swizzled st.shared into an NSTAGE ring (K-major SWIZZLE_128B k-blocks) -> mbarrier "full";
//                          pre-rounded / pre-split operands skip the registers: TMA (cp.async.bulk.tensor, one thread,
//                          SWIZZLE_128B boxes of 128 x 32 and BN x 32 floats, byte-counted mbarrier) or cp.async (batched GEMMs)
//   warp  12   MMA issuer: one thread; waits "full", issues tcgen05.mma (M=128, N=BN, K=8), tcgen05.commit -> "empty";
//                          after the last k-block commits the tile's accumulator -> "acc_full"
//   warps 4-11 epilogue  : tcgen05.ld of the finished accumulator (two TMEM buffers alternate, so the epilogue of tile i
//                          overlaps the loads + MMAs of tile i+1), bias / activation, then either a shared-memory
//                          transpose for coalesced 128-bit stores or the 32-row group max (redux.sync) -> "acc_empty"
#include <cuda.h>  // CUtensorMap (the encoder is fetched with cudaGetDriverEntryPoint: no link against libcuda)

#include "gemm.cuh"
#include "tc_ptx.cuh"

namespace lsdm {

namespace {

using namespace tc;

constexpr int WBM = 128, WBK = 32;
constexpr int W_A_STAGE = WBM * 128;  // bytes of one A k-block
constexpr int NPROD = 128;             // producer threads (4 warps)
constexpr int NEPI = 256;              // epilogue threads: 8 warps, two per TMEM lane quarter (each takes half of the columns);
                                       // with 4 warps the epilogue (not the MMA) bounded every tile
constexpr int WS_THREADS = NPROD + NEPI + 32;  // producers | epilogue warps | 1 MMA-issuer warp
constexpr int WSTG = 36;              // epilogue transpose row stride (floats)

// shared-memory plan: the (SPLIT, BN=256) ring is 2 x 96 KB, which leaves room for only four staging buffers
__host__ __device__ constexpr int ws_epi_warps(int bn, bool split) { return (split && bn == 256) ? 4 : 8; }
__host__ __device__ constexpr int ws_nstage(int bn, bool split, bool async) {
  return split ? (async ? (bn <= 32 ? 4 : (bn <= 64 ? 3 : 2)) : 2) : ((async && bn < 256) ? 4 : 3);
}

struct TmaMaps {
  CUtensorMap a, w, a_lo, w_lo;  // A / W (hi planes) and the 3xTF32 residual planes
};

template <int BN, bool SPLIT, bool ASYNC>
__global__ void __launch_bounds__(WS_THREADS, 1) gemm_ws_kernel(GemmArgs g, int tiles_m, int tiles_n, int total_tiles, int use_tma,
                                                                const __grid_constant__ TmaMaps maps) {
  constexpr int NSTAGE = ws_nstage(BN, SPLIT, ASYNC);
  constexpr int EPW = ws_epi_warps(BN, SPLIT);
  constexpr int W_STAGE = BN * 128;
  constexpr int HALF = W_A_STAGE + W_STAGE;          // [A | W]; SPLIT appends [A_lo | W_lo]
  constexpr int STAGE = HALF * (SPLIT ? 2 : 1);
  constexpr uint32_t TCOLS_PER = BN <= 32 ? 32 : (BN <= 64 ? 64 : (BN <= 128 ? 128 : 256));
  constexpr int A_PER = WBM * 8 / NPROD, W_PER = BN * 8 / NPROD;
  constexpr int PW = NPROD / 32;  // producer warps
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t s_full[NSTAGE], s_empty[NSTAGE], s_accf[2], s_acce[2];
  __shared__ uint32_t s_tmem;
  __shared__ float s_bias[1024];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
  const int nkb = g.K / WBK;

  if (warp == 0) tmem_alloc(smem_u32(&s_tmem), 2 * TCOLS_PER);
  if (tid == 0) {
    for (int i = 0; i < NSTAGE; ++i) {
      mbar_init(smem_u32(&s_full[i]), (ASYNC && use_tma) ? 1 : NPROD);
      mbar_init(smem_u32(&s_empty[i]), 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&s_accf[i]), 1);
      mbar_init(smem_u32(&s_acce[i]), EPW * 32);
    }
    fence_mbar_init();
  }
  if (g.bias_mode == 1)
    for (int i = tid; i < g.N; i += WS_THREADS) s_bias[i] = g.bias[i];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s_tmem;

  auto decode = [&](int t, int& m0, int& n0, int& z) {
    int per_z = tiles_m * tiles_n;
    z = t / per_z;
    int r = t - z * per_z;
    m0 = (r / tiles_n) * WBM;
    n0 = (r % tiles_n) * BN;
  };

  if (warp < PW && ASYNC && use_tma) {
    // ============ producer, pre-rounded operands, TMA: one thread issues the bulk tensor copies of every stage ============
    if (tid == 0) {
      tma_prefetch_desc(&maps.a);
      tma_prefetch_desc(&maps.w);
      int t = blockIdx.x, kb = 0;
      uint32_t it = 0;
      while (t < total_tiles) {
        const uint32_t s = it % NSTAGE;
        if (it >= (uint32_t)NSTAGE) mbar_wait(smem_u32(&s_empty[s]), ((it / NSTAGE) & 1u) ^ 1u);
        int m0, n0, z;
        decode(t, m0, n0, z);
        const uint32_t sA = base + s * STAGE, sW = sA + W_A_STAGE, full = smem_u32(&s_full[s]);
        const int k0 = kb * WBK;
        mbar_arrive_expect_tx(full, (uint32_t)STAGE);
        tma_load_2d(sA, &maps.a, k0, m0, full);  // rows past M are zero-filled
        tma_load_2d(sW, &maps.w, k0, n0, full);
        if (SPLIT) {
          tma_load_2d(sA + HALF, &maps.a_lo, k0, m0, full);
          tma_load_2d(sW + HALF, &maps.w_lo, k0, n0, full);
        }
        if (++kb == nkb) {
          kb = 0;
          t += gridDim.x;
        }
        ++it;
      }
    }
  } else if (warp < PW && ASYNC) {
    // ============ producers, pre-rounded operands: cp.async straight into the swizzled ring, NSTAGE k-blocks in flight ============
    int t = blockIdx.x, kb = 0;
    uint32_t it = 0;
    while (t < total_tiles) {
      const uint32_t s = it % NSTAGE;
      if (it >= (uint32_t)NSTAGE) mbar_wait(smem_u32(&s_empty[s]), ((it / NSTAGE) & 1u) ^ 1u);
      int m0, n0, z;
      decode(t, m0, n0, z);
      const float* __restrict__ A = g.A + (int64_t)z * g.strideA;
      const float* __restrict__ W = g.W + (int64_t)z * g.strideW;
      const uint32_t sA = base + s * STAGE, sW = sA + W_A_STAGE;
      const int k0 = kb * WBK;
#pragma unroll
      for (int i = 0; i < A_PER; ++i) {
        int q = tid + i * NPROD;
        int r = q >> 3, c = q & 7;
        int m = m0 + r;
        m = m < g.M ? m : g.M - 1;
        const int64_t go = (int64_t)m * g.lda + k0 + c * 4;
        const uint32_t so = sw128_off(r, c);
        cp_async16(sA + so, A + go);
        if (SPLIT) cp_async16(sA + HALF + so, g.A_lo + (int64_t)z * g.strideA + go);
      }
#pragma unroll
      for (int i = 0; i < W_PER; ++i) {
        int q = tid + i * NPROD;
        int r = q >> 3, c = q & 7;
        const int64_t go = (int64_t)(n0 + r) * g.ldw + k0 + c * 4;
        const uint32_t so = sw128_off(r, c);
        cp_async16(sW + so, W + go);
        if (SPLIT) cp_async16(sW + HALF + so, g.W_lo + (int64_t)z * g.strideW + go);
      }
      cp_async_mbar_arrive(smem_u32(&s_full[s]));
      if (++kb == nkb) {
        kb = 0;
        t += gridDim.x;
      }
      ++it;
    }
    cp_async_wait<0>();
  } else if (warp < PW) {
    // =============================== producers ===============================
    float4 ra[A_PER], rw[W_PER];
    auto fetch = [&](int t, int kb) {
      int m0, n0, z;
      decode(t, m0, n0, z);
      const float* __restrict__ A = g.A + (int64_t)z * g.strideA;
      const float* __restrict__ W = g.W + (int64_t)z * g.strideW;
      const int k0 = kb * WBK;
#pragma unroll
      for (int i = 0; i < A_PER; ++i) {
        int q = tid + i * NPROD;
        int r = q >> 3, c = q & 7;
        int m = m0 + r;
        m = m < g.M ? m : g.M - 1;
        ra[i] = *reinterpret_cast<const float4*>(A + (int64_t)m * g.lda + k0 + c * 4);
      }
#pragma unroll
      for (int i = 0; i < W_PER; ++i) {
        int q = tid + i * NPROD;
        int r = q >> 3, c = q & 7;
        rw[i] = *reinterpret_cast<const float4*>(W + (int64_t)(n0 + r) * g.ldw + k0 + c * 4);
      }
    };
    auto sub4 = [](float4 a, float4 b) { return make_float4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); };
    int t = blockIdx.x, kb = 0;
    if (t < total_tiles) fetch(t, 0);
    uint32_t it = 0;
    while (t < total_tiles) {
      const uint32_t s = it % NSTAGE;
      if (it >= (uint32_t)NSTAGE) mbar_wait(smem_u32(&s_empty[s]), ((it / NSTAGE) & 1u) ^ 1u);
      const uint32_t sA = base + s * STAGE, sW = sA + W_A_STAGE;
#pragma unroll
      for (int i = 0; i < A_PER; ++i) {
        int q = tid + i * NPROD;
        uint32_t off = sw128_off(q >> 3, q & 7);
        float4 hi = rna_tf32(ra[i]);
        st_shared_v4(sA + off, hi);
        if (SPLIT) st_shared_v4(sA + HALF + off, rna_tf32(sub4(ra[i], hi)));
      }
#pragma unroll
      for (int i = 0; i < W_PER; ++i) {
        int q = tid + i * NPROD;
        uint32_t off = sw128_off(q >> 3, q & 7);
        float4 hi = rna_tf32(rw[i]);
        st_shared_v4(sW + off, hi);
        if (SPLIT) st_shared_v4(sW + HALF + off, rna_tf32(sub4(rw[i], hi)));
      }
      // advance and prefetch the next k-block (possibly of the next tile) before publishing this one
      if (++kb == nkb) {
        kb = 0;
        t += gridDim.x;
      }
      if (t < total_tiles) fetch(t, kb);
      fence_proxy_async();
      mbar_arrive(smem_u32(&s_full[s]));
      ++it;
    }
  } else if (warp == PW + NEPI / 32) {
    // =============================== MMA issuer ===============================
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_tf32(WBM, BN);
      uint32_t it = 0, ti = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++ti) {
        const uint32_t b = ti & 1u;
        if (ti >= 2) mbar_wait(smem_u32(&s_acce[b]), ((ti >> 1) - 1u) & 1u);
        tc_fence_after();
        const uint32_t d = tmem + b * TCOLS_PER;
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const uint32_t s = it % NSTAGE;
          mbar_wait(smem_u32(&s_full[s]), (it / NSTAGE) & 1u);
          if (ASYNC) fence_proxy_async();  // cp.async (generic proxy) writes -> tensor core (async proxy)
          tc_fence_after();
          const uint32_t sA = base + s * STAGE, sW = sA + W_A_STAGE;
          const uint64_t da = umma_desc_sw128(sA), db = umma_desc_sw128(sW);
#pragma unroll
          for (int kk = 0; kk < WBK / 8; ++kk) {
            const uint64_t o = (uint64_t)(kk * 2);
            if (SPLIT) {
              const uint64_t dal = umma_desc_sw128(sA + HALF), dbl = umma_desc_sw128(sW + HALF);
              umma_tf32_ss(d, dal + o, db + o, idesc, (kb | kk) != 0 ? 1u : 0u);
              umma_tf32_ss(d, da + o, dbl + o, idesc, 1u);
              umma_tf32_ss(d, da + o, db + o, idesc, 1u);
            } else {
              umma_tf32_ss(d, da + o, db + o, idesc, (kb | kk) != 0 ? 1u : 0u);
            }
          }
          umma_commit(smem_u32(&s_empty[s]));
        }
        umma_commit(smem_u32(&s_accf[b]));
      }
    }
  } else if (warp - PW < EPW) {
    // =============================== epilogue ===============================
    const int ew = warp - PW;           // 0..EPW-1
    const int q4 = ew & 3;              // TMEM lane quarter == warp % 4 (PW is a multiple of 4)
    const int chalf = ew >> 2;          // which half of the tile's columns this warp drains
    float* stg = reinterpret_cast<float*>(base_ptr + NSTAGE * STAGE) + ew * 32 * WSTG;
    uint32_t ti = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++ti) {
      int m0, n0, z;
      decode(t, m0, n0, z);
      float* __restrict__ C = g.C + (int64_t)z * g.strideC;
      const uint32_t b = ti & 1u;
      mbar_wait(smem_u32(&s_accf[b]), (ti >> 1) & 1u);
      tc_fence_after();
      const int row = m0 + q4 * 32 + lane;
      const float rbias = (g.bias_mode == 2 && row < g.M) ? g.bias[row] : 0.0f;
      const uint32_t tl = tmem + b * TCOLS_PER + ((uint32_t)(q4 * 32) << 16);
      constexpr int CH = BN / 32;  // 32-column chunks; chunk i goes to the warp pair member (i * 2 / CH)
      const int ci0 = (EPW == 8) ? chalf * ((CH + 1) / 2) : 0;
      const int ci1 = (EPW == 8 && chalf == 0) ? (CH + 1) / 2 : CH;
#pragma unroll 1
      for (int ci = ci0; ci < ci1; ++ci) {
        const int c0 = ci * 32;
        uint32_t v[32];
        tmem_ld32(tl + (uint32_t)c0, v);
        tmem_ld_wait();
        float f[32];
        // group max: ReLU (and the TF32 rounding) commute with the max, so they are applied to the 1 pooled value per lane
        // instead of the 32 accumulator values
        epi_chunk(f, v, rbias, g.bias_mode == 1 ? &s_bias[n0 + c0] : nullptr, g.group_max ? (int)ACT_NONE : g.act,
                  g.group_max ? 0 : g.round_out);
        if (SPLIT && g.C_lo != nullptr) {
          // the consumer is another 3xTF32 GEMM: store x as its hi and lo planes (two passes through the staging tile)
          float* __restrict__ Cl = g.C_lo + (int64_t)z * g.strideC;
#pragma unroll
          for (int pass = 0; pass < 2; ++pass) {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              float o[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float hi = tf32_round_fin(f[4 * q + e]);
                o[e] = pass == 0 ? hi : tf32_round_fin(f[4 * q + e] - hi);
              }
              *reinterpret_cast<float4*>(stg + lane * WSTG + q * 4) = make_float4(o[0], o[1], o[2], o[3]);
            }
            __syncwarp();
            float* __restrict__ dst = pass == 0 ? C : Cl;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              int r = 4 * j + (lane >> 3), q = lane & 7;
              float4 val = *reinterpret_cast<const float4*>(stg + r * WSTG + q * 4);
              int m = m0 + q4 * 32 + r;
              if (m < g.M) *reinterpret_cast<float4*>(dst + (int64_t)m * g.ldc + n0 + c0 + q * 4) = val;
            }
            __syncwarp();
          }
        } else if (g.group_max) {
          // max of the biased values as SIGNED integers: exact whenever the true maximum is >= 0, some negative value
          // otherwise -- which the ReLU maps to the same 0
          int res = 0;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int mx = __reduce_max_sync(0xffffffffu, __float_as_int(f[j]));
            if (lane == j) res = mx;
          }
          float r = fmaxf(__int_as_float(res), 0.0f);
          if (g.round_out) r = tf32_round_fin(r);
          const int gm = (m0 >> 5) + q4;
          if (gm < (g.M >> 5)) C[(int64_t)gm * g.ldc + n0 + c0 + lane] = r;
        } else {
#pragma unroll
          for (int q = 0; q < 8; ++q)
            *reinterpret_cast<float4*>(stg + lane * WSTG + q * 4) = make_float4(f[4 * q], f[4 * q + 1], f[4 * q + 2], f[4 * q + 3]);
          __syncwarp();
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            int r = 4 * j + (lane >> 3), q = lane & 7;
            float4 val = *reinterpret_cast<const float4*>(stg + r * WSTG + q * 4);
            int m = m0 + q4 * 32 + r;
            if (m < g.M) *reinterpret_cast<float4*>(C + (int64_t)m * g.ldc + n0 + c0 + q * 4) = val;
          }
          __syncwarp();
        }
      }
      tc_fence_before();
      mbar_arrive(smem_u32(&s_acce[b]));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 2 * TCOLS_PER);
}

}  // namespace

using EncodeFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeFn tma_encoder() {
  static EncodeFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeFn>(p);
  }
  return fn;
}
// [rows x K] fp32 row-major with leading dimension ld (floats) -> boxes of box_rows x 32 floats, SWIZZLE_128B, zero fill out of bounds
bool tma_map_2d(CUtensorMap* m, const float* ptr, int64_t rows, int K, int64_t ld, int box_rows) {
  EncodeFn enc = tma_encoder();
  if (!enc || (reinterpret_cast<uintptr_t>(ptr) & 15) || (ld & 3) || rows <= 0) return false;
  const cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  const cuuint32_t box[2] = {32, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

namespace {

template <int BN, bool SPLIT, bool ASYNC>
int launch_ws2(const GemmArgs& g, cudaStream_t st) {
  constexpr int nstage = ws_nstage(BN, SPLIT, ASYNC);
  constexpr int smem = nstage * (W_A_STAGE + BN * 128) * (SPLIT ? 2 : 1) + ws_epi_warps(BN, SPLIT) * 32 * WSTG * 4 + 1024;
  static_assert(smem <= 227 * 1024 - 8 * 1024, "shared memory budget");
  static PerDeviceOnce attr_done;
  if (smem_opt_in(attr_done, gemm_ws_kernel<BN, SPLIT, ASYNC>, smem) != cudaSuccess) return -1;
  const int sms = device_sm_count();
  const int batch = g.batch > 0 ? g.batch : 1;
  const int tiles_m = (g.M + WBM - 1) / WBM, tiles_n = g.N / BN;
  const int64_t total = (int64_t)tiles_m * tiles_n * batch;
  if (total > 0x7fffffff) return -1;
  const int grid = total < sms ? (int)total : sms;
  TmaMaps maps;
  int use_tma = 0;
  if (ASYNC && g_gemm_tma && batch == 1) {
    use_tma = tma_map_2d(&maps.a, g.A, g.M, g.K, g.lda, WBM) && tma_map_2d(&maps.w, g.W, g.N, g.K, g.ldw, BN) &&
              (!SPLIT || (tma_map_2d(&maps.a_lo, g.A_lo, g.M, g.K, g.lda, WBM) && tma_map_2d(&maps.w_lo, g.W_lo, g.N, g.K, g.ldw, BN)));
  }
  gemm_ws_kernel<BN, SPLIT, ASYNC><<<grid, WS_THREADS, smem, st>>>(g, tiles_m, tiles_n, (int)total, use_tma, maps);
  return 1;
}
template <int BN>
int launch_ws(const GemmArgs& g, cudaStream_t st) {
  if (g.precision >= 2) {
    if (g.A_lo && g.W_lo && g_gemm_async) return launch_ws2<BN, true, true>(g, st);
    return launch_ws2<BN, true, false>(g, st);
  }
  if (g.a_rounded && g.w_rounded && g_gemm_async) return launch_ws2<BN, false, true>(g, st);
  return launch_ws2<BN, false, false>(g, st);
}

}  // namespace

int g_gemm_tma = 1;  // 1: pre-rounded / pre-split operands of un-batched GEMMs are fed by TMA; 0: cp.async

int launch_gemm_ws(const GemmArgs& g, cudaStream_t st) {
  if (!gemm_tc_eligible(g) || g.N > 1024) return -1;
  switch (g.N) {
    case 32: return launch_ws<32>(g, st);
    case 64: return launch_ws<64>(g, st);
    case 128: return launch_ws<128>(g, st);
    case 192: return launch_ws<192>(g, st);
    default: return launch_ws<256>(g, st);
  }
}

}  // namespace lsdm
