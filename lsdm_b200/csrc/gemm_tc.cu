// tcgen05 (5th-gen tensor core) implementation of the generic linear-layer GEMM (gemm.cuh):
//   C[M,N] = act(A[M,K] . W[N,K]^T + bias),  TF32 operands read as fp32 from shared memory, fp32 accumulate in TMEM.
// One CTA = one 128 x BN output tile (UMMA M=128, N=BN, K=8 per instruction).  K is consumed in 32-float
// k-blocks (one 128-byte swizzle row per matrix row); A and W k-blocks are register-staged (rounded to TF32 with
// cvt.rna, optionally split hi/lo for 3xTF32) into a 2-stage ring in the canonical K-major SWIZZLE_128B layout, one
// elected thread issues the MMAs and tcgen05.commit releases each stage through an mbarrier.  Epilogue: tcgen05.ld (thread = row) -> bias/activation ->
// either a shared-memory transpose for fully coalesced 128-bit stores, or the PointNet++ group max
// (max over the 32 rows a warp owns == redux.sync on the non-negative float bit patterns).
#include "gemm.cuh"
#include "tc_ptx.cuh"

namespace lsdm {

namespace {

using namespace tc;

constexpr int TBM = 128;      // rows per CTA tile == TMEM lanes
constexpr int TBK = 32;       // floats per k-block (128 B)
constexpr int TNT = 128;      // threads
constexpr int A_STAGE = TBM * 128;  // bytes
constexpr int STG_STRIDE = 36;      // floats; epilogue transpose row stride (conflict-free float4 both ways)

template <int BN, bool SPLIT>
__global__ void __launch_bounds__(TNT) gemm_tc_kernel(GemmArgs g) {
  constexpr int W_STAGE = BN * 128;
  constexpr int STAGE = (A_STAGE + W_STAGE) * (SPLIT ? 2 : 1);  // [A_hi | W_hi | A_lo | W_lo]
  constexpr uint32_t TCOLS = BN <= 32 ? 32 : (BN <= 64 ? 64 : (BN <= 128 ? 128 : 256));
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t s_bar[3];  // mma_done[2], acc_done
  __shared__ uint32_t s_tmem;
  __shared__ float s_bias[BN];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int m0 = blockIdx.x * TBM, n0 = blockIdx.y * BN;
  const float* __restrict__ A = g.A + (int64_t)blockIdx.z * g.strideA;
  const float* __restrict__ W = g.W + (int64_t)blockIdx.z * g.strideW;
  float* __restrict__ C = g.C + (int64_t)blockIdx.z * g.strideC;

  // 1024-byte aligned operand ring (SWIZZLE_128B atoms must start on 1024 B)
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t bar_mma0 = smem_u32(&s_bar[0]), bar_mma1 = smem_u32(&s_bar[1]), bar_acc = smem_u32(&s_bar[2]);

  if (warp == 0) tmem_alloc(smem_u32(&s_tmem), TCOLS);
  if (tid == 0) {
    mbar_init(bar_mma0, 1);
    mbar_init(bar_mma1, 1);
    mbar_init(bar_acc, 1);
    fence_mbar_init();
  }
  if (g.bias_mode == 1)
    for (int i = tid; i < BN; i += TNT) s_bias[i] = g.bias[n0 + i];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s_tmem;

  const int nkb = g.K / TBK;
  // Both operands are staged through registers: ld.global (next k-block in flight while the current one is multiplied)
  // -> cvt.rna.tf32 (operands are ROUNDED, never truncated) -> swizzled st.shared.  SPLIT additionally stores the
  // residual lo = rna(x - hi) of both operands: D += A_lo.W_hi + A_hi.W_lo + A_hi.W_hi ("3xTF32", ~fp32 accuracy).
  constexpr int A_PER = TBM * 8 / TNT, W_PER = BN * 8 / TNT;
  float4 ra[A_PER], rw[W_PER];
  auto fetch = [&](int kb) {
    const int k0 = kb * TBK;
#pragma unroll
    for (int i = 0; i < A_PER; ++i) {
      int q = tid + i * TNT;
      int r = q >> 3, c = q & 7;
      int m = m0 + r;
      m = m < g.M ? m : g.M - 1;  // clamp: rows past M are never stored
      ra[i] = *reinterpret_cast<const float4*>(A + (int64_t)m * g.lda + k0 + c * 4);
    }
#pragma unroll
    for (int i = 0; i < W_PER; ++i) {
      int q = tid + i * TNT;
      int r = q >> 3, c = q & 7;
      rw[i] = *reinterpret_cast<const float4*>(W + (int64_t)(n0 + r) * g.ldw + k0 + c * 4);
    }
  };
  auto sub4 = [](float4 a, float4 b) { return make_float4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); };
  auto store = [&](int kb) {
    const uint32_t sA = base + (kb & 1) * STAGE, sW = sA + A_STAGE;
#pragma unroll
    for (int i = 0; i < A_PER; ++i) {
      int q = tid + i * TNT;
      uint32_t off = sw128_off(q >> 3, q & 7);
      float4 hi = rna_tf32(ra[i]);
      st_shared_v4(sA + off, hi);
      if (SPLIT) st_shared_v4(sA + (A_STAGE + W_STAGE) + off, rna_tf32(sub4(ra[i], hi)));
    }
#pragma unroll
    for (int i = 0; i < W_PER; ++i) {
      int q = tid + i * TNT;
      uint32_t off = sw128_off(q >> 3, q & 7);
      float4 hi = rna_tf32(rw[i]);
      st_shared_v4(sW + off, hi);
      if (SPLIT) st_shared_v4(sW + (A_STAGE + W_STAGE) + off, rna_tf32(sub4(rw[i], hi)));
    }
  };

  constexpr uint32_t idesc = umma_idesc_tf32(TBM, BN);
  fetch(0);
  for (int kb = 0; kb < nkb; ++kb) {
    // stage kb&1 was last read by the MMAs of iteration kb-2: wait for their commit before overwriting it
    if (kb >= 2) mbar_wait((kb & 1) ? bar_mma1 : bar_mma0, (uint32_t)(((kb - 2) >> 1) & 1));
    store(kb);
    if (kb + 1 < nkb) fetch(kb + 1);
    fence_proxy_async();  // generic-proxy smem writes -> visible to the tensor core (async proxy)
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      const uint32_t sA = base + (kb & 1) * STAGE, sW = sA + A_STAGE;
      const uint64_t da = umma_desc_sw128(sA), db = umma_desc_sw128(sW);
#pragma unroll
      for (int kk = 0; kk < TBK / 8; ++kk) {
        const uint64_t o = (uint64_t)(kk * 2);
        if (SPLIT) {
          const uint64_t dal = umma_desc_sw128(sA + A_STAGE + W_STAGE), dbl = umma_desc_sw128(sW + A_STAGE + W_STAGE);
          umma_tf32_ss(tmem, dal + o, db + o, idesc, (kb | kk) != 0 ? 1u : 0u);
          umma_tf32_ss(tmem, da + o, dbl + o, idesc, 1u);
          umma_tf32_ss(tmem, da + o, db + o, idesc, 1u);
        } else {
          umma_tf32_ss(tmem, da + o, db + o, idesc, (kb | kk) != 0 ? 1u : 0u);
        }
      }
      umma_commit((kb & 1) ? bar_mma1 : bar_mma0);
      if (kb == nkb - 1) umma_commit(bar_acc);
    }
  }
  mbar_wait(bar_acc, 0);
  tc_fence_after();

  // ---- epilogue: warp w owns TMEM lanes / tile rows [32w, 32w+32) ----
  float* stg = reinterpret_cast<float*>(base_ptr) + warp * 32 * STG_STRIDE;  // aliases the (now idle) operand ring
  const int row = m0 + warp * 32 + lane;
  const float rbias = (g.bias_mode == 2 && row < g.M) ? g.bias[row] : 0.0f;
#pragma unroll 1
  for (int c0 = 0; c0 < BN; c0 += 32) {
    uint32_t v[32];
    tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
    tmem_ld_wait();
    float f[32];
    epi_chunk(f, v, rbias, g.bias_mode == 1 ? &s_bias[c0] : nullptr, g.act, g.round_out);
    if (g.group_max) {
      // rows of this warp = one 32-sample group; values are post-ReLU (>= 0), so uint order == float order
      uint32_t res = 0;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        uint32_t mx = __reduce_max_sync(0xffffffffu, __float_as_uint(f[j]));
        if (lane == j) res = mx;
      }
      const int gm = (m0 >> 5) + warp;
      if (gm < (g.M >> 5)) C[(int64_t)gm * g.ldc + n0 + c0 + lane] = __uint_as_float(res);
    } else {
#pragma unroll
      for (int q = 0; q < 8; ++q)
        *reinterpret_cast<float4*>(stg + lane * STG_STRIDE + q * 4) = make_float4(f[4 * q], f[4 * q + 1], f[4 * q + 2], f[4 * q + 3]);
      __syncwarp();
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        int r = 4 * j + (lane >> 3), q = lane & 7;
        float4 val = *reinterpret_cast<const float4*>(stg + r * STG_STRIDE + q * 4);
        int m = m0 + warp * 32 + r;
        if (m < g.M) *reinterpret_cast<float4*>(C + (int64_t)m * g.ldc + n0 + c0 + q * 4) = val;
      }
      __syncwarp();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, TCOLS);
}

template <int BN, bool SPLIT>
int launch_bn2(const GemmArgs& g, cudaStream_t st) {
  constexpr int smem = 2 * (A_STAGE + BN * 128) * (SPLIT ? 2 : 1) + 1024;
  static PerDeviceOnce attr_done;
  if (smem_opt_in(attr_done, gemm_tc_kernel<BN, SPLIT>, smem) != cudaSuccess) return -1;
  int batch = g.batch > 0 ? g.batch : 1;
  dim3 grid((g.M + TBM - 1) / TBM, g.N / BN, batch);
  gemm_tc_kernel<BN, SPLIT><<<grid, TNT, smem, st>>>(g);
  return 1;
}
template <int BN>
int launch_bn(const GemmArgs& g, cudaStream_t st) {
  return g.precision >= 2 ? launch_bn2<BN, true>(g, st) : launch_bn2<BN, false>(g, st);
}

}  // namespace

bool gemm_tc_eligible(const GemmArgs& g) {
  if (g.M <= 0 || g.K < 32 || (g.K % 32) != 0 || (g.lda & 3) || (g.ldw & 3) || (g.ldc & 3)) return false;
  if ((reinterpret_cast<uintptr_t>(g.A) | reinterpret_cast<uintptr_t>(g.W) | reinterpret_cast<uintptr_t>(g.C)) & 15) return false;
  if ((g.strideA & 3) || (g.strideW & 3) || (g.strideC & 3)) return false;
  if (g.group_max && (g.act != ACT_RELU || (g.M % 32) != 0)) return false;
  const int N = g.N;
  return N == 32 || N == 64 || N == 128 || N == 192 || (N >= 256 && N % 256 == 0);
}

int launch_gemm_tc(const GemmArgs& g, cudaStream_t st) {
  if (!gemm_tc_eligible(g)) return -1;
  switch (g.N) {
    case 32: return launch_bn<32>(g, st);
    case 64: return launch_bn<64>(g, st);
    case 128: return launch_bn<128>(g, st);
    case 192: return launch_bn<192>(g, st);
    default: return launch_bn<256>(g, st);
  }
}

}  // namespace lsdm
