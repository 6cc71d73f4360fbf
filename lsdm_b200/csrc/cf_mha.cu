// ContactFormer temporal multi-head attention layer (reference contact_former/transformer.py:44-103
// `MultiHeadAttention.forward`, :153-177 `PositionwiseFeedForward.forward`, eval mode: Dropout off).
//
//   x[bs, S, V, d_in]  (S = frames of the segment, V = mesh vertices, d_in = 64)
//   q|k|v = x W_{q,k,v}^T + b   -> per (head, vertex, sample) a sequence of S rows of 64
//   a     = softmax(q k^T / sqrt(d_k) [masked_fill(mask == 0, -inf)]) v        (attention ALONG TIME, per vertex)
//   out   = LayerNorm(fc(concat_heads(a)) + x)
//
// Kernels: the three projections and fc are the tcgen05 GEMM of gemm_tc.cu (TF32 / 3xTF32, fp32 accumulate in TMEM);
// softmax(QK^T)V is one fused kernel per (head, vertex, sample): K and V of the sequence staged once in shared memory,
// one warp per query row, scores -> softmax -> P.V without touching global memory (warp-level reductions); the residual
// add + LayerNorm is a warp-per-row kernel (ln.cuh).  The S x S score matrix is never materialised in HBM (the reference
// materialises n_head*V*bs of them).
#include <string>

#include "../../include/lsdm_b200.h"
#include "kernels.cuh"
#include "ln.cuh"
#include "tc_ptx.cuh"

using namespace lsdm;

namespace {

constexpr int CF_D = 64;      // d_in == d_k == d_v (contact_former.py:271-275,318: all three are `channels` = 64)
constexpr int CF_MAX_S = 256;  // frames per segment staged in shared memory
constexpr int CF_WARPS = 8;

// frame-major row r = (b*S + s)*V + v  ->  vertex-major row (b*V + v)*S + s
__device__ __forceinline__ int64_t vertex_major_row(int64_t r, int S, int V) {
  const int64_t bs_ = r / V;
  const int v = (int)(r - bs_ * V);
  const int64_t b = bs_ / S;
  const int s = (int)(bs_ - b * S);
  return (b * V + v) * S + s;
}

// xt[(b*V + v)*S + s] = x[(b*S + s)*V + v]: the time axis of one vertex becomes contiguous (what the reference's
// permute(3, 2, 0, 1, 4).contiguous() does to q, k and v, transformer.py:78-80 -- done once, on the 64-wide input instead)
__global__ void __launch_bounds__(256) cf_to_vertex_major_kernel(const float* __restrict__ x, int64_t rows, int S, int V, float* __restrict__ xt) {
  const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= rows) return;
  const float2 val = *reinterpret_cast<const float2*>(x + r * CF_D + lane * 2);
  *reinterpret_cast<float2*>(xt + vertex_major_row(r, S, V) * CF_D + lane * 2) = val;
}

// out[r] = LN(x[r] + y[r']); x is the layer input (left untouched); r' = r, or the vertex-major row of r when S > 0
template <int PER>
__global__ void __launch_bounds__(256) residual_ln_kernel(const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ g,
                                                          const float* __restrict__ b, int64_t rows, float* __restrict__ out, int S = 0, int V = 0) {
  constexpr int WIDTH = 32 * PER;
  const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= rows) return;
  const int64_t ry = S > 0 ? vertex_major_row(r, S, V) : r;
  float v[PER];
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    const int c = i * 32 + lane;
    v[i] = x[r * WIDTH + c] + y[ry * WIDTH + c];
  }
  layer_norm_row<PER>(v, g, b, lane, out + r * WIDTH);
}

// grid (H, V, bs).  Rows are VERTEX-MAJOR, r = (b*V + v)*S + s (cf_to_vertex_major_kernel), so that the S rows of one sequence
// are adjacent.  qkv[r][3*H*64]: q | k | v column blocks, head h = columns h*64.. of each.  att[r][H*64].  mask (nullable): uint8 [bs, S, S], nonzero = keep (transformer.py:89 fills -inf where mask == 0);
// all_masked: the reference's `mask.sum() == 0` branch (transformer.py:91-92): attention weights forced to 0.
__global__ void __launch_bounds__(CF_WARPS * 32) cf_attn_kernel(const float* __restrict__ qkv, const uint8_t* __restrict__ mask, int all_masked, int S, int V,
                                                                int H, float inv_temp, float* __restrict__ att) {
  extern __shared__ float sm[];
  float* sk = sm;                                   // [S][65]
  float* sv = sk + ((S * (CF_D + 1) + 3) & ~3);      // [S][64], 16-byte aligned
  float* sq = sv + S * CF_D;              // [warps][64]
  float* sp = sq + CF_WARPS * CF_D;       // [warps][S]
  const int h = blockIdx.x, v = blockIdx.y, bi = blockIdx.z, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t ld = 3 * (int64_t)H * CF_D;
  const int64_t row0 = ((int64_t)bi * V + v) * S, rstride = 1;  // rows are vertex-major (b, v, s): a sequence is contiguous
  const float* base = qkv + h * CF_D;
  for (int e = tid; e < S * (CF_D / 4); e += blockDim.x) {
    const int j = e / (CF_D / 4), d4 = (e % (CF_D / 4)) * 4;
    const float* src = base + (row0 + (int64_t)j * rstride) * ld;
    const float4 kk = *reinterpret_cast<const float4*>(src + H * CF_D + d4);
    const float4 vv = *reinterpret_cast<const float4*>(src + 2 * H * CF_D + d4);
    float* dk = sk + j * (CF_D + 1) + d4;
    dk[0] = kk.x; dk[1] = kk.y; dk[2] = kk.z; dk[3] = kk.w;
    *reinterpret_cast<float4*>(sv + j * CF_D + d4) = vv;
  }
  __syncthreads();
  float* q = sq + warp * CF_D;
  float* p = sp + warp * S;
  const uint8_t* mrow0 = mask ? mask + (int64_t)bi * S * S : nullptr;
  for (int i = warp; i < S; i += CF_WARPS) {
    const float* qsrc = base + (row0 + (int64_t)i * rstride) * ld;
    q[lane] = qsrc[lane];
    q[lane + 32] = qsrc[lane + 32];
    __syncwarp();
    float s[CF_MAX_S / 32], mx = -INFINITY;
#pragma unroll
    for (int c = 0; c < CF_MAX_S / 32; ++c) {
      const int j = c * 32 + lane;
      s[c] = -INFINITY;
      if (j < S) {
        float a = 0.f;
        const float* kr = sk + j * (CF_D + 1);
#pragma unroll 16
        for (int d = 0; d < CF_D; ++d) a = fmaf(q[d], kr[d], a);
        a *= inv_temp;  // attn / temperature (transformer.py:85)
        if (mrow0 && mrow0[(int64_t)i * S + j] == 0) a = -INFINITY;
        s[c] = a;
      }
      mx = fmaxf(mx, s[c]);
    }
    mx = warp_max(mx);
    float sum = 0.f;
#pragma unroll
    for (int c = 0; c < CF_MAX_S / 32; ++c) {
      // a fully masked row gives exp(-inf - -inf) = NaN, as torch's softmax over a row of -inf does
      s[c] = (c * 32 + lane) < S ? expf(s[c] - mx) : 0.f;
      sum += s[c];
    }
    const float inv = all_masked ? 0.f : 1.0f / warp_sum(sum);
#pragma unroll
    for (int c = 0; c < CF_MAX_S / 32; ++c)
      if (c * 32 + lane < S) p[c * 32 + lane] = all_masked ? 0.f : s[c] * inv;
    __syncwarp();
    float o0 = 0.f, o1 = 0.f;
    for (int j = 0; j < S; ++j) {
      const float pj = p[j];
      o0 = fmaf(pj, sv[j * CF_D + lane], o0);
      o1 = fmaf(pj, sv[j * CF_D + lane + 32], o1);
    }
    float* dst = att + (row0 + (int64_t)i * rstride) * ((int64_t)H * CF_D) + h * CF_D;
    dst[lane] = o0;
    dst[lane + 32] = o1;
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Tensor-core form of the same attention (seg_len a multiple of 32): one CTA per (head, vertex, sample), 256 threads.
//   K [S,64] and V^T [64,S] of the sequence are staged once in shared memory as TF32 K-major SWIZZLE_128B operands;
//   per 128-query tile:  D1[128,S] = (Q/8) K^T  (tcgen05, SS)  ->  softmax in place in TMEM (thread = query row, two warps per
//   TMEM lane quarter split the key chunks; row max / row sum exchanged through shared memory)  ->  O[128,64] = P V
//   (tcgen05, A = P read from TMEM)  ->  O / rowsum written to att.  The S x S matrix lives only in tensor memory.
// TMEM: columns [0,256) scores / probabilities, [256,320) output accumulator (512 allocated).
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256, 1) cf_attn_tc_kernel(const float* __restrict__ qkv, const uint8_t* __restrict__ mask, int all_masked, int S, int V,
                                                            int H, float* __restrict__ att) {
  using namespace tc;
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t s_bar;
  __shared__ uint32_t s_tmem;
  __shared__ float s_max[2][128], s_sum[2][128];
  const int h = blockIdx.x, v = blockIdx.y, bi = blockIdx.z, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int wq = warp & 3, hf = warp >> 2, row = wq * 32 + lane;
  const int64_t ld = 3 * (int64_t)H * CF_D;
  const int64_t row0 = ((int64_t)bi * V + v) * S, rstride = 1;  // rows are vertex-major (b, v, s)
  const float* base = qkv + h * CF_D;
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sK = sbase;                          // 2 k-blocks x [S rows x 128 B]
  const uint32_t sVt = sK + 2u * (uint32_t)S * 128u;  // S/32 k-blocks x [64 rows x 128 B]
  const uint32_t sQ = sVt + (uint32_t)S * 256u;       // 2 k-blocks x [128 rows x 128 B]
  const uint32_t bar = smem_u32(&s_bar);
  constexpr uint32_t COL_S = 0, COL_O = 256, TCOLS = 512;

  if (warp == 0) tmem_alloc(smem_u32(&s_tmem), TCOLS);
  if (tid == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  // Q tile of the first 128 queries: loads issued first, consumed after the K / V staging (latency hidden behind it)
  float4 qreg[8];
  auto q_fetch = [&](int m0) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int e = tid + u * 256, r = e >> 4, c16 = e & 15, i = m0 + r;
      qreg[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (i < S) qreg[u] = *reinterpret_cast<const float4*>(base + (row0 + (int64_t)i * rstride) * ld + c16 * 4);
    }
  };
  auto q_store = [&]() {  // scaled by 1/sqrt(64) = 0.125 (exact), rounded to TF32, rows past S are zero
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int e = tid + u * 256, r = e >> 4, c16 = e & 15;
      const float4 q = qreg[u];
      st_shared_v4(sQ + (uint32_t)(c16 >> 3) * (128u * 128u) + sw128_off(r, c16 & 7),
                   rna_tf32(make_float4(q.x * 0.125f, q.y * 0.125f, q.z * 0.125f, q.w * 0.125f)));
    }
  };
  q_fetch(0);
  // ---- stage K (row-major rows -> K-major operand) and V (transposed: row d of V^T holds the S keys); four iterations'
  //      loads are issued before their stores so that eight 128-bit loads per thread are in flight ----
  for (int e0 = tid; e0 < S * 16; e0 += 256 * 4) {
    float4 kk[4], vv[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int e = e0 + u * 256;
      if (e < S * 16) {
        const float* src = base + (row0 + (int64_t)(e >> 4) * rstride) * ld + (e & 15) * 4;
        kk[u] = *reinterpret_cast<const float4*>(src + H * CF_D);
        vv[u] = *reinterpret_cast<const float4*>(src + 2 * H * CF_D);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int e = e0 + u * 256;
      if (e < S * 16) {
        const int j = e >> 4, c16 = e & 15;  // key j, 16-byte chunk c16 of its 64 floats
        st_shared_v4(sK + (uint32_t)(c16 >> 3) * (uint32_t)S * 128u + sw128_off(j, c16 & 7), rna_tf32(kk[u]));
        const uint32_t vb = sVt + (uint32_t)(j >> 5) * (64u * 128u) + (uint32_t)(j & 3) * 4u;
        const int cj = (j & 31) >> 2;
        const float vals[4] = {vv[u].x, vv[u].y, vv[u].z, vv[u].w};
#pragma unroll
        for (int t = 0; t < 4; ++t)
          asm volatile("st.shared.f32 [%0], %1;" ::"r"(vb + sw128_off(c16 * 4 + t, cj)), "f"(rna_tf32(vals[t])) : "memory");
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s_tmem;
  const uint32_t tlane = tmem + ((uint32_t)(wq * 32) << 16);
  const uint32_t idesc1 = umma_idesc_tf32(128, S), idesc2 = umma_idesc_tf32(128, 64);
  const int nchunk = S >> 5;
  uint32_t phase = 0;
  const uint8_t* mb = mask ? mask + (int64_t)bi * S * S : nullptr;

  for (int m0 = 0; m0 < S; m0 += 128) {
    q_store();
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
#pragma unroll
      for (int kb = 0; kb < 2; ++kb) {
        const uint64_t da = umma_desc_sw128(sQ + kb * (128 * 128)), db = umma_desc_sw128(sK + (uint32_t)kb * (uint32_t)S * 128u);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) umma_tf32_ss(tmem + COL_S, da + (uint64_t)(kk * 2), db + (uint64_t)(kk * 2), idesc1, (kb | kk) != 0 ? 1u : 0u);
      }
      umma_commit(bar);
    }
    if (m0 + 128 < S) q_fetch(m0 + 128);  // next query tile: in flight during this tile's softmax / P.V / epilogue
    mbar_wait(bar, phase);
    phase ^= 1;
    tc_fence_after();
    // ---- softmax of row (m0 + row), this warp's key chunks: hf, hf + 2, ... ----
    const int i = m0 + row;
    const uint8_t* mrow = (mb && i < S) ? mb + (int64_t)i * S : nullptr;
    auto load_chunk = [&](int c, uint32_t (&sv)[32]) {
      tmem_ld32(tlane + COL_S + c * 32, sv);
      tmem_ld_wait();
      if (mrow) {
        const uint4 ma = *reinterpret_cast<const uint4*>(mrow + c * 32), mc = *reinterpret_cast<const uint4*>(mrow + c * 32 + 16);
        const uint32_t mw[8] = {ma.x, ma.y, ma.z, ma.w, mc.x, mc.y, mc.z, mc.w};
#pragma unroll
        for (int k = 0; k < 32; ++k)
          if (((mw[k >> 2] >> ((k & 3) * 8)) & 0xffu) == 0u) sv[k] = 0xff800000u;  // -inf
      }
    };
    float mx = -INFINITY;
    for (int c = hf; c < nchunk; c += 2) {
      uint32_t sv[32];
      load_chunk(c, sv);
#pragma unroll
      for (int k = 0; k < 32; ++k) mx = fmaxf(mx, __uint_as_float(sv[k]));
    }
    s_max[hf][row] = mx;
    __syncthreads();
    mx = fmaxf(s_max[0][row], s_max[1][row]);
    constexpr float LOG2E = 1.4426950408889634f;
    const float mxl = mx * LOG2E;
    float sum = 0.f;
    for (int c = hf; c < nchunk; c += 2) {
      uint32_t sv[32];
      load_chunk(c, sv);
#pragma unroll
      for (int k = 0; k < 32; ++k) {
        // exp(s - m) as exp2(s log2e - m log2e); a fully masked row gives -inf + inf = NaN, as torch's softmax does
        const float pexp = all_masked ? 0.f : exp2f(fmaf(__uint_as_float(sv[k]), LOG2E, -mxl));
        sum += pexp;
        sv[k] = rna_tf32_mma(pexp);  // only the MMA reads it (ignores the low 13 bits): round to nearest = one integer add
      }
      tmem_st32(tlane + COL_S + c * 32, sv);
    }
    tmem_st_wait();
    s_sum[hf][row] = sum;
    tc_fence_before();
    __syncthreads();
    // ---- O = P V ----
    if (tid == 0) {
      tc_fence_after();
      for (int kb = 0; kb < nchunk; ++kb) {
        const uint64_t db = umma_desc_sw128(sVt + (uint32_t)kb * (64u * 128u));
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
          umma_tf32_ts(tmem + COL_O, tmem + COL_S + kb * 32 + kk * 8, db + (uint64_t)(kk * 2), idesc2, (kb | kk) != 0 ? 1u : 0u);
      }
      umma_commit(bar);
    }
    mbar_wait(bar, phase);
    phase ^= 1;
    tc_fence_after();
    {
      uint32_t ov[32];
      tmem_ld32(tlane + COL_O + hf * 32, ov);
      tmem_ld_wait();
      const float inv = all_masked ? 0.f : 1.0f / (s_sum[0][row] + s_sum[1][row]);
      if (i < S) {
        float* dst = att + (row0 + (int64_t)i * rstride) * ((int64_t)H * CF_D) + h * CF_D + hf * 32;
#pragma unroll
        for (int k = 0; k < 32; k += 4)
          *reinterpret_cast<float4*>(dst + k) = make_float4(__uint_as_float(ov[k]) * inv, __uint_as_float(ov[k + 1]) * inv,
                                                            __uint_as_float(ov[k + 2]) * inv, __uint_as_float(ov[k + 3]) * inv);
      }
    }
    tc_fence_before();
    __syncthreads();  // TMEM scores / sQ / s_max / s_sum are reused by the next query tile
  }
  if (warp == 0) tmem_dealloc(tmem, TCOLS);
}

int cf_gemm(const float* A, int64_t lda, const float* W, int64_t ldw, const float* bias, float* C, int64_t ldc, int64_t M, int N, int K, int act,
            int precision, cudaStream_t st) {
  GemmArgs g{};
  g.A = A; g.lda = lda;
  g.W = W; g.ldw = ldw;
  g.C = C; g.ldc = ldc;
  g.bias = bias; g.bias_mode = bias ? 1 : 0;
  g.M = (int)M; g.N = N; g.K = K; g.batch = 1;
  g.act = act; g.precision = precision;
  return (precision >= 1 && gemm_tc_eligible(g)) ? launch_gemm_tc(g, st) : launch_gemm_simt(g, st);
}

struct CfWs {
  float *xt, *qkv, *att, *y;
  size_t bytes;
};
CfWs cf_carve(void* base, int64_t rows, int H) {
  char* p = static_cast<char*>(base);
  size_t off = 0;
  auto take = [&](size_t n) {
    off = (off + 255) & ~size_t(255);
    float* r = p ? reinterpret_cast<float*>(p + off) : nullptr;
    off += n * sizeof(float);
    return r;
  };
  CfWs w{};
  w.xt = take((size_t)rows * CF_D);
  w.qkv = take((size_t)rows * 3 * H * CF_D);
  w.att = take((size_t)rows * H * CF_D);
  w.y = take((size_t)rows * CF_D);
  w.bytes = (off + 255) & ~size_t(255);
  return w;
}

}  // namespace

int g_cf_attn_tc = 1;  // 1: tcgen05 attention when precision == 1 (TF32) and seg_len % 32 == 0; 0: CUDA-core fp32 kernel always

extern "C" {

LSDM_API int lsdm_cf_set_option(const char* name, int32_t value) {
  if (name && std::string(name) == "attn_tc" && (value == 0 || value == 1)) {
    g_cf_attn_tc = value;
    return LSDM_OK;
  }
  return set_error(LSDM_EINVAL, "lsdm_cf_set_option: unknown option");
}

LSDM_API size_t lsdm_cf_workspace_bytes(int32_t bs, int32_t seg_len, int32_t n_verts, int32_t n_head) {
  if (bs <= 0 || seg_len <= 0 || n_verts <= 0 || n_head <= 0) return 0;
  return cf_carve(nullptr, (int64_t)bs * seg_len * n_verts, n_head).bytes;
}

LSDM_API int lsdm_cf_mha_forward(const lsdm_cf_mha_weights* w, const float* x, const uint8_t* mask, int32_t all_masked, int32_t bs, int32_t seg_len,
                                 int32_t n_verts, int32_t n_head, int32_t precision, void* workspace, size_t workspace_bytes, float* out,
                                 void* stream) {
  if (!w || !x || !out || !workspace || bs <= 0 || n_verts <= 0 || n_head <= 0) return set_error(LSDM_EINVAL, "bad argument");
  if (!w->w_q || !w->w_k || !w->w_v || !w->b_q || !w->b_k || !w->b_v || !w->fc_w || !w->fc_b || !w->ln_w || !w->ln_b)
    return set_error(LSDM_EINVAL, "lsdm_cf_mha_forward: null weight pointer");
  if (seg_len < 1 || seg_len > CF_MAX_S) return set_error(LSDM_EINVAL, "lsdm_cf_mha_forward: 1 <= seg_len <= 256");
  if (precision < 0 || precision > 2) return set_error(LSDM_EINVAL, "precision must be 0 (fp32), 1 (tf32) or 2 (3xtf32)");
  if (((uintptr_t)workspace & 255) != 0) return set_error(LSDM_EINVAL, "workspace must be 256-byte aligned");
  const int64_t rows = (int64_t)bs * seg_len * n_verts;
  if (rows > 0x7fffffffll) return set_error(LSDM_EINVAL, "too many rows");
  const int H = n_head, HD = H * CF_D;
  CfWs ws = cf_carve(workspace, rows, H);
  if (ws.bytes > workspace_bytes) return set_error(LSDM_ENOMEM, "workspace too small (lsdm_cf_workspace_bytes)");
  cudaStream_t st = (cudaStream_t)stream;
  cf_to_vertex_major_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(x, rows, seg_len, n_verts, ws.xt);
  // q | k | v projections into one [rows, 3*H*64] buffer (vertex-major rows)
  const float* Ws[3] = {w->w_q, w->w_k, w->w_v};
  const float* Bs[3] = {w->b_q, w->b_k, w->b_v};
  for (int i = 0; i < 3; ++i)
    if (cf_gemm(ws.xt, CF_D, Ws[i], CF_D, Bs[i], ws.qkv + (int64_t)i * HD, 3 * (int64_t)HD, rows, HD, CF_D, ACT_NONE, precision, st) < 0)
      return set_error(LSDM_EINVAL, "lsdm_cf_mha_forward: projection GEMM (n_head * 64 must be 64, 128, 192 or a multiple of 256)");
  if (n_verts > 65535 || bs > 65535) return set_error(LSDM_EINVAL, "n_verts and bs must be <= 65535");
  if ((seg_len & 31) == 0 && g_cf_attn_tc && precision == 1) {  // single-pass TF32 operands: only where the caller asked for TF32
    // tensor-core attention: K + V^T + one Q tile as swizzled TF32 operands (+1 KB alignment slack)
    const size_t smem = (size_t)seg_len * 512 + 32 * 1024 + 1024;
    static PerDeviceOnce attr_tc;
  if (smem_opt_in(attr_tc, cf_attn_tc_kernel, 256 * 512 + 33 * 1024) != cudaSuccess) return set_error(LSDM_ECUDA, "cudaFuncSetAttribute(cf_attn_tc_kernel)");
    cf_attn_tc_kernel<<<dim3(H, n_verts, bs), 256, smem, st>>>(ws.qkv, mask, all_masked, seg_len, n_verts, H, ws.att);
  } else {
    const size_t smem = sizeof(float) * ((((size_t)seg_len * (CF_D + 1) + 3) & ~size_t(3)) + (size_t)seg_len * CF_D + CF_WARPS * CF_D + (size_t)CF_WARPS * seg_len);
    static PerDeviceOnce attr_done;
  if (smem_opt_in(attr_done, cf_attn_kernel, 160 * 1024) != cudaSuccess) return set_error(LSDM_ECUDA, "cudaFuncSetAttribute(cf_attn_kernel)");
    cf_attn_kernel<<<dim3(H, n_verts, bs), CF_WARPS * 32, smem, st>>>(ws.qkv, mask, all_masked, seg_len, n_verts, H, 1.0f / sqrtf((float)CF_D), ws.att);
  }
  if (cf_gemm(ws.att, HD, w->fc_w, HD, w->fc_b, ws.y, CF_D, rows, CF_D, HD, ACT_NONE, precision, st) < 0)
    return set_error(LSDM_EINVAL, "lsdm_cf_mha_forward: fc GEMM");
  residual_ln_kernel<2><<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(x, ws.y, w->ln_w, w->ln_b, rows, out, seg_len, n_verts);
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) return set_error(LSDM_ECUDA, std::string("lsdm_cf_mha_forward: ") + cudaGetErrorString(e));
  return LSDM_OK;
}

LSDM_API int lsdm_cf_ffn_forward(const lsdm_cf_ffn_weights* w, const float* x, int64_t rows, int32_t precision, void* workspace, size_t workspace_bytes,
                                 float* out, void* stream) {
  if (!w || !x || !out || !workspace || rows <= 0 || rows > 0x7fffffffll) return set_error(LSDM_EINVAL, "bad argument");
  if (!w->w1 || !w->b1 || !w->w2 || !w->b2 || !w->ln_w || !w->ln_b) return set_error(LSDM_EINVAL, "lsdm_cf_ffn_forward: null weight pointer");
  if (w->d_hid != 64 && w->d_hid != 128 && w->d_hid != 192 && (w->d_hid % 256) != 0)
    return set_error(LSDM_EINVAL, "lsdm_cf_ffn_forward: d_hid must be 64, 128, 192 or a multiple of 256");
  if (((uintptr_t)workspace & 255) != 0) return set_error(LSDM_EINVAL, "workspace must be 256-byte aligned");
  const size_t need = ((size_t)rows * (w->d_hid + CF_D) * sizeof(float) + 511) & ~size_t(255);
  if (need > workspace_bytes) return set_error(LSDM_ENOMEM, "workspace too small (rows * (d_hid + 64) floats + 512)");
  cudaStream_t st = (cudaStream_t)stream;
  float* hbuf = static_cast<float*>(workspace);
  float* y = hbuf + (((size_t)rows * w->d_hid + 63) & ~size_t(63));
  if (cf_gemm(x, CF_D, w->w1, CF_D, w->b1, hbuf, w->d_hid, rows, w->d_hid, CF_D, ACT_RELU, precision, st) < 0 ||
      cf_gemm(hbuf, w->d_hid, w->w2, w->d_hid, w->b2, y, CF_D, rows, CF_D, w->d_hid, ACT_NONE, precision, st) < 0)
    return set_error(LSDM_EINVAL, "lsdm_cf_ffn_forward: GEMM");
  residual_ln_kernel<2><<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(x, y, w->ln_w, w->ln_b, rows, out);
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) return set_error(LSDM_ECUDA, std::string("lsdm_cf_ffn_forward: ") + cudaGetErrorString(e));
  return LSDM_OK;
}

}  // extern "C"
