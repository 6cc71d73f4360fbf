// Condition-encoder kernels: the small per-sample MLPs of SceneDiffusionModel.forward
// (reference model/sdm.py:141-161,180-188), the timestep embedding (model/diffusion_utils.py:20-21)
// and the POSA human decoder (posa/posa_models.py:152-160,181-187,320-326).
#include "kernels.cuh"

namespace lsdm {

namespace {

// y[n] = act(W[n,:] . x + b[n]) for n < N; one warp per output, lanes stride K (coalesced weight rows).
template <int ACT>
__device__ __forceinline__ void cta_linear(const float* __restrict__ W, const float* __restrict__ b, const float* x,
                                           float* y, int N, int K) {
  // four outputs per warp iteration: their weight-row loads are issued together, so one L2 round trip covers four rows
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int n0 = warp * 4; n0 < N; n0 += nw * 4) {
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int k = lane; k < K; k += 32) {
      const float xv = x[k];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int n = n0 + j < N ? n0 + j : N - 1;
        acc[j] = fmaf(W[(int64_t)n * K + k], xv, acc[j]);
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float a = warp_sum(acc[j]);
      if (lane == 0 && n0 + j < N) y[n0 + j] = apply_act<ACT>(a + (b ? b[n0 + j] : 0.f));
    }
  }
  __syncthreads();
}

__global__ void __launch_bounds__(256) cond_kernel(CondWeights w, const float* __restrict__ text,
                                                   const float* __restrict__ cats, const float* __restrict__ mask_global,
                                                   int Bg, int b_off, int n_cats, float* __restrict__ enc_out,
                                                   float* __restrict__ out_cat, float* __restrict__ attn_w,
                                                   float* __restrict__ tr_out, float* __restrict__ qq_out) {
  __shared__ float s_in[CLIP];
  __shared__ float s_a[256], s_b[256];
  __shared__ float s_enc[LAT];
  __shared__ float s_ec[NOBJ][CATEMB];
  __shared__ float s_q[LAT];
  __shared__ float s_k[NOBJ][LAT];
  __shared__ float s_z[LAT + CATEMB];
  __shared__ float s_t1[LAT];
  __shared__ float s_tr[NOBJ][TRANS];
  __shared__ float s_logit[NHEAD][NOBJ];
  const int b = blockIdx.x, tid = threadIdx.x;

  for (int i = tid; i < CLIP; i += blockDim.x) s_in[i] = text[(int64_t)b * CLIP + i];
  __syncthreads();
  // embed_text: 512 -> 256 -> 256 -> 128, GELU after each (sdm.py:52-59)
  cta_linear<ACT_GELU>(w.et0_w, w.et0_b, s_in, s_a, 256, CLIP);
  cta_linear<ACT_GELU>(w.et2_w, w.et2_b, s_a, s_b, 256, 256);
  cta_linear<ACT_GELU>(w.et4_w, w.et4_b, s_b, s_enc, LAT, 256);
  for (int i = tid; i < LAT; i += blockDim.x) enc_out[(int64_t)b * LAT + i] = s_enc[i];
  // predict_cat: 128 -> 64 -> 32 -> C, GELU after each, softmax (sdm.py:68-76)
  cta_linear<ACT_GELU>(w.pc0_w, w.pc0_b, s_enc, s_a, 64, LAT);
  cta_linear<ACT_GELU>(w.pc2_w, w.pc2_b, s_a, s_b, 32, 64);
  cta_linear<ACT_GELU>(w.pc4_w, w.pc4_b, s_b, s_a, n_cats, 32);
  if (tid < 32) {
    float v = tid < n_cats ? s_a[tid] : -INFINITY;
    float mx = warp_max(v);
    float e = tid < n_cats ? expf(v - mx) : 0.f;
    float sum = warp_sum(e);
    if (tid < n_cats) out_cat[(int64_t)b * n_cats + tid] = e / sum;
  }
  // embed_cat: C -> 32 GELU per object (sdm.py:62-65,161)
  for (int i = tid; i < NOBJ * CATEMB; i += blockDim.x) {
    int o = i / CATEMB, j = i % CATEMB;
    const float* c = cats + ((int64_t)b * NOBJ + o) * n_cats;
    float acc = 0.f;
    for (int k = 0; k < n_cats; ++k) acc = fmaf(w.ec_w[j * n_cats + k], c[k], acc);
    s_ec[o][j] = gelu_erf(acc + w.ec_b[j]);
  }
  __syncthreads();
  // attn_layer weights only (sdm.py:79,180-182): q = Wq enc + bq, k_o = Wk ec_o + bk, 8 heads of 16
  cta_linear<ACT_NONE>(w.aq_w, w.a_inb, s_enc, s_q, LAT, LAT);
  for (int o = 0; o < NOBJ; ++o) cta_linear<ACT_NONE>(w.ak_w, w.a_inb + LAT, s_ec[o], s_k[o], LAT, CATEMB);
  if (tid < NHEAD * NOBJ) {
    int hd = tid / NOBJ, o = tid % NOBJ;
    float acc = 0.f;
    for (int d = 0; d < 16; ++d) acc = fmaf(s_q[hd * 16 + d], s_k[o][hd * 16 + d], acc);
    // additive float mask, batch-scrambled: head h of global sample bg reads mask[(bg*H+h) mod Bg] (SURVEY trap 2)
    int mrow = (int)((((int64_t)(b + b_off)) * NHEAD + hd) % Bg);
    s_logit[hd][o] = acc * 0.25f + mask_global[mrow * NOBJ + o];
  }
  __syncthreads();
  if (tid < NHEAD) {
    float mx = -INFINITY;
    for (int o = 0; o < NOBJ; ++o) mx = fmaxf(mx, s_logit[tid][o]);
    float sum = 0.f;
    for (int o = 0; o < NOBJ; ++o) {
      float e = expf(s_logit[tid][o] - mx);
      s_logit[tid][o] = e;
      sum += e;
    }
    for (int o = 0; o < NOBJ; ++o) s_logit[tid][o] /= sum;
  }
  __syncthreads();
  if (tid < NOBJ) {
    float acc = 0.f;
    for (int hd = 0; hd < NHEAD; ++hd) acc += s_logit[hd][tid];
    attn_w[(int64_t)b * NOBJ + tid] = acc / NHEAD;
  }
  // translation_layer: [ec_o || enc] 160 -> 128 -> 12, GELU after each (sdm.py:81-87,185-186)
  for (int o = 0; o < NOBJ; ++o) {
    for (int i = tid; i < LAT + CATEMB; i += blockDim.x) s_z[i] = i < CATEMB ? s_ec[o][i] : s_enc[i - CATEMB];
    __syncthreads();
    cta_linear<ACT_GELU>(w.tl0_w, w.tl0_b, s_z, s_t1, LAT, LAT + CATEMB);
    cta_linear<ACT_GELU>(w.tl2_w, w.tl2_b, s_t1, s_tr[o], TRANS, LAT);
  }
  // pcd_attention query projection (head_dim 1, scale 1): qq = Wq12 tr + bq (sdm.py:95,195)
  if (tid < NOBJ * TRANS) {
    int o = tid / TRANS, j = tid % TRANS;
    float acc = 0.f;
    for (int k = 0; k < TRANS; ++k) acc = fmaf(w.pq_w[j * TRANS + k], s_tr[o][k], acc);
    qq_out[((int64_t)b * NOBJ + o) * TRANS + j] = acc + w.p_inb[j];
    tr_out[((int64_t)b * NOBJ + o) * TRANS + j] = s_tr[o][j];
  }
}

// one CTA per sample: ts = W2 silu(W1 pe[t] + b1) + b2; s = [ts || enc]; H1[s, :] = gelu(s * w0 + b0) (sdm.py:108-110,164-166)
__global__ void __launch_bounds__(256) time_embed_kernel(const float* __restrict__ pe, const float* __restrict__ w1,
                                                         const float* __restrict__ b1, const float* __restrict__ w2,
                                                         const float* __restrict__ b2, const int64_t* __restrict__ t,
                                                         const float* __restrict__ enc, const float* __restrict__ up0_w,
                                                         const float* __restrict__ up0_b, float* __restrict__ s256,
                                                         float* __restrict__ H1, float* __restrict__ H1_lo) {
  __shared__ float s_pe[LAT], s_h[LAT], s_s[2 * LAT];
  const int b = blockIdx.x, tid = threadIdx.x;
  const int64_t tt = t[b];
  for (int i = tid; i < LAT; i += blockDim.x) s_pe[i] = pe[tt * LAT + i];
  __syncthreads();
  cta_linear<ACT_SILU>(w1, b1, s_pe, s_h, LAT, LAT);
  cta_linear<ACT_NONE>(w2, b2, s_h, s_s, LAT, LAT);
  for (int i = tid; i < LAT; i += blockDim.x) s_s[LAT + i] = enc ? enc[(int64_t)b * LAT + i] : 0.f;  // (enc == nullptr: time half only)
  __syncthreads();
  for (int i = tid; i < 2 * LAT; i += blockDim.x) s256[(int64_t)b * 2 * LAT + i] = s_s[i];
  for (int i = tid; i < 2 * LAT * 128; i += blockDim.x) {
    int s = i >> 7, j = i & 127;
    const float v = gelu_erf(fmaf(s_s[s], up0_w[j], up0_b[j]));
    const int64_t o = ((int64_t)b * 2 * LAT + s) * 128 + j;
    if (H1_lo != nullptr) {  // consumer is a pre-split 3xTF32 GEMM
      float hi, lo;
      split_tf32(v, hi, lo);
      H1[o] = hi;
      H1_lo[o] = lo;
    } else {
      H1[o] = v;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// POSA decoder.  GroupNorm(8 groups of 8 channels) normalises over ALL points of a sample, so each layer is one
// launch over (sample, 64-point slab) CTAs: the CTA normalises its slab of the previous layer's pre-norm output
// (statistics from the previous launch), applies ReLU and the 64->64 (or 3->64 / 64->3) linear layer out of shared
// memory, writes the pre-norm result and adds its partial (sum, sum of squares) per group in double precision.
// ---------------------------------------------------------------------------------------------
constexpr int HS = 64;  // points per CTA

// stats[b][layer][group][2] (double): sum, sumsq.  A thread holds channels ch = e*4 + chq (e = 0..15) of one point, i.e. two
// channels of every one of the 8 groups (group = ch / 8 = e / 2).
__device__ __forceinline__ void group_stats_accumulate(const float (&acc)[16], int p_valid, double* stats_out, double* s_red) {
  double s[8], q[8];
#pragma unroll
  for (int g = 0; g < 8; ++g) {
    const double a0 = p_valid ? (double)acc[2 * g] : 0.0, a1 = p_valid ? (double)acc[2 * g + 1] : 0.0;
    s[g] = a0 + a1;
    q[g] = a0 * a0 + a1 * a1;
  }
#pragma unroll
  for (int g = 0; g < 8; ++g)
    for (int o = 16; o > 0; o >>= 1) {
      s[g] += __shfl_xor_sync(0xffffffffu, s[g], o);
      q[g] += __shfl_xor_sync(0xffffffffu, q[g], o);
    }
  const int tid = threadIdx.x;
  if ((tid & 31) == 0) {
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      s_red[((tid >> 5) * 8 + g) * 2] = s[g];
      s_red[((tid >> 5) * 8 + g) * 2 + 1] = q[g];
    }
  }
  __syncthreads();
  if (tid < 16) {  // 8 groups x (sum, sumsq)
    double t = 0;
    for (int w = 0; w < 8; ++w) t += s_red[(w * 8 + (tid >> 1)) * 2 + (tid & 1)];
    atomicAdd(&stats_out[tid], t);
  }
}

// layer 0: y0 = W0 pts + b0 (3 -> 64), stats
__global__ void __launch_bounds__(256) human_l0_kernel(HumanWeights w, const float* __restrict__ objs, float* __restrict__ y0,
                                                       double* __restrict__ stats) {
  __shared__ float s_w[64 * 3], s_b[64];
  __shared__ double s_red[8 * 8 * 2];
  const int b = blockIdx.y, p0 = blockIdx.x * HS, tid = threadIdx.x;
  if (tid < 192) s_w[tid] = w.w0[tid];
  if (tid < 64) s_b[tid] = w.b0[tid];
  __syncthreads();
  const int p = p0 + (tid >> 2), chq = tid & 3;
  const float* pt = objs + ((int64_t)b * NOBJ * NPTS + p) * 3;
  const float x = pt[0], y = pt[1], z = pt[2];
  float acc[16];
#pragma unroll
  for (int e = 0; e < 16; ++e) {
    int ch = e * 4 + chq;
    acc[e] = fmaf(s_w[ch * 3 + 2], z, fmaf(s_w[ch * 3 + 1], y, fmaf(s_w[ch * 3], x, s_b[ch])));
  }
  float* o = y0 + ((int64_t)b * NPTS + p) * 64 + chq;
#pragma unroll
  for (int e = 0; e < 16; ++e) o[e * 4] = acc[e];
  group_stats_accumulate(acc, 1, stats + ((int64_t)b * 3 + 0) * 16, s_red);
}

// layers 1, 2: y = W relu(gn(prev)) + b over npts points (64 -> 64), stats.
// 128 threads per 64-point slab, a 4-point x 8-channel register tile per thread: thread (tp, tc) owns points 4 tp .. 4 tp + 3 and
// the eight channels of GroupNorm group tc.  Activations and weights are staged K-MAJOR ([k][point], [k][channel]) so that a
// k-step is three 128-bit shared loads (four points; eight channels) for 32 FMAs -- the first form (one point x sixteen channels
// per thread) issued 17 loads per 64 FMAs and was bound by shared-memory bandwidth.  Each output is still the k-ascending chain
// acc = fma(W[ch][k], a[k], acc) started from the bias.
constexpr int HM_T = 128;
__global__ void __launch_bounds__(HM_T) human_mid_kernel(const float* __restrict__ Wl, const float* __restrict__ bl,
                                                         const float* __restrict__ gamma, const float* __restrict__ beta,
                                                         const float* __restrict__ prev, float* __restrict__ out,
                                                         double* stats, int slot_in, int slot_out, int npts_in, int npts) {
  __shared__ __align__(16) float s_wt[64 * 68];  // [k][ch], row stride 68 floats
  __shared__ __align__(16) float s_at[64 * 68];  // normalised + ReLU input slab [k][p], row stride 68 floats
  __shared__ float s_b[64], s_scale[64], s_shift[64];
  __shared__ double s_red[(HM_T / 32) * 8 * 2];
  const int b = blockIdx.y, p0 = blockIdx.x * HS, tid = threadIdx.x;
  for (int i = tid; i < 64 * 64; i += HM_T) s_wt[(i & 63) * 68 + (i >> 6)] = Wl[i];  // Wl[ch][k] -> [k][ch]
  if (tid < 64) {
    const double* st = stats + ((int64_t)b * 3 + slot_in) * 16;
    int g = tid >> 3;
    double n = (double)npts_in * 8.0;
    double mean = st[g * 2] / n;
    double var = st[g * 2 + 1] / n - mean * mean;
    if (var < 0.0) var = 0.0;
    float rstd = (float)(1.0 / sqrt(var + 1e-5));
    float sc = rstd * gamma[tid];
    s_scale[tid] = sc;
    s_shift[tid] = beta[tid] - (float)mean * sc;
    s_b[tid] = bl[tid];
  }
  __syncthreads();
  for (int i = tid; i < HS * 64; i += HM_T) {
    int pl = i >> 6, k = i & 63, p = p0 + pl;
    float v = p < npts ? prev[((int64_t)b * NPTS + p) * 64 + k] : 0.f;
    s_at[k * 68 + pl] = fmaxf(fmaf(v, s_scale[k], s_shift[k]), 0.0f);
  }
  __syncthreads();
  const int tc = tid & 7, tp = tid >> 3;
  float acc[4][8];
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const float bb = s_b[tc * 8 + c];
#pragma unroll
    for (int q = 0; q < 4; ++q) acc[q][c] = bb;
  }
#pragma unroll 4
  for (int k = 0; k < 64; ++k) {
    const float4 a4 = *reinterpret_cast<const float4*>(&s_at[k * 68 + tp * 4]);
    const float4 w0 = *reinterpret_cast<const float4*>(&s_wt[k * 68 + tc * 8]);
    const float4 w1 = *reinterpret_cast<const float4*>(&s_wt[k * 68 + tc * 8 + 4]);
    const float av[4] = {a4.x, a4.y, a4.z, a4.w};
    const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
      for (int c = 0; c < 8; ++c) acc[q][c] = fmaf(wv[c], av[q], acc[q][c]);
  }
  double gs = 0.0, gq = 0.0;  // this thread's share of group tc: sum, sum of squares
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int p = p0 + tp * 4 + q;
    if (p < npts) {
      float* o = out + ((int64_t)b * NPTS + p) * 64 + tc * 8;
      *reinterpret_cast<float4*>(o) = make_float4(acc[q][0], acc[q][1], acc[q][2], acc[q][3]);
      *reinterpret_cast<float4*>(o + 4) = make_float4(acc[q][4], acc[q][5], acc[q][6], acc[q][7]);
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const double v = (double)acc[q][c];
        gs += v;
        gq += v * v;
      }
    }
  }
  // lanes of a warp: tc = lane & 7, four point groups -> reduce over lane bits 3, 4; then over the four warps
  gs += __shfl_xor_sync(0xffffffffu, gs, 8);
  gq += __shfl_xor_sync(0xffffffffu, gq, 8);
  gs += __shfl_xor_sync(0xffffffffu, gs, 16);
  gq += __shfl_xor_sync(0xffffffffu, gq, 16);
  if ((tid & 31) < 8) {
    s_red[((tid >> 5) * 8 + tc) * 2] = gs;
    s_red[((tid >> 5) * 8 + tc) * 2 + 1] = gq;
  }
  __syncthreads();
  if (tid < 16) {  // 8 groups x (sum, sumsq)
    double t = 0;
    for (int w = 0; w < HM_T / 32; ++w) t += s_red[(w * 8 + (tid >> 1)) * 2 + (tid & 1)];
    atomicAdd(&stats[((int64_t)b * 3 + slot_out) * 16 + tid], t);
  }
}

// final: hm[2p], hm[2p+1] = W3 relu(gn(y2[p])) + b3 for p < 512 (64 -> 3, nearest x2 upsample, first 1024 kept)
__global__ void __launch_bounds__(256) human_out_kernel(HumanWeights w, const float* __restrict__ y2,
                                                        const double* __restrict__ stats, float* __restrict__ hm) {
  __shared__ float s_w3[3 * 64 + 3], s_scale[64], s_shift[64];
  const int b = blockIdx.y, tid = threadIdx.x;
  if (tid < 192) s_w3[tid] = w.w3[tid];
  if (tid < 3) s_w3[192 + tid] = w.b3[tid];
  if (tid < 64) {
    const double* st = stats + ((int64_t)b * 3 + 2) * 16;
    int g = tid >> 3;
    double n = 655.0 * 8.0;
    double mean = st[g * 2] / n;
    double var = st[g * 2 + 1] / n - mean * mean;
    if (var < 0.0) var = 0.0;
    float sc = (float)(1.0 / sqrt(var + 1e-5)) * w.g2[tid];
    s_scale[tid] = sc;
    s_shift[tid] = w.be2[tid] - (float)mean * sc;
  }
  __syncthreads();
  // one warp per point: lanes hold 2 channels each
  const int lane = tid & 31, p = blockIdx.x * 8 + (tid >> 5);
  if (p >= 512) return;
  float2 v = *reinterpret_cast<const float2*>(y2 + ((int64_t)b * NPTS + p) * 64 + lane * 2);
  float a0 = fmaxf(fmaf(v.x, s_scale[lane * 2], s_shift[lane * 2]), 0.f);
  float a1 = fmaxf(fmaf(v.y, s_scale[lane * 2 + 1], s_shift[lane * 2 + 1]), 0.f);
  float r0 = warp_sum(a0 * s_w3[lane * 2] + a1 * s_w3[lane * 2 + 1]);
  float r1 = warp_sum(a0 * s_w3[64 + lane * 2] + a1 * s_w3[64 + lane * 2 + 1]);
  float r2 = warp_sum(a0 * s_w3[128 + lane * 2] + a1 * s_w3[128 + lane * 2 + 1]);
  if (lane < 6) {
    int d = lane % 3;
    float r = (d == 0 ? r0 : (d == 1 ? r1 : r2)) + s_w3[192 + d];
    hm[((int64_t)b * NPTS + 2 * p + lane / 3) * 3 + d] = r;
  }
}

}  // namespace

int launch_cond(const CondWeights& w, const float* text, const float* cats, const float* mask_global, int B, int Bg,
                int b_off, int n_cats, float* enc, float* out_cat, float* attn_w, float* tr, float* qq, cudaStream_t st) {
  cond_kernel<<<B, 256, 0, st>>>(w, text, cats, mask_global, Bg, b_off, n_cats, enc, out_cat, attn_w, tr, qq);
  return 1;
}

int launch_time_embed(const float* pe, const float* w1, const float* b1, const float* w2, const float* b2,
                      const int64_t* t, const float* enc, const float* up0_w, const float* up0_b, int B, float* s256,
                      float* H1, float* H1_lo, cudaStream_t st) {
  time_embed_kernel<<<B, 256, 0, st>>>(pe, w1, b1, w2, b2, t, enc, up0_w, up0_b, s256, H1, H1_lo);
  return 1;
}

int launch_human(const HumanWeights& w, const float* objs, int B, float* scratch, float* hm, cudaStream_t st) {
  // scratch: y0[B,1024,64] | y1[B,1024,64] | stats[B,3,8,2] doubles
  float* y0 = scratch;
  float* y1 = scratch + (size_t)B * NPTS * 64;
  double* stats = reinterpret_cast<double*>(scratch + (size_t)2 * B * NPTS * 64);
  cudaMemsetAsync(stats, 0, sizeof(double) * B * 3 * 16, st);
  human_l0_kernel<<<dim3(NPTS / HS, B), 256, 0, st>>>(w, objs, y0, stats);
  // layer 1 over 1024 points (stats slot 0 -> 1); layer 2 over the first 655 points (stats slot 1 -> 2)
  human_mid_kernel<<<dim3(NPTS / HS, B), HM_T, 0, st>>>(w.w1, w.b1, w.g0, w.be0, y0, y1, stats, 0, 1, NPTS, NPTS);
  human_mid_kernel<<<dim3((655 + HS - 1) / HS, B), HM_T, 0, st>>>(w.w2, w.b2, w.g1, w.be1, y1, y0, stats, 1, 2, NPTS, 655);
  human_out_kernel<<<dim3(512 / 8, B), 256, 0, st>>>(w, y0, stats, hm);
  return 4;
}

}  // namespace lsdm
