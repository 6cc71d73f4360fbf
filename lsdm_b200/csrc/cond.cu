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
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int n = warp; n < N; n += nw) {
    const float* w = W + (int64_t)n * K;
    float acc = 0.f;
    for (int k = lane; k < K; k += 32) acc = fmaf(w[k], x[k], acc);
    acc = warp_sum(acc);
    if (lane == 0) y[n] = apply_act<ACT>(acc + (b ? b[n] : 0.f));
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
                                                         float* __restrict__ H1) {
  __shared__ float s_pe[LAT], s_h[LAT], s_s[2 * LAT];
  const int b = blockIdx.x, tid = threadIdx.x;
  const int64_t tt = t[b];
  for (int i = tid; i < LAT; i += blockDim.x) s_pe[i] = pe[tt * LAT + i];
  __syncthreads();
  cta_linear<ACT_SILU>(w1, b1, s_pe, s_h, LAT, LAT);
  cta_linear<ACT_NONE>(w2, b2, s_h, s_s, LAT, LAT);
  for (int i = tid; i < LAT; i += blockDim.x) s_s[LAT + i] = enc[(int64_t)b * LAT + i];
  __syncthreads();
  for (int i = tid; i < 2 * LAT; i += blockDim.x) s256[(int64_t)b * 2 * LAT + i] = s_s[i];
  for (int i = tid; i < 2 * LAT * 128; i += blockDim.x) {
    int s = i >> 7, j = i & 127;
    H1[((int64_t)b * 2 * LAT + s) * 128 + j] = gelu_erf(fmaf(s_s[s], up0_w[j], up0_b[j]));
  }
}

// ---------------------------------------------------------------------------------------------
// POSA decoder: one CTA (256 threads) per sample.  Pre-norm activations go through a global scratch
// (L2-resident, 256 KB per sample); GroupNorm(8 groups of 8 channels) statistics over all points of the
// sample are accumulated in double and applied while the next layer reads the scratch.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void gn_stats(const float* __restrict__ y, int npts, float* s_mean, float* s_rstd,
                                         double* s_acc) {
  // y[npts][64]; thread owns channel tid%64 over points tid/64 + 4k  (256 threads)
  const int tid = threadIdx.x, ch = tid & 63, p0 = tid >> 6;
  double s = 0.0, ss = 0.0;
  for (int p = p0; p < npts; p += 4) {
    double v = (double)y[p * 64 + ch];
    s += v;
    ss += v * v;
  }
  // reduce the 8 channels of a group (adjacent lanes) then across the 4 point-slices
  for (int o = 4; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    ss += __shfl_xor_sync(0xffffffffu, ss, o);
  }
  if ((ch & 7) == 0) {
    s_acc[(p0 * 8 + (ch >> 3)) * 2 + 0] = s;
    s_acc[(p0 * 8 + (ch >> 3)) * 2 + 1] = ss;
  }
  __syncthreads();
  if (tid < 8) {
    double a = 0.0, q = 0.0;
    for (int k = 0; k < 4; ++k) {
      a += s_acc[(k * 8 + tid) * 2];
      q += s_acc[(k * 8 + tid) * 2 + 1];
    }
    double n = (double)npts * 8.0;
    double mean = a / n;
    double var = q / n - mean * mean;
    if (var < 0.0) var = 0.0;
    s_mean[tid] = (float)mean;
    s_rstd[tid] = (float)(1.0 / sqrt(var + 1e-5));
  }
  __syncthreads();
}

__global__ void __launch_bounds__(256) human_kernel(HumanWeights w, const float* __restrict__ objs,
                                                    float* __restrict__ scratch, float* __restrict__ hm) {
  __shared__ __align__(16) float s_w[64 * 64];
  __shared__ float s_b[64], s_g[64], s_be[64], s_mean[8], s_rstd[8];
  __shared__ double s_acc[64];
  __shared__ float s_w3[3 * 64 + 3];
  const int b = blockIdx.x, tid = threadIdx.x;
  const float* pts = objs + (int64_t)b * NOBJ * NPTS * 3;  // slot 0 = human
  float* y0 = scratch + (int64_t)b * 2 * NPTS * 64;
  float* y1 = y0 + NPTS * 64;

  // layer 0: 3 -> 64 (pre-norm) -> y0
  for (int i = tid; i < 64 * 3; i += 256) s_w[i] = w.w0[i];
  if (tid < 64) s_b[tid] = w.b0[tid];
  __syncthreads();
  for (int i = tid; i < NPTS * 64; i += 256) {
    int p = i >> 6, ch = i & 63;
    float v = s_b[ch];
    v = fmaf(s_w[ch * 3 + 0], pts[p * 3 + 0], v);
    v = fmaf(s_w[ch * 3 + 1], pts[p * 3 + 1], v);
    v = fmaf(s_w[ch * 3 + 2], pts[p * 3 + 2], v);
    y0[i] = v;
  }
  __syncthreads();
  gn_stats(y0, NPTS, s_mean, s_rstd, s_acc);

  // layers 1 and 2: relu(gn(prev)) -> 64 -> 64 (pre-norm)
  for (int layer = 1; layer <= 2; ++layer) {
    const float* wl = layer == 1 ? w.w1 : w.w2;
    const float* bl = layer == 1 ? w.b1 : w.b2;
    const float* gp = layer == 1 ? w.g0 : w.g1;
    const float* bp = layer == 1 ? w.be0 : w.be1;
    const float* src = layer == 1 ? y0 : y1;
    float* dst = layer == 1 ? y1 : y0;
    const int npts = layer == 1 ? NPTS : 655;
    for (int i = tid; i < 64 * 64; i += 256) s_w[i] = wl[i];
    if (tid < 64) {
      s_b[tid] = bl[tid];
      s_g[tid] = gp[tid];
      s_be[tid] = bp[tid];
    }
    __syncthreads();
    // thread = (point slice tid/64, out channel tid%64): warp reads one input row (broadcast) and 32 weight rows
    const int ch = tid & 63;
    for (int p = tid >> 6; p < npts; p += 4) {
      const float* xr = src + p * 64;
      float acc = s_b[ch];
#pragma unroll 8
      for (int k = 0; k < 64; ++k) {
        float a = fmaxf((xr[k] - s_mean[k >> 3]) * s_rstd[k >> 3] * s_g[k] + s_be[k], 0.0f);
        acc = fmaf(s_w[ch * 64 + k], a, acc);
      }
      dst[p * 64 + ch] = acc;
    }
    __syncthreads();
    gn_stats(dst, npts, s_mean, s_rstd, s_acc);
  }
  // final: relu(gn2(y0[:655])) -> 3, nearest x2 upsample, keep 1024: hm[p] = out[p/2]
  for (int i = tid; i < 3 * 64; i += 256) s_w3[i] = w.w3[i];
  if (tid < 3) s_w3[192 + tid] = w.b3[tid];
  if (tid < 64) {
    s_g[tid] = w.g2[tid];
    s_be[tid] = w.be2[tid];
  }
  __syncthreads();
  for (int i = tid; i < 512 * 3; i += 256) {
    int p = i / 3, d = i % 3;
    const float* xr = y0 + p * 64;
    float acc = s_w3[192 + d];
    for (int k = 0; k < 64; ++k) {
      float a = fmaxf((xr[k] - s_mean[k >> 3]) * s_rstd[k >> 3] * s_g[k] + s_be[k], 0.0f);
      acc = fmaf(s_w3[d * 64 + k], a, acc);
    }
    float* o = hm + ((int64_t)b * NPTS + 2 * p) * 3;
    o[d] = acc;
    o[3 + d] = acc;
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
                      float* H1, cudaStream_t st) {
  time_embed_kernel<<<B, 256, 0, st>>>(pe, w1, b1, w2, b2, t, enc, up0_w, up0_b, s256, H1);
  return 1;
}

int launch_human(const HumanWeights& w, const float* objs, int B, float* scratch, float* hm, cudaStream_t st) {
  human_kernel<<<B, 256, 0, st>>>(w, objs, scratch, hm);
  return 1;
}

}  // namespace lsdm
