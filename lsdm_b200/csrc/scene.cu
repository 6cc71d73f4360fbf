// Scene branch after the backbone (reference model/sdm.py:191-204): the two memory-reinterpreting
// reshapes ("scrambles"), the 12-head head_dim-1 point attention
// collapsed to its single distinct query row, the pointwise 15->3 translation and the masked object sum.
#include "kernels.cuh"

namespace lsdm {

namespace {

constexpr int PA_T = 256;

// One CTA per (b, o').  p1[b,o',p,d] = F[b, g%9, g/9] * w[b, g%9] with g = o'*3072 + p*3 + d   (scramble #1)
// logits_h[j] = qq_h * (Wk_h . p1_j + bk_h); a = softmax_j; ctx_h = sum_j a_hj (Wv_h . p1_j + bv_h); pa = Wo ctx + bo
// pw[b,o',p,:] = gelu(Wpt [p1_p || pa] + bpt)
__global__ void __launch_bounds__(PA_T, 2) point_attention_kernel(SceneWeights w, const float* __restrict__ backbone,
                                                               const float* __restrict__ attn_w,
                                                               const float* __restrict__ qq, float* __restrict__ pa_out,
                                                               float* __restrict__ pw, const int* __restrict__ remap) {
  __shared__ float s_p1[NPTS * 3];
  __shared__ float s_red[PA_T / 32][TRANS][3];
  __shared__ float s_max[TRANS], s_ctx[TRANS], s_pa[TRANS];
  __shared__ float s_wk[TRANS * 3], s_wv[TRANS * 3], s_bk[TRANS], s_bv[TRANS], s_qq[TRANS], s_aw[NOBJ];
  __shared__ int s_src[NOBJ];  // where cloud (b, o) sits in `backbone` (identity, or the de-duplicated order)
  const int bo = blockIdx.x, b = bo / NOBJ, op = bo % NOBJ, tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  if (tid < TRANS * 3) {
    s_wk[tid] = w.pk_w[tid];
    s_wv[tid] = w.pv_w[tid];
  }
  if (tid < TRANS) {
    s_bk[tid] = w.p_inb[TRANS + tid];
    s_bv[tid] = w.p_inb[2 * TRANS + tid];
    s_qq[tid] = qq[(int64_t)bo * TRANS + tid];
  }
  if (tid < NOBJ) {
    s_aw[tid] = attn_w[(int64_t)b * NOBJ + tid];
    s_src[tid] = remap ? remap[b * NOBJ + tid] : b * NOBJ + tid;
  }
  __syncthreads();
  for (int e = tid; e < NPTS * 3; e += PA_T) {
    int g = op * (NPTS * 3) + e;
    int o = g % NOBJ, c = g / NOBJ;
    s_p1[e] = backbone[(int64_t)s_src[o] * (NPTS * 3) + c] * s_aw[o];
  }
  __syncthreads();

  // pass 1: per-head max of the logits
  float lmax[TRANS];
#pragma unroll
  for (int h = 0; h < TRANS; ++h) lmax[h] = -INFINITY;
  for (int p = tid; p < NPTS; p += PA_T) {
    float x = s_p1[p * 3], y = s_p1[p * 3 + 1], z = s_p1[p * 3 + 2];
#pragma unroll
    for (int h = 0; h < TRANS; ++h) {
      float kk = fmaf(s_wk[h * 3 + 2], z, fmaf(s_wk[h * 3 + 1], y, s_wk[h * 3] * x)) + s_bk[h];
      lmax[h] = fmaxf(lmax[h], s_qq[h] * kk);
    }
  }
#pragma unroll
  for (int h = 0; h < TRANS; ++h) {
    float m = warp_max(lmax[h]);
    if (lane == 0) s_red[warp][h][0] = m;
  }
  __syncthreads();
  if (tid < TRANS) {
    float m = s_red[0][tid][0];
    for (int k = 1; k < PA_T / 32; ++k) m = fmaxf(m, s_red[k][tid][0]);
    s_max[tid] = m;
  }
  __syncthreads();
  // pass 2: sum exp and weighted values
  float lsum[TRANS], lacc[TRANS];
#pragma unroll
  for (int h = 0; h < TRANS; ++h) lsum[h] = 0.f, lacc[h] = 0.f;
  for (int p = tid; p < NPTS; p += PA_T) {
    float x = s_p1[p * 3], y = s_p1[p * 3 + 1], z = s_p1[p * 3 + 2];
#pragma unroll
    for (int h = 0; h < TRANS; ++h) {
      float kk = fmaf(s_wk[h * 3 + 2], z, fmaf(s_wk[h * 3 + 1], y, s_wk[h * 3] * x)) + s_bk[h];
      float vv = fmaf(s_wv[h * 3 + 2], z, fmaf(s_wv[h * 3 + 1], y, s_wv[h * 3] * x)) + s_bv[h];
      float e = expf(s_qq[h] * kk - s_max[h]);
      lsum[h] += e;
      lacc[h] = fmaf(e, vv, lacc[h]);
    }
  }
#pragma unroll
  for (int h = 0; h < TRANS; ++h) {
    float s = warp_sum(lsum[h]);
    float a = warp_sum(lacc[h]);
    if (lane == 0) {
      s_red[warp][h][1] = s;
      s_red[warp][h][2] = a;
    }
  }
  __syncthreads();
  if (tid < TRANS) {
    float s = 0.f, a = 0.f;
    for (int k = 0; k < PA_T / 32; ++k) {
      s += s_red[k][tid][1];
      a += s_red[k][tid][2];
    }
    s_ctx[tid] = a / s;
  }
  __syncthreads();
  if (tid < TRANS) {
    float acc = w.po_b[tid];
    for (int k = 0; k < TRANS; ++k) acc = fmaf(w.po_w[tid * TRANS + k], s_ctx[k], acc);
    s_pa[tid] = acc;
    pa_out[(int64_t)bo * TRANS + tid] = acc;
  }
  __syncthreads();
  // pointwise translate: 15 -> 3 GELU; the pa part is constant over the cloud
  float base[3];
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    float acc = w.pt_b[d];
    for (int k = 0; k < TRANS; ++k) acc = fmaf(w.pt_w[d * 15 + 3 + k], s_pa[k], acc);
    base[d] = acc;
  }
  float wx[9];
#pragma unroll
  for (int d = 0; d < 3; ++d)
#pragma unroll
    for (int k = 0; k < 3; ++k) wx[d * 3 + k] = w.pt_w[d * 15 + k];
  float* out = pw + (int64_t)bo * NPTS * 3;
  for (int p = tid; p < NPTS; p += PA_T) {
    float x = s_p1[p * 3], y = s_p1[p * 3 + 1], z = s_p1[p * 3 + 2];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      float v = fmaf(wx[d * 3 + 2], z, fmaf(wx[d * 3 + 1], y, fmaf(wx[d * 3], x, base[d])));
      out[p * 3 + d] = gelu_erf(v);
    }
  }
}

// pcd_out[b,p,d] = (sum_o pw[b,o,p,d] * mask[(f div 9) mod Bg, f mod 9] + hm[b,p,d]) / 2,
// f = (((b+b_off)*9 + o)*1024 + p)*3 + d   (scramble #2, global batch)
__global__ void __launch_bounds__(256) scene_mix_kernel(const float* __restrict__ pw, const float* __restrict__ hm,
                                                        const float* __restrict__ mask_global, int B, int Bg, int b_off,
                                                        float* __restrict__ pcd_out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)B * NPTS * 3) return;
  int b = (int)(i / (NPTS * 3));
  int e = (int)(i % (NPTS * 3));
  float acc = 0.f;
#pragma unroll
  for (int o = 0; o < NOBJ; ++o) {
    int64_t f = (((int64_t)(b + b_off) * NOBJ + o) * NPTS * 3) + e;
    float m = mask_global[((f / NOBJ) % Bg) * NOBJ + (f % NOBJ)];
    acc += pw[((int64_t)b * NOBJ + o) * NPTS * 3 + e] * m;
  }
  pcd_out[i] = (acc + hm[i]) / 2.0f;
}

}  // namespace

int launch_point_attention(const SceneWeights& w, const float* backbone, const float* attn_w, const float* qq, int B,
                           float* pa, float* pw, cudaStream_t st, const int* remap) {
  point_attention_kernel<<<B * NOBJ, PA_T, 0, st>>>(w, backbone, attn_w, qq, pa, pw, remap);
  return 1;
}

int launch_scene_mix(const float* pw, const float* hm, const float* mask_global, int B, int Bg, int b_off, float* pcd_out,
                     cudaStream_t st) {
  int64_t n = (int64_t)B * NPTS * 3;
  scene_mix_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(pw, hm, mask_global, B, Bg, b_off, pcd_out);
  return 1;
}

}  // namespace lsdm
