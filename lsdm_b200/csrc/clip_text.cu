// CLIP text tower (reference model/sdm.py:245-259 `_encode_text_clip` -> clip_model.encode_text; the model is openai/CLIP's
// `CLIP.encode_text`, ViT-B/32 text side: width 512, 12 pre-LN residual blocks, 8 heads x 64, causal mask, QuickGELU MLP,
// ln_final, feature of the EOT token (argmax of the token ids) times text_projection).
//
//   x = token_embedding[tokens] + positional_embedding
//   for each block:  x += out_proj(MHA_causal(ln_1(x)));  x += c_proj(quick_gelu(c_fc(ln_2(x))))
//   out[b] = ln_final(x)[b, argmax(tokens[b])] @ text_projection
//
// Only the first `seq_len` positions are computed: attention is causal, so positions after the EOT token never influence
// the EOT feature and dropping them is bit-identical (the reference tokenises to 22 positions and zero-pads to 77).
// Dense layers: the tcgen05 GEMM of gemm_tc.cu (3xTF32 by default: fp32-grade accuracy on the tensor cores); LayerNorm,
// the L x L causal attention per (sample, head), the embedding gather and the final projection are fp32 CUDA-core kernels.
#include <climits>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/lsdm_b200.h"
#include "kernels.cuh"
#include "ln.cuh"

using namespace lsdm;

struct lsdm_clip {
  int device = 0;
  std::unordered_map<std::string, float*> w;
  std::unordered_map<std::string, std::vector<int64_t>> shape;
  bool finalized = false;
  int width = 0, layers = 0, heads = 0, ctx = 0, vocab = 0, ff = 0, embed = 0;
  int precision = 2;
  int64_t launches = 0;
  const float* W(const std::string& k) const { return w.at(k); }
};

namespace {

constexpr int HEAD_DIM = 64;

// warp per row r = (b, p), p < L:  x[r] = tok_emb[tokens[b, p]] + pos[p];  h[r] = LN(x[r])
template <int PER>
__global__ void __launch_bounds__(256) clip_embed_ln_kernel(const int32_t* __restrict__ tokens, int ctx, int vocab, const float* __restrict__ tok_emb,
                                                            const float* __restrict__ pos, const float* __restrict__ g, const float* __restrict__ b,
                                                            int rows, int L, float* __restrict__ x, float* __restrict__ h) {
  constexpr int WIDTH = 32 * PER;
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (r >= rows) return;
  const int bi = r / L, p = r % L;
  int tok = tokens[(int64_t)bi * ctx + p];
  tok = tok < 0 ? 0 : (tok >= vocab ? vocab - 1 : tok);
  float v[PER];
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    const int c = i * 32 + lane;
    v[i] = tok_emb[(int64_t)tok * WIDTH + c] + pos[(int64_t)p * WIDTH + c];
    x[(int64_t)r * WIDTH + c] = v[i];
  }
  layer_norm_row<PER>(v, g, b, lane, h + (int64_t)r * WIDTH);
}

// Causal multi-head attention of one (head, sample): qkv[rows, 3*width] (q | k | v column blocks, head h = columns
// h*64..h*64+63 of each) -> att[rows, width].  L <= 77 keys: K (row stride 65: conflict-free column reads) and V in shared
// memory, one warp per query row: lanes own keys for the scores / softmax and output channels for P.V.
__global__ void __launch_bounds__(128) clip_attn_kernel(const float* __restrict__ qkv, int L, int width, float* __restrict__ att) {
  extern __shared__ float sm[];
  float* sk = sm;                      // [L][65]
  float* sv = sk + L * (HEAD_DIM + 1);  // [L][64]
  float* sq = sv + L * HEAD_DIM;        // [4 warps][64]
  float* sp = sq + 4 * HEAD_DIM;        // [4 warps][L]
  const int h = blockIdx.x, bi = blockIdx.y, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t ld = 3 * (int64_t)width;
  const float* base = qkv + (int64_t)bi * L * ld + h * HEAD_DIM;
  for (int e = tid; e < L * HEAD_DIM; e += blockDim.x) {
    const int j = e / HEAD_DIM, d = e % HEAD_DIM;
    sk[j * (HEAD_DIM + 1) + d] = base[j * ld + width + d];
    sv[j * HEAD_DIM + d] = base[j * ld + 2 * width + d];
  }
  __syncthreads();
  float* q = sq + warp * HEAD_DIM;
  float* p = sp + warp * L;
  for (int i = warp; i < L; i += 4) {
    q[lane] = base[i * ld + lane] * 0.125f;  // q / sqrt(64), as torch's MHA scales the projected query
    q[lane + 32] = base[i * ld + lane + 32] * 0.125f;
    __syncwarp();
    float s[3], mx = -INFINITY;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const int j = c * 32 + lane;
      s[c] = -INFINITY;
      if (j <= i) {  // causal: key j visible to query i iff j <= i
        float a = 0.f;
        const float* kr = sk + j * (HEAD_DIM + 1);
#pragma unroll 16
        for (int d = 0; d < HEAD_DIM; ++d) a = fmaf(q[d], kr[d], a);
        s[c] = a;
      }
      mx = fmaxf(mx, s[c]);
    }
    mx = warp_max(mx);
    float sum = 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      s[c] = (c * 32 + lane) <= i ? expf(s[c] - mx) : 0.f;
      sum += s[c];
    }
    const float inv = 1.0f / warp_sum(sum);
#pragma unroll
    for (int c = 0; c < 3; ++c)
      if (c * 32 + lane < L) p[c * 32 + lane] = s[c] * inv;
    __syncwarp();
    float o0 = 0.f, o1 = 0.f;
    for (int j = 0; j <= i; ++j) {
      const float pj = p[j];
      o0 = fmaf(pj, sv[j * HEAD_DIM + lane], o0);
      o1 = fmaf(pj, sv[j * HEAD_DIM + lane + 32], o1);
    }
    float* dst = att + ((int64_t)bi * L + i) * width + h * HEAD_DIM;
    dst[lane] = o0;
    dst[lane + 32] = o1;
    __syncwarp();
  }
}

// out[b, e] = sum_k hf[b*L + eot_b, k] * proj[k, e], eot_b = first argmax of tokens[b, :ctx]; NaN if eot_b >= L
__global__ void __launch_bounds__(256) clip_pool_project_kernel(const int32_t* __restrict__ tokens, int ctx, const float* __restrict__ hf, int L,
                                                                int width, const float* __restrict__ proj, int embed, float* __restrict__ out) {
  extern __shared__ float srow[];
  __shared__ int s_eot;
  const int bi = blockIdx.x, tid = threadIdx.x;
  if (tid < 32) {
    int best = INT_MIN, pos = 0;
    for (int p = tid; p < ctx; p += 32) {
      const int t = tokens[(int64_t)bi * ctx + p];
      if (t > best) {
        best = t;
        pos = p;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const int ob = __shfl_xor_sync(0xffffffffu, best, o), op = __shfl_xor_sync(0xffffffffu, pos, o);
      if (ob > best || (ob == best && op < pos)) {
        best = ob;
        pos = op;
      }
    }
    if (tid == 0) s_eot = pos;
  }
  __syncthreads();
  const int eot = s_eot;
  if (eot >= L) {
    for (int e = tid; e < embed; e += blockDim.x) out[(int64_t)bi * embed + e] = nanf("");
    return;
  }
  for (int k = tid; k < width; k += blockDim.x) srow[k] = hf[((int64_t)bi * L + eot) * width + k];
  __syncthreads();
  for (int e = tid; e < embed; e += blockDim.x) {
    float a = 0.f;
    for (int k = 0; k < width; ++k) a = fmaf(srow[k], proj[(int64_t)k * embed + e], a);
    out[(int64_t)bi * embed + e] = a;
  }
}

std::string blk(int l, const char* name) { return "transformer.resblocks." + std::to_string(l) + "." + name; }

int clip_gemm(lsdm_clip* h, const float* A, int64_t lda, const float* W, const float* bias, float* C, int M, int N, int K, int act,
              cudaStream_t st) {
  GemmArgs g{};
  g.A = A; g.lda = lda;
  g.W = W; g.ldw = K;
  g.C = C; g.ldc = N;
  g.bias = bias; g.bias_mode = bias ? 1 : 0;
  g.M = M; g.N = N; g.K = K; g.batch = 1;
  g.act = act; g.precision = h->precision;
  const int r = (h->precision >= 1 && gemm_tc_eligible(g)) ? launch_gemm_tc(g, st) : launch_gemm_simt(g, st);
  if (r > 0) h->launches += r;
  return r;
}

struct ClipWs {
  float *x, *hn, *qkv, *att, *y, *ffh;
  size_t bytes;
};
ClipWs carve(const lsdm_clip* h, void* base, int rows) {
  char* p = static_cast<char*>(base);
  size_t off = 0;
  auto take = [&](size_t n) {
    off = (off + 255) & ~size_t(255);
    float* r = p ? reinterpret_cast<float*>(p + off) : nullptr;
    off += n * sizeof(float);
    return r;
  };
  ClipWs w{};
  const size_t R = (size_t)rows, Wd = (size_t)h->width;
  w.x = take(R * Wd);
  w.hn = take(R * Wd);
  w.qkv = take(R * 3 * Wd);
  w.att = take(R * Wd);
  w.y = take(R * Wd);
  w.ffh = take(R * (size_t)h->ff);
  w.bytes = (off + 255) & ~size_t(255);
  return w;
}

}  // namespace

#define CCK(call)                                                                              \
  do {                                                                                         \
    cudaError_t e__ = (call);                                                                  \
    if (e__ != cudaSuccess) return set_error(LSDM_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e__)); \
  } while (0)

extern "C" {

LSDM_API int lsdm_clip_create(lsdm_clip** out, int32_t device) {
  if (!out) return set_error(LSDM_EINVAL, "null argument");
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || device < 0 || device >= n)
    return set_error(LSDM_ECUDA, "lsdm_clip_create: no such CUDA device (there is no CPU fallback)");
  lsdm_clip* h = new lsdm_clip();
  h->device = device;
  *out = h;
  return LSDM_OK;
}

LSDM_API void lsdm_clip_destroy(lsdm_clip* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  for (auto& kv : h->w) cudaFree(kv.second);
  delete h;
}

LSDM_API int lsdm_clip_load_weight(lsdm_clip* h, const char* key, const float* data, const int64_t* shape, int32_t ndim, void* stream) {
  if (!h || !key || !data || !shape || ndim < 1 || ndim > 4) return set_error(LSDM_EINVAL, "bad argument");
  CCK(cudaSetDevice(h->device));
  int64_t numel = 1;
  std::vector<int64_t> shp(shape, shape + ndim);
  for (int64_t d : shp) numel *= d;
  if (numel <= 0) return set_error(LSDM_EINVAL, std::string("empty tensor: ") + key);
  auto it = h->w.find(key);
  if (it != h->w.end() && h->shape[key] != shp) {
    cudaFree(it->second);
    h->w.erase(it);
    it = h->w.end();
  }
  float* dst = it != h->w.end() ? it->second : nullptr;
  if (!dst) {
    CCK(cudaMalloc(&dst, sizeof(float) * numel));
    h->w[key] = dst;
    h->shape[key] = shp;
  }
  CCK(cudaMemcpyAsync(dst, data, sizeof(float) * numel, cudaMemcpyDefault, (cudaStream_t)stream));
  h->finalized = false;
  return LSDM_OK;
}

// Infers the configuration from the tensor shapes (as openai/CLIP's build_model does) and checks every tensor of the text tower.
LSDM_API int lsdm_clip_finalize(lsdm_clip* h) {
  if (!h) return set_error(LSDM_EINVAL, "null handle");
  auto need = [&](const std::string& k, std::vector<int64_t> shp) -> bool {
    auto it = h->shape.find(k);
    return it != h->shape.end() && it->second == shp;
  };
  auto have = [&](const std::string& k) { return h->shape.count(k) != 0; };
  if (!have("ln_final.weight") || !have("token_embedding.weight") || !have("positional_embedding") || !have("text_projection"))
    return set_error(LSDM_ESTATE, "CLIP text tower: missing ln_final / token_embedding / positional_embedding / text_projection");
  const int64_t W = h->shape["ln_final.weight"][0];
  int L = 0;
  while (have(blk(L, "ln_1.weight"))) ++L;
  if (L == 0) return set_error(LSDM_ESTATE, "CLIP text tower: no transformer.resblocks.* tensors");
  if (h->shape["token_embedding.weight"].size() != 2 || h->shape["positional_embedding"].size() != 2 || h->shape["text_projection"].size() != 2)
    return set_error(LSDM_EINVAL, "CLIP text tower: bad embedding / projection rank");
  h->width = (int)W;
  h->layers = L;
  h->heads = (int)(W / HEAD_DIM);
  h->vocab = (int)h->shape["token_embedding.weight"][0];
  h->ctx = (int)h->shape["positional_embedding"][0];
  h->embed = (int)h->shape["text_projection"][1];
  h->ff = (int)(4 * W);
  if (W != 512 && W != 768 && W != 1024) return set_error(LSDM_EINVAL, "CLIP text tower: width must be 512, 768 or 1024 (head dim 64)");
  if (W % 256 != 0) return set_error(LSDM_EINVAL, "CLIP text tower: width must be a multiple of 256");
  if (h->ctx > 96) return set_error(LSDM_EINVAL, "CLIP text tower: context length above 96");
  bool ok = need("token_embedding.weight", {h->vocab, W}) && need("positional_embedding", {h->ctx, W}) && need("ln_final.weight", {W}) &&
            need("ln_final.bias", {W}) && need("text_projection", {W, h->embed});
  for (int l = 0; ok && l < L; ++l) {
    ok = need(blk(l, "ln_1.weight"), {W}) && need(blk(l, "ln_1.bias"), {W}) && need(blk(l, "ln_2.weight"), {W}) && need(blk(l, "ln_2.bias"), {W}) &&
         need(blk(l, "attn.in_proj_weight"), {3 * W, W}) && need(blk(l, "attn.in_proj_bias"), {3 * W}) &&
         need(blk(l, "attn.out_proj.weight"), {W, W}) && need(blk(l, "attn.out_proj.bias"), {W}) &&
         need(blk(l, "mlp.c_fc.weight"), {4 * W, W}) && need(blk(l, "mlp.c_fc.bias"), {4 * W}) &&
         need(blk(l, "mlp.c_proj.weight"), {W, 4 * W}) && need(blk(l, "mlp.c_proj.bias"), {W});
    if (!ok) return set_error(LSDM_ESTATE, "CLIP text tower: missing or mis-shaped tensor in resblock " + std::to_string(l));
  }
  if (!ok) return set_error(LSDM_ESTATE, "CLIP text tower: missing or mis-shaped embedding / final tensors");
  h->finalized = true;
  return LSDM_OK;
}

LSDM_API int lsdm_clip_set_precision(lsdm_clip* h, int32_t precision) {
  if (!h || precision < 0 || precision > 2) return set_error(LSDM_EINVAL, "precision must be 0 (fp32), 1 (tf32) or 2 (3xtf32)");
  h->precision = precision;
  return LSDM_OK;
}

LSDM_API size_t lsdm_clip_workspace_bytes(const lsdm_clip* h, int32_t batch, int32_t seq_len) {
  if (!h || !h->finalized || batch <= 0 || seq_len <= 0) return 0;
  return carve(h, nullptr, batch * seq_len).bytes;
}

LSDM_API int64_t lsdm_clip_launch_count(const lsdm_clip* h) { return h ? h->launches : 0; }

LSDM_API int lsdm_clip_dims(const lsdm_clip* h, int32_t* width, int32_t* layers, int32_t* heads, int32_t* ctx, int32_t* vocab, int32_t* embed) {
  if (!h || !h->finalized) return set_error(LSDM_ESTATE, "CLIP text tower not finalised");
  if (width) *width = h->width;
  if (layers) *layers = h->layers;
  if (heads) *heads = h->heads;
  if (ctx) *ctx = h->ctx;
  if (vocab) *vocab = h->vocab;
  if (embed) *embed = h->embed;
  return LSDM_OK;
}

LSDM_API int lsdm_clip_encode_text(lsdm_clip* h, const int32_t* tokens, int32_t batch, int32_t seq_len, void* workspace, size_t workspace_bytes,
                                   float* out, void* stream) {
  if (!h || !tokens || !out || !workspace || batch <= 0) return set_error(LSDM_EINVAL, "bad argument");
  if (!h->finalized) return set_error(LSDM_ESTATE, "CLIP text tower not finalised (lsdm_clip_finalize)");
  if (seq_len < 1 || seq_len > h->ctx) return set_error(LSDM_EINVAL, "seq_len must be in [1, context length]");
  const int rows = batch * seq_len, W = h->width, L = seq_len;
  if (((uintptr_t)workspace & 255) != 0) return set_error(LSDM_EINVAL, "workspace must be 256-byte aligned");
  ClipWs w = carve(h, workspace, rows);
  if (w.bytes > workspace_bytes) return set_error(LSDM_ENOMEM, "workspace too small (lsdm_clip_workspace_bytes)");
  cudaStream_t st = (cudaStream_t)stream;
  const int wblocks = (rows + 7) / 8;
  auto embed_ln = [&](const float* g, const float* b) {
    const float *te = h->W("token_embedding.weight"), *pe = h->W("positional_embedding");
    if (W == 512) clip_embed_ln_kernel<16><<<wblocks, 256, 0, st>>>(tokens, h->ctx, h->vocab, te, pe, g, b, rows, L, w.x, w.hn);
    else if (W == 768) clip_embed_ln_kernel<24><<<wblocks, 256, 0, st>>>(tokens, h->ctx, h->vocab, te, pe, g, b, rows, L, w.x, w.hn);
    else clip_embed_ln_kernel<32><<<wblocks, 256, 0, st>>>(tokens, h->ctx, h->vocab, te, pe, g, b, rows, L, w.x, w.hn);
    ++h->launches;
  };
  auto add_ln = [&](const float* g, const float* b) {
    if (W == 512) add_ln_kernel<16><<<wblocks, 256, 0, st>>>(w.x, w.y, g, b, rows, w.hn);
    else if (W == 768) add_ln_kernel<24><<<wblocks, 256, 0, st>>>(w.x, w.y, g, b, rows, w.hn);
    else add_ln_kernel<32><<<wblocks, 256, 0, st>>>(w.x, w.y, g, b, rows, w.hn);
    ++h->launches;
  };
  const size_t attn_smem = sizeof(float) * ((size_t)L * (HEAD_DIM + 1) + (size_t)L * HEAD_DIM + 4 * HEAD_DIM + 4 * (size_t)L);
  static PerDeviceOnce attr_done;
  CCK(smem_opt_in(attr_done, clip_attn_kernel, 64 * 1024));
  embed_ln(h->W(blk(0, "ln_1.weight")), h->W(blk(0, "ln_1.bias")));
  for (int l = 0; l < h->layers; ++l) {
    if (clip_gemm(h, w.hn, W, h->W(blk(l, "attn.in_proj_weight")), h->W(blk(l, "attn.in_proj_bias")), w.qkv, rows, 3 * W, W, ACT_NONE, st) < 0)
      return set_error(LSDM_EINVAL, "CLIP qkv gemm");
    clip_attn_kernel<<<dim3(h->heads, batch), 128, attn_smem, st>>>(w.qkv, L, W, w.att);
    ++h->launches;
    if (clip_gemm(h, w.att, W, h->W(blk(l, "attn.out_proj.weight")), h->W(blk(l, "attn.out_proj.bias")), w.y, rows, W, W, ACT_NONE, st) < 0)
      return set_error(LSDM_EINVAL, "CLIP out_proj gemm");
    add_ln(h->W(blk(l, "ln_2.weight")), h->W(blk(l, "ln_2.bias")));
    if (clip_gemm(h, w.hn, W, h->W(blk(l, "mlp.c_fc.weight")), h->W(blk(l, "mlp.c_fc.bias")), w.ffh, rows, h->ff, W, ACT_QGELU, st) < 0)
      return set_error(LSDM_EINVAL, "CLIP c_fc gemm");
    if (clip_gemm(h, w.ffh, h->ff, h->W(blk(l, "mlp.c_proj.weight")), h->W(blk(l, "mlp.c_proj.bias")), w.y, rows, W, h->ff, ACT_NONE, st) < 0)
      return set_error(LSDM_EINVAL, "CLIP c_proj gemm");
    if (l + 1 < h->layers) add_ln(h->W(blk(l + 1, "ln_1.weight")), h->W(blk(l + 1, "ln_1.bias")));
    else add_ln(h->W("ln_final.weight"), h->W("ln_final.bias"));
  }
  clip_pool_project_kernel<<<batch, 256, sizeof(float) * W, st>>>(tokens, h->ctx, w.hn, L, W, h->W("text_projection"), h->embed, out);
  ++h->launches;
  CCK(cudaPeekAtLastError());
  return LSDM_OK;
}

}  // extern "C"
