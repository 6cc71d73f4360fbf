// C ABI of lsdm_b200 (include/lsdm_b200.h): handle, weight registry, workspace carving and the
// orchestration of the kernels for encode_conditions / denoise_step / forward / sample_loop.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/lsdm_b200.h"
#include "kernels.cuh"
#include "train_bw.cuh"

using namespace lsdm;

namespace {

thread_local std::string g_err;
int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}
#define CK(call)                                                                                         \
  do {                                                                                                   \
    cudaError_t e__ = (call);                                                                            \
    if (e__ != cudaSuccess)                                                                              \
      return fail(LSDM_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e__));                      \
  } while (0)

}  // namespace
// error channel shared with the other translation units that export C ABI functions (clip_text.cu)
namespace lsdm {
int set_error(int code, const std::string& msg) { return fail(code, msg); }
}  // namespace lsdm
namespace {

struct WEntry {
  std::string key;
  std::vector<int64_t> shape;
  int64_t numel;
  int64_t off;  // float offset into the raw arena; -1 for ignored (int64) entries
  bool loaded;
};

struct SASpec { const char* name; int npoint; double radius; int cin; int mlp[3]; int N; };
struct FPSpec { const char* name; int cin; int nl; int mlp[3]; int Ca; int Cb; };
const SASpec kSA[4] = {{"sa1", 1024, 0.1, 6, {32, 32, 64}, 1024},
                       {"sa2", 256, 0.2, 67, {64, 64, 128}, 1024},
                       {"sa3", 64, 0.4, 131, {128, 128, 256}, 256},
                       {"sa4", 16, 0.8, 259, {256, 256, 512}, 64}};
const FPSpec kFP[4] = {{"fp4", 768, 2, {256, 256, 0}, 256, 512},
                       {"fp3", 384, 2, {256, 256, 0}, 128, 256},
                       {"fp2", 320, 2, {256, 128, 0}, 64, 256},
                       {"fp1", 128, 3, {128, 128, 128}, 0, 128}};

struct Bump {
  char* base;
  size_t off;
  explicit Bump(void* b) : base(static_cast<char*>(b)), off(0) {}
  template <typename T>
  T* take(size_t n) {
    off = (off + 255) & ~size_t(255);
    T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
    off += n * sizeof(T);
    return p;
  }
};

struct Workspace {
  // conditions
  float* human_scratch;
  double* bn_stats;
  int64_t* t_dev;
  // selection results (depend only on the clouds and the FPS start draws); three sets so that lsdm_sample_loop can run the
  // selection chain up to two steps ahead on a side stream while the dense layers and the x0 network of earlier steps run
  struct Sel {
    float *enc, *out_cat, *attn_w, *tr, *qq, *hm;  // step-invariant condition-MLP / human-decoder outputs
    int64_t* fps_start;
    int *idx[4], *grp[4], *nn_idx[4];
    float *xyz[5], *nn_w[4];
    int *plan_rows, *plan_used, *plan_tiles, *plan_off, *plan_n;  // sa1 on distinct rows: packed tiles of the level-0 groups
  } sel[3];  // NSEL
  int cur;  // set used by the last encode (debug taps)
  float* feat[5];
  // absent-cloud de-duplication (lsdm_sample_loop): compacted clouds, cloud -> compact position, compact -> cloud
  float* clouds_c;
  int *remap, *active, *n_active;
  float *tA, *tB, *tP, *g3, *g2, *g1, *backbone, *pa, *pw, *pcd_out[3];  // pcd_out: one per selection set (the step of k reads it while dense(k+1), dense(k+2) write the others)
  // step
  float *s256, *H1, *H2, *embpre, *cat, *h1, *c1, *c2, *f1, *x0, *guiding, *loss_scratch;
  // sa1 in cloud order (`loop_invariants` bit 2): level-1 features of every point, the ball-query groups and the distinct-row plan
  // they were computed from -- written once per lsdm_sample_loop call, permuted per step by the level-0 FPS order
  float *f1canon, *p2canon;
  int *c_grp, *c_plan_rows, *c_plan_used, *c_plan_tiles, *c_plan_off, *c_plan_n;
  int64_t* tb_t;  // time half of the split embedding for TIME_BATCH steps at once
  float *tb_s256, *tb_H1, *tb_H1_lo, *tb_H2, *tb_H2_lo, *tb_embpre, *tb_embpre_lo, *a_t_all;
  float *ctext, *a_t;  // hoisted loop: loop-invariant text half [rows,128] and batch-shared time half [1024,128] of the embedding pre-activation
  float *H1_lo, *H2_lo, *embpre_lo, *cat_lo, *h1_lo, *c1_lo, *c2_lo;  // 3xTF32 residual planes of the step network's activations
  size_t bytes;
};

}  // namespace

struct lsdm_handle {
  lsdm_config cfg;
  std::vector<WEntry> entries;
  std::unordered_map<std::string, int> index;
  float* arena = nullptr;    // raw state-dict tensors
  float* derived = nullptr;  // folded / split tensors
  int64_t round_delta = 0;   // float offset from a weight in (arena|derived) to its TF32-rounded copy (tensor path)
  int64_t lo_delta = 0;      // ... and to its 3xTF32 residual plane rna(w - rna(w))
  int64_t arena_floats = 0, derived_floats = 0;
  bool finalized = false, have_sched = false, have_cond = false;
  // folded weights
  float *sa_w[4][3], *sa_b[4][3], *sa_wx[4], *sa_wf[4];
  float *fp_w[4][3], *fp_b[4][3], *fp_wa[4], *fp_wb[4];
  float *sa_wx_raw[4], *sa_wf_raw[4], *fp_wa_raw[4], *fp_wb_raw[4];  // un-folded first-layer splits (train-mode BatchNorm)
  bool fold_dirty = false;
  lsdm_allreduce_fn allreduce = nullptr;  // SyncBN hook: sums the train-mode BatchNorm statistics over data-parallel shards
  void* allreduce_ctx = nullptr;  // running statistics changed (train-mode forward): re-fold before the next eval-mode encode
  float *head_w, *head_b;
  std::vector<float> host_tail;  // [b2|b3|bh|conv2.w|conv2.b] of the fused backbone tail
  std::vector<float> host_fp1_b1;  // folded bias of fp1's first conv (kernel parameter of the fused fp1 + head kernel)
  int select_grid = 9;           // bit mask: ball query level 0 (1), level 1 (2), 3-NN of fp2 (4), of fp1 (8) through a per-cloud cell grid (identical selections)
  int loop_invariants = 15;      // lsdm_sample_loop, STRICT: what is computed once per call instead of once per step because its inputs do not
                                 //    change over the loop (same kernels, same inputs -> same bits): 1 condition MLPs + human decoder,
                                 //    2 text half of the embedding (the time half once per step for the whole batch: every sample shares t),
                                 //    4 sa1 + level-0 ball query in cloud order (the level-0 FPS only permutes its rows),
                                 //    8 guiding points (second x0-network pass) on the call's last step only (earlier ones are never visible)
  int time_batch = 1;            // 1: the time half of the split embedding is evaluated for TIME_BATCH steps per launch sequence
  int hoist_split = 1;           // 1: the hoisted loop computes the time half of the embedding once per step for the whole batch and the text half once per loop
  int sa1_compact = 1;           // 1: sa1 runs on the distinct rows of every ball-query group only (bit-identical, ~6x fewer tiles)
  int x0_fused = 1;              // 1: the x0 network of a step runs as one persistent kernel (x0net_fused.cu); 0: one GEMM per layer
  int dedup_absent = 1;          // 1: lsdm_sample_loop encodes ONE all-zero (absent, zero-padded) cloud per step and shares its
                                 //    backbone output with every other absent cloud (bit-identical: eval-mode clouds are independent)
  int n_active = 0;              // clouds the encoder runs on in the current lsdm_sample_loop call
  std::vector<float> host_wx[2], host_wf[2], host_b1[2], host_b2[2];  // host copies for the v2 fused SA kernels (kernel params)
  // schedule
  float* sched = nullptr;  // 5 x T
  int T = 0;
  Workspace ws{};
  bool have_ws = false;
  int64_t launches = 0;
  cudaStream_t side = nullptr, dense_st = nullptr, cond_st = nullptr, nn_st = nullptr;
  cudaEvent_t ev_cfork = nullptr, ev_cjoin = nullptr, ev_fps = nullptr, ev_nn = nullptr;  // (nn_st: the 3-NN chain beside the ball queries, both only need the FPS result)  // condition MLPs + human decoder beside the selection chain (cond_stream)
  int cond_stream = 1;  // 1: in the pipelined loop the per-sample condition MLPs / human decoder run on their own stream, overlapping the FPS chain
  cudaEvent_t ev_fork = nullptr, ev_sel[3] = {nullptr, nullptr, nullptr}, ev_dense[3] = {nullptr, nullptr, nullptr},
              ev_step[3] = {nullptr, nullptr, nullptr};
  // optional per-class CUDA-event profiler (bench.py's kernel shares / roofline numerator)
  int precision = 0;       // dense layers of the condition encoder (PointNet++): 0 fp32, 1 tf32, 2 3xtf32
  int precision_step = 0;
  int sa_fused = 1;  // 1 (tensor-core builds): fused set-abstraction kernels for levels 1-3; 0: gather + one GEMM per layer
  bool profiling = false;
  struct ProfRec { int cls; cudaEvent_t a, b; std::string tag; double flops; };
  std::string prof_report;
  std::vector<ProfRec> prof;
  double gemm_flops = 0.0;

  const float* W(const std::string& k) const { return arena + entries[index.at(k)].off; }
};

namespace {

void add_entry(lsdm_handle* h, const std::string& key, std::vector<int64_t> shape, bool ignored = false) {
  WEntry e;
  e.key = key;
  e.shape = shape;
  e.numel = 1;
  for (auto s : shape) e.numel *= s;
  e.loaded = false;
  if (ignored) {
    e.off = -1;
  } else {
    e.off = h->arena_floats;
    h->arena_floats += (e.numel + 63) & ~int64_t(63);  // 256-byte aligned slots
  }
  h->index[key] = (int)h->entries.size();
  h->entries.push_back(e);
}

void add_linear(lsdm_handle* h, const std::string& p, int64_t o, int64_t i) {
  add_entry(h, p + ".weight", {o, i});
  add_entry(h, p + ".bias", {o});
}
void add_bn(lsdm_handle* h, const std::string& p, int64_t c) {
  add_entry(h, p + ".weight", {c});
  add_entry(h, p + ".bias", {c});
  add_entry(h, p + ".running_mean", {c});
  add_entry(h, p + ".running_var", {c});
  add_entry(h, p + ".num_batches_tracked", {}, true);
}

// The reference's model_state_dict contract (model/sdm.py:19-129).
void build_registry(lsdm_handle* h) {
  const int64_t C = h->cfg.n_cats;
  add_entry(h, "sequence_pos_encoder.pe", {5000, 1, LAT});
  add_entry(h, "embed_timestep.sequence_pos_encoder.pe", {5000, 1, LAT});
  add_linear(h, "embed_timestep.time_embed.0", LAT, LAT);
  add_linear(h, "embed_timestep.time_embed.2", LAT, LAT);
  add_linear(h, "embed_text.0", CLIP / 2, CLIP);
  add_linear(h, "embed_text.2", 2 * LAT, CLIP / 2);
  add_linear(h, "embed_text.4", LAT, 2 * LAT);
  add_linear(h, "embed_cat.0", CATEMB, C);
  add_linear(h, "predict_cat.0", LAT / 2, LAT);
  add_linear(h, "predict_cat.2", LAT / 4, LAT / 2);
  add_linear(h, "predict_cat.4", C, LAT / 4);
  add_entry(h, "attn_layer.q_proj_weight", {LAT, LAT});
  add_entry(h, "attn_layer.k_proj_weight", {LAT, CATEMB});
  add_entry(h, "attn_layer.v_proj_weight", {LAT, NPTS * 3});  // dead in the reference forward; kept for strict loading
  add_entry(h, "attn_layer.in_proj_bias", {3 * LAT});
  add_linear(h, "attn_layer.out_proj", LAT, LAT);
  add_linear(h, "translation_layer.0", LAT, LAT + CATEMB);
  add_linear(h, "translation_layer.2", TRANS, LAT);
  add_linear(h, "point_wise_trans_layer.0", 3, TRANS + 3);
  add_entry(h, "pcd_attention.q_proj_weight", {TRANS, TRANS});
  add_entry(h, "pcd_attention.k_proj_weight", {TRANS, 3});
  add_entry(h, "pcd_attention.v_proj_weight", {TRANS, 3});
  add_entry(h, "pcd_attention.in_proj_bias", {3 * TRANS});
  add_linear(h, "pcd_attention.out_proj", TRANS, TRANS);
  for (const auto& s : kSA) {
    int last = s.cin;
    for (int i = 0; i < 3; ++i) {
      std::string p = std::string("pcd_backbone.") + s.name + ".mlp_convs." + std::to_string(i);
      add_entry(h, p + ".weight", {s.mlp[i], last, 1, 1});
      add_entry(h, p + ".bias", {s.mlp[i]});
      last = s.mlp[i];
    }
    for (int i = 0; i < 3; ++i) add_bn(h, std::string("pcd_backbone.") + s.name + ".mlp_bns." + std::to_string(i), s.mlp[i]);
  }
  for (const auto& s : kFP) {
    int last = s.cin;
    for (int i = 0; i < s.nl; ++i) {
      std::string p = std::string("pcd_backbone.") + s.name + ".mlp_convs." + std::to_string(i);
      add_entry(h, p + ".weight", {s.mlp[i], last, 1});
      add_entry(h, p + ".bias", {s.mlp[i]});
      last = s.mlp[i];
    }
    for (int i = 0; i < s.nl; ++i) add_bn(h, std::string("pcd_backbone.") + s.name + ".mlp_bns." + std::to_string(i), s.mlp[i]);
  }
  add_entry(h, "pcd_backbone.conv1.weight", {128, 128, 1});
  add_entry(h, "pcd_backbone.conv1.bias", {128});
  add_bn(h, "pcd_backbone.bn1", 128);
  add_entry(h, "pcd_backbone.conv2.weight", {3, 128, 1});
  add_entry(h, "pcd_backbone.conv2.bias", {3});
  const int hc[3][2] = {{64, 3}, {64, 64}, {64, 64}};
  for (int i = 0; i < 3; ++i) {
    std::string p = "human_backbone.de_spiral." + std::to_string(i);
    add_linear(h, p + ".conv.layer", hc[i][0], hc[i][1]);
    add_entry(h, p + ".norm.weight", {64});
    add_entry(h, p + ".norm.bias", {64});
  }
  add_linear(h, "human_backbone.de_spiral.3.layer", 3, 64);
  add_linear(h, "upsampling_layer.0", 128, 1);
  add_linear(h, "upsampling_layer.2", 512, 128);
  add_linear(h, "upsampling_layer.4", NPTS, 512);
  add_linear(h, "combine_extraction.0", LAT, 2 * LAT);
  add_linear(h, "input_process.pose_embedding.0", LAT / 2, 3);
  add_linear(h, "input_process.pose_embedding.2", LAT, LAT / 2);
  add_linear(h, "input_process.combination_extraction.0", 192, 2 * LAT);
  add_linear(h, "input_process.combination_extraction.2", LAT, 192);
  add_linear(h, "output_process.pose_final.0", LAT / 2, LAT);
  add_linear(h, "output_process.pose_final.2", 3, LAT / 2);
}

size_t carve(const lsdm_handle* h, void* base, Workspace* w) {
  const size_t B = (size_t)h->cfg.batch_local, C = B * NOBJ, nc = (size_t)h->cfg.n_cats;
  Bump a(base);
  w->human_scratch = a.take<float>(B * 2 * NPTS * 64 + B * 128);
  w->t_dev = a.take<int64_t>(B);
  const size_t np[5] = {1024, 1024, 256, 64, 16};
  const size_t fc[5] = {3, 64, 128, 256, 512};
  w->feat[0] = nullptr;
  w->cur = 0;
  const size_t fn[4] = {64, 256, 1024, 1024};  // fine-point count of fp4, fp3, fp2, fp1
  for (int l = 1; l <= 4; ++l) w->feat[l] = a.take<float>(C * np[l] * fc[l]);
  for (int k = 0; k < 3; ++k) {
    Workspace::Sel& s = w->sel[k];
    s.enc = a.take<float>(B * LAT);
    s.out_cat = a.take<float>(B * nc);
    s.attn_w = a.take<float>(B * NOBJ);
    s.tr = a.take<float>(C * TRANS);
    s.qq = a.take<float>(C * TRANS);
    s.hm = a.take<float>(B * NPTS * 3);
    s.fps_start = a.take<int64_t>(4 * C);
    s.xyz[0] = nullptr;
    for (int l = 1; l <= 4; ++l) {
      s.idx[l - 1] = a.take<int>(C * np[l]);
      s.grp[l - 1] = a.take<int>(C * np[l] * 32);
      s.xyz[l] = a.take<float>(C * np[l] * 3);
    }
    for (int l = 0; l < 4; ++l) {
      s.nn_idx[l] = a.take<int>(C * fn[l] * 3);
      s.nn_w[l] = a.take<float>(C * fn[l] * 3);
    }
    s.plan_rows = a.take<int>(C * 256 * 128);
    s.plan_used = a.take<int>(C * 256);
    s.plan_tiles = a.take<int>(C);
    s.plan_off = a.take<int>(C + 1);
    s.plan_n = a.take<int>(4);
  }
  w->clouds_c = a.take<float>(C * NPTS * 3);
  w->f1canon = a.take<float>(C * 1024 * 64);
  w->p2canon = a.take<float>(C * 1024 * 64);
  w->c_grp = a.take<int>(C * 1024 * 32);
  w->c_plan_rows = a.take<int>(C * 256 * 128);
  w->c_plan_used = a.take<int>(C * 256);
  w->c_plan_tiles = a.take<int>(C);
  w->c_plan_off = a.take<int>(C + 1);
  w->c_plan_n = a.take<int>(4);
  w->remap = a.take<int>(C);
  w->active = a.take<int>(C);
  w->n_active = a.take<int>(4);
  w->tA = a.take<float>(C * 2097152);  // 2M floats/cloud: the train-mode path materialises sa1's [32768,64] pre-pool layer
  w->bn_stats = a.take<double>(2 * 1024);
  w->tB = a.take<float>(C * 1048576);
  w->tP = a.take<float>(C * 262144);
  w->g3 = a.take<float>(C * 64 * 256);
  w->g2 = a.take<float>(C * 256 * 256);
  w->g1 = a.take<float>(C * 1024 * 128);
  w->backbone = a.take<float>(C * NPTS * 3);
  w->pa = a.take<float>(C * TRANS);
  w->pw = a.take<float>(C * NPTS * 3);
  for (int k = 0; k < 3; ++k) w->pcd_out[k] = a.take<float>(B * NPTS * 3);
  const size_t rows = B * NPTS;
  w->s256 = a.take<float>(B * 256);
  w->H1 = a.take<float>(B * 256 * 128);
  w->H1_lo = a.take<float>(B * 256 * 128);
  w->H2 = a.take<float>(B * 256 * 512);
  w->H2_lo = a.take<float>(B * 256 * 512);
  w->embpre = a.take<float>(rows * 256);
  w->embpre_lo = a.take<float>(rows * 256);
  w->cat = a.take<float>(2 * rows * 256);
  w->cat_lo = a.take<float>(2 * rows * 256);
  w->h1 = a.take<float>(2 * rows * 64);
  w->h1_lo = a.take<float>(2 * rows * 64);
  w->c1 = a.take<float>(2 * rows * 192);
  w->c1_lo = a.take<float>(2 * rows * 192);
  w->c2 = a.take<float>(2 * rows * 128);
  w->c2_lo = a.take<float>(2 * rows * 128);
  w->f1 = a.take<float>(2 * rows * 64);
  w->ctext = a.take<float>(rows * 128);
  w->a_t = a.take<float>((size_t)NPTS * 128);
  {
    const size_t G = 16;  // TIME_BATCH
    w->tb_t = a.take<int64_t>(G);
    w->tb_s256 = a.take<float>(G * 256);
    w->tb_H1 = a.take<float>(G * 256 * 128);
    w->tb_H1_lo = a.take<float>(G * 256 * 128);
    w->tb_H2 = a.take<float>(G * 256 * 512);
    w->tb_H2_lo = a.take<float>(G * 256 * 512);
    w->tb_embpre = a.take<float>(G * NPTS * 128);
    w->tb_embpre_lo = a.take<float>(G * NPTS * 128);
    w->a_t_all = a.take<float>(G * NPTS * 128);
  }
  w->x0 = a.take<float>(rows * 3);
  w->guiding = a.take<float>(rows * 3);
  w->loss_scratch = a.take<float>(64);
  a.take<char>(256);
  w->bytes = a.off;
  return a.off;
}

enum KClass { K_GEMM = 0, K_FPS, K_BALL, K_GATHER, K_3NN, K_FPCOMB, K_HEAD, K_COND, K_SCENE, K_DENOISE, K_OTHER, K_NCLASS };

template <typename F>
int prof_launch(lsdm_handle* h, cudaStream_t st, int cls, F&& f, const char* tag = nullptr, double flops = 0.0) {
  if (!h->profiling) {
    int r = f();
    if (r > 0) h->launches += r;
    return r;
  }
  lsdm_handle::ProfRec rec;
  rec.cls = cls;
  rec.tag = tag ? tag : "";
  rec.flops = flops;
  cudaEventCreate(&rec.a);
  cudaEventCreate(&rec.b);
  cudaEventRecord(rec.a, st);
  int r = f();
  cudaEventRecord(rec.b, st);
  h->prof.push_back(rec);
  if (r > 0) h->launches += r;
  return r;
}

extern "C" LSDM_API int lsdm_finalize_weights(lsdm_handle* h, void* stream);

enum GemmFlags { GF_A_ROUNDED = 1, GF_ROUND_OUT = 2 };

int gemm(lsdm_handle* h, cudaStream_t st, const float* A, int64_t lda, const float* W, int64_t ldw, float* C,
         int64_t ldc, const float* bias, int M, int N, int K, int act, int group_max = 0, int prec = -1, int flags = 0,
         const float* A_lo = nullptr, float* C_lo = nullptr) {
  GemmArgs g{};
  g.A = A; g.lda = lda; g.strideA = 0;
  g.W = W; g.ldw = ldw; g.strideW = 0;
  g.C = C; g.ldc = ldc; g.strideC = 0;
  g.bias = bias; g.bias_mode = bias ? 1 : 0;
  g.M = M; g.N = N; g.K = K; g.batch = 1;
  g.act = act; g.group_max = group_max; g.precision = prec >= 0 ? prec : h->precision;
  if (g.precision == 1) {  // plain TF32: activations are rounded by their producers, weights have a rounded copy
    g.a_rounded = (flags & GF_A_ROUNDED) ? 1 : 0;
    g.round_out = (flags & GF_ROUND_OUT) ? 1 : 0;
    if (g.a_rounded) {
      g.W = W + h->round_delta;
      g.w_rounded = 1;
    }
  }
  if (g.precision == 2 && A_lo != nullptr) {  // pre-split activations (hi = A, lo = A_lo) x pre-split weights, all fed by cp.async
    g.A_lo = A_lo;
    g.W = W + h->round_delta;
    g.W_lo = W + h->lo_delta;
    g.C_lo = C_lo;
  }
  char tag[64];
  snprintf(tag, sizeof(tag), "gemm p%d N%d K%d%s", g.precision, N, K, group_max ? " gmax" : "");
  int r = prof_launch(h, st, K_GEMM, [&] { return launch_gemm(g, st); }, tag, 2.0 * M * (double)N * K);
  if (r < 0) return fail(LSDM_EINVAL, "gemm: unsupported shape M=" + std::to_string(M) + " N=" + std::to_string(N) +
                                          " K=" + std::to_string(K));
  if (h->profiling) h->gemm_flops += 2.0 * M * (double)N * K;
  return LSDM_OK;
}
#define GE(x)                  \
  do {                         \
    int r__ = (x);             \
    if (r__ != LSDM_OK) return r__; \
  } while (0)

int check_ready(lsdm_handle* h, bool need_cond) {
  if (!h) return fail(LSDM_EINVAL, "null handle");
  if (!h->finalized) return fail(LSDM_ESTATE, "weights not finalised (lsdm_finalize_weights)");
  if (!h->have_ws) return fail(LSDM_ESTATE, "no workspace (lsdm_set_workspace)");
  if (need_cond && !h->have_cond) return fail(LSDM_ESTATE, "conditions not encoded (lsdm_encode_conditions)");
  return LSDM_OK;
}

// FPS (4 levels), ball queries (4 levels), 3-NN weights (4 levels): everything that depends only on coordinates.
// Also the small per-sample condition MLPs and the POSA human decoder, which are equally independent of x / t.
__global__ void gather_fps_start_kernel(const int64_t* __restrict__ src, const int* __restrict__ active, int n_src, int n_dst,
                                        int64_t* __restrict__ dst) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 4 * n_dst) dst[i] = src[(int64_t)(i / n_dst) * n_src + active[i % n_dst]];
}

// `objs` are the caller's [B,9,1024,3] clouds (human decoder input); `clouds` / `C` are what the PointNet++ selection chain
// runs on: the same tensor, or the de-duplicated copy (Workspace::clouds_c, `active` = compact -> original cloud index).
int select_phase(lsdm_handle* h, Workspace::Sel& q, const float* text, const float* objs, const float* cats,
                 const float* mask_global, const int64_t* fps_start, cudaStream_t st, const float* clouds = nullptr, int C = -1,
                 const int* active = nullptr, bool fork_cond = false, bool skip_cond = false, bool skip_level0 = false) {
  // skip_cond (lsdm_sample_loop, `loop_invariants` bit 0): the condition MLPs and the human decoder are deterministic functions
  // of inputs that do not change over the loop; step 0 computed them and the loop copied the results into this set
  const bool fork_streams = fork_cond;
  if (skip_cond) fork_cond = false;
  const int B = h->cfg.batch_local;
  if (!clouds) clouds = objs, C = B * NOBJ;
  Workspace& w = h->ws;
  // fork_cond (pipelined loop): the condition MLPs and the human decoder do not depend on the selection chain -- they run on their
  // own stream beside it (both are latency-bound, few warps per SM) and join before the phase's completion event is recorded
  cudaStream_t sel_st = st;
  if (fork_cond) {
    CK(cudaEventRecord(h->ev_cfork, st));  // carries the buffer-reuse waits of this phase over to the condition stream
    CK(cudaStreamWaitEvent(h->cond_st, h->ev_cfork, 0));
    st = h->cond_st;
  }
  if (!skip_cond) {
    CondWeights cw{h->W("embed_text.0.weight"), h->W("embed_text.0.bias"), h->W("embed_text.2.weight"), h->W("embed_text.2.bias"),
                   h->W("embed_text.4.weight"), h->W("embed_text.4.bias"), h->W("predict_cat.0.weight"), h->W("predict_cat.0.bias"),
                   h->W("predict_cat.2.weight"), h->W("predict_cat.2.bias"), h->W("predict_cat.4.weight"), h->W("predict_cat.4.bias"),
                   h->W("embed_cat.0.weight"), h->W("embed_cat.0.bias"), h->W("attn_layer.q_proj_weight"),
                   h->W("attn_layer.k_proj_weight"), h->W("attn_layer.in_proj_bias"), h->W("translation_layer.0.weight"),
                   h->W("translation_layer.0.bias"), h->W("translation_layer.2.weight"), h->W("translation_layer.2.bias"),
                   h->W("pcd_attention.q_proj_weight"), h->W("pcd_attention.in_proj_bias")};
    prof_launch(h, st, K_COND, [&] { return launch_cond(cw, text, cats, mask_global, B, h->cfg.batch_global, h->cfg.batch_offset,
                                                        h->cfg.n_cats, q.enc, q.out_cat, q.attn_w, q.tr, q.qq, st); });
    HumanWeights hw{};
    hw.w0 = h->W("human_backbone.de_spiral.0.conv.layer.weight"); hw.b0 = h->W("human_backbone.de_spiral.0.conv.layer.bias");
    hw.g0 = h->W("human_backbone.de_spiral.0.norm.weight"); hw.be0 = h->W("human_backbone.de_spiral.0.norm.bias");
    hw.w1 = h->W("human_backbone.de_spiral.1.conv.layer.weight"); hw.b1 = h->W("human_backbone.de_spiral.1.conv.layer.bias");
    hw.g1 = h->W("human_backbone.de_spiral.1.norm.weight"); hw.be1 = h->W("human_backbone.de_spiral.1.norm.bias");
    hw.w2 = h->W("human_backbone.de_spiral.2.conv.layer.weight"); hw.b2 = h->W("human_backbone.de_spiral.2.conv.layer.bias");
    hw.g2 = h->W("human_backbone.de_spiral.2.norm.weight"); hw.be2 = h->W("human_backbone.de_spiral.2.norm.bias");
    hw.w3 = h->W("human_backbone.de_spiral.3.layer.weight"); hw.b3 = h->W("human_backbone.de_spiral.3.layer.bias");
    prof_launch(h, st, K_COND, [&] { return launch_human(hw, objs, B, w.human_scratch, q.hm, st); });
  }
  if (fork_cond) {
    CK(cudaEventRecord(h->ev_cjoin, h->cond_st));
    st = sel_st;
  }
  if (active) {
    gather_fps_start_kernel<<<(4 * C + 255) / 256, 256, 0, st>>>(fps_start, active, B * NOBJ, C, q.fps_start);
  } else if (fps_start != q.fps_start) {
    CK(cudaMemcpyAsync(q.fps_start, fps_start, sizeof(int64_t) * 4 * C, cudaMemcpyDefault, st));
  }
  prof_launch(h, st, K_FPS, [&] { return launch_fps4(clouds, q.fps_start, C, q.idx[0], q.idx[1], q.idx[2], q.idx[3], q.xyz[1], q.xyz[2],
                                                     q.xyz[3], q.xyz[4], st); });
  const float* xyz[5] = {clouds, q.xyz[1], q.xyz[2], q.xyz[3], q.xyz[4]};
  static const bool fork_nn_env = !(getenv("LSDM_NN_STREAM") && atoi(getenv("LSDM_NN_STREAM")) == 0);
  const bool fork_nn = fork_streams && fork_nn_env;  // ball queries and 3-NN searches both depend on the FPS result only: two streams
  if (fork_nn) {
    CK(cudaEventRecord(h->ev_fps, st));
    CK(cudaStreamWaitEvent(h->nn_st, h->ev_fps, 0));
  }
  // skip_level0 (`loop_invariants` bit 2): sa1 keeps every point, so its groups were found once per call in cloud order
  const bool want_plan = !skip_level0 && h->sa1_compact && h->precision >= 1 && h->sa_fused > 0;
  bool plan_done = false;
  for (int l = skip_level0 ? 1 : 0; l < 4; ++l)
    prof_launch(h, st, K_BALL, [&] {
      if (l <= 1 && ((h->select_grid >> l) & 1) && !(l == 1 && skip_level0)) {  // levels 0 and 1: cell grid instead of the 1024 x 1024 / 256 x 1024 scan (identical groups)
        const bool with_plan = l == 0 && want_plan;  // the level-0 grid kernel also writes the plan of sa1's distinct rows
        const int r = launch_ball_query_grid(xyz[l], xyz[l + 1], C, kSA[l].N, kSA[l].npoint, kSA[l].radius, q.grp[l], st,
                                             with_plan ? q.plan_rows : nullptr, q.plan_used, q.plan_tiles);
        if (r > 0) {
          if (with_plan) plan_done = true;
          return r;
        }
      }
      // (skip_level0: level 1 stays in cloud order over the loop -- its groups are stored as cloud indices, idx[0][position])
      return launch_ball_query(xyz[l], xyz[l + 1], C, kSA[l].N, kSA[l].npoint, kSA[l].radius, q.grp[l], st,
                               (l == 1 && skip_level0) ? q.idx[0] : nullptr);
    });
  if (want_plan)
    prof_launch(h, st, K_BALL, [&] {
      if (plan_done) return launch_sa1_plan_scan(q.plan_tiles, C, q.plan_off, q.plan_n, st);
      return launch_sa1_plan(q.grp[0], C, q.plan_rows, q.plan_used, q.plan_tiles, q.plan_off, q.plan_n, st);
    });
  const int fine[4] = {3, 2, 1, 0}, coarse[4] = {4, 3, 2, 1};
  const int fineN[4] = {64, 256, 1024, 1024}, coarseN[4] = {16, 64, 256, 1024};
  cudaStream_t nst = fork_nn ? h->nn_st : st;
  for (int l = 3; l >= 0; --l)  // (largest search first when it has its own stream)
    prof_launch(h, nst, K_3NN, [&] {
      if (l >= 2 && ((h->select_grid >> l) & 1)) {  // fp2 (1024 <- 256) and fp1 (1024 <- 1024)
        const int r = launch_three_nn_grid(xyz[fine[l]], xyz[coarse[l]], C, fineN[l], coarseN[l], q.nn_idx[l], q.nn_w[l], nst);
        if (r > 0) return r;
      }
      return launch_three_nn(xyz[fine[l]], xyz[coarse[l]], C, fineN[l], coarseN[l], q.nn_idx[l], q.nn_w[l], nst);
    });
  if (fork_nn) {
    CK(cudaEventRecord(h->ev_nn, h->nn_st));
    CK(cudaStreamWaitEvent(st, h->ev_nn, 0));
  }
  if (fork_cond) CK(cudaStreamWaitEvent(st, h->ev_cjoin, 0));
  CK(cudaPeekAtLastError());
  return LSDM_OK;
}

// sa1 + level-0 ball query in cloud order (centroids = the cloud's own points), once per lsdm_sample_loop call.
int sa1_cloud_order(lsdm_handle* h, const float* clouds, int C, cudaStream_t st) {
  Workspace& w = h->ws;
  int r = prof_launch(h, st, K_BALL, [&] {
    return launch_ball_query_grid(clouds, clouds, C, kSA[0].N, kSA[0].npoint, kSA[0].radius, w.c_grp, st, w.c_plan_rows, w.c_plan_used, w.c_plan_tiles);
  });
  if (r <= 0) return fail(LSDM_EINVAL, "cell-grid ball query unavailable");
  prof_launch(h, st, K_BALL, [&] { return launch_sa1_plan_scan(w.c_plan_tiles, C, w.c_plan_off, w.c_plan_n, st); });
  r = prof_launch(h, st, K_GEMM, [&] {
    return launch_sa1_compact(clouds, clouds, w.c_plan_rows, w.c_plan_used, w.c_plan_off, w.c_plan_n, h->host_wx[0].data(), h->host_wf[0].data(),
                              h->host_b1[0].data(), h->host_b2[0].data(), h->sa_w[0][1], h->sa_w[0][2], h->sa_b[0][2], C, w.f1canon,
                              h->precision == 1, st);
  }, "sa_fused sa1 (once per call)", 2.0 * C * 1024 * 32 * ((double)kSA[0].mlp[0] * kSA[0].mlp[1] + (double)kSA[0].mlp[1] * kSA[0].mlp[2]));
  if (r < 0) return fail(LSDM_EINVAL, "fused sa1 kernel unavailable");
  // sa2's first conv, feature half, once per source point -- in cloud order too
  GE(gemm(h, st, w.f1canon, kSA[1].cin - 3, h->sa_wf[1], kSA[1].cin - 3, w.p2canon, kSA[1].mlp[0], h->sa_b[1][0], C * 1024, kSA[1].mlp[0],
          kSA[1].cin - 3, ACT_NONE, 0, -1, GF_A_ROUNDED));
  CK(cudaPeekAtLastError());
  return LSDM_OK;
}

// Dense layers of PointNet++ given the selection results.
int dense_phase(lsdm_handle* h, const Workspace::Sel& q, const float* clouds, cudaStream_t st, int C, bool sa1_canon = false) {
  Workspace& w = h->ws;
  const float* xyz[5] = {clouds, q.xyz[1], q.xyz[2], q.xyz[3], q.xyz[4]};
  const float* feat[5] = {clouds, w.feat[1], w.feat[2], w.feat[3], w.feat[4]};
  for (int l = 0; l < 4; ++l) {
    const SASpec& s = kSA[l];
    const int N = s.N, S = s.npoint, C1 = s.mlp[0], C2 = s.mlp[1], C3 = s.mlp[2];
    const float* P = nullptr;
    // sa1_canon: level 1 (sa1's output features, sa2's projected rows P = W_f f + b, the coordinates) was computed once per call in
    // CLOUD order; this step's level-1 ball-query groups hold cloud indices and fp2 reads its fine rows through the level-0 FPS
    // indices, so nothing is permuted or recomputed here
    if (l == 0 && sa1_canon) continue;
    if (l == 1 && sa1_canon) {
      P = w.p2canon;
    } else if (l > 0) {  // first conv, feature half, once per source point
      GE(gemm(h, st, feat[l], s.cin - 3, h->sa_wf[l], s.cin - 3, w.tP, C1, h->sa_b[l][0], C * N, C1, s.cin - 3, ACT_NONE, 0, -1, GF_A_ROUNDED));
      P = w.tP;
    }
    if (h->precision >= 1 && h->sa_fused > 0 && l <= 2) {
      int r = prof_launch(h, st, K_GEMM, [&] {
        if (l == 0 && h->sa1_compact)
          return launch_sa1_compact(xyz[0], xyz[1], q.plan_rows, q.plan_used, q.plan_off, q.plan_n, h->host_wx[0].data(), h->host_wf[0].data(),
                                    h->host_b1[0].data(), h->host_b2[0].data(), h->sa_w[0][1], h->sa_w[0][2], h->sa_b[0][2], C, w.feat[1],
                                    h->precision == 1, st);
        if (l <= 1)
          return launch_sa_fused_v2(l, P, (l == 1 && sa1_canon) ? clouds : xyz[l], xyz[l + 1], q.grp[l], h->host_wx[l].data(), h->host_wf[l].data(), h->host_b1[l].data(),
                                    h->host_b2[l].data(), h->sa_wx[l], h->sa_b[l][0], h->sa_w[l][1], h->sa_w[l][2], h->sa_b[l][2], C, N, S, w.feat[l + 1],
                                    h->precision == 1, st);
        return launch_sa_fused(l, P, xyz[l], xyz[l + 1], q.grp[l], h->sa_wx[l], h->sa_wf[l], h->sa_b[l][0],
                               h->sa_w[l][1], h->sa_b[l][1], h->sa_w[l][2], h->sa_b[l][2], C, N, S, w.feat[l + 1], h->precision == 1, st);
      }, l == 0 ? "sa_fused sa1" : (l == 1 ? "sa_fused sa2" : "sa_fused sa3"), 2.0 * C * S * 32 * ((double)C1 * C2 + (double)C2 * C3));
      if (r < 0) return fail(LSDM_EINVAL, "fused SA kernel unavailable for this level");
      if (h->profiling) h->gemm_flops += 2.0 * C * S * 32 * ((double)C1 * C2 + (double)C2 * C3);
      continue;
    }
    prof_launch(h, st, K_GATHER, [&] { return launch_sa_gather(P, h->sa_wx[l], h->sa_wf[l], h->sa_b[l][0], xyz[l], xyz[l + 1], q.grp[l], C, N, S, C1,
                                    w.tA, h->precision == 1, st); });
    const int rows = C * S * 32;
    GE(gemm(h, st, w.tA, C1, h->sa_w[l][1], C1, w.tB, C2, h->sa_b[l][1], rows, C2, C1, ACT_RELU, 0, -1, GF_A_ROUNDED | GF_ROUND_OUT));
    GE(gemm(h, st, w.tB, C2, h->sa_w[l][2], C2, w.feat[l + 1], C3, h->sa_b[l][2], rows, C3, C2, ACT_RELU, 1, -1, GF_A_ROUNDED | GF_ROUND_OUT));
  }
  // feature propagation: fine level <- coarse level
  const int fine[4] = {3, 2, 1, 0}, coarse[4] = {4, 3, 2, 1};
  const int fineN[4] = {64, 256, 1024, 1024}, coarseN[4] = {16, 64, 256, 1024};
  const float* coarse_feat = w.feat[4];
  float* outs[4] = {w.g3, w.g2, w.g1, nullptr};
  for (int l = 0; l < 4; ++l) {
    const FPSpec& s = kFP[l];
    const int N = fineN[l], S = coarseN[l], C1 = s.mlp[0];
    const float* Pa = nullptr;
    const bool fused_level = h->precision >= 1 && l == 2;  // fp2: both weight matrices fit in shared memory
    if (s.Ca > 0 && !fused_level) {
      GE(gemm(h, st, feat[fine[l]], s.Ca, h->fp_wa[l], s.Ca, w.tA, C1, h->fp_b[l][0], C * N, C1, s.Ca, ACT_NONE, 0, -1, GF_A_ROUNDED));
      Pa = w.tA;
    }
    GE(gemm(h, st, coarse_feat, s.Cb, h->fp_wb[l], s.Cb, w.tB, C1, nullptr, C * S, C1, s.Cb, ACT_NONE, 0, -1, GF_A_ROUNDED));
    if (fused_level) {
      const double fl = 2.0 * C * N * ((double)s.Ca * C1 + (double)C1 * s.mlp[1]);
      int r = prof_launch(h, st, K_GEMM, [&] {
        return launch_fp_fused(sa1_canon ? w.f1canon : feat[fine[l]], s.Ca, h->fp_wa[l], h->fp_b[l][0], w.tB, q.nn_idx[l], q.nn_w[l],
                               h->fp_w[l][1], h->fp_b[l][1], C, N, S, C1, s.mlp[1], outs[l], h->precision == 1, st,
                               sa1_canon ? q.idx[0] : nullptr);
      }, "fp2_fused", fl);
      if (r < 0) return fail(LSDM_EINVAL, "fused FP kernel unavailable for this level");
      if (h->profiling) h->gemm_flops += fl;
      coarse_feat = outs[l];
      continue;
    }
    if (l == 3 && h->precision >= 1) {  // fp1 + head: interpolation gathered inside the fused kernel
      const double fl = 2.0 * C * N * 3.0 * 128 * 128;
      int r = prof_launch(h, st, K_GEMM, [&] {
        return launch_fp1_fused(w.tB, q.nn_idx[l], q.nn_w[l], h->host_fp1_b1.data(), h->fp_w[l][1], h->fp_w[l][2], h->head_w,
                                h->host_tail.data(), C, N, S, w.backbone, st);
      }, "fp1_fused", fl);
      if (r < 0) return fail(LSDM_EINVAL, "fused fp1 kernel unavailable");
      if (h->profiling) h->gemm_flops += fl;
      continue;
    }
    prof_launch(h, st, K_FPCOMB, [&] { return launch_fp_combine(Pa, h->fp_b[l][0], w.tB, q.nn_idx[l], q.nn_w[l], C, N, S, C1, w.tP, h->precision == 1, st); });
    if (l < 3) {
      GE(gemm(h, st, w.tP, C1, h->fp_w[l][1], C1, outs[l], s.mlp[1], h->fp_b[l][1], C * N, s.mlp[1], C1, ACT_RELU, 0, -1, GF_A_ROUNDED | GF_ROUND_OUT));
      coarse_feat = outs[l];
    } else {
      GE(gemm(h, st, w.tP, 128, h->fp_w[l][1], 128, w.tA, 128, h->fp_b[l][1], C * N, 128, 128, ACT_RELU, 0, -1, GF_A_ROUNDED | GF_ROUND_OUT));
      GE(gemm(h, st, w.tA, 128, h->fp_w[l][2], 128, w.tP, 128, h->fp_b[l][2], C * N, 128, 128, ACT_RELU, 0, -1, GF_A_ROUNDED | GF_ROUND_OUT));
      GE(gemm(h, st, w.tP, 128, h->head_w, 128, w.tA, 128, h->head_b, C * N, 128, 128, ACT_RELU, 0, -1, GF_A_ROUNDED));
      prof_launch(h, st, K_HEAD, [&] { return launch_head3(w.tA, h->W("pcd_backbone.conv2.weight"), h->W("pcd_backbone.conv2.bias"),
                                  (int64_t)C * N, w.backbone, st); });
    }
  }
  return LSDM_OK;
}

// emb[row, :] = gelu(A_t[row % 1024, :] + C_text[row, :]) as hi / lo TF32 planes with leading dimension 256 (thread = 4 channels)
__global__ void emb_combine_kernel(const float* __restrict__ a_t, const float* __restrict__ ctext, int64_t rows, float* __restrict__ hi,
                                   float* __restrict__ lo) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * 32) return;
  const int64_t row = i >> 5;
  const int q = (int)(i & 31);
  const float4 a = *reinterpret_cast<const float4*>(a_t + (row & (NPTS - 1)) * 128 + q * 4);
  const float4 c = *reinterpret_cast<const float4*>(ctext + row * 128 + q * 4);
  float v[4] = {gelu_erf(a.x + c.x), gelu_erf(a.y + c.y), gelu_erf(a.z + c.z), gelu_erf(a.w + c.w)}, l[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) split_tf32(v[j], v[j], l[j]);
  *reinterpret_cast<float4*>(hi + row * 256 + q * 4) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4*>(lo + row * 256 + q * 4) = make_float4(l[0], l[1], l[2], l[3]);
}

// The embedding chain emb[b,p,:] = act(Wc[:, s0:s0+ns] . u(b,p) + bias) for `nb` samples: timestep embedding -> [ts || enc] scalars ->
// upsampler 1 -> 128 -> 512 -> 1024 points -> combine (model/sdm.py:108-122,164-167,208), over the scalars [s0, s0 + ns) only.
struct UpsampleBufs { const int64_t* t; const float* enc; float *s256, *H1, *H1_lo, *H2, *H2_lo, *embpre, *embpre_lo; };
int upsample_chain(lsdm_handle* h, cudaStream_t st, const UpsampleBufs& u, bool sp, int ps, int nb, int s0, int ns, const float* combine_bias,
                   int combine_act, float* out, int64_t ld_out, float* out_lo) {
  prof_launch(h, st, K_COND, [&] { return launch_time_embed(h->W("embed_timestep.sequence_pos_encoder.pe"), h->W("embed_timestep.time_embed.0.weight"),
                                   h->W("embed_timestep.time_embed.0.bias"), h->W("embed_timestep.time_embed.2.weight"),
                                   h->W("embed_timestep.time_embed.2.bias"), u.t, u.enc, h->W("upsampling_layer.0.weight"),
                                   h->W("upsampling_layer.0.bias"), nb, u.s256, u.H1, sp ? u.H1_lo : nullptr, st); });
  GE(gemm(h, st, u.H1, 128, h->W("upsampling_layer.2.weight"), 128, u.H2, 512, h->W("upsampling_layer.2.bias"), nb * 256, 512,
          128, ACT_GELU, 0, ps, 0, sp ? u.H1_lo : nullptr, sp ? u.H2_lo : nullptr));
  {
    // embpre[b][p][s] = gelu(sum_k U4[p][k] H2[b][s0 + s][k] + b4[p]): the upsampler's last layer written point-major
    GemmArgs g{};
    g.A = h->W("upsampling_layer.4.weight"); g.lda = 512; g.strideA = 0;
    g.W = u.H2 + (size_t)s0 * 512; g.ldw = 512; g.strideW = 256 * 512;
    g.C = u.embpre; g.ldc = ns; g.strideC = (int64_t)NPTS * ns;
    g.bias = h->W("upsampling_layer.4.bias"); g.bias_mode = 2;
    g.M = NPTS; g.N = ns; g.K = 512; g.batch = nb; g.act = ACT_GELU; g.group_max = 0; g.precision = ps;
    if (sp) {
      g.A_lo = g.A + h->lo_delta;
      g.A = g.A + h->round_delta;
      g.W_lo = u.H2_lo + (size_t)s0 * 512;
      g.C_lo = u.embpre_lo;
    }
    int r = prof_launch(h, st, K_GEMM, [&] { return launch_gemm(g, st); }, "gemm p2 upsampler K512 batched",
                        2.0 * g.M * (double)g.N * g.K * g.batch);
    if (r < 0) return fail(LSDM_EINVAL, "upsampler gemm");
    if (h->profiling) h->gemm_flops += 2.0 * g.M * (double)g.N * g.K * g.batch;
  }
  GE(gemm(h, st, u.embpre, ns, h->W("combine_extraction.0.weight") + s0, 256, out, ld_out, combine_bias, nb * NPTS, 128, ns, combine_act, 0, ps, 0,
          sp ? u.embpre_lo : nullptr, out_lo));
  return LSDM_OK;
}

// Time half of the split embedding for the next `n` steps of a sampling loop (t_first, t_first - 1, ...): every sample shares t, so a
// step's A_t[p,:] = Wc[:, :128] . u_ts(p) is one "sample" of the chain; TIME_BATCH steps are evaluated per launch sequence instead of
// three latency-bound launches per step on the stream that is serial in x.  Row arithmetic does not depend on the batch: same bits.
constexpr int TIME_BATCH = 16;
__global__ void fill_t_run_kernel(int64_t* t, int n, int64_t t_first) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) t[i] = t_first - i;
}
int time_half_batch(lsdm_handle* h, int64_t t_first, int n, cudaStream_t st) {
  Workspace& w = h->ws;
  prof_launch(h, st, K_OTHER, [&] {
    fill_t_run_kernel<<<1, TIME_BATCH, 0, st>>>(w.tb_t, n, t_first);
    return 1;
  });
  UpsampleBufs u{w.tb_t, nullptr, w.tb_s256, w.tb_H1, w.tb_H1_lo, w.tb_H2, w.tb_H2_lo, w.tb_embpre, w.tb_embpre_lo};
  return upsample_chain(h, st, u, true, h->precision_step, n, 0, 128, nullptr, ACT_NONE, w.a_t_all, 128, nullptr);
}

// x/t-dependent part: timestep embedding, upsampler, x += pcd_out, Input/OutputProcess, optional posterior.
int step_core(lsdm_handle* h, float* x, const int64_t* t, const float* noise, float* sample_out, float* x0_out,
              float* guiding_out, bool want_guiding, int clip, cudaStream_t st, int si = -1, int hoist_split = 0,
              const float* a_t_pre = nullptr) {
  // hoist_split (hoisted sampling loop only, every sample shares t): 1 = first step (also builds the loop-invariant text half of
  // the embedding), 2 = later steps
  Workspace& w = h->ws;
  if (si < 0) si = w.cur;
  const int B = h->cfg.batch_local;
  const int rows = B * NPTS;
  if (t != w.t_dev) CK(cudaMemcpyAsync(w.t_dev, t, sizeof(int64_t) * B, cudaMemcpyDefault, st));
  // 3xTF32 step network with pre-split planes: every activation of the chain is stored as hi + lo by its producer, so the
  // GEMMs stage all four operand planes with cp.async (no per-tile splitting in registers)
  const bool sp = h->precision_step == 2 && g_gemm_async != 0;
  const int ps = h->precision_step;
  // ---- the step's embedding emb[b,p,:] = gelu(Wc . [u_ts(p) || u_text(b,p)] + bc)  (model/sdm.py:164-167,208) ----
  // `nb` samples, scalars [s0, s0 + ns) of [ts || enc], combine columns [s0, s0 + ns): the full chain is (B, 0, 256); the
  // hoisted loop splits it into a batch-shared time half (1, 0, 128) and a loop-invariant text half (B, 128, 128).
  const UpsampleBufs ub{w.t_dev, w.sel[si].enc, w.s256, w.H1, w.H1_lo, w.H2, w.H2_lo, w.embpre, w.embpre_lo};
  auto upsample = [&](int nb, int s0, int ns, const float* combine_bias, int combine_act, float* out, int64_t ld_out, float* out_lo) -> int {
    return upsample_chain(h, st, ub, sp, ps, nb, s0, ns, combine_bias, combine_act, out, ld_out, out_lo);
  };
  if (hoist_split && sp) {
    if (hoist_split == 1)   // once per loop: C_text[b,p,:] = Wc[:, 128:] . u_text(b,p) + bc
      GE(upsample(B, 128, 128, h->W("combine_extraction.0.bias"), ACT_NONE, w.ctext, 128, nullptr));
    // every step: A_t[p,:] = Wc[:, :128] . u_ts(p) for the ONE timestep all samples of the loop share, then emb = gelu(A_t + C_text)
    // (a_t_pre: the loop evaluated it for a run of steps at once, time_half_batch)
    if (!a_t_pre) {
      GE(upsample(1, 0, 128, nullptr, ACT_NONE, w.a_t, 128, nullptr));
      a_t_pre = w.a_t;
    }
    prof_launch(h, st, K_DENOISE, [&] {
      const int64_t n = (int64_t)rows * 32;
      emb_combine_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(a_t_pre, w.ctext, rows, w.cat + 128, w.cat_lo + 128);
      return 1;
    });
  } else {
    GE(upsample(B, 0, 256, h->W("combine_extraction.0.bias"), ACT_GELU, w.cat + 128, 256, sp ? w.cat_lo + 128 : nullptr));
  }
  if (sp && h->x0_fused && (rows & 127) == 0) {
    // the whole x0 network (both passes), `x += pcd_out`, the posterior and the ancestral noise in one persistent kernel
    auto hi = [&](const char* k) { return h->W(k) + h->round_delta; };
    auto lo = [&](const char* k) { return h->W(k) + h->lo_delta; };
    const int n_pass = want_guiding ? 2 : 1;
    const double fl = 2.0 * rows * n_pass * (64.0 * 128 + 256.0 * 192 + 192.0 * 128 + 128.0 * 64);
    float* gd = want_guiding ? (guiding_out ? guiding_out : w.guiding) : nullptr;
    int r = prof_launch(h, st, K_GEMM, [&] {
      return launch_x0net_fused(x, w.pcd_out[si], noise, sample_out, x0_out ? x0_out : w.x0, gd, w.t_dev, h->sched,
                                h->sched ? h->sched + h->T : nullptr, h->sched ? h->sched + 2 * h->T : nullptr, rows, n_pass, clip,
                                w.cat + 128, w.cat_lo + 128, 256, hi("input_process.pose_embedding.2.weight"),
                                lo("input_process.pose_embedding.2.weight"), hi("input_process.combination_extraction.0.weight"),
                                lo("input_process.combination_extraction.0.weight"), hi("input_process.combination_extraction.2.weight"),
                                lo("input_process.combination_extraction.2.weight"), hi("output_process.pose_final.0.weight"),
                                lo("output_process.pose_final.0.weight"), h->W("input_process.pose_embedding.0.weight"),
                                h->W("input_process.pose_embedding.0.bias"), h->W("input_process.pose_embedding.2.bias"),
                                h->W("input_process.combination_extraction.0.bias"), h->W("input_process.combination_extraction.2.bias"),
                                h->W("output_process.pose_final.0.bias"), h->W("output_process.pose_final.2.weight"),
                                h->W("output_process.pose_final.2.bias"), st);
    }, "x0net_fused", fl);
    if (r > 0) {
      if (h->profiling) h->gemm_flops += fl;
      CK(cudaPeekAtLastError());
      return LSDM_OK;
    }
    // TMA descriptors unavailable: the layer-by-layer chain below
  }
  const int M = want_guiding ? 2 * rows : rows;
  if (want_guiding) {
    CK(cudaMemcpy2DAsync(w.cat + (size_t)rows * 256 + 128, 256 * sizeof(float), w.cat + 128, 256 * sizeof(float),
                         128 * sizeof(float), rows, cudaMemcpyDeviceToDevice, st));
    if (sp)
      CK(cudaMemcpy2DAsync(w.cat_lo + (size_t)rows * 256 + 128, 256 * sizeof(float), w.cat_lo + 128, 256 * sizeof(float),
                           128 * sizeof(float), rows, cudaMemcpyDeviceToDevice, st));
  }
  prof_launch(h, st, K_DENOISE, [&] { return launch_pose_embed0(x, w.pcd_out[si], h->W("input_process.pose_embedding.0.weight"),
                                    h->W("input_process.pose_embedding.0.bias"), rows, w.h1, sp ? w.h1_lo : nullptr, st); });
  if (want_guiding)
    prof_launch(h, st, K_DENOISE, [&] { return launch_pose_embed0(w.pcd_out[si], nullptr, h->W("input_process.pose_embedding.0.weight"),
                                      h->W("input_process.pose_embedding.0.bias"), rows, w.h1 + (size_t)rows * 64,
                                      sp ? w.h1_lo + (size_t)rows * 64 : nullptr, st); });
  GE(gemm(h, st, w.h1, 64, h->W("input_process.pose_embedding.2.weight"), 64, w.cat, 256,
          h->W("input_process.pose_embedding.2.bias"), M, 128, 64, ACT_SIGMOID, 0, ps, 0, sp ? w.h1_lo : nullptr, sp ? w.cat_lo : nullptr));
  GE(gemm(h, st, w.cat, 256, h->W("input_process.combination_extraction.0.weight"), 256, w.c1, 192,
          h->W("input_process.combination_extraction.0.bias"), M, 192, 256, ACT_SIGMOID, 0, ps, 0, sp ? w.cat_lo : nullptr, sp ? w.c1_lo : nullptr));
  GE(gemm(h, st, w.c1, 192, h->W("input_process.combination_extraction.2.weight"), 192, w.c2, 128,
          h->W("input_process.combination_extraction.2.bias"), M, 128, 192, ACT_SIGMOID, 0, ps, 0, sp ? w.c1_lo : nullptr, sp ? w.c2_lo : nullptr));
  GE(gemm(h, st, w.c2, 128, h->W("output_process.pose_final.0.weight"), 128, w.f1, 64,
          h->W("output_process.pose_final.0.bias"), M, 64, 128, ACT_GELU, 0, ps, 0, sp ? w.c2_lo : nullptr, nullptr));
  float* x0 = x0_out ? x0_out : w.x0;
  prof_launch(h, st, K_DENOISE, [&] { return launch_final3(w.f1, h->W("output_process.pose_final.2.weight"), h->W("output_process.pose_final.2.bias"), rows,
                               x0, x, w.t_dev, h->sched, h->sched ? h->sched + h->T : nullptr, h->sched ? h->sched + 2 * h->T : nullptr, noise,
                               sample_out, clip, st); });
  if (want_guiding) {
    float* gd = guiding_out ? guiding_out : w.guiding;
    prof_launch(h, st, K_DENOISE, [&] { return launch_final3(w.f1 + (size_t)rows * 64, h->W("output_process.pose_final.2.weight"),
                                 h->W("output_process.pose_final.2.bias"), rows, gd, nullptr, nullptr, nullptr, nullptr,
                                 nullptr, nullptr, nullptr, 0, st); });
  }
  CK(cudaPeekAtLastError());
  return LSDM_OK;
}

__global__ void round_tf32_kernel(const float* __restrict__ src, float* __restrict__ dst, float* __restrict__ dst_lo, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    uint32_t r, l;
    const float x = src[i];
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    dst[i] = __uint_as_float(r);
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(l) : "f"(x - __uint_as_float(r)));
    dst_lo[i] = __uint_as_float(l);
  }
}

__global__ void add_planes_kernel(const float* __restrict__ hi, const float* __restrict__ lo, float* __restrict__ dst, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = hi[i] + lo[i];
}

// Absent objects are zero-padded by the dataset (reference posa/dataset.py:456): a cloud whose 3072 coordinates are all 0.0.
// In eval mode every cloud goes through PointNet++ independently and the output for such a cloud does not depend on its
// FPS start draws (every point, hence every gathered row, is the same), so all absent clouds of a batch share ONE result.
// (A -0.0 coordinate compares equal to +0.0 in every selection and can only flip the sign of an activation that is
// exactly zero: results are equal as values -- torch.equal -- which the tests and bench.py check against the full run.)
__global__ void classify_clouds_kernel(const float* __restrict__ clouds, int* __restrict__ absent) {
  const uint4* p = reinterpret_cast<const uint4*>(clouds + (int64_t)blockIdx.x * NPTS * 3);
  unsigned acc = 0;
  for (int i = threadIdx.x; i < NPTS * 3 / 4; i += blockDim.x) {
    const uint4 v = p[i];
    acc |= v.x | v.y | v.z | v.w;
  }
  acc &= 0x7fffffffu;  // value zero: +0.0 and -0.0 (an all-zero cloud multiplied by a 0 mask has both)
  const int any = __syncthreads_or(acc != 0);
  if (threadIdx.x == 0) absent[blockIdx.x] = any ? 0 : 1;
}
// active = [every present cloud in order, then the first absent cloud]; remap[c] = position of c's result in that list.
// `remap` holds the absent flags on entry.  One thread: C <= 9216, once per lsdm_sample_loop call.
__global__ void compact_clouds_kernel(int* __restrict__ remap, int* __restrict__ active, int* __restrict__ n_active, int C) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  int n = 0, first_absent = -1;
  for (int c = 0; c < C; ++c) {
    if (remap[c]) {
      if (first_absent < 0) first_absent = c;
    } else {
      active[n++] = c;
    }
  }
  const int rep = n;
  if (first_absent >= 0) active[n++] = first_absent;
  int k = 0;
  for (int c = 0; c < C; ++c) {
    if (remap[c]) remap[c] = rep;
    else remap[c] = k++;
  }
  *n_active = n;
}
__global__ void gather_clouds_kernel(const float* __restrict__ clouds, const int* __restrict__ active, float* __restrict__ out) {
  const float4* src = reinterpret_cast<const float4*>(clouds + (int64_t)active[blockIdx.x] * NPTS * 3);
  float4* dst = reinterpret_cast<float4*>(out + (int64_t)blockIdx.x * NPTS * 3);
  for (int i = threadIdx.x; i < NPTS * 3 / 4; i += blockDim.x) dst[i] = src[i];
}

__global__ void fill_t_kernel(int64_t* t, int n, int64_t v) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) t[i] = v;
}

}  // namespace

extern "C" {

#ifndef LSDM_SRC_HASH
#define LSDM_SRC_HASH "unknown"
#endif
// "... src:<first 16 hex digits of sha256(cat of csrc/*.cu, *.cuh and include/lsdm_b200.h in Makefile order)>"
LSDM_API const char* lsdm_version(void) { return "lsdm_b200 0.2 sm_100a src:" LSDM_SRC_HASH; }
LSDM_API const char* lsdm_last_error(void) { return g_err.c_str(); }

LSDM_API int lsdm_create(lsdm_handle** out, const lsdm_config* cfg) {
  if (!out || !cfg) return fail(LSDM_EINVAL, "null argument");
  if (cfg->batch_local <= 0 || cfg->batch_global < cfg->batch_local || cfg->batch_offset < 0 ||
      cfg->batch_offset + cfg->batch_local > cfg->batch_global)
    return fail(LSDM_EINVAL, "bad batch configuration");
  if (cfg->n_cats <= 0 || cfg->n_cats > 32) return fail(LSDM_EINVAL, "n_cats must be in 1..32");
  int prev_dev = 0;
  CK(cudaGetDevice(&prev_dev));
  CK(cudaSetDevice(cfg->device));
  struct Restore { int d; ~Restore() { cudaSetDevice(d); } } restore{prev_dev};  // the caller's current device is left as it was
  lsdm_handle* h = new lsdm_handle();
  h->cfg = *cfg;
  build_registry(h);
  // derived (folded) weights: generous upper bound = all backbone conv weights + biases again
  h->derived_floats = 2700000;
  cudaError_t e = cudaMalloc(&h->arena, sizeof(float) * 3 * (h->arena_floats + h->derived_floats));
  h->round_delta = h->arena_floats + h->derived_floats;
  h->lo_delta = 2 * h->round_delta;
  if (e != cudaSuccess) {
    delete h;
    return fail(LSDM_ENOMEM, std::string("cudaMalloc weights: ") + cudaGetErrorString(e));
  }
  h->derived = h->arena + h->arena_floats;
  // The selection chain (side stream) is a long dependent chain of small kernels (1360 FPS rounds per cloud): it gets the
  // greatest stream priority so that its CTAs are scheduled ahead of the persistent dense kernels' (LSDM_SIDE_PRIO=0: off).
  int prio_lo = 0, prio_hi = 0;
  cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
  const char* pe = getenv("LSDM_SIDE_PRIO");
  const int pmode = pe ? atoi(pe) : 0;  // measured (B = 64, three buffer sets): 0 -> 2.88 ms/step, 1 -> 2.97, 2 -> 2.89, 3 -> 2.91
  // modes: 0 no priorities, 1 side stream high, 2 side + dense high, 3 dense high only
  cudaStreamCreateWithPriority(&h->side, cudaStreamNonBlocking, (pmode == 1 || pmode == 2) ? prio_hi : prio_lo);
  cudaStreamCreateWithPriority(&h->dense_st, cudaStreamNonBlocking, pmode >= 2 ? prio_hi : prio_lo);
  cudaStreamCreateWithPriority(&h->cond_st, cudaStreamNonBlocking, prio_lo);
  cudaStreamCreateWithPriority(&h->nn_st, cudaStreamNonBlocking, prio_lo);
  cudaEventCreateWithFlags(&h->ev_fps, cudaEventDisableTiming);
  cudaEventCreateWithFlags(&h->ev_nn, cudaEventDisableTiming);
  cudaEventCreateWithFlags(&h->ev_cfork, cudaEventDisableTiming);
  cudaEventCreateWithFlags(&h->ev_cjoin, cudaEventDisableTiming);
  cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming);
  for (int i = 0; i < 3; ++i) {
    cudaEventCreateWithFlags(&h->ev_sel[i], cudaEventDisableTiming);
    cudaEventCreateWithFlags(&h->ev_dense[i], cudaEventDisableTiming);
    cudaEventCreateWithFlags(&h->ev_step[i], cudaEventDisableTiming);
  }
  *out = h;
  return LSDM_OK;
}

LSDM_API void lsdm_destroy(lsdm_handle* h) {
  if (!h) return;
  int prev_dev = 0;
  cudaGetDevice(&prev_dev);
  cudaSetDevice(h->cfg.device);
  struct Restore { int d; ~Restore() { cudaSetDevice(d); } } restore{prev_dev};
  if (h->arena) cudaFree(h->arena);
  if (h->sched) cudaFree(h->sched);
  if (h->side) cudaStreamDestroy(h->side);
  if (h->dense_st) cudaStreamDestroy(h->dense_st);
  if (h->cond_st) cudaStreamDestroy(h->cond_st);
  if (h->nn_st) cudaStreamDestroy(h->nn_st);
  if (h->ev_fps) cudaEventDestroy(h->ev_fps);
  if (h->ev_nn) cudaEventDestroy(h->ev_nn);
  if (h->ev_cfork) cudaEventDestroy(h->ev_cfork);
  if (h->ev_cjoin) cudaEventDestroy(h->ev_cjoin);
  if (h->ev_fork) cudaEventDestroy(h->ev_fork);
  for (int i = 0; i < 3; ++i) {
    if (h->ev_sel[i]) cudaEventDestroy(h->ev_sel[i]);
    if (h->ev_dense[i]) cudaEventDestroy(h->ev_dense[i]);
    if (h->ev_step[i]) cudaEventDestroy(h->ev_step[i]);
  }
  delete h;
}

LSDM_API int lsdm_set_batch(lsdm_handle* h, int32_t bl, int32_t bg, int32_t off) {
  if (!h) return fail(LSDM_EINVAL, "null handle");
  if (bl <= 0 || bg < bl || off < 0 || off + bl > bg) return fail(LSDM_EINVAL, "bad batch configuration");
  if (bl != h->cfg.batch_local) h->have_ws = false;
  h->cfg.batch_local = bl;
  h->cfg.batch_global = bg;
  h->cfg.batch_offset = off;
  h->have_cond = false;
  return LSDM_OK;
}

LSDM_API int lsdm_num_weights(const lsdm_handle* h) { return h ? (int)h->entries.size() : 0; }
LSDM_API const char* lsdm_weight_key(const lsdm_handle* h, int i) {
  if (!h || i < 0 || i >= (int)h->entries.size()) return nullptr;
  return h->entries[i].key.c_str();
}

LSDM_API int lsdm_load_weight(lsdm_handle* h, const char* key, const void* data, const int64_t* shape, int32_t ndim, void* stream) {
  if (!h || !key) return fail(LSDM_EINVAL, "null argument");
  if (strncmp(key, "clip_model.", 11) == 0) return LSDM_OK;  // external text tower, out of scope
  auto it = h->index.find(key);
  if (it == h->index.end()) return fail(LSDM_EINVAL, std::string("unexpected state-dict key: ") + key);
  WEntry& e = h->entries[it->second];
  int64_t n = 1;
  for (int i = 0; i < ndim; ++i) n *= shape[i];
  bool same = (int)e.shape.size() == ndim;
  for (int i = 0; same && i < ndim; ++i) same = e.shape[i] == shape[i];
  if (!same) return fail(LSDM_EINVAL, std::string("shape mismatch for ") + key);
  if (e.off >= 0) {
    if (!data) return fail(LSDM_EINVAL, "null data");
    CK(cudaMemcpyAsync(h->arena + e.off, data, sizeof(float) * n, cudaMemcpyDefault, (cudaStream_t)stream));
  }
  e.loaded = true;
  h->finalized = false;
  return LSDM_OK;
}

LSDM_API int lsdm_finalize_weights(lsdm_handle* h, void* stream) {
  if (!h) return fail(LSDM_EINVAL, "null handle");
  for (const auto& e : h->entries)
    if (!e.loaded && e.off >= 0) return fail(LSDM_ESTATE, "missing state-dict key: " + e.key);
  cudaStream_t st = (cudaStream_t)stream;
  int64_t off = 0;
  auto take = [&](int64_t n) {
    float* p = h->derived + off;
    off += (n + 63) & ~int64_t(63);
    return p;
  };
  const float eps = 1e-5f;
  auto fold = [&](const std::string& conv, const std::string& bn, int N, int K, float** Wf, float** bf) {
    *Wf = take((int64_t)N * K);
    *bf = take(N);
    prof_launch(h, st, K_OTHER, [&] { return launch_fold_bn(h->W(conv + ".weight"), h->W(conv + ".bias"), h->W(bn + ".weight"), h->W(bn + ".bias"),
                                  h->W(bn + ".running_mean"), h->W(bn + ".running_var"), N, K, eps, *Wf, *bf, st); });
  };
  for (int l = 0; l < 4; ++l) {
    const SASpec& s = kSA[l];
    int last = s.cin;
    for (int i = 0; i < 3; ++i) {
      std::string p = std::string("pcd_backbone.") + s.name;
      fold(p + ".mlp_convs." + std::to_string(i), p + ".mlp_bns." + std::to_string(i), s.mlp[i], last, &h->sa_w[l][i],
           &h->sa_b[l][i]);
      last = s.mlp[i];
    }
    h->sa_wx[l] = take((int64_t)s.mlp[0] * 3);
    h->sa_wf[l] = take((int64_t)s.mlp[0] * (s.cin - 3));
    prof_launch(h, st, K_OTHER, [&] { return launch_copy_cols(h->sa_w[l][0], s.cin, 0, 3, s.mlp[0], h->sa_wx[l], st); });
    prof_launch(h, st, K_OTHER, [&] { return launch_copy_cols(h->sa_w[l][0], s.cin, 3, s.cin - 3, s.mlp[0], h->sa_wf[l], st); });
  }
  for (int l = 0; l < 4; ++l) {
    const FPSpec& s = kFP[l];
    int last = s.cin;
    for (int i = 0; i < s.nl; ++i) {
      std::string p = std::string("pcd_backbone.") + s.name;
      fold(p + ".mlp_convs." + std::to_string(i), p + ".mlp_bns." + std::to_string(i), s.mlp[i], last, &h->fp_w[l][i],
           &h->fp_b[l][i]);
      last = s.mlp[i];
    }
    h->fp_wa[l] = nullptr;
    if (s.Ca > 0) {
      h->fp_wa[l] = take((int64_t)s.mlp[0] * s.Ca);
      prof_launch(h, st, K_OTHER, [&] { return launch_copy_cols(h->fp_w[l][0], s.cin, 0, s.Ca, s.mlp[0], h->fp_wa[l], st); });
    }
    h->fp_wb[l] = take((int64_t)s.mlp[0] * s.Cb);
    prof_launch(h, st, K_OTHER, [&] { return launch_copy_cols(h->fp_w[l][0], s.cin, s.Ca, s.Cb, s.mlp[0], h->fp_wb[l], st); });
  }
  fold("pcd_backbone.conv1", "pcd_backbone.bn1", 128, 128, &h->head_w, &h->head_b);
  for (int l = 0; l < 4; ++l) {  // raw (un-folded) splits of the first SA / FP layers for the train-mode path
    const SASpec& s = kSA[l];
    const float* w0 = h->W(std::string("pcd_backbone.") + s.name + ".mlp_convs.0.weight");
    h->sa_wx_raw[l] = take((int64_t)s.mlp[0] * 3);
    h->sa_wf_raw[l] = take((int64_t)s.mlp[0] * (s.cin - 3));
    prof_launch(h, st, K_OTHER, [&] { return launch_copy_cols(w0, s.cin, 0, 3, s.mlp[0], h->sa_wx_raw[l], st); });
    prof_launch(h, st, K_OTHER, [&] { return launch_copy_cols(w0, s.cin, 3, s.cin - 3, s.mlp[0], h->sa_wf_raw[l], st); });
    const FPSpec& f = kFP[l];
    const float* f0 = h->W(std::string("pcd_backbone.") + f.name + ".mlp_convs.0.weight");
    h->fp_wa_raw[l] = nullptr;
    if (f.Ca > 0) {
      h->fp_wa_raw[l] = take((int64_t)f.mlp[0] * f.Ca);
      prof_launch(h, st, K_OTHER, [&] { return launch_copy_cols(f0, f.cin, 0, f.Ca, f.mlp[0], h->fp_wa_raw[l], st); });
    }
    h->fp_wb_raw[l] = take((int64_t)f.mlp[0] * f.Cb);
    prof_launch(h, st, K_OTHER, [&] { return launch_copy_cols(f0, f.cin, f.Ca, f.Cb, f.mlp[0], h->fp_wb_raw[l], st); });
  }
  if (off > h->derived_floats) return fail(LSDM_ENOMEM, "derived weight arena too small");
  CK(cudaPeekAtLastError());
  // TF32-rounded (round-to-nearest) copy of every weight: the cp.async-fed tensor GEMM reads operands without touching them
  prof_launch(h, st, K_OTHER, [&] {
    int64_t n = h->round_delta;
    round_tf32_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(h->arena, h->arena + h->round_delta, h->arena + h->lo_delta, n);
    return 1;
  });
  // host copies of the small per-channel vectors of sa1 / sa2 (kernel parameters of the v2 fused kernels)
  for (int l = 0; l < 2; ++l) {
    const int C1 = kSA[l].mlp[0], C2 = kSA[l].mlp[1];
    h->host_wx[l].resize(C1 * 3);
    h->host_wf[l].assign(C1 * 3, 0.f);
    h->host_b1[l].resize(C1);
    h->host_b2[l].resize(C2);
    CK(cudaMemcpyAsync(h->host_wx[l].data(), h->sa_wx[l], sizeof(float) * C1 * 3, cudaMemcpyDeviceToHost, st));
    if (l == 0) CK(cudaMemcpyAsync(h->host_wf[l].data(), h->sa_wf[l], sizeof(float) * C1 * 3, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(h->host_b1[l].data(), h->sa_b[l][0], sizeof(float) * C1, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(h->host_b2[l].data(), h->sa_b[l][1], sizeof(float) * C2, cudaMemcpyDeviceToHost, st));
  }
  h->host_fp1_b1.resize(128);
  CK(cudaMemcpyAsync(h->host_fp1_b1.data(), h->fp_b[3][0], sizeof(float) * 128, cudaMemcpyDeviceToHost, st));
  h->host_tail.resize(771);
  CK(cudaMemcpyAsync(h->host_tail.data(), h->fp_b[3][1], sizeof(float) * 128, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(h->host_tail.data() + 128, h->fp_b[3][2], sizeof(float) * 128, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(h->host_tail.data() + 256, h->head_b, sizeof(float) * 128, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(h->host_tail.data() + 384, h->W("pcd_backbone.conv2.weight"), sizeof(float) * 384, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(h->host_tail.data() + 768, h->W("pcd_backbone.conv2.bias"), sizeof(float) * 3, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));  // one-off, load time only
  h->finalized = true;
  h->fold_dirty = false;
  return LSDM_OK;
}

LSDM_API int lsdm_set_schedule(lsdm_handle* h, const float* c1, const float* c2, const float* logvar, const float* sa, const float* s1a,
                      int32_t T, void* stream) {
  if (!h || !c1 || !c2 || !logvar || !sa || !s1a || T <= 0) return fail(LSDM_EINVAL, "bad schedule");
  cudaStream_t st = (cudaStream_t)stream;
  if (h->sched && h->T != T) {
    CK(cudaFree(h->sched));
    h->sched = nullptr;
  }
  if (!h->sched) CK(cudaMalloc(&h->sched, sizeof(float) * 5 * T));
  h->T = T;
  const float* src[5] = {c1, c2, logvar, sa, s1a};
  for (int i = 0; i < 5; ++i) CK(cudaMemcpyAsync(h->sched + (size_t)i * T, src[i], sizeof(float) * T, cudaMemcpyDefault, st));
  h->have_sched = true;
  return LSDM_OK;
}

LSDM_API size_t lsdm_workspace_bytes(const lsdm_handle* h) {
  if (!h) return 0;
  Workspace w;
  return carve(h, nullptr, &w);
}

LSDM_API int lsdm_set_workspace(lsdm_handle* h, void* workspace, size_t bytes) {
  if (!h || !workspace) return fail(LSDM_EINVAL, "null argument");
  if ((reinterpret_cast<uintptr_t>(workspace) & 255) != 0) return fail(LSDM_EINVAL, "workspace must be 256-byte aligned");
  size_t need = carve(h, workspace, &h->ws);
  if (bytes < need) {
    h->have_ws = false;
    return fail(LSDM_EINVAL, "workspace too small: need " + std::to_string(need));
  }
  h->have_ws = true;
  h->have_cond = false;
  return LSDM_OK;
}

// Train-mode PointNet++ dense layers (model.train(): reference pointnet2_utils.py:192-195,308-311, pointnet2.py:76): every
// conv is followed by BatchNorm with BATCH statistics over all 9B clouds, running statistics are updated in the handle's
// weight arena, and the backbone head applies the caller's Dropout(0.5) mask.  Layers run un-fused: conv (GEMM, raw weights)
// -> column statistics -> normalise + ReLU (+ max-pool).
int dense_phase_train(lsdm_handle* h, const Workspace::Sel& q, const float* clouds, const float* drop_mask, cudaStream_t st) {
  Workspace& w = h->ws;
  const int C = h->cfg.batch_local * NOBJ;
  const float* xyz[5] = {clouds, q.xyz[1], q.xyz[2], q.xyz[3], q.xyz[4]};
  const float* feat[5] = {clouds, w.feat[1], w.feat[2], w.feat[3], w.feat[4]};
  auto mut = [&](const std::string& k) { return const_cast<float*>(h->W(k)); };
  // statistics are over the GLOBAL batch: with a SyncBN hook the per-shard (sum, sum of squares) are all-reduced and the
  // row count is scaled by the number of (equal-sized) shards
  const int64_t shards = h->allreduce ? h->cfg.batch_global / h->cfg.batch_local : 1;
  int hook_err = 0;
  auto bn = [&](const std::string& key, float* y, int64_t M, int N, const float* mask, float* pooled) {
    prof_launch(h, st, K_OTHER, [&] { return launch_col_stats(y, M, N, w.bn_stats, st); });
    if (h->allreduce && h->allreduce(h->allreduce_ctx, w.bn_stats, 2 * N) != 0) hook_err = 1;
    prof_launch(h, st, K_OTHER, [&] {
      return launch_bn_apply(y, M, M * shards, N, w.bn_stats, h->W(key + ".weight"), h->W(key + ".bias"), mut(key + ".running_mean"),
                             mut(key + ".running_var"), mask, NPTS, pooled, 0, st);
    });
  };
  for (int l = 0; l < 4; ++l) {
    const SASpec& s = kSA[l];
    const std::string p = std::string("pcd_backbone.") + s.name;
    const int N = s.N, S = s.npoint, C1 = s.mlp[0], C2 = s.mlp[1], C3 = s.mlp[2];
    const float* b0 = h->W(p + ".mlp_convs.0.bias");
    const float* P = nullptr;
    if (l > 0) {
      GE(gemm(h, st, feat[l], s.cin - 3, h->sa_wf_raw[l], s.cin - 3, w.tP, C1, b0, C * N, C1, s.cin - 3, ACT_NONE));
      P = w.tP;
    }
    prof_launch(h, st, K_GATHER, [&] {
      return launch_sa_gather(P, h->sa_wx_raw[l], h->sa_wf_raw[l], b0, xyz[l], xyz[l + 1], q.grp[l], C, N, S, C1, w.tA, 0, st, 0);
    });
    const int rows = C * S * 32;
    bn(p + ".mlp_bns.0", w.tA, rows, C1, nullptr, nullptr);
    GE(gemm(h, st, w.tA, C1, h->W(p + ".mlp_convs.1.weight"), C1, w.tB, C2, h->W(p + ".mlp_convs.1.bias"), rows, C2, C1, ACT_NONE));
    bn(p + ".mlp_bns.1", w.tB, rows, C2, nullptr, nullptr);
    GE(gemm(h, st, w.tB, C2, h->W(p + ".mlp_convs.2.weight"), C2, w.tA, C3, h->W(p + ".mlp_convs.2.bias"), rows, C3, C2, ACT_NONE));
    bn(p + ".mlp_bns.2", w.tA, rows, C3, nullptr, w.feat[l + 1]);
  }
  const int fine[4] = {3, 2, 1, 0};
  const int fineN[4] = {64, 256, 1024, 1024}, coarseN[4] = {16, 64, 256, 1024};
  const float* coarse_feat = w.feat[4];
  float* outs[4] = {w.g3, w.g2, w.g1, w.tA};
  for (int l = 0; l < 4; ++l) {
    const FPSpec& s = kFP[l];
    const std::string p = std::string("pcd_backbone.") + s.name;
    const int N = fineN[l], S = coarseN[l], C1 = s.mlp[0];
    const float* b0 = h->W(p + ".mlp_convs.0.bias");
    const float* Pa = nullptr;
    if (s.Ca > 0) {
      GE(gemm(h, st, feat[fine[l]], s.Ca, h->fp_wa_raw[l], s.Ca, w.tA, C1, b0, C * N, C1, s.Ca, ACT_NONE));
      Pa = w.tA;
    }
    GE(gemm(h, st, coarse_feat, s.Cb, h->fp_wb_raw[l], s.Cb, w.tB, C1, nullptr, C * S, C1, s.Cb, ACT_NONE));
    prof_launch(h, st, K_FPCOMB, [&] { return launch_fp_combine(Pa, b0, w.tB, q.nn_idx[l], q.nn_w[l], C, N, S, C1, w.tP, 0, st, 0); });
    bn(p + ".mlp_bns.0", w.tP, (int64_t)C * N, C1, nullptr, nullptr);
    GE(gemm(h, st, w.tP, C1, h->W(p + ".mlp_convs.1.weight"), C1, outs[l], s.mlp[1], h->W(p + ".mlp_convs.1.bias"), C * N, s.mlp[1], C1,
            ACT_NONE));
    bn(p + ".mlp_bns.1", outs[l], (int64_t)C * N, s.mlp[1], nullptr, nullptr);
    coarse_feat = outs[l];
    if (l == 3) {
      GE(gemm(h, st, w.tA, 128, h->W(p + ".mlp_convs.2.weight"), 128, w.tP, 128, h->W(p + ".mlp_convs.2.bias"), C * N, 128, 128, ACT_NONE));
      bn(p + ".mlp_bns.2", w.tP, (int64_t)C * N, 128, nullptr, nullptr);
      GE(gemm(h, st, w.tP, 128, h->W("pcd_backbone.conv1.weight"), 128, w.tA, 128, h->W("pcd_backbone.conv1.bias"), C * N, 128, 128, ACT_NONE));
      bn("pcd_backbone.bn1", w.tA, (int64_t)C * N, 128, drop_mask, nullptr);
      prof_launch(h, st, K_HEAD, [&] {
        return launch_head3(w.tA, h->W("pcd_backbone.conv2.weight"), h->W("pcd_backbone.conv2.bias"), (int64_t)C * N, w.backbone, st);
      });
    }
  }
  h->fold_dirty = true;
  if (hook_err) return fail(LSDM_EINVAL, "all-reduce hook failed");
  CK(cudaPeekAtLastError());
  return LSDM_OK;
}

// Everything of the condition encoder except the selection chain (which the caller has already enqueued for set `si`).
static int encode_dense(lsdm_handle* h, const float* text, const float* objs, const float* cats, const float* mask_global, int si,
                        cudaStream_t st, const float* train_drop_mask = nullptr, const float* clouds = nullptr, int n_clouds = -1,
                        const int* remap = nullptr, bool sa1_canon = false) {
  Workspace& w = h->ws;
  const int B = h->cfg.batch_local;
  if (!clouds) clouds = objs, n_clouds = B * NOBJ;
  if (train_drop_mask) {
    GE(dense_phase_train(h, w.sel[si], objs, train_drop_mask, st));
  } else {
    if (h->fold_dirty) GE(lsdm_finalize_weights(h, st));  // running statistics moved since the last fold
    GE(dense_phase(h, w.sel[si], clouds, st, n_clouds, sa1_canon));
  }
  SceneWeights sw{h->W("pcd_attention.k_proj_weight"), h->W("pcd_attention.v_proj_weight"), h->W("pcd_attention.in_proj_bias"),
                  h->W("pcd_attention.out_proj.weight"), h->W("pcd_attention.out_proj.bias"),
                  h->W("point_wise_trans_layer.0.weight"), h->W("point_wise_trans_layer.0.bias")};
  prof_launch(h, st, K_SCENE, [&] { return launch_point_attention(sw, w.backbone, w.sel[si].attn_w, w.sel[si].qq, B, w.pa, w.pw, st, remap); });
  prof_launch(h, st, K_SCENE, [&] { return launch_scene_mix(w.pw, w.sel[si].hm, mask_global, B, h->cfg.batch_global, h->cfg.batch_offset, w.pcd_out[si], st); });
  CK(cudaPeekAtLastError());
  w.cur = si;
  h->have_cond = true;
  return LSDM_OK;
}

LSDM_API int lsdm_encode_conditions(lsdm_handle* h, const float* text, const float* objs, const float* cats, const float* mask_global,
                           const int64_t* fps_start, void* stream) {
  GE(check_ready(h, false));
  if (!text || !objs || !cats || !mask_global || !fps_start) return fail(LSDM_EINVAL, "null argument");
  cudaStream_t st = (cudaStream_t)stream;
  GE(select_phase(h, h->ws.sel[0], text, objs, cats, mask_global, fps_start, st));
  return encode_dense(h, text, objs, cats, mask_global, 0, st);
}

LSDM_API int lsdm_encode_conditions_train(lsdm_handle* h, const float* text, const float* objs, const float* cats,
                                          const float* mask_global, const int64_t* fps_start, const float* drop_mask, void* stream) {
  GE(check_ready(h, false));
  if (!text || !objs || !cats || !mask_global || !fps_start || !drop_mask) return fail(LSDM_EINVAL, "null argument");
  cudaStream_t st = (cudaStream_t)stream;
  GE(select_phase(h, h->ws.sel[0], text, objs, cats, mask_global, fps_start, st));
  return encode_dense(h, text, objs, cats, mask_global, 0, st, drop_mask);
}

LSDM_API int lsdm_set_allreduce(lsdm_handle* h, lsdm_allreduce_fn fn, void* ctx) {
  if (!h) return fail(LSDM_EINVAL, "null handle");
  h->allreduce = fn;
  h->allreduce_ctx = ctx;
  return LSDM_OK;
}

LSDM_API int lsdm_read_weight(lsdm_handle* h, const char* key, float* dst, int64_t numel, void* stream) {
  if (!h || !key || !dst) return fail(LSDM_EINVAL, "null argument");
  auto it = h->index.find(key);
  if (it == h->index.end() || h->entries[it->second].off < 0) return fail(LSDM_EINVAL, std::string("unknown key: ") + key);
  const WEntry& e = h->entries[it->second];
  if (numel != e.numel) return fail(LSDM_EINVAL, std::string("size mismatch for ") + key);
  CK(cudaMemcpyAsync(dst, h->arena + e.off, sizeof(float) * e.numel, cudaMemcpyDefault, (cudaStream_t)stream));
  return LSDM_OK;
}

LSDM_API int lsdm_denoise_step(lsdm_handle* h, float* x, const int64_t* t, const float* noise, float* sample_out, float* x0_out,
                      float* guiding_out, int32_t clip_denoised, void* stream) {
  GE(check_ready(h, true));
  if (!h->have_sched) return fail(LSDM_ESTATE, "no schedule (lsdm_set_schedule)");
  if (!x || !t || !noise || !sample_out) return fail(LSDM_EINVAL, "null argument");
  return step_core(h, x, t, noise, sample_out, x0_out, guiding_out, true, clip_denoised, (cudaStream_t)stream);
}

LSDM_API int lsdm_forward(lsdm_handle* h, float* x, const int64_t* t, float* out_cat, float* x0, float* guiding, void* stream) {
  GE(check_ready(h, true));
  if (!x || !t || !x0) return fail(LSDM_EINVAL, "null argument");
  cudaStream_t st = (cudaStream_t)stream;
  GE(step_core(h, x, t, nullptr, nullptr, x0, guiding, true, 0, st));
  if (out_cat)
    CK(cudaMemcpyAsync(out_cat, h->ws.sel[h->ws.cur].out_cat, sizeof(float) * h->cfg.batch_local * h->cfg.n_cats, cudaMemcpyDeviceToDevice, st));
  return LSDM_OK;
}

LSDM_API int lsdm_sample_loop(lsdm_handle* h, float* x, const float* text, const float* objs, const float* cats, const float* mask_global,
                     const int64_t* fps_start_all, const float* noise_all, int32_t t_first, int32_t n_steps, int32_t hoisted,
                     int32_t clip_denoised, float* x0_out, float* guiding_out, void* stream) {
  GE(check_ready(h, false));
  if (!h->have_sched) return fail(LSDM_ESTATE, "no schedule (lsdm_set_schedule)");
  if (!x || !fps_start_all || !noise_all) return fail(LSDM_EINVAL, "null argument");
  if (n_steps <= 0 || t_first >= h->T || t_first - n_steps + 1 < 0) return fail(LSDM_EINVAL, "bad timestep range");
  cudaStream_t st = (cudaStream_t)stream;
  const int B = h->cfg.batch_local, C = B * NOBJ;
  const size_t per = (size_t)B * NPTS * 3;
  // device-side timestep vector (the reference builds th.tensor([i]*B) on the host every step, gaussian_diffusion.py:737)
  int64_t* tvec = h->ws.t_dev;
  if (!text || !objs || !cats || !mask_global) return fail(LSDM_EINVAL, "null argument");
  // STRICT, three-stage software pipeline over steps (nothing in the condition encoder depends on x):
  //   side stream  : selection chain + condition MLPs + human decoder, up to two steps ahead   (FPS -> ball queries -> 3-NN)
  //   dense stream : PointNet++ dense layers + scene branch of step k               -> pcd_out[k%3]
  //   caller stream: x0 network + posterior of step k (the only part that is serial in x)
  // Three buffer sets (selection(k+1) only has to wait for the consumers of step k-2); events order producer -> consumer and
  // consumer -> reuse.
  // absent-cloud de-duplication: classify once per call (the clouds are constant over the loop), run the encoder on the
  // compacted list, read the shared result through `remap` in the scene branch
  const float* clouds = nullptr;
  const int *remap = nullptr, *active = nullptr;
  int nc = -1;
  if (h->dedup_absent) {
    Workspace& w = h->ws;
    classify_clouds_kernel<<<C, 256, 0, st>>>(objs, w.remap);
    compact_clouds_kernel<<<1, 32, 0, st>>>(w.remap, w.active, w.n_active, C);
    int n_act = 0;
    CK(cudaMemcpyAsync(&n_act, w.n_active, sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));  // once per call: grid sizes of the encoder depend on the count
    h->launches += 2;
    if (n_act < C) {
      gather_clouds_kernel<<<n_act, 256, 0, st>>>(objs, w.active, w.clouds_c);
      h->launches += 1;
      clouds = w.clouds_c;
      remap = w.remap;
      active = w.active;
      nc = n_act;
    }
    h->n_active = n_act;
  } else {
    h->n_active = C;
  }
  // LSDM_TIMELINE=1: timing events at the start / end of every stage of every step, printed (after a synchronize) to stderr
  static const bool timeline = getenv("LSDM_TIMELINE") && atoi(getenv("LSDM_TIMELINE")) != 0;
  struct TL { int k; char stage; cudaEvent_t a, b; };
  std::vector<TL> tl;
  cudaEvent_t tl0 = nullptr;
  auto tl_begin = [&](int k, char stage, cudaStream_t s) {
    if (!timeline) return;
    TL e{k, stage, nullptr, nullptr};
    cudaEventCreate(&e.a);
    cudaEventCreate(&e.b);
    cudaEventRecord(e.a, s);
    tl.push_back(e);
  };
  auto tl_end = [&](cudaStream_t s) {
    if (timeline) cudaEventRecord(tl.back().b, s);
  };
  if (timeline) {
    cudaEventCreate(&tl0);
    cudaEventRecord(tl0, st);
  }
  const bool pipelined = !hoisted && n_steps > 1 && !h->profiling;
  cudaStream_t side = pipelined ? h->side : st;
  cudaStream_t dst = pipelined ? h->dense_st : st;
  if (pipelined) {
    CK(cudaEventRecord(h->ev_fork, st));
    CK(cudaStreamWaitEvent(side, h->ev_fork, 0));
    CK(cudaStreamWaitEvent(dst, h->ev_fork, 0));
  }
  // loop invariants (STRICT): what reads nothing that changes over the loop is computed once per call.
  //  bit 0  text / category / translation MLPs, object-attention weights, human decoder: step 0's results are copied into the
  //         other two buffer sets and later steps skip those kernels;
  //  bit 2  sa1 keeps all 1024 points of a cloud (npoint = N), so the level-0 FPS draw only decides the ORDER of its centroids: a
  //         centroid's ball-query group and pooled feature are functions of its coordinates and the cloud.  They are computed once in
  //         cloud order (on the dense stream, which is idle during the first selection) and every step gathers its level-1 rows by
  //         its level-0 FPS indices (the FPS itself still runs: level 1 depends on the order)
  const int inv = (!hoisted && n_steps > 1) ? h->loop_invariants : 0;
  const bool skip_cond = (inv & 1) != 0;
  const bool sa1_canon = (inv & 4) != 0 && h->sa1_compact && h->precision >= 1 && h->sa_fused > 0 && (h->select_grid & 1);
  if (sa1_canon) {
    const int n_enc = clouds ? nc : C;
    GE(sa1_cloud_order(h, clouds ? clouds : objs, n_enc, dst));
    if (h->profiling) h->gemm_flops += 2.0 * n_enc * 1024 * 32 * ((double)kSA[0].mlp[0] * kSA[0].mlp[1] + (double)kSA[0].mlp[1] * kSA[0].mlp[2]);
  }
  tl_begin(0, 'S', side);
  const bool fork_cond = pipelined && h->cond_stream;
  GE(select_phase(h, h->ws.sel[0], text, objs, cats, mask_global, fps_start_all, side, clouds, nc, active, fork_cond, false, sa1_canon));
  if (skip_cond) {
    Workspace::Sel& s0 = h->ws.sel[0];
    for (int j = 1; j < 3; ++j) {
      Workspace::Sel& sj = h->ws.sel[j];
      CK(cudaMemcpyAsync(sj.enc, s0.enc, sizeof(float) * B * LAT, cudaMemcpyDeviceToDevice, side));
      CK(cudaMemcpyAsync(sj.out_cat, s0.out_cat, sizeof(float) * B * h->cfg.n_cats, cudaMemcpyDeviceToDevice, side));
      CK(cudaMemcpyAsync(sj.attn_w, s0.attn_w, sizeof(float) * B * NOBJ, cudaMemcpyDeviceToDevice, side));
      CK(cudaMemcpyAsync(sj.tr, s0.tr, sizeof(float) * C * TRANS, cudaMemcpyDeviceToDevice, side));
      CK(cudaMemcpyAsync(sj.qq, s0.qq, sizeof(float) * C * TRANS, cudaMemcpyDeviceToDevice, side));
      CK(cudaMemcpyAsync(sj.hm, s0.hm, sizeof(float) * B * NPTS * 3, cudaMemcpyDeviceToDevice, side));
    }
  }
  tl_end(side);
  if (pipelined) CK(cudaEventRecord(h->ev_sel[0], side));
  for (int k = 0; k < n_steps; ++k) {
    const int si = hoisted ? 0 : (k % 3);
    if (pipelined && k + 1 < n_steps) {
      const int sn = (k + 1) % 3;
      if (k >= 2) {  // set sn (selection results, condition MLP outputs, pcd_out[sn]) was last read by dense(k-2) and step(k-2)
        CK(cudaStreamWaitEvent(side, h->ev_dense[sn], 0));
        CK(cudaStreamWaitEvent(side, h->ev_step[sn], 0));
      }
      tl_begin(k + 1, 'S', side);
      GE(select_phase(h, h->ws.sel[sn], text, objs, cats, mask_global, fps_start_all + (size_t)(k + 1) * 4 * C, side, clouds, nc, active, fork_cond,
                      skip_cond, sa1_canon));
      tl_end(side);
      CK(cudaEventRecord(h->ev_sel[sn], side));
    }
    if (!pipelined && !hoisted && k >= 1)
      GE(select_phase(h, h->ws.sel[si], text, objs, cats, mask_global, fps_start_all + (size_t)k * 4 * C, st, clouds, nc, active, false, skip_cond, sa1_canon));
    if (!hoisted || k == 0) {
      if (pipelined) {
        CK(cudaStreamWaitEvent(dst, h->ev_sel[si], 0));  // (select(k) already waited for step(k-3), the last reader of pcd_out[si])
      }
      tl_begin(k, 'D', dst);
      GE(encode_dense(h, text, objs, cats, mask_global, si, dst, nullptr, clouds, nc, remap, sa1_canon));
      tl_end(dst);
      if (pipelined) {
        CK(cudaEventRecord(h->ev_dense[si], dst));
        CK(cudaStreamWaitEvent(st, h->ev_dense[si], 0));
      }
    }
    prof_launch(h, st, K_OTHER, [&] {
      fill_t_kernel<<<(B + 127) / 128, 128, 0, st>>>(tvec, B, (int64_t)(t_first - k));
      return 1;
    });
    const bool last = (k == n_steps - 1);
    // STRICT recomputes the guiding points every step like the reference; hoisted only needs them at the end
    // (`loop_invariants` bit 3: the guiding points of a step are only ever visible after the call's last step -- the model keeps the
    // latest `saved_guiding_points` --, so the second x0-network pass of the earlier steps is a dead store)
    const bool want_guiding = last || (!hoisted && (inv & 8) == 0);
    tl_begin(k, 'X', st);
    // the embedding's text half once per call, its time half once per step for the whole batch: the hoisted loop (`hoist_split`) and,
    // because text and the shared t are just as loop-invariant / batch-shared there, the STRICT loop (`loop_invariants` bit 1)
    const bool split_emb = hoisted ? (h->hoist_split != 0) : ((inv & 2) != 0);
    const int hs = (split_emb && h->x0_fused) ? (k == 0 ? 1 : 2) : 0;
    const float* a_t_pre = nullptr;
    if (hs && h->precision_step == 2 && g_gemm_async != 0 && h->time_batch) {
      if (k % TIME_BATCH == 0) GE(time_half_batch(h, (int64_t)t_first - k, std::min(TIME_BATCH, n_steps - k), st));
      a_t_pre = h->ws.a_t_all + (size_t)(k % TIME_BATCH) * NPTS * 128;
    }
    GE(step_core(h, x, tvec, noise_all + (size_t)k * per, x, last ? x0_out : nullptr, last ? guiding_out : nullptr,
                 want_guiding, clip_denoised, st, si, hs, a_t_pre));
    tl_end(st);
    if (pipelined) CK(cudaEventRecord(h->ev_step[si], st));
  }
  if (timeline) {
    cudaDeviceSynchronize();
    fprintf(stderr, "[lsdm timeline] n_steps=%d clouds=%d/%d  (stage S=selection D=dense X=x0-network; ms since call start)\n", n_steps, h->n_active, C);
    for (auto& e : tl) {
      float a = 0.f, b = 0.f;
      cudaEventElapsedTime(&a, tl0, e.a);
      cudaEventElapsedTime(&b, tl0, e.b);
      if (e.k < 6 || e.k >= n_steps - 2) fprintf(stderr, "[lsdm timeline] k=%3d %c  %8.3f -> %8.3f  (%.3f)\n", e.k, e.stage, a, b, b - a);
      cudaEventDestroy(e.a);
      cudaEventDestroy(e.b);
    }
    cudaEventDestroy(tl0);
  }
  return LSDM_OK;
}

LSDM_API int lsdm_get_out_cat(lsdm_handle* h, float* out_cat, void* stream) {
  GE(check_ready(h, true));
  CK(cudaMemcpyAsync(out_cat, h->ws.sel[h->ws.cur].out_cat, sizeof(float) * h->cfg.batch_local * h->cfg.n_cats, cudaMemcpyDeviceToDevice,
                     (cudaStream_t)stream));
  return LSDM_OK;
}
LSDM_API int lsdm_get_pcd_out(lsdm_handle* h, float* pcd_out, void* stream) {
  GE(check_ready(h, true));
  CK(cudaMemcpyAsync(pcd_out, h->ws.pcd_out[h->ws.cur], sizeof(float) * h->cfg.batch_local * NPTS * 3, cudaMemcpyDeviceToDevice,
                     (cudaStream_t)stream));
  return LSDM_OK;
}

LSDM_API int lsdm_q_sample(lsdm_handle* h, const float* x_start, const int64_t* t, const float* noise, float* x_t, void* stream) {
  GE(check_ready(h, false));
  if (!h->have_sched) return fail(LSDM_ESTATE, "no schedule (lsdm_set_schedule)");
  cudaStream_t st = (cudaStream_t)stream;
  CK(cudaMemcpyAsync(h->ws.t_dev, t, sizeof(int64_t) * h->cfg.batch_local, cudaMemcpyDefault, st));
  prof_launch(h, st, K_DENOISE, [&] { return launch_q_sample(x_start, h->ws.t_dev, noise, h->sched + 3 * h->T, h->sched + 4 * h->T, h->cfg.batch_local, x_t, st); });
  CK(cudaPeekAtLastError());
  return LSDM_OK;
}

LSDM_API int lsdm_chamfer(lsdm_handle* h, const float* x, const float* y, int32_t batch, int32_t n, int32_t m, float* sums, void* stream) {
  if (!h || !x || !y || !sums || batch <= 0 || n <= 0 || m <= 0 || n > 4096 || m > 4096) return fail(LSDM_EINVAL, "bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  prof_launch(h, st, K_OTHER, [&] { return launch_chamfer(x, y, batch, n, m, sums, st); });
  CK(cudaPeekAtLastError());
  return LSDM_OK;
}

// ---- evaluation metrics (handle-free: they use no weights and no workspace beyond what the caller passes) ----
LSDM_API int lsdm_eval_emd(const float* x, const float* y, int32_t batch, int32_t n, int32_t m, double* emd, int32_t* assignment,
                           int32_t* rounds, void* stream) {
  if (!x || !y || !emd || batch <= 0) return fail(LSDM_EINVAL, "bad argument");
  if (n != m) return fail(LSDM_EINVAL, "lsdm_eval_emd: clouds must have the same number of points (perfect matching)");
  if (n < 1 || n > 1024) return fail(LSDM_EINVAL, "lsdm_eval_emd: 1 <= n <= 1024");
  cudaStream_t st = (cudaStream_t)stream;
  if (launch_emd(x, y, batch, n, emd, assignment, rounds, st) < 0) return fail(LSDM_ECUDA, "emd launch");
  CK(cudaPeekAtLastError());
  return LSDM_OK;
}

LSDM_API int lsdm_eval_fscore(const float* gt, const float* pr, int32_t batch, int32_t n, int32_t m, double th, int32_t* counts,
                              double* out, void* stream) {
  if (!gt || !pr || !counts || !out || batch <= 0 || n <= 0 || m <= 0 || n > 4096 || m > 4096) return fail(LSDM_EINVAL, "bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  if (launch_fscore(gt, pr, batch, n, m, th, counts, out, st) < 0) return fail(LSDM_EINVAL, "fscore launch");
  CK(cudaPeekAtLastError());
  return LSDM_OK;
}

LSDM_API int lsdm_eval_chamfer(const float* x, const float* y, int32_t batch, int32_t n, int32_t m, float* per_sample, void* stream) {
  if (!x || !y || !per_sample || batch <= 0 || n <= 0 || m <= 0 || n > 4096 || m > 4096) return fail(LSDM_EINVAL, "bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  CK(cudaMemsetAsync(per_sample, 0, sizeof(float) * 2 * batch, st));
  launch_chamfer(x, y, batch, n, m, per_sample, st, 2);
  CK(cudaPeekAtLastError());
  return LSDM_OK;
}

LSDM_API int lsdm_eval_topk(const float* scores, const int64_t* target, int32_t batch, int32_t n_classes, const int32_t* ks, int32_t nk,
                            int32_t* correct, void* stream) {
  if (!scores || !target || !ks || !correct || batch <= 0 || n_classes <= 0 || nk <= 0) return fail(LSDM_EINVAL, "bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  launch_topk(scores, target, batch, n_classes, ks, nk, correct, st);
  CK(cudaPeekAtLastError());
  return LSDM_OK;
}

LSDM_API int lsdm_cat_loss(lsdm_handle* h, const float* probs, const float* target_cat, int32_t batch, float* sum, void* stream) {
  if (!h || !probs || !target_cat || !sum || batch <= 0) return fail(LSDM_EINVAL, "bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  prof_launch(h, st, K_OTHER, [&] { return launch_cat_loss(probs, target_cat, batch, h->cfg.n_cats, sum, st); });
  CK(cudaPeekAtLastError());
  return LSDM_OK;
}

// ---- training backward: reference run/train_sdm.py:78-84 `mp_trainer.backward(loss)` ----
LSDM_API size_t lsdm_train_tape_bytes(const lsdm_handle* h) { return h ? train_tape_bytes(h->cfg.batch_local, h->cfg.n_cats) : 0; }
LSDM_API int64_t lsdm_grad_floats(const lsdm_handle* h) { return h ? h->arena_floats : 0; }
LSDM_API int lsdm_weight_slot(const lsdm_handle* h, int32_t i, int64_t* offset, int64_t* numel) {
  if (!h || i < 0 || i >= (int)h->entries.size() || !offset || !numel) return fail(LSDM_EINVAL, "bad argument");
  *offset = h->entries[i].off;
  *numel = h->entries[i].numel;
  return LSDM_OK;
}

LSDM_API int lsdm_training_backward(lsdm_handle* h, const float* x_start, const int64_t* t, const float* noise, const float* text,
                                    const float* objs, const float* cats, const float* mask_global, const float* target_cat,
                                    const int64_t* fps_start, const float* drop_mask, float lambda_cat, float g_mse, float g_cat, void* tape,
                                    size_t tape_bytes, float* grads, float* losses_out, float* x0_out, void* stream) {
  GE(check_ready(h, false));
  if (!h->have_sched) return fail(LSDM_ESTATE, "no schedule (lsdm_set_schedule)");
  if (!x_start || !t || !noise || !text || !objs || !cats || !mask_global || !target_cat || !fps_start || !drop_mask || !tape || !grads)
    return fail(LSDM_EINVAL, "null argument");
  cudaStream_t st = (cudaStream_t)stream;
  Workspace& w = h->ws;
  const int B = h->cfg.batch_local;
  // selection chain (FPS order, ball-query groups, 3-NN) for these clouds and FPS starts: integer / piecewise-constant, no gradient
  GE(select_phase(h, w.sel[0], text, objs, cats, mask_global, fps_start, st));
  CK(cudaMemcpyAsync(w.t_dev, t, sizeof(int64_t) * B, cudaMemcpyDefault, st));
  TrainCtx ctx;
  ctx.W = [h](const std::string& k) { return h->W(k); };
  ctx.G = [h, grads](const std::string& k) { return grads + h->entries[h->index.at(k)].off; };
  ctx.B = B;
  ctx.Bg = h->cfg.batch_global;
  ctx.b_off = h->cfg.batch_offset;
  ctx.n_cats = h->cfg.n_cats;
  ctx.tape = tape;
  ctx.tape_bytes = tape_bytes;
  ctx.allreduce = h->allreduce;
  ctx.allreduce_ctx = h->allreduce_ctx;
  ctx.sched_sa = h->sched + 3 * h->T;
  ctx.sched_s1a = h->sched + 4 * h->T;
  TrainIO io;
  io.text = text; io.objs = objs; io.cats = cats; io.mask_global = mask_global; io.x_start = x_start; io.noise = noise;
  io.target_cat = target_cat; io.drop_mask = drop_mask; io.t = w.t_dev;
  const Workspace::Sel& q = w.sel[0];
  for (int l = 1; l <= 4; ++l) io.xyz[l] = q.xyz[l];
  for (int l = 0; l < 4; ++l) {
    io.grp[l] = q.grp[l];
    io.nn_idx[l] = q.nn_idx[l];
    io.nn_w[l] = q.nn_w[l];
  }
  io.g_mse = g_mse; io.g_cat = g_cat; io.lambda_cat = lambda_cat;
  io.losses_out = losses_out;
  io.x0_out = x0_out;
  int r = train_forward_backward(ctx, io, st);
  if (r != 0) return r;
  h->launches += 1;
  return LSDM_OK;
}

LSDM_API int lsdm_adamw_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr, float beta1, float beta2,
                             float eps, float weight_decay, int64_t step, float grad_scale, void* stream) {
  if (!param || !grad || !exp_avg || !exp_avg_sq || n <= 0 || step < 1) return fail(LSDM_EINVAL, "bad argument");
  launch_adamw(param, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, weight_decay, step, grad_scale, (cudaStream_t)stream);
  CK(cudaPeekAtLastError());
  return LSDM_OK;
}

LSDM_API int64_t lsdm_debug_tensor(lsdm_handle* h, const char* name, void* dst, size_t dst_bytes, void* stream) {
  if (!h || !name || !h->have_ws) return fail(LSDM_EINVAL, "bad argument");
  const Workspace& w = h->ws;
  const Workspace::Sel& q = w.sel[w.cur];
  const int64_t B = h->cfg.batch_local, C = B * NOBJ;
  struct Tap { const char* n; const void* p; int64_t count; size_t esz; };
  const Tap taps[] = {
      {"backbone", w.backbone, C * NPTS * 3, 4}, {"hm", q.hm, B * NPTS * 3, 4}, {"attn_w", q.attn_w, B * NOBJ, 4},
      {"tr", q.tr, C * TRANS, 4}, {"enc", q.enc, B * LAT, 4}, {"pa", w.pa, C * TRANS, 4}, {"pw", w.pw, C * NPTS * 3, 4},
      {"emb_cat", w.cat, B * NPTS * 256, 4}, {"pcd_out", w.pcd_out[w.cur], B * NPTS * 3, 4}, {"out_cat", q.out_cat, B * h->cfg.n_cats, 4},
      {"fps_idx0", q.idx[0], C * 1024, 4}, {"fps_idx1", q.idx[1], C * 256, 4}, {"fps_idx2", q.idx[2], C * 64, 4},
      {"fps_idx3", q.idx[3], C * 16, 4}, {"ball_idx0", q.grp[0], C * 1024 * 32, 4}, {"ball_idx1", q.grp[1], C * 256 * 32, 4},
      {"ball_idx2", q.grp[2], C * 64 * 32, 4}, {"ball_idx3", q.grp[3], C * 16 * 32, 4}, {"l1_feat", w.feat[1], C * 1024 * 64, 4},
      {"l2_feat", w.feat[2], C * 256 * 128, 4}, {"l3_feat", w.feat[3], C * 64 * 256, 4}, {"l4_feat", w.feat[4], C * 16 * 512, 4},
      {"fp4_feat", w.g3, C * 64 * 256, 4}, {"fp3_feat", w.g2, C * 256 * 256, 4}, {"fp2_feat", w.g1, C * 1024 * 128, 4},
      {"nn_idx3", q.nn_idx[3], C * 1024 * 3, 4}, {"nn_w3", q.nn_w[3], C * 1024 * 3, 4}, {"nn_idx0", q.nn_idx[0], C * 64 * 3, 4},
      {"nn_w0", q.nn_w[0], C * 64 * 3, 4}, {"nn_idx1", q.nn_idx[1], C * 256 * 3, 4}, {"nn_w1", q.nn_w[1], C * 256 * 3, 4},
      {"nn_idx2", q.nn_idx[2], C * 1024 * 3, 4}, {"nn_w2", q.nn_w[2], C * 1024 * 3, 4},
      {"x0", w.x0, B * NPTS * 3, 4}, {"guiding", w.guiding, B * NPTS * 3, 4}, {"s256", w.s256, B * 256, 4},
  };
  for (const Tap& t : taps) {
    if (strcmp(t.n, name) == 0) {
      size_t bytes = (size_t)t.count * t.esz;
      if (dst) {
        if (dst_bytes < bytes) return fail(LSDM_EINVAL, "destination too small");
        if (t.p == (const void*)w.cat && h->precision_step == 2 && g_gemm_async) {  // stored as hi + lo planes
          add_planes_kernel<<<(unsigned)((t.count + 255) / 256), 256, 0, (cudaStream_t)stream>>>(w.cat, w.cat_lo, (float*)dst, t.count);
          return t.count;
        }
        cudaError_t e = cudaMemcpyAsync(dst, t.p, bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream);
        if (e != cudaSuccess) return fail(LSDM_ECUDA, cudaGetErrorString(e));
      }
      return t.count;
    }
  }
  return fail(LSDM_EINVAL, std::string("unknown tap: ") + name);
}

LSDM_API int64_t lsdm_launch_count(const lsdm_handle* h) { return h ? h->launches : 0; }

LSDM_API const char* lsdm_profile_report(const lsdm_handle* h) { return h ? h->prof_report.c_str() : ""; }

LSDM_API int lsdm_set_option(lsdm_handle* h, const char* name, int32_t value) {
  if (!h || !name) return fail(LSDM_EINVAL, "null argument");
  if (strcmp(name, "sa_fused") == 0 && value >= 0 && value <= 3) {
    h->sa_fused = value > 0 ? 1 : 0;  // (values 1..3 named round-1 kernel variants; one fused form is left)
    return LSDM_OK;
  }
  if (strcmp(name, "cond_stream") == 0 && (value == 0 || value == 1)) {
    h->cond_stream = value;
    return LSDM_OK;
  }
  if (strcmp(name, "select_grid") == 0 && value >= 0 && value <= 15) {
    h->select_grid = value;
    return LSDM_OK;
  }
  if (strcmp(name, "loop_invariants") == 0 && value >= 0 && value <= 15) {
    h->loop_invariants = value;
    return LSDM_OK;
  }
  if (strcmp(name, "time_batch") == 0 && (value == 0 || value == 1)) {
    h->time_batch = value;
    return LSDM_OK;
  }
  if (strcmp(name, "hoist_split") == 0 && (value == 0 || value == 1)) {
    h->hoist_split = value;
    return LSDM_OK;
  }
  if (strcmp(name, "sa1_compact") == 0 && (value == 0 || value == 1)) {
    h->sa1_compact = value;
    return LSDM_OK;
  }
  if (strcmp(name, "x0_fused") == 0 && (value == 0 || value == 1)) {
    h->x0_fused = value;
    return LSDM_OK;
  }
  if (strcmp(name, "dedup_absent") == 0 && (value == 0 || value == 1)) {
    h->dedup_absent = value;
    return LSDM_OK;
  }
  if (strcmp(name, "fps_compact") == 0 && (value == 0 || value == 1)) {
    g_fps_compact = value;  // process-wide
    return LSDM_OK;
  }
  if (strcmp(name, "select_uniform") == 0 && (value == 0 || value == 1)) {
    g_select_uniform_shortcut = value;  // process-wide
    return LSDM_OK;
  }
  return fail(LSDM_EINVAL, std::string("unknown option or bad value: ") + name);
}

LSDM_API int lsdm_set_precision(lsdm_handle* h, int32_t precision_encoder, int32_t precision_step) {
  if (!h || precision_encoder < 0 || precision_encoder > 2 || precision_step < 0 || precision_step > 2)
    return fail(LSDM_EINVAL, "precision must be 0 (fp32), 1 (tf32) or 2 (3xtf32)");
  h->precision = precision_encoder;
  h->precision_step = precision_step;
  return LSDM_OK;
}

LSDM_API int lsdm_debug_gemm(lsdm_handle* h, const float* A, int64_t lda, const float* W, int64_t ldw, float* C, int64_t ldc,
                             const float* bias, int32_t bias_mode, int32_t M, int32_t N, int32_t K, int32_t act,
                             int32_t group_max, int32_t precision, void* stream) {
  if (!h || !A || !W || !C) return fail(LSDM_EINVAL, "null argument");
  GemmArgs g{};
  g.A = A; g.lda = lda; g.W = W; g.ldw = ldw; g.C = C; g.ldc = ldc;
  g.bias = bias; g.bias_mode = bias ? bias_mode : 0;
  g.M = M; g.N = N; g.K = K; g.batch = 1; g.act = act; g.group_max = group_max; g.precision = precision;
  int r = precision >= 1 ? (g_gemm_ws ? launch_gemm_ws(g, (cudaStream_t)stream) : launch_gemm_tc(g, (cudaStream_t)stream))
                         : launch_gemm_simt(g, (cudaStream_t)stream);
  if (r < 0) return fail(LSDM_EINVAL, "gemm shape not supported by the requested implementation");
  h->launches += r;
  CK(cudaPeekAtLastError());
  return LSDM_OK;
}

LSDM_API int lsdm_profile_begin(lsdm_handle* h) {
  if (!h) return fail(LSDM_EINVAL, "null handle");
  for (auto& r : h->prof) {
    cudaEventDestroy(r.a);
    cudaEventDestroy(r.b);
  }
  h->prof.clear();
  h->gemm_flops = 0.0;
  h->profiling = true;
  return LSDM_OK;
}

LSDM_API int lsdm_profile_end(lsdm_handle* h, double* ms_by_class, int64_t* launches_by_class, int32_t n_class,
                              double* gemm_flops) {
  if (!h || !ms_by_class || !launches_by_class || n_class < K_NCLASS) return fail(LSDM_EINVAL, "bad argument");
  h->profiling = false;
  for (int i = 0; i < n_class; ++i) ms_by_class[i] = 0.0, launches_by_class[i] = 0;
  struct Agg { double ms = 0, flops = 0; int64_t n = 0; };
  std::vector<std::pair<std::string, Agg>> aggs;
  for (auto& r : h->prof) {
    CK(cudaEventSynchronize(r.b));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, r.a, r.b));
    ms_by_class[r.cls] += ms;
    launches_by_class[r.cls] += 1;
    if (!r.tag.empty()) {
      Agg* a = nullptr;
      for (auto& kv : aggs)
        if (kv.first == r.tag) a = &kv.second;
      if (!a) {
        aggs.emplace_back(r.tag, Agg());
        a = &aggs.back().second;
      }
      a->ms += ms;
      a->flops += r.flops;
      a->n += 1;
    }
    cudaEventDestroy(r.a);
    cudaEventDestroy(r.b);
  }
  h->prof.clear();
  h->prof_report.clear();
  for (auto& kv : aggs) {
    char line[256];
    snprintf(line, sizeof(line), "%s\t%lld\t%.6f\t%.6e\n", kv.first.c_str(), (long long)kv.second.n, kv.second.ms, kv.second.flops);
    h->prof_report += line;
  }
  if (gemm_flops) *gemm_flops = h->gemm_flops;
  return LSDM_OK;
}

}  // extern "C"
