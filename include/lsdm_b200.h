/*
 * lsdm_b200 -- C ABI of the B200-native LSDM multi-conditional denoising path.
 *
 * The reference (andvg3/LSDM) has no plugin / operator / FFI layer: its boundary is the Python
 * surface of util/model_util.py, model/sdm.py and diffusion/{gaussian_diffusion,respace}.py.  This
 * header is the C boundary UNDER that surface: each entry point names the reference Python
 * function(s) it replaces (paths relative to the reference tree).  The Python mirror of the
 * reference API (lsdm_b200/) binds these symbols with ctypes; see INTEGRATION.md for the binding a
 * reference maintainer would add.
 *
 * Conventions
 *  - plain C types only; no torch / pybind types.  `stream` is a cudaStream_t passed as void*.
 *  - every pointer is a DEVICE pointer unless the parameter is documented "host or device"
 *    (those are copied with cudaMemcpyDefault on `stream`).
 *  - no entry point synchronises the stream or allocates caller-visible memory; the caller
 *    provides one workspace of lsdm_workspace_bytes() bytes (256-byte aligned) per handle.
 *  - return value: 0 on success, a negative LSDM_E* code on failure; lsdm_last_error() returns a
 *    thread-local message.  One stream per handle at a time; distinct handles are independent.
 *  - all floating-point tensors are fp32, row-major, contiguous.
 */
#ifndef LSDM_B200_H
#define LSDM_B200_H

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define LSDM_API __attribute__((visibility("default")))
#else
#define LSDM_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define LSDM_OK 0
#define LSDM_EINVAL (-1)   /* bad argument / shape / unknown key */
#define LSDM_ESTATE (-2)   /* call order violated (weights not finalised, conditions not encoded ...) */
#define LSDM_ECUDA (-3)    /* CUDA runtime error (message has cudaGetErrorString) */
#define LSDM_ENOMEM (-4)

#define LSDM_N_POINTS 1024 /* points per cloud (reference posa/dataset.py:456-474) */
#define LSDM_N_OBJ 9       /* object slots per scene, slot 0 = human */
#define LSDM_CLIP_DIM 512  /* text-embedding width fed to embed_text (util/model_util.py:31) */

typedef struct lsdm_handle lsdm_handle;

typedef struct lsdm_config {
  int32_t batch_local;  /* samples this handle processes per call (<= batch_global) */
  int32_t batch_global; /* global batch B: the scrambles of model/sdm.py:180-182,199-201 index the mask by it */
  int32_t batch_offset; /* global index of local sample 0 (data-parallel shard offset) */
  int32_t n_cats;       /* 13 (proxd) or 11 (humanise): util/model_util.py:26-73 */
  int32_t device;       /* CUDA device ordinal */
  int32_t reserved;
} lsdm_config;

/* Library / build identification ("lsdm_b200 <ver> sm_100a"). */
LSDM_API const char* lsdm_version(void);
LSDM_API const char* lsdm_last_error(void);

/* Replaces SceneDiffusionModel.__init__ + .to(device) (model/sdm.py:19-129). */
LSDM_API int lsdm_create(lsdm_handle** out, const lsdm_config* cfg);
LSDM_API void lsdm_destroy(lsdm_handle* h);
/* Re-shard an existing handle (same weights): changes batch_local/global/offset only. */
LSDM_API int lsdm_set_batch(lsdm_handle* h, int32_t batch_local, int32_t batch_global, int32_t batch_offset);

/* Replaces nn.Module.load_state_dict (run/test_sdm.py:123-124): one call per state-dict entry, with
 * the reference's exact key and shape.  `data`: host or device, fp32
 * (num_batches_tracked entries are accepted and ignored).  Unknown `clip_model.*` keys are ignored. */
LSDM_API int lsdm_load_weight(lsdm_handle* h, const char* key, const void* data, const int64_t* shape, int32_t ndim,
                     void* stream);
/* Number of state-dict entries the handle expects / i-th expected key (for the Python mirror). */
LSDM_API int lsdm_num_weights(const lsdm_handle* h);
LSDM_API const char* lsdm_weight_key(const lsdm_handle* h, int i);
/* Folds eval-mode BatchNorm into the 1x1 convs, splits the first SA/FP layers into their
 * coordinate / feature halves.  Must follow the last lsdm_load_weight. */
LSDM_API int lsdm_finalize_weights(lsdm_handle* h, void* stream);

/* Replaces GaussianDiffusion.__init__'s tables + _extract_into_tensor (diffusion/gaussian_diffusion.py:
 * 166-202,1585-1598): fp32 casts of the float64 tables, length T each (host or device). */
LSDM_API int lsdm_set_schedule(lsdm_handle* h, const float* posterior_mean_coef1, const float* posterior_mean_coef2,
                      const float* posterior_log_variance_clipped, const float* sqrt_alphas_cumprod,
                      const float* sqrt_one_minus_alphas_cumprod, int32_t T, void* stream);

LSDM_API size_t lsdm_workspace_bytes(const lsdm_handle* h);
LSDM_API int lsdm_set_workspace(lsdm_handle* h, void* workspace, size_t bytes);

/* Replaces the x/t-independent part of SceneDiffusionModel.forward (model/sdm.py:147-203): text / category
 * MLPs, object attention weights, PointNet++ on the 9 clouds, POSA human decoder, collapsed point attention,
 * pointwise translation, masked sum.  Leaves pcd_out / enc / out_cat in the workspace.
 *   text_emb[Bl,512]  given_objs[Bl,9,1024,3]  given_cats[Bl,9,C]  mask_global[Bg,9]
 *   fps_start[4][9*Bl] int64 (host or device): the four torch.randint(0,N,(9B,)) draws of
 *   farthest_point_sample (model/pcd_backbone/pointnet2_utils.py:72), LOCAL slice, N = 1024,1024,256,64. */
LSDM_API int lsdm_encode_conditions(lsdm_handle* h, const float* text_emb, const float* given_objs, const float* given_cats,
                           const float* mask_global, const int64_t* fps_start, void* stream);

/* model.train() variant of lsdm_encode_conditions (reference run/train_sdm.py:30-107 calls the model in train mode): the 22
 * BatchNorm layers of the backbone use batch statistics over all 9*Bl clouds and update running_mean / running_var inside the
 * handle (momentum 0.1, unbiased variance; read them back with lsdm_read_weight), and the backbone head applies
 * drop_mask[9*Bl,128,1024] (0 or 2: Dropout(0.5), reference model/pcd_backbone/pointnet2.py:76, in the reference's layout). */
LSDM_API int lsdm_encode_conditions_train(lsdm_handle* h, const float* text_emb, const float* given_objs, const float* given_cats,
                                          const float* mask_global, const int64_t* fps_start, const float* drop_mask, void* stream);

/* SyncBN hook for data-parallel train-mode forwards: called (on the host, while kernels are being enqueued) once per BatchNorm
 * layer with a DEVICE buffer of n doubles (per-channel sums and sums of squares of this shard); it must enqueue, on the stream
 * passed to lsdm_encode_conditions_train, an in-place SUM all-reduce over the shards and return 0.  Shards must be equal-sized
 * (batch_global / batch_local of them).  NULL disables it (per-shard statistics). */
typedef int (*lsdm_allreduce_fn)(void* ctx, double* device_buf, int32_t n);
LSDM_API int lsdm_set_allreduce(lsdm_handle* h, lsdm_allreduce_fn fn, void* ctx);

/* Current value of a state-dict entry held by the handle (e.g. BatchNorm running statistics after a train-mode forward);
 * dst host or device, numel must match. */
LSDM_API int lsdm_read_weight(lsdm_handle* h, const char* key, float* dst, int64_t numel, void* stream);

/* Replaces one GaussianDiffusion.p_sample (diffusion/gaussian_diffusion.py:501-561) given encoded conditions:
 * timestep embedding, upsampler, x += pcd_out (IN PLACE, model/sdm.py:204), Input/OutputProcess, posterior
 * mean with the mutated x, ancestral noise.  t[Bl] int64 (host or device) indexes the schedule tables AND the
 * positional table.  x[Bl,1024,3] in-out; noise[Bl,1024,3]; sample_out[Bl,1024,3] (may alias x);
 * x0_out, guiding_out nullable.  clip_denoised != 0 clamps x0 to [-1,1] before the posterior mean
 * (process_xstart, gaussian_diffusion.py:359-365; every reference caller passes False). */
LSDM_API int lsdm_denoise_step(lsdm_handle* h, float* x, const int64_t* t, const float* noise, float* sample_out, float* x0_out,
                      float* guiding_out, int32_t clip_denoised, void* stream);

/* Replaces SceneDiffusionModel.forward given encoded conditions (model/sdm.py:131-218): like lsdm_denoise_step
 * without the posterior.  x is mutated in place; out_cat[Bl,C], x0[Bl,1024,3], guiding[Bl,1024,3] (nullable). */
LSDM_API int lsdm_forward(lsdm_handle* h, float* x, const int64_t* t, float* out_cat, float* x0, float* guiding, void* stream);

/* Replaces the loop body of p_sample_loop_progressive (diffusion/gaussian_diffusion.py:736-759) for `n_steps`
 * consecutive timesteps t_first, t_first-1, ...: STRICT mode re-encodes the conditions every step from
 * fps_start_all[n_steps][4][9*Bl] (what the reference does); hoisted != 0 encodes once.
 * noise_all[n_steps][Bl,1024,3].  x in-out.  All pointers device. */
LSDM_API int lsdm_sample_loop(lsdm_handle* h, float* x, const float* text_emb, const float* given_objs, const float* given_cats,
                     const float* mask_global, const int64_t* fps_start_all, const float* noise_all, int32_t t_first,
                     int32_t n_steps, int32_t hoisted, int32_t clip_denoised, float* x0_out, float* guiding_out, void* stream);

/* Copies of cached per-loop outputs: out_cat[Bl,C] (model.saved_cat), pcd_out[Bl,1024,3]. */
LSDM_API int lsdm_get_out_cat(lsdm_handle* h, float* out_cat, void* stream);
LSDM_API int lsdm_get_pcd_out(lsdm_handle* h, float* pcd_out, void* stream);

/* Replaces q_sample (diffusion/gaussian_diffusion.py:238-256): x_t = sqrt(abar_t) x0 + sqrt(1-abar_t) noise. */
LSDM_API int lsdm_q_sample(lsdm_handle* h, const float* x_start, const int64_t* t, const float* noise, float* x_t, void* stream);

/* Replaces pytorch3d.loss.chamfer_distance with default arguments as called at diffusion/gaussian_diffusion.py:1334:
 * sums[0] += sum_b mean_p min_q |x-y|^2, sums[1] += the y->x direction (caller zeroes sums, divides by B). */
LSDM_API int lsdm_chamfer(lsdm_handle* h, const float* x, const float* y, int32_t batch, int32_t n, int32_t m, float* sums,
                 void* stream);

/* Replaces the categorical term of training_losses (diffusion/gaussian_diffusion.py:1297-1301):
 * sum_b CrossEntropy(probs[b,:] treated as logits, argmax(target_cat[b,:])) accumulated into *sum. */
LSDM_API int lsdm_cat_loss(lsdm_handle* h, const float* probs, const float* target_cat, int32_t batch, float* sum, void* stream);

/* ---- Training backward.  Replaces `mp_trainer.backward(loss)` of run/train_sdm.py:78-84 (autograd through
 * diffusion/gaussian_diffusion.py:1256-1342 and model/sdm.py:131-218 in model.train() mode) and the AdamW step of
 * diffusion/fp16_util.py:198-214 / run/train_sdm.py:268.
 *
 * lsdm_training_backward runs a taped fp32 forward of training_losses (BatchNorm batch statistics -- all-reduced through the
 * lsdm_set_allreduce hook when sharded --, the caller's q_sample noise, FPS starts and Dropout mask) and the reverse sweep, and
 * ACCUMULATES the gradient of  g_mse * chamfer + g_cat * lambda_cat * CE  into `grads`: a float buffer of lsdm_grad_floats()
 * elements in which state-dict entry i occupies [offset, offset + numel) as reported by lsdm_weight_slot() (same order and
 * keys as lsdm_weight_key; BatchNorm running statistics and the dead attn_layer v_proj / out_proj slots stay untouched).
 * `tape`: device scratch of lsdm_train_tape_bytes() bytes.  losses_out (device, nullable): [chamfer, mean cross-entropy].
 * x0_out (device, nullable): the model output [Bl,1024,3].  All pointers device except fps_start (host or device). */
LSDM_API size_t lsdm_train_tape_bytes(const lsdm_handle* h);
LSDM_API int64_t lsdm_grad_floats(const lsdm_handle* h);
LSDM_API int lsdm_weight_slot(const lsdm_handle* h, int32_t i, int64_t* offset, int64_t* numel);
LSDM_API int lsdm_training_backward(lsdm_handle* h, const float* x_start, const int64_t* t, const float* noise, const float* text_emb,
                                    const float* given_objs, const float* given_cats, const float* mask_global, const float* target_cat,
                                    const int64_t* fps_start, const float* drop_mask, float lambda_cat, float g_mse, float g_cat, void* tape,
                                    size_t tape_bytes, float* grads, float* losses_out, float* x0_out, void* stream);
/* torch.optim.AdamW update of one flat tensor (decoupled weight decay, bias correction with `step` >= 1; grad is multiplied by
 * grad_scale first, e.g. 1/world_size after a summing all-reduce). */
LSDM_API int lsdm_adamw_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr, float beta1,
                             float beta2, float eps, float weight_decay, int64_t step, float grad_scale, void* stream);

/* ---- Evaluation metrics of the sampling path (reference run/test_sdm.py:186-207).  Handle-free. ----
 *
 * lsdm_eval_emd replaces util/evaluation.py:5-11 `emd(x, y)` = scipy cdist + linear_sum_assignment: for each of `batch`
 * cloud pairs x[b,n,3], y[b,m,3] (n == m <= 1024) the mean Euclidean distance of the minimum-cost perfect matching, as a
 * double in emd[batch].  Device algorithm: epsilon-scaled forward auction on integer-scaled distances (deterministic); the
 * matching's cost is within 2^-26 x (cloud extent) per point of the optimum.  assignment[batch,n] (nullable) receives the
 * matched y index of every x point, rounds[batch] (nullable) the number of bidding rounds.  A sample whose auction
 * does not terminate within the round cap reports NaN. */
LSDM_API int lsdm_eval_emd(const float* x, const float* y, int32_t batch, int32_t n, int32_t m, double* emd, int32_t* assignment,
                           int32_t* rounds, void* stream);

/* Replaces util/evaluation.py:28-52 `calculate_fscore(gt, pr, th)` (open3d nearest-neighbour distances both ways, double):
 * out[batch,3] = {fscore, precision, recall}; counts[batch,2] is int32 scratch.  n, m <= 4096. */
LSDM_API int lsdm_eval_fscore(const float* gt, const float* pr, int32_t batch, int32_t n, int32_t m, double th, int32_t* counts,
                              double* out, void* stream);

/* pytorch3d.loss.chamfer_distance as called per sample at run/test_sdm.py:187: per_sample[batch,2] = the two directed
 * point-mean squared-NN terms of each pair (their sum is the reference's `loss` at batch 1). */
LSDM_API int lsdm_eval_chamfer(const float* x, const float* y, int32_t batch, int32_t n, int32_t m, float* per_sample, void* stream);

/* Replaces util/evaluation.py:13-26 `accuracy(output, target, topk)`: correct[k] = number of samples whose target class
 * (int64) is among the ks[k] highest of scores[batch,n_classes] (ks: device int32[nk]; ties ordered by class index). */
LSDM_API int lsdm_eval_topk(const float* scores, const int64_t* target, int32_t batch, int32_t n_classes, const int32_t* ks, int32_t nk,
                            int32_t* correct, void* stream);

/* ---- CLIP text tower ------------------------------------------------------------------
 * Replaces `clip_model.encode_text(tokens).float()` of model/sdm.py:245-259 (openai/CLIP `CLIP.encode_text`, ViT-B/32 text
 * side: 12 pre-LN causal transformer blocks of width 512, 8 heads, QuickGELU MLP, ln_final, EOT-token feature times
 * text_projection).  Tokenisation (clip.tokenize: BPE on the host) stays with the caller: `tokens` is its output.
 *
 * A separate handle: the tower is frozen and independent of the denoiser's weights.  Weight keys are the openai/CLIP
 * state-dict names WITHOUT the reference's `clip_model.` prefix: token_embedding.weight [V,W], positional_embedding [ctx,W],
 * transformer.resblocks.<l>.{ln_1,ln_2}.{weight,bias}, .attn.in_proj_{weight,bias}, .attn.out_proj.{weight,bias},
 * .mlp.c_fc.{weight,bias}, .mlp.c_proj.{weight,bias}, ln_final.{weight,bias}, text_projection [W,E]; fp32, host or device
 * pointers.  The configuration (width, layers, heads = W/64, context, vocabulary, embed dim) is inferred from the shapes at
 * lsdm_clip_finalize, which fails if any tensor is missing.  The handle owns its weight copies (cudaMalloc). */
typedef struct lsdm_clip lsdm_clip;
LSDM_API int lsdm_clip_create(lsdm_clip** out, int32_t device);
LSDM_API void lsdm_clip_destroy(lsdm_clip* h);
LSDM_API int lsdm_clip_load_weight(lsdm_clip* h, const char* key, const float* data, const int64_t* shape, int32_t ndim, void* stream);
LSDM_API int lsdm_clip_finalize(lsdm_clip* h);
LSDM_API int lsdm_clip_dims(const lsdm_clip* h, int32_t* width, int32_t* layers, int32_t* heads, int32_t* ctx, int32_t* vocab, int32_t* embed);
/* Dense-layer arithmetic: 0 fp32 CUDA cores, 1 TF32 tcgen05, 2 3xTF32 tcgen05 (default; fp32-grade accuracy). */
LSDM_API int lsdm_clip_set_precision(lsdm_clip* h, int32_t precision);
LSDM_API size_t lsdm_clip_workspace_bytes(const lsdm_clip* h, int32_t batch, int32_t seq_len);
/* tokens: device int32 [batch, ctx] (zero-padded after the EOT token, which carries the largest id: clip.tokenize's layout);
 * out: device fp32 [batch, embed].  Only positions [0, seq_len) are computed -- attention is causal, so this is bit-identical
 * to the full context as long as every sample's EOT position is < seq_len (a sample that violates this gets NaN).
 * Enqueues on `stream`, no synchronisation; workspace (256-byte aligned) of lsdm_clip_workspace_bytes(batch, seq_len). */
LSDM_API int lsdm_clip_encode_text(lsdm_clip* h, const int32_t* tokens, int32_t batch, int32_t seq_len, void* workspace, size_t workspace_bytes,
                                   float* out, void* stream);
LSDM_API int64_t lsdm_clip_launch_count(const lsdm_clip* h);

/* ---- ContactFormer temporal attention layer -------------------------------------------------------
 * lsdm_cf_mha_forward replaces contact_former/transformer.py:72-103 `MultiHeadAttention.forward(x, mask)` in eval mode
 * (both Dropouts off): x[bs, seg_len, n_verts, 64] -> LayerNorm(fc(softmax(q k^T / 8 [mask]) v) + x), attention along the
 * seg_len axis independently per (head, vertex, sample); d_in = d_k = d_v = 64 (contact_former.py:271-275,318), seg_len <= 256,
 * n_head * 64 in {64, 128, 192, k * 256}.  Weights are the module's tensors as stored (row-major [out, in]): w_q / w_k /
 * w_v [n_head*64, 64] + biases, fc [64, n_head*64] + bias, layer_norm weight / bias [64].
 * mask (nullable): uint8 [bs, seg_len, seg_len], entries equal to 0 are filled with -inf (transformer.py:87-89; a fully
 * masked row yields NaN exactly as torch.softmax does); all_masked = the reference's `mask.sum() == 0` branch (:91-92),
 * decided by the caller.  precision: 0 fp32 CUDA cores, 1 TF32 tcgen05, 2 3xTF32 tcgen05 for the four projections.
 * Workspace: lsdm_cf_workspace_bytes(bs, seg_len, n_verts, n_head), 256-byte aligned.  `out` may not alias `x`.
 * lsdm_cf_ffn_forward replaces transformer.py:167-177 `PositionwiseFeedForward.forward`: LayerNorm(w_2 relu(w_1 x) + x) on
 * rows x 64 (the two 1x1 Conv1d weights as [d_hid, 64] and [64, d_hid]); workspace rows * (d_hid + 64) floats + 512 bytes. */
typedef struct lsdm_cf_mha_weights {
  const float *w_q, *b_q, *w_k, *b_k, *w_v, *b_v, *fc_w, *fc_b, *ln_w, *ln_b;
} lsdm_cf_mha_weights;
typedef struct lsdm_cf_ffn_weights {
  const float *w1, *b1, *w2, *b2, *ln_w, *ln_b;
  int32_t d_hid;
  int32_t reserved;
} lsdm_cf_ffn_weights;
/* Process-wide knob: "attn_tc" = 1 (default) runs softmax(QK^T)V on the tensor cores (tcgen05 TF32, scores kept in tensor
 * memory) when precision == 1 and seg_len is a multiple of 32; 0 forces the fp32 CUDA-core kernel that serves every other case
 * (precision 0 / 2 always use it: their contract is fp32-grade accuracy). */
LSDM_API int lsdm_cf_set_option(const char* name, int32_t value);
LSDM_API size_t lsdm_cf_workspace_bytes(int32_t bs, int32_t seg_len, int32_t n_verts, int32_t n_head);
LSDM_API int lsdm_cf_mha_forward(const lsdm_cf_mha_weights* w, const float* x, const uint8_t* mask, int32_t all_masked, int32_t bs, int32_t seg_len,
                                 int32_t n_verts, int32_t n_head, int32_t precision, void* workspace, size_t workspace_bytes, float* out,
                                 void* stream);
LSDM_API int lsdm_cf_ffn_forward(const lsdm_cf_ffn_weights* w, const float* x, int64_t rows, int32_t precision, void* workspace, size_t workspace_bytes,
                                 float* out, void* stream);

/* Debug / parity taps: copy a named intermediate of the last encode/forward into `dst` (device).
 * Returns the element count, or a negative error.  Names: "backbone" [9Bl,1024,3], "hm" [Bl,1024,3],
 * "attn_w" [Bl,9], "tr" [Bl,9,12], "enc" [Bl,128], "pa" [Bl,9,12], "pw" [Bl,9,1024,3], "emb" [Bl,1024,128],
 * "fps_idx0..3", "ball_idx0..3" (int32), "l1_feat".. */
LSDM_API int64_t lsdm_debug_tensor(lsdm_handle* h, const char* name, void* dst, size_t dst_bytes, void* stream);

/* Number of kernels this library has launched on behalf of `h` since creation (bench.py's gpu_launches). */
LSDM_API int64_t lsdm_launch_count(const lsdm_handle* h);

/* Arithmetic of the dense layers: 0 = fp32 CUDA cores; 1 = TF32 tensor cores (tcgen05, operands rounded to nearest,
 * fp32 accumulate in TMEM); 2 = 3xTF32 (hi/lo split of both operands on the tensor cores, ~fp32 accuracy).  Set separately
 * for the condition encoder (PointNet++ etc., the bulk of the FLOPs) and for the per-step x0 network + upsampler (small,
 * accuracy-critical).  Layers whose shape the tensor path does not support, and all selection kernels (FPS, ball query,
 * 3-NN), are always exact fp32. */
LSDM_API int lsdm_set_precision(lsdm_handle* h, int32_t precision_encoder, int32_t precision_step);

/* Switches that exist for the on / off equality tests and bench.py's transparency legs (every "on" form is tested to give the
 * same result as its "off" form); defaults in brackets.
 * "sa_fused"      [1] 0 = set-abstraction levels 1-3 as gather + one GEMM launch per layer, > 0 = the fused tensor-core kernels.
 * "sa1_compact"   [1] sa1 runs on the distinct rows of every ball-query group only (bit-identical, ~6x fewer tiles).
 * "x0_fused"      [1] the x0 network of a step runs as one persistent kernel; 0 = one GEMM per layer.
 * "hoist_split"   [1] hoisted loop only: time half of the embedding once per step for the whole batch, text half once per loop.
 * "dedup_absent"  [1] lsdm_sample_loop encodes the all-zero cloud of absent objects once per step and shares the result.
 * "select_uniform" (process-wide) [1] a cloud whose points all coincide takes the closed-form FPS order {start, 0, 0, ...} and a
 *                 3-candidate 3-NN scan: the same integers as the full scans, which 0 forces.
 * "select_grid"   [9] bit mask: ball query of level 0 (1) / level 1 (2), 3-NN of fp2 (4) / fp1 (8) through a per-cloud cell grid
 *                 instead of the full scan (identical groups, indices and weights).
 * "cond_stream"   [1] pipelined loop: the per-sample condition MLPs and the human decoder run on their own stream beside the
 *                 selection chain.
 * "loop_invariants" [15] lsdm_sample_loop, STRICT: bit mask of what is computed once per call instead of once per step because
 *                 nothing it reads changes over the loop (reference p_sample_loop: conditions fixed, t shared by the batch):
 *                 1 condition MLPs + human decoder, 2 text half of the embedding once per call / time half once per step for the
 *                 whole batch (split 256-term sums: 1e-8-level differences; every other bit is bit-identical), 4 sa1, the level-0
 *                 ball query and sa2's projected rows in cloud order (sa1 keeps all points: the level-0 FPS only orders them),
 *                 8 guiding points on the call's last step only.  0 = everything every step.
 * "time_batch"    [1] the time half of the split embedding is evaluated for 16 steps per launch sequence (bit-identical).
 * "fps_compact"   (process-wide) [1] FPS level 0 drops the points at distance 0 every 128 rounds (identical selection order). */
LSDM_API int lsdm_set_option(lsdm_handle* h, const char* name, int32_t value);

/* Test hook: one linear layer C[M,N] = act(A[M,K] W[N,K]^T + bias) through the fp32 (precision 0) or tcgen05 TF32 / 3xTF32
 * (precision 1 / 2) GEMM.  act: 0 none, 1 relu, 2 gelu, 3 sigmoid, 4 silu.  group_max: C[M/32,N] = max over 32-row groups. */
LSDM_API int lsdm_debug_gemm(lsdm_handle* h, const float* A, int64_t lda, const float* W, int64_t ldw, float* C, int64_t ldc,
                             const float* bias, int32_t bias_mode, int32_t M, int32_t N, int32_t K, int32_t act,
                             int32_t group_max, int32_t precision, void* stream);

/* Per-kernel-class timing with CUDA events on the launching stream (measurement aid for bench.py; adds two event
 * records per launch, so it is used in a separate pass, never inside the timed region).  Classes, in order:
 * gemm, fps, ball_query, sa_gather, three_nn, fp_combine, head3, cond, scene, denoise, other (LSDM_N_KCLASS = 11).
 * lsdm_profile_end synchronises the recorded events; gemm_flops = sum of 2*M*N*K over the profiled GEMM launches. */
#define LSDM_N_KCLASS 11
LSDM_API int lsdm_profile_begin(lsdm_handle* h);
LSDM_API int lsdm_profile_end(lsdm_handle* h, double* ms_by_class, int64_t* launches_by_class, int32_t n_class,
                              double* gemm_flops);

/* Per-kernel detail of the last profiled pass for the tensor-core class: lines "tag\tlaunches\tms\tflops\n". */
LSDM_API const char* lsdm_profile_report(const lsdm_handle* h);

#ifdef __cplusplus
}
#endif
#endif /* LSDM_B200_H */
