// Launch prototypes of every non-GEMM kernel on the path.  Each launcher enqueues on `st` and
// returns the number of kernels launched.
#pragma once
#include <string>
#include "common.cuh"
#include "gemm.cuh"

namespace lsdm {

// ---- pointnet_select.cu ---------------------------------------------------------------------
int launch_fps4(const float* xyz0, const int64_t* start, int n_clouds, int* idx1, int* idx2, int* idx3, int* idx4,
                float* xyz1, float* xyz2, float* xyz3, float* xyz4, cudaStream_t st);
int launch_ball_query(const float* xyz, const float* new_xyz, int n_clouds, int N, int S, double radius, int* group,
                      cudaStream_t st, const int* compose = nullptr);
int launch_three_nn(const float* xyz1, const float* xyz2, int n_clouds, int N, int S, int* nn_idx, float* nn_w,
                    cudaStream_t st);

// ---- pointnet_grid.cu: cell-grid forms of the two scans for 1024 source points (identical results) ----
int launch_ball_query_grid(const float* xyz, const float* new_xyz, int n_clouds, int N, int S, double radius, int* group, cudaStream_t st,
                           int* plan_rows = nullptr, int* plan_used = nullptr, int* plan_tiles = nullptr);
int launch_three_nn_grid(const float* xyz1, const float* xyz2, int n_clouds, int N, int S, int* nn_idx, float* nn_w, cudaStream_t st);

// ---- pointnet_glue.cu -----------------------------------------------------------------------
// h1[(c,s,k), ch] = relu(P[c*N + j, ch] + Wx[ch,:] . (xyz[c,j] - new_xyz[c,s])),  j = group[c,s,k]
// P == nullptr (sa1): P is replaced by bias[ch] + Wf[ch,0:3] . xyz[c,j]   (features are the coordinates)
int launch_sa_gather(const float* P, const float* Wx, const float* Wf3, const float* bias, const float* xyz,
                     const float* new_xyz, const int* group, int n_clouds, int N, int S, int C1, float* h1,
                     int round_out, cudaStream_t st, int apply_relu = 1);
// h[(c,n), ch] = relu(Pa[(c,n), ch] + sum_k w[c,n,k] * Pb[c*S + idx[c,n,k], ch]);  Pa == nullptr -> bias[ch]
int launch_fp_combine(const float* Pa, const float* bias, const float* Pb, const int* nn_idx, const float* nn_w,
                      int n_clouds, int N, int S, int C1, float* h, int round_out, cudaStream_t st, int apply_relu = 1);
// out[row, 0:3] = h[row, 0:128] . W[3,128]^T + b
int launch_head3(const float* h, const float* W, const float* b, int64_t rows, float* out, cudaStream_t st);
// eval-mode BatchNorm folded into the preceding 1x1 conv: Wf = W * s, bf = (b - mean) * s + beta, s = gamma / sqrt(var + eps)
int launch_fold_bn(const float* W, const float* b, const float* gamma, const float* beta, const float* mean,
                   const float* var, int N, int K, float eps, float* Wf, float* bf, cudaStream_t st);
int launch_copy_cols(const float* src, int ld_src, int col0, int ncols, int rows, float* dst, cudaStream_t st);

// ---- sa_fused.cu ----------------------------------------------------------------------------
// Fused set-abstraction block on the tensor cores (gather + 2 dense layers + max-pool), levels 0..2.
// Returns -1 when the (level, variant) combination is not built.
int launch_sa_fused(int level, const float* P, const float* xyz, const float* new_xyz, const int* grp,
                    const float* Wx, const float* Wf3, const float* b1, const float* W2, const float* b2, const float* W3,
                    const float* b3, int n_clouds, int N, int S, float* out, int round_out, cudaStream_t st);

int launch_sa_fused_v2(int level, const float* P, const float* xyz, const float* new_xyz, const int* grp, const float* h_wx,
                       const float* h_wf, const float* h_b1, const float* h_b2, const float* d_wx, const float* d_b1, const float* W2,
                       const float* W3, const float* b3, int n_clouds, int N, int S, float* out, int round_out, cudaStream_t st);

// sa1 on the DISTINCT rows of every group only (the ball query pads with the first hit; the max-pool ignores duplicates):
// plan (selection stream) + the fused kernel on the packed tiles.  Bit-identical to launch_sa_fused_v2(level 0).
int launch_sa1_plan(const int* grp, int n_clouds, int* rows, int* tile_used, int* tiles, int* tile_off, int* n_tiles, cudaStream_t st);
int launch_sa1_plan_scan(const int* tiles, int n_clouds, int* tile_off, int* n_tiles, cudaStream_t st);  // when the ball query wrote the plan
int launch_sa1_compact(const float* xyz, const float* new_xyz, const int* rows, const int* tile_used, const int* tile_off, const int* n_tiles,
                       const float* h_wx, const float* h_wf, const float* h_b1, const float* h_b2, const float* W2, const float* W3, const float* b3,
                       int n_clouds, float* out, int round_out, cudaStream_t st);

// fused feature-propagation level (fp2): first conv (fine half on the tensor core + interpolated coarse projection) + second conv
int launch_fp_fused(const float* X, int CA, const float* Wa, const float* ba, const float* Pb, const int* nn_idx, const float* nn_w,
                    const float* W1, const float* b1, int n_clouds, int N, int S, int C1, int C2, float* out, int round_out,
                    cudaStream_t st, const int* x_perm = nullptr);

// fused last level + head: fp1 (interpolation gathered in-kernel, 3 x 128x128 convs) + conv2; h_b1 / h_consts are HOST arrays
int launch_fp1_fused(const float* Pb, const int* nn_idx, const float* nn_w, const float* h_b1, const float* W2, const float* W3,
                     const float* Wh, const float* h_consts, int n_clouds, int N, int S, float* out, cudaStream_t st);

// ---- x0net_fused.cu ---------------------------------------------------------------------------
// the x0 network of one step (3 -> 64 -> 128 || emb -> 192 -> 128 -> 64 -> 3, 3xTF32) + in-place x += pcd_out + posterior mean +
// ancestral noise, per 128-row tile; n_pass == 2 also runs the guiding-points pass.  Returns -1 if TMA descriptors cannot be built.
int launch_x0net_fused(float* x, const float* pcd_out, const float* noise, float* sample_out, float* x0_out, float* guiding_out,
                       const int64_t* t, const float* c1, const float* c2, const float* logvar, int rows, int n_pass, int clip,
                       const float* emb_hi, const float* emb_lo, int64_t ld_emb, const float* w1, const float* w1_lo, const float* w2,
                       const float* w2_lo, const float* w3, const float* w3_lo, const float* w4, const float* w4_lo, const float* w0,
                       const float* b0, const float* b1, const float* b2, const float* b3, const float* b4, const float* w5,
                       const float* b5, cudaStream_t st);

// ---- bn_train.cu ------------------------------------------------------------------------------
// train-mode BatchNorm: column statistics (double), then normalise + ReLU [+ dropout mask] [+ 32-row max-pool] and the
// running-statistics update (momentum 0.1, unbiased variance)
int launch_col_stats(const float* y, int64_t M, int N, double* sums, cudaStream_t st);
int launch_bn_apply(float* y, int64_t M, int64_t M_stat, int N, const double* sums, const float* gamma, const float* beta, float* run_mean,
                    float* run_var, const float* drop_mask, int mask_points, float* pooled, int round_out, cudaStream_t st);

// ---- cond.cu --------------------------------------------------------------------------------
struct CondWeights {
  const float *et0_w, *et0_b, *et2_w, *et2_b, *et4_w, *et4_b;  // embed_text 512->256->256->128
  const float *pc0_w, *pc0_b, *pc2_w, *pc2_b, *pc4_w, *pc4_b;  // predict_cat 128->64->32->C
  const float *ec_w, *ec_b;                                    // embed_cat C->32
  const float *aq_w, *ak_w, *a_inb;                            // attn_layer q_proj [128,128], k_proj [128,32], in_proj_bias [384]
  const float *tl0_w, *tl0_b, *tl2_w, *tl2_b;                  // translation_layer 160->128->12
  const float *pq_w, *p_inb;                                   // pcd_attention q_proj [12,12], in_proj_bias [36]
};
// one CTA per local sample: enc[B,128], out_cat[B,C], attn_w[B,9], tr[B,9,12], qq[B,9,12]
int launch_cond(const CondWeights& w, const float* text, const float* cats, const float* mask_global, int B, int Bg,
                int b_off, int n_cats, float* enc, float* out_cat, float* attn_w, float* tr, float* qq, cudaStream_t st);
// ts[B,128] = W2 silu(W1 pe[t] + b1) + b2 ; also writes s[B,256] = [ts || enc] and H1[B*256,128] = gelu(s*w0 + b0)
int launch_time_embed(const float* pe, const float* w1, const float* b1, const float* w2, const float* b2,
                      const int64_t* t, const float* enc, const float* up0_w, const float* up0_b, int B, float* s256,
                      float* H1, float* H1_lo /* nullable: 3xTF32 residual plane, H1 then holds the hi plane */, cudaStream_t st);
struct HumanWeights {
  const float *w0, *b0, *g0, *be0;  // 3->64 + GroupNorm(8)
  const float *w1, *b1, *g1, *be1;  // 64->64
  const float *w2, *b2, *g2, *be2;  // 64->64 on the first 655 points
  const float *w3, *b3;             // 64->3
};
// POSA decoder (4 launches); scratch: y0[B,1024,64] | y1[B,1024,64] | GroupNorm stats [B,3,8,2] doubles; hm[B,1024,3]
int launch_human(const HumanWeights& w, const float* objs /*[B,9,1024,3], slot 0 used*/, int B, float* scratch,
                 float* hm, cudaStream_t st);

// ---- scene.cu -------------------------------------------------------------------------------
struct SceneWeights {
  const float *pk_w, *pv_w, *p_inb;  // pcd_attention k_proj [12,3], v_proj [12,3], in_proj_bias [36]
  const float *po_w, *po_b;          // out_proj [12,12]
  const float *pt_w, *pt_b;          // point_wise_trans_layer [3,15]
};
// per (b,o'): scramble #1, collapsed 12-head attention, pointwise 15->3 GELU -> pw[B,9,1024,3], pa[B,9,12]
int launch_point_attention(const SceneWeights& w, const float* backbone /*[9B,1024,3]*/, const float* attn_w,
                           const float* qq, int B, float* pa, float* pw, cudaStream_t st, const int* remap = nullptr);
// scramble #2 (global mask), sum over objects, average with human feature -> pcd_out[B,1024,3]
int launch_scene_mix(const float* pw, const float* hm, const float* mask_global, int B, int Bg, int b_off, float* pcd_out,
                     cudaStream_t st);

// ---- denoise.cu -----------------------------------------------------------------------------
// xin = x (+ pcd_out, written back to x when add != nullptr); h1[row,64] = sigmoid(E0 xin + b)
int launch_pose_embed0(float* x, const float* add, const float* w, const float* b, int64_t rows, float* h1,
                       float* h1_lo /* nullable, as above */, cudaStream_t st);
// x0 = gelu(F2 f1 + b) (64->3); optional posterior + ancestral sample (gaussian_diffusion.py:266-269,545-560):
// sample = c1[t] x0 + c2[t] xin + [t != 0] exp(0.5 logvar[t]) noise
int launch_final3(const float* f1, const float* w, const float* b, int64_t rows, float* x0_out, const float* xin,
                  const int64_t* t, const float* c1, const float* c2, const float* logvar, const float* noise,
                  float* sample_out, int clip_denoised, cudaStream_t st);
int launch_q_sample(const float* x0, const int64_t* t, const float* noise, const float* sa, const float* s1a, int B,
                    float* xt, cudaStream_t st);

// ---- loss.cu --------------------------------------------------------------------------------
int launch_chamfer(const float* x, const float* y, int B, int n, int m, float* sums, cudaStream_t st, int sample_stride = 0);
int launch_cat_loss(const float* probs, const float* target, int B, int C, float* sum, cudaStream_t st);

extern int g_fps_compact;  // pointnet_select.cu: FPS level 0 drops the points at distance 0 every 128 rounds (same selection order)
extern int g_select_uniform_shortcut;  // pointnet_select.cu: closed-form selections for clouds whose points all coincide

// ---- api.cu: thread-local error message behind lsdm_last_error(); returns `code` ----
int set_error(int code, const std::string& msg);

// ---- eval_metrics.cu ------------------------------------------------------------------------
// EMD (min-cost perfect matching, auction algorithm) of x[B,n,3] vs y[B,n,3], n <= 1024: out_emd[B] (double) = mean matched distance
int launch_emd(const float* x, const float* y, int B, int n, double* out_emd, int* out_assign /*nullable [B,n]*/,
               int* out_rounds /*nullable [B]*/, cudaStream_t st);
// out[B,3] (double) = {fscore, precision, recall} at threshold th; counts[B,2] int scratch
int launch_fscore(const float* gt, const float* pr, int B, int n, int m, double th, int* counts, double* out, cudaStream_t st);
// correct[k] = number of samples whose target class is within the top ks[k] scores
int launch_topk(const float* out, const int64_t* target, int B, int C, const int* ks, int nk, int* correct, cudaStream_t st);

}  // namespace lsdm
