// Interface of the training backward pass (train_bw.cu); the C ABI wrapper lives in api.cu (lsdm_training_backward).
#pragma once
#include <functional>
#include <string>

#include "../../include/lsdm_b200.h"
#include "common.cuh"

namespace lsdm {

struct TrainCtx {
  std::function<const float*(const std::string&)> W;  // state-dict entry (raw, un-folded) on the device
  std::function<float*(const std::string&)> G;        // its slot in the flat gradient buffer (accumulated into)
  int B = 0, Bg = 0, b_off = 0, n_cats = 0;
  void* tape = nullptr;
  size_t tape_bytes = 0;
  lsdm_allreduce_fn allreduce = nullptr;  // SyncBN: sums (double) over the data-parallel shards
  void* allreduce_ctx = nullptr;
  const float *sched_sa = nullptr, *sched_s1a = nullptr;  // sqrt(abar_t), sqrt(1 - abar_t)
};

struct TrainIO {
  const float *text = nullptr, *objs = nullptr, *cats = nullptr, *mask_global = nullptr, *x_start = nullptr, *noise = nullptr,
              *target_cat = nullptr, *drop_mask = nullptr;
  const int64_t* t = nullptr;
  // selection results of the PointNet++ levels (from the handle's selection chain): centroid coordinates xyz[1..4],
  // ball-query groups grp[0..3] ([C,S,32] source indices), 3-NN indices / weights of fp4, fp3, fp2, fp1
  const float* xyz[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  const int* grp[4] = {nullptr, nullptr, nullptr, nullptr};
  const int* nn_idx[4] = {nullptr, nullptr, nullptr, nullptr};
  const float* nn_w[4] = {nullptr, nullptr, nullptr, nullptr};
  float g_mse = 1.f, g_cat = 1.f, lambda_cat = 0.1f;  // upstream gradients of the two loss terms
  float* losses_out = nullptr;                          // device [2]: chamfer, mean cross-entropy (before lambda_cat)
  float* x0_out = nullptr;                              // device [B,1024,3] model output (nullable)
};

size_t train_tape_bytes(int B, int n_cats);
// size_only != nullptr: writes the tape size and returns without launching anything.
int train_forward_backward(const TrainCtx& ctx, const TrainIO& io, cudaStream_t st, size_t* size_only = nullptr);

// multi-tensor AdamW on flat buffers (torch.optim.AdamW semantics: decoupled weight decay, bias correction)
int launch_adamw(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2, float eps, float weight_decay,
                 int64_t step, float grad_scale, cudaStream_t st);

}  // namespace lsdm
