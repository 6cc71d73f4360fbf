// Generic "linear layer" GEMM interface used by every dense layer of the path:
//   C[M,N] = act(A[M,K] . W[N,K]^T + bias)       (A and W row-major, K contiguous: PyTorch's layout)
// optionally batched (blockIdx.z) and optionally reduced by max over groups of 32 consecutive rows
// (the PointNet++ set-abstraction pooling, reference pointnet2_utils.py:197).
#pragma once
#include "common.cuh"

namespace lsdm {

struct GemmArgs {
  const float* A;
  int64_t lda, strideA;
  const float* W;
  int64_t ldw, strideW;
  float* C;
  int64_t ldc, strideC;
  const float* bias;
  int bias_mode;  // 0 none, 1 per output column, 2 per output row
  int M, N, K, batch;
  int act;        // lsdm::Act
  int group_max;  // 1: C has M/32 rows, row g = max over rows [32g, 32g+32) of the activated tile
  int precision;  // 0: fp32 CUDA cores (bit-faithful); 1: TF32 tensor cores (tcgen05, operands rounded-to-nearest);
                  // 2: 3xTF32 (hi/lo split of both operands, ~fp32 accuracy on the tensor cores)
  int a_rounded;  // A already holds TF32-representable values (its producer rounded them)  } both set: operands can be
  int w_rounded;  // W points at the TF32-rounded weight copy made at finalize                } staged by cp.async, no registers
  int round_out;  // epilogue rounds C to TF32 (round-to-nearest) because C feeds another tensor-core GEMM
  // 3xTF32 with PRE-SPLIT operands (precision 2): x == hi + lo with hi = rna_tf32(x), lo = rna_tf32(x - hi).  When both
  // A_lo and W_lo are given, A / W point at the hi planes and all four planes are staged by cp.async (same strides);
  // when C_lo is given the epilogue writes the result as a hi plane (C) and a lo plane (C_lo) for the next GEMM.
  const float* A_lo;
  const float* W_lo;
  float* C_lo;
};

// Launches the GEMM on `stream`; returns the number of kernels launched (1) or a negative value on bad arguments.
int launch_gemm(const GemmArgs& g, cudaStream_t stream);
int launch_gemm_simt(const GemmArgs& g, cudaStream_t stream);
int launch_gemm_tc(const GemmArgs& g, cudaStream_t stream);   // one CTA per tile (simple pipeline)
int launch_gemm_ws(const GemmArgs& g, cudaStream_t stream);   // persistent, warp-specialised (default tensor path)
extern int g_gemm_async;                                       // 1: pre-rounded operands take the cp.async-fed persistent kernel
extern int g_gemm_tma;                                         // 1: un-batched pre-rounded operands are fed by TMA instead of cp.async
extern int g_gemm_ws;                                          // 1: launch_gemm uses the warp-specialised kernel
bool gemm_tc_eligible(const GemmArgs& g);

}  // namespace lsdm
