// Shared device/host helpers for the lsdm_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include <atomic>

namespace lsdm {

// cudaFuncAttributeMaxDynamicSharedMemorySize is a PER-DEVICE attribute of a kernel: every launcher opts in once per
// (kernel, device), not once per process (a second GPU driven from the same process needs its own opt-in).
struct PerDeviceOnce {
  std::atomic<uint64_t> mask{0};
  bool first() {
    int d = 0;
    cudaGetDevice(&d);
    const uint64_t bit = 1ull << (d & 63);
    return !(mask.fetch_or(bit) & bit);
  }
};
template <typename Kernel>
inline cudaError_t smem_opt_in(PerDeviceOnce& once, Kernel kernel, int bytes) {
  if (!once.first()) return cudaSuccess;
  return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
}
inline int device_sm_count() {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return sms;
}

// Shared-memory accesses by explicit 32-bit shared-window address.  nvcc (12.9, sm_100a) re-derives the address of a shared
// object that is reached through a pointer (a struct of sub-array pointers, a device-function argument) from the generic
// address on every access -- S2R SR_CgaCtaId + LEA + adds, not hoisted out of loops -- which costs more than the access in the
// selection kernels' inner loops.  `smem_addr` converts once; the accessors below are plain ld/st.shared.
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ float4 lds_f4(uint32_t a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ float lds_f32(uint32_t a) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t a) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t lds_u16(uint32_t a) {
  uint32_t v;
  asm volatile("{ .reg .u16 t; ld.shared.u16 t, [%1]; cvt.u32.u16 %0, t; }" : "=r"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts_u32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }

constexpr int NPTS = 1024;  // points per cloud
constexpr int NOBJ = 9;     // object slots per scene
constexpr int CLIP = 512;
constexpr int LAT = 128;
constexpr int CATEMB = 32;
constexpr int TRANS = 12;
constexpr int NHEAD = 8;

enum Act : int { ACT_NONE = 0, ACT_RELU = 1, ACT_GELU = 2, ACT_SIGMOID = 3, ACT_SILU = 4, ACT_QGELU = 5 /* x sigmoid(1.702 x), CLIP */ };

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }
__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

template <int ACT>
__device__ __forceinline__ float apply_act(float x) {
  if (ACT == ACT_RELU) return fmaxf(x, 0.0f);
  if (ACT == ACT_GELU) return gelu_erf(x);
  if (ACT == ACT_SIGMOID) return sigmoidf_(x);
  if (ACT == ACT_SILU) return x * sigmoidf_(x);
  if (ACT == ACT_QGELU) return x * sigmoidf_(1.702f * x);
  return x;
}

__device__ __forceinline__ float apply_act_rt(float x, int act) {
  switch (act) {
    case ACT_RELU: return fmaxf(x, 0.0f);
    case ACT_GELU: return gelu_erf(x);
    case ACT_SIGMOID: return sigmoidf_(x);
    case ACT_SILU: return x * sigmoidf_(x);
    case ACT_QGELU: return x * sigmoidf_(1.702f * x);
    default: return x;
  }
}

// x == hi + lo (to ~2^-22 relative): hi = x rounded to nearest TF32, lo = the residual rounded to TF32 (finite inputs)
__device__ __forceinline__ float tf32_round_fin(float x) { return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u); }
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
  hi = tf32_round_fin(x);
  lo = tf32_round_fin(x - hi);
}

// Tensor-core epilogue forms of the transcendental activations: MUFU ex2 / rcp instead of the ~25-instruction libm expf and
// IEEE division (the epilogue warps, not the MMAs, bounded the sigmoid / GELU layers of the x0 network).  Relative error
// <= ~4e-7 (ex2.approx 2^-22, rcp.approx 1 ulp; erf: Abramowitz-Stegun 7.1.26, |error| <= 1.5e-7) -- far inside the 1e-3
// bound of the path; the fp32 CUDA-core build (gemm_simt.cu) keeps the libm forms.
__device__ __forceinline__ float fast_sigmoid(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float fast_erf(float x) {
  const float ax = fabsf(x);
  const float t = __fdividef(1.0f, fmaf(0.3275911f, ax, 1.0f));
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  const float r = 1.0f - p * t * __expf(-ax * ax);
  return copysignf(r, x);
}
template <int ACT>
__device__ __forceinline__ float apply_act_fast(float x) {
  if (ACT == ACT_GELU) return 0.5f * x * (1.0f + fast_erf(x * 0.70710678118654752440f));
  if (ACT == ACT_SIGMOID) return fast_sigmoid(x);
  if (ACT == ACT_SILU) return x * fast_sigmoid(x);
  if (ACT == ACT_QGELU) return x * fast_sigmoid(1.702f * x);
  return apply_act<ACT>(x);
}

// Epilogue transform of one 32-column accumulator chunk: f = [round_tf32](act(v + row_bias + col_bias)).
// The activation / rounding selectors are resolved ONCE per chunk (switch outside the unrolled element loop).
template <int ACT, bool ROUND>
__device__ __forceinline__ void epi_chunk_t(float (&f)[32], const uint32_t (&v)[32], float rbias, const float* cbias) {
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    float x = __uint_as_float(v[j]) + rbias;
    if (cbias) x += cbias[j];
    x = apply_act_fast<ACT>(x);
    if (ROUND) {
      // == cvt.rna.tf32.f32 for finite x (two integer ops instead of the FSETP + IADD + LOP3 the cvt expands to)
      x = __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);
    }
    f[j] = x;
  }
}
__device__ __forceinline__ void epi_chunk(float (&f)[32], const uint32_t (&v)[32], float rbias, const float* cbias, int act,
                                          int round_out) {
  switch (act * 2 + (round_out ? 1 : 0)) {
    case ACT_NONE * 2: epi_chunk_t<ACT_NONE, false>(f, v, rbias, cbias); break;
    case ACT_NONE * 2 + 1: epi_chunk_t<ACT_NONE, true>(f, v, rbias, cbias); break;
    case ACT_RELU * 2: epi_chunk_t<ACT_RELU, false>(f, v, rbias, cbias); break;
    case ACT_RELU * 2 + 1: epi_chunk_t<ACT_RELU, true>(f, v, rbias, cbias); break;
    case ACT_GELU * 2: epi_chunk_t<ACT_GELU, false>(f, v, rbias, cbias); break;
    case ACT_GELU * 2 + 1: epi_chunk_t<ACT_GELU, true>(f, v, rbias, cbias); break;
    case ACT_SIGMOID * 2: epi_chunk_t<ACT_SIGMOID, false>(f, v, rbias, cbias); break;
    case ACT_SIGMOID * 2 + 1: epi_chunk_t<ACT_SIGMOID, true>(f, v, rbias, cbias); break;
    case ACT_SILU * 2: epi_chunk_t<ACT_SILU, false>(f, v, rbias, cbias); break;
    case ACT_SILU * 2 + 1: epi_chunk_t<ACT_SILU, true>(f, v, rbias, cbias); break;
    case ACT_QGELU * 2: epi_chunk_t<ACT_QGELU, false>(f, v, rbias, cbias); break;
    default: epi_chunk_t<ACT_QGELU, true>(f, v, rbias, cbias); break;
  }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Squared distance in the reference's expanded form (model/pcd_backbone/pointnet2_utils.py:19-38):
// d = (-2 * (a.b)) + |a|^2 + |b|^2, accumulated in that order, no re-association.
__device__ __forceinline__ float sqdist_expanded(float ax, float ay, float az, float a2, float bx, float by, float bz,
                                                 float b2) {
  float dot = __fmaf_rn(az, bz, __fmaf_rn(ay, by, __fmul_rn(ax, bx)));
  float d = __fmul_rn(-2.0f, dot);
  d = __fadd_rn(d, a2);
  d = __fadd_rn(d, b2);
  return d;
}
// |p|^2 as torch.sum(p**2, -1): (x*x + y*y) + z*z without FMA contraction.
__device__ __forceinline__ float sqnorm3(float x, float y, float z) {
  return __fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z));
}

}  // namespace lsdm
