// Does ptxas contract mul.rn.f32x2 + add.rn.f32x2 (the __fmul2_rn / __fadd2_rn intrinsics) into a fused FFMA2?
// Replays the farthest-point-sampling distance update (pointnet_select.cu) in its packed form and in the scalar
// __fmul_rn/__fadd_rn form (never contracted) on random inputs and counts differing results.
// Build: nvcc -O3 -DVARIANT=0|1 -gencode arch=compute_100a,code=sm_100a -o packed_fma_probe packed_fma_probe.cu ; `cuobjdump -sass | grep FFMA2`.
// Measured on B200 (CUDA 12.9): VARIANT 0 -> 8 FFMA2 in SASS, 21 % of results differ; VARIANT 1 -> see DESIGN.md.
#ifndef VARIANT
#define VARIANT 0
#endif
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
constexpr int PER2 = 4;
__global__ void probe(const float* __restrict__ pts, const float* __restrict__ cen, int ncen, int* mism) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  float2 qx[PER2], qy[PER2], qz[PER2], qd[PER2];
  float sd[2 * PER2];
#pragma unroll
  for (int j = 0; j < PER2; ++j) {
    const float* p = pts + ((size_t)tid * PER2 + j) * 6;
    qx[j] = make_float2(p[0], p[3]); qy[j] = make_float2(p[1], p[4]); qz[j] = make_float2(p[2], p[5]);
    qd[j] = make_float2(1e10f, 1e10f);
    sd[2 * j] = sd[2 * j + 1] = 1e10f;
  }
  int bad = 0;
  for (int it = 0; it < ncen; ++it) {
    const float cx = cen[it * 3], cy = cen[it * 3 + 1], cz = cen[it * 3 + 2];
    const float2 ncx = make_float2(-cx, -cx), ncy = make_float2(-cy, -cy), ncz = make_float2(-cz, -cz);
#pragma unroll
    for (int j = 0; j < PER2; ++j) {
      float2 dx = __fadd2_rn(qx[j], ncx), dy = __fadd2_rn(qy[j], ncy), dz = __fadd2_rn(qz[j], ncz);
#if VARIANT == 0  // all packed: ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 when the product has one use
      float2 d = __fadd2_rn(__fadd2_rn(__fmul2_rn(dx, dx), __fmul2_rn(dy, dy)), __fmul2_rn(dz, dz));
#else              // packed products, scalar (explicitly rounded, never contracted) sums
      const float2 sxx = __fmul2_rn(dx, dx), syy = __fmul2_rn(dy, dy), szz = __fmul2_rn(dz, dz);
      float2 d = make_float2(__fadd_rn(__fadd_rn(sxx.x, syy.x), szz.x), __fadd_rn(__fadd_rn(sxx.y, syy.y), szz.y));
#endif
      qd[j] = make_float2(fminf(d.x, qd[j].x), fminf(d.y, qd[j].y));
      float ex = __fsub_rn(qx[j].x, cx), ey = __fsub_rn(qy[j].x, cy), ez = __fsub_rn(qz[j].x, cz);
      float e0 = __fadd_rn(__fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey)), __fmul_rn(ez, ez));
      ex = __fsub_rn(qx[j].y, cx); ey = __fsub_rn(qy[j].y, cy); ez = __fsub_rn(qz[j].y, cz);
      float e1 = __fadd_rn(__fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey)), __fmul_rn(ez, ez));
      bad += (d.x != e0) + (d.y != e1);
      sd[2 * j] = fminf(e0, sd[2 * j]); sd[2 * j + 1] = fminf(e1, sd[2 * j + 1]);
    }
  }
  if (bad) atomicAdd(mism, bad);
}
int main() {
  const int nthr = 1 << 14, ncen = 256;
  float *pts, *cen; int* m;
  cudaMallocManaged(&pts, (size_t)nthr * PER2 * 6 * 4); cudaMallocManaged(&cen, ncen * 3 * 4); cudaMallocManaged(&m, 8);
  srand(1);
  for (size_t i = 0; i < (size_t)nthr * PER2 * 6; ++i) pts[i] = rand() / (float)RAND_MAX - 0.5f;
  for (int i = 0; i < ncen * 3; ++i) cen[i] = rand() / (float)RAND_MAX - 0.5f;
  m[0] = m[1] = 0;
  probe<<<nthr / 128, 128>>>(pts, cen, ncen, m);
  cudaDeviceSynchronize();
  printf("variant %d vs scalar squared distances: %d of %lld differ\n", VARIANT, m[0], (long long)nthr * PER2 * 2 * ncen);
  return 0;
}
