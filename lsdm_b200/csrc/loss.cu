// Training-loss kernels (reference diffusion/gaussian_diffusion.py:1297-1301,1334).
//  - chamfer: pytorch3d.loss.chamfer_distance with default arguments (squared L2 nearest neighbour in both
//    directions, point mean, batch mean, directions summed).  K = 3 coordinates: CUDA-core min-reduce with
//    the opposite cloud staged in shared memory; no N x M matrix is materialised.
//  - cat loss: CrossEntropy applied to ALREADY-softmaxed probabilities (the reference's behaviour).
#include "kernels.cuh"

namespace lsdm {

namespace {

// grid (ceil(n/256), B, 2): z = 0 -> for each x_i min_j |x_i - y_j|^2 ; z = 1 -> roles swapped.
// sums[b * sample_stride + z] += (1/n_z) * sum_i min  (atomic per CTA; sample_stride 0 = batch total, 2 = per sample)
__global__ void __launch_bounds__(256) chamfer_kernel(const float* __restrict__ x, const float* __restrict__ y, int n,
                                                      int m, float* __restrict__ sums, int sample_stride) {
  extern __shared__ float sm[];  // other cloud, xyz interleaved -> SoA
  const int dir = blockIdx.z, b = blockIdx.y;
  const float* a = dir == 0 ? x + (int64_t)b * n * 3 : y + (int64_t)b * m * 3;
  const float* o = dir == 0 ? y + (int64_t)b * m * 3 : x + (int64_t)b * n * 3;
  const int na = dir == 0 ? n : m, no = dir == 0 ? m : n;
  float* sx = sm;
  float* sy = sm + no;
  float* sz = sm + 2 * no;
  for (int i = threadIdx.x; i < no; i += blockDim.x) {
    sx[i] = o[i * 3];
    sy[i] = o[i * 3 + 1];
    sz[i] = o[i * 3 + 2];
  }
  __syncthreads();
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  float best = 0.f;
  if (i < na) {
    float px = a[i * 3], py = a[i * 3 + 1], pz = a[i * 3 + 2];
    best = INFINITY;
    for (int j = 0; j < no; ++j) {
      float dx = px - sx[j], dy = py - sy[j], dz = pz - sz[j];
      float d = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
      best = fminf(best, d);
    }
  }
  __shared__ float s_part[8];
  float v = warp_sum(best);
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    float tot = 0.f;
    for (int k = 0; k < (int)(blockDim.x >> 5); ++k) tot += s_part[k];
    atomicAdd(&sums[b * sample_stride + dir], tot / (float)na);
  }
}

// one warp per sample: -log_softmax(probs)[argmax(target)]
__global__ void cat_loss_kernel(const float* __restrict__ probs, const float* __restrict__ target, int B, int C,
                                float* __restrict__ sum) {
  int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  int lane = threadIdx.x & 31;
  if (b >= B) return;
  float p = lane < C ? probs[(int64_t)b * C + lane] : -INFINITY;
  float tv = lane < C ? target[(int64_t)b * C + lane] : -INFINITY;
  // argmax(target), first maximum
  float tmax = warp_max(tv);
  unsigned cand = __ballot_sync(0xffffffffu, tv == tmax && lane < C);
  int cls = __ffs(cand) - 1;
  float mx = warp_max(p);
  float e = lane < C ? expf(p - mx) : 0.f;
  float lse = logf(warp_sum(e)) + mx;
  float pc = __shfl_sync(0xffffffffu, p, cls);
  if (lane == 0) atomicAdd(sum, lse - pc);
}

}  // namespace

int launch_chamfer(const float* x, const float* y, int B, int n, int m, float* sums, cudaStream_t st, int sample_stride) {
  int big = n > m ? n : m;
  dim3 grid((big + 255) / 256, B, 2);
  chamfer_kernel<<<grid, 256, 3 * big * sizeof(float), st>>>(x, y, n, m, sums, sample_stride);
  return 1;
}

int launch_cat_loss(const float* probs, const float* target, int B, int C, float* sum, cudaStream_t st) {
  cat_loss_kernel<<<(B + 7) / 8, 256, 0, st>>>(probs, target, B, C, sum);
  return 1;
}

}  // namespace lsdm
