// Evaluation metrics of the sampling path (reference run/test_sdm.py:186-207, util/evaluation.py:5-52): what the reference
// computes on the HOST right after p_sample_loop -- here on the device, batched over samples.
//  - EMD: util/evaluation.py:5-11 = scipy cdist + linear_sum_assignment (exact min-cost perfect matching), mean matched distance.
//    Device algorithm: Bertsekas forward auction with epsilon scaling, one CTA per sample, Jacobi bidding rounds, exact integer
//    arithmetic on distances scaled by a power of two (deterministic; the matching is within n*eps_final of the optimum, i.e.
//    <= 2^-26 of the cloud diameter per point).  The returned value is recomputed in double from the matching.
//  - F-score: util/evaluation.py:28-52 = open3d compute_point_cloud_distance both ways (nearest-neighbour Euclidean distance,
//    double), fraction below the threshold, harmonic mean.
//  - top-k accuracy: util/evaluation.py:13-26.
//  - per-sample Chamfer: pytorch3d.loss.chamfer_distance as called at run/test_sdm.py:187 (shares loss.cu's kernel).
#include <climits>

#include "kernels.cuh"

namespace lsdm {

namespace {

constexpr int EMD_MAXN = 1024;
constexpr int EMD_THREADS = 1024;
constexpr int EMD_WARPS = EMD_THREADS / 32;
constexpr int EMD_MAX_ROUNDS = 1 << 20;  // safety cap: a sample that hits it reports NaN instead of hanging the device

struct Top2 {
  long long v1, v2;
  int j1;
};

__device__ __forceinline__ long long shfl_xor_ll(long long v, int o) {
  int lo = __shfl_xor_sync(0xffffffffu, (int)(v & 0xffffffffll), o);
  int hi = __shfl_xor_sync(0xffffffffu, (int)(v >> 32), o);
  return ((long long)hi << 32) | (unsigned int)lo;
}

// merge two (best, index of best, second best) triples; ties on the best value go to the smaller object index
__device__ __forceinline__ Top2 merge_top2(const Top2& a, const Top2& b) {
  Top2 r;
  const bool a_first = a.v1 > b.v1 || (a.v1 == b.v1 && a.j1 < b.j1);
  if (a_first) {
    r.v1 = a.v1;
    r.j1 = a.j1;
    r.v2 = a.v2 > b.v1 ? a.v2 : b.v1;
  } else {
    r.v1 = b.v1;
    r.j1 = b.j1;
    r.v2 = b.v2 > a.v1 ? b.v2 : a.v1;
  }
  return r;
}

// One CTA per sample.  Persons = points of x, objects = points of y, both n <= 1024.
// Shared memory (dynamic): xs|ys SoA (6n floats), price[n] int64, bidkey[n] uint64, part_v1[n], part_v2[n] int64,
// part_j[n], owner[n], assigned[n], list[n] int32.
__global__ void __launch_bounds__(EMD_THREADS, 1) emd_auction_kernel(const float* __restrict__ X, const float* __restrict__ Y, int n,
                                                                     double* __restrict__ out_emd, int* __restrict__ out_assign,
                                                                     int* __restrict__ out_rounds) {
  extern __shared__ __align__(16) unsigned char smem[];
  long long* price = (long long*)smem;
  unsigned long long* bidkey = (unsigned long long*)(price + n);
  long long* part_v1 = (long long*)(bidkey + n);
  long long* part_v2 = part_v1 + EMD_MAXN;
  float* px = (float*)(part_v2 + EMD_MAXN);
  float* py = px + n;
  float* pz = py + n;
  float* ox = pz + n;
  float* oy = ox + n;
  float* oz = oy + n;
  int* part_j = (int*)(oz + n);
  int* owner = part_j + EMD_MAXN;
  int* assigned = owner + n;
  int* list = assigned + n;
  __shared__ int s_cnt;
  __shared__ float s_red[EMD_WARPS];
  __shared__ double s_dred[EMD_WARPS];

  const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const float* x = X + (int64_t)b * n * 3;
  const float* y = Y + (int64_t)b * n * 3;

  // ---- load, bounding-box extent -> power-of-two scale so that every scaled distance is < 2^30 ----
  float ext = 0.f;
  if (tid < n) {
    px[tid] = x[tid * 3];
    py[tid] = x[tid * 3 + 1];
    pz[tid] = x[tid * 3 + 2];
    ox[tid] = y[tid * 3];
    oy[tid] = y[tid * 3 + 1];
    oz[tid] = y[tid * 3 + 2];
    ext = fmaxf(fmaxf(fmaxf(fabsf(px[tid]), fabsf(py[tid])), fabsf(pz[tid])), fmaxf(fmaxf(fabsf(ox[tid]), fabsf(oy[tid])), fabsf(oz[tid])));
    price[tid] = 0;
    bidkey[tid] = 0ull;
  }
  ext = warp_max(ext);
  if (lane == 0) s_red[warp] = ext;
  __syncthreads();
  ext = 0.f;
  for (int k = 0; k < EMD_WARPS; ++k) ext = fmaxf(ext, s_red[k]);
  // every distance <= 2*sqrt(3)*ext < 4*ext: scale = 2^(28 - ceil(log2(ext)))  ->  scaled distance < 2^30
  int e2 = 0;
  if (ext > 0.f) (void)frexpf(ext, &e2);  // ext = m * 2^e2, m in [0.5, 1)  ->  ext <= 2^e2
  const float scale = ldexpf(1.0f, 28 - e2);
  const long long eps_final = 16;
  long long eps = 1ll << 27;  // first phase: an eighth of the largest possible cost

  int rounds = 0;
  bool failed = false;
  for (;;) {  // ---- epsilon-scaling phases: prices are kept, the assignment restarts ----
    if (tid < n) {
      owner[tid] = -1;
      assigned[tid] = -1;
    }
    __syncthreads();
    for (;;) {  // ---- Jacobi bidding rounds ----
      if (tid == 0) s_cnt = 0;
      __syncthreads();
      if (tid < n && assigned[tid] < 0) list[atomicAdd(&s_cnt, 1)] = tid;  // order is irrelevant: a round's result is order-free
      __syncthreads();
      const int u = s_cnt;
      if (u == 0) break;
      if (++rounds > EMD_MAX_ROUNDS) {
        failed = true;
        break;
      }
      // G warps share one bidder when few are left (u * G <= 32)
      int G = 1;
      while (G < EMD_WARPS && u * (G * 2) <= EMD_WARPS) G *= 2;
      const int items = u * G;
      for (int it = warp; it < items; it += EMD_WARPS) {
        const int i = list[it / G], slice = it % G;
        const float ax = px[i], ay = py[i], az = pz[i];
        Top2 t{LLONG_MIN, LLONG_MIN, 0x7fffffff};
        for (int j = slice * 32 + lane; j < n; j += 32 * G) {
          const float dx = ax - ox[j], dy = ay - oy[j], dz = az - oz[j];
          const float d = __fsqrt_rn(__fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx))));
          const long long v = -__float2ll_rn(d * scale) - price[j];
          if (v > t.v1) {
            t.v2 = t.v1;
            t.v1 = v;
            t.j1 = j;
          } else if (v > t.v2) {
            t.v2 = v;
          }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          Top2 q;
          q.v1 = shfl_xor_ll(t.v1, o);
          q.v2 = shfl_xor_ll(t.v2, o);
          q.j1 = __shfl_xor_sync(0xffffffffu, t.j1, o);
          t = merge_top2(t, q);
        }
        if (lane == 0) {
          part_v1[it] = t.v1;
          part_v2[it] = t.v2;
          part_j[it] = t.j1;
        }
      }
      __syncthreads();
      if (tid < u) {
        Top2 t{part_v1[tid * G], part_v2[tid * G], part_j[tid * G]};
        for (int s = 1; s < G; ++s) t = merge_top2(t, Top2{part_v1[tid * G + s], part_v2[tid * G + s], part_j[tid * G + s]});
        // bid = price[j1] + (best - second) + eps ; n == 1 has no second-best object
        const long long gap = t.v2 == LLONG_MIN ? 0 : t.v1 - t.v2;
        const long long bid = price[t.j1] + gap + eps;
        const int i = list[tid];
        atomicMax(&bidkey[t.j1], ((unsigned long long)bid << 10) | (unsigned long long)(1023 - i));  // ties -> smaller person index
      }
      __syncthreads();
      if (tid < n) {
        const unsigned long long k = bidkey[tid];
        if (k != 0ull) {
          const int winner = 1023 - (int)(k & 1023ull);
          const int prev = owner[tid];
          if (prev >= 0) assigned[prev] = -1;  // prev was assigned, hence not bidding: no other writer touches it
          owner[tid] = winner;
          assigned[winner] = tid;
          price[tid] = (long long)(k >> 10);
          bidkey[tid] = 0ull;
        }
      }
      // the list rebuild at the top of the loop is behind a barrier
    }
    if (failed || eps <= eps_final) break;
    eps = eps / 8 > eps_final ? eps / 8 : eps_final;
    __syncthreads();
  }

  // ---- matched distance, double, from the final matching ----
  double acc = 0.0;
  if (tid < n && !failed) {
    const int j = assigned[tid];
    const double dx = (double)px[tid] - (double)ox[j], dy = (double)py[tid] - (double)oy[j], dz = (double)pz[tid] - (double)oz[j];
    acc = sqrt(dx * dx + dy * dy + dz * dz);
    if (out_assign) out_assign[(int64_t)b * n + tid] = j;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) s_dred[warp] = acc;
  __syncthreads();
  if (tid == 0) {
    double tot = 0.0;
    for (int k = 0; k < EMD_WARPS; ++k) tot += s_dred[k];
    out_emd[b] = failed ? nan("") : tot / (double)n;
    if (out_rounds) out_rounds[b] = rounds;
  }
}

// grid (ceil(max(n,m)/256), B, 2): dir 0 -> for every gt point the distance to its nearest predicted point (d1 of
// util/evaluation.py:36), dir 1 -> the other way (d2, :37); counts[b][dir] += #(distance < th).  Double arithmetic as open3d.
__global__ void __launch_bounds__(256) fscore_count_kernel(const float* __restrict__ gt, const float* __restrict__ pr, int n, int m,
                                                           double th, int* __restrict__ counts) {
  extern __shared__ float sm[];
  const int dir = blockIdx.z, b = blockIdx.y;
  const float* a = dir == 0 ? gt + (int64_t)b * n * 3 : pr + (int64_t)b * m * 3;
  const float* o = dir == 0 ? pr + (int64_t)b * m * 3 : gt + (int64_t)b * n * 3;
  const int na = dir == 0 ? n : m, no = dir == 0 ? m : n;
  for (int i = threadIdx.x; i < no * 3; i += blockDim.x) sm[i] = o[i];
  __syncthreads();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  int hit = 0;
  if (i < na) {
    const double ax = a[i * 3], ay = a[i * 3 + 1], az = a[i * 3 + 2];
    double best = INFINITY;
    for (int j = 0; j < no; ++j) {
      const double dx = ax - (double)sm[j * 3], dy = ay - (double)sm[j * 3 + 1], dz = az - (double)sm[j * 3 + 2];
      best = fmin(best, dx * dx + dy * dy + dz * dz);
    }
    hit = sqrt(best) < th ? 1 : 0;
  }
  const unsigned m32 = __ballot_sync(0xffffffffu, hit);
  if ((threadIdx.x & 31) == 0 && m32) atomicAdd(&counts[b * 2 + dir], __popc(m32));
}

// out[b] = {fscore, precision, recall} (util/evaluation.py:39-52): precision from d1 (gt -> pr), recall from d2 (pr -> gt)
__global__ void fscore_final_kernel(const int* __restrict__ counts, int B, int n, int m, double* __restrict__ out) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const double precision = (double)counts[b * 2] / (double)n;
  const double recall = (double)counts[b * 2 + 1] / (double)m;
  out[b * 3] = recall + precision > 0.0 ? 2.0 * recall * precision / (recall + precision) : 0.0;
  out[b * 3 + 1] = precision;
  out[b * 3 + 2] = recall;
}

// one warp per sample: rank of the target class among the C scores (number of classes placed before it by a stable
// descending sort); correct[k] += rank < ks[k]
__global__ void topk_kernel(const float* __restrict__ out, const int64_t* __restrict__ target, int B, int C, const int* __restrict__ ks,
                            int nk, int* __restrict__ correct) {
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (b >= B) return;
  const int tc = (int)target[b];
  const float tv = out[(int64_t)b * C + tc];
  int before = 0;
  for (int c = lane; c < C; c += 32) {
    const float v = out[(int64_t)b * C + c];
    before += (v > tv || (v == tv && c < tc)) ? 1 : 0;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) before += __shfl_xor_sync(0xffffffffu, before, o);
  if (lane == 0)
    for (int k = 0; k < nk; ++k)
      if (before < ks[k]) atomicAdd(&correct[k], 1);
}

}  // namespace

size_t emd_smem_bytes(int n) {
  return (size_t)n * (8 + 8) + (size_t)EMD_MAXN * (8 + 8 + 4) + (size_t)n * (6 * 4 + 3 * 4) + 64;
}

int launch_emd(const float* x, const float* y, int B, int n, double* out_emd, int* out_assign, int* out_rounds, cudaStream_t st) {
  if (n < 1 || n > EMD_MAXN) return -1;
  const size_t smem = emd_smem_bytes(n);
  static PerDeviceOnce attr_done;
  if (smem_opt_in(attr_done, emd_auction_kernel, (int)emd_smem_bytes(EMD_MAXN)) != cudaSuccess) return -1;
  emd_auction_kernel<<<B, EMD_THREADS, smem, st>>>(x, y, n, out_emd, out_assign, out_rounds);
  return 1;
}

int launch_fscore(const float* gt, const float* pr, int B, int n, int m, double th, int* counts, double* out, cudaStream_t st) {
  const int big = n > m ? n : m;
  if (big > 4096) return -1;
  cudaMemsetAsync(counts, 0, sizeof(int) * 2 * B, st);
  dim3 grid((big + 255) / 256, B, 2);
  fscore_count_kernel<<<grid, 256, 3 * big * sizeof(float), st>>>(gt, pr, n, m, th, counts);
  fscore_final_kernel<<<(B + 127) / 128, 128, 0, st>>>(counts, B, n, m, out);
  return 2;
}

int launch_topk(const float* out, const int64_t* target, int B, int C, const int* ks, int nk, int* correct, cudaStream_t st) {
  cudaMemsetAsync(correct, 0, sizeof(int) * nk, st);
  topk_kernel<<<(B + 7) / 8, 256, 0, st>>>(out, target, B, C, ks, nk, correct);
  return 1;
}

}  // namespace lsdm
