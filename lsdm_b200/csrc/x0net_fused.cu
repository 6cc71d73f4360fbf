// The whole x0 network of one denoising step in ONE persistent kernel (reference model/sdm.py:204,211-217,
// model/diffusion_utils.py:66-78,107-122, diffusion/gaussian_diffusion.py:266-269,356-357,545-560):
//
//   z = x + pcd_out (x updated in place)  |  z = pcd_out (guiding-points pass)
//   h1   = sigmoid(E0 z + b)              3 -> 64     CUDA cores
//   pose = sigmoid(E2 h1 + b)            64 -> 128    tcgen05, 3xTF32
//   c1   = sigmoid(C0 [pose || emb] + b) 256 -> 192   tcgen05, 3xTF32 (emb: the step's embedding tile, by TMA)
//   c2   = sigmoid(C2 c1 + b)            192 -> 128   tcgen05, 3xTF32
//   f1   = gelu(F0 c2 + b)               128 -> 64    tcgen05, 3xTF32
//   x0   = gelu(F2 f1 + b)                64 -> 3     CUDA cores
//   sample = c1[t] x0 + c2[t] z + [t != 0] exp(logvar[t] / 2) noise
//
// per 128-row tile; no activation of the chain ever leaves the SM.  It replaces four GEMM launches, two pose-embedding
// launches, two final launches and a strided copy that moved every activation through HBM as two fp32 planes.
//
// Activations live in TENSOR MEMORY as the A operands of the next layer, split x == hi + lo (hi = rna_tf32(x),
// lo = rna_tf32(x - hi)); every layer issues lo.W_hi + hi.W_lo + hi.W_hi with fp32 accumulation (3xTF32, fp32-grade).
// TMEM column plan (512 columns = the whole tensor memory of the SM; buffers are recycled as the chain advances):
//   h1_hi [384,448) h1_lo [448,512) | D1 -> pose_lo [0,128)   pose_hi [128,256) | D2 -> c1_lo [320,512)  c1_hi [0,192)
//   D3 -> c2_lo [192,320)  c2_hi [0,128) | D4 [320,384)
// (an epilogue thread overwrites the accumulator columns of ITS OWN row with the lo plane after reading them).
//
// Weights (hi and lo planes, 704 KB in total: they do not fit in shared memory) are streamed from L2 by TMA, one
// [N x 32] k-block per 24 KB ring slot, eight slots; the embedding tile's k-blocks travel through the same ring.
// Ten warps: 0-7 layer 0 + epilogues (two per TMEM lane quarter, half of the columns each), 8 TMA producer,
// 9 MMA issuer.  The epilogue warps compute layer 0 of the NEXT tile while the last MMA layer of the current one runs.
#include <cuda.h>

#include "kernels.cuh"
#include "tc_ptx.cuh"

namespace lsdm {

bool tma_map_2d(CUtensorMap* m, const float* ptr, int64_t rows, int K, int64_t ld, int box_rows);  // gemm_ws.cu

namespace {

using namespace tc;

constexpr int XT = 320;           // threads
constexpr int NSLOT = 8;
constexpr int SLOT = 24 * 1024;   // one [192 x 32] fp32 k-block
constexpr int SLOTS_PER_TILE = 4 + 8 + 16 + 12 + 8;

constexpr uint32_t COL_H1_HI = 384, COL_H1_LO = 448, COL_D1 = 0, COL_POSE_HI = 128, COL_D2 = 320, COL_C1_HI = 0, COL_D3 = 192,
                   COL_C2_HI = 0, COL_D4 = 320;

struct X0Maps {
  CUtensorMap w1h, w1l, w2h, w2l, w3h, w3l, w4h, w4l, eh, el;
};

struct X0Consts {  // device pointers into the weight arena
  const float *w0, *b0, *b1, *b2, *b3, *b4, *w5, *b5;
};

struct X0Args {
  float* x;                // [rows,3] in-out
  const float* pcd_out;    // [rows,3]
  const float* noise;      // [rows,3] or null (forward only)
  float* sample_out;       // may alias x; null: x keeps z
  float* x0_out;           // nullable
  float* guiding_out;      // nullable (n_pass == 2 writes it)
  const int64_t* t;        // [B]
  const float *c1, *c2, *logvar;  // schedule tables (null when sample_out is null)
  int rows, n_pass, clip;
};

__device__ __forceinline__ void named_bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

__global__ void __launch_bounds__(XT, 1) x0net_fused_kernel(X0Args a, X0Consts kc, const __grid_constant__ X0Maps maps) {
  extern __shared__ uint8_t smem_raw[];
  // one barrier per hand-over point, each completing exactly one phase per tile (a shared barrier could be run over by a
  // thread that arrives for the next hand-over before a slow neighbour has arrived for / observed the current one):
  //   s_barH: h1 ready (256 epilogue threads -> MMA);  s_barA[i]: activation of layer i+1 ready;  s_barD[i]: accumulator of layer i+1 done
  __shared__ uint64_t s_full[NSLOT], s_empty[NSLOT], s_barH, s_barA[3], s_barD[4];
  __shared__ uint32_t s_tmem;
  __shared__ float s_w0[64 * 3], s_b0[64], s_b1[128], s_b2[192], s_b3[128], s_b4[64], s_w5[3 * 64], s_b5[4];
  __shared__ float s_part[128 * 3];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int n_rowtiles = a.rows / 128;
  const int n_tiles = n_rowtiles * a.n_pass;

  if (warp == 0) tmem_alloc(smem_u32(&s_tmem), 512);
  if (tid == 0) {
    for (int i = 0; i < NSLOT; ++i) {
      mbar_init(smem_u32(&s_full[i]), 1);
      mbar_init(smem_u32(&s_empty[i]), 1);
    }
    mbar_init(smem_u32(&s_barH), 256);
    for (int i = 0; i < 3; ++i) mbar_init(smem_u32(&s_barA[i]), 256);
    for (int i = 0; i < 4; ++i) mbar_init(smem_u32(&s_barD[i]), 1);
    fence_mbar_init();
  }
  for (int i = tid; i < 192; i += XT) {
    s_w0[i] = kc.w0[i];
    s_w5[i] = kc.w5[i];
    s_b2[i] = kc.b2[i];
  }
  if (tid < 128) {
    s_b1[tid] = kc.b1[tid];
    s_b3[tid] = kc.b3[tid];
  }
  if (tid < 64) {
    s_b0[tid] = kc.b0[tid];
    s_b4[tid] = kc.b4[tid];
  }
  if (tid < 3) s_b5[tid] = kc.b5[tid];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s_tmem;

  if (warp == 8) {
    // ======================================= TMA producer =======================================
    if (lane == 0) {
      const CUtensorMap* all[10] = {&maps.w1h, &maps.w1l, &maps.w2h, &maps.w2l, &maps.w3h, &maps.w3l, &maps.w4h, &maps.w4l, &maps.eh, &maps.el};
      for (int i = 0; i < 10; ++i) tma_prefetch_desc(all[i]);
      uint32_t it = 0;
      auto push = [&](const CUtensorMap* m, int k0, int r0, uint32_t bytes) {
        const uint32_t s = it % NSLOT;
        if (it >= (uint32_t)NSLOT) mbar_wait(smem_u32(&s_empty[s]), ((it / NSLOT) & 1u) ^ 1u);
        const uint32_t full = smem_u32(&s_full[s]);
        mbar_arrive_expect_tx(full, bytes);
        tma_load_2d(base + s * SLOT, m, k0, r0, full);
        ++it;
      };
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int r0 = (tile / a.n_pass) * 128;  // embedding rows of this tile (both passes of a row tile share them)
        for (int kb = 0; kb < 2; ++kb) {
          push(&maps.w1h, kb * 32, 0, 128 * 128);
          push(&maps.w1l, kb * 32, 0, 128 * 128);
        }
        for (int kb = 0; kb < 4; ++kb) {
          push(&maps.w2h, kb * 32, 0, 192 * 128);
          push(&maps.w2l, kb * 32, 0, 192 * 128);
        }
        for (int kb = 0; kb < 4; ++kb) {
          push(&maps.eh, kb * 32, r0, 128 * 128);
          push(&maps.el, kb * 32, r0, 128 * 128);
          push(&maps.w2h, 128 + kb * 32, 0, 192 * 128);
          push(&maps.w2l, 128 + kb * 32, 0, 192 * 128);
        }
        for (int kb = 0; kb < 6; ++kb) {
          push(&maps.w3h, kb * 32, 0, 128 * 128);
          push(&maps.w3l, kb * 32, 0, 128 * 128);
        }
        for (int kb = 0; kb < 4; ++kb) {
          push(&maps.w4h, kb * 32, 0, 64 * 128);
          push(&maps.w4l, kb * 32, 0, 64 * 128);
        }
      }
    }
  } else if (warp == 9) {
    // ======================================= MMA issuer =======================================
    if (lane == 0) {
      uint32_t it = 0, ph = 0;
      auto slot_wait = [&](uint32_t i) -> uint32_t {
        const uint32_t s = i % NSLOT;
        mbar_wait(smem_u32(&s_full[s]), (i / NSLOT) & 1u);
        return base + s * SLOT;
      };
      auto slot_free = [&](uint32_t i) { umma_commit(smem_u32(&s_empty[i % NSLOT])); };
      // one k-block (32 columns of K) of a layer whose A operand (hi, lo planes) sits in tensor memory
      auto kblock_ts = [&](uint32_t d, uint32_t a_hi, uint32_t a_lo, uint32_t idesc, bool first) {
        const uint32_t sh = slot_wait(it), sl = slot_wait(it + 1);
        tc_fence_after();
        const uint64_t bh = umma_desc_sw128(sh), bl = umma_desc_sw128(sl);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          const uint64_t o = (uint64_t)(kk * 2);
          umma_tf32_ts(d, a_lo + kk * 8, bh + o, idesc, (first && kk == 0) ? 0u : 1u);
          umma_tf32_ts(d, a_hi + kk * 8, bl + o, idesc, 1u);
          umma_tf32_ts(d, a_hi + kk * 8, bh + o, idesc, 1u);
        }
        slot_free(it);
        slot_free(it + 1);
        it += 2;
      };
      constexpr uint32_t id128 = umma_idesc_tf32(128, 128), id192 = umma_idesc_tf32(128, 192), id64 = umma_idesc_tf32(128, 64);
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ph ^= 1u) {
        // ---- layer 1: D1 = h1 . E2^T ----
        mbar_wait(smem_u32(&s_barH), ph);
        tc_fence_after();
        for (int kb = 0; kb < 2; ++kb) kblock_ts(tmem + COL_D1, tmem + COL_H1_HI + kb * 32, tmem + COL_H1_LO + kb * 32, id128, kb == 0);
        umma_commit(smem_u32(&s_barD[0]));
        // ---- layer 2: D2 = [pose || emb] . C0^T ----
        mbar_wait(smem_u32(&s_barA[0]), ph);
        tc_fence_after();
        for (int kb = 0; kb < 4; ++kb) kblock_ts(tmem + COL_D2, tmem + COL_POSE_HI + kb * 32, tmem + COL_D1 + kb * 32, id192, kb == 0);
        for (int kb = 0; kb < 4; ++kb) {  // embedding half of K: A operand from shared memory
          const uint32_t eh = slot_wait(it), el = slot_wait(it + 1), sh = slot_wait(it + 2), sl = slot_wait(it + 3);
          tc_fence_after();
          const uint64_t ah = umma_desc_sw128(eh), al = umma_desc_sw128(el), bh = umma_desc_sw128(sh), bl = umma_desc_sw128(sl);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            const uint64_t o = (uint64_t)(kk * 2);
            umma_tf32_ss(tmem + COL_D2, al + o, bh + o, id192, 1u);
            umma_tf32_ss(tmem + COL_D2, ah + o, bl + o, id192, 1u);
            umma_tf32_ss(tmem + COL_D2, ah + o, bh + o, id192, 1u);
          }
          for (int j = 0; j < 4; ++j) slot_free(it + j);
          it += 4;
        }
        umma_commit(smem_u32(&s_barD[1]));
        // ---- layer 3: D3 = c1 . C2^T ----
        mbar_wait(smem_u32(&s_barA[1]), ph);
        tc_fence_after();
        for (int kb = 0; kb < 6; ++kb) kblock_ts(tmem + COL_D3, tmem + COL_C1_HI + kb * 32, tmem + COL_D2 + kb * 32, id128, kb == 0);
        umma_commit(smem_u32(&s_barD[2]));
        // ---- layer 4: D4 = c2 . F0^T ----
        mbar_wait(smem_u32(&s_barA[2]), ph);
        tc_fence_after();
        for (int kb = 0; kb < 4; ++kb) kblock_ts(tmem + COL_D4, tmem + COL_C2_HI + kb * 32, tmem + COL_D3 + kb * 32, id64, kb == 0);
        umma_commit(smem_u32(&s_barD[3]));
      }
    }
  } else {
    // ======================================= layer 0 + epilogues =======================================
    const int rit = tid & 127, half = tid >> 7, wq = warp & 3;
    const uint32_t tl = tmem + ((uint32_t)(wq * 32) << 16);
    uint32_t ph = 0;

    // layer 0 of `tile` for this thread's row: z (kept by the caller for the posterior), h1 hi/lo -> tensor memory
    auto layer0 = [&](int tile, float (&z)[3]) {
      const int pass = tile % a.n_pass;
      const int64_t row = (int64_t)(tile / a.n_pass) * 128 + rit;
      const float* p = a.pcd_out + row * 3;
      z[0] = p[0]; z[1] = p[1]; z[2] = p[2];
      if (pass == 0) {
        float* xr = a.x + row * 3;
        z[0] += xr[0]; z[1] += xr[1]; z[2] += xr[2];
        if (half == 0 && a.sample_out != a.x) {  // the reference's in-place `x += pcd_out` stays visible to the caller
          xr[0] = z[0]; xr[1] = z[1]; xr[2] = z[2];
        }
      }
      uint32_t vh[32], vl[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const int ch = half * 32 + j;
        const float v = sigmoidf_(fmaf(s_w0[ch * 3 + 2], z[2], fmaf(s_w0[ch * 3 + 1], z[1], fmaf(s_w0[ch * 3], z[0], s_b0[ch]))));
        const float hi = tf32_round_fin(v);
        vh[j] = __float_as_uint(hi);
        vl[j] = rna_tf32_mma(v - hi);
      }
      tmem_st32(tl + COL_H1_HI + half * 32, vh);
      tmem_st32(tl + COL_H1_LO + half * 32, vl);
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(smem_u32(&s_barH));
    };
    // accumulator chunk -> bias + sigmoid -> hi plane to `col_hi`, lo plane in place
    auto epi_sigmoid = [&](uint32_t col_d, uint32_t col_hi, const float* bias, int chunk) {
      uint32_t v[32], vh[32];
      tmem_ld32(tl + col_d + chunk * 32, v);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float x = fast_sigmoid(__uint_as_float(v[j]) + bias[chunk * 32 + j]);
        const float hi = tf32_round_fin(x);
        vh[j] = __float_as_uint(hi);
        v[j] = rna_tf32_mma(x - hi);
      }
      tmem_st32(tl + col_hi + chunk * 32, vh);
      tmem_st32(tl + col_d + chunk * 32, v);
    };

    float z[3], zn[3];
    int tile = blockIdx.x;
    if (tile < n_tiles) layer0(tile, z);
    for (; tile < n_tiles; tile += gridDim.x, ph ^= 1u) {
      // ---- epilogue 1: pose = sigmoid(D1 + b1) ----
      mbar_wait(smem_u32(&s_barD[0]), ph);
      tc_fence_after();
#pragma unroll 1
      for (int c = 0; c < 2; ++c) epi_sigmoid(COL_D1, COL_POSE_HI, s_b1, half * 2 + c);
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(smem_u32(&s_barA[0]));
      // ---- epilogue 2: c1 = sigmoid(D2 + b2) ----
      mbar_wait(smem_u32(&s_barD[1]), ph);
      tc_fence_after();
#pragma unroll 1
      for (int c = 0; c < 3; ++c) epi_sigmoid(COL_D2, COL_C1_HI, s_b2, half * 3 + c);
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(smem_u32(&s_barA[1]));
      // ---- epilogue 3: c2 = sigmoid(D3 + b3) ----
      mbar_wait(smem_u32(&s_barD[2]), ph);
      tc_fence_after();
#pragma unroll 1
      for (int c = 0; c < 2; ++c) epi_sigmoid(COL_D3, COL_C2_HI, s_b3, half * 2 + c);
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(smem_u32(&s_barA[2]));
      // ---- layer 0 of the next tile while layer 4 runs on the tensor core ([384,512) is free once layer 3 has completed) ----
      const int next = tile + gridDim.x;
      if (next < n_tiles) layer0(next, zn);
      // ---- epilogue 4: f1 = gelu(D4 + b4); x0 = gelu(F2 f1 + b5); posterior ----
      mbar_wait(smem_u32(&s_barD[3]), ph);
      tc_fence_after();
      float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f;
      {
        uint32_t v[32];
        tmem_ld32(tl + COL_D4 + half * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int ch = half * 32 + j;
          const float f = apply_act_fast<ACT_GELU>(__uint_as_float(v[j]) + s_b4[ch]);
          acc0 = fmaf(f, s_w5[ch], acc0);
          acc1 = fmaf(f, s_w5[64 + ch], acc1);
          acc2 = fmaf(f, s_w5[128 + ch], acc2);
        }
      }
      tc_fence_before();
      if (half == 1) {
        s_part[rit * 3] = acc0;
        s_part[rit * 3 + 1] = acc1;
        s_part[rit * 3 + 2] = acc2;
      }
      named_bar_sync(1, 256);
      if (half == 0) {
        const int pass = tile % a.n_pass;
        const int64_t row = (int64_t)(tile / a.n_pass) * 128 + rit;
        float x0[3] = {gelu_erf(acc0 + s_part[rit * 3] + s_b5[0]), gelu_erf(acc1 + s_part[rit * 3 + 1] + s_b5[1]),
                       gelu_erf(acc2 + s_part[rit * 3 + 2] + s_b5[2])};
        if (pass == 1) {
          if (a.guiding_out) {
            a.guiding_out[row * 3] = x0[0];
            a.guiding_out[row * 3 + 1] = x0[1];
            a.guiding_out[row * 3 + 2] = x0[2];
          }
        } else {
          if (a.clip) {
#pragma unroll
            for (int d = 0; d < 3; ++d) x0[d] = fminf(fmaxf(x0[d], -1.0f), 1.0f);
          }
          if (a.x0_out) {
            a.x0_out[row * 3] = x0[0];
            a.x0_out[row * 3 + 1] = x0[1];
            a.x0_out[row * 3 + 2] = x0[2];
          }
          if (a.sample_out) {
            const int64_t tt = a.t[row / NPTS];
            const float k1 = a.c1[tt], k2 = a.c2[tt];
            const float nz = tt != 0 ? 1.0f : 0.0f, sg = expf(0.5f * a.logvar[tt]);
#pragma unroll
            for (int d = 0; d < 3; ++d) {
              const float mean = k1 * x0[d] + k2 * z[d];
              a.sample_out[row * 3 + d] = mean + nz * sg * a.noise[row * 3 + d];
            }
          }
        }
      }
      named_bar_sync(1, 256);  // s_part is rewritten by the next tile
      z[0] = zn[0]; z[1] = zn[1]; z[2] = zn[2];
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

}  // namespace

// emb_hi / emb_lo: the step's embedding [rows,128] as hi / lo planes with leading dimension ld_emb; w*: TF32-rounded (hi) weight
// planes, w*_lo = residual planes.  Returns 1, or -1 when the TMA descriptors cannot be built (caller falls back to the GEMM chain).
int launch_x0net_fused(float* x, const float* pcd_out, const float* noise, float* sample_out, float* x0_out, float* guiding_out,
                       const int64_t* t, const float* c1, const float* c2, const float* logvar, int rows, int n_pass, int clip,
                       const float* emb_hi, const float* emb_lo, int64_t ld_emb, const float* w1, const float* w1_lo, const float* w2,
                       const float* w2_lo, const float* w3, const float* w3_lo, const float* w4, const float* w4_lo, const float* w0,
                       const float* b0, const float* b1, const float* b2, const float* b3, const float* b4, const float* w5,
                       const float* b5, cudaStream_t st) {
  if (rows <= 0 || (rows & 127) != 0 || n_pass < 1 || n_pass > 2) return -1;
  X0Maps m;
  const bool ok = tma_map_2d(&m.w1h, w1, 128, 64, 64, 128) && tma_map_2d(&m.w1l, w1_lo, 128, 64, 64, 128) &&
                  tma_map_2d(&m.w2h, w2, 192, 256, 256, 192) && tma_map_2d(&m.w2l, w2_lo, 192, 256, 256, 192) &&
                  tma_map_2d(&m.w3h, w3, 128, 192, 192, 128) && tma_map_2d(&m.w3l, w3_lo, 128, 192, 192, 128) &&
                  tma_map_2d(&m.w4h, w4, 64, 128, 128, 64) && tma_map_2d(&m.w4l, w4_lo, 64, 128, 128, 64) &&
                  tma_map_2d(&m.eh, emb_hi, rows, 128, ld_emb, 128) && tma_map_2d(&m.el, emb_lo, rows, 128, ld_emb, 128);
  if (!ok) return -1;
  constexpr int smem = NSLOT * SLOT + 1024;
  static PerDeviceOnce attr_done;
  if (smem_opt_in(attr_done, x0net_fused_kernel, smem) != cudaSuccess) return -1;
  X0Args a{x, pcd_out, noise, sample_out, x0_out, guiding_out, t, c1, c2, logvar, rows, n_pass, clip};
  X0Consts kc{w0, b0, b1, b2, b3, b4, w5, b5};
  const int n_tiles = rows / 128 * n_pass;
  const int sms = device_sm_count();
  const int grid = n_tiles < sms ? n_tiles : sms;
  x0net_fused_kernel<<<grid, XT, smem, st>>>(a, kc, m);
  return 1;
}

}  // namespace lsdm
