// navppo_tc.cu — tcgen05 tensor-core path of the PPO update (NAVPPO_BF16 / NAVPPO_BF16X3)
// + one-CTA GEMM self-tests.
//
// mlp_grad_tc_kernel is the tensor-core twin of mlp_grad_kernel (navppo_kernels.cu): one epoch
// body (forward, losses, backward) of one network per CTA row over tiles of 128 samples, with
// every matrix product on the 5th-generation tensor cores:
//
//   per hidden chunk c of 128 units (4 per residual block), M = 128 everywhere
//   forward   Z  = X  Wa_c^T            [128 s x 128 j]   K = IN (16 | 32)
//             U += H  Wb_c^T            [128 s x IN]      K = 128
//   backward  Z  (recomputed), GH = GU Wb_c   [128 s x 128 j]   K = IN
//             GX += GZ Wa_c             [128 s x IN]      K = 128        (block 2 only)
//             dWa_c  = GZ^T [X | 1]     [128 j x 48]      K = 128 samples (col 32 = bias gradient)
//             dWbT_c = H^T  GU          [128 j x IN]      K = 128 samples
//
// Operands are BF16 in shared memory in the dual-use row-block tile format of tc_common.cuh
// (every activation / weight tile is stored ONCE and read K-major by the products that
// contract over features / hidden units and MN-major by those that contract over samples);
// accumulators are fp32 in TMEM (480 of 512 columns; the weight-gradient accumulators are double buffered).  Arithmetic modes: PASSES = 1 plain
// bf16 x bf16; PASSES = 3 split operands x = hi + lo (two bf16 tiles) and hi*hi + lo*hi +
// hi*lo, about 16 mantissa bits (5e-6 relative in the self-test) at 3 MMAs per step.
// Elementwise work between the products (bias, LeakyReLU and its derivative, heads, losses)
// runs on the CUDA cores straight out of TMEM (tcgen05.ld, thread = sample row) and writes
// the next operand tile.  Weight chunks arrive as one TMA bulk copy (cp.async.bulk) from a
// per-epoch pre-split, pre-tiled copy of the parameters (tc_prep_weights_kernel).
// Weight-gradient chunks are read from TMEM (thread = hidden unit) and accumulated into the
// CTA's own row of the partial-gradient workspace, exactly like the CUDA-core kernel, so the
// deterministic grad_reduce_kernel / Adam path downstream is shared.
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "../../include/navppo.h"
#include "nav_common.h"
#include "ppo_common.cuh"
#include "tc_common.cuh"

namespace {

using namespace ppo;

extern __shared__ __align__(128) unsigned char tc_smem[];

// ----------------------------------------------------------------------------------------
// Pre-split, pre-tiled weights: per network 4 chunk blobs of block 1 then 4 of block 2, each
//   [Wa hi | Wa lo | Wb hi | Wb lo | ba fp32[128]]
// Wa tile: rows = hidden unit j (128), columns = input feature (IN);  Wb tile: rows = output
// feature o (IN), columns = hidden unit j (128) — both as stored by torch ([out, in]).
// ----------------------------------------------------------------------------------------
constexpr int CHUNK = 128;
constexpr int NCHUNK = HID / CHUNK;                         // 4
__host__ __device__ constexpr uint32_t wtile_bytes(int IN) { return (uint32_t)(CHUNK * IN * 2); }
__host__ __device__ constexpr uint32_t chunk_blob_bytes(int IN) { return 4u * wtile_bytes(IN) + CHUNK * 4u; }
constexpr uint32_t NET_BLOB = NCHUNK * (chunk_blob_bytes(OBS) + chunk_blob_bytes(X1));   // 200,704 B
__host__ __device__ constexpr uint32_t blob_offset(int block2, int c) {
  return block2 ? NCHUNK * chunk_blob_bytes(OBS) + (uint32_t)c * chunk_blob_bytes(X1) : (uint32_t)c * chunk_blob_bytes(OBS);
}

__global__ void tc_prep_weights_kernel(const float* __restrict__ params, unsigned char* __restrict__ wprep) {
  const int net = blockIdx.y;
  const float* p = params + (net ? NAVPPO_CRITIC_OFFSET : 0);
  unsigned char* out = wprep + (size_t)net * NET_BLOB;
  // one thread per fc1 weight of both blocks: 512*16 + 512*32 = 24576, same count for fc2
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < HID * (OBS + X1); t += gridDim.x * blockDim.x) {
    const int block2 = t >= HID * OBS;
    const int IN = block2 ? X1 : OBS;
    const int r = block2 ? t - HID * OBS : t;
    const int j = r / IN, i = r % IN;              // fc1: hidden unit j, input feature i
    const int c = j / CHUNK, jl = j % CHUNK;
    unsigned char* blob = out + blob_offset(block2, c);
    uint16_t h, l;
    tc::split_bf16(p[(block2 ? O_W2A : O_W1A) + j * IN + i], &h, &l);
    *reinterpret_cast<uint16_t*>(blob + tc::rb16_off(CHUNK, jl, i)) = h;
    *reinterpret_cast<uint16_t*>(blob + wtile_bytes(IN) + tc::rb16_off(CHUNK, jl, i)) = l;
    // fc2 [IN][512]: element (o = i, j)
    tc::split_bf16(p[(block2 ? O_W2B : O_W1B) + i * HID + j], &h, &l);
    *reinterpret_cast<uint16_t*>(blob + 2 * wtile_bytes(IN) + tc::rb16_off(IN, i, jl)) = h;
    *reinterpret_cast<uint16_t*>(blob + 3 * wtile_bytes(IN) + tc::rb16_off(IN, i, jl)) = l;
    if (i == 0) *reinterpret_cast<float*>(blob + 4 * wtile_bytes(IN) + jl * 4) = p[(block2 ? O_B2A : O_B1A) + j];
  }
}

// ----------------------------------------------------------------------------------------
// Shared-memory plan of mlp_grad_tc_kernel (bytes).  Activation tiles have 128 rows (samples);
// each has a hi part and a lo part (the lo parts are only touched when PASSES == 3).
// ----------------------------------------------------------------------------------------
constexpr int XCOLS = 48;                                   // x0 | y1 | 1 0 0 ... (bias-gradient column)
constexpr uint32_t ROWG = 128 * 16;                         // bytes of one 8-column group of a 128-row tile
constexpr uint32_t SX_PART = (XCOLS / 8) * ROWG;            // 12 KB
constexpr uint32_t SGU_PART = (X1 / 8) * ROWG;              // 8 KB
constexpr uint32_t SH_PART = (CHUNK / 8) * ROWG;            // 32 KB
constexpr uint32_t OFF_SX = 0;
constexpr uint32_t OFF_SGU = OFF_SX + 2 * SX_PART;          // 24 KB
constexpr uint32_t OFF_SH = OFF_SGU + 2 * SGU_PART;         // 40 KB
constexpr uint32_t OFF_SGZ = OFF_SH + 2 * SH_PART;          // 104 KB
constexpr uint32_t SWA_BYTES = 2 * wtile_bytes(X1);         // one fc1 weight chunk, hi | lo (16 KB)
constexpr uint32_t OFF_SWA = OFF_SGZ + 2 * SH_PART;         // 168 KB: TWO fc1 chunk buffers (prefetch ring)
constexpr uint32_t OFF_SWB = OFF_SWA + 2 * SWA_BYTES;       // one fc2 chunk buffer, hi | lo (16 KB)
constexpr uint32_t OFF_BIAS = OFF_SWB + SWA_BYTES;          // fc1 biases of both blocks, resident: 2 x 512 floats
constexpr uint32_t OFF_RED = OFF_BIAS + 2 * HID * 4;        // head / bias-b gradient partials [4 warps][128] floats
constexpr uint32_t OFF_MISC = OFF_RED + 4 * 128 * 4;        // barriers, tmem base
constexpr uint32_t TC_SMEM_BYTES = OFF_MISC + 64;           // 227,392 B of the 232,448 B a CTA may have

// TMEM columns
// the weight-gradient accumulators are double buffered (chunk c -> buffer c & 1) so that chunk c - 1 can be
// flushed while the tensor pipe works on chunk c
constexpr uint32_t TM_Z = 0, TM_GH = 128, TM_DWA = 256, TM_DWB = 304, TM_U = 336, TM_GX = 368, TM_DW2 = 144,
                   TM_COLS = 512;   // second buffer: TM_DWA + 144 = 400, TM_DWB + 144 = 448 .. 480

struct TcGradArgs {
  GradArgs g;
  const unsigned char* wprep;
};

// issue one product D[tmem d_col .. + N) (+)= A B over `ksteps` instructions (16 k each)
template <int PASSES>
__device__ __forceinline__ void issue_gemm(uint32_t tmem_d, uint32_t a_hi, uint32_t a_lo, uint32_t a_lbo, uint32_t a_sbo,
                                           uint32_t a_step, int a_mn, uint32_t b_hi, uint32_t b_lo, uint32_t b_lbo,
                                           uint32_t b_sbo, uint32_t b_step, int b_mn, int N, int ksteps, bool accumulate) {
  const uint32_t idesc = tc::make_idesc_bf16(128, N, a_mn, b_mn);
  uint32_t acc = accumulate ? 1u : 0u;
  for (int kk = 0; kk < ksteps; ++kk) {
    const uint64_t ah = tc::make_desc(a_hi + kk * a_step, a_lbo, a_sbo);
    const uint64_t bh = tc::make_desc(b_hi + kk * b_step, b_lbo, b_sbo);
    if (PASSES == 3) {  // small terms first
      const uint64_t al = tc::make_desc(a_lo + kk * a_step, a_lbo, a_sbo);
      const uint64_t bl = tc::make_desc(b_lo + kk * b_step, b_lbo, b_sbo);
      tc::mma_bf16(tmem_d, al, bh, idesc, acc); acc = 1u;
      tc::mma_bf16(tmem_d, ah, bl, idesc, acc);
    }
    tc::mma_bf16(tmem_d, ah, bh, idesc, acc); acc = 1u;
  }
}

// 8 consecutive columns of one row -> one 16-byte granule of the hi tile (and of the lo tile)
template <int PASSES>
__device__ __forceinline__ void store8(unsigned char* tile_hi, uint32_t part_bytes, int row, int col0, const float* v) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) tc::split_bf16x2(v[2 * i], v[2 * i + 1], &h[i], &l[i]);
  unsigned char* dst = tile_hi + (uint32_t)(col0 >> 3) * ROWG + (uint32_t)row * 16;
  *reinterpret_cast<uint4*>(dst) = make_uint4(h[0], h[1], h[2], h[3]);
  if (PASSES == 3) *reinterpret_cast<uint4*>(dst + part_bytes) = make_uint4(l[0], l[1], l[2], l[3]);
}

__device__ __forceinline__ void tma_load(unsigned char* dst, const unsigned char* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(tc::smem_u32(bar)), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   tc::smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(tc::smem_u32(bar))
               : "memory");
}

// fire-and-forget accumulation into the CTA's own partial-gradient row: the L2 does the fp32 add,
// so the thread never waits for the old value (every address has ONE writer, in tile order ->
// the sum is the same sequence of IEEE additions as a load-add-store)
__device__ __forceinline__ void red_add4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void red_add1(float* addr, float a) {
  asm volatile("red.global.add.f32 [%0], %1;" ::"l"(addr), "f"(a) : "memory");
}
// LeakyReLU(0.2) as one multiply and one max (x > 0.2 x exactly when x > 0)
__device__ __forceinline__ float lrelu_fast(float x) { return fmaxf(x, LEAK * x); }

template <int PASSES>
__global__ void __launch_bounds__(256, 1) mlp_grad_tc_kernel(TcGradArgs ta) {
  const GradArgs& a = ta.g;
  unsigned char* smem = tc_smem;
  unsigned char* sX = smem + OFF_SX;
  unsigned char* sGU = smem + OFF_SGU;
  unsigned char* sH = smem + OFF_SH;
  unsigned char* sGZ = smem + OFF_SGZ;
  unsigned char* sWa = smem + OFF_SWA;                                  // two buffers of SWA_BYTES
  unsigned char* sWb = smem + OFF_SWB;
  float* sBias = reinterpret_cast<float*>(smem + OFF_BIAS);
  float* sRed = reinterpret_cast<float*>(smem + OFF_RED);
  uint64_t* mbar = reinterpret_cast<uint64_t*>(smem + OFF_MISC);        // MMA completion
  uint64_t* wabar = reinterpret_cast<uint64_t*>(smem + OFF_MISC + 8);   // fc1 chunk arrival, one per buffer (2)
  uint64_t* wbbar = reinterpret_cast<uint64_t*>(smem + OFF_MISC + 24);  // fc2 chunk arrival
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + OFF_MISC + 32);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q = warp & 3, half = warp >> 2;          // TMEM lane quarter / column half of this warp
  const int row = q * 32 + lane;                     // sample row (chain epilogues) or hidden row (dW flush)
  const bool owner = half == 0;                      // threads 0..127 own one sample row each
  const int net = blockIdx.y;
  const float* __restrict__ p = a.params + (net ? NAVPPO_CRITIC_OFFSET : 0);
  const unsigned char* __restrict__ wblob = ta.wprep + (size_t)net * NET_BLOB;
  float* __restrict__ grow = a.gpart + ((size_t)net * gridDim.x + blockIdx.x) * NET_ROW;

  if (tid == 0) { tc::mbar_init(mbar, 1); tc::mbar_init(wabar, 1); tc::mbar_init(wabar + 1, 1); tc::mbar_init(wbbar, 1); }
  if (warp == 0) tc::tmem_alloc(tmem_slot, TM_COLS);
  for (int i = tid; i < NET_ROW; i += 256) grow[i] = 0.f;
  __threadfence();   // the zeros are in L2 before any red.add of this CTA
  for (int i = tid; i < 2 * HID; i += 256) sBias[i] = p[(i < HID ? O_B1A : O_B2A - HID) + i];
  // constant part of the X tile: columns 32..47 = (1, 0, 0, ...) in every row
  if (owner) {
    float ones[8] = {1.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, zeros[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    store8<3>(sX, SX_PART, row, 32, ones);
    store8<3>(sX, SX_PART, row, 40, zeros);
  }
  tc::fence_smem_to_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  const uint32_t trow = tmem + ((uint32_t)(q * 32) << 16);   // this warp's lane quarter
  uint32_t mphase = 0;
  uint32_t wa_ph[2] = {0, 0}, wb_ph = 0;             // used by thread 0 only (it issues TMA and MMA)

  // shared-memory addresses of the operand tiles
  const uint32_t aX = tc::smem_u32(sX), aGU = tc::smem_u32(sGU), aH = tc::smem_u32(sH), aGZ = tc::smem_u32(sGZ),
                 aWa = tc::smem_u32(sWa), aWb = tc::smem_u32(sWb);
  // Weight pipeline (thread 0): a chunk blob is [Wa hi | Wa lo | Wb hi | Wb lo | ba]; the fc1 half
  // goes to ring buffer c & 1, the fc2 half to the single fc2 buffer, each as one TMA bulk copy on
  // its own mbarrier.  A region is refilled as soon as the last MMA that reads it has completed,
  // i.e. one chunk ahead, so the copies fly behind the CUDA-core epilogues.
  auto load_wa = [&](int blk, int c) {
    const uint32_t wt = wtile_bytes(blk ? X1 : OBS);
    tma_load(sWa + (c & 1) * SWA_BYTES, wblob + blob_offset(blk, c), 2 * wt, wabar + (c & 1));
  };
  auto load_wb = [&](int blk, int c) {
    const uint32_t wt = wtile_bytes(blk ? X1 : OBS);
    tma_load(sWb, wblob + blob_offset(blk, c) + 2 * wt, 2 * wt, wbbar);
  };
  auto wait_wa = [&](int c) { tc::mbar_wait(wabar + (c & 1), wa_ph[c & 1]); wa_ph[c & 1] ^= 1; };
  auto wait_wb = [&]() { tc::mbar_wait(wbbar, wb_ph); wb_ph ^= 1; };

  double macc[4] = {0.0, 0.0, 0.0, 0.0};
  const int ntiles = (a.T + 127) / 128;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int si = tile * 128 + row;
    const bool valid = owner && si < a.T;
    float x1[X1], u1[OBS], u2[X1], gu2[X1];
    // ------------------------------------------------------------------ load x0, publish it
    if (owner) {
      if (valid) {
        const float4* o = reinterpret_cast<const float4*>(a.obs + (size_t)si * OBS);
#pragma unroll
        for (int k4 = 0; k4 < OBS / 4; ++k4) {
          const float4 t = o[k4];
          x1[4 * k4] = t.x; x1[4 * k4 + 1] = t.y; x1[4 * k4 + 2] = t.z; x1[4 * k4 + 3] = t.w;
        }
      } else {
#pragma unroll
        for (int k = 0; k < OBS; ++k) x1[k] = 0.f;
      }
      store8<PASSES>(sX, SX_PART, row, 0, x1);
      store8<PASSES>(sX, SX_PART, row, 8, x1 + 8);
    }
    tc::fence_smem_to_async();
    tc::fence_before_sync();
    __syncthreads();

    // ================================================================== forward, both blocks
#pragma unroll 1
    for (int blk = 0; blk < 2; ++blk) {
      const int IN = blk ? X1 : OBS;
      const uint32_t wt = wtile_bytes(IN);
      // Z = X Wa^T : A = X (rows = samples), B = Wa (rows = hidden units), K = IN
      auto issue_z = [&](int c) {
        const uint32_t w = aWa + (c & 1) * SWA_BYTES;
        issue_gemm<PASSES>(tmem + TM_Z, aX, aX + SX_PART, ROWG, 128, 2 * ROWG, 0, w, w + wt, ROWG, 128, 2 * ROWG, 0, 128,
                           IN / 16, false);
      };
      if (tid == 0) {
        load_wa(blk, 0);
        load_wb(blk, 0);
        wait_wa(0);
        tc::fence_after_sync();
        issue_z(0);
        tc::mma_commit(mbar);
      }
#pragma unroll 1
      for (int c = 0; c < NCHUNK; ++c) {
        tc::mbar_wait(mbar, mphase); mphase ^= 1;   // Z(c) is in TMEM (and U(c-1) has been accumulated)
        tc::fence_after_sync();
        if (tid == 0) {
          if (c > 0) load_wb(blk, c);               // U(c-1) done: the fc2 buffer is free
          if (c + 1 < NCHUNK) load_wa(blk, c + 1);  // ring slot (c+1)&1 was last read by Z(c-1)
        }
        {  // H = lrelu(Z + ba): this warp's 32 rows x 64 columns
          const float* sBa = sBias + blk * HID + c * CHUNK;
#pragma unroll
          for (int c0 = half * 64; c0 < half * 64 + 64; c0 += 16) {
            float v[16];
            tc::tmem_ld16(trow + TM_Z + c0, v);
#pragma unroll
            for (int i4 = 0; i4 < 4; ++i4) {
              const float4 b4 = *reinterpret_cast<const float4*>(sBa + c0 + 4 * i4);
              v[4 * i4] = lrelu_fast(v[4 * i4] + b4.x); v[4 * i4 + 1] = lrelu_fast(v[4 * i4 + 1] + b4.y);
              v[4 * i4 + 2] = lrelu_fast(v[4 * i4 + 2] + b4.z); v[4 * i4 + 3] = lrelu_fast(v[4 * i4 + 3] + b4.w);
            }
            store8<PASSES>(sH, SH_PART, row, c0, v);
            store8<PASSES>(sH, SH_PART, row, c0 + 8, v + 8);
          }
        }
        tc::fence_smem_to_async();
        tc::fence_before_sync();
        __syncthreads();
        if (tid == 0) {
          tc::fence_after_sync();
          wait_wb();
          // U += H Wb^T : A = H (rows = samples), B = Wb (rows = output features, R = IN), K = 128
          issue_gemm<PASSES>(tmem + TM_U, aH, aH + SH_PART, ROWG, 128, 2 * ROWG, 0, aWb, aWb + wt, (uint32_t)IN * 16, 128,
                             2u * IN * 16, 0, IN, CHUNK / 16, c > 0);
          if (c + 1 < NCHUNK) {                     // next chunk's Z rides on the same commit
            wait_wa(c + 1);
            issue_z(c + 1);
          }
          tc::mma_commit(mbar);
        }
      }
      tc::mbar_wait(mbar, mphase); mphase ^= 1;     // U complete
      tc::fence_after_sync();
      // block output: u = x + U + bb
      if (owner) {
        if (blk == 0) {
          float acc[16];
          tc::tmem_ld16(trow + TM_U, acc);
#pragma unroll
          for (int k = 0; k < OBS; ++k) {
            u1[k] = x1[k] + acc[k] + p[O_B1B + k];
            x1[OBS + k] = lrelu(u1[k]);
          }
          store8<PASSES>(sX, SX_PART, row, 16, x1 + 16);
          store8<PASSES>(sX, SX_PART, row, 24, x1 + 24);
        } else {
          float acc[16];
          tc::tmem_ld16(trow + TM_U, acc);
#pragma unroll
          for (int k = 0; k < 16; ++k) u2[k] = x1[k] + acc[k] + p[O_B2B + k];
          tc::tmem_ld16(trow + TM_U + 16, acc);
#pragma unroll
          for (int k = 0; k < 16; ++k) u2[16 + k] = x1[16 + k] + acc[k] + p[O_B2B + 16 + k];
        }
      }
      tc::fence_smem_to_async();
      tc::fence_before_sync();
      __syncthreads();
    }

    // ================================================================== heads, losses, dL/du2
    float go1 = 0.f, go2 = 0.f;
    if (owner) {
      float y2[X1];
#pragma unroll
      for (int k = 0; k < X1; ++k) y2[k] = lrelu(u2[k]);
      if (net == 0) {
        float o1 = p[O_HEAD + X1], o2 = p[O_HEAD + 2 * X1 + 1];
#pragma unroll
        for (int k = 0; k < X1; ++k) { o1 = fmaf(p[O_HEAD + k], y2[k], o1); o2 = fmaf(p[O_HEAD + X1 + 1 + k], y2[k], o2); }
        const float m0 = sigmoidf_(o1), m1 = tanhf(o2);
        if (valid) {
          const float2 av = reinterpret_cast<const float2*>(a.act)[si];
          const float lp = gauss_logp(av.x, av.y, m0, m1, a.var);
          const float lr = lp - a.logp_old[si];
          const float ratio = expf(lr);                                          // ppo.py:316
          const float A = a.adv[si];
          const float s1 = ratio * A;                                            // ppo.py:319
          const float s2 = fminf(fmaxf(ratio, 1.f - a.clip), 1.f + a.clip) * A;  // ppo.py:320
          macc[0] += (double)(-fminf(s1, s2));                                   // ppo.py:342
          macc[2] += (double)((ratio - 1.f) - lr);                               // ppo.py:326
          macc[3] += (fabsf(ratio - 1.f) > a.clip) ? 1.0 : 0.0;                  // ppo.py:335
          const float g_lp = (s1 <= s2 ? -A : 0.f) * a.inv_n * ratio;
          const float gm0 = g_lp * (av.x - m0) / a.var, gm1 = g_lp * (av.y - m1) / a.var;
          go1 = gm0 * m0 * (1.f - m0);
          go2 = gm1 * (1.f - m1 * m1);
        }
      } else {
        float v = p[O_HEAD + X1];
#pragma unroll
        for (int k = 0; k < X1; ++k) v = fmaf(p[O_HEAD + k], y2[k], v);
        if (valid) {
          const float d = v - a.rtg[si];
          macc[1] += (double)(d * d);                                            // ppo.py:343
          go1 = 2.f * d * a.inv_n;
        }
      }
#pragma unroll
      for (int k = 0; k < X1; ++k) {
        const float gy = (net == 0) ? fmaf(p[O_HEAD + k], go1, p[O_HEAD + X1 + 1 + k] * go2) : p[O_HEAD + k] * go1;
        gu2[k] = gy * dlrelu(u2[k]);
      }
#pragma unroll
      for (int k = 0; k < X1; k += 8) store8<PASSES>(sGU, SGU_PART, row, k, gu2 + k);
      // per-warp partial sums over the 32 samples: head weights / biases, then the fc2 bias of block 2
      const int nh = (net == 0) ? 2 : 1;
      for (int hd = 0; hd < nh; ++hd) {
        const float g = hd ? go2 : go1;
        float bsum = g;
        for (int o = 16; o > 0; o >>= 1) bsum += __shfl_xor_sync(0xffffffffu, bsum, o);
#pragma unroll
        for (int k = 0; k < X1; ++k) {
          float t = g * y2[k];
          for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
          if (lane == 0) sRed[warp * 128 + hd * (X1 + 1) + k] = t;
        }
        if (lane == 0) sRed[warp * 128 + hd * (X1 + 1) + X1] = bsum;
      }
#pragma unroll
      for (int k = 0; k < X1; ++k) {
        float t = gu2[k];
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (lane == 0) sRed[warp * 128 + 72 + k] = t;
      }
    }
    tc::fence_smem_to_async();
    tc::fence_before_sync();
    __syncthreads();
    {
      const int nhead = (net == 0) ? ACTOR_HEAD : CRITIC_HEAD;
      if (tid < nhead) grow[O_HEAD + tid] += ((sRed[tid] + sRed[128 + tid]) + sRed[256 + tid]) + sRed[384 + tid];
      if (tid >= 72 && tid < 72 + X1)
        grow[O_B2B + tid - 72] += ((sRed[tid] + sRed[128 + tid]) + sRed[256 + tid]) + sRed[384 + tid];
    }

    // ================================================================== backward, block 2 then block 1
    float gu1[OBS];
#pragma unroll 1
    for (int blk = 1; blk >= 0; --blk) {
      const int IN = blk ? X1 : OBS;
      const uint32_t wt = wtile_bytes(IN);
      const int o_wa = blk ? O_W2A : O_W1A, o_ba = blk ? O_B2A : O_B1A, o_wb = blk ? O_W2B : O_W1B;
      // phase 1 of chunk c: Z (recomputed) and GH = GU Wb (Wb read MN-major: rows = K = output
      // features, R = IN; columns = hidden units = N)
      auto issue_phase1 = [&](int c) {
        const uint32_t w = aWa + (c & 1) * SWA_BYTES;
        issue_gemm<PASSES>(tmem + TM_Z, aX, aX + SX_PART, ROWG, 128, 2 * ROWG, 0, w, w + wt, ROWG, 128, 2 * ROWG, 0, 128,
                           IN / 16, false);
        issue_gemm<PASSES>(tmem + TM_GH, aGU, aGU + SGU_PART, ROWG, 128, 2 * ROWG, 0, aWb, aWb + wt, 128, (uint32_t)IN * 16,
                           256, 1, 128, IN / 16, false);
      };
      // weight-gradient chunk c out of TMEM (lane = hidden unit j of the chunk) into the CTA's row
      auto flush_dw = [&](int c) {
        const int j = c * CHUNK + row;
        if (half == 0) {
          float* ga = grow + o_wa + j * IN;
          for (int c0 = 0; c0 < IN; c0 += 16) {
            float v[16];
            tc::tmem_ld16(trow + TM_DWA + (c & 1) * TM_DW2 + c0, v);
#pragma unroll
            for (int i = 0; i < 4; ++i) red_add4(ga + c0 + 4 * i, v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
          }
          float v[16];
          tc::tmem_ld16(trow + TM_DWA + (c & 1) * TM_DW2 + 32, v);       // column 32 = sum over samples of g_z = bias gradient
          red_add1(grow + o_ba + j, v[0]);
        } else {
          float* gb = grow + o_wb + j * IN;            // kernel layout: fc2 transposed
          for (int c0 = 0; c0 < IN; c0 += 16) {
            float v[16];
            tc::tmem_ld16(trow + TM_DWB + (c & 1) * TM_DW2 + c0, v);
#pragma unroll
            for (int i = 0; i < 4; ++i) red_add4(gb + c0 + 4 * i, v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
          }
        }
      };
      if (tid == 0) {
        load_wa(blk, 0);
        load_wb(blk, 0);
        wait_wa(0);
        wait_wb();
        tc::fence_after_sync();
        issue_phase1(0);
        tc::mma_commit(mbar);
      }
#pragma unroll 1
      for (int c = 0; c < NCHUNK; ++c) {
        tc::mbar_wait(mbar, mphase); mphase ^= 1;   // phase 1 of c (and phase 2 of c-1) complete
        tc::fence_after_sync();
        if (tid == 0 && c + 1 < NCHUNK) {
          load_wb(blk, c + 1);                      // GH(c) done: the fc2 buffer is free
          load_wa(blk, c + 1);                      // ring slot (c+1)&1 was last read by Z / GX of c-1
        }
        {  // H = lrelu(z), GZ = GH * lrelu'(z), z = Z + ba
          const float* sBa = sBias + blk * HID + c * CHUNK;
#pragma unroll
          for (int c0 = half * 64; c0 < half * 64 + 64; c0 += 16) {
            float z[16], g[16];
            tc::tmem_ld16(trow + TM_Z + c0, z);
            tc::tmem_ld16(trow + TM_GH + c0, g);
#pragma unroll
            for (int i4 = 0; i4 < 4; ++i4) {
              const float4 b4 = *reinterpret_cast<const float4*>(sBa + c0 + 4 * i4);   // one broadcast load per 4 biases
              const float bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const int i = 4 * i4 + k;
                const float zz = z[i] + bb[k];
                const float sl = zz > 0.f ? 1.f : LEAK;       // LeakyReLU slope: h = zz * slope, g_z = g_h * slope
                z[i] = zz * sl;
                g[i] = g[i] * sl;
              }
            }
            store8<PASSES>(sH, SH_PART, row, c0, z);
            store8<PASSES>(sH, SH_PART, row, c0 + 8, z + 8);
            store8<PASSES>(sGZ, SH_PART, row, c0, g);
            store8<PASSES>(sGZ, SH_PART, row, c0 + 8, g + 8);
          }
        }
        tc::fence_smem_to_async();
        tc::fence_before_sync();
        __syncthreads();
        if (tid == 0) {
          tc::fence_after_sync();
          if (blk) {  // GX += GZ Wa : A = GZ (rows = samples, K = hidden), B = Wa read MN-major (rows = K = hidden)
            const uint32_t w = aWa + (c & 1) * SWA_BYTES;
            issue_gemm<PASSES>(tmem + TM_GX, aGZ, aGZ + SH_PART, ROWG, 128, 2 * ROWG, 0, w, w + wt, 128, ROWG, 256, 1, IN,
                               CHUNK / 16, c > 0);
          }
          // dWa = GZ^T [X | 1] : both operands MN-major (rows = K = samples)
          issue_gemm<PASSES>(tmem + TM_DWA + (c & 1) * TM_DW2, aGZ, aGZ + SH_PART, 128, ROWG, 256, 1, aX, aX + SX_PART, 128, ROWG, 256, 1, XCOLS,
                             128 / 16, false);
          // dWbT = H^T GU
          issue_gemm<PASSES>(tmem + TM_DWB + (c & 1) * TM_DW2, aH, aH + SH_PART, 128, ROWG, 256, 1, aGU, aGU + SGU_PART, 128, ROWG, 256, 1, IN,
                             128 / 16, false);
          if (c + 1 < NCHUNK) {                     // next chunk's phase 1 rides on the same commit
            wait_wa(c + 1);
            wait_wb();
            issue_phase1(c + 1);
          }
          tc::mma_commit(mbar);
        }
        // the previous chunk's weight gradients sit complete in the other TMEM buffer: flush them
        // while the tensor pipe runs the MMAs just issued
        if (c > 0) flush_dw(c - 1);
      }
      tc::mbar_wait(mbar, mphase); mphase ^= 1;     // phase 2 of the last chunk complete
      tc::fence_after_sync();
      flush_dw(NCHUNK - 1);
      tc::fence_before_sync();
      __syncthreads();
      if (blk) {
        // dL/dx1 = g_u2 (skip connection) + GX; dL/du1 = dL/dy1 * lrelu'(u1); publish GU1 for block 1
        if (owner) {
          float acc[16];
          tc::tmem_ld16(trow + TM_GX + 16, acc);
#pragma unroll
          for (int k = 0; k < OBS; ++k) gu1[k] = (gu2[OBS + k] + acc[k]) * dlrelu(u1[k]);
          store8<PASSES>(sGU, SGU_PART, row, 0, gu1);
          store8<PASSES>(sGU, SGU_PART, row, 8, gu1 + 8);
#pragma unroll
          for (int k = 0; k < OBS; ++k) {
            float t = gu1[k];
            for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
            if (lane == 0) sRed[warp * 128 + 104 + k] = t;
          }
        }
        tc::fence_smem_to_async();
        tc::fence_before_sync();
        __syncthreads();
        if (tid >= 104 && tid < 104 + OBS)
          grow[O_B1B + tid - 104] += ((sRed[tid] + sRed[128 + tid]) + sRed[256 + tid]) + sRed[384 + tid];
      }
    }
    __syncthreads();   // sRed / tiles are rewritten by the next tile
  }

  // per-CTA metric partials: fixed-order sum over the row owners
  __syncthreads();
  double* red = reinterpret_cast<double*>(sH);
  if (owner) {
#pragma unroll
    for (int k = 0; k < 4; ++k) red[k * 128 + tid] = macc[k];
  }
  __syncthreads();
  if (tid < 4) {
    double t = 0.0;
    for (int m = 0; m < 128; ++m) t += red[tid * 128 + m];
    a.mpart[((size_t)net * gridDim.x + blockIdx.x) * 4 + tid] = t;
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, TM_COLS);
}

// ----------------------------------------------------------------------------------------
// Self-test: D[128, N] = A[128, K] * B[N, K]^T with every operand role the fused kernel uses.
// a_mode / b_mode: 0 = rows are the M/N index ("K-major"), 1 = rows are K ("MN-major").
// The descriptor strides are passed in so a test can probe the hardware's conventions.
// ----------------------------------------------------------------------------------------
struct SelfTestArgs {
  const float* A;   // [128, K] row-major
  const float* B;   // [N, K] row-major
  float* D;         // [128, N]
  int N, K, a_mode, b_mode;
  uint32_t a_lbo, a_sbo, a_kstep, b_lbo, b_sbo, b_kstep, a_major, b_major;
};

__global__ void __launch_bounds__(128) tc_selftest_kernel(SelfTestArgs p) {
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  unsigned char* sA = tc_smem;
  unsigned char* sB = tc_smem + 128 * p.K * 4;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) tc::mbar_init(&bar, 1);
  if (warp == 0) tc::tmem_alloc(&tmem_base_s, 256);
  for (int i = tid; i < 128 * p.K; i += 128) {
    const int m = i / p.K, k = i % p.K;
    const uint32_t off = p.a_mode == 0 ? tc::rb_off(128, m, k) : p.a_mode == 1 ? tc::rb_off(p.K, k, m) : 4u * i;
    *reinterpret_cast<float*>(sA + off) = tc::to_tf32(p.A[i]);
  }
  for (int i = tid; i < p.N * p.K; i += 128) {
    const int n = i / p.K, k = i % p.K;
    const uint32_t off = p.b_mode == 0 ? tc::rb_off(p.N, n, k) : p.b_mode == 1 ? tc::rb_off(p.K, k, n) : 4u * i;
    *reinterpret_cast<float*>(sB + off) = tc::to_tf32(p.B[i]);
  }
  tc::fence_smem_to_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = tmem_base_s;
  if (tid == 0) {
    const uint32_t idesc = tc::make_idesc_tf32(128, p.N, p.a_major, p.b_major);
    const uint32_t a0 = tc::smem_u32(sA), b0 = tc::smem_u32(sB);
    for (int kk = 0; kk < p.K / 8; ++kk)
      tc::mma_tf32(tmem, tc::make_desc(a0 + kk * p.a_kstep, p.a_lbo, p.a_sbo),
                   tc::make_desc(b0 + kk * p.b_kstep, p.b_lbo, p.b_sbo), idesc, kk > 0);
    tc::mma_commit(&bar);
  }
  tc::mbar_wait(&bar, 0);
  tc::fence_after_sync();
  for (int c0 = 0; c0 < p.N; c0 += 16) {
    float v[16];
    tc::tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
#pragma unroll
    for (int i = 0; i < 16; ++i) p.D[(size_t)tid * p.N + c0 + i] = v[i];
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 256);
}

// Same self-test with BF16 operands (kind::f16): `passes` = 1 (hi * hi) or 3 (hi*hi + lo*hi +
// hi*lo on the split x = hi + lo).  a_mode / b_mode: 0 = rows are the M/N index, 1 = rows are K.
struct SelfTest16Args {
  const float* A; const float* B; float* D;
  int N, K, a_mode, b_mode, passes;
};

__global__ void __launch_bounds__(128) tc_selftest_bf16_kernel(SelfTest16Args p) {
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const uint32_t a_bytes = 128 * p.K * 2, b_bytes = p.N * p.K * 2;
  unsigned char* sAh = tc_smem;
  unsigned char* sAl = sAh + a_bytes;
  unsigned char* sBh = sAl + a_bytes;
  unsigned char* sBl = sBh + b_bytes;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) tc::mbar_init(&bar, 1);
  if (warp == 0) tc::tmem_alloc(&tmem_base_s, 256);
  for (int i = tid; i < 128 * p.K; i += 128) {
    const int m = i / p.K, k = i % p.K;
    const uint32_t off = p.a_mode == 0 ? tc::rb16_off(128, m, k) : tc::rb16_off(p.K, k, m);
    uint16_t h, l;
    tc::split_bf16(p.A[i], &h, &l);
    *reinterpret_cast<uint16_t*>(sAh + off) = h;
    *reinterpret_cast<uint16_t*>(sAl + off) = l;
  }
  for (int i = tid; i < p.N * p.K; i += 128) {
    const int n = i / p.K, k = i % p.K;
    const uint32_t off = p.b_mode == 0 ? tc::rb16_off(p.N, n, k) : tc::rb16_off(p.K, k, n);
    uint16_t h, l;
    tc::split_bf16(p.B[i], &h, &l);
    *reinterpret_cast<uint16_t*>(sBh + off) = h;
    *reinterpret_cast<uint16_t*>(sBl + off) = l;
  }
  tc::fence_smem_to_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = tmem_base_s;
  if (tid == 0) {
    const uint32_t idesc = tc::make_idesc_bf16(128, p.N, p.a_mode, p.b_mode);
    // rows = M/N index: SBO = 128, LBO = R*16, 16 k per instruction = 2 granule columns
    // rows = K:         SBO = R*16, LBO = 128, 16 k per instruction = 2 groups of 8 rows
    const uint32_t a_lbo = p.a_mode == 0 ? 128 * 16 : 128, a_sbo = p.a_mode == 0 ? 128 : p.K * 16;
    const uint32_t a_step = p.a_mode == 0 ? 2 * 128 * 16 : 256;
    const uint32_t b_lbo = p.b_mode == 0 ? p.N * 16 : 128, b_sbo = p.b_mode == 0 ? 128 : p.K * 16;
    const uint32_t b_step = p.b_mode == 0 ? 2 * p.N * 16 : 256;
    uint32_t acc = 0;
    for (int kk = 0; kk < p.K / 16; ++kk) {
      const uint64_t ah = tc::make_desc(tc::smem_u32(sAh) + kk * a_step, a_lbo, a_sbo);
      const uint64_t al = tc::make_desc(tc::smem_u32(sAl) + kk * a_step, a_lbo, a_sbo);
      const uint64_t bh = tc::make_desc(tc::smem_u32(sBh) + kk * b_step, b_lbo, b_sbo);
      const uint64_t bl = tc::make_desc(tc::smem_u32(sBl) + kk * b_step, b_lbo, b_sbo);
      if (p.passes == 3) {   // small terms first
        tc::mma_bf16(tmem, al, bh, idesc, acc); acc = 1;
        tc::mma_bf16(tmem, ah, bl, idesc, acc);
      }
      tc::mma_bf16(tmem, ah, bh, idesc, acc); acc = 1;
    }
    tc::mma_commit(&bar);
  }
  tc::mbar_wait(&bar, 0);
  tc::fence_after_sync();
  for (int c0 = 0; c0 < p.N; c0 += 16) {
    float v[16];
    tc::tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
#pragma unroll
    for (int i = 0; i < 16; ++i) p.D[(size_t)tid * p.N + c0 + i] = v[i];
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 256);
}

}  // namespace

// ---- internal entry points used by navppo_kernels.cu -----------------------------------
size_t navppo_tc_prep_bytes() { return (size_t)2 * NET_BLOB; }

int navppo_tc_init() {
  NAV_CUDA_TRY(cudaFuncSetAttribute(mlp_grad_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM_BYTES));
  NAV_CUDA_TRY(cudaFuncSetAttribute(mlp_grad_tc_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM_BYTES));
  return NAVSIM_OK;
}

// One gradient pass on the tensor cores: re-tile the current weights, then the fused kernel.
// `rows` CTAs per network; fills a.gpart / a.mpart like mlp_grad_kernel.
int navppo_tc_grad_launch(const ppo::GradArgs& a, int rows, int passes, float* wprep, cudaStream_t s) {
  tc_prep_weights_kernel<<<dim3(24, 2), 256, 0, s>>>(a.params, reinterpret_cast<unsigned char*>(wprep));
  TcGradArgs ta{a, reinterpret_cast<const unsigned char*>(wprep)};
  if (passes == 3) mlp_grad_tc_kernel<3><<<dim3(rows, 2), 256, TC_SMEM_BYTES, s>>>(ta);
  else mlp_grad_tc_kernel<1><<<dim3(rows, 2), 256, TC_SMEM_BYTES, s>>>(ta);
  NAV_CUDA_TRY(cudaGetLastError());
  return NAVSIM_OK;
}

namespace {
}  // namespace

extern "C" {

int navppo_tc_selftest_bf16(const float* A, const float* B, float* D, int32_t N, int32_t K, int32_t a_mode, int32_t b_mode,
                            int32_t passes, void* stream) {
  if (!A || !B || !D) return nav_fail(NAVSIM_EINVAL, "null buffer");
  if (N < 16 || N > 256 || N % 16 || K < 16 || K > 128 || K % 16) return nav_fail(NAVSIM_EINVAL, "N in 16..256 step 16, K in 16..128 step 16");
  if (passes != 1 && passes != 3) return nav_fail(NAVSIM_EINVAL, "passes is 1 or 3");
  SelfTest16Args p{A, B, D, N, K, a_mode, b_mode, passes};
  const size_t smem = (size_t)(128 + N) * K * 4;
  NAV_CUDA_TRY(cudaFuncSetAttribute(tc_selftest_bf16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  tc_selftest_bf16_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(p);
  NAV_CUDA_TRY(cudaGetLastError());
  return NAVSIM_OK;
}

int navppo_tc_selftest(const float* A, const float* B, float* D, int32_t N, int32_t K, int32_t a_mode, int32_t b_mode,
                       const uint32_t* strides6, void* stream) {
  if (!A || !B || !D || !strides6) return nav_fail(NAVSIM_EINVAL, "null buffer");
  if (N < 16 || N > 256 || N % 16 || K < 8 || K > 128 || K % 8) return nav_fail(NAVSIM_EINVAL, "N in 16..256 step 16, K in 8..128 step 8");
  // strides6[6], [7]: major bits of the instruction descriptor (0 = K-major, 1 = MN-major).
  // a_mode / b_mode == 2: the global buffer is copied verbatim into shared memory (raw image).
  SelfTestArgs p{A, B, D, N, K, a_mode, b_mode, strides6[0], strides6[1], strides6[2], strides6[3], strides6[4], strides6[5],
                 strides6[6], strides6[7]};
  const size_t smem = (size_t)(128 + N) * K * 4;
  NAV_CUDA_TRY(cudaFuncSetAttribute(tc_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  tc_selftest_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(p);
  NAV_CUDA_TRY(cudaGetLastError());
  return NAVSIM_OK;
}

}  // extern "C"
