// navppo_tcws.cu — warp-specialised tcgen05 gradient kernel of the PPO update (the default for
// NAVPPO_BF16 / NAVPPO_BF16X3; navppo_tc.cu keeps the first, single-role version as a bit-exact
// cross-check).
//
// Same arithmetic as mlp_grad_tc_kernel (one epoch body — forward, losses, backward — of one network
// per CTA row over tiles of 128 samples, ppo.py:305-397, net_actor.py:16-53), restructured so that the
// tensor pipe and the CUDA cores work at the same time:
//
//   warps 0-7   epilogue: TMEM -> bias / LeakyReLU / derivative -> bf16 hi|lo split -> operand tiles in
//               shared memory; warps 0-3 also own one sample row each (block outputs, heads, losses)
//   warps 8-11  weight-gradient flush: TMEM (lane = hidden unit) -> red.global.add into the CTA's
//               partial-gradient row
//   warp 12     the ONE thread that issues every tcgen05.mma
//   warp 13     the ONE thread that streams pre-tiled weight half-chunks through a 3-slot TMA ring
//
// The hidden layer is walked in half-chunks of 64 units.  Z / GH accumulators live in a 2-slot TMEM
// ring, so while the epilogue warps turn half-chunk h into operand tiles, the tensor pipe already
// computes Z / GH of half-chunk h + 1 and the U / GX products of half-chunk h - 1.  The weight-gradient
// products (M = 128 hidden units of a whole chunk, K = 128 samples) read the complete H / GZ tiles and
// are the one phase that cannot overlap the epilogue of the next chunk (the tiles are single
// buffered: 128 KB of the 227 KB); their accumulators are flushed by the flush warps under the next
// chunk's epilogue.  All hand-offs are mbarriers (tcgen05.commit on the tensor side, one elected
// arrive per warp on the CUDA-core side); there is no CTA-wide barrier inside the tile loop.
//
// Every product is issued in the same k-order and pass order as the first kernel, every reduction
// keeps its order, so the two kernels return bit-identical gradients (tests/test_ppo_tc_gpu.py).
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "../../include/navppo.h"
#include "nav_common.h"
#include "ppo_common.cuh"
#include "tc_common.cuh"

namespace {

using namespace ppo;

extern __shared__ __align__(128) unsigned char ws_smem[];

// ----------------------------------------------------------------------------------------
// Pre-split, pre-tiled weights: per network 8 half-chunk blobs of block 1 then 8 of block 2, each
//   [Wa hi | Wa lo | Wb hi | Wb lo]
// Wa tile: rows = hidden unit j (64), columns = input feature (IN);  Wb tile: rows = output feature o
// (IN), columns = hidden unit j (64) — the bf16 row-block format of tc_common.cuh.
// ----------------------------------------------------------------------------------------
constexpr int HC = 64;                                    // hidden units per half-chunk
constexpr int NHC = HID / HC;                             // 8
__host__ __device__ constexpr uint32_t wpart(int IN) { return (uint32_t)(HC * IN * 2); }
__host__ __device__ constexpr uint32_t wblob(int IN) { return 4u * wpart(IN); }
constexpr uint32_t WS_NET_BLOB = NHC * (wblob(OBS) + wblob(X1));   // 196,608 B
__host__ __device__ constexpr uint32_t wblob_off(int block2, int h) {
  return block2 ? NHC * wblob(OBS) + (uint32_t)h * wblob(X1) : (uint32_t)h * wblob(OBS);
}

__global__ void ws_prep_weights_kernel(const float* __restrict__ params, unsigned char* __restrict__ wprep) {
  const int net = blockIdx.y;
  const float* p = params + (net ? NAVPPO_CRITIC_OFFSET : 0);
  unsigned char* out = wprep + (size_t)net * WS_NET_BLOB;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < HID * (OBS + X1); t += gridDim.x * blockDim.x) {
    const int block2 = t >= HID * OBS;
    const int IN = block2 ? X1 : OBS;
    const int r = block2 ? t - HID * OBS : t;
    const int j = r / IN, i = r % IN;              // fc1: hidden unit j, input feature i
    const int h = j / HC, jl = j % HC;
    unsigned char* blob = out + wblob_off(block2, h);
    const uint32_t wp = wpart(IN);
    uint16_t hi, lo;
    tc::split_bf16(p[(block2 ? O_W2A : O_W1A) + j * IN + i], &hi, &lo);
    *reinterpret_cast<uint16_t*>(blob + tc::rb16_off(HC, jl, i)) = hi;
    *reinterpret_cast<uint16_t*>(blob + wp + tc::rb16_off(HC, jl, i)) = lo;
    // fc2 [IN][512]: element (o = i, j)
    tc::split_bf16(p[(block2 ? O_W2B : O_W1B) + i * HID + j], &hi, &lo);
    *reinterpret_cast<uint16_t*>(blob + 2 * wp + tc::rb16_off(IN, i, jl)) = hi;
    *reinterpret_cast<uint16_t*>(blob + 3 * wp + tc::rb16_off(IN, i, jl)) = lo;
  }
}

// ----------------------------------------------------------------------------------------
// Shared-memory plan (bytes).  Activation tiles have 128 rows (samples), a hi part and a lo part.
// ----------------------------------------------------------------------------------------
constexpr int XCOLS = 48;                                   // x0 | y1 | 1 0 0 ... (bias-gradient column)
constexpr uint32_t ROWG = 128 * 16;                         // bytes of one 8-column group of a 128-row tile
constexpr uint32_t SX_PART = (XCOLS / 8) * ROWG;            // 12 KB
constexpr uint32_t SGU_PART = (X1 / 8) * ROWG;              // 8 KB
constexpr uint32_t SH_PART = (128 / 8) * ROWG;              // 32 KB: 128 hidden columns = one chunk
constexpr uint32_t OFF_SX = 0;
constexpr uint32_t OFF_SGU = OFF_SX + 2 * SX_PART;          // 24 KB
constexpr uint32_t OFF_SH = OFF_SGU + 2 * SGU_PART;         // 40 KB
constexpr uint32_t OFF_SGZ = OFF_SH + 2 * SH_PART;          // 104 KB
constexpr uint32_t WSLOT = wblob(X1);                       // 16 KB: one weight half-chunk
constexpr int NWSLOT = 3;
constexpr uint32_t OFF_W = OFF_SGZ + 2 * SH_PART;           // 168 KB
constexpr uint32_t OFF_BIAS = OFF_W + NWSLOT * WSLOT;       // 216 KB: fc1 biases of both blocks, 2 x 512 floats
constexpr uint32_t OFF_RED = OFF_BIAS + 2 * HID * 4;        // head / bias-b gradient partials [4 warps][128] floats
constexpr uint32_t OFF_PAR = OFF_RED + 4 * 128 * 4;         // fc2 biases and head weights, 128 floats
constexpr int P_B1B = 0, P_B2B = OBS, P_HEAD = OBS + X1;    // offsets inside that block
constexpr uint32_t OFF_BAR = OFF_PAR + 128 * 4;             // mbarriers, tmem base
constexpr uint32_t WS_SMEM_BYTES = OFF_BAR + 256;           // 227,584 B of the 232,448 B a CTA may have

// mbarrier indices
enum { B_ZFULL = 0, B_EFULL = 2, B_HFREE = 4, B_WFULL = 6, B_WFREE = 9, B_DWFULL = 12, B_DWFREE = 13, B_ACC = 14,
       B_XREADY = 15, B_COUNT = 16 };

// TMEM columns: Z / GH ring of two 64-column slots, weight-gradient accumulators, U and GX
constexpr uint32_t TM_ZG = 0;        // slot s: Z at s * 128, GH at s * 128 + 64
constexpr uint32_t TM_DWA = 256;     // 48 columns
constexpr uint32_t TM_DWB = 352;     // 32 columns
constexpr uint32_t TM_U = 416;       // 32 columns
constexpr uint32_t TM_GX = 448;      // 32 columns
constexpr uint32_t TM_COLS = 512;

constexpr int WS_THREADS = 512;
constexpr int W_FLUSH0 = 8, W_MMA = 12, W_TMA = 13;

struct WsGradArgs {
  GradArgs g;
  const unsigned char* wprep;
  long long* prof;    // diagnostic cycle counters (PROF variant only)
};

// cycle accounting of one thread per role (CTA (0, 0) only): every PT(i) charges the cycles since the
// previous mark to category i
#define PT(i)                                   \
  do {                                          \
    if (PROF) {                                 \
      const long long t__ = clock64();          \
      pc[i] += t__ - plast;                     \
      plast = t__;                              \
      if (tbuf && tn < 500) tbuf[tn++] = ((long long)(i) << 56) | (t__ & 0x00FFFFFFFFFFFFFFLL); \
    }                                           \
  } while (0)

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc::smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void named_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
template <int N>
__device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N>
__device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }

// Shared-memory descriptors are assembled from 32-bit words at issue time so that the issuing thread only
// keeps tile base addresses in registers: low word = start >> 4 | (LBO >> 4) << 16, high word =
// SBO >> 4 | version 1 (bit 46 of the descriptor).  Addresses and strides below are in 16-byte units.
struct Opnd {
  uint32_t start;   // hi-part tile start >> 4
  uint32_t part;    // offset of the lo part >> 4
  uint32_t lbo, sbo, step;   // >> 4; step = advance per 16-k instruction
};
// one lane of a converged warp (the same lane every time): the tcgen05 issue pattern
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ uint64_t words_desc(uint32_t lo, uint32_t sbo) {
  return ((uint64_t)(sbo | 0x4000u) << 32) | (uint64_t)lo;
}

// D (+)= A B over KSTEPS instructions (16 k each); executed by a converged warp, issued by one lane
template <int PASSES, int KSTEPS>
__device__ __forceinline__ void gemm(uint32_t tmem_d, const Opnd& A, const Opnd& B, uint32_t idesc, bool accumulate) {
  if (elect_one()) {
    uint32_t acc = accumulate ? 1u : 0u;
    const uint32_t a0 = A.start | (A.lbo << 16), b0 = B.start | (B.lbo << 16);
#pragma unroll
    for (int kk = 0; kk < KSTEPS; ++kk) {
      const uint64_t ah = words_desc(a0 + kk * A.step, A.sbo), bh = words_desc(b0 + kk * B.step, B.sbo);
      if (PASSES == 3) {  // small terms first
        tc::mma_bf16(tmem_d, words_desc(a0 + kk * A.step + A.part, A.sbo), bh, idesc, acc); acc = 1u;
        tc::mma_bf16(tmem_d, ah, words_desc(b0 + kk * B.step + B.part, B.sbo), idesc, acc);
      }
      tc::mma_bf16(tmem_d, ah, bh, idesc, acc); acc = 1u;
    }
  }
  __syncwarp();
}
// D (+)= A B over the 64 hidden units of a half-chunk with A in tensor memory, as the epilogue warps leave
// it: per 32 hidden units [16 packed hi columns | 16 packed lo columns], 8 packed columns per instruction
template <int PASSES>
__device__ __forceinline__ void gemm_ts(uint32_t tmem_d, uint32_t tmem_a, const Opnd& B, uint32_t idesc, bool accumulate) {
  if (elect_one()) {
    uint32_t acc = accumulate ? 1u : 0u;
    const uint32_t b0 = B.start | (B.lbo << 16);
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      const uint32_t ah = tmem_a + 32 * (kk >> 1) + 8 * (kk & 1);
      const uint64_t bh = words_desc(b0 + kk * B.step, B.sbo);
      if (PASSES == 3) {  // small terms first
        tc::mma_bf16_ts(tmem_d, ah + 16, bh, idesc, acc); acc = 1u;
        tc::mma_bf16_ts(tmem_d, ah, words_desc(b0 + kk * B.step + B.part, B.sbo), idesc, acc);
      }
      tc::mma_bf16_ts(tmem_d, ah, bh, idesc, acc); acc = 1u;
    }
  }
  __syncwarp();
}
// tcgen05.commit by the lane that issues the MMAs
__device__ __forceinline__ void commit(uint64_t* bar) {
  if (elect_one()) tc::mma_commit(bar);
  __syncwarp();
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

__device__ __forceinline__ void red_add4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void red_add1(float* addr, float a) {
  asm volatile("red.global.add.f32 [%0], %1;" ::"l"(addr), "f"(a) : "memory");
}
__device__ __forceinline__ float lrelu_fast(float x) { return fmaxf(x, LEAK * x); }

// Sums over the 32 lanes of a warp of V per-lane values at once: every round hands half of the values to
// the partner lane (the pairs and their order are those of the xor butterfly 16, 8, 4, 2, 1, so each
// total is the same sequence of additions), and lane l ends up with the total of value l * V / 32.
template <int V>
__device__ __forceinline__ float colsum(float* v, int lane) {
  int n = V;
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) {
    if (n > 1) {
      n >>= 1;
      const bool up = (lane & m) != 0;
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        if (i < n) {
          const float send = up ? v[i] : v[i + n];
          const float keep = up ? v[i + n] : v[i];
          v[i] = keep + __shfl_xor_sync(0xffffffffu, send, m);
        }
      }
    } else {
      v[0] += __shfl_xor_sync(0xffffffffu, v[0], m);
    }
  }
  return v[0];
}

// pass order inside a tile: forward block 1, forward block 2, backward block 2, backward block 1
__device__ __forceinline__ int pass_block(int pass) { return (pass == 1 || pass == 2) ? 1 : 0; }

template <int PASSES, bool PROF>
__global__ void __launch_bounds__(WS_THREADS, 1) mlp_grad_ws_kernel(WsGradArgs ta) {
  long long pc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  long long plast = PROF ? clock64() : 0;
  long long* tbuf = nullptr;   // PROF: event trace of one tile (the 41st of CTA (0, 0)), 512 slots per role after the counters
  int tn = 0;
  constexpr int TRACE_TILE = 40;
#define TRACE_BEGIN(role, designated)                                                                          \
  if (PROF) {                                                                                                  \
    tbuf = (blockIdx.x == 0 && blockIdx.y == 0 && (designated) && tile == TRACE_TILE * (int)gridDim.x)        \
               ? ta.prof + 32 + (role) * 512 : nullptr;                                                        \
    if (tbuf) { tn = 1; tbuf[0] = clock64(); }                                                                 \
  }
#define TRACE_END() if (PROF && tbuf) { tbuf[511] = tn; tbuf = nullptr; }
  const GradArgs& a = ta.g;
  unsigned char* smem = ws_smem;
  unsigned char* sX = smem + OFF_SX;
  unsigned char* sGU = smem + OFF_SGU;
  unsigned char* sH = smem + OFF_SH;
  unsigned char* sGZ = smem + OFF_SGZ;
  float* sBias = reinterpret_cast<float*>(smem + OFF_BIAS);
  float* sRed = reinterpret_cast<float*>(smem + OFF_RED);
  float* sPar = reinterpret_cast<float*>(smem + OFF_PAR);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + OFF_BAR + B_COUNT * 8);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int net = blockIdx.y;
  const float* __restrict__ p = a.params + (net ? NAVPPO_CRITIC_OFFSET : 0);
  const unsigned char* __restrict__ wblob_g = ta.wprep + (size_t)net * WS_NET_BLOB;
  float* __restrict__ grow = a.gpart + ((size_t)net * gridDim.x + blockIdx.x) * NET_ROW;
  const int ntiles = (a.T + 127) / 128;

  if (tid == 0) {
    for (int i = 0; i < 2; ++i) { tc::mbar_init(bars + B_ZFULL + i, 1); tc::mbar_init(bars + B_EFULL + i, 8); tc::mbar_init(bars + B_HFREE + i, 1); }
    for (int i = 0; i < NWSLOT; ++i) { tc::mbar_init(bars + B_WFULL + i, 1); tc::mbar_init(bars + B_WFREE + i, 1); }
    tc::mbar_init(bars + B_DWFULL, 1);
    tc::mbar_init(bars + B_DWFREE, 4);
    tc::mbar_init(bars + B_ACC, 1);
    tc::mbar_init(bars + B_XREADY, 4);
  }
  if (warp == 0) tc::tmem_alloc(tmem_slot, TM_COLS);
  for (int i = tid; i < NET_ROW; i += WS_THREADS) grow[i] = 0.f;
  __threadfence();   // the zeros are in L2 before any red.add of this CTA
  for (int i = tid; i < 2 * HID; i += WS_THREADS) sBias[i] = p[(i < HID ? O_B1A : O_B2A - HID) + i];
  if (tid < OBS) sPar[P_B1B + tid] = p[O_B1B + tid];
  else if (tid < OBS + X1) sPar[tid] = p[O_B2B + tid - OBS];
  else if (tid < OBS + X1 + (net == 0 ? ACTOR_HEAD : CRITIC_HEAD)) sPar[tid] = p[O_HEAD + tid - OBS - X1];
  if (tid < 128) {   // constant part of the X tile: columns 32..47 = (1, 0, 0, ...) in every row
    float ones[8] = {1.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, zeros[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    store8<3>(sX, SX_PART, tid, 32, ones);
    store8<3>(sX, SX_PART, tid, 40, zeros);
  }
  tc::fence_smem_to_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = *tmem_slot;

  if (warp < 8) {
    // ====================================================================== epilogue warps
    reg_inc<184>();
    const int q = warp & 3, ch = warp >> 2;            // TMEM lane quarter / 32-column half of a half-chunk
    const int row = q * 32 + lane;                     // sample row
    const bool owner = ch == 0;                        // threads 0..127 own one sample row each
    const uint32_t trow = tmem + ((uint32_t)(q * 32) << 16);
    uint32_t pz = 0, ph = 0x3, pacc = 0;               // parity bits per ring slot (hfree starts "free")
    double macc[4] = {0.0, 0.0, 0.0, 0.0};

    // one arrive per warp once every lane's TMEM accesses (and, with SMEM, tile stores) are ordered
    auto publish = [&](uint64_t* bar, bool wrote_smem) {
      if (wrote_smem) tc::fence_smem_to_async();
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar);
    };
    // 32 rows x 4 packed words (8 bf16 columns) -> one 16-byte granule per row of a 128-row tile
    auto st_granule = [&](unsigned char* tile, int col0, const uint32_t* w) {
      *reinterpret_cast<uint4*>(tile + (uint32_t)(col0 >> 3) * ROWG + (uint32_t)row * 16) = make_uint4(w[0], w[1], w[2], w[3]);
    };

    float x1[X1];
    auto load_x0 = [&](int tile) {                     // this row's observation (zeros past the end of the batch)
      const int sj = tile * 128 + row;
      if (tile < ntiles && sj < a.T) {
        const float4* o = reinterpret_cast<const float4*>(a.obs + (size_t)sj * OBS);
#pragma unroll
        for (int k4 = 0; k4 < OBS / 4; ++k4) {
          const float4 t = o[k4];
          x1[4 * k4] = t.x; x1[4 * k4 + 1] = t.y; x1[4 * k4 + 2] = t.z; x1[4 * k4 + 3] = t.w;
        }
      } else {
#pragma unroll
        for (int k = 0; k < OBS; ++k) x1[k] = 0.f;
      }
    };
    if (owner) load_x0(blockIdx.x);

    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      TRACE_END();
      TRACE_BEGIN(warp >> 2, lane == 0 && (warp == 0 || warp == 4));
      const int si = tile * 128 + row;
      const bool valid = owner && si < a.T;
      float u2[X1], gsk[OBS];
      uint32_t u1pos = 0;                              // bit k: u1[k] > 0
      float2 av = make_float2(0.f, 0.f);
      float s_lpo = 0.f, s_adv = 0.f, s_rtg = 0.f;
      // ------------------------------------------------------------------ publish x0; fetch the row's scalars early
      if (owner) {
        store8<PASSES>(sX, SX_PART, row, 0, x1);
        store8<PASSES>(sX, SX_PART, row, 8, x1 + 8);
        publish(bars + B_XREADY, true);
        if (valid) {
          if (net == 0) {
            av = reinterpret_cast<const float2*>(a.act)[si];
            s_lpo = a.logp_old[si];
            s_adv = a.adv[si];
          } else {
            s_rtg = a.rtg[si];
          }
        }
      }

#pragma unroll 1
      for (int pass = 0; pass < 4; ++pass) {
        const int blk = pass_block(pass);
        const bool fwd = pass < 2;
        // -------------------------------------------------------------- the 8 half-chunks of this pass
#pragma unroll 1
        for (int h = 0; h < NHC; ++h) {
          const int s = h & 1;
          PT(5);
          tc::mbar_wait(bars + B_ZFULL + s, (pz >> s) & 1); pz ^= 1u << s;    // Z (and GH) of half-chunk h are in TMEM
          tc::fence_after_sync();
          PT(0);
          const float* sBa = sBias + blk * HID + h * HC + ch * 32;
          const uint32_t tz = trow + TM_ZG + s * 128 + ch * 32;
          const int tcol = s * 64 + ch * 32;                                   // column inside the 128-column tiles
          if (fwd) {
            // H = lrelu(Z + ba), written back over the columns this warp just read as the A operand of
            // U += H Wb^T: [hi: 16 packed columns | lo: 16 packed columns]
            float v[32];
            tc::tmem_ld16_nowait(tz, v);
            tc::tmem_ld16_nowait(tz + 16, v + 16);
            tc::tmem_ld_wait();
            uint32_t hi[16], lo[16];
#pragma unroll
            for (int i4 = 0; i4 < 8; ++i4) {
              const float4 b4 = *reinterpret_cast<const float4*>(sBa + 4 * i4);
              const float h0 = lrelu_fast(v[4 * i4] + b4.x), h1 = lrelu_fast(v[4 * i4 + 1] + b4.y);
              const float h2 = lrelu_fast(v[4 * i4 + 2] + b4.z), h3 = lrelu_fast(v[4 * i4 + 3] + b4.w);
              tc::split_bf16x2(h0, h1, &hi[2 * i4], &lo[2 * i4]);
              tc::split_bf16x2(h2, h3, &hi[2 * i4 + 1], &lo[2 * i4 + 1]);
            }
            tc::tmem_st16(tz, hi);
            if (PASSES == 3) tc::tmem_st16(tz + 16, lo);
            tc::tmem_st_wait();
            publish(bars + B_EFULL + s, false);
          } else {
            // H = lrelu(z), GZ = GH * lrelu'(z), z = Z + ba
            float z[32], g[32];
            tc::tmem_ld16_nowait(tz, z);
            tc::tmem_ld16_nowait(tz + 16, z + 16);
            tc::tmem_ld16_nowait(tz + 64, g);
            tc::tmem_ld16_nowait(tz + 80, g + 16);
            tc::tmem_ld_wait();
            uint32_t hh[16], hl[16], gh[16], gl[16];
#pragma unroll
            for (int i4 = 0; i4 < 8; ++i4) {
              const float4 b4 = *reinterpret_cast<const float4*>(sBa + 4 * i4);   // one broadcast load per 4 biases
              const float bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const int i = 4 * i4 + k;
                const float zz = z[i] + bb[k];
                const float sl = zz > 0.f ? 1.f : LEAK;       // LeakyReLU slope: h = zz * slope, g_z = g_h * slope
                z[i] = zz * sl;
                g[i] = g[i] * sl;
              }
              tc::split_bf16x2(z[4 * i4], z[4 * i4 + 1], &hh[2 * i4], &hl[2 * i4]);
              tc::split_bf16x2(z[4 * i4 + 2], z[4 * i4 + 3], &hh[2 * i4 + 1], &hl[2 * i4 + 1]);
              tc::split_bf16x2(g[4 * i4], g[4 * i4 + 1], &gh[2 * i4], &gl[2 * i4]);
              tc::split_bf16x2(g[4 * i4 + 2], g[4 * i4 + 3], &gh[2 * i4 + 1], &gl[2 * i4 + 1]);
            }
            if (blk) {   // GZ is also the A operand (in tensor memory, over the GH columns) of GX += GZ Wa
              tc::tmem_st16(tz + 64, gh);
              if (PASSES == 3) tc::tmem_st16(tz + 80, gl);
            }
            PT(3);
            tc::mbar_wait(bars + B_HFREE + s, (ph >> s) & 1); ph ^= 1u << s;    // the tile columns are no longer read
            PT(1);
#pragma unroll
            for (int c8 = 0; c8 < 4; ++c8) {
              st_granule(sH, tcol + 8 * c8, hh + 4 * c8);
              st_granule(sGZ, tcol + 8 * c8, gh + 4 * c8);
              if (PASSES == 3) {
                st_granule(sH + SH_PART, tcol + 8 * c8, hl + 4 * c8);
                st_granule(sGZ + SH_PART, tcol + 8 * c8, gl + 4 * c8);
              }
            }
            if (blk) tc::tmem_st_wait();
            publish(bars + B_EFULL + s, true);
          }
          PT(fwd ? 2 : 3);
        }

        // -------------------------------------------------------------- between the passes: row owners only
        if (!owner) continue;
        if (pass == 3) load_x0(tile + gridDim.x);       // the next tile's observation row flies behind the wait
        tc::mbar_wait(bars + B_ACC, pacc); pacc ^= 1;   // every product of this pass has completed
        tc::fence_after_sync();
        PT(4);
        if (pass == 0) {
          // block-1 output: u1 = x0 + U + bb, y1 = lrelu(u1) -> X columns 16..31
          float acc[16];
          tc::tmem_ld16(trow + TM_U, acc);
#pragma unroll
          for (int k = 0; k < OBS; ++k) {
            const float u = x1[k] + acc[k] + sPar[P_B1B + k];
            if (u > 0.f) u1pos |= 1u << k;
            x1[OBS + k] = lrelu(u);
          }
          store8<PASSES>(sX, SX_PART, row, 16, x1 + 16);
          store8<PASSES>(sX, SX_PART, row, 24, x1 + 24);
          publish(bars + B_XREADY, true);
        } else if (pass == 1) {
          {
            float acc[32];
            tc::tmem_ld16_nowait(trow + TM_U, acc);
            tc::tmem_ld16_nowait(trow + TM_U + 16, acc + 16);
            tc::tmem_ld_wait();
#pragma unroll
            for (int k = 0; k < X1; ++k) u2[k] = x1[k] + acc[k] + sPar[P_B2B + k];
          }
          // ================================================================ heads, losses, dL/du2
          float go1 = 0.f, go2 = 0.f;
          float y2[X1], gu2[X1];
#pragma unroll
          for (int k = 0; k < X1; ++k) y2[k] = lrelu(u2[k]);
          const float* hw = sPar + P_HEAD;
          if (net == 0) {
            float o1 = hw[X1], o2 = hw[2 * X1 + 1];
#pragma unroll
            for (int k = 0; k < X1; ++k) { o1 = fmaf(hw[k], y2[k], o1); o2 = fmaf(hw[X1 + 1 + k], y2[k], o2); }
            const float m0 = sigmoidf_(o1), m1 = tanhf(o2);
            if (valid) {
              const float lp = gauss_logp(av.x, av.y, m0, m1, a.var);
              const float lr = lp - s_lpo;
              const float ratio = expf(lr);                                          // ppo.py:316
              const float A = s_adv;
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
            float v = hw[X1];
#pragma unroll
            for (int k = 0; k < X1; ++k) v = fmaf(hw[k], y2[k], v);
            if (valid) {
              const float d = v - s_rtg;
              macc[1] += (double)(d * d);                                            // ppo.py:343
              go1 = 2.f * d * a.inv_n;
            }
          }
#pragma unroll
          for (int k = 0; k < X1; ++k) {
            const float gy = (net == 0) ? fmaf(hw[k], go1, hw[X1 + 1 + k] * go2) : hw[k] * go1;
            gu2[k] = gy * dlrelu(u2[k]);
          }
#pragma unroll
          for (int k = 0; k < X1; k += 8) store8<PASSES>(sGU, SGU_PART, row, k, gu2 + k);
          publish(bars + B_XREADY, true);               // GU is published: the tensor pipe starts the backward pass
#pragma unroll
          for (int k = 0; k < OBS; ++k) gsk[k] = gu2[OBS + k];   // the skip-connection share of dL/dx1
          // per-warp sums over the 32 samples (lane k ends up with column k): head weights / biases, then
          // the fc2 bias of block 2
          const int nh = (net == 0) ? 2 : 1;
          for (int hd = 0; hd < nh; ++hd) {
            const float g = hd ? go2 : go1;
            float bsum = g;
            for (int o = 16; o > 0; o >>= 1) bsum += __shfl_xor_sync(0xffffffffu, bsum, o);
            float t[X1];
#pragma unroll
            for (int k = 0; k < X1; ++k) t[k] = g * y2[k];
            const float cs = colsum<X1>(t, lane);
            sRed[warp * 128 + hd * (X1 + 1) + lane] = cs;
            if (lane == 0) sRed[warp * 128 + hd * (X1 + 1) + X1] = bsum;
          }
          {
            const float cs = colsum<X1>(gu2, lane);
            sRed[warp * 128 + 72 + lane] = cs;
          }
          named_sync(1, 128);
          {
            const int nhead = (net == 0) ? ACTOR_HEAD : CRITIC_HEAD;
            if (tid < nhead) red_add1(grow + O_HEAD + tid, ((sRed[tid] + sRed[128 + tid]) + sRed[256 + tid]) + sRed[384 + tid]);
            if (tid >= 72 && tid < 72 + X1)
              red_add1(grow + O_B2B + tid - 72, ((sRed[tid] + sRed[128 + tid]) + sRed[256 + tid]) + sRed[384 + tid]);
          }
        } else if (pass == 2) {
          // dL/dx1 = g_u2 (skip connection) + GX; dL/du1 = dL/dy1 * lrelu'(u1); publish GU1 for block 1
          float acc[16], gu1[OBS];
          tc::tmem_ld16(trow + TM_GX + 16, acc);
#pragma unroll
          for (int k = 0; k < OBS; ++k) gu1[k] = (gsk[k] + acc[k]) * (((u1pos >> k) & 1u) ? 1.f : LEAK);
          store8<PASSES>(sGU, SGU_PART, row, 0, gu1);
          store8<PASSES>(sGU, SGU_PART, row, 8, gu1 + 8);
          publish(bars + B_XREADY, true);
          const float cs = colsum<OBS>(gu1, lane);      // lanes 2k and 2k + 1 hold column k
          if (!(lane & 1)) sRed[warp * 128 + 104 + (lane >> 1)] = cs;
          named_sync(1, 128);
          if (tid >= 104 && tid < 104 + OBS)
            red_add1(grow + O_B1B + tid - 104, ((sRed[tid] + sRed[128 + tid]) + sRed[256 + tid]) + sRed[384 + tid]);
        }
        // pass == 3: the tile's last product has completed; X / GU may be overwritten
      }
    }
    PT(5);
    if (PROF && blockIdx.x == 0 && blockIdx.y == 0 && lane == 0 && (warp == 0 || warp == 4))
      for (int i = 0; i < 8; ++i) ta.prof[(warp ? 8 : 0) + i] = pc[i];
    // per-CTA metric partials of the row owners (every product has completed: the H tile is free)
    if (owner) {
      double* red = reinterpret_cast<double*>(sH);
#pragma unroll
      for (int k = 0; k < 4; ++k) red[k * 128 + tid] = macc[k];
    }
  } else if (warp < W_MMA) {
    // ====================================================================== weight-gradient flush warps
    reg_dec<56>();
    const int q = warp - W_FLUSH0;
    const int jrow = q * 32 + lane;                    // hidden unit inside the chunk = TMEM lane
    const uint32_t trow = tmem + ((uint32_t)(q * 32) << 16);
    uint32_t pdw = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      TRACE_END();
      TRACE_BEGIN(3, lane == 0 && warp == W_FLUSH0);
#pragma unroll 1
      for (int blk = 1; blk >= 0; --blk) {
        const int IN = blk ? X1 : OBS;
        const int o_wa = blk ? O_W2A : O_W1A, o_ba = blk ? O_B2A : O_B1A, o_wb = blk ? O_W2B : O_W1B;
#pragma unroll 1
        for (int c = 0; c < NHC / 2; ++c) {
          PT(1);
          tc::mbar_wait(bars + B_DWFULL, pdw); pdw ^= 1;
          tc::fence_after_sync();
          PT(0);
          const int j = c * 128 + jrow;
          float* ga = grow + o_wa + j * IN;
          float* gb = grow + o_wb + j * IN;            // kernel layout: fc2 transposed
          for (int c0 = 0; c0 < IN; c0 += 16) {
            float v[16];
            tc::tmem_ld16(trow + TM_DWA + c0, v);
#pragma unroll
            for (int i = 0; i < 4; ++i) red_add4(ga + c0 + 4 * i, v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
          }
          {
            float v[16];
            tc::tmem_ld16(trow + TM_DWA + 32, v);      // column 32 = sum over samples of g_z = bias gradient
            red_add1(grow + o_ba + j, v[0]);
          }
          for (int c0 = 0; c0 < IN; c0 += 16) {
            float v[16];
            tc::tmem_ld16(trow + TM_DWB + c0, v);
#pragma unroll
            for (int i = 0; i < 4; ++i) red_add4(gb + c0 + 4 * i, v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
          }
          tc::fence_before_sync();
          __syncwarp();
          if (lane == 0) mbar_arrive(bars + B_DWFREE);
        }
      }
    }
    PT(1);
    if (PROF && blockIdx.x == 0 && blockIdx.y == 0 && lane == 0 && warp == W_FLUSH0)
      for (int i = 0; i < 8; ++i) ta.prof[24 + i] = pc[i];
  } else {
    reg_dec<88>();
    if (warp == W_TMA && lane == 0) {
      // ==================================================================== weight producer
      uint32_t pfree = 0x7;                           // all three slots start free
      int slot = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
#pragma unroll 1
        for (int pass = 0; pass < 4; ++pass) {
          const int blk = pass_block(pass);
          const uint32_t bytes = wblob(blk ? X1 : OBS);
#pragma unroll 1
          for (int h = 0; h < NHC; ++h) {
            tc::mbar_wait(bars + B_WFREE + slot, (pfree >> slot) & 1); pfree ^= 1u << slot;
            tma_load(smem + OFF_W + slot * WSLOT, wblob_g + wblob_off(blk, h), bytes, bars + B_WFULL + slot);
            slot = slot == NWSLOT - 1 ? 0 : slot + 1;
          }
        }
      }
    } else if (warp == W_MMA) {
      // ==================================================================== the MMA-issuing warp (converged; one elected lane issues)
      const uint32_t aX = tc::smem_u32(sX), aGU = tc::smem_u32(sGU), aH = tc::smem_u32(sH), aGZ = tc::smem_u32(sGZ),
                     aW = tc::smem_u32(smem + OFF_W);
      // activation tiles: K-major (rows = samples are the M index) and MN-major (rows = samples are K)
      constexpr uint32_t RG = ROWG >> 4;
      const Opnd Xk{aX >> 4, SX_PART >> 4, RG, 8, 2 * RG}, Xm{aX >> 4, SX_PART >> 4, 8, RG, 16};
      const Opnd GUk{aGU >> 4, SGU_PART >> 4, RG, 8, 2 * RG}, GUm{aGU >> 4, SGU_PART >> 4, 8, RG, 16};
      const Opnd Hm{aH >> 4, SH_PART >> 4, 8, RG, 16}, GZm{aGZ >> 4, SH_PART >> 4, 8, RG, 16};
      uint32_t pe = 0, pw = 0, pdfree = 1, px = 0;
      int wslot_p1 = 0;       // ring slot of the next half-chunk whose Z / GH is to be issued
      int wslot_p2 = 0;       // ring slot of the next half-chunk whose U / GX is to be issued
      auto next_slot = [](int s) { return s == NWSLOT - 1 ? 0 : s + 1; };

      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        TRACE_END();
        TRACE_BEGIN(2, lane == 0);
#pragma unroll 1
        for (int pass = 0; pass < 4; ++pass) {
          const int blk = pass_block(pass);
          const bool fwd = pass < 2;
          const int IN = blk ? X1 : OBS;
          const uint32_t wp = wpart(IN);
          const uint32_t id_z = tc::make_idesc_bf16(128, HC, 0, 0), id_gh = tc::make_idesc_bf16(128, HC, 0, 1);
          const uint32_t id_u = tc::make_idesc_bf16(128, IN, 0, 0), id_gx = tc::make_idesc_bf16(128, IN, 0, 1);
          const uint32_t id_dwa = tc::make_idesc_bf16(128, XCOLS, 1, 1), id_dwb = tc::make_idesc_bf16(128, IN, 1, 1);

          // Z (and GH) of half-chunk h into ring slot h & 1
          auto issue_p1 = [&](int h) {
            const int s = h & 1;
            PT(4);
            tc::mbar_wait(bars + B_WFULL + wslot_p1, (pw >> wslot_p1) & 1); pw ^= 1u << wslot_p1;
            tc::fence_after_sync();
            PT(2);
            const uint32_t w = aW + wslot_p1 * WSLOT;
            // Z = X Wa^T : A = X (K-major), B = Wa (rows = hidden units, K-major), K = IN
            const Opnd Wk{w >> 4, wp >> 4, HC, 8, 2 * HC};
            if (blk) gemm<PASSES, 2>(tmem + TM_ZG + s * 128, Xk, Wk, id_z, false);
            else gemm<PASSES, 1>(tmem + TM_ZG + s * 128, Xk, Wk, id_z, false);
            if (!fwd) {
              // GH = GU Wb : A = GU (K-major), B = Wb read MN-major (rows = K = output features, R = IN)
              const Opnd Wbm{(w + 2 * wp) >> 4, wp >> 4, 8, (uint32_t)IN, 16};
              if (blk) gemm<PASSES, 2>(tmem + TM_ZG + s * 128 + 64, GUk, Wbm, id_gh, false);
              else gemm<PASSES, 1>(tmem + TM_ZG + s * 128 + 64, GUk, Wbm, id_gh, false);
            }
            PT(5);
            commit(bars + B_ZFULL + s);
            if (!fwd && !blk) commit(bars + B_WFREE + wslot_p1);   // backward block 1 has no GX: last reader
            wslot_p1 = next_slot(wslot_p1);
          };
          // U += H Wb^T (forward) or GX += GZ Wa (backward block 2) over the 64 hidden units of half-chunk h
          auto issue_p2 = [&](int h) {
            const int s = h & 1;
            const uint32_t w = aW + wslot_p2 * WSLOT;
            if (fwd) {   // A = H (tensor memory, over the Z columns), B = Wb (rows = output features, K-major)
              const Opnd Wbk{(w + 2 * wp) >> 4, wp >> 4, (uint32_t)IN, 8, 2u * IN};
              gemm_ts<PASSES>(tmem + TM_U, tmem + TM_ZG + s * 128, Wbk, id_u, h > 0);
            } else {     // A = GZ (tensor memory, over the GH columns), B = Wa read MN-major (rows = K = hidden units)
              const Opnd Wam{w >> 4, wp >> 4, 8, HC, 16};
              gemm_ts<PASSES>(tmem + TM_GX, tmem + TM_ZG + s * 128 + 64, Wam, id_gx, h > 0);
            }
            PT(6);
            commit(bars + B_WFREE + wslot_p2);
          };
          // weight gradients of the chunk whose two half-chunks sit complete in the H / GZ tiles
          auto issue_dw = [&]() {
            PT(4);
            tc::mbar_wait(bars + B_DWFREE, pdfree); pdfree ^= 1;   // the previous chunk's accumulators were flushed
            tc::fence_after_sync();
            PT(3);
            // dWa = GZ^T [X | 1], dWbT = H^T GU : both operands MN-major (rows = K = samples)
            gemm<PASSES, 8>(tmem + TM_DWA, GZm, Xm, id_dwa, false);
            gemm<PASSES, 8>(tmem + TM_DWB, Hm, GUm, id_dwb, false);
            PT(7);
            commit(bars + B_DWFULL);
            commit(bars + B_HFREE + 0);
            commit(bars + B_HFREE + 1);
          };

          PT(4);
          tc::mbar_wait(bars + B_XREADY, px); px ^= 1;    // the pass's input tile (X or GU) is published
          tc::fence_after_sync();
          PT(0);
          issue_p1(0);
          issue_p1(1);
#pragma unroll 1
          for (int h = 0; h < NHC; ++h) {
            const int s = h & 1;
            PT(4);
            tc::mbar_wait(bars + B_EFULL + s, (pe >> s) & 1); pe ^= 1u << s;   // operand tiles of half-chunk h are written
            tc::fence_after_sync();
            PT(1);
            if (fwd) {
              issue_p2(h);
              wslot_p2 = next_slot(wslot_p2);
              if (h + 2 < NHC) issue_p1(h + 2);
            } else {
              if (s == 0) {
                if (blk) issue_p2(h);
                wslot_p2 = next_slot(wslot_p2);
                if (h + 2 < NHC) issue_p1(h + 2);
              } else {
                if (blk) issue_p2(h);                    // before Z / GH of h + 2 overwrite the slot GZ sits in
                wslot_p2 = next_slot(wslot_p2);
                if (h + 2 < NHC) issue_p1(h + 2);
                issue_dw();
              }
            }
          }
          commit(bars + B_ACC);                   // everything issued so far: U / GX / the tile is complete
        }
      }
      PT(4);
      if (PROF && blockIdx.x == 0 && blockIdx.y == 0 && lane == 0)
        for (int i = 0; i < 8; ++i) ta.prof[16 + i] = pc[i];
    }
  }

  // per-CTA metric partials: fixed-order sum over the row owners
  tc::fence_before_sync();
  __syncthreads();
  if (tid < 4) {
    const double* red = reinterpret_cast<const double*>(sH);
    double t = 0.0;
    for (int m = 0; m < 128; ++m) t += red[tid * 128 + m];
    a.mpart[((size_t)net * gridDim.x + blockIdx.x) * 4 + tid] = t;
  }
  __syncthreads();
  if (warp == 0) {
    tc::fence_after_sync();
    tc::tmem_dealloc(tmem, TM_COLS);
  }
}

}  // namespace

// ---- internal entry points used by navppo_kernels.cu -----------------------------------
size_t navppo_tcws_prep_bytes() { return (size_t)2 * WS_NET_BLOB; }

static long long* g_prof = nullptr;   // navppo_tc_profile: per-role cycle counters + one-tile event trace of CTA (0, 0)

int navppo_tcws_init() {
  NAV_CUDA_TRY(cudaFuncSetAttribute(mlp_grad_ws_kernel<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WS_SMEM_BYTES));
  NAV_CUDA_TRY(cudaFuncSetAttribute(mlp_grad_ws_kernel<3, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WS_SMEM_BYTES));
  NAV_CUDA_TRY(cudaFuncSetAttribute(mlp_grad_ws_kernel<3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WS_SMEM_BYTES));
  return NAVSIM_OK;
}

// One gradient pass on the tensor cores: re-tile the current weights, then the fused kernel.
// `rows` CTAs per network; fills a.gpart / a.mpart like mlp_grad_kernel.
int navppo_tcws_grad_launch(const ppo::GradArgs& a, int rows, int passes, float* wprep, cudaStream_t s) {
  ws_prep_weights_kernel<<<dim3(24, 2), 256, 0, s>>>(a.params, reinterpret_cast<unsigned char*>(wprep));
  WsGradArgs ta{a, reinterpret_cast<const unsigned char*>(wprep), g_prof};
  if (passes == 3 && g_prof) mlp_grad_ws_kernel<3, true><<<dim3(rows, 2), WS_THREADS, WS_SMEM_BYTES, s>>>(ta);
  else if (passes == 3) mlp_grad_ws_kernel<3, false><<<dim3(rows, 2), WS_THREADS, WS_SMEM_BYTES, s>>>(ta);
  else mlp_grad_ws_kernel<1, false><<<dim3(rows, 2), WS_THREADS, WS_SMEM_BYTES, s>>>(ta);
  NAV_CUDA_TRY(cudaGetLastError());
  return NAVSIM_OK;
}

extern "C" int navppo_tc_profile(long long* device_counters) {
  g_prof = device_counters;
  return NAVSIM_OK;
}
