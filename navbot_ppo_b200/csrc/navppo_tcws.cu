// navppo_tcws.cu — warp-specialised tcgen05 gradient kernel of the PPO update (the default for
// NAVPPO_BF16 / NAVPPO_BF16X3; navppo_tc.cu keeps the first, single-role version as a bit-exact
// cross-check).
//
// Same arithmetic as mlp_grad_tc_kernel (one epoch body — forward, losses, backward — of one network
// per CTA row over tiles of 128 samples, ppo.py:305-397, net_actor.py:16-53), restructured so that the
// tensor pipe and the CUDA cores work at the same time and so that as few operands as possible
// travel through shared memory (operand fetch from shared memory, 128 B/clk, is what bounds these
// skinny products: profiles/r02a_tcgen05_issue_rate.txt):
//
//   warps 0-7   epilogue: TMEM -> bias / LeakyReLU / derivative -> bf16 hi|lo split -> written back IN
//               PLACE into tensor memory as the A operand of the next products; warps 0-3 also own one
//               sample row each (block outputs, heads, losses)
//   warps 8-11  weight-gradient flush: TMEM (lane = hidden unit) -> red.global.add into the CTA's
//               partial-gradient row
//   warp 12     issues the tcgen05.mma that FILL the TMEM ring (Z; Z^T / GH^T) — converged warp, one elected lane
//   warp 14     issues the tcgen05.mma that CONSUME what the epilogue packed (U; dWa / dWb, GX): two issuers
//               because one thread's waits, descriptor moves and commits were the longest chain of the tile
//   warp 13     one thread streams pre-tiled weights through one TMA ring of 7 x 16 KB slots
//
// Forward (per block, chunks of 128 hidden units, 2-slot TMEM ring of 128 columns):
//     Z = X Wa^T  [128 samples x 128]  (operands in shared memory)
//     H = lrelu(Z + ba) -> packed over Z;   U += H Wb^T  (A = H in tensor memory)
// Backward (per block, chunks of 128 hidden units x halves of 64 samples, TRANSPOSED: lane = hidden unit; 2-slot
// TMEM ring of 128 columns):
//     Z^T = Wa X^T, GH^T = Wb^T GU^T  [128 hidden x 64 samples]
//     H^T = lrelu(.), GZ^T = GH^T * lrelu'(.) -> packed over Z^T / GH^T
//     dWa += GZ^T [X | 1],  dWb^T += H^T GU   (A in tensor memory, K = the 64 samples; double-buffered
//                                              accumulators, flushed per chunk under the next chunk)
//     GX += GZ Wa  (block 2 only; A = the GZ^T tile the epilogue also leaves in shared memory)
// Only X, GU, the weights and (block 2) GZ^T ever sit in shared memory.  All hand-offs are mbarriers
// (tcgen05.commit on the tensor side, one elected arrive per warp on the CUDA-core side); there is no
// CTA-wide barrier inside the tile loop.
//
// Every product is issued in the same k-order and pass order as the first kernel, every reduction
// keeps its order, so the two kernels return bit-identical gradients (tests/test_ppo_tc_gpu.py).
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "../../include/navppo.h"
#include "nav_common.h"
#include "navsim_device.cuh"
#include "navsim_math.h"
#include "ppo_common.cuh"
#include "tc_common.cuh"

namespace {

using namespace ppo;

extern __shared__ __align__(128) unsigned char ws_smem[];

// ----------------------------------------------------------------------------------------
// Pre-split, pre-tiled weights: per network 4 chunk blobs (128 hidden units) of block 1 then 4 of block 2, each
//   [Wa hi | Wa lo | Wb hi | Wb lo]
// Wa tile: rows = hidden unit j (128), columns = input feature (IN);  Wb tile: rows = output feature o (IN), columns
// = hidden unit j (128) — the bf16 row-block format of tc_common.cuh.  The same blob serves the forward pass
// (Z = X Wa^T reads the Wa tile K-major, U += H Wb^T the Wb tile K-major) and, through the other descriptor strides,
// the transposed backward pass.
// ----------------------------------------------------------------------------------------
constexpr int CH = 128;
constexpr int NCH = HID / CH;                              // 4
__host__ __device__ constexpr uint32_t cpart(int IN) { return (uint32_t)(CH * IN * 2); }
__host__ __device__ constexpr uint32_t cblob(int IN) { return 4u * cpart(IN); }
__host__ __device__ constexpr uint32_t cblob_off(int block2, int c) {
  return block2 ? NCH * cblob(OBS) + (uint32_t)c * cblob(X1) : (uint32_t)c * cblob(OBS);
}
constexpr uint32_t WS_NET_BLOB = NCH * (cblob(OBS) + cblob(X1));   // 196,608 B

__global__ void ws_prep_weights_kernel(const float* __restrict__ params, unsigned char* __restrict__ wprep) {
  const int net = blockIdx.y;
  const float* p = params + (net ? NAVPPO_CRITIC_OFFSET : 0);
  unsigned char* out = wprep + (size_t)net * WS_NET_BLOB;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < HID * (OBS + X1); t += gridDim.x * blockDim.x) {
    const int block2 = t >= HID * OBS;
    const int IN = block2 ? X1 : OBS;
    const int r = block2 ? t - HID * OBS : t;
    const int j = r / IN, i = r % IN;              // fc1: hidden unit j, input feature i
    const int c = j / CH, jc = j % CH;
    unsigned char* cb = out + cblob_off(block2, c);
    const uint32_t cp = cpart(IN);
    uint16_t hi, lo;
    tc::split_bf16(p[(block2 ? O_W2A : O_W1A) + j * IN + i], &hi, &lo);
    *reinterpret_cast<uint16_t*>(cb + tc::rb16_off(CH, jc, i)) = hi;
    *reinterpret_cast<uint16_t*>(cb + cp + tc::rb16_off(CH, jc, i)) = lo;
    // fc2 [IN][512]: element (o = i, j)
    tc::split_bf16(p[(block2 ? O_W2B : O_W1B) + i * HID + j], &hi, &lo);
    *reinterpret_cast<uint16_t*>(cb + 2 * cp + tc::rb16_off(IN, i, jc)) = hi;
    *reinterpret_cast<uint16_t*>(cb + 3 * cp + tc::rb16_off(IN, i, jc)) = lo;
  }
}

// ----------------------------------------------------------------------------------------
// Shared-memory plan (bytes).  Activation tiles have 128 rows (samples), a hi part and a lo part.
// ----------------------------------------------------------------------------------------
constexpr int XCOLS = 48;                                   // x0 | y1 | 1 0 0 ... (bias-gradient column)
constexpr uint32_t ROWG = 128 * 16;                         // bytes of one 8-column group of a 128-row tile
constexpr uint32_t SX_PART = (XCOLS / 8) * ROWG;            // 12 KB
constexpr uint32_t SGU_PART = (X1 / 8) * ROWG;              // 8 KB
constexpr uint32_t SGZ_PART = (128 / 8) * ROWG;             // 32 KB: GZ^T of one chunk, 128 hidden rows x 128 samples
constexpr uint32_t OFF_SX = 0;
constexpr uint32_t OFF_SGU = OFF_SX + 2 * SX_PART;          // 24 KB
constexpr uint32_t OFF_SGZT = OFF_SGU + 2 * SGU_PART;       // 40 KB
// One weight ring for both directions: 7 slots of 16 KB.  A chunk blob takes two slots ([Wa hi | Wa lo] and
// [Wb hi | Wb lo]), so every pass consumes exactly 8 slots and the producer runs up to 7 slots (3 chunks) ahead of
// the products.
constexpr uint32_t WSLOT = 2 * cpart(X1);                   // 16 KB: half a chunk blob of block 2
constexpr int NWSLOT = 7;
constexpr uint32_t OFF_WF = OFF_SGZT + 2 * SGZ_PART;        // 104 KB: the ring
constexpr uint32_t OFF_BIAS = OFF_WF + NWSLOT * WSLOT;      // 216 KB: fc1 biases of both blocks, 2 x 512 floats
constexpr uint32_t OFF_RED = OFF_BIAS + 2 * HID * 4;        // head / bias-b gradient partials [4 warps][128] floats
constexpr uint32_t OFF_PAR = OFF_RED + 4 * 128 * 4;         // fc2 biases and head weights, 128 floats
constexpr int P_B1B = 0, P_B2B = OBS, P_HEAD = OBS + X1;    // offsets inside that block
constexpr uint32_t OFF_BAR = OFF_PAR + 128 * 4;             // mbarriers, tmem base
constexpr uint32_t WS_SMEM_BYTES = OFF_BAR + 384;   // (34 mbarriers + the TMEM base)           // 228,224 B of the 232,448 B a CTA may have

// mbarrier indices
enum { B_ZFULL = 0, B_EFULL = 4, B_SFREE = 8, B_WFULL = 12, B_WFREE = 19, B_DWFULL = 26, B_DWFREE = 28,
       B_GZFREE = 30, B_ACC = 31, B_XREADY = 32, B_GZFULL = 33, B_COUNT = 34 };

// TMEM columns: product ring of two 128-column slots (forward: Z of a chunk; backward: Z^T | GH^T of a unit), two
// weight-gradient accumulator buffers (dWa 48 | dWb 32 columns each), U and GX
constexpr uint32_t TM_ZG = 0;
constexpr uint32_t TM_DW = 256;      // buffer b at + 80 b: dWa at + 0, dWb at + 48
constexpr uint32_t TM_U = 416;       // 32 columns
constexpr uint32_t TM_GX = 448;      // 32 columns
constexpr uint32_t TM_COLS = 512;

constexpr int WS_THREADS = 512;
constexpr int W_FLUSH0 = 8, W_MMA = 12, W_TMA = 13, W_MMA2 = 14;

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

// D (+)= A B over KSTEPS instructions (16 k each); executed by a converged warp, issued by one lane.
// SWAP issues the two small terms of the split product in the other order: a transposed product
// (A = weights, B = activations) then adds (activation lo x weight hi) first, like the product it mirrors.
template <int PASSES, int KSTEPS, bool SWAP = false>
__device__ __forceinline__ void gemm(uint32_t tmem_d, const Opnd& A, const Opnd& B, uint32_t idesc, bool accumulate) {
  {   // the caller is the one elected lane of the issuing warp
    uint32_t acc = accumulate ? 1u : 0u;
    const uint32_t a0 = A.start | (A.lbo << 16), b0 = B.start | (B.lbo << 16);
#pragma unroll
    for (int kk = 0; kk < KSTEPS; ++kk) {
      const uint64_t ah = words_desc(a0 + kk * A.step, A.sbo), bh = words_desc(b0 + kk * B.step, B.sbo);
      if (PASSES == 3) {  // small terms first
        const uint64_t al = words_desc(a0 + kk * A.step + A.part, A.sbo), bl = words_desc(b0 + kk * B.step + B.part, B.sbo);
        if (SWAP) {
          tc::mma_bf16(tmem_d, ah, bl, idesc, acc); acc = 1u;
          tc::mma_bf16(tmem_d, al, bh, idesc, acc);
        } else {
          tc::mma_bf16(tmem_d, al, bh, idesc, acc); acc = 1u;
          tc::mma_bf16(tmem_d, ah, bl, idesc, acc);
        }
      }
      tc::mma_bf16(tmem_d, ah, bh, idesc, acc); acc = 1u;
    }
  }
}
// D (+)= A B over K = 64 with A in tensor memory, as the epilogue warps leave it: per 32 k [16 packed hi
// columns | 16 packed lo columns] (or, L16, per 16 k [8 packed hi | 8 packed lo]), 8 packed columns per instruction
template <int PASSES, bool L16 = false, int KSTEPS = 4>
__device__ __forceinline__ void gemm_ts(uint32_t tmem_d, uint32_t tmem_a, const Opnd& B, uint32_t idesc, bool accumulate) {
  {   // the caller is the one elected lane of the issuing warp
    uint32_t acc = accumulate ? 1u : 0u;
    const uint32_t b0 = B.start | (B.lbo << 16);
#pragma unroll
    for (int kk = 0; kk < KSTEPS; ++kk) {
      const uint32_t ah = L16 ? tmem_a + 16 * kk : tmem_a + 32 * (kk >> 1) + 8 * (kk & 1);
      const uint64_t bh = words_desc(b0 + kk * B.step, B.sbo);
      if (PASSES == 3) {  // small terms first
        tc::mma_bf16_ts(tmem_d, ah + (L16 ? 8 : 16), bh, idesc, acc); acc = 1u;
        tc::mma_bf16_ts(tmem_d, ah, words_desc(b0 + kk * B.step + B.part, B.sbo), idesc, acc);
      }
      tc::mma_bf16_ts(tmem_d, ah, bh, idesc, acc); acc = 1u;
    }
  }
}
// tcgen05.commit by the lane that issued the MMAs (the caller is that lane)
__device__ __forceinline__ void commit(uint64_t* bar) { tc::mma_commit(bar); }

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
  unsigned char* sGZT = smem + OFF_SGZT;
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
    for (int i = 0; i < 4; ++i) {
      tc::mbar_init(bars + B_ZFULL + i, 1); tc::mbar_init(bars + B_EFULL + i, 8); tc::mbar_init(bars + B_SFREE + i, 1);
    }
    for (int i = 0; i < 2; ++i) { tc::mbar_init(bars + B_DWFULL + i, 1); tc::mbar_init(bars + B_DWFREE + i, 4); }
    for (int i = 0; i < NWSLOT; ++i) { tc::mbar_init(bars + B_WFULL + i, 1); tc::mbar_init(bars + B_WFREE + i, 1); }
    tc::mbar_init(bars + B_GZFREE, 1);
    tc::mbar_init(bars + B_GZFULL, 16);     // 8 epilogue warps x the 2 sample halves of a chunk
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
    uint32_t pz = 0, pgz = 1, pacc = 0;                // parity bits (per ring slot); the GZ^T tile starts free
    double macc[4] = {0.0, 0.0, 0.0, 0.0};

    // one arrive per warp once every lane's TMEM accesses (and, with SMEM, tile stores) are ordered
    auto publish = [&](uint64_t* bar, bool wrote_smem) {
      if (wrote_smem) tc::fence_smem_to_async();
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar);
    };
    // 4 packed words (8 bf16 columns) of this thread's row -> one 16-byte granule of a 128-row tile
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
        // -------------------------------------------------------------- the steps of this pass: 4 chunks of 128 hidden
        // units (forward) / 8 units = chunk x sample half (backward), through the two 128-column ring slots
#pragma unroll 1
        for (int h = 0; h < (fwd ? NCH : 2 * NCH); ++h) {
          const int s = h & 1;
          PT(5);
          tc::mbar_wait(bars + B_ZFULL + s, (pz >> s) & 1); pz ^= 1u << s;    // Z (and GH) of step h are in TMEM
          tc::fence_after_sync();
          PT(0);
          if (fwd) {
            // chunk h of 128 hidden units; lane = sample row; this warp's 64 columns in two rounds of 32.
            // H = lrelu(Z + ba), written back over the columns just read as the A operand of U += H Wb^T:
            // per 32 columns [hi: 16 packed columns | lo: 16]
#pragma unroll
            for (int r2 = 0; r2 < 2; ++r2) {
              const uint32_t tz = trow + TM_ZG + s * 128 + ch * 64 + r2 * 32;
              const float* sBa = sBias + blk * HID + h * CH + ch * 64 + r2 * 32;
              float v[32];
              tc::tmem_ld16_nowait(tz, v);
              tc::tmem_ld16_nowait(tz + 16, v + 16);
              tc::tmem_ld_wait();
              uint32_t hi[16], lo[16];
#pragma unroll
              for (int i4 = 0; i4 < 8; ++i4)
                tc::bias_lrelu_split4(v + 4 * i4, *reinterpret_cast<const float4*>(sBa + 4 * i4), LEAK, hi + 2 * i4, lo + 2 * i4);
              tc::tmem_st16(tz, hi);
              if (PASSES == 3) tc::tmem_st16(tz + 16, lo);
            }
            tc::tmem_st_wait();
            publish(bars + B_EFULL + s, false);
          } else {
            const uint32_t tz = trow + TM_ZG + s * 128 + ch * 32;
            // unit h = (chunk c of 128 hidden units, half sh of 64 samples), TRANSPOSED: lane = hidden unit, columns
            // = samples.  H^T = lrelu(z), GZ^T = GH^T * lrelu'(z), z = Z^T + ba, written back over Z^T / GH^T as the
            // A operands of the weight-gradient products; block 2 also leaves GZ^T in shared memory for GX.
            const int c = h >> 1, sh = h & 1;
            const float ba = sBias[blk * HID + c * CH + row];
            float z[32], g[32];
            tc::tmem_ld16_nowait(tz, z);
            tc::tmem_ld16_nowait(tz + 16, z + 16);
            tc::tmem_ld16_nowait(tz + 64, g);
            tc::tmem_ld16_nowait(tz + 80, g + 16);
            tc::tmem_ld_wait();
            uint32_t hh[16], hl[16], gh[16], gl[16];
            const uint64_t ba2 = tc::pack2(ba, ba);
#pragma unroll
            for (int i2 = 0; i2 < 16; ++i2) {
              // LeakyReLU slope per value: h = zz * slope, g_z = g_h * slope (two values per instruction)
              const uint64_t zz = tc::add2(tc::pack2(z[2 * i2], z[2 * i2 + 1]), ba2);
              float za, zb;
              tc::unpack2(zz, za, zb);
              const uint64_t sl = tc::pack2(za > 0.f ? 1.f : LEAK, zb > 0.f ? 1.f : LEAK);
              float ha, hb, ga, gb;
              tc::unpack2(tc::mul2(zz, sl), ha, hb);
              tc::unpack2(tc::mul2(tc::pack2(g[2 * i2], g[2 * i2 + 1]), sl), ga, gb);
              tc::split_bf16x2(ha, hb, &hh[i2], &hl[i2]);
              tc::split_bf16x2(ga, gb, &gh[i2], &gl[i2]);
            }
            tc::tmem_st16(tz, hh);
            tc::tmem_st16(tz + 64, gh);
            if (PASSES == 3) {
              tc::tmem_st16(tz + 16, hl);
              tc::tmem_st16(tz + 80, gl);
            }
            PT(3);
            tc::tmem_st_wait();
            publish(bars + B_EFULL + s, false);                // the weight-gradient products may start
            if (blk) {
              // ... while GZ^T goes to shared memory for GX (issued once both halves of the chunk are there): the
              // wait for GX of the previous chunk no longer holds up the weight-gradient products of this unit
              if (sh == 0) {   // GX of the previous chunk has finished reading the tile
                tc::mbar_wait(bars + B_GZFREE, pgz); pgz ^= 1;
              }
              PT(1);
              const int scol = sh * 64 + ch * 32;              // sample column inside the [128 hidden x 128 samples] tile
#pragma unroll
              for (int c8 = 0; c8 < 4; ++c8) {
                st_granule(sGZT, scol + 8 * c8, gh + 4 * c8);
                if (PASSES == 3) st_granule(sGZT + SGZ_PART, scol + 8 * c8, gl + 4 * c8);
              }
              tc::fence_smem_to_async();
              __syncwarp();
              if (lane == 0) mbar_arrive(bars + B_GZFULL);
            }
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
    // per-CTA metric partials of the row owners (every product has completed: the GZ^T tile is free)
    if (owner) {
      double* red = reinterpret_cast<double*>(sGZT);
#pragma unroll
      for (int k = 0; k < 4; ++k) red[k * 128 + tid] = macc[k];
    }
  } else if (warp < W_MMA) {
    // ====================================================================== weight-gradient flush warps
    reg_dec<56>();
    const int q = warp - W_FLUSH0;
    const int jrow = q * 32 + lane;                    // hidden unit inside the chunk = TMEM lane
    const uint32_t trow = tmem + ((uint32_t)(q * 32) << 16);
    uint32_t pdw = 0;                                  // parity bit per accumulator buffer
    // fire-and-forget accumulation into the CTA's own partial-gradient row: the L2 does the fp32 add, so the
    // thread never waits for the old value (every address has ONE writer, in tile order -> the sum is the same
    // sequence of IEEE additions as a load-add-store; a load-add-store flush was tried and is slower: its five
    // dependent round trips per chunk take 5.5 k cycles against 2 k for the reds; pausing between the reds to leave
    // the load/store unit to the epilogue warps changes nothing, nor does a nanosleep back-off in this role's waits)
    auto accum16 = [&](float* dst, uint32_t taddr) {
      float v[16];
      tc::tmem_ld16(taddr, v);
#pragma unroll
      for (int i = 0; i < 4; ++i) red_add4(dst + 4 * i, v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
    };
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      TRACE_END();
      TRACE_BEGIN(3, lane == 0 && warp == W_FLUSH0);
#pragma unroll 1
      for (int blk = 1; blk >= 0; --blk) {
        const int IN = blk ? X1 : OBS;
        const int o_wa = blk ? O_W2A : O_W1A, o_ba = blk ? O_B2A : O_B1A, o_wb = blk ? O_W2B : O_W1B;
#pragma unroll 1
        for (int c = 0; c < NCH; ++c) {
          const int b = c & 1;
          PT(1);
          tc::mbar_wait(bars + B_DWFULL + b, (pdw >> b) & 1); pdw ^= 1u << b;
          tc::fence_after_sync();
          PT(0);
          const uint32_t td = trow + TM_DW + b * 80;
          const int j = c * CH + jrow;
          float* ga = grow + o_wa + j * IN;
          float* gb = grow + o_wb + j * IN;            // kernel layout: fc2 transposed
          for (int c0 = 0; c0 < IN; c0 += 16) accum16(ga + c0, td + c0);
          {
            float v[16];
            tc::tmem_ld16(td + 32, v);                 // column 32 = sum over samples of g_z = bias gradient
            red_add1(grow + o_ba + j, v[0]);
          }
          for (int c0 = 0; c0 < IN; c0 += 16) accum16(gb + c0, td + 48 + c0);
          tc::fence_before_sync();
          __syncwarp();
          if (lane == 0) mbar_arrive(bars + B_DWFREE + b);
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
      uint32_t pfree = (1u << NWSLOT) - 1;            // every ring slot starts free
      int slot = 0;
      auto load = [&](const unsigned char* src, uint32_t bytes) {
        tc::mbar_wait(bars + B_WFREE + slot, (pfree >> slot) & 1); pfree ^= 1u << slot;
        tma_load(smem + OFF_WF + slot * WSLOT, src, bytes, bars + B_WFULL + slot);
        slot = slot == NWSLOT - 1 ? 0 : slot + 1;
      };
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
#pragma unroll 1
        for (int pass = 0; pass < 4; ++pass) {
          const int blk = pass_block(pass);
          // a chunk blob (128 hidden units) as two slots, [Wa hi | Wa lo] then [Wb hi | Wb lo], forward and backward
          const uint32_t half = 2 * cpart(blk ? X1 : OBS);
#pragma unroll 1
          for (int c = 0; c < NCH; ++c) {
            load(wblob_g + cblob_off(blk, c), half);
            load(wblob_g + cblob_off(blk, c) + half, half);
          }
        }
      }
    } else if (warp == W_MMA) {
      // ==================================================================== MMA warp 1 (converged; one elected lane issues):
      // the products that FILL the TMEM ring — Z (forward), Z^T / GH^T (backward) — as soon as their weights
      // have landed and the slot's previous consumer product has completed
      const uint32_t aX = tc::smem_u32(sX), aGU = tc::smem_u32(sGU), aWF = tc::smem_u32(smem + OFF_WF);
      constexpr uint32_t RG = ROWG >> 4;
      const Opnd Xk{aX >> 4, SX_PART >> 4, RG, 8, 2 * RG};   // K-major: rows = samples are the M / N index
      uint32_t pw = 0, psf = 0, px = 0, pacc = 0;
      bool first_pass = true;
      int wslot = 0;          // weight ring slot of the next load in sequence
      auto next_slot = [](int s) { return s == NWSLOT - 1 ? 0 : s + 1; };
      auto wait_bar = [&](uint64_t* bar, uint32_t& parity_bits, int bit) {
        tc::mbar_wait(bar, (parity_bits >> bit) & 1);
        parity_bits ^= 1u << bit;
      };
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
#pragma unroll 1
        for (int pass = 0; pass < 4; ++pass) {
          const int blk = pass_block(pass);
          const int IN = blk ? X1 : OBS;
          if (!first_pass) { tc::mbar_wait(bars + B_ACC, pacc); pacc ^= 1; }   // the previous pass has left the TMEM ring
          first_pass = false;
          if (pass < 2) {
            tc::mbar_wait(bars + B_XREADY, px); px ^= 1;  // the pass's input tile X is published
            // ============================================================== forward: chunks of 128 hidden units
            const uint32_t cp = cpart(IN);
            const uint32_t id_z = tc::make_idesc_bf16(128, CH, 0, 0);
#pragma unroll 1
            for (int c = 0; c < NCH; ++c) {
              const int s = c & 1;
              const int sa = wslot, sb = next_slot(sa);                   // the chunk's [Wa] load; [Wb] is the consumer's
              wslot = next_slot(sb);
              if (c >= 2) wait_bar(bars + B_SFREE + s, psf, s);           // U of chunk c - 2 has read the slot
              wait_bar(bars + B_WFULL + sa, pw, sa);
              pw ^= 1u << sb;
              tc::fence_after_sync();
              if (elect_one()) {
                // Z of chunk c into ring slot c & 1: A = X (K-major), B = Wa (rows = the chunk's hidden units, K-major), K = IN
                const uint32_t w = aWF + sa * WSLOT;
                const Opnd Wk{w >> 4, cp >> 4, CH, 8, 2 * CH};
                if (blk) gemm<PASSES, 2>(tmem + TM_ZG + s * 128, Xk, Wk, id_z, false);
                else gemm<PASSES, 1>(tmem + TM_ZG + s * 128, Xk, Wk, id_z, false);
                commit(bars + B_ZFULL + s);
                commit(bars + B_WFREE + sa);                              // Z is the only reader of the [Wa] load
              }
              __syncwarp();
            }
          } else {
            // ============================================================== backward: chunks of 128 hidden units x halves
            // of 64 samples, transposed (D lanes = hidden units); the chunk's weights sit in ring slots sa
            // ([Wa hi | Wa lo]) and sb ([Wb hi | Wb lo])
            const uint32_t cp = cpart(IN);
            const uint32_t id_zt = tc::make_idesc_bf16(128, 64, 0, 0), id_ght = tc::make_idesc_bf16(128, 64, 1, 0);
#pragma unroll 1
            for (int u = 0; u < 2 * NCH; ++u) {
              const int sh = u & 1;
              const int sa = sh ? (wslot >= 2 ? wslot - 2 : wslot + NWSLOT - 2) : wslot;
              const int sb = next_slot(sa);
              if (sh == 0) {
                wait_bar(bars + B_WFULL + sa, pw, sa);
                wait_bar(bars + B_WFULL + sb, pw, sb);
                wslot = next_slot(sb);
              }
              if (u >= 2) wait_bar(bars + B_SFREE + sh, psf, sh);         // dW of unit u - 2 has read the slot
              tc::fence_after_sync();
              // Z^T = Wa X^T and GH^T = Wb^T GU^T of unit u = (chunk u / 2, sample half u % 2) into ring slot u & 1.
              // Z^T of the first two units only needs X, which has been there since the forward pass: it is issued
              // while the row owners still work on GU (the pass boundary), GH^T follows once GU is published
              const uint32_t wa = aWF + sa * WSLOT, wb = aWF + sb * WSLOT;
              const Opnd Wa_k{wa >> 4, cp >> 4, CH, 8, 2 * CH};                               // rows = hidden units = M
              const Opnd Wb_m{wb >> 4, cp >> 4, 8, (uint32_t)IN, 16};                         // rows = K = output features
              const Opnd Xs{(aX >> 4) + 64 * sh, SX_PART >> 4, RG, 8, 2 * RG};                // the 64 sample rows = N
              const Opnd GUs{(aGU >> 4) + 64 * sh, SGU_PART >> 4, RG, 8, 2 * RG};
              const uint32_t d = tmem + TM_ZG + sh * 128;
              if (u < 2) {
                if (elect_one()) {
                  if (blk) gemm<PASSES, 2, true>(d, Wa_k, Xs, id_zt, false);
                  else gemm<PASSES, 1, true>(d, Wa_k, Xs, id_zt, false);
                }
                __syncwarp();
                if (u == 0) continue;
                tc::mbar_wait(bars + B_XREADY, px); px ^= 1;              // GU is published
                tc::fence_after_sync();
                if (elect_one()) {
                  const Opnd GU0{aGU >> 4, SGU_PART >> 4, RG, 8, 2 * RG};
                  const uint32_t d0 = tmem + TM_ZG;
                  if (blk) gemm<PASSES, 2, true>(d0 + 64, Wb_m, GU0, id_ght, false);
                  else gemm<PASSES, 1, true>(d0 + 64, Wb_m, GU0, id_ght, false);
                  commit(bars + B_ZFULL + 0);
                  if (blk) gemm<PASSES, 2, true>(d + 64, Wb_m, GUs, id_ght, false);
                  else gemm<PASSES, 1, true>(d + 64, Wb_m, GUs, id_ght, false);
                  commit(bars + B_ZFULL + 1);
                  commit(bars + B_WFREE + sb);                        // last reader of Wb; block 1 has no GX: of Wa too
                  if (!blk) commit(bars + B_WFREE + sa);
                }
                __syncwarp();
                continue;
              }
              if (elect_one()) {
                if (blk) {
                  gemm<PASSES, 2, true>(d, Wa_k, Xs, id_zt, false);
                  gemm<PASSES, 2, true>(d + 64, Wb_m, GUs, id_ght, false);
                } else {
                  gemm<PASSES, 1, true>(d, Wa_k, Xs, id_zt, false);
                  gemm<PASSES, 1, true>(d + 64, Wb_m, GUs, id_ght, false);
                }
                commit(bars + B_ZFULL + sh);
                if (sh == 1) {                                      // last reader of Wb; block 1 has no GX: of Wa too
                  commit(bars + B_WFREE + sb);
                  if (!blk) commit(bars + B_WFREE + sa);
                }
              }
              __syncwarp();
            }
          }
        }
      }
    } else if (warp == W_MMA2) {
      // ==================================================================== MMA warp 2 (converged; one elected lane issues):
      // the products that CONSUME what the epilogue warps packed into the ring — U (forward), GX and the weight
      // gradients (backward)
      const uint32_t aX = tc::smem_u32(sX), aGU = tc::smem_u32(sGU), aGZT = tc::smem_u32(sGZT),
                     aWF = tc::smem_u32(smem + OFF_WF);
      constexpr uint32_t RG = ROWG >> 4;
      const Opnd GZTm{aGZT >> 4, SGZ_PART >> 4, 8, RG, 16};   // GZ^T tile read as A = GZ: rows = K = hidden units
      uint32_t pe = 0, pw = 0, pdfree = 0x3, px = 0, pgzf = 0;
      int wslot = 0;          // weight ring slot of the next load in sequence
      auto next_slot = [](int s) { return s == NWSLOT - 1 ? 0 : s + 1; };
      auto wait_bar = [&](uint64_t* bar, uint32_t& parity_bits, int bit) {
        tc::mbar_wait(bar, (parity_bits >> bit) & 1);
        parity_bits ^= 1u << bit;
      };
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        TRACE_END();
        TRACE_BEGIN(2, lane == 0);
#pragma unroll 1
        for (int pass = 0; pass < 4; ++pass) {
          const int blk = pass_block(pass);
          const int IN = blk ? X1 : OBS;
          PT(4);
          tc::mbar_wait(bars + B_XREADY, px); px ^= 1;    // the pass's input tile (X or GU) is published
          PT(0);
          if (pass < 2) {
            const uint32_t cp = cpart(IN);
            const uint32_t id_u = tc::make_idesc_bf16(128, IN, 0, 0);
#pragma unroll 1
            for (int c = 0; c < NCH; ++c) {
              const int s = c & 1;
              const int sa = wslot, sb = next_slot(sa);                  // the chunk's [Wb] load is the second one
              wslot = next_slot(sb);
              PT(4);
              wait_bar(bars + B_EFULL + s, pe, s);                       // H of chunk c sits packed in the slot
              PT(1);
              pw ^= 1u << sa;
              wait_bar(bars + B_WFULL + sb, pw, sb);
              tc::fence_after_sync();
              PT(2);
              if (elect_one()) {
                // U += H Wb^T over the chunk's 128 hidden units: A = H (tensor memory), B = Wb (rows = output features, K-major)
                const uint32_t w = aWF + sb * WSLOT;
                const Opnd Wbk{w >> 4, cp >> 4, (uint32_t)IN, 8, 2u * IN};
                gemm_ts<PASSES, false, 8>(tmem + TM_U, tmem + TM_ZG + s * 128, Wbk, id_u, c > 0);
                PT(6);
                commit(bars + B_WFREE + sb);
                if (c + 2 < NCH) commit(bars + B_SFREE + s);
                if (c == NCH - 1) commit(bars + B_ACC);                   // U is complete
              }
              __syncwarp();
            }
          } else {
            const uint32_t cp = cpart(IN);
            const uint32_t id_dwa = tc::make_idesc_bf16(128, XCOLS, 0, 1), id_dwb = tc::make_idesc_bf16(128, IN, 0, 1);
            const uint32_t id_gx = tc::make_idesc_bf16(128, IN, 1, 1);
#pragma unroll 1
            for (int u = 0; u < 2 * NCH; ++u) {
              const int c = u >> 1, sh = u & 1, b = c & 1;
              PT(4);
              wait_bar(bars + B_EFULL + sh, pe, sh);                     // H^T / GZ^T of unit u sit packed in the slot
              PT(1);
              int sa = 0;
              if (sh == 0) {
                wait_bar(bars + B_DWFREE + b, pdfree, b);                // accumulator buffer b was flushed
                PT(3);
              } else {                                                   // the chunk's two loads, in sequence
                sa = wslot;
                const int sb = next_slot(sa);
                if (blk) tc::mbar_wait(bars + B_WFULL + sa, (pw >> sa) & 1);   // (long complete: Z^T read the same load)
                pw ^= (1u << sa) | (1u << sb);
                wslot = next_slot(sb);
                PT(2);
              }
              tc::fence_after_sync();
              if (elect_one()) {
                // dWa += GZ^T [X | 1], dWbT += H^T GU over the unit's 64 samples: A in tensor memory, B MN-major
                const Opnd Xs{(aX >> 4) + 64 * sh, SX_PART >> 4, 8, RG, 16};
                const Opnd GUs{(aGU >> 4) + 64 * sh, SGU_PART >> 4, 8, RG, 16};
                const uint32_t slot = tmem + TM_ZG + sh * 128, d = tmem + TM_DW + b * 80;
                gemm_ts<PASSES>(d, slot + 64, Xs, id_dwa, sh > 0);
                gemm_ts<PASSES>(d + 48, slot, GUs, id_dwb, sh > 0);
                PT(7);
                if (u + 2 < 2 * NCH) commit(bars + B_SFREE + sh);
                if (sh == 1) commit(bars + B_DWFULL + b);
                if (!blk && u == 2 * NCH - 1) commit(bars + B_ACC);       // the tile is complete
              }
              __syncwarp();
              if (sh == 1 && blk) {
                // GX += GZ Wa over the chunk's 128 hidden units once both halves of GZ^T are in shared memory: both
                // operands MN-major (rows = K)
                tc::mbar_wait(bars + B_GZFULL, pgzf); pgzf ^= 1;
                tc::fence_after_sync();
                if (elect_one()) {
                  const uint32_t w = aWF + sa * WSLOT;
                  const Opnd Wa_m{w >> 4, cp >> 4, 8, CH, 16};
                  gemm<PASSES, 8>(tmem + TM_GX, GZTm, Wa_m, id_gx, c > 0);
                  PT(6);
                  commit(bars + B_GZFREE);
                  commit(bars + B_WFREE + sa);
                  if (u == 2 * NCH - 1) commit(bars + B_ACC);             // GX is complete
                }
                __syncwarp();
              }
            }
          }
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
    const double* red = reinterpret_cast<const double*>(sGZT);
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

// ========================================================================================
// Inference on the tensor cores: the forward half of the gradient kernel (same products, same epilogue
// arithmetic, same order — a sample's mean / value is bit-identical to what the update recomputes for it),
// one tile of 128 samples per CTA, + the get_action / evaluate epilogues of mlp_infer_kernel
// (ppo.py:690-705, 725-737).  Used by the fused rollout and by the update's first evaluate when the handle's
// precision is NAVPPO_BF16X3 / NAVPPO_BF16.
// A lone tile is a latency chain (wait, TMEM load, pack, TMEM store, arrive: ~0.45 us per step whatever its width),
// so the steps are as wide as tensor memory allows: chunks of 128 hidden units (the backward pass's weight blobs,
// which hold Wa with 128 rows and Wb with 128 k), 4 steps per block, three 128-column Z slots.  The k-order of
// every accumulation is unchanged.
//   warps 0 .. 4 EW - 1   epilogue: EW warps per TMEM lane quarter, 128 / EW columns of a chunk each;
//                         warps 0-3 also own one sample row each
//   warp 4 EW             issues every tcgen05.mma;   warp 4 EW + 1: one thread streams the weights (TMA ring)
// ========================================================================================
constexpr int INF_EW = 4;
constexpr int INF_CW = CH / INF_EW;                          // columns of a chunk per epilogue warp
constexpr int INF_ZSLOTS = 3;
constexpr int INF_EWARPS = 4 * INF_EW;
constexpr int INF_W_TMA = INF_EWARPS + 1;                    // the MMA warp is the one before it
constexpr int INF_THREADS = 32 * (INF_EWARPS + 2);
constexpr uint32_t IOFF_WF = 2 * SX_PART;                    // 24 KB: weight ring
constexpr uint32_t IOFF_BIAS = IOFF_WF + NWSLOT * WSLOT;     // 136 KB
constexpr uint32_t IOFF_PAR = IOFF_BIAS + 2 * HID * 4;
constexpr uint32_t IOFF_BAR = IOFF_PAR + 128 * 4;
constexpr uint32_t INF_SMEM_BYTES = IOFF_BAR + 256;
enum { I_ZFULL = 0, I_EFULL = 4, I_WFULL = 8, I_WFREE = 15, I_ACC = 22, I_XREADY = 23, I_COUNT = 24 };
static_assert(INF_ZSLOTS * CH <= (int)TM_U, "Z slots must end below the U accumulator");

struct WsInferArgs {
  static constexpr bool kFused = false;
  InferArgs g;
  const unsigned char* wprep;
  int mode;
  int weights_ready;  // chained launch: `wprep` was written before the predecessor started (not by it): stream it at once
  long long* prof;    // diagnostic: nanosecond timestamps of CTA (0, 0)'s first row owner (navppo_tc_profile), else null
};
// The fused rollout: the same kernel keeps its 128 robots for H steps — policy forward, sample, Env.step by the
// same 512 threads (4 lanes per robot, navsim_device.cuh), next observation straight from shared memory.  Robots are
// independent, so the CTAs never synchronise with each other and nothing but the rollout rows leaves the SM.
struct WsFusedArgs : WsInferArgs {
  static constexpr bool kFused = true;
  navsim_dev::SimConst c;
  navsim_dev::SimState st;
  const float* g_map;
  const uint16_t* rt_tab;
  navsim_dev::DevStats* stats;
  int H;
  float* obs_rows;     // [H, N, 16]: row t = the observation step t acts on (row 0 filled by the caller)
  float* next_obs;     // [N, 16]
  float* rew; uint8_t* done; uint8_t* arrive; uint8_t* trunc;   // [H, N]
  float* ep_ret; float* ep_path; int32_t* ep_len;               // [H, N] or null
};
constexpr uint32_t IOFF_SIM = IOFF_BAR + 256;                  // fused: [map mbarrier 16 B][actions 128 x 8 B][obs tile][map]
constexpr uint32_t FUSED_MAP_MAX = 8192;
constexpr uint32_t IOFF_SACT = IOFF_SIM + 16;
constexpr uint32_t IOFF_SOBS = IOFF_SACT + 128 * 8;
constexpr uint32_t IOFF_SMAP = IOFF_SOBS + 128 * navsim_dev::kObsPad * 4;
constexpr uint32_t FUSED_SMEM_BYTES = IOFF_SMAP + FUSED_MAP_MAX;
constexpr int FUSED_G = 4;                                     // lanes per robot: 128 robots x 4 = the 16 epilogue warps
__device__ __forceinline__ long long global_ns() {
  long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
#define ISTAMP(k) do { if (ta.prof && blockIdx.x == 0 && blockIdx.y == 0 && tid == 0) ta.prof[4096 + (k)] = global_ns(); } while (0)

template <int PASSES, class Args>
__global__ void __launch_bounds__(INF_THREADS, 1) mlp_infer_ws_kernel(Args ta) {
  constexpr bool FUSED = Args::kFused;
  const InferArgs& a = ta.g;
  int H = 1;
  if constexpr (FUSED) H = ta.H;
  const int net = blockIdx.y;
  if (ta.mode == INFER_FORWARD && ((net == 0 && !a.mu) || (net == 1 && !a.v))) return;
  unsigned char* smem = ws_smem;
  unsigned char* sX = smem + OFF_SX;
  float* sBias = reinterpret_cast<float*>(smem + IOFF_BIAS);
  float* sPar = reinterpret_cast<float*>(smem + IOFF_PAR);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + IOFF_BAR);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + IOFF_BAR + I_COUNT * 8);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const float* __restrict__ p = a.params + (net ? NAVPPO_CRITIC_OFFSET : 0);
  const unsigned char* __restrict__ wblob_g = ta.wprep + (size_t)net * WS_NET_BLOB;
  const int tile = blockIdx.x;
  ISTAMP(0);

  if (tid == 0) {
    for (int i = 0; i < INF_ZSLOTS; ++i) { tc::mbar_init(bars + I_ZFULL + i, 1); tc::mbar_init(bars + I_EFULL + i, INF_EWARPS); }
    for (int i = 0; i < NWSLOT; ++i) { tc::mbar_init(bars + I_WFULL + i, 1); tc::mbar_init(bars + I_WFREE + i, 1); }
    tc::mbar_init(bars + I_ACC, 1);
    tc::mbar_init(bars + I_XREADY, 4);
    if constexpr (FUSED) {   // the obstacle set: one bulk copy, waited on before the first Env.step
      uint64_t* map_bar = reinterpret_cast<uint64_t*>(smem + IOFF_SIM);
      tc::mbar_init(map_bar, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      tma_load(smem + IOFF_SMAP, reinterpret_cast<const unsigned char*>(ta.g_map), navsim_dev::map_bytes_of(ta.c.B, ta.c.S), map_bar);
    }
  }
  if (warp == 0) tc::tmem_alloc(tmem_slot, TM_COLS);
  for (int i = tid; i < 2 * HID; i += INF_THREADS) sBias[i] = p[(i < HID ? O_B1A : O_B2A - HID) + i];
  if (tid < OBS) sPar[P_B1B + tid] = p[O_B1B + tid];
  else if (tid < OBS + X1) sPar[tid] = p[O_B2B + tid - OBS];
  else if (tid < OBS + X1 + (net == 0 ? ACTOR_HEAD : CRITIC_HEAD)) sPar[tid] = p[O_HEAD + tid - OBS - X1];
  tc::fence_smem_to_async();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  ISTAMP(1);
  // chained launch (rollout): everything above — it reads only the parameters — ran under the previous kernel's
  // tail; the observations (and, first step, the re-tiled weights) are touched after this wait
  if (!(warp == INF_W_TMA && ta.weights_ready)) asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  ISTAMP(2);

  if (warp < INF_EWARPS) {
    // ====================================================================== epilogue warps
    const int q = warp & 3, ch = warp >> 2;            // TMEM lane quarter / column share of a half-chunk
    const int row = q * 32 + lane;                     // sample row
    const bool owner = ch == 0;
    const uint32_t trow = tmem + ((uint32_t)(q * 32) << 16);
    const int si = tile * 128 + row;
    const bool valid = owner && si < a.T;
    uint32_t pz = 0;
    auto publish = [&](uint64_t* bar, bool wrote_smem) {
      if (wrote_smem) tc::fence_smem_to_async();
      tc::fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar);
    };
    // fused rollout: this thread is also lane g of robot `slot` (4 lanes per robot)
    [[maybe_unused]] const int slot = tid >> 2, g = tid & (FUSED_G - 1);
    [[maybe_unused]] int ai = 0;
    [[maybe_unused]] bool sim_valid = false, sim_writer = false, goal_dirty = false;
    [[maybe_unused]] navsim_dev::Agent ag;
    [[maybe_unused]] navsim_dev::MapView mv;
    [[maybe_unused]] float* s_obs = reinterpret_cast<float*>(smem + IOFF_SOBS);
    [[maybe_unused]] float2* s_act = reinterpret_cast<float2*>(smem + IOFF_SACT);
    if constexpr (FUSED) {
      const int ai_raw = tile * 128 + slot;
      sim_valid = ai_raw < ta.c.N;
      ai = sim_valid ? ai_raw : ta.c.N - 1;            // surplus lanes shadow the last robot and store nothing
      sim_writer = sim_valid && g == 0;
      navsim_dev::load_agent(ta.st, ai, &ag);
      mv = navsim_dev::map_view(reinterpret_cast<const float*>(smem + IOFF_SMAP), ta.c.B, ta.c.S);
      tc::mbar_wait(reinterpret_cast<uint64_t*>(smem + IOFF_SIM), 0);
    }
    float x1[X1];
    if (owner) {
      if (si < a.T) {
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
    }
#pragma unroll 1
    for (int t = 0; t < H; ++t) {
    if (owner) {
      store8<PASSES>(sX, SX_PART, row, 0, x1);
      store8<PASSES>(sX, SX_PART, row, 8, x1 + 8);
      publish(bars + I_XREADY, true);
      ISTAMP(3);
    }
#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) {
      const int blk = pass;
#pragma unroll 1
      for (int c = 0; c < NCH; ++c) {
        const int s = c % INF_ZSLOTS;
        tc::mbar_wait(bars + I_ZFULL + s, (pz >> s) & 1); pz ^= 1u << s;
        tc::fence_after_sync();
        if (c == 0) ISTAMP(4 + 4 * pass);
        // H = lrelu(Z + ba) of this warp's INF_CW columns, packed over the columns it just read as the A operand
        // of U += H Wb^T: per 16 columns [hi: 8 packed words | lo: 8 packed words]
        const int c0 = ch * INF_CW;
        const uint32_t tslot = trow + TM_ZG + s * CH;
        const float* sBa = sBias + blk * HID + c * CH + c0;
        float v[INF_CW];
#pragma unroll
        for (int i = 0; i < INF_CW; i += 16) tc::tmem_ld16_nowait(tslot + c0 + i, v + i);
        tc::tmem_ld_wait();
        uint32_t hi[INF_CW / 2], lo[INF_CW / 2];
#pragma unroll
        for (int i4 = 0; i4 < INF_CW / 4; ++i4)
          tc::bias_lrelu_split4(v + 4 * i4, *reinterpret_cast<const float4*>(sBa + 4 * i4), LEAK, hi + 2 * i4, lo + 2 * i4);
#pragma unroll
        for (int j = 0; j < INF_CW / 16; ++j) {
          tc::tmem_st8(tslot + c0 + 16 * j, hi + 8 * j);
          if (PASSES == 3) tc::tmem_st8(tslot + c0 + 16 * j + 8, lo + 8 * j);
        }
        tc::tmem_st_wait();
        publish(bars + I_EFULL + s, false);
      }
      if (!owner) continue;
      ISTAMP(5 + 4 * pass);
      tc::mbar_wait(bars + I_ACC, pass);              // every product of this pass has completed
      tc::fence_after_sync();
      ISTAMP(6 + 4 * pass);
      if (pass == 0) {
        // block-1 output: u1 = x0 + U + bb, y1 = lrelu(u1) -> X columns 16..31
        float acc[16];
        tc::tmem_ld16(trow + TM_U, acc);
#pragma unroll
        for (int k = 0; k < OBS; ++k) x1[OBS + k] = lrelu(x1[k] + acc[k] + sPar[P_B1B + k]);
        store8<PASSES>(sX, SX_PART, row, 16, x1 + 16);
        store8<PASSES>(sX, SX_PART, row, 24, x1 + 24);
        publish(bars + I_XREADY, true);
        ISTAMP(7);
      } else {
        float y2[X1];
        {
          float acc[32];
          tc::tmem_ld16_nowait(trow + TM_U, acc);
          tc::tmem_ld16_nowait(trow + TM_U + 16, acc + 16);
          tc::tmem_ld_wait();
#pragma unroll
          for (int k = 0; k < X1; ++k) y2[k] = lrelu(x1[k] + acc[k] + sPar[P_B2B + k]);
        }
        const float* hw = sPar + P_HEAD;
        if (net == 1) {                        // critic: V = out(X), net_critic.py:129
          float v = hw[X1];
#pragma unroll
          for (int k = 0; k < X1; ++k) v = fmaf(hw[k], y2[k], v);
          if (valid) a.v[si] = v;
        } else {
          float o1 = hw[X1], o2 = hw[2 * X1 + 1];
#pragma unroll
          for (int k = 0; k < X1; ++k) { o1 = fmaf(hw[k], y2[k], o1); o2 = fmaf(hw[X1 + 1 + k], y2[k], o2); }
          const float m0 = sigmoidf_(o1), m1 = tanhf(o2);  // net_actor.py:141-142
          if constexpr (FUSED) if (!valid) s_act[row] = make_float2(0.f, 0.f);
          if (valid) {
            if (ta.mode == INFER_FORWARD) {
              reinterpret_cast<float2*>(a.mu)[si] = make_float2(m0, m1);
            } else if (ta.mode == INFER_ACT) {
              const float var_ = a.dyn ? __uint_as_float(a.dyn[0]) : a.var;
              const uint32_t draw_ = (a.dyn ? a.draw + a.dyn[1] : a.draw) + (uint32_t)t;
              float e0, e1;
              if (a.noise_in) {
                const float2 e = reinterpret_cast<const float2*>(a.noise_in)[si];
                e0 = e.x; e1 = e.y;
              } else {  // Box-Muller on two Philox words
                uint32_t r[4];
                const uint64_t agent = (uint64_t)(a.agent_off + si);
                nv_philox4x32_10(draw_, 2u, (uint32_t)agent, (uint32_t)(agent >> 32), (uint32_t)a.seed,
                                 (uint32_t)(a.seed >> 32), r);
                const float uu = ((float)(r[0] >> 8) + 0.5f) * (1.0f / 16777216.0f);  // (0, 1)
                const float vv = ((float)(r[1] >> 8) + 0.5f) * (1.0f / 16777216.0f);
                const float rad = sqrtf(-2.0f * logf(uu));
                float sn, cs;
                sincosf(6.283185307179586f * vv, &sn, &cs);
                e0 = rad * cs; e1 = rad * sn;
              }
              const float sd = sqrtf(var_);
              float a0 = fmaf(sd, e0, m0), a1 = fmaf(sd, e1, m1);       // dist.sample(), ppo.py:698-699
              a0 = fminf(fmaxf(a0, 0.f), 1.f);                          // ppo.py:701
              a1 = fminf(fmaxf(a1, -1.f), 1.f);                         // ppo.py:702
              reinterpret_cast<float2*>(a.act)[(size_t)t * a.T + si] = make_float2(a0, a1);
              a.logp[(size_t)t * a.T + si] = gauss_logp(a0, a1, m0, m1, var_);   // ppo.py:704 (at the clamped action)
              if (a.mu) reinterpret_cast<float2*>(a.mu)[si] = make_float2(m0, m1);
              if constexpr (FUSED) s_act[row] = make_float2(a0, a1);
            } else {
              const float2 av = reinterpret_cast<const float2*>(a.act_in)[si];
              a.logp[si] = gauss_logp(av.x, av.y, m0, m1, a.var);       // ppo.py:734-735
            }
          }
        }
      }
    }
    if constexpr (FUSED) {
      // ------------------------------------------------------------------ Env.step of this CTA's 128 robots
      named_sync(2, 32 * INF_EWARPS);                 // every row's action is in shared memory
      const float2 av = s_act[slot];
      navsim_dev::StepRows rows;
      {
        const size_t vo = (size_t)t * ta.c.N;
        rows.rew = ta.rew + vo; rows.done = ta.done + vo; rows.arrive = ta.arrive + vo;
        rows.trunc = ta.trunc ? ta.trunc + vo : nullptr;
        rows.ep_ret = ta.ep_ret ? ta.ep_ret + vo : nullptr;
        rows.ep_path = ta.ep_path ? ta.ep_path + vo : nullptr;
        rows.ep_len = ta.ep_len ? ta.ep_len + vo : nullptr;
      }
      navsim_dev::agent_step<FUSED_G, NAVSIM_LIDAR_FEATS, false>(
          ta.c, mv, ta.rt_tab, ta.stats, ag, av.x, av.y, g, ai, (uint64_t)(ta.c.agent_off + ai), sim_valid, sim_writer,
          s_obs + slot * navsim_dev::kObsPad, nullptr, nullptr, rows, nullptr, goal_dirty);
      named_sync(2, 32 * INF_EWARPS);                 // every robot's next observation row is in shared memory
      {
        // each warp writes out the rows of its own 8 robots as 128-bit stores; the row owners keep theirs for the
        // next forward pass
        float* dst = (t + 1 < H) ? ta.obs_rows + (size_t)(t + 1) * ta.c.N * OBS : ta.next_obs;
        const int r = warp * 8 + (lane >> 2), qq = (lane & 3) * 4;
        if (tile * 128 + r < ta.c.N) {
          const float* src = s_obs + r * navsim_dev::kObsPad + qq;
          reinterpret_cast<float4*>(dst + (size_t)(tile * 128 + r) * OBS)[lane & 3] = make_float4(src[0], src[1], src[2], src[3]);
        }
        if (owner) {
#pragma unroll
          for (int k = 0; k < OBS; ++k) x1[k] = s_obs[row * navsim_dev::kObsPad + k];
        }
      }
    }
    }   // step t
    if constexpr (FUSED) {
      if (sim_writer) navsim_dev::store_agent(ta.st, ai, ag, goal_dirty);
      if (blockIdx.x == 0 && tid == 0) atomicAdd(&ta.stats->steps, (unsigned long long)ta.c.N * (unsigned long long)H);
    }
  } else if (warp == INF_W_TMA) {
    if (lane == 0) {
      int slot = 0;
      uint32_t pfree = (1u << NWSLOT) - 1;
#pragma unroll 1
      for (int tt = 0; tt < H; ++tt)
#pragma unroll 1
      for (int i = 0; i < 4 * NCH; ++i) {       // per block and chunk: [Wa hi | Wa lo], then [Wb hi | Wb lo]
        const int blk = i >> 3, c = (i >> 1) & 3, part = i & 1;
        const uint32_t half = 2 * cpart(blk ? X1 : OBS);
        tc::mbar_wait(bars + I_WFREE + slot, (pfree >> slot) & 1); pfree ^= 1u << slot;
        tma_load(smem + IOFF_WF + slot * WSLOT, wblob_g + cblob_off(blk, c) + part * half, half, bars + I_WFULL + slot);
        slot = slot == NWSLOT - 1 ? 0 : slot + 1;
      }
    }
  } else {
    // ====================================================================== the MMA-issuing warp
    const uint32_t aX = tc::smem_u32(sX), aWF = tc::smem_u32(smem + IOFF_WF);
    constexpr uint32_t RG = ROWG >> 4;
    const Opnd Xk{aX >> 4, SX_PART >> 4, RG, 8, 2 * RG};
    uint32_t pe = 0, pw = 0;
    int wz = 0, wu = 1;     // weight ring slot of the next Z (a chunk's [Wa] load) / of the next U (its [Wb] load)
    auto next2 = [](int s) { return s + 2 >= NWSLOT ? s + 2 - NWSLOT : s + 2; };
    auto wait_bar = [&](uint64_t* bar, uint32_t& parity_bits, int bit) {
      tc::mbar_wait(bar, (parity_bits >> bit) & 1);
      parity_bits ^= 1u << bit;
    };
#pragma unroll 1
    for (int tt = 0; tt < H; ++tt)
#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) {
      const int blk = pass;
      const int IN = blk ? X1 : OBS;
      const uint32_t cp = cpart(IN);
      const uint32_t id_z = tc::make_idesc_bf16(128, CH, 0, 0), id_u = tc::make_idesc_bf16(128, IN, 0, 0);
      // Z of chunk c into slot c % 3: A = X (K-major), B = Wa (rows = the chunk's 128 hidden units, K-major), K = IN
      auto issue_z = [&](int c, int wslot) {
        const int s = c % INF_ZSLOTS;
        const uint32_t w = aWF + wslot * WSLOT;
        const Opnd Wk{w >> 4, cp >> 4, CH, 8, 2 * CH};
        if (blk) gemm<PASSES, 2>(tmem + TM_ZG + s * CH, Xk, Wk, id_z, false);
        else gemm<PASSES, 1>(tmem + TM_ZG + s * CH, Xk, Wk, id_z, false);
        commit(bars + I_ZFULL + s);
        commit(bars + I_WFREE + wslot);          // Z is the only reader of the chunk's [Wa] load
      };
      tc::mbar_wait(bars + I_XREADY, pass);
      {
        int w = wz;
        for (int i = 0; i < INF_ZSLOTS; ++i) { wait_bar(bars + I_WFULL + w, pw, w); w = next2(w); }
      }
      tc::fence_after_sync();
      if (elect_one()) {
        int w = wz;
        for (int i = 0; i < INF_ZSLOTS; ++i) { issue_z(i, w); w = next2(w); }
      }
      __syncwarp();
      for (int i = 0; i < INF_ZSLOTS; ++i) wz = next2(wz);
#pragma unroll 1
      for (int c = 0; c < NCH; ++c) {
        const int s = c % INF_ZSLOTS;
        wait_bar(bars + I_EFULL + s, pe, s);
        wait_bar(bars + I_WFULL + wu, pw, wu);
        if (c + INF_ZSLOTS < NCH) wait_bar(bars + I_WFULL + wz, pw, wz);
        tc::fence_after_sync();
        if (elect_one()) {
          // U += H Wb^T over the chunk's 128 hidden units: A = H (tensor memory), B = Wb (rows = output features, K-major)
          const uint32_t w = aWF + wu * WSLOT;
          const Opnd Wbk{w >> 4, cp >> 4, (uint32_t)IN, 8, 2u * IN};
          gemm_ts<PASSES, true, 8>(tmem + TM_U, tmem + TM_ZG + s * CH, Wbk, id_u, c > 0);
          commit(bars + I_WFREE + wu);
          if (c + INF_ZSLOTS < NCH) issue_z(c + INF_ZSLOTS, wz);
          else if (c == NCH - 1) commit(bars + I_ACC);
        }
        __syncwarp();
        wu = next2(wu);
        if (c + INF_ZSLOTS < NCH) wz = next2(wz);
      }
    }
  }
  ISTAMP(11);
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) {
    tc::fence_after_sync();
    tc::tmem_dealloc(tmem, TM_COLS);
  }
  ISTAMP(12);
}

}  // namespace

// ---- internal entry points used by navppo_kernels.cu -----------------------------------
size_t navppo_tcws_prep_bytes() { return (size_t)2 * WS_NET_BLOB; }

static long long* g_prof = nullptr;   // navppo_tc_profile: per-role cycle counters + one-tile event trace of CTA (0, 0)

int navppo_tcws_init() {
  NAV_CUDA_TRY(cudaFuncSetAttribute(mlp_grad_ws_kernel<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WS_SMEM_BYTES));
  NAV_CUDA_TRY(cudaFuncSetAttribute(mlp_grad_ws_kernel<3, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WS_SMEM_BYTES));
  NAV_CUDA_TRY(cudaFuncSetAttribute(mlp_grad_ws_kernel<3, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WS_SMEM_BYTES));
  NAV_CUDA_TRY(cudaFuncSetAttribute(mlp_infer_ws_kernel<1, WsInferArgs>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)INF_SMEM_BYTES));
  NAV_CUDA_TRY(cudaFuncSetAttribute(mlp_infer_ws_kernel<3, WsInferArgs>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)INF_SMEM_BYTES));
  NAV_CUDA_TRY(cudaFuncSetAttribute(mlp_infer_ws_kernel<1, WsFusedArgs>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FUSED_SMEM_BYTES));
  NAV_CUDA_TRY(cudaFuncSetAttribute(mlp_infer_ws_kernel<3, WsFusedArgs>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FUSED_SMEM_BYTES));
  return NAVSIM_OK;
}

// Re-tile the current weights for the tensor-core kernels (one launch; the rollout does it once for all its steps).
int navppo_tcws_prep_launch(const float* params, float* wprep, cudaStream_t s) {
  ws_prep_weights_kernel<<<dim3(24, 2), 256, 0, s>>>(params, reinterpret_cast<unsigned char*>(wprep));
  NAV_CUDA_TRY(cudaGetLastError());
  return NAVSIM_OK;
}

// One inference pass on the tensor cores over weights already re-tiled into `wprep`.
// `chained`: programmatic dependent launch — the kernel may start under its predecessor's tail (it waits for the
// predecessor before touching anything but the parameters and, with `weights_ready` — the predecessor is not the
// kernel that re-tiled them — the weights).
int navppo_tcws_infer_launch(const ppo::InferArgs& a, int mode, bool both_nets, int passes, const float* wprep, bool chained,
                             bool weights_ready, cudaStream_t s) {
  WsInferArgs ta;
  ta.g = a; ta.wprep = reinterpret_cast<const unsigned char*>(wprep); ta.mode = mode; ta.weights_ready = weights_ready ? 1 : 0;
  ta.prof = g_prof;
  cudaLaunchConfig_t lc{};
  lc.gridDim = dim3((a.T + 127) / 128, both_nets ? 2 : 1);
  lc.blockDim = dim3(INF_THREADS);
  lc.dynamicSmemBytes = INF_SMEM_BYTES;
  lc.stream = s;
  cudaLaunchAttribute attr{};
  attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr.val.programmaticStreamSerializationAllowed = 1;
  if (chained) { lc.attrs = &attr; lc.numAttrs = 1; }
  if (passes == 3) NAV_CUDA_TRY(cudaLaunchKernelEx(&lc, mlp_infer_ws_kernel<3, WsInferArgs>, ta));
  else NAV_CUDA_TRY(cudaLaunchKernelEx(&lc, mlp_infer_ws_kernel<1, WsInferArgs>, ta));
  return NAVSIM_OK;
}

// The whole step loop of PPO.rollout as ONE launch (weights already re-tiled into `wprep`): returns NAVSIM_EINVAL
// with `*unsupported = 1` when the simulator is in a configuration the fused kernel does not step (beam count other
// than 10, compacted-wall maps, noise / wheel-ramp options, start / goal tables) — the caller then chains kernels.
int navppo_tcws_rollout_launch(navsim* sim, const ppo::InferArgs& a, int H, int passes, const float* wprep, float* obs_rows,
                               float* next_obs, float* rew, uint8_t* done, uint8_t* arrive, uint8_t* trunc, float* ep_ret,
                               float* ep_path, int32_t* ep_len, int* unsupported, cudaStream_t s) {
  navsim_dev::DeviceView dv;
  if (int rc = navsim_device_view(sim, s, &dv)) return rc;
  *unsupported = (dv.variant != 0 || dv.c.S > navsim_dev::kCompactWalls || dv.c.B != NAVSIM_LIDAR_FEATS ||
                  navsim_dev::map_bytes_of(dv.c.B, dv.c.S) > FUSED_MAP_MAX) ? 1 : 0;
  if (*unsupported) return NAVSIM_EINVAL;
  WsFusedArgs ta;
  ta.g = a; ta.wprep = reinterpret_cast<const unsigned char*>(wprep); ta.mode = ppo::INFER_ACT; ta.weights_ready = 0; ta.prof = nullptr;
  ta.c = dv.c; ta.st = dv.st; ta.g_map = dv.map; ta.rt_tab = dv.rt; ta.stats = dv.stats;
  ta.H = H; ta.obs_rows = obs_rows; ta.next_obs = next_obs; ta.rew = rew; ta.done = done; ta.arrive = arrive; ta.trunc = trunc;
  ta.ep_ret = ep_ret; ta.ep_path = ep_path; ta.ep_len = ep_len;
  const dim3 grid((a.T + 127) / 128, 1);
  if (passes == 3) mlp_infer_ws_kernel<3, WsFusedArgs><<<grid, INF_THREADS, FUSED_SMEM_BYTES, s>>>(ta);
  else mlp_infer_ws_kernel<1, WsFusedArgs><<<grid, INF_THREADS, FUSED_SMEM_BYTES, s>>>(ta);
  NAV_CUDA_TRY(cudaGetLastError());
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
