// dmip_surrogate_tc.cu — K4 on the tensor cores: energy / score of the scatterometry posterior through the frozen ReLU
// surrogate [in <= 3] -> 256 -> 256 -> 256 -> [out <= 32]  (utils_scatterometry.py:9-16, :30-38; models/SNF.py:234-237;
// losses.py:349-371), forward and reverse sweep of 128 rows per CTA in ONE persistent kernel.
//
// Orientation: rows are M (one TMEM lane = one row = one thread of a row warp), features are N.  The two 256 x 256 layers
// are bf16x3 split products (x = hi + lo, both bf16: hi hi + hi lo + lo hi, fp32 accumulate — 16 mantissa bits per
// operand; the surrogate fixture is met with a 10x margin on E and the gradient, tools/study_k4_split.py) of
// tcgen05.mma M = 128, N = 128 / 256, K = 16 with the A operand in TENSOR MEMORY:
//   TMEM = two regions X, Y of 256 columns.  A GEMM reads its A operand from one region and accumulates into the other;
//   the epilogue turns each unit of 16 fp32 accumulator columns IN PLACE into 8 columns of bf16 hi pairs + 8 columns
//   of bf16 lo pairs — the A operand of one k-step of the next GEMM — so the regions swap roles layer by layer and no
//   activation ever touches shared memory: 192 of its 219 KB are the weight ring.
//   Weights: ONE packed image per matrix (128 out x 64 in tiles, K-major, 128-byte swizzle, hi and lo), read K-major by
//   the forward GEMMs and MN-major — the same bytes — by the reverse sweep (W^T), streamed from L2 by bulk-TMA in
//   32 KB stages.
//   Layer 0 (K = 3) and the input gradient (N = 3) are fp32 FFMA in the row threads; the output layer (N = 32) and its
//   transpose (K = 32) are small tcgen05 GEMMs; the per-row energy, its cotangent dE/df or the merged
//   likelihood cotangent stay in the row's threads (the ReLU patterns: one word per layer, chunk and thread in shared memory).
// Per tile: P0 | G1 P1 | G2 P2 | G3 P3 | G4 P4 | G5 P5 | G6 P6.  Each 256 x 256 GEMM runs its stages in the order
// (kb0 kb1) x both chunks (N = 256 instructions), (kb2 kb3) x c0, (kb2 kb3) x c1 (kb = 64-wide K block, c = 128-wide
// N chunk): chunk 0 is complete after 3/4 of the GEMM and its epilogue (all row warps) runs under the rest; the next GEMM starts on the
// K blocks chunk 0 produced while the row warps convert chunk 1.  P0 of the next tile runs inside P6 of the current one.
// Warps: 0-15 rows (warp & 3 = TMEM lane quarter; 32 features of each chunk per thread), 16 producer, 17 MMA issuer.
#include <stdlib.h>
#include <string.h>

#include "dmip_common.h"
#include "dmip_ptx.cuh"

namespace dmip {
namespace {

constexpr int kTRows = 128;
constexpr int kTThreads = 576;               // warps 0-15 rows, 16 producer, 17 MMA issuer
constexpr int kTStage = 32768;                 // hi 16 KB + lo 16 KB
constexpr int kTRing = 6;
constexpr int kImgW = 262144;                  // one 256 x 256 matrix: [part][c][kb] tiles of 16 KB
constexpr int kImgW3 = 32768;                  // output layer: [part][kb] tiles of 32 rows x 128 B
constexpr size_t kImgBytes = 2 * static_cast<size_t>(kImgW) + kImgW3;

constexpr int kOffW0 = kTRing * kTStage;       // float4[256]: W0[f][0..2], b0[f]
constexpr int kOffB1 = kOffW0 + 4096;          // float[256]
constexpr int kOffB2 = kOffB1 + 1024;
constexpr int kOffB3 = kOffB2 + 1024;          // float[32]
constexpr int kOffPart = kOffB3 + 128;         // float4[4][128]: the four feature quarters' shares of the input gradient
constexpr int kOffE = kOffPart + 8192;         // float[3][128]: energy shares of outputs 8.., 16.., 24.. (threads sub 1-3 of a row)
constexpr int kOffMask = kOffE + 2048;          // uint32[3 layers][2 words][512 row threads]: ReLU patterns of the thread's 64 features
constexpr int kOffBars = kOffMask + 12288;     // full[6] empty[6] a_ready[2] acc_full[2], tmem holder
constexpr int kTSmem = kOffBars + 256;

struct SurrTc {
  int mode, in_dim, out_dim;
  long long n, rpo;     // rows, rows per observation (0: one y row per x row)
  const float *W0, *b0, *W1, *b1, *W2, *b2, *W3, *b3;
  uint8_t* img;
  float a2, bb2, lambd;
  const float *x, *y;
  float *energy, *grad, *fx;
};

__device__ __forceinline__ void split_pair(float x0, float x1, uint32_t& hi, uint32_t& lo) {   // x0 in bits [0,16)
  hi = pack_bf16x2(x0, x1);
  lo = pack_bf16x2(x0 - __uint_as_float(hi << 16), x1 - __uint_as_float(hi & 0xFFFF0000u));
}

// ------------------------------------------------------------------------------------------------ weight images
// blocks 0-31: tile (matrix, part, c, kb) of W1 / W2; blocks 32-39: tile (part, kb) of the output layer (rows >= out_dim zero)
__global__ void __launch_bounds__(256) k_surr_pack(const __grid_constant__ SurrTc P) {
  const int blk = blockIdx.x;
  if (blk < 32) {
    const int m = blk >> 4, p = (blk >> 3) & 1, c = (blk >> 2) & 1, kb = blk & 3;
    const float* W = m ? P.W2 : P.W1;
    uint8_t* dst = P.img + static_cast<size_t>(m) * kImgW + ((p * 2 + c) * 4 + kb) * 16384;
    for (int e = threadIdx.x; e < 128 * 64; e += 256) {
      const int r = e >> 6, kk = e & 63;
      const float v = W[static_cast<size_t>(c * 128 + r) * 256 + kb * 64 + kk];
      uint32_t hi, lo;
      split_pair(v, 0.f, hi, lo);
      *reinterpret_cast<unsigned short*>(dst + sw128_offset(r, kk, 0)) = static_cast<unsigned short>((p ? lo : hi) & 0xFFFFu);
    }
  } else {
    const int i = blk - 32, p = i >> 2, kb = i & 3;
    uint8_t* dst = P.img + 2 * static_cast<size_t>(kImgW) + (p * 4 + kb) * 4096;
    for (int e = threadIdx.x; e < 32 * 64; e += 256) {
      const int r = e >> 6, kk = e & 63;
      const float v = r < P.out_dim ? P.W3[static_cast<size_t>(r) * 256 + kb * 64 + kk] : 0.f;
      uint32_t hi, lo;
      split_pair(v, 0.f, hi, lo);
      *reinterpret_cast<unsigned short*>(dst + sw128_offset(r, kk, 0)) = static_cast<unsigned short>((p ? lo : hi) & 0xFFFFu);
    }
  }
}

// ------------------------------------------------------------------------------------------------ row-warp pieces
struct TBars {
  uint64_t *full, *empty, *a_ready, *acc_full;
};

// m |= (v > 0) << e as one predicated OR (the C expression made ptxas park the values in local memory)
__device__ __forceinline__ void mask_bit(uint32_t& m, float v, int e) {
  asm("{\n\t.reg .pred p;\n\tsetp.gt.f32 p, %1, 0f00000000;\n\t@p or.b32 %0, %0, %2;\n\t}" : "+r"(m) : "f"(v), "r"(1u << e));
}
// One UNIT = 16 features = the K of one MMA = 16 fp32 accumulator columns, which the epilogue replaces IN PLACE by 8
// columns of bf16 hi pairs + 8 columns of bf16 lo pairs: the A operand of k-step s of the next GEMM sits at columns
// 16 s (hi) and 16 s + 8 (lo) of the region.
__device__ __forceinline__ void store_unit(uint32_t taddr, const float (&v)[16]) {
  uint32_t hi[8], lo[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) split_pair(v[2 * e], v[2 * e + 1], hi[e], lo[e]);
  tmem_st8(taddr, hi);
  tmem_st8(taddr + 8, lo);
}
// the A operand of this warp's rows and features is complete
__device__ __forceinline__ void publish(uint64_t* a_ready, int lane) {
  tc_wait_st();
  tc_fence_before();
  __syncwarp();
  if (lane == 0) mbar_arrive(a_ready);
}
// A row thread owns 32 features of chunk 0 and 32 features of chunk 1 of every 256-wide layer (two units each): ALL
// sixteen row warps convert chunk 0 as soon as it is complete (3/4 into the GEMM) and then chunk 1 — the next GEMM waits
// for half an epilogue, not a whole one.  The ReLU patterns live in shared memory (one word per layer, part and thread);
// the phases are small loops: fully unrolled over 128 features per thread, the row warps walked 200 KB of code at
// different places and spent half of their issue slots waiting for instructions (ncu: stall_no_inst).

// forward epilogue of one part (two units) of a hidden layer: z = acc + bias, ReLU pattern, relu, in place
__device__ __forceinline__ void fwd_part(uint32_t cols, const float* bias, uint32_t* mask_word) {
  uint32_t u[16];
  tmem_ld16(cols, u);
  uint32_t m = 0;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    tc_wait_ld();
    float v[16];
#pragma unroll
    for (int e4 = 0; e4 < 4; ++e4) {
      const float4 b = *reinterpret_cast<const float4*>(bias + 16 * i + 4 * e4);
      v[4 * e4 + 0] = fmaxf(__uint_as_float(u[4 * e4 + 0]) + b.x, 0.f);
      v[4 * e4 + 1] = fmaxf(__uint_as_float(u[4 * e4 + 1]) + b.y, 0.f);
      v[4 * e4 + 2] = fmaxf(__uint_as_float(u[4 * e4 + 2]) + b.z, 0.f);
      v[4 * e4 + 3] = fmaxf(__uint_as_float(u[4 * e4 + 3]) + b.w, 0.f);
    }
    if (i == 0) tmem_ld16(cols + 16, u);
#pragma unroll
    for (int e = 0; e < 16; ++e) mask_bit(m, v[e], 16 * i + e);
    store_unit(cols + 16 * i, v);
  }
  *mask_word = m;
}
// reverse epilogue of one part of a hidden layer: hbar masked by the layer's ReLU pattern, in place
__device__ __forceinline__ void bwd_part(uint32_t cols, uint32_t m) {
  uint32_t u[16];
  tmem_ld16(cols, u);
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    tc_wait_ld();
    float v[16];
#pragma unroll
    for (int e = 0; e < 16; ++e) v[e] = ((m >> (16 * i + e)) & 1u) ? __uint_as_float(u[e]) : 0.f;
    if (i == 0) tmem_ld16(cols + 16, u);
    store_unit(cols + 16 * i, v);
  }
}

__constant__ int c_kb[8] = {0, 1, 0, 1, 2, 3, 2, 3};
__constant__ int c_ch[8] = {0, 0, 1, 1, 0, 0, 1, 1};

__global__ void __launch_bounds__(kTThreads, 1) k_surrogate_tc(const __grid_constant__ SurrTc P) {
  extern __shared__ __align__(1024) uint8_t smem[];
  float4* sW0 = reinterpret_cast<float4*>(smem + kOffW0);
  float* sB1 = reinterpret_cast<float*>(smem + kOffB1);
  float* sB2 = reinterpret_cast<float*>(smem + kOffB2);
  float* sB3 = reinterpret_cast<float*>(smem + kOffB3);
  float4* sPart = reinterpret_cast<float4*>(smem + kOffPart);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBars);
  TBars B = {bars, bars + kTRing, bars + 2 * kTRing, bars + 2 * kTRing + 2};
  uint32_t* holder = reinterpret_cast<uint32_t*>(bars + 2 * kTRing + 4);
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  if ((smem_u32(smem) & 1023u) != 0u) __trap();
  if (threadIdx.x == 0) {
    for (int i = 0; i < kTRing; ++i) {
      mbar_init(&B.full[i], 1);
      mbar_init(&B.empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&B.a_ready[i], 16);
      mbar_init(&B.acc_full[i], 1);
    }
    fence_barrier_init();
  }
  for (int f = threadIdx.x; f < 256; f += kTThreads) {
    float4 w = make_float4(0.f, 0.f, 0.f, P.b0[f]);
    w.x = P.W0[f * P.in_dim];
    if (P.in_dim > 1) w.y = P.W0[f * P.in_dim + 1];
    if (P.in_dim > 2) w.z = P.W0[f * P.in_dim + 2];
    sW0[f] = w;
    sB1[f] = P.b1[f];
    sB2[f] = P.b2[f];
    if (f < 32) sB3[f] = f < P.out_dim ? P.b3[f] : 0.f;
  }
  if (warp == 17) tmem_alloc<512>(holder);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *holder;
  const long long n_tiles = (P.n + kTRows - 1) / kTRows;

  if (warp == 16) {
    // ---------------------------------------------------------------------------------------------- producer
    int s = 0;
    uint32_t ph = 0;
    const uint8_t* img3 = P.img + 2 * static_cast<size_t>(kImgW);
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
#pragma unroll 1
      for (int g = 1; g <= 6; ++g) {
        if (g == 3 || g == 4) {
          mbar_wait(&B.empty[s], ph ^ 1u, 0x100 + s);
          if (elect_one()) {
            uint8_t* st = smem + s * kTStage;
            mbar_arrive_expect_tx(&B.full[s], kTStage);
            bulk_g2s(st, img3, 16384, &B.full[s]);
            bulk_g2s(st + 16384, img3 + 16384, 16384, &B.full[s]);
          }
          __syncwarp();
          if (++s == kTRing) { s = 0; ph ^= 1u; }
          continue;
        }
        const uint8_t* img = P.img + ((g == 1 || g == 6) ? 0 : kImgW);
        const bool bwd = g >= 5;
        // K blocks 0, 1 (they need only chunk 0 of the operand): DOUBLE stages (two ring slots, 64 KB) for N = 256
        // instructions; slots are consumed in even numbers per GEMM, so a double stage starts on an even slot
#pragma unroll 1
        for (int kb = 0; kb < 2; ++kb) {
          mbar_wait(&B.empty[s], ph ^ 1u, 0x100 + s);
          mbar_wait(&B.empty[s + 1], ph ^ 1u, 0x101 + s);
          if (elect_one()) {
            uint8_t* st = smem + s * kTStage;
            mbar_arrive_expect_tx(&B.full[s], 2 * kTStage);
            mbar_arrive(&B.full[s + 1]);   // nobody waits for it: the slot's barrier just keeps step with the ring's phase
            if (!bwd) {
              // forward: all 256 out features x in features [64 kb, +64): [hi c0 | hi c1 | lo c0 | lo c1]
#pragma unroll
              for (int p = 0; p < 2; ++p)
#pragma unroll
                for (int c = 0; c < 2; ++c)
                  bulk_g2s(st + p * 32768 + c * 16384, img + ((p * 2 + c) * 4 + kb) * 16384, 16384, &B.full[s]);
            } else {
              // reverse: K = out features [64 kb, +64), N = all 256 in features = four MN groups of 64, 8 KB apart
              const int ct = kb >> 1, half = kb & 1;
#pragma unroll
              for (int p = 0; p < 2; ++p)
#pragma unroll
                for (int gi = 0; gi < 4; ++gi)
                  bulk_g2s(st + p * 32768 + gi * 8192, img + ((p * 2 + ct) * 4 + gi) * 16384 + half * 8192, 8192, &B.full[s]);
            }
          }
          __syncwarp();
          s += 2;
          if (s == kTRing) { s = 0; ph ^= 1u; }
        }
#pragma unroll 1
        for (int i = 4; i < 8; ++i) {
          const int kb = c_kb[i], c = c_ch[i];
          mbar_wait(&B.empty[s], ph ^ 1u, 0x100 + s);
          if (elect_one()) {
            uint8_t* st = smem + s * kTStage;
            mbar_arrive_expect_tx(&B.full[s], kTStage);
            if (!bwd) {
              // forward: out features [128 c, +128) x in features [64 kb, +64): one tile per part
              bulk_g2s(st, img + ((0 * 2 + c) * 4 + kb) * 16384, 16384, &B.full[s]);
              bulk_g2s(st + 16384, img + ((1 * 2 + c) * 4 + kb) * 16384, 16384, &B.full[s]);
            } else {
              // reverse: K = out features [64 kb, +64) = rows [64 (kb & 1), +64) of the tiles of row chunk kb >> 1,
              // N = in features [128 c, +128) = the tiles of K blocks 2 c, 2 c + 1 (two MN groups of 64, 8 KB apart)
              const int ct = kb >> 1, half = kb & 1;
#pragma unroll
              for (int p = 0; p < 2; ++p)
#pragma unroll
                for (int i2 = 0; i2 < 2; ++i2)
                  bulk_g2s(st + p * 16384 + i2 * 8192, img + ((p * 2 + ct) * 4 + 2 * c + i2) * 16384 + half * 8192, 8192,
                           &B.full[s]);
            }
          }
          __syncwarp();
          if (++s == kTRing) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 17) {
    // ---------------------------------------------------------------------------------------------- MMA issuer
    int s = 0;
    uint32_t ph = 0;
    uint32_t ka0 = 0, ka1 = 0;   // waits done on a_ready[0], a_ready[1]
    uint32_t nf0 = 0, nf1 = 0;   // commits made on acc_full[0], acc_full[1]
    int last_full = -1;          // barrier of the last commit of the previous GEMM
    const uint32_t base16 = (smem_u32(smem) & 0x3FFFFu) >> 4;
    const uint64_t dK = umma_smem_desc_sw128(0);
    const uint64_t dMN = umma_smem_desc(0, 8192, 1024);
    const uint64_t dMN3 = umma_smem_desc(0, 4096, 1024);
    constexpr uint32_t iF = umma_idesc_bf16(128, 128);
    constexpr uint32_t iF3 = umma_idesc_bf16(128, 32);
    constexpr uint32_t iB = umma_idesc_bf16_major(128, 128, 0, 1);
    constexpr uint32_t iF2 = umma_idesc_bf16(128, 256);
    constexpr uint32_t iB2 = umma_idesc_bf16_major(128, 256, 0, 1);
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
#pragma unroll 1
      for (int g = 1; g <= 6; ++g) {
        const uint32_t A = tmem_base + ((g & 1) ? 0u : 256u);
        const uint32_t D = tmem_base + ((g & 1) ? 256u : 0u);
        // The accumulator region of this GEMM is the operand region of the previous one: every MMA of it has retired.
        // (-DDMIP_K4_RELAXED drops this wait and relies on tcgen05.mma executing in issue order: same results on
        // 100,003 rows, 1.3 % faster — not kept, the order of an operand read against a later accumulator write is not
        // a documented guarantee.)
#ifndef DMIP_K4_RELAXED
        if (last_full == 0) mbar_wait(&B.acc_full[0], (nf0 - 1u) & 1u, 0x200);
        else if (last_full == 1) mbar_wait(&B.acc_full[1], (nf1 - 1u) & 1u, 0x201);
#endif
        if (g == 3) {
          mbar_wait(&B.a_ready[0], ka0 & 1u, 0x210); ++ka0;
          mbar_wait(&B.full[s], ph, 0x220 + s);
          tc_fence_after();
          const uint32_t st16 = base16 + static_cast<uint32_t>(s) * (kTStage >> 4);
#pragma unroll 1
          for (int h = 0; h < 2; ++h) {   // K = features of chunk 0, then of chunk 1
            if (h == 1) {
              mbar_wait(&B.a_ready[1], ka1 & 1u, 0x211); ++ka1;
              tc_fence_after();
            }
            if (elect_one()) {
#pragma unroll
              for (int kk = 0; kk < 8; ++kk) {
                const int k16 = 8 * h + kk;
                const uint32_t a_hi = A + 16u * k16, a_lo = a_hi + 8u;
                const uint64_t b_hi = dK | (st16 + (k16 >> 2) * 256u + (k16 & 3) * 2u), b_lo = b_hi + 1024u;
                umma_ts(D, a_hi, b_hi, iF3, k16 > 0 ? 1u : 0u);
                umma_ts(D, a_hi, b_lo, iF3, 1u);
                umma_ts(D, a_lo, b_hi, iF3, 1u);
              }
              if (h == 1) {
                tc_commit(&B.empty[s]);
                tc_commit(&B.acc_full[0]);
              }
            }
            __syncwarp();
          }
          ++nf0;
          last_full = 0;
          if (++s == kTRing) { s = 0; ph ^= 1u; }
          continue;
        }
        if (g == 4) {
          mbar_wait(&B.a_ready[0], ka0 & 1u, 0x210); ++ka0;
          mbar_wait(&B.full[s], ph, 0x220 + s);
          tc_fence_after();
          const uint32_t st16 = base16 + static_cast<uint32_t>(s) * (kTStage >> 4);
          if (elect_one()) {
#pragma unroll
            for (int c = 0; c < 2; ++c) {
#pragma unroll
              for (int ks = 0; ks < 2; ++ks) {
                const uint32_t a_hi = A + 32u + 16u * ks, a_lo = a_hi + 8u;   // fbar sits beside f: columns 32-63
                const uint64_t b_hi = dMN3 | (st16 + c * 512u + ks * 128u), b_lo = b_hi + 1024u;
                umma_ts(D + 128u * c, a_hi, b_hi, iB, ks > 0 ? 1u : 0u);
                umma_ts(D + 128u * c, a_hi, b_lo, iB, 1u);
                umma_ts(D + 128u * c, a_lo, b_hi, iB, 1u);
              }
              tc_commit(&B.acc_full[c]);
            }
            tc_commit(&B.empty[s]);
          }
          __syncwarp();
          ++nf0; ++nf1;
          last_full = 1;
          if (++s == kTRing) { s = 0; ph ^= 1u; }
          continue;
        }
        const bool bwd = g >= 5;
        mbar_wait(&B.a_ready[0], ka0 & 1u, 0x210); ++ka0;
#pragma unroll 1
        for (int kb = 0; kb < 2; ++kb) {   // K blocks 0, 1 for both chunks at once: N = 256 from a double stage
          mbar_wait(&B.full[s], ph, 0x220 + s);
          tc_fence_after();
          const uint32_t st16 = base16 + static_cast<uint32_t>(s) * (kTStage >> 4);
          if (elect_one()) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              const int sg = 4 * kb + ks;
              const uint32_t a_hi = A + 16u * sg, a_lo = a_hi + 8u;
              const uint64_t b_hi = bwd ? (dMN | (st16 + ks * 128u)) : (dK | (st16 + ks * 2u));
              const uint64_t b_lo = b_hi + 2048u;
              const uint32_t idesc = bwd ? iB2 : iF2;
              umma_ts(D, a_hi, b_hi, idesc, (kb | ks) != 0 ? 1u : 0u);
              umma_ts(D, a_hi, b_lo, idesc, 1u);
              umma_ts(D, a_lo, b_hi, idesc, 1u);
            }
            tc_commit(&B.empty[s]);
            tc_commit(&B.empty[s + 1]);
          }
          __syncwarp();
          s += 2;
          if (s == kTRing) { s = 0; ph ^= 1u; }
        }
#pragma unroll 1
        for (int i = 4; i < 8; ++i) {
          const int kb = c_kb[i], c = c_ch[i];
          if (i == 4) { mbar_wait(&B.a_ready[1], ka1 & 1u, 0x211); ++ka1; }
          mbar_wait(&B.full[s], ph, 0x220 + s);
          tc_fence_after();
          const uint32_t st16 = base16 + static_cast<uint32_t>(s) * (kTStage >> 4);
          if (elect_one()) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              const int sg = 4 * kb + ks;
              const uint32_t a_hi = A + 16u * sg, a_lo = a_hi + 8u;
              const uint64_t b_hi = bwd ? (dMN | (st16 + ks * 128u)) : (dK | (st16 + ks * 2u));
              const uint64_t b_lo = b_hi + 1024u;
              const uint32_t idesc = bwd ? iB : iF;
              umma_ts(D + 128u * c, a_hi, b_hi, idesc, 1u);
              umma_ts(D + 128u * c, a_hi, b_lo, idesc, 1u);
              umma_ts(D + 128u * c, a_lo, b_hi, idesc, 1u);
            }
            tc_commit(&B.empty[s]);
            if (i == 5) tc_commit(&B.acc_full[0]);
            if (i == 7) tc_commit(&B.acc_full[1]);
          }
          __syncwarp();
          if (++s == kTRing) { s = 0; ph ^= 1u; }
        }
        ++nf0; ++nf1;
        last_full = 1;
      }
    }
  } else {
    // ---------------------------------------------------------------------------------------------- row warps
    // warp = 4 sub + q: q = TMEM lane quarter (rows 32 q ..); the thread owns features [32 sub, +32) of chunk 0 (part 0)
    // and [128 + 32 sub, +32) of chunk 1 (part 1) of every 256-wide layer
    const int q = warp & 3, sub = warp >> 2;
    const int row = q * 32 + lane;
    const uint32_t lt = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const uint32_t X = lt + 32u * sub, Y = lt + 256u + 32u * sub;   // part p of a region: + 128 p
    uint32_t* mk = reinterpret_cast<uint32_t*>(smem + kOffMask) + threadIdx.x;   // [layer * 1024 + part * 512]
    float* sE = reinterpret_cast<float*>(smem + kOffE);
    uint32_t nw0 = 0, nw1 = 0;   // waits done on acc_full[0], acc_full[1]
    const int od = P.out_dim, idim = P.in_dim, mode = P.mode;
    const float a2 = P.a2, bb2 = P.bb2, lambd = P.lambd;
    const float4* w0 = sW0 + 32 * sub;
    auto wait_part = [&](int p, uint32_t tag) {
      if (p == 0) { mbar_wait(&B.acc_full[0], nw0 & 1u, tag); ++nw0; }
      else { mbar_wait(&B.acc_full[1], nw1 & 1u, tag + 1); ++nw1; }
      tc_fence_after();
    };
    // P0 of one part: h1 = relu(W0 x + b0) for this thread's 32 features of chunk p, straight into region X
    auto p0_part = [&](int p, float x0, float x1, float x2) {
      uint32_t m = 0;
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        float v[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) {
          const float4 w = w0[128 * p + 16 * i + e];
          v[e] = fmaxf(fmaf(x2, w.z, fmaf(x1, w.y, fmaf(x0, w.x, w.w))), 0.f);
          mask_bit(m, v[e], 16 * i + e);
        }
        store_unit(X + 128 * p + 16 * i, v);
      }
      mk[p * 512] = m;
      publish(&B.a_ready[p], lane);
    };
    auto load_x = [&](long long g, float& x0, float& x1, float& x2) {
      x0 = x1 = x2 = 0.f;
      if (g < P.n) {
        x0 = P.x[g * idim];
        if (idim > 1) x1 = P.x[g * idim + 1];
        if (idim > 2) x2 = P.x[g * idim + 2];
      }
    };
    float x0, x1, x2;
    load_x(static_cast<long long>(blockIdx.x) * kTRows + row, x0, x1, x2);
    // ---- P0 of the CTA's first tile; every later tile's P0 runs inside the previous tile's P6 (below)
#pragma unroll 1
    for (int p = 0; p < 2; ++p) p0_part(p, x0, x1, x2);
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const long long grow = tile * kTRows + row;
      const bool live = grow < P.n;
      // ---- P1: h2 = relu(acc + b1) (G1 accumulates into Y);  P2: h3 = relu(acc + b2) (G2 -> X)
#pragma unroll 1
      for (int l = 1; l <= 2; ++l) {
#pragma unroll 1
        for (int p = 0; p < 2; ++p) {
          wait_part(p, 0x300);
          fwd_part((l == 1 ? Y : X) + 128 * p, (l == 1 ? sB1 : sB2) + 128 * p + 32 * sub, mk + l * 1024 + p * 512);
          publish(&B.a_ready[p], lane);
        }
      }
      // ---- P3: f = acc + b3 (G3 -> Y[0, 32)), energy, cotangent: the four threads of a row take 8 outputs each and write
      //      their share of the cotangent operand (two units: hi 8 + lo 8 columns each) BESIDE f, into Y[32, 64) — nobody
      //      overwrites a column another thread of the row still has to read
      {
        // the observation's 8 values first: their latency runs under the wait for the output layer's GEMM
        const float* yrow = P.y + (P.rpo > 0 ? grow / P.rpo : grow) * od;
        float yv8[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) yv8[e] = (live && 8 * sub + e < od) ? yrow[8 * sub + e] : 0.f;
        wait_part(0, 0x320);
        uint32_t u[8];
        tmem_ld8(lt + 256u + 8u * sub, u);
        tc_wait_ld();
        float w[8];
        float E = 0.f;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          float we = 0.f;
          const int o = 8 * sub + e;
          if (o < od) {
            const float f = __uint_as_float(u[e]) + sB3[o];
            const float yv = yv8[e];
            const float p = a2 * f * f + bb2;
            const float ip = __frcp_rn(p);
            const float r = yv - f;
            if (mode == 0) {
              E += 0.5f * __logf(p) + 0.5f * r * r * ip;
              we = (a2 * f - r - a2 * f * r * r * ip) * ip;                   // dE/df  
            } else {
              we = (-a2 * f + r + a2 * r * r * f) * ip;                       // -a^2 v1 + v2 + a^2 v3  (losses.py:354-368)
            }
            if (live && P.fx) P.fx[grow * od + o] = f;
          }
          w[e] = we;
        }
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) split_pair(w[2 * e], w[2 * e + 1], hi[e], lo[e]);
        const uint32_t dst = lt + 256u + 32u + 16u * (sub >> 1) + 4u * (sub & 1);   // unit sub / 2, pairs 4 (sub % 2) ..
        tmem_st4(dst, hi[0], hi[1], hi[2], hi[3]);
        tmem_st4(dst + 8u, lo[0], lo[1], lo[2], lo[3]);
        publish(&B.a_ready[0], lane);
        if (mode == 0 && P.energy) {
          if (sub != 0) sE[(sub - 1) * 128 + row] = E;
          asm volatile("bar.sync 2, 512;" ::: "memory");
          if (sub == 0 && live) {
            E += sE[row] + sE[128 + row] + sE[256 + row];
            E += lambd * (fmaxf(x0 - 1.f, 0.f) + fmaxf(-1.f - x0, 0.f) + fmaxf(x1 - 1.f, 0.f) + fmaxf(-1.f - x1, 0.f) +
                          fmaxf(x2 - 1.f, 0.f) + fmaxf(-1.f - x2, 0.f));
            P.energy[grow] = E;
          }
        }
      }
      // ---- P4: h3bar = acc masked by h3 > 0 (G4 -> X);  P5: h2bar (G5 -> Y)
#pragma unroll 1
      for (int l = 2; l >= 1; --l) {
#pragma unroll 1
        for (int p = 0; p < 2; ++p) {
          wait_part(p, 0x330);
          bwd_part((l == 2 ? X : Y) + 128 * p, mk[l * 1024 + p * 512]);
          publish(&B.a_ready[p], lane);
        }
      }
      // ---- P6: h1bar (G6 -> X) masked, xbar = W0^T h1bar — and, part by part, the NEXT tile's P0 into the columns this
      //      thread has just read (its own 32 columns of the part), so that G1 of the next tile starts under this tile's tail
      const long long ntile = tile + gridDim.x;
      const bool has_next = ntile < n_tiles;
      float nx0, nx1, nx2;
      load_x(has_next ? ntile * kTRows + row : P.n, nx0, nx1, nx2);
      float g0 = 0.f, g1 = 0.f, g2 = 0.f;
#pragma unroll 1
      for (int p = 0; p < 2; ++p) {
        wait_part(p, 0x350);
        const uint32_t m = mk[p * 512];
        uint32_t u[16];
        tmem_ld16(X + 128 * p, u);
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          tc_wait_ld();
          float v[16];
#pragma unroll
          for (int e = 0; e < 16; ++e) v[e] = ((m >> (16 * i + e)) & 1u) ? __uint_as_float(u[e]) : 0.f;
          if (i == 0) tmem_ld16(X + 128 * p + 16, u);
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            const float4 w = w0[128 * p + 16 * i + e];
            g0 = fmaf(v[e], w.x, g0);
            g1 = fmaf(v[e], w.y, g1);
            g2 = fmaf(v[e], w.z, g2);
          }
        }
        if (has_next) p0_part(p, nx0, nx1, nx2);
      }
      tc_fence_before();
      if (sub != 0) sPart[(sub - 1) * 128 + row] = make_float4(g0, g1, g2, 0.f);
      asm volatile("bar.sync 1, 512;" ::: "memory");
      if (sub == 0 && live) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
          const float4 o = sPart[k * 128 + row];
          g0 += o.x; g1 += o.y; g2 += o.z;
        }
        if (mode == 0) {
          g0 += lambd * ((x0 > 1.f ? 1.f : 0.f) - (x0 < -1.f ? 1.f : 0.f));
          g1 += lambd * ((x1 > 1.f ? 1.f : 0.f) - (x1 < -1.f ? 1.f : 0.f));
          g2 += lambd * ((x2 > 1.f ? 1.f : 0.f) - (x2 < -1.f ? 1.f : 0.f));
        }
        P.grad[grow * idim] = g0;
        if (idim > 1) P.grad[grow * idim + 1] = g1;
        if (idim > 2) P.grad[grow * idim + 2] = g2;
      }
      x0 = nx0; x1 = nx1; x2 = nx2;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 17) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace

// The tensor-core path serves the reference's surrogate shape ([in <= 3] -> 256 -> 256 -> 256 -> [out <= 32]); any other
// net runs the fp32 FFMA kernel of dmip_surrogate.cu.  DMIP_SURROGATE_PATH=ffma forces the FFMA kernel (A/B runs, tests).
bool surrogate_tc_supported(const DmipMlp& net) {
  const char* force = getenv("DMIP_SURROGATE_PATH");
  if (force && strcmp(force, "ffma") == 0) return false;
  return net.n_layers == 4 && net.in_dim >= 1 && net.in_dim <= 3 && net.width[0] == 256 && net.width[1] == 256 &&
         net.width[2] == 256 && net.out_dim >= 1 && net.out_dim <= 32 && net.width[3] == net.out_dim;
}

size_t surrogate_tc_workspace() { return kImgBytes; }

int launch_surrogate_tc(const DmipSurrogate* d, void* images, cudaStream_t s) {
  static int n_sm = 0;
  static bool ready[64] = {};
  int dev = 0;
  DMIP_CHECK_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64 || !ready[dev]) {
    DMIP_CHECK_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
    DMIP_CHECK_CUDA(cudaFuncSetAttribute(k_surrogate_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, kTSmem));
    if (dev >= 0 && dev < 64) ready[dev] = true;
  }
  DMIP_REQUIRE((reinterpret_cast<uintptr_t>(images) & 127u) == 0, "surrogate workspace must be 128-byte aligned");
  const DmipMlp& net = d->net;
  SurrTc P = {};
  P.mode = d->mode;
  P.in_dim = net.in_dim;
  P.out_dim = net.out_dim;
  P.n = d->n;
  P.rpo = d->rows_per_obs;
  P.W0 = net.W[0]; P.b0 = net.b[0];
  P.W1 = net.W[1]; P.b1 = net.b[1];
  P.W2 = net.W[2]; P.b2 = net.b[2];
  P.W3 = net.W[3]; P.b3 = net.b[3];
  P.img = static_cast<uint8_t*>(images);
  P.a2 = d->a * d->a;
  P.bb2 = d->b * d->b;
  P.lambd = d->lambd_bd;
  P.x = d->x;
  P.y = d->y;
  P.energy = d->energy;
  P.grad = d->grad;
  P.fx = d->fx;
  k_surr_pack<<<40, 256, 0, s>>>(P);
  DMIP_CHECK_CUDA(cudaGetLastError());
  count_launch();
  const long long tiles = (d->n + kTRows - 1) / kTRows;
  k_surrogate_tc<<<static_cast<unsigned>(tiles < n_sm ? tiles : n_sm), kTThreads, kTSmem, s>>>(P);
  DMIP_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return DMIP_OK;
}

}  // namespace dmip
