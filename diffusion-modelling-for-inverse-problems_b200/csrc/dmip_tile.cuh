// dmip_tile.cuh — the fp32 32-row tile engine shared by the loss and surrogate kernels: activations live in shared
// memory transposed ([k][row], row stride kLd), a layer is a 32 x N x K register-tiled FFMA GEMM against weights
// stored [contraction][output] in global memory (L2/L1 resident).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace dmip {
namespace {

constexpr int kRows = 32;
constexpr int kLd = 36;
constexpr int kMaxW = 512;
constexpr int kThreadsL = 512;   // 16 warps: 4 per scheduler hide the L1/L2 latency of the weight loads

// out[n*kLd + r] = sum_k in[k*kLd + r] * Wt[k*N + n], r < 32, n < N.  512 threads: thread (rg, ng) owns rows 8 rg .. 8 rg + 7
// and columns ng + 128 i, i < 4; the operands of step k+1 are fetched while step k's 32 FMAs issue.  Ends with
// __syncthreads().
__device__ void tile_gemm_direct(const float* in, float* out, const float* __restrict__ Wt, int K, int N) {
  const int t = threadIdx.x;
  const int rg = t >> 7, ng = t & 127;
  for (int nb = 0; nb < N; nb += 512) {
    float acc[8][4];
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[r][i] = 0.f;
    const int ncols = (N - nb - ng + 127) >> 7;   // columns of this thread inside [nb, nb + 512)
    if (ncols > 0) {
      const float* wp = Wt + nb + ng;
      float4 a0 = *reinterpret_cast<const float4*>(in + rg * 8);
      float4 a1 = *reinterpret_cast<const float4*>(in + rg * 8 + 4);
      float w[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) w[i] = (i < ncols) ? __ldg(wp + 128 * i) : 0.f;
      for (int k = 0; k < K; ++k) {
        const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        float wc[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) wc[i] = w[i];
        if (k + 1 < K) {   // prefetch step k+1
          a0 = *reinterpret_cast<const float4*>(in + (k + 1) * kLd + rg * 8);
          a1 = *reinterpret_cast<const float4*>(in + (k + 1) * kLd + rg * 8 + 4);
          const float* wn = wp + static_cast<size_t>(k + 1) * N;
#pragma unroll
          for (int i = 0; i < 4; ++i) w[i] = (i < ncols) ? __ldg(wn + 128 * i) : 0.f;
        }
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
          for (int i = 0; i < 4; ++i) acc[r][i] = fmaf(a[r], wc[i], acc[r][i]);
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int n = nb + ng + 128 * i;
      if (n < N) {
#pragma unroll
        for (int r = 0; r < 8; ++r) out[n * kLd + rg * 8 + r] = acc[r][i];
      }
    }
  }
  __syncthreads();
}


// ---- staged variant: the weights stream through shared memory in 16-k slabs (cp.async, double buffered), so their L2
// latency is paid one slab (16 k-steps) ahead instead of one k-step ahead, and every thread reads its four weights with
// one conflict-free LDS.128.  Thread (rg, ng) owns R rows and columns 4 ng .. 4 ng + 3; N = 64 R (512 or 256).  A warp
// covers ALL 32 rows x 4 R columns (lane = row group x column group), so per k-step it touches 128 B of activations
// and 16 R B of weights: 3 shared-memory wavefronts per 32 FFMA instead of 6 with one row group per warp.
constexpr int kSlabK = 16;
constexpr int kWbufFloats = 2 * kSlabK * kMaxW;   // 64 KB

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst))),
               "l"(gsrc)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int kN>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(kN) : "memory"); }

template <int R>
__device__ void tile_gemm_staged(const float* in, float* out, const float* __restrict__ Wt, int K, float* wbuf) {
  constexpr int N = 64 * R;              // 512 threads x (4 columns x R rows) = 32 rows x N columns
  constexpr int kGroups = N / 4;         // column groups
  const int t = threadIdx.x;
  const int rg = (t & 31) / R, ng = (t >> 5) * R + (t & 31) % R;   // 32 / R row groups x R column groups per warp
  static_assert(kGroups == (kThreadsL / 32) * R, "16 warps x R column groups cover the N / 4 groups");
  float acc[R][4];
#pragma unroll
  for (int r = 0; r < R; ++r)
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[r][i] = 0.f;
  const int n_slabs = K / kSlabK;
  constexpr int kChunks = kSlabK * N / 4;   // 16-byte chunks per slab
  auto issue = [&](int s) {
    const float* src = Wt + static_cast<size_t>(s) * kSlabK * N;
    float* dst = wbuf + (s & 1) * kSlabK * N;
    for (int c = t; c < kChunks; c += kThreadsL) cp_async16(dst + c * 4, src + c * 4);
    cp_async_commit();
  };
  issue(0);
  for (int s = 0; s < n_slabs; ++s) {
    if (s + 1 < n_slabs) {
      issue(s + 1);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();   // slab s has landed for every thread
    const float* w = wbuf + (s & 1) * kSlabK * N + ng * 4;
    const float* a = in + s * kSlabK * kLd + rg * R;
#pragma unroll
    for (int k = 0; k < kSlabK; ++k) {
      const float4 wv = *reinterpret_cast<const float4*>(w + k * N);
      float av[R];
      if (R == 8) {
        const float4 a0 = *reinterpret_cast<const float4*>(a + k * kLd);
        const float4 a1 = *reinterpret_cast<const float4*>(a + k * kLd + 4);
        av[0] = a0.x; av[1] = a0.y; av[2] = a0.z; av[3] = a0.w;
        av[R > 4 ? 4 : 0] = a1.x; av[R > 4 ? 5 : 0] = a1.y; av[R > 4 ? 6 : 0] = a1.z; av[R > 4 ? 7 : 0] = a1.w;
      } else {
        const float4 a0 = *reinterpret_cast<const float4*>(a + k * kLd);
        av[0] = a0.x; av[1] = a0.y; av[2] = a0.z; av[3] = a0.w;
      }
#pragma unroll
      for (int r = 0; r < R; ++r) {
        acc[r][0] = fmaf(av[r], wv.x, acc[r][0]);
        acc[r][1] = fmaf(av[r], wv.y, acc[r][1]);
        acc[r][2] = fmaf(av[r], wv.z, acc[r][2]);
        acc[r][3] = fmaf(av[r], wv.w, acc[r][3]);
      }
    }
    __syncthreads();   // everyone is done with this buffer before slab s+2 is copied into it
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float* o = out + (ng * 4 + i) * kLd + rg * R;
    if (R == 8) {
      *reinterpret_cast<float4*>(o) = make_float4(acc[0][i], acc[1][i], acc[2][i], acc[3][i]);
      *reinterpret_cast<float4*>(o + 4) = make_float4(acc[R > 4 ? 4 : 0][i], acc[R > 4 ? 5 : 0][i], acc[R > 4 ? 6 : 0][i], acc[R > 4 ? 7 : 0][i]);
    } else {
      *reinterpret_cast<float4*>(o) = make_float4(acc[0][i], acc[1][i], acc[2][i], acc[3][i]);
    }
  }
  __syncthreads();
}

// ---- skinny outputs (N <= 32 with a long contraction: the nets' last layer, the first layer of a reverse sweep).  The
// direct path would leave all but 4 N threads idle and walk K serially behind one global load per step (ncu: 30 % of
// k_jets_fwd's warp samples sat at its closing barrier).  Here a row's contraction is cut into 16 k-slices held by one
// half-warp and summed with shuffles; the weights (<= 64 KB) are L1 hits.
__device__ void tile_gemm_skinny(const float* in, float* out, const float* __restrict__ Wt, int K, int N) {
  const int t = threadIdx.x;
  const int r = t >> 4, ks = t & 15;   // 32 rows x 16 k-slices
  for (int nb = 0; nb < N; nb += 8) {
    const int nn = N - nb < 8 ? N - nb : 8;
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
    for (int k = ks; k < K; k += 16) {
      const float a = in[k * kLd + r];
      const float* w = Wt + static_cast<size_t>(k) * N + nb;
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (i < nn) acc[i] = fmaf(a, __ldg(w + i), acc[i]);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float v = acc[i];
      v += __shfl_xor_sync(0xffffffffu, v, 8);
      v += __shfl_xor_sync(0xffffffffu, v, 4);
      v += __shfl_xor_sync(0xffffffffu, v, 2);
      v += __shfl_xor_sync(0xffffffffu, v, 1);
      if (ks == 0 && i < nn) out[(nb + i) * kLd + r] = v;
    }
  }
  __syncthreads();
}

// out[n*kLd + r] = sum_k in[k*kLd + r] * Wt[k*N + n]: staged path for the big square layers, skinny path for few
// outputs behind a long contraction, direct path otherwise
__device__ void tile_gemm(const float* in, float* out, const float* __restrict__ Wt, int K, int N, float* wbuf) {
  if (N <= 32 && K >= 64) { tile_gemm_skinny(in, out, Wt, K, N); return; }
  if ((K % kSlabK) == 0 && K >= 64 && (reinterpret_cast<uintptr_t>(Wt) & 15) == 0) {
    if (N == 512) { tile_gemm_staged<8>(in, out, Wt, K, wbuf); return; }
    if (N == 256) { tile_gemm_staged<4>(in, out, Wt, K, wbuf); return; }
  }
  tile_gemm_direct(in, out, Wt, K, N);
}


__global__ void k_transpose_l(const float* __restrict__ W, float* __restrict__ Wt, int rows, int cols) {
  __shared__ float tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (r < rows && c < cols) ? W[static_cast<size_t>(r) * cols + c] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (r < rows && c < cols) Wt[static_cast<size_t>(c) * rows + r] = tile[threadIdx.x][i];
  }
}


inline size_t align_up(size_t v) { return (v + 255) & ~size_t(255); }

}  // namespace
}  // namespace dmip
