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
__device__ void tile_gemm(const float* in, float* out, const float* __restrict__ Wt, int K, int N) {
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
