// dmip_pack.cu — re-tile an nn.Linear MLP ([in] -> 512 -> 512 -> 512 -> [out]) into the operand images the
// tcgen05 kernels stream: 16-bit (layer 0 bf16, layers 1-3 f16), K-major, 128-byte swizzle, one 16 KB stage = 128 out-features x 64 k, stored in
// the exact order the MMA warp consumes them (layer 0: chunk-major, then layers 1, 2, output layer), followed by
// an fp32 tail (biases and the row-constant columns of W0).  Source layout: state_dict keys 0/3/5/7.
#include "dmip_common.h"
#include "dmip_ptx.cuh"

namespace dmip {

int tc_net_geom(const DmipMlp* net, int n_varying, int out_rows, int split, TcNetGeom* g) {
  DMIP_REQUIRE(net != nullptr, "net is NULL");
  DMIP_REQUIRE(net->n_layers == 4 && net->width[0] == 512 && net->width[1] == 512 && net->width[2] == 512,
               "tcgen05 path needs hidden_layers == [512,512,512] (got %d layers, widths %d,%d,%d); use DMIP_PREC_F32",
               net->n_layers, net->width[0], net->width[1], net->width[2]);
  DMIP_REQUIRE(split >= 1 && split <= 4, "l0_split must be 1, 2, 3 or 4 (got %d)", split);
  g->l0_f16 = split == 4 ? 1 : 0;
  if (split == 4) split = 1;
  DMIP_REQUIRE(n_varying >= 1 && n_varying <= net->in_dim, "n_varying %d out of range (in_dim %d)", n_varying,
               net->in_dim);
  DMIP_REQUIRE(out_rows >= 1 && out_rows <= net->out_dim && out_rows <= 128,
               "out_rows %d out of range (out_dim %d, max 128)", out_rows, net->out_dim);
  g->n_varying = n_varying;
  g->split = split;
  g->dvp = round_up(n_varying, 8);
  g->k0 = (split - 1) * g->dvp + n_varying;
  g->k0pad = round_up(g->k0, 16);
  DMIP_REQUIRE(g->k0pad <= 512, "layer-0 GEMM depth %d exceeds 512", g->k0pad);
  g->kb0 = ceil_div(g->k0pad, 64);
  g->out_rows = out_rows;
  g->outpad = round_up(out_rows, 16);
  g->n_stages = 4 * g->kb0 + 32 + 32 + 8;
  g->n_const = net->in_dim - n_varying;
  return DMIP_OK;
}

namespace {

struct PackParams {
  const float* W[4];
  const float* b[4];
  int in_dim, dv, dvp, split, l0_f16, k0, kb0, out_rows, n_const, n_stages;
  uint8_t* stages;
  float* tail;
};

__global__ void k_pack(const PackParams p) {
  const int st = blockIdx.x;
  if (st < p.n_stages) {
    int l, c, kb;
    if (st < 4 * p.kb0) {
      l = 0; c = st / p.kb0; kb = st % p.kb0;
    } else {
      const int s2 = st - 4 * p.kb0;
      if (s2 < 32) { l = 1; c = s2 >> 3; kb = s2 & 7; }
      else if (s2 < 64) { l = 2; c = (s2 - 32) >> 3; kb = s2 & 7; }
      else { l = 3; c = 0; kb = s2 - 64; }
    }
    uint8_t* img = p.stages + static_cast<size_t>(st) * 16384;
    for (int i = threadIdx.x; i < 128 * 64; i += blockDim.x) {
      const int r = i >> 6, k = i & 63;
      const int n = c * 128 + r;
      const int kg = kb * 64 + k;
      float v = 0.f;
      if (l == 0) {
        const int part = kg / p.dvp, idx = kg - part * p.dvp;
        if (part < p.split && idx < p.dv) {
          // operand parts [x_hi | x_lo | x_hi] meet weight parts [W_hi | W_hi | W_lo]
          const float w = p.W[0][static_cast<size_t>(n) * p.in_dim + idx];
          const float hi = p.l0_f16 ? w : bf16_round(w);
          v = (part == 2) ? (w - hi) : hi;
        }
      } else if (l < 3) {
        v = p.W[l][static_cast<size_t>(n) * 512 + kg];
      } else if (r < p.out_rows) {
        v = p.W[3][static_cast<size_t>(r) * 512 + kg];
      }
#ifndef DMIP_H_F16
      // l0_split = 4: layer 0 is an f16 x f16 product (kind::f16 wants one format for both operands), |w| clamped to the f16 range
      const unsigned short h = static_cast<unsigned short>(
          ((l == 0 && p.l0_f16) ? pack_f16x2(fminf(fmaxf(v, -65504.f), 65504.f), 0.f) : pack_bf16x2(v, 0.f)) & 0xFFFFu);
#else
      // layers 1-3 multiply f16 activations (tanh values): their weights are f16 too (kind::f16 wants one format for
      // both operands; 11 mantissa bits instead of 8, |w| clamped to the f16 range); layer 0 meets the bf16 hi/lo state
      const unsigned short h = static_cast<unsigned short>(
          (l == 0 ? pack_bf16x2(v, 0.f) : pack_f16x2(fminf(fmaxf(v, -65504.f), 65504.f), 0.f)) & 0xFFFFu);
#endif
      *reinterpret_cast<unsigned short*>(img + sw128_offset(r, k, 16384)) = h;
    }
  } else {
    // fp32 tail: b0 b1 b2 [512 each], b3 [128], W0const [512][n_const]
    for (int i = threadIdx.x; i < 1536; i += blockDim.x) p.tail[i] = p.b[i >> 9][i & 511];
    for (int i = threadIdx.x; i < 128; i += blockDim.x) p.tail[1536 + i] = i < p.out_rows ? p.b[3][i] : 0.f;
    for (int i = threadIdx.x; i < 512 * p.n_const; i += blockDim.x) {
      const int n = i / p.n_const, j = i - n * p.n_const;
      p.tail[1664 + i] = p.W[0][static_cast<size_t>(n) * p.in_dim + p.dv + j];
    }
  }
}

}  // namespace

int launch_pack(const DmipMlp* net, const TcNetGeom& g, void* packed, cudaStream_t s) {
  PackParams p;
  for (int l = 0; l < 4; ++l) {
    DMIP_REQUIRE(net->W[l] && net->b[l], "net layer %d has a NULL pointer", l);
    p.W[l] = net->W[l];
    p.b[l] = net->b[l];
  }
  p.in_dim = net->in_dim;
  p.dv = g.n_varying;
  p.dvp = g.dvp;
  p.split = g.split;
  p.l0_f16 = g.l0_f16;
  p.k0 = g.k0;
  p.kb0 = g.kb0;
  p.out_rows = g.out_rows;
  p.n_const = g.n_const;
  p.n_stages = g.n_stages;
  p.stages = static_cast<uint8_t*>(packed);
  p.tail = reinterpret_cast<float*>(p.stages + g.stage_bytes());
  k_pack<<<g.n_stages + 1, 256, 0, s>>>(p);
  DMIP_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return DMIP_OK;
}

}  // namespace dmip
