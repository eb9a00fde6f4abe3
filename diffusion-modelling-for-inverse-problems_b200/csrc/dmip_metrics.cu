// dmip_metrics.cu — evaluation metrics next to the sampler: the 75-bin d-dimensional sample histogram
// and the histogram KL divergence the reference reports (main_diffusion_linear.py:84-117,
// main_diffusion_scatterometry.py:72-101), computed where the samples already are instead of after a device->host copy
// of 30,000 x repeats x observations samples and an np.histogramdd call per repeat.
//
// Binning reproduces np.histogramdd exactly: per dimension the bin of x is searchsorted(edges, x, side='right') - 1 on
// the caller's edges (np.linspace(lo, hi, bins + 1)), a sample equal to the last edge belongs to the last bin, samples
// outside the range are dropped.  Integer counts, so the result is bit-exact and order independent.
#include "dmip_common.h"

namespace dmip {

namespace {

constexpr int kMaxHistDim = 4;

struct HistParams {
  const float* x;      // (n, dim)
  long long n;
  int dim;
  int bins[kMaxHistDim];
  const double* edges[kMaxHistDim];   // bins[d] + 1 ascending edges
  unsigned long long* counts;         // prod(bins), row-major (C order), accumulated into
};

__global__ void __launch_bounds__(256) k_histogramdd(const __grid_constant__ HistParams P) {
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < P.n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    long long flat = 0;
    bool inside = true;
    for (int d = 0; d < P.dim; ++d) {
      const double v = static_cast<double>(P.x[i * P.dim + d]);
      const double* e = P.edges[d];
      const int nb = P.bins[d];
      // searchsorted(e, v, side='right'): number of edges <= v
      int lo = 0, hi = nb + 1;
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (e[mid] <= v) lo = mid + 1; else hi = mid;
      }
      int b = lo - 1;
      if (v == e[nb]) b = nb - 1;          // the right-most edge is inclusive
      if (!(b >= 0 && b < nb)) inside = false;   // also drops NaN
      flat = flat * nb + b;
    }
    if (inside) atomicAdd(&P.counts[flat], 1ull);
  }
}

// KL(p || q) = sum rel_entr(p, q) of two count histograms normalised as the reference does:
// h / sum(h), + epsilon, / sum again.  One block; out[0] = KL, out[1] = sum(p counts), out[2] = sum(q counts).
__global__ void __launch_bounds__(1024) k_hist_kl(const unsigned long long* __restrict__ hp, const unsigned long long* __restrict__ hq,
                                                   long long m, double epsilon, double* out) {
  __shared__ double red[1024];
  __shared__ double tot[2];
  const int t = threadIdx.x;
  double sp = 0.0, sq = 0.0;
  for (long long i = t; i < m; i += 1024) { sp += static_cast<double>(hp[i]); sq += static_cast<double>(hq[i]); }
  auto block_sum = [&](double v) {
    red[t] = v;
    __syncthreads();
    for (int s = 512; s > 0; s >>= 1) {
      if (t < s) red[t] += red[t + s];
      __syncthreads();
    }
    const double r = red[0];
    __syncthreads();
    return r;
  };
  const double Sp = block_sum(sp), Sq = block_sum(sq);
  if (t == 0) { tot[0] = Sp; tot[1] = Sq; }
  // after "+ epsilon" each histogram sums to 1 + m * epsilon
  const double zp = 1.0 + static_cast<double>(m) * epsilon, zq = zp;
  double kl = 0.0;
  for (long long i = t; i < m; i += 1024) {
    const double p = (static_cast<double>(hp[i]) / Sp + epsilon) / zp;
    const double q = (static_cast<double>(hq[i]) / Sq + epsilon) / zq;
    if (p > 0.0) kl += p * log(p / q);     // scipy.special.rel_entr
  }
  const double K = block_sum(kl);
  if (t == 0) { out[0] = K; out[1] = Sp; out[2] = Sq; }
}

}  // namespace

int launch_histogramdd(const DmipHistogram* d, cudaStream_t s) {
  DMIP_REQUIRE(d->dim >= 1 && d->dim <= kMaxHistDim, "histogram dimension must be 1..%d", kMaxHistDim);
  DMIP_REQUIRE(d->n >= 0, "negative sample count");
  long long m = 1;
  HistParams P = {};
  for (int k = 0; k < d->dim; ++k) {
    DMIP_REQUIRE(d->bins[k] >= 1 && d->edges[k] != nullptr, "dimension %d: bins must be >= 1 and edges non-NULL", k);
    P.bins[k] = d->bins[k];
    P.edges[k] = d->edges[k];
    m *= d->bins[k];
  }
  DMIP_REQUIRE(d->counts != nullptr, "counts is NULL");
  if (d->n == 0) return DMIP_OK;
  DMIP_REQUIRE(d->x != nullptr, "x is NULL");
  P.x = d->x; P.n = d->n; P.dim = d->dim;
  P.counts = reinterpret_cast<unsigned long long*>(d->counts);
  const long long want = (d->n + 255) / 256;
  const unsigned grid = static_cast<unsigned>(want < 148 * 16 ? want : 148 * 16);
  k_histogramdd<<<grid, 256, 0, s>>>(P);
  DMIP_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return DMIP_OK;
}

int launch_hist_kl(const void* hp, const void* hq, long long m, double epsilon, double* out, cudaStream_t s) {
  DMIP_REQUIRE(hp && hq && out && m >= 1, "histogram KL: NULL pointer or empty histogram");
  k_hist_kl<<<1, 1024, 0, s>>>(static_cast<const unsigned long long*>(hp), static_cast<const unsigned long long*>(hq), m,
                               epsilon, out);
  DMIP_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return DMIP_OK;
}

}  // namespace dmip
