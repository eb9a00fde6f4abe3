// dmip_tsample.cu — the training-time draw of t on the device: BaseClassDiffusionModel.sample_t
// (models/diffusion.py:48-58) over VariancePreservingSDE.sample_debiasing_t (sdes.py:51-57 -> sdeflow-light's
// sample_vp_truncated_q, restated in dmip/sdes.py vp_truncated_inverse_cdf; SURVEY.md App. A.2: parity unpinned, the
// distribution is checked against its analytic CDF).  The reference draws u on the CPU and copies t to the GPU every batch;
// at its own batch size (1000) that copy and the ten small CPU ops are a third of a training step's host time.
#include <math.h>

#include "dmip_common.h"

namespace dmip {

namespace {

struct TSample {
  const float* u;   // uniforms in [0, 1)
  float* t;
  long long n;
  int debias;
  double bmin, db, T, eps_add, r_eps, a_eps, Z, t_eps;
};

// double precision: n is a batch size, and the host mirror evaluates the constants in double as well
__global__ void __launch_bounds__(256) k_sample_t(const TSample P) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= P.n) return;
  const double u = static_cast<double>(P.u[i]);
  double t;
  if (P.debias) {
    // inverse CDF of q(t) ~ beta(t) / var(t), constant below t_eps
    if (u <= P.t_eps * P.r_eps / P.Z) {
      t = P.Z / P.r_eps * u;
    } else {
      const double e = P.Z * u + P.a_eps - P.r_eps * P.t_eps;
      const double l1p = e > 30.0 ? e + log1p(exp(-e)) : log1p(exp(e));
      t = (-P.bmin + sqrt(P.bmin * P.bmin + 2.0 * P.db * l1p)) / P.db;
    }
    t += P.eps_add;                                  // models/diffusion.py:51-53
    if (static_cast<float>(t) > static_cast<float>(P.T)) t -= P.eps_add;
  } else {
    t = P.eps_add + u * P.T;                         // models/diffusion.py:55-56
    if (static_cast<float>(t) > static_cast<float>(P.T)) t = P.T - P.eps_add;
  }
  P.t[i] = static_cast<float>(t);
}

}  // namespace

int launch_sample_t(const float* u, float* t, long long n, int debias, float beta_min, float beta_max, float t_epsilon,
                    float T, float eps_add, cudaStream_t s) {
  DMIP_REQUIRE(n >= 0, "negative count");
  if (n == 0) return DMIP_OK;
  DMIP_REQUIRE(u != nullptr && t != nullptr, "u / t is NULL");
  DMIP_REQUIRE(beta_max > beta_min && beta_min > 0.f && t_epsilon > 0.f && T > t_epsilon, "bad VP-SDE schedule");
  TSample P = {};
  P.u = u; P.t = t; P.n = n; P.debias = debias;
  P.bmin = beta_min; P.db = static_cast<double>(beta_max) - beta_min; P.T = T; P.eps_add = eps_add; P.t_eps = t_epsilon;
  auto big_b = [&](double x) { return 0.5 * x * x * P.db + x * P.bmin; };
  auto antider = [&](double x) { const double b = big_b(x); return log(1.0 - exp(-b)) + b; };
  P.r_eps = (P.bmin + P.db * P.t_eps) / (1.0 - exp(-big_b(P.t_eps)));
  P.a_eps = antider(P.t_eps);
  P.Z = P.r_eps * P.t_eps + antider(P.T) - P.a_eps;
  k_sample_t<<<static_cast<unsigned>((n + 255) / 256), 256, 0, s>>>(P);
  DMIP_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return DMIP_OK;
}

}  // namespace dmip
