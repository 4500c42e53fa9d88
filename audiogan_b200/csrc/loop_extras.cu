// Kernels of the reference-faithful training step beyond the core step:
//   REINFORCE gradient of the stop head            audiogan.py:873-908
//   calc_dists feature-matching statistics          audiogan.py:336-359, 848-855 (further down)
#include "common.cuh"

namespace ag {

// ---------------------------------------------------------------- REINFORCE (audiogan.py:873-908)
// reward_b = -loss_ps[b]; baseline' = mean(reward) (first call) or 0.5 baseline + 0.5 mean(reward) (:874-875);
// the stochastic stop node is a 2-way multinomial over (1 - sigma(s), sigma(s)) (:445-450): the score-function gradient of
// -sum_{b,t} (reward_b - baseline') * [t < glen_b] * log p(stop_bt) with respect to the stop logit is
//   out[b,t] = -(reward_b - baseline') * (stop_bt - sigma(s_bt))   inside the mask, 0 outside.
// One block per sample; every block recomputes the batch mean (B values) so that no second launch is needed.
__global__ void __launch_bounds__(128) reinforce_dlogit_kernel(const float* __restrict__ s, int64_t s_ld,
                                                               const int32_t* __restrict__ stop, int64_t stop_ld,
                                                               const float* __restrict__ loss_ps, const int32_t* __restrict__ glen,
                                                               const float* __restrict__ baseline_in, float* __restrict__ baseline_out,
                                                               float* __restrict__ out, int64_t out_ld, int64_t B, int64_t T) {
  __shared__ float red[32];
  const int64_t b = blockIdx.x;
  float acc = 0.f;
  for (int64_t i = threadIdx.x; i < B; i += blockDim.x) acc -= loss_ps[i];
  const float rmean = block_sum(acc, red) / (float)B;
  const float base = baseline_in ? 0.5f * baseline_in[0] + 0.5f * rmean : rmean;
  if (b == 0 && threadIdx.x == 0 && baseline_out) baseline_out[0] = base;
  const float adv = -loss_ps[b] - base;
  const int n = glen[b];
  for (int64_t t = threadIdx.x; t < T; t += blockDim.x) {
    float v = 0.f;
    if (t < n) v = -adv * ((stop[b * stop_ld + t] != 0 ? 1.f : 0.f) - sigmoidf_(s[b * s_ld + t]));
    out[b * out_ld + t] = v;
  }
}

}  // namespace ag

using namespace ag;
extern "C" {

int ag_reinforce_dlogit(const float* s, int64_t s_ld, const int32_t* stop, int64_t stop_ld, const float* loss_ps,
                        const int32_t* glen, const float* baseline_in, float* baseline_out, float* out, int64_t out_ld,
                        int64_t B, int64_t T, void* stream) {
  AG_CHECK_ARG(s && stop && loss_ps && glen && out && B > 0 && T > 0 && s_ld >= T && stop_ld >= T && out_ld >= T,
               "ag_reinforce_dlogit: bad args");
  AG_CHECK_ARG(baseline_in != baseline_out || !baseline_in, "ag_reinforce_dlogit: baseline_in and baseline_out must not alias");
  reinforce_dlogit_kernel<<<(unsigned)B, 128, 0, (cudaStream_t)stream>>>(s, s_ld, stop, stop_ld, loss_ps, glen, baseline_in,
                                                                         baseline_out, out, out_ld, B, T);
  AG_LAUNCH_CHECK();
  return AG_OK;
}

}
