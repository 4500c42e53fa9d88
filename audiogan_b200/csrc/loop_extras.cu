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


// ---------------------------------------------------------------- calc_dists time moments (audiogan.py:341-348)
// For every (sample b, channel c) of a channel-last activation h[b, t, c] (fp32 or bf16 storage, rows past len[b] hold zeros):
//   m = sum_t h / l        s = sqrt(sum_t (h - m)^2) / l        f = (sum_t (h - m)^4)^(1/4) / l          (t < l = len[b])
// HBM-bound streams over the activation.  Two passes, as the reference does it (the centred sums are NOT derived from raw power
// sums: that cancels in fp32): pass 1 accumulates S1, pass 2 the centred sums Q2, Q3, Q4 (Q3 is what the backward needs).
// Block = 256 threads = CW channels x (256 / CW) row lanes over a chunk of TCH rows; partial sums leave through one
// atomicAdd per (block, channel).  Grid (time chunks, channel tiles, B).
template <int PASS>
__global__ void __launch_bounds__(256) time_moments_kernel(const void* __restrict__ h, int dtype, int64_t h_bs, int64_t h_rs,
                                                           const int32_t* __restrict__ len, int C, int CW, int TCH,
                                                           float* __restrict__ S1, float* __restrict__ Q) {
  __shared__ float red[3][256];
  const int b = blockIdx.z;
  const int l = len[b];
  const int t0 = blockIdx.x * TCH;
  if (t0 >= l) return;
  const int cl = threadIdx.x % CW, rl = threadIdx.x / CW, RL = 256 / CW;
  const int c = blockIdx.y * CW + cl;
  const int t1 = min(l, t0 + TCH);
  const char* base = reinterpret_cast<const char*>(h);
  const int64_t es = dtype ? 2 : 4;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f;
  float m = 0.f;
  if (PASS == 2 && c < C) m = S1[(int64_t)b * C + c] / (float)l;
  if (c < C) {
    for (int t = t0 + rl; t < t1; t += RL) {
      const float v = ld_any(base + ((int64_t)b * h_bs + (int64_t)t * h_rs) * es, c, dtype);
      if (PASS == 1) {
        a0 += v;
      } else {
        const float d = v - m, d2 = d * d;
        a0 += d2; a1 += d2 * d; a2 += d2 * d2;
      }
    }
  }
  red[0][threadIdx.x] = a0;
  if (PASS == 2) { red[1][threadIdx.x] = a1; red[2][threadIdx.x] = a2; }
  __syncthreads();
  if (rl == 0 && c < C) {
    for (int r = 1; r < RL; ++r) {
      a0 += red[0][r * CW + cl];
      if (PASS == 2) { a1 += red[1][r * CW + cl]; a2 += red[2][r * CW + cl]; }
    }
    const int64_t o = (int64_t)b * C + c;
    if (PASS == 1) {
      atomicAdd(&S1[o], a0);
    } else {
      const int64_t BC = (int64_t)gridDim.z * C;
      atomicAdd(&Q[o], a0); atomicAdd(&Q[BC + o], a1); atomicAdd(&Q[2 * BC + o], a2);
    }
  }
}

// backward: dh[b,t,c] = gm/l + gs * (h-m) / (l sqrt(Q2)) + gf * Q4^(-3/4) / l * ((h-m)^3 - Q3/l)   for t < l, 0 elsewhere
// (sum_t (h - m) = 0 removes the dependence of Q2 on m; Q4 depends on m through -4 Q3 dm).
__global__ void __launch_bounds__(256) time_moments_bwd_kernel(const void* __restrict__ h, int dtype, int64_t h_bs, int64_t h_rs,
                                                               const int32_t* __restrict__ len, int C, int T,
                                                               const float* __restrict__ S1, const float* __restrict__ Q,
                                                               const float* __restrict__ gm, const float* __restrict__ gs,
                                                               const float* __restrict__ gf, void* __restrict__ dh,
                                                               int dh_dtype, int64_t BC) {
  const int b = blockIdx.z;
  const int l = len[b];
  const int64_t n = (int64_t)T * C;
  const float fl = (float)l;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int t = (int)(i / C), c = (int)(i % C);
    float out = 0.f;
    if (t < l) {
      const int64_t o = (int64_t)b * C + c;
      const float m = S1[o] / fl, q2 = Q[o], q3 = Q[BC + o], q4 = Q[2 * BC + o];
      const float v = ld_any(reinterpret_cast<const char*>(h) + ((int64_t)b * h_bs + (int64_t)t * h_rs) * (dtype ? 2 : 4), c, dtype);
      const float d = v - m;
      out = gm[o] / fl;
      if (q2 > 0.f) out += gs[o] * d / (fl * sqrtf(q2));
      if (q4 > 0.f) out += gf[o] * (d * d * d - q3 / fl) / (fl * powf(q4, 0.75f));
    }
    st_any(dh, (int64_t)b * n + i, out, dh_dtype);
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

static int moments_cw(int64_t C) { int cw = 1; while (cw < C && cw < 64) cw <<= 1; return cw; }

int ag_time_moments_fwd(const void* h, int32_t dtype, int64_t h_bs, int64_t h_rs, const int32_t* len, int64_t B, int64_t T,
                        int64_t C, float* S1, float* Q, void* stream) {
  AG_CHECK_ARG(h && len && S1 && Q && B > 0 && T > 0 && C > 0 && B <= 65535, "ag_time_moments_fwd: bad args");
  const int CW = moments_cw(C), TCH = 256;
  dim3 grid((unsigned)((T + TCH - 1) / TCH), (unsigned)((C + CW - 1) / CW), (unsigned)B);
  AG_CUDA(cudaMemsetAsync(S1, 0, sizeof(float) * B * C, (cudaStream_t)stream));
  AG_CUDA(cudaMemsetAsync(Q, 0, sizeof(float) * 3 * B * C, (cudaStream_t)stream));
  time_moments_kernel<1><<<grid, 256, 0, (cudaStream_t)stream>>>(h, dtype, h_bs, h_rs, len, (int)C, CW, TCH, S1, Q);
  time_moments_kernel<2><<<grid, 256, 0, (cudaStream_t)stream>>>(h, dtype, h_bs, h_rs, len, (int)C, CW, TCH, S1, Q);
  AG_LAUNCH_CHECK();
  return AG_OK;
}

int ag_time_moments_bwd(const void* h, int32_t dtype, int64_t h_bs, int64_t h_rs, const int32_t* len, int64_t B, int64_t T,
                        int64_t C, const float* S1, const float* Q, const float* gm, const float* gs, const float* gf,
                        void* dh, int32_t dh_dtype, void* stream) {
  AG_CHECK_ARG(h && len && S1 && Q && gm && gs && gf && dh && B > 0 && T > 0 && C > 0 && B <= 65535,
               "ag_time_moments_bwd: bad args");
  const int64_t n = T * C;
  dim3 grid((unsigned)((n + 256 * 8 - 1) / (256 * 8)), 1, (unsigned)B);
  time_moments_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(h, dtype, h_bs, h_rs, len, (int)C, (int)T, S1, Q, gm, gs, gf, dh,
                                                                  dh_dtype, B * C);
  AG_LAUNCH_CHECK();
  return AG_OK;
}

}
