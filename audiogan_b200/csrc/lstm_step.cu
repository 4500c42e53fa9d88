// Step-wise generator recurrence for hidden sizes whose weights cannot stay resident on chip (audiogan.py:428-460 with
// --gstatesize 2048: [whh | wx] is 36.8 MB in bf16, more than the GPU's shared memory).  The gate product of one frame is then
// a small-M tensor-core GEMM over the bf16 weights (they stay in the 126 MB L2 between frames) issued per frame by the host side
// (engine._gen_stepwise_*: ag_gemm_nt_tc per frame, all inside the step's CUDA graph); the kernels here are the point-wise parts
// between those GEMMs -- LSTM cell forward / backward, projection finish (tanh, stop logit), dpx -- each writing the operand
// row block of the NEXT GEMM in bf16 beside the saved state.  HBM/L2-bound streams over [B, H] rows.
#include "common.cuh"

namespace ag {

__device__ __forceinline__ float sigm(float x) { return 1.f / (1.f + expf(-x)); }

// gates = act(gpre + pre_t); c_t = f c_{t-1} + i g; h_t = o tanh(c_t).   One thread per (sample, unit).
__global__ void __launch_bounds__(256) step_cell_fwd_kernel(const float* __restrict__ gpre, const float* __restrict__ pre, int64_t pre_bs,
                                                            const float* __restrict__ cprev, int64_t c_bs, float* __restrict__ gates,
                                                            int64_t g_bs, float* __restrict__ cout, float* __restrict__ h32, int64_t h_bs,
                                                            __nv_bfloat16* __restrict__ h16, __nv_bfloat16* __restrict__ hx, int64_t hx_ld,
                                                            int B, int H) {
  const int64_t n = (int64_t)B * H;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int b = (int)(i / H), j = (int)(i - (int64_t)b * H);
    const float* gp = gpre + (int64_t)b * 4 * H + j;
    const float* pp = pre + (int64_t)b * pre_bs + j;
    const float gi = sigm(gp[0] + pp[0]), gf = sigm(gp[H] + pp[H]), gg = tanhf(gp[2 * H] + pp[2 * H]), go = sigm(gp[3 * H] + pp[3 * H]);
    const float cp = cprev ? cprev[(int64_t)b * c_bs + j] : 0.f;
    const float c = gf * cp + gi * gg, h = go * tanhf(c);
    if (gates) {
      float* gq = gates + (int64_t)b * g_bs + j;
      gq[0] = gi; gq[H] = gf; gq[2 * H] = gg; gq[3 * H] = go;
    }
    cout[(int64_t)b * c_bs + j] = c;
    if (h32) h32[(int64_t)b * h_bs + j] = h;
    const __nv_bfloat16 hb = __float2bfloat16(h);
    if (h16) h16[(int64_t)b * h_bs + j] = hb;
    hx[(int64_t)b * hx_ld + j] = hb;
  }
}

// px [B, FP] = [wp ; ws] h_t + b2 (from the GEMM): x_t = tanh(px[:, :F]) -> xbuf (fp32 + bf16) and the x part of the next
// gate GEMM's operand row; stop logit px[:, F] -> sbuf.
__global__ void __launch_bounds__(256) step_proj_finish_kernel(const float* __restrict__ px, int FP, float* __restrict__ x32,
                                                               __nv_bfloat16* __restrict__ x16, int64_t x_bs, __nv_bfloat16* __restrict__ hx,
                                                               int64_t hx_ld, int hx_off, float* __restrict__ sbuf, int64_t s_bs, int B, int F) {
  const int64_t n = (int64_t)B * (F + 1);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int b = (int)(i / (F + 1)), p = (int)(i - (int64_t)b * (F + 1));
    const float v = px[(int64_t)b * FP + p];
    if (p < F) {
      const float xv = tanhf(v);
      x32[(int64_t)b * x_bs + p] = xv;
      const __nv_bfloat16 xb = __float2bfloat16(xv);
      if (x16) x16[(int64_t)b * x_bs + p] = xb;
      hx[(int64_t)b * hx_ld + hx_off + p] = xb;
    } else if (sbuf) {
      sbuf[(int64_t)b * s_bs] = v;
    }
  }
}

// dpx_t[:, p] = (dx_ext_t + dxpre)(1 - x_t^2) for p < F, dpx_t[:, F] = ds_ext_t, pad columns 0; also into the dpx part of the
// operand row [dgates_{t+1} | dpx_t] of the dh GEMM
__global__ void __launch_bounds__(256) step_dpx_kernel(const float* __restrict__ dxpre, int64_t dxpre_ld, const float* __restrict__ dx_ext,
                                                       int64_t dx_bs, const float* __restrict__ ds_ext, int64_t ds_bs,
                                                       const float* __restrict__ xt, int64_t x_bs, float* __restrict__ dpx,
                                                       __nv_bfloat16* __restrict__ dpx16, int64_t dpx_bs, __nv_bfloat16* __restrict__ dgp,
                                                       int64_t dgp_ld, int dgp_off, int B, int F, int FP) {
  const int64_t n = (int64_t)B * FP;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int b = (int)(i / FP), p = (int)(i - (int64_t)b * FP);
    float v = 0.f;
    if (p < F) {
      const float x = xt[(int64_t)b * x_bs + p];
      v = ((dx_ext ? dx_ext[(int64_t)b * dx_bs + p] : 0.f) + (dxpre ? dxpre[(int64_t)b * dxpre_ld + p] : 0.f)) * (1.f - x * x);
    } else if (p == F) {
      v = ds_ext ? ds_ext[(int64_t)b * ds_bs] : 0.f;
    }
    dpx[(int64_t)b * dpx_bs + p] = v;
    const __nv_bfloat16 vb = __float2bfloat16(v);
    if (dpx16) dpx16[(int64_t)b * dpx_bs + p] = vb;
    dgp[(int64_t)b * dgp_ld + dgp_off + p] = vb;
  }
}

// cell backward of frame t: dh_t (from the GEMM) -> dgates_t, dc carried in `dc` [B, H]
__global__ void __launch_bounds__(256) step_cell_bwd_kernel(const float* __restrict__ dh, const float* __restrict__ gates, int64_t g_bs,
                                                            const float* __restrict__ c, const float* __restrict__ cprev, int64_t c_bs,
                                                            float* __restrict__ dc, float* __restrict__ dg32, __nv_bfloat16* __restrict__ dg16,
                                                            int64_t dg_bs, __nv_bfloat16* __restrict__ dgp, int64_t dgp_ld, int B, int H) {
  const int64_t n = (int64_t)B * H;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int b = (int)(i / H), j = (int)(i - (int64_t)b * H);
    const float* gq = gates + (int64_t)b * g_bs + j;
    const float gi = gq[0], gf = gq[H], gg = gq[2 * H], go = gq[3 * H];
    const float cc = c[(int64_t)b * c_bs + j], cp = cprev ? cprev[(int64_t)b * c_bs + j] : 0.f;
    const float dht = dh[(int64_t)b * H + j];
    const float tch = tanhf(cc);
    const float dcv = dc[i] + dht * go * (1.f - tch * tch);
    dc[i] = dcv * gf;
    const float o0 = dcv * gg * gi * (1.f - gi), o1 = dcv * cp * gf * (1.f - gf), o2 = dcv * gi * (1.f - gg * gg), o3 = dht * tch * go * (1.f - go);
    const float o[4] = {o0, o1, o2, o3};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int64_t off = (int64_t)b * dg_bs + (int64_t)q * H + j;
      if (dg32) dg32[off] = o[q];
      const __nv_bfloat16 ob = __float2bfloat16(o[q]);
      if (dg16) dg16[off] = ob;
      dgp[(int64_t)b * dgp_ld + (int64_t)q * H + j] = ob;
    }
  }
}

static unsigned step_grid(int64_t n) {
  int64_t g = (n + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 8;
  return (unsigned)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace ag

using namespace ag;
extern "C" {

int ag_lstm_step_cell_fwd(const float* gpre, const float* pre, int64_t pre_bs, const float* cprev, int64_t c_bs, float* gates, int64_t g_bs,
                          float* cout, float* h32, int64_t h_bs, void* h16, void* hx, int64_t hx_ld, int32_t B, int32_t H, void* stream) {
  AG_CHECK_ARG(gpre && pre && cout && hx && B > 0 && H > 0, "ag_lstm_step_cell_fwd: bad args");
  step_cell_fwd_kernel<<<step_grid((int64_t)B * H), 256, 0, (cudaStream_t)stream>>>(gpre, pre, pre_bs, cprev, c_bs, gates, g_bs, cout, h32, h_bs,
                                                                                    (__nv_bfloat16*)h16, (__nv_bfloat16*)hx, hx_ld, B, H);
  AG_LAUNCH_CHECK();
  return AG_OK;
}

int ag_gen_step_proj_finish(const float* px, int32_t FP, float* x32, void* x16, int64_t x_bs, void* hx, int64_t hx_ld, int32_t hx_off,
                            float* sbuf, int64_t s_bs, int32_t B, int32_t F, void* stream) {
  AG_CHECK_ARG(px && x32 && hx && B > 0 && F > 0 && FP > F, "ag_gen_step_proj_finish: bad args");
  step_proj_finish_kernel<<<step_grid((int64_t)B * (F + 1)), 256, 0, (cudaStream_t)stream>>>(px, FP, x32, (__nv_bfloat16*)x16, x_bs,
                                                                                             (__nv_bfloat16*)hx, hx_ld, hx_off, sbuf, s_bs, B, F);
  AG_LAUNCH_CHECK();
  return AG_OK;
}

int ag_gen_step_dpx(const float* dxpre, int64_t dxpre_ld, const float* dx_ext, int64_t dx_bs, const float* ds_ext, int64_t ds_bs, const float* xt,
                    int64_t x_bs, float* dpx, void* dpx16, int64_t dpx_bs, void* dgp, int64_t dgp_ld, int32_t dgp_off, int32_t B, int32_t F,
                    int32_t FP, void* stream) {
  AG_CHECK_ARG(xt && dpx && dgp && B > 0 && F > 0 && FP > F, "ag_gen_step_dpx: bad args");
  step_dpx_kernel<<<step_grid((int64_t)B * FP), 256, 0, (cudaStream_t)stream>>>(dxpre, dxpre_ld, dx_ext, dx_bs, ds_ext, ds_bs, xt, x_bs, dpx,
                                                                                (__nv_bfloat16*)dpx16, dpx_bs, (__nv_bfloat16*)dgp, dgp_ld, dgp_off,
                                                                                B, F, FP);
  AG_LAUNCH_CHECK();
  return AG_OK;
}

int ag_lstm_step_cell_bwd(const float* dh, const float* gates, int64_t g_bs, const float* c, const float* cprev, int64_t c_bs, float* dc,
                          float* dg32, void* dg16, int64_t dg_bs, void* dgp, int64_t dgp_ld, int32_t B, int32_t H, void* stream) {
  AG_CHECK_ARG(dh && gates && c && dc && dgp && B > 0 && H > 0, "ag_lstm_step_cell_bwd: bad args");
  step_cell_bwd_kernel<<<step_grid((int64_t)B * H), 256, 0, (cudaStream_t)stream>>>(dh, gates, g_bs, c, cprev, c_bs, dc, dg32,
                                                                                    (__nv_bfloat16*)dg16, dg_bs, (__nv_bfloat16*)dgp, dgp_ld, B, H);
  AG_LAUNCH_CHECK();
  return AG_OK;
}

}
