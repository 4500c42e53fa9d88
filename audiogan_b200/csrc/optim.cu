// Fused multi-tensor optimizer with the reference's per-tensor clip
// (check_grad audiogan.py:232-240, clip_grad :243-253, RMSprop :693-694 / :788 / :921).
// HBM-bound: pass 1 reads g (4 B/param), pass 2 reads p,g,state and writes p,state (20 B/param
// RMSprop, 28 B/param Adam).  One block per fixed-size chunk; 128-bit accesses where aligned.
#include "common.cuh"

namespace ag {

__global__ void __launch_bounds__(256) mt_sqnorm_kernel(const ag_mt_entry* __restrict__ table,
                                                        const int32_t* __restrict__ chunk_tensor,
                                                        const int64_t* __restrict__ chunk_off, int chunk,
                                                        float* __restrict__ sqnorm, int32_t* __restrict__ flags, float big) {
  __shared__ float red[32];
  const int ti = chunk_tensor[blockIdx.x];
  const int64_t off = chunk_off[blockIdx.x];
  const ag_mt_entry e = table[ti];
  const int64_t n = min((int64_t)chunk, e.n - off);
  const float* g = e.g + off;
  float acc = 0.f;
  int bad = 0;
  if ((reinterpret_cast<uintptr_t>(g) & 15) == 0) {
    const int64_t n4 = n >> 2;
    const float4* g4 = reinterpret_cast<const float4*>(g);
    for (int64_t i = threadIdx.x; i < n4; i += blockDim.x) {
      const float4 v = g4[i];
      acc += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
      bad |= (v.x != v.x) | (v.y != v.y) | (v.z != v.z) | (v.w != v.w);
      bad |= ((fabsf(v.x) > big) | (fabsf(v.y) > big) | (fabsf(v.z) > big) | (fabsf(v.w) > big)) << 1;
    }
    for (int64_t i = (n4 << 2) + threadIdx.x; i < n; i += blockDim.x) {
      const float v = g[i];
      acc += v * v;
      bad |= (v != v) | ((fabsf(v) > big) << 1);
    }
  } else {
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
      const float v = g[i];
      acc += v * v;
      bad |= (v != v) | ((fabsf(v) > big) << 1);
    }
  }
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) atomicAdd(&sqnorm[ti], acc);
  if (bad & 1) atomicOr(&flags[0], 1);
  if (bad & 2) atomicOr(&flags[1], 1);
}

template <bool ADAM>
__global__ void __launch_bounds__(256) mt_update_kernel(const ag_mt_entry* __restrict__ table,
                                                        const int32_t* __restrict__ chunk_tensor,
                                                        const int64_t* __restrict__ chunk_off, int chunk,
                                                        const float* __restrict__ sqnorm, float clip, float gscale,
                                                        float lr, float a1, float a2, float om1, float om2, float eps, float bc1, float bc2) {
  const int ti = chunk_tensor[blockIdx.x];
  const int64_t off = chunk_off[blockIdx.x];
  const ag_mt_entry e = table[ti];
  const int64_t n = min((int64_t)chunk, e.n - off);
  float sc = gscale;
  if (clip > 0.f) {
    const float nrm = sqrtf(sqnorm[ti]) * gscale;       // norm of the (already scaled) gradient
    if (nrm > clip) sc = gscale / (nrm / clip);         // audiogan.py:251-252
  }
  float* p = e.p + off;
  const float* g = e.g + off;
  float* s1 = e.s1 + off;
  float* s2 = ADAM ? e.s2 + off : nullptr;
  auto upd = [&](float& pv, float gv, float& s1v, float& s2v) {
    gv *= sc;
    if (ADAM) {
      s1v = a1 * s1v + om1 * gv;
      s2v = a2 * s2v + om2 * gv * gv;
      pv -= lr * (s1v / bc1) / (sqrtf(s2v / bc2) + eps);
    } else {
      s1v = a1 * s1v + om1 * gv * gv;
      pv -= lr * gv / (sqrtf(s1v) + eps);
    }
  };
  const bool al = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(s1) |
                    (ADAM ? reinterpret_cast<uintptr_t>(s2) : 0)) & 15) == 0;
  int64_t done = 0;
  if (al) {
    const int64_t n4 = n >> 2;
    for (int64_t i = threadIdx.x; i < n4; i += blockDim.x) {
      float4 pv = reinterpret_cast<float4*>(p)[i];
      const float4 gv = reinterpret_cast<const float4*>(g)[i];
      float4 sv = reinterpret_cast<float4*>(s1)[i];
      float4 tv = ADAM ? reinterpret_cast<float4*>(s2)[i] : make_float4(0, 0, 0, 0);
      upd(pv.x, gv.x, sv.x, tv.x); upd(pv.y, gv.y, sv.y, tv.y);
      upd(pv.z, gv.z, sv.z, tv.z); upd(pv.w, gv.w, sv.w, tv.w);
      reinterpret_cast<float4*>(p)[i] = pv;
      reinterpret_cast<float4*>(s1)[i] = sv;
      if (ADAM) reinterpret_cast<float4*>(s2)[i] = tv;
    }
    done = n4 << 2;
  }
  for (int64_t i = done + threadIdx.x; i < n; i += blockDim.x) {
    float pv = p[i], sv = s1[i], tv = ADAM ? s2[i] : 0.f;
    upd(pv, g[i], sv, tv);
    p[i] = pv; s1[i] = sv;
    if (ADAM) s2[i] = tv;
  }
}

// In-place per-tensor clip (audiogan.py:243-253): g *= clip/||g|| where ||g|| > clip.
__global__ void __launch_bounds__(256) mt_clip_kernel(const ag_mt_entry* __restrict__ table,
                                                      const int32_t* __restrict__ chunk_tensor,
                                                      const int64_t* __restrict__ chunk_off, int chunk,
                                                      const float* __restrict__ sqnorm, float clip) {
  const int ti = chunk_tensor[blockIdx.x];
  const float nrm = sqrtf(sqnorm[ti]);
  if (!(nrm > clip)) return;
  const float sc = 1.f / (nrm / clip);
  const int64_t off = chunk_off[blockIdx.x];
  const ag_mt_entry e = table[ti];
  const int64_t n = min((int64_t)chunk, e.n - off);
  float* g = const_cast<float*>(e.g) + off;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) g[i] *= sc;
}

}  // namespace ag

using namespace ag;
extern "C" {
int ag_mt_sqnorm(const ag_mt_entry* table, const int32_t* ct, const int64_t* co, int32_t nchunks, int32_t chunk,
                 float* sqnorm, int32_t* flags, float big, void* stream) {
  AG_CHECK_ARG(table && ct && co && nchunks > 0 && chunk > 0 && sqnorm && flags, "ag_mt_sqnorm: bad args");
  mt_sqnorm_kernel<<<nchunks, 256, 0, (cudaStream_t)stream>>>(table, ct, co, chunk, sqnorm, flags, big > 0.f ? big : 1e5f);
  AG_LAUNCH_CHECK();
  return AG_OK;
}
int ag_mt_clip(const ag_mt_entry* table, const int32_t* ct, const int64_t* co, int32_t nchunks, int32_t chunk,
               const float* sqnorm, float clip, void* stream) {
  AG_CHECK_ARG(table && ct && co && nchunks > 0 && chunk > 0 && sqnorm && clip > 0.f, "ag_mt_clip: bad args");
  mt_clip_kernel<<<nchunks, 256, 0, (cudaStream_t)stream>>>(table, ct, co, chunk, sqnorm, clip);
  AG_LAUNCH_CHECK();
  return AG_OK;
}
int ag_mt_rmsprop(const ag_mt_entry* table, const int32_t* ct, const int64_t* co, int32_t nchunks, int32_t chunk,
                  const float* sqnorm, float clip, float gscale, double lr, double alpha, double eps, void* stream) {
  AG_CHECK_ARG(table && ct && co && nchunks > 0 && chunk > 0 && (clip <= 0.f || sqnorm), "ag_mt_rmsprop: bad args");
  mt_update_kernel<false><<<nchunks, 256, 0, (cudaStream_t)stream>>>(table, ct, co, chunk, sqnorm, clip,
                                                                      gscale == 0.f ? 1.f : gscale, (float)lr, (float)alpha, 0.f,
                                                                      (float)(1.0 - alpha), 0.f, (float)eps, 1.f, 1.f);
  AG_LAUNCH_CHECK();
  return AG_OK;
}
int ag_mt_adam(const ag_mt_entry* table, const int32_t* ct, const int64_t* co, int32_t nchunks, int32_t chunk,
               const float* sqnorm, float clip, float gscale, double lr, double b1, double b2, double eps, int32_t step,
               void* stream) {
  AG_CHECK_ARG(table && ct && co && nchunks > 0 && chunk > 0 && step > 0 && (clip <= 0.f || sqnorm), "ag_mt_adam: bad args");
  // hyper-parameters arrive as doubles so that 1 - beta and the bias corrections are rounded ONCE, as torch.optim does
  const float bc1 = (float)(1.0 - pow(b1, (double)step)), bc2 = (float)(1.0 - pow(b2, (double)step));
  mt_update_kernel<true><<<nchunks, 256, 0, (cudaStream_t)stream>>>(table, ct, co, chunk, sqnorm, clip,
                                                                     gscale == 0.f ? 1.f : gscale, (float)lr, (float)b1, (float)b2,
                                                                     (float)(1.0 - b1), (float)(1.0 - b2), (float)eps, bc1, bc2);
  AG_LAUNCH_CHECK();
  return AG_OK;
}
}
