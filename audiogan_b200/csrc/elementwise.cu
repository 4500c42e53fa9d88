// HBM-bound kernels: weight-norm (multi-tensor), gather/pack, framing+noise, BCE, small reductions.
#include "common.cuh"

namespace ag {

// ---------------------------------------------------------------- weight norm (audiogan.py:77-80)
// One block per (tensor,row).  Finds its tensor by binary search over row_start.
__device__ __forceinline__ int find_tensor(const int32_t* row_start, int nt, int row) {
  int lo = 0, hi = nt - 1;
  while (lo < hi) {
    int mid = (lo + hi + 1) >> 1;
    if (row_start[mid] <= row) lo = mid; else hi = mid - 1;
  }
  return lo;
}

__global__ void __launch_bounds__(128) wn_fwd_kernel(const ag_wn_entry* __restrict__ table,
                                                     const int32_t* __restrict__ row_start, int nt) {
  __shared__ float red[32];
  const int row_g = blockIdx.x;
  const int ti = find_tensor(row_start, nt, row_g);
  const ag_wn_entry e = table[ti];
  const int row = row_g - row_start[ti];
  const int64_t base = (int64_t)row * e.cols;
  if (e.kind == 1) {
    for (int c = threadIdx.x; c < e.cols; c += blockDim.x) e.w[base + c] = e.v[base + c];
    return;
  }
  float ss = 0.f;
  for (int c = threadIdx.x; c < e.cols; c += blockDim.x) { float x = e.v[base + c]; ss += x * x; }
  ss = block_sum(ss, red);
  const float nrm = sqrtf(ss);
  const float sc = e.g[row] / nrm;
  if (threadIdx.x == 0) e.norm[row] = nrm;
  for (int c = threadIdx.x; c < e.cols; c += blockDim.x) e.w[base + c] = e.v[base + c] * sc;
}

// dg = <dw,v>/||v|| ; dv = g/||v|| * (dw - v*<dw,v>/||v||^2)
__global__ void __launch_bounds__(128) wn_bwd_kernel(const ag_wn_entry* __restrict__ table,
                                                     const int32_t* __restrict__ row_start, int nt) {
  __shared__ float red[32];
  const int row_g = blockIdx.x;
  const int ti = find_tensor(row_start, nt, row_g);
  const ag_wn_entry e = table[ti];
  const int row = row_g - row_start[ti];
  const int64_t base = (int64_t)row * e.cols;
  if (e.kind == 1) {
    for (int c = threadIdx.x; c < e.cols; c += blockDim.x) e.dv[base + c] = e.dw[base + c];
    return;
  }
  float dot = 0.f;
  for (int c = threadIdx.x; c < e.cols; c += blockDim.x) dot += e.dw[base + c] * e.v[base + c];
  dot = block_sum(dot, red);
  const float nrm = e.norm[row], g = e.g[row];
  const float inv = 1.f / nrm;
  if (threadIdx.x == 0) e.dg[row] = dot * inv;
  const float sc = g * inv, k = dot * inv * inv;
  for (int c = threadIdx.x; c < e.cols; c += blockDim.x)
    e.dv[base + c] = sc * (e.dw[base + c] - e.v[base + c] * k);
}

__global__ void gather_kernel(void* __restrict__ dst, const float* __restrict__ src,
                              const int32_t* __restrict__ idx, int64_t n, int dtype) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int32_t j = idx[i];
    st_any(dst, i, j >= 0 ? src[j] : 0.f, dtype);
  }
}

// ---------------------------------------------------------------- framing + noise
__global__ void frame_noise_kernel(void* __restrict__ dst, int64_t dst_ld, int64_t pad_l,
                                   const float* __restrict__ src, int64_t src_ld,
                                   const float* __restrict__ noise, float nscale, int64_t B, int64_t L, int dtype) {
  const int64_t total = B * dst_ld;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = i / dst_ld, j = i - b * dst_ld - pad_l;
    float v = 0.f;
    if (j >= 0 && j < L) {
      v = src[b * src_ld + j];
      if (noise) v += nscale * noise[b * L + j];
    }
    st_any(dst, i, v, dtype);
  }
}

// ---------------------------------------------------------------- BCE (audiogan.py:187-197)
__device__ __forceinline__ float bce_elem(float x, float t) {
  const float mx = fmaxf(-x, 0.f);
  return x - x * t + mx + logf(expf(-mx) + expf(-x - mx));
}
__global__ void __launch_bounds__(256) bce_fwd_kernel(const float* __restrict__ x, const float* __restrict__ tgt,
                                                      const float* __restrict__ w, float* __restrict__ loss, int64_t T) {
  __shared__ float red[32];
  const int64_t b = blockIdx.x;
  float acc = 0.f;
  for (int64_t t = threadIdx.x; t < T; t += blockDim.x) {
    const int64_t i = b * T + t;
    float l = bce_elem(x[i], tgt[i]);
    if (w) l *= w[i];
    acc += l;
  }
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) loss[b] = acc;
}
__global__ void bce_bwd_kernel(const float* __restrict__ x, const float* __restrict__ tgt, const float* __restrict__ w,
                               const float* __restrict__ gout, float* __restrict__ dx, int64_t B, int64_t T) {
  const int64_t n = B * T;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float g = gout[i / T] * (sigmoidf_(x[i]) - tgt[i]);
    if (w) g *= w[i];
    dx[i] = g;
  }
}
__global__ void __launch_bounds__(256) bce_const_kernel(const float* __restrict__ x, int64_t ld, const int32_t* __restrict__ len,
                                                        float target, float sign, float* __restrict__ loss_mean,
                                                        float* __restrict__ loss_ps, float* __restrict__ dlogits,
                                                        float* __restrict__ stats, int64_t B, int64_t T) {
  __shared__ float red[32];
  const int64_t b = blockIdx.x;
  const int n = len[b];
  const float invn = 1.f / (float)n, invnb = invn / (float)B;
  float acc = 0.f, corr = 0.f;
  for (int64_t t = threadIdx.x; t < T; t += blockDim.x) {
    const float xv = x[b * ld + t];
    const bool in = t < n;
    if (in) { acc += bce_elem(xv, target); corr += (sign * xv > 0.f) ? 1.f : 0.f; }
    if (dlogits) dlogits[b * ld + t] = in ? (sigmoidf_(xv) - target) * invnb : 0.f;
  }
  acc = block_sum(acc, red);
  corr = block_sum(corr, red);
  if (threadIdx.x == 0) {
    if (loss_ps) loss_ps[b] = acc * invn;
    if (loss_mean) atomicAdd(loss_mean, acc * invnb);
    if (stats) { atomicAdd(stats, corr); atomicAdd(stats + 1, (float)(n < T ? n : (int)T)); }
  }
}


// ---------------------------------------------------------------- activation-gradient assembly
__global__ void ew_grad_kernel(const ag_ew_desc d) {
  const int64_t Tp = d.pad_l + d.T + d.pad_r;
  const int64_t total = d.B * Tp * d.C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t c = i % d.C, r = (i / d.C) % Tp, b = i / (d.C * Tp);
    const int64_t t = r - d.pad_l;
    float v = 0.f;
    if (t >= 0 && t < d.T && (!d.len || t < d.len[b])) {
      if (d.g1) v += ld_any(d.g1, b * d.g1_bs + t * d.g1_rs + c * d.g1_cs, d.g1_dtype);
      if (d.g2) v += ld_any(d.g2, b * d.g2_bs + t * d.g2_rs + c * d.g2_cs, d.g2_dtype);
      if (d.act) v *= (ld_any(d.act, b * d.a_bs + t * d.a_rs + c, d.act_dtype) > 0.f) ? 1.f : d.slope;
      if (d.acc) {
        const int64_t ai = b * d.acc_bs + t * d.acc_rs + c;
        st_any(d.acc, ai, ld_any(d.acc, ai, d.acc_dtype) + v, d.acc_dtype);
      }
    }
    if (d.out) st_any(d.out, i, v, d.out_dtype);
  }
}

// 4 channels per thread (16-byte fp32 / 8-byte bf16 accesses everywhere), 32-bit index arithmetic: the scalar kernel above
// spends its time in three 64-bit divisions per element, not on the memory system.  Requires C % 4 == 0, unit channel
// strides, strides % 4 == 0.
__global__ void __launch_bounds__(256) ew_grad_vec4_kernel(const ag_ew_desc d) {
  const uint32_t C4 = (uint32_t)(d.C >> 2), Tp = (uint32_t)(d.pad_l + d.T + d.pad_r);
  const uint32_t total = (uint32_t)d.B * Tp * C4;
  float4 csum = make_float4(0.f, 0.f, 0.f, 0.f);     // d.colsum: this thread's column group is fixed (256 % C4 == 0, host-checked)
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const uint32_t rb = i / C4, c = (i - rb * C4) << 2, b = rb / Tp, r = rb - b * Tp;
    const int64_t t = (int64_t)r - d.pad_l;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (t >= 0 && t < d.T && (!d.len || t < d.len[b])) {
      if (d.g1) v = ldg4_any(d.g1, b * d.g1_bs + t * d.g1_rs + c, d.g1_dtype);
      if (d.g2) {
        const float4 w = ldg4_any(d.g2, b * d.g2_bs + t * d.g2_rs + c, d.g2_dtype);
        v.x += w.x; v.y += w.y; v.z += w.z; v.w += w.w;
      }
      if (d.act) {
        const float4 a = ldg4_any(d.act, b * d.a_bs + t * d.a_rs + c, d.act_dtype);
        v.x *= a.x > 0.f ? 1.f : d.slope; v.y *= a.y > 0.f ? 1.f : d.slope;
        v.z *= a.z > 0.f ? 1.f : d.slope; v.w *= a.w > 0.f ? 1.f : d.slope;
      }
      if (d.acc) {
        const int64_t ai = b * d.acc_bs + t * d.acc_rs + c;
        float4 o = ld4_plain_any(d.acc, ai, d.acc_dtype);
        o.x += v.x; o.y += v.y; o.z += v.z; o.w += v.w;
        st4_any(d.acc, ai, o, d.acc_dtype);
      }
    }
    if (d.out) st4_any(d.out, (int64_t)i << 2, v, d.out_dtype);
    csum.x += v.x; csum.y += v.y; csum.z += v.z; csum.w += v.w;
  }
  if (d.colsum) {                                   // fused bias gradient: column sums of v, one set of atomics per block
    __shared__ float4 red[256];
    red[threadIdx.x] = csum;
    __syncthreads();
    if (threadIdx.x < C4) {
      float4 a = red[threadIdx.x];
      for (uint32_t r = threadIdx.x + C4; r < 256; r += C4) { const float4 q = red[r]; a.x += q.x; a.y += q.y; a.z += q.z; a.w += q.w; }
      float* o = d.colsum + 4 * threadIdx.x;
      atomicAdd(o, a.x); atomicAdd(o + 1, a.y); atomicAdd(o + 2, a.z); atomicAdd(o + 3, a.w);
    }
  }
}

// out[c] += sum over rows; block = 32 columns x 8 row lanes, grid.y splits the rows.
__global__ void __launch_bounds__(256) colsum_kernel(const void* __restrict__ in, int dtype, int64_t bs, int64_t rs, int64_t T,
                                                     int64_t M, int64_t C, float* __restrict__ out, int64_t rows_per) {
  __shared__ float red[8][33];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int64_t c = (int64_t)blockIdx.x * 32 + cx;
  const int64_t m0 = (int64_t)blockIdx.y * rows_per, m1 = min(M, m0 + rows_per);
  float acc = 0.f;
  if (c < C)
    for (int64_t m = m0 + ry; m < m1; m += 8) {
      const int64_t b = m / T;
      acc += ld_any(in, b * bs + (m - b * T) * rs + c, dtype);
    }
  red[ry][cx] = acc;
  __syncthreads();
  if (ry == 0 && c < C) {
    float v = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) v += red[i][cx];
    atomicAdd(&out[c], v);
  }
}


// C % 4 == 0, C <= 128, aligned rows: one block per (row chunk, batch) -- no index division; thread = (4 channels, row lane),
// float4 loads, the row lanes meet in shared memory, one set of atomics per block.
__global__ void __launch_bounds__(256) colsum_vec4_kernel(const void* __restrict__ in, int dtype, int64_t bs, int64_t rs, int64_t T, int C4,
                                                          float* __restrict__ out, int rows_per) {
  __shared__ float4 red[256];
  const int c4 = threadIdx.x % C4, rl = threadIdx.x / C4, nrl = 256 / C4;
  const int64_t t0 = (int64_t)blockIdx.x * rows_per, t1 = min(T, t0 + rows_per);
  const int64_t p = (int64_t)blockIdx.y * bs + 4 * c4;      // element index of (batch, row 0, this thread's channels)
  float4 a = make_float4(0.f, 0.f, 0.f, 0.f), a2 = a;
  if (rl < nrl) {
    int64_t t = t0 + rl;
    for (; t + nrl < t1; t += 2 * nrl) {                 // two independent loads in flight per thread
      const float4 v = ldg4_any(in, p + t * rs, dtype);
      const float4 w = ldg4_any(in, p + (t + nrl) * rs, dtype);
      a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
      a2.x += w.x; a2.y += w.y; a2.z += w.z; a2.w += w.w;
    }
    if (t < t1) {
      const float4 v = ldg4_any(in, p + t * rs, dtype);
      a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
    }
    a.x += a2.x; a.y += a2.y; a.z += a2.z; a.w += a2.w;
  }
  red[threadIdx.x] = a;
  __syncthreads();
  if (threadIdx.x < C4) {
    float4 v = red[threadIdx.x];
    for (int r = 1; r < nrl; ++r) { const float4 w = red[r * C4 + threadIdx.x]; v.x += w.x; v.y += w.y; v.z += w.z; v.w += w.w; }
    float* q = out + 4 * threadIdx.x;
    atomicAdd(q, v.x); atomicAdd(q + 1, v.y); atomicAdd(q + 2, v.z); atomicAdd(q + 3, v.w);
  }
}

// Rank-1 data gradient (the classifier's last layer, audiogan.py:508-512 backward): out[m, n] = g[m] * w[n] * lrelu'(act[m, n]).
// HBM-bound: 16-byte accesses, act / out fp32 or bf16.
__global__ void __launch_bounds__(256) outer_dact_kernel(const float* __restrict__ g, const float* __restrict__ w, const void* act,
                                                         int act_dtype, void* out, int out_dtype, int64_t M, int N4, float slope) {
  const int64_t total = M * N4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t m = i / N4;
    const int n = (int)(i - m * N4) * 4;
    const float gv = __ldg(g + m);
    const float4 wv = make_float4(__ldg(w + n), __ldg(w + n + 1), __ldg(w + n + 2), __ldg(w + n + 3));   // w: any alignment
    float4 a;
    if (act_dtype == 0) a = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(act) + m * N4 * 4 + n));
    else {
      const uint2 u = __ldg(reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(act) + m * N4 * 4 + n));
      const __nv_bfloat162 lo = *reinterpret_cast<const __nv_bfloat162*>(&u.x), hi = *reinterpret_cast<const __nv_bfloat162*>(&u.y);
      a = make_float4(__low2float(lo), __high2float(lo), __low2float(hi), __high2float(hi));
    }
    float4 o;
    o.x = gv * wv.x * (a.x > 0.f ? 1.f : slope); o.y = gv * wv.y * (a.y > 0.f ? 1.f : slope);
    o.z = gv * wv.z * (a.z > 0.f ? 1.f : slope); o.w = gv * wv.w * (a.w > 0.f ? 1.f : slope);
    if (out_dtype == 0) *reinterpret_cast<float4*>(reinterpret_cast<float*>(out) + m * N4 * 4 + n) = o;
    else {
      __nv_bfloat162 lo = __floats2bfloat162_rn(o.x, o.y), hi = __floats2bfloat162_rn(o.z, o.w);
      uint2 u;
      u.x = *reinterpret_cast<uint32_t*>(&lo); u.y = *reinterpret_cast<uint32_t*>(&hi);
      *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(out) + m * N4 * 4 + n) = u;
    }
  }
}

// generic strided 3-D copy / accumulate: dst[b,t,c] (+)= src[b,t,c]
__global__ void copy3d_kernel(void* __restrict__ dst, int64_t d_bs, int64_t d_rs, int64_t d_cs,
                              const void* __restrict__ src, int64_t s_bs, int64_t s_rs, int64_t s_cs,
                              int64_t B, int64_t T, int64_t Cn, int accumulate, int src_dtype, int dst_dtype) {
  const int64_t total = B * T * Cn;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t c = i % Cn, t = (i / Cn) % T, b = i / (Cn * T);
    const float v = ld_any(src, b * s_bs + t * s_rs + c * s_cs, src_dtype);
    const int64_t q = b * d_bs + t * d_rs + c * d_cs;
    st_any(dst, q, accumulate ? (ld_any(dst, q, dst_dtype) + v) : v, dst_dtype);
  }
}

// ---------------------------------------------------------------- small reductions / layout
__global__ void rowgroup_sum_kernel(const float* __restrict__ in, float* __restrict__ out, int64_t T, int64_t N) {
  const int64_t b = blockIdx.y;
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const float* p = in + b * T * N + n;
  float acc = 0.f;
  for (int64_t t = 0; t < T; ++t) acc += p[t * N];
  out[b * N + n] = acc;
}
// bf16 input: two columns per thread (32-bit loads), four rows in flight
__global__ void rowgroup_sum_bf16_kernel(const __nv_bfloat16* __restrict__ in, float* __restrict__ out, int64_t T, int64_t N) {
  const int64_t b = blockIdx.y;
  const int64_t n = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 2;
  if (n >= N) return;
  const uint32_t* p = reinterpret_cast<const uint32_t*>(in + b * T * N + n);
  const int64_t st = N / 2;
  float a0 = 0.f, a1 = 0.f;
  int64_t t = 0;
  for (; t + 4 <= T; t += 4) {
    const uint32_t u0 = __ldg(p + t * st), u1 = __ldg(p + (t + 1) * st), u2 = __ldg(p + (t + 2) * st), u3 = __ldg(p + (t + 3) * st);
    a0 += (__uint_as_float(u0 << 16) + __uint_as_float(u1 << 16)) + (__uint_as_float(u2 << 16) + __uint_as_float(u3 << 16));
    a1 += (__uint_as_float(u0 & 0xffff0000u) + __uint_as_float(u1 & 0xffff0000u)) + (__uint_as_float(u2 & 0xffff0000u) + __uint_as_float(u3 & 0xffff0000u));
  }
  for (; t < T; ++t) { const uint32_t u = __ldg(p + t * st); a0 += __uint_as_float(u << 16); a1 += __uint_as_float(u & 0xffff0000u); }
  out[b * N + n] = a0;
  out[b * N + n + 1] = a1;
}

// src [B,C,T] <-> dst channel-last with strides; 32x32 smem tile transpose.
__global__ void transpose_bct_kernel(const float* __restrict__ src, float* __restrict__ dst, int64_t C, int64_t T,
                                     int64_t dst_bs, int64_t dst_rs, int to_cl) {
  __shared__ float tile[32][33];
  const int64_t b = blockIdx.z;
  const int64_t c0 = (int64_t)blockIdx.y * 32, t0 = (int64_t)blockIdx.x * 32;
  const float* s = src + b * C * T;
  float* d = dst + b * dst_bs;
  if (to_cl) {
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
      const int64_t c = c0 + i, t = t0 + threadIdx.x;
      tile[i][threadIdx.x] = (c < C && t < T) ? s[c * T + t] : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
      const int64_t t = t0 + i, c = c0 + threadIdx.x;
      if (c < C && t < T) d[t * dst_rs + c] = tile[threadIdx.x][i];
    }
  } else {  // channel-last (dst arg is the strided one) -> [B,C,T] written to src arg
    float* so = const_cast<float*>(s);
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
      const int64_t t = t0 + i, c = c0 + threadIdx.x;
      tile[i][threadIdx.x] = (c < C && t < T) ? d[t * dst_rs + c] : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
      const int64_t c = c0 + i, t = t0 + threadIdx.x;
      if (c < C && t < T) so[c * T + t] = tile[threadIdx.x][i];
    }
  }
}

// frame assembly (audiogan.py:462-464) into channel 0 of a channel-last buffer whose first slot has `slot` channels: writes
// (x, 0, ..., 0) per row in one pass (the pad channels of the slot must read as zeros for the conv GEMMs)
__global__ void frames_to_slot_kernel(void* __restrict__ dst, int ddt, int64_t d_bs, int64_t d_rs, int slot,
                                      const float* __restrict__ src, int64_t s_bs, int64_t B, int64_t L) {
  const int64_t total = B * L;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = i / L, t = i - b * L;
    const float v = __ldg(src + b * s_bs + t);
    const int64_t o = b * d_bs + t * d_rs;
    if (ddt == 1 && slot == 8) {
      *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(dst) + o) = make_uint4((uint32_t)__bfloat16_as_ushort(__float2bfloat16(v)), 0u, 0u, 0u);
    } else if (ddt == 1 && slot == 16) {                     // one 32-byte sector per row
      uint4* q = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(dst) + o);
      q[0] = make_uint4((uint32_t)__bfloat16_as_ushort(__float2bfloat16(v)), 0u, 0u, 0u);
      q[1] = make_uint4(0u, 0u, 0u, 0u);
    } else if (ddt == 0 && slot == 8) {
      float4* q = reinterpret_cast<float4*>(reinterpret_cast<float*>(dst) + o);
      q[0] = make_float4(v, 0.f, 0.f, 0.f);
      q[1] = make_float4(0.f, 0.f, 0.f, 0.f);
    } else {
      st_any(dst, o, v, ddt);
      for (int c = 1; c < slot; ++c) st_any(dst, o + c, 0.f, ddt);
    }
  }
}

// zero rows [0, head) and [tail0, rows) of every batch of a packed [B, rows, row_bytes] buffer; one store of sizeof(U) bytes
// per thread and iteration (U = uint4 when rows are 16-byte multiples and the buffer is aligned, else uint32_t / uint16_t)
template <typename U>
__global__ void zero_pads_kernel(U* __restrict__ p, int64_t B, int64_t rows, int64_t rowu, int64_t head, int64_t tail0) {
  const int64_t per = (head + rows - tail0) * rowu, total = B * per;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = i / per, j = i - b * per;
    const int64_t off = j < head * rowu ? j : tail0 * rowu + (j - head * rowu);
    p[b * rows * rowu + off] = U{};
  }
}

// Up to 8 buffers per launch (blockIdx.y = buffer): the conv stacks allocate 2-5 padded activation buffers per pass, and a 4 us
// launch per buffer is all latency.  unit[e] = bytes per store (16 / 4 / 2, by alignment).
struct PadBatch {
  void* p[8];
  long long B[8], rows[8], rowb[8], head[8], tail0[8];
  int unit[8];
};
template <typename U>
__device__ __forceinline__ void zero_pads_one(void* pv, int64_t B, int64_t rows, int64_t rowu, int64_t head, int64_t tail0) {
  U* p = reinterpret_cast<U*>(pv);
  const int64_t per = (head + rows - tail0) * rowu, total = B * per;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = i / per, j = i - b * per;
    const int64_t off = j < head * rowu ? j : tail0 * rowu + (j - head * rowu);
    p[b * rows * rowu + off] = U{};
  }
}
__global__ void zero_pads_multi_kernel(const PadBatch pb) {
  const int e = blockIdx.y, u = pb.unit[e];
  if (u == 16) zero_pads_one<uint4>(pb.p[e], pb.B[e], pb.rows[e], pb.rowb[e] / 16, pb.head[e], pb.tail0[e]);
  else if (u == 4) zero_pads_one<uint32_t>(pb.p[e], pb.B[e], pb.rows[e], pb.rowb[e] / 4, pb.head[e], pb.tail0[e]);
  else zero_pads_one<uint16_t>(pb.p[e], pb.B[e], pb.rows[e], pb.rowb[e] / 2, pb.head[e], pb.tail0[e]);
}

static inline int grid_for(int64_t n, int threads) {
  int64_t b = (n + threads - 1) / threads;
  int64_t cap = (int64_t)sm_count() * 16;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace ag

using namespace ag;
extern "C" {

int ag_wn_fwd_multi(const ag_wn_entry* table, const int32_t* row_start, int32_t nt, int32_t total_rows, void* stream) {
  AG_CHECK_ARG(table && row_start && nt > 0 && total_rows > 0, "ag_wn_fwd_multi: bad args");
  wn_fwd_kernel<<<total_rows, 128, 0, (cudaStream_t)stream>>>(table, row_start, nt);
  AG_LAUNCH_CHECK();
  return AG_OK;
}
int ag_wn_bwd_multi(const ag_wn_entry* table, const int32_t* row_start, int32_t nt, int32_t total_rows, void* stream) {
  AG_CHECK_ARG(table && row_start && nt > 0 && total_rows > 0, "ag_wn_bwd_multi: bad args");
  wn_bwd_kernel<<<total_rows, 128, 0, (cudaStream_t)stream>>>(table, row_start, nt);
  AG_LAUNCH_CHECK();
  return AG_OK;
}
int ag_gather(void* dst, const float* src, const int32_t* idx, int64_t n, int32_t dtype, void* stream) {
  AG_CHECK_ARG(dst && src && idx && n >= 0, "ag_gather: bad args");
  if (n == 0) return AG_OK;
  gather_kernel<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(dst, src, idx, n, dtype);
  AG_LAUNCH_CHECK();
  return AG_OK;
}
int ag_frame_noise(void* dst, int64_t dst_ld, int64_t pad_l, const float* src, int64_t src_ld, const float* noise,
                   float nscale, int64_t B, int64_t L, int32_t dtype, void* stream) {
  AG_CHECK_ARG(dst && src && B > 0 && L > 0 && dst_ld >= pad_l + L, "ag_frame_noise: bad args");
  frame_noise_kernel<<<grid_for(B * dst_ld, 256), 256, 0, (cudaStream_t)stream>>>(dst, dst_ld, pad_l, src, src_ld, noise,
                                                                                 nscale, B, L, dtype);
  AG_LAUNCH_CHECK();
  return AG_OK;
}
int ag_bce_fwd(const float* x, const float* tgt, const float* w, float* loss, int64_t B, int64_t T, void* stream) {
  AG_CHECK_ARG(x && tgt && loss && B > 0 && T > 0, "ag_bce_fwd: bad args");
  bce_fwd_kernel<<<(unsigned)B, 256, 0, (cudaStream_t)stream>>>(x, tgt, w, loss, T);
  AG_LAUNCH_CHECK();
  return AG_OK;
}
int ag_bce_bwd(const float* x, const float* tgt, const float* w, const float* gout, float* dx, int64_t B, int64_t T,
               void* stream) {
  AG_CHECK_ARG(x && tgt && gout && dx && B > 0 && T > 0, "ag_bce_bwd: bad args");
  bce_bwd_kernel<<<grid_for(B * T, 256), 256, 0, (cudaStream_t)stream>>>(x, tgt, w, gout, dx, B, T);
  AG_LAUNCH_CHECK();
  return AG_OK;
}
int ag_bce_const_fused(const float* x, int64_t ld, const int32_t* len, float target, float sign, float* loss_mean,
                       float* loss_ps, float* dlogits, float* stats, int64_t B, int64_t T, void* stream) {
  AG_CHECK_ARG(x && len && B > 0 && T > 0 && ld >= T, "ag_bce_const_fused: bad args");
  bce_const_kernel<<<(unsigned)B, 256, 0, (cudaStream_t)stream>>>(x, ld, len, target, sign, loss_mean, loss_ps, dlogits,
                                                                  stats, B, T);
  AG_LAUNCH_CHECK();
  return AG_OK;
}
int ag_ew_grad(const ag_ew_desc* d, void* stream) {
  AG_CHECK_ARG(d && d->B > 0 && d->T > 0 && d->C > 0 && (d->out || d->acc), "ag_ew_grad: bad args");
  const int64_t total = d->B * (d->pad_l + d->T + d->pad_r) * d->C;
  auto al = [](const void* p, int dt) { return (reinterpret_cast<uintptr_t>(p) & (dt ? 7 : 15)) == 0; };   // 4 elements
  const bool vec = d->C % 4 == 0 && total / 4 < (1ll << 31) &&
                   (!d->g1 || (d->g1_cs == 1 && d->g1_bs % 4 == 0 && d->g1_rs % 4 == 0 && al(d->g1, d->g1_dtype))) &&
                   (!d->g2 || (d->g2_cs == 1 && d->g2_bs % 4 == 0 && d->g2_rs % 4 == 0 && al(d->g2, d->g2_dtype))) &&
                   (!d->act || (d->a_bs % 4 == 0 && d->a_rs % 4 == 0 && al(d->act, d->act_dtype))) &&
                   (!d->acc || (d->acc_bs % 4 == 0 && d->acc_rs % 4 == 0 && al(d->acc, d->acc_dtype))) &&
                   (!d->out || al(d->out, d->out_dtype));
  AG_CHECK_ARG(!d->colsum || (vec && d->C / 4 <= 256 && 256 % (d->C / 4) == 0),
               "ag_ew_grad: the fused column sum needs the vector path and a channel count of 4 * 2^k <= 1024");
  if (vec) {
    ew_grad_vec4_kernel<<<grid_for(total / 4, 256), 256, 0, (cudaStream_t)stream>>>(*d);
    AG_LAUNCH_CHECK();
    return AG_OK;
  }
  ew_grad_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(*d);
  AG_LAUNCH_CHECK();
  return AG_OK;
}
int ag_colsum(const void* in, int32_t dtype, int64_t bs, int64_t rs, int64_t B, int64_t T, int64_t C, float* out, void* stream) {
  AG_CHECK_ARG(in && out && B > 0 && T > 0 && C > 0, "ag_colsum: bad args");
  if (C % 4 == 0 && C <= 128 && bs % 4 == 0 && rs % 4 == 0 && B < 65536 && (reinterpret_cast<uintptr_t>(in) & (dtype ? 7 : 15)) == 0) {
    const int rows_per = 2048;
    dim3 grid((unsigned)((T + rows_per - 1) / rows_per), (unsigned)B);
    colsum_vec4_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(in, dtype, bs, rs, T, (int)(C / 4), out, rows_per);
    AG_LAUNCH_CHECK();
    return AG_OK;
  }
  const int64_t M = B * T;
  const int64_t gx = (C + 31) / 32;
  int64_t gy = (int64_t)sm_count() * 8 / gx;
  if (gy < 1) gy = 1;
  int64_t rows_per = (M + gy - 1) / gy;
  if (rows_per < 64) rows_per = 64;
  gy = (M + rows_per - 1) / rows_per;
  colsum_kernel<<<dim3((unsigned)gx, (unsigned)gy), 256, 0, (cudaStream_t)stream>>>(in, dtype, bs, rs, T, M, C, out, rows_per);
  AG_LAUNCH_CHECK();
  return AG_OK;
}
int ag_copy3d(void* dst, int64_t d_bs, int64_t d_rs, int64_t d_cs, const void* src, int64_t s_bs, int64_t s_rs,
              int64_t s_cs, int64_t B, int64_t T, int64_t Cn, int32_t accumulate, int32_t src_dtype, int32_t dst_dtype,
              void* stream) {
  AG_CHECK_ARG(dst && src && B > 0 && T > 0 && Cn > 0, "ag_copy3d: bad args");
  copy3d_kernel<<<grid_for(B * T * Cn, 256), 256, 0, (cudaStream_t)stream>>>(dst, d_bs, d_rs, d_cs, src, s_bs, s_rs, s_cs, B, T,
                                                                            Cn, accumulate, src_dtype, dst_dtype);
  AG_LAUNCH_CHECK();
  return AG_OK;
}
int ag_frames_to_slot(void* dst, int32_t dst_dtype, int64_t d_bs, int64_t d_rs, int32_t slot, const float* src, int64_t s_bs, int64_t B,
                      int64_t L, void* stream) {
  AG_CHECK_ARG(dst && src && B > 0 && L > 0 && slot > 0, "ag_frames_to_slot: bad args");
  if (slot == 8 || slot == 16)
    AG_CHECK_ARG((reinterpret_cast<uintptr_t>(dst) & (dst_dtype ? 15 : 31)) % 16 == 0 && d_bs % 8 == 0 && d_rs % 8 == 0, "ag_frames_to_slot: unaligned");
  frames_to_slot_kernel<<<grid_for(B * L, 256), 256, 0, (cudaStream_t)stream>>>(dst, dst_dtype, d_bs, d_rs, slot, src, s_bs, B, L);
  AG_LAUNCH_CHECK();
  return AG_OK;
}
int ag_zero_pads(void* buf, int64_t B, int64_t rows, int64_t row_bytes, int64_t head, int64_t tail0, void* stream) {
  AG_CHECK_ARG(buf && B > 0 && rows > 0 && row_bytes > 0 && row_bytes % 2 == 0 && head >= 0 && tail0 >= head && tail0 <= rows &&
                   (reinterpret_cast<uintptr_t>(buf) & 1) == 0, "ag_zero_pads: bad args");
  const uintptr_t a = reinterpret_cast<uintptr_t>(buf);
  const int unit = (row_bytes % 16 == 0 && (a & 15) == 0) ? 16 : (row_bytes % 4 == 0 && (a & 3) == 0) ? 4 : 2;
  const int64_t n = B * (head + rows - tail0) * (row_bytes / unit);
  if (n == 0) return AG_OK;
  cudaStream_t s = (cudaStream_t)stream;
  if (unit == 16) zero_pads_kernel<uint4><<<grid_for(n, 256), 256, 0, s>>>(reinterpret_cast<uint4*>(buf), B, rows, row_bytes / 16, head, tail0);
  else if (unit == 4) zero_pads_kernel<uint32_t><<<grid_for(n, 256), 256, 0, s>>>(reinterpret_cast<uint32_t*>(buf), B, rows, row_bytes / 4, head, tail0);
  else zero_pads_kernel<uint16_t><<<grid_for(n, 256), 256, 0, s>>>(reinterpret_cast<uint16_t*>(buf), B, rows, row_bytes / 2, head, tail0);
  AG_LAUNCH_CHECK();
  return AG_OK;
}
int ag_zero_pads_multi(const ag_pad_entry* e, int32_t n, void* stream) {
  AG_CHECK_ARG(e && n > 0 && n <= 8, "ag_zero_pads_multi: 1..8 entries");
  PadBatch pb;
  int64_t nmax = 0;
  for (int i = 0; i < n; ++i) {
    AG_CHECK_ARG(e[i].buf && e[i].B > 0 && e[i].rows > 0 && e[i].row_bytes > 0 && e[i].row_bytes % 2 == 0 && e[i].head >= 0 &&
                     e[i].tail0 >= e[i].head && e[i].tail0 <= e[i].rows && (reinterpret_cast<uintptr_t>(e[i].buf) & 1) == 0,
                 "ag_zero_pads_multi: bad entry %d", i);
    const uintptr_t a = reinterpret_cast<uintptr_t>(e[i].buf);
    const int unit = (e[i].row_bytes % 16 == 0 && (a & 15) == 0) ? 16 : (e[i].row_bytes % 4 == 0 && (a & 3) == 0) ? 4 : 2;
    pb.p[i] = e[i].buf; pb.B[i] = e[i].B; pb.rows[i] = e[i].rows; pb.rowb[i] = e[i].row_bytes; pb.head[i] = e[i].head;
    pb.tail0[i] = e[i].tail0; pb.unit[i] = unit;
    const int64_t cnt = e[i].B * (e[i].head + e[i].rows - e[i].tail0) * (e[i].row_bytes / unit);
    if (cnt > nmax) nmax = cnt;
  }
  if (nmax == 0) return AG_OK;
  zero_pads_multi_kernel<<<dim3((unsigned)grid_for(nmax, 256), (unsigned)n), 256, 0, (cudaStream_t)stream>>>(pb);
  AG_LAUNCH_CHECK();
  return AG_OK;
}
int ag_rowgroup_sum(const void* in, int32_t dtype, float* out, int64_t B, int64_t T, int64_t N, void* stream) {
  AG_CHECK_ARG(in && out && B > 0 && T > 0 && N > 0 && B < 65536, "ag_rowgroup_sum: bad args");
  if (dtype == 1) {
    AG_CHECK_ARG(N % 2 == 0 && (reinterpret_cast<uintptr_t>(in) & 3) == 0, "ag_rowgroup_sum: bf16 input needs an even N");
    dim3 g2((unsigned)((N / 2 + 127) / 128), (unsigned)B);
    rowgroup_sum_bf16_kernel<<<g2, 128, 0, (cudaStream_t)stream>>>(reinterpret_cast<const __nv_bfloat16*>(in), out, T, N);
    AG_LAUNCH_CHECK();
    return AG_OK;
  }
  dim3 grid((unsigned)((N + 127) / 128), (unsigned)B);
  rowgroup_sum_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float*>(in), out, T, N);
  AG_LAUNCH_CHECK();
  return AG_OK;
}
int ag_transpose_bct(const float* src, float* dst, int64_t B, int64_t C, int64_t T, int64_t dst_bs, int64_t dst_rs,
                     int32_t to_cl, void* stream) {
  AG_CHECK_ARG(src && dst && B > 0 && C > 0 && T > 0 && B < 65536, "ag_transpose_bct: bad args");
  dim3 grid((unsigned)((T + 31) / 32), (unsigned)((C + 31) / 32), (unsigned)B);
  transpose_bct_kernel<<<grid, dim3(32, 8), 0, (cudaStream_t)stream>>>(src, dst, C, T, dst_bs, dst_rs, to_cl);
  AG_LAUNCH_CHECK();
  return AG_OK;
}
int ag_outer_dact(const float* g, const float* w, const void* act, int32_t act_dtype, void* out, int32_t out_dtype, int64_t M,
                  int64_t N, float slope, void* stream) {
  AG_CHECK_ARG(g && w && act && out && M > 0 && N > 0 && N % 4 == 0, "ag_outer_dact: bad args");
  AG_CHECK_ARG(((reinterpret_cast<uintptr_t>(act) | reinterpret_cast<uintptr_t>(out)) & 15) == 0, "ag_outer_dact: unaligned");
  outer_dact_kernel<<<grid_for(M * (N / 4), 256), 256, 0, (cudaStream_t)stream>>>(g, w, act, act_dtype, out, out_dtype, M, (int)(N / 4), slope);
  AG_LAUNCH_CHECK();
  return AG_OK;
}
}
