// The generator's last layer, Conv1d(C -> 1, k = 3) over the dense channel-last buffer (audiogan.py:403-407, :467),
// and its two gradients.  With one output channel there is nothing for tensor cores to do: each kernel streams the
// [B, T + k - 1, C] buffer once -- HBM-bound (4 B per buffer element), coalesced 16-byte accesses.
#include "common.cuh"

namespace ag {

// out[b,t] = bias + sum_{j<k,c<C} X[b, t+j, c] * w[j*C + c].  One warp per output row; the k*C window is contiguous.
// XDT: 0 = fp32 storage (4 elements per 16-byte load), 1 = bf16 storage (8 elements per 16-byte load, K % 8 == 0).
template <int XDT>
__global__ void __launch_bounds__(256) conv1out_fwd_kernel(const void* __restrict__ X, int64_t x_bs, int C, int k,
                                                           const float* __restrict__ w, const float* __restrict__ bias,
                                                           float* __restrict__ out, int64_t B, int64_t T) {
  extern __shared__ float ws[];                      // k*C weights
  const int K = k * C;
  for (int i = threadIdx.x; i < K; i += blockDim.x) ws[i] = w[i];
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const float bv = bias ? bias[0] : 0.f;
  const int64_t rows = B * T;
  for (int64_t m = (int64_t)blockIdx.x * nw + wid; m < rows; m += (int64_t)gridDim.x * nw) {
    const int64_t b = m / T, t = m - b * T;
    float acc = 0.f;
    if (XDT == 0) {
      const float4* p = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(X) + b * x_bs + t * C);
      for (int k4 = lane; k4 < K / 4; k4 += 32) {
        const float4 v = __ldg(p + k4);
        const float4 q = *reinterpret_cast<const float4*>(ws + 4 * k4);
        acc += v.x * q.x + v.y * q.y + v.z * q.z + v.w * q.w;
      }
    } else {
      const uint4* p = reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(X) + b * x_bs + t * C);
      for (int k8 = lane; k8 < K / 8; k8 += 32) {
        const uint4 u = __ldg(p + k8);
        const float4 q0 = *reinterpret_cast<const float4*>(ws + 8 * k8), q1 = *reinterpret_cast<const float4*>(ws + 8 * k8 + 4);
        const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&u.x), b2 = *reinterpret_cast<const __nv_bfloat162*>(&u.y);
        const __nv_bfloat162 c2 = *reinterpret_cast<const __nv_bfloat162*>(&u.z), d2 = *reinterpret_cast<const __nv_bfloat162*>(&u.w);
        acc += __low2float(a) * q0.x + __high2float(a) * q0.y + __low2float(b2) * q0.z + __high2float(b2) * q0.w +
               __low2float(c2) * q1.x + __high2float(c2) * q1.y + __low2float(d2) * q1.z + __high2float(d2) * q1.w;
      }
    }
    acc = warp_sum(acc);
    if (lane == 0) out[m] = acc + bv;
  }
}

// Streaming version for C <= 128 (the dense buffer has 120 channels): a warp walks a strip of rows, lane = 4 channels, so
// every buffer row is loaded ONCE (one coalesced 8/16-byte load per lane, 4 rows in flight).  The tap partials of a row are
// folded -- per lane, before any reduction (the window sum is linear) -- into a rolling window of output partials; each batch
// of 4 rows completes 4 outputs, which are reduced over the warp with a halving butterfly: 6 shuffles per 4 rows.
template <int XDT, int KK>
__global__ void __launch_bounds__(256) conv1out_fwd_strip_kernel(const void* __restrict__ X, int64_t x_bs, int C,
                                                                 const float* __restrict__ w, const float* __restrict__ bias,
                                                                 float* __restrict__ out, int64_t B, int64_t T, int spb, int S) {
  const int lane = threadIdx.x & 31;
  const int64_t strip = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (strip >= B * spb) return;
  const int64_t b = strip / spb;
  const int64_t t0 = (strip - b * spb) * S, t1 = min(T, t0 + S);
  const bool active = 4 * lane < C;
  float4 wv[KK];
#pragma unroll
  for (int j = 0; j < KK; ++j)
    wv[j] = active ? *reinterpret_cast<const float4*>(w + j * C + 4 * lane) : make_float4(0.f, 0.f, 0.f, 0.f);
  const float bv = bias ? bias[0] : 0.f;
  float carry[KK - 1];
#pragma unroll
  for (int j = 0; j < KK - 1; ++j) carry[j] = 0.f;
  const int64_t xb = b * x_bs + 4 * lane;
  const int64_t rend = t1 + KK - 1;                 // input rows [t0, rend)
  const bool hi = (lane & 16) != 0, hi2 = (lane & 8) != 0;
  const int idx = ((lane >> 4) & 1) * 2 + ((lane >> 3) & 1);
  for (int64_t rr = t0; rr < rend; rr += 8) {
    float4 x8[8];                                     // 8 rows in flight per lane (the kernel is bound by bytes in flight)
#pragma unroll
    for (int i = 0; i < 8; ++i)
      x8[i] = (active && rr + i < rend) ? ldg4_any(X, xb + (rr + i) * C, XDT) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
    const int64_t r0 = rr + 4 * h;
    if (r0 >= rend) break;
    float4 x[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) x[i] = x8[4 * h + i];
    float o[4 + KK - 1];                              // o[m]: this lane's partial of output r0 - (KK-1) + m
#pragma unroll
    for (int m = 0; m < 4 + KK - 1; ++m) o[m] = m < KK - 1 ? carry[m] : 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < KK; ++j)                     // row r0 + i is tap j of output r0 + i - j
        o[i - j + KK - 1] += x[i].x * wv[j].x + x[i].y * wv[j].y + x[i].z * wv[j].z + x[i].w * wv[j].w;
#pragma unroll
    for (int j = 0; j < KK - 1; ++j) carry[j] = o[4 + j];
    // outputs o[0..3] are complete: halving butterfly (lanes < 16 keep 0,1; then bit 3 picks one), then 3 plain steps
    const float s0 = __shfl_xor_sync(0xffffffffu, hi ? o[0] : o[2], 16), s1 = __shfl_xor_sync(0xffffffffu, hi ? o[1] : o[3], 16);
    const float k0 = (hi ? o[2] : o[0]) + s0, k1 = (hi ? o[3] : o[1]) + s1;
    float v = (hi2 ? k1 : k0) + __shfl_xor_sync(0xffffffffu, hi2 ? k0 : k1, 8);
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    const int64_t t = r0 - (KK - 1) + idx;
    if ((lane & 7) == 0 && t >= t0 && t < t1) out[b * T + t] = v + bv;
    }
  }
}

// dX[b, t', c] = sum_j g[b, t'-j] * w[j*C + c], t' in [0, T+k-1).  Thread = 4 channels, block walks rows.
template <int KMAX, int dxdt>
__global__ void __launch_bounds__(128) conv1out_dgrad_kernel(const float* __restrict__ g, const float* __restrict__ w,
                                                             void* __restrict__ dX, int64_t dx_bs, int C, int k, int64_t T,
                                                             int rows_per_block) {
  const int c4 = threadIdx.x;                        // channel group
  if (c4 * 4 >= C) return;
  const int64_t b = blockIdx.y;
  const int64_t t0 = (int64_t)blockIdx.x * rows_per_block, t1 = min(T + k - 1, t0 + rows_per_block);
  float4 wv[KMAX];
#pragma unroll
  for (int j = 0; j < KMAX; ++j) wv[j] = j < k ? *reinterpret_cast<const float4*>(w + j * C + 4 * c4) : make_float4(0.f, 0.f, 0.f, 0.f);
  const float* gb = g + b * T;
  for (int64_t t = t0; t < t1; ++t) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int j = 0; j < KMAX; ++j) {
      const int64_t tg = t - j;
      const float gv = (j < k && tg >= 0 && tg < T) ? __ldg(gb + tg) : 0.f;
      acc.x += gv * wv[j].x; acc.y += gv * wv[j].y; acc.z += gv * wv[j].z; acc.w += gv * wv[j].w;
    }
    st4_any(dX, b * dx_bs + t * C + 4 * c4, acc, dxdt);
  }
}

// dw[j*C + c] += sum_{b,t} g[b,t] * X[b, t+j, c];  dw[k*C] += sum g.   Thread = (4 channels, row lane): the block's 4 warps
// walk interleaved rows (4 independent load streams per block instead of one); every X element is read once.
template <int KMAX, int xdt>
__global__ void __launch_bounds__(128) conv1out_wgrad_kernel(const float* __restrict__ g, const void* __restrict__ X,
                                                             int64_t x_bs, int C, int k, float* __restrict__ dw, int64_t T,
                                                             int rows_per_block) {
  const int c4 = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int64_t b = blockIdx.y;
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_block, r1 = min(T + k - 1, r0 + rows_per_block);   // X rows
  const float* gb = g + b * T;
  for (int cb = 0; cb * 128 < C; ++cb) {               // 32 channel groups (128 channels) per pass
    const int cc = min(cb * 128 + 4 * c4, C - 4);      // lanes past C redo the last group (no divergent barriers); not added
    const bool live = cb * 128 + 4 * c4 < C;
    float4 acc[KMAX];
#pragma unroll
    for (int j = 0; j < KMAX; ++j) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int64_t r = r0 + rl; r < r1; r += 16) {        // 4 rows of this lane in flight before the first use
      float4 x[4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
        x[i] = (r + 4 * i < r1) ? ldg4_any(X, b * x_bs + (r + 4 * i) * C + cc, xdt) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
#pragma unroll
        for (int j = 0; j < KMAX; ++j) {
          const int64_t t = r + 4 * i - j;             // X row r is tap j of output t = r - j
          const float gv = (j < k && t >= 0 && t < T && r + 4 * i < r1) ? __ldg(gb + t) : 0.f;
          acc[j].x += gv * x[i].x; acc[j].y += gv * x[i].y; acc[j].z += gv * x[i].z; acc[j].w += gv * x[i].w;
        }
      }
    }
    // the 4 row lanes meet in shared memory: same-address L2 atomics serialise, so one set per block, not per warp
    __shared__ float4 red[3][KMAX][32];
    if (rl > 0) {
#pragma unroll
      for (int j = 0; j < KMAX; ++j) red[rl - 1][j][c4] = acc[j];
    }
    __syncthreads();
    if (rl == 0 && live) {
#pragma unroll
      for (int j = 0; j < KMAX; ++j) {
        if (j < k) {
          float4 a = acc[j];
#pragma unroll
          for (int o = 0; o < 3; ++o) { const float4 v = red[o][j][c4]; a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w; }
          float* q = dw + j * C + cc;
          atomicAdd(q, a.x); atomicAdd(q + 1, a.y); atomicAdd(q + 2, a.z); atomicAdd(q + 3, a.w);
        }
      }
    }
    __syncthreads();
  }
  {                                                    // the bias gradient over this block's outputs: one atomic per warp
    float s = 0.f;
    for (int64_t t = r0 + threadIdx.x; t < min(T, r0 + (int64_t)rows_per_block); t += 128) s += __ldg(gb + t);
    s = warp_sum(s);
    if (c4 == 0) atomicAdd(dw + k * C, s);
  }
}

// Streaming version for C <= 128 and a compile-time tap count: a warp owns a strip of S consecutive buffer rows, lane = 4
// channels, 8 rows in flight; the KK + 7 gradient values a batch of 8 rows needs are loaded once (the kernel above re-derives
// a bounds-checked g for every (row, tap) pair: ~200 instructions per KB of X, issue-bound at 1.4 TB/s).  X row r is tap j of
// output t = r - j.  The block's 8 warps meet in shared memory: one set of atomics per block.
template <int KK, int xdt>
__global__ void __launch_bounds__(256) conv1out_wgrad_strip_kernel(const float* __restrict__ g, const void* __restrict__ X, int64_t x_bs, int C,
                                                                   float* __restrict__ dw, int64_t B, int64_t T, int spb, int S) {
  __shared__ float4 red[7][KK][32];
  __shared__ float redb[8];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t strip = (int64_t)blockIdx.x * 8 + w;
  const bool active = 4 * lane < C;
  float4 acc[KK];
#pragma unroll
  for (int j = 0; j < KK; ++j) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
  float gsum = 0.f;
  if (strip < B * spb) {
    const int64_t b = strip / spb;
    const int64_t r0 = (strip - b * spb) * S, r1 = min(T + KK - 1, r0 + (int64_t)S);     // buffer rows of this strip
    const float* gb = g + b * T;
    const int64_t xb = b * x_bs + 4 * lane;
    for (int64_t r = r0; r < r1; r += 8) {
      float4 x[8];
#pragma unroll
      for (int i = 0; i < 8; ++i)
        x[i] = (active && r + i < r1) ? ldg4_any(X, xb + (r + i) * C, xdt) : make_float4(0.f, 0.f, 0.f, 0.f);
      float gv[8 + KK - 1];                            // gv[m] = g[r - (KK-1) + m], zero outside [0, T) and past the strip's rows
#pragma unroll
      for (int m = 0; m < 8 + KK - 1; ++m) {
        const int64_t t = r - (KK - 1) + m;
        gv[m] = (t >= 0 && t < T) ? __ldg(gb + t) : 0.f;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < KK; ++j) {                 // row r + i, tap j: output t = r + i - j -> gv[i - j + KK - 1]
          const float gq = gv[i - j + KK - 1];
          acc[j].x += gq * x[i].x; acc[j].y += gq * x[i].y; acc[j].z += gq * x[i].z; acc[j].w += gq * x[i].w;
        }
    }
    // bias gradient: the outputs t in [r0, min(T, r1)) belong to this strip
    for (int64_t t = r0 + lane; t < min(T, r1); t += 32) gsum += __ldg(gb + t);
  }
  gsum = warp_sum(gsum);
  if (w > 0) {
#pragma unroll
    for (int j = 0; j < KK; ++j) red[w - 1][j][lane] = acc[j];
  }
  if (lane == 0) redb[w] = gsum;
  __syncthreads();
  if (w == 0) {
    if (active) {
#pragma unroll
      for (int j = 0; j < KK; ++j) {
        float4 a = acc[j];
#pragma unroll
        for (int o = 0; o < 7; ++o) { const float4 v = red[o][j][lane]; a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w; }
        float* q = dw + j * C + 4 * lane;
        atomicAdd(q, a.x); atomicAdd(q + 1, a.y); atomicAdd(q + 2, a.z); atomicAdd(q + 3, a.w);
      }
    }
    if (lane == 0) {
      float sb = 0.f;
#pragma unroll
      for (int o = 0; o < 8; ++o) sb += redb[o];
      atomicAdd(dw + KK * C, sb);
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// The discriminator's FIRST layer, Conv1d(1 -> C, k, stride s) on the raw waveform (audiogan.py:527-536 with C_in = 1).
// As a GEMM it has K = k = 7: nothing for tensor cores to do, the layer is a stream over the [B, T, C] output (HBM-bound).
// out[b, t, c] = t < len[b] ? lrelu(bias[c] + sum_j w[c*k + j] * x[b*x_ld + s*t + j]) : 0.   Thread = (row, 4 channels).
template <int KMAX>
__global__ void __launch_bounds__(256) conv1in_fwd_kernel(const float* __restrict__ x, int64_t x_ld, const float* __restrict__ w,
                                                          const float* __restrict__ bias, void* __restrict__ out, int odt, int64_t out_bs,
                                                          int k, int s, int C4, int64_t T, const int32_t* __restrict__ len, float slope) {
  const int c4 = threadIdx.x % C4, rl = threadIdx.x / C4, nrl = 256 / C4;
  if (rl >= nrl) return;
  const int64_t b = blockIdx.y;
  float wv[4][KMAX], bv[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    bv[e] = bias ? bias[4 * c4 + e] : 0.f;
#pragma unroll
    for (int j = 0; j < KMAX; ++j) wv[e][j] = j < k ? w[(4 * c4 + e) * k + j] : 0.f;
  }
  const int64_t Lb = len ? len[b] : T;
  const float* xb = x + b * x_ld;
  const int64_t ob = b * out_bs + 4 * c4;
  const int64_t t1 = min(T, ((int64_t)blockIdx.x + 1) * 1024);
  for (int64_t t = (int64_t)blockIdx.x * 1024 + rl; t < t1; t += nrl) {
    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
    if (t < Lb) {
      float xv[KMAX];
#pragma unroll
      for (int j = 0; j < KMAX; ++j) xv[j] = j < k ? __ldg(xb + s * t + j) : 0.f;
      float a[4] = {bv[0], bv[1], bv[2], bv[3]};
#pragma unroll
      for (int j = 0; j < KMAX; ++j)
#pragma unroll
        for (int e = 0; e < 4; ++e) a[e] = fmaf(wv[e][j], xv[j], a[e]);
      o = make_float4(a[0] > 0.f ? a[0] : a[0] * slope, a[1] > 0.f ? a[1] : a[1] * slope, a[2] > 0.f ? a[2] : a[2] * slope,
                      a[3] > 0.f ? a[3] : a[3] * slope);
    }
    st4_any(out, ob + t * (4 * C4), o, odt);
  }
}

// dw[c*(k+1) + j] += sum_{b,t} dy[b,t,c] * x[b*x_ld + s*t + j],  dw[c*(k+1) + k] += sum dy[b,t,c]   (weight + bias gradient).
template <int KMAX>
__global__ void __launch_bounds__(256) conv1in_wgrad_kernel(const void* __restrict__ dy, int ydt, int64_t dy_bs, const float* __restrict__ x,
                                                            int64_t x_ld, float* __restrict__ dw, int k, int s, int C4, int64_t T) {
  __shared__ float red[256];
  const int c4 = threadIdx.x % C4, rl = threadIdx.x / C4, nrl = 256 / C4;
  const int64_t b = blockIdx.y;
  float acc[4][KMAX + 1];
#pragma unroll
  for (int e = 0; e < 4; ++e)
#pragma unroll
    for (int j = 0; j <= KMAX; ++j) acc[e][j] = 0.f;
  if (rl < nrl) {
    const float* xb = x + b * x_ld;
    const int64_t yb = b * dy_bs + 4 * c4;
    const int64_t t1 = min(T, ((int64_t)blockIdx.x + 1) * 2048);
    for (int64_t t = (int64_t)blockIdx.x * 2048 + rl; t < t1; t += nrl) {
      const float4 g = ldg4_any(dy, yb + t * (4 * C4), ydt);
      const float gv[4] = {g.x, g.y, g.z, g.w};
#pragma unroll
      for (int j = 0; j < KMAX; ++j) {
        const float xv = j < k ? __ldg(xb + s * t + j) : 0.f;
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[e][j] = fmaf(gv[e], xv, acc[e][j]);
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[e][KMAX] += gv[e];
    }
  }
  if ((C4 & (C4 - 1)) == 0 && C4 <= 8) {
    // C <= 32 channels (the default first layer has 16): lanes with the same channel group meet by shuffles, the 8 warps in shared
    // memory, one atomic per value and block.  (The generic path below runs 4 (k + 1) block reductions with a serial 256/C4-term
    // sum each: ~30 us per block at C4 = 4.)
    __shared__ float sred[8][4 * (KMAX + 1) * 8];
    const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
#pragma unroll
    for (int e = 0; e < 4; ++e)
#pragma unroll
      for (int j = 0; j <= KMAX; ++j) {
        float v = acc[e][j];
        for (int o = C4; o < 32; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane < C4) sred[wp][(e * (KMAX + 1) + j) * C4 + lane] = v;
      }
    __syncthreads();
    for (int idx = threadIdx.x; idx < 4 * (KMAX + 1) * C4; idx += 256) {
      const int cc = idx % C4, ej = idx / C4, e = ej / (KMAX + 1), j = ej - e * (KMAX + 1);
      if (j < k || j == KMAX) {
        float v = 0.f;
#pragma unroll
        for (int q = 0; q < 8; ++q) v += sred[q][idx];
        atomicAdd(dw + (4 * cc + e) * (k + 1) + (j == KMAX ? k : j), v);
      }
    }
    return;
  }
  // block reduction over the row lanes, one value at a time (k + 1 values x 4 channels), then one atomic per value
#pragma unroll
  for (int e = 0; e < 4; ++e) {
#pragma unroll
    for (int j = 0; j <= KMAX; ++j) {
      if (j < k || j == KMAX) {
        __syncthreads();
        red[threadIdx.x] = rl < nrl ? acc[e][j] : 0.f;
        __syncthreads();
        if (threadIdx.x < C4) {
          float v = 0.f;
          for (int r = 0; r < nrl; ++r) v += red[r * C4 + threadIdx.x];
          atomicAdd(dw + (4 * threadIdx.x + e) * (k + 1) + (j == KMAX ? k : j), v);
        }
      }
    }
  }
}

// Data gradient of that first layer (the gradient of the raw waveform, audiogan.py:769, :145, :131 -- x_grad_norm, FGSM and the
// generator update need it): dx[b, u] = sum_{t, j : s*t + j == u} sum_c dy[b, t, c] * w[c*k + j] in padded coordinates
// u in [0, Tin + 2p), zero outside the real samples.  As a GEMM it has N = s = 2 columns; here: one thread per sample.
__global__ void __launch_bounds__(256) conv1in_dgrad_kernel(const void* __restrict__ dy, int ydt, int64_t dy_bs, const float* __restrict__ w,
                                                            float* __restrict__ dx, int64_t dx_ld, int k, int s, int p, int C,
                                                            int64_t T, int64_t Tin) {
  extern __shared__ float wsm[];                    // [k][C] (transposed: the channel loop reads consecutive floats)
  for (int i = threadIdx.x; i < k * C; i += blockDim.x) { const int c = i / k, j = i - c * k; wsm[j * C + c] = w[i]; }
  __syncthreads();
  const int64_t b = blockIdx.y;
  const int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= dx_ld) return;
  float acc = 0.f;
  if (u >= p && u < p + Tin) {
    for (int j = 0; j < k; ++j) {
      const int64_t r = u - j;
      if (r < 0 || r % s != 0) continue;
      const int64_t t = r / s;
      if (t >= T) continue;
      const int64_t row = b * dy_bs + t * C;
      const float* wj = wsm + j * C;
      for (int c = 0; c < C; c += 4) {
        const float4 g = ldg4_any(dy, row + c, ydt);
        acc += g.x * wj[c] + g.y * wj[c + 1] + g.z * wj[c + 2] + g.w * wj[c + 3];
      }
    }
  }
  dx[b * dx_ld + u] = acc;
}

// out[k] += sum_m g[m] * X[m, k], out[K] += sum_m g[m]: weight + bias gradient of a Linear(K -> 1) (the classifier's last layer,
// audiogan.py:508-512) -- a GEMM with N = 1, here a weighted column sum over the packed [M, K] activation (HBM-bound).
__global__ void __launch_bounds__(256) wcolsum_kernel(const float* __restrict__ g, const void* __restrict__ X, int xdt, int64_t M, int K4,
                                                      float* __restrict__ out, int64_t rows_per) {
  __shared__ float4 red[256];
  const int c4 = threadIdx.x % K4, rl = threadIdx.x / K4, nrl = 256 / K4;
  const int64_t m0 = (int64_t)blockIdx.x * rows_per, m1 = min(M, m0 + rows_per);
  float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
  float gs = 0.f;
  if (rl < nrl)
    for (int64_t m = m0 + rl; m < m1; m += 4 * nrl) {      // 4 rows of this thread in flight (K = 512: only 2 row lanes per block)
      float4 x[4];
      float gv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int64_t mm = m + (int64_t)i * nrl;
        const bool ok = mm < m1;
        gv[i] = ok ? __ldg(g + mm) : 0.f;
        x[i] = ok ? ldg4_any(X, mm * (4 * K4) + 4 * c4, xdt) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        a.x += gv[i] * x[i].x; a.y += gv[i] * x[i].y; a.z += gv[i] * x[i].z; a.w += gv[i] * x[i].w;
        gs += gv[i];
      }
    }
  red[threadIdx.x] = a;
  __syncthreads();
  if (threadIdx.x < K4) {
    float4 v = red[threadIdx.x];
    for (int r = 1; r < nrl; ++r) { const float4 q = red[r * K4 + threadIdx.x]; v.x += q.x; v.y += q.y; v.z += q.z; v.w += q.w; }
    float* o = out + 4 * threadIdx.x;
    atomicAdd(o, v.x); atomicAdd(o + 1, v.y); atomicAdd(o + 2, v.z); atomicAdd(o + 3, v.w);
  }
  if (c4 == 0 && rl < nrl) atomicAdd(out + 4 * K4, gs);          // every row lane of column group 0 saw a disjoint set of rows
}


// out[m] = sum_k X[m, k] * w[k] + bias[0]: forward of a Linear(K -> 1) (the classifier's last layer, audiogan.py:508-512) -- a GEMM with
// N = 1 wastes a 128 x 16 tensor-core tile per 128 rows; this is one warp per row streaming the packed [M, K] activation (HBM-bound).
__global__ void __launch_bounds__(256) rowdot_kernel(const void* __restrict__ X, int xdt, const float* __restrict__ w,
                                                     const float* __restrict__ bias, int64_t M, int K4, float* __restrict__ out) {
  __shared__ float4 ws[256];
  for (int i = threadIdx.x; i < K4; i += 256) ws[i] = make_float4(__ldg(w + 4 * i), __ldg(w + 4 * i + 1), __ldg(w + 4 * i + 2), __ldg(w + 4 * i + 3));
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const float b = bias ? __ldg(bias) : 0.f;
  for (int64_t m = (int64_t)blockIdx.x * 8 + wid; m < M; m += (int64_t)gridDim.x * 8) {
    float acc = 0.f;
    for (int k4 = lane; k4 < K4; k4 += 32) {
      const float4 x = ldg4_any(X, m * (4 * (int64_t)K4) + 4 * k4, xdt);
      const float4 q = ws[k4];
      acc += x.x * q.x + x.y * q.y + x.z * q.z + x.w * q.w;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) out[m] = acc + b;
  }
}

}  // namespace ag

using namespace ag;
extern "C" {

int ag_conv1in_fwd(const float* x, int64_t x_ld, const float* w, const float* bias, void* out, int32_t out_dtype, int64_t out_bs,
                   int32_t k, int32_t s, int64_t C, int64_t B, int64_t T, const int32_t* len, float slope, void* stream) {
  AG_CHECK_ARG(x && w && out && B > 0 && B < 65536 && T > 0 && C > 0 && C % 4 == 0 && C <= 1024 && k > 0 && k <= 8 && s > 0 &&
                   out_bs % 4 == 0 && (reinterpret_cast<uintptr_t>(out) & (out_dtype ? 7 : 15)) == 0, "ag_conv1in_fwd: bad args");
  dim3 grid((unsigned)((T + 1023) / 1024), (unsigned)B);
  conv1in_fwd_kernel<8><<<grid, 256, 0, (cudaStream_t)stream>>>(x, x_ld, w, bias, out, out_dtype, out_bs, k, s, (int)(C / 4), T, len, slope);
  AG_LAUNCH_CHECK();
  return AG_OK;
}

int ag_conv1in_wgrad(const void* dy, int32_t dy_dtype, int64_t dy_bs, const float* x, int64_t x_ld, float* dw, int32_t k, int32_t s, int64_t C,
                     int64_t B, int64_t T, void* stream) {
  AG_CHECK_ARG(dy && x && dw && B > 0 && B < 65536 && T > 0 && C > 0 && C % 4 == 0 && C <= 1024 && k > 0 && k <= 8 && s > 0 &&
                   dy_bs % 4 == 0 && (reinterpret_cast<uintptr_t>(dy) & (dy_dtype ? 7 : 15)) == 0, "ag_conv1in_wgrad: bad args");
  dim3 grid((unsigned)((T + 2047) / 2048), (unsigned)B);
  conv1in_wgrad_kernel<8><<<grid, 256, 0, (cudaStream_t)stream>>>(dy, dy_dtype, dy_bs, x, x_ld, dw, k, s, (int)(C / 4), T);
  AG_LAUNCH_CHECK();
  return AG_OK;
}

int ag_conv1out_fwd(const void* X, int32_t x_dtype, int64_t x_bs, int64_t C, int32_t k, const float* w, const float* bias, float* out,
                    int64_t B, int64_t T, void* stream) {
  AG_CHECK_ARG(X && w && out && B > 0 && T > 0 && C > 0 && C % 4 == 0 && k > 0 && x_bs % 4 == 0, "ag_conv1out_fwd: bad args");
  AG_CHECK_ARG((reinterpret_cast<uintptr_t>(X) & 15) == 0 && (reinterpret_cast<uintptr_t>(w) & 15) == 0 &&
                   (!x_dtype || (C % 8 == 0 && x_bs % 8 == 0)), "ag_conv1out_fwd: unaligned");
  if (C <= 128 && k == 3) {
    const int S = 128;
    const int64_t spb = (T + S - 1) / S, strips = B * spb;
    const unsigned g = (unsigned)((strips + 7) / 8);
    if (x_dtype) conv1out_fwd_strip_kernel<1, 3><<<g, 256, 0, (cudaStream_t)stream>>>(X, x_bs, (int)C, w, bias, out, B, T, (int)spb, S);
    else conv1out_fwd_strip_kernel<0, 3><<<g, 256, 0, (cudaStream_t)stream>>>(X, x_bs, (int)C, w, bias, out, B, T, (int)spb, S);
    AG_LAUNCH_CHECK();
    return AG_OK;
  }
  const int64_t rows = B * T;
  int64_t grid = (rows + 7) / 8;
  const int64_t cap = (int64_t)sm_count() * 8;
  if (grid > cap) grid = cap;
  if (x_dtype) conv1out_fwd_kernel<1><<<(unsigned)grid, 256, (size_t)k * C * 4, (cudaStream_t)stream>>>(X, x_bs, (int)C, k, w, bias, out, B, T);
  else conv1out_fwd_kernel<0><<<(unsigned)grid, 256, (size_t)k * C * 4, (cudaStream_t)stream>>>(X, x_bs, (int)C, k, w, bias, out, B, T);
  AG_LAUNCH_CHECK();
  return AG_OK;
}

int ag_conv1out_dgrad(const float* g, const float* w, void* dX, int32_t dx_dtype, int64_t dx_bs, int64_t C, int32_t k, int64_t B, int64_t T,
                      void* stream) {
  AG_CHECK_ARG(g && w && dX && B > 0 && B < 65536 && T > 0 && C > 0 && C % 4 == 0 && C <= 512 && k > 0 && k <= 4 && dx_bs % 4 == 0,
               "ag_conv1out_dgrad: bad args");
  AG_CHECK_ARG((reinterpret_cast<uintptr_t>(dX) & (dx_dtype ? 7 : 15)) == 0 && (reinterpret_cast<uintptr_t>(w) & 15) == 0,
               "ag_conv1out_dgrad: unaligned");
  const int rpb = 64;
  dim3 grid((unsigned)((T + k - 1 + rpb - 1) / rpb), (unsigned)B);
  if (dx_dtype) conv1out_dgrad_kernel<4, 1><<<grid, 128, 0, (cudaStream_t)stream>>>(g, w, dX, dx_bs, (int)C, k, T, rpb);
  else conv1out_dgrad_kernel<4, 0><<<grid, 128, 0, (cudaStream_t)stream>>>(g, w, dX, dx_bs, (int)C, k, T, rpb);
  AG_LAUNCH_CHECK();
  return AG_OK;
}

int ag_conv1out_wgrad(const float* g, const void* X, int32_t x_dtype, int64_t x_bs, int64_t C, int32_t k, float* dw, int64_t B, int64_t T,
                      void* stream) {
  AG_CHECK_ARG(g && X && dw && B > 0 && B < 65536 && T > 0 && C > 0 && C % 4 == 0 && C <= 508 && k > 0 && k <= 4 && x_bs % 4 == 0,
               "ag_conv1out_wgrad: bad args");
  AG_CHECK_ARG((reinterpret_cast<uintptr_t>(X) & (x_dtype ? 7 : 15)) == 0, "ag_conv1out_wgrad: unaligned");
  if (C <= 128 && k == 3) {
    const int S = 128;
    const int64_t spb = (T + k - 1 + S - 1) / S, strips = B * spb;
    const unsigned gr = (unsigned)((strips + 7) / 8);
    if (x_dtype) conv1out_wgrad_strip_kernel<3, 1><<<gr, 256, 0, (cudaStream_t)stream>>>(g, X, x_bs, (int)C, dw, B, T, (int)spb, S);
    else conv1out_wgrad_strip_kernel<3, 0><<<gr, 256, 0, (cudaStream_t)stream>>>(g, X, x_bs, (int)C, dw, B, T, (int)spb, S);
    AG_LAUNCH_CHECK();
    return AG_OK;
  }
  const int rpb = 512;
  dim3 grid((unsigned)((T + k - 1 + rpb - 1) / rpb), (unsigned)B);
  if (x_dtype) conv1out_wgrad_kernel<4, 1><<<grid, 128, 0, (cudaStream_t)stream>>>(g, X, x_bs, (int)C, k, dw, T, rpb);
  else conv1out_wgrad_kernel<4, 0><<<grid, 128, 0, (cudaStream_t)stream>>>(g, X, x_bs, (int)C, k, dw, T, rpb);
  AG_LAUNCH_CHECK();
  return AG_OK;
}
int ag_conv1in_dgrad(const void* dy, int32_t dy_dtype, int64_t dy_bs, const float* w, float* dx, int64_t dx_ld, int32_t k, int32_t s,
                     int32_t p, int64_t C, int64_t B, int64_t T, int64_t Tin, void* stream) {
  AG_CHECK_ARG(dy && w && dx && B > 0 && B < 65536 && T > 0 && Tin > 0 && C > 0 && C % 4 == 0 && C <= 1024 && k > 0 && k <= 64 && s > 0 && p >= 0 &&
                   dy_bs % 4 == 0 && dx_ld >= Tin + p && (reinterpret_cast<uintptr_t>(dy) & (dy_dtype ? 7 : 15)) == 0,
               "ag_conv1in_dgrad: bad args");
  dim3 grid((unsigned)((dx_ld + 255) / 256), (unsigned)B);
  conv1in_dgrad_kernel<<<grid, 256, (size_t)k * C * 4, (cudaStream_t)stream>>>(dy, dy_dtype, dy_bs, w, dx, dx_ld, k, s, p, (int)C, T, Tin);
  AG_LAUNCH_CHECK();
  return AG_OK;
}

int ag_wcolsum(const float* g, const void* X, int32_t x_dtype, int64_t M, int64_t K, float* out, void* stream) {
  AG_CHECK_ARG(g && X && out && M > 0 && K > 0 && K % 4 == 0 && K <= 1024 && (reinterpret_cast<uintptr_t>(X) & (x_dtype ? 7 : 15)) == 0,
               "ag_wcolsum: bad args");
  int64_t blocks = (int64_t)sm_count() * 4;
  int64_t rows_per = (M + blocks - 1) / blocks;
  if (rows_per < 64) rows_per = 64;
  blocks = (M + rows_per - 1) / rows_per;
  wcolsum_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(g, X, x_dtype, M, (int)(K / 4), out, rows_per);
  AG_LAUNCH_CHECK();
  return AG_OK;
}

int ag_rowdot(const void* X, int32_t x_dtype, const float* w, const float* bias, int64_t M, int64_t K, float* out, void* stream) {
  AG_CHECK_ARG(X && w && out && M > 0 && K > 0 && K % 4 == 0 && K <= 1024 && (reinterpret_cast<uintptr_t>(X) & (x_dtype ? 7 : 15)) == 0,
               "ag_rowdot: bad args");
  int64_t blocks = (M + 7) / 8;
  const int64_t cap = (int64_t)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  rowdot_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(X, x_dtype, w, bias, M, (int)(K / 4), out);
  AG_LAUNCH_CHECK();
  return AG_OK;
}
}
