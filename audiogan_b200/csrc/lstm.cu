// Persistent recurrent kernels, fp32 ("fp32 mode": <=1e-5 parity with the reference's fp32 path).
//
// Replaces the per-frame Python loop of Generator.forward (audiogan.py:437-460: LSTMCell + proj +
// tanh + stopper + Bernoulli stop + host-side early exit) and NN.LSTM(bidirectional) under
// dynamic_rnn (audiogan.py:214-229, :498-503, :543), plus their BPTT.
//
// One cooperative launch per sequence pass.  CTA (dir, slice) owns HS hidden units of one
// direction: the 4*HS gate rows of [whh | wx] (forward) or the HS rows of [whh^T | wp^T ws]
// (backward) stay resident in shared memory for all T steps when they fit (RES), the state
// slices c / dc live in shared memory, and the only per-step global traffic is the [B, K]
// activation vector every CTA re-reads from L2 (cp.async.cg, double buffered) plus the saved
// gates / c / h.  Steps are separated by a per-direction grid barrier (monotonic counter).
// The feedback variant adds a second phase per step: x_t = tanh(wp h_t + bp), the stop logit,
// the Bernoulli stop draw from supplied uniforms and the device-side early-exit flag -- no
// host synchronisation per frame.
#include "common.cuh"

namespace ag {

constexpr int LT = 256;          // threads per CTA
constexpr int KC = 64;           // k-chunk (floats) staged per pipeline stage
constexpr int BTILE = 64;        // batches per tile: 8 warps x 8
constexpr int SLD = KC + 4;      // staged row stride (floats): 16-byte rows, conflict-free float4 reads

__device__ __forceinline__ void cp_async16(void* smem, const void* g) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// Monotonic-counter grid barrier over `n` co-resident CTAs (cooperative launch guarantees residency).
__device__ __forceinline__ void grid_barrier(unsigned* ctr, unsigned target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(ctr, 1u);
    while (ld_acquire_u32(ctr) < target) { __nanosleep(20); }
    __threadfence();
  }
  __syncthreads();
}

// The [B, K] activation matrix one phase consumes: row b is the concatenation of two strided rows.
struct Seg {
  const float* p0; int64_t s0; int n0;
  const float* p1; int64_t s1; int n1;
};

__device__ __forceinline__ void stage_chunk(float* stage, int buf, const Seg& sg, int k0, int K, int b0, int B) {
#pragma unroll
  for (int i = 0; i < (BTILE * KC / 4) / LT; ++i) {
    const int idx = threadIdx.x + LT * i;
    const int bl = idx >> 4, kk = (idx & 15) << 2;
    float* dst = stage + (buf * BTILE + bl) * SLD + kk;
    const int b = b0 + bl, k = k0 + kk;
    const float* src = nullptr;
    if (b < B && k < K) {
      if (k < sg.n0) { if (sg.p0) src = sg.p0 + b * sg.s0 + k; }
      else if (sg.p1) src = sg.p1 + b * sg.s1 + (k - sg.n0);
    }
    if (src) cp_async16(dst, src);
    else *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

// acc[rr][nb] += sum_k W[row(rr)][k] * in[b0 + warp*8 + nb][k]   (this lane's k-subset when KS > 1)
template <int RPT, int KS>
__device__ __forceinline__ void slice_gemm(float (&acc)[RPT][8], const float* const (&wrow)[RPT], const Seg& sg,
                                           int K, int b0, int B, float* stage) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int ks = (KS > 1) ? lane / (32 / KS) : 0;
  const int nch = (K + KC - 1) / KC;
  stage_chunk(stage, 0, sg, 0, K, b0, B);
  cp_async_commit();
  for (int c = 0; c < nch; ++c) {
    if (c + 1 < nch) {
      stage_chunk(stage, (c + 1) & 1, sg, (c + 1) * KC, K, b0, B);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const float* sb = stage + ((c & 1) * BTILE + w * 8) * SLD;
#pragma unroll 4
    for (int i = ks; i < KC / 4; i += KS) {
      const int k = c * KC + 4 * i;
      if (k < K) {
        float4 wv[RPT];
#pragma unroll
        for (int rr = 0; rr < RPT; ++rr) wv[rr] = *reinterpret_cast<const float4*>(wrow[rr] + k);
#pragma unroll
        for (int nb = 0; nb < 8; ++nb) {
          const float4 a = *reinterpret_cast<const float4*>(sb + nb * SLD + 4 * i);
#pragma unroll
          for (int rr = 0; rr < RPT; ++rr) {
            acc[rr][nb] = fmaf(wv[rr].x, a.x, acc[rr][nb]);
            acc[rr][nb] = fmaf(wv[rr].y, a.y, acc[rr][nb]);
            acc[rr][nb] = fmaf(wv[rr].z, a.z, acc[rr][nb]);
            acc[rr][nb] = fmaf(wv[rr].w, a.w, acc[rr][nb]);
          }
        }
      }
    }
    __syncthreads();
  }
  if (KS > 1) {
#pragma unroll
    for (int off = 32 / KS; off < 32; off <<= 1)
#pragma unroll
      for (int rr = 0; rr < RPT; ++rr)
#pragma unroll
        for (int nb = 0; nb < 8; ++nb) acc[rr][nb] += __shfl_xor_sync(0xffffffffu, acc[rr][nb], off);
  }
}

__host__ __device__ inline int pad_ld(int K) { return (K % 8 == 0) ? K + 4 : K; }   // K % 4 == 0 -> ld % 8 == 4

// out[r] (r < nr <= 4) = <W2s[r], vec> with the K range split over the warp's lanes; every lane gets the sums.
__device__ __forceinline__ void warp_rows_dot(float (&out)[4], const float* Wrows, int ldw, int nr, const float* vec, int K4) {
  const int lane = threadIdx.x & 31;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  for (int k4 = lane; k4 < K4; k4 += 32) {
    const float4 v = __ldcg(reinterpret_cast<const float4*>(vec) + k4);
    const float4 w0 = *reinterpret_cast<const float4*>(Wrows + 4 * k4);
    a0 += w0.x * v.x + w0.y * v.y + w0.z * v.z + w0.w * v.w;
    if (nr > 1) { const float4 q = *reinterpret_cast<const float4*>(Wrows + ldw + 4 * k4); a1 += q.x * v.x + q.y * v.y + q.z * v.z + q.w * v.w; }
    if (nr > 2) { const float4 q = *reinterpret_cast<const float4*>(Wrows + 2 * ldw + 4 * k4); a2 += q.x * v.x + q.y * v.y + q.z * v.z + q.w * v.w; }
    if (nr > 3) { const float4 q = *reinterpret_cast<const float4*>(Wrows + 3 * ldw + 4 * k4); a3 += q.x * v.x + q.y * v.y + q.z * v.z + q.w * v.w; }
  }
  out[0] = warp_sum(a0); out[1] = warp_sum(a1); out[2] = warp_sum(a2); out[3] = warp_sum(a3);
}

// =====================================================================================  forward
template <int HS, bool FB, bool RES>
__global__ void __launch_bounds__(LT, 1) lstm_fwd_kernel(const ag_lstm_desc d, const int ncta_dir, const int PR) {
  constexpr int ROWS = 4 * HS;
  constexpr int RL = ROWS < 32 ? ROWS : 32;
  constexpr int KS = 32 / RL;
  constexpr int RPT = ROWS / RL;
  constexpr int IPT = (BTILE * HS) / LT;            // cell-update items per thread per batch tile
  extern __shared__ __align__(16) float smem[];
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int H = d.H, F = FB ? d.F : 0, K1 = H + F, ndir = d.ndir, B = d.B, T = d.T, Tcap = d.Tcap;
  const int dir = blockIdx.x / ncta_dir, cta = blockIdx.x % ncta_dir, j0 = cta * HS;
  const int ldw = pad_ld(K1), ld2 = pad_ld(H);

  float* Ws = smem;
  float* stage = Ws + (RES ? ROWS * ldw : 0);
  float* gs = stage + 2 * BTILE * SLD;
  float* cs = gs + BTILE * ROWS;
  float* W2s = cs + B * HS;
  int* gen = reinterpret_cast<int*>(W2s + (FB ? PR * ld2 : 0));
  int* cnt = gen + B;

  // local row lr = jj*4 + q  <->  global gate row q*H + j0 + jj
  const float* w1d = d.w1 + (int64_t)dir * 4 * H * K1;
  if (RES) {
    const int K4 = K1 / 4;
    for (int idx = tid; idx < ROWS * K4; idx += LT) {
      const int lr = idx / K4, k4 = idx - lr * K4;
      const int grow = (lr & 3) * H + j0 + (lr >> 2);
      *reinterpret_cast<float4*>(Ws + lr * ldw + 4 * k4) = *reinterpret_cast<const float4*>(w1d + (int64_t)grow * K1 + 4 * k4);
    }
  }
  const float* wrow[RPT];
#pragma unroll
  for (int rr = 0; rr < RPT; ++rr) {
    const int lr = (lane % RL) + RL * rr;
    const int grow = (lr & 3) * H + j0 + (lr >> 2);
    wrow[rr] = RES ? (Ws + lr * ldw) : (w1d + (int64_t)grow * K1);
  }
  for (int i = tid; i < B * HS; i += LT) cs[i] = 0.f;
  // phase-2 rows owned by this CTA (feedback only; ndir == 1)
  int p0 = 0, np = 0;
  bool owns_logit = false;
  if (FB) {
    p0 = blockIdx.x * PR;
    np = min(PR, F + 1 - p0);
    if (np < 0) np = 0;
    owns_logit = (np > 0) && (p0 + np == F + 1);
    for (int idx = tid; idx < np * (H / 4); idx += LT) {
      const int r = idx / (H / 4), k4 = idx - r * (H / 4);
      *reinterpret_cast<float4*>(W2s + r * ld2 + 4 * k4) = *reinterpret_cast<const float4*>(d.w2 + (int64_t)(p0 + r) * H + 4 * k4);
    }
    for (int b = tid; b < B; b += LT) { gen[b] = 1; cnt[b] = 0; }
  }
  __syncthreads();

  unsigned* bar = d.barrier + dir;
  unsigned nbar = 0;
  const int64_t hstr = (int64_t)(Tcap + 2) * ndir * H;   // hbuf batch stride
  const int64_t gstr = (int64_t)Tcap * ndir * 4 * H;     // pre / gates batch stride
  const int64_t cstr = (int64_t)Tcap * ndir * H;         // cbuf batch stride
  int steps_run = T;

  for (int s = 0; s < T; ++s) {
    const int t = dir ? (T - 1 - s) : s;
    const int prow = dir ? (t + 2) : t;                  // hbuf row holding the previous h
    Seg sg;
    sg.p0 = d.hbuf + (int64_t)prow * ndir * H + dir * H; sg.s0 = hstr; sg.n0 = H;
    sg.p1 = FB ? (d.xbuf + (int64_t)t * F) : nullptr; sg.s1 = (int64_t)(Tcap + 1) * F; sg.n1 = F;

    for (int b0 = 0; b0 < B; b0 += BTILE) {
      // prefetch this tile's input projections (independent of the recurrence)
      float pre[IPT > 0 ? IPT : 1][4];
#pragma unroll
      for (int ii = 0; ii < IPT; ++ii) {
        const int it = tid + LT * ii, bl = it / HS, jj = it - bl * HS, b = b0 + bl;
#pragma unroll
        for (int q = 0; q < 4; ++q)
          pre[ii][q] = (b < B) ? d.pre[b * gstr + (int64_t)t * ndir * 4 * H + dir * 4 * H + q * H + j0 + jj] : 0.f;
      }
      float acc[RPT][8];
#pragma unroll
      for (int rr = 0; rr < RPT; ++rr)
#pragma unroll
        for (int nb = 0; nb < 8; ++nb) acc[rr][nb] = 0.f;
      slice_gemm<RPT, KS>(acc, wrow, sg, K1, b0, B, stage);
      if (lane < RL) {
#pragma unroll
        for (int rr = 0; rr < RPT; ++rr)
#pragma unroll
          for (int nb = 0; nb < 8; ++nb) gs[(w * 8 + nb) * ROWS + lane + RL * rr] = acc[rr][nb];
      }
      __syncthreads();
#pragma unroll
      for (int ii = 0; ii < IPT; ++ii) {
        const int it = tid + LT * ii, bl = it / HS, jj = it - bl * HS, b = b0 + bl;
        if (b >= B) continue;
        const int j = j0 + jj;
        const float4 a = *reinterpret_cast<const float4*>(gs + bl * ROWS + jj * 4);
        const bool valid = !d.len || t < d.len[b];
        float gi = 0.f, gf = 0.f, gg = 0.f, go = 0.f, c = 0.f, h = 0.f;
        if (valid) {
          gi = sigmoidf_(a.x + pre[ii][0]);
          gf = sigmoidf_(a.y + pre[ii][1]);
          gg = tanhf(a.z + pre[ii][2]);
          go = sigmoidf_(a.w + pre[ii][3]);
          c = gf * cs[b * HS + jj] + gi * gg;
          h = go * tanhf(c);
          cs[b * HS + jj] = c;
        }
        d.hbuf[b * hstr + (int64_t)(t + 1) * ndir * H + dir * H + j] = h;
        if (d.cbuf) d.cbuf[b * cstr + (int64_t)t * ndir * H + dir * H + j] = c;
        if (d.gates) {
          float* gp = d.gates + b * gstr + (int64_t)t * ndir * 4 * H + dir * 4 * H + j;
          gp[0] = gi; gp[H] = gf; gp[2 * H] = gg; gp[3 * H] = go;
        }
      }
      __syncthreads();
    }
    grid_barrier(bar, (++nbar) * ncta_dir);

    if (FB) {
      // phase 2: x_t = tanh(wp h_t + bp), logit = ws h_t + bs, stop draw, early-exit flag
      if (np > 0) {
        for (int b = w; b < B; b += LT / 32) {
          const float* hrow = d.hbuf + b * hstr + (int64_t)(t + 1) * H;
          for (int pp = 0; pp < np; pp += 4) {
            float o[4];
            const int nr = min(4, np - pp);
            warp_rows_dot(o, W2s + pp * ld2, ld2, nr, hrow, H / 4);
            if (lane == 0) {
              for (int r = 0; r < nr; ++r) {
                const int p = p0 + pp + r;
                const float v = o[r] + d.b2[p];
                if (p < F) {
                  d.xbuf[(b * (int64_t)(Tcap + 1) + t + 1) * F + p] = tanhf(v);
                } else {
                  if (d.sbuf) d.sbuf[b * (int64_t)Tcap + t] = v;
                  const int stop = (d.u && d.u[b * (int64_t)Tcap + t] < sigmoidf_(v)) ? 1 : 0;
                  if (d.stop) d.stop[b * (int64_t)Tcap + t] = stop;
                  if (gen[b]) cnt[b] += 1;
                  if (stop) gen[b] = 0;
                }
              }
            }
          }
        }
      }
      if (owns_logit) {
        __syncthreads();
        int g = 0;
        for (int b = tid; b < B; b += LT) g |= gen[b];
        const int any = __syncthreads_or(g);
        if (!any && tid == 0) *reinterpret_cast<volatile int*>(d.t_end) = t + 1;
      }
      grid_barrier(bar, (++nbar) * ncta_dir);
      const int te = *reinterpret_cast<volatile int*>(d.t_end);
      if (te != 0) { steps_run = te; break; }
    }
  }
  if (FB && owns_logit) {
    __syncthreads();
    for (int b = tid; b < B; b += LT) if (d.glen) d.glen[b] = cnt[b];
    if (tid == 0 && steps_run == T) *reinterpret_cast<volatile int*>(d.t_end) = T;
  }
}

// ====================================================================================  backward
template <int HS, bool FB, bool RES>
__global__ void __launch_bounds__(LT, 1) lstm_bwd_kernel(const ag_lstm_desc d, const int ncta_dir, const int PR) {
  constexpr int RL = HS;            // 4, 8 or 16 rows -> k split over the rest of the warp
  constexpr int KS = 32 / RL;
  constexpr int IPT = (BTILE * HS) / LT;
  extern __shared__ __align__(16) float smem[];
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int H = d.H, F = FB ? d.F : 0, FP = FB ? ((F + 1 + 3) / 4) * 4 : 0, K = 4 * H + FP;
  const int ndir = d.ndir, B = d.B, T = d.T, Tcap = d.Tcap;
  const int dir = blockIdx.x / ncta_dir, cta = blockIdx.x % ncta_dir, j0 = cta * HS;
  const int ldw = pad_ld(K), ldx = pad_ld(4 * H);

  float* Ws = smem;
  float* stage = Ws + (RES ? HS * ldw : 0);
  float* dhs = stage + 2 * BTILE * SLD;
  float* dcs = dhs + BTILE * HS;
  float* Wxs = dcs + B * HS;

  const float* w1d = d.w1t + ((int64_t)dir * H + j0) * K;
  if (RES) {
    const int K4 = K / 4;
    for (int idx = tid; idx < HS * K4; idx += LT) {
      const int lr = idx / K4, k4 = idx - lr * K4;
      *reinterpret_cast<float4*>(Ws + lr * ldw + 4 * k4) = *reinterpret_cast<const float4*>(w1d + (int64_t)lr * K + 4 * k4);
    }
  }
  const float* wrow[1];
  wrow[0] = RES ? (Ws + (lane % RL) * ldw) : (w1d + (int64_t)(lane % RL) * K);
  for (int i = tid; i < B * HS; i += LT) dcs[i] = 0.f;
  int p0 = 0, np = 0;
  if (FB) {
    p0 = blockIdx.x * PR;
    np = min(PR, F - p0);
    if (np < 0) np = 0;
    for (int idx = tid; idx < np * H; idx += LT) {       // H float4 per row of wxt (4H floats)
      const int r = idx / H, k4 = idx - r * H;
      *reinterpret_cast<float4*>(Wxs + r * ldx + 4 * k4) = *reinterpret_cast<const float4*>(d.wxt + (int64_t)(p0 + r) * 4 * H + 4 * k4);
    }
  }
  __syncthreads();

  unsigned* bar = d.barrier + dir;
  unsigned nbar = 0;
  const int64_t gstr = (int64_t)Tcap * ndir * 4 * H;
  const int64_t cstr = (int64_t)Tcap * ndir * H;

  for (int s = 0; s < T; ++s) {
    const int t = dir ? s : (T - 1 - s);                 // reverse of the forward order
    const int tn = dir ? (t - 1) : (t + 1);              // the step processed just before this one
    const bool has_next = s > 0;

    if (FB) {
      // phase A: dpx[b,t,p] = (dx_ext + wx^T dgates_{t+1})[p] * (1 - x_t[p]^2); column F = ds_ext
      if (np > 0) {
        for (int b = w; b < B; b += LT / 32) {
          const float* dgrow = d.dgates + b * gstr + (int64_t)tn * 4 * H;
          for (int pp = 0; pp < np; pp += 4) {
            float o[4] = {0.f, 0.f, 0.f, 0.f};
            const int nr = min(4, np - pp);
            if (has_next) warp_rows_dot(o, Wxs + pp * ldx, ldx, nr, dgrow, H);
            if (lane == 0) {
              for (int r = 0; r < nr; ++r) {
                const int p = p0 + pp + r;
                float dx = o[r];
                if (d.dx_ext) dx += d.dx_ext[(b * (int64_t)Tcap + t) * F + p];
                const float x = d.xbuf[(b * (int64_t)(Tcap + 1) + t + 1) * F + p];
                d.dpx[(b * (int64_t)Tcap + t) * FP + p] = dx * (1.f - x * x);
              }
            }
          }
        }
      }
      if (blockIdx.x == gridDim.x - 1) {
        for (int b = tid; b < B; b += LT) {
          float* q = d.dpx + (b * (int64_t)Tcap + t) * FP;
          q[F] = d.ds_ext ? d.ds_ext[b * (int64_t)Tcap + t] : 0.f;
          for (int p = F + 1; p < FP; ++p) q[p] = 0.f;
        }
      }
      grid_barrier(bar, (++nbar) * ncta_dir);
    }

    // phase B: dh = dh_ext + whh^T dgates_next (+ wp^T dpx_t + ws ds_t), then the cell backward
    Seg sg;
    sg.p0 = has_next ? (d.dgates + (int64_t)tn * ndir * 4 * H + dir * 4 * H) : nullptr; sg.s0 = gstr; sg.n0 = 4 * H;
    sg.p1 = FB ? (d.dpx + (int64_t)t * FP) : nullptr; sg.s1 = (int64_t)Tcap * FP; sg.n1 = FP;
    for (int b0 = 0; b0 < B; b0 += BTILE) {
      float acc[1][8];
#pragma unroll
      for (int nb = 0; nb < 8; ++nb) acc[0][nb] = 0.f;
      if (has_next || FB) slice_gemm<1, KS>(acc, wrow, sg, K, b0, B, stage);
      if (lane < RL) {
#pragma unroll
        for (int nb = 0; nb < 8; ++nb) dhs[(w * 8 + nb) * HS + lane] = acc[0][nb];
      }
      __syncthreads();
#pragma unroll
      for (int ii = 0; ii < IPT; ++ii) {
        const int it = tid + LT * ii, bl = it / HS, jj = it - bl * HS, b = b0 + bl;
        if (b >= B) continue;
        const int j = j0 + jj;
        const int Lb = d.len ? min(d.len[b], T) : T;
        float* dg = d.dgates + b * gstr + (int64_t)t * ndir * 4 * H + dir * 4 * H + j;
        if (t >= Lb) {
          dg[0] = 0.f; dg[H] = 0.f; dg[2 * H] = 0.f; dg[3 * H] = 0.f;
          continue;
        }
        const float* gp = d.gates + b * gstr + (int64_t)t * ndir * 4 * H + dir * 4 * H + j;
        const float gi = gp[0], gf = gp[H], gg = gp[2 * H], go = gp[3 * H];
        const float c = d.cbuf[b * cstr + (int64_t)t * ndir * H + dir * H + j];
        const int tp = dir ? (t + 1) : (t - 1);
        const bool has_prev = dir ? (tp < Lb) : (tp >= 0);
        const float cprev = has_prev ? d.cbuf[b * cstr + (int64_t)tp * ndir * H + dir * H + j] : 0.f;
        float dh = dhs[bl * HS + jj];
        if (d.dh_ext) dh += d.dh_ext[b * (d.dh_ext_bs ? d.dh_ext_bs : cstr) + (int64_t)t * ndir * H + dir * H + j];
        const float tc = tanhf(c);
        const float dc = dcs[b * HS + jj] + dh * go * (1.f - tc * tc);
        dcs[b * HS + jj] = dc * gf;
        dg[0] = dc * gg * gi * (1.f - gi);
        dg[H] = dc * cprev * gf * (1.f - gf);
        dg[2 * H] = dc * gi * (1.f - gg * gg);
        dg[3 * H] = dh * tc * go * (1.f - go);
      }
      __syncthreads();
    }
    grid_barrier(bar, (++nbar) * ncta_dir);
  }
}

// ---------------------------------------------------------------------------------- host side
struct Plan { int HS, ncta_dir, PR; bool res; size_t smem; };

static int pick_hs(int H, int ndir) {
  const int hs_opts[3] = {4, 8, 16};
  for (int i = 0; i < 3; ++i) {
    const int hs = hs_opts[i];
    if (H % hs == 0 && (int64_t)ndir * (H / hs) <= sm_count()) return hs;
  }
  return 0;
}

static size_t fwd_smem(const ag_lstm_desc* d, int HS, int PR, bool res) {
  const int F = d->F, K1 = d->H + F;
  size_t fl = (res ? (size_t)4 * HS * pad_ld(K1) : 0) + 2 * BTILE * SLD + (size_t)BTILE * 4 * HS + (size_t)d->B * HS +
              (F > 0 ? (size_t)PR * pad_ld(d->H) : 0);
  return fl * 4 + (size_t)2 * d->B * 4 + 16;
}
static size_t bwd_smem(const ag_lstm_desc* d, int HS, int PR, bool res) {
  const int F = d->F, FP = F > 0 ? ((F + 1 + 3) / 4) * 4 : 0, K = 4 * d->H + FP;
  size_t fl = (res ? (size_t)HS * pad_ld(K) : 0) + 2 * BTILE * SLD + (size_t)BTILE * HS + (size_t)d->B * HS +
              (F > 0 ? (size_t)PR * pad_ld(4 * d->H) : 0);
  return fl * 4 + 16;
}

static int check_lstm(const ag_lstm_desc* d, const char* who, bool bwd) {
  AG_CHECK_ARG(d, "%s: null descriptor", who);
  AG_CHECK_ARG(d->B > 0 && d->T > 0 && d->Tcap >= d->T && d->H > 0 && d->H % 4 == 0, "%s: bad B/T/Tcap/H", who);
  AG_CHECK_ARG(d->ndir == 1 || d->ndir == 2, "%s: ndir must be 1 or 2", who);
  AG_CHECK_ARG(d->F >= 0 && d->F % 4 == 0 && (d->F == 0 || d->ndir == 1), "%s: bad F", who);
  AG_CHECK_ARG(d->barrier, "%s: null barrier", who);
  if (!bwd) {
    AG_CHECK_ARG(d->pre && d->w1 && d->hbuf, "%s: null pre/w1/hbuf", who);
    if (d->F > 0) AG_CHECK_ARG(d->w2 && d->b2 && d->xbuf && d->t_end, "%s: feedback needs w2,b2,xbuf,t_end", who);
  } else {
    AG_CHECK_ARG(d->w1t && d->gates && d->cbuf && d->dgates, "%s: null w1t/gates/cbuf/dgates", who);
    if (d->F > 0) AG_CHECK_ARG(d->wxt && d->xbuf && d->dpx, "%s: feedback needs wxt,xbuf,dpx", who);
  }
  return AG_OK;
}

template <typename KernT>
static int launch_coop(KernT kern, const ag_lstm_desc* d, const Plan& p, cudaStream_t s) {
  AG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem));
  ag_lstm_desc dd = *d;
  int ncta_dir = p.ncta_dir, PR = p.PR;
  void* args[3] = {&dd, &ncta_dir, &PR};
  AG_CUDA(cudaLaunchCooperativeKernel((const void*)kern, dim3((unsigned)(p.ncta_dir * d->ndir)), dim3(LT), args, p.smem, s));
  return AG_OK;
}

#define AG_LSTM_DISPATCH(KERN)                                                                      \
  do {                                                                                              \
    const bool fb = d->F > 0;                                                                       \
    if (p.HS == 4) {                                                                                \
      if (fb) return p.res ? launch_coop(KERN<4, true, true>, d, p, s) : launch_coop(KERN<4, true, false>, d, p, s);     \
      return p.res ? launch_coop(KERN<4, false, true>, d, p, s) : launch_coop(KERN<4, false, false>, d, p, s);           \
    } else if (p.HS == 8) {                                                                         \
      if (fb) return p.res ? launch_coop(KERN<8, true, true>, d, p, s) : launch_coop(KERN<8, true, false>, d, p, s);     \
      return p.res ? launch_coop(KERN<8, false, true>, d, p, s) : launch_coop(KERN<8, false, false>, d, p, s);           \
    } else {                                                                                        \
      if (fb) return p.res ? launch_coop(KERN<16, true, true>, d, p, s) : launch_coop(KERN<16, true, false>, d, p, s);   \
      return p.res ? launch_coop(KERN<16, false, true>, d, p, s) : launch_coop(KERN<16, false, false>, d, p, s);         \
    }                                                                                               \
  } while (0)

static int run_fwd(const ag_lstm_desc* d, const Plan& p, cudaStream_t s) { AG_LSTM_DISPATCH(lstm_fwd_kernel); }
static int run_bwd(const ag_lstm_desc* d, const Plan& p, cudaStream_t s) { AG_LSTM_DISPATCH(lstm_bwd_kernel); }

}  // namespace ag

using namespace ag;
extern "C" {

int ag_lstm_fwd(const ag_lstm_desc* d, void* stream) {
  int rc = check_lstm(d, "ag_lstm_fwd", false);
  if (rc) return rc;
  Plan p;
  p.HS = pick_hs(d->H, d->ndir);
  AG_CHECK_ARG(p.HS > 0, "ag_lstm_fwd: H=%d (ndir %d) does not map onto %d SMs", d->H, d->ndir, sm_count());
  p.ncta_dir = d->H / p.HS;
  const int ncta = p.ncta_dir * d->ndir;
  p.PR = d->F > 0 ? (d->F + 1 + ncta - 1) / ncta : 0;
  p.res = true;
  p.smem = fwd_smem(d, p.HS, p.PR, true);
  if (p.smem > (size_t)smem_optin()) { p.res = false; p.smem = fwd_smem(d, p.HS, p.PR, false); }
  AG_CHECK_ARG(p.smem <= (size_t)smem_optin(), "ag_lstm_fwd: needs %zu B of shared memory", p.smem);
  cudaStream_t s = (cudaStream_t)stream;
  AG_CUDA(cudaMemsetAsync(d->barrier, 0, 8 * sizeof(unsigned), s));
  if (d->F > 0) AG_CUDA(cudaMemsetAsync(d->t_end, 0, sizeof(int), s));
  return run_fwd(d, p, s);
}

int ag_lstm_bwd(const ag_lstm_desc* d, void* stream) {
  int rc = check_lstm(d, "ag_lstm_bwd", true);
  if (rc) return rc;
  Plan p;
  p.HS = pick_hs(d->H, d->ndir);
  AG_CHECK_ARG(p.HS > 0, "ag_lstm_bwd: H=%d (ndir %d) does not map onto %d SMs", d->H, d->ndir, sm_count());
  p.ncta_dir = d->H / p.HS;
  const int ncta = p.ncta_dir * d->ndir;
  p.PR = d->F > 0 ? (d->F + ncta - 1) / ncta : 0;
  p.res = true;
  p.smem = bwd_smem(d, p.HS, p.PR, true);
  if (p.smem > (size_t)smem_optin()) { p.res = false; p.smem = bwd_smem(d, p.HS, p.PR, false); }
  AG_CHECK_ARG(p.smem <= (size_t)smem_optin(), "ag_lstm_bwd: needs %zu B of shared memory", p.smem);
  cudaStream_t s = (cudaStream_t)stream;
  AG_CUDA(cudaMemsetAsync(d->barrier, 0, 8 * sizeof(unsigned), s));
  return run_bwd(d, p, s);
}
}
