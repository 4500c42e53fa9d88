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
#include "tc_common.cuh"
#include <algorithm>

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
    while (ld_acquire_u32(ctr) < target) { }
    __threadfence();
  }
  __syncthreads();
}

// The [B, K] activation matrix one phase consumes: row b is the concatenation of two strided rows.
struct Seg {
  const float* p0; int64_t s0; int n0;
  const float* p1; int64_t s1; int n1;
  const __nv_bfloat16* q0; const __nv_bfloat16* q1;   // bf16 shadow copies (same strides), bf16 mode only
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
  // every CTA walks the K chunks in a different rotation: right after a grid barrier all CTAs want the same [B, K]
  // vector, and identical request streams serialise on the same L2 lines
  const int rot = (int)((blockIdx.x * 5u) % (unsigned)nch);
  auto kof = [&](int c) { int cc = c + rot; if (cc >= nch) cc -= nch; return cc * KC; };
  stage_chunk(stage, 0, sg, kof(0), K, b0, B);
  cp_async_commit();
  for (int c = 0; c < nch; ++c) {
    if (c + 1 < nch) {
      stage_chunk(stage, (c + 1) & 1, sg, kof(c + 1), K, b0, B);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const float* sb = stage + ((c & 1) * BTILE + w * 8) * SLD;
#pragma unroll 4
    for (int i = ks; i < KC / 4; i += KS) {
      const int k = kof(c) + 4 * i;
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


// Batched twin of warp_rows_dot: NB vectors at once (all their loads in flight together), nr <= 4 weight rows.
// out[i][r] = <Wrows[r], vec[i]>.  fp32 weights/vectors; vec[i] == nullptr gives zeros.
template <int NB>
__device__ __forceinline__ void warp_rows_dot_nb(float (&out)[NB][4], const float* Wrows, int ldw, int nr,
                                                 const float* const (&vec)[NB], int K4) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int i = 0; i < NB; ++i) { out[i][0] = 0.f; out[i][1] = 0.f; out[i][2] = 0.f; out[i][3] = 0.f; }
  for (int k4 = lane; k4 < K4; k4 += 32) {
    float4 v[NB];
#pragma unroll
    for (int i = 0; i < NB; ++i)
      v[i] = vec[i] ? __ldcg(reinterpret_cast<const float4*>(vec[i]) + k4) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      if (r < nr) {
        const float4 q = *reinterpret_cast<const float4*>(Wrows + r * ldw + 4 * k4);
#pragma unroll
        for (int i = 0; i < NB; ++i) out[i][r] += q.x * v[i].x + q.y * v[i].y + q.z * v[i].z + q.w * v[i].w;
      }
    }
  }
#pragma unroll
  for (int i = 0; i < NB; ++i)
#pragma unroll
    for (int r = 0; r < 4; ++r) out[i][r] = warp_sum(out[i][r]);
}
// bf16 weights (smem) and bf16 vectors (global, L2)
template <int NB>
__device__ __forceinline__ void warp_rows_dot16_nb(float (&out)[NB][4], const __nv_bfloat16* Wrows, int ldw, int nr,
                                                   const __nv_bfloat16* const (&vec)[NB], int K8) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int i = 0; i < NB; ++i) { out[i][0] = 0.f; out[i][1] = 0.f; out[i][2] = 0.f; out[i][3] = 0.f; }
  for (int k8 = lane; k8 < K8; k8 += 32) {
    uint4 v[NB];
#pragma unroll
    for (int i = 0; i < NB; ++i)
      v[i] = vec[i] ? __ldcg(reinterpret_cast<const uint4*>(vec[i]) + k8) : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      if (r < nr) {
        const uint4 q = *reinterpret_cast<const uint4*>(Wrows + r * ldw + 8 * k8);
        const __nv_bfloat162* qp = reinterpret_cast<const __nv_bfloat162*>(&q);
        float2 y[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) y[e] = __bfloat1622float2(qp[e]);
#pragma unroll
        for (int i = 0; i < NB; ++i) {
          const __nv_bfloat162* vp = reinterpret_cast<const __nv_bfloat162*>(&v[i]);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 x = __bfloat1622float2(vp[e]);
            out[i][r] = fmaf(x.x, y[e].x, out[i][r]);
            out[i][r] = fmaf(x.y, y[e].y, out[i][r]);
          }
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < NB; ++i)
#pragma unroll
    for (int r = 0; r < 4; ++r) out[i][r] = warp_sum(out[i][r]);
}

struct PhaseClock {
  long long t0, acc[8];
  __device__ __forceinline__ void start() { t0 = clock64(); }
  __device__ __forceinline__ void lap(int i) { const long long t = clock64(); acc[i] += t - t0; t0 = t; }
};

// ------------------------------------------------------------------------------------ bf16 tensor-core variant
// bf16 mode: the per-step [B, K] operand is read from bf16 shadow buffers through a 4-deep cp.async ring (the
// fp32 variant's 2-deep ring exposes the L2 latency of every 64-column chunk), the resident weight slice is bf16,
// and the products run on mma.sync.m16n8k16 (bf16 x bf16 -> fp32).  The recurrent state stays fp32.
constexpr int NST_MIN = 4;        // cp.async ring depth: 4, 8 or 12 stages, the deepest that fits (runtime `nst`)
constexpr int SLD16 = KC + 8;     // staged bf16 row stride at kc = 64 (elements): 144 B rows, conflict-free reads
constexpr int SLOT16 = BTILE * SLD16;   // ring slot size in elements (>= rb * (kc + 8) for every (rb, kc) pair)

__host__ __device__ inline int pad_ld16(int K) {          // >= K rounded to a whole sub-chunk, == 8 (mod 64): conflict-free
  return (K + KC - 1) / KC * KC + 8;
}
__host__ __device__ inline int ring_kc(int bper) {         // columns per ring slot for a CTA owning `bper` batches
  return bper <= 16 ? 256 : (bper <= 32 ? 128 : 64);
}

// Per-thread view of the [B, K] bf16 operand for one (step, batch tile).  A ring slot holds `rb` batch rows x `kc`
// columns with rb * kc == 4096 (rb = 64/32/16 -> kc = 64/128/256): CTAs that own a narrow batch range stage wide
// chunks, so the number of block-wide syncs per step shrinks with the batch split.  Each thread always copies the
// same two (row, 16-byte column) cells of every chunk: row pointers are resolved once per step.
struct Stager16 {
  const __nv_bfloat16* r0[2];   // row base inside segment 0 (nullptr = zeros)
  const __nv_bfloat16* r1[2];   // row base inside segment 1, already shifted by -n0
  int dst[2];                   // element offset inside a ring slot
  int kk[2], n0, K;
  __device__ __forceinline__ void init(const Seg& sg, int b0, int B, int K_, int kc) {
    n0 = sg.n0;
    K = K_;
    const int cpr = kc >> 3, sld = kc + 8;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int idx = threadIdx.x + LT * i;
      const int bl = idx / cpr, b = b0 + bl;
      kk[i] = (idx - bl * cpr) << 3;
      dst[i] = bl * sld + kk[i];
      const bool ok = b < B;
      r0[i] = (ok && sg.q0) ? sg.q0 + b * sg.s0 : nullptr;
      r1[i] = (ok && sg.q1) ? sg.q1 + b * sg.s1 - sg.n0 : nullptr;
    }
  }
  __device__ __forceinline__ void issue(__nv_bfloat16* slot, int k0) const {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int k = k0 + kk[i];
      const __nv_bfloat16* base = (k < n0) ? r0[i] : r1[i];
      if (base != nullptr && k < K) cp_async16(slot + dst[i], base + k);
      else *reinterpret_cast<uint4*>(slot + dst[i]) = make_uint4(0u, 0u, 0u, 0u);
    }
  }
};

__device__ __forceinline__ void mma_bf16(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

__device__ __forceinline__ void cp_async_wait_dyn(int pending) {
  if (pending >= 10) cp_async_wait<10>();
  else if (pending >= 6) cp_async_wait<6>();
  else if (pending >= 2) cp_async_wait<2>();
  else if (pending == 1) cp_async_wait<1>();
  else cp_async_wait<0>();
}

// MT 16-row tiles of the weight slice x up to 8 batch tiles of 8.  Warp w owns row tile w % MT and the batch tiles
// w / MT + (8 / MT) * i, i < MT, skipping those past the last valid batch.  acc[i] is an m16n8 fragment:
// [0],[1] -> row g, batches 2t, 2t+1; [2],[3] -> row g+8.  All fragment loads of a 64-column sub-chunk are issued
// before its MMAs; `kc` columns (rb = 4096 / kc rows) per ring slot, one block-wide sync per slot.
template <int MT>
__device__ __forceinline__ void slice_gemm_mma(float (&acc)[MT][4], const __nv_bfloat16* Ws, int ldw, const Seg& sg, int K,
                                               int b0, int B, __nv_bfloat16* stage, const int NST, const int kc,
                                               PhaseClock* pc = nullptr) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int mt = w % MT, nt0 = w / MT;
  constexpr int NSTEP = 8 / MT;                         // batch-tile stride between a warp's tiles
  const int rb = 4096 / kc, sld = kc + 8, slot = rb * sld;
  const int ntv = (min(min(B - b0, BTILE), rb) + 7) >> 3;   // valid batch tiles held by a slot
  const int nch = (K + kc - 1) / kc;
  const int rot = (int)((blockIdx.x * 5u) % (unsigned)nch);     // per-CTA chunk rotation (see slice_gemm)
  auto kof = [&](int c) { int cc = c + rot; if (cc >= nch) cc -= nch; return cc * kc; };
  Stager16 st;
  st.init(sg, b0, B, K, kc);
  for (int s = 0; s < NST - 1; ++s) {
    if (s < nch) st.issue(stage + s * slot, kof(s));
    cp_async_commit();
  }
  const __nv_bfloat16* wa = Ws + (mt * 16 + g) * ldw + 2 * t;
  // independent mma.sync dependency chains per warp: MT tiles x NCH accumulator sets (k16 steps interleaved)
  constexpr int NCH = MT >= 4 ? 1 : (MT == 2 ? 2 : 4);
  float part[NCH][MT][4];
#pragma unroll
  for (int q = 0; q < NCH; ++q)
#pragma unroll
    for (int i = 0; i < MT; ++i) { part[q][i][0] = 0.f; part[q][i][1] = 0.f; part[q][i][2] = 0.f; part[q][i][3] = 0.f; }
  for (int c = 0; c < nch; ++c) {
    long long q0 = 0;
    if (pc) q0 = clock64();
    cp_async_wait_dyn(NST - 2);
    __syncthreads();
    if (pc) { const long long q1 = clock64(); pc->acc[5] += q1 - q0; q0 = q1; }
    if (c + NST - 1 < nch) st.issue(stage + ((c + NST - 1) % NST) * slot, kof(c + NST - 1));
    cp_async_commit();
    if (pc) { const long long q1 = clock64(); pc->acc[6] += q1 - q0; q0 = q1; }
    const int kbase = kof(c);
    const __nv_bfloat16* sb0 = stage + (c % NST) * slot + g * sld + 2 * t;
    for (int sub = 0; sub < kc && kbase + sub < K; sub += KC) {
      const __nv_bfloat16* wk = wa + kbase + sub;
      const __nv_bfloat16* sb = sb0 + sub;
      uint32_t af[KC / 16][4];
#pragma unroll
      for (int kk = 0; kk < KC / 16; ++kk) {
        af[kk][0] = *reinterpret_cast<const uint32_t*>(wk + kk * 16);
        af[kk][1] = *reinterpret_cast<const uint32_t*>(wk + 8 * ldw + kk * 16);
        af[kk][2] = *reinterpret_cast<const uint32_t*>(wk + kk * 16 + 8);
        af[kk][3] = *reinterpret_cast<const uint32_t*>(wk + 8 * ldw + kk * 16 + 8);
      }
#pragma unroll
      for (int i = 0; i < MT; ++i) {
        const int nt = nt0 + NSTEP * i;
        if (nt < ntv) {
          uint32_t bf[KC / 16][2];
#pragma unroll
          for (int kk = 0; kk < KC / 16; ++kk) {
            bf[kk][0] = *reinterpret_cast<const uint32_t*>(sb + nt * 8 * sld + kk * 16);
            bf[kk][1] = *reinterpret_cast<const uint32_t*>(sb + nt * 8 * sld + kk * 16 + 8);
          }
#pragma unroll
          for (int kk = 0; kk < KC / 16; ++kk)
            mma_bf16(part[kk % NCH][i], af[kk][0], af[kk][1], af[kk][2], af[kk][3], bf[kk][0], bf[kk][1]);
        }
      }
    }
    if (pc) { const long long q1 = clock64(); pc->acc[7] += q1 - q0; }
  }
#pragma unroll
  for (int i = 0; i < MT; ++i)
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float v = part[0][i][e];
#pragma unroll
      for (int q = 1; q < NCH; ++q) v += part[q][i][e];
      acc[i][e] += v;
    }
  cp_async_wait<0>();
  __syncthreads();
}


// ------------------------------------------------------------------------------------ tcgen05 variant (prec = 2)
// The per-step product on 5th-generation tensor cores: D[batch (M = 64 TMEM lanes), rows (N = NR columns)] +=
// act[batch, K] . Wslice[rows, K]^T.  The weight slice is the B operand, resident in shared memory for the whole
// sequence in the canonical K-major SWIZZLE_128B layout (one [NR x 128 B] tile per 64 columns); the [B, K] activation
// chunk is the A operand, staged by cp.async straight into the same swizzled layout (64-row tiles; rows past the
// CTA's batch range stay zero).  One elected thread issues the MMAs of a chunk; tcgen05.commit recycles the ring slot
// and, after the last chunk, publishes the accumulator, which warps 0-3 read back with tcgen05.ld.
struct TcState {
  uint8_t* Wt;            // [nkb][NR][128 B]
  uint8_t* ring;          // [NST][kcb][64][128 B]
  uint64_t* slot_free;    // [NST]
  uint64_t* acc_ready;
  uint32_t tmem;
  uint32_t gc;            // chunks issued so far (all steps): slot = gc % NST, use = gc / NST
  uint32_t steps;         // accumulator hand-offs so far
  int NST, kcb, rb;
};

struct StagerTc {
  const __nv_bfloat16* r0[2];
  const __nv_bfloat16* r1[2];
  int dst[2], kk[2], n0, K;
  bool on[2];
  __device__ __forceinline__ void init(const Seg& sg, int b0, int B, int K_, int kc, int rb) {
    n0 = sg.n0;
    K = K_;
    const int cpr = kc >> 3;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int idx = threadIdx.x + LT * i;
      on[i] = idx < rb * cpr;
      const int bl = idx / cpr, b = b0 + bl, c = idx - bl * cpr;        // c: 16-byte column inside the chunk
      kk[i] = c << 3;
      dst[i] = (c >> 3) * 8192 + bl * 128 + (((c & 7) ^ (bl & 7)) << 4);   // bytes: k-block tile, row, swizzled 16-byte cell
      const bool ok = on[i] && b < B;
      r0[i] = (ok && sg.q0) ? sg.q0 + b * sg.s0 : nullptr;
      r1[i] = (ok && sg.q1) ? sg.q1 + b * sg.s1 - sg.n0 : nullptr;
    }
  }
  __device__ __forceinline__ void issue(uint8_t* slot, int k0) const {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      if (!on[i]) continue;
      const int k = k0 + kk[i];
      const __nv_bfloat16* base = (k < n0) ? r0[i] : r1[i];
      if (base != nullptr && k < K) cp_async16(slot + dst[i], base + k);
      else *reinterpret_cast<uint4*>(slot + dst[i]) = make_uint4(0u, 0u, 0u, 0u);
    }
  }
};

// out[bl * ldo + n] = sum_k act[b0 + bl, k] * Wslice[n, k] for bl < rb, n < NR
template <int NR>
__device__ __forceinline__ void slice_gemm_tc(TcState& st, const Seg& sg, int K, int b0, int B, float* out, int ldo,
                                              PhaseClock* pc = nullptr) {
  using namespace tc;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int kc = st.kcb * 64, slot_bytes = st.kcb * 8192;
  const int nch = (K + kc - 1) / kc;
  StagerTc sgr;
  sgr.init(sg, b0, B, K, kc, st.rb);
  auto acquire_and_issue = [&](int c) {             // chunk c of this step goes to ring slot (gc + c) % NST
    const uint32_t g = st.gc + (uint32_t)c, slot = g % (uint32_t)st.NST, use = g / (uint32_t)st.NST;
    if (use > 0) mbar_wait(&st.slot_free[slot], (use - 1) & 1);        // the MMAs that last read this slot are done
    sgr.issue(st.ring + slot * slot_bytes, c * kc);
  };
  for (int s = 0; s < st.NST - 1; ++s) {
    if (s < nch) acquire_and_issue(s);
    cp_async_commit();
  }
  const uint32_t idesc = umma_idesc(64, NR, 0, 0);
  const uint64_t da0 = umma_desc(smem_u32(st.ring), 16, 1024), db0 = umma_desc(smem_u32(st.Wt), 16, 1024);
  for (int c = 0; c < nch; ++c) {
    long long q0 = 0;
    if (pc) q0 = clock64();
    cp_async_wait_dyn(st.NST - 2);
    fence_proxy_async();                            // this thread's landed cp.async / zero stores -> visible to the MMA proxy
    __syncthreads();
    if (pc) { const long long q1 = clock64(); pc->acc[5] += q1 - q0; q0 = q1; }
    if (c + st.NST - 1 < nch) acquire_and_issue(c + st.NST - 1);
    cp_async_commit();
    if (pc) { const long long q1 = clock64(); pc->acc[6] += q1 - q0; q0 = q1; }
    if (tid == 0) {
      tc_fence_after();
      const uint32_t slot = (st.gc + (uint32_t)c) % (uint32_t)st.NST;
      // descriptors differ only in the 14-bit start-address field: advance it with plain adds
      uint64_t da = da0 + (uint64_t)(slot * (uint32_t)(slot_bytes >> 4));
      uint64_t db = db0 + (uint64_t)((uint32_t)(c * st.kcb) * (uint32_t)(NR * 8));
      const int nkbv = min(st.kcb, (K - c * kc + 63) >> 6);
      for (int kb = 0; kb < nkbv; ++kb) {
        tc_mma(st.tmem, da, db, idesc, (c | kb) ? 1u : 0u);
        tc_mma(st.tmem, da + 2, db + 2, idesc, 1u);
        tc_mma(st.tmem, da + 4, db + 4, idesc, 1u);
        tc_mma(st.tmem, da + 6, db + 6, idesc, 1u);
        da += 512;                  // next 64-column A tile: 8192 B
        db += NR * 8;               // next 64-column B tile: NR * 128 B
      }
      tc_commit(&st.slot_free[slot]);
      if (c == nch - 1) tc_commit(st.acc_ready);
    }
    if (pc) { const long long q1 = clock64(); pc->acc[7] += q1 - q0; }
  }
  st.gc += (uint32_t)nch;
  cp_async_wait<0>();
  long long q2 = 0;
  if (pc) q2 = clock64();
  mbar_wait(st.acc_ready, st.steps & 1);
  if (pc) pc->acc[3] += clock64() - q2;
  st.steps += 1;
  tc_fence_after();
  // accumulator row m (batch) lives in TMEM lane (m % 16) + 32 * (m / 16): warp q < 4 holds batches 16q .. 16q+15
  if (w < 4) {
    const int bl = 16 * w + lane;
#pragma unroll 1
    for (int c0 = 0; c0 < NR; c0 += 16) {
      uint32_t v[16];
      tc_ld16(st.tmem + ((uint32_t)(w * 32) << 16) + (uint32_t)c0, v);
      if (lane < 16 && bl < st.rb) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          *reinterpret_cast<uint4*>(out + bl * ldo + c0 + 4 * j) = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
}

// one-time setup: barriers, TMEM columns, zeroed ring (rows past the batch range must read as zeros), swizzled weights
template <int NR>
__device__ __forceinline__ void tc_setup(TcState& st, uint32_t* tmem_slot, int nkb) {
  using namespace tc;
  const int tid = threadIdx.x, w = tid >> 5;
  if (tid == 0) {
    for (int s = 0; s < st.NST; ++s) mbar_init(&st.slot_free[s], 1);
    mbar_init(st.acc_ready, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  constexpr uint32_t COLS = NR < 32 ? 32 : (NR <= 32 ? 32 : (NR <= 64 ? 64 : (NR <= 128 ? 128 : 256)));
  if (w == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  const int ring_bytes = st.NST * st.kcb * 8192;
  for (int i = tid * 16; i < ring_bytes; i += LT * 16) *reinterpret_cast<uint4*>(st.ring + i) = make_uint4(0u, 0u, 0u, 0u);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  st.tmem = *tmem_slot;
  st.gc = 0;
  st.steps = 0;
  (void)nkb;
}
template <int NR>
__device__ __forceinline__ void tc_teardown(TcState& st) {
  using namespace tc;
  constexpr uint32_t COLS = NR < 32 ? 32 : (NR <= 32 ? 32 : (NR <= 64 ? 64 : (NR <= 128 ? 128 : 256)));
  tc_fence_before();
  __syncthreads();
  if ((threadIdx.x >> 5) == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(st.tmem), "r"(COLS) : "memory");
  }
}
// fp32 global row `src` (K columns) -> row lr of the swizzled B-operand tiles (zero past K up to nkb * 64)
template <int NR>
__device__ __forceinline__ void fill_slice_tc(uint8_t* Wt, int lr, const float* src, int K, int nkb) {
  for (int k8 = threadIdx.x; k8 < nkb * 8; k8 += LT) {          // one 16-byte cell (8 columns) per iteration
    uint32_t pk[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int k = k8 * 8 + 2 * e;
      const float a = k < K ? src[k] : 0.f, b = k + 1 < K ? src[k + 1] : 0.f;
      pk[e] = tc::pack_bf16(a, b);
    }
    const int kb = k8 >> 3, c = k8 & 7;
    *reinterpret_cast<uint4*>(Wt + (size_t)kb * (NR * 128) + lr * 128 + ((c ^ (lr & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
  }
}

// fp32 rows of a weight matrix -> bf16 resident slice (tail columns up to ldw zeroed)
__device__ __forceinline__ void fill_slice16(__nv_bfloat16* Ws, int ldw, int lr, const float* src, int K) {
  // called by all threads with the same (lr, src): columns strided over the block
  for (int k = threadIdx.x; k < ldw; k += LT) Ws[lr * ldw + k] = __float2bfloat16(k < K ? src[k] : 0.f);
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

// bf16 twin of warp_rows_dot: weights rows (smem) and the vector (global, L2) are bf16, 8 elements per lane per step
__device__ __forceinline__ void warp_rows_dot16(float (&out)[4], const __nv_bfloat16* Wrows, int ldw, int nr,
                                                const __nv_bfloat16* vec, int K8) {
  const int lane = threadIdx.x & 31;
  float a[4] = {0.f, 0.f, 0.f, 0.f};
  for (int k8 = lane; k8 < K8; k8 += 32) {
    const uint4 v = __ldcg(reinterpret_cast<const uint4*>(vec) + k8);
    const __nv_bfloat162* vp = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      if (r < nr) {
        const uint4 q = *reinterpret_cast<const uint4*>(Wrows + r * ldw + 8 * k8);
        const __nv_bfloat162* qp = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 x = __bfloat1622float2(vp[e]), y = __bfloat1622float2(qp[e]);
          a[r] = fmaf(x.x, y.x, a[r]);
          a[r] = fmaf(x.y, y.y, a[r]);
        }
      }
    }
  }
  out[0] = warp_sum(a[0]); out[1] = warp_sum(a[1]); out[2] = warp_sum(a[2]); out[3] = warp_sum(a[3]);
}

// =====================================================================================  forward
template <int HS, bool FB, bool RES, int BF>
__global__ void __launch_bounds__(LT, 1) lstm_fwd_kernel(const ag_lstm_desc d, const int ncta_dir, const int PR, const int NST, const int bsplit,
                                                         const int kcb, const int nbg2) {
  constexpr int ROWS = 4 * HS;
  constexpr int GSLD = BF == 2 ? ROWS + 4 : ROWS;   // gate exchange tile row stride
  constexpr int MT = ROWS / 16 > 0 ? ROWS / 16 : 1;
  constexpr int RL = ROWS < 32 ? ROWS : 32;
  constexpr int KS = 32 / RL;
  constexpr int RPT = ROWS / RL;
  constexpr int IPT = (BTILE * HS) / LT;            // cell-update items per thread per batch tile
  extern __shared__ __align__(16) float smem[];
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int H = d.H, F = FB ? d.F : 0, K1 = H + F, ndir = d.ndir, B = d.B, T = d.T, Tcap = d.Tcap;
  // CTA = (direction, slice of HS hidden units, batch range): blockIdx.x = (dir * ncta_dir + slice) * bsplit + bs
  const int grp = blockIdx.x / bsplit, bsi = blockIdx.x - grp * bsplit;
  const int dir = grp / ncta_dir, cta = grp % ncta_dir, j0 = cta * HS;
  const int Bper = (((B + bsplit - 1) / bsplit) + 7) & ~7;
  const int blo = min(B, bsi * Bper), bhi = min(B, blo + Bper);
  const int ldw = pad_ld(K1), ld2 = pad_ld(H);

  const int ldw16 = pad_ld16(K1);
  // fp32: [Ws fp32][stage 2 x 64 x 68 fp32];  bf16: [Ws16 bf16][stage16 4 x 64 x 72 bf16]  (both multiples of 16 B)
  float* Ws = smem;
  __nv_bfloat16* Ws16 = reinterpret_cast<__nv_bfloat16*>(smem);
  float* stage = Ws + (RES ? ROWS * ldw : 0);
  __nv_bfloat16* stage16 = Ws16 + ROWS * ldw16;
  // bf16: the gate exchange tile aliases the (idle) cp.async ring
  float* gs = BF ? reinterpret_cast<float*>(stage16) : stage + 2 * BTILE * SLD;
  float* cs = BF ? reinterpret_cast<float*>(stage16 + NST * SLOT16) : gs + BTILE * ROWS;
  TcState tcs;
  uint32_t* tmem_slot = nullptr;
  const int nkb = (K1 + 63) / 64;
  if (BF == 2) {
    // tcgen05: [Wt nkb x ROWS x 128 B | ring NST x kcb x 8 KB | barriers 64 B | gs 64 x GSLD fp32 | cs ...], 1 KB aligned
    uint8_t* sm8 = reinterpret_cast<uint8_t*>(smem);
    sm8 += (1024u - (tc::smem_u32(sm8) & 1023u)) & 1023u;
    tcs.Wt = sm8;
    tcs.ring = sm8 + (size_t)nkb * ROWS * 128;
    uint64_t* bars = reinterpret_cast<uint64_t*>(tcs.ring + (size_t)NST * kcb * 8192);
    tcs.slot_free = bars;
    tcs.acc_ready = bars + 6;
    tmem_slot = reinterpret_cast<uint32_t*>(bars + 7);
    tcs.NST = NST; tcs.kcb = kcb;
    tcs.rb = min(64, (Bper + 15) & ~15);
    gs = reinterpret_cast<float*>(bars + 8);
    cs = gs + BTILE * GSLD;
  }
  float* W2s = cs + Bper * HS;
  // phase-2 weights: fp32 rows (stride ld2) or, in the bf16 modes, bf16 rows (stride H + 8)
  int* gen = reinterpret_cast<int*>(W2s + (FB ? (BF ? (PR * (H + 8) + 1) / 2 : PR * ld2) : 0));
  int* cnt = gen + B;

  // local row lr = jj*4 + q  <->  global gate row q*H + j0 + jj
  const float* w1d = d.w1 + (int64_t)dir * 4 * H * K1;
  if (BF == 2) {
    for (int lr = 0; lr < ROWS; ++lr) fill_slice_tc<ROWS>(tcs.Wt, lr, w1d + (int64_t)((lr & 3) * H + j0 + (lr >> 2)) * K1, K1, nkb);
    tc_setup<ROWS>(tcs, tmem_slot, nkb);
  } else if (BF) {
    for (int lr = 0; lr < ROWS; ++lr) fill_slice16(Ws16, ldw16, lr, w1d + (int64_t)((lr & 3) * H + j0 + (lr >> 2)) * K1, K1);
  } else if (RES) {
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
  for (int i = tid; i < Bper * HS; i += LT) cs[i] = 0.f;
  // phase-2 rows owned by this CTA (feedback only; ndir == 1)
  // phase 2 is tiled 2-D: CTA = (row group of PR proj/stop rows, batch group) -> each CTA reads only its batch
  // group's h rows instead of all of h
  int p0 = 0, np = 0, b2lo = 0, b2hi = 0;
  bool owns_logit = false;
  if (FB) {
    const int rg = blockIdx.x / nbg2, bg = blockIdx.x - rg * nbg2;
    const int B2per = (B + nbg2 - 1) / nbg2;
    b2lo = min(B, bg * B2per);
    b2hi = min(B, b2lo + B2per);
    p0 = rg * PR;
    np = min(PR, F + 1 - p0);
    if (np < 0 || b2hi <= b2lo) np = 0;
    owns_logit = (np > 0) && (p0 + np == F + 1);
    if (BF) {
      __nv_bfloat16* W2s16 = reinterpret_cast<__nv_bfloat16*>(W2s);
      for (int idx = tid; idx < np * H; idx += LT) {
        const int r = idx / H, k = idx - r * H;
        W2s16[r * (H + 8) + k] = __float2bfloat16(d.w2[(int64_t)(p0 + r) * H + k]);
      }
    } else {
      for (int idx = tid; idx < np * (H / 4); idx += LT) {
        const int r = idx / (H / 4), k4 = idx - r * (H / 4);
        *reinterpret_cast<float4*>(W2s + r * ld2 + 4 * k4) = *reinterpret_cast<const float4*>(d.w2 + (int64_t)(p0 + r) * H + 4 * k4);
      }
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
  PhaseClock pc;
  for (int i = 0; i < 8; ++i) pc.acc[i] = 0;
  const long long tstart = clock64();

  for (int s = 0; s < T; ++s) {
    pc.start();
    const int t = dir ? (T - 1 - s) : s;
    const int prow = dir ? (t + 2) : t;                  // hbuf row holding the previous h
    Seg sg;
    sg.p0 = d.hbuf + (int64_t)prow * ndir * H + dir * H; sg.s0 = hstr; sg.n0 = H;
    sg.p1 = FB ? (d.xbuf + (int64_t)t * F) : nullptr; sg.s1 = (int64_t)(Tcap + 1) * F; sg.n1 = F;
    sg.q0 = BF ? reinterpret_cast<const __nv_bfloat16*>(d.hbuf16) + (int64_t)prow * ndir * H + dir * H : nullptr;
    sg.q1 = (BF && FB) ? reinterpret_cast<const __nv_bfloat16*>(d.xbuf16) + (int64_t)t * F : nullptr;

    for (int b0 = blo; b0 < bhi; b0 += BTILE) {
      // prefetch this tile's input projections (independent of the recurrence)
      float pre[IPT > 0 ? IPT : 1][4];
#pragma unroll
      for (int ii = 0; ii < IPT; ++ii) {
        const int it = tid + LT * ii, bl = it / HS, jj = it - bl * HS, b = b0 + bl;
#pragma unroll
        for (int q = 0; q < 4; ++q)
          pre[ii][q] = (b < bhi) ? d.pre[b * gstr + (int64_t)t * ndir * 4 * H + dir * 4 * H + q * H + j0 + jj] : 0.f;
      }
      if (BF == 2) {
        slice_gemm_tc<ROWS>(tcs, sg, K1, b0, bhi, gs, GSLD, d.dbg ? &pc : nullptr);
        pc.lap(0);
      } else if (BF) {
        float accm[MT][4];
#pragma unroll
        for (int i = 0; i < MT; ++i) { accm[i][0] = 0.f; accm[i][1] = 0.f; accm[i][2] = 0.f; accm[i][3] = 0.f; }
        slice_gemm_mma<MT>(accm, Ws16, ldw16, sg, K1, b0, bhi, stage16, NST, ring_kc(Bper), d.dbg ? &pc : nullptr);
        pc.lap(0);
        const int g = lane >> 2, tq = lane & 3, mt = w % MT, nt0 = w / MT;
#pragma unroll
        for (int i = 0; i < MT; ++i) {
          const int bl = (nt0 + (8 / MT) * i) * 8 + 2 * tq, lr = mt * 16 + g;
          gs[bl * ROWS + lr] = accm[i][0];
          gs[(bl + 1) * ROWS + lr] = accm[i][1];
          gs[bl * ROWS + lr + 8] = accm[i][2];
          gs[(bl + 1) * ROWS + lr + 8] = accm[i][3];
        }
      } else {
        float acc[RPT][8];
#pragma unroll
        for (int rr = 0; rr < RPT; ++rr)
#pragma unroll
          for (int nb = 0; nb < 8; ++nb) acc[rr][nb] = 0.f;
        slice_gemm<RPT, KS>(acc, wrow, sg, K1, b0, bhi, stage);
        if (lane < RL) {
#pragma unroll
          for (int rr = 0; rr < RPT; ++rr)
#pragma unroll
            for (int nb = 0; nb < 8; ++nb) gs[(w * 8 + nb) * ROWS + lane + RL * rr] = acc[rr][nb];
        }
      }
      __syncthreads();
#pragma unroll
      for (int ii = 0; ii < IPT; ++ii) {
        const int it = tid + LT * ii, bl = it / HS, jj = it - bl * HS, b = b0 + bl;
        if (b >= bhi) continue;
        const int j = j0 + jj;
        const float4 a = *reinterpret_cast<const float4*>(gs + bl * GSLD + jj * 4);
        const bool valid = !d.len || t < d.len[b];
        float gi = 0.f, gf = 0.f, gg = 0.f, go = 0.f, c = 0.f, h = 0.f;
        if (valid) {
          gi = sigmoidf_(a.x + pre[ii][0]);
          gf = sigmoidf_(a.y + pre[ii][1]);
          gg = tanhf(a.z + pre[ii][2]);
          go = sigmoidf_(a.w + pre[ii][3]);
          c = gf * cs[(b - blo) * HS + jj] + gi * gg;
          h = go * tanhf(c);
          cs[(b - blo) * HS + jj] = c;
        }
        d.hbuf[b * hstr + (int64_t)(t + 1) * ndir * H + dir * H + j] = h;
        if (d.hbuf16) reinterpret_cast<__nv_bfloat16*>(d.hbuf16)[b * hstr + (int64_t)(t + 1) * ndir * H + dir * H + j] = __float2bfloat16(h);
        if (d.cbuf) d.cbuf[b * cstr + (int64_t)t * ndir * H + dir * H + j] = c;
        if (d.gates) {
          float* gp = d.gates + b * gstr + (int64_t)t * ndir * 4 * H + dir * 4 * H + j;
          gp[0] = gi; gp[H] = gf; gp[2 * H] = gg; gp[3 * H] = go;
        }
      }
      __syncthreads();
    }
    pc.lap(1);
    grid_barrier(bar, (++nbar) * ncta_dir * bsplit);
    pc.lap(2);

    if (FB) {
      // phase 2: x_t = tanh(wp h_t + bp), logit = ws h_t + bs, stop draw, early-exit flag
      if (s > 0 && blockIdx.x == 0 && tid == 0) reinterpret_cast<volatile unsigned*>(d.barrier)[4 + ((s + 1) & 1)] = 0u;
      if (np > 0) {
        constexpr int NB = 8;                                      // batches in flight per warp
        const int nb2 = b2hi - b2lo, rotb = (int)(blockIdx.x / nbg2) % nb2;
        for (int bi0 = w; bi0 < nb2; bi0 += NB * (LT / 32)) {
          const float* hrow[NB];
          const __nv_bfloat16* hrow16[NB];
          int bb[NB];
#pragma unroll
          for (int i = 0; i < NB; ++i) {
            const int bi = bi0 + i * (LT / 32);
            bb[i] = bi < nb2 ? b2lo + (bi + rotb) % nb2 : -1;     // per-CTA rotation of the batch order
            hrow[i] = bb[i] >= 0 ? d.hbuf + bb[i] * hstr + (int64_t)(t + 1) * H : nullptr;
            hrow16[i] = (BF && bb[i] >= 0) ? reinterpret_cast<const __nv_bfloat16*>(d.hbuf16) + bb[i] * hstr + (int64_t)(t + 1) * H : nullptr;
          }
          for (int pp = 0; pp < np; pp += 4) {
            float o[NB][4];
            const int nr = min(4, np - pp);
            if (BF) warp_rows_dot16_nb<NB>(o, reinterpret_cast<const __nv_bfloat16*>(W2s) + pp * (H + 8), H + 8, nr, hrow16, H / 8);
            else warp_rows_dot_nb<NB>(o, W2s + pp * ld2, ld2, nr, hrow, H / 4);
            if (lane == 0) {
#pragma unroll
              for (int i = 0; i < NB; ++i) {
                const int b = bb[i];
                if (b < 0) continue;
                for (int r = 0; r < nr; ++r) {
                  const int p = p0 + pp + r;
                  const float v = o[i][r] + d.b2[p];
                  if (p < F) {
                    const float xv = tanhf(v);
                    d.xbuf[(b * (int64_t)(Tcap + 1) + t + 1) * F + p] = xv;
                    if (d.xbuf16) reinterpret_cast<__nv_bfloat16*>(d.xbuf16)[(b * (int64_t)(Tcap + 1) + t + 1) * F + p] = __float2bfloat16(xv);
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
      }
      if (owns_logit) {                 // one owner per batch group: count its samples that are still generating
        __syncthreads();
        int g = 0;
        for (int b = b2lo + tid; b < b2hi; b += LT) g += gen[b] ? 1 : 0;
        g = (int)block_sum((float)g, reinterpret_cast<float*>(cnt + B));
        if (tid == 0 && g > 0) atomicAdd(d.barrier + 4 + (s & 1), (unsigned)g);
      }
      pc.lap(3);
      grid_barrier(bar, (++nbar) * ncta_dir * bsplit);
      pc.lap(2);
      if (d.u && *reinterpret_cast<volatile unsigned*>(d.barrier + 4 + (s & 1)) == 0u) { steps_run = t + 1; break; }
    }
  }
  if (d.dbg && tid == 0) {
    long long* q = d.dbg + (int64_t)blockIdx.x * 8;
    for (int i = 0; i < 8; ++i) q[i] = pc.acc[i];
    q[4] = clock64() - tstart;
  }
  if (BF == 2) tc_teardown<ROWS>(tcs);
  if (FB) {
    __syncthreads();
    if (owns_logit)
      for (int b = b2lo + tid; b < b2hi; b += LT) if (d.glen) d.glen[b] = cnt[b];
    if (blockIdx.x == 0 && tid == 0) *reinterpret_cast<volatile int*>(d.t_end) = steps_run;
  }
}

// ====================================================================================  backward
template <int HS, bool FB, bool RES, int BF>
__global__ void __launch_bounds__(LT, 1) lstm_bwd_kernel(const ag_lstm_desc d, const int ncta_dir, const int PR, const int NST, const int bsplit,
                                                         const int kcb, const int nbg2) {
  constexpr int DHLD = BF == 2 ? HS + 4 : HS;       // dh exchange tile row stride
  constexpr int RL = HS;            // 4, 8 or 16 rows -> k split over the rest of the warp
  constexpr int KS = 32 / RL;
  constexpr int IPT = (BTILE * HS) / LT;
  extern __shared__ __align__(16) float smem[];
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int H = d.H, F = FB ? d.F : 0, FP = FB ? ((F + 1 + 7) / 8) * 8 : 0, K = 4 * H + FP;
  const int ndir = d.ndir, B = d.B, T = d.T, Tcap = d.Tcap;
  const int grp = blockIdx.x / bsplit, bsi = blockIdx.x - grp * bsplit;
  const int dir = grp / ncta_dir, cta = grp % ncta_dir, j0 = cta * HS;
  const int Bper = (((B + bsplit - 1) / bsplit) + 7) & ~7;
  const int blo = min(B, bsi * Bper), bhi = min(B, blo + Bper);
  constexpr int MT = HS / 16 > 0 ? HS / 16 : 1;
  const int ldw = pad_ld(K), ldx = pad_ld(4 * H);

  const int ldw16 = pad_ld16(K);
  float* Ws = smem;
  __nv_bfloat16* Ws16 = reinterpret_cast<__nv_bfloat16*>(smem);
  float* stage = Ws + (RES ? HS * ldw : 0);
  __nv_bfloat16* stage16 = Ws16 + HS * ldw16;
  float* dhs = BF ? reinterpret_cast<float*>(stage16) : stage + 2 * BTILE * SLD;      // bf16: aliases the idle ring
  float* dcs = BF ? reinterpret_cast<float*>(stage16 + NST * SLOT16) : dhs + BTILE * HS;
  TcState tcs;
  uint32_t* tmem_slot = nullptr;
  const int nkb = (K + 63) / 64;
  if (BF == 2) {
    uint8_t* sm8 = reinterpret_cast<uint8_t*>(smem);
    sm8 += (1024u - (tc::smem_u32(sm8) & 1023u)) & 1023u;
    tcs.Wt = sm8;
    tcs.ring = sm8 + (size_t)nkb * HS * 128;
    uint64_t* bars = reinterpret_cast<uint64_t*>(tcs.ring + (size_t)NST * kcb * 8192);
    tcs.slot_free = bars;
    tcs.acc_ready = bars + 6;
    tmem_slot = reinterpret_cast<uint32_t*>(bars + 7);
    tcs.NST = NST; tcs.kcb = kcb;
    tcs.rb = min(64, (Bper + 15) & ~15);
    dhs = reinterpret_cast<float*>(bars + 8);
    dcs = dhs + BTILE * DHLD;
  }
  float* Wxs = dcs + Bper * HS;

  const float* w1d = d.w1t + ((int64_t)dir * H + j0) * K;
  if (BF == 2) {
    for (int lr = 0; lr < HS; ++lr) fill_slice_tc<HS>(tcs.Wt, lr, w1d + (int64_t)lr * K, K, nkb);
    tc_setup<HS>(tcs, tmem_slot, nkb);
  } else if (BF) {
    for (int lr = 0; lr < HS; ++lr) fill_slice16(Ws16, ldw16, lr, w1d + (int64_t)lr * K, K);
  } else if (RES) {
    const int K4 = K / 4;
    for (int idx = tid; idx < HS * K4; idx += LT) {
      const int lr = idx / K4, k4 = idx - lr * K4;
      *reinterpret_cast<float4*>(Ws + lr * ldw + 4 * k4) = *reinterpret_cast<const float4*>(w1d + (int64_t)lr * K + 4 * k4);
    }
  }
  const float* wrow[1];
  wrow[0] = RES ? (Ws + (lane % RL) * ldw) : (w1d + (int64_t)(lane % RL) * K);
  for (int i = tid; i < Bper * HS; i += LT) dcs[i] = 0.f;
  int p0 = 0, np = 0, b2lo = 0, b2hi = 0;
  if (FB) {                                              // phase A tiled 2-D: (row group of wx^T rows, batch group)
    const int rg = blockIdx.x / nbg2, bg = blockIdx.x - rg * nbg2;
    const int B2per = (B + nbg2 - 1) / nbg2;
    b2lo = min(B, bg * B2per);
    b2hi = min(B, b2lo + B2per);
    p0 = rg * PR;
    np = min(PR, F - p0);
    if (np < 0 || b2hi <= b2lo) np = 0;
    if (BF) {                                            // bf16 rows of wx^T, stride 4H + 8
      __nv_bfloat16* Wxs16 = reinterpret_cast<__nv_bfloat16*>(Wxs);
      for (int idx = tid; idx < np * 4 * H; idx += LT) {
        const int r = idx / (4 * H), k = idx - r * 4 * H;
        Wxs16[r * (4 * H + 8) + k] = __float2bfloat16(d.wxt[(int64_t)(p0 + r) * 4 * H + k]);
      }
    } else {
      for (int idx = tid; idx < np * H; idx += LT) {     // H float4 per row of wxt (4H floats)
        const int r = idx / H, k4 = idx - r * H;
        *reinterpret_cast<float4*>(Wxs + r * ldx + 4 * k4) = *reinterpret_cast<const float4*>(d.wxt + (int64_t)(p0 + r) * 4 * H + 4 * k4);
      }
    }
  }
  __syncthreads();

  unsigned* bar = d.barrier + dir;
  unsigned nbar = 0;
  const int64_t gstr = (int64_t)Tcap * ndir * 4 * H;
  const int64_t cstr = (int64_t)Tcap * ndir * H;

  PhaseClock pc;
  for (int i = 0; i < 8; ++i) pc.acc[i] = 0;
  const long long tstart = clock64();
  for (int s = 0; s < T; ++s) {
    pc.start();
    const int t = dir ? s : (T - 1 - s);                 // reverse of the forward order
    const int tn = dir ? (t - 1) : (t + 1);              // the step processed just before this one
    const bool has_next = s > 0;

    if (FB) {
      // phase A: dpx[b,t,p] = (dx_ext + wx^T dgates_{t+1})[p] * (1 - x_t[p]^2); column F = ds_ext
      if (np > 0) {
        constexpr int NB = 8;
        const int nb2 = b2hi - b2lo, rotb = (int)(blockIdx.x / nbg2) % nb2;
        for (int bi0 = w; bi0 < nb2; bi0 += NB * (LT / 32)) {
          const float* dgrow[NB];
          const __nv_bfloat16* dgrow16[NB];
          int bb[NB];
#pragma unroll
          for (int i = 0; i < NB; ++i) {
            const int bi = bi0 + i * (LT / 32);
            bb[i] = bi < nb2 ? b2lo + (bi + rotb) % nb2 : -1;
            dgrow[i] = (bb[i] >= 0 && has_next) ? d.dgates + bb[i] * gstr + (int64_t)tn * 4 * H : nullptr;
            dgrow16[i] = (BF && bb[i] >= 0 && has_next)
                             ? reinterpret_cast<const __nv_bfloat16*>(d.dgates16) + bb[i] * gstr + (int64_t)tn * 4 * H : nullptr;
          }
          for (int pp = 0; pp < np; pp += 4) {
            float o[NB][4];
            const int nr = min(4, np - pp);
            if (BF) warp_rows_dot16_nb<NB>(o, reinterpret_cast<const __nv_bfloat16*>(Wxs) + pp * (4 * H + 8), 4 * H + 8, nr, dgrow16, H / 2);
            else warp_rows_dot_nb<NB>(o, Wxs + pp * ldx, ldx, nr, dgrow, H);
            if (lane == 0) {
#pragma unroll
              for (int i = 0; i < NB; ++i) {
                const int b = bb[i];
                if (b < 0) continue;
                for (int r = 0; r < nr; ++r) {
                  const int p = p0 + pp + r;
                  float dx = o[i][r];
                  if (d.dx_ext) dx += d.dx_ext[(b * (int64_t)Tcap + t) * F + p];
                  const float x = d.xbuf[(b * (int64_t)(Tcap + 1) + t + 1) * F + p];
                  const float dpv = dx * (1.f - x * x);
                  d.dpx[(b * (int64_t)Tcap + t) * FP + p] = dpv;
                  if (d.dpx16) reinterpret_cast<__nv_bfloat16*>(d.dpx16)[(b * (int64_t)Tcap + t) * FP + p] = __float2bfloat16(dpv);
                }
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
          if (d.dpx16) {
            __nv_bfloat16* q16 = reinterpret_cast<__nv_bfloat16*>(d.dpx16) + (b * (int64_t)Tcap + t) * FP;
            for (int p = F; p < FP; ++p) q16[p] = __float2bfloat16(q[p]);
          }
        }
      }
      pc.lap(3);
      grid_barrier(bar, (++nbar) * ncta_dir * bsplit);
      pc.lap(2);
    }

    // phase B: dh = dh_ext + whh^T dgates_next (+ wp^T dpx_t + ws ds_t), then the cell backward
    Seg sg;
    sg.p0 = has_next ? (d.dgates + (int64_t)tn * ndir * 4 * H + dir * 4 * H) : nullptr; sg.s0 = gstr; sg.n0 = 4 * H;
    sg.p1 = FB ? (d.dpx + (int64_t)t * FP) : nullptr; sg.s1 = (int64_t)Tcap * FP; sg.n1 = FP;
    sg.q0 = (BF && has_next) ? reinterpret_cast<const __nv_bfloat16*>(d.dgates16) + (int64_t)tn * ndir * 4 * H + dir * 4 * H : nullptr;
    sg.q1 = (BF && FB) ? reinterpret_cast<const __nv_bfloat16*>(d.dpx16) + (int64_t)t * FP : nullptr;
    for (int b0 = blo; b0 < bhi; b0 += BTILE) {
      // prefetch what the cell backward needs (saved gates, c_t, c_prev, external dh): independent of this step's GEMM
      float pg[IPT > 0 ? IPT : 1][7];
      bool pv[IPT > 0 ? IPT : 1];
#pragma unroll
      for (int ii = 0; ii < IPT; ++ii) {
        const int it = tid + LT * ii, bl = it / HS, jj = it - bl * HS, b = b0 + bl, j = j0 + jj;
        pv[ii] = false;
#pragma unroll
        for (int e = 0; e < 7; ++e) pg[ii][e] = 0.f;
        if (b < bhi) {
          const int Lb = d.len ? min(d.len[b], T) : T;
          if (t < Lb) {
            pv[ii] = true;
            const float* gp = d.gates + b * gstr + (int64_t)t * ndir * 4 * H + dir * 4 * H + j;
            pg[ii][0] = gp[0]; pg[ii][1] = gp[H]; pg[ii][2] = gp[2 * H]; pg[ii][3] = gp[3 * H];
            pg[ii][4] = d.cbuf[b * cstr + (int64_t)t * ndir * H + dir * H + j];
            const int tp = dir ? (t + 1) : (t - 1);
            const bool has_prev = dir ? (tp < Lb) : (tp >= 0);
            pg[ii][5] = has_prev ? d.cbuf[b * cstr + (int64_t)tp * ndir * H + dir * H + j] : 0.f;
            pg[ii][6] = d.dh_ext ? d.dh_ext[b * (d.dh_ext_bs ? d.dh_ext_bs : cstr) + (int64_t)t * ndir * H + dir * H + j] : 0.f;
          }
        }
      }
      if (BF == 2) {
        if (has_next || FB) {
          slice_gemm_tc<HS>(tcs, sg, K, b0, bhi, dhs, DHLD, d.dbg ? &pc : nullptr);
        } else {
          for (int i = tid; i < BTILE * DHLD; i += LT) dhs[i] = 0.f;
        }
        pc.lap(0);
      } else if (BF) {
        float accm[MT][4];
#pragma unroll
        for (int i = 0; i < MT; ++i) { accm[i][0] = 0.f; accm[i][1] = 0.f; accm[i][2] = 0.f; accm[i][3] = 0.f; }
        if (has_next || FB) slice_gemm_mma<MT>(accm, Ws16, ldw16, sg, K, b0, bhi, stage16, NST, ring_kc(Bper), d.dbg ? &pc : nullptr);
        pc.lap(0);
        if (HS >= 16) {
          const int g = lane >> 2, tq = lane & 3, mt = w % MT, nt0 = w / MT;
#pragma unroll
          for (int i = 0; i < MT; ++i) {
            const int bl = (nt0 + (8 / MT) * i) * 8 + 2 * tq, lr = mt * 16 + g;
            dhs[bl * HS + lr] = accm[i][0];
            dhs[(bl + 1) * HS + lr] = accm[i][1];
            dhs[bl * HS + lr + 8] = accm[i][2];
            dhs[(bl + 1) * HS + lr + 8] = accm[i][3];
          }
        }
      } else {
        float acc[1][8];
#pragma unroll
        for (int nb = 0; nb < 8; ++nb) acc[0][nb] = 0.f;
        if (has_next || FB) slice_gemm<1, KS>(acc, wrow, sg, K, b0, bhi, stage);
        if (lane < RL) {
#pragma unroll
          for (int nb = 0; nb < 8; ++nb) dhs[(w * 8 + nb) * HS + lane] = acc[0][nb];
        }
      }
      __syncthreads();
#pragma unroll
      for (int ii = 0; ii < IPT; ++ii) {
        const int it = tid + LT * ii, bl = it / HS, jj = it - bl * HS, b = b0 + bl;
        if (b >= bhi) continue;
        const int j = j0 + jj;
        float* dg = d.dgates + b * gstr + (int64_t)t * ndir * 4 * H + dir * 4 * H + j;
        __nv_bfloat16* dg16 = d.dgates16 ? reinterpret_cast<__nv_bfloat16*>(d.dgates16) + b * gstr + (int64_t)t * ndir * 4 * H + dir * 4 * H + j : nullptr;
        if (!pv[ii]) {
          dg[0] = 0.f; dg[H] = 0.f; dg[2 * H] = 0.f; dg[3 * H] = 0.f;
          if (dg16) { const __nv_bfloat16 z = __float2bfloat16(0.f); dg16[0] = z; dg16[H] = z; dg16[2 * H] = z; dg16[3 * H] = z; }
          continue;
        }
        const float gi = pg[ii][0], gf = pg[ii][1], gg = pg[ii][2], go = pg[ii][3], c = pg[ii][4], cprev = pg[ii][5];
        const float dh = dhs[bl * DHLD + jj] + pg[ii][6];
        const float tc = tanhf(c);
        const float dc = dcs[(b - blo) * HS + jj] + dh * go * (1.f - tc * tc);
        dcs[(b - blo) * HS + jj] = dc * gf;
        const float di = dc * gg * gi * (1.f - gi), df = dc * cprev * gf * (1.f - gf);
        const float dgg = dc * gi * (1.f - gg * gg), dgo = dh * tc * go * (1.f - go);
        dg[0] = di; dg[H] = df; dg[2 * H] = dgg; dg[3 * H] = dgo;
        if (dg16) {
          dg16[0] = __float2bfloat16(di); dg16[H] = __float2bfloat16(df);
          dg16[2 * H] = __float2bfloat16(dgg); dg16[3 * H] = __float2bfloat16(dgo);
        }
      }
      __syncthreads();
    }
    pc.lap(1);
    grid_barrier(bar, (++nbar) * ncta_dir * bsplit);
    pc.lap(2);
  }
  if (d.dbg && tid == 0) {
    long long* q = d.dbg + (int64_t)blockIdx.x * 8;
    for (int i = 0; i < 8; ++i) q[i] = pc.acc[i];
    q[4] = clock64() - tstart;
  }
  if (BF == 2) tc_teardown<HS>(tcs);
}

// ---------------------------------------------------------------------------------- host side
struct Plan { int HS, ncta_dir, PR, nst, bsplit, kcb, bf, nbg2; bool res; size_t smem; };

static int pick_hs(int H, int ndir) {
  const int hs_opts[3] = {4, 8, 16};
  for (int i = 0; i < 3; ++i) {
    const int hs = hs_opts[i];
    if (H % hs == 0 && (int64_t)ndir * (H / hs) <= sm_count()) return hs;
  }
  return 0;
}

static size_t fwd_smem(const ag_lstm_desc* d, int HS, int PR, bool res, bool bf = false, int NST = NST_MIN) {
  const int F = d->F, K1 = d->H + F;
  size_t fl = (bf ? 0 : (size_t)BTILE * 4 * HS) + (size_t)(d->B + 8) * HS +
              (F > 0 ? (bf ? ((size_t)PR * (d->H + 8) + 1) / 2 : (size_t)PR * pad_ld(d->H)) : 0);
  size_t head = bf ? ((size_t)4 * HS * pad_ld16(K1) + (size_t)NST * SLOT16) * 2
                   : ((res ? (size_t)4 * HS * pad_ld(K1) : 0) + 2 * BTILE * SLD) * 4;
  return head + fl * 4 + (size_t)2 * d->B * 4 + 64 * 4 + 16;
}
static size_t bwd_smem(const ag_lstm_desc* d, int HS, int PR, bool res, bool bf = false, int NST = NST_MIN) {
  const int F = d->F, FP = F > 0 ? ((F + 1 + 7) / 8) * 8 : 0, K = 4 * d->H + FP;
  size_t fl = (bf ? 0 : (size_t)BTILE * HS) + (size_t)(d->B + 8) * HS + (F > 0 ? (bf ? (size_t)PR * (2 * d->H + 4) : (size_t)PR * pad_ld(4 * d->H)) : 0);
  size_t head = bf ? ((size_t)HS * pad_ld16(K) + (size_t)NST * SLOT16) * 2
                   : ((res ? (size_t)HS * pad_ld(K) : 0) + 2 * BTILE * SLD) * 4;
  return head + fl * 4 + 16;
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

// batch groups of the 2-D phase-2 / phase-A tiling.  Measured (profiles/r1_lstm_phase_cycles.txt): these phases are
// bound by the per-warp dot-product / reduction latency, not by the bytes of h they read -- more rows per CTA (what a
// batch split costs) made phase 2 slower (13.4 k -> 22.6 k cycles per step), so the 1-D row split stays the default.
static int pick_nbg2(int B, int ncta) {
  (void)B; (void)ncta;
  return 1;
}
static int bper_of(int B, int bsplit) { return (((B + bsplit - 1) / bsplit) + 7) & ~7; }
static size_t fwd_smem_tc(const ag_lstm_desc* d, int HS, int PR, int NST, int kcb, int bsplit) {
  const int F = d->F, K1 = d->H + F, nkb = (K1 + 63) / 64, rows = 4 * HS;
  return 1024 + (size_t)nkb * rows * 128 + (size_t)NST * kcb * 8192 + 64 + (size_t)BTILE * (rows + 4) * 4 +
         (size_t)bper_of(d->B, bsplit) * HS * 4 + (F > 0 ? (size_t)PR * (d->H + 8) * 2 + 4 : 0) + (size_t)2 * d->B * 4 + 64 * 4 + 16;
}
static size_t bwd_smem_tc(const ag_lstm_desc* d, int HS, int PR, int NST, int kcb, int bsplit) {
  const int F = d->F, FP = F > 0 ? ((F + 1 + 7) / 8) * 8 : 0, K = 4 * d->H + FP, nkb = (K + 63) / 64;
  return 1024 + (size_t)nkb * HS * 128 + (size_t)NST * kcb * 8192 + 64 + (size_t)BTILE * (HS + 4) * 4 +
         (size_t)bper_of(d->B, bsplit) * HS * 4 + (F > 0 ? (size_t)PR * (2 * d->H + 4) * 4 : 0) + 16;
}

template <typename KernT>
static int launch_coop(KernT kern, const ag_lstm_desc* d, const Plan& p, cudaStream_t s) {
  AG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem));
  ag_lstm_desc dd = *d;
  int ncta_dir = p.ncta_dir, PR = p.PR, nst = p.nst, bsplit = p.bsplit, kcb = p.kcb, nbg2 = p.nbg2;
  void* args[7] = {&dd, &ncta_dir, &PR, &nst, &bsplit, &kcb, &nbg2};
  AG_CUDA(cudaLaunchCooperativeKernel((const void*)kern, dim3((unsigned)(p.ncta_dir * d->ndir * p.bsplit)), dim3(LT), args, p.smem, s));
  return AG_OK;
}

#define AG_LSTM_DISPATCH_HS(KERN, HS_)                                                                               \
  do {                                                                                                              \
    if (p.bf == 1) return fb ? launch_coop(KERN<HS_, true, true, 1>, d, p, s) : launch_coop(KERN<HS_, false, true, 1>, d, p, s); \
    if (fb) return p.res ? launch_coop(KERN<HS_, true, true, 0>, d, p, s) : launch_coop(KERN<HS_, true, false, 0>, d, p, s); \
    return p.res ? launch_coop(KERN<HS_, false, true, 0>, d, p, s) : launch_coop(KERN<HS_, false, false, 0>, d, p, s);     \
  } while (0)
#define AG_LSTM_DISPATCH(KERN)          \
  do {                                  \
    const bool fb = d->F > 0;           \
    if (p.bf == 2 && p.HS == 16) return fb ? launch_coop(KERN<16, true, true, 2>, d, p, s) : launch_coop(KERN<16, false, true, 2>, d, p, s); \
    if (p.bf == 2 && p.HS == 32) return fb ? launch_coop(KERN<32, true, true, 2>, d, p, s) : launch_coop(KERN<32, false, true, 2>, d, p, s); \
    if (p.HS == 4) AG_LSTM_DISPATCH_HS(KERN, 4);  \
    if (p.HS == 8) AG_LSTM_DISPATCH_HS(KERN, 8);  \
    if (p.HS == 16) AG_LSTM_DISPATCH_HS(KERN, 16);  \
    AG_LSTM_DISPATCH_HS(KERN, 32);      \
  } while (0)

static int run_fwd(const ag_lstm_desc* d, const Plan& p, cudaStream_t s) { AG_LSTM_DISPATCH(lstm_fwd_kernel); }
static int run_bwd(const ag_lstm_desc* d, const Plan& p, cudaStream_t s) { AG_LSTM_DISPATCH(lstm_bwd_kernel); }

}  // namespace ag

namespace ag { namespace lc {
int cluster_fwd(const ag_lstm_desc* d, cudaStream_t s, int* launched);   // lstm_cluster.cu
int cluster_bwd(const ag_lstm_desc* d, cudaStream_t s, int* launched);
} namespace lg {
int gen_fwd(const ag_lstm_desc* d, cudaStream_t s, int* launched);       // lstm_gen.cu
int gen_bwd(const ag_lstm_desc* d, cudaStream_t s, int* launched);
int64_t gen_fwd_ws_bytes(const ag_lstm_desc* d);
int64_t gen_bwd_ws_bytes(const ag_lstm_desc* d);
int gen_batch_cap(const ag_lstm_desc* d, int bwd);
} }

using namespace ag;
extern "C" {

static const char* grid_family(const Plan& p) {
  return p.bf == 2 ? "grid-tcgen05" : p.bf == 1 ? "grid-bf16" : p.res ? "grid-fp32" : "grid-fp32-streamed";
}

int64_t ag_lstm_workspace_bytes(const ag_lstm_desc* d, int32_t bwd) {
  if (!d) return AG_EINVAL;
  return bwd ? lg::gen_bwd_ws_bytes(d) : lg::gen_fwd_ws_bytes(d);
}

int32_t ag_lstm_batch_cap(const ag_lstm_desc* d, int32_t bwd) {
  if (!d) return 0;
  return lg::gen_batch_cap(d, bwd);
}

int ag_lstm_fwd(const ag_lstm_desc* d, void* stream) {
  int rc = check_lstm(d, "ag_lstm_fwd", false);
  if (rc) return rc;
  clear_decline();
  {
    int launched = 0;
    rc = lc::cluster_fwd(d, (cudaStream_t)stream, &launched);
    if (rc || launched) return rc;
    rc = lg::gen_fwd(d, (cudaStream_t)stream, &launched);
    if (rc || launched) return rc;
  }
  Plan p;
  p.res = true; p.bf = 0; p.nst = NST_MIN; p.bsplit = 1; p.HS = 0; p.kcb = 1; p.nbg2 = 1;
  const int nsm = sm_count();
  auto pr_for = [&](int ncta) { return d->F > 0 ? (d->F + 1 + ncta - 1) / ncta : 0; };
  if (d->prec == 2 && d->H % 8 == 0 && d->F % 8 == 0) {
    // tcgen05 variant: 64 or 128 gate rows per CTA (the MMA's N), batch on the 64 TMEM lanes, 3-deep ring
    const int opts[2] = {32, 16};
    for (int i = 0; i < 2 && !p.bf; ++i) {
      const int hs = opts[i];
      if (d->H % hs || (int64_t)d->ndir * (d->H / hs) > nsm) continue;
      const int groups = d->ndir * (d->H / hs);
      int bs = std::max(1, std::min(nsm / groups, (d->B + 15) / 16));
      const int nbg2 = pick_nbg2(d->B, groups * bs);
      const int pr = pr_for((groups * bs) / nbg2);
      const int rb = std::min(64, (bper_of(d->B, bs) + 15) & ~15);
      for (int kcb = 64 / rb; kcb >= 1 && !p.bf; kcb >>= 1) {
        const size_t sm = fwd_smem_tc(d, hs, pr, 3, kcb, bs);
        if (sm > (size_t)smem_optin()) continue;
        AG_CHECK_ARG(d->hbuf16 && (d->F == 0 || d->xbuf16), "ag_lstm_fwd: bf16 mode needs hbuf16 / xbuf16");
        p.bf = 2; p.HS = hs; p.ncta_dir = d->H / hs; p.bsplit = bs; p.PR = pr; p.smem = sm; p.nst = 3; p.kcb = kcb; p.nbg2 = nbg2;
      }
    }
  }
  if (!p.bf && d->prec >= 1 && d->H % 8 == 0 && d->F % 8 == 0) {
    // tensor-core variant: the largest slice whose bf16 weights stay resident (fewest re-reads of the [B, K] operand
    // per step), then split the batch over the remaining SMs (>= 8 samples per CTA)
    const int opts[4] = {32, 16, 8, 4};
    for (int i = 0; i < 4 && !p.bf; ++i) {
      const int hs = opts[i];
      if (d->H % hs || (int64_t)d->ndir * (d->H / hs) > nsm) continue;
      const int groups = d->ndir * (d->H / hs);
      int bs = nsm / groups;
      bs = std::max(1, std::min(bs, (d->B + 7) / 8));
      const int nbg2 = pick_nbg2(d->B, groups * bs);
      const int pr = pr_for((groups * bs) / nbg2);
      const size_t sm16 = fwd_smem(d, hs, pr, true, true);
      if (sm16 > (size_t)smem_optin()) continue;
      AG_CHECK_ARG(d->hbuf16 && (d->F == 0 || d->xbuf16), "ag_lstm_fwd: bf16 mode needs hbuf16 / xbuf16");
      p.bf = 1; p.HS = hs; p.ncta_dir = d->H / hs; p.bsplit = bs; p.PR = pr; p.smem = sm16; p.nbg2 = nbg2;
    }
  }
  if (!p.bf) {
    p.HS = pick_hs(d->H, d->ndir);
    AG_CHECK_ARG(p.HS > 0, "ag_lstm_fwd: H=%d (ndir %d) does not map onto %d SMs", d->H, d->ndir, nsm);
    p.ncta_dir = d->H / p.HS;
    p.PR = pr_for(p.ncta_dir * d->ndir);
    p.smem = fwd_smem(d, p.HS, p.PR, true);
    if (p.smem > (size_t)smem_optin()) { p.res = false; p.smem = fwd_smem(d, p.HS, p.PR, false); }
  }
  AG_CHECK_ARG(p.smem <= (size_t)smem_optin(), "ag_lstm_fwd: needs %zu B of shared memory", p.smem);
  cudaStream_t s = (cudaStream_t)stream;
  AG_CUDA(cudaMemsetAsync(d->barrier, 0, 8 * sizeof(unsigned), s));
  if (d->F > 0) AG_CUDA(cudaMemsetAsync(d->t_end, 0, sizeof(int), s));
  set_path(grid_family(p));
  return run_fwd(d, p, s);
}

int ag_lstm_bwd(const ag_lstm_desc* d, void* stream) {
  int rc = check_lstm(d, "ag_lstm_bwd", true);
  if (rc) return rc;
  clear_decline();
  {
    int launched = 0;
    rc = lc::cluster_bwd(d, (cudaStream_t)stream, &launched);
    if (rc || launched) return rc;
    rc = lg::gen_bwd(d, (cudaStream_t)stream, &launched);
    if (rc || launched) return rc;
  }
  Plan p;
  p.res = true; p.bf = 0; p.nst = NST_MIN; p.bsplit = 1; p.HS = 0; p.kcb = 1; p.nbg2 = 1;
  const int nsm = sm_count();
  auto pr_for = [&](int ncta) { return d->F > 0 ? (d->F + ncta - 1) / ncta : 0; };
  if (d->prec == 2) {
    const int opts[2] = {32, 16};
    for (int i = 0; i < 2 && !p.bf; ++i) {
      const int hs = opts[i];
      if (d->H % hs || (int64_t)d->ndir * (d->H / hs) > nsm) continue;
      const int groups = d->ndir * (d->H / hs);
      int bs = std::max(1, std::min(nsm / groups, (d->B + 15) / 16));
      const int nbg2 = pick_nbg2(d->B, groups * bs);
      const int pr = pr_for((groups * bs) / nbg2);
      const int rb = std::min(64, (bper_of(d->B, bs) + 15) & ~15);
      for (int kcb = 64 / rb; kcb >= 1 && !p.bf; kcb >>= 1) {
        const size_t sm = bwd_smem_tc(d, hs, pr, 3, kcb, bs);
        if (sm > (size_t)smem_optin()) continue;
        AG_CHECK_ARG(d->dgates16 && (d->F == 0 || d->dpx16), "ag_lstm_bwd: bf16 mode needs dgates16 / dpx16");
        p.bf = 2; p.HS = hs; p.ncta_dir = d->H / hs; p.bsplit = bs; p.PR = pr; p.smem = sm; p.nst = 3; p.kcb = kcb; p.nbg2 = nbg2;
      }
    }
  }
  if (!p.bf && d->prec >= 1) {
    const int opts[2] = {32, 16};
    for (int i = 0; i < 2 && !p.bf; ++i) {
      const int hs = opts[i];
      if (d->H % hs || (int64_t)d->ndir * (d->H / hs) > nsm) continue;
      const int groups = d->ndir * (d->H / hs);
      int bs = nsm / groups;
      bs = std::max(1, std::min(bs, (d->B + 7) / 8));
      const int nbg2 = pick_nbg2(d->B, groups * bs);
      const int pr = pr_for((groups * bs) / nbg2);
      const size_t sm16 = bwd_smem(d, hs, pr, true, true);
      if (sm16 > (size_t)smem_optin()) continue;
      AG_CHECK_ARG(d->dgates16 && (d->F == 0 || d->dpx16), "ag_lstm_bwd: bf16 mode needs dgates16 / dpx16");
      p.bf = 1; p.HS = hs; p.ncta_dir = d->H / hs; p.bsplit = bs; p.PR = pr; p.smem = sm16; p.nbg2 = nbg2;
    }
  }
  if (!p.bf) {
    p.HS = pick_hs(d->H, d->ndir);
    AG_CHECK_ARG(p.HS > 0, "ag_lstm_bwd: H=%d (ndir %d) does not map onto %d SMs", d->H, d->ndir, nsm);
    p.ncta_dir = d->H / p.HS;
    p.PR = pr_for(p.ncta_dir * d->ndir);
    p.smem = bwd_smem(d, p.HS, p.PR, true);
    if (p.smem > (size_t)smem_optin()) { p.res = false; p.smem = bwd_smem(d, p.HS, p.PR, false); }
  }
  AG_CHECK_ARG(p.smem <= (size_t)smem_optin(), "ag_lstm_bwd: needs %zu B of shared memory", p.smem);
  cudaStream_t s = (cudaStream_t)stream;
  AG_CUDA(cudaMemsetAsync(d->barrier, 0, 8 * sizeof(unsigned), s));
  set_path(grid_family(p));
  return run_bwd(d, p, s);
}
}
