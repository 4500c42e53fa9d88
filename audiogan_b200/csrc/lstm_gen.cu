// Generator recurrence (audiogan.py:428-460: LSTMCell with output feedback + proj/tanh + stopper + Bernoulli stop) in
// bf16 mode on tcgen05 tensor cores with the recurrent weights RESIDENT ON CHIP for the whole sequence.
//
// CTA (slice, batch group): 32 hidden units = 128 gate rows (the MMA's M) of [whh | wx] (K = H + F, 1224 for the default
// net) and NB = 16 or 32 samples (the MMA's N).  The first KT columns of the weight slice live in TENSOR MEMORY
// (tcgen05.st once; lane = gate row, two k per 32-bit column; A-from-TMEM MMAs), the rest in shared memory (K-major
// core-matrix tiles, A-from-smem MMAs into the same accumulators).  Back-to-back MMAs into one accumulator serialise
// (~45-60 cycles each, measured in lstm_cluster.cu), so the K range is cut into independent chains, one issuing thread
// and one TMEM accumulator each, summed in the epilogue.
// Per step: gates = pre_t + [whh | wx] . [h_{t-1} ; x_{t-1}]  ->  cell update (c in registers)  ->  h_t to global (L2)
//           ->  flag exchange inside the batch group  ->  every CTA pulls h_t [NB, H] into its B operand (it is also
//           next step's operand) and computes ITS rows of x_t = tanh(wp h_t + bp) / the stop logit from shared memory
//           ->  flag exchange  ->  x_t [NB, F] pulled into the B operand.
// The H/32 CTAs of a batch group synchronise through per-CTA release/acquire flags (one L2 round trip, no atomics);
// batch groups are independent unless stop sampling is on, in which case the second exchange is grid-wide so that
// every CTA sees how many samples are still generating (the reference's per-frame host sync, audiogan.py:458-460).
#include "common.cuh"
#include "tc_common.cuh"
#include <algorithm>
#include <stdlib.h>

namespace ag {
namespace lg {

using namespace tc;

constexpr int LT = 256;
constexpr int UPC = 32;                 // hidden units per CTA

__device__ __forceinline__ uint64_t umma_desc_nosw(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= 1ull << 46;
  return d;
}
__device__ __forceinline__ void tc_mma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_st8(uint32_t taddr, const uint32_t (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(v[0]), "r"(v[1]),
               "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
__device__ __forceinline__ void tc_ld8_nowait(uint32_t taddr, uint32_t (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr));
}
__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void cp_async16(void* smem, const void* g) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(smem_u32(smem)), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\n" ::: "memory");
  asm volatile("cp.async.wait_group 0;\n" ::: "memory");
}
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_u32(unsigned* p, unsigned v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;\n" ::"l"(p), "r"(v) : "memory");
}
// publish: everything this CTA wrote before is visible to whoever acquires flag >= v
__device__ __forceinline__ void flag_publish(unsigned* flag, unsigned v) {
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) st_release_u32(flag, v);
}
// the same with ONE fence: bar.sync orders the CTA's stores before thread 0, whose gpu-scope fence + release store is cumulative
__device__ __forceinline__ void flag_publish_light(unsigned* flag, unsigned v) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    st_release_u32(flag, v);
  }
}
// wait until flags[first + i*stride] >= v for i < n (warp 0 polls, one flag per lane and pass), then block-wide sync
__device__ __forceinline__ void flags_wait(const unsigned* flags, int first, int stride, int n, unsigned v) {
  if (threadIdx.x < 32) {
    for (int i = threadIdx.x; i < n; i += 32) {
      const unsigned* p = flags + first + i * stride;
      while (ld_acquire_u32(p) < v) { }
    }
  }
  __syncthreads();
}

// global -> shared bulk copy (TMA, no tensor map) that completes on an mbarrier of this CTA
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t mbar_smem) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem), "l"(src),
               "r"(bytes), "r"(mbar_smem)
               : "memory");
}
__device__ __forceinline__ void fence_proxy_async_full() { asm volatile("fence.proxy.async;" ::: "memory"); }

struct Clk {
  long long acc[7], t0;
  bool on;
  __device__ __forceinline__ void init(bool o) {
    on = o;
    for (int i = 0; i < 7; ++i) acc[i] = 0;
    t0 = 0;
  }
  __device__ __forceinline__ void start() { if (on) t0 = clock64(); }
  __device__ __forceinline__ void lap(int i) {
    if (on) { const long long t = clock64(); acc[i] += t - t0; t0 = t; }
  }
};

struct FwdGeom {
  int KP;        // H + F padded to a multiple of 16 (MMA K steps)
  int KT;        // columns resident in TMEM (multiple of 64, <= H)
  int PR;        // proj / stop rows per CTA in phase 2 (<= 8)
  int nsl;       // CTAs per batch group = H / 32
  int ngroups;   // batch groups
  int xchg;      // h_t / x_t exchange: 1 = flags + TMA bulk copies of bf16 blocks (default), 0 = tagged "LL" words
};

// "LL" exchange (the NCCL low-latency protocol): every 8-byte word carries 4 bytes of payload and the 4-byte step tag,
// written with one 8-byte store and polled by the consumers with 16-byte volatile loads -- data and "it is there" arrive
// in ONE L2 round trip, with no fence, flag or barrier on the critical path (a fence + flag + poll + pull sequence
// costs ~7 k cycles per exchange, measured; two exchanges per step).  Two slots per buffer: a CTA can only be one
// exchange ahead of its slowest peer.
//   LLh [2][B_pad][H/2]  (bf16x2 of h_t, tag)      LLx [2][B_pad][F] (fp32 x_t, tag)     LLa [2][16] (alive count, tag)
// Strong (relaxed, gpu scope) accesses on purpose: weak .cg stores of a lone word can sit in the SM's store-combining
// path indefinitely and weak polling loads then never see them (observed: a hang with stop sampling on).
__device__ __forceinline__ void ll_store2(void* p, uint32_t d0, uint32_t tag) {
  asm volatile("st.relaxed.gpu.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(d0), "r"(tag) : "memory");
}
__device__ __forceinline__ void ll_store4(void* p, uint32_t d0, uint32_t d1, uint32_t tag) {
  asm volatile("st.relaxed.gpu.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(d0), "r"(tag), "r"(d1), "r"(tag) : "memory");
}
__device__ __forceinline__ uint4 ll_load4(const void* p) {
  uint4 v;
  asm volatile("ld.relaxed.gpu.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// Shared memory (1 KB aligned):
//   As  [(KP-KT)/8][128 rows][16 B]   weight columns KT.. (K-major core matrices, LBO = 2048, SBO = 128)
//   Bs  [KP/8][NB rows][16 B]         [h_{t-1} ; x_{t-1} ; 0-pad] bf16 (LBO = NB*16, SBO = 128)
//   W2s [16][H + 8] bf16              this CTA's rows of [wp ; ws] (rows >= np zero: the mma.sync tile is 16 rows)
//   gs  [4][NB][32] fp32              gate exchange (also the phase-2 partial sums [8 warps][8 rows][NB])
//   gen / cnt [NB] int, mbarrier, tmem slot
template <int NB>
__global__ void __launch_bounds__(LT, 1) lstm_gen_fwd_kernel(const ag_lstm_desc d, const FwdGeom gm) {
  constexpr int NCH = NB == 16 ? 4 : 2;            // accumulator chains (NCH * NB = 64 TMEM columns)
  constexpr int JV = NB * UPC / LT;                // consecutive units per thread in the cell update (2 or 4)
  constexpr int TPS = UPC / JV;                    // threads per sample
  constexpr int CSTR = NB * 16 + 16;               // k-chunk stride of Bs: +16 B keeps the chunk-major stores conflict-free
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, q = w & 3, hf = w >> 2;
  const int H = d.H, F = d.F, K1 = H + F, B = d.B, T = d.T, Tcap = d.Tcap;
  const int KP = gm.KP, KT = gm.KT, PR = gm.PR, nsl = gm.nsl;
  const int slice = blockIdx.x % nsl, grp = blockIdx.x / nsl, j0 = slice * UPC, b0 = grp * NB;
  const int nkc = KP / 8, nks = (KP - KT) / 8;
  const int Bpad = gm.ngroups * NB;

  uint8_t* sm = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* As = sm;
  uint8_t* Bs = As + (size_t)nks * 2048;
  __nv_bfloat16* W2s = reinterpret_cast<__nv_bfloat16*>(Bs + (size_t)nkc * CSTR);
  float* gs = reinterpret_cast<float*>(W2s + (size_t)16 * (H + 8));
  int* gen = reinterpret_cast<int*>(gs + 4 * NB * UPC);
  int* cnt = gen + NB;
  uint64_t* mma_done = reinterpret_cast<uint64_t*>(cnt + NB);
  uint64_t* hfull = mma_done + 1;                  // TMA exchange: h_t / x_{t-1} block has landed in Bs
  uint64_t* xfull = mma_done + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mma_done + 3);

  const bool tma_x = gm.xchg != 0;
  // k-chunk stride of the B operand: contiguous chunks when TMA writes it (one bulk copy moves a whole block), padded by
  // 16 B when threads store chunk-major (bank conflicts)
  const uint32_t bstr = tma_x ? (uint32_t)NB * 16u : (uint32_t)CSTR;
  uint2* LLh = reinterpret_cast<uint2*>(d.ll_ws);
  uint2* LLx = LLh + (size_t)2 * Bpad * (H / 2);
  uint2* LLa = LLx + (size_t)2 * Bpad * F;
  // TMA exchange: Hx [2 slots][group][H/8 k-chunks][NB rows][16 B] bf16 -- the image of the B operand's h part, so a consumer
  // fetches one k-chunk (NB * 16 bytes) per bulk copy; Xx likewise [2][group][F/8][NB][16 B]; one release flag per producer
  // CTA, slot and exchange (value = step tag)
  uint8_t* Hx = reinterpret_cast<uint8_t*>(d.ll_ws);
  uint8_t* Xx = Hx + (size_t)2 * Bpad * H * 2;
  unsigned* flagH = nullptr;
  unsigned* flagX = nullptr;
  if (tma_x) {
    LLa = reinterpret_cast<uint2*>(Xx + (size_t)2 * Bpad * F * 2);
    flagH = reinterpret_cast<unsigned*>(LLa + 32);
    flagX = flagH + 2 * gm.ngroups * 32;
  }

  if (tid == 0) {
    mbar_init(mma_done, NCH);
    mbar_init(hfull, 1);
    mbar_init(xfull, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // TMEM: columns [0, KT/2) = A (weights), then NCH accumulators of NB columns
  if (w == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tmem_d = tmem + (uint32_t)(KT / 2);
  {
    // local row lr = 32 q + lane  <->  gate row q*H + j0 + lane;  warps w and w + 4 split the column range
    const float* src = d.w1 + ((int64_t)q * H + j0 + lane) * K1;
    // 4 column blocks per pass: 16 loads in flight before the first tcgen05.st (the passes are latency-bound otherwise)
    for (int k0 = hf * 16; k0 < KT; k0 += 128) {
      float4 a[4][4];
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int e = 0; e < 4; ++e)
          a[u][e] = (k0 + 32 * u < KT) ? __ldg(reinterpret_cast<const float4*>(src + k0 + 32 * u) + e) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (k0 + 32 * u < KT) {
          uint32_t v[8];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            v[2 * e] = pack_bf16(a[u][e].x, a[u][e].y);
            v[2 * e + 1] = pack_bf16(a[u][e].z, a[u][e].w);
          }
          tc_st8(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)((k0 + 32 * u) >> 1), v);
        }
      }
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
#pragma unroll 4
    for (int cell = tid; cell < 128 * nks; cell += LT) {
      const int cc = cell >> 7, lr = cell & 127, k = KT + cc * 8;
      const float* s2 = d.w1 + ((int64_t)(lr >> 5) * H + j0 + (lr & 31)) * K1 + k;
      float4 lo = make_float4(0.f, 0.f, 0.f, 0.f), hi = lo;
      if (k + 8 <= K1) {                       // rows are 16-byte aligned (K1 % 4 == 0) and k % 8 == 0
        lo = __ldg(reinterpret_cast<const float4*>(s2));
        hi = __ldg(reinterpret_cast<const float4*>(s2) + 1);
      } else if (k + 4 <= K1) {
        lo = __ldg(reinterpret_cast<const float4*>(s2));
      }
      *reinterpret_cast<uint4*>(As + (size_t)cc * 2048 + lr * 16) =
          make_uint4(pack_bf16(lo.x, lo.y), pack_bf16(lo.z, lo.w), pack_bf16(hi.x, hi.y), pack_bf16(hi.z, hi.w));
    }
  }
  // phase-2 rows of this CTA
  const int p0 = slice * PR;
  int np = min(PR, F + 1 - p0);
  if (np < 0) np = 0;
  const bool owns_logit = np > 0 && p0 + np == F + 1;
  for (int idx = tid; idx < 16 * (H + 8) / 8; idx += LT) *reinterpret_cast<uint4*>(W2s + idx * 8) = make_uint4(0u, 0u, 0u, 0u);
  __syncthreads();
#pragma unroll 4
  for (int idx = tid; idx < np * (H / 4); idx += LT) {
    const int r = idx / (H / 4), k4 = idx - r * (H / 4);
    const float4 a = __ldg(reinterpret_cast<const float4*>(d.w2 + (int64_t)(p0 + r) * H) + k4);
    *reinterpret_cast<uint2*>(W2s + r * (H + 8) + 4 * k4) = make_uint2(pack_bf16(a.x, a.y), pack_bf16(a.z, a.w));
  }
  for (uint32_t i = tid * 16; i < (uint32_t)nkc * bstr; i += LT * 16) *reinterpret_cast<uint4*>(Bs + i) = make_uint4(0u, 0u, 0u, 0u);
  for (int b = tid; b < NB; b += LT) { gen[b] = 1; cnt[b] = 0; }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  const uint32_t idesc = umma_idesc(128, NB, 0, 0);
  const uint32_t as_local = smem_u32(As), bs_local = smem_u32(Bs);
  const int64_t hstr = (int64_t)(Tcap + 2) * H, gstr = (int64_t)Tcap * 4 * H, cstr = (int64_t)Tcap * H, xstr = (int64_t)(Tcap + 1) * F;
  __nv_bfloat16* hb16 = reinterpret_cast<__nv_bfloat16*>(d.hbuf16);
  __nv_bfloat16* xb16 = reinterpret_cast<__nv_bfloat16*>(d.xbuf16);
  const bool want_h32 = !(d.flags & AG_LSTM_BF16_H_ONLY);
  // cell-update items: sample bl, units jv .. jv + JV - 1
  const int bl = tid / TPS, jv = (tid % TPS) * JV, bme = b0 + bl;
  float cst[JV];
#pragma unroll
  for (int e = 0; e < JV; ++e) cst[e] = 0.f;
  int steps_run = T;
  uint32_t nmma = 0;
  Clk ck;
  ck.init(d.dbg != nullptr);
  const long long tstart = clock64();

  float pre[4][JV];
  auto load_pre = [&](int t) {
    const float* pp = d.pre + bme * gstr + (int64_t)t * 4 * H + j0 + jv;
#pragma unroll
    for (int qq = 0; qq < 4; ++qq)
#pragma unroll
      for (int e = 0; e < JV; ++e) pre[qq][e] = bme < B ? __ldg(pp + qq * H + e) : 0.f;
  };
  load_pre(0);
  // chain w of the gate product: TMEM k-steps [w*nt, (w+1)*nt) and shared-memory k-steps w, w + NCH, ...; the columns of
  // h (k < H) are issued as soon as h_{t-1} is in place, the columns of x (k >= H) once x_{t-1} has arrived
  const int nt = (KT / 16) / NCH, nsh = (H - KT) / 16, nsx = (KP - H) / 16;
  auto issue_h_part = [&]() {
    if (lane == 0 && w < NCH) {
      tc_fence_after();
      uint32_t ta = tmem + (uint32_t)(w * nt * 8);
      uint64_t db = umma_desc_nosw(bs_local, bstr, 128) + (uint64_t)(w * nt * (2 * bstr / 16));
      for (int kk = 0; kk < nt; ++kk) {
        tc_mma_ts(tmem_d + NB * w, ta, db, idesc, kk ? 1u : 0u);
        ta += 8;
        db += 2 * bstr / 16;
      }
      for (int kk = w; kk < nsh; kk += NCH)
        tc_mma(tmem_d + NB * w, umma_desc_nosw(as_local + (uint32_t)kk * 4096, 2048, 128),
               umma_desc_nosw(bs_local + (uint32_t)(KT / 8 + 2 * kk) * bstr, bstr, 128), idesc, 1u);
    }
  };
  auto issue_x_part = [&]() {
    if (lane == 0 && w < NCH) {
      tc_fence_after();
      for (int kk = nsh + w; kk < nsh + nsx; kk += NCH)
        tc_mma(tmem_d + NB * w, umma_desc_nosw(as_local + (uint32_t)kk * 4096, 2048, 128),
               umma_desc_nosw(bs_local + (uint32_t)(KT / 8 + 2 * kk) * bstr, bstr, 128), idesc, 1u);
      tc_commit(mma_done);
    }
  };
  issue_h_part();                                   // step 0: h_{-1} = 0 (keeps the accumulators defined)
  const float b2_me = (tid < np * NB) ? d.b2[p0 + tid / NB] : 0.f;

  for (int s = 0; s < T; ++s) {
    ck.start();
    const int t = s;
    const uint32_t tag = (uint32_t)(s + 1), ptag = (uint32_t)s, slot = (uint32_t)s & 1u, pslot = slot ^ 1u;
    if (s > 0) {
      if (d.u) {
        // samples still generating after step s - 1, summed over the batch groups (audiogan.py:458-460)
        unsigned a = 0;
        for (int g2 = 0; g2 < gm.ngroups; ++g2) {
          uint4 v;
          do { v = ll_load4(LLa + (size_t)pslot * 16 + (g2 & ~1)); } while (((g2 & 1) ? v.w : v.y) != ptag);
          a += (g2 & 1) ? v.z : v.x;
        }
        if (a == 0u) { steps_run = s; break; }
      }
      if (tma_x) {
        // x_{t-1}: wait for the flags of the CTAs that own projection rows, then one bulk copy per k-chunk
        if (w == 0) {
          for (int pr = lane; pr * PR < F; pr += 32) {
            const unsigned* fp = flagX + ((size_t)pslot * gm.ngroups + grp) * 32 + pr;
            while (ld_acquire_u32(fp) < ptag) { }
          }
          __syncwarp();
          __threadfence();
          fence_proxy_async_full();
          if (lane == 0) {
            const uint8_t* xs = Xx + ((size_t)pslot * gm.ngroups + grp) * (F / 8) * NB * 16;
            mbar_arrive_expect_tx(xfull, (uint32_t)(F / 8) * NB * 16);
            bulk_g2s(bs_local + (uint32_t)(H / 8) * bstr, xs, (uint32_t)(F / 8) * NB * 16, smem_u32(xfull));
          }
        }
        mbar_wait(xfull, (uint32_t)(s - 1) & 1u);
      } else {
      // x_{t-1} [NB, F]: 8 consecutive rows p of one sample = one 16-byte k-chunk of the B operand
      const uint2* src = LLx + ((size_t)pslot * Bpad + b0) * F;
      if (w == 0) {                                     // one word per producer CTA first (see the h exchange)
        for (int pr = lane; pr * PR < F; pr += 32) {
          const int pl = min(F, (pr + 1) * PR) - 1;     // last proj row of producer pr
          const uint2* p = src + (size_t)(NB - 1) * F + (pl & ~1);
          uint4 v;
          do { v = ll_load4(p); } while (((pl & 1) ? v.w : v.y) != ptag);
        }
      }
      __syncthreads();
      constexpr int XC = (NB * 25 + LT - 1) / LT;       // chunks per thread for F = 200 (more passes for a larger F)
      for (int base = 0; base < NB * (F / 8); base += LT * XC) {
        // all loads of a pass are issued before the first tag is looked at: one L2 round trip per pass, not per chunk
        uint4 v[XC][4];
        unsigned pending = 0;
#pragma unroll
        for (int i = 0; i < XC; ++i)
          if (base + i * LT + tid < NB * (F / 8)) pending |= 1u << i;
        while (pending) {
#pragma unroll
          for (int i = 0; i < XC; ++i) {
            if (pending & (1u << i)) {
              const int idx = base + i * LT + tid, b = idx / (F / 8), c = idx - b * (F / 8);   // chunk fastest: coalesced loads
              const uint2* p = src + (size_t)b * F + c * 8;
              v[i][0] = ll_load4(p); v[i][1] = ll_load4(p + 2); v[i][2] = ll_load4(p + 4); v[i][3] = ll_load4(p + 6);
            }
          }
#pragma unroll
          for (int i = 0; i < XC; ++i) {
            if (pending & (1u << i)) {
              bool ok = true;
#pragma unroll
              for (int e = 0; e < 4; ++e) ok = ok && v[i][e].y == ptag && v[i][e].w == ptag;
              if (ok) {
                const int idx = base + i * LT + tid, b = idx / (F / 8), c = idx - b * (F / 8);
                *reinterpret_cast<uint4*>(Bs + (size_t)(H / 8 + c) * bstr + b * 16) = make_uint4(
                    pack_bf16(__uint_as_float(v[i][0].x), __uint_as_float(v[i][0].z)), pack_bf16(__uint_as_float(v[i][1].x), __uint_as_float(v[i][1].z)),
                    pack_bf16(__uint_as_float(v[i][2].x), __uint_as_float(v[i][2].z)), pack_bf16(__uint_as_float(v[i][3].x), __uint_as_float(v[i][3].z)));
                pending &= ~(1u << i);
              }
            }
          }
        }
      }
      fence_proxy_async();
      __syncthreads();
      }
    }
    ck.lap(0);
    issue_x_part();
    mbar_wait(mma_done, nmma & 1u);
    ++nmma;
    tc_fence_after();
    ck.lap(1);
    {
      // warps w and w + 4 share TMEM lanes 32 (w & 3) ..: each takes half of the NB columns of every chain
      constexpr int CW = NB / 2;
      float acc[CW];
#pragma unroll
      for (int i = 0; i < CW; ++i) acc[i] = 0.f;
#pragma unroll
      for (int ch = 0; ch < NCH; ++ch) {
#pragma unroll
        for (int c8 = 0; c8 < CW / 8; ++c8) {
          uint32_t v[8];
          tc_ld8_nowait(tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)(ch * NB + hf * CW + c8 * 8), v);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          for (int i = 0; i < 8; ++i) acc[c8 * 8 + i] += __uint_as_float(v[i]);
        }
      }
#pragma unroll
      for (int i = 0; i < CW; ++i) gs[(q * NB + hf * CW + i) * UPC + lane] = acc[i];
    }
    tc_fence_before();
    __syncthreads();
    float gq[4][JV], cv[JV], hv[JV];
#pragma unroll
    for (int e = 0; e < JV; ++e) {
      const float gi = fmaf(tanh_approx((gs[(0 * NB + bl) * UPC + jv + e] + pre[0][e]) * 0.5f), 0.5f, 0.5f);
      const float gf = fmaf(tanh_approx((gs[(1 * NB + bl) * UPC + jv + e] + pre[1][e]) * 0.5f), 0.5f, 0.5f);
      const float gg = tanh_approx(gs[(2 * NB + bl) * UPC + jv + e] + pre[2][e]);
      const float go = fmaf(tanh_approx((gs[(3 * NB + bl) * UPC + jv + e] + pre[3][e]) * 0.5f), 0.5f, 0.5f);
      const float c = gf * cst[e] + gi * gg;
      cst[e] = c;
      gq[0][e] = gi; gq[1][e] = gf; gq[2][e] = gg; gq[3][e] = go;
      cv[e] = c;
      hv[e] = go * tanh_approx(c);
    }
    if (tma_x) {
      // h_t slice -> Hx in the consumers' operand layout (k-chunk major), then this CTA's release flag
      uint8_t* dst = Hx + (((size_t)slot * gm.ngroups + grp) * (H / 8) + (size_t)((j0 + jv) >> 3)) * NB * 16 + bl * 16 + ((j0 + jv) & 7) * 2;
      if (JV == 2) *reinterpret_cast<uint32_t*>(dst) = pack_bf16(hv[0], hv[1]);
      else *reinterpret_cast<uint2*>(dst) = make_uint2(pack_bf16(hv[0], hv[1]), pack_bf16(hv[JV - 2], hv[JV - 1]));
      flag_publish_light(flagH + ((size_t)slot * gm.ngroups + grp) * 32 + slice, tag);
    } else {
      // h_t slice -> LL buffer (what the peers wait for), samples past B included (zeros keep their operand rows clean)
      uint2* dst = LLh + ((size_t)slot * Bpad + bme) * (H / 2) + (j0 + jv) / 2;
      if (JV == 2) ll_store2(dst, pack_bf16(hv[0], hv[1]), tag);
      else ll_store4(dst, pack_bf16(hv[0], hv[1]), pack_bf16(hv[JV - 2], hv[JV - 1]), tag);
    }
    ck.lap(2);
    // off the critical path: h_t and the saved state for BPTT, next step's input projections
    if (bme < B) {
      const int64_t ho = bme * hstr + (int64_t)(t + 1) * H + j0 + jv;
#pragma unroll
      for (int e = 0; e < JV; e += 2) {
        if (want_h32) *reinterpret_cast<float2*>(d.hbuf + ho + e) = make_float2(hv[e], hv[e + 1]);
        *reinterpret_cast<uint32_t*>(hb16 + ho + e) = pack_bf16(hv[e], hv[e + 1]);
      }
      if (d.cbuf) {
#pragma unroll
        for (int e = 0; e < JV; e += 2) *reinterpret_cast<float2*>(d.cbuf + bme * cstr + (int64_t)t * H + j0 + jv + e) = make_float2(cv[e], cv[e + 1]);
      }
      if (d.gates) {
        float* gp = d.gates + bme * gstr + (int64_t)t * 4 * H + j0 + jv;
#pragma unroll
        for (int qq = 0; qq < 4; ++qq)
#pragma unroll
          for (int e = 0; e < JV; e += 2) *reinterpret_cast<float2*>(gp + qq * H + e) = make_float2(gq[qq][e], gq[qq][e + 1]);
      }
    }
    if (s + 1 < T) load_pre(s + 1);
    ck.lap(3);
    // h_t [NB, H] of the whole batch group
    if (tma_x) {
      if (w == 0) {
        for (int pr = lane; pr < nsl; pr += 32) {
          const unsigned* fp = flagH + ((size_t)slot * gm.ngroups + grp) * 32 + pr;
          while (ld_acquire_u32(fp) < tag) { }
        }
        __syncwarp();
        __threadfence();
        fence_proxy_async_full();
        if (lane == 0) mbar_arrive_expect_tx(hfull, (uint32_t)(H / 8) * NB * 16);
        __syncwarp();
        if (lane < 4) {                                      // four bulk copies of a quarter of the block each
          const uint32_t q4 = (uint32_t)(H / 8) * NB * 16 / 4;
          const uint8_t* hsrc = Hx + ((size_t)slot * gm.ngroups + grp) * (H / 8) * NB * 16;
          bulk_g2s(bs_local + lane * q4, hsrc + (size_t)lane * q4, q4, smem_u32(hfull));
        }
      }
      mbar_wait(hfull, (uint32_t)s & 1u);
    } else {
      // 8 units = 4 LL words = one 16-byte k-chunk
      const uint2* src = LLh + ((size_t)slot * Bpad + b0) * (H / 2);
      // 128 CTAs re-polling 64 KB each saturate the L2 (measured: the exchange took ~8 k cycles that way), so one warp
      // first waits on ONE word per producer CTA; the tagged bulk read below then succeeds on its first pass almost always
      if (w == 0) {
        for (int pr = lane; pr < nsl; pr += 32) {
          const uint2* p = src + (size_t)(NB - 1) * (H / 2) + pr * (UPC / 2) + (UPC / 2 - 2);
          uint4 v;
          do { v = ll_load4(p); } while (v.w != tag);
        }
      }
      __syncthreads();
      constexpr int HC = 8;                             // chunks in flight per thread and pass
      for (int base = 0; base < NB * (H / 8); base += LT * HC) {
        uint4 v[HC][2];
        unsigned pending = 0;
#pragma unroll
        for (int i = 0; i < HC; ++i)
          if (base + i * LT + tid < NB * (H / 8)) pending |= 1u << i;
        while (pending) {
#pragma unroll
          for (int i = 0; i < HC; ++i) {
            if (pending & (1u << i)) {
              const int idx = base + i * LT + tid, b = idx / (H / 8), c = idx - b * (H / 8);
              const uint2* p = src + (size_t)b * (H / 2) + c * 4;
              v[i][0] = ll_load4(p); v[i][1] = ll_load4(p + 2);
            }
          }
#pragma unroll
          for (int i = 0; i < HC; ++i) {
            if ((pending & (1u << i)) && v[i][0].y == tag && v[i][0].w == tag && v[i][1].y == tag && v[i][1].w == tag) {
              const int idx = base + i * LT + tid, b = idx / (H / 8), c = idx - b * (H / 8);
              *reinterpret_cast<uint4*>(Bs + (size_t)c * bstr + b * 16) = make_uint4(v[i][0].x, v[i][0].z, v[i][1].x, v[i][1].z);
              pending &= ~(1u << i);
            }
          }
        }
      }
      fence_proxy_async();
      __syncthreads();
    }
    ck.lap(4);
    if (s + 1 < T) issue_h_part();                  // next step's whh . h_t runs under phase 2 and the x exchange
    // phase 2: rows p0 .. p0 + np of [wp ; ws] against h_t on mma.sync (m16n8k16; warp w takes k in [w H/8, (w+1) H/8))
    if (np > 0) {
      const int g8 = lane >> 2, tq = lane & 3, kw = H / 8;
      float acc2[NB / 8][4], acc3[NB / 8][4];          // even / odd k-steps: independent mma.sync chains
#pragma unroll
      for (int n = 0; n < NB / 8; ++n) {
        acc2[n][0] = 0.f; acc2[n][1] = 0.f; acc2[n][2] = 0.f; acc2[n][3] = 0.f;
        acc3[n][0] = 0.f; acc3[n][1] = 0.f; acc3[n][2] = 0.f; acc3[n][3] = 0.f;
      }
      const __nv_bfloat16* wr = W2s + g8 * (H + 8) + 2 * tq;
      for (int k0 = w * kw; k0 < (w + 1) * kw; k0 += 16) {
        uint32_t a[4];
        a[0] = *reinterpret_cast<const uint32_t*>(wr + k0);
        a[1] = *reinterpret_cast<const uint32_t*>(wr + 8 * (H + 8) + k0);
        a[2] = *reinterpret_cast<const uint32_t*>(wr + k0 + 8);
        a[3] = *reinterpret_cast<const uint32_t*>(wr + 8 * (H + 8) + k0 + 8);
        const uint8_t* bp = Bs + (size_t)(k0 / 8) * bstr + g8 * 16 + tq * 4;
#pragma unroll
        for (int n = 0; n < NB / 8; ++n) {
          const uint32_t bb0 = *reinterpret_cast<const uint32_t*>(bp + n * 128);
          const uint32_t bb1 = *reinterpret_cast<const uint32_t*>(bp + n * 128 + bstr);
          if ((k0 >> 4) & 1) mma_bf16_16816(acc3[n], a, bb0, bb1);
          else mma_bf16_16816(acc2[n], a, bb0, bb1);
        }
      }
      // rows 0..7 of the tile (c0, c1) are this CTA's rows; rows 8..15 are padding
#pragma unroll
      for (int n = 0; n < NB / 8; ++n) {
        gs[(w * 8 + g8) * NB + n * 8 + 2 * tq] = acc2[n][0] + acc3[n][0];
        gs[(w * 8 + g8) * NB + n * 8 + 2 * tq + 1] = acc2[n][1] + acc3[n][1];
      }
    }
    __syncthreads();
    if (tid < np * NB) {
      const int r = tid / NB, b2 = tid - r * NB, b = b0 + b2, p = p0 + r;
      float v = b2_me;
#pragma unroll
      for (int ww = 0; ww < 8; ++ww) v += gs[(ww * 8 + r) * NB + b2];
      if (p < F) {
        const float xv = 1.f - __fdividef(2.f, 1.f + __expf(2.f * v));     // tanh, abs error ~1e-7
        if (tma_x)
          *reinterpret_cast<__nv_bfloat16*>(Xx + (((size_t)slot * gm.ngroups + grp) * (F / 8) + (size_t)(p >> 3)) * NB * 16 + b2 * 16 + (p & 7) * 2) =
              __float2bfloat16(xv);
        else
          ll_store2(LLx + ((size_t)slot * Bpad + b) * F + p, __float_as_uint(xv), tag);
        if (b < B) {
          d.xbuf[b * xstr + (int64_t)(t + 1) * F + p] = xv;
          xb16[b * xstr + (int64_t)(t + 1) * F + p] = __float2bfloat16(xv);
        }
      } else if (b < B) {
        if (d.sbuf) d.sbuf[b * (int64_t)Tcap + t] = v;
        const int stop = (d.u && d.u[b * (int64_t)Tcap + t] < sigmoidf_(v)) ? 1 : 0;
        if (d.stop) d.stop[b * (int64_t)Tcap + t] = stop;
        if (gen[b2]) cnt[b2] += 1;
        if (stop) gen[b2] = 0;
      }
    }
    if (tma_x && np > 0 && p0 < F) flag_publish_light(flagX + ((size_t)slot * gm.ngroups + grp) * 32 + slice, tag);
    if (owns_logit && d.u) {
      __syncthreads();
      if (tid == 0) {
        unsigned a = 0;
        for (int b2 = 0; b2 < NB; ++b2) a += (b0 + b2 < B && gen[b2]) ? 1u : 0u;
        ll_store2(LLa + (size_t)slot * 16 + grp, a, tag);
      }
    }
    ck.lap(5);
  }
  __syncthreads();
  if (owns_logit && d.glen)
    for (int b2 = tid; b2 < NB; b2 += LT) if (b0 + b2 < B) d.glen[b0 + b2] = cnt[b2];
  if (blockIdx.x == 0 && tid == 0) *reinterpret_cast<volatile int*>(d.t_end) = steps_run;
  if (d.dbg && tid == 0) {
    long long* qd = d.dbg + (int64_t)blockIdx.x * 8;
    for (int i = 0; i < 7; ++i) qd[i] = ck.acc[i];
    qd[7] = clock64() - tstart;
  }
  // the MMAs issued for a step that never ran (early exit) must retire before the tensor memory is released
  if (lane == 0 && w < NCH) tc_commit(mma_done);
  mbar_wait(mma_done, nmma & 1u);
  tc_fence_before();
  __syncthreads();
  if (w == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}

// ====================================================================================== backward (BPTT)
// Per step t (descending):  [dh_pre ; dx_pre] = [whh | wx]^T dgates_{t+1}          (M = H + F rows, K = 4H)
//                           dpx_t = (dx_ext_t + dx_pre) (1 - x_t^2),  dpx_t[F] = ds_ext_t
//                           dh_t  = dh_pre + [wp ; ws]^T dpx_t   ->  cell backward  ->  dgates_t
// The CTA that owns 32 hidden units owns their 128 gate rows, i.e. a K-SLICE of the big product: its A operand is
// [whh | wx]^T restricted to those 128 columns ((H + F) x 128, ten 128-row M-tiles: five resident in tensor memory,
// five in shared memory) and its B operand is its OWN dgates_{t+1} slice -- no all-gather of the gate gradients.  The
// (H + F) x NB partial sums are reduce-scattered through the L2 (bf16 pairs + step tag, "LL" words): every CTA adds the
// 32 incoming blocks of its 32 units, the owners of the x rows add theirs, form dpx_t and publish it (second, small LL
// exchange); [wp ; ws]^T dpx_t for the CTA's own units runs on mma.sync from shared memory.
//   LLp [2][ngroups][nsl sources][MR rows][8]  (bf16x2 of two samples, tag)      LLd [2][B_pad][F] (fp32 dpx, tag)
struct BwdGeom {
  int MR;        // H + F rows of the transposed product
  int PR;        // x rows per owning CTA
  int nsl, ngroups;
  int FP;        // F + 1 rounded up to 8
};
constexpr int BT_TMEM = 5;                      // M-tiles of A in tensor memory (64 columns each); the rest in shared memory
constexpr int BNB = 16;                         // samples per CTA

__device__ __forceinline__ uint2 ll_load2(const void* p) {
  uint2 v;
  asm volatile("ld.relaxed.gpu.global.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p) : "memory");
  return v;
}

__global__ void __launch_bounds__(LT, 1) lstm_gen_bwd_kernel(const ag_lstm_desc d, const BwdGeom gm) {
  constexpr int NB = BNB;
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, q = w & 3, hf = w >> 2;
  const int H = d.H, F = d.F, FP = gm.FP, KW = 4 * H + FP, B = d.B, T = d.T, Tcap = d.Tcap;
  const int MR = gm.MR, PR = gm.PR, nsl = gm.nsl;
  const int slice = blockIdx.x % nsl, grp = blockIdx.x / nsl, j0 = slice * UPC, b0 = grp * NB;
  const int NMT = (MR + 127) / 128, nst = NMT - BT_TMEM;     // M-tiles, of which in shared memory
  const int Bpad = gm.ngroups * NB;
  const int WLD = FP + 8;                                    // row stride (bf16) of WpT_s / dpxs

  uint8_t* sm = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* As = sm;                                                         // [nst][16 chunks][128 rows][16 B]
  uint8_t* Bop = As + (size_t)nst * 32768;                                  // [16 chunks][NB rows][16 B]
  __nv_bfloat16* WpT = reinterpret_cast<__nv_bfloat16*>(Bop + 16 * NB * 16);   // [32 units][WLD]: [wp[:, j] | ws[j] | 0]
  __nv_bfloat16* dpxs = WpT + (size_t)UPC * WLD;                            // [NB][WLD]: dpx_t (col F = ds_t)
  float* dhs = reinterpret_cast<float*>(dpxs + (size_t)NB * WLD);           // [NB][33]
  uint64_t* mma_done = reinterpret_cast<uint64_t*>(dhs + NB * 33);     // NB * 33 floats: 8-byte aligned
  uint64_t* mma_x = mma_done + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mma_x + 1);

  uint2* LLp = reinterpret_cast<uint2*>(d.ll_ws);
  uint2* LLd = LLp + (size_t)2 * gm.ngroups * nsl * MR * 8;

  if (tid == 0) {
    mbar_init(mma_done, 4);
    mbar_init(mma_x, 4);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (w == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tmem_d = tmem + BT_TMEM * 64;               // accumulator of M-tile mt: columns mt * NB
  // A[m][k = lr]: m < H -> whh[grow(lr)][m] = w1t[m][grow(lr)];  m >= H -> wx[grow(lr)][m - H] = wxt[m - H][grow(lr)],
  // grow(lr) = (lr / 32) * H + j0 + lr % 32: 16 consecutive lr are 16 consecutive floats
  auto arow = [&](int m, int lr0) -> const float* {
    const int go = (lr0 >> 5) * H + j0 + (lr0 & 31);
    return m < H ? d.w1t + (int64_t)m * KW + go : d.wxt + (int64_t)(m - H) * 4 * H + go;
  };
  for (int it = hf; it < BT_TMEM * 8; it += 8) {             // (tile, k-step) pairs, 4 per pass: 16 loads in flight
    float4 a[4][4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int id = it + 2 * u, mt = id >> 3, kk = id & 7, m = mt * 128 + q * 32 + lane;
      const float* src = arow(min(m, MR - 1), kk * 16);
#pragma unroll
      for (int e = 0; e < 4; ++e) a[u][e] = (id < BT_TMEM * 8 && m < MR) ? __ldg(reinterpret_cast<const float4*>(src) + e) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int id = it + 2 * u, mt = id >> 3, kk = id & 7;
      if (id < BT_TMEM * 8) {
        uint32_t v[8];
#pragma unroll
        for (int e = 0; e < 4; ++e) { v[2 * e] = pack_bf16(a[u][e].x, a[u][e].y); v[2 * e + 1] = pack_bf16(a[u][e].z, a[u][e].w); }
        tc_st8(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(mt * 64 + kk * 8), v);
      }
    }
  }
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
#pragma unroll 4
  for (int cell = tid; cell < nst * 16 * 128; cell += LT) {
    const int r = cell & 127, cc = (cell >> 7) & 15, ts = cell >> 11, m = (BT_TMEM + ts) * 128 + r;
    float4 lo = make_float4(0.f, 0.f, 0.f, 0.f), hi = lo;
    if (m < MR) {
      const float* src = arow(m, cc * 8);
      lo = __ldg(reinterpret_cast<const float4*>(src));
      hi = __ldg(reinterpret_cast<const float4*>(src) + 1);
    }
    *reinterpret_cast<uint4*>(As + (size_t)ts * 32768 + cc * 2048 + r * 16) =
        make_uint4(pack_bf16(lo.x, lo.y), pack_bf16(lo.z, lo.w), pack_bf16(hi.x, hi.y), pack_bf16(hi.z, hi.w));
  }
  for (int idx = tid; idx < UPC * WLD; idx += LT) {
    const int j = idx / WLD, k = idx - j * WLD;
    WpT[idx] = __float2bfloat16(k < FP ? d.w1t[(int64_t)(j0 + j) * KW + 4 * H + k] : 0.f);
  }
  for (int idx = tid; idx < NB * WLD; idx += LT) dpxs[idx] = __float2bfloat16(0.f);
  for (int idx = tid; idx < 16 * NB * 16 / 4; idx += LT) reinterpret_cast<uint32_t*>(Bop)[idx] = 0u;
  // x rows owned by this CTA
  const int p0 = slice * PR;
  int np = min(PR, F - p0);
  if (np < 0) np = 0;
  const bool owns_last = np > 0 && p0 + np == F;             // also writes column F (ds) and the pad columns of dpx
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  const uint32_t idesc = umma_idesc(128, NB, 0, 0);
  const uint32_t as_local = smem_u32(As), bop_local = smem_u32(Bop);
  const int64_t gstr = (int64_t)Tcap * 4 * H, cstr = (int64_t)Tcap * H, xstr = (int64_t)(Tcap + 1) * F;
  __nv_bfloat16* dg16 = reinterpret_cast<__nv_bfloat16*>(d.dgates16);
  const bool want_f32 = !(dg16 && (d.flags & AG_LSTM_BF16_DGATES_ONLY));
  __nv_bfloat16* dp16 = reinterpret_cast<__nv_bfloat16*>(d.dpx16);
  // cell-backward items: sample bl, units jv, jv + 1;   reduce items: row rr, sample pair sp
  const int bl = tid >> 4, jv = (tid & 15) * 2, bme = b0 + bl;
  const int rr = tid >> 3, sp = tid & 7;
  float dcs[2] = {0.f, 0.f};
  uint32_t nmma = 0;
  Clk ck;
  ck.init(d.dbg != nullptr);
  const long long tstart = clock64();

  for (int s = 0; s < T; ++s) {
    ck.start();
    const int t = T - 1 - s;
    const uint32_t tag = (uint32_t)(s + 1), slot = (uint32_t)s & 1u;
    // inputs of the cell backward and of dpx: independent of the recurrence, in flight under the MMAs
    float2 pg[4], pc = make_float2(0.f, 0.f), pcp = pc;
#pragma unroll
    for (int qq = 0; qq < 4; ++qq) pg[qq] = pc;
    if (bme < B) {
      const float* gp = d.gates + bme * gstr + (int64_t)t * 4 * H + j0 + jv;
#pragma unroll
      for (int qq = 0; qq < 4; ++qq) pg[qq] = __ldg(reinterpret_cast<const float2*>(gp + qq * H));
      pc = __ldg(reinterpret_cast<const float2*>(d.cbuf + bme * cstr + (int64_t)t * H + j0 + jv));
      if (t > 0) pcp = __ldg(reinterpret_cast<const float2*>(d.cbuf + bme * cstr + (int64_t)(t - 1) * H + j0 + jv));
    }
    float xo[2] = {0.f, 0.f}, dxe[2] = {0.f, 0.f};
    if (tid < np * 8) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int b = b0 + 2 * sp + e;
        if (b < B) {
          xo[e] = __ldg(d.xbuf + b * xstr + (int64_t)(t + 1) * F + p0 + rr);
          if (d.dx_ext) dxe[e] = __ldg(d.dx_ext + (b * (int64_t)Tcap + t) * F + p0 + rr);
        }
      }
    }
    float dsv = 0.f;
    if (tid < NB && b0 + tid < B && d.ds_ext) dsv = __ldg(d.ds_ext + (b0 + tid) * (int64_t)Tcap + t);

    uint2* outp = LLp + (((size_t)slot * gm.ngroups + grp) * nsl + slice) * MR * 8;
    // partial sums of one M-tile -> LL blocks: row m, 8 words of (two samples as bf16x2, tag)
    auto emit_tile = [&](int mt) {
      uint32_t v[8], v2[8];
      tc_ld8_nowait(tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)(mt * NB), v);
      tc_ld8_nowait(tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)(mt * NB + 8), v2);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      const int m = mt * 128 + q * 32 + lane;
      if (m < MR) {
        uint2* o = outp + (size_t)m * 8;
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          ll_store4(o + 2 * e, pack_bf16(__uint_as_float(v[4 * e]), __uint_as_float(v[4 * e + 1])),
                    pack_bf16(__uint_as_float(v[4 * e + 2]), __uint_as_float(v[4 * e + 3])), tag);
          ll_store4(o + 4 + 2 * e, pack_bf16(__uint_as_float(v2[4 * e]), __uint_as_float(v2[4 * e + 1])),
                    pack_bf16(__uint_as_float(v2[4 * e + 2]), __uint_as_float(v2[4 * e + 3])), tag);
        }
      }
    };
    const int MTX = H / 128;                                   // first M-tile of the x rows
    if (s > 0) {
      if (lane == 0 && w < 4) {
        // the x rows first (they head the longer dependency chain: dx_pre -> dpx_t -> exchange -> [wp ; ws]^T dpx_t), their
        // own barrier; then the unit rows.  One accumulator per M-tile, each thread interleaves its tiles k-step by k-step.
        tc_fence_after();
        const uint64_t db0 = umma_desc_nosw(bop_local, NB * 16, 128);
        auto issue_tile_step = [&](int mt, int kk) {
          if (mt < BT_TMEM) tc_mma_ts(tmem_d + mt * NB, tmem + (uint32_t)(mt * 64 + kk * 8), db0 + (uint64_t)(kk * 2 * NB), idesc, kk ? 1u : 0u);
          else tc_mma(tmem_d + mt * NB, umma_desc_nosw(as_local + (uint32_t)(mt - BT_TMEM) * 32768 + kk * 4096, 2048, 128),
                      db0 + (uint64_t)(kk * 2 * NB), idesc, kk ? 1u : 0u);
        };
        for (int kk = 0; kk < 8; ++kk)
          for (int mt = MTX + w; mt < NMT; mt += 4) issue_tile_step(mt, kk);
        tc_commit(mma_x);
      }
      mbar_wait(mma_x, nmma & 1u);
      tc_fence_after();
      for (int mt = MTX + hf; mt < NMT; mt += 2) emit_tile(mt);
      // the unit rows run on the tensor pipe while the x-row owners wait for their blocks (tcgen05.ld of the x tiles
      // would otherwise queue behind these MMAs)
      // issued by warps 4-7: the issuing thread blocks while the pipe's queue is full, and warps 0-1 hold the x-row reduce
      if (lane == 0 && w >= 4) {
        tc_fence_after();
        const uint64_t db0 = umma_desc_nosw(bop_local, NB * 16, 128);
        for (int kk = 0; kk < 8; ++kk) {
          for (int mt = w - 4; mt < MTX; mt += 4) {
            if (mt < BT_TMEM) tc_mma_ts(tmem_d + mt * NB, tmem + (uint32_t)(mt * 64 + kk * 8), db0 + (uint64_t)(kk * 2 * NB), idesc, kk ? 1u : 0u);
            else tc_mma(tmem_d + mt * NB, umma_desc_nosw(as_local + (uint32_t)(mt - BT_TMEM) * 32768 + kk * 4096, 2048, 128),
                        db0 + (uint64_t)(kk * 2 * NB), idesc, kk ? 1u : 0u);
          }
        }
        tc_commit(mma_done);
      }
      ck.lap(0);
    }
    // ---- owners of the x rows: dx_pre = sum of the nsl incoming blocks, dpx_t, publish
    const uint2* inb = LLp + ((size_t)slot * gm.ngroups + grp) * nsl * MR * 8;
    if (tid < np * 8) {
      float s0 = 0.f, s1 = 0.f;
      if (s > 0) {
        const uint2* p = inb + (size_t)(H + p0 + rr) * 8 + sp;
        uint2 v[32];
        unsigned pending = nsl >= 32 ? 0xffffffffu : ((1u << nsl) - 1u);
        while (pending) {
#pragma unroll
          for (int i = 0; i < 32; ++i) if (pending & (1u << i)) v[i] = ll_load2(p + (size_t)i * MR * 8);
#pragma unroll
          for (int i = 0; i < 32; ++i) if ((pending & (1u << i)) && v[i].y == tag) pending &= ~(1u << i);
        }
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (i < nsl) { s0 += __uint_as_float(v[i].x << 16); s1 += __uint_as_float(v[i].x & 0xffff0000u); }
      }
      const float dv[2] = {(dxe[0] + s0) * (1.f - xo[0] * xo[0]), (dxe[1] + s1) * (1.f - xo[1] * xo[1])};
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int b = b0 + 2 * sp + e;
        ll_store2(LLd + ((size_t)slot * Bpad + b) * F + p0 + rr, __float_as_uint(dv[e]), tag);
        if (b < B) {
          d.dpx[(b * (int64_t)Tcap + t) * FP + p0 + rr] = dv[e];
          if (dp16) dp16[(b * (int64_t)Tcap + t) * FP + p0 + rr] = __float2bfloat16(dv[e]);
        }
      }
    }
    if (owns_last && tid < NB && b0 + tid < B) {
      float* qd = d.dpx + ((b0 + tid) * (int64_t)Tcap + t) * FP;
      qd[F] = dsv;
      for (int p = F + 1; p < FP; ++p) qd[p] = 0.f;
      if (dp16) {
        __nv_bfloat16* q16 = dp16 + ((b0 + tid) * (int64_t)Tcap + t) * FP;
        q16[F] = __float2bfloat16(dsv);
        for (int p = F + 1; p < FP; ++p) q16[p] = __float2bfloat16(0.f);
      }
    }
    if (tid < NB) dpxs[tid * WLD + F] = __float2bfloat16(dsv);
    ck.lap(1);
    if (s > 0) {
      mbar_wait(mma_done, nmma & 1u);
      ++nmma;
      tc_fence_after();
      for (int mt = hf; mt < MTX; mt += 2) emit_tile(mt);
      tc_fence_before();
    }
    ck.lap(2);
    // ---- every CTA: dh_pre of its 32 units
    {
      float s0 = 0.f, s1 = 0.f;
      if (s > 0) {
        const uint2* p = inb + (size_t)(j0 + rr) * 8 + sp;
        uint2 v[32];
        unsigned pending = nsl >= 32 ? 0xffffffffu : ((1u << nsl) - 1u);
        while (pending) {
#pragma unroll
          for (int i = 0; i < 32; ++i) if (pending & (1u << i)) v[i] = ll_load2(p + (size_t)i * MR * 8);
#pragma unroll
          for (int i = 0; i < 32; ++i) if ((pending & (1u << i)) && v[i].y == tag) pending &= ~(1u << i);
        }
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (i < nsl) { s0 += __uint_as_float(v[i].x << 16); s1 += __uint_as_float(v[i].x & 0xffff0000u); }
      }
      dhs[(2 * sp) * 33 + rr] = s0;
      dhs[(2 * sp + 1) * 33 + rr] = s1;
    }
    ck.lap(3);
    // ---- dpx_t [NB, F] of the batch group -> shared memory (bf16)
    {
      const uint2* src = LLd + ((size_t)slot * Bpad + b0) * F;
      constexpr int XC = 2;
      for (int base = 0; base < NB * (F / 8); base += LT * XC) {
        uint4 v[XC][4];
        unsigned pending = 0;
#pragma unroll
        for (int i = 0; i < XC; ++i)
          if (base + i * LT + tid < NB * (F / 8)) pending |= 1u << i;
        while (pending) {
#pragma unroll
          for (int i = 0; i < XC; ++i) {
            if (pending & (1u << i)) {
              const int idx = base + i * LT + tid, b = idx / (F / 8), c = idx - b * (F / 8);
              const uint2* p = src + (size_t)b * F + c * 8;
              v[i][0] = ll_load4(p); v[i][1] = ll_load4(p + 2); v[i][2] = ll_load4(p + 4); v[i][3] = ll_load4(p + 6);
            }
          }
#pragma unroll
          for (int i = 0; i < XC; ++i) {
            if (pending & (1u << i)) {
              bool ok = true;
#pragma unroll
              for (int e = 0; e < 4; ++e) ok = ok && v[i][e].y == tag && v[i][e].w == tag;
              if (ok) {
                const int idx = base + i * LT + tid, b = idx / (F / 8), c = idx - b * (F / 8);
                *reinterpret_cast<uint4*>(dpxs + (size_t)b * WLD + c * 8) = make_uint4(
                    pack_bf16(__uint_as_float(v[i][0].x), __uint_as_float(v[i][0].z)), pack_bf16(__uint_as_float(v[i][1].x), __uint_as_float(v[i][1].z)),
                    pack_bf16(__uint_as_float(v[i][2].x), __uint_as_float(v[i][2].z)), pack_bf16(__uint_as_float(v[i][3].x), __uint_as_float(v[i][3].z)));
                pending &= ~(1u << i);
              }
            }
          }
        }
      }
    }
    __syncthreads();
    // ---- dh += [wp ; ws]^T dpx_t for the CTA's own units: mma.sync m16n8k16, warp = (m-tile, n-tile, k half)
    {
      const int g8 = lane >> 2, tq = lane & 3, mtile = w & 1, ntile = (w >> 1) & 1, kh = w >> 2;
      const int nks2 = FP / 16, k_lo = kh * ((nks2 + 1) / 2), k_hi = min(nks2, k_lo + (nks2 + 1) / 2);
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
      const __nv_bfloat16* ar = WpT + (size_t)(mtile * 16 + g8) * WLD + 2 * tq;
      const __nv_bfloat16* br = dpxs + (size_t)(ntile * 8 + g8) * WLD + 2 * tq;
      for (int ks = k_lo; ks < k_hi; ++ks) {
        uint32_t a[4];
        a[0] = *reinterpret_cast<const uint32_t*>(ar + ks * 16);
        a[1] = *reinterpret_cast<const uint32_t*>(ar + 8 * WLD + ks * 16);
        a[2] = *reinterpret_cast<const uint32_t*>(ar + ks * 16 + 8);
        a[3] = *reinterpret_cast<const uint32_t*>(ar + 8 * WLD + ks * 16 + 8);
        mma_bf16_16816(acc, a, *reinterpret_cast<const uint32_t*>(br + ks * 16), *reinterpret_cast<const uint32_t*>(br + ks * 16 + 8));
      }
      const int u0 = mtile * 16 + g8, c0 = ntile * 8 + 2 * tq;
      atomicAdd(&dhs[c0 * 33 + u0], acc[0]);
      atomicAdd(&dhs[(c0 + 1) * 33 + u0], acc[1]);
      atomicAdd(&dhs[c0 * 33 + u0 + 8], acc[2]);
      atomicAdd(&dhs[(c0 + 1) * 33 + u0 + 8], acc[3]);
    }
    __syncthreads();
    ck.lap(4);
    // ---- cell backward for (sample bl, units jv, jv + 1)
    {
      const float gi[2] = {pg[0].x, pg[0].y}, gf[2] = {pg[1].x, pg[1].y}, gg[2] = {pg[2].x, pg[2].y}, go[2] = {pg[3].x, pg[3].y};
      const float cc[2] = {pc.x, pc.y}, cp[2] = {pcp.x, pcp.y};
      float o[4][2];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const float dht = dhs[bl * 33 + jv + e];
        const float tch = tanh_approx(cc[e]);                  // the function the forward applied
        const float dc = dcs[e] + dht * go[e] * (1.f - tch * tch);
        dcs[e] = dc * gf[e];
        o[0][e] = dc * gg[e] * gi[e] * (1.f - gi[e]);
        o[1][e] = dc * cp[e] * gf[e] * (1.f - gf[e]);
        o[2][e] = dc * gi[e] * (1.f - gg[e] * gg[e]);
        o[3][e] = dht * tch * go[e] * (1.f - go[e]);
      }
#pragma unroll
      for (int qq = 0; qq < 4; ++qq) {
        const uint32_t pk = pack_bf16(o[qq][0], o[qq][1]);
        const int lr = qq * 32 + jv;
        *reinterpret_cast<uint32_t*>(Bop + (lr >> 3) * (NB * 16) + bl * 16 + (lr & 7) * 2) = pk;
        if (bme < B) {
          const int64_t off = bme * gstr + (int64_t)t * 4 * H + qq * H + j0 + jv;
          if (want_f32) *reinterpret_cast<float2*>(d.dgates + off) = make_float2(o[qq][0], o[qq][1]);
          if (dg16) *reinterpret_cast<uint32_t*>(dg16 + off) = pk;
        }
      }
    }
    fence_proxy_async();
    __syncthreads();
    ck.lap(5);
  }
  if (d.dbg && tid == 0) {
    long long* qd = d.dbg + (int64_t)blockIdx.x * 8;
    for (int i = 0; i < 7; ++i) qd[i] = ck.acc[i];
    qd[7] = clock64() - tstart;
  }
  tc_fence_before();
  __syncthreads();
  if (w == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}

// ---------------------------------------------------------------------------------- host side
static size_t fwd_smem(const ag_lstm_desc* d, const FwdGeom& g, int NB) {
  return 1024 + (size_t)((g.KP - g.KT) / 8) * 2048 + (size_t)(g.KP / 8) * (NB * 16 + 16) + (size_t)16 * (d->H + 8) * 2 +
         (size_t)4 * NB * UPC * 4 + 2 * NB * 4 + 64;
}

// geometry of the forward launch for (B, H, F); returns a reason string when the shape cannot run here, else NULL
static const char* fwd_geom(const ag_lstm_desc* d, FwdGeom& g, int& NB, size_t& ll_need, size_t& smem, char* why, size_t nwhy) {
  if (d->H % 128 != 0 || d->F % 8 != 0) { snprintf(why, nwhy, "H=%d %% 128 or F=%d %% 8 != 0", d->H, d->F); return why; }
  g.nsl = d->H / UPC;
  NB = (d->B + 15) / 16 * g.nsl <= sm_count() ? 16 : 32;
  g.ngroups = (d->B + NB - 1) / NB;
  if (g.ngroups * g.nsl > sm_count() || g.ngroups > 16) {
    snprintf(why, nwhy, "B=%d: %d batch groups x %d slices > %d SMs", d->B, g.ngroups, g.nsl, sm_count());
    return why;
  }
  g.KP = (d->H + d->F + 15) / 16 * 16;
  g.KT = std::min(d->H / 64 * 64, 896);             // 448 TMEM columns of weights + 64 of accumulators; k < KT <= H
  g.PR = (d->F + 1 + g.nsl - 1) / g.nsl;
  if (g.PR > 8) { snprintf(why, nwhy, "F=%d: %d projection rows per CTA > 8", d->F, g.PR); return why; }
  ll_need = ((size_t)2 * g.ngroups * NB * (d->H / 2) + (size_t)2 * g.ngroups * NB * d->F + 32) * 8 + 16384;   // + exchange flags
  {
    static int mode = -1;                  // AUDIOGAN_GEN_XCHG=ll: the tagged-word exchange (A/B knob); default: flags + TMA bulk copies
    if (mode < 0) { const char* e = getenv("AUDIOGAN_GEN_XCHG"); mode = (e && e[0] == 'l') ? 0 : 1; }
    g.xchg = mode;
  }
  smem = fwd_smem(d, g, NB);
  if (smem > (size_t)smem_optin()) {
    snprintf(why, nwhy, "H=%d F=%d: weight slice needs %zu B of shared memory beside tensor memory (> %d)", d->H, d->F, smem, smem_optin());
    return why;
  }
  if (smem < (size_t)116 * 1024) { snprintf(why, nwhy, "H=%d: slice too small to pin one CTA per SM", d->H); return why; }
  return nullptr;
}

// *launched = 1 when this kernel took the call, 0 -> the caller falls back to lstm.cu
int gen_fwd(const ag_lstm_desc* d, cudaStream_t s, int* launched) {
  *launched = 0;
  if (d->F <= 0 || d->ndir != 1 || d->prec < 1 || !(d->flags & 2) || (d->flags & 1)) return AG_OK;
  FwdGeom g;
  int NB = 16;
  size_t ll_need = 0, smem = 0;
  char why[160];
  if (fwd_geom(d, g, NB, ll_need, smem, why, sizeof(why))) { set_decline("tmem declined: %s", why); return AG_OK; }
  if (!d->hbuf16 || !d->xbuf16 || !d->ll_ws) { set_decline("tmem declined: hbuf16 / xbuf16 / ll_ws missing"); return AG_OK; }
  if ((size_t)d->ll_ws_bytes < ll_need) { set_decline("tmem declined: ll_ws %lld B < %zu B", (long long)d->ll_ws_bytes, ll_need); return AG_OK; }
  const void* kern = NB == 16 ? (const void*)lstm_gen_fwd_kernel<16> : (const void*)lstm_gen_fwd_kernel<32>;
  AG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  AG_CUDA(cudaMemsetAsync(d->ll_ws, 0, ll_need, s));                  // tags: step numbers start at 1
  AG_CUDA(cudaMemsetAsync(d->t_end, 0, sizeof(int), s));
  ag_lstm_desc dd = *d;
  void* args[2] = {&dd, &g};
  AG_CUDA(cudaLaunchCooperativeKernel(kern, dim3((unsigned)(g.ngroups * g.nsl)), dim3(LT), args, smem, s));
  *launched = 1;
  set_path("tmem");
  return AG_OK;
}
static const char* bwd_geom(const ag_lstm_desc* d, BwdGeom& g, size_t& ll_need, size_t& smem, char* why, size_t nwhy);
// largest batch one launch of the TMEM-resident kernels takes for this (H, F): whole batch groups that fit the SMs
int gen_batch_cap(const ag_lstm_desc* d, int bwd) {
  if (d->F <= 0 || d->ndir != 1 || d->prec < 1 || d->H % 128 != 0) return 0;
  const int nsl = d->H / UPC;
  const int groups = std::min(sm_count() / std::max(nsl, 1), 16);
  if (groups <= 0) return 0;
  ag_lstm_desc t = *d;
  t.B = groups * (bwd ? BNB : 32);
  char why[160];
  size_t ll = 0, smem = 0;
  if (bwd) {
    BwdGeom g;
    if (bwd_geom(&t, g, ll, smem, why, sizeof(why))) return 0;
  } else {
    FwdGeom g;
    int NB = 16;
    if (fwd_geom(&t, g, NB, ll, smem, why, sizeof(why))) return 0;
  }
  return t.B;
}
int64_t gen_fwd_ws_bytes(const ag_lstm_desc* d) {
  if (d->F <= 0 || d->ndir != 1 || d->prec < 1) return 0;
  FwdGeom g;
  int NB = 16;
  size_t ll_need = 0, smem = 0;
  char why[160];
  if (fwd_geom(d, g, NB, ll_need, smem, why, sizeof(why))) return 0;
  return (int64_t)ll_need;
}


static size_t bwd_smem(const ag_lstm_desc* d, const BwdGeom& g) {
  const int NMT = (g.MR + 127) / 128, WLD = g.FP + 8;
  return 1024 + (size_t)(NMT - BT_TMEM) * 32768 + 16 * BNB * 16 + (size_t)(UPC + BNB) * WLD * 2 + (size_t)(BNB * 33) * 4 + 64;
}

static const char* bwd_geom(const ag_lstm_desc* d, BwdGeom& g, size_t& ll_need, size_t& smem, char* why, size_t nwhy) {
  if (d->H % 128 != 0 || d->F % 8 != 0) { snprintf(why, nwhy, "H=%d %% 128 or F=%d %% 8 != 0", d->H, d->F); return why; }
  g.nsl = d->H / UPC;
  g.ngroups = (d->B + BNB - 1) / BNB;
  if (g.nsl > 32 || g.ngroups * g.nsl > sm_count()) {
    snprintf(why, nwhy, "B=%d H=%d: %d batch groups x %d slices > %d SMs (or > 32 slices)", d->B, d->H, g.ngroups, g.nsl, sm_count());
    return why;
  }
  g.MR = d->H + d->F;
  g.FP = (d->F + 1 + 7) / 8 * 8;
  g.PR = (d->F + g.nsl - 1) / g.nsl;
  const int NMT = (g.MR + 127) / 128;
  if (g.PR * 8 > LT || NMT <= BT_TMEM || BT_TMEM * 64 + NMT * BNB > 512 || g.FP % 16 != 0) {
    snprintf(why, nwhy, "H=%d F=%d: %d M-tiles do not fit tensor memory / FP=%d %% 16", d->H, d->F, NMT, g.FP);
    return why;
  }
  ll_need = ((size_t)2 * g.ngroups * g.nsl * g.MR * 8 + (size_t)2 * g.ngroups * BNB * d->F + 32) * 8;
  smem = bwd_smem(d, g);
  if (smem > (size_t)smem_optin()) {
    snprintf(why, nwhy, "H=%d F=%d: needs %zu B of shared memory (> %d)", d->H, d->F, smem, smem_optin());
    return why;
  }
  if (smem < (size_t)116 * 1024) { snprintf(why, nwhy, "H=%d: slice too small to pin one CTA per SM", d->H); return why; }
  return nullptr;
}

int gen_bwd(const ag_lstm_desc* d, cudaStream_t s, int* launched) {
  *launched = 0;
  if (d->F <= 0 || d->ndir != 1 || d->prec < 1 || !(d->flags & 2) || (d->flags & 1)) return AG_OK;
  BwdGeom g;
  size_t ll_need = 0, smem = 0;
  char why[160];
  if (bwd_geom(d, g, ll_need, smem, why, sizeof(why))) { set_decline("tmem declined: %s", why); return AG_OK; }
  if (!d->ll_ws || d->len || d->dh_ext) { set_decline("tmem declined: ll_ws missing or len / dh_ext given"); return AG_OK; }
  if ((size_t)d->ll_ws_bytes < ll_need) { set_decline("tmem declined: ll_ws %lld B < %zu B", (long long)d->ll_ws_bytes, ll_need); return AG_OK; }
  const void* kern = (const void*)lstm_gen_bwd_kernel;
  AG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  AG_CUDA(cudaMemsetAsync(d->ll_ws, 0, ll_need, s));
  ag_lstm_desc dd = *d;
  void* args[2] = {&dd, &g};
  AG_CUDA(cudaLaunchCooperativeKernel(kern, dim3((unsigned)(g.ngroups * g.nsl)), dim3(LT), args, smem, s));
  *launched = 1;
  set_path("tmem");
  return AG_OK;
}
int64_t gen_bwd_ws_bytes(const ag_lstm_desc* d) {
  if (d->F <= 0 || d->ndir != 1 || d->prec < 1) return 0;
  BwdGeom g;
  size_t ll_need = 0, smem = 0;
  char why[160];
  if (bwd_geom(d, g, ll_need, smem, why, sizeof(why))) return 0;
  return (int64_t)ll_need;
}

}  // namespace lg
}  // namespace ag
