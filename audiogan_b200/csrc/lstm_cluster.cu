// Cluster-resident recurrent kernels (bf16 mode) for the discriminator's bidirectional LSTM
// (audiogan.py:214-229, :498-503, :543 -- NN.LSTM under dynamic_rnn) and its BPTT.
//
// The grid-barrier kernels in lstm.cu pay ~3.2 k cycles per step for the grid barrier plus an L2 round trip for the
// [B, H] state vector (profiles/r1_lstm_phase_cycles.txt).  Samples are independent, so nothing forces one copy of
// the weights per GPU: here ONE THREAD-BLOCK CLUSTER of CS = H/32 CTAs (16 for the default H = 512) holds a full copy
// of one direction's recurrent weights in its distributed shared memory (128 gate rows x H bf16 = 128 KB per CTA) and
// runs a slice of <= 32 samples through all T steps on its own.  Per step and CTA:
//   forward   D[128 gate rows (TMEM lanes), NB samples] = Whh_slice[128, H] . h_{t-1}[NB, H]^T      (H/16 tcgen05.mma, M128 N16/32)
//             tcgen05.ld -> + input projection -> gate non-linearities -> smem exchange -> cell update (c in registers)
//             -> h_t slice [NB, 32] bf16 pushed into every peer's next-step B operand through DSMEM (st.shared::cluster)
//             -> barrier.cluster (release / acquire).  No grid barrier, no L2 round trip on the critical path.
//   backward  partial dh[H units (4 M-tiles), NB] = Whh_slice^T[H, 128] . dgates_{t+1, slice}[NB, 128]^T: the B operand is
//             the CTA's OWN gate gradients (no all-gather); the partial sums are reduce-scattered through DSMEM (fp32),
//             each owner adds its 16 incoming blocks, runs the cell backward (dc in registers) and refills its B operand.
// Clusters never talk to each other: cluster c serves direction c % ndir and every (nclusters / ndir)-th sample slice.
#include "common.cuh"
#include "tc_common.cuh"
#include <algorithm>

namespace ag {
namespace lc {

using namespace tc;

constexpr int LT = 256;     // threads per CTA
constexpr int UPC = 32;     // hidden units per CTA -> 4 * 32 = 128 gate rows = the MMA's M

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t cluster_nctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ uint32_t mapa(uint32_t saddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_cluster_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void tc_ld8(uint32_t taddr, uint32_t (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tc_ld16_nowait(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
}
template <int CW>
__device__ __forceinline__ void tc_ldw(uint32_t taddr, uint32_t (&v)[CW]);
template <>
__device__ __forceinline__ void tc_ldw<8>(uint32_t taddr, uint32_t (&v)[8]) { tc_ld8(taddr, v); }
template <>
__device__ __forceinline__ void tc_ldw<16>(uint32_t taddr, uint32_t (&v)[16]) { tc_ld16(taddr, v); }

// 2 MUFU each; absolute error ~1e-7, far inside the bf16 mode's tolerance
__device__ __forceinline__ float fsig(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
__device__ __forceinline__ float ftanh(float x) { return 2.f * fsig(2.f * x) - 1.f; }
// MUFU.TANH: abs error ~5e-4 (2^-11), below the bf16 rounding of h that the recurrent product sees anyway
__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <uint32_t COLS>
__device__ __forceinline__ uint32_t tmem_alloc(uint32_t* slot) {
  if ((threadIdx.x >> 5) == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  return *slot;
}
template <uint32_t COLS>
__device__ __forceinline__ void tmem_free(uint32_t tmem) {
  tc_fence_before();
  __syncthreads();
  if ((threadIdx.x >> 5) == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(COLS) : "memory");
  }
}

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

// ======================================================================================= forward
// CTA = G warp groups of 128 threads; group g carries its own 16-sample sub-slice through the sequence, so one group's
// exchange latency is covered by the other groups' MMAs / cell updates (the groups share the TMEM-resident weights).
//   TMEM   columns [0, H/2): A = the CTA's 128 gate rows x H bf16 (lane = row, 2 k per column), written once with
//          tcgen05.st;  columns H/2 + 16 g ..: group g's accumulator D[128 rows, 16 samples].
//   smem   per group: Bt[2] = h_{t-1} [16 samples, H] bf16, K-major NO-swizzle core-matrix layout
//          [k-chunk of 8][16 rows][16 B] -- source CTA r owns k-chunks 4r..4r+3 = one contiguous 1 KB block, so the
//          exchange is ONE bulk DSMEM copy per peer (cp.async.bulk.shared::cluster.shared::cta) that completes on the
//          receiver's mbarrier (complete_tx): no cluster barrier inside the time loop.
//          gs [4][16][32] fp32 gate exchange, hs[2] [4][16][8] bf16 staging of the outgoing h slice, 3 mbarriers.
constexpr int NBG = 16;                       // samples per warp group = the MMA's N
constexpr int GT = 128;                       // threads per warp group
constexpr int MAXG = 4;

__device__ __forceinline__ uint64_t umma_desc_nosw(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= 1ull << 46;
  return d;                                   // layout type 0: SWIZZLE_NONE
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
__device__ __forceinline__ void group_sync(int g) { asm volatile("bar.sync %0, %1;" ::"r"(g + 1), "r"(GT) : "memory"); }
__device__ __forceinline__ void bulk_copy_to_peer(uint32_t dst_cluster, uint32_t src_cta, uint32_t bytes, uint32_t mbar_cluster) {
  asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_cluster),
               "r"(src_cta), "r"(bytes), "r"(mbar_cluster)
               : "memory");
}

__host__ __device__ inline uint32_t fwd_group_bytes(int H) {
  return 2u * NBG * H * 2 + 4u * NBG * UPC * 4 + 2u * NBG * UPC * 2 + 64;
}

__global__ void __launch_bounds__(GT * MAXG, 1) lstm_cl_fwd_kernel(const ag_lstm_desc d, const int nss, const int cpd, const int G) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const int tid = threadIdx.x, g = tid >> 7, gt = tid & (GT - 1), lane = tid & 31, q = (tid >> 5) & 3;
  const int H = d.H, ndir = d.ndir, B = d.B, T = d.T, Tcap = d.Tcap;
  const int CS = (int)cluster_nctarank(), rank = (int)cluster_ctarank();
  const int cid = blockIdx.x / CS, dir = cid % ndir, cl = cid / ndir, j0 = rank * UPC;

  uint8_t* sm = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm);
  uint8_t* gbase = sm + 64 + (size_t)g * fwd_group_bytes(H);
  const uint32_t bt_bytes = (uint32_t)NBG * H * 2;
  uint8_t* Bt = gbase;
  float* gs = reinterpret_cast<float*>(Bt + 2 * (size_t)bt_bytes);
  uint8_t* hs = reinterpret_cast<uint8_t*>(gs + 4 * NBG * UPC);
  uint64_t* full = reinterpret_cast<uint64_t*>(hs + 2 * NBG * UPC * 2);      // full[0], full[1], mma_done
  uint64_t* mma_done = full + 2;

  if (gt == 0) {
    mbar_init(&full[0], 1);
    mbar_init(&full[1], 1);
    mbar_init(mma_done, 4);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  const uint32_t tmem = tmem_alloc<512>(tmem_slot);
  // resident weights -> TMEM: this thread's row lr = 32 q + lane (gate q, unit j0 + lane), column range split over the groups
  {
    const float* src = d.w1 + ((int64_t)dir * 4 * H + (int64_t)q * H + j0 + lane) * H;
    // 4 column blocks per pass: all 16 loads in flight before the first tcgen05.st (latency-bound otherwise)
    for (int k0 = g * 16; k0 < H; k0 += 64 * G) {
      float4 a[4][4];
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int e = 0; e < 4; ++e)
          a[u][e] = (k0 + 16 * G * u < H) ? __ldg(reinterpret_cast<const float4*>(src + k0 + 16 * G * u) + e) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (k0 + 16 * G * u < H) {
          uint32_t v[8];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            v[2 * e] = pack_bf16(a[u][e].x, a[u][e].y);
            v[2 * e + 1] = pack_bf16(a[u][e].z, a[u][e].w);
          }
          tc_st8(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)((k0 + 16 * G * u) >> 1), v);
        }
      }
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_arrive();          // every CTA's barriers are initialised (and the weights in place) before any remote completion
  cluster_wait();
  tc_fence_after();

  const uint32_t tmem_d = tmem + (uint32_t)(H / 2 + 4 * NBG * g);   // 4 accumulators (one per K quarter)
  const uint32_t idesc = umma_idesc(128, NBG, 0, 0);
  const int64_t hstr = (int64_t)(Tcap + 2) * ndir * H, gstr = (int64_t)Tcap * ndir * 4 * H, cstr = (int64_t)Tcap * ndir * H;
  __nv_bfloat16* hb16 = reinterpret_cast<__nv_bfloat16*>(d.hbuf16);
  const bool want_h32 = !(d.flags & AG_LSTM_BF16_H_ONLY);       // bf16 mode: every consumer of h reads the bf16 copy
  const uint32_t bt_local = smem_u32(Bt), full_local = smem_u32(full), hs_local = smem_u32(hs);
  uint32_t nuse0 = 0, nuse1 = 0, nmma = 0;     // completed phases of full[0], full[1], mma_done
  Clk ck;
  ck.init(d.dbg != nullptr);
  const long long tstart = clock64();

  for (int ss0 = 0; ss0 < nss; ss0 += cpd * G) {
    const int ss = ss0 + g * cpd + cl;          // this group's sub-slice of direction `dir`
    if (ss < nss) {
      const int b0 = ss * NBG;
      for (uint32_t i = gt * 16; i < bt_bytes; i += GT * 16) *reinterpret_cast<uint4*>(Bt + i) = make_uint4(0u, 0u, 0u, 0u);
      // cell-update items of this thread: sample bl = gt / 8, units j4 .. j4 + 3 (float4 loads / stores everywhere)
      const int bl = gt >> 3, j4 = (gt & 7) * 4, bme = b0 + bl;
      float cst[4] = {0.f, 0.f, 0.f, 0.f};
      const int len_me = bme < B ? (d.len ? d.len[bme] : T) : 0;
      fence_proxy_async();
      group_sync(g);
      if (gt == 0 && T > 1) mbar_arrive_expect_tx(&full[1], bt_bytes);

      float4 pre[4];
      auto load_pre = [&](int t) {
        const float* pp = d.pre + bme * gstr + (int64_t)t * ndir * 4 * H + dir * 4 * H + j0 + j4;
#pragma unroll
        for (int qq = 0; qq < 4; ++qq)
          pre[qq] = bme < B ? __ldg(reinterpret_cast<const float4*>(pp + qq * H)) : make_float4(0.f, 0.f, 0.f, 0.f);
      };
      load_pre(dir ? T - 1 : 0);

      for (int s = 0; s < T; ++s) {
        ck.start();
        const int t = dir ? (T - 1 - s) : s;
        const uint32_t buf = (uint32_t)s & 1u;
        if (s > 0) {
          if (buf) { mbar_wait(&full[1], nuse1 & 1u); ++nuse1; }
          else { mbar_wait(&full[0], nuse0 & 1u); ++nuse0; }
        }
        ck.lap(0);
        if (gt == 0 && s + 2 < T) mbar_arrive_expect_tx(&full[buf], bt_bytes);     // h_{s+1} lands here for step s + 2
        if (lane == 0) {
          // 4 issuing threads (one per warp), each with its own accumulator and a quarter of K: back-to-back MMAs into
          // ONE accumulator serialise on the TMEM read-modify-write (~63 cycles each, measured), independent chains do not
          tc_fence_after();
          const int nk = H / 64;                                          // MMAs per chain
          uint64_t db = umma_desc_nosw(bt_local + buf * bt_bytes, NBG * 16, 128) + (uint64_t)(q * nk * 2 * NBG);
          uint32_t ta = tmem + (uint32_t)(q * nk * 8);
          for (int kk = 0; kk < nk; ++kk) {
            tc_mma_ts(tmem_d + NBG * q, ta, db, idesc, kk ? 1u : 0u);
            ta += 8;                 // 16 k = 8 packed columns
            db += 2 * NBG;           // 2 k-chunks x (16 rows x 16 B) = 512 B, in 16-byte units
          }
          tc_commit(mma_done);
        }
        mbar_wait(mma_done, nmma & 1u);
        ++nmma;
        tc_fence_after();
        ck.lap(1);
        {
          uint32_t v[NBG], v1[NBG], v2[NBG], v3[NBG];
          const uint32_t ta = tmem_d + ((uint32_t)(q * 32) << 16);
          tc_ld16_nowait(ta, v);
          tc_ld16_nowait(ta + NBG, v1);
          tc_ld16_nowait(ta + 2 * NBG, v2);
          tc_ld16_nowait(ta + 3 * NBG, v3);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          ck.lap(2);
#pragma unroll
          for (int i = 0; i < NBG; ++i)
            gs[(q * NBG + i) * UPC + lane] = (__uint_as_float(v[i]) + __uint_as_float(v1[i])) + (__uint_as_float(v2[i]) + __uint_as_float(v3[i]));
        }
        tc_fence_before();
        group_sync(g);
        ck.lap(3);
        uint8_t* hsb = hs + buf * (NBG * UPC * 2);
        float4 gq[4], cv = make_float4(0.f, 0.f, 0.f, 0.f), hv = cv;
#pragma unroll
        for (int qq = 0; qq < 4; ++qq) gq[qq] = cv;
        if (t < len_me) {
          // sigmoid(x) = 0.5 tanh(0.5 x) + 0.5: one MUFU per gate value
#pragma unroll
          for (int qq = 0; qq < 4; ++qq) {
            const float4 a = *reinterpret_cast<const float4*>(gs + (qq * NBG + bl) * UPC + j4);
            const float sc = (qq == 2) ? 1.f : 0.5f, of = (qq == 2) ? 0.f : 0.5f;
            gq[qq].x = fmaf(tanh_approx((a.x + pre[qq].x) * sc), sc, of);
            gq[qq].y = fmaf(tanh_approx((a.y + pre[qq].y) * sc), sc, of);
            gq[qq].z = fmaf(tanh_approx((a.z + pre[qq].z) * sc), sc, of);
            gq[qq].w = fmaf(tanh_approx((a.w + pre[qq].w) * sc), sc, of);
          }
          cv.x = gq[1].x * cst[0] + gq[0].x * gq[2].x;
          cv.y = gq[1].y * cst[1] + gq[0].y * gq[2].y;
          cv.z = gq[1].z * cst[2] + gq[0].z * gq[2].z;
          cv.w = gq[1].w * cst[3] + gq[0].w * gq[2].w;
          cst[0] = cv.x; cst[1] = cv.y; cst[2] = cv.z; cst[3] = cv.w;
          hv.x = gq[3].x * tanh_approx(cv.x);
          hv.y = gq[3].y * tanh_approx(cv.y);
          hv.z = gq[3].z * tanh_approx(cv.z);
          hv.w = gq[3].w * tanh_approx(cv.w);
        }
        const uint2 h16 = make_uint2(pack_bf16(hv.x, hv.y), pack_bf16(hv.z, hv.w));
        *reinterpret_cast<uint2*>(hsb + (j4 >> 3) * (NBG * 16) + bl * 16 + (j4 & 7) * 2) = h16;
        fence_proxy_async();         // only shared-memory stores are outstanding here: the global stores come after the push
        group_sync(g);
        ck.lap(4);
        if (s + 1 < T && lane < 4) {
          // this CTA's [16, 32] slice of h_t -> k-chunks 4 rank .. 4 rank + 3 of every peer's next-step B operand;
          // the issue of one bulk copy costs ~60 cycles per lane (measured), so the 16 copies are spread over the 4 warps
          const int peer = lane * 4 + q;
          if (peer < CS) {
            const uint32_t dst = mapa(bt_local + (buf ^ 1u) * bt_bytes + (uint32_t)rank * (NBG * UPC * 2), (uint32_t)peer);
            const uint32_t bar = mapa(full_local + (buf ^ 1u) * 8, (uint32_t)peer);
            bulk_copy_to_peer(dst, hs_local + buf * (NBG * UPC * 2), NBG * UPC * 2, bar);
          }
        }
        ck.lap(5);
        // off the critical path: next step's input projections, then this step's saved state
        if (s + 1 < T) load_pre(dir ? (T - 2 - s) : (s + 1));
        if (bme < B) {
          const int64_t ho = bme * hstr + (int64_t)(t + 1) * ndir * H + dir * H + j0 + j4;
          if (want_h32) *reinterpret_cast<float4*>(d.hbuf + ho) = hv;
          *reinterpret_cast<uint2*>(hb16 + ho) = h16;
          if (d.cbuf) *reinterpret_cast<float4*>(d.cbuf + bme * cstr + (int64_t)t * ndir * H + dir * H + j0 + j4) = cv;
          if (d.gates) {
            float* gp = d.gates + bme * gstr + (int64_t)t * ndir * 4 * H + dir * 4 * H + j0 + j4;
#pragma unroll
            for (int qq = 0; qq < 4; ++qq) *reinterpret_cast<float4*>(gp + qq * H) = gq[qq];
          }
        }
        ck.lap(6);
      }
    }
    if (ss0 + cpd * G < nss) {   // another round: nobody may still be reading what the next round overwrites
      __syncthreads();
      cluster_arrive();
      cluster_wait();
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_arrive();            // no CTA leaves while a peer may still copy into its shared memory
  cluster_wait();
  if (d.dbg && gt == 0) {
    long long* qd = d.dbg + ((int64_t)blockIdx.x * MAXG + g) * 8;
    for (int i = 0; i < 7; ++i) qd[i] = ck.acc[i];
    qd[7] = clock64() - tstart;
  }
  tmem_free<512>(tmem);
}

// ====================================================================================== backward
// Same decomposition (CTA = 32 hidden units of one direction, G warp groups x 16 samples).  Per step and group:
//   partial dh[H units, 16] = Whh_slice^T [H, 128] . dgates_{prev step, slice}[16, 128]^T  -- A = the transposed weight slice,
//   H/128 M-tiles x 64 TMEM columns, resident; B = the CTA's OWN gate gradients of the step it just finished (no
//   all-gather).  The M-tiles are independent accumulator chains, issued by one thread per warp.
//   Reduce-scatter: warp q holds units 128 m + 32 q .. of every M-tile m = exactly owner CTA 4m + q's units; the partial
//   sums go (bf16, sample pairs packed) into a staging block per owner and leave as ONE 1 KB bulk DSMEM copy per owner,
//   completing on the owner's mbarrier.  Each owner adds its CS incoming blocks, runs the cell backward (dc in
//   registers) and refills its B operand.
//   smem per group: Bop [16 k-chunks][16 rows][16 B] | stage[2] [CS owners][8 sample pairs][32 units] bf16x2 |
//                   red[2] [CS sources][8][32] bf16x2 | full[2], mma_done.
constexpr int MAXG_BWD = 3;
__host__ __device__ inline uint32_t bwd_group_bytes(int H) {
  const uint32_t CS = (uint32_t)H / UPC;
  return 16u * NBG * 16 + 4u * CS * (NBG / 2) * UPC * 4 + 64;
}

__global__ void __launch_bounds__(GT * MAXG_BWD, 1) lstm_cl_bwd_kernel(const ag_lstm_desc d, const int nss, const int cpd, const int G) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  const int tid = threadIdx.x, g = tid >> 7, gt = tid & (GT - 1), lane = tid & 31, q = (tid >> 5) & 3;
  const int H = d.H, ndir = d.ndir, B = d.B, T = d.T, Tcap = d.Tcap;
  const int CS = (int)cluster_nctarank(), rank = (int)cluster_ctarank();
  const int cid = blockIdx.x / CS, dir = cid % ndir, cl = cid / ndir, j0 = rank * UPC;
  const int MT = H / 128;

  uint8_t* sm = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm);
  uint8_t* gbase = sm + 64 + (size_t)g * bwd_group_bytes(H);
  const uint32_t blk_bytes = (NBG / 2) * UPC * 4;            // one (owner | source) block: 8 sample pairs x 32 units x bf16x2 = 1 KB
  const uint32_t red_bytes = (uint32_t)CS * blk_bytes;
  uint8_t* Bop = gbase;
  uint8_t* stage = Bop + 16 * NBG * 16;
  uint8_t* red = stage + 2 * (size_t)red_bytes;
  uint64_t* full = reinterpret_cast<uint64_t*>(red + 2 * (size_t)red_bytes);
  uint64_t* mma_done = full + 2;

  if (gt == 0) {
    mbar_init(&full[0], 1);
    mbar_init(&full[1], 1);
    mbar_init(mma_done, (uint32_t)MT);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  const uint32_t tmem = tmem_alloc<512>(tmem_slot);
  // A[unit j][k = lr] = whh[gate row (lr/32)*H + j0 + lr%32][j] = w1t[dir][j][...]; this thread's rows: 128 m + 32 q + lane
  {
    const float* wt = d.w1t + (int64_t)dir * H * 4 * H;
    for (int idx0 = g; idx0 < MT * 8; idx0 += 4 * G) {
      float4 a[4][4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int idx = idx0 + u * G, m = idx >> 3, lr0 = (idx & 7) * 16;
        const float* src = wt + (int64_t)(m * 128 + q * 32 + lane) * 4 * H + (lr0 >> 5) * H + j0 + (lr0 & 31);
#pragma unroll
        for (int e = 0; e < 4; ++e)
          a[u][e] = idx < MT * 8 ? __ldg(reinterpret_cast<const float4*>(src) + e) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int idx = idx0 + u * G, m = idx >> 3, kk = idx & 7;
        if (idx < MT * 8) {
          uint32_t v[8];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            v[2 * e] = pack_bf16(a[u][e].x, a[u][e].y);
            v[2 * e + 1] = pack_bf16(a[u][e].z, a[u][e].w);
          }
          tc_st8(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(m * 64 + kk * 8), v);
        }
      }
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  cluster_arrive();
  cluster_wait();
  tc_fence_after();

  const uint32_t tmem_d = tmem + (uint32_t)(MT * 64 + MT * NBG * g);
  const uint32_t idesc = umma_idesc(128, NBG, 0, 0);
  const int64_t gstr = (int64_t)Tcap * ndir * 4 * H, cstr = (int64_t)Tcap * ndir * H;
  const int64_t dhbs = d.dh_ext_bs ? d.dh_ext_bs : cstr;
  __nv_bfloat16* dg16 = reinterpret_cast<__nv_bfloat16*>(d.dgates16);
  const bool want_f32 = !(dg16 && (d.flags & AG_LSTM_BF16_DGATES_ONLY));   // bf16 mode: every consumer reads the bf16 copy
  const uint32_t bop_local = smem_u32(Bop), stage_local = smem_u32(stage), red_local = smem_u32(red), full_local = smem_u32(full);
  uint32_t nuse0 = 0, nuse1 = 0, nmma = 0;
  Clk ck;
  ck.init(d.dbg != nullptr);
  const long long tstart = clock64();

  for (int ss0 = 0; ss0 < nss; ss0 += cpd * G) {
    const int ss = ss0 + g * cpd + cl;
    if (ss < nss) {
      const int b0 = ss * NBG;
      const int bl = gt >> 3, j4 = (gt & 7) * 4, bme = b0 + bl;      // cell-backward items: sample bl, units j4 .. j4 + 3
      const int len_me = bme < B ? (d.len ? min(d.len[bme], T) : T) : 0;
      float dcs[4] = {0.f, 0.f, 0.f, 0.f};
      if (gt == 0) {
        if (T > 1) mbar_arrive_expect_tx(&full[1], red_bytes);
        if (T > 2) mbar_arrive_expect_tx(&full[0], red_bytes);
      }

      for (int s = 0; s < T; ++s) {
        ck.start();
        const int t = dir ? s : (T - 1 - s);               // reverse of the forward order
        const uint32_t buf = (uint32_t)s & 1u;
        // what the cell backward needs (saved gates, c_t, c_prev, external dh): independent of the recurrence
        const bool valid = t < len_me;
        float4 pg[4], pc = make_float4(0.f, 0.f, 0.f, 0.f), pcp = pc, pdh = pc;
#pragma unroll
        for (int qq = 0; qq < 4; ++qq) pg[qq] = pc;
        if (valid) {
          const float* gp = d.gates + bme * gstr + (int64_t)t * ndir * 4 * H + dir * 4 * H + j0 + j4;
#pragma unroll
          for (int qq = 0; qq < 4; ++qq) pg[qq] = __ldg(reinterpret_cast<const float4*>(gp + qq * H));
          const float* cp = d.cbuf + bme * cstr + dir * H + j0 + j4;
          pc = __ldg(reinterpret_cast<const float4*>(cp + (int64_t)t * ndir * H));
          const int tp = dir ? (t + 1) : (t - 1);
          if (dir ? (tp < len_me) : (tp >= 0)) pcp = __ldg(reinterpret_cast<const float4*>(cp + (int64_t)tp * ndir * H));
          if (d.dh_ext) pdh = __ldg(reinterpret_cast<const float4*>(d.dh_ext + bme * dhbs + (int64_t)t * ndir * H + dir * H + j0 + j4));
        }
        float dh[4] = {0.f, 0.f, 0.f, 0.f};
        if (s > 0) {
          if (lane == 0 && q < MT) {
            tc_fence_after();
            uint64_t db = umma_desc_nosw(bop_local, NBG * 16, 128);
            uint32_t ta = tmem + (uint32_t)(q * 64);
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) {
              tc_mma_ts(tmem_d + NBG * q, ta, db, idesc, kk ? 1u : 0u);
              ta += 8;
              db += 2 * NBG;
            }
            tc_commit(mma_done);
          }
          mbar_wait(mma_done, nmma & 1u);
          ++nmma;
          tc_fence_after();
          ck.lap(0);
          // reduce-scatter, outgoing side: M-tile m, lanes 32 q .. = owner CTA 4 m + q's units
          uint8_t* stb = stage + buf * red_bytes;
          for (int m = 0; m < MT; ++m) {
            uint32_t v[NBG];
            tc_ld16(tmem_d + ((uint32_t)(q * 32) << 16) + (uint32_t)(m * NBG), v);
            uint32_t* dst = reinterpret_cast<uint32_t*>(stb + (uint32_t)(4 * m + q) * blk_bytes) + lane;
#pragma unroll
            for (int p = 0; p < NBG / 2; ++p) dst[p * UPC] = pack_bf16(__uint_as_float(v[2 * p]), __uint_as_float(v[2 * p + 1]));
          }
          tc_fence_before();
          fence_proxy_async();
          group_sync(g);
          ck.lap(1);
          if (lane < 4) {
            const int owner = lane * 4 + q;
            if (owner < CS) {
              const uint32_t dst = mapa(red_local + buf * red_bytes + (uint32_t)rank * blk_bytes, (uint32_t)owner);
              const uint32_t bar = mapa(full_local + buf * 8, (uint32_t)owner);
              bulk_copy_to_peer(dst, stage_local + buf * red_bytes + (uint32_t)owner * blk_bytes, blk_bytes, bar);
            }
          }
          ck.lap(2);
          if (buf) { mbar_wait(&full[1], nuse1 & 1u); ++nuse1; }
          else { mbar_wait(&full[0], nuse0 & 1u); ++nuse0; }
          ck.lap(3);
          // incoming side: sum the CS blocks; pair bl/2 holds samples (bl & ~1, bl | 1) as (lo, hi) bf16
          const uint8_t* rb = red + buf * red_bytes + ((bl >> 1) * UPC + j4) * 4;
          for (int k = 0; k < CS; ++k) {
            const uint4 x = *reinterpret_cast<const uint4*>(rb + (uint32_t)k * blk_bytes);
            const uint32_t xs[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) dh[e] += __uint_as_float((bl & 1) ? (xs[e] & 0xffff0000u) : (xs[e] << 16));
          }
        }
        float4 dq[4];
#pragma unroll
        for (int qq = 0; qq < 4; ++qq) dq[qq] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (valid) {
          const float gi[4] = {pg[0].x, pg[0].y, pg[0].z, pg[0].w}, gf[4] = {pg[1].x, pg[1].y, pg[1].z, pg[1].w};
          const float gg[4] = {pg[2].x, pg[2].y, pg[2].z, pg[2].w}, go[4] = {pg[3].x, pg[3].y, pg[3].z, pg[3].w};
          const float cc[4] = {pc.x, pc.y, pc.z, pc.w}, cp[4] = {pcp.x, pcp.y, pcp.z, pcp.w}, de[4] = {pdh.x, pdh.y, pdh.z, pdh.w};
          float o[4][4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float dht = dh[e] + de[e];
            const float tch = tanh_approx(cc[e]);      // the same function the forward applied
            const float dc = dcs[e] + dht * go[e] * (1.f - tch * tch);
            dcs[e] = dc * gf[e];
            o[0][e] = dc * gg[e] * gi[e] * (1.f - gi[e]);
            o[1][e] = dc * cp[e] * gf[e] * (1.f - gf[e]);
            o[2][e] = dc * gi[e] * (1.f - gg[e] * gg[e]);
            o[3][e] = dht * tch * go[e] * (1.f - go[e]);
          }
#pragma unroll
          for (int qq = 0; qq < 4; ++qq) dq[qq] = make_float4(o[qq][0], o[qq][1], o[qq][2], o[qq][3]);
        }
        // next step's B operand (own gate gradients, bf16): element (row bl, k = 32 qq + j4 + e), no-swizzle core matrices
        uint2 d16[4];
#pragma unroll
        for (int qq = 0; qq < 4; ++qq) {
          d16[qq] = make_uint2(pack_bf16(dq[qq].x, dq[qq].y), pack_bf16(dq[qq].z, dq[qq].w));
          const int lr = qq * 32 + j4;
          *reinterpret_cast<uint2*>(Bop + (lr >> 3) * (NBG * 16) + bl * 16 + (lr & 7) * 2) = d16[qq];
        }
        fence_proxy_async();
        group_sync(g);
        ck.lap(4);
        if (bme < B) {
          const int64_t o = bme * gstr + (int64_t)t * ndir * 4 * H + dir * 4 * H + j0 + j4;
#pragma unroll
          for (int qq = 0; qq < 4; ++qq) {
            if (want_f32) *reinterpret_cast<float4*>(d.dgates + o + qq * H) = dq[qq];
            if (dg16) *reinterpret_cast<uint2*>(dg16 + o + qq * H) = d16[qq];
          }
        }
        // re-arm the barrier this step used: partial sums of step s + 2 land there (peers cannot send them before they
        // have this CTA's step s + 1 block, which leaves after this point)
        if (gt == 0 && s > 0 && s + 2 < T) mbar_arrive_expect_tx(&full[buf], red_bytes);
        ck.lap(5);
      }
    }
    if (ss0 + cpd * G < nss) {
      __syncthreads();
      cluster_arrive();
      cluster_wait();
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_arrive();
  cluster_wait();
  if (d.dbg && gt == 0) {
    long long* qd = d.dbg + ((int64_t)blockIdx.x * MAXG + g) * 8;
    for (int i = 0; i < 7; ++i) qd[i] = ck.acc[i];
    qd[7] = clock64() - tstart;
  }
  tmem_free<512>(tmem);
}

// ---------------------------------------------------------------------------------- host side
static size_t fwd_smem_bytes(int H, int G) {
  // >= 120 KB keeps the clusters at one CTA per SM (every CTA allocates all 512 TMEM columns)
  return std::max((size_t)120 * 1024, 1024 + 64 + (size_t)G * fwd_group_bytes(H));
}
static size_t bwd_smem_bytes(int H, int G) {
  return std::max((size_t)120 * 1024, 1024 + 64 + (size_t)G * bwd_group_bytes(H));
}

static bool eligible(const ag_lstm_desc* d) {
  if (d->F != 0 || d->prec < 1 || (d->flags & 1)) return false;
  const int H = d->H;
  if (H == 128 || H == 256 || H == 512) return true;        // H/32 CTAs per cluster (4, 8, 16), whole 128-unit M-tiles
  set_decline("cluster declined: H=%d not in {128, 256, 512}", H);
  return false;
}

static void cluster_cfg(cudaLaunchConfig_t* cfg, cudaLaunchAttribute* at, int CS, int threads, size_t smem, cudaStream_t s) {
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = (unsigned)CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg->attrs = at; cfg->numAttrs = 1;
  cfg->blockDim = dim3((unsigned)threads); cfg->dynamicSmemBytes = smem; cfg->stream = s;
  cfg->gridDim = dim3((unsigned)(CS * 2));
}

// Returns AG_OK with *launched = 1 when the cluster kernel took the call; *launched = 0 -> the caller uses lstm.cu.
template <typename KernT>
static int launch_groups(KernT kern, const ag_lstm_desc* d, int maxg, size_t (*smem_of)(int, int), cudaStream_t s, int* launched) {
  *launched = 0;
  const int CS = d->H / UPC;
  size_t smem = smem_of(d->H, maxg);
  if (smem > (size_t)smem_optin()) { set_decline("cluster declined: %zu B of shared memory", smem); return AG_OK; }
  if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) { cudaGetLastError(); return AG_OK; }
  if (CS > 8 && cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) { cudaGetLastError(); return AG_OK; }
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute at[1];
  cluster_cfg(&cfg, at, CS, GT * maxg, smem, s);
  int nmax = 0;
  if (cudaOccupancyMaxActiveClusters(&nmax, kern, &cfg) != cudaSuccess) { cudaGetLastError(); return AG_OK; }
  if (nmax < d->ndir) {                                     // cannot host one cluster per direction: caller falls back
    set_decline("cluster declined: %d co-resident clusters of %d CTAs < %d directions", nmax, CS, d->ndir);
    return AG_OK;
  }
  // sub-slices of 16 samples per direction -> G warp groups on each of cpd clusters (fewest rounds, then fewest groups)
  const int nss = (d->B + NBG - 1) / NBG, cpd_max = nmax / d->ndir;
  const int G = std::min(maxg, (nss + cpd_max - 1) / cpd_max);
  const int cpd = std::min(cpd_max, (nss + G - 1) / G);
  smem = smem_of(d->H, G);
  cluster_cfg(&cfg, at, CS, GT * G, smem, s);
  cfg.gridDim = dim3((unsigned)(CS * d->ndir * cpd));
  ag_lstm_desc dd = *d;
  AG_CUDA(cudaLaunchKernelEx(&cfg, kern, dd, nss, cpd, G));
  *launched = 1;
  set_path("cluster");
  return AG_OK;
}
int cluster_fwd(const ag_lstm_desc* d, cudaStream_t s, int* launched) {
  *launched = 0;
  if (!eligible(d)) return AG_OK;
  if (!d->hbuf16) { set_decline("cluster declined: hbuf16 missing"); return AG_OK; }
  return launch_groups(lstm_cl_fwd_kernel, d, MAXG, fwd_smem_bytes, s, launched);
}
int cluster_bwd(const ag_lstm_desc* d, cudaStream_t s, int* launched) {
  *launched = 0;
  if (!eligible(d)) return AG_OK;
  return launch_groups(lstm_cl_bwd_kernel, d, MAXG_BWD, bwd_smem_bytes, s, launched);
}

}  // namespace lc
}  // namespace ag

extern "C" int ag_lstm_cluster_max_active(int H, int bwd) {
  using namespace ag::lc;
  const int CS = H / UPC;
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = (unsigned)CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  cfg.blockDim = dim3(LT); cfg.gridDim = dim3((unsigned)(CS * 8));
  int nmax = 0;
  cudaError_t e;
  if (bwd) {
    cfg.dynamicSmemBytes = bwd_smem_bytes(H, MAXG_BWD);
    cfg.blockDim = dim3(GT * MAXG_BWD);
    cudaFuncSetAttribute(lstm_cl_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cfg.dynamicSmemBytes);
    if (CS > 8) cudaFuncSetAttribute(lstm_cl_bwd_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    e = cudaOccupancyMaxActiveClusters(&nmax, lstm_cl_bwd_kernel, &cfg);
  } else {
    cfg.dynamicSmemBytes = fwd_smem_bytes(H, MAXG);
    cfg.blockDim = dim3(GT * MAXG);
    cudaFuncSetAttribute(lstm_cl_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cfg.dynamicSmemBytes);
    if (CS > 8) cudaFuncSetAttribute(lstm_cl_fwd_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    e = cudaOccupancyMaxActiveClusters(&nmax, lstm_cl_fwd_kernel, &cfg);
  }
  if (e != cudaSuccess) { cudaGetLastError(); return -1; }
  return nmax;
}
