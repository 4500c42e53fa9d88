// bf16 tensor-core view-GEMM for sm_100a: tcgen05.mma with the accumulator in TMEM, B (packed bf16 weights)
// staged by TMA (cp.async.bulk.tensor, 128B swizzle), A staged by producer warps that read the strided
// activation *view* (fp32 or bf16, any im2col / channel-prefix addressing of include/audiogan_b200.h),
// convert to bf16 and store the canonical K-major SWIZZLE_128B tile; mbarrier full/empty ring of 4 stages;
// epilogue TMEM -> registers (tcgen05.ld) -> bias / row-bias / skip / LeakyReLU / LeakyReLU' / mask -> global.
//
//   NT: C(m,n) = epilogue( sum_k A(m,k) B(n,k) )           forward + data gradients ("bf16 mode")
//   TN: dW[n,k] += sum_m Y(m,n) A(m,k)                     weight gradients: both operands MN-major
//
// Warp roles (160 threads): warps 0-3 = A producers, then the epilogue (warp w owns TMEM lanes 32w..32w+31);
// warp 4 = TMEM allocator + single-thread MMA issuer.  One 128 x BN output tile per CTA.
#include "common.cuh"
#include "tc_common.cuh"
#include <cuda.h>
#include <mutex>
#include <cstdlib>

namespace ag {
namespace tc {

// profiling aid (ag_gemm_dbg_*): per-phase cycle totals over all CTAs of the NT kernel, off unless enabled
__device__ unsigned long long g_nt_dbg[16];
__device__ int g_nt_dbg_on = 0;

constexpr int BM = 128, BK = 64, NTHREADS = 160, NPROD = 128;
constexpr int NT_NPROD = 256, NT_THREADS = 320;   // NT kernel: 8 producer / epilogue warps, MMA warp, TMA warp (B tiles)
constexpr int NT_NCHT = (BM * 8) / NT_NPROD;      // NT: 16-byte A chunks per producer thread and k-block
constexpr int TN_NPROD = 256, TN_THREADS = 288;   // TN kernel: 8 producer / epilogue warps + the MMA warp

// Column c of a row: element offset (c / inner) * outer_stride + c % inner (32-bit division: columns < 2^31).
__device__ __forceinline__ int64_t col_off(int64_t c, int64_t inner, int64_t outer_stride) {
  const uint32_t c1 = (uint32_t)c / (uint32_t)inner;
  return (int64_t)c1 * outer_stride + (int64_t)((uint32_t)c - c1 * (uint32_t)inner);
}

// NCH 16-byte chunks (8 consecutive columns each) -> packed bf16, with every global load issued before the first
// use so the whole batch is in flight at once (the load latency is paid once per stage, not once per chunk).
// roff[i] < 0 or c[i] >= ncols gives zeros; column `ones_at` reads 1.0 (the bias-gradient column of the TN GEMM).
// MODE 0: fp32 rows, 16-byte aligned, chunks never straddle `inner`;  MODE 1: same for bf16;  MODE 2: any view.
template <int MODE, int NCH>
__device__ __forceinline__ void load_chunks(uint4 (&out)[NCH], const void* base, int dtype, const int64_t (&roff)[NCH],
                                            const int64_t (&c)[NCH], int64_t ncols, int64_t inner, int64_t outer_stride,
                                            int64_t ones_at) {
  if (MODE == 0) {
    float4 lo[NCH], hi[NCH];
    bool ok[NCH];
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      ok[i] = roff[i] >= 0 && c[i] + 8 <= ncols;
      const int64_t off = ok[i] ? roff[i] + col_off(c[i], inner, outer_stride) : 0;
      const float4* p = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(base) + off);
      lo[i] = __ldg(p);
      hi[i] = __ldg(p + 1);
    }
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      out[i].x = pack_bf16(lo[i].x, lo[i].y); out[i].y = pack_bf16(lo[i].z, lo[i].w);
      out[i].z = pack_bf16(hi[i].x, hi[i].y); out[i].w = pack_bf16(hi[i].z, hi[i].w);
      if (!ok[i]) out[i] = make_uint4(0u, 0u, 0u, 0u);
    }
  } else if (MODE == 1) {
    bool ok[NCH];
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      ok[i] = roff[i] >= 0 && c[i] + 8 <= ncols;
      const int64_t off = ok[i] ? roff[i] + col_off(c[i], inner, outer_stride) : 0;
      out[i] = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(base) + off));
    }
#pragma unroll
    for (int i = 0; i < NCH; ++i)
      if (!ok[i]) out[i] = make_uint4(0u, 0u, 0u, 0u);
  }
  if (MODE == 2) {
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      float v[8];
      uint32_t c1 = (uint32_t)c[i] / (uint32_t)inner;
      int64_t cr = c[i] - (int64_t)c1 * inner;
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int64_t cc = c[i] + e;
        const bool okv = roff[i] >= 0 && cc < ncols;
        const float x = ld_any(base, okv ? roff[i] + (int64_t)c1 * outer_stride + cr : 0, dtype);
        v[e] = okv ? x : ((roff[i] >= 0 && cc == ones_at) ? 1.f : 0.f);
        if (++cr == inner) { cr = 0; ++c1; }
      }
      out[i].x = pack_bf16(v[0], v[1]); out[i].y = pack_bf16(v[2], v[3]);
      out[i].z = pack_bf16(v[4], v[5]); out[i].w = pack_bf16(v[6], v[7]);
    }
  }
}

// Two-phase variant for the vector modes: issue() puts the global loads of a stage in flight, get() converts them.
// The NT producers issue stage k+1 before they store stage k, so a stage's load latency overlaps the slot wait, the
// swizzled stores and the proxy fence of the previous one.
template <int MODE, int NCH>
struct ChunkLoader {
  float4 lo[MODE == 0 ? NCH : 1], hi[MODE == 0 ? NCH : 1];
  uint4 raw[MODE == 1 ? NCH : 1];
  bool ok[NCH];
  __device__ __forceinline__ void issue(const void* base, const int64_t (&roff)[NCH], const int64_t (&c)[NCH], int64_t ncols,
                                        int64_t inner, int64_t outer_stride) {
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      ok[i] = roff[i] >= 0 && c[i] + 8 <= ncols;
      const int64_t off = ok[i] ? roff[i] + col_off(c[i], inner, outer_stride) : 0;
      if (MODE == 0) {
        const float4* p = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(base) + off);
        lo[i] = __ldg(p);
        hi[i] = __ldg(p + 1);
      } else {
        raw[i] = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(base) + off));
      }
    }
  }
  // column offsets (and their validity) computed by the caller, once: the producer warps are instruction-bound (one warp
  // per scheduler and CTA), and col_off's division per chunk and stage was most of what they executed
  __device__ __forceinline__ void issue_pre(const void* base, const int64_t (&roff)[NCH], const int32_t (&coff)[NCH], uint32_t cmask) {
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      ok[i] = roff[i] >= 0 && ((cmask >> i) & 1u);
      const int64_t off = ok[i] ? roff[i] + coff[i] : 0;
      if (MODE == 0) {
        const float4* p = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(base) + off);
        lo[i] = __ldg(p);
        hi[i] = __ldg(p + 1);
      } else {
        raw[i] = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(base) + off));
      }
    }
  }
  __device__ __forceinline__ void issue_col(const void* base, const int64_t (&roff)[NCH], int64_t coff, bool cok) {
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      ok[i] = roff[i] >= 0 && cok;
      const int64_t off = ok[i] ? roff[i] + coff : 0;
      if (MODE == 0) {
        const float4* p = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(base) + off);
        lo[i] = __ldg(p);
        hi[i] = __ldg(p + 1);
      } else {
        raw[i] = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(base) + off));
      }
    }
  }
  __device__ __forceinline__ void get(uint4 (&out)[NCH]) const {
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      if (MODE == 0) {
        out[i].x = pack_bf16(lo[i].x, lo[i].y); out[i].y = pack_bf16(lo[i].z, lo[i].w);
        out[i].z = pack_bf16(hi[i].x, hi[i].y); out[i].w = pack_bf16(hi[i].z, hi[i].w);
      } else {
        out[i] = raw[i];
      }
      if (!ok[i]) out[i] = make_uint4(0u, 0u, 0u, 0u);
    }
  }
};

// pipeline depth per tile width: BN = 256 -> 2 stages (96 KB) so that TWO CTAs are resident per SM and one CTA's
// epilogue overlaps the other's main loop (TMEM: 2 x 256 columns); narrower tiles get 3-4 stages, still 2 CTAs/SM.
__host__ __device__ constexpr int nt_stages(int BN) { return BN >= 256 ? 2 : (BN >= 128 ? 3 : 4); }
constexpr int TRLD = 36;    // row stride (floats) of the 32x32 epilogue transpose tiles: 16-byte rows, conflict-free

__device__ __forceinline__ float4 ld4_any(const void* p, int64_t i, int dtype) {
  if (dtype == 0) return *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p) + i);
  const uint2 u = *reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(p) + i);
  const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&u.x), b = *reinterpret_cast<const __nv_bfloat162*>(&u.y);
  return make_float4(__low2float(a), __high2float(a), __low2float(b), __high2float(b));
}

template <int BN, int MODE, bool VECC>
__global__ void __launch_bounds__(NT_THREADS, 2) gemm_nt_tc_kernel(const ag_gemm_desc d, const __grid_constant__ CUtensorMap mapB) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  constexpr int STG = nt_stages(BN);
  // 1024-byte alignment for the 128B-swizzled tiles (pointer arithmetic on the __shared__ array keeps LDS/STS)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  constexpr int A_BYTES = BM * BK * 2, B_BYTES = BN * BK * 2, STAGE_BYTES = A_BYTES + B_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + STG * STAGE_BYTES);
  uint64_t* empty = full + STG;
  uint64_t* tmem_full = empty + STG;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);
  int64_t* rowoff = reinterpret_cast<int64_t*>(tmem_slot + 2);   // [BM] A row offsets, later C row offsets
  int* s_b = reinterpret_cast<int*>(rowoff + BM);                 // [BM] batch, [BM] t, [BM] mask length
  int* s_t = s_b + BM;
  int* s_ml = s_t + BM;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const int nkb = (int)((d.K + BK - 1) / BK);

  if (tid < BM) {
    const int64_t m = m0 + tid;
    int64_t off = -1;
    if (m < d.M) { const int64_t b = m / d.a_rpb; off = b * d.a_bs + (m - b * d.a_rpb) * d.a_rs; }
    rowoff[tid] = off;
  }
  if (tid == 0) {
    for (int s = 0; s < STG; ++s) { mbar_init(&full[s], NT_NPROD / 32 + 1); mbar_init(&empty[s], 1); }
    mbar_init(tmem_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  constexpr uint32_t TMEM_COLS = BN < 32 ? 32 : BN;     // power of two >= 32 (BN in {16,32,64,128,256})
  if (warp == NT_NPROD / 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < NT_NPROD / 32) {
    // ------------------------------------------------------------------ producers (8 warps: they are instruction-bound)
    constexpr int NCHT = NT_NCHT;
    int64_t ro[NCHT];
#pragma unroll
    for (int i = 0; i < NCHT; ++i) ro[i] = rowoff[(i * NT_NPROD + tid) >> 3];
    auto cols_of = [&](int kb, int64_t (&cc)[NT_NCHT]) {
#pragma unroll
      for (int i = 0; i < NCHT; ++i) cc[i] = (int64_t)kb * BK + ((i * NT_NPROD + tid) & 7) * 8;
    };
    ChunkLoader<MODE == 2 ? 0 : MODE, NCHT> ld;
    // all chunks of a thread sit in the same 8-column group ((i*256 + tid) & 7 == tid & 7): one column offset per k-block
    auto issue_kb = [&](int kb) {
      const int64_t c = (int64_t)kb * BK + (tid & 7) * 8;
      ld.issue_col(d.A, ro, col_off(c, d.a_kin, d.a_k1s), c + 8 <= d.K);
    };
    if (MODE != 2) issue_kb(0);
    const bool dbg = g_nt_dbg_on != 0 && tid == 0;
    long long t_get = 0, t_empty = 0, t_store = 0, t0 = 0;
    const long long t_begin = clock64();
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb % STG;
      const uint32_t ph = (kb / STG) & 1;
      uint4 ch[NCHT];
      if (dbg) t0 = clock64();
      if (MODE != 2) {
        ld.get(ch);                                   // stage kb has landed (or we wait for it here)
        if (kb + 1 < nkb) issue_kb(kb + 1);           // put stage kb+1 in flight before touching shared memory
      } else {
        int64_t cc[NCHT];
        cols_of(kb, cc);
        load_chunks<2, NCHT>(ch, d.A, d.a_dtype, ro, cc, d.K, d.a_kin, d.a_k1s, -1);
      }
      if (dbg) { asm volatile("" :: "r"(ch[0].x), "r"(ch[NCHT - 1].w)); const long long t1 = clock64(); t_get += t1 - t0; t0 = t1; }
      mbar_wait(&empty[s], ph ^ 1);
      if (dbg) { const long long t1 = clock64(); t_empty += t1 - t0; t0 = t1; }
      uint8_t* sa = smem + s * STAGE_BYTES;
#pragma unroll
      for (int i = 0; i < NCHT; ++i) {
        const int cid = i * NT_NPROD + tid;
        const int r = cid >> 3, c = cid & 7;
        *reinterpret_cast<uint4*>(sa + r * 128 + ((c ^ (r & 7)) << 4)) = ch[i];
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&full[s]);
      if (dbg) { const long long t1 = clock64(); t_store += t1 - t0; }
    }
    const long long t_main = clock64();
    // ------------------------------------------------------------------ epilogue
    // TMEM lane = output row.  Each warp transposes 32x32 blocks through shared memory (the pipeline stages are idle
    // by now) so that a warp instruction covers contiguous columns of a row: coalesced (vector) stores and reads.
    {
      const int64_t m = m0 + tid;               // every producer thread has passed its last rowoff read: reuse it
      __syncwarp();
      asm volatile("bar.sync 1, %0;" ::"n"(NT_NPROD) : "memory");
      if (tid >= BM) {
      } else if (m < d.M) {
        const int64_t b = m / d.c_rpb, t = m - b * d.c_rpb;
        rowoff[tid] = b * d.c_bs + t * d.c_rs;
        s_b[tid] = (int)b;
        s_t[tid] = (int)t;
        s_ml[tid] = d.mask_len ? d.mask_len[b] : 0;
      } else {
        rowoff[tid] = -1;
      }
    }
    asm volatile("bar.sync 1, %0;" ::"n"(NT_NPROD) : "memory");       // C row offsets visible to all 8 epilogue warps
    mbar_wait(tmem_full, 0);
    tc_fence_after();
    __syncwarp();
    float* tr = reinterpret_cast<float*>(smem) + warp * (32 * TRLD);
    const int wq = warp & 3, wh = warp >> 2;        // TMEM lane quarter; warps w and w + 4 take alternate column chunks
    const float alpha = d.alpha == 0.f ? 1.f : d.alpha;
    constexpr int CH = BN < 32 ? BN : 32;
    const uint32_t tlane = tmem_base + ((uint32_t)(wq * 32) << 16);
#pragma unroll 1
    for (int c0 = wh * CH; c0 < BN; c0 += 2 * CH) {
      if (n0 + c0 >= d.N) break;
      {
        uint32_t v[16];
        tc_ld16(tlane + (uint32_t)c0, v);
#pragma unroll
        for (int j = 0; j < 4; ++j)
          *reinterpret_cast<uint4*>(tr + lane * TRLD + 4 * j) = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        if (CH == 32) {
          tc_ld16(tlane + (uint32_t)(c0 + 16), v);
#pragma unroll
          for (int j = 0; j < 4; ++j)
            *reinterpret_cast<uint4*>(tr + lane * TRLD + 16 + 4 * j) = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        }
      }
      __syncwarp();
      if (VECC) {
        const int q = lane & 7, rs = lane >> 3;
        const int64_t n = n0 + c0 + 4 * q;
        const bool nv = 4 * q < CH && n < d.N;
        const int64_t n1 = (nv && n >= d.c_nin) ? (int64_t)((uint32_t)n / (uint32_t)d.c_nin) : 0;   // plain 2-D C: no division
        const int64_t coff = n1 * d.c_n1s + (n - n1 * d.c_nin);
        float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
        if (nv && d.bias) bv = *reinterpret_cast<const float4*>(d.bias + ((d.bias_mod > 0 && n >= d.bias_mod) ? n % d.bias_mod : n));
        const int64_t mpos = n1 * d.mask_n1mul + d.mask_toff;
        // all global reads of a half-chunk's epilogue operands first (4 rows x {row-bias, skip, dact}): one exposed latency
        // per half (8 rows at once cost 112 registers of staging and spilled under the 96-register budget of 640 threads/SM)
#pragma unroll 1
        for (int hb = 0; hb < 8; hb += 4) {
        float4 rbv[4], skv[4], dav[4];
        int64_t civ[4];
#pragma unroll
        for (int it = 0; it < 4; ++it) {
          const int row = wq * 32 + (hb + it) * 4 + rs;
          const int64_t crow = rowoff[row];
          civ[it] = (crow < 0 || !nv) ? -1 : crow + coff;
          rbv[it] = skv[it] = make_float4(0.f, 0.f, 0.f, 0.f);
          dav[it] = make_float4(1.f, 1.f, 1.f, 1.f);
          if (civ[it] >= 0) {
            if (d.rowbias) rbv[it] = *reinterpret_cast<const float4*>(d.rowbias + (int64_t)s_b[row] * d.rowbias_ld + n);
            if (d.skip) skv[it] = ld4_any(d.skip, civ[it], d.aux_dtype);
            if (d.dact) dav[it] = ld4_any(d.dact, civ[it], d.aux_dtype);
          }
        }
#pragma unroll
        for (int it = 0; it < 4; ++it) {
          const int r = (hb + it) * 4 + rs, row = wq * 32 + r;
          if (civ[it] < 0) continue;
          const int64_t ci = civ[it];
          float4 x = *reinterpret_cast<const float4*>(tr + r * TRLD + 4 * q);
          x.x = x.x * alpha + bv.x + rbv[it].x + skv[it].x; x.y = x.y * alpha + bv.y + rbv[it].y + skv[it].y;
          x.z = x.z * alpha + bv.z + rbv[it].z + skv[it].z; x.w = x.w * alpha + bv.w + rbv[it].w + skv[it].w;
          if (d.act == 1) {
            x.x = x.x > 0.f ? x.x : x.x * d.slope; x.y = x.y > 0.f ? x.y : x.y * d.slope;
            x.z = x.z > 0.f ? x.z : x.z * d.slope; x.w = x.w > 0.f ? x.w : x.w * d.slope;
          }
          if (d.dact) {
            x.x *= dav[it].x > 0.f ? 1.f : d.slope; x.y *= dav[it].y > 0.f ? 1.f : d.slope;
            x.z *= dav[it].z > 0.f ? 1.f : d.slope; x.w *= dav[it].w > 0.f ? 1.f : d.slope;
          }
          if (d.mask_len) {
            const int64_t pos = (int64_t)s_t[row] * d.mask_tmul + mpos;
            if (pos < 0 || pos >= s_ml[row]) x = make_float4(0.f, 0.f, 0.f, 0.f);
          }
          if (d.c_dtype == 0) {
            *reinterpret_cast<float4*>(reinterpret_cast<float*>(d.C) + ci) = x;
          } else {
            uint2 o;
            o.x = pack_bf16(x.x, x.y); o.y = pack_bf16(x.z, x.w);
            *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(d.C) + ci) = o;
          }
        }
        }
      } else {
        const int64_t n = n0 + c0 + lane;
        const bool nv = lane < CH && n < d.N;
        const int64_t n1 = nv ? n / d.c_nin : 0;
        const int64_t coff = n1 * d.c_n1s + (n - n1 * d.c_nin);
        const float bv = (nv && d.bias) ? d.bias[d.bias_mod > 0 ? n % d.bias_mod : n] : 0.f;
#pragma unroll 2
        for (int r = 0; r < 32; ++r) {
          const int row = wq * 32 + r;
          const int64_t crow = rowoff[row];
          if (crow < 0 || !nv) continue;
          const int64_t ci = crow + coff;
          float x = tr[r * TRLD + lane] * alpha + bv;
          if (d.rowbias) x += d.rowbias[(int64_t)s_b[row] * d.rowbias_ld + n];
          if (d.skip) x += ld_any(d.skip, ci, d.aux_dtype);
          if (d.act == 1) x = x > 0.f ? x : x * d.slope;
          if (d.dact) x *= (ld_any(d.dact, ci, d.aux_dtype) > 0.f) ? 1.f : d.slope;
          if (d.mask_len) {
            const int64_t pos = (int64_t)s_t[row] * d.mask_tmul + n1 * d.mask_n1mul + d.mask_toff;
            if (pos < 0 || pos >= s_ml[row]) x = 0.f;
          }
          st_any(d.C, ci, x, d.c_dtype);
        }
      }
      __syncwarp();
    }
    tc_fence_before();
    if (dbg) {
      const long long t_end = clock64();
      atomicAdd(&g_nt_dbg[0], 1ull);
      atomicAdd(&g_nt_dbg[1], (unsigned long long)t_get);
      atomicAdd(&g_nt_dbg[2], (unsigned long long)t_empty);
      atomicAdd(&g_nt_dbg[3], (unsigned long long)t_store);
      atomicAdd(&g_nt_dbg[4], (unsigned long long)(t_main - t_begin));
      atomicAdd(&g_nt_dbg[5], (unsigned long long)(t_end - t_main));
      atomicAdd(&g_nt_dbg[6], (unsigned long long)nkb);
    }
  } else if (warp == NT_NPROD / 32 + 1) {
    // ------------------------------------------------------------------ TMA issuer for the B (weight) tiles: its own
    // thread, so a tile's TMA latency starts the moment the slot is free instead of after the producers' global loads
    // of the same k-block have landed (the two latencies used to add up on the per-k-block critical path)
    if (lane == 0) {
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % STG;
        mbar_wait(&empty[s], ((kb / STG) & 1) ^ 1);
        mbar_arrive_expect_tx(&full[s], B_BYTES);
        tma_load_2d(smem + s * STAGE_BYTES + A_BYTES, &mapB, &full[s], kb * BK, n0);
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ MMA issuer (one thread)
    if (lane == 0) {
      const uint32_t idesc = umma_idesc(BM, BN, 0, 0);
      const bool dbgm = g_nt_dbg_on != 0;
      long long t_full = 0;
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % STG;
        const long long q0 = dbgm ? clock64() : 0;
        mbar_wait(&full[s], (kb / STG) & 1);
        if (dbgm) t_full += clock64() - q0;
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + s * STAGE_BYTES);
        const uint64_t da = umma_desc(sa, 16, 1024), db = umma_desc(sa + A_BYTES, 16, 1024);
#pragma unroll
        for (int k = 0; k < BK / 16; ++k)          // +32 B per UMMA_K inside the 128 B swizzle row
          tc_mma(tmem_base, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (kb | k) ? 1u : 0u);
        tc_commit(&empty[s]);
      }
      tc_commit(tmem_full);
      if (dbgm) atomicAdd(&g_nt_dbg[7], (unsigned long long)t_full);
    }
    __syncwarp();
  }
  __syncthreads();
  if (warp == NT_NPROD / 32) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// ---------------------------------------------------------------------------------------- host
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  });
  return fn;
}

// 2-D bf16 map: inner dim `cols` (contiguous), outer dim `rows` with stride ld elements; box {box_c, box_r}, 128B swizzle.
static int make_map_2d(CUtensorMap* map, const void* ptr, int64_t rows, int64_t cols, int64_t ld, int box_c, int box_r) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) { set_error("cuTensorMapEncodeTiled entry point not available"); return AG_ENOTSUP; }
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)box_c, (cuuint32_t)box_r};
  cuuint32_t es[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d): rows %lld cols %lld ld %lld", (int)r, (long long)rows, (long long)cols, (long long)ld); return AG_ECUDA; }
  return AG_OK;
}

template <int BN, int MODE, bool VECC>
static int launch_nt2(const ag_gemm_desc* d, cudaStream_t s) {
  CUtensorMap mapB;
  int rc = make_map_2d(&mapB, d->B, d->N, d->K, d->ldb, BK, BN);
  if (rc) return rc;
  constexpr int STG = nt_stages(BN);
  constexpr int smem = STG * (BM * BK * 2 + BN * BK * 2) + 1024 /*align*/ + (2 * STG + 1) * 8 + 16 + BM * 8 + 3 * BM * 4;
  static_assert(STG * (BM * BK * 2 + BN * BK * 2) >= 8 * 32 * TRLD * 4, "epilogue transpose tiles must fit in the stages");
  auto kern = gemm_nt_tc_kernel<BN, MODE, VECC>;
  static bool attr_set = false;      // once per instantiation (a driver call per launch otherwise)
  if (!attr_set) { AG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); attr_set = true; }
  dim3 grid((unsigned)((d->M + BM - 1) / BM), (unsigned)((d->N + BN - 1) / BN));
  kern<<<grid, NT_THREADS, smem, s>>>(*d, mapB);
  AG_LAUNCH_CHECK();
  return AG_OK;
}
template <int BN, int MODE>
static int launch_nt(const ag_gemm_desc* d, cudaStream_t s) {
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  const bool cal = d->c_dtype == 0 ? al16(d->C) : (reinterpret_cast<uintptr_t>(d->C) & 7) == 0;
  const bool auxal = d->aux_dtype == 0 ? (al16(d->skip) && al16(d->dact))
                                       : ((reinterpret_cast<uintptr_t>(d->skip) & 7) == 0 && (reinterpret_cast<uintptr_t>(d->dact) & 7) == 0);
  const bool vecc = d->N % 4 == 0 && d->c_nin % 4 == 0 && d->c_bs % 4 == 0 && d->c_rs % 4 == 0 && d->c_n1s % 4 == 0 && cal && auxal &&
                    al16(d->bias) && (d->bias_mod == 0 || d->bias_mod % 4 == 0) && al16(d->rowbias) && d->rowbias_ld % 4 == 0;
  return vecc ? launch_nt2<BN, MODE, true>(d, s) : launch_nt2<BN, MODE, false>(d, s);
}

// ======================================================================================== NT, persistent, TMA-fed A
// For bf16 activations whose im2col row is ONE contiguous window (every conv / transposed-conv / linear view but the
// generator's channel-prefix reads), A is a 3-D tensor map {k, t, batch} with strides {a_rs, a_bs} -- overlapping rows
// are fine for TMA -- and a 128 x 64 box lands in shared memory as the canonical K-major SWIZZLE_128B tile: no producer
// warps, no register staging.  One persistent CTA per SM walks the tiles; the accumulator is double-buffered in TMEM
// (2 x BN columns) so the 8 epilogue warps drain tile i while the MMA warp runs tile i+1.  Rows are tiled per batch
// (row tile = (batch, t0): rows past the batch's end are zero-filled by TMA and skipped by the epilogue).
//   warps 0-7: epilogue (warp w owns TMEM lanes 32*(w&3).., warps w and w+4 take alternate 32-column chunks)
//   warp 8: TMA producer (one thread)      warp 9: TMEM allocator + MMA issuer (one thread)
__host__ __device__ constexpr int tma_stages(int BN) { return BN >= 256 ? 3 : (BN >= 128 ? 5 : (BN >= 64 ? 6 : 8)); }
constexpr int TM_THREADS = 320, TM_NEPI = 256;

__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

// Lean vector epilogue of one 32-row x CH-column chunk (the persistent kernel is epilogue-bound: the general code above
// costs ~30 instructions per element in 64-bit index arithmetic and per-row tables).  Everything per row comes from ONE
// 16-byte table entry {C row offset | -1, row-bias offset, mask position base, mask length}; offsets are 32-bit element
// offsets (the host checks the spans); the epilogue operands of all 8 rows of a lane are in flight before the first use.
// F: compile-time feature set (bit 0 skip, 1 LeakyReLU' operand, 2 row bias, 3 length mask, 4 LeakyReLU; alpha == 1), so a
// launch only executes the instructions of the epilogue it asked for; F < 0: every feature decided at run time.
enum { EPI_SKIP = 1, EPI_DACT = 2, EPI_RB = 4, EPI_MASK = 8, EPI_ACT = 16 };
__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

__device__ __forceinline__ void tma_prefetch_4d(const CUtensorMap* map, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global.tile [%0, {%1, %2, %3, %4}];" ::"l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}

template <bool CBF, bool AUXBF, int CH, int F>
__device__ __forceinline__ void epi_chunk_vec(const ag_gemm_desc& d, const float* __restrict__ tr, const int4* __restrict__ rowtab,
                                              int wq, int lane, int nc, float alpha) {
  const int q = lane & 7, rs = lane >> 3;
  const int n = nc + 4 * q, N = (int)d.N, cnin = (int)d.c_nin;
  const bool nv = 4 * q < CH && n < N;
  const int n1 = (nv && n >= cnin) ? (int)((uint32_t)n / (uint32_t)cnin) : 0;
  const int coff = n1 * (int)d.c_n1s + (n - n1 * cnin);
  float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
  if (nv && d.bias) {
    const int bm = (int)d.bias_mod;
    bv = *reinterpret_cast<const float4*>(d.bias + ((bm > 0 && n >= bm) ? (int)((uint32_t)n % (uint32_t)bm) : n));
  }
  const int mposl = n1 * (int)d.mask_n1mul;
  const bool has_skip = F < 0 ? d.skip != nullptr : (F & EPI_SKIP) != 0, has_dact = F < 0 ? d.dact != nullptr : (F & EPI_DACT) != 0;
  const bool has_rb = F < 0 ? d.rowbias != nullptr : (F & EPI_RB) != 0, has_mask = F < 0 ? d.mask_len != nullptr : (F & EPI_MASK) != 0;
  const bool has_act = F < 0 ? d.act == 1 : (F & EPI_ACT) != 0;
  const float slope = d.slope;
  int ci[8];
  uint32_t keep = 0;
  float4 rb[8];
  uint2 sk2[AUXBF ? 8 : 1], da2[AUXBF ? 8 : 1];
  float4 sk4[AUXBF ? 1 : 8], da4[AUXBF ? 1 : 8];
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int4 rt = rowtab[wq * 32 + it * 4 + rs];
    const bool ok = nv && rt.x >= 0;
    ci[it] = ok ? rt.x + coff : -1;
    keep |= ((!has_mask || (uint32_t)(rt.z + mposl) < (uint32_t)rt.w) ? 1u : 0u) << it;
    rb[it] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (has_rb && ok) rb[it] = *reinterpret_cast<const float4*>(d.rowbias + (rt.y + n));
    if (AUXBF) {
      sk2[it] = make_uint2(0u, 0u);
      da2[it] = make_uint2(0x3f803f80u, 0x3f803f80u);
      if (has_skip && ok) sk2[it] = *reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(d.skip) + ci[it]);
      if (has_dact && ok) da2[it] = *reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(d.dact) + ci[it]);
    } else {
      sk4[it] = make_float4(0.f, 0.f, 0.f, 0.f);
      da4[it] = make_float4(1.f, 1.f, 1.f, 1.f);
      if (has_skip && ok) sk4[it] = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(d.skip) + ci[it]);
      if (has_dact && ok) da4[it] = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(d.dact) + ci[it]);
    }
  }
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    if (ci[it] < 0) continue;
    float4 x = *reinterpret_cast<const float4*>(tr + (it * 4 + rs) * TRLD + 4 * q);
    float4 sk, da;
    if (AUXBF) {
      sk = make_float4(__uint_as_float(sk2[it].x << 16), __uint_as_float(sk2[it].x & 0xffff0000u),
                       __uint_as_float(sk2[it].y << 16), __uint_as_float(sk2[it].y & 0xffff0000u));
      da = make_float4(__uint_as_float(da2[it].x << 16), __uint_as_float(da2[it].x & 0xffff0000u),
                       __uint_as_float(da2[it].y << 16), __uint_as_float(da2[it].y & 0xffff0000u));
    } else {
      sk = sk4[it];
      da = da4[it];
    }
    if (F < 0) { x.x *= alpha; x.y *= alpha; x.z *= alpha; x.w *= alpha; }
    x.x += bv.x; x.y += bv.y; x.z += bv.z; x.w += bv.w;
    if (has_rb) { x.x += rb[it].x; x.y += rb[it].y; x.z += rb[it].z; x.w += rb[it].w; }
    if (has_skip) { x.x += sk.x; x.y += sk.y; x.z += sk.z; x.w += sk.w; }
    if (has_act) {
      x.x = x.x > 0.f ? x.x : x.x * slope; x.y = x.y > 0.f ? x.y : x.y * slope;
      x.z = x.z > 0.f ? x.z : x.z * slope; x.w = x.w > 0.f ? x.w : x.w * slope;
    }
    if (has_dact) {
      x.x *= da.x > 0.f ? 1.f : slope; x.y *= da.y > 0.f ? 1.f : slope;
      x.z *= da.z > 0.f ? 1.f : slope; x.w *= da.w > 0.f ? 1.f : slope;
    }
    if (has_mask && !((keep >> it) & 1u)) x = make_float4(0.f, 0.f, 0.f, 0.f);
    if (CBF) {
      uint2 o;
      o.x = pack_bf16(x.x, x.y); o.y = pack_bf16(x.z, x.w);
      *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(d.C) + ci[it]) = o;
    } else {
      *reinterpret_cast<float4*>(reinterpret_cast<float*>(d.C) + ci[it]) = x;
    }
  }
}

template <int BN, bool VECC>
__global__ void __launch_bounds__(TM_THREADS, 1)
gemm_nt_tma_kernel(const ag_gemm_desc d, const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                   int R, int tpb, int ntn, int total_tiles, int pfx_G, int pfx_KT, const __grid_constant__ CUtensorMap mapS,
                   const __grid_constant__ CUtensorMap mapD, int pf_flags, int pf_cn) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  constexpr int STG = tma_stages(BN);
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  constexpr int A_BYTES = BM * BK * 2, B_BYTES = BN * BK * 2, STAGE_BYTES = A_BYTES + B_BYTES;
  float* stage_tr = reinterpret_cast<float*>(smem + STG * STAGE_BYTES);          // 8 warps x 32 x TRLD floats
  int4* rowtab = reinterpret_cast<int4*>(stage_tr + 8 * 32 * TRLD);              // [BM] the vector epilogue's per-row entry
  uint64_t* full = reinterpret_cast<uint64_t*>(rowtab + BM);
  uint64_t* empty = full + STG;
  uint64_t* tfull = empty + STG;          // [2] accumulator ready
  uint64_t* tempty = tfull + 2;           // [2] accumulator drained
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  int64_t* rowoff = reinterpret_cast<int64_t*>(tmem_slot + 2);   // [BM] C row offsets of the tile in the epilogue
  int* s_b = reinterpret_cast<int*>(rowoff + BM);
  int* s_t = s_b + BM;
  int* s_ml = s_t + BM;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // channel-prefix views.  a_layout 1 (pfx_G > 0): k-block = (channel group of 8, block of 8 taps), a 4-D box {8 c, 128 rows, 8 taps};
  // a_layout 2 (pfx_G < 0: -pfx_G groups of 64 channels, pfx_KT taps): k-block = (tap, channel group), a box {64 c, 128 rows, 1 tap}
  const int pfxG = pfx_G < 0 ? -pfx_G : pfx_G;
  const int nkb = pfx_G ? pfxG * pfx_KT : (int)((d.K + BK - 1) / BK);
  if (tid == 0) {
    for (int s = 0; s < STG; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], TM_NEPI / 32); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  constexpr uint32_t ACC_COLS = BN < 32 ? 32 : BN;
  constexpr uint32_t TMEM_COLS = 2 * ACC_COLS;
  if (warp == 9) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 8) {
    // ------------------------------------------------------------------ epilogue warps
    float* tr = stage_tr + warp * (32 * TRLD);
    const int wq = warp & 3, wh = warp >> 2;
    const float alpha = d.alpha == 0.f ? 1.f : d.alpha;
    constexpr int CH = BN < 32 ? BN : 32;
    int lt = 0;
    // specialised epilogues: bf16 output, bf16 (or no) epilogue operands, alpha == 1; -1 = the run-time-flag version
    const bool aux_any = d.skip != nullptr || d.dact != nullptr;
    const int epi_sel = (d.c_dtype == 1 && (!aux_any || d.aux_dtype == 1) && alpha == 1.f)
                            ? ((d.skip ? EPI_SKIP : 0) | (d.dact ? EPI_DACT : 0) | (d.rowbias ? EPI_RB : 0) | (d.mask_len ? EPI_MASK : 0) |
                               (d.act == 1 ? EPI_ACT : 0))
                            : -1;
    const int epi_sel32 = (d.c_dtype == 0 && (!aux_any || d.aux_dtype == 1) && alpha == 1.f)
                              ? ((d.skip ? EPI_SKIP : 0) | (d.dact ? EPI_DACT : 0) | (d.rowbias ? EPI_RB : 0) | (d.mask_len ? EPI_MASK : 0) |
                                 (d.act == 1 ? EPI_ACT : 0))
                              : -1;
    const int dbgf = g_nt_dbg_on;
    const bool dbg = dbgf != 0 && tid == 0;
    long long t_wait = 0, t_work = 0, q0 = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++lt) {
      const int mt = tile / ntn, nt = tile - mt * ntn;
      const int bt = mt / tpb, t0 = (mt - bt * tpb) * BM;
      const int n0 = nt * BN;
      const int acc = lt & 1;
      if (dbg) q0 = clock64();
      asm volatile("bar.sync 1, %0;" ::"n"(TM_NEPI) : "memory");       // the previous tile's row table is no longer read
      if (tid < BM) {
        const int t = t0 + tid;
        if (t < R) {
          const int64_t m = (int64_t)bt * R + t;
          const int64_t b = m / d.c_rpb, tc_ = m - b * d.c_rpb;
          const int ml = d.mask_len ? d.mask_len[b] : 0;
          rowoff[tid] = b * d.c_bs + tc_ * d.c_rs;
          s_b[tid] = (int)b;
          s_t[tid] = (int)tc_;
          s_ml[tid] = ml;
          if (VECC) rowtab[tid] = make_int4((int)(b * d.c_bs + tc_ * d.c_rs), (int)(b * d.rowbias_ld),
                                            (int)(tc_ * d.mask_tmul + d.mask_toff), ml);
        } else {
          rowoff[tid] = -1;
          if (VECC) rowtab[tid] = make_int4(-1, 0, 0, 0);
        }
      }
      asm volatile("bar.sync 1, %0;" ::"n"(TM_NEPI) : "memory");
      mbar_wait(&tfull[acc], (lt >> 1) & 1);
      tc_fence_after();
      if (dbg) { const long long q1 = clock64(); t_wait += q1 - q0; q0 = q1; }
      const uint32_t tlane = tmem_base + (uint32_t)(acc * ACC_COLS) + ((uint32_t)(wq * 32) << 16);
#pragma unroll 1
      for (int c0 = wh * CH; c0 < BN; c0 += 2 * CH) {
        if (n0 + c0 >= d.N) break;
        if (!(dbgf & 2)) {
          uint32_t v[16];
          tc_ld16(tlane + (uint32_t)c0, v);
#pragma unroll
          for (int j = 0; j < 4; ++j)
            *reinterpret_cast<uint4*>(tr + lane * TRLD + 4 * j) = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          if (CH == 32) {
            tc_ld16(tlane + (uint32_t)(c0 + 16), v);
#pragma unroll
            for (int j = 0; j < 4; ++j)
              *reinterpret_cast<uint4*>(tr + lane * TRLD + 16 + 4 * j) = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          }
        }
        __syncwarp();
        if (dbgf & 8) {
        } else if (VECC) {
#define AG_EPI(FL) case FL: epi_chunk_vec<true, true, CH, FL>(d, tr, rowtab, wq, lane, n0 + c0, alpha); break;
          if (epi_sel >= 0) {
            switch (epi_sel) {       // bf16 output (+ bf16 operands): the feature sets the training step uses, specialised
              AG_EPI(0) AG_EPI(EPI_ACT) AG_EPI(EPI_SKIP | EPI_ACT) AG_EPI(EPI_SKIP | EPI_MASK | EPI_ACT) AG_EPI(EPI_MASK | EPI_ACT)
              AG_EPI(EPI_DACT) AG_EPI(EPI_SKIP) AG_EPI(EPI_MASK) AG_EPI(EPI_SKIP | EPI_DACT) AG_EPI(EPI_DACT | EPI_MASK)
              default: epi_chunk_vec<true, true, CH, -1>(d, tr, rowtab, wq, lane, n0 + c0, alpha);
            }
          } else if (epi_sel32 >= 0) {
            switch (epi_sel32) {     // fp32 output (LSTM input projection: row bias; data gradient into the recurrent state: skip)
#define AG_EPI32(FL) case FL: epi_chunk_vec<false, true, CH, FL>(d, tr, rowtab, wq, lane, n0 + c0, alpha); break;
              AG_EPI32(0) AG_EPI32(EPI_RB) AG_EPI32(EPI_SKIP)
#undef AG_EPI32
              default: epi_chunk_vec<false, true, CH, -1>(d, tr, rowtab, wq, lane, n0 + c0, alpha);
            }
          } else if (d.c_dtype) {
            epi_chunk_vec<true, false, CH, -1>(d, tr, rowtab, wq, lane, n0 + c0, alpha);
          } else {
            if (d.aux_dtype) epi_chunk_vec<false, true, CH, -1>(d, tr, rowtab, wq, lane, n0 + c0, alpha);
            else epi_chunk_vec<false, false, CH, -1>(d, tr, rowtab, wq, lane, n0 + c0, alpha);
          }
#undef AG_EPI
        } else {
          const int64_t n = n0 + c0 + lane;
          const bool nv = lane < CH && n < d.N;
          const int64_t n1 = nv ? n / d.c_nin : 0;
          const int64_t coff = n1 * d.c_n1s + (n - n1 * d.c_nin);
          const float bv = (nv && d.bias) ? d.bias[d.bias_mod > 0 ? n % d.bias_mod : n] : 0.f;
#pragma unroll 2
          for (int r = 0; r < 32; ++r) {
            const int row = wq * 32 + r;
            const int64_t crow = rowoff[row];
            if (crow < 0 || !nv) continue;
            const int64_t ci = crow + coff;
            float x = tr[r * TRLD + lane] * alpha + bv;
            if (d.rowbias) x += d.rowbias[(int64_t)s_b[row] * d.rowbias_ld + n];
            if (d.skip) x += ld_any(d.skip, ci, d.aux_dtype);
            if (d.act == 1) x = x > 0.f ? x : x * d.slope;
            if (d.dact) x *= (ld_any(d.dact, ci, d.aux_dtype) > 0.f) ? 1.f : d.slope;
            if (d.mask_len) {
              const int64_t pos = (int64_t)s_t[row] * d.mask_tmul + n1 * d.mask_n1mul + d.mask_toff;
              if (pos < 0 || pos >= s_ml[row]) x = 0.f;
            }
            st_any(d.C, ci, x, d.c_dtype);
          }
        }
        __syncwarp();
      }
      // this warp's share of the accumulator is in registers / memory: hand the TMEM buffer back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);
      if (dbg) t_work += clock64() - q0;
    }
    if (dbg) {
      atomicAdd(&g_nt_dbg[0], (unsigned long long)lt);
      atomicAdd(&g_nt_dbg[1], (unsigned long long)t_wait);
      atomicAdd(&g_nt_dbg[2], (unsigned long long)t_work);
    }
  } else if (warp == 8) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int it = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int mt = tile / ntn, nt = tile - mt * ntn;
        const int bt = mt / tpb, t0 = (mt - bt * tpb) * BM;
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % STG;
          mbar_wait(&empty[s], ((it / STG) & 1) ^ 1);
          mbar_arrive_expect_tx(&full[s], STAGE_BYTES);
          if (pfx_G > 0) {
            const int g = kb / pfx_KT;
            tma_load_4d(smem + s * STAGE_BYTES, &mapA, &full[s], g * 8, t0, (kb - g * pfx_KT) * 8, bt);
          } else if (pfx_G < 0) {
            const int j = kb / pfxG;
            tma_load_4d(smem + s * STAGE_BYTES, &mapA, &full[s], (kb - j * pfxG) * 64, t0, j, bt);
          } else {
            tma_load_3d(smem + s * STAGE_BYTES, &mapA, &full[s], kb * BK, t0, bt);
          }
          tma_load_2d(smem + s * STAGE_BYTES + A_BYTES, &mapB, &full[s], kb * BK, nt * BN);
        }
        // The epilogue's skip / LeakyReLU' operands are read once, 8 bytes per lane by 8 warps (~16 KB in flight per SM): their
        // DRAM latency, not bandwidth, sets the tile time of the thin GEMMs that have them.  The producer runs STG stages ahead of
        // the MMAs: one TMA prefetch per operand and tile (the whole box of C-addressed runs) puts the lines in the L2 early.
        if (pf_flags & 3) {
          const int n0 = nt * BN, run = n0 / pf_cn;
          const int c0 = (pf_flags & 4) ? n0 - run * pf_cn : 0;
          const int rowc = (pf_flags & 8) ? bt * R + t0 : t0, batc = (pf_flags & 8) ? 0 : bt;
          if (pf_flags & 1) tma_prefetch_4d(&mapS, c0, run, rowc, batc);
          if (pf_flags & 2) tma_prefetch_4d(&mapD, c0, run, rowc, batc);
        }
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      const uint32_t idesc = umma_idesc(BM, BN, 0, 0);
      int it = 0, lt = 0;
      const bool dbgm = g_nt_dbg_on != 0;
      long long t_full = 0, t_tempty = 0;
      const long long t_begin = clock64();
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++lt) {
        const int acc = lt & 1;
        const long long q0 = dbgm ? clock64() : 0;
        mbar_wait(&tempty[acc], ((lt >> 1) & 1) ^ 1);
        if (dbgm) t_tempty += clock64() - q0;
        tc_fence_after();
        const uint32_t tacc = tmem_base + (uint32_t)(acc * ACC_COLS);
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % STG;
          const long long q1 = dbgm ? clock64() : 0;
          mbar_wait(&full[s], (it / STG) & 1);
          if (dbgm) t_full += clock64() - q1;
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + s * STAGE_BYTES);
          // channel-prefix stage: un-swizzled [tap][row][16 B] (2 taps = 4096 B per UMMA_K); else the SWIZZLE_128B tile
          const uint64_t da = pfx_G > 0 ? umma_desc_ns(sa, BM * 16, 128) : umma_desc(sa, 16, 1024), db = umma_desc(sa + A_BYTES, 16, 1024);
          const uint64_t astep = pfx_G > 0 ? (uint64_t)((2 * BM * 16) >> 4) : 2ull;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)
            tc_mma(tacc, da + (uint64_t)k * astep, db + (uint64_t)(k * 2), idesc, (kb | k) ? 1u : 0u);
          tc_commit(&empty[s]);
        }
        tc_commit(&tfull[acc]);
      }
      if (dbgm) {
        atomicAdd(&g_nt_dbg[3], (unsigned long long)t_full);
        atomicAdd(&g_nt_dbg[4], (unsigned long long)t_tempty);
        atomicAdd(&g_nt_dbg[5], (unsigned long long)(clock64() - t_begin));
        atomicAdd(&g_nt_dbg[6], (unsigned long long)it);
        atomicAdd(&g_nt_dbg[7], 1ull);
      }
    }
    __syncwarp();
  }
  __syncthreads();
  if (warp == 9) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// 3-D bf16 map {cols (contiguous), rows (stride rs elements), batches (stride bs elements)}; box {box_c, box_r, 1}.
static int make_map_3d(CUtensorMap* map, const void* ptr, int64_t cols, int64_t rows, int64_t nb, int64_t rs, int64_t bs,
                       int box_c, int box_r) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) { set_error("cuTensorMapEncodeTiled entry point not available"); return AG_ENOTSUP; }
  cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)nb};
  cuuint64_t strides[2] = {(cuuint64_t)rs * 2, (cuuint64_t)bs * 2};
  cuuint32_t box[3] = {(cuuint32_t)box_c, (cuuint32_t)box_r, 1};
  cuuint32_t es[3] = {1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (3-D) failed (%d): cols %lld rows %lld nb %lld rs %lld bs %lld", (int)r, (long long)cols,
              (long long)rows, (long long)nb, (long long)rs, (long long)bs);
    return AG_ECUDA;
  }
  return AG_OK;
}

static bool nt_vec_epilogue(const ag_gemm_desc* d);
// 4-D bf16 map of a channel-prefix im2col view {channel (contiguous), row (stride rs), tap (stride ts), batch (stride bs)}.
//   a_layout 1: box {8 channels, box_r rows, 8 taps, 1}, NO swizzle (measured: SWIZZLE_128B faults unless the box's inner
//     dimension is 128 bytes): shared memory holds [tap][row][16 bytes], i.e. per tap a column of 8-row x 16-byte core matrices --
//     the canonical un-swizzled UMMA operand (K-major for the NT kernel: LBO = box_r*16 between taps, SBO = 128 between row groups).
//   a_layout 2: box {64 channels, box_r rows, 1 tap, 1}, SWIZZLE_128B: one tap's 64-channel group of box_r rows = the same
//     [row][128 B] swizzled tile the plain 3-D map delivers (K-major block of the NT kernel, MN-major block of the TN kernel).
//     Channels past the prefix (the last group of a prefix that is not a multiple of 64) are out of the map's bounds: zero-filled
//     by the TMA unit, never read from memory.  128-byte requests instead of 16-byte ones: the request rate no longer bounds it.
static int make_map_4d(CUtensorMap* map, const void* ptr, int64_t cin, int64_t taps, int64_t rows, int64_t nb, int64_t ts, int64_t rs,
                       int64_t bs, int box_r, bool wide) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) { set_error("cuTensorMapEncodeTiled entry point not available"); return AG_ENOTSUP; }
  cuuint64_t dims[4] = {(cuuint64_t)cin, (cuuint64_t)rows, (cuuint64_t)taps, (cuuint64_t)nb};
  cuuint64_t strides[3] = {(cuuint64_t)rs * 2, (cuuint64_t)ts * 2, (cuuint64_t)bs * 2};
  cuuint32_t box[4] = {wide ? 64u : 8u, (cuuint32_t)box_r, wide ? 1u : 8u, 1};
  cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, wide ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (4-D) failed (%d): cin %lld taps %lld rows %lld nb %lld ts %lld rs %lld bs %lld", (int)r,
              (long long)cin, (long long)taps, (long long)rows, (long long)nb, (long long)ts, (long long)rs, (long long)bs);
    return AG_ECUDA;
  }
  return AG_OK;
}
// Channel-prefix views.  a_layout 1: B / dW columns ordered (channel group of 8, tap padded to 8, channel): G groups, KT tap blocks.
// a_layout 2: columns ordered (tap, channel group of 64, channel): G = ceil(prefix / 64) groups per tap, KT = taps.
struct PfxGeom { int64_t taps, G, KT, Kq; int wide; };
static bool pfx_geom(const ag_gemm_desc* d, PfxGeom* g) {
  if ((d->a_layout != 1 && d->a_layout != 2) || d->a_kin <= 0 || d->a_kin % 8 != 0 || d->K % d->a_kin != 0 || d->a_k1s % 8 != 0 || d->a_k1s <= 0)
    return false;
  g->taps = d->K / d->a_kin;
  g->wide = d->a_layout == 2;
  if (g->wide) {
    g->G = (d->a_kin + 63) / 64;
    g->KT = g->taps;
  } else {
    g->G = d->a_kin / 8;
    g->KT = (g->taps + 7) / 8;
  }
  g->Kq = g->G * g->KT * 64;
  return true;
}

// Row-tile height R (rows per batch) when the descriptor can take the TMA-fed kernel, else 0.
static int64_t tma_rows_per_batch(const ag_gemm_desc* d) {
  static int off = -1;
  if (off < 0) { const char* e = getenv("AUDIOGAN_NT"); off = (e && e[0] == 'o') ? 1 : 0; }     // AUDIOGAN_NT=old: A/B switch
  if (off) return 0;
  PfxGeom pg;
  const bool pfx = pfx_geom(d, &pg);
  if (d->a_dtype != 1 || d->b_dtype != 1 || (!pfx && d->a_kin < d->K) || !nt_vec_epilogue(d)) return 0;
  if (pfx && (d->a_rpb > d->M || d->ldb < pg.Kq)) return 0;
  if ((reinterpret_cast<uintptr_t>(d->A) & 15) != 0 || d->a_rs % 8 != 0 || d->a_rs <= 0) return 0;
  const bool a_flat = d->a_rpb >= d->M, c_flat = d->c_rpb >= d->M;
  int64_t R = d->M;
  if (!a_flat) R = d->a_rpb;
  if (!c_flat) { if (!a_flat && d->c_rpb != d->a_rpb) return 0; R = d->c_rpb; }
  if (R <= 0 || d->M % R != 0) return 0;
  if (!a_flat && d->M / R > 1 && (d->a_bs % 8 != 0 || d->a_bs <= 0)) return 0;
  if (R >= (1ll << 31) || d->M / R >= (1ll << 31)) return 0;
  // the vector epilogue works with 32-bit element offsets
  const int64_t cb = c_flat ? 1 : d->M / d->c_rpb, cr = c_flat ? d->M : d->c_rpb;
  const int64_t span = (cb - 1) * d->c_bs + (cr - 1) * d->c_rs + ((d->N - 1) / d->c_nin) * d->c_n1s + d->c_nin;
  const int64_t lim = (1ll << 31) - 1;
  if (span >= lim || d->c_bs < 0 || d->c_rs < 0 || d->c_n1s < 0 || cb * d->rowbias_ld + d->N >= lim || d->rowbias_ld < 0) return 0;
  if (d->mask_len && (cr * (d->mask_tmul < 0 ? -d->mask_tmul : d->mask_tmul) + ((d->N - 1) / d->c_nin + 1) * (d->mask_n1mul < 0 ? -d->mask_n1mul : d->mask_n1mul) +
                          (d->mask_toff < 0 ? -d->mask_toff : d->mask_toff) >= lim)) return 0;
  return R;
}

static int epi_prefetch_on() {          // A/B knob: AUDIOGAN_EPI_PF=0 switches the epilogue-operand L2 prefetch off
  static int v = -1;
  if (v < 0) { const char* e = getenv("AUDIOGAN_EPI_PF"); v = (e && e[0] == '0') ? 0 : 1; }
  return v;
}
// 4-D bf16 map of an epilogue operand in C addressing {column within a run of c_nin, run (stride c_n1s), row (stride c_rs), batch
// (stride c_bs)}; box = the runs one BN-column tile touches x 128 rows.  Only used by cp.async.bulk.prefetch.tensor (no shared-memory
// destination).  Returns false when the geometry does not fit a tensor map (16-byte strides / box rows): no prefetch then.
template <int BN>
static bool make_map_epi(CUtensorMap* map, const ag_gemm_desc* d, const void* ptr, int* flags, int* cn_out) {
  EncodeTiledFn fn = encode_fn();
  if (!fn || !ptr || (reinterpret_cast<uintptr_t>(ptr) & 15) != 0) return false;
  const bool c_flat = d->c_rpb >= d->M;
  const int64_t cn = d->c_nin < d->N ? d->c_nin : d->N;
  const int64_t nruns = (d->N + cn - 1) / cn;
  const bool wide = cn >= BN;
  const int64_t box_c = wide ? BN : cn;
  int64_t box_runs = wide ? 1 : ((BN + cn - 1) / cn + ((BN % cn) ? 1 : 0));
  if (box_runs > 256) box_runs = 256;
  const int64_t rows = c_flat ? d->M : d->c_rpb, nbat = c_flat ? 1 : d->M / d->c_rpb;
  const int64_t rstr = d->c_rs, nstr = nruns > 1 ? d->c_n1s : d->c_rs, bstr = nbat > 1 ? d->c_bs : rows * d->c_rs;
  if (box_c > 256 || box_c % 8 != 0 || rstr <= 0 || rstr % 8 != 0 || nstr <= 0 || nstr % 8 != 0 || bstr <= 0 || bstr % 8 != 0) return false;
  if (rows >= (1ll << 31) || nbat >= (1ll << 31)) return false;
  cuuint64_t dims[4] = {(cuuint64_t)cn, (cuuint64_t)nruns, (cuuint64_t)rows, (cuuint64_t)nbat};
  cuuint64_t strides[3] = {(cuuint64_t)nstr * 2, (cuuint64_t)rstr * 2, (cuuint64_t)bstr * 2};
  cuuint32_t box[4] = {(cuuint32_t)box_c, (cuuint32_t)box_runs, (cuuint32_t)BM, 1};
  cuuint32_t es[4] = {1, 1, 1, 1};
  if (fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return false;
  *flags |= (wide ? 4 : 0) | (c_flat ? 8 : 0);
  *cn_out = (int)cn;
  return true;
}
template <int BN, bool VECC>
static int launch_nt_tma2(const ag_gemm_desc* d, int64_t R, cudaStream_t s) {
  CUtensorMap mapA, mapB;
  const bool a_flat = d->a_rpb >= d->M;
  const int64_t nb = d->M / R;
  PfxGeom pg = {0, 0, 0, 0, 0};
  const bool pfx = pfx_geom(d, &pg);
  int rc = pfx ? make_map_4d(&mapA, d->A, d->a_kin, pg.taps, R, nb, d->a_k1s, d->a_rs, nb == 1 ? R * d->a_rs : d->a_bs, BM, pg.wide != 0)
               : make_map_3d(&mapA, d->A, d->K, R, nb, d->a_rs, (a_flat || nb == 1) ? R * d->a_rs : d->a_bs, BK, BM);
  if (rc) return rc;
  rc = make_map_2d(&mapB, d->B, d->N, pfx ? pg.Kq : d->K, d->ldb, BK, BN);
  if (rc) return rc;
  constexpr int STG = tma_stages(BN);
  constexpr int smem = STG * (BM * BK * 2 + BN * BK * 2) + 8 * 32 * TRLD * 4 + 1024 + (2 * STG + 4) * 8 + 16 + BM * 8 + 3 * BM * 4 + BM * 16;
  static_assert(smem <= 227 * 1024, "shared-memory budget");
  auto kern = gemm_nt_tma_kernel<BN, VECC>;
  static bool attr_set = false;      // once per instantiation (a driver call per launch otherwise)
  if (!attr_set) { AG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); attr_set = true; }
  const int64_t tpb = (R + BM - 1) / BM, ntn = (d->N + BN - 1) / BN;
  const int64_t total = nb * tpb * ntn;
  AG_CHECK_ARG(total < (1ll << 31), "ag_gemm_nt_tc: too many tiles");
  // L2 prefetch of the bf16 epilogue operands (skip, LeakyReLU' operand), one TMA prefetch per tile
  CUtensorMap mapS = mapA, mapD = mapA;
  int pf_flags = 0, pf_cn = 1;
  if (epi_prefetch_on() && d->aux_dtype == 1 && (d->skip || d->dact)) {
    const bool c_flat = d->c_rpb >= d->M;
    if (c_flat || d->c_rpb == R) {                     // the row tiles are the C operand's (batch, row) tiles
      if (d->skip && make_map_epi<BN>(&mapS, d, d->skip, &pf_flags, &pf_cn)) pf_flags |= 1;
      if (d->dact && make_map_epi<BN>(&mapD, d, d->dact, &pf_flags, &pf_cn)) pf_flags |= 2;
    }
  }
  const int grid = (int)(total < sm_count() ? total : sm_count());
  kern<<<grid, TM_THREADS, smem, s>>>(*d, mapA, mapB, (int)R, (int)tpb, (int)ntn, (int)total,
                                          (int)(pg.wide ? -pg.G : pg.G), (int)pg.KT, mapS, mapD, pf_flags, pf_cn);
  AG_LAUNCH_CHECK();
  return AG_OK;
}
static bool nt_vec_epilogue(const ag_gemm_desc* d) {
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  const bool cal = d->c_dtype == 0 ? al16(d->C) : (reinterpret_cast<uintptr_t>(d->C) & 7) == 0;
  const bool auxal = d->aux_dtype == 0 ? (al16(d->skip) && al16(d->dact))
                                       : ((reinterpret_cast<uintptr_t>(d->skip) & 7) == 0 && (reinterpret_cast<uintptr_t>(d->dact) & 7) == 0);
  return d->N % 4 == 0 && d->c_nin % 4 == 0 && d->c_bs % 4 == 0 && d->c_rs % 4 == 0 && d->c_n1s % 4 == 0 && cal && auxal &&
         al16(d->bias) && (d->bias_mod == 0 || d->bias_mod % 4 == 0) && al16(d->rowbias) && d->rowbias_ld % 4 == 0;
}
template <int BN>
static int launch_nt_tma(const ag_gemm_desc* d, int64_t R, cudaStream_t s) {
  return launch_nt_tma2<BN, true>(d, R, s);       // scalar-epilogue shapes (N < 4, odd strides) stay on the 2-CTA/SM kernel
}

// ======================================================================================== TN (weight gradient)
// D[n, k] (+)= sum_m Y(m, n) * A(m, k): the reduction index m is the row index of both global operands, so both
// MMA operands are MN-major: a stage holds 64 m-rows; operand "A" = Y^T as two 64-wide n blocks, operand "B" =
// the activation window as BNK/64 k blocks, each block [64 m-rows][128 B] with the 128B swizzle.
__host__ __device__ constexpr int tn_rm(int my, int ma) { return (my == 1 && ma == 1) ? 64 : 32; }

template <int BNK, int MODEY, int MODEA>
__global__ void __launch_bounds__(TN_THREADS, 2) gemm_tn_tc_kernel(const ag_gemm_desc d, float* __restrict__ dw, int64_t ldw,
                                                                int ones_col, int64_t rows_per_split) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  // reduction rows per stage: 64 with bf16 operands, 32 otherwise (fp32 chunks cost 8 registers while in flight, and all
  // loads of a stage are issued before the first shared-memory store); same bytes in the ring either way (<= 96 KB)
  constexpr int RM = tn_rm(MODEY, MODEA);
  constexpr int STAGES = nt_stages(BNK) * (64 / RM);
  constexpr int A_BYTES = 2 * RM * 128, B_BYTES = (BNK / 64) * RM * 128, STAGE_BYTES = A_BYTES + B_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* tmem_full = empty + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);
  int64_t* yo = reinterpret_cast<int64_t*>(tmem_slot + 2);  // [STAGES][RM]
  int64_t* ao = yo + STAGES * RM;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t n0 = (int64_t)blockIdx.y * BM, k0 = (int64_t)blockIdx.x * BNK;
  const int64_t mbeg = (int64_t)blockIdx.z * rows_per_split;
  const int64_t mend = min(d.M, mbeg + rows_per_split);
  const int nst = (int)((mend - mbeg + RM - 1) / RM);
  const int64_t ones_at = ones_col ? d.K : -1;

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], TN_NPROD / 32); mbar_init(&empty[s], 1); }
    mbar_init(tmem_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  constexpr uint32_t TMEM_COLS = BNK < 32 ? 32 : BNK;
  if (warp == TN_NPROD / 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < TN_NPROD / 32) {
    // per-thread constants of the vector path: column offsets / validity of this thread's chunks (the same every stage)
    constexpr int NYc = (RM * 16) / TN_NPROD, CPRc = BNK / 8, NTOTc = (RM * CPRc) / TN_NPROD;
    int32_t ycoff[NYc], acoff[NTOTc];
    uint32_t ymask = 0, amask = 0;
    if (MODEY != 2 && MODEA != 2) {
#pragma unroll
      for (int i = 0; i < NYc; ++i) {
        const int64_t c = n0 + ((i * TN_NPROD + tid) & 15) * 8;
        const bool okc = c + 8 <= d.N;
        ycoff[i] = okc ? (int32_t)col_off(c, d.c_nin, d.c_n1s) : 0;
        ymask |= (okc ? 1u : 0u) << i;
      }
#pragma unroll
      for (int i = 0; i < NTOTc; ++i) {
        const int64_t c = k0 + ((i * TN_NPROD + tid) % CPRc) * 8;
        const bool okc = c + 8 <= d.K;
        acoff[i] = okc ? (int32_t)col_off(c, d.a_kin, d.a_k1s) : 0;
        amask |= (okc ? 1u : 0u) << i;
      }
    }
    const bool dbg = g_nt_dbg_on != 0 && tid == 0;
    long long t_empty = 0, t_off = 0, t_ld = 0, t0 = 0;
    const long long t_begin = clock64();
    auto put_offsets = [&](int it2) {
      if (tid < RM) {
        const int64_t m = mbeg + (int64_t)it2 * RM + tid;
        int64_t y = -1, a = -1;
        if (m < mend) {
          const uint32_t mu = (uint32_t)m, by = mu / (uint32_t)d.c_rpb, ba = mu / (uint32_t)d.a_rpb;
          y = (int64_t)by * d.c_bs + (int64_t)(mu - by * (uint32_t)d.c_rpb) * d.c_rs;
          a = (int64_t)ba * d.a_bs + (int64_t)(mu - ba * (uint32_t)d.a_rpb) * d.a_rs;
        }
        yo[(it2 % STAGES) * RM + tid] = y;
        ao[(it2 % STAGES) * RM + tid] = a;
      }
    };
    put_offsets(0);
    asm volatile("bar.sync 1, %0;" ::"n"(TN_NPROD) : "memory");
    for (int it = 0; it < nst; ++it) {
      const int s = it % STAGES;
      if (dbg) t0 = clock64();
      mbar_wait(&empty[s], ((it / STAGES) & 1) ^ 1);
      if (dbg) { const long long t1 = clock64(); t_empty += t1 - t0; t0 = t1; }
      // row offsets of this stage were written during the previous one (stage 0: before the loop); the ones of the NEXT
      // stage go out now, in the shadow of this stage's loads (32-bit divisions: M < 2^31 is checked on the host)
      if (it + 1 < nst) put_offsets(it + 1);
      const int64_t mr0 = mbeg + (int64_t)it * RM;
      (void)mr0;
      if (dbg) { const long long t1 = clock64(); t_off += t1 - t0; t0 = t1; }
      uint8_t* sa = smem + s * STAGE_BYTES;
      uint8_t* sb = sa + A_BYTES;
      // Y^T: 64 rows x 16 chunks (two 64-wide n blocks); activation window: 64 rows x BNK/8 chunks
      constexpr int NY = (RM * 16) / TN_NPROD;
      constexpr int CPR = BNK / 8;                          // chunks per row of the activation window
      constexpr int NTOT = (RM * CPR) / TN_NPROD;              // chunks per thread: 4 / 8 / 16
      const bool edge = ones_at >= 0 && k0 + BNK > d.K;     // the chunk holding the ones column goes through the generic path
      if (MODEY != 2 && MODEA != 2 && !edge) {
        // ALL loads of the stage (Y and the whole activation window) are in flight before the first shared-memory store:
        // one exposed global-load latency per stage instead of three
        uint4 chy[NY], cha[NTOT];
        int64_t ro[NY], roa[NTOT];
#pragma unroll
        for (int i = 0; i < NY; ++i) ro[i] = yo[s * RM + ((i * TN_NPROD + tid) >> 4)];
#pragma unroll
        for (int i = 0; i < NTOT; ++i) roa[i] = ao[s * RM + (i * TN_NPROD + tid) / CPR];
        ChunkLoader<MODEY == 2 ? 0 : MODEY, NY> ly;
        ChunkLoader<MODEA == 2 ? 0 : MODEA, NTOT> la;
        ly.issue_pre(d.C, ro, ycoff, ymask);
        la.issue_pre(d.A, roa, acoff, amask);
        ly.get(chy);
        la.get(cha);
#pragma unroll
        for (int i = 0; i < NY; ++i) {
          const int cid = i * TN_NPROD + tid;
          const int r = cid >> 4, c = cid & 15;
          *reinterpret_cast<uint4*>(sa + (c >> 3) * (RM * 128) + r * 128 + (((c & 7) ^ (r & 7)) << 4)) = chy[i];
        }
#pragma unroll
        for (int i = 0; i < NTOT; ++i) {
          const int cid = i * TN_NPROD + tid;
          const int r = cid / CPR, c = cid % CPR;
          *reinterpret_cast<uint4*>(sb + (c >> 3) * (RM * 128) + r * 128 + (((c & 7) ^ (r & 7)) << 4)) = cha[i];
        }
      } else {
      {
        uint4 ch[NY];
        int64_t ro[NY], cc[NY];
#pragma unroll
        for (int i = 0; i < NY; ++i) {
          const int cid = i * TN_NPROD + tid;
          ro[i] = yo[s * RM + (cid >> 4)];
          cc[i] = n0 + (cid & 15) * 8;
        }
        load_chunks<MODEY, NY>(ch, d.C, d.c_dtype, ro, cc, d.N, d.c_nin, d.c_n1s, -1);
#pragma unroll
        for (int i = 0; i < NY; ++i) {
          const int cid = i * TN_NPROD + tid;
          const int r = cid >> 4, c = cid & 15;
          *reinterpret_cast<uint4*>(sa + (c >> 3) * (RM * 128) + r * 128 + (((c & 7) ^ (r & 7)) << 4)) = ch[i];
        }
      }
      constexpr int NA = NTOT < 8 ? NTOT : 8;
#pragma unroll 1
      for (int i0 = 0; i0 < NTOT; i0 += NA) {
        uint4 ch[NA];
        int64_t ro[NA], cc[NA];
#pragma unroll
        for (int i = 0; i < NA; ++i) {
          const int cid = (i0 + i) * TN_NPROD + tid;
          ro[i] = ao[s * RM + cid / CPR];
          cc[i] = k0 + (cid % CPR) * 8;
        }
        if (MODEA != 2 && edge)
          load_chunks<2, NA>(ch, d.A, d.a_dtype, ro, cc, d.K, d.a_kin, d.a_k1s, ones_at);
        else
          load_chunks<MODEA, NA>(ch, d.A, d.a_dtype, ro, cc, d.K, d.a_kin, d.a_k1s, ones_at);
#pragma unroll
        for (int i = 0; i < NA; ++i) {
          const int cid = (i0 + i) * TN_NPROD + tid;
          const int r = cid / CPR, c = cid % CPR;
          *reinterpret_cast<uint4*>(sb + (c >> 3) * (RM * 128) + r * 128 + (((c & 7) ^ (r & 7)) << 4)) = ch[i];
        }
      }
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&full[s]);
      asm volatile("bar.sync 1, %0;" ::"n"(TN_NPROD) : "memory");      // next stage's row offsets are visible to all producers
      if (dbg) t_ld += clock64() - t0;
    }
    const long long t_main = clock64();
    // epilogue: TMEM lane = n, column = k; 32x32 transposes through shared memory so that one warp instruction
    // accumulates 32 consecutive k of one weight row (coalesced red.global.add.f32)
    mbar_wait(tmem_full, 0);
    tc_fence_after();
    const int64_t ktot = d.K + (ones_col ? 1 : 0);
    float* tr = reinterpret_cast<float*>(smem) + warp * (32 * 33);
    constexpr int CH = 32;
    const int wq = warp & 3, wh = warp >> 2;           // TMEM lane quarter; warps w and w + 4 split the columns
#pragma unroll 1
    for (int c0 = wh * (BNK / 2); c0 < (wh + 1) * (BNK / 2); c0 += CH) {
      if (k0 + c0 >= ktot) break;
      uint32_t v[16];
      tc_ld16(tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)c0, v);
#pragma unroll
      for (int j = 0; j < 16; ++j) tr[lane * 33 + j] = __uint_as_float(v[j]);
      tc_ld16(tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(c0 + 16), v);
#pragma unroll
      for (int j = 0; j < 16; ++j) tr[lane * 33 + 16 + j] = __uint_as_float(v[j]);
      __syncwarp();
      const int64_t k = k0 + c0 + lane;
      if (k < ktot) {
#pragma unroll 4
        for (int r = 0; r < 32; ++r) {
          const int64_t n = n0 + wq * 32 + r;
          if (n < d.N) atomicAdd(&dw[n * ldw + k], tr[r * 33 + lane]);
        }
      }
      __syncwarp();
    }
    tc_fence_before();
    if (dbg) {
      const long long t_end = clock64();
      atomicAdd(&g_nt_dbg[8], 1ull);
      atomicAdd(&g_nt_dbg[9], (unsigned long long)t_empty);
      atomicAdd(&g_nt_dbg[10], (unsigned long long)t_off);
      atomicAdd(&g_nt_dbg[11], (unsigned long long)t_ld);
      atomicAdd(&g_nt_dbg[12], (unsigned long long)(t_main - t_begin));
      atomicAdd(&g_nt_dbg[13], (unsigned long long)(t_end - t_main));
      atomicAdd(&g_nt_dbg[14], (unsigned long long)nst);
    }
  } else {
    if (lane == 0) {
      const uint32_t idesc = umma_idesc(BM, BNK, 1, 1);
      for (int it = 0; it < nst; ++it) {
        const int s = it % STAGES;
        mbar_wait(&full[s], (it / STAGES) & 1);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + s * STAGE_BYTES);
        const uint64_t da = umma_desc(sa, RM * 128, 1024), db = umma_desc(sa + A_BYTES, RM * 128, 1024);
#pragma unroll
        for (int k = 0; k < RM / 16; ++k)          // 16 m-rows = two 8-row groups = 2048 B per UMMA_K
          tc_mma(tmem_base, da + (uint64_t)(k * 128), db + (uint64_t)(k * 128), idesc, (it | k) ? 1u : 0u);
        tc_commit(&empty[s]);
      }
      tc_commit(tmem_full);
    }
    __syncwarp();
  }
  __syncthreads();
  if (warp == TN_NPROD / 32) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// ---------------------------------------------------------------------------------------- TN, TMA-fed
// Both operands bf16 with contiguous columns (Y: plain [.., N] rows; A: one contiguous im2col window per row): a stage is
// 64 reduction rows of ONE batch, fetched as {64 columns x 64 rows} boxes of two 3-D tensor maps {column, t, batch} --
// the box IS the MN-major SWIZZLE_128B operand block, rows past the batch's end arrive as zeros and add nothing.  No
// producer warps: warp 8 issues the TMA loads, warp 9 the MMAs, warps 0-7 only run the epilogue (2 CTAs / SM, so one
// CTA's epilogue overlaps the other's main loop).  The bias gradient (the "ones" column) is a separate column-sum kernel.
constexpr int TNT_THREADS = 320;
template <int BNK>
__global__ void __launch_bounds__(TNT_THREADS, 2)
gemm_tn_tma_kernel(const __grid_constant__ CUtensorMap mapY, const __grid_constant__ CUtensorMap mapA, float* __restrict__ dw, int64_t ldw,
                   int N, int K, int spb, int total_stages, int stages_per_split, int pfx_KT) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  constexpr int RM = 64;
  constexpr int STAGES = nt_stages(BNK);
  constexpr int A_BYTES = 2 * RM * 128, B_BYTES = (BNK / 64) * RM * 128, STAGE_BYTES = A_BYTES + B_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* tmem_full = empty + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n0 = blockIdx.y * BM, k0 = blockIdx.x * BNK;
  const int sbeg = blockIdx.z * stages_per_split;
  const int send = min(total_stages, sbeg + stages_per_split);
  const int nst = send - sbeg;
  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(tmem_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  constexpr uint32_t TMEM_COLS = BNK < 32 ? 32 : BNK;
  if (warp == 9) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (warp < 8) {
    // epilogue: TMEM lane = n, column = k; 32x32 transposes through shared memory so that one warp instruction accumulates
    // 32 consecutive k of one weight row (coalesced red.global.add.f32)
    mbar_wait(tmem_full, 0);
    tc_fence_after();
    float* tr = reinterpret_cast<float*>(smem) + warp * (32 * 33);
    const int wq = warp & 3, wh = warp >> 2;
#pragma unroll 1
    for (int c0 = wh * (BNK / 2); c0 < (wh + 1) * (BNK / 2); c0 += 32) {
      if (k0 + c0 >= K) break;
      uint32_t v[16];
      tc_ld16(tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)c0, v);
#pragma unroll
      for (int j = 0; j < 16; ++j) tr[lane * 33 + j] = __uint_as_float(v[j]);
      tc_ld16(tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(c0 + 16), v);
#pragma unroll
      for (int j = 0; j < 16; ++j) tr[lane * 33 + 16 + j] = __uint_as_float(v[j]);
      __syncwarp();
      const int k = k0 + c0 + lane;
      if (k < K) {
#pragma unroll 4
        for (int r = 0; r < 32; ++r) {
          const int n = n0 + wq * 32 + r;
          if (n < N) atomicAdd(&dw[(int64_t)n * ldw + k], tr[r * 33 + lane]);
        }
      }
      __syncwarp();
    }
    tc_fence_before();
  } else if (warp == 8) {
    if (lane == 0) {
      for (int it = 0; it < nst; ++it) {
        const int s = it % STAGES;
        const int g = sbeg + it, b = g / spb, t0 = (g - b * spb) * RM;
        mbar_wait(&empty[s], ((it / STAGES) & 1) ^ 1);
        mbar_arrive_expect_tx(&full[s], STAGE_BYTES);
        uint8_t* sa = smem + s * STAGE_BYTES;
        tma_load_3d(sa, &mapY, &full[s], n0, t0, b);
        tma_load_3d(sa + RM * 128, &mapY, &full[s], n0 + 64, t0, b);
#pragma unroll
        for (int j = 0; j < BNK / 64; ++j) {
          if (pfx_KT > 0) {  // channel-prefix view: 64-column block kb = (channel group, tap block); past the end -> zero fill
            const int kb = k0 / 64 + j, g = kb / pfx_KT;
            tma_load_4d(sa + A_BYTES + j * (RM * 128), &mapA, &full[s], g * 8, t0, (kb - g * pfx_KT) * 8, b);
          } else if (pfx_KT < 0) {   // a_layout 2 (-pfx_KT groups of 64 channels per tap): block kb = (tap, channel group); a tap past
            const int kb = k0 / 64 + j, tap = kb / (-pfx_KT);                          // the last one is out of bounds -> zero fill
            tma_load_4d(sa + A_BYTES + j * (RM * 128), &mapA, &full[s], (kb + tap * pfx_KT) * 64, t0, tap, b);
          } else {
            tma_load_3d(sa + A_BYTES + j * (RM * 128), &mapA, &full[s], k0 + j * 64, t0, b);
          }
        }
      }
    }
    __syncwarp();
  } else {
    if (lane == 0) {
      const uint32_t idesc = umma_idesc(BM, BNK, 1, 1);
      for (int it = 0; it < nst; ++it) {
        const int s = it % STAGES;
        mbar_wait(&full[s], (it / STAGES) & 1);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + s * STAGE_BYTES);
        // channel-prefix stage of the activation operand: un-swizzled [tap][row][16 B] = MN-major core-matrix columns
        // (LBO = 128 B between 8-row groups along the reduction, SBO = RM*16 B between 8-column groups)
        const uint64_t da = umma_desc(sa, RM * 128, 1024);
        const uint64_t db = pfx_KT > 0 ? umma_desc_ns(sa + A_BYTES, 128, RM * 16) : umma_desc(sa + A_BYTES, RM * 128, 1024);
        const uint64_t bstep = pfx_KT > 0 ? 16ull : 128ull;           // 16 reduction rows: 2 x 128 B, or 2048 B in the swizzled block
#pragma unroll
        for (int k = 0; k < RM / 16; ++k)
          tc_mma(tmem_base, da + (uint64_t)(k * 128), db + (uint64_t)k * bstep, idesc, (it | k) ? 1u : 0u);
        tc_commit(&empty[s]);
      }
      tc_commit(tmem_full);
    }
    __syncwarp();
  }
  __syncthreads();
  if (warp == 9) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

// dw[n*ldw] += sum_m Y(m, n): the bias gradient beside the TMA-fed weight-gradient kernel.  Thread = 8 columns (16 bytes of
// bf16); a block covers NCG column groups x 256/NCG row lanes (NCG = power of two >= N/8, at most 32: narrow Y matrices -- the
// generator's 32..128 hidden channels -- keep all 256 threads loading); grid.y splits the rows.
__global__ void __launch_bounds__(256) tn_bias_kernel(const __nv_bfloat16* __restrict__ Y, int64_t rpb, int64_t bs, int64_t rs, int64_t M,
                                                      int N, float* __restrict__ out, int64_t ldw, int64_t rows_per, int ncg_log2) {
  __shared__ float red[256][9];
  const int ncg = 1 << ncg_log2, nrl = 256 >> ncg_log2;
  const int cx = threadIdx.x & (ncg - 1), ry = threadIdx.x >> ncg_log2;
  const int n = (blockIdx.x * ncg + cx) * 8;
  const int64_t m0 = (int64_t)blockIdx.y * rows_per, m1 = min(M, m0 + rows_per);
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (n < N) {
    const uint32_t rpb32 = (uint32_t)rpb;
    for (int64_t m = m0 + ry; m < m1; m += 4 * nrl) {           // four rows in flight per thread
      uint4 u[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int64_t mi = m + (int64_t)i * nrl;
        const uint32_t bi = (uint32_t)mi / rpb32;
        u[i] = mi < m1 ? __ldg(reinterpret_cast<const uint4*>(Y + bi * bs + (mi - (int64_t)bi * rpb) * rs + n)) : make_uint4(0u, 0u, 0u, 0u);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const uint32_t w[4] = {u[i].x, u[i].y, u[i].z, u[i].w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          acc[2 * e] += __uint_as_float(w[e] << 16);
          acc[2 * e + 1] += __uint_as_float(w[e] & 0xffff0000u);
        }
      }
    }
  }
  // row lanes with the same column group: shuffles inside a warp (cx = lane & (ncg - 1)), then the 8 warps through shared memory
  // (a serial 256/ncg-term sum per value by ncg threads took longer than the loads: 8 us of a 16 us launch at N = 32)
  const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    float v = acc[e];
    for (int o = ncg; o < 32; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane < ncg) red[wp * 32 + lane][e] = v;
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < 8 * ncg; idx += 256) {
    const int c = idx >> 3, e = idx & 7, nn = (blockIdx.x * ncg + c) * 8;
    if (nn < N) {
      float v = 0.f;
#pragma unroll
      for (int q = 0; q < 8; ++q) v += red[q * 32 + c][e];
      atomicAdd(&out[(int64_t)(nn + e) * ldw], v);
    }
  }
}

static int tn_split_factor() {          // experiment knob: CTAs per SM worth of row splits (AUDIOGAN_TN_SPLIT, default 4)
  static int v = 0;
  if (!v) { const char* e = getenv("AUDIOGAN_TN_SPLIT"); v = e ? atoi(e) : 4; if (v < 1) v = 4; }
  return v;
}
template <int BNK>
static int launch_tn(const ag_gemm_desc* d, float* dw, int64_t ldw, int ones_col, bool vy, bool va, cudaStream_t s) {
  constexpr int RM = 64;                                   // row granularity of the splits (both stage depths divide it)
  // ring bytes do not depend on the stage depth (32 rows x 2S stages = 64 rows x S stages); barriers / offsets sized for 2S
  constexpr int STAGES = 2 * nt_stages(BNK);
  constexpr int smem = nt_stages(BNK) * (2 * 64 * 128 + (BNK / 64) * 64 * 128) + 1024 + (2 * STAGES + 1) * 8 + 16 + 2 * STAGES * RM * 8;
  const int64_t ktot = d->K + (ones_col ? 1 : 0);
  const int64_t gx = (ktot + BNK - 1) / BNK, gy = (d->N + BM - 1) / BM;
  int64_t want = (int64_t)sm_count() * tn_split_factor() / (gx * gy);
  if (want < 1) want = 1;
  int64_t rows = (d->M + want - 1) / want;
  if (rows < 512) rows = 512;
  rows = (rows + RM - 1) / RM * RM;
  const int64_t gz = (d->M + rows - 1) / rows;
  AG_CHECK_ARG(gy < 65536 && gz < 65536, "ag_gemm_tn_tc: grid too large");
  dim3 grid((unsigned)gx, (unsigned)gy, (unsigned)gz);
#define AG_TN_LAUNCH(MY, MA)                                                                               \
  do {                                                                                                     \
    auto kern = gemm_tn_tc_kernel<BNK, MY, MA>;                                                             \
    AG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));               \
    kern<<<grid, TN_THREADS, smem, s>>>(*d, dw, ldw, ones_col, rows);                                        \
  } while (0)
  const int my = vy ? (d->c_dtype == 0 ? 0 : 1) : 2, ma = va ? (d->a_dtype == 0 ? 0 : 1) : 2;
  if (my == 0 && ma == 0) AG_TN_LAUNCH(0, 0);
  else if (my == 1 && ma == 1) AG_TN_LAUNCH(1, 1);
  else if (my == 0 && ma == 2) AG_TN_LAUNCH(0, 2);
  else if (my == 2 && ma == 0) AG_TN_LAUNCH(2, 0);
  else if (my == 0 && ma == 1) AG_TN_LAUNCH(0, 1);
  else if (my == 1 && ma == 0) AG_TN_LAUNCH(1, 0);
  else AG_TN_LAUNCH(2, 2);
#undef AG_TN_LAUNCH
  AG_LAUNCH_CHECK();
  return AG_OK;
}

// Rows per batch R when the weight-gradient descriptor can take the TMA-fed kernel, else 0.
static int64_t tn_tma_rows_per_batch(const ag_gemm_desc* d) {
  static int off = -1;
  if (off < 0) { const char* e = getenv("AUDIOGAN_TN"); off = (e && e[0] == 'o') ? 1 : 0; }     // AUDIOGAN_TN=old: A/B switch
  if (off) return 0;
  PfxGeom pg;
  const bool pfx = pfx_geom(d, &pg);
  if (d->a_dtype != 1 || d->c_dtype != 1 || (!pfx && d->a_kin < d->K) || d->c_nin < d->N) return 0;
  if (pfx && d->a_rpb > d->M) return 0;
  if (((reinterpret_cast<uintptr_t>(d->A) | reinterpret_cast<uintptr_t>(d->C)) & 15) != 0) return 0;
  if (d->a_rs % 8 != 0 || d->a_rs <= 0 || d->c_rs % 8 != 0 || d->c_rs <= 0 || d->N % 8 != 0) return 0;
  const bool a_flat = d->a_rpb >= d->M, y_flat = d->c_rpb >= d->M;
  int64_t R = d->M;
  if (!a_flat) R = d->a_rpb;
  if (!y_flat) { if (!a_flat && d->c_rpb != d->a_rpb) return 0; R = d->c_rpb; }
  if (R <= 0 || d->M % R != 0) return 0;
  const int64_t nb = d->M / R;
  if (nb > 1 && ((!a_flat && (d->a_bs % 8 != 0 || d->a_bs <= 0)) || (!y_flat && (d->c_bs % 8 != 0 || d->c_bs <= 0)))) return 0;
  if (R >= (1ll << 31) || nb >= (1ll << 31) || d->N >= (1ll << 31) || d->K >= (1ll << 31)) return 0;
  return R;
}

template <int BNK>
static int launch_tn_tma(const ag_gemm_desc* d, float* dw, int64_t ldw, int ones_col, int64_t R, cudaStream_t s) {
  CUtensorMap mapY, mapA;
  const bool a_flat = d->a_rpb >= d->M, y_flat = d->c_rpb >= d->M;
  const int64_t nb = d->M / R;
  int rc = make_map_3d(&mapY, d->C, d->N, R, nb, d->c_rs, (y_flat || nb == 1) ? R * d->c_rs : d->c_bs, 64, 64);
  if (rc) return rc;
  PfxGeom pg = {0, 0, 0, 0, 0};
  const bool pfx = pfx_geom(d, &pg);
  const int64_t Kd = pfx ? pg.Kq : d->K;                  // columns of dw the kernel produces
  AG_CHECK_ARG(ldw >= Kd + (ones_col ? 1 : 0), "ag_gemm_tn_tc: bad ldw for the channel-prefix layout");
  rc = pfx ? make_map_4d(&mapA, d->A, d->a_kin, pg.taps, R, nb, d->a_k1s, d->a_rs, nb == 1 ? R * d->a_rs : d->a_bs, 64, pg.wide != 0)
           : make_map_3d(&mapA, d->A, d->K, R, nb, d->a_rs, (a_flat || nb == 1) ? R * d->a_rs : d->a_bs, 64, 64);
  if (rc) return rc;
  constexpr int STAGES = nt_stages(BNK);
  constexpr int smem = STAGES * (2 * 64 * 128 + (BNK / 64) * 64 * 128) + 1024 + (2 * STAGES + 1) * 8 + 16;
  static_assert(STAGES * (2 * 64 * 128 + (BNK / 64) * 64 * 128) >= 8 * 32 * 33 * 4, "epilogue transpose tiles must fit in the stages");
  const int64_t spb = (R + 63) / 64, total = nb * spb;
  const int64_t gx = (Kd + BNK - 1) / BNK, gy = (d->N + BM - 1) / BM;
  int64_t want = (int64_t)sm_count() * tn_split_factor() / (gx * gy);
  if (want < 1) want = 1;
  int64_t sps = (total + want - 1) / want;
  if (sps < 8) sps = 8;
  const int64_t gz = (total + sps - 1) / sps;
  AG_CHECK_ARG(gy < 65536 && gz < 65536 && total < (1ll << 31), "ag_gemm_tn_tc: grid too large");
  auto kern = gemm_tn_tma_kernel<BNK>;
  static bool attr_set = false;      // once per instantiation (a driver call per launch otherwise)
  if (!attr_set) { AG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); attr_set = true; }
  kern<<<dim3((unsigned)gx, (unsigned)gy, (unsigned)gz), TNT_THREADS, smem, s>>>(mapY, mapA, dw, ldw, (int)d->N, (int)Kd, (int)spb,
                                                                                (int)total, (int)sps, (int)(pg.wide ? -pg.G : pg.KT));
  AG_LAUNCH_CHECK();
  if (ones_col) {
    int lg = 0;
    while ((1 << lg) < d->N / 8 && lg < 5) ++lg;
    const int64_t bx = (d->N / 8 + (1 << lg) - 1) >> lg;
    int64_t by = (int64_t)sm_count() * 8 / bx;
    if (by < 1) by = 1;
    int64_t rows_per = (d->M + by - 1) / by;
    if (rows_per < 128) rows_per = 128;
    by = (d->M + rows_per - 1) / rows_per;
    tn_bias_kernel<<<dim3((unsigned)bx, (unsigned)by), 256, 0, s>>>(reinterpret_cast<const __nv_bfloat16*>(d->C), d->c_rpb, d->c_bs, d->c_rs,
                                                                    d->M, (int)d->N, dw + Kd, ldw, rows_per, lg);
    AG_LAUNCH_CHECK();
  }
  return AG_OK;
}

}  // namespace tc
}  // namespace ag

using namespace ag;
extern "C" {

int ag_gemm_nt_tc(const ag_gemm_desc* d, void* stream) {
  AG_CHECK_ARG(d && d->M > 0 && d->N > 0 && d->K > 0 && d->A && d->B && d->C, "ag_gemm_nt_tc: bad descriptor");
  AG_CHECK_ARG(d->b_dtype == 1, "ag_gemm_nt_tc: B must be bf16 (packed weights)");
  AG_CHECK_ARG(d->ldb % 8 == 0 && (reinterpret_cast<uintptr_t>(d->B) & 15) == 0,
               "ag_gemm_nt_tc: B rows must be 16-byte aligned (ldb %lld)", (long long)d->ldb);
  AG_CHECK_ARG(d->a_rpb > 0 && d->a_kin > 0 && d->c_rpb > 0 && d->c_nin > 0, "ag_gemm_nt_tc: bad view fields");
  AG_CHECK_ARG((d->M + tc::BM - 1) / tc::BM < 2147483647LL && (d->N + 15) / 16 < 65536, "ag_gemm_nt_tc: grid too large");
  cudaStream_t s = (cudaStream_t)stream;
  const int al = d->a_dtype == 0 ? 4 : 8;       // elements per 16 bytes
  const bool vec = d->a_kin % 8 == 0 && d->K % 8 == 0 && d->a_bs % al == 0 && d->a_rs % al == 0 && d->a_k1s % al == 0 &&
                   (reinterpret_cast<uintptr_t>(d->A) & 15) == 0;
  const int64_t R = vec ? tc::tma_rows_per_batch(d) : 0;
  AG_CHECK_ARG(d->a_layout == 0 || ((d->a_layout == 1 || d->a_layout == 2) && R > 0),
               "ag_gemm_nt_tc: the channel-prefix layouts (a_layout 1, 2) need bf16 operands, 16-byte aligned strides and a vector epilogue");
#define AG_TC_NT(BN)                                                                                         \
  return R > 0 ? tc::launch_nt_tma<BN>(d, R, s)                                                              \
               : (!vec ? tc::launch_nt<BN, 2>(d, s) : (d->a_dtype == 0 ? tc::launch_nt<BN, 0>(d, s) : tc::launch_nt<BN, 1>(d, s)))
  // Column-tile width: the widest tile N allows -- unless that leaves most SMs without a tile.  Skinny products (a handful of
  // 128-row tiles: the per-frame GEMMs of the step-wise recurrence, M = batch) are bound by how fast the weights stream from
  // L2, i.e. by the number of CTAs pulling: narrow the tile until ~2/3 of the SMs have one.
  int bn = d->N > 128 ? 256 : d->N > 64 ? 128 : d->N > 32 ? 64 : d->N > 16 ? 32 : 16;
  {
    const int64_t rpb = R > 0 ? R : (d->a_rpb < d->M ? d->a_rpb : d->M);
    const int64_t row_tiles = (d->M / (rpb > 0 ? rpb : 1)) * ((rpb + tc::BM - 1) / tc::BM);
    const int64_t want = (int64_t)sm_count() * 2 / 3;
    while (bn > 16 && row_tiles * ((d->N + bn - 1) / bn) < want && row_tiles <= 8) bn >>= 1;
  }
  // N = 384 (the widest generator data gradient, 4 x 96 prefix channels): two 256-column tiles leave the second half empty and cost
  // its MMAs and B loads anyway; three 128-column tiles are exact (AUDIOGAN_BN_EXACT=0: A/B knob)
  {
    static int exact = -1;
    if (exact < 0) { const char* e = getenv("AUDIOGAN_BN_EXACT"); exact = (e && e[0] == '0') ? 0 : 1; }
    if (exact && bn == 256 && d->N > 256 && d->N <= 512 && d->N % 256 != 0 && d->N % 256 <= 128 && d->N % 128 == 0) bn = 128;
  }
  if (bn == 256) { AG_TC_NT(256); }
  if (bn == 128) { AG_TC_NT(128); }
  if (bn == 64) { AG_TC_NT(64); }
  if (bn == 32) { AG_TC_NT(32); }
  AG_TC_NT(16);
#undef AG_TC_NT
}
int ag_gemm_tn_tc(const ag_gemm_desc* d, float* dw, int64_t ldw, int32_t ones_col, void* stream) {
  AG_CHECK_ARG(d && d->M > 0 && d->N > 0 && d->K > 0 && d->A && d->C && dw, "ag_gemm_tn_tc: bad descriptor");
  AG_CHECK_ARG(d->M < (1ll << 31) && d->a_rpb < (1ll << 31) && d->c_rpb < (1ll << 31), "ag_gemm_tn_tc: M / rows_per_batch must be < 2^31");
  AG_CHECK_ARG(d->a_rpb > 0 && d->a_kin > 0 && d->c_rpb > 0 && d->c_nin > 0, "ag_gemm_tn_tc: bad view fields");
  AG_CHECK_ARG(ldw >= d->K + (ones_col ? 1 : 0), "ag_gemm_tn_tc: bad ldw");
  cudaStream_t s = (cudaStream_t)stream;
  const int ala = d->a_dtype == 0 ? 4 : 8, aly = d->c_dtype == 0 ? 4 : 8;
  const bool va = d->a_kin % 8 == 0 && d->a_bs % ala == 0 && d->a_rs % ala == 0 && d->a_k1s % ala == 0 &&
                  (reinterpret_cast<uintptr_t>(d->A) & 15) == 0;
  const bool vy = d->c_nin % 8 == 0 && d->c_bs % aly == 0 && d->c_rs % aly == 0 && d->c_n1s % aly == 0 &&
                  (reinterpret_cast<uintptr_t>(d->C) & 15) == 0;
  const int64_t Rt = (vy && va) ? tc::tn_tma_rows_per_batch(d) : 0;
  if (Rt > 0) {
    tc::PfxGeom pg;
    const int64_t Kd = tc::pfx_geom(d, &pg) ? pg.Kq : d->K;
    if (Kd > 128) return tc::launch_tn_tma<256>(d, dw, ldw, ones_col, Rt, s);
    if (Kd > 64) return tc::launch_tn_tma<128>(d, dw, ldw, ones_col, Rt, s);
    return tc::launch_tn_tma<64>(d, dw, ldw, ones_col, Rt, s);
  }
  AG_CHECK_ARG(d->a_layout == 0, "ag_gemm_tn_tc: the channel-prefix layouts (a_layout 1, 2) need bf16 operands with 16-byte aligned strides");
  const int64_t ktot = d->K + (ones_col ? 1 : 0);
  if (ktot > 128) return tc::launch_tn<256>(d, dw, ldw, ones_col, vy, va, s);
  if (ktot > 64) return tc::launch_tn<128>(d, dw, ldw, ones_col, vy, va, s);
  return tc::launch_tn<64>(d, dw, ldw, ones_col, vy, va, s);
}
}

// profiling aid: enable/reset (on != 0) the NT kernel's phase counters, read them back (16 values)
extern "C" int ag_gemm_dbg_enable(int on) {
  unsigned long long z[16] = {0};
  AG_CUDA(cudaMemcpyToSymbol(ag::tc::g_nt_dbg, z, sizeof(z)));
  AG_CUDA(cudaMemcpyToSymbol(ag::tc::g_nt_dbg_on, &on, sizeof(int)));
  return AG_OK;
}
extern "C" int ag_gemm_dbg_read(unsigned long long* out) {
  AG_CUDA(cudaDeviceSynchronize());
  AG_CUDA(cudaMemcpyFromSymbol(out, ag::tc::g_nt_dbg, 16 * sizeof(unsigned long long)));
  return AG_OK;
}
