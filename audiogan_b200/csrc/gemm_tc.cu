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
#include <cuda.h>
#include <mutex>
#include <unordered_map>

namespace ag {
namespace tc {

constexpr int BM = 128, BK = 64, STAGES = 4, NTHREADS = 160, NPROD = 128;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t a = smem_u32(bar);
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(a), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// UMMA shared-memory descriptor, SWIZZLE_128B, bf16.  K-major: rows of 128 B, 8-row groups SBO apart (LBO unused).
// MN-major: 64-element (128 B) MN blocks LBO apart, 8-k-row groups SBO apart.
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= 1ull << 46;          // descriptor version (Blackwell)
  d |= 2ull << 61;          // SWIZZLE_128B
  return d;
}
// kind::f16 instruction descriptor: fp32 accumulate, bf16 x bf16, M x N, per-operand major-ness.
__host__ __device__ inline uint32_t umma_idesc(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}

struct SmemLayout {
  // [STAGES][A 16 KB][B BN*128 B] then barriers
  static __host__ __device__ constexpr int a_bytes() { return BM * BK * 2; }
  static __host__ __device__ constexpr int b_bytes(int BN) { return BN * BK * 2; }
};

// 16-byte chunk (8 consecutive k) of row `roff` starting at column k -> packed bf16
template <bool VEC>
__device__ __forceinline__ uint4 load_chunk(const ag_gemm_desc& d, int64_t roff, int64_t k) {
  uint4 out = make_uint4(0u, 0u, 0u, 0u);
  if (roff < 0 || k >= d.K) return out;
  if (VEC) {
    const int64_t k1 = k / d.a_kin;
    const int64_t off = roff + k1 * d.a_k1s + (k - k1 * d.a_kin);
    if (d.a_dtype == 0) {
      const float4 lo = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(d.A) + off);
      const float4 hi = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(d.A) + off + 4);
      out.x = pack_bf16(lo.x, lo.y); out.y = pack_bf16(lo.z, lo.w);
      out.z = pack_bf16(hi.x, hi.y); out.w = pack_bf16(hi.z, hi.w);
    } else {
      out = *reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(d.A) + off);
    }
  } else {
    float v[8];
    int64_t k1 = k / d.a_kin, kr = k - k1 * d.a_kin;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      v[e] = (k + e < d.K) ? ld_any(d.A, roff + k1 * d.a_k1s + kr, d.a_dtype) : 0.f;
      if (++kr == d.a_kin) { kr = 0; ++k1; }
    }
    out.x = pack_bf16(v[0], v[1]); out.y = pack_bf16(v[2], v[3]);
    out.z = pack_bf16(v[4], v[5]); out.w = pack_bf16(v[6], v[7]);
  }
  return out;
}

template <int BN, bool VEC>
__global__ void __launch_bounds__(NTHREADS, 1) gemm_nt_tc_kernel(const ag_gemm_desc d, const __grid_constant__ CUtensorMap mapB) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // 1024-byte alignment for the 128B-swizzled tiles
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  constexpr int A_BYTES = BM * BK * 2, B_BYTES = BN * BK * 2, STAGE_BYTES = A_BYTES + B_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* tmem_full = empty + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);
  int64_t* rowoff = reinterpret_cast<int64_t*>(tmem_slot + 2);

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
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], NPROD / 32 + 1); mbar_init(&empty[s], 1); }
    mbar_init(tmem_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  constexpr uint32_t TMEM_COLS = BN < 32 ? 32 : BN;     // power of two >= 32 (BN in {16,32,64,128,256})
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 4) {
    // ------------------------------------------------------------------ producers
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb % STAGES;
      const uint32_t ph = (kb / STAGES) & 1;
      mbar_wait(&empty[s], ph ^ 1);
      uint8_t* sa = smem + s * STAGE_BYTES;
      if (tid == 0) {
        mbar_arrive_expect_tx(&full[s], B_BYTES);
        tma_load_2d(sa + A_BYTES, &mapB, &full[s], kb * BK, n0);
      }
      uint4 ch[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int cid = i * NPROD + tid;
        const int r = cid >> 3, c = cid & 7;
        ch[i] = load_chunk<VEC>(d, rowoff[r], (int64_t)kb * BK + c * 8);
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int cid = i * NPROD + tid;
        const int r = cid >> 3, c = cid & 7;
        *reinterpret_cast<uint4*>(sa + r * 128 + ((c ^ (r & 7)) << 4)) = ch[i];
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&full[s]);
    }
    // ------------------------------------------------------------------ epilogue
    mbar_wait(tmem_full, 0);
    tc_fence_after();
    const int row = warp * 32 + lane;
    const int64_t m = m0 + row;
    const bool mv = m < d.M;
    const int64_t b = mv ? m / d.c_rpb : 0, t = mv ? m - b * d.c_rpb : 0;
    const int64_t crow = b * d.c_bs + t * d.c_rs;
    const int mlen = (mv && d.mask_len) ? d.mask_len[b] : 0;
    const float alpha = d.alpha == 0.f ? 1.f : d.alpha;
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 16) {
      uint32_t v[16];
      tc_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
      if (!mv) continue;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int64_t n = n0 + c0 + j;
        if (n >= d.N) break;
        const int64_t n1 = n / d.c_nin;
        const int64_t ci = crow + n1 * d.c_n1s + (n - n1 * d.c_nin);
        float x = __uint_as_float(v[j]) * alpha;
        if (d.bias) x += d.bias[d.bias_mod > 0 ? n % d.bias_mod : n];
        if (d.rowbias) x += d.rowbias[b * d.rowbias_ld + n];
        if (d.skip) x += ld_any(d.skip, ci, d.aux_dtype);
        if (d.act == 1) x = x > 0.f ? x : x * d.slope;
        if (d.dact) x *= (ld_any(d.dact, ci, d.aux_dtype) > 0.f) ? 1.f : d.slope;
        if (d.mask_len) {
          const int64_t pos = t * d.mask_tmul + n1 * d.mask_n1mul + d.mask_toff;
          if (pos < 0 || pos >= mlen) x = 0.f;
        }
        st_any(d.C, ci, x, d.c_dtype);
      }
    }
    tc_fence_before();
  } else {
    // ------------------------------------------------------------------ MMA issuer (one thread)
    if (lane == 0) {
      const uint32_t idesc = umma_idesc(BM, BN, 0, 0);
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % STAGES;
        mbar_wait(&full[s], (kb / STAGES) & 1);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + s * STAGE_BYTES);
        const uint64_t da = umma_desc(sa, 16, 1024), db = umma_desc(sa + A_BYTES, 16, 1024);
#pragma unroll
        for (int k = 0; k < BK / 16; ++k)          // +32 B per UMMA_K inside the 128 B swizzle row
          tc_mma(tmem_base, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (kb | k) ? 1u : 0u);
        tc_commit(&empty[s]);
      }
      tc_commit(tmem_full);
    }
    __syncwarp();
  }
  __syncthreads();
  if (warp == 4) {
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

template <int BN, bool VEC>
static int launch_nt(const ag_gemm_desc* d, cudaStream_t s) {
  CUtensorMap mapB;
  int rc = make_map_2d(&mapB, d->B, d->N, d->K, d->ldb, BK, BN);
  if (rc) return rc;
  constexpr int smem = STAGES * (BM * BK * 2 + BN * BK * 2) + 1024 /*align*/ + (2 * STAGES + 1) * 8 + 16 + BM * 8;
  auto kern = gemm_nt_tc_kernel<BN, VEC>;
  AG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  dim3 grid((unsigned)((d->M + BM - 1) / BM), (unsigned)((d->N + BN - 1) / BN));
  kern<<<grid, NTHREADS, smem, s>>>(*d, mapB);
  AG_LAUNCH_CHECK();
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
#define AG_TC_NT(BN) return vec ? tc::launch_nt<BN, true>(d, s) : tc::launch_nt<BN, false>(d, s)
  if (d->N > 128) { AG_TC_NT(256); }
  if (d->N > 64) { AG_TC_NT(128); }
  if (d->N > 32) { AG_TC_NT(64); }
  if (d->N > 16) { AG_TC_NT(32); }
  AG_TC_NT(16);
#undef AG_TC_NT
}
}
