// fp32 FFMA view-GEMM ("fp32 mode": <=1e-5 parity with the reference's fp32 path).
//
//   NT: C(m,n) = epilogue( sum_k A(m,k) * B(n,k) )          forward + data-gradient
//   TN: dW[n,k] += sum_m Y(m,n) * A(m,k)                    weight gradient (+ bias column)
//
// Operands are strided *views* (see include/audiogan_b200.h): the im2col matrix of a
// channel-last activation is addressed in place, so Conv1d / ConvTranspose1d / Linear of
// audiogan.py:256-283, :465-467, :531-536, :547-549 all run through these two kernels.
// CUDA-core roofline: 148 SMs x 128 FFMA/clk.  The tcgen05 path (gemm_tc.cu) is the bf16 mode.
#include "common.cuh"

namespace ag {

constexpr int BM = 128, BK = 16, NT_THREADS = 256;

__device__ __forceinline__ int64_t c_col_off(const ag_gemm_desc& d, int64_t n, int64_t* n1_out) {
  const int64_t n1 = n / d.c_nin;
  if (n1_out) *n1_out = n1;
  return n1 * d.c_n1s + (n - n1 * d.c_nin);
}

template <int BN>
__global__ void __launch_bounds__(NT_THREADS) gemm_nt_kernel(const ag_gemm_desc d) {
  constexpr int TN = BN / 16;
  constexpr int NB_LD = (BN * BK) / NT_THREADS;   // B elements per thread per k-tile
  __shared__ __align__(16) float As[2][BK][BM + 4];
  __shared__ __align__(16) float Bs[2][BK][BN + 4];
  __shared__ int64_t rowoffA[BM];

  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int64_t m0 = (int64_t)blockIdx.y * BM, n0 = (int64_t)blockIdx.x * BN;

  if (tid < BM) {
    const int64_t m = m0 + tid;
    int64_t off = -1;
    if (m < d.M) {
      const int64_t b = m / d.a_rpb;
      off = b * d.a_bs + (m - b * d.a_rpb) * d.a_rs;
    }
    rowoffA[tid] = off;
  }
  __syncthreads();

  const int kk_ld = tid & 15;       // this thread's k lane inside a k-tile (fixed)
  const int r_ld = tid >> 4;        // first row/col it loads; then +16 per step
  float ra[8], rb[NB_LD > 0 ? NB_LD : 1];
  float acc[8][TN];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  auto load_tile = [&](int64_t k0) {
    const int64_t k = k0 + kk_ld;
    const bool kv = k < d.K;
    int64_t koff = 0;
    if (kv) {
      const int64_t k1 = k / d.a_kin;
      koff = k1 * d.a_k1s + (k - k1 * d.a_kin);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int64_t ro = rowoffA[r_ld + 16 * i];
      ra[i] = (kv && ro >= 0) ? ld_any(d.A, ro + koff, d.a_dtype) : 0.f;
    }
#pragma unroll
    for (int i = 0; i < NB_LD; ++i) {
      const int64_t n = n0 + r_ld + 16 * i;
      rb[i] = (kv && n < d.N) ? ld_any(d.B, n * d.ldb + k, d.b_dtype) : 0.f;
    }
  };
  auto store_tile = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 8; ++i) As[buf][kk_ld][r_ld + 16 * i] = ra[i];
#pragma unroll
    for (int i = 0; i < NB_LD; ++i) Bs[buf][kk_ld][r_ld + 16 * i] = rb[i];
  };

  const int64_t nk = (d.K + BK - 1) / BK;
  load_tile(0);
  store_tile(0);
  __syncthreads();
  for (int64_t kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) load_tile((kt + 1) * BK);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[8], b[TN];
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 8 + 4]);
      a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w;
      a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
#pragma unroll
      for (int j = 0; j < TN; ++j) b[j] = Bs[buf][kk][tx * TN + j];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (kt + 1 < nk) store_tile(buf ^ 1);
    __syncthreads();
  }

  // ---------------- epilogue
  const float alpha = d.alpha == 0.f ? 1.f : d.alpha;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int64_t m = m0 + ty * 8 + i;
    if (m >= d.M) continue;
    const int64_t b = m / d.c_rpb, t = m - b * d.c_rpb;
    const int64_t crow = b * d.c_bs + t * d.c_rs;
    const int mlen = d.mask_len ? d.mask_len[b] : 0;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int64_t n = n0 + tx * TN + j;
      if (n >= d.N) continue;
      int64_t n1;
      const int64_t ci = crow + c_col_off(d, n, &n1);
      float v = acc[i][j] * alpha;
      if (d.bias) v += d.bias[d.bias_mod > 0 ? n % d.bias_mod : n];
      if (d.rowbias) v += d.rowbias[b * d.rowbias_ld + n];
      if (d.skip) v += ld_any(d.skip, ci, d.aux_dtype);
      if (d.act == 1) v = v > 0.f ? v : v * d.slope;
      if (d.dact) v *= (ld_any(d.dact, ci, d.aux_dtype) > 0.f) ? 1.f : d.slope;
      if (d.mask_len) {
        const int64_t pos = t * d.mask_tmul + n1 * d.mask_n1mul + d.mask_toff;
        if (pos < 0 || pos >= mlen) v = 0.f;
      }
      st_any(d.C, ci, v, d.c_dtype);
    }
  }
}

// ------------------------------------------------------------------------------------------
// TN: dW[n, k] += sum_m Y(m,n) A(m,k).  64x64 output tile, 16-row reduction chunks, split over m.
constexpr int TO = 64, TR = 16;

__global__ void __launch_bounds__(256) gemm_tn_kernel(const ag_gemm_desc d, float* __restrict__ dw, int64_t ldw,
                                                      int ones_col, int64_t rows_per_split) {
  __shared__ __align__(16) float Ys[2][TR][TO + 4];
  __shared__ __align__(16) float As[2][TR][TO + 4];
  __shared__ int64_t yoff[2][TR], aoff[2][TR];

  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;     // tx -> k, ty -> n
  const int64_t n0 = (int64_t)blockIdx.y * TO, k0 = (int64_t)blockIdx.x * TO;
  const int64_t mbeg = (int64_t)blockIdx.z * rows_per_split;
  const int64_t mend = min(d.M, mbeg + rows_per_split);
  if (mbeg >= mend) return;
  const int64_t Ktot = d.K + (ones_col ? 1 : 0);

  const int c_ld = tid & 63;        // column (n or k) this thread loads
  const int r_ld = tid >> 6;        // first reduction row; +4 per step (4 steps)
  const int64_t n_ld = n0 + c_ld, k_ld = k0 + c_ld;
  const bool nv = n_ld < d.N;
  const bool kv = k_ld < d.K;
  const bool kone = ones_col && k_ld == d.K;
  int64_t ycol = 0, acol = 0;
  if (nv) ycol = c_col_off(d, n_ld, nullptr);
  if (kv) { const int64_t k1 = k_ld / d.a_kin; acol = k1 * d.a_k1s + (k_ld - k1 * d.a_kin); }

  auto rowoffs = [&](int buf, int64_t mb) {
    if (tid < TR) {
      const int64_t m = mb + tid;
      int64_t o = -1;
      if (m < mend) { const int64_t b = m / d.c_rpb; o = b * d.c_bs + (m - b * d.c_rpb) * d.c_rs; }
      yoff[buf][tid] = o;
    } else if (tid < 2 * TR) {
      const int64_t m = mb + tid - TR;
      int64_t o = -1;
      if (m < mend) { const int64_t b = m / d.a_rpb; o = b * d.a_bs + (m - b * d.a_rpb) * d.a_rs; }
      aoff[buf][tid - TR] = o;
    }
  };
  float ry[4], ra[4];
  auto load_tile = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = r_ld + 4 * i;
      const int64_t yo = yoff[buf][r], ao = aoff[buf][r];
      ry[i] = (nv && yo >= 0) ? ld_any(d.C, yo + ycol, d.c_dtype) : 0.f;
      ra[i] = (ao >= 0) ? (kv ? ld_any(d.A, ao + acol, d.a_dtype) : (kone ? 1.f : 0.f)) : 0.f;
    }
  };
  auto store_tile = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 4; ++i) { Ys[buf][r_ld + 4 * i][c_ld] = ry[i]; As[buf][r_ld + 4 * i][c_ld] = ra[i]; }
  };

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int64_t nch = (mend - mbeg + TR - 1) / TR;
  rowoffs(0, mbeg);
  __syncthreads();
  load_tile(0);
  store_tile(0);
  if (nch > 1) rowoffs(1, mbeg + TR);
  __syncthreads();
  for (int64_t c = 0; c < nch; ++c) {
    const int buf = c & 1;
    if (c + 1 < nch) load_tile(buf ^ 1);          // row offsets for chunk c+1 were written last iteration
#pragma unroll
    for (int r = 0; r < TR; ++r) {
      const float4 y = *reinterpret_cast<const float4*>(&Ys[buf][r][ty * 4]);
      const float4 a = *reinterpret_cast<const float4*>(&As[buf][r][tx * 4]);
      const float yy[4] = {y.x, y.y, y.z, y.w}, aa[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(yy[i], aa[j], acc[i][j]);
    }
    __syncthreads();                               // everyone done with buf and with yoff/aoff[buf^1] reads
    if (c + 1 < nch) {
      store_tile(buf ^ 1);
      if (c + 2 < nch) rowoffs(buf, mbeg + (c + 2) * TR);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t n = n0 + ty * 4 + i;
    if (n >= d.N) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t k = k0 + tx * 4 + j;
      if (k < Ktot) atomicAdd(&dw[n * ldw + k], acc[i][j]);
    }
  }
}

static int check_desc(const ag_gemm_desc* d, const char* who) {
  AG_CHECK_ARG(d, "%s: null descriptor", who);
  AG_CHECK_ARG(d->M > 0 && d->N > 0 && d->K > 0, "%s: bad M,N,K = %lld,%lld,%lld", who, (long long)d->M,
               (long long)d->N, (long long)d->K);
  AG_CHECK_ARG(d->A && d->C, "%s: null operand", who);
  AG_CHECK_ARG(d->a_rpb > 0 && d->a_kin > 0 && d->c_rpb > 0 && d->c_nin > 0, "%s: bad view fields", who);
  return AG_OK;
}

}  // namespace ag

using namespace ag;
extern "C" {

int ag_gemm_nt_f32(const ag_gemm_desc* d, void* stream) {
  int rc = check_desc(d, "ag_gemm_nt_f32");
  if (rc) return rc;
  AG_CHECK_ARG(d->B && d->ldb >= d->K, "ag_gemm_nt_f32: bad B");
  const int64_t gy = (d->M + BM - 1) / BM;
  AG_CHECK_ARG(gy < 65536 * 32767LL, "ag_gemm_nt_f32: M too large");
  cudaStream_t s = (cudaStream_t)stream;
  auto launch = [&](auto kern, int bn) {
    dim3 grid((unsigned)((d->N + bn - 1) / bn), (unsigned)gy);
    kern<<<grid, NT_THREADS, 0, s>>>(*d);
  };
  AG_CHECK_ARG(gy <= 2147483647LL, "ag_gemm_nt_f32: grid too large");
  if (gy > 65535) {
    // grid.y limit: swap roles is not needed in practice (M <= 8.3M rows); guard anyway.
    AG_CHECK_ARG(false, "ag_gemm_nt_f32: M=%lld exceeds 65535 row tiles", (long long)d->M);
  }
  if (d->N > 64) launch(gemm_nt_kernel<128>, 128);
  else if (d->N > 32) launch(gemm_nt_kernel<64>, 64);
  else if (d->N > 16) launch(gemm_nt_kernel<32>, 32);
  else launch(gemm_nt_kernel<16>, 16);
  AG_LAUNCH_CHECK();
  return AG_OK;
}

int ag_gemm_tn_f32(const ag_gemm_desc* d, float* dw, int64_t ldw, int32_t ones_col, void* stream) {
  int rc = check_desc(d, "ag_gemm_tn_f32");
  if (rc) return rc;
  const int64_t Ktot = d->K + (ones_col ? 1 : 0);
  AG_CHECK_ARG(dw && ldw >= Ktot, "ag_gemm_tn_f32: bad dw");
  const int64_t gx = (Ktot + TO - 1) / TO, gy = (d->N + TO - 1) / TO;
  // split the reduction so that ~4 waves of CTAs are in flight, at least 256 rows per split
  int64_t want = (int64_t)sm_count() * 4 / (gx * gy);
  if (want < 1) want = 1;
  int64_t rows = (d->M + want - 1) / want;
  if (rows < 256) rows = 256;
  rows = (rows + TR - 1) / TR * TR;
  const int64_t gz = (d->M + rows - 1) / rows;
  AG_CHECK_ARG(gy < 65536 && gz < 65536, "ag_gemm_tn_f32: grid too large");
  dim3 grid((unsigned)gx, (unsigned)gy, (unsigned)gz);
  gemm_tn_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(*d, dw, ldw, ones_col, rows);
  AG_LAUNCH_CHECK();
  return AG_OK;
}
}
