// Data-parallel gradient all-reduce over NVLink PEER MEMORY (one process per GPU, the buffers of all ranks mapped into every
// process: torch.distributed._symmetric_memory hands out the mappings, audiogan_b200/dist.py).
//
// Replaces the bucketed ncclAllReduce between backward and the fused clip + RMSprop launches (audiogan.py has
// NN.DataParallel's gather there, :379-410).  The step's dependency chain puts that all-reduce on the critical path twice per
// step (D's 8.5 M and G's 6.3 M gradients: nothing else can run between a net's backward and its optimizer step), so what
// matters is latency: NCCL needs ~0.33 ms per 34 MB call on 2 GPUs here (launch + proxy/stream hand-off + ring protocol).
// Two-shot all-reduce written directly against peer pointers:
//   barrier (flags in peer memory)  ->  rank r sums element range r of ALL ranks' buffers (16-byte loads over NVLink) and
//   stores the sum into ALL ranks' buffers (16-byte peer stores)  ->  barrier.
// Every element crosses NVLink (N-1)/N times in each direction per rank -- the reduce-scatter + all-gather minimum -- and
// no element is read while someone writes it: range q of any rank's buffer is read only by rank q (phase 1) and written
// only by rank q.  The kernels are ordinary launches: they are captured into the step's CUDA graph.
#include "common.cuh"
#include <algorithm>

namespace ag {

__device__ __forceinline__ void st_release_sys(int32_t* p, int32_t v) {
  asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ int32_t ld_acquire_sys(const int32_t* p) {
  int32_t v;
  asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// One block, one thread per peer.  sig[r] (rank r's signal pad, mapped everywhere): slot [src] = last epoch rank `src`
// arrived at; slot [32] of the LOCAL pad = this rank's epoch counter (kept on the device so that a replayed CUDA graph,
// whose kernel arguments are frozen, still advances it).  Every rank runs the same sequence of barriers, so the counters agree.
__global__ void peer_barrier_kernel(int32_t* const* __restrict__ sig, int rank, int world) {
  __shared__ int32_t epoch;
  if (threadIdx.x == 0) {
    epoch = sig[rank][32] + 1;
    sig[rank][32] = epoch;
  }
  __syncthreads();
  __threadfence_system();
  const int p = threadIdx.x;
  if (p < world) {
    st_release_sys(sig[p] + rank, epoch);                 // "rank has arrived" into peer p's pad
    while (ld_acquire_sys(sig[rank] + p) < epoch) { }     // wait until peer p has arrived here
  }
  __syncthreads();
  __threadfence_system();
}

// n4 = number of float4 elements of the range [lo4, hi4) this rank owns; bufs[r] = rank r's gradient buffer
__global__ void __launch_bounds__(256) peer_reduce_push_kernel(float4* const* __restrict__ bufs, int rank, int world,
                                                               int64_t lo4, int64_t hi4) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = lo4 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < hi4; i += stride) {
    float4 acc = bufs[rank][i];
    for (int r = 1; r < world; ++r) {
      const int src = (rank + r) % world;                  // every rank starts on a different peer: the links share the load
      float4 v;                                            // peer data written by another GPU: never through a stale L1 line
      asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(bufs[src] + i) : "memory");
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    for (int r = 0; r < world; ++r) bufs[(rank + r) % world][i] = acc;
  }
}

}  // namespace ag

using namespace ag;
extern "C" {

int ag_peer_barrier(void* const* sig_ptrs_dev, int32_t rank, int32_t world, void* stream) {
  AG_CHECK_ARG(sig_ptrs_dev && world >= 1 && world <= 32 && rank >= 0 && rank < world, "ag_peer_barrier: bad args");
  peer_barrier_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(reinterpret_cast<int32_t* const*>(sig_ptrs_dev), rank, world);
  AG_LAUNCH_CHECK();
  return AG_OK;
}

int ag_peer_allreduce(void* const* buf_ptrs_dev, void* const* sig_ptrs_dev, int32_t rank, int32_t world, int64_t n,
                      int32_t nblocks, void* stream) {
  AG_CHECK_ARG(buf_ptrs_dev && sig_ptrs_dev && world >= 1 && world <= 32 && rank >= 0 && rank < world && n >= 0 && n % 4 == 0,
               "ag_peer_allreduce: bad args (n must be a multiple of 4 floats, buffers 16-byte aligned)");
  if (world == 1 || n == 0) return AG_OK;
  const int64_t n4 = n / 4, per = (n4 + world - 1) / world;
  const int64_t lo = std::min<int64_t>(n4, per * rank), hi = std::min<int64_t>(n4, lo + per);
  int nb = nblocks > 0 ? nblocks : 2 * sm_count();
  peer_barrier_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(reinterpret_cast<int32_t* const*>(sig_ptrs_dev), rank, world);
  if (hi > lo)
    peer_reduce_push_kernel<<<nb, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<float4* const*>(buf_ptrs_dev), rank, world, lo, hi);
  peer_barrier_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(reinterpret_cast<int32_t* const*>(sig_ptrs_dev), rank, world);
  AG_LAUNCH_CHECK();
  return AG_OK;
}

}
