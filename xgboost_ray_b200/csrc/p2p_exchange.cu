// p2p_exchange.cu -- histogram exchange over NVLink peer memory (EXPERIMENTAL, opt-in: B2_EXCHANGE_P2P=1).
//
// Replaces, for one tree level, the pair  ncclReduceScatter(int64 histograms) ... ncclAllGather(candidates)  that
// stands in for the reference's Rabit allreduce (xgboost_ray/main.py:745-752 -> xgboost collective; SURVEY.md 8a
// row a11, 8e).  Every rank maps the build buffer, the candidate table and a small flag array of every peer
// (cudaIpcOpenMemHandle, engine.cu) and then
//   1. p2p_signal(slot HIST)     after its histogram kernel: "my partial histograms of this level are complete";
//   2. p2p_reduce_kernel         waits for all ranks' HIST flags and sums the W partial copies of the slice it owns
//                                 straight out of the peers' memory (exact int64, any order) into its level buffer;
//   3. p2p_signal(slot READ)     "I have finished reading your build buffers" (a rank waits for this before it
//                                 zeroes its build buffer for the next level);
//   4. p2p_push_cands_kernel     after the split scan: stores its candidates into every peer's candidate table,
//      p2p_signal(slot CAND), p2p_wait(slot CAND) in front of the decide kernel.
// Flags are monotonically increasing epochs written with st.release.sys and polled with ld.acquire.sys; a bounded
// spin (about ten seconds) raises an error word instead of hanging the GPU when a peer has died.
//
// Status: compiles, NOT yet validated on hardware (the round-1 GPU budget was spent); the default exchange is NCCL.
#include "common.cuh"

namespace b2 {

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ ulonglong2 ld_volatile_v2(const long long* p) {
  ulonglong2 v;
  asm volatile("ld.volatile.global.v2.u64 {%0,%1}, [%2];" : "=l"(v.x), "=l"(v.y) : "l"(p) : "memory");
  return v;
}

// thread w < world: wait until rank w has published `epoch` in `slot` of MY flag array
__device__ __forceinline__ void wait_flags(const B2P2P& pp, int slot, uint32_t epoch, uint32_t* err) {
  if ((int)threadIdx.x < pp.world) {
    const uint32_t* f = pp.flags[pp.rank] + slot * pp.world + threadIdx.x;
    long long spins = 0;
    while ((int32_t)(ld_acquire_sys(f) - epoch) < 0) {
      __nanosleep(200);
      if (++spins > 50000000LL) { atomicExch(err, 1u + (uint32_t)slot); break; }   // ~10 s: a peer is gone
    }
  }
  __syncthreads();
}

__global__ void p2p_signal_kernel(B2P2P pp, int slot, uint32_t epoch) {
  if ((int)threadIdx.x < pp.world) {
    __threadfence_system();
    st_release_sys(pp.flags[threadIdx.x] + slot * pp.world + pp.rank, epoch);
  }
}
__global__ void p2p_wait_kernel(B2P2P pp, int slot, uint32_t epoch, uint32_t* err) { wait_flags(pp, slot, epoch, err); }

// level_buf[i] = sum over ranks w of build_w[rank * shard_stride + i], i < n_elems (n_elems even: slices are 16-byte multiples)
__global__ void __launch_bounds__(256)
p2p_reduce_kernel(B2P2P pp, uint32_t epoch, long long* __restrict__ level_buf, size_t n_elems, size_t shard_stride,
                  uint32_t* err) {
  wait_flags(pp, kSlotHist, epoch, err);
  const size_t base = (size_t)pp.rank * shard_stride;
  for (size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 2; i < n_elems; i += (size_t)gridDim.x * blockDim.x * 2) {
    unsigned long long a = 0, b = 0;
    for (int w = 0; w < pp.world; ++w) {
      const ulonglong2 v = ld_volatile_v2(pp.build[w] + base + i);
      a += v.x; b += v.y;
    }
    level_buf[i] = (long long)a; level_buf[i + 1] = (long long)b;
  }
}

// my candidates [n] -> region `rank` of every rank's table (own table included)
__global__ void p2p_push_cands_kernel(B2P2P pp, const B2SplitCand* __restrict__ local, int n, int cand_cap) {
  constexpr int kWords = sizeof(B2SplitCand) / 8;
  static_assert(sizeof(B2SplitCand) % 8 == 0, "candidates are copied as 64-bit words");
  const size_t total = (size_t)pp.world * n * kWords;
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
    const int w = (int)(t / ((size_t)n * kWords));
    const size_t r = t - (size_t)w * n * kWords;
    reinterpret_cast<unsigned long long*>(pp.cands[w] + (size_t)pp.rank * cand_cap)[r] =
        reinterpret_cast<const unsigned long long*>(local)[r];
  }
}

}  // namespace b2

extern "C" {
int b2_p2p_struct_bytes() { return (int)sizeof(B2P2P); }
int b2_p2p_flag_words(int world) { return kP2PSlots * world; }
int b2_launch_p2p_signal(const void* pp, int slot, uint32_t epoch, cudaStream_t s) {
  b2::p2p_signal_kernel<<<1, 32, 0, s>>>(*reinterpret_cast<const B2P2P*>(pp), slot, epoch);
  return (int)cudaGetLastError();
}
int b2_launch_p2p_wait(const void* pp, int slot, uint32_t epoch, uint32_t* err, cudaStream_t s) {
  b2::p2p_wait_kernel<<<1, 32, 0, s>>>(*reinterpret_cast<const B2P2P*>(pp), slot, epoch, err);
  return (int)cudaGetLastError();
}
int b2_launch_p2p_reduce(const void* pp, uint32_t epoch, long long* level_buf, size_t n_elems, size_t shard_stride, uint32_t* err,
                         int num_sms, cudaStream_t s) {
  if (n_elems == 0) return 0;
  size_t want = (n_elems / 2 + 255) / 256;
  int grid = (int)(want < (size_t)num_sms * 4 ? want : (size_t)num_sms * 4);
  if (grid < 1) grid = 1;
  b2::p2p_reduce_kernel<<<grid, 256, 0, s>>>(*reinterpret_cast<const B2P2P*>(pp), epoch, level_buf, n_elems, shard_stride, err);
  return (int)cudaGetLastError();
}
int b2_launch_p2p_push_cands(const void* pp, const B2SplitCand* local, int n, int cand_cap, int num_sms, cudaStream_t s) {
  if (n <= 0) return 0;
  b2::p2p_push_cands_kernel<<<num_sms, 256, 0, s>>>(*reinterpret_cast<const B2P2P*>(pp), local, n, cand_cap);
  return (int)cudaGetLastError();
}
}
