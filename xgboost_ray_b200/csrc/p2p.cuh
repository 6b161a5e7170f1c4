// p2p.cuh -- device-side primitives of the NVLink peer-memory exchange (B2P2P in common.cuh).
//
// Flags are monotonically increasing epochs, written with st.release.sys into the PEER's flag array after a
// __threadfence_system() and polled with ld.acquire.sys in local memory.  Peer histogram data is read with
// ld.volatile (never the read-only / L1 path: the same addresses carry new data every level).  A wait gives up after
// spin_limit polls or when the communicator was aborted, and raises *err instead of hanging the GPU.
#pragma once
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
__device__ __forceinline__ uint32_t ld_volatile_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ ulonglong2 ld_volatile_v2(const long long* p) {
  ulonglong2 v;
  asm volatile("ld.volatile.global.v2.u64 {%0,%1}, [%2];" : "=l"(v.x), "=l"(v.y) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long ld_volatile_u64(const long long* p) {
  unsigned long long v;
  asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_volatile_u64(long long* p, long long v) {
  asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// epoch this exchange of `slot` runs at (the same number on every rank)
__device__ __forceinline__ uint32_t p2p_next_epoch(const B2P2P& pp, int slot) { return ld_volatile_u32(pp.epoch + slot) + 1u; }

// threads w < world of ONE CTA: publish `epoch` of `slot` to every rank (own flag array included).  Everything this
// rank wrote before (its own kernels earlier in the stream, and the calling CTA's stores before the preceding
// __syncthreads) is visible to a peer that observes the flag.
__device__ __forceinline__ void p2p_signal(const B2P2P& pp, int slot, uint32_t epoch) {
  if ((int)threadIdx.x < pp.world) {
    __threadfence_system();
    st_release_sys(pp.flags[threadIdx.x] + slot * pp.world + pp.rank, epoch);
  }
}
// every CTA that reads peer data: wait until all ranks published `epoch` of `slot` (threads w < world poll), then barrier
__device__ __forceinline__ void p2p_wait(const B2P2P& pp, int slot, uint32_t epoch) {
  if ((int)threadIdx.x < pp.world) {
    const uint32_t* f = pp.flags[pp.rank] + slot * pp.world + threadIdx.x;
    long long spins = 0;
    while ((int32_t)(ld_acquire_sys(f) - epoch) < 0) {
      __nanosleep(100);
      if ((++spins & 1023) == 0 && (ld_volatile_u32(pp.abort_flag) != 0u || spins > pp.spin_limit)) {
        atomicExch(pp.err, 1u + (uint32_t)slot);
        break;
      }
    }
  }
  __syncthreads();
}
// single-CTA exchange kernels: record the completed epoch (thread 0, after the CTA's last use of the slot)
__device__ __forceinline__ void p2p_finish_single(const B2P2P& pp, int slot, uint32_t epoch) {
  if (threadIdx.x == 0) pp.epoch[slot] = epoch;
}
// multi-CTA exchange kernels: the last CTA to arrive records the epoch and re-arms the counter
__device__ __forceinline__ void p2p_finish_grid(const B2P2P& pp, int slot, uint32_t epoch, unsigned n_ctas) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(pp.done + slot, 1u) == n_ctas - 1u) { pp.done[slot] = 0u; pp.epoch[slot] = epoch; }
  }
}

}  // namespace b2
