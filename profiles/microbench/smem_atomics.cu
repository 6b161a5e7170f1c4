// smem_atomics.cu -- microbenchmark behind the histogram kernel's shared-memory atomic pattern (round 2, batch 8).
// ncu of the production kernel shows 1.27 shared-atomic wavefronts per instruction although the 32 lanes of every
// red.shared.add address 32 different banks (profiles/r02/b7_hist_traffic.json: 80.0M instructions, 101.5M wavefronts,
// 21.5M "bank conflicts" on the 10M-row root launch).  Each mode below issues the same number of atomics with a
// different address pattern; the host prints cycles per warp-level atomic per SM, ncu adds wavefronts / conflicts.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/smem_atomics profiles/microbench/smem_atomics.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void red_add(uint32_t a, int v) { asm volatile("red.shared.add.s32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t mix(uint32_t x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; }

// mode 0: production pattern (pair of groups, 4 lanes per row, g then h 128 B apart), random bins
// mode 1: same, every lane of the warp uses the SAME bin in a step
// mode 2: production, h plane on the opposite 16 banks ((a ^ 64) + 128)
// mode 3: g plane only (half the atomics)
// mode 4: both groups folded into ONE 64 KiB region (tests the 64 KiB distance between the two histograms)
// mode 5: one group, 2 lanes per row (16 rows per warp instruction)
// mode 6: production, steps issued g(j) g(j+1) h(j) h(j+1)
// mode 7: production, bins restricted to 0..15 (hot cells: same-address traffic between warps)
// mode 8: lane-private column: slot = lane, random bins (the plainest conflict-free pattern)
// mode 9: like 8 but all 32 lanes in ONE 128-byte row (same bin): no row spread at all
template <int kMode>
__global__ void __launch_bounds__(1024, 1) k(int iters, long long* cycles, int* sink) {
  extern __shared__ __align__(16) int32_t s[];
  for (int e = threadIdx.x; e < 32768; e += blockDim.x) s[e] = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int sub = lane >> 2, gsel = (lane & 3) >> 1, half = lane & 1;
  const int rot = kMode == 5 ? (lane >> 1) : sub * 2 + gsel;
  const uint32_t s0 = (uint32_t)__cvta_generic_to_shared(s);
  const uint32_t smem_g = s0 + ((kMode == 4 || kMode == 5 || kMode >= 8) ? 0u : (uint32_t)gsel * 65536u);
  const uint32_t base = smem_g + half * 64;
  uint32_t seed = mix(blockIdx.x * 1024u + threadIdx.x + 1u);
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    uint32_t w[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) { seed = seed * 1664525u + 1013904223u; w[q] = mix(seed); }
    if (kMode == 1 || kMode == 9) {
#pragma unroll
      for (int q = 0; q < 4; ++q) w[q] = __shfl_sync(0xffffffffu, w[q], 0);
    }
    if (kMode == 7) {
#pragma unroll
      for (int q = 0; q < 4; ++q) w[q] &= 0x0f0f0f0fu;
    }
    const int gx = (int)(seed & 0xffff) - 32768, hx = (int)(seed >> 20);
    if (kMode == 6) {
#pragma unroll
      for (int j = 0; j < 16; j += 2) {
        const uint32_t b0 = __byte_perm(w[j >> 2], 0u, 0x4404u | ((uint32_t)(j & 3) << 4));
        const uint32_t b1 = __byte_perm(w[j >> 2], 0u, 0x4404u | ((uint32_t)((j + 1) & 3) << 4));
        const uint32_t a0 = base + b0 + (((uint32_t)(j + rot) & 15u) << 2), a1 = base + b1 + (((uint32_t)(j + 1 + rot) & 15u) << 2);
        red_add(a0, gx); red_add(a1, gx); red_add(a0 + 128, hx); red_add(a1 + 128, hx);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const uint32_t bin256 = __byte_perm(w[j >> 2], 0u, 0x4404u | ((uint32_t)(j & 3) << 4));
        uint32_t a;
        if (kMode >= 8) a = s0 + bin256 + (uint32_t)lane * 4u;
        else a = base + bin256 + (((uint32_t)(j + rot) & 15u) << 2);
        red_add(a, gx);
        if (kMode == 2) red_add((a ^ 64u) + 128u, hx);
        else if (kMode != 3) red_add(a + 128u, hx);
      }
    }
  }
  const long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  if (s[threadIdx.x] == 0x7fffffff) sink[0] = 1;
}

template <int kMode>
void run(int iters, int sms) {
  long long* d_c; int* d_s;
  cudaMalloc(&d_c, sms * sizeof(long long)); cudaMalloc(&d_s, 4);
  cudaFuncSetAttribute(k<kMode>, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<kMode><<<sms, 1024, 131072>>>(iters / 10, d_c, d_s);
  cudaEventRecord(e0);
  k<kMode><<<sms, 1024, 131072>>>(iters, d_c, d_s);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
  long long c[256]; cudaMemcpy(c, d_c, sms * sizeof(long long), cudaMemcpyDeviceToHost);
  long long mx = 0; for (int i = 0; i < sms; ++i) mx = c[i] > mx ? c[i] : mx;
  const double per_lane = (kMode == 3 ? 16.0 : 32.0) * iters;        // atomics per lane
  const double warp_instr_per_sm = per_lane * 32;                       // 32 warps per SM
  printf("mode %d: %.3f ms, %.3f cycles per warp-level atomic per SM (%s)\n", kMode, ms, (double)mx / warp_instr_per_sm,
         cudaGetErrorString(cudaGetLastError()));
  cudaFree(d_c); cudaFree(d_s);
}

int main(int argc, char** argv) {
  int iters = argc > 1 ? atoi(argv[1]) : 4000;
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int sms = p.multiProcessorCount;
  printf("%s, %d SMs, %d iterations\n", p.name, sms, iters);
  run<0>(iters, sms); run<1>(iters, sms); run<2>(iters, sms); run<3>(iters, sms); run<4>(iters, sms);
  run<5>(iters, sms); run<6>(iters, sms); run<7>(iters, sms); run<8>(iters, sms); run<9>(iters, sms);
  return 0;
}
