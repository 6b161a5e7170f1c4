// microbench.cu -- shared-memory atomic throughput probes that decide the histogram kernel design.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench microbench.cu ; run on B200.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ void red_s32(uint32_t a, int v) { asm volatile("red.shared.add.s32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ int atom_s32(uint32_t a, int v) { int o; asm volatile("atom.shared.add.s32 %0, [%1], %2;" : "=r"(o) : "r"(a), "r"(v) : "memory"); return o; }

// mode 0: conflict-free rotated slots, no return; 1: conflict-free, returning + overflow check;
// mode 2: random slot (bank conflicts); 3: same address all lanes (worst case); 4: g only (1 atomic/step)
template <int MODE>
__global__ void __launch_bounds__(256, 3) probe(int iters, unsigned long long* out_cycles, int* sink) {
  extern __shared__ int s[];
  for (int e = threadIdx.x; e < 16384; e += blockDim.x) s[e] = 0;
  __syncthreads();
  const uint32_t base = (uint32_t)__cvta_generic_to_shared(s);
  const int lane = threadIdx.x & 31, rot = lane >> 1, half = lane & 1;
  uint32_t x = threadIdx.x * 2654435761u + blockIdx.x * 40503u + 12345u;
  int spill = 0;
  unsigned long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    uint32_t w[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) { x = x * 1664525u + 1013904223u; w[k] = x ^ (x >> 15); }
    int g = (int)(x >> 20) - 2048, h = (int)(x >> 22);
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      uint32_t bin = (w[j >> 2] >> (8 * (j & 3))) & 0xffu;
      uint32_t slot;
      if (MODE == 2) slot = (w[(j + 1) & 3] >> (5 * (j & 3))) & 31u;
      else if (MODE == 3) { slot = 0; bin = 7; }
      else slot = (uint32_t)half * 16u + ((uint32_t)(j + rot) & 15u);
      uint32_t a = base + bin * 128u + slot * 4u;
      if (MODE == 1) {
        int og = atom_s32(a, g); int oh = atom_s32(a + 32768u, h);
        if ((uint32_t)(og + g + (1 << 29)) >= (1u << 30) || (uint32_t)(oh + h + (1 << 29)) >= (1u << 30)) spill++;
      } else if (MODE == 4) {
        red_s32(a, g);
      } else {
        red_s32(a, g); red_s32(a + 32768u, h);
      }
    }
  }
  __syncthreads();
  unsigned long long t1 = clock64();
  if (threadIdx.x == 0) out_cycles[blockIdx.x] = t1 - t0;
  if (spill == 123456789) sink[0] = spill + s[threadIdx.x];
}

template <int MODE>
void run(const char* name, int ctas_per_sm, int threads, int iters) {
  int dev = 0; cudaDeviceProp prop; cudaGetDeviceProperties(&prop, dev);
  int grid = prop.multiProcessorCount * ctas_per_sm;
  unsigned long long* d_cyc; int* d_sink;
  cudaMalloc(&d_cyc, grid * sizeof(unsigned long long)); cudaMalloc(&d_sink, 4);
  cudaFuncSetAttribute(probe<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  probe<MODE><<<grid, threads, 65536>>>(iters / 10 + 1, d_cyc, d_sink);  // warm-up
  cudaEventRecord(e0);
  probe<MODE><<<grid, threads, 65536>>>(iters, d_cyc, d_sink);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
  unsigned long long* h = new unsigned long long[grid];
  cudaMemcpy(h, d_cyc, grid * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < grid; ++i) avg += (double)h[i]; avg /= grid;
  double updates_per_cta = (double)threads * iters * 16;  // slot updates (g+h pair counts as one)
  double upc_sm = updates_per_cta * ctas_per_sm / avg;    // per SM per clock
  double total = updates_per_cta * grid;
  printf("%-34s ctas/sm=%d thr=%d: %.2f slot-updates/clk/SM (cycles %.0f), %.3f ms, %.1f G updates/s, eff clock %.0f MHz, err=%s\n",
         name, ctas_per_sm, threads, upc_sm, avg, ms, total / ms * 1e-6, avg / ms * 1e-3, cudaGetErrorString(cudaGetLastError()));
  delete[] h; cudaFree(d_cyc); cudaFree(d_sink);
}

int main() {
  for (int c = 1; c <= 3; ++c) {
    run<0>("conflict-free red g+h", c, 256, 4000);
    run<4>("conflict-free red g only", c, 256, 4000);
    run<1>("conflict-free atom(ret)+check g+h", c, 256, 4000);
    run<2>("random-slot red g+h", c, 256, 4000);
  }
  run<3>("same-address red g+h", 3, 256, 500);
  run<0>("conflict-free red g+h", 3, 512, 2000);
  run<0>("conflict-free red g+h", 2, 1024, 2000);
  return 0;
}
