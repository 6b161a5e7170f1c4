// microbench_loads.cu -- how fast can an SM pull row slices?  Decides the row-staging design of the
// histogram kernel.  Patterns over a [n][128 B] row-major byte matrix:
//   slice   : 2 lanes x 16 B = one 32 B sector per row (what a one-group CTA reads), 16 rows / warp load
//   line    : 8 lanes x 16 B = the whole 128 B row, 4 rows / warp load
// each with sequential rows or rows through a random permutation (gather), 1 or 3 loads in flight.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ uint4 ldg_nc_v4(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}

template <int LANES_PER_ROW, bool GATHER, int DEPTH>
__global__ void __launch_bounds__(256) probe(const uint8_t* __restrict__ bins, const int* __restrict__ ridx, long long n,
                                             int n_slices, unsigned* sink) {
  // CTA b reads slice (b % n_slices) of every row it visits; CTAs with the same b / n_slices visit the same rows
  const int slice_off = (blockIdx.x % n_slices) * LANES_PER_ROW * 16;
  constexpr int ROWS_PER_WARP = 32 / LANES_PER_ROW;
  const int lane = threadIdx.x & 31;
  const long long warp_global = ((long long)(blockIdx.x / n_slices) * blockDim.x + threadIdx.x) >> 5;
  const long long n_warps = ((long long)(gridDim.x / n_slices) * blockDim.x) >> 5;
  const int sub = lane / LANES_PER_ROW, part = lane % LANES_PER_ROW;
  unsigned acc = 0;
  for (long long base = warp_global * ROWS_PER_WARP * DEPTH; base < n; base += n_warps * ROWS_PER_WARP * DEPTH) {
    uint4 v[DEPTH];
#pragma unroll
    for (int d = 0; d < DEPTH; ++d) {
      long long r = base + d * ROWS_PER_WARP + sub;
      if (r < n) {
        long long rid = GATHER ? (long long)__ldg(ridx + r) : r;
        v[d] = ldg_nc_v4(bins + rid * 128 + slice_off + part * 16);
      } else v[d] = make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int d = 0; d < DEPTH; ++d) acc += v[d].x ^ v[d].y ^ v[d].z ^ v[d].w;
  }
  if (acc == 0x12345678u) sink[0] = acc;
}

template <int LPR, bool G, int D>
void run(const char* name, const uint8_t* bins, const int* ridx, long long n, int ctas_per_sm, int n_slices = 1) {
  int dev = 0; cudaDeviceProp prop; cudaGetDeviceProperties(&prop, dev);
  unsigned* sink; cudaMalloc(&sink, 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  int grid = prop.multiProcessorCount * ctas_per_sm; grid -= grid % n_slices;
  probe<LPR, G, D><<<grid, 256>>>(bins, ridx, n, n_slices, sink);
  cudaEventRecord(e0);
  probe<LPR, G, D><<<grid, 256>>>(bins, ridx, n, n_slices, sink);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double bytes = (double)n * LPR * 16 * n_slices;
  printf("%-28s lanes/row=%d gather=%d depth=%d ctas/sm=%d slices=%d: %.3f ms  %.0f GB/s useful, %.2f G rows/s  %s\n", name, LPR, (int)G, D,
         ctas_per_sm, n_slices, ms, bytes / ms * 1e-6, n / ms * 1e-6, cudaGetErrorString(cudaGetLastError()));
  cudaFree(sink);
}

int main() {
  const long long n = 10000000;
  uint8_t* bins; int* ridx;
  cudaMalloc(&bins, n * 128); cudaMemset(bins, 1, n * 128);
  int* h = (int*)malloc(n * sizeof(int));
  for (long long i = 0; i < n; ++i) h[i] = (int)i;
  srand(1);
  // "node at depth 3": random 1/8 subset kept in ascending order is what a deep node looks like; here a
  // full random permutation is the worst case and a stride-8 ascending list the typical one
  int* h2 = (int*)malloc(n * sizeof(int));
  for (long long i = 0; i < n; ++i) h2[i] = (int)((i * 8) % n + (i * 8) / n);
  for (long long i = n - 1; i > 0; --i) { long long j = ((long long)rand() * RAND_MAX + rand()) % (i + 1); int t = h[i]; h[i] = h[j]; h[j] = t; }
  cudaMalloc(&ridx, n * sizeof(int));
  for (int pass = 0; pass < 2; ++pass) {
    cudaMemcpy(ridx, pass == 0 ? h2 : h, n * sizeof(int), cudaMemcpyHostToDevice);
    printf("--- ridx = %s\n", pass == 0 ? "ascending stride-8" : "random permutation");
    run<2, false, 3>("slice x1 seq", bins, ridx, n, 3, 1);
    run<2, false, 3>("slice x4 seq (real pattern)", bins, ridx, n, 3, 4);
    run<2, true, 3>("slice x4 gather", bins, ridx, n, 3, 4);
    run<4, false, 3>("half x2 seq", bins, ridx, n, 3, 2);
    run<4, true, 3>("half x2 gather", bins, ridx, n, 3, 2);
    run<4, true, 3>("half x2 gather", bins, ridx, n, 1, 2);
    run<8, false, 3>("line x1 seq", bins, ridx, n, 3, 1);
    run<8, true, 3>("line x1 gather", bins, ridx, n, 3, 1);
  }
  return 0;
}
