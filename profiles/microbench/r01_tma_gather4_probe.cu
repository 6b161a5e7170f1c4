// tma_gather_test.cu -- probe: how does cp.async.bulk.tensor.2d tile::gather4 want its tensor map, and
// what is the shared-memory layout?  Matrix: uint8 [n_rows][128], we gather the 32-byte slice [32,64) of 4 rows.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__global__ void gather_kernel(const __grid_constant__ CUtensorMap tmap, const int* rows, uint8_t* out, int col) {
  __shared__ __align__(128) uint8_t tile[4 * 32];
  __shared__ __align__(8) uint64_t bar;
  const uint32_t bar_a = (uint32_t)__cvta_generic_to_shared(&bar), tile_a = (uint32_t)__cvta_generic_to_shared(tile);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(128) : "memory");
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        ::"r"(tile_a), "l"(&tmap), "r"(bar_a), "r"(col), "r"(rows[0]), "r"(rows[1]), "r"(rows[2]), "r"(rows[3])
        : "memory");
  }
  // wait phase 0
  uint32_t ok = 0; int spins = 0;
  while (!ok && spins < 2000000) {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(ok) : "r"(bar_a), "r"(0) : "memory");
    ++spins;
  }
  if (threadIdx.x < 128) out[threadIdx.x] = ok ? tile[threadIdx.x] : 0xEE;
  if (threadIdx.x == 0) out[128] = (uint8_t)ok;
}

int main(int argc, char** argv) {
  const int box_rows = argc > 1 ? atoi(argv[1]) : 1;
  const int n = 1000;
  uint8_t* h = (uint8_t*)malloc(n * 128);
  for (int r = 0; r < n; ++r) for (int c = 0; c < 128; ++c) h[r * 128 + c] = (uint8_t)((r * 7 + c) & 0xff);
  uint8_t *d, *d_out; int* d_rows;
  cudaMalloc(&d, n * 128); cudaMemcpy(d, h, n * 128, cudaMemcpyHostToDevice);
  cudaMalloc(&d_out, 256); cudaMemset(d_out, 0xDD, 256);
  int rows[4] = {5, 900, 17, 333};
  cudaMalloc(&d_rows, 16); cudaMemcpy(d_rows, rows, 16, cudaMemcpyHostToDevice);
  EncodeFn encode = nullptr; cudaDriverEntryPointQueryResult q;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &q);
  printf("entry point: %s %p\n", cudaGetErrorString(e), (void*)encode);
  CUtensorMap tmap;
  cuuint64_t gdim[2] = {128, (cuuint64_t)n}; cuuint64_t gstride[1] = {128};
  cuuint32_t box[2] = {32, (cuuint32_t)box_rows}; cuuint32_t estr[2] = {1, 1};
  CUresult r = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode box_rows=%d -> %d\n", box_rows, (int)r);
  gather_kernel<<<1, 128>>>(tmap, d_rows, d_out, 32);
  e = cudaDeviceSynchronize();
  printf("kernel: %s\n", cudaGetErrorString(e));
  uint8_t out[256]; cudaMemcpy(out, d_out, 256, cudaMemcpyDeviceToHost);
  printf("barrier ok=%d\n", out[128]);
  int good = 1;
  for (int k = 0; k < 4; ++k) {
    int bad = 0;
    for (int c = 0; c < 32; ++c) if (out[k * 32 + c] != h[rows[k] * 128 + 32 + c]) bad++;
    printf("row slot %d (row %d): %s first bytes %d %d %d expected %d %d %d\n", k, rows[k], bad ? "MISMATCH" : "ok", out[k * 32], out[k * 32 + 1],
           out[k * 32 + 2], h[rows[k] * 128 + 32], h[rows[k] * 128 + 33], h[rows[k] * 128 + 34]);
    if (bad) good = 0;
  }
  printf("RESULT box_rows=%d %s\n", box_rows, good ? "PASS" : "FAIL");
  return 0;
}
