// auc_kernel.cu -- binary ROC AUC of one worker's rows (the `auc` eval_metric the reference's tests use).
//
// Restates xgboost's BinaryAUC (src/metric/auc.cc: BinaryROCAUC / EvalBinary): rows are sorted by descending
// prediction, true / false positive weights are accumulated, and the area is the sum of the trapezoids between
// consecutive DISTINCT prediction values (ties form one step).  Distributed: every worker contributes its local
// (area, fp_total * tp_total) pair, the pairs are summed over the workers and AUC = sum(area) / sum(fp * tp) -- the
// allreduce that xgboost applies to `std::array<double, 2>{auc, fp * tp}`.  The sort is cub::DeviceRadixSort (library
// sort on a once-per-evaluation path); scans and the area reduction are cub primitives / a plain kernel.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "common.cuh"

namespace b2 {

__device__ __forceinline__ uint32_t desc_key(float v) {
  uint32_t u = __float_as_uint(v);
  u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);   // ascending order-preserving key
  return ~u;                                        // descending
}

__global__ void auc_keys_kernel(const float* __restrict__ pred, int64_t n, uint32_t* __restrict__ keys, int32_t* __restrict__ idx) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    keys[i] = desc_key(pred[i]); idx[i] = (int32_t)i;
  }
}
// per sorted position: positive / negative weight and "last row of its tie group" marker
__global__ void auc_weights_kernel(const uint32_t* __restrict__ keys, const int32_t* __restrict__ idx, const float* __restrict__ label,
                                   const float* __restrict__ weight, int64_t n, double* __restrict__ pos, double* __restrict__ neg,
                                   int32_t* __restrict__ end_mark) {
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x) {
    const int32_t i = idx[p];
    const double w = weight ? (double)weight[i] : 1.0, y = (double)label[i];
    pos[p] = y * w; neg[p] = (1.0 - y) * w;
    end_mark[p] = (p == n - 1 || keys[p] != keys[p + 1]) ? (int32_t)p : -1;
  }
}
__global__ void auc_area_kernel(const double* __restrict__ tp, const double* __restrict__ fp, const int32_t* __restrict__ end_mark,
                                const int32_t* __restrict__ last_end /* inclusive max-scan of end_mark */, int64_t n,
                                double* __restrict__ out /* [2]: area, fp_total * tp_total */) {
  double area = 0.0;
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x) {
    if (end_mark[p] < 0) continue;
    const int32_t prev = p > 0 ? last_end[p - 1] : -1;
    const double tp0 = prev >= 0 ? tp[prev] : 0.0, fp0 = prev >= 0 ? fp[prev] : 0.0;
    area += fabs(fp[p] - fp0) * (tp[p] + tp0) * 0.5;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) area += __shfl_xor_sync(0xffffffffu, area, o);
  if ((threadIdx.x & 31) == 0 && area != 0.0) atomicAdd(&out[0], area);
  if (blockIdx.x == 0 && threadIdx.x == 0 && n > 0) out[1] = fp[n - 1] * tp[n - 1];
}

struct MaxOp {
  __device__ __forceinline__ int32_t operator()(int32_t a, int32_t b) const { return a > b ? a : b; }
};

}  // namespace b2

extern "C" {
// bytes of scratch memory b2_auc_binary needs for n rows
size_t b2_auc_temp_bytes(int64_t n) {
  if (n < 1) n = 1;
  size_t sort_bytes = 0, scan_d = 0, scan_i = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, (const uint32_t*)nullptr, (uint32_t*)nullptr, (const int32_t*)nullptr,
                                  (int32_t*)nullptr, (int)n);
  cub::DeviceScan::InclusiveSum(nullptr, scan_d, (const double*)nullptr, (double*)nullptr, (int)n);
  cub::DeviceScan::InclusiveScan(nullptr, scan_i, (const int32_t*)nullptr, (int32_t*)nullptr, b2::MaxOp(), (int)n);
  size_t cub_bytes = sort_bytes > scan_d ? sort_bytes : scan_d;
  if (scan_i > cub_bytes) cub_bytes = scan_i;
  cub_bytes = (cub_bytes + 255) & ~(size_t)255;
  const size_t per_row = 2 * sizeof(uint32_t) + 2 * sizeof(int32_t) + 4 * sizeof(double) + 2 * sizeof(int32_t);
  return cub_bytes + (size_t)n * per_row + 4096;
}
// pred: transformed predictions [n]; out [2] doubles (device), zeroed here.  Returns a cudaError.
int b2_auc_binary(const float* pred, const float* label, const float* weight, int64_t n, void* temp, size_t temp_bytes,
                  double* out, int num_sms, cudaStream_t s) {
  cudaMemsetAsync(out, 0, 2 * sizeof(double), s);
  if (n <= 0) return (int)cudaGetLastError();
  if (temp_bytes < b2_auc_temp_bytes(n)) return (int)cudaErrorInvalidValue;
  size_t sort_bytes = 0, scan_d = 0, scan_i = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, (const uint32_t*)nullptr, (uint32_t*)nullptr, (const int32_t*)nullptr,
                                  (int32_t*)nullptr, (int)n);
  cub::DeviceScan::InclusiveSum(nullptr, scan_d, (const double*)nullptr, (double*)nullptr, (int)n);
  cub::DeviceScan::InclusiveScan(nullptr, scan_i, (const int32_t*)nullptr, (int32_t*)nullptr, b2::MaxOp(), (int)n);
  size_t cub_bytes = sort_bytes > scan_d ? sort_bytes : scan_d;
  if (scan_i > cub_bytes) cub_bytes = scan_i;
  cub_bytes = (cub_bytes + 255) & ~(size_t)255;
  char* p = (char*)temp;
  void* cub_tmp = p; p += cub_bytes;
  double* pos = (double*)p; p += (size_t)n * 8;
  double* neg = (double*)p; p += (size_t)n * 8;
  double* tp = (double*)p; p += (size_t)n * 8;
  double* fp = (double*)p; p += (size_t)n * 8;
  uint32_t* keys = (uint32_t*)p; p += (size_t)n * 4;
  uint32_t* keys_s = (uint32_t*)p; p += (size_t)n * 4;
  int32_t* idx = (int32_t*)p; p += (size_t)n * 4;
  int32_t* idx_s = (int32_t*)p; p += (size_t)n * 4;
  int32_t* end_mark = (int32_t*)p; p += (size_t)n * 4;
  int32_t* last_end = (int32_t*)p; p += (size_t)n * 4;
  int64_t g = (n + 255) / 256; if (g > (int64_t)num_sms * 16) g = (int64_t)num_sms * 16;
  b2::auc_keys_kernel<<<(int)g, 256, 0, s>>>(pred, n, keys, idx);
  size_t tb = cub_bytes;
  cub::DeviceRadixSort::SortPairs(cub_tmp, tb, keys, keys_s, idx, idx_s, (int)n, 0, 32, s);
  b2::auc_weights_kernel<<<(int)g, 256, 0, s>>>(keys_s, idx_s, label, weight, n, pos, neg, end_mark);
  tb = cub_bytes; cub::DeviceScan::InclusiveSum(cub_tmp, tb, pos, tp, (int)n, s);
  tb = cub_bytes; cub::DeviceScan::InclusiveSum(cub_tmp, tb, neg, fp, (int)n, s);
  tb = cub_bytes; cub::DeviceScan::InclusiveScan(cub_tmp, tb, end_mark, last_end, b2::MaxOp(), (int)n, s);
  b2::auc_area_kernel<<<(int)g, 256, 0, s>>>(tp, fp, end_mark, last_end, n, out);
  return (int)cudaGetLastError();
}
}
