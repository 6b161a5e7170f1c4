// partition_kernel.cu -- row partition (UpdatePosition), leaf sums and prediction-cache update.
//
// Replaces XGBoost's ApplySplit/UpdatePosition and UpdatePredictionCache stages reached through
// xgb.train() (xgboost_ray/main.py:745-752; Appendix A.8/A.9).
// Rows of a node live in a contiguous segment of a row-index list; a split rewrites the segment
// as [left rows | right rows] into the other (ping-pong) list.  Histogram sums are exact integers,
// so the order of rows inside a child segment is irrelevant to the model.
#include <stdlib.h>

#include "common.cuh"

namespace b2 {

constexpr int kPartThreads = 256;
constexpr int kPartChunk = 2048;  // rows per CTA work item

typedef B2SegWork SegWork;

// kCat = false: the matrix has no categorical feature; the row loop is then straight-line code whose 8 row-id loads
// and 8 bin-byte loads the compiler batches (48 registers).  With the category test in the body it software-pipelines
// only 2-3 deep and the kernel runs 1.7x slower (ncu launch lists in profiles/), so numeric matrices keep their own
// instantiation.
// kMode 0: numeric only.  1: category set in shared memory (the validated categorical path).  2 (experimental,
// B2_PART_CAT_MODE=2): category set in eight uniform registers selected with a 3-level select tree, no shared-memory
// load in the row loop.
template <int kMode>
__global__ void __launch_bounds__(kPartThreads)
partition_kernel(const uint8_t* __restrict__ bins_col, int64_t col_stride, const int32_t* __restrict__ ridx_in,
                 int32_t* __restrict__ ridx_out, const B2SplitWork* __restrict__ work, const B2LevelCtl* __restrict__ ctl,
                 int32_t* __restrict__ counters /* [2*n_work]: left, right */) {
  const int n_work = ctl->n_split, total_chunks = ctl->part_chunks;
  __shared__ int s_warp_left[kPartThreads / 32][kPartChunk / kPartThreads];
  __shared__ int s_base_left, s_base_right;
  __shared__ int s_pref[kPartThreads / 32][kPartChunk / kPartThreads];
  __shared__ uint32_t s_cat[8];   // category set of the chunk's split (all zero for a numeric split)
  constexpr int kIters = kPartChunk / kPartThreads;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int chunk = blockIdx.x; chunk < total_chunks; chunk += gridDim.x) {
    int lo = 0, hi = n_work - 1;
    while (lo < hi) {
      int mid = (lo + hi + 1) >> 1;
      if (__ldg(&work[mid].chunk_begin) <= chunk) lo = mid; else hi = mid - 1;
    }
    const B2SplitWork w = work[lo];
    const int row0 = (chunk - w.chunk_begin) * kPartChunk;
    const int nrows = min(kPartChunk, w.seg_count - row0);
    // the category set goes through shared memory so that the row loop below stays branch-free straight-line code
    // (a divergent global load in its body kept the compiler from batching the 8 row-id / bin-byte loads: the
    // kernel ran 1.7x slower, profiles/r01_summary.md)
    constexpr bool kCat = kMode != 0;
    if (kMode == 1) {
      if (threadIdx.x < 8) s_cat[threadIdx.x] = w.is_cat ? __ldg(&work[lo].cat_bits[threadIdx.x]) : 0u;
      __syncthreads();
    }
    uint32_t c0 = 0, c1 = 0, c2 = 0, c3 = 0, c4 = 0, c5 = 0, c6 = 0, c7 = 0;
    if (kMode == 2 && w.is_cat) {
      const uint4 lo4 = __ldg(reinterpret_cast<const uint4*>(work[lo].cat_bits)), hi4 = __ldg(reinterpret_cast<const uint4*>(work[lo].cat_bits) + 1);
      c0 = lo4.x; c1 = lo4.y; c2 = lo4.z; c3 = lo4.w; c4 = hi4.x; c5 = hi4.y; c6 = hi4.z; c7 = hi4.w;
    }
    const bool is_cat = kCat && w.is_cat != 0, has_missing = w.has_missing != 0, default_left = w.default_left != 0;
    int rid[kIters]; bool left[kIters]; unsigned bal[kIters];
#pragma unroll
    for (int it = 0; it < kIters; ++it) {
      const int r = it * kPartThreads + threadIdx.x;
      const bool valid = r < nrows;
      rid[it] = valid ? (ridx_in ? __ldg(ridx_in + w.seg_begin + row0 + r) : w.seg_begin + row0 + r) : 0;   // root: identity
      int b = valid ? (int)__ldg(bins_col + (int64_t)w.feature * col_stride + rid[it]) : 0;
      bool l;
      if (kCat) {
        uint32_t word;
        if (kMode == 1) word = s_cat[b >> 5];
        else {
          const uint32_t w01 = (b & 32) ? c1 : c0, w23 = (b & 32) ? c3 : c2, w45 = (b & 32) ? c5 : c4, w67 = (b & 32) ? c7 : c6;
          const uint32_t w03 = (b & 64) ? w23 : w01, w47 = (b & 64) ? w67 : w45;
          word = (b & 128) ? w47 : w03;
        }
        const bool in_set = ((word >> (b & 31)) & 1u) != 0u;                     // category in the set -> right
        const bool go_left = is_cat ? !in_set : (b <= w.split_bin);
        l = (has_missing && b == B2_MISSING_BIN) ? default_left : go_left;
      } else {
        l = (w.has_missing && b == B2_MISSING_BIN) ? (w.default_left != 0) : (b <= w.split_bin);
      }
      left[it] = valid && l;
      bal[it] = __ballot_sync(0xffffffffu, left[it]);
      if (lane == 0) s_warp_left[warp][it] = __popc(bal[it]);
    }
    __syncthreads();
    // exclusive prefix over (it, warp) in row order: index = it*8 + warp.  One warp scans the 64 counts with shuffles
    // (a single thread walking them serially kept the other 255 threads of the CTA waiting for ~2000 cycles per chunk)
    if (warp == 0) {
      constexpr int kCounts = kIters * (kPartThreads / 32);           // 64
      static_assert(kCounts == 64, "two counts per lane");
      const int i0 = 2 * lane, i1 = 2 * lane + 1;                       // index = it * 8 + wp
      const int c0 = s_warp_left[i0 & 7][i0 >> 3], c1 = s_warp_left[i1 & 7][i1 >> 3];
      int incl = c0 + c1;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
      const int excl = incl - (c0 + c1);
      s_pref[i0 & 7][i0 >> 3] = excl;
      s_pref[i1 & 7][i1 >> 3] = excl + c0;
      if (lane == 31) {
        // ONE 64-bit atomic per chunk claims both output ranges: low word = left rows so far (what finalize_level reads
        // as the left child's size), high word = rows so far; rights so far = rows - lefts.  At the shallow levels all
        // chunks hit the same one or two counters and the same-address atomics serialise (level 0: 9.8K of them).
        const int acc = incl;
        const unsigned long long old = atomicAdd(reinterpret_cast<unsigned long long*>(counters + 2 * lo),
                                                 ((unsigned long long)(unsigned)nrows << 32) | (unsigned long long)(unsigned)acc);
        s_base_left = (int)(unsigned)(old & 0xffffffffull);
        s_base_right = (int)(unsigned)(old >> 32) - s_base_left;
      }
    }
    __syncthreads();
    const int base_l = s_base_left, base_r = s_base_right;
#pragma unroll
    for (int it = 0; it < kIters; ++it) {
      const int r = it * kPartThreads + threadIdx.x;
      if (r < nrows) {
        const int lrank = s_pref[warp][it] + __popc(bal[it] & ((1u << lane) - 1u));
        if (left[it]) ridx_out[w.seg_begin + base_l + lrank] = rid[it];
        else {
          const int rrank = r - lrank;  // rights before this row inside the chunk
          ridx_out[w.seg_begin + w.seg_count - 1 - (base_r + rrank)] = rid[it];
        }
      }
    }
    __syncthreads();
  }
}

// ---- leaf refinement: 40-bit fixed-point sums of the fp32 gradients per leaf (exact int64)
__global__ void __launch_bounds__(256)
leaf_sums_kernel(const float2* __restrict__ gh, const int32_t* __restrict__ ridx0, const int32_t* __restrict__ ridx1,
                 const SegWork* __restrict__ work, const B2LevelCtl* __restrict__ ctl, const int32_t* __restrict__ qexp,
                 int leaf_bits, long long* __restrict__ sums /* [n_leaves][2] */,
                 uint16_t* __restrict__ pos /* nullable: row -> leaf index, read by margin_update_kernel */) {
  const int n_work = ctl->hist_n_work, total_chunks = ctl->hist_total_chunks;
  __shared__ long long sg[8], sh[8];
  const double kg = ldexp(1.0, leaf_bits - qexp[0]), kh = ldexp(1.0, leaf_bits - qexp[1]);
  for (int chunk = blockIdx.x; chunk < total_chunks; chunk += gridDim.x) {
    int lo = 0, hi = n_work - 1;
    while (lo < hi) {
      int mid = (lo + hi + 1) >> 1;
      if (__ldg(&work[mid].chunk_begin) <= chunk) lo = mid; else hi = mid - 1;
    }
    const SegWork w = work[lo];
    const int32_t* ridx = w.pad0 ? nullptr : (w.buf ? ridx1 : ridx0);   // pad0: the leaf is the root (rows = identity)
    const int row0 = (chunk - w.chunk_begin) * kPartChunk;
    const int nrows = min(kPartChunk, w.seg_count - row0);
    long long ag = 0, ah = 0;
    for (int r = threadIdx.x; r < nrows; r += blockDim.x) {
      const int row = ridx ? __ldg(ridx + w.seg_begin + row0 + r) : (w.seg_begin + row0 + r);
      if (pos) pos[row] = (uint16_t)w.id;
      const float2 v = __ldg(gh + row);
      ag += __double2ll_rn(__dmul_rn((double)v.x, kg));
      ah += __double2ll_rn(__dmul_rn((double)v.y, kh));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { ag += __shfl_xor_sync(0xffffffffu, ag, o); ah += __shfl_xor_sync(0xffffffffu, ah, o); }
    if ((threadIdx.x & 31) == 0) { sg[threadIdx.x >> 5] = ag; sh[threadIdx.x >> 5] = ah; }
    __syncthreads();
    if (threadIdx.x == 0) {
      long long tg = 0, th = 0;
      for (int i = 0; i < 8; ++i) { tg += sg[i]; th += sh[i]; }
      atomicAdd((unsigned long long*)&sums[2 * w.id], (unsigned long long)tg);
      atomicAdd((unsigned long long*)&sums[2 * w.id + 1], (unsigned long long)th);
    }
    __syncthreads();
  }
}

// margin[row*K + k] += leaf_value[leaf]
__global__ void __launch_bounds__(256)
pred_update_kernel(float* __restrict__ margin, int K, int k, const int32_t* __restrict__ ridx0,
                   const int32_t* __restrict__ ridx1, const SegWork* __restrict__ work, const B2LevelCtl* __restrict__ ctl,
                   const float* __restrict__ leaf_value) {
  const int n_work = ctl->hist_n_work, total_chunks = ctl->hist_total_chunks;
  for (int chunk = blockIdx.x; chunk < total_chunks; chunk += gridDim.x) {
    int lo = 0, hi = n_work - 1;
    while (lo < hi) {
      int mid = (lo + hi + 1) >> 1;
      if (__ldg(&work[mid].chunk_begin) <= chunk) lo = mid; else hi = mid - 1;
    }
    const SegWork w = work[lo];
    const int32_t* ridx = w.pad0 ? nullptr : (w.buf ? ridx1 : ridx0);
    const int row0 = (chunk - w.chunk_begin) * kPartChunk;
    const int nrows = min(kPartChunk, w.seg_count - row0);
    const float v = __ldg(leaf_value + w.id);
    for (int r = threadIdx.x; r < nrows; r += blockDim.x) {
      const int64_t row = ridx ? __ldg(ridx + w.seg_begin + row0 + r) : (w.seg_begin + row0 + r);
      margin[row * K + k] += v;
    }
  }
}

// ---- last split level (d = max_depth - 1): the children are leaves, so their rows need no ordered index list any
// more.  Instead of partition -> finalize -> leaf_sums -> pred_update (three scattered passes over all rows) one pass
// decides left/right, adds the row's fp32 gradient pair to the 40-bit fixed-point sums of the child LEAF and records
// the leaf index of the row; the margin is then updated by a streaming kernel (margin_update_kernel).
// Leaf index of child `side` of split j: leaf_base + 2 j + side -- exactly the index decide_kernel gives the node at
// the next level (leaves are numbered in node order; leaf_base = leaves that existed before this level's children).
template <bool kCat>
__global__ void __launch_bounds__(kPartThreads)
final_assign_kernel(const uint8_t* __restrict__ bins_col, int64_t col_stride, const int32_t* __restrict__ ridx_in,
                    const B2SplitWork* __restrict__ work, const B2LevelCtl* __restrict__ ctl, const float2* __restrict__ gh,
                    const int32_t* __restrict__ qexp, int leaf_bits, long long* __restrict__ sums,
                    uint16_t* __restrict__ pos) {
  const int n_work = ctl->n_split, total_chunks = ctl->part_chunks, leaf_base = ctl->leaf_base_next;
  __shared__ long long s_acc[kPartThreads / 32][4];
  const double kg = ldexp(1.0, leaf_bits - qexp[0]), kh = ldexp(1.0, leaf_bits - qexp[1]);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int chunk = blockIdx.x; chunk < total_chunks; chunk += gridDim.x) {
    int lo = 0, hi = n_work - 1;
    while (lo < hi) {
      int mid = (lo + hi + 1) >> 1;
      if (__ldg(&work[mid].chunk_begin) <= chunk) lo = mid; else hi = mid - 1;
    }
    const B2SplitWork w = work[lo];
    const int row0 = (chunk - w.chunk_begin) * kPartChunk;
    const int nrows = min(kPartChunk, w.seg_count - row0);
    const bool is_cat = kCat && w.is_cat != 0;
    const int leaf_l = leaf_base + 2 * lo;
    long long lg = 0, lh = 0, rg = 0, rh = 0;
#pragma unroll 4
    for (int r = threadIdx.x; r < nrows; r += kPartThreads) {
      const int rid = ridx_in ? __ldg(ridx_in + w.seg_begin + row0 + r) : w.seg_begin + row0 + r;
      const int b = (int)__ldg(bins_col + (int64_t)w.feature * col_stride + rid);
      bool go_left = b <= w.split_bin;
      if (kCat && is_cat) go_left = ((__ldg(&work[lo].cat_bits[b >> 5]) >> (b & 31)) & 1u) == 0u;   // category in the set -> right
      const bool l = (w.has_missing && b == B2_MISSING_BIN) ? (w.default_left != 0) : go_left;
      const float2 v = __ldg(gh + rid);
      const long long qg = __double2ll_rn(__dmul_rn((double)v.x, kg)), qh = __double2ll_rn(__dmul_rn((double)v.y, kh));
      if (l) { lg += qg; lh += qh; } else { rg += qg; rh += qh; }
      pos[rid] = (uint16_t)(leaf_l + (l ? 0 : 1));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      lg += __shfl_xor_sync(0xffffffffu, lg, o); lh += __shfl_xor_sync(0xffffffffu, lh, o);
      rg += __shfl_xor_sync(0xffffffffu, rg, o); rh += __shfl_xor_sync(0xffffffffu, rh, o);
    }
    if (lane == 0) { s_acc[warp][0] = lg; s_acc[warp][1] = lh; s_acc[warp][2] = rg; s_acc[warp][3] = rh; }
    __syncthreads();
    if (threadIdx.x < 4) {
      long long t = 0;
      for (int i = 0; i < kPartThreads / 32; ++i) t += s_acc[i][threadIdx.x];
      atomicAdd((unsigned long long*)&sums[2 * leaf_l + threadIdx.x], (unsigned long long)t);   // [leaf_l][g,h], [leaf_l+1][g,h]
    }
    __syncthreads();
  }
}

// margin[row*K + k] += leaf_value[pos[row]] -- streaming, rows in natural order
__global__ void __launch_bounds__(256)
margin_update_kernel(float* __restrict__ margin, int K, int k, const uint16_t* __restrict__ pos,
                     const float* __restrict__ leaf_value, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    margin[i * K + k] += __ldg(leaf_value + pos[i]);
}

__global__ void iota_kernel(int32_t* out, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) out[i] = (int32_t)i;
}

}  // namespace b2

extern "C" {
int b2_part_chunk_rows() { return b2::kPartChunk; }

int b2_launch_partition(const uint8_t* bins_col, int64_t col_stride, const int32_t* ridx_in, int32_t* ridx_out,
                        const B2SplitWork* work, const B2LevelCtl* ctl, int max_chunks, int32_t* counters, int any_categorical,
                        int num_sms, cudaStream_t stream) {
  if (max_chunks <= 0) return 0;
  int grid = max_chunks < num_sms * 8 ? max_chunks : num_sms * 8;
  static int cat_mode = -1;
  if (cat_mode < 0) { const char* e = getenv("B2_PART_CAT_MODE"); cat_mode = (e && atoi(e) == 2) ? 2 : 1; }
  if (any_categorical && cat_mode == 2)
    b2::partition_kernel<2><<<grid, b2::kPartThreads, 0, stream>>>(bins_col, col_stride, ridx_in, ridx_out, work, ctl, counters);
  else if (any_categorical)
    b2::partition_kernel<1><<<grid, b2::kPartThreads, 0, stream>>>(bins_col, col_stride, ridx_in, ridx_out, work, ctl, counters);
  else
    b2::partition_kernel<0><<<grid, b2::kPartThreads, 0, stream>>>(bins_col, col_stride, ridx_in, ridx_out, work, ctl, counters);
  return (int)cudaGetLastError();
}
int b2_launch_leaf_sums(const float2* gh, const int32_t* ridx0, const int32_t* ridx1, const void* work, const B2LevelCtl* ctl,
                        int max_chunks, const int32_t* qexp, int leaf_bits, long long* sums, uint16_t* pos, int num_sms,
                        cudaStream_t stream) {
  if (max_chunks <= 0) return 0;
  int grid = max_chunks < num_sms * 8 ? max_chunks : num_sms * 8;
  b2::leaf_sums_kernel<<<grid, 256, 0, stream>>>(gh, ridx0, ridx1, (const b2::SegWork*)work, ctl, qexp, leaf_bits, sums, pos);
  return (int)cudaGetLastError();
}
int b2_launch_final_assign(const uint8_t* bins_col, int64_t col_stride, const int32_t* ridx_in, const B2SplitWork* work,
                           const B2LevelCtl* ctl, int max_chunks, const float2* gh, const int32_t* qexp, int leaf_bits,
                           long long* sums, uint16_t* pos, int any_categorical, int num_sms, cudaStream_t stream) {
  if (max_chunks <= 0) return 0;
  int grid = max_chunks < num_sms * 8 ? max_chunks : num_sms * 8;
  if (any_categorical)
    b2::final_assign_kernel<true><<<grid, b2::kPartThreads, 0, stream>>>(bins_col, col_stride, ridx_in, work, ctl, gh, qexp, leaf_bits, sums, pos);
  else
    b2::final_assign_kernel<false><<<grid, b2::kPartThreads, 0, stream>>>(bins_col, col_stride, ridx_in, work, ctl, gh, qexp, leaf_bits, sums, pos);
  return (int)cudaGetLastError();
}
int b2_launch_margin_update(float* margin, int K, int k, const uint16_t* pos, const float* leaf_value, int64_t n, int num_sms,
                            cudaStream_t stream) {
  if (n <= 0) return 0;
  int64_t want = (n + 255) / 256;
  int grid = (int)(want < (int64_t)num_sms * 16 ? want : (int64_t)num_sms * 16);
  b2::margin_update_kernel<<<grid, 256, 0, stream>>>(margin, K, k, pos, leaf_value, n);
  return (int)cudaGetLastError();
}
int b2_launch_pred_update(float* margin, int K, int k, const int32_t* ridx0, const int32_t* ridx1, const void* work,
                          const B2LevelCtl* ctl, int max_chunks, const float* leaf_value, int num_sms, cudaStream_t stream) {
  if (max_chunks <= 0) return 0;
  int grid = max_chunks < num_sms * 8 ? max_chunks : num_sms * 8;
  b2::pred_update_kernel<<<grid, 256, 0, stream>>>(margin, K, k, ridx0, ridx1, (const b2::SegWork*)work, ctl, leaf_value);
  return (int)cudaGetLastError();
}
int b2_launch_iota(int32_t* out, int64_t n, cudaStream_t stream) {
  if (n <= 0) return 0;
  int grid = (int)((n + 1023) / 1024);
  if (grid > 4096) grid = 4096;
  b2::iota_kernel<<<grid, 256, 0, stream>>>(out, n);
  return (int)cudaGetLastError();
}
}
