// hist_kernel.cu -- feature x bin gradient/hessian histogram build for sm_100a.
//
// Replaces the BuildHist stage that the reference reaches through xgb.train()
// (xgboost_ray/main.py:745-752).  Bandwidth-bound scatter-reduce:
// no tensor cores.  Design (DESIGN.md "Histogram kernel"):
//
//  * A CTA owns ONE feature group (<= 32 features -> 32 "slots") of ONE node at a time and keeps
//    its histogram in shared memory as two int32 planes [256 bins][32 slots] (g and h): 64 KiB, so
//    three CTAs are resident per SM.  Because a bin row is exactly 32 words, the bank of an update
//    is its SLOT and does not depend on the bin value.
//  * A lane loads 16 bin bytes of one row (LDG.128 straight to registers; staging rows through
//    shared memory would spend the shared-memory bandwidth that the atomics are bound by).  Two
//    lanes cover the 32-byte group slice of a row, 16 rows per warp.  Lane l pre-rotates its bytes
//    by (l>>1) so that at step j the 32 lanes of a warp touch 32 DIFFERENT slots: every ATOMS.ADD
//    is bank-conflict free for any data, one wavefront per instruction.
//  * Sums are exact integers (fixed-point gradients), so the result is independent of the order
//    of rows, CTAs and GPUs.  A CTA flushes its planes to the global int64 histogram when it moves
//    to another node; every `window_rows` = 2^(30-qbits) rows it flushes just the cells that reached
//    2^30 (none for well spread bins), which is what keeps the int32 cells from overflowing.
//  * Rows of a node are addressed through the row-index segment list (gather) except at the root.
#include <stdlib.h>

#include "hist_common.cuh"

#ifndef B2_HIST_DEFAULT_VARIANT
#define B2_HIST_DEFAULT_VARIANT 2
#endif

namespace b2 {

// kGPC = feature groups per CTA.  1: 64 KiB of histogram, two lanes per row, 16 rows per warp step.
// 2: 128 KiB (one 1024-thread CTA per SM); FOUR lanes read 64 contiguous bytes of a row (one L1 wavefront
// instead of two) and the row id / gradient pair loads are shared by both groups, which removes about a
// quarter of the L1TEX wavefronts per row -- the pipe this kernel is bound by (profiles/r01_summary.md).
template <bool kGather, int kThreads, int kMinBlocks, int kGPC>
__global__ void __launch_bounds__(kThreads, kMinBlocks)
hist_build_kernel(const uint8_t* __restrict__ bins, int row_stride, const int2* __restrict__ gpair,
                  const int32_t* __restrict__ ridx, const B2HistWork* __restrict__ work, int n_work,
                  int total_chunks, int chunk_rows, int window_rows, int n_groups, long long* __restrict__ hist,
                  const B2LevelCtl* __restrict__ ctl, int log2_shards, int node_cap, int debug_mode) {
  constexpr int kLanesPerRow = 2 * kGPC;
  constexpr int kRowsPerWarp = 32 / kLanesPerRow;
  HistTarget target; target.base = (unsigned long long*)hist; target.log2_shards = log2_shards; target.node_cap = node_cap;
  target.n_groups = n_groups;
  if (ctl) { n_work = ctl->hist_n_work; total_chunks = ctl->hist_total_chunks; chunk_rows = ctl->hist_chunk_rows; }
  extern __shared__ __align__(16) int32_t s_hist[];  // [kGPC][256][2][32]
  const int n_cta_groups = (n_groups + kGPC - 1) / kGPC;
  const int group0 = (blockIdx.x % n_cta_groups) * kGPC;
  const int stream = blockIdx.x / n_cta_groups;
  const int n_streams = gridDim.x / n_cta_groups;
  if (total_chunks <= 0) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
  // lane -> (row of the warp step, group of the CTA, 16-byte half of the group slice).  rot is distinct for the
  // 16 (row, group) pairs of a warp, so at every step the 32 lanes hit 32 different banks (bank = slot).
  const int sub = lane / kLanesPerRow, gsel = (lane % kLanesPerRow) >> 1, half = lane & 1;
  const int rot = sub * kGPC + gsel;
  const bool active = group0 + gsel < n_groups;
  const int lane_byte_off = (group0 + gsel) * 32 + half * 16;
  const uint32_t smem_g = (uint32_t)__cvta_generic_to_shared(s_hist) + gsel * (B2_GROUP_ELEMS * 4);

  for (int e = threadIdx.x; e < kGPC * B2_GROUP_ELEMS; e += blockDim.x) s_hist[e] = 0;
  __syncthreads();

  int cur = -1;          // work index whose partial sums are in shared memory
  int rows_in_window = 0;
  // A stream takes a CONTIGUOUS range of the chunk list, so consecutive chunks of a CTA mostly belong to the same
  // node and the 2 x 16K-cell node flush happens once per node per CTA instead of once per chunk (with the strided
  // assignment of round 1 every chunk of a deep level was a node change: levels 6-7 cost 1.5x the root per row).
  // debug_mode bit 2 (B2_HIST_DEBUG_MODE=4) restores the strided assignment for A/B timing.
  const bool strided = (debug_mode & 4) != 0;
  debug_mode &= 3;
  const int c_begin = strided ? stream : (int)(((long long)stream * total_chunks) / n_streams);
  const int c_end = strided ? total_chunks : (int)(((long long)(stream + 1) * total_chunks) / n_streams);
  const int c_step = strided ? n_streams : 1;
  for (int chunk = c_begin; chunk < c_end; chunk += c_step) {
    // locate the node of this chunk (uniform across the CTA): last w with chunk_begin <= chunk
    int lo = 0, hi = n_work - 1;
    while (lo < hi) {
      int mid = (lo + hi + 1) >> 1;
      if (__ldg(&work[mid].chunk_begin) <= chunk) lo = mid; else hi = mid - 1;
    }
    const int w = lo;
    const int seg_begin = __ldg(&work[w].seg_begin), seg_count = __ldg(&work[w].seg_count);
    const int row0 = (chunk - __ldg(&work[w].chunk_begin)) * chunk_rows;
    const int nrows = min(chunk_rows, seg_count - row0);
    if (cur >= 0 && w != cur) {
      // node change: add the partial sums to the global int64 histogram
      __syncthreads();
#pragma unroll
      for (int gs = 0; gs < kGPC; ++gs)
        if (group0 + gs < n_groups) flush_planes(s_hist + gs * B2_GROUP_ELEMS, target, __ldg(&work[cur].hist_index), group0 + gs);
      __syncthreads();
      rows_in_window = 0;
    } else if (cur >= 0 && rows_in_window + nrows > window_rows) {
      // same node, overflow guard interval reached: flush only the (rare) cells at or above 2^30
      // (measured alternatives, profiles/r01_summary.md: flushing all cells with RED.64 every window cost 17 % of the
      // kernel; a CTA-private int64 scratch with plain read-modify-write was 30 % slower still)
      __syncthreads();
#pragma unroll
      for (int gs = 0; gs < kGPC; ++gs)
        if (group0 + gs < n_groups) flush_large_cells(s_hist + gs * B2_GROUP_ELEMS, target, __ldg(&work[cur].hist_index), group0 + gs);
      __syncthreads();
      rows_in_window = 0;
    }
    cur = w;
    rows_in_window += nrows;
    const int64_t pos0 = (int64_t)seg_begin + row0;
    const int iter_rows = n_warps * kRowsPerWarp;
    // 3-stage register pipeline: while stage k is accumulated, the loads of the next two
    // iterations are in flight, and (gather) the row ids of three more iterations behind them,
    // so no load waits on the ridx -> bins dependency.
    const int rbase = warp * kRowsPerWarp;
    const int r0 = rbase + sub;
    const int lim = active ? nrows : 0;   // lanes of a group past the last one (odd group count) load nothing
    int64_t id0 = fetch_rid<kGather>(ridx, pos0, r0, lim);
    int64_t id1 = fetch_rid<kGather>(ridx, pos0, r0 + iter_rows, lim);
    int64_t id2 = fetch_rid<kGather>(ridx, pos0, r0 + 2 * iter_rows, lim);
    unsigned sink = 0;
#define B2_LOAD(id) (debug_mode == 2 ? fake_row(id) : load_row_id(bins, gpair, id, row_stride, lane_byte_off))
    RowData s0 = B2_LOAD(id0);
    id0 = fetch_rid<kGather>(ridx, pos0, r0 + 3 * iter_rows, lim);
    RowData s1 = B2_LOAD(id1);
    id1 = fetch_rid<kGather>(ridx, pos0, r0 + 4 * iter_rows, lim);
    RowData s2 = B2_LOAD(id2);
    id2 = fetch_rid<kGather>(ridx, pos0, r0 + 5 * iter_rows, lim);
    for (int r = rbase; r < nrows; r += 3 * iter_rows) {   // warp-uniform trip count
      if (kGPC == 1 || active) accumulate_row(s0, smem_g, rot, half, debug_mode, &sink);
      s0 = B2_LOAD(id0);
      id0 = fetch_rid<kGather>(ridx, pos0, r + sub + 6 * iter_rows, lim);
      if (r + iter_rows < nrows && (kGPC == 1 || active)) accumulate_row(s1, smem_g, rot, half, debug_mode, &sink);
      s1 = B2_LOAD(id1);
      id1 = fetch_rid<kGather>(ridx, pos0, r + sub + 7 * iter_rows, lim);
      if (r + 2 * iter_rows < nrows && (kGPC == 1 || active)) accumulate_row(s2, smem_g, rot, half, debug_mode, &sink);
      s2 = B2_LOAD(id2);
      id2 = fetch_rid<kGather>(ridx, pos0, r + sub + 8 * iter_rows, lim);
    }
#undef B2_LOAD
    if (sink == 0x9e3779b9u) s_hist[threadIdx.x] = (int)sink;
  }
  if (cur >= 0) {
    __syncthreads();
#pragma unroll
    for (int gs = 0; gs < kGPC; ++gs)
      if (group0 + gs < n_groups) flush_planes(s_hist + gs * B2_GROUP_ELEMS, target, __ldg(&work[cur].hist_index), group0 + gs);
  }
}

// ---------------------------------------------------------------- sibling = parent - built
__global__ void hist_subtract_kernel(const long long* __restrict__ parent_level, long long* __restrict__ level,
                                     const int32_t* __restrict__ triples, int n_pairs, int64_t node_elems,
                                     const B2LevelCtl* __restrict__ ctl) {
  if (ctl) n_pairs = ctl->n_pairs;
  // triples[3*p] = parent slot (prev level), built slot, sibling slot (this level)
  const int p = blockIdx.y;
  if (p >= n_pairs) return;
  const long long* par = parent_level + (size_t)triples[3 * p] * node_elems;
  const long long* built = level + (size_t)triples[3 * p + 1] * node_elems;
  long long* sib = level + (size_t)triples[3 * p + 2] * node_elems;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < node_elems; i += (int64_t)gridDim.x * blockDim.x)
    sib[i] = par[i] - built[i];
}

}  // namespace b2

extern "C" {

// Launch on `stream`.  grid = n_groups * n_streams persistent CTAs; returns the cudaError.
int b2_launch_hist(const uint8_t* bins, int row_stride, const int2* gpair, const int32_t* ridx,
                   const B2HistWork* work, int n_work, int total_chunks, int chunk_rows, int window_rows,
                   int n_groups, long long* hist, const B2LevelCtl* ctl, int log2_shards, int node_cap, int num_sms,
                   cudaStream_t stream) {
  static bool attr_set = false;
  static int debug_mode = -1, variant = -1;
  if (debug_mode < 0) { const char* e = getenv("B2_HIST_DEBUG_MODE"); debug_mode = e ? atoi(e) : 0; }
  // variants (B2_HIST_VARIANT): 0 = 256 threads x 3 CTAs/SM, one group per CTA
  //                             1 = 512 threads x 2 CTAs/SM, one group per CTA
  //                             2 = 1024 threads x 1 CTA/SM, two groups per CTA (default; A/B in profiles/r01_summary.md)
  if (variant < 0) {
    const char* e = getenv("B2_HIST_VARIANT");
    variant = e ? atoi(e) : B2_HIST_DEFAULT_VARIANT;
    const char* t = getenv("B2_HIST_THREADS");   // older spelling of variant 0
    if (!e && t && atoi(t) == 256) variant = 0;
    if (variant < 0 || variant > 2) variant = B2_HIST_DEFAULT_VARIANT;
  }
  const int gpc = variant == 2 ? 2 : 1;
  const int smem = gpc * B2_GROUP_ELEMS * (int)sizeof(int32_t);  // 64 KiB per group
  if (!attr_set) {
    cudaFuncSetAttribute(b2::hist_build_kernel<true, 256, 3, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, B2_GROUP_ELEMS * 4);
    cudaFuncSetAttribute(b2::hist_build_kernel<false, 256, 3, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, B2_GROUP_ELEMS * 4);
    cudaFuncSetAttribute(b2::hist_build_kernel<true, 512, 2, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, B2_GROUP_ELEMS * 4);
    cudaFuncSetAttribute(b2::hist_build_kernel<false, 512, 2, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, B2_GROUP_ELEMS * 4);
    cudaFuncSetAttribute(b2::hist_build_kernel<true, 1024, 1, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, B2_GROUP_ELEMS * 8);
    cudaFuncSetAttribute(b2::hist_build_kernel<false, 1024, 1, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, B2_GROUP_ELEMS * 8);
    attr_set = true;
  }
  // with ctl the work list / chunk counts live in device memory (sync-free level loop) and the grid is the
  // full persistent grid; without it they are host values
  if (!ctl && (total_chunks <= 0 || n_work <= 0)) return 0;
  if (!ctl && chunk_rows > window_rows) return (int)cudaErrorInvalidValue;  // a chunk must fit one int32 window
  const int ctas_per_sm = variant == 0 ? 3 : variant == 1 ? 2 : 1;
  const int n_cta_groups = (n_groups + gpc - 1) / gpc;
  int n_streams = (num_sms * ctas_per_sm) / n_cta_groups;
  if (n_streams < 1) n_streams = 1;
  if (!ctl && n_streams > total_chunks) n_streams = total_chunks;
  dim3 grid(n_cta_groups * n_streams), block(variant == 0 ? 256 : variant == 1 ? 512 : 1024);
#define B2_HIST_ARGS bins, row_stride, gpair, ridx, work, n_work, total_chunks, chunk_rows, window_rows, n_groups, hist, ctl, log2_shards, \
                     node_cap, debug_mode
  if (variant == 2) {
    if (ridx) b2::hist_build_kernel<true, 1024, 1, 2><<<grid, block, smem, stream>>>(B2_HIST_ARGS);
    else b2::hist_build_kernel<false, 1024, 1, 2><<<grid, block, smem, stream>>>(B2_HIST_ARGS);
  } else if (variant == 1) {
    if (ridx) b2::hist_build_kernel<true, 512, 2, 1><<<grid, block, smem, stream>>>(B2_HIST_ARGS);
    else b2::hist_build_kernel<false, 512, 2, 1><<<grid, block, smem, stream>>>(B2_HIST_ARGS);
  } else {
    if (ridx) b2::hist_build_kernel<true, 256, 3, 1><<<grid, block, smem, stream>>>(B2_HIST_ARGS);
    else b2::hist_build_kernel<false, 256, 3, 1><<<grid, block, smem, stream>>>(B2_HIST_ARGS);
  }
#undef B2_HIST_ARGS
  return (int)cudaGetLastError();
}

int b2_launch_hist_subtract(const long long* parent_level, long long* level, const int32_t* triples, int n_pairs,
                            int64_t node_elems, const B2LevelCtl* ctl, cudaStream_t stream) {
  if (n_pairs <= 0) return 0;   // with ctl: n_pairs is the upper bound, the real count is read on the device
  int bx = (int)((node_elems + 256 * 8 - 1) / (256 * 8));
  if (bx < 1) bx = 1;
  dim3 grid(bx, n_pairs);
  b2::hist_subtract_kernel<<<grid, 256, 0, stream>>>(parent_level, level, triples, n_pairs, node_elems, ctl);
  return (int)cudaGetLastError();
}
}
