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

#include "common.cuh"

namespace b2 {

constexpr int kHistThreads = 256;
constexpr int kRowsPerWarpIter = 16;

__device__ __forceinline__ uint4 ldg_nc_v4(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void red_shared_add(uint32_t saddr, int v) {
  asm volatile("red.shared.add.s32 [%0], %1;" ::"r"(saddr), "r"(v) : "memory");
}

// result byte i = source byte (i + rot) & 15
__device__ __forceinline__ uint4 rotate_bytes(uint4 v, int rot) {
  uint32_t w0 = v.x, w1 = v.y, w2 = v.z, w3 = v.w;
  if (rot & 4) { uint32_t t = w0; w0 = w1; w1 = w2; w2 = w3; w3 = t; }
  if (rot & 8) { uint32_t t0 = w0, t1 = w1; w0 = w2; w1 = w3; w2 = t0; w3 = t1; }
  int bs = (rot & 3) * 8;
  uint4 r;
  r.x = __funnelshift_r(w0, w1, bs);
  r.y = __funnelshift_r(w1, w2, bs);
  r.z = __funnelshift_r(w2, w3, bs);
  r.w = __funnelshift_r(w3, w0, bs);
  return r;
}

struct RowData {
  uint4 bins;
  int2 gp;
};

template <bool kGather>
__device__ __forceinline__ RowData load_row(const uint8_t* __restrict__ bins, const int2* __restrict__ gpair,
                                            const int32_t* __restrict__ ridx, int64_t pos, bool valid,
                                            int row_stride, int lane_byte_off) {
  RowData d;
  d.bins = make_uint4(0, 0, 0, 0);
  d.gp = make_int2(0, 0);
  if (valid) {
    int64_t rid = kGather ? (int64_t)__ldg(ridx + pos) : pos;
    d.bins = ldg_nc_v4(bins + rid * row_stride + lane_byte_off);
    d.gp = __ldg(gpair + rid);
  }
  return d;
}

// row id of chunk-row r (or -1 past the end of the chunk); root level: identity
template <bool kGather>
__device__ __forceinline__ int64_t fetch_rid(const int32_t* __restrict__ ridx, int64_t pos0, int r, int nrows) {
  if (r >= nrows) return -1;
  return kGather ? (int64_t)__ldg(ridx + pos0 + r) : pos0 + r;
}
__device__ __forceinline__ RowData load_row_id(const uint8_t* __restrict__ bins, const int2* __restrict__ gpair,
                                               int64_t rid, int row_stride, int lane_byte_off) {
  RowData d;
  d.bins = make_uint4(0, 0, 0, 0);
  d.gp = make_int2(0, 0);
  if (rid >= 0) {
    d.bins = ldg_nc_v4(bins + rid * row_stride + lane_byte_off);
    d.gp = __ldg(gpair + rid);
  }
  return d;
}
// diagnostic (debug_mode 2): synthesise the row instead of loading it
__device__ __forceinline__ RowData fake_row(int64_t rid) {
  RowData d;
  uint32_t x = (uint32_t)rid * 2654435761u + 12345u;
  d.bins = make_uint4(x, x * 1664525u + 1013904223u, x ^ (x >> 13), x * 22695477u + 1u);
  d.gp = make_int2(rid >= 0 ? 3 : 0, rid >= 0 ? 1 : 0);
  return d;
}

// 16 steps: one byte (= one feature slot) per step, two conflict-free shared atomics per step.
// Shared layout: int32 [256 bins][2 planes (g,h)][32 slots] = 256 B per bin, so bin*256 is the byte
// placed at byte position 1 by ONE prmt; the cell address is that plus a per-lane, per-step offset.
__device__ __forceinline__ void accumulate_row(const RowData& d, uint32_t smem_g, int rot, int half, int debug_mode = 0,
                                               unsigned* sink = nullptr) {
  if (debug_mode == 1) {  // diagnostic: consume the loads without touching shared memory
    *sink += d.bins.x ^ d.bins.y ^ d.bins.z ^ d.bins.w ^ (unsigned)d.gp.x ^ (unsigned)d.gp.y;
    return;
  }
  uint4 b = rotate_bytes(d.bins, rot);
  const uint32_t w[4] = {b.x, b.y, b.z, b.w};
  const uint32_t base = smem_g + half * 64;
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const uint32_t bin256 = __byte_perm(w[j >> 2], 0u, 0x4404u | ((uint32_t)(j & 3) << 4));
    const uint32_t slot_off = ((uint32_t)(j + rot) & 15u) * 4u;
    const uint32_t a = base + bin256 + slot_off;
    red_shared_add(a, d.gp.x);
    red_shared_add(a + B2_GROUP_SLOTS * 4, d.gp.y);
  }
}

// shared cell e = bin*64 + plane*32 + slot  ->  global cell plane*8192 + bin*32 + slot
// Global histogram layout (feature-slot sharded for the reduce-scatter):
//   int64 [shards][node_cap][group][plane][256 bins][sp]   sp = 32 / shards, slot s lives on shard s % shards at s / shards.
// shards == 1 degenerates to [node][group][plane][bin][32].
struct HistTarget {
  unsigned long long* base;   // build buffer
  int log2_shards;            // shards = 1 << log2_shards
  int node_cap;               // node slots per shard in this launch's buffer
  int n_groups;
};
__device__ __forceinline__ size_t target_index(const HistTarget& t, int node_slot, int group, int e) {
  const int bin = e >> 6, plane = (e >> 5) & 1, slot = e & 31;
  const int shards = 1 << t.log2_shards, sp = B2_GROUP_SLOTS >> t.log2_shards;
  const int r = slot & (shards - 1), sl = slot >> t.log2_shards;
  const size_t slice_elems = (size_t)t.n_groups * 2 * B2_BINS * sp;
  return ((size_t)r * t.node_cap + node_slot) * slice_elems + ((size_t)(group * 2 + plane) * B2_BINS + bin) * sp + sl;
}

// window flush: move the int32 partial sums into this CTA's PRIVATE int64 scratch (plain coalesced
// read-modify-write in L2, no atomics: only this CTA touches its scratch block)
__device__ __forceinline__ void flush_to_scratch(int32_t* s_hist, long long* scratch) {
  for (int e = threadIdx.x * 4; e < B2_GROUP_ELEMS; e += blockDim.x * 4) {
    int4 v = *reinterpret_cast<int4*>(s_hist + e);
    if ((v.x | v.y | v.z | v.w) != 0) {
      longlong2 a = *reinterpret_cast<longlong2*>(scratch + e), b = *reinterpret_cast<longlong2*>(scratch + e + 2);
      a.x += v.x; a.y += v.y; b.x += v.z; b.y += v.w;
      *reinterpret_cast<longlong2*>(scratch + e) = a; *reinterpret_cast<longlong2*>(scratch + e + 2) = b;
      *reinterpret_cast<int4*>(s_hist + e) = make_int4(0, 0, 0, 0);
    }
  }
}
// lazy window flush: only cells whose magnitude reached 2^30 are moved to the global histogram.  Called
// (between barriers) at least every `window_rows` = 2^(30 - qbits) rows, during which a cell can grow by
// less than 2^30, so no int32 cell can overflow; for well spread bins nothing is flushed at all.
__device__ __forceinline__ void flush_large_cells(int32_t* s_hist, const HistTarget& t, int node_slot, int group) {
  for (int e = threadIdx.x * 4; e < B2_GROUP_ELEMS; e += blockDim.x * 4) {
    const int4 v = *reinterpret_cast<const int4*>(s_hist + e);
    const int vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (vv[k] >= (1 << 30) || vv[k] <= -(1 << 30)) {
        atomicAdd(t.base + target_index(t, node_slot, group, e + k), (unsigned long long)(long long)vv[k]);
        s_hist[e + k] = 0;
      }
    }
  }
}
// node flush: shared (+ scratch if it was used) -> global int64 histogram with atomics
__device__ __forceinline__ void flush_planes(int32_t* s_hist, long long* scratch, bool scratch_dirty, const HistTarget& t,
                                             int node_slot, int group) {
  for (int e = threadIdx.x; e < B2_GROUP_ELEMS; e += blockDim.x) {
    long long v = s_hist[e];
    if (scratch_dirty) { v += scratch[e]; scratch[e] = 0; }
    if (v != 0) atomicAdd(t.base + target_index(t, node_slot, group, e), (unsigned long long)v);
    s_hist[e] = 0;
  }
}

template <bool kGather>
__global__ void __launch_bounds__(kHistThreads, 3)
hist_build_kernel(const uint8_t* __restrict__ bins, int row_stride, const int2* __restrict__ gpair,
                  const int32_t* __restrict__ ridx, const B2HistWork* __restrict__ work, int n_work,
                  int total_chunks, int chunk_rows, int window_rows, int n_groups, long long* __restrict__ hist,
                  const B2LevelCtl* __restrict__ ctl, long long* __restrict__ scratch_all, int log2_shards, int node_cap,
                  int debug_mode) {
  HistTarget target; target.base = (unsigned long long*)hist; target.log2_shards = log2_shards; target.node_cap = node_cap;
  target.n_groups = n_groups;
  if (ctl) { n_work = ctl->hist_n_work; total_chunks = ctl->hist_total_chunks; chunk_rows = ctl->hist_chunk_rows; }
  extern __shared__ __align__(16) int32_t s_hist[];  // [2][256][32]
  __shared__ int s_cur_work;
  const int group = blockIdx.x % n_groups;
  const int stream = blockIdx.x / n_groups;
  const int n_streams = gridDim.x / n_groups;
  if (stream >= total_chunks) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
  const int rot = lane >> 1, half = lane & 1;
  const int lane_byte_off = group * 32 + half * 16;
  const uint32_t smem_g = (uint32_t)__cvta_generic_to_shared(s_hist);

  for (int e = threadIdx.x; e < B2_GROUP_ELEMS; e += blockDim.x) s_hist[e] = 0;
  __syncthreads();

  int cur = -1;          // work index whose partial sums are in shared memory
  int rows_in_window = 0;
  bool scratch_dirty = false;
  long long* scratch = scratch_all + (size_t)blockIdx.x * B2_GROUP_ELEMS;
  for (int chunk = stream; chunk < total_chunks; chunk += n_streams) {
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
      flush_planes(s_hist, scratch, false, target, __ldg(&work[cur].hist_index), group);
      __syncthreads();
      rows_in_window = 0;
    } else if (cur >= 0 && rows_in_window + nrows > window_rows) {
      // same node, overflow guard interval reached: flush only the (rare) cells at or above 2^30
      // (measured alternatives: flushing all cells with RED.64 every window cost 17 % of the kernel; a
      // CTA-private int64 scratch with plain read-modify-write was 30 % slower still)
      __syncthreads();
      flush_large_cells(s_hist, target, __ldg(&work[cur].hist_index), group);
      __syncthreads();
      rows_in_window = 0;
    }
    cur = w;
    rows_in_window += nrows;
    const int64_t pos0 = (int64_t)seg_begin + row0;
    const int iter_rows = n_warps * kRowsPerWarpIter;
    // 3-stage register pipeline: while stage k is accumulated, the loads of the next two
    // iterations are in flight, and (gather) the row ids of three more iterations behind them,
    // so no load waits on the ridx -> bins dependency.
    const int r0 = warp * kRowsPerWarpIter + rot;
    int64_t id0 = fetch_rid<kGather>(ridx, pos0, r0, nrows);
    int64_t id1 = fetch_rid<kGather>(ridx, pos0, r0 + iter_rows, nrows);
    int64_t id2 = fetch_rid<kGather>(ridx, pos0, r0 + 2 * iter_rows, nrows);
    unsigned sink = 0;
#define B2_LOAD(id) (debug_mode == 2 ? fake_row(id) : load_row_id(bins, gpair, id, row_stride, lane_byte_off))
    RowData s0 = B2_LOAD(id0);
    id0 = fetch_rid<kGather>(ridx, pos0, r0 + 3 * iter_rows, nrows);
    RowData s1 = B2_LOAD(id1);
    id1 = fetch_rid<kGather>(ridx, pos0, r0 + 4 * iter_rows, nrows);
    RowData s2 = B2_LOAD(id2);
    id2 = fetch_rid<kGather>(ridx, pos0, r0 + 5 * iter_rows, nrows);
    for (int r = r0 - rot; r < nrows; r += 3 * iter_rows) {   // warp-uniform trip count
      accumulate_row(s0, smem_g, rot, half, debug_mode, &sink);
      s0 = B2_LOAD(id0);
      id0 = fetch_rid<kGather>(ridx, pos0, r + rot + 6 * iter_rows, nrows);
      if (r + iter_rows < nrows) accumulate_row(s1, smem_g, rot, half, debug_mode, &sink);
      s1 = B2_LOAD(id1);
      id1 = fetch_rid<kGather>(ridx, pos0, r + rot + 7 * iter_rows, nrows);
      if (r + 2 * iter_rows < nrows) accumulate_row(s2, smem_g, rot, half, debug_mode, &sink);
      s2 = B2_LOAD(id2);
      id2 = fetch_rid<kGather>(ridx, pos0, r + rot + 8 * iter_rows, nrows);
    }
#undef B2_LOAD
    if (sink == 0x9e3779b9u) s_hist[threadIdx.x] = (int)sink;
  }
  if (cur >= 0) {
    __syncthreads();
    flush_planes(s_hist, scratch, scratch_dirty, target, __ldg(&work[cur].hist_index), group);
  }
  (void)s_cur_work;
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
// scratch: zero-initialised int64 [b2_hist_scratch_elems(n_groups, num_sms)], kept all-zero between launches
size_t b2_hist_scratch_elems(int n_groups, int num_sms) {
  int n_streams = (num_sms * 3) / n_groups;
  if (n_streams < 1) n_streams = 1;
  return (size_t)n_groups * n_streams * B2_GROUP_ELEMS;
}

int b2_launch_hist(const uint8_t* bins, int row_stride, const int2* gpair, const int32_t* ridx,
                   const B2HistWork* work, int n_work, int total_chunks, int chunk_rows, int window_rows,
                   int n_groups, long long* hist, const B2LevelCtl* ctl, long long* scratch, int log2_shards, int node_cap,
                   int num_sms, cudaStream_t stream) {
  static bool attr_set = false;
  static int debug_mode = -1;
  if (debug_mode < 0) { const char* e = getenv("B2_HIST_DEBUG_MODE"); debug_mode = e ? atoi(e) : 0; }
  const int smem = B2_GROUP_ELEMS * (int)sizeof(int32_t);  // 64 KiB
  if (!attr_set) {
    cudaFuncSetAttribute(b2::hist_build_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(b2::hist_build_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    attr_set = true;
  }
  // with ctl the work list / chunk counts live in device memory (sync-free level loop) and the grid is the
  // full persistent grid; without it they are host values
  if (!ctl && (total_chunks <= 0 || n_work <= 0)) return 0;
  if (!ctl && chunk_rows > window_rows) return (int)cudaErrorInvalidValue;  // a chunk must fit one int32 window
  int n_streams = (num_sms * 3) / n_groups;
  if (n_streams < 1) n_streams = 1;
  if (!ctl && n_streams > total_chunks) n_streams = total_chunks;
  dim3 grid(n_groups * n_streams), block(b2::kHistThreads);
  if (ridx)
    b2::hist_build_kernel<true><<<grid, block, smem, stream>>>(bins, row_stride, gpair, ridx, work, n_work, total_chunks,
                                                              chunk_rows, window_rows, n_groups, hist, ctl, scratch, log2_shards, node_cap, debug_mode);
  else
    b2::hist_build_kernel<false><<<grid, block, smem, stream>>>(bins, row_stride, gpair, ridx, work, n_work, total_chunks,
                                                               chunk_rows, window_rows, n_groups, hist, ctl, scratch, log2_shards, node_cap, debug_mode);
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
