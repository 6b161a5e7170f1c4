// hist_tma_kernel.cu -- TMA-staged variant of the histogram build (sm_100a).
//
// Same work distribution, shared-memory layout, conflict-free atomics and flush rules as
// hist_kernel.cu, but the 32-byte row slices of a CTA's feature group are brought into shared
// memory by the TMA engine (cp.async.bulk.tensor.2d tile::gather4: four arbitrary rows per
// instruction, addressed through the node's row-index segment), so the global data no longer
// returns through the LSU write-back path that the shared atomics are bound by
// (profiles/r01_summary.md).  One producer warp issues the loads into a 4-stage ring guarded by
// mbarriers; eight consumer warps read their 16 bytes with LDS.128 and accumulate.
#include <cuda.h>
#include <stdlib.h>

#include "hist_common.cuh"

namespace b2 {

constexpr int kTmaConsumerWarps = 8;
constexpr int kTmaThreads = (kTmaConsumerWarps + 1) * 32;
constexpr int kStageRows = 64;
constexpr int kStages = 4;
constexpr int kStageBytes = kStageRows * B2_GROUP_SLOTS;   // 2048
constexpr int kTmaSmemBytes = B2_GROUP_ELEMS * 4 + kStages * kStageBytes + 2 * kStages * 8;

__device__ __forceinline__ void mbar_init(uint32_t a, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a), "r"(count)); }
__device__ __forceinline__ void mbar_arrive(uint32_t a) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(a) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t a, int bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(a), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t a, int parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(ok) : "r"(a), "r"(parity) : "memory");
  }
}
__device__ __forceinline__ void tma_gather4(uint32_t dst, const CUtensorMap* tmap, uint32_t bar, int col, int r0, int r1, int r2, int r3) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(dst), "l"(tmap), "r"(bar), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3) : "memory");
}
__device__ __forceinline__ void tma_tile(uint32_t dst, const CUtensorMap* tmap, uint32_t bar, int col, int row) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(dst), "l"(tmap), "r"(bar), "r"(col), "r"(row) : "memory");
}
__device__ __forceinline__ void consumer_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kTmaConsumerWarps * 32) : "memory"); }

// the CTA's sequence of 64-row blocks; every warp walks it identically
struct BlockIter {
  const B2HistWork* work; int n_work, total_chunks, chunk_rows, n_streams;
  int chunk, w, blk, n_blocks, nrows, k;   // k = running block counter (ring position)
  int64_t pos0;
  __device__ void load_chunk() {
    if (chunk >= total_chunks) return;
    int lo = 0, hi = n_work - 1;
    while (lo < hi) { int mid = (lo + hi + 1) >> 1; if (__ldg(&work[mid].chunk_begin) <= chunk) lo = mid; else hi = mid - 1; }
    w = lo;
    const int row0 = (chunk - __ldg(&work[w].chunk_begin)) * chunk_rows;
    nrows = min(chunk_rows, __ldg(&work[w].seg_count) - row0);
    pos0 = (int64_t)__ldg(&work[w].seg_begin) + row0;
    n_blocks = (nrows + kStageRows - 1) / kStageRows;
    blk = 0;
  }
  __device__ void init(const B2HistWork* wk, int nw, int tc, int cr, int stream, int ns) {
    work = wk; n_work = nw; total_chunks = tc; chunk_rows = cr; n_streams = ns; chunk = stream; k = 0; w = -1; blk = 0; n_blocks = 0;
    nrows = 0; pos0 = 0;
    load_chunk();
  }
  __device__ bool done() const { return chunk >= total_chunks; }
  __device__ void next() {   // to the next block; crosses chunk boundaries
    ++k; ++blk;
    if (blk >= n_blocks) { chunk += n_streams; load_chunk(); }
  }
};

template <bool kGather>
__global__ void __launch_bounds__(kTmaThreads, 3)
hist_build_tma_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ CUtensorMap tmap_tile,
                      const int2* __restrict__ gpair, const int32_t* __restrict__ ridx,
                      const B2HistWork* __restrict__ work, int n_work, int total_chunks, int chunk_rows, int window_rows,
                      int n_groups, long long* __restrict__ hist, const B2LevelCtl* __restrict__ ctl, int log2_shards, int node_cap,
                      int64_t n_rows_total) {
  if (ctl) { n_work = ctl->hist_n_work; total_chunks = ctl->hist_total_chunks; chunk_rows = ctl->hist_chunk_rows; }
  extern __shared__ __align__(128) uint8_t smem[];
  int32_t* s_hist = reinterpret_cast<int32_t*>(smem);
  const uint32_t smem_a = (uint32_t)__cvta_generic_to_shared(smem);
  const uint32_t stage_a = smem_a + B2_GROUP_ELEMS * 4;
  const uint32_t full_a = stage_a + kStages * kStageBytes, empty_a = full_a + kStages * 8;
  const int group = blockIdx.x % n_groups, stream = blockIdx.x / n_groups, n_streams = gridDim.x / n_groups;
  if (stream >= total_chunks) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  HistTarget target; target.base = (unsigned long long*)hist; target.log2_shards = log2_shards; target.node_cap = node_cap;
  target.n_groups = n_groups;

  for (int e = threadIdx.x; e < B2_GROUP_ELEMS; e += blockDim.x) s_hist[e] = 0;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(full_a + s * 8, 1); mbar_init(empty_a + s * 8, kTmaConsumerWarps / 2); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  BlockIter it;
  it.init(work, n_work, total_chunks, chunk_rows, stream, n_streams);

  if (warp == kTmaConsumerWarps) {
    // ===================== producer warp
    // root (contiguous rows): ONE tile-mode instruction per 64-row stage (rows past the matrix end are
    // zero filled and still counted); gathered levels: 16 lanes x gather4, with the row ids of the NEXT
    // stage already in flight so the producer never waits on the ridx load.
    auto load_idx = [&](const BlockIter& b, int* idx) {
      if (b.done() || lane >= kStageRows / 4) { idx[0] = idx[1] = idx[2] = idx[3] = 0; return; }
      const int r = b.blk * kStageRows + lane * 4;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        int rr = r + j; if (rr >= b.nrows) rr = b.nrows - 1;   // tail rows repeat the last row (consumers add zero)
        idx[j] = __ldg(ridx + b.pos0 + rr);
      }
    };
    int idx[4] = {0, 0, 0, 0}, nidx[4];
    if (kGather) load_idx(it, idx);
    BlockIter nx = it;
    for (; !it.done(); it.next()) {
      if (kGather) { nx.next(); load_idx(nx, nidx); }
      const int s = it.k % kStages, use = it.k / kStages;
      mbar_wait(empty_a + s * 8, (use & 1) ^ 1);           // first use of a slot passes immediately
      if (lane == 0) mbar_expect_tx(full_a + s * 8, kStageBytes);
      __syncwarp();
      if (kGather) {
        if (lane < kStageRows / 4)
          tma_gather4(stage_a + s * kStageBytes + lane * 128, &tmap, full_a + s * 8, group * B2_GROUP_SLOTS, idx[0], idx[1], idx[2], idx[3]);
#pragma unroll
        for (int j = 0; j < 4; ++j) idx[j] = nidx[j];
      } else if (lane == 0) {
        tma_tile(stage_a + s * kStageBytes, &tmap_tile, full_a + s * 8, group * B2_GROUP_SLOTS, (int)(it.pos0 + it.blk * kStageRows));
      }
    }
    return;
  }

  // ===================== consumer warps: even warps' group (0-3) takes even blocks, 4-7 odd blocks
  const int parity_sel = warp >> 2, quarter = warp & 3;
  const int rot = lane >> 1, half = lane & 1;
  const int row_in_block = quarter * 16 + rot;
  int cur = -1, rows_in_window = 0, last_chunk = -1;
  // prefetch iterator: runs ahead to this warp's next owned block to hide the rid -> gpair latency
  BlockIter pf = it;
  while (!pf.done() && (pf.k & 1) != parity_sel) pf.next();
  auto fetch_rid = [&](const BlockIter& b) -> int64_t {
    if (b.done()) return -1;
    const int r = b.blk * kStageRows + row_in_block;
    if (r >= b.nrows) return -1;
    return kGather ? (int64_t)__ldg(ridx + b.pos0 + r) : b.pos0 + r;
  };
  int64_t rid_cur = fetch_rid(pf);
  int2 gp_cur = rid_cur >= 0 ? __ldg(gpair + rid_cur) : make_int2(0, 0);
  for (; !it.done(); it.next()) {
    if (it.chunk != last_chunk) {   // chunk boundary: uniform flush decisions for all consumer warps
      last_chunk = it.chunk;
      if (cur >= 0 && it.w != cur) {
        consumer_sync();
        for (int e = threadIdx.x; e < B2_GROUP_ELEMS; e += kTmaConsumerWarps * 32) {
          const long long v = s_hist[e];
          if (v != 0) atomicAdd(target.base + target_index(target, __ldg(&work[cur].hist_index), group, e), (unsigned long long)v);
          s_hist[e] = 0;
        }
        consumer_sync();
        rows_in_window = 0;
      } else if (cur >= 0 && rows_in_window + it.nrows > window_rows) {
        consumer_sync();
        for (int e = threadIdx.x * 4; e < B2_GROUP_ELEMS; e += kTmaConsumerWarps * 32 * 4) {
          const int4 v = *reinterpret_cast<const int4*>(s_hist + e);
          const int vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
          for (int q = 0; q < 4; ++q)
            if (vv[q] >= (1 << 30) || vv[q] <= -(1 << 30)) {
              atomicAdd(target.base + target_index(target, __ldg(&work[cur].hist_index), group, e + q), (unsigned long long)(long long)vv[q]);
              s_hist[e + q] = 0;
            }
        }
        consumer_sync();
        rows_in_window = 0;
      }
      cur = it.w;
      rows_in_window += it.nrows;
    }
    if ((it.k & 1) != parity_sel) continue;
    // this warp owns block it.k (== pf): start the loads of its NEXT owned block first
    pf.next();
    while (!pf.done() && (pf.k & 1) != parity_sel) pf.next();
    const int64_t rid_nxt = fetch_rid(pf);
    const int s = it.k % kStages, use = it.k / kStages;
    mbar_wait(full_a + s * 8, use & 1);
    RowData d;
    {
      const uint32_t a = stage_a + s * kStageBytes + row_in_block * B2_GROUP_SLOTS + half * 16;
      asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(d.bins.x), "=r"(d.bins.y), "=r"(d.bins.z), "=r"(d.bins.w) : "r"(a));
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(empty_a + s * 8);
    const int2 gp_nxt = rid_nxt >= 0 ? __ldg(gpair + rid_nxt) : make_int2(0, 0);
    d.gp = gp_cur;                       // zero for tail rows (rid_cur < 0): their duplicated bins add nothing
    accumulate_row(d, smem_a, rot, half);
    rid_cur = rid_nxt; gp_cur = gp_nxt;
  }
  if (cur >= 0) {
    consumer_sync();
    for (int e = threadIdx.x; e < B2_GROUP_ELEMS; e += kTmaConsumerWarps * 32) {
      const long long v = s_hist[e];
      if (v != 0) atomicAdd(target.base + target_index(target, __ldg(&work[cur].hist_index), group, e), (unsigned long long)v);
    }
  }
  (void)n_rows_total;
}

}  // namespace b2

extern "C" {

typedef CUresult (*B2EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// tensor map over the row-major bin matrix uint8 [n_rows][row_stride], box {32 bytes, 1 row} for tile::gather4.
// out must point to 128 bytes, 64-byte aligned.  box_rows = 1 for tile::gather4, 64 for the tile-mode root stage.
// Returns 0 on success.
int b2_make_bins_tensor_map(void* out, const uint8_t* bins, int64_t n_rows, int row_stride, int box_rows) {
  static B2EncodeFn encode = nullptr;
  if (!encode) {
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &q) != cudaSuccess || !encode) return -1;
  }
  cuuint64_t gdim[2] = {(cuuint64_t)row_stride, (cuuint64_t)(n_rows > 0 ? n_rows : 1)};
  cuuint64_t gstride[1] = {(cuuint64_t)row_stride};
  cuuint32_t box[2] = {B2_GROUP_SLOTS, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = encode((CUtensorMap*)out, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, (void*)bins, gdim, gstride, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return (int)r;
}

int b2_launch_hist_tma(const void* tmap, const void* tmap_tile, const int2* gpair, const int32_t* ridx, const B2HistWork* work, int n_work,
                       int total_chunks, int chunk_rows, int window_rows, int n_groups, long long* hist, const B2LevelCtl* ctl,
                       int log2_shards, int node_cap, int64_t n_rows_total, int num_sms, cudaStream_t stream) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(b2::hist_build_tma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, b2::kTmaSmemBytes);
    cudaFuncSetAttribute(b2::hist_build_tma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, b2::kTmaSmemBytes);
    attr_set = true;
  }
  if (!ctl && (total_chunks <= 0 || n_work <= 0)) return 0;
  if (!ctl && chunk_rows > window_rows) return (int)cudaErrorInvalidValue;
  int n_streams = (num_sms * 3) / n_groups;
  if (n_streams < 1) n_streams = 1;
  if (!ctl && n_streams > total_chunks) n_streams = total_chunks;
  dim3 grid(n_groups * n_streams), block(b2::kTmaThreads);
  const CUtensorMap* tm = (const CUtensorMap*)tmap;
  const CUtensorMap* tt = (const CUtensorMap*)tmap_tile;
  if (ridx)
    b2::hist_build_tma_kernel<true><<<grid, block, b2::kTmaSmemBytes, stream>>>(*tm, *tt, gpair, ridx, work, n_work, total_chunks, chunk_rows,
                                                                           window_rows, n_groups, hist, ctl, log2_shards, node_cap,
                                                                           n_rows_total);
  else
    b2::hist_build_tma_kernel<false><<<grid, block, b2::kTmaSmemBytes, stream>>>(*tm, *tt, gpair, ridx, work, n_work, total_chunks, chunk_rows,
                                                                            window_rows, n_groups, hist, ctl, log2_shards, node_cap,
                                                                            n_rows_total);
  return (int)cudaGetLastError();
}
}
