// split_kernel.cu -- split-gain scan over the (allreduced) level histograms.
//
// Replaces XGBoost's EvaluateSplits stage reached through xgb.train()
// (xgboost_ray/main.py:745-752).  One CTA per
// (node, feature group): 1024 threads = 32 slots (features) x 32 bin chunks of 8 bins (the scan is
// latency bound, so short per-thread load chains matter more than thread count).
// Prefix sums are exact int64, gains are IEEE fp64 with explicit round-to-nearest ops (no fma
// contraction), so every rank and the CPU oracle compute identical candidates.
#include "common.cuh"

namespace b2 {

__device__ __forceinline__ double thr_l1(double g, double a) {
  if (g > a) return __dadd_rn(g, -a);
  if (g < -a) return __dadd_rn(g, a);
  return 0.0;
}
__device__ __forceinline__ double calc_gain(double G, double H, const B2TrainParamDev& p) {
  if (H < p.min_child_weight || H <= 0.0) return 0.0;
  double t = (p.alpha == 0.0) ? G : thr_l1(G, p.alpha);
  return __ddiv_rn(__dmul_rn(t, t), __dadd_rn(H, p.lambda));
}

struct Best {
  unsigned long long key;  // (loss_chg bits << 32) | ~order ; 0 = none
  int32_t bin, default_left;
  long long lg, lh;
};

__device__ __forceinline__ void consider(Best& b, float chg, uint32_t order, int bin, int dl, long long lg, long long lh) {
  if (!(chg > 0.0f) || isinf(chg)) return;
  unsigned long long key = ((unsigned long long)__float_as_uint(chg) << 32) | (unsigned long long)(0xffffffffu - order);
  if (key > b.key) { b.key = key; b.bin = bin; b.default_left = dl; b.lg = lg; b.lh = lh; }
}

// feat_meta: per group: first feature id, size; per feature: nbins, has_missing
constexpr int kEvalChunks = 32;                 // bin chunks per feature
constexpr int kEvalBinsPerChunk = 256 / kEvalChunks;

__global__ void __launch_bounds__(32 * kEvalChunks)
eval_splits_kernel(const long long* __restrict__ level_hist, int n_groups, const B2EvalNode* __restrict__ nodes,
                   const int32_t* __restrict__ group_first, const int32_t* __restrict__ group_size,
                   const int32_t* __restrict__ nbins, const uint8_t* __restrict__ has_missing,
                   const int32_t* __restrict__ qexp, int qbits, B2TrainParamDev p, B2SplitCand* __restrict__ cands,
                   const B2LevelCtl* __restrict__ ctl, int log2_shards, int shard_rank) {
  // This rank owns sp = 32 >> log2_shards slots of every group (slot s is owned by s % shards): the
  // G*sp owned "virtual slots" of a node are covered by cpn = ceil(G*sp/32) CTAs.
  const int sp = B2_GROUP_SLOTS >> log2_shards;
  const int cpn = (n_groups * sp + 31) >> 5;
  const int node = blockIdx.x / cpn, cta_in_node = blockIdx.x % cpn;
  if (ctl && node >= ctl->n_nodes) return;
  const int s = threadIdx.x & 31, q = threadIdx.x >> 5;
  __shared__ long long cs_g[kEvalChunks][32], cs_h[kEvalChunks][32];
  __shared__ unsigned long long wkey[kEvalChunks];
  const B2EvalNode nd = nodes[node];
  // inverse scales: 2^(e - qbits)
  p.inv_scale_g = ldexp(1.0, qexp[0] - qbits);
  p.inv_scale_h = ldexp(1.0, qexp[1] - qbits);
  const int v = cta_in_node * 32 + s;                      // virtual slot
  const bool v_ok = v < n_groups * sp;
  const int group = v_ok ? v / sp : 0, sl = v_ok ? v % sp : 0;
  const int slot = (sl << log2_shards) + shard_rank;       // real slot inside the group
  const size_t slice_elems = (size_t)n_groups * 2 * B2_BINS * sp;
  const long long* hg = level_hist + (size_t)nd.hist_index * slice_elems + (size_t)(group * 2) * B2_BINS * sp + sl;
  const long long* hh = hg + (size_t)B2_BINS * sp;
  const bool active = v_ok && slot < group_size[group];
  const int f = group_first[group] + slot;
  const int nf = active ? nbins[f] : 0;
  const bool fmiss = active ? (has_missing[f] != 0) : false;

  long long sg = 0, sh = 0;
  if (active) {
#pragma unroll
    for (int i = 0; i < kEvalBinsPerChunk; ++i) {
      int b = q * kEvalBinsPerChunk + i;
      if (b < nf) { sg += hg[b * sp]; sh += hh[b * sp]; }
    }
  }
  cs_g[q][s] = sg; cs_h[q][s] = sh;
  __syncthreads();
  long long pg = 0, ph = 0, real_g = 0, real_h = 0;
#pragma unroll 8
  for (int k = 0; k < kEvalChunks; ++k) {
    long long a = cs_g[k][s], b = cs_h[k][s];
    if (k < q) { pg += a; ph += b; }
    real_g += a; real_h += b;
  }
  const long long tot_g = nd.sum_g, tot_h = nd.sum_h;
  const double G = __dmul_rn(__ll2double_rn(tot_g), p.inv_scale_g), H = __dmul_rn(__ll2double_rn(tot_h), p.inv_scale_h);
  const double root_gain = (double)nd.root_gain;
  const bool node_has_missing = fmiss && (real_g != tot_g || real_h != tot_h);
  Best best; best.key = 0; best.bin = 0; best.default_left = 0; best.lg = 0; best.lh = 0;
  if (active) {
#pragma unroll 2
    for (int i = 0; i < kEvalBinsPerChunk; ++i) {
      const int b = q * kEvalBinsPerChunk + i;
      if (b >= nf) break;
      const long long eg_excl = pg, eh_excl = ph;
      pg += hg[b * sp]; ph += hh[b * sp];
      {  // forward: left = prefix inclusive, missing -> right
        const double lh_d = __dmul_rn(__ll2double_rn(ph), p.inv_scale_h);
        if (lh_d >= p.min_child_weight) {
          const double rh_d = __dadd_rn(H, -lh_d);
          if (rh_d >= p.min_child_weight) {
            const double lg_d = __dmul_rn(__ll2double_rn(pg), p.inv_scale_g);
            const double rg_d = __dadd_rn(G, -lg_d);
            const double gain = __dadd_rn(__dadd_rn(calc_gain(lg_d, lh_d, p), calc_gain(rg_d, rh_d, p)), -root_gain);
            consider(best, __double2float_rn(gain), (uint32_t)f * 1024u + (uint32_t)b, b, 0, pg, ph);
          }
        }
      }
      if (node_has_missing) {  // backward at bin b: right' = sum of real bins >= b, missing -> left
        const long long rg_i = real_g - eg_excl, rh_i = real_h - eh_excl;
        const double rh_d = __dmul_rn(__ll2double_rn(rh_i), p.inv_scale_h);
        if (rh_d >= p.min_child_weight) {
          const double lh_d = __dadd_rn(H, -rh_d);
          if (lh_d >= p.min_child_weight) {
            const double rg_d = __dmul_rn(__ll2double_rn(rg_i), p.inv_scale_g);
            const double lg_d = __dadd_rn(G, -rg_d);
            const double gain = __dadd_rn(__dadd_rn(calc_gain(lg_d, lh_d, p), calc_gain(rg_d, rh_d, p)), -root_gain);
            consider(best, __double2float_rn(gain), (uint32_t)f * 1024u + 512u + (uint32_t)(255 - b), b - 1, 1,
                     tot_g - rg_i, tot_h - rh_i);
          }
        }
      }
    }
  }
  // block argmax on key
  unsigned long long k = best.key;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    unsigned long long other = __shfl_xor_sync(0xffffffffu, k, o);
    k = other > k ? other : k;
  }
  if (s == 0) wkey[q] = k;
  __syncthreads();
  unsigned long long kmax = 0;
#pragma unroll 8
  for (int i = 0; i < kEvalChunks; ++i) kmax = wkey[i] > kmax ? wkey[i] : kmax;
  B2SplitCand* out = cands + (size_t)node * cpn + cta_in_node;
  if (kmax == 0) {
    if (threadIdx.x == 0) {
      out->loss_chg = 0.0f; out->feature = -1; out->bin = 0; out->default_left = 0; out->left_g = 0; out->left_h = 0;
      out->order = 0xffffffffu; out->pad = 0;
    }
  } else if (best.key == kmax) {
    out->loss_chg = __uint_as_float((uint32_t)(kmax >> 32));
    out->feature = f; out->bin = best.bin; out->default_left = best.default_left;
    out->left_g = best.lg; out->left_h = best.lh; out->order = 0xffffffffu - (uint32_t)(kmax & 0xffffffffu); out->pad = 0;
  }
}

// root totals: sum of all 256 bins of slot 0 / group 0 (every row lands in exactly one bin, the
// missing sentinel included) -> nodes[0].sum_g/h and root_gain
__global__ void root_totals_kernel(const long long* __restrict__ level_hist, int n_groups, B2EvalNode* nodes,
                                   const int32_t* __restrict__ qexp, int qbits, B2TrainParamDev p, int log2_shards) {
  __shared__ long long sg[256], sh[256];
  const int sp = B2_GROUP_SLOTS >> log2_shards;
  const size_t slice_elems = (size_t)n_groups * 2 * B2_BINS * sp;
  // any owned slot works (padding slots too): every row lands in exactly one bin of every slot
  const long long* hg = level_hist + (size_t)nodes[0].hist_index * slice_elems;
  sg[threadIdx.x] = hg[(size_t)threadIdx.x * sp];
  sh[threadIdx.x] = hg[(size_t)(B2_BINS + threadIdx.x) * sp];
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) { sg[threadIdx.x] += sg[threadIdx.x + o]; sh[threadIdx.x] += sh[threadIdx.x + o]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    p.inv_scale_g = ldexp(1.0, qexp[0] - qbits);
    p.inv_scale_h = ldexp(1.0, qexp[1] - qbits);
    nodes[0].sum_g = sg[0]; nodes[0].sum_h = sh[0];
    double G = __dmul_rn(__ll2double_rn(sg[0]), p.inv_scale_g), H = __dmul_rn(__ll2double_rn(sh[0]), p.inv_scale_h);
    nodes[0].root_gain = __double2float_rn(calc_gain(G, H, p));
  }
}

}  // namespace b2

extern "C" {
int b2_launch_eval_splits(const long long* level_hist, int n_groups, const B2EvalNode* nodes, int n_nodes,
                          const int32_t* group_first, const int32_t* group_size, const int32_t* nbins,
                          const uint8_t* has_missing, const int32_t* qexp, int qbits, B2TrainParamDev p,
                          B2SplitCand* cands, const B2LevelCtl* ctl, int log2_shards, int shard_rank, cudaStream_t stream) {
  if (n_nodes <= 0) return 0;   // with ctl: n_nodes is the upper bound of the level
  const int sp = B2_GROUP_SLOTS >> log2_shards, cpn = (n_groups * sp + 31) >> 5;
  b2::eval_splits_kernel<<<n_nodes * cpn, 32 * b2::kEvalChunks, 0, stream>>>(level_hist, n_groups, nodes, group_first, group_size,
                                                                         nbins, has_missing, qexp, qbits, p, cands, ctl,
                                                                         log2_shards, shard_rank);
  return (int)cudaGetLastError();
}
int b2_launch_root_totals(const long long* level_hist, int n_groups, B2EvalNode* nodes, const int32_t* qexp, int qbits,
                          B2TrainParamDev p, int log2_shards, cudaStream_t stream) {
  b2::root_totals_kernel<<<1, 256, 0, stream>>>(level_hist, n_groups, nodes, qexp, qbits, p, log2_shards);
  return (int)cudaGetLastError();
}
}
