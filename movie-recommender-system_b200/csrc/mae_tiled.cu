// mae_tiled.cu -- fused baseline prediction + |error| reduction (P:69-86 over P:217-236) on an item-tiled test layout.
//
// The generic kernel (baseline.cu predict_mae_kernel) gathers avg[u] and dev[i] from global memory: scattered 8-byte
// gathers cost one L1 tag cycle per distinct line (capture A: the test pass was bound by that, not by HBM).  Here the
// test entries are grouped by item tile (kMaeTileItems items): a CTA stages the tile's item deviations in shared
// memory (64 KB) and its warps stream contiguous runs of 32-entry rows through private TMA rings (cp.async.bulk +
// mbarrier), exactly like the item pass of the fit (tiled.cu).  Inside a tile the entries stay in (user, item) order,
// so the remaining global gather -- the user average -- touches one or two lines per warp.
// An entry is one 8-byte word: int32 user | 16-bit item id local to the tile | half-star code (0xFF = padding).
#include <cub/cub.cuh>

#include <algorithm>
#include <vector>

#include "common.cuh"
#include "tma.cuh"

namespace mrs {
namespace {

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__global__ void mae_keys_kernel(const int32_t* __restrict__ items, int64_t n, uint16_t* __restrict__ key, int32_t* __restrict__ pos) {
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x) {
    key[p] = (uint16_t)(items[p] / kMaeTileItems);
    pos[p] = (int32_t)p;
  }
}

// sorted tile keys -> first position of every tile
__global__ void tile_ptr_kernel(const uint16_t* __restrict__ key, int64_t n, int32_t n_tiles, int32_t* __restrict__ ptr) {
  for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < n; q += (int64_t)gridDim.x * blockDim.x) {
    const int32_t t = key[q];
    const int32_t prev = q ? (int32_t)key[q - 1] : -1;
    for (int32_t s = prev + 1; s <= t; ++s) ptr[s] = (int32_t)q;
    if (q == n - 1)
      for (int32_t s = t + 1; s <= n_tiles; ++s) ptr[s] = (int32_t)n;
  }
}

__global__ void mae_fill_kernel(uint2* __restrict__ entry, int64_t n_slots) {
  for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < n_slots; q += (int64_t)gridDim.x * blockDim.x)
    entry[q] = make_uint2(0u, 0xffu << 16);  // padding slot
}

__global__ void mae_scatter_kernel(const uint16_t* __restrict__ key, const int32_t* __restrict__ perm, const int32_t* __restrict__ tile_ptr,
                                   const int32_t* __restrict__ tile_row_ptr, const int32_t* __restrict__ users,
                                   const int32_t* __restrict__ items, const uint8_t* __restrict__ codes, int64_t n, uint2* __restrict__ entry) {
  for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < n; q += (int64_t)gridDim.x * blockDim.x) {
    const int32_t t = key[q], p = perm[q];
    const int64_t dst = ((int64_t)tile_row_ptr[t] << 5) + (q - tile_ptr[t]);
    entry[dst] = make_uint2((uint32_t)users[p], (uint32_t)(items[p] - t * kMaeTileItems) | ((uint32_t)codes[p] << 16));
  }
}

constexpr int kMaeThreads = 1024;
#ifndef MRS_K3_ROWS
#define MRS_K3_ROWS 8
#endif
#ifndef MRS_K3_STAGES
#define MRS_K3_STAGES 2
#endif
constexpr int kMaeRows = MRS_K3_ROWS;      // rows (32 entries of 8 bytes = 256 B) per ring stage
constexpr int kMaeStages = MRS_K3_STAGES;
constexpr size_t kMaeSmem = (size_t)kMaeTileItems * 8 + (size_t)(kMaeThreads / 32) * kMaeStages * kMaeRows * 256 + (size_t)(kMaeThreads / 32) * kMaeStages * 8;

// FOLD: the test pass finishes the fit itself (single-GPU closure MeanAbsoluteErrorSpark(baselinePredictorSpark(train), test),
// mrs_fit_mae_async): the item pass' integer accumulators go straight into the tile's deviation table in shared memory, so
// the finishing kernel K2b (5.7 us on the critical path between the item pass and this kernel, tools/timeline.py) is gone;
// the model's arrays are written by all CTAs together (an equal share of the item ids each, independent of the tiling).
struct FoldArgs {
  long long* xdev_fix;              // [2][n_items] per-item deviation sums on the 2^-40 grid: buffer *parity is complete once K2 has finished
  unsigned int* parity;             // which buffer this pass uses; the other one is re-armed here, the last block flips the parity
  const int32_t* icolp;             // [n_items+1] train column pointer (rating counts)
  unsigned long long* k1_part;      // [4] sum of all train codes (K1) + the hand-over counts of the pass (common.cuh flag_wait)
  int32_t n_k2_ctas;                // > 0: wait for that many item-pass CTAs to count themselves off instead of for the grid
  int32_t deliver;                  // FOLD == 2: 1 = this kernel's CTAs deliver the partial sums themselves, 0 = the fit's push kernel did
  double n_fit;                     // number of train ratings
  double* idevavg;                  // model outputs
  double* xbuf;
  double* gavg;
  uint32_t* usum;                   // consumed by K2: re-armed here for the next pass' K1
  // FOLD == 2 (sharded run, mrs_fit_mae_push_async): the per-item exchange of the fit happens in this kernel's prologue
  const int32_t* known;             // [K] items that occur on some rank, ascending (compact slot j <-> item known[j])
  const int32_t* item_slot;         // [n_items] inverse: compact slot of an item, -1 if it occurs on no rank
  int32_t K;
  PushDev big;                      // exchange handle of the per-item partial sums (2K + 2 doubles per rank)
};

// one CTA = (item tile, share of the tile's rows); 32 warps, one CTA per SM
template <int FOLD>  // 0: the model is finished; 1: finish it here (single GPU); 2: exchange across ranks + finish it here
__global__ void __launch_bounds__(kMaeThreads, 1) predict_mae_tiled_kernel(const uint2* __restrict__ entry, const int32_t* __restrict__ tile_row_ptr,
                                                                          const int3* __restrict__ cta_desc, int32_t n_users, int32_t n_items,
                                                                          const double* __restrict__ uavg, const double* __restrict__ idevavg,
                                                                          const double* __restrict__ gavg_p, double n_total,
                                                                          double* __restrict__ part, unsigned int* __restrict__ counter,
                                                                          double* __restrict__ out2, unsigned long long* __restrict__ tl,
                                                                          int use_push, const PushDev x, const FoldArgs f) {
  tl_begin(tl, 3);
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* s_dev = reinterpret_cast<double*>(smem_raw);                                     // [kMaeTileItems]
  uint2* s_ring = reinterpret_cast<uint2*>(smem_raw + (size_t)kMaeTileItems * 8);          // [warps][stages][rows*32]
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(smem_raw + (size_t)kMaeTileItems * 8 + (size_t)(kMaeThreads / 32) * kMaeStages * kMaeRows * 256);
  __shared__ double sh[kMaeThreads / 32];
  __shared__ bool is_last;
  const int3 cd = cta_desc[blockIdx.x];
  const int32_t tile = cd.x, share = cd.y, ctas_per_tile = cd.z;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  constexpr int32_t wpb = kMaeThreads >> 5;
  uint64_t* bar = s_bar + wid * kMaeStages;
  uint2* ring = s_ring + (size_t)wid * kMaeStages * kMaeRows * 32;
  if (lane == 0) {
#pragma unroll
    for (int st = 0; st < kMaeStages; ++st) tma::mbar_init(bar + st, 1);
    tma::fence_barrier_init();
  }
  __syncwarp();
  // ---- this warp's rows: an equal share of the tile's rows
  const int32_t ra = tile_row_ptr[tile], rb = tile_row_ptr[tile + 1];
  const int32_t nw = ctas_per_tile * wpb, w = share * wpb + wid;
  const int32_t r0 = ra + (int32_t)(((int64_t)(rb - ra) * w) / nw);
  const int32_t r_end = ra + (int32_t)(((int64_t)(rb - ra) * (w + 1)) / nw);
  const int32_t n_chunks = (r_end - r0 + kMaeRows - 1) / kMaeRows;
  if (lane == 0) {
#pragma unroll
    for (int st = 0; st < kMaeStages; ++st) {
      if (st < n_chunks) {
        const int32_t rr = r0 + st * kMaeRows;
        const uint32_t bytes = (uint32_t)min(kMaeRows, r_end - rr) * 256u;
        tma::mbar_arrive_expect_tx(bar + st, bytes);
        tma::bulk_g2s(ring + st * kMaeRows * 32, entry + ((int64_t)rr << 5), bytes, bar + st);
      }
    }
  }
  // ---- the tile's item deviations (unknown item -> 0.0, P:226-227)
  const int32_t i0 = tile * kMaeTileItems;
  constexpr int kPerTile = kMaeTileItems / kMaeThreads;
  int32_t cnt_of[FOLD == 1 ? kPerTile : 1];
  if (FOLD == 1) {  // rating counts of the tile's items: layout data, fetched before the wait
#pragma unroll
    for (int k = 0; k < kPerTile; ++k) {
      const int32_t i = i0 + k * kMaeThreads + threadIdx.x;
      cnt_of[k] = (i < n_items) ? __ldg(f.icolp + i + 1) - __ldg(f.icolp + i) : 0;
    }
  }
  // barriers, partition, the first ring stages and the counts overlapped the end of the fit; its outputs are complete from here on
  if (FOLD && f.n_k2_ctas > 0) flag_wait(f.k1_part + 2, (unsigned long long)f.n_k2_ctas); else pdl_wait();
  double gavg;
  bool shard_ok = true;
  unsigned long long big_epoch = 0;
  if (FOLD == 2) {
    // ---- sharded closure: every CTA delivers an equal share of this rank's per-item partial sums into every rank's
    // receive buffer (NVLink stores), the last CTA to finish raises this rank's flag everywhere, then every CTA waits for
    // all ranks' flags, builds its tile's deviations from the deliveries in its OWN memory (rank order: bit-identical
    // totals on every rank) and writes its share of the model's arrays.  No separate exchange or finishing kernel.
    constexpr double kInvFix = 1.0 / 1099511627776.0;  // 2^-40
    __shared__ int s_last_cta;
    const PushDev& x2 = f.big;
    big_epoch = *x2.epoch + 1;
    const int xpar = (int)(big_epoch & 1);
    const unsigned int par = *f.parity & 1u;
    const long long* __restrict__ fix = f.xdev_fix + (size_t)par * n_items;
    long long* __restrict__ fix_other = f.xdev_fix + (size_t)(par ^ 1u) * n_items;
    const int32_t K = f.K;
    const int32_t perK = (K + gridDim.x - 1) / gridDim.x;
    const int32_t jlo = blockIdx.x * perK, jhi = min(K, jlo + perK);
    if (f.deliver) {
    for (int32_t j = jlo + threadIdx.x; j < jhi; j += kMaeThreads) {
      const int32_t i = __ldg(f.known + j);
      const double ds = (double)__ldcg(fix + i) * kInvFix;
      fix_other[i] = 0;  // re-arm the buffer of the NEXT pass
      const double cnt = (double)(__ldg(f.icolp + i + 1) - __ldg(f.icolp + i));
      for (int q = 0; q < x2.world; ++q) {
        double* slot = push_slot(x2, push_peer(x2, q), xpar, x2.rank);
        slot[j] = ds;
        slot[(size_t)K + j] = cnt;
      }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
      const double gs = 0.5 * (double)__ldcg(f.k1_part);
      for (int p = 0; p < x2.world; ++p) {
        double* slot = push_slot(x2, p, xpar, x2.rank);
        slot[2 * (size_t)K] = gs;
        slot[2 * (size_t)K + 1] = f.n_fit;
      }
    }
    for (int32_t u = blockIdx.x * kMaeThreads + threadIdx.x; u < n_users; u += gridDim.x * kMaeThreads) f.usum[u] = 0;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
      const int last = (atomicAdd(x2.done, 1u) + 1u == gridDim.x);
      if (last) *x2.done = 0;
      __threadfence();
      s_last_cta = last;
    }
    __syncthreads();
    if (s_last_cta && (int)threadIdx.x < x2.world) {
      __threadfence_system();
      push_flag_raise(x2, threadIdx.x, big_epoch);
    }
    }  // (f.deliver)
    int good = 1;
    if ((int)threadIdx.x < x2.world) good = push_flag_wait(x2, threadIdx.x, big_epoch) ? 1 : 0;  // every rank has delivered
    shard_ok = __syncthreads_and(good) != 0;
    const double bad = nan("");
    {
      double gs = 0.0, gc = 0.0;
      for (int p = 0; p < x2.world; ++p) {  // rank order
        const double* slot = push_slot(x2, x2.rank, xpar, p);
        gs += slot[2 * (size_t)K];
        gc += slot[2 * (size_t)K + 1];
      }
      gavg = shard_ok ? (gc > 0.0 ? gs / gc : 0.0) : bad;
      if (blockIdx.x == 0 && threadIdx.x == 0) {
        f.xbuf[2 * (size_t)n_items] = shard_ok ? gs : bad;
        f.xbuf[2 * (size_t)n_items + 1] = shard_ok ? gc : bad;
        f.gavg[0] = gavg;
      }
    }
    constexpr int kPer = kMaeTileItems / kMaeThreads;
#pragma unroll 2
    for (int k = 0; k < kPer; ++k) {
      const int32_t i = i0 + k * kMaeThreads + threadIdx.x;
      const int32_t j = (i < n_items) ? __ldg(f.item_slot + i) : -1;
      double ds = 0.0, cnt = 0.0;
      if (j >= 0) {
        for (int p = 0; p < x2.world; ++p) {
          const double* slot = push_slot(x2, x2.rank, xpar, p);
          ds += slot[j];
          cnt += slot[(size_t)K + j];
        }
      }
      s_dev[k * kMaeThreads + threadIdx.x] = shard_ok ? (cnt > 0.0 ? ds / cnt : 0.0) : bad;
    }
    for (int32_t j = jlo + threadIdx.x; j < jhi; j += kMaeThreads) {  // this CTA's share of the model's arrays
      double ds = 0.0, cnt = 0.0;
      for (int p = 0; p < x2.world; ++p) {
        const double* slot = push_slot(x2, x2.rank, xpar, p);
        ds += slot[j];
        cnt += slot[(size_t)K + j];
      }
      const int32_t i = __ldg(f.known + j);
      f.xbuf[i] = shard_ok ? ds : bad;
      f.xbuf[(size_t)n_items + i] = shard_ok ? cnt : bad;
      f.idevavg[i] = shard_ok ? (cnt > 0.0 ? ds / cnt : 0.0) : bad;
    }
  } else if (FOLD == 1) {
    constexpr double kInvFix = 1.0 / 1099511627776.0;  // 2^-40
    constexpr int kPer = kMaeTileItems / kMaeThreads;
    const unsigned int par = *f.parity & 1u;
    const long long* __restrict__ fix = f.xdev_fix + (size_t)par * n_items;
    long long* __restrict__ fix_other = f.xdev_fix + (size_t)(par ^ 1u) * n_items;
    long long fx[kPer];
#pragma unroll
    for (int k = 0; k < kPer; ++k) {  // all loads of the tile's 8 items per thread go out together
      const int32_t i = i0 + k * kMaeThreads + threadIdx.x;
      fx[k] = (i < n_items) ? __ldcg(fix + i) : 0;
    }
    const double gs = 0.5 * (double)__ldcg(f.k1_part);  // integer sum of codes: exact
    gavg = f.n_fit > 0.0 ? gs / f.n_fit : 0.0;           // P:18 mean of an empty Seq is 0.0
#pragma unroll
    for (int k = 0; k < kPer; ++k) {
      const double cnt = (double)cnt_of[k];
      s_dev[k * kMaeThreads + threadIdx.x] = cnt > 0.0 ? ((double)fx[k] * kInvFix) / cnt : 0.0;  // P:185; unknown item -> 0.0 (P:197)
    }
    // the model's arrays: every CTA an equal share of the item ids
    const int32_t per = (n_items + gridDim.x - 1) / gridDim.x;
    const int32_t lo = blockIdx.x * per, hi = min(n_items, lo + per);
    for (int32_t i = lo + threadIdx.x; i < hi; i += kMaeThreads) {
      const double ds = (double)__ldcg(fix + i) * kInvFix;
      fix_other[i] = 0;  // re-arm the buffer of the NEXT pass (its last reader, the previous pass, is long done)
      const double cnt = (double)(__ldg(f.icolp + i + 1) - __ldg(f.icolp + i));
      f.xbuf[i] = ds;
      f.xbuf[(size_t)n_items + i] = cnt;
      f.idevavg[i] = cnt > 0.0 ? ds / cnt : 0.0;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
      f.xbuf[2 * (size_t)n_items] = gs;
      f.xbuf[2 * (size_t)n_items + 1] = f.n_fit;
      f.gavg[0] = gavg;
    }
    for (int32_t u = blockIdx.x * kMaeThreads + threadIdx.x; u < n_users; u += gridDim.x * kMaeThreads) f.usum[u] = 0;
  } else {
#pragma unroll 4
    for (int32_t x = threadIdx.x; x < kMaeTileItems; x += kMaeThreads) {
      const int32_t i = i0 + x;
      s_dev[x] = (i < n_items) ? __ldg(idevavg + i) : 0.0;
    }
    gavg = gavg_p[0];
  }
  __syncthreads();
  tl_cta(tl, 2);

  double acc = 0.0;
  const double* ua_base = uavg;
  asm volatile("" : "+l"(ua_base));  // keep the table base in registers (ptxas otherwise reloads it from the constant bank per gather)
  for (int32_t c = 0; c < n_chunks; ++c) {
    const int st = c % kMaeStages;
    const int32_t r = r0 + c * kMaeRows;
    const int32_t nrows = min(kMaeRows, r_end - r);
    tma::mbar_wait(bar + st, (uint32_t)(c / kMaeStages) & 1u);
    const uint2* rp = ring + st * kMaeRows * 32 + lane;
    uint2 ev[kMaeRows];
    if (nrows == kMaeRows) {  // whole stage (all but a warp's last chunk): no per-row predicates
#pragma unroll
      for (int k = 0; k < kMaeRows; ++k) ev[k] = rp[k * 32];
    } else {
#pragma unroll
      for (int k = 0; k < kMaeRows; ++k) ev[k] = (k < nrows) ? rp[k * 32] : make_uint2(0u, 0xffu << 16);
    }
    __syncwarp();
    if (lane == 0 && c + kMaeStages < n_chunks) {
      const int32_t rr = r + kMaeStages * kMaeRows;
      const uint32_t bytes = (uint32_t)min(kMaeRows, r_end - rr) * 256u;
      tma::mbar_arrive_expect_tx(bar + st, bytes);
      tma::bulk_g2s(ring + st * kMaeRows * 32, entry + ((int64_t)rr << 5), bytes, bar + st);
    }
    double ua[kMaeRows];
#pragma unroll
    for (int k = 0; k < kMaeRows; ++k)  // the user averages: all gathers of the stage go out together (unknown user: -1.0)
      ua[k] = (ev[k].x < (uint32_t)n_users) ? __ldg(ua_base + ev[k].x) : -1.0;
#pragma unroll
    for (int k = 0; k < kMaeRows; ++k) {  // branch free: selects only, the eight rows overlap in the pipeline
      const uint32_t code = ev[k].y >> 16;
      const double d = s_dev[ev[k].y & 0xffffu];
      const double a = ua[k];
      const double x = __dadd_rn(a, d);                                      // P:229: the branch is taken on the rounded sum
      const double sc = x > a ? 5.0 - a : (x < a ? a - 1.0 : 1.0);           // scale(x, a), P:57-61
      const double pb = __dadd_rn(a, __dmul_rn(d, sc));                      // no FMA: the JVM multiplies, then adds
      const double p = a < 0.0 ? gavg : pb;                                  // unknown user -> global average (P:222-224)
      const double err = fabs(fma((double)code, 0.5, -p));                   // |r - p|, P:71
      acc += (code != 0xffu) ? err : 0.0;                                    // 0xFF = padding slot
    }
  }
  acc = warp_sum(acc);
  if (lane == 0) sh[wid] = acc;
  __syncthreads();
  tl_cta(tl, 3);
  if (threadIdx.x < 32) {
    double t = (threadIdx.x < wpb) ? sh[threadIdx.x] : 0.0;
    t = warp_sum(t);
    if (threadIdx.x == 0) {
      part[blockIdx.x] = t;
      __threadfence();
      is_last = (atomicAdd(counter, 1u) == gridDim.x - 1);
    }
  }
  __syncthreads();
  if (is_last) {  // per-block partials are combined in block order by the last block: deterministic
    __threadfence();
    double t = 0.0;
    for (int b = threadIdx.x; b < (int)gridDim.x; b += blockDim.x) t += __ldcg(&part[b]);
    t = warp_sum(t);
    if (lane == 0) sh[wid] = t;
    __syncthreads();
    if (!use_push) {
      if (threadIdx.x == 0) {
        double s = 0.0;
        for (int k = 0; k < wpb; ++k) s += sh[k];
        out2[0] = s;
        out2[1] = n_total;
        *counter = 0;
        if (FOLD) {
          f.k1_part[0] = 0; f.k1_part[1] = 0; f.k1_part[2] = 0;  // every CTA has read the code sum and passed its wait: re-arm for the next pass
          *f.parity ^= 1u;   // ... and the accumulators: the next pass uses the buffer re-armed above
          if (FOLD == 2) *f.big.epoch = big_epoch;  // every CTA is done with the deliveries of this exchange
        }
      }
    } else {
      // sharded run, fused exchange of {sum |err|, n}: this last block delivers the rank's pair into every rank's receive
      // slot (NVLink stores), raises the flags, waits for the other ranks' pairs and adds them in rank order
      const unsigned long long epoch = *x.epoch + 1;
      const int parity = (int)(epoch & 1);
      if (threadIdx.x == 0) {
        double s = 0.0;
        for (int k = 0; k < wpb; ++k) s += sh[k];
        for (int p = 0; p < x.world; ++p) {
          double* slot = push_slot(x, p, parity, x.rank);
          slot[0] = s;
          slot[1] = n_total;
        }
        __threadfence_system();
        for (int p = 0; p < x.world; ++p) push_flag_raise(x, p, epoch);
      }
      __syncthreads();
      int good = 1;
      if ((int)threadIdx.x < x.world) good = push_flag_wait(x, threadIdx.x, epoch) ? 1 : 0;
      const bool ok = __syncthreads_and(good) != 0;
      if (threadIdx.x == 0) {
        double s = 0.0, c = 0.0;
        for (int p = 0; p < x.world; ++p) {
          const double* slot = push_slot(x, x.rank, parity, p);
          s += slot[0];
          c += slot[1];
        }
        out2[0] = (ok && shard_ok) ? s : nan("");  // a peer never delivered: the MAE becomes NaN
        out2[1] = (ok && shard_ok) ? c : nan("");
        *x.epoch = epoch;
        *counter = 0;
        if (FOLD) {
          f.k1_part[0] = 0; f.k1_part[1] = 0; f.k1_part[2] = 0;
          *f.parity ^= 1u;
          if (FOLD == 2) *f.big.epoch = big_epoch;
        }
      }
    }
  }
  tl_end(tl, 3);
}

int grid_for(int64_t n, int block, int sm_count) {
  return (int)std::max<int64_t>(1, std::min<int64_t>((n + block - 1) / block, (int64_t)sm_count * 16));
}

}  // namespace

void free_mae_layout(const mrs_ratings* T) {
  auto& L = T->ml;
  dev_free(L.entry); dev_free(L.tile_row_ptr); dev_free(L.cta_desc);
  L = mrs_ratings::mae_layout();
}

int32_t build_mae_layout(const mrs_ratings* T) {
  auto& L = T->ml;
  if (L.built) return MRS_OK;
  MRS_REQUIRE(T->value_kind == kValueCode && T->n > 0, MRS_ERR_INVALID, "item-tiled test layout needs a non-empty set of half-star codes");
  mrs_engine* e = T->eng;
  cudaStream_t st = e->stream;
  const int64_t n = T->n;
  const int32_t NT = (T->n_items + kMaeTileItems - 1) / kMaeTileItems;
  MRS_REQUIRE(NT < 65536, MRS_ERR_UNSUPPORTED, "too many item tiles (%d)", NT);
  L.n_tiles = NT;
  const int grid = grid_for(n, 256, e->sm_count);
  uint16_t *k_in = nullptr, *k_out = nullptr;
  int32_t *p_in = nullptr, *perm = nullptr, *tile_ptr = nullptr;
  MRS_TRY(dev_alloc(&k_in, (size_t)n)); MRS_TRY(dev_alloc(&k_out, (size_t)n));
  MRS_TRY(dev_alloc(&p_in, (size_t)n)); MRS_TRY(dev_alloc(&perm, (size_t)n));
  MRS_TRY(dev_alloc(&tile_ptr, (size_t)NT + 1));
  MRS_TRY(dev_alloc(&L.tile_row_ptr, (size_t)NT + 1));
  mae_keys_kernel<<<grid, 256, 0, st>>>(T->ucol, n, k_in, p_in);
  int tbits = 1;
  while ((1 << tbits) < NT) ++tbits;
  size_t tmp = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, tmp, k_in, k_out, p_in, perm, (int)n, 0, tbits, st);
  MRS_TRY(ensure_scratch(e, tmp));
  cub::DeviceRadixSort::SortPairs(e->scratch, tmp, k_in, k_out, p_in, perm, (int)n, 0, tbits, st);  // stable: (user, item) order kept
  tile_ptr_kernel<<<grid, 256, 0, st>>>(k_out, n, NT, tile_ptr);
  std::vector<int32_t> h_ptr((size_t)NT + 1), h_rows((size_t)NT + 1, 0);
  MRS_CUDA(cudaMemcpyAsync(h_ptr.data(), tile_ptr, sizeof(int32_t) * ((size_t)NT + 1), cudaMemcpyDeviceToHost, st));
  MRS_CUDA(cudaStreamSynchronize(st));
  for (int32_t t = 0; t < NT; ++t) h_rows[t + 1] = h_rows[t] + (h_ptr[t + 1] - h_ptr[t] + 31) / 32;
  L.n_rows = h_rows[NT];
  MRS_TRY(dev_alloc(&L.entry, (size_t)L.n_rows * 32 + 32));
  MRS_CUDA(cudaMemcpyAsync(L.tile_row_ptr, h_rows.data(), sizeof(int32_t) * ((size_t)NT + 1), cudaMemcpyHostToDevice, st));
  mae_fill_kernel<<<grid_for(L.n_rows * 32, 256, e->sm_count), 256, 0, st>>>(L.entry, L.n_rows * 32);
  mae_scatter_kernel<<<grid, 256, 0, st>>>(k_out, perm, tile_ptr, L.tile_row_ptr, T->coo_u, T->ucol, (const uint8_t*)T->uval, n, L.entry);
  count_launch(6);
  MRS_CUDA(cudaGetLastError());
  {  // CTAs of the test pass, dealt out to the item tiles in proportion to their rows
    std::vector<int64_t> cost((size_t)NT);
    for (int32_t t = 0; t < NT; ++t) cost[(size_t)t] = h_rows[(size_t)t + 1] - h_rows[(size_t)t];
    const std::vector<int3> desc = deal_ctas(cost, e->sm_count);
    L.n_ctas = (int32_t)desc.size();
    MRS_TRY(dev_alloc(&L.cta_desc, std::max<size_t>(1, desc.size())));
    MRS_CUDA(cudaMemcpyAsync(L.cta_desc, desc.data(), sizeof(int3) * desc.size(), cudaMemcpyHostToDevice, st));
    MRS_CUDA(cudaStreamSynchronize(st));  // `desc` must outlive the copy
  }
  MRS_CUDA(cudaStreamSynchronize(st));
  dev_free(k_in); dev_free(k_out); dev_free(p_in); dev_free(perm); dev_free(tile_ptr);
  L.built = true;
  return MRS_OK;
}

int32_t launch_mae_tiled_baseline(const mrs_model* m, const mrs_ratings* T, double* d_out2, const PushDev* push, bool fold, const PushDev* big,
                                  bool deliver) {
  MRS_TRY(build_mae_layout(T));
  const auto& L = T->ml;
  mrs_engine* e = m->eng;
  if (!(e->smem_attr_done & 2u)) {
    MRS_CUDA(cudaFuncSetAttribute(predict_mae_tiled_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaeSmem));
    MRS_CUDA(cudaFuncSetAttribute(predict_mae_tiled_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaeSmem));
    MRS_CUDA(cudaFuncSetAttribute(predict_mae_tiled_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaeSmem));
    e->smem_attr_done |= 2u;
  }
  const int32_t grid = L.n_ctas;
  MRS_REQUIRE(grid > 0 && grid <= m->mae_part_cap, MRS_ERR_UNSUPPORTED, "test set needs %d CTAs, more than the %d partial slots of the model", grid,
              m->mae_part_cap);
  FoldArgs f = {};
  if (fold) {
    const mrs_ratings* R = m->train;
    f.xdev_fix = m->xdev_fix; f.icolp = R->icolp; f.k1_part = m->k1_part; f.n_fit = (double)R->n;
    f.n_k2_ctas = m->flag_sync ? R->tl.n_ctas : 0;
    f.idevavg = m->idevavg; f.xbuf = m->xbuf; f.gavg = m->gavg; f.usum = m->usum; f.parity = m->counters + 4;
    if (big) {  // sharded closure: both exchanges inside this kernel
      MRS_REQUIRE(push && grid <= e->sm_count, MRS_ERR_UNSUPPORTED, "sharded closure: the test pass must be one wave (%d CTAs)", grid);
      f.known = m->slot_of_item; f.item_slot = m->item_slot; f.K = m->n_slots_known; f.big = *big; f.deliver = deliver ? 1 : 0;
      MRS_CUDA(launch_pdl(predict_mae_tiled_kernel<2>, dim3(grid), dim3(kMaeThreads), kMaeSmem, e->stream, L.entry, L.tile_row_ptr, L.cta_desc, m->n_users,
                          m->n_items, m->uavg, m->idevavg, m->gavg, (double)T->n, m->mae_part, m->counters, d_out2, e->d_timeline, 1, *push, f));
    } else {
      MRS_CUDA(launch_pdl(predict_mae_tiled_kernel<1>, dim3(grid), dim3(kMaeThreads), kMaeSmem, e->stream, L.entry, L.tile_row_ptr, L.cta_desc, m->n_users,
                          m->n_items, m->uavg, m->idevavg, m->gavg, (double)T->n, m->mae_part, m->counters, d_out2, e->d_timeline, 0, PushDev{}, f));
    }
  } else {
    MRS_CUDA(launch_pdl(predict_mae_tiled_kernel<0>, dim3(grid), dim3(kMaeThreads), kMaeSmem, e->stream, L.entry, L.tile_row_ptr, L.cta_desc, m->n_users,
                        m->n_items, m->uavg, m->idevavg, m->gavg, (double)T->n, m->mae_part, m->counters, d_out2, e->d_timeline, push ? 1 : 0,
                        push ? *push : PushDev{}, f));
  }
  mark(e, "predict_mae_tiled");
  MRS_CUDA(cudaGetLastError());
  return MRS_OK;
}

}  // namespace mrs
