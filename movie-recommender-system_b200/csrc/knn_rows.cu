// knn_rows.cu -- the row-block path of getNeighbors / getSimilarity / predictor (P:596-649, P:489-586) for user counts
// whose similarity matrix does not fit (ml-25m shape: 162,541 users -> 211 GB as fp64), and for a rank that owns only a
// range of the rows (config 5: similarity row-blocks sharded across the GPUs of a box).
//
// Nothing n_known x n_known is ever held -- not even one row leaves the SM: a CTA produces the row of its user range by
// range in shared memory, reduces it on the fly to the first k neighbours in the order (similarity desc, user id asc)
// and writes only that list; prediction walks the list of u and looks the item up in every neighbour's row.
//
//   R1  similarity rows, ITEM-DRIVEN: the work is sum_i cnt_i^2 pair products (1.3e11 at ml-25m shape) instead of the
//       n_users x nnz (3.3e12) of the dense-staged kernel in knn.cu.  A CTA owns a row user u and visits the ranges of 16,384
//       compact user indices one after the other; each of its 32 warps owns 512 of them with private fp64 accumulators.  The warp walks
//       the items of u in ascending order and, for each, the slice of that item's column (users ascending) that falls into
//       its 512 indices -- a per-column segment table gives the slice bounds.  acc[v] = acc[v] + r~(u,i) * r~(v,i) with
//       non-fused, correctly rounded ops: for every (u, v) exactly the oracle's sequence (ascending item id over the
//       intersection), so the similarity bits -- and with them the neighbour order -- are the oracle's.
//   R2  selection, same CTA: after each range the 16,384 similarities pass a filter that keeps candidates beating the
//       running k-th best (sim, id) in a shared-memory buffer, compacted by a bitonic sort when it fills; exact, ties by
//       ascending id.  (A first version wrote the rows to HBM for a separate selection kernel: 2 x 211 GB of traffic at
//       ml-25m shape that also evicted the hot columns from L2 -- 45 % hit rate in profiles/r01_ncu_full_raw_knnrows.csv.)
//   R3  prediction / MAE from the lists: one warp per (u, i), lanes over the first k neighbours, binary search of i in the
//       neighbour's row (items ascending).
#include <algorithm>
#include <climits>
#include <cmath>
#include <vector>

#include "common.cuh"

namespace mrs {
namespace {

#ifndef MRS_ROWS_SUB
#define MRS_ROWS_SUB 512
#endif
#ifndef MRS_ROWS_WARPS
#define MRS_ROWS_WARPS 32
#endif
constexpr int kRowsSub = MRS_ROWS_SUB;      // compact user indices per warp (4 KB of fp64 accumulators at 512)
constexpr int kRowsWarps = MRS_ROWS_WARPS;  // warps per CTA -> 16,384 indices, 128 KB of shared memory, one CTA per SM
static_assert(kRowsSub * kRowsWarps == 16384, "a range is 16,384 compact indices (128 KB of accumulators)");
constexpr int kRowsDepth = 4;     // column slices requested ahead of the one being accumulated
constexpr int kSelTile = 2048;    // row elements examined between two checks of the candidate buffer
constexpr int kSelCap = 4096;     // candidate buffer (a tile can add at most kSelTile to at most kSelCap - kSelTile)
constexpr int kMaxListK = 1024;   // neighbours kept per user at most

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ bool before(double ka, int32_t ia, double kb, int32_t ib) {
  return (ka > kb) || (ka == kb && ia < ib);  // P:610 sortBy(-sim) is stable over ascending user ids
}

// ---------------- layout: compact user index of every CSC entry, row lengths, per-column segment table ----------------
__global__ void rows_ccv_kernel(const int32_t* __restrict__ irow, const int32_t* __restrict__ cidx, int64_t n, int32_t* __restrict__ ccv) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < n; p += stride) ccv[p] = cidx[irow[p]];
}

__global__ void rows_seg_kernel(const int32_t* __restrict__ icolp, const int32_t* __restrict__ ccv, int32_t n_items, int32_t n_sub,
                                int32_t* __restrict__ seg) {
  const int64_t total = (int64_t)n_items * (n_sub + 1);
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += stride) {
    const int32_t i = (int32_t)(t / (n_sub + 1)), b = (int32_t)(t % (n_sub + 1));
    int32_t lo = icolp[i], hi = icolp[i + 1];
    const int32_t want = b * kRowsSub;
    while (lo < hi) {  // first entry of the column with compact index >= want
      const int32_t mid = (lo + hi) >> 1;
      if (ccv[mid] < want) lo = mid + 1; else hi = mid;
    }
    seg[t] = lo;
  }
}

int32_t build_rows_layout(const mrs_ratings* R) {
  auto& L = R->sl;
  if (L.rows_built) return MRS_OK;
  cudaStream_t st = R->eng->stream;
  const int sms = R->eng->sm_count;
  const int32_t subs = (L.n_known + kRowsSub - 1) / kRowsSub;
  L.n_sub = std::max(1, (subs + kRowsWarps - 1) / kRowsWarps) * kRowsWarps;
  const int64_t seg_n = (int64_t)R->n_items * (L.n_sub + 1);
  MRS_REQUIRE(seg_n < ((int64_t)1 << 31), MRS_ERR_UNSUPPORTED, "similarity rows: segment table of %lld entries is too large", (long long)seg_n);
  MRS_TRY(dev_alloc(&L.clen, (size_t)L.n_known));
  MRS_TRY(dev_alloc(&L.ccv, (size_t)R->n));
  MRS_TRY(dev_alloc(&L.seg, (size_t)seg_n));
  MRS_CUDA(cudaMemcpyAsync(L.clen, L.h_len.data(), sizeof(int32_t) * L.n_known, cudaMemcpyHostToDevice, st));
  rows_ccv_kernel<<<sms * 8, 256, 0, st>>>(R->irow, L.cidx, R->n, L.ccv);
  rows_seg_kernel<<<sms * 8, 256, 0, st>>>(R->icolp, L.ccv, R->n_items, L.n_sub, L.seg);
  count_launch(2);
  MRS_CUDA(cudaGetLastError());
  MRS_CUDA(cudaStreamSynchronize(st));
  L.rows_built = true;
  return MRS_OK;
}

__global__ void rows_cpre_kernel(const double* __restrict__ upre, const int32_t* __restrict__ csc_src, int64_t n, double* __restrict__ cpre) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < n; p += stride) cpre[p] = upre[csc_src[p]];
}

// ---------------- R1 + R2: similarity row of one user, range by range, reduced on the fly to its first k neighbours -----
__device__ void bitonic_sort_shared(double* key, int32_t* id, int32_t P) {
  for (int32_t size = 2; size <= P; size <<= 1) {
    for (int32_t stride = size >> 1; stride > 0; stride >>= 1) {
      __syncthreads();
      for (int32_t t = threadIdx.x; t < (P >> 1); t += blockDim.x) {
        const int32_t lo = 2 * t - (t & (stride - 1));
        const int32_t hi = lo + stride;
        const bool up = ((lo & size) == 0);
        const double ka = key[lo], kb = key[hi];
        const int32_t ia = id[lo], ib = id[hi];
        const bool swap = up ? before(kb, ib, ka, ia) : before(ka, ia, kb, ib);
        if (swap) { key[lo] = kb; key[hi] = ka; id[lo] = ib; id[hi] = ia; }
      }
    }
  }
  __syncthreads();
}

// One CTA per row user u.  For each range of 16,384 compact user indices: (R1) every warp accumulates its 512 similarities
// in shared memory, (R2) the CTA passes the 16,384 values through the running top-k filter.  A row never leaves the SM:
// only the k neighbours are written.
template <int MODE>  // 1: cosine (P:424-426), 2: jaccard (P:454-458)
__global__ void __launch_bounds__(kRowsWarps * 32, 1)
    knn_rows_kernel(const int32_t* __restrict__ urow, const int32_t* __restrict__ ucol, const double* __restrict__ upre,
                    const int32_t* __restrict__ known_user, const int32_t* __restrict__ clen, const int32_t* __restrict__ rows,
                    const int32_t* __restrict__ seg, int32_t seg_stride, const int32_t* __restrict__ ccv,
                    const double* __restrict__ cpre, int32_t n_known, int32_t n_ranges, int32_t kk, int32_t row_lo,
                    int32_t* __restrict__ nbr_id, double* __restrict__ nbr_sim, const int32_t* __restrict__ tie_rank,
                    const int32_t* __restrict__ tie_inv) {
  extern __shared__ __align__(16) double acc_all[];             // [kRowsWarps * kRowsSub] accumulators
  double* c_key = acc_all + kRowsWarps * kRowsSub;              // [kSelCap] candidate similarities
  int32_t* c_id = (int32_t*)(c_key + kSelCap);                  // [kSelCap] candidate compact indices
  __shared__ int32_t count;
  __shared__ double tau_key;
  __shared__ int32_t tau_id;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  double* acc = acc_all + w * kRowsSub;
  const int32_t cu = rows[blockIdx.x];
  const int32_t u = known_user[cu];
  const int32_t b = urow[u], e = urow[u + 1];
  const int32_t nu = e - b;
  if (threadIdx.x == 0) { count = 0; tau_key = -INFINITY; tau_id = INT_MAX; }

  for (int32_t range = 0; range < n_ranges; ++range) {
    const int32_t sub = range * kRowsWarps + w;
    const int32_t base_cv = sub * kRowsSub;
#pragma unroll
    for (int x = 0; x < kRowsSub / 32; ++x) acc[x * 32 + lane] = 0.0;
    __syncwarp();
    if (base_cv < n_known) {
      const int32_t* segp = seg + sub;
      for (int32_t j0 = b; j0 < e; j0 += 32) {
        const int32_t p = j0 + lane;
        const bool valid = p < e;
        const int32_t item = valid ? __ldg(ucol + p) : 0;
        const double ru = (MODE == 1) ? (valid ? __ldg(upre + p) : 0.0) : 1.0;
        const int64_t so = (int64_t)item * seg_stride;
        const int32_t lo = valid ? __ldg(segp + so) : 0;
        const int32_t hi = valid ? __ldg(segp + so + 1) : 0;
        const int cnt = min(32, e - j0);
        // stage d holds the first 32 entries of the slice of item (t0 + d), requested kRowsDepth items ahead of their use
        int32_t s_lo[kRowsDepth], s_hi[kRowsDepth], s_cv[kRowsDepth];
        double s_x[kRowsDepth];
#pragma unroll
        for (int d = 0; d < kRowsDepth; ++d) {
          s_lo[d] = __shfl_sync(0xffffffffu, lo, d);
          s_hi[d] = __shfl_sync(0xffffffffu, hi, d);
          s_cv[d] = 0;
          s_x[d] = 0.0;
          const int32_t q = s_lo[d] + lane;
          if (d < cnt && q < s_hi[d]) {
            s_cv[d] = __ldg(ccv + q);
            if (MODE == 1) s_x[d] = __ldg(cpre + q);
          }
        }
        for (int t0 = 0; t0 < cnt; t0 += kRowsDepth) {
#pragma unroll
          for (int d = 0; d < kRowsDepth; ++d) {
            const int t = t0 + d;
            if (t < cnt) {  // warp-uniform
              const double ru_t = __shfl_sync(0xffffffffu, ru, t);
              const int32_t clo = s_lo[d], chi = s_hi[d];
              const int32_t cv0 = s_cv[d];
              const double x0 = s_x[d];
              const int tn = t + kRowsDepth;
              if (tn < cnt) {  // refill the stage
                s_lo[d] = __shfl_sync(0xffffffffu, lo, tn);
                s_hi[d] = __shfl_sync(0xffffffffu, hi, tn);
                const int32_t q = s_lo[d] + lane;
                if (q < s_hi[d]) {
                  s_cv[d] = __ldg(ccv + q);
                  if (MODE == 1) s_x[d] = __ldg(cpre + q);
                }
              }
              if (clo + lane < chi) {
                double* a = acc + (cv0 - base_cv);
                *a = __dadd_rn(*a, (MODE == 1) ? __dmul_rn(ru_t, x0) : 1.0);  // ascending item id, no FMA
              }
              for (int32_t q = clo + 32 + lane; q < chi; q += 32) {  // slices longer than a warp (popular items)
                double* a = acc + (__ldg(ccv + q) - base_cv);
                *a = __dadd_rn(*a, (MODE == 1) ? __dmul_rn(ru_t, __ldg(cpre + q)) : 1.0);
              }
              __syncwarp();  // the next item may update the same user from another lane
            }
          }
        }
      }
    }
    __syncthreads();  // the range's 16,384 similarities are complete (and count / tau initialised before the first use)

    // ---- R2: candidates that beat the running k-th best (sim, id) go to the buffer; a full buffer is sorted and cut to k
    const int32_t range_base = range * kRowsWarps * kRowsSub;
    for (int32_t tile0 = 0; tile0 < kRowsWarps * kRowsSub && range_base + tile0 < n_known; tile0 += kSelTile) {
      const double tk = tau_key;
      const int32_t ti = tau_id;
#pragma unroll
      for (int r = 0; r < kSelTile / (kRowsWarps * 32); ++r) {
        const int32_t xl = tile0 + r * (kRowsWarps * 32) + threadIdx.x;
        const int32_t x = range_base + xl;
        double s = 0.0;
        bool ok = false;
        int32_t x_key = x;
        if (x < n_known && x != cu) {  // P:608 allUsers - u
          s = acc_all[xl];
          if (MODE == 2) s = s / (double)(nu + clen[x] - (int32_t)s);  // P:458
          if (s != s) s = -INFINITY;  // NaN has no place in the (sim desc, id asc) order: it ranks last (see knn.cu sort_key)
          if (tie_rank) x_key = tie_rank[x];  // ties are ordered by the user's place in the tie order
          ok = before(s, x_key, tk, ti);
        }
        const unsigned m = __ballot_sync(0xffffffffu, ok);
        if (m) {
          int32_t pos = 0;
          if (lane == 0) pos = atomicAdd(&count, __popc(m));
          pos = __shfl_sync(0xffffffffu, pos, 0) + __popc(m & ((1u << lane) - 1u));
          if (ok) { c_key[pos] = s; c_id[pos] = x_key; }
        }
      }
      __syncthreads();
      const int32_t c = count;
      if (c > kSelCap - kSelTile) {  // compact: keep the first kk, raise the bar to the kk-th
        int32_t P = 2;
        while (P < c) P <<= 1;
        for (int32_t x = c + threadIdx.x; x < P; x += blockDim.x) { c_key[x] = -INFINITY; c_id[x] = INT_MAX; }
        bitonic_sort_shared(c_key, c_id, P);
        if (threadIdx.x == 0) {
          count = min(c, kk);
          if (c >= kk) { tau_key = c_key[kk - 1]; tau_id = c_id[kk - 1]; }
        }
      }
      __syncthreads();
    }
  }
  const int32_t c = count;
  int32_t P = 2;
  while (P < c) P <<= 1;
  for (int32_t x = c + threadIdx.x; x < P; x += blockDim.x) { c_key[x] = -INFINITY; c_id[x] = INT_MAX; }
  bitonic_sort_shared(c_key, c_id, P);
  const int64_t off = (int64_t)(cu - row_lo) * kk;
  for (int32_t j = threadIdx.x; j < kk; j += blockDim.x) {
    nbr_id[off + j] = known_user[tie_inv ? tie_inv[c_id[j]] : c_id[j]];
    nbr_sim[off + j] = c_key[j];
  }
}

// ---------------- R3: weighted-sum deviation + prediction from the lists, one warp per (u, i) -------------------------
template <bool WSD>
__device__ __forceinline__ double predict_pair_lists(int32_t u, int32_t i, int lane, int32_t n_users, int32_t n_items,
                                                     const double* __restrict__ uavg, double gavg, const int32_t* __restrict__ urow,
                                                     const int32_t* __restrict__ ucol, const double* __restrict__ udev,
                                                     const int32_t* __restrict__ icolp, const int32_t* __restrict__ cidx,
                                                     const int32_t* __restrict__ nbr_id, const double* __restrict__ nbr_sim,
                                                     int32_t row_lo, int32_t row_hi, int32_t k_fit, int32_t kk) {
  const double ua = (u >= 0 && u < n_users) ? uavg[u] : -1.0;
  if (ua < 0.0) return WSD ? 0.0 : gavg;  // P:572-573; wsd of a user without ratings: every similarity is 0 (P:527-529)
  const int32_t cu = cidx[u];
  if (cu < row_lo || cu >= row_hi) return nan("");  // rows of another rank: fail loudly
  double num = 0.0, den = 0.0;
  if (i >= 0 && i < n_items && icolp[i + 1] > icolp[i]) {
    const int64_t off = (int64_t)(cu - row_lo) * k_fit;
    for (int32_t j = lane; j < kk; j += 32) {
      const int32_t v = __ldg(nbr_id + off + j);
      int32_t lo = urow[v], hi = urow[v + 1];
      while (lo < hi) {  // items ascending inside a row
        const int32_t mid = (lo + hi) >> 1;
        if (__ldg(ucol + mid) < i) lo = mid + 1; else hi = mid;
      }
      if (lo < urow[v + 1] && __ldg(ucol + lo) == i) {
        const double s = __ldg(nbr_sim + off + j);
        num += udev[lo] * s;  // P:522
        den += fabs(s);
      }
    }
  }
  num = warp_sum(num);
  den = warp_sum(den);
  const double w = den > 0.0 ? num / den : 0.0;  // P:527-529
  if (WSD) return w;
  return combine_fn(ua, w);                       // P:578
}

template <typename VT>
__global__ void __launch_bounds__(256) lists_mae_kernel(const int32_t* __restrict__ tu, const int32_t* __restrict__ ti,
                                                       const VT* __restrict__ tv, int64_t n, int32_t n_users, int32_t n_items,
                                                       const double* __restrict__ uavg, const double* __restrict__ gavg_p,
                                                       const int32_t* __restrict__ urow, const int32_t* __restrict__ ucol,
                                                       const double* __restrict__ udev, const int32_t* __restrict__ icolp,
                                                       const int32_t* __restrict__ cidx, const int32_t* __restrict__ nbr_id,
                                                       const double* __restrict__ nbr_sim, int32_t row_lo, int32_t row_hi, int32_t k_fit,
                                                       int32_t kk, double* __restrict__ part, unsigned int* __restrict__ counter,
                                                       double* __restrict__ out2) {
  __shared__ double sh[8];
  __shared__ bool is_last;
  const double gavg = gavg_p[0];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  double acc = 0.0;  // identical in every lane of the warp
  for (int64_t q = blockIdx.x * (int64_t)wpb + wid; q < n; q += (int64_t)gridDim.x * wpb) {
    const double pr = predict_pair_lists<false>(tu[q], ti[q], lane, n_users, n_items, uavg, gavg, urow, ucol, udev, icolp, cidx, nbr_id,
                                                nbr_sim, row_lo, row_hi, k_fit, kk);
    acc += fabs(decode_value(tv[q]) - pr);  // P:71
  }
  if (lane == 0) sh[wid] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < wpb; ++w) t += sh[w];
    part[blockIdx.x] = t;
    __threadfence();
    is_last = (atomicAdd(counter, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (is_last) {
    __threadfence();
    double t = 0.0;
    for (int b = threadIdx.x; b < (int)gridDim.x; b += blockDim.x) t += __ldcg(&part[b]);
    t = warp_sum(t);
    if (lane == 0) sh[wid] = t;
    __syncthreads();
    if (threadIdx.x == 0) {
      double s = 0.0;
      for (int w = 0; w < wpb; ++w) s += sh[w];
      out2[0] = s;
      out2[1] = (double)n;
      *counter = 0;
    }
  }
}

template <bool WSD>
__global__ void __launch_bounds__(256) lists_pairs_kernel(const int32_t* __restrict__ us, const int32_t* __restrict__ is, int64_t n,
                                                         int32_t n_users, int32_t n_items, const double* __restrict__ uavg,
                                                         const double* __restrict__ gavg_p, const int32_t* __restrict__ urow,
                                                         const int32_t* __restrict__ ucol, const double* __restrict__ udev,
                                                         const int32_t* __restrict__ icolp, const int32_t* __restrict__ cidx,
                                                         const int32_t* __restrict__ nbr_id, const double* __restrict__ nbr_sim,
                                                         int32_t row_lo, int32_t row_hi, int32_t k_fit, int32_t kk, double* __restrict__ out) {
  const double gavg = gavg_p[0];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  for (int64_t q = blockIdx.x * (int64_t)wpb + wid; q < n; q += (int64_t)gridDim.x * wpb) {
    const double pr = predict_pair_lists<WSD>(us[q], is[q], lane, n_users, n_items, uavg, gavg, urow, ucol, udev, icolp, cidx, nbr_id,
                                              nbr_sim, row_lo, row_hi, k_fit, kk);
    if (lane == 0) out[q] = pr;
  }
}

}  // namespace

void free_rows_layout(const mrs_ratings* r) {
  auto& L = r->sl;
  dev_free(L.clen); dev_free(L.ccv); dev_free(L.seg);
  L.clen = nullptr; L.ccv = nullptr; L.seg = nullptr;
  L.rows_built = false;
}

void rows_free(mrs_sim* s) {
  dev_free(s->row_order); dev_free(s->cpre);  // the lists themselves are freed with the handle
  s->row_order = nullptr; s->cpre = nullptr;
}

// allocate the buffers of a row range: users with original id in [user_lo, user_hi) own lists
int32_t rows_alloc(mrs_model* m, mrs_sim* s, int32_t user_lo, int32_t user_hi) {
  const mrs_ratings* R = m->train;
  MRS_TRY(build_rows_layout(R));
  const auto& L = R->sl;
  MRS_REQUIRE(user_lo <= user_hi, MRS_ERR_INVALID, "mrs_fit_similarity_rows: empty or reversed user range [%d, %d)", user_lo, user_hi);
  MRS_REQUIRE(s->k <= kMaxListK, MRS_ERR_UNSUPPORTED, "mrs_fit_similarity: the row-block path keeps at most %d neighbours per user (k = %d)", kMaxListK, s->k);
  s->row_lo = (int32_t)(std::lower_bound(L.h_known.begin(), L.h_known.end(), user_lo) - L.h_known.begin());
  s->row_hi = (int32_t)(std::lower_bound(L.h_known.begin(), L.h_known.end(), user_hi) - L.h_known.begin());
  s->k_fit = std::max(0, std::min(s->k, L.n_known - 1));
  const int32_t n_rows = s->row_hi - s->row_lo;
  std::vector<int32_t> order;
  order.reserve((size_t)n_rows);
  for (int32_t c : L.h_order)
    if (c >= s->row_lo && c < s->row_hi) order.push_back(c);  // longest rows first: CTAs of a batch cost about the same
  MRS_TRY(dev_alloc(&s->row_order, (size_t)n_rows));
  MRS_TRY(dev_alloc(&s->cpre, (size_t)R->n));
  MRS_TRY(dev_alloc(&s->nbr_id, (size_t)n_rows * (size_t)std::max(s->k_fit, 1)));
  MRS_TRY(dev_alloc(&s->nbr_sim, (size_t)n_rows * (size_t)std::max(s->k_fit, 1)));
  if (n_rows > 0) {
    MRS_CUDA(cudaMemcpyAsync(s->row_order, order.data(), sizeof(int32_t) * n_rows, cudaMemcpyHostToDevice, m->eng->stream));
    MRS_CUDA(cudaStreamSynchronize(m->eng->stream));  // `order` goes out of scope
  }
  return MRS_OK;
}

int32_t rows_fit_async(mrs_model* m, mrs_sim* s, bool /*first*/) {
  const mrs_ratings* R = m->train;
  mrs_engine* e = m->eng;
  const auto& L = R->sl;
  cudaStream_t st = e->stream;
  const int32_t n_rows = s->row_hi - s->row_lo;
  if (n_rows == 0 || s->k_fit == 0) return MRS_OK;
  if (s->kind == MRS_SIM_COSINE) {
    rows_cpre_kernel<<<e->sm_count * 8, 256, 0, st>>>(s->upre, R->csc_src, R->n, s->cpre);
    mark(e, "rows_cpre");
  }
  const size_t smem = (size_t)kRowsWarps * kRowsSub * sizeof(double) + (size_t)kSelCap * (sizeof(double) + sizeof(int32_t));
  MRS_CUDA(cudaFuncSetAttribute(knn_rows_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  MRS_CUDA(cudaFuncSetAttribute(knn_rows_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int32_t n_ranges = L.n_sub / kRowsWarps;
  // rows are listed longest first: the hardware hands the next CTA to the first SM that frees up (longest-processing-time order)
  if (s->kind == MRS_SIM_COSINE)
    knn_rows_kernel<1><<<n_rows, kRowsWarps * 32, smem, st>>>(R->urow, R->ucol, s->upre, L.known_user, L.clen, s->row_order, L.seg, L.n_sub + 1,
                                                            L.ccv, s->cpre, L.n_known, n_ranges, s->k_fit, s->row_lo, s->nbr_id, s->nbr_sim, m->tie_rank, m->tie_inv);
  else
    knn_rows_kernel<2><<<n_rows, kRowsWarps * 32, smem, st>>>(R->urow, R->ucol, s->upre, L.known_user, L.clen, s->row_order, L.seg, L.n_sub + 1,
                                                            L.ccv, s->cpre, L.n_known, n_ranges, s->k_fit, s->row_lo, s->nbr_id, s->nbr_sim, m->tie_rank, m->tie_inv);
  mark(e, "knn_rows");
  MRS_CUDA(cudaGetLastError());
  return MRS_OK;
}

int32_t mae_lists_async(const mrs_model* m, const mrs_sim* s, const mrs_ratings* T, double* d_out2) {
  const mrs_ratings* R = m->train;
  const auto& L = R->sl;
  cudaStream_t st = m->eng->stream;
  const int wpb = 8;
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((T->n + wpb - 1) / wpb, (int64_t)s->mae_part_cap));
  const int32_t kk = std::min(s->k, s->k_fit);
  if (T->value_kind == kValueCode)
    lists_mae_kernel<uint8_t><<<grid, 256, 0, st>>>(T->coo_u, T->ucol, (const uint8_t*)T->uval, T->n, m->n_users, m->n_items, m->uavg, m->gavg,
                                                    R->urow, R->ucol, s->udev, R->icolp, L.cidx, s->nbr_id, s->nbr_sim, s->row_lo,
                                                    s->row_hi, s->k_fit, kk, s->mae_part, s->counter, d_out2);
  else
    lists_mae_kernel<double><<<grid, 256, 0, st>>>(T->coo_u, T->ucol, (const double*)T->uval, T->n, m->n_users, m->n_items, m->uavg, m->gavg,
                                                   R->urow, R->ucol, s->udev, R->icolp, L.cidx, s->nbr_id, s->nbr_sim, s->row_lo,
                                                   s->row_hi, s->k_fit, kk, s->mae_part, s->counter, d_out2);
  mark(m->eng, "lists_mae");
  MRS_CUDA(cudaGetLastError());
  return MRS_OK;
}

int32_t predict_lists_async(const mrs_model* m, const mrs_sim* s, const int32_t* d_users, const int32_t* d_items, int64_t n, double* d_out,
                            bool wsd_only) {
  const mrs_ratings* R = m->train;
  const auto& L = R->sl;
  cudaStream_t st = m->eng->stream;
  const int wpb = 8;
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((n + wpb - 1) / wpb, (int64_t)m->eng->sm_count * 16));
  const int32_t kk = std::min(s->k, s->k_fit);
  if (wsd_only)
    lists_pairs_kernel<true><<<grid, 256, 0, st>>>(d_users, d_items, n, m->n_users, m->n_items, m->uavg, m->gavg, R->urow, R->ucol, s->udev,
                                                   R->icolp, L.cidx, s->nbr_id, s->nbr_sim, s->row_lo, s->row_hi, s->k_fit, kk, d_out);
  else
    lists_pairs_kernel<false><<<grid, 256, 0, st>>>(d_users, d_items, n, m->n_users, m->n_items, m->uavg, m->gavg, R->urow, R->ucol, s->udev,
                                                    R->icolp, L.cidx, s->nbr_id, s->nbr_sim, s->row_lo, s->row_hi, s->k_fit, kk, d_out);
  mark(m->eng, "lists_pairs");
  MRS_CUDA(cudaGetLastError());
  return MRS_OK;
}

}  // namespace mrs
