// baseline.cu -- the baseline family of shared/predictions.scala as segmented-reduction kernels:
//   global / per-user / per-item averages (P:94-148, Spark twins P:265-309), scale()-normalised deviations
//   (P:155-169, P:316-329), per-item average deviation (P:176-198, P:336-355), the baseline predictor
//   (P:205-237, P:362-391) and the fused prediction + |error| reduction of MAE (P:69-86, P:256-258).
//
// Every kernel here is HBM/L2-bound integer/byte + fp64 work: no tensor cores.  Accumulation is fp64; rating
// sums of half-star data are exact in any order, deviation sums are reduced in a fixed order
// (lane-strided inside a chunk, xor-shuffle tree, chunks of a column in ascending order), so results are
// run-to-run bit-reproducible.
#include <algorithm>
#include <cmath>
#include <cstdlib>

#include "common.cuh"

namespace mrs {
namespace {

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__global__ void fill_f64_kernel(double* __restrict__ p, int64_t n, double v) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = v;
}

// ---- K1: per-user chunk sums of ratings (user-major values only: 1 B/rating for half-star codes)
template <typename VT>
__global__ void __launch_bounds__(256) user_chunk_sum_kernel(const VT* __restrict__ uval, const int32_t* __restrict__ urow,
                                                            const int32_t* __restrict__ chunk_seg,
                                                            const int32_t* __restrict__ chunk_begin, int32_t n_chunks,
                                                            double* __restrict__ upart) {
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  for (int32_t c = blockIdx.x * warps_per_block + (threadIdx.x >> 5); c < n_chunks; c += gridDim.x * warps_per_block) {
    const int32_t seg = chunk_seg[c];
    const int32_t b = chunk_begin[c];
    const int32_t e = min(b + kUserChunk, urow[seg + 1]);
    double s;
    if (sizeof(VT) == 1) {
      uint32_t acc = 0;  // <= 2048 * 255: exact
      for (int32_t p = b + lane; p < e; p += 32) acc += (uint32_t)uval[p];
      acc = __reduce_add_sync(0xffffffffu, acc);
      s = 0.5 * (double)acc;
    } else {
      double acc = 0.0;
      for (int32_t p = b + lane; p < e; p += 32) acc += (double)uval[p];
      s = warp_sum(acc);
    }
    if (lane == 0) upart[c] = s;
  }
}

// ---- K1 (half-star codes): streaming of the padded user-major code array, one 128-bit load (16 codes of ONE user) per
// thread and vector; lanes of the same user are contiguous, so a segmented shuffle reduction leaves one integer
// atomicAdd per (warp, user).  Integer sums are exact: the result does not depend on the order of the atomics.
// Persistent grid (8 blocks of 256 threads per SM, V = 2 vectors per thread in flight = 80 KB per SM): 5.8 us for 26.5 MB at
// ml-25m shape; the one-shot grid of 2,590 blocks this replaced took 11-13 us, 4 blocks per SM with V = 4 take 7.5 us
// (tools/timeline.py, same box).
#ifndef MRS_K1_V
#define MRS_K1_V 2
#endif
__global__ void __launch_bounds__(256) user_sum_kernel(const uint8_t* __restrict__ uval16, const int32_t* __restrict__ vec_row, int32_t n_vec,
                                                      uint32_t* __restrict__ usum, unsigned long long* __restrict__ gsum_codes,
                                                      unsigned long long* __restrict__ tl, int count_off) {
  __shared__ uint32_t sh[8];
  tl_begin(tl, 0);
  pdl_trigger();  // the item pass may start its prologue (ring prefetch) while this kernel runs

  const int lane = threadIdx.x & 31;
  constexpr int V = MRS_K1_V;  // vectors per thread and iteration (all loads are issued before the first use)
  const int32_t T = gridDim.x * blockDim.x;
  uint32_t total = 0;
  for (int32_t t0 = blockIdx.x * blockDim.x + threadIdx.x; t0 - lane < n_vec; t0 += V * T) {  // warp-uniform trip count
    uint4 x[V];
    int32_t row[V];
#pragma unroll
    for (int k = 0; k < V; ++k) {
      const int64_t t = (int64_t)t0 + (int64_t)k * T;
      const bool in = t < n_vec;
      x[k] = in ? __ldg(reinterpret_cast<const uint4*>(uval16) + t) : make_uint4(0u, 0u, 0u, 0u);
      row[k] = in ? __ldg(vec_row + t) : -1;
    }
#pragma unroll
    for (int k = 0; k < V; ++k) {
      uint32_t s = __dp4a(x[k].x, 0x01010101u, __dp4a(x[k].y, 0x01010101u, __dp4a(x[k].z, 0x01010101u, __dp4a(x[k].w, 0x01010101u, 0u))));
      total += s;
      // segmented suffix sums over runs of equal row ids (lanes of one user are contiguous)
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t v = __shfl_down_sync(0xffffffffu, s, o);
        const int32_t r = __shfl_down_sync(0xffffffffu, row[k], o);
        if (lane + o < 32 && r == row[k]) s += v;
      }
      const int32_t prev = __shfl_up_sync(0xffffffffu, row[k], 1);
      if (row[k] >= 0 && (lane == 0 || prev != row[k])) atomicAdd(usum + row[k], s);
    }
  }
  total = __reduce_add_sync(0xffffffffu, total);
  if (lane == 0) sh[threadIdx.x >> 5] = total;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long a = 0;
    for (int k = 0; k < (int)(blockDim.x >> 5); ++k) a += sh[k];
    atomicAdd(gsum_codes, a);  // integer: exact, order independent
  }
  if (count_off) flag_count_off(gsum_codes + 1);  // opt-in flag protocol: the item pass waits for the count, not for the grid
  tl_end(tl, 0);
}

// Users-only refit (mrs_fit_users_async): user averages and the global average from K1's integer sums, in the item pass'
// own words (one correctly rounded division per user, -1.0 = no ratings, P:222); re-arms the sums for the next K1.
__global__ void __launch_bounds__(256) user_avg_kernel(uint32_t* __restrict__ usum, const int32_t* __restrict__ urow, int32_t n_users,
                                                      unsigned long long* __restrict__ k1_part, double n_total, double* __restrict__ uavg,
                                                      double* __restrict__ gavg, double* __restrict__ xbuf_tail) {
  pdl_trigger();
  pdl_wait();  // K1's sums are complete
  for (int32_t u = blockIdx.x * blockDim.x + threadIdx.x; u < n_users; u += gridDim.x * blockDim.x) {
    const uint32_t S = usum[u];
    const uint32_t cnt = (uint32_t)(urow[u + 1] - urow[u]);
    usum[u] = 0;
    uavg[u] = cnt ? (0.5 * (double)S) / (double)cnt : -1.0;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    const double gs = 0.5 * (double)k1_part[0];  // integer sum of codes: exact, order independent
    k1_part[0] = 0; k1_part[1] = 0; k1_part[2] = 0;
    xbuf_tail[0] = gs;
    xbuf_tail[1] = n_total;
    gavg[0] = n_total > 0.0 ? gs / n_total : 0.0;  // P:18 mean of an empty Seq is 0.0
  }
}

// ---- K1b: per-user average (-1.0 sentinel for users without ratings) + global rating sum / count
__global__ void __launch_bounds__(256) user_finalize_kernel(const int32_t* __restrict__ urow, const int32_t* __restrict__ seg_chunk_ptr,
                                                           const double* __restrict__ upart, int32_t n_users,
                                                           double* __restrict__ uavg, double* __restrict__ gsum_gcnt) {
  __shared__ double sh[8];
  const int32_t u = blockIdx.x * blockDim.x + threadIdx.x;
  double s = 0.0;
  if (u < n_users) {
    const int32_t cnt = urow[u + 1] - urow[u];
    for (int32_t c = seg_chunk_ptr[u]; c < seg_chunk_ptr[u + 1]; ++c) s += upart[c];
    uavg[u] = cnt ? s / (double)cnt : -1.0;  // P:18 sum/length; unknown user marked like P:222 getOrElse(user,-1.0)
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    double t = (threadIdx.x < (blockDim.x >> 5)) ? sh[threadIdx.x] : 0.0;
    t = warp_sum(t);
    if (threadIdx.x == 0 && t != 0.0) atomicAdd(&gsum_gcnt[0], t);  // exact for half-star data => order-independent
  }
  if (u == 0) gsum_gcnt[1] = (double)urow[n_users];
}

// ---- K2: per-item chunk sums of normalised deviations and of ratings (item-major: user id + value),
//      gathering the rater's average from the per-user table
template <typename VT>
__global__ void __launch_bounds__(256) item_chunk_dev_kernel(const int32_t* __restrict__ irow, const VT* __restrict__ ival,
                                                            const int32_t* __restrict__ icolp,
                                                            const int32_t* __restrict__ chunk_seg,
                                                            const int32_t* __restrict__ chunk_begin, int32_t n_chunks,
                                                            const double* __restrict__ uavg, double* __restrict__ ipart) {
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  for (int32_t c = blockIdx.x * warps_per_block + (threadIdx.x >> 5); c < n_chunks; c += gridDim.x * warps_per_block) {
    const int32_t seg = chunk_seg[c];
    const int32_t b = chunk_begin[c];
    const int32_t e = min(b + kItemChunk, icolp[seg + 1]);
    double dsum = 0.0, rsum = 0.0;
#pragma unroll 4
    for (int32_t p = b + lane; p < e; p += 32) {
      const double r = decode_value(ival[p]);
      const double a = __ldg(&uavg[irow[p]]);
      dsum += deviation_fn(r, a);  // P:167 / P:327
      rsum += r;
    }
    dsum = warp_sum(dsum);
    rsum = warp_sum(rsum);
    if (lane == 0) {
      ipart[c] = dsum;
      ipart[n_chunks + c] = rsum;
    }
  }
}

// ---- K2b: chunk partials -> exchange buffer [devsum | count | gsum gcount | ratesum] (P:176-186 (sum,count); P:267 reduceByKey)
__global__ void __launch_bounds__(256) item_partial_kernel(const int32_t* __restrict__ icolp, const int32_t* __restrict__ seg_chunk_ptr,
                                                          const double* __restrict__ ipart, int32_t n_chunks, int32_t n_items,
                                                          double* __restrict__ xbuf) {
  const int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_items) return;
  double ds = 0.0, rs = 0.0;
  for (int32_t c = seg_chunk_ptr[i]; c < seg_chunk_ptr[i + 1]; ++c) {
    ds += ipart[c];
    rs += ipart[n_chunks + c];
  }
  xbuf[i] = ds;
  xbuf[(size_t)n_items + i] = (double)(icolp[i + 1] - icolp[i]);
  xbuf[2 * (size_t)n_items + 2 + i] = rs;
}

// ---- K2c: (after the optional cross-rank sum of xbuf) per-item averages and the global average
__global__ void __launch_bounds__(256) item_finalize_kernel(const double* __restrict__ xbuf, int32_t n_items,
                                                           double* __restrict__ idevavg, double* __restrict__ iavg,
                                                           double* __restrict__ gavg) {
  pdl_trigger();  // the MAE pass may set up its rings while this runs
  pdl_wait();     // xbuf is complete (local finalisation or the cross-rank exchange)
  const int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) {
    const double gs = xbuf[2 * (size_t)n_items], gc = xbuf[2 * (size_t)n_items + 1];
    gavg[0] = gc > 0.0 ? gs / gc : 0.0;  // P:18 mean of an empty Seq is 0.0
  }
  if (i >= n_items) return;
  const double cnt = xbuf[(size_t)n_items + i];
  idevavg[i] = cnt > 0.0 ? xbuf[i] / cnt : 0.0;                                     // P:185 x._1/x._2 ; unknown item -> 0.0 (P:197)
  iavg[i] = cnt > 0.0 ? xbuf[2 * (size_t)n_items + 2 + i] / cnt : nan("");          // unknown item -> global average at query time (P:147)
}

// ---- prediction of one (u,i) for the five closed-form predictors
template <int KIND>
__device__ __forceinline__ double predict_one(int32_t u, int32_t i, int32_t n_users, int32_t n_items,
                                              const double* __restrict__ uavg, const double* __restrict__ idevavg,
                                              const double* __restrict__ iavg, double gavg) {
  if (KIND == MRS_PRED_GLOBAL) return gavg;  // P:105
  if (KIND == MRS_PRED_USER) {               // P:126
    const double a = (u >= 0 && u < n_users) ? __ldg(&uavg[u]) : -1.0;
    return a < 0.0 ? gavg : a;
  }
  if (KIND == MRS_PRED_ITEM) {  // P:147
    const double a = (i >= 0 && i < n_items) ? __ldg(&iavg[i]) : nan("");
    return isnan(a) ? gavg : a;
  }
  if (KIND == MRS_PRED_ITEMDEV) return (i >= 0 && i < n_items) ? __ldg(&idevavg[i]) : 0.0;  // P:197
  // baseline, P:217-236
  const double a = (u >= 0 && u < n_users) ? __ldg(&uavg[u]) : -1.0;
  if (a < 0.0) return gavg;  // P:222-224
  const double d = (i >= 0 && i < n_items) ? __ldg(&idevavg[i]) : 0.0;
  return combine_fn(a, d);  // P:229
}

// ---- K3: fused predict + |r - p| + reduction over the sorted COO of the test set (P:69-86).
// Per-block partials are combined by the last block in block order: deterministic.
template <typename VT, int KIND>
__global__ void __launch_bounds__(256) predict_mae_kernel(const int32_t* __restrict__ tu, const int32_t* __restrict__ ti,
                                                         const VT* __restrict__ tv, int64_t n, int32_t n_users, int32_t n_items,
                                                         const double* __restrict__ uavg, const double* __restrict__ idevavg,
                                                         const double* __restrict__ iavg, const double* __restrict__ gavg_p,
                                                         double* __restrict__ part, unsigned int* __restrict__ counter,
                                                         double* __restrict__ out2) {
  __shared__ double sh[8];
  __shared__ bool is_last;
  const double gavg = gavg_p[0];
  double acc = 0.0;
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x) {
    const double r = decode_value(tv[p]);
    const double pr = predict_one<KIND>(tu[p], ti[p], n_users, n_items, uavg, idevavg, iavg, gavg);
    acc += fabs(r - pr);
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    double t = (threadIdx.x < (blockDim.x >> 5)) ? sh[threadIdx.x] : 0.0;
    t = warp_sum(t);
    if (threadIdx.x == 0) {
      part[blockIdx.x] = t;
      __threadfence();
      const unsigned int done = atomicAdd(counter, 1u);
      is_last = (done == gridDim.x - 1);
    }
  }
  __syncthreads();
  if (is_last) {
    __threadfence();
    double t = 0.0;
    for (int b = threadIdx.x; b < (int)gridDim.x; b += blockDim.x) t += __ldcg(&part[b]);
    t = warp_sum(t);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = t;
    __syncthreads();
    if (threadIdx.x == 0) {
      double s = 0.0;
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += sh[w];
      out2[0] = s;
      out2[1] = (double)n;
      *counter = 0;  // re-arm for the next launch
    }
  }
}

// ---- K3 (half-star codes): same as predict_mae_kernel with 4 consecutive test entries per thread (128-bit loads of the
// user and item ids, 32-bit load of the codes) so that enough bytes are in flight; the tail (n % 4) is done by block 0.
template <int KIND>
__global__ void __launch_bounds__(256) predict_mae4_kernel(const int32_t* __restrict__ tu, const int32_t* __restrict__ ti,
                                                          const uint8_t* __restrict__ tv, int64_t n, int32_t n_users, int32_t n_items,
                                                          const double* __restrict__ uavg, const double* __restrict__ idevavg,
                                                          const double* __restrict__ iavg, const double* __restrict__ gavg_p,
                                                          double* __restrict__ part, unsigned int* __restrict__ counter,
                                                          double* __restrict__ out2) {
  __shared__ double sh[8];
  __shared__ bool is_last;
  const double gavg = gavg_p[0];
  const int64_t n4 = n >> 2;
  double acc = 0.0;
  constexpr int Q = 2;  // quads per thread per iteration: 2 x (16 + 16 + 4) bytes requested before the first use
  const int64_t T = (int64_t)gridDim.x * blockDim.x;
  for (int64_t q0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q0 < n4; q0 += Q * T) {
    int4 u4[Q], i4[Q];
    uchar4 c4[Q];
#pragma unroll
    for (int k = 0; k < Q; ++k) {
      const int64_t q = min(q0 + k * T, n4 - 1);
      u4[k] = __ldg(reinterpret_cast<const int4*>(tu) + q);
      i4[k] = __ldg(reinterpret_cast<const int4*>(ti) + q);
      c4[k] = __ldg(reinterpret_cast<const uchar4*>(tv) + q);
    }
    double pr[Q][4];
#pragma unroll
    for (int k = 0; k < Q; ++k) {
      pr[k][0] = predict_one<KIND>(u4[k].x, i4[k].x, n_users, n_items, uavg, idevavg, iavg, gavg);
      pr[k][1] = predict_one<KIND>(u4[k].y, i4[k].y, n_users, n_items, uavg, idevavg, iavg, gavg);
      pr[k][2] = predict_one<KIND>(u4[k].z, i4[k].z, n_users, n_items, uavg, idevavg, iavg, gavg);
      pr[k][3] = predict_one<KIND>(u4[k].w, i4[k].w, n_users, n_items, uavg, idevavg, iavg, gavg);
    }
#pragma unroll
    for (int k = 0; k < Q; ++k) {
      if (q0 + k * T < n4) {
        acc += fabs(0.5 * (double)c4[k].x - pr[k][0]);
        acc += fabs(0.5 * (double)c4[k].y - pr[k][1]);
        acc += fabs(0.5 * (double)c4[k].z - pr[k][2]);
        acc += fabs(0.5 * (double)c4[k].w - pr[k][3]);
      }
    }
  }
  if (blockIdx.x == 0) {
    const int64_t p = (n4 << 2) + threadIdx.x;
    if (p < n) acc += fabs(0.5 * (double)tv[p] - predict_one<KIND>(tu[p], ti[p], n_users, n_items, uavg, idevavg, iavg, gavg));
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    double t = (threadIdx.x < (blockDim.x >> 5)) ? sh[threadIdx.x] : 0.0;
    t = warp_sum(t);
    if (threadIdx.x == 0) {
      part[blockIdx.x] = t;
      __threadfence();
      is_last = (atomicAdd(counter, 1u) == gridDim.x - 1);
    }
  }
  __syncthreads();
  if (is_last) {
    __threadfence();
    double t = 0.0;
    for (int b = threadIdx.x; b < (int)gridDim.x; b += blockDim.x) t += __ldcg(&part[b]);
    t = warp_sum(t);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = t;
    __syncthreads();
    if (threadIdx.x == 0) {
      double s = 0.0;
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += sh[w];
      out2[0] = s;
      out2[1] = (double)n;
      *counter = 0;
    }
  }
}

template <int KIND>
__global__ void __launch_bounds__(256) predict_pairs_kernel(const int32_t* __restrict__ us, const int32_t* __restrict__ is, int64_t n,
                                                           int32_t n_users, int32_t n_items, const double* __restrict__ uavg,
                                                           const double* __restrict__ idevavg, const double* __restrict__ iavg,
                                                           const double* __restrict__ gavg_p, double* __restrict__ out) {
  const double gavg = gavg_p[0];
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x)
    out[p] = predict_one<KIND>(us[p], is[p], n_users, n_items, uavg, idevavg, iavg, gavg);
}

inline int chunk_grid(int32_t n_chunks, int sm_count) {
  const int per_block = 8;  // warps per 256-thread block
  int g = (n_chunks + per_block - 1) / per_block;
  return std::max(1, std::min(g, sm_count * 8));
}

// code path: K1 (user code sums) -> K2 tiled item pass (forms the user averages on the way) -> K2b finalize
int32_t launch_fit_codes(mrs_engine* e, const mrs_ratings* R, mrs_model* m, bool fused, const PushDev* push, bool no_finalize) {
  // (the user sums are zero on entry: cleared when the model is created and re-armed by the kernel that consumes them last --
  // K2b, the push kernel or the fused test pass -- so no memset node opens the pass)
  user_sum_kernel<<<m->k1_blocks, 256, 0, e->stream>>>(R->uval16, R->vec_row, R->n_vec, m->usum, m->k1_part, e->d_timeline, m->flag_sync ? 1 : 0);
  mark(e, "user_sum");
  MRS_CUDA(cudaGetLastError());
  return launch_item_tiled(e, R, m, fused, push, no_finalize);
}

template <typename VT>
int32_t launch_fit_local(mrs_engine* e, const mrs_ratings* R, mrs_model* m) {
  cudaStream_t st = e->stream;
  const int32_t NI = R->n_items, NU = R->n_users;
  MRS_CUDA(cudaMemsetAsync(m->xbuf + 2 * (size_t)NI, 0, 2 * sizeof(double), st));
  if (R->uch.n_chunks > 0) {
    user_chunk_sum_kernel<VT><<<chunk_grid(R->uch.n_chunks, e->sm_count), 256, 0, st>>>(
        (const VT*)R->uval, R->urow, R->uch.chunk_seg, R->uch.chunk_begin, R->uch.n_chunks, m->upart);
    mark(e, "user_chunk_sum");
  }
  user_finalize_kernel<<<(NU + 255) / 256, 256, 0, st>>>(R->urow, R->uch.seg_chunk_ptr, m->upart, NU, m->uavg, m->xbuf + 2 * (size_t)NI);
  mark(e, "user_finalize");
  if (R->ich.n_chunks > 0) {
    item_chunk_dev_kernel<VT><<<chunk_grid(R->ich.n_chunks, e->sm_count), 256, 0, st>>>(
        R->irow, (const VT*)R->ival, R->icolp, R->ich.chunk_seg, R->ich.chunk_begin, R->ich.n_chunks, m->uavg, m->ipart);
    mark(e, "item_chunk_dev");
  }
  item_partial_kernel<<<(NI + 255) / 256, 256, 0, st>>>(R->icolp, R->ich.seg_chunk_ptr, m->ipart, R->ich.n_chunks, NI, m->xbuf);
  mark(e, "item_partial");
  MRS_CUDA(cudaGetLastError());
  return MRS_OK;
}

template <typename VT, int KIND>
int32_t launch_mae(const mrs_model* m, const mrs_ratings* T, double* d_out2) {
  mrs_engine* e = m->eng;
  int grid = (int)std::min<int64_t>((T->n + 255) / 256, (int64_t)m->mae_part_cap);
  if (grid < 1) grid = 1;
  if (sizeof(VT) == 1) {
    grid = (int)std::max<int64_t>(1, std::min<int64_t>(((T->n >> 3) + 255) / 256, (int64_t)m->mae_part_cap));
    predict_mae4_kernel<KIND><<<grid, 256, 0, e->stream>>>(T->coo_u, T->ucol, (const uint8_t*)T->uval, T->n, m->n_users, m->n_items,
                                                           m->uavg, m->idevavg, m->iavg, m->gavg, m->mae_part, m->counters, d_out2);
    mark(e, "predict_mae");
    MRS_CUDA(cudaGetLastError());
    return MRS_OK;
  }
  predict_mae_kernel<VT, KIND><<<grid, 256, 0, e->stream>>>(T->coo_u, T->ucol, (const VT*)T->uval, T->n, m->n_users, m->n_items,
                                                            m->uavg, m->idevavg, m->iavg, m->gavg, m->mae_part, m->counters, d_out2);
  mark(e, "predict_mae");
  MRS_CUDA(cudaGetLastError());
  return MRS_OK;
}

template <typename VT>
int32_t dispatch_mae(const mrs_model* m, int32_t kind, const mrs_ratings* T, double* d_out2) {
  switch (kind) {
    case MRS_PRED_GLOBAL: return launch_mae<VT, MRS_PRED_GLOBAL>(m, T, d_out2);
    case MRS_PRED_USER: return launch_mae<VT, MRS_PRED_USER>(m, T, d_out2);
    case MRS_PRED_ITEM: return launch_mae<VT, MRS_PRED_ITEM>(m, T, d_out2);
    case MRS_PRED_ITEMDEV: return launch_mae<VT, MRS_PRED_ITEMDEV>(m, T, d_out2);
    case MRS_PRED_BASELINE: return launch_mae<VT, MRS_PRED_BASELINE>(m, T, d_out2);
    default: set_error("mrs_mae: predictor kind %d needs a similarity handle", kind); return MRS_ERR_INVALID;
  }
}

}  // namespace

int32_t fit_local(mrs_engine* e, const mrs_ratings* R, mrs_model** inout, bool fused_finalize, const PushDev* push, bool no_finalize) {
  MRS_REQUIRE(e && R && inout, MRS_ERR_INVALID, "mrs_fit: NULL argument");
  use_engine(e);
  const bool codes = (R->value_kind == kValueCode);
  if (codes) MRS_TRY(build_tiled_layout(R));
  mrs_model* m = *inout;
  if (m && m->train != R) {
    set_error("mrs_fit_local: the model passed for reuse was fitted on a different rating set");
    return MRS_ERR_INVALID;
  }
  if (!m) {
    m = new mrs_model();
    m->eng = e;
    m->train = R;
    m->n_users = R->n_users;
    m->n_items = R->n_items;
    m->mae_part_cap = e->sm_count * 16;
    {  // persistent K1: MRS_K1_BPS blocks per SM (default 8), never more blocks than vectors need
      const int bps = getenv("MRS_K1_BPS") ? std::max(1, atoi(getenv("MRS_K1_BPS"))) : 8;
      m->k1_blocks = std::max(1, std::min(e->sm_count * bps, (R->n_vec + 255) / 256));
    }
    int32_t s = MRS_OK;
    if (codes) {
      if (s == MRS_OK) s = dev_alloc(&m->usum, (size_t)R->n_users);
      if (s == MRS_OK && cudaMemsetAsync(m->usum, 0, sizeof(uint32_t) * (size_t)R->n_users, e->stream) != cudaSuccess) s = MRS_ERR_CUDA;
      if (s == MRS_OK) s = dev_alloc(&m->k1_part, 4);
      if (s == MRS_OK && cudaMemsetAsync(m->k1_part, 0, 4 * sizeof(unsigned long long), e->stream) != cudaSuccess) s = MRS_ERR_CUDA;
      m->flag_sync = getenv("MRS_FLAGSYNC") && atoi(getenv("MRS_FLAGSYNC")) != 0;  // opt-in: measured slower (profiles/r02_summary.md)
      if (s == MRS_OK) s = dev_alloc(&m->xdev_fix, 2 * (size_t)R->n_items);  // two buffers, used alternately by the folded closure (tiled.cu)
      if (s == MRS_OK) s = dev_alloc(&m->xcode_sum, (size_t)R->n_items);
      if (s == MRS_OK && cudaMemsetAsync(m->xdev_fix, 0, 2 * sizeof(long long) * (size_t)R->n_items, e->stream) != cudaSuccess) s = MRS_ERR_CUDA;
      if (s == MRS_OK && cudaMemsetAsync(m->xcode_sum, 0, sizeof(unsigned long long) * (size_t)R->n_items, e->stream) != cudaSuccess) s = MRS_ERR_CUDA;
    }
    if (s == MRS_OK) s = dev_alloc(&m->upart, (size_t)R->uch.n_chunks);
    // padded to whole user tiles: the tiled kernel stages a tile's averages with one 64 KB bulk copy
    const size_t uavg_len = ((size_t)R->n_users + kTileUsers - 1) / kTileUsers * kTileUsers + kTileUsers;
    if (s == MRS_OK) s = dev_alloc(&m->uavg, uavg_len);
    if (s == MRS_OK && codes) {
      // the item pass writes the averages of the tiles that have ratings; user tiles without any (the id range of another
      // rank in a sharded run) keep this "no ratings" mark (P:222 getOrElse(user, -1.0)) for the life of the model
      fill_f64_kernel<<<std::max(1, std::min((int)((uavg_len + 255) / 256), e->sm_count * 8)), 256, 0, e->stream>>>(m->uavg, (int64_t)uavg_len, -1.0);
      count_launch();
      if (cudaGetLastError() != cudaSuccess) s = MRS_ERR_CUDA;
    }
    if (s == MRS_OK) s = dev_alloc(&m->ipart, 2 * (size_t)R->ich.n_chunks);
    if (s == MRS_OK) s = dev_alloc(&m->xbuf, 3 * (size_t)R->n_items + 2);
    if (s == MRS_OK) s = dev_alloc(&m->idevavg, (size_t)R->n_items);
    if (s == MRS_OK) s = dev_alloc(&m->iavg, (size_t)R->n_items);
    if (s == MRS_OK) s = dev_alloc(&m->gavg, 1);
    if (s == MRS_OK) s = dev_alloc(&m->mae_part, (size_t)m->mae_part_cap);
    if (s == MRS_OK) s = dev_alloc(&m->counters, 8);
    if (s == MRS_OK && cudaMemsetAsync(m->counters, 0, 8 * sizeof(unsigned int), e->stream) != cudaSuccess) s = MRS_ERR_CUDA;
    if (s != MRS_OK) { mrs_model_destroy(m); return s; }
    *inout = m;
  }
  m->finished = false;
  m->host_valid = false;
  if (codes) {
    MRS_REQUIRE(!push || !m->want_item_avg, MRS_ERR_UNSUPPORTED, "mrs_fit_local_push: switch item averages off (mrs_model_set_item_averages): they do not travel in the fused exchange");
    MRS_REQUIRE(!no_finalize || !m->want_item_avg, MRS_ERR_UNSUPPORTED, "mrs_fit_mae_async: switch item averages off first (mrs_model_set_item_averages)");
    MRS_TRY(launch_fit_codes(e, R, m, fused_finalize, push, no_finalize));
    if (fused_finalize) m->finished = true;
    return MRS_OK;
  }
  MRS_REQUIRE(!push, MRS_ERR_UNSUPPORTED, "mrs_fit_local_push: the fused exchange needs a half-star coded rating set");
  MRS_TRY(launch_fit_local<double>(e, R, m));
  if (fused_finalize) return fit_finish(m);
  return MRS_OK;
}

// Users-only refit of an existing model: what the personalized / kNN predictors need from the fit (P:557-586 uses the user
// averages and the global average; the per-item average deviation belongs to the baseline predictor only and keeps the
// values of the last full fit -- the train set of a model never changes).  Two small kernels instead of the three of a
// full fit.
int32_t fit_users(mrs_engine* e, const mrs_ratings* R, mrs_model** inout) {
  MRS_REQUIRE(e && R && inout && *inout, MRS_ERR_INVALID, "mrs_fit_users_async: NULL argument (pass a fitted model of this train set)");
  mrs_model* m = *inout;
  MRS_REQUIRE(m->train == R && m->finished, MRS_ERR_INVALID, "mrs_fit_users_async: the model must have been fitted on this rating set");
  if (R->value_kind != kValueCode || R->n == 0) return fit_local(e, R, inout, true, nullptr, false);  // fp64-valued sets: the full fit
  use_engine(e);
  m->host_valid = false;
  user_sum_kernel<<<m->k1_blocks, 256, 0, e->stream>>>(R->uval16, R->vec_row, R->n_vec, m->usum, m->k1_part, e->d_timeline, 0);
  mark(e, "user_sum");
  MRS_CUDA(cudaGetLastError());
  const int grid = std::max(1, std::min((R->n_users + 255) / 256, e->sm_count * 4));
  MRS_CUDA(launch_pdl(user_avg_kernel, dim3(grid), dim3(256), 0, e->stream, m->usum, R->urow, R->n_users, m->k1_part, (double)R->n, m->uavg, m->gavg,
                      m->xbuf + 2 * (size_t)R->n_items));
  mark(e, "user_avg");
  MRS_CUDA(cudaGetLastError());
  return MRS_OK;
}

int32_t fit_finish(mrs_model* m) {
  MRS_REQUIRE(m, MRS_ERR_INVALID, "mrs_fit_finish: NULL model");
  use_engine(m->eng);
  MRS_CUDA(launch_pdl(item_finalize_kernel, dim3((m->n_items + 255) / 256), dim3(256), 0, m->eng->stream, m->xbuf, m->n_items, m->idevavg,
                      m->iavg, m->gavg));
  mark(m->eng, "item_finalize");
  MRS_CUDA(cudaGetLastError());
  m->finished = true;
  m->host_valid = false;
  return MRS_OK;
}

int32_t mae_baseline_async(const mrs_model* m, int32_t kind, const mrs_ratings* T, double* d_out2) {
  MRS_REQUIRE(m && T && d_out2, MRS_ERR_INVALID, "mrs_mae: NULL argument");
  use_engine(m->eng);
  MRS_REQUIRE(m->finished, MRS_ERR_INVALID, "mrs_mae: model not finished (call mrs_fit_finish)");
  MRS_REQUIRE(kind != MRS_PRED_ITEM || m->want_item_avg, MRS_ERR_INVALID, "mrs_mae: item averages were switched off for this model");
  if (kind == MRS_PRED_BASELINE && T->value_kind == kValueCode && T->n > 0) return launch_mae_tiled_baseline(m, T, d_out2);
  return T->value_kind == kValueCode ? dispatch_mae<uint8_t>(m, kind, T, d_out2) : dispatch_mae<double>(m, kind, T, d_out2);
}

int32_t predict_baseline_async(const mrs_model* m, int32_t kind, const int32_t* d_users, const int32_t* d_items, int64_t n,
                               double* d_out) {
  MRS_REQUIRE(m && m->finished, MRS_ERR_INVALID, "mrs_predict: model missing or not finished");
  use_engine(m->eng);
  MRS_REQUIRE(kind != MRS_PRED_ITEM || m->want_item_avg, MRS_ERR_INVALID, "mrs_predict: item averages were switched off for this model");
  if (n == 0) return MRS_OK;
  int grid = (int)std::min<int64_t>((n + 255) / 256, (int64_t)m->eng->sm_count * 16);
  cudaStream_t st = m->eng->stream;
#define MRS_LAUNCH_PRED(K) \
  predict_pairs_kernel<K><<<grid, 256, 0, st>>>(d_users, d_items, n, m->n_users, m->n_items, m->uavg, m->idevavg, m->iavg, m->gavg, d_out)
  switch (kind) {
    case MRS_PRED_GLOBAL: MRS_LAUNCH_PRED(MRS_PRED_GLOBAL); break;
    case MRS_PRED_USER: MRS_LAUNCH_PRED(MRS_PRED_USER); break;
    case MRS_PRED_ITEM: MRS_LAUNCH_PRED(MRS_PRED_ITEM); break;
    case MRS_PRED_ITEMDEV: MRS_LAUNCH_PRED(MRS_PRED_ITEMDEV); break;
    case MRS_PRED_BASELINE: MRS_LAUNCH_PRED(MRS_PRED_BASELINE); break;
    default: set_error("mrs_predict: predictor kind %d needs a similarity handle", kind); return MRS_ERR_INVALID;
  }
#undef MRS_LAUNCH_PRED
  mark(m->eng, "predict_pairs");
  MRS_CUDA(cudaGetLastError());
  return MRS_OK;
}

}  // namespace mrs
