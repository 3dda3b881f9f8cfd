// knn.cu -- personalized / kNN predictors of shared/predictions.scala:
//   preprocessedRating (P:470-481), adjustedCosineSimilarityFunction (P:407-433), jaccardCoefficient (P:440-464),
//   getNeighbors / getSimilarity (P:596-649), weightedSumDeviation (P:489-549), predictor (P:557-586).
//
// This file is the "similarity matrix fits on one GPU" path (ml-100k shape: U <= 16384 known users): S is
// materialised as a dense n_known x n_known fp64 matrix, every row is fully sorted once (so any k is a prefix,
// SURVEY A.6) and prediction looks similarities up by (u, v).
//
// Bit-exact neighbour sets need the similarity bits to equal the oracle's, so everything that feeds the ranking
// uses correctly-rounded, non-fused fp64 ops in ONE canonical order: deviations (sub, div), squared
// norm (sequential, ascending item id), r~ (div), dot product (sequential, ascending item id).  The similarity
// kernel is a shared-memory-staged SpGEMM: the dense r~ rows of a block of 32 users sit in shared memory (one lane per
// row user) and the sparse rows of the other users are streamed past them, one entry per warp step; items the row
// user did not rate contribute an exact +/-0 product, which leaves the running sum unchanged, so the result is
// bit-identical to a sum over the item intersection.  s(u,v) and s(v,u) are the same sequence of products: only one
// of the two is computed, the other one is its copy.
#include <algorithm>
#include <climits>
#include <cmath>
#include <numeric>
#include <vector>

#include "common.cuh"

namespace mrs {
namespace {

constexpr int32_t kMaxDenseUsers = 16384;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---------------- work partition of the similarity kernel (sparsity pattern only; cached on the rating set) --------
#ifndef MRS_WIDE_THREADS
#define MRS_WIDE_THREADS 512
#endif
constexpr int kWideThreads = MRS_WIDE_THREADS;
constexpr int kWideRows = 16;       // row users per CTA: one per lane of a half warp
constexpr int kWideUserCost = 8;   // what picking up a column user costs per phase, in entries (work partition)
// items per phase: the tile of items x 16 fp64 values and a 512-byte staging buffer per warp share the SM's 227 KB with one
// CTA (1,747 with 16 warps: ml-100k's 1,682 items are one phase)
constexpr int kWideMaxItems = (227 * 1024 - 512 - (kWideThreads / 32) * 512) / (kWideRows * 8) - 1;  // (- the all-zero row)
#ifndef MRS_WIDE_BATCH
#define MRS_WIDE_BATCH 8
#endif

// vmeta[p*npos + pos] = {compact index of the user at position pos of the length-sorted order (-1: padding), first and
// one-past-last CSR position of that user's entries with p*ic <= item < (p+1)*ic, 0}: everything a warp needs to know
// about a column user in phase p, in ONE 16-byte load
__global__ void vmeta_kernel(const int32_t* __restrict__ urow, const int32_t* __restrict__ ucol, const int32_t* __restrict__ known_user,
                             const int32_t* __restrict__ order, int32_t npos, int32_t ic, int32_t P, int4* __restrict__ vmeta) {
  const int32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= npos * P) return;
  const int32_t p = t / npos, pos = t % npos;
  const int32_t c = order[pos];
  int4 out = make_int4(-1, 0, 0, 0);
  if (c >= 0) {
    const int32_t u = known_user[c];
    const int32_t b = urow[u], e = urow[u + 1];
    int32_t bound[2];
    for (int k = 0; k < 2; ++k) {
      const int64_t want = (int64_t)(p + k) * ic;
      int32_t lo = b, hi = e;
      while (lo < hi) {  // items ascending inside a row
        const int32_t mid = (lo + hi) >> 1;
        if (ucol[mid] < want) lo = mid + 1; else hi = mid;
      }
      bound[k] = lo;
    }
    out = make_int4(c, bound[0], (p == P - 1) ? e : bound[1], 0);
  }
  vmeta[t] = out;
}

// compact user index of every CSC entry (what the prediction kernels look a rater's similarity up by)
__global__ void ccd_kernel(const int32_t* __restrict__ irow, const int32_t* __restrict__ cidx, int64_t n, int32_t* __restrict__ ccd) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < n; p += stride) ccd[p] = cidx[irow[p]];
}

}  // namespace

int32_t build_sim_layout(const mrs_ratings* R, bool with_wide) {
  mrs_ratings::sim_layout& L = R->sl;
  cudaStream_t st = R->eng->stream;
  if (!L.built) {
    std::vector<int32_t> urow((size_t)R->n_users + 1);
    MRS_CUDA(cudaMemcpyAsync(urow.data(), R->urow, sizeof(int32_t) * urow.size(), cudaMemcpyDeviceToHost, st));
    MRS_CUDA(cudaStreamSynchronize(st));
    std::vector<int32_t>& known = L.h_known;
    std::vector<int32_t> cidx((size_t)R->n_users, -1);
    known.clear();
    for (int32_t u = 0; u < R->n_users; ++u)
      if (urow[u + 1] > urow[u]) { cidx[u] = (int32_t)known.size(); known.push_back(u); }
    const int32_t nk = (int32_t)known.size();
    L.h_len.resize((size_t)nk);
    for (int32_t c = 0; c < nk; ++c) L.h_len[c] = urow[known[c] + 1] - urow[known[c]];
    L.h_order.resize((size_t)nk);
    std::iota(L.h_order.begin(), L.h_order.end(), 0);
    std::stable_sort(L.h_order.begin(), L.h_order.end(), [&](int32_t a, int32_t b) { return L.h_len[a] > L.h_len[b]; });
    L.n_known = nk;
    MRS_TRY(dev_alloc(&L.known_user, (size_t)nk));
    MRS_TRY(dev_alloc(&L.cidx, (size_t)R->n_users));
    MRS_CUDA(cudaMemcpyAsync(L.known_user, known.data(), sizeof(int32_t) * nk, cudaMemcpyHostToDevice, st));
    MRS_CUDA(cudaMemcpyAsync(L.cidx, cidx.data(), sizeof(int32_t) * cidx.size(), cudaMemcpyHostToDevice, st));
    MRS_CUDA(cudaStreamSynchronize(st));  // cidx goes out of scope
    L.built = true;
  }
  if (with_wide && !L.wide_built) {
    // Row users in blocks of kWideRows positions of the length-sorted order; block b is compared with the users at
    // positions >= kWideRows * b (s is symmetric: the rest is copied), so a block's work is the number of entries of those users.  Every
    // block gets CTAs in proportion to its work, each CTA a contiguous range of positions with an equal share of the
    // entries; inside a range the rows get shorter, so the warps that pick them up one by one finish together.
    const int32_t nk = L.n_known, nb = (nk + kWideRows - 1) / kWideRows, npos = (nk + 31) / 32 * 32;
    const int32_t P = std::max(1, (R->n_items + kWideMaxItems - 1) / kWideMaxItems);
    const int32_t ic = std::max(1, (R->n_items + P - 1) / P);
    std::vector<int32_t> order((size_t)npos, -1);
    for (int32_t t = 0; t < nk; ++t) order[t] = L.h_order[t];
    // (a user costs its entries plus a fixed amount per phase: descriptor, partial sum in and out)
    const int user_cost = getenv("MRS_WIDE_USERCOST") ? atoi(getenv("MRS_WIDE_USERCOST")) : kWideUserCost;
    std::vector<int64_t> suffix((size_t)nk + 1, 0);
    for (int32_t t = nk - 1; t >= 0; --t) suffix[t] = suffix[(size_t)t + 1] + L.h_len[order[t]] + (int64_t)user_cost * P;
    int64_t W = 0;
    for (int32_t b = 0; b < nb; ++b) W += suffix[(size_t)b * kWideRows];
    const int64_t G = R->eng->sm_count;
    auto ctas_of = [&](int32_t b, int64_t target) { return std::max<int64_t>(1, (suffix[(size_t)b * kWideRows] + target - 1) / target); };
    int64_t target = std::max<int64_t>(1, (W + G - 1) / G);
    if (nb <= G) {  // one wave: raise the share until the CTAs fit the SMs
      for (int it = 0; it < 200; ++it) {
        int64_t count = 0;
        for (int32_t b = 0; b < nb; ++b) count += ctas_of(b, target);
        if (count <= G) break;
        target += std::max<int64_t>(1, target / 50);
      }
    }
    std::vector<int4> desc;
    for (int32_t b = 0; b < nb; ++b) {
      const int64_t k = ctas_of(b, target), work = suffix[(size_t)b * kWideRows];
      int32_t pos = b * kWideRows;
      for (int64_t j = 0; j < k && pos < nk; ++j) {
        const int64_t stop = work - (work * (j + 1)) / k;  // entries left behind this CTA's range
        int32_t end = pos + 1;
        while (end < nk && suffix[end] > stop) ++end;
        if (j == k - 1) end = nk;
        desc.push_back(make_int4(b, pos, end, 0));
        pos = end;
      }
    }
    L.n_wide = (int32_t)desc.size();
    L.wide_ic = ic;
    L.wide_p = P;
    int32_t* d_order = nullptr;
    MRS_TRY(dev_alloc(&d_order, order.size()));
    MRS_TRY(dev_alloc(&L.wide_desc, std::max<size_t>(1, desc.size())));
    MRS_TRY(dev_alloc(&L.vmeta, std::max<size_t>(1, order.size() * (size_t)P)));
    MRS_CUDA(cudaMemcpyAsync(d_order, order.data(), sizeof(int32_t) * order.size(), cudaMemcpyHostToDevice, st));
    if (!desc.empty()) MRS_CUDA(cudaMemcpyAsync(L.wide_desc, desc.data(), sizeof(int4) * desc.size(), cudaMemcpyHostToDevice, st));
    MRS_TRY(dev_alloc(&L.ccd, std::max<size_t>(1, (size_t)R->n)));
    if (R->n > 0) {
      ccd_kernel<<<(int)std::max<int64_t>(1, std::min<int64_t>((R->n + 255) / 256, (int64_t)R->eng->sm_count * 8)), 256, 0, st>>>(R->irow, L.cidx, R->n, L.ccd);
      count_launch();
    }
    if (nk > 0) {
      vmeta_kernel<<<(npos * P + 255) / 256, 256, 0, st>>>(R->urow, R->ucol, L.known_user, d_order, npos, ic, P, L.vmeta);
      count_launch();
    }
    L.wide_npos = npos;
    MRS_CUDA(cudaGetLastError());
    MRS_CUDA(cudaStreamSynchronize(st));  // host vectors go out of scope
    dev_free(d_order);
    L.wide_built = true;
  }
  return MRS_OK;
}

namespace {

// ---------------- P1: deviations, squared norm, r~ (one warp per known user, canonical sequential order) -------------
template <typename VT>
__global__ void __launch_bounds__(256) dev_pre_kernel(const VT* __restrict__ uval, const int32_t* __restrict__ urow,
                                                     const int32_t* __restrict__ known_user, int32_t n_known,
                                                     const double* __restrict__ uavg, double* __restrict__ udev,
                                                     double* __restrict__ upre, double* __restrict__ unorm) {
  pdl_trigger();  // every kernel of the kNN closure is launched with programmatic dependent launch: its launch latency
  pdl_wait();     // overlaps the tail of its predecessor (7 kernels of 7-100 us each); its inputs are complete from here on
  const int lane = threadIdx.x & 31;
  const int32_t c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (c >= n_known) return;
  const int32_t u = known_user[c];
  const int32_t b = urow[u], e = urow[u + 1];
  const double a = uavg[u];
  double ss = 0.0;
  for (int32_t base = b; base < e; base += 32) {
    const int32_t p = base + lane;
    double d = 0.0;
    if (p < e) {
      d = deviation_fn(decode_value(uval[p]), a);  // P:167
      udev[p] = d;
    }
    const double d2 = __dmul_rn(d, d);  // math.pow(y, 2), P:474
    const int cnt = min(32, e - base);
    if (cnt == 32) {  // full chunk: the 32 broadcasts are independent and go out together, only the additions form a chain
      double t[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) t[j] = __shfl_sync(0xffffffffu, d2, j);
#pragma unroll
      for (int j = 0; j < 32; ++j) ss = __dadd_rn(ss, t[j]);  // ascending item id
    } else {
      for (int j = 0; j < cnt; ++j) ss = __dadd_rn(ss, __shfl_sync(0xffffffffu, d2, j));  // ascending item id
    }
  }
  const double w = __dsqrt_rn(ss);
  if (lane == 0) unorm[c] = w;
  for (int32_t p = b + lane; p < e; p += 32) upre[p] = (w != 0.0) ? __ddiv_rn(udev[p], w) : 0.0;  // P:478-479
}

// ---------------- P2: values in the layouts the next kernels stream (packed CSR for similarity, CSC for prediction) ----
template <int MODE>  // 0: deviations only (uniform); 1: cosine; 2: jaccard
__global__ void gather_values_kernel(const double* __restrict__ udev, const double* __restrict__ upre,
                                     const int32_t* __restrict__ csc_src, int64_t n, const int32_t* __restrict__ ucol,
                                     double* __restrict__ cdev, int4* __restrict__ pk) {
  pdl_trigger();
  pdl_wait();
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < n; p += stride) {
    cdev[p] = udev[csc_src[p]];
    if (MODE != 0) {  // one 16-byte record per rating: a warp fetches it with a single uniform load
      const double x = (MODE == 1) ? upre[p] : 1.0;
      pk[p] = make_int4(ucol[p], ucol[p] * (kWideRows * 8), __double2loint(x), __double2hiint(x));  // .y: byte offset of the item's tile row
    }
  }
}

// ---------------- P3: user-user similarity, shared-memory-staged SpGEMM --------------------------------------------
// One lane per ROW user, one column user per half warp: the tile holds r~ of the block's 16 row users as [item][16] (0.0
// where unrated); a half warp takes a column user v and walks v's entries in ascending item order -- the entry (item,
// r~_v) comes from a 16-byte broadcast read of the warp's staging buffer (filled 16 records at a time by one coalesced
// load), the 16 row values of that item from one conflict-free 128-byte shared-memory row -- so 32 products cost three
// memory wavefronts and no lane ever waits for another's item.  (The kernel this replaces held 8 row users per CTA and
// gave a lane one column user: every lane gathered the 64 bytes of its own item, 0.16 wavefronts per product at the
// measured conflict rate against 0.09 here, and it was bound by exactly that; profiles/r01_ncu_full_raw_knn_final.csv.)
// If 16 rows x all items do not fit shared memory the items are cut into wide_p phases of wide_ic items; the partial sums
// of a phase wait in S (the pair (row block, v) belongs to one CTA, the order of the additions is unchanged).  ml-100k's
// 1,682 items are one phase.
// Only v at or behind the block's own positions are visited: s(v,u) is the same sequence of products as s(u,v) and is
// written as its copy.  Positions are places in the length-sorted order, so long rows meet long rows once, not twice.
template <int MODE>  // 1: cosine (P:424-426), 2: jaccard (P:454-458)
__global__ void __launch_bounds__(kWideThreads, 1) similarity_wide_kernel(const int32_t* __restrict__ urow, const int4* __restrict__ pk,
                                                                        const int32_t* __restrict__ known_user, int32_t n_known,
                                                                        const int4* __restrict__ vmeta, int32_t npos,
                                                                        const int4* __restrict__ desc, int32_t ic, int32_t P,
                                                                        double* __restrict__ S, unsigned long long* __restrict__ tl) {
  extern __shared__ __align__(16) double tile[];  // [ic + 1][16] (the last row stays all zero), then one staging buffer of 2 x 16 records per warp
  tl_cta(tl, 0);  // (MRS_TIMELINE=1: per-CTA stamps, tools/knn_once.py)
  constexpr int nwarps = kWideThreads / 32;
  constexpr int kR = kWideRows;
  constexpr int kB = MRS_WIDE_BATCH;
  const int4 d = desc[blockIdx.x];
  const int32_t pos0 = d.x * kR, v0 = d.y, v1 = d.z;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int t = lane & (kR - 1), h = lane >> 4;  // row user of this lane, half warp
  int4* __restrict__ buf = reinterpret_cast<int4*>(tile + (size_t)(ic + 1) * kR) + wid * 32;
  const int32_t my_cu = vmeta[pos0 + t].x;  // this lane's row user (-1: padding behind the last user)
  const bool have = my_cu >= 0;
  int32_t nu = 0;
  if (MODE == 2 && have) { const int32_t u = known_user[my_cu]; nu = urow[u + 1] - urow[u]; }
  double* __restrict__ Srow = S + (int64_t)(have ? my_cu : 0) * n_known;
  const int4 none = make_int4(-1, 0, 0, 0);
  pdl_trigger();
  {
    double2* t2 = reinterpret_cast<double2*>(tile);
    for (int32_t x = threadIdx.x; x < (ic + 1) * (kR / 2); x += kWideThreads) t2[x] = make_double2(0.0, 0.0);
  }
  pdl_wait();  // the records are complete from here on; zeroing the tile overlapped the previous kernel
  for (int32_t p = 0; p < P; ++p) {
    const int32_t i0 = p * ic;
    const bool last = (p == P - 1);
    const int4* __restrict__ vm = vmeta + (size_t)p * npos;
    __syncthreads();
    for (int r = wid; r < kR; r += nwarps) {  // the row users' entries of this phase
      const int4 m = vm[pos0 + r];
      if (m.x >= 0) {
        for (int32_t q = m.y + lane; q < m.z; q += 32) {
          const int4 e = pk[q];
          tile[(e.x - i0) * kR + r] = __hiloint2double(e.w, e.z);
        }
      }
    }
    __syncthreads();
    if (p == 0) tl_cta(tl, 1);
    const unsigned char* __restrict__ col = reinterpret_cast<const unsigned char*>(tile + t) - (int64_t)i0 * (kR * 8);
    // what a half warp reads behind the end of its user: the all-zero row times 0.0 -- an exact +0 that leaves the sum as it is
    const int4 safe = make_int4(i0 + ic, (i0 + ic) * (kR * 8), 0, 0);
    // Column users are dealt out to the half warps round robin (the CTA's positions hold ever shorter rows, so every warp
    // gets the same mix and its two halves rows of nearly the same length).  The addresses of everything a warp will need
    // are known in advance: the descriptor of the user after next and the first records of the next one are in flight
    // during the sums of the current one, so no load latency is paid between two users.
    auto meta_of = [&](int32_t pos) { return (pos < v1) ? __ldg(vm + pos) : none; };
    auto chunk_of = [&](const int4& m, int32_t q0) { return (q0 < m.z) ? __ldg(pk + min(q0 + t, m.z - 1)) : safe; };
    int32_t pos_a = v0 + 2 * wid + h;
    int4 ma = meta_of(pos_a), mb = meta_of(pos_a + 2 * nwarps), mc = meta_of(pos_a + 4 * nwarps);
    int4 ent_a = chunk_of(ma, ma.y);
    while (__any_sync(0xffffffffu, pos_a < v1)) {
      const int4 ent_b = chunk_of(mb, mb.y);
      const int4 md = meta_of(pos_a + 6 * nwarps);
      const int32_t cv = ma.x, qa = ma.y, qb = ma.z;  // (a half without a user: -1, 0, 0)
      const bool skip = (qa == qb && p > 0 && !last);  // nothing to add in this phase: the partial sum stays where it is
      double acc = 0.0;
      if (p > 0 && have && cv >= 0 && !skip) acc = __ldcg(Srow + cv);
      int4 ent = ent_a;
      const int32_t len = qb - qa;
      const int32_t len_w = max(len, __shfl_xor_sync(0xffffffffu, len, 16));  // the longer of the warp's two rows
      for (int32_t c0 = 0; c0 < len_w; c0 += kR) {
        const int n = min(kR, len - c0);  // this half's records in the chunk (<= 0: none left)
        __syncwarp();
        buf[t * 2 + h] = (t < n) ? ent : safe;
        __syncwarp();
        ent = chunk_of(ma, qa + c0 + kR);  // the next chunk, requested ahead
#pragma unroll
        for (int k0 = 0; k0 < kR; k0 += kB) {
          if (k0 < len_w - c0) {  // warp-uniform
            int4 e[kB];
            double x[kB];
#pragma unroll
            for (int j = 0; j < kB; ++j) e[j] = buf[(k0 + j) * 2 + h];  // one address per half warp: one wavefront
#pragma unroll
            for (int j = 0; j < kB; ++j) x[j] = *reinterpret_cast<const double*>(col + e[j].y);  // the item's 16 row values: 128 contiguous bytes
#pragma unroll
            for (int j = 0; j < kB; ++j) acc = __dadd_rn(acc, __dmul_rn(x[j], __hiloint2double(e[j].w, e[j].z)));  // ascending item id, no FMA
          }
        }
      }
      if (have && cv >= 0 && !skip) {
        double s = acc;
        if (MODE == 2 && last) {
          const int32_t v = known_user[cv];
          const int32_t nv = urow[v + 1] - urow[v];
          s = acc / (double)(nu + nv - (int32_t)acc);  // P:458
        }
        __stcg(Srow + cv, s);
        if (last && pos_a >= pos0 + kR) __stcg(S + (int64_t)cv * n_known + my_cu, s);  // s(v,u): the same products in the same order
      }
      pos_a += 2 * nwarps; ma = mb; ent_a = ent_b;
      mb = mc; mc = md;
    }
    __syncthreads();  // every warp is done with the tile
    if (p == 0) tl_cta(tl, 2);
    if (!last) {  // take the row users' entries out again: the tile is all zero for the next phase
      for (int r = wid; r < kR; r += nwarps) {
        const int4 m = vm[pos0 + r];
        if (m.x >= 0)
          for (int32_t q = m.y + lane; q < m.z; q += 32) tile[(pk[q].x - i0) * kR + r] = 0.0;
      }
    }
  }
  tl_cta(tl, 3);
}

// ---------------- P4: full sort of every row by (similarity desc, user id asc); P:610 + SURVEY A.6 ----------------
__device__ __forceinline__ bool before(double ka, int32_t ia, double kb, int32_t ib) {
  return (ka > kb) || (ka == kb && ia < ib);
}
// A NaN similarity (a user whose average is exactly 1.0 has scale 0 below it: deviation -inf, r~ NaN, like the JVM) is not
// ordered by `before`; the sorting networks need a total order or they duplicate / drop elements and padding ids
// (INT_MAX) reach the output.  NaN keys therefore rank as -inf (last, ties by id); the reference's sortWith has no
// defined result for them either.
__device__ __forceinline__ double sort_key(double s) { return s != s ? -INFINITY : s; }

// in-shared-memory bitonic sort of P (power of two) (key,id) pairs into `before` order
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

__global__ void __launch_bounds__(512) sort_rank_kernel(const double* __restrict__ S, const int32_t* __restrict__ known_user,
                                                       int32_t n_known, int32_t P, int32_t* __restrict__ rank,
                                                       int32_t* __restrict__ nbr_id, double* __restrict__ nbr_sim,
                                                       const int32_t* __restrict__ tie_rank, const int32_t* __restrict__ tie_inv) {
  extern __shared__ double sh_key[];
  int32_t* sh_id = (int32_t*)(sh_key + P);
  pdl_trigger();
  pdl_wait();
  const int32_t cu = blockIdx.x;
  for (int32_t x = threadIdx.x; x < P; x += blockDim.x) {
    const bool cand = (x < n_known && x != cu);  // P:608 allUsers - u
    sh_key[x] = cand ? sort_key(S[(int64_t)cu * n_known + x]) : -INFINITY;
    sh_id[x] = cand ? (tie_rank ? tie_rank[x] : x) : INT_MAX;  // the tie key: the user's place in the tie order
  }
  bitonic_sort_shared(sh_key, sh_id, P);
  const int32_t nn = n_known - 1;
  for (int32_t j = threadIdx.x; j < nn; j += blockDim.x) {
    const int32_t cv = tie_inv ? tie_inv[sh_id[j]] : sh_id[j];
    nbr_id[(int64_t)cu * nn + j] = known_user[cv];
    nbr_sim[(int64_t)cu * nn + j] = sh_key[j];
    rank[(int64_t)cu * n_known + cv] = j;
  }
  if (threadIdx.x == 0) rank[(int64_t)cu * n_known + cu] = INT_MAX;  // s_k(u,u) = 0 (A.5)
}

// Same result for P <= 2048 with two elements per thread held in registers: compare-exchange partners at element
// distance 1 live in the same thread, at distance 2..32 in the same warp (shuffles), only distance >= 64 goes through
// shared memory -- 10 of the 55 stages of a 1024-element network need a barrier instead of all of them.
template <int P>
__global__ void __launch_bounds__(P / 2) sort_rank_reg_kernel(const double* __restrict__ S, const int32_t* __restrict__ known_user,
                                                             int32_t n_known, int32_t* __restrict__ rank, int32_t* __restrict__ nbr_id,
                                                             double* __restrict__ nbr_sim, const int32_t* __restrict__ tie_rank,
                                                             const int32_t* __restrict__ tie_inv) {
  __shared__ double sk[P];
  __shared__ int32_t si[P];
  pdl_trigger();
  pdl_wait();
  const int32_t cu = blockIdx.x;
  const int32_t e0 = 2 * threadIdx.x, e1 = e0 + 1;
  const bool c0 = (e0 < n_known && e0 != cu), c1 = (e1 < n_known && e1 != cu);  // P:608 allUsers - u
  double k0 = c0 ? sort_key(S[(int64_t)cu * n_known + e0]) : -INFINITY, k1 = c1 ? sort_key(S[(int64_t)cu * n_known + e1]) : -INFINITY;
  int32_t i0 = c0 ? (tie_rank ? tie_rank[e0] : e0) : INT_MAX, i1 = c1 ? (tie_rank ? tie_rank[e1] : e1) : INT_MAX;  // tie keys
#pragma unroll 1
  for (int32_t size = 2; size <= P; size <<= 1) {
    const bool up = ((e0 & size) == 0);
#pragma unroll 1
    for (int32_t stride = size >> 1; stride >= 64; stride >>= 1) {
      sk[e0] = k0; si[e0] = i0; sk[e1] = k1; si[e1] = i1;
      __syncthreads();
      const double pk0 = sk[e0 ^ stride], pk1 = sk[e1 ^ stride];
      const int32_t pi0 = si[e0 ^ stride], pi1 = si[e1 ^ stride];
      __syncthreads();
      // (similarity, id) pairs are distinct, so "partner after mine" is the negation of "partner before mine"; the only
      // equal pairs are padding slots, and swapping two of those changes nothing
      const bool keep_first = (((e0 & stride) == 0) == up);
      if (before(pk0, pi0, k0, i0) == keep_first) { k0 = pk0; i0 = pi0; }
      if (before(pk1, pi1, k1, i1) == keep_first) { k1 = pk1; i1 = pi1; }
    }
#pragma unroll
    for (int32_t stride = 32; stride >= 2; stride >>= 1) {
      if (stride < size) {
        const int m = stride >> 1;
        const double pk0 = __shfl_xor_sync(0xffffffffu, k0, m), pk1 = __shfl_xor_sync(0xffffffffu, k1, m);
        const int32_t pi0 = __shfl_xor_sync(0xffffffffu, i0, m), pi1 = __shfl_xor_sync(0xffffffffu, i1, m);
        const bool keep_first = (((e0 & stride) == 0) == up);
        if (before(pk0, pi0, k0, i0) == keep_first) { k0 = pk0; i0 = pi0; }
        if (before(pk1, pi1, k1, i1) == keep_first) { k1 = pk1; i1 = pi1; }
      }
    }
    // distance 1: both elements are this thread's
    if (up ? before(k1, i1, k0, i0) : before(k0, i0, k1, i1)) {
      const double tk = k0; k0 = k1; k1 = tk;
      const int32_t ti = i0; i0 = i1; i1 = ti;
    }
  }
  const int32_t nn = n_known - 1;
  if (e0 < nn) {
    const int32_t cv = tie_inv ? tie_inv[i0] : i0;
    nbr_id[(int64_t)cu * nn + e0] = known_user[cv];
    nbr_sim[(int64_t)cu * nn + e0] = k0;
    rank[(int64_t)cu * n_known + cv] = e0;
  }
  if (e1 < nn) {
    const int32_t cv = tie_inv ? tie_inv[i1] : i1;
    nbr_id[(int64_t)cu * nn + e1] = known_user[cv];
    nbr_sim[(int64_t)cu * nn + e1] = k1;
    rank[(int64_t)cu * n_known + cv] = e1;
  }
  if (threadIdx.x == 0) rank[(int64_t)cu * n_known + cu] = INT_MAX;  // s_k(u,u) = 0 (A.5)
}

// ---------------- P5: weighted-sum deviation + prediction (+ |error| reduction), one warp per (u,i) ----------------
// SIMMODE 0: uniform (s == 1, P:400); 1: matrix, no neighbourhood; 2: matrix restricted to the first k neighbours of u
template <int SIMMODE, bool WSD = false>
__device__ __forceinline__ double predict_pair(int32_t u, int32_t i, int lane, int32_t n_users, int32_t n_items,
                                               const double* __restrict__ uavg, double gavg, const int32_t* __restrict__ icolp,
                                               const int32_t* __restrict__ ccv, const double* __restrict__ cdev,
                                               const int32_t* __restrict__ cidx, const double* __restrict__ S,
                                               const int32_t* __restrict__ rank, int32_t n_known, int32_t k) {
  const double ua = (u >= 0 && u < n_users) ? uavg[u] : -1.0;
  if (ua < 0.0) {
    if (!WSD) return gavg;              // P:572-573
    if (SIMMODE != 0) return 0.0;       // wsd of a user without ratings: every similarity is 0 -> denominator 0 (P:527-529)
  }
  double num = 0.0, den = 0.0;
  if (i >= 0 && i < n_items) {
    const int32_t b = icolp[i], e = icolp[i + 1];
    if (SIMMODE == 0) {
      for (int32_t p = b + lane; p < e; p += 32) {
        num += cdev[p];  // P:522 with s == 1
        den += 1.0;
      }
    } else {
      // raters of the item, 4 x 32 at a time: the compact indices of all four batches are requested together, then the four
      // similarities and ranks -- two dependent L2 round trips per 128 raters instead of three per 32 (the user of a CSC
      // entry used to be looked up through irow and cidx).  Every lane adds its entries in the same order as before.
      const int64_t rowbase = (int64_t)cidx[u] * n_known;
      const double* __restrict__ Su = S + rowbase;
      const int32_t* __restrict__ ranku = rank + rowbase;
      for (int32_t p0 = b + lane; p0 < e; p0 += 128) {
        int32_t cv[4];
        double dv[4], sv[4];
        int32_t rk[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int32_t p = p0 + 32 * j;
          const bool ok = p < e;
          cv[j] = ok ? __ldg(ccv + p) : -1;
          dv[j] = ok ? cdev[p] : 0.0;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          sv[j] = cv[j] >= 0 ? Su[cv[j]] : 0.0;
          rk[j] = (SIMMODE == 2 && cv[j] >= 0) ? ranku[cv[j]] : 0;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (cv[j] >= 0) {
            const double sj = (SIMMODE == 2 && rk[j] >= k) ? 0.0 : sv[j];  // P:638-641
            num += dv[j] * sj;  // P:522
            den += fabs(sj);
          }
        }
      }
    }
  }
  num = warp_sum(num);
  den = warp_sum(den);
  const double w = den > 0.0 ? num / den : 0.0;  // P:527-529
  if (WSD) return w;
  return combine_fn(ua, w);                       // P:578
}

template <typename VT, int SIMMODE>
__global__ void __launch_bounds__(256) pers_mae_kernel(const int32_t* __restrict__ tu, const int32_t* __restrict__ ti,
                                                      const VT* __restrict__ tv, int64_t n, int32_t n_users, int32_t n_items,
                                                      const double* __restrict__ uavg, const double* __restrict__ gavg_p,
                                                      const int32_t* __restrict__ icolp, const int32_t* __restrict__ ccd,
                                                      const double* __restrict__ cdev, const int32_t* __restrict__ cidx,
                                                      const double* __restrict__ S, const int32_t* __restrict__ rank,
                                                      int32_t n_known, int32_t k, double* __restrict__ part,
                                                      unsigned int* __restrict__ counter, double* __restrict__ out2) {
  __shared__ double sh[8];
  __shared__ bool is_last;
  pdl_trigger();
  pdl_wait();
  const double gavg = gavg_p[0];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  double acc = 0.0;  // identical in every lane of the warp
  for (int64_t q = blockIdx.x * (int64_t)wpb + wid; q < n; q += (int64_t)gridDim.x * wpb) {
    const double pr = predict_pair<SIMMODE>(tu[q], ti[q], lane, n_users, n_items, uavg, gavg, icolp, ccd, cdev, cidx, S, rank, n_known, k);
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

template <int SIMMODE, bool WSD>
__global__ void __launch_bounds__(256) pers_pairs_kernel(const int32_t* __restrict__ us, const int32_t* __restrict__ is, int64_t n,
                                                        int32_t n_users, int32_t n_items, const double* __restrict__ uavg,
                                                        const double* __restrict__ gavg_p, const int32_t* __restrict__ icolp,
                                                        const int32_t* __restrict__ ccd, const double* __restrict__ cdev,
                                                        const int32_t* __restrict__ cidx, const double* __restrict__ S,
                                                        const int32_t* __restrict__ rank, int32_t n_known, int32_t k,
                                                        double* __restrict__ out) {
  const double gavg = gavg_p[0];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  for (int64_t q = blockIdx.x * (int64_t)wpb + wid; q < n; q += (int64_t)gridDim.x * wpb) {
    const double pr = predict_pair<SIMMODE, WSD>(us[q], is[q], lane, n_users, n_items, uavg, gavg, icolp, ccd, cdev, cidx, S, rank, n_known, k);
    if (lane == 0) out[q] = pr;
  }
}

int sim_mode(const mrs_sim* s) {
  if (s->kind == MRS_SIM_UNIFORM) return 0;
  return s->k > 0 ? 2 : 1;
}

template <int MODE>
int32_t dispatch_similarity(const mrs_ratings* R, mrs_sim* s, cudaStream_t st) {
  const auto& L = R->sl;
  if (L.n_wide == 0) return MRS_OK;
  const size_t smem = (size_t)(L.wide_ic + 1) * kWideRows * sizeof(double) + (size_t)(kWideThreads / 32) * 32 * sizeof(int4);
  MRS_CUDA(cudaFuncSetAttribute(similarity_wide_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  MRS_CUDA(launch_pdl(similarity_wide_kernel<MODE>, dim3(L.n_wide), dim3(kWideThreads), smem, st, R->urow, s->pk, L.known_user, L.n_known, L.vmeta,
                      L.wide_npos, L.wide_desc, L.wide_ic, L.wide_p, s->S, R->eng->d_timeline));
  mark(R->eng, "similarity");
  MRS_CUDA(cudaGetLastError());
  return MRS_OK;
}

}  // namespace

void free_sim_layout(const mrs_ratings* r) {
  auto& L = r->sl;
  dev_free(L.known_user); dev_free(L.cidx); dev_free(L.vmeta); dev_free(L.wide_desc); dev_free(L.ccd);
  free_rows_layout(r);
  L = mrs_ratings::sim_layout();
}

static int32_t sim_fit_impl(mrs_model* m, int32_t kind, int32_t k, mrs_sim** inout, bool force_rows, int32_t user_lo, int32_t user_hi) {
  MRS_REQUIRE(m && inout, MRS_ERR_INVALID, "mrs_fit_similarity: NULL argument");
  MRS_REQUIRE(m->finished, MRS_ERR_INVALID, "mrs_fit_similarity: model not finished");
  MRS_REQUIRE(kind == MRS_SIM_UNIFORM || kind == MRS_SIM_COSINE || kind == MRS_SIM_JACCARD, MRS_ERR_INVALID,
              "mrs_fit_similarity: unknown similarity kind %d", kind);
  MRS_REQUIRE(!(kind == MRS_SIM_UNIFORM && k > 0), MRS_ERR_UNSUPPORTED,
              "mrs_fit_similarity: a neighbourhood over the uniform similarity is not supported (no call site in the reference)");
  const mrs_ratings* R = m->train;
  mrs_engine* e = m->eng;
  cudaStream_t st = e->stream;
  use_engine(e);
  const bool matrix = (kind != MRS_SIM_UNIFORM);
  MRS_REQUIRE(!(force_rows && !matrix), MRS_ERR_INVALID, "mrs_fit_similarity_rows: the uniform similarity has no neighbour lists");
  MRS_TRY(build_sim_layout(R, false));
  const bool rows = matrix && (force_rows || R->sl.n_known > kMaxDenseUsers);
  if (matrix && !rows) MRS_TRY(build_sim_layout(R, true));
  const auto& L = R->sl;
  if (rows) {
    MRS_REQUIRE(k > 0, MRS_ERR_UNSUPPORTED,
                "mrs_fit_similarity: with %d users only the first k neighbours of a user are kept (row-block path); k must be > 0",
                L.n_known);
  }
  mrs_sim* s = *inout;
  if (s && (s->model != m || s->kind != kind || s->lists != rows)) {
    set_error("mrs_fit_similarity_async: the handle passed for reuse belongs to another model, similarity kind or path");
    return MRS_ERR_INVALID;
  }
  const bool first = !s;
  if (!s) {
    s = new mrs_sim();
    s->model = m;
    s->kind = kind;
    s->n_known = L.n_known;
    s->mae_part_cap = e->sm_count * 8;
    s->lists = rows;
    s->k = k;
    const size_t nk = (size_t)L.n_known;
    int32_t rc = MRS_OK;
    if (rc == MRS_OK) rc = dev_alloc(&s->udev, (size_t)R->n);
    if (rc == MRS_OK) rc = dev_alloc(&s->upre, (size_t)R->n);
    if (rc == MRS_OK) rc = dev_alloc(&s->unorm, nk);
    if (rc == MRS_OK && !rows) rc = dev_alloc(&s->cdev, (size_t)R->n);
    if (rc == MRS_OK) rc = dev_alloc(&s->mae_part, (size_t)s->mae_part_cap);
    if (rc == MRS_OK) rc = dev_alloc(&s->counter, 4);
    if (rc == MRS_OK && cudaMemsetAsync(s->counter, 0, 4 * sizeof(unsigned int), st) != cudaSuccess) rc = MRS_ERR_CUDA;
    if (matrix && !rows) {
      if (rc == MRS_OK) rc = dev_alloc(&s->pk, (size_t)R->n);
      if (rc == MRS_OK) rc = dev_alloc(&s->S, nk * nk);
      if (rc == MRS_OK) rc = dev_alloc(&s->rank, nk * nk);
      if (rc == MRS_OK) rc = dev_alloc(&s->nbr_id, nk * (nk ? nk - 1 : 0));
      if (rc == MRS_OK) rc = dev_alloc(&s->nbr_sim, nk * (nk ? nk - 1 : 0));
    }
    if (rc == MRS_OK && rows) rc = rows_alloc(m, s, user_lo, user_hi);
    if (rc != MRS_OK) { mrs_sim_destroy(s); return rc; }
    *inout = s;
  } else if (rows) {
    MRS_REQUIRE(k <= s->k_fit || s->k_fit >= s->n_known - 1, MRS_ERR_INVALID,
                "mrs_fit_similarity_async: the handle keeps %d neighbours per user, k = %d does not fit", s->k_fit, k);
  }
  s->k = k;
  if (L.n_known == 0) return MRS_OK;
  // P1
  const int wpb = 8;
  if (R->value_kind == kValueCode)
    MRS_CUDA(launch_pdl(dev_pre_kernel<uint8_t>, dim3((L.n_known + wpb - 1) / wpb), dim3(256), 0, st, (const uint8_t*)R->uval, R->urow, L.known_user,
                        L.n_known, m->uavg, s->udev, s->upre, s->unorm));
  else
    MRS_CUDA(launch_pdl(dev_pre_kernel<double>, dim3((L.n_known + wpb - 1) / wpb), dim3(256), 0, st, (const double*)R->uval, R->urow, L.known_user,
                        L.n_known, m->uavg, s->udev, s->upre, s->unorm));
  mark(e, "dev_pre");
  if (rows) {
    MRS_CUDA(cudaGetLastError());
    return rows_fit_async(m, s, first);
  }
  // P2
  const int64_t work = R->n;
  const int g2 = (int)std::max<int64_t>(1, std::min<int64_t>((work + 255) / 256, (int64_t)e->sm_count * 8));
  if (kind == MRS_SIM_UNIFORM) MRS_CUDA(launch_pdl(gather_values_kernel<0>, dim3(g2), dim3(256), 0, st, s->udev, s->upre, R->csc_src, R->n, R->ucol, s->cdev, s->pk));
  else if (kind == MRS_SIM_COSINE) MRS_CUDA(launch_pdl(gather_values_kernel<1>, dim3(g2), dim3(256), 0, st, s->udev, s->upre, R->csc_src, R->n, R->ucol, s->cdev, s->pk));
  else MRS_CUDA(launch_pdl(gather_values_kernel<2>, dim3(g2), dim3(256), 0, st, s->udev, s->upre, R->csc_src, R->n, R->ucol, s->cdev, s->pk));
  mark(e, "gather_values");
  MRS_CUDA(cudaGetLastError());
  if (!matrix) return MRS_OK;
  // P3
  if (kind == MRS_SIM_COSINE) MRS_TRY(dispatch_similarity<1>(R, s, st));
  else MRS_TRY(dispatch_similarity<2>(R, s, st));
  // P4
  int32_t P = 2;
  while (P < L.n_known) P <<= 1;
  if (P <= 512) {
    MRS_CUDA(launch_pdl(sort_rank_reg_kernel<512>, dim3(L.n_known), dim3(256), 0, st, s->S, L.known_user, L.n_known, s->rank, s->nbr_id, s->nbr_sim, m->tie_rank, m->tie_inv));
  } else if (P == 1024) {
    MRS_CUDA(launch_pdl(sort_rank_reg_kernel<1024>, dim3(L.n_known), dim3(512), 0, st, s->S, L.known_user, L.n_known, s->rank, s->nbr_id, s->nbr_sim, m->tie_rank, m->tie_inv));
  } else if (P == 2048) {
    MRS_CUDA(launch_pdl(sort_rank_reg_kernel<2048>, dim3(L.n_known), dim3(1024), 0, st, s->S, L.known_user, L.n_known, s->rank, s->nbr_id, s->nbr_sim, m->tie_rank, m->tie_inv));
  } else {
    const size_t smem = (size_t)P * (sizeof(double) + sizeof(int32_t));
    MRS_CUDA(cudaFuncSetAttribute(sort_rank_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    MRS_CUDA(launch_pdl(sort_rank_kernel, dim3(L.n_known), dim3(512), smem, st, s->S, L.known_user, L.n_known, P, s->rank, s->nbr_id, s->nbr_sim, m->tie_rank, m->tie_inv));
  }
  mark(e, "sort_rank");
  MRS_CUDA(cudaGetLastError());
  return MRS_OK;
}

int32_t sim_fit_async(mrs_model* m, int32_t kind, int32_t k, mrs_sim** inout) {
  return sim_fit_impl(m, kind, k, inout, false, 0, INT_MAX);
}

int32_t mae_personalized_async(const mrs_model* m, const mrs_sim* s, const mrs_ratings* T, double* d_out2) {
  MRS_REQUIRE(m && s && T && d_out2, MRS_ERR_INVALID, "mrs_mae: NULL argument");
  MRS_REQUIRE(s->model == m, MRS_ERR_INVALID, "mrs_mae: similarity handle belongs to another model");
  if (s->lists) return mae_lists_async(m, s, T, d_out2);
  const mrs_ratings* R = m->train;
  const auto& L = R->sl;
  cudaStream_t st = m->eng->stream;
  const int wpb = 8;
  int grid = (int)std::max<int64_t>(1, std::min<int64_t>((T->n + wpb - 1) / wpb, (int64_t)s->mae_part_cap));
  const int mode = sim_mode(s);
#define MRS_PERS_MAE(VT, MODE)                                                                                                   \
  MRS_CUDA(launch_pdl(pers_mae_kernel<VT, MODE>, dim3(grid), dim3(256), 0, st, T->coo_u, T->ucol, (const VT*)T->uval, T->n, m->n_users,   \
                      m->n_items, m->uavg, m->gavg, R->icolp, L.ccd, s->cdev, L.cidx, s->S, s->rank, L.n_known, s->k, s->mae_part, \
                      s->counter, d_out2))
  if (T->value_kind == kValueCode) {
    if (mode == 0) MRS_PERS_MAE(uint8_t, 0); else if (mode == 1) MRS_PERS_MAE(uint8_t, 1); else MRS_PERS_MAE(uint8_t, 2);
  } else {
    if (mode == 0) MRS_PERS_MAE(double, 0); else if (mode == 1) MRS_PERS_MAE(double, 1); else MRS_PERS_MAE(double, 2);
  }
#undef MRS_PERS_MAE
  mark(m->eng, "pers_mae");
  MRS_CUDA(cudaGetLastError());
  return MRS_OK;
}

int32_t predict_personalized_async(const mrs_model* m, const mrs_sim* s, const int32_t* d_users, const int32_t* d_items, int64_t n,
                                   double* d_out, bool wsd_only) {
  MRS_REQUIRE(m && s, MRS_ERR_INVALID, "mrs_predict: NULL argument");
  MRS_REQUIRE(s->model == m, MRS_ERR_INVALID, "mrs_predict: similarity handle belongs to another model");
  if (n == 0) return MRS_OK;
  if (s->lists) return predict_lists_async(m, s, d_users, d_items, n, d_out, wsd_only);
  const mrs_ratings* R = m->train;
  const auto& L = R->sl;
  cudaStream_t st = m->eng->stream;
  const int wpb = 8;
  int grid = (int)std::max<int64_t>(1, std::min<int64_t>((n + wpb - 1) / wpb, (int64_t)m->eng->sm_count * 16));
  const int mode = sim_mode(s);
#define MRS_PERS_PAIRS(MODE, W)                                                                                               \
  pers_pairs_kernel<MODE, W><<<grid, 256, 0, st>>>(d_users, d_items, n, m->n_users, m->n_items, m->uavg, m->gavg, R->icolp, L.ccd, \
                                                   s->cdev, L.cidx, s->S, s->rank, L.n_known, s->k, d_out)
  if (wsd_only) {
    if (mode == 0) MRS_PERS_PAIRS(0, true); else if (mode == 1) MRS_PERS_PAIRS(1, true); else MRS_PERS_PAIRS(2, true);
  } else {
    if (mode == 0) MRS_PERS_PAIRS(0, false); else if (mode == 1) MRS_PERS_PAIRS(1, false); else MRS_PERS_PAIRS(2, false);
  }
#undef MRS_PERS_PAIRS
  mark(m->eng, "pers_pairs");
  MRS_CUDA(cudaGetLastError());
  return MRS_OK;
}

}  // namespace mrs

// ------------------------------------------------------------------ C ABI
using namespace mrs;

extern "C" int32_t mrs_fit_similarity_async(mrs_model* m, int32_t sim_kind, int32_t k, mrs_sim** inout) {
  return sim_fit_async(m, sim_kind, k, inout);
}

extern "C" int32_t mrs_fit_similarity_rows_async(mrs_model* m, int32_t sim_kind, int32_t k, int32_t user_lo, int32_t user_hi, mrs_sim** inout) {
  return sim_fit_impl(m, sim_kind, k, inout, true, user_lo, user_hi);
}

extern "C" int32_t mrs_fit_similarity(mrs_model* m, int32_t sim_kind, int32_t k, mrs_sim** out) {
  MRS_REQUIRE(out, MRS_ERR_INVALID, "mrs_fit_similarity: NULL output");
  *out = nullptr;
  MRS_TRY(sim_fit_async(m, sim_kind, k, out));
  cudaError_t ce = cudaStreamSynchronize(m->eng->stream);
  if (ce != cudaSuccess) {
    set_error("mrs_fit_similarity: %s", cudaGetErrorString(ce));
    mrs_sim_destroy(*out);
    *out = nullptr;
    return MRS_ERR_CUDA;
  }
  return MRS_OK;
}

// place of an Int in the iteration order of a Scala 2.11 immutable.HashSet[Int]: the hash trie is walked in
// ascending order of successive 5-bit groups, least significant first, of improve(hashCode) -- recalled from the 2.11
// standard library (scala.collection.immutable.HashSet.improve), not verifiable here without a JVM
static uint64_t hashset_order_key(int32_t id) {
  uint32_t h = (uint32_t)id + ~((uint32_t)id << 9);
  h ^= h >> 14;
  h += h << 4;
  h ^= h >> 10;
  uint64_t key = 0;
  for (int k = 0; k < 7; ++k) key = (key << 5) | ((h >> (5 * k)) & 31u);
  return key;
}

extern "C" int32_t mrs_model_set_tie_order(mrs_model* m, int32_t mode) {
  MRS_REQUIRE(m && (mode == 0 || mode == 1), MRS_ERR_INVALID, "mrs_model_set_tie_order: mode must be 0 (user id) or 1 (Scala 2.11 HashSet order)");
  use_engine(m->eng);
  dev_free(m->tie_rank); dev_free(m->tie_inv);
  m->tie_rank = m->tie_inv = nullptr;
  m->tie_mode = mode;
  if (mode == 0) return MRS_OK;
  MRS_TRY(build_sim_layout(m->train, false));
  const auto& L = m->train->sl;
  const int32_t nk = L.n_known;
  std::vector<int32_t> order((size_t)nk), rank((size_t)nk);
  for (int32_t c = 0; c < nk; ++c) order[(size_t)c] = c;
  std::sort(order.begin(), order.end(), [&](int32_t a, int32_t b) { return hashset_order_key(L.h_known[(size_t)a]) < hashset_order_key(L.h_known[(size_t)b]); });
  for (int32_t r = 0; r < nk; ++r) rank[(size_t)order[(size_t)r]] = r;
  MRS_TRY(dev_alloc(&m->tie_rank, (size_t)std::max(nk, 1)));
  MRS_TRY(dev_alloc(&m->tie_inv, (size_t)std::max(nk, 1)));
  MRS_CUDA(cudaMemcpy(m->tie_rank, rank.data(), sizeof(int32_t) * (size_t)nk, cudaMemcpyHostToDevice));
  MRS_CUDA(cudaMemcpy(m->tie_inv, order.data(), sizeof(int32_t) * (size_t)nk, cudaMemcpyHostToDevice));
  return MRS_OK;
}

extern "C" int32_t mrs_sim_set_k(mrs_sim* s, int32_t k) {
  MRS_REQUIRE(s, MRS_ERR_INVALID, "mrs_sim_set_k: NULL handle");
  MRS_REQUIRE(!s->lists || (k > 0 && (k <= s->k_fit || s->k_fit >= s->n_known - 1)), MRS_ERR_UNSUPPORTED,
              "mrs_sim_set_k: this handle keeps the first %d neighbours of each user; k = %d is outside (0, %d]", s->k_fit, k, s->k_fit);
  s->k = k;
  return MRS_OK;
}

extern "C" int32_t mrs_sim_entry_values(const mrs_sim* s, int32_t which, int32_t* users_out, int32_t* items_out, double* vals_out,
                                        int64_t cap, int64_t* n_out) {
  MRS_REQUIRE(s && n_out, MRS_ERR_INVALID, "mrs_sim_entry_values: NULL argument");
  use_engine(s->model->eng);
  MRS_REQUIRE(which == 0 || which == 1, MRS_ERR_INVALID, "mrs_sim_entry_values: which must be 0 (deviation) or 1 (preprocessed)");
  const mrs_ratings* R = s->model->train;
  *n_out = R->n;
  if (!users_out && !items_out && !vals_out) return MRS_OK;
  MRS_REQUIRE(cap >= R->n, MRS_ERR_INVALID, "mrs_sim_entry_values: capacity %lld < %lld", (long long)cap, (long long)R->n);
  cudaStream_t st = R->eng->stream;
  if (users_out) MRS_CUDA(cudaMemcpyAsync(users_out, R->coo_u, sizeof(int32_t) * R->n, cudaMemcpyDeviceToHost, st));
  if (items_out) MRS_CUDA(cudaMemcpyAsync(items_out, R->ucol, sizeof(int32_t) * R->n, cudaMemcpyDeviceToHost, st));
  if (vals_out) MRS_CUDA(cudaMemcpyAsync(vals_out, which == 0 ? s->udev : s->upre, sizeof(double) * R->n, cudaMemcpyDeviceToHost, st));
  MRS_CUDA(cudaStreamSynchronize(st));
  return MRS_OK;
}

extern "C" void mrs_sim_destroy(mrs_sim* s) {
  if (!s) return;
  if (s->model && s->model->eng) use_engine(s->model->eng);
  dev_free(s->udev); dev_free(s->upre); dev_free(s->unorm); dev_free(s->cdev); dev_free(s->pk); dev_free(s->S);
  dev_free(s->rank); dev_free(s->nbr_id); dev_free(s->nbr_sim); dev_free(s->mae_part); dev_free(s->counter);
  rows_free(s);
  delete s;
}

static int32_t host_user_len(const mrs_ratings* R, int32_t u, int32_t* len) {
  *len = 0;
  if (u < 0 || u >= R->n_users) return MRS_OK;
  int32_t p[2];
  MRS_CUDA(cudaMemcpyAsync(p, R->urow + u, 2 * sizeof(int32_t), cudaMemcpyDeviceToHost, R->eng->stream));
  MRS_CUDA(cudaStreamSynchronize(R->eng->stream));
  *len = p[1] - p[0];
  return MRS_OK;
}

static int32_t host_cidx(const mrs_ratings* R, int32_t u, int32_t* c) {
  *c = -1;
  if (u < 0 || u >= R->n_users) return MRS_OK;
  MRS_CUDA(cudaMemcpyAsync(c, R->sl.cidx + u, sizeof(int32_t), cudaMemcpyDeviceToHost, R->eng->stream));
  MRS_CUDA(cudaStreamSynchronize(R->eng->stream));
  return MRS_OK;
}

extern "C" int32_t mrs_similarity(const mrs_sim* s, int32_t u, int32_t v, double* out) {
  MRS_REQUIRE(s && out, MRS_ERR_INVALID, "mrs_similarity: NULL argument");
  use_engine(s->model->eng);
  const mrs_ratings* R = s->model->train;
  cudaStream_t st = R->eng->stream;
  int32_t cu = -1, cv = -1;
  MRS_TRY(host_cidx(R, u, &cu));
  MRS_TRY(host_cidx(R, v, &cv));
  if (s->kind == MRS_SIM_UNIFORM) { *out = 1.0; return MRS_OK; }  // P:400
  if (cu < 0 || cv < 0) {
    if (s->kind == MRS_SIM_JACCARD && s->k <= 0) {
      int32_t lu = 0, lv = 0;
      MRS_TRY(host_user_len(R, u, &lu));
      MRS_TRY(host_user_len(R, v, &lv));
      *out = (lu + lv) ? 0.0 : nan("");  // 0/0 on the JVM (P:458)
    } else {
      *out = 0.0;  // empty intersection -> empty sum (P:424-426); zero-similarity fillers weigh 0 either way
    }
    return MRS_OK;
  }
  if (s->lists) {  // s_k(u, v): the similarity if v is among the first k neighbours of u, else 0 (P:638-641)
    MRS_REQUIRE(cu >= s->row_lo && cu < s->row_hi, MRS_ERR_INVALID, "mrs_similarity: user %d is outside the row range of this handle", u);
    const int32_t kk = std::min(s->k, s->k_fit);
    std::vector<int32_t> ids((size_t)std::max(kk, 1));
    std::vector<double> sims((size_t)std::max(kk, 1));
    const int64_t off = (int64_t)(cu - s->row_lo) * s->k_fit;
    MRS_CUDA(cudaMemcpyAsync(ids.data(), s->nbr_id + off, sizeof(int32_t) * kk, cudaMemcpyDeviceToHost, st));
    MRS_CUDA(cudaMemcpyAsync(sims.data(), s->nbr_sim + off, sizeof(double) * kk, cudaMemcpyDeviceToHost, st));
    MRS_CUDA(cudaStreamSynchronize(st));
    *out = 0.0;
    for (int32_t j = 0; j < kk; ++j)
      if (ids[(size_t)j] == v) { *out = sims[(size_t)j]; break; }
    return MRS_OK;
  }
  double val = 0.0;
  int32_t rk = 0;
  const int64_t off = (int64_t)cu * s->n_known + cv;
  MRS_CUDA(cudaMemcpyAsync(&val, s->S + off, sizeof(double), cudaMemcpyDeviceToHost, st));
  MRS_CUDA(cudaMemcpyAsync(&rk, s->rank + off, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  MRS_CUDA(cudaStreamSynchronize(st));
  *out = (s->k > 0 && rk >= s->k) ? 0.0 : val;
  return MRS_OK;
}

extern "C" int32_t mrs_neighbors(const mrs_sim* s, int32_t u, int32_t k, int32_t* ids_out, double* sims_out, int32_t cap, int32_t* n_out) {
  MRS_REQUIRE(s && n_out, MRS_ERR_INVALID, "mrs_neighbors: NULL argument");
  use_engine(s->model->eng);
  const mrs_ratings* R = s->model->train;
  cudaStream_t st = R->eng->stream;
  int32_t cu = -1;
  MRS_TRY(host_cidx(R, u, &cu));
  const int32_t nk = s->n_known;
  const int32_t cand = nk - (cu >= 0 ? 1 : 0);
  int32_t w = std::max(0, std::min(std::min(k, cand), cap));
  *n_out = w;
  if (w == 0) return MRS_OK;
  if (s->lists && cu >= 0) {
    MRS_REQUIRE(cu >= s->row_lo && cu < s->row_hi, MRS_ERR_INVALID, "mrs_neighbors: user %d is outside the row range of this handle", u);
    MRS_REQUIRE(k <= s->k_fit || s->k_fit >= nk - 1, MRS_ERR_UNSUPPORTED,
                "mrs_neighbors: this handle keeps the first %d neighbours of each user, %d were asked for", s->k_fit, k);
    w = std::min(w, s->k_fit);
    *n_out = w;
    const int64_t off = (int64_t)(cu - s->row_lo) * s->k_fit;
    if (ids_out) MRS_CUDA(cudaMemcpyAsync(ids_out, s->nbr_id + off, sizeof(int32_t) * w, cudaMemcpyDeviceToHost, st));
    if (sims_out) MRS_CUDA(cudaMemcpyAsync(sims_out, s->nbr_sim + off, sizeof(double) * w, cudaMemcpyDeviceToHost, st));
    MRS_CUDA(cudaStreamSynchronize(st));
    return MRS_OK;
  }
  if (s->kind != MRS_SIM_UNIFORM && cu >= 0) {
    const int64_t off = (int64_t)cu * (nk - 1);
    if (ids_out) MRS_CUDA(cudaMemcpyAsync(ids_out, s->nbr_id + off, sizeof(int32_t) * w, cudaMemcpyDeviceToHost, st));
    if (sims_out) MRS_CUDA(cudaMemcpyAsync(sims_out, s->nbr_sim + off, sizeof(double) * w, cudaMemcpyDeviceToHost, st));
    MRS_CUDA(cudaStreamSynchronize(st));
    return MRS_OK;
  }
  // all similarities equal (uniform: 1.0; unknown u: 0.0) -> stable sort keeps ascending user id
  std::vector<int32_t> known((size_t)std::min(nk, w + 1));
  MRS_CUDA(cudaMemcpyAsync(known.data(), R->sl.known_user, sizeof(int32_t) * known.size(), cudaMemcpyDeviceToHost, st));
  MRS_CUDA(cudaStreamSynchronize(st));
  const double val = (s->kind == MRS_SIM_UNIFORM) ? 1.0 : 0.0;
  int32_t j = 0;
  for (size_t t = 0; t < known.size() && j < w; ++t) {
    if (known[t] == u) continue;
    if (ids_out) ids_out[j] = known[t];
    if (sims_out) sims_out[j] = val;
    ++j;
  }
  *n_out = j;
  return MRS_OK;
}
