// loader.cu -- host COO (Rating(user,item,rating), P:9) -> device-resident user-major CSR, item-major CSC
// and sorted COO.  Replaces `load(...).collect()` materialisation (P:35-49).  Not on the timed hot path of the
// reference either (predict/Baseline.scala:40-42 load before timing); it is inside bench.py's e2e figure.
//
// The three stable sorts (by item, by user, and by item again for the item-major order) use cub::DeviceRadixSort (library
// code, like calling cuBLAS); everything else is ours.
#include <cub/cub.cuh>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "common.cuh"

namespace mrs {

namespace {

// passes over the raw input: id ranges (needed before the sorts) and whether every rating is a half-star code (needed when
// the values are gathered); separate kernels because the items arrive before the users and the ids before the ratings
// stats of one id array: maximum and minimum (the items travel first and are scanned while the users are still on the link)
__global__ void scan_id_kernel(const int32_t* __restrict__ a, int64_t n, int32_t* __restrict__ max_out, int32_t* __restrict__ min_out) {
  int32_t mx = -1, mn = 0x7fffffff;
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x) {
    const int32_t v = a[p];
    mx = max(mx, v);
    mn = min(mn, v);
  }
  for (int o = 16; o > 0; o >>= 1) {
    mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMax(max_out, mx);
    atomicMin(min_out, mn);
  }
}

// user of every entry of the item-sorted order: the key of the second (stable) sort
__global__ void user_keys_kernel(const int32_t* __restrict__ d_u, const int32_t* __restrict__ perm_i, int64_t n, uint32_t* __restrict__ keys) {
  for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < n; q += (int64_t)gridDim.x * blockDim.x) keys[q] = (uint32_t)d_u[perm_i[q]];
}

// sorted users + raw position of every user-major entry -> user array, item array, row pointer, duplicate count (ids only: runs
// before the ratings have arrived)
__global__ void scatter_users_kernel(const uint32_t* __restrict__ users, const int32_t* __restrict__ perm_u, const int32_t* __restrict__ d_i, int64_t n,
                                     int32_t n_users, int32_t* __restrict__ coo_u, int32_t* __restrict__ ucol, int32_t* __restrict__ urow,
                                     int32_t* __restrict__ dup_count) {
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x) {
    const int32_t user = (int32_t)users[p], item = d_i[perm_u[p]];
    coo_u[p] = user;
    ucol[p] = item;
    int32_t prev = -1;
    if (p > 0) {
      prev = (int32_t)users[p - 1];
      if (prev == user && d_i[perm_u[p - 1]] == item) atomicAdd(dup_count, 2);  // (counted twice: the message halves it)
    }
    for (int32_t s = prev + 1; s <= user; ++s) urow[s] = (int32_t)p;
    if (p == n - 1)
      for (int32_t s = user + 1; s <= n_users; ++s) urow[s] = (int32_t)n;
  }
}

__global__ void scan_values_kernel(const double* __restrict__ r, int64_t n, int32_t* __restrict__ stats) {
  // stats[3] = number of ratings that are not k*0.5 in [0,127]: code 0xFF is the padding mark of the item-tiled test
  // layout (mae_tiled.cu), so a rating of 127.5 sends the set to the fp64-valued kernels instead of being dropped
  int32_t bad = 0;
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x) {
    double v = r[p] * 2.0;
    if (!(v >= 0.0 && v <= 254.0 && v == floor(v))) bad++;
  }
  for (int o = 16; o > 0; o >>= 1) bad += __shfl_xor_sync(0xffffffffu, bad, o);
  if ((threadIdx.x & 31) == 0 && bad) atomicAdd(&stats[3], bad);
}

// stats[3] = number of codes equal to 0xFF (the padding mark of the item-tiled test layout: not a legal code)
__global__ void scan_codes_kernel(const uint8_t* __restrict__ c, int64_t n, int32_t* __restrict__ stats) {
  int32_t bad = 0;
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x) bad += (c[p] == 0xFF);
  for (int o = 16; o > 0; o >>= 1) bad += __shfl_xor_sync(0xffffffffu, bad, o);
  if ((threadIdx.x & 31) == 0 && bad) atomicAdd(&stats[3], bad);
}

// item-major order from the user-major one: a STABLE sort on the item alone keeps the users of an item ascending
__global__ void item_keys_kernel(const int32_t* __restrict__ ucol, int64_t n, uint32_t* __restrict__ keys, int32_t* __restrict__ idx) {
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x) {
    keys[p] = (uint32_t)ucol[p];
    idx[p] = (int32_t)p;
  }
}
// sorted item keys + positions in the user-major arrays -> rater of every item-major entry and the column pointer
__global__ void scatter_items_kernel(const uint32_t* __restrict__ keys, const int32_t* __restrict__ src, const int32_t* __restrict__ coo_u, int64_t n,
                                     int32_t n_items, int32_t* __restrict__ irow, int32_t* __restrict__ icolp) {
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x) {
    const int32_t item = (int32_t)keys[p];
    irow[p] = coo_u[src[p]];
    const int32_t prev = p > 0 ? (int32_t)keys[p - 1] : -1;
    for (int32_t s = prev + 1; s <= item; ++s) icolp[s] = (int32_t)p;
    if (p == n - 1)
      for (int32_t s = item + 1; s <= n_items; ++s) icolp[s] = (int32_t)n;
  }
}

// values into sorted order: kEncode reads the raw fp64 ratings (and encodes half-star codes), else copies VT values
template <typename VT, bool kEncode>
__global__ void gather_sorted_kernel(const int32_t* __restrict__ src, int64_t n, const void* __restrict__ values_in, VT* __restrict__ values_out) {
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x) {
    if (kEncode) {
      double v = ((const double*)values_in)[src[p]];
      if (sizeof(VT) == 1) values_out[p] = (VT)(v * 2.0); else values_out[p] = (VT)v;
    } else {
      values_out[p] = ((const VT*)values_in)[src[p]];
    }
  }
}

__global__ void chunk_count_kernel(const int32_t* __restrict__ seg_ptr, int32_t n_seg, int32_t chunk, int32_t* __restrict__ counts) {
  int32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s < n_seg) {
    int32_t len = seg_ptr[s + 1] - seg_ptr[s];
    counts[s] = (len + chunk - 1) / chunk;
  }
}

__global__ void chunk_fill_kernel(const int32_t* __restrict__ seg_ptr, const int32_t* __restrict__ seg_chunk_ptr, int32_t n_seg,
                                  int32_t chunk, int32_t* __restrict__ chunk_seg, int32_t* __restrict__ chunk_begin) {
  int32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s < n_seg) {
    int32_t c0 = seg_chunk_ptr[s], c1 = seg_chunk_ptr[s + 1], b = seg_ptr[s];
    for (int32_t c = c0; c < c1; ++c) {
      chunk_seg[c] = s;
      chunk_begin[c] = b + (c - c0) * chunk;
    }
  }
}

int grid_for(int64_t n, int block, int sm_count) {
  int64_t g = (n + block - 1) / block;
  int64_t cap = (int64_t)sm_count * 16;
  return (int)std::max<int64_t>(1, std::min(g, cap));
}

int bits_for(uint32_t v) {
  int b = 0;
  while (v) { ++b; v >>= 1; }
  return std::max(b, 1);
}

__global__ void vec_count_kernel(const int32_t* __restrict__ urow, int32_t n_users, int32_t* __restrict__ cnt) {
  const int32_t u = blockIdx.x * blockDim.x + threadIdx.x;
  if (u < n_users) cnt[u] = (urow[u + 1] - urow[u] + 15) >> 4;
}

__global__ void pad_fill_kernel(const uint8_t* __restrict__ uval, const int32_t* __restrict__ coo_u, const int32_t* __restrict__ urow,
                                const int32_t* __restrict__ vrow, int64_t n, uint8_t* __restrict__ uval16) {
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x) {
    const int32_t u = coo_u[p];
    uval16[((int64_t)vrow[u] << 4) + (p - urow[u])] = uval[p];
  }
}

// user of every 16-code vector of the padded array (largest u with vrow[u] <= t)
__global__ void vec_row_kernel(const int32_t* __restrict__ vrow, int32_t n_users, int32_t n_vec, int32_t* __restrict__ vec_row) {
  for (int32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < n_vec; t += gridDim.x * blockDim.x) {
    int32_t lo = 0, hi = n_users;
    while (hi - lo > 1) {
      const int32_t mid = (lo + hi) >> 1;
      if (vrow[mid] <= t) lo = mid; else hi = mid;
    }
    vec_row[t] = lo;
  }
}

int grid_for(int64_t n, int block, int sm_count);

int32_t build_padded_codes(mrs_engine* e, mrs_ratings* R) {
  cudaStream_t st = e->stream;
  const int32_t NU = R->n_users;
  int32_t *cnt = nullptr, *vrow = nullptr;
  MRS_TRY(dev_alloc(&cnt, (size_t)NU + 1));
  MRS_TRY(dev_alloc(&vrow, (size_t)NU + 1));
  MRS_CUDA(cudaMemsetAsync(cnt, 0, sizeof(int32_t) * ((size_t)NU + 1), st));
  vec_count_kernel<<<(NU + 255) / 256, 256, 0, st>>>(R->urow, NU, cnt);
  size_t tmp = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, tmp, cnt, vrow, NU + 1, st);
  MRS_TRY(ensure_scratch(e, tmp));
  cub::DeviceScan::ExclusiveSum(e->scratch, tmp, cnt, vrow, NU + 1, st);
  int32_t n_vec = 0;
  MRS_CUDA(cudaMemcpyAsync(&n_vec, vrow + NU, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  MRS_CUDA(cudaStreamSynchronize(st));
  R->n_vec = n_vec;
  MRS_TRY(dev_alloc(&R->uval16, (size_t)n_vec * 16));
  MRS_TRY(dev_alloc(&R->vec_row, (size_t)n_vec));
  if (n_vec) {
    MRS_CUDA(cudaMemsetAsync(R->uval16, 0, (size_t)n_vec * 16, st));
    pad_fill_kernel<<<grid_for(R->n, 256, e->sm_count), 256, 0, st>>>((const uint8_t*)R->uval, R->coo_u, R->urow, vrow, R->n, R->uval16);
    vec_row_kernel<<<grid_for(n_vec, 256, e->sm_count), 256, 0, st>>>(vrow, NU, n_vec, R->vec_row);
    count_launch(4);
  }
  MRS_CUDA(cudaGetLastError());
  MRS_CUDA(cudaStreamSynchronize(st));
  dev_free(cnt); dev_free(vrow);
  return MRS_OK;
}

int32_t build_chunks(mrs_engine* e, const int32_t* seg_ptr, int32_t n_seg, int32_t chunk, mrs_chunks* out) {
  cudaStream_t st = e->stream;
  int32_t* counts = nullptr;
  MRS_TRY(dev_alloc(&counts, (size_t)n_seg + 1));
  MRS_TRY(dev_alloc(&out->seg_chunk_ptr, (size_t)n_seg + 1));
  MRS_CUDA(cudaMemsetAsync(counts, 0, sizeof(int32_t) * ((size_t)n_seg + 1), st));
  chunk_count_kernel<<<(n_seg + 255) / 256, 256, 0, st>>>(seg_ptr, n_seg, chunk, counts);
  count_launch();
  size_t tmp = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, tmp, counts, out->seg_chunk_ptr, n_seg + 1, st);
  MRS_TRY(ensure_scratch(e, tmp));
  cub::DeviceScan::ExclusiveSum(e->scratch, tmp, counts, out->seg_chunk_ptr, n_seg + 1, st);
  count_launch();
  int32_t total = 0;
  MRS_CUDA(cudaMemcpyAsync(&total, out->seg_chunk_ptr + n_seg, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  MRS_CUDA(cudaStreamSynchronize(st));
  out->n_chunks = total;
  MRS_TRY(dev_alloc(&out->chunk_seg, (size_t)total));
  MRS_TRY(dev_alloc(&out->chunk_begin, (size_t)total));
  chunk_fill_kernel<<<(n_seg + 255) / 256, 256, 0, st>>>(seg_ptr, out->seg_chunk_ptr, n_seg, chunk, out->chunk_seg, out->chunk_begin);
  count_launch();
  MRS_CUDA(cudaGetLastError());
  dev_free(counts);
  return MRS_OK;
}

struct sort_buffers {
  uint64_t *k_in = nullptr, *k_out = nullptr;
  int32_t *v_in = nullptr, *perm_u = nullptr;  // perm_u: raw position of every user-major entry (kept until the values arrive)
  int32_t* v2_in = nullptr;
};

// Everything that needs the ids only -- the sorts, both index structures, the duplicate check -- so that it runs while the
// ratings are still being copied.  The user-major order (items ascending inside a user) comes from two STABLE radix sorts with
// 32-bit keys, least significant key first: by item (sort_items: needs the items only, so it runs while the users are still
// on the PCIe link) and then by user (sort_ids: 3 passes at ml-25m shape).  One sort on a packed (user, item) key took 5 passes
// on 64-bit keys, all of them behind the arrival of both id arrays: 1.24 ms against 0.67 ms on the critical path of an
// end-to-end step (ncu launch lists, profiles/r02_summary.md).
int32_t sort_items(mrs_engine* e, mrs_ratings* R, const int32_t* d_i, sort_buffers* B) {
  cudaStream_t st = e->stream;
  const int64_t n = R->n;
  MRS_TRY(dev_alloc(&B->k_in, (size_t)n));
  MRS_TRY(dev_alloc(&B->k_out, (size_t)n));
  MRS_TRY(dev_alloc(&B->v_in, (size_t)n));
  MRS_TRY(dev_alloc(&B->perm_u, (size_t)n));
  MRS_TRY(dev_alloc(&B->v2_in, (size_t)n));
  if (n == 0) return MRS_OK;
  const int grid = grid_for(n, 256, e->sm_count);
  const int ibits = bits_for((uint32_t)(R->n_items - 1));
  uint32_t* k32_in = reinterpret_cast<uint32_t*>(B->k_in);   // the 64-bit buffers hold two 32-bit key arrays each
  uint32_t* k32_out = reinterpret_cast<uint32_t*>(B->k_out);
  item_keys_kernel<<<grid, 256, 0, st>>>(d_i, n, k32_in, B->v_in);  // key = item, payload = position in the raw input
  size_t tmp = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, tmp, k32_in, k32_out, B->v_in, B->v2_in, (int)n, 0, ibits, st);
  MRS_TRY(ensure_scratch(e, tmp));
  cub::DeviceRadixSort::SortPairs(e->scratch, tmp, k32_in, k32_out, B->v_in, B->v2_in, (int)n, 0, ibits, st);  // v2_in: raw positions, item order
  count_launch(4);
  MRS_CUDA(cudaGetLastError());
  return MRS_OK;
}

int32_t sort_ids(mrs_engine* e, mrs_ratings* R, const int32_t* d_u, const int32_t* d_i, int32_t* d_dup, sort_buffers* B) {
  cudaStream_t st = e->stream;
  const int64_t n = R->n;
  const int block = 256;
  const int grid = grid_for(n, block, e->sm_count);
  MRS_TRY(dev_alloc(&R->ucol, (size_t)n));
  MRS_TRY(dev_alloc(&R->coo_u, (size_t)n));
  MRS_TRY(dev_alloc(&R->irow, (size_t)n));
  MRS_TRY(dev_alloc(&R->csc_src, (size_t)n));
  MRS_TRY(dev_alloc(&R->urow, (size_t)R->n_users + 1));
  MRS_TRY(dev_alloc(&R->icolp, (size_t)R->n_items + 1));
  if (n == 0) {
    MRS_CUDA(cudaMemsetAsync(R->urow, 0, sizeof(int32_t) * ((size_t)R->n_users + 1), st));
    MRS_CUDA(cudaMemsetAsync(R->icolp, 0, sizeof(int32_t) * ((size_t)R->n_items + 1), st));
    return MRS_OK;
  }
  const int ubits = bits_for((uint32_t)(R->n_users - 1)), ibits = bits_for((uint32_t)(R->n_items - 1));
  uint32_t* k32_in = reinterpret_cast<uint32_t*>(B->k_in);
  uint32_t* k32_out = reinterpret_cast<uint32_t*>(B->k_out);
  uint32_t* ukey_in = k32_in + n;
  uint32_t* ukey_out = k32_out + n;
  // ---- user-major: the item-sorted entries (sort_items) sorted again, stably, by user; payload = position in the raw input
  user_keys_kernel<<<grid, block, 0, st>>>(d_u, B->v2_in, n, ukey_in);
  size_t tmp = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, tmp, ukey_in, ukey_out, B->v2_in, B->perm_u, (int)n, 0, ubits, st);
  MRS_TRY(ensure_scratch(e, tmp));
  cub::DeviceRadixSort::SortPairs(e->scratch, tmp, ukey_in, ukey_out, B->v2_in, B->perm_u, (int)n, 0, ubits, st);
  scatter_users_kernel<<<grid, block, 0, st>>>(ukey_out, B->perm_u, d_i, n, R->n_users, R->coo_u, R->ucol, R->urow, d_dup);
  // ---- item-major: a stable sort of the user-major entries on the item alone (the users of an item stay ascending);
  // payload = position in the user-major arrays
  item_keys_kernel<<<grid, block, 0, st>>>(R->ucol, n, k32_in, B->v2_in);
  cub::DeviceRadixSort::SortPairs(nullptr, tmp, k32_in, k32_out, B->v2_in, R->csc_src, (int)n, 0, ibits, st);
  MRS_TRY(ensure_scratch(e, tmp));
  cub::DeviceRadixSort::SortPairs(e->scratch, tmp, k32_in, k32_out, B->v2_in, R->csc_src, (int)n, 0, ibits, st);
  scatter_items_kernel<<<grid, block, 0, st>>>(k32_out, R->csc_src, R->coo_u, n, R->n_items, R->irow, R->icolp);
  count_launch(10);
  MRS_CUDA(cudaGetLastError());
  return MRS_OK;
}

template <typename VT>
int32_t finish_values(mrs_engine* e, mrs_ratings* R, const double* d_r, sort_buffers* B, const uint8_t* d_codes = nullptr) {
  cudaStream_t st = e->stream;
  const int64_t n = R->n;
  const int block = 256;
  const int grid = grid_for(n, block, e->sm_count);
  VT *uval = nullptr, *ival = nullptr;
  MRS_TRY(dev_alloc(&uval, (size_t)n + 32));  // +32: 128-bit loads may read past the last rating (masked)
  MRS_TRY(dev_alloc(&ival, (size_t)n + 32));
  R->uval = uval;
  R->ival = ival;
  if (n) {
    if (d_codes) gather_sorted_kernel<VT, false><<<grid, block, 0, st>>>(B->perm_u, n, d_codes, uval);   // compact form: codes as given
    else gather_sorted_kernel<VT, true><<<grid, block, 0, st>>>(B->perm_u, n, d_r, uval);
    gather_sorted_kernel<VT, false><<<grid, block, 0, st>>>(R->csc_src, n, uval, ival);
    count_launch(2);
    MRS_CUDA(cudaGetLastError());
  }
  if (sizeof(VT) == 1) MRS_TRY(build_padded_codes(e, R));
  if (sizeof(VT) != 1) {  // chunk lists feed the generic (fp64-valued) fit kernels only
    MRS_TRY(build_chunks(e, R->urow, R->n_users, kUserChunk, &R->uch));
    MRS_TRY(build_chunks(e, R->icolp, R->n_items, kItemChunk, &R->ich));
  }
  MRS_CUDA(cudaStreamSynchronize(st));
  return MRS_OK;
}

}  // namespace

void free_upload(mrs_upload* up) {
  if (!up) return;
  use_engine(up->eng);
  if (up->ev_values) cudaEventSynchronize(up->ev_values);  // the copies read host memory and write these buffers
  dev_free(up->d_u); dev_free(up->d_i); dev_free(up->d_r); dev_free(up->d_c);
  if (up->ev_items) cudaEventDestroy(up->ev_items);
  if (up->ev_ids) cudaEventDestroy(up->ev_ids);
  if (up->ev_values) cudaEventDestroy(up->ev_values);
  delete up;
}

// Enqueue the three host -> device copies on the engine's copy stream and return at once.  The ids go first: the first
// sort of the build needs only them and runs while the (twice as large) rating array is still on its way.
int32_t upload_begin(mrs_engine* e, const int32_t* users, const int32_t* items, const double* ratings, int64_t n, mrs_upload** out,
                     const uint8_t* codes = nullptr) {
  MRS_REQUIRE(e && out, MRS_ERR_INVALID, "mrs_upload_begin: NULL engine or output");
  MRS_REQUIRE(n >= 0 && n < (int64_t)0x7fffffff, MRS_ERR_UNSUPPORTED, "mrs_ratings_from_coo: n=%lld outside [0, 2^31)", (long long)n);
  MRS_REQUIRE(n == 0 || (users && items && (ratings || codes)), MRS_ERR_INVALID, "mrs_ratings_from_coo: NULL input array");
  use_engine(e);
  mrs_upload* up = new mrs_upload();
  up->eng = e;
  up->n = n;
  int32_t rc = dev_alloc(&up->d_u, (size_t)n);
  if (rc == MRS_OK) rc = dev_alloc(&up->d_i, (size_t)n);
  if (rc == MRS_OK && !codes) rc = dev_alloc(&up->d_r, (size_t)n);
  if (rc == MRS_OK && codes) rc = dev_alloc(&up->d_c, (size_t)n + 16);
  cudaError_t ce = cudaSuccess;
  if (rc == MRS_OK) ce = cudaEventCreateWithFlags(&up->ev_ids, cudaEventDisableTiming);
  if (rc == MRS_OK && ce == cudaSuccess) ce = cudaEventCreateWithFlags(&up->ev_items, cudaEventDisableTiming);
  if (rc == MRS_OK && ce == cudaSuccess) ce = cudaEventCreateWithFlags(&up->ev_values, cudaEventDisableTiming);
  // the staging buffers may be recycled blocks: stay behind whatever the compute stream still has queued on them
  if (rc == MRS_OK && ce == cudaSuccess) ce = cudaEventRecord(e->ev_order, e->stream);
  if (rc == MRS_OK && ce == cudaSuccess) ce = cudaStreamWaitEvent(e->copy_stream, e->ev_order, 0);
  // the items travel first: they are the least significant sort key, and their sort runs while the users follow
  if (rc == MRS_OK && ce == cudaSuccess && n) ce = cudaMemcpyAsync(up->d_i, items, sizeof(int32_t) * n, cudaMemcpyHostToDevice, e->copy_stream);
  if (rc == MRS_OK && ce == cudaSuccess) ce = cudaEventRecord(up->ev_items, e->copy_stream);
  if (rc == MRS_OK && ce == cudaSuccess && n) ce = cudaMemcpyAsync(up->d_u, users, sizeof(int32_t) * n, cudaMemcpyHostToDevice, e->copy_stream);
  if (rc == MRS_OK && ce == cudaSuccess) ce = cudaEventRecord(up->ev_ids, e->copy_stream);
  if (rc == MRS_OK && ce == cudaSuccess && n && !codes) ce = cudaMemcpyAsync(up->d_r, ratings, sizeof(double) * n, cudaMemcpyHostToDevice, e->copy_stream);
  if (rc == MRS_OK && ce == cudaSuccess && n && codes) ce = cudaMemcpyAsync(up->d_c, codes, (size_t)n, cudaMemcpyHostToDevice, e->copy_stream);
  if (rc == MRS_OK && ce == cudaSuccess) ce = cudaEventRecord(up->ev_values, e->copy_stream);
  if (rc == MRS_OK && ce != cudaSuccess) { set_error("mrs_upload_begin: %s", cudaGetErrorString(ce)); rc = MRS_ERR_CUDA; }
  if (rc != MRS_OK) { free_upload(up); return rc; }
  *out = up;
  return MRS_OK;
}

// Build the device-resident rating set from a staged upload (consumes it).
int32_t ratings_from_upload(mrs_upload* up, int32_t n_users_dim, int32_t n_items_dim, mrs_ratings** out) {
  MRS_REQUIRE(up && out, MRS_ERR_INVALID, "mrs_ratings_from_upload: NULL argument");
  mrs_engine* e = up->eng;
  use_engine(e);
  cudaStream_t st = e->stream;
  const int64_t n = up->n;
  int32_t* d_stats = nullptr;
  int32_t s = dev_alloc(&d_stats, 8);
  if (s != MRS_OK) { free_upload(up); return s; }
  int32_t h_stats[5] = {-1, -1, 0x7fffffff, 0, 0};
  mrs_ratings* R = nullptr;
  sort_buffers B;
  auto fail = [&](int32_t code) {
    dev_free(B.k_in); dev_free(B.k_out); dev_free(B.v_in); dev_free(B.perm_u); dev_free(B.v2_in);
    dev_free(d_stats);
    free_upload(up);
    if (R) mrs_ratings_destroy(R);
    return code;
  };
  cudaError_t ce = cudaMemcpyAsync(d_stats, h_stats, sizeof(h_stats), cudaMemcpyHostToDevice, st);
  // ---- the items (they arrive first): id range, then their sort -- enqueued before we wait for the users
  if (ce == cudaSuccess) ce = cudaStreamWaitEvent(st, up->ev_items ? up->ev_items : up->ev_ids, 0);
  if (ce == cudaSuccess && n) {
    scan_id_kernel<<<grid_for(n, 256, e->sm_count), 256, 0, st>>>(up->d_i, n, d_stats + 1, d_stats + 2);
    count_launch();
  }
  if (ce == cudaSuccess) ce = cudaMemcpyAsync(h_stats, d_stats, 3 * sizeof(int32_t), cudaMemcpyDeviceToHost, st);
  if (ce == cudaSuccess) ce = cudaStreamSynchronize(st);
  if (ce != cudaSuccess) { set_error("mrs_ratings_from_coo: %s", cudaGetErrorString(ce)); return fail(MRS_ERR_CUDA); }
  if (n && h_stats[2] < 0) {
    set_error("mrs_ratings_from_coo: negative user or item id (%d)", h_stats[2]);
    return fail(MRS_ERR_INVALID);
  }
  R = new mrs_ratings();
  R->eng = e;
  R->n = n;
  R->n_items = std::max(n_items_dim, h_stats[1] + 1);
  if (R->n_items < 1) R->n_items = 1;
  R->n_users = 1;
  s = sort_items(e, R, up->d_i, &B);
  if (s != MRS_OK) return fail(s);
  // ---- the users
  ce = cudaStreamWaitEvent(st, up->ev_ids, 0);
  if (ce == cudaSuccess && n) {
    scan_id_kernel<<<grid_for(n, 256, e->sm_count), 256, 0, st>>>(up->d_u, n, d_stats, d_stats + 2);
    count_launch();
  }
  if (ce == cudaSuccess) ce = cudaMemcpyAsync(h_stats, d_stats, 3 * sizeof(int32_t), cudaMemcpyDeviceToHost, st);
  if (ce == cudaSuccess) ce = cudaStreamSynchronize(st);
  if (ce != cudaSuccess) { set_error("mrs_ratings_from_coo: %s", cudaGetErrorString(ce)); return fail(MRS_ERR_CUDA); }
  if (n && h_stats[2] < 0) {
    set_error("mrs_ratings_from_coo: negative user or item id (%d)", h_stats[2]);
    return fail(MRS_ERR_INVALID);
  }
  R->n_users = std::max(n_users_dim, h_stats[0] + 1);
  if (R->n_users < 1) R->n_users = 1;
  int32_t* d_dup = d_stats + 4;
  s = sort_ids(e, R, up->d_u, up->d_i, d_dup, &B);  // enqueued before we wait for the ratings
  if (s != MRS_OK) return fail(s);
  ce = cudaStreamWaitEvent(st, up->ev_values, 0);
  if (ce == cudaSuccess && n) {
    if (up->d_c) scan_codes_kernel<<<grid_for(n, 256, e->sm_count), 256, 0, st>>>(up->d_c, n, d_stats);
    else scan_values_kernel<<<grid_for(n, 256, e->sm_count), 256, 0, st>>>(up->d_r, n, d_stats);
    count_launch();
  }
  if (ce == cudaSuccess) ce = cudaMemcpyAsync(h_stats + 3, d_stats + 3, sizeof(int32_t), cudaMemcpyDeviceToHost, st);
  if (ce == cudaSuccess) ce = cudaStreamSynchronize(st);
  if (ce != cudaSuccess) { set_error("mrs_ratings_from_coo: %s", cudaGetErrorString(ce)); return fail(MRS_ERR_CUDA); }
  if (up->d_c && h_stats[3] != 0) {
    set_error("mrs_upload_begin_codes: %d rating code(s) are 255 (ratings of 127.5 are not representable in the compact form)", h_stats[3]);
    return fail(MRS_ERR_INVALID);
  }
  R->value_kind = (h_stats[3] == 0) ? kValueCode : kValueF64;
  s = (R->value_kind == kValueCode) ? finish_values<uint8_t>(e, R, up->d_r, &B, up->d_c) : finish_values<double>(e, R, up->d_r, &B);
  int32_t dup = 0;
  if (s == MRS_OK) {
    ce = cudaMemcpy(&dup, d_dup, sizeof(int32_t), cudaMemcpyDeviceToHost);
    if (ce != cudaSuccess) { set_error("cudaMemcpy failed: %s", cudaGetErrorString(ce)); s = MRS_ERR_CUDA; }
  }
  if (s == MRS_OK && dup != 0) {
    // each duplicate pair is seen once in the CSR pass and once in the CSC pass
    set_error("mrs_ratings_from_coo: %d duplicate (user,item) pair(s); the reference keeps the last one (P:168), this engine rejects them", dup / 2);
    s = MRS_ERR_DUPLICATE;
  }
  if (s != MRS_OK) return fail(s);
  dev_free(B.k_in); dev_free(B.k_out); dev_free(B.v_in); dev_free(B.perm_u); dev_free(B.v2_in);
  dev_free(d_stats);
  free_upload(up);
  *out = R;
  return MRS_OK;
}

int32_t build_ratings(mrs_engine* e, const int32_t* users, const int32_t* items, const double* ratings, int64_t n,
                      int32_t n_users_dim, int32_t n_items_dim, mrs_ratings** out) {
  MRS_REQUIRE(e && out, MRS_ERR_INVALID, "mrs_ratings_from_coo: NULL engine or output");
  mrs_upload* up = nullptr;
  MRS_TRY(upload_begin(e, users, items, ratings, n, &up));
  return ratings_from_upload(up, n_users_dim, n_items_dim, out);
}

}  // namespace mrs

// ------------------------------------------------------------------ C ABI
using namespace mrs;

extern "C" int32_t mrs_upload_begin(mrs_engine* e, const int32_t* users, const int32_t* items, const double* ratings, int64_t n,
                                    mrs_upload** out) {
  return upload_begin(e, users, items, ratings, n, out);
}

extern "C" int32_t mrs_upload_begin_codes(mrs_engine* e, const int32_t* users, const int32_t* items, const uint8_t* codes, int64_t n,
                                          mrs_upload** out) {
  MRS_REQUIRE(n == 0 || codes, MRS_ERR_INVALID, "mrs_upload_begin_codes: NULL code array");
  return upload_begin(e, users, items, nullptr, n, out, codes ? codes : reinterpret_cast<const uint8_t*>(""));
}

extern "C" int32_t mrs_ratings_from_coo_codes(mrs_engine* e, const int32_t* users, const int32_t* items, const uint8_t* codes, int64_t n,
                                              int32_t n_users_dim, int32_t n_items_dim, mrs_ratings** out) {
  MRS_REQUIRE(e && out, MRS_ERR_INVALID, "mrs_ratings_from_coo_codes: NULL engine or output");
  mrs_upload* up = nullptr;
  MRS_TRY(mrs_upload_begin_codes(e, users, items, codes, n, &up));
  return ratings_from_upload(up, n_users_dim, n_items_dim, out);
}

extern "C" int32_t mrs_ratings_from_upload(mrs_upload* up, int32_t n_users_dim, int32_t n_items_dim, mrs_ratings** out) {
  return ratings_from_upload(up, n_users_dim, n_items_dim, out);
}

extern "C" void mrs_upload_destroy(mrs_upload* up) { free_upload(up); }

extern "C" int32_t mrs_ratings_from_coo(mrs_engine* e, const int32_t* users, const int32_t* items, const double* ratings,
                                        int64_t n, int32_t n_users_dim, int32_t n_items_dim, mrs_ratings** out) {
  return build_ratings(e, users, items, ratings, n, n_users_dim, n_items_dim, out);
}

// ---------------------------------------------------------------------------------------------------------------------
// Text loader: `load` (P:35-49) = textFile -> split(sep) -> trim -> keep the row iff column 0 is an Int.
// The text is parsed ON THE DEVICE (one thread per line); the host parser below states the same rules and is used only
// when a rating is written in a form the device parser does not take (exponent, hex, NaN, more than 15 digits).

// Java semantics of Integer.parseInt after trim (P:27-33 toInt): optional sign, then decimal digits only.
__host__ __device__ static bool parse_java_int(const char* b, const char* e, int32_t* out) {
  while (b < e && (unsigned char)*b <= ' ') ++b;  // String.trim strips chars <= U+0020
  while (e > b && (unsigned char)e[-1] <= ' ') --e;
  if (b == e) return false;
  bool neg = false;
  if (*b == '-' || *b == '+') { neg = (*b == '-'); ++b; }
  if (b == e) return false;
  int64_t v = 0;
  for (; b < e; ++b) {
    if (*b < '0' || *b > '9') return false;
    v = v * 10 + (*b - '0');
    if (v > (int64_t)0x80000000LL) return false;
  }
  v = neg ? -v : v;
  if (v > 0x7fffffffLL || v < -(int64_t)0x80000000LL) return false;
  *out = (int32_t)v;
  return true;
}

// first three columns of a line split on the literal separator; returns the number of columns found (1..3, 4 = more)
__host__ __device__ static int split3(const char* p, const char* le, const char* sep, int seplen, const char** c, const char** ce) {
  c[0] = p; c[1] = c[2] = nullptr;
  ce[0] = ce[1] = ce[2] = le;
  int nc = 1;
  const char* q = p;
  while (nc < 4) {
    const char* hit = nullptr;
    for (const char* t = q; t + seplen <= le; ++t) {
      bool eq = true;
      for (int k = 0; k < seplen; ++k)
        if (t[k] != sep[k]) { eq = false; break; }
      if (eq) { hit = t; break; }
    }
    if (!hit) break;
    ce[nc - 1] = hit;
    if (nc < 3) c[nc] = hit + seplen;
    q = hit + seplen;
    ++nc;
  }
  return nc;
}

namespace mrs {
namespace {

constexpr int kMaxSep = 8;
struct sep_arg { char s[kMaxSep]; int len; };

struct is_newline {
  const char* text;
  __device__ bool operator()(int32_t p) const { return text[p] == '\n'; }
};

// [sign] digits [. digits], at most 15 significant digits: mantissa and power of ten are exact doubles, one correctly
// rounded division gives the correctly rounded value (what Double.parseDouble returns).  0 = parsed, 2 = another form.
__device__ int parse_plain_decimal(const char* b, const char* e, double* out) {
  while (b < e && (unsigned char)*b <= ' ') ++b;
  while (e > b && (unsigned char)e[-1] <= ' ') --e;
  if (b == e) return 2;
  bool neg = false;
  if (*b == '-' || *b == '+') { neg = (*b == '-'); ++b; }
  unsigned long long mant = 0;
  int nd = 0, frac = 0;
  bool digit = false, dot = false;
  for (; b < e; ++b) {
    const char ch = *b;
    if (ch >= '0' && ch <= '9') {
      digit = true;
      if (nd > 15) return 2;
      mant = mant * 10ull + (unsigned)(ch - '0');
      if (mant) ++nd;
      if (dot) ++frac;
    } else if (ch == '.' && !dot) {
      dot = true;
    } else {
      return 2;
    }
  }
  if (!digit || nd > 15 || frac > 22) return 2;
  double p10 = 1.0;
  for (int k = 0; k < frac; ++k) p10 *= 10.0;  // exact up to 1e22
  const double v = __ddiv_rn((double)mant, p10);
  *out = neg ? -v : v;
  return 0;
}

// status[0] = first line (0-based) with a malformed row, status[1] = number of ratings in another number form
__global__ void parse_lines_kernel(const char* __restrict__ text, const int32_t* __restrict__ nl_pos, int32_t n_lines, sep_arg sep,
                                   int32_t* __restrict__ us, int32_t* __restrict__ is, double* __restrict__ rs,
                                   unsigned char* __restrict__ keep, int32_t* __restrict__ status) {
  const int32_t l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= n_lines) return;
  const char* p = text + (l ? nl_pos[l - 1] + 1 : 0);
  const char* le = text + nl_pos[l];
  if (le > p && le[-1] == '\r') --le;  // textFile strips \r\n as well as \n
  const char *c[3], *ce[3];
  const int nc = split3(p, le, sep.s, sep.len, c, ce);
  int32_t uu = 0, ii = 0;
  double rr = 0.0;
  unsigned char k = 0;
  if (parse_java_int(c[0], ce[0], &uu)) {  // P:40: rows whose column 0 is not an Int are dropped (header skip)
    if (nc < 3 || !parse_java_int(c[1], ce[1], &ii)) {
      atomicMin(status, l);                // the reference throws at P:41
    } else if (parse_plain_decimal(c[2], ce[2], &rr) != 0) {
      atomicAdd(status + 1, 1);
    } else {
      k = 1;
    }
  }
  us[l] = uu; is[l] = ii; rs[l] = rr; keep[l] = k;
}

// the same rules on the host (strtod takes every number form); returns the 1-based line of the first malformed row or 0
int64_t parse_text_host(const char* text, size_t len, const char* sep, std::vector<int32_t>& us, std::vector<int32_t>& is,
                        std::vector<double>& rs, bool* bad_number) {
  const int seplen = (int)strlen(sep);
  const char* p = text;
  const char* end = text + len;
  int64_t lineno = 0;
  *bad_number = false;
  while (p < end) {
    const char* nl = (const char*)memchr(p, '\n', (size_t)(end - p));
    if (!nl) nl = end;
    const char* le = nl;
    if (le > p && le[-1] == '\r') --le;
    ++lineno;
    const char *c[3], *ce[3];
    const int nc = split3(p, le, sep, seplen, c, ce);
    int32_t uu = 0, ii = 0;
    if (parse_java_int(c[0], ce[0], &uu)) {
      if (nc < 3 || !parse_java_int(c[1], ce[1], &ii)) return lineno;
      std::string tok(c[2], ce[2]);
      char* endp = nullptr;
      double rr = strtod(tok.c_str(), &endp);
      while (endp && *endp && (unsigned char)*endp <= ' ') ++endp;
      if (endp == tok.c_str() || (endp && *endp)) { *bad_number = true; return lineno; }
      us.push_back(uu); is.push_back(ii); rs.push_back(rr);
    }
    p = nl + 1;
  }
  return 0;
}

// text (host memory, not retained) -> rating set
int32_t ratings_from_text(mrs_engine* e, const char* text, int64_t nbytes, const char* sep, const char* what, mrs_ratings** out) {
  MRS_REQUIRE(e && sep && out && sep[0] && (text || nbytes == 0), MRS_ERR_INVALID, "mrs_ratings_from_text: NULL/empty argument");
  const size_t seplen = strlen(sep);
  MRS_REQUIRE(seplen <= (size_t)kMaxSep, MRS_ERR_UNSUPPORTED, "mrs_ratings_from_text: separator longer than %d bytes", kMaxSep);
  MRS_REQUIRE(nbytes >= 0 && nbytes < (int64_t)0x7ffffff0, MRS_ERR_UNSUPPORTED, "mrs_ratings_from_text: %lld bytes outside [0, 2^31)", (long long)nbytes);
  use_engine(e);
  cudaStream_t st = e->stream;
  const int32_t n = (int32_t)nbytes + 1;  // + a final '\n' so that the last line is terminated
  char* d_text = nullptr;
  int32_t *d_nl = nullptr, *d_num = nullptr, *d_status = nullptr, *l_u = nullptr, *l_i = nullptr;
  double* l_r = nullptr;
  unsigned char* d_keep = nullptr;
  mrs_upload* up = nullptr;
  auto cleanup = [&]() {
    dev_free(d_text); dev_free(d_nl); dev_free(d_num); dev_free(d_status); dev_free(l_u); dev_free(l_i); dev_free(l_r); dev_free(d_keep);
  };
#define MRS_TEXT_TRY(expr)                                                      \
  do {                                                                          \
    int32_t s__ = (expr);                                                       \
    if (s__ != MRS_OK) { cleanup(); if (up) free_upload(up); return s__; }      \
  } while (0)
#define MRS_TEXT_CUDA(expr)                                                                             \
  do {                                                                                                  \
    cudaError_t c__ = (expr);                                                                           \
    if (c__ != cudaSuccess) {                                                                           \
      set_error("mrs_ratings_from_text: %s", cudaGetErrorString(c__));                                  \
      cleanup(); if (up) free_upload(up);                                                               \
      return MRS_ERR_CUDA;                                                                              \
    }                                                                                                   \
  } while (0)
  MRS_TEXT_TRY(dev_alloc(&d_text, (size_t)n));
  MRS_TEXT_TRY(dev_alloc(&d_num, 4));
  MRS_TEXT_TRY(dev_alloc(&d_status, 4));
  if (nbytes) MRS_TEXT_CUDA(cudaMemcpyAsync(d_text, text, (size_t)nbytes, cudaMemcpyHostToDevice, st));
  MRS_TEXT_CUDA(cudaMemsetAsync(d_text + nbytes, '\n', 1, st));
  // ---- line ends
  cub::CountingInputIterator<int32_t> pos(0);
  size_t tmp = 0;
  MRS_TEXT_TRY(dev_alloc(&d_nl, (size_t)n));  // upper bound: every byte a newline
  cub::DeviceSelect::If(nullptr, tmp, pos, d_nl, d_num, n, is_newline{d_text}, st);
  MRS_TEXT_TRY(ensure_scratch(e, tmp));
  cub::DeviceSelect::If(e->scratch, tmp, pos, d_nl, d_num, n, is_newline{d_text}, st);
  count_launch(2);
  int32_t n_lines = 0;
  MRS_TEXT_CUDA(cudaMemcpyAsync(&n_lines, d_num, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  MRS_TEXT_CUDA(cudaStreamSynchronize(st));
  // ---- one thread per line
  MRS_TEXT_TRY(dev_alloc(&l_u, (size_t)n_lines));
  MRS_TEXT_TRY(dev_alloc(&l_i, (size_t)n_lines));
  MRS_TEXT_TRY(dev_alloc(&l_r, (size_t)n_lines));
  MRS_TEXT_TRY(dev_alloc(&d_keep, (size_t)n_lines));
  int32_t h_status[2] = {0x7fffffff, 0};
  MRS_TEXT_CUDA(cudaMemcpyAsync(d_status, h_status, sizeof(h_status), cudaMemcpyHostToDevice, st));
  sep_arg sa;
  memset(&sa, 0, sizeof(sa));
  memcpy(sa.s, sep, seplen);
  sa.len = (int)seplen;
  parse_lines_kernel<<<(n_lines + 255) / 256, 256, 0, st>>>(d_text, d_nl, n_lines, sa, l_u, l_i, l_r, d_keep, d_status);
  count_launch();
  MRS_TEXT_CUDA(cudaMemcpyAsync(h_status, d_status, sizeof(h_status), cudaMemcpyDeviceToHost, st));
  MRS_TEXT_CUDA(cudaStreamSynchronize(st));
  if (h_status[0] != 0x7fffffff) {
    cleanup();
    set_error("mrs_ratings_from_file: %s:%lld: malformed row (the reference throws at predictions.scala:41)", what, (long long)h_status[0] + 1);
    return MRS_ERR_IO;
  }
  if (h_status[1] != 0) {  // some rating is not a plain decimal: the host parser (strtod) decides
    cleanup();
    std::vector<int32_t> us, is;
    std::vector<double> rs;
    bool bad_number = false;
    const int64_t bad = parse_text_host(text, (size_t)nbytes, sep, us, is, rs, &bad_number);
    if (bad) {
      set_error(bad_number ? "mrs_ratings_from_file: %s:%lld: rating is not a number"
                           : "mrs_ratings_from_file: %s:%lld: malformed row (the reference throws at predictions.scala:41)", what, (long long)bad);
      return MRS_ERR_IO;
    }
    return build_ratings(e, us.data(), is.data(), rs.data(), (int64_t)us.size(), 0, 0, out);
  }
  // ---- kept rows, compacted straight into the staging buffers of the build
  up = new mrs_upload();
  up->eng = e;
  MRS_TEXT_TRY(dev_alloc(&up->d_u, (size_t)n_lines));
  MRS_TEXT_TRY(dev_alloc(&up->d_i, (size_t)n_lines));
  MRS_TEXT_TRY(dev_alloc(&up->d_r, (size_t)n_lines));
  MRS_TEXT_CUDA(cudaEventCreateWithFlags(&up->ev_ids, cudaEventDisableTiming));
  MRS_TEXT_CUDA(cudaEventCreateWithFlags(&up->ev_values, cudaEventDisableTiming));
  cub::DeviceSelect::Flagged(nullptr, tmp, l_r, d_keep, up->d_r, d_num, n_lines, st);
  MRS_TEXT_TRY(ensure_scratch(e, tmp));
  cub::DeviceSelect::Flagged(e->scratch, tmp, l_u, d_keep, up->d_u, d_num, n_lines, st);
  cub::DeviceSelect::Flagged(e->scratch, tmp, l_i, d_keep, up->d_i, d_num, n_lines, st);
  cub::DeviceSelect::Flagged(e->scratch, tmp, l_r, d_keep, up->d_r, d_num, n_lines, st);
  count_launch(6);
  int32_t kept = 0;
  MRS_TEXT_CUDA(cudaMemcpyAsync(&kept, d_num, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  MRS_TEXT_CUDA(cudaEventRecord(up->ev_ids, st));
  MRS_TEXT_CUDA(cudaEventRecord(up->ev_values, st));
  MRS_TEXT_CUDA(cudaStreamSynchronize(st));
  up->n = kept;
  cleanup();
#undef MRS_TEXT_TRY
#undef MRS_TEXT_CUDA
  return ratings_from_upload(up, 0, 0, out);
}

}  // namespace
}  // namespace mrs

extern "C" int32_t mrs_ratings_from_text(mrs_engine* e, const char* text, int64_t nbytes, const char* sep, mrs_ratings** out) {
  return ratings_from_text(e, text, nbytes, sep, "<text>", out);
}

extern "C" int32_t mrs_ratings_from_file(mrs_engine* e, const char* path, const char* sep, mrs_ratings** out) {
  MRS_REQUIRE(e && path && sep && out && sep[0], MRS_ERR_INVALID, "mrs_ratings_from_file: NULL/empty argument");
  FILE* f = fopen(path, "rb");
  MRS_REQUIRE(f != nullptr, MRS_ERR_IO, "mrs_ratings_from_file: cannot open %s", path);
  std::vector<char> buf;
  {
    fseek(f, 0, SEEK_END);
    long sz = ftell(f);
    fseek(f, 0, SEEK_SET);
    buf.resize((size_t)std::max(0L, sz));
    size_t got = buf.empty() ? 0 : fread(buf.data(), 1, buf.size(), f);
    buf.resize(got);
    fclose(f);
  }
  return ratings_from_text(e, buf.data(), (int64_t)buf.size(), sep, path, out);
}

extern "C" int32_t mrs_ratings_info(const mrs_ratings* r, int64_t* n, int32_t* n_users_dim, int32_t* n_items_dim, int32_t* value_kind) {
  MRS_REQUIRE(r, MRS_ERR_INVALID, "mrs_ratings_info: NULL handle");
  if (n) *n = r->n;
  if (n_users_dim) *n_users_dim = r->n_users;
  if (n_items_dim) *n_items_dim = r->n_items;
  if (value_kind) *value_kind = r->value_kind;
  return MRS_OK;
}

extern "C" int32_t mrs_ratings_bytes(const mrs_ratings* r, int64_t* b) {
  MRS_REQUIRE(r && b, MRS_ERR_INVALID, "mrs_ratings_bytes: NULL argument");
  const int64_t vs = (int64_t)r->value_size();
  b[0] = r->n * vs;          // user-major pass reads the values only (the user is implicit in the row pointer)
  b[1] = r->n * (4 + vs);    // item-major pass: user id + value
  b[2] = r->n * (8 + vs);    // sorted COO pass: user id + item id + value
  if (r->uval16) b[0] = (int64_t)r->n_vec * 20;  // padded codes (16 per vector) + one user id per vector
  if (r->tl.built) b[1] = r->n * 4;              // tiled item-major layout: one packed 32-bit word per rating (padding not counted)
  if (r->ml.built) b[2] = r->n * 8;              // item-tiled test layout: one packed 8-byte word per rating (padding not counted)
  return MRS_OK;
}

extern "C" int32_t mrs_ratings_layout_info(const mrs_ratings* r, int64_t* o) {
  MRS_REQUIRE(r && o, MRS_ERR_INVALID, "mrs_ratings_layout_info: NULL argument");
  o[0] = r->tl.n_tiles; o[1] = r->tl.n_units; o[2] = r->tl.n_slices; o[3] = r->tl.n_slots;
  o[4] = r->ml.n_tiles; o[5] = r->ml.n_rows; o[6] = r->ml.n_rows * 32; o[7] = r->n_vec;
  return MRS_OK;
}

static void free_chunks(mrs_chunks* c) {
  dev_free(c->chunk_seg); dev_free(c->chunk_begin); dev_free(c->seg_chunk_ptr);
}

extern "C" void mrs_ratings_destroy(mrs_ratings* r) {
  if (!r) return;
  if (r->eng) use_engine(r->eng);
  dev_free(r->urow); dev_free(r->ucol); dev_free(r->uval); dev_free(r->coo_u); dev_free(r->vec_row); dev_free(r->uval16);
  dev_free(r->icolp); dev_free(r->irow); dev_free(r->ival); dev_free(r->csc_src);
  free_chunks(&r->uch); free_chunks(&r->ich);
  free_sim_layout(r);
  free_tiled_layout(r);
  free_mae_layout(r);
  delete r;
}
