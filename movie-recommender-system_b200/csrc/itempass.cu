// itempass.cu -- the item-deviation pass of the fit (P:155-186 / P:316-343): per item, the sum over its raters of
//     (r - avg_u) / scale(r, avg_u)                                                            (P:57-61, P:167)
// on a layout made for B200.  What round 1 taught (profiles/r01_summary.md): the pass is not byte bound but bound by
// (a) instruction issue -- a per-rating fp64 division costs ~25 instructions -- and (b) the load/store unit: every
// random 8-byte shared-memory gather costs ~6 LSU cycles per warp and every per-unit atomic ~1.3 cycles per lane.
// Hence two structures, split by ITEM POPULARITY when the layout is built:
//
//   * POPULAR items (>= ~6 ratings per 2,048 users; two thirds of the ratings at ml-25m shape).  A user has at most
//     kMaxCodes distinct deviations, so a CTA stages the table dev[user][code] of a tile of kPopTileUsers users in
//     shared memory (2,048 x 10 x 8 B = 160 KB, built from K1's per-user code sums with two reciprocals per user) and
//     the deviation of a rating is ONE shared-memory load: an entry is the 16-bit table index user*n_codes + code.
//     Small tiles are affordable here because popular items still have long (tile, item) runs.
//   * RARE items.  Small tiles would cut their few ratings into one-entry runs and every run costs an atomic, so
//     they use tiles of kRareTileUsers users with 8 bytes per user in shared memory ((code sum, count), 128 KB) and
//     compute the deviation per rating in exact integer form N/D (one reciprocal seed + 3 DFMA).
//
// Both are user-tiled sliced-ELL: inside a tile the entries are item-major; every (tile, item) run is cut into units
// of <= kUnitLen entries, units are sorted by length and packed 32 to a slice; lane l of a warp owns unit l of the
// slice and walks it sequentially (rows of 128 bytes: one 32-bit word per lane = one rare entry or two popular
// entries), so there is no cross-lane reduction and a fixed summation order.  A warp streams a contiguous run of rows
// through a private ring of TMA bulk copies (cp.async.bulk + mbarrier).  Unit sums are fp64; they are combined across
// units with integer atomics on a 2^-40 grid, which is exact, so the item sums do not depend on the order of arrival
// and the pass is bit-reproducible.  One launch covers both parts: CTAs are dealt out to the tiles of both parts in
// proportion to their cost (static partition, laid down with the layout).
#include <cub/cub.cuh>

#include <algorithm>
#include <climits>
#include <cstdlib>
#include <vector>

#include "common.cuh"
#include "tma.cuh"

namespace mrs {
namespace {

struct MaxOp {
  __device__ __forceinline__ int32_t operator()(int32_t a, int32_t b) const { return a > b ? a : b; }
};

int grid_for(int64_t n, int block, int sm_count) {
  return (int)std::max<int64_t>(1, std::min<int64_t>((n + block - 1) / block, (int64_t)sm_count * 16));
}

// ------------------------------------------------------------------------------------------------ layout kernels
// item of every CSC position (binary search in the column pointer)
__global__ void item_of_kernel(const int32_t* __restrict__ icolp, int32_t n_items, int64_t n, int32_t* __restrict__ item_of) {
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x) {
    int32_t lo = 0, hi = n_items;  // largest i with icolp[i] <= p
    while (hi - lo > 1) {
      const int32_t mid = (lo + hi) >> 1;
      if (icolp[mid] <= (int32_t)p) lo = mid; else hi = mid;
    }
    item_of[p] = lo;
  }
}

// stats[0] = min code, stats[1] = max code
__global__ void code_range_kernel(const uint8_t* __restrict__ val, int64_t n, int32_t* __restrict__ stats) {
  int32_t lo = 255, hi = 0;
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x) {
    const int32_t c = val[p];
    lo = min(lo, c); hi = max(hi, c);
  }
  lo = __reduce_min_sync(0xffffffffu, lo);
  hi = __reduce_max_sync(0xffffffffu, hi);
  if ((threadIdx.x & 31) == 0) { atomicMin(&stats[0], lo); atomicMax(&stats[1], hi); }
}

// flag[i] = 1 for popular items; count[0] += ratings of popular items
__global__ void pop_flag_kernel(const int32_t* __restrict__ icolp, int32_t n_items, int32_t threshold, uint8_t* __restrict__ flag,
                                unsigned long long* __restrict__ count) {
  const int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  int32_t c = 0;
  if (i < n_items) {
    c = icolp[i + 1] - icolp[i];
    const bool pop = c >= threshold;
    flag[i] = pop ? 1 : 0;
    if (!pop) c = 0;
  }
  c = __reduce_add_sync(0xffffffffu, c);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(count, (unsigned long long)c);
}

// sort key of every CSC position for one part: its user tile, or n_tiles (sorts behind everything) for entries of the
// other part; the radix sort is stable, so (item, user) order survives inside a tile
__global__ void part_keys_kernel(const int32_t* __restrict__ irow, const int32_t* __restrict__ item_of, const uint8_t* __restrict__ flag,
                                 int32_t want, int32_t tile_users, int32_t n_tiles, int64_t n, uint16_t* __restrict__ key,
                                 int32_t* __restrict__ pos) {
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x) {
    key[p] = (flag[item_of[p]] == want) ? (uint16_t)(irow[p] / tile_users) : (uint16_t)n_tiles;
    pos[p] = (int32_t)p;
  }
}

// q = position in (tile, item, user) order.  head_pos[q] = q at the first entry of a (tile,item) run, else 0
__global__ void seg_head_kernel(const int32_t* __restrict__ perm, const int32_t* __restrict__ item_of, const int32_t* __restrict__ irow,
                                int32_t tile_users, int64_t n, int32_t* __restrict__ head_pos) {
  for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < n; q += (int64_t)gridDim.x * blockDim.x) {
    bool head = (q == 0);
    if (!head) {
      const int32_t p = perm[q], pp = perm[q - 1];
      head = (item_of[p] != item_of[pp]) || (irow[p] / tile_users != irow[pp] / tile_users);
    }
    head_pos[q] = head ? (int32_t)q : 0;
  }
}

__global__ void unit_flag_kernel(const int32_t* __restrict__ seg_start, int64_t n, int32_t* __restrict__ flag) {
  for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < n; q += (int64_t)gridDim.x * blockDim.x)
    flag[q] = ((q - seg_start[q]) % kUnitLen == 0) ? 1 : 0;
}

__global__ void unit_scatter_kernel(const int32_t* __restrict__ flag, const int32_t* __restrict__ uid, const int32_t* __restrict__ perm,
                                    const int32_t* __restrict__ item_of, const int32_t* __restrict__ irow, int32_t tile_users, int64_t n,
                                    int32_t* __restrict__ unit_begin, int32_t* __restrict__ unit_item, int32_t* __restrict__ unit_tile) {
  for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < n; q += (int64_t)gridDim.x * blockDim.x) {
    if (flag[q]) {
      const int32_t id = uid[q], p = perm[q];
      unit_begin[id] = (int32_t)q;
      unit_item[id] = item_of[p];
      unit_tile[id] = irow[p] / tile_users;
    }
  }
}

// length of each unit, its (tile, kUnitLen - len) sort key and the first unit of every tile (unit ids ascend with the tile)
__global__ void unit_len_kernel(const int32_t* __restrict__ unit_begin, const int32_t* __restrict__ unit_tile, int32_t n_units, int64_t n,
                                int32_t* __restrict__ unit_len, uint32_t* __restrict__ sort_key, int32_t* __restrict__ ids,
                                int32_t* __restrict__ tile_first) {
  const int32_t id = blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= n_units) return;
  const int32_t b = unit_begin[id];
  const int32_t e = (id + 1 < n_units) ? unit_begin[id + 1] : (int32_t)n;
  const int32_t len = e - b;
  unit_len[id] = len;
  sort_key[id] = ((uint32_t)unit_tile[id] << kUnitBits) | (uint32_t)(kUnitLen - len);
  ids[id] = id;
  if (id == 0 || unit_tile[id - 1] != unit_tile[id]) tile_first[unit_tile[id]] = id;
}

// sorted index j -> slot (slice*32 + lane); lane 0 of a slice holds its longest unit: the slice is as many rows high as
// that unit needs (per_row entries of a unit share a row)
__global__ void slot_assign_kernel(const int32_t* __restrict__ sorted_id, const int32_t* __restrict__ unit_tile,
                                   const int32_t* __restrict__ unit_len, const int32_t* __restrict__ tile_unit_ptr,
                                   const int32_t* __restrict__ tile_slice_ptr, int32_t n_units, int32_t per_row,
                                   int32_t* __restrict__ unit_slot, int32_t* __restrict__ slice_rows) {
  const int32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n_units) return;
  const int32_t id = sorted_id[j];
  const int32_t t = unit_tile[id];
  const int32_t r = j - tile_unit_ptr[t];
  const int32_t slice = tile_slice_ptr[t] + (r >> 5), lane = r & 31;
  unit_slot[id] = slice * 32 + lane;
  if (lane == 0) slice_rows[slice] = (unit_len[id] + per_row - 1) / per_row;
}

__global__ void fill_u32_kernel(uint32_t* __restrict__ p, int64_t n, uint32_t v) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = v;
}

// POPULAR entry: 16 bits, index (code - code_min) * kPopTileUsers + (user local to the tile) into the tile's deviation table
// dev[code][user]; two entries of a unit per 32-bit word (entry 2j in the low half, 2j+1 in the high half of row j)
__global__ void entry_fill_pop_kernel(const int32_t* __restrict__ unit_begin, const int32_t* __restrict__ unit_len,
                                      const int32_t* __restrict__ unit_slot, const int32_t* __restrict__ unit_tile, int32_t n_units,
                                      const int32_t* __restrict__ perm, const int32_t* __restrict__ irow, const uint8_t* __restrict__ ival,
                                      const int32_t* __restrict__ slice_off, int32_t code_min, int32_t n_codes, uint16_t* __restrict__ entry) {
  const int32_t id = blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= n_units) return;
  const int32_t b = unit_begin[id], len = unit_len[id], slot = unit_slot[id];
  const int32_t slice = slot >> 5, lane = slot & 31;
  const int64_t row0 = slice_off[slice];
  const int32_t ubase = unit_tile[id] * kPopTileUsers;
  for (int32_t j = 0; j < len; ++j) {
    const int32_t p = perm[b + j];
    entry[((((row0 + (j >> 1)) << 5) + lane) << 1) + (j & 1)] = (uint16_t)((((int32_t)ival[p] - code_min) * kPopTileUsers) + (irow[p] - ubase));
  }
}

// RARE entry: 32 bits, code << 20 | (user local to the tile) << 3 -- the low 20 bits are the byte offset of the user's
// (code sum, count) pair in shared memory
__global__ void entry_fill_rare_kernel(const int32_t* __restrict__ unit_begin, const int32_t* __restrict__ unit_len,
                                       const int32_t* __restrict__ unit_slot, const int32_t* __restrict__ unit_tile, int32_t n_units,
                                       const int32_t* __restrict__ perm, const int32_t* __restrict__ irow, const uint8_t* __restrict__ ival,
                                       const int32_t* __restrict__ slice_off, uint32_t* __restrict__ entry) {
  const int32_t id = blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= n_units) return;
  const int32_t b = unit_begin[id], len = unit_len[id], slot = unit_slot[id];
  const int32_t slice = slot >> 5, lane = slot & 31;
  const int64_t row0 = slice_off[slice];
  const int32_t ubase = unit_tile[id] * kRareTileUsers;
  for (int32_t j = 0; j < len; ++j) {
    const int32_t p = perm[b + j];
    entry[((row0 + j) << 5) + lane] = ((uint32_t)ival[p] << 20) | ((uint32_t)(irow[p] - ubase) << 3);
  }
}

__global__ void slot_item_kernel(const int32_t* __restrict__ unit_slot, const int32_t* __restrict__ unit_item, int32_t n_units,
                                 int32_t* __restrict__ slot_item) {
  const int32_t id = blockIdx.x * blockDim.x + threadIdx.x;
  if (id < n_units) slot_item[unit_slot[id]] = unit_item[id];
}

// item-major codes padded to 16-byte vectors of one item each (for the per-item rating sums): one warp per item
__global__ void ivec_count_kernel(const int32_t* __restrict__ icolp, int32_t n_items, int32_t* __restrict__ cnt) {
  const int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_items) cnt[i] = (icolp[i + 1] - icolp[i] + 15) >> 4;
}
__global__ void ivec_fill_kernel(const uint8_t* __restrict__ ival, const int32_t* __restrict__ icolp, const int32_t* __restrict__ vcol,
                                 int32_t n_items, uint8_t* __restrict__ ival16, int32_t* __restrict__ vec_col) {
  const int lane = threadIdx.x & 31;
  for (int32_t i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); i < n_items; i += gridDim.x * (blockDim.x >> 5)) {
    const int32_t b = icolp[i], e = icolp[i + 1];
    const int64_t dst = (int64_t)vcol[i] << 4;
    for (int32_t p = b + lane; p < e; p += 32) ival16[dst + (p - b)] = ival[p];
    for (int32_t v = vcol[i] + lane; v < vcol[i + 1]; v += 32) vec_col[v] = i;
  }
}

// ------------------------------------------------------------------------------------------------ the pass
constexpr int kPassThreads = 1024;
constexpr int kPassWarps = kPassThreads / 32;
constexpr int kRows = 8;     // 128-byte rows per ring stage (1 KB)
constexpr int kStages = 2;   // ring depth per warp
constexpr double kFixScale = 1099511627776.0;  // 2^40
// shared memory: table (the larger of the two parts' tables + one dummy slot) | rings | barriers
constexpr size_t kPopTableBytes = ((size_t)kPopTileUsers * kMaxCodes + 1) * 8;
constexpr size_t kRareTableBytes = ((size_t)kRareTileUsers + 1) * 8;
constexpr size_t kTableBytes = ((kPopTableBytes > kRareTableBytes ? kPopTableBytes : kRareTableBytes) + 127) / 128 * 128;
constexpr size_t kRingBytes = (size_t)kPassWarps * kStages * kRows * 128;
constexpr size_t kPassSmem = kTableBytes + kRingBytes + (size_t)kPassWarps * kStages * 8;
static_assert(kPassSmem <= 232448, "item pass: shared memory budget of one CTA");

struct PassArgs {
  const uint32_t *entry_pop, *entry_rare;
  const int32_t *slice_off_pop, *slice_off_rare;
  const int32_t *slot_item_pop, *slot_item_rare;
  const int2* warp_part;
  const int3* cta_desc;
  int32_t n_pop_tiles;
  int32_t code_min, n_codes;
  const uint32_t* usum;
  const int32_t* urow;
  int32_t n_users;
  double* uavg;
  long long* xdev_fix;
};

// full-precision reciprocal of a double that holds an integer of at most 20 significant bits: the hardware seed reads
// only the high word (exact for such values) and is good to ~2^-21; r(1 + e + e^2) leaves e^3 ~ 2^-63
__device__ __forceinline__ double rcp_small_int(double d) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
  const double e = fma(-d, r, 1.0);
  return fma(r, fma(e, e, e), r);
}
// int32 -> double without the conversion unit (I2F is a quarter-rate XU instruction: 36 % XU pipe load in the round-1
// kernel): the integer goes into the low mantissa word of 2^52 + 2^31 and the constant is subtracted
__device__ __forceinline__ double int_to_double(int32_t v) {
  return __hiloint2double(0x43300000, (int)((uint32_t)v ^ 0x80000000u)) - 4503601774854144.0;
}

// Deviation of one rating in exact integer form.  With r = code/2 and avg = S/(2c) (S = the user's code sum, c = its
// rating count):   r - avg = (c*code - S)/(2c),   5 - avg = (10c - S)/(2c),   avg - 1 = (S - 2c)/(2c)
// so (r - avg)/scale(r, avg) (P:57-61, P:167) = N/D with N = c*code - S and D = 10c - S (N > 0), S - 2c (N < 0) or
// 1 (N = 0: r == avg, the reference's 0/1): two small integers and ONE rounding (the reference rounds the average, the
// difference and the quotient: <= 2 ulp apart).  r > avg <=> N > 0 exactly, so the branch is the reference's.
__device__ __forceinline__ double dev_from_counts(uint32_t S, uint32_t c, uint32_t code) {
  const int32_t N = (int32_t)(c * code) - (int32_t)S;
  int32_t D = (int32_t)S - 2 * (int32_t)c;
  if (N > 0) D = 10 * (int32_t)c - (int32_t)S;
  if (N == 0) D = 1;
  return int_to_double(N) * rcp_small_int(int_to_double(D));
}

__device__ __forceinline__ void hand_over(int32_t item, double acc, long long* __restrict__ xdev_fix) {
  if (item >= 0) atomicAdd(reinterpret_cast<unsigned long long*>(xdev_fix + item), (unsigned long long)__double2ll_rn(acc * kFixScale));
}

__global__ void __launch_bounds__(kPassThreads, 1) item_pass_kernel(const PassArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* s_tab = reinterpret_cast<double*>(smem_raw);                       // popular: dev[user][code]; rare: (S, c) pairs
  uint32_t* s_ring = reinterpret_cast<uint32_t*>(smem_raw + kTableBytes);    // [warps][kStages][kRows*32]
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(smem_raw + kTableBytes + kRingBytes);
  const int3 cd = a.cta_desc[blockIdx.x];
  const bool is_pop = cd.x < a.n_pop_tiles;
  const int32_t tile = is_pop ? cd.x : cd.x - a.n_pop_tiles;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  uint64_t* bar = s_bar + wid * kStages;
  uint32_t* ring = s_ring + (size_t)wid * kStages * kRows * 32;
  const uint32_t* __restrict__ entry = is_pop ? a.entry_pop : a.entry_rare;
  const int32_t* __restrict__ slice_off = is_pop ? a.slice_off_pop : a.slice_off_rare;
  const int32_t* __restrict__ slot_item = is_pop ? a.slot_item_pop : a.slot_item_rare;

  if (lane == 0) {
#pragma unroll
    for (int st = 0; st < kStages; ++st) tma::mbar_init(bar + st, 1);
    tma::fence_barrier_init();
  }
  __syncwarp();

  // ---- this warp's slices (static partition) and the first ring stages: everything here reads the layout only, so it
  // overlaps the tail of K1 (programmatic dependent launch)
  const int2 wp = __ldg(a.warp_part + (size_t)blockIdx.x * kPassWarps + wid);
  int32_t cur = wp.x;
  const int32_t s_hi = wp.y;
  const int32_t r0 = (cur < s_hi) ? __ldg(slice_off + cur) : 0;
  const int32_t r_end = (cur < s_hi) ? __ldg(slice_off + s_hi) : 0;
  const int32_t n_chunks = (r_end - r0 + kRows - 1) / kRows;
  if (lane == 0) {
#pragma unroll
    for (int st = 0; st < kStages; ++st) {
      if (st < n_chunks) {
        const int32_t rr = r0 + st * kRows;
        const uint32_t bytes = (uint32_t)min(kRows, r_end - rr) * 128u;
        tma::mbar_arrive_expect_tx(bar + st, bytes);
        tma::bulk_g2s(ring + st * kRows * 32, entry + ((int64_t)rr << 5), bytes, bar + st);
      }
    }
  }
  int32_t end1 = (cur < s_hi) ? __ldg(slice_off + cur + 1) : 0x7fffffff;        // end row of the current slice
  int32_t end2 = (cur + 1 < s_hi) ? __ldg(slice_off + cur + 2) : 0x7fffffff;    // ... of the next one (prefetched)
  int32_t item1 = (cur < s_hi) ? __ldg(slot_item + cur * 32 + lane) : -1;       // item of this lane's unit in the current slice
  int32_t item2 = (cur + 1 < s_hi) ? __ldg(slot_item + (cur + 1) * 32 + lane) : -1;

  pdl_trigger();  // K2b may be scheduled as SMs free up
  pdl_wait();     // K1's per-user code sums are complete from here on

  // ---- per-user averages (P:113 / P:274): every CTA writes an equal share of the user table (exact sum, one correctly
  // rounded division, P:18; -1.0 = no ratings, the reference's own sentinel P:222)
  {
    const int32_t per = (a.n_users + gridDim.x - 1) / gridDim.x;
    const int32_t u_lo = blockIdx.x * per, u_hi = min(a.n_users, u_lo + per);
    for (int32_t u = u_lo + threadIdx.x; u < u_hi; u += kPassThreads) {
      const uint32_t S = __ldg(a.usum + u);
      const uint32_t cnt = (uint32_t)(__ldg(a.urow + u + 1) - __ldg(a.urow + u));
      a.uavg[u] = cnt ? (0.5 * (double)S) / (double)cnt : -1.0;
    }
  }

  double acc = 0.0;
  if (is_pop) {
    // ---- deviation table of the tile: dev[j][x] for code code_min + j and user u0 + x.  One thread per user: the two
    // reciprocals of its scale() values once, then one multiplication per code; consecutive threads write consecutive
    // slots of a column (no bank conflicts)
    const int32_t u0 = tile * kPopTileUsers;
    const int32_t nc = a.n_codes;
    const int32_t n_slots = kPopTileUsers * nc;
#pragma unroll
    for (int32_t x = threadIdx.x; x < kPopTileUsers; x += kPassThreads) {
      const int32_t u = u0 + x;
      const bool in = u < a.n_users;
      const int32_t S = in ? (int32_t)__ldg(a.usum + u) : 0;
      const int32_t cnt = in ? __ldg(a.urow + u + 1) - __ldg(a.urow + u) : 0;
      const double inv_hi = rcp_small_int(int_to_double(10 * cnt - S));  // 1 / (5 - avg) up to the common factor 2c
      const double inv_lo = rcp_small_int(int_to_double(S - 2 * cnt));   // 1 / (avg - 1)
      for (int32_t j = 0; j < nc; ++j) {
        const int32_t N = cnt * (a.code_min + j) - S;
        double dev = int_to_double(N) * (N > 0 ? inv_hi : inv_lo);
        if (N == 0 || cnt == 0) dev = 0.0;  // r == avg: the reference's 0/1 (also keeps 0 * inf out when avg is exactly 1 or 5)
        s_tab[j * kPopTileUsers + x] = dev;
      }
    }
    if (threadIdx.x == 0) s_tab[n_slots] = 0.0;  // padding entries point here
    __syncthreads();

    for (int32_t c = 0; c < n_chunks; ++c) {
      const int st = c % kStages;
      const int32_t r = r0 + c * kRows;
      const int32_t nrows = min(kRows, r_end - r);
      tma::mbar_wait(bar + st, (uint32_t)(c / kStages) & 1u);
      const uint32_t* rp = ring + st * kRows * 32 + lane;  // conflict free: lane l reads word l of a row
      const uint32_t pad = (uint32_t)n_slots | ((uint32_t)n_slots << 16);
      uint32_t w[kRows];
      if (nrows == kRows) {
#pragma unroll
        for (int k = 0; k < kRows; ++k) w[k] = rp[k * 32];
      } else {
#pragma unroll
        for (int k = 0; k < kRows; ++k) w[k] = (k < nrows) ? rp[k * 32] : pad;
      }
      __syncwarp();
      if (lane == 0 && c + kStages < n_chunks) {  // the stage is free again: request the chunk kStages ahead
        const int32_t rr = r + kStages * kRows;
        const uint32_t bytes = (uint32_t)min(kRows, r_end - rr) * 128u;
        tma::mbar_arrive_expect_tx(bar + st, bytes);
        tma::bulk_g2s(ring + st * kRows * 32, entry + ((int64_t)rr << 5), bytes, bar + st);
      }
#pragma unroll
      for (int h = 0; h < kRows; h += 4) {
        double d0[4], d1[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {  // all gathers of the half stage go out together
          d0[k] = *reinterpret_cast<const double*>(smem_raw + ((w[h + k] << 3) & 0x7fff8u));
          d1[k] = *reinterpret_cast<const double*>(smem_raw + ((w[h + k] >> 13) & 0x7fff8u));
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {  // ordered accumulation + slice boundaries (warp-uniform branch)
          if (r + h + k == end1) {     // the slice ended with the previous row: hand this lane's unit sum over
            hand_over(item1, acc, a.xdev_fix);
            acc = 0.0;
            ++cur;
            end1 = end2; item1 = item2;
            end2 = (cur + 1 < s_hi) ? __ldg(slice_off + cur + 2) : 0x7fffffff;
            item2 = (cur + 1 < s_hi) ? __ldg(slot_item + (cur + 1) * 32 + lane) : -1;
          }
          acc += d0[k];
          acc += d1[k];
        }
      }
    }
  } else {
    // ---- (code sum, count) of the tile's users; slot kRareTileUsers is the dummy user of padding entries: S = 0, c = 1
    // gives N = 0 for code 0, i.e. a deviation of exactly 0
    uint2* s_sc = reinterpret_cast<uint2*>(s_tab);
    const int32_t u0 = tile * kRareTileUsers;
    constexpr int kPer = kRareTileUsers / kPassThreads;
#pragma unroll
    for (int k0 = 0; k0 < kPer; k0 += 8) {  // 8 users per thread at a time: their 24 loads go out together
      uint32_t S[8];
      int32_t b0[8], b1[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int32_t u = u0 + (k0 + k) * kPassThreads + threadIdx.x;
        const bool in = u < a.n_users;
        S[k] = in ? __ldg(a.usum + u) : 0u;
        b0[k] = in ? __ldg(a.urow + u) : 0;
        b1[k] = in ? __ldg(a.urow + u + 1) : 0;
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) s_sc[(k0 + k) * kPassThreads + threadIdx.x] = make_uint2(S[k], (uint32_t)(b1[k] - b0[k]));
    }
    if (threadIdx.x == 0) s_sc[kRareTileUsers] = make_uint2(0u, 1u);
    __syncthreads();
    const unsigned char* tab = reinterpret_cast<const unsigned char*>(s_tab);
    constexpr uint32_t kPadEntry = (uint32_t)kRareTileUsers << 3;

    for (int32_t c = 0; c < n_chunks; ++c) {
      const int st = c % kStages;
      const int32_t r = r0 + c * kRows;
      const int32_t nrows = min(kRows, r_end - r);
      tma::mbar_wait(bar + st, (uint32_t)(c / kStages) & 1u);
      const uint32_t* rp = ring + st * kRows * 32 + lane;
      uint32_t ev[kRows];
      if (nrows == kRows) {
#pragma unroll
        for (int k = 0; k < kRows; ++k) ev[k] = rp[k * 32];
      } else {
#pragma unroll
        for (int k = 0; k < kRows; ++k) ev[k] = (k < nrows) ? rp[k * 32] : kPadEntry;
      }
      __syncwarp();
      if (lane == 0 && c + kStages < n_chunks) {
        const int32_t rr = r + kStages * kRows;
        const uint32_t bytes = (uint32_t)min(kRows, r_end - rr) * 128u;
        tma::mbar_arrive_expect_tx(bar + st, bytes);
        tma::bulk_g2s(ring + st * kRows * 32, entry + ((int64_t)rr << 5), bytes, bar + st);
      }
      double dv[kRows];
#pragma unroll
      for (int k = 0; k < kRows; ++k) {  // heavy part: no branches, 8 independent chains
        const uint2 sc = *reinterpret_cast<const uint2*>(tab + (ev[k] & 0xffff8u));
        dv[k] = dev_from_counts(sc.x, sc.y, ev[k] >> 20);
      }
#pragma unroll
      for (int k = 0; k < kRows; ++k) {
        if (r + k == end1) {
          hand_over(item1, acc, a.xdev_fix);
          acc = 0.0;
          ++cur;
          end1 = end2; item1 = item2;
          end2 = (cur + 1 < s_hi) ? __ldg(slice_off + cur + 2) : 0x7fffffff;
          item2 = (cur + 1 < s_hi) ? __ldg(slot_item + (cur + 1) * 32 + lane) : -1;
        }
        acc += dv[k];
      }
    }
  }
  if (cur < s_hi) hand_over(item1, acc, a.xdev_fix);  // last slice of the range
}

// K2b: per item, integer accumulators -> exchange buffer (and re-arm them for the next pass); optionally finish the fit
__global__ void __launch_bounds__(256) item_tiled_finalize_kernel(long long* __restrict__ xdev_fix, uint32_t* __restrict__ xcode_sum,
                                                                 const int32_t* __restrict__ icolp, int32_t n_items,
                                                                 unsigned long long* __restrict__ k1_part, double n_total,
                                                                 double* __restrict__ xbuf, int fused, double* __restrict__ idevavg,
                                                                 double* __restrict__ iavg, double* __restrict__ gavg) {
  pdl_trigger();
  pdl_wait();  // the accumulators are complete once the item pass has finished
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    const double gs = 0.5 * (double)k1_part[0];  // integer sum of codes: exact, order independent
    k1_part[0] = 0;                              // re-arm for the next pass
    xbuf[2 * (size_t)n_items] = gs;
    xbuf[2 * (size_t)n_items + 1] = n_total;
    if (fused) gavg[0] = n_total > 0.0 ? gs / n_total : 0.0;
  }
  const int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_items) return;
  const double ds = (double)xdev_fix[i] * (1.0 / kFixScale);
  const double rs = 0.5 * (double)xcode_sum[i];
  xdev_fix[i] = 0;
  xcode_sum[i] = 0;
  const double cnt = (double)(icolp[i + 1] - icolp[i]);
  xbuf[i] = ds;
  xbuf[(size_t)n_items + i] = cnt;
  xbuf[2 * (size_t)n_items + 2 + i] = rs;
  if (fused) {
    idevavg[i] = cnt > 0.0 ? ds / cnt : 0.0;
    iavg[i] = cnt > 0.0 ? rs / cnt : nan("");
  }
}

// ------------------------------------------------------------------------------------------------ host side
struct PartTemps {
  int32_t* item_of = nullptr;
  uint8_t* pop_flag = nullptr;
};

void free_part(mrs_ratings::ell_part& P) {
  dev_free(P.entry); dev_free(P.slice_off); dev_free(P.slot_item);
  P = mrs_ratings::ell_part();
}

// one sliced-ELL part: the entries of the items with pop_flag == want, tiled by `tile_users`
int32_t build_part(const mrs_ratings* R, const PartTemps& tmp_in, int32_t want, int32_t tile_users, int64_t n_part, int32_t code_min,
                   int32_t n_codes, mrs_ratings::ell_part& P) {
  mrs_engine* e = R->eng;
  cudaStream_t st = e->stream;
  const int64_t n = R->n;
  const bool pop = (want == 1);
  const int32_t per_row = pop ? 2 : 1;
  const int32_t NT = (R->n_users + tile_users - 1) / tile_users;
  MRS_REQUIRE(NT < 65535, MRS_ERR_UNSUPPORTED, "too many user tiles (%d)", NT);
  P.tile_users = tile_users;
  P.n_tiles = NT;
  P.n_entries = n_part;
  P.h_tile_slice.assign((size_t)NT + 1, 0);
  P.h_slice_off.assign(1, 0);
  const uint32_t pad_word = pop ? (uint32_t)(kPopTileUsers * n_codes) * 0x10001u : ((uint32_t)kRareTileUsers << 3);
  if (n_part == 0) {
    MRS_TRY(dev_alloc(&P.slice_off, 1));
    MRS_CUDA(cudaMemsetAsync(P.slice_off, 0, sizeof(int32_t), st));
    MRS_TRY(dev_alloc(&P.entry, 32));
    MRS_TRY(dev_alloc(&P.slot_item, 32));
    return MRS_OK;
  }
  const int block = 256;
  const int grid = grid_for(n, block, e->sm_count);
  const int pgrid = grid_for(n_part, block, e->sm_count);
  // ---- (tile, item, user) order of the part's entries: stable radix sort of the CSC positions on the tile id
  uint16_t *tk_in = nullptr, *tk_out = nullptr;
  int32_t *pos_in = nullptr, *perm = nullptr, *head = nullptr, *seg_start = nullptr, *flag = nullptr, *uid = nullptr;
  MRS_TRY(dev_alloc(&tk_in, (size_t)n)); MRS_TRY(dev_alloc(&tk_out, (size_t)n));
  MRS_TRY(dev_alloc(&pos_in, (size_t)n)); MRS_TRY(dev_alloc(&perm, (size_t)n));
  MRS_TRY(dev_alloc(&head, (size_t)n_part)); MRS_TRY(dev_alloc(&seg_start, (size_t)n_part));
  MRS_TRY(dev_alloc(&flag, (size_t)n_part)); MRS_TRY(dev_alloc(&uid, (size_t)n_part + 1));
  part_keys_kernel<<<grid, block, 0, st>>>(R->irow, tmp_in.item_of, tmp_in.pop_flag, want, tile_users, NT, n, tk_in, pos_in);
  int tbits = 1;
  while ((1 << tbits) < NT + 1) ++tbits;
  size_t tmp = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, tmp, tk_in, tk_out, pos_in, perm, (int)n, 0, tbits, st);
  MRS_TRY(ensure_scratch(e, tmp));
  cub::DeviceRadixSort::SortPairs(e->scratch, tmp, tk_in, tk_out, pos_in, perm, (int)n, 0, tbits, st);
  // ---- units (the first n_part sorted positions are this part's)
  seg_head_kernel<<<pgrid, block, 0, st>>>(perm, tmp_in.item_of, R->irow, tile_users, n_part, head);
  cub::DeviceScan::InclusiveScan(nullptr, tmp, head, seg_start, MaxOp(), (int)n_part, st);
  MRS_TRY(ensure_scratch(e, tmp));
  cub::DeviceScan::InclusiveScan(e->scratch, tmp, head, seg_start, MaxOp(), (int)n_part, st);
  unit_flag_kernel<<<pgrid, block, 0, st>>>(seg_start, n_part, flag);
  cub::DeviceScan::ExclusiveSum(nullptr, tmp, flag, uid, (int)n_part, st);
  MRS_TRY(ensure_scratch(e, tmp));
  cub::DeviceScan::ExclusiveSum(e->scratch, tmp, flag, uid, (int)n_part, st);
  int32_t last_uid = 0, last_flag = 0;
  MRS_CUDA(cudaMemcpyAsync(&last_uid, uid + (n_part - 1), sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  MRS_CUDA(cudaMemcpyAsync(&last_flag, flag + (n_part - 1), sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  MRS_CUDA(cudaStreamSynchronize(st));
  const int32_t NUN = last_uid + last_flag;
  P.n_units = NUN;
  int32_t *unit_begin = nullptr, *unit_item = nullptr, *unit_tile = nullptr, *unit_len = nullptr, *ids = nullptr, *sorted_id = nullptr;
  int32_t *tile_first = nullptr, *unit_slot = nullptr, *slice_rows = nullptr, *d_tile_unit_ptr = nullptr, *d_tile_slice = nullptr;
  uint32_t *skey = nullptr, *skey_out = nullptr;
  MRS_TRY(dev_alloc(&unit_begin, (size_t)NUN)); MRS_TRY(dev_alloc(&unit_item, (size_t)NUN)); MRS_TRY(dev_alloc(&unit_tile, (size_t)NUN));
  MRS_TRY(dev_alloc(&unit_len, (size_t)NUN)); MRS_TRY(dev_alloc(&ids, (size_t)NUN)); MRS_TRY(dev_alloc(&sorted_id, (size_t)NUN));
  MRS_TRY(dev_alloc(&skey, (size_t)NUN)); MRS_TRY(dev_alloc(&skey_out, (size_t)NUN));
  MRS_TRY(dev_alloc(&tile_first, (size_t)NT + 1)); MRS_TRY(dev_alloc(&unit_slot, (size_t)NUN));
  MRS_TRY(dev_alloc(&d_tile_unit_ptr, (size_t)NT + 1)); MRS_TRY(dev_alloc(&d_tile_slice, (size_t)NT + 1));
  MRS_CUDA(cudaMemsetAsync(tile_first, 0xff, sizeof(int32_t) * ((size_t)NT + 1), st));  // -1: tile without units
  unit_scatter_kernel<<<pgrid, block, 0, st>>>(flag, uid, perm, tmp_in.item_of, R->irow, tile_users, n_part, unit_begin, unit_item, unit_tile);
  const int ugrid = (NUN + block - 1) / block;
  unit_len_kernel<<<ugrid, block, 0, st>>>(unit_begin, unit_tile, NUN, n_part, unit_len, skey, ids, tile_first);
  // ---- sort units by (tile, length desc); stable => canonical order among equal lengths
  cub::DeviceRadixSort::SortPairs(nullptr, tmp, skey, skey_out, ids, sorted_id, NUN, 0, kUnitBits + tbits, st);
  MRS_TRY(ensure_scratch(e, tmp));
  cub::DeviceRadixSort::SortPairs(e->scratch, tmp, skey, skey_out, ids, sorted_id, NUN, 0, kUnitBits + tbits, st);
  std::vector<int32_t> h_first((size_t)NT + 1, 0), h_unit_ptr((size_t)NT + 1, 0);
  MRS_CUDA(cudaMemcpyAsync(h_first.data(), tile_first, sizeof(int32_t) * (size_t)NT, cudaMemcpyDeviceToHost, st));
  MRS_CUDA(cudaStreamSynchronize(st));
  h_first[(size_t)NT] = NUN;  // first unit of every tile; tiles without units take the next tile's first
  for (int32_t t = NT - 1; t >= 0; --t)
    if (h_first[(size_t)t] < 0) h_first[(size_t)t] = h_first[(size_t)t + 1];
  for (int32_t t = 0; t < NT; ++t) {
    const int32_t cnt = h_first[(size_t)t + 1] - h_first[(size_t)t];
    h_unit_ptr[(size_t)t + 1] = h_unit_ptr[(size_t)t] + cnt;
    P.h_tile_slice[(size_t)t + 1] = P.h_tile_slice[(size_t)t] + (cnt + 31) / 32;
  }
  const int32_t NS = P.h_tile_slice[(size_t)NT];
  P.n_slices = NS;
  MRS_CUDA(cudaMemcpyAsync(d_tile_unit_ptr, h_unit_ptr.data(), sizeof(int32_t) * ((size_t)NT + 1), cudaMemcpyHostToDevice, st));
  MRS_CUDA(cudaMemcpyAsync(d_tile_slice, P.h_tile_slice.data(), sizeof(int32_t) * ((size_t)NT + 1), cudaMemcpyHostToDevice, st));
  MRS_TRY(dev_alloc(&slice_rows, (size_t)NS + 1));
  MRS_TRY(dev_alloc(&P.slice_off, (size_t)NS + 1));
  MRS_CUDA(cudaMemsetAsync(slice_rows, 0, sizeof(int32_t) * ((size_t)NS + 1), st));
  slot_assign_kernel<<<ugrid, block, 0, st>>>(sorted_id, unit_tile, unit_len, d_tile_unit_ptr, d_tile_slice, NUN, per_row, unit_slot, slice_rows);
  cub::DeviceScan::ExclusiveSum(nullptr, tmp, slice_rows, P.slice_off, NS + 1, st);
  MRS_TRY(ensure_scratch(e, tmp));
  cub::DeviceScan::ExclusiveSum(e->scratch, tmp, slice_rows, P.slice_off, NS + 1, st);
  P.h_slice_off.assign((size_t)NS + 1, 0);
  MRS_CUDA(cudaMemcpyAsync(P.h_slice_off.data(), P.slice_off, sizeof(int32_t) * ((size_t)NS + 1), cudaMemcpyDeviceToHost, st));
  MRS_CUDA(cudaStreamSynchronize(st));  // also keeps h_unit_ptr / h_tile_slice alive until their copies are done
  P.n_rows = P.h_slice_off[(size_t)NS];
  MRS_TRY(dev_alloc(&P.entry, (size_t)P.n_rows * 32 + 32));
  fill_u32_kernel<<<grid_for(P.n_rows * 32, block, e->sm_count), block, 0, st>>>(P.entry, P.n_rows * 32, pad_word);
  if (pop)
    entry_fill_pop_kernel<<<ugrid, block, 0, st>>>(unit_begin, unit_len, unit_slot, unit_tile, NUN, perm, R->irow, (const uint8_t*)R->ival, P.slice_off,
                                                  code_min, n_codes, reinterpret_cast<uint16_t*>(P.entry));
  else
    entry_fill_rare_kernel<<<ugrid, block, 0, st>>>(unit_begin, unit_len, unit_slot, unit_tile, NUN, perm, R->irow, (const uint8_t*)R->ival, P.slice_off,
                                                   P.entry);
  // ---- item of every slot (empty slots: -1)
  MRS_TRY(dev_alloc(&P.slot_item, (size_t)NS * 32));
  MRS_CUDA(cudaMemsetAsync(P.slot_item, 0xff, sizeof(int32_t) * (size_t)NS * 32, st));
  slot_item_kernel<<<ugrid, block, 0, st>>>(unit_slot, unit_item, NUN, P.slot_item);
  count_launch(18);
  MRS_CUDA(cudaGetLastError());
  MRS_CUDA(cudaStreamSynchronize(st));
  for (void* p : {(void*)tk_in, (void*)tk_out, (void*)pos_in, (void*)perm, (void*)head, (void*)seg_start, (void*)flag, (void*)uid,
                  (void*)unit_begin, (void*)unit_item, (void*)unit_tile, (void*)unit_len, (void*)ids, (void*)sorted_id, (void*)skey,
                  (void*)skey_out, (void*)tile_first, (void*)unit_slot, (void*)slice_rows, (void*)d_tile_unit_ptr, (void*)d_tile_slice})
    dev_free(p);
  return MRS_OK;
}

int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return (v && *v) ? atoi(v) : dflt;
}

// first slice s in [lo, hi) whose cost prefix (rows before it + slice_cost * slices before it) is >= v
int32_t lower_bound_cost(const std::vector<int32_t>& slice_off, int32_t lo, int32_t hi, int32_t slice_cost, int64_t v) {
  const int32_t s0 = lo;
  const int64_t row_base = slice_off[(size_t)lo];
  while (lo < hi) {
    const int32_t mid = (lo + hi) >> 1;
    if ((int64_t)slice_off[(size_t)mid] - row_base + (int64_t)slice_cost * (mid - s0) < v) lo = mid + 1; else hi = mid;
  }
  return lo;
}

}  // namespace

void free_tiled_layout(const mrs_ratings* R) {
  auto& T = R->tl;
  free_part(T.pop); free_part(T.rare);
  dev_free(T.cta_desc); dev_free(T.warp_part); dev_free(T.ival16); dev_free(T.vec_col);
  T = mrs_ratings::tiled_layout();
}

int32_t build_tiled_layout(const mrs_ratings* R) {
  auto& T = R->tl;
  if (T.built) return MRS_OK;
  MRS_REQUIRE(R->value_kind == kValueCode, MRS_ERR_INVALID, "tiled layout needs half-star codes");
  mrs_engine* e = R->eng;
  cudaStream_t st = e->stream;
  const int64_t n = R->n;
  const int32_t NI = R->n_items;
  PartTemps tmp;
  int32_t* d_stats = nullptr;
  unsigned long long* d_count = nullptr;
  MRS_TRY(dev_alloc(&tmp.item_of, (size_t)std::max<int64_t>(n, 1)));
  MRS_TRY(dev_alloc(&tmp.pop_flag, (size_t)NI + 1));
  MRS_TRY(dev_alloc(&d_stats, 2));
  MRS_TRY(dev_alloc(&d_count, 1));
  const int32_t init_stats[2] = {255, 0};
  MRS_CUDA(cudaMemcpyAsync(d_stats, init_stats, sizeof(init_stats), cudaMemcpyHostToDevice, st));
  MRS_CUDA(cudaMemsetAsync(d_count, 0, sizeof(unsigned long long), st));
  int32_t h_stats[2] = {255, 0};
  if (n > 0) {
    item_of_kernel<<<grid_for(n, 256, e->sm_count), 256, 0, st>>>(R->icolp, NI, n, tmp.item_of);
    code_range_kernel<<<grid_for(n, 256, e->sm_count), 256, 0, st>>>((const uint8_t*)R->ival, n, d_stats);
    MRS_CUDA(cudaMemcpyAsync(h_stats, d_stats, sizeof(h_stats), cudaMemcpyDeviceToHost, st));
    MRS_CUDA(cudaStreamSynchronize(st));
  }
  T.code_min = (n > 0) ? h_stats[0] : 0;
  T.n_codes = (n > 0) ? h_stats[1] - h_stats[0] + 1 : 1;
  // popular = long enough (tile, item) runs with tiles of kPopTileUsers users: about 6 ratings per tile on average.
  // A set with more than kMaxCodes distinct codes has no popular part (its deviation table would not fit).
  const int32_t pop_tiles = (R->n_users + kPopTileUsers - 1) / kPopTileUsers;
  int32_t thr = env_int("MRS_POP_THRESHOLD", 0);
  if (thr <= 0) thr = std::max(32, env_int("MRS_POP_PER_TILE", 6) * pop_tiles);
  if (T.n_codes > kMaxCodes) thr = INT_MAX;
  T.pop_threshold = thr;
  unsigned long long h_pop = 0;
  pop_flag_kernel<<<(NI + 255) / 256, 256, 0, st>>>(R->icolp, NI, thr, tmp.pop_flag, d_count);
  MRS_CUDA(cudaMemcpyAsync(&h_pop, d_count, sizeof(h_pop), cudaMemcpyDeviceToHost, st));
  MRS_CUDA(cudaStreamSynchronize(st));
  count_launch(3);
  MRS_TRY(build_part(R, tmp, 1, kPopTileUsers, (int64_t)h_pop, T.code_min, T.n_codes, T.pop));
  MRS_TRY(build_part(R, tmp, 0, kRareTileUsers, n - (int64_t)h_pop, T.code_min, T.n_codes, T.rare));
  dev_free(tmp.item_of); dev_free(tmp.pop_flag); dev_free(d_stats); dev_free(d_count);

  // ---- static work partition: CTAs dealt out to the tiles of both parts by cost, then slices to the warps of each CTA.
  // Cost model (issue slots, from the SASS): a popular row (64 ratings) ~ 14 instructions per lane, a rare row (32
  // ratings) ~ 34, handing a slice over ~ 40 (the atomics occupy the load/store unit for about that long).
  const int w_pop_row = env_int("MRS_W_POP_ROW", 14), w_rare_row = env_int("MRS_W_RARE_ROW", 34), w_slice = env_int("MRS_W_SLICE", 40);
  const int32_t pop_slice_cost = std::max(1, w_slice / w_pop_row), rare_slice_cost = std::max(1, w_slice / w_rare_row);
  const int32_t NTP = T.pop.n_tiles, NTR = T.rare.n_tiles;
  std::vector<int64_t> cost((size_t)NTP + NTR, 0);
  auto tile_cost = [](const mrs_ratings::ell_part& P, int32_t t, int w_row, int32_t slice_cost) -> int64_t {
    if (P.n_slices == 0) return 0;
    const int32_t s0 = P.h_tile_slice[(size_t)t], s1 = P.h_tile_slice[(size_t)t + 1];
    return ((int64_t)(P.h_slice_off[(size_t)s1] - P.h_slice_off[(size_t)s0]) + (int64_t)slice_cost * (s1 - s0)) * w_row;
  };
  for (int32_t t = 0; t < NTP; ++t) cost[(size_t)t] = tile_cost(T.pop, t, w_pop_row, pop_slice_cost);
  for (int32_t t = 0; t < NTR; ++t) cost[(size_t)NTP + t] = tile_cost(T.rare, t, w_rare_row, rare_slice_cost);
  const std::vector<int3> desc = deal_ctas(cost, e->sm_count);
  T.n_ctas = (int32_t)desc.size();
  std::vector<int2> wpart(std::max<size_t>(1, desc.size()) * kPassWarps, make_int2(0, 0));
  for (size_t b = 0; b < desc.size(); ++b) {
    const bool pop = desc[b].x < NTP;
    const mrs_ratings::ell_part& P = pop ? T.pop : T.rare;
    const int32_t t = pop ? desc[b].x : desc[b].x - NTP;
    const int32_t slice_cost = pop ? pop_slice_cost : rare_slice_cost;
    const int32_t ts0 = P.h_tile_slice[(size_t)t], ts1 = P.h_tile_slice[(size_t)t + 1];
    const int64_t total = (int64_t)(P.h_slice_off[(size_t)ts1] - P.h_slice_off[(size_t)ts0]) + (int64_t)slice_cost * (ts1 - ts0);
    const int64_t nw = (int64_t)desc[b].z * kPassWarps;
    for (int w = 0; w < kPassWarps; ++w) {
      const int64_t gw = (int64_t)desc[b].y * kPassWarps + w;
      const int32_t lo = lower_bound_cost(P.h_slice_off, ts0, ts1, slice_cost, total * gw / nw);
      const int32_t hi = lower_bound_cost(P.h_slice_off, ts0, ts1, slice_cost, total * (gw + 1) / nw);
      wpart[b * kPassWarps + w] = make_int2(lo, hi);
    }
  }
  MRS_TRY(dev_alloc(&T.cta_desc, std::max<size_t>(1, desc.size())));
  MRS_TRY(dev_alloc(&T.warp_part, wpart.size()));
  if (!desc.empty()) MRS_CUDA(cudaMemcpyAsync(T.cta_desc, desc.data(), sizeof(int3) * desc.size(), cudaMemcpyHostToDevice, st));
  MRS_CUDA(cudaMemcpyAsync(T.warp_part, wpart.data(), sizeof(int2) * wpart.size(), cudaMemcpyHostToDevice, st));
  MRS_CUDA(cudaStreamSynchronize(st));  // the host vectors must outlive the copies; the pass reads the tables in its prologue
  T.built = true;
  return MRS_OK;
}

// item-major codes as 16-byte vectors of one item each (only when per-item rating averages are wanted, P:134)
int32_t build_item_vectors(const mrs_ratings* R) {
  auto& T = R->tl;
  if (T.ival16 || R->n == 0) return MRS_OK;
  mrs_engine* e = R->eng;
  cudaStream_t st = e->stream;
  const int32_t NI = R->n_items;
  int32_t *cnt = nullptr, *vcol = nullptr;
  MRS_TRY(dev_alloc(&cnt, (size_t)NI + 1));
  MRS_TRY(dev_alloc(&vcol, (size_t)NI + 1));
  MRS_CUDA(cudaMemsetAsync(cnt, 0, sizeof(int32_t) * ((size_t)NI + 1), st));
  ivec_count_kernel<<<(NI + 255) / 256, 256, 0, st>>>(R->icolp, NI, cnt);
  size_t tmp = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, tmp, cnt, vcol, NI + 1, st);
  MRS_TRY(ensure_scratch(e, tmp));
  cub::DeviceScan::ExclusiveSum(e->scratch, tmp, cnt, vcol, NI + 1, st);
  int32_t n_vec = 0;
  MRS_CUDA(cudaMemcpyAsync(&n_vec, vcol + NI, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  MRS_CUDA(cudaStreamSynchronize(st));
  T.n_ivec = n_vec;
  MRS_TRY(dev_alloc(&T.ival16, (size_t)std::max(n_vec, 1) * 16));
  MRS_TRY(dev_alloc(&T.vec_col, (size_t)std::max(n_vec, 1)));
  MRS_CUDA(cudaMemsetAsync(T.ival16, 0, (size_t)std::max(n_vec, 1) * 16, st));
  ivec_fill_kernel<<<std::max(1, std::min((NI + 7) / 8, e->sm_count * 8)), 256, 0, st>>>((const uint8_t*)R->ival, R->icolp, vcol, NI, T.ival16, T.vec_col);
  count_launch(3);
  MRS_CUDA(cudaGetLastError());
  MRS_CUDA(cudaStreamSynchronize(st));
  dev_free(cnt); dev_free(vcol);
  return MRS_OK;
}

int32_t launch_item_tiled(mrs_engine* e, const mrs_ratings* R, mrs_model* m, bool fused) {
  const auto& T = R->tl;
  cudaStream_t st = e->stream;
  if (!(e->smem_attr_done & 1u)) {
    MRS_CUDA(cudaFuncSetAttribute(item_pass_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPassSmem));
    e->smem_attr_done |= 1u;
  }
  if (T.n_ctas > 0) {
    PassArgs a;
    a.entry_pop = T.pop.entry; a.entry_rare = T.rare.entry;
    a.slice_off_pop = T.pop.slice_off; a.slice_off_rare = T.rare.slice_off;
    a.slot_item_pop = T.pop.slot_item; a.slot_item_rare = T.rare.slot_item;
    a.warp_part = T.warp_part; a.cta_desc = T.cta_desc;
    a.n_pop_tiles = T.pop.n_tiles;
    a.code_min = T.code_min; a.n_codes = T.n_codes;
    a.usum = m->usum; a.urow = R->urow; a.n_users = R->n_users; a.uavg = m->uavg; a.xdev_fix = m->xdev_fix;
    // one CTA of 1024 threads per SM: the grid (T.n_ctas <= SM count unless there are more busy tiles than SMs) is one wave
    MRS_CUDA(launch_pdl(item_pass_kernel, dim3(T.n_ctas), dim3(kPassThreads), kPassSmem, st, a));
    mark(e, "item_tiled");
  }
  MRS_CUDA(launch_pdl(item_tiled_finalize_kernel, dim3((R->n_items + 255) / 256), dim3(256), 0, st, m->xdev_fix, m->xcode_sum, R->icolp, R->n_items,
                      m->k1_part, (double)R->n, m->xbuf, fused ? 1 : 0, m->idevavg, m->iavg, m->gavg));
  mark(e, "item_tiled_finalize");
  MRS_CUDA(cudaGetLastError());
  return MRS_OK;
}

}  // namespace mrs
