// tiled.cu -- the item-deviation pass of the fit (P:155-186 / P:316-343) on a layout made for B200:
//
//   * users are cut into tiles of at most kTileCap users; a CTA builds the tile's 8-byte user records in shared memory
//     (128 KB) from K1's integer code sums and the row pointer, so the 20 M random 8-byte gathers of the pass hit
//     shared-memory banks instead of L2 sectors (first capture: 640 MB of L2 sector traffic, L1 hit rate 11 %);
//   * inside a tile the entries are item-major; every (tile, item) segment is cut into units of <= kUnitLen
//     entries, units are sorted by length and packed 32 to a "slice" in sliced-ELL order (entry j of lane l at
//     row j, column l); a warp streams a run of 128-byte rows through its own TMA ring (cp.async.bulk + mbarrier) and
//     every lane accumulates ITS unit sequentially: no cross-lane reduction, a fixed summation order inside a unit;
//   * an entry is 4 bytes.  Form 1 (every code <= kAlphaMaxCode): the top 15 bits of the fp64 number (code - 2)/8 and the
//     byte offset of the user's record; form 0: valid bit | half-star code | local user.
//
// Deviation = N / D with N = c*code - S and D = 10c - S (N > 0) or S - 2c (N < 0): two exact small integers formed in
// fp64 from the entry and the record (S = the user's code sum, c = rating count), one hardware reciprocal seed refined by
// three DFMA, ONE rounding -- within 2 ulp of the reference's three (average, difference, quotient), far inside the 1e-6
// parity bound; r > avg <=> N > 0 exactly, so the branch of scale() is the reference's.  The kNN path, whose neighbour
// ranking needs the exact bits, computes its own deviations with true divisions (knn.cu).
// Unit sums are handed over as integer atomics on a 2^-40 grid: exact, so the per-item sums do not depend on the order
// in which warps, CTAs or ranks deliver them and the pass is bit-reproducible.  The rows of a tile are cut over its warps
// at row granularity by a static partition laid down with the layout (item_partition_kernel).
#include <cub/cub.cuh>

#include <algorithm>
#include <cstdlib>
#include <vector>

#include "common.cuh"
#include "tma.cuh"

namespace mrs {
namespace {

struct MaxOp {
  __device__ __forceinline__ int32_t operator()(int32_t a, int32_t b) const { return a > b ? a : b; }
};

// item of every CSC position: one warp per item writes its id over the item's column (a binary search per entry in the
// column pointer cost 250 us at ml-25m shape)
__global__ void item_expand_kernel(const int32_t* __restrict__ icolp, int32_t n_items, int32_t* __restrict__ item_of) {
  const int lane = threadIdx.x & 31;
  for (int32_t i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); i < n_items; i += gridDim.x * (blockDim.x >> 5)) {
    const int32_t b = icolp[i], e = icolp[i + 1];
    for (int32_t p = b + lane; p < e; p += 32) item_of[p] = i;
  }
}
// tile of every user: tiles are contiguous user ranges [ubegin[t], ubegin[t+1]) (users between two tiles have no ratings)
__global__ void user_tile_kernel(const int32_t* __restrict__ ubegin, int32_t n_tiles, uint16_t* __restrict__ utile) {
  const int32_t t = blockIdx.x;
  for (int32_t u = ubegin[t] + threadIdx.x; u < ubegin[t + 1]; u += blockDim.x) utile[u] = (uint16_t)t;
}
// the sort key is the user tile only; the radix sort is stable so (item, user) order survives inside a tile
__global__ void tile_keys_kernel(const int32_t* __restrict__ irow, const uint16_t* __restrict__ utile, int64_t n, uint16_t* __restrict__ tile_key,
                                 int32_t* __restrict__ pos) {
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x) {
    tile_key[p] = utile[irow[p]];
    pos[p] = (int32_t)p;
  }
}

// q = position in (tile, item, user) order.  head_pos[q] = q at the first entry of a (tile,item) segment, else 0
__global__ void seg_head_kernel(const int32_t* __restrict__ perm, const int32_t* __restrict__ item_of, const int32_t* __restrict__ irow,
                                const uint16_t* __restrict__ utile, int64_t n, int32_t* __restrict__ head_pos) {
  for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < n; q += (int64_t)gridDim.x * blockDim.x) {
    bool head = (q == 0);
    if (!head) {
      const int32_t p = perm[q], pp = perm[q - 1];
      head = (item_of[p] != item_of[pp]) || (utile[irow[p]] != utile[irow[pp]]);
    }
    head_pos[q] = head ? (int32_t)q : 0;
  }
}

__global__ void unit_flag_kernel(const int32_t* __restrict__ seg_start, int64_t n, int32_t* __restrict__ flag) {
  for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < n; q += (int64_t)gridDim.x * blockDim.x)
    flag[q] = ((q - seg_start[q]) % kUnitLen == 0) ? 1 : 0;
}

__global__ void unit_scatter_kernel(const int32_t* __restrict__ flag, const int32_t* __restrict__ uid, const int32_t* __restrict__ perm,
                                    const int32_t* __restrict__ item_of, const int32_t* __restrict__ irow, const uint16_t* __restrict__ utile,
                                    int64_t n, int32_t* __restrict__ unit_begin, int32_t* __restrict__ unit_item, int32_t* __restrict__ unit_tile) {
  for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < n; q += (int64_t)gridDim.x * blockDim.x) {
    if (flag[q]) {
      const int32_t id = uid[q], p = perm[q];
      unit_begin[id] = (int32_t)q;
      unit_item[id] = item_of[p];
      unit_tile[id] = utile[irow[p]];
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

// sorted index j -> slot (slice*32 + lane); lane 0 of a slice holds its longest unit
__global__ void slot_assign_kernel(const int32_t* __restrict__ sorted_id, const int32_t* __restrict__ unit_tile,
                                   const int32_t* __restrict__ unit_len, const int32_t* __restrict__ tile_unit_ptr,
                                   const int32_t* __restrict__ tile_slice_ptr, int32_t n_units, int32_t* __restrict__ unit_slot,
                                   int32_t* __restrict__ slice_width) {
  const int32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n_units) return;
  const int32_t id = sorted_id[j];
  const int32_t t = unit_tile[id];
  const int32_t r = j - tile_unit_ptr[t];
  const int32_t slice = tile_slice_ptr[t] + (r >> 5), lane = r & 31;
  unit_slot[id] = slice * 32 + lane;
  if (lane == 0) slice_width[slice] = unit_len[id];
}

// Entry word of one rating.  Form 0 (any half-star code): valid bit | code << 16 | local user.  Form 1 (codes <= kAlphaMaxCode,
// i.e. every MovieLens-style scale): the pass works in fp64 on exact small integers, so the word carries what the inner loop
// would otherwise have to compute or convert per rating: bits 31..17 = the top 15 bits (sign, exponent, 3 mantissa bits) of
// the fp64 number alpha = (code - 2)/8 -- all of it, the rest of the mantissa is zero -- and bits 15..3 = the local user,
// already scaled to the byte offset of its 8-byte table record.  A padding slot names the last record of the table, which
// is all zero: that makes its deviation exactly 0 without a test.
constexpr int kAlphaMaxCode = 17;
constexpr uint32_t kUserMask = (uint32_t)(kTileUsers - 1) << 3;  // byte offset of a user record inside the tile's table
constexpr uint32_t kAlphaPadding = kUserMask;                    // alpha = 0, the all-zero record behind the tile's users
static_assert(kTileUsers <= 16384 && (kTileUsers & (kTileUsers - 1)) == 0, "the user field of an entry is bits 16..3");
__device__ __forceinline__ uint32_t make_entry(int form, uint32_t code, uint32_t local_user) {
  if (form == 0) return 0x80000000u | (code << 16) | local_user;
  const double alpha = (double)((int32_t)code - 2) * 0.125;
  return ((uint32_t)__double2hiint(alpha) & 0xfffe0000u) | (local_user << 3);
}

__global__ void fill_u32_kernel(uint32_t* __restrict__ p, int64_t n, uint32_t v) {
  for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < n; q += (int64_t)gridDim.x * blockDim.x) p[q] = v;
}

// largest code of the set and largest rating count of a user: decide the entry form
__global__ void form_scan_kernel(const uint8_t* __restrict__ ival, int64_t n, const int32_t* __restrict__ urow, int32_t n_users, int32_t* __restrict__ out2) {
  int32_t mc = 0, mu = 0;
  for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < n; q += (int64_t)gridDim.x * blockDim.x) mc = max(mc, (int32_t)ival[q]);
  for (int32_t u = blockIdx.x * blockDim.x + threadIdx.x; u < n_users; u += gridDim.x * blockDim.x) mu = max(mu, urow[u + 1] - urow[u]);
  mc = __reduce_max_sync(0xffffffffu, mc);
  mu = __reduce_max_sync(0xffffffffu, mu);
  if ((threadIdx.x & 31) == 0) { atomicMax(out2, mc); atomicMax(out2 + 1, mu); }
}

__global__ void entry_fill_kernel(const int32_t* __restrict__ unit_begin, const int32_t* __restrict__ unit_len,
                                  const int32_t* __restrict__ unit_slot, const int32_t* __restrict__ unit_tile, int32_t n_units,
                                  const int32_t* __restrict__ perm, const int32_t* __restrict__ irow, const uint8_t* __restrict__ ival,
                                  const int32_t* __restrict__ slice_off, uint32_t* __restrict__ entry, int form, const int32_t* __restrict__ tile_ubegin) {
  const int32_t id = blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= n_units) return;
  const int32_t b = unit_begin[id], len = unit_len[id], slot = unit_slot[id];
  const int32_t slice = slot >> 5, lane = slot & 31;
  const int64_t row0 = slice_off[slice];
  const int32_t ubase = tile_ubegin[unit_tile[id]];
  for (int32_t j = 0; j < len; ++j) {
    const int32_t p = perm[b + j];
    entry[((row0 + j) << 5) + lane] = make_entry(form, (uint32_t)ival[p], (uint32_t)(irow[p] - ubase));
  }
}

__global__ void slot_item_kernel(const int32_t* __restrict__ unit_slot, const int32_t* __restrict__ unit_item, int32_t n_units,
                                 int32_t* __restrict__ slot_item, int32_t* __restrict__ slot_unit) {
  const int32_t id = blockIdx.x * blockDim.x + threadIdx.x;
  if (id < n_units) {
    slot_item[unit_slot[id]] = unit_item[id];
    if (slot_unit) slot_unit[unit_slot[id]] = id;
  }
}

// Entry placement that avoids shared-memory bank conflicts in the pass (MRS_REORDER=1; A/B in profiles/r02_summary.md).  The
// pass gathers the 8-byte (code sum, count) pair of every entry's user; a 64-bit shared-memory load is served 16 lanes at a
// time and is conflict free only if those lanes hit 16 different bank pairs = (user & 15).  With entries in (item, user)
// order the keys of a row are random: ~3.2 wavefronts per half warp instead of 1.  The order of the entries INSIDE a unit is
// free (it only fixes the summation order), so one warp per slice deals them out position by position: every lane offers
// an entry whose key is still free in its half warp (lowest lane wins a contested key, the others offer another key in the
// next round); a lane without such an entry leaves the position to padding if it has slack, else takes a conflict.
__global__ void __launch_bounds__(128) entry_fill_ordered_kernel(const int32_t* __restrict__ slot_unit, const int32_t* __restrict__ unit_begin,
                                                                const int32_t* __restrict__ unit_len, const int32_t* __restrict__ unit_tile,
                                                                int32_t n_slices, const int32_t* __restrict__ perm,
                                                                const int32_t* __restrict__ irow, const uint8_t* __restrict__ ival,
                                                                const int32_t* __restrict__ slice_off, uint32_t* __restrict__ entry, int form,
                                                                const int32_t* __restrict__ tile_ubegin) {
  __shared__ uint32_t s_ent[4][32 * kUnitLen];   // entries of the lane's unit, grouped by key
  __shared__ uint8_t s_next[4][32][16];          // next unplaced entry of every key group
  __shared__ uint8_t s_end[4][32][16];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int32_t slice = blockIdx.x * 4 + w;
  if (slice >= n_slices) return;
  const int32_t id = slot_unit[slice * 32 + lane];
  const int32_t len = id >= 0 ? unit_len[id] : 0;
  const int32_t b = id >= 0 ? unit_begin[id] : 0;
  const int32_t ubase = id >= 0 ? tile_ubegin[unit_tile[id]] : 0;
  uint32_t* ent = s_ent[w] + lane * kUnitLen;
  uint8_t* nxt = s_next[w][lane];
  uint8_t* end = s_end[w][lane];
  int32_t cnt[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) cnt[k] = 0;
  for (int32_t j = 0; j < len; ++j) {
    const int32_t key = (irow[perm[b + j]] - ubase) & 15;
#pragma unroll
    for (int k = 0; k < 16; ++k) cnt[k] += (key == k);
  }
  uint32_t avail = 0;
  {
    int32_t run = 0;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      nxt[k] = (uint8_t)run;
      run += cnt[k];
      end[k] = (uint8_t)run;
      if (cnt[k]) avail |= 1u << k;
    }
  }
  for (int32_t j = 0; j < len; ++j) {
    const int32_t p = perm[b + j];
    const int32_t x = irow[p] - ubase;
    ent[nxt[x & 15]++] = make_entry(form, (uint32_t)ival[p], (uint32_t)x);
  }
  {
    int32_t run = 0;
#pragma unroll
    for (int k = 0; k < 16; ++k) { nxt[k] = (uint8_t)run; run += cnt[k]; }
  }
  __syncwarp();
  const int64_t row0 = slice_off[slice];
  const int32_t n_pos = slice_off[slice + 1] - slice_off[slice];
  const int half = lane >> 4;
  int32_t remaining = len;
  for (int32_t pos = 0; pos < n_pos; ++pos) {
    uint32_t taken = 0;
    int32_t mine = -1;
    const bool need = remaining > 0;
    const int rot = (lane + pos) & 15;
    while (true) {
      const uint32_t cand = (need && mine < 0) ? (avail & ~taken) : 0u;
      int32_t prop = -1;
      if (cand) {
        const uint32_t r = ((cand >> rot) | (cand << (16 - rot))) & 0xffffu;
        prop = (__ffs(r) - 1 + rot) & 15;
      }
      if (!__any_sync(0xffffffffu, prop >= 0)) break;
      const uint32_t same = __match_any_sync(0xffffffffu, prop >= 0 ? (prop | (half << 4)) : (64 + lane));
      const bool win = prop >= 0 && (__ffs(same) - 1 == lane);
      if (win) mine = prop;
      const uint32_t won = __reduce_or_sync(0xffffffffu, win ? (1u << (prop + 16 * half)) : 0u);
      taken |= (won >> (16 * half)) & 0xffffu;
    }
    if (need && mine < 0 && remaining >= n_pos - pos) mine = __ffs(avail) - 1;  // no slack left: take a conflict
    if (mine >= 0) {
      const uint32_t e = ent[nxt[mine]++];
      if (nxt[mine] == end[mine]) avail &= ~(1u << mine);
      --remaining;
      entry[((row0 + pos) << 5) + lane] = e;
    }
  }
}

int grid_for(int64_t n, int block, int sm_count) {
  return (int)std::max<int64_t>(1, std::min<int64_t>((n + block - 1) / block, (int64_t)sm_count * 16));
}

// ----------------------------------------------------------------------------------------------------------------
// full-precision reciprocal without the IEEE division sequence: hardware seed (good to ~2^-21) + r(1 + e + e^2), which
// leaves e^3 ~ 2^-63 with 3 DFMA (two Newton steps take 4).  Measured on the same box (tools/ab.sh): replacing the two I2F
// of the deviation by mantissa tricks (LOP + DADD) made the pass 3 us SLOWER -- the fp64 pipe, not the conversion unit,
// is the scarce one here.
__device__ __forceinline__ double fast_rcp(double s) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(s));
#ifdef MRS_AB_NEWTON2
  r = fma(r, fma(-s, r, 1.0), r);
  r = fma(r, fma(-s, r, 1.0), r);
  return r;
#else
  const double e = fma(-s, r, 1.0);
  return fma(r, fma(e, e, e), r);
#endif
}

// K2: one CTA = (tile, share of the tile's work), 32 warps, one CTA per SM.
// Shared memory: (code sum, rating count) of the tile's users (64 KB, code sums from K1) + a private ring of
// kStages x kRows 128-byte rows per warp filled by 1-D bulk asynchronous copies (TMA) and signalled through mbarriers:
// ~100 KB of entries are in flight per SM without costing registers.  Slices of a tile are contiguous in memory, so a
// warp that owns a contiguous range of slices streams one contiguous run of rows; slice boundaries only decide when a
// lane's unit sum is handed over.  Unit sums are fp64 (fixed order inside the unit); they are combined across units
// with integer atomics on a 2^-40 grid, which is exact, so the item sums do not depend on the order of arrival.
constexpr int kTiledThreads = 1024;
#ifndef MRS_K2_ROWS
#define MRS_K2_ROWS 8
#endif
#ifndef MRS_K2_STAGES
#define MRS_K2_STAGES 2
#endif
constexpr int kRows = MRS_K2_ROWS;      // rows per ring stage (128 B each)
constexpr int kStages = MRS_K2_STAGES;  // ring depth per warp
constexpr int kSliceCost = 4;  // rows a slice boundary is worth (per-warp stamps, least squares: 0.092 us per row, 0.33 us per slice)
constexpr double kFixScale = 1099511627776.0;  // 2^40
constexpr size_t kTabBytes = (size_t)kTileUsers * 8;  // the tile's user records; the last one is the all-zero record of the padding slots
constexpr size_t kTiledSmem = kTabBytes + (size_t)(kTiledThreads / 32) * kStages * kRows * 128 + (size_t)(kTiledThreads / 32 * kStages) * 8;

// first slice s in [lo, hi) whose cost prefix (rows before it + kSliceCost * slices before it) is >= v
__device__ __forceinline__ int32_t lower_bound_cost(const int32_t* __restrict__ slice_off, int32_t lo, int32_t hi, int32_t row_base, int32_t v) {
  const int32_t s0 = lo;
  while (lo < hi) {
    const int32_t mid = (lo + hi) >> 1;
    if (__ldg(slice_off + mid) - row_base + kSliceCost * (mid - s0) < v) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// Deviation of one entry in exact integer form; branch free so that the rows of a batch overlap in the pipeline.
// With r = code/2 and avg = S/(2c) (S = the user's code sum, c = its rating count):
//     r - avg = (c*code - S)/(2c),   5 - avg = (10c - S)/(2c),   avg - 1 = (S - 2c)/(2c)
// so (r - avg)/scale(r, avg) (P:57-61, P:167) = N/D with N = c*code - S and D = 10c - S (N > 0) or S - 2c (N < 0):
// two small integers, ONE rounding (the reference rounds the average, the difference and the quotient: <= 2 ulp apart).
// r > avg <=> N > 0 exactly (|avg - r| >= 1/(2c) whenever they differ, far above an ulp), so the branch is the reference's.
__device__ __forceinline__ double tiled_dev(uint32_t e, const uint2* __restrict__ s_sc) {
  const uint32_t code = (e >> 16) & 0xffu;              // 0 for padding
  const uint2 sc = s_sc[e & 0xffffu];                   // .x = S, .y = c
  const int32_t N = (int32_t)(sc.y * code) - (int32_t)sc.x;
  const int32_t D = N > 0 ? (int32_t)(10u * sc.y) - (int32_t)sc.x : (int32_t)sc.x - (int32_t)(2u * sc.y);
  const double dev = (double)N * fast_rcp((double)D);
  return (N != 0 && (int32_t)e < 0) ? dev : 0.0;        // r == avg -> 0/1 = 0; padding (valid bit clear) contributes nothing
}

// Form 1: the same quotient N/D with every per-rating integer operation and conversion moved out of the loop.  The record
// of a user holds the HIGH WORDS of the two fp64 numbers Dup = 10c - S and Ddn = S - 2c (integers below 2^21: their low
// words are zero), the entry holds the high word of alpha = (code - 2)/8.  With t = Dup + Ddn = 8c,
//     N = c*code - S = alpha*t - Ddn          (one DADD, one DFMA, both exact: small integers)
// the sign of N is read off its high word, D is picked by a 32-bit select, and N == 0 (r == avg, P:60 scale 1, and every
// padding slot, whose record is all zero) takes D = 1.0 so that the product is an exact 0.  Per rating: 2 shared-memory
// loads, 1 MUFU, 7 fp64 operations and a handful of integer ones, against 2 I2F + MUFU + 5 fp64 + ~20 integer before.
__device__ __forceinline__ double tiled_dev_alpha(uint32_t e, const unsigned char* __restrict__ s_tab) {
#ifdef MRS_AB_NOGATHER  // timing experiment only: every entry reads the record of its lane (conflict free)
  const uint2 rec = *reinterpret_cast<const uint2*>(s_tab + ((threadIdx.x & 31u) << 3) + (e & 0x100u));
#else
  const uint2 rec = *reinterpret_cast<const uint2*>(s_tab + (e & kUserMask));
#endif
#ifdef MRS_AB_NOMATH    // timing experiment only: no reciprocal
  return __hiloint2double((int)rec.x, (int)(e & 0xfffe0000u));
#endif
  const double dup = __hiloint2double((int)rec.x, 0), ddn = __hiloint2double((int)rec.y, 0);
  const double alpha = __hiloint2double((int)(e & 0xfffe0000u), 0);
  const double N = fma(alpha, dup + ddn, -ddn);
  const int nh = __double2hiint(N);
  int dh = nh > 0 ? (int)rec.x : (int)rec.y;
  dh = nh == 0 ? 0x3ff00000 : dh;  // N is an exact integer and never -0.0: alpha*t is +0 or non-zero, and +0 - (+0) = +0
  return N * fast_rcp(__hiloint2double(dh, 0));
}
// code of a form-1 entry (only the pass that also sums the ratings per item needs it): 8*alpha + 2, 0 for padding
__device__ __forceinline__ uint32_t alpha_code(uint32_t e) {
  const double alpha = __hiloint2double((int)(e & 0xfffe0000u), 0);
  return ((e & kUserMask) == kAlphaPadding) ? 0u : (uint32_t)(__double2int_rn(alpha * 8.0) + 2);
}

template <bool WITH_SUM>
__device__ __forceinline__ void hand_over(int32_t item, double acc, uint32_t csum, long long* __restrict__ xdev_fix,
                                          unsigned long long* __restrict__ xcode_sum) {
  if (item >= 0) {
    atomicAdd(reinterpret_cast<unsigned long long*>(xdev_fix + item), (unsigned long long)__double2ll_rn(acc * kFixScale));
    if (WITH_SUM) atomicAdd(xcode_sum + item, (unsigned long long)csum);
  }
}

// Rows [x, y) of every warp of the item pass, .z = the slice that holds row x, .w = one past the slice that holds row y-1: an
// equal share of its tile's cost (rows + kSliceCost per slice), cut at ROW granularity.  A range may begin or end inside a
// slice: the two warps then hand over one partial sum each for the units of that slice (exact integer atomics, so the item
// sums stay deterministic).  Cutting at whole slices left the warps with 2 or 3 of the 64-row slices of the popular items --
// 128 against 192 rows, and every CTA as slow as its 192-row warps (per-warp stamps: 33 us against 42 us).
__device__ __forceinline__ int2 cost_position(const int32_t* __restrict__ slice_off, int32_t ts0, int32_t ts1, int32_t ra, int32_t v) {
  const int32_t f = lower_bound_cost(slice_off, ts0, ts1, ra, v);  // first slice whose cost prefix is >= v
  if (f < ts1 && __ldg(slice_off + f) - ra + kSliceCost * (f - ts0) == v) return make_int2(f, __ldg(slice_off + f));
  const int32_t sl = f - 1;                                        // the slice that holds cost value v
  const int32_t b = __ldg(slice_off + sl), e = __ldg(slice_off + sl + 1);
  const int32_t off = v - (b - ra + kSliceCost * (sl - ts0)) - kSliceCost;  // the slice's fixed cost is paid at its start
  const int32_t row = b + max(0, min(off, e - b));
  return row == e ? make_int2(sl + 1, row) : make_int2(sl, row);
}

__global__ void __launch_bounds__(kTiledThreads) item_partition_kernel(const int32_t* __restrict__ slice_off, const int32_t* __restrict__ tile_slice_ptr,
                                                                      const int3* __restrict__ cta_desc, int4* __restrict__ warp_part) {
  if ((threadIdx.x & 31) != 0) return;
  const int3 cd = cta_desc[blockIdx.x];
  const int32_t tile = cd.x, share = cd.y, ctas_per_tile = cd.z;
  constexpr int32_t wpb = kTiledThreads >> 5;
  const int32_t wid = threadIdx.x >> 5;
  const int32_t ts0 = tile_slice_ptr[tile], ts1 = tile_slice_ptr[tile + 1];
  const int32_t ra = __ldg(slice_off + ts0), rb = __ldg(slice_off + ts1);
  const int32_t nw = ctas_per_tile * wpb, w = share * wpb + wid;
  const int64_t total_cost = (int64_t)(rb - ra) + (int64_t)kSliceCost * (ts1 - ts0);
  const int32_t c_lo = (int32_t)((total_cost * w) / nw), c_hi = (int32_t)((total_cost * (w + 1)) / nw);
  const int2 lo = (w == 0) ? make_int2(ts0, ra) : cost_position(slice_off, ts0, ts1, ra, c_lo);
  const int2 hi = (w == nw - 1) ? make_int2(ts1, rb) : cost_position(slice_off, ts0, ts1, ra, c_hi);
  const int32_t s_hi = (hi.x < ts1 && hi.y > __ldg(slice_off + hi.x)) ? hi.x + 1 : hi.x;
  warp_part[(size_t)blockIdx.x * wpb + wid] = (lo.y < hi.y) ? make_int4(lo.y, hi.y, lo.x, s_hi) : make_int4(0, 0, 0, 0);
}

template <bool WITH_SUM, int FORM>
__global__ void __launch_bounds__(kTiledThreads, 1) item_tiled_kernel(const uint32_t* __restrict__ entry, const int32_t* __restrict__ slice_off,
                                                                     const int4* __restrict__ warp_part, const int3* __restrict__ cta_desc,
                                                                     const uint32_t* __restrict__ usum, const int32_t* __restrict__ urow,
                                                                     int32_t n_users, const int32_t* __restrict__ slot_item,
                                                                     double* __restrict__ uavg, long long* __restrict__ xdev_fix_both,
                                                                     unsigned long long* __restrict__ xcode_sum, unsigned long long* __restrict__ tl,
                                                                     const unsigned int* __restrict__ parity, int32_t n_items_acc,
                                                                     const int32_t* __restrict__ tile_ubegin, unsigned long long* __restrict__ flags,
                                                                     int32_t k1_blocks) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  tl_begin(tl, 1);
  // the accumulator buffer of this pass (two buffers alternate when the test pass finishes the fit itself: it cannot re-arm
  // the buffer it reads, so it re-arms the other one; the parity only changes between passes)
  long long* __restrict__ xdev_fix = xdev_fix_both + (size_t)(*parity & 1u) * n_items_acc;
  uint2* s_sc = reinterpret_cast<uint2*>(smem_raw);                           // [kTileUsers (+ the padding record)] (code sum, count) or (Dup, Ddn)
  uint32_t* s_ring = reinterpret_cast<uint32_t*>(smem_raw + kTabBytes);       // [warps][kStages][kRows*32]
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(smem_raw + kTabBytes + (size_t)(kTiledThreads / 32) * kStages * kRows * 128);
  const int3 cd = cta_desc[blockIdx.x];
  const int32_t tile = cd.x, share = cd.y;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  constexpr int32_t wpb = kTiledThreads >> 5;
  uint64_t* bar = s_bar + wid * kStages;           // this warp's stage barriers
  uint32_t* ring = s_ring + (size_t)wid * kStages * kRows * 32;

  if (lane == 0) {
#pragma unroll
    for (int st = 0; st < kStages; ++st) tma::mbar_init(bar + st, 1);
    tma::fence_barrier_init();
  }
  __syncwarp();

  // ---- this warp's rows: an equal share of the tile's cost (a row costs 1, a slice boundary kSliceCost: handing over 32 unit
  // sums and fetching the next slice's items), cut at row granularity.  The partition is static: item_partition_kernel ran
  // the binary searches once, when the layout was built -- 28 dependent loads that used to open every pass.
  const int4 part = __ldg(warp_part + (size_t)blockIdx.x * wpb + wid);
  const int32_t r0 = part.x, r_end = part.y;   // rows of this warp; the first and the last slice may be shared with a neighbour
  int32_t cur = part.z;
  const int32_t s_hi = part.w;
  const int32_t n_chunks = (r_end - r0 + kRows - 1) / kRows;
  if (lane == 0) {  // the first kStages chunks go out before the averages are formed
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

  pdl_trigger();  // the next kernel may be scheduled as SMs free up
  // ---- the tile's users: (code sum from K1, rating count); the first CTA of the tile also publishes the averages.
  // The rating counts are layout data: they are fetched before the wait, only the code sums behind it.
  const int32_t u0 = __ldg(tile_ubegin + tile), u1 = __ldg(tile_ubegin + tile + 1);
  constexpr int kPer = kTileUsers / kTiledThreads;
  uint32_t cnt_of[kPer];
#pragma unroll
  for (int k = 0; k < kPer; ++k) {
    const int32_t u = u0 + k * kTiledThreads + (int32_t)threadIdx.x;
    cnt_of[k] = (u < u1) ? (uint32_t)(__ldg(urow + u + 1) - __ldg(urow + u)) : 0u;
  }
  // everything above (barriers, partition, first ring stages, counts) overlapped K1; the user sums are complete once all its
  // blocks have counted themselves off (k1_blocks > 0) or, without the flag protocol, once its grid has completed
  if (k1_blocks > 0) flag_wait(flags + 1, (unsigned long long)k1_blocks); else pdl_wait();
  {
    uint32_t S_of[kPer];
#pragma unroll
    for (int k = 0; k < kPer; ++k) {
      const int32_t u = u0 + k * kTiledThreads + (int32_t)threadIdx.x;
      S_of[k] = (u < u1) ? __ldcg(usum + u) : 0u;  // written by K1's atomics while this kernel was already resident: L2, not L1
    }
#pragma unroll
    for (int k = 0; k < kPer; ++k) {
      const int32_t x = k * kTiledThreads + threadIdx.x;
      const int32_t u = u0 + x;
      const bool in = u < u1;
      const uint32_t S = S_of[k], cnt = cnt_of[k];
      if (FORM == 0) {
        s_sc[x] = make_uint2(S, cnt);
      } else {  // high words of the fp64 numbers 10c - S and S - 2c (below 2^21 in magnitude: the low words are zero)
        s_sc[x] = make_uint2((uint32_t)__double2hiint((double)((int32_t)(10u * cnt) - (int32_t)S)),
                             (uint32_t)__double2hiint((double)((int32_t)S - (int32_t)(2u * cnt))));
      }
      if (share == 0 && in)  // exact sum, one correctly rounded division (P:18); -1.0: no ratings (the reference's sentinel, P:222)
        uavg[u] = cnt ? (0.5 * (double)S) / (double)cnt : -1.0;
    }
    // (the last record, kTileCap, belongs to no user -- a tile holds at most kTileCap of them -- and stays all zero: padding)
  }
  __syncthreads();
  tl_cta(tl, 0);

  double acc = 0.0;
  uint32_t csum = 0;
  for (int32_t c = 0; c < n_chunks; ++c) {
    const int st = c % kStages;
    const int32_t r = r0 + c * kRows;
    const int32_t nrows = min(kRows, r_end - r);
    tma::mbar_wait(bar + st, (uint32_t)(c / kStages) & 1u);
    const uint32_t* rp = ring + st * kRows * 32 + lane;  // conflict free: lane l reads word l of a row
#ifdef MRS_AB_NOBOUND  // timing experiment only (wrong sums): no slice boundaries, one hand-over per warp
    const bool plain = (nrows == kRows);
#else
    const bool plain = (nrows == kRows) && (end1 > r) && (end1 >= r + kRows);  // whole chunk inside the current slice
#endif
    uint32_t ev[kRows];
    if (plain) {
#pragma unroll
      for (int k = 0; k < kRows; ++k) ev[k] = rp[k * 32];
    } else {
#pragma unroll
      for (int k = 0; k < kRows; ++k) ev[k] = (k < nrows) ? rp[k * 32] : (FORM == 1 ? kAlphaPadding : 0u);
    }
    __syncwarp();
    if (lane == 0 && c + kStages < n_chunks) {  // the stage is free again: request the chunk kStages ahead
      const int32_t rr = r + kStages * kRows;
      const uint32_t bytes = (uint32_t)min(kRows, r_end - rr) * 128u;
      tma::mbar_arrive_expect_tx(bar + st, bytes);
      tma::bulk_g2s(ring + st * kRows * 32, entry + ((int64_t)rr << 5), bytes, bar + st);
    }
    double dv[kRows];
#pragma unroll
    for (int k = 0; k < kRows; ++k)  // heavy part: no branches, 8 independent chains
      dv[k] = FORM == 1 ? tiled_dev_alpha(ev[k], smem_raw) : tiled_dev(ev[k], s_sc);
    if (plain) {
#pragma unroll
      for (int k = 0; k < kRows; ++k) {
        acc += dv[k];
        if (WITH_SUM) csum += FORM == 1 ? alpha_code(ev[k]) : ((ev[k] >> 16) & 0xffu);
      }
    } else {
#pragma unroll
      for (int k = 0; k < kRows; ++k) {  // ordered accumulation + slice boundaries (warp-uniform branch)
        if (r + k == end1) {             // the slice ended with the previous row: hand this lane's unit sum over
          hand_over<WITH_SUM>(item1, acc, csum, xdev_fix, xcode_sum);
          acc = 0.0; csum = 0;
          ++cur;
          end1 = end2; item1 = item2;
          end2 = (cur + 1 < s_hi) ? __ldg(slice_off + cur + 2) : 0x7fffffff;
          item2 = (cur + 1 < s_hi) ? __ldg(slot_item + (cur + 1) * 32 + lane) : -1;
        }
        acc += dv[k];
        if (WITH_SUM) csum += FORM == 1 ? alpha_code(ev[k]) : ((ev[k] >> 16) & 0xffu);
      }
    }
  }
  if (cur < s_hi) hand_over<WITH_SUM>(item1, acc, csum, xdev_fix, xcode_sum);  // last slice of the range
#ifdef MRS_WARP_STAMPS
  if (tl && lane == 0 && blockIdx.x < 256) {
    tl[32 + 1024 + blockIdx.x * 32 + wid] = gtimer_ns();
    tl[32 + 1024 + 8192 + blockIdx.x * 32 + wid] = ((unsigned long long)(r_end - r0) << 32) | (unsigned int)(s_hi - part.z);
  }
#endif
  if (k1_blocks > 0) flag_count_off(flags + 2);  // opt-in flag protocol: the fused test pass waits for the count
  if (tl) { tl_cta(tl, 1); tl_end(tl, 1); }
}

// K2b: per item, integer accumulators -> exchange buffer (and re-arm them for the next pass); optionally finish the fit
__global__ void __launch_bounds__(256) item_tiled_finalize_kernel(long long* __restrict__ xdev_fix, unsigned long long* __restrict__ xcode_sum,
                                                                 const int32_t* __restrict__ icolp, int32_t n_items,
                                                                 unsigned long long* __restrict__ k1_part, int32_t n_k1, double n_total,
                                                                 double* __restrict__ xbuf, int fused, double* __restrict__ idevavg,
                                                                 double* __restrict__ iavg, double* __restrict__ gavg,
                                                                 uint32_t* __restrict__ usum, int32_t n_users,
                                                                 unsigned long long* __restrict__ tl, const unsigned int* __restrict__ parity) {
  tl_begin(tl, 2);
  xdev_fix += (size_t)(*parity & 1u) * n_items;
  pdl_trigger();
  pdl_wait();  // the accumulators are complete once the item pass has finished
  for (int32_t u = blockIdx.x * blockDim.x + threadIdx.x; u < n_users; u += gridDim.x * blockDim.x) usum[u] = 0;  // consumed by K2: re-arm for K1
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    const double gs = 0.5 * (double)k1_part[0];  // integer sum of codes: exact, order independent
    k1_part[0] = 0; k1_part[1] = 0; k1_part[2] = 0;  // re-arm the sum and the hand-over counts for the next pass
    xbuf[2 * (size_t)n_items] = gs;
    xbuf[2 * (size_t)n_items + 1] = n_total;
    if (fused) gavg[0] = n_total > 0.0 ? gs / n_total : 0.0;
  }
  const int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_items) { tl_end(tl, 2); return; }
  const double ds = (double)xdev_fix[i] * (1.0 / kFixScale);
  const double rs = 0.5 * (double)xcode_sum[i];
  xdev_fix[i] = 0;  // re-arm for the next pass
  xcode_sum[i] = 0;
  const double cnt = (double)(icolp[i + 1] - icolp[i]);
  xbuf[i] = ds;
  xbuf[(size_t)n_items + i] = cnt;
  xbuf[2 * (size_t)n_items + 2 + i] = rs;
  if (fused) {
    idevavg[i] = cnt > 0.0 ? ds / cnt : 0.0;
    iavg[i] = cnt > 0.0 ? rs / cnt : nan("");
  }
  tl_end(tl, 2);
}

// ---- fused compute + collective (sharded runs): K2b delivers this rank's per-item partial sums straight into every
// rank's symmetric receive buffer with NVLink stores (fire and forget) and raises a flag; the finishing kernel of every
// rank then waits for the flags and adds the deliveries FROM ITS OWN MEMORY in rank order (bit-identical totals
// everywhere).  Compared with K2b -> all-reduce kernel -> finalize kernel this saves one launch, the publish copy and
// the round trip of the remote loads.  Only the K items that occur on some rank travel (slot_of_item).
// One thread per KNOWN item (compact slot j, item known[j]): the 32 lanes of a warp store 256 contiguous bytes into every
// rank's buffer (a thread per item id would leave 9 of 32 lanes active at ml-25m shape, whose ids are sparse, and send
// 72-byte fragments over NVLink: 36 us for the kernel at 8 ranks against 10 us unfused).
__global__ void __launch_bounds__(256) item_tiled_push_kernel(long long* __restrict__ xdev_fix, const int32_t* __restrict__ icolp,
                                                             unsigned long long* __restrict__ k1_part, double n_total,
                                                             const int32_t* __restrict__ known, int32_t K, const PushDev x,
                                                             uint32_t* __restrict__ usum, int32_t n_users, unsigned long long* __restrict__ tl,
                                                             const unsigned int* __restrict__ acc_parity, int32_t n_items) {
  __shared__ int s_last;
  tl_begin(tl, 2);
  xdev_fix += (size_t)(*acc_parity & 1u) * n_items;
  pdl_trigger();
  pdl_wait();  // the accumulators are complete once the item pass has finished
  for (int32_t u = blockIdx.x * blockDim.x + threadIdx.x; u < n_users; u += gridDim.x * blockDim.x) usum[u] = 0;  // consumed by K2: re-arm for K1
  const unsigned long long epoch = *x.epoch + 1;  // stable until the finishing kernel of this exchange has run
  const int parity = (int)(epoch & 1);
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    const double gs = 0.5 * (double)k1_part[0];  // integer sum of codes: exact, order independent
    k1_part[0] = 0; k1_part[1] = 0; k1_part[2] = 0;  // re-arm the sum and the hand-over counts for the next pass
    for (int q = 0; q < x.world; ++q) {
      double* slot = push_slot(x, push_peer(x, q), parity, x.rank);
      slot[2 * (size_t)K] = gs;
      slot[2 * (size_t)K + 1] = n_total;
    }
  }
  const int32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < K) {
    const int32_t i = __ldg(known + j);
    const double ds = (double)xdev_fix[i] * (1.0 / kFixScale);
    xdev_fix[i] = 0;  // re-arm (items known on no rank never leave zero)
    const double cnt = (double)(__ldg(icolp + i + 1) - __ldg(icolp + i));
    for (int q = 0; q < x.world; ++q) {  // every rank starts with its own successor: no port of the switch takes all senders at once
      double* slot = push_slot(x, push_peer(x, q), parity, x.rank);
      slot[j] = ds;
      slot[(size_t)K + j] = cnt;
    }
  }
  // the last block of this rank makes all deliveries visible system-wide, then raises its flag in every rank's memory
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const int last = (atomicAdd(x.done, 1u) + 1u == gridDim.x);
    if (last) *x.done = 0;
    __threadfence();
    s_last = last;
  }
  __syncthreads();
  if (s_last && (int)threadIdx.x < x.world) {
    __threadfence_system();
    push_flag_raise(x, threadIdx.x, epoch);
  }
  tl_end(tl, 2);
}

__global__ void __launch_bounds__(256) item_finish_pull_kernel(int32_t n_items, const int32_t* __restrict__ known, int32_t K, const PushDev x,
                                                              double* __restrict__ xbuf, double* __restrict__ idevavg, double* __restrict__ gavg,
                                                              unsigned long long* __restrict__ tl) {
  tl_begin(tl, 4);
  pdl_trigger();  // the test pass may set up its rings while this runs
  pdl_wait();     // this rank's own delivery is complete (stream order)
  const unsigned long long epoch = *x.epoch + 1;
  const int parity = (int)(epoch & 1);
  int good = 1;
  if ((int)threadIdx.x < x.world) good = push_flag_wait(x, threadIdx.x, epoch) ? 1 : 0;  // every rank has delivered
  const bool ok = __syncthreads_and(good) != 0;
  const double bad = nan("");
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    double gs = 0.0, gc = 0.0;
    for (int p = 0; p < x.world; ++p) {  // rank order: identical totals on every rank
      const double* slot = push_slot(x, x.rank, parity, p);
      gs += slot[2 * (size_t)K];
      gc += slot[2 * (size_t)K + 1];
    }
    if (!ok) gs = gc = bad;  // a peer never delivered: nothing downstream may pass for a result
    xbuf[2 * (size_t)n_items] = gs;
    xbuf[2 * (size_t)n_items + 1] = gc;
    gavg[0] = ok ? (gc > 0.0 ? gs / gc : 0.0) : bad;  // P:18 mean of an empty Seq is 0.0
  }
  const int32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < K) {
    double ds = 0.0, cnt = 0.0;
    for (int p = 0; p < x.world; ++p) {
      const double* slot = push_slot(x, x.rank, parity, p);
      ds += slot[j];
      cnt += slot[(size_t)K + j];
    }
    if (!ok) ds = cnt = bad;
    const int32_t i = __ldg(known + j);
    xbuf[i] = ds;
    xbuf[(size_t)n_items + i] = cnt;
    idevavg[i] = ok ? (cnt > 0.0 ? ds / cnt : 0.0) : bad;  // P:185; items known on no rank keep 0.0 (P:197) from the model's creation
  }
  __syncthreads();
  if (threadIdx.x == 0 && atomicAdd(x.done + 1, 1u) + 1u == gridDim.x) {  // last block out: this exchange is complete
    x.done[1] = 0;
    *x.epoch = epoch;
  }
  tl_end(tl, 4);
}

}  // namespace

int32_t launch_finish_pull(mrs_model* m, const PushDev& push) {
  mrs_engine* e = m->eng;
  MRS_CUDA(launch_pdl(item_finish_pull_kernel, dim3((m->n_slots_known + 255) / 256 + 1), dim3(256), 0, e->stream, m->n_items, m->slot_of_item,
                      m->n_slots_known, push, m->xbuf, m->idevavg, m->gavg, e->d_timeline));
  mark(e, "item_finish_pull");
  MRS_CUDA(cudaGetLastError());
  m->finished = true;
  m->host_valid = false;
  return MRS_OK;
}

void free_tiled_layout(const mrs_ratings* R) {
  auto& T = R->tl;
  dev_free(T.entry); dev_free(T.slice_off); dev_free(T.tile_slice_ptr); dev_free(T.tile_ubegin); dev_free(T.slot_item); dev_free(T.warp_part); dev_free(T.cta_desc);
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
  // ---- user tiles: contiguous user ranges of at most kTileUsers users, CUT BY COST so that every CTA of the pass gets the
  // same number of ratings: a tile takes the users that hold k quotas of n / #SMs ratings (k = as many as fit into
  // kTileUsers users) and is worked on by k CTAs.  Equal-sized tiles of 8,192 users got 7 or 8 CTAs each at ml-25m shape:
  // 161 against 144 rows per warp, and the pass as slow as its 7-CTA tiles (per-CTA stamps: 38 us against 35 us).
  // OPT-IN (MRS_TILES=1): on most boxes the pass is no faster than with equal-sized tiles and CTAs dealt out by cost
  // (68.2 against 68.0-68.8 us, tools/timeline.py), while the host-side cut below (one D2H copy of the row pointer, a
  // synchronisation and 162 k binary searches) costs 0.2 ms in every end-to-end step (tools/e2e_ab.py: 7.80 against 7.58 ms).
  std::vector<int32_t> h_ubegin, h_kctas;
  bool cost_tiles = getenv("MRS_TILES") && atoi(getenv("MRS_TILES")) == 1 && n > 0;
  if (cost_tiles) {
    const int32_t NU = R->n_users, C = e->sm_count;
    std::vector<int32_t> h_urow((size_t)NU + 1);
    MRS_CUDA(cudaMemcpyAsync(h_urow.data(), R->urow, sizeof(int32_t) * ((size_t)NU + 1), cudaMemcpyDeviceToHost, st));
    MRS_CUDA(cudaStreamSynchronize(st));
    auto skip_empty = [&](int32_t u) { while (u < NU && h_urow[(size_t)u + 1] == h_urow[(size_t)u]) ++u; return u; };
    // first user index whose rating prefix reaches m quotas of n / C (m = C: behind the last user)
    int32_t end_all = NU;  // one past the last user with ratings
    while (end_all > 0 && h_urow[(size_t)end_all] == h_urow[(size_t)end_all - 1]) --end_all;
    auto cut = [&](int64_t m) -> int32_t {
      if (m >= C) return end_all;
      const int64_t target = (n * m + C - 1) / C;
      return (int32_t)(std::lower_bound(h_urow.begin(), h_urow.end(), (int32_t)target) - h_urow.begin());
    };
    // fewest tiles any cut needs: greedily as many quotas per tile as fit into kTileUsers users
    int32_t nt_min = 0;
    {
      int32_t u = skip_empty(0);
      int64_t done = 0;
      while (u < NU && done < C) {
        int64_t k = 1;
        while (done + k < C && cut(done + k + 1) - u <= kTileCap) ++k;
        done += k;
        u = skip_empty(std::max(cut(done), u + 1));
        ++nt_min;
      }
    }
    // then the same number of CTAs per tile (+-1): tiles of very different sizes would not cost the same per rating
    bool ok = false;
    for (int32_t nt = std::max(1, nt_min); nt <= nt_min + 8 && !ok; ++nt) {
      const int32_t base = C / nt;
      int32_t extra = C % nt;
      if (base == 0) break;
      h_ubegin.clear(); h_kctas.clear();
      int32_t u = skip_empty(0);
      int64_t done = 0;
      bool fail = false;
      for (int32_t t = 0; t < nt && !fail; ++t) {
        if (u >= NU) { fail = true; break; }
        int32_t k = base + (extra > 0 ? 1 : 0);
        int32_t end = std::max(cut(done + k), u + 1);
        if (end - u > kTileCap && k > base && extra < nt - t) { k = base; end = std::max(cut(done + k), u + 1); }
        if (end - u > kTileCap) { fail = true; break; }
        if (k > base) --extra;
        h_ubegin.push_back(u);
        h_kctas.push_back(k);
        done += k;
        u = skip_empty(end);
      }
      ok = !fail && done == C && u >= NU;
    }
    if (ok) h_ubegin.push_back(NU); else cost_tiles = false;
  }
  if (!cost_tiles) {
    h_ubegin.clear(); h_kctas.clear();
    const int32_t nt = (R->n_users + kTileCap - 1) / kTileCap;
    for (int32_t t = 0; t <= nt; ++t) h_ubegin.push_back((int32_t)std::min<int64_t>((int64_t)t * kTileCap, R->n_users));
  }
  const int32_t NT = (int32_t)h_ubegin.size() - 1;
  MRS_REQUIRE(NT < 65536, MRS_ERR_UNSUPPORTED, "too many user tiles (%d)", NT);
  T.n_tiles = NT;
  const int block = 256;
  const int grid = grid_for(n, block, e->sm_count);
  std::vector<int32_t> h_tile_slice((size_t)NT + 1, 0);
  MRS_TRY(dev_alloc(&T.tile_slice_ptr, (size_t)NT + 1));
  MRS_TRY(dev_alloc(&T.tile_ubegin, (size_t)NT + 1));
  MRS_CUDA(cudaMemcpyAsync(T.tile_ubegin, h_ubegin.data(), sizeof(int32_t) * ((size_t)NT + 1), cudaMemcpyHostToDevice, st));
  if (n == 0) {
    MRS_CUDA(cudaMemsetAsync(T.tile_slice_ptr, 0, sizeof(int32_t) * ((size_t)NT + 1), st));
    MRS_TRY(dev_alloc(&T.slice_off, 1));
    MRS_CUDA(cudaMemsetAsync(T.slice_off, 0, sizeof(int32_t), st));
    MRS_TRY(dev_alloc(&T.entry, 1));
    MRS_TRY(dev_alloc(&T.slot_item, 1));
    MRS_CUDA(cudaStreamSynchronize(st));
    T.built = true;
    return MRS_OK;
  }
  // ---- (tile, item, user) order: stable radix sort of the CSC positions on the tile id only
  uint16_t *tk_in = nullptr, *tk_out = nullptr;
  int32_t *pos_in = nullptr, *perm = nullptr, *item_of = nullptr, *head = nullptr, *seg_start = nullptr, *flag = nullptr, *uid = nullptr;
  MRS_TRY(dev_alloc(&tk_in, (size_t)n)); MRS_TRY(dev_alloc(&tk_out, (size_t)n));
  MRS_TRY(dev_alloc(&pos_in, (size_t)n)); MRS_TRY(dev_alloc(&perm, (size_t)n));
  MRS_TRY(dev_alloc(&item_of, (size_t)n)); MRS_TRY(dev_alloc(&head, (size_t)n));
  MRS_TRY(dev_alloc(&seg_start, (size_t)n)); MRS_TRY(dev_alloc(&flag, (size_t)n)); MRS_TRY(dev_alloc(&uid, (size_t)n + 1));
  uint16_t* utile = nullptr;
  MRS_TRY(dev_alloc(&utile, (size_t)R->n_users + 1));
  user_tile_kernel<<<NT, 256, 0, st>>>(T.tile_ubegin, NT, utile);
  item_expand_kernel<<<std::max(1, std::min((NI + 7) / 8, e->sm_count * 32)), 256, 0, st>>>(R->icolp, NI, item_of);
  tile_keys_kernel<<<grid, block, 0, st>>>(R->irow, utile, n, tk_in, pos_in);
  int tbits = 1;
  while ((1 << tbits) < NT) ++tbits;
  size_t tmp = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, tmp, tk_in, tk_out, pos_in, perm, (int)n, 0, tbits, st);
  MRS_TRY(ensure_scratch(e, tmp));
  cub::DeviceRadixSort::SortPairs(e->scratch, tmp, tk_in, tk_out, pos_in, perm, (int)n, 0, tbits, st);
  // ---- units
  seg_head_kernel<<<grid, block, 0, st>>>(perm, item_of, R->irow, utile, n, head);
  cub::DeviceScan::InclusiveScan(nullptr, tmp, head, seg_start, MaxOp(), (int)n, st);
  MRS_TRY(ensure_scratch(e, tmp));
  cub::DeviceScan::InclusiveScan(e->scratch, tmp, head, seg_start, MaxOp(), (int)n, st);
  unit_flag_kernel<<<grid, block, 0, st>>>(seg_start, n, flag);
  cub::DeviceScan::ExclusiveSum(nullptr, tmp, flag, uid, (int)n, st);
  MRS_TRY(ensure_scratch(e, tmp));
  cub::DeviceScan::ExclusiveSum(e->scratch, tmp, flag, uid, (int)n, st);
  int32_t last_uid = 0, last_flag = 0;
  MRS_CUDA(cudaMemcpyAsync(&last_uid, uid + (n - 1), sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  MRS_CUDA(cudaMemcpyAsync(&last_flag, flag + (n - 1), sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  MRS_CUDA(cudaStreamSynchronize(st));
  const int32_t NUN = last_uid + last_flag;
  T.n_units = NUN;
  int32_t *unit_begin = nullptr, *unit_item = nullptr, *unit_tile = nullptr, *unit_len = nullptr, *ids = nullptr, *sorted_id = nullptr;
  int32_t *tile_count = nullptr, *unit_slot = nullptr, *slice_width = nullptr, *d_tile_unit_ptr = nullptr;
  uint32_t *skey = nullptr, *skey_out = nullptr;
  MRS_TRY(dev_alloc(&unit_begin, (size_t)NUN)); MRS_TRY(dev_alloc(&unit_item, (size_t)NUN)); MRS_TRY(dev_alloc(&unit_tile, (size_t)NUN));
  MRS_TRY(dev_alloc(&unit_len, (size_t)NUN)); MRS_TRY(dev_alloc(&ids, (size_t)NUN)); MRS_TRY(dev_alloc(&sorted_id, (size_t)NUN));
  MRS_TRY(dev_alloc(&skey, (size_t)NUN)); MRS_TRY(dev_alloc(&skey_out, (size_t)NUN));
  MRS_TRY(dev_alloc(&tile_count, (size_t)NT + 1)); MRS_TRY(dev_alloc(&unit_slot, (size_t)NUN));
  MRS_TRY(dev_alloc(&d_tile_unit_ptr, (size_t)NT + 1));
  MRS_CUDA(cudaMemsetAsync(tile_count, 0xff, sizeof(int32_t) * ((size_t)NT + 1), st));  // -1: tile without units
  unit_scatter_kernel<<<grid, block, 0, st>>>(flag, uid, perm, item_of, R->irow, utile, n, unit_begin, unit_item, unit_tile);
  const int ugrid = (NUN + block - 1) / block;
  unit_len_kernel<<<ugrid, block, 0, st>>>(unit_begin, unit_tile, NUN, n, unit_len, skey, ids, tile_count);
  // ---- sort units by (tile, length desc); stable => canonical order among equal lengths
  cub::DeviceRadixSort::SortPairs(nullptr, tmp, skey, skey_out, ids, sorted_id, NUN, 0, kUnitBits + tbits, st);
  MRS_TRY(ensure_scratch(e, tmp));
  cub::DeviceRadixSort::SortPairs(e->scratch, tmp, skey, skey_out, ids, sorted_id, NUN, 0, kUnitBits + tbits, st);
  std::vector<int32_t> h_count((size_t)NT + 1, 0), h_unit_ptr((size_t)NT + 1, 0);
  MRS_CUDA(cudaMemcpyAsync(h_count.data(), tile_count, sizeof(int32_t) * (size_t)NT, cudaMemcpyDeviceToHost, st));
  MRS_CUDA(cudaStreamSynchronize(st));
  h_count[NT] = NUN;  // h_count holds the first unit of every tile; tiles without units take the next tile's first
  for (int32_t t = NT - 1; t >= 0; --t)
    if (h_count[t] < 0) h_count[t] = h_count[t + 1];
  for (int32_t t = 0; t < NT; ++t) {
    const int32_t cnt = h_count[t + 1] - h_count[t];
    h_unit_ptr[t + 1] = h_unit_ptr[t] + cnt;
    h_tile_slice[t + 1] = h_tile_slice[t] + (cnt + 31) / 32;
  }
  const int32_t NS = h_tile_slice[NT];
  T.n_slices = NS;
  MRS_CUDA(cudaMemcpyAsync(d_tile_unit_ptr, h_unit_ptr.data(), sizeof(int32_t) * ((size_t)NT + 1), cudaMemcpyHostToDevice, st));
  MRS_CUDA(cudaMemcpyAsync(T.tile_slice_ptr, h_tile_slice.data(), sizeof(int32_t) * ((size_t)NT + 1), cudaMemcpyHostToDevice, st));
  MRS_TRY(dev_alloc(&slice_width, (size_t)NS + 1));
  MRS_TRY(dev_alloc(&T.slice_off, (size_t)NS + 1));
  MRS_CUDA(cudaMemsetAsync(slice_width, 0, sizeof(int32_t) * ((size_t)NS + 1), st));
  slot_assign_kernel<<<ugrid, block, 0, st>>>(sorted_id, unit_tile, unit_len, d_tile_unit_ptr, T.tile_slice_ptr, NUN, unit_slot, slice_width);
  cub::DeviceScan::ExclusiveSum(nullptr, tmp, slice_width, T.slice_off, NS + 1, st);
  MRS_TRY(ensure_scratch(e, tmp));
  cub::DeviceScan::ExclusiveSum(e->scratch, tmp, slice_width, T.slice_off, NS + 1, st);
  int32_t rows = 0;
  MRS_CUDA(cudaMemcpyAsync(&rows, T.slice_off + NS, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  MRS_CUDA(cudaStreamSynchronize(st));
  T.n_slots = (int64_t)rows * 32;
  MRS_TRY(dev_alloc(&T.entry, (size_t)T.n_slots));
  {  // entry form: the fp64 form needs codes <= kAlphaMaxCode and |10c - S|, |S - 2c| < 2^21 (MRS_FORM=0 forces the integer form)
    int32_t* d_mx = nullptr;
    int32_t h_mx[2] = {0, 0};
    MRS_TRY(dev_alloc(&d_mx, 2));
    MRS_CUDA(cudaMemsetAsync(d_mx, 0, 2 * sizeof(int32_t), st));
    form_scan_kernel<<<grid, block, 0, st>>>((const uint8_t*)R->ival, n, R->urow, R->n_users, d_mx);
    MRS_CUDA(cudaMemcpyAsync(h_mx, d_mx, sizeof(h_mx), cudaMemcpyDeviceToHost, st));
    MRS_CUDA(cudaStreamSynchronize(st));
    dev_free(d_mx);
    const bool forced0 = getenv("MRS_FORM") && atoi(getenv("MRS_FORM")) == 0;
    T.form = (!forced0 && h_mx[0] <= kAlphaMaxCode && h_mx[1] < (1 << 17)) ? 1 : 0;
  }
  if (T.form == 1)
    fill_u32_kernel<<<grid_for(T.n_slots, block, e->sm_count), block, 0, st>>>(T.entry, T.n_slots, kAlphaPadding);
  else
    MRS_CUDA(cudaMemsetAsync(T.entry, 0, sizeof(uint32_t) * (size_t)T.n_slots, st));
  const bool reorder = getenv("MRS_REORDER") && atoi(getenv("MRS_REORDER"));
  if (!reorder)
    entry_fill_kernel<<<ugrid, block, 0, st>>>(unit_begin, unit_len, unit_slot, unit_tile, NUN, perm, R->irow, (const uint8_t*)R->ival, T.slice_off, T.entry,
                                               T.form, T.tile_ubegin);
  // ---- item of every slot (empty slots: -1)
  MRS_TRY(dev_alloc(&T.slot_item, (size_t)NS * 32));
  MRS_CUDA(cudaMemsetAsync(T.slot_item, 0xff, sizeof(int32_t) * (size_t)NS * 32, st));
  int32_t* slot_unit = nullptr;
  if (reorder) {
    MRS_TRY(dev_alloc(&slot_unit, (size_t)NS * 32));
    MRS_CUDA(cudaMemsetAsync(slot_unit, 0xff, sizeof(int32_t) * (size_t)NS * 32, st));
  }
  slot_item_kernel<<<ugrid, block, 0, st>>>(unit_slot, unit_item, NUN, T.slot_item, slot_unit);
  if (reorder) {
    entry_fill_ordered_kernel<<<(NS + 3) / 4, 128, 0, st>>>(slot_unit, unit_begin, unit_len, unit_tile, NS, perm, R->irow, (const uint8_t*)R->ival, T.slice_off,
                                                         T.entry, T.form, T.tile_ubegin);
    MRS_CUDA(cudaStreamSynchronize(st));
    dev_free(slot_unit);
  }
  // ---- static work partition of the item pass: CTAs dealt out to the tiles by cost, then slices to the warps of each CTA.
  // Laid down here, behind the layout's final synchronisation: the item pass reads warp_part in its prologue, before
  // griddepcontrol.wait, so the table must never be written by a kernel of the pass itself.
  {
    std::vector<int32_t> h_slice_off((size_t)NS + 1);
    MRS_CUDA(cudaMemcpyAsync(h_slice_off.data(), T.slice_off, sizeof(int32_t) * ((size_t)NS + 1), cudaMemcpyDeviceToHost, st));
    MRS_CUDA(cudaStreamSynchronize(st));
    std::vector<int64_t> cost((size_t)NT, 0);
    for (int32_t t = 0; t < NT; ++t) {
      const int32_t s0 = h_tile_slice[t], s1 = h_tile_slice[t + 1];
      cost[(size_t)t] = (int64_t)(h_slice_off[(size_t)s1] - h_slice_off[(size_t)s0]) + (int64_t)kSliceCost * (s1 - s0);
    }
    std::vector<int3> desc;
    if (cost_tiles) {  // every tile was cut to hold k quotas: k CTAs
      for (int32_t t = 0; t < NT; ++t)
        for (int32_t k = 0; k < h_kctas[(size_t)t]; ++k)
          if (cost[(size_t)t] > 0) desc.push_back(make_int3(t, k, h_kctas[(size_t)t]));
    } else {
      desc = deal_ctas(cost, e->sm_count);
    }
    T.n_ctas = (int32_t)desc.size();
    MRS_TRY(dev_alloc(&T.cta_desc, std::max<size_t>(1, desc.size())));
    MRS_TRY(dev_alloc(&T.warp_part, std::max<size_t>(1, desc.size()) * (kTiledThreads / 32)));
    if (T.n_ctas > 0) {
      MRS_CUDA(cudaMemcpyAsync(T.cta_desc, desc.data(), sizeof(int3) * desc.size(), cudaMemcpyHostToDevice, st));
      item_partition_kernel<<<T.n_ctas, kTiledThreads, 0, st>>>(T.slice_off, T.tile_slice_ptr, T.cta_desc, T.warp_part);
      MRS_CUDA(cudaStreamSynchronize(st));  // `desc` (pageable host memory) must outlive the copy
    }
  }
  count_launch(23);
  MRS_CUDA(cudaGetLastError());
  MRS_CUDA(cudaStreamSynchronize(st));
  for (void* p : {(void*)tk_in, (void*)tk_out, (void*)pos_in, (void*)perm, (void*)item_of, (void*)head, (void*)seg_start, (void*)flag,
                  (void*)uid, (void*)unit_begin, (void*)unit_item, (void*)unit_tile, (void*)unit_len, (void*)ids, (void*)sorted_id,
                  (void*)skey, (void*)skey_out, (void*)tile_count, (void*)unit_slot, (void*)slice_width,
                  (void*)d_tile_unit_ptr, (void*)utile})
    dev_free(p);
  T.built = true;
  return MRS_OK;
}

int32_t launch_item_tiled(mrs_engine* e, const mrs_ratings* R, mrs_model* m, bool fused, const PushDev* push, bool no_finalize) {
  const auto& T = R->tl;
  cudaStream_t st = e->stream;
  if (!(e->smem_attr_done & 1u)) {
    MRS_CUDA(cudaFuncSetAttribute(item_tiled_kernel<true, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTiledSmem));
    MRS_CUDA(cudaFuncSetAttribute(item_tiled_kernel<false, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTiledSmem));
    MRS_CUDA(cudaFuncSetAttribute(item_tiled_kernel<true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTiledSmem));
    MRS_CUDA(cudaFuncSetAttribute(item_tiled_kernel<false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTiledSmem));
    e->smem_attr_done |= 1u;
  }
  // one CTA of 1024 threads per SM: the grid (T.n_ctas <= SM count unless there are more busy tiles than SMs) is one wave
  const dim3 grid2(T.n_ctas), block2(kTiledThreads);
  if (T.n_ctas > 0) {
#define MRS_LAUNCH_K2(SUM, FORM)                                                                                                               \
  MRS_CUDA(launch_pdl(item_tiled_kernel<SUM, FORM>, grid2, block2, kTiledSmem, st, T.entry, T.slice_off, T.warp_part, T.cta_desc, m->usum, R->urow, \
                      R->n_users, T.slot_item, m->uavg, m->xdev_fix, m->xcode_sum, e->d_timeline, m->counters + 4, R->n_items, T.tile_ubegin, m->k1_part, \
                      m->flag_sync ? m->k1_blocks : 0))
    if (m->want_item_avg) {
      if (T.form == 1) MRS_LAUNCH_K2(true, 1); else MRS_LAUNCH_K2(true, 0);
    } else {
      if (T.form == 1) MRS_LAUNCH_K2(false, 1); else MRS_LAUNCH_K2(false, 0);
    }
#undef MRS_LAUNCH_K2
  }
  mark(e, "item_tiled");
  if (no_finalize) return MRS_OK;  // mrs_fit_mae_async: the test pass finishes the fit itself
  if (push) {  // sharded run with the fused exchange: K2b delivers the partial sums to every rank itself
    MRS_CUDA(launch_pdl(item_tiled_push_kernel, dim3((m->n_slots_known + 255) / 256 + 1), dim3(256), 0, st, m->xdev_fix, R->icolp, m->k1_part,
                        (double)R->n, m->slot_of_item, m->n_slots_known, *push, m->usum, R->n_users, e->d_timeline, m->counters + 4, R->n_items));
    mark(e, "item_tiled_push");
    MRS_CUDA(cudaGetLastError());
    return MRS_OK;
  }
  MRS_CUDA(launch_pdl(item_tiled_finalize_kernel, dim3((R->n_items + 255) / 256), dim3(256), 0, st, m->xdev_fix, m->xcode_sum, R->icolp, R->n_items,
                      m->k1_part, m->k1_blocks, (double)R->n, m->xbuf, fused ? 1 : 0, m->idevavg, m->iavg, m->gavg, m->usum, R->n_users, e->d_timeline, m->counters + 4));
  mark(e, "item_tiled_finalize");
  MRS_CUDA(cudaGetLastError());
  return MRS_OK;
}

}  // namespace mrs
