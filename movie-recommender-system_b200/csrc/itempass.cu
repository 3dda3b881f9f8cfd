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

int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return (v && *v) ? atoi(v) : dflt;
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
                                 int32_t* __restrict__ slot_item, int32_t* __restrict__ slot_unit) {
  const int32_t id = blockIdx.x * blockDim.x + threadIdx.x;
  if (id < n_units) {
    slot_item[unit_slot[id]] = unit_item[id];
    slot_unit[unit_slot[id]] = id;
  }
}

// Entry placement that avoids shared-memory bank conflicts in the pass.  The pass gathers one 8-byte table slot per
// entry; a 64-bit shared-memory load is served 16 lanes at a time and conflict free only if those 16 lanes hit 16
// different 8-byte bank pairs -- for both tables the bank pair is (user id & 15).  With entries in (item, user) order the
// keys of a row are random: 3.2 wavefronts per half warp instead of 1 (ncu, round 2: 2.7 M of 4.7 M shared-memory
// wavefronts of the pass were conflicts).  The order of the entries INSIDE a unit is free (it only fixes the summation
// order), so one warp per slice deals them out position by position: every lane offers an entry whose key is still free
// in its half warp at this position (lowest lane wins a contested key, the others offer another key in the next round);
// a lane that has no such entry leaves the position to padding if it still has slack, else takes a conflict.
template <bool POP>
__global__ void __launch_bounds__(128) entry_fill_ordered_kernel(const int32_t* __restrict__ slot_unit, const int32_t* __restrict__ unit_begin,
                                                                const int32_t* __restrict__ unit_len, const int32_t* __restrict__ unit_tile,
                                                                int32_t n_slices, const int32_t* __restrict__ perm,
                                                                const int32_t* __restrict__ irow, const uint8_t* __restrict__ ival,
                                                                const int32_t* __restrict__ slice_off, int32_t code_min,
                                                                uint32_t* __restrict__ entry) {
  __shared__ uint32_t s_ent[4][32 * kUnitLen];   // entries of the lane's unit, grouped by key
  __shared__ uint8_t s_next[4][32][16];          // next unplaced entry of every key group
  __shared__ uint8_t s_end[4][32][16];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int32_t slice = blockIdx.x * 4 + w;
  if (slice >= n_slices) return;
  constexpr int32_t TU = POP ? kPopTileUsers : kRareTileUsers;
  const int32_t id = slot_unit[slice * 32 + lane];
  const int32_t len = id >= 0 ? unit_len[id] : 0;
  const int32_t b = id >= 0 ? unit_begin[id] : 0;
  const int32_t ubase = id >= 0 ? unit_tile[id] * TU : 0;
  uint32_t* ent = s_ent[w] + lane * kUnitLen;
  uint8_t* nxt = s_next[w][lane];
  uint8_t* end = s_end[w][lane];
  // counting sort of the unit's entries by key
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
    const uint32_t e = POP ? (uint32_t)(((int32_t)ival[p] - code_min) * kPopTileUsers + x) : (((uint32_t)ival[p] << 20) | ((uint32_t)x << 3));
    ent[nxt[x & 15]++] = e;
  }
  {  // rewind the group cursors
    int32_t run = 0;
#pragma unroll
    for (int k = 0; k < 16; ++k) { nxt[k] = (uint8_t)run; run += cnt[k]; }
  }
  __syncwarp();
  const int64_t row0 = slice_off[slice];
  const int32_t n_pos = (slice_off[slice + 1] - slice_off[slice]) * (POP ? 2 : 1);
  const int half = lane >> 4;
  int32_t remaining = len;
  uint16_t* entry16 = reinterpret_cast<uint16_t*>(entry);
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
      if (POP) entry16[((((row0 + (pos >> 1)) << 5) + lane) << 1) + (pos & 1)] = (uint16_t)e;
      else entry[((row0 + pos) << 5) + lane] = e;
    }
  }
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
constexpr size_t kPopTableBytes = ((size_t)kPopTileUsers * kMaxCodes + 2) * 8;   // + zero slot for padding entries (+1: bulk copies move 16-byte units)
constexpr size_t kRareTableBytes = ((size_t)kRareTileUsers + 2) * 8;             // + dummy user of padding entries
constexpr size_t kTableBytes = ((kPopTableBytes > kRareTableBytes ? kPopTableBytes : kRareTableBytes) + 127) / 128 * 128;
constexpr size_t kRingBytes = (size_t)kPassWarps * kStages * kRows * 128;
constexpr size_t kPassSmem = kTableBytes + kRingBytes + (size_t)kPassWarps * kStages * 8 + 16;
static_assert(kPassSmem <= 232448, "item pass: shared memory budget of one CTA");

// One segment of a CTA's work: a run of slices of ONE tile of one part (the table in shared memory belongs to a tile)
// (int2: x = tile of the part, bit 30 set = rare part; y = first of the 32 per-warp slice ranges of the segment in warp_part)
struct PassSeg {
  int32_t tile;
  int32_t wp;
};
static_assert(sizeof(PassSeg) == sizeof(int2), "PassSeg is stored as int2");

struct PassArgs {
  const uint32_t *entry_pop, *entry_rare;
  const int32_t *slice_off_pop, *slice_off_rare;
  const int32_t *slot_item_pop, *slot_item_rare;
  const int32_t* cta_seg_ptr;    // [n_ctas+1] segments of every CTA
  const PassSeg* seg;
  const int2* warp_part;         // [n_segments * 32] slices [x, y) of every warp, in the slice numbering of the part
  int32_t n_codes;
  const double* pop_img;         // table images written by K1b
  const uint2* rare_img;
  long long* xdev_fix;
  unsigned long long* tl;        // diagnostics (MRS_TIMELINE=1)
  long long* dbg;                // diagnostics (MRS_PASS_DEBUG=1): 16 clock64 stamps per CTA, else NULL
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

// state of a warp's slice walk inside a segment
struct Walk {
  int32_t cur, s_hi, end1, end2, item1, item2;
};

__device__ __forceinline__ void walk_begin(Walk& wk, const int2 wp, const int32_t* __restrict__ slice_off, const int32_t* __restrict__ slot_item,
                                           int lane) {
  wk.cur = wp.x; wk.s_hi = wp.y;
  wk.end1 = (wk.cur < wk.s_hi) ? __ldg(slice_off + wk.cur + 1) : 0x7fffffff;        // end row of the current slice
  wk.end2 = (wk.cur + 1 < wk.s_hi) ? __ldg(slice_off + wk.cur + 2) : 0x7fffffff;    // ... of the next one (prefetched)
  wk.item1 = (wk.cur < wk.s_hi) ? __ldg(slot_item + wk.cur * 32 + lane) : -1;       // item of this lane's unit in the current slice
  wk.item2 = (wk.cur + 1 < wk.s_hi) ? __ldg(slot_item + (wk.cur + 1) * 32 + lane) : -1;
}
// the slice ended with the previous row: hand this lane's unit sum over and move to the next slice
__device__ __forceinline__ void walk_next(Walk& wk, double& acc, const int32_t* __restrict__ slice_off, const int32_t* __restrict__ slot_item,
                                          int lane, long long* __restrict__ xdev_fix) {
  hand_over(wk.item1, acc, xdev_fix);
  acc = 0.0;
  ++wk.cur;
  wk.end1 = wk.end2; wk.item1 = wk.item2;
  wk.end2 = (wk.cur + 1 < wk.s_hi) ? __ldg(slice_off + wk.cur + 2) : 0x7fffffff;
  wk.item2 = (wk.cur + 1 < wk.s_hi) ? __ldg(slot_item + (wk.cur + 1) * 32 + lane) : -1;
}

__device__ __forceinline__ void ring_request(uint32_t* ring, uint64_t* bar, const uint32_t* __restrict__ entry, int32_t cc, int32_t rr, int32_t r_end) {
  const int st = cc % kStages;
  const uint32_t bytes = (uint32_t)min(kRows, r_end - rr) * 128u;
  tma::mbar_arrive_expect_tx(bar + st, bytes);
  tma::bulk_g2s(ring + st * kRows * 32, entry + ((int64_t)rr << 5), bytes, bar + st);
}

__global__ void __launch_bounds__(kPassThreads, 1) item_pass_kernel(const PassArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* s_tab = reinterpret_cast<double*>(smem_raw);                       // popular: dev[code][user]; rare: (S, c) pairs
  uint32_t* s_ring = reinterpret_cast<uint32_t*>(smem_raw + kTableBytes);    // [warps][kStages][kRows*32]
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(smem_raw + kTableBytes + kRingBytes);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  tl_begin(a.tl, 2);
  uint64_t* bar = s_bar + wid * kStages;
  uint64_t* tabbar = s_bar + kPassWarps * kStages;   // completion of the table's bulk copy
  uint32_t* ring = s_ring + (size_t)wid * kStages * kRows * 32;
  if (lane == 0) {
#pragma unroll
    for (int st = 0; st < kStages; ++st) tma::mbar_init(bar + st, 1);
    if (wid == 0) tma::mbar_init(tabbar, 1);
    tma::fence_barrier_init();
  }
  __syncthreads();
  const int32_t seg_lo = __ldg(a.cta_seg_ptr + blockIdx.x), seg_hi = __ldg(a.cta_seg_ptr + blockIdx.x + 1);
  long long* dbg = a.dbg ? a.dbg + (size_t)blockIdx.x * 16 : nullptr;
  int dbg_n = 0;
#define MRS_STAMP() do { if (dbg && threadIdx.x == 0 && dbg_n < 13) dbg[dbg_n++] = clock64(); } while (0)
  MRS_STAMP();
  if (dbg && threadIdx.x == 0) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); dbg[13] = (long long)t; }
  int32_t cc = 0;  // chunks this warp has consumed so far (ring stage and barrier phase follow from it)

  // the rows of the first segment start travelling before K1 has finished (they only depend on the layout)
  PassSeg sg = (seg_lo < seg_hi) ? a.seg[seg_lo] : PassSeg{0, 0};
  int2 wp = (seg_lo < seg_hi) ? __ldg(a.warp_part + sg.wp + wid) : make_int2(0, 0);
  {
    const bool rare = (sg.tile >> 30) & 1;
    const uint32_t* entry = rare ? a.entry_rare : a.entry_pop;
    const int32_t* slice_off = rare ? a.slice_off_rare : a.slice_off_pop;
    if (wp.x < wp.y && lane == 0) {
      const int32_t r0 = __ldg(slice_off + wp.x), r_end = __ldg(slice_off + wp.y);
#pragma unroll
      for (int k = 0; k < kStages; ++k)
        if (r0 + k * kRows < r_end) ring_request(ring, bar, entry, k, r0 + k * kRows, r_end);
    }
  }
  // ... and the rows of the later segments are asked into the L2 now: a warp's ring holds 2 KB, far too little to cover
  // HBM latency at the rate the rows are consumed (stamps, round 2: 1.5 us per 1 KB chunk), but its whole run of a
  // segment is contiguous and only a few KB -- one bulk prefetch per (warp, segment), in flight while K1 finishes and
  // the first table is built
  if (lane == 0) {
    for (int32_t si = seg_lo; si < seg_hi; ++si) {
      const PassSeg s2 = a.seg[si];
      const int2 w2 = __ldg(a.warp_part + s2.wp + wid);
      if (w2.x >= w2.y) continue;
      const bool rare = (s2.tile >> 30) & 1;
      const int32_t* so = rare ? a.slice_off_rare : a.slice_off_pop;
      const int32_t ra = __ldg(so + w2.x), rb = __ldg(so + w2.y);
      const int32_t skip = (si == seg_lo) ? kStages * kRows : 0;  // (already requested into shared memory)
      if (rb - ra > skip) tma::prefetch_l2((rare ? a.entry_rare : a.entry_pop) + ((int64_t)(ra + skip) << 5), (uint32_t)(rb - ra - skip) * 128u);
    }
  }
  pdl_trigger();  // K2b may be scheduled as SMs free up
  pdl_wait();     // K1b's table images are complete from here on
  MRS_STAMP();
  MRS_STAMP();
  for (int32_t si = seg_lo; si < seg_hi; ++si) {
    if (si > seg_lo) {
      sg = a.seg[si];
      wp = __ldg(a.warp_part + sg.wp + wid);
      __syncthreads();  // every warp is done with the previous segment's table
    }
    const bool rare = (sg.tile >> 30) & 1;
    const int32_t tile = sg.tile & 0x3fffffff;
    const uint32_t* __restrict__ entry = rare ? a.entry_rare : a.entry_pop;
    const int32_t* __restrict__ slice_off = rare ? a.slice_off_rare : a.slice_off_pop;
    const int32_t* __restrict__ slot_item = rare ? a.slot_item_rare : a.slot_item_pop;
    const int32_t r0 = (wp.x < wp.y) ? __ldg(slice_off + wp.x) : 0;
    const int32_t r_end = (wp.x < wp.y) ? __ldg(slice_off + wp.y) : 0;
    const int32_t n_chunks = (r_end - r0 + kRows - 1) / kRows;
    if (si > seg_lo && lane == 0) {  // (the first segment's rows were requested in the prologue)
#pragma unroll
      for (int k = 0; k < kStages; ++k)
        if (k < n_chunks) ring_request(ring, bar, entry, cc + k, r0 + k * kRows, r_end);
    }
    Walk wk;
    walk_begin(wk, wp, slice_off, slot_item, lane);
    double acc = 0.0;

    // ---- the tile's table: one image written by K1b, pulled into shared memory by bulk copies (4 in flight)
    {
      const uint32_t bytes = rare ? (uint32_t)kRareTableBytes : (uint32_t)((size_t)kPopTileUsers * a.n_codes + 2) * 8u;
      if (threadIdx.x == 0) {
        const unsigned char* img = rare ? reinterpret_cast<const unsigned char*>(a.rare_img) + (size_t)tile * kRareTableBytes
                                        : reinterpret_cast<const unsigned char*>(a.pop_img) + (size_t)tile * bytes;
        tma::mbar_arrive_expect_tx(tabbar, bytes);
        const uint32_t piece = ((bytes / 4) + 15u) & ~15u;
        for (uint32_t off = 0; off < bytes; off += piece) tma::bulk_g2s(smem_raw + off, img + off, min(piece, bytes - off), tabbar);
      }
      tma::mbar_wait(tabbar, (uint32_t)(si - seg_lo) & 1u);
    }
    MRS_STAMP();

    if (!rare) {
      const int32_t n_slots = kPopTileUsers * a.n_codes;
      const uint32_t pad = (uint32_t)n_slots | ((uint32_t)n_slots << 16);
      for (int32_t c = 0; c < n_chunks; ++c, ++cc) {
        const int st = cc % kStages;
        const int32_t r = r0 + c * kRows;
        const int32_t nrows = min(kRows, r_end - r);
        tma::mbar_wait(bar + st, (uint32_t)(cc / kStages) & 1u);
        const uint32_t* rp = ring + st * kRows * 32 + lane;  // conflict free: lane l reads word l of a row
        uint32_t w[kRows];
        if (nrows == kRows) {
#pragma unroll
          for (int k = 0; k < kRows; ++k) w[k] = rp[k * 32];
        } else {
#pragma unroll
          for (int k = 0; k < kRows; ++k) w[k] = (k < nrows) ? rp[k * 32] : pad;
        }
        __syncwarp();
        if (lane == 0 && c + kStages < n_chunks)  // the stage is free again: request the chunk kStages ahead
          ring_request(ring, bar, entry, cc + kStages, r + kStages * kRows, r_end);
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
            if (r + h + k == wk.end1) walk_next(wk, acc, slice_off, slot_item, lane, a.xdev_fix);
            acc += d0[k];
            acc += d1[k];
          }
        }
      }
    } else {
      constexpr uint32_t kPadEntry = (uint32_t)kRareTileUsers << 3;

      for (int32_t c = 0; c < n_chunks; ++c, ++cc) {
        const int st = cc % kStages;
        const int32_t r = r0 + c * kRows;
        const int32_t nrows = min(kRows, r_end - r);
        tma::mbar_wait(bar + st, (uint32_t)(cc / kStages) & 1u);
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
        if (lane == 0 && c + kStages < n_chunks) ring_request(ring, bar, entry, cc + kStages, r + kStages * kRows, r_end);
        double dv[kRows];
#pragma unroll
        for (int k = 0; k < kRows; ++k) {  // heavy part: no branches, 8 independent chains
          const uint2 sc = *reinterpret_cast<const uint2*>(smem_raw + (ev[k] & 0xffff8u));
          dv[k] = dev_from_counts(sc.x, sc.y, ev[k] >> 20);
        }
#pragma unroll
        for (int k = 0; k < kRows; ++k) {
          if (r + k == wk.end1) walk_next(wk, acc, slice_off, slot_item, lane, a.xdev_fix);
          acc += dv[k];
        }
      }
    }
    if (wk.cur < wk.s_hi) hand_over(wk.item1, acc, a.xdev_fix);  // last slice of the range
    MRS_STAMP();  // (warp 0's own end of the segment)
  }
  if (a.tl) { __syncthreads(); tl_end(a.tl, 2); }
  if (dbg) {
    __syncthreads();
    MRS_STAMP();
    if (threadIdx.x == 0) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); dbg[14] = (long long)t; dbg[15] = dbg_n; }
  }
#undef MRS_STAMP
}

// K1b: per user, everything the item pass and the test pass need from K1's code sums: the average (P:113 / P:274: exact
// sum, one correctly rounded division, P:18; -1.0 = no ratings, the reference's own sentinel P:222), the (code sum, count)
// pair in the rare-part table image and the column of deviations dev[code] in the popular-part table image (two
// reciprocals per user, one multiplication per code).  The images have the shared-memory layout of the item pass, so a
// CTA of the pass pulls a tile's table in with one bulk copy instead of building it (stamps, round 2: building cost
// 3-5 us per CTA and segment, a third of the pass).
__global__ void __launch_bounds__(256) user_table_kernel(uint32_t* __restrict__ usum, const int32_t* __restrict__ urow, int32_t u_lo,
                                                        int32_t u_hi, int32_t code_min, int32_t nc, double* __restrict__ uavg,
                                                        double* __restrict__ pop_img, uint2* __restrict__ rare_img,
                                                        unsigned long long* __restrict__ tl) {
  tl_begin(tl, 1);
  pdl_trigger();
  pdl_wait();  // K1's sums are complete
  const int32_t u = u_lo + blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= u_hi) { tl_end(tl, 1); return; }
  const int32_t S = (int32_t)usum[u];
  usum[u] = 0;  // re-armed for K1's integer atomics of the next pass (no memset node in front of every pass)
  const int32_t cnt = __ldg(urow + u + 1) - __ldg(urow + u);
  uavg[u] = cnt ? (0.5 * (double)S) / (double)cnt : -1.0;
  rare_img[(size_t)(u / kRareTileUsers) * (kRareTileUsers + 2) + (u % kRareTileUsers)] = make_uint2((uint32_t)S, (uint32_t)cnt);
  if (pop_img) {
    double* col = pop_img + (size_t)(u / kPopTileUsers) * ((size_t)kPopTileUsers * nc + 2) + (u % kPopTileUsers);
    const double inv_hi = rcp_small_int(int_to_double(10 * cnt - S));  // 1 / (5 - avg) up to the common factor 2c
    const double inv_lo = rcp_small_int(int_to_double(S - 2 * cnt));   // 1 / (avg - 1)
    for (int32_t j = 0; j < nc; ++j) {
      const int32_t N = cnt * (code_min + j) - S;
      double dev = int_to_double(N) * (N > 0 ? inv_hi : inv_lo);
      if (N == 0 || cnt == 0) dev = 0.0;  // r == avg: the reference's 0/1 (also keeps 0 * inf out when avg is exactly 1 or 5)
      col[(size_t)j * kPopTileUsers] = dev;
    }
  }
  tl_end(tl, 1);
}

// one-time initialisation of a model's table images: zeros, the dummy user (S = 0, c = 1: deviation 0 for code 0) of every
// rare tile, and "no ratings" for every user average (users outside [user_lo, user_hi) never get anything else)
__global__ void table_init_kernel(uint2* __restrict__ rare_img, int32_t n_rare_tiles, double* __restrict__ uavg, int64_t n_uavg) {
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n_rare_tiles) rare_img[(size_t)i * (kRareTileUsers + 2) + kRareTileUsers] = make_uint2(0u, 1u);
  for (int64_t k = i; k < n_uavg; k += (int64_t)gridDim.x * blockDim.x) uavg[k] = -1.0;
}

// K2b: per item, integer accumulators -> exchange buffer (and re-arm them for the next pass); optionally finish the fit
__global__ void __launch_bounds__(256) item_tiled_finalize_kernel(long long* __restrict__ xdev_fix, uint32_t* __restrict__ xcode_sum,
                                                                 const int32_t* __restrict__ icolp, int32_t n_items,
                                                                 unsigned long long* __restrict__ k1_part, double n_total,
                                                                 double* __restrict__ xbuf, int fused, double* __restrict__ idevavg,
                                                                 double* __restrict__ iavg, double* __restrict__ gavg,
                                                                 unsigned long long* __restrict__ tl) {
  tl_begin(tl, 3);
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
  if (i >= n_items) { tl_end(tl, 3); return; }
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
  tl_end(tl, 3);
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
  // ---- item and unit of every slot (empty slots: -1)
  int32_t* slot_unit = nullptr;
  MRS_TRY(dev_alloc(&P.slot_item, (size_t)NS * 32));
  MRS_TRY(dev_alloc(&slot_unit, (size_t)NS * 32));
  MRS_CUDA(cudaMemsetAsync(P.slot_item, 0xff, sizeof(int32_t) * (size_t)NS * 32, st));
  MRS_CUDA(cudaMemsetAsync(slot_unit, 0xff, sizeof(int32_t) * (size_t)NS * 32, st));
  slot_item_kernel<<<ugrid, block, 0, st>>>(unit_slot, unit_item, NUN, P.slot_item, slot_unit);
  if (env_int("MRS_NO_REORDER", 0)) {  // entries in (item, user) order: the A/B switch for the bank-conflict-aware placement
    if (pop)
      entry_fill_pop_kernel<<<ugrid, block, 0, st>>>(unit_begin, unit_len, unit_slot, unit_tile, NUN, perm, R->irow, (const uint8_t*)R->ival,
                                                    P.slice_off, code_min, n_codes, reinterpret_cast<uint16_t*>(P.entry));
    else
      entry_fill_rare_kernel<<<ugrid, block, 0, st>>>(unit_begin, unit_len, unit_slot, unit_tile, NUN, perm, R->irow, (const uint8_t*)R->ival,
                                                     P.slice_off, P.entry);
  } else if (pop) {
    entry_fill_ordered_kernel<true><<<(NS + 3) / 4, 128, 0, st>>>(slot_unit, unit_begin, unit_len, unit_tile, NS, perm, R->irow, (const uint8_t*)R->ival,
                                                                 P.slice_off, code_min, P.entry);
  } else {
    entry_fill_ordered_kernel<false><<<(NS + 3) / 4, 128, 0, st>>>(slot_unit, unit_begin, unit_len, unit_tile, NS, perm, R->irow, (const uint8_t*)R->ival,
                                                                  P.slice_off, code_min, P.entry);
  }
  count_launch(18);
  MRS_CUDA(cudaGetLastError());
  MRS_CUDA(cudaStreamSynchronize(st));
  for (void* p : {(void*)tk_in, (void*)tk_out, (void*)pos_in, (void*)perm, (void*)head, (void*)seg_start, (void*)flag, (void*)uid,
                  (void*)unit_begin, (void*)unit_item, (void*)unit_tile, (void*)unit_len, (void*)ids, (void*)sorted_id, (void*)skey,
                  (void*)skey_out, (void*)tile_first, (void*)unit_slot, (void*)slice_rows, (void*)d_tile_unit_ptr, (void*)d_tile_slice, (void*)slot_unit})
    dev_free(p);
  return MRS_OK;
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

long long* g_pass_dbg = nullptr;

void free_tiled_layout(const mrs_ratings* R) {
  auto& T = R->tl;
  free_part(T.pop); free_part(T.rare);
  dev_free(T.cta_seg_ptr); dev_free(T.seg); dev_free(T.warp_part); dev_free(T.ival16); dev_free(T.vec_col);
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
  if (n > 0) {  // the sorted COO starts with the first user that has ratings and ends with the last
    MRS_CUDA(cudaMemcpyAsync(&T.user_lo, R->coo_u, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    MRS_CUDA(cudaMemcpyAsync(&T.user_hi, R->coo_u + (n - 1), sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    MRS_CUDA(cudaStreamSynchronize(st));
    T.user_hi += 1;
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

  // ---- static work partition.  The slices of both parts form one sequence (popular tiles, then rare tiles); every CTA
  // takes a contiguous run of it, cut into segments at tile boundaries (the table in shared memory belongs to a tile),
  // each segment split over the 32 warps.  Cost of a run, in units of one 128-byte row (7 ns per CTA for both kinds of
  // rows -- stamps, round 2): rows + slice_cost per slice (handing a slice over) + table_cost per segment (pulling a
  // tile's table into shared memory: ~3 us).  The runs are made equal in cost by bisection on the cost per CTA, so most
  // CTAs work on ONE tile and load one table.
  const int32_t n_ctas = (n > 0) ? e->sm_count : 0;
  const int32_t slice_cost[2] = {std::max(1, env_int("MRS_POP_SLICE_COST", 3)), std::max(1, env_int("MRS_RARE_SLICE_COST", 2))};
  const int64_t table_cost = env_int("MRS_TABLE_COST", 400);
  std::vector<std::vector<int2>> cta_segs((size_t)std::max(n_ctas, 1));  // per CTA: (tile | kind, first slice); end slices in a parallel vector
  std::vector<std::vector<int32_t>> cta_seg_end((size_t)std::max(n_ctas, 1));
  if (n_ctas > 0) {
    struct Piece { int part, tile, s0, s1; };   // the non-empty tiles in sequence order
    std::vector<Piece> pieces;
    for (int part = 0; part < 2; ++part) {
      const mrs_ratings::ell_part& P = part == 0 ? T.pop : T.rare;
      for (int32_t t = 0; t < P.n_tiles && P.n_slices > 0; ++t)
        if (P.h_tile_slice[(size_t)t + 1] > P.h_tile_slice[(size_t)t]) pieces.push_back({part, t, P.h_tile_slice[(size_t)t], P.h_tile_slice[(size_t)t + 1]});
    }
    auto slice_c = [&](const Piece& pc, int32_t s_) -> int64_t {
      const mrs_ratings::ell_part& P = pc.part == 0 ? T.pop : T.rare;
      return (int64_t)(P.h_slice_off[(size_t)s_ + 1] - P.h_slice_off[(size_t)s_]) + slice_cost[pc.part];
    };
    // greedy sweep for a cost cap: returns the number of CTAs used (and the cuts when `emit`)
    auto sweep = [&](int64_t cap, bool emit) -> int64_t {
      int64_t used = 0, acc = 0;
      bool open = false;
      for (const Piece& pc : pieces) {
        int32_t s0 = pc.s0;
        while (s0 < pc.s1) {
          if (!open) { ++used; acc = 0; open = true; }
          acc += table_cost;
          int32_t s1 = s0;
          while (s1 < pc.s1 && (acc + slice_c(pc, s1) <= cap || s1 == s0)) { acc += slice_c(pc, s1); ++s1; }
          if (emit) {
            const size_t b_ = (size_t)std::min<int64_t>(used - 1, n_ctas - 1);
            cta_segs[b_].push_back(make_int2(pc.tile | (pc.part << 30), s0));
            cta_seg_end[b_].push_back(s1);
          }
          if (s1 < pc.s1) open = false;                                  // the cap was reached inside the tile: next CTA
          else if (acc + table_cost + table_cost / 2 > cap) open = false;  // not worth opening another tile for a sliver
          s0 = s1;
        }
      }
      return used;
    };
    int64_t lo = 1, hi = 0;
    for (const Piece& pc : pieces) {
      const mrs_ratings::ell_part& P = pc.part == 0 ? T.pop : T.rare;
      hi += (int64_t)(P.h_slice_off[(size_t)pc.s1] - P.h_slice_off[(size_t)pc.s0]) + (int64_t)slice_cost[pc.part] * (pc.s1 - pc.s0) + table_cost;
    }
    hi = std::max<int64_t>(hi, 2);
    while (lo < hi) {  // smallest cap that needs at most n_ctas CTAs
      const int64_t mid = (lo + hi) / 2;
      if (sweep(mid, false) <= n_ctas) hi = mid; else lo = mid + 1;
    }
    sweep(lo, true);
  }
  std::vector<int32_t> seg_ptr((size_t)n_ctas + 1, 0);
  std::vector<int2> segs, wpart;
  for (int32_t b_ = 0; b_ < n_ctas; ++b_) {
    for (size_t k = 0; k < cta_segs[(size_t)b_].size(); ++k) {
      const int2 sgm = cta_segs[(size_t)b_][k];
      const int part = (sgm.x >> 30) & 1;
      const mrs_ratings::ell_part& P = part == 0 ? T.pop : T.rare;
      const int32_t sc = slice_cost[part];
      const int32_t s0 = sgm.y, s1 = cta_seg_end[(size_t)b_][k];
      const int64_t total = (int64_t)(P.h_slice_off[(size_t)s1] - P.h_slice_off[(size_t)s0]) + (int64_t)sc * (s1 - s0);
      segs.push_back(make_int2(sgm.x, (int32_t)wpart.size()));
      for (int w = 0; w < kPassWarps; ++w) {
        const int32_t lo = lower_bound_cost(P.h_slice_off, s0, s1, sc, total * w / kPassWarps);
        const int32_t hi = lower_bound_cost(P.h_slice_off, s0, s1, sc, total * (w + 1) / kPassWarps);
        wpart.push_back(make_int2(lo, hi));
      }
    }
    seg_ptr[(size_t)b_ + 1] = (int32_t)segs.size();
  }
  T.n_ctas = n_ctas;
  T.n_segs = (int32_t)segs.size();
  MRS_TRY(dev_alloc(&T.cta_seg_ptr, seg_ptr.size()));
  MRS_TRY(dev_alloc(&T.seg, std::max<size_t>(1, segs.size())));
  MRS_TRY(dev_alloc(&T.warp_part, std::max<size_t>(1, wpart.size())));
  MRS_CUDA(cudaMemcpyAsync(T.cta_seg_ptr, seg_ptr.data(), sizeof(int32_t) * seg_ptr.size(), cudaMemcpyHostToDevice, st));
  if (!segs.empty()) {
    MRS_CUDA(cudaMemcpyAsync(T.seg, segs.data(), sizeof(int2) * segs.size(), cudaMemcpyHostToDevice, st));
    MRS_CUDA(cudaMemcpyAsync(T.warp_part, wpart.data(), sizeof(int2) * wpart.size(), cudaMemcpyHostToDevice, st));
  }
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

// table images of a new model (sizes follow the layout) and their one-time initialisation
int32_t alloc_table_images(mrs_engine* e, const mrs_ratings* R, mrs_model* m, size_t uavg_len) {
  const auto& T = R->tl;
  const size_t pop_doubles = (size_t)std::max(T.pop.n_tiles, 1) * ((size_t)kPopTileUsers * T.n_codes + 2);
  const size_t rare_pairs = (size_t)std::max(T.rare.n_tiles, 1) * (kRareTileUsers + 2);
  MRS_TRY(dev_alloc(&m->pop_img, pop_doubles));
  MRS_TRY(dev_alloc(&m->rare_img, rare_pairs));
  MRS_CUDA(cudaMemsetAsync(m->pop_img, 0, pop_doubles * sizeof(double), e->stream));
  MRS_CUDA(cudaMemsetAsync(m->rare_img, 0, rare_pairs * sizeof(uint2), e->stream));
  table_init_kernel<<<std::max(1, std::min((int)((uavg_len + 255) / 256), e->sm_count * 8)), 256, 0, e->stream>>>(m->rare_img, T.rare.n_tiles, m->uavg,
                                                                                                            (int64_t)uavg_len);
  count_launch();
  MRS_CUDA(cudaGetLastError());
  return MRS_OK;
}

int32_t launch_item_tiled(mrs_engine* e, const mrs_ratings* R, mrs_model* m, bool fused) {
  const auto& T = R->tl;
  cudaStream_t st = e->stream;
  if (!(e->smem_attr_done & 1u)) {
    MRS_CUDA(cudaFuncSetAttribute(item_pass_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPassSmem));
    e->smem_attr_done |= 1u;
  }
  prefer_max_smem(e, user_table_kernel, 1u << 7);
  prefer_max_smem(e, item_tiled_finalize_kernel, 1u << 8);
  if (T.user_hi > T.user_lo) {  // K1b: averages + the table images of the tiles that have ratings
    MRS_CUDA(launch_pdl(user_table_kernel, dim3((T.user_hi - T.user_lo + 255) / 256), dim3(256), 0, st, m->usum, R->urow, T.user_lo, T.user_hi,
                        T.code_min, T.n_codes, m->uavg, T.pop.n_slices > 0 ? m->pop_img : (double*)nullptr, m->rare_img, e->d_timeline));
    mark(e, "user_tables");
  }
  if (T.n_ctas > 0) {
    PassArgs a;
    a.entry_pop = T.pop.entry; a.entry_rare = T.rare.entry;
    a.slice_off_pop = T.pop.slice_off; a.slice_off_rare = T.rare.slice_off;
    a.slot_item_pop = T.pop.slot_item; a.slot_item_rare = T.rare.slot_item;
    a.cta_seg_ptr = T.cta_seg_ptr; a.seg = reinterpret_cast<const PassSeg*>(T.seg); a.warp_part = T.warp_part;
    a.n_codes = T.n_codes;
    a.pop_img = m->pop_img; a.rare_img = m->rare_img; a.xdev_fix = m->xdev_fix;
    a.dbg = nullptr;
    a.tl = e->d_timeline;
    static long long* g_dbg = nullptr;  // diagnostics only (tools/pass_stamps.py): one buffer per process
    if (env_int("MRS_PASS_DEBUG", 0) == 1) {
      if (!g_dbg) MRS_CUDA(cudaMalloc((void**)&g_dbg, sizeof(long long) * 16 * 1024));
      MRS_CUDA(cudaMemsetAsync(g_dbg, 0, sizeof(long long) * 16 * 1024, st));
      a.dbg = g_dbg;
      g_pass_dbg = g_dbg;
    }
    // one CTA of 1024 threads per SM: the grid (T.n_ctas <= SM count unless there are more busy tiles than SMs) is one wave
    MRS_CUDA(launch_pdl(item_pass_kernel, dim3(T.n_ctas), dim3(kPassThreads), kPassSmem, st, a));
    mark(e, "item_tiled");
  }
  MRS_CUDA(launch_pdl(item_tiled_finalize_kernel, dim3((R->n_items + 255) / 256), dim3(256), 0, st, m->xdev_fix, m->xcode_sum, R->icolp, R->n_items,
                      m->k1_part, (double)R->n, m->xbuf, fused ? 1 : 0, m->idevavg, m->iavg, m->gavg, e->d_timeline));
  mark(e, "item_tiled_finalize");
  MRS_CUDA(cudaGetLastError());
  return MRS_OK;
}

}  // namespace mrs

// diagnostics: clock64 stamps of the last item pass launched with MRS_PASS_DEBUG=1 (16 per CTA; [15] = number of stamps)
extern "C" int32_t mrs_debug_pass_stamps(int64_t* out, int32_t n_ctas) {
  MRS_REQUIRE(out && n_ctas > 0 && n_ctas <= 1024, MRS_ERR_INVALID, "mrs_debug_pass_stamps: bad argument");
  MRS_REQUIRE(mrs::g_pass_dbg, MRS_ERR_INVALID, "mrs_debug_pass_stamps: no pass has run with MRS_PASS_DEBUG=1");
  MRS_CUDA(cudaDeviceSynchronize());
  MRS_CUDA(cudaMemcpy(out, mrs::g_pass_dbg, sizeof(int64_t) * 16 * (size_t)n_ctas, cudaMemcpyDeviceToHost));
  return MRS_OK;
}
