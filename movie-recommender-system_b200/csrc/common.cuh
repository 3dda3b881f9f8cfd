// common.cuh -- shared declarations of libmrs_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <atomic>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/mrs_b200.h"

namespace mrs {

// ---------- error plumbing (no exceptions across the C ABI) ----------
void set_error(const char* fmt, ...);
extern std::atomic<long long> g_launches;
inline void count_launch(int n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

#define MRS_CUDA(call)                                                                     \
  do {                                                                                     \
    cudaError_t _e = (call);                                                               \
    if (_e != cudaSuccess) {                                                               \
      mrs::set_error("%s failed at %s:%d: %s", #call, __FILE__, __LINE__, cudaGetErrorString(_e)); \
      return MRS_ERR_CUDA;                                                                 \
    }                                                                                      \
  } while (0)

#define MRS_TRY(expr)                 \
  do {                                \
    int32_t _s = (expr);              \
    if (_s != MRS_OK) return _s;      \
  } while (0)

#define MRS_REQUIRE(cond, code, ...)  \
  do {                                \
    if (!(cond)) {                    \
      mrs::set_error(__VA_ARGS__);    \
      return (code);                  \
    }                                 \
  } while (0)

constexpr int kWarp = 32;
// rows (users) / columns (items) longer than this are split into chunks handled by different warps
constexpr int kUserChunk = 2048;
constexpr int kItemChunk = 1024;

// value kinds of a rating set
enum : int32_t { kValueCode = 0, kValueF64 = 1 };

__host__ __device__ inline double decode_value(uint8_t c) { return 0.5 * (double)c; }
__host__ __device__ inline double decode_value(double v) { return v; }

// Programmatic dependent launch (PDL): a kernel launched with launch_pdl() may start while its predecessor in the
// stream is still running; it must call pdl_wait() before touching anything the predecessor writes.  pdl_trigger()
// in the predecessor lets the dependent start being scheduled (its prologue then overlaps the predecessor's tail and
// the launch latency disappears from the critical path: capture r01 showed 4-5 us of idle SMs per kernel).
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// diagnostics (MRS_TIMELINE=1): slot 2k = earliest block start, 2k+1 = latest block end of kernel k of a pass, %globaltimer ns
__device__ __forceinline__ unsigned long long gtimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void tl_begin(unsigned long long* tl, int k) {
  if (tl && threadIdx.x == 0) atomicMin(tl + 2 * k, gtimer_ns());
}
// ---- kernel-to-kernel hand-over inside the baseline pass without waiting for grid completion: the producer's CTAs count
// themselves off in flags[k] behind a fence once their results are out (release), the consumer -- launched early by
// programmatic dependent launch -- has one thread spin on the count (acquire) and then releases its CTA.  Measured against
// OPT-IN (MRS_FLAGSYNC=1): measured SLOWER than griddepcontrol.wait at ml-25m shape -- the fence behind a CTA's ~10^4
// outstanding reductions costs more than the hand-over saves (70.5 us against 62.8 us per step, tools/timeline.py).
// flags = mrs_model::k1_part: [0] sum of all codes, [1] user-sum blocks done, [2] item-pass CTAs done; all re-armed together.
__device__ __forceinline__ void flag_count_off(unsigned long long* flag) {  // all threads of the CTA
  __syncthreads();  // the CTA's writes and atomics happen before thread 0's fence (cumulative): one fence, not one per thread
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(flag, 1ull);
  }
}
__device__ __forceinline__ void flag_wait(const unsigned long long* flag, unsigned long long target) {  // all threads of the CTA
  if (threadIdx.x == 0) {
    unsigned long long v;
    do {
      asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(flag) : "memory");
    } while (v < target);
  }
  __syncthreads();
}
// per-CTA stamp (slot 0..3, see mrs_debug_cta_stamps)
__device__ __forceinline__ void tl_cta(unsigned long long* tl, int slot) {
  if (tl && threadIdx.x == 0 && blockIdx.x < 256) tl[32 + slot * 256 + blockIdx.x] = gtimer_ns();
}
__device__ __forceinline__ void tl_end(unsigned long long* tl, int k) {
  if (tl && threadIdx.x == 0) atomicMax(tl + 2 * k + 1, gtimer_ns());
}

// ---- fused compute + collective over NVLink peer memory ("push" exchange): what a kernel needs to deliver its partial
// sums straight into every rank's symmetric receive buffer and to wait for the other ranks' deliveries (exchange.cu)
struct PushDev {
  double* const* peer;           // device array: base of every rank's symmetric allocation (own = local memory)
  int32_t rank, world;
  int64_t cap;                   // doubles per (parity, rank) receive slot
  int64_t recv_off;              // offset, in doubles from a base, of recv[2][world][cap]
  int64_t flag_off;              // offset, in 8-byte words from a base, of the delivery flags [world] (epoch numbers)
  unsigned long long* epoch;     // device: completed push exchanges on this handle (graph replays read it on the device)
  unsigned int* done;            // device: [2] block counters (deliveries, completions)
  int32_t* error;                // device: set when a peer did not deliver in time
  long long timeout_cycles;
};
__device__ __forceinline__ double* push_slot(const PushDev& x, int p, int parity, int from) {
  return x.peer[p] + x.recv_off + ((size_t)parity * x.world + from) * x.cap;
}
// q-th destination of this rank's deliveries: the ranks start with their own successor and go round, so that the stores of
// all ranks do not converge on one NVSwitch port at a time (with p = 0, 1, ... every rank writes to rank 0 first)
__device__ __forceinline__ int push_peer(const PushDev& x, int q) {
#ifdef MRS_PUSH_NOROTATE
  return q;
#else
  const int p = x.rank + 1 + q;
  return p >= x.world ? p - x.world : p;
#endif
}
__device__ __forceinline__ void push_flag_raise(const PushDev& x, int p, unsigned long long epoch) {
  unsigned long long* f = reinterpret_cast<unsigned long long*>(x.peer[p]) + x.flag_off + x.rank;
  asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(f), "l"(epoch) : "memory");
}
// thread p (< world) waits until rank p has delivered epoch `epoch`; returns false on time-out (error word set)
__device__ __forceinline__ bool push_flag_wait(const PushDev& x, int p, unsigned long long epoch) {
  const unsigned long long* f = reinterpret_cast<const unsigned long long*>(x.peer[x.rank]) + x.flag_off + p;
  const long long t0 = clock64();
  unsigned long long v;
  do {
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(f) : "memory");
    if (v >= epoch) return true;
  } while (clock64() - t0 <= x.timeout_cycles);
  atomicExch(x.error, 1);
  return false;
}

// P:57-61 -- strict comparisons, equality -> 1
__host__ __device__ inline double scale_fn(double x, double y) {
  if (x > y) return 5.0 - y;
  if (x < y) return y - 1.0;
  return 1.0;
}
// P:229 / P:383 / P:578: the branch is taken on the rounded sum avg+dev.  No FMA contraction:
// the JVM evaluates the multiply and the add separately.
__device__ inline double combine_fn(double avg, double dev) {
  return __dadd_rn(avg, __dmul_rn(dev, scale_fn(__dadd_rn(avg, dev), avg)));
}
// P:167 / P:327 / P:516: (r - avg) / scale(r, avg), correctly rounded subtract and divide
__device__ inline double deviation_fn(double r, double avg) { return __ddiv_rn(__dsub_rn(r, avg), scale_fn(r, avg)); }

}  // namespace mrs

// ---------- handle layouts ----------
struct mrs_engine {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  int sm_count = 148;
  void* scratch = nullptr;  // reusable device scratch (CUB temp storage etc.)
  size_t scratch_bytes = 0;
  double* h_pinned = nullptr;  // small pinned staging area for scalar read-backs
  cudaStream_t copy_stream = nullptr;  // host->device copies of staged uploads (mrs_upload_begin) run beside the kernels
  cudaEvent_t ev_order = nullptr;      // orders the copy stream behind work already enqueued on `stream`
  // device block cache: buffers released by destroyed handles are kept (exact size match) and handed out again, so
  // rebuilding a rating set or a model does not go back to the driver (cudaMalloc/cudaFree of ~1 GB of layouts cost
  // 100+ ms per build).  Everything runs on one stream, so reuse is ordered.
  std::unordered_multimap<size_t, void*> free_blocks;
  std::unordered_map<void*, size_t> live_blocks;
  size_t cached_bytes = 0;
  // cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is a per-device setting: remembered per engine, not per process
  // (bit 0: item pass kernels, bit 1: test pass kernel)
  uint32_t smem_attr_done = 0;
  unsigned long long* d_timeline = nullptr;  // [32 + 1024] diagnostics, allocated when MRS_TIMELINE=1 (mrs_debug_timeline, mrs_debug_cta_stamps)
  // diagnostics: event after every launch while profiling (mrs_profile_begin/end)
  bool profiling = false;
  std::vector<cudaEvent_t> prof_events;
  std::vector<const char*> prof_names;
};

// host -> device copies of one rating set in flight on the engine's copy stream (mrs_upload_begin)
struct mrs_upload {
  mrs_engine* eng = nullptr;
  int64_t n = 0;
  int32_t* d_u = nullptr;
  int32_t* d_i = nullptr;
  double* d_r = nullptr;
  uint8_t* d_c = nullptr;           // compact form: half-star codes (2 x rating) instead of fp64 ratings
  cudaEvent_t ev_items = nullptr;   // the items have arrived (they travel first: their sort runs while the users are on the link); may be null
  cudaEvent_t ev_ids = nullptr;     // users and items have arrived
  cudaEvent_t ev_values = nullptr;  // ratings have arrived
};

struct mrs_chunks {
  // Segments ("rows" = users or item columns) longer than the chunk size are split; one warp per chunk.
  int32_t n_chunks = 0;
  int32_t* chunk_seg = nullptr;    // [n_chunks] owning segment
  int32_t* chunk_begin = nullptr;  // [n_chunks] first entry
  int32_t* seg_chunk_ptr = nullptr;  // [n_seg+1] chunks of a segment are contiguous
};

struct mrs_ratings {
  mrs_engine* eng = nullptr;
  int64_t n = 0;
  int32_t n_users = 0;  // table size = max user id + 1 (or the caller's dim)
  int32_t n_items = 0;
  int32_t value_kind = mrs::kValueCode;
  // user-major CSR, items ascending inside a row
  int32_t* urow = nullptr;  // [n_users+1]
  int32_t* ucol = nullptr;  // [n] item id
  void* uval = nullptr;     // [n] uint8 code or double
  int32_t* coo_u = nullptr;  // [n] user id of each CSR entry (sorted COO = coo_u, ucol, uval)
  // code path, fit kernel K1: every user's codes padded with zeros to a multiple of 16 so that each 128-bit vector
  // belongs to exactly one user
  uint8_t* uval16 = nullptr;   // [n_vec * 16]
  int32_t* vec_row = nullptr;  // [n_vec] user of each vector
  int32_t n_vec = 0;
  // item-major CSC, users ascending inside a column
  int32_t* icolp = nullptr;  // [n_items+1]
  int32_t* irow = nullptr;   // [n] user id
  void* ival = nullptr;      // [n]
  int32_t* csc_src = nullptr;  // [n] position of the same rating in the CSR arrays
  mrs_chunks uch, ich;
  size_t value_size() const { return value_kind == mrs::kValueCode ? 1 : 8; }
  // ---- lazily built layout for the user-user similarity kernels (knn.cu): depends only on the sparsity pattern
  struct sim_layout {
    bool built = false;
    int32_t n_known = 0;           // users with at least one rating
    int32_t* known_user = nullptr; // [n_known] original id, ascending            (compact index c -> user)
    int32_t* cidx = nullptr;       // [n_users] compact index or -1
    // ---- dense-matrix path (knn.cu similarity_wide_kernel): work partition, cached with the sparsity pattern
    int4* vmeta = nullptr;         // [wide_p][wide_npos] user at each position of the length-sorted order (longest row first, -1 =
                                   // padding) and the CSR range of its entries that fall into phase p
    int32_t wide_npos = 0;         // positions: n_known rounded up to a multiple of 32
    int4* wide_desc = nullptr;     // [n_wide] CTA b: row block .x (32 positions), column users = positions [.y, .z)
    int32_t n_wide = 0;
    int32_t* ccd = nullptr;        // [n] compact user index of each CSC entry (prediction kernels)
    int32_t wide_ic = 0, wide_p = 0;  // items per phase (the tile holds wide_ic x 32 fp64 values), number of phases
    bool wide_built = false;       // built only for the dense-matrix path
    std::vector<int32_t> h_known;  // host copies used when a row range is laid out (knn_rows.cu)
    std::vector<int32_t> h_order;  // compact indices sorted by row length, longest first
    std::vector<int32_t> h_len;    // [n_known] row length
    // ---- row-block path (knn_rows.cu): item-major arrays in compact user indices + per-column segment table
    bool rows_built = false;
    int32_t n_sub = 0;             // sub-ranges of kRowsSub compact indices (padded to a multiple of the warps per CTA)
    int32_t* clen = nullptr;       // [n_known] row length of each compact index
    int32_t* ccv = nullptr;        // [n] compact user index of each CSC entry
    int32_t* seg = nullptr;        // [n_items * (n_sub+1)] first CSC entry of column i whose compact index is >= b * kRowsSub
  };
  mutable sim_layout sl;
  // ---- lazily built tiled item-major layout for the fit kernel (tiled.cu); half-star codes only
  struct tiled_layout {
    bool built = false;
    int32_t n_tiles = 0;
    int32_t n_units = 0;
    int32_t n_slices = 0;
    int64_t n_slots = 0;               // 32 * (rows of all slices)
    uint32_t* entry = nullptr;         // [n_slots] form 0: bit31 valid | code << 16 | user id local to the tile
                                       //           form 1: top 15 bits of the fp64 (code-2)/8 | bit16 padding | local user << 3
    int32_t form = 0;                  // 1 when every code is <= kAlphaMaxCode and every user has < 2^17 ratings (tiled.cu)
    int32_t* slice_off = nullptr;      // [n_slices+1] first 32-wide row of each slice
    int32_t* tile_slice_ptr = nullptr; // [n_tiles+1]
    int32_t* tile_ubegin = nullptr;    // [n_tiles+1] first user of every tile (tiles are user ranges of <= kTileUsers users, cut by cost)
    int32_t* slot_item = nullptr;      // [n_slices*32] item of the unit held by each slot, -1 = empty slot
    // static work partition of the item pass (depends on the layout and the SM count only; laid down with the layout):
    // CTA b works on tile cta_desc[b].x as share .y of .z -- CTAs are dealt out to the tiles in proportion to their
    // cost, tiles without ratings (a rank of a sharded run owns a user range) get none
    int32_t n_ctas = 0;
    int3* cta_desc = nullptr;          // [n_ctas]
    int4* warp_part = nullptr;         // [n_ctas * 32] rows [x, y) of every warp, first slice, one past the last slice
  };
  mutable tiled_layout tl;
  // ---- lazily built item-tiled layout for the fused predict + |error| kernel (mae_tiled.cu); half-star codes only
  struct mae_layout {
    bool built = false;
    int32_t n_tiles = 0;        // item tiles of kMaeTileItems
    int64_t n_rows = 0;         // 32-entry rows incl. padding (every tile starts at a row boundary)
    uint2* entry = nullptr;     // [n_rows*32] .x = user id, .y = local item | code << 16 (code 0xFF = padding)
    int32_t* tile_row_ptr = nullptr;  // [n_tiles+1] first row of every tile
    int32_t n_ctas = 0;         // CTAs dealt out to the item tiles in proportion to their rows (see tiled_layout)
    int3* cta_desc = nullptr;   // [n_ctas] (tile, share, shares of the tile)
  };
  mutable mae_layout ml;
};

struct mrs_model {
  mrs_engine* eng = nullptr;
  const mrs_ratings* train = nullptr;
  int32_t n_users = 0, n_items = 0;
  uint32_t* usum = nullptr;     // [n_users] sum of half-star codes per user (code path)
  unsigned long long* k1_part = nullptr;  // [4] sum of all half-star codes (integer atomics in K1), user-sum blocks done, item-pass CTAs
                                          // done (flag_count_off / flag_wait); re-armed together by the last consumer of a pass
  bool flag_sync = true;                  // MRS_FLAGSYNC=0: hand-overs by griddepcontrol.wait only
  int32_t k1_blocks = 0;
  long long* xdev_fix = nullptr;            // [n_items] per-item deviation sums in units of 2^-40 (exact integer accumulation)
  unsigned long long* xcode_sum = nullptr;  // [n_items] per-item sums of half-star codes
  bool want_item_avg = true;                // also accumulate per-item rating sums during the fit (P:134; not needed by P:362)
  double* upart = nullptr;      // [uch.n_chunks] chunk partial sums of ratings (fp64-value path)
  double* uavg = nullptr;       // [n_users]  average, -1.0 for unknown users (the reference's own sentinel, P:222)
  double* uinv_hi = nullptr;    // [n_users]  1 / (5 - avg)   (code path: reciprocal of scale() for ratings above the average)
  double* uinv_lo = nullptr;    // [n_users]  1 / (avg - 1)   (                              ... below the average)
  double* ipart = nullptr;      // [2 * ich.n_chunks] chunk partials: deviations | ratings
  double* xbuf = nullptr;       // [3*n_items + 2] exchange buffer: devsum | count | gsum gcount | ratesum
  double* idevavg = nullptr;    // [n_items]  0.0 for unknown items (P:197)
  double* iavg = nullptr;       // [n_items]  NaN for unknown items (replaced by the global average at query time, P:147)
  double* gavg = nullptr;       // [1] device scalar
  double* mae_part = nullptr;   // per-block partials of |err| sums
  unsigned int* counters = nullptr;  // small set of device counters (last-block-done patterns)
  int32_t mae_part_cap = 0;
  // fused push exchange (mrs_fit_local_push): the items that occur on some rank, ascending (compact slot j <-> item slot_of_item[j])
  int32_t* slot_of_item = nullptr;  // [n_slots_known]
  int32_t* item_slot = nullptr;     // [n_items] inverse map (-1: the item occurs on no rank); mrs_fit_mae_push_async
  int32_t n_slots_known = 0;        // K: items that occur on some rank; a delivery is [K dev sums | K counts | sum, n]
  // order of neighbours with EXACTLY equal similarity: 0 = ascending user id, 1 = iteration order of a Scala 2.11
  // immutable.HashSet[Int] (what the reference's stable sort keeps, P:608-610).  tie_rank[c] = place of compact user index c in
  // that order, tie_inv = its inverse; NULL for mode 0
  int32_t tie_mode = 0;
  int32_t* tie_rank = nullptr;
  int32_t* tie_inv = nullptr;
  bool finished = false;
  // host mirrors, filled lazily by queries
  mutable bool host_valid = false;
  mutable double h_gavg = 0.0;
};

struct mrs_sim {
  mrs_model* model = nullptr;
  int32_t kind = MRS_SIM_COSINE;
  int32_t k = 0;
  int32_t n_known = 0;
  double* udev = nullptr;   // [n] deviation of each CSR entry (exact ops, P:167)
  double* upre = nullptr;   // [n] preprocessed rating r~ (P:470-481)
  double* unorm = nullptr;  // [n_known]
  double* cdev = nullptr;   // [n] deviation of each CSC entry
  int4* pk = nullptr;       // [n] (item, r~ (cosine) or 1.0 (jaccard)) of each CSR entry as one 16-byte record {item, 0, lo, hi}
  double* S = nullptr;      // [n_known^2] similarity matrix in compact indices (cosine / jaccard)
  int32_t* rank = nullptr;  // [n_known^2] position of v in u's sorted neighbour list (INT_MAX for v == u)
  int32_t* nbr_id = nullptr;  // [n_known * (n_known-1)] neighbours of each user, original ids, (sim desc, id asc)
  double* nbr_sim = nullptr;
  double* mae_part = nullptr;
  int32_t mae_part_cap = 0;
  unsigned int* counter = nullptr;
  // ---- row-block path (knn_rows.cu): no matrix; only the first k_fit neighbours of the users of a row range are kept
  bool lists = false;
  int32_t k_fit = 0;      // neighbours kept per user = min(k at fit time, n_known - 1)
  int32_t row_lo = 0, row_hi = 0;  // compact indices [row_lo, row_hi) own lists (a rank's user shard)
  int32_t* row_order = nullptr;    // [row_hi - row_lo] compact indices of the range, longest row first
  double* cpre = nullptr;          // [n] r~ (cosine) of each CSC entry
};

namespace mrs {
extern thread_local mrs_engine* tls_engine;  // engine whose block cache serves dev_alloc / dev_free on this thread
inline void use_engine(const mrs_engine* e) {
  cudaSetDevice(e->device);
  tls_engine = const_cast<mrs_engine*>(e);
}
void* cache_alloc(size_t bytes);
void cache_free(void* p);
// called right after each kernel launch: counts it and, while profiling, records an event behind it
inline void mark(mrs_engine* e, const char* name, int n = 1) {
  count_launch(n);
  if (e->profiling) {
    cudaEvent_t ev;
    if (cudaEventCreate(&ev) == cudaSuccess) {
      cudaEventRecord(ev, e->stream);
      e->prof_events.push_back(ev);
      e->prof_names.push_back(name);
    }
  }
}
int32_t ensure_scratch(mrs_engine* e, size_t bytes);
// launch with the programmatic-stream-serialization attribute (PDL); the kernel must call pdl_wait()
template <typename... KArgs, typename... Args>
cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
template <typename T>
int32_t dev_alloc(T** p, size_t count) {
  if (count == 0) count = 1;
  *p = (T*)cache_alloc(count * sizeof(T));
  return *p ? MRS_OK : MRS_ERR_NOMEM;
}
inline void dev_free(void* p) {
  if (p) cache_free(p);
}

// loader.cu
int32_t build_ratings(mrs_engine* e, const int32_t* users, const int32_t* items, const double* ratings, int64_t n,
                      int32_t n_users_dim, int32_t n_items_dim, mrs_ratings** out);
// Deal `n_ctas` CTAs out to tiles in proportion to their cost (largest remainder; every tile with cost > 0 gets at least
// one as long as there are enough CTAs): desc[b] = (tile, share, shares of that tile), tiles ascending.
std::vector<int3> deal_ctas(const std::vector<int64_t>& tile_cost, int32_t n_ctas);
// tiled.cu
#ifndef MRS_TILE_USERS
#define MRS_TILE_USERS 16384
#endif
// user records per tile of the item pass: 8 bytes each in shared memory (128 KB).  The last record is the all-zero one the
// padding slots point at, so a tile holds at most kTileCap users.  Larger tiles mean fewer (tile, item) segments -- fewer
// units, slice boundaries and hand-over atomics: 1.10 M segments with 8,192 users per tile at ml-25m shape.
constexpr int kTileUsers = MRS_TILE_USERS;
constexpr int kTileCap = kTileUsers - 1;
constexpr int kUnitLen = 64;      // (tile,item) segments are cut into units of at most this many entries
constexpr int kUnitBits = 7;      // bits of (kUnitLen - len) in the unit sort key
int32_t build_tiled_layout(const mrs_ratings* R);
void free_tiled_layout(const mrs_ratings* R);
int32_t launch_item_tiled(mrs_engine* e, const mrs_ratings* R, mrs_model* m, bool fused_finalize, const PushDev* push = nullptr, bool no_finalize = false);
int32_t launch_finish_pull(mrs_model* m, const PushDev& push);
// mae_tiled.cu
constexpr int kMaeTileItems = 8192;  // items per tile: 64 KB of fp64 item deviations in shared memory
int32_t build_mae_layout(const mrs_ratings* T);
void free_mae_layout(const mrs_ratings* T);
int32_t launch_mae_tiled_baseline(const mrs_model* m, const mrs_ratings* T, double* d_out2, const PushDev* push = nullptr, bool fold = false,
                                  const PushDev* big = nullptr, bool deliver = true);
// baseline.cu
int32_t fit_local(mrs_engine* e, const mrs_ratings* train, mrs_model** inout, bool fused_finalize, const PushDev* push = nullptr, bool no_finalize = false);
int32_t fit_users(mrs_engine* e, const mrs_ratings* train, mrs_model** inout);
int32_t fit_finish(mrs_model* m);
int32_t mae_baseline_async(const mrs_model* m, int32_t pred_kind, const mrs_ratings* test, double* d_out2);
int32_t predict_baseline_async(const mrs_model* m, int32_t pred_kind, const int32_t* d_users, const int32_t* d_items,
                               int64_t n, double* d_out);
// exchange.cu
int32_t exchange_push_dev(mrs_exchange* x, int64_t n_doubles, PushDev* out);
// knn.cu
int32_t sim_fit_async(mrs_model* m, int32_t sim_kind, int32_t k, mrs_sim** inout);
int32_t mae_personalized_async(const mrs_model* m, const mrs_sim* s, const mrs_ratings* test, double* d_out2);
int32_t predict_personalized_async(const mrs_model* m, const mrs_sim* s, const int32_t* d_users, const int32_t* d_items,
                                   int64_t n, double* d_out, bool wsd_only);
void free_sim_layout(const mrs_ratings* r);
int32_t build_sim_layout(const mrs_ratings* R, bool with_wide);
// knn_rows.cu
int32_t rows_fit_async(mrs_model* m, mrs_sim* s, bool first);
int32_t rows_alloc(mrs_model* m, mrs_sim* s, int32_t user_lo, int32_t user_hi);
void rows_free(mrs_sim* s);
void free_rows_layout(const mrs_ratings* r);
int32_t mae_lists_async(const mrs_model* m, const mrs_sim* s, const mrs_ratings* test, double* d_out2);
int32_t predict_lists_async(const mrs_model* m, const mrs_sim* s, const int32_t* d_users, const int32_t* d_items, int64_t n,
                            double* d_out, bool wsd_only);
}  // namespace mrs
