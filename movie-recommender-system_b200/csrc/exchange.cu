// exchange.cu -- the one collective of the sharded baseline pass as OUR kernel over NVLink peer memory.
//
// The reference combines per-key partial sums with Spark's reduceByKey + collect (P:267-268) and sum/count (P:247).
// On B200s of one box every rank (one process per GPU) maps the others' exchange buffers through CUDA IPC; a single
// kernel then (1) publishes this rank's partial sums in its own symmetric buffer, (2) raises a flag in every peer's
// memory (NVLink store, release at system scope), (3) waits for the peers' flags, (4) reads all partial sums with
// 128-bit peer loads and adds them in RANK ORDER, so every rank ends with bit-identical totals.  A 3.3 MB buffer is
// latency bound: NCCL's all-reduce costs 30-40 us per call here, this kernel one peer round trip plus the reads.
// Large buffers take the two-shot form so that the NVLink traffic per rank does not grow with the number of ranks: after
// the first barrier every rank reduces only ITS slice (reading that slice from every peer) into a symmetric result
// buffer, a second barrier follows, and every rank collects the reduced slices.  Small buffers ({sum |err|, n}) stay
// one-shot (one barrier).  Buffers are double-buffered by epoch parity, which removes the trailing barrier: a rank
// overwrites parity p again two epochs later, and its peers enter the next epoch only after they finished reading.
//
// The waits are bounded (about 2 s): a missing peer makes the kernel give up and set an error word instead of hanging.
#include <algorithm>
#include <cstring>
#include <vector>

#include "common.cuh"

struct mrs_exchange {
  mrs_engine* eng = nullptr;
  int32_t rank = 0, world = 1;
  int64_t n = 0;                 // capacity in doubles of one parity buffer
  unsigned char* base = nullptr; // own symmetric allocation: publish [2][n] | result [2][n] doubles | flags [2][world] u64 |
                                 // push mode: recv [2][world][n] doubles | delivery flags [world] u64
  unsigned long long* d_pepoch = nullptr;  // completed push exchanges
  unsigned int* d_pdone = nullptr;         // [2]
  std::vector<void*> peer_base;  // mapped bases of all ranks (own = base)
  double** d_peer = nullptr;     // device array of the mapped bases
  unsigned long long* d_epoch = nullptr;  // number of completed exchanges (device side, so that launches can be graph-captured)
  unsigned int* d_done = nullptr;  // [3] blocks-published / blocks-reduced / blocks-finished counters of this rank
  int32_t* d_error = nullptr;
  unsigned long long* d_stamps = nullptr;  // [8] globaltimer stamps of block 0 in the last exchange (diagnostics)
  bool connected = false;
  bool local = false;            // peers live in this process: plain pointers, nothing to unmap
  bool failed = false;           // a timed-out exchange was observed by the host: the handle refuses further exchanges
  long long timeout_cycles = 4000000000LL;  // bound of every flag wait (~2 s at 1.9 GHz); mrs_exchange_set_timeout_ms
};

namespace mrs {
namespace {

constexpr int kExThreads = 512;

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

constexpr int kMaxWorld = 16;

__device__ __forceinline__ void st_relaxed_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// Barrier across the ranks, entered by ALL threads of every block.  The last arriving block of this rank makes this
// rank's data visible (one system-scope fence) and then thread p raises flag `which` in rank p's flag row -- the `world`
// NVLink stores go out together instead of one release-store after the other (8 ranks: 20 us -> a few us).  Then thread
// p of every block polls rank p's flag in OUR row.  Returns false if a peer did not show up within ~2 s.
__device__ __forceinline__ bool rank_barrier(double* const* __restrict__ peer, int32_t rank, int32_t world, int64_t cap, int which,
                                             unsigned long long epoch, unsigned int* __restrict__ done, int32_t* __restrict__ error,
                                             long long timeout_cycles) {
  __shared__ int s_last;
  if (threadIdx.x == 0) {
    const int last = (atomicAdd(done, 1u) + 1u == gridDim.x);
    if (last) *done = 0;
    __threadfence();  // the other blocks' data (fenced before their arrival) is ordered before what follows
    s_last = last;
  }
  __syncthreads();
  if (s_last && (int)threadIdx.x < world) {
    __threadfence_system();
    unsigned long long* flags = reinterpret_cast<unsigned long long*>(peer[threadIdx.x] + 4 * (size_t)cap) + (size_t)which * world;
    st_relaxed_sys(flags + rank, epoch);
  }
  int good = 1;
  if ((int)threadIdx.x < world) {
    const unsigned long long* mine = reinterpret_cast<const unsigned long long*>(peer[rank] + 4 * (size_t)cap) + (size_t)which * world;
    const long long t0 = clock64();
    while (ld_acquire_sys(mine + threadIdx.x) < epoch) {
      if (clock64() - t0 > timeout_cycles) { atomicExch(error, 1); good = 0; break; }  // a peer is missing
    }
  }
  return __syncthreads_and(good) != 0;
}

// result pair i of the compact array -> its place in the caller's buffer
__device__ __forceinline__ void put2(double* __restrict__ inout, const int32_t* __restrict__ idx, int64_t i, double2 v) {
  if (idx) {
    inout[__ldg(idx + 2 * i)] = v.x;
    inout[__ldg(idx + 2 * i + 1)] = v.y;
  } else {
    reinterpret_cast<double2*>(inout)[i] = v;
  }
}

// inout[0..n): this rank's partial sums on entry, the sum over all ranks (added in rank order) on exit
__global__ void __launch_bounds__(kExThreads) peer_allreduce_kernel(double* const* __restrict__ peer, int32_t rank, int32_t world, int64_t n,
                                                                   int64_t cap, int two_shot, unsigned long long* __restrict__ epoch_done,
                                                                   unsigned int* __restrict__ done, int32_t* __restrict__ error,
                                                                   unsigned long long* __restrict__ stamps, double* __restrict__ inout,
                                                                   const int32_t* __restrict__ idx, long long timeout_cycles) {
  // idx != nullptr: the exchange covers the n positions idx[0..n) of `inout` only (the slots that can be non-zero on some
  // rank -- 28 % of the exchange buffer at ml-25m shape, whose item ids are sparse); everything below works on the compact
  // array, only the first read and the last write go through the index list
  pdl_trigger();  // the kernel behind this one may be scheduled as SMs free up (it waits for our completion itself)
  pdl_wait();     // `inout` is complete
  bool ok;
  const bool stamp = (blockIdx.x == 0 && threadIdx.x == 0);
  if (stamp) stamps[0] = gtime();
  const unsigned long long epoch = *epoch_done + 1;  // stable for the whole kernel: only its last block advances it
  const int parity = (int)(epoch & 1);
  const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x, nth = (int64_t)gridDim.x * blockDim.x;
  double* pub = peer[rank] + (size_t)parity * cap;
  // (1) publish (128-bit copies, 4 in flight per thread: the buffer is a few MB and latency bound)
  if (idx) {
    for (int64_t j = tid; j < n; j += nth) pub[j] = inout[__ldg(idx + j)];
  } else {
    const int64_t n2 = n >> 1;
    const double2* src = reinterpret_cast<const double2*>(inout);
    double2* dst = reinterpret_cast<double2*>(pub);
    int64_t i = tid;
    for (; i + 3 * nth < n2; i += 4 * nth) {
      const double2 a = src[i], b = src[i + nth], c = src[i + 2 * nth], d = src[i + 3 * nth];
      dst[i] = a; dst[i + nth] = b; dst[i + 2 * nth] = c; dst[i + 3 * nth] = d;
    }
    for (; i < n2; i += nth) dst[i] = src[i];
    if ((n & 1) && tid == 0) pub[n - 1] = inout[n - 1];
  }
  __threadfence();  // device scope: the barrier's last block makes everything visible system-wide before it raises the flags
  __syncthreads();
  if (stamp) stamps[1] = gtime();
  ok = rank_barrier(peer, rank, world, cap, 0, epoch, done, error, timeout_cycles);  // (2) everybody has published
  if (stamp) stamps[2] = gtime();
  if (ok && !two_shot) {
    // one-shot: every rank adds all partial sums, in rank order (identical everywhere); 2 x 128-bit peer loads per
    // rank in flight per thread
    const int64_t n2 = n >> 1;
    int64_t i = tid;
    for (; i + nth < n2; i += 2 * nth) {
      double2 a0 = make_double2(0.0, 0.0), a1 = a0;
      for (int p = 0; p < world; ++p) {
        const double2* src = reinterpret_cast<const double2*>(peer[p] + (size_t)parity * cap);
        const double2 v0 = src[i], v1 = src[i + nth];
        a0.x += v0.x; a0.y += v0.y; a1.x += v1.x; a1.y += v1.y;
      }
      put2(inout, idx, i, a0);
      put2(inout, idx, i + nth, a1);
    }
    for (; i < n2; i += nth) {
      double2 a0 = make_double2(0.0, 0.0);
      for (int p = 0; p < world; ++p) {
        const double2 v0 = reinterpret_cast<const double2*>(peer[p] + (size_t)parity * cap)[i];
        a0.x += v0.x; a0.y += v0.y;
      }
      put2(inout, idx, i, a0);
    }
    if ((n & 1) && tid == 0) {
      double a = 0.0;
      for (int p = 0; p < world; ++p) a += (peer[p] + (size_t)parity * cap)[n - 1];
      inout[idx ? idx[n - 1] : n - 1] = a;
    }
  } else if (ok) {
    // two-shot, n even: (3) reduce my slice in rank order into my result buffer ...
    const int64_t n2 = n >> 1;
    const int64_t lo = n2 * rank / world, hi = n2 * (rank + 1) / world;
    double2* res = reinterpret_cast<double2*>(peer[rank] + (size_t)(2 + parity) * cap);
    for (int64_t i = lo + tid; i < hi; i += nth) {
      double2 v[kMaxWorld];
#pragma unroll
      for (int p = 0; p < kMaxWorld; ++p)  // all peer loads of the element go out together
        if (p < world) v[p] = reinterpret_cast<const double2*>(peer[p] + (size_t)parity * cap)[i];
      double2 acc = make_double2(0.0, 0.0);
#pragma unroll
      for (int p = 0; p < kMaxWorld; ++p)
        if (p < world) { acc.x += v[p].x; acc.y += v[p].y; }  // rank order
      res[i] = acc;
    }
    __threadfence();
    __syncthreads();
    if (stamp) stamps[3] = gtime();
    ok = rank_barrier(peer, rank, world, cap, 1, epoch, done + 1, error, timeout_cycles);  // (4) every slice is reduced
    if (stamp) stamps[4] = gtime();
    if (ok) {
      // (5) ... and collect the reduced slices of all ranks
      // element i lives in the result buffer of its owner; the owner is recovered from the slice bounds, so one
      // flat loop keeps loads from all peers in flight
      for (int64_t i0 = tid; i0 < n2; i0 += 4 * nth) {
        double2 v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int64_t i = i0 + k * nth;
          if (i < n2) {
            int p = (int)((i * world) / n2);                 // candidate owner, then fix the rounding of the bounds
            while (p + 1 < world && i >= n2 * (p + 1) / world) ++p;
            while (p > 0 && i < n2 * p / world) --p;
            v[k] = reinterpret_cast<const double2*>(peer[p] + (size_t)(2 + parity) * cap)[i];
          }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int64_t i = i0 + k * nth;
          if (i < n2) put2(inout, idx, i, v[k]);
        }
      }
    }
  }
  __syncthreads();
  if (stamp) stamps[5] = gtime();
  // A peer that never showed up (error word set by whichever block gave up): the sums are incomplete.  Poison the
  // caller's buffer so that nothing downstream can pass for a result (the MAE becomes NaN), whatever this block saw.
  if (*reinterpret_cast<volatile int32_t*>(error) != 0)
    for (int64_t j = tid; j < n; j += nth) inout[idx ? __ldg(idx + j) : j] = nan("");
  if (threadIdx.x == 0 && atomicAdd(done + 2, 1u) + 1u == gridDim.x) {  // last block out: this exchange is complete
    done[2] = 0;
    done[1] = 0;  // blocks that gave up in the first barrier never entered the second one: re-arm it
    *epoch_done = epoch;
  }
}

}  // namespace

// device-side view of an exchange handle for the kernels that deliver their own partial sums (tiled.cu, mae_tiled.cu)
int32_t exchange_push_dev(mrs_exchange* x, int64_t n_doubles, PushDev* out) {
  MRS_REQUIRE(x && out, MRS_ERR_INVALID, "push exchange: NULL argument");
  MRS_REQUIRE(x->connected, MRS_ERR_INVALID, "push exchange: call mrs_exchange_connect first");
  MRS_REQUIRE(!x->failed, MRS_ERR_CUDA, "push exchange: an earlier exchange on this handle timed out; the handle is dead");
  MRS_REQUIRE(n_doubles > 0 && n_doubles <= x->n, MRS_ERR_INVALID, "push exchange: %lld doubles exceed the capacity %lld", (long long)n_doubles,
              (long long)x->n);
  out->peer = x->d_peer;
  out->rank = x->rank; out->world = x->world;
  out->cap = x->n;
  out->recv_off = 4 * x->n + 2 * (int64_t)x->world;
  out->flag_off = out->recv_off + 2 * (int64_t)x->world * x->n;
  out->epoch = x->d_pepoch;
  out->done = x->d_pdone;
  out->error = x->d_error;
  out->timeout_cycles = x->timeout_cycles;
  return MRS_OK;
}

}  // namespace mrs

using namespace mrs;

// ---- C ABI (declared in include/mrs_b200.h)
extern "C" int32_t mrs_exchange_create(mrs_engine* e, int64_t n_doubles, int32_t rank, int32_t world, void* ipc_handle_out64,
                                       mrs_exchange** out) {
  MRS_REQUIRE(e && out && ipc_handle_out64 && n_doubles > 0 && world >= 1 && world <= 16 && rank >= 0 && rank < world, MRS_ERR_INVALID,
              "mrs_exchange_create: bad argument (1 <= world <= 16)");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  use_engine(e);
  mrs_exchange* x = new mrs_exchange();
  x->eng = e; x->rank = rank; x->world = world;
  x->n = (n_doubles + 1) & ~(int64_t)1;
  const size_t bytes = 4 * (size_t)x->n * sizeof(double) + 2 * (size_t)world * sizeof(unsigned long long) +
                       2 * (size_t)world * (size_t)x->n * sizeof(double) + (size_t)world * sizeof(unsigned long long) + 64;
  // IPC-shareable memory must come from cudaMalloc directly (not from the engine's block cache)
  cudaError_t ce = cudaMalloc((void**)&x->base, bytes);
  if (ce != cudaSuccess) { delete x; set_error("mrs_exchange_create: cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(ce)); return MRS_ERR_NOMEM; }
  MRS_CUDA(cudaMemset(x->base, 0, bytes));
  MRS_CUDA(cudaMalloc((void**)&x->d_done, 4 * sizeof(unsigned int)));
  MRS_CUDA(cudaMemset(x->d_done, 0, 4 * sizeof(unsigned int)));
  MRS_CUDA(cudaMalloc((void**)&x->d_pepoch, sizeof(unsigned long long)));
  MRS_CUDA(cudaMemset(x->d_pepoch, 0, sizeof(unsigned long long)));
  MRS_CUDA(cudaMalloc((void**)&x->d_pdone, 4 * sizeof(unsigned int)));
  MRS_CUDA(cudaMemset(x->d_pdone, 0, 4 * sizeof(unsigned int)));
  MRS_CUDA(cudaMalloc((void**)&x->d_epoch, sizeof(unsigned long long)));
  MRS_CUDA(cudaMemset(x->d_epoch, 0, sizeof(unsigned long long)));
  MRS_CUDA(cudaMalloc((void**)&x->d_error, sizeof(int32_t)));
  MRS_CUDA(cudaMemset(x->d_error, 0, sizeof(int32_t)));
  MRS_CUDA(cudaMalloc((void**)&x->d_stamps, 8 * sizeof(unsigned long long)));
  MRS_CUDA(cudaMemset(x->d_stamps, 0, 8 * sizeof(unsigned long long)));
  MRS_CUDA(cudaMalloc((void**)&x->d_peer, sizeof(double*) * (size_t)world));
  cudaIpcMemHandle_t h;
  MRS_CUDA(cudaIpcGetMemHandle(&h, x->base));
  memcpy(ipc_handle_out64, &h, sizeof(h));
  *out = x;
  return MRS_OK;
}

extern "C" int32_t mrs_exchange_connect(mrs_exchange* x, const void* all_handles_world_x_64) {
  MRS_REQUIRE(x && all_handles_world_x_64, MRS_ERR_INVALID, "mrs_exchange_connect: NULL argument");
  use_engine(x->eng);
  x->peer_base.assign((size_t)x->world, nullptr);
  for (int p = 0; p < x->world; ++p) {
    if (p == x->rank) { x->peer_base[p] = x->base; continue; }
    cudaIpcMemHandle_t h;
    memcpy(&h, (const unsigned char*)all_handles_world_x_64 + (size_t)p * 64, sizeof(h));
    void* ptr = nullptr;
    cudaError_t ce = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess);
    if (ce != cudaSuccess) { set_error("mrs_exchange_connect: cannot map the buffer of rank %d: %s", p, cudaGetErrorString(ce)); return MRS_ERR_CUDA; }
    x->peer_base[p] = ptr;
  }
  MRS_CUDA(cudaMemcpy(x->d_peer, x->peer_base.data(), sizeof(void*) * (size_t)x->world, cudaMemcpyHostToDevice));
  x->connected = true;
  return MRS_OK;
}

// the ranks are devices of THIS process (mrs_multi_*): with peer access enabled their buffers are plain pointers
extern "C" int32_t mrs_exchange_connect_local(mrs_exchange** xs, int32_t world) {
  MRS_REQUIRE(xs && world >= 1, MRS_ERR_INVALID, "mrs_exchange_connect_local: bad argument");
  for (int r = 0; r < world; ++r) MRS_REQUIRE(xs[r] && xs[r]->world == world && xs[r]->rank == r, MRS_ERR_INVALID, "mrs_exchange_connect_local: handle %d does not match", r);
  for (int r = 0; r < world; ++r) {
    mrs_exchange* x = xs[r];
    use_engine(x->eng);
    x->peer_base.assign((size_t)world, nullptr);
    for (int p = 0; p < world; ++p) x->peer_base[(size_t)p] = xs[p]->base;
    MRS_CUDA(cudaMemcpy(x->d_peer, x->peer_base.data(), sizeof(void*) * (size_t)world, cudaMemcpyHostToDevice));
    x->connected = true;
    x->local = true;
  }
  return MRS_OK;
}

static int32_t allreduce_impl(mrs_exchange* x, void* device_inout, int64_t n_doubles, const int32_t* device_idx);

extern "C" int32_t mrs_exchange_allreduce_async(mrs_exchange* x, void* device_inout, int64_t n_doubles) {
  return allreduce_impl(x, device_inout, n_doubles, nullptr);
}

extern "C" int32_t mrs_exchange_allreduce_indexed_async(mrs_exchange* x, void* device_inout, const int32_t* device_idx, int64_t n_idx) {
  MRS_REQUIRE(device_idx, MRS_ERR_INVALID, "mrs_exchange_allreduce_indexed_async: NULL index list");
  return allreduce_impl(x, device_inout, n_idx, device_idx);
}

static int32_t allreduce_impl(mrs_exchange* x, void* device_inout, int64_t n_doubles, const int32_t* device_idx) {
  MRS_REQUIRE(x && device_inout, MRS_ERR_INVALID, "mrs_exchange_allreduce_async: NULL argument");
  MRS_REQUIRE(x->connected, MRS_ERR_INVALID, "mrs_exchange_allreduce_async: call mrs_exchange_connect first");
  MRS_REQUIRE(n_doubles > 0 && n_doubles <= x->n, MRS_ERR_INVALID, "mrs_exchange_allreduce_async: %lld doubles exceed the capacity %lld",
              (long long)n_doubles, (long long)x->n);
  MRS_REQUIRE(((uintptr_t)device_inout & 15) == 0, MRS_ERR_INVALID, "mrs_exchange_allreduce_async: buffer must be 16-byte aligned");
  MRS_REQUIRE(!x->failed, MRS_ERR_CUDA, "mrs_exchange_allreduce_async: an earlier exchange on this handle timed out (a peer was missing); "
              "its results were poisoned with NaN and the handle is dead -- destroy it and create a new one");
  use_engine(x->eng);
  // every block waits on the flags: the grid must be co-resident (one CTA per SM at most)
  const int64_t work = (n_doubles + 1) / 2;
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((work + kExThreads - 1) / kExThreads, (int64_t)x->eng->sm_count));
  // two-shot (reduce my slice, barrier, collect) once the buffer is large enough to be bandwidth relevant
  // (its second barrier, fence and phase cost about 12 us, i.e. about 10 MB of NVLink time: below that every rank simply
  // reads all the peers' buffers)
  const int two_shot = (x->world > 2 && (n_doubles & 1) == 0 && (int64_t)(x->world - 1) * n_doubles * 8 > (int64_t)12 << 20) ? 1 : 0;
  MRS_CUDA(launch_pdl(peer_allreduce_kernel, dim3(grid), dim3(kExThreads), 0, x->eng->stream, x->d_peer, x->rank, x->world, n_doubles, x->n,
                      two_shot, x->d_epoch, x->d_done, x->d_error, x->d_stamps, (double*)device_inout, device_idx, x->timeout_cycles));
  mark(x->eng, "peer_allreduce");
  MRS_CUDA(cudaGetLastError());
  return MRS_OK;
}

// ---- fused compute + collective entry points (the exchange happens inside the fit's and the test pass' own kernels)
extern "C" int32_t mrs_fit_local_push(mrs_engine* e, const mrs_ratings* train, mrs_model** inout, mrs_exchange* x, const int32_t* device_known_items,
                                      int32_t n_known) {
  MRS_REQUIRE(e && train && inout && x && device_known_items && n_known >= 0, MRS_ERR_INVALID, "mrs_fit_local_push: NULL argument");
  MRS_REQUIRE(*inout, MRS_ERR_INVALID, "mrs_fit_local_push: pass a model of this train set (mrs_fit_local once, with item averages switched off)");
  mrs_model* m = *inout;
  use_engine(e);
  if (!m->slot_of_item || m->n_slots_known != n_known) {  // first use (never inside a graph capture): keep our own copy of the list
    dev_free(m->slot_of_item);
    m->slot_of_item = nullptr;
    MRS_TRY(dev_alloc(&m->slot_of_item, (size_t)std::max(n_known, 1)));
    MRS_CUDA(cudaMemcpyAsync(m->slot_of_item, device_known_items, sizeof(int32_t) * (size_t)n_known, cudaMemcpyDeviceToDevice, e->stream));
    // items known on no rank are never written by the finishing kernel: their deviation is 0.0 (P:197), their count 0
    MRS_CUDA(cudaMemsetAsync(m->idevavg, 0, sizeof(double) * (size_t)m->n_items, e->stream));
    MRS_CUDA(cudaMemsetAsync(m->xbuf, 0, sizeof(double) * (2 * (size_t)m->n_items + 2), e->stream));
    m->n_slots_known = n_known;
  }
  PushDev pd;
  MRS_TRY(exchange_push_dev(x, 2 * (int64_t)n_known + 2, &pd));
  return fit_local(e, train, inout, false, &pd);
}

extern "C" int32_t mrs_fit_finish_pull(mrs_model* m, mrs_exchange* x) {
  MRS_REQUIRE(m && x && m->slot_of_item, MRS_ERR_INVALID, "mrs_fit_finish_pull: call mrs_fit_local_push first");
  use_engine(m->eng);
  PushDev pd;
  MRS_TRY(exchange_push_dev(x, 2 * (int64_t)m->n_slots_known + 2, &pd));
  return launch_finish_pull(m, pd);
}

extern "C" int32_t mrs_mae_push_async(const mrs_model* m, const mrs_ratings* test, mrs_exchange* x, void* device_out2) {
  MRS_REQUIRE(m && test && x && device_out2, MRS_ERR_INVALID, "mrs_mae_push_async: NULL argument");
  MRS_REQUIRE(m->finished, MRS_ERR_INVALID, "mrs_mae_push_async: model not finished");
  MRS_REQUIRE(m->eng == test->eng, MRS_ERR_INVALID, "mrs_mae_push_async: model and test set live on different engines");
  MRS_REQUIRE(test->value_kind == kValueCode && test->n > 0, MRS_ERR_UNSUPPORTED, "mrs_mae_push_async: needs a non-empty half-star coded test set");
  use_engine(m->eng);
  PushDev pd;
  MRS_TRY(exchange_push_dev(x, 2, &pd));
  return launch_mae_tiled_baseline(m, test, (double*)device_out2, &pd);
}

namespace mrs {
namespace {
__global__ void item_slot_kernel(const int32_t* __restrict__ known, int32_t K, int32_t* __restrict__ item_slot) {
  const int32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < K) item_slot[known[j]] = j;
}
}  // namespace
}  // namespace mrs

// The sharded closure MeanAbsoluteErrorSpark(baselinePredictorSpark(train), test) (distributed/DistributedBaseline.scala:45-47)
// of one rank as THREE kernels: user sums, item pass, and a test pass whose prologue is the exchange of the fit -- every CTA
// delivers an equal share of the per-item partial sums to every rank, waits for all ranks' flags, builds its tile's
// deviations from the deliveries in its own memory -- and whose last block exchanges {sum |err|, n}.
extern "C" int32_t mrs_fit_mae_push_async(mrs_engine* e, const mrs_ratings* train, mrs_model** inout, const mrs_ratings* test, mrs_exchange* x_items,
                                          mrs_exchange* x_pair, const int32_t* device_known_items, int32_t n_known, void* device_out2) {
  MRS_REQUIRE(e && train && inout && *inout && test && x_items && x_pair && device_known_items && device_out2 && n_known >= 0, MRS_ERR_INVALID,
              "mrs_fit_mae_push_async: NULL argument (pass a model of this train set)");
  MRS_REQUIRE(x_items != x_pair, MRS_ERR_INVALID, "mrs_fit_mae_push_async: the two exchanges need their own handles (own epoch counters)");
  MRS_REQUIRE(train->eng == e && test->eng == e, MRS_ERR_INVALID, "mrs_fit_mae_push_async: the rating sets live on another engine");
  MRS_REQUIRE(train->value_kind == kValueCode && test->value_kind == kValueCode && test->n > 0 && train->n > 0, MRS_ERR_UNSUPPORTED,
              "mrs_fit_mae_push_async: needs non-empty half-star coded train and test shards");
  mrs_model* m = *inout;
  MRS_REQUIRE(!m->want_item_avg, MRS_ERR_UNSUPPORTED, "mrs_fit_mae_push_async: switch item averages off first (mrs_model_set_item_averages)");
  use_engine(e);
  if (!m->slot_of_item || !m->item_slot || m->n_slots_known != n_known) {  // first use (never inside a graph capture)
    dev_free(m->slot_of_item); m->slot_of_item = nullptr;
    dev_free(m->item_slot); m->item_slot = nullptr;
    MRS_TRY(dev_alloc(&m->slot_of_item, (size_t)std::max(n_known, 1)));
    MRS_TRY(dev_alloc(&m->item_slot, (size_t)std::max(m->n_items, 1)));
    MRS_CUDA(cudaMemcpyAsync(m->slot_of_item, device_known_items, sizeof(int32_t) * (size_t)n_known, cudaMemcpyDeviceToDevice, e->stream));
    MRS_CUDA(cudaMemsetAsync(m->item_slot, 0xff, sizeof(int32_t) * (size_t)m->n_items, e->stream));  // -1: occurs on no rank
    if (n_known > 0) {
      item_slot_kernel<<<(n_known + 255) / 256, 256, 0, e->stream>>>(m->slot_of_item, n_known, m->item_slot);
      count_launch();
      MRS_CUDA(cudaGetLastError());
    }
    // items known on no rank are never written by the test pass: their deviation is 0.0 (P:197), their count 0
    MRS_CUDA(cudaMemsetAsync(m->idevavg, 0, sizeof(double) * (size_t)m->n_items, e->stream));
    MRS_CUDA(cudaMemsetAsync(m->xbuf, 0, sizeof(double) * (2 * (size_t)m->n_items + 2), e->stream));
    m->n_slots_known = n_known;
  }
  PushDev big, pair;
  MRS_TRY(exchange_push_dev(x_items, 2 * (int64_t)n_known + 2, &big));
  MRS_TRY(exchange_push_dev(x_pair, 2, &pair));
  // Who delivers the per-item partial sums: the test pass' CTAs in their prologue (default: three kernels per step) or the
  // fit's own push kernel, with the test pass still waiting for the flags and summing the deliveries itself
  // (MRS_CLOSURE_DELIVER=kernel: four kernels, no finishing kernel).  Measured at 2 ranks, weak step: 93.0 us (default),
  // 99.0 us (kernel delivery), 92.5 us for the five-kernel fused form (mrs_fit_local_push / mrs_fit_finish_pull / mrs_mae_push_async).
  const char* dv = getenv("MRS_CLOSURE_DELIVER");
  const bool by_test_pass = !(dv && strcmp(dv, "kernel") == 0);
  MRS_TRY(fit_local(e, train, inout, false, by_test_pass ? nullptr : &big, by_test_pass));
  MRS_TRY(launch_mae_tiled_baseline(m, test, (double*)device_out2, &pair, true, &big, by_test_pass));
  m->finished = true;  // the test pass has written the model's arrays
  m->host_valid = false;
  return MRS_OK;
}

extern "C" int32_t mrs_exchange_status(mrs_exchange* x, int32_t* timed_out) {
  MRS_REQUIRE(x && timed_out, MRS_ERR_INVALID, "mrs_exchange_status: NULL argument");
  use_engine(x->eng);
  MRS_CUDA(cudaMemcpyAsync(timed_out, x->d_error, sizeof(int32_t), cudaMemcpyDeviceToHost, x->eng->stream));
  MRS_CUDA(cudaStreamSynchronize(x->eng->stream));
  if (*timed_out) x->failed = true;  // sticky: see allreduce_impl
  return MRS_OK;
}

extern "C" int32_t mrs_exchange_set_timeout_ms(mrs_exchange* x, int64_t milliseconds) {
  MRS_REQUIRE(x && milliseconds > 0, MRS_ERR_INVALID, "mrs_exchange_set_timeout_ms: bad argument");
  x->timeout_cycles = (long long)milliseconds * 2000000LL;  // clock64 ticks at <= 2 GHz: the bound is at least `milliseconds`
  return MRS_OK;
}

extern "C" int32_t mrs_exchange_stamps(mrs_exchange* x, uint64_t* out8) {
  MRS_REQUIRE(x && out8, MRS_ERR_INVALID, "mrs_exchange_stamps: NULL argument");
  use_engine(x->eng);
  MRS_CUDA(cudaMemcpyAsync(out8, x->d_stamps, 8 * sizeof(uint64_t), cudaMemcpyDeviceToHost, x->eng->stream));
  MRS_CUDA(cudaStreamSynchronize(x->eng->stream));
  return MRS_OK;
}

extern "C" void mrs_exchange_destroy(mrs_exchange* x) {
  if (!x) return;
  if (x->eng) { cudaSetDevice(x->eng->device); cudaStreamSynchronize(x->eng->stream); }
  for (int p = 0; p < (int)x->peer_base.size(); ++p)
    if (!x->local && p != x->rank && x->peer_base[p]) cudaIpcCloseMemHandle(x->peer_base[p]);
  if (x->d_peer) cudaFree(x->d_peer);
  if (x->d_done) cudaFree(x->d_done);
  if (x->d_epoch) cudaFree(x->d_epoch);
  if (x->d_pepoch) cudaFree(x->d_pepoch);
  if (x->d_pdone) cudaFree(x->d_pdone);
  if (x->d_error) cudaFree(x->d_error);
  if (x->d_stamps) cudaFree(x->d_stamps);
  if (x->base) cudaFree(x->base);
  delete x;
}
