// exchange.cu -- the one collective of the sharded baseline pass as OUR kernel over NVLink peer memory.
//
// The reference combines per-key partial sums with Spark's reduceByKey + collect (P:267-268) and sum/count (P:247).
// On B200s of one box every rank (one process per GPU) maps the others' exchange buffers through CUDA IPC; a single
// kernel then (1) publishes this rank's partial sums in its own symmetric buffer, (2) raises a flag in every peer's
// memory (NVLink store, release at system scope), (3) waits for the peers' flags, (4) reads all partial sums with
// 128-bit peer loads and adds them in RANK ORDER, so every rank ends with bit-identical totals.  A 3.3 MB buffer is
// latency bound: NCCL's all-reduce costs 30-40 us per call here, this kernel one peer round trip plus the reads.
// Buffers are double-buffered by epoch parity, which removes the trailing barrier: a rank overwrites parity p again two
// epochs later, and its peers enter the next epoch only after they finished reading.
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
  unsigned char* base = nullptr; // own symmetric allocation: [2][n] doubles | flags[world] u64 | error word
  std::vector<void*> peer_base;  // mapped bases of all ranks (own = base)
  double** d_peer = nullptr;     // device array of the mapped bases
  unsigned long long* d_epoch = nullptr;  // number of completed exchanges (device side, so that launches can be graph-captured)
  unsigned int* d_done = nullptr;  // [2] blocks-published / blocks-finished counters of this rank
  int32_t* d_error = nullptr;
  bool connected = false;
};

namespace mrs {
namespace {

constexpr int kExThreads = 512;

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// inout[0..n): this rank's partial sums on entry, the sum over all ranks (added in rank order) on exit
__global__ void __launch_bounds__(kExThreads) peer_allreduce_kernel(double* const* __restrict__ peer, int32_t rank, int32_t world, int64_t n,
                                                                   int64_t cap, unsigned long long* __restrict__ epoch_done,
                                                                   unsigned int* __restrict__ done, int32_t* __restrict__ error,
                                                                   double* __restrict__ inout) {
  __shared__ bool ok;
  const unsigned long long epoch = *epoch_done + 1;  // stable for the whole kernel: only its last block advances it
  const int parity = (int)(epoch & 1);
  double* mine = peer[rank] + (size_t)parity * cap;
  // (1) publish
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) mine[i] = inout[i];
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    // (2) the last publishing block of this rank raises the flag in every rank's flag row (own included)
    if (atomicAdd(done, 1u) + 1u == gridDim.x) {
      done[0] = 0;
      __threadfence_system();
      for (int p = 0; p < world; ++p) {
        unsigned long long* flags = reinterpret_cast<unsigned long long*>(peer[p] + 2 * (size_t)cap);
        st_release_sys(flags + rank, epoch);
      }
    }
    // (3) wait for every rank's flag in OUR flag row
    const unsigned long long* my_flags = reinterpret_cast<const unsigned long long*>(peer[rank] + 2 * (size_t)cap);
    bool good = true;
    const long long t0 = clock64();
    for (int p = 0; p < world && good; ++p) {
      while (ld_acquire_sys(my_flags + p) < epoch) {
        if (clock64() - t0 > 4000000000LL) { good = false; atomicExch(error, 1); break; }  // ~2 s at 1.9 GHz
      }
    }
    ok = good;
  }
  __syncthreads();
  if (ok) {
    // (4) reduce in rank order (identical on every rank)
    const int64_t n2 = n >> 1;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n2; i += (int64_t)gridDim.x * blockDim.x) {
      double2 acc = make_double2(0.0, 0.0);
      for (int p = 0; p < world; ++p) {
        const double2 v = reinterpret_cast<const double2*>(peer[p] + (size_t)parity * cap)[i];
        acc.x += v.x;
        acc.y += v.y;
      }
      reinterpret_cast<double2*>(inout)[i] = acc;
    }
    if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
      double a = 0.0;
      for (int p = 0; p < world; ++p) a += (peer[p] + (size_t)parity * cap)[n - 1];
      inout[n - 1] = a;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0 && atomicAdd(done + 1, 1u) + 1u == gridDim.x) {  // last block out: this exchange is complete
    done[1] = 0;
    *epoch_done = epoch;
  }
}

}  // namespace
}  // namespace mrs

using namespace mrs;

// ---- C ABI (declared in include/mrs_b200.h)
extern "C" int32_t mrs_exchange_create(mrs_engine* e, int64_t n_doubles, int32_t rank, int32_t world, void* ipc_handle_out64,
                                       mrs_exchange** out) {
  MRS_REQUIRE(e && out && ipc_handle_out64 && n_doubles > 0 && world >= 1 && rank >= 0 && rank < world, MRS_ERR_INVALID,
              "mrs_exchange_create: bad argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  use_engine(e);
  mrs_exchange* x = new mrs_exchange();
  x->eng = e; x->rank = rank; x->world = world;
  x->n = (n_doubles + 1) & ~(int64_t)1;
  const size_t bytes = 2 * (size_t)x->n * sizeof(double) + (size_t)world * sizeof(unsigned long long) + 64;
  // IPC-shareable memory must come from cudaMalloc directly (not from the engine's block cache)
  cudaError_t ce = cudaMalloc((void**)&x->base, bytes);
  if (ce != cudaSuccess) { delete x; set_error("mrs_exchange_create: cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(ce)); return MRS_ERR_NOMEM; }
  MRS_CUDA(cudaMemset(x->base, 0, bytes));
  MRS_CUDA(cudaMalloc((void**)&x->d_done, 2 * sizeof(unsigned int)));
  MRS_CUDA(cudaMemset(x->d_done, 0, 2 * sizeof(unsigned int)));
  MRS_CUDA(cudaMalloc((void**)&x->d_epoch, sizeof(unsigned long long)));
  MRS_CUDA(cudaMemset(x->d_epoch, 0, sizeof(unsigned long long)));
  MRS_CUDA(cudaMalloc((void**)&x->d_error, sizeof(int32_t)));
  MRS_CUDA(cudaMemset(x->d_error, 0, sizeof(int32_t)));
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

extern "C" int32_t mrs_exchange_allreduce_async(mrs_exchange* x, void* device_inout, int64_t n_doubles) {
  MRS_REQUIRE(x && device_inout, MRS_ERR_INVALID, "mrs_exchange_allreduce_async: NULL argument");
  MRS_REQUIRE(x->connected, MRS_ERR_INVALID, "mrs_exchange_allreduce_async: call mrs_exchange_connect first");
  MRS_REQUIRE(n_doubles > 0 && n_doubles <= x->n, MRS_ERR_INVALID, "mrs_exchange_allreduce_async: %lld doubles exceed the capacity %lld",
              (long long)n_doubles, (long long)x->n);
  MRS_REQUIRE(((uintptr_t)device_inout & 15) == 0, MRS_ERR_INVALID, "mrs_exchange_allreduce_async: buffer must be 16-byte aligned");
  use_engine(x->eng);
  // every block waits on the flags: the grid must be co-resident (one CTA per SM at most)
  const int64_t work = (n_doubles + 1) / 2;
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((work + kExThreads - 1) / kExThreads, (int64_t)x->eng->sm_count));
  peer_allreduce_kernel<<<grid, kExThreads, 0, x->eng->stream>>>(x->d_peer, x->rank, x->world, n_doubles, x->n, x->d_epoch, x->d_done, x->d_error,
                                                                 (double*)device_inout);
  mark(x->eng, "peer_allreduce");
  MRS_CUDA(cudaGetLastError());
  return MRS_OK;
}

extern "C" int32_t mrs_exchange_status(mrs_exchange* x, int32_t* timed_out) {
  MRS_REQUIRE(x && timed_out, MRS_ERR_INVALID, "mrs_exchange_status: NULL argument");
  use_engine(x->eng);
  MRS_CUDA(cudaMemcpyAsync(timed_out, x->d_error, sizeof(int32_t), cudaMemcpyDeviceToHost, x->eng->stream));
  MRS_CUDA(cudaStreamSynchronize(x->eng->stream));
  return MRS_OK;
}

extern "C" void mrs_exchange_destroy(mrs_exchange* x) {
  if (!x) return;
  if (x->eng) { cudaSetDevice(x->eng->device); cudaStreamSynchronize(x->eng->stream); }
  for (int p = 0; p < (int)x->peer_base.size(); ++p)
    if (p != x->rank && x->peer_base[p]) cudaIpcCloseMemHandle(x->peer_base[p]);
  if (x->d_peer) cudaFree(x->d_peer);
  if (x->d_done) cudaFree(x->d_done);
  if (x->d_epoch) cudaFree(x->d_epoch);
  if (x->d_error) cudaFree(x->d_error);
  if (x->base) cudaFree(x->base);
  delete x;
}
