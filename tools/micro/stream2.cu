// stream2.cu -- how fast can one launch read 80 MB that are cold in a CLEAN L2?  (calibration for the item pass; not product code)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o stream2 stream2.cu && ./stream2
// Variants: plain 16-byte loads (persistent grid); per-warp TMA rings (cp.async.bulk + mbarrier) with contiguous per-warp ranges
// (what the item pass does) or chunks dealt out cyclically over all warps of the grid; stage sizes 1/2/4 KB; ring depth 2..8.
// Every kernel stamps first-CTA-start / last-CTA-end with %globaltimer, so launch overhead is separated from streaming time.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__device__ __forceinline__ unsigned long long gt() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ void mbar_init(uint64_t* b, int c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(b)), "r"(c)); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t ph) {
  asm volatile("{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D;\nbra W;\nD:\n}" ::"r"((uint32_t)__cvta_generic_to_shared(b)), "r"(ph) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* b) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src),
               "r"(bytes), "r"((uint32_t)__cvta_generic_to_shared(b)) : "memory");
}

__global__ void sum16(const uint4* __restrict__ p, size_t n4, unsigned long long* out, unsigned long long* tl) {
  if (threadIdx.x == 0) atomicMin(tl, gt());
  unsigned long long acc = 0;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    uint4 v = p[i];
    acc += (unsigned long long)v.x + v.y + v.z + v.w;
  }
  acc += __shfl_xor_sync(~0u, acc, 16); acc += __shfl_xor_sync(~0u, acc, 8); acc += __shfl_xor_sync(~0u, acc, 4);
  acc += __shfl_xor_sync(~0u, acc, 2); acc += __shfl_xor_sync(~0u, acc, 1);
  if ((threadIdx.x & 31) == 0) atomicAdd(out, acc);
  __syncthreads();
  if (threadIdx.x == 0) atomicMax(tl + 1, gt());
}

// per-warp ring: STAGES x (ROWS*128 B).  CYCLIC=0: warp w streams one contiguous range; CYCLIC=1: chunk c belongs to global warp c % W
template <int ROWS, int STAGES, int CYCLIC, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1) ring(const uint32_t* __restrict__ p, int64_t n_rows, unsigned long long* out, unsigned long long* tl) {
  extern __shared__ __align__(128) unsigned char smem[];
  uint32_t* s_ring = reinterpret_cast<uint32_t*>(smem);
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(smem + (size_t)WARPS * STAGES * ROWS * 128);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (threadIdx.x == 0) atomicMin(tl, gt());
  uint64_t* bar = s_bar + wid * STAGES;
  uint32_t* rg = s_ring + (size_t)wid * STAGES * ROWS * 32;
  if (lane == 0) {
    for (int s = 0; s < STAGES; ++s) mbar_init(bar + s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  const int64_t W = (int64_t)gridDim.x * WARPS, w = (int64_t)blockIdx.x * WARPS + wid;
  const int64_t n_chunks_all = (n_rows + ROWS - 1) / ROWS;
  int64_t c0, nc, stride;
  if (CYCLIC) { c0 = w; stride = W; nc = (n_chunks_all - w + W - 1) / W; if (nc < 0) nc = 0; }
  else { const int64_t a = n_chunks_all * w / W, b = n_chunks_all * (w + 1) / W; c0 = a; stride = 1; nc = b - a; }
  if (lane == 0)
    for (int s = 0; s < STAGES; ++s)
      if (s < nc) {
        const int64_t r = (c0 + s * stride) * ROWS;
        const uint32_t bytes = (uint32_t)min((int64_t)ROWS, n_rows - r) * 128u;
        mbar_expect(bar + s, bytes);
        bulk_g2s(rg + s * ROWS * 32, p + (r << 5), bytes, bar + s);
      }
  unsigned long long acc = 0;
  for (int64_t c = 0; c < nc; ++c) {
    const int s = (int)(c % STAGES);
    const int64_t r = (c0 + c * stride) * ROWS;
    const int nrows = (int)min((int64_t)ROWS, n_rows - r);
    mbar_wait(bar + s, (uint32_t)(c / STAGES) & 1u);
    const uint32_t* rp = rg + s * ROWS * 32 + lane;
    uint32_t v[ROWS];
#pragma unroll
    for (int k = 0; k < ROWS; ++k) v[k] = (k < nrows) ? rp[k * 32] : 0u;
    __syncwarp();
    if (lane == 0 && c + STAGES < nc) {
      const int64_t rr = (c0 + (c + STAGES) * stride) * ROWS;
      const uint32_t bytes = (uint32_t)min((int64_t)ROWS, n_rows - rr) * 128u;
      mbar_expect(bar + s, bytes);
      bulk_g2s(rg + s * ROWS * 32, p + (rr << 5), bytes, bar + s);
    }
#pragma unroll
    for (int k = 0; k < ROWS; ++k) acc += v[k];
  }
  acc += __shfl_xor_sync(~0u, acc, 16); acc += __shfl_xor_sync(~0u, acc, 8); acc += __shfl_xor_sync(~0u, acc, 4);
  acc += __shfl_xor_sync(~0u, acc, 2); acc += __shfl_xor_sync(~0u, acc, 1);
  if (lane == 0) atomicAdd(out, acc);
  __syncthreads();
  if (threadIdx.x == 0) atomicMax(tl + 1, gt());
}

static void* g_flush_w; static void* g_flush_r; static unsigned long long* g_sink;
__global__ void read_all(const uint4* p, size_t n4, unsigned long long* sink) {
  unsigned long long acc = 0;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) { uint4 v = p[i]; acc += v.x ^ v.y ^ v.z ^ v.w; }
  if (acc == 0x1234567ull) *sink = acc;
}
static void flush_clean() {  // 256 MiB write, then 256 MiB read of another buffer: cold AND clean L2
  cudaMemsetAsync(g_flush_w, 1, 256u << 20);
  read_all<<<1184, 256>>>((const uint4*)g_flush_r, (256u << 20) / 16, g_sink);
}

template <typename F>
static void time_it(const char* name, F f, unsigned long long* d_tl, double mb) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  float best = 1e9f; double best_in = 1e18;
  for (int r = 0; r < 7; ++r) {
    unsigned long long init[2] = {~0ull, 0ull};
    cudaMemcpy(d_tl, init, sizeof(init), cudaMemcpyHostToDevice);
    flush_clean();
    cudaDeviceSynchronize();
    cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
    unsigned long long h[2]; cudaMemcpy(h, d_tl, sizeof(h), cudaMemcpyDeviceToHost);
    const double in = (double)(h[1] - h[0]) / 1e3; if (in < best_in) best_in = in;
  }
  printf("%-44s events %6.1f us | in-kernel %6.1f us (%5.0f GB/s)\n", name, best * 1e3, best_in, mb / best_in * 1e3);
  cudaError_t e = cudaGetLastError(); if (e != cudaSuccess) printf("  error: %s\n", cudaGetErrorString(e));
}

template <int ROWS, int STAGES, int CYCLIC, int WARPS>
static void run_ring(const char* name, const uint32_t* d, int64_t n_rows, unsigned long long* o, unsigned long long* tl, int grid, double mb) {
  const size_t smem = (size_t)WARPS * STAGES * ROWS * 128 + (size_t)WARPS * STAGES * 8;
  cudaFuncSetAttribute(ring<ROWS, STAGES, CYCLIC, WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  time_it(name, [&] { ring<ROWS, STAGES, CYCLIC, WARPS><<<grid, WARPS * 32, smem>>>(d, n_rows, o, tl); }, tl, mb);
}

int main() {
  const size_t n = 20u * 1000 * 1000;  // 80 MB of uint32
  const int64_t n_rows = n / 32;
  uint32_t* d; unsigned long long* o; unsigned long long* tl;
  CK(cudaMalloc(&d, n * 4)); CK(cudaMalloc(&o, 8)); CK(cudaMalloc(&tl, 16)); CK(cudaMalloc(&g_flush_w, 256u << 20)); CK(cudaMalloc(&g_flush_r, 256u << 20));
  CK(cudaMalloc(&g_sink, 8));
  CK(cudaMemset(d, 1, n * 4)); CK(cudaMemset(g_flush_r, 0, 256u << 20));
  const double mb = n * 4.0 / 1e6;
  int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  printf("80 MB cold read, clean L2, %d SMs.  in-kernel = last CTA end - first CTA start (%%globaltimer)\n", sms);
  for (int g : {2, 4, 8}) {
    char nm[64]; snprintf(nm, sizeof nm, "sum16 persistent grid %dx256", sms * g);
    time_it(nm, [&] { sum16<<<sms * g, 256>>>((const uint4*)d, n / 4, o, tl); }, tl, mb);
  }
  time_it("sum16 grid 148x1024", [&] { sum16<<<sms, 1024>>>((const uint4*)d, n / 4, o, tl); }, tl, mb);
  time_it("sum16 grid 296x1024", [&] { sum16<<<sms * 2, 1024>>>((const uint4*)d, n / 4, o, tl); }, tl, mb);
  run_ring<8, 4, 0, 32>("ring/warp 1KBx4 contiguous (item pass)", d, n_rows, o, tl, sms, mb);
  run_ring<8, 4, 1, 32>("ring/warp 1KBx4 cyclic", d, n_rows, o, tl, sms, mb);
  run_ring<8, 6, 0, 32>("ring/warp 1KBx6 contiguous", d, n_rows, o, tl, sms, mb);
  run_ring<8, 6, 1, 32>("ring/warp 1KBx6 cyclic", d, n_rows, o, tl, sms, mb);
  run_ring<16, 3, 0, 32>("ring/warp 2KBx3 contiguous", d, n_rows, o, tl, sms, mb);
  run_ring<16, 3, 1, 32>("ring/warp 2KBx3 cyclic", d, n_rows, o, tl, sms, mb);
  run_ring<32, 2, 0, 32>("ring/warp 4KBx2 contiguous", d, n_rows, o, tl, sms, mb);
  run_ring<32, 2, 1, 32>("ring/warp 4KBx2 cyclic", d, n_rows, o, tl, sms, mb);
  run_ring<8, 2, 0, 32>("ring/warp 1KBx2 contiguous", d, n_rows, o, tl, sms, mb);
  run_ring<8, 2, 1, 32>("ring/warp 1KBx2 cyclic", d, n_rows, o, tl, sms, mb);
  run_ring<8, 4, 0, 16>("ring/warp 1KBx4 contiguous, 16 warps", d, n_rows, o, tl, sms, mb);
  run_ring<8, 8, 1, 16>("ring/warp 1KBx8 cyclic, 16 warps", d, n_rows, o, tl, sms, mb);
  run_ring<32, 4, 1, 8>("ring/warp 4KBx4 cyclic, 8 warps", d, n_rows, o, tl, sms, mb);
  run_ring<32, 4, 0, 8>("ring/warp 4KBx4 contiguous, 8 warps", d, n_rows, o, tl, sms, mb);
  return 0;
}
